"""GPU suite, multi-GPU: partitioned BFS over NCCL equals the single-GPU result (needs >= 2 GPUs on the box)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (CPU gloo test covers the host logic)")
@pytest.mark.parametrize("loop", ["native", "native-nccl", "python"])
def test_partitioned_bfs_two_gpus(loop):
    """The drivers of the 1-D partitioned BFS — ess_dist_bfs with its peer-memory exchange kernels, the same loop
    over NCCL collectives, and the torch.distributed loop that the gloo tests cover — must reproduce the
    single-GPU depths."""
    port = {"native": "29517", "native-nccl": "29516", "python": "29518"}[loop]
    extra = {"native": [], "native-nccl": ["--nccl-exchange"], "python": ["--python-loop"]}[loop]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", port, os.path.join(ROOT, "scripts", "dist_check.py"), "--scale", "18"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    if loop.startswith("native"):
        assert ("exchange: nccl" if loop == "native-nccl" else "exchange: peer-memory") in out.stdout, out.stdout[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (CPU gloo test covers the host logic)")
@pytest.mark.parametrize("loop", ["native", "native-nccl", "python"])
def test_partitioned_sssp_two_gpus(loop):
    """ess_dist_sssp with the fused peer-memory reduce+collect, the same loop over ncclReduceScatter(min), and
    PartitionedSSSP (torch.distributed) against the single-GPU ess_sssp distances, bit for bit."""
    port = {"native": "29519", "native-nccl": "29521", "python": "29520"}[loop]
    extra = {"native": [], "native-nccl": ["--nccl-exchange"], "python": ["--python-loop"]}[loop]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", port, os.path.join(ROOT, "scripts", "dist_check.py"), "--scale", "17", "--alg",
           "sssp"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    if loop.startswith("native"):
        assert ("exchange=nccl" if loop == "native-nccl" else "exchange=peer-memory") in out.stdout, out.stdout[-2000:]


@pytest.mark.parametrize("ranks", [1, 2])
def test_partitioned_enactor_contract(ranks):
    """gunrock::bfs::run and gunrock::sssp::run with a partitioned gcuda::multi_context_t: the unchanged enactor loop
    (prepare_frontier -> while(!is_converged) loop()), advance::execute<lb> on the owned rows and
    operators::exchange::execute routing each level's frontier to the owners, must give the single-GPU depths and
    distances bit for bit for three balancers. One rank exercises the whole path (binning, records, absorb, global
    convergence) on a single-GPU box; two ranks add the NCCL all-to-all."""
    if torch.cuda.device_count() < ranks:
        pytest.skip("needs %d GPUs" % ranks)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ranks), "--master-addr",
           "127.0.0.1", "--master-port", str(29530 + ranks), os.path.join(ROOT, "scripts", "dist_check.py"), "--scale",
           "15", "--alg", "bfs,sssp", "--enactor", "--sources", "2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST_CHECK_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("enactor bfs") == 9 and out.stdout.count("enactor sssp") == 9
    assert "equal=False" not in out.stdout
