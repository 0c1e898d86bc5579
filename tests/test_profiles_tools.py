"""CPU suite: the scripts that turn ncu output into the committed summaries keep working on the committed data."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ncu_summarise_reproduces_committed_traffic(tmp_path):
    csv = os.path.join(ROOT, "profiles", "r02e_launches_bfs_kron26.csv")
    out_json = str(tmp_path / "traffic.json")
    out_txt = str(tmp_path / "shares.txt")
    run = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summarise.py"), csv, "--shares", out_txt,
                          "--traffic", out_json], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stderr[-1000:]
    got = json.load(open(out_json))
    committed = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    for cls in ("pull_step", "push_expand"):
        assert abs(got[cls]["bytes_per_launch"] - committed[cls]["bytes_per_launch"]) < 1.0, cls
        assert got[cls]["launches"] == committed[cls]["launches"]
    table = open(out_txt).read()
    assert "pull_step" in table and "push_expand" in table and "counters" in table


def test_bench_reads_the_committed_traffic():
    sys.path.insert(0, ROOT)
    import bench
    t = bench.ncu_traffic("pull_step")
    assert t and t > 1e6
