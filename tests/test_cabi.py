"""CPU suite: the C-ABI library loads and exports every symbol include/essentials_b200.h declares, the
binding table matches the header, and the product refuses to run without a device (no CPU fallback)."""
import os
import re

import pytest

import essentials_b200 as ess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "essentials_b200.h")).read()
    return sorted(set(re.findall(r"ESS_API\s+[\w\s\*]+?\b(ess_\w+)\s*\(", text)))


def test_header_declares_expected_surface():
    names = _declared()
    for must in ("ess_bfs", "ess_sssp", "ess_pagerank", "ess_ppr", "ess_kcore", "ess_color", "ess_advance_probe",
                 "ess_filter_probe", "ess_graph_create", "ess_context_create", "ess_bfs_partition_step"):
        assert must in names
    assert len(names) >= 20


def test_library_exports_every_declared_symbol():
    L = ess.lib()
    for name in _declared():
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert sorted(ess._SIGNATURES) == _declared(), "ctypes table and header disagree"
    assert L.ess_version() >= 100


def test_enum_values_match_reference_order():
    text = open(os.path.join(ROOT, "include", "gunrock", "framework", "operators", "configs.hxx")).read()
    lb = re.search(r"enum load_balance_t \{([^}]*)\}", text).group(1).replace(" ", "").split(",")
    assert lb == list(ess.LOAD_BALANCE)
    fa = re.search(r"enum filter_algorithm_t \{([^}]*)\}", text).group(1).replace(" ", "").split(",")
    assert fa == list(ess.FILTER)
    di = re.search(r"enum advance_direction_t \{([^}]*)\}", text).group(1).replace(" ", "").split(",")
    assert di == list(ess.DIRECTION)


def test_no_cpu_fallback():
    import torch
    from essentials_b200 import graphgen
    g = graphgen.rmat_csr(6)
    with pytest.raises(ess.EssentialsError):
        ess.Graph(g)  # host tensors are rejected: there is no CPU path
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            ess.Context(0)


def test_product_never_imports_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "essentials_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".hxx", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "liboracle" not in text and "libref_cpu" not in text, f
    for base, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            assert "oracle" not in open(os.path.join(base, f)).read().replace("the oracle", "").replace(
                "Jacobi oracle", "").replace("/oracle stream", ""), f


def test_tune_knobs_are_validated_on_the_host():
    """ess_tune is host-only state: unknown knobs and a non-positive peer timeout are refused with a message."""
    ess.tune("dist_peer_timeout_ms", 2500)
    ess.tune("dist_peer_timeout_ms", 4000)
    with pytest.raises(ess.EssentialsError, match="positive"):
        ess.tune("dist_peer_timeout_ms", 0)
    with pytest.raises(ess.EssentialsError, match="unknown knob"):
        ess.tune("no_such_knob", 1)


def test_graph_from_host_validates_arrays_before_touching_the_device():
    """Graph.from_host refuses non-contiguous arrays and wrong element types on the host side (no context needed)."""
    import torch
    from essentials_b200 import graphgen
    g = graphgen.rmat_csr(6)
    wrong_ids = graphgen.CSR(g.n, g.m, g.offsets, g.indices.long(), None, "", True)
    with pytest.raises(ess.EssentialsError, match="int32"):
        ess.Graph.from_host(None, wrong_ids)
    strided = graphgen.CSR(g.n, g.m, g.offsets, torch.stack([g.indices, g.indices], 1)[:, 0], None, "", True)
    with pytest.raises(ess.EssentialsError, match="contiguous"):
        ess.Graph.from_host(None, strided)
    wrong_w = graphgen.CSR(g.n, g.m, g.offsets, g.indices, torch.ones(g.m, dtype=torch.float64), "", True)
    with pytest.raises(ess.EssentialsError, match="float32"):
        ess.Graph.from_host(None, wrong_w)
