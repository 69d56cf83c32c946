"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the host mirror has the reference's parameter inventory, and the product never
touches the oracle or computes on the CPU."""
import ctypes as C
import os
import re

import pytest
import torch

from tests.util import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "giga_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(giga_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import giga_b200
    from giga_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 14
    raw = C.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(raw, s), f"{s} declared in include/giga_b200.h but not exported"
        assert s in _lib.SYMBOLS, f"{s} has no ctypes prototype"
    assert sorted(_lib.SYMBOLS) == declared
    assert raw.giga_version() == 100
    assert os.path.dirname(_lib.LIB_PATH) == os.path.join(ROOT, "giga_b200")  # in-tree, not site-packages


def test_state_dict_matches_reference_inventory(golden):
    import giga_b200

    for name in ["giga", "giga_aff", "giga_geo", "giga_detach"]:
        net = giga_b200.get_network(name)
        sd = net.state_dict()
        assert list(sd.keys()) == list(golden[f"keys_{name}"])
        assert [str(tuple(v.shape)) for v in sd.values()] == list(golden[f"shapes_{name}"])
    assert sum(p.numel() for p in giga_b200.get_network("giga").parameters()) == 581863
    with pytest.raises(NotImplementedError):
        giga_b200.get_network("vgn")
    with pytest.raises(KeyError):
        giga_b200.get_network("nope")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import giga_b200
    from giga_b200._lib import lib

    net = giga_b200.get_network("giga")
    with pytest.raises(giga_b200.GigaError):
        net.encode_inputs(torch.zeros(1, 40, 40, 40))
    with pytest.raises(giga_b200.GigaError):
        net(torch.zeros(1, 40, 40, 40), torch.zeros(1, 4, 3))
    h = C.c_void_p()
    assert lib.giga_ctx_create(C.byref(h), 0) == -2  # GIGA_ENODEV
    assert b"no CPU path" in lib.giga_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "giga_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|oracle/|oracle\.)", txt), f"{f} uses the oracle"
    for f in ["bench.py", "__graft_entry__.py"]:
        p = os.path.join(ROOT, f)
        if os.path.exists(p):
            assert "/root/reference" not in open(p).read(), f"{f} must not read /root/reference at run time"
