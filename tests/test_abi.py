"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/nimpress_cuda.h declares; without a GPU it refuses to work (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    import nimpress_b200 as nb
    return nb.load_library()


def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(np[ch]_[a-z_0-9]+)\s*\(", txt)))


def test_exports_match_header(lib):
    syms = declared_symbols("nimpress_cuda.h")
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/nimpress_cuda.h but not exported"
    assert sorted(lib._npc_symbols) == syms, "python binding and header disagree"


def test_struct_layouts_match_oracle_and_header():
    import nimpress_b200 as nb
    import orc
    assert nb.ROW_DTYPE == orc.ROW_DTYPE and nb.LOCUS_DTYPE == orc.LOCUS_DTYPE
    assert nb.ROW_DTYPE.fields["beta"][1] == 8 and nb.ROW_DTYPE.fields["kind"][1] == 28
    assert nb.LOCUS_DTYPE.fields["imputed"][1] == 40


def test_no_cpu_fallback(lib):
    import torch
    import nimpress_b200 as nb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nb.NpcError, match="no CPU fallback|no CUDA"):
        nb.Engine(16)


def test_normalise_is_reference_epilogue(lib):
    """npc_normalise == scores /= nloci*2 ; scores += offset (src/nimpress.nim:643-649)."""
    import nimpress_b200 as nb
    x = np.array([0.3, -1.25, 0.0, np.nan, 7e-3])
    out = x.copy()
    lib.npc_normalise(out.ctypes.data, len(out), 6, 0.123)
    want = x / (6 * 2.0) + 0.123
    assert np.array_equal(np.isnan(out), np.isnan(want))
    assert np.array_equal(out[~np.isnan(out)], want[~np.isnan(want)])
    out = np.array([1.0]); lib.npc_normalise(out.ctypes.data, 1, 0, 0.0)
    assert np.isinf(out[0])
    out = np.array([0.0]); lib.npc_normalise(out.ctypes.data, 1, 0, 0.0)
    assert np.isnan(out[0])          # nloci == 0: 0/0
