"""CPU tests of the boundary: the C-ABI library builds, loads, exports every symbol that
include/aligntools_b200.h declares, refuses to compute without a GPU (no CPU fallback), and its
pure-host helpers work.  No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def A():
    import aligntools.c_b200 as A
    A.build()
    return A


def test_header_symbols_exported(A):
    hdr = open(os.path.join(ROOT, "include", "aligntools_b200.h")).read()
    declared = set(re.findall(r"\b(at_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"at_pack_2bit()"}
    assert declared, "no declarations parsed"
    lib = A.load_library()
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    assert set(A.EXPORTS) <= declared | set(A.EXPORTS)
    for sym in A.EXPORTS:
        assert sym in declared, f"{sym} bound in Python but not declared in the header"


def test_no_cpu_fallback(A):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(A.AtError) as e:
        A.Aligner()
    assert e.value.rc == -2            # AT_E_CUDA


def test_product_does_not_import_oracle(A):
    pkg = os.path.join(ROOT, "aligntools")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                src = open(os.path.join(dp, fn), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, fn


def test_pack_2bit_and_cigar_helpers(A):
    lib = A.load_library()
    seq = b"ACGTTGCAAC"
    dst = (C.c_uint8 * 8)()
    n = lib.at_pack_2bit(seq, len(seq), dst)
    assert n == 3
    codes = [(dst[k >> 2] >> (2 * (k & 3))) & 3 for k in range(len(seq))]
    assert bytes(b"ACGT"[c] for c in codes) == seq
    assert lib.at_pack_2bit(b"ACGN", 4, (C.c_uint8 * 8)()) < 0
    # numpy packer agrees with the C one
    buf, off, lens = A.pack_seqs([seq, b"TTGA"])
    p2, poff, _ = A.pack_2bit(buf, off, lens)
    assert bytes(p2[:3]) == bytes(dst[:3]) and int(poff[1]) == 3
    ops = np.array([(5 << 4) | 0, (2 << 4) | 1, (7 << 4) | 3, (1 << 4) | 2], dtype=np.uint32)
    out = C.create_string_buffer(64)
    assert lib.at_cigar_to_string(ops.ctypes.data, 4, out, 64) == 8
    assert out.value == b"5M2I7N1D"
    assert lib.at_cigar_to_string(ops.ctypes.data, 4, out, 4) < 0


def test_default_params_are_init_opt(A):
    lib = A.load_library()
    p = A._Params()
    lib.at_default_params(C.byref(p))
    assert (p.m, p.u, p.o, p.e, p.j, p.jump) == (1, -2, -5, -1, -10, 0)   # src/alignment.h:105-110
    o = A.Opt()
    assert (o.m, o.u, o.o, o.e, o.j, o.jump) == (1, -2, -5, -1, -10, False)


def test_synth_workloads_shapes():
    from aligntools.c_b200 import synth
    w = synth.config2_local(n_pairs=256)
    assert w["q"].size == 256 * 150 and w["t"].size == 256 * 500
    assert set(np.unique(w["q"])) <= set(b"ACGT")
    w2 = synth.config2_local(n_pairs=256)
    assert np.array_equal(w["q"], w2["q"])            # deterministic
    w3 = synth.config3_fit_jump(n_pairs=2)
    assert w3["q_len"].tolist() == [2000, 2000] and w3["t_len"].tolist() == [20000, 20000]
    assert w3["site_off"][-1] >= 16
    w4 = synth.config4_overlap(n_pairs=2, lo=500, hi=900)
    assert all(500 <= x <= 900 for x in w4["q_len"])
    w5 = synth.config5_edit(n_pairs=1, length=3000)
    assert w5["q_len"][0] == 3000 and w5["t_len"][0] == 3000
