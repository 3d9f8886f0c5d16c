"""The fill kernels' cell arithmetic (aligntools/c_b200/csrc/at_cell.cuh: tagged values, pointer nibbles as P - Q + bias),
compiled for the host by g++ and compared with the oracle on random cases of every affine mode -- int32 lanes and packed
s16x2 lanes, parameters with flipped signs included.  Pins the recurrence / tie rules / pointer algebra without a GPU."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libcell_model.so")
SRC = os.path.join(HERE, "cell_model.cpp")
HDR = os.path.join(HERE, "..", "aligntools", "c_b200", "csrc", "at_cell.cuh")
MODES = {"global": 0, "local": 1, "fit": 2, "overlap": 3}


@pytest.fixture(scope="module")
def model():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-o", SO, SRC], check=True)
    lib = C.CDLL(SO)
    lib.cell_model_run.restype = C.c_int
    lib.cell_model_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_char_p, C.c_char_p]
    return lib


def run_model(lib, mode, jump, packed, s1a, s1b, s2a, s2b, prm, sites):
    out = (C.c_int * 14)()
    oa = C.create_string_buffer(len(s1a) + len(s2a) + 1)
    ob = C.create_string_buffer(len(s1b) + len(s2b) + 1)
    sa = np.asarray(sites or [0], dtype=np.int32)
    rc = lib.cell_model_run(MODES[mode], int(jump), int(packed), s1a, len(s1a), s1b, len(s1b), s2a, s2b, len(s2a),
                            prm["m"], prm["u"], prm["o"], prm["e"], prm["j"], sa.ctypes.data, len(sites or []), out, oa, ob)
    assert rc == 0
    res = []
    for h in range(2 if packed else 1):
        r = out[7 * h:7 * h + 7]
        res.append(dict(score=r[0], end=(r[1], r[2]), beg=(r[4], r[5]), ops=(oa, ob)[h].raw[:r[6]]))
    return res


def rand_pair(rng, l1_max, l2_max, fit, alphabet=b"ACGT", l2=None):
    l2 = l2 or rng.randint(2, l2_max)
    l1 = rng.randint(1, min(l1_max, l2) if fit else l1_max)
    s2 = bytes(rng.choice(alphabet) for _ in range(l2))
    st = rng.randrange(0, max(1, l2 - l1 + 1))
    base = s2[st:st + l1]
    base = base + bytes(rng.choice(alphabet) for _ in range(l1 - len(base)))
    s1 = bytearray()
    for ch in base:
        r = rng.random()
        if r < 0.08:
            s1.append(rng.choice(alphabet))
        elif r < 0.11:
            continue
        elif r < 0.14:
            s1.append(ch); s1.append(rng.choice(alphabet))
        else:
            s1.append(ch)
    s1 = bytes(s1[:l1]) or bytes([rng.choice(alphabet)])
    return s1, s2


def rand_params(rng, flipped):
    if flipped:
        return dict(m=rng.randint(-2, 5), u=rng.randint(-5, 2), o=rng.randint(-8, 2), e=rng.randint(-4, 2), j=rng.randint(-12, 2))
    return dict(m=rng.randint(1, 5), u=rng.randint(-5, 0), o=rng.randint(-8, 0), e=rng.randint(-4, 0), j=rng.randint(-12, 0))


@pytest.mark.parametrize("mode", ["global", "local", "fit", "fitjump"])
def test_int32_lanes_vs_oracle(model, oracle_mod, mode):
    rng = random.Random(1000 + len(mode))
    md = "fit" if mode == "fitjump" else mode
    for k in range(700):
        s1, s2 = rand_pair(rng, 60, 140, md == "fit", alphabet=b"ACGT" if k % 3 else b"AC")
        prm = rand_params(rng, flipped=(k % 4 == 3))
        sites = sorted(rng.randrange(len(s2)) for _ in range(rng.choice([0, 1, 3, 8]))) if mode == "fitjump" else None
        p = oracle_mod.Params(prm["m"], prm["u"], prm["o"], prm["e"], prm["j"], mode == "fitjump")
        ref = oracle_mod.port_align(md, s1, s2, p, sites)
        got = run_model(model, md, mode == "fitjump", False, s1, s1, s2, s2, prm, sites)[0]
        assert got["score"] == ref.score, (mode, k, prm, s1, s2, sites)
        assert got["end"] == tuple(ref.coords[:2]) and got["beg"] == tuple(ref.coords[2:]), (mode, k, prm, s1, s2, sites)
        assert got["ops"] == ref.ops, (mode, k, prm, s1, s2, sites)


@pytest.mark.parametrize("mode", ["local", "global", "fit"])
def test_packed_lanes_vs_oracle(model, oracle_mod, mode):
    """Two pairs in the halves of every register (shared l2, different reads / read lengths); global and fit carry
    -inf as AT_NEG16 inside the 16 bits."""
    rng = random.Random(77 + len(mode))
    for k in range(700):
        l2 = rng.randint(2, 160)
        s1a, s2a = rand_pair(rng, 70, 160, mode == "fit", l2=l2)
        s1b, s2b = rand_pair(rng, 70, 160, mode == "fit", l2=l2)
        prm = rand_params(rng, flipped=(k % 4 == 3))
        if 8 * (max(len(s1a), len(s1b)) + l2 + 2) * max(abs(v) for v in prm.values()) >= 24000:
            continue
        p = oracle_mod.Params(prm["m"], prm["u"], prm["o"], prm["e"], prm["j"], False)
        got = run_model(model, mode, False, True, s1a, s1b, s2a, s2b, prm, None)
        for h, (s1, s2) in enumerate(((s1a, s2a), (s1b, s2b))):
            ref = oracle_mod.port_align(mode, s1, s2, p)
            assert got[h]["score"] == ref.score, (k, h, prm, s1, s2)
            assert got[h]["end"] == tuple(ref.coords[:2]) and got[h]["beg"] == tuple(ref.coords[2:]), (k, h, prm)
            assert got[h]["ops"] == ref.ops, (k, h, prm, s1, s2)


def test_overlap_cell_vs_oracle(model, oracle_mod):
    """Single-plane cell (lin_update): tags LEFT 2 / DIAGONAL 1 / RIGHT 0, 2-bit pointers as 0xAAAAAAAA - x."""
    rng = random.Random(5)
    for k in range(900):
        l1, l2 = rng.randint(1, 90), rng.randint(1, 90)
        s1 = bytes(rng.choice(b"ACGT") for _ in range(l1))
        ov = rng.randint(0, min(l1, l2))
        s2 = bytes(c if rng.random() > 0.1 else rng.choice(b"ACGT") for c in s1[l1 - ov:]) + bytes(rng.choice(b"ACGT") for _ in range(l2 - ov))
        prm = rand_params(rng, flipped=(k % 4 == 3))
        p = oracle_mod.Params(prm["m"], prm["u"], prm["o"], prm["e"], prm["j"], False)
        ref = oracle_mod.port_align("overlap", s1, s2, p)
        got = run_model(model, "overlap", False, False, s1, s1, s2, s2, prm, None)[0]
        assert got["score"] == ref.score, (k, prm, s1, s2)
        assert got["end"] == tuple(ref.coords[:2]) and got["beg"] == tuple(ref.coords[2:]), (k, prm, s1, s2)
        assert got["ops"].replace(b"N", b"D") == ref.ops, (k, prm, s1, s2)
