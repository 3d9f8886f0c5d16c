"""GPU parity at the FULL shapes of BASELINE.json (run with -m gpu on a B200), against the compiled,
unmodified reference (oracle/_ref) wherever its 48 B/cell matrices fit in host memory:

  C2  a random 10 000-pair subset of the 1 Mi-pair batch, taken from the FULL-batch run     (reference)
  C3  fit -s -j -10, 2 kbp x 20 kbp                                                          (reference)
  C4  overlap, 10-20 kbp pairs                                                               (reference)
  C5  edit -u 1, 100 kbp x 100 kbp (the reference would need 480 GB: integer port)           (port)

plus the score-range boundary of the int32 lanes (a -inf stand-in must never beat a finite value).
Everything goes through the C-ABI library; the oracle is only the checker."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    import aligntools.c_b200 as A
    return A


@pytest.fixture(scope="module")
def aligner(A):
    al = A.Aligner()
    yield al
    al.close()


def _ref(oracle_mod):
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref not built (it is compiled where /root/reference exists and travels with the snapshot)")
    return oracle_mod


def _params(oracle_mod, prm):
    return oracle_mod.Params(prm["m"], prm["u"], prm["o"], prm["e"], prm["j"], prm["jump"])


def _off1(off):
    return np.append(off, 0).astype(np.uint64)


def test_c3_full_shape_vs_reference(A, aligner, oracle_mod):
    """fit with jump state at BASELINE's shape (2 000 x 20 000): score, r1 and r2 of every pair equal the reference's."""
    from aligntools.c_b200 import synth
    orc = _ref(oracle_mod)
    w = synth.config3_fit_jump(n_pairs=3, stream=7)
    ref = orc.ref_batch("fit", _params(orc, w["params"]), w["q"], _off1(w["q_off"]), w["q_len"], w["t"], _off1(w["t_off"]), w["t_len"],
                        w["sites"], w["site_off"], want_aln=True, threads=1)
    b = aligner.batch("fit", A.Opt(**w["params"]), w["q"], w["q_off"], w["q_len"], w["t"], w["t_off"], w["t_len"],
                      sites=w["sites"], site_off=w["site_off"])
    b.run()
    res = b.fetch()
    b.free()
    assert np.array_equal(res.score.astype(np.int64), ref.score)
    n_jump = 0
    for k in range(3):
        assert res.aln(k) == ref.aln(k), k
        n_jump += res.cigar_string(k).count("N")
    assert n_jump > 0                              # the workload does exercise the jump state


def test_c4_full_shape_vs_reference(A, aligner, oracle_mod):
    """overlap at BASELINE's shape (10-20 kbp long-read pairs) against the reference itself (one pair at a time:
    its matrices are 48 B per cell)."""
    from aligntools.c_b200 import synth
    orc = _ref(oracle_mod)
    w = synth.config4_overlap(n_pairs=4, stream=9)
    assert int(w["q_len"].min()) >= 10000 and int(w["t_len"].min()) >= 10000
    ref = orc.ref_batch("overlap", _params(orc, w["params"]), w["q"], _off1(w["q_off"]), w["q_len"], w["t"], _off1(w["t_off"]), w["t_len"],
                        want_aln=True, threads=1)
    b = aligner.batch("overlap", A.Opt(**w["params"]), w["q"], w["q_off"], w["q_len"], w["t"], w["t_off"], w["t_len"])
    b.run()
    res = b.fetch()
    b.free()
    assert np.array_equal(res.score.astype(np.int64), ref.score)
    assert int(res.score.max()) > 1000
    for k in range(4):
        assert res.aln(k) == ref.aln(k), k


def test_c5_full_shape_vs_port(A, aligner, oracle_mod):
    """edit -u 1 at BASELINE's shape, 100 kbp x 100 kbp, one pair, bit-parallel and cell-by-cell kernels
    against the integer port (about 40 s of one host core)."""
    from aligntools.c_b200 import synth
    w = synth.config5_edit(n_pairs=1, length=100000, stream=3)
    ref = oracle_mod.port_batch("edit", _params(oracle_mod, w["params"]), w["q"], _off1(w["q_off"]), w["q_len"], w["t"], _off1(w["t_off"]),
                                w["t_len"], want_aln=False, threads=1)
    res = aligner.align_arrays("edit", A.Opt(**w["params"]), w["q"], w["q_off"], w["q_len"], w["t"], w["t_off"], w["t_len"], out_flags=0)
    assert int(res.score[0]) == int(ref.score[0])
    os.environ["AT_NO_BITPAR"] = "1"
    try:
        res2 = aligner.align_arrays("edit", A.Opt(**w["params"]), w["q"], w["q_off"], w["q_len"], w["t"], w["t_off"], w["t_len"], out_flags=0)
    finally:
        del os.environ["AT_NO_BITPAR"]
    assert int(res2.score[0]) == int(ref.score[0])


def _replay(cigar_ops, s1, s2, i, j):
    """r1 / r2 from a CIGAR and the cell the traceback stopped in."""
    r1, r2 = bytearray(), bytearray()
    for op in cigar_ops:
        n, c = int(op) >> 4, int(op) & 15
        if c == 0:
            r1 += s1[i:i + n]; r2 += s2[j:j + n]; i += n; j += n
        elif c == 1:
            r1 += s1[i:i + n]; r2 += b"-" * n; i += n
        else:
            r1 += b"-" * n; r2 += s2[j:j + n]; j += n
    return bytes(r1), bytes(r2)


def test_c2_full_batch_random_subset_vs_reference(A, aligner, oracle_mod):
    """BASELINE config 2 at its full size: ONE run over all 1 Mi pairs (resident path and the pipelined one-shot
    path with 2-bit input), then 10 000 randomly chosen pairs of THAT run against the compiled reference --
    score and both gapped strings, rebuilt from the run's CIGARs and start cells."""
    from aligntools.c_b200 import synth
    orc = _ref(oracle_mod)
    n = 1 << 20
    w = synth.config2_local(n_pairs=n)
    opt = A.Opt(**w["params"])
    b = aligner.batch("local", opt, w["q"], w["q_off"], w["q_len"], w["t"], w["t_off"], w["t_len"], out_flags=A.OUT_CIGAR)
    tm = b.run()
    res = b.fetch()
    b.free()
    assert tm.cells == n * 150 * 500
    pick = np.sort(np.random.default_rng(20261018).choice(n, 10000, replace=False))
    ref = orc.ref_batch("local", _params(orc, w["params"]), w["q"], w["q_off"][pick].copy(), w["q_len"][pick].copy(),
                        w["t"], w["t_off"][pick].copy(), w["t_len"][pick].copy(), want_aln=True, threads=os.cpu_count() or 4)
    assert np.array_equal(res.score[pick].astype(np.int64), ref.score)
    qb, tb = w["q"].tobytes(), w["t"].tobytes()
    for x, p in enumerate(pick):
        s1 = qb[int(w["q_off"][p]):int(w["q_off"][p]) + 150]
        s2 = tb[int(w["t_off"][p]):int(w["t_off"][p]) + 500]
        assert _replay(res.cigar_ops(p), s1, s2, int(res.beg_i[p]), int(res.beg_j[p])) == ref.aln(x), p
    # the one-shot path on 2-bit input gives the same million answers
    q2, qo2, _ = A.pack_2bit(w["q"], w["q_off"], w["q_len"])
    t2, to2, _ = A.pack_2bit(w["t"], w["t_off"], w["t_len"])
    nops = int(res.cigar_off[n])
    one = aligner.align_arrays("local", opt, q2, qo2, w["q_len"], t2, to2, w["t_len"], out_flags=A.OUT_CIGAR, encoding=A.SEQ_2BIT,
                               cigar_cap=nops + 16)
    assert np.array_equal(one.score, res.score) and np.array_equal(one.cigar_off, res.cigar_off)
    assert np.array_equal(one.cigar[:nops], res.cigar[:nops])
    assert np.array_equal(one.beg_i, res.beg_i) and np.array_equal(one.beg_j, res.beg_j)


def test_score_range_boundary(A, aligner, oracle_mod):
    """Largest |e| x length the int32 lanes accept: long gaps drive finite scores to -2.6e7 (x8 in the lanes), where
    they must still beat every -inf stand-in; one step further the batch is refused (AT_E_RANGE), never mis-scored."""
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    l1, l2 = 30000, 3500
    s1 = acgt[rng.integers(0, 4, l1)]
    s2 = s1[:l2].copy()
    q, t = np.ascontiguousarray(s1), np.ascontiguousarray(s2)
    z = np.zeros(1, np.uint64)
    for mode, prm in (("global", dict(m=1, u=-2, o=-5, e=-1000, j=-10, jump=False)),
                      ("global", dict(m=1000, u=-1000, o=-1000, e=-1, j=-10, jump=False)),
                      ("local", dict(m=1, u=-2, o=-1000, e=-1000, j=-10, jump=False))):
        assert (l1 + l2 + 2) * 1000 < (1 << 25)
        ref = oracle_mod.port_batch(mode, _params(oracle_mod, prm), q, np.zeros(2, np.uint64), np.array([l1], np.uint32),
                                    t, np.zeros(2, np.uint64), np.array([l2], np.uint32), want_aln=True, threads=1)
        res = aligner.align_arrays(mode, A.Opt(**prm), q, z, np.array([l1], np.uint32), t, z, np.array([l2], np.uint32),
                                   out_flags=A.OUT_CIGAR | A.OUT_ALN)
        assert int(res.score[0]) == int(ref.score[0]), (mode, prm)
        assert res.aln(0) == ref.aln(0), (mode, prm)
    # ADVICE r1's example: global, l1 = 70 000, l2 = 1, e = -1000 passed the old check and was mis-scored
    s1 = acgt[rng.integers(0, 4, 70000)]
    with pytest.raises(A.AtError) as e:
        aligner.align_arrays("global", A.Opt(1, -2, -5, -1000), np.ascontiguousarray(s1), z, np.array([70000], np.uint32),
                             np.ascontiguousarray(s1[:1]), z, np.array([1], np.uint32), out_flags=0)
    assert e.value.rc == -6
