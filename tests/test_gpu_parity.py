"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the C-ABI library
(aligntools/c_b200/libaligntools_b200.so); the oracle is only the checker."""
import hashlib
import random

import numpy as np
import pytest

from helpers import (expected_stdout, load_cli, load_fuzz, pack_batch, parse_cli_argv,
                     sites_from_comment)

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


def rle(ops: bytes) -> str:
    out, k = [], 0
    while k < len(ops):
        j = k
        while j < len(ops) and ops[j] == ops[k]:
            j += 1
        out.append(f"{j - k}{chr(ops[k])}")
        k = j
    return "".join(out)


def check_batch_vs_port(A, aligner, oracle_mod, mode, prm, q, qo, ql, t, to, tl, sites=None, site_off=None,
                        encoding=0, threads=8, q_dev=None, qo_dev=None, t_dev=None, to_dev=None):
    opt = A.Opt(**prm)
    p = oracle_mod.Params(prm["m"], prm["u"], prm["o"], prm["e"], prm["j"], prm["jump"])
    ref = oracle_mod.port_batch(mode, p, q, np.append(qo, 0).astype(np.uint64), ql, t, np.append(to, 0).astype(np.uint64), tl,
                                sites, site_off, want_aln=(mode != "edit"), want_ops=(mode != "edit"), threads=threads)
    b = aligner.batch(mode, opt, q if q_dev is None else q_dev, qo if qo_dev is None else qo_dev, ql,
                      t if t_dev is None else t_dev, to if to_dev is None else to_dev, tl,
                      sites=sites, site_off=site_off, encoding=encoding)
    tm = b.run()
    res = b.fetch()
    b.free()
    n = len(ql)
    assert tm.cells == int((ql.astype(np.uint64) * tl.astype(np.uint64)).sum())
    bad = np.nonzero(res.score.astype(np.int64) != ref.score)[0]
    assert bad.size == 0, (mode, "score mismatch at", bad[:5], res.score[bad[:5]], ref.score[bad[:5]])
    if mode == "edit":
        return tm
    assert np.array_equal(res.end_i, ref.coords[:, 0].astype(np.uint32))
    assert np.array_equal(res.end_j, ref.coords[:, 1].astype(np.uint32))
    for k in range(n):
        assert res.aln(k) == ref.aln(k), (mode, k)
        assert res.cigar_string(k) == rle(ref.op(k)), (mode, k)
    return tm


def test_golden_cli_vectors_through_cabi(A, aligner):
    """Config 1 (B1) and the other 26 rc==0 golden commands: stdout rebuilt from the GPU result
    must hash to the reference's md5."""
    gold = load_cli()
    n = 0
    for v in gold["vectors"]:
        if v["rc"] != 0:
            continue
        mode, prm, fname = parse_cli_argv(v["argv"])
        recs = gold["files"][fname.split("/")[-1]]["records"]
        s1, s2 = recs[0]["seq"].encode(), recs[1]["seq"].encode()
        comment = recs[1]["comment"] or ""
        opt = A.Opt(**prm)
        sites = [sites_from_comment(comment)] if (mode == "fit" and prm["jump"]) else None
        res = aligner.align(mode, [s1], [s2], opt, sites=sites, out_flags=0 if mode == "edit" else 3)
        r1, r2 = (b"", b"") if mode == "edit" else res.aln(0)
        out = expected_stdout(mode, prm, comment, int(res.score[0]), r1, r2)
        assert hashlib.md5(out).hexdigest() == v["stdout_md5"], (v["id"], out[:80])
        n += 1
    assert n >= 27


def test_fuzz_fixtures(A, aligner):
    """900 reference-answered cases, all six mode variants, random (also sign-flipped) params."""
    for c in load_fuzz():
        opt = A.Opt(c["m"], c["u"], c["o"], c["e"], c["j"], bool(c["jump"]))
        sites = [c["sites"] or []] if c["jump"] else None
        res = aligner.align(c["mode"], [c["s1"].encode("latin-1")], [c["s2"].encode("latin-1")], opt, sites=sites,
                            out_flags=0 if c["mode"] == "edit" else 3)
        assert int(res.score[0]) == c["score"], c
        if c["mode"] != "edit":
            assert res.aln(0) == (c["r1"].encode("latin-1"), c["r2"].encode("latin-1")), c


def _random_batch(rng, n, l1_rng, extra_rng, alphabet=b"ACGT", fit=False):
    q, t = [], []
    for _ in range(n):
        l1 = rng.randint(*l1_rng)
        s1 = bytes(rng.choice(alphabet) for _ in range(l1))
        mut = bytearray()
        for ch in s1:
            r = rng.random()
            if r < 0.06:
                mut.append(rng.choice(alphabet))
            elif r < 0.08:
                continue
            elif r < 0.10:
                mut.append(ch); mut.append(rng.choice(alphabet))
            else:
                mut.append(ch)
        s2 = bytes(rng.choice(alphabet) for _ in range(rng.randint(*extra_rng))) + bytes(mut) + \
            bytes(rng.choice(alphabet) for _ in range(rng.randint(1, extra_rng[1] + 1)))
        if fit and len(s1) > len(s2):
            s1, s2 = s2, s1
        q.append(s1); t.append(s2)
    return q, t


@pytest.mark.parametrize("mode", ["global", "local", "fit", "fitjump", "overlap", "edit"])
def test_ragged_batches_all_row_classes(A, aligner, oracle_mod, mode):
    """Ragged batch covering every rows-per-lane class (l1 1..256) and multi-stripe reads
    (l1 up to 900), each pair checked against the oracle: score, end cell, r1/r2, CIGAR."""
    rng = random.Random(1234 + len(mode))
    q, t = _random_batch(rng, 300, (1, 300), (0, 120), fit=mode.startswith("fit"))
    q2, t2 = _random_batch(rng, 40, (257, 900), (0, 300), fit=mode.startswith("fit"))
    q += q2; t += t2
    qb, qo, ql = pack_batch(q)
    tb, to, tl = pack_batch(t)
    prm = dict(m=2, u=-3, o=-4, e=-1, j=-7, jump=(mode == "fitjump"))
    sites = site_off = None
    if mode == "fitjump":
        ss, so = [], [0]
        for s2 in t:
            k = rng.choice([0, 1, 3, 8])
            ss += sorted(rng.randrange(len(s2)) for _ in range(k)); so.append(len(ss))
        sites = np.array(ss + [0], dtype=np.int32); site_off = np.array(so, dtype=np.uint64)
    md = "fit" if mode == "fitjump" else mode
    check_batch_vs_port(A, aligner, oracle_mod, md, prm, qb, qo[:-1].copy(), ql, tb, to[:-1].copy(), tl, sites, site_off)


def test_config2_shape_20k_pairs(A, aligner, oracle_mod):
    """BASELINE config 2 shape (local, 150 x 500, README local params) on 20 000 pairs."""
    from aligntools.c_b200 import synth
    w = synth.config2_local(n_pairs=20000)
    tm = check_batch_vs_port(A, aligner, oracle_mod, "local", w["params"], w["q"], w["q_off"], w["q_len"],
                             w["t"], w["t_off"], w["t_len"])
    assert tm.launches >= 3 and tm.ptr_bytes > 0


def test_config2_2bit_encoding(A, aligner, oracle_mod):
    """Same workload handed over as 2-bit packed device buffers (AT_SEQ_2BIT)."""
    from aligntools.c_b200 import synth
    w = synth.config2_local(n_pairs=3000, stream=1)
    q2, qo2, _ = A.pack_2bit(w["q"], w["q_off"], w["q_len"])
    t2, to2, _ = A.pack_2bit(w["t"], w["t_off"], w["t_len"])
    check_batch_vs_port(A, aligner, oracle_mod, "local", w["params"], w["q"], w["q_off"], w["q_len"],
                        w["t"], w["t_off"], w["t_len"], encoding=1, q_dev=q2, qo_dev=qo2, t_dev=t2, to_dev=to2)


@pytest.mark.parametrize("mode", ["local", "global", "fit"])
@pytest.mark.parametrize("align", [1, 16])
def test_2bit_resident_sequences(A, aligner, oracle_mod, mode, align):
    """AT_SEQ_2BIT input whose pairs all run on K1 stays 2-bit packed in HBM: the fill reads the codes directly (128-bit
    loads when the records start on 16-byte boundaries, byte loads otherwise) and the traceback decodes them for r1 / r2.
    Ragged lengths, packed s16x2 and int32 lanes, score + end cell + alignment strings + CIGAR against the oracle."""
    rng = random.Random(100 + align + len(mode))
    q, t = [], []
    for k in range(400):
        l2 = 200 if k < 250 else rng.randint(2, 700)                # equal l2 -> packed jobs in local mode
        l1 = rng.randint(1, min(l2, 256))
        s2 = bytes(rng.choice(b"ACGT") for _ in range(l2))
        st = rng.randrange(0, l2 - l1 + 1)
        s1 = bytes(c if rng.random() > 0.08 else rng.choice(b"ACGT") for c in s2[st:st + l1])
        q.append(s1); t.append(s2)
    qb, qo, ql = pack_batch(q)
    tb, to, tl = pack_batch(t)
    q2, qo2, _ = A.pack_2bit(qb, qo[:-1].copy(), ql, align=align)
    t2, to2, _ = A.pack_2bit(tb, to[:-1].copy(), tl, align=align)
    prm = dict(m=2, u=-3, o=-4, e=-1, j=-7, jump=False)
    tm = check_batch_vs_port(A, aligner, oracle_mod, mode, prm, qb, qo[:-1].copy(), ql, tb, to[:-1].copy(), tl,
                             encoding=1, q_dev=q2, qo_dev=qo2, t_dev=t2, to_dev=to2)
    assert tm.fill_kernel_flags & 4, "the batch should have stayed 2-bit resident"


def test_config3_shape_fit_jump(A, aligner, oracle_mod):
    """BASELINE config 3 shape: fit -s -j -10, 2 kbp transcripts vs 20 kbp two-gene targets."""
    from aligntools.c_b200 import synth
    w = synth.config3_fit_jump(n_pairs=6)
    check_batch_vs_port(A, aligner, oracle_mod, "fit", w["params"], w["q"], w["q_off"], w["q_len"],
                        w["t"], w["t_off"], w["t_len"], w["sites"], w["site_off"])


def test_config4_shape_overlap(A, aligner, oracle_mod):
    """BASELINE config 4 shape at reduced length (3-6 kbp long-read pairs)."""
    from aligntools.c_b200 import synth
    w = synth.config4_overlap(n_pairs=6, lo=3000, hi=6000)
    check_batch_vs_port(A, aligner, oracle_mod, "overlap", w["params"], w["q"], w["q_off"], w["q_len"],
                        w["t"], w["t_off"], w["t_len"])


def test_config5_shape_edit(A, aligner, oracle_mod):
    """BASELINE config 5 shape at reduced length (20 kbp pairs)."""
    from aligntools.c_b200 import synth
    w = synth.config5_edit(n_pairs=4, length=20000)
    check_batch_vs_port(A, aligner, oracle_mod, "edit", w["params"], w["q"], w["q_off"], w["q_len"],
                        w["t"], w["t_off"], w["t_len"])


def test_chunked_pointer_arena(A, aligner, oracle_mod, monkeypatch):
    """Force several chunks (tiny pointer arena) and scattered, non-monotonic input offsets."""
    from aligntools.c_b200 import synth
    monkeypatch.setenv("AT_PTR_BUDGET_MB", "8")
    w = synth.config2_local(n_pairs=1500, stream=2)
    perm = np.random.default_rng(5).permutation(1500)
    check_batch_vs_port(A, aligner, oracle_mod, "local", w["params"], w["q"], w["q_off"][perm].copy(), w["q_len"],
                        w["t"], w["t_off"][perm].copy(), w["t_len"])


def test_error_codes(A, aligner):
    with pytest.raises(A.AtError) as e:
        aligner.align("fit", [b"ACGTACGT"], [b"ACG"])
    assert e.value.rc == -4          # AT_E_FITLEN, reference dies at :599
    with pytest.raises(A.AtError) as e:
        aligner.align("local", [b""], [b"ACG"])
    assert e.value.rc == -7
    with pytest.raises(A.AtError) as e:
        aligner.align("fit", [b"A"], [b"A"])
    assert e.value.rc == -7


def test_error_codes_large_batch(A, aligner):
    """Batches of 2^17 pairs and more are validated by several host threads over ranges of pairs;
    the error reported must still be the FIRST failing pair's, as a serial scan finds it."""
    n = (1 << 17) + 11
    q = np.frombuffer(b"ACGT" * 2, np.uint8).copy()
    q_off = np.zeros(n, np.uint64); t_off = np.zeros(n, np.uint64)
    q_len = np.full(n, 3, np.uint32); t_len = np.full(n, 5, np.uint32)
    q_len[n - 7] = 0                           # last range: empty record (AT_E_UNDEF)
    q_len[n // 2 + 3] = 6                      # third range: l1 > l2 (AT_E_FITLEN) -- the first in pair order
    with pytest.raises(A.AtError) as e:
        aligner.align_arrays("fit", A.Opt(), q, q_off, q_len, q, t_off, t_len, out_flags=0)
    assert e.value.rc == -4 and ("pair %d:" % (n // 2 + 3)) in str(e.value)
    q_len[5] = 0                               # first range
    with pytest.raises(A.AtError) as e:
        aligner.align_arrays("fit", A.Opt(), q, q_off, q_len, q, t_off, t_len, out_flags=0)
    assert e.value.rc == -7 and "pair 5:" in str(e.value)
    q_len[:] = 3                               # and a clean batch of that size goes through (global 3 x 5, identical pairs)
    res = aligner.align_arrays("global", A.Opt(), q, q_off, q_len, q, t_off, t_len, out_flags=0)
    assert (res.score == res.score[0]).all()


def test_reference_named_operators(A):
    """Single-pair operators with the reference's names (README examples)."""
    sc, r1, r2 = A.align_local_affine(b"PLEASANTLY", b"MEANLY", A.Opt(m=2, u=-2, o=-5, e=-2))
    assert (sc, r1, r2) == (4.0, b"LEA", b"MEA")                      # golden B5
    sc, r1, r2 = A.align_gla(b"PLEASANTLY", b"MEANLY")
    assert (sc, r1, r2) == (-12.0, b"PLEASANTLY", b"M-EAN---LY")      # golden B3
    assert A.edit_dist(b"PLEASANTLY", b"MEANLY", A.Opt(u=1)) == 5      # golden B22
    sc, r1, r2 = A.align_overlap(b"PLEASANTLY", b"MEANLY")
    assert (sc, r1, r2) == (0.0, b"", b"")                             # golden B16
    with pytest.raises(ValueError):
        A.align_fit_affine_jump(b"PLEASANTLY", b"MEANLY")


@pytest.mark.parametrize("mode", ["local", "fitjump", "overlap", "edit"])
def test_pipelined_one_shot_equals_three_call_path(A, aligner, monkeypatch, mode):
    """at_batch_align cuts a large batch into sub-slices that overlap H2D / kernels / D2H on
    several streams; results (scores, cells, dense CIGAR + alignment strings and their offsets)
    must be identical to create + run + fetch on the whole batch."""
    rng = random.Random(77)
    q, t = _random_batch(rng, 600, (20, 500), (0, 150), fit=mode.startswith("fit"))
    qb, qo, ql = pack_batch(q)
    tb, to, tl = pack_batch(t)
    prm = dict(m=2, u=-3, o=-4, e=-1, j=-7, jump=(mode == "fitjump"))
    sites = site_off = None
    if mode == "fitjump":
        ss, so = [], [0]
        for s2 in t:
            ss += sorted(rng.randrange(len(s2)) for _ in range(rng.choice([0, 2, 5]))); so.append(len(ss))
        sites = np.array(ss + [0], dtype=np.int32); site_off = np.array(so, dtype=np.uint64)
    md = "fit" if mode == "fitjump" else mode
    opt = A.Opt(**prm)
    flags = 0 if md == "edit" else 3
    b = aligner.batch(md, opt, qb, qo[:-1].copy(), ql, tb, to[:-1].copy(), tl, sites=sites, site_off=site_off, out_flags=flags)
    b.run()
    ref = b.fetch()
    b.free()
    monkeypatch.setenv("AT_PIPE_MIN_CELLS", "1")
    monkeypatch.setenv("AT_PIPE_SLICE_CELLS", "3000000")        # about 10 sub-slices
    monkeypatch.setenv("AT_PIPE_MIN_TASKS", "1")                # (K2 workloads are otherwise cut only into GPU-filling sub-slices)
    res = aligner.align_arrays(md, opt, qb, qo[:-1].copy(), ql, tb, to[:-1].copy(), tl, sites=sites, site_off=site_off, out_flags=flags)
    assert res.timing.launches > ref.n // 100
    assert np.array_equal(res.score, ref.score)
    if md != "edit":
        for name in ("end_i", "end_j", "beg_i", "beg_j", "cigar_off", "aln_off"):
            assert np.array_equal(getattr(res, name), getattr(ref, name)), name
        no, nc = int(ref.cigar_off[-1]), int(ref.aln_off[-1])
        assert np.array_equal(res.cigar[:no], ref.cigar[:no])
        assert np.array_equal(res.aln1[:nc], ref.aln1[:nc]) and np.array_equal(res.aln2[:nc], ref.aln2[:nc])
    # too small an output buffer is an error, not a silent truncation
    if md != "edit":
        with pytest.raises(A.AtError) as e:
            aligner.align_arrays(md, opt, qb, qo[:-1].copy(), ql, tb, to[:-1].copy(), tl, sites=sites, site_off=site_off,
                                 out_flags=1, cigar_cap=16)
        assert e.value.rc == -5


@pytest.mark.parametrize("mode", ["global", "local", "fit", "fitjump"])
@pytest.mark.parametrize("alphabet", [b"ACGT", b"AC", b"ACGTN", b"ACDEFGHIKLMNPQRSTVWY"])
def test_k1_query_profile_and_fallback_variants(A, aligner, oracle_mod, monkeypatch, mode, alphabet):
    """K1 has two instruction streams: the shared-memory query profile (targets with at most four
    distinct bytes) and the xor/min fallback (any byte alphabet).  Both must equal the oracle, on
    int32 lanes and -- local mode, equal target lengths -- on packed s16x2 lanes."""
    rng = random.Random(4242 + len(alphabet))
    q, t = [], []
    for k in range(160):
        l2 = 180 if k < 120 else rng.randint(40, 260)            # equal l2 -> packed jobs in local mode
        l1 = rng.randint(1, min(l2, 256))
        s2 = bytes(rng.choice(alphabet) for _ in range(l2))
        st = rng.randrange(0, l2 - l1 + 1)
        s1 = bytes(c if rng.random() > 0.1 else rng.choice(alphabet) for c in s2[st:st + l1])
        q.append(s1); t.append(s2)
    qb, qo, ql = pack_batch(q)
    tb, to, tl = pack_batch(t)
    prm = dict(m=2, u=-3, o=-4, e=-1, j=-7, jump=(mode == "fitjump"))
    sites = site_off = None
    if mode == "fitjump":
        ss, so = [], [0]
        for s2 in t:
            ss += sorted(rng.randrange(len(s2)) for _ in range(rng.choice([0, 2, 6]))); so.append(len(ss))
        sites = np.array(ss + [0], dtype=np.int32); site_off = np.array(so, dtype=np.uint64)
    md = "fit" if mode == "fitjump" else mode
    for no_profile in (False, True):
        if no_profile:
            monkeypatch.setenv("AT_NO_PROFILE", "1")
        else:
            monkeypatch.delenv("AT_NO_PROFILE", raising=False)
        check_batch_vs_port(A, aligner, oracle_mod, md, prm, qb, qo[:-1].copy(), ql, tb, to[:-1].copy(), tl, sites, site_off)


def _k2_edge_batch(mode, seed=31337):
    rng = random.Random(seed)
    if mode.startswith("fit"):      # l1 <= l2, l2 >= 2
        shapes = [(257, 257), (257, 300), (513, 513), (1000, 1001), (256, 3000), (258, 259), (2049, 2100), (300, 4097)]
    else:
        shapes = [(257, 1), (300, 5), (1000, 31), (1000, 32), (1000, 33), (513, 64), (2049, 300), (256, 3000),
                  (4000, 17), (258, 255), (258, 256), (258, 257), (1, 1), (1, 700), (700, 1), (257, 4097)]
    q, t = [], []
    for l1, l2 in shapes:
        base = bytes(rng.choice(b"ACGT") for _ in range(max(l1, l2)))
        s1 = bytes(c if rng.random() > 0.07 else rng.choice(b"ACGT") for c in base[:l1])
        s2 = bytes(c if rng.random() > 0.07 else rng.choice(b"ACGT") for c in base[:l2])
        q.append(s1); t.append(s2)
    qb, qo, ql = pack_batch(q)
    tb, to, tl = pack_batch(t)
    prm = dict(m=1, u=-2, o=-3, e=-1, j=-5, jump=(mode == "fitjump"))
    sites = site_off = None
    if mode == "fitjump":
        ss, so = [], [0]
        for s2 in t:
            ss += sorted(rng.randrange(len(s2)) for _ in range(4)); so.append(len(ss))
        sites = np.array(ss + [0], dtype=np.int32); site_off = np.array(so, dtype=np.uint64)
    return ("fit" if mode == "fitjump" else mode), prm, qb, qo[:-1].copy(), ql, tb, to[:-1].copy(), tl, sites, site_off


@pytest.mark.parametrize("mode", ["global", "local", "fit", "fitjump", "overlap", "edit"])
def test_k2_edge_shapes(A, aligner, oracle_mod, mode):
    """Stripe-pipelined K2 at its corners: targets shorter than one 32-column hand-off block or one
    TMA tile, reads of exactly one / one-plus-one stripes, single-column targets, long thin and
    short wide matrices."""
    md, prm, qb, qo, ql, tb, to, tl, sites, site_off = _k2_edge_batch(mode)
    check_batch_vs_port(A, aligner, oracle_mod, md, prm, qb, qo, ql, tb, to, tl, sites, site_off)


@pytest.mark.parametrize("align", [1, 16])
@pytest.mark.parametrize("mode", ["global", "local", "fit", "fitjump", "overlap", "edit", "edit-cells"])
def test_k2_2bit_resident(A, aligner, oracle_mod, monkeypatch, mode, align):
    """AT_SEQ_2BIT input stays 2-bit packed in HBM for the stripe-pipelined kernels too: TMA stages 64-byte tiles of
    packed target, the warp expands them into its shared-memory ring, reads are decoded when a stripe starts, the jump
    masks follow the expanded targets' ring alignment (records at arbitrary byte offsets: align=1)."""
    if mode == "edit-cells":
        monkeypatch.setenv("AT_NO_BITPAR", "1")      # the cell-by-cell single-plane kernel instead of the bit-parallel one
    else:
        monkeypatch.delenv("AT_NO_BITPAR", raising=False)
    md, prm, qb, qo, ql, tb, to, tl, sites, site_off = _k2_edge_batch("edit" if mode == "edit-cells" else mode, seed=4242 + align)
    q2, qo2, _ = A.pack_2bit(qb, qo, ql, align=align)
    t2, to2, _ = A.pack_2bit(tb, to, tl, align=align)
    tm = check_batch_vs_port(A, aligner, oracle_mod, md, prm, qb, qo, ql, tb, to, tl, sites, site_off,
                             encoding=A.SEQ_2BIT, q_dev=q2, qo_dev=qo2, t_dev=t2, to_dev=to2)
    assert tm.fill_kernel_flags & 4, "the batch was expanded to bytes instead of staying 2-bit resident"
    if mode == "fitjump":      # the same lists as a whitelist
        prm = dict(prm, jump=2)
        opt = A.Opt(m=prm["m"], u=prm["u"], o=prm["o"], e=prm["e"], j=prm["j"], jump=True, whitelist=True)
        p = oracle_mod.Params(prm["m"], prm["u"], prm["o"], prm["e"], prm["j"], 2)
        ref = oracle_mod.port_batch("fit", p, qb, np.append(qo, 0).astype(np.uint64), ql, tb, np.append(to, 0).astype(np.uint64), tl,
                                    sites, site_off, want_aln=True, want_ops=True, threads=8)
        res = aligner.align_arrays("fit", opt, q2, qo2, ql, t2, to2, tl, sites=sites, site_off=site_off, encoding=A.SEQ_2BIT)
        assert np.array_equal(res.score.astype(np.int64), ref.score)
        for k in range(len(ql)):
            assert res.cigar_string(k) == rle(ref.op(k)), k


@pytest.mark.parametrize("mode", ["global", "local", "fit", "overlap", "edit"])
def test_k2_2bit_ring_shift(A, aligner, oracle_mod, mode):
    """2-bit targets whose packed records start 15 / 14 bytes past a 16-byte boundary: the expanded target sits up to 60
    columns into K2's ring, so lane 0 enters the next 256-column tile up to 60 steps before the step counter does
    (regression: the tile used to be awaited 32 steps ahead only)."""
    rng = random.Random(77)
    shapes = [(300, 60), (300, 700), (300, 300), (520, 1030), (300, 2000)]      # packed targets: 15, 175, 75, 258, 500 bytes
    if mode == "fit":
        shapes = [(30, 60)] + shapes[1:]
    q, t = [], []
    for l1, l2 in shapes:
        base = bytes(rng.choice(b"ACGT") for _ in range(max(l1, l2)))
        q.append(bytes(c if rng.random() > 0.1 else rng.choice(b"ACGT") for c in base[:l1]))
        t.append(bytes(c if rng.random() > 0.1 else rng.choice(b"ACGT") for c in base[:l2]))
    qb, qo, ql = pack_batch(q)
    tb, to, tl = pack_batch(t)
    q2, qo2, _ = A.pack_2bit(qb, qo[:-1].copy(), ql)
    t2, to2, _ = A.pack_2bit(tb, to[:-1].copy(), tl)
    assert [int(x) & 15 for x in to2[:3]] == [0, 15, 14]
    prm = dict(m=1, u=-2, o=-3, e=-1, j=-5, jump=False)
    check_batch_vs_port(A, aligner, oracle_mod, mode, prm, qb, qo[:-1].copy(), ql, tb, to[:-1].copy(), tl,
                        encoding=A.SEQ_2BIT, q_dev=q2, qo_dev=qo2, t_dev=t2, to_dev=to2)


@pytest.mark.parametrize("alphabet", [b"ACGT", b"A", b"ACGTNRYK", b"ACGTNRYKM", b"ACDEFGHIKLMNPQRSTVWY"])
def test_edit_unit_cost_bit_parallel(A, aligner, oracle_mod, monkeypatch, alphabet):
    """`edit -u 1` runs the bit-parallel (Myers) kernel when the reads use at most 8 distinct bytes,
    the cell-by-cell kernel otherwise; both must give the reference's distance.  Shapes cover one
    block, partial blocks, 1 / 2 / 4 blocks per lane, several stripes, targets shorter than a
    hand-off block, and target symbols that never occur in the reads."""
    rng = random.Random(99 + len(alphabet))
    shapes = [(1, 1), (1, 40), (40, 1), (31, 33), (32, 32), (33, 31), (100, 5), (5, 100), (1024, 900), (1025, 1100),
              (2048, 2000), (2049, 70), (4096, 4096), (4097, 4200), (9000, 3000), (3000, 9000), (700, 17), (17, 700)]
    q, t = [], []
    for l1, l2 in shapes:
        base = bytes(rng.choice(alphabet) for _ in range(max(l1, l2)))
        s1 = bytes(c if rng.random() > 0.1 else rng.choice(alphabet) for c in base[:l1])
        s2 = bytearray(c if rng.random() > 0.1 else rng.choice(alphabet) for c in base[:l2])
        for k in range(0, l2, 37):
            s2[k] = ord("#")                                   # a byte no read contains
        q.append(s1); t.append(bytes(s2))
    for _ in range(60):
        l1, l2 = rng.randint(1, 300), rng.randint(1, 300)
        q.append(bytes(rng.choice(alphabet) for _ in range(l1))); t.append(bytes(rng.choice(alphabet) for _ in range(l2)))
    qb, qo, ql = pack_batch(q)
    tb, to, tl = pack_batch(t)
    prm = dict(m=1, u=1, o=-5, e=-1, j=-10, jump=False)
    check_batch_vs_port(A, aligner, oracle_mod, "edit", prm, qb, qo[:-1].copy(), ql, tb, to[:-1].copy(), tl)
    monkeypatch.setenv("AT_NO_BITPAR", "1")
    check_batch_vs_port(A, aligner, oracle_mod, "edit", prm, qb, qo[:-1].copy(), ql, tb, to[:-1].copy(), tl)


def test_config2_full_size_checksums(A, aligner):
    """BASELINE config 2 at its full size (1 Mi pairs): a checksum of checksums against the run that was
    compared with the oracle pair by pair (tests/golden/c2_full_checksums.json), through both the resident
    and the pipelined one-shot path; plus size-independent properties of every alignment -- CIGAR column
    counts consistent with the cells the traceback started and stopped in."""
    import json
    import os
    from aligntools.c_b200 import synth
    from helpers import GOLD
    with open(os.path.join(GOLD, "c2_full_checksums.json")) as f:
        gold = json.load(f)
    w = synth.config2_local(n_pairs=gold["pairs"])
    opt = A.Opt(**w["params"])
    b = aligner.batch("local", opt, w["q"], w["q_off"], w["q_len"], w["t"], w["t_off"], w["t_len"], out_flags=A.OUT_CIGAR)
    tm = b.run()
    res = b.fetch()
    b.free()
    n = gold["pairs"]
    assert tm.cells == gold["cells"]
    assert int(res.score.astype(np.int64).sum()) == gold["score_sum"]
    assert int(res.cigar_off[n]) == gold["cigar_ops"]
    ops = res.cigar[:gold["cigar_ops"]]
    runs = (ops >> 4).astype(np.int64)
    code = ops & 3
    assert int(runs.sum()) == gold["alignment_columns"]
    # per pair: rows consumed = M + I columns = end_i - beg_i, target columns consumed = M + D = end_j - beg_j
    starts = res.cigar_off[:n].astype(np.int64)
    nonempty = res.cigar_off[1:n + 1] > res.cigar_off[:n]
    rows = np.add.reduceat(np.where(code != 2, runs, 0), starts[nonempty])
    cols = np.add.reduceat(np.where(code != 1, runs, 0), starts[nonempty])
    assert np.array_equal(rows, (res.end_i.astype(np.int64) - res.beg_i)[nonempty])
    assert np.array_equal(cols, (res.end_j.astype(np.int64) - res.beg_j)[nonempty])
    assert np.all(res.score >= 0) and np.all(res.end_i <= 150) and np.all(res.end_j <= 500)
    one = aligner.align_arrays("local", opt, w["q"], w["q_off"], w["q_len"], w["t"], w["t_off"], w["t_len"], out_flags=A.OUT_CIGAR,
                               cigar_cap=gold["cigar_ops"] + 16)
    assert np.array_equal(one.score, res.score) and np.array_equal(one.cigar_off, res.cigar_off)
    assert np.array_equal(one.cigar[:gold["cigar_ops"]], ops)


def test_junction_whitelist_mode(A, aligner, oracle_mod):
    """at_params.jump == 2 -- entering the jump state only ON the listed target indices (the
    semantics of the comment at src/alignment.h:542-544) -- on K1 (short reads) and K2 (multi-stripe reads) against
    the oracle's restatement, which tests/test_oracle.py pins on the compiled reference."""
    rng = random.Random(2468)
    q, t, ss, so = [], [], [], [0]
    for k in range(120):
        l2 = rng.randint(40, 300) if k < 100 else rng.randint(1500, 4000)
        s2 = bytes(rng.choice(b"ACGT") for _ in range(l2))
        a, b = sorted(rng.sample(range(10, l2 - 10), 2))
        ex = rng.randint(5, 120) if k < 100 else rng.randint(200, 700)
        s1 = (s2[max(0, a - ex):a] + s2[b:b + ex]) or b"A"
        s1 = bytes(c if rng.random() > 0.04 else rng.choice(b"ACGT") for c in s1)
        q.append(s1); t.append(s2)
        ss += sorted(set([a, b - 1, b] + [rng.randrange(l2) for _ in range(rng.choice([0, 3]))])); so.append(len(ss))
    qb, qo, ql = pack_batch(q)
    tb, to, tl = pack_batch(t)
    sites = np.array(ss + [0], dtype=np.int32); site_off = np.array(so, dtype=np.uint64)
    prm = dict(m=1, u=-2, o=-5, e=-1, j=-6)
    ref = oracle_mod.port_batch("fit", oracle_mod.Params(**prm, jump=2), qb, qo, ql, tb, to, tl, sites, site_off, want_aln=True, want_ops=True, threads=4)
    b = aligner.batch("fit", A.Opt(**prm, jump=True, whitelist=True), qb, qo[:-1].copy(), ql, tb, to[:-1].copy(), tl, sites=sites, site_off=site_off)
    b.run()
    res = b.fetch()
    b.free()
    assert np.array_equal(res.score.astype(np.int64), ref.score)
    n_jump = 0
    for k in range(len(q)):
        assert res.aln(k) == ref.aln(k), k
        assert res.cigar_string(k) == rle(ref.op(k)), k
        n_jump += "N" in res.cigar_string(k)
    assert n_jump > 30
    # and the same list as a blacklist gives different alignments (the modes are not confused)
    res_b = aligner.align("fit", q[:40], t[:40], A.Opt(**prm, jump=True), sites=[ss[so[k]:so[k + 1]] for k in range(40)])
    assert any(res_b.aln(k) != res.aln(k) for k in range(40))


@pytest.mark.parametrize("mode", ["global", "fit"])
@pytest.mark.parametrize("alphabet", [b"ACGT", b"ACDEFGHIKLMNPQRSTVWY"])
def test_k1_packed_lanes_global_and_fit(A, aligner, oracle_mod, mode, alphabet):
    """Global and fit on packed s16x2 lanes (two pairs per warp when they share l1 and l2; -inf is AT_NEG16 inside the
    16 bits): uniform shapes of every rows-per-lane class, query-profile and xor/min variants, sign-flipped parameters,
    against the oracle -- and a batch whose score range does not fit 16 bits, which must fall back to int32 lanes."""
    rng = random.Random(606 + len(alphabet) + len(mode))
    for shape_k, (l1, l2) in enumerate([(150, 150), (150, 400), (31, 64), (97, 97), (200, 333), (256, 256), (1, 9)]):
        if mode == "fit" and l1 > l2:
            continue
        q, t = [], []
        for _ in range(37):
            s2 = bytes(rng.choice(alphabet) for _ in range(l2))
            st = rng.randrange(0, l2 - l1 + 1)
            s1 = bytearray(c if rng.random() > 0.1 else rng.choice(alphabet) for c in s2[st:st + l1])
            if rng.random() < 0.5 and l1 > 20:      # an indel, keeping the length
                k = rng.randrange(5, l1 - 5)
                del s1[k]; s1.append(rng.choice(alphabet))
            q.append(bytes(s1)); t.append(s2)
        qb, qo, ql = pack_batch(q)
        tb, to, tl = pack_batch(t)
        prms = [dict(m=1, u=-1, o=-4, e=-1, j=-10, jump=False), dict(m=2, u=-3, o=-5, e=-2, j=-10, jump=False)]
        if shape_k < 3:
            prms.append(dict(m=2, u=-1, o=1, e=-1, j=-10, jump=False))      # a profitable gap opening
        for prm in prms:
            tm = check_batch_vs_port(A, aligner, oracle_mod, mode, prm, qb, qo[:-1].copy(), ql, tb, to[:-1].copy(), tl)
            if 8 * (l1 + l2 + 2) * max(abs(v) for k, v in prm.items() if k in "muoe") < 24000:
                assert tm.fill_kernel_kind == 1, "expected packed s16x2 lanes"
    # too wide a score range for 16 bits: int32 lanes, same answers
    l1, l2 = 250, 900
    q = [bytes(rng.choice(alphabet) for _ in range(l1)) for _ in range(8)]
    t = [bytes(rng.choice(alphabet) for _ in range(l2)) for _ in range(8)]
    qb, qo, ql = pack_batch(q)
    tb, to, tl = pack_batch(t)
    tm = check_batch_vs_port(A, aligner, oracle_mod, mode, dict(m=5, u=-4, o=-9, e=-3, j=-10, jump=False), qb, qo[:-1].copy(), ql, tb, to[:-1].copy(), tl)
    assert tm.fill_kernel_kind == 0
