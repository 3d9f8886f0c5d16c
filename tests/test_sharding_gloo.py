"""The N>1 host path on CPU: the C-ABI sharding plan (at_plan_slices) and the multi-process gather
(aligntools.c_b200.sharding) under torch.distributed `gloo`, world_size 2.  The per-rank compute is
stubbed with the oracle port here (tests may use the oracle; the product has no CPU compute path) --
what is tested is the plan, the slice hand-out and the rebased gather, which are the same on a GPU box."""
import os
import random

import numpy as np
import pytest

from helpers import pack_batch


def _batch(n=90, seed=3):
    rng = random.Random(seed)
    q, t = [], []
    for _ in range(n):
        l1 = rng.randint(5, 120)
        s1 = bytes(rng.choice(b"ACGT") for _ in range(l1))
        s2 = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 20))) + s1 + bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 200)))
        q.append(s1); t.append(s2)
    return q, t


def test_plan_slices_is_contiguous_and_balanced():
    import aligntools.c_b200 as A
    A.build()
    q, t = _batch(500, seed=11)
    _, _, ql = pack_batch(q)
    _, _, tl = pack_batch(t)
    cells = ql.astype(np.uint64) * tl.astype(np.uint64)
    for parts in (1, 2, 3, 8, 64):
        cut = A.plan_slices(ql, tl, parts)
        assert cut[0] == 0 and cut[-1] == len(ql) and np.all(np.diff(cut.astype(np.int64)) >= 0)
        per = [int(cells[int(cut[r]):int(cut[r + 1])].sum()) for r in range(parts)]
        assert sum(per) == int(cells.sum())
        if parts <= 8:
            assert max(per) - min(per) <= 2 * int(cells.max()), (parts, per)
    # more parts than pairs: empty slices, still a cover
    cut = A.plan_slices(ql[:3], tl[:3], 8)
    assert cut[0] == 0 and cut[-1] == 3 and len(cut) == 9


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist
    import aligntools.c_b200 as A
    from aligntools.c_b200 import sharding
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q, t = _batch()
    qb, qo, ql = pack_batch(q)
    tb, to, tl = pack_batch(t)
    cut = sharding.plan_slices(ql, tl, world)
    lo, hi = sharding.rank_slice(cut, rank)
    p = oracle.Params(2, -3, -4, -1, -10, False)

    def compute(lo, hi):      # stands in for aligner.align_arrays on this rank's GPU
        ref = oracle.port_batch("local", p, qb, qo[lo:hi + 1].copy(), ql[lo:hi].copy(), tb, to[lo:hi + 1].copy(), tl[lo:hi].copy(),
                                want_aln=True, want_ops=True, threads=1)
        n = hi - lo
        r = A.BatchResult(n)
        r.score[:] = ref.score
        r.end_i[:], r.end_j[:] = ref.coords[:, 0], ref.coords[:, 1]
        r.beg_i[:], r.beg_j[:] = ref.coords[:, 2], ref.coords[:, 3]
        r.aln_off = np.zeros(n + 1, np.uint64)
        np.cumsum(ref.aln_len, out=r.aln_off[1:])
        r.aln1 = np.concatenate([np.frombuffer(ref.aln(k)[0], np.uint8) for k in range(n)] + [np.zeros(1, np.uint8)])
        r.aln2 = np.concatenate([np.frombuffer(ref.aln(k)[1], np.uint8) for k in range(n)] + [np.zeros(1, np.uint8)])
        return r

    mine = compute(lo, hi)
    full = sharding.gather_results(mine, cut, rank, world)
    if rank == 0:
        whole = compute(0, len(ql))
        assert np.array_equal(full.score, whole.score)
        for name in ("end_i", "end_j", "beg_i", "beg_j", "aln_off"):
            assert np.array_equal(getattr(full, name), getattr(whole, name)), name
        for k in range(len(ql)):
            assert full.aln(k) == whole.aln(k)
        open(os.path.join(tmpdir, "ok"), "w").write("1")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_gather(tmp_path, oracle_mod):
    import torch.multiprocessing as mp
    import aligntools.c_b200 as A
    A.build()
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def _worker_shm(rank, world, port, tmpdir):
    """HostGather: every rank writes its slice straight into rank 0's shared-memory output arrays."""
    import torch.distributed as dist
    import aligntools.c_b200 as A
    from aligntools.c_b200 import sharding
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q, t = _batch(70, seed=21)
    qb, qo, ql = pack_batch(q)
    tb, to, tl = pack_batch(t)
    n = len(ql)
    cut = sharding.plan_slices(ql, tl, world)
    lo, hi = sharding.rank_slice(cut, rank)
    p = oracle.Params(2, -3, -4, -1, -10, False)
    cap = 4096
    hg = sharding.HostGather(dist, rank, world, n, cap, "t")
    out = hg.slice_out(lo, hi)

    def rle_ops(ops):         # per-column op letters -> packed run-length ops (AT_CIG_*)
        res, k = [], 0
        while k < len(ops):
            j = k
            while j < len(ops) and ops[j] == ops[k]:
                j += 1
            res.append(((j - k) << 4) | b"MIDN".index(ops[k]))
            k = j
        return res

    # stands in for at_batch_align on this rank's GPU: fills the caller-provided (shared) output views
    ref = oracle.port_batch("local", p, qb, qo[lo:hi + 1].copy(), ql[lo:hi].copy(), tb, to[lo:hi + 1].copy(), tl[lo:hi].copy(),
                            want_aln=True, want_ops=True, threads=1)
    out.score[:] = ref.score
    out.end_i[:], out.end_j[:] = ref.coords[:, 0], ref.coords[:, 1]
    out.beg_i[:], out.beg_j[:] = ref.coords[:, 2], ref.coords[:, 3]
    pos = 0
    for k in range(hi - lo):
        out.cigar_off[k] = pos
        for op in rle_ops(ref.op(k)):
            out.cigar[pos] = op; pos += 1
    out.cigar_off[hi - lo] = pos
    dist.barrier()
    if rank == 0:
        full_c = np.zeros(cap * world, np.uint32); full_off = np.zeros(n + 1, np.uint64)
        tot = hg.merge(cut, full_c, full_off)
        whole = oracle.port_batch("local", p, qb, qo, ql, tb, to, tl, want_aln=True, want_ops=True, threads=1)
        assert np.array_equal(hg.arr["score"][:n].astype(np.int64), whole.score)
        assert np.array_equal(hg.arr["end_j"][:n], whole.coords[:, 1].astype(np.uint32))
        pos = 0
        for k in range(n):
            ops = rle_ops(whole.op(k))
            assert int(full_off[k]) == pos and list(full_c[pos:pos + len(ops)]) == ops, k
            pos += len(ops)
        assert tot == pos == int(full_off[n])
        open(os.path.join(tmpdir, "ok_shm"), "w").write("1")
    dist.barrier()
    hg.close()
    dist.destroy_process_group()


def test_two_rank_shared_memory_host_gather(tmp_path, oracle_mod):
    import torch.multiprocessing as mp
    import aligntools.c_b200 as A
    A.build()
    port = 31000 + os.getpid() % 2000
    mp.spawn(_worker_shm, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok_shm").exists()
