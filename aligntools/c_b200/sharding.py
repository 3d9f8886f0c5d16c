"""Multi-process sharding of a batch: one rank per GPU (torch.distributed), each rank aligns a
CONTIGUOUS slice of the pairs and rank 0 gathers the results -- no collective on the data path,
because pairs never communicate.  torch.distributed is plumbing only: a barrier-free
gather of the per-rank result arrays (`gloo` on CPU hosts, `nccl`/`gloo` on GPU boxes).

    cut = plan_slices(q_len, t_len, world)            # C-ABI at_plan_slices: balanced by DP cells
    lo, hi = rank_slice(cut, rank)
    res = aligner.align_arrays(mode, opt, q, q_off[lo:hi], q_len[lo:hi], t, t_off[lo:hi], t_len[lo:hi], ...)
    full = gather_results(res, cut, rank, world)       # rank 0: BatchResult of the whole batch
"""
from __future__ import annotations

import numpy as np

from . import BatchResult, plan_slices  # noqa: F401  (re-exported)


def rank_slice(cut, rank):
    return int(cut[rank]), int(cut[rank + 1])


def merge_results(parts, cut):
    """Concatenate per-slice BatchResults (slice r = pairs [cut[r], cut[r+1])) into one, rebasing
    the dense CIGAR / alignment offsets -- the host-side gather of the sharded batch."""
    n = int(cut[-1])
    out = BatchResult(n)
    have_cig = all(p.cigar_off is not None for p in parts)
    have_aln = all(p.aln_off is not None for p in parts)
    if have_cig:
        out.cigar_off = np.zeros(n + 1, np.uint64)
        out.cigar = np.zeros(sum(int(p.cigar_off[p.n]) for p in parts) + 1, np.uint32)
    if have_aln:
        out.aln_off = np.zeros(n + 1, np.uint64)
        tot = sum(int(p.aln_off[p.n]) for p in parts)
        out.aln1 = np.zeros(tot + 1, np.uint8)
        out.aln2 = np.zeros(tot + 1, np.uint8)
    base_ops = base_cols = 0
    for r, p in enumerate(parts):
        lo, hi = int(cut[r]), int(cut[r + 1])
        assert p.n == hi - lo, "slice size does not match the plan"
        for name in ("score", "end_i", "end_j", "beg_i", "beg_j"):
            getattr(out, name)[lo:hi] = getattr(p, name)[:p.n]
        if have_cig:
            k = int(p.cigar_off[p.n])
            out.cigar_off[lo:hi] = p.cigar_off[:p.n] + np.uint64(base_ops)
            out.cigar[base_ops:base_ops + k] = p.cigar[:k]
            base_ops += k
        if have_aln:
            k = int(p.aln_off[p.n])
            out.aln_off[lo:hi] = p.aln_off[:p.n] + np.uint64(base_cols)
            out.aln1[base_cols:base_cols + k] = p.aln1[:k]
            out.aln2[base_cols:base_cols + k] = p.aln2[:k]
            base_cols += k
    if have_cig:
        out.cigar_off[n] = base_ops
    if have_aln:
        out.aln_off[n] = base_cols
    return out


def gather_results(res, cut, rank, world, dst=0):
    """Gather every rank's BatchResult on rank `dst` (torch.distributed.gather_object) and merge
    them in pair order.  Returns the merged result on `dst`, None elsewhere."""
    import torch.distributed as dist
    payload = {k: getattr(res, k) for k in ("n", "score", "end_i", "end_j", "beg_i", "beg_j", "cigar", "cigar_off",
                                            "aln1", "aln2", "aln_off")}
    box = [None] * world if rank == dst else None
    dist.gather_object(payload, box, dst=dst)
    if rank != dst:
        return None
    parts = []
    for d in box:
        p = BatchResult(d["n"])
        for k, v in d.items():
            setattr(p, k, v)
        parts.append(p)
    return merge_results(parts, cut)
