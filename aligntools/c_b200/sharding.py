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

import ctypes as C
import os

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


class HostGather:
    """Results of a sharded batch gathered in rank 0's host memory WITHOUT a collective (ranks of one box): rank 0
    owns shared-memory output arrays of the whole batch; every rank's at_batch_align writes its slice's scores /
    cells straight into them at the slice offset and its CIGARs into its own shared region, which rank 0 then
    concatenates in pair order (one memcpy per rank).  torch.distributed only carries the segment names (and the
    caller's barrier); `tag` keeps the segment names of concurrent gathers apart."""

    def __init__(self, dist, rank, world, n_pairs, cigar_cap_per_rank, tag, pin=False):
        from multiprocessing import shared_memory
        self.rank, self.world, self.n = rank, world, n_pairs
        self.cap = int(cigar_cap_per_rank)
        self.sizes = {"score": 4 * n_pairs, "end_i": 4 * n_pairs, "end_j": 4 * n_pairs, "beg_i": 4 * n_pairs, "beg_j": 4 * n_pairs,
                      "cigar_off": 8 * (n_pairs + world), "cigar": 4 * self.cap * world, "nops": 8 * world}
        names = [None]
        self.shm = {}
        if rank == 0:
            tag = "".join(ch for ch in str(tag) if ch.isalnum())[:24]
            names = [{k: f"atb2_{tag}_{os.getpid()}_{k}" for k in self.sizes}]
            for k, sz in self.sizes.items():
                self.shm[k] = shared_memory.SharedMemory(name=names[0][k], create=True, size=max(sz, 8))
        if world > 1:
            dist.broadcast_object_list(names, src=0)
        if rank != 0:
            from multiprocessing import resource_tracker
            for k in self.sizes:
                self.shm[k] = shared_memory.SharedMemory(name=names[0][k])
                try:      # the segments belong to rank 0: keep this process's tracker from unlinking them at exit
                    resource_tracker.unregister(self.shm[k]._name, "shared_memory")
                except Exception:
                    pass
        # page-lock the segments in THIS process (every rank's device copies land in them directly); best effort
        self.registered = []
        if pin:
            from . import load_library
            lib = load_library()
            for k in self.sizes:
                buf = (C.c_char * self.shm[k].size).from_buffer(self.shm[k].buf)
                addr = C.addressof(buf)
                if lib.at_host_register(addr, self.shm[k].size) == 0:
                    self.registered.append((addr, buf))
        self.arr = {k: np.ndarray((self.sizes[k] // (8 if k in ("cigar_off", "nops") else 4),),
                                  dtype=np.uint64 if k in ("cigar_off", "nops") else (np.int32 if k == "score" else np.uint32),
                                  buffer=self.shm[k].buf) for k in self.sizes}

    def slice_out(self, lo, hi):
        """BatchResult whose arrays are views into the shared outputs at this rank's slice."""
        r = BatchResult(0)
        r.n = hi - lo
        for k in ("score", "end_i", "end_j", "beg_i", "beg_j"):
            setattr(r, k, self.arr[k][lo:hi])
        base = lo + self.rank                      # every rank's offsets take n_slice + 1 entries
        r.cigar_off = self.arr["cigar_off"][base:base + (hi - lo) + 1]
        r.cigar = self.arr["cigar"][self.rank * self.cap:(self.rank + 1) * self.cap]
        return r

    def merge(self, cut, dst_cigar, dst_off):
        """rank 0: dense CIGARs of the whole batch in pair order."""
        base = 0
        for r in range(self.world):
            lo, hi = int(cut[r]), int(cut[r + 1])
            offs = self.arr["cigar_off"][lo + r:lo + r + (hi - lo) + 1]
            k = int(offs[-1]) if hi > lo else 0
            dst_off[lo:hi] = offs[:-1] + np.uint64(base)
            dst_cigar[base:base + k] = self.arr["cigar"][r * self.cap:r * self.cap + k]
            base += k
        dst_off[self.n] = base
        return base

    def close(self):
        self.arr = None
        if self.registered:
            from . import load_library
            lib = load_library()
            for addr, buf in self.registered:
                lib.at_host_unregister(addr)
            self.registered = []
        for s in self.shm.values():
            if self.rank == 0:      # unlink first: it works while views of the mapping are still alive, close() does not
                try:
                    s.unlink()
                except Exception:
                    pass
            try:
                s.close()
            except Exception:
                pass
