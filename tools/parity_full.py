#!/usr/bin/env python
"""Full-size parity run (BASELINE.json north star: "bit-exact ... on 1M synthetic pairs").

C2: every pair of the 1 Mi-pair local batch -- score, end cell, traceback start cell and every
alignment column (the GPU's CIGAR expanded to columns against the oracle port's op string) --
slice by slice, the port running on all host cores.  C3 / C4 / C5 at their full shapes on a few
pairs each.  The oracle is only the checker here.  Writes one JSON document (default
gpurun_out/parity_full.json; copy into profiles/ to keep it)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aligntools.c_b200 as A  # noqa: E402
from aligntools.c_b200 import synth  # noqa: E402
import oracle  # noqa: E402

CODE = np.frombuffer(b"MIDN", dtype=np.uint8)


def gpu_columns(res, n):
    """GPU CIGARs -> (columns per pair, dense op letters of all pairs in pair order)."""
    tot = int(res.cigar_off[n])
    ops = res.cigar[:tot]
    runs = (ops >> 4).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(runs)])
    ncols = csum[res.cigar_off[1:n + 1].astype(np.int64)] - csum[res.cigar_off[:n].astype(np.int64)]
    return ncols, np.repeat(CODE[ops & 3], runs)


def port_columns(ref, n):
    lens = ref.aln_len.astype(np.int64)
    start = ref.aln_off[:n].astype(np.int64)
    dense_start = np.concatenate([[0], np.cumsum(lens)[:-1]])
    idx = np.repeat(start - dense_start, lens) + np.arange(int(lens.sum()), dtype=np.int64)
    return lens, ref.ops[idx]


def check(al, w, threads, with_cols=True, twobit=False):
    mode = w["mode"]
    prm = w["params"]
    n = len(w["q_len"])
    opt = A.Opt(**prm)
    flags = 0 if mode == "edit" else A.OUT_CIGAR
    if twobit:      # the layout bench.py's e2e leg hands over: 2-bit codes, records on 16-byte boundaries
        q2, qo2, _ = A.pack_2bit(w["q"], w["q_off"], w["q_len"], align=16)
        t2, to2, _ = A.pack_2bit(w["t"], w["t_off"], w["t_len"], align=16)
        b = al.batch(mode, opt, q2, qo2, w["q_len"], t2, to2, w["t_len"], sites=w["sites"], site_off=w["site_off"],
                     out_flags=flags, encoding=A.SEQ_2BIT)
    else:
        b = al.batch(mode, opt, w["q"], w["q_off"], w["q_len"], w["t"], w["t_off"], w["t_len"],
                     sites=w["sites"], site_off=w["site_off"], out_flags=flags)
    tm = b.run()
    res = b.fetch()
    b.free()
    p = oracle.Params(prm["m"], prm["u"], prm["o"], prm["e"], prm["j"], prm["jump"])
    t0 = time.perf_counter()
    ref = oracle.port_batch(mode, p, w["q"], np.append(w["q_off"], 0).astype(np.uint64), w["q_len"], w["t"],
                            np.append(w["t_off"], 0).astype(np.uint64), w["t_len"], w["sites"], w["site_off"],
                            want_aln=(mode != "edit"), want_ops=(mode != "edit"), threads=threads)
    cpu_s = time.perf_counter() - t0
    out = {"pairs": n, "cells": int(tm.cells), "gpu_device_ms": tm.device_ms, "cpu_port_s": cpu_s,
           "score_mismatch": int((res.score.astype(np.int64) != ref.score).sum())}
    if mode != "edit":
        out["end_cell_mismatch"] = int(((res.end_i != ref.coords[:, 0].astype(np.uint32)) | (res.end_j != ref.coords[:, 1].astype(np.uint32))).sum())
        out["begin_cell_mismatch"] = int(((res.beg_i != ref.coords[:, 2].astype(np.uint32)) | (res.beg_j != ref.coords[:, 3].astype(np.uint32))).sum())
        if with_cols:
            gl, gc = gpu_columns(res, n)
            pl, pc = port_columns(ref, n)
            out["alignment_length_mismatch"] = int((gl != pl).sum())
            out["alignment_columns"] = int(pl.sum())
            out["alignment_column_mismatch"] = int((gc != pc).sum()) if gc.size == pc.size else -1
    return out


def check_bulk(al, w, threads, n_sample, seed=1):
    """A batch too large to re-run on the CPU (several pointer-arena chunks): run it whole on the GPU,
    then re-run a random sample of its pairs through the port and compare score, cells and columns."""
    mode, prm = w["mode"], w["params"]
    n = len(w["q_len"])
    b = al.batch(mode, A.Opt(**prm), w["q"], w["q_off"], w["q_len"], w["t"], w["t_off"], w["t_len"],
                 sites=w["sites"], site_off=w["site_off"], out_flags=A.OUT_CIGAR)
    tm = b.run()
    res = b.fetch()
    b.free()
    rng = np.random.default_rng(seed)
    pick = np.unique(np.concatenate([[0, n - 1], rng.integers(0, n, size=n_sample)]))
    p = oracle.Params(prm["m"], prm["u"], prm["o"], prm["e"], prm["j"], prm["jump"])
    sites = site_off = None
    if w["sites"] is not None:
        so = [0]
        ss = []
        for k in pick:
            ss.append(w["sites"][int(w["site_off"][k]):int(w["site_off"][k + 1])]); so.append(so[-1] + len(ss[-1]))
        sites = np.concatenate(ss + [np.zeros(1, np.int32)]).astype(np.int32); site_off = np.array(so, dtype=np.uint64)
    ref = oracle.port_batch(mode, p, w["q"], np.append(w["q_off"][pick], 0).astype(np.uint64), np.ascontiguousarray(w["q_len"][pick]), w["t"],
                            np.append(w["t_off"][pick], 0).astype(np.uint64), np.ascontiguousarray(w["t_len"][pick]), sites, site_off,
                            want_aln=True, want_ops=True, threads=threads)
    bad = 0
    for x, k in enumerate(pick):
        ok = int(res.score[k]) == int(ref.score[x]) and int(res.end_i[k]) == int(ref.coords[x, 0]) and int(res.end_j[k]) == int(ref.coords[x, 1])
        ops = res.cigar_ops(int(k))
        cols = np.repeat(CODE[ops & 3], (ops >> 4).astype(np.int64)).tobytes()
        ok = ok and cols == ref.op(x)
        bad += 0 if ok else 1
    return {"pairs": n, "cells": int(tm.cells), "gpu_device_ms": tm.device_ms, "gcups": tm.cells / tm.device_ms / 1e6, "ptr_GB": tm.ptr_bytes / 1e9,
            "sampled_pairs": int(len(pick)), "sample_mismatch": bad}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-bulk", type=int, default=0, help="pairs of a multi-chunk C3 batch (sample-checked)")
    ap.add_argument("--c4-bulk", type=int, default=0)
    ap.add_argument("--bulk-sample", type=int, default=48)
    ap.add_argument("--pairs", type=int, default=1 << 20)
    ap.add_argument("--slice", type=int, default=1 << 17)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--max-seconds", type=float, default=900.0, help="stop starting new C2 slices after this long")
    ap.add_argument("--twobit", action="store_true", help="hand every other C2 slice over as 2-bit records (the HBM-resident packed path)")
    ap.add_argument("--c3", type=int, default=8)
    ap.add_argument("--c4", type=int, default=6)
    ap.add_argument("--c5", type=int, default=2)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_full.json"))
    args = ap.parse_args()
    oracle.build()
    al = A.Aligner()
    doc = {"host_threads": args.threads, "c2": {"pairs_requested": args.pairs, "slices": []}}
    t_start = time.perf_counter()
    w = synth.config2_local(n_pairs=args.pairs)
    agg = {}
    for lo in range(0, args.pairs, args.slice):
        if time.perf_counter() - t_start > args.max_seconds:
            break
        hi = min(args.pairs, lo + args.slice)
        sl = dict(w)
        for k in ("q_off", "q_len", "t_off", "t_len"):
            sl[k] = np.ascontiguousarray(w[k][lo:hi])
        r = check(al, sl, args.threads, twobit=args.twobit and (lo // args.slice) % 2 == 1)   # odd slices 2-bit, even slices bytes
        r["first_pair"] = lo
        r["twobit"] = bool(args.twobit and (lo // args.slice) % 2 == 1)
        doc["c2"]["slices"].append(r)
        for k, v in r.items():
            if k.endswith("mismatch") or k in ("pairs", "cells", "alignment_columns"):
                agg[k] = agg.get(k, 0) + v
        print(json.dumps(r), flush=True)
    doc["c2"]["total"] = agg
    doc["c2"]["workload"] = "C2 local 150x500 -m 2 -u -2 -o -5 -e -2, synth.config2_local stream 0"
    if args.c3:
        doc["c3"] = check(al, synth.config3_fit_jump(n_pairs=args.c3), args.threads)
        print("c3", json.dumps(doc["c3"]), flush=True)
    if args.c4:
        doc["c4"] = check(al, synth.config4_overlap(n_pairs=args.c4), args.threads)
        print("c4", json.dumps(doc["c4"]), flush=True)
    if args.c5:
        doc["c5"] = check(al, synth.config5_edit(n_pairs=args.c5), args.threads)
        print("c5", json.dumps(doc["c5"]), flush=True)
    if args.c3_bulk:
        doc["c3_bulk"] = check_bulk(al, synth.config3_fit_jump(n_pairs=args.c3_bulk, stream=7), args.threads, args.bulk_sample)
        print("c3_bulk", json.dumps(doc["c3_bulk"]), flush=True)
    if args.c4_bulk:
        doc["c4_bulk"] = check_bulk(al, synth.config4_overlap(n_pairs=args.c4_bulk, stream=7), args.threads, args.bulk_sample)
        print("c4_bulk", json.dumps(doc["c4_bulk"]), flush=True)
    al.close()
    bad = sum(v for k, v in agg.items() if k.endswith("mismatch"))
    for c in ("c3_bulk", "c4_bulk"):
        if c in doc:
            bad += doc[c]["sample_mismatch"]
    for c in ("c3", "c4", "c5"):
        if c in doc:
            bad += sum(v for k, v in doc[c].items() if k.endswith("mismatch"))
    doc["all_bit_exact"] = bad == 0
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(doc, f, indent=1)
    print("all_bit_exact", doc["all_bit_exact"])
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
