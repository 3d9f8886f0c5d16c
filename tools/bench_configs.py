#!/usr/bin/env python
"""Secondary measurements: device-timed GCUPS of the other BASELINE.json shapes (C3 fit+jump,
C4 overlap, C5 edit, plus global) at reduced pair counts.  Not the headline bench (bench.py = C2);
results go to profiles/ as context for DESIGN.md."""
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


def run(al, name, w, flags, reps=2):
    opt = A.Opt(**w["params"])
    t0 = time.perf_counter()
    b = al.batch(w["mode"], opt, w["q"], w["q_off"], w["q_len"], w["t"], w["t_off"], w["t_len"],
                 sites=w["sites"], site_off=w["site_off"], out_flags=flags)
    t_create = time.perf_counter() - t0
    best = None
    for _ in range(reps):
        tm = b.run()
        if best is None or tm.device_ms < best.device_ms:
            best = tm
    res = b.fetch()
    b.free()
    out = {"config": name, "pairs": int(len(w["q_len"])), "cells": int(best.cells), "fill_ms": best.fill_ms,
           "traceback_ms": best.traceback_ms, "device_ms": best.device_ms,
           "gcups": best.cells / (best.device_ms * 1e-3) / 1e9, "fill_gcups": best.cells / (best.fill_ms * 1e-3) / 1e9,
           "ptr_GB": best.ptr_bytes / 1e9, "launches": int(best.launches), "create_s": t_create,
           "score_sum": int(res.score.astype(np.int64).sum())}
    print(json.dumps(out), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3", type=int, default=1024)
    ap.add_argument("--c4", type=int, default=256)
    ap.add_argument("--c5", type=int, default=16)
    ap.add_argument("--c5len", type=int, default=100000)
    ap.add_argument("--glob", type=int, default=65536)
    args = ap.parse_args()
    al = A.Aligner()
    if args.glob:
        run(al, "global 150x150 -m1 -u-1 -o-4 -e-1, score+CIGAR", synth.global_short(n_pairs=args.glob), A.OUT_CIGAR)
    if args.glob:
        w = synth.global_short(n_pairs=args.glob, l1=150, l2=400)
        w["mode"] = "fit"; w["params"] = dict(m=1, u=-2, o=-5, e=-1, j=-10, jump=False)
        run(al, "fit 150x400 (defaults), score+CIGAR", w, A.OUT_CIGAR)
    if args.c3:
        run(al, "C3 fit -s -j -10, 2k x 20k, score+CIGAR", synth.config3_fit_jump(n_pairs=args.c3), A.OUT_CIGAR)
    if args.c4:
        run(al, "C4 overlap 10-20 kbp, score+CIGAR", synth.config4_overlap(n_pairs=args.c4), A.OUT_CIGAR)
    if args.c5:
        run(al, f"C5 edit -u 1, {args.c5len} x {args.c5len}", synth.config5_edit(n_pairs=args.c5, length=args.c5len), 0, reps=1)


if __name__ == "__main__":
    main()
