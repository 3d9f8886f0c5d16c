#!/usr/bin/env python
"""Smallest program that launches one BASELINE shape's kernels a few times -- the command line ncu wraps
(`ncu --set full -k regex:... python tools/prof_run.py c2 --pairs 1048576 --reps 2`) after it has exited 0 unprofiled."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aligntools.c_b200 as A  # noqa: E402
from aligntools.c_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c1", "c2", "c3", "c4", "c5", "global", "fit"])
    ap.add_argument("--pairs", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--twobit", action="store_true")
    args = ap.parse_args()
    mk = {"c1": lambda n: synth.config1_global(n or 1), "c2": lambda n: synth.config2_local(n_pairs=n or (1 << 20)),
          "c3": lambda n: synth.config3_fit_jump(n_pairs=n or 256), "c4": lambda n: synth.config4_overlap(n_pairs=n or 64),
          "c5": lambda n: synth.config5_edit(n_pairs=n or 8), "global": lambda n: synth.global_short(n_pairs=n or 65536),
          "fit": lambda n: dict(synth.global_short(n_pairs=n or 65536, l1=150, l2=400), mode="fit",
                                params=dict(m=1, u=-2, o=-5, e=-1, j=-10, jump=False))}
    w = mk[args.config](args.pairs)
    al = A.Aligner()
    enc = A.SEQ_BYTES
    q, qo, t, to = w["q"], w["q_off"], w["t"], w["t_off"]
    if args.twobit:
        enc = A.SEQ_2BIT
        q, qo, _ = A.pack_2bit(w["q"], w["q_off"], w["q_len"])
        t, to, _ = A.pack_2bit(w["t"], w["t_off"], w["t_len"])
    b = al.batch(w["mode"], A.Opt(**w["params"]), q, qo, w["q_len"], t, to, w["t_len"], sites=w["sites"], site_off=w["site_off"],
                 out_flags=0 if w["mode"] == "edit" else A.OUT_CIGAR, encoding=enc)
    for _ in range(args.reps):
        tm = b.run()
    res = b.fetch()
    b.free()
    print(json.dumps({"config": args.config, "pairs": int(len(w["q_len"])), "cells": int(tm.cells), "fill_ms": tm.fill_ms, "traceback_ms": tm.traceback_ms,
                      "kernel_ms": tm.fill_kernel_ms, "gcups": tm.cells / (tm.device_ms * 1e-3) / 1e9, "kind": tm.fill_kernel_kind, "rows": tm.fill_kernel_rows,
                      "score_sum": int(res.score.astype("int64").sum())}))
    al.close()


if __name__ == "__main__":
    main()
