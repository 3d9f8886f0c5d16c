#!/usr/bin/env python
"""How long does the first large pointer-arena allocation take? (at_batch_create on a fresh handle)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aligntools.c_b200 as A
from aligntools.c_b200 import synth
al = A.Aligner()
for n in (int(x) for x in sys.argv[1:] or ["512", "2048", "3000", "3000"]):
    w = synth.config3_fit_jump(n_pairs=n)
    t0 = time.perf_counter()
    b = al.batch("fit", A.Opt(**w["params"]), w["q"], w["q_off"], w["q_len"], w["t"], w["t_off"], w["t_len"], sites=w["sites"], site_off=w["site_off"], out_flags=1)
    t1 = time.perf_counter()
    tm = b.run()
    t2 = time.perf_counter()
    b.free()
    print(f"pairs {n}: create {t1 - t0:.3f} s (arena {tm.ptr_bytes / 1e9:.1f} GB), run {t2 - t1:.3f} s, device {tm.device_ms:.1f} ms", flush=True)
