#!/usr/bin/env python
"""Single-stripe speed of the bit-parallel edit kernel: reads of <= 1024 rows against long targets
(one task per pair, no stripe hand-off); device time / columns = time per column step of one warp."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aligntools.c_b200 as A
al = A.Aligner()
rng = np.random.default_rng(1)
for n, l1, l2 in ((148, 1000, 100000), (592, 1000, 100000), (2368, 1000, 100000), (148, 4000, 100000)):
    q = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=n * l1)]
    t = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=n * l2)]
    qo = (np.arange(n, dtype=np.uint64) * l1); to = (np.arange(n, dtype=np.uint64) * l2)
    ql = np.full(n, l1, np.uint32); tl = np.full(n, l2, np.uint32)
    b = al.batch("edit", A.Opt(u=1), q, qo, ql, t, to, tl, out_flags=0)
    b.run(); tm = b.run(); b.free()
    print(f"pairs {n} {l1}x{l2}: {tm.device_ms:.3f} ms, {tm.device_ms * 1e6 / l2:.1f} ns per column step, {tm.cells / tm.device_ms / 1e6:.0f} GCUPS", flush=True)
