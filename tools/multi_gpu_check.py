#!/usr/bin/env python
"""One process, several GPUs: at_create(devices[]) shards a ragged batch over the devices
(contiguous slices balanced by cells) -- check against the oracle and report the timing."""
import json, os, random, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aligntools.c_b200 as A
from aligntools.c_b200 import synth
import oracle
import torch

n_dev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
al = A.Aligner(devices=list(range(n_dev)))
out = {"devices": n_dev}
for name, w in (("local", synth.config2_local(n_pairs=40000)), ("overlap", synth.config4_overlap(n_pairs=12, lo=2000, hi=5000)),
                ("fitjump", synth.config3_fit_jump(n_pairs=8))):
    opt = A.Opt(**w["params"])
    for oneshot in (False, True):
        if oneshot:
            res = al.align_arrays(w["mode"], opt, w["q"], w["q_off"], w["q_len"], w["t"], w["t_off"], w["t_len"], sites=w["sites"], site_off=w["site_off"], out_flags=3)
            tm = res.timing
        else:
            b = al.batch(w["mode"], opt, w["q"], w["q_off"], w["q_len"], w["t"], w["t_off"], w["t_len"], sites=w["sites"], site_off=w["site_off"], out_flags=3)
            tm = b.run(); res = b.fetch(); b.free()
        p = oracle.Params(**{k: w["params"][k] for k in "muoej"}, jump=w["params"]["jump"])
        n = len(w["q_len"])
        ref = oracle.port_batch(w["mode"], p, w["q"], np.append(w["q_off"], 0).astype(np.uint64), w["q_len"], w["t"], np.append(w["t_off"], 0).astype(np.uint64), w["t_len"],
                                w["sites"], w["site_off"], want_aln=True, threads=16)
        ok = bool(np.array_equal(res.score.astype(np.int64), ref.score)) and all(res.aln(k) == ref.aln(k) for k in range(0, n, max(1, n // 500)))
        out[f"{name}{'_oneshot' if oneshot else ''}"] = {"pairs": n, "bit_exact": ok, "device_ms": tm.device_ms, "gcups": tm.cells / tm.device_ms / 1e6}
        assert ok, name
al.close()
print(json.dumps(out))
