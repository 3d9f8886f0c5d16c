#!/usr/bin/env python
"""First-contact GPU check: every mode on small inputs against the oracle port, printing all
mismatches instead of stopping at the first (used while bringing kernels up)."""
import os
import random
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aligntools.c_b200 as A  # noqa: E402
import oracle  # noqa: E402
from helpers import pack_batch  # noqa: E402


def main():
    al = A.Aligner()
    rng = random.Random(3)
    total_bad = 0
    for mode in ["local", "global", "fit", "fitjump", "overlap", "edit"]:
        for (lo, hi, n) in [(1, 32, 40), (33, 160, 40), (161, 256, 20), (257, 700, 10)]:
            q, t = [], []
            for _ in range(n):
                l1 = rng.randint(lo, hi)
                s1 = bytes(rng.choice(b"ACGT") for _ in range(l1))
                s2 = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 50))) + \
                    bytes(c if rng.random() > 0.1 else rng.choice(b"ACGT") for c in s1) + \
                    bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 60)))
                q.append(s1); t.append(s2)
            prm = dict(m=2, u=-3, o=-4, e=-1, j=-7, jump=(mode == "fitjump"))
            md = "fit" if mode == "fitjump" else mode
            sites = None
            if mode == "fitjump":
                sites = [sorted(rng.randrange(len(x)) for _ in range(rng.choice([0, 2, 5]))) for x in t]
            try:
                res = al.align(md, q, t, A.Opt(**prm), sites=sites, out_flags=0 if md == "edit" else 3)
            except Exception:
                traceback.print_exc()
                total_bad += 1
                continue
            bad = 0
            for k in range(n):
                p = oracle.Params(prm["m"], prm["u"], prm["o"], prm["e"], prm["j"], prm["jump"])
                o = oracle.port_align(md, q[k], t[k], p, sites[k] if sites else None)
                ok = int(res.score[k]) == o.score
                if ok and md != "edit":
                    ok = res.aln(k) == (o.r1, o.r2)
                if not ok:
                    bad += 1
                    if bad <= 2:
                        print(f"  MISMATCH {mode} l1={len(q[k])} l2={len(t[k])}: gpu score {int(res.score[k])} vs {o.score}; "
                              f"end=({int(res.end_i[k])},{int(res.end_j[k])}) vs {o.coords[:2]}")
                        if md != "edit":
                            print("   gpu:", res.aln(k)[0][:70], "\n       ", res.aln(k)[1][:70])
                            print("   ref:", o.r1[:70], "\n       ", o.r2[:70])
            print(f"{mode:8s} l1 in [{lo},{hi}] n={n}: {'OK' if not bad else str(bad) + ' BAD'}", flush=True)
            total_bad += bad
    print("TOTAL BAD:", total_bad)
    return 1 if total_bad else 0


if __name__ == "__main__":
    sys.exit(main())
