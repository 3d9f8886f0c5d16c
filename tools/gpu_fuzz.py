#!/usr/bin/env python
"""Randomised GPU-vs-oracle stress: random modes, scoring parameters (also zero / sign-flipped), alphabets,
batch shapes (K1 classes, packed pairs, K2 stripes, bit-parallel edit) and both entry paths; every pair is
compared with the oracle port (score, cells, alignment strings, CIGAR).  Runs for --seconds."""
import argparse, json, os, random, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import aligntools.c_b200 as A
import oracle
from helpers import pack_batch

ALPHABETS = [b"ACGT", b"AC", b"ACGTN", b"ACGTNRYK", b"ACDEFGHIKLMNPQRSTVWY", b"A"]


def rle(ops):
    out, k = [], 0
    while k < len(ops):
        j = k
        while j < len(ops) and ops[j] == ops[k]:
            j += 1
        out.append(f"{j - k}{chr(ops[k])}"); k = j
    return "".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--one-shot", type=float, default=0.3, help="share of rounds that go through the pipelined at_batch_align")
    args = ap.parse_args()
    rng = random.Random(args.seed)
    al = A.Aligner()
    t_end = time.time() + args.seconds
    rounds = pairs = bad = 0
    while time.time() < t_end:
        mode = rng.choice(["global", "local", "fit", "fitjump", "overlap", "edit"])
        alpha = rng.choice(ALPHABETS + [b"ACGT"] * 3)
        shape = rng.choice(["short", "short", "uniform", "uniform", "mixed", "long", "thin"])
        n = {"short": rng.randint(1, 400), "uniform": rng.randint(1, 300), "mixed": rng.randint(1, 120), "long": rng.randint(1, 12), "thin": rng.randint(1, 30)}[shape]
        l1_fixed = rng.randint(1, 256)
        same_l2 = rng.random() < 0.5
        l2_fixed = rng.randint(1, 400)
        q, t = [], []
        for _ in range(n):
            if shape == "short": l1 = rng.randint(1, 256)
            elif shape == "uniform": l1 = l1_fixed
            elif shape == "mixed": l1 = rng.choice([rng.randint(1, 256), rng.randint(257, 1200)])
            elif shape == "long": l1 = rng.randint(257, 2600)
            else: l1 = rng.randint(1, 3000)
            s1 = bytes(rng.choice(alpha) for _ in range(l1))
            if shape == "thin":
                s2 = bytes(rng.choice(alpha) for _ in range(rng.randint(1, 40)))
            else:
                core = bytes(c if rng.random() > 0.12 else rng.choice(alpha) for c in s1)
                if rng.random() < 0.3: core = core[: rng.randint(0, len(core))]
                s2 = bytes(rng.choice(alpha) for _ in range(rng.randint(0, 60))) + core + bytes(rng.choice(alpha) for _ in range(rng.randint(0, 200)))
                if (same_l2 and shape == "short") or shape == "uniform":      # uniform: one shape for the whole batch (packed global / fit lanes, device-side plan)
                    lf = max(l2_fixed, l1_fixed) if shape == "uniform" else l2_fixed
                    s2 = (s2 + bytes(rng.choice(alpha) for _ in range(lf)))[:lf]
            if not s2: s2 = bytes([rng.choice(alpha)])
            if mode.startswith("fit"):
                if len(s1) > len(s2) and shape != "uniform": s1, s2 = s2, s1
                if len(s2) < 2: s2 = s2 + bytes([rng.choice(alpha)])
            q.append(s1); t.append(s2)
        if rng.random() < 0.25:
            prm = dict(m=rng.randint(-2, 5), u=rng.randint(-5, 2), o=rng.randint(-8, 2), e=rng.randint(-4, 2), j=rng.randint(-12, 2))
        else:
            prm = dict(m=rng.randint(1, 5), u=rng.randint(-5, 0), o=rng.randint(-8, 0), e=rng.randint(-4, 0), j=rng.randint(-12, 0))
        if mode == "edit" and rng.random() < 0.6: prm["u"] = 1
        prm["jump"] = mode == "fitjump"
        white = mode == "fitjump" and rng.random() < 0.3      # junction list as a whitelist (at_params.jump == 2)
        md = "fit" if mode == "fitjump" else mode
        sites = site_off = None
        if mode == "fitjump":
            ss, so = [], [0]
            for s2 in t:
                ss += sorted(rng.randrange(len(s2)) for _ in range(rng.choice([0, 0, 1, 4, 12]))); so.append(len(ss))
            sites = np.array(ss + [0], dtype=np.int32); site_off = np.array(so, dtype=np.uint64)
        qb, qo, ql = pack_batch(q); tb, to, tl = pack_batch(t)
        opt = A.Opt(**prm, whitelist=white)
        p = oracle.Params(prm["m"], prm["u"], prm["o"], prm["e"], prm["j"], 2 if white else prm["jump"])
        enc, qd, qod, td, tod = A.SEQ_BYTES, qb, qo[:-1].copy(), tb, to[:-1].copy()
        if alpha == b"ACGT" and rng.random() < 0.5:             # 2-bit input (stays packed in HBM when the whole batch runs on K1)
            enc = A.SEQ_2BIT
            al2 = rng.choice([1, 16])
            qd, qod, _ = A.pack_2bit(qb, qo[:-1].copy(), ql, align=al2)
            td, tod, _ = A.pack_2bit(tb, to[:-1].copy(), tl, align=al2)
        ref = oracle.port_batch(md, p, qb, qo, ql, tb, to, tl, sites, site_off, want_aln=(md != "edit"), want_ops=(md != "edit"), threads=16)
        flags = 0 if md == "edit" else 3
        if rng.random() < args.one_shot:
            os.environ["AT_PIPE_MIN_CELLS"] = "1"; os.environ["AT_PIPE_SLICE_CELLS"] = str(rng.choice([20000, 300000, 5000000]))
            os.environ["AT_PIPE_MIN_TASKS"] = "1"; os.environ["AT_ONE_SLICE_MIN_CELLS"] = str(rng.choice([1, 1 << 40]))
            if rng.random() < 0.3: os.environ["AT_PTR_BUDGET_MB"] = "48"       # several chunks per sub-slice
            else: os.environ.pop("AT_PTR_BUDGET_MB", None)
            res = al.align_arrays(md, opt, qd, qod, ql, td, tod, tl, sites=sites, site_off=site_off, out_flags=flags, encoding=enc)
        else:
            os.environ.pop("AT_PTR_BUDGET_MB", None)
            b = al.batch(md, opt, qd, qod, ql, td, tod, tl, sites=sites, site_off=site_off, out_flags=flags, encoding=enc)
            b.run(); res = b.fetch(); b.free()
        ok = np.array_equal(res.score.astype(np.int64), ref.score)
        if ok and md != "edit":
            ok = np.array_equal(res.end_i, ref.coords[:, 0].astype(np.uint32)) and np.array_equal(res.end_j, ref.coords[:, 1].astype(np.uint32))
            for k in range(n):
                if not ok: break
                ok = res.aln(k) == ref.aln(k) and res.cigar_string(k) == rle(ref.op(k))
        rounds += 1; pairs += n
        if not ok:
            bad += 1
            print("MISMATCH", json.dumps(dict(mode=mode, prm=prm, shape=shape, n=n, alpha=alpha.decode(), seed=args.seed, round=rounds, enc=enc, white=white)), flush=True)
            if bad >= 5: break
    print(json.dumps({"rounds": rounds, "pairs": pairs, "mismatching_rounds": bad, "seconds": args.seconds, "seed": args.seed}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
