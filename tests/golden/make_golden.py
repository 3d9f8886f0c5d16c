#!/usr/bin/env python
"""Regenerate the golden fixtures from the UNMODIFIED reference (run in the build
container only; needs /root/reference and `make -C oracle`).

Writes
  tests/golden/cli_vectors.json   -- the 24 (+extra) CLI commands of SURVEY.md Appendix B
                                     over the reference's test/*.fa: inputs (records as
                                     parsed by kseq), full stdout, md5, rc
  tests/golden/fuzz_vectors.json  -- random small pairs, all six mode variants and random
                                     scoring parameters, answered by the compiled
                                     reference in-process (score, r1, r2)
Nothing under /root/reference is copied verbatim: the FASTA records are stored as JSON
fields so the GPU box (which has no /root/reference) can rebuild equivalent inputs.
"""
import gzip
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

REF_TEST = "/root/reference/test"

CLI = [  # (id, argv after the binary; "$T/x.fa" is substituted)
    ("B1", ["global", "-m", "1", "-u", "-1", "-o", "-4", "-e", "-1", "$T/test_global.fa"]),
    ("B2", ["global", "$T/test_global.fa"]),
    ("B3", ["global", "$T/test_local.fa"]),
    ("B4", ["global", "$T/test_edit.fa"]),
    ("B5", ["local", "-m", "2", "-u", "-2", "-o", "-5", "-e", "-2", "$T/test_local.fa"]),
    ("B6", ["local", "$T/test_local.fa"]),
    ("B7", ["local", "$T/test_global.fa"]),
    ("B8", ["local", "$T/test_edit.fa"]),
    ("B9", ["fit", "-m", "2", "-u", "-2", "-s", "$T/test_fit.fa"]),
    ("B10", ["fit", "-s", "$T/test_fit.fa"]),
    ("B11", ["fit", "$T/test_fit.fa"]),
    ("B12", ["fit", "-m", "2", "-u", "-2", "$T/test_fit.fa"]),
    ("B13", ["fit", "$T/test_global.fa"]),
    ("B14", ["fit", "$T/test_edit.fa"]),
    ("B15", ["overlap", "$T/test_global.fa"]),
    ("B16", ["overlap", "$T/test_local.fa"]),
    ("B17", ["overlap", "$T/test_edit.fa"]),
    ("B18", ["overlap", "$T/test_fit.fa"]),
    ("B19", ["edit", "-u", "1", "-o", "2", "$T/test_edit.fa"]),
    ("B20", ["edit", "$T/test_edit.fa"]),
    ("B21", ["edit", "-u", "1", "$T/test_global.fa"]),
    ("B22", ["edit", "-u", "1", "$T/test_local.fa"]),
    ("B23", ["edit", "-u", "1", "$T/test_fit.fa"]),
    ("B24a", ["fit", "$T/tmp.fa"]),
    ("B24b", ["fit", "-s", "$T/tmp.fa"]),
    # extra: CLI contract rows of SURVEY A.5
    ("X1", ["fit", "-s", "-j", "-10", "$T/test_fit.fa"]),
    ("X2", ["global", "-s", "$T/test_global.fa"]),
    ("X3", ["overlap", "-m", "2", "$T/test_global.fa"]),
    ("X4", ["global"]),
    ("X5", ["bogus"]),
    ("X6", []),
    ("X7", ["fit", "$T/test_local.fa"]),       # l1 > l2 -> FATAL
    ("X8", ["local", "-m", "3", "-u", "-1", "-o", "-2", "-e", "-2", "$T/test_fit.fa"]),
    ("X9", ["global", "-o", "0", "-e", "0", "$T/test_local.fa"]),
    ("X10", ["edit", "-u", "3", "$T/test_local.fa"]),
]


def parse_fasta(path):
    """Minimal FASTA reader matching kseq's name / comment / sequence split."""
    recs = []
    op = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    with op(path, "rt") as f:
        name = comment = None
        seq = []
        for line in f:
            line = line.rstrip("\n").rstrip("\r")
            if line.startswith(">"):
                if name is not None:
                    recs.append({"name": name, "comment": comment, "seq": "".join(seq)})
                hdr = line[1:]
                parts = hdr.split(None, 1)
                name = parts[0] if parts else ""
                comment = parts[1] if len(parts) > 1 else None
                seq = []
            else:
                seq.append(line.strip())
        if name is not None:
            recs.append({"name": name, "comment": comment, "seq": "".join(seq)})
    return recs


def main():
    if not oracle.have_ref():
        oracle.build(quiet=False)
    files = {}
    for fn in sorted(os.listdir(REF_TEST)):
        if fn.endswith(".fa"):
            files[fn] = {"records": parse_fasta(os.path.join(REF_TEST, fn)),
                         "md5": hashlib.md5(open(os.path.join(REF_TEST, fn), "rb").read()).hexdigest()}
    vectors = []
    for vid, argv in CLI:
        real = [a.replace("$T", REF_TEST) for a in argv]
        rc, out, err = oracle.run_ref_cli(real)
        small = len(out) <= 4096
        vectors.append({
            "id": vid, "argv": argv, "rc": rc,
            "stdout_md5": hashlib.md5(out).hexdigest(), "stdout_len": len(out),
            "stdout": out.decode("latin-1"),
            "stderr": err.decode("latin-1").replace(REF_TEST, "$T").replace(oracle.REF_CLI, "$BIN"),
        })
        print(vid, rc, vectors[-1]["stdout_md5"], out[:40])
    with open(os.path.join(HERE, "cli_vectors.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "files": files, "vectors": vectors}, f, indent=0)

    # ---- fuzz fixtures answered by the in-process reference ----
    rng = random.Random(20261018)
    fuzz = []
    alph_sets = ["ACGT", "ACGT", "AC", "ACGTacgtN", "ARNDCQEGHILKMFPSTWYV"]

    def rseq(n, alph):
        return "".join(rng.choice(alph) for _ in range(n))

    def mutate(s, alph, rate):
        out = []
        for ch in s:
            r = rng.random()
            if r < rate:
                out.append(rng.choice(alph))
            elif r < rate * 1.5:
                continue
            elif r < rate * 2:
                out.append(ch)
                out.append(rng.choice(alph))
            else:
                out.append(ch)
        return "".join(out)

    variants = ["global", "local", "fit", "fitjump", "overlap", "edit"]
    for case in range(900):
        var = variants[case % 6]
        alph = rng.choice(alph_sets)
        wild = case % 5 == 4   # sign-flipped / zero parameters
        if wild:
            p = oracle.Params(m=rng.randint(-2, 5), u=rng.randint(-5, 2), o=rng.randint(-8, 2),
                              e=rng.randint(-4, 2), j=rng.randint(-12, 2))
        else:
            p = oracle.Params(m=rng.randint(1, 5), u=rng.randint(-5, 0), o=rng.randint(-8, 0),
                              e=rng.randint(-4, 0), j=rng.randint(-12, 0))
        l1 = rng.randint(1, 70) if case % 7 else rng.randint(60, 300)
        core = rseq(l1, alph)
        kind = rng.random()
        if kind < 0.25:
            s1, s2 = core, rseq(rng.randint(max(2, l1), l1 + 120), alph)
        else:
            s1 = core
            mid = mutate(core, alph, rng.choice([0.0, 0.05, 0.15, 0.3]))
            if var == "fitjump" and len(mid) > 6:
                cut = rng.randint(2, len(mid) - 2)
                mid = mid[:cut] + rseq(rng.randint(5, 60), alph.lower() if alph.isupper() else alph) + mid[cut:]
            s2 = rseq(rng.randint(0, 40), alph) + mid + rseq(rng.randint(0, 40), alph)
        if not s2:
            s2 = rseq(3, alph)
        if var in ("fit", "fitjump") and len(s1) > len(s2):
            s1, s2 = s2, s1
        if var in ("fit", "fitjump") and len(s2) < 2:
            s2 = s2 + rseq(2, alph)
        sites = None
        mode = var
        if var == "fitjump":
            mode = "fit"
            p.jump = True
            k = rng.choice([0, 1, 2, 4, 8, len(s2)])
            if k == len(s2):
                sites = list(range(len(s2)))
            else:
                sites = sorted(rng.randrange(0, len(s2) + 3) for _ in range(k))
        r = oracle.ref_align(mode, s1.encode(), s2.encode(), p, sites)
        fuzz.append({"mode": mode, "m": p.m, "u": p.u, "o": p.o, "e": p.e, "j": p.j, "jump": int(p.jump),
                     "sites": sites, "s1": s1, "s2": s2, "score": r.score,
                     "r1": r.r1.decode("latin-1"), "r2": r.r2.decode("latin-1")})
    with open(os.path.join(HERE, "fuzz_vectors.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "cases": fuzz}, f, indent=0)
    print("fuzz cases:", len(fuzz))


if __name__ == "__main__":
    main()
