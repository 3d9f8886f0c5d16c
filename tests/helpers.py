"""Shared helpers for the parity tests (golden loaders, reference stdout formatting)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def load_cli():
    with open(os.path.join(GOLD, "cli_vectors.json")) as f:
        return json.load(f)


def load_fuzz():
    with open(os.path.join(GOLD, "fuzz_vectors.json")) as f:
        return json.load(f)["cases"]


def parse_cli_argv(argv):
    """Mirror of the reference's getopt loops (src/alignment.h:323, 481, 703, 856, 971):
    returns (mode, params dict, file name) for the well-formed golden commands."""
    mode = argv[0]
    prm = dict(m=1, u=-2, o=-5, e=-1, j=-10, jump=False)
    i = 1
    fname = None
    while i < len(argv):
        a = argv[i]
        if a == "-s":
            prm["jump"] = True
            i += 1
        elif a in ("-m", "-u", "-o", "-e", "-j"):
            prm[a[1]] = int(argv[i + 1])
            i += 2
        else:
            fname = a
            i += 1
    if mode == "overlap":          # main_overlap reads argv[1] (:994): options unusable
        prm = dict(m=1, u=-2, o=-5, e=-1, j=-10, jump=False)
    return mode, prm, fname


def expected_stdout(mode, prm, comment, score, r1, r2):
    """The reference's stdout for one run."""
    out = b""
    if mode == "edit":
        return b"edit_distance=%d\n" % score
    if mode == "fit":
        if prm["jump"]:
            out += comment.encode("latin-1") + b"\n"
        out += b"asDAsdaSDAsdasDAsdaSD\n"
    if mode == "overlap":
        out += b"%d.000000\n" % score
    else:
        out += b"score=%d.000000\n" % score
    return out + r1 + b"\n" + r2 + b"\n"


def sites_from_comment(comment):
    """kstring_read (:245-253): split on '|', atoi each field."""
    out = []
    for tok in comment.split("|"):
        tok = tok.strip()
        num = ""
        for k, ch in enumerate(tok):
            if ch.isdigit() or (k == 0 and ch in "+-"):
                num += ch
            else:
                break
        try:
            out.append(int(num))
        except ValueError:
            out.append(0)
    return out


def pack_batch(seqs):
    """list[bytes] -> (uint8 concat, uint64 offsets[n+1], uint32 lens[n])."""
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
    if buf.size == 0:
        buf = np.zeros(1, np.uint8)
    return buf, off, lens
