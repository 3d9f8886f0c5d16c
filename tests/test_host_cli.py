"""The C host (bin/alignTools, host/): the reference's command line over the C-ABI.

CPU part: the host's FASTA/FASTQ reader against what the reference's kseq parser makes of the
same bytes (tests/golden/fasta_cases.json; live against oracle/_ref when it is present), and every
golden CLI vector that ends before the alignment step (usage, option and input errors).
GPU part (-m gpu): every golden command byte-for-byte (stdout md5, stderr trailer, exit code), the
`batch` sub-command against the single-pair blocks, and live runs beside the compiled reference
CLI (oracle/_ref/alignTools_ref) on crafted inputs."""
import base64
import gzip
import hashlib
import json
import os
import subprocess

import pytest

from helpers import GOLD, load_cli

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "bin", "alignTools")
DUMP = os.path.join(ROOT, "bin", "at_fasta_dump")


@pytest.fixture(scope="module")
def built():
    import aligntools.c_b200 as A
    A.build()
    assert os.path.exists(CLI) and os.path.exists(DUMP)
    return A


@pytest.fixture(scope="module")
def fasta_dir(tmp_path_factory):
    """The reference's test/*.fa rebuilt from the records stored with the golden vectors."""
    d = tmp_path_factory.mktemp("fa")
    for fname, info in load_cli()["files"].items():
        with open(d / fname, "w") as f:
            for r in info["records"]:
                f.write(">" + r["name"] + (" " + r["comment"] if r["comment"] else "") + "\n" + r["seq"] + "\n")
    return str(d)


def run_cli(argv, fasta_dir):
    args = [a.replace("$T", fasta_dir) for a in argv]
    pr = subprocess.run([CLI] + args, capture_output=True)
    return pr.returncode, pr.stdout, pr.stderr


def expected_stderr(v, fasta_dir):
    return v["stderr"].replace("$BIN", CLI).replace("$T", fasta_dir).encode()


# ------------------------------------------------------------------ CPU ----
def test_fasta_reader_matches_kseq_fixtures(built, tmp_path):
    with open(os.path.join(GOLD, "fasta_cases.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) >= 30
    for c in cases:
        p = tmp_path / (c["name"] + (".gz" if c["gz"] else ".fa"))
        data = base64.b64decode(c["input_b64"])
        with (gzip.open(p, "wb") if c["gz"] else open(p, "wb")) as f:
            f.write(data)
        got = subprocess.run([DUMP, str(p)], capture_output=True).stdout
        assert got == base64.b64decode(c["dump_b64"]), c["name"]


def test_fasta_reader_streams_across_window_boundaries(built, tmp_path):
    """The reader inflates its input through a fixed window; rebuilt with windows of 1, 3 and 7 bytes every token of
    every fixture straddles a refill, and the records must still come out as kseq's."""
    with open(os.path.join(GOLD, "fasta_cases.json")) as f:
        cases = json.load(f)["cases"]
    host = os.path.join(ROOT, "host")
    for win in (1, 3, 7):
        exe = tmp_path / f"dump_w{win}"
        subprocess.run(["gcc", "-std=gnu11", "-O1", f"-DAT_FASTA_WINDOW={win}u", "-I", host, os.path.join(host, "at_fasta_dump.c"),
                        os.path.join(host, "at_fasta.c"), "-o", str(exe), "-lz"], check=True)
        for c in cases:
            p = tmp_path / (c["name"] + f"_w{win}" + (".gz" if c["gz"] else ".fa"))
            data = base64.b64decode(c["input_b64"])
            with (gzip.open(p, "wb") if c["gz"] else open(p, "wb")) as f:
                f.write(data)
            got = subprocess.run([str(exe), str(p)], capture_output=True).stdout
            assert got == base64.b64decode(c["dump_b64"]), (win, c["name"])


def test_fasta_reader_matches_live_reference(built, oracle_mod, tmp_path):
    """Random record soups through both parsers (only where oracle/_ref was built)."""
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref not built")
    import random
    rng = random.Random(7)
    pieces = [b">", b"@", b"+", b"\n", b"\r\n", b" ", b"\t", b"ACGT", b"acgtn", b"|", b"12", b"name", b"IIII", b"\n\n", b"x"]
    for k in range(300):
        data = b"".join(rng.choice(pieces) for _ in range(rng.randint(0, 60)))
        p = tmp_path / f"soup{k}.fa"
        p.write_bytes(data)
        got = subprocess.run([DUMP, str(p)], capture_output=True).stdout
        assert got == oracle_mod.ref_kseq_dump(str(p)), data


def test_cli_paths_that_end_before_the_alignment(built, fasta_dir):
    """Usage texts, unknown command, option errors, unreadable file, fit with l1 > l2 (golden
    vectors X2-X7): stdout, stderr and exit code equal the reference's, no GPU involved."""
    n = 0
    for v in load_cli()["vectors"]:
        if v["rc"] == 0:
            continue
        rc, out, err = run_cli(v["argv"], fasta_dir)
        assert rc == v["rc"], (v["id"], rc, err)
        assert hashlib.md5(out).hexdigest() == v["stdout_md5"], v["id"]
        assert err == expected_stderr(v, fasta_dir), (v["id"], err)
        n += 1
    assert n >= 6


def test_cli_input_errors(built, tmp_path):
    """kstring_read's failure modes (src/alignment.h:229-244)."""
    three = tmp_path / "three.fa"; three.write_text(">a\nAC\n>b\nGT\n>c\nTT\n")
    one = tmp_path / "one.fa"; one.write_text(">a\nACGT\n")
    nocomment = tmp_path / "nc.fa"; nocomment.write_text(">a\nACGT\n>b\nACGTACGT\n")
    for args, msg in ((["global", str(three)], b"FATAL ERROR: input fasta file has more than 2 sequences\n"),
                      (["edit", str(one)], b"FATAL ERROR: read_kstring: fail to read sequence\n"),
                      (["fit", "-s", str(nocomment)], b"FATAL ERROR: fail to read junction sites\n"),
                      (["local", str(tmp_path / "missing.fa")], b"FATAL ERROR: Can't open " + str(tmp_path / "missing.fa").encode() + b"\n\n")):
        pr = subprocess.run([CLI] + args, capture_output=True)
        assert pr.returncode == 255 and pr.stderr == msg and pr.stdout == b"", (args, pr.stderr)
    for args in (["local", "-j", "3", str(one)], ["edit", "-s", str(one)], ["overlap", "-x", str(one)]):
        pr = subprocess.run([CLI] + args, capture_output=True)
        assert pr.returncode == 1 and pr.stdout == b"", args
    pr = subprocess.run([CLI, "fit"], capture_output=True)
    assert pr.returncode == 1 and b"-j INT   jump penality [-10]" in pr.stderr and b"-s       weather jump state include" in pr.stderr
    pr = subprocess.run([CLI, "edit"], capture_output=True)
    assert pr.returncode == 1 and pr.stderr == (b"\nUsage:   alignTools edit [options] <target.fa>\n\n"
                                                b"Options: -u INT   mismatch penalty [-2]\n         -o INT   gap penalty [-5]\n\n")


def test_host_sources_do_not_touch_the_oracle():
    for fn in os.listdir(os.path.join(ROOT, "host")):
        src = open(os.path.join(ROOT, "host", fn), errors="ignore").read()
        assert "oracle" not in src.replace("the oracle", ""), fn


# ------------------------------------------------------------------ GPU ----
@pytest.mark.gpu
def test_cli_golden_vectors_on_gpu(built, fasta_dir):
    """All 35 golden commands (SURVEY.md Appendix B + contract rows): stdout md5, stderr, rc."""
    n = 0
    for v in load_cli()["vectors"]:
        rc, out, err = run_cli(v["argv"], fasta_dir)
        assert rc == v["rc"], (v["id"], rc, err[-300:])
        assert hashlib.md5(out).hexdigest() == v["stdout_md5"], (v["id"], out[:120])
        assert err == expected_stderr(v, fasta_dir), (v["id"], err[-300:])
        n += 1
    assert n >= 35


def _write_pairs(path, pairs):
    with open(path, "w") as f:
        for k, (s1, s2, com) in enumerate(pairs):
            f.write(f">r{k}\n{s1}\n>t{k}" + (f" {com}" if com is not None else "") + f"\n{s2}\n")


@pytest.mark.gpu
@pytest.mark.parametrize("mode,opts", [("global", ["-m", "2", "-u", "-3", "-o", "-4", "-e", "-1"]), ("local", []),
                                      ("fit", []), ("fit", ["-s", "-j", "-6"]), ("overlap", []), ("edit", ["-u", "1"])])
def test_batch_subcommand_equals_single_pair_runs(built, tmp_path, mode, opts):
    """`alignTools batch <mode> ... pairs.fa` prints, pair after pair, exactly what the legacy
    sub-command prints for each pair alone; -c gives one TSV line per pair."""
    import random
    rng = random.Random(99)
    pairs = []
    for k in range(40):
        l1 = rng.randint(5, 400)
        s1 = "".join(rng.choice("ACGT") for _ in range(l1))
        s2 = "".join(rng.choice("ACGT") for _ in range(rng.randint(0, 30))) + \
            "".join(c if rng.random() > 0.08 else rng.choice("ACGT") for c in s1) + \
            "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 300)))
        com = "|".join(str(rng.randrange(len(s2))) for _ in range(rng.randint(1, 5)))
        pairs.append((s1, s2, com))
    allp = tmp_path / "pairs.fa"
    _write_pairs(allp, pairs)
    pr = subprocess.run([CLI, "batch", mode] + opts + [str(allp)], capture_output=True)
    assert pr.returncode == 0, pr.stderr[-300:]
    want = b""
    for k, p in enumerate(pairs[:3]):
        one = tmp_path / f"p{k}.fa"
        _write_pairs(one, [p])
        single = subprocess.run([CLI, mode] + (opts if mode != "overlap" else []) + [str(one)], capture_output=True)
        assert single.returncode == 0, single.stderr[-300:]
        want += single.stdout
    assert pr.stdout.startswith(want)
    assert pr.stdout.count(b"\n") == len(pairs) * (1 if mode == "edit" else (3 + (mode == "fit") + ("-s" in opts)))
    tsv = subprocess.run([CLI, "batch", mode] + opts + ["-c", str(allp)], capture_output=True)
    assert tsv.returncode == 0
    rows = tsv.stdout.decode().strip().split("\n")
    assert len(rows) == len(pairs) and all(len(r.split("\t")) == 8 for r in rows)
    # two-file form: reads and targets in separate files
    rf, tf = tmp_path / "reads.fa", tmp_path / "targets.fa"
    with open(rf, "w") as f:
        f.writelines(f">r{k}\n{p[0]}\n" for k, p in enumerate(pairs))
    with open(tf, "w") as f:
        f.writelines(f">t{k} {p[2]}\n{p[1]}\n" for k, p in enumerate(pairs))
    two = subprocess.run([CLI, "batch", mode] + opts + [str(rf), str(tf)], capture_output=True)
    assert two.returncode == 0 and two.stdout == pr.stdout


@pytest.mark.gpu
def test_cli_beside_the_compiled_reference(built, oracle_mod, tmp_path):
    """Live: the host and the reference CLI (oracle/_ref/alignTools_ref, shipped as a binary) on
    crafted inputs -- multi-line FASTA, FASTQ, gz, lower-case, junction comments."""
    if not os.path.exists(oracle_mod.REF_CLI):
        pytest.skip("oracle/_ref/alignTools_ref not built")
    import random
    rng = random.Random(5)
    files = []
    for k in range(3):
        l1 = rng.randint(20, 300)
        s1 = "".join(rng.choice("ACGTacgt") for _ in range(l1))
        s2 = "".join(rng.choice("ACGT") for _ in range(rng.randint(5, 50))) + s1.upper() + "".join(rng.choice("ACGT") for _ in range(rng.randint(5, 200)))
        wrap = lambda s, w: "\n".join(s[i:i + w] for i in range(0, len(s), w))
        com = "|".join(str(rng.randrange(len(s2))) for _ in range(4))
        if k % 3 == 0:
            body = f">read{k} some text\n{wrap(s1, 60)}\n>gene{k} {com}\n{wrap(s2, 70)}\n"
        elif k % 3 == 1:
            body = f"@read{k}\n{s1}\n+\n{'I' * len(s1)}\n@gene{k} {com}\n{s2}\n+\n{'J' * len(s2)}\n"
        else:
            body = f">read{k}\r\n{wrap(s1, 50)}\r\n\r\n>gene{k}\t{com}\r\n{s2}\r\n"
        p = tmp_path / (f"c{k}.fa" + (".gz" if k == 1 else ""))
        with (gzip.open(p, "wb") if k == 1 else open(p, "wb")) as f:
            f.write(body.encode())
        files.append(str(p))
    cmds = [["global", "-m", "3", "-u", "-1", "-o", "-6", "-e", "-2"], ["local", "-m", "2", "-u", "-2", "-o", "-5", "-e", "-2"],
            ["fit"], ["fit", "-s", "-j", "-3", "-m", "2"], ["overlap"], ["edit", "-u", "1"]]
    for fn in files:
        for cmd in cmds:
            ours = subprocess.run([CLI] + cmd + [fn], capture_output=True)
            ref = subprocess.run([oracle_mod.REF_CLI] + cmd + [fn], capture_output=True)
            assert ours.returncode == ref.returncode, (cmd, fn, ours.stderr[-200:])
            assert ours.stdout == ref.stdout, (cmd, fn)
            assert ours.stderr == ref.stderr.replace(oracle_mod.REF_CLI.encode(), CLI.encode()), (cmd, fn)


def _random_pairs(rng, n, jump_sites=True):
    pairs = []
    for k in range(n):
        l1 = rng.randint(5, 300)
        s1 = "".join(rng.choice("ACGT") for _ in range(l1))
        s2 = "".join(rng.choice("ACGT") for _ in range(rng.randint(0, 30))) + \
            "".join(c if rng.random() > 0.08 else rng.choice("ACGT") for c in s1) + \
            "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 200)))
        com = "|".join(str(rng.randrange(len(s2))) for _ in range(rng.randint(1, 5))) if jump_sites else None
        pairs.append((s1, s2, com))
    return pairs


@pytest.mark.gpu
@pytest.mark.parametrize("mode,opts", [("local", ["-m", "2", "-u", "-2", "-o", "-5", "-e", "-2"]), ("fit", ["-s", "-j", "-6"]), ("edit", ["-u", "1"])])
def test_batch_streams_blocks(built, tmp_path, mode, opts):
    """The batch sub-command streams its input in blocks (-B pairs per block: parse block k+1 while block k is on
    the GPU, print block k-1): the output must not depend on the block size, also for gzip'ed input."""
    import random
    pairs = _random_pairs(random.Random(4), 157)
    fa = tmp_path / "pairs.fa"
    _write_pairs(fa, pairs)
    gz = tmp_path / "pairs.fa.gz"
    with gzip.open(gz, "wb") as f:
        f.write(fa.read_bytes())
    whole = subprocess.run([CLI, "batch", mode] + opts + [str(fa)], capture_output=True)
    assert whole.returncode == 0, whole.stderr[-300:]
    for blk, src in (("1", fa), ("7", gz), ("64", fa), ("157", fa)):
        for fmt in ([], ["-c"], ["-S"]):
            ref = whole.stdout if not fmt else subprocess.run([CLI, "batch", mode] + opts + fmt + [str(fa)], capture_output=True).stdout
            pr = subprocess.run([CLI, "batch", mode] + opts + fmt + ["-B", blk, str(src)], capture_output=True)
            assert pr.returncode == 0, pr.stderr[-300:]
            assert pr.stdout.replace(str(src).encode(), b"F").replace(b" -B " + blk.encode(), b"") == ref.replace(str(fa).encode(), b"F"), (blk, fmt)


@pytest.mark.gpu
def test_batch_sam_records(built, oracle_mod, tmp_path):
    """-S: SAM-like records.  Every record is checked against the oracle: POS is the first aligned target base, the
    CIGAR (with S for the read's unaligned ends) consumes exactly the read, and replaying it gives the oracle's
    alignment columns; AS is the score.  The header carries the @PG line the reference builds (src/main.c:36-38)."""
    import random
    pairs = _random_pairs(random.Random(12), 60, jump_sites=False)
    fa = tmp_path / "pairs.fa"
    _write_pairs(fa, pairs)
    import re
    for mode, prm in (("local", oracle_mod.Params(2, -2, -5, -2, -10, False)), ("global", oracle_mod.Params()), ("overlap", oracle_mod.Params())):
        opts = ["-m", "2", "-u", "-2", "-o", "-5", "-e", "-2"] if mode == "local" else []
        pr = subprocess.run([CLI, "batch", mode] + opts + ["-S", str(fa)], capture_output=True)
        assert pr.returncode == 0, pr.stderr[-300:]
        lines = pr.stdout.decode().strip().split("\n")
        assert lines[0] == "@HD\tVN:1.6\tSO:unsorted" and lines[1].startswith("@PG\tID:alignTools\tPN:alignTools\tVN:0.7.23-r15\tCL:")
        recs = [ln.split("\t") for ln in lines[2:]]
        assert len(recs) == len(pairs)
        for k, (f, (s1, s2, _)) in enumerate(zip(recs, pairs)):
            ref = oracle_mod.port_align(mode, s1.encode(), s2.encode(), prm)
            assert f[0] == f"r{k}" and f[2] == f"t{k}" and f[9] == s1 and f[11] == f"AS:i:{ref.score}", (mode, k)
            if not ref.ops:
                assert f[1] == "4" and f[5] == "*"
                continue
            ops = re.findall(r"(\d+)([MIDNS])", f[5])
            assert sum(int(n) for n, c in ops if c in "MIS") == len(s1), (mode, k, f[5])
            body = "".join(c * int(n) for n, c in ops if c != "S")
            assert body == ref.ops.decode(), (mode, k)
            assert int(f[3]) == (1 if mode == "global" else ref.coords[3] + 1), (mode, k)


@pytest.mark.gpu
def test_batch_fit_junction_lists(built, oracle_mod, tmp_path):
    """fit -s in batch mode: every target brings its own junction list; a target header without one is an error (the
    kseq quirk of inheriting the previous record's comment must not leak another pair's junctions); -w reads the list
    as a whitelist."""
    import random
    rng = random.Random(3)
    pairs = []
    for k in range(30):
        l2 = rng.randint(60, 300)
        s2 = "".join(rng.choice("ACGT") for _ in range(l2))
        a, b = sorted(rng.sample(range(10, l2 - 10), 2))
        s1 = s2[max(0, a - 25):a] + s2[b:b + 25]
        pairs.append((s1, s2, f"{a}|{b - 1}|{b}"))
    fa = tmp_path / "j.fa"
    _write_pairs(fa, pairs)
    for flag, jump in (([], True), (["-w"], 2)):
        pr = subprocess.run([CLI, "batch", "fit", "-s", "-j", "-4", "-c"] + flag + [str(fa)], capture_output=True)
        assert pr.returncode == 0, pr.stderr[-300:]
        rows = [r.split("\t") for r in pr.stdout.decode().strip().split("\n")]
        for k, (s1, s2, com) in enumerate(pairs):
            ref = oracle_mod.port_align("fit", s1.encode(), s2.encode(), oracle_mod.Params(1, -2, -5, -1, -4, jump), [int(x) for x in com.split("|")])
            assert int(rows[k][2]) == ref.score, (flag, k)
            assert "".join(c * int(n) for n, c in __import__("re").findall(r"(\d+)([MIDN])", rows[k][7])) == ref.ops.decode(), (flag, k)
    bad = tmp_path / "bad.fa"
    _write_pairs(bad, pairs[:2] + [(pairs[2][0], pairs[2][1], None)] + pairs[3:5])
    pr = subprocess.run([CLI, "batch", "fit", "-s", str(bad)], capture_output=True)
    assert pr.returncode == 255 and pr.stderr == b"FATAL ERROR: fail to read junction sites\n"


@pytest.mark.gpu
def test_reference_main_with_five_call_sites_bound_to_the_library(built, oracle_mod, fasta_dir):
    """INTEGRATION.md 1 made literal: oracle/_ref/alignTools_dropin is the reference's OWN main.c / alignment.h /
    kstring.c (compiled by oracle/Makefile from a scratch copy) in which only the five call sites (:345, :509, :736, :885,
    :1000) call the at_* shims, linked against libaligntools_b200.so.  Every golden command must come out byte for byte."""
    exe = os.path.join(os.path.dirname(oracle_mod.REF_CLI), "alignTools_dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/alignTools_dropin not built (needs /root/reference at build time)")
    n = 0
    for v in load_cli()["vectors"]:
        args = [a.replace("$T", fasta_dir) for a in v["argv"]]
        pr = subprocess.run([exe] + args, capture_output=True)
        assert pr.returncode == v["rc"], (v["id"], pr.returncode, pr.stderr[-300:])
        assert hashlib.md5(pr.stdout).hexdigest() == v["stdout_md5"], (v["id"], pr.stdout[:120])
        assert pr.stderr == v["stderr"].replace("$BIN", exe).replace("$T", fasta_dir).encode(), (v["id"], pr.stderr[-300:])
        n += 1
    assert n >= 35
