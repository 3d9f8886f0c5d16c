#!/usr/bin/env python
"""Generate tests/golden/fasta_cases.json: tricky FASTA / FASTQ / gz inputs and what the
REFERENCE's kseq parser (src/kseq.h:189-229, driven as in kstring_read src/alignment.h:229-237)
makes of them, obtained from oracle/_ref (ref_kseq_dump in oracle/ref_shim.c).  Run in the
container that has /root/reference; the fixture pins host/at_fasta.c on the GPU box too."""
import base64
import gzip
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

CASES = {
    "basic_comment": b">r1 first read\nACGTACGT\n>t1 10|20|30\nACGTTTACGTAA\n",
    "multiline_blank_lines": b">a\nACGT\n\nTTGA\n\n\n>b desc here\nAC\nGT\nAC\n",
    "crlf": b">a some comment\r\nACGT\r\nTTGA\r\n>b\r\nGGCC\r\n",
    "fastq_pair": b"@q1 c1\nACGTAC\n+\nIIIIII\n@q2\nTTGACA\n+q2\nJJJJJJ\n",
    "fastq_multiline": b"@q1\nACGT\nACGT\n+\nIIII\nIIII\n@q2 x|y\nTT\n+\nII\n",
    "fastq_truncated_qual": b"@q1\nACGTAC\n+\nIII\n",
    "fastq_no_qual": b"@q1\nACGTAC\n+",
    "fastq_qual_starts_with_at": b"@q1\nACGT\n+\n@III\n@q2\nGGCC\n+\nIIII\n",
    "stale_comment": b">a 5|9\nACGTACGTACGT\n>b\nACGTACGAACGTTT\n",
    "stale_comment_three": b">a zzz\nAC\n>b\nGT\n>c\nTT\n",
    "tab_comment": b">a\tcomment with\ttabs\nACGT\n>b \nTTTT\n",
    "empty_comment_space": b">a \nACGT\n>b  two spaces\nTTTT\n",
    "leading_junk": b"junk line\nmore junk\n>a\nACGT\n>b\nTT\n",
    "no_trailing_newline": b">a\nACGT\n>b\nTTGA",
    "header_only_at_end": b">a\nACGT\n>",
    "header_name_at_eof": b">a\nACGT\n>b",
    "name_space_at_eof": b">a c\nACGT\n>b ",
    "empty_sequence": b">a\n>b\nACGT\n",
    "plus_line_in_fasta": b">a\nACGT\n+\nIIII\n>b\nTT\n",
    "at_inside_sequence_line": b">a\nAC@GT\nTT>A\n>b\nGG+CC\n",
    "at_starts_sequence_line": b">a\nACGT\n@b\nTTGA\n",
    "lowercase_protein": b">p1\nPAKKfqifWEKQ\n>p2 x\nmeanly\n",
    "spaces_in_sequence": b">a\nAC GT\n TT\n>b\nG\tC\n",
    "lone_cr_line": b">a\nACGT\n\r\nTT\n>b\n\r\nGG\n",
    "cr_single_char": b">a\n\r\n>b\nA\r\n",
    "three_records": b">a\nAC\n>b\nGT\n>c\nTT\n",
    "one_record": b">a\nACGT\n",
    "empty_file": b"",
    "no_header": b"ACGT\nTTGA\n",
    "embedded_nul": b">a\nAC\x00GT\n>b\nTTGA\n",
    "long_lines": b">a long\n" + b"ACGT" * 9000 + b"\n" + b"TTGA" * 5000 + b"\n>b 1|2\n" + b"G" * 40000 + b"\n",
    "pipe_sites_odd": b">a\nACGT\n>b |12||7|x|-3|  9|4junk|\nACGTAC\n",
}


def main():
    if not oracle.have_ref():
        raise SystemExit("oracle/_ref missing: run `make -C oracle` where /root/reference exists")
    out = {"generator": "tests/golden/make_fasta_golden.py (reference kseq via oracle/_ref)", "cases": []}
    with tempfile.TemporaryDirectory() as td:
        for name, data in CASES.items():
            for gz in (False, True):
                if gz and name not in ("basic_comment", "fastq_pair", "long_lines"):
                    continue
                path = os.path.join(td, name + (".fa.gz" if gz else ".fa"))
                with (gzip.open(path, "wb") if gz else open(path, "wb")) as f:
                    f.write(data)
                dump = oracle.ref_kseq_dump(path)
                out["cases"].append({"name": name + ("_gz" if gz else ""), "gz": gz,
                                     "input_b64": base64.b64encode(data).decode(),
                                     "dump_b64": base64.b64encode(dump).decode()})
    with open(os.path.join(ROOT, "tests", "golden", "fasta_cases.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
