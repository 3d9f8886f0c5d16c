/* at_fasta_dump.c -- test driver of the host's FASTA/FASTQ reader: prints every record as
 *     <name>\t<comment or "(null)">\t<seq_len>\t<seq>\n
 * so the CPU tests can compare the reader with kseq's behaviour (tests/test_host_cli.py). */
#include <stdio.h>
#include "at_fasta.h"

int main(int argc, char **argv)
{
	if (argc < 2) { fprintf(stderr, "usage: at_fasta_dump <file>\n"); return 1; }
	at_fasta *f = at_fasta_open(argv[1]);
	if (!f) { fprintf(stderr, "Can't open %s\n", argv[1]); return 2; }
	at_fasta_rec r;
	while (at_fasta_next(f, &r))
		printf("%s\t%s\t%zu\t%s\n", r.name, r.comment ? r.comment : "(null)", r.seq_len, r.seq);
	at_fasta_close(f);
	return 0;
}
