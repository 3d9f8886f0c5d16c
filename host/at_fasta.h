/* at_fasta.h -- FASTA/FASTQ (optionally gzip'ed) record reader of the alignTools host.
 *
 * The reference reads its input through klib's kseq stream parser (KSEQ_INIT(gzFile, gzread),
 * src/alignment.h:23; kseq_read src/kseq.h:189-229) inside kstring_read (src/alignment.h:217-262).
 * This is an independent reader with the same observable record semantics (what counts as a
 * header, name / comment split, multi-line sequences, FASTQ quality skipping, CR stripping,
 * the comment carried over from the previous record when a header has none), streaming the
 * inflated bytes through a 256 KiB window.
 */
#ifndef AT_FASTA_H
#define AT_FASTA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct at_fasta_rec {
	const char *name;      /* NUL-terminated, owned by the reader, valid until the next call */
	const char *comment;   /* NULL when no header so far carried a comment (see at_fasta_next) */
	int         own_comment; /* 1 when THIS record's header carried the comment (0: inherited from an earlier record, the kseq quirk) */
	const char *seq;       /* NUL-terminated; seq_len == strlen(seq) as the reference strdup()s it */
	size_t      seq_len;
} at_fasta_rec;

typedef struct at_fasta at_fasta;

/* NULL when the file cannot be opened (the reference: die("Can't open %s\n"), :229). */
at_fasta *at_fasta_open(const char *path);
/* 1 = a record was read, 0 = end of input (also a truncated FASTQ record, as kseq's -2). */
int       at_fasta_next(at_fasta *f, at_fasta_rec *rec);
void      at_fasta_close(at_fasta *f);

/* junction sites of a header comment: fields separated by '|', empty fields skipped, each
 * field through atoi (kstring_read :250-253 via ksplit_core, src/kstring.c:89-131).
 * Returns the number of sites; *out is malloc'ed (NULL when there are none). */
size_t at_parse_sites(const char *comment, int **out);

#ifdef __cplusplus
}
#endif
#endif
