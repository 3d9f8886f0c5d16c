/* at_fasta.c -- see at_fasta.h.  Record grammar (behaviour of kseq_read, src/kseq.h:189-229):
 *
 *   header   : first '>' or '@' found when scanning forward (anything before it is skipped)
 *   name     : header bytes up to the first isspace() byte
 *   comment  : if that byte is not '\n', the rest of the header line (one trailing CR dropped when
 *              the comment is longer than one byte); a header WITHOUT a comment leaves the previous
 *              record's comment in place -- the reference tests `seq->comment.s`, which kseq never
 *              clears
 *   sequence : following lines, concatenated, until a line STARTS with '>', '@' or '+'; empty lines
 *              are skipped; after every line one trailing CR is dropped while the sequence is longer
 *              than one byte
 *   '+'      : FASTQ -- the rest of that line is skipped and quality lines are consumed until they
 *              cover the sequence length; a quality string of a different length (or a missing one)
 *              ends the input without yielding the record (kseq returns -2, the caller's loop stops)
 *
 * The input is STREAMED: the (possibly gzip'ed) file is inflated through a 256 KiB window, so a batch of any
 * size is parsed in constant memory while earlier blocks are already on the GPU (host/alignTools.c, `batch`).
 */
#include "at_fasta.h"

#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#ifndef AT_FASTA_WINDOW
#define AT_FASTA_WINDOW (1u << 18)      /* the tests also build the reader with a window of a few bytes */
#endif

typedef struct { char *s; size_t l, m; } strbuf;

struct at_fasta {
	gzFile fp;
	unsigned char *buf;     /* the window of inflated bytes */
	size_t n, pos;          /* valid bytes / read position in the window */
	int eof;                /* the stream is exhausted (or unreadable: whatever was inflated so far is the input) */
	int header_seen;        /* the next record's '>' / '@' has already been consumed */
	int have_comment;       /* comment.s holds a string (possibly a previous record's) */
	strbuf name, comment, seq;
};

static int sb_reserve(strbuf *b, size_t need)
{
	if (need <= b->m) return 0;
	size_t m = b->m ? b->m : 256;
	while (m < need) m *= 2;
	char *p = (char *)realloc(b->s, m);
	if (!p) return -1;
	b->s = p; b->m = m;
	return 0;
}

static int sb_append(strbuf *b, const unsigned char *src, size_t len)
{
	if (sb_reserve(b, b->l + len + 1)) return -1;
	memcpy(b->s + b->l, src, len);
	b->l += len; b->s[b->l] = 0;
	return 0;
}

static int sb_clear(strbuf *b)
{
	if (sb_reserve(b, 1)) return -1;
	b->l = 0; b->s[0] = 0;
	return 0;
}

/* make sure the window holds an unread byte; 0 at the end of the input */
static int more(at_fasta *f)
{
	if (f->pos < f->n) return 1;
	if (f->eof) return 0;
	const int got = gzread(f->fp, f->buf, AT_FASTA_WINDOW);
	if (got <= 0) { f->eof = 1; f->n = f->pos = 0; return 0; }
	f->n = (size_t)got; f->pos = 0;
	return 1;
}

at_fasta *at_fasta_open(const char *path)
{
	gzFile fp = path ? gzopen(path, "r") : NULL;
	if (!fp) return NULL;
	at_fasta *f = (at_fasta *)calloc(1, sizeof *f);
	if (f) f->buf = (unsigned char *)malloc(AT_FASTA_WINDOW);
	if (!f || !f->buf) { free(f); gzclose(fp); return NULL; }
	gzbuffer(fp, 1u << 17);
	f->fp = fp;
	return f;
}

void at_fasta_close(at_fasta *f)
{
	if (!f) return;
	if (f->fp) gzclose(f->fp);
	free(f->buf); free(f->name.s); free(f->comment.s); free(f->seq.s);
	free(f);
}

/* append the rest of the current line to b (b == NULL: skip it) and consume its '\n'; CR rule of the line reader.
 * Returns 0 when the input was already exhausted (nothing appended, no CR rule), 1 when the line ended with
 * '\n', 2 when it ended with the input. */
static int take_line(at_fasta *f, strbuf *b)
{
	if (!more(f)) return 0;
	int ended = 2;
	while (more(f)) {
		const unsigned char *p = f->buf + f->pos;
		const unsigned char *nl = (const unsigned char *)memchr(p, '\n', f->n - f->pos);
		const size_t len = nl ? (size_t)(nl - p) : f->n - f->pos;
		if (b) sb_append(b, p, len);
		f->pos += len;
		if (nl) { ++f->pos; ended = 1; break; }
	}
	if (b && b->l > 1 && b->s[b->l - 1] == '\r') b->s[--b->l] = 0;
	return ended;
}

int at_fasta_next(at_fasta *f, at_fasta_rec *rec)
{
	if (!f || !rec) return 0;
	if (!f->header_seen) {
		for (;;) {
			if (!more(f)) return 0;
			const unsigned char c = f->buf[f->pos++];
			if (c == '>' || c == '@') break;
		}
	}
	f->header_seen = 0;
	if (!more(f)) return 0;                       /* a lone header character at the very end */
	/* name: up to the first isspace() byte */
	if (sb_clear(&f->name)) return 0;
	int delim = 0;
	while (more(f)) {
		size_t k = f->pos;
		while (k < f->n && !isspace(f->buf[k])) ++k;
		if (sb_append(&f->name, f->buf + f->pos, k - f->pos)) return 0;
		f->pos = k;
		if (k < f->n) { delim = f->buf[f->pos++]; break; }
	}
	/* comment */
	int own_comment = 0;
	if (delim != '\n' && more(f)) {
		if (sb_clear(&f->comment)) return 0;
		take_line(f, &f->comment);
		f->have_comment = 1;
		own_comment = 1;
	}
	/* sequence */
	if (sb_clear(&f->seq)) return 0;
	int c = -1;
	while (more(f)) {
		c = f->buf[f->pos++];
		if (c == '>' || c == '+' || c == '@') break;
		if (c != '\n') {
			const unsigned char ch = (unsigned char)c;
			sb_append(&f->seq, &ch, 1);
			take_line(f, &f->seq);
		}
		c = -1;
	}
	if (c == '>' || c == '@') f->header_seen = 1;
	if (c == '+') {                               /* FASTQ: skip the '+' line, then the quality string */
		if (take_line(f, NULL) != 1) return 0;    /* no quality string */
		strbuf qual = {0, 0, 0};
		while (take_line(f, &qual) && qual.l < f->seq.l) {}
		const size_t ql = qual.l;
		free(qual.s);
		if (ql != f->seq.l) return 0;             /* truncated quality string */
	}
	rec->name = f->name.s;
	rec->comment = f->have_comment ? f->comment.s : NULL;
	rec->own_comment = own_comment;
	rec->seq = f->seq.s;
	rec->seq_len = strlen(f->seq.s);              /* the reference strdup()s: an embedded NUL ends the sequence */
	return 1;
}

size_t at_parse_sites(const char *comment, int **out)
{
	*out = NULL;
	if (!comment) return 0;
	size_t n = 0, cap = 0;
	const char *p = comment;
	while (*p) {
		while (*p == '|') ++p;
		if (!*p) break;
		if (n == cap) {
			cap = cap ? 2 * cap : 8;
			int *q = (int *)realloc(*out, cap * sizeof(int));
			if (!q) { free(*out); *out = NULL; return 0; }
			*out = q;
		}
		(*out)[n++] = atoi(p);
		while (*p && *p != '|') ++p;
	}
	return n;
}
