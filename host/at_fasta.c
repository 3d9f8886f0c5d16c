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
 */
#include "at_fasta.h"

#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct { char *s; size_t l, m; } strbuf;

struct at_fasta {
	unsigned char *buf;     /* the whole inflated file */
	size_t n, pos;
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

static int sb_set(strbuf *b, const unsigned char *src, size_t len)
{
	if (sb_reserve(b, len + 1)) return -1;
	memcpy(b->s, src, len);
	b->l = len; b->s[len] = 0;
	return 0;
}

static int sb_append(strbuf *b, const unsigned char *src, size_t len)
{
	if (sb_reserve(b, b->l + len + 1)) return -1;
	memcpy(b->s + b->l, src, len);
	b->l += len; b->s[b->l] = 0;
	return 0;
}

at_fasta *at_fasta_open(const char *path)
{
	gzFile fp = path ? gzopen(path, "r") : NULL;
	if (!fp) return NULL;
	at_fasta *f = (at_fasta *)calloc(1, sizeof *f);
	if (!f) { gzclose(fp); return NULL; }
	size_t cap = 1 << 16;
	f->buf = (unsigned char *)malloc(cap);
	while (f->buf) {
		if (f->n == cap) {
			cap *= 2;
			unsigned char *p = (unsigned char *)realloc(f->buf, cap);
			if (!p) { free(f->buf); f->buf = NULL; break; }
			f->buf = p;
		}
		const size_t want = cap - f->n;
		int got = gzread(fp, f->buf + f->n, (unsigned)(want > (1u << 30) ? (1u << 30) : want));
		if (got <= 0) break;      /* end of file, or an unreadable stream: whatever was inflated so far is the input */
		f->n += (size_t)got;
	}
	gzclose(fp);
	if (!f->buf) { free(f); return NULL; }
	return f;
}

void at_fasta_close(at_fasta *f)
{
	if (!f) return;
	free(f->buf); free(f->name.s); free(f->comment.s); free(f->seq.s);
	free(f);
}

/* end of the line that starts at `from`: index of its '\n', or n */
static size_t line_end(const at_fasta *f, size_t from)
{
	const unsigned char *p = from < f->n ? (const unsigned char *)memchr(f->buf + from, '\n', f->n - from) : NULL;
	return p ? (size_t)(p - f->buf) : f->n;
}

/* append the rest of the current line to b; CR rule of the line reader.  Returns 0 when the
 * input was already exhausted (nothing appended, no CR rule), 1 otherwise. */
static int take_line(at_fasta *f, strbuf *b)
{
	if (f->pos >= f->n) return 0;
	const size_t e = line_end(f, f->pos);
	sb_append(b, f->buf + f->pos, e - f->pos);
	f->pos = e < f->n ? e + 1 : f->n;
	if (b->l > 1 && b->s[b->l - 1] == '\r') b->s[--b->l] = 0;
	return 1;
}

int at_fasta_next(at_fasta *f, at_fasta_rec *rec)
{
	if (!f || !rec) return 0;
	if (!f->header_seen) {
		while (f->pos < f->n && f->buf[f->pos] != '>' && f->buf[f->pos] != '@') ++f->pos;
		if (f->pos >= f->n) return 0;
		++f->pos;
	}
	f->header_seen = 0;
	if (f->pos >= f->n) return 0;                 /* a lone header character at the very end */
	/* name */
	size_t k = f->pos;
	while (k < f->n && !isspace(f->buf[k])) ++k;
	if (sb_set(&f->name, f->buf + f->pos, k - f->pos)) return 0;
	const int delim = k < f->n ? f->buf[k] : 0;
	f->pos = k < f->n ? k + 1 : f->n;
	/* comment */
	if (delim != '\n' && f->pos < f->n) {
		f->comment.l = 0;
		if (f->comment.s) f->comment.s[0] = 0;
		take_line(f, &f->comment);
		if (!f->comment.s) sb_set(&f->comment, (const unsigned char *)"", 0);
		f->have_comment = 1;
	}
	/* sequence */
	if (sb_set(&f->seq, (const unsigned char *)"", 0)) return 0;
	int c = -1;
	while (f->pos < f->n) {
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
		const size_t e = line_end(f, f->pos);
		if (e >= f->n) return 0;                  /* no quality string */
		f->pos = e + 1;
		strbuf qual = {0, 0, 0};
		while (take_line(f, &qual) && qual.l < f->seq.l) {}
		const size_t ql = qual.l;
		free(qual.s);
		if (ql != f->seq.l) return 0;             /* truncated quality string */
	}
	rec->name = f->name.s;
	rec->comment = f->have_comment ? f->comment.s : NULL;
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
