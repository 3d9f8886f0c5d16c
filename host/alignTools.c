/* alignTools.c -- C host of the B200-native alignTools: the reference's command line
 *     alignTools <global|local|fit|overlap|edit> [-m -u -o -e -j -s] <target.fa>
 * (src/main.c:32-57 and the five main_<mode> drivers, src/alignment.h:318-350, 476-517, 698-744,
 * 851-892, 967-1008) kept as a drop-in -- same option strings, usage text, stdout / stderr bytes
 * and exit codes -- with the DP itself running on the GPU through the C-ABI
 * (include/aligntools_b200.h).  The legacy sub-commands call the single-pair entry points that
 * carry the reference's own signatures (at_align_gla, ...); the new `batch` sub-command streams
 * the record pairs of its input through at_batch_align block by block (text, TSV or SAM-like output).  There is no CPU alignment code here:
 * without a B200 the alignment step dies with the library's error.
 *
 * Deliberate differences from the reference binary (crashes are not reproduced):
 *   - `edit -e ...` dereferences a NULL optarg in the reference (optstring "m:u:o:e", :323); here
 *     it is treated like any unhandled option (exit code 1);
 *   - the heap overflow in strrev (:178-182) that aborts some runs does not exist;
 *   - empty records and `fit` against a 1-base target (undefined in the reference) are errors.
 */
#define _GNU_SOURCE
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "aligntools_b200.h"
#include "at_fasta.h"

#ifndef PACKAGE_VERSION
#define PACKAGE_VERSION "0.7.23-r15"      /* the trailer is part of the output contract (src/main.c:6-8, 50) */
#endif

enum { JUMP_ON = 0, JUMP_OFF = 1 };        /* the reference's `bool`: true == 0 (src/alignment.h:24) */

static void die(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	fprintf(stderr, "FATAL ERROR: ");
	vfprintf(stderr, fmt, ap);
	fprintf(stderr, "\n");
	va_end(ap);
	exit(-1);
}

static void *xcalloc(size_t n, size_t sz)
{
	void *p = calloc(n ? n : 1, sz);
	if (!p) die("mycalloc failure requesting %d of size %d bytes", (int)n, (int)sz);
	return p;
}

static char *xstrdup(const char *s)
{
	char *p = strdup(s);
	if (!p) die("mycalloc failure requesting %d of size %d bytes", (int)strlen(s) + 1, 1);
	return p;
}

static void default_opt(at_opt_t *o)       /* init_opt, src/alignment.h:102-114 */
{
	memset(o, 0, sizeof *o);
	o->o = -5; o->e = -1; o->m = 1; o->u = -2; o->j = -10; o->s = JUMP_OFF;
}

/* ---------------------------------------------------------------- option parsing ---- */
enum { M_GLOBAL, M_LOCAL, M_FIT, M_OVERLAP, M_EDIT };
static const char *const mode_name[] = {"global", "local", "fit", "overlap", "edit"};

/* One getopt loop per mode as in the reference: the option STRING is shared ("m:u:o:e:j:s",
 * edit: "m:u:o:e"), but only fit acts on -j / -s; everything else falls to `default: return 1`.
 * `extra` adds the batch-only flags (collected in `bo`).  Returns 0, or 1 when the driver must return 1. */
typedef struct { int tsv, sam, whitelist, gpus, block; } batch_opts;

static int parse_opts(int mode, int argc, char **argv, at_opt_t *opt, const char *extra, batch_opts *bo)
{
	char optstr[32];
	snprintf(optstr, sizeof optstr, "%s%s", mode == M_EDIT ? "m:u:o:e" : "m:u:o:e:j:s", extra ? extra : "");
	int c;
	while ((c = getopt(argc, argv, optstr)) >= 0) {
		switch (c) {
		case 'm': opt->m = atoi(optarg); break;
		case 'u': opt->u = atoi(optarg); break;
		case 'o': opt->o = atoi(optarg); break;
		case 'e': if (!optarg) return 1; opt->e = atoi(optarg); break;
		case 'j': if (mode != M_FIT) return 1; opt->j = atoi(optarg); break;
		case 's': if (mode != M_FIT) return 1; opt->s = JUMP_ON; break;
		case 'c': if (!bo) return 1; bo->tsv = 1; break;
		case 'S': if (!bo) return 1; bo->sam = 1; break;
		case 'w': if (!bo || mode != M_FIT) return 1; bo->whitelist = 1; break;
		case 'g': if (!bo) return 1; bo->gpus = atoi(optarg); break;
		case 'B': if (!bo) return 1; bo->block = atoi(optarg); break;
		default: return 1;
		}
	}
	return 0;
}

static void mode_usage(int mode, const at_opt_t *opt, const char *prefix)
{
	fprintf(stderr, "\n");
	fprintf(stderr, "Usage:   alignTools %s%s [options] <target.fa>\n\n", prefix, mode_name[mode]);
	if (mode == M_EDIT) {
		fprintf(stderr, "Options: -u INT   mismatch penalty [%d]\n", opt->u);
		fprintf(stderr, "         -o INT   gap penalty [%d]\n", opt->o);
	} else {
		fprintf(stderr, "Options: -m INT   score for a match [%d]\n", opt->m);
		fprintf(stderr, "         -u INT   mismatch penalty [%d]\n", opt->u);
		fprintf(stderr, "         -o INT   gap open penalty [%d]\n", opt->o);
		fprintf(stderr, "         -e INT   gap extension penalty [%d]\n", opt->e);
		if (mode == M_FIT) {
			fprintf(stderr, "         -j INT   jump penality [%d]\n", opt->j);
			fprintf(stderr, "         -s       weather jump state include\n");
		}
	}
	fprintf(stderr, "\n");
}

/* ------------------------------------------------------- legacy two-record input ---- */
/* kstring_read (src/alignment.h:217-262): at most two records, the third is fatal; with -s the
 * second record's comment is echoed on stdout and split into junction sites. */
static void read_pair(const char *fname, at_kstring_t *s1, at_kstring_t *s2, at_opt_t *opt)
{
	at_fasta *fa = at_fasta_open(fname);
	if (!fa) die("Can't open %s\n", fname);
	char *seq[2] = {NULL, NULL}, *comment[2] = {NULL, NULL};
	at_fasta_rec rec;
	int n = 0;
	while (at_fasta_next(fa, &rec)) {
		if (n >= 2) die("input fasta file has more than 2 sequences");
		seq[n] = xstrdup(rec.seq);
		if (rec.comment) comment[n] = xstrdup(rec.comment);
		++n;
	}
	at_fasta_close(fa);
	if (!seq[0] || !seq[1]) die("read_kstring: fail to read sequence");
	s1->s = seq[0]; s1->l = strlen(seq[0]); s1->m = s1->l + 1;
	s2->s = seq[1]; s2->l = strlen(seq[1]); s2->m = s2->l + 1;
	if (opt->s == JUMP_ON) {
		if (!comment[1]) die("fail to read junction sites");
		printf("%s\n", comment[1]);
		opt->sites.size = at_parse_sites(comment[1], &opt->sites.pos);
	}
	free(comment[0]); free(comment[1]);
}

static int main_legacy(int mode, int argc, char **argv)
{
	at_opt_t opt;
	default_opt(&opt);
	if (parse_opts(mode, argc, argv, &opt, NULL, NULL)) return 1;
	if (optind + 1 > argc) { mode_usage(mode, &opt, ""); return 1; }
	at_kstring_t ks1 = {0, 0, NULL}, ks2 = {0, 0, NULL}, r1 = {0, 0, NULL}, r2 = {0, 0, NULL};
	/* overlap opens argv[1], every other mode argv[argc-1] (src/alignment.h:994 vs :503) */
	read_pair(mode == M_OVERLAP ? argv[1] : argv[argc - 1], &ks1, &ks2, &opt);
	if (mode == M_EDIT) {
		printf("edit_distance=%d\n", at_edit_dist(&ks1, &ks2, &opt));
		free(ks1.s); free(ks2.s);
		return 0;
	}
	if (mode == M_FIT && ks1.l > ks2.l) die("first sequence must be shorter than the second\n");
	r1.s = (char *)xcalloc(ks1.l + ks2.l + 1, 1);
	r2.s = (char *)xcalloc(ks1.l + ks2.l + 1, 1);
	double score = 0;
	switch (mode) {
	case M_GLOBAL:  score = at_align_gla(&ks1, &ks2, &r1, &r2, &opt); break;
	case M_LOCAL:   score = at_align_local_affine(&ks1, &ks2, &r1, &r2, &opt); break;
	case M_FIT:     printf("asDAsdaSDAsdasDAsdaSD\n");      /* debug line of align_fit_affine_jump (:602) */
	                score = at_align_fit_affine_jump(&ks1, &ks2, &r1, &r2, &opt); break;
	default:        score = at_align_overlap(&ks1, &ks2, &r1, &r2, &opt); break;
	}
	if (mode == M_OVERLAP) printf("%f\n", score);            /* bare number (:1000) */
	else printf("score=%f\n", score);
	printf("%s\n%s\n", r1.s, r2.s);
	free(ks1.s); free(ks2.s); free(r1.s); free(r2.s); free(opt.sites.pos);
	return 0;
}

/* ------------------------------------------------------------ batch sub-command ---- */
/* alignTools batch <mode> [mode options] [-c | -S] [-w] [-g N] [-B pairs] <pairs.fa> [<targets.fa>]
 *   one file : records 2k and 2k+1 form pair k (read, target);
 *   two files: record k of the first file is the read, record k of the second the target.
 * The input is STREAMED (the generalisation of kstring_read, src/alignment.h:217-262, from two records to N):
 * records are parsed into blocks of up to -B pairs held in pinned host buffers; while block k is on the GPU(s)
 * through one at_batch_align call (which pipelines H2D / kernels / D2H inside), the main thread parses block k+1
 * and prints block k-1 -- memory stays bounded by three blocks whatever the size of the file.
 * Default output: per pair exactly the block the legacy sub-command prints for that pair.
 * -c: one TSV line per pair (read, target, score, beg_i, end_i, beg_j, end_j, CIGAR).
 * -S: SAM-like records (header @HD + the @PG line the reference builds but never prints, src/main.c:36-38): QNAME, FLAG,
 *     RNAME = target name, POS = first aligned target base (1-based), MAPQ 255, CIGAR with M / I / D, N for jump
 *     columns and S for the unaligned ends of the read, SEQ = the read, tags AS:i:<score> (edit: NM:i:<distance>).
 * -w: fit -s with the junction list as a WHITELIST (at_params.jump = 2; the semantics of the comment at :542-544).
 * -g N: spread every block over N GPUs (contiguous slices, no communication). */
#include <pthread.h>

enum { OUT_TEXT, OUT_TSV, OUT_SAM };

typedef struct {
	/* inputs: pinned (at_host_alloc) when the library can provide it */
	uint8_t *q, *t; size_t q_n, q_cap, t_n, t_cap;
	uint64_t *q_off, *t_off, *site_off; uint32_t *q_len, *t_len; size_t cap_pairs;
	int32_t *sites; size_t n_sites, cap_sites;
	char **q_name, **t_name, **t_comment;
	size_t n;
	/* outputs */
	at_batch_output out; size_t out_pairs, cig_cap, aln_cap;
	int rc; char err[512];
} block_t;

typedef struct { at_handle *h; int mode; at_params prm; uint32_t flags; block_t *blk; } job_t;

static void *pin_alloc(size_t bytes)
{
	void *p = at_host_alloc(bytes ? bytes : 1);      /* pinned: the library's copies overlap its kernels */
	if (!p) die("mycalloc failure requesting %d of size %d bytes", (int)bytes, 1);
	return p;
}

static void *pin_grow(void *old, size_t old_bytes, size_t new_bytes)
{
	void *p = pin_alloc(new_bytes);
	if (old) { memcpy(p, old, old_bytes); at_host_free(old); }
	return p;
}

static void block_reserve_pairs(block_t *b, size_t n)
{
	if (n <= b->cap_pairs) return;
	size_t cap = b->cap_pairs ? b->cap_pairs : 4096;
	while (cap < n) cap *= 2;
	b->q_off = (uint64_t *)pin_grow(b->q_off, b->cap_pairs * 8, cap * 8);
	b->t_off = (uint64_t *)pin_grow(b->t_off, b->cap_pairs * 8, cap * 8);
	b->site_off = (uint64_t *)pin_grow(b->site_off, (b->cap_pairs + 1) * 8, (cap + 1) * 8);
	b->q_len = (uint32_t *)pin_grow(b->q_len, b->cap_pairs * 4, cap * 4);
	b->t_len = (uint32_t *)pin_grow(b->t_len, b->cap_pairs * 4, cap * 4);
	b->q_name = (char **)realloc(b->q_name, cap * sizeof(char *));
	b->t_name = (char **)realloc(b->t_name, cap * sizeof(char *));
	b->t_comment = (char **)realloc(b->t_comment, cap * sizeof(char *));
	if (!b->q_name || !b->t_name || !b->t_comment) die("mycalloc failure requesting %d of size %d bytes", (int)cap, 8);
	if (!b->cap_pairs) b->site_off[0] = 0;
	b->cap_pairs = cap;
}

static void block_push_seq(uint8_t **buf, size_t *n, size_t *cap, const at_fasta_rec *r)
{
	if (*n + r->seq_len + 1 > *cap) {
		size_t c = *cap ? *cap : (size_t)1 << 20;
		while (*n + r->seq_len + 1 > c) c *= 2;
		*buf = (uint8_t *)pin_grow(*buf, *n, c);
		*cap = c;
	}
	memcpy(*buf + *n, r->seq, r->seq_len);
	*n += r->seq_len;
}

static void block_clear(block_t *b)
{
	for (size_t k = 0; k < b->n; ++k) { free(b->q_name[k]); free(b->t_name[k]); free(b->t_comment[k]); }
	b->n = 0; b->q_n = b->t_n = 0; b->n_sites = 0;
	if (b->site_off) b->site_off[0] = 0;
}

/* a read and its target join the block; `jump`: the junction list comes from the target record's OWN comment */
static void block_push_pair(block_t *b, const at_fasta_rec *read, const at_fasta_rec *target, int mode, int jump)
{
	block_reserve_pairs(b, b->n + 1);
	const size_t k = b->n;
	b->q_off[k] = b->q_n; b->q_len[k] = (uint32_t)read->seq_len;
	b->q_name[k] = xstrdup(read->name);
	block_push_seq(&b->q, &b->q_n, &b->q_cap, read);
	b->t_off[k] = b->t_n; b->t_len[k] = (uint32_t)target->seq_len;
	b->t_name[k] = xstrdup(target->name);
	b->t_comment[k] = NULL;
	block_push_seq(&b->t, &b->t_n, &b->t_cap, target);
	if (mode == M_FIT && read->seq_len > target->seq_len) die("first sequence must be shorter than the second\n");
	if (jump) {
		/* the legacy two-record path inherits kseq's quirk -- a header without a comment keeps the previous record's
		 * (SURVEY.md A.5); across the pairs of a batch that would hand one target another pair's junctions, so here
		 * the comment must be the target record's own */
		if (!target->comment || !target->own_comment) die("fail to read junction sites");
		b->t_comment[k] = xstrdup(target->comment);
		int *pos = NULL;
		const size_t ns = at_parse_sites(target->comment, &pos);
		if (b->n_sites + ns + 1 > b->cap_sites) {
			size_t c = b->cap_sites ? b->cap_sites : 4096;
			while (b->n_sites + ns + 1 > c) c *= 2;
			b->sites = (int32_t *)pin_grow(b->sites, b->n_sites * 4, c * 4);
			b->cap_sites = c;
		}
		for (size_t x = 0; x < ns; ++x) b->sites[b->n_sites + x] = pos[x];
		free(pos);
		b->n_sites += ns;
	}
	b->site_off[k + 1] = b->n_sites;
	b->n = k + 1;
}

typedef struct { at_fasta *fa[2]; int n_files; at_fasta_rec held; char *held_name, *held_comment, *held_seq; int have_held; } pair_reader;

/* next (read, target) pair of the input; 0 at its end.  One file: consecutive records; two files: one record of each */
static int next_pair(pair_reader *pr, at_fasta_rec *read, at_fasta_rec *target)
{
	if (pr->n_files == 2) {
		/* the read's strings must survive the second reader's call: they live in the FIRST reader's buffers */
		const int a = at_fasta_next(pr->fa[0], read), b = at_fasta_next(pr->fa[1], target);
		if (a != b) die("read_kstring: fail to read sequence");
		return a;
	}
	at_fasta_rec r;
	if (!at_fasta_next(pr->fa[0], &r)) return 0;
	free(pr->held_name); free(pr->held_comment); free(pr->held_seq);
	pr->held_name = xstrdup(r.name); pr->held_comment = r.comment ? xstrdup(r.comment) : NULL; pr->held_seq = xstrdup(r.seq);
	read->name = pr->held_name; read->comment = pr->held_comment; read->own_comment = r.own_comment;
	read->seq = pr->held_seq; read->seq_len = r.seq_len;
	if (!at_fasta_next(pr->fa[0], target)) die("read_kstring: fail to read sequence");      /* an odd number of records */
	return 1;
}

static void block_size_outputs(block_t *b, uint32_t flags, size_t cig_per_pair)
{
	const size_t n = b->n;
	if (n > b->out_pairs) {
		size_t cap = b->out_pairs ? b->out_pairs : 4096;
		while (cap < n) cap *= 2;
		free(b->out.score); free(b->out.end_i); free(b->out.end_j); free(b->out.beg_i); free(b->out.beg_j); free(b->out.cigar_off); free(b->out.aln_off);
		b->out.score = (int32_t *)xcalloc(cap, 4);
		b->out.end_i = (uint32_t *)xcalloc(cap, 4); b->out.end_j = (uint32_t *)xcalloc(cap, 4);
		b->out.beg_i = (uint32_t *)xcalloc(cap, 4); b->out.beg_j = (uint32_t *)xcalloc(cap, 4);
		b->out.cigar_off = (uint64_t *)xcalloc(cap + 1, 8); b->out.aln_off = (uint64_t *)xcalloc(cap + 1, 8);
		b->out_pairs = cap;
	}
	if (flags & AT_OUT_CIGAR) {
		size_t need = cig_per_pair * n + 1024;
		if (need > b->q_n + b->t_n + 16) need = b->q_n + b->t_n + 16;      /* never more than one op per column */
		if (need > b->cig_cap) { if (b->out.cigar) at_host_free(b->out.cigar); b->out.cigar = (uint32_t *)pin_alloc(need * 4); b->cig_cap = need; }
		b->out.cigar_cap = b->cig_cap;
	}
	if (flags & AT_OUT_ALN) {
		const size_t need = b->q_n + b->t_n + 16;                           /* an alignment has at most l1 + l2 columns */
		if (need > b->aln_cap) {
			if (b->out.aln1) at_host_free(b->out.aln1);
			if (b->out.aln2) at_host_free(b->out.aln2);
			b->out.aln1 = (char *)pin_alloc(need); b->out.aln2 = (char *)pin_alloc(need); b->aln_cap = need;
		}
		b->out.aln_cap = b->aln_cap;
	}
}

static void *align_block(void *arg)
{
	job_t *j = (job_t *)arg;
	block_t *b = j->blk;
	at_batch_input in;
	memset(&in, 0, sizeof in);
	in.n_pairs = b->n; in.encoding = AT_SEQ_BYTES;
	in.q = b->q; in.q_off = b->q_off; in.q_len = b->q_len;
	in.t = b->t; in.t_off = b->t_off; in.t_len = b->t_len;
	if (j->prm.jump) { in.sites = b->sites ? b->sites : (const int32_t *)b->site_off; in.site_off = b->site_off; }
	size_t per_pair = 64;
	for (;;) {      /* CIGAR capacity: a guess per pair, enlarged when the library reports AT_E_NOSPACE */
		block_size_outputs(b, j->flags, per_pair);
		b->rc = at_batch_align(j->h, j->mode, &j->prm, &in, j->flags, &b->out, NULL);
		if (b->rc != AT_E_NOSPACE || !(j->flags & AT_OUT_CIGAR) || b->cig_cap >= b->q_n + b->t_n + 16) break;
		per_pair *= 8;
	}
	if (b->rc) snprintf(b->err, sizeof b->err, "%s (%s)", at_strerror(b->rc), at_last_error(j->h));
	return NULL;
}

static char *g_cmdline = NULL;      /* the full command line, for the SAM @PG record */

static void print_block(const block_t *b, int mode, int jump, int fmt)
{
	static char *cig = NULL; static size_t cig_cap = 0;
	const at_batch_output *o = &b->out;
	for (size_t k = 0; k < b->n; ++k) {
		const char *cs = "*";
		if (fmt != OUT_TEXT && mode != M_EDIT) {
			const uint64_t o0 = o->cigar_off[k], o1 = o->cigar_off[k + 1];
			const size_t need = 12 * (size_t)(o1 - o0) + 40;
			if (need > cig_cap) { cig_cap = 2 * need; cig = (char *)realloc(cig, cig_cap); if (!cig) die("mycalloc failure requesting %d of size %d bytes", (int)cig_cap, 1); }
			size_t pos = 0;
			if (o1 > o0) {
				/* SAM: the read's unaligned ends are soft clips (global aligns end to end: its CIGAR covers the flush too) */
				if (fmt == OUT_SAM && mode != M_GLOBAL && o->beg_i[k]) pos += (size_t)snprintf(cig + pos, cig_cap - pos, "%uS", o->beg_i[k]);
				pos += (size_t)at_cigar_to_string(o->cigar + o0, o1 - o0, cig + pos, cig_cap - pos);
				if (fmt == OUT_SAM && mode != M_GLOBAL && o->end_i[k] < b->q_len[k]) pos += (size_t)snprintf(cig + pos, cig_cap - pos, "%uS", b->q_len[k] - o->end_i[k]);
				cs = cig;
			}
		}
		if (fmt == OUT_TSV) {
			printf("%s\t%s\t%d\t%u\t%u\t%u\t%u\t%s\n", b->q_name[k], b->t_name[k], o->score[k], o->beg_i[k], o->end_i[k], o->beg_j[k], o->end_j[k], cs);
		} else if (fmt == OUT_SAM) {
			const int mapped = mode != M_EDIT && cs[0] != '*';
			const unsigned pos1 = !mapped ? 0u : (mode == M_GLOBAL ? 1u : o->beg_j[k] + 1u);
			printf("%s\t%d\t%s\t%u\t%d\t%s\t*\t0\t0\t", b->q_name[k], mapped ? 0 : 4, b->t_name[k], pos1, mapped ? 255 : 0, cs);
			fwrite(b->q + b->q_off[k], 1, b->q_len[k], stdout);
			printf("\t*\t%s:i:%d\n", mode == M_EDIT ? "NM" : "AS", o->score[k]);
		} else if (mode == M_EDIT) {
			printf("edit_distance=%d\n", o->score[k]);
		} else {
			if (mode == M_FIT) {
				if (jump) printf("%s\n", b->t_comment[k]);
				printf("asDAsdaSDAsdasDAsdaSD\n");
			}
			if (mode == M_OVERLAP) printf("%f\n", (double)o->score[k]);
			else printf("score=%f\n", (double)o->score[k]);
			const uint64_t a0 = o->aln_off[k], a1 = o->aln_off[k + 1];
			fwrite(o->aln1 + a0, 1, (size_t)(a1 - a0), stdout); fputc('\n', stdout);
			fwrite(o->aln2 + a0, 1, (size_t)(a1 - a0), stdout); fputc('\n', stdout);
		}
	}
}

static void batch_usage(void)
{
	fprintf(stderr, "\n");
	fprintf(stderr, "Usage:   alignTools batch <global|local|fit|overlap|edit> [options] <pairs.fa> [<targets.fa>]\n\n");
	fprintf(stderr, "Options: -m -u -o -e (-j -s for fit) as in the single-pair commands\n");
	fprintf(stderr, "         -w       fit -s: the junction list names the ONLY positions where a jump may start\n");
	fprintf(stderr, "         -c       one TSV line per pair with the CIGAR instead of the text blocks\n");
	fprintf(stderr, "         -S       SAM-like records\n");
	fprintf(stderr, "         -g INT   number of GPUs [1]\n");
	fprintf(stderr, "         -B INT   pairs per streamed block [262144]\n");
	fprintf(stderr, "\n");
}

static int main_batch(int argc, char **argv)
{
	if (argc < 2) { batch_usage(); return 1; }
	int mode = -1;
	for (int k = 0; k < 5; ++k) if (strcmp(argv[1], mode_name[k]) == 0) mode = k;
	if (mode < 0) { fprintf(stderr, "[main] unrecognized command '%s'\n", argv[1]); return 1; }
	at_opt_t opt;
	default_opt(&opt);
	batch_opts bo = {0, 0, 0, 1, 262144};
	if (parse_opts(mode, argc - 1, argv + 1, &opt, "cSwg:B:", &bo)) return 1;
	char **files = argv + 1 + optind;
	const int n_files = argc - 1 - optind;
	if (n_files < 1 || n_files > 2) { batch_usage(); return 1; }
	const int jump = mode == M_FIT && opt.s == JUMP_ON;
	const int fmt = bo.sam ? OUT_SAM : (bo.tsv ? OUT_TSV : OUT_TEXT);
	if (bo.block < 1) bo.block = 1;
	const size_t max_symbols = (size_t)1 << 28;      /* a block also closes at 256 Mi symbols */

	pair_reader pr;
	memset(&pr, 0, sizeof pr);
	pr.n_files = n_files;
	for (int fi = 0; fi < n_files; ++fi) {
		pr.fa[fi] = at_fasta_open(files[fi]);
		if (!pr.fa[fi]) die("Can't open %s\n", files[fi]);
	}
	int devs[64];
	int gpus = bo.gpus < 1 ? 1 : (bo.gpus > 64 ? 64 : bo.gpus);
	for (int k = 0; k < gpus; ++k) devs[k] = k;
	at_handle *h = NULL;
	int rc = at_create(devs, gpus, &h);
	if (rc) die("aligntools-b200: %s", at_strerror(rc));
	job_t job;
	job.h = h; job.mode = mode;
	job.prm.m = opt.m; job.prm.u = opt.u; job.prm.o = opt.o; job.prm.e = opt.e; job.prm.j = opt.j; job.prm.jump = jump ? (bo.whitelist ? 2 : 1) : 0;
	job.flags = mode == M_EDIT ? 0u : (fmt == OUT_TEXT ? AT_OUT_ALN : AT_OUT_CIGAR);
	if (fmt == OUT_SAM) printf("@HD\tVN:1.6\tSO:unsorted\n@PG\tID:alignTools\tPN:alignTools\tVN:%s\tCL:%s\n", PACKAGE_VERSION, g_cmdline ? g_cmdline : "alignTools");

	/* three blocks in rotation: one being parsed, one on the GPU, one being printed */
	block_t blocks[3];
	memset(blocks, 0, sizeof blocks);
	size_t total = 0;
	int cur = 0, running = -1, more = 1;
	pthread_t th;
	while (more || running >= 0) {
		block_t *b = &blocks[cur];
		int filled = 0;
		if (more) {
			block_clear(b);
			at_fasta_rec read, target;
			while (b->n < (size_t)bo.block && b->q_n + b->t_n < max_symbols) {
				if (!next_pair(&pr, &read, &target)) { more = 0; break; }
				block_push_pair(b, &read, &target, mode, jump);
			}
			filled = b->n > 0;
			total += b->n;
		}
		int done = -1;
		if (running >= 0) {      /* the block on the GPU: wait for it, then hand the next one over before printing */
			pthread_join(th, NULL);
			done = running; running = -1;
			if (blocks[done].rc) die("%s", blocks[done].err);
		}
		if (filled) {
			job.blk = b;
			if (pthread_create(&th, NULL, align_block, &job)) die("pthread_create failed");
			running = cur;
		}
		if (done >= 0) print_block(&blocks[done], mode, jump, fmt);
		/* next block to fill: neither the one on the GPU nor the one just printed is touched until it has been printed */
		cur = (cur + 1) % 3;
	}
	if (total == 0) die("read_kstring: fail to read sequence");
	for (int fi = 0; fi < n_files; ++fi) at_fasta_close(pr.fa[fi]);
	at_destroy(h);
	return 0;
}

/* ------------------------------------------------------------------------ main ---- */
static int usage(void)
{
	fprintf(stderr, "\n");
	fprintf(stderr, "Program: alignTools (pairwise DNA sequence alignment)\n");
	fprintf(stderr, "Version: %s\n", PACKAGE_VERSION);
	fprintf(stderr, "Contact: Rongxin Fang <r3fang@ucsd.edu>\n\n");
	fprintf(stderr, "Usage:   alignTools <command> [options]\n\n");
	fprintf(stderr, "Command: global     global (needle) alignment allows affine gap\n");
	fprintf(stderr, "         local      smith-waterman with affine gap\n");
	fprintf(stderr, "         fit        fit alingment allows affine gap plus jump state\n");
	fprintf(stderr, "         overlap    overlap alignment\n");
	fprintf(stderr, "         edit       edit distance\n");
	fprintf(stderr, "\n");
	return 1;
}

int main(int argc, char *argv[])
{
	int ret = -1;
	if (argc < 2) return usage();
	{      /* the command line as the reference's @PG record spells it (src/main.c:36-38) */
		size_t len = 1;
		for (int i = 0; i < argc; ++i) len += strlen(argv[i]) + 1;
		g_cmdline = (char *)xcalloc(len, 1);
		for (int i = 0; i < argc; ++i) { if (i) strcat(g_cmdline, " "); strcat(g_cmdline, argv[i]); }
	}
	for (int k = 0; k < 5 && ret < 0; ++k)
		if (strcmp(argv[1], mode_name[k]) == 0) ret = main_legacy(k, argc - 1, argv + 1);
	if (ret < 0 && strcmp(argv[1], "batch") == 0) ret = main_batch(argc - 1, argv + 1);
	if (ret < 0) {
		fprintf(stderr, "[main] unrecognized command '%s'\n", argv[1]);
		return 1;
	}
	if (ret == 0) {      /* trailer on stderr (src/main.c:49-55); argv as getopt left it */
		fprintf(stderr, "[main] Version: %s\n", PACKAGE_VERSION);
		fprintf(stderr, "[main] CMD:");
		for (int i = 0; i < argc; ++i) fprintf(stderr, " %s", argv[i]);
		fprintf(stderr, "\n");
	}
	return ret;
}
