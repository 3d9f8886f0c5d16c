/* alignTools.c -- C host of the B200-native alignTools: the reference's command line
 *     alignTools <global|local|fit|overlap|edit> [-m -u -o -e -j -s] <target.fa>
 * (src/main.c:32-57 and the five main_<mode> drivers, src/alignment.h:318-350, 476-517, 698-744,
 * 851-892, 967-1008) kept as a drop-in -- same option strings, usage text, stdout / stderr bytes
 * and exit codes -- with the DP itself running on the GPU through the C-ABI
 * (include/aligntools_b200.h).  The legacy sub-commands call the single-pair entry points that
 * carry the reference's own signatures (at_align_gla, ...); the new `batch` sub-command packs
 * every record pair of its input into one at_batch_* call.  There is no CPU alignment code here:
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
 * `extra` adds the batch-only flags.  Returns 0, or 1 when the driver must return 1. */
static int parse_opts(int mode, int argc, char **argv, at_opt_t *opt, const char *extra, int *tsv, int *gpus)
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
		case 'c': if (!tsv) return 1; *tsv = 1; break;
		case 'g': if (!gpus) return 1; *gpus = atoi(optarg); break;
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
	if (parse_opts(mode, argc, argv, &opt, NULL, NULL, NULL)) return 1;
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
/* alignTools batch <mode> [mode options] [-c] [-g N] <pairs.fa> [<targets.fa>]
 *   one file : records 2k and 2k+1 form pair k (read, target);
 *   two files: record k of the first file is the read, record k of the second the target.
 * Default output: per pair exactly the block the legacy sub-command prints for that pair.
 * -c: one TSV line per pair (read, target, score, beg_i, end_i, beg_j, end_j, CIGAR).
 * -g N: spread the batch over N GPUs (contiguous slices, no communication). */
typedef struct { uint8_t *bytes; size_t n, cap; uint64_t *off; uint32_t *len; char **name; char **comment; size_t cnt, rcap; } seqset;

static void seqset_push(seqset *s, const at_fasta_rec *r, int keep_comment)
{
	if (s->cnt == s->rcap) {
		s->rcap = s->rcap ? 2 * s->rcap : 1024;
		s->off = (uint64_t *)realloc(s->off, s->rcap * sizeof *s->off);
		s->len = (uint32_t *)realloc(s->len, s->rcap * sizeof *s->len);
		s->name = (char **)realloc(s->name, s->rcap * sizeof *s->name);
		s->comment = (char **)realloc(s->comment, s->rcap * sizeof *s->comment);
		if (!s->off || !s->len || !s->name || !s->comment) die("mycalloc failure requesting %d of size %d bytes", (int)s->rcap, 8);
	}
	if (s->n + r->seq_len + 1 > s->cap) {
		while (s->n + r->seq_len + 1 > s->cap) s->cap = s->cap ? 2 * s->cap : 1 << 20;
		s->bytes = (uint8_t *)realloc(s->bytes, s->cap);
		if (!s->bytes) die("mycalloc failure requesting %d of size %d bytes", (int)s->cap, 1);
	}
	memcpy(s->bytes + s->n, r->seq, r->seq_len);
	s->off[s->cnt] = s->n; s->len[s->cnt] = (uint32_t)r->seq_len;
	s->name[s->cnt] = xstrdup(r->name);
	s->comment[s->cnt] = (keep_comment && r->comment) ? xstrdup(r->comment) : NULL;
	s->n += r->seq_len; s->cnt++;
}

static void batch_usage(void)
{
	fprintf(stderr, "\n");
	fprintf(stderr, "Usage:   alignTools batch <global|local|fit|overlap|edit> [options] <pairs.fa> [<targets.fa>]\n\n");
	fprintf(stderr, "Options: -m -u -o -e (-j -s for fit) as in the single-pair commands\n");
	fprintf(stderr, "         -c       one TSV line per pair with the CIGAR instead of the text blocks\n");
	fprintf(stderr, "         -g INT   number of GPUs [1]\n");
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
	int tsv = 0, gpus = 1;
	if (parse_opts(mode, argc - 1, argv + 1, &opt, "cg:", &tsv, &gpus)) return 1;
	char **files = argv + 1 + optind;
	const int n_files = argc - 1 - optind;
	if (n_files < 1 || n_files > 2) { batch_usage(); return 1; }
	const int jump = mode == M_FIT && opt.s == JUMP_ON;

	seqset reads, targets;
	memset(&reads, 0, sizeof reads); memset(&targets, 0, sizeof targets);
	for (int fi = 0; fi < n_files; ++fi) {
		at_fasta *fa = at_fasta_open(files[fi]);
		if (!fa) die("Can't open %s\n", files[fi]);
		at_fasta_rec rec;
		size_t k = 0;
		while (at_fasta_next(fa, &rec)) {
			const int is_target = n_files == 2 ? fi == 1 : (int)(k & 1);
			seqset_push(is_target ? &targets : &reads, &rec, is_target);
			++k;
		}
		at_fasta_close(fa);
	}
	if (reads.cnt == 0 || reads.cnt != targets.cnt) die("read_kstring: fail to read sequence");
	const size_t n = reads.cnt;

	/* junction sites: one list per pair from the target record's own comment */
	int32_t *sites = NULL; uint64_t *site_off = NULL;
	if (jump) {
		size_t tot = 0, cap = 0;
		site_off = (uint64_t *)xcalloc(n + 1, sizeof *site_off);
		for (size_t k = 0; k < n; ++k) {
			if (!targets.comment[k]) die("fail to read junction sites");
			int *pos = NULL;
			const size_t ns = at_parse_sites(targets.comment[k], &pos);
			if (tot + ns + 1 > cap) { cap = 2 * (tot + ns + 1); sites = (int32_t *)realloc(sites, cap * sizeof *sites); if (!sites) die("mycalloc failure requesting %d of size %d bytes", (int)cap, 4); }
			for (size_t x = 0; x < ns; ++x) sites[tot + x] = pos[x];
			free(pos);
			tot += ns; site_off[k + 1] = tot;
		}
		if (!sites) sites = (int32_t *)xcalloc(1, sizeof *sites);
	}
	if (mode == M_FIT)
		for (size_t k = 0; k < n; ++k)
			if (reads.len[k] > targets.len[k]) die("first sequence must be shorter than the second\n");

	int devs[64];
	if (gpus < 1) gpus = 1;
	if (gpus > 64) gpus = 64;
	for (int k = 0; k < gpus; ++k) devs[k] = k;
	at_handle *h = NULL;
	int rc = at_create(devs, gpus, &h);
	if (rc) die("aligntools-b200: %s", at_strerror(rc));
	at_params prm = {opt.m, opt.u, opt.o, opt.e, opt.j, jump};
	at_batch_input in;
	memset(&in, 0, sizeof in);
	in.n_pairs = n; in.encoding = AT_SEQ_BYTES;
	in.q = reads.bytes; in.q_off = reads.off; in.q_len = reads.len;
	in.t = targets.bytes; in.t_off = targets.off; in.t_len = targets.len;
	in.sites = sites; in.site_off = site_off;
	const uint32_t flags = mode == M_EDIT ? 0u : (tsv ? AT_OUT_CIGAR : AT_OUT_ALN);
	at_batch *b = NULL;
	rc = at_batch_create(h, mode, &prm, &in, flags, &b);
	if (!rc) rc = at_batch_run(b, NULL);
	if (rc) die("%s (%s)", at_strerror(rc), at_last_error(h));
	uint64_t n_ops = 0, n_cols = 0;
	at_batch_sizes(b, &n_ops, &n_cols);
	at_batch_output out;
	memset(&out, 0, sizeof out);
	out.score = (int32_t *)xcalloc(n, sizeof(int32_t));
	out.end_i = (uint32_t *)xcalloc(n, 4); out.end_j = (uint32_t *)xcalloc(n, 4);
	out.beg_i = (uint32_t *)xcalloc(n, 4); out.beg_j = (uint32_t *)xcalloc(n, 4);
	if (flags & AT_OUT_CIGAR) { out.cigar = (uint32_t *)xcalloc(n_ops + 1, 4); out.cigar_cap = n_ops + 1; out.cigar_off = (uint64_t *)xcalloc(n + 1, 8); }
	if (flags & AT_OUT_ALN) { out.aln1 = (char *)xcalloc(n_cols + 1, 1); out.aln2 = (char *)xcalloc(n_cols + 1, 1); out.aln_cap = n_cols + 1; out.aln_off = (uint64_t *)xcalloc(n + 1, 8); }
	rc = at_batch_fetch(b, &out);
	if (rc) die("%s (%s)", at_strerror(rc), at_last_error(h));
	at_batch_free(b);

	char *cig = NULL; size_t cig_cap = 0;
	for (size_t k = 0; k < n; ++k) {
		if (tsv) {
			const char *cs = "*";
			if (mode != M_EDIT) {
				const uint64_t o0 = out.cigar_off[k], o1 = out.cigar_off[k + 1];
				const size_t need = 12 * (size_t)(o1 - o0) + 2;
				if (need > cig_cap) { cig_cap = 2 * need; cig = (char *)realloc(cig, cig_cap); if (!cig) die("mycalloc failure requesting %d of size %d bytes", (int)cig_cap, 1); }
				if (o1 > o0) { at_cigar_to_string(out.cigar + o0, o1 - o0, cig, cig_cap); cs = cig; }
			}
			printf("%s\t%s\t%d\t%u\t%u\t%u\t%u\t%s\n", reads.name[k], targets.name[k], out.score[k],
			       out.beg_i[k], out.end_i[k], out.beg_j[k], out.end_j[k], cs);
			continue;
		}
		if (mode == M_EDIT) { printf("edit_distance=%d\n", out.score[k]); continue; }
		if (mode == M_FIT) {
			if (jump) printf("%s\n", targets.comment[k]);
			printf("asDAsdaSDAsdasDAsdaSD\n");
		}
		if (mode == M_OVERLAP) printf("%f\n", (double)out.score[k]);
		else printf("score=%f\n", (double)out.score[k]);
		const uint64_t a0 = out.aln_off[k], a1 = out.aln_off[k + 1];
		fwrite(out.aln1 + a0, 1, (size_t)(a1 - a0), stdout); fputc('\n', stdout);
		fwrite(out.aln2 + a0, 1, (size_t)(a1 - a0), stdout); fputc('\n', stdout);
	}
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
