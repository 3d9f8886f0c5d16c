/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked into the product library.
 *
 * One translation unit that #includes the UNMODIFIED reference header
 * (<reference>/src/alignment.h, found through -I) and exports a plain C entry
 * point around its five static-inline DP functions:
 *     align_gla               src/alignment.h:417-473
 *     align_local_affine      src/alignment.h:805-847
 *     align_fit_affine_jump   src/alignment.h:596-694
 *     align_overlap           src/alignment.h:926-964
 *     edit_dist               src/alignment.h:291-315
 * The reference prints a debug line from align_fit_affine_jump (:602); printf is
 * routed to a sink for this TU so the harness' stdout stays clean.
 * Built by oracle/Makefile into oracle/_ref/libaligntools_ref.so with
 *   gcc -std=gnu17 -O2 -include oracle/pad.h -I<reference>/src
 */
#include <stdio.h>
#include <stdarg.h>
static int ref_shim_sink(const char *fmt, ...) { (void)fmt; return 0; }
#define printf ref_shim_sink
#include "alignment.h"
#undef printf

enum { REF_GLOBAL = 0, REF_LOCAL = 1, REF_FIT = 2, REF_OVERLAP = 3, REF_EDIT = 4 };

/* Returns 0 on success.  r1/r2 must hold l1+l2+1 bytes.  score_out receives the
 * reference's double (integer valued).  For REF_EDIT only score_out is written. */
int ref_align(int mode, const char *s1, size_t l1, const char *s2, size_t l2,
              int m, int u, int o, int e, int j, int jump,
              const int *sites, size_t n_sites,
              double *score_out, char *r1_out, char *r2_out, size_t *aln_len)
{
	kstring_t ks1, ks2, r1, r2;
	opt_t opt;
	double score = 0;
	memset(&opt, 0, sizeof(opt));
	opt.m = m; opt.u = u; opt.o = o; opt.e = e; opt.j = j;
	opt.s = jump ? true : false;          /* reference enum: true == 0 (:24) */
	opt.sites.size = n_sites;
	opt.sites.pos = (int *)sites;
	ks1.l = l1; ks1.m = l1 + 1; ks1.s = (char *)s1;
	ks2.l = l2; ks2.m = l2 + 1; ks2.s = (char *)s2;
	if (mode == REF_EDIT) {
		*score_out = (double)edit_dist(&ks1, &ks2, &opt);
		if (aln_len) *aln_len = 0;
		return 0;
	}
	memset(&r1, 0, sizeof(r1)); memset(&r2, 0, sizeof(r2));
	r1.s = mycalloc(l1 + l2, char);       /* as the mode drivers do (:507-508) */
	r2.s = mycalloc(l1 + l2, char);
	switch (mode) {
	case REF_GLOBAL:  score = align_gla(&ks1, &ks2, &r1, &r2, &opt); break;
	case REF_LOCAL:   score = align_local_affine(&ks1, &ks2, &r1, &r2, &opt); break;
	case REF_FIT:     score = align_fit_affine_jump(&ks1, &ks2, &r1, &r2, &opt); break;
	case REF_OVERLAP: score = align_overlap(&ks1, &ks2, &r1, &r2, &opt); break;
	default: free(r1.s); free(r2.s); return -1;
	}
	*score_out = score;
	if (aln_len) *aln_len = r1.l;
	memcpy(r1_out, r1.s, r1.l); r1_out[r1.l] = 0;
	memcpy(r2_out, r2.s, r2.l); r2_out[r2.l] = 0;
	free(r1.s); free(r2.s);
	return 0;
}

/* Time `reps` back-to-back calls of the reference DP for one pair (CPU baseline leg). */
int ref_threads(void) { return 1; }   /* the reference is single-threaded */

/* ---- batch driver: the reference DP over many pairs, one pthread per slice.  Used for
 * bulk parity and as the CPU baseline ("kind": "reference").  Each call is exactly the
 * reference's own path including its matrix alloc/free (:119-170). ---- */
#include <pthread.h>
#include <stdint.h>
typedef struct {
	int mode, m, u, o, e, j, jump; size_t lo, hi;
	const char *q; const uint64_t *q_off; const uint32_t *q_len;
	const char *t; const uint64_t *t_off; const uint32_t *t_len;
	const int *sites; const uint64_t *site_off;
	double *score; char *r1, *r2; const uint64_t *aln_off; uint32_t *aln_len; int rc;
} ref_job;

static void *ref_worker(void *arg)
{
	ref_job *b = (ref_job *)arg; size_t p;
	for (p = b->lo; p < b->hi; p++) {
		size_t al = 0, cap = (size_t)b->q_len[p] + b->t_len[p] + 1; int rc;
		char *tmp = NULL, *r1, *r2;
		if (b->r1) { r1 = b->r1 + b->aln_off[p]; r2 = b->r2 + b->aln_off[p]; }
		else { tmp = (char *)malloc(cap * 2); r1 = tmp; r2 = tmp + cap; }
		rc = ref_align(b->mode, b->q + b->q_off[p], b->q_len[p], b->t + b->t_off[p], b->t_len[p],
		               b->m, b->u, b->o, b->e, b->j, b->jump,
		               b->sites ? b->sites + b->site_off[p] : NULL,
		               b->sites ? (size_t)(b->site_off[p + 1] - b->site_off[p]) : 0,
		               &b->score[p], r1, r2, &al);
		if (b->aln_len) b->aln_len[p] = (uint32_t)al;
		free(tmp);
		if (rc) b->rc = rc;
	}
	return NULL;
}

int ref_align_batch(int mode, int m, int u, int o, int e, int j, int jump, size_t n,
                    const char *q, const uint64_t *q_off, const uint32_t *q_len,
                    const char *t, const uint64_t *t_off, const uint32_t *t_len,
                    const int *sites, const uint64_t *site_off,
                    double *score, char *r1, char *r2, const uint64_t *aln_off, uint32_t *aln_len,
                    int n_threads)
{
	pthread_t th[256]; ref_job jobs[256]; int k, rc = 0;
	if (n_threads < 1) n_threads = 1; if (n_threads > 256) n_threads = 256;
	for (k = 0; k < n_threads; k++) {
		ref_job *b = &jobs[k];
		b->mode = mode; b->m = m; b->u = u; b->o = o; b->e = e; b->j = j; b->jump = jump;
		b->lo = n * (size_t)k / n_threads; b->hi = n * (size_t)(k + 1) / n_threads;
		b->q = q; b->q_off = q_off; b->q_len = q_len; b->t = t; b->t_off = t_off; b->t_len = t_len;
		b->sites = sites; b->site_off = site_off; b->score = score;
		b->r1 = r1; b->r2 = r2; b->aln_off = aln_off; b->aln_len = aln_len; b->rc = 0;
		if (n_threads == 1) ref_worker(b); else pthread_create(&th[k], NULL, ref_worker, b);
	}
	for (k = 0; k < n_threads; k++) { if (n_threads > 1) pthread_join(th[k], NULL); if (jobs[k].rc) rc = jobs[k].rc; }
	return rc;
}

/* ---- the reference's FASTA path: kseq (src/kseq.h:189-229) exactly as kstring_read drives it
 * (src/alignment.h:229-237), dumped one record per line as
 *     <name>\t<comment.s or "(null)">\t<strlen(seq)>\t<seq>\n
 * so tests can pin the host's own reader (host/at_fasta.c) on it.  Returns the number of bytes
 * written (output truncated at cap), or -1 when the file cannot be opened. ---- */
long ref_kseq_dump(const char *fname, char *out, size_t cap)
{
	gzFile fp = gzopen(fname, "r");
	if (fp == NULL) return -1;
	kseq_t *seq = kseq_init(fp);
	size_t pos = 0;
	while (kseq_read(seq) >= 0) {
		int n = snprintf(out + pos, pos < cap ? cap - pos : 0, "%s\t%s\t%zu\t%s\n", seq->name.s,
		                 seq->comment.s ? seq->comment.s : "(null)", strlen(seq->seq.s), seq->seq.s);
		if (n < 0 || pos + (size_t)n >= cap) break;
		pos += (size_t)n;
	}
	kseq_destroy(seq);
	gzclose(fp);
	return (long)pos;
}
