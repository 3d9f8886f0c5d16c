// at_shim.cu -- section B of include/aligntools_b200.h: single-pair entry points with the
// reference's own signatures (src/alignment.h:292, 418, 597, 806, 926) so that the five
// main_<mode> drivers can call the B200 path unchanged.  Errors follow die()
// (src/alignment.h:69-79): "FATAL ERROR: ..." on stderr, exit(-1).
#include "../../../include/aligntools_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

static at_handle *g_handle = nullptr;
static std::once_flag g_once;

static void shim_die(const char *fmt, ...)
{
	va_list ap; va_start(ap, fmt);
	fprintf(stderr, "FATAL ERROR: "); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n");
	va_end(ap);
	exit(-1);
}

static at_handle *shim_handle()
{
	std::call_once(g_once, [] {
		int dev = 0;
		if (const char *e = getenv("AT_DEVICE")) dev = atoi(e);
		int rc = at_create(&dev, 1, &g_handle);
		if (rc) shim_die("aligntools-b200: %s", at_strerror(rc));
	});
	return g_handle;
}

static double shim_align(int mode, at_kstring_t *s1, at_kstring_t *s2, at_kstring_t *r1, at_kstring_t *r2, at_opt_t *opt, bool want_aln)
{
	at_handle *h = shim_handle();
	at_params p;
	p.m = opt->m; p.u = opt->u; p.o = opt->o; p.e = opt->e; p.j = opt->j;
	p.jump = (mode == AT_FIT && opt->s == 0) ? 1 : 0;      /* `true` is 0 in the reference's enum (:24) */
	const uint64_t q_off = 0, t_off = 0;
	const uint32_t q_len = (uint32_t)s1->l, t_len = (uint32_t)s2->l;
	uint64_t site_off[2] = {0, p.jump ? (uint64_t)opt->sites.size : 0};
	int32_t dummy_site = -1;
	at_batch_input in;
	memset(&in, 0, sizeof in);
	in.n_pairs = 1; in.encoding = AT_SEQ_BYTES;
	in.q = (const uint8_t *)s1->s; in.q_off = &q_off; in.q_len = &q_len;
	in.t = (const uint8_t *)s2->s; in.t_off = &t_off; in.t_len = &t_len;
	if (p.jump) { in.sites = opt->sites.size ? (const int32_t *)opt->sites.pos : &dummy_site; in.site_off = site_off; }
	int32_t score = 0;
	uint64_t aln_off[2] = {0, 0};
	std::vector<char> a1, a2;
	at_batch_output out;
	memset(&out, 0, sizeof out);
	out.score = &score;
	if (want_aln) {
		a1.resize((size_t)q_len + t_len + 1); a2.resize((size_t)q_len + t_len + 1);
		out.aln1 = a1.data(); out.aln2 = a2.data(); out.aln_cap = a1.size(); out.aln_off = aln_off;
	}
	int rc = at_batch_align(h, mode, &p, &in, want_aln ? AT_OUT_ALN : 0, &out, nullptr);
	if (rc == AT_E_FITLEN) shim_die("first sequence must be shorter than the second to do fitting alignment");
	if (rc) shim_die("%s (%s)", at_strerror(rc), at_last_error(h));
	if (want_aln) {
		const size_t n = (size_t)(aln_off[1] - aln_off[0]);
		/* ownership as strrev (:176-183): the caller's block is freed and replaced */
		free(r1->s); free(r2->s);
		r1->s = (char *)calloc(n + 1, 1); r2->s = (char *)calloc(n + 1, 1);
		if (!r1->s || !r2->s) shim_die("mycalloc failure requesting %d of size %d bytes", (int)(n + 1), 1);
		memcpy(r1->s, a1.data(), n); memcpy(r2->s, a2.data(), n);
		r1->l = r2->l = n; r1->m = r2->m = n + 1;
	}
	return (double)score;
}

extern "C" double at_align_gla(at_kstring_t *s1, at_kstring_t *s2, at_kstring_t *r1, at_kstring_t *r2, at_opt_t *opt)
{
	if (!s1 || !s2 || !r1 || !r2 || !opt) shim_die("align: parameter error\n");
	return shim_align(AT_GLOBAL, s1, s2, r1, r2, opt, true);
}
extern "C" double at_align_local_affine(at_kstring_t *s1, at_kstring_t *s2, at_kstring_t *r1, at_kstring_t *r2, at_opt_t *opt)
{
	if (!s1 || !s2 || !r1 || !r2 || !opt) shim_die("align: parameter error\n");
	return shim_align(AT_LOCAL, s1, s2, r1, r2, opt, true);
}
extern "C" double at_align_fit_affine_jump(at_kstring_t *s1, at_kstring_t *s2, at_kstring_t *r1, at_kstring_t *r2, at_opt_t *opt)
{
	if (!s1 || !s2 || !r1 || !r2 || !opt) shim_die("align: parameter error\n");
	if (s1->l > s2->l) shim_die("first sequence must be shorter than the second to do fitting alignment");
	return shim_align(AT_FIT, s1, s2, r1, r2, opt, true);
}
extern "C" double at_align_overlap(at_kstring_t *s1, at_kstring_t *s2, at_kstring_t *r1, at_kstring_t *r2, at_opt_t *opt)
{
	if (!s1 || !s2 || !r1 || !r2 || !opt) shim_die("align_overlap: parameter error\n");
	return shim_align(AT_OVERLAP, s1, s2, r1, r2, opt, true);
}
extern "C" int at_edit_dist(at_kstring_t *s1, at_kstring_t *s2, at_opt_t *opt)
{
	if (!s1 || !s2 || !opt) shim_die("edit_dist: parameter error\n");
	return (int)shim_align(AT_EDIT, s1, s2, nullptr, nullptr, opt, false);
}
