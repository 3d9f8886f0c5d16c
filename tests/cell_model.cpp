// cell_model.cpp -- TEST INFRASTRUCTURE: host-side model of the fill kernels' arithmetic.
//
// Compiles aligntools/c_b200/csrc/at_cell.cuh (the cell update K1 / K2 run on the GPU: tagged values, pointer
// nibbles assembled as P - Q + bias) with g++ and drives it over whole matrices in plain column order -- no warp
// geometry -- then walks the stored nibbles with K3's rules.  tests/test_cell_model.py compares the result with
// the oracle on random cases of every mode, on int32 and packed s16x2 lanes, so the recurrence, the tie rules and
// the pointer algebra are pinned on the CPU before any GPU time is spent.  Not part of the product.
//   g++ -O2 -shared -fPIC -o tests/_build/libcell_model.so tests/cell_model.cpp
#include "../aligntools/c_b200/csrc/at_cell.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

using namespace atb2;

enum { MODE_GLOBAL = 0, MODE_LOCAL = 1, MODE_FIT = 2 };
enum { ST_LOW = 0, ST_MID = 1, ST_UPP = 2, ST_JUMP = 3 };

struct PairIO {
	const uint8_t *s1, *s2; int l1, l2;
	const int *sites; int n_sites;
	int score, end_i, end_j, end_state;
	std::vector<uint8_t> nib, jbit;      // (l1 + 1) x (l2 + 1)
};

template <int MODE, bool JUMP, bool PACKED>
static void fill(PairIO *pp /* 1 or 2 pairs, same l2 */, int m, int u, int o, int e, int jp, int garbage_head)
{
	typedef Lanes<PACKED> V;
	typedef typename V::T T;
	constexpr int NP = PACKED ? 2 : 1;
	constexpr uint32_t SPW = V::STEPS_PER_WORD;
	const int l2 = pp[0].l2;
	int l1 = 0;
	for (int h = 0; h < NP; ++h) { if (pp[h].l1 > l1) l1 = pp[h].l1; pp[h].nib.assign((size_t)(pp[h].l1 + 1) * (l2 + 1), 0xff); pp[h].jbit.assign((size_t)(pp[h].l1 + 1) * (l2 + 1), 0); }
	CellConst<PACKED> c;
	c.set(o, e, jp, cell_k_and<PACKED>(), cell_k_or<PACKED>());
	const T ZERO = V::value(0), NEGV = PACKED ? (T)((uint32_t)(AT_NEG16 * 0x10001) + 0x80008000u) : (T)AT_NEG;
	std::vector<RowState<PACKED, JUMP>> row(l1 + 1);
	std::vector<T> lcol0(l1 + 1);      // L(i, 0): what row i hands down in column 0 (only used through lup of column >= 1? no: L flows within a column)
	// column 0 (left border) -- reference: :432-436 global, calloc zeros local, :612-617 fit
	for (int i = 1; i <= l1; ++i) {
		RowState<PACKED, JUMP> &st = row[i];
		if (MODE == MODE_GLOBAL)     { st.mo = NEGV | V::rep(3); st.u = NEGV | V::rep(1); st.h = V::value(o + e * i) | V::rep(TAG_L); }
		else if (MODE == MODE_LOCAL) { st.mo = ZERO + c.o_m + V::rep(TAG_M); st.u = ZERO | V::rep(1); st.h = ZERO | V::rep(TAG_L); }
		else                         { st.mo = NEGV | V::rep(3); st.u = NEGV | V::rep(1); st.h = NEGV | V::rep(TAG_M); }
		st.j = NEGV; st.x = st.xj = 0;
	}
	(void)garbage_head;
	// results
	long best_key[2] = {-1, -1}; int best_i[2] = {0, 0}, best_j[2] = {0, 0};
	long capM[2] = {-(1L << 40), -(1L << 40)}, capL[2] = {-(1L << 40), -(1L << 40)}; int capMj[2] = {0, 0}, capLj[2] = {0, 0};
	auto half = [&](T v, int h) -> long { return PACKED ? (long)(((uint32_t)v >> (16 * h)) & 0xffffu) - 0x8000 : (long)(int32_t)v; };
	for (int j = 1; j <= l2; ++j) {
		// matrix row 0 at columns j-1 (diagonal of row 1) and j
		T d, lup, mo_up;
		auto h_row0 = [&](int col) -> T {
			if (MODE == MODE_GLOBAL) return col == 0 ? (V::value(o < 0 ? 0 : o) | V::rep(o < 0 ? TAG_M : TAG_L)) : (V::value(o + e * col) | V::rep(TAG_U));   // :437-441
			if (MODE == MODE_LOCAL) return ZERO | V::rep(TAG_L);
			return ZERO | V::rep(TAG_M);                                                             // :619-624: M[0][j] = U[0][j] = 0
		};
		d = h_row0(j - 1);
		if (MODE == MODE_GLOBAL)     { lup = NEGV | V::rep(3); mo_up = NEGV | V::rep(3); }
		else if (MODE == MODE_LOCAL) { lup = ZERO | V::rep(3); mo_up = ZERO + c.o_m + V::rep(TAG_M); }
		else                         { lup = NEGV | V::rep(3); mo_up = ZERO + c.o_m + V::rep(TAG_M); }
		T jadd = 0;
		if (JUMP) {
			bool barred = false;
			for (int k = 0; k < pp[0].n_sites; ++k) if (pp[0].sites[k] == j - 1) barred = true;   // blacklist (:659)
			jadd = barred ? c.j_barred : c.j_enter;
		}
		const uint32_t mul = ((j - 1) % SPW == 0 && PACKED) ? 0u : 16u, mulj = 2u;
		for (int i = 1; i <= l1; ++i) {
			// substitution score: raw byte equality (:449, :632, :824)
			int s[2];
			for (int h = 0; h < NP; ++h) s[h] = (i <= pp[h].l1 && pp[h].s1[i - 1] == pp[h].s2[j - 1]) ? m : u;
			T pw;
			if (PACKED) pw = (MODE == MODE_LOCAL) ? (T)(((uint32_t)(8 * s[0]) & 0xffffu) | ((uint32_t)(8 * s[1]) << 16)) : (T)(8 * s[0] + 8 * s[1] * 65536);
			else pw = (T)(8 * s[0]);
			CellOut<PACKED> out;
			d = cell_update<MODE == MODE_LOCAL, JUMP, PACKED, true>(c, row[i], d, pw, lup, mo_up, jadd, mul, mulj, out);
			lup = out.lk; mo_up = out.mo;
			// nibble of this cell: low nibble of the running word
			const uint32_t w = ptr_word(row[i].x);
			for (int h = 0; h < NP; ++h) {
				if (i > pp[h].l1) continue;
				pp[h].nib[(size_t)i * (l2 + 1) + j] = (uint8_t)((w >> (16 * h)) & 15u);
				if (JUMP) pp[h].jbit[(size_t)i * (l2 + 1) + j] = (uint8_t)(jump_word(row[i].xj) & 1u);
				const long mk = half(out.mk, h) >> 3, lk = half(out.lk, h) >> 3;
				if (MODE == MODE_LOCAL) {          // running first maximum in row-major order (:830-833): larger score, then smaller row, then smaller column
					const long key = mk;
					if (key > best_key[h] || (key == best_key[h] && i < best_i[h])) { best_key[h] = key; best_i[h] = i; best_j[h] = j; }
				}
				if (MODE == MODE_FIT && i == pp[h].l1 && j < l2) {      // column l2 excluded (:677, :684)
					if (mk > capM[h]) { capM[h] = mk; capMj[h] = j; }
					if (lk > capL[h]) { capL[h] = lk; capLj[h] = j; }
				}
				if (MODE == MODE_GLOBAL && i == pp[h].l1 && j == l2) { pp[h].score = (int)(half(out.h, h) >> 3); pp[h].end_state = 3 - (int)(half(out.h, h) & 3); pp[h].end_i = i; pp[h].end_j = j; }
			}
		}
	}
	for (int h = 0; h < NP; ++h) {
		if (MODE == MODE_LOCAL) { pp[h].score = (int)best_key[h]; pp[h].end_i = best_i[h]; pp[h].end_j = best_j[h]; pp[h].end_state = ST_MID; }
		if (MODE == MODE_FIT) {
			const bool useL = capL[h] > capM[h];       // L replaces M only when strictly greater (:685)
			pp[h].score = (int)(useL ? capL[h] : capM[h]); pp[h].end_i = pp[h].l1; pp[h].end_j = useL ? capLj[h] : capMj[h];
			pp[h].end_state = useL ? ST_LOW : ST_MID;
		}
	}
}

// K3's walk (at_kernels.cuh) over the nibble matrix; ops out in forward order: 'M' 'I' 'D' 'N'
static int walk(const PairIO &p, int mode, int jump, char *ops, int *beg_i, int *beg_j)
{
	int i = p.end_i, j = p.end_j, state = p.end_state, n = 0;
	const int l2 = p.l2;
	std::vector<char> rev;
	bool home = false;
	for (;;) {
		const bool go = mode == MODE_FIT ? (i > 0) : (i > 0 && j > 0);
		if (!go || home) break;
		if (j == 0 && state != ST_LOW) break;
		const uint32_t nb = p.nib[(size_t)i * (l2 + 1) + j];
		if (state == ST_LOW)      { state = (nb & 4u) ? ST_MID : ST_LOW; --i; rev.push_back('I'); }
		else if (state == ST_MID) { const uint32_t pm = nb & 3u; --i; --j; rev.push_back('M'); if (pm == 3 && mode == MODE_LOCAL) home = true; else state = (int)pm; }
		else if (state == ST_UPP) { state = (nb & 8u) ? ST_UPP : ST_MID; --j; rev.push_back('D'); }
		else                      { state = (jump && p.jbit[(size_t)i * (l2 + 1) + j]) ? ST_JUMP : ST_MID; --j; rev.push_back('N'); }
	}
	*beg_i = i; *beg_j = j;
	if (mode == MODE_GLOBAL) { while (j > 0) { --j; rev.push_back('D'); } while (i > 0) { --i; rev.push_back('I'); } }
	for (size_t k = rev.size(); k-- > 0;) ops[n++] = rev[k];
	return n;
}


// ---- overlap (single plane, lin_update) ----
static void fill_overlap(PairIO &p, int m, int u, int o)
{
	const int l1 = p.l1, l2 = p.l2, gap = 4 * o;
	p.nib.assign((size_t)(l1 + 1) * (l2 + 1), 0xff);
	std::vector<LinRow> row(l1 + 1);
	for (int i = 1; i <= l1; ++i) { row[i].a2 = gap + 2; row[i].x = 0; }      // M[i][0] = 0 (:938): A = 4 o
	long capV = gap; int capJ = 0;                                            // M[l1][0] = 0 seeds the search (:954-959)
	for (int j = 1; j <= l2; ++j) {
		int d2 = (j == 1 ? gap : AT_NEGL) + 2;                                // A(0, j-1): M[0][0] = 0, M[0][j] = -inf (:937)
		int a_up = AT_NEGL;
		for (int i = 1; i <= l1; ++i) {
			const int s = p.s1[i - 1] == p.s2[j - 1] ? m : u;
			int d2n;
			a_up = lin_update(row[i], d2, 4 * (s - o) - 1, a_up, gap, d2n);
			d2 = d2n;
			p.nib[(size_t)i * (l2 + 1) + j] = (uint8_t)(lin_word(row[i].x) & 3u);
			if (i == l1 && j < l2 && a_up > capV) { capV = a_up; capJ = j; }  // column l2 excluded (:955)
		}
	}
	p.score = (int)((capV - gap) / 4); p.end_i = l1; p.end_j = capJ; p.end_state = ST_MID;
}

static int walk_overlap(const PairIO &p, char *ops, int *beg_i, int *beg_j)
{
	int i = p.end_i, j = p.end_j, n = 0;
	std::vector<char> rev;
	while (j > 0) {                                                           // :899
		if (i == 0) break;
		const uint32_t c = p.nib[(size_t)i * (p.l2 + 1) + j];
		if (c & 2u)      { --i; rev.push_back('I'); }
		else if (c & 1u) { --i; --j; rev.push_back('M'); }
		else             { --j; rev.push_back('D'); }
	}
	*beg_i = i; *beg_j = j;
	for (size_t k = rev.size(); k-- > 0;) ops[n++] = rev[k];
	return n;
}

// One pair (packed = 0) or two pairs sharing l2 (packed = 1; local only).  ops_* must hold l1 + l2 bytes.
// out[h] = {score, end_i, end_j, end_state, beg_i, beg_j, n_ops}
extern "C" int cell_model_run(int mode, int jump, int packed, const uint8_t *s1a, int l1a, const uint8_t *s1b, int l1b,
                              const uint8_t *s2a, const uint8_t *s2b, int l2, int m, int u, int o, int e, int jp,
                              const int *sites, int n_sites, int *out /* [2][7] */, char *ops_a, char *ops_b)
{
	PairIO pp[2];
	pp[0].s1 = s1a; pp[0].l1 = l1a; pp[0].s2 = s2a; pp[0].l2 = l2; pp[0].sites = sites; pp[0].n_sites = n_sites;
	pp[1].s1 = s1b; pp[1].l1 = l1b; pp[1].s2 = s2b; pp[1].l2 = l2; pp[1].sites = sites; pp[1].n_sites = n_sites;
	if (mode == 3) {
		fill_overlap(pp[0], m, u, o);
		int bi = 0, bj = 0;
		const int n = walk_overlap(pp[0], ops_a, &bi, &bj);
		out[0] = pp[0].score; out[1] = pp[0].end_i; out[2] = pp[0].end_j; out[3] = pp[0].end_state; out[4] = bi; out[5] = bj; out[6] = n;
		return 0;
	}
	if (packed) {
		if (jump) return -1;
		if (mode == MODE_LOCAL) fill<MODE_LOCAL, false, true>(pp, m, u, o, e, jp, 0);
		else if (mode == MODE_GLOBAL) fill<MODE_GLOBAL, false, true>(pp, m, u, o, e, jp, 0);
		else fill<MODE_FIT, false, true>(pp, m, u, o, e, jp, 0);
	} else if (mode == MODE_GLOBAL) fill<MODE_GLOBAL, false, false>(pp, m, u, o, e, jp, 0);
	else if (mode == MODE_LOCAL) fill<MODE_LOCAL, false, false>(pp, m, u, o, e, jp, 0);
	else if (mode == MODE_FIT && jump) fill<MODE_FIT, true, false>(pp, m, u, o, e, jp, 0);
	else if (mode == MODE_FIT) fill<MODE_FIT, false, false>(pp, m, u, o, e, jp, 0);
	else return -1;
	for (int h = 0; h < (packed ? 2 : 1); ++h) {
		int bi = 0, bj = 0;
		const int n = walk(pp[h], mode, jump, h ? ops_b : ops_a, &bi, &bj);
		int *r = out + 7 * h;
		r[0] = pp[h].score; r[1] = pp[h].end_i; r[2] = pp[h].end_j; r[3] = pp[h].end_state; r[4] = bi; r[5] = bj; r[6] = n;
	}
	return 0;
}
