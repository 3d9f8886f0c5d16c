// at_fill_affine.cuh -- K1: Gotoh M/L/U(/J) fill for global / local / fit (+jump), sm_100a.
// (included by at_kernels.cuh after the constants)
//
// Inter-pair kernel for SHORT reads (l1 <= 256 rows = 32 lanes x R <= 8 rows); longer reads go
// to the stripe-pipelined K2 (at_wavefront.cuh).  One kernel template, two kinds of score lanes:
//   Lanes<false>  int32   : one pair per warp, every mode.
//   Lanes<true>   s16x2   : TWO pairs per warp, pair A in bits 0-15 and pair B in bits 16-31 of
//                           every register (VIMNMX.U16x2 / VIADDMNMX.U16x2 DPX instructions);
//                           local mode, single stripe, both pairs share l2.
//
// Geometry: lane k owns R consecutive rows; at step t it works on column j = t - k (anti-diagonal
// of R-row blocks).  Per step each lane hands the last row of its strip -- M+o, L, H+m and the
// 2-bit argmax code of H = max(L,M,U[,J]) -- to lane k+1 with four __shfl_up_sync.  Target symbols
// are staged through a per-warp shared-memory ring (one 32-bit entry per column).
//
// Arithmetic (what makes the instruction mix cheap; ncu: the ALU pipe is the binding unit):
//   * scores are kept x8; a and b being multiples of 8, a != b  =>  |a-b| >= 8, so a traceback
//     flag is  min(max(a,b) - b, 1|3|4|8)  : one subtract (FMA-pipe IMAD.IADD) + one VIMNMX, and
//     it lands on its bit of the pointer nibble without shifts;
//     (packed lanes WITH the query profile -- the C2 path -- go further and carry every argmax in
//     the spare low bits of the values themselves: see TAG below, one VIADDMNMX per flag);
//   * packed lanes are BIASED by 0x8000 per half and compared unsigned, so every add/subtract is
//     an ordinary 32-bit integer instruction (no carry can cross the halves while the values
//     stay in range, which the host checks) and can be issued on the FMA pipe;
//   * PROF variant (target alphabet of the shard has at most 4 distinct bytes -- DNA): a per-warp QUERY
//     PROFILE in shared memory, rebuilt per job: word[comb][lane][r] = 8*s(read row, target symbol) for
//     pair A (+ the same for pair B << 16, comb = codeA*4 + codeB), so  M = H(i-1,j-1) + word  is one
//     LDS (64-bit, two rows at a time) and one add; the target ring holds the byte offset of `comb`;
//   * fallback variant (any byte alphabet): symbols are pre-shifted (<<8 / <<16), so
//     min(a ^ b, 8|m-u|)  is 0 on a match and the whole substitution penalty otherwise;
//     M = (H+m) - that;
//   * local mode keeps the lane's running maximum of  8*M + (7 - row_in_lane)  with one
//     VIADDMNMX per cell: larger score first, then the smaller row -- together with "first step
//     at which the key reached its final value" this is the reference's first maximum in
//     row-major order (src/alignment.h:830-833).
//
// Reference recurrences: src/alignment.h:451-462 (global), :635-667 (fit), :825-841 (local);
// tie rules SURVEY.md A.0:  M: first strictly greater in order L,M,U,(J|HOME);  L: extend wins
// ties;  U: open wins ties;  J: enter wins ties, entering forbidden on listed target indices.
#pragma once

namespace atb2 {

#define AT_RING 512        // target ring entries per warp (two 256-column blocks)
#define AT_FILL_WARPS 4

// query-profile geometry: words per lane (>= R, even, half of it odd: conflict-free 64-bit loads)
__host__ __device__ constexpr int prof_lane_stride(int R) { return R <= 2 ? 2 : (R <= 6 ? 6 : 10); }
// K1 only: an odd R keeps its rows unpadded and reads them with 32-bit loads (an odd lane stride is
// conflict-free too); the smaller table lets a fifth CTA fit an SM for R = 5
__host__ __device__ constexpr int k1_lane_stride(int R) { return (R & 1) ? R : prof_lane_stride(R); }
__host__ __device__ constexpr int prof_combs(bool packed) { return packed ? 16 : 4; }
// dynamic shared memory of one CTA of at_fill_affine<.., R, .., PACKED, PROF>
__host__ __device__ constexpr size_t fill_smem_bytes(int R, bool packed, bool prof)
{
	return (size_t)AT_FILL_WARPS * (prof ? (size_t)AT_RING * 2 + (size_t)prof_combs(packed) * 32 * k1_lane_stride(R) * 4
	                                     : (size_t)AT_RING * 4);
}

template <bool PACKED> struct Lanes;

template <> struct Lanes<false> {
	typedef int32_t T;
	static constexpr int CSHIFT = 16;
	static constexpr uint32_t STEPS_PER_WORD = 8;
	__device__ __forceinline__ static T vmax(T a, T b) { return max(a, b); }
	__device__ __forceinline__ static T vmax3(T a, T b, T c) { return __vimax3_s32(a, b, c); }
	__device__ __forceinline__ static T flag(T d, uint32_t k) { return (T)min((uint32_t)d, k); }     // d >= 0
	__device__ __forceinline__ static T addmax(T a, T b, T c) { return __viaddmax_s32(a, b, c); }
	__device__ __forceinline__ static T delta(int v) { return 8 * v; }
	__device__ __forceinline__ static T delta_h(int v) { return 8 * v; }
	__device__ __forceinline__ static T value(int v) { return 8 * v; }
	__device__ __forceinline__ static T raw(int v) { return v; }
};

template <> struct Lanes<true> {
	typedef uint32_t T;
	static constexpr int CSHIFT = 8;
	static constexpr uint32_t STEPS_PER_WORD = 4;
	__device__ __forceinline__ static T vmax(T a, T b) { return __vmaxu2(a, b); }
	__device__ __forceinline__ static T vmax3(T a, T b, T c) { return __vimax3_u16x2(a, b, c); }
	__device__ __forceinline__ static T flag(T d, uint32_t k) { return __vminu2(d, k * 0x10001u); }
	__device__ __forceinline__ static T addmax(T a, T b, T c) { return __viaddmax_u16x2(a, b, c); }
	__device__ __forceinline__ static T delta(int v) { return (uint32_t)(8 * v * 0x10001); }           // exact 32-bit sum of both halves: for IADD/IMAD
	__device__ __forceinline__ static T delta_h(int v) { return ((uint32_t)(8 * v) & 0xffffu) * 0x10001u; } // two's complement per half: for VIADDMNMX.U16x2
	__device__ __forceinline__ static T value(int v) { return (uint32_t)(8 * v * 0x10001) + 0x80008000u; }
	__device__ __forceinline__ static T raw(int v) { return (uint32_t)(v * 0x10001); }
};

struct FillJob { uint32_t a, b; };     // pair indices; b == a for int32 lanes or a packed job without partner

struct FillArgs2 {
	const uint8_t  *q;       const uint64_t *q_off;  const uint32_t *q_len;
	const uint8_t  *t;       const uint64_t *t_off;  const uint32_t *t_len;
	const uint8_t  *jmask;   // fit+jump: 1 where entering J is forbidden; indexed like t
	const uint8_t  *symmap;  // PROF: byte -> code 0..3 of the shard's target alphabet (256 entries)
	uint32_t        syms;    // PROF: the byte of code c in bits 8c..8c+7
	const FillJob  *jobs;
	uint32_t        n_jobs;
	uint32_t       *counter;
	uint32_t       *ptr;     const uint64_t *ptr_off;   uint32_t pair_base;
	int32_t        *score;   uint32_t *end_i;  uint32_t *end_j;  uint8_t *end_state;
	int             m, u, o, e, jp;
	int             want_ptr;
};

template <int MODE, int R, bool JUMP, bool PACKED, bool PROF>
__global__ void __launch_bounds__(32 * AT_FILL_WARPS) at_fill_affine(const FillArgs2 a)
{
	typedef Lanes<PACKED> V;
	typedef typename V::T T;
	static_assert(!PACKED || (MODE == MODE_LOCAL && !JUMP), "packed lanes: local mode");
	constexpr uint32_t SPW = V::STEPS_PER_WORD;
	// TAG (packed lanes + query profile): the argmax of every max rides in the three spare low bits of the
	// x8-scaled values instead of being derived with subtract + min pairs.  Candidates carry a tie-break tag
	// into the max (one VIADDMNMX each), one LOP3 per propagated value drops it again:
	//   H' = clean | code of the winner (L 2, M 1, U 0: L wins ties, then M -- the reference's order)
	//   L' = clean | 2,  U' = clean,  Mo' = M + o + 2
	//   Ln_t = max(L' + e + 4, Mo')      low bits 6: extended (ties included, :456), 2: opened
	//   Un_t = max(U' + e, Mo')          low bits 0: extended, 2: opened (ties included, :460)
	//   Mn_t = max(H'diag + s + 1, ZERO) low bits 0: HOME (0.0 strictly greater, :825), else diagonal code + 1
	// The pointer nibble is the complement of (Mn_t & 3) | (Ln_t & 4) | (Un_t & 2) << 2; words are inverted at the store.
	constexpr bool TAG = PACKED && PROF;
	constexpr int RPP = 32 * R;
	constexpr int LS = k1_lane_stride(R), NC = prof_combs(PACKED);
	constexpr uint32_t COMB_BYTES = 32u * LS * 4u;             // one comb's slice of the profile
	constexpr size_t WARP_BYTES = fill_smem_bytes(R, PACKED, PROF) / AT_FILL_WARPS;

	extern __shared__ __align__(16) unsigned char fill_smem[];
	const int lane = threadIdx.x & 31;
	unsigned char *warp_smem = fill_smem + (threadIdx.x >> 5) * WARP_BYTES;
	uint32_t *ring = (uint32_t *)warp_smem;                                   // !PROF: one u32 per column
	uint16_t *ring16 = (uint16_t *)warp_smem;                                 // PROF: byte offset of the column's comb (| blacklist bit)
	uint32_t *prof = (uint32_t *)(warp_smem + AT_RING * 2);                   // PROF: [NC][32][LS]
	const unsigned char *prof_lane = (const unsigned char *)(prof + lane * LS);

	const int m = a.m, u = a.u, o = a.o, e = a.e;
	const T m8 = V::delta(m), o8 = V::delta(o), e8 = V::delta(e), e8h = V::delta_h(e);
	const T HM = PROF ? (T)0 : m8;                                            // the carried H holds H + m in the fallback variant
	T nz = lane ? 1 : 0;                                                      // lane 0 takes matrix row 0 instead of a neighbour
	asm volatile("" : "+r"(nz));                                              // opaque: keep x * nz + b0 a multiply-add (FMA pipe), not a SEL
	const uint32_t mu8 = (uint32_t)(8 * (m >= u ? m - u : u - m));     // per half; < 1 << CSHIFT (host-checked)
	const int nsg = m >= u ? -1 : 1;                                    // M = (H+m) - penalty  (or + when u > m)
	const T ZERO = V::value(0);
	const T e8h4 = V::delta_h(e) + V::raw(4), o8p1 = o8 + V::raw(1);              // TAG: extension tag of L, M' (low bits 1) -> Mo' (low bits 2)
	const T NEGV = PACKED ? ZERO : (T)AT_NEG;                           // -inf stand-in (int32 lanes only)
	const bool want_ptr = a.want_ptr != 0;

	for (;;) {
		uint32_t job = 0;
		if (lane == 0) job = atomicAdd(a.counter, 1u);
		job = __shfl_sync(0xffffffffu, job, 0);
		if (job >= a.n_jobs) break;
		const FillJob jb = a.jobs[job];
		const uint32_t pA = jb.a, pB = PACKED ? jb.b : jb.a;
		const uint32_t l1A = a.q_len[pA], l1B = a.q_len[pB], l2 = a.t_len[pA];
		const uint8_t *__restrict__ qA = a.q + a.q_off[pA], *__restrict__ qB = a.q + a.q_off[pB];
		const uint8_t *__restrict__ tA = a.t + a.t_off[pA], *__restrict__ tB = a.t + a.t_off[pB];
		const uint8_t *__restrict__ jm = JUMP ? a.jmask + a.t_off[pA] : nullptr;
		uint32_t *__restrict__ ptr = a.ptr + a.ptr_off[pA - a.pair_base];
		const uint32_t t_last = (l2 + 31u) | (JUMP ? 31u : (SPW - 1u));
		const uint32_t G = t_last / SPW + 1, GJ = (t_last >> 5) + 1;
		constexpr uint32_t n_stripes = 1;
		uint32_t *__restrict__ ptrJ = ptr + (size_t)n_stripes * G * RPP;

		// ring[(j-1) & 511]: packed (tA << 8) | (tB << 24); int32 (tA << 16) | blacklist bit;
		// PROF: byte offset of the column's comb in the profile | blacklist bit
		auto load_block = [&](uint32_t blk) {
			const uint32_t base = blk * 256u;
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				const uint32_t idx = base + k * 32u + lane;
				if (PROF) {
					uint32_t v = 0;                            // past the end: any comb (those cells are never read back)
					if (idx < l2) {
						v = __ldg(a.symmap + __ldg(tA + idx));
						if (PACKED) v = v * 4u + __ldg(a.symmap + __ldg(tB + idx));
						v *= COMB_BYTES;
						if (JUMP) v |= __ldg(jm + idx) ? 1u : 0u;
					}
					ring16[idx & (AT_RING - 1)] = (uint16_t)v;
				} else {
					uint32_t v = PACKED ? 0x00010001u : 0x2u;      // past the end: never equals a symbol
					if (idx < l2) {
						if (PACKED) v = ((uint32_t)__ldg(tA + idx) << 8) | ((uint32_t)__ldg(tB + idx) << 24);
						else { v = (uint32_t)__ldg(tA + idx) << 16; if (JUMP) v |= __ldg(jm + idx) ? 1u : 0u; }
					}
					ring[idx & (AT_RING - 1)] = v;
				}
			}
		};

		// results
		int lbest_sc[2] = {-1, -1}, lbest_i[2] = {0x7fffffff, 0x7fffffff}, lbest_j[2] = {0, 0};   // local
		int capM = AT_NEG_INIT, capMj = 0, capL = AT_NEG_INIT, capLj = 0;                     // fit
		int gH = 0, gC = 0;                                                                    // global

		for (uint32_t stripe = 0; stripe < n_stripes; ++stripe) {
			const uint32_t row0 = stripe * RPP + lane * R;
			const bool last_stripe = stripe + 1 == n_stripes;
			__syncwarp();
			load_block(0);
			load_block(1);
			__syncwarp();

			T Mol[R], Ul[R], Hl[R], Cl[R], Jl[R], crow[R];
			uint32_t ac[R], acc[R], accJ[R];
#pragma unroll
			for (int r = 0; r < R; ++r) {
				const uint32_t ri = row0 + r;
				const int i = (int)ri + 1;
				if (PACKED) {
					const uint32_t ca = ri < l1A ? ((uint32_t)qA[ri] << 8) : 0x0002u;
					const uint32_t cb = ri < l1B ? ((uint32_t)qB[ri] << 8) : 0x0002u;
					ac[r] = ca | (cb << 16);
					// running-max key offset: 7 - r for real rows; -0x8000 sinks padded rows below every real key
					// (TAG: the key is built from M' = M | 1, so the offsets are one less -- per half, two's complement)
					const uint32_t ka = (ri < l1A ? (uint32_t)(7 - r) : 0x8000u) - (TAG ? 1u : 0u), kb = (ri < l1B ? (uint32_t)(7 - r) : 0x8000u) - (TAG ? 1u : 0u);
					crow[r] = (T)((ka & 0xffffu) | ((kb & 0xffffu) << 16));
				} else {
					ac[r] = ri < l1A ? ((uint32_t)qA[ri] << 16) : 0x4u;
					crow[r] = (T)(ri < l1A ? 7 - r : -(1 << 28));
				}
				// column 0 (left border)
				if (MODE == MODE_GLOBAL)     { Mol[r] = NEGV; Ul[r] = NEGV; Hl[r] = V::value(o + e * i) + HM; Cl[r] = V::raw(ST_LOW); }   // :432-436
				else if (MODE == MODE_LOCAL) { Mol[r] = ZERO + o8; Ul[r] = ZERO; Hl[r] = ZERO + HM; Cl[r] = V::raw(ST_LOW); }             // calloc zeros
				else                         { Mol[r] = NEGV; Ul[r] = NEGV; Hl[r] = NEGV; Cl[r] = V::raw(ST_MID); }                       // :612-617
				if (TAG) { Mol[r] = ZERO + o8 + V::raw(2); Hl[r] = ZERO + V::raw(2); }                                                    // Mo', H' = 0 | LOW
				Jl[r] = NEGV; acc[r] = 0; accJ[r] = 0;
				if (PROF) {     // this lane's rows of the query profile: 8*s(read symbol, target symbol of code c)
					const uint32_t qa = ri < l1A ? (uint32_t)qA[ri] : 0x100u, qb = (PACKED && ri < l1B) ? (uint32_t)qB[ri] : 0x100u;
					int sa[4], sb[4];
#pragma unroll
					for (int c = 0; c < 4; ++c) {
						const uint32_t sy = (a.syms >> (8 * c)) & 255u;
						sa[c] = 8 * (qa == sy ? m : u); sb[c] = 8 * (qb == sy ? m : u);
					}
#pragma unroll
					for (int c = 0; c < NC; ++c)
						prof[(c * 32 + lane) * LS + r] = TAG ? (((uint32_t)(sa[c >> 2] + 1) & 0xffffu) | ((uint32_t)(sb[c & 3] + 1) << 16))   // 8 s + 1 per half, two's complement (VIADDMNMX.U16x2 adds per half)
						                                 : PACKED ? (uint32_t)(sa[c >> 2] + sb[c & 3] * 65536) : (uint32_t)sa[c];
				}
			}
			if (PROF) __syncwarp();
			T sM = Mol[R - 1], sH = Hl[R - 1], sC = Cl[R - 1];
			T sL = MODE == MODE_GLOBAL ? V::value(o + e * (int)(row0 + R)) : (MODE == MODE_LOCAL ? ZERO : NEGV);
			if (TAG) sL = ZERO + V::raw(2);
			T pH, pC;      // H(row0, 0) + m and its code
			if (row0 == 0) {
				if (MODE == MODE_GLOBAL)     { pH = V::value(o < 0 ? 0 : o) + HM; pC = V::raw(o < 0 ? ST_MID : ST_LOW); }   // max5(L=o, M=0, U=o)
				else if (MODE == MODE_LOCAL) { pH = ZERO + HM; pC = V::raw(ST_LOW); }
				else                         { pH = ZERO + HM; pC = V::raw(ST_MID); }                                       // M[0][0]=U[0][0]=0
				if (TAG) pH = ZERO + V::raw(2);
			} else {
				if (MODE == MODE_GLOBAL)     { pH = V::value(o + e * (int)row0) + HM; pC = V::raw(ST_LOW); }
				else if (MODE == MODE_LOCAL) { pH = ZERO + HM; pC = V::raw(ST_LOW); }
				else                         { pH = NEGV; pC = V::raw(ST_MID); }
				if (TAG) pH = ZERO + V::raw(2);
			}
			// matrix row 0 as lane 0 sees it (zero in every other lane: x = neighbour * nz + b0)
			T b0M, b0L, b0H, b0C, b0E = 0;
			if (MODE == MODE_GLOBAL)     { b0M = NEGV; b0L = NEGV; b0H = V::value(o) + HM; b0C = V::raw(ST_UPP); b0E = e8; }         // :437-441, U[0][j] = o + e j
			else if (MODE == MODE_LOCAL) { b0M = ZERO + o8; b0L = ZERO; b0H = ZERO + HM; b0C = V::raw(ST_LOW); }
			else                         { b0M = ZERO + o8; b0L = NEGV; b0H = ZERO + HM; b0C = V::raw(ST_MID); }                     // :619-624
			if (TAG) { b0M = ZERO + o8 + V::raw(2); b0L = ZERO + V::raw(2); b0H = ZERO + V::raw(2); }     // Mo', L', H' of matrix row 0
			if (lane) { b0M = 0; b0L = 0; b0H = 0; b0C = 0; b0E = 0; }
			const int cap_r = (!PACKED && last_stripe && lane == (int)(((l1A - 1) % RPP) / R)) ? (int)((l1A - 1) % R) : -1;
			int hot[R];
#pragma unroll
			for (int r = 0; r < R; ++r) hot[r] = r == cap_r ? 1 : 0;
			T kbest = PACKED ? (T)0 : (T)AT_NEG_INIT;     // below every real key
			T tbest = 0;

			// first step at which the running key took its (so far) final value -> column of the running maximum
			auto note_best = [&](const T before, const T after, const uint32_t t) {
				if (PACKED) { const T chg = V::flag(after ^ before, 1) * 0xffffu; tbest = (tbest & ~chg) | ((T)(t * 0x10001u) & chg); }
				else if (after != before) tbest = (T)t;
			};
			// `first`: first step of a pointer word -- the accumulator restarts instead of shifting
			auto step = [&](const uint32_t t, const bool checked, const bool first) {
				const int j = (int)t - lane;
				// neighbour's last row, or matrix row 0 at column j = t in lane 0 (multiply-add: keeps the ALU pipe free)
				T rM = __shfl_up_sync(0xffffffffu, sM, 1) * nz + b0M;
				T rL = __shfl_up_sync(0xffffffffu, sL, 1) * nz + b0L;
				T rH = __shfl_up_sync(0xffffffffu, sH, 1) * nz + b0H;
				T rC = 0;
				if (!TAG) rC = __shfl_up_sync(0xffffffffu, sC, 1) * nz + b0C;
				if (MODE == MODE_GLOBAL) rH += b0E * (T)t;
				if (checked && t == 0) { rH = pH; rC = pC; }   // step 0 only primes the pipeline: keep H(row0, 0)
				T D = pH, DC = pC;
				pH = rH; pC = rC;
				if (!checked || (j >= 1 && j <= (int)l2)) {
					uint32_t c = PROF ? (uint32_t)ring16[(uint32_t)(j - 1) & (AT_RING - 1)] : ring[(uint32_t)(j - 1) & (AT_RING - 1)];
					T jadd = 0;
					if (JUMP) { jadd = (c & 1u) ? (T)AT_NEG : V::delta(a.jp - o); c &= ~1u; }   // M[i][j-1] + jump, or barred (:659-665)
					uint32_t pw[LS + 1];                            // PROF: the profile words of this lane's rows for the column's comb
					if (PROF) {
						if (LS & 1) {
							const uint32_t *pp = (const uint32_t *)(prof_lane + c);
#pragma unroll
							for (int r2 = 0; r2 < R; ++r2) pw[r2] = pp[r2];
						} else {
							const uint2 *pp = (const uint2 *)(prof_lane + c);
#pragma unroll
							for (int r2 = 0; r2 < (R + 1) / 2; ++r2) { const uint2 v2 = pp[r2]; pw[2 * r2] = v2.x; pw[2 * r2 + 1] = v2.y; }
						}
					}
					T Lup = rL, MoUp = rM, Mo = 0, Ln = 0, Hm = 0, code = 0;
					int rowM = 0, rowL = 0;                        // fit: M, L of the pair's last row
					const T kold = kbest;
#pragma unroll
					for (int r = 0; r < R; ++r) {
						if (TAG) {
							const T Mt = V::addmax(D, (T)pw[r], ZERO);          // HOME or diagonal code + 1 in bits 0-1
							const T Lt = V::addmax(Lup, e8h4, MoUp);            // bit 2: extended
							const T Ut = V::addmax(Ul[r], e8h, Mol[r]);         // bit 1: opened
							const T Lk = Lt & ~V::raw(4);                       // clean | 2
							const T Uk = Ut & ~V::raw(2);                       // clean
							const T Mk = (Mt & ~V::raw(3)) | V::raw(1);         // clean | 1
							Mo = Mk + o8p1;                                     // M + o, low bits 2
							Hm = V::vmax3(Lk, Mk, Uk);                          // clean | code
							// un-negated nibble into the accumulator (the word is inverted at the store); the shifted-in
							// bits of the other half's oldest nibble are overwritten, so the word never needs a reset
							const T s1 = (Mt & V::raw(3)) | (Lt & ~V::raw(3));
							const T z = (s1 & V::raw(7)) | ((Ut * 4u) & ~V::raw(7));
							acc[r] = ((acc[r] * 16u) & ~V::raw(15)) | (z & V::raw(15));
							kbest = V::addmax(Mk, crow[r], kbest);
							D = Hl[r];
							Hl[r] = Hm; Ul[r] = Uk; Mol[r] = Mo;
							Lup = Lk; MoUp = Mo; Ln = Lk;
							continue;
						}
						T Mraw;                                                 // H(i-1,j-1) + s
						if (PROF) Mraw = D + (T)pw[r];
						else { const T tt = V::flag((T)(ac[r] ^ c), mu8); Mraw = tt * (T)nsg + D; }   // tt: 0 on a match, 8|m-u| otherwise
						T Mn = Mraw, pm = DC;
						if (MODE == MODE_LOCAL) { Mn = V::vmax(Mraw, ZERO); pm = DC | V::flag(Mn - Mraw, 3); }   // HOME: 0.0 strictly greater (:825)
						const T Lext = Lup + e8;
						Ln = V::vmax(Lext, MoUp);
						const T fL = V::flag(Ln - Lext, 4);                     // gap opened only when strictly better (:456)
						const T Un = V::addmax(Ul[r], e8h, Mol[r]);
						const T fU = V::flag(Un - Mol[r], 8);                   // gap extended only when strictly better (:460)
						T Jn = 0, fJ = 0;
						if (JUMP) {
							const T ent = Mol[r] + jadd;
							Jn = V::vmax(ent, Jl[r]);
							fJ = V::flag(Jn - ent, 1);                          // stay in J only when strictly better (:660)
						}
						Mo = Mn + o8;
						T H = V::vmax3(Ln, Mn, Un);
						// 0 LOW, 1 MID, 2 UPP: first strictly greater in the order L, M, U.  notL / notM are 0 or 3:
						// L wins -> 0;  M wins (H > L, H == M) -> 3 & 1;  U wins (H > L, H > M) -> 3 & 2
						code = V::flag(H - Ln, 3) & (V::flag(H - Mn, 3) ^ V::raw(1));
						if (JUMP) { const T H4 = V::vmax(H, Jn); code = V::vmax(code, V::flag(H4 - H, 3)); H = H4; }
						Hm = H + HM;
						acc[r] = (first ? 0u : acc[r] * 16u) + (uint32_t)(pm | fL) + (uint32_t)fU;
						if (JUMP) accJ[r] = accJ[r] * 2u + (uint32_t)fJ;
						if (MODE == MODE_LOCAL) kbest = V::addmax(Mn, crow[r], kbest);
						// fit: the pair's last row is row cap_r of ONE lane: pick its values with a one-hot multiply-add per
						// row (FMA pipe) instead of compares and selects in every row; the search itself runs once per step
						if (MODE == MODE_FIT) { rowM += (int)Mn * hot[r]; rowL += (int)Ln * hot[r]; }
						if (MODE == MODE_GLOBAL) { if (r == cap_r && j == (int)l2) { gH = (int)H; gC = (int)code; } }   // (cheaper than the one-hot form here)
						D = Hl[r]; DC = Cl[r];
						Hl[r] = Hm; Cl[r] = code; Ul[r] = Un; Mol[r] = Mo; if (JUMP) Jl[r] = Jn;
						Lup = Ln; MoUp = Mo;
					}
					sM = Mo; sL = Ln; sH = Hm; sC = code;
					if (MODE == MODE_FIT && cap_r >= 0 && j < (int)l2) {       // column l2 excluded (:677, :684)
						if (rowM > capM) { capM = rowM; capMj = j; }
						if (rowL > capL) { capL = rowL; capLj = j; }
					}
					if (MODE == MODE_LOCAL) note_best(kold, kbest, t);
				} else {
#pragma unroll
					for (int r = 0; r < R; ++r) { acc[r] = (first && !TAG) ? 0u : acc[r] * 16u; if (JUMP) accJ[r] *= 2u; }
				}
			};

			for (uint32_t tb = 0; tb <= t_last; tb += SPW) {
				if ((tb & 255u) == 32u && tb > 32u) {     // block tb/256 - 1 is dead: refill its slots two blocks ahead
					__syncwarp();
					load_block(tb / 256u + 1u);
					__syncwarp();
				}
				if (tb >= 32u && tb + SPW - 1u <= l2) {
					if (PACKED) {
#pragma unroll
						for (uint32_t k = 0; k < SPW; ++k) step(tb + k, false, k == 0);
					} else {      // int32 lanes: 8 steps per word; unroll by 4 only (instruction-cache footprint)
#pragma unroll 4
						for (uint32_t k = 0; k < SPW; ++k) step(tb + k, false, false);
					}
				} else {
#pragma unroll 1
					for (uint32_t k = 0; k < SPW; ++k) step(tb + k, true, PACKED && k == 0);
				}
				if (want_ptr) {
					uint32_t *w = ptr + ((size_t)(stripe * G + tb / SPW) * 32 + lane) * R;
#pragma unroll
					for (int r = 0; r < R; ++r) w[r] = TAG ? ~acc[r] : acc[r];
				}
				if (JUMP && want_ptr && ((tb + SPW - 1u) & 31u) == 31u) {
					uint32_t *w = ptrJ + ((size_t)(stripe * GJ + (tb >> 5)) * 32 + lane) * R;
#pragma unroll
					for (int r = 0; r < R; ++r) w[r] = accJ[r];
				}
			}

			if (MODE == MODE_LOCAL) {   // fold this stripe's lane maximum into (score, row, column) per pair half
#pragma unroll
				for (int h = 0; h < (PACKED ? 2 : 1); ++h) {
					int key, tcol;
					if (PACKED) { key = (int)((kbest >> (16 * h)) & 0xffffu) - 0x8000; tcol = (int)((tbest >> (16 * h)) & 0xffffu) - lane; }
					else { key = (int)kbest; tcol = (int)tbest - lane; }
					const bool real = key >= 0;
					if (real) {
						const int sc = key >> 3, row = (int)row0 + (7 - (key & 7)) + 1;
						if (sc > lbest_sc[h]) { lbest_sc[h] = sc; lbest_i[h] = row; lbest_j[h] = tcol; }   // later stripes hold larger rows: strict >
					}
				}
			}
			__syncwarp();
		}

		// ---- end cell (reference: :466-469 global, :673-690 fit, running max :830-833 local) ----
		if (MODE == MODE_LOCAL) {
#pragma unroll
			for (int h = 0; h < (PACKED ? 2 : 1); ++h) {
				int sc = lbest_sc[h], row = lbest_i[h], col = lbest_j[h];
#pragma unroll
				for (int d = 16; d >= 1; d >>= 1) {
					const int osc = __shfl_xor_sync(0xffffffffu, sc, d);
					const int orow = __shfl_xor_sync(0xffffffffu, row, d);
					const int ocol = __shfl_xor_sync(0xffffffffu, col, d);
					if (osc > sc || (osc == sc && orow < row)) { sc = osc; row = orow; col = ocol; }
				}
				const uint32_t p = h ? pB : pA;
				if (lane == 0 && (h == 0 || pB != pA)) { a.score[p] = sc; a.end_i[p] = row; a.end_j[p] = col; a.end_state[p] = ST_MID; }
			}
		} else {
			const int owner = (int)(((l1A - 1) % RPP) / R);
			if (lane == owner) {
				if (MODE == MODE_GLOBAL) { a.score[pA] = gH >> 3; a.end_i[pA] = l1A; a.end_j[pA] = l2; a.end_state[pA] = (uint8_t)gC; }
				else {
					const bool useL = capL > capM;            // L replaces M only when strictly greater (:685)
					a.score[pA] = (useL ? capL : capM) >> 3; a.end_i[pA] = l1A; a.end_j[pA] = useL ? capLj : capMj;
					a.end_state[pA] = useL ? ST_LOW : ST_MID;
				}
			}
		}
	}
}

}  // namespace atb2
