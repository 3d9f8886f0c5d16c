// at_kernels_p16.cuh -- packed int16x2 ("s16x2") inter-pair fill for short reads.
//
// K1p  at_fill_local_p16<R> : local (Smith-Waterman, affine) fill, TWO pairs per warp: pair A
//      lives in the low 16 bits and pair B in the high 16 bits of every score register, so each
//      DPX instruction (VIMNMX.S16x2 / VIADD.16x2 / VIADDMNMX.S16x2, sm_90+) advances two cells.
//      Same systolic geometry as at_fill_affine (lane k owns R rows, step t works on column
//      t-k, __shfl_up_sync hands the strip's last row to lane k+1); single stripe (l1 <= 32R),
//      both pairs of a job share l2.  The target symbols of both pairs are staged through a
//      per-warp shared-memory ring (one 32-bit entry per column, symbol<<8 per half).
//
// Scores are kept multiplied by 8 (the host checks they fit int16).  That buys two things:
//   * a != b  ->  (a ^ b) >= 8, so  min(a ^ b, 4|8|1)  lands a pointer flag directly on its
//     bit of the nibble (one LOP3 + one VIMNMX.U16x2 per flag for both pairs);
//   * the three free low bits carry (7 - row_in_lane), so one VIADDMNMX keeps the lane's
//     running maximum keyed by (score, smaller row first) -- the reference's first-maximum
//     in row-major order (src/alignment.h:830-833) without per-row bookkeeping.
// Pointer nibble layout is the one of at_kernels.cuh; a word holds 4 steps x 2 pairs:
//   word[((step/4)*R + r)*32 + lane],  bits 16h + 4*(3 - step%4) .. +3  for pair half h.
#pragma once
#include "at_kernels.cuh"

namespace at {

struct FillArgsP16 {
	const uint8_t  *q;       const uint64_t *q_off;  const uint32_t *q_len;
	const uint8_t  *t;       const uint64_t *t_off;  const uint32_t *t_len;
	const uint2    *jobs;    // (pair A, pair B); B == A when a pair has no partner
	uint32_t        n_jobs;
	uint32_t       *counter;
	uint32_t       *ptr;     const uint64_t *ptr_off;   uint32_t pair_base;
	int32_t        *score;   uint32_t *end_i;  uint32_t *end_j;  uint8_t *end_state;
	int             m, u, o, e;
	int             want_ptr;
};

__device__ __forceinline__ uint32_t pk(int v) { return ((uint32_t)v & 0xffffu) * 0x10001u; }

#define AT_P16_RING 512
#define AT_P16_WARPS 4

template <int R>
__global__ void __launch_bounds__(32 * AT_P16_WARPS) at_fill_local_p16(const FillArgsP16 a)
{
	__shared__ uint32_t ring_all[AT_P16_WARPS][AT_P16_RING];
	const int lane = threadIdx.x & 31;
	uint32_t *ring = ring_all[threadIdx.x >> 5];
	const uint32_t m8 = pk(8 * a.m), o8 = pk(8 * a.o), e8 = pk(8 * a.e), mu8 = pk(8 * (a.m - a.u));
	const bool want_ptr = a.want_ptr != 0;

	for (;;) {
		uint32_t job = 0;
		if (lane == 0) job = atomicAdd(a.counter, 1u);
		job = __shfl_sync(0xffffffffu, job, 0);
		if (job >= a.n_jobs) break;
		const uint2 jb = a.jobs[job];
		const uint32_t pA = jb.x, pB = jb.y;
		const uint32_t l1A = a.q_len[pA], l1B = a.q_len[pB], l2 = a.t_len[pA];
		const uint8_t *__restrict__ qA = a.q + a.q_off[pA], *__restrict__ qB = a.q + a.q_off[pB];
		const uint8_t *__restrict__ tA = a.t + a.t_off[pA], *__restrict__ tB = a.t + a.t_off[pB];
		uint32_t *__restrict__ ptr = a.ptr + a.ptr_off[pA - a.pair_base];
		const uint32_t t_last = (l2 + 31u) | 3u;

		// ---- stage target columns: ring[(j-1) & 511] = (tA[j-1] << 8) | (tB[j-1] << 24) ----
		auto load_block = [&](uint32_t blk) {          // 256 columns per block
			const uint32_t base = blk * 256u;
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				const uint32_t idx = base + k * 32u + lane;
				uint32_t v = 0x00010001u;              // past the end: never equals a symbol
				if (idx < l2) v = ((uint32_t)__ldg(tA + idx) << 8) | ((uint32_t)__ldg(tB + idx) << 24);
				ring[idx & (AT_P16_RING - 1)] = v;
			}
		};
		__syncwarp();
		load_block(0);
		load_block(1);
		__syncwarp();

		// ---- per-row state: column 0 of local mode is all zeros (calloc) ----
		uint32_t ac[R], Mol[R], Ul[R], Hl[R], Cl[R], acc[R], crow[R];
#pragma unroll
		for (int r = 0; r < R; ++r) {
			const uint32_t ri = lane * R + r;
			const uint32_t ca = ri < l1A ? ((uint32_t)qA[ri] << 8) : 0x0002u;
			const uint32_t cb = ri < l1B ? ((uint32_t)qB[ri] << 8) : 0x0002u;
			ac[r] = ca | (cb << 16);
			Mol[r] = o8; Ul[r] = 0; Hl[r] = m8; Cl[r] = 0; acc[r] = 0;
			const uint32_t ka = ri < l1A ? (uint32_t)(7 - r) : (uint32_t)(-30000) & 0xffffu;
			const uint32_t kb_ = ri < l1B ? (uint32_t)(7 - r) : (uint32_t)(-30000) & 0xffffu;
			crow[r] = ka | (kb_ << 16);
		}
		uint32_t sM = o8, sL = 0, sH = m8, sC = 0;     // column 0 of the strip's last row: M+o, L, H+m, code
		uint32_t pH = m8, pC = 0;                      // H(row0, 0) + m, code LOW (all-zero tie -> first argument)
		uint32_t kbest = 0xffffffffu;                  // -1 per half: any real cell (key >= 0) beats it
		uint32_t tbest = 0;

		auto step = [&](const uint32_t t, const bool checked) {
			const int j = (int)t - lane;
			uint32_t rM = __shfl_up_sync(0xffffffffu, sM, 1);
			uint32_t rL = __shfl_up_sync(0xffffffffu, sL, 1);
			uint32_t rH = __shfl_up_sync(0xffffffffu, sH, 1);
			uint32_t rC = __shfl_up_sync(0xffffffffu, sC, 1);
			if (lane == 0) { rM = o8; rL = 0; rH = m8; rC = 0; }   // matrix row 0: zeros
			uint32_t D = pH, DC = pC;
			pH = rH; pC = rC;
			if (!checked || (j >= 1 && j <= (int)l2)) {
				const uint32_t c = ring[(uint32_t)(j - 1) & (AT_P16_RING - 1)];
				uint32_t Lup = rL, MoUp = rM, Mo = 0, Ln = 0, Hm = 0, code = 0;
				const uint32_t kold = kbest;
#pragma unroll
				for (int r = 0; r < R; ++r) {
					const uint32_t x = ac[r] ^ c;
					const uint32_t tt = __vminu2(x, mu8);                 // 0 on a match, 8(m-u) otherwise
					const uint32_t Mraw = __vsub2(D, tt);                 // H(i-1,j-1) + s
					const uint32_t Mn = __vmaxs2(Mraw, 0u);               // 0-floor
					const uint32_t h3 = __vminu2(Mn ^ Mraw, 0x00030003u); // HOME: 0.0 strictly greater (:825)
					const uint32_t Lext = __vadd2(Lup, e8);
					Ln = __vmaxs2(Lext, MoUp);
					const uint32_t fL = __vminu2(Ln ^ Lext, 0x00040004u); // gap opened only when strictly better (:456)
					const uint32_t Un = __viaddmax_s16x2(Ul[r], e8, Mol[r]);
					const uint32_t fU = __vminu2(Un ^ Mol[r], 0x00080008u); // gap extended only when strictly better (:460)
					Mo = __vadd2(Mn, o8);
					const uint32_t t1 = __vmaxs2(Ln, Mn);
					const uint32_t H = __vmaxs2(t1, Un);
					code = __vminu2(H ^ Ln, 0x00010001u) + __vminu2(H ^ t1, 0x00010001u);   // 0 LOW, 1 MID, 2 UPP
					Hm = __vadd2(H, m8);
					acc[r] = ((acc[r] << 4) + (DC | h3 | fL)) | fU;
					kbest = __viaddmax_s16x2(Mn, crow[r], kbest);
					D = Hl[r]; DC = Cl[r];
					Hl[r] = Hm; Cl[r] = code; Ul[r] = Un; Mol[r] = Mo;
					Lup = Ln; MoUp = Mo;
				}
				sM = Mo; sL = Ln; sH = Hm; sC = code;
				// first step at which each half's running key reached its (so far) final value
				const uint32_t chg = __vminu2(kbest ^ kold, 0x00010001u) * 0xffffu;
				tbest = (tbest & ~chg) | (pk((int)t) & chg);
			} else {
#pragma unroll
				for (int r = 0; r < R; ++r) acc[r] <<= 4;
			}
		};

		for (uint32_t tb = 0; tb <= t_last; tb += 4) {
			if ((tb & 255u) == 32u && tb > 32u) {     // block (tb/256 - 1) is dead: refill its slots two blocks ahead
				__syncwarp();
				load_block(tb / 256u + 1u);
				__syncwarp();
			}
			if (tb >= 32u && tb + 3u <= l2) {
#pragma unroll
				for (int k = 0; k < 4; ++k) step(tb + k, false);
			} else {
#pragma unroll
				for (int k = 0; k < 4; ++k) step(tb + k, true);
			}
			if (want_ptr) {
				uint32_t *w = ptr + ((size_t)(tb >> 2) * R) * 32 + lane;
#pragma unroll
				for (int r = 0; r < R; ++r) w[r * 32] = acc[r];
			}
#pragma unroll
			for (int r = 0; r < R; ++r) acc[r] = 0;   // 4 nibbles per half: the next shift must not spill A's bits into B's half
		}

		// ---- end cell per half: maximise (score, smaller row) over lanes; column = step - lane ----
#pragma unroll
		for (int h = 0; h < 2; ++h) {
			const int key = (int)(short)((kbest >> (16 * h)) & 0xffffu);
			const int tcol = (int)((tbest >> (16 * h)) & 0xffffu) - lane;
			int sc = key >> 3, row = lane * R + (7 - (key & 7)) + 1, col = tcol;
			if (key < 0) { sc = -1; row = 0x7fffffff; }
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
	}
}

}  // namespace at
