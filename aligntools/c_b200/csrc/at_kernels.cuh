// at_kernels.cuh -- sm_100a device code of the alignTools DP core.
//
// K1  at_fill_affine<MODE,R,JUMP>  : Gotoh M/L/U(/J) fill, one pair per warp, int32 lanes.
//     at_fill_linear<MODE,R>       : single-plane fill (overlap max-plus / edit min-plus).
// K3  at_traceback_walk            : device traceback, one thread per pair chases the pointers
//                                    once and leaves the reversed run-length ops in scratch;
//     at_traceback_emit            : one warp per pair writes the dense CIGAR and replays it over
//                                    the sequences into the gapped strings r1 / r2.
//
// Geometry.  Rows i <-> read s1, columns j <-> target s2.  A warp
// sweeps a STRIPE of 32*R rows over all columns as a systolic array: lane k owns rows
// row0+k*R .. +R-1 and at step t works on column j = t - k, so the 32 lanes sit on one
// anti-diagonal of R-row blocks.  The last row of each lane (M+o, L, H=max(L,M,U[,J]) and
// the 2-bit argmax code of H) moves to lane k+1 with __shfl_up_sync once per step; lane 31
// parks it in a boundary row that lane 0 reads back on the next stripe (reads longer than
// 32*R rows).  Per-row state (M+o, U, H, code, J of column j-1) stays in registers.
//
// Traceback pointers: one nibble per cell, bits 0-1 = pointerM (0 LOW, 1 MID, 2 UPP,
// 3 HOME|JUMP), bit 2 = pointerL is MID (gap opened), bit 3 = pointerU is UPP (gap
// extended); fit+jump adds a 1-bit plane (pointerJ is JUMP).  Reference planes:
// src/alignment.h:44-47, values :27-34.  Nibbles of 8 consecutive steps of one row are
// packed into a 32-bit word (earliest step in the top nibble) and stored as
//   word[((stripe*G + step/8)*32 + lane)*R + r]
// i.e. in SKEWED coordinates (step = column + lane) with the R words of a lane's 8-step block
// contiguous: a warp flush writes one dense 128*R-byte run, and a diagonal traceback move
// (row-1, column-1) stays inside the same 32-byte sector most of the time.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace at {

enum { MODE_GLOBAL = 0, MODE_LOCAL = 1, MODE_FIT = 2, MODE_OVERLAP = 3, MODE_EDIT = 4 };
enum { ST_LOW = 0, ST_MID = 1, ST_UPP = 2, ST_JUMP = 3 };   // also the 2-bit pointerM codes
enum { CIG_M = 0, CIG_I = 1, CIG_D = 2, CIG_N = 3 };

// -INFINITY stand-in: a NEG-like value never beats a finite one as long
// as (l1+l2+2)*max|param| < 2^27, which the host checks (AT_E_RANGE).
#define AT_NEG (-(1 << 29))
#define AT_NEG_INIT (-(1 << 30) - (1 << 29))

struct FillArgs {
	const uint8_t  *q;       const uint64_t *q_off;  const uint32_t *q_len;
	const uint8_t  *t;       const uint64_t *t_off;  const uint32_t *t_len;
	const uint8_t  *jmask;   // fit+jump: 1 where entering J is forbidden; indexed like t
	const uint32_t *jobs;    // FillJob records {a, b} (largest pairs first); the single-plane kernels read .a
	uint32_t        n_jobs;
	uint32_t       *counter; // dynamic job queue
	uint32_t       *ptr;     // traceback pointer arena (32-bit words)
	const uint64_t *ptr_off; // [pair] word offset of the pair's pointer block (chunk-local index)
	uint32_t        pair_base; // first pair of the chunk (ptr_off is indexed pair - pair_base)
	int4           *bnd;     // stripe boundary rows, one slab per resident warp
	uint32_t        bnd_stride;
	int32_t        *score;   uint32_t *end_i;  uint32_t *end_j;  uint8_t *end_state;
	int             m, u, o, e, jp;
	int             want_ptr;
};

__device__ __forceinline__ uint32_t steps_last(uint32_t l2, int align_mask) { return (l2 + 31u) | (uint32_t)align_mask; }

}  // namespace at
#include "at_fill_affine.cuh"
namespace at {

// ------------------------------------------------------------------------------------
// Single-plane fill: overlap (src/alignment.h:926-964, linear gap `o`, order
// LEFT > DIAGONAL > RIGHT, 2-bit pointers, 16 steps per word) and edit distance
// (:291-315, min-plus, unit gaps, no pointers).
// ------------------------------------------------------------------------------------
template <int MODE, int R>
__device__ __forceinline__ void fill_linear_pair(const FillArgs &a, const uint32_t p, const int lane, const uint32_t warp_slot)
{
	const uint32_t l1 = a.q_len[p], l2 = a.t_len[p];
	const uint8_t *__restrict__ q  = a.q + a.q_off[p];
	const uint8_t *__restrict__ tg = a.t + a.t_off[p];
	uint32_t *__restrict__ ptr = MODE == MODE_OVERLAP ? a.ptr + a.ptr_off[p - a.pair_base] : nullptr;
	const int m = a.m, u = a.u, o = a.o;
	constexpr int RPP = 32 * R;
	const uint32_t n_stripes = (l1 + RPP - 1) / RPP;
	const uint32_t t_last = steps_last(l2, 15);
	const uint32_t G = (t_last >> 4) + 1;
	int4 *__restrict__ bnd = a.bnd + (size_t)warp_slot * a.bnd_stride;
	const bool want_ptr = a.want_ptr != 0 && MODE == MODE_OVERLAP;
	int capM = 0, capMj = 0;     // overlap: M[l1][0] = 0 seeds the search (:954-959)
	int eH = 0;                  // edit: M[l1][l2]

	for (uint32_t stripe = 0; stripe < n_stripes; ++stripe) {
		const uint32_t row0 = stripe * RPP + lane * R;
		const bool last_stripe = stripe + 1 == n_stripes;
		int ac[R], Ml[R];
		uint32_t acc[R];
#pragma unroll
		for (int r = 0; r < R; ++r) {
			const uint32_t ri = row0 + r;
			ac[r] = ri < l1 ? (int)q[ri] : 0x100;
			Ml[r] = MODE == MODE_OVERLAP ? 0 : (int)ri + 1;     // M[i][0] = 0 (:938) | i (:301)
			acc[r] = 0;
		}
		int sM = Ml[R - 1];
		int pH = MODE == MODE_OVERLAP ? 0 : (int)row0;           // M[row0][0]
		const int cap_r = (last_stripe && lane == (int)(((l1 - 1) % RPP) / R)) ? (int)((l1 - 1) % R) : -1;
		int top_next = 0;
		if (stripe > 0 && lane == 0) top_next = bnd[1].x;

		for (uint32_t t = 1; t <= t_last; ++t) {
			const int j = (int)t - lane;
			int rM = __shfl_up_sync(0xffffffffu, sM, 1);
			if (lane == 0) {
				if (stripe == 0) rM = MODE == MODE_OVERLAP ? AT_NEG : j;      // M[0][j] = -inf (:937) | j (:302)
				else { rM = top_next; if (t + 1 <= l2) top_next = bnd[t + 1].x; }
			}
			int D = pH;
			pH = rM;
			if (j >= 1 && j <= (int)l2) {
				const int c = (int)__ldg(tg + (j - 1));
				int Mup = rM, v = 0;
#pragma unroll
				for (int r = 0; r < R; ++r) {
					const bool eq = ac[r] == c;
					if (MODE == MODE_OVERLAP) {
						const int left = Ml[r] + o, diag = D + (eq ? m : u), up = Mup + o;
						v = left; uint32_t code = 1;                       // 1 LEFT, 2 DIAGONAL, 3 RIGHT (0 = never set)
						if (diag > v) { v = diag; code = 2; }
						if (up > v)   { v = up;   code = 3; }
						acc[r] = (acc[r] << 2) | code;
						if (r == cap_r && j < (int)l2) { if (v > capM) { capM = v; capMj = j; } }
					} else {
						const int left = Ml[r] + 1, diag = D + (eq ? 0 : u), up = Mup + 1;
						v = min(min(left, diag), up);                     // min3 (:280-286)
						if (r == cap_r && j == (int)l2) eH = v;
					}
					D = Ml[r]; Ml[r] = v; Mup = v;
				}
				sM = v;
				if (lane == 31 && !last_stripe) bnd[j].x = sM;
			} else if (MODE == MODE_OVERLAP) {
#pragma unroll
				for (int r = 0; r < R; ++r) acc[r] <<= 2;
			}
			if (want_ptr && (t & 15u) == 15u) {
				uint32_t *w = ptr + ((size_t)(stripe * G + (t >> 4)) * 32 + lane) * R;
#pragma unroll
				for (int r = 0; r < R; ++r) w[r] = acc[r];
			}
		}
		__syncwarp();
	}
	const int owner = (int)(((l1 - 1) % RPP) / R);
	if (lane == owner) {
		if (MODE == MODE_OVERLAP) { a.score[p] = capM; a.end_i[p] = l1; a.end_j[p] = capMj; a.end_state[p] = ST_MID; }
		else { a.score[p] = eH; a.end_i[p] = l1; a.end_j[p] = l2; a.end_state[p] = ST_MID; }
	}
}

template <int MODE, int R>
__global__ void __launch_bounds__(128) at_fill_linear(const FillArgs a)
{
	const int lane = threadIdx.x & 31;
	const uint32_t warp_slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	for (;;) {
		uint32_t job = 0;
		if (lane == 0) job = atomicAdd(a.counter, 1u);
		job = __shfl_sync(0xffffffffu, job, 0);
		if (job >= a.n_jobs) break;
		fill_linear_pair<MODE, R>(a, a.jobs[2 * job], lane, warp_slot);   // jobs are FillJob{a,b} records; int32 lanes use .a
	}
}

// ------------------------------------------------------------------------------------
// K3 traceback (reference: trace_back_gla :372-412, trace_back_fit_affine_jump :558-592,
// trace_back_local_affine :766-800, trace_back_overlap :896-922).  The pointer is read at
// the cell BEFORE the move; local emits the HOME cell's column; global flushes the
// remaining target then read symbols after the loop.
// ------------------------------------------------------------------------------------
struct TraceArgs {
	const uint8_t  *q;  const uint64_t *q_off;  const uint32_t *q_len;
	const uint8_t  *t;  const uint64_t *t_off;  const uint32_t *t_len;
	const uint32_t *ptr; const uint64_t *ptr_off;
	const uint8_t  *rclass;    // [pair] bits 0-3: rows-per-lane R the fill used; bits 4-5: 0 = int32 layout,
	                           //        1 / 2 = packed s16x2 layout, pair in the low / high half
	uint32_t        pair_base, n_pairs;   // chunk
	const uint32_t *end_i, *end_j; const uint8_t *end_state;
	uint32_t       *beg_i, *beg_j;
	uint32_t       *n_ops, *n_cols;       // [pair]
	uint32_t       *scratch;              // reversed run-length ops, slot of l1+l2 ops per pair
	const uint64_t *scratch_off;          // [n_chunk] chunk-local slot offsets
	const uint64_t *ops_off, *cols_off;   // [n_chunk+1] exclusive scans (chunk-local index)
	uint32_t       *cigar;                // dense ops of the chunk
	uint8_t        *aln1, *aln2;          // dense columns of the chunk
	int             mode, jump;
};

struct PtrView {
	const uint32_t *ptr, *ptrJ;
	uint32_t R, RPP, G, GJ, half;
	__device__ __forceinline__ uint32_t nib(uint32_t i, uint32_t j) const {
		const uint32_t ri = i - 1, stripe = ri / RPP, rem = ri - stripe * RPP, lane = rem / R, r = rem - lane * R, t = j + lane;
		if (half) {   // packed s16x2: 4 steps x 2 pairs per word, single stripe
			const uint32_t w = __ldg(ptr + ((size_t)(t >> 2) * 32 + lane) * R + r);
			return (w >> (16 * (half - 1) + 4 * (3 - (t & 3)))) & 15u;
		}
		const uint32_t w = __ldg(ptr + ((size_t)(stripe * G + (t >> 3)) * 32 + lane) * R + r);
		return (w >> (4 * (7 - (t & 7)))) & 15u;
	}
	__device__ __forceinline__ uint32_t jbit(uint32_t i, uint32_t j) const {
		const uint32_t ri = i - 1, stripe = ri / RPP, rem = ri - stripe * RPP, lane = rem / R, r = rem - lane * R, t = j + lane;
		const uint32_t w = __ldg(ptrJ + ((size_t)(stripe * GJ + (t >> 5)) * 32 + lane) * R + r);
		return (w >> (31 - (t & 31))) & 1u;
	}
	__device__ __forceinline__ uint32_t two(uint32_t i, uint32_t j) const {   // overlap: 2 bits, 16 steps per word
		const uint32_t ri = i - 1, stripe = ri / RPP, rem = ri - stripe * RPP, lane = rem / R, r = rem - lane * R, t = j + lane;
		const uint32_t w = __ldg(ptr + ((size_t)(stripe * G + (t >> 4)) * 32 + lane) * R + r);
		return (w >> (2 * (15 - (t & 15)))) & 3u;
	}
};

// run-length encoder writing ops in walk (= reverse) order
struct RunWriter {
	uint32_t *dst; uint32_t n_ops, n_cols, run, op;
	__device__ __forceinline__ void flush() { if (run) dst[n_ops++] = (run << 4) | op; run = 0; }
	__device__ __forceinline__ void col(uint32_t o) { if (run && o != op) flush(); op = o; ++run; ++n_cols; }
};

// One thread per pair: chase the pointers once, leave the reversed CIGAR in the pair's scratch slot.
__global__ void __launch_bounds__(128) at_traceback_walk(const TraceArgs a)
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= a.n_pairs) return;
	const uint32_t p = a.pair_base + k;
	const uint32_t l1 = a.q_len[p], l2 = a.t_len[p];
	PtrView pv;
	pv.R = a.rclass[p] & 15u; pv.RPP = 32 * pv.R; pv.half = (a.rclass[p] >> 4) & 3u;
	const uint32_t n_stripes = (l1 + pv.RPP - 1) / pv.RPP;
	const bool jump = a.jump != 0;
	if (a.mode == MODE_OVERLAP) { const uint32_t tl = steps_last(l2, 15); pv.G = (tl >> 4) + 1; pv.GJ = 0; }
	else if (pv.half) { const uint32_t tl = steps_last(l2, 3); pv.G = (tl >> 2) + 1; pv.GJ = 0; }
	else { const uint32_t tl = steps_last(l2, jump ? 31 : 7); pv.G = (tl >> 3) + 1; pv.GJ = (tl >> 5) + 1; }
	pv.ptr = a.ptr + a.ptr_off[k];
	pv.ptrJ = pv.ptr + (size_t)n_stripes * pv.G * pv.RPP;
	RunWriter w;
	w.dst = a.scratch + a.scratch_off[k]; w.n_ops = w.n_cols = w.run = 0; w.op = 0;
	uint32_t i = a.end_i[p], j = a.end_j[p], state = a.end_state[p];

	if (a.mode == MODE_OVERLAP) {
		while (j > 0) {                                   // :899
			const uint32_t c = pv.two(i, j);
			if (c == 1)      { --j; w.col(CIG_D); }                  // LEFT
			else if (c == 2) { --i; --j; w.col(CIG_M); }             // DIAGONAL
			else if (c == 3) { --i; w.col(CIG_I); }                  // RIGHT
			else break;                                   // unset pointer: unreachable on a finite path
		}
	} else {
		bool home = false;
		for (;;) {
			const bool go = a.mode == MODE_FIT ? (i > 0) : (i > 0 && j > 0);   // :562 | :377, :771
			if (!go || home) break;
			if (j == 0 && state != ST_LOW) break;         // only reachable with corrupt pointers: never index s2[-1]
			const uint32_t nb = pv.nib(i, j);
			if (state == ST_LOW)      { state = (nb & 4u) ? ST_MID : ST_LOW; --i; w.col(CIG_I); }
			else if (state == ST_MID) {
				const uint32_t pm = nb & 3u;
				--i; --j; w.col(CIG_M);
				if (pm == 3 && a.mode == MODE_LOCAL) home = true;              // HOME: column emitted, then stop (:788-791)
				else state = pm;
			}
			else if (state == ST_UPP) { state = (nb & 8u) ? ST_UPP : ST_MID; --j; w.col(CIG_D); }
			else                      { state = pv.jbit(i, j) ? ST_JUMP : ST_MID; --j; w.col(CIG_N); }
		}
	}
	a.beg_i[p] = i; a.beg_j[p] = j;
	if (a.mode == MODE_GLOBAL) {                          // flush (:398-407): rest of the target, then rest of the read
		while (j > 0) { --j; w.col(CIG_D); }
		while (i > 0) { --i; w.col(CIG_I); }
	}
	w.flush();
	a.n_ops[p] = w.n_ops; a.n_cols[p] = w.n_cols;
}

// One warp per pair: reverse the scratch ops into the dense CIGAR and, on request, replay them
// forward over the two sequences to write the gapped strings r1 / r2 (coalesced).
__global__ void __launch_bounds__(128) at_traceback_emit(const TraceArgs a)
{
	const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (k >= a.n_pairs) return;
	const uint32_t p = a.pair_base + k;
	const uint32_t n = a.n_ops[p];
	const uint32_t *src = a.scratch + a.scratch_off[k];
	if (a.cigar) {
		uint32_t *dst = a.cigar + a.ops_off[k];
		for (uint32_t x = lane; x < n; x += 32) dst[x] = src[n - 1 - x];
	}
	if (a.aln1) {
		const uint8_t *q = a.q + a.q_off[p], *tg = a.t + a.t_off[p];
		uint8_t *a1 = a.aln1 + a.cols_off[k], *a2 = a.aln2 + a.cols_off[k];
		uint32_t i = a.mode == MODE_GLOBAL ? 0 : a.beg_i[p], j = a.mode == MODE_GLOBAL ? 0 : a.beg_j[p], col = 0;
		for (uint32_t x = 0; x < n; ++x) {
			const uint32_t op = src[n - 1 - x], len = op >> 4, code = op & 15u;
			const bool gap1 = code == CIG_D || code == CIG_N, gap2 = code == CIG_I;
			for (uint32_t c = lane; c < len; c += 32) {
				a1[col + c] = gap1 ? (uint8_t)'-' : q[i + c];
				a2[col + c] = gap2 ? (uint8_t)'-' : tg[j + c];
			}
			col += len; if (!gap1) i += len; if (!gap2) j += len;
		}
	}
}

// fit+jump: expand the per-pair blacklists into a byte mask aligned with the target bytes.
__global__ void at_build_jmask(const int32_t *sites, const uint64_t *site_off, const uint64_t *t_off,
                               const uint32_t *t_len, uint32_t n_pairs, uint8_t *jmask)
{
	const uint32_t p = blockIdx.x;
	if (p >= n_pairs) return;
	const uint64_t lo = site_off[p], hi = site_off[p + 1];
	for (uint64_t k = lo + threadIdx.x; k < hi; k += blockDim.x) {
		const int32_t s = sites[k];
		if (s >= 0 && (uint32_t)s < t_len[p]) jmask[t_off[p] + s] = 1;
	}
}

}  // namespace at
