// at_runtime.cu -- host runtime behind include/aligntools_b200.h (sections A and C).
//
// A batch is sharded over the handle's devices as contiguous slices of pairs (no
// inter-GPU traffic).  Each shard keeps its sequences resident in HBM,
// cuts its pairs into CHUNKS whose traceback-pointer blocks fit the pointer arena, and
// per chunk runs  fill (one launch per rows-per-lane class R)  ->  traceback count walk
// ->  exclusive scans  ->  traceback emit walk, all on one stream, timed with CUDA events.
// There is no CPU fallback anywhere in this file: without a usable sm_100 device every
// entry returns AT_E_CUDA.
#include "../../../include/aligntools_b200.h"
#include "at_kernels.cuh"
#include "at_devmem.h"

#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace atb2;

// ------------------------------------------------------------------ handle ----
struct at_device {
	int id = 0;
	int sm_count = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t pipe[4] = {nullptr, nullptr, nullptr, nullptr};   // at_batch_align's pipeline streams (created on first use)
	std::shared_ptr<BigCache> big = std::make_shared<BigCache>();
};

struct at_batch;
struct at_handle {
	std::vector<at_device> devs;
	std::string err;
	std::mutex mu;
	std::atomic<uint64_t> launches{0};
	// at_batch_align's pipeline: one workspace per (device, worker), kept for the life of the handle so
	// that after the first call a pipelined batch makes no CUDA allocator call at all
	std::vector<at_batch *> pipe_ws;
	std::vector<uint64_t> pipe_prefix;     // running cell counts of the batch in hand (kept: no page faults per call)
	std::mutex align_mu;
};

static void set_err(at_handle *h, const char *fmt, ...)
{
	char buf[512];
	va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
	if (h) { std::lock_guard<std::mutex> g(h->mu); h->err = buf; }
}

#define CU(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	set_err(h, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return AT_E_CUDA; } } while (0)

extern "C" void at_default_params(at_params *p)
{
	if (!p) return;
	p->m = 1; p->u = -2; p->o = -5; p->e = -1; p->j = -10; p->jump = 0;   // init_opt, src/alignment.h:105-110
}

extern "C" const char *at_strerror(int rc)
{
	switch (rc) {
	case AT_OK: return "ok";
	case AT_E_ARG: return "parameter error";
	case AT_E_CUDA: return "CUDA error / no usable sm_100 device (there is no CPU fallback)";
	case AT_E_NOMEM: return "allocation failure";
	case AT_E_FITLEN: return "first sequence must be shorter than the second to do fitting alignment";
	case AT_E_NOSPACE: return "output buffer too small";
	case AT_E_RANGE: return "score range exceeds the int32 lanes";
	case AT_E_UNDEF: return "input undefined in the reference (empty record, or fit with l2 < 2)";
	default: return "unknown error";
	}
}

extern "C" const char *at_version(void) { return "aligntools-b200 0.1 (sm_100a)"; }

__global__ void at_unpack_2bit(const uint8_t *src, const uint64_t *src_off, const uint64_t *dst_off,
                               const uint32_t *len, uint32_t n_pairs, uint8_t *dst);

extern "C" int at_create(const int *devices, int n_devices, at_handle **out)
{
	if (!out) return AT_E_ARG;
	*out = nullptr;
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return AT_E_CUDA;
	at_handle *h = new at_handle();
	std::vector<int> ids;
	if (!devices || n_devices <= 0) ids.push_back(0);
	else ids.assign(devices, devices + n_devices);
	for (int id : ids) {
		if (id < 0 || id >= count) { delete h; return AT_E_ARG; }
		cudaDeviceProp prop;
		if (cudaGetDeviceProperties(&prop, id) != cudaSuccess || prop.major < 10) { delete h; return AT_E_CUDA; }
		at_device d; d.id = id; d.sm_count = prop.multiProcessorCount;
		// host threads waiting for the device yield their core between polls: the pipelined path runs three
		// workers per GPU, and a multi-rank job (one process per GPU) must not starve them on a small host
		if (cudaSetDevice(id) == cudaSuccess && cudaSetDeviceFlags(cudaDeviceScheduleYield) != cudaSuccess) cudaGetLastError();
		if (cudaSetDevice(id) != cudaSuccess || cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return AT_E_CUDA; }
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, id) == cudaSuccess) {
			uint64_t keep = UINT64_MAX;      // keep freed blocks cached in the pool (trimmed in at_destroy)
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
		// The helper kernels ask for the same (maximal) shared-memory carve-out as the fill kernels: an
		// SM runs kernels of different carve-outs only one after the other, so without this every small
		// kernel of the pipelined path would wait for a persistent fill grid of another stream to drain.
		const void *helpers[] = {(const void *)at_traceback_walk<true>, (const void *)at_traceback_emit<true>, (const void *)at_scan_offsets,
		                         (const void *)at_symbol_set, (const void *)at_build_jmask, (const void *)at_unpack_2bit, (const void *)at_plan_uniform};
		for (const void *fn : helpers) cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
		h->devs.push_back(d);
	}
	*out = h;
	return AT_OK;
}

extern "C" void at_batch_free(at_batch *b);

extern "C" void at_destroy(at_handle *h)
{
	if (!h) return;
	for (at_batch *ws : h->pipe_ws) if (ws) at_batch_free(ws);
	h->pipe_ws.clear();
	for (auto &d : h->devs) {
		cudaSetDevice(d.id);
		if (d.stream) { cudaStreamSynchronize(d.stream); cudaStreamDestroy(d.stream); }
		for (auto &ps : d.pipe) if (ps) { cudaStreamSynchronize(ps); cudaStreamDestroy(ps); }
		d.big->drop_all();
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, d.id) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
	}
	delete h;
}

extern "C" const char *at_last_error(const at_handle *h)
{
	if (!h) return "";
	static thread_local std::string copy;      // another thread of the handle may be rewriting h->err
	{ std::lock_guard<std::mutex> g(const_cast<at_handle *>(h)->mu); copy = h->err; }
	return copy.c_str();
}
extern "C" int at_device_count(const at_handle *h) { return h ? (int)h->devs.size() : 0; }
extern "C" uint64_t at_launch_count(const at_handle *h) { return h ? h->launches.load() : 0; }

// ------------------------------------------------------------------- batch ----
static const int MAXR = 8;
static const size_t AT_SEQ_SLACK = 1024;   // K2 stages 256-byte target tiles by TMA: the last tile may run past the last record

// One kernel launch of a chunk: the jobs of one (kernel kind, rows-per-lane) class.
enum { LK_INT32 = 0, LK_PACKED = 1, LK_WAVE = 2, LK_BITS = 3 };   // LK_BITS: bit-parallel edit distance (K2 task geometry, 1024*r rows per stripe)
struct Launch {
	int kind = LK_INT32, r = 1;
	uint64_t cells = 0;
	std::vector<FillJob> h_jobs;        // LK_INT32: one pair per warp (a == b); LK_PACKED: two pairs per warp
	DevBuf<FillJob> d_jobs;
	std::vector<WaveTask> h_tasks;      // LK_WAVE / LK_BITS: (pair, stripe) in queue order (order_wave_tasks)
	DevBuf<WaveTask> d_tasks;
	uint64_t prog_base = 0;             // first progress word of this launch in Shard::d_prog
	size_t n_planned = 0;               // uniform shards: the job list exists on the device only (at_plan_uniform)
	size_t n_jobs() const { return n_planned ? n_planned : (kind >= LK_WAVE ? h_tasks.size() : h_jobs.size()); }
};

struct Chunk {
	uint32_t k0 = 0, k1 = 0;            // shard-local pair range
	uint64_t ptr_words = 0;
	std::vector<uint64_t> h_ptr_off;    // chunk-local
	DevBuf<uint64_t> d_ptr_off;
	std::vector<Launch> launches;
	std::vector<uint64_t> h_bnd_off;    // chunk-local: element offset of the pair's two boundary slabs (K2)
	DevBuf<uint64_t> d_bnd_off;
	uint64_t bnd_elems = 0, prog_words = 0;
	DevBuf<uint64_t> d_ops_off, d_cols_off, d_scratch_off;
	std::vector<uint64_t> h_scratch_off; uint64_t scratch_words = 0;
	DevBuf<uint32_t> d_cigar; DevBuf<uint8_t> d_aln1, d_aln2;   // exact-size outputs of this chunk ...
	uint32_t *cigar = nullptr; uint8_t *aln1 = nullptr, *aln2 = nullptr;   // ... or the workspace's grow-only buffers (pipelined path); what fetch reads
	uint64_t tot_ops = 0, tot_cols = 0;
};

// Sub-slices of one device take turns (in pair order) on the host->device copy of their sequences.
struct UploadGate {
	std::mutex mu; std::condition_variable cv; uint64_t turn = 0;
	cudaEvent_t last = nullptr;      // recorded behind the previous turn's copies: the next turn's stream waits for it, not the host
	void wait_for(uint64_t k) { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return turn >= k; }); }
	void pass(uint64_t k) { std::lock_guard<std::mutex> lk(mu); if (turn < k + 1) { turn = k + 1; cv.notify_all(); } }    // idempotent
};

struct Shard {
	at_device *dev = nullptr;
	cudaStream_t stream = nullptr;      // the device's main stream, or a pipeline stream (at_batch_align)
	uint64_t p0 = 0, p1 = 0;            // pair range within the batch's input arrays
	uint64_t out_base = 0;              // index of pair p0 in the caller's output arrays
	uint32_t n = 0;
	DevBuf<uint8_t> d_q, d_t, d_jmask, d_rclass, d_end_state, d_q2, d_t2;
	DevBuf<uint64_t> d_q_off, d_t_off, d_site_off;
	DevBuf<uint32_t> d_q_len, d_t_len, d_end_i, d_end_j, d_beg_i, d_beg_j, d_n_ops, d_n_cols, d_counter;
	DevBuf<int32_t> d_score, d_sites;
	DevBuf<uint32_t> ws_cigar; DevBuf<uint8_t> ws_aln1, ws_aln2;     // pipeline workspace: outputs of the sub-slice, sized by their upper bound
	DevBuf<uint32_t> d_ptr, d_scratch, d_prog; DevBuf<uint8_t> d_bnd, d_scan_tmp; DevBuf<int32_t> d_chain;
	DevBuf<uint8_t> d_symmap; DevBuf<uint32_t> d_symset;
	bool prof = false; uint32_t syms = 0;      // query-profile variant of K1: the targets use <= 4 distinct bytes
	bool bits = false;                         // bit-parallel edit distance: `-u 1` and reads of <= 8 distinct bytes
	bool twobit = false;                       // the sequences stay 2-bit packed in HBM (AT_SEQ_2BIT input consumed directly by K1 / K2 / K3)
	std::vector<uint64_t> h_t_rel;             // twobit + jump: device offsets of the packed targets (the jump masks copy their alignment)
	DevBuf<uint64_t> d_j_off;                  // twobit + jump: per-pair offsets into d_jmask (byte-encoded targets use d_t_off)
	BufCache cache;                            // released device blocks, reused by this shard's next allocations
	struct UploadGate *gate = nullptr; uint64_t gate_turn = 0;   // pipelined path: sub-slices upload their sequences one at a time, in pair order
	bool workspace = false;                    // pipeline workspace: reused for many sub-slices, buffers get head-room
	std::vector<uint8_t> h_rclass;
	std::vector<Chunk> chunks;
	uint64_t cells = 0, ptr_bytes = 0, t_span = 0;   // t_span: bytes of d_t that hold the caller's target span
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	cudaEvent_t evk[2] = {nullptr, nullptr};
	cudaEvent_t ev_up = nullptr;               // behind this shard's sequence upload (pipelined path: orders the sub-slices' copies)
	cudaEvent_t ev_sync = nullptr;             // shard_sync(): a sleeping wait for the stream
	// last-run timing
	double fill_ms = 0, tb_ms = 0, dev_ms = 0, domk_ms = 0; uint64_t domk_cells = 0, launches = 0; uint32_t domk_kind = 0, domk_r = 0, domk_flags = 0;
	int rc = AT_OK;
};

// How host threads wait for the device.  Default: spin with yields (the device's cudaDeviceScheduleYield flag) -- the lowest
// latency, and measured best even with 8 ranks x 3 pipeline workers on 32 cores (e2e C2 at N = 8: 37.0 ms spinning,
// 38.8 ms sleeping; profiles/bench_r02_n8*.json).  AT_SYNC=block makes the waits sleep on events created with
// cudaEventBlockingSync instead, for hosts where the cores are needed elsewhere.
static bool wait_blocking()
{
	static const bool v = [] { const char *e = getenv("AT_SYNC"); return e && (e[0] == 'b' || e[0] == 'B'); }();
	return v;
}
static cudaError_t event_create_timed(cudaEvent_t *e) { return cudaEventCreateWithFlags(e, wait_blocking() ? cudaEventBlockingSync : cudaEventDefault); }
static cudaError_t shard_sync(Shard &s)
{
	if (!wait_blocking()) return cudaStreamSynchronize(s.stream);
	if (!s.ev_sync) { cudaError_t e = cudaEventCreateWithFlags(&s.ev_sync, cudaEventBlockingSync | cudaEventDisableTiming); if (e != cudaSuccess) return e; }
	cudaError_t e = cudaEventRecord(s.ev_sync, s.stream);
	return e != cudaSuccess ? e : cudaEventSynchronize(s.ev_sync);
}

struct at_batch {
	at_handle *h = nullptr;
	int mode = 0; at_params prm; uint32_t out_flags = 0;
	uint64_t n = 0;
	bool traceback = false;
	std::vector<Shard> shards;
	bool ran = false;
};

static inline uint32_t rclass_of(uint32_t l1) { uint32_t r = (l1 + 31) / 32; return r < 1 ? 1 : (r > (uint32_t)MAXR ? MAXR : r); }

static uint64_t ptr_words_of(int mode, bool jump, uint32_t l1, uint32_t l2)
{
	const uint32_t R = rclass_of(l1), RPP = 32 * R;
	const uint64_t stripes = (l1 + RPP - 1) / RPP;
	if (mode == AT_EDIT) return 0;
	if (mode == AT_OVERLAP) { const uint32_t tl = (l2 + 31u) | 15u; return stripes * ((tl >> 4) + 1) * RPP; }
	const uint32_t tl = (l2 + 31u) | (jump ? 31u : 7u);
	uint64_t w = stripes * ((tl >> 3) + 1) * RPP;
	if (jump) w += stripes * ((tl >> 5) + 1) * RPP;
	return w;
}

__global__ void at_unpack_2bit(const uint8_t *src, const uint64_t *src_off, const uint64_t *dst_off,
                               const uint32_t *len, uint32_t n_pairs, uint8_t *dst)
{
	const uint32_t p = blockIdx.x;
	if (p >= n_pairs) return;
	const uint8_t *s = src + src_off[p];
	uint8_t *d = dst + dst_off[p];
	for (uint32_t k = threadIdx.x; k < len[p]; k += blockDim.x) {
		const uint32_t c = (s[k >> 2] >> (2 * (k & 3))) & 3u;
		d[k] = (uint8_t)("ACGT"[c]);
	}
}

// upload one side (reads or targets) of a shard; rewrites offsets relative to the device buffer
// `resident` (2-bit input only): the packed records stay as they are -- d_bytes receives them, d_off their byte
// offsets -- and the kernels read the codes directly; otherwise 2-bit input is expanded to bytes on the device.
static int upload_side(at_handle *h, Shard &s, uint32_t encoding, const uint8_t *src, const uint64_t *off,
                       const uint32_t *len, DevBuf<uint8_t> &d_bytes, DevBuf<uint8_t> &d_packed, DevBuf<uint64_t> &d_off,
                       DevBuf<uint32_t> &d_len, uint64_t *span_bytes, bool resident, std::vector<uint64_t> *keep_rel = nullptr)
{
	const uint32_t n = s.n;
	cudaStream_t st = s.stream;
	std::vector<uint64_t> rel(n), unp(n);
	uint64_t lo = UINT64_MAX, hi = 0, tot = 0;
	bool monotonic = true;
	for (uint32_t k = 0; k < n; ++k) {
		const uint64_t o = off[s.p0 + k];
		const uint64_t nbytes = encoding == AT_SEQ_2BIT ? ((uint64_t)len[s.p0 + k] + 3) / 4 : len[s.p0 + k];
		if (k && o < hi) monotonic = false;
		lo = std::min(lo, o); hi = std::max(hi, o + nbytes);
		unp[k] = tot; tot += len[s.p0 + k];
	}
	DevBuf<uint8_t> &raw = (encoding == AT_SEQ_2BIT && !resident) ? d_packed : d_bytes;
	if (monotonic && hi - lo <= 2 * tot + 64) {          // one bulk copy of the caller's span
		CU(h, raw.alloc(hi - lo + AT_SEQ_SLACK));
		CU(h, cudaMemcpyAsync(raw.p, src + lo, hi - lo, cudaMemcpyHostToDevice, st));
		for (uint32_t k = 0; k < n; ++k) rel[k] = off[s.p0 + k] - lo;
		*span_bytes = hi - lo;
	} else {                                             // scattered records: repack on the host first
		std::vector<uint8_t> stage;
		uint64_t pos = 0;
		for (uint32_t k = 0; k < n; ++k) {
			const uint64_t nbytes = encoding == AT_SEQ_2BIT ? ((uint64_t)len[s.p0 + k] + 3) / 4 : len[s.p0 + k];
			rel[k] = pos; stage.resize(pos + nbytes);
			memcpy(stage.data() + pos, src + off[s.p0 + k], nbytes); pos += nbytes;
		}
		CU(h, raw.alloc(pos + AT_SEQ_SLACK));
		CU(h, cudaMemcpyAsync(raw.p, stage.data(), pos, cudaMemcpyHostToDevice, st));      // pageable source: staged before the call returns
		*span_bytes = pos;
	}
	CU(h, d_off.alloc(n)); CU(h, d_len.alloc(n));
	CU(h, cudaMemcpyAsync(d_len.p, len + s.p0, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
	if (encoding == AT_SEQ_2BIT && !resident) {
		DevBuf<uint64_t> d_src_off;
		CU(h, d_src_off.alloc(n));
		CU(h, cudaMemcpyAsync(d_src_off.p, rel.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
		CU(h, cudaMemcpyAsync(d_off.p, unp.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
		CU(h, d_bytes.alloc(tot + AT_SEQ_SLACK));
		at_unpack_2bit<<<n, 128, 0, st>>>(d_packed.p, d_src_off.p, d_off.p, d_len.p, n, d_bytes.p);
		CU(h, cudaGetLastError());
		h->launches++;
		d_src_off.release();      // stream-ordered: reused only by later work of this stream
		*span_bytes = tot;
	} else {
		CU(h, cudaMemcpyAsync(d_off.p, rel.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
		if (keep_rel) keep_rel->swap(rel);      // (the copy above has staged the pageable vector already)
	}
	// No synchronisation here: `rel` / `unp` / `stage` are pageable, so the runtime has staged them by the time
	// cudaMemcpyAsync returns; the caller's (possibly pinned) sequence buffers stay valid until setup_shard's one
	// synchronisation at its end.
	return AT_OK;
}

// give the per-chunk buffers back to the pool (stream-ordered: the same stream reuses them at once)
static void release_chunks(Shard &s)
{
	for (auto &c : s.chunks) {
		c.d_ptr_off.release(); c.d_bnd_off.release(); for (auto &l : c.launches) { l.d_jobs.release(); l.d_tasks.release(); }
		c.d_ops_off.release(); c.d_cols_off.release(); c.d_scratch_off.release(); c.d_cigar.release(); c.d_aln1.release(); c.d_aln2.release();
	}
	s.chunks.clear();
}

static void free_shard(Shard &s)
{
	if (s.dev) { cudaSetDevice(s.dev->id); tl_stream = s.stream; tl_big = s.dev->big.get(); }
	tl_cache = &s.cache;
	s.d_q.release(); s.d_t.release(); s.d_jmask.release(); s.d_j_off.release(); s.d_rclass.release(); s.d_end_state.release();
	s.d_q2.release(); s.d_t2.release();
	s.d_q_off.release(); s.d_t_off.release(); s.d_site_off.release();
	s.d_q_len.release(); s.d_t_len.release(); s.d_end_i.release(); s.d_end_j.release(); s.d_beg_i.release();
	s.d_beg_j.release(); s.d_n_ops.release(); s.d_n_cols.release(); s.d_counter.release();
	s.ws_cigar.release(); s.ws_aln1.release(); s.ws_aln2.release();
	s.d_score.release(); s.d_sites.release(); s.d_ptr.release(); s.d_scratch.release(); s.d_bnd.release(); s.d_scan_tmp.release();
	s.d_prog.release(); s.d_chain.release(); s.d_symmap.release(); s.d_symset.release();
	release_chunks(s);
	for (auto &e : s.ev) if (e) cudaEventDestroy(e);
	for (auto &e : s.evk) if (e) cudaEventDestroy(e);
	if (s.ev_up) cudaEventDestroy(s.ev_up);
	if (s.ev_sync) cudaEventDestroy(s.ev_sync);
	s.cache.flush(s.stream);
	tl_cache = nullptr; tl_big = nullptr;
}

extern "C" void at_batch_free(at_batch *b)
{
	if (!b) return;
	for (auto &s : b->shards) free_shard(s);
	delete b;
}

// Which bytes occur in a device buffer (at_symbol_set): 256-bit set, read back to the host.
static int scan_alphabet(at_handle *h, Shard &s, const uint8_t *d_bytes, uint64_t n_bytes, uint32_t set8[8])
{
	cudaStream_t st = s.stream;
	CU(h, s.d_symset.alloc(8));
	CU(h, cudaMemsetAsync(s.d_symset.p, 0, 8 * sizeof(uint32_t), st));
	at_symbol_set<<<(int)std::min<uint64_t>(s.dev->sm_count * 8, (n_bytes + 4095) / 4096 + 1), 256, 0, st>>>(d_bytes, n_bytes, s.d_symset.p);
	CU(h, cudaGetLastError());
	h->launches++;
	CU(h, cudaMemcpyAsync(set8, s.d_symset.p, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	CU(h, shard_sync(s));
	return AT_OK;
}

static int upload_symmap(at_handle *h, Shard &s, const uint8_t map[256])
{
	CU(h, s.d_symmap.alloc(256));
	CU(h, cudaMemcpyAsync(s.d_symmap.p, map, 256, cudaMemcpyHostToDevice, s.stream));      // `map` is pageable stack memory: staged before the call returns
	return AT_OK;
}

// Kernel variants of the shard, from the alphabets of its sequences:
//   s.prof  query-profile variant of K1 / K2 -- the TARGETS use at most 4 distinct bytes (symmap: target byte -> 0..3);
//   s.bits  bit-parallel kernel for `edit -u 1` -- the READS use at most 8 distinct bytes (symmap: read byte -> 0..7,
//           anything else -> 8).  The two are exclusive: symmap describes one side.
static int choose_variants(at_batch *b, Shard &s, const at_batch_input *in, uint64_t q_span)
{
	at_handle *h = b->h;
	int rc;
	s.prof = false; s.bits = false; s.syms = 0;
	const bool two_bit = in->encoding == AT_SEQ_2BIT;                  // the alphabet is ACGT by construction
	uint32_t set8[8];
	uint8_t map[256];
	if (b->mode == AT_EDIT && b->prm.u == 1 && !getenv("AT_NO_BITPAR")) {
		memset(map, 8, sizeof map);
		int nsym = 0;
		if (two_bit) { map['A'] = 0; map['C'] = 1; map['G'] = 2; map['T'] = 3; nsym = 4; }
		else {
			if ((rc = scan_alphabet(h, s, s.d_q.p, q_span, set8))) return rc;
			for (int c = 0; c < 256; ++c)
				if (set8[c >> 5] >> (c & 31) & 1u) { if (nsym < 8) map[c] = (uint8_t)nsym; ++nsym; }
		}
		if (nsym >= 1 && nsym <= 8) { s.bits = true; return upload_symmap(h, s, map); }
		return AT_OK;                                                  // large read alphabet: cell-by-cell kernel, xor/min variant
	}
	if (getenv("AT_NO_PROFILE")) return AT_OK;
	memset(map, 0, sizeof map);
	int nsym = 0; uint32_t syms = 0;
	if (two_bit) { map['C'] = 1; map['G'] = 2; map['T'] = 3; nsym = 4; syms = (uint32_t)'A' | ((uint32_t)'C' << 8) | ((uint32_t)'G' << 16) | ((uint32_t)'T' << 24); }
	else {
		// d_t holds the caller's span: every byte of it is a target byte or a gap between records (which can only add symbols)
		if ((rc = scan_alphabet(h, s, s.d_t.p, s.t_span, set8))) return rc;
		for (int c = 0; c < 256; ++c)
			if (set8[c >> 5] >> (c & 31) & 1u) { if (nsym < 4) { map[c] = (uint8_t)nsym; syms |= (uint32_t)c << (8 * nsym); } ++nsym; }
	}
	if (nsym < 1 || nsym > 4) return AT_OK;
	for (int c = nsym; c < 4; ++c) syms |= (syms & 255u) << (8 * c);   // unused codes repeat a used symbol (they are never looked up)
	s.syms = syms; s.prof = true;
	return upload_symmap(h, s, map);
}

// fit+jump: the per-pair blacklists as a byte mask aligned with the target bytes (at_build_jmask)
static int build_jump_mask(at_batch *b, Shard &s, const at_batch_input *in)
{
	at_handle *h = b->h;
	cudaStream_t st = s.stream;
	const uint32_t n = s.n;
	// params.jump == 2 (junction WHITELIST, the semantics of the comment at src/alignment.h:542-544): entering J is
	// barred everywhere except on the listed indices -- the mask starts as all ones and the sites clear it
	const bool whitelist = b->prm.jump == 2;
	// byte-encoded targets: the mask is indexed like the targets (same offsets, same 16-byte alignment for the TMA tiles of
	// K2).  2-bit resident targets: one mask byte per SYMBOL, each pair's mask placed so that it has the alignment K2's
	// ring gives the expanded target -- 4 x (packed address mod 16) past a 16-byte boundary.
	uint64_t mask_bytes = s.d_t.n;
	if (s.twobit) {
		std::vector<uint64_t> jo(n);
		uint64_t cur = 0;
		for (uint32_t k = 0; k < n; ++k) {
			jo[k] = cur + 4 * (((uintptr_t)s.d_t.p + s.h_t_rel[k]) & 15u);
			cur = (jo[k] + in->t_len[s.p0 + k] + 15u) & ~(uint64_t)15u;
		}
		mask_bytes = cur + AT_SEQ_SLACK;
		CU(h, s.d_j_off.alloc(n));
		CU(h, cudaMemcpyAsync(s.d_j_off.p, jo.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
	}
	const uint64_t *d_j_off = s.twobit ? s.d_j_off.p : s.d_t_off.p;
	CU(h, s.d_jmask.alloc(mask_bytes));
	CU(h, cudaMemsetAsync(s.d_jmask.p, whitelist ? 1 : 0, mask_bytes, st));
	if (!in->sites || !in->site_off) return AT_OK;
	const uint64_t lo = in->site_off[s.p0], hi = in->site_off[s.p1];
	std::vector<uint64_t> so(n + 1);
	for (uint32_t k = 0; k <= n; ++k) so[k] = in->site_off[s.p0 + k] - lo;
	CU(h, s.d_sites.alloc(hi - lo + 1)); CU(h, s.d_site_off.alloc(n + 1));
	if (hi > lo) CU(h, cudaMemcpyAsync(s.d_sites.p, in->sites + lo, (hi - lo) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
	CU(h, cudaMemcpyAsync(s.d_site_off.p, so.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
	at_build_jmask<<<n, 64, 0, st>>>(s.d_sites.p, s.d_site_off.p, d_j_off, s.d_t_len.p, n, s.d_jmask.p, whitelist ? 0 : 1);
	CU(h, cudaGetLastError());
	h->launches++;
	return AT_OK;
}


// Queue order of the K2 tasks of one launch.  The persistent warps claim tasks in queue order, and a stripe can only run
// as far as the stripe above it has got.  Pair after pair (round 1) puts a pair's stripes next to each other in the
// queue: they are claimed within microseconds of each other and run as ONE tightly coupled chain, each stripe 64 columns
// behind the one above -- the chain advances at the pace of its momentarily slowest warp and the others poll (ncu on
// C3, 2 048 pairs: 21 % of all warp samples in the progress poll and its warp-sync).  Stripe after stripe ACROSS the
// launch's pairs -- all stripes 0, then all stripes 1, ... -- puts as many other tasks between a stripe and its
// predecessor as the launch has pairs: with more pairs than resident warps the predecessor has finished (or is far
// ahead) when the stripe is claimed and nothing ever polls; with fewer pairs the chains are as long as they have to be
// to fill the GPU and no longer.  A task's predecessor still precedes it in the queue, so a claimed task only ever
// waits for a running one (deadlock freedom as before).  In: tasks collected pair by pair, stripes ascending.
// AT_WAVE_ORDER=pair keeps the old order (A/B runs).
// AT_WAVE_START_LAG (columns, default 0): a K2 stripe starts only once the stripe above it is this far ahead (A/B knob)
static uint32_t wave_start_lag() { static const uint32_t v = [] { const char *e = getenv("AT_WAVE_START_LAG"); return e && *e ? (uint32_t)strtoul(e, nullptr, 10) : 0u; }(); return v; }

static void order_wave_tasks(std::vector<WaveTask> &t)
{
	static const bool pair_major = [] { const char *e = getenv("AT_WAVE_ORDER"); return e && !strcmp(e, "pair"); }();
	const size_t n = t.size();
	if (pair_major || n == 0) {
		for (size_t x = 0; x < n; ++x) t[x].prev = (uint32_t)(t[x].stripe ? x - 1 : x);
		return;
	}
	struct Run { uint32_t pair, n, last; };
	std::vector<Run> active;
	for (size_t x = 0; x < n;) { size_t y = x; while (y < n && t[y].pair == t[x].pair) ++y; active.push_back(Run{t[x].pair, (uint32_t)(y - x), 0}); x = y; }
	std::vector<WaveTask> out;
	out.reserve(n);
	for (uint32_t s = 0; !active.empty(); ++s) {
		size_t keep = 0;
		for (Run &r : active) {
			const uint32_t idx = (uint32_t)out.size();
			out.push_back(WaveTask{r.pair, s, s ? r.last : idx, 0});
			r.last = idx;
			if (s + 1 < r.n) active[keep++] = r;
		}
		active.resize(keep);
	}
	t.swap(out);
}

static int setup_shard(at_batch *b, Shard &s, const at_batch_input *in)
{
	at_handle *h = b->h;
	CU(h, cudaSetDevice(s.dev->id));
	cudaStream_t st = s.stream;
	tl_stream = st;
	tl_cache = &s.cache; tl_big = s.dev->big.get();
	const uint32_t n = s.n;
	int rc;
	// AT_PIPE_TRACE=2: host timeline of this function's phases (ms since entry) on stderr
	static const bool trace_setup = getenv("AT_PIPE_TRACE") && atoi(getenv("AT_PIPE_TRACE")) >= 2;
	const auto t_enter = std::chrono::steady_clock::now();
	std::string tl_;
	auto mark = [&](const char *what) {
		if (!trace_setup) return;
		char buf[64];
		snprintf(buf, sizeof buf, " %s %.2f", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enter).count());
		tl_ += buf;
	};
	uint64_t q_span = 0;
	{
		// a pipeline worker waits for the uploads of the sub-slices before its own: the first (small) sub-slice's
		// sequences are not held up behind a later, larger one sharing the copy engine, and its fill starts early
		// 2-bit input stays packed in HBM: K1 reads the codes with 128-bit / byte loads (the code IS the query-profile
		// index), K2 stages 64-byte tiles by TMA and expands them into its shared-memory ring once per 256 columns, the
		// traceback decodes them; no expansion kernel, a quarter of the sequence bytes in HBM and over PCIe.  (The alphabet
		// is ACGT by construction, so the query-profile variants always apply; AT_NO_PROFILE falls back to expanding.)
		s.twobit = in->encoding == AT_SEQ_2BIT && !getenv("AT_NO_PROFILE") && !getenv("AT_NO_2BIT_RESIDENT");
		if (s.gate) {
			s.gate->wait_for(s.gate_turn);
			if (s.gate->last) CU(h, cudaStreamWaitEvent(st, s.gate->last, 0));      // copy engine: earlier sub-slices first
		}
		rc = upload_side(h, s, in->encoding, in->q, in->q_off, in->q_len, s.d_q, s.d_q2, s.d_q_off, s.d_q_len, &q_span, s.twobit);
		if (!rc) rc = upload_side(h, s, in->encoding, in->t, in->t_off, in->t_len, s.d_t, s.d_t2, s.d_t_off, s.d_t_len, &s.t_span, s.twobit,
		                          s.twobit && b->mode == AT_FIT && b->prm.jump ? &s.h_t_rel : nullptr);
		if (!s.ev_up) CU(h, cudaEventCreateWithFlags(&s.ev_up, cudaEventDisableTiming | (wait_blocking() ? cudaEventBlockingSync : 0)));
		if (!rc) CU(h, cudaEventRecord(s.ev_up, st));      // behind the sequence copies: what must complete before the caller's buffers are free
		if (s.gate) {
			if (!rc) s.gate->last = s.ev_up;
			s.gate->pass(s.gate_turn);
		}
		if (rc) return rc;
	}
	s.d_q2.release(); s.d_t2.release();
	mark("upload");
	const bool jump = b->mode == AT_FIT && b->prm.jump;
	if ((rc = choose_variants(b, s, in, q_span))) return rc;
	if (jump && (rc = build_jump_mask(b, s, in))) return rc;
	mark("alphabet+jmask");
	// a UNIFORM shard (every pair of the same shape, all on K1) needs no per-pair work on the host at all: see below
	const uint32_t l1u = in->q_len[s.p0], l2u = in->t_len[s.p0];
	bool uniform = b->mode <= AT_FIT && l1u <= 32u * MAXR && !getenv("AT_NO_UNIFORM_PLAN");
	for (uint32_t k = 1; k < n && uniform; ++k) uniform = in->q_len[s.p0 + k] == l1u && in->t_len[s.p0 + k] == l2u;
	// per-pair class, result arrays
	s.cells = 0;
	if (uniform) s.cells = (uint64_t)n * l1u * l2u;
	else {
		s.h_rclass.resize(n);
		for (uint32_t k = 0; k < n; ++k) {
			const uint32_t l1 = in->q_len[s.p0 + k], l2 = in->t_len[s.p0 + k];
			s.h_rclass[k] = (uint8_t)rclass_of(l1);
			s.cells += (uint64_t)l1 * l2;
		}
	}
	CU(h, s.d_score.alloc(n)); CU(h, s.d_end_i.alloc(n)); CU(h, s.d_end_j.alloc(n)); CU(h, s.d_end_state.alloc(n));
	CU(h, s.d_beg_i.alloc(n)); CU(h, s.d_beg_j.alloc(n)); CU(h, s.d_n_ops.alloc(n + 1)); CU(h, s.d_n_cols.alloc(n + 1));
	CU(h, cudaMemsetAsync(s.d_n_ops.p, 0, (n + 1) * sizeof(uint32_t), st));
	CU(h, cudaMemsetAsync(s.d_n_cols.p, 0, (n + 1) * sizeof(uint32_t), st));
	CU(h, cudaMemsetAsync(s.d_beg_i.p, 0, n * sizeof(uint32_t), st));
	CU(h, cudaMemsetAsync(s.d_beg_j.p, 0, n * sizeof(uint32_t), st));
	CU(h, s.d_counter.alloc(64));

	mark("results");
	// ---- chunking by pointer-arena budget ----
	// The whole shard's pointer blocks in one chunk if they fit the arena this shard already owns (a
	// pipeline workspace after its first sub-slice); otherwise ask the driver what is free.
	// cudaMemGetInfo is kept off the steady-state path: with other threads driving the device it was
	// seen to block for tens of milliseconds.
	uint64_t need_words = 0;
	auto pair_words = [&](uint32_t k) -> uint64_t {
		const uint32_t l1 = in->q_len[s.p0 + k], l2 = in->t_len[s.p0 + k];
		return b->traceback ? ptr_words_of(b->mode, jump, l1, l2) + 64ull * rclass_of(l1) : 0;   // + rounding slack of the packed layout
	};
	if (uniform) need_words = (uint64_t)n * pair_words(0);
	else for (uint32_t k = 0; k < n; ++k) need_words += pair_words(k);
	size_t free_b = 0, total_b = 0;
	const bool owned = need_words <= s.d_ptr.n && !getenv("AT_PTR_BUDGET_MB");
	if (!owned && need_words) {
		CU(h, cudaMemGetInfo(&free_b, &total_b));
		// blocks cached in the stream-ordered pool are reusable by this batch: count them as free
		cudaMemPool_t pool; uint64_t reserved = 0, used = 0;
		if (cudaDeviceGetDefaultMemPool(&pool, s.dev->id) == cudaSuccess &&
		    cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
		    cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
			free_b += (size_t)(reserved - used);
		free_b += s.d_ptr.bytes;      // the arena this shard holds is re-sized, not added to
		free_b += s.dev->big->cached_bytes();      // arena-sized blocks cached in the handle are dropped on demand
	}
	uint64_t budget_words = owned ? std::max<uint64_t>(s.d_ptr.n, 1) : (uint64_t)(free_b * 0.45) / 4;
	if (!need_words) budget_words = UINT64_MAX;
	if (const char *env = getenv("AT_PTR_BUDGET_MB")) budget_words = (uint64_t)atoll(env) * (1ull << 20) / 4;
	// packed s16x2 lanes (two pairs per warp): local mode, scores x8 must fit int16, 8|m-u| < 256 (symbols are << 8)
	const int64_t maxabs = std::max<int64_t>({llabs((long long)b->prm.m), llabs((long long)b->prm.u), llabs((long long)b->prm.o), llabs((long long)b->prm.e), 1});
	// global / fit (without jump state) also carry -inf inside the 16 bits (AT_NEG16 = -30000, at_cell.cuh): finite values must
	// stay above -24000 and a -inf value may drift by a few steps' worth of the largest parameter
	const bool p16_local = b->mode == AT_LOCAL;
	const bool p16_mode = (p16_local || b->mode == AT_GLOBAL || (b->mode == AT_FIT && !jump)) && llabs((long long)b->prm.m - b->prm.u) <= 31 &&
	                      (p16_local || maxabs <= 31) && !getenv("AT_NO_P16");
	auto p16_ok = [&](uint32_t l1, uint32_t l2) {
		return p16_mode && l1 <= 32u * MAXR && l2 <= 60000u && 8 * (int64_t)(l1 + l2 + 42) * maxabs < (p16_local ? 32000 : 24000);      // + 40: K1 also computes up to 34 columns past l2
	};
	release_chunks(s);      // a pipeline worker reuses its shard (and the shard-level buffers) for every sub-slice
	// ---- uniform shard: every pair has the same shape and runs on K1 in one chunk -> the plan is closed form ----
	{
		if (uniform && need_words && need_words > budget_words) {      // several chunks: the general plan (it needs the per-pair classes)
			uniform = false;
			s.h_rclass.assign(n, (uint8_t)rclass_of(l1u));
		}
		if (uniform) {
			const bool packed = p16_ok(l1u, l2u) && n >= 2;
			const uint32_t R = rclass_of(l1u);
			const uint32_t tl = (l2u + 31u) | (packed ? 3u : (jump ? 31u : 7u));
			const uint64_t words_per_job = !b->traceback ? 0 : packed ? (uint64_t)((tl >> 2) + 1) * R * 32 : ptr_words_of(b->mode, jump, l1u, l2u);
			const uint64_t n_jobs = packed ? (n + 1) / 2 : n;
			s.chunks.emplace_back();
			Chunk &c = s.chunks.back();
			c.k0 = 0; c.k1 = n;
			c.ptr_words = words_per_job * n_jobs;
			c.scratch_words = b->traceback ? (uint64_t)n * (l1u + l2u) : 0;
			c.launches.emplace_back();
			Launch &l = c.launches.back();
			l.kind = packed ? LK_PACKED : LK_INT32; l.r = (int)R; l.cells = (uint64_t)n * l1u * l2u; l.n_planned = n_jobs;
			CU(h, l.d_jobs.alloc(n_jobs)); CU(h, c.d_ptr_off.alloc(n)); CU(h, c.d_bnd_off.alloc(n)); CU(h, s.d_rclass.alloc(n));
			if (b->traceback) { CU(h, c.d_ops_off.alloc(n + 1)); CU(h, c.d_cols_off.alloc(n + 1)); CU(h, c.d_scratch_off.alloc(n)); }
			// on this stream, behind the uploads; the host does not wait for it (it may sit behind another sub-slice's fill)
			at_plan_uniform<<<(n + 255) / 256, 256, 0, st>>>(n, packed ? 1 : 0, R, words_per_job, (uint64_t)l1u + l2u, l.d_jobs.p, c.d_ptr_off.p, c.d_bnd_off.p,
			                                                 b->traceback ? c.d_scratch_off.p : nullptr, s.d_rclass.p);
			CU(h, cudaGetLastError());
			h->launches++;
			s.ptr_bytes = c.ptr_words * 4;
			CU(h, s.d_bnd.alloc(64)); CU(h, s.d_prog.alloc(1)); CU(h, s.d_chain.alloc(4ull * n + 4));
			if (c.scratch_words && s.d_scratch.alloc(c.scratch_words) != cudaSuccess) { set_err(h, "traceback scratch of %llu MB", (unsigned long long)(c.scratch_words >> 18)); return AT_E_NOMEM; }
			if (c.ptr_words) {
				cudaError_t e = s.d_ptr.alloc(c.ptr_words + (s.workspace && c.ptr_words > s.d_ptr.n ? c.ptr_words / 16 : 0));
				if (e != cudaSuccess) { set_err(h, "pointer arena of %llu MB: %s", (unsigned long long)(c.ptr_words >> 18), cudaGetErrorString(e)); return AT_E_NOMEM; }
			}
			for (auto &e : s.ev) if (!e) CU(h, event_create_timed(&e));
			for (auto &e : s.evk) if (!e) CU(h, event_create_timed(&e));
			mark("uniform plan");
			CU(h, cudaEventSynchronize(s.ev_up));      // the caller's buffers have been read; the plan kernel may still be queued
			mark("sync");
			if (trace_setup) fprintf(stderr, "[at setup] pairs %u:%s\n", n, tl_.c_str());
			return AT_OK;
		}
	}
	uint64_t max_chunk_words = 0, max_scratch_words = 0;
	{
		Chunk cur; cur.k0 = 0;
		uint64_t words = 0;
		for (uint32_t k = 0; k < n; ++k) {
			const uint64_t w = pair_words(k);
			if (w > budget_words) { set_err(h, "pair %llu needs %llu MB of traceback pointers; arena budget is %llu MB",
			                                (unsigned long long)(s.p0 + k), (unsigned long long)(w >> 18), (unsigned long long)(budget_words >> 18)); return AT_E_NOMEM; }
			if (words + w > budget_words && k > cur.k0) {
				cur.k1 = k; s.chunks.push_back(std::move(cur));
				cur = Chunk(); cur.k0 = k; words = 0;
			}
			words += w;
		}
		cur.k1 = n; s.chunks.push_back(std::move(cur));
	}
	mark("meminfo+chunks");
	s.ptr_bytes = 0;
	const bool linear = b->mode >= AT_OVERLAP;          // single-plane modes run on K2 at every length
	uint64_t max_bnd_elems = 0, max_prog_words = 0;
	for (auto &c : s.chunks) {
		const uint32_t nc = c.k1 - c.k0;
		auto cells_of = [&](uint32_t k) { return (uint64_t)in->q_len[s.p0 + k] * in->t_len[s.p0 + k]; };
		auto by_cells = [&](std::vector<uint32_t> &v) {   // largest pairs first (dynamic queue => good tail balance)
			bool ragged = false;
			for (size_t x = 1; x < v.size() && !ragged; ++x) ragged = cells_of(v[x]) != cells_of(v[0]);
			if (ragged) std::stable_sort(v.begin(), v.end(), [&](uint32_t x, uint32_t y) { return cells_of(x) > cells_of(y); });
		};
		Launch l32[MAXR + 1], l16[MAXR + 1], lwv[MAXR + 1], lbit[MAXR + 1];
		std::vector<uint32_t> scalar_pairs, wave_pairs, cand;
		for (uint32_t k = c.k0; k < c.k1; ++k) {
			const uint32_t l1 = in->q_len[s.p0 + k], l2 = in->t_len[s.p0 + k];
			if (linear || l1 > 32u * MAXR) wave_pairs.push_back(k);
			else if (p16_ok(l1, l2)) cand.push_back(k);
			else scalar_pairs.push_back(k);
		}
		// ---- packed jobs (K1, s16x2): partners must share the rows-per-lane class and l2 ----
		{
			// local: the partners share the rows-per-lane class and l2; global / fit: l1 as well (one lane holds the last row of both)
			auto key = [&](uint32_t k) { return ((uint64_t)(p16_local ? s.h_rclass[k] : in->q_len[s.p0 + k]) << 32) | in->t_len[s.p0 + k]; };
			bool sorted = true;
			for (size_t x = 1; x < cand.size() && sorted; ++x) sorted = key(cand[x - 1]) >= key(cand[x]);
			if (!sorted) std::stable_sort(cand.begin(), cand.end(), [&](uint32_t x, uint32_t y) { return key(x) > key(y); });
			for (size_t x = 0; x < cand.size();) {
				if (x + 1 < cand.size() && key(cand[x]) == key(cand[x + 1])) {
					const int r = s.h_rclass[cand[x]] & 15;
					l16[r].h_jobs.push_back(FillJob{cand[x], cand[x + 1]});
					l16[r].cells += cells_of(cand[x]) + cells_of(cand[x + 1]);
					x += 2;
				} else { scalar_pairs.push_back(cand[x]); x += 1; }
			}
		}
		// ---- int32 jobs (K1) per R class ----
		by_cells(scalar_pairs);
		for (uint32_t k : scalar_pairs) { const int r = s.h_rclass[k] & 15; l32[r].h_jobs.push_back(FillJob{k, k}); l32[r].cells += cells_of(k); }
		// ---- wavefront tasks (K2): (pair, stripe), collected pair by pair, then put in queue order (order_wave_tasks) ----
		by_cells(wave_pairs);
		c.h_bnd_off.assign(nc, 0);
		c.bnd_elems = 0; c.prog_words = 0;
		// bit-parallel edit distance: r = 32-row blocks per lane, 1024*r rows per stripe.  One column step of a
		// lane is a serial chain through its r blocks, so few long pairs want r = 1 (more stripes in flight);
		// 4 blocks per lane amortise the per-step overhead once there are tasks for every resident warp.
		int bits_r = 1;
		if (s.bits) {
			const uint64_t warp_slots = (uint64_t)s.dev->sm_count * 8 * AT_WAVE_WARPS;
			for (int r : {4, 2}) {
				uint64_t tasks = 0;
				for (uint32_t k : wave_pairs) tasks += (in->q_len[s.p0 + k] + 1024u * r - 1) / (1024u * r);
				if (tasks >= 2 * warp_slots) { bits_r = r; break; }
			}
		}
		for (uint32_t k : wave_pairs) {
			const uint32_t l1 = in->q_len[s.p0 + k], l2 = in->t_len[s.p0 + k];
			const int r = s.bits ? bits_r : (s.h_rclass[k] & 15);
			const uint32_t rows = s.bits ? 1024u * r : 32u * r;
			Launch &lw = s.bits ? lbit[r] : lwv[r];
			const uint32_t n_stripes = (l1 + rows - 1) / rows;
			for (uint32_t st2 = 0; st2 < n_stripes; ++st2) lw.h_tasks.push_back(WaveTask{k, st2, 0, 0});
			lw.cells += cells_of(k);
			if (n_stripes > 1) { c.h_bnd_off[k - c.k0] = c.bnd_elems; c.bnd_elems += 2ull * ((l2 + 4u) & ~3u); }
		}
		for (int r = 1; r <= MAXR; ++r) {
			if (!l32[r].h_jobs.empty()) { l32[r].kind = LK_INT32; l32[r].r = r; c.launches.push_back(std::move(l32[r])); }
			if (!l16[r].h_jobs.empty()) { l16[r].kind = LK_PACKED; l16[r].r = r; c.launches.push_back(std::move(l16[r])); }
			order_wave_tasks(lwv[r].h_tasks); order_wave_tasks(lbit[r].h_tasks);
			if (!lwv[r].h_tasks.empty()) {
				lwv[r].kind = LK_WAVE; lwv[r].r = r; lwv[r].prog_base = c.prog_words; c.prog_words += lwv[r].h_tasks.size();
				c.launches.push_back(std::move(lwv[r]));
			}
			if (!lbit[r].h_tasks.empty()) {
				lbit[r].kind = LK_BITS; lbit[r].r = r; lbit[r].prog_base = c.prog_words; c.prog_words += lbit[r].h_tasks.size();
				c.launches.push_back(std::move(lbit[r]));
			}
		}
		max_bnd_elems = std::max(max_bnd_elems, c.bnd_elems);
		max_prog_words = std::max(max_prog_words, c.prog_words);
		// ---- pointer blocks: one per int32 / wavefront pair, one per packed job (shared by its two pairs) ----
		c.h_ptr_off.assign(nc, 0);
		uint64_t words = 0;
		if (b->traceback) {
			for (const Launch &l : c.launches) {
				if (l.kind == LK_PACKED) {
					for (const FillJob &jb : l.h_jobs) {
						const uint32_t tl = (in->t_len[s.p0 + jb.a] + 31u) | 3u;
						c.h_ptr_off[jb.a - c.k0] = words; c.h_ptr_off[jb.b - c.k0] = words;
						s.h_rclass[jb.a] = (uint8_t)(l.r | (1 << 4)); s.h_rclass[jb.b] = (uint8_t)(l.r | (2 << 4));
						words += (uint64_t)((tl >> 2) + 1) * l.r * 32;
					}
				} else if (l.kind == LK_INT32) {
					for (const FillJob &jb : l.h_jobs) { const uint32_t k = jb.a; c.h_ptr_off[k - c.k0] = words; words += ptr_words_of(b->mode, jump, in->q_len[s.p0 + k], in->t_len[s.p0 + k]); }
				} else {
					for (const WaveTask &tk : l.h_tasks) if (tk.stripe == 0) { const uint32_t k = tk.pair; c.h_ptr_off[k - c.k0] = words; words += ptr_words_of(b->mode, jump, in->q_len[s.p0 + k], in->t_len[s.p0 + k]); }
				}
			}
		}
		c.ptr_words = words;
		max_chunk_words = std::max(max_chunk_words, c.ptr_words);
		s.ptr_bytes += c.ptr_words * 4;
		for (Launch &l : c.launches) {
			if (l.kind >= LK_WAVE) {
				CU(h, l.d_tasks.alloc(l.h_tasks.size()));
				CU(h, cudaMemcpyAsync(l.d_tasks.p, l.h_tasks.data(), l.h_tasks.size() * sizeof(WaveTask), cudaMemcpyHostToDevice, st));
			} else {
				CU(h, l.d_jobs.alloc(l.h_jobs.size()));
				CU(h, cudaMemcpyAsync(l.d_jobs.p, l.h_jobs.data(), l.h_jobs.size() * sizeof(FillJob), cudaMemcpyHostToDevice, st));
			}
		}
		CU(h, c.d_ptr_off.alloc(nc));
		CU(h, cudaMemcpyAsync(c.d_ptr_off.p, c.h_ptr_off.data(), nc * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
		CU(h, c.d_bnd_off.alloc(nc));
		CU(h, cudaMemcpyAsync(c.d_bnd_off.p, c.h_bnd_off.data(), nc * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
		if (b->traceback) {
			CU(h, c.d_ops_off.alloc(nc + 1)); CU(h, c.d_cols_off.alloc(nc + 1));
			// traceback scratch: a slot of l1+l2 run-length ops per pair (an alignment has at most l1+l2 columns)
			c.h_scratch_off.resize(nc);
			uint64_t sw = 0;
			for (uint32_t k = 0; k < nc; ++k) { c.h_scratch_off[k] = sw; sw += (uint64_t)in->q_len[s.p0 + c.k0 + k] + in->t_len[s.p0 + c.k0 + k]; }
			c.scratch_words = sw;
			max_scratch_words = std::max(max_scratch_words, sw);
			CU(h, c.d_scratch_off.alloc(nc));
			CU(h, cudaMemcpyAsync(c.d_scratch_off.p, c.h_scratch_off.data(), nc * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
		}
	}
	mark("plan+jobs");
	if (max_bnd_elems && s.d_bnd.alloc(max_bnd_elems * (linear ? sizeof(uint64_t) : sizeof(int4)) + 64) != cudaSuccess) { set_err(h, "stripe boundary slabs of %llu MB", (unsigned long long)((max_bnd_elems * (linear ? 8 : 16)) >> 20)); return AT_E_NOMEM; }
	if (!max_bnd_elems) CU(h, s.d_bnd.alloc(64));
	CU(h, s.d_prog.alloc(max_prog_words + 1));
	{
		uint32_t max_nc = 0;
		for (auto &c : s.chunks) max_nc = std::max(max_nc, c.k1 - c.k0);
		CU(h, s.d_chain.alloc(4ull * max_nc + 4));
	}
	if (max_scratch_words && s.d_scratch.alloc(max_scratch_words) != cudaSuccess) { set_err(h, "traceback scratch of %llu MB", (unsigned long long)(max_scratch_words >> 18)); return AT_E_NOMEM; }
	CU(h, s.d_rclass.alloc(n));
	CU(h, cudaMemcpyAsync(s.d_rclass.p, s.h_rclass.data(), n, cudaMemcpyHostToDevice, st));
	if (max_chunk_words) {
		cudaError_t e = s.d_ptr.alloc(max_chunk_words + (s.workspace && max_chunk_words > s.d_ptr.n ? max_chunk_words / 16 : 0));
		if (e != cudaSuccess) { set_err(h, "pointer arena of %llu MB: %s", (unsigned long long)(max_chunk_words >> 18), cudaGetErrorString(e)); return AT_E_NOMEM; }
	}
	for (auto &e : s.ev) if (!e) CU(h, event_create_timed(&e));
	for (auto &e : s.evk) if (!e) CU(h, event_create_timed(&e));
	mark("arena");
	// the ONE synchronisation of the set-up: every upload has left the caller's buffers (they may be reused or freed once
	// at_batch_create / the sub-slice's set-up returns) and the shard's own host vectors
	CU(h, shard_sync(s));
	mark("sync");
	if (trace_setup) fprintf(stderr, "[at setup] pairs %u:%s\n", n, tl_.c_str());
	return AT_OK;
}

// argument checks shared by at_batch_create and the pipelined at_batch_align; mirrors the reference's own failure modes
// `prefix` (optional): receives the running cell counts, prefix[k] = sum of l1*l2 over pairs < k (n_pairs + 1 entries),
// so that the one-shot path validates, counts and slices a million-pair batch in one pass
// `wave_tasks` (optional): number of K2 tasks -- (pair, stripe of 256 rows) -- the batch will run as, 0 for pairs that go to K1:
// the one-shot path sizes its sub-slices so that each still fills the GPU several times over
static int validate_batch(at_handle *h, int mode, const at_params *p, const at_batch_input *in, std::vector<uint64_t> *prefix = nullptr,
                          uint64_t *wave_tasks = nullptr)
{
	if (mode < AT_GLOBAL || mode > AT_EDIT) { set_err(h, "bad mode %d", mode); return AT_E_ARG; }
	if (!in->n_pairs || !in->q || !in->q_off || !in->q_len || !in->t || !in->t_off || !in->t_len) { set_err(h, "align: parameter error"); return AT_E_ARG; }
	if (in->encoding != AT_SEQ_BYTES && in->encoding != AT_SEQ_2BIT) return AT_E_ARG;
	if ((in->sites == nullptr) != (in->site_off == nullptr)) return AT_E_ARG;
	if (in->n_pairs >= (1ull << 31)) return AT_E_ARG;
	if (in->site_off) {      // per-pair slices of `sites`: ascending offsets (build_jump_mask indexes with their differences)
		for (uint64_t k = 0; k < in->n_pairs; ++k)
			if (in->site_off[k + 1] < in->site_off[k] || in->site_off[k + 1] - in->site_off[0] > (1ull << 40)) { set_err(h, "site_off is not ascending at pair %llu", (unsigned long long)k); return AT_E_ARG; }
	}
	int64_t maxabs = std::max<int64_t>({llabs((long long)p->m), llabs((long long)p->u), llabs((long long)p->o), llabs((long long)p->e), llabs((long long)p->j), 1});
	// Score lanes hold values x8.  With B = 8 (l1 + l2 + 2) maxabs bounding every finite value, a -inf stand-in
	// (AT_NEG = -2^29, at_kernels.cuh) that has drifted by up to B must still lose against every finite value:
	// AT_NEG + B < -B  <=>  (l1 + l2 + 2) * maxabs < 2^25.  (AT_NEG_INIT - B stays above INT32_MIN with room to spare.)
	const uint64_t max_sum = (uint64_t)(((1ll << 25) - 1) / maxabs);
	uint64_t *pre = nullptr;
	if (prefix) { prefix->resize(in->n_pairs + 1); pre = prefix->data(); pre[0] = 0; }
	// one pair's checks, in the reference's order of failure; `report` false = only say whether it fails
	auto check = [&](uint64_t k, uint64_t l1, uint64_t l2, bool report) -> int {
		if (l1 == 0 || l2 == 0) { if (report) set_err(h, "pair %llu: empty record", (unsigned long long)k); return AT_E_UNDEF; }
		if (mode == AT_FIT && l1 > l2) { if (report) set_err(h, "pair %llu: first sequence must be shorter than the second to do fitting alignment", (unsigned long long)k); return AT_E_FITLEN; }
		if (mode == AT_FIT && l2 < 2) { if (report) set_err(h, "pair %llu: fit with l2 < 2 is undefined in the reference", (unsigned long long)k); return AT_E_UNDEF; }
		/* + 40: K1's lanes run up to 34 columns past l2 (their cells obey the same bounds) */
		if (l1 + l2 + 42 > max_sum) { if (report) set_err(h, "pair %llu: score range", (unsigned long long)k); return AT_E_RANGE; }
		return AT_OK;
	};
	// ranges of pairs: checks and the range's cell total in one pass (a few host threads on a large batch),
	// the running cell counts in a second pass once every range knows its base
	const uint64_t n = in->n_pairs;
	const size_t parts = n >= (1u << 17) ? 4 : 1;
	std::vector<uint64_t> lo(parts + 1), sum(parts, 0), bad(parts, UINT64_MAX), tasks(parts, 0);
	const bool all_wave = mode >= AT_OVERLAP;
	for (size_t r = 0; r <= parts; ++r) lo[r] = n * r / parts;
	auto pass1 = [&](size_t r) {
		uint64_t acc = 0, tk = 0;
		for (uint64_t k = lo[r]; k < lo[r + 1]; ++k) {
			const uint64_t l1 = in->q_len[k], l2 = in->t_len[k];
			if (check(k, l1, l2, false)) { bad[r] = k; break; }
			acc += l1 * l2;
			if (all_wave || l1 > 32u * MAXR) tk += (l1 + 32u * MAXR - 1) / (32u * MAXR);
		}
		sum[r] = acc; tasks[r] = tk;
	};
	auto pass2 = [&](size_t r, uint64_t acc) {
		for (uint64_t k = lo[r]; k < lo[r + 1]; ++k) { acc += (uint64_t)in->q_len[k] * in->t_len[k]; pre[k + 1] = acc; }
	};
	auto each_range = [&](const std::function<void(size_t)> &fn) {
		if (parts == 1) { fn(0); return; }
		std::vector<std::thread> th;
		for (size_t r = 1; r < parts; ++r) th.emplace_back(fn, r);
		fn(0);
		for (auto &t : th) t.join();
	};
	each_range(pass1);
	for (size_t r = 0; r < parts; ++r)
		if (bad[r] != UINT64_MAX) return check(bad[r], in->q_len[bad[r]], in->t_len[bad[r]], true);   // the first failing pair, as a serial scan would report
	if (wave_tasks) { *wave_tasks = 0; for (size_t r = 0; r < parts; ++r) *wave_tasks += tasks[r]; }
	if (pre) {
		std::vector<uint64_t> base(parts, 0);
		for (size_t r = 1; r < parts; ++r) base[r] = base[r - 1] + sum[r - 1];
		each_range([&](size_t r) { pass2(r, base[r]); });
	}
	return AT_OK;
}

// the same slicing rule as cut_by_cells, from running cell counts (binary searches instead of passes over the pairs)
static void cut_by_prefix(const std::vector<uint64_t> &prefix, uint64_t lo, uint64_t hi, size_t parts, std::vector<uint64_t> &cut)
{
	cut.assign(parts + 1, lo);
	cut[parts] = hi;
	if (parts <= 1) return;
	const uint64_t base = prefix[lo], total = prefix[hi] - base;
	for (size_t d = 1; d < parts; ++d) {
		// smallest k in [lo, hi) with (prefix[k+1] - base) * parts >= total * d; the slice ends after it
		uint64_t a = lo, b = hi;
		while (a < b) {
			const uint64_t mid = a + (b - a) / 2;
			if ((unsigned __int128)(prefix[mid + 1] - base) * parts >= (unsigned __int128)total * d) b = mid; else a = mid + 1;
		}
		cut[d] = a < hi ? a + 1 : hi;
	}
	for (size_t d = 1; d <= parts; ++d) cut[d] = std::max(cut[d], cut[d - 1]);
}

// contiguous slices of [lo, hi) with (nearly) equal numbers of cells; cut has parts + 1 entries
static void cut_by_cells(const at_batch_input *in, uint64_t lo, uint64_t hi, size_t parts, std::vector<uint64_t> &cut)
{
	cut.assign(parts + 1, lo);
	cut[parts] = hi;
	if (parts <= 1) return;
	uint64_t total = 0;                       // sum of l1*l2 <= 2^31 pairs x 2^54 would overflow; validate_batch bounds l1+l2 < 2^27
	for (uint64_t k = lo; k < hi; ++k) total += (uint64_t)in->q_len[k] * in->t_len[k];
	uint64_t acc = 0; size_t d = 1;
	for (uint64_t k = lo; k < hi && d < parts; ++k) {
		acc += (uint64_t)in->q_len[k] * in->t_len[k];
		while (d < parts && (unsigned __int128)acc * parts >= (unsigned __int128)total * d) { cut[d++] = k + 1; }
	}
	for (; d < parts; ++d) cut[d] = hi;
	cut[parts] = hi;
}

extern "C" int at_plan_slices(const uint32_t *q_len, const uint32_t *t_len, uint64_t n_pairs, uint32_t parts, uint64_t *cut)
{
	if (!q_len || !t_len || !cut || parts == 0) return AT_E_ARG;
	at_batch_input in;
	memset(&in, 0, sizeof in);
	in.n_pairs = n_pairs; in.q_len = q_len; in.t_len = t_len;
	std::vector<uint64_t> c;
	cut_by_cells(&in, 0, n_pairs, parts, c);
	for (uint32_t k = 0; k <= parts; ++k) cut[k] = c[k];
	return AT_OK;
}

static at_batch *new_batch(at_handle *h, int mode, const at_params *p, uint32_t out_flags, uint64_t n)
{
	at_batch *b = new at_batch();
	b->h = h; b->mode = mode; b->prm = *p; b->out_flags = out_flags; b->n = n;
	b->traceback = mode != AT_EDIT && (out_flags & (AT_OUT_CIGAR | AT_OUT_ALN));
	return b;
}

extern "C" int at_batch_create(at_handle *h, int mode, const at_params *p, const at_batch_input *in,
                               uint32_t out_flags, at_batch **out)
{
	if (!h || !p || !in || !out) return AT_E_ARG;
	*out = nullptr;
	if (int rc = validate_batch(h, mode, p, in)) return rc;
	at_batch *b = new_batch(h, mode, p, out_flags, in->n_pairs);
	// contiguous slices balanced by cells
	const size_t nd = h->devs.size();
	std::vector<uint64_t> cut;
	cut_by_cells(in, 0, b->n, nd, cut);
	b->shards.resize(nd);
	std::vector<std::thread> th;
	for (size_t d = 0; d < nd; ++d) {
		Shard &s = b->shards[d];
		s.dev = &h->devs[d]; s.stream = s.dev->stream; s.p0 = cut[d]; s.p1 = cut[d + 1]; s.out_base = s.p0; s.n = (uint32_t)(s.p1 - s.p0);
	}
	auto work = [&](size_t d) { Shard &s = b->shards[d]; s.rc = s.n ? setup_shard(b, s, in) : AT_OK; };
	if (nd == 1) work(0);
	else { for (size_t d = 0; d < nd; ++d) th.emplace_back(work, d); for (auto &t : th) t.join(); }
	for (auto &s : b->shards) if (s.rc) { int rc = s.rc; at_batch_free(b); return rc; }
	*out = b;
	return AT_OK;
}

// ------------------------------------------------------------------ launch ----
// Kernel tables.  K1 (at_fill_affine): mode variant x rows-per-lane x lanes; kind: 0 global, 1 local,
// 2 fit, 3 fit+jump on int32 lanes, 4 / 5 / 6 local / global / fit on packed s16x2 lanes.  K2 (at_wave_*): affine modes
// with R = 8, single-plane modes (overlap / edit) with R = 1..8.
typedef void (*fill2_fn)(const FillArgs2);
typedef void (*wave_fn)(const WaveArgs);

template <int R, bool PROF> static fill2_fn affine_fn(int kind)
{
	switch (kind) {
	case 0: return at_fill_affine<MODE_GLOBAL, R, false, false, PROF>;
	case 1: return at_fill_affine<MODE_LOCAL, R, false, false, PROF>;
	case 2: return at_fill_affine<MODE_FIT, R, false, false, PROF>;
	case 3: return at_fill_affine<MODE_FIT, R, true, false, PROF>;
	case 4: return at_fill_affine<MODE_LOCAL, R, false, true, PROF>;
	case 5: return at_fill_affine<MODE_GLOBAL, R, false, true, PROF>;
	default: return at_fill_affine<MODE_FIT, R, false, true, PROF>;
	}
}
template <bool PROF> static fill2_fn affine_kernel_r(int kind, int R)
{
	switch (R) {
	case 1: return affine_fn<1, PROF>(kind); case 2: return affine_fn<2, PROF>(kind); case 3: return affine_fn<3, PROF>(kind);
	case 4: return affine_fn<4, PROF>(kind); case 5: return affine_fn<5, PROF>(kind); case 6: return affine_fn<6, PROF>(kind);
	case 7: return affine_fn<7, PROF>(kind); default: return affine_fn<8, PROF>(kind);
	}
}
static fill2_fn affine_kernel(int kind, int R, bool prof) { return prof ? affine_kernel_r<true>(kind, R) : affine_kernel_r<false>(kind, R); }
template <int MODE, bool PROF> static wave_fn wave_linear_fn(int R)
{
	switch (R) {
	case 1: return at_wave_linear<MODE, 1, PROF>; case 2: return at_wave_linear<MODE, 2, PROF>; case 3: return at_wave_linear<MODE, 3, PROF>;
	case 4: return at_wave_linear<MODE, 4, PROF>; case 5: return at_wave_linear<MODE, 5, PROF>; case 6: return at_wave_linear<MODE, 6, PROF>;
	case 7: return at_wave_linear<MODE, 7, PROF>; default: return at_wave_linear<MODE, 8, PROF>;
	}
}
static wave_fn bits_kernel(int R) { return R == 1 ? at_wave_edit_bits<1> : (R == 2 ? at_wave_edit_bits<2> : at_wave_edit_bits<4>); }
static wave_fn wave_kernel(int mode, bool jump, int R, bool prof)
{
	switch (mode) {
	case AT_GLOBAL: return prof ? at_wave_affine<MODE_GLOBAL, false, true> : at_wave_affine<MODE_GLOBAL, false, false>;
	case AT_LOCAL: return prof ? at_wave_affine<MODE_LOCAL, false, true> : at_wave_affine<MODE_LOCAL, false, false>;
	case AT_FIT: return jump ? (prof ? at_wave_affine<MODE_FIT, true, true> : at_wave_affine<MODE_FIT, true, false>)
	                         : (prof ? at_wave_affine<MODE_FIT, false, true> : at_wave_affine<MODE_FIT, false, false>);
	case AT_OVERLAP: return prof ? wave_linear_fn<MODE_OVERLAP, true>(R) : wave_linear_fn<MODE_OVERLAP, false>(R);
	default: return prof ? wave_linear_fn<MODE_EDIT, true>(R) : wave_linear_fn<MODE_EDIT, false>(R);
	}
}

struct CastU64 { __host__ __device__ uint64_t operator()(const uint32_t &x) const { return (uint64_t)x; } };

// `fills_done` (optional) is called once, when the next sub-slice of the pipelined one-shot path may
// launch its fill: after this shard's last traceback kernel has been enqueued (score-only: after its
// fills completed), so the device runs fill, traceback, next fill back to back while this shard's
// D2H copies and host work overlap the next fill.
static int run_shard(at_batch *b, Shard &s, const std::function<void()> *fills_done = nullptr)
{
	at_handle *h = b->h;
	CU(h, cudaSetDevice(s.dev->id));
	cudaStream_t st = s.stream;
	tl_stream = st;
	tl_cache = &s.cache; tl_big = s.dev->big.get();
	const bool jump = b->mode == AT_FIT && b->prm.jump;
	static const bool tb_overlap = getenv("AT_PIPE_TB_OVERLAP") != nullptr;
	s.fill_ms = s.tb_ms = s.dev_ms = s.domk_ms = 0; s.domk_cells = 0; s.launches = 0;
	// dominant (most cells) fill launch of the whole shard -> per-launch timing for the roofline
	int dom_chunk = -1, dom_launch = -1; uint64_t dom_cells = 0;
	for (size_t ci = 0; ci < s.chunks.size(); ++ci)
		for (size_t li = 0; li < s.chunks[ci].launches.size(); ++li)
			if (s.chunks[ci].launches[li].cells > dom_cells) { dom_cells = s.chunks[ci].launches[li].cells; dom_chunk = (int)ci; dom_launch = (int)li; }

	cudaEvent_t e_begin = s.ev[0], e_fill = s.ev[1], e_tb = s.ev[2];
	float ms = 0;
	bool first = true;
	cudaEvent_t e_first = s.ev[3];
	for (size_t ci = 0; ci < s.chunks.size(); ++ci) {
		Chunk &c = s.chunks[ci];
		const uint32_t nc = c.k1 - c.k0;
		CU(h, cudaMemsetAsync(s.d_counter.p, 0, 64 * sizeof(uint32_t), st));
		if (c.prog_words) CU(h, cudaMemsetAsync(s.d_prog.p, 0, c.prog_words * sizeof(uint32_t), st));
		if (b->mode >= AT_OVERLAP && c.bnd_elems) CU(h, cudaMemsetAsync(s.d_bnd.p, 0, c.bnd_elems * sizeof(uint64_t), st));   // tagged hand-off words of the single-plane kernels
		CU(h, cudaEventRecord(e_begin, st));
		if (first) { CU(h, cudaEventRecord(e_first, st)); first = false; }
		for (size_t li = 0; li < c.launches.size(); ++li) {
			Launch &l = c.launches[li];
			const void *fn = l.kind == LK_BITS ? (const void *)bits_kernel(l.r) : l.kind == LK_WAVE ? (const void *)wave_kernel(b->mode, jump, l.r, s.prof)
			                                   : (const void *)affine_kernel(l.kind == LK_PACKED ? (b->mode == AT_LOCAL ? 4 : b->mode == AT_GLOBAL ? 5 : 6) : (b->mode == AT_FIT ? (jump ? 3 : 2) : b->mode), l.r, s.prof);
			const int warps = l.kind >= LK_WAVE ? AT_WAVE_WARPS : AT_FILL_WARPS;
			const size_t dyn_smem = l.kind >= LK_WAVE ? 0 : fill_smem_bytes(l.r, l.kind == LK_PACKED, s.prof);
			if (dyn_smem > 48 * 1024) CU(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
			int occ = 0;
			CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 32 * warps, dyn_smem));
			if (occ < 1) { set_err(h, "fill kernel (mode %d, R %d) cannot be resident: not an sm_100 build?", b->mode, l.r); return AT_E_CUDA; }
			// AT_PIPE_TB_OVERLAP (older scheme, kept for A/B runs): traceback kernels run beside the next sub-slice's
			// fill, which must then leave them shared memory and registers (a K1 grid at full occupancy fills an SM)
			if (tb_overlap && s.workspace && l.kind != LK_WAVE && l.kind != LK_BITS && occ > 4) occ = 4;
			int blocks = (int)std::min<uint64_t>((uint64_t)s.dev->sm_count * occ, (l.n_jobs() + warps - 1) / warps);
			if (blocks < 1) blocks = 1;
			const bool dom = (int)ci == dom_chunk && (int)li == dom_launch;
			if (dom) CU(h, cudaEventRecord(s.evk[0], st));
			if (l.kind >= LK_WAVE) {
				WaveArgs wa;
				wa.symmap = s.d_symmap.p; wa.syms = s.syms;
				wa.q = s.d_q.p; wa.q_off = s.d_q_off.p; wa.q_len = s.d_q_len.p;
				wa.t = s.d_t.p; wa.t_off = s.d_t_off.p; wa.t_len = s.d_t_len.p;
				wa.jmask = s.d_jmask.p; wa.j_off = s.twobit && jump ? s.d_j_off.p : s.d_t_off.p; wa.twobit = s.twobit ? 1 : 0; wa.start_lag = wave_start_lag(); wa.tasks = l.d_tasks.p; wa.n_tasks = (uint32_t)l.h_tasks.size();
				wa.counter = s.d_counter.p + (l.kind == LK_BITS ? 48 : 32) + l.r; wa.prog = s.d_prog.p + l.prog_base;
				wa.ptr = s.d_ptr.p; wa.ptr_off = c.d_ptr_off.p; wa.pair_base = c.k0;
				wa.bnd = s.d_bnd.p; wa.bnd_off = c.d_bnd_off.p; wa.chain = s.d_chain.p;
				wa.score = s.d_score.p; wa.end_i = s.d_end_i.p; wa.end_j = s.d_end_j.p; wa.end_state = s.d_end_state.p;
				wa.m = b->prm.m; wa.u = b->prm.u; wa.o = b->prm.o; wa.e = b->prm.e; wa.jp = b->prm.j;
				wa.want_ptr = b->traceback ? 1 : 0;
				wa.k_and = cell_k_and<false>(); wa.k_or = cell_k_or<false>();
				void *kargs[] = {(void *)&wa};
				CU(h, cudaLaunchKernel(fn, dim3(blocks), dim3(32 * warps), kargs, 0, st));
			} else {
				FillArgs2 fa;
				fa.q = s.d_q.p; fa.q_off = s.d_q_off.p; fa.q_len = s.d_q_len.p;
				fa.t = s.d_t.p; fa.t_off = s.d_t_off.p; fa.t_len = s.d_t_len.p;
				fa.jmask = s.d_jmask.p; fa.j_off = s.twobit && jump ? s.d_j_off.p : s.d_t_off.p; fa.symmap = s.d_symmap.p; fa.syms = s.syms;
				fa.jobs = l.d_jobs.p; fa.n_jobs = (uint32_t)l.n_jobs();
				fa.counter = s.d_counter.p + (l.kind == LK_PACKED ? 16 : 0) + l.r; fa.ptr = s.d_ptr.p; fa.ptr_off = c.d_ptr_off.p; fa.pair_base = c.k0;
				fa.score = s.d_score.p; fa.end_i = s.d_end_i.p; fa.end_j = s.d_end_j.p; fa.end_state = s.d_end_state.p;
				fa.m = b->prm.m; fa.u = b->prm.u; fa.o = b->prm.o; fa.e = b->prm.e; fa.jp = b->prm.j;
				fa.want_ptr = b->traceback ? 1 : 0; fa.twobit = s.twobit ? 1 : 0;
				fa.k_and = l.kind == LK_PACKED ? cell_k_and<true>() : cell_k_and<false>(); fa.k_or = l.kind == LK_PACKED ? cell_k_or<true>() : cell_k_or<false>();
				void *kargs[] = {(void *)&fa};
				CU(h, cudaLaunchKernel(fn, dim3(blocks), dim3(32 * warps), kargs, dyn_smem, st));
			}
			if (dom) CU(h, cudaEventRecord(s.evk[1], st));
			s.launches++;
		}
		CU(h, cudaEventRecord(e_fill, st));
		const bool last_chunk = ci + 1 == s.chunks.size();
		// pipelined one-shot path: the SMs go to the next sub-slice once this one's traceback kernels are in the
		// queue ahead of its fill -- the fill runs at full occupancy and the (short) traceback is not squeezed
		// beside it; copies and host round trips still overlap the neighbours' fills
		const bool tb_first = fills_done && last_chunk && b->traceback && !tb_overlap;
		if (fills_done && last_chunk && !tb_first) { CU(h, cudaEventSynchronize(e_fill)); (*fills_done)(); }
		if (b->traceback) {
			TraceArgs ta;
			ta.q = s.d_q.p; ta.q_off = s.d_q_off.p; ta.q_len = s.d_q_len.p;
			ta.t = s.d_t.p; ta.t_off = s.d_t_off.p; ta.t_len = s.d_t_len.p;
			ta.ptr = s.d_ptr.p; ta.ptr_off = c.d_ptr_off.p; ta.rclass = s.d_rclass.p;
			ta.pair_base = c.k0; ta.n_pairs = nc;
			ta.end_i = s.d_end_i.p; ta.end_j = s.d_end_j.p; ta.end_state = s.d_end_state.p;
			ta.beg_i = s.d_beg_i.p; ta.beg_j = s.d_beg_j.p; ta.n_ops = s.d_n_ops.p; ta.n_cols = s.d_n_cols.p;
			ta.ops_off = c.d_ops_off.p; ta.cols_off = c.d_cols_off.p;
			ta.scratch = s.d_scratch.p; ta.scratch_off = c.d_scratch_off.p;
			ta.cigar = nullptr; ta.aln1 = nullptr; ta.aln2 = nullptr;
			ta.mode = b->mode; ta.jump = jump ? 1 : 0; ta.twobit = s.twobit ? 1 : 0;
			// few long walks: one walker per warp, prefetching ahead; with many walks, or short ones (a few hundred
			// steps: nothing to prefetch far ahead of), the chase is throughput-bound and the prefetches only add traffic
			ta.lookahead = nc < 32768u && c.scratch_words / nc >= 2048 ? 1 : 0;
			const int walk_blocks = ta.lookahead ? (int)(((uint64_t)nc * 32 + 127) / 128) : (int)((nc + 127) / 128);   // one walker per warp | per thread
			if (s.workspace && tb_overlap) at_traceback_walk<true><<<walk_blocks, 128, 0, st>>>(ta);
			else at_traceback_walk<false><<<walk_blocks, 128, 0, st>>>(ta);
			CU(h, cudaGetLastError());
			s.launches++;
			// exclusive offsets of n_ops / n_cols: offsets[0] = 0, offsets[1..nc] inclusive sums
			if (s.workspace && nc <= (1u << 18)) {      // pipelined path: must be able to run beside another stream's fill
				at_scan_offsets<<<2, 256, 0, st>>>(s.d_n_ops.p + c.k0, s.d_n_cols.p + c.k0, nc, c.d_ops_off.p, c.d_cols_off.p);
				CU(h, cudaGetLastError());
				s.launches++;
			} else {
				size_t tmp_bytes = 0;
				auto it_ops = thrust::make_transform_iterator((const uint32_t *)(s.d_n_ops.p + c.k0), CastU64());
				auto it_cols = thrust::make_transform_iterator((const uint32_t *)(s.d_n_cols.p + c.k0), CastU64());
				CU(h, cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, it_ops, c.d_ops_off.p + 1, (int)nc, st));
				CU(h, s.d_scan_tmp.alloc(tmp_bytes + 16));
				CU(h, cudaMemsetAsync(c.d_ops_off.p, 0, sizeof(uint64_t), st));
				CU(h, cudaMemsetAsync(c.d_cols_off.p, 0, sizeof(uint64_t), st));
				CU(h, cub::DeviceScan::InclusiveSum(s.d_scan_tmp.p, tmp_bytes, it_ops, c.d_ops_off.p + 1, (int)nc, st));
				CU(h, cub::DeviceScan::InclusiveSum(s.d_scan_tmp.p, tmp_bytes, it_cols, c.d_cols_off.p + 1, (int)nc, st));
			}
			const bool want_cig = b->out_flags & AT_OUT_CIGAR, want_aln = b->out_flags & AT_OUT_ALN;
			auto out_buffers = [&](uint64_t ops, uint64_t cols) -> int {
				// tb_first: the workspace's own buffers, kept across sub-slices and calls (grow-only) -- upper-bound sized
				// blocks are too large to go through the allocator in the steady state
				DevBuf<uint32_t> &cg = tb_first ? s.ws_cigar : c.d_cigar;
				DevBuf<uint8_t> &a1 = tb_first ? s.ws_aln1 : c.d_aln1, &a2 = tb_first ? s.ws_aln2 : c.d_aln2;
				if (want_cig) { if (cg.alloc(ops + 1) != cudaSuccess) { set_err(h, "cigar buffer"); return AT_E_NOMEM; } }
				if (want_aln) { if (a1.alloc(cols + 1) != cudaSuccess || a2.alloc(cols + 1) != cudaSuccess) { set_err(h, "alignment buffer"); return AT_E_NOMEM; } }
				ta.cigar = c.cigar = want_cig ? cg.p : nullptr;
				ta.aln1 = c.aln1 = want_aln ? a1.p : nullptr; ta.aln2 = c.aln2 = want_aln ? a2.p : nullptr;
				return AT_OK;
			};
			uint64_t tot[2] = {0, 0};
			if (tb_first) {
				// no host round trip between the walk and the emit: the outputs are sized by their upper bound
				// (an alignment has at most l1 + l2 columns; the workspace keeps the buffers), the totals follow later
				// (rounded up generously: sub-slices differ by a few pairs, and a request just above every cached
				// block would go to the CUDA allocator, which waits for the fills already in the queue)
				const uint64_t cap = (c.scratch_words + c.scratch_words / 8 + (1u << 22)) & ~(uint64_t)((1u << 22) - 1);
				if (int rc = out_buffers(cap, cap)) return rc;
			} else {
				CU(h, cudaMemcpyAsync(&tot[0], c.d_ops_off.p + nc, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
				CU(h, cudaMemcpyAsync(&tot[1], c.d_cols_off.p + nc, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
				CU(h, shard_sync(s));
				if (int rc = out_buffers(tot[0], tot[1])) return rc;
			}
			if (s.workspace && tb_overlap) at_traceback_emit<true><<<(int)(((uint64_t)nc * 32 + 127) / 128), 128, 0, st>>>(ta);
			else at_traceback_emit<false><<<(int)(((uint64_t)nc * 32 + 127) / 128), 128, 0, st>>>(ta);
			CU(h, cudaGetLastError());
			s.launches++;
			if (tb_first) {
				// the next fill is launched once this one has COMPLETED: launched earlier it would take the SMs
				// as this fill's blocks retire, and this sub-slice's traceback (and its workspace) would wait a whole
				// fill behind it -- the workers then fall into lockstep
				CU(h, cudaEventSynchronize(e_fill));
				(*fills_done)();
				CU(h, cudaMemcpyAsync(&tot[0], c.d_ops_off.p + nc, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
				CU(h, cudaMemcpyAsync(&tot[1], c.d_cols_off.p + nc, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
				CU(h, shard_sync(s));
			}
			c.tot_ops = tot[0]; c.tot_cols = tot[1];
		}
		CU(h, cudaEventRecord(e_tb, st));
		CU(h, shard_sync(s));
		CU(h, cudaEventElapsedTime(&ms, e_begin, e_fill)); s.fill_ms += ms;
		CU(h, cudaEventElapsedTime(&ms, e_fill, e_tb)); s.tb_ms += ms;
		if ((int)ci == dom_chunk) {
			CU(h, cudaEventElapsedTime(&ms, s.evk[0], s.evk[1])); s.domk_ms = ms; s.domk_cells = dom_cells;
			const Launch &dl = c.launches[dom_launch];
			s.domk_kind = (uint32_t)dl.kind; s.domk_r = (uint32_t)dl.r; s.domk_flags = (s.prof ? 1u : 0u) | (jump ? 2u : 0u) | (s.twobit ? 4u : 0u);
		}
		if (ci + 1 == s.chunks.size()) { CU(h, cudaEventElapsedTime(&ms, e_first, e_tb)); s.dev_ms = ms; }
	}
	h->launches += s.launches;
	return AT_OK;
}

extern "C" int at_batch_run(at_batch *b, at_timing *timing)
{
	if (!b) return AT_E_ARG;
	const size_t nd = b->shards.size();
	auto work = [&](size_t d) { Shard &s = b->shards[d]; s.rc = s.n ? run_shard(b, s) : AT_OK; };
	if (nd == 1) work(0);
	else { std::vector<std::thread> th; for (size_t d = 0; d < nd; ++d) th.emplace_back(work, d); for (auto &t : th) t.join(); }
	for (auto &s : b->shards) if (s.rc) return s.rc;
	b->ran = true;
	if (timing) {
		memset(timing, 0, sizeof *timing);
		for (auto &s : b->shards) {
			timing->fill_ms = std::max(timing->fill_ms, s.fill_ms);
			timing->traceback_ms = std::max(timing->traceback_ms, s.tb_ms);
			timing->device_ms = std::max(timing->device_ms, s.dev_ms);
			timing->cells += s.cells; timing->launches += s.launches; timing->ptr_bytes += b->traceback ? s.ptr_bytes : 0;
			if (s.domk_cells > timing->fill_kernel_cells) { timing->fill_kernel_cells = s.domk_cells; timing->fill_kernel_ms = s.domk_ms; timing->fill_kernel_kind = s.domk_kind; timing->fill_kernel_rows = s.domk_r; timing->fill_kernel_flags = s.domk_flags; }
		}
	}
	return AT_OK;
}

extern "C" int at_batch_sizes(const at_batch *b, uint64_t *cigar_ops, uint64_t *aln_bytes)
{
	if (!b || !b->ran) return AT_E_ARG;
	uint64_t o = 0, c = 0;
	for (auto &s : b->shards) for (auto &ch : s.chunks) { o += ch.tot_ops; c += ch.tot_cols; }
	if (cigar_ops) *cigar_ops = (b->out_flags & AT_OUT_CIGAR) ? o : 0;
	if (aln_bytes) *aln_bytes = (b->out_flags & AT_OUT_ALN) ? c : 0;
	return AT_OK;
}

// D2H of one shard's results into the caller's arrays.  base_ops / base_cols: position of the
// shard's first op / column in the dense outputs (advanced past this shard on return).  The
// offset arrays receive entries [out_base, out_base + n) here; the final entry [n_total] is the
// caller's job.
static int fetch_shard(at_batch *b, Shard &s, at_batch_output *out, bool want_cig, bool want_aln,
                       uint64_t &base_ops, uint64_t &base_cols)
{
	at_handle *h = b->h;
	if (!s.n) return AT_OK;
	CU(h, cudaSetDevice(s.dev->id));
	cudaStream_t st = s.stream;
	const uint64_t ob = s.out_base;
	CU(h, cudaMemcpyAsync(out->score + ob, s.d_score.p, s.n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
	if (out->end_i) CU(h, cudaMemcpyAsync(out->end_i + ob, s.d_end_i.p, s.n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	if (out->end_j) CU(h, cudaMemcpyAsync(out->end_j + ob, s.d_end_j.p, s.n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	if (out->beg_i) CU(h, cudaMemcpyAsync(out->beg_i + ob, s.d_beg_i.p, s.n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	if (out->beg_j) CU(h, cudaMemcpyAsync(out->beg_j + ob, s.d_beg_j.p, s.n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	for (auto &c : s.chunks) {
		const uint32_t nc = c.k1 - c.k0;
		if (want_cig) {
			CU(h, cudaMemcpyAsync(out->cigar_off + ob + c.k0, c.d_ops_off.p, nc * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
			if (c.tot_ops) CU(h, cudaMemcpyAsync(out->cigar + base_ops, c.cigar, c.tot_ops * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
		}
		if (want_aln) {
			CU(h, cudaMemcpyAsync(out->aln_off + ob + c.k0, c.d_cols_off.p, nc * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
			if (c.tot_cols) {
				CU(h, cudaMemcpyAsync(out->aln1 + base_cols, c.aln1, c.tot_cols, cudaMemcpyDeviceToHost, st));
				CU(h, cudaMemcpyAsync(out->aln2 + base_cols, c.aln2, c.tot_cols, cudaMemcpyDeviceToHost, st));
			}
		}
		CU(h, shard_sync(s));
		// chunk-local offsets -> global
		if (want_cig && base_ops) for (uint32_t k = 0; k < nc; ++k) out->cigar_off[ob + c.k0 + k] += base_ops;
		if (want_aln && base_cols) for (uint32_t k = 0; k < nc; ++k) out->aln_off[ob + c.k0 + k] += base_cols;
		base_ops += c.tot_ops; base_cols += c.tot_cols;
	}
	CU(h, shard_sync(s));
	return AT_OK;
}

extern "C" int at_batch_fetch(at_batch *b, at_batch_output *out)
{
	if (!b || !out || !b->ran || !out->score) return AT_E_ARG;
	at_handle *h = b->h;
	const bool want_cig = b->traceback && (b->out_flags & AT_OUT_CIGAR) && out->cigar;
	const bool want_aln = b->traceback && (b->out_flags & AT_OUT_ALN) && out->aln1 && out->aln2;
	if (want_cig && !out->cigar_off) return AT_E_ARG;
	if (want_aln && !out->aln_off) return AT_E_ARG;
	uint64_t tot_ops = 0, tot_cols = 0;
	at_batch_sizes(b, &tot_ops, &tot_cols);
	if (want_cig && tot_ops > out->cigar_cap) { set_err(h, "cigar buffer too small: need %llu ops", (unsigned long long)tot_ops); return AT_E_NOSPACE; }
	if (want_aln && tot_cols > out->aln_cap) { set_err(h, "alignment buffer too small: need %llu bytes", (unsigned long long)tot_cols); return AT_E_NOSPACE; }
	uint64_t base_ops = 0, base_cols = 0;
	for (auto &s : b->shards)
		if (int rc = fetch_shard(b, s, out, want_cig, want_aln, base_ops, base_cols)) return rc;
	if (want_cig) out->cigar_off[b->n] = base_ops;
	if (want_aln) out->aln_off[b->n] = base_cols;
	if (!b->traceback) {
		if (out->cigar_off) for (uint64_t k = 0; k <= b->n; ++k) out->cigar_off[k] = 0;
		if (out->aln_off) for (uint64_t k = 0; k <= b->n; ++k) out->aln_off[k] = 0;
	}
	return AT_OK;
}

#include "at_pipeline.inl"      // at_batch_align: the pipelined one-shot path (same translation unit)

extern "C" void *at_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return p;
}

extern "C" void at_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" int at_host_register(void *p, size_t bytes)
{
	if (!p || !bytes) return AT_E_ARG;
	if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return AT_E_CUDA; }
	return AT_OK;
}

extern "C" int at_host_unregister(void *p)
{
	if (!p) return AT_E_ARG;
	if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return AT_E_CUDA; }
	return AT_OK;
}

extern "C" int64_t at_pack_2bit(const char *seq, uint64_t n, uint8_t *dst)
{
	if (!seq || !dst) return -1;
	const uint64_t nb = (n + 3) / 4;
	memset(dst, 0, nb);
	for (uint64_t k = 0; k < n; ++k) {
		uint8_t c;
		switch (seq[k]) { case 'A': c = 0; break; case 'C': c = 1; break; case 'G': c = 2; break; case 'T': c = 3; break; default: return -1; }
		dst[k >> 2] |= (uint8_t)(c << (2 * (k & 3)));
	}
	return (int64_t)nb;
}

extern "C" int64_t at_cigar_to_string(const uint32_t *ops, uint64_t n_ops, char *dst, uint64_t cap)
{
	if (!dst || (!ops && n_ops)) return -1;
	uint64_t pos = 0;
	for (uint64_t k = 0; k < n_ops; ++k) {
		char buf[16];
		const uint32_t code = ops[k] & 15u;      // the format reserves four bits for the op code
		if (code > AT_CIG_N) return -1;
		const int len = snprintf(buf, sizeof buf, "%u%c", ops[k] >> 4, "MIDN"[code]);
		if (pos + len + 1 > cap) return -1;
		memcpy(dst + pos, buf, len); pos += len;
	}
	if (pos + 1 > cap) return -1;
	dst[pos] = 0;
	return (int64_t)pos;
}
