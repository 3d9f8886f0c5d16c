// at_devmem.h -- device-memory plumbing of the host runtime (at_runtime.cu): typed device buffers, the
// per-shard cache of released blocks and the per-device cache of arena-sized blocks.  Host code only.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <vector>

// Large device blocks (the traceback-pointer arena: tens of GB) come from plain cudaMalloc and are kept
// in a per-device cache for the life of the handle.  The stream-ordered pool maps such a block at about
// 20 GB/s the first time (seconds for one arena); cudaMalloc takes milliseconds.
struct BigCache {
	struct Blk { void *p; size_t bytes; };
	std::mutex mu;
	std::vector<Blk> free_blocks;
	void *take(size_t bytes, size_t *got) {
		std::lock_guard<std::mutex> g(mu);
		size_t best = SIZE_MAX;
		for (size_t k = 0; k < free_blocks.size(); ++k)
			if (free_blocks[k].bytes >= bytes && free_blocks[k].bytes <= 2 * bytes &&
			    (best == SIZE_MAX || free_blocks[k].bytes < free_blocks[best].bytes)) best = k;
		if (best == SIZE_MAX) return nullptr;
		void *p = free_blocks[best].p; *got = free_blocks[best].bytes;
		free_blocks.erase(free_blocks.begin() + best);
		return p;
	}
	void give(void *p, size_t bytes) { std::lock_guard<std::mutex> g(mu); free_blocks.push_back(Blk{p, bytes}); }
	void drop_all() { std::lock_guard<std::mutex> g(mu); for (auto &b : free_blocks) cudaFree(b.p); free_blocks.clear(); }
	size_t cached_bytes() { std::lock_guard<std::mutex> g(mu); size_t t = 0; for (auto &b : free_blocks) t += b.bytes; return t; }
};
static const size_t AT_BIG_BLOCK = 256ull << 20;

// Device buffers come from the device's stream-ordered memory pool (cudaMallocAsync); at_create
// raises the pool's release threshold so that the 40+ GB pointer arena of one batch is handed to
// the next batch without going back to the driver.  tl_stream is the calling shard's stream.
static thread_local cudaStream_t tl_stream = nullptr;

// A shard (in the pipelined one-shot path: a worker's workspace) keeps the blocks it releases in a
// small cache and reuses them for its next allocations, so a steady-state sub-slice makes NO call into
// the CUDA allocator: cudaMallocAsync / cudaFreeAsync from several streams make the pool insert
// cross-stream dependencies (or map new memory), which coupled the pipeline workers' streams.
struct BufCache {
	struct Blk { void *p; size_t bytes; };
	std::vector<Blk> free_blocks;
	void *take(size_t bytes) {
		size_t best = SIZE_MAX;
		for (size_t k = 0; k < free_blocks.size(); ++k)
			if (free_blocks[k].bytes >= bytes && free_blocks[k].bytes <= 4 * bytes + 4096 &&
			    (best == SIZE_MAX || free_blocks[k].bytes < free_blocks[best].bytes)) best = k;
		if (best == SIZE_MAX) return nullptr;
		void *p = free_blocks[best].p;
		last_bytes = free_blocks[best].bytes;
		free_blocks.erase(free_blocks.begin() + best);
		return p;
	}
	size_t last_bytes = 0;
	void give(void *p, size_t bytes) { free_blocks.push_back(Blk{p, bytes}); }
	void flush(cudaStream_t st) { for (auto &b : free_blocks) cudaFreeAsync(b.p, st); free_blocks.clear(); }
};
static thread_local BufCache *tl_cache = nullptr;
static thread_local BigCache *tl_big = nullptr;

template <class T> struct DevBuf {
	T *p = nullptr; size_t n = 0; size_t bytes = 0; bool big = false;
	cudaError_t alloc(size_t count) {
		if (count <= n && p) return cudaSuccess;
		release();
		const size_t want = std::max<size_t>(count, 1) * sizeof(T);
		if (want >= AT_BIG_BLOCK && tl_big) {          // arena-sized: cudaMalloc, cached per device
			size_t got = 0;
			void *q = tl_big->take(want, &got);
			cudaError_t e = cudaSuccess;
			if (!q) {
				got = want;
				e = cudaMalloc(&q, want);
				if (e != cudaSuccess) { cudaGetLastError(); tl_big->drop_all(); e = cudaMalloc(&q, want); }   // cached blocks may be in the way
			}
			if (e != cudaSuccess) { cudaGetLastError(); p = nullptr; return e; }
			p = (T *)q; bytes = got; n = got / sizeof(T); big = true;
			return cudaSuccess;
		}
		if (tl_cache) { if (void *q = tl_cache->take(want)) { p = (T *)q; bytes = tl_cache->last_bytes; n = bytes / sizeof(T); return cudaSuccess; } }
		cudaError_t e = cudaMallocAsync((void **)&p, want, tl_stream);
		if (e != cudaSuccess && tl_big) { cudaGetLastError(); tl_big->drop_all(); e = cudaMallocAsync((void **)&p, want, tl_stream); }   // cached arenas may be in the way
		if (e == cudaSuccess) { n = count; bytes = want; } else { p = nullptr; cudaGetLastError(); }
		return e;
	}
	void release() {
		if (p) {
			if (big && tl_big) tl_big->give(p, bytes);
			else if (big) cudaFree(p);
			else if (tl_cache) tl_cache->give(p, bytes);
			else cudaFreeAsync(p, tl_stream);
		}
		p = nullptr; n = 0; bytes = 0; big = false;
	}
};

