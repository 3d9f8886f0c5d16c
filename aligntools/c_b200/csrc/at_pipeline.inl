// at_pipeline.inl -- the one-shot entry at_batch_align and its pipelined path; part of at_runtime.cu's
// translation unit (included there: it uses the internal Shard / at_batch types and setup_shard /
// run_shard / fetch_shard).
// ------------------------------------------------------------- one-shot, pipelined ----
// at_batch_align on a large batch: every device's slice is cut into sub-slices that go through
// create (H2D) -> run (fill + traceback) -> fetch (D2H) on AT_PIPE_STREAMS streams, one host
// thread per stream, so the copies of one sub-slice overlap the kernels of another.  Sub-slices
// are claimed in pair order; a sub-slice's position in the dense CIGAR / alignment outputs is
// known once every earlier sub-slice has finished its run (totals are published under a mutex).
#define AT_PIPE_STREAMS 3
static uint64_t env_u64(const char *name, uint64_t dflt) { const char *e = getenv(name); return e && *e ? (uint64_t)strtoull(e, nullptr, 10) : dflt; }
static uint64_t pipe_min_cells() { return env_u64("AT_PIPE_MIN_CELLS", 1ull << 31); }      // below this a batch is not worth cutting up
static uint64_t pipe_slice_cells() { return env_u64("AT_PIPE_SLICE_CELLS", 1ull << 33); }  // target cells per full-size sub-slice (about 5 ms of fill)

struct PipeSlice { uint64_t lo = 0, hi = 0; size_t dev = 0; uint64_t tot_ops = 0, tot_cols = 0; bool known = false; };

static int align_pipelined(at_handle *h, int mode, const at_params *p, const at_batch_input *in, uint32_t out_flags,
                           at_batch_output *out, at_timing *timing, const std::vector<uint64_t> &prefix, uint64_t slice_cells)
{
	const uint64_t total_cells = prefix[in->n_pairs];
	const auto t_func = std::chrono::steady_clock::now();
	const size_t nd = h->devs.size();
	if (h->pipe_ws.size() < nd * AT_PIPE_STREAMS) h->pipe_ws.resize(nd * AT_PIPE_STREAMS, nullptr);
	std::vector<uint64_t> dcut;
	cut_by_prefix(prefix, 0, in->n_pairs, nd, dcut);
	std::vector<PipeSlice> slices;
	std::vector<std::vector<size_t>> per_dev(nd);
	for (size_t d = 0; d < nd; ++d) {
		if (dcut[d + 1] == dcut[d]) continue;
		const uint64_t cells = prefix[dcut[d + 1]] - prefix[dcut[d]];
		// graduated sub-slices: quarter-size units, grouped 1, 2, 4, 4, 4, ... so that the first kernel starts
		// early (short first upload) while the bulk runs in full-size sub-slices (fewer kernel tails)
		size_t units = (size_t)std::max<uint64_t>(1, (unsigned __int128)4 * cells / slice_cells);
		units = std::min<size_t>(units, 256);
		units = std::min<size_t>(units, (size_t)(dcut[d + 1] - dcut[d]));
		std::vector<uint64_t> cut;
		cut_by_prefix(prefix, dcut[d], dcut[d + 1], units, cut);
		const size_t max_group = (size_t)std::max<uint64_t>(1, env_u64("AT_PIPE_MAX_GROUP", 4));
		// ... and graduated again at the end (..., 4, 2, 1): what is left to do after the last fill -- the traceback of the
		// last sub-slices and their D2H copies -- shrinks with them (C2: 1.2 ms after the last kernel with a full-size
		// sub-slice second to last)
		std::vector<size_t> tail;
		for (size_t g = max_group / 2; g >= 1; g /= 2) tail.push_back(g);
		size_t tail_units = 0;
		for (size_t g : tail) tail_units += g;
		if (units < 2 * tail_units + max_group || getenv("AT_PIPE_NO_RAMP_DOWN")) { tail.clear(); tail_units = 0; }
		std::vector<size_t> groups;
		for (size_t left = units - tail_units, step = 1; left > 0; step = std::min<size_t>(2 * step, max_group)) {
			const size_t g = std::min(step, left);
			groups.push_back(g); left -= g;
		}
		groups.insert(groups.end(), tail.begin(), tail.end());
		size_t u0 = 0;
		for (size_t g : groups) {
			const size_t u1 = std::min(units, u0 + g);
			if (cut[u1] != cut[u0]) {
				PipeSlice sl; sl.lo = cut[u0]; sl.hi = cut[u1]; sl.dev = d;
				per_dev[d].push_back(slices.size());
				slices.push_back(sl);
			}
			u0 = u1;
		}
	}
	// pipeline streams (created once per device)
	for (size_t d = 0; d < nd; ++d) {
		at_device &dv = h->devs[d];
		if (!dv.pipe[0]) {
			CU(h, cudaSetDevice(dv.id));
			for (int w = 0; w < AT_PIPE_STREAMS; ++w) CU(h, cudaStreamCreateWithFlags(&dv.pipe[w], cudaStreamNonBlocking));
		}
	}
	const bool traceback = mode != AT_EDIT && (out_flags & (AT_OUT_CIGAR | AT_OUT_ALN));
	const bool want_cig = traceback && (out_flags & AT_OUT_CIGAR) && out->cigar;
	const bool want_aln = traceback && (out_flags & AT_OUT_ALN) && out->aln1 && out->aln2;
	if (want_cig && !out->cigar_off) return AT_E_ARG;
	if (want_aln && !out->aln_off) return AT_E_ARG;

	std::mutex mu; std::condition_variable cv;
	std::atomic<int> failed{0};
	// one sub-slice's FILL at a time owns a device's SMs: concurrent persistent fills would share them,
	// finish together and leave the GPU idle while all workers prepare their next sub-slice in lockstep.
	// The lock is handed on when the fill has completed and the sub-slice's traceback kernels are already in
	// the queue behind it: the device runs fill, traceback, next fill back to back (run_shard, `fills_done`).
	std::vector<std::mutex> run_mu(nd);
	std::vector<UploadGate> gate(nd);          // uploads of a device's sub-slices go one at a time, in pair order
	std::vector<std::atomic<size_t>> next(nd);
	for (auto &x : next) x = 0;
	struct Acc { double fill = 0, tb = 0, dev = 0, domk = 0; uint64_t domc = 0, launches = 0, ptr = 0; uint32_t kind = 0, r = 0, flags = 0; };
	std::vector<Acc> acc(nd * AT_PIPE_STREAMS);
	const int n_workers = (int)std::min<uint64_t>(AT_PIPE_STREAMS, std::max<uint64_t>(1, env_u64("AT_PIPE_WORKERS", AT_PIPE_STREAMS)));   // diagnosis: fewer workers

	const bool trace = getenv("AT_PIPE_TRACE") != nullptr;      // host timeline of every sub-slice on stderr
	if (trace) fprintf(stderr, "[at pipe] slicing: %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_func).count());
	const auto t_origin = std::chrono::steady_clock::now();
	auto now_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_origin).count(); };
	auto worker = [&](size_t d, int w) {
		at_device &dv = h->devs[d];
		Acc &a = acc[d * AT_PIPE_STREAMS + w];
		// one workspace per worker, owned by the handle: the sequence buffers, pointer arena and scratch
		// are reused by every sub-slice of this and of later calls (no allocator traffic in the pipeline)
		at_batch *&slot = h->pipe_ws[d * AT_PIPE_STREAMS + w];
		if (!slot) { slot = new_batch(h, mode, p, out_flags, 0); slot->shards.resize(1); }
		at_batch *b = slot;
		b->mode = mode; b->prm = *p; b->out_flags = out_flags;
		b->traceback = mode != AT_EDIT && (out_flags & (AT_OUT_CIGAR | AT_OUT_ALN));
		for (;;) {
			const size_t k = next[d].fetch_add(1);
			if (k >= per_dev[d].size()) break;
			if (failed.load()) { gate[d].pass(k); break; }      // a later sub-slice may already wait for this turn
			const size_t si = per_dev[d][k];
			PipeSlice &sl = slices[si];
			at_batch_input sub = *in;
			sub.n_pairs = sl.hi - sl.lo;
			sub.q_off = in->q_off + sl.lo; sub.q_len = in->q_len + sl.lo;
			sub.t_off = in->t_off + sl.lo; sub.t_len = in->t_len + sl.lo;
			if (in->site_off) sub.site_off = in->site_off + sl.lo;
			b->n = sub.n_pairs;
			Shard &s = b->shards[0];
			s.dev = &dv; s.stream = dv.pipe[w]; s.workspace = true; s.p0 = 0; s.p1 = sub.n_pairs; s.n = (uint32_t)sub.n_pairs; s.out_base = sl.lo;
			s.gate = &gate[d]; s.gate_turn = k;
			const double t_a = now_ms();
			int rc = setup_shard(b, s, &sub);
			gate[d].pass(k);                        // also when the set-up returned before its upload
			s.gate = nullptr;
			const double t_b = now_ms();
			double t_lock = 0, t_hand = 0;
			if (!rc) {
				std::unique_lock<std::mutex> own(run_mu[d]);
				t_lock = now_ms();
				const std::function<void()> hand_on = [&] { if (own.owns_lock()) { own.unlock(); t_hand = now_ms(); } };
				rc = run_shard(b, s, &hand_on);
			}
			const double t_c = now_ms();
			uint64_t to = 0, tc = 0;
			if (!rc) for (auto &c : s.chunks) { to += c.tot_ops; tc += c.tot_cols; }
			uint64_t base_ops = 0, base_cols = 0;
			{
				std::unique_lock<std::mutex> lk(mu);
				sl.tot_ops = to; sl.tot_cols = tc; sl.known = true;
				if (rc) failed = rc;
				cv.notify_all();
				cv.wait(lk, [&] { if (failed.load()) return true; for (size_t x = 0; x < si; ++x) if (!slices[x].known) return false; return true; });
				for (size_t x = 0; x < si; ++x) { base_ops += slices[x].tot_ops; base_cols += slices[x].tot_cols; }
			}
			if (!rc && !failed.load()) {
				if ((want_cig && base_ops + to > out->cigar_cap) || (want_aln && base_cols + tc > out->aln_cap)) {
					set_err(h, "output buffer too small: need at least %llu ops / %llu bytes", (unsigned long long)(base_ops + to), (unsigned long long)(base_cols + tc));
					rc = AT_E_NOSPACE;
				} else rc = fetch_shard(b, s, out, want_cig, want_aln, base_ops, base_cols);
			}
			if (trace) fprintf(stderr, "[at pipe] dev %zu stream %d slice %zu pairs %llu: setup %.2f-%.2f SMs %.2f-%.2f run -%.2f fetch -%.2f ms (fill %.2f tb %.2f)\n",
			                   d, w, si, (unsigned long long)sub.n_pairs, t_a, t_b, t_lock, t_hand, t_c, now_ms(), s.fill_ms, s.tb_ms);
			a.fill += s.fill_ms; a.tb += s.tb_ms; a.dev += s.dev_ms; a.launches += s.launches; a.ptr += traceback ? s.ptr_bytes : 0;
			if (s.domk_cells > a.domc) { a.domc = s.domk_cells; a.domk = s.domk_ms; a.kind = s.domk_kind; a.r = s.domk_r; a.flags = s.domk_flags; }
			if (rc) { std::lock_guard<std::mutex> lk(mu); failed = rc; cv.notify_all(); break; }
		}
	};
	// never more workers than sub-slices: a device's single sub-slice must always land in worker 0's workspace, which
	// holds the arena from the previous call (another worker's would be cold: 100 ms to allocate a 50 GB arena)
	std::vector<std::thread> th;
	for (size_t d = 0; d < nd; ++d)
		for (int w = 0; w < n_workers && (size_t)w < per_dev[d].size(); ++w) th.emplace_back(worker, d, w);
	for (auto &t : th) t.join();
	if (failed.load()) return failed.load();
	uint64_t all_ops = 0, all_cols = 0;
	for (auto &sl : slices) { all_ops += sl.tot_ops; all_cols += sl.tot_cols; }
	if (want_cig) out->cigar_off[in->n_pairs] = all_ops;
	if (want_aln) out->aln_off[in->n_pairs] = all_cols;
	if (!traceback) {
		if (out->cigar_off) for (uint64_t k = 0; k <= in->n_pairs; ++k) out->cigar_off[k] = 0;
		if (out->aln_off) for (uint64_t k = 0; k <= in->n_pairs; ++k) out->aln_off[k] = 0;
	}
	if (timing) {
		memset(timing, 0, sizeof *timing);
		timing->cells = total_cells;
		for (size_t d = 0; d < nd; ++d) {
			double fill = 0, tb = 0, dev = 0;
			for (int w = 0; w < AT_PIPE_STREAMS; ++w) {
				const Acc &a = acc[d * AT_PIPE_STREAMS + w];
				fill += a.fill; tb += a.tb; dev += a.dev; timing->launches += a.launches; timing->ptr_bytes += a.ptr;
				if (a.domc > timing->fill_kernel_cells) { timing->fill_kernel_cells = a.domc; timing->fill_kernel_ms = a.domk; timing->fill_kernel_kind = a.kind; timing->fill_kernel_rows = a.r; timing->fill_kernel_flags = a.flags; }
			}
			timing->fill_ms = std::max(timing->fill_ms, fill); timing->traceback_ms = std::max(timing->traceback_ms, tb);
			timing->device_ms = std::max(timing->device_ms, dev);     // sum of the sub-slices' device times (they overlap)
		}
	}
	return AT_OK;
}

extern "C" int at_batch_align(at_handle *h, int mode, const at_params *p, const at_batch_input *in,
                              uint32_t out_flags, at_batch_output *out, at_timing *timing)
{
	if (!h || !p || !in || !out || !out->score) return AT_E_ARG;
	const auto t_entry = std::chrono::steady_clock::now();
	std::unique_lock<std::mutex> one_at_a_time(h->align_mu);      // the handle's pipeline workspaces serve one call at a time
	std::vector<uint64_t> &prefix = h->pipe_prefix;
	uint64_t wave_tasks = 0;
	if (int rc = validate_batch(h, mode, p, in, &prefix, &wave_tasks)) return rc;
	const uint64_t cells = prefix[in->n_pairs];
	// Sub-slice size: about 2^33 cells (5 ms of fill) for short pairs; where the pairs run as K2 tasks (stripes of 256 rows,
	// one warp each) a sub-slice must still hold several waves of them -- a launch with fewer tasks than resident warps
	// leaves SMs idle and its stripe pipelines never fill.  A batch that gives fewer than two such sub-slices is not cut up.
	uint64_t slice_cells = pipe_slice_cells();
	bool worth = cells >= pipe_min_cells() && in->n_pairs >= 16;
	if (wave_tasks) {
		uint64_t min_tasks = 0;
		for (auto &d : h->devs) min_tasks = std::max<uint64_t>(min_tasks, 4ull * 16 * d.sm_count);      // four waves of 16 warps per SM
		min_tasks = env_u64("AT_PIPE_MIN_TASKS", min_tasks);
		const uint64_t per_dev = wave_tasks / h->devs.size();
		if (per_dev < 2 * min_tasks) worth = false;
		else slice_cells = std::max<uint64_t>(slice_cells, (uint64_t)((double)cells / (double)wave_tasks * (double)min_tasks));
	}
	// A batch not worth cutting up but still sizeable (a few hundred long pairs, say) runs as ONE sub-slice per device through the
	// same machinery: the workers' workspaces keep their pointer arena and buffers from call to call, where a fresh at_batch
	// asks the driver for its arena and for the free-memory figure every time (10-25 ms on C3 / C4-sized calls).
	const bool one_slice = !worth && cells >= env_u64("AT_ONE_SLICE_MIN_CELLS", 1ull << 28);
	if ((worth || one_slice) && !getenv("AT_NO_PIPELINE")) {
		if (getenv("AT_PIPE_TRACE")) fprintf(stderr, "[at pipe] validation + cell count of %llu pairs: %.2f ms\n", (unsigned long long)in->n_pairs,
		                                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count());
		return align_pipelined(h, mode, p, in, out_flags, out, timing, prefix, one_slice ? UINT64_MAX / 8 : slice_cells);
	}
	// small batches: the plain three-call path, still under the handle's lock (include/aligntools_b200.h: a handle
	// serves one batch operation at a time; concurrent callers of at_batch_align are serialised)
	at_batch *b = nullptr;
	int rc = at_batch_create(h, mode, p, in, out_flags, &b);
	if (rc) return rc;
	rc = at_batch_run(b, timing);
	if (!rc) rc = at_batch_fetch(b, out);
	at_batch_free(b);
	return rc;
}

