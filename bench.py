#!/usr/bin/env python
"""bench.py -- GCUPS (fill + traceback, device-timed) of the alignTools DP core on B200.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): batched local
(Smith-Waterman) affine-gap alignment of 1 Mi synthetic 150 bp reads against 500 bp target
windows, score + CIGAR, parameters -m 2 -u -2 -o -5 -e -2.  A "step" is one pass of the hot path
(fill kernels + device traceback) over that batch.  Weak scaling: every rank runs the full batch
on its own GPU (pairs are independent; no collective on the data path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--pairs P]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0.  The oracle (oracle/) is executed only for the `cpu_baseline`
object and for `--impl reference`; it is never on the measured GPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OPS_PER_CELL = {"global": 10, "local": 12, "fit": 10, "fitjump": 14, "overlap": 6, "edit": 6}
# Integer roofline denominator (BASELINE.md 4 / SURVEY.md 8d): the ALU pipe issues 64 int32 lanes per
# clock per SM (ncu: sm__inst_executed_pipe_alu is the binding unit of the fill; tools/int_peak.cu
# measures 56-59 lane-ops/clk/SM for VIMNMX / LOP3 / VIADDMNMX), and one packed s16x2 instruction
# counts as two operations.  profiles/int_peak_r01.txt has the raw numbers.
INT32_LANES_PER_SM = 64
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per PAIR, from the ncu --set full capture of ONE
# launch on the full 1 Mi-pair batch (profiles/ncu_fill_local_final_r01.csv: 739.2 MB read + 44.591 GB written); a launch
# of the bench moves this times its number of pairs.  Algorithmic bytes: 0.5 B/cell pointers + the sequences.
NCU_TRAFFIC_BYTES_PER_PAIR = (739.219456e6 + 44.591399e9) / 1048576
PACKED_FACTOR = 2          # one s16x2 instruction advances two cells ("x2 counted for packed s16x2")


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        inside = [ln for ts, ln in self.lines if t_begin is None or (t_begin <= ts <= t_end + 0.12)]
        if not inside:                       # region shorter than one sampling period: take the sample nearest to it
            inside = [ln for ts, ln in self.lines[-1:]]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        hot = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(hot) if hot else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned_like(arr):
    """Copy a numpy array into pinned host memory (torch allocator) and return a numpy view."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr))
    try:
        t = t.pin_memory()
    except Exception:
        pass
    return t.numpy(), t


def make_workload(pairs, stream):
    from aligntools.c_b200 import synth
    return synth.config2_local(n_pairs=pairs, stream=stream)


def cpu_baseline_sample(w, n_sample, threads, kind_pref="reference"):
    """Time the reference's own CPU implementation (oracle/_ref) -- or the port when the compiled
    reference is absent -- on the first n_sample pairs of the workload."""
    import oracle
    p = oracle.Params(**{k: w["params"][k] for k in "muoej"}, jump=w["params"]["jump"])
    sl = slice(0, n_sample)
    args = (w["mode"], p, w["q"], w["q_off"][sl], w["q_len"][sl], w["t"], w["t_off"][sl], w["t_len"][sl])
    cells = int((w["q_len"][sl].astype(np.uint64) * w["t_len"][sl].astype(np.uint64)).sum())
    if kind_pref == "reference" and oracle.have_ref():
        kind = "reference"
        t0 = time.perf_counter()
        out = oracle.ref_batch(*args, want_aln=True, threads=threads)
        dt = time.perf_counter() - t0
    else:
        kind = "port"
        oracle.build()
        t0 = time.perf_counter()
        out = oracle.port_batch(*args, want_aln=True, threads=threads)
        dt = time.perf_counter() - t0
    return kind, cells, dt, out


def run_reference_arm(args):
    rank, local_rank, world = env_rank()
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    w = make_workload(max(args.ref_pairs, 1024), 0)
    per_step = args.ref_pairs
    times = []
    for k in range(args.warmup + args.steps):
        kind, cells, dt, _ = cpu_baseline_sample(w, per_step, threads)
        if k >= args.warmup:
            times.append(dt)
    tot = sum(times)
    gcups = cells * len(times) / tot / 1e9
    line = {
        "impl": "reference", "metric": "GCUPS (fill+traceback) of the reference CPU path", "value": gcups, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 batched local SW affine 150x500 (-m 2 -u -2 -o -5 -e -2), score+alignment",
                   "pairs_per_step": per_step},
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": threads, "kind": kind,
                         "sample": f"{per_step} pairs of the C2 workload per step, one pthread per host core"},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=1 << 20, help="pairs per GPU (C2 = 1 Mi)")
    ap.add_argument("--ref-pairs", type=int, default=4096, help="pairs per step of the --impl reference arm")
    ap.add_argument("--cpu-sample", type=int, default=12000, help="pairs timed for cpu_baseline (about 14 s on one core)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    rank, local_rank, world = env_rank()
    import torch
    import torch.distributed as dist
    use_dist = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if use_dist:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL logs (its version banner) go to stdout by default: rank 0 prints ONE JSON line there
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import aligntools.c_b200 as A
    al = A.Aligner(devices=[local_rank])
    w = make_workload(args.pairs, rank)
    opt = A.Opt(**w["params"])
    # pinned host copies (the e2e leg copies from these every step)
    q, _kq = pinned_like(w["q"]); t, _kt = pinned_like(w["t"])
    qo, _kqo = pinned_like(w["q_off"]); to, _kto = pinned_like(w["t_off"])
    ql, _kql = pinned_like(w["q_len"]); tl, _ktl = pinned_like(w["t_len"])

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM, device-timed ----------------
    batch = al.batch("local", opt, q, qo, ql, t, to, tl, out_flags=A.OUT_CIGAR)
    sampler = ClockSampler(local_rank)
    sampler.start()                      # nvidia-smi needs a moment to start: begin before the warm-up, keep only the timed region
    for _ in range(args.warmup):
        batch.run()
    barrier()
    dev_ms, fill_ms, tb_ms, launches, kern_ms = 0.0, 0.0, 0.0, 0, []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tm = batch.run()
        dev_ms += tm.device_ms; fill_ms += tm.fill_ms; tb_ms += tm.traceback_ms; launches += tm.launches
        kern_ms.append(tm.fill_kernel_ms)
    barrier()
    t_end = time.perf_counter()
    wall_ms = (t_end - t0) * 1e3
    clocks = sampler.stop(t0, t_end)
    res = batch.fetch()
    cells = tm.cells
    if use_dist:
        tt = torch.tensor([dev_ms, wall_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(tt[0]), float(tt[1])
        tl_ = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(tl_, op=dist.ReduceOp.SUM)
        launches = int(tl_[0])
    value = cells * world * args.steps / (dev_ms * 1e-3) / 1e9
    score_sum = int(res.score.astype(np.int64).sum())
    cigar_ops = int(res.cigar_off[-1])
    batch.free()

    # ---------------- e2e: host buffers through the public C-ABI call ----------------
    # at_batch_align (one-shot): H2D of the sequences from pinned host memory, fill + traceback,
    # D2H of scores / cells / CIGARs into host buffers, all inside the timed region.  The library
    # pipelines sub-slices over several streams so the copies overlap the kernels.
    e2e = None
    if not args.no_e2e:
        h2d = q.nbytes + t.nbytes + qo.nbytes + to.nbytes + ql.nbytes + tl.nbytes
        d2h = 0
        e2e_times = []
        out_bufs = A.BatchResult(args.pairs)
        cig_cap = max(2 * cigar_ops + 1024, 1 << 16)
        out_bufs.cigar, _kc = pinned_like(np.zeros(cig_cap, np.uint32))
        out_bufs.cigar_off = np.zeros(args.pairs + 1, np.uint64)
        e2e_warm = max(3, args.warmup)             # the first calls size the library's workspaces
        for k in range(e2e_warm + args.e2e_steps):
            barrier()
            t1 = time.perf_counter()
            r2 = al.align_arrays("local", opt, q, qo, ql, t, to, tl, out_flags=A.OUT_CIGAR, out=out_bufs, cigar_cap=cig_cap)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t1
            if k >= e2e_warm:
                e2e_times.append(dt)
            if os.environ.get("AT_BENCH_VERBOSE"):
                print(f"[bench] e2e iteration {k}: {1e3 * dt:.2f} ms", file=sys.stderr, flush=True)
            d2h = r2.score.nbytes + 4 * r2.end_i.nbytes + (r2.cigar_off.nbytes - 8) + int(r2.cigar_off[-1]) * 4
        e2e_ms = 1e3 * sum(e2e_times) / len(e2e_times)
        if use_dist:
            tt = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_ms = float(tt[0])
        assert int(r2.score.astype(np.int64).sum()) == score_sum and int(r2.cigar_off[-1]) == cigar_ops
        assert np.array_equal(r2.cigar[:cigar_ops], res.cigar[:cigar_ops]), "pipelined e2e CIGARs differ from the resident run"
        e2e = {"value": cells * world / (e2e_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "kernel_ms_sum": r2.timing.fill_ms + r2.timing.traceback_ms, "launches_per_step": int(r2.timing.launches),
               "api": "at_batch_align (one call: pinned host buffers in, host buffers out; sub-slices pipelined over 3 streams)"}

    # ---------------- roofline of the dominant kernel (the local fill) ----------------
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    k_ms = sum(kern_ms) / len(kern_ms)
    int_peak = 148 * INT32_LANES_PER_SM * PACKED_FACTOR * sm_max_mhz * 1e6 / 1e12     # T int-op/s
    achieved = tm.fill_kernel_cells * OPS_PER_CELL["local"] / (k_ms * 1e-3) / 1e12
    roofline = {"bound": "int_alu", "achieved": achieved, "peak": int_peak, "unit": "Tiop/s", "frac": achieved / int_peak,
                "traffic": NCU_TRAFFIC_BYTES_PER_PAIR * args.pairs, "traffic_note": "ncu dram read+write of one launch (profiles/ncu_fill_local_final_r01.csv), scaled by pairs",
                "algorithmic_bytes": int(0.5 * tm.fill_kernel_cells) + int(q.nbytes + t.nbytes),      # 4-bit pointer per cell + the sequences
                "kernel": "at_fill_affine<LOCAL,R=5,JUMP=0,PACKED=s16x2,PROF=1>", "kernel_ms": k_ms,
                "kernel_gcups": tm.fill_kernel_cells / (k_ms * 1e-3) / 1e9,
                "ops_per_cell": OPS_PER_CELL["local"],
                "peak_note": (f"148 SM x {INT32_LANES_PER_SM} ALU lanes x {PACKED_FACTOR} (s16x2) x {sm_max_mhz:.0f} MHz nominal "
                              "(BASELINE.md 4); algorithmic ops/cell from SURVEY.md 8(d)"),
                "hbm": {"bound": "hbm", "achieved": tm.ptr_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": tm.ptr_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak, "what": "traceback-pointer writes", "peak_src": peak_src}}

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        kind, ccells, dt, out = cpu_baseline_sample(w, args.cpu_sample, 1)
        assert np.array_equal(out.score[:args.cpu_sample], res.score[:args.cpu_sample].astype(np.int64)), "GPU/CPU score mismatch"
        cpu = {"value": ccells / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": kind,
               "sample": f"first {args.cpu_sample} pairs of the same C2 batch, single thread ({dt:.1f} s); scores cross-checked against the GPU run"}

    if rank == 0:
        line = {
            "metric": "GCUPS (fill+traceback, device-timed)", "value": value, "unit": "GCUPS",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s16x2", "data": "synthetic",
            "config": {"workload": "C2 batched local SW affine: 150 bp reads vs 500 bp windows (-m 2 -u -2 -o -5 -e -2), score + CIGAR",
                       "pairs_per_gpu": args.pairs, "l1": 150, "l2": 500, "parallelism": f"pairs sharded x{world}, no collective",
                       "l2_policy": "inputs (650 MB sequences + 42 GB pointer arena) are larger than L2"},
            "fill_ms_per_step": fill_ms / args.steps, "traceback_ms_per_step": tb_ms / args.steps,
            "wall_ms_per_step": wall_ms / args.steps,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu,
            "checks": {"score_sum": score_sum, "cigar_ops": cigar_ops},
        }
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
