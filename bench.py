#!/usr/bin/env python3
"""bench.py -- acquisition cells/s of the B200 engine on BASELINE.json's workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2] [--impl ours|reference]

A "step" is one pass of the hot path (front end + PRN x Doppler x code-phase search + best-Doppler pick)
over one batch of synthetic captures.  Default workload = BASELINE configs[1] (cfg2: 32 GPS PRNs, +-10 kHz at
half-bin spacing = 161 Doppler indices, 20 non-coherent 4 ms blocks); every rank searches its own
independent capture (weak scaling, no data-path collective; the 768-byte record arrays are gathered with
NCCL inside the timed region when N > 1).

One JSON line on stdout (rank 0):
  value      whole-job cells/s with the captures already resident in HBM (device timing, CUDA events, max over ranks)
  e2e        the same metric through the reference-facing C-ABI call acq_search() with HOST buffers
             (pinned), host->device and device->host copies inside the timed region
  roofline   dominant kernel (fused correlate + inverse FFT + peak search) against the on-SM roofs measured
             by micro-benchmark in this run (FP32 issue, shared-memory bandwidth) and against measured HBM
  cpu_baseline  the oracle port timed on this box's host cores on a bounded sample (N = 1, rank 0)
`--impl reference` times the reference's own CPU path instead (see reference_arm()).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "acquisition cells/sec (PRN x Doppler x code-phase)"
UNIT = "cells/s"
# algorithmic work per tile (one inverse FFT of one (sat, Doppler, block))
FLOP_PER_TILE = {4092: 6 * 16384 + 5 * 16384 * 14 + 3 * 4092, 16368: 6 * 16384 + 5 * 16384 * 14 + 3 * 16368}
SMEM_BYTES_PER_TILE = {4092: 131072 + 1048576 + 4092 * 8, 16368: 131072 + 1048576 + 16368 * 8}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).
    nvidia-smi needs a few hundred ms to deliver its first sample, longer than a short timed region: it is started
    before the warm-up, samples carry timestamps, and only those between mark_start() and mark_end() are reported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    @staticmethod
    def _epoch(stamp):
        import datetime
        try:
            return datetime.datetime.strptime(stamp, "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        rows = []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                row = (float(f[1]), float(f[2]))
            except ValueError:
                continue
            rs = {name for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9])
                  if v.lower().startswith("active")}
            rows.append((self._epoch(f[9]) if len(f) > 9 else None, row, rs))
        if not rows:
            return None
        inside = [r for r in rows if r[0] is not None and self.t0 is not None and self.t1 is not None
                  and self.t0 <= r[0] <= self.t1]
        use = inside or rows  # a region shorter than the sampling period: fall back to every sample of the run
        reasons = set().union(*[r[2] for r in use])
        return {"sm_mhz": statistics.median([r[1][0] for r in use]), "sm_max_mhz": max(r[1][1] for r in use),
                "reasons": sorted(reasons), "samples": len(use), "samples_in_timed_region": len(inside),
                "samples_total": len(rows), "period_ms": 20}


# ------------------------------------------------------------------------------------------ workload
def build_workload(cfg, rank, captures_per_gpu):
    """Synthetic captures for this rank (numpy, host).  Returns table, params kwargs, packed bytes [n_cap, bytes]."""
    from flydog_sdr_gps_b200 import scenarios, synth
    table = scenarios.table(cfg)
    kw = scenarios.params_kw(cfg)
    k = kw.get("k_noncoh", 1)
    caps = [synth.make_capture(10_000 * rank + c, k, table, scenarios.signals(cfg if cfg != "cfg5" else "cfg1", rank * 1000 + c))
            for c in range(captures_per_gpu)]
    return table, kw, np.stack(caps)


def cpu_baseline(cfg, table, kw, packed, budget_s=12.0):
    """Oracle port on the host cores, bounded sample of the same workload (same capture bytes)."""
    from oracle import oracle_py as O
    O.build(ref=False)
    nthreads = os.cpu_count() or 1
    prm = O.default_params(**{k: v for k, v in kw.items()})
    n_dop = prm.dop_hi - prm.dop_lo + 1
    # size the sample: ~0.35 ms per tile per core
    tiles_per_sat = n_dop * prm.k_noncoh
    n_sats = max(1, min(len(table), int(budget_s * nthreads / (tiles_per_sat * 0.35e-3))))
    if n_sats >= nthreads:
        n_sats -= n_sats % nthreads
    sel = np.arange(n_sats, dtype=np.int32)
    O.search(packed[0], table, sel=sel[:1], params=prm, nthreads=1)  # plans, page-in
    t0 = time.perf_counter()
    O.search(packed[0], table, sel=sel, params=prm, nthreads=nthreads)
    dt = time.perf_counter() - t0
    cells = sum(n_dop * (16368 if table[s][3] == 3 else 4092) for s in sel)
    return {"value": cells / dt, "unit": UNIT, "cores": min(nthreads, n_sats), "kind": "port",
            "sample": "%d of %d PRNs of one %s capture, all %d Doppler indices, K=%d (%.1f s); "
                      "oracle = search.cpp restated + in-repo FFT (FFTW unavailable)" % (
                          n_sats, len(table), cfg, n_dop, prm.k_noncoh, dt),
            "tiles_per_s": n_sats * tiles_per_sat / dt}


def _ref_worker(args):
    packed, sats = args
    from oracle import oracle_py as O
    t0 = time.perf_counter()
    O.ref_search(packed, np.asarray(sats, np.int32))
    return time.perf_counter() - t0


def literal_reference_rate(packed_block, n_procs, sats_per_proc=4):
    """The unmodified search.cpp (oracle/_ref): Sample()+Correlate() per sat, forked over host cores
    (processes, not threads: the reference keeps its buffers in file statics)."""
    from oracle import oracle_py as O
    if not O.have_ref():
        return None
    import multiprocessing as mp
    O.ref()  # SearchInit once, inherited by fork
    jobs = [(packed_block, [(p * sats_per_proc + k) % 32 for k in range(sats_per_proc)]) for p in range(n_procs)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(n_procs) as pool:
        pool.map(_ref_worker, jobs)
    dt = time.perf_counter() - t0
    cells = n_procs * sats_per_proc * 41 * 4092
    return {"value": cells / dt, "unit": UNIT, "cores": n_procs,
            "sample": "%d x %d Navstar sats, reference defaults (41 bins), one block" % (n_procs, sats_per_proc)}


# ------------------------------------------------------------------------------------------ reference arm
def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores.
    The default workload (cfg2) uses extensions search.cpp cannot express (half-bins, K=20), so the timed
    code is the oracle port of search.cpp (asserted bit-identical to the unmodified search.cpp at its
    defaults by tests/test_oracle_cpu.py); the literal search.cpp rate on its own config is attached."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    table, kw, packed = build_workload(args.config, 0, 1)
    steps, warm = args.steps, args.warmup
    from oracle import oracle_py as O
    O.build(ref=False)
    nthreads = os.cpu_count() or 1
    prm = O.default_params(**kw)
    n_dop = prm.dop_hi - prm.dop_lo + 1
    tiles_per_sat = n_dop * prm.k_noncoh
    # each step = a bounded sample sized for ~3 s on all cores
    n_sats = max(1, min(len(table), int(3.0 * nthreads / (tiles_per_sat * 0.35e-3))))
    sel = np.arange(n_sats, dtype=np.int32)
    cells = sum(n_dop * (16368 if table[s][3] == 3 else 4092) for s in sel)
    for _ in range(min(warm, 1)):
        O.search(packed[0], table, sel=sel, params=prm, nthreads=nthreads)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.search(packed[0], table, sel=sel, params=prm, nthreads=nthreads)
    dt = (time.perf_counter() - t0) / steps
    value = cells / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.config, "detail": __import__("flydog_sdr_gps_b200").scenarios.CONFIGS[args.config]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(nthreads, n_sats), "kind": "port",
                             "sample": "each step: %d of %d PRNs x %d Doppler indices x K=%d of one capture" % (
                                 n_sats, len(table), n_dop, prm.k_noncoh)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    lit = literal_reference_rate(packed[0][:8192], nthreads)
    if lit:
        line["literal_search_cpp"] = lit
    print(json.dumps(line), flush=True)
    return 0


def cufft_point(torch, stream, batch=256, reps=20):
    """cuFFT through torch.fft.ifft: `batch` unnormalised-size 16384-point complex64 inverse transforms per call,
    input and output resident (32 MiB each way, L2-sized).  This is ONLY the transform: the engine's tile also
    forms the product, the power, the non-coherent sum and the peak search, and never writes the lags."""
    x = torch.randn(batch, 16384, dtype=torch.complex64, device="cuda")
    for _ in range(3):
        y = torch.fft.ifft(x, norm="forward")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        y = torch.fft.ifft(x, norm="forward")
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    del y
    return {"what": "torch.fft.ifft (cuFFT) %d x 16384 complex64, out-of-place, transform only" % batch,
            "ms_per_call": ms, "transforms_per_s": batch / (ms * 1e-3)}


# ------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--captures-per-gpu", type=int, default=0, help="0 = 1 (cfg5: 128)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cufft", action="store_true",
                    help="also time torch.fft.ifft (cuFFT) on a batch of 16384-point transforms: the permitted "
                         "timed comparison point (inverse FFT only, no product / power / peak search)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import flydog_sdr_gps_b200 as F

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n_cap = args.captures_per_gpu or (128 if args.config == "cfg5" else 1)
    table, kw, packed = build_workload(args.config, rank, n_cap)
    eng = F.AcqEngine(table, F.default_params(**kw), device=local)
    n_sel = len(table)
    cells_step = eng.cells_per_search() * n_cap      # per GPU per step
    tiles_step = eng.tiles_per_search() * n_cap
    lags = {16368 if r[3] == 3 else 4092 for r in table}

    # resident inputs / outputs
    d_in = torch.from_numpy(packed.reshape(-1)).cuda()
    d_out = torch.zeros(n_cap * n_sel * 24, dtype=torch.uint8, device="cuda")
    gathered = torch.zeros(world * d_out.numel(), dtype=torch.uint8, device="cuda") if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    h_in = torch.from_numpy(packed.reshape(-1)).pin_memory()
    h_out = torch.zeros(n_cap * n_sel * 24, dtype=torch.uint8).pin_memory()
    # a non-default stream: its handle is non-NULL, so the engine launches on exactly this stream and
    # torch.cuda.Event timing brackets the engine's kernels (NULL would select the engine's own stream)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    def step_device():
        eng.search_device(d_in.data_ptr(), d_out.data_ptr(), n_cap, stream_ptr=stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_out)

    def step_e2e():
        eng.search_ptr(h_in.data_ptr(), n_cap, h_out.data_ptr())
        if world > 1:
            # records are on the host already; a host gather of 768 B/rank is what a receiver farm would do
            pass

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()

    # ---- device-resident timing: K steps of the product path, L2 flushed between steps (flush outside the
    # event pairs).  No events between the kernels here: the four launches of a step are chained by
    # programmatic dependent launch, which an event record in between would switch off.
    sampler.mark_start()
    launches0 = eng.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in evs:
        flush.fill_(1)
        a.record()
        step_device()
        b.record()
    barrier()
    launches = eng.launch_count - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    # ---- the same K steps once more with CUDA events between the kernels (on the launching stream): per-kernel
    # durations for the roofline.  Still inside the clock-sampling window.
    eng.set_profiling(True)
    step_device()
    barrier()
    kern_ms = []
    evs_p = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evs_p:
        flush.fill_(1)
        a.record()
        step_device()
        b.record()
        kern_ms.append(eng.kernel_ms())  # waits for this step (the flush of the next step is not timed anyway)
    barrier()
    eng.set_profiling(False)
    prof_ms = sum(a.elapsed_time(b) for a, b in evs_p) / args.steps
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = world * cells_step / (ms_per_step * 1e-3)

    # ---- end-to-end timing through acq_search (host buffers, copies inside)
    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * cells_step * args.steps / e2e_s

    # results sanity: device path == host path, bitwise
    same = bool(torch.equal(d_out.cpu(), h_out))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel
    search_ms = statistics.mean(k["search"] for k in kern_ms)
    step_kern_ms = statistics.mean(sum(k.values()) for k in kern_ms)
    mb = F.microbench(local)
    flop = sum(eng.params.k_noncoh * eng.n_dop * FLOP_PER_TILE[16368 if r[3] == 3 else 4092] for r in table) * n_cap
    smem_b = sum(eng.params.k_noncoh * eng.n_dop * SMEM_BYTES_PER_TILE[16368 if r[3] == 3 else 4092] for r in table) * n_cap
    hbm_b = n_cap * eng.params.k_noncoh * 8192 + len(table) * 16384 * 8 + 24 * n_sel * n_cap
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    fp32_ach = flop / (search_ms * 1e-3) / 1e12
    smem_ach = smem_b / (search_ms * 1e-3) / 1e12
    hbm_ach = hbm_b / (search_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "search_kernel_traffic.json"))).get(args.config)
        traffic = traffic and traffic.get("bytes")
    except Exception:
        pass
    # The binding roof is on-SM: ncu shows the L1/shared-memory data pipe as the busiest unit
    # (profiles/), the FMA pipe second, HBM idle.  `achieved` uses SURVEY's algorithmic shared-memory byte
    # model per tile; the kernel itself moves fewer bytes (3 exchange passes + output pruning).
    roofline = {
        "bound": "smem",
        "kernel": "k_search_l1 / k_search_e1b (conj-multiply + 16384-pt inverse FFT + |.|^2 + peak search)",
        "achieved": smem_ach * 1e3, "peak": mb["smem_tbs"] * 1e3, "unit": "GB/s", "frac": smem_ach / mb["smem_tbs"],
        "peak_source": "shared-memory micro-benchmark in this run (acq_microbench: conflict-free 8-byte LDS+STS); "
                       "nominal 148 SM x 128 B/clk",
        "algorithmic_bytes_per_launch": smem_b, "model": "SURVEY 8(d): %d B of shared-memory traffic per tile" %
        SMEM_BYTES_PER_TILE[max(lags)],
        "kernel_ms": search_ms, "kernel_share_of_step": search_ms / step_kern_ms,
        "traffic": traffic,
        "achieved_smem": smem_ach, "achieved_fp32": fp32_ach, "achieved_hbm": hbm_ach,
        "fp32": {"bound": "fp32", "achieved": fp32_ach, "peak": mb["ffma_tflops"], "unit": "TFLOP/s",
                 "frac": fp32_ach / mb["ffma_tflops"], "algorithmic_flop_per_launch": flop,
                 "peak_source": "FFMA micro-benchmark in this run; nominal 148 SM x 128 lanes x 2 x clock"},
        "hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": hbm_b},
        "microbench": mb,
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.config, "detail": F.scenarios.CONFIGS[args.config], "captures_per_gpu": n_cap,
                   "sats": n_sel, "doppler_indices": eng.n_dop, "k_noncoh": eng.params.k_noncoh,
                   "cells_per_step_per_gpu": cells_step, "tiles_per_step_per_gpu": tiles_step,
                   "l2": "256 MiB flush write between timed steps, outside the event pairs",
                   "sharding": "independent captures per rank; records gathered with NCCL all_gather" if world > 1
                   else "single GPU"},
        "tiles_per_s": world * tiles_step / (ms_per_step * 1e-3),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h_in.numel()),
                "d2h_bytes_per_step": int(h_out.numel()), "ms_per_step": e2e_s / args.steps * 1e3,
                "api": "acq_search (C ABI, pinned host buffers)"},
        "gpu_launches": int(launches),
        "kernel_ms": {k: statistics.mean(x[k] for x in kern_ms) for k in kern_ms[0]},
        "kernel_ms_pass": {"what": "second pass of the same K steps with CUDA events between the kernels "
                                   "(which disables programmatic dependent launch between them)",
                           "ms_per_step": prof_ms},
        "device_equals_host_path": same,
        "clocks": clocks,
        "roofline": roofline,
    }
    if args.cufft:
        line["cufft_comparison"] = cufft_point(torch, stream)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.config, table, kw, packed)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
