#!/usr/bin/env python
"""bench.py -- acquisition evaluations / second on the RF + EI hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c1|c5rf] [--candidates n]

Default workload = BASELINE.json configs[2] ("C3"): 2-objective mixed ordinal/categorical space (D_enc = 23),
256-tree Hutter forests, random_scalarizations EI + tchebyshev over 16 777 216 candidates per GPU, stable top-20.
One "step" = one pass of the hot path over the whole pool: fused forest traversal + EI scalarisation (hm_rf_eval)
-> stable top-k (hm_topk) [-> all-gather + merge of the per-rank top-k when N > 1].  Weak scaling: every rank
scores its own 16M-candidate shard (seed + rank), model replicated, the only collective is the k-pair all-gather.

value  : evaluations/s with the encoded candidates already resident in HBM (CUDA events, max over ranks)
e2e    : the same through the host-buffer C-ABI call (hm_rf_score_host): pinned host candidates -> H2D -> kernels ->
         D2H of every score + the top-k, all inside the timed region
roofline: the dominant kernel (hm_forest_kernel) against the measured HBM copy bandwidth, algorithmic bytes =
         (4*D_enc + 8) per evaluation (fp32 encoded candidate in, fp64 score out)
cpu_baseline: the CPU oracle port (oracle/hm_oracle.c, OpenMP over all host threads + numpy EI) on a bounded sample
--impl reference: times that CPU port alone (the reference itself is Python on sklearn and cannot travel to the box)
"""
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

METRIC = "acquisition_evals_per_sec"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c1", "c5rf"])
    ap.add_argument("--candidates", type=int, default=0, help="candidates per GPU (default: the workload's)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="candidates in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def make_workload(name):
    from hypermapper_b200 import workloads
    if name == "c3":
        return workloads.RFWorkload(workloads.scenario_c3(), n_train=2000, seed=5), 1 << 24, \
            "C3: 2-objective mixed ordinal/categorical (D_enc=23), 256-tree RF x2, EI+tchebyshev, top-20"
    if name == "c1":
        return workloads.RFWorkload(workloads.scenario_c1(), n_train=15, seed=0), 1 << 26, \
            "C1 shape: 2 real inputs (Branin), 10-tree RF, EI+tchebyshev, top-20"
    if name == "c5rf":
        return workloads.RFWorkload(workloads.scenario_c5(), n_train=4000, seed=9), 1 << 24, \
            "C5-RF: 20-D mixed space (D_enc=32), 256-tree RF, EI+tchebyshev, top-20"
    raise SystemExit("unknown workload")


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [x.strip() for x in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(wl, sample, repeats=1):
    """The CPU port of the hot path on `sample` candidates with all host threads: OpenMP forest traversal +
    Hutter reduction (oracle/hm_oracle.c) per objective, numpy/scipy EI scalarisation, stable top-k."""
    from oracle import oracle
    from hypermapper_b200.forest import forest_tables
    fas = {}
    for o, m in wl.models.items():
        fas[o] = oracle.forest_arrays_from_model(m)
    Xcm = wl.candidates_encoded(sample, seed=1234)
    X = np.ascontiguousarray(Xcm.T)          # the reference's row-major float32 layout
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        means, variances = {}, {}
        for o, fa in fas.items():
            means[o], variances[o], _ = oracle.rf_mean_var_omp(fa, X)
        vals = oracle.EI(means, variances, wl.data_array, wl.objective_weights, wl.scalarization, wl.objective_limits)
        oracle.get_min_indices_fast(vals, 20)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample / best, best, oracle.num_threads()


def run_reference(args, rank, world):
    """--impl reference: the CPU port alone (rank 0 only)."""
    if rank != 0:
        return
    wl, n_default, desc = make_workload(args.workload)
    sample = args.cpu_sample or 200000
    times = []
    for s in range(args.warmup + args.steps):
        rate, dt, threads = cpu_port_rate(wl, sample)
        if s >= args.warmup:
            times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = sample / (ms / 1000.0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": desc, "candidates_per_step": sample, "note": "bounded sample of the same workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d candidates per step, OpenMP C port of models.py:174-275 + numpy EI" % sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from hypermapper_b200 import forest, _lib
    from hypermapper_b200.distributed import allgather_topk
    from hypermapper_b200.scoring import HostScorer

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl, n_default, desc = make_workload(args.workload)
    N = args.candidates or n_default
    k = 20
    forests = [forest.device_forest(wl.models[o]) for o in wl.objectives]
    acq = wl.acq_params()
    D = wl.D_enc

    # candidates of this rank: pinned host copy (for e2e) + HBM-resident copy (for value)
    x_host = torch.empty((D, N), dtype=torch.float32).pin_memory()
    wl.candidates_encoded(N, seed=6 + rank, out=x_host.numpy())
    x_dev = x_host.cuda(non_blocking=False)
    index_base = rank * N
    L = _lib.lib()
    ws = torch.empty((int(L.hm_topk_workspace_bytes(k)),), dtype=torch.uint8, device="cuda")

    kernel_ms = []
    ev_pairs = []

    def step(record):
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        out = forest.rf_eval(forests, x_dev, acq=acq)
        if record:
            e1.record()
            ev_pairs.append((e0, e1))
        tv, ti = forest.topk(out["score"], k, index_base=index_base)
        if world > 1:
            tv, ti = allgather_topk(tv, ti, k)
        return tv, ti

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        res = step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        res = step(True)
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = t_start.elapsed_time(t_end)
    kernel_ms = [a.elapsed_time(b) for a, b in ev_pairs]
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * N / (ms_per_step / 1000.0)
    launches_per_step = 3 + (2 if world > 1 else 0)    # forest kernel + top-k stage 1/2 (+ merge stage 1/2)

    # e2e: host buffers through the C-ABI host entry point
    e2e = None
    if not args.no_e2e:
        scorer = HostScorer(D, chunk=1 << 21)
        scores_host = torch.empty((N,), dtype=torch.float64).pin_memory()
        for _ in range(max(1, min(args.warmup, 2))):
            scorer.score(forests, x_host, acq, k=k, scores_out=scores_host)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            _, tv_h, ti_h = scorer.score(forests, x_host, acq, k=k, scores_out=scores_host)
            if world > 1:
                gv, gi = allgather_topk(torch.from_numpy(tv_h).cuda(), torch.from_numpy(ti_h + index_base).cuda(), k)
                gi.cpu()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        td = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        dt = float(td.item())
        e2e = {"value": world * N * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": 4 * D * N * world,
               "d2h_bytes_per_step": (8 * N + 16 * k) * world, "ms_per_step": 1000.0 * dt / args.steps,
               "api": "hm_rf_score_host (C-ABI, pinned host candidates, 2M-candidate double-buffered chunks)"}
        # cross-check: same top-k as the HBM-resident path
        if world == 1:
            assert np.array_equal(ti_h, res[1].cpu().numpy()), "e2e and device-resident top-k differ"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_hbm_peak()
    kms = statistics.mean(kernel_ms)
    alg_bytes = (4 * D + 8) * N
    achieved = alg_bytes / (kms / 1000.0) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "hm_forest_kernel<false>", "kernel_ms": kms,
                "algorithmic_bytes_per_eval": 4 * D + 8, "peak_source": peak_src,
                "share_of_step": kms / ms_per_step,
                "note": "large forests are shared-memory/issue bound, not HBM bound (SURVEY 8(d)); see DESIGN.md, profiles/"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "b200",
        "config": {"workload": desc, "candidates_per_gpu": N, "D_enc": D, "trees": wl.n_trees,
                   "objectives": len(wl.objectives), "k": k, "parallelism": "candidate-sharded x%d" % world,
                   "l2": "inputs (%.2f GB per GPU) larger than L2, no flush needed" % (4 * D * N / 1e9)},
        "roofline": roofline,
        "clocks": clocks,
        "gpu_launches": launches_per_step * args.steps,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample or 200000
        rate, dt, threads = cpu_port_rate(wl, sample)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d candidates of the same pool (%.1f s), OpenMP C port of models.py:174-275 "
                                          "+ numpy EI + top-20" % (sample, dt)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
