#!/usr/bin/env python
"""bench.py -- acquisition evaluations / second on the surrogate-prediction + acquisition hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload c3|c3f32|c1|c5rf|c2|c5gp|c4|c3raw|c3pool] [--candidates n]

Default workload = BASELINE.json configs[2] ("C3"), the configuration the north-star's RF+EI target is quoted on:
2-objective mixed ordinal/categorical space (D_enc = 23), two 256-tree Hutter forests, random_scalarizations
EI + tchebyshev over 16 777 216 candidates per GPU, stable top-20.
One "step" = one pass of the hot path over the whole pool:
  RF : fused forest traversal + Hutter reduction + EI scalarisation (hm_rf_eval) -> stable top-k (hm_topk)
       c3 (all-discrete space): the candidates are held as packed 64-bit codes (hm_rf_eval_packed, csrc/hm_pack.cuh);
       c3f32 is the same workload through the fp32 encoded matrix (the round-1 path), bit-identical results
  GP : hm_gp_mean_var -> hm_acq_score -> hm_topk                      (c2: Hartmann-6, n=256; c5gp: 20-D, n=512)
  c4 : hm_pareto_filter on 10^7 x 3 points (metric: points/s)
  c3raw : C3 with the candidates as RAW parameter values, the form the reference API takes: hm_encode -> RF step
  c3pool: C3 with the pool drawn on the device: hm_sample_pool (generate + encode) -> RF step
[-> all-gather + merge of the per-rank top-k when N > 1].  Weak scaling: every rank owns its own shard (seed + rank),
model replicated, the only collective is the k-pair all-gather.

value   : units/s with the inputs already resident in HBM (CUDA events on the launching stream, max over ranks)
e2e     : the same through host buffers: pinned host inputs -> H2D -> kernels -> D2H of the results, all timed
roofline: the dominant kernel against the measured peak (HBM copy bandwidth from MEASURED_PEAKS.json; FP64 DFMA rate
          measured in-process for the GP kernel)
cpu_baseline / --impl reference: the CPU oracle port on all host threads, bounded sample (the reference itself is
          Python on sklearn/GPy and cannot travel to the GPU box; kind = "port")
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

WORKLOADS = ["c3", "c3f32", "c1", "c5rf", "c2", "c5gp", "c4", "c3raw", "c3pool"]

# roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel from an
# `ncu --set full` capture of the same command at the same size (profiles/<file>); None where no capture at that size exists
# on-chip utilisation of the same launch in the same capture (per cent of peak): what actually bounds the kernel
NCU_ONCHIP = {
    ("c3", 1 << 24): {"l1tex_data_pipe_pct": 76.1, "issue_slots_pct": 55.3, "dram_pct": 0.3,
                      "source": "profiles/r02_final_forest_c3_ncu_raw.csv"},
    ("c3f32", 1 << 24): {"l1tex_data_pipe_pct": 87.9, "issue_slots_pct": 63.5, "dram_pct": 0.7,
                         "source": "profiles/r01_forest_c3_default_ncu_raw.csv"},
    ("c1", 1 << 26): {"l1tex_data_pipe_pct": 88.3, "issue_slots_pct": 77.0, "dram_pct": 5.1,
                      "source": "profiles/r02_final_forest_c1_ncu_raw.csv"},
}
NCU_TRAFFIC = {
    ("c3", 1 << 24): (0.661674e9 + 0.496786e9, "profiles/r02_final_forest_c3_ncu_raw.csv"),
    ("c3f32", 1 << 24): (2.234378e9 + 0.545148e9, "profiles/r01_forest_c3_default_ncu_raw.csv"),
    ("c1", 1 << 26): (0.538545e9 + 0.498207e9, "profiles/r02_final_forest_c1_ncu_raw.csv"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=WORKLOADS)
    ap.add_argument("--candidates", type=int, default=0, help="units per GPU (default: the workload's)")
    ap.add_argument("--total-candidates", type=int, default=0,
                    help="STRONG scaling: units of the whole job, split evenly over the GPUs (BASELINE configs[4]: "
                         "67108864 over 2/4/8 GPUs); the JSON line then says \"scaling\": \"strong\"")
    ap.add_argument("--cpu-sample", type=int, default=0, help="units in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--fused-topk", action="store_true",
                    help="RF workloads: device-timed step = hm_rf_score_topk (no score array) instead of hm_rf_eval + hm_topk")
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
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


# =====================================================================================================================
# workloads: each exposes  cpu_port(sample) -> seconds  (no GPU),  and on the GPU side  setup_device(N, rank),
# step(record) -> ((topk values, topk indices), events),  e2e_step() -> (values, indices)
# =====================================================================================================================
class RFBench:
    metric, unit, dtype = "acquisition_evals_per_sec", "evals/s", "f64"
    cpu_desc = "OpenMP C port of models.py:174-275 + numpy/scipy EI + top-20"
    e2e_api = "hm_rf_score_host (C-ABI, pinned host candidates, 2M-candidate double-buffered chunks)"

    def __init__(self, name):
        from bench_support import workloads
        self.name = name
        cache = os.path.join(os.environ.get("HM_BENCH_CACHE", "/tmp"),
                             "hm_bench_wl_v3_%s.pkl" % ("c3" if name.startswith("c3") else name))
        self.packed = name == "c3"
        self.fused_topk = False
        cached = None
        if os.path.exists(cache):
            import pickle
            try:
                with open(cache, "rb") as f:
                    cached = pickle.load(f)
            except Exception:
                cached = None
        if cached is not None:
            self.wl = cached
        if name in ("c3", "c3f32", "c3raw", "c3pool"):
            self.wl = cached or workloads.RFWorkload(workloads.scenario_c3(), n_train=2000, seed=5)
            self.n_default = 1 << 24
            self.desc = "C3: 2-objective mixed ordinal/categorical (D_enc=23), 256-tree RF x2, EI+tchebyshev, top-20"
            if name == "c3":
                self.desc += "; candidates as packed 64-bit codes"
                self.e2e_api = ("hm_rf_score_packed_host (C-ABI, pinned host codes, 8 B per candidate, 2M-candidate "
                                "double-buffered chunks)")
            elif name == "c3f32":
                self.desc += "; candidates as the fp32 encoded matrix"
        elif name == "c1":
            self.wl = cached or workloads.RFWorkload(workloads.scenario_c1(), n_train=15, seed=0)
            self.n_default = 1 << 26
            self.desc = "C1 shape: 2 real inputs (Branin), 10-tree RF, EI+tchebyshev, top-20"
        else:
            self.wl = cached or workloads.RFWorkload(workloads.scenario_c5(), n_train=4000, seed=9)
            self.n_default = 1 << 24
            self.desc = "C5-RF: 20-D mixed space (D_enc=32), 256-tree RF, EI+tchebyshev, top-20"
        if cached is None:
            try:
                import pickle
                tmp = "%s.%d.tmp" % (cache, os.getpid())        # atomic: concurrent ranks never read a partial file
                with open(tmp, "wb") as f:
                    pickle.dump(self.wl, f)
                os.replace(tmp, cache)
            except Exception:
                pass
        self.k = 20
        self.cpu_sample_default = 200000 if name != "c1" else 2000000

    def config(self, N, world):
        wl = self.wl
        return {"workload": self.desc, "candidates_per_gpu": N, "D_enc": wl.D_enc, "trees": wl.n_trees,
                "objectives": len(wl.objectives), "k": self.k, "parallelism": "candidate-sharded x%d" % world,
                "l2": "inputs (%.2f GB per GPU) larger than L2, no flush needed" % (4 * wl.D_enc * N / 1e9)}

    def cpu_port(self, sample):
        """OpenMP forest traversal + Hutter reduction (oracle/hm_oracle.c) per objective, numpy/scipy EI, top-k."""
        from oracle import oracle
        wl = self.wl
        if not hasattr(self, "_fas"):
            self._fas = {o: oracle.forest_arrays_from_model(m) for o, m in wl.models.items()}
            self._Xcpu = {}
        if sample not in self._Xcpu:
            self._Xcpu = {sample: np.ascontiguousarray(wl.candidates_encoded(sample, seed=1234).T)}
            self._raw_sample = wl.candidates_raw(sample, seed=1234) if self.packed else None
        X = self._Xcpu[sample]
        t0 = time.perf_counter()
        means, variances = {}, {}
        for o, fa in self._fas.items():
            means[o], variances[o], _ = oracle.rf_mean_var_omp(fa, X)
        vals = oracle.EI(means, variances, wl.data_array, wl.objective_weights, wl.scalarization, wl.objective_limits)
        top = oracle.get_min_indices_fast(vals, self.k)
        dt = time.perf_counter() - t0
        self._cpu_out = (sample, means, vals, top)
        return dt

    def parity(self):
        """BASELINE.md section 3.6 gate: the GPU path on the very sample the CPU leg just scored with the oracle --
        leaf ids and means bit-exact, acquisition values within 1e-5 relative, top-k indices identical (near ties
        re-scored with the reference's arithmetic, hypermapper_b200/exact.py).  The oracle is the checker here."""
        import torch
        from oracle import oracle
        from hypermapper_b200 import exact, forest
        wl = self.wl
        sample, means, vals, top = self._cpu_out
        X = self._Xcpu[sample]                                       # [sample][D_enc] float32 rows
        X32 = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
        if self.packed:
            # the product path of this workload: packed codes of the same candidates
            codes = torch.from_numpy(self.ds.pack_codes_host(self._raw_sample).view(np.int64)).cuda()
            out = forest.rf_eval_packed(self.forests, codes, want_mean=True, acq=self.acq)
        else:
            out = forest.rf_eval(self.forests, X32, want_mean=True, acq=self.acq)
        mean_exact = all(np.array_equal(out["mean"][m].cpu().numpy(), means[o]) for m, o in enumerate(wl.objectives))
        sc = out["score"].cpu().numpy()
        with np.errstate(all="ignore"):
            err = np.abs(sc - vals) / np.maximum(np.abs(vals), 1e-300)
        err[sc == vals] = 0
        n_leaf = min(4096, sample)
        if self.packed:
            lv = forest.rf_eval_packed(self.forests, codes[:n_leaf].contiguous(), want_leaves=True)["leaves"]
        else:
            lv = forest.rf_eval(self.forests, X32[:, :n_leaf].contiguous(), want_leaves=True)["leaves"]
        leaf_exact = all(np.array_equal(lv[m].cpu().numpy().astype(np.int64),
                                        oracle.rf_apply(self._fas[o], np.ascontiguousarray(X[:n_leaf])))
                         for m, o in enumerate(wl.objectives))

        def exact_rows(idx):
            sub = X32[:, torch.as_tensor(np.asarray(idx, dtype=np.int64), device="cuda")].contiguous()
            leaves = [t.cpu().numpy() for t in forest.rf_eval(self.forests, sub, want_leaves=True)["leaves"]]
            return exact.reference_rf_scores_from_leaves(wl.acquisition, wl.scalarization, wl.models, leaves,
                                                         wl.objective_weights, wl.objective_limits,
                                                         wl.iteration_number, wl.data_array)
        got = exact.exact_min_indices(out["score"], self.k, exact_rows)
        raw = forest.topk(out["score"], self.k)[1].cpu().numpy()
        res = {"sample": sample, "path": "packed codes (hm_rf_eval_packed)" if self.packed else "fp32 matrix (hm_rf_eval)",
               "leaves_bit_exact": bool(leaf_exact), "leaves_checked": n_leaf * wl.n_trees * len(wl.objectives),
               "mean_bit_exact": bool(mean_exact), "score_max_rel_err": float(err.max()), "score_tol": 1e-5,
               "topk_identical": bool(np.array_equal(got, top)), "topk_identical_without_rescore": bool(np.array_equal(raw, top)),
               "near_tie_rescores": dict(exact.stats_counters)}
        res["ok"] = bool(leaf_exact and mean_exact and err.max() <= 1e-5 and res["topk_identical"])
        return res

    def setup_device(self, N, rank):
        import torch
        from hypermapper_b200 import forest
        from hypermapper_b200.scoring import HostScorer
        wl = self.wl
        self.N, self.rank = N, rank
        self.forests = [forest.device_forest(wl.models[o]) for o in wl.objectives]
        self.acq = wl.acq_params()
        self.scores_host = torch.empty((N,), dtype=torch.float64).pin_memory()
        self.scorer = HostScorer(wl.D_enc, chunk=int(os.environ.get("HM_BENCH_CHUNK", 1 << 21)))
        if self.packed:
            from hypermapper_b200.space_device import device_space_for
            self.ds = device_space_for(wl.space)
            forest.attach_packing(self.forests, self.ds)
            self.codes_host = torch.empty((N,), dtype=torch.int64).pin_memory()
            raw = np.empty((min(N, 1 << 21), self.ds.D), dtype=np.float64)
            rng = np.random.default_rng(6 + rank)
            from bench_support import workloads
            inputs = wl.space.get_input_parameters()
            ch = self.codes_host.numpy().view(np.uint64)
            for lo in range(0, N, 1 << 21):                      # the candidates candidates_encoded(N, 6 + rank) encodes
                hi = min(N, lo + (1 << 21))
                cols = workloads.raw_uniform_columns(wl.space, hi - lo, rng)
                for j, p in enumerate(inputs):
                    raw[:hi - lo, j] = cols[p]
                self.ds.pack_codes_host(raw[:hi - lo], out=ch[lo:hi])
            self.codes_dev = self.codes_host.cuda()
            # the fp32 form of the first candidates, for the per-evaluation shared-memory byte count
            self.x_dev = torch.from_numpy(wl.candidates_encoded(min(N, 1 << 16), seed=6 + rank)).cuda()
            n_groups = sum(f.packed_groups for f in self.forests)
        else:
            self.x_host = torch.empty((wl.D_enc, N), dtype=torch.float32).pin_memory()
            wl.candidates_encoded(N, seed=6 + rank, out=self.x_host.numpy())
            self.x_dev = self.x_host.cuda()
            n_groups = sum(f.n_groups for f in self.forests if f.staged)
        # hm_forest_kernel + hm_topk_stage1/2; forests that stream through shared memory add the coherence pre-pass
        # (hm_forest_sortkey, hm_forest_scan, hm_forest_scatter_rows / _codes) for pools of >= 2^17 candidates
        self.prepass = n_groups > 2 and N >= (1 << 17)
        self.smem_bytes_per_eval = self._smem_bytes_per_eval()
        self.launches_per_step = 3 + (3 if self.prepass else 0)

    def _smem_bytes_per_eval(self, sample=1 << 16):
        """Bytes one evaluation moves through the shared-memory / L1 data pipe: per tree, 12 B per internal node on the
        candidate's path (8 B node + 4 B feature value) + the 8 B leaf node.  Path lengths come from the leaves the
        kernel itself reports (hm_rf_eval apply output) on a sample of the pool and the depth of every sklearn node."""
        import torch
        from hypermapper_b200 import forest
        wl = self.wl
        n = min(sample, self.N, self.x_dev.shape[1])
        per_visit = 8 if self.packed else 12           # packed codes: the node only; fp32 path: node + feature value
        leaves = forest.rf_eval(self.forests, self.x_dev[:, :n].contiguous(), want_leaves=True)["leaves"]
        total = 0.0
        for m, o in enumerate(wl.objectives):
            lv = leaves[m].cpu().numpy()
            for t, est in enumerate(wl.models[o].estimators_):
                tr = est.tree_
                depth = np.zeros(tr.node_count, dtype=np.int64)
                for node in range(tr.node_count):           # sklearn ids are pre-order: parents come first
                    l, r = tr.children_left[node], tr.children_right[node]
                    if l != r:
                        depth[l] = depth[node] + 1
                        depth[r] = depth[node] + 1
                total += float(depth[lv[t]].mean()) * per_visit + 8
        return total

    def step(self, record):
        import torch
        from hypermapper_b200 import forest
        ev = None
        if record:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if self.fused_topk:
            # hm_rf_score_topk: prefix pass + filtered pass + merge of the short list; no score array (what local_search
            # needs from its random pools); the events bracket the whole call
            tv, ti = forest.score_topk(self.forests, self.codes_dev if self.packed else self.x_dev, self.acq, self.k)
            if record:
                ev[1].record()
            return (tv, ti + self.rank * self.N), ev
        if self.packed:
            out = forest.rf_eval_packed(self.forests, self.codes_dev, acq=self.acq)
        else:
            out = forest.rf_eval(self.forests, self.x_dev, acq=self.acq)
        if record:
            ev[1].record()
        tv, ti = forest.topk(out["score"], self.k, index_base=self.rank * self.N)
        return (tv, ti), ev

    def e2e_step(self):
        import torch
        if self.packed:
            _, tv, ti = self.scorer.score_packed(self.forests, self.codes_host, self.acq, k=self.k,
                                                 scores_out=self.scores_host)
        else:
            _, tv, ti = self.scorer.score(self.forests, self.x_host, self.acq, k=self.k, scores_out=self.scores_host)
        return torch.from_numpy(tv).cuda(), torch.from_numpy(ti + self.rank * self.N).cuda()

    def e2e_bytes(self, N):
        return (8 if self.packed else 4 * self.wl.D_enc) * N, 8 * N + 16 * self.k

    def roofline(self, N, kernel_ms):
        peak, src = measured_peaks()
        b = 4 * self.wl.D_enc + 8
        ach = b * N / (kernel_ms / 1e3) / 1e9
        # secondary, physical roofline: the shared-memory / L1 data pipe moves 128 B per clock per SM
        import torch
        prop = torch.cuda.get_device_properties(0)
        smem_peak = 128.0 * prop.multi_processor_count * 1.965e9 / 1e9      # GB/s at clocks.max.sm = 1965 MHz
        smem_ach = self.smem_bytes_per_eval * N / (kernel_ms / 1e3) / 1e9
        traffic, traffic_src = NCU_TRAFFIC.get((self.name, N), (None, None))
        compact = None
        if self.packed:
            compact = {"bytes_per_eval": 16, "achieved": 16 * N / (kernel_ms / 1e3) / 1e9,
                       "how": "what this path actually has to move: 8 B packed code in + 8 B fp64 score out (SURVEY 8(d): "
                              "secondary figure; `achieved` above stays on the contract's 4*D_enc + 8)"}
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "compact": compact,
                "traffic_source": traffic_src, "ncu_onchip": NCU_ONCHIP.get((self.name, N)),
                "smem_pipe": {"achieved": smem_ach, "peak": smem_peak, "unit": "GB/s", "frac": smem_ach / smem_peak,
                              "bytes_per_eval": self.smem_bytes_per_eval,
                              "how": ("per tree: 8 B per internal node on the path (the node; the candidate is a register "
                                      "pair) + 8 B leaf node" if self.packed else
                                      "per tree: 12 B per internal node on the path (8 B node + 4 B feature) + 8 B leaf "
                                      "node") + ", path lengths measured on 65536 pool candidates; peak = 128 B/clk/SM x "
                                     "SMs x 1965 MHz"},
                "kernel": ("hm_forest_kernel<packed>" if self.packed else "hm_forest_kernel<fp32>") + (" (+ coherence pre-pass hm_forest_sortkey/scan/scatter_rows, inside "
                                                       "kernel_ms)" if getattr(self, "prepass", False) else "") +
                          (" -- fused mode: kernel_ms = the whole hm_rf_score_topk call (65536-candidate prefix pass + hm_topk, "
                           "filtered pass over the pool, hm_topk_merge of the list)" if self.fused_topk else ""),
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_eval": b, "peak_source": src,
                "note": "forests with thousands of node visits per evaluation are bound by shared-memory load latency "
                        "and the L1/shared-memory data pipe, not HBM (profiles/, DESIGN.md section 3)"}


class SpaceRFBench(RFBench):
    """C3 with the search-space kernels in front of the forest kernel (csrc/hm_space.cu):
      c3raw  : inputs are the RAW parameter values [N][D] fp64 (what the reference API takes); step = hm_encode ->
               hm_rf_eval -> hm_topk; e2e = pinned raw host array -> H2D -> the same -> D2H scores + top-k
      c3pool : the pool is drawn on the device; step = hm_sample_pool (generate + encode) -> hm_rf_eval -> hm_topk;
               nothing but the top-k crosses PCIe, so e2e == the device path + the D2H of the top-k and their rows
    roofline.space_kernel reports hm_space_kernel alone against the HBM peak (it is the HBM-bound part of the step)."""
    cpu_desc = ("numpy restatement of models.py:941-984 (encode) + OpenMP C port of models.py:174-275 + numpy/scipy EI "
                "+ top-20")

    def __init__(self, name):
        RFBench.__init__(self, name)
        self.desc = {"c3raw": "C3 from raw parameter values: device encode (D=10 -> D_enc=23) + 256-tree RF x2, "
                              "EI+tchebyshev, top-20",
                     "c3pool": "C3 with device-generated pools: Philox sample+encode (D=10 -> D_enc=23) + 256-tree RF "
                               "x2, EI+tchebyshev, top-20"}[name]
        self.e2e_api = {"c3raw": "hm_rf_score_raw_host (C-ABI, pinned raw host values, 2M-candidate double-buffered chunks: "
                                 "H2D -> hm_encode -> hm_rf_eval -> D2H scores; hm_topk -> D2H top-k)",
                        "c3pool": "DevicePool + acquisition_scores_device: hm_sample_pool -> hm_rf_eval -> hm_topk -> "
                                  "D2H top-k + hm_sample_rows of the winners (no candidate crosses PCIe)"}[name]

    def raw_rows(self, n, seed):
        from bench_support import workloads
        wl = self.wl
        rng = np.random.default_rng(seed)
        inputs = wl.space.get_input_parameters()
        out = np.empty((n, len(inputs)), dtype=np.float64)
        chunk = 1 << 21
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            cols = workloads.raw_uniform_columns(wl.space, hi - lo, rng)
            for j, p in enumerate(inputs):
                out[lo:hi, j] = cols[p]
        return out

    def cpu_port(self, sample):
        from oracle import oracle
        wl = self.wl
        if not hasattr(self, "_fas"):
            self._fas = {o: oracle.forest_arrays_from_model(m) for o, m in wl.models.items()}
            self._raw_cpu = {}
        if sample not in self._raw_cpu:
            self._raw_cpu = {sample: self.raw_rows(sample, 1234)}
        raw = self._raw_cpu[sample]
        t0 = time.perf_counter()
        X = np.ascontiguousarray(wl.encoder.encode(raw).T)
        means, variances = {}, {}
        for o, fa in self._fas.items():
            means[o], variances[o], _ = oracle.rf_mean_var_omp(fa, X)
        vals = oracle.EI(means, variances, wl.data_array, wl.objective_weights, wl.scalarization, wl.objective_limits)
        top = oracle.get_min_indices_fast(vals, self.k)
        dt = time.perf_counter() - t0
        self._Xcpu = {sample: X.astype(np.float32)}
        self._cpu_out = (sample, means, vals, top)
        return dt

    def setup_device(self, N, rank):
        import torch
        from hypermapper_b200 import forest
        from hypermapper_b200.space_device import device_space_for
        wl = self.wl
        self.N, self.rank = N, rank
        self.forests = [forest.device_forest(wl.models[o]) for o in wl.objectives]
        self.acq = wl.acq_params()
        self.ds = device_space_for(wl.space)
        self.D = self.ds.D
        self.seed = 6 + rank
        if self.name == "c3raw":
            self.raw_host = torch.from_numpy(self.raw_rows(N, 6 + rank)).pin_memory()
            self.raw_dev = self.raw_host.cuda()
            from hypermapper_b200.scoring import HostScorer
            self.scorer = HostScorer(wl.D_enc, chunk=1 << 21)
            self.x_dev = self.ds.encode(self.raw_dev)[0]
        else:
            self.x_dev = self.ds.sample_pool(self.seed, N, False)["x32"]
        self.scores_host = torch.empty((N,), dtype=torch.float64).pin_memory()
        self.prepass = sum(f.n_groups for f in self.forests if f.staged) > 2 and N >= (1 << 17)
        self.smem_bytes_per_eval = self._smem_bytes_per_eval()
        self.launches_per_step = 4 + (3 if self.prepass else 0)
        self.space_ms = []

    def _front(self, raw_dev):
        if self.name == "c3raw":
            return self.ds.encode(raw_dev)[0]
        return self.ds.sample_pool(self.seed, self.N, False)["x32"]

    def step(self, record):
        import torch
        from hypermapper_b200 import forest
        ev = None
        e0 = e1 = None
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            e0.record()
        x32 = self._front(getattr(self, "raw_dev", None))
        if record:
            e1.record()
            self.space_ms.append((e0, e1))
            ev[0].record()
        out = forest.rf_eval(self.forests, x32, acq=self.acq)
        if record:
            ev[1].record()
        tv, ti = forest.topk(out["score"], self.k, index_base=self.rank * self.N)
        return (tv, ti), ev

    def e2e_step(self):
        import torch
        from hypermapper_b200 import forest
        if self.name == "c3raw":
            _, tv, ti = self.scorer.score_raw(self.ds, self.forests, self.raw_host, self.acq, k=self.k,
                                              scores_out=self.scores_host)
            return torch.from_numpy(tv).cuda(), torch.from_numpy(ti + self.rank * self.N).cuda()
        x32 = self._front(None)
        out = forest.rf_eval(self.forests, x32, acq=self.acq)
        tv, ti = forest.topk(out["score"], self.k, index_base=self.rank * self.N)
        self.winner_rows = self.ds.sample_rows(self.seed, False, ti - self.rank * self.N)
        return tv, ti

    def e2e_bytes(self, N):
        if self.name == "c3raw":
            return 8 * self.D * N, 8 * N + 16 * self.k
        return 0, 16 * self.k + 8 * self.D * self.k

    def roofline(self, N, kernel_ms):
        roof = RFBench.roofline(self, N, kernel_ms)
        peak, src = measured_peaks()
        ms = statistics.mean(a.elapsed_time(b) for a, b in self.space_ms)
        wl = self.wl
        b = (8 * self.D if self.name == "c3raw" else 0) + 4 * wl.D_enc
        roof["space_kernel"] = {"kernel": "hm_space_kernel<%s>" % ("false" if self.name == "c3raw" else "true"),
                                "bound": "hbm", "kernel_ms": ms, "algorithmic_bytes_per_candidate": b,
                                "achieved": b * N / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": b * N / (ms / 1e3) / 1e9 / peak,
                                "how": "raw fp64 values read (c3raw) + fp32 encoded columns written, per candidate"}
        return roof


class GPBench:
    metric, unit, dtype = "acquisition_evals_per_sec", "evals/s", "f64"
    cpu_desc = "numpy/scipy (BLAS) restatement of GPy predict, not GPy itself + EI + top-20"
    e2e_api = ("hm_gp_score_host (C-ABI, pinned host candidates, 512k-candidate double-buffered chunks: H2D -> "
               "hm_gp_mean_var -> hm_acq_score -> D2H scores; hm_topk -> D2H top-k)")

    def __init__(self, name):
        from hypermapper_b200 import gp
        from bench_support import workloads
        self.name = name
        if name == "c2":
            n, D, self.n_default = 256, 6, 1 << 20
            X = np.random.default_rng(2).random((n, D))
            y = workloads.hartmann6(X)
            scale = 1.0
            self.desc = "C2: Hartmann-6, GP Matern-5/2 ARD (n=256), EI over a 1M-candidate pool, top-20"
        else:
            n, D, self.n_default = 512, 32, 1 << 22
            X = np.random.default_rng(12).random((n, D))
            y = np.sin(X[:, :8].sum(1)) + (X[:, 8:] ** 2).sum(1) * 0.1
            scale = 3.0
            self.desc = "C5-GP: 20-D mixed space encoded to D_enc=32, GP Matern-5/2 ARD (n=512), EI, top-20"
        self.n, self.D = n, D
        self.X, self.y = X, y
        self.ls = np.random.default_rng(3).uniform(0.3, 1.0, D) * scale
        self.state = gp.GPState.from_hyperparameters(X, y, self.ls, 1.0, 1e-4)
        self.k = 20
        self.data = {"Value": [float(v) for v in y]}
        self.weights = {"Value": 1.0}
        self.limits = {"Value": [float(y.min()), float(y.max())]}
        self.cpu_sample_default = 200000 if name == "c2" else 50000

    def small(self, N):
        return 8 * self.D * N < 200e6

    def config(self, N, world):
        return {"workload": self.desc, "candidates_per_gpu": N, "D_enc": self.D, "n_train": self.n, "k": self.k,
                "parallelism": "candidate-sharded x%d" % world,
                "l2": "flushed between timed steps (256 MB write)" if self.small(N) else "inputs larger than L2"}

    def candidates(self, N, seed):
        return np.ascontiguousarray(np.random.default_rng(seed).random((self.D, N)))

    def cpu_port(self, sample):
        """numpy/scipy restatement of GPy's exact-inference predict (BLAS, all threads) + EI + top-k"""
        from oracle import oracle
        if not hasattr(self, "_st"):
            self._st = oracle.gp_fit(self.X, self.y, self.ls, 1.0, 1e-4)
            self._Xs = np.ascontiguousarray(self.candidates(sample, 1234).T)
        t0 = time.perf_counter()
        mean, var = oracle.gp_predict(self._st, self._Xs)
        vals = oracle.EI({"Value": mean}, {"Value": var}, self.data, self.weights, "tchebyshev", self.limits)
        top = oracle.get_min_indices_fast(vals, self.k)
        dt = time.perf_counter() - t0
        self._cpu_out = (sample, mean, var, vals, top)
        return dt

    def parity(self):
        """GPU posterior mean / variance / EI on the CPU leg's sample against the oracle (pinned to scikit-learn's GPR,
        tests/golden/gp_sklearn.npz): 1e-5 relative (mean: relative to max(|mean|, 1e-3 std(y)); variance: plus the
        rounding floor of the cancellation sf2 - |L^-1 k|^2), top-k identical."""
        import torch
        from hypermapper_b200 import forest, gp, models
        sample, mean_o, var_o, vals, top = self._cpu_out
        X64 = torch.from_numpy(np.ascontiguousarray(self._Xs.T)).cuda()
        mean = torch.empty((1, sample), dtype=torch.float64, device="cuda")
        var = torch.empty_like(mean)
        gp.gp_mean_var(self.dgp, X64, mean[0], var[0])
        sc = models.acq_score_device(self.acq, mean, var)
        m, v, s = mean[0].cpu().numpy(), var[0].cpu().numpy(), sc.cpu().numpy()
        ystd = float(np.std(self.y))
        em = float(np.max(np.abs(m - mean_o) / np.maximum(np.abs(mean_o), 1e-3 * ystd)))
        floor = 64 * self.n * np.finfo(np.float64).eps * self.state.variance * self.state.y_std ** 2
        ev = float(np.max(np.maximum(np.abs(v - var_o) - floor, 0) / var_o))
        with np.errstate(all="ignore"):
            es = np.abs(s - vals) / np.maximum(np.abs(vals), 1e-300)
        es[s == vals] = 0
        got = forest.topk(sc, self.k)[1].cpu().numpy()
        res = {"sample": sample, "mean_max_rel_err": em, "var_max_rel_err": ev, "score_max_rel_err": float(es.max()),
               "tol": 1e-5, "topk_identical": bool(np.array_equal(got, top))}
        res["ok"] = bool(em <= 1e-5 and ev <= 1e-5 and es.max() <= 1e-5 and res["topk_identical"])
        return res

    def setup_device(self, N, rank):
        import torch
        from hypermapper_b200 import gp, _lib
        from hypermapper_b200.random_scalarizations import acquisition_constants
        self.N, self.rank = N, rank
        self.dgp = gp.device_gp(self.state)
        w, fmin, beta = acquisition_constants("EI", "tchebyshev", ["Value"], self.weights, self.limits, self.data, 5)
        self.acq = _lib.make_acq_params("EI", "tchebyshev", w, fmin, beta)
        self.x_host = torch.from_numpy(self.candidates(N, 6 + rank)).pin_memory()
        self.x_dev = self.x_host.cuda()
        self.x_stage = torch.empty_like(self.x_dev)
        self.mean = torch.empty((1, N), dtype=torch.float64, device="cuda")
        self.var = torch.empty_like(self.mean)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if self.small(N) else None
        self.scores_host = torch.empty((N,), dtype=torch.float64).pin_memory()
        self.launches_per_step = 4
        self.fp64_peak = float(_lib.lib().hm_fp64_peak_tflops())

    def step(self, record):
        import torch
        from hypermapper_b200 import forest, gp, models
        if self.flush is not None:
            self.flush.fill_(1)                    # evict the (small) candidate matrix from L2 between steps
        ev = None
        if record:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        gp.gp_mean_var(self.dgp, self.x_dev, self.mean[0], self.var[0])
        if record:
            ev[1].record()
        scores = models.acq_score_device(self.acq, self.mean, self.var)
        tv, ti = forest.topk(scores, self.k, index_base=self.rank * self.N)
        return (tv, ti), ev

    def e2e_step(self):
        import torch
        if not hasattr(self, "scorer"):
            from hypermapper_b200.scoring import HostScorer
            self.scorer = HostScorer(self.D, chunk=1 << 19)
        _, tv, ti = self.scorer.score_gp([self.dgp], self.x_host, self.acq, k=self.k, scores_out=self.scores_host)
        return torch.from_numpy(tv).cuda(), torch.from_numpy(ti + self.rank * self.N).cuda()

    def e2e_bytes(self, N):
        return 8 * self.D * N, 8 * N + 16 * self.k

    def roofline(self, N, kernel_ms):
        n_pad = (self.n + 15) // 16 * 16
        # cross-covariance (2D-flop dot + ~40 for the Matern transform) + lower-triangular contraction (n^2 flops) + squares
        flops = n_pad * (2 * self.D + 40) + n_pad * n_pad + 17 * n_pad
        ach = flops * N / (kernel_ms / 1e3) / 1e12
        peak = self.fp64_peak if self.fp64_peak > 0 else 37.0
        return {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "kernel": "hm_gp_kernel", "kernel_ms": kernel_ms, "algorithmic_flops_per_eval": flops,
                "peak_source": "FP64 DFMA burst measured in-process (hm_fp64_peak_tflops)" if self.fp64_peak > 0
                else "datasheet 37 TFLOP/s",
                "note": "tcgen05 has no f64 kind and tf32/3xtf32 miss the 1e-5 tolerance (DESIGN.md section 3, K2): the "
                        "contraction runs on the FP64 pipe"}


class ParetoBench:
    metric, unit, dtype = "pareto_points_per_sec", "points/s", "f64"
    cpu_desc = "single-thread C port of compute_pareto.py:89-117 (the sweep is sequential by definition)"
    e2e_api = "compute_pareto path: pinned host costs -> H2D -> hm_pareto_filter -> D2H mask"

    def __init__(self, name):
        self.name = name
        self.M = 3
        self.n_default = 10_000_000
        self.desc = "C4: compute_pareto dominance filter, 10M uniform 3-objective points (both passes)"
        self.cpu_sample_default = 10_000_000
        self.k = 0

    def config(self, N, world):
        return {"workload": self.desc, "points_per_gpu": N, "objectives": self.M,
                "parallelism": ("single GPU" if world == 1 else
                                "row blocks x%d: local pass 1, all-gather of the survivors, full filter on the union "
                                "(distributed.pareto_mask_sharded)" % world),
                "l2": "inputs (%.2f GB) larger than L2" % (8 * self.M * N / 1e9)}

    def cpu_port(self, sample):
        from oracle import oracle
        c = np.random.default_rng(7).random((sample, self.M))
        t0 = time.perf_counter()
        mask = oracle.pareto(c)
        dt = time.perf_counter() - t0
        self._cpu_out = (sample, c, mask)
        return dt

    def parity(self):
        """the GPU mask on the CPU leg's points == the oracle's mask, bit for bit"""
        import torch
        from hypermapper_b200 import compute_pareto
        sample, c, mask = self._cpu_out
        got = compute_pareto.pareto_mask_device(torch.from_numpy(c).cuda()).cpu().numpy().astype(bool)
        same = bool(np.array_equal(got, np.asarray(mask).astype(bool)))
        return {"sample": sample, "mask_bit_exact": same, "front_size": int(got.sum()), "ok": same}

    def setup_device(self, N, rank):
        import torch
        import torch.distributed as dist
        self.N, self.rank = N, rank
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        c = np.random.default_rng(7 + rank).random((N, self.M))
        self.c_host = torch.from_numpy(c).pin_memory()
        self.c_dev = self.c_host.cuda()
        self.c_stage = torch.empty_like(self.c_dev)
        self.mask_host = torch.empty((N,), dtype=torch.uint8).pin_memory()
        # fast path (one CUDA graph): ctl_init, 2 x (prep, stream), exact, tail_small, sort_front, pass2_a, pass2_b, publish
        # sharded: pass 1 on the block (ctl_init, 2 x (prep, stream), exact, publish = 7) + sort_front + gather_rows + the
        # full filter on the union (11)
        self.launches_per_step = 11 if self.world == 1 else 20

    def step(self, record):
        import torch
        from hypermapper_b200 import compute_pareto
        ev = None
        if record:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if self.world > 1:
            from hypermapper_b200.distributed import pareto_mask_sharded
            mask = pareto_mask_sharded(self.c_dev)
        else:
            mask = compute_pareto.pareto_mask_device(self.c_dev)
        if record:
            ev[1].record()
        return (mask, None), ev

    def e2e_step(self):
        from hypermapper_b200 import compute_pareto
        self.c_stage.copy_(self.c_host, non_blocking=True)
        if self.world > 1:
            from hypermapper_b200.distributed import pareto_mask_sharded
            mask = pareto_mask_sharded(self.c_stage)
        else:
            mask = compute_pareto.pareto_mask_device(self.c_stage)
        self.mask_host.copy_(mask, non_blocking=True)
        return mask, None

    def e2e_bytes(self, N):
        return 8 * self.M * N, N

    def roofline(self, N, kernel_ms):
        peak, src = measured_peaks()
        b = 8 * self.M + 1
        ach = b * N / (kernel_ms / 1e3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "kernel": "hm_px_* (whole filter call, one CUDA graph: 2 x {pivot preparation, stream filter}, exact all-pairs, "
                          "one-CTA pass 2, publish; the level-1 hm_px_stream launch is the only one that touches all N points)",
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_point": b, "peak_source": src}


def make_bench(name):
    if name in ("c3raw", "c3pool"):
        return SpaceRFBench(name)
    if name in ("c3", "c3f32", "c1", "c5rf"):
        return RFBench(name)
    if name in ("c2", "c5gp"):
        return GPBench(name)
    return ParetoBench(name)


def cpu_threads(workload):
    from oracle import oracle
    n = oracle.use_all_host_threads()
    return 1 if workload == "c4" else n


def run_reference(args, rank):
    """--impl reference: the CPU port alone, rank 0 only, bounded sample per step."""
    if rank != 0:
        return
    b = make_bench(args.workload)
    threads = cpu_threads(args.workload)
    sample = args.cpu_sample or b.cpu_sample_default
    times = []
    for s in range(args.warmup + args.steps):
        dt = b.cpu_port(sample)
        if s >= args.warmup:
            times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = sample / (ms / 1000.0)
    line = {
        "metric": b.metric, "value": value, "unit": b.unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": b.dtype, "data": "synthetic", "impl": "reference",
        "config": {"workload": b.desc, "units_per_step": sample, "note": "bounded sample of the same workload"},
        "cpu_baseline": {"value": value, "unit": b.unit, "cores": threads, "kind": "port",
                         "sample": "%d units per step; %s" % (sample, b.cpu_desc)},
        "e2e": {"value": value, "unit": b.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from hypermapper_b200.distributed import allgather_topk

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    b = make_bench(args.workload)
    if args.fused_topk:
        assert hasattr(b, "fused_topk"), "--fused-topk applies to the forest workloads"
        b.fused_topk = True
        b.desc += "; fused scoring + selection (hm_rf_score_topk: no score array)"
    N = args.candidates or b.n_default
    strong = args.total_candidates > 0
    if strong:
        assert args.total_candidates % world == 0, "--total-candidates must be a multiple of the number of GPUs"
        N = args.total_candidates // world
    b.setup_device(N, rank)
    ev_pairs = []

    def full_step(record):
        (tv, ti), ev = b.step(record)
        if ev is not None:
            ev_pairs.append(ev)
        if world > 1 and ti is not None:
            tv, ti = allgather_topk(tv, ti, b.k)
        return tv, ti

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        res = full_step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        res = full_step(True)
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = t_start.elapsed_time(t_end)
    kernel_ms = statistics.mean(a.elapsed_time(c) for a, c in ev_pairs)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * N / (ms_per_step / 1000.0)
    launches = b.launches_per_step + (2 if (world > 1 and b.k) else 0)

    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            b.e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            tv, ti = b.e2e_step()
            if world > 1 and ti is not None:
                tv, ti = allgather_topk(tv, ti, b.k)
                ti.cpu()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        td = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        dt = float(td.item())
        h2d, d2h = b.e2e_bytes(N)
        e2e = {"value": world * N * args.steps / dt, "unit": b.unit, "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": 1000.0 * dt / args.steps, "api": b.e2e_api}
        if world == 1 and ti is not None:
            assert torch.equal(ti.cpu(), res[1].cpu()), "e2e and device-resident top-k differ"

    parity_failed = False
    if rank == 0:
        roof = b.roofline(N, kernel_ms)
        roof["share_of_step"] = kernel_ms / ms_per_step
        line = {
            "metric": b.metric, "value": value, "unit": b.unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": b.dtype, "data": "synthetic", "impl": "b200", "config": b.config(N, world),
            "roofline": roof, "clocks": clocks, "gpu_launches": launches * args.steps,
        }
        if strong:
            line["config"]["total_candidates"] = args.total_candidates
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            threads = cpu_threads(args.workload)
            sample = args.cpu_sample or b.cpu_sample_default
            dt = b.cpu_port(sample)
            line["cpu_baseline"] = {"value": sample / dt, "unit": b.unit, "cores": threads, "kind": "port",
                                    "sample": "%d units of the same workload (%.1f s); %s" % (sample, dt, b.cpu_desc)}
            # parity gate: no throughput number counts unless the GPU path reproduces the oracle
            # on the sample the CPU leg just scored
            line["parity"] = b.parity()
            parity_failed = not line["parity"]["ok"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if parity_failed:
        sys.stderr.write("bench.py: PARITY GATE FAILED: %s\n" % json.dumps(line["parity"]))
        sys.exit(1)


if __name__ == "__main__":
    main()
