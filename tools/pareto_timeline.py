#!/usr/bin/env python
"""In-graph timeline of one hm_pareto_filter call (CUPTI through torch.profiler, the kernels are launched by a CUDA graph):
per kernel start offset, duration and the gap to the previous kernel.   usage: python tools/pareto_timeline.py [N] [M]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    from torch.profiler import profile, ProfilerActivity
    from hypermapper_b200 import compute_pareto
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    c = torch.from_numpy(np.random.default_rng(7).random((N, M))).cuda()
    for _ in range(5):
        compute_pareto.pareto_mask_device(c)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            compute_pareto.pareto_mask_device(c)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # the last call = the events after the last hm_px_ctl_init / first kernel of the chain
    starts = [i for i, e in enumerate(evs) if "ctl_init" in e.name or "hm_px_head" in e.name]
    first = starts[-1] if starts else 0
    evs = evs[first:]
    t0 = evs[0].time_range.start
    prev_end = t0
    total_k = 0.0
    print("%-56s %9s %9s %9s" % ("kernel", "start us", "dur us", "gap us"))
    for e in evs:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        print("%-56s %9.2f %9.2f %9.2f" % (e.name[:56], s, d, e.time_range.start - prev_end))
        prev_end = e.time_range.end
        total_k += d
    print("span %.2f us, kernel time %.2f us, gaps %.2f us" % (prev_end - t0, total_k, prev_end - t0 - total_k))


if __name__ == "__main__":
    main()
