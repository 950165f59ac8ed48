#!/usr/bin/env python
"""K2 precision study: what the predictive-variance contraction  T = Kx . L^-T  loses when its
operands go through the tcgen05 tensor path (tf32 / bf16 inputs, fp32 accumulation in TMEM) instead of the FP64
tensor path (mma.sync.m8n8k4.f64, what csrc/hm_gp.cu runs).  CPU only (numpy emulation of the operand roundings and of
fp32 accumulation), on the bench's own C2 model (Hartmann-6, n = 256, Matern-5/2 ARD) and the C5-GP shape (n = 512,
D = 32), with and without observation noise.

    python tools/gp_precision_probe.py            # prints the table committed as profiles/r02_gp_precision_probe.md

For each scheme: max relative error of the de-normalised predictive variance against the fp64 oracle over 20 000
candidates (the north-star tolerance is 1e-5), the number of tensor-core GEMM passes it needs, and a time projection
for the contraction alone  =  passes x 2 n^2 / (tensor rate)  against  n^2 / (measured DMMA rate).

Schemes
  tf32            operands rounded to tf32 (10 mantissa bits), fp32 accumulation               1 pass
  3xtf32          hi/lo tf32 split of both operands, the three large products                  3 passes
  bf16 x k        greedy k-slice bf16 split (each slice = bf16(remainder)), products i+j<=k+1   k(k+1)/2 passes
  ozaki bf16 x k  Ozaki scheme: per-row common exponent, k slices of 8 bits, every slice product and every fp32
                  accumulation EXACT (8+8 bits x n <= 2^8 terms <= 24 bits), recombined in fp64    k(k+1)/2 passes
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402  (test infrastructure: the fp64 reference of this study)

DMMA_TFLOPS = 22.0        # achieved by hm_gp_kernel on C2 (profiles/r01_gp_c2_summary.md)
BF16_TFLOPS = 1656.5      # MEASURED_PEAKS.json, dense bf16 burst
TF32_TFLOPS = BF16_TFLOPS / 2


def round_mantissa(x, bits):
    """round fp32 values to `bits` explicit mantissa bits (round to nearest even): tf32 = 10, bf16 = 7"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    half = (1 << (drop - 1)) - 1
    u = u + half + ((u >> drop) & 1)
    u = (u >> drop) << drop
    return u.astype(np.uint32).view(np.float32)


def greedy_slices(x64, bits, k):
    out, rem = [], np.array(x64, dtype=np.float64)
    for _ in range(k):
        s = round_mantissa(rem.astype(np.float32), bits).astype(np.float64)
        out.append(s)
        rem = rem - s
    return out


def ozaki_slices(x64, k, slice_bits=8):
    """rows share one exponent; slice i holds bits [i*slice_bits, (i+1)*slice_bits) below it as small integers scaled
    by a power of two -> every slice is exactly representable in bf16 and slice products accumulate exactly in fp32"""
    x = np.array(x64, dtype=np.float64)
    e = np.ceil(np.log2(np.maximum(np.abs(x).max(axis=1, keepdims=True), 1e-300)))
    out, rem = [], x.copy()
    for i in range(k):
        scale = np.exp2(e - (i + 1) * slice_bits + 1)
        s = np.trunc(rem / scale) * scale
        out.append(s)
        rem = rem - s
    return out


def fp32_gemm(a, b):
    """a [N][n] . b [n][m] with fp32 operands and fp32 accumulation (what a tensor-core pass with an fp32 TMEM
    accumulator returns)"""
    return (a.astype(np.float32) @ b.astype(np.float32)).astype(np.float64)


def contraction(Kx, LinvT, scheme, k=0):
    if scheme == "fp64":
        return Kx @ LinvT, 0
    if scheme == "tf32":
        return fp32_gemm(round_mantissa(Kx.astype(np.float32), 10), round_mantissa(LinvT.astype(np.float32), 10)), 1
    if scheme == "3xtf32":
        a = greedy_slices(Kx, 10, 2)
        b = greedy_slices(LinvT, 10, 2)
        return fp32_gemm(a[0], b[0]) + fp32_gemm(a[0], b[1]) + fp32_gemm(a[1], b[0]), 3
    if scheme in ("bf16", "ozaki"):
        a = greedy_slices(Kx, 7, k) if scheme == "bf16" else ozaki_slices(Kx, k)
        b = greedy_slices(LinvT.T, 7, k) if scheme == "bf16" else ozaki_slices(LinvT.T, k)
        total, passes = 0.0, 0
        for i in range(k):
            for j in range(k - i):
                total = total + fp32_gemm(a[i], b[j].T)
                passes += 1
        return total, passes
    raise ValueError(scheme)


def model(n, D, noise, seed, scale):
    from bench_support import workloads
    rng = np.random.default_rng(seed)
    X = rng.random((n, D))
    y = workloads.hartmann6(X) if D == 6 else np.sin(X[:, :8].sum(1)) + (X[:, 8:] ** 2).sum(1) * 0.1
    ls = np.random.default_rng(seed + 1).uniform(0.3, 1.0, D) * scale
    return oracle.gp_fit(X, y, ls, 1.0, noise), X


def study(tag, n, D, noise, seed, scale, N=20000):
    st, X = model(n, D, noise, seed, scale)
    Xs = np.random.default_rng(4).random((N, D))
    Xs[:5] = X[:5]                                   # candidates on training points: the hard end of the cancellation
    Kx = oracle.matern52(st.X, Xs, st.lengthscale, st.variance).T          # [N][n]
    LinvT = np.linalg.inv(st.L).T
    rows = []
    ref = None
    schemes = [("fp64", 0), ("tf32", 0), ("3xtf32", 0)] + [("bf16", k) for k in (2, 3, 4, 5, 6)] + \
              [("ozaki", k) for k in (4, 5, 6, 7)]
    for scheme, k in schemes:
        T, passes = contraction(Kx, LinvT, scheme, k)
        var = np.clip(st.variance - np.sum(T * T, axis=1), 1e-15, np.inf) + st.noise
        var = np.maximum(var * st.y_std ** 2, 1e-11)
        if ref is None:
            ref = var
        err = np.abs(var - ref) / ref
        t_ratio = None if not passes else passes * 2.0 / ((TF32_TFLOPS if "tf32" in scheme else BF16_TFLOPS) / DMMA_TFLOPS)
        rows.append((scheme + (" x %d" % k if k else ""), passes, float(err.max()), float(np.median(err)), t_ratio))
    print("\n### %s: n = %d, D = %d, noise = %g, %d candidates (5 of them on training points)\n" % (tag, n, D, noise, N))
    print("| scheme | tensor passes | max rel. variance error | median | meets 1e-5 | contraction time vs DMMA (projected) |")
    print("|---|---|---|---|---|---|")
    for name, passes, emax, emed, t in rows:
        print("| %s | %s | %.2e | %.2e | %s | %s |" % (name, passes or "-", emax, emed, "yes" if emax <= 1e-5 else "NO",
                                                      "-" if t is None else "%.2f" % t))


if __name__ == "__main__":
    print("# K2 precision probe: the variance contraction on the tcgen05 path (numpy emulation)\n")
    print("DMMA rate %.1f TFLOP/s (achieved, C2), bf16 tensor %.1f TFLOP/s (MEASURED_PEAKS.json), tf32 = half of it; the "
          "projection counts the contraction only (no operand splitting, no TMEM read-back, no fp64 recombination) and a "
          "tensor pass costs 2 n^2 flops per candidate against n^2 for the triangular DMMA product (the tensor path could "
          "skip the zero half too; not assumed)." % (DMMA_TFLOPS, BF16_TFLOPS))
    study("C2", 256, 6, 1e-4, 2, 1.0)
    study("C2, noise fixed to 0 (models.py:451-453)", 256, 6, 0.0, 2, 1.0)
    study("C5-GP", 512, 32, 1e-4, 12, 3.0)
