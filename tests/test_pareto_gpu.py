"""GPU parity for K4: the Pareto kernels against the reference's golden masks (incl. the known-answer CSVs of the
reference's own tests), the C oracle on seeded inputs, and size-independent properties at 10^7 points."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

CASES = ["uniform3", "uniform2", "uniform4", "ties3", "ties2", "anticorr2", "dups3", "nan3", "single", "quirk3",
         "quirk3_swapped", "blackscholes"]


def golden():
    return np.load(os.path.join(GOLDEN, "pareto.npz"), allow_pickle=False)


@pytest.mark.parametrize("case", CASES)
def test_mask_bit_exact_vs_reference_golden(case):
    from hypermapper_b200 import compute_pareto
    z = golden()
    got = compute_pareto.sequential_is_pareto_efficient(z["costs_" + case])
    assert got.dtype == bool
    assert np.array_equal(got, z["mask_" + case])


def test_reference_known_answer_blackscholes():
    from hypermapper_b200 import compute_pareto
    z = golden()
    costs = z["costs_blackscholes"]
    got = np.array(sorted(costs[compute_pareto.sequential_is_pareto_efficient(costs)].tolist()))
    assert np.array_equal(got, z["kat_blackscholes"])        # tests/data/BlackScholes_real_pareto.csv, 88 rows


@pytest.mark.parametrize("kind", ["branin_grid", "ordinal_branin"])
def test_reference_known_answer_grids(kind):
    """tests/test_system.py:36-64 of the reference: exhaustive 100 x 100 / 16 x 16 Branin grids -> the fronts in
    tests/data/branin_grid_search_pareto.csv (98 rows) / ordinal_branin_real_pareto.csv (12 rows)"""
    import pareto_kat
    from hypermapper_b200 import compute_pareto
    pareto_kat.check(golden(), kind, compute_pareto.sequential_is_pareto_efficient)


def _branin_pair(x1, x2):
    a, b, c, r, s, t = 1.0, 5.1 / (4.0 * np.pi * np.pi), 5.0 / np.pi, 6.0, 10.0, 1.0 / (8.0 * np.pi)
    return a * (x2 - b * x1 * x1 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s


@pytest.mark.parametrize("n,m,kind", [(1, 3, "u"), (2, 2, "u"), (1023, 3, "u"), (1025, 3, "u"), (50000, 3, "u"),
                                      (50000, 2, "ties"), (30000, 5, "u"), (20000, 8, "u"), (40000, 3, "nan"),
                                      (6000, 2, "front"), (200000, 3, "dups"), (3000, 1, "u")])
def test_vs_oracle_seeded(n, m, kind):
    from hypermapper_b200 import compute_pareto
    from oracle import oracle
    rng = np.random.default_rng(n * 10 + m)
    c = rng.random((n, m))
    if kind == "ties":
        c = np.floor(c * 64) / 64
    elif kind == "nan":
        c[rng.integers(0, n, 40), rng.integers(0, m, 40)] = np.nan
        c[5] = np.nan
    elif kind == "front":
        c[:, 1] = 1 - c[:, 0]                 # every point on the front: many rounds, all pivots
    elif kind == "dups":
        c[n // 2:] = c[: n - n // 2]
    want = oracle.pareto(c)
    got = compute_pareto.sequential_is_pareto_efficient(c)
    assert np.array_equal(got, want)
    p1 = compute_pareto.pareto_mask_device(__import__("torch").from_numpy(c).cuda(), pass2=False).cpu().numpy().astype(bool)
    assert np.array_equal(p1, oracle.pareto_pass1(c))


def test_string_input_and_empty():
    from hypermapper_b200 import compute_pareto
    c = np.array([["1.0", "2.0"], ["2.0", "1.0"], ["3.0", "3.0"]])
    assert compute_pareto.sequential_is_pareto_efficient(c).tolist() == [True, True, False]
    assert compute_pareto.sequential_is_pareto_efficient(np.zeros((0, 3))).shape == (0,)


def test_full_size_properties_10M_x3():
    """BASELINE config 4: 10^7 uniform 3-objective points.  Properties: idempotence (the front of the front is the
    front), the front is mutually non-dominated, a random sample of dropped points is dominated by a kept one, and
    the survivor set matches the digest the C oracle produced for this seed in the build container (125 points)."""
    import hashlib
    import torch
    from hypermapper_b200 import compute_pareto
    c = np.random.default_rng(7).random((10_000_000, 3))
    cd = torch.from_numpy(c).cuda()
    mask = compute_pareto.pareto_mask_device(cd).cpu().numpy().astype(bool)
    front = c[mask]
    assert mask.sum() == 125
    assert hashlib.sha256(np.nonzero(mask)[0].tobytes()).hexdigest()[:16] == "b689d6ffe3f04453"
    assert compute_pareto.sequential_is_pareto_efficient(front).all()
    rng = np.random.default_rng(0)
    for i in rng.choice(np.nonzero(~mask)[0], 300, replace=False):
        assert np.any(np.all(front < c[i], axis=1) | ((np.any(front == c[i], axis=1)) & np.any(front < c[i], axis=1)))


def test_parallel_is_pareto_efficient_two_stage():
    from hypermapper_b200 import compute_pareto
    from oracle import oracle
    rng = np.random.default_rng(3)
    c = rng.random((20000, 2))
    preds = {"a": c[:, 0].copy(), "b": c[:, 1].copy()}
    keep2 = compute_pareto.parallel_is_pareto_efficient(False, preds, c, number_of_cpus=4)
    bounds = compute_pareto._chunk_bounds(len(c), 4)
    keep = np.concatenate([oracle.pareto(c[lo:hi]) for lo, hi in bounds])
    assert np.array_equal(preds["a"], c[keep, 0])
    assert np.array_equal(keep2, oracle.pareto(c[keep]))


SHARDED_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %r)
from hypermapper_b200.distributed import pareto_mask_sharded, shard_bounds
from hypermapper_b200.compute_pareto import pareto_mask_device
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)      # both ranks share cuda:0; collectives via the host
rng = np.random.default_rng(5)
cases = {"uniform3": rng.random((400_000, 3)), "ties3": np.floor(rng.random((200_000, 3)) * 64) / 64,
         "uniform2": rng.random((300_001, 2)), "nan3": rng.random((50_001, 3))}
cases["nan3"][[7, 30_000, 49_999], [0, 2, 1]] = np.nan          # NaN costs: the all-rows fallback (exact path on every rank)
for name, c in cases.items():
    whole = pareto_mask_device(torch.from_numpy(c).cuda()).cpu().numpy()
    lo, hi = shard_bounds(len(c), rank, world)
    got = pareto_mask_sharded(torch.from_numpy(np.ascontiguousarray(c[lo:hi])).cuda()).cpu().numpy()
    assert np.array_equal(got, whole[lo:hi]), (name, rank, int(got.sum()), int(whole[lo:hi].sum()))
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_sharded_pareto_two_ranks_equals_single_gpu(tmp_path):
    """row blocks on two ranks (sharing this GPU; gloo carries the survivor exchange), K4 kernels on both
    -> each rank's slice equals the single-GPU mask."""
    import subprocess
    import sys
    from conftest import ROOT
    script = tmp_path / "worker_pareto_gpu.py"
    script.write_text(SHARDED_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29535", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
