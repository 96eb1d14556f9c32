"""The j-split wave planner (csrc/stream.cuh plan_splits) through its host-only probe nb_plan_splits — no GPU needed.

Invariants every plan must satisfy (the kernels rely on them), the plans of the benchmark shapes that were MEASURED
(profiles/r02/splits_shapes_ab.log, int_splits_ab.log), and the round-2 rule for single-split multi-wave grids."""
import ctypes

import pytest

from nbody_cosmological_simulation_b200 import _lib as L


def plan(n_targets, n_chunks, ctas_per_sm, max_splits, targets_per_block=512):
    lib = L.load()
    s, c, b = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib.nb_plan_splits(n_targets, n_chunks, targets_per_block, ctas_per_sm, max_splits, ctypes.byref(s), ctypes.byref(c), ctypes.byref(b))
    assert rc == 0
    return s.value, c.value, b.value


@pytest.mark.parametrize("n_targets", [1, 511, 512, 513, 5000, 131072, 1 << 20, 4194304])
@pytest.mark.parametrize("n_chunks", [1, 2, 7, 40, 513, 4096, 16384])
@pytest.mark.parametrize("ctas_per_sm,max_splits", [(3, 10), (2, 10), (4, 16), (3, 32), (3, 1)])
def test_plan_covers_every_chunk_exactly_once(n_targets, n_chunks, ctas_per_sm, max_splits):
    s, cps, blocks = plan(n_targets, n_chunks, ctas_per_sm, max_splits)
    assert blocks == (n_targets + 511) // 512
    assert 1 <= s <= min(max_splits, n_chunks)
    assert s * cps >= n_chunks                      # the splits cover all chunks ...
    assert (s - 1) * cps < n_chunks                 # ... and none of them is empty


def test_measured_plans_of_the_benchmark_shapes():
    # N = 2^20, D = 3, float32: 2048 target blocks, 444 resident CTAs -> 8 splits = 36.9 waves (measured best within the 256 MiB cap)
    assert plan(1 << 20, 4096, 3, 10)[:2] == (8, 512)
    # the same tick sharded 8 ways: 256 blocks x 26 splits = 14.99 waves
    assert plan(1 << 17, 4096, 3, 32)[0] == 26
    # argument errors
    lib = L.load()
    assert lib.nb_plan_splits(0, 1, 512, 3, 1, None, None, None) != 0
    assert lib.nb_plan_splits(1, 1, 512, 3, 0, None, None, None) != 0


def test_single_split_multi_wave_grid_is_split_further():
    """Fast-lookup kernel at N = 2^20 (2 CTAs/SM): 2048 x s CTAs fill 296 slots equally badly for every s, the wave model alone
    keeps s = 1 (84 ms CTAs, ragged end: 585 ms); the rule takes the largest equally-rated s whose CTAs still stream >= 256
    chunks (measured 572 ms)."""
    s, cps, _ = plan(1 << 20, 4096, 2, 10)
    assert (s, cps) == (8, 512)
    # not triggered when the grid is a single wave, or when the model already splits
    assert plan(20000, 79, 2, 32)[0] > 1
    assert plan(100000, 4096, 2, 1)[0] == 1         # cap of one split is honoured
    # never below 256 chunks per split through this rule
    s, cps, _ = plan(4194304, 16384, 2, 32)
    assert cps >= 256 and s > 1
