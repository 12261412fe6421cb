"""Kernel-logic tests of the tile kernel WITHOUT a GPU: the body of gadget3_b200/csrc/g3_tile.cuh compiled by g++
(tests/emu/g3emu.cpp), one OS thread per warp with a real barrier, on the tables the product's planner builds,
against the CPU oracle. This checks barrier placement, table layouts, the column -> warp assignment and the index
arithmetic; the device arithmetic itself (exp, reciprocal, FMA contraction) is what the `-m gpu` tests pin.
The harness is test infrastructure: the product library never loads it."""
import numpy as np
import pytest

from gadget3_b200.configs import CONFIGS, parameter_batch
from oracle.oracle_lib import Oracle
from _util import US_WEIGHT, us_tolerance

emu = pytest.importorskip("emu.emu_lib")


def _setup(name, B, seed=2026):
    model, p = CONFIGS[name]()
    ints, reals = model.pack(p)
    opt = p.optimised_indices()
    full = np.tile(p.values(), (B, 1))
    full[:, opt] = parameter_batch(p, B, seed=seed)
    return model, p, ints, reals, full, opt


@pytest.mark.parametrize("smem", [True, False], ids=["state_in_smem", "state_in_slab"])
@pytest.mark.parametrize("name", ["C1", "C1_MULTI", "C2", "C2_MATAGE", "C2_STRAT", "C4", "C5", "LING", "MINI_ANDERSEN", "MINI_OTHERFOOD", "MINI_PREFS", "MINI_SUITS", "MINI_FLEETS", "MINI_MIGRATE", "MINI_TRANSFORM", "SPAWN_RIG", "MINI_SPAWN", "MINI_TIMEVAR", "TIMEVAR_RIG", "RANDOM_RIG", "C1_RANDOM", "SIX_FLEETS", "MINI_MATVAR", "RENEWAL_RIG", "WLOSS_RIG", "MINI_STOMACH"])
def test_values_and_components(oracle_lib, name, smem):
    model, p, ints, reals, full, _ = _setup(name, 6)
    o_nll, o_comp, _ = Oracle(ints, reals).eval_batch(full, want_comp=True, nthreads=4)
    g_nll, g_comp, _ = emu.eval_batch(ints, reals, full, want_comp=True, smem_state=smem, lanes_per_tile=4, ncta=2)
    ncomp = len(model.comp_names)
    assert (np.abs(g_comp[:, :ncomp] - o_comp[:, :ncomp]) <= 1e-10 * np.abs(o_comp[:, :ncomp])).all()
    assert (np.abs(g_comp[:, -1] - o_comp[:, -1]) <= 1e-10 * np.abs(o_comp[:, -1])).all()
    assert (np.abs(g_comp[:, ncomp:-2] - o_comp[:, ncomp:-2]) <= 1e-12 * np.abs(o_comp[:, ncomp:-2])).all()     # random-effect terms
    us_tol = us_tolerance(model, p, o_comp[:, -2])
    assert (np.abs(g_comp[:, -2] - o_comp[:, -2]) <= us_tol).all()
    assert (np.abs(g_nll - o_nll) <= 1e-10 * np.abs(o_nll) + US_WEIGHT * us_tol).all()
    if name in ("C1", "C1_MULTI", "C2", "C2_MATAGE", "C2_STRAT", "C4", "C5", "LING"):
        assert (np.abs(g_nll - o_nll) / np.abs(o_nll)).max() < 1e-10


@pytest.mark.parametrize("name", ["C1", "C1_MULTI", "C2", "C2_MATAGE", "C5", "LING", "MINI_SPAWN", "SPAWN_RIG", "MINI_TIMEVAR", "RANDOM_RIG", "C1_RANDOM", "MINI_MATVAR", "RENEWAL_RIG", "WLOSS_RIG", "MINI_STOMACH"])
def test_gradients(oracle_lib, name):
    model, p, ints, reals, full, opt = _setup(name, 2, seed=77)
    tidx = opt[:6].astype(np.int32) if name != "C1" else opt[:10].astype(np.int32)       # 2-3 direction tiles
    o_nll, _, o_grad = Oracle(ints, reals).eval_batch(full, tangent_idx=tidx, nthreads=4)
    g_nll, g_comp, g_grad = emu.eval_batch(ints, reals, full, tangent_idx=tidx, want_comp=True, smem_state=False,
                                           lanes_per_tile=3, ncta=2)
    assert (np.abs(g_nll - o_nll) / np.abs(o_nll)).max() < 1e-10
    scale = np.abs(o_grad).max(axis=1, keepdims=True)
    assert (np.abs(g_grad - o_grad) / scale).max() < 1e-8


def test_warp_count_only_regroups_sums(oracle_lib):
    """Units and collect tasks are dealt to the warps differently; the partition into units does not depend on the warp
    count, and every sum has a fixed order for a given count -- across counts only the likelihood shares (one per warp)
    are grouped differently."""
    model, p, ints, reals, full, _ = _setup("C2", 3)
    ref = emu.eval_batch(ints, reals, full, want_comp=True)
    again = emu.eval_batch(ints, reals, full, want_comp=True)
    assert (again[0] == ref[0]).all() and (again[1] == ref[1]).all()
    for W in (1, 4, 9):
        got = emu.eval_batch(ints, reals, full, want_comp=True, W=W)
        assert (np.abs(got[0] - ref[0]) <= 1e-13 * np.abs(ref[0])).all()
        assert (np.abs(got[1][:, :4] - ref[1][:, :4]) <= 1e-13 * np.abs(ref[1][:, :4])).all()
        assert (got[1][:, 4:] == ref[1][:, 4:]).all()          # understocking and bounds: per-unit sums, keeper's order
    info = emu.plan_info(ints, reals)
    # state: NW - 1 zero rows, num and wgt of the 260 cells, the maturing fish of 3 columns, NBX - 1 slack rows
    assert info["W"] == 15 and info["ncols"] == 13 and info["n_s"] == 4 + 2 * 260 + 2 * 3 * 20 + 3
    assert info["seglen"] == [8, 8] and info["units"] == 13 * 3


@pytest.mark.parametrize("name", ["C2", "C4", "C5"])
def test_partition_into_units_only_regroups_sums(oracle_lib, name):
    """Whole columns (seg = -1), the planner's row segments and a finer partition: the ghost rows hand every unit the rows
    below it, so the results differ only by the grouping of the per-unit partial sums."""
    model, p, ints, reals, full, _ = _setup(name, 3)
    whole = emu.eval_batch(ints, reals, full, want_comp=True, seg=-1)
    assert emu.plan_info(ints, reals, seg=-1)["units"] == emu.plan_info(ints, reals)["ncols"]
    for seg in (0, 4 if name != "C4" else 16, 12 if name != "C4" else 20):
        got = emu.eval_batch(ints, reals, full, want_comp=True, seg=seg, smem_state=(seg != 0))
        ncomp = len(model.comp_names)
        assert (np.abs(got[1][:, :ncomp] - whole[1][:, :ncomp]) <= 1e-12 * np.abs(whole[1][:, :ncomp])).all()
        us_tol = us_tolerance(model, p, whole[1][:, -2], rtol=0.0)
        assert (np.abs(got[0] - whole[0]) <= 1e-12 * np.abs(whole[0]) + US_WEIGHT * us_tol).all()


def test_tiles_lanes_and_ragged_end(oracle_lib):
    """Lane, tile and CTA indexing: 7 vectors in tiles of 3 over 2 persistent CTAs == the same vectors alone."""
    model, p, ints, reals, full, _ = _setup("C1", 7)
    ref = emu.eval_batch(ints, reals, full, lanes_per_tile=1, ncta=1)[0]
    got = emu.eval_batch(ints, reals, full, lanes_per_tile=3, ncta=2)[0]
    assert (got == ref).all()


def test_nan_for_negative_retro_years_is_per_lane(oracle_lib):
    """R/action_time.R:78-79: a vector with retro_years < 0 returns NaN; its neighbours in the tile do not."""
    model, p, ints, reals, full, _ = _setup("C1", 4)
    full[2, p.index("retro_years")] = -1
    full[3, p.index("retro_years")] = 2          # a shorter run in the same tile
    o = Oracle(ints, reals).eval_batch(full)[0]
    g = emu.eval_batch(ints, reals, full, lanes_per_tile=4, ncta=1)[0]
    assert np.isnan(g[2]) and np.isnan(o[2])
    ok = [0, 1, 3]
    assert (np.abs(g[ok] - o[ok]) / np.abs(o[ok])).max() < 1e-10


def test_plan_tables_from_global_memory(oracle_lib):
    """The plan's record tables are read from shared memory when they fit beside the state and from global memory otherwise:
    same tables, same results to the bit."""
    model, p, ints, reals, full, _ = _setup("C5", 2)
    a = emu.eval_batch(ints, reals, full, want_comp=True, meta_smem=True)
    b = emu.eval_batch(ints, reals, full, want_comp=True, meta_smem=False)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


@pytest.mark.parametrize("name", ["C2", "C4"])
def test_hot_tables_in_shared_memory_or_in_the_slab(oracle_lib, name):
    """Growth bands and walpha l^wbeta live in shared memory where the planner found room and in the slab otherwise (always
    for gradient runs): same values either way."""
    model, p, ints, reals, full, _ = _setup(name, 2)
    a = emu.eval_batch(ints, reals, full, want_comp=True, hot_tables=True)
    b = emu.eval_batch(ints, reals, full, want_comp=True, hot_tables=False)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


@pytest.mark.parametrize("name", ["C1", "C2", "C4", "C5"])
def test_lean_instantiations_are_bit_identical(oracle_lib, name):
    """A model runs on the instantiation of the kernel that has the features it does not use compiled out (LEAN masks of
    g3_tile.cuh: C1 / C2 everything, C4 all but constant maturity, C5 constant maturity and the rarer likelihood functions);
    the arithmetic that remains is the same code: identical results."""
    model, p, ints, reals, full, _ = _setup(name, 3)
    a = emu.eval_batch(ints, reals, full, want_comp=True, lean=True)
    b = emu.eval_batch(ints, reals, full, want_comp=True, lean=False)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


def test_random_terms_follow_retro_years(oracle_lib):
    """g3l_random_dnorm with period 'single' scores at the last step of each vector's own run (retro_years differs by lane)."""
    model, p, ints, reals, full, _ = _setup("RANDOM_RIG", 4)
    full[1, p.index("retro_years")], full[2, p.index("retro_years")] = 2.0, 1.0
    o_nll, o_comp, _ = Oracle(ints, reals).eval_batch(full, want_comp=True, nthreads=2)
    g_nll, g_comp, _ = emu.eval_batch(ints, reals, full, want_comp=True, smem_state=True, lanes_per_tile=4, ncta=1)
    assert (np.abs(g_comp[:, :4] - o_comp[:, :4]) <= 1e-12 * np.abs(o_comp[:, :4])).all() and (o_comp[:, :2] != 0).all()
    assert (np.abs(g_nll - o_nll) <= 1e-12 * np.abs(o_nll)).all()


def test_projection_years_differ_by_lane(oracle_lib):
    """project_years is a parameter: the lanes of one tile run 0..3 years past end_year (g3_to_cuda(max_project_years = 3))."""
    model, p, ints, reals, full, _ = _setup("C1_PROJ", 5)
    pj, rt = p.index("project_years"), p.index("retro_years")
    full[1, pj], full[2, pj], full[3, pj], full[4, rt] = 1.0, 2.0, 3.0, 1.0
    o_nll, o_comp, _ = Oracle(ints, reals).eval_batch(full, want_comp=True, nthreads=2)
    for smem in (True, False):
        g_nll, g_comp, _ = emu.eval_batch(ints, reals, full, want_comp=True, smem_state=smem, lanes_per_tile=5, ncta=1)
        assert (np.abs(g_nll - o_nll) / np.abs(o_nll)).max() < 1e-10
        assert (np.abs(g_comp[:, :3] - o_comp[:, :3]) <= 1e-10 * np.abs(o_comp[:, :3])).all()
