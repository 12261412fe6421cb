"""Parity of the CUDA path (through the C ABI) with the CPU oracle. Needs a B200: `-m gpu`.

Tolerances are the contract's (BASELINE.json north_star): 1e-10 relative on nll and per-component
reports, 1e-8 on gradients (relative to the largest gradient entry of the vector).
"""
import numpy as np
import pytest

from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import g3_cuda_adfun, unpack_report
from oracle.oracle_lib import Oracle

pytestmark = pytest.mark.gpu

NLL_RTOL = 1e-10
GRAD_RTOL = 1e-8


@pytest.fixture(scope="module")
def setups(oracle_lib):
    out = {}
    for n in ("C1", "C1_MULTI", "C2", "C2_MATAGE", "C2_STRAT", "C4", "C5"):
        model, p = CONFIGS[n]()
        fn = g3_cuda_adfun(model, p, device=0)
        out[n] = (model, p, fn, Oracle(*model.pack(p)))
    return out


def _rel(a, b, floor=0.0):
    return np.abs(a - b) / np.maximum(np.abs(b), floor if floor > 0 else 1e-300)


from _util import US_WEIGHT, check_detail_reports, golden, us_tolerance as _us_tolerance  # noqa: E402


def _us_weight(model):
    return US_WEIGHT


KERNELS = pytest.mark.parametrize("kernel", [4, 3], ids=["tile_kernel", "thread_kernel"])


@KERNELS
@pytest.mark.parametrize("name", ["C1", "C1_MULTI", "C2", "C2_MATAGE", "C2_STRAT", "C4", "C5"])
def test_nll_and_components(setups, name, kernel):
    """C1_MULTI: g3l_distribution_multinomial + g3l_distribution_surveyindices_linear (R/likelihood_distribution.R:41-60,
    82-125; the oracle's versions are re-derived independently in tests/test_oracle_likelihoods.py).
    C2_STRAT: stratified sumofsquares, proportions within each length group (R/likelihood_distribution.R:8-16).
    C2_MATAGE: g3a_mature_continuous with its age term, beta * cur_step_size and a50 (R/action_mature.R:18,44-74; the
    oracle's share against a hand value in tests/test_oracle_kat.py)."""
    model, p, fn, orc = setups[name]
    if kernel == 3 and not fn.handle.tile_info()["thread_ok"]:
        pytest.skip("the thread-per-evaluation kernel does not take this model (tile kernel only)")
    fn.handle.set_kernel(kernel)
    try:
        _check_nll_and_components(model, p, fn, orc)
    finally:
        fn.handle.set_kernel(0)


def _check_nll_and_components(model, p, fn, orc):
    nvec = 63 if len(model.stock_names) and sum(L * A_ * R for L, A_, R in model.stock_shapes) > 400 else 255
    P = np.vstack([fn.par[None, :], parameter_batch(p, nvec, seed=2026)])
    full = fn.full_par(P)
    g_nll, g_comp, _ = fn.handle.eval_batch(full, want_comp=True)
    o_nll, o_comp, _ = orc.eval_batch(full, want_comp=True, nthreads=8)
    assert np.isfinite(g_nll).all()
    ncomp = len(model.comp_names)
    assert _rel(g_comp[:, :ncomp], o_comp[:, :ncomp]).max() < NLL_RTOL
    assert _rel(g_comp[:, -1], o_comp[:, -1]).max() < NLL_RTOL            # bounds penalty
    # the objective: the contract's 1e-10, plus the conditioning of g3l_understocking's difference of sums (a few ulps of
    # the consumed biomass, _us_tolerance; measured sizes in tools/gpu_parity_diag.py: at most 3e-11 of nll)
    w = _us_weight(model)
    assert (np.abs(g_nll - o_nll) <= NLL_RTOL * np.abs(o_nll) + w * _us_tolerance(model, p, o_comp[:, -2], rtol=0.0)).all()
    # everything except the understocking term agrees far better
    assert (np.abs((g_nll - w * g_comp[:, -2]) - (o_nll - w * o_comp[:, -2])) / np.abs(o_nll)).max() < 2e-11
    assert np.median(np.abs((g_nll - w * g_comp[:, -2]) - (o_nll - w * o_comp[:, -2])) / np.abs(o_nll)) < 1e-12
    # the understocking term itself: within its conditioning (see _us_tolerance)
    assert (np.abs(g_comp[:, -2] - o_comp[:, -2]) <= _us_tolerance(model, p, o_comp[:, -2])).all()


@KERNELS
@pytest.mark.parametrize("name", ["C1", "C1_MULTI", "C2", "C2_MATAGE", "C2_STRAT", "C4", "C5"])
def test_gradients(setups, name, kernel):
    model, p, fn, orc = setups[name]
    if kernel == 3 and not fn.handle.tile_info()["thread_ok"]:
        pytest.skip("the thread-per-evaluation kernel does not take this model (tile kernel only)")
    fn.handle.set_kernel(kernel)
    try:
        _check_gradients(model, p, fn, orc)
    finally:
        fn.handle.set_kernel(0)


def _check_gradients(model, p, fn, orc):
    nv = 2 if sum(L * A_ * R for L, A_, R in model.stock_shapes) > 400 else 6
    P = parameter_batch(p, nv, seed=77)
    full = fn.full_par(P)
    g_nll, g_grad = fn.gr_batch(P)
    _, _, o_grad = orc.eval_batch(full, tangent_idx=fn.opt_idx.astype(np.int32), nthreads=8)
    o_nll, o_comp, _ = orc.eval_batch(full, want_comp=True)
    assert g_grad.shape == (nv, len(fn.opt_idx))
    assert (np.abs(g_nll - o_nll) <= NLL_RTOL * np.abs(o_nll) + US_WEIGHT * _us_tolerance(model, p, o_comp[:, -2])).all()
    scale = np.abs(o_grad).max(axis=1, keepdims=True)
    # 1e-8, plus the relative conditioning of the understocking term where it is active (its gradient is
    # 2e8 (sum tp - sum tp') d(sum tp - sum tp'): the same difference of near-equal sums as the value)
    gtol = GRAD_RTOL + US_WEIGHT * _us_tolerance(model, p, o_comp[:, -2]) / np.abs(o_nll)
    assert ((np.abs(g_grad - o_grad) / scale).max(axis=1) < gtol).all()


@pytest.mark.parametrize("name", ["C1", "C1_MULTI", "C2", "C2_MATAGE", "C2_STRAT", "C4", "C5"])
def test_report_arrays(setups, name):
    model, p, fn, orc = setups[name]
    par = parameter_batch(p, 1, seed=5)[0]
    rg = fn.report(par)
    ro = unpack_report(model, orc.report(fn.full_par(par[None, :])[0]))
    assert set(rg) == set(ro)
    for k in ro:
        a, b = np.asarray(rg[k], dtype=float), np.asarray(ro[k], dtype=float)
        assert a.shape == b.shape, k
        assert (np.isnan(a) == np.isnan(b)).all(), k
        m = ~np.isnan(b)
        if not m.any():
            continue
        if k.startswith("nll_understocking") or k == "nll":
            us_tol = _us_tolerance(model, p, ro["nll_understocking__wgt"].sum())
            assert np.abs(a - b)[m].sum() <= (NLL_RTOL * abs(ro["nll"]) if k == "nll" else 0.0) + (US_WEIGHT if k == "nll" else 1.0) * us_tol, k
        else:
            scale = np.abs(b[m]).max() if m.any() else 1.0
            if scale == 0.0:
                assert (a[m] == 0.0).all(), k
                continue
            assert (np.abs(a - b)[m] / np.maximum(np.abs(b[m]), 1e-12 * scale)).max() < 1e-9, k


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_detail_reports_against_reference_generated_code(setups, name):
    """detail_*__renewalnum/suit/cons/predby and suit_*__report from g3cuda_report() vs the arrays the
    reference's generated code REPORT()s (tests/golden/reference_*.npz)."""
    model, p, fn, orc = setups[name]
    g = golden(name)
    for vi in (0, 5):
        rep = unpack_report(model, fn.handle.report(g["par"][vi]))
        assert check_detail_reports(model, rep, g, vi) >= 5
        for sn in model.stock_names:
            for what in ("num", "wgt"):
                np.testing.assert_allclose(rep[f"dstart_{sn}__{what}"][..., g["report_times"]],
                                           g[f"v{vi}__dstart_{sn}__{what}"], rtol=1e-9, atol=1e-300)


def test_single_vector_fn_gr(setups):
    model, p, fn, orc = setups["C1"]
    v = fn.fn()
    assert v == pytest.approx(orc.eval_batch(fn.full_par(fn.par[None, :]))[0][0], rel=1e-9)
    g = fn.gr()
    assert g.shape == (1, len(fn.par))
    assert np.isfinite(g).all()


def test_nan_for_negative_retro_years(setups):
    model, p, fn, orc = setups["C1"]
    full = fn.full_par(parameter_batch(p, 4, seed=1))
    full[2, p.index("retro_years")] = -1
    nll = fn.handle.eval_batch(full)[0]
    assert np.isnan(nll[2]) and np.isfinite(np.delete(nll, 2)).all()


def test_empty_and_single_batches(setups):
    model, p, fn, orc = setups["C1"]
    nll, _, _ = fn.handle.eval_batch(np.empty((0, fn.handle.n_par)))
    assert nll.shape == (0,)
    one = fn.full_par(fn.par[None, :])
    assert fn.handle.eval_batch(one)[0].shape == (1,)
    with pytest.raises(ValueError):
        fn.handle.eval_batch(np.zeros((3, fn.handle.n_par - 1)))


def test_full_size_batch_properties(setups):
    """BASELINE.json config 2 at full size (65,536 vectors): the oracle checks a slice; the rest
    is covered by size-independent properties -- results do not depend on the position in the
    batch (permutation) or on the batch a vector travels in, and repeated rows agree to the bit."""
    model, p, fn, orc = setups["C2"]
    B = 65536
    P = parameter_batch(p, B, seed=4242)
    P[B // 2:B // 2 + 64] = P[:64]                         # repeated rows
    full = fn.full_par(P)
    nll = fn.handle.eval_batch(full)[0]
    assert np.isfinite(nll).all()
    assert (nll[B // 2:B // 2 + 64] == nll[:64]).all()
    o, oc, _ = orc.eval_batch(full[:96], want_comp=True, nthreads=8)
    assert (np.abs(nll[:96] - o) <= NLL_RTOL * np.abs(o) + _us_weight(model) * _us_tolerance(model, p, oc[:, -2])).all()
    perm = np.random.default_rng(0).permutation(B)
    nll_p = fn.handle.eval_batch(full[perm])[0]
    assert (nll_p == nll[perm]).all()
    small = fn.handle.eval_batch(full[1000:1100])[0]        # same kernel, smaller batch (partly filled tiles): identical to the bit
    assert (small == nll[1000:1100]).all()
    one = fn.handle.eval_batch(full[1234:1235])[0]
    assert one[0] == nll[1234]
    # the thread-per-evaluation kernel on the same vectors: same values to rounding
    fn.handle.set_kernel(3)
    try:
        thr = fn.handle.eval_batch(full[1000:1100], want_comp=True)
    finally:
        fn.handle.set_kernel(0)
    tol = NLL_RTOL * np.abs(small) + _us_weight(model) * _us_tolerance(model, p, thr[1][:, -2])
    assert (np.abs(thr[0] - small) <= tol).all()
    # the warp count of a tile only regroups the likelihood shares (one per warp); whole columns instead of row segments
    # regroup the per-unit partial sums
    fn.handle.set_tile_warps(7)
    try:
        w7 = fn.handle.eval_batch(full[1000:1100])[0]
    finally:
        fn.handle.set_tile_warps(0)
    assert _rel(w7, small).max() < 1e-13
    fn.handle.set_tile_partition(0, -1)
    try:
        whole = fn.handle.eval_batch(full[1000:1100])[0]
    finally:
        fn.handle.set_tile_partition(0, 0)
    assert (np.abs(whole - small) <= 1e-12 * np.abs(small) + _us_weight(model) * _us_tolerance(model, p, thr[1][:, -2], rtol=0.0)).all()


def test_ragged_batch_larger_than_one_wave(setups):
    """More vectors than are resident at once, and not a multiple of the tile size: the persistent CTAs walk several
    tiles, the last one ragged; every row must still equal the same row evaluated in a small batch."""
    model, p, fn, orc = setups["C1"]
    B = 200003
    full = fn.full_par(parameter_batch(p, B, seed=99))
    nll, comp, _ = fn.handle.eval_batch(full, want_comp=True)
    assert nll.shape == (B,) and np.isfinite(nll).all()
    for lo in (0, 77777, 155640, B - 37):
        small = fn.handle.eval_batch(full[lo:lo + 37], want_comp=True)
        assert (small[0] == nll[lo:lo + 37]).all() and (small[1] == comp[lo:lo + 37]).all()
    fn.handle.set_kernel(3)         # the thread kernel: more vectors than its workspace holds at once (152 x 1024 slots)
    try:
        t_nll = fn.handle.eval_batch(full)[0]
    finally:
        fn.handle.set_kernel(0)
    # two summation orders of the understocking sums: each within that term's conditioning of the other
    assert (np.abs(t_nll - nll) <= NLL_RTOL * np.abs(nll) + 2 * US_WEIGHT * _us_tolerance(model, p, comp[:, -2], rtol=0.0)).all()
    o, oc, _ = orc.eval_batch(full[B - 64:], want_comp=True, nthreads=8)
    assert (np.abs(nll[B - 64:] - o) <= NLL_RTOL * np.abs(o) + US_WEIGHT * _us_tolerance(model, p, oc[:, -2])).all()


def test_device_pointer_entry_point(setups):
    import torch
    model, p, fn, orc = setups["C2"]
    full = fn.full_par(parameter_batch(p, 512, seed=8))
    d_par = torch.from_numpy(full).cuda()
    d_nll = torch.empty(512, dtype=torch.float64, device="cuda")
    d_comp = torch.empty((512, fn.handle.n_nll), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream()
    n0 = fn.handle.launch_count()
    fn.handle.eval_batch_device(d_par.data_ptr(), 512, d_nll.data_ptr(), d_comp=d_comp.data_ptr(), stream=st.cuda_stream)
    torch.cuda.synchronize()
    assert fn.handle.launch_count() == n0 + 1
    assert fn.handle.last_kernel_ms() > 0
    host = fn.handle.eval_batch(full, want_comp=True)
    assert (d_nll.cpu().numpy() == host[0]).all()
    assert (d_comp.cpu().numpy() == host[1]).all()


def test_growth_beyond_maxlengthgroupgrowth(setups):
    """Mean length jump larger than maxlengthgroupgrowth (alpha < 0 in growth_bbinom): the reference's
    lgamma form yields exp(log|Gamma| ...) -- the same regime as the KAT tables of
    tests/test-action_grow.R:34-62. The lgamma-free product form on the device must agree."""
    model, p, fn, orc = setups["C1"]
    full = fn.full_par(parameter_batch(p, 8, seed=3))
    full[:, p.index("fish.Linf")] = np.linspace(300, 600, 8)
    full[:, p.index("fish.K")] = 1.2
    g = fn.handle.eval_batch(full)[0]
    o = orc.eval_batch(full)[0]
    ok = np.isfinite(o)          # lgamma poles (alpha + x a non-positive integer) make the reference NaN
    assert ok.sum() >= 2
    assert _rel(g[ok], o[ok]).max() < 1e-9


@KERNELS
def test_ling_against_reference_generated_code(oracle_lib, kernel):
    """CUDA vs inttest/codegeneration/ling.cpp (constant maturity, maxlengthgroupgrowth 15, two renewals,
    age -> mature movement): golden nll from the reference's generated code, gradient vs the oracle."""
    model, p = CONFIGS["LING"]()
    fn = g3_cuda_adfun(model, p, device=0)
    g = golden("LING")
    assert list(g["switch"]) == list(p.switch)
    fn.handle.set_kernel(kernel)
    try:
        nll, comp, _ = fn.handle.eval_batch(g["par"], want_comp=True)
        _check_gradients(model, p, fn, Oracle(*model.pack(p)))
    finally:
        fn.handle.set_kernel(0)
    tol = NLL_RTOL * np.abs(g["nll"]) + US_WEIGHT * _us_tolerance(model, p, comp[:, -2])
    assert (np.abs(nll - g["nll"]) <= tol).all()
    assert np.median(_rel(nll, g["nll"])) < 1e-12


@KERNELS
@pytest.mark.parametrize("name", ["C1", "C2"])
def test_against_reference_generated_code(setups, name, kernel):
    """CUDA nll vs the reference's OWN generated objective (baseline/*.cpp compiled against
    oracle/tmb_shim; vectors and results in tests/golden/reference_*.npz, tools/make_golden_reference.py)."""
    model, p, fn, orc = setups[name]
    g = golden(name)
    assert list(g["switch"]) == list(p.switch)
    fn.handle.set_kernel(kernel)
    try:
        nll, comp, _ = fn.handle.eval_batch(g["par"], want_comp=True)
    finally:
        fn.handle.set_kernel(0)
    tol = NLL_RTOL * np.abs(g["nll"]) + US_WEIGHT * _us_tolerance(model, p, comp[:, -2])
    assert (np.abs(nll - g["nll"]) <= tol).all()
    assert np.median(_rel(nll, g["nll"])) < 1e-12


@KERNELS
@pytest.mark.parametrize("name", ["MINI_ANDERSEN", "MINI_OTHERFOOD", "MINI_PREFS", "MINI_SUITS", "MINI_FLEETS", "MINI_MIGRATE", "MINI_TRANSFORM", "SPAWN_RIG", "MINI_SPAWN", "MINI_TIMEVAR", "TIMEVAR_RIG", "SIX_FLEETS", "MINI_MATVAR", "RENEWAL_RIG", "WLOSS_RIG", "MINI_STOMACH"])
def test_andersen_predator_model(oracle_lib, name, kernel):
    """g3_suitability_andersen (predator-length dependent), one-step-only migration: CUDA vs oracle.
    MINI_SUITS: richards for the predator, gamma / andersenfleet / straightline fleets, reports included.
    MINI_FLEETS: numberfleet, linearfleet and effortfleet catchabilities.
    MINI_OTHERFOOD: the predator also eats a g3a_otherfood stock (assigned every step, R/action_renewal.R:276-308).
    MINI_MATVAR: time-varying maturity parameters (alpha by year, l50 a g3_timevariable; tests/test_oracle_matvar.py).
    RENEWAL_RIG: renewal mean length by year, standard deviation a g3_timevariable (one record per shape; the oracle's
    renewal against the normal density by hand in tests/test_oracle_kat.py).
    WLOSS_RIG: spawning weight loss, relative and with a floor (by hand in tests/test_oracle_spawn.py).
    MINI_STOMACH: stomach contents -- g3l_catchdistribution(predators = ...) on the stock predator's consumption.
    SIX_FLEETS: three stocks, six fleets, 15 likelihood components, 9 collect vectors, five predators on one stock."""
    model, p = CONFIGS[name]()
    fn = g3_cuda_adfun(model, p, device=0)
    orc = Oracle(*model.pack(p))
    P = np.vstack([fn.par[None, :], parameter_batch(p, 63, seed=12)])
    full = fn.full_par(P)
    fn.handle.set_kernel(kernel)        # (the handle is this test's own: no need to restore)
    g_nll, g_comp, _ = fn.handle.eval_batch(full, want_comp=True)
    o_nll, o_comp, _ = orc.eval_batch(full, want_comp=True, nthreads=8)
    assert np.isfinite(o_nll).all()
    ncomp = len(model.comp_names)
    assert _rel(g_comp[:, :ncomp], o_comp[:, :ncomp]).max() < NLL_RTOL
    us_tol = _us_tolerance(model, p, o_comp[:, -2])
    assert (np.abs(g_comp[:, -2] - o_comp[:, -2]) <= us_tol).all()
    assert (np.abs(g_nll - o_nll) <= NLL_RTOL * np.abs(o_nll) + US_WEIGHT * us_tol).all()
    g2, grad = fn.gr_batch(P[:3])
    _, _, og = orc.eval_batch(full[:3], tangent_idx=fn.opt_idx.astype(np.int32))
    # bounded_vec(1000 (p1 - log(L_pred / l)), 0, 1) overflows exp() for distant length pairs: the p1
    # derivative is NaN there (0 * inf, as in the reference's reverse-mode tape); every other direction is finite
    assert (np.isnan(grad) == np.isnan(og)).all()
    ok = ~np.isnan(og)
    assert ok.sum() >= og.size - 3 * 2
    scale = np.nanmax(np.abs(og), axis=1, keepdims=True)
    gtol = GRAD_RTOL + US_WEIGHT * us_tol[:3] / np.abs(o_nll[:3])
    err = np.where(ok, np.abs(grad - og) / scale, 0.0)
    assert (err.max(axis=1) < gtol).all()
    rg, ro = fn.report(P[1]), unpack_report(model, orc.report(full[1]))
    for k in ro:
        if k.startswith("suit_") or k.startswith("detail_"):
            np.testing.assert_allclose(rg[k], ro[k], rtol=1e-9, atol=1e-12 * max(np.abs(ro[k]).max(), 1e-300), err_msg=k)


def test_profile_sweep_c5(setups):
    """Config 5's workload shape: a likelihood-profile grid over two parameters (here 64 x 64; the bench
    runs 131,072 vectors per GPU). Oracle on a slice; the grid is symmetric in nothing, but every point
    must equal the same vector evaluated alone, and neighbouring points must vary smoothly."""
    from gadget3_b200.configs import profile_sweep
    model, p, fn, orc = setups["C5"]
    P = profile_sweep(p, ["prey.K", "prey.pred.l50"], 64)
    full = fn.full_par(P)
    nll, comp, _ = fn.handle.eval_batch(full, want_comp=True)
    assert nll.shape == (4096,) and np.isfinite(nll).all()
    pick = np.random.default_rng(5).choice(4096, 24, replace=False)
    o, oc, _ = orc.eval_batch(full[pick], want_comp=True, nthreads=8)
    tol = NLL_RTOL * np.abs(o) + US_WEIGHT * _us_tolerance(model, p, oc[:, -2])
    assert (np.abs(nll[pick] - o) <= tol).all()
    alone = fn.handle.eval_batch(full[pick])[0]
    assert (alone == nll[pick]).all()


def test_multi_device_handle(setups):
    """One handle over several devices (g3cuda_multi_*): the rows of a batch are cut into one
    contiguous block per device and evaluated concurrently, results land in the caller's buffers. With one GPU visible
    the same device is listed twice (two handles, two host threads); with more, every device takes part. Values,
    components and gradients must equal the single-device call to the bit (same kernels, same tiles of 32 rows or not:
    a vector's result does not depend on its neighbours)."""
    import torch
    model, p, fn, orc = setups["C1"]
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0]
    multi = g3_cuda_adfun(model, p, devices=devices)
    assert multi.multi is not None and len(multi.multi.handles) == len(devices)
    P = parameter_batch(p, 1003, seed=31)             # uneven split
    nll1, comp1 = fn.nll_components_batch(P)
    nllm, compm = multi.nll_components_batch(P)
    assert (nllm == nll1).all() and (compm == comp1).all()
    g1 = fn.gr_batch(P[:65])
    gm = multi.gr_batch(P[:65])
    assert (gm[0] == g1[0]).all() and (gm[1] == g1[1]).all()
    assert multi.fn(P[3]) == fn.fn(P[3])
    # an error on any device is the call's error
    bad = fn.full_par(P[:10])[:, :-1]
    with pytest.raises(ValueError):
        multi.multi.eval_batch(bad)


def test_projection_years_are_refused_not_nan(setups):
    """project_years >= 1 would need tables beyond the model's years: G3CUDA_EUNSUPP, not a NaN that looks like a value."""
    from gadget3_b200._lib import G3CudaError
    model, p, fn, orc = setups["C1"]
    full = fn.full_par(parameter_batch(p, 4, seed=1))
    full[2, p.index("project_years")] = 1
    with pytest.raises(G3CudaError, match="project_years"):
        fn.handle.eval_batch(full)
    assert np.isfinite(fn.handle.eval_batch(full[:2])[0]).all()


@pytest.mark.parametrize("name", ["C1", "C2", "C4", "C5"])
def test_lean_instantiations_are_bit_identical(setups, name):
    """By default a model runs on the instantiation of the tile kernel that has the features it does not use compiled out
    (LEAN masks, g3_tile.cuh); g3cuda_set_kernel(m, 5) selects the full instantiation. The same source for what remains (the
    CPU emulation gives identical bits, tests/test_tile_emu.py); on the device the two compilations contract and order a few
    operations differently, so they agree to rounding."""
    model, p, fn, orc = setups[name]
    full = fn.full_par(parameter_batch(p, 96, seed=44))
    try:
        fn.handle.set_kernel(4)
        a = fn.handle.eval_batch(full, want_comp=True)
        fn.handle.set_kernel(5)
        b = fn.handle.eval_batch(full, want_comp=True)
    finally:
        fn.handle.set_kernel(0)
    ncomp = len(model.comp_names)
    assert _rel(a[1][:, :ncomp], b[1][:, :ncomp]).max() < 1e-12
    tol = 1e-12 * np.abs(b[0]) + US_WEIGHT * _us_tolerance(model, p, b[1][:, -2])
    assert (np.abs(a[0] - b[0]) <= tol).all()
    assert np.median(_rel(a[0], b[0])) < 1e-14


@pytest.mark.parametrize("name", ["C2", "MINI_SPAWN", "SPAWN_RIG", "MINI_OTHERFOOD", "MINI_PREFS", "MINI_TIMEVAR", "TIMEVAR_RIG", "C1_RANDOM", "MINI_MATVAR"])
def test_results_do_not_depend_on_scheduling(oracle_lib, name):
    """The work queues hand units to whichever warp is free and the spawning / hand-over phases run beside each other: every
    sum has a fixed order all the same, so repeated launches, and a vector alone or among 300 others, give identical bits
    (a missing barrier or a racing update would show here)."""
    model, p = CONFIGS[name]()
    fn = g3_cuda_adfun(model, p, device=0)
    full = fn.full_par(parameter_batch(p, 300, seed=9))
    ref = fn.handle.eval_batch(full, want_comp=True)
    for _ in range(4):
        again = fn.handle.eval_batch(full, want_comp=True)
        assert (again[0] == ref[0]).all() and (again[1] == ref[1]).all()
    for i in (0, 131, 299):
        alone = fn.handle.eval_batch(full[i:i + 1], want_comp=True)
        assert alone[0][0] == ref[0][i] and (alone[1][0] == ref[1][i]).all()
    tidx = fn.opt_idx[:5].astype(np.int32)
    g1 = fn.handle.eval_batch(full[:40], tangent_idx=tidx)
    g2 = fn.handle.eval_batch(full[:40], tangent_idx=tidx)
    assert (g1[2] == g2[2]).all() or (np.isnan(g1[2]) == np.isnan(g2[2])).all() and np.array_equal(np.nan_to_num(g1[2]), np.nan_to_num(g2[2]))


@KERNELS
@pytest.mark.parametrize("name", ["RANDOM_RIG", "C1_RANDOM"])
def test_random_effect_terms(oracle_lib, name, kernel):
    """g3l_random_dnorm / g3l_random_walk (R/likelihood_random.R:14-109): the rig of the reference's own test
    (tests/test-likelihood_random.R:6-34; the oracle's terms against the normal density by hand in
    tests/test_oracle_random.py) and the vignette model with recruitment deviations: values, every nll row, gradients
    and the reported per-step rows, CUDA vs oracle."""
    model, p = CONFIGS[name]()
    fn = g3_cuda_adfun(model, p, device=0)
    orc = Oracle(*model.pack(p))
    P = np.vstack([fn.par[None, :], parameter_batch(p, 63, seed=21)])
    full = fn.full_par(P)
    full[7, p.index("retro_years")], full[9, p.index("retro_years")] = 2.0, 1.0      # 'single' terms score at the run's own last step
    fn.handle.set_kernel(kernel)
    g_nll, g_comp, _ = fn.handle.eval_batch(full, want_comp=True)
    o_nll, o_comp, _ = orc.eval_batch(full, want_comp=True, nthreads=8)
    ncomp, nrand = len(model.comp_names), len(model.random_names)
    assert nrand >= 2 and g_comp.shape[1] == ncomp + nrand + 2
    assert _rel(g_comp[:, ncomp:ncomp + nrand], o_comp[:, ncomp:ncomp + nrand]).max() < 1e-13
    if ncomp:
        assert _rel(g_comp[:, :ncomp], o_comp[:, :ncomp]).max() < NLL_RTOL
    us_tol = _us_tolerance(model, p, o_comp[:, -2]) if model.pp_info else np.zeros(len(o_nll))
    assert (np.abs(g_nll - o_nll) <= NLL_RTOL * np.abs(o_nll) + US_WEIGHT * us_tol).all()
    if name == "RANDOM_RIG":        # by hand, as the reference's test does
        import math
        v = dict(zip(p.switch, full[5]))
        ld = lambda x, mu, sd: -0.5 * math.log(2 * math.pi) - math.log(sd) - 0.5 * ((x - mu) / sd) ** 2
        wy = [v[f"walk_year.{y}"] for y in range(1990, 1995)]
        hand = (-ld(v["dnorm_log"], v["dnorm_log_mean"], v["dnorm_log_sigma"]) + math.exp(ld(v["dnorm_lin"], v["dnorm_lin_mean"], v["dnorm_lin_sigma"]))
                + sum(-ld(wy[i + 1], wy[i], v["walk_year_sigma"]) for i in range(4)))
        ws = [v[f"walk_step.{y}.{s}"] for y in range(1990, 1995) for s in range(1, 5)]
        hand += sum(-ld(ws[i + 1], ws[i], v["walk_step_sigma"]) for i in range(19))
        assert g_nll[5] == pytest.approx(hand, rel=1e-12)
    g2, grad = fn.gr_batch(P[:4])
    _, _, og = orc.eval_batch(full[:4], tangent_idx=fn.opt_idx.astype(np.int32))
    scale = np.abs(og).max(axis=1, keepdims=True)
    gtol = GRAD_RTOL + US_WEIGHT * us_tol[:4] / np.abs(o_nll[:4])
    assert ((np.abs(grad - og) / scale).max(axis=1) < gtol).all()
    rg, ro = fn.report(P[1]), unpack_report(model, orc.report(full[1]))
    for n in model.random_names:
        np.testing.assert_allclose(rg[n + "__dnorm"], ro[n + "__dnorm"], rtol=1e-13, atol=0.0, err_msg=n)


@KERNELS
def test_projection_years(oracle_lib, kernel):
    """project_years up to the room the model was lowered with (g3_to_cuda(max_project_years = 3); R/action_time.R:21-78; the
    oracle's projected years by hand in tests/test_oracle_projection.py): vectors that project 0..3 years side by side, values,
    gradients and the reported history against the oracle; one year more is refused, not answered with a NaN."""
    from gadget3_b200._lib import G3CudaError
    model, p = CONFIGS["C1_PROJ"]()
    fn = g3_cuda_adfun(model, p, device=0)
    orc = Oracle(*model.pack(p))
    P = np.vstack([fn.par[None, :], parameter_batch(p, 63, seed=33)])
    full = fn.full_par(P)
    pj = p.index("project_years")
    full[:, pj] = np.arange(64) % 4
    full[4, p.index("retro_years")] = 1.0         # (with project_years = 0: a retro year that is projected again has its data back)
    fn.handle.set_kernel(kernel)
    g_nll, g_comp, _ = fn.handle.eval_batch(full, want_comp=True)
    o_nll, o_comp, _ = orc.eval_batch(full, want_comp=True, nthreads=8)
    ncomp = len(model.comp_names)
    assert _rel(g_comp[:, :ncomp], o_comp[:, :ncomp]).max() < NLL_RTOL
    us_tol = _us_tolerance(model, p, o_comp[:, -2])
    assert (np.abs(g_nll - o_nll) <= NLL_RTOL * np.abs(o_nll) + US_WEIGHT * us_tol).all()
    # no observation and no landing in the projected years: the objective does not move
    full0 = full.copy(); full0[:, pj] = 0
    assert np.array_equal(fn.handle.eval_batch(full0)[0], g_nll)
    tix = fn.opt_idx.astype(np.int32)[:12]
    _, _, gg = fn.handle.eval_batch(full[:4], tangent_idx=tix)
    _, _, og = orc.eval_batch(full[:4], tangent_idx=tix)
    assert ((np.abs(gg - og) / np.abs(og).max(axis=1, keepdims=True)).max(axis=1) < GRAD_RTOL + US_WEIGHT * us_tol[:4] / np.abs(o_nll[:4])).all()
    rg, ro = unpack_report(model, fn.handle.report(full[3])), unpack_report(model, orc.report(full[3]))      # projects 3 years
    for k in ("dstart_fish__num", "dstart_fish__wgt", "detail_fish__renewalnum"):
        assert ro[k][..., -1].sum() > 0, k
        np.testing.assert_allclose(rg[k], ro[k], rtol=1e-9, atol=1e-12 * np.abs(ro[k]).max(), err_msg=k)
    full[2, pj] = 4
    with pytest.raises(G3CudaError, match="max_project_years"):
        fn.handle.eval_batch(full)
