"""g3a_predate_catchability_numberfleet / linearfleet / effortfleet in the oracle (R/action_predate.R:7-18,117-126).

The reference's tests hold no numbers for these (tests/test-action_predate-numberfleet.R only checks
R == TMB), so the restatement is checked through what the formulas promise, read back from the
detail_* report arrays: a numberfleet lands E individuals, a linearfleet takes E * cur_step_size of the
suitable biomass, an effortfleet catchability * E * cur_step_size of it.
"""
import numpy as np
import pytest

from gadget3_b200.configs import CONFIGS
from gadget3_b200.run_cuda import unpack_report
from oracle.oracle_lib import Oracle


@pytest.fixture(scope="module")
def run(oracle_lib):
    model, p = CONFIGS["MINI_FLEETS"]()
    rep = unpack_report(model, Oracle(*model.pack(p)).report(p.values()))
    acts = {a.predstock.name: a for a in model.actions if getattr(a, "kind", "") == "predate"}
    return model, p, rep, acts


def _E(act, t, area_id):
    return act.catchability_f.E.lookup(2000 + t // 4, t % 4 + 1, area_id)


def test_numberfleet_lands_E_individuals(run):
    model, p, rep, acts = run
    cons, wgt = rep["detail_fish_f_num__cons"], rep["dstart_fish__wgt"]          # [length, age, area, time]
    assert float(rep["nll_understocking__wgt"].sum()) < 1e-6                      # no overconsumption in this run
    for t in range(16):
        for r, aid in enumerate((1, 2)):
            E = _E(acts["f_num"], t, aid)
            caught = (cons[:, :, r, t] / wgt[:, :, r, t]).sum()
            if E == 0:
                assert caught == 0
            else:
                assert caught == pytest.approx(E, rel=2e-3)       # (cells with ~0 consumption lose a little to avoid_zero)


@pytest.mark.parametrize("fleet,scale", [("f_lin", None), ("f_eff", "fish.f_eff.catchability")])
def test_rate_fleets_take_a_share_of_the_suitable_biomass(run, fleet, scale):
    model, p, rep, acts = run
    cons, suit = rep[f"detail_fish_{fleet}__cons"], rep[f"detail_fish_{fleet}__suit"]
    c = 1.0 if scale is None else p.value[p.index(scale)]
    seen = 0
    for t in range(16):
        for r, aid in enumerate((1, 2)):
            E = _E(acts[fleet], t, aid)
            big = suit[:, :, r, t] * c * E * 0.25 > 1.0           # cells whose consumption is far from avoid_zero's knee
            if E == 0:
                assert (cons[:, :, r, t] == 0).all()
                continue
            np.testing.assert_allclose(cons[:, :, r, t][big], (suit[:, :, r, t] * (c * E * 0.25))[big], rtol=1e-9)
            seen += int(big.sum())
    assert seen > 100


def test_sumofsquaredlogratios_from_the_reported_model_array(run, oracle_lib):
    """g3l_distribution_sumofsquaredlogratios (R/likelihood_distribution.R:127-136): the component total must equal
    sum((log(obs + 10) - log(model + 10))^2) over the reported model array."""
    model, p, rep, acts = run
    name = "cdist_sumofsquaredlogratios_kilos_f_lin"
    act = [a for a in model.actions if getattr(a, "nll_name", "") == name][0]
    obs = np.asarray(act.ld.obs).reshape(rep[f"{name}_model__wgt"].shape)
    want = ((np.log(obs + 10.0) - np.log(rep[f"{name}_model__wgt"] + 10.0)) ** 2).sum()
    assert rep[f"nll_{name}__wgt"].sum() == pytest.approx(want, rel=1e-12)
    assert want > 0
