"""The TMB truth-bundle reader (gadget3_b200/tmb_bundle.py, tools/compare_with_tmb.R): a bundle written
with known tables and parameters must rebuild the identical packed model, whichever RNG made the data."""
import numpy as np
import pytest

from gadget3_b200 import configs
from gadget3_b200.configs import parameter_batch
from gadget3_b200.tmb_bundle import TmbBundle, write_bundle
from oracle.oracle_lib import Oracle


@pytest.mark.parametrize("kind", ["single", "multi"])
def test_bundle_roundtrip_rebuilds_the_same_model(tmp_path, oracle_lib, kind):
    rng = np.random.default_rng(99)                 # NOT the configs' seed: the tables travel in the bundle
    years = list(range(1990, 2024))
    if kind == "single":
        data = configs._fleet_data(rng, years, 1000, 100, 100, range(0, 81, 20), 2000,
                                   lambda r, n: np.floor(r.uniform(1, 5, n)), lambda a: 30 + a * 10)
        model, p = configs.build_c1(data=data)
    else:
        data = configs._fleet_data(rng, years, 1e5, 10, 300, range(0, 111, 5), 2000,
                                   lambda r, n: np.floor(np.minimum(r.normal(6, 1, n), 10)), lambda a: 30 + a * 2,
                                   maturity=True)
        model, p = configs.build_c2(data=data)
    p.value[p.index("fish.K" if kind == "single" else "fish_imm.K")] = 0.27          # a non-default template value
    orc = Oracle(*model.pack(p))
    full = np.tile(p.values(), (3, 1))
    full[1:, p.optimised_indices()] = parameter_batch(p, 2, seed=5)
    tidx = np.asarray(p.optimised_indices(), dtype=np.int32)
    nll, _, grad = orc.eval_batch(full, tangent_idx=tidx)
    write_bundle(str(tmp_path), kind, data, p, full, nll, grad)

    b = TmbBundle(str(tmp_path))
    assert list(b.parameters.switch) == list(p.switch)
    np.testing.assert_array_equal(b.parameters.values(), p.values())
    np.testing.assert_array_equal(b.par, full)
    ints_a, reals_a = model.pack(p)
    ints_b, reals_b = b.model.pack(b.parameters)
    np.testing.assert_array_equal(ints_a, ints_b)
    np.testing.assert_array_equal(reals_a, reals_b)          # identical packed spec => identical objective
    np.testing.assert_array_equal(b.fn, nll)
    np.testing.assert_array_equal(b.gr, grad)


def test_bundle_rejects_a_different_parameter_order(tmp_path, oracle_lib):
    model, p = configs.build_c1()
    d = configs._fleet_data(np.random.default_rng(1), list(range(1990, 2024)), 1000, 100, 50, range(0, 81, 20), 500,
                            lambda r, n: np.floor(r.uniform(1, 5, n)), lambda a: 30 + a * 10)
    q = p.copy()
    q.switch[3], q.switch[4] = q.switch[4], q.switch[3]
    write_bundle(str(tmp_path), "single", d, q, np.zeros((1, len(p))), np.zeros(1), np.zeros((1, 1)))
    with pytest.raises(ValueError, match="Parameters not in expected order"):
        TmbBundle(str(tmp_path))
