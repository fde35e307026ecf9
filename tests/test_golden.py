"""The oracle against the committed golden vectors (tests/golden/golden.npz, outputs
of the unmodified reference run by tests/golden/make_golden.py) and against the
survey's learning-curve anchors.  CPU only."""
import os

import numpy as np
import pytest

from bidinet_b200 import config as cf
from oracle.pyoracle import OracleHierarchy, c1_frame

from util import N_RANDOM, run_random_case, small_cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("case", small_cases(), ids=lambda c: c[0])
def test_oracle_reproduces_golden(case, golden):
    import sys
    sys.path.insert(0, os.path.dirname(GOLDEN))
    from make_golden import run_case
    name, cfg, ld, frame, reward = case
    preds, keys, digests = run_case(OracleHierarchy, cfg, ld, frame, reward)
    assert np.array_equal(preds, golden[name + "/pred"])
    assert list(keys) == list(golden[name + "/keys"])
    assert list(digests) == list(golden[name + "/digests"])


@pytest.mark.parametrize("seed", range(N_RANDOM))
def test_oracle_reproduces_random_golden(seed):
    """tests/util.random_case geometries: the oracle against outputs of the unmodified reference
    (tests/golden/golden_random.npz), predictions after every step and every state array at the end."""
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "golden_random.npz"))
    preds, keys, digests = run_random_case(OracleHierarchy, seed)
    assert np.array_equal(preds, g["r%d/pred" % seed])
    assert list(keys) == list(g["r%d/keys" % seed])
    assert list(digests) == list(g["r%d/digests" % seed])


# MSE over the last 100 steps of (getPrediction before the step - input)^2
@pytest.mark.parametrize("variant,steps,mse", [(cf.KWINNERS, 1000, 0.028279), (cf.ITERATIVE, 300, 0.003827),
                                               (cf.NEO_HIERARCHY, 300, 0.011272)])
def test_survey_learning_curve_anchor(variant, steps, mse):
    cfg, ld = cf.config_c1(variant)
    o = OracleHierarchy(cfg, ld, 1234)
    errs = []
    for t in range(steps):
        x = c1_frame(t)
        errs.append(float(np.mean((o.get_predictions() - x) ** 2)))
        o.set_inputs(x)
        o.step()
    assert abs(np.mean(errs[-100:]) - mse) < 5e-6
    if variant == cf.KWINNERS:  # "last-step winners per layer 1,1,1"
        assert [int(o.get_array("HID_STATE", l).sum()) for l in range(3)] == [1, 1, 1]
