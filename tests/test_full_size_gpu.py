"""Parity at BASELINE.json's FULL sizes (the small-case tests are in test_parity_gpu.py).

  C3  4096 neo::Agent hierarchies in one engine: sampled agents equal their own single-agent
      oracle runs bit for bit (agents never interact, so the rest of the batch is free to differ)
  C4  the large six-layer hierarchy (input 128x128, every radius 6): 50 steps of the k-winners and 5
      of the iterative variant against the oracle, every state array equal; the VideoTest
      configuration (input feedback radius 16) trained and free-run
  C5  the largest single layer the reference (16-bit indices) can hold, 256x256, and the full
      1024x1024 against the oracle; and a size-independent property at 1024x1024 -- splitting the
      layer into row strips changes nothing
All comparisons are exact (np.array_equal)."""
import numpy as np
import pytest

from bidinet_b200 import Engine, StripGroup
from bidinet_b200 import config as cf
from oracle.pyoracle import OracleHierarchy

from util import diff_states

pytestmark = pytest.mark.gpu


def test_c3_sampled_agents_of_the_full_batch():
    cfg, ld = cf.config_c2()
    A, sample, steps = 4096, (0, 1, 2047, 4095), 3
    eng = Engine(cfg, ld, n_agents=A, seed=1234)
    orcs = {a: OracleHierarchy(cfg, ld, 1234 + a) for a in sample}
    rng = np.random.RandomState(5)
    n = eng.n_noise
    for t in range(steps):
        x = rng.rand(A, 64).astype(np.float32)
        r = rng.randn(A).astype(np.float32)
        u, v = rng.rand(A, n).astype(np.float32), (0.05 * rng.randn(A, n)).astype(np.float32)
        for a, o in orcs.items():
            u[a], v[a] = o.peek_noise()
            o.set_inputs(x[a])
            o.step(float(r[a]))
        eng.set_inputs(x)
        eng.step(rewards=r, noise=(u, v))
    pred = eng.get_predictions()
    for a, o in orcs.items():
        assert np.array_equal(o.get_predictions(), pred[a]), a
        assert diff_states(o.state(), eng.state(agent=a)) == [], a


def _bars(t, n=128, period=16):
    x = np.arange(n)[None, :] + np.zeros((n, 1), dtype=int)
    return (((x + t) % period) < period // 2).astype(np.float32).ravel()


@pytest.mark.parametrize("variant,steps", [(cf.KWINNERS, 50), (cf.ITERATIVE, 5)])
def test_c4_steps(variant, steps):
    """C4 at full size: 50 steps of the k-winners variant (the winners' learning rule and the prediction delta rule
    have long since left their initial regime), 5 of the iterative one (35 coder iterations each); predictions
    after every step, every state array at the end."""
    cfg, ld = cf.config_c4(variant)
    orc = OracleHierarchy(cfg, ld, 1234)
    eng = Engine(cfg, ld, n_agents=1, seed=1234)   # the engine replays createRandom itself
    assert diff_states(orc.state(), eng.state()) == []
    for t in range(steps):
        x = _bars(t)
        orc.set_inputs(x)
        eng.set_inputs(x)
        orc.step()
        eng.step()
        assert np.array_equal(orc.get_predictions(), eng.get_predictions()[0]), t
    bad = diff_states(orc.state(), eng.state())
    assert bad == [], bad[:6]
    if variant == cf.KWINNERS:   # the run learned: weights of winners moved away from their initial values
        fresh = OracleHierarchy(cfg, ld, 1234)
        assert not np.array_equal(fresh.get_array("FF_W", 0), orc.get_array("FF_W", 0))
        assert not np.array_equal(fresh.get_array("PRED_W", 2), orc.get_array("PRED_W", 2))


def test_videotest_configuration_trains_then_free_runs():
    """VideoTest.cpp:52-65,136-139,180-186: in 128x128, layers 32^2/24^2/16^2, input feedback radius 16 (33x33
    windows into layer 0 for every one of the 16384 input nodes); a few training steps, then the free run on its
    own predictions with learn = false -- fed back through setInput(x, y), which indexes with the HIDDEN width
    (sdr/IPredictiveRSDR.h:115-117: the caller's loop writes input[x + 32*y], x outer / y inner)."""
    cfg, ld = cf.config_videotest()
    orc = OracleHierarchy(cfg, ld, 1234)
    eng = Engine(cfg, ld, n_agents=1, seed=1234)
    assert eng.array_len("FB_W", 3) == 9821956           # N_infb of the VideoTest row
    assert diff_states(orc.state(), eng.state()) == []
    hid_w = ld[0].width
    xs, ys = np.meshgrid(np.arange(128), np.arange(128), indexing="ij")   # x outer, y inner
    dst, src = (xs + hid_w * ys).ravel(), (xs + 128 * ys).ravel()
    x = None
    for t in range(8):
        train = t < 4
        if train:
            x = _bars(t)
        else:               # prsdr.setInput(x, y, prsdr.getPrediction(x, y)) for every x, y: later writes win
            pred = orc.get_predictions()
            x = x.copy()
            x[dst] = pred[src]
        orc.set_inputs(x)
        eng.set_inputs(x)
        orc.step(0.0, train)
        eng.step(learn=train)
        assert np.array_equal(orc.get_predictions(), eng.get_predictions()[0]), t
    bad = diff_states(orc.state(), eng.state())
    assert bad == [], bad[:6]


def _sparse(t, n, density=0.05, period=8):
    return (np.random.RandomState(7 + t % period).rand(n * n) < density).astype(np.float32)


def test_c5_256_grid_kernels():
    n = 256
    cfg, ld = cf.config_c5(n)
    orc = OracleHierarchy(cfg, ld, 1234)
    eng = Engine(cfg, ld, n_agents=1, seed=1234)
    for t in range(3):
        x = _sparse(t, n)
        orc.set_inputs(x)
        eng.set_inputs(x)
        orc.step()
        eng.step()
    assert np.array_equal(orc.get_predictions(), eng.get_predictions()[0])
    for key in (("HID_STATE", 0), ("HID_ACTIVATION", 0), ("VIS_RECON", 0), ("P_STATE", 0), ("FF_W", 0), ("REC_W", 0), ("PRED_W", 0),
                ("HID_THRESHOLD", 0)):
        assert np.array_equal(orc.get_array(*key), eng.get_array(*key)), key
    assert orc.get_array("HID_STATE", 0).sum() > 100  # winners existed, so the learning rule ran


def test_c5_1024_against_the_oracle():
    """C5 at its full size, unsplit, against the oracle (the reference itself cannot index this layer: 16-bit
    connection indices; the oracle is the same restatement that is pinned to the reference at every size the
    reference can hold).  Two steps, learning on: unit planes and every weight list equal."""
    n = 1024
    cfg, ld = cf.config_c5(n)
    orc = OracleHierarchy(cfg, ld, 1234)
    eng = Engine(cfg, ld, n_agents=1, seed=1234)
    for t in range(2):
        x = _sparse(t, n)
        orc.set_inputs(x)
        eng.set_inputs(x)
        orc.step()
        eng.step()
        assert np.array_equal(orc.get_predictions(), eng.get_predictions()[0]), t
    for key in (("HID_STATE", 0), ("HID_ACTIVATION", 0), ("VIS_RECON", 0), ("HID_RECON", 0), ("P_STATE", 0), ("P_ACTIVATION", 0),
                ("P_BASELINE", 0), ("HID_THRESHOLD", 0), ("FF_W", 0), ("REC_W", 0), ("PRED_W", 0)):
        assert np.array_equal(orc.get_array(*key), eng.get_array(*key)), key
    assert orc.get_array("HID_STATE", 0).sum() > 1000
    eng.close()
    orc.close()


def test_c5_full_size_split_equals_unsplit():
    """1024x1024, every radius 6: two row strips (same-process transport, one device) against the unsplit
    engine from the same createRandom stream -- a property that needs no CPU run of a layer this size."""
    from bidinet_b200 import assemble_rows
    n = 1024
    cfg, ld = cf.config_c5(n)
    whole = Engine(cfg, ld, n_agents=1, seed=1234)
    parts = StripGroup(cfg, ld, devices=[0, 0], seed=1234)
    for t in range(3):
        x = _sparse(t, n)
        whole.set_inputs(x)
        parts.set_inputs(x)
        whole.step()
        parts.step()
    assert np.array_equal(whole.get_predictions(), parts.get_predictions())
    for key, span in ((("HID_STATE", 0), "OWN_HID"), (("HID_ACTIVATION", 0), "OWN_HID"), (("VIS_RECON", 0), "OWN_VIS"),
                      (("HID_RECON", 0), "OWN_HID"), (("P_STATE", 0), "OWN_HID"), (("P_ACTIVATION", 0), "OWN_HID"),
                      (("HID_THRESHOLD", 0), "OWN_HID"), (("P_BASELINE", 0), "OWN_HID")):
        got = assemble_rows([e.get_array(*key) for e in parts.parts], [e.spans[span] for e in parts.parts], n)
        assert np.array_equal(whole.get_array(*key), got), key
    assert whole.get_array("HID_STATE", 0).sum() > 1000
    parts.close()
    whole.close()


@pytest.mark.parametrize("geom", ["ragged_r3", "wide_r5_norec"])
def test_persistent_coder_of_a_single_iterative_agent(geom):
    """bidi_persist.cuh: the whole iterative coder of a large layer of ONE agent as one cooperative launch (weights of a
    tile of units resident in shared memory for all 35 iterations, grid barrier between sweep and reconstruction).
    Geometries whose grids are not multiples of the tile sizes, with ratios above and below 1, with and without a
    recurrent family; the second layer is too small for it and runs the per-phase kernels.  Exact against the oracle,
    every array, and equal to the per-phase kernels of the same engine build (option persist_coder = 0)."""
    if geom == "ragged_r3":
        cfg, ld = cf.make_config(cf.ITERATIVE, 40, 28, [dict(width=27, height=21), dict(width=14, height=12)], r_infb=5)
        n_in = 40 * 28
    else:
        layers = [dict(width=36, height=30, r_ff=5, r_lat=5, r_rec=-1, r_pred=2, r_fb=2), dict(width=20, height=24, r_ff=4)]
        cfg, ld = cf.make_config(cf.ITERATIVE, 30, 33, layers, r_infb=3)
        n_in = 30 * 33
    orc = OracleHierarchy(cfg, ld, 77)
    eng = Engine(cfg, ld, n_agents=1, seed=77)
    ref = Engine(cfg, ld, n_agents=1, seed=77)
    ref.set_option("persist_coder", 0)
    assert eng.launch_count == ref.launch_count == 0
    rng = np.random.RandomState(3)
    for t in range(6):
        x = (rng.rand(n_in) < 0.3).astype(np.float32) * rng.rand(n_in).astype(np.float32)
        for e in (orc, eng, ref):
            e.set_inputs(x)
        orc.step()
        eng.step()
        ref.step()
        assert np.array_equal(orc.get_predictions(), eng.get_predictions()[0]), t
    assert eng.launch_count < ref.launch_count, "the persistent kernel was not the path taken"
    assert diff_states(orc.state(), eng.state()) == []
    assert diff_states(ref.state(), eng.state()) == []
    for e in (eng, ref):
        e.close()
