"""CUDA engine (through the C ABI) against the oracle on the same seeded inputs.

Bar: every state array -- activations, spikes, k-winner sets, predictions, weights,
traces, column state -- EQUAL (np.array_equal, i.e. bit-exact up to the sign of
zero) after every step.  The kernels keep the reference's fp32 operation order and
are built without FMA contraction, and bidi_math.h restates the C library's expf,
so no tolerance is needed."""
import numpy as np
import pytest

from bidinet_b200 import Engine
from bidinet_b200.engine import BidiError
from bidinet_b200 import config as cf
from oracle.pyoracle import OracleHierarchy

from util import (N_CSRL_RANDOM, RANDOM_STEPS, csrl_random_case, diff_states, random_case, random_schedule, small_cases,
                  state_digests)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", small_cases(), ids=lambda c: c[0])
def test_upload_download_roundtrip(case):
    """bidi_set_array / bidi_get_array: serialised reference order <-> slot planes."""
    name, cfg, ld, _, _ = case
    orc = OracleHierarchy(cfg, ld, 99)
    eng = Engine(cfg, ld, n_agents=2)
    st = orc.state()
    rng = np.random.RandomState(1)
    st = {k: rng.randn(*v.shape).astype(np.float32) for k, v in st.items()}
    eng.load_state(st, agent=1)
    assert diff_states(st, eng.state(agent=1)) == []
    zero = eng.state(agent=0)  # the other agent is untouched
    assert all(not np.any(v) or k[0] in ("HID_THRESHOLD", "COL_THRESHOLD") for k, v in zero.items())


@pytest.mark.parametrize("case", small_cases(), ids=lambda c: c[0])
def test_init_random_matches_oracle(case):
    """bidi_init_random replays createRandom's std::mt19937 draw order."""
    name, cfg, ld, _, _ = case
    eng = Engine(cfg, ld, n_agents=3, seed=1234)
    for a in (0, 2):
        orc = OracleHierarchy(cfg, ld, 1234 + a)
        assert diff_states(orc.state(), eng.state(agent=a)) == [], "agent %d" % a


@pytest.mark.parametrize("case", small_cases(), ids=lambda c: c[0])
def test_init_random_from_the_callers_generator(case):
    """bidi_init_random_generator: createRandom's draws taken from a caller-side std::mt19937 handed over as text (the
    deep::CSRL facade's createRandom).  A generator seeded like bidi_init_random's gives the same weights; the state
    that comes back is a valid generator that has moved on; a malformed state is refused."""
    name, cfg, ld, _, _ = case
    seeded = Engine(cfg, ld, n_agents=1, seed=1234)
    eng = Engine(cfg, ld, n_agents=2)
    key = np.random.RandomState(1234).get_state()[1]       # init_genrand(1234): the state of std::mt19937(1234)
    text = " ".join(str(int(k)) for k in key) + " 624"
    after = eng.init_random_generator(text, agent=1)
    assert diff_states(seeded.state(agent=0), eng.state(agent=1)) == []
    words = after.split()
    assert len(words) == 625 and after != text
    eng.init_random_generator(after, agent=0)               # the advanced state is accepted again
    assert diff_states(eng.state(agent=0), eng.state(agent=1)) != []
    with pytest.raises(BidiError):
        eng.init_random_generator("1 2 3 not a generator")


@pytest.mark.parametrize("path", ["default", "fused_general", "per_phase_tiles", "per_phase_threads", "per_phase_grid"])
@pytest.mark.parametrize("case", small_cases(), ids=lambda c: c[0])
def test_free_running_parity(case, path):
    """N steps from an identical state, no re-synchronisation: exact after every step.
    Every engine path is held to it: the default (a CUDA graph; fused per-agent coder kernel where the
    layer fits in shared memory -- its dense-row form where the windows span the grid --
    and cooperative tile kernels for small launches), the general fused coder kernel
    everywhere, one
    kernel per phase with the tile kernels (large single agents), one kernel per
    phase with one thread per unit (large agent batches), and the 2-D grid-tile kernels
    (very large k-winners layers)."""
    name, cfg, ld, frame, reward = case
    orc = OracleHierarchy(cfg, ld, 1234)
    eng = Engine(cfg, ld, n_agents=1)
    if path == "fused_general":
        eng.set_option("dense_coder", False)
    elif path != "default":
        eng.set_option("force_phase_kernels", True)
    if path == "per_phase_threads":
        eng.set_option("tile_kernels", False)
        eng.set_option("grid_kernels", 0)
    if path == "per_phase_grid":  # the 2-D tile kernels for large layers, forced onto every layer
        if cfg.variant == cf.NEO_AGENT:
            pytest.skip("no grid kernels for the column hierarchy")
        eng.set_option("grid_kernels", 2)
    eng.load_state(orc.state())
    steps = 25
    for t in range(steps):
        x = frame(t)
        orc.set_inputs(x)
        eng.set_inputs(x)
        noise = orc.peek_noise() if orc.num_noise() else None
        learn = t != 9
        orc.step(reward(t), learn)
        eng.step(rewards=[reward(t)], noise=noise, learn=learn)
        bad = diff_states(orc.state(), eng.state())
        assert bad == [], "step %d: %s" % (t, bad[:6])
        assert np.array_equal(orc.get_predictions(), eng.get_predictions()[0])


@pytest.mark.parametrize("path", ["default", "fused_general", "per_phase_tiles", "per_phase_threads", "per_phase_grid"])
@pytest.mark.parametrize("case", small_cases(), ids=lambda c: c[0])
def test_batched_free_running_parity(case, path):
    """Three agents in one engine, every variant on every engine path: each agent against its OWN oracle (own
    weights, own inputs, own reward) after every step -- the agent decomposition of the tile kernels (tiles_of),
    the 2-D grid tiles, the Q network's one-CTA-per-agent launch and the pitched copy of bidi_get_actions."""
    name, cfg, ld, frame, reward = case
    A, steps = 3, 10
    eng = Engine(cfg, ld, n_agents=A, seed=1234)
    orcs = [OracleHierarchy(cfg, ld, 1234 + a) for a in range(A)]
    if path == "fused_general":
        eng.set_option("dense_coder", False)
    elif path != "default":
        eng.set_option("force_phase_kernels", True)
    if path == "per_phase_threads":
        eng.set_option("tile_kernels", False)
        eng.set_option("grid_kernels", 0)
    if path == "per_phase_grid":
        if cfg.variant == cf.NEO_AGENT:
            pytest.skip("no grid kernels for the column hierarchy")
        eng.set_option("grid_kernels", 2)
    for a in range(A):
        assert diff_states(orcs[a].state(), eng.state(agent=a)) == [], a
    for t in range(steps):
        x = np.stack([frame(t + 3 * a) for a in range(A)])
        r = np.array([reward(t + a) for a in range(A)], np.float32)
        noise = None
        if orcs[0].num_noise():
            nz = [o.peek_noise() for o in orcs]
            noise = (np.stack([n[0] for n in nz]), np.stack([n[1] for n in nz]))
        learn = t != 4
        for a, o in enumerate(orcs):
            o.set_inputs(x[a])
            o.step(float(r[a]), learn)
        eng.set_inputs(x)
        eng.step(rewards=r, noise=noise, learn=learn)
        pred = eng.get_predictions()
        for a, o in enumerate(orcs):
            assert np.array_equal(o.get_predictions(), pred[a]), (t, a)
            bad = diff_states(o.state(), eng.state(agent=a))
            assert bad == [], "step %d agent %d: %s" % (t, a, bad[:6])
    if cfg.variant == cf.QPRSDR:
        acts = eng.get_actions()
        for a, o in enumerate(orcs):
            assert np.array_equal(acts[a], o.get_array("ACT_EXPL", 0)), a


def test_batched_agents_are_independent():
    """Agents in one engine evolve exactly like separate single-agent runs."""
    cfg, ld = cf.config_c2()
    A = 5
    eng = Engine(cfg, ld, n_agents=A, seed=1234)
    orcs = [OracleHierarchy(cfg, ld, 1234 + a) for a in range(A)]
    rng = np.random.RandomState(3)
    for t in range(6):
        x = rng.rand(A, 64).astype(np.float32)
        r = rng.randn(A).astype(np.float32)
        nz = [o.peek_noise() for o in orcs]
        for a, o in enumerate(orcs):
            o.set_inputs(x[a])
            o.step(float(r[a]))
        eng.set_inputs(x)
        eng.step(rewards=r, noise=(np.stack([n[0] for n in nz]), np.stack([n[1] for n in nz])))
        pred = eng.get_predictions()
        for a, o in enumerate(orcs):
            assert np.array_equal(o.get_predictions(), pred[a]), (t, a)
    for a in (0, A - 1):
        assert diff_states(orcs[a].state(), eng.state(agent=a)) == []


def test_host_buffers_in_place():
    """bidi_host_inputs / bidi_host_predictions: stepping through the engine's own page-locked buffers gives
    the same predictions as the copying calls."""
    cfg, ld = cf.config_c2()
    rng = np.random.RandomState(5)
    frames = rng.rand(6, 3, 64).astype(np.float32)
    a, b = Engine(cfg, ld, n_agents=3), Engine(cfg, ld, n_agents=3)
    for e in (a, b):
        e.init_random(77)
    x, pred = b.host_inputs(), b.host_predictions()
    assert x.shape == (3, 64) and pred.shape == (3, 64)
    for t in range(6):
        a.set_inputs(frames[t])
        a.step(rewards=[0.1, 0.2, 0.3])
        np.copyto(x, frames[t])
        b.set_inputs(x)
        b.step(rewards=[0.1, 0.2, 0.3])
        b.wait_inputs()
        x[:] = -1.0  # the copy has read the buffer: rewriting it now must not matter
        out = b.get_predictions(out=pred)
        assert out is pred
        assert np.array_equal(a.get_predictions(), pred), "step %d" % t
    a.close()
    b.close()


def test_device_noise_mode_runs_and_explores():
    """noise=None: the engine's own generator; actions stay in [0,1] and differ between agents."""
    cfg, ld = cf.config_c2()
    eng = Engine(cfg, ld, n_agents=4, seed=1)
    x = np.random.RandomState(0).rand(4, 64).astype(np.float32)
    for t in range(3):
        eng.set_inputs(x)
        eng.step(rewards=np.zeros(4, np.float32))
    act = eng.get_column_actions()
    assert np.all(act >= 0) and np.all(act <= 1) and np.isfinite(eng.get_predictions()).all()
    assert not np.array_equal(act[0], act[1])


def test_snapshot_resume_is_bit_identical(tmp_path):
    """bidi_save_snapshot / bidi_load_snapshot: a run resumed from a file continues exactly like the
    uninterrupted one (device noise included: the generator's step counter is part of the snapshot);
    a snapshot of a different configuration is refused."""
    import bidinet_b200
    cfg, ld = cf.config_c2()
    x = np.random.RandomState(0).rand(8, 3, 64).astype(np.float32)
    a = Engine(cfg, ld, n_agents=3, seed=11)
    for t in range(3):
        a.set_inputs(x[t])
        a.step(rewards=np.full(3, 0.1 * t, np.float32))
    path = str(tmp_path / "c2.snap")
    a.save_snapshot(path)
    b = Engine(cfg, ld, n_agents=3)
    b.load_snapshot(path)
    for t in range(3, 6):
        for e in (a, b):
            e.set_inputs(x[t])
            e.step(rewards=np.full(3, 0.1 * t, np.float32))
    assert np.array_equal(a.get_predictions(), b.get_predictions())
    assert np.array_equal(a.get_column_actions(), b.get_column_actions())
    for agent in (0, 2):
        assert diff_states(a.state(agent), b.state(agent)) == []
    other = Engine(*cf.config_c1(cf.KWINNERS))
    with pytest.raises(bidinet_b200.BidiError, match="not a snapshot"):
        other.load_snapshot(path)


@pytest.mark.parametrize("path", ["default", "fused_general", "per_phase_tiles", "per_phase_threads", "per_phase_grid"])
@pytest.mark.parametrize("case", [c for c in small_cases() if c[1].variant == cf.NEO_HIERARCHY], ids=lambda c: c[0])
def test_step_generate_parity(case, path):
    """neo::PredictiveHierarchy::simStepGenerate (noise-driven coders, no learning) interleaved with ordinary
    steps: every state array equal after every step, the N(0,1) draws taken from the oracle's generator."""
    name, cfg, ld, frame, reward = case
    orc = OracleHierarchy(cfg, ld, 1234)
    eng = Engine(cfg, ld, n_agents=1)
    if path == "fused_general":
        eng.set_option("dense_coder", False)
    elif path != "default":
        eng.set_option("force_phase_kernels", True)
    if path == "per_phase_threads":
        eng.set_option("tile_kernels", False)
        eng.set_option("grid_kernels", 0)
    if path == "per_phase_grid":
        eng.set_option("grid_kernels", 2)
    eng.load_state(orc.state())
    generated = 0
    for t in range(16):
        x = frame(t)
        orc.set_inputs(x)
        eng.set_inputs(x)
        if t % 3 == 2:
            z = orc.peek_generate()
            orc.step_generate(0.3)
            eng.step_generate(0.3, normals=z)
            generated += 1
        else:
            orc.step(reward(t))
            eng.step()
        bad = diff_states(orc.state(), eng.state())
        assert bad == [], "step %d: %s" % (t, bad[:6])
        assert np.array_equal(orc.get_predictions(), eng.get_predictions()[0])
    assert generated == 5
    eng.step_generate(0.3)  # the engine's own generator: runs, stays finite
    assert np.isfinite(eng.get_predictions()).all()


@pytest.mark.parametrize("seed", range(40))
def test_random_geometries(seed):
    """Randomly drawn shapes (all five variants; 1x1 grids, up- and down-sampling ratios, windows from radius 0 to
    wider than the grid, 3..16-cell columns): exact parity with the oracle over 6 steps on the default engine path,
    and for the k-winners variant also on the 2-D grid-tile kernels.  The final state is also held against the
    digests of the unmodified reference's own run (tests/golden/golden_random.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_random.npz"))
    cfg, ld, frames = random_case(seed)
    paths = ["default"] + (["grid"] if cfg.variant == cf.KWINNERS else [])
    for path in paths:
        orc = OracleHierarchy(cfg, ld, 4321 + seed)
        eng = Engine(cfg, ld, n_agents=1)
        if path == "grid":
            eng.set_option("grid_kernels", 2)
        eng.load_state(orc.state())
        for t in range(RANDOM_STEPS):
            x = frames[t % len(frames)]
            orc.set_inputs(x)
            eng.set_inputs(x)
            noise = orc.peek_noise() if orc.num_noise() else None
            reward, learn = random_schedule(t)
            orc.step(reward, learn)
            eng.step(rewards=[reward], noise=noise, learn=learn)
            bad = diff_states(orc.state(), eng.state())
            assert bad == [], "seed %d path %s step %d: %s" % (seed, path, t, bad[:6])
        keys, digests = state_digests(eng.state())
        assert list(keys) == list(g["r%d/keys" % seed])
        assert list(digests) == list(g["r%d/digests" % seed]), "seed %d path %s: differs from the reference's golden digests" % (seed, path)
        eng.close()


@pytest.mark.parametrize("seed", range(N_CSRL_RANDOM))
def test_csrl_random_geometries(seed):
    """deep::CSRL on randomly drawn shapes (the oracle is pinned to the compiled reference on the same shapes by
    tests/test_oracle_vs_ref.py): every state array -- coders, prediction nodes with their traces, every SDRRL unit --
    equal after every step, two agents in the engine."""
    cfg, ld, frames = csrl_random_case(seed)
    eng = Engine(cfg, ld, n_agents=2, seed=99 + seed)
    orcs = [OracleHierarchy(cfg, ld, 99 + seed + a) for a in range(2)]
    for a in range(2):
        assert diff_states(orcs[a].state(), eng.state(agent=a)) == [], a
    for t in range(RANDOM_STEPS):
        x = np.stack([frames[(t + a) % len(frames)] for a in range(2)])
        nz = [o.peek_noise() for o in orcs]
        reward, learn = random_schedule(t)
        for a, o in enumerate(orcs):
            o.set_inputs(x[a])
            o.step(reward + a, learn)
        eng.set_inputs(x)
        eng.step(rewards=[reward, reward + 1], noise=(np.stack([n[0] for n in nz]), np.stack([n[1] for n in nz])), learn=learn)
        pred = eng.get_predictions()
        for a, o in enumerate(orcs):
            assert np.array_equal(o.get_predictions(), pred[a]), (seed, t, a)
            bad = diff_states(o.state(), eng.state(agent=a))
            assert bad == [], "seed %d step %d agent %d: %s" % (seed, t, a, bad[:6])
    eng.step(rewards=[0.0, 0.0])  # the engine's own exploration draws: runs, stays finite
    assert np.isfinite(eng.get_predictions()).all()
