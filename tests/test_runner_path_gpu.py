"""The path bench.py times, held to the oracle: device-resident frames (bidi_load_frames), action feedback and
reward on the device (bidi_set_feedback -> k_feedback), bidi_step_frame -- against the host-driven calls
(bidi_set_inputs + bidi_step) and against the oracle stepping the headless RunnerMain loop
(RunnerMain.cpp:235-251, neo/Agent.cpp:141-284) with the environment's own noise switched off and the
columns' exploration draws supplied.  Exact (np.array_equal) on every state array.

Also here: the steady state the bench is timed in (600 free-running steps, past its 400-step burn-in, where
every column cell spikes and every learning rule does its full work), and the distribution of the engine's
own counter-based exploration and feedback noise (k_noise, k_feedback), which the timed runs use."""
import numpy as np
import pytest

from bidinet_b200 import Engine
from bidinet_b200 import config as cf
from bidinet_b200.sharding import FB_IDX, base_frames
from oracle.pyoracle import OracleHierarchy

from util import diff_states

pytestmark = pytest.mark.gpu
F = np.float32


def clamp01(v):
    return np.minimum(F(1), np.maximum(F(0), v))


def host_feedback(pred, x):
    """The caller's half of the runner loop in fp32, operation for operation what k_feedback (noise_std = 0) and
    tools/runner_env.c do: action = prediction*0.5 + 0.5 fed back clamped, reward = mean(clamp01(action)) - 0.5."""
    act = pred[FB_IDX] * F(0.5) + F(0.5)
    s = F(0)
    for k in range(len(FB_IDX)):
        s = F(s + clamp01(act[k]))
    x[FB_IDX] = clamp01(act)
    return F(F(s / F(len(FB_IDX))) - F(0.5))


def exploration_draws(rng, shape, brk=0.01, std=0.05):
    """(u, v) as neo/Column.cpp:126-131 draws them: u uniform; v uniform where u < break chance, else N(0, std)."""
    u = rng.rand(*shape).astype(np.float32)
    v = np.where(u < brk, rng.rand(*shape), std * rng.randn(*shape)).astype(np.float32)
    return u, v


def test_frame_feedback_path_equals_host_path_and_oracle():
    """60 steps, 3 agents: engine A on the device-resident path, engine B on host buffers, one oracle per agent."""
    cfg, ld = cf.config_c2()
    A, steps, T = 3, 60, 16
    dev = Engine(cfg, ld, n_agents=A, seed=1234)
    host = Engine(cfg, ld, n_agents=A, seed=1234)
    orcs = [OracleHierarchy(cfg, ld, 1234 + a) for a in range(A)]
    frames = base_frames(T, A, 0)
    u, v = exploration_draws(np.random.RandomState(11), (steps, A, dev.n_noise))
    assert (u < 0.01).sum() > 100  # exploration breaks are exercised
    dev.load_frames(frames)
    dev.load_frame_noise(u, v)
    dev.set_feedback(FB_IDX, FB_IDX, 0.5, 0.5, 0.0, True)
    pred = np.zeros((A, 64), np.float32)
    for t in range(steps):
        x = frames[t % T].copy()
        r = np.zeros(A, np.float32)
        for a in range(A):
            r[a] = host_feedback(pred[a], x[a])
            orcs[a].set_inputs(x[a])
            orcs[a].step_noise(float(r[a]), u[t, a], v[t, a], learn=t != 20)
        host.set_inputs(x)
        host.step(rewards=r, noise=(u[t], v[t]), learn=t != 20)
        dev.step_frame(t, learn=t != 20)
        pred = host.get_predictions()
        assert np.array_equal(pred, dev.get_predictions()), t
        for a in range(A):
            assert np.array_equal(orcs[a].get_predictions(), pred[a]), (t, a)
        if t % 10 == 9 or t == steps - 1:
            for a in range(A):
                so = orcs[a].state()
                assert diff_states(so, dev.state(agent=a)) == [], (t, a)
                assert diff_states(so, host.state(agent=a)) == [], (t, a)
    assert np.array_equal(dev.get_column_actions(), host.get_column_actions())
    dev.load_frame_noise(None, None)   # back to the engine's own generator: the two engines now part ways
    dev.step_frame(steps)
    assert np.isfinite(dev.get_predictions()).all()


@pytest.mark.parametrize("agents", [1, 4096])
def test_steady_state_600_steps(agents):
    """The regime bench.py times (it burns in 400 steps first): 600 free-running steps on the benchmarked path,
    sampled agents equal to their own oracle on every state array, and the run did reach the regime in which
    (nearly) every column cell has state > 0 -- the column kernel's parking batches, all-cells-live learning and
    long-row tails only run hot there."""
    cfg, ld = cf.config_c2()
    steps, T, NT = 600, 64, 8
    sample = [0] if agents == 1 else [0, 1, 2, 777, 2047, 2048, 4000, 4095]
    eng = Engine(cfg, ld, n_agents=agents, seed=1234)
    orcs = {a: OracleHierarchy(cfg, ld, 1234 + a) for a in sample}
    frames = base_frames(T, agents, 0)
    u, v = exploration_draws(np.random.RandomState(3), (NT, agents, eng.n_noise))
    eng.load_frames(frames)
    eng.load_frame_noise(u, v)
    eng.set_feedback(FB_IDX, FB_IDX, 0.5, 0.5, 0.0, True)
    pred = {a: np.zeros(64, np.float32) for a in sample}
    for t in range(steps):
        eng.step_frame(t)
        for a, o in orcs.items():
            x = frames[t % T, a].copy()
            r = host_feedback(pred[a], x)
            o.set_inputs(x)
            o.step_noise(float(r), u[t % NT, a], v[t % NT, a])
            pred[a] = o.get_predictions()
        if t % 100 == 99:
            got = eng.get_predictions()
            for a in sample:
                assert np.array_equal(pred[a], got[a]), (t, a)
    for a, o in orcs.items():
        so = o.state()
        assert diff_states(so, eng.state(agent=a)) == [], a
        live = sum(int((so[("COL_STATE", l)] > 0).sum()) for l in range(cfg.n_layers + 1))
        cells = sum(so[("COL_STATE", l)].size for l in range(cfg.n_layers + 1))
        assert cells == 2304 and live >= 0.95 * cells, (a, live)


def _ks_normal(x, std):
    """Kolmogorov-Smirnov distance of a sample from N(0, std)."""
    from math import erf, sqrt
    x = np.sort(np.asarray(x, dtype=np.float64))
    cdf = 0.5 * (1.0 + np.vectorize(erf)(x / (std * sqrt(2.0))))
    n = x.size
    return max(np.max(np.arange(1, n + 1) / n - cdf), np.max(cdf - np.arange(0, n) / n))


def test_device_exploration_noise_distribution():
    """k_noise over > 10^6 draws: u uniform on [0,1), break frequency 0.01 within 3 sigma, v ~ N(0, 0.05) off the
    breaks (KS test) and uniform on them, no correlation across agents, columns or steps, and different shards
    (bidi_set_noise_stream) draw different streams."""
    cfg, ld = cf.config_c2()
    A = 1024
    eng = Engine(cfg, ld, n_agents=A, seed=1)
    x = np.zeros((A, 64), np.float32)
    us, vs = [], []
    for t in range(3):
        eng.set_inputs(x)
        eng.step(rewards=np.zeros(A, np.float32))
        u, v = eng.get_noise()
        us.append(u)
        vs.append(v)
    u, v = np.stack(us), np.stack(vs)           # [step][agent][column*3]
    n = u.size
    assert n > 1.3e6
    assert 0.0 <= u.min() and u.max() < 1.0
    brk = u < 0.01
    p = 0.01
    assert abs(brk.mean() - p) < 3.0 * np.sqrt(p * (1 - p) / n), brk.mean()
    assert abs(u.mean() - 0.5) < 4.0 / np.sqrt(12.0 * n)
    hist = np.histogram(u, bins=20, range=(0, 1))[0] / n
    assert np.abs(hist - 0.05).max() < 5 * np.sqrt(0.05 * 0.95 / n)
    normal = v[~brk].astype(np.float64)
    sub = normal[:: max(1, normal.size // 200000)]
    assert _ks_normal(sub, 0.05) < 1.63 / np.sqrt(sub.size)      # 1 % critical value
    assert abs(normal.std() - 0.05) < 0.05 * 4.0 / np.sqrt(2 * normal.size)
    assert abs(normal.mean()) < 0.05 * 4.0 / np.sqrt(normal.size)
    on = v[brk]
    assert on.size > 10000 and 0.0 <= on.min() and on.max() < 1.0 and abs(on.mean() - 0.5) < 4.0 / np.sqrt(12.0 * on.size)

    def corr(a, b):
        a, b = a.ravel().astype(np.float64), b.ravel().astype(np.float64)
        return abs(np.corrcoef(a, b)[0, 1])

    lim = 4.5 / np.sqrt(u[0, 0].size * (A - 1))
    assert corr(u[0, :-1], u[0, 1:]) < lim and corr(v[0, :-1], v[0, 1:]) < lim         # neighbouring agents
    assert corr(u[0], u[1]) < lim and corr(v[1], v[2]) < lim                           # consecutive steps
    assert corr(u[0, :, :-1], u[0, :, 1:]) < lim                                       # neighbouring draws
    assert corr(u[~brk], v[~brk]) < 4.5 / np.sqrt((~brk).sum())                        # u and the normal drawn after it
    # a second shard of the same population: same seed, agents offset by A
    other = Engine(cfg, ld, n_agents=A, seed=1)
    other.set_noise_stream(0x5DEECE66D, A)
    other.set_inputs(x)
    other.step(rewards=np.zeros(A, np.float32))
    u2, _ = other.get_noise()
    assert not np.array_equal(u2, u[0]) and corr(u2, u[0]) < lim
    same = Engine(cfg, ld, n_agents=A, seed=1)   # and the default stream is reproducible
    same.set_inputs(x)
    same.step(rewards=np.zeros(A, np.float32))
    assert np.array_equal(same.get_noise()[0], u[0])


def test_device_feedback_noise():
    """k_feedback with its noise on, 1.2 * 10^6 draws: the fed-back inputs are clamp01(action + N(0, 0.05)) around
    the action the previous prediction encodes (the reward arithmetic is the noiseless one the parity tests above
    hold exactly); draws are independent across agents and steps."""
    cfg, ld = cf.config_c2()
    A, steps = 4096, 30
    eng = Engine(cfg, ld, n_agents=A, seed=5)
    eng.load_frames(base_frames(8, A, 0))
    eng.set_feedback(FB_IDX, FB_IDX, 0.5, 0.5, 0.05, True)
    pred = eng.get_predictions()
    draws = []
    for t in range(steps):
        eng.step_frame(t)
        fed = eng.get_inputs()[:, FB_IDX]               # the inputs the step just consumed
        act = pred[:, FB_IDX] * F(0.5) + F(0.5)
        assert fed.min() >= 0.0 and fed.max() <= 1.0
        draws.append(np.where((fed > 0) & (fed < 1) & (act > 0.25) & (act < 0.75), fed - act, np.nan))
        pred = eng.get_predictions()
    d = np.stack(draws).astype(np.float64)
    ok = np.isfinite(d)
    assert ok.sum() >= 1.0e6
    z = d[ok]
    sub = z[:: z.size // 200000]
    assert _ks_normal(sub, 0.05) < 1.63 / np.sqrt(sub.size)
    assert abs(z.std() - 0.05) < 0.05 * 4 / np.sqrt(2 * z.size) and abs(z.mean()) < 0.05 * 4 / np.sqrt(z.size)

    def corr(a, b):
        m = np.isfinite(a) & np.isfinite(b)
        return abs(np.corrcoef(a[m], b[m])[0, 1]), 4.5 / np.sqrt(m.sum())

    c, lim = corr(d[0].ravel(), d[1].ravel())
    assert c < lim                                      # steps are independent
    c, lim = corr(d[0, :-1].ravel(), d[0, 1:].ravel())
    assert c < lim                                      # agents are independent
    c, lim = corr(d[0, :, :-1].ravel(), d[0, :, 1:].ravel())
    assert c < lim                                      # taps are independent
