"""The batched headless environment on the device (bidi_env_*, BIDI_ENV_RUNNER: a planar stand-in for the Box2D
quadruped of runner/Runner.cpp) closing the RunnerMain loop (RunnerMain.cpp:204-251), against the oracle hierarchy
driven by the environment's CPU twin (oracle/runner_env_oracle.c): predictions, every state array and the
environment's own state equal after every step."""
import numpy as np
import pytest

from bidinet_b200 import Engine
from bidinet_b200 import config as cf
from oracle.pyoracle import OracleHierarchy, RunnerEnv

from util import diff_states

F = np.float32


def clamp01(v):
    return np.minimum(F(1), np.maximum(F(0), v))


def test_cpu_twin_walks_and_is_deterministic():
    """(CPU) a trot-like open-loop gait moves the body forward; two runs agree bit for bit; limits hold."""
    def run():
        e = RunnerEnv()
        xs = []
        for t in range(240):
            ph = 2 * np.pi * t / 30.0
            a = np.full(10, 0.5, np.float32)
            for hip, knee, off in ((0, 1, 0.0), (3, 4, np.pi), (6, 7, np.pi), (8, 9, 0.0)):
                a[hip] = 0.5 - 0.4 * np.cos(ph + off)          # sweep the leg
                a[knee] = 0.9 if np.sin(ph + off) > 0 else 0.3   # extended while it sweeps backwards
            e.motor_update(a)
            xs.append(e.dump())
        return np.stack(xs)
    a, b = run(), run()
    assert np.array_equal(a, b)
    assert np.isfinite(a).all()
    assert a[:, 11:15].min() == 0.0 and a[:, 11:15].max() == 1.0       # feet lift and land
    assert abs(a[-1, 17]) > 0.05                                       # the body went somewhere
    assert np.all(np.abs(a[:, 10]) < 0.5)                              # and stayed upright
    lo = np.array([-0.6, -1.0, -0.4, -0.6, -1.0, -0.4, -0.6, -1.0, -0.6, -1.0], np.float32)
    hi = np.array([0.6, 0.2, 0.4, 0.6, 0.2, 0.4, 0.6, 0.2, 0.6, 0.2], np.float32)
    assert np.all(a[:, :10] >= lo) and np.all(a[:, :10] <= hi)


@pytest.mark.gpu
def test_closed_loop_equals_oracle_with_cpu_twin():
    cfg, ld = cf.config_c2()
    A, steps = 3, 80
    eng = Engine(cfg, ld, n_agents=A, seed=1234)
    eng.env_attach()
    orcs = [OracleHierarchy(cfg, ld, 1234 + a) for a in range(A)]
    envs = [RunnerEnv() for _ in range(A)]
    rng = np.random.RandomState(2)
    assert np.array_equal(eng.env_state(), np.stack([e.dump() for e in envs]))
    pred = np.zeros((A, 64), np.float32)
    moved = 0.0
    for t in range(steps):
        z = (0.05 * rng.randn(A, 10)).astype(np.float32)
        nz = [o.peek_noise() for o in orcs]
        noise = (np.stack([n[0] for n in nz]), np.stack([n[1] for n in nz]))
        for a in range(A):
            x = orcs[a].get_array("VIS_INPUT", 0).copy()      # inputs the loop does not write keep their value
            x[:15] = envs[a].state()
            x[15:19] = envs[a].clocks()
            x[20:30] = clamp01(pred[a, 20:30] * F(0.5) + F(0.5) + z[a])
            r = envs[a].reward()
            orcs[a].set_inputs(x)
            orcs[a].step(r, t != 7)
            pred[a] = orcs[a].get_predictions()
            envs[a].motor_update(clamp01(pred[a, 20:30] * F(0.5) + F(0.5)))
        eng.env_step(noise=noise, env_noise=z, learn=t != 7)
        assert np.array_equal(pred, eng.get_predictions()), t
        want = np.stack([e.dump() for e in envs])
        assert np.array_equal(want, eng.env_state()), t
        moved = max(moved, float(np.abs(want[:, 15]).max()))
    for a in range(A):
        assert diff_states(orcs[a].state(), eng.state(agent=a)) == [], a
    assert moved > 0.0   # the agents' motor commands did move the bodies


@pytest.mark.gpu
def test_closed_loop_with_engine_noise_runs_a_large_batch():
    cfg, ld = cf.config_c2()
    A = 512
    eng = Engine(cfg, ld, n_agents=A, seed=5)
    eng.env_attach()
    for t in range(60):
        eng.env_step()
    st = eng.env_state()
    assert np.isfinite(st).all() and np.isfinite(eng.get_predictions()).all()
    assert np.abs(st[:, 17]).max() > 0 and len(np.unique(st[:, 17])) > A // 2   # every agent lives in its own world
    eng.env_reset()
    assert np.array_equal(eng.env_state()[0], eng.env_state()[A - 1])
