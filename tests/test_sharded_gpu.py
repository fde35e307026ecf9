"""bidi_create_sharded: one process, several shards (devices), agents partitioned contiguously, every call fanned
out.  A sharded population must evolve exactly like the same agents in ONE engine and like separate single-agent
oracles: global agent g is created from std::mt19937(seed + g) whatever the number of shards, and the engines'
own exploration noise is keyed by the global agent index."""
import numpy as np
import pytest

from bidinet_b200 import Engine, ShardedEngine
from bidinet_b200 import config as cf
from bidinet_b200.sharding import FB_IDX, base_frames
from oracle.pyoracle import OracleHierarchy

from util import diff_states

pytestmark = pytest.mark.gpu


def _drive(eng_factory, steps=8):
    cfg, ld = cf.config_c2()
    A = 7
    whole = Engine(cfg, ld, n_agents=A, seed=1234)
    parts = eng_factory(cfg, ld, A)
    orcs = [OracleHierarchy(cfg, ld, 1234 + a) for a in range(A)]
    rng = np.random.RandomState(9)
    for t in range(steps):
        x = rng.rand(A, 64).astype(np.float32)
        r = rng.randn(A).astype(np.float32)
        nz = [o.peek_noise() for o in orcs]
        noise = (np.stack([n[0] for n in nz]), np.stack([n[1] for n in nz]))
        for a, o in enumerate(orcs):
            o.set_inputs(x[a])
            o.step(float(r[a]), t != 3)
        for e in (whole, parts):
            e.set_inputs(x)
            e.step(rewards=r, noise=noise, learn=t != 3)
        pw, pp = whole.get_predictions(), parts.get_predictions()
        assert np.array_equal(pw, pp), t
        for a, o in enumerate(orcs):
            assert np.array_equal(o.get_predictions(), pp[a]), (t, a)
    assert np.array_equal(whole.get_column_actions(), parts.get_column_actions())
    for a in range(A):
        assert diff_states(orcs[a].state(), parts.state(agent=a)) == [], a
    return whole, parts


def test_three_shards_on_one_device_equal_one_engine_and_the_oracle():
    whole, parts = _drive(lambda cfg, ld, A: ShardedEngine(cfg, ld, A, devices=[0, 0, 0], seed=1234))
    assert parts.count == [2, 2, 3] and parts.first == [0, 2, 4]
    # device noise: keyed by the global agent, so the sharded run still equals the single engine
    x = np.zeros((7, 64), np.float32)
    for e in (whole, parts):
        e.set_inputs(x)
        e.step(rewards=np.zeros(7, np.float32))
    assert np.array_equal(whole.get_predictions(), parts.get_predictions())
    assert np.array_equal(whole.get_column_actions(), parts.get_column_actions())
    parts.close()
    whole.close()


def test_sharded_frame_path_equals_one_engine():
    """The device-resident path (frames, feedback taps, engine noise) through the sharded calls."""
    cfg, ld = cf.config_c2()
    A, T = 6, 8
    whole = Engine(cfg, ld, n_agents=A, seed=77)
    parts = ShardedEngine(cfg, ld, A, devices=[0, 0], seed=77)
    frames = base_frames(T, A, 0)
    for e in (whole, parts):
        e.load_frames(frames)
        e.set_feedback(FB_IDX, FB_IDX, 0.5, 0.5, 0.05, True)
    parts.timer_start()
    for t in range(12):
        whole.step_frame(t)
        parts.step_frame(t)
    assert parts.timer_stop_ms() > 0.0
    assert np.array_equal(whole.get_predictions(), parts.get_predictions())
    assert np.array_equal(whole.get_column_actions(), parts.get_column_actions())
    assert parts.launch_count > 0 and parts.device_bytes > whole.device_bytes * 0.9


@pytest.mark.gpu2
def test_two_devices_from_one_process():
    """Two shards on two GPUs driven by one host thread: exact parity with one engine holding all agents."""
    from conftest import require_two_gpus
    require_two_gpus()
    whole, parts = _drive(lambda cfg, ld, A: ShardedEngine(cfg, ld, A, devices=[0, 1], seed=1234))
    parts.close()
    whole.close()


def test_bad_arguments():
    import bidinet_b200
    cfg, ld = cf.config_c2()
    with pytest.raises(bidinet_b200.BidiError):
        ShardedEngine(cfg, ld, 1, devices=[0, 0])   # fewer agents than shards
    with pytest.raises(bidinet_b200.BidiError):
        ShardedEngine(cfg, ld, 4, devices=[0, 99])  # no such device
