"""tools/runner_env.c (the compiled environment of bench.py's end-to-end loop) against the same arithmetic
in numpy: RunnerMain.cpp:235-251 as restated by oracle/ref_harness.cpp ref_run_runner.  CPU only."""
import ctypes

import numpy as np

from bidinet_b200.build import build_runner_env


def test_runner_env_step_matches_numpy():
    lib = ctypes.CDLL(build_runner_env())
    fp = ctypes.POINTER(ctypes.c_float)
    lib.runner_env_step.argtypes = [ctypes.c_int] * 4 + [fp] * 4
    lib.runner_env_step.restype = None
    rng = np.random.RandomState(3)
    A, n_in, fb0, n_fb = 37, 65, 48, 16
    pred = (rng.rand(A, n_in).astype(np.float32) * 4.0 - 2.0)
    noise = rng.normal(0.0, 0.05, (A, n_fb)).astype(np.float32)
    x = rng.rand(A, n_in).astype(np.float32)
    x0 = x.copy()
    rewards = np.zeros(A, dtype=np.float32)
    lib.runner_env_step(A, n_in, fb0, n_fb, *(a.ctypes.data_as(fp) for a in (pred, noise, x, rewards)))
    act = pred[:, fb0:fb0 + n_fb] * np.float32(0.5) + np.float32(0.5)
    assert np.array_equal(x[:, fb0:fb0 + n_fb], np.clip(act + noise, 0.0, 1.0))
    assert np.array_equal(x[:, :fb0], x0[:, :fb0]) and np.array_equal(x[:, fb0 + n_fb:], x0[:, fb0 + n_fb:])
    want = np.zeros(A, dtype=np.float32)
    clipped = np.clip(act, 0.0, 1.0)
    for k in range(n_fb):  # the reference's left-to-right sum
        want += clipped[:, k]
    assert np.array_equal(rewards, want / np.float32(n_fb) - np.float32(0.5))
