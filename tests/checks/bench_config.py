#!/usr/bin/env python
"""Single-agent timings of the BASELINE.json configurations other than the bench headline:
C1 (small), C4 (large 6-layer, radius 6), C5-size single layer -- engine ms/step on the GPU,
the compiled reference (oracle/_ref) ms/step on one host core, and an exact-parity check
of the first steps at full size.  usage: python tests/checks/bench_config.py [c1|c4|c5] ...
Prints one JSON object per configuration."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bidinet_b200 import Engine, config as cf  # noqa: E402
from oracle.pyoracle import OracleHierarchy, RefHierarchy, have_ref, c1_frame  # noqa: E402


def bars(t, n=128, period=16):
    """C4 input: 128x128 moving bars, period 16, values {0,1}."""
    x = np.arange(n)[None, :] + np.zeros((n, 1), dtype=int)
    return (((x + t) % period) < period // 2).astype(np.float32).ravel()


def sparse_frames(t, n, density=0.05, period=8):
    rng = np.random.RandomState(7 + t % period)
    return (rng.rand(n * n) < density).astype(np.float32)


def run(name, cfg, ld, frame, gpu_steps, cpu_steps, parity_steps):
    cls = RefHierarchy if have_ref() else OracleHierarchy
    ref = cls(cfg, ld, 1234)
    eng = Engine(cfg, ld, n_agents=1, seed=1234)
    for kv in filter(None, os.environ.get("BIDI_OPTS", "").split(",")):  # e.g. BIDI_OPTS=grid_kernels=2
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    # exact parity of the first steps at full size (the engine replays createRandom itself)
    ok = True
    for t in range(parity_steps):
        x = frame(t)
        ref.set_inputs(x); eng.set_inputs(x)
        noise = ref.peek_noise() if ref.num_noise() else None  # the exploration draws the reference is about to make
        ref.step(); eng.step(rewards=[0.0], noise=noise)
        ok = ok and np.array_equal(ref.get_predictions(), eng.get_predictions()[0])
    if parity_steps:
        keys = [("HID_STATE", 0), ("FF_W", 0), ("P_STATE", cfg.n_layers - 1)]
        if cfg.variant == cf.QPRSDR:
            keys += [("Q_FF_W", 0), ("ACT_EXPL", 0), ("QCONN_TRACE", 0)]
        for key in keys:
            ok = ok and np.array_equal(ref.get_array(*key), eng.get_array(*key))
    # CPU: one core
    frames = np.stack([frame(t) for t in range(16)])
    t0 = time.perf_counter()
    ref.run(frames, cpu_steps)
    cpu_ms = 1e3 * (time.perf_counter() - t0) / max(1, cpu_steps)
    # GPU: resident frames, CUDA events
    eng.load_frames(frames[:, None, :])
    for t in range(5):
        eng.step_frame(t)
    eng.sync()
    l0 = eng.launch_count
    eng.timer_start()
    for t in range(gpu_steps):
        eng.step_frame(t)
    gpu_ms = eng.timer_stop_ms() / gpu_steps
    launches = (eng.launch_count - l0) / gpu_steps
    # end to end through the host API
    t0 = time.perf_counter()
    for t in range(gpu_steps):
        eng.set_inputs(frames[t % 16]); eng.step(); eng.get_predictions()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / gpu_steps
    out = {"config": name, "variant": cf.VARIANT_NAMES[cfg.variant], "gpu_ms_per_step": gpu_ms, "e2e_ms_per_step": e2e_ms,
           "cpu_ms_per_step_1core": cpu_ms, "cpu_kind": "reference" if have_ref() else "port",
           "speedup_resident": cpu_ms / gpu_ms, "speedup_e2e": cpu_ms / e2e_ms, "launches_per_step": launches,
           "parity_exact_first_steps": bool(ok), "parity_steps": parity_steps, "device_MB": eng.device_bytes / 1e6,
           "options": os.environ.get("BIDI_OPTS", "")}
    print(json.dumps(out), flush=True)
    eng.close()


def main(which):
    if "c1" in which:
        for v in (cf.KWINNERS, cf.ITERATIVE, cf.NEO_HIERARCHY):
            cfg, ld = cf.config_c1(v)
            run("C1", cfg, ld, c1_frame, 200, 100, 20)
        cfg, ld = cf.config_c2()
        run("C2", cfg, ld, lambda t: np.random.RandomState(t).rand(64).astype(np.float32), 200, 100, 0)
    if "q" in which:  # sdr::QPRSDR over the C1 hierarchy; the reward follows the frame index
        cfg, ld = cf.config_q_c1()
        run("C1+Q", cfg, ld, c1_frame, 200, 100, 20)
    if "c4" in which:
        cfg, ld = cf.config_c4(cf.KWINNERS)
        run("C4", cfg, ld, bars, 100, 20, 5)
        cfg, ld = cf.config_c4(cf.ITERATIVE)
        run("C4", cfg, ld, bars, 20, 3, 2)
    if "c5" in which:
        n = 256  # the unmodified reference's unsigned-short indices cap a layer at 65536 units
        cfg, ld = cf.config_c5(n)
        run("C5-256 (largest layer the reference's 16-bit indices allow)", cfg, ld, lambda t: sparse_frames(t, n), 50, 5, 3)


if __name__ == "__main__":
    main(sys.argv[1:] or ["c1", "c4", "c5"])
