#!/usr/bin/env python
"""Per-kernel device time of the C3 step at steady state, CUDA events inside the step graph
(bidi_profile_select), no profiler attached.  usage: python tools/kernel_times.py [agents] [burn_in]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bidinet_b200 import Engine, config as cf  # noqa: E402
from bidinet_b200.sharding import FB_IDX, base_frames  # noqa: E402

A = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 400
cfg, ld = cf.config_c2()
eng = Engine(cfg, ld, n_agents=A, seed=1234)
for kv in filter(None, os.environ.get("BIDI_OPTS", "").split(",")):
    k, v = kv.split("=")
    eng.set_option(k, int(v))
eng.load_frames(base_frames(64, A, 0, 64))
eng.set_feedback(FB_IDX, FB_IDX, 0.5, 0.5, 0.05, True)
t = 0
for _ in range(burn):
    eng.step_frame(t); t += 1
eng.sync()
eng.timer_start()
for _ in range(20):
    eng.step_frame(t); t += 1
print("step: %.3f ms after %d burn-in steps, %d agents" % (eng.timer_stop_ms() / 20, burn, A))
tot = 0.0
for name in ("coder", "columns", "predict", "predict_input", "neo_learn", "step_end"):
    eng.profile_select(name)
    for _ in range(3):
        eng.step_frame(t); t += 1
    eng.profile_collect()
    ms0, n0 = eng.profile_collect()
    for _ in range(10):
        eng.step_frame(t); t += 1
        eng.profile_collect()
    ms, n = eng.profile_collect()
    per_step = (ms - ms0) / 10
    tot += per_step
    print("%-14s %7.3f ms per step  (%d launches per step)" % (name, per_step, (n - n0) // 10))
eng.profile_select("")
print("sum: %.3f ms" % tot)
