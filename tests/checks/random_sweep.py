#!/usr/bin/env python
"""One-off parity sweep on the GPU over many more random geometries than the test suite holds
(tests/util.random_case: all five variants, 1-3 layers, grids 1-20 wide, radii 0 to wider than the grid,
3-16-cell columns): every state array after each of 6 steps against the oracle, default engine path.
usage: python tests/checks/random_sweep.py first_seed last_seed max_seconds"""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np
from bidinet_b200 import Engine, config as cf
from oracle.pyoracle import OracleHierarchy
from util import random_case, random_schedule, diff_states
t0 = time.time(); bad = []; n = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    if time.time() - t0 > float(sys.argv[3]): break
    cfg, ld, frames = random_case(seed)
    orc = OracleHierarchy(cfg, ld, 4321 + seed)
    eng = Engine(cfg, ld, n_agents=1)
    eng.load_state(orc.state())
    ok = True
    for t in range(6):
        x = frames[t % len(frames)]
        orc.set_inputs(x); eng.set_inputs(x)
        noise = orc.peek_noise() if orc.num_noise() else None
        r, learn = random_schedule(t)
        orc.step(r, learn); eng.step(rewards=[r], noise=noise, learn=learn)
        d = diff_states(orc.state(), eng.state())
        if d: ok = False; bad.append((seed, t, d[:3])); break
    eng.close(); n += 1
print("cases", n, "failures", len(bad), bad[:5], "%.1fs" % (time.time() - t0))
