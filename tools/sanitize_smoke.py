#!/usr/bin/env python
"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck):
  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
C2 neo::Agent (column kernels, dense coder), C1 k-winners / iterative (tile kernels, fused coder,
per-phase kernels), a QPRSDR, a split layer (grid kernels + same-process halo copies); 2 steps each."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bidinet_b200 import Engine, StripGroup, config as cf  # noqa: E402


def run(cfg, ld, agents=2, opts=(), steps=2):
    eng = Engine(cfg, ld, n_agents=agents, seed=1234)
    for k, v in opts:
        eng.set_option(k, v)
    rng = np.random.RandomState(0)
    for t in range(steps):
        eng.set_inputs(rng.rand(agents, eng.n_in).astype(np.float32))
        eng.step(rewards=np.zeros(agents, np.float32))
    p = eng.get_predictions()
    assert np.isfinite(p).all()
    eng.close()


run(*cf.config_c2())
run(*cf.config_c1(cf.KWINNERS), agents=1)
run(*cf.config_c1(cf.KWINNERS), agents=1, opts=(("grid_kernels", 2),))
run(*cf.config_c1(cf.ITERATIVE), agents=1)
run(*cf.config_c1(cf.ITERATIVE), agents=1, opts=(("persist_coder", 0),))
run(*cf.config_c1(cf.NEO_HIERARCHY), agents=1, opts=(("force_phase_kernels", 1), ("tile_kernels", 0)))
run(*cf.config_q_small())
cfg, ld = cf.make_config(cf.KWINNERS, 24, 24, [dict(width=24, height=24, r_ff=3, r_rec=3, r_lat=3, r_pred=3, sparsity=0.15)])
grp = StripGroup(cfg, ld, devices=[0, 0, 0], seed=5)
x = (np.random.RandomState(1).rand(24 * 24) < 0.3).astype(np.float32)
for t in range(2):
    grp.set_inputs(x)
    grp.step()
assert np.isfinite(grp.get_predictions()).all()
grp.close()
print("sanitize_smoke: ok")
