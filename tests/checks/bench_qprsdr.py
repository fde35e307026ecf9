#!/usr/bin/env python
"""sdr::QPRSDR batched: A agents of the small QPRSDR configuration (tests/util q_small) on one GPU against the
compiled reference on the host cores.  usage: python tests/checks/bench_qprsdr.py [agents]"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bidinet_b200 import Engine, config as cf  # noqa: E402
from oracle.pyoracle import RefHierarchy, have_ref, OracleHierarchy  # noqa: E402

A = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cfg, ld = cf.config_q_small()
eng = Engine(cfg, ld, n_agents=A, seed=1234)
rng = np.random.RandomState(0)
frames = rng.rand(16, A, 64).astype(np.float32)
eng.load_frames(frames)
for t in range(20):
    eng.step_frame(t)
eng.sync()
eng.timer_start()
K = 50
for t in range(K):
    eng.step_frame(t)
ms = eng.timer_stop_ms() / K
cls = RefHierarchy if have_ref() else OracleHierarchy
cores = len(os.sched_getaffinity(0))
refs = [cls(cfg, ld, 1234 + k) for k in range(cores)]
steps = 300


def work(k):
    refs[k].run(frames[:, k, :], steps)


th = [threading.Thread(target=work, args=(k,)) for k in range(cores)]
t0 = time.perf_counter()
for t in th:
    t.start()
for t in th:
    t.join()
dt = time.perf_counter() - t0
print(json.dumps({"config": "QPRSDR 8x8 -> 8x8 -> 4x4, 8 actions, 32 derive iterations", "agents": A, "gpu_ms_per_step": ms,
                  "gpu_agent_steps_per_s": A / (ms * 1e-3), "cpu_agent_steps_per_s": cores * steps / dt, "cpu_threads": cores,
                  "cpu_kind": "reference" if have_ref() else "port"}))
