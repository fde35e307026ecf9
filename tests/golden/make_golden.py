"""Generates tests/golden/golden.npz from the UNMODIFIED reference
(oracle/_ref/libbidiref.so, built from /root/reference by oracle/Makefile).

Run in the build container:  python tests/golden/make_golden.py
For every case of tests/util.small_cases() it records, over STEPS steps from seed
1234: getPrediction of every input after every step, and a SHA-256 of every state
array at the end (-0.0 normalised to +0.0).  golden_random.npz holds the same for
the N_RANDOM randomly drawn geometries of tests/util.random_case (all five variants,
6 steps each).  The reference itself ships no golden vectors;
these are outputs of the reference run here."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.pyoracle import RefHierarchy  # noqa: E402
from util import N_RANDOM, run_random_case, small_cases  # noqa: E402

STEPS = 40
SEED = 1234


def array_digest(a):
    return hashlib.sha256((np.asarray(a, dtype="<f4") + np.float32(0.0)).tobytes()).hexdigest()


def run_case(cls, cfg, ld, frame, reward):
    h = cls(cfg, ld, SEED)
    preds = []
    for t in range(STEPS):
        h.set_inputs(frame(t))
        h.step(reward(t), t % 13 != 12)
        preds.append(h.get_predictions())
    st = h.state()
    keys = sorted("%s:%d" % k for k in st)
    digests = [array_digest(st[(k.split(":")[0], int(k.split(":")[1]))]) for k in keys]
    return np.stack(preds), keys, digests


def main():
    out = {}
    for name, cfg, ld, frame, reward in small_cases():
        preds, keys, digests = run_case(RefHierarchy, cfg, ld, frame, reward)
        out[name + "/pred"] = preds
        out[name + "/keys"] = np.array(keys)
        out[name + "/digests"] = np.array(digests)
        print(name, preds.shape, len(keys))
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    out = {}
    for seed in range(N_RANDOM):
        preds, keys, digests = run_random_case(RefHierarchy, seed)
        out["r%d/pred" % seed] = preds
        out["r%d/keys" % seed] = np.array(keys)
        out["r%d/digests" % seed] = np.array(digests)
    np.savez_compressed(os.path.join(HERE, "golden_random.npz"), **out)
    print("random geometries:", N_RANDOM)


if __name__ == "__main__":
    main()
