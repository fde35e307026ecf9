"""Synthetic inputs and algorithmic-byte counts of the BASELINE.json configurations.

Shared by bench.py (both arms) and the tests; pure numpy, no device and no checker code.
"""
import numpy as np

from . import config as cf


def c1_frames():
    """C1: 10-frame repeating binary pattern on 16x16, x(x,y,t) = ((x+p)%8<2 && (y+2p)%8<3), p = t%10
    (the shape of Experiment_Prediction.cpp:93-103 widened to the grid)."""
    out = []
    for p in range(10):
        x = np.arange(16)[None, :]
        y = np.arange(16)[:, None]
        out.append((((x + p) % 8 < 2) & ((y + 2 * p) % 8 < 3)).astype(np.float32).ravel())
    return np.stack(out)


def c4_frames(n=128, period=16):
    """C4: 128x128 moving bars, period 16, values {0,1}."""
    x = np.arange(n)[None, :] + np.zeros((n, 1), dtype=int)
    return np.stack([(((x + t) % period) < period // 2).astype(np.float32).ravel() for t in range(period)])


def c5_frames(size, n_frames=8, density=0.05, seed=7):
    """C5: sparse random binary frames (density 5 %, seed 7), period 8."""
    rng = np.random.RandomState(seed)
    return (rng.rand(n_frames, size * size) < density).astype(np.float32)


WORKLOADS = {
    # name: (config builder taking the variant, frames, label)
    "c1": (cf.config_c1, c1_frames, "C1: input 16x16, layers 16^2 -> 8^2 -> 4^2, class defaults, repeating 10-frame pattern"),
    "c4": (cf.config_c4, c4_frames, "C4: input 128x128, layers 96^2 -> 64^2 -> 48^2 -> 32^2 -> 24^2 -> 16^2, every radius 6, moving bars"),
}
VARIANTS = {"kwinners": cf.KWINNERS, "iterative": cf.ITERATIVE}
CLASS_NAMES = {cf.KWINNERS: "sdr::PredictiveRSDR", cf.ITERATIVE: "sdr::IPredictiveRSDR"}


def hierarchy_bytes_per_step(eng, variant, rho=None):
    """Algorithmic bytes of one simStep of one agent, SURVEY 8d:
    KWINNERS   4*(N_ff+N_rec+N_fb+N_pred) + 4*rho*(N_ff+N_rec)        rho: winners / units (~0.012)
    ITERATIVE  4*(1+rho)*(N_ff+N_rec) + 8*(N_lat+N_fb+N_pred+N_infb)  rho -> 1: the coder update is effectively dense
    (every weight read once per step, 4 B more where the learning rule writes it back; perfect on-chip reuse over
    the coder's iterations assumed)."""
    n = {k: 0 for k in ("FF_W", "REC_W", "LAT_W", "FB_W", "PRED_W")}
    for l in range(eng.n_layers + 1):
        for k in n:
            n[k] += eng.array_len(k, l)
    if variant == cf.KWINNERS:
        rho = 0.012 if rho is None else rho
        return int(4 * (n["FF_W"] + n["REC_W"] + n["FB_W"] + n["PRED_W"]) + 4 * rho * (n["FF_W"] + n["REC_W"]))
    rho = 1.0 if rho is None else rho
    return int(4 * (1 + rho) * (n["FF_W"] + n["REC_W"]) + 8 * (n["LAT_W"] + n["FB_W"] + n["PRED_W"]))
