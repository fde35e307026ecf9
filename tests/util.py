"""Shared helpers of the parity tests."""
import numpy as np

from bidinet_b200 import config as cf
from oracle.pyoracle import c1_frame

EXPERIMENT_SEQ = np.array(  # Experiment_Prediction.cpp:93-103, 3 values on a 2x2 input
    [[0, 1, 0, 0], [0, 0, 0, 0], [1, 1, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0],
     [0, 0, 1, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 1, 0, 0], [0, 1, 1, 0]], dtype=np.float32)


def diff_states(sa, sb, rtol=0.0, atol=0.0):
    """List of (key, max abs diff) where two state dicts differ (exact by default)."""
    bad = []
    if sa.keys() != sb.keys():
        return [("keys", sorted(set(sa) ^ set(sb)))]
    for k in sa:
        a, b = sa[k], sb[k]
        if a.shape != b.shape:
            bad.append((k, "shape %s vs %s" % (a.shape, b.shape)))
        elif rtol == 0.0 and atol == 0.0:
            if not np.array_equal(a, b):
                bad.append((k, float(np.nanmax(np.abs(a - b)))))
        elif not np.allclose(a, b, rtol=rtol, atol=atol):
            bad.append((k, float(np.nanmax(np.abs(a - b)))))
    return bad


def small_cases():
    """(name, cfg, layers, frame_fn, reward_fn): configs the oracle steps in milliseconds."""
    rng = np.random.RandomState(0)
    frames64 = rng.rand(64, 64).astype(np.float32)
    frames35 = (rng.rand(64, 35) < 0.3).astype(np.float32)
    cases = []
    for v in (cf.KWINNERS, cf.ITERATIVE, cf.NEO_HIERARCHY):
        cfg, ld = cf.config_c1(v)
        cases.append(("c1_" + cf.VARIANT_NAMES[v], cfg, ld, c1_frame, lambda t: 0.0))
    cfg, ld = cf.config_experiment_prediction()
    cases.append(("experiment_prediction", cfg, ld, lambda t: EXPERIMENT_SEQ[t % 10], lambda t: 0.0))
    cfg, ld = cf.config_c2()
    cases.append(("c2_neo_agent", cfg, ld, lambda t: frames64[t % 64], lambda t: float(np.sin(0.3 * t))))
    cfg, ld = cf.config_q_small()   # sdr::QPRSDR: hierarchy + Q network + action derivation
    cases.append(("q_small_qprsdr", cfg, ld, lambda t: frames64[t % 64], lambda t: float(np.sin(0.3 * t))))
    # ragged: non-square grids, non-integer ratios, radius -1 recurrent, windows wider than the grid
    for v in (cf.KWINNERS, cf.ITERATIVE, cf.NEO_HIERARCHY, cf.NEO_AGENT):
        layers = [dict(width=6, height=4, r_ff=2, r_rec=1, r_lat=2, r_pred=1, r_fb=1),
                  dict(width=3, height=5, r_ff=1, r_rec=-1 if v == cf.KWINNERS else 2, r_lat=1, r_pred=3, r_fb=2),
                  dict(width=2, height=2, r_ff=4, r_rec=1, r_lat=1, r_pred=1, r_fb=1)]
        if v == cf.KWINNERS:
            for l in layers:
                l["sparsity"] = 0.3
        if v == cf.NEO_AGENT:
            layers[1]["col_cells"] = 5
            layers[1]["col_iter"] = 3
        cfg, ld = cf.make_config(v, 7, 5, layers, r_infb=0 if v == cf.KWINNERS else 2)
        cases.append(("ragged_" + cf.VARIANT_NAMES[v], cfg, ld, lambda t: frames35[t % 64],
                      lambda t: float(np.cos(0.7 * t))))
    # degenerate sizes: a 1x1 input, a 1x1 top layer, radius 0 windows (each unit sees one cell)
    for v in (cf.KWINNERS, cf.ITERATIVE, cf.NEO_AGENT):
        layers = [dict(width=3, height=2, r_ff=0, r_rec=0, r_lat=1, r_pred=0, r_fb=0),
                  dict(width=1, height=1, r_ff=1, r_rec=-1 if v == cf.KWINNERS else 0, r_lat=0, r_pred=0, r_fb=0)]
        if v == cf.KWINNERS:
            for l in layers:
                l["sparsity"] = 0.6
        cfg, ld = cf.make_config(v, 1, 1, layers, r_infb=0 if v == cf.KWINNERS else 1, init=(-0.5, 0.5, 0.01, 0.05, 0.0))
        cases.append(("tiny_" + cf.VARIANT_NAMES[v], cfg, ld, lambda t: np.array([float(t % 3 == 0)], np.float32),
                      lambda t: float(t % 2)))
    # sdr::QPRSDR on ragged grids: three Q layers, actions scattered over the input, a window wider than the grid
    layers = [dict(width=6, height=4, r_ff=2, r_rec=1, r_lat=2, r_pred=1, r_fb=1),
              dict(width=3, height=5, r_ff=1, r_rec=2, r_lat=1, r_pred=3, r_fb=2),
              dict(width=2, height=2, r_ff=4, r_rec=1, r_lat=1, r_pred=1, r_fb=1)]
    cfg, ld = cf.make_config(cf.QPRSDR, 7, 5, layers, r_infb=2, actions=[3, 10, 17, 30], anti_actions=[4, 11, 18, 31],
                             init=(-0.3, 0.3, 0.01, 0.05, 0.0), q_derive_iterations=9)
    cases.append(("ragged_qprsdr", cfg, ld, lambda t: frames35[t % 64], lambda t: float(np.cos(0.7 * t))))
    return cases
