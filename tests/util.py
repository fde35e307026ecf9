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
    # deep::CSRL: sdr::IRSDR coders + one deep::SDRRL unit per node; a small one, and a ragged three-layer one (a layer
    # without recurrent SDRRL inputs, windows wider than the grid, a 1x1 top layer, no action inputs in the first)
    cfg, ld = cf.config_csrl_small()
    cases.append(("csrl_small", cfg, ld, lambda t: frames35[t % 64][:30], lambda t: float(np.sin(0.3 * t))))
    layers = [dict(width=5, height=4, r_ff=2, r_rec=1, r_lat=2, r_pred=1, r_fb=1, csrl_n_recurrent=0, csrl_derive_iterations=3,
                   iter_settle=3, iter_measure=1, col_cells=3),
              dict(width=3, height=2, r_ff=1, r_rec=2, r_lat=1, r_pred=3, r_fb=2, csrl_n_recurrent=1, csrl_derive_iterations=2,
                   iter_settle=2, iter_measure=2, col_cells=6),
              dict(width=1, height=1, r_ff=4, r_rec=0, r_lat=0, r_pred=0, r_fb=3, csrl_n_recurrent=5, csrl_derive_iterations=0,
                   iter_settle=0, iter_measure=3, col_cells=2)]
    cfg, ld = cf.make_config(cf.CSRL, 7, 5, layers, r_infb=2, actions=[3, 10, 17, 30], init=(-0.4, 0.4, 0.01, 0.05, 0.02),
                             csrl_n_recurrent=1, csrl_derive_iterations=3, csrl_iter_settle=1, csrl_iter_measure=2, col_cells=7)
    cases.append(("ragged_csrl", cfg, ld, lambda t: frames35[t % 64], lambda t: float(np.cos(0.7 * t))))
    return cases


N_RANDOM = 40
N_CSRL_RANDOM = 10


def csrl_random_case(seed):
    """A random small deep::CSRL: grids 1..10 wide, 1..3 layers, radii 0..3, 1..8 cells per SDRRL unit."""
    rng = np.random.RandomState(7000 + seed)
    n_layers = int(rng.randint(1, 4))
    iw, ih = int(rng.randint(1, 9)), int(rng.randint(1, 9))
    layers = []
    for l in range(n_layers):
        layers.append(dict(width=int(rng.randint(1, 11)), height=int(rng.randint(1, 11)), r_ff=int(rng.randint(0, 4)),
                           r_rec=int(rng.randint(0, 3)), r_lat=int(rng.randint(0, 3)), r_pred=int(rng.randint(0, 3)),
                           r_fb=int(rng.randint(0, 3)), csrl_n_recurrent=int(rng.randint(0, 4)),
                           csrl_derive_iterations=int(rng.randint(0, 5)), iter_settle=int(rng.randint(0, 4)),
                           iter_measure=int(rng.randint(1, 4)), col_cells=int(rng.randint(1, 9))))
    n_in = iw * ih
    k = int(rng.randint(0, min(6, n_in) + 1))
    actions = sorted(int(c) for c in rng.permutation(n_in)[:k])
    cfg, ld = cf.make_config(cf.CSRL, iw, ih, layers, r_infb=int(rng.randint(0, 4)), actions=actions,
                             init=(-0.4, 0.4, 0.01, 0.05, float(rng.choice([0.0, 0.05]))), csrl_n_recurrent=int(rng.randint(0, 4)),
                             csrl_derive_iterations=int(rng.randint(0, 5)), csrl_iter_settle=int(rng.randint(0, 4)),
                             csrl_iter_measure=int(rng.randint(1, 4)), col_cells=int(rng.randint(1, 9)))
    frames = (rng.rand(8, n_in) < 0.4).astype(np.float32) * rng.rand(8, n_in).astype(np.float32)
    return cfg, ld, frames


def random_case(seed):
    """A random small hierarchy: grids 1..20 wide, ratios above and below 1, radii from 0 to wider than the grid."""
    rng = np.random.RandomState(1000 + seed)
    variant = [cf.KWINNERS, cf.ITERATIVE, cf.NEO_HIERARCHY, cf.NEO_AGENT, cf.QPRSDR][seed % 5]
    n_layers = int(rng.randint(1, 4))
    iw, ih = int(rng.randint(1, 21)), int(rng.randint(1, 21))
    layers = []
    for l in range(n_layers):
        d = dict(width=int(rng.randint(1, 21)), height=int(rng.randint(1, 21)), r_ff=int(rng.randint(0, 6)),
                 r_rec=int(rng.randint(-1, 4)) if variant == cf.KWINNERS else int(rng.randint(0, 4)),
                 r_lat=int(rng.randint(0, 5)), r_pred=int(rng.randint(0, 5)), r_fb=int(rng.randint(0, 5)))
        if variant == cf.KWINNERS:
            d["sparsity"] = float(rng.choice([0.05, 0.2, 0.5]))
        else:
            d["iter_settle"] = int(rng.randint(1, 7))
            if variant == cf.ITERATIVE or variant == cf.QPRSDR:
                d["iter_measure"] = int(rng.randint(1, 4))
        if variant == cf.NEO_AGENT:
            d["col_cells"] = int(rng.choice([16, 16, 7, 3]))
            d["col_iter"] = int(rng.randint(1, 5))
        layers.append(d)
    kw = dict(init=(-0.4, 0.4, 0.01, 0.05, float(rng.choice([0.0, 0.05]))))
    if variant == cf.QPRSDR:
        n_in = iw * ih
        if n_in < 2:
            iw, n_in = 2, 2 * ih
        k = int(rng.randint(1, min(8, n_in // 2) + 1))
        cells = rng.permutation(n_in)[:2 * k]
        kw.update(actions=[int(c) for c in cells[:k]], anti_actions=[int(c) for c in cells[k:]], q_derive_iterations=int(rng.randint(1, 6)))
    if variant == cf.NEO_AGENT:
        kw.update(col_cells=int(rng.choice([16, 5])), col_iter=int(rng.randint(1, 4)))
    cfg, ld = cf.make_config(variant, iw, ih, layers, r_infb=0 if variant == cf.KWINNERS else int(rng.randint(0, 5)), **kw)
    frames = (rng.rand(8, iw * ih) < 0.4).astype(np.float32) * rng.rand(8, iw * ih).astype(np.float32)
    return cfg, ld, frames


RANDOM_STEPS = 6


def random_schedule(t):
    """(reward, learn) of step t in the random-geometry runs: one no-learn step in the middle."""
    return 0.3 * t - 0.5, t != 3


def state_digests(st):
    """sorted "ARRAY:layer" keys and the SHA-256 of every state array (-0.0 normalised to +0.0)."""
    import hashlib
    keys = sorted("%s:%d" % k for k in st)
    out = []
    for k in keys:
        name, layer = k.split(":")
        a = np.asarray(st[(name, int(layer))], dtype="<f4") + np.float32(0.0)
        out.append(hashlib.sha256(a.tobytes()).hexdigest())
    return keys, out


def run_random_case(cls, seed):
    """Steps hierarchy class `cls` (reference, oracle) through random case `seed` on its own noise draws:
    predictions after every step, digests of every state array at the end."""
    cfg, ld, frames = random_case(seed)
    h = cls(cfg, ld, 4321 + seed)
    preds = []
    for t in range(RANDOM_STEPS):
        h.set_inputs(frames[t % len(frames)])
        h.step(*random_schedule(t))
        preds.append(h.get_predictions())
    keys, digests = state_digests(h.state())
    return np.stack(preds), keys, digests
