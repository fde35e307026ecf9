"""ctypes mirror of include/bidinet_b200.h's POD structs plus the reference's defaults.

The default values restate the four reference ``LayerDesc`` constructors
(sdr/PredictiveRSDR.h:32-41, sdr/IPredictiveRSDR.h:35-45,
neo/PredictiveHierarchy.h:34-44, neo/Agent.h:49-64) and the public hyper-parameter
members of the hierarchy objects (sdr/IPredictiveRSDR.h:101-105, neo/Agent.h:137-144).
``_columnGamma`` / ``_columnGammaLambda`` are never initialised by the reference
(neo/Agent.h:25-26,130-131); this build defaults them to 0.99 / 0.95 explicitly.
"""
import ctypes as C

KWINNERS, ITERATIVE, NEO_HIERARCHY, NEO_AGENT, QPRSDR, CSRL = 0, 1, 2, 3, 4, 5
MAX_ACTIONS = 64
VARIANT_NAMES = {KWINNERS: "kwinners", ITERATIVE: "iterative",
                 NEO_HIERARCHY: "neo_hierarchy", NEO_AGENT: "neo_agent", QPRSDR: "qprsdr", CSRL: "csrl"}
COLUMN_ACTIONS = 3

ARRAY_NAMES = [
    "VIS_INPUT", "VIS_RECON",
    "HID_THRESHOLD", "HID_STATE", "HID_STATE_PREV", "HID_ACTIVATION", "HID_RECON",
    "HID_SPIKE", "HID_SPIKE_PREV",
    "FF_W", "REC_W", "LAT_W", "FF_TRACE", "REC_TRACE",
    "P_STATE", "P_STATE_PREV", "P_ACTIVATION", "P_ACTIVATION_PREV", "P_BASELINE",
    "FB_W", "PRED_W",
    "COL_INPUT", "COL_RECON_ERR", "COL_FF_W", "COL_LAT_W",
    "COL_THRESHOLD", "COL_ACTIVATION", "COL_SPIKE", "COL_SPIKE_PREV", "COL_STATE",
    "COL_Q_W", "COL_Q_TRACE", "COL_ACT_STATE", "COL_ACT_EXPL", "COL_ACT_W", "COL_ACT_TRACE",
    "COL_PREV_VALUE", "PREDICTION",
    "Q_STATE", "Q_ERROR", "Q_BIAS_W", "Q_BIAS_TRACE", "Q_FF_W", "Q_FF_TRACE",
    "QCONN_W", "QCONN_TRACE", "ACT_PREDICTED", "ACT_DERIVE", "ACT_EXPL", "ACT_ERROR", "Q_PREV_VALUE",
    "FB_TRACE", "PRED_TRACE", "P_LOCAL_REWARD", "COL_ACT_ERROR", "COL_CELL_ACT_STATE", "COL_CELL_ACT_ERROR",
    "COL_AVG_SURPRISE", "REWARD_OFFSETS",
]
ARRAY = {n: i for i, n in enumerate(ARRAY_NAMES)}
# arrays that hold a per-unit connection list (bidi_conn_counts is defined for them)
LIST_ARRAYS = ("FF_W", "REC_W", "LAT_W", "FB_W", "PRED_W", "COL_INPUT", "Q_FF_W")


class LayerDesc(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int),
        ("r_ff", C.c_int), ("r_rec", C.c_int), ("r_lat", C.c_int), ("r_pred", C.c_int), ("r_fb", C.c_int),
        ("learn_ff", C.c_float), ("learn_rec", C.c_float), ("learn_lat", C.c_float),
        ("learn_threshold", C.c_float),
        ("learn_fb", C.c_float), ("learn_pred", C.c_float),
        ("sparsity", C.c_float),
        ("avg_surprise_decay", C.c_float), ("attention_factor", C.c_float),
        ("iter_settle", C.c_int), ("iter_measure", C.c_int),
        ("leak", C.c_float), ("lambda_", C.c_float),
        ("weight_decay", C.c_float), ("max_weight_delta", C.c_float),
        ("baseline_decay", C.c_float), ("sensitivity", C.c_float),
        ("col_cells", C.c_int), ("col_iter", C.c_int),
        ("col_sparsity", C.c_float), ("col_leak", C.c_float),
        ("col_gamma", C.c_float), ("col_gamma_lambda", C.c_float),
        ("col_alpha_ff", C.c_float), ("col_alpha_lat", C.c_float), ("col_alpha_threshold", C.c_float),
        ("col_alpha_q", C.c_float), ("col_alpha_action", C.c_float),
        ("col_expl_stddev", C.c_float), ("col_expl_break", C.c_float),
        ("csrl_n_recurrent", C.c_int), ("csrl_derive_iterations", C.c_int),
        ("csrl_derive_alpha", C.c_float), ("csrl_surprise_factor", C.c_float),
    ]


class Config(C.Structure):
    _fields_ = [
        ("variant", C.c_int),
        ("in_width", C.c_int), ("in_height", C.c_int),
        ("r_infb", C.c_int), ("n_layers", C.c_int),
        ("learn_infb", C.c_float),
        ("col_cells", C.c_int), ("col_iter", C.c_int),
        ("col_sparsity", C.c_float), ("col_leak", C.c_float),
        ("col_gamma", C.c_float), ("col_gamma_lambda", C.c_float),
        ("col_alpha_ff", C.c_float), ("col_alpha_lat", C.c_float), ("col_alpha_threshold", C.c_float),
        ("col_alpha_q", C.c_float), ("col_alpha_action", C.c_float),
        ("col_expl_stddev", C.c_float), ("col_expl_break", C.c_float),
        ("init_min_w", C.c_float), ("init_max_w", C.c_float),
        ("init_min_inh", C.c_float), ("init_max_inh", C.c_float), ("init_threshold", C.c_float),
        ("q_alpha", C.c_float), ("q_action_alpha", C.c_float), ("q_derive_alpha", C.c_float),
        ("q_expl_break", C.c_float), ("q_expl_stddev", C.c_float), ("q_gamma", C.c_float), ("q_gamma_lambda", C.c_float),
        ("q_derive_iterations", C.c_int), ("q_n_actions", C.c_int),
        ("q_action_idx", C.c_int * MAX_ACTIONS), ("q_anti_idx", C.c_int * MAX_ACTIONS),
        ("csrl_n_recurrent", C.c_int), ("csrl_derive_iterations", C.c_int),
        ("csrl_iter_settle", C.c_int), ("csrl_iter_measure", C.c_int),
        ("csrl_derive_alpha", C.c_float), ("csrl_surprise_factor", C.c_float),
        ("csrl_avg_surprise_decay", C.c_float), ("csrl_leak", C.c_float),
    ]


_COLUMN_DEFAULTS = dict(
    col_cells=16, col_iter=7, col_sparsity=0.125, col_leak=0.1,
    col_gamma=0.99, col_gamma_lambda=0.95,
    col_alpha_ff=0.01, col_alpha_lat=0.05, col_alpha_threshold=0.01,
    col_alpha_q=0.01, col_alpha_action=0.1,
    col_expl_stddev=0.05, col_expl_break=0.01)

_LAYER_DEFAULTS = {
    KWINNERS: dict(
        width=16, height=16, r_ff=8, r_rec=4, r_lat=3, r_pred=4, r_fb=8,
        learn_ff=0.02, learn_rec=0.02, learn_lat=0.2, learn_threshold=0.12,
        learn_fb=0.05, learn_pred=0.05,
        avg_surprise_decay=0.01, attention_factor=4.0, sparsity=0.01),
    ITERATIVE: dict(
        width=16, height=16, r_ff=3, r_rec=3, r_lat=3, r_pred=3, r_fb=3,
        learn_ff=0.01, learn_rec=0.01, learn_lat=0.2, learn_fb=0.05, learn_pred=0.05,
        iter_settle=30, iter_measure=5, leak=0.3, lambda_=0.95,
        weight_decay=0.0, max_weight_delta=0.5, sparsity=0.2, learn_threshold=0.02,
        baseline_decay=0.01, sensitivity=8.0),
    NEO_HIERARCHY: dict(
        width=16, height=16, r_ff=4, r_rec=4, r_lat=4, r_pred=4, r_fb=4,
        learn_ff=0.01, learn_rec=0.01, learn_lat=0.05, learn_fb=0.1, learn_pred=0.03,
        iter_settle=30, iter_measure=0, leak=0.1, lambda_=0.95,
        weight_decay=0.0, max_weight_delta=0.5, sparsity=0.02, learn_threshold=0.01,
        baseline_decay=0.01, sensitivity=6.0),
}
_LAYER_DEFAULTS[NEO_AGENT] = dict(_LAYER_DEFAULTS[NEO_HIERARCHY], **_COLUMN_DEFAULTS)
# deep/CSRL.h:77-101 (LayerDesc) and :213-242 (the CSRL object's members for the input nodes)
_LAYER_DEFAULTS[CSRL] = dict(
    width=16, height=16, r_ff=5, r_rec=4, r_lat=4, r_pred=4, r_fb=5,
    learn_ff=0.01, learn_rec=0.01, learn_lat=0.2, learn_fb=0.05, learn_pred=0.05,
    iter_settle=17, iter_measure=4, leak=0.1, lambda_=0.95, weight_decay=0.0001, max_weight_delta=0.5,
    sparsity=0.08, learn_threshold=0.01, avg_surprise_decay=0.01,
    col_cells=8, col_sparsity=0.2, col_gamma=0.99, col_gamma_lambda=0.95,
    col_alpha_ff=0.05, col_alpha_lat=0.2, col_alpha_threshold=0.01, col_alpha_q=0.02, col_alpha_action=0.2,
    col_expl_stddev=0.1, col_expl_break=0.01,
    csrl_n_recurrent=4, csrl_derive_iterations=30, csrl_derive_alpha=0.05, csrl_surprise_factor=2.0)
_CSRL_CONFIG_DEFAULTS = dict(
    learn_infb=0.05, col_cells=8, col_sparsity=0.2, col_gamma=0.99, col_gamma_lambda=0.95,
    col_alpha_ff=0.05, col_alpha_lat=0.1, col_alpha_threshold=0.005, col_alpha_q=0.02, col_alpha_action=0.2,
    col_expl_stddev=0.1, col_expl_break=0.01,
    csrl_n_recurrent=4, csrl_derive_iterations=30, csrl_iter_settle=17, csrl_iter_measure=4,
    csrl_derive_alpha=0.05, csrl_surprise_factor=2.0, csrl_avg_surprise_decay=0.01, csrl_leak=0.1)
_LAYER_DEFAULTS[QPRSDR] = _LAYER_DEFAULTS[ITERATIVE]  # sdr::QPRSDR owns an sdr::IPredictiveRSDR

_LEARN_INFB = {KWINNERS: 0.0, ITERATIVE: 0.05, NEO_HIERARCHY: 0.1, NEO_AGENT: 0.1, QPRSDR: 0.05, CSRL: 0.05}
# sdr/QPRSDR.h:96-106
_Q_DEFAULTS = dict(q_alpha=0.01, q_action_alpha=0.1, q_derive_iterations=32, q_derive_alpha=0.05,
                   q_expl_break=0.01, q_expl_stddev=0.05, q_gamma=0.99, q_gamma_lambda=0.98)


def layer_desc(variant, **overrides):
    """A LayerDesc holding the reference defaults of ``variant``, then ``overrides``."""
    d = LayerDesc()
    vals = dict(_LAYER_DEFAULTS[variant])
    for k, v in overrides.items():
        vals["lambda_" if k == "lambda" else k] = v
    for k, v in vals.items():
        if not hasattr(d, k):
            raise AttributeError("LayerDesc has no field %r" % k)
        setattr(d, k, v)
    return d


def make_config(variant, in_width, in_height, layers, r_infb=0,
                init=(-0.01, 0.01, 0.01, 0.05, 0.1), actions=(), anti_actions=(), **overrides):
    """(Config, LayerDesc array) for ``variant``.

    ``layers`` is a list of dicts of LayerDesc overrides (at least width/height) or
    LayerDesc objects.  ``init`` = (minW, maxW, minInh, maxInh, threshold), the
    createRandom arguments (RunnerMain.cpp:154, Experiment_Prediction.cpp:126).
    """
    descs = (LayerDesc * len(layers))()
    for i, l in enumerate(layers):
        descs[i] = l if isinstance(l, LayerDesc) else layer_desc(variant, **l)
    cfg = Config()
    cfg.variant = variant
    cfg.in_width, cfg.in_height = in_width, in_height
    cfg.r_infb = r_infb
    cfg.n_layers = len(layers)
    cfg.learn_infb = _LEARN_INFB[variant]
    for k, v in _COLUMN_DEFAULTS.items():
        setattr(cfg, k, v)
    for k, v in _Q_DEFAULTS.items():
        setattr(cfg, k, v)
    if variant == CSRL:  # actions = the input cells of type _action; no anti-action cells
        for k, v in _CSRL_CONFIG_DEFAULTS.items():
            setattr(cfg, k, v)
        anti_actions = list(actions)
    assert len(actions) == len(anti_actions) <= MAX_ACTIONS
    cfg.q_n_actions = len(actions)
    for k, (a, b) in enumerate(zip(actions, anti_actions)):
        cfg.q_action_idx[k], cfg.q_anti_idx[k] = a, b
    (cfg.init_min_w, cfg.init_max_w, cfg.init_min_inh, cfg.init_max_inh,
     cfg.init_threshold) = init
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError("Config has no field %r" % k)
        setattr(cfg, k, v)
    return cfg, descs


# ---- the BASELINE.json configurations (SURVEY section 8 shorthand) -------------
def config_c1(variant=KWINNERS):
    """C1: input 16x16, layers 16^2 -> 8^2 -> 4^2, class defaults."""
    layers = [dict(width=16, height=16), dict(width=8, height=8), dict(width=4, height=4)]
    return make_config(variant, 16, 16, layers, r_infb=0 if variant == KWINNERS else 16)


def config_experiment_prediction():
    """Experiment_Prediction.cpp:113-126: input 2x2, 16^2 -> 12^2 -> 8^2, inFB 16, threshold 0."""
    layers = [dict(width=16, height=16), dict(width=12, height=12), dict(width=8, height=8)]
    return make_config(ITERATIVE, 2, 2, layers, r_infb=16, init=(-0.01, 0.01, 0.01, 0.05, 0.0))


def config_q_small():
    """A small sdr::QPRSDR: input 8x8 (cells 40..47 actions, 48..55 their complements), layers 8^2 -> 4^2."""
    layers = [dict(width=8, height=8), dict(width=4, height=4)]
    return make_config(QPRSDR, 8, 8, layers, r_infb=4, actions=list(range(40, 48)), anti_actions=list(range(48, 56)),
                       init=(-0.3, 0.3, 0.01, 0.05, 0.0))  # lively from the first steps: non-zero Q values, errors and traces


def config_q_c1():
    """sdr::QPRSDR over the C1 hierarchy (input 16x16, 16^2 -> 8^2 -> 4^2): 16 action cells + their complements."""
    layers = [dict(width=16, height=16), dict(width=8, height=8), dict(width=4, height=4)]
    return make_config(QPRSDR, 16, 16, layers, r_infb=16, actions=list(range(192, 208)), anti_actions=list(range(208, 224)),
                       init=(-0.3, 0.3, 0.01, 0.05, 0.0))


def config_csrl_small():
    """A small deep::CSRL: input 6x5 (cells 20..25 of type _action), layers 6x6 -> 3x3, short SDRRL loops."""
    layers = [dict(width=6, height=6, r_ff=2, r_rec=1, r_lat=2, r_pred=1, r_fb=1, csrl_n_recurrent=2, csrl_derive_iterations=5,
                   iter_settle=5, iter_measure=2, col_cells=4),
              dict(width=3, height=3, r_ff=2, r_rec=1, r_lat=1, r_pred=1, r_fb=1, csrl_n_recurrent=3, csrl_derive_iterations=4,
                   iter_settle=4, iter_measure=3, col_cells=5)]
    return make_config(CSRL, 6, 5, layers, r_infb=2, actions=[20, 21, 22, 23, 24, 25], init=(-0.3, 0.3, 0.01, 0.05, 0.05),
                       csrl_n_recurrent=2, csrl_derive_iterations=6, csrl_iter_settle=4, csrl_iter_measure=2, col_cells=3)


def config_c2():
    """C2/C3: neo::Agent of RunnerMain.cpp:141-154: input 8x8, layers 8^2 -> 4^2, inFB 8."""
    layers = [dict(width=8, height=8), dict(width=4, height=4)]
    return make_config(NEO_AGENT, 8, 8, layers, r_infb=8)


def config_c4(variant=KWINNERS):
    """C4: input 128x128, layers 96^2..16^2, every radius 6, inFB 6."""
    r = dict(r_ff=6, r_rec=6, r_lat=6, r_pred=6, r_fb=6)
    layers = [dict(width=w, height=w, **r) for w in (96, 64, 48, 32, 24, 16)]
    return make_config(variant, 128, 128, layers, r_infb=0 if variant == KWINNERS else 6)


def config_videotest():
    """VideoTest.cpp:52-65: sdr::IPredictiveRSDR, input 128x128, layers 32^2 -> 24^2 -> 16^2 (class defaults: every
    radius 3), input feedback radius 16 -> 9.8 M input-feedback connections; free-runs on its own predictions with
    learn = false (VideoTest.cpp:180-186)."""
    layers = [dict(width=32, height=32), dict(width=24, height=24), dict(width=16, height=16)]
    return make_config(ITERATIVE, 128, 128, layers, r_infb=16)


def config_c5(size=1024, variant=KWINNERS):
    """C5: one same-size layer (size x size over size x size), every radius 6."""
    r = dict(r_ff=6, r_rec=6, r_lat=6, r_pred=6, r_fb=6)
    return make_config(variant, size, size, [dict(width=size, height=size, **r)],
                       r_infb=0 if variant == KWINNERS else 6)
