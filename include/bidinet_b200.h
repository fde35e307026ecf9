/*
 * bidinet_b200.h -- C ABI of the B200-native BIDInet hierarchy-update engine.
 *
 * The reference (222464/BIDInet) has no FFI layer: its experiment mains call
 * plain C++ value classes (sdr::PredictiveRSDR, sdr::IPredictiveRSDR,
 * neo::PredictiveHierarchy, neo::Agent).  This header is the boundary a
 * drop-in facade for those classes binds to (see the headers under include/bidinet/ and
 * INTEGRATION.md).  Every entry point cites the reference interface it
 * replaces; paths are relative to BIDInet/source/ of the reference.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary
 *   - every call returns 0 on success, a negative BIDI_E* code otherwise;
 *     bidi_last_error() gives the message (no exceptions cross the ABI)
 *   - one engine = one device + one CUDA stream; calls on one engine must be
 *     serialised by the caller (the reference is single threaded as well)
 *   - "arrays" are the reference's per-layer state vectors, serialised in the
 *     REFERENCE order: unit index = x + y*width, connection lists clipped to
 *     the grid and concatenated unit after unit, inside a unit dx outer / dy
 *     inner (sdr/RSDR.cpp:48-62).  The engine converts to and from its own
 *     HBM layout.
 */
#ifndef BIDINET_B200_H
#define BIDINET_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define BIDI_MAX_LAYERS 16
#define BIDI_COLUMN_ACTIONS 3 /* neo/Agent.h:9-11 _attention,_learnPrediction,_signal */
#define BIDI_MAX_ACTIONS 64  /* action inputs of an sdr::QPRSDR */

/* Which reference hierarchy the engine reproduces. */
enum bidi_variant {
  BIDI_KWINNERS = 0,      /* sdr::PredictiveRSDR    sdr/PredictiveRSDR.cpp:100-194 */
  BIDI_ITERATIVE = 1,     /* sdr::IPredictiveRSDR   sdr/IPredictiveRSDR.cpp:138-240 */
  BIDI_NEO_HIERARCHY = 2, /* neo::PredictiveHierarchy neo/PredictiveHierarchy.cpp:137-245 */
  BIDI_NEO_AGENT = 3,     /* neo::Agent             neo/Agent.cpp:141-284 */
  BIDI_QPRSDR = 4,        /* sdr::QPRSDR: sdr::IPredictiveRSDR + the Q network and action derivation, sdr/QPRSDR.cpp:99-341 */
  BIDI_CSRL = 5           /* deep::CSRL: sdr::IRSDR coders + one deep::SDRRL unit per prediction node, deep/CSRL.cpp:189-449, deep/SDRRL.cpp:46-277 */
};

enum bidi_status {
  BIDI_OK = 0,
  BIDI_EINVAL = -1,  /* bad argument / unsupported configuration */
  BIDI_ECUDA = -2,   /* CUDA runtime error (message has the CUDA string) */
  BIDI_ENOMEM = -3,
  BIDI_ENODEV = -4   /* no usable CUDA device: there is NO CPU fallback */
};

/*
 * One layer of the hierarchy.  Union of the fields of the four reference
 * LayerDesc structs (sdr/PredictiveRSDR.h:14-42, sdr/IPredictiveRSDR.h:14-46,
 * neo/PredictiveHierarchy.h:14-45, neo/Agent.h:19-65); a variant ignores the
 * fields its reference class does not have.
 */
typedef struct bidi_layer_desc {
  int width, height;
  int r_ff, r_rec, r_lat, r_pred, r_fb; /* receptive, recurrent, lateral, predictive, feedback radii */
  float learn_ff, learn_rec, learn_lat;
  float learn_threshold;                /* _learnThreshold / _sdrLearnThreshold */
  float learn_fb, learn_pred;
  float sparsity;                       /* _sparsity / _sdrSparsity */
  /* k-winners (sdr/PredictiveRSDR.h:27-30) */
  float avg_surprise_decay, attention_factor;
  /* iterative coders */
  int iter_settle;                      /* _sdrIterSettle, or _sdrIter for the neo classes */
  int iter_measure;                     /* _sdrIterMeasure (ITERATIVE only) */
  float leak;                           /* _sdrLeak */
  float lambda;                         /* _sdrLambda (trace decay, neo classes) */
  float weight_decay, max_weight_delta; /* _sdrWeightDecay, _sdrMaxWeightDelta */
  float baseline_decay, sensitivity;    /* _sdrBaselineDecay, _sdrSensitivity */
  /* per-node columns (neo/Agent.h:22-31) */
  int col_cells, col_iter;
  float col_sparsity, col_leak, col_gamma, col_gamma_lambda;
  float col_alpha_ff, col_alpha_lat, col_alpha_threshold, col_alpha_q, col_alpha_action;
  float col_expl_stddev, col_expl_break;
  /* deep::CSRL only (deep/CSRL.h:29-75).  It also uses: learn_fb/learn_pred = _learnFeedBackRL/_learnPredictionRL,
   * avg_surprise_decay, sparsity, iter_settle/iter_measure/leak (coder AND SDRRL units), lambda, weight_decay,
   * max_weight_delta, learn_threshold, and the col_* fields for the SDRRL units (col_cells = _cellsPerColumn,
   * col_sparsity = _cellSparsity, col_alpha_ff/lat/threshold = the _gate*Alpha, col_alpha_q = _qAlpha,
   * col_alpha_action = _actionAlpha, col_gamma, col_gamma_lambda, col_expl_*) */
  int csrl_n_recurrent;       /* _numRecurrentInputs */
  int csrl_derive_iterations; /* _actionDeriveIterations */
  float csrl_derive_alpha;    /* _actionDeriveAlpha */
  float csrl_surprise_factor; /* _surpriseLearnFactor */
} bidi_layer_desc;

typedef struct bidi_config {
  int variant; /* enum bidi_variant */
  int in_width, in_height;
  int r_infb;        /* inputFeedBackRadius (not KWINNERS) */
  int n_layers;
  float learn_infb;  /* _learnInputFeedBack  sdr/IPredictiveRSDR.h:101 */
  /* columns of the input prediction nodes (public members, neo/Agent.h:125-135) */
  int col_cells, col_iter;
  float col_sparsity, col_leak, col_gamma, col_gamma_lambda;
  float col_alpha_ff, col_alpha_lat, col_alpha_threshold, col_alpha_q, col_alpha_action;
  float col_expl_stddev, col_expl_break;
  /* createRandom arguments (sdr/IPredictiveRSDR.h:107) */
  float init_min_w, init_max_w, init_min_inh, init_max_inh, init_threshold;
  /* sdr::QPRSDR only: public members (sdr/QPRSDR.h:84-107) and the action lists of its
   * createRandom (sdr/QPRSDR.h:109): q_n_actions input cells carry actions, as many carry
   * their complements ("anti-actions", value 1 - action) */
  float q_alpha, q_action_alpha, q_derive_alpha, q_expl_break, q_expl_stddev, q_gamma, q_gamma_lambda;
  int q_derive_iterations;
  int q_n_actions;
  int q_action_idx[BIDI_MAX_ACTIONS], q_anti_idx[BIDI_MAX_ACTIONS];
  /* deep::CSRL only: the public members for the input nodes' SDRRL units (deep/CSRL.h:183-243); the input cells of
   * type _action are q_action_idx[0..q_n_actions) (deep/CSRL.h:11-13, createRandom's inputTypes).  It also uses
   * learn_infb = _learnFeedBackRL and the col_* fields as the layers do. */
  int csrl_n_recurrent, csrl_derive_iterations, csrl_iter_settle, csrl_iter_measure;
  float csrl_derive_alpha, csrl_surprise_factor, csrl_avg_surprise_decay, csrl_leak;
} bidi_config;

/*
 * State arrays.  `layer` is 0..n_layers-1; for the P_*, FB_W and COL_* arrays
 * layer == n_layers addresses the INPUT prediction nodes
 * (sdr/IPredictiveRSDR.h:68-81) and their columns.  Arrays a variant does not
 * have report length 0.
 */
enum bidi_array {
  /* sparse coder of the layer (sdr/RSDR.h:15-41, sdr/IRSDR.h:21-49) */
  BIDI_VIS_INPUT = 0, BIDI_VIS_RECON,
  BIDI_HID_THRESHOLD, BIDI_HID_STATE, BIDI_HID_STATE_PREV, BIDI_HID_ACTIVATION, BIDI_HID_RECON,
  BIDI_HID_SPIKE, BIDI_HID_SPIKE_PREV,
  BIDI_FF_W, BIDI_REC_W, BIDI_LAT_W, BIDI_FF_TRACE, BIDI_REC_TRACE,
  /* prediction nodes (sdr/PredictiveRSDR.h:44-61) */
  BIDI_P_STATE, BIDI_P_STATE_PREV, BIDI_P_ACTIVATION, BIDI_P_ACTIVATION_PREV,
  BIDI_P_BASELINE, /* _averageSurprise (KWINNERS) or _baseline */
  BIDI_FB_W, BIDI_PRED_W,
  /* columns (neo/Column.h:9-58); concatenated node after node */
  BIDI_COL_INPUT, BIDI_COL_RECON_ERR,             /* [sum nIn] */
  BIDI_COL_FF_W,                                  /* [node][cell][input] */
  BIDI_COL_LAT_W,                                 /* [node][cell][cell] */
  BIDI_COL_THRESHOLD, BIDI_COL_ACTIVATION, BIDI_COL_SPIKE, BIDI_COL_SPIKE_PREV, BIDI_COL_STATE, /* [node][cell] */
  BIDI_COL_Q_W, BIDI_COL_Q_TRACE,                 /* [node][cell] */
  BIDI_COL_ACT_STATE, BIDI_COL_ACT_EXPL,          /* [node][action] */
  BIDI_COL_ACT_W, BIDI_COL_ACT_TRACE,             /* [node][action][cell] */
  BIDI_COL_PREV_VALUE,                            /* [node] */
  /* sdr::PredictiveRSDR::_prediction (sdr/PredictiveRSDR.h:77), layer 0 only */
  BIDI_PREDICTION,
  /* sdr::QPRSDR (sdr/QPRSDR.h:10-55).  Per Q layer (= hierarchy layer): */
  BIDI_Q_STATE, BIDI_Q_ERROR, BIDI_Q_BIAS_W, BIDI_Q_BIAS_TRACE, /* [node] */
  BIDI_Q_FF_W, BIDI_Q_FF_TRACE,                  /* kept connections only, node after node, list order */
  /* layer 0 only: */
  BIDI_QCONN_W, BIDI_QCONN_TRACE,                /* _qConnections, one per node of the top layer */
  BIDI_ACT_PREDICTED, BIDI_ACT_DERIVE, BIDI_ACT_EXPL, BIDI_ACT_ERROR, /* _actionNodes: actions, then anti-actions */
  BIDI_Q_PREV_VALUE,                             /* [1] */
  /* deep::CSRL (deep/CSRL.h:77-136) and its deep::SDRRL units (deep/SDRRL.h:8-62).  CSRL reuses the names above:
   * P_ACTIVATION(_PREV) = _state(Prev), P_STATE(_PREV) = _stateOutput(Prev); the COL_* arrays are the SDRRL unit of
   * every node ([node][cell]... as for neo::Column, except COL_ACT_W / COL_ACT_TRACE = _actionConnections
   * [node][cell][action] and COL_ACT_STATE / COL_ACT_EXPL [node][2*(3 + recurrent inputs)]) */
  BIDI_FB_TRACE, BIDI_PRED_TRACE,                /* traces of the feedback / predictive connections */
  BIDI_P_LOCAL_REWARD,                           /* [node] _localReward */
  BIDI_COL_ACT_ERROR,                            /* [node][action] Action::_error */
  BIDI_COL_CELL_ACT_STATE, BIDI_COL_CELL_ACT_ERROR, /* [node][cell] Cell::_actionState, _actionError */
  BIDI_COL_AVG_SURPRISE,                         /* [node] */
  BIDI_REWARD_OFFSETS,                           /* _lastLayerRewardOffsets, layer n_layers-1 only */
  BIDI_ARRAY_COUNT
};

typedef struct bidi_engine bidi_engine;

/* Message of the last failed call; h may be NULL for a failed bidi_create. */
const char* bidi_last_error(const bidi_engine* h);

/*
 * Allocate an engine for n_agents independent hierarchies of identical shape on
 * CUDA device `device` (state zero-initialised as the reference's constructors
 * do, thresholds = cfg->init_threshold).  Replaces the allocation half of
 * createRandom (sdr/PredictiveRSDR.cpp:8, sdr/IPredictiveRSDR.cpp:8,
 * neo/PredictiveHierarchy.cpp:8, neo/Agent.cpp:7).  Fails with BIDI_ENODEV when
 * no CUDA device is usable.
 */
int bidi_create(const bidi_config* cfg, const bidi_layer_desc* layers, int n_agents, int device,
                bidi_engine** out);
void bidi_destroy(bidi_engine* h);

/*
 * Random half of createRandom: agent a (first <= a < first+count) gets exactly
 * the weights std::mt19937(seed_base + a) produces in the reference's draw
 * order (sdr/RSDR.cpp:48-116, sdr/PredictiveRSDR.cpp:38-91,
 * neo/Column.cpp:22-43).
 */
int bidi_init_random(bidi_engine* h, int first, int count, unsigned seed_base);
/*
 * The same replay for ONE agent, drawing from the CALLER's generator: `state_in` is the text form of a std::mt19937
 * (operator<<: 624 words and the position), `state_out` (capacity state_out_cap, 8 KiB is enough; may be NULL) receives
 * the generator's state after createRandom's draws, so that a caller-side std::mt19937 continues exactly where the
 * reference's would (deep/CSRL.cpp:11-187 draws ~1e5 values per hierarchy in an order only the engine tabulates).
 */
int bidi_init_random_generator(bidi_engine* h, int agent, const char* state_in, char* state_out, int state_out_cap);

/* Shape queries. */
int bidi_num_agents(const bidi_engine* h);
int bidi_num_inputs(const bidi_engine* h);          /* in_width*in_height */
long long bidi_array_len(const bidi_engine* h, int which, int layer); /* <0 on error */
/*
 * Per-unit list lengths of a connection-list array (FF_W, REC_W, LAT_W, FB_W,
 * PRED_W: the clipped window sizes; COL_INPUT: nIn per node).  `counts` has one
 * entry per unit of the layer.  What the reference exposes through
 * getHiddenNode(i)._feedForwardConnections.size() (sdr/RSDR.h:116).
 */
int bidi_conn_counts(const bidi_engine* h, int which, int layer, int* counts, int n_units);

/* Serialised-state access in reference order. */
int bidi_set_array(bidi_engine* h, int agent, int which, int layer, const float* src, long long len);
int bidi_get_array(bidi_engine* h, int agent, int which, int layer, float* dst, long long len);

/*
 * setInput(index, value) for every input of every agent at once
 * (sdr/PredictiveRSDR.h:84, sdr/IPredictiveRSDR.h:111, neo/Agent.h:150):
 * in[agent][index], host memory; copied host->device on the engine's stream.
 */
int bidi_set_inputs(bidi_engine* h, const float* in);
/*
 * Zero-copy variant of the two host-buffer calls: the engine's own page-locked staging buffers,
 * [agent][index] like the arguments of bidi_set_inputs / bidi_get_predictions.  A caller that
 * writes its inputs into bidi_host_inputs() and passes that pointer to bidi_set_inputs (or passes
 * bidi_host_predictions() to bidi_get_predictions) saves the staging memcpy.  The input buffer may
 * be rewritten once bidi_wait_inputs returns (the host->device copy has read it); the prediction
 * buffer is valid until the next bidi_get_predictions.
 */
/* The layer-0 inputs as the last step consumed them (after bidi_set_feedback's taps), out[agent][index], host. */
int bidi_get_inputs(bidi_engine* h, float* out);
float* bidi_host_inputs(bidi_engine* h);
const float* bidi_host_predictions(bidi_engine* h);
int bidi_wait_inputs(bidi_engine* h);
int bidi_set_input(bidi_engine* h, int agent, int index, float value);

/*
 * simStep (sdr/PredictiveRSDR.h:82, sdr/IPredictiveRSDR.h:109, neo/Agent.h:148)
 * for all agents: one pass bottom-up coding -> top-down prediction -> learning.
 *   rewards  [n_agents] host, NEO_AGENT and QPRSDR (NULL = 0)
 *   noise_u, noise_v  [n_agents][bidi_num_noise] host; NEO_AGENT ([n_columns][3]): the values
 *            the reference draws per column visit (neo/Column.cpp:126-131):
 *            u = dist01(gen); v = u < breakChance ? dist01(gen) : pertDist(gen).
 *            Column visit order = layers top -> bottom, nodes ascending, then the
 *            input nodes (neo/Agent.cpp:155-156,212).  NULL = the engine draws
 *            from its own counter-based per-column generator.
 *   learn    the reference's `learn` flag.
 * Asynchronous: returns after enqueueing on the engine's stream.
 */
int bidi_step(bidi_engine* h, const float* rewards, const float* noise_u, const float* noise_v,
              int learn);
/*
 * neo::PredictiveHierarchy::simStepGenerate(generator, noise) (neo/PredictiveHierarchy.h:110,
 * neo/PredictiveHierarchy.cpp:247-315): a step WITHOUT learning in which every coder unit's
 * excitation starts, in every iteration, from noiseDist(generator) * noise instead of 0
 * (SparseCoder::activateNoise, neo/SparseCoder.cpp:178-240).  BIDI_NEO_HIERARCHY only.
 *   normals  [n_agents][bidi_num_generate_noise] host: the N(0,1) values in the reference's draw
 *            order -- layer 0 first, per layer iteration outer / unit inner, one
 *            std::normal_distribution per layer; NULL = the engine's own counter-based generator.
 */
int bidi_num_generate_noise(const bidi_engine* h);
int bidi_step_generate(bidi_engine* h, const float* normals, float noise);
int bidi_num_columns(const bidi_engine* h);
/*
 * Floats per agent in each of bidi_step's noise_u / noise_v: 3 per column for NEO_AGENT;
 * one per action for QPRSDR, whose simStep draws, per action i, u = dist01(gen) and then
 * v = u < _explorationBreak ? dist01(gen) : pertDist(gen) (sdr/QPRSDR.cpp:205-213).
 */
int bidi_num_noise(const bidi_engine* h);
/* sdr::QPRSDR::getActionRel(i) (sdr/QPRSDR.h:124) of every action node: out[agent][2*q_n_actions]. */
int bidi_get_actions(bidi_engine* h, float* out);
/*
 * QPRSDR: which of a layer's feed-forward window slots (the in-bounds ones, list order, node after
 * node -- bidi_conn_counts(BIDI_FF_W) of them per node) the Q network keeps: createRandom draws a
 * weight for every such slot and keeps it only on a path from an action input (sdr/QPRSDR.cpp:70-81).
 */
int bidi_q_kept(const bidi_engine* h, int layer, unsigned char* kept, long long len);

/*
 * getPrediction(index) for every input of every agent (sdr/PredictiveRSDR.h:92,
 * sdr/IPredictiveRSDR.h:119, neo/Agent.h:158): out[agent][index], host memory.
 * Synchronises the engine's stream.
 */
int bidi_get_predictions(bidi_engine* h, float* out);
/* Column::getAction(a) (neo/Column.h:85) of every column: out[agent][column][3]. */
int bidi_get_column_actions(bidi_engine* h, float* out);

int bidi_sync(bidi_engine* h);

/*
 * Device-resident stepping for throughput runs: frames[T][agent][index] (and
 * optionally rewards[T][agent]) are uploaded once; bidi_step_frame(t) is
 * bidi_set_inputs(frame t % T) + bidi_step without any host<->device traffic
 * (exploration noise from the engine's own generator).
 */
int bidi_load_frames(bidi_engine* h, const float* frames, const float* rewards, int n_frames);
int bidi_step_frame(bidi_engine* h, int t, int learn);
/*
 * Exploration draws for device-resident stepping: noise_u / noise_v [n_frames][n_agents][bidi_num_noise],
 * the same values bidi_step takes per call (neo/Column.cpp:126-131, sdr/QPRSDR.cpp:205-213).  Once loaded,
 * bidi_step_frame(t) uses frame t % n_frames of them instead of the engine's own generator, so that the
 * device-resident path can be held bit for bit against the host-driven one.  n_frames = 0 unloads.
 */
int bidi_load_frame_noise(bidi_engine* h, const float* noise_u, const float* noise_v, int n_frames);
/*
 * The engine's own counter-based generator (exploration, simStepGenerate and feedback noise when the caller
 * supplies none) is keyed by (seed, step, first_global_agent + local agent, ...): engines that hold different
 * shards of one population pass their first global agent index so that no two agents share a noise stream.
 * Part of the snapshot.
 */
int bidi_set_noise_stream(bidi_engine* h, unsigned long long seed, long long first_global_agent);
/* The exploration draws the last step consumed, [n_agents][bidi_num_noise] each (either may be NULL). */
int bidi_get_noise(bidi_engine* h, float* noise_u, float* noise_v);

/*
 * The caller's half of the RunnerMain loop for device-resident runs
 * (RunnerMain.cpp:241-249): before every bidi_step_frame, for each tap k,
 * input[dst[k]] = clamp01(prediction[src[k]]*scale + bias + N(0,noise_std)); with
 * reward_from_actions the step's reward is mean_k clamp01(prediction[src[k]]*scale
 * + bias) - 0.5, the headless stand-in for the Box2D body velocity.  n = 0 disables.
 */
int bidi_set_feedback(bidi_engine* h, int n, const int* src, const int* dst, float scale, float bias,
                      float noise_std, int reward_from_actions);

/*
 * A batched headless environment on the device, one instance per agent, so that the RunnerMain loop runs closed
 * without leaving the GPU.  BIDI_ENV_RUNNER stands in for the Box2D quadruped of runner/Runner.cpp
 * (not installable here): the same interface -- Runner::getStateVector's 15 floats (runner/Runner.cpp:260-339: ten
 * joint angles, the body angle, four foot contacts), Runner::motorUpdate's ten targets (:341), reward = body
 * velocity x (RunnerMain.cpp:207-210) -- over a small deterministic planar surrogate of the rigid-body world
 * (bidi_env.cuh).  bidi_env_step is one pass of RunnerMain.cpp:204-251 for every agent:
 *   reward <- body vx; inputs 0..14 <- state, 15..18 <- the four clock sines, 20..29 <- clamp01(prediction*0.5 +
 *   0.5 + N(0, 0.05)); simStep(reward); action <- clamp01(prediction*0.5 + 0.5); motorUpdate; physics tick.
 * noise_u / noise_v as for bidi_step; env_noise [n_agents][10] host: the N(0, 0.05) draws of the action feedback
 * (NULL: the engine's generator).  The engine needs at least 30 inputs.  bidi_env_get_state: out[agent][24] =
 * 10 joint angles, body angle, 4 contacts, body vx, angular velocity, body x, 4 foot positions, pad.
 */
#define BIDI_ENV_RUNNER 1
#define BIDI_ENV_STATE_FLOATS 24
int bidi_env_attach(bidi_engine* h, int kind);
int bidi_env_reset(bidi_engine* h);
int bidi_env_step(bidi_engine* h, const float* noise_u, const float* noise_v, const float* env_noise, int learn);
int bidi_env_get_state(bidi_engine* h, float* out);

/*
 * Kernel-level timing inside the step graph: the launches of the named kernel
 * ("columns", "ic_sweep", "reconstruct", "predict", ...) are bracketed by CUDA
 * events on the engine's stream.  bidi_profile_collect, called after a step,
 * adds that step's launches to the running totals (it synchronises).
 */
int bidi_profile_select(bidi_engine* h, const char* kernel_name);
int bidi_profile_collect(bidi_engine* h, double* total_ms, long long* launches);

/*
 * Checkpoint / resume: every agent's complete state (weights, traces, unit state, inputs,
 * predictions, the step counter of the engine's own exploration generator) to one file and back
 * into an engine created with the SAME configuration and agent count.  The reference has no such
 * facility for its hierarchies.  A resumed run continues bit-identically.
 */
int bidi_save_snapshot(bidi_engine* h, const char* path);
int bidi_load_snapshot(bidi_engine* h, const char* path);

/*
 * Engine options: which of several kernel families, all with identical results, runs a layer (the parity tests hold
 * every one to the oracle; the defaults are the fastest choice per layer shape).
 *   "force_phase_kernels" 1: one kernel per phase even where a fused per-agent coder kernel applies
 *   "dense_coder"         0: the general fused coder kernel where the dense-row one was chosen
 *   "tile_kernels"        0: thread-per-unit kernels where the tile kernels (small launches) were chosen
 *   "grid_kernels"        0 never / 1 wide, untiled k-winners layers (default) / 2 every k-winners layer
 *   "persist_coder"       0: per-phase kernels for the large layers of a single iterative agent
 *   "tma_box"             0: k_g_kw_activate stages its input box with plain loads instead of the tiled TMA load
 *   "overlap_learn"       0: learning and reconstruction phases stay on the step's main graph branch
 *   "strip_graph"         0: a split layer's step is issued launch by launch instead of as one graph
 * Changing an option re-records the step graph and converts the affected layers' weight layout in place.
 */
int bidi_set_option(bidi_engine* h, const char* name, int value);

/*
 * Spatial split of ONE oversized layer over several GPUs (BASELINE config 5;
 * the reference cannot even index such a layer: its connection indices are unsigned
 * short, sdr/RSDR.h:10).  A single-layer BIDI_KWINNERS hierarchy is cut into n_parts
 * horizontal row strips; part p owns hidden rows [p*H/n, (p+1)*H/n) and the matching
 * visible rows.  Weights never move: a unit's weights live on the part that owns the
 * unit (plus identical, redundantly updated copies of the few halo rows whose weights
 * the gather-form reconstruction of the owned rows reads).  Per step the parts exchange
 * halo rows of five unit-state planes (activations before the k-winner rank, winners,
 * reconstructions, prediction activations, predicted winners); results are bit-identical
 * to the unsplit engine.
 *
 * Every part is a full bidi_engine created with the SAME config on its own device and
 * given the same inputs (bidi_set_inputs / bidi_load_frames take the whole frame).
 * bidi_strip_configure must precede bidi_init_random / bidi_set_array.  Two transports:
 *   - same process (one host thread drives all parts, any mix of devices, peer copies
 *     over NVLink): bidi_strip_link_local, then bidi_strip_step_local per step;
 *   - one process per GPU (torchrun), peer memory: every rank FIRST initialises its state
 *     (bidi_init_random / bidi_set_array -- refused once linked, because from then on the
 *     neighbours store into this memory at their own pace), then calls bidi_strip_ipc_export,
 *     the 128-byte records are gathered to all ranks (any host channel), then
 *     bidi_strip_link_ipc and plain bidi_step / bidi_step_frame.  One small kernel per
 *     exchange point stores the halo rows straight into the neighbours' memory over
 *     NVLink, raises a flag there and waits for theirs; the whole step is one CUDA graph;
 *   - one process per GPU, NCCL: bidi_strip_nccl_unique_id on rank 0, ship the 128 bytes
 *     to every rank, bidi_strip_link_nccl; halos go through ncclSend/ncclRecv on the
 *     engine's stream (also inside the step's graph).
 * bidi_get_predictions writes ONLY the rows the part owns into the caller's whole-grid array (the
 * rest is left untouched), bidi_set_inputs reads only the rows its windows reach, and
 * bidi_get_array returns whole-grid arrays of which only the owned rows are meaningful
 * (bidi_strip_plan tells which).
 */
enum bidi_strip_span { /* row intervals [lo,hi) of part p, see bidi_strip_plan */
  BIDI_SPAN_OWN_HID = 0,  /* hidden rows the part owns */
  BIDI_SPAN_OWN_VIS,      /* visible (input) rows the part owns */
  BIDI_SPAN_EXT_HID,      /* hidden rows whose weights it holds and updates (owned + halo) */
  BIDI_SPAN_STATE,        /* hidden rows on which it keeps the winners (state) valid */
  BIDI_SPAN_ACT,          /* hidden rows of activations it needs for the k-winner rank */
  BIDI_SPAN_VIS_RECON,    /* visible rows of the reconstruction its learning rule reads */
  BIDI_SPAN_HID_RECON,    /* hidden rows of the recurrent reconstruction it reads */
  BIDI_SPAN_PRED_STATE,   /* hidden rows of predicted winners the input prediction reads */
  BIDI_SPAN_COUNT
};
/* Pure geometry, no device needed: spans[BIDI_SPAN_COUNT][2] of part `part` of `n_parts`. */
int bidi_strip_plan(const bidi_config* cfg, const bidi_layer_desc* layers, int part, int n_parts, int* spans);
int bidi_strip_configure(bidi_engine* h, int part, int n_parts);
int bidi_strip_link_local(bidi_engine** parts, int n_parts);
int bidi_strip_step_local(bidi_engine** parts, int n_parts, int learn);
/* frames variant of the above: every part steps on frame t of its loaded frames */
int bidi_strip_step_frame_local(bidi_engine** parts, int n_parts, int t, int learn);
int bidi_strip_ipc_export(bidi_engine* h, void* handle128);
int bidi_strip_link_ipc(bidi_engine* h, const void* handles /* [n_parts][128] */);
int bidi_strip_nccl_unique_id(void* id128);
int bidi_strip_link_nccl(bidi_engine* h, const void* id128);
/* floats this part sends per step over the transport (all exchange points, learn on) */
long long bidi_strip_halo_floats(const bidi_engine* h);

/*
 * Many hierarchies over several devices of ONE process.  In the reference a second agent is a
 * second object (RunnerMain.cpp:100-106) and agents never interact, so the population of n_agents identical
 * hierarchies is cut into contiguous blocks: shard k = global agents [k*n/n_devices, (k+1)*n/n_devices) on
 * CUDA device devices[k] (a device may repeat), one bidi_engine each.  Every call fans out: the shards
 * enqueue on their own devices' streams and run concurrently; there is no data-path collective.  All arrays
 * are indexed by GLOBAL agent: in / out [n_agents][n_in], rewards [n_agents], noise [n_agents][bidi_num_noise],
 * frames [n_frames][n_agents][n_in].  Agent g is created from std::mt19937(seed_base + g) and draws the
 * engine's own exploration noise from stream g (bidi_set_noise_stream), whatever the number of shards.
 * bidi_sharded_shard gives the engine that holds an agent range, e.g. for bidi_get_array / bidi_set_array
 * with the agent index local to the shard.
 */
typedef struct bidi_sharded bidi_sharded;
int bidi_create_sharded(const bidi_config* cfg, const bidi_layer_desc* layers, int n_agents, const int* devices, int n_devices,
                        bidi_sharded** out);
void bidi_sharded_destroy(bidi_sharded* s);
const char* bidi_sharded_last_error(const bidi_sharded* s);
int bidi_sharded_num_shards(const bidi_sharded* s);
int bidi_sharded_num_agents(const bidi_sharded* s);
bidi_engine* bidi_sharded_shard(bidi_sharded* s, int k, int* first_agent, int* n_agents);
int bidi_sharded_init_random(bidi_sharded* s, unsigned seed_base);
int bidi_sharded_set_inputs(bidi_sharded* s, const float* in);
int bidi_sharded_step(bidi_sharded* s, const float* rewards, const float* noise_u, const float* noise_v, int learn);
int bidi_sharded_get_predictions(bidi_sharded* s, float* out);
int bidi_sharded_get_column_actions(bidi_sharded* s, float* out);
int bidi_sharded_load_frames(bidi_sharded* s, const float* frames, const float* rewards, int n_frames);
int bidi_sharded_step_frame(bidi_sharded* s, int t, int learn);
int bidi_sharded_set_feedback(bidi_sharded* s, int n, const int* src, const int* dst, float scale, float bias, float noise_std,
                              int reward_from_actions);
int bidi_sharded_sync(bidi_sharded* s);
/* CUDA-event timing on every shard's stream; stop returns the slowest shard's milliseconds */
int bidi_sharded_timer_start(bidi_sharded* s);
int bidi_sharded_timer_stop_ms(bidi_sharded* s, float* ms);

/* Timing helpers on the engine's stream (CUDA events), for bench.py. */
int bidi_timer_start(bidi_engine* h);
int bidi_timer_stop_ms(bidi_engine* h, float* ms);
/* Number of kernel launches (graph kernel nodes included) issued so far. */
long long bidi_launch_count(const bidi_engine* h);
/* Bytes of device memory held by the engine. */
long long bidi_device_bytes(const bidi_engine* h);
/* sizeof(bidi_config), sizeof(bidi_layer_desc): lets a binding check its mirror. */
int bidi_abi_sizes(int* config_bytes, int* layer_desc_bytes);

#ifdef __cplusplus
}
#endif
#endif /* BIDINET_B200_H */
