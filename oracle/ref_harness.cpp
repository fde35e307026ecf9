// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Headless C wrapper around the UNMODIFIED reference classes, compiled from the
// sources where they lie under /root/reference (see oracle/Makefile) into
// oracle/_ref/libbidiref.so.  It gives the tests and bench.py's reference arm
//   - creation from a fixed seed (the mains seed from time(), RunnerMain.cpp:36)
//   - setInput / simStep / getPrediction exactly as the mains call them
//   - a dump of every state array in the serialised order include/bidinet_b200.h
//     defines, so the CUDA engine and the C restatement (oracle/bidi_oracle.c)
//     can be started from, and compared with, an identical state.
// Nothing here is copied from the reference; it only calls it.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>
#include <vector>
#include <iostream>

// The coders' node vectors, neo::Column internals and _inputPredictionNodes are
// private (sdr/RSDR.h:43-52, neo/Column.h:8-58, neo/Agent.h:116-120); access
// specifiers do not change layout, so this TU may see them as public.
#define private public
#include <sdr/PredictiveRSDR.h>
#include <sdr/IPredictiveRSDR.h>
#include <sdr/QPRSDR.h>
#include <neo/PredictiveHierarchy.h>
#include <neo/Agent.h>
#include <deep/CSRL.h>
#undef private

#include "bidinet_b200.h"

namespace {

struct Ref {
  bidi_config cfg;
  std::vector<bidi_layer_desc> ld;
  std::mt19937 gen;
  sdr::PredictiveRSDR kw;
  sdr::IPredictiveRSDR it;
  neo::PredictiveHierarchy nh;
  neo::Agent ag;
  sdr::QPRSDR qp;
  deep::CSRL cs;
};
// sdr::QPRSDR::simStep prints q on every step (sdr/QPRSDR.cpp:259); keep the harness quiet
struct QuietCout {
  std::streambuf* old;
  QuietCout() : old(std::cout.rdbuf(nullptr)) {}
  ~QuietCout() { std::cout.rdbuf(old); }
};
sdr::IPredictiveRSDR& hierarchy_of(Ref& r) { return r.cfg.variant == BIDI_QPRSDR ? r.qp._prsdr : r.it; }

template <class D> void fill_common(D& d, const bidi_layer_desc& s) {
  d._width = s.width; d._height = s.height;
  d._receptiveRadius = s.r_ff; d._recurrentRadius = s.r_rec; d._lateralRadius = s.r_lat;
  d._predictiveRadius = s.r_pred; d._feedBackRadius = s.r_fb;
  d._learnFeedForward = s.learn_ff; d._learnRecurrent = s.learn_rec; d._learnLateral = s.learn_lat;
  d._learnFeedBack = s.learn_fb; d._learnPrediction = s.learn_pred;
}
template <class D> void fill_neo(D& d, const bidi_layer_desc& s) {
  fill_common(d, s);
  d._sdrIter = s.iter_settle; d._sdrLeak = s.leak; d._sdrLambda = s.lambda;
  d._sdrWeightDecay = s.weight_decay; d._sdrMaxWeightDelta = s.max_weight_delta;
  d._sdrSparsity = s.sparsity; d._sdrLearnThreshold = s.learn_threshold;
  d._sdrBaselineDecay = s.baseline_decay; d._sdrSensitivity = s.sensitivity;
}

// ---- generic element visitors (serialised order of include/bidinet_b200.h) ----
template <class Coder, class F> bool visit_coder(Coder& c, int which, F&& f) {
  switch (which) {
    case BIDI_VIS_INPUT: for (auto& v : c._visible) f(v._input); return true;
    case BIDI_VIS_RECON: for (auto& v : c._visible) f(v._reconstruction); return true;
    case BIDI_HID_THRESHOLD: for (auto& h : c._hidden) f(h._threshold); return true;
    case BIDI_HID_STATE: for (auto& h : c._hidden) f(h._state); return true;
    case BIDI_HID_STATE_PREV: for (auto& h : c._hidden) f(h._statePrev); return true;
    case BIDI_HID_ACTIVATION: for (auto& h : c._hidden) f(h._activation); return true;
    case BIDI_HID_RECON: for (auto& h : c._hidden) f(h._reconstruction); return true;
    case BIDI_FF_W: for (auto& h : c._hidden) for (auto& k : h._feedForwardConnections) f(k._weight); return true;
    case BIDI_REC_W: for (auto& h : c._hidden) for (auto& k : h._recurrentConnections) f(k._weight); return true;
    default: return false;
  }
}
template <class Coder, class F> bool visit_icoder(Coder& c, int which, bool traces, F&& f) {
  if (visit_coder(c, which, f)) return true;
  switch (which) {
    case BIDI_HID_SPIKE: for (auto& h : c._hidden) f(h._spike); return true;
    case BIDI_HID_SPIKE_PREV: for (auto& h : c._hidden) f(h._spikePrev); return true;
    case BIDI_LAT_W: for (auto& h : c._hidden) for (auto& k : h._lateralConnections) f(k._weight); return true;
    case BIDI_FF_TRACE: if (!traces) return false;
      for (auto& h : c._hidden) for (auto& k : h._feedForwardConnections) f(k._trace); return true;
    case BIDI_REC_TRACE: if (!traces) return false;
      for (auto& h : c._hidden) for (auto& k : h._recurrentConnections) f(k._trace); return true;
    default: return false;
  }
}
template <class Nodes, class F> bool visit_pnodes(Nodes& nodes, int which, F&& f) {
  switch (which) {
    case BIDI_P_STATE: for (auto& p : nodes) f(p._state); return true;
    case BIDI_P_STATE_PREV: for (auto& p : nodes) f(p._statePrev); return true;
    case BIDI_P_ACTIVATION: for (auto& p : nodes) f(p._activation); return true;
    case BIDI_P_ACTIVATION_PREV: for (auto& p : nodes) f(p._activationPrev); return true;
    case BIDI_FB_W: for (auto& p : nodes) for (auto& k : p._feedBackConnections) f(k._weight); return true;
    default: return false;
  }
}
template <class Nodes, class F> bool visit_pred_w(Nodes& nodes, F&& f) {
  for (auto& p : nodes) for (auto& k : p._predictiveConnections) f(k._weight);
  return true;
}
template <class Nodes, class F> bool visit_columns(Nodes& nodes, int which, F&& f) {
  switch (which) {
    case BIDI_COL_INPUT: for (auto& p : nodes) for (auto& v : p._column._inputs) f(v); return true;
    case BIDI_COL_RECON_ERR: for (auto& p : nodes) for (auto& v : p._column._reconstructionError) f(v); return true;
    case BIDI_COL_FF_W: for (auto& p : nodes) for (auto& c : p._column._cells) for (auto& k : c._feedForwardConnections) f(k._weight); return true;
    case BIDI_COL_LAT_W: for (auto& p : nodes) for (auto& c : p._column._cells) for (auto& k : c._lateralConnections) f(k._weight); return true;
    case BIDI_COL_THRESHOLD: for (auto& p : nodes) for (auto& c : p._column._cells) f(c._threshold); return true;
    case BIDI_COL_ACTIVATION: for (auto& p : nodes) for (auto& c : p._column._cells) f(c._activation); return true;
    case BIDI_COL_SPIKE: for (auto& p : nodes) for (auto& c : p._column._cells) f(c._spike); return true;
    case BIDI_COL_SPIKE_PREV: for (auto& p : nodes) for (auto& c : p._column._cells) f(c._spikePrev); return true;
    case BIDI_COL_STATE: for (auto& p : nodes) for (auto& c : p._column._cells) f(c._state); return true;
    case BIDI_COL_Q_W: for (auto& p : nodes) for (auto& k : p._column._qConnections) f(k._weight); return true;
    case BIDI_COL_Q_TRACE: for (auto& p : nodes) for (auto& k : p._column._qConnections) f(k._trace); return true;
    case BIDI_COL_ACT_STATE: for (auto& p : nodes) for (auto& a : p._column._actions) f(a._state); return true;
    case BIDI_COL_ACT_EXPL: for (auto& p : nodes) for (auto& a : p._column._actions) f(a._exploratoryState); return true;
    case BIDI_COL_ACT_W: for (auto& p : nodes) for (auto& a : p._column._actions) for (auto& k : a._connections) f(k._weight); return true;
    case BIDI_COL_ACT_TRACE: for (auto& p : nodes) for (auto& a : p._column._actions) for (auto& k : a._connections) f(k._trace); return true;
    case BIDI_COL_PREV_VALUE: for (auto& p : nodes) f(p._column._prevValue); return true;
    default: return false;
  }
}

// deep::CSRL's nodes: P_ACTIVATION(_PREV) = _state(Prev), P_STATE(_PREV) = _stateOutput(Prev); the COL_* arrays
// are the node's deep::SDRRL unit
template <class Nodes, class F> bool visit_csrl_nodes(Nodes& nodes, int which, bool has_pred, F&& f) {
  switch (which) {
    case BIDI_P_STATE: for (auto& p : nodes) f(p._stateOutput); return true;
    case BIDI_P_STATE_PREV: for (auto& p : nodes) f(p._stateOutputPrev); return true;
    case BIDI_P_ACTIVATION: for (auto& p : nodes) f(p._state); return true;
    case BIDI_P_ACTIVATION_PREV: for (auto& p : nodes) f(p._statePrev); return true;
    case BIDI_P_BASELINE: if (!has_pred) return false; for (auto& p : nodes) f(p._baseline); return true;
    case BIDI_P_LOCAL_REWARD: for (auto& p : nodes) f(p._localReward); return true;
    case BIDI_FB_W: for (auto& p : nodes) for (auto& k : p._feedBackConnections) f(k._weight); return true;
    case BIDI_FB_TRACE: for (auto& p : nodes) for (auto& k : p._feedBackConnections) f(k._trace); return true;
    case BIDI_COL_INPUT: for (auto& p : nodes) for (auto& v : p._sdrrl._inputs) f(v); return true;
    case BIDI_COL_RECON_ERR: for (auto& p : nodes) for (auto& v : p._sdrrl._reconstructionError) f(v); return true;
    case BIDI_COL_FF_W: for (auto& p : nodes) for (auto& c : p._sdrrl._cells) for (auto& k : c._feedForwardConnections) f(k._weight); return true;
    case BIDI_COL_LAT_W: for (auto& p : nodes) for (auto& c : p._sdrrl._cells) for (auto& k : c._lateralConnections) f(k._weight); return true;
    case BIDI_COL_THRESHOLD: for (auto& p : nodes) for (auto& c : p._sdrrl._cells) f(c._threshold); return true;
    case BIDI_COL_ACTIVATION: for (auto& p : nodes) for (auto& c : p._sdrrl._cells) f(c._activation); return true;
    case BIDI_COL_SPIKE: for (auto& p : nodes) for (auto& c : p._sdrrl._cells) f(c._spike); return true;
    case BIDI_COL_SPIKE_PREV: for (auto& p : nodes) for (auto& c : p._sdrrl._cells) f(c._spikePrev); return true;
    case BIDI_COL_STATE: for (auto& p : nodes) for (auto& c : p._sdrrl._cells) f(c._state); return true;
    case BIDI_COL_CELL_ACT_STATE: for (auto& p : nodes) for (auto& c : p._sdrrl._cells) f(c._actionState); return true;
    case BIDI_COL_CELL_ACT_ERROR: for (auto& p : nodes) for (auto& c : p._sdrrl._cells) f(c._actionError); return true;
    case BIDI_COL_Q_W: for (auto& p : nodes) for (auto& k : p._sdrrl._qConnections) f(k._weight); return true;
    case BIDI_COL_Q_TRACE: for (auto& p : nodes) for (auto& k : p._sdrrl._qConnections) f(k._trace); return true;
    case BIDI_COL_ACT_STATE: for (auto& p : nodes) for (auto& a : p._sdrrl._actions) f(a._state); return true;
    case BIDI_COL_ACT_EXPL: for (auto& p : nodes) for (auto& a : p._sdrrl._actions) f(a._exploratoryState); return true;
    case BIDI_COL_ACT_ERROR: for (auto& p : nodes) for (auto& a : p._sdrrl._actions) f(a._error); return true;
    case BIDI_COL_ACT_W: for (auto& p : nodes) for (auto& c : p._sdrrl._cells) for (auto& k : c._actionConnections) f(k._weight); return true;
    case BIDI_COL_ACT_TRACE: for (auto& p : nodes) for (auto& c : p._sdrrl._cells) for (auto& k : c._actionConnections) f(k._trace); return true;
    case BIDI_COL_PREV_VALUE: for (auto& p : nodes) f(p._sdrrl._prevValue); return true;
    case BIDI_COL_AVG_SURPRISE: for (auto& p : nodes) f(p._sdrrl._averageSurprise); return true;
    default: return false;
  }
}
// deep::SDRRL leaves Cell::_activation/_spike/_state/_threshold-independent scratch and Action::_error uninitialised until
// the first simStep (deep/SDRRL.h:25-41); zero them so a dump taken before the first step is defined
template <class Nodes> void zero_sdrrl_scratch(Nodes& nodes) {
  for (auto& p : nodes) {
    for (auto& c : p._sdrrl._cells) { c._activation = 0.0f; c._spike = 0.0f; c._state = 0.0f; }
    for (auto& a : p._sdrrl._actions) a._error = 0.0f;
  }
}

template <class F> bool visit(Ref& r, int which, int layer, F&& f) {
  const int L = r.cfg.n_layers;
  if (layer < 0 || layer > L) return false;
  switch (r.cfg.variant) {
    case BIDI_CSRL: {
      deep::CSRL& cs = r.cs;
      if (which == BIDI_REWARD_OFFSETS) {
        if (layer != L - 1) return false;
        for (auto& v : cs._lastLayerRewardOffsets) f(v);
        return true;
      }
      if (layer == L) return visit_csrl_nodes(cs._inputPredictionNodes, which, false, f);
      auto& ly = cs._layers[layer];
      if (visit_icoder(ly._sdr, which, true, f)) return true;
      if (which == BIDI_PRED_W) return visit_pred_w(ly._predictionNodes, f);
      if (which == BIDI_PRED_TRACE) { for (auto& p : ly._predictionNodes) for (auto& k : p._predictiveConnections) f(k._trace); return true; }
      return visit_csrl_nodes(ly._predictionNodes, which, true, f);
    }
    case BIDI_KWINNERS: {
      if (which == BIDI_PREDICTION) {
        if (layer != 0) return false;
        for (auto& v : r.kw._prediction) f(v);
        return true;
      }
      if (layer >= L) return false;
      auto& ly = const_cast<sdr::PredictiveRSDR::Layer&>(r.kw.getLayers()[layer]);
      if (visit_coder(ly._sdr, which, f)) return true;
      if (which == BIDI_P_BASELINE) { for (auto& p : ly._predictionNodes) f(p._averageSurprise); return true; }
      if (which == BIDI_PRED_W) return visit_pred_w(ly._predictionNodes, f);
      return visit_pnodes(ly._predictionNodes, which, f);
    }
    case BIDI_QPRSDR:
      if (which >= BIDI_Q_STATE) {
        sdr::QPRSDR& q = r.qp;
        if (layer < L) {
          auto& nodes = q._qFunctionLayers[layer]._qFunctionNodes;
          switch (which) {
            case BIDI_Q_STATE: for (auto& n : nodes) f(n._state); return true;
            case BIDI_Q_ERROR: for (auto& n : nodes) f(n._error); return true;
            case BIDI_Q_BIAS_W: for (auto& n : nodes) f(n._bias._weight); return true;
            case BIDI_Q_BIAS_TRACE: for (auto& n : nodes) f(n._bias._trace); return true;
            case BIDI_Q_FF_W: for (auto& n : nodes) for (auto& c : n._feedForwardConnections) f(c._weight); return true;
            case BIDI_Q_FF_TRACE: for (auto& n : nodes) for (auto& c : n._feedForwardConnections) f(c._trace); return true;
          }
        }
        if (layer != 0) return false;
        switch (which) {
          case BIDI_QCONN_W: for (auto& c : q._qConnections) f(c._weight); return true;
          case BIDI_QCONN_TRACE: for (auto& c : q._qConnections) f(c._trace); return true;
          case BIDI_ACT_PREDICTED: for (auto& a : q._actionNodes) f(a._predictedAction); return true;
          case BIDI_ACT_DERIVE: for (auto& a : q._actionNodes) f(a._deriveAction); return true;
          case BIDI_ACT_EXPL: for (auto& a : q._actionNodes) f(a._exploratoryAction); return true;
          case BIDI_ACT_ERROR: for (auto& a : q._actionNodes) f(a._error); return true;
          case BIDI_Q_PREV_VALUE: f(q._prevValue); return true;
        }
        return false;
      }
      /* fall through: the hierarchy the QPRSDR owns */
    case BIDI_ITERATIVE: {
      sdr::IPredictiveRSDR& it = hierarchy_of(r);
      if (which >= BIDI_Q_STATE) return false;
      if (layer == L) return visit_pnodes(it._inputPredictionNodes, which, f);
      auto& ly = const_cast<sdr::IPredictiveRSDR::Layer&>(it.getLayers()[layer]);
      if (visit_icoder(ly._sdr, which, false, f)) return true;
      if (which == BIDI_P_BASELINE) { for (auto& p : ly._predictionNodes) f(p._baseline); return true; }
      if (which == BIDI_PRED_W) return visit_pred_w(ly._predictionNodes, f);
      return visit_pnodes(ly._predictionNodes, which, f);
    }
    case BIDI_NEO_HIERARCHY: {
      if (layer == L) return visit_pnodes(r.nh._inputPredictionNodes, which, f);
      auto& ly = const_cast<neo::PredictiveHierarchy::Layer&>(r.nh.getLayers()[layer]);
      if (visit_icoder(ly._sdr, which, true, f)) return true;
      if (which == BIDI_P_BASELINE) { for (auto& p : ly._predictionNodes) f(p._baseline); return true; }
      if (which == BIDI_PRED_W) return visit_pred_w(ly._predictionNodes, f);
      return visit_pnodes(ly._predictionNodes, which, f);
    }
    case BIDI_NEO_AGENT: {
      if (layer == L) {
        if (visit_pnodes(r.ag._inputPredictionNodes, which, f)) return true;
        return visit_columns(r.ag._inputPredictionNodes, which, f);
      }
      auto& ly = const_cast<neo::Agent::Layer&>(r.ag.getLayers()[layer]);
      if (visit_icoder(ly._sdr, which, true, f)) return true;
      if (which == BIDI_P_BASELINE) { for (auto& p : ly._predictionNodes) f(p._baseline); return true; }
      if (which == BIDI_PRED_W) return visit_pred_w(ly._predictionNodes, f);
      if (visit_pnodes(ly._predictionNodes, which, f)) return true;
      return visit_columns(ly._predictionNodes, which, f);
    }
  }
  return false;
}

// neo::Column leaves Cell::_activation/_spike/_state and Action::_error
// uninitialised until the first simStep overwrites them (neo/Column.h:31-33);
// zero them so a dump taken before the first step is defined.
template <class Nodes> void zero_column_scratch(Nodes& nodes) {
  for (auto& p : nodes) {
    for (auto& c : p._column._cells) { c._activation = 0.0f; c._spike = 0.0f; c._state = 0.0f; }
    for (auto& a : p._column._actions) a._error = 0.0f;
  }
}

} // namespace

extern "C" {

void* ref_create(const bidi_config* cfg, const bidi_layer_desc* layers, unsigned seed) {
  Ref* r = new Ref();
  r->cfg = *cfg;
  r->ld.assign(layers, layers + cfg->n_layers);
  r->gen.seed(seed);
  const int L = cfg->n_layers;
  switch (cfg->variant) {
    case BIDI_KWINNERS: {
      std::vector<sdr::PredictiveRSDR::LayerDesc> d(L);
      for (int l = 0; l < L; l++) {
        fill_common(d[l], layers[l]);
        d[l]._learnThreshold = layers[l].learn_threshold;
        d[l]._averageSurpriseDecay = layers[l].avg_surprise_decay;
        d[l]._attentionFactor = layers[l].attention_factor;
        d[l]._sparsity = layers[l].sparsity;
      }
      r->kw.createRandom(cfg->in_width, cfg->in_height, d, cfg->init_min_w, cfg->init_max_w,
                         cfg->init_threshold, r->gen);
      break;
    }
    case BIDI_QPRSDR:
    case BIDI_ITERATIVE: {
      std::vector<sdr::IPredictiveRSDR::LayerDesc> d(L);
      for (int l = 0; l < L; l++) {
        fill_common(d[l], layers[l]);
        d[l]._sdrIterSettle = layers[l].iter_settle; d[l]._sdrIterMeasure = layers[l].iter_measure;
        d[l]._sdrLeak = layers[l].leak; d[l]._sdrLambda = layers[l].lambda;
        d[l]._sdrWeightDecay = layers[l].weight_decay; d[l]._sdrMaxWeightDelta = layers[l].max_weight_delta;
        d[l]._sdrSparsity = layers[l].sparsity; d[l]._sdrLearnThreshold = layers[l].learn_threshold;
        d[l]._sdrBaselineDecay = layers[l].baseline_decay; d[l]._sdrSensitivity = layers[l].sensitivity;
      }
      if (cfg->variant == BIDI_QPRSDR) {
        sdr::QPRSDR& q = r->qp;
        q._prsdr._learnInputFeedBack = cfg->learn_infb;
        q._qAlpha = cfg->q_alpha; q._actionAlpha = cfg->q_action_alpha; q._actionDeriveIterations = cfg->q_derive_iterations;
        q._actionDeriveAlpha = cfg->q_derive_alpha; q._explorationBreak = cfg->q_expl_break; q._explorationStdDev = cfg->q_expl_stddev;
        q._gamma = cfg->q_gamma; q._gammaLambda = cfg->q_gamma_lambda;
        q._prevValue = 0.0f;  // not initialised by the reference (sdr/QPRSDR.h:66)
        std::vector<int> act(cfg->q_action_idx, cfg->q_action_idx + cfg->q_n_actions), anti(cfg->q_anti_idx, cfg->q_anti_idx + cfg->q_n_actions);
        q.createRandom(cfg->in_width, cfg->in_height, cfg->r_infb, act, anti, d, cfg->init_min_w, cfg->init_max_w,
                       cfg->init_min_inh, cfg->init_max_inh, cfg->init_threshold, r->gen);
        break;
      }
      r->it._learnInputFeedBack = cfg->learn_infb;
      r->it.createRandom(cfg->in_width, cfg->in_height, cfg->r_infb, d, cfg->init_min_w, cfg->init_max_w,
                         cfg->init_min_inh, cfg->init_max_inh, cfg->init_threshold, r->gen);
      break;
    }
    case BIDI_NEO_HIERARCHY: {
      std::vector<neo::PredictiveHierarchy::LayerDesc> d(L);
      for (int l = 0; l < L; l++) fill_neo(d[l], layers[l]);
      r->nh._learnInputFeedBack = cfg->learn_infb;
      r->nh.createRandom(cfg->in_width, cfg->in_height, cfg->r_infb, d, cfg->init_min_w, cfg->init_max_w,
                         cfg->init_min_inh, cfg->init_max_inh, cfg->init_threshold, r->gen);
      break;
    }
    case BIDI_NEO_AGENT: {
      std::vector<neo::Agent::LayerDesc> d(L);
      for (int l = 0; l < L; l++) {
        fill_neo(d[l], layers[l]);
        const bidi_layer_desc& s = layers[l];
        d[l]._cellsPerColumn = s.col_cells; d[l]._columnSparsity = s.col_sparsity; d[l]._columnIter = s.col_iter;
        d[l]._columnLeak = s.col_leak; d[l]._columnGamma = s.col_gamma; d[l]._columnGammaLambda = s.col_gamma_lambda;
        d[l]._columnFeedForwardAlpha = s.col_alpha_ff; d[l]._columnLateralAlpha = s.col_alpha_lat;
        d[l]._columnThresholdAlpha = s.col_alpha_threshold; d[l]._columnQAlpha = s.col_alpha_q;
        d[l]._columnActionAlpha = s.col_alpha_action; d[l]._columnExplorationStdDev = s.col_expl_stddev;
        d[l]._columnExplorationBreakChance = s.col_expl_break;
      }
      neo::Agent& a = r->ag;
      a._cellsPerColumn = cfg->col_cells; a._columnSparsity = cfg->col_sparsity; a._columnIter = cfg->col_iter;
      a._columnLeak = cfg->col_leak; a._columnGamma = cfg->col_gamma; a._columnGammaLambda = cfg->col_gamma_lambda;
      a._columnFeedForwardAlpha = cfg->col_alpha_ff; a._columnLateralAlpha = cfg->col_alpha_lat;
      a._columnThresholdAlpha = cfg->col_alpha_threshold; a._columnQAlpha = cfg->col_alpha_q;
      a._columnActionAlpha = cfg->col_alpha_action; a._columnExplorationStdDev = cfg->col_expl_stddev;
      a._columnExplorationBreakChance = cfg->col_expl_break;
      a._learnInputFeedBack = cfg->learn_infb;
      a.createRandom(cfg->in_width, cfg->in_height, cfg->r_infb, d, cfg->init_min_w, cfg->init_max_w,
                     cfg->init_min_inh, cfg->init_max_inh, cfg->init_threshold, r->gen);
      for (auto& ly : a._layers) zero_column_scratch(ly._predictionNodes);
      zero_column_scratch(a._inputPredictionNodes);
      break;
    }
    case BIDI_CSRL: {
      std::vector<deep::CSRL::LayerDesc> d(L);
      for (int l = 0; l < L; l++) {
        const bidi_layer_desc& s = layers[l];
        deep::CSRL::LayerDesc& t = d[l];
        t._width = s.width; t._height = s.height;
        t._receptiveRadius = s.r_ff; t._recurrentRadius = s.r_rec; t._lateralRadius = s.r_lat;
        t._predictiveRadius = s.r_pred; t._feedBackRadius = s.r_fb;
        t._numRecurrentInputs = s.csrl_n_recurrent;
        t._learnFeedForward = s.learn_ff; t._learnRecurrent = s.learn_rec; t._learnLateral = s.learn_lat;
        t._learnFeedBackRL = s.learn_fb; t._learnPredictionRL = s.learn_pred;
        t._sdrIterSettle = s.iter_settle; t._sdrIterMeasure = s.iter_measure; t._sdrLeak = s.leak;
        t._sdrLambda = s.lambda; t._sdrWeightDecay = s.weight_decay; t._sparsity = s.sparsity; t._sdrLearnThreshold = s.learn_threshold;
        t._averageSurpriseDecay = s.avg_surprise_decay; t._surpriseLearnFactor = s.csrl_surprise_factor;
        t._cellsPerColumn = s.col_cells; t._cellSparsity = s.col_sparsity; t._gamma = s.col_gamma; t._gammaLambda = s.col_gamma_lambda;
        t._gateFeedForwardAlpha = s.col_alpha_ff; t._gateLateralAlpha = s.col_alpha_lat; t._gateThresholdAlpha = s.col_alpha_threshold;
        t._qAlpha = s.col_alpha_q; t._actionAlpha = s.col_alpha_action; t._actionDeriveAlpha = s.csrl_derive_alpha;
        t._actionDeriveIterations = s.csrl_derive_iterations;
        t._explorationStdDev = s.col_expl_stddev; t._explorationBreak = s.col_expl_break;
      }
      deep::CSRL& c = r->cs;
      c._learnFeedBackRL = cfg->learn_infb;
      c._averageSurpriseDecay = cfg->csrl_avg_surprise_decay; c._surpriseLearnFactor = cfg->csrl_surprise_factor;
      c._numRecurrentInputs = cfg->csrl_n_recurrent;
      c._cellsPerColumn = cfg->col_cells; c._cellSparsity = cfg->col_sparsity; c._gamma = cfg->col_gamma; c._gammaLambda = cfg->col_gamma_lambda;
      c._gateFeedForwardAlpha = cfg->col_alpha_ff; c._gateLateralAlpha = cfg->col_alpha_lat; c._gateThresholdAlpha = cfg->col_alpha_threshold;
      c._qAlpha = cfg->col_alpha_q; c._actionAlpha = cfg->col_alpha_action; c._actionDeriveAlpha = cfg->csrl_derive_alpha;
      c._actionDeriveIterations = cfg->csrl_derive_iterations;
      c._explorationStdDev = cfg->col_expl_stddev; c._explorationBreak = cfg->col_expl_break;
      c._sdrIterSettle = cfg->csrl_iter_settle; c._sdrIterMeasure = cfg->csrl_iter_measure; c._sdrLeak = cfg->csrl_leak;
      std::vector<deep::CSRL::InputType> types((size_t)cfg->in_width * cfg->in_height, deep::CSRL::_state);
      for (int k = 0; k < cfg->q_n_actions; k++) types[(size_t)cfg->q_action_idx[k]] = deep::CSRL::_action;
      c.createRandom(cfg->in_width, cfg->in_height, cfg->r_infb, types, d, cfg->init_min_w, cfg->init_max_w,
                     cfg->init_min_inh, cfg->init_max_inh, cfg->init_threshold, r->gen);
      for (auto& ly : c._layers) zero_sdrrl_scratch(ly._predictionNodes);
      zero_sdrrl_scratch(c._inputPredictionNodes);
      break;
    }
    default: delete r; return nullptr;
  }
  return r;
}

void ref_destroy(void* h) { delete static_cast<Ref*>(h); }

long long ref_array_len(void* h, int which, int layer) {
  long long n = 0;
  if (!visit(*static_cast<Ref*>(h), which, layer, [&](float&) { n++; })) return 0;
  return n;
}
int ref_get_array(void* h, int which, int layer, float* dst) {
  return visit(*static_cast<Ref*>(h), which, layer, [&](float& v) { *dst++ = v; }) ? 0 : -1;
}
int ref_set_array(void* h, int which, int layer, const float* src) {
  return visit(*static_cast<Ref*>(h), which, layer, [&](float& v) { v = *src++; }) ? 0 : -1;
}

void ref_set_input(void* h, int i, float v) {
  Ref& r = *static_cast<Ref*>(h);
  switch (r.cfg.variant) {
    case BIDI_KWINNERS: r.kw.setInput(i, v); break;
    case BIDI_ITERATIVE: r.it.setInput(i, v); break;
    case BIDI_QPRSDR: r.qp.setState(i, v); break;
    case BIDI_NEO_HIERARCHY: r.nh.setInput(i, v); break;
    case BIDI_NEO_AGENT: r.ag.setInput(i, v); break;
    case BIDI_CSRL: r.cs._layers.front()._sdr.setVisibleState(i, v); break;  // setInput asserts the cell is of type _state; the harness writes any cell
  }
}
void ref_set_inputs(void* h, const float* in, int n) {
  for (int i = 0; i < n; i++) ref_set_input(h, i, in[i]);
}
float ref_get_prediction(void* h, int i) {
  Ref& r = *static_cast<Ref*>(h);
  switch (r.cfg.variant) {
    case BIDI_KWINNERS: return r.kw.getPrediction(i);
    case BIDI_ITERATIVE: return r.it.getPrediction(i);
    case BIDI_QPRSDR: return r.qp._prsdr.getPrediction(i);
    case BIDI_NEO_HIERARCHY: return r.nh.getPrediction(i);
    case BIDI_NEO_AGENT: return r.ag.getPrediction(i);
    case BIDI_CSRL: return r.cs.getPrediction(i);
  }
  return 0.0f;
}
void ref_get_predictions(void* h, float* out, int n) {
  for (int i = 0; i < n; i++) out[i] = ref_get_prediction(h, i);
}

void ref_step(void* h, float reward, int learn) {
  Ref& r = *static_cast<Ref*>(h);
  switch (r.cfg.variant) {
    case BIDI_KWINNERS: r.kw.simStep(learn != 0); break;
    case BIDI_ITERATIVE: r.it.simStep(r.gen, learn != 0); break;
    case BIDI_QPRSDR: { QuietCout quiet; r.qp.simStep(reward, r.gen, learn != 0); break; }
    case BIDI_NEO_HIERARCHY: r.nh.simStep(r.gen, learn != 0); break;
    case BIDI_NEO_AGENT: r.ag.simStep(reward, r.gen, learn != 0); break;
    case BIDI_CSRL: r.cs.simStep(reward, r.gen, learn != 0); break;
  }
}

// n steps back to back on one frame sequence frames[t % n_frames][n_in]; returns
// nothing -- used by the timing legs so the Python call overhead stays outside.
void ref_run(void* h, const float* frames, int n_frames, int n_in, const float* rewards, int steps, int learn) {
  for (int t = 0; t < steps; t++) {
    ref_set_inputs(h, frames + (size_t)(t % n_frames) * n_in, n_in);
    ref_step(h, rewards ? rewards[t % n_frames] : 0.0f, learn);
  }
}

// The RunnerMain loop, headless (RunnerMain.cpp:235-251): per step the open-loop
// part of the input comes from frames[t % n_frames]; inputs dst[k] are the fed-back
// actions clamp01(getPrediction(src[k])*0.5+0.5+N(0,0.05)) with the noise drawn
// from the agent's generator through a per-step normal_distribution, as the main
// does; reward = mean_k clamp01(getPrediction(src[k])*0.5+0.5) - 0.5 stands in for
// the Box2D body velocity.
void ref_run_runner(void* h, const float* frames, int n_frames, int n_in, int steps, int n_fb, const int* src,
                    const int* dst, int learn) {
  Ref& r = *static_cast<Ref*>(h);
  for (int t = 0; t < steps; t++) {
    ref_set_inputs(h, frames + (size_t)(t % n_frames) * n_in, n_in);
    std::normal_distribution<float> noiseDist(0.0f, 0.05f);
    float sum = 0.0f;
    for (int k = 0; k < n_fb; k++) {
      const float a = ref_get_prediction(h, src[k]) * 0.5f + 0.5f;
      sum += std::min(1.0f, std::max(0.0f, a));
      ref_set_input(h, dst[k], std::min(1.0f, std::max(0.0f, a + noiseDist(r.gen))));
    }
    ref_step(h, n_fb ? sum / n_fb - 0.5f : 0.0f, learn);
  }
}

int ref_num_noise(void* h);
int ref_num_columns(void* h) {
  Ref& r = *static_cast<Ref*>(h);
  if (r.cfg.variant == BIDI_CSRL) {
    int n = (int)r.cs._inputPredictionNodes.size();
    for (auto& ly : r.cs._layers) n += (int)ly._predictionNodes.size();
    return n;
  }
  if (r.cfg.variant != BIDI_NEO_AGENT) return 0;
  int n = (int)r.ag._inputPredictionNodes.size();
  for (auto& ly : r.ag._layers) n += (int)ly._predictionNodes.size();
  return n;
}

// The exploration draws the NEXT simStep will make (neo/Column.cpp:126-131), in
// column visit order (neo/Agent.cpp:155-156,212), obtained by replaying a COPY of
// the generator.  The draws do not depend on the network state, so the copy
// yields exactly what Column::simStep will see.  u,v: [n_columns][3].
int ref_peek_noise(void* h, float* u, float* v) {
  Ref& r = *static_cast<Ref*>(h);
  if (r.cfg.variant == BIDI_QPRSDR) {  // sdr/QPRSDR.cpp:202-213; the hierarchy's simStep draws nothing
    std::mt19937 g = r.gen;
    std::uniform_real_distribution<float> dist01(0.0f, 1.0f);
    std::normal_distribution<float> pertDist(0.0f, r.cfg.q_expl_stddev);
    for (int i = 0; i < r.cfg.q_n_actions; i++) {
      u[i] = dist01(g);
      v[i] = (u[i] < r.cfg.q_expl_break) ? dist01(g) : pertDist(g);
    }
    return 0;
  }
  if (r.cfg.variant == BIDI_CSRL) {  // deep/CSRL.cpp:205,232,246-253 then the SDRRL units in visit order, deep/SDRRL.cpp:53-54,203-208
    std::mt19937 g = r.gen;
    size_t k = 0;
    std::uniform_real_distribution<float> dist01(0.0f, 1.0f);
    std::normal_distribution<float> shared(0.0f, r.cfg.col_expl_stddev);
    for (size_t pi = 0; pi < r.cs._inputPredictionNodes.size(); pi++)
      if (r.cs._inputTypes[pi] == deep::CSRL::_action) {
        u[k] = dist01(g);
        v[k] = (u[k] < r.cfg.col_expl_break) ? dist01(g) : shared(g);
        k++;
      }
    auto unit = [&](int na, float stddev, float brk) {
      std::uniform_real_distribution<float> d01(0.0f, 1.0f);
      std::normal_distribution<float> pert(0.0f, stddev);
      for (int a = 0; a < na; a++, k++) {
        u[k] = d01(g);
        v[k] = (u[k] < brk) ? d01(g) : pert(g);
      }
    };
    for (int l = r.cfg.n_layers - 1; l >= 0; l--)
      for (auto& p : r.cs._layers[l]._predictionNodes) unit(p._sdrrl.getNumActions(), r.ld[l].col_expl_stddev, r.ld[l].col_expl_break);
    for (auto& p : r.cs._inputPredictionNodes) unit(p._sdrrl.getNumActions(), r.cfg.col_expl_stddev, r.cfg.col_expl_break);
    return 0;
  }
  if (r.cfg.variant != BIDI_NEO_AGENT) return -1;
  std::mt19937 g = r.gen;
  const int L = r.cfg.n_layers;
  size_t k = 0;
  auto visit_col = [&](float stddev, float brk) {
    std::uniform_real_distribution<float> dist01(0.0f, 1.0f);
    std::normal_distribution<float> pertDist(0.0f, stddev);
    for (int a = 0; a < BIDI_COLUMN_ACTIONS; a++, k++) {
      u[k] = dist01(g);
      v[k] = (u[k] < brk) ? dist01(g) : pertDist(g);
    }
  };
  for (int l = L - 1; l >= 0; l--)
    for (size_t pi = 0; pi < r.ag._layers[l]._predictionNodes.size(); pi++)
      visit_col(r.ld[l].col_expl_stddev, r.ld[l].col_expl_break);
  for (size_t pi = 0; pi < r.ag._inputPredictionNodes.size(); pi++)
    visit_col(r.cfg.col_expl_stddev, r.cfg.col_expl_break);
  return 0;
}

// neo::PredictiveHierarchy::simStepGenerate (neo/PredictiveHierarchy.cpp:247-315): the N(0,1) values its
// coders will draw (neo/SparseCoder.cpp:179,200: one normal_distribution per activateNoise call, iteration
// outer, hidden unit inner), from a COPY of the generator; and the call itself.
int ref_num_generate_noise(void* h) {
  Ref& r = *static_cast<Ref*>(h);
  if (r.cfg.variant != BIDI_NEO_HIERARCHY) return 0;
  int n = 0;
  for (int l = 0; l < r.cfg.n_layers; l++) n += r.ld[l].iter_settle * r.ld[l].width * r.ld[l].height;
  return n;
}
int ref_peek_generate(void* h, float* out) {
  Ref& r = *static_cast<Ref*>(h);
  if (r.cfg.variant != BIDI_NEO_HIERARCHY) return -1;
  std::mt19937 g = r.gen;
  size_t k = 0;
  for (int l = 0; l < r.cfg.n_layers; l++) {
    std::normal_distribution<float> noiseDist(0.0f, 1.0f);
    const int n = r.ld[l].iter_settle * r.ld[l].width * r.ld[l].height;
    for (int j = 0; j < n; j++) out[k++] = noiseDist(g);
  }
  return 0;
}
void ref_step_generate(void* h, float noise) {
  Ref& r = *static_cast<Ref*>(h);
  if (r.cfg.variant == BIDI_NEO_HIERARCHY) r.nh.simStepGenerate(r.gen, noise);
}

int ref_num_noise(void* h) {
  Ref& r = *static_cast<Ref*>(h);
  if (r.cfg.variant == BIDI_CSRL) {
    int n = r.cfg.q_n_actions;
    for (auto& ly : r.cs._layers) for (auto& p : ly._predictionNodes) n += p._sdrrl.getNumActions();
    for (auto& p : r.cs._inputPredictionNodes) n += p._sdrrl.getNumActions();
    return n;
  }
  return r.cfg.variant == BIDI_QPRSDR ? r.cfg.q_n_actions : ref_num_columns(h) * BIDI_COLUMN_ACTIONS;
}

// Per-unit list lengths, as bidi_conn_counts.
int ref_conn_counts(void* h, int which, int layer, int* counts) {
  Ref& r = *static_cast<Ref*>(h);
  const int L = r.cfg.n_layers;
  int k = 0;
#define COUNT_CODER(C)                                                                              \
  if (which == BIDI_FF_W) { for (auto& n : (C)._hidden) counts[k++] = (int)n._feedForwardConnections.size(); return 0; } \
  if (which == BIDI_REC_W) { for (auto& n : (C)._hidden) counts[k++] = (int)n._recurrentConnections.size(); return 0; }  \
  if (which == BIDI_LAT_W) { for (auto& n : (C)._hidden) counts[k++] = (int)n._lateralConnections.size(); return 0; }
#define COUNT_NODES(N)                                                                              \
  if (which == BIDI_FB_W) { for (auto& p : (N)) counts[k++] = (int)p._feedBackConnections.size(); return 0; }
#define COUNT_PRED(N)                                                                               \
  if (which == BIDI_PRED_W) { for (auto& p : (N)) counts[k++] = (int)p._predictiveConnections.size(); return 0; }
  switch (r.cfg.variant) {
    case BIDI_KWINNERS:
      if (layer >= L) return -1;
      COUNT_CODER(const_cast<sdr::RSDR&>(r.kw._layers[layer]._sdr))
      COUNT_NODES(r.kw._layers[layer]._predictionNodes) COUNT_PRED(r.kw._layers[layer]._predictionNodes)
      break;
    case BIDI_QPRSDR:
      if (which == BIDI_Q_FF_W) {
        if (layer >= L) return -1;
        for (auto& n : r.qp._qFunctionLayers[layer]._qFunctionNodes) counts[k++] = (int)n._feedForwardConnections.size();
        return 0;
      }
      /* fall through */
    case BIDI_ITERATIVE: {
      sdr::IPredictiveRSDR& it = hierarchy_of(r);
      if (layer == L) { COUNT_NODES(it._inputPredictionNodes) break; }
      COUNT_CODER(it._layers[layer]._sdr)
      COUNT_NODES(it._layers[layer]._predictionNodes) COUNT_PRED(it._layers[layer]._predictionNodes)
      break;
    }
    case BIDI_NEO_HIERARCHY:
      if (layer == L) { COUNT_NODES(r.nh._inputPredictionNodes) break; }
      COUNT_CODER(r.nh._layers[layer]._sdr)
      COUNT_NODES(r.nh._layers[layer]._predictionNodes) COUNT_PRED(r.nh._layers[layer]._predictionNodes)
      break;
    case BIDI_CSRL:
      if (layer == L) {
        COUNT_NODES(r.cs._inputPredictionNodes)
        if (which == BIDI_COL_INPUT) { for (auto& p : r.cs._inputPredictionNodes) counts[k++] = p._sdrrl.getNumStates(); return 0; }
        break;
      }
      COUNT_CODER(r.cs._layers[layer]._sdr)
      COUNT_NODES(r.cs._layers[layer]._predictionNodes) COUNT_PRED(r.cs._layers[layer]._predictionNodes)
      if (which == BIDI_COL_INPUT) { for (auto& p : r.cs._layers[layer]._predictionNodes) counts[k++] = p._sdrrl.getNumStates(); return 0; }
      break;
    case BIDI_NEO_AGENT:
      if (layer == L) {
        COUNT_NODES(r.ag._inputPredictionNodes)
        if (which == BIDI_COL_INPUT) { for (auto& p : r.ag._inputPredictionNodes) counts[k++] = p._column.getNumStates(); return 0; }
        break;
      }
      COUNT_CODER(r.ag._layers[layer]._sdr)
      COUNT_NODES(r.ag._layers[layer]._predictionNodes) COUNT_PRED(r.ag._layers[layer]._predictionNodes)
      if (which == BIDI_COL_INPUT) { for (auto& p : r.ag._layers[layer]._predictionNodes) counts[k++] = p._column.getNumStates(); return 0; }
      break;
  }
  return -1;
}

} // extern "C"
