// bidinet/detail/Hierarchy.h -- shared plumbing of the drop-in C++ facade.
//
// The reference's hierarchy classes (sdr::PredictiveRSDR, sdr::IPredictiveRSDR,
// neo::PredictiveHierarchy, neo::Agent) are plain value classes an experiment main
// drives through createRandom / setInput / simStep / getPrediction
// (RunnerMain.cpp:139-154,239-249; Experiment_Prediction.cpp:124-148).  The facade
// classes keep those names and signatures and forward to the extern "C" engine
// (include/bidinet_b200.h).  Differences a caller can observe are listed in
// INTEGRATION.md (getLayers() returns a read-only host snapshot of the state, the weight lists come
// through per-array accessors; objects are move-only because they own device memory).
#pragma once

#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "bidinet_b200.h"

namespace bidinet {
namespace detail {

inline void check(int rc, bidi_engine* h, const char* what) {
  if (rc != BIDI_OK) throw std::runtime_error(std::string(what) + ": " + bidi_last_error(h));
}

// Host mirror of what the reference's callers read through getLayers() (sdr/IPredictiveRSDR.h:83-131,
// sdr/PredictiveRSDR.h:63-104, neo/PredictiveHierarchy.h:82-132, neo/Agent.h:106-170; read by sdr/QPRSDR.cpp:123,
// sdr/ICSRL.cpp:89, vis/CSRLVisualizer.cpp:58): per layer the coder's read accessors (`_sdr`) and the prediction nodes'
// scalars.  The state lives in device memory; the mirror is downloaded on the first getLayers() after a step.
struct CoderMirror {
  int _visibleWidth = 0, _visibleHeight = 0, _hiddenWidth = 0, _hiddenHeight = 0;
  std::vector<float> _visibleState, _visibleRecon, _hiddenState, _hiddenStatePrev;
  float getVisibleRecon(int index) const { return _visibleRecon[(size_t)index]; }
  float getVisibleRecon(int x, int y) const { return getVisibleRecon(x + y * _visibleWidth); }
  float getVisibleState(int index) const { return _visibleState[(size_t)index]; }
  float getVisibleState(int x, int y) const { return getVisibleState(x + y * _visibleWidth); }
  float getHiddenState(int index) const { return _hiddenState[(size_t)index]; }
  float getHiddenState(int x, int y) const { return getHiddenState(x + y * _hiddenWidth); }
  float getHiddenStatePrev(int index) const { return _hiddenStatePrev[(size_t)index]; }
  float getHiddenStatePrev(int x, int y) const { return getHiddenStatePrev(x + y * _hiddenWidth); }
  int getNumVisible() const { return _visibleWidth * _visibleHeight; }
  int getNumHidden() const { return _hiddenWidth * _hiddenHeight; }
  int getVisibleWidth() const { return _visibleWidth; }
  int getVisibleHeight() const { return _visibleHeight; }
  int getHiddenWidth() const { return _hiddenWidth; }
  int getHiddenHeight() const { return _hiddenHeight; }
};
struct PredictionNodeMirror {
  float _state = 0.0f, _statePrev = 0.0f, _activation = 0.0f, _activationPrev = 0.0f, _baseline = 0.0f;
};
struct LayerMirror {
  CoderMirror _sdr;
  std::vector<PredictionNodeMirror> _predictionNodes;
};

class Hierarchy {
public:
  typedef LayerMirror Layer;
  typedef PredictionNodeMirror PredictionNode;
  // getLayers() of the reference classes: a read-only snapshot of the current state (see LayerMirror)
  const std::vector<Layer>& getLayers() const {
    if (!_layersFresh) {
      const int L = _cfg.n_layers;
      _layers.resize((size_t)L);
      for (int l = 0; l < L; l++) {
        Layer& y = _layers[(size_t)l];
        y._sdr._visibleWidth = l == 0 ? _cfg.in_width : _descs[(size_t)l - 1].width;
        y._sdr._visibleHeight = l == 0 ? _cfg.in_height : _descs[(size_t)l - 1].height;
        y._sdr._hiddenWidth = _descs[(size_t)l].width; y._sdr._hiddenHeight = _descs[(size_t)l].height;
        y._sdr._visibleState = optArray(BIDI_VIS_INPUT, l); y._sdr._visibleRecon = optArray(BIDI_VIS_RECON, l);
        y._sdr._hiddenState = optArray(BIDI_HID_STATE, l); y._sdr._hiddenStatePrev = optArray(BIDI_HID_STATE_PREV, l);
        const std::vector<float> s = optArray(BIDI_P_STATE, l), sp = optArray(BIDI_P_STATE_PREV, l), a = optArray(BIDI_P_ACTIVATION, l),
                                 ap = optArray(BIDI_P_ACTIVATION_PREV, l), b = optArray(BIDI_P_BASELINE, l);
        y._predictionNodes.assign(s.size(), PredictionNode());
        for (size_t i = 0; i < s.size(); i++) {
          PredictionNode& n = y._predictionNodes[i];
          n._state = s[i];
          if (i < sp.size()) n._statePrev = sp[i];
          if (i < a.size()) n._activation = a[i];
          if (i < ap.size()) n._activationPrev = ap[i];
          if (i < b.size()) n._baseline = b[i];
        }
      }
      _layersFresh = true;
    }
    return _layers;
  }

  Hierarchy() = default;
  Hierarchy(const Hierarchy&) = delete;
  Hierarchy& operator=(const Hierarchy&) = delete;
  Hierarchy(Hierarchy&& o) noexcept { swap(o); }
  Hierarchy& operator=(Hierarchy&& o) noexcept { swap(o); return *this; }
  ~Hierarchy() { if (_h) bidi_destroy(_h); }

  // Serialised state in the reference's order (bidi_array in bidinet_b200.h): everything getLayers() does not
  // mirror (the connection lists' weights and traces, thresholds, the columns' arrays).
  std::vector<float> getArray(int which, int layer) const {
    const long long n = bidi_array_len(_h, which, layer);
    std::vector<float> out(n > 0 ? (size_t)n : 0);
    if (n > 0) check(bidi_get_array(_h, 0, which, layer, out.data(), n), _h, "bidi_get_array");
    return out;
  }
  bidi_engine* engine() const { return _h; }

protected:
  mutable std::vector<Layer> _layers;
  mutable bool _layersFresh = false;
  std::vector<float> optArray(int which, int layer) const {  // empty when the variant does not have the array
    const long long n = bidi_array_len(_h, which, layer);
    std::vector<float> out(n > 0 ? (size_t)n : 0);
    if (n > 0) check(bidi_get_array(_h, 0, which, layer, out.data(), n), _h, "bidi_get_array");
    return out;
  }
  bidi_engine* _h = nullptr;
  bidi_config _cfg{};
  std::vector<bidi_layer_desc> _descs;
  std::vector<float> _inputs;
  mutable std::vector<float> _predictions;
  mutable bool _predictionsFresh = false;

  void swap(Hierarchy& o) {
    std::swap(_h, o._h); std::swap(_cfg, o._cfg); _descs.swap(o._descs); _inputs.swap(o._inputs);
    _predictions.swap(o._predictions); std::swap(_predictionsFresh, o._predictionsFresh);
    _layers.swap(o._layers); std::swap(_layersFresh, o._layersFresh);
  }

  void create() {
    if (_h) { bidi_destroy(_h); _h = nullptr; }
    _cfg.n_layers = (int)_descs.size();
    bidi_engine* h = nullptr;
    const int rc = bidi_create(&_cfg, _descs.data(), 1, 0, &h);
    if (rc != BIDI_OK) throw std::runtime_error(std::string("bidi_create: ") + bidi_last_error(nullptr));
    _h = h;
    _inputs.assign((size_t)_cfg.in_width * _cfg.in_height, 0.0f);
    _predictions.assign(_inputs.size(), 0.0f);
    _predictionsFresh = true;
    _layersFresh = false;
  }

  std::vector<int> counts(int which, int layer, int units) const {
    std::vector<int> c((size_t)units);
    check(bidi_conn_counts(_h, which, layer, c.data(), units), _h, "bidi_conn_counts");
    return c;
  }
  void upload(int which, int layer, const std::vector<float>& v) {
    if (!v.empty()) check(bidi_set_array(_h, 0, which, layer, v.data(), (long long)v.size()), _h, "bidi_set_array");
  }

  // createRandom's weight draws, consumed from the CALLER's generator in the
  // reference's order, so the generator ends in the same state as
  // after the reference's createRandom.
  void drawWeights(std::mt19937& generator) {
    const int L = _cfg.n_layers, V = _cfg.variant;
    const bool iterative = V != BIDI_KWINNERS, agent = V == BIDI_NEO_AGENT;
    std::uniform_real_distribution<float> weightDist(_cfg.init_min_w, _cfg.init_max_w);
    std::uniform_real_distribution<float> inhibitionDist(_cfg.init_min_inh, _cfg.init_max_inh);
    for (int l = 0; l <= L; l++) {
      if (l == L && !iterative) break;
      const int units = l < L ? _descs[l].width * _descs[l].height : _cfg.in_width * _cfg.in_height;
      const int cells = l < L ? _descs[l].col_cells : _cfg.col_cells;
      if (l < L) {  // coder: per unit ff, recurrent, (lateral)  -- sdr/RSDR.cpp:48-116, sdr/IRSDR.cpp:50-121
        const std::vector<int> nff = counts(BIDI_FF_W, l, units), nrec = counts(BIDI_REC_W, l, units);
        const std::vector<int> nlat = iterative ? counts(BIDI_LAT_W, l, units) : std::vector<int>();
        std::vector<float> ff, rec, lat;
        for (int i = 0; i < units; i++) {
          for (int k = 0; k < nff[i]; k++) ff.push_back(weightDist(generator));
          for (int k = 0; k < nrec[i]; k++) rec.push_back(weightDist(generator));
          if (iterative) for (int k = 0; k < nlat[i]; k++) lat.push_back(inhibitionDist(generator));
        }
        upload(BIDI_FF_W, l, ff); upload(BIDI_REC_W, l, rec); upload(BIDI_LAT_W, l, lat);
      }
      // prediction nodes: bias (drawn, never used), feedback, predictive, then the column
      const bool hasFb = l == L || l < L - 1;
      const std::vector<int> nfb = hasFb ? counts(BIDI_FB_W, l, units) : std::vector<int>((size_t)units, 0);
      const std::vector<int> npred = l < L ? counts(BIDI_PRED_W, l, units) : std::vector<int>((size_t)units, 0);
      std::vector<float> fb, pred, cff, clat, cq, cact, cthr;
      for (int i = 0; i < units; i++) {
        (void)weightDist(generator);
        for (int k = 0; k < nfb[i]; k++) fb.push_back(weightDist(generator));
        for (int k = 0; k < npred[i]; k++) pred.push_back(weightDist(generator));
        if (agent) {  // neo/Column.cpp:22-43
          const int nin = npred[i] + 2 * nfb[i];
          for (int c = 0; c < cells; c++) {
            cthr.push_back(_cfg.init_threshold);
            for (int j = 0; j < nin; j++) cff.push_back(weightDist(generator));
            for (int j = 0; j < cells; j++) clat.push_back(inhibitionDist(generator));
            cq.push_back(weightDist(generator));
          }
          for (int a = 0; a < BIDI_COLUMN_ACTIONS; a++)
            for (int k = 0; k < cells; k++) cact.push_back(weightDist(generator));
        }
      }
      upload(BIDI_FB_W, l, fb); upload(BIDI_PRED_W, l, pred);
      if (agent) {
        upload(BIDI_COL_FF_W, l, cff); upload(BIDI_COL_LAT_W, l, clat); upload(BIDI_COL_Q_W, l, cq);
        upload(BIDI_COL_ACT_W, l, cact); upload(BIDI_COL_THRESHOLD, l, cthr);
      }
    }
  }

  void step(const float* reward, const float* u, const float* v, bool learn) {
    check(bidi_set_inputs(_h, _inputs.data()), _h, "bidi_set_inputs");
    check(bidi_step(_h, reward, u, v, learn ? 1 : 0), _h, "bidi_step");
    _predictionsFresh = false;
    _layersFresh = false;
  }
  float prediction(int index) const {
    if (!_predictionsFresh) {
      check(bidi_get_predictions(_h, _predictions.data()), _h, "bidi_get_predictions");
      _predictionsFresh = true;
    }
    return _predictions[(size_t)index];
  }
};

}  // namespace detail
}  // namespace bidinet
