// Drop-in for the reference's neo/PredictiveHierarchy.h.  Same LayerDesc fields and
// defaults (neo/PredictiveHierarchy.h:14-45) and signatures (:106-130).
// simStepGenerate (noise-driven dreaming, neo/PredictiveHierarchy.h:110) included.
#pragma once
#include "../detail/Hierarchy.h"

namespace neo {
class PredictiveHierarchy : public bidinet::detail::Hierarchy {
public:
  struct LayerDesc {
    int _width, _height;
    int _receptiveRadius, _recurrentRadius, _lateralRadius, _predictiveRadius, _feedBackRadius;
    float _learnFeedForward, _learnRecurrent, _learnLateral;
    float _learnFeedBack, _learnPrediction;
    int _sdrIter;
    float _sdrLeak;
    float _sdrLambda;
    float _sdrHiddenDecay;
    float _sdrWeightDecay;
    float _sdrMaxWeightDelta;
    float _sdrSparsity;
    float _sdrLearnThreshold;
    float _sdrBaselineDecay;
    float _sdrSensitivity;
    LayerDesc()
        : _width(16), _height(16),
          _receptiveRadius(4), _recurrentRadius(4), _lateralRadius(4), _predictiveRadius(4), _feedBackRadius(4),
          _learnFeedForward(0.01f), _learnRecurrent(0.01f), _learnLateral(0.05f),
          _learnFeedBack(0.1f), _learnPrediction(0.03f),
          _sdrIter(30),
          _sdrLeak(0.1f), _sdrLambda(0.95f), _sdrHiddenDecay(0.01f), _sdrWeightDecay(0.0f), _sdrMaxWeightDelta(0.5f),
          _sdrSparsity(0.02f), _sdrLearnThreshold(0.01f),
          _sdrBaselineDecay(0.01f), _sdrSensitivity(6.0f) {}
  };

  float _learnInputFeedBack;
  PredictiveHierarchy() : _learnInputFeedBack(0.1f) {}

  void createRandom(int inputWidth, int inputHeight, int inputFeedBackRadius, const std::vector<LayerDesc>& layerDescs,
                    float initMinWeight, float initMaxWeight, float initMinInhibition, float initMaxInhibition,
                    float initThreshold, std::mt19937& generator) {
    _layerDescs = layerDescs;
    _cfg = bidi_config{};
    _cfg.variant = BIDI_NEO_HIERARCHY;
    _cfg.in_width = inputWidth; _cfg.in_height = inputHeight; _cfg.r_infb = inputFeedBackRadius;
    _cfg.learn_infb = _learnInputFeedBack;
    _cfg.init_min_w = initMinWeight; _cfg.init_max_w = initMaxWeight;
    _cfg.init_min_inh = initMinInhibition; _cfg.init_max_inh = initMaxInhibition; _cfg.init_threshold = initThreshold;
    _descs.assign(layerDescs.size(), bidi_layer_desc{});
    for (size_t l = 0; l < layerDescs.size(); l++) fill(_descs[l], layerDescs[l]);
    create();
    drawWeights(generator);
  }

  void simStep(std::mt19937& /*generator*/, bool learn = true) { step(nullptr, nullptr, nullptr, learn); }

  // neo/PredictiveHierarchy.cpp:247-315.  The N(0,1) values come from the CALLER's generator exactly as
  // SparseCoder::activateNoise draws them (neo/SparseCoder.cpp:179,200): layer 0 first, one
  // normal_distribution per layer, iteration outer, hidden unit inner.
  void simStepGenerate(std::mt19937& generator, float noise) {
    std::vector<float> normals;
    normals.reserve((size_t)bidi_num_generate_noise(_h));
    for (const LayerDesc& d : _layerDescs) {
      std::normal_distribution<float> noiseDist(0.0f, 1.0f);
      const int n = d._sdrIter * d._width * d._height;
      for (int k = 0; k < n; k++) normals.push_back(noiseDist(generator));
    }
    bidinet::detail::check(bidi_set_inputs(_h, _inputs.data()), _h, "bidi_set_inputs");
    bidinet::detail::check(bidi_step_generate(_h, normals.data(), noise), _h, "bidi_step_generate");
    _predictionsFresh = false;
    _layersFresh = false;
  }

  void setInput(int index, float value) { _inputs[(size_t)index] = value; }
  void setInput(int x, int y, float value) { setInput(x + y * _layerDescs.front()._width, value); }  // neo/PredictiveHierarchy.h:116-118
  float getPrediction(int index) const { return prediction(index); }
  float getPrediction(int x, int y) const { return getPrediction(x + y * _cfg.in_width); }
  const std::vector<LayerDesc>& getLayerDescs() const { return _layerDescs; }

  template <class Desc> static void fill(bidi_layer_desc& d, const Desc& s) {
    d.width = s._width; d.height = s._height;
    d.r_ff = s._receptiveRadius; d.r_rec = s._recurrentRadius; d.r_lat = s._lateralRadius;
    d.r_pred = s._predictiveRadius; d.r_fb = s._feedBackRadius;
    d.learn_ff = s._learnFeedForward; d.learn_rec = s._learnRecurrent; d.learn_lat = s._learnLateral;
    d.learn_fb = s._learnFeedBack; d.learn_pred = s._learnPrediction;
    d.iter_settle = s._sdrIter; d.iter_measure = 0; d.leak = s._sdrLeak; d.lambda = s._sdrLambda;
    d.weight_decay = s._sdrWeightDecay; d.max_weight_delta = s._sdrMaxWeightDelta; d.sparsity = s._sdrSparsity;
    d.learn_threshold = s._sdrLearnThreshold; d.baseline_decay = s._sdrBaselineDecay; d.sensitivity = s._sdrSensitivity;
  }

private:
  std::vector<LayerDesc> _layerDescs;
};
}  // namespace neo
