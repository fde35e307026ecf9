// Drop-in for the reference's sdr/IPredictiveRSDR.h (iterative spiking coder
// hierarchy; Experiment_Prediction.cpp:113-148, VideoTest.cpp:52-139).  Same
// LayerDesc fields and defaults (sdr/IPredictiveRSDR.h:14-46) and signatures (:107-125).
#pragma once
#include "../detail/Hierarchy.h"

namespace sdr {
class IPredictiveRSDR : public bidinet::detail::Hierarchy {
public:
  struct LayerDesc {
    int _width, _height;
    int _receptiveRadius, _recurrentRadius, _lateralRadius, _predictiveRadius, _feedBackRadius;
    float _learnFeedForward, _learnRecurrent, _learnLateral;
    float _learnFeedBack, _learnPrediction;
    int _sdrIterSettle, _sdrIterMeasure;
    float _sdrLeak;
    float _sdrLambda;
    float _sdrHiddenDecay;
    float _sdrWeightDecay;
    float _sdrMaxWeightDelta;
    float _sdrSparsity;
    float _sdrLearnThreshold;
    float _sdrNoise;
    float _sdrBaselineDecay;
    float _sdrSensitivity;
    LayerDesc()
        : _width(16), _height(16),
          _receptiveRadius(3), _recurrentRadius(3), _lateralRadius(3), _predictiveRadius(3), _feedBackRadius(3),
          _learnFeedForward(0.01f), _learnRecurrent(0.01f), _learnLateral(0.2f),
          _learnFeedBack(0.05f), _learnPrediction(0.05f),
          _sdrIterSettle(30), _sdrIterMeasure(5),
          _sdrLeak(0.3f), _sdrLambda(0.95f), _sdrHiddenDecay(0.01f), _sdrWeightDecay(0.0f), _sdrMaxWeightDelta(0.5f),
          _sdrSparsity(0.2f), _sdrLearnThreshold(0.02f), _sdrNoise(0.01f),
          _sdrBaselineDecay(0.01f), _sdrSensitivity(8.0f) {}
  };

  // LayerDesc -> the C ABI's layer descriptor
  static void fill(bidi_layer_desc& d, const LayerDesc& s) {
    d.width = s._width; d.height = s._height;
    d.r_ff = s._receptiveRadius; d.r_rec = s._recurrentRadius; d.r_lat = s._lateralRadius;
    d.r_pred = s._predictiveRadius; d.r_fb = s._feedBackRadius;
    d.learn_ff = s._learnFeedForward; d.learn_rec = s._learnRecurrent; d.learn_lat = s._learnLateral;
    d.learn_fb = s._learnFeedBack; d.learn_pred = s._learnPrediction;
    d.iter_settle = s._sdrIterSettle; d.iter_measure = s._sdrIterMeasure; d.leak = s._sdrLeak; d.lambda = s._sdrLambda;
    d.weight_decay = s._sdrWeightDecay; d.max_weight_delta = s._sdrMaxWeightDelta; d.sparsity = s._sdrSparsity;
    d.learn_threshold = s._sdrLearnThreshold; d.baseline_decay = s._sdrBaselineDecay; d.sensitivity = s._sdrSensitivity;
  }

  float _learnInputFeedBack;
  IPredictiveRSDR() : _learnInputFeedBack(0.05f) {}

  void createRandom(int inputWidth, int inputHeight, int inputFeedBackRadius, const std::vector<LayerDesc>& layerDescs,
                    float initMinWeight, float initMaxWeight, float initMinInhibition, float initMaxInhibition,
                    float initThreshold, std::mt19937& generator) {
    _layerDescs = layerDescs;
    _cfg = bidi_config{};
    _cfg.variant = BIDI_ITERATIVE;
    _cfg.in_width = inputWidth; _cfg.in_height = inputHeight; _cfg.r_infb = inputFeedBackRadius;
    _cfg.learn_infb = _learnInputFeedBack;
    _cfg.init_min_w = initMinWeight; _cfg.init_max_w = initMaxWeight;
    _cfg.init_min_inh = initMinInhibition; _cfg.init_max_inh = initMaxInhibition; _cfg.init_threshold = initThreshold;
    _descs.assign(layerDescs.size(), bidi_layer_desc{});
    for (size_t l = 0; l < layerDescs.size(); l++) fill(_descs[l], layerDescs[l]);
    create();
    drawWeights(generator);
  }

  // The reference constructs a normal_distribution from the generator's argument and
  // never samples it (sdr/IRSDR.cpp:126): the generator is left untouched here too.
  void simStep(std::mt19937& /*generator*/, bool learn = true) { step(nullptr, nullptr, nullptr, learn); }

  void setInput(int index, float value) { _inputs[(size_t)index] = value; }
  // The reference indexes with the first HIDDEN layer's width (sdr/IPredictiveRSDR.h:115-117); kept as is.
  void setInput(int x, int y, float value) { setInput(x + y * _layerDescs.front()._width, value); }
  float getPrediction(int index) const { return prediction(index); }
  float getPrediction(int x, int y) const { return getPrediction(x + y * _cfg.in_width); }
  const std::vector<LayerDesc>& getLayerDescs() const { return _layerDescs; }

private:
  std::vector<LayerDesc> _layerDescs;
};
}  // namespace sdr
