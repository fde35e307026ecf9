// Drop-in for the reference's sdr/PredictiveRSDR.h (k-winners hierarchy).  Same
// namespace, class name, LayerDesc fields and defaults (sdr/PredictiveRSDR.h:14-42)
// and method signatures (:80-106); the work happens in the CUDA engine.
#pragma once
#include "../detail/Hierarchy.h"

namespace sdr {
class PredictiveRSDR : public bidinet::detail::Hierarchy {
public:
  struct LayerDesc {
    int _width, _height;
    int _receptiveRadius, _recurrentRadius, _lateralRadius, _predictiveRadius, _feedBackRadius;
    float _learnFeedForward, _learnRecurrent, _learnLateral, _learnThreshold;
    float _learnFeedBack, _learnPrediction;
    int _subIterSettle;
    int _subIterMeasure;
    float _leak;
    float _averageSurpriseDecay;
    float _attentionFactor;
    float _sparsity;
    LayerDesc()
        : _width(16), _height(16),
          _receptiveRadius(8), _recurrentRadius(4), _lateralRadius(3), _predictiveRadius(4), _feedBackRadius(8),
          _learnFeedForward(0.02f), _learnRecurrent(0.02f), _learnLateral(0.2f), _learnThreshold(0.12f),
          _learnFeedBack(0.05f), _learnPrediction(0.05f),
          _subIterSettle(17), _subIterMeasure(5), _leak(0.1f),
          _averageSurpriseDecay(0.01f), _attentionFactor(4.0f), _sparsity(0.01f) {}
  };

  void createRandom(int inputWidth, int inputHeight, const std::vector<LayerDesc>& layerDescs, float initMinWeight,
                    float initMaxWeight, float initThreshold, std::mt19937& generator) {
    _layerDescs = layerDescs;
    _cfg = bidi_config{};
    _cfg.variant = BIDI_KWINNERS;
    _cfg.in_width = inputWidth; _cfg.in_height = inputHeight;
    _cfg.init_min_w = initMinWeight; _cfg.init_max_w = initMaxWeight; _cfg.init_threshold = initThreshold;
    _descs.assign(layerDescs.size(), bidi_layer_desc{});
    for (size_t l = 0; l < layerDescs.size(); l++) {
      const LayerDesc& s = layerDescs[l];
      bidi_layer_desc& d = _descs[l];
      d.width = s._width; d.height = s._height;
      d.r_ff = s._receptiveRadius; d.r_rec = s._recurrentRadius; d.r_lat = s._lateralRadius;
      d.r_pred = s._predictiveRadius; d.r_fb = s._feedBackRadius;
      d.learn_ff = s._learnFeedForward; d.learn_rec = s._learnRecurrent; d.learn_lat = s._learnLateral;
      d.learn_threshold = s._learnThreshold; d.learn_fb = s._learnFeedBack; d.learn_pred = s._learnPrediction;
      d.avg_surprise_decay = s._averageSurpriseDecay; d.attention_factor = s._attentionFactor; d.sparsity = s._sparsity;
    }
    create();
    drawWeights(generator);
  }

  void simStep(bool learn = true) { step(nullptr, nullptr, nullptr, learn); }

  void setInput(int index, float value) { _inputs[(size_t)index] = value; }
  void setInput(int x, int y, float value) { setInput(x + y * _cfg.in_width, value); }  // visible width, sdr/PredictiveRSDR.h:88-90
  float getPrediction(int index) const { return prediction(index); }
  float getPrediction(int x, int y) const { return getPrediction(x + y * _cfg.in_width); }
  const std::vector<LayerDesc>& getLayerDescs() const { return _layerDescs; }

private:
  std::vector<LayerDesc> _layerDescs;
};
}  // namespace sdr
