// Drop-in for the reference's neo/Agent.h, the hierarchy RunnerMain.cpp:139-154,239-249
// drives.  Same LayerDesc fields and defaults (neo/Agent.h:19-65), public
// hyper-parameter members (:125-135) and signatures (:146-170).
//
// _columnGamma / _columnGammaLambda are never initialised by the reference (undefined
// behaviour there); this facade defaults them to 0.99 / 0.95.
#pragma once
#include "PredictiveHierarchy.h"

namespace neo {
class Agent : public bidinet::detail::Hierarchy {
public:
  enum ColumnAction { _attention = 0, _learnPrediction, _signal, _numColumnActions };

  struct LayerDesc {
    int _width, _height;
    int _cellsPerColumn;
    float _columnSparsity;
    int _columnIter;
    float _columnLeak;
    float _columnGamma;
    float _columnGammaLambda;
    float _columnFeedForwardAlpha, _columnLateralAlpha, _columnThresholdAlpha;
    float _columnQAlpha, _columnActionAlpha;
    float _columnExplorationStdDev, _columnExplorationBreakChance;
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
          _cellsPerColumn(16), _columnSparsity(0.125f), _columnIter(7),
          _columnLeak(0.1f), _columnGamma(0.99f), _columnGammaLambda(0.95f),
          _columnFeedForwardAlpha(0.01f), _columnLateralAlpha(0.05f), _columnThresholdAlpha(0.01f),
          _columnQAlpha(0.01f), _columnActionAlpha(0.1f),
          _columnExplorationStdDev(0.05f), _columnExplorationBreakChance(0.01f),
          _receptiveRadius(4), _recurrentRadius(4), _lateralRadius(4), _predictiveRadius(4), _feedBackRadius(4),
          _learnFeedForward(0.01f), _learnRecurrent(0.01f), _learnLateral(0.05f),
          _learnFeedBack(0.1f), _learnPrediction(0.03f),
          _sdrIter(30),
          _sdrLeak(0.1f), _sdrLambda(0.95f), _sdrHiddenDecay(0.01f), _sdrWeightDecay(0.0f), _sdrMaxWeightDelta(0.5f),
          _sdrSparsity(0.02f), _sdrLearnThreshold(0.01f),
          _sdrBaselineDecay(0.01f), _sdrSensitivity(6.0f) {}
  };

  // First layer (input node) columns
  int _cellsPerColumn;
  float _columnSparsity;
  int _columnIter;
  float _columnLeak;
  float _columnGamma;
  float _columnGammaLambda;
  float _columnFeedForwardAlpha, _columnLateralAlpha, _columnThresholdAlpha;
  float _columnQAlpha, _columnActionAlpha;
  float _columnExplorationStdDev, _columnExplorationBreakChance;
  float _learnInputFeedBack;

  Agent()
      : _cellsPerColumn(16), _columnSparsity(0.125f), _columnIter(7),
        _columnLeak(0.1f), _columnGamma(0.99f), _columnGammaLambda(0.95f),
        _columnFeedForwardAlpha(0.01f), _columnLateralAlpha(0.05f), _columnThresholdAlpha(0.01f),
        _columnQAlpha(0.01f), _columnActionAlpha(0.1f),
        _columnExplorationStdDev(0.05f), _columnExplorationBreakChance(0.01f),
        _learnInputFeedBack(0.1f) {}

  void createRandom(int inputWidth, int inputHeight, int inputFeedBackRadius, const std::vector<LayerDesc>& layerDescs,
                    float initMinWeight, float initMaxWeight, float initMinInhibition, float initMaxInhibition,
                    float initThreshold, std::mt19937& generator) {
    describe(inputWidth, inputHeight, inputFeedBackRadius, layerDescs, initMinWeight, initMaxWeight, initMinInhibition,
             initMaxInhibition, initThreshold);
    create();
    drawWeights(generator);
    _columns = bidi_num_columns(_h);
    _noiseU.assign((size_t)_columns * 3, 0.0f);
    _noiseV.assign((size_t)_columns * 3, 0.0f);
  }

  // The engine-side description (bidi_config + bidi_layer_desc list) of this agent's hyper-parameters and of
  // createRandom's arguments, without creating anything: neo::AgentBatch builds many agents from one description.
  void describe(int inputWidth, int inputHeight, int inputFeedBackRadius, const std::vector<LayerDesc>& layerDescs,
                float initMinWeight, float initMaxWeight, float initMinInhibition, float initMaxInhibition, float initThreshold) {
    _layerDescs = layerDescs;
    _cfg = bidi_config{};
    _cfg.variant = BIDI_NEO_AGENT;
    _cfg.in_width = inputWidth; _cfg.in_height = inputHeight; _cfg.r_infb = inputFeedBackRadius;
    _cfg.learn_infb = _learnInputFeedBack;
    _cfg.col_cells = _cellsPerColumn; _cfg.col_iter = _columnIter; _cfg.col_sparsity = _columnSparsity;
    _cfg.col_leak = _columnLeak; _cfg.col_gamma = _columnGamma; _cfg.col_gamma_lambda = _columnGammaLambda;
    _cfg.col_alpha_ff = _columnFeedForwardAlpha; _cfg.col_alpha_lat = _columnLateralAlpha;
    _cfg.col_alpha_threshold = _columnThresholdAlpha; _cfg.col_alpha_q = _columnQAlpha; _cfg.col_alpha_action = _columnActionAlpha;
    _cfg.col_expl_stddev = _columnExplorationStdDev; _cfg.col_expl_break = _columnExplorationBreakChance;
    _cfg.init_min_w = initMinWeight; _cfg.init_max_w = initMaxWeight;
    _cfg.init_min_inh = initMinInhibition; _cfg.init_max_inh = initMaxInhibition; _cfg.init_threshold = initThreshold;
    _descs.assign(layerDescs.size(), bidi_layer_desc{});
    for (size_t l = 0; l < layerDescs.size(); l++) {
      const LayerDesc& s = layerDescs[l];
      bidi_layer_desc& d = _descs[l];
      PredictiveHierarchy::fill(d, s);
      d.col_cells = s._cellsPerColumn; d.col_iter = s._columnIter; d.col_sparsity = s._columnSparsity; d.col_leak = s._columnLeak;
      d.col_gamma = s._columnGamma; d.col_gamma_lambda = s._columnGammaLambda;
      d.col_alpha_ff = s._columnFeedForwardAlpha; d.col_alpha_lat = s._columnLateralAlpha; d.col_alpha_threshold = s._columnThresholdAlpha;
      d.col_alpha_q = s._columnQAlpha; d.col_alpha_action = s._columnActionAlpha;
      d.col_expl_stddev = s._columnExplorationStdDev; d.col_expl_break = s._columnExplorationBreakChance;
    }
    _cfg.n_layers = (int)_descs.size();
  }
  const bidi_config& config() const { return _cfg; }
  const std::vector<bidi_layer_desc>& descs() const { return _descs; }

  // The exploration samples are drawn from the CALLER's generator exactly as
  // neo::Column::simStep draws them (neo/Column.cpp:47-48,126-131: per column visit a
  // fresh normal_distribution; per action dist01, then dist01 or the perturbation),
  // in the reference's visit order: layers top -> bottom, nodes ascending, then the
  // input nodes (neo/Agent.cpp:155-156,212).  The draws do not depend on the network
  // state, so they can all be taken before the step is launched.
  void simStep(float reward, std::mt19937& generator, bool learn = true) {
    drawExploration(generator, _noiseU.data(), _noiseV.data());
    step(&reward, _noiseU.data(), _noiseV.data(), learn);
  }
  // u, v: [columns][3] in the reference's column visit order
  void drawExploration(std::mt19937& generator, float* u, float* v) const {
    size_t k = 0;
    auto visit = [&](int nodes, float stddev, float breakChance) {
      for (int n = 0; n < nodes; n++) {
        std::uniform_real_distribution<float> dist01(0.0f, 1.0f);
        std::normal_distribution<float> pertDist(0.0f, stddev);
        for (int a = 0; a < _numColumnActions; a++, k++) {
          u[k] = dist01(generator);
          v[k] = u[k] < breakChance ? dist01(generator) : pertDist(generator);
        }
      }
    };
    for (int l = (int)_layerDescs.size() - 1; l >= 0; l--)
      visit(_layerDescs[l]._width * _layerDescs[l]._height, _layerDescs[l]._columnExplorationStdDev,
            _layerDescs[l]._columnExplorationBreakChance);
    visit(_cfg.in_width * _cfg.in_height, _columnExplorationStdDev, _columnExplorationBreakChance);
  }

  void setInput(int index, float value) { _inputs[(size_t)index] = value; }
  void setInput(int x, int y, float value) { setInput(x + y * _layerDescs.front()._width, value); }  // neo/Agent.h:154-156
  float getPrediction(int index) const { return prediction(index); }
  float getPrediction(int x, int y) const { return getPrediction(x + y * _cfg.in_width); }
  const std::vector<LayerDesc>& getLayerDescs() const { return _layerDescs; }

private:
  std::vector<LayerDesc> _layerDescs;
  int _columns = 0;
  std::vector<float> _noiseU, _noiseV;
};
}  // namespace neo
