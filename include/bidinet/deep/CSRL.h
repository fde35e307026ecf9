// Drop-in for the reference's deep/CSRL.h (continuous-action RL hierarchy: sdr::IRSDR coders bottom-up, clamped
// predictions top-down, one deep::SDRRL unit per prediction node; deep/CSRL.cpp:11-449, deep/SDRRL.cpp:9-277).
// Same LayerDesc fields and defaults (deep/CSRL.h:29-95), public members (:183-243) and signatures (:245-267),
// forwarding to the engine's BIDI_CSRL variant through the C ABI (include/bidinet_b200.h).
#pragma once
#include <cassert>
#include <sstream>

#include "../detail/Hierarchy.h"

namespace deep {
class CSRL : public bidinet::detail::Hierarchy {
public:
  enum InputType { _state, _action };
  enum ActionType { _attention = 0, _learn, _reward, _numActionTypes };

  struct LayerDesc {
    int _width, _height;
    int _receptiveRadius, _recurrentRadius, _lateralRadius, _predictiveRadius, _feedBackRadius;
    int _numRecurrentInputs;
    float _learnFeedForward, _learnRecurrent, _learnLateral;
    float _learnFeedBackPred, _learnPredictionPred;
    float _learnFeedBackRL, _learnPredictionRL;
    float _drift;
    int _sdrIterSettle;
    int _sdrIterMeasure;
    float _sdrLeak;
    float _sdrStepSize;
    float _sdrLambda;
    float _sdrHiddenDecay;
    float _sdrWeightDecay;
    float _sparsity;
    float _sdrLearnThreshold;
    float _sdrNoise;
    float _sdrBaselineDecay;
    float _sdrSensitivity;
    float _averageSurpriseDecay;
    float _surpriseLearnFactor;
    int _cellsPerColumn;
    float _cellSparsity;
    float _gamma;
    float _gammaLambda;
    float _gateFeedForwardAlpha;
    float _gateLateralAlpha;
    float _gateThresholdAlpha;
    int _gateSolveIter;
    float _qAlpha;
    float _actionAlpha, _actionDeriveAlpha;
    int _actionDeriveIterations;
    float _explorationStdDev;
    float _explorationBreak;
    float _epsilon;
    LayerDesc()
        : _width(16), _height(16),
          _receptiveRadius(5), _recurrentRadius(4), _lateralRadius(4), _predictiveRadius(4), _feedBackRadius(5),
          _numRecurrentInputs(4),
          _learnFeedForward(0.01f), _learnRecurrent(0.01f), _learnLateral(0.2f),
          _learnFeedBackPred(0.05f), _learnPredictionPred(0.05f),
          _learnFeedBackRL(0.05f), _learnPredictionRL(0.05f),
          _drift(0.0f),
          _sdrIterSettle(17), _sdrIterMeasure(4), _sdrLeak(0.1f),
          _sdrStepSize(0.04f), _sdrLambda(0.95f), _sdrHiddenDecay(0.01f), _sdrWeightDecay(0.0001f),
          _sparsity(0.08f), _sdrLearnThreshold(0.01f), _sdrNoise(0.01f),
          _sdrBaselineDecay(0.01f), _sdrSensitivity(10.0f),
          _averageSurpriseDecay(0.01f), _surpriseLearnFactor(2.0f),
          _cellsPerColumn(8), _cellSparsity(0.2f),
          _gamma(0.99f), _gammaLambda(0.95f),
          _gateFeedForwardAlpha(0.05f), _gateLateralAlpha(0.2f), _gateThresholdAlpha(0.01f), _gateSolveIter(5),
          _qAlpha(0.02f),
          _actionAlpha(0.2f), _actionDeriveAlpha(0.05f), _actionDeriveIterations(30),
          _explorationStdDev(0.1f), _explorationBreak(0.01f), _epsilon(0.05f) {}
  };

  // the CSRL object's own members: the hyper-parameters of the INPUT nodes' SDRRL units (deep/CSRL.h:183-243)
  float _learnFeedBackPred;
  float _learnFeedBackRL;
  float _drift;
  float _averageSurpriseDecay;
  float _surpriseLearnFactor;
  int _numRecurrentInputs;
  int _cellsPerColumn;
  float _cellSparsity;
  float _gamma;
  float _gammaLambda;
  float _gateFeedForwardAlpha;
  float _gateLateralAlpha;
  float _gateThresholdAlpha;
  int _gateSolveIter;
  float _qAlpha;
  float _actionAlpha, _actionDeriveAlpha;
  int _actionDeriveIterations;
  float _explorationStdDev;
  float _explorationBreak;
  float _epsilon;
  float _sdrBaselineDecay;
  float _sdrSensitivity;
  int _sdrIterSettle;
  int _sdrIterMeasure;
  float _sdrLeak;

  CSRL()
      : _learnFeedBackPred(0.05f), _learnFeedBackRL(0.05f), _drift(0.0f),
        _averageSurpriseDecay(0.01f), _surpriseLearnFactor(2.0f),
        _numRecurrentInputs(4), _cellsPerColumn(8), _cellSparsity(0.2f),
        _gamma(0.99f), _gammaLambda(0.95f),
        _gateFeedForwardAlpha(0.05f), _gateLateralAlpha(0.1f), _gateThresholdAlpha(0.005f), _gateSolveIter(5),
        _qAlpha(0.02f),
        _actionAlpha(0.2f), _actionDeriveAlpha(0.05f), _actionDeriveIterations(30),
        _explorationStdDev(0.1f), _explorationBreak(0.01f), _epsilon(0.05f),
        _sdrBaselineDecay(0.01f), _sdrSensitivity(10.0f),
        _sdrIterSettle(17), _sdrIterMeasure(4), _sdrLeak(0.1f) {}

  void createRandom(int inputWidth, int inputHeight, int inputFeedBackRadius, const std::vector<InputType>& inputTypes,
                    const std::vector<LayerDesc>& layerDescs, float initMinWeight, float initMaxWeight, float initMinInhibition,
                    float initMaxInhibition, float initThreshold, std::mt19937& generator) {
    _layerDescs = layerDescs;
    _inputTypes = inputTypes;
    _cfg = bidi_config{};
    _cfg.variant = BIDI_CSRL;
    _cfg.in_width = inputWidth; _cfg.in_height = inputHeight; _cfg.r_infb = inputFeedBackRadius;
    _cfg.learn_infb = _learnFeedBackRL;
    _cfg.col_cells = _cellsPerColumn; _cfg.col_sparsity = _cellSparsity; _cfg.col_gamma = _gamma; _cfg.col_gamma_lambda = _gammaLambda;
    _cfg.col_alpha_ff = _gateFeedForwardAlpha; _cfg.col_alpha_lat = _gateLateralAlpha; _cfg.col_alpha_threshold = _gateThresholdAlpha;
    _cfg.col_alpha_q = _qAlpha; _cfg.col_alpha_action = _actionAlpha;
    _cfg.col_expl_stddev = _explorationStdDev; _cfg.col_expl_break = _explorationBreak;
    _cfg.csrl_n_recurrent = _numRecurrentInputs; _cfg.csrl_derive_iterations = _actionDeriveIterations;
    _cfg.csrl_derive_alpha = _actionDeriveAlpha; _cfg.csrl_surprise_factor = _surpriseLearnFactor;
    _cfg.csrl_avg_surprise_decay = _averageSurpriseDecay;
    _cfg.csrl_iter_settle = _sdrIterSettle; _cfg.csrl_iter_measure = _sdrIterMeasure; _cfg.csrl_leak = _sdrLeak;
    _cfg.init_min_w = initMinWeight; _cfg.init_max_w = initMaxWeight;
    _cfg.init_min_inh = initMinInhibition; _cfg.init_max_inh = initMaxInhibition; _cfg.init_threshold = initThreshold;
    _cfg.q_n_actions = 0;
    for (size_t i = 0; i < inputTypes.size(); i++)
      if (inputTypes[i] == _action) {
        if (_cfg.q_n_actions >= BIDI_MAX_ACTIONS) throw std::runtime_error("deep::CSRL: more action inputs than BIDI_MAX_ACTIONS");
        _cfg.q_action_idx[_cfg.q_n_actions++] = (int)i;
      }
    _descs.assign(layerDescs.size(), bidi_layer_desc{});
    for (size_t l = 0; l < layerDescs.size(); l++) {
      const LayerDesc& s = layerDescs[l];
      bidi_layer_desc& d = _descs[l];
      d.width = s._width; d.height = s._height;
      d.r_ff = s._receptiveRadius; d.r_rec = s._recurrentRadius; d.r_lat = s._lateralRadius;
      d.r_pred = s._predictiveRadius; d.r_fb = s._feedBackRadius;
      d.csrl_n_recurrent = s._numRecurrentInputs;
      d.learn_ff = s._learnFeedForward; d.learn_rec = s._learnRecurrent; d.learn_lat = s._learnLateral;
      d.learn_fb = s._learnFeedBackRL; d.learn_pred = s._learnPredictionRL;
      d.iter_settle = s._sdrIterSettle; d.iter_measure = s._sdrIterMeasure; d.leak = s._sdrLeak;
      d.lambda = s._sdrLambda; d.weight_decay = s._sdrWeightDecay; d.sparsity = s._sparsity; d.learn_threshold = s._sdrLearnThreshold;
      d.max_weight_delta = 0.5f;  // deep/CSRL.cpp:423 calls IRSDR::learn without it: the default argument of sdr/IRSDR.h:73
      d.avg_surprise_decay = s._averageSurpriseDecay; d.csrl_surprise_factor = s._surpriseLearnFactor;
      d.col_cells = s._cellsPerColumn; d.col_sparsity = s._cellSparsity; d.col_gamma = s._gamma; d.col_gamma_lambda = s._gammaLambda;
      d.col_alpha_ff = s._gateFeedForwardAlpha; d.col_alpha_lat = s._gateLateralAlpha; d.col_alpha_threshold = s._gateThresholdAlpha;
      d.col_alpha_q = s._qAlpha; d.col_alpha_action = s._actionAlpha; d.csrl_derive_alpha = s._actionDeriveAlpha;
      d.csrl_derive_iterations = s._actionDeriveIterations;
      d.col_expl_stddev = s._explorationStdDev; d.col_expl_break = s._explorationBreak;
    }
    create();
    // createRandom's draws (deep/CSRL.cpp:11-187, deep/SDRRL.cpp:9-44) come from the caller's generator: its state goes to
    // the engine, which tabulates the draw order, and comes back advanced
    std::ostringstream os;
    os << generator;
    std::vector<char> back(16384);
    bidinet::detail::check(bidi_init_random_generator(_h, 0, os.str().c_str(), back.data(), (int)back.size()), _h,
                           "bidi_init_random_generator");
    std::istringstream is(back.data());
    is >> generator;
    _noiseU.assign((size_t)bidi_num_noise(_h), 0.0f);
    _noiseV.assign(_noiseU.size(), 0.0f);
  }

  // The exploration samples never depend on the network state, so they are drawn from the CALLER's generator before
  // the step is launched, in the generator's order of use: the input nodes of type _action, ascending, with ONE
  // perturbation distribution for the whole loop (deep/CSRL.cpp:232,246-253); then every SDRRL unit in visit order --
  // top layer first, nodes ascending, the input nodes last (deep/CSRL.cpp:262-330) -- one pair per action, a fresh
  // distribution per unit (deep/SDRRL.cpp:53-54,203-208).
  void simStep(float reward, std::mt19937& generator, bool learn = true) {
    size_t k = 0;
    {
      std::uniform_real_distribution<float> dist01(0.0f, 1.0f);
      std::normal_distribution<float> pertDist(0.0f, _cfg.col_expl_stddev);
      for (size_t pi = 0; pi < _inputTypes.size(); pi++)
        if (_inputTypes[pi] == _action) {
          _noiseU[k] = dist01(generator);
          _noiseV[k] = _noiseU[k] < _cfg.col_expl_break ? dist01(generator) : pertDist(generator);
          k++;
        }
    }
    auto units = [&](int nodes, int nRecurrent, float stddev, float breakChance) {
      const int na = 2 * ((int)_numActionTypes + nRecurrent);
      for (int n = 0; n < nodes; n++) {
        std::uniform_real_distribution<float> dist01(0.0f, 1.0f);
        std::normal_distribution<float> pertDist(0.0f, stddev);
        for (int a = 0; a < na; a++, k++) {
          _noiseU[k] = dist01(generator);
          _noiseV[k] = _noiseU[k] < breakChance ? dist01(generator) : pertDist(generator);
        }
      }
    };
    for (int l = (int)_descs.size() - 1; l >= 0; l--)
      units(_descs[(size_t)l].width * _descs[(size_t)l].height, _descs[(size_t)l].csrl_n_recurrent, _descs[(size_t)l].col_expl_stddev,
            _descs[(size_t)l].col_expl_break);
    units(_cfg.in_width * _cfg.in_height, _cfg.csrl_n_recurrent, _cfg.col_expl_stddev, _cfg.col_expl_break);
    if (k != _noiseU.size()) throw std::runtime_error("deep::CSRL::simStep: exploration draw count does not match the engine's");
    step(&reward, _noiseU.data(), _noiseV.data(), learn);
    // the step ends with the first coder's visible states set to the input nodes' outputs (deep/CSRL.cpp:441-448): an
    // input of type _action keeps that value, one of type _state is overwritten by the caller's next setInput
    (void)prediction(0);
    for (size_t pi = 0; pi < _inputs.size(); pi++) _inputs[pi] = _predictions[pi];
  }

  void setInput(int index, float value) {
    assert(_inputTypes[(size_t)index] == _state);
    _inputs[(size_t)index] = value;
  }
  void setInput(int x, int y, float value) { setInput(x + y * _cfg.in_width, value); }  // deep/CSRL.h:250-252: the visible width
  float getPrediction(int index) const { return prediction(index); }
  float getPrediction(int x, int y) const { return getPrediction(x + y * _cfg.in_width); }
  const std::vector<LayerDesc>& getLayerDescs() const { return _layerDescs; }

private:
  std::vector<LayerDesc> _layerDescs;
  std::vector<InputType> _inputTypes;
  std::vector<float> _noiseU, _noiseV;
};
}  // namespace deep
