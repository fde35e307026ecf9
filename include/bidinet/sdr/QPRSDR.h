// Drop-in for the reference's sdr/QPRSDR.h: an sdr::IPredictiveRSDR hierarchy plus a Q
// network over its grids, sign-gradient action derivation, exploration and TD(lambda)
// (sdr/QPRSDR.cpp:8-341).  Same public members and defaults (sdr/QPRSDR.h:84-107) and
// signatures (:109-126).
#pragma once
#include <iostream>

#include "IPredictiveRSDR.h"

namespace sdr {
class QPRSDR : public bidinet::detail::Hierarchy {
public:
  float _qAlpha;
  float _actionAlpha;
  int _actionDeriveIterations;
  float _actionDeriveAlpha;
  float _reluLeak;  // unused by the reference's simStep as well
  float _explorationBreak;
  float _explorationStdDev;
  float _gamma;
  float _gammaLambda;
  float _learnInputFeedBack;  // the owned hierarchy's member (sdr/IPredictiveRSDR.h:101)

  QPRSDR()
      : _qAlpha(0.01f), _actionAlpha(0.1f), _actionDeriveIterations(32), _actionDeriveAlpha(0.05f), _reluLeak(0.01f),
        _explorationBreak(0.01f), _explorationStdDev(0.05f), _gamma(0.99f), _gammaLambda(0.98f), _learnInputFeedBack(0.05f) {}

  void createRandom(int inputWidth, int inputHeight, int firstLayerPredictionRadius, const std::vector<int>& actionIndices,
                    const std::vector<int>& antiActionIndices, const std::vector<IPredictiveRSDR::LayerDesc>& layerDescs,
                    float initMinWeight, float initMaxWeight, float initMinInhibition, float initMaxInhibition,
                    float initThreshold, std::mt19937& generator) {
    if (actionIndices.size() != antiActionIndices.size() || actionIndices.empty() || actionIndices.size() > BIDI_MAX_ACTIONS)
      throw std::runtime_error("sdr::QPRSDR::createRandom: need 1..BIDI_MAX_ACTIONS actions and as many anti-actions");
    _layerDescs = layerDescs;
    _cfg = bidi_config{};
    _cfg.variant = BIDI_QPRSDR;
    _cfg.in_width = inputWidth; _cfg.in_height = inputHeight; _cfg.r_infb = firstLayerPredictionRadius;
    _cfg.learn_infb = _learnInputFeedBack;
    _cfg.init_min_w = initMinWeight; _cfg.init_max_w = initMaxWeight;
    _cfg.init_min_inh = initMinInhibition; _cfg.init_max_inh = initMaxInhibition; _cfg.init_threshold = initThreshold;
    _cfg.q_alpha = _qAlpha; _cfg.q_action_alpha = _actionAlpha; _cfg.q_derive_iterations = _actionDeriveIterations;
    _cfg.q_derive_alpha = _actionDeriveAlpha; _cfg.q_expl_break = _explorationBreak; _cfg.q_expl_stddev = _explorationStdDev;
    _cfg.q_gamma = _gamma; _cfg.q_gamma_lambda = _gammaLambda;
    _cfg.q_n_actions = (int)actionIndices.size();
    for (size_t k = 0; k < actionIndices.size(); k++) { _cfg.q_action_idx[k] = actionIndices[k]; _cfg.q_anti_idx[k] = antiActionIndices[k]; }
    _actionNodeIndices.assign((size_t)inputWidth * inputHeight, -1);
    for (size_t k = 0; k < actionIndices.size(); k++) _actionNodeIndices[(size_t)actionIndices[k]] = (int)k;
    _descs.assign(layerDescs.size(), bidi_layer_desc{});
    for (size_t l = 0; l < layerDescs.size(); l++) IPredictiveRSDR::fill(_descs[l], layerDescs[l]);
    create();
    drawWeights(generator);  // _prsdr.createRandom (sdr/QPRSDR.cpp:11)
    // the Q network's draws (sdr/QPRSDR.cpp:30-31, :50-83): one per top node, then per layer and node a
    // bias and one weight per in-bounds window slot, kept only on a path from an action input
    std::uniform_real_distribution<float> weightDist(initMinWeight, initMaxWeight);
    const int L = (int)layerDescs.size();
    const int top = layerDescs.back()._width * layerDescs.back()._height;
    std::vector<float> qc((size_t)top);
    for (int i = 0; i < top; i++) qc[(size_t)i] = weightDist(generator) / top;
    upload(BIDI_QCONN_W, 0, qc);
    for (int l = 0; l < L; l++) {
      const int units = layerDescs[l]._width * layerDescs[l]._height;
      const std::vector<int> nff = counts(BIDI_FF_W, l, units);
      long long total = 0;
      for (int n : nff) total += n;
      std::vector<unsigned char> kept((size_t)total);
      bidinet::detail::check(bidi_q_kept(_h, l, kept.data(), total), _h, "bidi_q_kept");
      std::vector<float> bias, ff;
      size_t k = 0;
      for (int i = 0; i < units; i++) {
        bias.push_back(weightDist(generator));
        for (int c = 0; c < nff[(size_t)i]; c++, k++) {
          const float w = weightDist(generator);
          if (kept[k]) ff.push_back(w);
        }
      }
      upload(BIDI_Q_BIAS_W, l, bias); upload(BIDI_Q_FF_W, l, ff);
    }
    _noiseU.assign(actionIndices.size(), 0.0f);
    _noiseV.assign(actionIndices.size(), 0.0f);
    _actions.assign(2 * actionIndices.size(), 0.0f);
    _actionsFresh = true;
  }

  // The exploration samples come from the CALLER's generator exactly as the reference draws
  // them (sdr/QPRSDR.cpp:202-213: one normal_distribution for the whole loop); the hierarchy's
  // own simStep draws nothing (sdr/IRSDR.cpp:126).  Like the reference, the step prints Q
  // (sdr/QPRSDR.cpp:259).
  void simStep(float reward, std::mt19937& generator, bool learn = true) {
    std::uniform_real_distribution<float> dist01(0.0f, 1.0f);
    std::normal_distribution<float> pertDist(0.0f, _explorationStdDev);
    for (size_t i = 0; i < _noiseU.size(); i++) {
      _noiseU[i] = dist01(generator);
      _noiseV[i] = _noiseU[i] < _explorationBreak ? dist01(generator) : pertDist(generator);
    }
    step(&reward, _noiseU.data(), _noiseV.data(), learn);
    _actionsFresh = false;
    std::cout << getArray(BIDI_Q_PREV_VALUE, 0)[0] << std::endl;
  }

  void setState(int index, float state) { _inputs[(size_t)index] = state; }
  // forwards to IPredictiveRSDR::setInput(x, y), which indexes with the first HIDDEN layer's width (sdr/IPredictiveRSDR.h:115-117)
  void setState(int x, int y, float state) { setState(x + y * _layerDescs.front()._width, state); }
  float getAction(int index) const { return action(_actionNodeIndices[(size_t)index]); }
  float getAction(int x, int y) const { return getAction(x + y * _cfg.in_width); }
  float getActionRel(int index) const { return action(index); }

private:
  std::vector<IPredictiveRSDR::LayerDesc> _layerDescs;
  std::vector<int> _actionNodeIndices;
  std::vector<float> _noiseU, _noiseV;
  mutable std::vector<float> _actions;
  mutable bool _actionsFresh = false;

  float action(int node) const {
    if (!_actionsFresh) {
      bidinet::detail::check(bidi_get_actions(_h, _actions.data()), _h, "bidi_get_actions");
      _actionsFresh = true;
    }
    return _actions[(size_t)node];
  }
};
}  // namespace sdr
