// neo::AgentBatch -- many neo::Agent hierarchies of one shape stepped together, over one or several GPUs of
// this process (bidi_create_sharded in bidinet_b200.h).  The reference has no such class: there a second agent is
// a second neo::Agent object (RunnerMain.cpp:100-106), stepped one after the other on one CPU thread.  The batch
// keeps Agent's vocabulary -- the same LayerDesc, the same public hyper-parameter members (set them on `params`
// before createRandom), setInput / simStep / getPrediction with an agent index in front.
//
// Agent g is created from std::mt19937(seedBase + g) in createRandom's draw order, i.e. exactly the agent
// `std::mt19937 generator(seedBase + g); agent.createRandom(..., generator);` builds.
#pragma once
#include "Agent.h"

namespace neo {
class AgentBatch {
public:
  Agent params;  // hyper-parameters of every agent of the batch (neo/Agent.h:125-135)

  AgentBatch() = default;
  AgentBatch(const AgentBatch&) = delete;
  AgentBatch& operator=(const AgentBatch&) = delete;
  ~AgentBatch() { if (_s) bidi_sharded_destroy(_s); }

  // devices: CUDA device of every shard (agents are split into devices.size() contiguous blocks)
  void createRandom(int numAgents, const std::vector<int>& devices, int inputWidth, int inputHeight, int inputFeedBackRadius,
                    const std::vector<Agent::LayerDesc>& layerDescs, float initMinWeight, float initMaxWeight,
                    float initMinInhibition, float initMaxInhibition, float initThreshold, unsigned seedBase) {
    if (_s) { bidi_sharded_destroy(_s); _s = nullptr; }
    params.describe(inputWidth, inputHeight, inputFeedBackRadius, layerDescs, initMinWeight, initMaxWeight, initMinInhibition,
                    initMaxInhibition, initThreshold);
    bidi_sharded* s = nullptr;
    if (bidi_create_sharded(&params.config(), params.descs().data(), numAgents, devices.data(), (int)devices.size(), &s) != BIDI_OK)
      throw std::runtime_error(std::string("bidi_create_sharded: ") + bidi_sharded_last_error(nullptr));
    _s = s;
    check(bidi_sharded_init_random(_s, seedBase), "bidi_sharded_init_random");
    _agents = numAgents;
    _numInputs = inputWidth * inputHeight;
    _noise = bidi_num_noise(bidi_sharded_shard(_s, 0, nullptr, nullptr));
    _inputs.assign((size_t)_agents * _numInputs, 0.0f);
    _predictions.assign(_inputs.size(), 0.0f);
    _predictionsFresh = true;
  }

  int getNumAgents() const { return _agents; }
  int getNumInputs() const { return _numInputs; }
  void setInput(int agent, int index, float value) { _inputs[(size_t)agent * _numInputs + index] = value; }
  // neo/Agent.h:154-156 indexes with the first hidden layer's width; kept
  void setInput(int agent, int x, int y, float value) { setInput(agent, x + y * params.getLayerDescs().front()._width, value); }
  float* inputs(int agent) { return _inputs.data() + (size_t)agent * _numInputs; }

  // one simStep of every agent; exploration from the engine's counter-based generator (stream = global agent index)
  void simStep(const std::vector<float>& rewards, bool learn = true) { step(rewards, nullptr, nullptr, learn); }
  // ... or drawn from one caller-owned generator per agent, exactly as neo::Agent::simStep draws them
  void simStep(const std::vector<float>& rewards, std::vector<std::mt19937>& generators, bool learn = true) {
    _noiseU.resize((size_t)_agents * _noise); _noiseV.resize(_noiseU.size());
    for (int a = 0; a < _agents; a++) params.drawExploration(generators[(size_t)a], &_noiseU[(size_t)a * _noise], &_noiseV[(size_t)a * _noise]);
    step(rewards, _noiseU.data(), _noiseV.data(), learn);
  }

  float getPrediction(int agent, int index) const {
    if (!_predictionsFresh) {
      check(bidi_sharded_get_predictions(_s, _predictions.data()), "bidi_sharded_get_predictions");
      _predictionsFresh = true;
    }
    return _predictions[(size_t)agent * _numInputs + index];
  }
  float getPrediction(int agent, int x, int y) const { return getPrediction(agent, x + y * params.config().in_width); }
  bidi_sharded* handle() const { return _s; }

private:
  bidi_sharded* _s = nullptr;
  int _agents = 0, _numInputs = 0, _noise = 0;
  std::vector<float> _inputs, _noiseU, _noiseV;
  mutable std::vector<float> _predictions;
  mutable bool _predictionsFresh = false;

  void check(int rc, const char* what) const {
    if (rc != BIDI_OK) throw std::runtime_error(std::string(what) + ": " + bidi_sharded_last_error(_s));
  }
  void step(const std::vector<float>& rewards, const float* u, const float* v, bool learn) {
    if ((int)rewards.size() != _agents) throw std::runtime_error("AgentBatch::simStep: one reward per agent");
    check(bidi_sharded_set_inputs(_s, _inputs.data()), "bidi_sharded_set_inputs");
    check(bidi_sharded_step(_s, rewards.data(), u, v, learn ? 1 : 0), "bidi_sharded_step");
    _predictionsFresh = false;
  }
};
}  // namespace neo
