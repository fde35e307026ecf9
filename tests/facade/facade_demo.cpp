// tests/facade/facade_demo.cpp -- one source, two builds:
//   reference build : -I/root/reference/BIDInet/source + the reference's own .cpp files
//   facade build    : -Iinclude/bidinet + libbidinet_b200.so (the CUDA engine)
// It drives the four hierarchy classes and sdr::QPRSDR exactly the way the reference's experiment mains
// do (Experiment_Prediction.cpp:113-153, RunnerMain.cpp:139-154,235-251) with fixed
// seeds and prints a hash of every prediction vector; the two builds must print the
// same text.  (The reference leaves _columnGamma/_columnGammaLambda uninitialised, so
// the demo sets them explicitly in both builds.)
#include <sdr/PredictiveRSDR.h>
#include <sdr/IPredictiveRSDR.h>
#include <sdr/QPRSDR.h>
#include <neo/PredictiveHierarchy.h>
#include <neo/Agent.h>
#include <deep/CSRL.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <random>
#include <vector>

static unsigned long long g_hash = 0xcbf29ce484222325ULL;
static void hashFloat(float v) {
  v += 0.0f;  // -0 -> +0
  unsigned char b[4];
  std::memcpy(b, &v, 4);
  for (int i = 0; i < 4; i++) { g_hash ^= b[i]; g_hash *= 0x100000001b3ULL; }
}

template <class H> static void report(const char* tag, int step, const H& h, int n) {
  for (int i = 0; i < n; i++) hashFloat(h.getPrediction(i));
  std::printf("%s %d %016llx\n", tag, step, g_hash);
}

// what the reference's own callers read through getLayers() (sdr/QPRSDR.cpp:123, sdr/ICSRL.cpp:89,
// vis/CSRLVisualizer.cpp:58): the coder's geometry and hidden states, the prediction nodes' states
template <class H> static void reportLayers(const char* tag, const H& h) {
  for (size_t l = 0; l < h.getLayers().size(); l++) {
    const auto& layer = h.getLayers()[l];
    for (int i = 0; i < layer._sdr.getNumHidden(); i++) hashFloat(layer._sdr.getHiddenState(i));
    for (size_t i = 0; i < layer._predictionNodes.size(); i++) {
      hashFloat(layer._predictionNodes[i]._state);
      hashFloat(layer._predictionNodes[i]._statePrev);
    }
    std::printf("%s layer %d: visible %dx%d hidden %dx%d, %d prediction nodes, %016llx\n", tag, (int)l, layer._sdr.getVisibleWidth(),
                layer._sdr.getVisibleHeight(), layer._sdr.getHiddenWidth(), layer._sdr.getHiddenHeight(), (int)layer._predictionNodes.size(), g_hash);
  }
}

int main() {
  {  // Experiment_Prediction.cpp:91-153
    std::mt19937 generator(1234);
    const float series[10][3] = {{0, 1, 0}, {0, 0, 0}, {1, 1, 0}, {0, 0, 1}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}, {0, 0, 0}, {0, 1, 0}, {0, 1, 1}};
    std::vector<sdr::IPredictiveRSDR::LayerDesc> layerDescs(3);
    layerDescs[0]._width = 16; layerDescs[0]._height = 16;
    layerDescs[1]._width = 12; layerDescs[1]._height = 12;
    layerDescs[2]._width = 8; layerDescs[2]._height = 8;
    sdr::IPredictiveRSDR prsdr;
    prsdr.createRandom(2, 2, 16, layerDescs, -0.01f, 0.01f, 0.01f, 0.05f, 0.0f, generator);
    float avgError = 1.0f;
    for (int iter = 0; iter < 6; iter++)
      for (int i = 0; i < 10; i++) {
        float error = 0.0f;
        for (int j = 0; j < 3; j++) {
          error += std::pow(prsdr.getPrediction(j) - series[i][j], 2);
          prsdr.setInput(j, series[i][j]);
        }
        avgError = (1.0f - 0.01f) * avgError + 0.01f * error;
        prsdr.simStep(generator);
        report("iprsdr", iter * 10 + i, prsdr, 4);
      }
    hashFloat(avgError);
    reportLayers("iprsdr", prsdr);
    std::printf("iprsdr avgError %.9g generator %u\n", avgError, (unsigned)generator());
  }
  {  // sdr::PredictiveRSDR on a 16x16 moving pattern, setInput(x, y, v)
    std::mt19937 generator(99);
    std::vector<sdr::PredictiveRSDR::LayerDesc> layerDescs(3);
    layerDescs[1]._width = 8; layerDescs[1]._height = 8;
    layerDescs[2]._width = 4; layerDescs[2]._height = 4;
    sdr::PredictiveRSDR prsdr;
    prsdr.createRandom(16, 16, layerDescs, -0.01f, 0.01f, 0.1f, generator);
    for (int t = 0; t < 60; t++) {
      const int p = t % 10;
      for (int y = 0; y < 16; y++)
        for (int x = 0; x < 16; x++) prsdr.setInput(x, y, ((x + p) % 8 < 2 && (y + 2 * p) % 8 < 3) ? 1.0f : 0.0f);
      prsdr.simStep(t % 7 != 6);
      report("prsdr", t, prsdr, 256);
    }
    reportLayers("prsdr", prsdr);
    std::printf("prsdr generator %u\n", (unsigned)generator());
  }
  {  // neo::PredictiveHierarchy, non-square input, prediction fed back when not learning
    std::mt19937 generator(7);
    std::vector<neo::PredictiveHierarchy::LayerDesc> layerDescs(2);
    layerDescs[0]._width = 10; layerDescs[0]._height = 6; layerDescs[0]._sdrIter = 12;
    layerDescs[1]._width = 5; layerDescs[1]._height = 4; layerDescs[1]._sdrIter = 9;
    neo::PredictiveHierarchy ph;
    ph.createRandom(9, 7, 3, layerDescs, -0.01f, 0.01f, 0.01f, 0.05f, 0.1f, generator);
    for (int t = 0; t < 40; t++) {
      for (int i = 0; i < 63; i++) ph.setInput(i, t < 30 ? 0.5f + 0.5f * std::sin(0.3f * t + 0.7f * i) : ph.getPrediction(i));
      ph.simStep(generator, t < 30);
      report("neoph", t, ph, 63);
    }
    for (int t = 0; t < 6; t++) {  // noise-driven dreaming on its own predictions (neo/PredictiveHierarchy.h:110)
      for (int i = 0; i < 63; i++) ph.setInput(i, ph.getPrediction(i));
      ph.simStepGenerate(generator, 0.2f);
      report("neoph-generate", t, ph, 63);
    }
    reportLayers("neoph", ph);
    std::printf("neoph generator %u\n", (unsigned)generator());
  }
  {  // RunnerMain.cpp:139-154,235-251, headless: sine state vector instead of the Box2D runner
    std::mt19937 generator(4321);
    const int clockCount = 4, inputCount = 3 + 3 + 2 + 2 + 1 + 2 + 2 + clockCount + 1, outputCount = 3 + 3 + 2 + 2;
    neo::Agent agent;
    agent._columnGamma = 0.99f; agent._columnGammaLambda = 0.95f;
    std::vector<neo::Agent::LayerDesc> layerDescs(2);
    layerDescs[0]._width = 8; layerDescs[0]._height = 8;
    layerDescs[1]._width = 4; layerDescs[1]._height = 4;
    for (auto& d : layerDescs) { d._columnGamma = 0.99f; d._columnGammaLambda = 0.95f; }
    agent.createRandom(8, 8, 8, layerDescs, -0.01f, 0.01f, 0.01f, 0.05f, 0.1f, generator);
    for (int steps = 0; steps < 40; steps++) {
      std::normal_distribution<float> noiseDist(0.0f, 0.05f);
      std::vector<float> state(15);
      for (int i = 0; i < 15; i++) state[i] = 0.5f + 0.5f * std::sin(0.05f * steps * (1 + i));
      for (int a = 0; a < clockCount; a++) state.push_back(std::sin(steps / 60.0f * 2.0f * a * 2.0f * 3.141596f) * 0.5f + 0.5f);
      for (size_t i = 0; i < state.size(); i++) agent.setInput((int)i, state[i]);
      float reward = 0.0f;
      for (int i = 0; i < outputCount; i++) {
        const float a = agent.getPrediction(inputCount + i) * 0.5f + 0.5f;
        reward += std::min(1.0f, std::max(0.0f, a));
        agent.setInput(inputCount + i, std::min(1.0f, std::max(0.0f, a + noiseDist(generator))));
      }
      agent.simStep(reward / outputCount - 0.5f, generator);
      report("agent", steps, agent, 64);
    }
    reportLayers("agent", agent);
    std::printf("agent generator %u\n", (unsigned)generator());
  }
  {  // sdr::QPRSDR: states in, actions fed back as inputs, reward, the way its createRandom/simStep/getAction are meant
     // to be driven (sdr/QPRSDR.h:109-126).  The reference never initialises _prevValue (sdr/QPRSDR.h:66): the
     // object is built in zeroed storage so that both builds start from 0.
    std::mt19937 generator(4321);
    std::vector<sdr::IPredictiveRSDR::LayerDesc> layerDescs(2);
    layerDescs[0]._width = 8; layerDescs[0]._height = 8;
    layerDescs[1]._width = 4; layerDescs[1]._height = 4;
    std::vector<int> actionIndices, antiActionIndices;
    for (int i = 0; i < 6; i++) { actionIndices.push_back(40 + i); antiActionIndices.push_back(48 + i); }
    void* storage = std::calloc(1, sizeof(sdr::QPRSDR));
    sdr::QPRSDR* agent = new (storage) sdr::QPRSDR();
    agent->createRandom(8, 8, 4, actionIndices, antiActionIndices, layerDescs, -0.3f, 0.3f, 0.01f, 0.05f, 0.0f, generator);
    for (int t = 0; t < 25; t++) {
      for (int i = 0; i < 40; i++) agent->setState(i, 0.5f + 0.5f * std::sin(0.1f * t * (1 + i)));
      for (int i = 0; i < 6; i++) {
        agent->setState(40 + i, agent->getAction(40 + i));
        agent->setState(48 + i, 1.0f - agent->getAction(40 + i));
      }
      const float reward = agent->getActionRel(0) - 0.5f;
      agent->simStep(reward, generator, t % 7 != 6);
      for (int i = 0; i < 12; i++) hashFloat(agent->getActionRel(i));
      std::printf("qprsdr %d %016llx\n", t, g_hash);
    }
    std::printf("qprsdr generator %u\n", (unsigned)generator());
    agent->~QPRSDR();
    std::free(storage);
  }
  {  // deep::CSRL: state inputs set by the caller, inputs of type _action left to the hierarchy (it feeds its own exploratory
     // outputs back, deep/CSRL.cpp:441-448), reward from the first action (deep/CSRL.h:245-267)
    std::mt19937 generator(777);
    std::vector<deep::CSRL::LayerDesc> layerDescs(2);
    layerDescs[0]._width = 6; layerDescs[0]._height = 6;
    layerDescs[1]._width = 3; layerDescs[1]._height = 3;
    for (size_t l = 0; l < layerDescs.size(); l++) {
      deep::CSRL::LayerDesc& d = layerDescs[l];
      d._receptiveRadius = 2; d._recurrentRadius = 1; d._lateralRadius = 2; d._predictiveRadius = 1; d._feedBackRadius = 1;
      d._numRecurrentInputs = 2 + (int)l; d._sdrIterSettle = 5; d._sdrIterMeasure = 2; d._cellsPerColumn = 4 + (int)l;
      d._actionDeriveIterations = 5;
    }
    std::vector<deep::CSRL::InputType> inputTypes(30, deep::CSRL::_state);
    for (int i = 20; i < 26; i++) inputTypes[i] = deep::CSRL::_action;
    deep::CSRL csrl;
    csrl._numRecurrentInputs = 2; csrl._cellsPerColumn = 3; csrl._actionDeriveIterations = 6; csrl._sdrIterSettle = 4; csrl._sdrIterMeasure = 2;
    csrl.createRandom(6, 5, 2, inputTypes, layerDescs, -0.3f, 0.3f, 0.01f, 0.05f, 0.05f, generator);
    std::printf("csrl created, generator %u\n", (unsigned)generator());
    for (int t = 0; t < 30; t++) {
      for (int i = 0; i < 30; i++)
        if (inputTypes[i] == deep::CSRL::_state) csrl.setInput(i, 0.5f + 0.5f * std::sin(0.2f * t + 0.7f * i));
      const float reward = csrl.getPrediction(20) - csrl.getPrediction(21);
      csrl.simStep(reward, generator, t % 6 != 5);
      if (std::getenv("DEMO_VERBOSE")) {
        std::printf("  reward %.9g predictions", reward);
        for (int i = 0; i < 30; i++) std::printf(" %.9g", csrl.getPrediction(i));
        std::printf("\n");
      }
      report("csrl", t, csrl, 30);
    }
    std::printf("csrl generator %u\n", (unsigned)generator());
  }
  return 0;
}
