// tests/facade/batch_demo.cpp -- neo::AgentBatch (include/bidinet/neo/AgentBatch.h) against the same number of
// separate neo::Agent objects of the facade, which tests/facade/facade_demo.cpp in turn holds to the reference:
// agent g of the batch must be the agent `std::mt19937 gen(seedBase + g); agent.createRandom(..., gen)` builds,
// and stepping the batch with one caller-owned generator per agent must give, bit for bit, what stepping the
// separate agents with those generators gives (RunnerMain.cpp:139-154,235-251 drives one agent this way).
// usage: batch_demo [device of shard 0] [device of shard 1] ...   (default: two shards on device 0)
#include <neo/AgentBatch.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

int main(int argc, char** argv) {
  std::vector<int> devices;
  for (int k = 1; k < argc; k++) devices.push_back(std::atoi(argv[k]));
  if (devices.empty()) devices = {0, 0};
  const int numAgents = 5, steps = 12;
  const unsigned seedBase = 4242;
  std::vector<neo::Agent::LayerDesc> layerDescs(2);
  layerDescs[0]._width = 8; layerDescs[0]._height = 8;
  layerDescs[1]._width = 4; layerDescs[1]._height = 4;

  neo::AgentBatch batch;
  batch.params._columnGamma = 0.99f; batch.params._columnGammaLambda = 0.95f;
  batch.createRandom(numAgents, devices, 8, 8, 8, layerDescs, -0.01f, 0.01f, 0.01f, 0.05f, 0.1f, seedBase);

  std::vector<neo::Agent> agents(numAgents);
  std::vector<std::mt19937> gens, batchGens;
  for (int g = 0; g < numAgents; g++) {
    gens.emplace_back(seedBase + g);
    agents[g]._columnGamma = 0.99f; agents[g]._columnGammaLambda = 0.95f;
    agents[g].createRandom(8, 8, 8, layerDescs, -0.01f, 0.01f, 0.01f, 0.05f, 0.1f, gens[g]);
  }
  batchGens = gens;  // the generators as createRandom left them: the batch draws its exploration from copies

  long long mismatches = 0;
  std::vector<float> rewards(numAgents);
  for (int t = 0; t < steps; t++) {
    for (int g = 0; g < numAgents; g++) {
      for (int i = 0; i < 64; i++) {
        const float x = 0.5f + 0.5f * std::sin(0.05f * t * (1.0f + i) + 0.1f * g);
        agents[g].setInput(i, x);
        batch.setInput(g, i, x);
      }
      rewards[g] = std::cos(0.3f * t + g);
    }
    for (int g = 0; g < numAgents; g++) agents[g].simStep(rewards[g], gens[g], t != 5);
    batch.simStep(rewards, batchGens, t != 5);
    for (int g = 0; g < numAgents; g++)
      for (int i = 0; i < 64; i++) {
        const float a = agents[g].getPrediction(i), b = batch.getPrediction(g, i);
        if (std::memcmp(&a, &b, 4) != 0 && !(a == 0.0f && b == 0.0f)) mismatches++;
      }
  }
  for (int g = 0; g < numAgents; g++)
    if (gens[g]() != batchGens[g]()) mismatches++;
  batch.simStep(rewards);  // the engine's own exploration noise: runs
  if (!std::isfinite(batch.getPrediction(numAgents - 1, 63))) mismatches++;
  std::printf("batch_demo agents %d shards %d steps %d mismatches %lld\n", numAgents, (int)devices.size(), steps, mismatches);
  return mismatches == 0 ? 0 : 1;
}
