// bidi_coder.cuh -- the whole up pass of one layer's iterative sparse coder in ONE
// launch: sdr::IRSDR::activate (sdr/IRSDR.cpp:125-216) and
// neo::SparseCoder::activate (neo/SparseCoder.cpp:116-176).
//
// One CTA per agent.  The layer's feed-forward and recurrent slot planes are staged
// into shared memory once and stay there for all 30-35 iterations, so the weights
// cross HBM once per step instead of once per iteration; the unit-state vectors
// (errors, reconstructions, activations, spikes) never leave shared memory until the
// end.  Every iteration is two block barriers: sweep | reconstruct+errors.
//   * sweep: thread i owns hidden unit i and walks its slot planes in the
//     reference's list order (coalesced, conflict-free: slot plane s is w[s][0..U)).
//   * lateral inhibition is sparse: spikePrev is 0 or 1, so only the spiking
//     neighbours contribute, exactly their weight.  Warp 0 compacts the spikes into a
//     list in the lateral lists' own order (x-major: dx outer, dy inner).
//   * reconstruction is the gather form of the reference's scatter (gather_recon).
// Used when the planes fit in shared memory (small per-agent layers: C1, C2/C3,
// Experiment_Prediction); larger layers run the per-phase kernels of
// bidi_kernels.cuh.
#pragma once
#include "bidi_kernels.cuh"

namespace bidi {

// floats of dynamic shared memory k_coder_cta needs for layer L
__host__ __device__ inline long long coder_cta_smem_floats(const LayerDev& L) {
  const long long nff = (long long)L.ff.nslots * L.nh, nrec = (long long)(L.rec.r < 0 ? 0 : L.rec.nslots) * L.nh;
  return nff + nrec + 3LL * L.nv + 9LL * L.nh + 8;
}

template <int NEO>
__global__ void __launch_bounds__(512) k_coder_cta(EngineDev P, LayerDev L, int l, long long next_vis_input) {
  extern __shared__ __align__(16) float sm[];
  const int a = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  float* B = blob_of(P, a);
  const int nv = L.nv, nh = L.nh, hw = L.hw, hh = L.hh;
  const int nff = L.ff.nslots * nh, nrec = (L.rec.r < 0 ? 0 : L.rec.nslots) * nh;
  float* wff = sm;
  float* wrec = wff + nff;
  float* x = wrec + nrec;
  float* ev = x + nv;
  float* vrec = ev + nv;
  float* sprev = vrec + nv;   // statePrev
  float* eh = sprev + nh;
  float* hrec = eh + nh;
  float* act = hrec + nh;
  float* state = act + nh;
  float* spike = state + nh;
  float* thr = spike + nh;
  int* list = reinterpret_cast<int*>(thr + nh);  // spiking units, lateral-list order
  int* list_n = list + nh;

  // ---- stage the planes and the unit state
  {
    const float* gff = B + L.ff.w_off;
    for (int k = tid; k < nff; k += T) wff[k] = gff[k];
    if (nrec) {
      const float* grec = B + L.rec.w_off;
      for (int k = tid; k < nrec; k += T) wrec[k] = grec[k];
    }
    const float* gx = vis_input_of(P, L, l, a);
    for (int k = tid; k < nv; k += T) { x[k] = gx[k]; vrec[k] = B[L.vis_recon + k]; }
    for (int k = tid; k < nh; k += T) {
      sprev[k] = B[L.state_prev + k]; hrec[k] = B[L.recon + k]; thr[k] = B[L.thr + k];
      spike[k] = B[L.spike_prev + k];  // spikePrev persists across steps (sdr/IRSDR.cpp:131-135 resets only act, state)
      act[k] = 0.0f; state[k] = 0.0f;
    }
  }
  __syncthreads();
  const float* wlat = B + L.lat.w_off;
  const int total = NEO ? L.iter_settle : L.iter_settle + L.iter_measure;
  const float measure_inv = NEO ? 0.0f : 1.0f / (float)L.iter_measure;
  float counter = 0.0f, mult = 1.0f;

  for (int it = -1; it < total; it++) {
    if (it >= 0) {
      // ---- leaky integrate-and-fire sweep (sdr/IRSDR.cpp:144-168)
      const int n_spk = *list_n;
      for (int i = tid; i < nh; i += T) {
        const int ux = i % hw, uy = i / hw;
        float excitation = 0.0f;
        for_each_conn(L.ff, ux, uy, [&](int s, int t) { excitation += ev[t] * wff[s * nh + i]; });
        for_each_conn(L.rec, ux, uy, [&](int s, int t) { excitation += eh[t] * wrec[s * nh + i]; });
        float inhibition = 0.0f;
        const int r = L.lat.r;
        for (int e = 0; e < n_spk; e++) {
          const int u = list[e], tx = u % hw, ty = u / hw, dx = tx - ux, dy = ty - uy;
          if (dx >= -r && dx <= r && dy >= -r && dy <= r && (dx | dy) != 0) {
            const int sx = L.lat.absx ? tx : dx + r, sy = L.lat.absy ? ty : dy + r;
            inhibition += wlat[(long long)(sx * L.lat.dyn + sy) * nh + i];
          }
        }
        float av = (1.0f - L.leak) * act[i] + excitation - inhibition;
        float sp;
        if (av > thr[i]) { av = 0.0f; sp = 1.0f; } else sp = 0.0f;
        act[i] = av;
        spike[i] = sp;
        if (NEO) state[i] += sp;
        else if (it >= L.iter_settle) state[i] += measure_inv * sp;
      }
      counter += 1.0f;
      mult = 1.0f / counter;
      __syncthreads();
    }
    // ---- spikePrev list for the next sweep (warp 0), x-major = the lateral lists' order
    if (tid < 32) {
      int base = 0;
      for (int p0 = 0; p0 < nh; p0 += 32) {
        const int p = p0 + tid;
        int u = 0;
        bool f = false;
        if (p < nh) { u = (p / hh) + (p % hh) * hw; f = spike[u] != 0.0f; }
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (f) list[base + __popc(m & ((1u << tid) - 1u))] = u;
        base += __popc(m);
      }
      if (tid == 0) *list_n = base;
    }
    // ---- reconstruction + errors (sdr/IRSDR.cpp:139-142, :218-235; neo/SparseCoder.cpp:164-168)
    const float* drive = NEO ? state : spike;
    for (int j = tid; j < nv + nh; j += T) {
      if (j < nv) {
        if (it >= 0) vrec[j] = gather_recon(L.ff, wff, drive, j % L.vw, j / L.vw, NEO != 0, mult);
        ev[j] = x[j] - vrec[j];
      } else {
        const int h = j - nv;
        if (it >= 0) hrec[h] = gather_recon(L.rec, wrec, drive, h % hw, h / hw, NEO != 0, mult);
        eh[h] = sprev[h] - hrec[h];
      }
    }
    __syncthreads();
  }

  if (!NEO) {  // reconstructFromStates, sdr/IRSDR.cpp:215
    for (int j = tid; j < nv + nh; j += T) {
      if (j < nv) B[L.vis_recon + j] = gather_recon(L.ff, wff, state, j % L.vw, j / L.vw, false, 0.0f);
      else B[L.recon + (j - nv)] = gather_recon(L.rec, wrec, state, (j - nv) % hw, (j - nv) / hw, false, 0.0f);
    }
  } else {
    for (int j = tid; j < nv; j += T) B[L.vis_recon + j] = vrec[j];
    for (int j = tid; j < nh; j += T) B[L.recon + j] = hrec[j];
  }
  const bool has_next = next_vis_input >= 0;
  for (int i = tid; i < nh; i += T) {
    float st = state[i];
    if (NEO) st *= mult;  // neo/SparseCoder.cpp:171-175
    B[L.state + i] = st;
    B[L.act + i] = act[i];
    B[L.spike + i] = spike[i];
    B[L.spike_prev + i] = spike[i];
    if (has_next) {  // next layer's input (neo/Agent.cpp:147-151)
      if (P.variant == BIDI_NEO_AGENT) st = st * B[L.col.a_expl + i * 3 + 0];
      B[next_vis_input + i] = st;
    }
  }
}

} // namespace bidi
