// bidi_coder.cuh -- the whole up pass of one layer's iterative sparse coder in ONE
// launch: sdr::IRSDR::activate (sdr/IRSDR.cpp:125-216) and
// neo::SparseCoder::activate (neo/SparseCoder.cpp:116-176).
//
// One CTA per agent.  The layer's feed-forward and recurrent slot planes are staged
// into shared memory once and stay there for all 30-35 iterations, so the weights
// cross HBM once per step instead of once per iteration; the unit-state vectors
// (errors, reconstructions, activations, spikes) never leave shared memory until the
// end.  Every iteration is two block barriers: sweep | reconstruct+errors.
//
//   * sweep: thread i owns hidden unit i and walks its slot planes in the
//     reference's list order (plane s is w[s][0..U): coalesced, conflict-free).  The
//     error maps are stored with a zero halo of the window radius and the slot planes
//     hold 0 where a unit has no connection, so the inner loop is a bare
//     load-load-multiply-add with no bounds test: a slot the reference does not have
//     contributes (+-0) to a sum that starts at +0, which leaves it unchanged.
//   * lateral inhibition is sparse: spikePrev is 0 or 1, so only the spiking
//     neighbours contribute, exactly their weight.  Warp 0 compacts the spikes into a
//     list in the lateral lists' own order (x-major: dx outer, dy inner).
//   * reconstruction is the gather form of the reference's scatter: each visible /
//     hidden cell sums over the units with a non-zero drive (a bit mask kept by the
//     sweep), ascending unit index = the order the scatter adds them.
// Used when the planes fit in shared memory (small per-agent layers: C1, C2/C3,
// Experiment_Prediction); larger layers run the per-phase kernels of
// bidi_kernels.cuh.
#pragma once
#include "bidi_kernels.cuh"

namespace bidi {

struct CoderCtaShape {
  int fhx, fhy, fpitch, fn;  // feed-forward error map: halo, row pitch, floats
  int rhx, rhy, rpitch, rn;  // recurrent error map
  int nff, nrec, words;
  long long floats;          // dynamic shared memory, in floats
};
__host__ __device__ inline CoderCtaShape coder_cta_shape(const LayerDev& L) {
  CoderCtaShape s;
  const bool has_rec = L.rec.r >= 0;
  s.fhx = L.ff.absx ? 0 : L.ff.r; s.fhy = L.ff.absy ? 0 : L.ff.r;
  s.fpitch = L.vw + 2 * s.fhx + 1; s.fn = s.fpitch * (L.vh + 2 * s.fhy + 1);
  s.rhx = has_rec && !L.rec.absx ? L.rec.r : 0; s.rhy = has_rec && !L.rec.absy ? L.rec.r : 0;
  s.rpitch = L.hw + 2 * s.rhx + 1; s.rn = s.rpitch * (L.hh + 2 * s.rhy + 1);
  s.nff = L.ff.nslots * L.nh; s.nrec = has_rec ? L.rec.nslots * L.nh : 0;
  s.words = (L.nh + 31) / 32;
  s.floats = (long long)s.nff + s.nrec + s.fn + s.rn + 2LL * L.nv + 6LL * L.nh  // planes, error maps, x/vrec, unit state
             + 2LL * L.nh + L.nh + 2LL * s.words + 8;                            // centre tables, lateral list, masks
  return s;
}

// sum over the units in `mask` whose window contains target (tx,ty), ascending unit index
__device__ __forceinline__ float gather_masked(const WinDev& w, const float* __restrict__ wp, const float* __restrict__ drive,
                                               const unsigned* __restrict__ mask, int words, const int* __restrict__ centre,
                                               int tx, int ty, bool scaled, float mult) {
  float sum = 0.0f;
  const int r = w.r;
  for (int wd = 0; wd < words; wd++) {
    for (unsigned m = mask[wd]; m; m &= m - 1) {
      const int u = wd * 32 + __ffs(m) - 1;
      const int c = centre[u], dx = tx - (c & 0xffff), dy = ty - (c >> 16);
      if (dx >= -r && dx <= r && dy >= -r && dy <= r && !(w.excl && (dx | dy) == 0)) {
        const int sx = w.absx ? tx : dx + r, sy = w.absy ? ty : dy + r;
        const float wv = wp[(sx * w.dyn + sy) * w.U + u], d = drive[u];
        sum += scaled ? wv * d * mult : wv * d;
      }
    }
  }
  return sum;
}

template <int NEO>
__global__ void __launch_bounds__(512) k_coder_cta(EngineDev P, LayerDev L, int l, long long next_vis_input) {
  extern __shared__ __align__(16) float sm[];
  const int a = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  float* B = blob_of(P, a);
  const int nv = L.nv, nh = L.nh, hw = L.hw, hh = L.hh, vw = L.vw;
  const CoderCtaShape S = coder_cta_shape(L);
  const bool has_rec = L.rec.r >= 0;
  float* wff = sm;
  float* wrec = wff + S.nff;
  float* ev = wrec + S.nrec;   // visible error, halo padded
  float* eh = ev + S.fn;       // hidden error, halo padded
  float* x = eh + S.rn;
  float* vrec = x + nv;
  float* sprev = vrec + nv;    // statePrev
  float* hrec = sprev + nh;
  float* act = hrec + nh;
  float* state = act + nh;
  float* spike = state + nh;
  float* thr = spike + nh;
  int* c_ff = reinterpret_cast<int*>(thr + nh);   // per unit: ff window centre, cx | cy << 16
  int* c_self = c_ff + nh;                         // per unit: own coordinates
  int* list = c_self + nh;                         // spiking units in lateral-list order
  unsigned* m_drive = reinterpret_cast<unsigned*>(list + nh);  // units with a non-zero drive
  unsigned* m_state = m_drive + S.words;           // units with a non-zero state (ITERATIVE's final reconstruction)
  int* list_n = reinterpret_cast<int*>(m_state + S.words);

  // ---- stage the planes and the unit state
  {
    const float* gff = B + L.ff.w_off;
    for (int k = tid; k < S.nff; k += T) wff[k] = gff[k];
    if (has_rec) {
      const float* grec = B + L.rec.w_off;
      for (int k = tid; k < S.nrec; k += T) wrec[k] = grec[k];
    }
    for (int k = tid; k < S.fn; k += T) ev[k] = 0.0f;
    for (int k = tid; k < S.rn; k += T) eh[k] = 0.0f;
    const float* gx = vis_input_of(P, L, l, a);
    for (int k = tid; k < nv; k += T) { x[k] = gx[k]; vrec[k] = B[L.vis_recon + k]; }
    for (int k = tid; k < nh; k += T) {
      sprev[k] = B[L.state_prev + k]; hrec[k] = B[L.recon + k]; thr[k] = B[L.thr + k];
      spike[k] = B[L.spike_prev + k];  // spikePrev persists across steps (sdr/IRSDR.cpp:131-135 resets only act, state)
      act[k] = 0.0f; state[k] = 0.0f;
      const int ux = k % hw, uy = k / hw;
      c_ff[k] = L.ff.cx[ux] | (L.ff.cy[uy] << 16);
      c_self[k] = ux | (uy << 16);
    }
    for (int k = tid; k < 2 * S.words; k += T) m_drive[k] = 0u;
  }
  __syncthreads();
  const float* wlat = B + L.lat.w_off;
  const bool gen = NEO && P.gen_on;
  const float* gen_nz = gen ? P.gen_normals + (long long)a * P.gen_per_agent + L.gen_off : nullptr;
  const float gen_scale = gen ? *P.gen_scale : 0.0f;
  const int total = NEO ? L.iter_settle : L.iter_settle + L.iter_measure;
  const float measure_inv = NEO ? 0.0f : 1.0f / (float)L.iter_measure;
  float counter = 0.0f, mult = 1.0f;

  for (int it = -1; it < total; it++) {
    if (it >= 0) {
      // ---- leaky integrate-and-fire sweep (sdr/IRSDR.cpp:144-168)
      const int n_spk = *list_n;
      const bool measuring = !NEO && it >= L.iter_settle;
      for (int i0 = 0; i0 < nh; i0 += T) {
        const int i = i0 + tid;
        float sp = 0.0f;
        if (i < nh) {
          const int cs = c_self[i], ux = cs & 0xffff, uy = cs >> 16;
          float excitation = gen ? gen_nz[(long long)it * nh + i] * gen_scale : 0.0f;  // simStepGenerate (neo/SparseCoder.cpp:200)
          {
            const int cf = c_ff[i];
            const float* ep = ev + (L.ff.absx ? 0 : (cf & 0xffff)) + (L.ff.absy ? 0 : (cf >> 16)) * S.fpitch;
            const float* wp = wff + i;
            for (int sx = 0; sx < L.ff.dxn; sx++)
#pragma unroll 4
              for (int sy = 0; sy < L.ff.dyn; sy++) { excitation += ep[sx + sy * S.fpitch] * wp[0]; wp += nh; }
          }
          if (has_rec) {
            const float* ep = eh + (L.rec.absx ? 0 : ux) + (L.rec.absy ? 0 : uy) * S.rpitch;
            const float* wp = wrec + i;
            for (int sx = 0; sx < L.rec.dxn; sx++)
#pragma unroll 4
              for (int sy = 0; sy < L.rec.dyn; sy++) { excitation += ep[sx + sy * S.rpitch] * wp[0]; wp += nh; }
          }
          float inhibition = 0.0f;
          const int r = L.lat.r;
          for (int e = 0; e < n_spk; e++) {
            const int c = list[e], tx = c & 0xffff, ty = c >> 16, dx = tx - ux, dy = ty - uy;
            if (dx >= -r && dx <= r && dy >= -r && dy <= r && (dx | dy) != 0) {
              const int sx = L.lat.absx ? tx : dx + r, sy = L.lat.absy ? ty : dy + r;
              inhibition += wlat[(long long)(sx * L.lat.dyn + sy) * nh + i];
            }
          }
          float av = (1.0f - L.leak) * act[i] + excitation - inhibition;
          if (av > thr[i]) { av = 0.0f; sp = 1.0f; }
          act[i] = av;
          spike[i] = sp;
          if (NEO) state[i] += sp;
          else if (measuring) state[i] += measure_inv * sp;
        }
        const unsigned m = __ballot_sync(0xffffffffu, sp != 0.0f);
        if ((tid & 31) == 0) {
          const int wd = (i0 + tid) >> 5;
          if (wd < S.words) {
            if (NEO) m_drive[wd] |= m; else m_drive[wd] = m;
            if (measuring) m_state[wd] |= m;
          }
        }
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
        int c = 0;
        bool f = false;
        if (p < nh) { const int ux = p / hh, uy = p % hh; c = ux | (uy << 16); f = spike[ux + uy * hw] != 0.0f; }
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (f) list[base + __popc(m & ((1u << tid) - 1u))] = c;
        base += __popc(m);
      }
      if (tid == 0) *list_n = base;
    }
    // ---- reconstruction + errors (sdr/IRSDR.cpp:139-142, :218-235; neo/SparseCoder.cpp:164-168)
    const float* drive = NEO ? state : spike;
    for (int j = tid; j < nv + nh; j += T) {
      if (j < nv) {
        const int tx = j % vw, ty = j / vw;
        if (it >= 0) vrec[j] = gather_masked(L.ff, wff, drive, m_drive, S.words, c_ff, tx, ty, NEO != 0, mult);
        ev[(tx + S.fhx) + (ty + S.fhy) * S.fpitch] = x[j] - vrec[j];
      } else {
        const int h = j - nv, tx = h % hw, ty = h / hw;
        if (it >= 0) hrec[h] = has_rec ? gather_masked(L.rec, wrec, drive, m_drive, S.words, c_self, tx, ty, NEO != 0, mult) : 0.0f;
        if (has_rec) eh[(tx + S.rhx) + (ty + S.rhy) * S.rpitch] = sprev[h] - hrec[h];
      }
    }
    __syncthreads();
  }

  if (!NEO) {  // reconstructFromStates, sdr/IRSDR.cpp:215
    for (int j = tid; j < nv + nh; j += T) {
      if (j < nv) B[L.vis_recon + j] = gather_masked(L.ff, wff, state, m_state, S.words, c_ff, j % vw, j / vw, false, 0.0f);
      else {
        const int h = j - nv;
        B[L.recon + h] = has_rec ? gather_masked(L.rec, wrec, state, m_state, S.words, c_self, h % hw, h / hw, false, 0.0f) : 0.0f;
      }
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
      else if (P.csrl) st = st * B[L.sd.a_expl + (long long)i * L.sd.na + 0];
      B[next_vis_input + i] = st;
    }
  }
}

} // namespace bidi
