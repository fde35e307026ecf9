// bidi_coder_dense.cuh -- the fused per-agent coder kernel (bidi_coder.cuh) for layers
// whose receptive windows span the whole target grid (2r+1 >= grid size in both axes):
// the RunnerMain hierarchy (RunnerMain.cpp:141-154: 8x8 -> 8x8 -> 4x4, every radius 4) is
// of this kind.  There a slot IS a target cell, every unit sees every cell (edge units
// through zero weights), and the up pass is a dense mat-vec per iteration whose order of
// summation is the slot order.  So, unlike the general kernel:
//   * the slot planes w[slot][unit] are transposed into rows w[unit][slot] in shared
//     memory (row pitch = 4 mod 8: conflict-free 128-bit loads, one row per lane) and the
//     error vectors are kept in slot order, so the sweep is the column kernel's loop:
//     two 128-bit loads, two packed products, four chained adds per four connections;
//   * the gather-form reconstruction needs no window test (a unit outside the window has
//     weight 0 and adds +-0): each warp compacts the units with a non-zero drive into a
//     private (unit, drive) list, ascending unit index = the reference's scatter order,
//     and thread s sums w[unit][s]*drive over the list.
// Same arithmetic, same order, same bits as bidi_coder.cuh / bidi_kernels.cuh.
#pragma once
#include "bidi_coder.cuh"

namespace bidi {

struct CoderDenseShape {
  int Sf, Sr;        // row pitch of the feed-forward / recurrent weight rows (floats)
  int warps;
  long long floats;  // dynamic shared memory, in floats
};
__host__ __device__ inline int dense_pitch(int n) {
  int s = (n + 3) / 4 * 4;
  if ((s & 7) != 4) s += 4;
  return s;
}
__host__ __device__ inline bool coder_dense_applies(const LayerDev& L) {
  return L.ff.absx && L.ff.absy && (L.rec.r < 0 || (L.rec.absx && L.rec.absy));
}
__host__ __device__ inline int coder_dense_threads(const LayerDev& L) {
#ifdef BIDI_CODER_SWEEP_THREADS
  const int t = (L.nh + 31) / 32 * 32;  // experiment: only as many warps as the sweep uses
#else
  const int t = (L.nv + L.nh + 31) / 32 * 32;
#endif
  return t < 64 ? 64 : (t > 512 ? 512 : t);
}
__host__ __device__ inline CoderDenseShape coder_dense_shape(const LayerDev& L) {
  CoderDenseShape s;
  const bool has_rec = L.rec.r >= 0;
  s.Sf = dense_pitch(L.nv); s.Sr = has_rec ? dense_pitch(L.nh) : 0;
  s.warps = coder_dense_threads(L) / 32;
  s.floats = (long long)L.nh * (s.Sf + s.Sr)     // weight rows
             + s.Sf + s.Sr                        // errors, slot order
             + 2LL * L.nv + 2LL * L.nh            // x, vrec, statePrev, hrec: slot order
             + 4LL * L.nh                         // act, state, spike, thr: unit order
             + 3LL * s.warps * L.nh + s.warps     // per warp: lateral list, drive list (unit, value), counts
             + 8;
  return s;
}

// the units with src != 0, ascending unit index, into this warp's private list
__device__ __forceinline__ int compact_drive(const float* __restrict__ src, int nh, int lane, int* lu, float* ld) {
  int base = 0;
  for (int p0 = 0; p0 < nh; p0 += 32) {
    const int p = p0 + lane;
    const float d = p < nh ? src[p] : 0.0f;
    const unsigned m = __ballot_sync(0xffffffffu, d != 0.0f);
    if (d != 0.0f) { const int k = base + __popc(m & ((1u << lane) - 1u)); lu[k] = p; ld[k] = d; }
    base += __popc(m);
  }
  __syncwarp();
  return base;
}

template <int NEO>
__global__ void __launch_bounds__(512) k_coder_dense(EngineDev P, LayerDev L, int l, long long next_vis_input) {
  extern __shared__ __align__(16) float sm[];
  const int a = blockIdx.x, tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const Pair F{P.k_one2, P.k_negzero2, P.k_negone2};
  float* B = blob_of(P, a);
  const int nv = L.nv, nh = L.nh, hw = L.hw, hh = L.hh, vw = L.vw, vh = L.vh;
  const CoderDenseShape S = coder_dense_shape(L);
  const bool has_rec = L.rec.r >= 0;
  float* wff = sm;                       // [nh][Sf]
  float* wrec = wff + (size_t)nh * S.Sf; // [nh][Sr]
  float* ef = wrec + (size_t)nh * S.Sr;  // [Sf] visible error, slot order, pad 0
  float* er = ef + S.Sf;                 // [Sr]
  float* x = er + S.Sr;                  // slot order
  float* vrec = x + nv;
  float* sprev = vrec + nv;              // statePrev, slot order
  float* hrec = sprev + nh;
  float* act = hrec + nh;                // unit order from here
  float* state = act + nh;
  float* spike = state + nh;
  float* thr = spike + nh;
  int* lat_list = reinterpret_cast<int*>(thr + nh) + warp * nh;      // this warp's spiking units, lateral-list order
  int* lu = reinterpret_cast<int*>(thr + nh) + S.warps * nh + warp * nh;
  float* ld = thr + nh + 2 * S.warps * nh + warp * nh;

  // ---- stage: slot planes -> rows, unit state
  {
    const float* gff = B + L.ff.w_off;
    for (int k = tid; k < nv * nh; k += T) { const int s = k / nh, u = k - s * nh; wff[u * S.Sf + s] = gff[k]; }
    for (int k = tid; k < nh * (S.Sf - nv); k += T) { const int u = k / (S.Sf - nv), s = nv + k - u * (S.Sf - nv); wff[u * S.Sf + s] = 0.0f; }
    if (has_rec) {
      const float* grec = B + L.rec.w_off;
      for (int k = tid; k < nh * nh; k += T) { const int s = k / nh, u = k - s * nh; wrec[u * S.Sr + s] = grec[k]; }
      for (int k = tid; k < nh * (S.Sr - nh); k += T) { const int u = k / (S.Sr - nh), s = nh + k - u * (S.Sr - nh); wrec[u * S.Sr + s] = 0.0f; }
    }
    const float* gx = vis_input_of(P, L, l, a);
    for (int s = tid; s < S.Sf; s += T) {
      float e = 0.0f;
      if (s < nv) {
        const int t = s / vh + (s % vh) * vw;  // slot (tx, ty) = tx*vh + ty <-> cell tx + ty*vw
        x[s] = gx[t]; vrec[s] = B[L.vis_recon + t];
        e = x[s] - vrec[s];
      }
      ef[s] = e;
    }
    for (int s = tid; s < S.Sr; s += T) {
      float e = 0.0f;
      if (s < nh) {
        const int u = s / hh + (s % hh) * hw;
        sprev[s] = B[L.state_prev + u]; hrec[s] = B[L.recon + u];
        e = sprev[s] - hrec[s];
      }
      er[s] = e;
    }
    for (int k = tid; k < nh; k += T) {
      thr[k] = B[L.thr + k];
      spike[k] = B[L.spike_prev + k];  // spikePrev persists across steps (sdr/IRSDR.cpp:131-135 resets only act, state)
      act[k] = 0.0f; state[k] = 0.0f;
    }
  }
  __syncthreads();
  // spikePrev in the lateral lists' order (x-major: dx outer, dy inner), one private copy per
  // sweep warp; rebuilt after every sweep, once all units' new spikes are in shared memory
  const int sweep_warps = (nh + 31) >> 5;
  int n_spk = 0;
  auto build_lat_list = [&]() {
    n_spk = 0;
    for (int p0 = 0; p0 < nh; p0 += 32) {
      const int p = p0 + lane;
      int cc = 0;
      bool f = false;
      if (p < nh) { const int ux = p / hh, uy = p - ux * hh; cc = ux | (uy << 16); f = spike[ux + uy * hw] != 0.0f; }
      const unsigned m = __ballot_sync(0xffffffffu, f);
      if (f) lat_list[n_spk + __popc(m & ((1u << lane) - 1u))] = cc;
      n_spk += __popc(m);
    }
    __syncwarp();
  };
  if (warp < sweep_warps) build_lat_list();
  const float* wlat = B + L.lat.w_off;
  const bool gen = NEO && P.gen_on;
  const float* gen_nz = gen ? P.gen_normals + (long long)a * P.gen_per_agent + L.gen_off : nullptr;
  const float gen_scale = gen ? *P.gen_scale : 0.0f;
  const int total = NEO ? L.iter_settle : L.iter_settle + L.iter_measure;
  const float measure_inv = NEO ? 0.0f : 1.0f / (float)L.iter_measure;
  float counter = 0.0f, mult = 1.0f;
  const int nf4 = S.Sf >> 2, nr4 = S.Sr >> 2;
  const ulonglong2* ef4 = reinterpret_cast<const ulonglong2*>(ef);
  const ulonglong2* er4 = reinterpret_cast<const ulonglong2*>(er);

  for (int it = 0; it < total; it++) {
    // ---- leaky integrate-and-fire sweep (sdr/IRSDR.cpp:144-168, neo/SparseCoder.cpp:129-158)
    if (warp < sweep_warps) {
      const int i = tid;
      const bool measuring = !NEO && it >= L.iter_settle;
      float sp = 0.0f, av = 0.0f, stv = 0.0f;
      if (i < nh) {
        // the lateral weights of the first few spiking neighbours are fetched from HBM/L2 now,
        // so that their latency hides behind the excitation sums (an absent entry adds +0)
        constexpr int LATQ = 8;
        const int r = L.lat.r, ux = i % hw, uy = i / hw;
        float lw[LATQ];
#pragma unroll
        for (int e = 0; e < LATQ; e++) {
          lw[e] = 0.0f;
          if (e < n_spk) {
            const int cc = lat_list[e], tx = cc & 0xffff, ty = cc >> 16, dx = tx - ux, dy = ty - uy;
            if (dx >= -r && dx <= r && dy >= -r && dy <= r && (dx | dy) != 0) {
              const int sx = L.lat.absx ? tx : dx + r, sy = L.lat.absy ? ty : dy + r;
              lw[e] = wlat[(long long)(sx * L.lat.dyn + sy) * nh + i];
            }
          }
        }
        // simStepGenerate: the excitation starts from noiseDist(generator) * noise (neo/SparseCoder.cpp:200)
        float excitation = gen ? gen_nz[(long long)it * nh + i] * gen_scale : 0.0f;
        const ulonglong2* w4 = reinterpret_cast<const ulonglong2*>(wff + i * S.Sf);
#pragma unroll 4
        for (int j = 0; j < nf4; j++) {
          const ulonglong2 w = w4[j], e = ef4[j];
          const u64 p01 = F.mul(e.x, w.x), p23 = F.mul(e.y, w.y);
          excitation += lo32(p01); excitation += hi32(p01); excitation += lo32(p23); excitation += hi32(p23);
        }
        if (has_rec) {
          const ulonglong2* r4 = reinterpret_cast<const ulonglong2*>(wrec + i * S.Sr);
#pragma unroll 4
          for (int j = 0; j < nr4; j++) {
            const ulonglong2 w = r4[j], e = er4[j];
            const u64 p01 = F.mul(e.x, w.x), p23 = F.mul(e.y, w.y);
            excitation += lo32(p01); excitation += hi32(p01); excitation += lo32(p23); excitation += hi32(p23);
          }
        }
        float inhibition = 0.0f;
#pragma unroll
        for (int e = 0; e < LATQ; e++) inhibition += lw[e];
        for (int e = LATQ; e < n_spk; e++) {
          const int cc = lat_list[e], tx = cc & 0xffff, ty = cc >> 16, dx = tx - ux, dy = ty - uy;
          if (dx >= -r && dx <= r && dy >= -r && dy <= r && (dx | dy) != 0) {
            const int sx = L.lat.absx ? tx : dx + r, sy = L.lat.absy ? ty : dy + r;
            inhibition += wlat[(long long)(sx * L.lat.dyn + sy) * nh + i];
          }
        }
        av = (1.0f - L.leak) * act[i] + excitation - inhibition;
        if (av > thr[i]) { av = 0.0f; sp = 1.0f; }
        stv = state[i];
        if (NEO) stv += sp;
        else if (measuring) stv += measure_inv * sp;
      }
      if (i < nh) { act[i] = av; spike[i] = sp; state[i] = stv; }
    }
    counter += 1.0f;
    mult = 1.0f / counter;
    __syncthreads();
    // ---- reconstruction + errors (sdr/IRSDR.cpp:139-142, :218-235; neo/SparseCoder.cpp:164-168)
    {
      if (warp < sweep_warps) build_lat_list();
      const int n = compact_drive(NEO ? state : spike, nh, lane, lu, ld);
      for (int j = tid; j < nv + nh; j += T) {
        if (j < nv) {
          const float* col = wff + j;
          float sum = 0.0f;
          for (int k = 0; k < n; k++) { const float wv = col[lu[k] * S.Sf], d = ld[k]; sum += NEO ? wv * d * mult : wv * d; }
          vrec[j] = sum;
          ef[j] = x[j] - sum;
        } else if (has_rec) {
          const int s = j - nv;
          const float* col = wrec + s;
          float sum = 0.0f;
          for (int k = 0; k < n; k++) { const float wv = col[lu[k] * S.Sr], d = ld[k]; sum += NEO ? wv * d * mult : wv * d; }
          hrec[s] = sum;
          er[s] = sprev[s] - sum;
        }
      }
    }
    __syncthreads();
  }

  // ---- results back to the blob, unit order
  if (!NEO) {  // reconstructFromStates, sdr/IRSDR.cpp:215
    const int n = compact_drive(state, nh, lane, lu, ld);
    for (int j = tid; j < nv + nh; j += T) {
      float sum = 0.0f;
      if (j < nv) {
        for (int k = 0; k < n; k++) sum += wff[lu[k] * S.Sf + j] * ld[k];
        B[L.vis_recon + j / vh + (j % vh) * vw] = sum;
      } else {
        const int s = j - nv;
        if (has_rec) for (int k = 0; k < n; k++) sum += wrec[lu[k] * S.Sr + s] * ld[k];
        B[L.recon + s / hh + (s % hh) * hw] = sum;
      }
    }
  } else {
    for (int s = tid; s < nv; s += T) B[L.vis_recon + s / vh + (s % vh) * vw] = vrec[s];
    for (int s = tid; s < nh; s += T) B[L.recon + s / hh + (s % hh) * hw] = has_rec ? hrec[s] : 0.0f;
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
