// bidi_coder_dense.cuh -- the fused per-agent coder kernel (bidi_coder.cuh) for layers
// whose receptive windows span the whole target grid (2r+1 >= grid size in both axes):
// the RunnerMain hierarchy (RunnerMain.cpp:141-154: 8x8 -> 8x8 -> 4x4, every radius 4) is
// of this kind.  There a slot IS a target cell, every unit sees every cell (edge units
// through zero weights), and the up pass is a dense mat-vec per iteration whose order of
// summation is the slot order.  So, unlike the general kernel:
//   * a layer that runs this kernel keeps its weights (and eligibility traces) in HBM as ROWS
//     w[unit][slot] (WinDev::row_pitch), feed-forward and recurrent part in ONE row per unit
//     (row pitch = 4 mod 8: conflict-free 128-bit loads, one row per lane) -- exactly the
//     shared-memory image, so staging is ONE TMA bulk copy (cp.async.bulk + mbarrier) and the
//     updated rows go back with one bulk store.  The error vectors are kept in slot order, so
//     the sweep is the column kernel's loop: two 128-bit loads, two packed products, four
//     chained adds per four connections, software-pipelined four quads deep;
//   * the gather-form reconstruction needs no window test (a unit outside the window has
//     weight 0 and adds +-0): each reconstructing warp compacts the units with a non-zero
//     drive into a private list of 8-byte entries (row offset, drive), ascending
//     unit index = the reference's scatter order, and a lane owns a PAIR (or four) of
//     adjacent slots: per listed unit one broadcast 64-bit load, one 64-bit weight load,
//     one or two packed products and two independent adds.  A CTA has as many warps as the
//     larger of the two phases needs (RunnerMain layers: two and one), so none idles at the
//     barriers;
//   * the lateral lists index a table with the precomputed slot-plane offset of every unit;
//   * everything but the weight rows is kept small (activations and thresholds in registers,
//     the reconstructions written over the errors at the end), so that six CTAs of the
//     RunnerMain layer 0 fit an SM: 4096 agents are then five waves instead of six;
//   * the layer's LEARNING is the kernel's tail (`learn`): everything the rules of sdr/IRSDR.cpp:287-319 and
//     neo/SparseCoder.cpp:326-361 (+ the reward of neo/Agent.cpp:252-265) read is known once the layer's own
//     up pass is over -- the final states, the last reconstructions, the PREVIOUS step's prediction -- and
//     nothing the down pass reads is written by them.  The weight rows are still in shared memory, so the
//     rule reads only the traces from HBM and writes weights and traces back: no second sweep of the
//     weights, no separate learning launch competing with the column kernels for HBM.
// Same arithmetic, same order, same bits as bidi_coder.cuh / bidi_kernels.cuh.
#pragma once
#include "bidi_coder.cuh"

namespace bidi {

// Shared-memory row of unit u: [ feed-forward slots, padded to 4 | recurrent slots, padded to 4 | pad ],
// pitch = 4 (mod 8) floats.  The error, input and reconstruction vectors have the same layout.
struct CoderDenseShape {
  int nf, nr, pitch;  // floats of the feed-forward / recurrent part of a row; row pitch
  int sweep_warps, recon_warps, warps;
  int quad;           // the reconstruction gives a lane four adjacent slots instead of two
  int tasks;          // slot pairs (or quads) the reconstruction covers
  long long floats;   // dynamic shared memory, in floats
};
__host__ __device__ inline int dense_pitch(int n) {
  int s = (n + 3) / 4 * 4;
  if ((s & 7) != 4) s += 4;
  return s;
}
__host__ __device__ inline bool coder_dense_applies(const LayerDev& L) {
  return L.ff.absx && L.ff.absy && (L.rec.r < 0 || (L.rec.absx && L.rec.absy));
}
__host__ __device__ inline CoderDenseShape coder_dense_shape(const LayerDev& L) {
  CoderDenseShape s;
  s.nf = (L.nv + 3) & ~3; s.nr = L.rec.r >= 0 ? (L.nh + 3) & ~3 : 0;
  s.pitch = dense_pitch(s.nf + s.nr);
  s.sweep_warps = (L.nh + 31) / 32;                 // one lane per unit
  const int pairs = (s.nf + s.nr) / 2;
  s.quad = (pairs + 31) / 32 > s.sweep_warps;       // pairs while they need no more warps than the sweep
  s.tasks = s.quad ? pairs / 2 : pairs;
  s.recon_warps = (s.tasks + 31) / 32;
  if (s.recon_warps > 16) s.recon_warps = 16;
  s.warps = s.sweep_warps > s.recon_warps ? s.sweep_warps : s.recon_warps;
  s.floats = 4                                      // the staging copy's mbarrier (keeps the rows 16-byte aligned)
             + (long long)L.nh * s.pitch            // weight rows
             + 2LL * s.pitch                        // errors (the reconstructions at the end), inputs (x | statePrev)
             + 2LL * s.recon_warps * L.nh           // per reconstructing warp: drive list, 8-byte entries
             + 2LL * L.nh                           // state, spike: unit order (act and thr live in registers)
             + 2LL * L.nh                           // lateral-order table: (coordinates, slot plane offset)
             + 2LL * ((L.nh + 31) / 32) + 2         // the spiking units as bit masks in lateral-list order, this sweep's and the next's
             + 8;                                   // RunnerMain layer 0: 37.4 KB, six CTAs per SM
  return s;
}
__host__ __device__ inline int coder_dense_threads(const LayerDev& L) { return 32 * coder_dense_shape(L).warps; }

// the units with src != 0, ascending unit index, into this warp's private list: entry = (row offset, drive)
__device__ __forceinline__ int compact_drive(const float* __restrict__ src, int nh, int lane, int pitch, int2* ent) {
  int base = 0;
  for (int p0 = 0; p0 < nh; p0 += 32) {
    const int p = p0 + lane;
    const float d = p < nh ? src[p] : 0.0f;
    const unsigned m = __ballot_sync(0xffffffffu, d != 0.0f);
    if (d != 0.0f) ent[base + __popc(m & ((1u << lane) - 1u))] = make_int2(p * pitch, __float_as_int(d));
    base += __popc(m);
  }
  __syncwarp();
  return base;
}

// acc + sum of w[j]*e[j] over n4 quads, added strictly in order; the 128-bit loads of the next four
// quads are issued before the sixteen chained adds of the current four.  LIGHT: two quads per trip, no
// explicit double buffering (the one-warp CTAs of small layers trade it for registers: 28 CTAs per SM)
template <int LIGHT>
__device__ __forceinline__ float dot_row_seq(const Pair& F, const ulonglong2* __restrict__ w4, const ulonglong2* __restrict__ e4, int n4, float acc) {
  if (LIGHT) {
#pragma unroll 2
    for (int j = 0; j < n4; j++) {
      const ulonglong2 w = w4[j], e = e4[j];
      const u64 p01 = F.mul(e.x, w.x), p23 = F.mul(e.y, w.y);
      acc += lo32(p01); acc += hi32(p01); acc += lo32(p23); acc += hi32(p23);
    }
    return acc;
  }
  ulonglong2 wa[4], ea[4], wb[4], eb[4];
  const int nb = n4 >> 2;
  auto load = [&](ulonglong2* w, ulonglong2* e, int j) {
#pragma unroll
    for (int t = 0; t < 4; t++) { w[t] = w4[j + t]; e[t] = e4[j + t]; }
  };
  auto eat = [&](const ulonglong2* w, const ulonglong2* e) {
    u64 pr[8];
#pragma unroll
    for (int t = 0; t < 4; t++) { pr[2 * t] = F.mul(e[t].x, w[t].x); pr[2 * t + 1] = F.mul(e[t].y, w[t].y); }
#pragma unroll
    for (int t = 0; t < 8; t++) { acc += lo32(pr[t]); acc += hi32(pr[t]); }
  };
  if (nb > 0) load(wa, ea, 0);
  int b = 0;
  for (; b + 2 <= nb; b += 2) {
    load(wb, eb, 4 * (b + 1));
    eat(wa, ea);
    if (b + 2 < nb) load(wa, ea, 4 * (b + 2));
    eat(wb, eb);
  }
  if (b < nb) eat(wa, ea);
  for (int j = nb * 4; j < n4; j++) {
    const ulonglong2 w = w4[j], e = e4[j];
    const u64 p01 = F.mul(e.x, w.x), p23 = F.mul(e.y, w.y);
    acc += lo32(p01); acc += hi32(p01); acc += lo32(p23); acc += hi32(p23);
  }
  return acc;
}

// LIGHT: the one-warp launch of a small layer (coder_dense_threads == 32), compiled for 28 CTAs per SM
template <int NEO, int LIGHT>
__global__ void __launch_bounds__(LIGHT ? 32 : 512, LIGHT ? 28 : 0) k_coder_dense(EngineDev P, LayerDev L, int l, long long next_vis_input, int learn) {
  extern __shared__ __align__(16) float sm[];
  const int a = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Pair F{P.k_one2, P.k_negzero2, P.k_negone2};
  float* B = blob_of(P, a);
  const int nv = L.nv, nh = L.nh, hw = L.hw, hh = L.hh, vw = L.vw, vh = L.vh;
  const CoderDenseShape S = coder_dense_shape(L);
  const int W = S.warps, T = 32 * W, pitch = S.pitch;
  const bool has_rec = L.rec.r >= 0;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
  float* w = sm + 4;                          // [nh][pitch]
  float* err = w + (size_t)nh * pitch;        // [pitch] errors, slot order, pads 0; after the last iteration: the reconstructions
  float* xin = err + pitch;                   // [pitch] x | statePrev
  int2* ent = reinterpret_cast<int2*>(xin + pitch) + (warp < S.recon_warps ? warp : 0) * nh;  // this warp's drive list
  float* state = xin + pitch + 2 * S.recon_warps * nh;  // unit order
  float* spike = state + nh;
  int2* lat_tab = reinterpret_cast<int2*>(spike + nh);  // [nh], lateral-list order: (x | y << 16, slot plane offset)
  const int nw = (nh + 31) >> 5;
  unsigned* msk = reinterpret_cast<unsigned*>(lat_tab + nh);  // [2][nw] bit p = the unit at position p of the lateral lists' order spiked

  // ---- stage: the layer's rows [unit][feed-forward | recurrent | pad] are one contiguous image in HBM: one TMA
  //      bulk copy, overlapped with the staging of the vectors below
  const uint32_t row_bytes = (uint32_t)(nh * pitch) * 4u;
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, row_bytes);
    tma_load_1d(w, B + L.ff.w_off, row_bytes, bar);
    // the learning tail reads the eligibility traces once, ~all iterations from now: bring them into the L2 meanwhile
    if (NEO && learn) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(B + L.ff.tr_off), "r"(row_bytes) : "memory");
    // ... and every lateral plane (the sweeps touch only the planes of spiking neighbours)
    if (learn && L.lat.r >= 0) {
      const uint32_t lb = ((uint32_t)(L.lat.nslots * nh) * 4u) & ~15u;
      if (lb) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(B + L.lat.w_off), "r"(lb) : "memory");
    }
  }
  {
    const float* gx = vis_input_of(P, L, l, a);
    constexpr int SB = LIGHT ? 2 : 4;  // loads of SB trips in flight before the first store
    for (int s0 = tid; s0 < pitch; s0 += SB * T) {
      float xv[SB], rv[SB];
#pragma unroll
      for (int h = 0; h < SB; h++) {
        const int s = s0 + h * T;
        xv[h] = 0.0f; rv[h] = 0.0f;
        if (s < nv) {
          const int t = s / vh + (s % vh) * vw;  // slot (tx, ty) = tx*vh + ty <-> cell tx + ty*vw
          xv[h] = gx[t]; rv[h] = B[L.vis_recon + t];
        } else if (s >= S.nf && s - S.nf < nh && has_rec) {
          const int k = s - S.nf, u = k / hh + (k % hh) * hw;
          xv[h] = B[L.state_prev + u]; rv[h] = B[L.recon + u];
        }
      }
#pragma unroll
      for (int h = 0; h < SB; h++) {
        const int s = s0 + h * T;
        if (s < pitch) { xin[s] = xv[h]; err[s] = xv[h] - rv[h]; }
      }
    }
    if (tid < 2 * nw) msk[tid] = 0u;
    __syncthreads();
    for (int k = tid; k < nh; k += T) {
      const float sp0 = B[L.spike_prev + k];
      spike[k] = sp0;  // spikePrev persists across steps (sdr/IRSDR.cpp:131-135 resets only act, state)
      state[k] = 0.0f;
      const int tx = k / hh, ty = k - tx * hh;  // position k of the lateral lists' order (x-major: dx outer, dy inner)
      lat_tab[k] = make_int2(tx | (ty << 16), (tx * L.lat.dyn + ty) * nh);
      if (sp0 != 0.0f) { const int pp = (k % hw) * hh + k / hw; atomicOr(&msk[pp >> 5], 1u << (pp & 31)); }
    }
  }
  mbar_wait(bar, 0);
  __syncthreads();
  // reconstruction of both sides from the units with a non-zero drive, ascending unit index
  // (sdr/IRSDR.cpp:218-235, neo/SparseCoder.cpp:164-168): a lane owns adjacent slots of the row layout.
  // It leaves the errors input - reconstruction in err[]; `last`: the reconstruction itself (the errors
  // are not read again), which is what goes back to the blob
  auto reconstruct = [&](const float* drive, bool scaled, float mult, bool last) {
    if (warp >= S.recon_warps) return;
    const int n = compact_drive(drive, nh, lane, pitch, ent);
    const u64 mult2 = pk(mult, mult);
    if (S.quad) {
      for (int j = tid; j < S.tasks; j += 32 * S.recon_warps) {
        const float* col = w + 4 * j;
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
        for (int k = 0; k < n; k++) {
          const int2 e = ent[k];
          const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(col + e.x);
          const u64 dd = pk(__int_as_float(e.y), __int_as_float(e.y));
          u64 p01 = F.mul(wv.x, dd), p23 = F.mul(wv.y, dd);
          if (scaled) { p01 = F.mul(p01, mult2); p23 = F.mul(p23, mult2); }
          s0 += lo32(p01); s1 += hi32(p01); s2 += lo32(p23); s3 += hi32(p23);
        }
        ulonglong2 out = make_ulonglong2(pk(s0, s1), pk(s2, s3));
        if (!last) {
          const ulonglong2 xv = reinterpret_cast<const ulonglong2*>(xin)[j];
          out = make_ulonglong2(F.sub2(xv.x, out.x), F.sub2(xv.y, out.y));
        }
        reinterpret_cast<ulonglong2*>(err)[j] = out;
      }
    } else {
      for (int j = tid; j < S.tasks; j += 32 * S.recon_warps) {
        const float* col = w + 2 * j;
        float s0 = 0.0f, s1 = 0.0f;
        for (int k = 0; k < n; k++) {
          const int2 e = ent[k];
          u64 pr = F.mul(*reinterpret_cast<const u64*>(col + e.x), pk(__int_as_float(e.y), __int_as_float(e.y)));
          if (scaled) pr = F.mul(pr, mult2);
          s0 += lo32(pr); s1 += hi32(pr);
        }
        const u64 sum = pk(s0, s1);
        reinterpret_cast<u64*>(err)[j] = last ? sum : F.sub2(reinterpret_cast<const u64*>(xin)[j], sum);
      }
    }
  };
  const float* wlat = B + L.lat.w_off;
  const bool gen = NEO && P.gen_on;
  const float* gen_nz = gen ? P.gen_normals + (long long)a * P.gen_per_agent + L.gen_off : nullptr;
  const float gen_scale = gen ? *P.gen_scale : 0.0f;
  const int total = NEO ? L.iter_settle : L.iter_settle + L.iter_measure;
  const float measure_inv = NEO ? 0.0f : 1.0f / (float)L.iter_measure;
  float counter = 0.0f, mult = 1.0f;
  // this lane's unit in the sweeps: window test and slot plane of a lateral-list entry (cc, base) are
  //   |tx - ux| <= r  <=>  unsigned(tx + r - ux) <= 2r ;  wlat[base + lat_off]
  const int i = tid;
  const int r = L.lat.r, ux = i % hw, uy = i / hw, own = ux | (uy << 16), tcx = r - ux, tcy = r - uy;
  const int my_pos = ux * hh + uy;
  const unsigned r2 = 2u * (unsigned)r;
  const bool lat_on = r >= 0;
  const int lat_off = ((L.lat.absx ? 0 : tcx) * L.lat.dyn + (L.lat.absy ? 0 : tcy)) * nh + i;
  float act_r = 0.0f;                                         // this lane's unit: activation and threshold stay in registers
  const float thr_r = (warp < S.sweep_warps && i < nh) ? B[L.thr + i] : 0.0f;
  auto lat_weight = [&](int2 t) -> float {
    const bool in = lat_on && (unsigned)((t.x & 0xffff) + tcx) <= r2 && (unsigned)((t.x >> 16) + tcy) <= r2 && t.x != own;
    return in ? wlat[t.y + lat_off] : 0.0f;
  };

  for (int it = 0; it < total; it++) {
    // ---- leaky integrate-and-fire sweep (sdr/IRSDR.cpp:144-168, neo/SparseCoder.cpp:129-158)
    if (warp < S.sweep_warps && i < nh) {
      const bool measuring = !NEO && it >= L.iter_settle;
      // the lateral weights of the first few spiking neighbours are fetched from HBM/L2 now,
      // so that their latency hides behind the excitation sums (an absent entry adds +0)
#ifndef BIDI_LATQ
#define BIDI_LATQ 4   // measured on C3: 2..4 within 0.1 %, 8 is 0.5 % slower, 12 is 1.3 % slower
#endif
      constexpr int LATQ = BIDI_LATQ;
      // the spiking units, ascending position in the lateral lists (the order of the reference's sum): set bits of cur
      const unsigned* cur = msk + (it & 1) * nw;
      int mw = 0;
      unsigned bits = cur[0];
      auto next_pos = [&]() -> int {
        while (bits == 0u) { if (++mw >= nw) return -1; bits = cur[mw]; }
        const int b = __ffs(bits) - 1;
        bits &= bits - 1u;
        return mw * 32 + b;
      };
      float lw[LATQ];
#pragma unroll
      for (int e = 0; e < LATQ; e++) { const int pp = next_pos(); lw[e] = pp >= 0 ? lat_weight(lat_tab[pp]) : 0.0f; }
      // simStepGenerate: the excitation starts from noiseDist(generator) * noise (neo/SparseCoder.cpp:200)
      float excitation = gen ? gen_nz[(long long)it * nh + i] * gen_scale : 0.0f;
      excitation = dot_row_seq<LIGHT>(F, reinterpret_cast<const ulonglong2*>(w + (size_t)i * pitch), reinterpret_cast<const ulonglong2*>(err),
                               pitch >> 2, excitation);
      float inhibition = 0.0f;
#pragma unroll
      for (int e = 0; e < LATQ; e++) inhibition += lw[e];
      for (int pp = next_pos(); pp >= 0; pp = next_pos()) {
        const int2 t = lat_tab[pp];
        const bool in = lat_on && (unsigned)((t.x & 0xffff) + tcx) <= r2 && (unsigned)((t.x >> 16) + tcy) <= r2 && t.x != own;
        if (in) inhibition += wlat[t.y + lat_off];
      }
      float av = (1.0f - L.leak) * act_r + excitation - inhibition, sp = 0.0f;
      if (av > thr_r) { av = 0.0f; sp = 1.0f; }
      float stv = state[i];
      if (NEO) stv += sp;
      else if (measuring) stv += measure_inv * sp;
      act_r = av; spike[i] = sp; state[i] = stv;
      if (sp != 0.0f) atomicOr(&msk[((it + 1) & 1) * nw + (my_pos >> 5)], 1u << (my_pos & 31));
    }
    counter += 1.0f;
    mult = 1.0f / counter;
    __syncthreads();
    // ---- reconstruction + errors (sdr/IRSDR.cpp:139-142, :218-235; neo/SparseCoder.cpp:164-168)
    const bool last = NEO && it == total - 1;
    if (tid < nw) msk[(it & 1) * nw + tid] = 0u;  // read by this sweep, written by the next but one
    reconstruct(NEO ? state : spike, NEO != 0, mult, last);
    __syncthreads();
  }

  // ---- results back to the blob, unit order
  if (!NEO) {  // reconstructFromStates, sdr/IRSDR.cpp:215 (the errors it also rewrites are not used again)
    reconstruct(state, false, 1.0f, true);
    __syncthreads();
  }
  if (!NEO || total > 0) {  // (no iteration: the blob keeps the reconstructions it has)
    for (int s = tid; s < nv; s += T) B[L.vis_recon + s / vh + (s % vh) * vw] = err[s];
    for (int s = tid; s < nh; s += T) B[L.recon + s / hh + (s % hh) * hw] = has_rec ? err[S.nf + s] : 0.0f;
  }
  const bool has_next = next_vis_input >= 0;
  if (warp < S.sweep_warps && i < nh) B[L.act + i] = act_r;
  for (int k = tid; k < nh; k += T) {
    float st = state[k];
    if (NEO) st *= mult;  // neo/SparseCoder.cpp:171-175
    B[L.state + k] = st;
    B[L.spike + k] = spike[k];
    B[L.spike_prev + k] = spike[k];
    if (has_next) {  // next layer's input (neo/Agent.cpp:147-151)
      if (P.variant == BIDI_NEO_AGENT) st = st * B[L.col.a_expl + k * 3 + 0];
      else if (P.csrl) st = st * B[L.sd.a_expl + (long long)k * L.sd.na + 0];
      B[next_vis_input + k] = st;
    }
  }
  if (!learn) return;

  // ---- learning (sdr/IRSDR.cpp:287-319; neo/Agent.cpp:252-265 + neo/SparseCoder.cpp:326-361), the operations of
  //      k_ic_learn / k_neo_learn in their order.  err[] holds the final reconstructions: xin becomes the error
  //      vector (slot order), state[] the learning factor of every unit, spike[] its reward
  __syncthreads();
  for (int s = tid; s < pitch; s += T) xin[s] = xin[s] - err[s];
  float* lrn = state;
  float* rwd = spike;
  for (int k = tid; k < nh; k += T) {
    const float lv = NEO ? state[k] * mult : state[k];
    if (NEO) {
      const float perr = lv - B[L.p_state_prev + k];
      const float error2 = perr * perr, base = B[L.p_baseline + k];
      rwd[k] = bidi_sigmoid(L.sensitivity * (error2 - base));
      B[L.p_baseline + k] = (1.0f - L.baseline_decay) * base + L.baseline_decay * error2;
    }
    lrn[k] = lv;
  }
  __syncthreads();
  {  // feed-forward and recurrent weights, row by row: an element is four adjacent slots of one unit (128-bit
     // accesses, consecutive lanes = consecutive quads of a row).  Which of the four are the unit's own connections
     // (inside the radius, the recurrent centre excluded) comes from the layer's mask table.  The traces stream
     // through registers (HBM rows like the weights; prefetched into the L2 at kernel start), the new weights go
     // into the shared-memory rows, which one bulk copy then takes back to HBM.
    constexpr int NB = LIGHT ? 2 : 8;
    const int pq = pitch >> 2, nq_ff = S.nf >> 2, total_q = nh * pq;
    const float maxd = L.max_weight_delta, decay = L.weight_decay, lambda = L.lambda;
    const int* __restrict__ mk = L.dense_mask;
    float4* tr4 = NEO ? reinterpret_cast<float4*>(B + L.ff.tr_off) : nullptr;
    float4* w4 = reinterpret_cast<float4*>(w);
    const float4* e4 = reinterpret_cast<const float4*>(xin);
    int u = tid / pq, q = tid - u * pq;
    const int du = T / pq, dq = T - du * pq;   // one step of T elements
    for (int idx = tid; idx < total_q; idx += NB * T) {
      float4 trv[NB];
      int msk[NB], uu[NB], qq[NB];
#pragma unroll
      for (int h = 0; h < NB; h++) {
        const int id = idx + h * T;
        uu[h] = u; qq[h] = q;
        msk[h] = id < total_q ? mk[id] : 0;
        if (NEO && msk[h]) trv[h] = tr4[id];
        u += du; q += dq;
        if (q >= pq) { q -= pq; u++; }
      }
#pragma unroll
      for (int h = 0; h < NB; h++) {
        if (!msk[h]) continue;
        const int id = idx + h * T;
        const float lv = lrn[uu[h]];
        const float k = (qq[h] < nq_ff ? L.learn_ff : L.learn_rec) * (NEO ? rwd[uu[h]] : lv);
        const float4 wv4 = w4[id], ev4 = e4[qq[h]];
        float wv[4] = {wv4.x, wv4.y, wv4.z, wv4.w};
        const float ev[4] = {ev4.x, ev4.y, ev4.z, ev4.w};
        float tv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (NEO) { tv[0] = trv[h].x; tv[1] = trv[h].y; tv[2] = trv[h].z; tv[3] = trv[h].w; }
#pragma unroll
        for (int e = 0; e < 4; e++)
          if ((msk[h] >> e) & 1) {
            if (NEO) {  // delta = lr*reward*trace - decay*w; w += clamp(delta); trace = lambda*trace + state*e
              const float delta = k * tv[e] - decay * wv[e];
              wv[e] = wv[e] + bidi_min(maxd, bidi_max(-maxd, delta));
              tv[e] = lambda * tv[e] + lv * ev[e];
            } else {    // delta = lr*state*e - decay*w
              const float delta = k * ev[e] - decay * wv[e];
              wv[e] = wv[e] + bidi_min(maxd, bidi_max(-maxd, delta));
            }
          }
        w4[id] = make_float4(wv[0], wv[1], wv[2], wv[3]);
        if (NEO) tr4[id] = make_float4(tv[0], tv[1], tv[2], tv[3]);
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) tma_store_1d(B + L.ff.w_off, w, row_bytes);  // waited for at the end of the kernel
  }
  // lateral weights (sdr/IRSDR.cpp:313-317): w = max(0, w + lr*(state_i*state_j - sparsity^2)), every connection, every step.
  // A layer whose lateral windows span the grid keeps one slot plane per TARGET unit, so the rule is a flat pass over
  // [slot][unit quad] with 128-bit accesses, all loads of a batch in flight before the first store; a unit owns the
  // slot of target t iff t is inside its radius and is not the unit itself
  const float sp2 = L.sparsity * L.sparsity;
  if (L.lat.r >= 0 && L.lat.absx && L.lat.absy && (hw & 3) == 0) {
    constexpr int NL = LIGHT ? 4 : 8;
    const int nq = nh >> 2, total_l = L.lat.nslots * nq, rl = L.lat.r;
    float4* wl4 = reinterpret_cast<float4*>(B + L.lat.w_off);
    const float ll = L.learn_lat;
    int s = tid / nq, q = tid - s * nq;
    const int ds = T / nq, dq = T - ds * nq;
    for (int idx = tid; idx < total_l; idx += NL * T) {
      float4 v[NL];
      int ss[NL], qq[NL];
#pragma unroll
      for (int h = 0; h < NL; h++) {
        const int id = idx + h * T;
        ss[h] = s; qq[h] = q;
        if (id < total_l) v[h] = wl4[id];
        s += ds; q += dq;
        if (q >= nq) { q -= nq; s++; }
      }
#pragma unroll
      for (int h = 0; h < NL; h++) {
        const int id = idx + h * T;
        if (id >= total_l) continue;
        const int tx = ss[h] / hh, ty = ss[h] - tx * hh, t = tx + ty * hw;
        const int u0 = 4 * qq[h], uy0 = u0 / hw, ux0 = u0 - uy0 * hw;
        const float st = lrn[t];
        const float4 su = *reinterpret_cast<const float4*>(lrn + u0);
        const bool oky = (unsigned)(ty - uy0 + rl) <= 2u * (unsigned)rl;
        float wv[4] = {v[h].x, v[h].y, v[h].z, v[h].w};
        const float sv[4] = {su.x, su.y, su.z, su.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const bool ok = oky && (unsigned)(tx - ux0 - e + rl) <= 2u * (unsigned)rl && t != u0 + e;
          if (ok) wv[e] = bidi_max(0.0f, wv[e] + ll * (sv[e] * st - sp2));
        }
        wl4[id] = make_float4(wv[0], wv[1], wv[2], wv[3]);
      }
    }
    if (warp < S.sweep_warps && i < nh) B[L.thr + i] = bidi_max(0.0f, thr_r + (lrn[i] - L.sparsity) * L.learn_threshold);
  } else if (warp < S.sweep_warps && i < nh) {
    float* __restrict__ wl = B + L.lat.w_off;
    const float si = lrn[i];
    for_each_slot_batched<LIGHT ? 8 : 16, WE>(L.lat, ux, uy,
        [&](int s, int t) { return WE{wl[(long long)s * nh + i], lrn[t]}; },
        [&](int s, const WE& v) { wl[(long long)s * nh + i] = bidi_max(0.0f, v.w + L.learn_lat * (si * v.e - sp2)); });
    B[L.thr + i] = bidi_max(0.0f, thr_r + (si - L.sparsity) * L.learn_threshold);
  }
  if (tid == 0) tma_store_commit_and_wait();  // the rows in shared memory must outlive the bulk store's reads
}

} // namespace bidi
