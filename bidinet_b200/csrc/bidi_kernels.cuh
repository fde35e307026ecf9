// bidi_kernels.cuh -- sm_100a kernels of the per-timestep hierarchy update.
//
// Mapping: one thread per (agent, unit); adjacent threads own adjacent units, so a
// warp reads one slot plane w[slot][unit..unit+31] as a single coalesced request
// and walks the slots in the reference's list order.  All floating point is
// written as plain a*b+c and compiled with -fmad=false: the reference's x86-64
// build has no FMA, and spikes / k-winner ranks are discontinuous in the last ulp
// (SURVEY section 7 "hard parts").
#pragma once
#include <cuda_runtime.h>
#include "bidi_types.h"
#include "bidi_math.h"

namespace bidi {

// ------------------------------------------------------------------ addressing
__device__ __forceinline__ float* blob_of(const EngineDev& P, int a) { return P.blob + (long long)a * P.stride; }
__device__ __forceinline__ float* vis_input_of(const EngineDev& P, const LayerDev& L, int l, int a) {
  return l == 0 ? P.in0 + (long long)a * P.n_in : blob_of(P, a) + L.vis_input;
}
// prediction-node state of node layer l (l == P.L: the input nodes, dense output array)
__device__ __forceinline__ float* p_state_of(const EngineDev& P, const LayerDev& L, int l, int a) {
  return l == P.L ? P.pred_out + (long long)a * P.n_in : blob_of(P, a) + L.p_state;
}

// Valid slot box of unit (ux,uy): the window clipped to the target grid
// (sdr/RSDR.cpp:53: vx >= 0 && vx < visibleWidth && ...).  Slots outside
// [sx0,sx1]x[sy0,sy1] correspond to connections the reference never creates.
struct SlotBox { int cx, cy, sx0, sx1, sy0, sy1; };
__device__ __forceinline__ SlotBox slot_box(const WinDev& w, int ux, int uy) {
  SlotBox b;
  b.cx = w.cx[ux]; b.cy = w.cy[uy];
  if (w.absx) { b.sx0 = max(0, b.cx - w.r); b.sx1 = min(w.tw - 1, b.cx + w.r); }
  else { b.sx0 = max(0, w.r - b.cx); b.sx1 = min(w.dxn - 1, w.tw - 1 - b.cx + w.r); }
  if (w.absy) { b.sy0 = max(0, b.cy - w.r); b.sy1 = min(w.th - 1, b.cy + w.r); }
  else { b.sy0 = max(0, w.r - b.cy); b.sy1 = min(w.dyn - 1, w.th - 1 - b.cy + w.r); }
  return b;
}
__device__ __forceinline__ int slot_tx(const WinDev& w, const SlotBox& b, int sx) { return w.absx ? sx : b.cx - w.r + sx; }
__device__ __forceinline__ int slot_ty(const WinDev& w, const SlotBox& b, int sy) { return w.absy ? sy : b.cy - w.r + sy; }
// number of list entries of the unit (clipped window, minus the excluded centre)
__device__ __forceinline__ int slot_count(const WinDev& w, const SlotBox& b) {
  if (w.r < 0) return 0;
  int n = max(0, b.sx1 - b.sx0 + 1) * max(0, b.sy1 - b.sy0 + 1);
  return (w.excl && n > 0) ? n - 1 : n;
}

// f(slot, target) for every connection of the unit, in the reference's list order.
// The loops run over the whole slot box so that a warp stays converged on one
// slot plane at a time; `valid` masks slots the unit does not own.
template <class F>
__device__ __forceinline__ void for_each_conn(const WinDev& w, int ux, int uy, F&& f) {
  if (w.r < 0) return;
  const SlotBox b = slot_box(w, ux, uy);
  for (int sx = 0; sx < w.dxn; sx++) {
    const bool okx = sx >= b.sx0 && sx <= b.sx1;
    const int tx = slot_tx(w, b, sx);
    for (int sy = 0; sy < w.dyn; sy++) {
      const int ty = slot_ty(w, b, sy);
      bool ok = okx && sy >= b.sy0 && sy <= b.sy1;
      if (w.excl && tx == b.cx && ty == b.cy) ok = false;
      if (ok) f(sx * w.dyn + sy, tx + ty * w.tw);
    }
  }
}

// Element-wise learning rules over EVERY slot of the box, NB slots at a time: all of a
// batch's loads (load(slot, clamped target) -> V) are issued before the first store
// (store(slot, V, valid)), so one memory round trip serves NB connections instead of one.
// Loads are unconditional (the target index is clamped into the grid); only the stores are
// predicated, which keeps the slots a unit does not own at 0.
template <int NB, class V, class LoadF, class StoreF>
__device__ __forceinline__ void for_each_slot_batched(const WinDev& w, int ux, int uy, LoadF&& load, StoreF&& store) {
  if (w.r < 0) return;
  const SlotBox b = slot_box(w, ux, uy);
  for (int sx = 0; sx < w.dxn; sx++) {
    const bool okx = sx >= b.sx0 && sx <= b.sx1;
    const int txr = slot_tx(w, b, sx), tx = min(max(txr, 0), w.tw - 1);
    for (int sy0 = 0; sy0 < w.dyn; sy0 += NB) {
      V v[NB];
#pragma unroll
      for (int q = 0; q < NB; q++) {
        const int sy = min(sy0 + q, w.dyn - 1);
        const int ty = min(max(slot_ty(w, b, sy), 0), w.th - 1);
        v[q] = load(sx * w.dyn + sy, tx + ty * w.tw);
      }
#pragma unroll
      for (int q = 0; q < NB; q++) {
        const int sy = sy0 + q, tyr = slot_ty(w, b, sy);
        bool ok = okx && sy < w.dyn && sy >= b.sy0 && sy <= b.sy1;
        if (w.excl && txr == b.cx && tyr == b.cy) ok = false;
        if (ok) store(sx * w.dyn + sy, v[q]);
      }
    }
  }
}
struct WE { float w, e; };
struct WTE { float w, tr, e; };

// learn_dot (below) for windows that span the target grid (w.absx && w.absy: the RunnerMain hierarchy): slot (sx, sy) IS
// target cell (sx, sy), so there are no clamps and no coordinate translation, and every address is a running pointer --
// the generic body spends ~50 instructions per connection on index arithmetic, which made k_predict / k_predict_input
// issue-bound at 1.9 TB/s.  Same loads, same operations, same order; prev / now are the source planes themselves.
// DYN: the window height when it is 4 or 8 (the RunnerMain layers: every column of slots is then one full batch with no
// remainder test), else 0
template <int NB, bool SKIP_ZERO, int DYN>
__device__ __forceinline__ float learn_dot_abs_n(const WinDev& w, int ux, int uy, float* __restrict__ plane, int U, int i, bool do_learn,
                                                 float k, float sum, const float* __restrict__ prev, const float* __restrict__ now) {
  const SlotBox b = slot_box(w, ux, uy);
  float* wc = plane + i;                       // slot (sx, 0) of this unit; a slot further down the column is U floats on
  const int dyn = DYN ? DYN : w.dyn, tw = w.tw;
  for (int sx = 0; sx < w.dxn; sx++, wc += (size_t)dyn * U) {
    const bool okx = do_learn && sx >= b.sx0 && sx <= b.sx1;
    const bool xcl = w.excl && sx == b.cx;
    const float* np = now + sx;                // target (sx, sy) = plane[sx + sy*tw]
    const float* pp = prev + sx;
    float* wq = wc;
    for (int sy0 = 0; sy0 < dyn; sy0 += NB, wq += (size_t)NB * U, np += NB * tw, pp += NB * tw) {
      float wv[NB], pv[NB], nv[NB];
      if (sy0 + NB <= dyn) {
#pragma unroll
        for (int q = 0; q < NB; q++) { wv[q] = wq[(size_t)q * U]; nv[q] = np[q * tw]; pv[q] = do_learn ? pp[q * tw] : 0.0f; }
      } else {
#pragma unroll
        for (int q = 0; q < NB; q++) {
          const int qq = min(q, dyn - 1 - sy0);
          wv[q] = wq[(size_t)qq * U]; nv[q] = np[qq * tw]; pv[q] = do_learn ? pp[qq * tw] : 0.0f;
        }
      }
#pragma unroll
      for (int q = 0; q < NB; q++) {
        const int sy = sy0 + q;
        if (sy < dyn) {
          bool ok = okx && sy >= b.sy0 && sy <= b.sy1 && !(xcl && sy == b.cy);
          if (SKIP_ZERO && pv[q] == 0.0f) ok = false;
          float wn = wv[q];
          if (ok) { wn = wv[q] + k * pv[q]; wq[(size_t)q * U] = wn; }
          sum += nv[q] * wn;
        }
      }
    }
  }
  return sum;
}
template <int NB, bool SKIP_ZERO>
__device__ __forceinline__ float learn_dot_abs(const WinDev& w, int ux, int uy, float* __restrict__ plane, int U, int i, bool do_learn,
                                               float k, float sum, const float* __restrict__ prev, const float* __restrict__ now) {
  if (w.dyn == 8) return learn_dot_abs_n<8, SKIP_ZERO, 8>(w, ux, uy, plane, U, i, do_learn, k, sum, prev, now);
  if (w.dyn == 4) return learn_dot_abs_n<4, SKIP_ZERO, 4>(w, ux, uy, plane, U, i, do_learn, k, sum, prev, now);
  return learn_dot_abs_n<NB, SKIP_ZERO, 0>(w, ux, uy, plane, U, i, do_learn, k, sum, prev, now);
}

// The delta rule w += k*prev(target) and the product sum += now(target)*w that uses the updated weight, over
// the whole slot box in ONE pass: every weight is read once, NB slots' loads in flight together.  The rule
// touches only the unit's own connections (and, with SKIP_ZERO, only those whose source is not zero: k*0 = +-0
// leaves w as it was); the sum runs over every slot like dot_conn (the others hold 0).
template <int NB, bool SKIP_ZERO, class Prev, class Now>
__device__ __forceinline__ float learn_dot(const WinDev& w, int ux, int uy, float* __restrict__ plane, long long U, int i, bool do_learn,
                                           float k, float sum, Prev&& prev, Now&& now) {
  if (w.r < 0) return sum;
  const SlotBox b = slot_box(w, ux, uy);
  float* wp = plane + i;
  for (int sx = 0; sx < w.dxn; sx++) {
    const bool okx = sx >= b.sx0 && sx <= b.sx1;
    const int txr = slot_tx(w, b, sx), tx = min(max(txr, 0), w.tw - 1);
    for (int sy0 = 0; sy0 < w.dyn; sy0 += NB) {
      float wv[NB], pv[NB], nv[NB];
#pragma unroll
      for (int q = 0; q < NB; q++) {
        const int sy = min(sy0 + q, w.dyn - 1);
        const int t = tx + min(max(slot_ty(w, b, sy), 0), w.th - 1) * w.tw;
        wv[q] = wp[(long long)(sx * w.dyn + sy) * U];
        nv[q] = now(t);
        pv[q] = do_learn ? prev(t) : 0.0f;
      }
#pragma unroll
      for (int q = 0; q < NB; q++) {
        const int sy = sy0 + q, tyr = slot_ty(w, b, sy);
        if (sy < w.dyn) {
          bool ok = do_learn && okx && sy >= b.sy0 && sy <= b.sy1;
          if (w.excl && txr == b.cx && tyr == b.cy) ok = false;
          if (SKIP_ZERO && pv[q] == 0.0f) ok = false;
          float wn = wv[q];
          if (ok) { wn = wv[q] + k * pv[q]; wp[(long long)(sx * w.dyn + sy) * U] = wn; }
          sum += nv[q] * wn;
        }
      }
    }
  }
  return sum;
}

// sum += src(target) * w[slot][unit] over the WHOLE slot box, in the reference's list
// order, with no per-slot test: the slot planes hold 0 wherever the unit has no
// connection (off-grid, outside the radius, excluded centre), the target index is
// clamped into the grid, and (+-0) added to the running sum leaves it unchanged.  The
// loop body is then a bare load-load-multiply-add the compiler can unroll and batch,
// which is what keeps a single large agent (C4) from being bound by load latency.
template <class Src>
__device__ __forceinline__ float dot_conn(const WinDev& w, int ux, int uy, const float* __restrict__ plane, long long U, int i,
                                          float sum, Src&& src) {
  if (w.r < 0) return sum;
  const int cx = w.cx[ux], cy = w.cy[uy];
  const float* wp = plane + i;
#ifdef BIDI_DOT_BATCH
  // slots in list order (sx outer, sy inner), B at a time: all 2B loads of a batch are
  // issued before the first add, so one memory round trip serves B connections
  constexpr int B = 16;
  int sx = 0, sy = 0;
  for (int s0 = 0; s0 < w.nslots; s0 += B) {
    float wv[B], xv[B];
#pragma unroll
    for (int k = 0; k < B; k++) {
      const bool in = s0 + k < w.nslots;
      const int tx = w.absx ? sx : min(max(cx - w.r + sx, 0), w.tw - 1);
      const int ty = w.absy ? sy : min(max(cy - w.r + sy, 0), w.th - 1);
      wv[k] = in ? wp[(long long)(s0 + k) * U] : 0.0f;
      xv[k] = in ? src(tx + ty * w.tw) : 0.0f;
      if (++sy == w.dyn) { sy = 0; sx++; }
    }
#pragma unroll
    for (int k = 0; k < B; k++)
      if (s0 + k < w.nslots) sum += xv[k] * wv[k];
  }
#else
  for (int sx = 0; sx < w.dxn; sx++) {
    const int tx = w.absx ? sx : min(max(cx - w.r + sx, 0), w.tw - 1);
#pragma unroll 8
    for (int sy = 0; sy < w.dyn; sy++) {
      const int ty = w.absy ? sy : min(max(cy - w.r + sy, 0), w.th - 1);
      sum += src(tx + ty * w.tw) * wp[0];
      wp += U;
    }
  }
#endif
  return sum;
}

// Gather form of the reference's scatter reconstruction (sdr/RSDR.cpp:175-187):
// sum over the units whose window contains target (tx,ty) of w[unit->target]*drive,
// visited in ascending unit index = the order the scatter adds them.  A zero drive
// contributes +-0 to a sum that starts at +0, so skipping it is exact.
__device__ __forceinline__ float gather_recon(const WinDev& w, const float* __restrict__ wp,
                                              const float* __restrict__ drive, int tx, int ty, bool scaled, float mult) {
  float sum = 0.0f;
  if (w.r < 0) return sum;
  const int y0 = w.gy_lo[ty], y1 = w.gy_hi[ty], x0 = w.gx_lo[tx], x1 = w.gx_hi[tx];
  for (int uy = y0; uy <= y1; uy++) {
    const int cy = w.cy[uy];
    const int sy = w.absy ? ty : ty - cy + w.r;
    for (int ux = x0; ux <= x1; ux++) {
      const int cx = w.cx[ux];
      if (w.excl && cx == tx && cy == ty) continue;
      const int u = ux + uy * w.uw;
      const float d = drive[u];
      if (d != 0.0f) {
        const int sx = w.absx ? tx : tx - cx + w.r;
        const float wv = wp[(long long)(sx * w.dyn + sy) * w.U + u];
        sum += scaled ? wv * d * mult : wv * d;
      }
    }
  }
  return sum;
}

// one thread per (agent, hidden unit of the launch's row strip)
#define BIDI_THREAD_AGENT_HIDDEN                                            \
  const long long gid_ = (long long)blockIdx.x * blockDim.x + threadIdx.x; \
  const int n_ = (L.row1 - L.row0) * L.hw;                                  \
  if (gid_ >= (long long)P.A * n_) return;                                  \
  const int a = (int)(gid_ / n_);                                           \
  const int i = L.row0 * L.hw + (int)(gid_ % n_);

#define BIDI_THREAD_AGENT_UNIT(N)                                        \
  const long long gid_ = (long long)blockIdx.x * blockDim.x + threadIdx.x; \
  if (gid_ >= (long long)P.A * (N)) return;                              \
  const int a = (int)(gid_ / (N));                                        \
  int i = (int)(gid_ % (N));

// ------------------------------------------------------------------ k-winners coder
// sdr/RSDR.cpp:123-133: a = -theta + sum_ff x*w + sum_rec sPrev*w
__global__ void k_kw_activate(EngineDev P, LayerDev L, int l) {
  BIDI_THREAD_AGENT_HIDDEN
  float* B = blob_of(P, a);
  const float* x = vis_input_of(P, L, l, a);
  const float* sprev = B + L.state_prev;
  const float* wff = B + L.ff.w_off;
  const float* wrec = B + L.rec.w_off;
  const int ux = i % L.hw, uy = i / L.hw, U = L.nh;
  float sum = -B[L.thr + i];
  sum = dot_conn(L.ff, ux, uy, wff, U, i, sum, [&](int t) { return x[t]; });
  sum = dot_conn(L.rec, ux, uy, wrec, U, i, sum, [&](int t) { return sprev[t]; });
  B[L.act + i] = sum;
}

// sdr/RSDR.cpp:135-145 and :148-163: rank among the lateral neighbours.  `src`
// and `dst` are blob offsets; dst_next (>=0) also receives the state as the next
// layer's visible input (sdr/PredictiveRSDR.cpp:108-111).
__global__ void k_kw_inhibit(EngineDev P, LayerDev L, int l, long long src, long long dst, long long dst_next) {
  BIDI_THREAD_AGENT_HIDDEN
  float* B = blob_of(P, a);
  const float* act = B + src;
  const int ux = i % L.hw, uy = i / L.hw;
  const float mine = act[i];
  int n = 0, cnt = 0;
  for_each_conn(L.lat, ux, uy, [&](int, int t) { n++; cnt += act[t] >= mine ? 1 : 0; });
  const float num_active = L.sparsity * (float)n;
  const float s = (float)cnt < num_active ? 1.0f : 0.0f;
  B[dst + i] = s;
  if (dst_next >= 0) B[dst_next + i] = s;
}

// Reconstruction of the visible layer and of the previous hidden state from a
// per-unit drive (state or spike): sdr/RSDR.cpp:165-194, sdr/IRSDR.cpp:218-254,
// neo/SparseCoder.cpp:242-259 (scaled: drive*multiplier).  Threads [0,nv) own a
// visible cell, [nv,nv+nh) a hidden unit.  copy_spike: also spikePrev = spike
// (sdr/IRSDR.cpp:170-171), race-free here because this kernel never reads spikePrev.
__global__ void k_reconstruct(EngineDev P, LayerDev L, int l, long long drive_off, int scaled, float mult, int copy_spike) {
  const int nvs = (L.vrow1 - L.vrow0) * L.vw, nhs = (L.row1 - L.row0) * L.hw;
  BIDI_THREAD_AGENT_UNIT(nvs + nhs)
  float* B = blob_of(P, a);
  const float* drive = B + drive_off;
  if (i < nvs) {
    const int v = L.vrow0 * L.vw + i;
    B[L.vis_recon + v] = gather_recon(L.ff, B + L.ff.w_off, drive, v % L.vw, v / L.vw, scaled != 0, mult);
  } else {
    const int h = L.row0 * L.hw + (i - nvs);
    B[L.recon + h] = gather_recon(L.rec, B + L.rec.w_off, drive, h % L.hw, h / L.hw, scaled != 0, mult);
    if (copy_spike) B[L.spike_prev + h] = B[L.spike + h];
  }
}

// sdr/RSDR.cpp:196-212 applied to layer 0's predicted states -> _prediction
// (sdr/PredictiveRSDR.cpp:188-193)
__global__ void k_kw_predict_input(EngineDev P, LayerDev L) {
  const int nvs = (L.vrow1 - L.vrow0) * L.vw;
  BIDI_THREAD_AGENT_UNIT(nvs)
  i += L.vrow0 * L.vw;
  float* B = blob_of(P, a);
  P.pred_out[(long long)a * P.n_in + i] = gather_recon(L.ff, B + L.ff.w_off, B + L.p_state, i % L.vw, i / L.vw, false, 0.0f);
}

// sdr/RSDR.cpp:241-266: winners only; theta for every unit
__global__ void k_kw_learn(EngineDev P, LayerDev L, int l) {
  BIDI_THREAD_AGENT_HIDDEN
  float* B = blob_of(P, a);
  const float* x = vis_input_of(P, L, l, a);
  const float* vrec = B + L.vis_recon;
  const float* sprev = B + L.state_prev;
  const float* hrec = B + L.recon;
  float* wff = B + L.ff.w_off;
  float* wrec = B + L.rec.w_off;
  const int ux = i % L.hw, uy = i / L.hw, U = L.nh;
  const float learn = B[L.state + i];
  if (learn > 0.0f) {
    const float att = B[L.p_mod + i];
    const float kf = L.learn_ff * att * learn, kr = L.learn_rec * att * learn;
    for_each_conn(L.ff, ux, uy, [&](int s, int t) { wff[(long long)s * U + i] += kf * (x[t] - vrec[t]); });
    for_each_conn(L.rec, ux, uy, [&](int s, int t) { wrec[(long long)s * U + i] += kr * (sprev[t] - hrec[t]); });
  }
  B[L.thr + i] += L.learn_threshold * (learn - L.sparsity);
}

// ------------------------------------------------------------------ iterative coders
// One leaky-integrate-and-fire sweep: sdr/IRSDR.cpp:138-168 / :177-209 /
// neo/SparseCoder.cpp:129-158.  first: activation and state start from 0
// (sdr/IRSDR.cpp:131-135).  acc: 0 = settle (no state change), 1 = state +=
// scale*spike (measure; scale = 1/measureIter), 2 = state += spike (neo).
__global__ void k_ic_sweep(EngineDev P, LayerDev L, int l, int first, int acc, float scale, int it) {
  BIDI_THREAD_AGENT_HIDDEN
  float* B = blob_of(P, a);
  const float* x = vis_input_of(P, L, l, a);
  const float* vrec = B + L.vis_recon;
  const float* sprev = B + L.state_prev;
  const float* hrec = B + L.recon;
  const float* spk_prev = B + L.spike_prev;
  const float* wff = B + L.ff.w_off;
  const float* wrec = B + L.rec.w_off;
  const float* wlat = B + L.lat.w_off;
  const int ux = i % L.hw, uy = i / L.hw, U = L.nh;
  // simStepGenerate: starts from noiseDist(generator) * noise (neo/SparseCoder.cpp:200)
  float excitation = P.gen_on ? P.gen_normals[(long long)a * P.gen_per_agent + L.gen_off + (long long)it * L.nh + i] * *P.gen_scale : 0.0f;
  excitation = dot_conn(L.ff, ux, uy, wff, U, i, excitation, [&](int t) { return x[t] - vrec[t]; });
  excitation = dot_conn(L.rec, ux, uy, wrec, U, i, excitation, [&](int t) { return sprev[t] - hrec[t]; });
  float inhibition = 0.0f;
  inhibition = dot_conn(L.lat, ux, uy, wlat, U, i, inhibition, [&](int t) { return spk_prev[t]; });
  float act = first ? 0.0f : B[L.act + i];
  float st = first ? 0.0f : B[L.state + i];
  act = (1.0f - L.leak) * act + excitation - inhibition;
  float spike;
  if (act > B[L.thr + i]) { act = 0.0f; spike = 1.0f; } else spike = 0.0f;
  if (acc == 1) st += scale * spike;
  else if (acc == 2) st += spike;
  B[L.act + i] = act;
  B[L.spike + i] = spike;
  B[L.state + i] = st;
}

// End of the up pass of one layer: neo: state *= 1/counter (neo/SparseCoder.cpp:171-175);
// then the next layer's input = state (x the column's _attention action for
// neo::Agent, neo/Agent.cpp:147-151).
__device__ __forceinline__ void d_ic_finish(const EngineDev& P, const LayerDev& L, int l, int scale_state, float mult, long long next_vis_input, long long gid_) {
  const int n_ = (L.row1 - L.row0) * L.hw;
  if (gid_ >= (long long)P.A * n_) return;
  const int a = (int)(gid_ / n_);
  const int i = L.row0 * L.hw + (int)(gid_ % n_);
  float* B = blob_of(P, a);
  float st = B[L.state + i];
  if (scale_state) { st *= mult; B[L.state + i] = st; }
  if (next_vis_input >= 0) {
    if (P.variant == BIDI_NEO_AGENT) st = st * B[L.col.a_expl + i * 3 + 0];
    else if (P.csrl) st = st * B[L.sd.a_expl + (long long)i * L.sd.na + 0];  // the unit's _attention action, deep/CSRL.cpp:198
    B[next_vis_input + i] = st;
  }
}
__global__ void k_ic_finish(EngineDev P, LayerDev L, int l, int scale_state, float mult, long long next_vis_input) {
  d_ic_finish(P, L, l, scale_state, mult, next_vis_input, (long long)blockIdx.x * blockDim.x + threadIdx.x);
}

// lateral anti-Hebbian + threshold homeostasis: sdr/IRSDR.cpp:313-317
__device__ __forceinline__ void ic_learn_tail(const LayerDev& L, float* B, int i, int ux, int uy) {
  const float* __restrict__ state = B + L.state;
  float* __restrict__ wlat = B + L.lat.w_off;
  const float si = state[i], sp2 = L.sparsity * L.sparsity;
  const int U = L.nh;
  for_each_slot_batched<8, WE>(L.lat, ux, uy,
      [&](int s, int t) { return WE{wlat[(long long)s * U + i], state[t]}; },
      [&](int s, const WE& v) { wlat[(long long)s * U + i] = bidi_max(0.0f, v.w + L.learn_lat * (si * v.e - sp2)); });
  B[L.thr + i] = bidi_max(0.0f, B[L.thr + i] + (si - L.sparsity) * L.learn_threshold);
}

// sdr/IRSDR.cpp:287-319 (Hebbian, clamped delta)
__global__ void k_ic_learn(EngineDev P, LayerDev L, int l) {
  BIDI_THREAD_AGENT_HIDDEN
  float* B = blob_of(P, a);
  const float* __restrict__ x = vis_input_of(P, L, l, a);
  const float* __restrict__ vrec = B + L.vis_recon;
  const float* __restrict__ sprev = B + L.state_prev;
  const float* __restrict__ hrec = B + L.recon;
  float* __restrict__ wff = B + L.ff.w_off;
  float* __restrict__ wrec = B + L.rec.w_off;
  const int ux = i % L.hw, uy = i / L.hw, U = L.nh;
  const float learn = B[L.state + i], maxd = L.max_weight_delta, decay = L.weight_decay;
  for_each_slot_batched<8, WE>(L.ff, ux, uy,
      [&](int s, int t) { return WE{wff[(long long)s * U + i], x[t] - vrec[t]}; },
      [&](int s, const WE& v) {
        const float delta = L.learn_ff * learn * v.e - decay * v.w;
        wff[(long long)s * U + i] = v.w + bidi_min(maxd, bidi_max(-maxd, delta));
      });
  for_each_slot_batched<8, WE>(L.rec, ux, uy,
      [&](int s, int t) { return WE{wrec[(long long)s * U + i], sprev[t] - hrec[t]}; },
      [&](int s, const WE& v) {
        const float delta = L.learn_rec * learn * v.e - decay * v.w;
        wrec[(long long)s * U + i] = v.w + bidi_min(maxd, bidi_max(-maxd, delta));
      });
  ic_learn_tail(L, B, i, ux, uy);
}

// neo/Agent.cpp:252-265 (reward from the prediction error) followed by
// neo/SparseCoder.cpp:326-361 (reward-modulated update through eligibility traces)
__global__ void k_neo_learn(EngineDev P, LayerDev L, int l) {
  BIDI_THREAD_AGENT_HIDDEN
  float* B = blob_of(P, a);
  const float* __restrict__ x = vis_input_of(P, L, l, a);
  const float* __restrict__ vrec = B + L.vis_recon;
  const float* __restrict__ sprev = B + L.state_prev;
  const float* __restrict__ hrec = B + L.recon;
  float* __restrict__ wff = B + L.ff.w_off; float* __restrict__ tff = B + L.ff.tr_off;
  float* __restrict__ wrec = B + L.rec.w_off; float* __restrict__ trec = B + L.rec.tr_off;
  const int ux = i % L.hw, uy = i / L.hw, U = L.nh;
  const float learn = B[L.state + i], maxd = L.max_weight_delta, decay = L.weight_decay;
  float reward;
  if (P.csrl) reward = B[L.p_mod + i];  // deep::CSRL: the node's _reward action (deep/CSRL.cpp:278-281,421-424); rule sdr/IRSDR.cpp:321-356
  else {
    const float err = learn - B[L.p_state_prev + i];
    const float error2 = err * err;
    const float base = B[L.p_baseline + i];
    reward = bidi_sigmoid(L.sensitivity * (error2 - base));
    B[L.p_baseline + i] = (1.0f - L.baseline_decay) * base + L.baseline_decay * error2;
  }
  for_each_slot_batched<8, WTE>(L.ff, ux, uy,
      [&](int s, int t) { const long long k = (long long)s * U + i; return WTE{wff[k], tff[k], x[t] - vrec[t]}; },
      [&](int s, const WTE& v) {
        const long long k = (long long)s * U + i;
        const float delta = L.learn_ff * reward * v.tr - decay * v.w;
        wff[k] = v.w + bidi_min(maxd, bidi_max(-maxd, delta));
        tff[k] = L.lambda * v.tr + learn * v.e;
      });
  for_each_slot_batched<8, WTE>(L.rec, ux, uy,
      [&](int s, int t) { const long long k = (long long)s * U + i; return WTE{wrec[k], trec[k], sprev[t] - hrec[t]}; },
      [&](int s, const WTE& v) {
        const long long k = (long long)s * U + i;
        const float delta = L.learn_rec * reward * v.tr - decay * v.w;
        wrec[k] = v.w + bidi_min(maxd, bidi_max(-maxd, delta));
        trec[k] = L.lambda * v.tr + learn * v.e;
      });
  ic_learn_tail(L, B, i, ux, uy);
}

// ------------------------------------------------------------------ prediction nodes
// Down pass of node layer l < L: sdr/PredictiveRSDR.cpp:122-160,
// sdr/IPredictiveRSDR.cpp:156-194, neo/PredictiveHierarchy.cpp:151-181,
// neo/Agent.cpp:177-208.  The delta rule runs BEFORE the activation that uses the
// weights.
__global__ void k_predict(EngineDev P, LayerDev L, int l, int learn, long long up_p_state, long long up_p_state_prev) {
  BIDI_THREAD_AGENT_HIDDEN
  float* B = blob_of(P, a);
  const bool top = l == P.L - 1;
  const float* up_state = B + up_p_state;
  const float* up_prev = B + up_p_state_prev;
  const float* state = B + L.state;
  const float* sprev = B + L.state_prev;
  float* wfb = B + L.fb.w_off;
  float* wpr = B + L.pred.w_off;
  const int ux = i % L.hw, uy = i / L.hw, U = L.pn;
  float kfb = 0.0f, kpr = 0.0f;
  bool upd = false;
  if (learn) {
    float err = state[i] - B[L.p_state_prev + i];
    if (P.variant == BIDI_KWINNERS) {
      const float surprise = err * err, avg = B[L.p_baseline + i];
      B[L.p_mod + i] = bidi_sigmoid(L.attention_factor * (surprise - avg));
      B[L.p_baseline + i] = (1.0f - L.avg_surprise_decay) * avg + L.avg_surprise_decay * surprise;
    } else if (P.variant == BIDI_ITERATIVE) {
      const float error2 = err * err, base = B[L.p_baseline + i];
      B[L.p_baseline + i] = (1.0f - L.baseline_decay) * base + L.baseline_decay * error2;
    } else if (P.variant == BIDI_NEO_AGENT) {
      err = B[L.col.a_expl + i * 3 + 1] * err; // _learnPrediction gate, neo/Agent.cpp:181
    }
    kfb = L.learn_fb * err; kpr = L.learn_pred * err;
    // w += k*src: a zero error or a zero source adds (+-0), which leaves w as it was, so those
    // weights are not rewritten (binary, sparse states: almost all of them)
    upd = err != 0.0f;
  }
  float activation = 0.0f;
  if (!top && L.fb.r >= 0) {
    if (L.fb.absx && L.fb.absy) activation = learn_dot_abs<8, true>(L.fb, ux, uy, wfb, U, i, upd, kfb, activation, up_prev, up_state);
    else activation = learn_dot<8, true>(L.fb, ux, uy, wfb, U, i, upd, kfb, activation, [&](int t) { return up_prev[t]; }, [&](int t) { return up_state[t]; });
  }
  if (L.pred.r >= 0) {
    if (L.pred.absx && L.pred.absy) activation = learn_dot_abs<8, true>(L.pred, ux, uy, wpr, U, i, upd, kpr, activation, sprev, state);
    else activation = learn_dot<8, true>(L.pred, ux, uy, wpr, U, i, upd, kpr, activation, [&](int t) { return sprev[t]; }, [&](int t) { return state[t]; });
  }
  B[L.p_act + i] = activation;
  if (P.variant != BIDI_KWINNERS) B[L.p_state + i] = bidi_clamp01(activation);
}

// Input prediction nodes: sdr/IPredictiveRSDR.cpp:198-218, neo/Agent.cpp:232-246
__global__ void k_predict_input(EngineDev P, LayerDev L, int learn, long long p0_state, long long p0_state_prev) {
  BIDI_THREAD_AGENT_UNIT(L.pn)
  float* B = blob_of(P, a);
  const float* p0 = B + p0_state;
  const float* p0_prev = B + p0_state_prev;
  float* wfb = B + L.fb.w_off;
  const int ux = i % L.fb.uw, uy = i / L.fb.uw, U = L.pn;
  float k = 0.0f;
  if (learn) {
    float err = P.in0[(long long)a * P.n_in + i] - B[L.p_state_prev + i];
    if (P.variant == BIDI_NEO_AGENT) err = B[L.col.a_expl + i * 3 + 1] * err;
    k = L.learn_fb * err; // learn_fb of the input layer = _learnInputFeedBack
  }
  float activation = 0.0f;
  if (L.fb.r >= 0 && L.fb.absx && L.fb.absy) activation = learn_dot_abs<8, false>(L.fb, ux, uy, wfb, U, i, learn != 0, k, activation, p0_prev, p0);
  else activation = learn_dot<8, false>(L.fb, ux, uy, wfb, U, i, learn != 0, k, activation, [&](int t) { return p0_prev[t]; }, [&](int t) { return p0[t]; });
  B[L.p_act + i] = activation;
  P.pred_out[(long long)a * P.n_in + i] = activation;
}

// stepEnd of every layer + the *_Prev copies: sdr/PredictiveRSDR.cpp:177-184,
// sdr/IPredictiveRSDR.cpp:224-239
__device__ __forceinline__ void d_step_end(const EngineDev& P, int max_units, long long gid_) {
  if (gid_ >= (long long)P.A * max_units) return;
  const int a = (int)(gid_ / max_units);
  const int i = (int)(gid_ % max_units);
  float* B = blob_of(P, a);
  const int nl = P.variant == BIDI_KWINNERS ? P.L : P.L + 1;
  for (int l = 0; l < nl; l++) {
    const LayerDev& L = P.layers[l];
    if (l < P.L && i < L.nh) B[L.state_prev + i] = B[L.state + i];
    if (i < L.pn) {
      B[L.p_state_prev + i] = p_state_of(P, L, l, a)[i];
      B[L.p_act_prev + i] = B[L.p_act + i];
    }
  }
  // deep::CSRL: every input cell takes its own prediction (the caller then overwrites the state cells), deep/CSRL.cpp:445
  if (P.csrl && i < P.n_in) P.in0[(long long)a * P.n_in + i] = P.pred_out[(long long)a * P.n_in + i];
}
__global__ void k_step_end(EngineDev P, int max_units) { d_step_end(P, max_units, (long long)blockIdx.x * blockDim.x + threadIdx.x); }

// The same for one row strip of a split layer (bidi_strip.inl): the winners are kept on
// hidden rows [s0,s1) (owned rows plus the halo later phases read), the prediction
// nodes only on the owned rows [L.row0,L.row1).
__global__ void k_strip_step_end(EngineDev P, LayerDev L, int s0, int s1, unsigned* epoch) {
  const int n = (s1 - s0) * L.hw;
  if (epoch && blockIdx.x == 0 && threadIdx.x == 0) *epoch += 1u;  // the step's exchanges (k_strip_push) are behind us
  BIDI_THREAD_AGENT_UNIT(n)
  i += s0 * L.hw;
  float* B = blob_of(P, a);
  B[L.state_prev + i] = B[L.state + i];
  const int row = i / L.hw;
  if (row >= L.row0 && row < L.row1) {
    B[L.p_state_prev + i] = B[L.p_state + i];
    B[L.p_act_prev + i] = B[L.p_act + i];
  }
}

// Halo exchange of a split layer over peer-mapped memory (NVLink): ONE small kernel per
// exchange point pushes the rows the neighbours need straight into their blobs (plain
// coalesced stores to peer memory), publishes an epoch number in each neighbour's flag
// word, and waits for the neighbours' flags -- transfer, signal and wait in one launch,
// no library call, capturable in the step's CUDA graph.  One CTA: the halos are a few
// rows (r x width x 4 B = 24.6 kB per direction for C5).
struct StripXferDev { int peer; int pad; long long off, count; };
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void __launch_bounds__(1024) k_strip_push(EngineDev P, const StripXferDev* __restrict__ sends, int n_sends, float* const* __restrict__ peer_blob,
                                                     unsigned* const* __restrict__ peer_flags, unsigned* my_flags, const unsigned* epoch_ctr, int point,
                                                     int me, int n_parts, const int* __restrict__ notify, int n_notify, const int* __restrict__ await, int n_await) {
  const unsigned epoch = *epoch_ctr * 8u + (unsigned)point + 1u;
  for (int k = 0; k < n_sends; k++) {
    const StripXferDev x = sends[k];
    for (int a = 0; a < P.A; a++) {
      const float* src = P.blob + (long long)a * P.stride + x.off;
      float* dst = peer_blob[x.peer] + (long long)a * P.stride + x.off;
      if (((x.off | x.count | P.stride) & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (long long j = threadIdx.x; j < (x.count >> 2); j += blockDim.x) d4[j] = s4[j];
      } else {
        for (long long j = threadIdx.x; j < x.count; j += blockDim.x) dst[j] = src[j];
      }
    }
  }
  __threadfence_system();  // this thread's stores to peer memory are ordered before the flag
  __syncthreads();
  if ((int)threadIdx.x < n_notify) st_release_sys(peer_flags[notify[threadIdx.x]] + point * n_parts + me, epoch);
  if ((int)threadIdx.x < n_await) {
    const unsigned* f = my_flags + point * n_parts + await[threadIdx.x];
    while ((int)(ld_acquire_sys(f) - epoch) < 0) {}
  }
  __syncthreads();
}

// ------------------------------------------------------------------ neo::Column
// Exploration noise when the caller supplies none: a counter-based generator keyed
// by (seed, step, agent, column, action); u uniform, v uniform or N(0,std).
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(unsigned long long h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }

// `first`: global index of this engine's agent 0 (bidi_set_noise_stream), so that shards of one population never share a stream
__global__ void k_noise(EngineDev P, float* u, float* v, unsigned long long seed, unsigned long long step, long long first) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)P.A * P.ncol;
  if (gid >= n) return;
  const unsigned long long key0 = (unsigned long long)(gid + first * P.ncol) * 3;
  const int c = (int)(gid % P.ncol);
  float std = 0.0f, brk = 0.0f;
  for (int l = 0; l <= P.L; l++) {
    const ColumnsDev& C = P.layers[l].col;
    if (c >= C.noise_base && c < C.noise_base + C.n) { std = C.expl_std; brk = C.expl_break; }
  }
  for (int k = 0; k < 3; k++) {
    unsigned long long h = mix64(seed ^ mix64(step * 0x100000001B3ULL + key0 + k));
    const float uu = u01(h);
    h = mix64(h);
    float vv;
    if (uu < brk) vv = u01(h);
    else {
      const float r1 = fmaxf(u01(h), 5.9604645e-8f), r2 = u01(mix64(h));
      vv = sqrtf(-2.0f * logf(r1)) * cospif(2.0f * r2) * std;
    }
    u[gid * 3 + k] = uu;
    v[gid * 3 + k] = vv;
  }
}

// N(0,1) draws for simStepGenerate when the caller supplies none (counter-based, Box-Muller)
__global__ void k_gen_normals(float* out, long long n, unsigned long long seed, unsigned long long step, long long first_key) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n) return;
  const unsigned long long h = mix64(seed ^ mix64(step * 0x100000001B3ULL + (unsigned long long)(gid + first_key)));
  const float r1 = fmaxf(u01(h), 5.9604645e-8f), r2 = u01(mix64(h));
  out[gid] = sqrtf(-2.0f * logf(r1)) * cospif(2.0f * r2);
}

// The caller's side of the RunnerMain loop, for device-resident runs: input[dst] =
// clamp01(prediction[src]*scale + bias + N(0,std)) (RunnerMain.cpp:241-242) and,
// if asked, reward = mean(clamp01(prediction[src]*scale + bias)) - 0.5, the
// headless stand-in for the Box2D body velocity.
__global__ void k_feedback(EngineDev P, int n, const int* __restrict__ src, const int* __restrict__ dst, float scale,
                           float bias, float std, unsigned long long seed, unsigned long long step, long long first, float* rewards) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= P.A) return;
  const float* pred = P.pred_out + (long long)a * P.n_in;
  float* in = P.in0 + (long long)a * P.n_in;
  float sum = 0.0f;
  for (int k = 0; k < n; k++) {
    const float act = pred[src[k]] * scale + bias;
    sum += bidi_clamp01(act);
    unsigned long long h = mix64(seed ^ mix64(step * 0x9E3779B1ULL + (unsigned long long)(a + first) * n + k));
    const float r1 = fmaxf(u01(h), 5.9604645e-8f), r2 = u01(mix64(h));
    in[dst[k]] = bidi_clamp01(act + sqrtf(-2.0f * logf(r1)) * cospif(2.0f * r2) * std);
  }
  if (rewards) rewards[a] = sum / (float)n - 0.5f;
}

} // namespace bidi

#include "bidi_columns.cuh"
#include "bidi_columns16.cuh"
