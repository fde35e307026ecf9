// bidi_tiles.cuh -- the per-phase kernels for SMALL launches (one large agent: C1, C4).
//
// With one thread per unit a single agent gives the GPU a few hundred warps, each
// walking 100-500 connections one memory round trip after another: pure latency.
// Here a CTA owns a tile of 32 units and 8 warps share each unit's window: warp g takes
// the slots s = g, g+8, ... and lane u the unit, so every load is a coalesced 128-byte
// row of a slot plane and all of a thread's loads are independent.  The PRODUCTS (or
// updated weights) are formed in parallel and parked in shared memory as P[slot][unit];
// then one warp adds them up slot by slot -- the reference's list order, which is the
// only part of the computation whose order matters for bit-exact results.
// Element-wise learning rules need no order and simply run on the (unit, slot-group)
// threads.  Same arithmetic as bidi_kernels.cuh, so both paths give identical bits.
#pragma once
#include "bidi_kernels.cuh"

namespace bidi {

#define TILE_U 32
#define TILE_G 8

struct TileId { int a, u0, ul, g, unit; bool ok; };
__device__ __forceinline__ TileId tile_id(int n_units, int bid) {
  TileId t;
  const int tiles = (n_units + TILE_U - 1) / TILE_U;
  t.a = bid / tiles;
  t.u0 = (bid % tiles) * TILE_U;
  t.ul = threadIdx.x & 31;
  t.g = threadIdx.x >> 5;
  t.unit = t.u0 + t.ul;
  t.ok = t.unit < n_units;
  return t;
}

// P[(base + s)*32 + ul] = src(target of slot s) * w[s][unit] for this thread's slots.
// f(s, t, valid, w_ref) may first update the weight (delta rule) -- the thread owns it.
// do_upd: first w += k * usrc(target) for the connections the unit owns (the delta rule; the thread owns the weight).
// The slots are taken NB at a time: every load of a batch -- weight, source, delta-rule source -- is issued before
// the first product, so a batch costs one memory round trip (one agent's step is a chain of such round trips;
// one load at a time made a C4 sweep tile 23 us of pure latency).
template <class USrc, class Src>
__device__ __forceinline__ void tile_products(const WinDev& w, const TileId& T, int uw, float* __restrict__ plane, long long U,
                                              float* __restrict__ Pbuf, int base, bool do_upd, float k, USrc&& usrc, Src&& src) {
  if (w.r < 0) return;
  const int ux = T.ok ? T.unit % uw : 0, uy = T.ok ? T.unit / uw : 0;
  const SlotBox b = slot_box(w, ux, uy);
  int sx = T.g / w.dyn, sy = T.g - sx * w.dyn;
  constexpr int NB = 8;
  for (int s0 = T.g; s0 < w.nslots; s0 += TILE_G * NB) {
    float wv[NB], xv[NB], uv[NB];
    bool valid[NB];
#pragma unroll
    for (int q = 0; q < NB; q++) {
      const int s = s0 + q * TILE_G;
      const int txr = slot_tx(w, b, sx), tyr = slot_ty(w, b, sy);
      valid[q] = T.ok && s < w.nslots && sx >= b.sx0 && sx <= b.sx1 && sy >= b.sy0 && sy <= b.sy1 && !(w.excl && txr == b.cx && tyr == b.cy);
      const int t = min(max(txr, 0), w.tw - 1) + min(max(tyr, 0), w.th - 1) * w.tw;
      wv[q] = 0.0f; xv[q] = 0.0f; uv[q] = 0.0f;
      if (T.ok && s < w.nslots) {
        wv[q] = plane[(long long)s * U + T.unit];
        xv[q] = src(t);
        if (do_upd) uv[q] = usrc(t);
      }
      sy += TILE_G;
      while (sy >= w.dyn) { sy -= w.dyn; sx++; }
    }
#pragma unroll
    for (int q = 0; q < NB; q++) {
      const int s = s0 + q * TILE_G;
      if (s < w.nslots) {
        if (do_upd && valid[q]) { wv[q] += k * uv[q]; plane[(long long)s * U + T.unit] = wv[q]; }
        Pbuf[(base + s) * TILE_U + T.ul] = xv[q] * wv[q];  // slots the unit does not own hold 0: the product is +-0
      }
    }
  }
}
struct NoSrc { __device__ __forceinline__ float operator()(int) const { return 0.0f; } };

// f(slot, target) for this thread's share of the unit's connections (element-wise rules)
template <class F>
__device__ __forceinline__ void tile_for_each(const WinDev& w, const TileId& T, int uw, F&& f) {
  if (w.r < 0 || !T.ok) return;
  const int ux = T.unit % uw, uy = T.unit / uw;
  const SlotBox b = slot_box(w, ux, uy);
  int sx = T.g / w.dyn, sy = T.g - sx * w.dyn;
  for (int s = T.g; s < w.nslots; s += TILE_G) {
    const int tx = slot_tx(w, b, sx), ty = slot_ty(w, b, sy);
    if (sx >= b.sx0 && sx <= b.sx1 && sy >= b.sy0 && sy <= b.sy1 && !(w.excl && tx == b.cx && ty == b.cy)) f(s, tx + ty * w.tw);
    sy += TILE_G;
    while (sy >= w.dyn) { sy -= w.dyn; sx++; }
  }
}

__device__ __forceinline__ float tile_sum(const float* __restrict__ Pbuf, int base, int n, int ul, float sum) {
#pragma unroll 8
  for (int s = 0; s < n; s++) sum += Pbuf[(base + s) * TILE_U + ul];
  return sum;
}

// ---- sdr/RSDR.cpp:123-133
__device__ __forceinline__ void t_kw_activate(const EngineDev& P, const LayerDev& L, int l, int bid) {
  extern __shared__ float Pbuf[];
  const TileId T = tile_id(L.nh, bid);
  float* B = blob_of(P, T.a);
  const float* x = vis_input_of(P, L, l, T.a);
  const float* sprev = B + L.state_prev;
  const int nff = L.ff.nslots, nrec = L.rec.r < 0 ? 0 : L.rec.nslots;
  tile_products(L.ff, T, L.hw, B + L.ff.w_off, L.nh, Pbuf, 0, false, 0.0f, NoSrc(), [&](int t) { return x[t]; });
  tile_products(L.rec, T, L.hw, B + L.rec.w_off, L.nh, Pbuf, nff, false, 0.0f, NoSrc(), [&](int t) { return sprev[t]; });
  __syncthreads();
  if (T.g == 0 && T.ok) B[L.act + T.unit] = tile_sum(Pbuf, 0, nff + nrec, T.ul, -B[L.thr + T.unit]);
}

// ---- sdr/RSDR.cpp:135-163
__device__ __forceinline__ void t_kw_inhibit(const EngineDev& P, const LayerDev& L, int l, long long src, long long dst, long long dst_next, int bid) {
  __shared__ int s_n[TILE_G][TILE_U], s_c[TILE_G][TILE_U];
  const TileId T = tile_id(L.nh, bid);
  float* B = blob_of(P, T.a);
  const float* act = B + src;
  const float mine = T.ok ? act[T.unit] : 0.0f;
  int n = 0, cnt = 0;
  tile_for_each(L.lat, T, L.hw, [&](int, int t) { n++; cnt += act[t] >= mine ? 1 : 0; });
  s_n[T.g][T.ul] = n; s_c[T.g][T.ul] = cnt;
  __syncthreads();
  if (T.g == 0 && T.ok) {
    n = 0; cnt = 0;
    for (int g = 0; g < TILE_G; g++) { n += s_n[g][T.ul]; cnt += s_c[g][T.ul]; }
    const float s = (float)cnt < L.sparsity * (float)n ? 1.0f : 0.0f;
    B[dst + T.unit] = s;
    if (dst_next >= 0) B[dst_next + T.unit] = s;
  }
}

// ---- reconstruction (gather): sdr/RSDR.cpp:165-212, sdr/IRSDR.cpp:218-254, neo/SparseCoder.cpp:242-259
// mode 0: tiles [0,tv) reconstruct the visible layer, the rest the hidden one; mode 1:
// _prediction = reconstructFeedForward(layer-0 predicted states) (sdr/PredictiveRSDR.cpp:188-193)
__device__ __forceinline__ void t_reconstruct(const EngineDev& P, const LayerDev& L, int l, long long drive_off, int scaled, float mult, int copy_spike, int mode, int bid) {
  extern __shared__ float Pbuf[];
  const int tv = (L.nv + TILE_U - 1) / TILE_U, th = mode ? 0 : (L.nh + TILE_U - 1) / TILE_U;
  const int a = bid / (tv + th), tile = bid % (tv + th);
  const bool vis = tile < tv;
  const int ul = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int n_targets = vis ? L.nv : L.nh, target = (vis ? tile : tile - tv) * TILE_U + ul;
  const bool ok = target < n_targets;
  const WinDev& w = vis ? L.ff : L.rec;
  float* B = blob_of(P, a);
  const float* drive = B + drive_off;
  const float* wp = B + w.w_off;
  int ncand = 0;
  if (ok && w.r >= 0) {
    const int tx = target % w.tw, ty = target / w.tw;
    const int x0 = w.gx_lo[tx], x1 = w.gx_hi[tx], y0 = w.gy_lo[ty], y1 = w.gy_hi[ty];
    const int nx = max(0, x1 - x0 + 1), ny = max(0, y1 - y0 + 1);
    ncand = nx * ny;
    // candidates NB at a time: first all their drives, then the weights of those that drive, then the products
    constexpr int NB = 8;
    for (int c0 = g; c0 < ncand; c0 += TILE_G * NB) {
      float dv[NB], wv[NB];
      long long wk[NB];
#pragma unroll
      for (int q = 0; q < NB; q++) {
        const int c = c0 + q * TILE_G;
        dv[q] = 0.0f; wk[q] = -1;
        if (c < ncand) {
          const int uy = y0 + c / nx, ux = x0 + c % nx;
          const int cx = w.cx[ux], cy = w.cy[uy], u = ux + uy * w.uw;
          if (!(w.excl && cx == tx && cy == ty)) {
            const int sx = w.absx ? tx : tx - cx + w.r, sy = w.absy ? ty : ty - cy + w.r;
            dv[q] = drive[u];
            wk[q] = (long long)(sx * w.dyn + sy) * w.U + u;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < NB; q++) wv[q] = (dv[q] != 0.0f && wk[q] >= 0) ? wp[wk[q]] : 0.0f;
#pragma unroll
      for (int q = 0; q < NB; q++) {
        const int c = c0 + q * TILE_G;
        if (c < ncand) Pbuf[c * TILE_U + ul] = dv[q] != 0.0f ? (scaled ? wv[q] * dv[q] * mult : wv[q] * dv[q]) : 0.0f;
      }
    }
  }
  __syncthreads();
  if (g == 0 && ok) {
    const float sum = tile_sum(Pbuf, 0, ncand, ul, 0.0f);
    if (mode) P.pred_out[(long long)a * P.n_in + target] = sum;
    else if (vis) B[L.vis_recon + target] = sum;
    else {
      B[L.recon + target] = sum;
      if (copy_spike) B[L.spike_prev + target] = B[L.spike + target];
    }
  }
}

// ---- sdr/RSDR.cpp:241-266 (winners only)
__device__ __forceinline__ void t_kw_learn(const EngineDev& P, const LayerDev& L, int l, int bid) {
  const TileId T = tile_id(L.nh, bid);
  float* B = blob_of(P, T.a);
  if (!T.ok) return;
  const float* x = vis_input_of(P, L, l, T.a);
  const float* vrec = B + L.vis_recon;
  const float* sprev = B + L.state_prev;
  const float* hrec = B + L.recon;
  float* wff = B + L.ff.w_off;
  float* wrec = B + L.rec.w_off;
  const long long U = L.nh;
  const float learn = B[L.state + T.unit];
  if (learn > 0.0f) {
    const float att = B[L.p_mod + T.unit];
    const float kf = L.learn_ff * att * learn, kr = L.learn_rec * att * learn;
    tile_for_each(L.ff, T, L.hw, [&](int s, int t) { wff[s * U + T.unit] += kf * (x[t] - vrec[t]); });
    tile_for_each(L.rec, T, L.hw, [&](int s, int t) { wrec[s * U + T.unit] += kr * (sprev[t] - hrec[t]); });
  }
  if (T.g == 0) B[L.thr + T.unit] += L.learn_threshold * (learn - L.sparsity);
}

// ---- one sweep of the iterative coders, per phase: sdr/IRSDR.cpp:138-168
__device__ __forceinline__ void t_ic_sweep(const EngineDev& P, const LayerDev& L, int l, int first, int acc, float scale, int it, int bid) {
  extern __shared__ float Pbuf[];
  const TileId T = tile_id(L.nh, bid);
  float* B = blob_of(P, T.a);
  const float* x = vis_input_of(P, L, l, T.a);
  const float* vrec = B + L.vis_recon;
  const float* sprev = B + L.state_prev;
  const float* hrec = B + L.recon;
  const float* spk_prev = B + L.spike_prev;
  const int nff = L.ff.nslots, nrec = L.rec.r < 0 ? 0 : L.rec.nslots, nlat = L.lat.nslots;
  tile_products(L.ff, T, L.hw, B + L.ff.w_off, L.nh, Pbuf, 0, false, 0.0f, NoSrc(), [&](int t) { return x[t] - vrec[t]; });
  tile_products(L.rec, T, L.hw, B + L.rec.w_off, L.nh, Pbuf, nff, false, 0.0f, NoSrc(), [&](int t) { return sprev[t] - hrec[t]; });
  tile_products(L.lat, T, L.hw, B + L.lat.w_off, L.nh, Pbuf, nff + nrec, false, 0.0f, NoSrc(), [&](int t) { return spk_prev[t]; });
  __syncthreads();
  if (T.g == 0 && T.ok) {
    const int i = T.unit;
    const float start = P.gen_on ? P.gen_normals[(long long)T.a * P.gen_per_agent + L.gen_off + (long long)it * L.nh + i] * *P.gen_scale : 0.0f;
    const float excitation = tile_sum(Pbuf, 0, nff + nrec, T.ul, start);
    const float inhibition = tile_sum(Pbuf, nff + nrec, nlat, T.ul, 0.0f);
    float av = first ? 0.0f : B[L.act + i];
    float st = first ? 0.0f : B[L.state + i];
    av = (1.0f - L.leak) * av + excitation - inhibition;
    float spike;
    if (av > B[L.thr + i]) { av = 0.0f; spike = 1.0f; } else spike = 0.0f;
    if (acc == 1) st += scale * spike;
    else if (acc == 2) st += spike;
    B[L.act + i] = av;
    B[L.spike + i] = spike;
    B[L.state + i] = st;
  }
}

// ---- coder learning of the iterative variants: sdr/IRSDR.cpp:287-319, neo/SparseCoder.cpp:326-361
// (+ the reward of neo/Agent.cpp:252-265).  traced = 0: Hebbian; 1: eligibility traces.
__device__ __forceinline__ void t_ic_learn(const EngineDev& P, const LayerDev& L, int l, int traced, int bid) {
  const TileId T = tile_id(L.nh, bid);
  float* B = blob_of(P, T.a);
  const float* x = vis_input_of(P, L, l, T.a);
  const float* vrec = B + L.vis_recon;
  const float* sprev = B + L.state_prev;
  const float* hrec = B + L.recon;
  const float* state = B + L.state;
  float* wff = B + L.ff.w_off; float* wrec = B + L.rec.w_off; float* wlat = B + L.lat.w_off;
  const long long U = L.nh;
  const int i = T.ok ? T.unit : 0;
  const float learn = state[i], maxd = L.max_weight_delta, decay = L.weight_decay;
  float reward = 0.0f, base = 0.0f, error2 = 0.0f;
  if (traced) {
    const float err = learn - B[L.p_state_prev + i];
    error2 = err * err;
    base = B[L.p_baseline + i];
    reward = bidi_sigmoid(L.sensitivity * (error2 - base));
  }
  __syncthreads();  // every thread of the unit has read the old baseline
  if (!T.ok) return;
  if (traced) {
    float* tff = B + L.ff.tr_off; float* trec = B + L.rec.tr_off;
    if (T.g == 0) B[L.p_baseline + i] = (1.0f - L.baseline_decay) * base + L.baseline_decay * error2;
    tile_for_each(L.ff, T, L.hw, [&](int s, int t) {
      const long long k = s * U + i;
      const float tr = tff[k];
      const float delta = L.learn_ff * reward * tr - decay * wff[k];
      wff[k] += bidi_min(maxd, bidi_max(-maxd, delta));
      tff[k] = L.lambda * tr + learn * (x[t] - vrec[t]);
    });
    tile_for_each(L.rec, T, L.hw, [&](int s, int t) {
      const long long k = s * U + i;
      const float tr = trec[k];
      const float delta = L.learn_rec * reward * tr - decay * wrec[k];
      wrec[k] += bidi_min(maxd, bidi_max(-maxd, delta));
      trec[k] = L.lambda * tr + learn * (sprev[t] - hrec[t]);
    });
  } else {
    tile_for_each(L.ff, T, L.hw, [&](int s, int t) {
      const long long k = s * U + i;
      const float delta = L.learn_ff * learn * (x[t] - vrec[t]) - decay * wff[k];
      wff[k] += bidi_min(maxd, bidi_max(-maxd, delta));
    });
    tile_for_each(L.rec, T, L.hw, [&](int s, int t) {
      const long long k = s * U + i;
      const float delta = L.learn_rec * learn * (sprev[t] - hrec[t]) - decay * wrec[k];
      wrec[k] += bidi_min(maxd, bidi_max(-maxd, delta));
    });
  }
  const float sp2 = L.sparsity * L.sparsity;
  tile_for_each(L.lat, T, L.hw, [&](int s, int t) {
    const long long k = s * U + i;
    wlat[k] = bidi_max(0.0f, wlat[k] + L.learn_lat * (learn * state[t] - sp2));
  });
  if (T.g == 0) B[L.thr + i] = bidi_max(0.0f, B[L.thr + i] + (learn - L.sparsity) * L.learn_threshold);
}

// ---- down pass of node layer l (all variants but neo::Agent, whose columns come first):
// sdr/PredictiveRSDR.cpp:122-160, sdr/IPredictiveRSDR.cpp:156-194.  l == P.L: the input nodes.
__device__ __forceinline__ void t_predict(const EngineDev& P, const LayerDev& L, int l, int learn, long long up_p_state, long long up_p_state_prev, int bid) {
  extern __shared__ float Pbuf[];
  const TileId T = tile_id(L.pn, bid);
  float* B = blob_of(P, T.a);
  const bool input_layer = l == P.L, top = l == P.L - 1;
  const float* up_state = B + up_p_state;
  const float* up_prev = B + up_p_state_prev;
  const float* state = B + L.state;
  const float* sprev = B + L.state_prev;
  const int i = T.ok ? T.unit : 0;
  const int nfb = (top || L.fb.r < 0) ? 0 : L.fb.nslots, npr = input_layer ? 0 : L.pred.nslots;
  float err = 0.0f, mod = 0.0f, new_base = 0.0f;
  if (learn) {
    const float actual = input_layer ? P.in0[(long long)T.a * P.n_in + i] : state[i];
    err = actual - B[L.p_state_prev + i];
    if (!input_layer && P.variant == BIDI_KWINNERS) {
      const float surprise = err * err, avg = B[L.p_baseline + i];
      mod = bidi_sigmoid(L.attention_factor * (surprise - avg));
      new_base = (1.0f - L.avg_surprise_decay) * avg + L.avg_surprise_decay * surprise;
    } else if (!input_layer && P.variant == BIDI_ITERATIVE) {
      const float error2 = err * err, base = B[L.p_baseline + i];
      new_base = (1.0f - L.baseline_decay) * base + L.baseline_decay * error2;
    }
  }
  __syncthreads();  // the old baseline has been read by every thread of the unit
  if (learn && T.ok && T.g == 0 && !input_layer) {
    if (P.variant == BIDI_KWINNERS) { B[L.p_mod + i] = mod; B[L.p_baseline + i] = new_base; }
    else if (P.variant == BIDI_ITERATIVE) B[L.p_baseline + i] = new_base;
  }
  const float kfb = L.learn_fb * err, kpr = L.learn_pred * err;
  const int uw = input_layer ? L.fb.uw : L.hw;
  if (nfb)
    tile_products(L.fb, T, uw, B + L.fb.w_off, L.pn, Pbuf, 0, learn != 0, kfb, [&](int t) { return up_prev[t]; },
                  [&](int t) { return up_state[t]; });
  if (npr)
    tile_products(L.pred, T, uw, B + L.pred.w_off, L.pn, Pbuf, nfb, learn != 0, kpr, [&](int t) { return sprev[t]; },
                  [&](int t) { return state[t]; });
  __syncthreads();
  if (T.g == 0 && T.ok) {
    const float activation = tile_sum(Pbuf, 0, nfb + npr, T.ul, 0.0f);
    B[L.p_act + i] = activation;
    if (input_layer) P.pred_out[(long long)T.a * P.n_in + i] = activation;
    else if (P.variant != BIDI_KWINNERS) B[L.p_state + i] = bidi_clamp01(activation);
  }
}

// ---- kernel wrappers: one CTA per tile
__global__ void __launch_bounds__(TILE_U* TILE_G) k_t_kw_activate(EngineDev P, LayerDev L, int l) { t_kw_activate(P, L, l, (int)blockIdx.x); }
__global__ void __launch_bounds__(TILE_U* TILE_G) k_t_kw_inhibit(EngineDev P, LayerDev L, int l, long long src, long long dst, long long dst_next) { t_kw_inhibit(P, L, l, src, dst, dst_next, (int)blockIdx.x); }
__global__ void __launch_bounds__(TILE_U* TILE_G) k_t_reconstruct(EngineDev P, LayerDev L, int l, long long drive_off, int scaled, float mult, int copy_spike, int mode) { t_reconstruct(P, L, l, drive_off, scaled, mult, copy_spike, mode, (int)blockIdx.x); }
__global__ void __launch_bounds__(TILE_U* TILE_G) k_t_kw_learn(EngineDev P, LayerDev L, int l) { t_kw_learn(P, L, l, (int)blockIdx.x); }
__global__ void __launch_bounds__(TILE_U* TILE_G) k_t_ic_sweep(EngineDev P, LayerDev L, int l, int first, int acc, float scale, int it) { t_ic_sweep(P, L, l, first, acc, scale, it, (int)blockIdx.x); }
__global__ void __launch_bounds__(TILE_U* TILE_G) k_t_ic_learn(EngineDev P, LayerDev L, int l, int traced) { t_ic_learn(P, L, l, traced, (int)blockIdx.x); }
__global__ void __launch_bounds__(TILE_U* TILE_G) k_t_predict(EngineDev P, LayerDev L, int l, int learn, long long up_p_state, long long up_p_state_prev) { t_predict(P, L, l, learn, up_p_state, up_p_state_prev, (int)blockIdx.x); }

} // namespace bidi
