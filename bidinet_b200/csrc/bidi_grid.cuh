// bidi_grid.cuh -- the k-winners per-phase kernels for LARGE layers (C4, C5, row strips
// of a split layer): one CTA per 32x8 tile of the 2-D unit grid.
//
// The thread-per-unit kernels of bidi_kernels.cuh pay ~14 instructions per connection
// (index clamps, a global load of the source value next to every weight load) and reach
// 2.6 TB/s on C5.  Here a CTA first stages the part of the source plane its tile's windows
// cover -- the window box plus halo, zero outside the grid -- into shared memory; after
// that a connection costs one coalesced weight load (a warp = 32 adjacent units reads
// w[slot][unit..unit+31], 128 bytes), one shared-memory load, a multiply and an add, with
// no bounds test: slots a unit does not own hold weight 0 and halo cells hold 0, and
// (+-0) added to a running sum leaves it unchanged.  Every sum still runs over the slots
// in the reference's list order (dx outer, dy inner), so the bits are those of
// bidi_kernels.cuh.  The weight stream is the only HBM traffic; unit planes are
// L2-resident.
//
// Launch geometry: blockIdx.x = (agent * tiles_y + tile_y) * tiles_x + tile_x over the
// rows [row0,row1) the launch works on (a row strip, or the whole grid).
#pragma once
#include "bidi_kernels.cuh"

namespace bidi {

#define GRID_TX 32
#define GRID_TY 8

struct GridTile { int a, ux0, uy0, ux, uy, unit, row_end; bool ok; };
__host__ __device__ inline int grid_tiles(int uw, int rows) { return ((uw + GRID_TX - 1) / GRID_TX) * ((rows + GRID_TY - 1) / GRID_TY); }
__device__ __forceinline__ GridTile grid_tile(int b, int uw, int row0, int row1) {
  GridTile t;
  const int tiles_x = (uw + GRID_TX - 1) / GRID_TX, tiles_y = (row1 - row0 + GRID_TY - 1) / GRID_TY;
  const int tx = b % tiles_x;
  b /= tiles_x;
  t.a = b / tiles_y;
  t.ux0 = tx * GRID_TX; t.uy0 = row0 + (b % tiles_y) * GRID_TY;
  t.ux = t.ux0 + (threadIdx.x & 31); t.uy = t.uy0 + (threadIdx.x >> 5);
  t.row_end = row1;
  t.ok = t.ux < uw && t.uy < row1;
  t.unit = t.ux + t.uy * uw;
  return t;
}

// The target cells the windows of a tile's units cover: [x0, x0+w) x [y0, y0+h).
struct Box { int x0, y0, w, h; };
__device__ __forceinline__ Box window_box(const WinDev& w, const GridTile& t) {
  Box b;
  if (w.absx) { b.x0 = 0; b.w = w.tw; }
  else { b.x0 = w.cx[t.ux0] - w.r; b.w = w.cx[min(t.ux0 + GRID_TX, w.uw) - 1] + w.r - b.x0 + 1; }
  if (w.absy) { b.y0 = 0; b.h = w.th; }
  else { b.y0 = w.cy[t.uy0] - w.r; b.h = w.cy[min(t.uy0 + GRID_TY, t.row_end) - 1] + w.r - b.y0 + 1; }
  return b;
}
// sm[(tx-x0) + (ty-y0)*w] = src(tx + ty*tw) inside the grid, 0 in the halo
template <class Src>
__device__ __forceinline__ void stage_box(float* __restrict__ sm, const Box& b, int tw, int th, Src&& src) {
  const int n = b.w * b.h;
  for (int k = threadIdx.x; k < n; k += GRID_TX * GRID_TY) {
    const int by = k / b.w, gx = b.x0 + k - by * b.w, gy = b.y0 + by;
    sm[k] = (gx >= 0 && gx < tw && gy >= 0 && gy < th) ? src(gx + gy * tw) : 0.0f;
  }
}
// sum += sm(target) * w[slot][unit] over the whole slot box, list order.  DYN = the
// window's height when it is one of the usual 2r+1 (the inner loop is then fully unrolled:
// all DYN weight loads of a window column are in flight together), 0 = any height.
template <int DYN, int NC>
__device__ __forceinline__ float box_dot_n(const WinDev& w, const Box& b, const float* __restrict__ s0, const float* __restrict__ wp,
                                           long long U, float sum) {
  const int dyn = DYN ? DYN : w.dyn;
  if (DYN) {
    // NC window columns at a time: all NC*DYN weight loads are in flight before the first add
    for (int sx0 = 0; sx0 < w.dxn; sx0 += NC) {
      float wv[NC][DYN ? DYN : 1];
#pragma unroll
      for (int c = 0; c < NC; c++)
#pragma unroll
        for (int sy = 0; sy < DYN; sy++) wv[c][sy] = sx0 + c < w.dxn ? wp[(long long)(c * DYN + sy) * U] : 0.0f;
#pragma unroll
      for (int c = 0; c < NC; c++)
        if (sx0 + c < w.dxn) {
          const float* sp = s0 + sx0 + c;
#pragma unroll
          for (int sy = 0; sy < DYN; sy++) sum += sp[sy * b.w] * wv[c][sy];
        }
      wp += (long long)NC * DYN * U;
    }
  } else {
    for (int sx = 0; sx < w.dxn; sx++) {
      const float* sp = s0 + sx;
#pragma unroll 4
      for (int sy = 0; sy < dyn; sy++) {
        sum += sp[sy * b.w] * wp[0];
        wp += U;
      }
    }
  }
  return sum;
}
__device__ __forceinline__ float box_dot(const WinDev& w, const Box& b, const float* __restrict__ sm, int ux, int uy,
                                         const float* __restrict__ wp, long long U, float sum) {
  if (w.r < 0) return sum;
  const float* s0 = sm + (w.absx ? 0 : w.cx[ux] - w.r - b.x0) + (w.absy ? 0 : w.cy[uy] - w.r - b.y0) * b.w;
  switch (w.dyn) {
    case 5: return box_dot_n<5, 5>(w, b, s0, wp, U, sum);
    case 7: return box_dot_n<7, 4>(w, b, s0, wp, U, sum);
    case 9: return box_dot_n<9, 3>(w, b, s0, wp, U, sum);
    case 13: return box_dot_n<13, 2>(w, b, s0, wp, U, sum);
    case 17: return box_dot_n<17, 2>(w, b, s0, wp, U, sum);
    default: return box_dot_n<0, 1>(w, b, s0, wp, U, sum);
  }
}
// largest window box of any tile, in floats (host: shared memory to reserve)
inline int grid_box_floats(const WinDev& w, const int* cx, const int* cy) {
  if (w.r < 0) return 0;
  int bw = w.tw, bh = w.th;
  if (!w.absx) { bw = 0; for (int u = 0; u < w.uw; u += GRID_TX) bw = max(bw, cx[min(u + GRID_TX, w.uw) - 1] - cx[u] + 2 * w.r + 1); }
  if (!w.absy) { bh = 0; for (int u = 0; u < w.uh; u++) bh = max(bh, cy[min(u + GRID_TY, w.uh) - 1] - cy[u] + 2 * w.r + 1); }
  return bw * bh;
}

// ---- sdr/RSDR.cpp:123-133: a = -theta + sum_ff x*w + sum_rec sPrev*w
__global__ void __launch_bounds__(GRID_TX* GRID_TY) k_g_kw_activate(EngineDev P, LayerDev L, int l, int ff_floats) {
  extern __shared__ float sm[];
  const GridTile T = grid_tile(blockIdx.x, L.hw, L.row0, L.row1);
  float* B = blob_of(P, T.a);
  const float* x = vis_input_of(P, L, l, T.a);
  const float* sprev = B + L.state_prev;
  const Box bf = window_box(L.ff, T), br = L.rec.r >= 0 ? window_box(L.rec, T) : Box{0, 0, 0, 0};
  float* sm_rec = sm + ff_floats;
  stage_box(sm, bf, L.ff.tw, L.ff.th, [&](int t) { return x[t]; });
  if (L.rec.r >= 0) stage_box(sm_rec, br, L.rec.tw, L.rec.th, [&](int t) { return sprev[t]; });
  __syncthreads();
  if (!T.ok) return;
  float sum = -B[L.thr + T.unit];
  sum = box_dot(L.ff, bf, sm, T.ux, T.uy, B + L.ff.w_off + T.unit, L.nh, sum);
  sum = box_dot(L.rec, br, sm_rec, T.ux, T.uy, B + L.rec.w_off + T.unit, L.nh, sum);
  B[L.act + T.unit] = sum;
}

// ---- sdr/RSDR.cpp:135-145, :148-163: rank among the lateral neighbours
__global__ void __launch_bounds__(GRID_TX* GRID_TY) k_g_kw_inhibit(EngineDev P, LayerDev L, int l, long long src, long long dst, long long dst_next) {
  extern __shared__ float sm[];
  const GridTile T = grid_tile(blockIdx.x, L.hw, L.row0, L.row1);
  float* B = blob_of(P, T.a);
  const float* act = B + src;
  const WinDev& w = L.lat;
  const Box b = window_box(w, T);
  stage_box(sm, b, w.tw, w.th, [&](int t) { return act[t]; });
  __syncthreads();
  if (!T.ok) return;
  const SlotBox sb = slot_box(w, T.ux, T.uy);
  const float mine = sm[(T.ux - b.x0) + (T.uy - b.y0) * b.w];
  int cnt = 0;
  const int ny = sb.sy1 - sb.sy0 + 1;
  for (int sx = sb.sx0; sx <= sb.sx1; sx++) {
    const float* sp = sm + (slot_tx(w, sb, sx) - b.x0) + (slot_ty(w, sb, sb.sy0) - b.y0) * b.w;
    if (ny == 13) {  // an interior unit of a radius-6 layer: the whole column at once
#pragma unroll
      for (int q = 0; q < 13; q++) cnt += sp[q * b.w] >= mine ? 1 : 0;
    } else if (ny == 7) {
#pragma unroll
      for (int q = 0; q < 7; q++) cnt += sp[q * b.w] >= mine ? 1 : 0;
    } else {
#pragma unroll 4
      for (int q = 0; q < ny; q++) cnt += sp[q * b.w] >= mine ? 1 : 0;
    }
  }
  cnt -= 1;  // the loop counted the unit itself (mine >= mine); the lateral list excludes the centre
  const float s = (float)cnt < L.sparsity * (float)slot_count(w, sb) ? 1.0f : 0.0f;
  B[dst + T.unit] = s;
  if (dst_next >= 0) B[dst_next + T.unit] = s;
}

// ---- gather-form reconstruction (sdr/RSDR.cpp:165-212): tiles [0,vt) of an agent own
// visible cells (feed-forward family), the rest hidden units (recurrent family);
// to_pred: the visible result is the input-space prediction (sdr/PredictiveRSDR.cpp:188-193).
__global__ void __launch_bounds__(GRID_TX* GRID_TY) k_g_reconstruct(EngineDev P, LayerDev L, int l, long long drive_off, int hidden_too, int to_pred,
                                                                      int cap, int scaled, float mult, int copy_spike) {
  extern __shared__ float sm[];
  const int vt = grid_tiles(L.vw, L.vrow1 - L.vrow0), ht = hidden_too ? grid_tiles(L.hw, L.row1 - L.row0) : 0;
  const int a = blockIdx.x / (vt + ht), tile = blockIdx.x % (vt + ht);
  const bool vis = tile < vt;
  const WinDev& w = vis ? L.ff : L.rec;
  const GridTile T = vis ? grid_tile(tile, w.tw, L.vrow0, L.vrow1) : grid_tile(tile - vt, w.tw, L.row0, L.row1);
  float* B = blob_of(P, a);
  const float* drive = B + drive_off;
  float* out = vis ? (to_pred ? P.pred_out + (long long)a * P.n_in : B + L.vis_recon) : B + L.recon;
  if (w.r < 0) {  // no recurrent family: the reconstruction is 0
    if (T.ok) {
      out[T.unit] = 0.0f;
      if (!vis && copy_spike) B[L.spike_prev + T.unit] = B[L.spike + T.unit];
    }
    return;
  }
  // the units whose windows can contain a cell of this tile: a box of the unit grid
  const int tx1 = min(T.ux0 + GRID_TX, w.tw) - 1, ty1 = min(T.uy0 + GRID_TY, T.row_end) - 1;
  int bx0 = w.uw, bx1 = -1, by0 = w.uh, by1 = -1;
  for (int t = T.ux0; t <= tx1; t++) { bx0 = min(bx0, w.gx_lo[t]); bx1 = max(bx1, w.gx_hi[t]); }
  for (int t = T.uy0; t <= ty1; t++) { by0 = min(by0, w.gy_lo[t]); by1 = max(by1, w.gy_hi[t]); }
  const int bw = max(0, bx1 - bx0 + 1), bh = max(0, by1 - by0 + 1), nb = bw * bh;
  // The drive is sparse (k-winners: ~1 % of the units), so instead of every cell walking
  // its (2r+1)^2 candidates the CTA compacts the non-zero units of the box into a list,
  // in ascending unit index -- the order the reference's scatter adds them -- and every cell
  // walks the list.  Warp g compacts the g-th eighth of the box; a scan over the eight
  // counts gives each warp its place.
  int* l_unit = reinterpret_cast<int*>(sm);   // [nb] unit index
  int* l_ctr = l_unit + cap;                  // [nb] window centre, cx | cy << 16
  float* l_d = sm + 2 * cap;                  // [nb] drive
  __shared__ int s_cnt[GRID_TY + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = ((nb + GRID_TY - 1) / GRID_TY + 31) / 32 * 32, k0 = warp * chunk, k1 = min(nb, k0 + chunk);
  int mine = 0;
  for (int k = k0 + lane; k < k0 + chunk; k += 32) {
    float d = 0.0f;
    if (k < k1) { const int by = k / bw; d = drive[(bx0 + k - by * bw) + (by0 + by) * w.uw]; }
    mine += __popc(__ballot_sync(0xffffffffu, d != 0.0f));
  }
  if (lane == 0) s_cnt[warp + 1] = mine;
  __syncthreads();
  if (threadIdx.x == 0) { s_cnt[0] = 0; for (int g = 0; g < GRID_TY; g++) s_cnt[g + 1] += s_cnt[g]; }
  __syncthreads();
  int pos = s_cnt[warp];
  for (int k = k0 + lane; k < k0 + chunk; k += 32) {
    float d = 0.0f;
    int ux = 0, uy = 0;
    if (k < k1) { const int by = k / bw; ux = bx0 + k - by * bw; uy = by0 + by; d = drive[ux + uy * w.uw]; }
    const unsigned m = __ballot_sync(0xffffffffu, d != 0.0f);
    if (d != 0.0f) {
      const int q = pos + __popc(m & ((1u << lane) - 1u));
      l_unit[q] = ux + uy * w.uw; l_ctr[q] = w.cx[ux] | (w.cy[uy] << 16); l_d[q] = d;
    }
    pos += __popc(m);
  }
  __syncthreads();
  if (!T.ok) return;
  const int tx = T.ux, ty = T.uy, n = s_cnt[GRID_TY], r = w.r;
  const float* wp = B + w.w_off;
  float sum = 0.0f;
  // list entries whose window contains this cell, NQ at a time: the weight loads of a batch are all
  // issued before the batch is added up, in list order (neo/SparseCoder.cpp:242-259: w*state*multiplier)
  constexpr int NQ = 8;
  float wv[NQ], dv[NQ];
  int q = 0;
  for (int k = 0; k < n; k++) {
    const int c = l_ctr[k], dx = tx - (c & 0xffff), dy = ty - (c >> 16);
    if (dx >= -r && dx <= r && dy >= -r && dy <= r && !(w.excl && (dx | dy) == 0)) {
      const int sx = w.absx ? tx : dx + r, sy = w.absy ? ty : dy + r;
      wv[q] = wp[(long long)(sx * w.dyn + sy) * w.U + l_unit[k]];
      dv[q] = l_d[k];
      if (++q == NQ) {
#pragma unroll
        for (int j = 0; j < NQ; j++) sum += scaled ? wv[j] * dv[j] * mult : wv[j] * dv[j];
        q = 0;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NQ; j++)
    if (j < q) sum += scaled ? wv[j] * dv[j] * mult : wv[j] * dv[j];
  out[T.unit] = sum;
  if (!vis && copy_spike) B[L.spike_prev + T.unit] = B[L.spike + T.unit];  // spikePrev = spike (sdr/IRSDR.cpp:170-171)
}
// host: largest candidate box of any tile of targets, in floats
inline int grid_gather_floats(const WinDev& w, const int* gxlo, const int* gxhi, const int* gylo, const int* gyhi) {
  if (w.r < 0) return 0;
  int bw = 0, bh = 0;
  for (int t0 = 0; t0 < w.tw; t0 += GRID_TX) {
    int lo = w.uw, hi = -1;
    for (int t = t0; t < min(t0 + GRID_TX, w.tw); t++) { lo = min(lo, gxlo[t]); hi = max(hi, gxhi[t]); }
    bw = max(bw, hi - lo + 1);
  }
  for (int t0 = 0; t0 < w.th; t0++) {
    int lo = w.uh, hi = -1;
    for (int t = t0; t < min(t0 + GRID_TY, w.th); t++) { lo = min(lo, gylo[t]); hi = max(hi, gyhi[t]); }
    bh = max(bh, hi - lo + 1);
  }
  return max(0, bw) * max(0, bh);
}

// ---- one leaky-integrate-and-fire sweep of the iterative coders (sdr/IRSDR.cpp:138-168, neo/SparseCoder.cpp:129-158)
// for large layers: the three source boxes (visible error, hidden error, previous spikes) staged once per tile
__global__ void __launch_bounds__(GRID_TX* GRID_TY) k_g_ic_sweep(EngineDev P, LayerDev L, int l, int first, int acc, float scale, int it,
                                                                   int ff_floats, int rec_floats) {
  extern __shared__ float sm[];
  const GridTile T = grid_tile(blockIdx.x, L.hw, L.row0, L.row1);
  float* B = blob_of(P, T.a);
  const float* x = vis_input_of(P, L, l, T.a);
  const float* vrec = B + L.vis_recon;
  const float* sprev = B + L.state_prev;
  const float* hrec = B + L.recon;
  const float* spk_prev = B + L.spike_prev;
  const bool has_rec = L.rec.r >= 0;
  const Box bf = window_box(L.ff, T), br = has_rec ? window_box(L.rec, T) : Box{0, 0, 0, 0}, bl = window_box(L.lat, T);
  float* sm_rec = sm + ff_floats;
  float* sm_lat = sm_rec + rec_floats;
  stage_box(sm, bf, L.ff.tw, L.ff.th, [&](int t) { return x[t] - vrec[t]; });
  if (has_rec) stage_box(sm_rec, br, L.rec.tw, L.rec.th, [&](int t) { return sprev[t] - hrec[t]; });
  stage_box(sm_lat, bl, L.lat.tw, L.lat.th, [&](int t) { return spk_prev[t]; });
  __syncthreads();
  if (!T.ok) return;
  const int i = T.unit;
  float excitation = P.gen_on ? P.gen_normals[(long long)T.a * P.gen_per_agent + L.gen_off + (long long)it * L.nh + i] * *P.gen_scale : 0.0f;
  excitation = box_dot(L.ff, bf, sm, T.ux, T.uy, B + L.ff.w_off + i, L.nh, excitation);
  excitation = box_dot(L.rec, br, sm_rec, T.ux, T.uy, B + L.rec.w_off + i, L.nh, excitation);
  const float inhibition = box_dot(L.lat, bl, sm_lat, T.ux, T.uy, B + L.lat.w_off + i, L.nh, 0.0f);
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

// ---- prediction nodes of a k-winners layer: sdr/PredictiveRSDR.cpp:122-160 (delta rule
// first -- it only touches the few units whose prediction was wrong -- then the activation)
__global__ void __launch_bounds__(GRID_TX* GRID_TY) k_g_kw_predict(EngineDev P, LayerDev L, int l, int learn, long long up_p_state,
                                                                     long long up_p_state_prev, int fb_floats) {
  extern __shared__ float sm[];
  const GridTile T = grid_tile(blockIdx.x, L.hw, L.row0, L.row1);
  float* B = blob_of(P, T.a);
  const bool top = l == P.L - 1;
  const float* up_state = B + up_p_state;
  const float* up_prev = B + up_p_state_prev;
  const float* state = B + L.state;
  const float* sprev = B + L.state_prev;
  float* wfb = B + L.fb.w_off;
  float* wpr = B + L.pred.w_off;
  const long long U = L.pn;
  const Box bf = top ? Box{0, 0, 0, 0} : window_box(L.fb, T), bp = window_box(L.pred, T);
  float* sm_pred = sm + fb_floats;
  if (!top) stage_box(sm, bf, L.fb.tw, L.fb.th, [&](int t) { return up_state[t]; });
  stage_box(sm_pred, bp, L.pred.tw, L.pred.th, [&](int t) { return state[t]; });
  __syncthreads();
  const int i = T.unit;
  float kfb = 0.0f, kpr = 0.0f;
  bool wrong = false;
  if (learn && T.ok) {
    const float err = state[i] - B[L.p_state_prev + i];
    const float surprise = err * err, avg = B[L.p_baseline + i];
    B[L.p_mod + i] = bidi_sigmoid(L.attention_factor * (surprise - avg));
    B[L.p_baseline + i] = (1.0f - L.avg_surprise_decay) * avg + L.avg_surprise_decay * surprise;
    kfb = L.learn_fb * err; kpr = L.learn_pred * err;
    wrong = err != 0.0f;
  }
  // delta rule: w += k*source for the few units whose prediction was wrong (a zero error or
  // a zero source adds +-0).  The rule is element-wise, so the 32 lanes of the warp share the
  // connections of each such unit instead of one lane walking all (2r+1)^2 of them.
  for (unsigned todo = __ballot_sync(0xffffffffu, wrong); todo; todo &= todo - 1) {
    const int src = __ffs(todo) - 1, lane = threadIdx.x & 31;
    const int ux = T.ux0 + src, uy = T.uy, unit = ux + uy * L.hw;
    const float kfb_w = __shfl_sync(0xffffffffu, kfb, src), kpr_w = __shfl_sync(0xffffffffu, kpr, src);
    for (int fam = top ? 1 : 0; fam < 2; fam++) {
      const WinDev& w = fam ? L.pred : L.fb;
      if (w.r < 0) continue;
      float* wq = (fam ? wpr : wfb) + unit;
      const float* from = fam ? sprev : up_prev;
      const float k = fam ? kpr_w : kfb_w;
      const SlotBox sb = slot_box(w, ux, uy);
      for (int sl = lane; sl < w.nslots; sl += 32) {
        const int sx = sl / w.dyn, sy = sl - sx * w.dyn;
        const int tx = slot_tx(w, sb, sx), ty = slot_ty(w, sb, sy);
        if (sx >= sb.sx0 && sx <= sb.sx1 && sy >= sb.sy0 && sy <= sb.sy1 && !(w.excl && tx == sb.cx && ty == sb.cy)) {
          const float v = from[tx + ty * w.tw];
          if (v != 0.0f) wq[(long long)sl * U] += k * v;
        }
      }
    }
  }
  __syncwarp();
  if (!T.ok) return;
  float activation = 0.0f;
  if (!top) activation = box_dot(L.fb, bf, sm, T.ux, T.uy, wfb + i, U, activation);
  activation = box_dot(L.pred, bp, sm_pred, T.ux, T.uy, wpr + i, U, activation);
  B[L.p_act + i] = activation;
}

// ---- sdr/RSDR.cpp:241-266: winners only; theta for every unit.  A warp owns 32 adjacent
// units; for each winner among them the 32 lanes share its connections (the rule is
// element-wise), so a winner's 2*(2r+1)^2 weights take a dozen round trips, not hundreds.
__global__ void __launch_bounds__(GRID_TX* GRID_TY) k_g_kw_learn(EngineDev P, LayerDev L, int l) {
  const GridTile T = grid_tile(blockIdx.x, L.hw, L.row0, L.row1);
  float* B = blob_of(P, T.a);
  const float* x = vis_input_of(P, L, l, T.a);
  const float* vrec = B + L.vis_recon;
  const float* sprev = B + L.state_prev;
  const float* hrec = B + L.recon;
  const int lane = threadIdx.x & 31;
  const long long U = L.nh;
  const float learn = T.ok ? B[L.state + T.unit] : 0.0f;
  const float att = learn > 0.0f ? B[L.p_mod + T.unit] : 0.0f;
  const float kf = L.learn_ff * att * learn, kr = L.learn_rec * att * learn;
  unsigned winners = __ballot_sync(0xffffffffu, learn > 0.0f);
  for (; winners; winners &= winners - 1) {
    const int src = __ffs(winners) - 1;
    const int ux = T.ux0 + src, uy = T.uy, unit = ux + uy * L.hw;
    const float kf_w = __shfl_sync(0xffffffffu, kf, src), kr_w = __shfl_sync(0xffffffffu, kr, src);
    for (int fam = 0; fam < 2; fam++) {
      const WinDev& w = fam ? L.rec : L.ff;
      if (w.r < 0) continue;
      float* wp = B + w.w_off + unit;
      const float* a = fam ? sprev : x;
      const float* b = fam ? hrec : vrec;
      const float k = fam ? kr_w : kf_w;
      const SlotBox sb = slot_box(w, ux, uy);
      for (int s = lane; s < w.nslots; s += 32) {
        const int sx = s / w.dyn, sy = s - sx * w.dyn;
        const int tx = slot_tx(w, sb, sx), ty = slot_ty(w, sb, sy);
        if (sx >= sb.sx0 && sx <= sb.sx1 && sy >= sb.sy0 && sy <= sb.sy1 && !(w.excl && tx == sb.cx && ty == sb.cy)) {
          const int t = tx + ty * w.tw;
          wp[(long long)s * U] += k * (a[t] - b[t]);
        }
      }
    }
  }
  if (T.ok) B[L.thr + T.unit] += L.learn_threshold * (learn - L.sparsity);
}

} // namespace bidi
