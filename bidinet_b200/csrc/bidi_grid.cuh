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
#include "bidi_columns.cuh"  // mbarrier / bulk-copy helpers

namespace bidi {

#define GRID_TX 32
#define GRID_TY 8

// Element (slot, unit) of a connection family in either layout (bidi_types.h, WinDev::row_pitch).
__device__ __forceinline__ long long widx(const WinDev& w, int slot, int unit) {
  return w.row_pitch ? (long long)unit * w.row_pitch + slot : (long long)slot * w.U + unit;
}

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

// host: largest window box of a 32x1 row segment, in floats
inline int grid_rowbox_floats(const WinDev& w, const int* cx) {
  if (w.r < 0) return 0;
  int bw = w.tw;
  if (!w.absx) { bw = 0; for (int u = 0; u < w.uw; u += GRID_TX) bw = max(bw, cx[min(u + GRID_TX, w.uw) - 1] - cx[u] + 2 * w.r + 1); }
  const int bh = w.absy ? w.th : 2 * w.r + 1;
  return bw * bh;
}

// ---- Large k-winners layers keep every family as ROWS, w[unit][slot] (pitch = 4 mod 8 floats): a unit's window is
// contiguous in HBM.  What that buys, given that the sources of a k-winners hierarchy are SPARSE (the states and
// predictions are binary with ~1 % ones; only the input of layer 0 is dense):
//   * a sum over a sparse source walks the few non-zero cells of the tile's window box (compacted once per CTA, in
//     x-major order = the order of every unit's connection list) and reads ONE weight per hit -- the other
//     (2r+1)^2 - hits products are (+-0) and leave the running sum unchanged, so they are not formed and their
//     weights never leave HBM;
//   * the dense feed-forward sum of layer 0 stages the 32 rows of a row segment with ONE TMA bulk copy
//     (cp.async.bulk + mbarrier) and reads them with conflict-free 128-bit shared-memory loads;
//   * the winners-only learning rule and the gather reconstructions touch a winner's row as consecutive sectors
//     instead of one 32-byte sector per weight.
// Non-zero cells of plane[] inside box b, in x-major order (tx outer, ty inner), as (tx | ty << 16, value bits).
// One warp: returns the count.  The box may reach outside the grid (cells there are 0 and never listed).
__device__ __forceinline__ int compact_box_warp(const float* __restrict__ plane, const Box& b, int tw, int th, int2* __restrict__ ent, int lane) {
  const int n = b.w * b.h;
  int base = 0;
  for (int k0 = 0; k0 < n; k0 += 128) {  // four independent loads in flight per lane
    float v[4]; int cc[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int k = k0 + 32 * t + lane, bx = k / b.h, gx = b.x0 + bx, gy = b.y0 + k - bx * b.h;
      const bool in = k < n && gx >= 0 && gx < tw && gy >= 0 && gy < th;
      v[t] = in ? plane[gx + gy * tw] : 0.0f;
      cc[t] = gx | (gy << 16);
    }
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const unsigned m = __ballot_sync(0xffffffffu, v[t] != 0.0f);
      if (v[t] != 0.0f) ent[base + __popc(m & ((1u << lane) - 1u))] = make_int2(cc[t], __float_as_int(v[t]));
      base += __popc(m);
    }
  }
  __syncwarp();
  return base;
}
// The same by a whole CTA of GRID_TY warps (two barriers inside): warp g lists the g-th part of the traversal.
__device__ __forceinline__ int compact_box_cta(const float* __restrict__ plane, const Box& b, int tw, int th, int2* __restrict__ ent, int* s_cnt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = b.w * b.h;
  const int chunk = ((n + GRID_TY - 1) / GRID_TY + 31) / 32 * 32, k0 = warp * chunk, k1 = min(n, k0 + chunk);
  int mine = 0;
  for (int k = k0 + lane; k < k0 + chunk; k += 32) {
    float d = 0.0f;
    if (k < k1) {
      const int bx = k / b.h, gx = b.x0 + bx, gy = b.y0 + k - bx * b.h;
      if (gx >= 0 && gx < tw && gy >= 0 && gy < th) d = plane[gx + gy * tw];
    }
    mine += __popc(__ballot_sync(0xffffffffu, d != 0.0f));
  }
  __syncthreads();  // (the previous list's readers are done with s_cnt)
  if (lane == 0) s_cnt[warp + 1] = mine;
  __syncthreads();
  int pos = 0;
  for (int g = 0; g < warp; g++) pos += s_cnt[g + 1];
  int total = 0;
  for (int g = 0; g < GRID_TY; g++) total += s_cnt[g + 1];
  for (int k = k0 + lane; k < k0 + chunk; k += 32) {
    float d = 0.0f;
    int cc = 0;
    if (k < k1) {
      const int bx = k / b.h, gx = b.x0 + bx, gy = b.y0 + k - bx * b.h;
      cc = gx | (gy << 16);
      if (gx >= 0 && gx < tw && gy >= 0 && gy < th) d = plane[gx + gy * tw];
    }
    const unsigned m = __ballot_sync(0xffffffffu, d != 0.0f);
    if (d != 0.0f) ent[pos + __popc(m & ((1u << lane) - 1u))] = make_int2(cc, __float_as_int(d));
    pos += __popc(m);
  }
  __syncthreads();
  return total;
}
// sum += value * w[unit][slot] over the listed cells inside the unit's window, list order.  Pass 1 marks the hits among
// the list entries (64 at a time, a bit each); pass 2 loads the weights of four hits together before adding them in order.
__device__ __forceinline__ float list_dot(const WinDev& w, const int2* __restrict__ ent, int n, int ux, int uy, const float* __restrict__ wrow, float sum) {
  if (w.r < 0) return sum;
  const int cx = w.cx[ux], cy = w.cy[uy], r = w.r;
  const unsigned r2 = 2u * (unsigned)r;
  auto slot_of = [&](int c) { const int tx = c & 0xffff, ty = c >> 16; return (w.absx ? tx : tx - cx + r) * w.dyn + (w.absy ? ty : ty - cy + r); };
  for (int k0 = 0; k0 < n; k0 += 64) {
    unsigned long long hits = 0ull;
    const int kn = min(64, n - k0);
    for (int k = 0; k < kn; k++) {
      const int c = ent[k0 + k].x, dx = (c & 0xffff) - cx, dy = (c >> 16) - cy;
      const bool in = (unsigned)(dx + r) <= r2 && (unsigned)(dy + r) <= r2 && !(w.excl && (dx | dy) == 0);
      hits |= (unsigned long long)in << k;
    }
    while (hits) {
      float wv[4], dv[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        wv[j] = 0.0f; dv[j] = 0.0f;
        if (hits) {
          const int k = __ffsll((long long)hits) - 1;
          hits &= hits - 1ull;
          const int2 e = ent[k0 + k];
          wv[j] = wrow[slot_of(e.x)]; dv[j] = __int_as_float(e.y);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; j++) sum += dv[j] * wv[j];  // (an absent hit adds +0)
    }
  }
  return sum;
}

// ---- sdr/RSDR.cpp:123-133: a = -theta + sum_ff x*w + sum_rec sPrev*w.  One WARP per row segment of 32 units.
// dense_ff: the visible plane is dense (layer 0): its 32 weight rows are staged by one bulk copy and the window box
// by the warp.  Otherwise (the visible plane is the binary state of the layer below) the feed-forward sum is sparse too.
// smem: [32*pitch slab | ff box] (dense) , list entries, mbarrier
template <int DYN>
__global__ void __launch_bounds__(32) k_g_kw_activate(EngineDev P, LayerDev L, int l, int dense_ff, int slab_floats, int box_floats, int cap) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x;
  const int nseg = (L.hw + 31) >> 5, rows = L.row1 - L.row0;
  int bi = blockIdx.x;
  const int seg = bi % nseg; bi /= nseg;
  const int a = bi / rows, uy = L.row0 + bi % rows;
  GridTile T;
  T.a = a; T.ux0 = seg * 32; T.uy0 = uy; T.ux = T.ux0 + lane; T.uy = uy; T.row_end = uy + 1;
  T.ok = T.ux < L.hw; T.unit = T.ux + uy * L.hw;
  float* B = blob_of(P, a);
  const float* x = vis_input_of(P, L, l, a);
  float* slab = sm;
  float* xbox = sm + slab_floats;
  int2* ent = reinterpret_cast<int2*>(xbox + box_floats);
  uint64_t* bar = reinterpret_cast<uint64_t*>(ent + cap);
  const int nun = min(32, L.hw - T.ux0);
  const WinDev& wf = L.ff;
  if (dense_ff) {
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(nun * wf.row_pitch) * 4u;
      mbar_expect_tx(bar, bytes);
      tma_load_1d(slab, B + wf.w_off + (long long)(T.ux0 + uy * L.hw) * wf.row_pitch, bytes, bar);
    }
  }
  float sum = T.ok ? -B[L.thr + T.unit] : 0.0f;
  const Box bf = window_box(wf, T);
  if (dense_ff) {
    // the window box of the segment: coalesced rows, zero outside the grid
    const int n = bf.w * bf.h;
    for (int k0 = 0; k0 < n; k0 += 256) {
      float v[8];
#pragma unroll
      for (int t = 0; t < 8; t++) {
        const int k = k0 + 32 * t + lane, by = k / bf.w, gx = bf.x0 + k - by * bf.w, gy = bf.y0 + by;
        v[t] = (k < n && gx >= 0 && gx < wf.tw && gy >= 0 && gy < wf.th) ? x[gx + gy * wf.tw] : 0.0f;
      }
#pragma unroll
      for (int t = 0; t < 8; t++) { const int k = k0 + 32 * t + lane; if (k < n) xbox[k] = v[t]; }
    }
    __syncwarp();
    mbar_wait(bar, 0);
    if (T.ok) {
      const float* s0 = xbox + (wf.absx ? 0 : wf.cx[T.ux] - wf.r - bf.x0) + (wf.absy ? 0 : wf.cy[uy] - wf.r - bf.y0) * bf.w;
      const float* wrow = slab + lane * wf.row_pitch;
      if (DYN > 0) {
        // whole window unrolled: DYN*DYN slots, 128-bit weight loads (rows 4 mod 8 floats apart: conflict-free)
        constexpr int DD = DYN > 0 ? DYN : 1, NS = DYN * DYN, NQ4 = (NS + 3) / 4;
#pragma unroll
        for (int q = 0; q < NQ4; q++) {
          const float4 w4 = reinterpret_cast<const float4*>(wrow)[q];
          const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int s = 4 * q + e;
            if (s < NS) sum += s0[(s / DD) + (s % DD) * bf.w] * wv[e];
          }
        }
      } else {
        for (int sx = 0; sx < wf.dxn; sx++)
          for (int sy = 0; sy < wf.dyn; sy++) sum += s0[sx + sy * bf.w] * wrow[sx * wf.dyn + sy];
      }
    }
  } else if (wf.r >= 0) {
    const int n = compact_box_warp(x, bf, wf.tw, wf.th, ent, lane);
    if (T.ok) sum = list_dot(wf, ent, n, T.ux, uy, B + wf.w_off + (long long)T.unit * wf.row_pitch, sum);
    __syncwarp();
  }
  if (L.rec.r >= 0) {
    const Box br = window_box(L.rec, T);
    const int n = compact_box_warp(B + L.state_prev, br, L.rec.tw, L.rec.th, ent, lane);
    if (T.ok) sum = list_dot(L.rec, ent, n, T.ux, uy, B + L.rec.w_off + (long long)T.unit * L.rec.row_pitch, sum);
  }
  if (T.ok) B[L.act + T.unit] = sum;
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
      wv[q] = wp[widx(w, sx * w.dyn + sy, l_unit[k])];
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

// ---- prediction nodes of a k-winners layer: sdr/PredictiveRSDR.cpp:122-160 (delta rule first -- it only touches
// the few units whose prediction was wrong -- then the activation).  Every source here is a binary k-winners plane
// (the layer's states, the prediction of the layer above), so both the rule and the sums walk the lists of the
// non-zero cells of the tile's window boxes: rule  w += k*source  is a no-op where the source is 0 (adds +-0), the
// activation  sum w*source  likewise.  Rows layout; a lane owns its unit's two rows.
// smem: four lists of `cap` entries (state, previous state; upper prediction, its previous value)
__global__ void __launch_bounds__(GRID_TX* GRID_TY) k_g_kw_predict(EngineDev P, LayerDev L, int l, int learn, long long up_p_state,
                                                                     long long up_p_state_prev, int cap_pred, int cap_fb) {
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_cnt[GRID_TY + 1];
  const GridTile T = grid_tile(blockIdx.x, L.hw, L.row0, L.row1);
  float* B = blob_of(P, T.a);
  const bool top = l == P.L - 1 || L.fb.r < 0;
  int2* l_state = reinterpret_cast<int2*>(sm);
  int2* l_sprev = l_state + cap_pred;
  int2* l_up = l_sprev + cap_pred;
  int2* l_upprev = l_up + cap_fb;
  const Box bp = window_box(L.pred, T);
  const int n_state = compact_box_cta(B + L.state, bp, L.pred.tw, L.pred.th, l_state, s_cnt);
  const int n_sprev = learn ? compact_box_cta(B + L.state_prev, bp, L.pred.tw, L.pred.th, l_sprev, s_cnt) : 0;
  int n_up = 0, n_upprev = 0;
  if (!top) {
    const Box bf = window_box(L.fb, T);
    n_up = compact_box_cta(B + up_p_state, bf, L.fb.tw, L.fb.th, l_up, s_cnt);
    if (learn) n_upprev = compact_box_cta(B + up_p_state_prev, bf, L.fb.tw, L.fb.th, l_upprev, s_cnt);
  }
  if (!T.ok) return;
  const int i = T.unit;
  float* wfb = top ? nullptr : B + L.fb.w_off + (long long)i * L.fb.row_pitch;
  float* wpr = B + L.pred.w_off + (long long)i * L.pred.row_pitch;
  if (learn) {
    const float err = B[L.state + i] - B[L.p_state_prev + i];
    const float surprise = err * err, avg = B[L.p_baseline + i];
    B[L.p_mod + i] = bidi_sigmoid(L.attention_factor * (surprise - avg));
    B[L.p_baseline + i] = (1.0f - L.avg_surprise_decay) * avg + L.avg_surprise_decay * surprise;
    if (err != 0.0f) {
      const float kfb = L.learn_fb * err, kpr = L.learn_pred * err;
      for (int fam = top ? 1 : 0; fam < 2; fam++) {
        const WinDev& w = fam ? L.pred : L.fb;
        if (w.r < 0) continue;
        float* wr = fam ? wpr : wfb;
        const int2* ent = fam ? l_sprev : l_upprev;
        const int n = fam ? n_sprev : n_upprev;
        const float k = fam ? kpr : kfb;
        const int cx = w.cx[T.ux], cy = w.cy[T.uy], r = w.r;
        for (int q = 0; q < n; q++) {
          const int2 e = ent[q];
          const int tx = e.x & 0xffff, ty = e.x >> 16, dx = tx - cx, dy = ty - cy;
          if (dx >= -r && dx <= r && dy >= -r && dy <= r && !(w.excl && (dx | dy) == 0)) {
            const int sl = (w.absx ? tx : dx + r) * w.dyn + (w.absy ? ty : dy + r);
            wr[sl] += k * __int_as_float(e.y);
          }
        }
      }
    }
  }
  float activation = 0.0f;
  if (!top) activation = list_dot(L.fb, l_up, n_up, T.ux, T.uy, wfb, activation);
  activation = list_dot(L.pred, l_state, n_state, T.ux, T.uy, wpr, activation);
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
      float* wp = B + w.w_off;
      const float* a = fam ? sprev : x;
      const float* b = fam ? hrec : vrec;
      const float k = fam ? kr_w : kf_w;
      const SlotBox sb = slot_box(w, ux, uy);
      for (int s = lane; s < w.nslots; s += 32) {
        const int sx = s / w.dyn, sy = s - sx * w.dyn;
        const int tx = slot_tx(w, sb, sx), ty = slot_ty(w, sb, sy);
        if (sx >= sb.sx0 && sx <= sb.sx1 && sy >= sb.sy0 && sy <= sb.sy1 && !(w.excl && tx == sb.cx && ty == sb.cy)) {
          const int t = tx + ty * w.tw;
          wp[widx(w, s, unit)] += k * (a[t] - b[t]);
        }
      }
    }
  }
  if (T.ok) B[L.thr + T.unit] += L.learn_threshold * (learn - L.sparsity);
}

} // namespace bidi
