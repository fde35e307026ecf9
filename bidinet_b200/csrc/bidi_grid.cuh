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
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched from the driver at run time)
#include "bidi_kernels.cuh"
#include "bidi_columns.cuh"  // mbarrier / bulk-copy helpers

namespace bidi {

#define GRID_TX 32
#define GRID_TY 8

// k / d for 0 <= k < 65536, 1 <= d <= 256 without an integer division: the quotient of (k + 0.5) / d is at least
// 0.5/d away from an integer, far more than the float error of the product.
__device__ __forceinline__ int fast_div(int k, float inv_d) { return (int)(((float)k + 0.5f) * inv_d); }

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

// host: largest window box of a 32x1 row segment (width, height; floats)
inline void grid_rowbox_dims(const WinDev& w, const int* cx, int& bw, int& bh) {
  bw = bh = 0;
  if (w.r < 0) return;
  bw = w.tw;
  if (!w.absx) { bw = 0; for (int u = 0; u < w.uw; u += GRID_TX) bw = max(bw, cx[min(u + GRID_TX, w.uw) - 1] - cx[u] + 2 * w.r + 1); }
  bh = w.absy ? w.th : 2 * w.r + 1;
}
inline int grid_rowbox_floats(const WinDev& w, const int* cx) {
  int bw, bh;
  grid_rowbox_dims(w, cx, bw, bh);
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
// Non-zero cells of NL planes inside NL boxes, compacted at once by a CTA of GRID_TY warps.  XMAJOR: traversal tx outer,
// ty inner (the order of every unit's connection list); else ty outer, tx inner (ascending unit index, the order of
// the reference's scatter).  A box may reach outside its grid (cells there are 0 and never listed).  Every list's loads
// are issued before any value is consumed -- one memory round trip for all lists -- and the values stay in registers
// between the counting and the filling pass.  emit(list, position, gx, gy, value) stores an entry.  s_cnt: [NL][GRID_TY+1].
template <int NL, bool XMAJOR, class Emit>
__device__ __forceinline__ void compact_boxes_cta(const float* const (&plane)[NL], const Box (&b)[NL], const int (&tw)[NL], const int (&th)[NL],
                                                  int (&n_out)[NL], int (*s_cnt)[GRID_TY + 1], Emit&& emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int VM = 4, ROUND = GRID_TY * 32 * VM;  // cells of a list one round covers
  int nmax = 0;
  float inv[NL];
#pragma unroll
  for (int l = 0; l < NL; l++) {
    n_out[l] = 0;
    nmax = max(nmax, b[l].w * b[l].h);
    inv[l] = 1.0f / (float)max(1, XMAJOR ? b[l].h : b[l].w);
  }
  for (int r0 = 0; r0 < nmax; r0 += ROUND) {
    float v[NL][VM];
#pragma unroll
    for (int l = 0; l < NL; l++) {
      const int n = b[l].w * b[l].h;
#pragma unroll
      for (int t = 0; t < VM; t++) {
        const int k = r0 + (warp * VM + t) * 32 + lane;
        const int q = fast_div(k, inv[l]);
        const int gx = b[l].x0 + (XMAJOR ? q : k - q * b[l].w), gy = b[l].y0 + (XMAJOR ? k - q * b[l].h : q);
        // (through the L2: in the persistent layer kernel other CTAs rewrite these planes between two barriers)
        v[l][t] = (plane[l] != nullptr && k < n && gx >= 0 && gx < tw[l] && gy >= 0 && gy < th[l]) ? __ldcg(plane[l] + gx + gy * tw[l]) : 0.0f;
      }
    }
    int mine[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) {
      mine[l] = 0;
#pragma unroll
      for (int t = 0; t < VM; t++) mine[l] += __popc(__ballot_sync(0xffffffffu, v[l][t] != 0.0f));
    }
    __syncthreads();  // (the readers of the previous counts are done)
    if (lane == 0) {
#pragma unroll
      for (int l = 0; l < NL; l++) s_cnt[l][warp + 1] = mine[l];
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < NL; l++) {
      int pos = n_out[l], tot = 0;
      for (int g = 0; g < GRID_TY; g++) { const int c = s_cnt[l][g + 1]; tot += c; if (g < warp) pos += c; }
#pragma unroll
      for (int t = 0; t < VM; t++) {
        const unsigned m = __ballot_sync(0xffffffffu, v[l][t] != 0.0f);
        if (v[l][t] != 0.0f) {
          const int k = r0 + (warp * VM + t) * 32 + lane;
          const int q = fast_div(k, inv[l]);
          emit(l, pos + __popc(m & ((1u << lane) - 1u)), b[l].x0 + (XMAJOR ? q : k - q * b[l].w), b[l].y0 + (XMAJOR ? k - q * b[l].h : q), v[l][t]);
        }
        pos += __popc(m);
      }
      n_out[l] += tot;
    }
  }
  __syncthreads();
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

// list_dot in two halves, so that the weight loads of the first hits are in flight while something else runs:
// list_hits marks the hits among the first 64 entries and loads the weights of the first NP of them.
template <int NP> struct ListHits { unsigned long long rest; float w[NP], d[NP]; };
template <int NP>
__device__ __forceinline__ ListHits<NP> list_hits(const WinDev& w, const int2* __restrict__ ent, int n, int ux, int uy, const float* __restrict__ wrow) {
  ListHits<NP> H;
  H.rest = 0ull;
#pragma unroll
  for (int j = 0; j < NP; j++) { H.w[j] = 0.0f; H.d[j] = 0.0f; }
  if (w.r < 0) return H;
  const int cx = w.cx[ux], cy = w.cy[uy], r = w.r;
  const unsigned r2 = 2u * (unsigned)r;
  const int kn = min(64, n);
  unsigned long long hits = 0ull;
  for (int k = 0; k < kn; k++) {
    const int c = ent[k].x, dx = (c & 0xffff) - cx, dy = (c >> 16) - cy;
    const bool in = (unsigned)(dx + r) <= r2 && (unsigned)(dy + r) <= r2 && !(w.excl && (dx | dy) == 0);
    hits |= (unsigned long long)in << k;
  }
#pragma unroll
  for (int j = 0; j < NP; j++)
    if (hits) {
      const int k = __ffsll((long long)hits) - 1;
      hits &= hits - 1ull;
      const int2 e = ent[k];
      const int tx = e.x & 0xffff, ty = e.x >> 16;
      H.w[j] = wrow[(w.absx ? tx : tx - cx + r) * w.dyn + (w.absy ? ty : ty - cy + r)];
      H.d[j] = __int_as_float(e.y);
    }
  H.rest = hits;
  return H;
}
template <int NP>
__device__ __forceinline__ float list_finish(const WinDev& w, const ListHits<NP>& H, const int2* __restrict__ ent, int n, int ux, int uy,
                                             const float* __restrict__ wrow, float sum) {
  if (w.r < 0) return sum;
#pragma unroll
  for (int j = 0; j < NP; j++) sum += H.d[j] * H.w[j];  // (an absent hit adds +0)
  if (H.rest) {  // more than NP hits among the first 64 entries
    const int cx = w.cx[ux], cy = w.cy[uy], r = w.r;
    for (unsigned long long hits = H.rest; hits; hits &= hits - 1ull) {
      const int2 e = ent[__ffsll((long long)hits) - 1];
      const int tx = e.x & 0xffff, ty = e.x >> 16;
      sum += __int_as_float(e.y) * wrow[(w.absx ? tx : tx - cx + r) * w.dyn + (w.absy ? ty : ty - cy + r)];
    }
  }
  if (n > 64) sum = list_dot(w, ent + 64, n - 64, ux, uy, wrow, sum);
  return sum;
}

// ---- sdr/RSDR.cpp:123-133: a = -theta + sum_ff x*w + sum_rec sPrev*w.  One WARP per row segment of 32 units.
// dense_ff: the visible plane is dense (layer 0): its 32 weight rows are staged by one bulk copy and the window box
// by the warp.  Otherwise (the visible plane is the binary state of the layer below) the feed-forward sum is sparse too.
// Every global load the segment needs -- the weight rows, the input box, the box of previous states -- is issued
// before the first one is consumed, and the recurrent weights of the first hits are fetched under the dense sum.
// smem: [32*pitch slab | ff box] (dense) , list entries, mbarrier
// One 3-D tiled TMA load (cp.async.bulk.tensor, SASS UTMALDG): box (bw x bh x 1) of the [agent][y][x] input array at
// (x0, y0, a) into shared memory, rows bw floats apart; cells outside the grid arrive as zeros (no index tests, no
// per-cell loads: the halo is the TMA unit's out-of-bounds fill).
__device__ __forceinline__ void tma_load_box(void* dst_smem, const void* tmap, int x0, int y0, int a, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst_smem)),
               "l"(tmap), "r"(x0), "r"(y0), "r"(a), "r"(smem_u32(bar))
               : "memory");
}
// tmap_bw > 0: the input box is brought in by the TMA unit through `tmap` (rows tmap_bw floats apart, tmap_bw x tmap_bh cells)
template <int DYN>
__global__ void __launch_bounds__(32) k_g_kw_activate(EngineDev P, LayerDev L, int l, int dense_ff, int slab_floats, int box_floats, int cap,
                                                      const __grid_constant__ CUtensorMap tmap, int tmap_bw, int tmap_bh) {
  extern __shared__ __align__(128) float sm[];
  constexpr int NB = 20;  // box cells per lane per trip: a 44x13 box (radius 6, 32 units) is one trip
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
  float* xbox = sm;                 // (first: the tensor load wants 128-byte alignment)
  float* slab = sm + box_floats;
  int2* ent = reinterpret_cast<int2*>(slab + slab_floats);
  uint64_t* bar = reinterpret_cast<uint64_t*>(ent + cap);
  const int nun = min(32, L.hw - T.ux0);
  const WinDev& wf = L.ff;
  const WinDev& wr = L.rec;
  if (dense_ff) {
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
  }
  Box bf = window_box(wf, T);
  if (dense_ff) {
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(nun * wf.row_pitch) * 4u;
      mbar_expect_tx(bar, bytes + (tmap_bw > 0 ? (uint32_t)(tmap_bw * tmap_bh) * 4u : 0u));
      tma_load_1d(slab, B + wf.w_off + (long long)(T.ux0 + uy * L.hw) * wf.row_pitch, bytes, bar);
      if (tmap_bw > 0) tma_load_box(xbox, &tmap, bf.x0 & ~3, bf.y0, a, bar);
    }
    // (the innermost coordinate of a tiled load must be a multiple of 16 bytes -- anything else faults with "illegal
    // instruction" on this GPU: the box starts at the multiple of four cells at or below x0 and is up to three wider)
    if (tmap_bw > 0) { bf.x0 &= ~3; bf.w = tmap_bw; }  // the staged rows are tmap_bw apart
  }
  float sum = T.ok ? -B[L.thr + T.unit] : 0.0f;
  const Box br = wr.r >= 0 ? window_box(wr, T) : Box{0, 0, 0, 0};
  const float* sprev = B + L.state_prev;
  // non-zero cells of a box, x-major, from a plane: all loads of a trip first, then the ballots
  auto compact = [&](const float* __restrict__ plane, const Box& b, int tw, int th) -> int {
    const int n = b.w * b.h;
    const float inv_h = 1.0f / (float)b.h;
    int base = 0;
    for (int k0 = 0; k0 < n; k0 += 32 * NB) {
      float v[NB];
#pragma unroll
      for (int t = 0; t < NB; t++) {
        const int k = k0 + 32 * t + lane, bx = fast_div(k, inv_h), gx = b.x0 + bx, gy = b.y0 + k - bx * b.h;
        v[t] = (k < n && gx >= 0 && gx < tw && gy >= 0 && gy < th) ? plane[gx + gy * tw] : 0.0f;
      }
#pragma unroll
      for (int t = 0; t < NB; t++) {
        const unsigned m = __ballot_sync(0xffffffffu, v[t] != 0.0f);
        if (m) {
          if (v[t] != 0.0f) {
            const int k = k0 + 32 * t + lane, bx = fast_div(k, inv_h);
            ent[base + __popc(m & ((1u << lane) - 1u))] = make_int2((b.x0 + bx) | ((b.y0 + k - bx * b.h) << 16), __float_as_int(v[t]));
          }
          base += __popc(m);
        }
      }
    }
    __syncwarp();
    return base;
  };
  if (dense_ff) {
    // the window box of the segment (coalesced rows, zero outside the grid) and the box of previous states: the loads
    // of both are in flight together with the bulk copy of the rows
    const int n = tmap_bw > 0 ? 0 : bf.w * bf.h;  // (nothing to stage by hand when the TMA unit brings the box)
    const float inv_w = 1.0f / (float)bf.w;
    float xv[NB];
#pragma unroll
    for (int t = 0; t < NB; t++) {
      const int k = 32 * t + lane, by = fast_div(k, inv_w), gx = bf.x0 + k - by * bf.w, gy = bf.y0 + by;
      xv[t] = (k < n && gx >= 0 && gx < wf.tw && gy >= 0 && gy < wf.th) ? x[gx + gy * wf.tw] : 0.0f;
    }
    const int n_rec = (wr.r >= 0) ? compact(sprev, br, wr.tw, wr.th) : 0;
#pragma unroll
    for (int t = 0; t < NB; t++) { const int k = 32 * t + lane; if (k < n) xbox[k] = xv[t]; }
    for (int k0 = 32 * NB; k0 < n; k0 += 32 * NB) {  // (larger boxes: the rest, trip by trip)
#pragma unroll
      for (int t = 0; t < NB; t++) {
        const int k = k0 + 32 * t + lane, by = fast_div(k, inv_w), gx = bf.x0 + k - by * bf.w, gy = bf.y0 + by;
        xv[t] = (k < n && gx >= 0 && gx < wf.tw && gy >= 0 && gy < wf.th) ? x[gx + gy * wf.tw] : 0.0f;
      }
#pragma unroll
      for (int t = 0; t < NB; t++) { const int k = k0 + 32 * t + lane; if (k < n) xbox[k] = xv[t]; }
    }
    __syncwarp();
    const float* rrow = B + wr.w_off + (long long)(T.ok ? T.unit : 0) * wr.row_pitch;
    const ListHits<8> H = (T.ok && wr.r >= 0) ? list_hits<8>(wr, ent, n_rec, T.ux, uy, rrow) : ListHits<8>{};
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
      if (wr.r >= 0) sum = list_finish<8>(wr, H, ent, n_rec, T.ux, uy, rrow, sum);
      B[L.act + T.unit] = sum;
    }
    return;
  }
  if (wf.r >= 0) {
    const int n = compact(x, bf, wf.tw, wf.th);
    if (T.ok) sum = list_dot(wf, ent, n, T.ux, uy, B + wf.w_off + (long long)T.unit * wf.row_pitch, sum);
    __syncwarp();
  }
  if (wr.r >= 0) {
    const int n = compact(sprev, br, wr.tw, wr.th);
    if (T.ok) sum = list_dot(wr, ent, n, T.ux, uy, B + wr.w_off + (long long)T.unit * wr.row_pitch, sum);
  }
  if (T.ok) B[L.act + T.unit] = sum;
}

// ---- sdr/RSDR.cpp:135-145, :148-163: rank among the lateral neighbours: s = #{j in lat(i): a_j >= a_i} < sparsity*|lat(i)|.
// The count only matters up to that limit (1-2 for the usual 1 %), and it can only grow: a lane scans the first two
// columns of its window alone -- after them ~95 % of the units are known losers -- and the few lanes still in the
// race get the rest of their window counted by the whole warp, 32 cells per ballot, stopping at the limit.
__global__ void __launch_bounds__(GRID_TX* GRID_TY) k_g_kw_inhibit(EngineDev P, LayerDev L, int l, long long src, long long dst, long long dst_next) {
  extern __shared__ float sm[];
  const GridTile T = grid_tile(blockIdx.x, L.hw, L.row0, L.row1);
  float* B = blob_of(P, T.a);
  const float* act = B + src;
  const WinDev& w = L.lat;
  const Box b = window_box(w, T);
  stage_box(sm, b, w.tw, w.th, [&](int t) { return act[t]; });
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float mine = 0.0f, lim = 0.0f;
  int cnt = 0, col = 0, sx1 = -1, ny = 0, bx = 0, by = 0, selfcol = 0;  // bx, by: box coordinates of the window's column `col`, row sy0
  if (T.ok) {
    const SlotBox sb = slot_box(w, T.ux, T.uy);
    mine = sm[(T.ux - b.x0) + (T.uy - b.y0) * b.w];
    lim = L.sparsity * (float)slot_count(w, sb);
    col = sb.sx0; sx1 = sb.sx1; ny = sb.sy1 - sb.sy0 + 1;
    by = slot_ty(w, sb, sb.sy0) - b.y0;
    selfcol = w.absx ? sb.cx : w.r;
    constexpr int NC1 = 2;
    for (int c = 0; c < NC1 && col <= sx1; c++, col++) {
      const float* sp = sm + (slot_tx(w, sb, col) - b.x0) + by * b.w;
      if (ny == 13) {
#pragma unroll
        for (int q = 0; q < 13; q++) cnt += sp[q * b.w] >= mine ? 1 : 0;
      } else {
#pragma unroll 4
        for (int q = 0; q < ny; q++) cnt += sp[q * b.w] >= mine ? 1 : 0;
      }
    }
    bx = slot_tx(w, sb, col) - b.x0;
  }
  // the unit itself is inside its window (mine >= mine) but not in the lateral list: it counts once the scan has passed its column
  const bool lost = T.ok && (float)(cnt - (selfcol < col ? 1 : 0)) >= lim;
  for (unsigned todo = __ballot_sync(0xffffffffu, T.ok && !lost && col <= sx1); todo; todo &= todo - 1) {
    const int s = __ffs(todo) - 1;
    const float mine_s = __shfl_sync(0xffffffffu, mine, s), lim_s = __shfl_sync(0xffffffffu, lim, s);
    const int ny_s = __shfl_sync(0xffffffffu, ny, s), bx_s = __shfl_sync(0xffffffffu, bx, s), by_s = __shfl_sync(0xffffffffu, by, s);
    const int cols_s = __shfl_sync(0xffffffffu, sx1 - col + 1, s), cnt_s = __shfl_sync(0xffffffffu, cnt, s);
    const int rem = cols_s * ny_s;
    int c = 0;
    for (int j0 = 0; j0 < rem; j0 += 32) {
      const int j = j0 + lane, jx = j / ny_s, jy = j - jx * ny_s;
      const bool ge = j < rem && sm[(bx_s + jx) + (by_s + jy) * b.w] >= mine_s;
      c += __popc(__ballot_sync(0xffffffffu, ge));
      if ((float)(cnt_s + c - 1) >= lim_s) break;  // (at most one of them is the unit itself)
    }
    if (lane == s) { cnt += c; col = sx1 + 1; }
  }
  if (!T.ok) return;
  const float st = (!lost && (float)(cnt - 1) < lim) ? 1.0f : 0.0f;
  B[dst + T.unit] = st;
  if (dst_next >= 0) B[dst_next + T.unit] = st;
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
  __shared__ int s_cnt[1][GRID_TY + 1];
  const float* const planes[1] = {drive};
  const Box boxes[1] = {Box{bx0, by0, bw, bh}};
  const int tws[1] = {w.uw}, ths[1] = {w.uh};
  int cnt[1];
  compact_boxes_cta<1, false>(planes, boxes, tws, ths, cnt, s_cnt, [&](int, int q, int ux, int uy, float d) {
    l_unit[q] = ux + uy * w.uw; l_ctr[q] = w.cx[ux] | (w.cy[uy] << 16); l_d[q] = d;
  });
  if (!T.ok) return;
  const int tx = T.ux, ty = T.uy, n = cnt[0], r = w.r;
  const unsigned r2 = 2u * (unsigned)r;
  const float* wp = B + w.w_off;
  float sum = 0.0f;
  // list entries whose window contains this cell: marked 64 entries at a time, then the weights of four hits are
  // loaded together and added in list order (neo/SparseCoder.cpp:242-259: w*state*multiplier)
  for (int k0 = 0; k0 < n; k0 += 64) {
    unsigned long long hits = 0ull;
    const int kn = min(64, n - k0);
    for (int k = 0; k < kn; k++) {
      const int c = l_ctr[k0 + k], dx = tx - (c & 0xffff), dy = ty - (c >> 16);
      const bool in = (unsigned)(dx + r) <= r2 && (unsigned)(dy + r) <= r2 && !(w.excl && (dx | dy) == 0);
      hits |= (unsigned long long)in << k;
    }
    while (hits) {
      float wv[4], dv[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        wv[j] = 0.0f; dv[j] = 0.0f;
        if (hits) {
          const int k = k0 + __ffsll((long long)hits) - 1;
          hits &= hits - 1ull;
          const int c = l_ctr[k];
          const int sx = w.absx ? tx : tx - (c & 0xffff) + r, sy = w.absy ? ty : ty - (c >> 16) + r;
          wv[j] = wp[widx(w, sx * w.dyn + sy, l_unit[k])];
          dv[j] = l_d[k];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; j++) sum += scaled ? wv[j] * dv[j] * mult : wv[j] * dv[j];  // (an absent hit adds +0)
    }
  }
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
  __shared__ int s_cnt[4][GRID_TY + 1];
  const GridTile T = grid_tile(blockIdx.x, L.hw, L.row0, L.row1);
  float* B = blob_of(P, T.a);
  const bool top = l == P.L - 1 || L.fb.r < 0;
  int2* l_state = reinterpret_cast<int2*>(sm);
  int2* l_sprev = l_state + cap_pred;
  int2* l_up = l_sprev + cap_pred;
  int2* l_upprev = l_up + cap_fb;
  const Box bp = window_box(L.pred, T), bfb = top ? Box{0, 0, 0, 0} : window_box(L.fb, T);
  int cnt[4] = {0, 0, 0, 0};
  int2* const lists[4] = {l_state, l_sprev, l_up, l_upprev};
  auto emit = [&](int li, int q, int gx, int gy, float v) { lists[li][q] = make_int2(gx | (gy << 16), __float_as_int(v)); };
  if (top) {
    const float* const planes[2] = {B + L.state, learn ? B + L.state_prev : nullptr};
    const Box boxes[2] = {bp, bp};
    const int tws[2] = {L.pred.tw, L.pred.tw}, ths[2] = {L.pred.th, L.pred.th};
    int c2[2];
    compact_boxes_cta<2, true>(planes, boxes, tws, ths, c2, s_cnt, emit);
    cnt[0] = c2[0]; cnt[1] = c2[1];
  } else {
    const float* const planes[4] = {B + L.state, learn ? B + L.state_prev : nullptr, B + up_p_state, learn ? B + up_p_state_prev : nullptr};
    const Box boxes[4] = {bp, bp, bfb, bfb};
    const int tws[4] = {L.pred.tw, L.pred.tw, L.fb.tw, L.fb.tw}, ths[4] = {L.pred.th, L.pred.th, L.fb.th, L.fb.th};
    compact_boxes_cta<4, true>(planes, boxes, tws, ths, cnt, s_cnt, emit);
  }
  const int n_state = cnt[0], n_sprev = cnt[1], n_up = cnt[2], n_upprev = cnt[3];
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

// ---- sdr/RSDR.cpp:241-266: winners only; theta for every unit.  The tile's winners are listed in shared memory and
// dealt to its warps one each: the 32 lanes share a winner's connections (the rule is element-wise and a winner's
// row is contiguous), every load of a family in flight before the first store.
__global__ void __launch_bounds__(GRID_TX* GRID_TY) k_g_kw_learn(EngineDev P, LayerDev L, int l) {
  __shared__ int s_n;
  __shared__ int s_who[GRID_TX * GRID_TY];
  __shared__ float s_kf[GRID_TX * GRID_TY], s_kr[GRID_TX * GRID_TY];
  const GridTile T = grid_tile(blockIdx.x, L.hw, L.row0, L.row1);
  float* B = blob_of(P, T.a);
  const float* x = vis_input_of(P, L, l, T.a);
  const float* vrec = B + L.vis_recon;
  const float* sprev = B + L.state_prev;
  const float* hrec = B + L.recon;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const float learn = T.ok ? B[L.state + T.unit] : 0.0f;
  if (learn > 0.0f) {
    const float att = B[L.p_mod + T.unit];
    const int q = atomicAdd(&s_n, 1);
    s_who[q] = T.ux | (T.uy << 16);
    s_kf[q] = L.learn_ff * att * learn; s_kr[q] = L.learn_rec * att * learn;
  }
  if (T.ok) B[L.thr + T.unit] += L.learn_threshold * (learn - L.sparsity);
  __syncthreads();
  const int nwin = s_n;
  for (int q = warp; q < nwin; q += GRID_TY) {
    const int ux = s_who[q] & 0xffff, uy = s_who[q] >> 16, unit = ux + uy * L.hw;
    for (int fam = 0; fam < 2; fam++) {
      const WinDev& w = fam ? L.rec : L.ff;
      if (w.r < 0) continue;
      float* wp = B + w.w_off;
      const float* a = fam ? sprev : x;
      const float* b = fam ? hrec : vrec;
      const float k = fam ? s_kr[q] : s_kf[q];
      const SlotBox sb = slot_box(w, ux, uy);
      const float inv_dyn = 1.0f / (float)w.dyn;
      constexpr int NJ = 6;  // 192 slots per trip: a radius-6 window (169 slots) is one trip of independent loads
      for (int s0 = 0; s0 < w.nslots; s0 += 32 * NJ) {
        float wv[NJ], av[NJ], bv[NJ];
        bool ok[NJ];
#pragma unroll
        for (int j = 0; j < NJ; j++) {
          const int s = s0 + 32 * j + lane, sx = fast_div(s, inv_dyn), sy = s - sx * w.dyn;
          const int tx = slot_tx(w, sb, sx), ty = slot_ty(w, sb, sy);
          ok[j] = s < w.nslots && sx >= sb.sx0 && sx <= sb.sx1 && sy >= sb.sy0 && sy <= sb.sy1 && !(w.excl && tx == sb.cx && ty == sb.cy);
          wv[j] = 0.0f; av[j] = 0.0f; bv[j] = 0.0f;
          if (ok[j]) { const int t = tx + ty * w.tw; wv[j] = wp[widx(w, s, unit)]; av[j] = a[t]; bv[j] = b[t]; }
        }
#pragma unroll
        for (int j = 0; j < NJ; j++)
          if (ok[j]) wp[widx(w, s0 + 32 * j + lane, unit)] = wv[j] + k * (av[j] - bv[j]);
      }
    }
  }
}

} // namespace bidi
