// bidi_persist.cuh -- the iterative coder of ONE large layer of a single agent (sdr::IRSDR::activate,
// sdr/IRSDR.cpp:125-216: 30 settle + 5 measure iterations of sweep -> spikes -> reconstruction) as ONE cooperative
// launch instead of 2 x 35 + 1 dependent launches.
//
// A single agent's large layer (C4: 96x96 units, radius 6, 18.6 MB of feed-forward | recurrent | lateral weights)
// is far too big for one CTA and its iterations are strictly sequential, so the per-phase kernels of bidi_tiles.cuh
// / bidi_grid.cuh pay a launch, a drain and a re-read of every weight from the L2 per phase.  Here the layer is cut
// into one tile of hidden units per CTA (all CTAs co-resident), and
//   * a CTA copies its tile's weights of all three families into SHARED MEMORY once and keeps them there for all
//     35 iterations, as w[slot][unit of the tile] -- lanes = adjacent units, conflict-free;
//   * a sweep stages the three source boxes of the tile (visible error, hidden error, previous spikes) and runs
//     two sequential sums per unit (excitation, inhibition) on separate threads, slot order = the reference's list
//     order; activation and state stay in registers from the first iteration to the last;
//   * only unit-state planes cross CTAs (spikes, reconstructions), through the L2 (ld.global.cg), with a grid-wide
//     barrier (one atomic per CTA, acquire spin) in place of the launch boundary;
//   * the gather-form reconstruction (ascending unit index = the order of the reference's scatter) walks the
//     compacted list of the spiking units of the tile's candidate box, as bidi_grid.cuh does.
// Same operations in the same order as k_g_ic_sweep / k_g_reconstruct: the bits are theirs.
#pragma once
#include <cooperative_groups.h>
#include "bidi_grid.cuh"

namespace bidi {

struct PersistShape {
  int on;
  int htx, hty, hnx, hny;        // hidden tile size, tile counts
  int vtx, vty, vnx, vny;        // visible tiles: which CTA owns which part of the visible reconstruction
  int ut;                        // units per hidden tile = row pitch of the weight slab
  int box_ff, box_rec, box_lat;  // floats of the staged source boxes
  int cap_ff, cap_rec;           // cells of the candidate boxes of the two gathers
  int blocks;                    // CTAs (one agent)
  int dyn;                       // common window height of the three families when it is one of the unrolled sizes, else 0
  long long smem_floats;
};
#define BIDI_PERSIST_THREADS 256

// grid-wide barrier over `nblocks` co-resident CTAs: ctr counts arrivals since the launch (zeroed before it)
__device__ __forceinline__ void persist_sync(unsigned* ctr, unsigned& target, unsigned nblocks) {
#ifdef BIDI_PERSIST_CG
  cooperative_groups::this_grid().sync();
  return;
#endif
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    // arrival = a release at GPU scope: with the block barrier above it orders every thread's stores before it
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
  }
  __syncthreads();
}

// sum += box(target) * w[slot][u] over the whole slot box of one family, list order; weights from the shared slab
template <int DYN>
__device__ __forceinline__ float persist_dot(const WinDev& w, const float* __restrict__ s0, int bw, const float* __restrict__ wsm, int ut, float sum) {
  if (w.r < 0) return sum;
  if (DYN > 0) {
    // one window column (DYN slots) per trip, software-pipelined: the 2 x DYN shared-memory loads of the next column are
    // issued before the DYN chained adds of the current one (a chain thread has its scheduler nearly to itself, so
    // nothing else hides the load latency); the additions keep the list order
    constexpr int DN = DYN > 0 ? DYN : 1;
    float sa[DN], wa[DN], sb[DN], wb[DN];
    const int n = w.dxn;
    auto ld = [&](float* s, float* ww, int sx) {
      const float* sp = s0 + sx;
      const float* wp = wsm + (size_t)sx * DYN * ut;
#pragma unroll
      for (int sy = 0; sy < DYN; sy++) { s[sy] = sp[sy * bw]; ww[sy] = wp[sy * ut]; }
    };
    auto eat = [&](const float* s, const float* ww) {
#pragma unroll
      for (int sy = 0; sy < DYN; sy++) sum += s[sy] * ww[sy];
    };
    if (n > 0) ld(sa, wa, 0);
    int sx = 0;
    for (; sx + 2 <= n; sx += 2) {
      ld(sb, wb, sx + 1);
      eat(sa, wa);
      if (sx + 2 < n) ld(sa, wa, sx + 2);
      eat(sb, wb);
    }
    if (sx < n) eat(sa, wa);
  } else {
    for (int sx = 0; sx < w.dxn; sx++)
      for (int sy = 0; sy < w.dyn; sy++) sum += s0[sx + sy * bw] * wsm[(size_t)(sx * w.dyn + sy) * ut];
  }
  return sum;
}

// the window box of a tx x ty tile of units at (ux0, uy0) (clipped at the grid's end)
__device__ __forceinline__ Box tile_box(const WinDev& w, int ux0, int uy0, int tx, int ty) {
  Box b;
  if (w.r < 0) return Box{0, 0, 0, 0};
  if (w.absx) { b.x0 = 0; b.w = w.tw; }
  else { b.x0 = w.cx[ux0] - w.r; b.w = w.cx[min(ux0 + tx, w.uw) - 1] + w.r - b.x0 + 1; }
  if (w.absy) { b.y0 = 0; b.h = w.th; }
  else { b.y0 = w.cy[uy0] - w.r; b.h = w.cy[min(uy0 + ty, w.uh) - 1] + w.r - b.y0 + 1; }
  return b;
}
// the units whose windows can contain a cell of the tx x ty tile of targets at (x0, y0): a box of the unit grid
__device__ __forceinline__ Box cand_box(const WinDev& w, int x0, int y0, int tx, int ty) {
  if (w.r < 0) return Box{0, 0, 0, 0};
  const int x1 = min(x0 + tx, w.tw) - 1, y1 = min(y0 + ty, w.th) - 1;
  int bx0 = w.uw, bx1 = -1, by0 = w.uh, by1 = -1;
  for (int t = x0; t <= x1; t++) { bx0 = min(bx0, w.gx_lo[t]); bx1 = max(bx1, w.gx_hi[t]); }
  for (int t = y0; t <= y1; t++) { by0 = min(by0, w.gy_lo[t]); by1 = max(by1, w.gy_hi[t]); }
  return Box{bx0, by0, max(0, bx1 - bx0 + 1), max(0, by1 - by0 + 1)};
}

template <int DYN>
__global__ void __launch_bounds__(BIDI_PERSIST_THREADS, 1) k_ic_layer_persist(EngineDev P, LayerDev L, int l, PersistShape S, unsigned* ctr) {
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_cnt[2][GRID_TY + 1];
  const int tid = threadIdx.x, b = blockIdx.x;
  const int a = 0;  // one agent per launch
  float* B = blob_of(P, a);
  const float* x = vis_input_of(P, L, l, a);
  const int ut = S.ut;
  const bool has_rec = L.rec.r >= 0;
  const int nsf = L.ff.nslots, nsr = has_rec ? L.rec.nslots : 0, nsl = L.lat.r >= 0 ? L.lat.nslots : 0;
  float* wsm = sm;                                       // [nsf + nsr + nsl][ut]
  float* box_f = wsm + (size_t)(nsf + nsr + nsl) * ut;   // visible errors over the tile's feed-forward box
  float* box_r = box_f + S.box_ff;                       // hidden errors
  float* box_l = box_r + S.box_rec;                      // previous spikes
  float* inh_s = box_l + S.box_lat;                      // [ut] inhibition sums, handed to the excitation threads
  int* lf_unit = reinterpret_cast<int*>(inh_s + ut);     // gather lists: unit index, window centre, drive
  int* lf_ctr = lf_unit + S.cap_ff;
  float* lf_d = reinterpret_cast<float*>(lf_ctr + S.cap_ff);
  int* lr_unit = reinterpret_cast<int*>(lf_d + S.cap_ff);
  int* lr_ctr = lr_unit + S.cap_rec;
  float* lr_d = reinterpret_cast<float*>(lr_ctr + S.cap_rec);
  int* ctab = reinterpret_cast<int*>(lr_d + S.cap_rec);  // window centres: ff.cx[hw], ff.cy[hh], rec.cx[hw], rec.cy[hh]
  for (int k = tid; k < 2 * (L.hw + L.hh); k += BIDI_PERSIST_THREADS) {
    const int q = k < L.hw + L.hh ? k : k - (L.hw + L.hh);
    const WinDev& w = k < L.hw + L.hh ? L.ff : L.rec;
    ctab[k] = w.r < 0 ? 0 : (q < L.hw ? w.cx[q] : w.cy[q - L.hw]);
  }

  // ---- this CTA's tiles
  const bool has_h = b < S.hnx * S.hny, has_v = b < S.vnx * S.vny;
  const int hx0 = (b % S.hnx) * S.htx, hy0 = (b / S.hnx) * S.hty;
  const int vx0 = (b % S.vnx) * S.vtx, vy0 = (b / S.vnx) * S.vty;
  // sweep roles: thread t < ut runs the excitation sum of tile unit t and owns its state; ut <= t < 2 ut the inhibition sum
  const int role = tid / ut, tu = tid - role * ut;
  const int lux = tu % S.htx, luy = tu / S.htx, ux = hx0 + lux, uy = hy0 + luy;
  const bool unit_ok = has_h && role < 2 && ux < L.hw && uy < L.hh;
  const int unit = ux + uy * L.hw;

  // ---- the tile's weights: HBM/L2 -> shared memory, once
  if (has_h) {
    const int total = (nsf + nsr + nsl) * ut;
    const float inv_ut = 1.0f / (float)ut, inv_tx = 1.0f / (float)S.htx;
    for (int k0 = 0; k0 < total; k0 += 8 * BIDI_PERSIST_THREADS) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int k = k0 + j * BIDI_PERSIST_THREADS + tid;
        const int s = fast_div(k, inv_ut), u = k - s * ut, ly = fast_div(u, inv_tx);
        const int gx = hx0 + u - ly * S.htx, gy = hy0 + ly;
        v[j] = 0.0f;
        if (k < total && gx < L.hw && gy < L.hh) {
          const int g = gx + gy * L.hw;
          v[j] = s < nsf ? B[L.ff.w_off + (long long)s * L.nh + g]
                         : (s < nsf + nsr ? B[L.rec.w_off + (long long)(s - nsf) * L.nh + g] : B[L.lat.w_off + (long long)(s - nsf - nsr) * L.nh + g]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; j++) { const int k = k0 + j * BIDI_PERSIST_THREADS + tid; if (k < total) wsm[k] = v[j]; }
    }
  }
  const Box bf = has_h ? tile_box(L.ff, hx0, hy0, S.htx, S.hty) : Box{0, 0, 0, 0};
  const Box br = has_h && has_rec ? tile_box(L.rec, hx0, hy0, S.htx, S.hty) : Box{0, 0, 0, 0};
  const Box bl = has_h ? tile_box(L.lat, hx0, hy0, S.htx, S.hty) : Box{0, 0, 0, 0};
  const Box cf = has_v ? cand_box(L.ff, vx0, vy0, S.vtx, S.vty) : Box{0, 0, 0, 0};
  const Box cr = has_h && has_rec ? cand_box(L.rec, hx0, hy0, S.htx, S.hty) : Box{0, 0, 0, 0};
  const float* vrec = B + L.vis_recon;
  const float* sprev = B + L.state_prev;
  const float* hrec = B + L.recon;
  const float* spk_prev = B + L.spike_prev;
  // (source offsets of this thread's unit inside the three boxes)
  const int of = unit_ok ? (L.ff.absx ? 0 : L.ff.cx[ux] - L.ff.r - bf.x0) + (L.ff.absy ? 0 : L.ff.cy[uy] - L.ff.r - bf.y0) * bf.w : 0;
  const int orr = unit_ok && has_rec ? (L.rec.absx ? 0 : L.rec.cx[ux] - L.rec.r - br.x0) + (L.rec.absy ? 0 : L.rec.cy[uy] - L.rec.r - br.y0) * br.w : 0;
  const int ol = unit_ok && L.lat.r >= 0 ? (L.lat.absx ? 0 : L.lat.cx[ux] - L.lat.r - bl.x0) + (L.lat.absy ? 0 : L.lat.cy[uy] - L.lat.r - bl.y0) * bl.w : 0;
  const float thr = unit_ok && role == 0 ? B[L.thr + unit] : 0.0f;
  float act = 0.0f, st = 0.0f;
  unsigned target = 0;
  const unsigned nblocks = gridDim.x;
  const int settle = L.iter_settle, total_it = L.iter_settle + L.iter_measure;
  const float inv_measure = 1.0f / (float)L.iter_measure;

#ifdef BIDI_PERSIST_TIMING
  unsigned long long dbg_compact = 0, dbg_nf = 0, dbg_nr = 0;
#endif
  // gather-form reconstruction of this CTA's visible tile and hidden tile from drive[] (spikes, or the states at the end)
  auto reconstruct = [&](const float* drive, bool copy_spike) {
    const float* const planes[2] = {has_v ? drive : nullptr, (has_h && has_rec) ? drive : nullptr};
    const Box boxes[2] = {cf, cr};
    const int tws[2] = {L.ff.uw, L.rec.uw}, ths[2] = {L.ff.uh, L.rec.uh};
    int cnt[2];
#ifdef BIDI_PERSIST_TIMING
    unsigned long long tc0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tc0));
#endif
    compact_boxes_cta<2, false>(planes, boxes, tws, ths, cnt, s_cnt, [&](int li, int q, int gx, int gy, float d) {
      if (li == 0) { lf_unit[q] = gx + gy * L.ff.uw; lf_ctr[q] = ctab[gx] | (ctab[L.hw + gy] << 16); lf_d[q] = d; }
      else { lr_unit[q] = gx + gy * L.rec.uw; lr_ctr[q] = ctab[L.hw + L.hh + gx] | (ctab[2 * L.hw + L.hh + gy] << 16); lr_d[q] = d; }
    });
#ifdef BIDI_PERSIST_TIMING
    { unsigned long long tc1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tc1)); dbg_compact += tc1 - tc0; dbg_nf += cnt[0]; dbg_nr += cnt[1]; }
#endif
    // threads [0, 128): visible cells; [128, 256): hidden units
    const bool vis = tid < BIDI_PERSIST_THREADS / 2;
    const int t = vis ? tid : tid - BIDI_PERSIST_THREADS / 2;
    const WinDev& w = vis ? L.ff : L.rec;
    const int tx = vis ? vx0 + t % S.vtx : hx0 + t % S.htx, ty = vis ? vy0 + t / S.vtx : hy0 + t / S.htx;
    const bool ok = vis ? (has_v && t < S.vtx * S.vty && tx < L.vw && ty < L.vh) : (has_h && t < ut && tx < L.hw && ty < L.hh);
    if (!ok) return;
    const int cell = tx + ty * w.tw;
    float sum = 0.0f;
    if (w.r >= 0) {
      const int n = vis ? cnt[0] : cnt[1], r = w.r;
      const unsigned r2 = 2u * (unsigned)r;
      const int* l_unit = vis ? lf_unit : lr_unit;
      const int* l_ctr = vis ? lf_ctr : lr_ctr;
      const float* l_d = vis ? lf_d : lr_d;
      const float* wp = B + w.w_off;
      for (int k0 = 0; k0 < n; k0 += 64) {
        unsigned long long hits = 0ull;
        const int kn = min(64, n - k0);
        for (int k = 0; k < kn; k++) {
          const int c = l_ctr[k0 + k], dx = tx - (c & 0xffff), dy = ty - (c >> 16);
          const bool in = (unsigned)(dx + r) <= r2 && (unsigned)(dy + r) <= r2 && !(w.excl && (dx | dy) == 0);
          hits |= (unsigned long long)in << k;
        }
        while (hits) {
          // sixteen weight loads in flight per round trip to the L2
          constexpr int NG = 16;
          float wv[NG], dv[NG];
#pragma unroll
          for (int j = 0; j < NG; j++) {
            wv[j] = 0.0f; dv[j] = 0.0f;
            if (hits) {
              const int k = k0 + __ffsll((long long)hits) - 1;
              hits &= hits - 1ull;
              const int c = l_ctr[k];
              const int sx = w.absx ? tx : tx - (c & 0xffff) + r, sy = w.absy ? ty : ty - (c >> 16) + r;
              wv[j] = wp[(long long)(sx * w.dyn + sy) * w.U + l_unit[k]];
              dv[j] = l_d[k];
            }
          }
#pragma unroll
          for (int j = 0; j < NG; j++) sum += wv[j] * dv[j];  // (an absent hit adds +0)
        }
      }
    }
    if (vis) B[L.vis_recon + cell] = sum;
    else {
      B[L.recon + cell] = sum;
      if (copy_spike) B[L.spike_prev + cell] = __ldcg(B + L.spike + cell);  // spikePrev = spike (sdr/IRSDR.cpp:170-171)
    }
  };

#ifdef BIDI_PERSIST_TIMING
  unsigned long long tacc[6] = {0, 0, 0, 0, 0, 0}, tprev;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tprev));
#define PT(k) do { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tacc[k] += t_ - tprev; tprev = t_; } while (0)
#else
#define PT(k)
#endif
  __syncthreads();
  PT(0);
  for (int it = 0; it < total_it; it++) {
    // ---- sweep (sdr/IRSDR.cpp:138-168)
    if (has_h) {
      // the three source boxes in ONE pass: every load (through the L2) is issued before the first value is stored
      const int n1 = bf.w * bf.h, n2 = n1 + br.w * br.h, n3 = n2 + bl.w * bl.h;
      const float iw1 = 1.0f / (float)max(1, bf.w), iw2 = 1.0f / (float)max(1, br.w), iw3 = 1.0f / (float)max(1, bl.w);
      for (int k0 = 0; k0 < n3; k0 += 6 * BIDI_PERSIST_THREADS) {
        float v[6];
#pragma unroll
        for (int j = 0; j < 6; j++) {
          const int k = k0 + j * BIDI_PERSIST_THREADS + tid;
          v[j] = 0.0f;
          if (k < n1) {
            const int by = fast_div(k, iw1), gx = bf.x0 + k - by * bf.w, gy = bf.y0 + by;
            if (gx >= 0 && gx < L.ff.tw && gy >= 0 && gy < L.ff.th) { const int t = gx + gy * L.ff.tw; v[j] = x[t] - __ldcg(vrec + t); }
          } else if (k < n2) {
            const int q = k - n1, by = fast_div(q, iw2), gx = br.x0 + q - by * br.w, gy = br.y0 + by;
            if (gx >= 0 && gx < L.rec.tw && gy >= 0 && gy < L.rec.th) { const int t = gx + gy * L.rec.tw; v[j] = sprev[t] - __ldcg(hrec + t); }
          } else if (k < n3) {
            const int q = k - n2, by = fast_div(q, iw3), gx = bl.x0 + q - by * bl.w, gy = bl.y0 + by;
            if (gx >= 0 && gx < L.lat.tw && gy >= 0 && gy < L.lat.th) v[j] = __ldcg(spk_prev + gx + gy * L.lat.tw);
          }
        }
#pragma unroll
        for (int j = 0; j < 6; j++) {
          const int k = k0 + j * BIDI_PERSIST_THREADS + tid;
          if (k < n1) box_f[k] = v[j];
          else if (k < n2) box_r[k - n1] = v[j];
          else if (k < n3) box_l[k - n2] = v[j];
        }
      }
    }
    __syncthreads();
    PT(1);
    float part = 0.0f;
    if (unit_ok) {
      if (role == 0) {
        part = persist_dot<DYN>(L.ff, box_f + of, bf.w, wsm + tu, ut, 0.0f);
        part = persist_dot<DYN>(L.rec, box_r + orr, br.w, wsm + (size_t)nsf * ut + tu, ut, part);
      } else {
        part = persist_dot<DYN>(L.lat, box_l + ol, bl.w, wsm + (size_t)(nsf + nsr) * ut + tu, ut, 0.0f);
        inh_s[tu] = part;
      }
    }
    __syncthreads();
    if (unit_ok && role == 0) {
      float av = (1.0f - L.leak) * act + part - inh_s[tu], spike = 0.0f;
      if (av > thr) { av = 0.0f; spike = 1.0f; }
      if (it >= settle) st += inv_measure * spike;
      act = av;
      B[L.spike + unit] = spike;
      if (it == total_it - 1) { B[L.act + unit] = act; B[L.state + unit] = st; }
    }
    __syncthreads();
    PT(2);
    persist_sync(ctr, target, nblocks);
    PT(3);
    // ---- reconstruction from the spikes (sdr/IRSDR.cpp:218-235); spikePrev = spike
    reconstruct(B + L.spike, true);
    __syncthreads();
    PT(4);
    persist_sync(ctr, target, nblocks);
    PT(5);
  }
#ifdef BIDI_PERSIST_TIMING
  if (tid == 0 && l <= 1)
    printf("persist l=%d b=%d blocks=%d ut=%d: stage %llu chains %llu sync1 %llu recon %llu (compact %llu nf %llu nr %llu) sync2 %llu\n", l, b, (int)gridDim.x, ut,
           tacc[1], tacc[2], tacc[3], tacc[4], dbg_compact, dbg_nf / 35, dbg_nr / 35, tacc[5]);
#endif
  if (total_it == 0 && unit_ok && role == 0) { B[L.act + unit] = 0.0f; B[L.state + unit] = 0.0f; }
  if (total_it == 0) persist_sync(ctr, target, nblocks);
  // ---- reconstructFromStates (sdr/IRSDR.cpp:215)
  reconstruct(B + L.state, false);
}

} // namespace bidi
