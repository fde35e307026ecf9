// bidi_columns.cuh -- neo::Column::simStep (neo/Column.cpp:46-169) for every
// prediction node of one layer, all agents.
//
// Work split: one warp owns one column PAIR (nodes 2p, 2p+1).  Lane r < cells is cell r
// of BOTH columns: every per-column float of the reference is carried as a pair
// (x = column 2p, y = column 2p+1), which gives each lane two independent dependency
// chains to interleave.  During the integrate-and-fire sweep the lane walks its weight
// row sequentially, which is the reference's summation order; all 32 lanes share the
// reconstruction and the element-wise weight updates.
//
// Rounding: the reference's x86-64 build multiplies, rounds, then adds, and spikes are
// discontinuous in the last ulp, so nothing here may be a fused multiply-add.
//
// Data movement: the pair's record (bidi_types.h, ColumnsDev) is one contiguous block
// in HBM.  One cp.async.bulk (TMA, SASS UBLKCP) brings it into shared memory with an
// mbarrier transaction count, the iterations and all learning rules run out of shared
// memory, and one cp.async.bulk writes it back: every weight crosses HBM exactly once
// in each direction per step.  Rows are padded so that lane i's 128-bit loads of row
// i never bank-conflict (stride/2 odd).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "bidi_types.h"
#include "bidi_math.h"

namespace bidi {

#define BIDI_COL_WARPS 1

#ifdef BIDI_COL_TIMING
__device__ unsigned long long g_col_timing[16];
#define BIDI_T(k) do { __syncwarp(); if (lane == 0) { long long t_ = clock64(); atomicAdd(&g_col_timing[k], (unsigned long long)(t_ - t_prev_)); t_prev_ = t_; } } while (0)
#else
#define BIDI_T(k)
#endif

typedef unsigned long long u64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// global -> shared, completion signalled on the mbarrier (bytes and both addresses multiples of 16)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  // the shared-memory source must stay valid until the copy has READ it; the global
  // writes themselves complete before the next kernel of the stream starts
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (TMA store)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- packed pairs
__device__ __forceinline__ u64 pk(float x, float y) {
  u64 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y));
  return d;
}
__device__ __forceinline__ void upk(u64 v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
// Exact IEEE product / sum / difference of two packed pairs.
//
// B200's fp32 pipe issues a scalar FMUL/FADD/FFMA every 2 cycles per scheduler and a
// packed FFMA2 (PTX fma.rn.f32x2) at the same rate, i.e. twice the work per slot, but
// a dependent FFMA2 was measured at 17 cycles against 6 for FADD
// (profiles/r1_columns_notes.md).  So: everything that is NOT on a sequential
// dependency chain (all products, the element-wise learning updates) uses FFMA2, and
// the running sums -- whose order is the reference's and cannot be split -- use two
// interleaved scalar FADD chains per lane.
//
// ptxas fuses mul.f32x2 + add.f32x2 even when both carry .rn and --fmad=false, so the
// exact product is written fma(a, b, -0) and the exact packed sum fma(a, 1, c), with
// -0 / 1 / -1 taken from kernel parameters so that ptxas cannot fold and re-fuse them.
struct Pair {
  u64 one, negzero, negone;
  __device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) const {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
  }
  __device__ __forceinline__ u64 mul(u64 a, u64 b) const { return fma2(a, b, negzero); }
  // independent (not chained) packed sum / difference
  __device__ __forceinline__ u64 add2(u64 a, u64 b) const { return fma2(a, one, b); }
  __device__ __forceinline__ u64 sub2(u64 a, u64 b) const { return fma2(b, negone, a); }
  // chained sums: two scalar operations (built with -fmad=false: never fused)
  __device__ __forceinline__ u64 add(u64 a, u64 b) const {
    float ax, ay, bx, by; upk(a, ax, ay); upk(b, bx, by);
    return pk(ax + bx, ay + by);
  }
  __device__ __forceinline__ u64 sub(u64 a, u64 b) const {
    float ax, ay, bx, by; upk(a, ax, ay); upk(b, bx, by);
    return pk(ax - bx, ay - by);
  }
};
__device__ __forceinline__ u64 shfl64(u64 v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// floats of shared memory one column pair needs: its record, the gathered inputs, the final cell states and learning factors
// (+ the scratch of the 16-cell kernel: its parked rows, which first serve as the staging of the input planes)
__host__ __device__ inline int column_scratch_floats(const ColumnsDev& C) {
  const int park = C.rh ? 8 * (C.rh + 4) : 0, stage = (C.stage_n + 3) & ~3;
  return park > stage ? park : stage;
}
__host__ __device__ inline int column_slab_floats(const ColumnsDev& C) { return C.max_rec + 2 * C.max_s + 4 * C.cq + column_scratch_floats(C); }

template <int LPC>
__global__ void __launch_bounds__(BIDI_COL_WARPS * 32) k_columns(EngineDev P, LayerDev L, int l) {
  extern __shared__ __align__(128) float smem[];
  constexpr int PPW = 32 / LPC;  // column pairs per warp
  const ColumnsDev& C = L.col;
  const Pair F{P.k_one2, P.k_negzero2, P.k_negone2};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPC, r = lane % LPC, lane0 = sub * LPC;
  const int n_pairs = (C.n + 1) >> 1;
  const int tasks_per_agent = (n_pairs + PPW - 1) / PPW;
  const long long task = (long long)blockIdx.x * BIDI_COL_WARPS + warp;
  if (task >= (long long)P.A * tasks_per_agent) return;  // whole warps leave together
  const int a = (int)(task / tasks_per_agent);
  const int pair = (int)(task % tasks_per_agent) * PPW + sub;
  const bool active = pair < n_pairs;
  const int node_a = 2 * pair;
  const bool has_b = active && node_a + 1 < C.n;
  float* B = blob_of(P, a);

  const int Cc = C.cells, Cq = C.cq;
  const int slab = column_slab_floats(C);
  float* rec = smem + (size_t)(warp * PPW + sub) * slab;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)BIDI_COL_WARPS * PPW * slab) + warp;

  const int pr = active ? pair : 0;
  const int S = active ? C.stride[pr] : 0;
  const int R = active ? C.rec_off[pr + 1] - C.rec_off[pr] : 0;
  float* g_rec = B + C.rec_base + C.rec_off[pr];

#ifdef BIDI_COL_TIMING
  long long t_prev_ = clock64();
#endif
  // ---- start the record on its way in
  const int n_active = min(PPW, n_pairs - (int)(task % tasks_per_agent) * PPW);
  if (lane == 0) mbar_init(bar, n_active);
  __syncwarp();
  if (r == 0 && active) {
    mbar_expect_tx(bar, (uint32_t)R * 4u);
    tma_load_1d(rec, g_rec, (uint32_t)R * 4u, bar);
  }

  // record sections (all pair-interleaved)
  u64* W2 = reinterpret_cast<u64*>(rec);  // [Cc][S]
  u64* latT2 = W2 + Cc * S;               // [Cc][Cq]
  u64* v_thr = latT2 + Cc * Cq;
  u64* v_qw = v_thr + Cq; u64* v_qtr = v_qw + Cq;
  u64* v_aw = v_qtr + Cq;                 // [3][Cq]
  u64* v_atr = v_aw + 3 * Cq;
  u64* v_sprev = v_atr + 3 * Cq;
  u64* err2 = v_sprev + Cq;               // [S]
  u64* scal2 = err2 + S;                  // [4]
  u64* in2 = reinterpret_cast<u64*>(rec + C.max_rec);  // gathered inputs [S], padded with 0
  u64* st2s = in2 + C.max_s;                           // final cell states [Cq]
  u64* k2s = st2s + Cq;                                // feed-forward learning factors [Cq]

  // ---- gather both columns' inputs while the copy is in flight (neo/Agent.cpp:159-169,
  // :217-220).  C.in_src[j] is the blob offset input j is read from: the upper layer's
  // _signal actions and states for the feedback connections, then this layer's coder
  // states for the predictive ones -- resolved once on the host.
  // The exploration draws and the reward are fetched here too, long before their use.
  float nz_u = 0.0f, nz_v = 0.0f, nz_u_b = 0.0f, nz_v_b = 0.0f;
  const float reward = active ? P.rewards[a] : 0.0f;
  if (active && r < 3) {
    const long long nk = ((long long)a * P.ncol + C.noise_base + node_a) * 3 + r;
    nz_u = P.noise_u[nk]; nz_v = P.noise_v[nk];
    if (has_b) { nz_u_b = P.noise_u[nk + 3]; nz_v_b = P.noise_v[nk + 3]; }
  }
  // With one pair per warp and <= 32*GR inputs the gathered values wait in registers
  // until the first reconstruction needs them, so their load latency hides behind the
  // bulk copy and the first sweep.
  constexpr int GR = 4;
  const bool reg_gather = LPC == 32 && C.max_in <= 32 * GR;
  float gv[2][GR];
  int g_nin[2] = {0, 0};
  if (active) {
    float* in2f = reinterpret_cast<float*>(in2);
    for (int hf = 0; hf < 2; hf++) {
      const int node = node_a + hf;
      int nin = 0;
      if (node < C.n) {
        const int in0 = C.in_off[node];
        nin = C.in_off[node + 1] - in0;
        const int* src = C.in_src + in0;
        if (reg_gather) {
#pragma unroll
          for (int t = 0; t < GR; t++) { const int j = lane + 32 * t; gv[hf][t] = j < nin ? B[src[j]] : 0.0f; }
        } else {
          float* g_in = B + C.inputs + in0;
#pragma unroll 4
          for (int j = r; j < nin; j += LPC) {
            const float v = B[src[j]];
            in2f[2 * j + hf] = v;
            g_in[j] = v;
          }
        }
      } else if (reg_gather) {
#pragma unroll
        for (int t = 0; t < GR; t++) gv[hf][t] = 0.0f;
      }
      g_nin[hf] = nin;
      if (!reg_gather) for (int j = nin + r; j < S; j += LPC) in2f[2 * j + hf] = 0.0f;
    }
  }
  __syncwarp();
  BIDI_T(0);
  mbar_wait(bar, 0);
  BIDI_T(1);

#ifdef BIDI_COL_COPYONLY
  if (active) {  // experiment: record in, record out, nothing else
    const float4* s4 = reinterpret_cast<const float4*>(rec);
    float4* g4 = reinterpret_cast<float4*>(g_rec);
    for (int k = r; k < (R >> 2); k += LPC) g4[k] = s4[k];
  }
  return;
#endif
  // ---- iterate (neo/Column.cpp:58-103)
  const bool cell = active && r < Cc;
  float thr_a = 0.0f, thr_b = 0.0f, sp_a = 0.0f, sp_b = 0.0f;
  if (cell) { upk(v_thr[r], thr_a, thr_b); upk(v_sprev[r], sp_a, sp_b); }
  const unsigned gmask = LPC == 32 ? 0xffffffffu : 0xffffu;
  unsigned m_a = (__ballot_sync(0xffffffffu, sp_a != 0.0f) >> lane0) & gmask;
  unsigned m_b = (__ballot_sync(0xffffffffu, sp_b != 0.0f) >> lane0) & gmask;
  u64 act2 = 0ull;
  float st_a = 0.0f, st_b = 0.0f, counter = 0.0f, mult = 0.0f;
  const u64 oml2 = pk(1.0f - C.leak, 1.0f - C.leak);
  const ulonglong2* w4 = reinterpret_cast<const ulonglong2*>(W2 + (cell ? r : 0) * S);
  ulonglong2* e4 = reinterpret_cast<ulonglong2*>(err2);
  const int n2 = S >> 1;
  constexpr int UB = 4;
  for (int it = 0; it < C.iter; it++) {
    if (cell) {
      // excitation += w * err, sequentially over the inputs; loads run one block ahead of the math
      u64 exc2 = 0ull;
      int j = 0;
      ulonglong2 wc[UB], ec[UB];
      if (n2 >= UB) {
#pragma unroll
        for (int u = 0; u < UB; u++) { wc[u] = w4[u]; ec[u] = e4[u]; }
      }
      for (; j + 2 * UB <= n2; j += UB) {
        ulonglong2 wn[UB], en[UB];
#pragma unroll
        for (int u = 0; u < UB; u++) { wn[u] = w4[j + UB + u]; en[u] = e4[j + UB + u]; }
#pragma unroll
        for (int u = 0; u < UB; u++) {
          exc2 = F.add(F.mul(wc[u].x, ec[u].x), exc2);
          exc2 = F.add(F.mul(wc[u].y, ec[u].y), exc2);
        }
#pragma unroll
        for (int u = 0; u < UB; u++) { wc[u] = wn[u]; ec[u] = en[u]; }
      }
      if (j + UB <= n2) {
#pragma unroll
        for (int u = 0; u < UB; u++) {
          exc2 = F.add(F.mul(wc[u].x, ec[u].x), exc2);
          exc2 = F.add(F.mul(wc[u].y, ec[u].y), exc2);
        }
        j += UB;
      }
      for (; j < n2; j++) {
        const ulonglong2 w = w4[j], e = e4[j];
        exc2 = F.add(F.mul(w.x, e.x), exc2);
        exc2 = F.add(F.mul(w.y, e.y), exc2);
      }
      // inhibition += lat * spikePrev: only cells that spiked in either column can contribute
      u64 inh2 = 0ull;
      for (unsigned m = m_a | m_b; m; m &= m - 1) {
        const int k = __ffs(m) - 1;
        const u64 s2 = pk(((m_a >> k) & 1u) ? 1.0f : 0.0f, ((m_b >> k) & 1u) ? 1.0f : 0.0f);
        inh2 = F.add(F.mul(latT2[k * Cq + r], s2), inh2);
      }
      float av_a, av_b;
      upk(F.sub(F.add(F.mul(oml2, act2), exc2), inh2), av_a, av_b);
      if (av_a > thr_a) { av_a = 0.0f; sp_a = 1.0f; } else sp_a = 0.0f;
      if (av_b > thr_b) { av_b = 0.0f; sp_b = 1.0f; } else sp_b = 0.0f;
      st_a += sp_a; st_b += sp_b;
      act2 = pk(av_a, av_b);
    }
    BIDI_T(2);
    m_a = (__ballot_sync(0xffffffffu, cell && sp_a != 0.0f) >> lane0) & gmask;  // also orders the err reads above
    m_b = (__ballot_sync(0xffffffffu, cell && sp_b != 0.0f) >> lane0) & gmask;
    counter += 1.0f;
    mult = 1.0f / counter;
    if (it == 0 && reg_gather && active) {  // the gathered inputs land now (each lane reads back only its own entries)
      float* in2f = reinterpret_cast<float*>(in2);
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        float* g_in = B + C.inputs + (node_a + hf < C.n ? C.in_off[node_a + hf] : 0);
#pragma unroll
        for (int t = 0; t < GR; t++) {
          const int j = lane + 32 * t;
          if (j < S) in2f[2 * j + hf] = gv[hf][t];
          if (j < g_nin[hf]) g_in[j] = gv[hf][t];
        }
      }
      for (int j = 32 * GR + lane; j < S; j += 32) in2[j] = 0ull;
    }
    BIDI_T(3);
    // reconstruction from the spiking cells, ascending cell index (neo/Column.cpp:93-102):
    // recon += (spike*multiplier) * w ; error = input - recon
    constexpr int NP = 4;
    for (int j0 = 0; j0 < S; j0 += NP * LPC) {
      u64 rc[NP];
#pragma unroll
      for (int p = 0; p < NP; p++) rc[p] = 0ull;
      for (unsigned m = m_a | m_b; m; m &= m - 1) {
        const int k = __ffs(m) - 1;
        const u64 m2 = pk(((m_a >> k) & 1u) ? mult : 0.0f, ((m_b >> k) & 1u) ? mult : 0.0f);
        const u64* row = W2 + k * S + j0 + r;
#pragma unroll
        for (int p = 0; p < NP; p++)
          if (j0 + r + p * LPC < S) rc[p] = F.add(F.mul(m2, row[p * LPC]), rc[p]);
      }
#pragma unroll
      for (int p = 0; p < NP; p++) {
        const int j = j0 + r + p * LPC;
        if (j < S) err2[j] = F.sub2(in2[j], rc[p]);
      }
    }
    __syncwarp();
    BIDI_T(4);
  }
  st_a *= mult; st_b *= mult;
  const u64 st2 = pk(st_a, st_b);
  if (cell) st2s[r] = st2;
  __syncwarp();

  // ---- value and actions, sequential over the cells (neo/Column.cpp:111-123):
  // lanes 0..2 of the group accumulate one action each, lane 3 the value
  u64 accum2 = 0ull;
  if (active) {
    const u64* wv = r < 3 ? v_aw + r * Cq : v_qw;
    for (int k = 0; k < Cc; k++) accum2 = F.add(F.mul(wv[k], st2s[k]), accum2);
  }
  const u64 q2 = shfl64(accum2, lane0 + 3);
  u64 as2 = 0ull, ae2 = 0ull;
  if (active && r < 3) {
    float s_a, s_b;
    upk(accum2, s_a, s_b);
    const float as_a = bidi_sigmoid(s_a), as_b = bidi_sigmoid(s_b);
    float ae_a, ae_b = as_b;
    const long long nk = ((long long)a * P.ncol + C.noise_base + node_a) * 3 + r;
    ae_a = nz_u < C.expl_break ? nz_v : bidi_clamp01(as_a + nz_v);
    B[C.a_state + node_a * 3 + r] = as_a;
    B[C.a_expl + node_a * 3 + r] = ae_a;
    P.col_actions[nk] = ae_a;
    if (has_b) {
      ae_b = nz_u_b < C.expl_break ? nz_v_b : bidi_clamp01(as_b + nz_v_b);
      B[C.a_state + (node_a + 1) * 3 + r] = as_b;
      B[C.a_expl + (node_a + 1) * 3 + r] = ae_b;
      P.col_actions[nk + 3] = ae_b;
    }
    as2 = pk(as_a, as_b);
    ae2 = pk(ae_a, ae_b);
  }
  // tdError = reward + gamma*q - prevValue
  const u64 td2 = active ? F.sub(F.add(F.mul(pk(C.gamma, C.gamma), q2), pk(reward, reward)), scal2[0]) : 0ull;
  const u64 q_alpha_td2 = F.mul(pk(C.a_q, C.a_q), td2), act_alpha_td2 = F.mul(pk(C.a_act, C.a_act), td2);
  const u64 gl2 = pk(C.gamma_lambda, C.gamma_lambda);

  BIDI_T(5);
  // ---- learning (neo/Column.cpp:139-166), all inside the shared-memory record
  if (cell) {
    const u64 tr = v_qtr[r];
    v_qw[r] = F.add2(F.mul(q_alpha_td2, tr), v_qw[r]);
    v_qtr[r] = F.add2(F.mul(tr, gl2), st2);
  }
  for (int ai = 0; ai < 3; ai++) {
    const u64 as_i = shfl64(as2, lane0 + ai), ae_i = shfl64(ae2, lane0 + ai);
    if (cell) {
      const u64 tr = v_atr[ai * Cq + r];
      v_aw[ai * Cq + r] = F.add2(F.mul(act_alpha_td2, tr), v_aw[ai * Cq + r]);
      v_atr[ai * Cq + r] = F.add2(F.mul(tr, gl2), F.mul(F.sub2(ae_i, as_i), st2));
    }
  }
  const u64 sp2 = pk(C.sparsity * C.sparsity, C.sparsity * C.sparsity);
  const u64 aff2 = pk(C.a_ff, C.a_ff), alat2 = pk(C.a_lat, C.a_lat);
  // feed-forward: w[i][j] += (alpha*state_i)*err_j for the cells with state_i > 0.  A
  // cell that stays put gets the factor +0: w + (+-0) leaves w as it was.
  if (cell) k2s[r] = F.mul(aff2, pk(st_a > 0.0f ? st_a : 0.0f, st_b > 0.0f ? st_b : 0.0f));
  __syncwarp();
#ifdef BIDI_COL_FLAT_LEARN
  if (active) {
    // one flat, independent pass over the whole weight block (rows are contiguous)
    ulonglong2* wf = reinterpret_cast<ulonglong2*>(W2);
    const int tot = Cc * n2;
    int row = lane / n2, col = lane - row * n2;
#pragma unroll 4
    for (int e = lane; e < tot; e += 32) {
      ulonglong2 w = wf[e];
      const ulonglong2 er = e4[col];
      const u64 k2 = k2s[row];
      w.x = F.add2(F.mul(k2, er.x), w.x);
      w.y = F.add2(F.mul(k2, er.y), w.y);
      wf[e] = w;
      col += 32;
      while (col >= n2) { col -= n2; row++; }
    }
    // lateral: lat[r][i] = max(0, lat[r][i] + alpha*(state_r*state_i - sparsity^2)), every (i, r)
    for (int e = lane; e < Cc * Cq; e += 32) {
      const int i = e / Cq, rr = e - i * Cq;
      if (rr < Cc) {
        float la, lb;
        upk(F.add2(F.mul(alat2, F.sub2(F.mul(st2s[rr], st2s[i]), sp2)), latT2[e]), la, lb);
        latT2[e] = pk(bidi_max(0.0f, la), bidi_max(0.0f, lb));
      }
    }
  }
#else
  if (active) {
    for (int i = 0; i < Cc; i++) {
      const u64 si2 = st2s[i];
      float si_a, si_b;
      upk(si2, si_a, si_b);
      if (si_a > 0.0f || si_b > 0.0f) {
        const u64 k2 = k2s[i];
        ulonglong2* wr = reinterpret_cast<ulonglong2*>(W2 + i * S);
        for (int j = r; j < n2; j += LPC) {
          ulonglong2 w = wr[j];
          const ulonglong2 e = e4[j];
          w.x = F.add2(F.mul(k2, e.x), w.x);
          w.y = F.add2(F.mul(k2, e.y), w.y);
          wr[j] = w;
        }
      }
      if (cell) {  // lateral: lat[r][i] = max(0, lat[r][i] + alpha*(state_r*state_i - sparsity^2))
        u64* p = latT2 + i * Cq + r;
        float la, lb;
        upk(F.add2(F.mul(alat2, F.sub2(F.mul(st2, si2), sp2)), *p), la, lb);
        *p = pk(bidi_max(0.0f, la), bidi_max(0.0f, lb));
      }
    }
  }
#endif
  if (cell) {
    // threshold += alpha*(state - sparsity)
    v_thr[r] = F.add2(F.mul(pk(C.a_thr, C.a_thr), F.sub2(st2, pk(C.sparsity, C.sparsity))), pk(thr_a, thr_b));
    v_sprev[r] = pk(sp_a, sp_b);
    float av_a, av_b;
    upk(act2, av_a, av_b);
    const int cb = node_a * Cc + r;
    B[C.act + cb] = av_a; B[C.spike + cb] = sp_a; B[C.state + cb] = st_a;
    if (has_b) { B[C.act + cb + Cc] = av_b; B[C.spike + cb + Cc] = sp_b; B[C.state + cb + Cc] = st_b; }
  }
  __syncwarp();
  if (active && r == 0) scal2[0] = q2;

  BIDI_T(6);
  // ---- record back to HBM
#ifdef BIDI_COL_PLAIN_STORE
  __syncwarp();
  if (active) {
    const float4* s4 = reinterpret_cast<const float4*>(rec);
    float4* g4 = reinterpret_cast<float4*>(g_rec);
    const int n4 = R >> 2;
#pragma unroll 4
    for (int k = r; k < n4; k += LPC) g4[k] = s4[k];
  }
#else
  fence_async_smem();
  __syncwarp();
  if (r == 0 && active) {
    tma_store_1d(g_rec, rec, (uint32_t)R * 4u);
    tma_store_commit_and_wait();
  }
#endif
  BIDI_T(7);
}

} // namespace bidi
