// bidi_columns.cuh -- neo::Column::simStep (neo/Column.cpp:46-169) for every
// prediction node of one layer, all agents.
//
// Work split: a half-warp (16 lanes, LPC = 16) owns one column, so a warp runs two
// columns at once; columns with more than 16 cells take the whole warp (LPC = 32).
// Lane r of the group is cell r during the integrate-and-fire sweep -- it walks its
// weight row sequentially, which is the reference's summation order -- and the
// lanes share the reconstruction and the element-wise weight updates.
//
// Data movement: the column's record (bidi_types.h, ColumnsDev) is one contiguous
// block in HBM.  One cp.async.bulk (TMA, SASS UBLKCP) brings it into shared memory
// with an mbarrier transaction count, the 7 iterations and all learning rules run
// out of shared memory, and one cp.async.bulk writes it back: every weight crosses
// HBM exactly once in each direction per step.  Rows are padded so that lane i's
// 128-bit loads of row i never bank-conflict (stride/4 odd).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "bidi_types.h"
#include "bidi_math.h"

namespace bidi {

#define BIDI_COL_WARPS 4

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
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (TMA store)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// floats of shared memory one column needs: its record plus the gathered inputs
__host__ __device__ inline int column_slab_floats(int max_rec, int max_s) { return max_rec + max_s; }

template <int LPC>
__global__ void __launch_bounds__(BIDI_COL_WARPS * 32) k_columns(EngineDev P, LayerDev L, int l, long long up_a_expl, long long up_p_state) {
  extern __shared__ __align__(128) float smem[];
  constexpr int CPW = 32 / LPC;  // columns per warp
  const ColumnsDev& C = L.col;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPC, r = lane % LPC, lane0 = sub * LPC;
  const int tasks_per_agent = (C.n + CPW - 1) / CPW;
  const long long task = (long long)blockIdx.x * BIDI_COL_WARPS + warp;
  if (task >= (long long)P.A * tasks_per_agent) return;  // whole warps leave together
  const int a = (int)(task / tasks_per_agent);
  const int node = (int)(task % tasks_per_agent) * CPW + sub;
  const bool active = node < C.n;
  float* B = blob_of(P, a);

  const int Cc = C.cells, Cq = C.cq;
  const int slab = column_slab_floats(C.max_rec, C.max_s);
  float* rec = smem + (size_t)(warp * CPW + sub) * slab;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)BIDI_COL_WARPS * CPW * slab) + warp;

  const int nd = active ? node : 0;
  const int in0 = C.in_off[nd], nin = active ? C.in_off[nd + 1] - in0 : 0;
  const int S = active ? C.stride[nd] : 0, S4 = S >> 2;
  const int R = active ? C.rec_off[nd + 1] - C.rec_off[nd] : 0;
  float* g_rec = B + C.rec_base + C.rec_off[nd];

  // ---- start the record on its way in
  const int n_active = min(CPW, C.n - (int)(task % tasks_per_agent) * CPW);
  if (lane == 0) mbar_init(bar, n_active);
  __syncwarp();
  if (r == 0 && active) {
    mbar_expect_tx(bar, (uint32_t)R * 4u);
    tma_load_1d(rec, g_rec, (uint32_t)R * 4u, bar);
  }

  // record sections
  float* W = rec;
  float* latT = W + Cc * S;
  float* v_thr = latT + Cc * Cq;
  float* v_qw = v_thr + Cq; float* v_qtr = v_qw + Cq;
  float* v_aw = v_qtr + Cq;  // [3][Cq]
  float* v_atr = v_aw + 3 * Cq;
  float* v_sprev = v_atr + 3 * Cq;
  float* err = v_sprev + Cq;
  float* scal = err + S;
  float* s_in = rec + C.max_rec;  // gathered inputs, padded with 0 up to S

  // ---- gather the column's inputs while the copy is in flight (neo/Agent.cpp:159-169, :217-220)
  const bool input_layer = l == P.L;
  if (active) {
    const int ux = node % L.fb.uw, uy = node / L.fb.uw;
    int nfb = 0;
    if (L.fb.r >= 0 && (input_layer || l < P.L - 1)) {
      const SlotBox b = slot_box(L.fb, ux, uy);
      const int ny = b.sy1 - b.sy0 + 1;
      nfb = slot_count(L.fb, b);
      const float* up_sig = B + up_a_expl;
      const float* up_state = B + up_p_state;
      for (int k = r; k < nfb; k += LPC) {
        const int sx = b.sx0 + k / ny, sy = b.sy0 + k % ny;
        const int t = slot_tx(L.fb, b, sx) + slot_ty(L.fb, b, sy) * L.fb.tw;
        s_in[2 * k] = up_sig[t * 3 + 2];  // _signal
        s_in[2 * k + 1] = up_state[t];
      }
    }
    if (!input_layer) {
      const SlotBox b = slot_box(L.pred, ux, uy);
      const int ny = b.sy1 - b.sy0 + 1;
      const int npr = slot_count(L.pred, b);
      const float* state = B + L.state;
      for (int k = r; k < npr; k += LPC) {
        const int sx = b.sx0 + k / ny, sy = b.sy0 + k % ny;
        s_in[2 * nfb + k] = state[slot_tx(L.pred, b, sx) + slot_ty(L.pred, b, sy) * L.pred.tw];
      }
    }
    for (int j = nin + r; j < S; j += LPC) s_in[j] = 0.0f;
  }
  __syncwarp();
  if (active) {
    float* g_in = B + C.inputs + in0;
    for (int j = r; j < nin; j += LPC) g_in[j] = s_in[j];
  }
  mbar_wait(bar, 0);

  // ---- iterate (neo/Column.cpp:58-103)
  const bool cell = active && r < Cc;
  const float thr = cell ? v_thr[r] : 0.0f;
  const unsigned gmask = LPC == 32 ? 0xffffffffu : 0xffffu;
  unsigned spikes = (__ballot_sync(0xffffffffu, cell && v_sprev[r] != 0.0f) >> lane0) & gmask;
  float act = 0.0f, st = 0.0f, spike = 0.0f, counter = 0.0f, mult = 0.0f;
  const float4* w4 = reinterpret_cast<const float4*>(W + (cell ? r : 0) * S);
  float4* e4 = reinterpret_cast<float4*>(err);
  const float4* in4 = reinterpret_cast<const float4*>(s_in);
  for (int it = 0; it < C.iter; it++) {
    if (cell) {
      float excitation = 0.0f;
#pragma unroll 4
      for (int j = 0; j < S4; j++) {
        const float4 w = w4[j], e = e4[j];
        excitation += w.x * e.x; excitation += w.y * e.y; excitation += w.z * e.z; excitation += w.w * e.w;
      }
      // lateral inhibition: spikePrev is 0 or 1, so only the spiking cells contribute, exactly their weight
      float inhibition = 0.0f;
      for (unsigned m = spikes; m; m &= m - 1) inhibition += latT[(__ffs(m) - 1) * Cq + r];
      float activation = (1.0f - C.leak) * act + excitation - inhibition;
      if (activation > thr) { activation = 0.0f; spike = 1.0f; } else spike = 0.0f;
      st += spike;
      act = activation;
    }
    spikes = (__ballot_sync(0xffffffffu, cell && spike != 0.0f) >> lane0) & gmask;  // also orders the err reads above
    counter += 1.0f;
    mult = 1.0f / counter;
    // reconstruction from the spiking cells, ascending cell index (neo/Column.cpp:93-102)
    for (int j = r; j < S4; j += LPC) {
      float4 rc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      for (unsigned m = spikes; m; m &= m - 1) {
        const float4 w = reinterpret_cast<const float4*>(W + (__ffs(m) - 1) * S)[j];
        rc.x += mult * w.x; rc.y += mult * w.y; rc.z += mult * w.z; rc.w += mult * w.w;
      }
      const float4 x = in4[j];
      e4[j] = make_float4(x.x - rc.x, x.y - rc.y, x.z - rc.z, x.w - rc.w);
    }
    __syncwarp();
  }
  st *= mult;

  // ---- value and actions, sequential over the cells (neo/Column.cpp:111-123):
  // lanes 0..2 of the group accumulate one action each, lane 3 the value
  float accum = 0.0f;
  {
    const float* wv = r < 3 ? v_aw + r * Cq : v_qw;
    for (int k = 0; k < Cc; k++) {
      const float sk = __shfl_sync(0xffffffffu, st, lane0 + k);
      if (active) accum += wv[k] * sk;
    }
  }
  const float q = __shfl_sync(0xffffffffu, accum, lane0 + 3);
  float a_state = 0.0f, a_expl = 0.0f;
  if (active && r < 3) {
    a_state = bidi_sigmoid(accum);
    const long long nk = ((long long)a * P.ncol + C.noise_base + node) * 3 + r;
    const float u = P.noise_u[nk], v = P.noise_v[nk];
    a_expl = u < C.expl_break ? v : bidi_clamp01(a_state + v);
    B[C.a_state + node * 3 + r] = a_state;
    B[C.a_expl + node * 3 + r] = a_expl;
    P.col_actions[nk] = a_expl;
  }
  const float td = active ? P.rewards[a] + C.gamma * q - scal[0] : 0.0f;
  const float q_alpha_td = C.a_q * td, act_alpha_td = C.a_act * td;

  // ---- learning (neo/Column.cpp:139-166), all inside the shared-memory record
  if (cell) {
    const float tr = v_qtr[r];
    v_qw[r] += q_alpha_td * tr;
    v_qtr[r] = tr * C.gamma_lambda + st;
  }
  for (int ai = 0; ai < 3; ai++) {
    const float as = __shfl_sync(0xffffffffu, a_state, lane0 + ai), ae = __shfl_sync(0xffffffffu, a_expl, lane0 + ai);
    if (cell) {
      const float tr = v_atr[ai * Cq + r];
      v_aw[ai * Cq + r] += act_alpha_td * tr;
      v_atr[ai * Cq + r] = tr * C.gamma_lambda + (ae - as) * st;
    }
  }
  const float sp2 = C.sparsity * C.sparsity;
  for (int i = 0; i < Cc; i++) {
    const float si = __shfl_sync(0xffffffffu, st, lane0 + i);
    if (active && si > 0.0f) {  // feed-forward: w += (alpha*state_i)*err_j
      const float k = C.a_ff * si;
      float4* wr = reinterpret_cast<float4*>(W + i * S);
      for (int j = r; j < S4; j += LPC) {
        float4 w = wr[j];
        const float4 e = e4[j];
        w.x += k * e.x; w.y += k * e.y; w.z += k * e.z; w.w += k * e.w;
        wr[j] = w;
      }
    }
    if (cell) {  // lateral: lat[r][i] = max(0, lat[r][i] + alpha*(state_r*state_i - sparsity^2))
      float* p = latT + i * Cq + r;
      *p = bidi_max(0.0f, *p + C.a_lat * (st * si - sp2));
    }
  }
  if (cell) {
    v_thr[r] = thr + C.a_thr * (st - C.sparsity);
    v_sprev[r] = spike;
    const int cb = node * Cc + r;
    B[C.act + cb] = act;
    B[C.spike + cb] = spike;
    B[C.state + cb] = st;
  }
  __syncwarp();
  if (active && r == 0) scal[0] = q;

  // ---- record back to HBM
  fence_async_smem();
  __syncwarp();
  if (r == 0 && active) {
    tma_store_1d(g_rec, rec, (uint32_t)R * 4u);
    tma_store_commit_and_wait();
  }
}

} // namespace bidi
