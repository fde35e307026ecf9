// bidi_csrl.cuh -- deep::CSRL (deep/CSRL.cpp:189-449): the parts of its step the other hierarchies do not have.
// The coders are sdr::IRSDR (the iterative kernels), the top-down pass is k_predict / k_predict_input without
// their delta rule; here are
//   k_csrl_explore      exploration of the action-type input predictions                 deep/CSRL.cpp:246-253
//   k_sdrrl             one deep::SDRRL unit per prediction node (deep/SDRRL.cpp:46-277): inputs gathered from the
//                       rewards above / the neighbourhood's coder states / its own recurrent actions, a small
//                       spiking coder, 30 sign-gradient iterations deriving the action, exploration, TD(lambda)
//   k_csrl_pred_learn   feedback / predictive weights through eligibility traces, gated by the unit's _learn action
//                                                                                         deep/CSRL.cpp:336-419
//   k_csrl_noise        the exploration draws when the caller supplies none
// The reward-modulated coder rule (sdr/IRSDR.cpp:321-356) is k_neo_learn with the units' _reward actions.
// One warp per (agent, node) in k_sdrrl: lanes are cells in the cell phases, inputs in the reconstruction, actions
// in the derive loop; every sum runs in the reference's order.  This variant is not on the benchmarked path
// (SURVEY 8f.4): written for exact parity, not tuned.
#pragma once
#include "bidi_kernels.cuh"

namespace bidi {

#define BIDI_SDRRL_WARPS 4
#define BIDI_SDRRL_MAX_ACTIONS 64  /* 2*(3 + recurrent inputs) */

// shared memory per warp, in floats
__host__ __device__ inline int sdrrl_warp_floats(int max_in) { return 2 * max_in + 4 * 32 + 3 * BIDI_SDRRL_MAX_ACTIONS; }

__global__ void k_csrl_explore(EngineDev P) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)P.A * P.csrl_n_actions) return;
  const int a = (int)(gid / P.csrl_n_actions), k = (int)(gid % P.csrl_n_actions);
  const int pi = P.csrl_action_idx[k];
  const long long nk = (long long)a * P.n_noise + k;
  float* out = P.pred_out + (long long)a * P.n_in + pi;
  const float u = P.noise_u[nk], v = P.noise_v[nk];
  if (u < P.csrl_expl_break) *out = v * 2.0f - 1.0f;
  else *out = bidi_min(1.0f, bidi_max(-1.0f, bidi_min(1.0f, bidi_max(-1.0f, *out)) + v));
}

__global__ void __launch_bounds__(32 * BIDI_SDRRL_WARPS) k_sdrrl(EngineDev P, LayerDev L, int l) {
  extern __shared__ float sm[];
  const SdrrlDev& S = L.sd;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long task = (long long)blockIdx.x * BIDI_SDRRL_WARPS + warp;
  if (task >= (long long)P.A * S.n) return;  // whole warps leave together
  const int a = (int)(task / S.n), node = (int)(task % S.n);
  float* B = blob_of(P, a);
  const int cells = S.cells, na = S.na, half = na >> 1, in0 = S.in_off[node], nin = S.in_off[node + 1] - in0;
  float* in_s = sm + (size_t)warp * sdrrl_warp_floats(S.max_in);
  float* re_s = in_s + S.max_in;      // _reconstructionError
  float* spk_s = re_s + S.max_in;     // [32] this iteration's spikes, then the cells' states
  float* st_s = spk_s + 32;           // [32] cell states
  float* as_s = st_s + 32;            // [32] Cell::_actionState
  float* ae_s = as_s + 32;            // [32] Cell::_actionError
  float* act_s = ae_s + 32;           // [na] Action::_state
  float* expl_s = act_s + BIDI_SDRRL_MAX_ACTIONS;
  float* aerr_s = expl_s + BIDI_SDRRL_MAX_ACTIONS;
  const float reward = P.rewards[a];
  const unsigned full = 0xffffffffu;

  // ---- inputs (deep/CSRL.cpp:267-276, 295-304, 320-324): reward from above (or the offsets + reward at the top),
  //      the neighbourhood's coder states, the unit's own recurrent actions of the previous step
  for (int j = lane; j < nin; j += 32) {
    const int s = S.in_src[in0 + j];
    const float v = s >= 0 ? B[s] : B[P.csrl_reward_offsets + (-1 - s)] + reward;
    in_s[j] = v;
    B[S.inputs + in0 + j] = v;
    re_s[j] = B[S.recon_err + in0 + j];
  }
  for (int i = lane; i < na; i += 32) { act_s[i] = B[S.a_state + (long long)node * na + i]; aerr_s[i] = B[S.a_err + (long long)node * na + i]; }
  __syncwarp();

  const bool cell = lane < cells;
  const int ck = node * cells + lane;
  float* ff = B + S.ff_w + (long long)in0 * cells;          // [cell][nin]
  float* lat = B + S.lat_w + (long long)node * cells * cells;  // [cell][cell]
  float thr = cell ? B[S.thr + ck] : 0.0f;
  float sprev = cell ? B[S.spike_prev + ck] : 0.0f;
  float act = 0.0f, st = 0.0f, spike = 0.0f;
  const float oml = 1.0f - S.leak, measure_inv = 1.0f / (float)S.measure;

  auto reconstruct = [&](const float* drive) {  // _reconstructionError = input - sum_j drive_j * w_ji, ascending j
    for (int i = lane; i < nin; i += 32) {
      float recon = 0.0f;
      for (int j = 0; j < cells; j++) recon += drive[j] * ff[(long long)j * nin + i];
      re_s[i] = in_s[i] - recon;
    }
    __syncwarp();
  };

  // ---- settle, then measure (deep/SDRRL.cpp:63-140)
  for (int iter = 0; iter < S.settle + S.measure; iter++) {
    float exc = 0.0f, inh = 0.0f;
    if (cell) for (int j = 0; j < nin; j++) exc += ff[(long long)lane * nin + j] * re_s[j];
    for (int j = 0; j < cells; j++) {
      const float sj = __shfl_sync(full, sprev, j);
      if (cell) inh += lat[lane * cells + j] * sj;
    }
    float av = oml * act + exc - inh;
    if (av > thr) { av = 0.0f; spike = 1.0f; } else spike = 0.0f;
    if (iter >= S.settle) st += spike * measure_inv;
    act = av;
    sprev = spike;
    __syncwarp();  // everybody has read the previous errors
    if (cell) spk_s[lane] = spike;
    __syncwarp();
    reconstruct(spk_s);
  }
  if (cell) st_s[lane] = st;
  __syncwarp();
  reconstruct(st_s);  // final state reconstruction, :143-151

  // ---- action derivation (:153-200)
  const float qw = cell ? B[S.q_w + ck] : 0.0f;
  float* aw = B + S.a_w + (long long)node * cells * na;   // _actionConnections [cell][action]
  float* atr = B + S.a_tr + (long long)node * cells * na;
  float cas = cell ? B[S.cell_as + ck] : 0.0f, cae = cell ? B[S.cell_ae + ck] : 0.0f;
  auto forward = [&](const float* actions) {  // Cell::_actionState of this lane's cell
    if (!cell) return;
    if (st > 0.0f) {
      float sum = 0.0f;
      for (int vi = 0; vi < na; vi++) sum += aw[lane * na + vi] * actions[vi];
      cas = bidi_sigmoid(sum) * st;
    } else cas = 0.0f;
  };
  for (int i = lane; i < half; i += 32) act_s[i + half] = 1.0f - act_s[i];
  __syncwarp();
  for (int iter = 0; iter < S.derive_iter; iter++) {
    forward(act_s);
    cae = qw * cas * (1.0f - cas);
    if (cell) ae_s[lane] = cae;
    __syncwarp();
    for (int i = lane; i < na; i += 32) {
      float sum = 0.0f;
      for (int k = 0; k < cells; k++) sum += aw[k * na + i] * ae_s[k];
      aerr_s[i] = sum;
    }
    __syncwarp();
    for (int i = lane; i < half; i += 32)
      act_s[i] = bidi_min(1.0f, bidi_max(0.0f, act_s[i] + S.derive_alpha * ((aerr_s[i] - aerr_s[i + half]) > 0.0f ? 1.0f : -1.0f)));
    __syncwarp();
    for (int i = lane; i < half; i += 32) act_s[i + half] = 1.0f - act_s[i];
    __syncwarp();
  }
  // ---- exploration (:203-211); the draws come from the caller's generator or from k_csrl_noise
  const long long nk = (long long)a * P.n_noise + S.noise_base + (long long)node * na;
  for (int i = lane; i < na; i += 32) {
    const float u = P.noise_u[nk + i], v = P.noise_v[nk + i];
    expl_s[i] = u < S.expl_break ? v : bidi_min(1.0f, bidi_max(0.0f, act_s[i] + v));
  }
  __syncwarp();
  for (int i = lane; i < half; i += 32) expl_s[i + half] = 1.0f - expl_s[i];
  __syncwarp();
  forward(expl_s);
  if (cell) { as_s[lane] = cas; spk_s[lane] = qw; }
  __syncwarp();
  float q = 0.0f;
  for (int k = 0; k < cells; k++) if (st_s[k] > 0.0f) q += spk_s[k] * as_s[k];  // every lane: the same sequential sum
  // ---- TD(lambda) (:231-259)
  const float td = reward + S.gamma * q - B[S.prev_value + node];
  const float q_alpha_td = S.a_q * td, act_alpha_td = S.a_act * td, surprise = td * td;
  const float avg = B[S.avg_surprise + node];
  __syncwarp();
  if (lane == 0) {
    B[S.avg_surprise + node] = (1.0f - S.surprise_decay) * avg + S.surprise_decay * surprise;
    B[S.prev_value + node] = q;
    B[L.p_mod + node] = expl_s[2];  // _localReward = getAction(_reward)
  }
  if (cell) {
    const float error = qw * cas * (1.0f - cas);
    for (int vi = 0; vi < na; vi++) {
      const float tr = atr[lane * na + vi];
      aw[lane * na + vi] = aw[lane * na + vi] + act_alpha_td * tr;
      atr[lane * na + vi] = tr * S.gamma_lambda + error * expl_s[vi];
    }
    const float qtr = B[S.q_tr + ck];
    B[S.q_w + ck] = qw + q_alpha_td * qtr;
    B[S.q_tr + ck] = qtr * S.gamma_lambda + cas;
    // ---- the unit's own coder learns (:262-275)
    if (st > 0.0f)
      for (int j = 0; j < nin; j++) ff[(long long)lane * nin + j] += S.a_ff * st * re_s[j];
    const float sp2 = S.sparsity * S.sparsity;
    for (int j = 0; j < cells; j++) lat[lane * cells + j] = bidi_max(0.0f, lat[lane * cells + j] + S.a_lat * (st * st_s[j] - sp2));
    B[S.thr + ck] = thr + S.a_thr * (st - S.sparsity);
    B[S.act + ck] = act; B[S.spike + ck] = spike; B[S.spike_prev + ck] = sprev; B[S.state + ck] = st;
    B[S.cell_as + ck] = cas; B[S.cell_ae + ck] = cae;
  }
  for (int j = lane; j < nin; j += 32) B[S.recon_err + in0 + j] = re_s[j];
  for (int i = lane; i < na; i += 32) {
    B[S.a_state + (long long)node * na + i] = act_s[i];
    B[S.a_expl + (long long)node * na + i] = expl_s[i];
    B[S.a_err + (long long)node * na + i] = aerr_s[i];
  }
}

struct WT { float w, tr, src; };

// deep/CSRL.cpp:336-419.  l < L: the layer's nodes (feedback to the layer above unless top, predictive); l == L: the
// input nodes (feedback into layer 0; gate = 1 for cells that are not of type _action)
__global__ void k_csrl_pred_learn(EngineDev P, LayerDev L, int l, long long up_state_prev) {
  BIDI_THREAD_AGENT_UNIT(L.pn)
  float* B = blob_of(P, a);
  const SdrrlDev& S = L.sd;
  const bool input = l == P.L;
  const float target = input ? P.in0[(long long)a * P.n_in + i] : B[L.state + i];
  const float err = target - B[L.p_act_prev + i];
  float gate = B[S.a_expl + (long long)i * S.na + 1];
  if (input && !P.csrl_is_action[i]) gate = 1.0f;
  const int ux = i % L.fb.uw, uy = i / L.fb.uw, U = L.pn;
  const float gl = S.gamma_lambda;
  if (input || l < P.L - 1) {
    float* w = B + L.fb.w_off; float* tr = B + L.fb.tr_off;
    const float* src = B + up_state_prev;
    const float k = L.learn_fb * gate;
    for_each_slot_batched<8, WT>(L.fb, ux, uy,
        [&](int s, int t) { const long long g = (long long)s * U + i; return WT{w[g], tr[g], src[t]}; },
        [&](int s, const WT& v) {
          const long long g = (long long)s * U + i;
          const float nt = gl * v.tr + err * v.src;
          tr[g] = nt;
          w[g] = v.w + k * nt;
        });
  }
  if (!input) {
    float* w = B + L.pred.w_off; float* tr = B + L.pred.tr_off;
    const float* src = B + L.state_prev;
    const float k = L.learn_pred * gate;
    for_each_slot_batched<8, WT>(L.pred, ux, uy,
        [&](int s, int t) { const long long g = (long long)s * U + i; return WT{w[g], tr[g], src[t]}; },
        [&](int s, const WT& v) {
          const long long g = (long long)s * U + i;
          const float nt = gl * v.tr + err * v.src;
          tr[g] = nt;
          w[g] = v.w + k * nt;
        });
  }
}

// exploration draws when the caller supplies none (counter-based, like k_noise): entry e of an agent belongs to the
// action inputs (e < n_actions) or to a unit of some node layer (SdrrlDev::noise_base)
__global__ void k_csrl_noise(EngineDev P, float* u, float* v, unsigned long long seed, unsigned long long step, long long first) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)P.A * P.n_noise) return;
  const int a = (int)(gid / P.n_noise), e = (int)(gid % P.n_noise);
  float std = P.csrl_expl_std, brk = P.csrl_expl_break;
  for (int l = 0; l <= P.L; l++) {
    const SdrrlDev& S = P.layers[l].sd;
    if (e >= S.noise_base && e < S.noise_base + S.n * S.na) { std = S.expl_std; brk = S.expl_break; }
  }
  unsigned long long h = mix64(seed ^ mix64(step * 0x100000001B3ULL + (unsigned long long)(a + first) * P.n_noise + e));
  const float uu = u01(h);
  h = mix64(h);
  float vv;
  if (uu < brk) vv = u01(h);
  else {
    const float r1 = fmaxf(u01(h), 5.9604645e-8f), r2 = u01(mix64(h));
    vv = sqrtf(-2.0f * logf(r1)) * cospif(2.0f * r2) * std;
  }
  u[gid] = uu;
  v[gid] = vv;
}

} // namespace bidi
