// bidi_qnet.cuh -- the part of sdr::QPRSDR::simStep that follows the hierarchy step
// (sdr/QPRSDR.cpp:102-341): a sparse Q network over the hierarchy's grids (sigmoid units
// gated by the prediction states), 32 sign-gradient iterations that derive the action,
// exploration, and TD(lambda) updates through eligibility traces.
//
// One CTA per agent runs the whole sequence -- 2L+2 dependent phases per derive iteration
// -- with block barriers between phases; all state lives in the agent's blob.
//   * The Q layers share the window geometry of the coder's feed-forward family; the
//     reference keeps only connections on a path from an action input
//     (sdr/QPRSDR.cpp:72-81).  Here every window slot has a plane entry; slots the
//     reference drops hold weight 0 and trace 0 for ever (the learning rule is masked), so
//     they add (+-0) to the forward sums and to the back-propagated errors.
//   * Back-propagation is the gather form of the reference's scatter
//     (prev._error += q._error * w, nodes ascending): each lower node sums over the upper
//     nodes whose window contains it, ascending -- the same order of additions.
// The derive step moves each action by +-alpha according to the SIGN of its error, so as
// everywhere else in this engine the arithmetic keeps the reference's operation order.
#pragma once
#include "bidi_kernels.cuh"

namespace bidi {

__device__ __forceinline__ float clamp11(float x) { return bidi_min(1.0f, bidi_max(-1.0f, x)); }

// forward pass from the action values `actions` (derive or exploratory), sdr/QPRSDR.cpp:111-144
__device__ __forceinline__ void q_forward(const EngineDev& P, float* B, int a, const float* actions) {
  for (int l = 0; l < P.L; l++) {
    const LayerDev& L = P.layers[l];
    const float* below = l ? B + P.layers[l - 1].q_state : nullptr;
    for (int i = threadIdx.x; i < L.nh; i += blockDim.x) {
      const int ux = i % L.hw, uy = i / L.hw;
      float sum = 0.0f;
      if (l == 0) sum = dot_conn(L.ff, ux, uy, B + L.q_w_off, L.nh, i, sum, [&](int t) { const int k = P.q_act_index[t]; return k >= 0 ? actions[k] : 0.0f; });
      else sum = dot_conn(L.ff, ux, uy, B + L.q_w_off, L.nh, i, sum, [&](int t) { return below[t]; });
      B[L.q_state + i] = bidi_sigmoid(sum) * B[L.p_state + i];
      B[L.q_err + i] = 0.0f;
    }
    __syncthreads();
  }
}
// back-propagation of dQ/dstate (sdr/QPRSDR.cpp:154-190, :268-301)
__device__ __forceinline__ void q_backward(const EngineDev& P, float* B, bool to_actions, float* a_err) {
  {
    const LayerDev& T = P.layers[P.L - 1];
    for (int i = threadIdx.x; i < P.q_top; i += blockDim.x) B[T.q_err + i] = B[P.q_qc_w + i];
    __syncthreads();
  }
  for (int l = P.L - 1; l >= 0; l--) {
    const LayerDev& L = P.layers[l];
    for (int i = threadIdx.x; i < L.nh; i += blockDim.x) {
      const float s = B[L.q_state + i];
      B[L.q_err + i] = B[L.q_err + i] * s * (1.0f - s);
    }
    __syncthreads();
    if (l > 0) {
      const LayerDev& D = P.layers[l - 1];
      for (int t = threadIdx.x; t < D.nh; t += blockDim.x)
        B[D.q_err + t] = gather_recon(L.ff, B + L.q_w_off, B + L.q_err, t % L.vw, t / L.vw, false, 0.0f);
    } else if (to_actions) {
      for (int k = threadIdx.x; k < P.q_half; k += blockDim.x) {
        const int t = P.q_a_input[k];
        a_err[k] = gather_recon(L.ff, B + L.q_w_off, B + L.q_err, t % L.vw, t / L.vw, false, 0.0f);
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_qnet(EngineDev P, int learn) {
  const int a = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  float* B = blob_of(P, a);
  const float* pred = P.pred_out + (long long)a * P.n_in;  // the input prediction nodes' states
  const int half = P.q_half, nn = 2 * half;
  float* a_pred = B + P.q_act; float* a_derive = a_pred + nn; float* a_expl = a_derive + nn; float* a_err = a_expl + nn;
  __shared__ float s_q;
  for (int i = tid; i < nn; i += T) { const float v = clamp11(pred[P.q_a_input[i]]); a_pred[i] = v; a_derive[i] = v; }
  __syncthreads();
  for (int it = 0; it < P.q_iters; it++) {
    q_forward(P, B, a, a_derive);
    for (int i = tid; i < nn; i += T) a_err[i] = 0.0f;
    __syncthreads();
    q_backward(P, B, true, a_err);
    for (int i = tid; i < half; i += T) {
      const float d = clamp11(a_derive[i] + (a_err[i] > 0.0f ? 1.0f : -1.0f) * P.q_derive_alpha);
      a_derive[i] = d;
      a_derive[half + i] = 1.0f - d;
    }
    __syncthreads();
  }
  // explore (sdr/QPRSDR.cpp:202-217); the draws come with the step (bidi_step noise_u / noise_v)
  for (int i = tid; i < half; i += T) {
    const float u = P.noise_u[(long long)a * half + i], v = P.noise_v[(long long)a * half + i];
    const float e = u < P.q_expl_break ? v : clamp11(a_derive[i] + v);
    a_expl[i] = e;
    a_expl[half + i] = 1.0f - e;
  }
  __syncthreads();
  q_forward(P, B, a, a_expl);
  const LayerDev& Top = P.layers[P.L - 1];
  if (tid == 0) {
    float q = 0.0f;
    for (int i = 0; i < P.q_top; i++) q += B[P.q_qc_w + i] * B[Top.q_state + i];
    s_q = q;
  }
  __syncthreads();
  const float q = s_q;
  const float td = P.rewards[a] + P.q_gamma * q - B[P.q_prev];
  const float q_alpha_td = P.q_alpha * td, action_alpha_td = P.q_action_alpha * td;
  __syncthreads();  // every thread has read the previous value
  if (tid == 0) B[P.q_prev] = q;
  if (learn) {  // sdr/QPRSDR.cpp:266-331
    q_backward(P, B, false, a_err);
    for (int l = 0; l < P.L; l++) {
      const LayerDev& L = P.layers[l];
      const float* below = l ? B + P.layers[l - 1].q_state : nullptr;
      float* w = B + L.q_w_off; float* tr = B + L.q_tr_off;
      for (int i = tid; i < L.nh; i += T) {
        const float err = B[L.q_err + i];
        const float bt = B[L.q_bias_tr + i];
        B[L.q_bias_w + i] += action_alpha_td * bt;
        B[L.q_bias_tr + i] = P.q_gamma_lambda * bt + err;
        for_each_conn(L.ff, i % L.hw, i / L.hw, [&](int s, int t) {
          const long long k = (long long)s * L.nh + i;
          if (L.q_mask[k] != 0.0f) {
            const float src = l ? below[t] : a_expl[P.q_act_index[t]];
            const float trk = tr[k];
            w[k] += action_alpha_td * trk;
            tr[k] = P.q_gamma_lambda * trk + err * src;
          }
        });
      }
    }
  }
  // Q connections, also when not learning (sdr/QPRSDR.cpp:334-340)
  for (int i = tid; i < P.q_top; i += T) {
    const float trk = B[P.q_qc_tr + i];
    B[P.q_qc_w + i] += q_alpha_td * trk;
    B[P.q_qc_tr + i] = P.q_gamma_lambda * trk + B[Top.q_state + i];
  }
}

// exploration draws when the caller supplies none (counter-based, like k_noise)
__global__ void k_noise_q(EngineDev P, float* u, float* v, unsigned long long seed, unsigned long long step, long long first) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)P.A * P.q_half) return;
  unsigned long long h = mix64(seed ^ mix64(step * 0x100000001B3ULL + (unsigned long long)(gid + first * P.q_half)));
  const float uu = u01(h);
  h = mix64(h);
  float vv;
  if (uu < P.q_expl_break) vv = u01(h);
  else {
    const float r1 = fmaxf(u01(h), 5.9604645e-8f), r2 = u01(mix64(h));
    vv = sqrtf(-2.0f * logf(r1)) * cospif(2.0f * r2) * P.q_expl_std;
  }
  u[gid] = uu;
  v[gid] = vv;
}

} // namespace bidi
