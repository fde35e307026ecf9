// bidi_env.cuh -- a batched, headless stand-in for the reference's Box2D quadruped (runner/Runner.cpp), one
// instance per agent, stepped on the device so that the RunnerMain loop (RunnerMain.cpp:204-251) runs CLOSED:
//   reward            = body velocity x                                      (:207-210)
//   inputs 0..14      = Runner::getStateVector (runner/Runner.cpp:260-339): 10 joint angles (left back 3,
//                       left front 3, right back 2, right front 2), the body angle, 4 foot contacts
//   inputs 15..18     = the four clock sines                                  (:235-236; computed by the host)
//   inputs 20..29     = clamp01(prediction*0.5 + 0.5 + N(0, 0.05))            (:241-242)
//   simStep(reward)                                                           (:246)
//   action            = clamp01(prediction*0.5 + 0.5) -> Runner::motorUpdate(action, 12) (runner/Runner.cpp:341),
//                       the keep-upright rule (:248-249), one physics tick of 1/60 s
// Box2D is not part of this build (nor installable here), so the rigid-body world is replaced by a small
// deterministic planar surrogate with the same interface: position-controlled joints with limits, two-segment
// legs whose feet touch the ground when extended, a body pushed forward by the feet that sweep backwards while in
// contact.  Plain fp32 +,-,*,/ in a fixed order and polynomial sin/cos, so that the CPU twin the tests hold it
// against (an independent restatement among the test checkers) agrees bit for bit.
#pragma once
#include "bidi_kernels.cuh"

namespace bidi {

#define BIDI_ENV_JOINTS 10
#define BIDI_ENV_STATE 15      /* floats of Runner::getStateVector */
#define BIDI_ENV_FLOATS 24     /* per agent: 10 angles, body angle, 4 contacts, vx, angular velocity, x, 4 foot x, pad */

struct EnvParams {
  float dt, interp, max_speed, thigh, shin, ankle_gain, contact_height, traction, drag, max_body_angle, upright_gain;
  float jmin[BIDI_ENV_JOINTS], jmax[BIDI_ENV_JOINTS];
  int hip[4], knee[4], ankle[4];  // joint index of every leg's segments (ankle -1: the right legs have two segments)
};

__host__ __device__ inline EnvParams env_default_params() {
  EnvParams p;
  p.dt = 1.0f / 60.0f; p.interp = 12.0f; p.max_speed = 8.0f; p.thigh = 0.5f; p.shin = 0.5f; p.ankle_gain = 0.1f;
  p.contact_height = 0.85f; p.traction = 0.2f; p.drag = 0.98f; p.max_body_angle = 0.3f; p.upright_gain = 10.0f;
  const float lo[BIDI_ENV_JOINTS] = {-0.6f, -1.0f, -0.4f, -0.6f, -1.0f, -0.4f, -0.6f, -1.0f, -0.6f, -1.0f};
  const float hi[BIDI_ENV_JOINTS] = {0.6f, 0.2f, 0.4f, 0.6f, 0.2f, 0.4f, 0.6f, 0.2f, 0.6f, 0.2f};
  for (int j = 0; j < BIDI_ENV_JOINTS; j++) { p.jmin[j] = lo[j]; p.jmax[j] = hi[j]; }
  const int hip[4] = {0, 3, 6, 8}, knee[4] = {1, 4, 7, 9}, ankle[4] = {2, 5, -1, -1};
  for (int k = 0; k < 4; k++) { p.hip[k] = hip[k]; p.knee[k] = knee[k]; p.ankle[k] = ankle[k]; }
  return p;
}

// odd / even polynomials standing in for sin / cos on the joint range (the surrogate's own definition of them)
__host__ __device__ inline float env_sin(float x) { const float x2 = x * x; return x * (1.0f + x2 * (-1.0f / 6.0f + x2 * (1.0f / 120.0f))); }
__host__ __device__ inline float env_cos(float x) { const float x2 = x * x; return 1.0f + x2 * (-0.5f + x2 * (1.0f / 24.0f)); }

// foot of leg k relative to the hip: x forward, reach downward
__host__ __device__ inline void env_foot(const EnvParams& p, const float* ang, int k, float& fx, float& reach) {
  const float a0 = ang[p.hip[k]], a01 = a0 + ang[p.knee[k]];
  fx = p.thigh * env_sin(a0) + p.shin * env_sin(a01);
  reach = p.thigh * env_cos(a0) + p.shin * env_cos(a01);
  if (p.ankle[k] >= 0) reach = reach + p.ankle_gain * env_sin(ang[p.ankle[k]]);
}

// E: one agent's BIDI_ENV_FLOATS
__host__ __device__ inline void env_reset(const EnvParams& p, float* E) {
  for (int j = 0; j < BIDI_ENV_FLOATS; j++) E[j] = 0.0f;
  for (int k = 0; k < 4; k++) {
    float fx, reach;
    env_foot(p, E, k, fx, reach);
    E[18 + k] = fx;
    E[11 + k] = reach >= p.contact_height ? 1.0f : 0.0f;
  }
}

// Runner::motorUpdate(action, interpolateFactor) + the keep-upright rule + one tick
__host__ __device__ inline void env_tick(const EnvParams& p, float* E, const float* action) {
  for (int j = 0; j < BIDI_ENV_JOINTS; j++) {
    const float target = action[j] * (p.jmax[j] - p.jmin[j]) + p.jmin[j];
    float speed = p.interp * (target - E[j]);
    speed = bidi_min(p.max_speed, bidi_max(-p.max_speed, speed));
    E[j] = bidi_min(p.jmax[j], bidi_max(p.jmin[j], E[j] + p.dt * speed));
  }
  float push = 0.0f, touching = 0.0f, front = 0.0f, back = 0.0f;
  for (int k = 0; k < 4; k++) {
    float fx, reach;
    env_foot(p, E, k, fx, reach);
    const float c = reach >= p.contact_height ? 1.0f : 0.0f;
    const float vf = (fx - E[18 + k]) / p.dt;   // the foot's horizontal velocity relative to the body
    if (c != 0.0f && E[11 + k] != 0.0f) { push = push - vf; touching = touching + 1.0f; }  // in contact through the whole tick
    if (k == 1 || k == 3) front = front + c; else back = back + c;
    E[18 + k] = fx; E[11 + k] = c;
  }
  float vx = E[15];
  if (touching > 0.0f) vx = (1.0f - p.traction) * vx + p.traction * (push / touching);
  else vx = p.drag * vx;
  E[15] = vx;
  E[17] = E[17] + p.dt * vx;
  // the body pitches towards the side that lost its support; RunnerMain.cpp:248-249 keeps it upright
  float ang = E[10], av = E[16];
  av = 0.9f * av + p.dt * (3.0f * (0.1f * (back - front) - ang));
  if (fabsf(ang) > p.max_body_angle) av = -p.upright_gain * ang;
  E[16] = av;
  E[10] = ang + p.dt * av;
}

__global__ void k_env_reset(EnvParams p, float* env, int A) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < A) env_reset(p, env + (long long)a * BIDI_ENV_FLOATS);
}

// before simStep: reward, state inputs, clocks, fed-back noisy actions (RunnerMain.cpp:207-242).
// noise: [A][n_act] supplied N(0, 0.05) draws, or nullptr -> the engine's counter-based generator
__global__ void k_env_pre(EngineDev P, const float* __restrict__ env, float4 clocks, int clock0, int act0, int n_act,
                          const float* __restrict__ noise, float std, unsigned long long seed, unsigned long long step, long long first,
                          float* rewards) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= P.A) return;
  const float* E = env + (long long)a * BIDI_ENV_FLOATS;
  const float* pred = P.pred_out + (long long)a * P.n_in;
  float* in = P.in0 + (long long)a * P.n_in;
  rewards[a] = E[15];
  for (int j = 0; j < BIDI_ENV_STATE; j++) in[j] = E[j];
  in[clock0] = clocks.x; in[clock0 + 1] = clocks.y; in[clock0 + 2] = clocks.z; in[clock0 + 3] = clocks.w;
  for (int k = 0; k < n_act; k++) {
    float z;
    if (noise) z = noise[(long long)a * n_act + k];
    else {
      const unsigned long long h = mix64(seed ^ mix64(step * 0x9E3779B1ULL + (unsigned long long)(a + first) * n_act + k));
      const float r1 = fmaxf(u01(h), 5.9604645e-8f), r2 = u01(mix64(h));
      z = sqrtf(-2.0f * logf(r1)) * cospif(2.0f * r2) * std;
    }
    in[act0 + k] = bidi_clamp01(pred[act0 + k] * 0.5f + 0.5f + z);
  }
}

// after simStep: the new predictions are the motor targets (RunnerMain.cpp:244-249)
__global__ void k_env_post(EngineDev P, EnvParams p, float* env, int act0) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= P.A) return;
  const float* pred = P.pred_out + (long long)a * P.n_in;
  float action[BIDI_ENV_JOINTS];
  for (int k = 0; k < BIDI_ENV_JOINTS; k++) action[k] = bidi_clamp01(pred[act0 + k] * 0.5f + 0.5f);
  env_tick(p, env + (long long)a * BIDI_ENV_FLOATS, action);
}

} // namespace bidi
