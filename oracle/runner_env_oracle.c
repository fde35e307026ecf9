/*
 * oracle/runner_env_oracle.c -- TEST INFRASTRUCTURE.  CPU twin of the batched headless runner environment
 * (bidinet_b200/csrc/bidi_env.cuh), restated independently from its definition: a planar stand-in for the Box2D
 * quadruped of the reference (runner/Runner.cpp: getStateVector :260-339, motorUpdate :341) under the RunnerMain
 * loop (RunnerMain.cpp:204-251).  One instance = one agent's world.  Only tests/ use it; the product never does.
 *
 * Definition (every operation in fp32, in this order):
 *   joints j = 0..9 (left back hip/knee/ankle, left front hip/knee/ankle, right back hip/knee, right front hip/knee),
 *     limits: hips [-0.6, 0.6], knees [-1.0, 0.2], ankles [-0.4, 0.4]
 *   motorUpdate: target = action*(max - min) + min; speed = clamp(12*(target - angle), +-8); angle = clamp(angle + dt*speed)
 *   leg k: fx = 0.5*S(hip) + 0.5*S(hip + knee); reach = 0.5*C(hip) + 0.5*C(hip + knee) (+ 0.1*S(ankle) if it has one)
 *     with S(x) = x*(1 + x^2*(-1/6 + x^2/120)), C(x) = 1 + x^2*(-1/2 + x^2/24); contact = reach >= 0.85
 *   feet in contact before AND after the tick push the body: push = sum -(fx - fx_prev)/dt over them;
 *     vx = 0.8*vx + 0.2*push/count, or 0.98*vx when none touches; x += dt*vx
 *   pitch: av = 0.9*av + dt*3*(0.1*(back contacts - front contacts) - angle); if |angle| > 0.3: av = -10*angle
 *     (RunnerMain.cpp:248-249); angle += dt*av
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float angle[10];
  float body_angle, body_av, vx, x;
  float contact[4], foot_x[4];
  int steps;
} renv;

static const float kLo[10] = {-0.6f, -1.0f, -0.4f, -0.6f, -1.0f, -0.4f, -0.6f, -1.0f, -0.6f, -1.0f};
static const float kHi[10] = {0.6f, 0.2f, 0.4f, 0.6f, 0.2f, 0.4f, 0.6f, 0.2f, 0.6f, 0.2f};
static const int kHip[4] = {0, 3, 6, 8}, kKnee[4] = {1, 4, 7, 9}, kAnkle[4] = {2, 5, -1, -1};

static float S(float x) { float x2 = x * x; return x * (1.0f + x2 * (-1.0f / 6.0f + x2 * (1.0f / 120.0f))); }
static float Cc(float x) { float x2 = x * x; return 1.0f + x2 * (-0.5f + x2 * (1.0f / 24.0f)); }
static float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void foot(const renv* e, int k, float* fx, float* reach) {
  float h = e->angle[kHip[k]], hk = h + e->angle[kKnee[k]];
  *fx = 0.5f * S(h) + 0.5f * S(hk);
  *reach = 0.5f * Cc(h) + 0.5f * Cc(hk);
  if (kAnkle[k] >= 0) *reach = *reach + 0.1f * S(e->angle[kAnkle[k]]);
}

void* renv_create(void) {
  renv* e = (renv*)calloc(1, sizeof(renv));
  for (int k = 0; k < 4; k++) {
    float fx, reach;
    foot(e, k, &fx, &reach);
    e->foot_x[k] = fx;
    e->contact[k] = reach >= 0.85f ? 1.0f : 0.0f;
  }
  return e;
}
void renv_destroy(void* h) { free(h); }

float renv_reward(void* h) { return ((renv*)h)->vx; }

/* Runner::getStateVector: 10 joint angles, body angle, 4 contacts */
void renv_state(void* h, float* out15) {
  renv* e = (renv*)h;
  memcpy(out15, e->angle, sizeof(float) * 10);
  out15[10] = e->body_angle;
  memcpy(out15 + 11, e->contact, sizeof(float) * 4);
}

/* the four clock inputs of RunnerMain.cpp:235-236 at the environment's step counter */
void renv_clocks(void* h, float* out4) {
  renv* e = (renv*)h;
  for (int a = 0; a < 4; a++) out4[a] = sinf(e->steps / 60.0f * 2.0f * a * 2.0f * 3.141596f) * 0.5f + 0.5f;
}

/* Runner::motorUpdate(action, 12) + keep upright + one tick of 1/60 s */
void renv_motor_update(void* h, const float* action10) {
  renv* e = (renv*)h;
  const float dt = 1.0f / 60.0f;
  for (int j = 0; j < 10; j++) {
    float target = action10[j] * (kHi[j] - kLo[j]) + kLo[j];
    float speed = clampf(12.0f * (target - e->angle[j]), -8.0f, 8.0f);
    e->angle[j] = clampf(e->angle[j] + dt * speed, kLo[j], kHi[j]);
  }
  float push = 0.0f, n = 0.0f, front = 0.0f, back = 0.0f;
  for (int k = 0; k < 4; k++) {
    float fx, reach;
    foot(e, k, &fx, &reach);
    float c = reach >= 0.85f ? 1.0f : 0.0f;
    float vf = (fx - e->foot_x[k]) / dt;
    if (c != 0.0f && e->contact[k] != 0.0f) { push = push - vf; n = n + 1.0f; }
    if (k == 1 || k == 3) front = front + c; else back = back + c;
    e->foot_x[k] = fx;
    e->contact[k] = c;
  }
  if (n > 0.0f) e->vx = (1.0f - 0.2f) * e->vx + 0.2f * (push / n);
  else e->vx = 0.98f * e->vx;
  e->x = e->x + dt * e->vx;
  float ang = e->body_angle, av = e->body_av;
  av = 0.9f * av + dt * (3.0f * (0.1f * (back - front) - ang));
  if (fabsf(ang) > 0.3f) av = -10.0f * ang;
  e->body_av = av;
  e->body_angle = ang + dt * av;
  e->steps++;
}

/* the layout bidi_env_get_state returns: 10 angles, body angle, 4 contacts, vx, angular velocity, x, 4 foot x, pad */
void renv_dump(void* h, float* out24) {
  renv* e = (renv*)h;
  memset(out24, 0, sizeof(float) * 24);
  memcpy(out24, e->angle, sizeof(float) * 10);
  out24[10] = e->body_angle;
  memcpy(out24 + 11, e->contact, sizeof(float) * 4);
  out24[15] = e->vx; out24[16] = e->body_av; out24[17] = e->x;
  memcpy(out24 + 18, e->foot_x, sizeof(float) * 4);
}
