/* The caller's side of the reference's headless runner loop (RunnerMain.cpp:235-251; the same
 * arithmetic as oracle/ref_harness.cpp ref_run_runner), for every agent of a batch in one pass:
 *   action_k     = getPrediction(fb0 + k) * 0.5 + 0.5
 *   input[fb0+k] = clamp01(action_k + noise_k)          (the noisy action is fed back)
 *   reward       = mean_k clamp01(action_k) - 0.5        (stands in for the Box2D body velocity)
 * Harness code for bench.py's end-to-end loop (the environment is not part of the engine): the
 * reference does this in compiled code, so the GPU arm's host side does too.
 * pred, x: [agent][n_in]; noise: [agent][n_fb] N(0, 0.05) draws; rewards: [agent]. */
static float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

void runner_env_step(int n_agents, int n_in, int fb0, int n_fb, const float* pred, const float* noise, float* x, float* rewards) {
  for (int a = 0; a < n_agents; a++) {
    const float* p = pred + (long long)a * n_in + fb0;
    const float* z = noise + (long long)a * n_fb;
    float* xi = x + (long long)a * n_in + fb0;
    float sum = 0.0f;
    for (int k = 0; k < n_fb; k++) {
      const float act = p[k] * 0.5f + 0.5f;
      sum += clamp01(act);
      xi[k] = clamp01(act + z[k]);
    }
    rewards[a] = n_fb ? sum / (float)n_fb - 0.5f : 0.0f;
  }
}
