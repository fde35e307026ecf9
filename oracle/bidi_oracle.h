/*
 * oracle/bidi_oracle.h -- TEST INFRASTRUCTURE.  C API of the CPU restatement of
 * the reference's per-timestep hierarchy update (bidi_oracle.c).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may use it;
 * the product (bidinet_b200/) never does.
 */
#ifndef BIDI_ORACLE_H
#define BIDI_ORACLE_H
#include "bidinet_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
void* orc_create(const bidi_config* cfg, const bidi_layer_desc* layers, unsigned seed);
void orc_destroy(void* h);
long long orc_array_len(void* h, int which, int layer);
int orc_get_array(void* h, int which, int layer, float* dst);
int orc_set_array(void* h, int which, int layer, const float* src);
int orc_conn_counts(void* h, int which, int layer, int* counts);
void orc_set_input(void* h, int i, float v);
void orc_set_inputs(void* h, const float* in, int n);
float orc_get_prediction(void* h, int i);
void orc_get_predictions(void* h, float* out, int n);
int orc_num_columns(void* h);
/* floats per step in each of u, v: 3 per column (neo::Agent), one per action (sdr::QPRSDR) */
int orc_num_noise(void* h);
/* exploration draws of the NEXT step: u,v [n_columns][3], column visit order */
int orc_peek_noise(void* h, float* u, float* v);
/* simStep drawing the exploration noise from the oracle's own mt19937 */
void orc_step(void* h, float reward, int learn);
/* simStep with the exploration noise supplied (the generator is not advanced) */
void orc_step_noise(void* h, float reward, const float* u, const float* v, int learn);
/* neo::PredictiveHierarchy::simStepGenerate: the N(0,1) draws of the next call, and the call itself */
int orc_num_generate_noise(void* h);
int orc_peek_generate(void* h, float* out);
void orc_step_generate(void* h, float noise);
void orc_run(void* h, const float* frames, int n_frames, int n_in, const float* rewards, int steps, int learn);
void orc_run_runner(void* h, const float* frames, int n_frames, int n_in, int steps, int n_fb, const int* src,
                    const int* dst, int learn);
#ifdef __cplusplus
}
#endif
#endif
