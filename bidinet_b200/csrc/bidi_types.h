// bidi_types.h -- plain structs shared by the host engine and the CUDA kernels.
//
// HBM layout.  Every agent owns one contiguous "blob" of
// floats; agent a's blob starts at blob + a*stride.  Inside a blob:
//   * per-unit state arrays are dense planes [unit], unit = x + y*width
//     (the reference's node order, sdr/RSDR.cpp:36-38)
//   * a connection family (feed-forward, recurrent, lateral, feedback,
//     predictive) is a stack of "slot planes" w[slot][unit]: slot (sx,sy) is a
//     position inside the receptive window, so the weights one warp of adjacent
//     units needs for a slot are adjacent in memory, and walking the slots
//     sx-outer / sy-inner visits a unit's weights in exactly the reference's
//     dx-outer / dy-inner list order (sdr/RSDR.cpp:48-62).  Slots that fall off
//     the target grid hold 0 and are never read.
//   * the per-node columns of neo::Agent keep the reference's dense
//     [node][cell][input] order (neo/Column.cpp:22-43), one contiguous run per
//     column.
// Layer-0 inputs and the input-space predictions live OUTSIDE the blobs in dense
// [agent][index] arrays so that setInput/getPrediction traffic is one contiguous
// host<->device copy.
#ifndef BIDI_TYPES_H
#define BIDI_TYPES_H

#include "bidinet_b200.h"

#define BIDI_MAX_NODE_LAYERS (BIDI_MAX_LAYERS + 1)
#define BIDI_COL_ROW 20 /* floats per cell row of a planar column block: 16 lateral weights + 4 scalars */

// One connection family: units of a uw x uh grid look at a window of radius r
// around a centre in a tw x th target grid.
struct WinDev {
  int uw, uh, tw, th;
  int r;             // -1: family absent
  int dxn, dyn;      // slot box: dxn = min(2r+1, tw), dyn = min(2r+1, th)
  int absx, absy;    // slot coordinate = absolute target coordinate (window spans the grid)
  int excl;          // centre excluded (recurrent, lateral)
  int nslots;        // dxn*dyn
  int U;             // uw*uh
  int max_cand;      // largest number of units whose window can contain one target (gather form)
  const int* cx;     // [uw] window centre per unit column, round_f32(ux * tw/uw)  (sdr/RSDR.cpp:40)
  const int* cy;     // [uh]
  // reverse map for gather-form reconstruction: target column tx is inside the
  // windows of unit columns gx_lo[tx]..gx_hi[tx] (cx is monotone, so it is a range)
  const int* gx_lo; const int* gx_hi;  // [tw]
  const int* gy_lo; const int* gy_hi;  // [th]
  long long w_off;   // float offset of w[slot][unit] inside the agent blob, -1 if none
  long long tr_off;  // eligibility traces, same shape, -1 if none
  // 0: slot planes, element (slot, unit) at w_off + slot*U + unit.  > 0: ROWS, element (slot, unit) at
  // w_off + unit*row_pitch + slot -- the layout of a layer that runs the dense fused coder (bidi_coder_dense.cuh):
  // one row per unit holding its feed-forward slots, then (rec.w_off = ff.w_off + padded feed-forward length) its
  // recurrent ones, identical to the kernel's shared-memory image
  int row_pitch;
};

// The columns (neo::Column) of one node layer.  Columns are stored in PAIRS (nodes
// 2p and 2p+1): everything the two columns read AND write each step lives in one
// contiguous, 16-byte aligned "pair record" with every element interleaved [..][2]
// (x = column 2p, y = column 2p+1), so that one TMA bulk copy brings the pair into
// shared memory, one takes it back, and a lane works on both columns at once with
// packed f32x2 arithmetic:
//   W2[cells][S][2]     feed-forward weights, common row stride S >= both nIn, S even
//                       with S/2 odd (conflict-free 128-bit shared loads when lane i
//                       walks row i); pad entries are 0
//   latT2[cells][Cq][2] lateral weights TRANSPOSED (latT[k][i] = lat[i][k]), Cq = cells
//                       rounded up to 4
//   10 vectors [Cq][2]  threshold, q w, q trace, 3 action w, 3 action traces, spikePrev
//   err2[S][2]          _reconstructionError
//   scal2[4][2]         prevValue, pad
// A layer with an odd number of nodes ends with a phantom column of zeros.
// Columns of 16 cells (the reference's default, neo/Agent.h:22) use the PLANAR variant of
// the record instead: column 2p's block followed by column 2p+1's, without the [2]
// interleave.  A block is
//   W[16][S - rh]          weight rows (the part not kept in registers), S = 4 (mod 8)
//   cell[16][BIDI_COL_ROW] per cell i: lat[i][0..15] (row-major), threshold, q w, q trace, spikePrev
//                          -- the lane that owns the cell reads its lateral row and its four scalars
//                          with 128-bit accesses, rows 20 floats apart fall in distinct banks
//   3 action w [16], 3 action traces [16], err[S], scal[4]
// bidi_columns16.cuh works on it with lane = (column, cell).
// Write-only per-step outputs (inputs, activation, spike, state, action states) stay
// in plain planes.
struct ColumnsDev {
  int n, cells, cq, tot_in; // nodes, cells per column, cells rounded up to 4, sum of nIn
  int max_in;               // largest nIn
  int planar;               // record layout: 0 pair-interleaved (k_columns), 1 planar (k_columns16)
  int rh;                   // planar only: the first rh inputs of every weight row form the record's lane-major
                            // "register part" [rh/4][32 lanes][4] that precedes the two column blocks (0, 32 or 64)
  int max_s, max_rec;       // largest pair stride / largest shared-memory part of a pair record (floats)
  const int* in_off;        // [n+1] prefix sums of nIn (same for every agent)
  const int* in_src;        // [tot_in] blob offset each column input is gathered from (neo/Agent.cpp:159-169)
  // The inputs of a layer's columns come from at most three short planes of the agent (the upper columns' _signal
  // actions, the upper prediction states, the layer's own coder states).  stage_n > 0: bidi_columns16.cuh copies those
  // planes into shared memory with independent coalesced loads and picks the inputs there (in_loc[j] = position of
  // input j in the staging), instead of a dependent index -> value chain of global loads per input.
  int stage_n;              // floats staged (0: gather straight from the blob through in_src)
  int stage_len[3], stage_step[3];
  long long stage_off[3];   // segment k: stage[base_k + q] = blob[stage_off[k] + q*stage_step[k]], q < stage_len[k]
  const unsigned short* in_loc;  // [tot_in]
  const int* rec_off;       // [n_pairs+1] pair record offsets (floats) relative to rec_base
  const int* stride;        // [n_pairs] S
  long long rec_base;       // blob offset of the first record
  long long inputs, act, spike, state, a_state, a_expl; // planes
  int noise_base;           // first column index of this layer in the visit order (x3 for the noise arrays)
  // neo/Column.cpp:46 simStep arguments
  float sparsity, gamma, leak, a_ff, a_lat, a_thr, a_q, a_act, gamma_lambda, expl_std, expl_break;
  int iter;
};

// The deep::SDRRL units (deep/SDRRL.h:8-62) of one node layer of a deep::CSRL: plain planes in the reference's
// serialised order ([node][cell][input], [node][cell][cell], [node][cell], [node][action], [node][cell][action]).
// Per-pair geometry of one layer's columns as a kernel argument (constant bank): the record offset, the row stride and
// the input range of a pair are then known without a global load, so the record's bulk copy starts at once.
// n = 0: the layer has more pairs than fit; the kernel reads the ColumnsDev tables instead.
#define BIDI_PAIRTAB_MAX 48
struct ColPairTab {
  int n;
  int rec_off[BIDI_PAIRTAB_MAX + 1];
  int in_off[2 * BIDI_PAIRTAB_MAX + 1];
  short stride[BIDI_PAIRTAB_MAX];
};

struct SdrrlDev {
  int n, cells, na, n_rec;   // nodes, cells per unit, 2*(3 + recurrent inputs) action entries, recurrent inputs
  int tot_in, max_in;
  const int* in_off;         // [n+1] prefix sums of numStates
  const int* in_src;         // [tot_in] blob offset an input is gathered from; -(1+k): _lastLayerRewardOffsets[k] + reward
  long long inputs, recon_err, ff_w, lat_w, thr, act, spike, spike_prev, state, cell_as, cell_ae, q_w, q_tr;
  long long a_state, a_expl, a_err, a_w, a_tr, prev_value, avg_surprise;
  int noise_base;            // first entry of this layer's units in an agent's exploration-noise arrays
  float sparsity, gamma, leak, a_ff, a_lat, a_thr, a_q, a_act, derive_alpha, gamma_lambda, expl_std, expl_break, surprise_decay;
  int settle, measure, derive_iter;
};

// One layer: the sparse coder (visible -> hidden) and its prediction nodes.
// Index n_layers holds only the input prediction nodes (pn_*, fb, col).
struct LayerDev {
  int vw, vh, hw, hh, nv, nh;
  // Row strip a launch works on (hidden rows [row0,row1), visible rows [vrow0,vrow1)).
  // The whole grid except when one oversized layer is split across GPUs (bidi_strip_*).
  int row0, row1, vrow0, vrow1;
  int gen_off;                      // simStepGenerate: offset of this layer's N(0,1) draws inside an agent's block
  // coder unit planes (offsets into the blob)
  long long vis_input, vis_recon;   // vis_input is -1 for layer 0 (dense input array)
  long long thr, state, state_prev, act, recon, spike, spike_prev;
  WinDev ff, rec, lat;
  // prediction nodes
  int pn;                           // number of prediction nodes (nh, or n_in for the input nodes)
  long long p_state, p_state_prev, p_act, p_act_prev, p_baseline, p_mod; // p_mod: attention / reward scratch
  WinDev fb, pred;
  // sdr::QPRSDR's Q layer over this layer's grid (bidi_qnet.cuh): same window geometry as `ff`;
  // weight/trace slot planes, per-node planes; q_mask[slot][unit] = 1 where the reference keeps the
  // connection (one copy for all agents, outside the blobs)
  long long q_state, q_err, q_bias_w, q_bias_tr, q_w_off, q_tr_off;
  const float* q_mask;
  ColumnsDev col;
  SdrrlDev sd;
  // layers whose windows span the grid (bidi_coder_dense.cuh): per (unit u, row quad q) the four bits "slot 4q+e of
  // the row [feed-forward | recurrent | pad] is one of u's own connections", [unit][pitch/4]; one table for all agents
  const int* dense_mask;
  // hyper-parameters (bidi_layer_desc)
  float learn_ff, learn_rec, learn_lat, learn_threshold, learn_fb, learn_pred, sparsity;
  float avg_surprise_decay, attention_factor;
  int iter_settle, iter_measure;
  float leak, lambda, weight_decay, max_weight_delta, baseline_decay, sensitivity;
};

struct EngineDev {
  int variant, L, A, n_in;
  float* blob;            // [A][stride]
  long long stride;       // floats per agent
  float* in0;             // [A][n_in] layer-0 visible input
  float* pred_out;        // [A][n_in] getPrediction values
  const float* rewards;   // [A]
  const float* noise_u;   // [A][ncol*3]
  const float* noise_v;
  float* col_actions;     // [A][ncol*3] exploratory actions in visit order (getAction)
  int ncol;
  const LayerDev* layers; // device array [L+1]
  // {1,1}, {-0,-0}, {-1,-1} as packed f32x2 bit patterns.  Passed at run time so that
  // ptxas cannot fold fma(a,b,-0) / fma(a,1,c) back into a fused multiply-add.
  unsigned long long k_one2, k_negzero2, k_negone2;
  // neo::PredictiveHierarchy::simStepGenerate: the coders' excitation starts from gen_normals[...] * *gen_scale
  int gen_on, gen_per_agent;
  const float* gen_normals;               // [A][gen_per_agent]: layer 0 first; iteration outer, unit inner
  const float* gen_scale;                 // the `noise` argument, in device memory so that the step graph need not change
  // sdr::QPRSDR (bidi_qnet.cuh)
  int q_on, q_half, q_top, q_iters;       // 0/1; actions; nodes of the top layer; _actionDeriveIterations
  long long q_qc_w, q_qc_tr, q_act, q_prev; // blob offsets: _qConnections (w, trace), action nodes [4][2*half], _prevValue
  const int* q_act_index;                 // [n_in] input cell -> action node, -1 if none
  const int* q_a_input;                   // [2*half] action node -> input cell
  float q_alpha, q_action_alpha, q_derive_alpha, q_expl_break, q_gamma, q_gamma_lambda, q_expl_std;
  // deep::CSRL (bidi_csrl.cuh): the hierarchy runs the ITERATIVE coders; every node owns a deep::SDRRL unit
  int csrl, n_noise, csrl_n_actions;
  long long csrl_reward_offsets;          // blob offset of _lastLayerRewardOffsets
  const int* csrl_action_idx;             // [csrl_n_actions] input cells of type _action, ascending
  const int* csrl_is_action;              // [n_in]
  float csrl_expl_std, csrl_expl_break;   // the CSRL object's own (input nodes)
};

#endif
