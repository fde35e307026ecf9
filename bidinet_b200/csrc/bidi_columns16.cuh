// bidi_columns16.cuh -- neo::Column::simStep (neo/Column.cpp:46-169) specialised for the
// reference's default of 16 cells per column (neo/Agent.h:22), on the PLANAR pair record
// (bidi_types.h, ColumnsDev).
//
// One warp owns one column pair and every lane is busy in every phase:
//   lane = (column c = lane / 16, cell or input slot i = lane % 16).
//   * sweep: lane (c, i) walks row i of column c's weights sequentially -- the
//     reference's summation order -- four inputs per 128-bit shared-memory load; the
//     products are packed (FFMA2), the running sum is one scalar FADD chain per lane.
//     Rows are S = 4 (mod 8) floats apart, so the eight rows a quarter warp reads at once
//     fall in distinct banks.
//   * reconstruction and the element-wise weight update: lane (c, i) owns the input
//     PAIRS i, i+16, ... of column c (64-bit accesses, packed arithmetic), looping only
//     over its own column's spiking cells in ascending order.
// The kernel is bound by instruction issue, not by HBM (profiles/r1_columns_notes.md), so
// what matters is the number of warp instructions per pair: the cell count is a compile
// time constant, and nothing is computed for a column that did not spike.
//
// Residency: the kernel is bound by how many records an SM holds (its time goes as
// 0.7 ms + 22 ms / resident warps per SM on C3, measured by padding shared memory), and a
// record is 80 % weights.  So the first RH inputs of every weight row live in REGISTERS of
// the lane that owns the row (RH = 64 or 32 floats per lane), loaded and stored with
// coalesced 128-bit global accesses from a lane-major part of the record, and only the
// rest of the row goes through shared memory with the TMA copy.  The sweep and the
// element-wise update use the registers directly; for the reconstruction the few rows that
// spiked are parked in a small shared scratch (four rows per column at a time).
//
// Rounding rules are those of bidi_columns.cuh: no fused multiply-add anywhere.
#pragma once
#include "bidi_columns.cuh"

namespace bidi {

__device__ __forceinline__ float lo32(u64 v) { float x, y; upk(v, x, y); return x; }
__device__ __forceinline__ float hi32(u64 v) { float x, y; upk(v, x, y); return y; }

// SMALL: every row of the layer is at most 32 inputs long (the top layer of the RunnerMain hierarchy: 16 + pad): the
// gather and the reconstruction are unrolled for one trip instead of nine / five (their loops still cover any length).
template <int RH, bool SMALL = false>
__global__ void __launch_bounds__(32) k_columns16(EngineDev P, LayerDev L, int l, ColPairTab T) {
  extern __shared__ __align__(128) float smem[];
  constexpr int CC = 16;
  constexpr int RB = 4;              // spiking rows per column parked in the scratch at a time
  constexpr int RQ = RH / 4, RP = RH / 2, RT = RH / 32;  // register part: 128-bit chunks, input pairs, 16-pair trips
  constexpr int RS = RH ? RH + 4 : 0;         // scratch row pitch: +16 bytes, so that rows parked by different lanes do not share banks
  const ColumnsDev& C = L.col;
  const Pair F{P.k_one2, P.k_negzero2, P.k_negone2};
  const int lane = threadIdx.x, c = lane >> 4, i = lane & 15;
  const int a = blockIdx.z * gridDim.y + blockIdx.y, pair = blockIdx.x;  // consecutive CTAs are the pairs of one agent
  if (a >= P.A) return;
  const int node = 2 * pair + c;
  const bool has = node < C.n;
  float* B = blob_of(P, a);

  // the pair's geometry: from the constant bank where the layer's table fits there (no load between launch and bulk copy)
  const bool tab = T.n > 0;
  const int S = tab ? (int)T.stride[pair] : C.stride[pair], SH = S - RH;  // row length; the part of it that lives in shared memory
  const int ro0 = tab ? T.rec_off[pair] : C.rec_off[pair], ro1 = tab ? T.rec_off[pair + 1] : C.rec_off[pair + 1];
  const int R = ro1 - ro0 - 32 * RH;                     // floats of the shared-memory part of the record
  const int blk = R >> 1;
  float* rec = smem;
  float* in_s = smem + C.max_rec + c * C.max_s;          // gathered inputs of column c, [S], pad 0
  float* st_s = smem + C.max_rec + 2 * C.max_s + c * CC;  // final cell states of column c
  float* kf_s = st_s + 2 * CC;                            // feed-forward learning factors of column c
  float* stg = smem + C.max_rec + 2 * C.max_s + 4 * CC;   // staging of the input planes, then the scratch rows
  float* scr = stg + c * (RB * RS);                       // parked rows of column c, [RB][RS]
  uint64_t* bar = reinterpret_cast<uint64_t*>(stg + column_scratch_floats(C));
  float* g_reg = B + C.rec_base + ro0;                   // register part: [RQ][32 lanes][4]
  float* g_rec = g_reg + 32 * RH;

  // ---- every global load of the prologue is issued before the first one is waited for: the positions of the inputs and
  // the planes they come from first (their addresses hang on nothing but the constant bank), then the record's bulk copy
  // and the register part.  (Issued after the register part, the positions waited behind it -- a second global round trip
  // per pair, 6-8 % of all warp samples.)
  if (lane == 0) mbar_init(bar, 1);
  const int in0 = !has ? 0 : (tab ? T.in_off[node] : C.in_off[node]);
  const int nin = !has ? 0 : (tab ? T.in_off[node + 1] : C.in_off[node + 1]) - in0;
  constexpr int GT = SMALL ? 2 : 9;  // inputs per lane of a gather trip: 144 (32) inputs per trip
  const bool staged = C.stage_n > 0;
  const int n0 = C.stage_len[0], n01 = n0 + C.stage_len[1], sn = C.stage_n;
  auto plane_off = [&](int k) -> long long {
    return k < n0 ? C.stage_off[0] + (long long)k * C.stage_step[0]
         : k < n01 ? C.stage_off[1] + (long long)(k - n0) * C.stage_step[1]
                   : C.stage_off[2] + (long long)(k - n01) * C.stage_step[2];
  };
  int lv[GT];
  float sv[4];
  if (staged) {
    const unsigned short* loc = C.in_loc + in0;
#pragma unroll
    for (int t = 0; t < GT; t++) { const int j = i + 16 * t; lv[t] = j < nin ? (int)loc[j] : -1; }
#pragma unroll
    for (int t = 0; t < 4; t++) { const int k = lane + 32 * t; sv[t] = k < sn ? B[plane_off(k)] : 0.0f; }
  }
  __syncwarp();
  if (lane == 0) {
    mbar_expect_tx(bar, (uint32_t)R * 4u);
    tma_load_1d(rec, g_rec, (uint32_t)R * 4u, bar);
  }
  // the first RH inputs of this lane's weight row, straight from HBM into registers
  u64 wr[RP > 0 ? RP : 1];
#pragma unroll
  for (int q = 0; q < RQ; q++) {
    const ulonglong2 v = reinterpret_cast<const ulonglong2*>(g_reg)[q * 32 + lane];
    wr[2 * q] = v.x; wr[2 * q + 1] = v.y;
  }
  // sections of column c's block
  float* W = rec + c * blk;          // [16][SH]
  float* cellr = W + CC * SH + i * BIDI_COL_ROW;  // this lane's cell row: lat[i][0..15], threshold, q w, q trace, spikePrev
  float* v_aw = W + CC * SH + CC * BIDI_COL_ROW;  // [3][16]
  float* v_atr = v_aw + 3 * CC;
  float* err = v_atr + 3 * CC;       // [S]
  float* scal = err + S;             // [4]

  // ---- gather column c's inputs while the copy is in flight (neo/Agent.cpp:159-169, :217-220)
  const float reward = P.rewards[a];
  float nz_u = 0.0f, nz_v = 0.0f;
  const long long nk = ((long long)a * P.ncol + C.noise_base + node) * 3 + i;
  if (has && i < 3) { nz_u = P.noise_u[nk]; nz_v = P.noise_v[nk]; }
  if (staged) {
    // the few planes every input of this layer comes from go to shared memory with independent, coalesced loads
    // and the inputs are picked there
    const unsigned short* loc = C.in_loc + in0;
#pragma unroll
    for (int t = 0; t < 4; t++) { const int k = lane + 32 * t; if (k < sn) stg[k] = sv[t]; }
    for (int k0 = 128; k0 < sn; k0 += 128) {
#pragma unroll
      for (int t = 0; t < 4; t++) { const int k = k0 + lane + 32 * t; sv[t] = k < sn ? B[plane_off(k)] : 0.0f; }
#pragma unroll
      for (int t = 0; t < 4; t++) { const int k = k0 + lane + 32 * t; if (k < sn) stg[k] = sv[t]; }
    }
    __syncwarp();
    for (int j0 = 0; j0 < S; j0 += 16 * GT) {
      if (j0 > 0) {
#pragma unroll
        for (int t = 0; t < GT; t++) { const int j = j0 + i + 16 * t; lv[t] = j < nin ? (int)loc[j] : -1; }
      }
#pragma unroll
      for (int t = 0; t < GT; t++) {
        const int j = j0 + i + 16 * t;
        if (j < S) in_s[j] = lv[t] >= 0 ? stg[lv[t]] : 0.0f;
      }
    }
  } else {
    const int* src = C.in_src + in0;
    for (int j0 = 0; j0 < S; j0 += 16 * GT) {
      float gv[GT];
#pragma unroll
      for (int t = 0; t < GT; t++) { const int j = j0 + i + 16 * t; gv[t] = j < nin ? B[src[j]] : 0.0f; }
#pragma unroll
      for (int t = 0; t < GT; t++) {
        const int j = j0 + i + 16 * t;
        if (j < S) in_s[j] = gv[t];  // (the COL_INPUT plane is not written: the host regathers it on demand)
      }
    }
  }
  __syncwarp();
  mbar_wait(bar, 0);

  // ---- iterate (neo/Column.cpp:58-103)
  float thr, sp;
  {
    const float4 c4 = *reinterpret_cast<const float4*>(cellr + CC);
    thr = c4.x; sp = c4.w;
  }
  const ulonglong2* lat4 = reinterpret_cast<const ulonglong2*>(cellr);
  unsigned m = (__ballot_sync(0xffffffffu, sp != 0.0f) >> (16 * c)) & 0xffffu;  // column c's spiking cells
  float act = 0.0f, st = 0.0f, counter = 0.0f, mult = 0.0f;
  const float oml = 1.0f - C.leak;
  const ulonglong2* w4 = reinterpret_cast<const ulonglong2*>(W + i * SH);
  const ulonglong2* e4 = reinterpret_cast<const ulonglong2*>(err);
  const int n4 = SH >> 2, n2 = S >> 1;
  const u64* in2 = reinterpret_cast<const u64*>(in_s);
  u64* err2 = reinterpret_cast<u64*>(err);
  constexpr int NP = SMALL ? 1 : 5;  // input pairs per lane per trip: 16*5*2 = 160 (32) inputs
  for (int it = 0; it < C.iter; it++) {
    // inhibition += lat * spikePrev over the cells that spiked (the others add +-0; an absent entry adds +0,
    // which leaves a sum that started at +0 unchanged).  Done before the sweep, four loads in flight, so that
    // their latency is not on the path between the sweep and the threshold test
    float inh = 0.0f;
    if (m) {
#pragma unroll
      for (int q = 0; q < 4; q++) {  // the lane's lateral row, four 128-bit loads; ascending cell index
        const ulonglong2 v = lat4[q];
        if (m & (1u << (4 * q))) inh += lo32(v.x);
        if (m & (2u << (4 * q))) inh += hi32(v.x);
        if (m & (4u << (4 * q))) inh += lo32(v.y);
        if (m & (8u << (4 * q))) inh += hi32(v.y);
      }
    }
    // excitation += w * err, sequentially over the inputs.  The shared-memory loads run two quads ahead of
    // the add chain (a quad's four chained adds are 16 cycles, a load is ~30); the last two prefetches of a
    // row read past its end (inside the slab) and are never used
    float exc = 0.0f;
    ulonglong2 eq0 = e4[0], eq1 = e4[1], wq0, wq1;
    if (RH == 0) { wq0 = w4[0]; wq1 = w4[1]; }
#pragma unroll
    for (int q = 0; q < RQ; q += 2) {
      {
        const u64 p01 = F.mul(wr[2 * q], eq0.x), p23 = F.mul(wr[2 * q + 1], eq0.y);
        eq0 = e4[q + 2];
        if (q + 2 == RQ) wq0 = w4[0];
        exc += lo32(p01); exc += hi32(p01); exc += lo32(p23); exc += hi32(p23);
      }
      {
        const u64 p01 = F.mul(wr[2 * q + 2], eq1.x), p23 = F.mul(wr[2 * q + 3], eq1.y);
        eq1 = e4[q + 3];
        if (q + 2 == RQ) wq1 = w4[1];
        exc += lo32(p01); exc += hi32(p01); exc += lo32(p23); exc += hi32(p23);
      }
    }
    {
      int j = 0;
      for (; j + 2 <= n4; j += 2) {
        {
          const u64 p01 = F.mul(wq0.x, eq0.x), p23 = F.mul(wq0.y, eq0.y);
          wq0 = w4[j + 2]; eq0 = e4[RQ + j + 2];
          exc += lo32(p01); exc += hi32(p01); exc += lo32(p23); exc += hi32(p23);
        }
        {
          const u64 p01 = F.mul(wq1.x, eq1.x), p23 = F.mul(wq1.y, eq1.y);
          wq1 = w4[j + 3]; eq1 = e4[RQ + j + 3];
          exc += lo32(p01); exc += hi32(p01); exc += lo32(p23); exc += hi32(p23);
        }
      }
      if (j < n4) {
        const u64 p01 = F.mul(wq0.x, eq0.x), p23 = F.mul(wq0.y, eq0.y);
        exc += lo32(p01); exc += hi32(p01); exc += lo32(p23); exc += hi32(p23);
      }
    }
    float av = oml * act + exc - inh;
    if (av > thr) { av = 0.0f; sp = 1.0f; } else sp = 0.0f;
    st += sp;
    act = av;
    m = (__ballot_sync(0xffffffffu, sp != 0.0f) >> (16 * c)) & 0xffffu;  // also orders the err reads above
    counter += 1.0f;
    mult = 1.0f / counter;
    // reconstruction from the spiking cells, ascending cell index (neo/Column.cpp:93-102):
    // recon += (spike*multiplier) * w ; error = input - recon
    const u64 mult2 = pk(mult, mult);
    const int ns = n2 - RP;  // input pairs in the shared-memory part of a row
    // the sums over the (few) spiking rows are packed: NP independent accumulators per lane hide the packed add's latency
    u64 r2[NP];
#pragma unroll
    for (int t = 0; t < NP; t++) r2[t] = 0ull;
    if (RH > 0) {
      // spiking rows RB at a time per column: their owners park the register part in the scratch,
      // then every lane adds them to its input pairs, ascending cell index
      const unsigned bal = __ballot_sync(0xffffffffu, sp != 0.0f);
      const int most = max(__popc(bal & 0xffffu), __popc(bal >> 16));
      const int rank = __popc(m & ((1u << i) - 1u));
      unsigned mm = m;
      for (int b0 = 0; b0 < most; b0 += RB) {
        if (sp != 0.0f && rank >= b0 && rank < b0 + RB) {
          ulonglong2* dst = reinterpret_cast<ulonglong2*>(scr + (rank - b0) * RS);
#pragma unroll
          for (int q = 0; q < RQ; q++) dst[q] = make_ulonglong2(wr[2 * q], wr[2 * q + 1]);
        }
        __syncwarp();
        for (int b = 0; b < RB && mm; b++, mm &= mm - 1) {
          const int k = __ffs(mm) - 1;
          const u64* rrow = reinterpret_cast<const u64*>(scr + b * RS) + i;
#pragma unroll
          for (int t = 0; t < RT; t++) r2[t] = F.add2(F.mul(mult2, rrow[16 * t]), r2[t]);
          const u64* row = reinterpret_cast<const u64*>(W + k * SH) + i;
#pragma unroll
          for (int t = RT; t < NP; t++)
            if (i + 16 * (t - RT) < ns) r2[t] = F.add2(F.mul(mult2, row[16 * (t - RT)]), r2[t]);
        }
        __syncwarp();
      }
    } else {
      for (unsigned mm = m; mm; mm &= mm - 1) {
        const u64* row = reinterpret_cast<const u64*>(W + (__ffs(mm) - 1) * SH) + i;
#pragma unroll
        for (int t = 0; t < NP; t++)
          if (i + 16 * t < ns) r2[t] = F.add2(F.mul(mult2, row[16 * t]), r2[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < NP; t++) {
      const int p = i + 16 * t;
      if (p < n2) err2[p] = F.sub2(in2[p], r2[t]);
    }
    // rows longer than 16*NP pairs: the rest of the shared-memory part
    for (int p0 = 16 * NP; p0 < n2; p0 += 16 * NP) {
      u64 s2[NP];
#pragma unroll
      for (int t = 0; t < NP; t++) s2[t] = 0ull;
      for (unsigned mm = m; mm; mm &= mm - 1) {
        const u64* row = reinterpret_cast<const u64*>(W + (__ffs(mm) - 1) * SH) + (p0 - RP) + i;
#pragma unroll
        for (int t = 0; t < NP; t++)
          if (p0 + i + 16 * t < n2) s2[t] = F.add2(F.mul(mult2, row[16 * t]), s2[t]);
      }
#pragma unroll
      for (int t = 0; t < NP; t++) {
        const int p = p0 + i + 16 * t;
        if (p < n2) err2[p] = F.sub2(in2[p], s2[t]);
      }
    }
    __syncwarp();
  }
  st *= mult;
  st_s[i] = st;
  __syncwarp();

  // ---- value and actions, sequential over the cells (neo/Column.cpp:111-123):
  // lanes 0..2 of each half accumulate one action each, lane 3 the value
  float accum = 0.0f;
  if (i < 4) {  // q weights are element 17 of the cell rows
    const float* wv = i < 3 ? v_aw + i * CC : cellr - i * BIDI_COL_ROW + CC + 1;
    const int ws = i < 3 ? 1 : BIDI_COL_ROW;
#pragma unroll
    for (int k = 0; k < CC; k++) accum += wv[k * ws] * st_s[k];
  }
  const float q = __shfl_sync(0xffffffffu, accum, 16 * c + 3);
  float as = 0.0f, ae = 0.0f;
  if (i < 3) {
    as = bidi_sigmoid(accum);
    ae = nz_u < C.expl_break ? nz_v : bidi_clamp01(as + nz_v);
    if (has) {
      B[C.a_state + node * 3 + i] = as;
      B[C.a_expl + node * 3 + i] = ae;
      P.col_actions[nk] = ae;
    }
  }
  // tdError = reward + gamma*q - prevValue
  const float td = reward + C.gamma * q - scal[0];
  const float q_alpha_td = C.a_q * td, act_alpha_td = C.a_act * td;
  const float gl = C.gamma_lambda;

  // ---- learning (neo/Column.cpp:139-166), all inside the shared-memory record
  {  // the cell's four scalars: threshold += alpha*(state - sparsity), q weight and trace, spikePrev
    float4 c4 = *reinterpret_cast<const float4*>(cellr + CC);
    const float tr = c4.z;
    c4.x = thr + C.a_thr * (st - C.sparsity);
    c4.y = c4.y + q_alpha_td * tr;
    c4.z = tr * gl + st;
    c4.w = sp;
    *reinterpret_cast<float4*>(cellr + CC) = c4;
  }
#pragma unroll
  for (int ai = 0; ai < 3; ai++) {
    const float as_i = __shfl_sync(0xffffffffu, as, 16 * c + ai), ae_i = __shfl_sync(0xffffffffu, ae, 16 * c + ai);
    const float tr = v_atr[ai * CC + i];
    v_aw[ai * CC + i] = v_aw[ai * CC + i] + act_alpha_td * tr;
    v_atr[ai * CC + i] = tr * gl + (ae_i - as_i) * st;
  }
  // feed-forward: w[r][j] += (alpha*state_r)*err_j for the cells with state_r > 0
  kf_s[i] = C.a_ff * st;
  __syncwarp();
  const unsigned live = (__ballot_sync(0xffffffffu, st > 0.0f) >> (16 * c)) & 0xffffu;
  if (RH > 0 && st > 0.0f) {  // the register part of this lane's own row
    const float kf = kf_s[i];
    const u64 k2 = pk(kf, kf);
#pragma unroll
    for (int q = 0; q < RQ; q++) {
      const ulonglong2 e = e4[q];
      wr[2 * q] = F.add2(F.mul(k2, e.x), wr[2 * q]);
      wr[2 * q + 1] = F.add2(F.mul(k2, e.y), wr[2 * q + 1]);
    }
  }
  if (st > 0.0f) {  // the shared-memory part of this lane's own row: 128-bit accesses, rows 4 (mod 8) floats apart
    const float kf = kf_s[i];
    const u64 k2 = pk(kf, kf);
    ulonglong2* wrow = reinterpret_cast<ulonglong2*>(W + i * SH);
#pragma unroll 4
    for (int j = 0; j < n4; j++) {
      ulonglong2 wv = wrow[j];
      const ulonglong2 e = e4[RQ + j];
      wv.x = F.add2(F.mul(k2, e.x), wv.x);
      wv.y = F.add2(F.mul(k2, e.y), wv.y);
      wrow[j] = wv;
    }
  }
  // lateral: lat[i][r] = max(0, lat[i][r] + alpha*(state_i*state_r - sparsity^2)), every (i, r): the lane's own row,
  // packed (x - s == x + (-s) exactly)
  {
    const float sp2 = C.sparsity * C.sparsity;
    const u64 st2 = pk(st, st), nsp2 = pk(-sp2, -sp2), al2 = pk(C.a_lat, C.a_lat);
    ulonglong2* lrow = reinterpret_cast<ulonglong2*>(cellr);
    const ulonglong2* s4 = reinterpret_cast<const ulonglong2*>(st_s);
    auto upd = [&](u64 lv, u64 sr) {
      const u64 nv = F.add2(lv, F.mul(al2, F.add2(F.mul(st2, sr), nsp2)));
      float x, y; upk(nv, x, y);
      return pk(bidi_max(0.0f, x), bidi_max(0.0f, y));
    };
#pragma unroll
    for (int q = 0; q < 4; q++) {
      ulonglong2 v = lrow[q];
      const ulonglong2 sr = s4[q];
      v.x = upd(v.x, sr.x); v.y = upd(v.y, sr.y);
      lrow[q] = v;
    }
  }
  if (has) {
    const int cb = node * CC + i;
    B[C.act + cb] = act; B[C.spike + cb] = sp; B[C.state + cb] = st;
  }
  __syncwarp();  // every lane has read prevValue
  if (i == 0) scal[0] = q;

  // ---- record back to HBM: the register part with coalesced 128-bit stores, the rest as one TMA bulk copy
#pragma unroll
  for (int q = 0; q < RQ; q++) reinterpret_cast<ulonglong2*>(g_reg)[q * 32 + lane] = make_ulonglong2(wr[2 * q], wr[2 * q + 1]);
  fence_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_1d(g_rec, rec, (uint32_t)R * 4u);
    tma_store_commit_and_wait();
  }
}

} // namespace bidi
