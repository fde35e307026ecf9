// bidi_coop.cuh -- a whole simStep of ONE (or a few) large agents as a single cooperative
// launch.
//
// A single agent's step is a long chain of small dependent phases (C4, iterative coder:
// 445 of them), each a kernel of a few hundred CTAs that runs for microseconds; even
// inside a CUDA graph every link of the chain costs a launch and a drain.  Here one
// persistent kernel, sized to be co-resident (cudaLaunchCooperativeKernel), interprets the
// step as a program: for every phase each CTA works through its share of the phase's
// tiles -- the very tile bodies of bidi_tiles.cuh, so the bits are the same -- and a
// grid-wide barrier replaces the launch boundary.
#pragma once
#include <cooperative_groups.h>
#include "bidi_tiles.cuh"

namespace bidi {

enum CoopCode { OP_KW_ACTIVATE = 1, OP_KW_INHIBIT, OP_RECONSTRUCT, OP_KW_LEARN, OP_IC_SWEEP, OP_IC_LEARN, OP_PREDICT, OP_IC_FINISH, OP_STEP_END };
// arguments after (P, L, l), packed by type in order of appearance: ints -> i[], floats -> f[], long longs -> a[]
struct CoopOp { int code, l, blocks, ni, i[4]; float f[2]; int nf, na; long long a[3]; };

__global__ void __launch_bounds__(TILE_U* TILE_G) k_step_coop(EngineDev P, const CoopOp* __restrict__ ops, int n_ops) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  __shared__ LayerDev sL;
  __shared__ CoopOp sOp;
  for (int k = 0; k < n_ops; k++) {
    __syncthreads();
    for (int w = threadIdx.x; w < (int)(sizeof(CoopOp) / 4); w += blockDim.x) reinterpret_cast<int*>(&sOp)[w] = reinterpret_cast<const int*>(ops + k)[w];
    __syncthreads();
    for (int w = threadIdx.x; w < (int)(sizeof(LayerDev) / 4); w += blockDim.x)
      reinterpret_cast<int*>(&sL)[w] = reinterpret_cast<const int*>(P.layers + sOp.l)[w];
    __syncthreads();
    const CoopOp& o = sOp;
    const LayerDev& L = sL;
    for (int b = blockIdx.x; b < o.blocks; b += gridDim.x) {
      switch (o.code) {
        case OP_KW_ACTIVATE: t_kw_activate(P, L, o.l, b); break;
        case OP_KW_INHIBIT: t_kw_inhibit(P, L, o.l, o.a[0], o.a[1], o.a[2], b); break;
        case OP_RECONSTRUCT: t_reconstruct(P, L, o.l, o.a[0], o.i[0], o.f[0], o.i[1], o.i[2], b); break;
        case OP_KW_LEARN: t_kw_learn(P, L, o.l, b); break;
        case OP_IC_SWEEP: t_ic_sweep(P, L, o.l, o.i[0], o.i[1], o.f[0], o.i[2], b); break;
        case OP_IC_LEARN: t_ic_learn(P, L, o.l, o.i[0], b); break;
        case OP_PREDICT: t_predict(P, L, o.l, o.i[0], o.a[0], o.a[1], b); break;
        case OP_IC_FINISH: d_ic_finish(P, L, o.l, o.i[0], o.f[0], o.a[0], (long long)b * blockDim.x + threadIdx.x); break;
        case OP_STEP_END: d_step_end(P, o.i[0], (long long)b * blockDim.x + threadIdx.x); break;
      }
      __syncthreads();  // the tile's shared-memory scratch is free again
    }
    grid.sync();
  }
}

} // namespace bidi
