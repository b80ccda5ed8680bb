// Compile-mode glue.  The per-lane solver logic in rti_core.cuh is written once; it is compiled
//   * by nvcc for sm_100a as the product (the only shipped path), and
//   * by g++ with -DNMPC_HOST_EMUL *only* inside tests/host_emul/ as a lane-by-lane emulation
//     used to debug the kernels in the GPU-less build container.  The emulation is test
//     infrastructure: nothing in the product library or the Python package can reach it.
#pragma once
#include <math.h>
#include "../../include/nmpc_horizon.h"

#if defined(__CUDACC__)
#define NMPC_HD __host__ __device__ __forceinline__
#define NMPC_D __device__ __forceinline__
#else
#define NMPC_HD inline
#define NMPC_D inline
#endif

namespace nmpc {

NMPC_HD void nmpc_sincos(double a, double* s, double* c) {
#if defined(__CUDA_ARCH__)
    sincos(a, s, c);
#else
    *s = sin(a); *c = cos(a);
#endif
}

// per-lane scratch columns live in shared memory on the device, interleaved over the threads of a
// sweep CTA (bank-conflict free); the host emulation uses a plain array
#if defined(__CUDACC__)
#ifndef NMPC_FDF_MINB
#define NMPC_FDF_MINB 16                // resident single-warp CTAs per SM the solve-sweep kernel is compiled for (16: 128 registers)
#endif
#ifndef NMPC_B_MINB
#define NMPC_B_MINB 8                   // resident single-warp CTAs per SM the factorising sweep is compiled for (8: 255 registers)
#endif
#ifndef NMPC_FDF_PREFETCH
#define NMPC_FDF_PREFETCH 0             // 1: the solve sweeps prefetch the next stage's fields into L2 (rti_core.cuh sweep_lane)
#endif
#ifndef NMPC_SW_TILES
#define NMPC_SW_TILES 1                 // tiles (warps) per CTA of a sweep kernel (1: the omni4 factorising sweep fits more warps per SM)
#endif
#ifndef NMPC_B_STAGE_IMAGE
#define NMPC_B_STAGE_IMAGE 1            // factorising sweep of diff / tric: iterate and step rows of the next stage copied asynchronously into shared memory
#endif
#ifndef NMPC_B_L2_PREFETCH
#define NMPC_B_L2_PREFETCH 1            // the factorising sweeps prefetch the rows of the next stage into L2 between their two halves
#endif
#ifndef NMPC_B_KEEP_LIN
#define NMPC_B_KEEP_LIN 1               // factorising sweep of diff / tric: the rows of [A B] stay in registers from the update half to the Riccati half
#endif
#define NMPC_SCRATCH_STRIDE (32 * NMPC_SW_TILES)
#else
#define NMPC_SCRATCH_STRIDE 1
#endif

// keeps the compiler from hoisting the loads of a later phase of a stage above an earlier one
// (register pressure); no instruction is emitted
#if defined(__CUDA_ARCH__)
#define NMPC_PHASE_FENCE() asm volatile("" ::: "memory")
#else
#define NMPC_PHASE_FENCE() ((void)0)
#endif

// Register economy of the factorising per-lane sweep for models with at least this many channels (omni4; rti_core.cuh):
// the update half of a stage is streamed channel by channel, and compiler fences between the columns of the Riccati
// congruence bound how far the loads of later columns are hoisted.  omni4: 3.5 KB -> 0.9 KB of spill stores per thread,
// 6.95 -> 5.75 ms per launch of 65,536 instances; diff / tric are faster in the whole-stage formulation (profiles/README_r02_notes.txt)
#ifndef NMPC_B_LEAN_MINNV
#define NMPC_B_LEAN_MINNV 4
#endif

constexpr int LANES = 32;     // instances per tile = lanes per warp
constexpr int NSTAGE = NMPC_N;    // N and dt: scripts/<m>/common.py:5-9, emitted into include/nmpc_horizon.h by emit.py
constexpr double OCP_DT = NMPC_DT;

// warp-wide "does any lane still want this sweep"; per-lane emulation on the host
#if defined(__CUDA_ARCH__)
#define NMPC_ANY(pred) (__any_sync(0xffffffffu, (pred)))
#else
#define NMPC_ANY(pred) (pred)
#endif

}  // namespace nmpc
