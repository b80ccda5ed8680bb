// libnmpc_b200.so — sm_100a kernels and the batched C ABI (include/nmpc_b200.h).
//
// Kernels (one thread = one OCP instance; a warp owns a contiguous "tile" of 32 instances):
//   k_lti_setup  : closed RK4 sensitivities of the lag/integrator rows, per stage (set-up, tiny)
//   k_linearize  : K1 (RK4 + forward sensitivities) + K2 (Gauss-Newton LS gradient), thread per
//                  (instance, stage) -> writes the QP records of the tile workspace
//   k_sweep<KIND>: K3, Riccati-based primal-dual interior point, thread per instance.  One launch =
//                  one horizon sweep (factorise / predictor / delta-backward / delta-forward) of
//                  every instance that is still iterating; the host enqueues the sweeps of all
//                  iterations back to back, finished instances and finished launches fall through
//   k_step       : K4, full SQP-RTI step (+ optional shift), thread per (instance, stage)
// There is no CPU path in this library: without a CUDA device nmpc_create fails.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <new>
#include <vector>

#include "rti_core.cuh"
#include "rti_coop.cuh"
#include "rti_solo.cuh"
#include "ctrl_glue.cuh"
#include "path_disc.cuh"
#include "rollout.cuh"
#include "../../include/nmpc_b200.h"

using namespace nmpc;

static thread_local char g_err[512] = "";
static int set_err(int code, const char* what, cudaError_t e = cudaSuccess)
{
    if (e != cudaSuccess) snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    else snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}
#define CK(call)                                                            \
    do {                                                                    \
        cudaError_t e_ = (call);                                            \
        if (e_ != cudaSuccess) return set_err(NMPC_E_CUDA, #call, e_);      \
    } while (0)

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
template <class M>
__global__ void k_lti_setup(const double* __restrict__ p, double dt, double* __restrict__ lti, double* __restrict__ thr)
{
    using S = Rti<M>;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= NSTAGE) return;
    double x0[S::NX], u0[S::NU], xn[S::NX], Ep[3][S::NC], pk[S::NP], out[4 * S::NV];
    for (int i = 0; i < S::NX; i++) x0[i] = 0.0;
    for (int i = 0; i < S::NU; i++) u0[i] = 0.0;
    for (int i = 0; i < S::NP; i++) pk[i] = p[k * S::NP + i];
    S::rk4_sens(x0, u0, pk, dt, xn, Ep, out);
    for (int i = 0; i < 4 * S::NV; i++) lti[k * 4 * S::NV + i] = out[i];
    for (int c = 0; c < S::NC; c++) thr[k * S::NC + c] = Ep[2][c];      // theta row of the sensitivities (state independent for diff / omni4)
}

// stage table of the group path: row k = [av|ar|au|ru (4 NV), 0, 1, pad | diagonal of W (NY), pad]; row N is zero
__global__ void k_stage_table(int nv, int ny, int t_w, int trow, const double* __restrict__ lti, const double* __restrict__ W,
                              double* __restrict__ stg)
{
    const int k = blockIdx.x, i = threadIdx.x;
    if (i >= trow) return;
    double v = 0.0;
    if (k < NSTAGE) {
        if (i < 4 * nv) v = lti[k * 4 * nv + i];
        else if (i == 4 * nv + 1) v = 1.0;
        else if (i >= t_w && i < t_w + ny) v = W[k * ny + (i - t_w)];
    }
    stg[(size_t)k * trow + i] = v;
}

constexpr int LIN_BLOCK = 128;

// grid: (ceil(nchunk/LIN_BLOCK), N+1).  i0 = first instance of the chunk.
template <class M>
__global__ void __launch_bounds__(LIN_BLOCK)
k_linearize(int B, int i0, int nchunk, const double* __restrict__ x0bar, const double* __restrict__ yref, int nyref,
            const double* __restrict__ We_inst, const double* __restrict__ x, const double* __restrict__ u, int ld,
            Tables tb, double* __restrict__ ws)
{
    using S = Rti<M>;
    using R = typename S::R;
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nchunk) return;
    const int i = i0 + li, k = blockIdx.y;
    double* tl = ws + (size_t)(li / LANES) * R::tile_doubles + (li % LANES);
    double* lin = tl + R::OFF_LIN + (size_t)k * R::NF_LIN * LANES;
    double* it = tl + R::OFF_IT + (size_t)k * R::NF_IT * LANES;
    double xk[S::NX], uk[S::NU], xk1[S::NX], yr[S::NY], xb[S::NX], We[S::NX];
#pragma unroll
    for (int j = 0; j < S::NX; j++) xk[j] = x[((size_t)k * S::NX + j) * ld + i];
    if (k < NSTAGE) {
#pragma unroll
        for (int c = 0; c < S::NU; c++) uk[c] = u[((size_t)k * S::NU + c) * ld + i];
#pragma unroll
        for (int j = 0; j < S::NX; j++) xk1[j] = x[((size_t)(k + 1) * S::NX + j) * ld + i];
    } else {
#pragma unroll
        for (int j = 0; j < S::NX; j++) We[j] = We_inst ? We_inst[(size_t)j * B + i] : tb.We[j];
    }
    for (int j = 0; j < nyref; j++) yr[j] = yref[((size_t)k * nyref + j) * B + i];
    if (k == 0) {
#pragma unroll
        for (int j = 0; j < S::NX; j++) xb[j] = x0bar[(size_t)j * B + i];
    }
    S::linearize_stage(k, xk, uk, xk1, yr, nyref, xb, tb, We, lin, it);
}

// K3 control block of one chunk: per-lane scalars of the interior-point loop (SoA, leading dim ldc)
// and per-iteration counters.  ctl_d rows: nrm[4], mu, alpha, sigmu, mu_aff0, lin_res, mcw;
// ctl_i rows: done, iter, status, fb, nfb.
constexpr int NCTL_D = 10, NCTL_I = 5;
constexpr int SW_TILES = NMPC_SW_TILES;  // tiles (warps) per CTA of a sweep kernel
static_assert(LANES * SW_TILES == NMPC_SCRATCH_STRIDE, "scratch columns are interleaved over the CTA's threads");

template <class S>
__device__ __forceinline__ void ctl_load(typename S::LaneCtl& c, const double* cd, const int* ci, int ldc, int li)
{
#pragma unroll
    for (int q = 0; q < 4; q++) c.nrm[q] = cd[(size_t)q * ldc + li];
    c.mu = cd[(size_t)4 * ldc + li]; c.alpha = cd[(size_t)5 * ldc + li]; c.sigmu = cd[(size_t)6 * ldc + li];
    c.mu_aff0 = cd[(size_t)7 * ldc + li]; c.lin_res = cd[(size_t)8 * ldc + li]; c.mcw = cd[(size_t)9 * ldc + li];
    c.done = ci[(size_t)0 * ldc + li]; c.iter = ci[(size_t)1 * ldc + li]; c.status = ci[(size_t)2 * ldc + li];
    c.fb = ci[(size_t)3 * ldc + li]; c.nfb = ci[(size_t)4 * ldc + li];
}
template <class S>
__device__ __forceinline__ void ctl_store(const typename S::LaneCtl& c, double* cd, int* ci, int ldc, int li)
{
#pragma unroll
    for (int q = 0; q < 4; q++) cd[(size_t)q * ldc + li] = c.nrm[q];
    cd[(size_t)4 * ldc + li] = c.mu; cd[(size_t)5 * ldc + li] = c.alpha; cd[(size_t)6 * ldc + li] = c.sigmu;
    cd[(size_t)7 * ldc + li] = c.mu_aff0; cd[(size_t)8 * ldc + li] = c.lin_res; cd[(size_t)9 * ldc + li] = c.mcw;
    ci[(size_t)0 * ldc + li] = c.done; ci[(size_t)1 * ldc + li] = c.iter; ci[(size_t)2 * ldc + li] = c.status;
    ci[(size_t)3 * ldc + li] = c.fb; ci[(size_t)4 * ldc + li] = c.nfb;
}

// K3, one kernel = one or three horizon sweeps for every lane of the chunk that is still iterating.
//   KIND: Rti::SW_B_FIRST (cold start + residuals + factorise), SW_FDF (predictor forward, delta
//         backward, delta forward [+ centering repeat]), SW_B (apply step + residuals + factorise).
//   gate: number of lanes still iterating (act[it]); 0 -> the whole grid returns at once.
//   cnt_out (B kernels): lanes that continue -> act[it+1].
template <class M, int KIND>
__global__ void __launch_bounds__(LANES * SW_TILES, (KIND == Rti<M>::SW_FDF ? NMPC_FDF_MINB : NMPC_B_MINB) / SW_TILES)
k_sweep(int B, int i0, int nchunk, int ldc, Tables tb, const double* __restrict__ We_inst, IpmOpts o, double* __restrict__ ws,
        double* __restrict__ ctl_d, int* __restrict__ ctl_i, const int* __restrict__ gate, int* __restrict__ cnt_out, int gate_min,
        int defer_fb)
{
    using S = Rti<M>;
    using R = typename S::R;
    if (KIND != S::SW_B_FIRST && *gate < gate_min) return;     // gate_min = 1: run while any lane iterates
    extern __shared__ double scratch_buf[];       // B sweeps: CarryB::SC_N columns of NMPC_SCRATCH_STRIDE doubles
    double* scratch = scratch_buf;
    const int li = blockIdx.x * (LANES * SW_TILES) + threadIdx.x;
    const bool active = li < nchunk;
    typename S::LaneCtl c;
    bool run = false;
    if (KIND == S::SW_B_FIRST) { c.init(active); run = active; }
    else if (active) {
        run = ctl_i[li] == 0;                      // row 0 = done
        if (run) ctl_load<S>(c, ctl_d, ctl_i, ldc, li);
    }
    if (run) {
        double We[S::NX];
        if (KIND == S::SW_B_FIRST || KIND == S::SW_B) {
#pragma unroll
            for (int j = 0; j < S::NX; j++) We[j] = We_inst ? We_inst[(size_t)j * B + i0 + li] : tb.We[j];
        }
        double* tile_lane = ws + (size_t)(li / LANES) * R::tile_doubles + (li % LANES);
        S::template run_phase<KIND>(tile_lane, tb, We, o, KIND == S::SW_FDF && defer_fb != 0, c, scratch + threadIdx.x);
    }
    if (active && (run || KIND == S::SW_B_FIRST)) ctl_store<S>(c, ctl_d, ctl_i, ldc, li);
    if (KIND == S::SW_B_FIRST || KIND == S::SW_B) {
        const unsigned m = __ballot_sync(0xffffffffu, run && c.done == 0);
        if ((threadIdx.x & (LANES - 1)) == 0 && m) atomicAdd(cnt_out, __popc(m));
    }
}

// ---- persistent path (rti_coop.cuh, rti_records.cuh): per-instance contiguous records -----------
// K1+K2 and the interior-point cold start into the group layout: block = LING_BLOCK instances of one
// stage; the [Q, LHD) head and the [MC, NREC) tail of every record are staged in shared memory and
// written out as contiguous runs (368 + 496 bytes for diff).
constexpr int LING_BLOCK = 64;
template <class M>
__global__ void __launch_bounds__(LING_BLOCK)
k_linearize_g(int B, int i0, int nchunk, const double* __restrict__ x0bar, const double* __restrict__ yref, int nyref,
              const double* __restrict__ We_inst, const double* __restrict__ x, const double* __restrict__ u, int ld,
              Tables tb, IpmOpts o, double* __restrict__ ws, const int* __restrict__ active)
{
    using S = Rti<M>;
    using GR = GRec<S::NV>;
    constexpr int HEAD = GR::LHD, TAIL = GR::NREC - GR::MC, ROW = (HEAD + TAIL) | 1;   // odd row stride: conflict-free
    extern __shared__ double lin_sm[];
    const int li = blockIdx.x * LING_BLOCK + threadIdx.x, k = blockIdx.y;
    const bool on = li < nchunk && (!active || active[i0 + li] != 0);      // SQP passes: instances that stopped iterating keep their records
    if (on) {
        const int i = i0 + li;
        double xk[S::NX], uk[S::NU], xk1[S::NX], yr[S::NY], xb[S::NX], We[S::NX];
#pragma unroll
        for (int j = 0; j < S::NX; j++) xk[j] = x[((size_t)k * S::NX + j) * ld + i];
        if (k < NSTAGE) {
#pragma unroll
            for (int c = 0; c < S::NU; c++) uk[c] = u[((size_t)k * S::NU + c) * ld + i];
#pragma unroll
            for (int j = 0; j < S::NX; j++) xk1[j] = x[((size_t)(k + 1) * S::NX + j) * ld + i];
        } else {
#pragma unroll
            for (int j = 0; j < S::NX; j++) We[j] = We_inst ? We_inst[(size_t)j * B + i] : tb.We[j];
        }
        for (int j = 0; j < nyref; j++) yr[j] = yref[((size_t)k * nyref + j) * B + i];
        if (k == 0) {
#pragma unroll
            for (int j = 0; j < S::NX; j++) xb[j] = x0bar[(size_t)j * B + i];
        }
        double* row = lin_sm + (size_t)threadIdx.x * ROW;
        for (int d = 0; d < HEAD + TAIL; d++) row[d] = 0.0;     // stages 0 and N fill only part; padding stays defined
        double* tail = row + HEAD - GR::MC;                     // record offsets >= MC land in the tail part of the row
        S::template linearize_stage<GR, 1>(k, xk, uk, xk1, yr, nyref, xb, tb, We, row, tail);
        S::template coldstart_stage<GR, 1>(k, o, row, tail);
    }
    __syncthreads();
    const int nrow = min(LING_BLOCK, nchunk - blockIdx.x * LING_BLOCK);
    for (int idx = threadIdx.x; idx < nrow * (HEAD + TAIL); idx += LING_BLOCK) {
        const int r = idx / (HEAD + TAIL), d = idx - r * (HEAD + TAIL);
        const int off = d < HEAD ? d : GR::MC + (d - HEAD);
        if (active && active[i0 + blockIdx.x * LING_BLOCK + r] == 0) continue;
        ws[(size_t)(blockIdx.x * LING_BLOCK + r) * GR::inst_doubles + (size_t)k * GR::NREC + off] = lin_sm[(size_t)r * ROW + d];
    }
}

#ifndef NMPC_GRP_WARPS
#define NMPC_GRP_WARPS 4
#endif
constexpr int GRP_WARPS = NMPC_GRP_WARPS;           // warps per CTA of the persistent K3 kernel

// K3, persistent lane-cooperative kernel (rti_coop.cuh): each warp runs the whole interior-point loop of 32/G instances
// at a time and refills converged slots from the queue *next (instances [0, n) of the chunk, or the hand-over list)
#ifndef NMPC_COOP_MINB
#define NMPC_COOP_MINB 3
#endif
template <class M, int G, int MINB>
__global__ void __launch_bounds__(GRP_WARPS * 32, MINB)
k_ipm_coop(int i0, int n, Tables tb, const double* __restrict__ We_inst, int ldWe, IpmOpts o, double* __restrict__ ws,
           int* __restrict__ next, GrpOut out, GrpResume rs)
{
    using GP = Coop<M, G>;
    extern __shared__ __align__(16) double grp_sm[];
    typename GP::Lane L;
    GP::init_lane(L, threadIdx.x & 31, threadIdx.x >> 5);
    GP::run_warp(&L, grp_sm, ws, i0, n, next, tb, We_inst, ldWe, o, out, rs);
}

// K3, block-per-instance mapping (rti_solo.cuh): the interior point of ONE instance in the shared memory of one SM - the
// latency path (the ROS drop-in at batch 1, small fleets).  Reads the QP from the tile layout K1 / K2 wrote, leaves the QP
// solution in IT.Z of the tile for K4.
template <class M>
__global__ void __launch_bounds__(Solo<M>::THREADS, 1)
k_ipm_solo(int B, int i0, Tables tb, const double* __restrict__ We_inst, IpmOpts o, double* __restrict__ ws,
           int* __restrict__ qp_status, int* __restrict__ qp_iter, double* __restrict__ stats, double* __restrict__ scr)
{
    using S = Rti<M>;
    using R = typename S::R;
    static_assert(Solo<M>::THREADS > NSTAGE, "one thread per stage");
    static_assert(Solo<M>::SM_BYTES <= 227 * 1024, "the interior-point state of one instance must fit the shared memory of an SM");
    extern __shared__ __align__(16) double solo_sm[];
    const int li = blockIdx.x, i = i0 + li;
    double* tile_lane = ws + (size_t)(li / LANES) * R::tile_doubles + (li % LANES);
    const double* We = We_inst ? We_inst + i : tb.We;
    typename S::LaneStats st;
    Solo<M>::run(solo_sm, typename Solo<M>::TileIO{tile_lane}, tb, We, We_inst ? (size_t)B : 1, o, &st,
                 scr ? scr + (size_t)blockIdx.x * Solo<M>::GSCR_DOUBLES : nullptr);
    if (threadIdx.x == 0) {
        qp_status[i] = st.status; qp_iter[i] = st.iter;
        if (stats) {
#pragma unroll
            for (int q = 0; q < 4; q++) stats[(size_t)q * B + i] = st.res[q];
            stats[(size_t)4 * B + i] = st.mu; stats[(size_t)5 * B + i] = st.lin_res;
            stats[(size_t)6 * B + i] = (double)st.cond_fallbacks; stats[(size_t)7 * B + i] = (double)st.status;
        }
    }
}

// the same on the per-instance records of the lane-cooperative path (k_linearize_g / k_step_g around it): block b takes entry b
// of the queue (all n instances of the chunk, or the compacted list of the instances an SQP pass still iterates)
template <class M>
__global__ void __launch_bounds__(Solo<M>::THREADS, 1)
k_ipm_solo_g(int i0, int n, Tables tb, const double* __restrict__ We_inst, int ldWe, IpmOpts o, double* __restrict__ ws,
             GrpOut out, GrpResume rs, double* __restrict__ scr)
{
    using S = Rti<M>;
    using GR = GRec<S::NV>;
    extern __shared__ __align__(16) double solo_sm[];
    const int cnt = rs.n_dev ? *rs.n_dev : n;
    if ((int)blockIdx.x >= cnt) return;
    const int gi = rs.list ? rs.list[blockIdx.x] : (int)blockIdx.x;
    const int i = i0 + gi;
    const double* We = We_inst ? We_inst + i : tb.We;
    typename S::LaneStats st;
    Solo<M>::run(solo_sm, typename Solo<M>::RecIO{ws + (size_t)gi * GR::inst_doubles}, tb, We, We_inst ? (size_t)ldWe : 1, o, &st,
                 scr ? scr + (size_t)blockIdx.x * Solo<M>::GSCR_DOUBLES : nullptr);
    if (threadIdx.x == 0) {
        out.qp_status[i] = st.status; out.qp_iter[i] = st.iter;
        if (out.stats) {
#pragma unroll
            for (int q = 0; q < 4; q++) out.stats[(size_t)q * out.B + i] = st.res[q];
            out.stats[(size_t)4 * out.B + i] = st.mu; out.stats[(size_t)5 * out.B + i] = st.lin_res;
            out.stats[(size_t)6 * out.B + i] = (double)st.cond_fallbacks; out.stats[(size_t)7 * out.B + i] = (double)st.status;
        }
    }
}

// K4 from the group layout
template <class M>
__global__ void __launch_bounds__(LIN_BLOCK)
k_step_g(int B, int i0, int nchunk, const double* __restrict__ x0bar, double* __restrict__ x, double* __restrict__ u, int ld,
         const double* __restrict__ ws, const int* __restrict__ qp_status, int* __restrict__ status,
         const int* __restrict__ active, unsigned long long* __restrict__ stepn)
{
    using S = Rti<M>;
    using GR = GRec<S::NV>;
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nchunk) return;
    const int i = i0 + li, k = blockIdx.y;
    if (active && active[i] == 0) return;
    const int qs = qp_status[i];
    if (qs != 0 && qs != 1) { if (k == 0) status[i] = NMPC_QP_FAILURE; return; }
    const double* rec = ws + (size_t)li * GR::inst_doubles + (size_t)k * GR::NREC;
    double xk[S::NX], uk[S::NU], xb[S::NX];
#pragma unroll
    for (int j = 0; j < S::NX; j++) { xk[j] = x[((size_t)k * S::NX + j) * ld + i]; xb[j] = (k == 0) ? x0bar[(size_t)j * B + i] : 0.0; }
    if (k < NSTAGE) {
#pragma unroll
        for (int c = 0; c < S::NU; c++) uk[c] = u[((size_t)k * S::NU + c) * ld + i];
    }
    if (stepn) {
        // inf-norm of this stage's step (SQP convergence test): x_0 <- x0bar counts with its distance to the old x_0
        double xo[S::NX], uo[S::NU];
#pragma unroll
        for (int j = 0; j < S::NX; j++) xo[j] = xk[j];
#pragma unroll
        for (int c = 0; c < S::NU; c++) uo[c] = k < NSTAGE ? uk[c] : 0.0;
        S::template step_stage<GR, 1>(k, rec, xb, xk, uk);
        double nrm = 0.0;
#pragma unroll
        for (int j = 0; j < S::NX; j++) nrm = fmax(nrm, fabs(xk[j] - xo[j]));
        if (k < NSTAGE) {
#pragma unroll
            for (int c = 0; c < S::NU; c++) nrm = fmax(nrm, fabs(uk[c] - uo[c]));
        }
        if (!(nrm == nrm)) nrm = 1e300;                                  // NaN: never "converged"
        atomicMax(&stepn[i], (unsigned long long)__double_as_longlong(nrm));   // non-negative doubles order like their bit patterns
    } else S::template step_stage<GR, 1>(k, rec, xb, xk, uk);
    bool bad = false;
#pragma unroll
    for (int j = 0; j < S::NX; j++) { x[((size_t)k * S::NX + j) * ld + i] = xk[j]; bad |= (xk[j] != xk[j]); }
    if (k < NSTAGE) {
#pragma unroll
        for (int c = 0; c < S::NU; c++) u[((size_t)k * S::NU + c) * ld + i] = uk[c];
    }
    if (bad) atomicMax(&status[i], NMPC_NAN_DETECTED);
}

// ---- BASELINE config 4 (north star kernel (4)): SQP bookkeeping and the warm-start shift ------------------------------
// per-instance state of an SQP solve: active (still iterating), RTI steps taken, QP iterations summed, step inf-norm
__global__ void k_sqp_begin(int B, int* __restrict__ active, int* __restrict__ sqp_iter, int* __restrict__ qp_total,
                            unsigned long long* __restrict__ stepn)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    active[i] = 1; sqp_iter[i] = 0; qp_total[i] = 0; stepn[i] = 0ull;
}
// queue of one chunk's active instances (order = arrival; an instance's result does not depend on its slot)
__global__ void k_sqp_list(int i0, int nchunk, const int* __restrict__ active, int* __restrict__ count, int* __restrict__ list)
{
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = li < nchunk && active[i0 + li] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (on) list[base + __popc(m & ((1u << lane) - 1u))] = li;
}
// after an RTI pass: count it, stop the instances that converged (step <= tol), failed or reached max_iter
__global__ void k_sqp_update(int B, int max_iter, double tol, int* __restrict__ active, const int* __restrict__ status,
                             const int* __restrict__ qp_iter, int* __restrict__ sqp_iter, int* __restrict__ qp_total,
                             unsigned long long* __restrict__ stepn)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B || active[i] == 0) return;
    const int it = sqp_iter[i] + 1;
    sqp_iter[i] = it; qp_total[i] += qp_iter[i];
    const double nrm = __longlong_as_double((long long)stepn[i]);
    stepn[i] = 0ull;
    if (status[i] != 0 || nrm <= tol || it >= max_iter) active[i] = 0;
}
// warm-start shift (scripts/test_scripts/casadi_sim_diff.py:104-106 warm-starts from the previous
// solution): row r of the stage-major iterate takes the value one stage later, the last stage is kept.
// grid: (ceil(B / 128), rows) with rows = nx for x (nstages = N + 1) or nu for u (nstages = N)
__global__ void k_shift(int B, int ld, int rows, int nstages, double* __restrict__ a, const int* __restrict__ mask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= B || (mask && mask[i] == 0)) return;
    double nxt = a[((size_t)1 * rows + r) * ld + i];
    for (int k = 0; k + 1 < nstages; k++) {
        a[((size_t)k * rows + r) * ld + i] = nxt;
        if (k + 2 < nstages) nxt = a[((size_t)(k + 2) * rows + r) * ld + i];
    }
}

// ---- hybrid schedule: hand-over of the instances the per-sweep kernels did not finish ----------------------
// compaction in two passes: unfinished instance li -> record q = map[li] of the group workspace, list[q] = li, its
// control block.  Records are ordered by decreasing complementarity gap mu (HB_NB buckets of two binary orders
// of magnitude): log10(mu) at the hand-over predicts the remaining iterations (correlation 0.9 on the synthetic
// inputs), and the queue of the group kernel is consumed front to back, so the instances that will iterate
// longest start first and the few stragglers do not run alone at the end (longest-processing-time-first).
constexpr int HB_NB = 32;
__device__ __forceinline__ int handover_bucket(double mu)
{
    if (!(mu > 0.0)) return 0;                              // NaN / non-positive: treat as "far from converged"
    int e; frexp(mu, &e);                                   // mu = f * 2^e, f in [0.5, 1)
    int b = (4 - e) / 2;                                    // mu >= 8 -> 0, then one bucket per factor 4
    return b < 0 ? 0 : (b >= HB_NB ? HB_NB - 1 : b);
}
// pass 1: bucket counts; map[li] = bucket or -1
__global__ void k_handover_count(int nchunk, int ldc, const double* __restrict__ ctl_d, const int* __restrict__ ctl_i,
                                 int* __restrict__ bcnt, int* __restrict__ map)
{
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nchunk) return;
    int b = -1;
    if (ctl_i[li] != 1) {                                   // row 0 = done: 0 iterating, 2 left the lockstep path at a centering repeat
        b = handover_bucket(ctl_d[(size_t)4 * ldc + li]);   // row 4 = mu
        atomicAdd(&bcnt[b], 1);
    }
    map[li] = b;
}
// pass 2: position inside the bucket, list / map / control block; bcnt[HB_NB..2*HB_NB) = fill counters
template <class M>
__global__ void k_handover_assign(int nchunk, int ldc, const double* __restrict__ ctl_d, const int* __restrict__ ctl_i,
                                  int* __restrict__ bcnt, int* __restrict__ nres, int* __restrict__ list, int* __restrict__ map,
                                  void* __restrict__ ctl_out)
{
    using S = Rti<M>;
    static_assert(sizeof(typename S::LaneCtl) <= 128, "d_ctl_g is sized 128 bytes per instance");
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li == 0) { int t = 0; for (int b = 0; b < HB_NB; b++) t += bcnt[b]; *nres = t; }
    if (li >= nchunk) return;
    const int b = map[li];
    if (b < 0) return;
    int off = 0;
    for (int bb = 0; bb < b; bb++) off += bcnt[bb];
    const int q = off + atomicAdd(&bcnt[HB_NB + b], 1);
    list[q] = li;
    map[li] = q;
    typename S::LaneCtl c;
    ctl_load<S>(c, ctl_d, ctl_i, ldc, li);
    c.done = 0;
    reinterpret_cast<typename S::LaneCtl*>(ctl_out)[q] = c;
}
// state of the unfinished instances from the tile layout into their group records; grid (instances / CV_BLOCK, stages).
// A thread gathers its instance's fields (coalesced over the lanes of a tile) into a shared-memory row, the block then
// writes the rows out coalesced, a quarter of the record at a time (19 KB of shared memory per block instead of 71 KB:
// four times the resident threads).  Measured on the diff headline batch (ncu, 65,536 instances, 26 %
// handed over): a thread storing its record 8 bytes at a time 2.77 ms; staged through shared memory (this) 2.46 ms;
// staged, but a thread per UNFINISHED instance in list order (all threads busy, tile reads scattered) 3.17 ms; four
// field slices per instance with one block per tile 3.0 ms.  The kernel reads 4.3 GB - nearly the whole tile workspace,
// since a 32-byte sector holds four lanes and 26 % of the lanes are wanted.
constexpr int CV_BLOCK = 64, CV_PIECES = 4;
template <class M>
__global__ void __launch_bounds__(CV_BLOCK)
k_handover_convert(int nchunk, const int* __restrict__ map, const double* __restrict__ ws_tile, double* __restrict__ ws_grp,
                   const double* __restrict__ thr)
{
    using GP = RecOps<M>;
    using R = typename Rti<M>::R;
    using GR = typename GP::R;
    constexpr int PW = ((GR::NREC + CV_PIECES - 1) / CV_PIECES + 1) & ~1;     // piece width, even
    constexpr int ROW = PW | 1;                                              // odd row stride: conflict-free
    __shared__ double cv_sm[CV_BLOCK * ROW];
    __shared__ int cv_q[CV_BLOCK];
    const int li = blockIdx.x * CV_BLOCK + threadIdx.x, k = blockIdx.y;
    int q = -1;
    if (li < nchunk) q = map[li];
    cv_q[threadIdx.x] = q;
    const double* tl = ws_tile + (size_t)(li / LANES) * R::tile_doubles + (li % LANES);
    const double* thr_k = thr + (size_t)(k < NSTAGE ? k : 0) * Rti<M>::NC;
    auto piece = [&](auto pc) {
        constexpr int D0 = decltype(pc)::value * PW, D1 = D0 + PW < GR::NREC ? D0 + PW : GR::NREC, W = D1 - D0;
        if (q >= 0) GP::template tile_to_record_range<D0, D1>(tl, k, cv_sm + threadIdx.x * ROW, thr_k);
        __syncthreads();
        for (int idx = threadIdx.x; idx < CV_BLOCK * W; idx += CV_BLOCK) {
            const int r = idx / W, d = idx - r * W;
            const int qr = cv_q[r];
            if (qr >= 0) GP::rec_of(ws_grp, qr, k)[D0 + d] = cv_sm[r * ROW + d];
        }
        __syncthreads();
    };
    static_assert(CV_PIECES == 4, "one call per piece below");
    piece(std::integral_constant<int, 0>{}); piece(std::integral_constant<int, 1>{});
    piece(std::integral_constant<int, 2>{}); piece(std::integral_constant<int, 3>{});
}
// K4 for the hybrid schedule: the step of an instance comes from its tile or, if it was handed over, from its group record
template <class M>
__global__ void __launch_bounds__(LIN_BLOCK)
k_step_h(int B, int i0, int nchunk, const double* __restrict__ x0bar, double* __restrict__ x, double* __restrict__ u, int ld,
         const double* __restrict__ ws_tile, const double* __restrict__ ws_grp, const int* __restrict__ map,
         const int* __restrict__ qp_status, int* __restrict__ status)
{
    using S = Rti<M>;
    using R = typename S::R;
    using GR = GRec<S::NV>;
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nchunk) return;
    const int i = i0 + li, k = blockIdx.y;
    const int qs = qp_status[i];
    if (qs != 0 && qs != 1) { if (k == 0) status[i] = NMPC_QP_FAILURE; return; }
    double xk[S::NX], uk[S::NU], xb[S::NX];
#pragma unroll
    for (int j = 0; j < S::NX; j++) { xk[j] = x[((size_t)k * S::NX + j) * ld + i]; xb[j] = (k == 0) ? x0bar[(size_t)j * B + i] : 0.0; }
    if (k < NSTAGE) {
#pragma unroll
        for (int c = 0; c < S::NU; c++) uk[c] = u[((size_t)k * S::NU + c) * ld + i];
    }
    const int q = map[li];
    if (q >= 0) S::template step_stage<GR, 1>(k, ws_grp + (size_t)q * GR::inst_doubles + (size_t)k * GR::NREC, xb, xk, uk);
    else S::step_stage(k, ws_tile + (size_t)(li / LANES) * R::tile_doubles + R::OFF_IT + (size_t)k * R::NF_IT * LANES + (li % LANES), xb, xk, uk);
    bool bad = false;
#pragma unroll
    for (int j = 0; j < S::NX; j++) { x[((size_t)k * S::NX + j) * ld + i] = xk[j]; bad |= (xk[j] != xk[j]); }
    if (k < NSTAGE) {
#pragma unroll
        for (int c = 0; c < S::NU; c++) u[((size_t)k * S::NU + c) * ld + i] = uk[c];
    }
    if (bad) atomicMax(&status[i], NMPC_NAN_DETECTED);
}

// end of K3: per-instance QP status / iteration count / statistics out of the control block
__global__ void k_ipm_finish(int B, int i0, int nchunk, int ldc, const double* __restrict__ ctl_d, const int* __restrict__ ctl_i,
                             int* __restrict__ qp_status, int* __restrict__ qp_iter, double* __restrict__ stats)
{
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nchunk) return;
    const int i = i0 + li;
    const int status = ctl_i[(size_t)2 * ldc + li];
    qp_status[i] = status;
    qp_iter[i] = ctl_i[(size_t)1 * ldc + li];
    if (stats) {
#pragma unroll
        for (int q = 0; q < 4; q++) stats[(size_t)q * B + i] = ctl_d[(size_t)q * ldc + li];
        stats[(size_t)4 * B + i] = ctl_d[(size_t)4 * ldc + li];
        stats[(size_t)5 * B + i] = ctl_d[(size_t)8 * ldc + li];
        stats[(size_t)6 * B + i] = (double)ctl_i[(size_t)4 * ldc + li];
        stats[(size_t)7 * B + i] = (double)status;
    }
}

// K4: x += dx, u += du (acados ocp_nlp_update_variables_sqp with alpha = 1).
// A QP that failed other than by max-iter leaves the iterate untouched and yields status 4.
template <class M>
__global__ void __launch_bounds__(LIN_BLOCK)
k_step(int B, int i0, int nchunk, const double* __restrict__ x0bar, double* __restrict__ x, double* __restrict__ u, int ld,
       const double* __restrict__ ws, const int* __restrict__ qp_status, int* __restrict__ status)
{
    using S = Rti<M>;
    using R = typename S::R;
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nchunk) return;
    const int i = i0 + li, k = blockIdx.y;
    const int qs = qp_status[i];
    if (qs != 0 && qs != 1) { if (k == 0) status[i] = NMPC_QP_FAILURE; return; }
    const double* rec = ws + (size_t)(li / LANES) * R::tile_doubles + R::OFF_IT + (size_t)k * R::NF_IT * LANES + (li % LANES);
    double xk[S::NX], uk[S::NU], xb[S::NX];
#pragma unroll
    for (int j = 0; j < S::NX; j++) { xk[j] = x[((size_t)k * S::NX + j) * ld + i]; xb[j] = (k == 0) ? x0bar[(size_t)j * B + i] : 0.0; }
    if (k < NSTAGE) {
#pragma unroll
        for (int c = 0; c < S::NU; c++) uk[c] = u[((size_t)k * S::NU + c) * ld + i];
    }
    S::step_stage(k, rec, xb, xk, uk);
    bool bad = false;
#pragma unroll
    for (int j = 0; j < S::NX; j++) { x[((size_t)k * S::NX + j) * ld + i] = xk[j]; bad |= (xk[j] != xk[j]); }
    if (k < NSTAGE) {
#pragma unroll
        for (int c = 0; c < S::NU; c++) u[((size_t)k * S::NU + c) * ld + i] = uk[c];
    }
    if (bad) atomicMax(&status[i], NMPC_NAN_DETECTED);
}

__global__ void k_fill_int(int n, int* a, int v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

// instance-major [B][R] -> SoA [R][B] (and back), 32x32 tiles through shared memory
__global__ void k_aos_to_soa(int B, int R, const double* __restrict__ in, double* __restrict__ out)
{
    __shared__ double t[32][33];
    const int b0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int b = b0 + y, r = r0 + threadIdx.x;
        if (b < B && r < R) t[y][threadIdx.x] = in[(size_t)b * R + r];
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int r = r0 + y, b = b0 + threadIdx.x;
        if (b < B && r < R) out[(size_t)r * B + b] = t[threadIdx.x][y];
    }
}
// SoA with leading dim ld [R][ld] -> instance-major [B][R]
__global__ void k_soa_to_aos(int B, int R, int ld, const double* __restrict__ in, double* __restrict__ out)
{
    __shared__ double t[32][33];
    const int b0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int r = r0 + y, b = b0 + threadIdx.x;
        if (b < B && r < R) t[y][threadIdx.x] = in[(size_t)r * ld + b];
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int b = b0 + y, r = r0 + threadIdx.x;
        if (b < B && r < R) out[(size_t)b * R + r] = t[threadIdx.x][y];
    }
}
// instance-major [B][R] -> SoA with leading dim ld
__global__ void k_aos_to_soa_ld(int B, int R, int ld, const double* __restrict__ in, double* __restrict__ out)
{
    __shared__ double t[32][33];
    const int b0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int b = b0 + y, r = r0 + threadIdx.x;
        if (b < B && r < R) t[y][threadIdx.x] = in[(size_t)b * R + r];
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int r = r0 + y, b = b0 + threadIdx.x;
        if (b < B && r < R) out[(size_t)r * ld + b] = t[threadIdx.x][y];
    }
}
// default iterate of a freshly created solver: x_k = x0_default for all k, u = 0
__global__ void k_init_iterate(int cap, int nx, int nu, const double* __restrict__ x0def, double* x, double* u)
{
    const size_t n = (size_t)cap * (NSTAGE + 1) * nx;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x)
        x[t] = x0def[(t / cap) % nx];
    const size_t m = (size_t)cap * NSTAGE * nu;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < m; t += (size_t)gridDim.x * blockDim.x) u[t] = 0.0;
}

// fp64 FMA peak: 8 independent chains per thread
__global__ void k_dfma(int iters, double* out)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-6;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct ModelInfo {
    int nx, nu, np, nv;
    double p[3], Q[11], Rw[4], QN[11], lbx[4], ubx[4], lbu[4], ubu[4];
};
// config/nmpc_nav_control_acados_models.yaml:2-75 through scripts/<m>/generate_c_code.py:30-60; the table is generated by
// `python -m nmpc_nav_control_b200.emit <yaml>`, the committed copy is the reference's YAML
#ifndef NMPC_MODEL_DEFAULTS_INC
#define NMPC_MODEL_DEFAULTS_INC "model_defaults.inc"
#endif
static const ModelInfo g_models[3] = {
#include NMPC_MODEL_DEFAULTS_INC
};

// ---- plant step and nearest path parameter, one thread per robot (rollout.cuh) ---
template <class M>
__global__ void k_plant_step(int B, double* __restrict__ xp, const double* __restrict__ u, int ldu, const double* __restrict__ noise,
                             const double* __restrict__ p, double dt, double* __restrict__ pose, double* __restrict__ vel,
                             double* __restrict__ steer)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    Rollout<M>::plant_step(xp + i, (size_t)B, u + i, (size_t)ldu, noise ? noise + i : nullptr, p, dt, pose + i, vel + i,
                           steer ? steer + i : nullptr);
}
__global__ void k_path_nearest(int B, const double* __restrict__ segs, const int* __restrict__ path_off, int n_paths,
                               const int* __restrict__ path_id, const double* __restrict__ pose, double back, double ahead,
                               double* __restrict__ u)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    int p = path_id ? path_id[i] : 0;
    p = p < 0 ? 0 : (p >= n_paths ? n_paths - 1 : p);
    const int s0 = path_off[p], s1 = path_off[p + 1];
    u[i] = PathNearest::nearest_u(segs + (size_t)s0 * PathDisc::SEG, s1 - s0, u[i], pose[i], pose[(size_t)B + i], back, ahead);
}

// ---- batched path discretisation, one thread per robot (path_disc.cuh) ----------
__global__ void k_path_discretize(int B, const double* __restrict__ segs, const int* __restrict__ path_off, int n_paths,
                                  const int* __restrict__ path_id, const double* __restrict__ u0, double period, int num_poses,
                                  int holonomic, double* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    int p = path_id ? path_id[i] : 0;
    p = p < 0 ? 0 : (p >= n_paths ? n_paths - 1 : p);
    const int s0 = path_off[p], s1 = path_off[p + 1];
    PathDisc::next_poses(segs + (size_t)s0 * PathDisc::SEG, s1 - s0, u0[i], period, num_poses, holonomic != 0, out + i, (size_t)B);
}

// ---- controller glue around the solve, one thread per instance (ctrl_glue.cuh) ----
template <class M>
__global__ void k_ctrl_pre(int B, const double* __restrict__ pose, const double* __restrict__ vel, const double* __restrict__ steer,
                           const double* __restrict__ refs, const int* __restrict__ nref, int nref_max, const double* __restrict__ vref,
                           int ldv, const double* __restrict__ p, const double* __restrict__ W0, const double* __restrict__ We_tab,
                           double* __restrict__ x0bar, double* __restrict__ yref, double* __restrict__ We)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    int n = nref ? nref[i] : nref_max;
    if (n > nref_max) n = nref_max;
    CtrlGlue<M>::pre(pose + i, vel + i, steer ? steer + i : nullptr, refs + i, n, vref + i, (size_t)ldv, p, W0, We_tab,
                     x0bar + i, yref + i, We ? We + i : nullptr, (size_t)B);
}
template <class M>
__global__ void k_ctrl_post(int B, const int* __restrict__ status, const double* __restrict__ x0bar, const double* __restrict__ u,
                            int ldu, double dt, const double* __restrict__ p, double* __restrict__ vref, int ldv,
                            double* __restrict__ cmd)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    CtrlGlue<M>::post(status[i], x0bar + i, (size_t)B, u + i, (size_t)ldu, dt, p, vref + i, (size_t)ldv, cmd + i);
}

struct nmpc_solver {
    int model, cap, device, chunk;
    int k3_group = 2;            // K3 schedule: 0 per-sweep kernels (rti_core.cuh), 1 persistent lane-cooperative kernel (rti_coop.cuh),
                                 // 2 hybrid: per-sweep kernels while most instances iterate, then the group kernel for the rest
    double *d_ws_g = nullptr;    // group workspace (schedules 1, 2)
    double *d_solo_scr = nullptr; // block-per-instance kernel, models whose state exceeds the shared memory of an SM (omni4): [B A], Phi, Phi' per instance
    int *d_list = nullptr, *d_map = nullptr;
    void* d_ctl_g = nullptr;
    int hyb_kmax = 12; double hyb_frac = 0.75;  // hand over once fewer than 75 % of the chunk iterate: a lockstep launch costs the same at any
                                 // active fraction (17.5 k instance-iterations / ms when full), the lane-cooperative kernel sustains 12.7 k
    int hyb_min = 24576;         // smaller batches go to the group kernel alone (its per-iteration latency is 3-4x lower)
    int grp_blocks = 0;
    // function attributes (dynamic shared memory opt-in) are per device: remembered per solver, not per process
    int grp_blocks_per_sm = 0; bool sweep_attr_set = false, lin_attr_set = false, solo_attr_set = false, solo_g_attr_set = false;
    int solo_max = -1;           // batches up to this size go to the block-per-instance kernel (diff, tric; 0: never; -1: four blocks per SM,
                                 // where the lane-cooperative kernel catches up: 592 instances 4.0 ms against 4.7 ms on B200)
    size_t ws_doubles_per_inst = 0;
    ModelInfo mi;
    nmpc_ipm_opts opts;
    // host mirrors of the tables
    std::vector<double> W, We, lbx, ubx, lbu, ubu, p;
    // device
    double *d_tab = nullptr;      // W | We | lbx | ubx | lbu | ubu | p | lti
    size_t off_W, off_We, off_lbx, off_ubx, off_lbu, off_ubu, off_p, off_lti, off_thr, off_stg, tab_doubles;
    int t_w = 0, trow = 0;
    bool tab_dirty = true, p_dirty = true;
    double* h_tab = nullptr;      // pinned mirror of the uploaded part of d_tab
    cudaEvent_t ev_tab = nullptr; // the last upload has left the mirror
    double *d_x = nullptr, *d_u = nullptr;       // persisted iterate, SoA, ld = cap
    double *d_ws = nullptr;                      // tile workspace for one chunk
    double *d_ctl_d = nullptr;                   // K3 control block of one chunk
    int *d_ctl_i = nullptr, *d_cnt = nullptr;    // d_cnt: act[iter_max+2], lanes entering iteration it
    int cnt_cap = 0;
    size_t tile_doubles = 0;
    int *d_qp_status = nullptr;
    // host-call staging
    double *d_stage_in = nullptr, *d_x0bar = nullptr, *d_yref = nullptr, *d_We = nullptr, *d_out = nullptr, *d_out_aos = nullptr;
    // small-batch host path (nmpc_rti_solve_host with B <= SMALL_HOST_B): one pinned staging buffer each way, one copy each way
    double *h_small = nullptr, *d_small = nullptr;
    int *d_status = nullptr, *d_iter = nullptr;
    double *d_stats = nullptr;                   // [8][cap] statistics of the last host call
    int last_host_B = 0;
    int *d_sqp_active = nullptr, *d_sqp_iter = nullptr, *d_sqp_qp = nullptr;   // SQP solve: per-instance bookkeeping [cap]
    unsigned long long* d_sqp_stepn = nullptr;
    double *d_roll_refs = nullptr, *d_roll_cmd = nullptr;                      // rollout engine: reference poses, commands
    double *d_vref = nullptr;                    // controller glue: carried reference states [nv][cap]
    double *d_cin = nullptr; int *d_cnref = nullptr;   // controller glue, host call: pose | vel | steer | cmd | refs (SoA, ld = B)
    cudaStream_t own_stream = nullptr;
    std::vector<cudaEvent_t> ev;                 // 4 per chunk + 2
    int n_ev_chunks = 0;
    int last_chunks = 0, last_launches = 0;
    cudaEvent_t ev_total[2] = {nullptr, nullptr};
};

extern "C" int nmpc_dims(int model, nmpc_dims_t* out)
{
    if (model < 0 || model > 2 || !out) return set_err(NMPC_E_ARG, "nmpc_dims: bad model");
    const ModelInfo& m = g_models[model];
    out->nx = m.nx; out->nu = m.nu; out->np = m.np; out->ny = m.nx + m.nu; out->nyn = m.nx;
    out->nbx = m.nv; out->nbu = m.nv; out->n = NSTAGE;
    return 0;
}

extern "C" void nmpc_default_opts(nmpc_ipm_opts* o)
{
    o->mu0 = 1.0; o->alpha_min = 1e-8;
    o->res_g_max = 1e-6; o->res_b_max = 1e-8; o->res_d_max = 1e-8; o->res_m_max = 1e-8;
    o->reg_prim = 1e-15; o->lam_min = 1e-16; o->t_min = 1e-16; o->tau_min = 1e-16; o->thr0 = 0.1;
    o->iter_max = 50; o->cond_pred_corr = 1;
}

extern "C" const char* nmpc_last_error(void) { return g_err; }

static size_t tile_doubles_of(int model)
{
    switch (model) {
        case 0: return Rti<DiffModel>::R::tile_doubles;
        case 1: return Rti<Omni4Model>::R::tile_doubles;
        default: return Rti<TricModel>::R::tile_doubles;
    }
}

extern "C" int nmpc_create(int model, int max_batch, int device, nmpc_solver** out)
{
    if (model < 0 || model > 2 || max_batch < 1 || !out) return set_err(NMPC_E_ARG, "nmpc_create: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_err(NMPC_E_NODEVICE, "nmpc_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return set_err(NMPC_E_ARG, "nmpc_create: bad device index");
    CK(cudaSetDevice(device));
    nmpc_solver* s = new (std::nothrow) nmpc_solver();
    if (!s) return set_err(NMPC_E_ARG, "nmpc_create: out of host memory");
    s->model = model; s->cap = max_batch; s->device = device; s->mi = g_models[model];
    nmpc_default_opts(&s->opts);
    const ModelInfo& m = s->mi;
    const int nx = m.nx, nu = m.nu, ny = nx + nu, nv = m.nv, n = NSTAGE;
    s->W.resize((size_t)n * ny); s->We.assign(m.QN, m.QN + nx);
    s->lbx.resize((size_t)n * nv); s->ubx.resize((size_t)n * nv); s->lbu.resize((size_t)n * nv); s->ubu.resize((size_t)n * nv);
    s->p.resize((size_t)n * m.np);
    for (int k = 0; k < n; k++) {
        for (int j = 0; j < nx; j++) s->W[(size_t)k * ny + j] = m.Q[j];
        for (int c = 0; c < nu; c++) s->W[(size_t)k * ny + nx + c] = m.Rw[c];
        for (int c = 0; c < nv; c++) {
            s->lbx[(size_t)k * nv + c] = m.lbx[c]; s->ubx[(size_t)k * nv + c] = m.ubx[c];
            s->lbu[(size_t)k * nv + c] = m.lbu[c]; s->ubu[(size_t)k * nv + c] = m.ubu[c];
        }
        for (int q = 0; q < m.np; q++) s->p[(size_t)k * m.np + q] = m.p[q];
    }
    size_t off = 0;
    s->off_W = off; off += (size_t)n * ny;
    s->off_We = off; off += nx;
    s->off_lbx = off; off += (size_t)n * nv;
    s->off_ubx = off; off += (size_t)n * nv;
    s->off_lbu = off; off += (size_t)n * nv;
    s->off_ubu = off; off += (size_t)n * nv;
    s->off_p = off; off += (size_t)n * m.np;
    s->off_lti = off; off += (size_t)n * 4 * nv;
    s->off_thr = off; off += (size_t)n * (1 + 3 * nv);
    s->t_w = (4 * nv + 2 + 1) & ~1; s->trow = s->t_w + ((ny + 1) & ~1);
    off = (off + 1) & ~(size_t)1;      // 16-byte aligned rows (cp.async)
    s->off_stg = off; off += (size_t)(n + 1) * s->trow;
    s->tab_doubles = off;
    // chunking bounds the workspace: NMPC_CHUNK instances per launch group (multiple of 32)
    int chunk = 131072;
    if (const char* e = getenv("NMPC_CHUNK")) { int v = atoi(e); if (v >= 32) chunk = v; }
    chunk = (chunk + LANES - 1) / LANES * LANES;
    const int cap_pad = (max_batch + LANES - 1) / LANES * LANES;
    s->chunk = chunk < cap_pad ? chunk : cap_pad;
    s->tile_doubles = tile_doubles_of(model);
    if (const char* e = getenv("NMPC_K3")) s->k3_group = !strcmp(e, "sweep") ? 0 : !strcmp(e, "group") ? 1 : 2;
    if (const char* e = getenv("NMPC_SOLO_MAX")) { int v = atoi(e); if (v >= 0) s->solo_max = v; }
    if (const char* e = getenv("NMPC_HYB_KMAX")) { int v = atoi(e); if (v >= 0 && v <= 1000) s->hyb_kmax = v; }
    if (const char* e = getenv("NMPC_HYB_MIN")) { int v = atoi(e); if (v >= 0) s->hyb_min = v; }
    if (model == 1) s->hyb_frac = 0.6;     // omni4: 103.4 ms per 65,536 instances against 104.9 at 0.75 (tools/gpu/sweep_frac.sh)
    if (const char* e = getenv("NMPC_HYB_FRAC")) { double v = atof(e); if (v >= 0.0 && v <= 1.0) s->hyb_frac = v; }
    {
        s->ws_doubles_per_inst = s->tile_doubles / LANES;
    }
    cudaError_t e;
#define CKC(call) do { e = (call); if (e != cudaSuccess) { set_err(NMPC_E_CUDA, #call, e); nmpc_destroy(s); return NMPC_E_CUDA; } } while (0)
    if (s->solo_max < 0) { int nsm = 0; CKC(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, s->device)); s->solo_max = (model == 1 ? 2 : 4) * nsm; }
    if (s->solo_max > max_batch) s->solo_max = max_batch;
    if (model == 1 && s->k3_group == 2 && s->solo_max > 0)
        CKC(cudaMalloc(&s->d_solo_scr, (size_t)s->solo_max * Solo<Omni4Model>::GSCR_DOUBLES * sizeof(double)));
    CKC(cudaMalloc(&s->d_tab, s->tab_doubles * sizeof(double)));
    CKC(cudaMalloc(&s->d_x, (size_t)max_batch * (n + 1) * nx * sizeof(double)));
    CKC(cudaMalloc(&s->d_u, (size_t)max_batch * n * nu * sizeof(double)));
    if (s->k3_group != 1) CKC(cudaMalloc(&s->d_ws, (size_t)s->chunk * s->ws_doubles_per_inst * sizeof(double)));
    if (s->k3_group != 0) {
        const size_t per_group = (model == 1) ? GRec<4>::inst_doubles : GRec<2>::inst_doubles;
        CKC(cudaMalloc(&s->d_ws_g, (size_t)s->chunk * per_group * sizeof(double)));
        CKC(cudaMalloc(&s->d_list, (size_t)s->chunk * sizeof(int)));
        CKC(cudaMalloc(&s->d_map, (size_t)s->chunk * sizeof(int)));
        CKC(cudaMalloc(&s->d_ctl_g, (size_t)s->chunk * 128));
    }
    CKC(cudaMalloc(&s->d_qp_status, (size_t)max_batch * sizeof(int)));
    CKC(cudaMalloc(&s->d_ctl_d, (size_t)NCTL_D * s->chunk * sizeof(double)));
    CKC(cudaMalloc(&s->d_ctl_i, (size_t)NCTL_I * s->chunk * sizeof(int)));
    s->cnt_cap = 1008 + 2 * 32;      // act[0..iter_max] of the sweep schedule (iter_max <= 1000), then the hand-over count and the group queue
    CKC(cudaMalloc(&s->d_cnt, (size_t)s->cnt_cap * sizeof(int)));
    CKC(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
    CKC(cudaEventCreate(&s->ev_total[0])); CKC(cudaEventCreate(&s->ev_total[1]));
    {   // default iterate: x_k = default x0 (scripts/<m>/generate_c_code.py:58-60), u = 0
        double x0def[11] = {0, 0, 3.14159265358979323846, 0, 0, 0, 0, 0, 0, 0, 0};
        double* d_def = s->d_tab;   // scratch: tables are uploaded later
        CKC(cudaMemcpy(d_def, x0def, sizeof(double) * nx, cudaMemcpyHostToDevice));
        k_init_iterate<<<1024, 256>>>(max_batch, nx, nu, d_def, s->d_x, s->d_u);
        CKC(cudaGetLastError());
        CKC(cudaDeviceSynchronize());
    }
#undef CKC
    *out = s;
    return 0;
}

extern "C" int nmpc_destroy(nmpc_solver* s)
{
    if (!s) return 0;
    cudaSetDevice(s->device);
    cudaFree(s->d_vref); cudaFree(s->d_cin); cudaFree(s->d_cnref); cudaFree(s->d_tab); cudaFree(s->d_x); cudaFree(s->d_u); cudaFree(s->d_ws); cudaFree(s->d_qp_status);
    cudaFree(s->d_ctl_d); cudaFree(s->d_ctl_i); cudaFree(s->d_cnt);
    cudaFree(s->d_solo_scr); cudaFree(s->d_ws_g); cudaFree(s->d_list); cudaFree(s->d_map); cudaFree(s->d_ctl_g);
    if (s->h_small) cudaFreeHost(s->h_small);
    cudaFree(s->d_small);
    cudaFree(s->d_stage_in); cudaFree(s->d_x0bar); cudaFree(s->d_yref); cudaFree(s->d_We); cudaFree(s->d_out); cudaFree(s->d_out_aos);
    cudaFree(s->d_status); cudaFree(s->d_iter); cudaFree(s->d_stats);
    cudaFree(s->d_sqp_active); cudaFree(s->d_sqp_iter); cudaFree(s->d_sqp_qp); cudaFree(s->d_sqp_stepn);
    cudaFree(s->d_roll_refs); cudaFree(s->d_roll_cmd);
    for (auto e : s->ev) cudaEventDestroy(e);
    if (s->ev_total[0]) cudaEventDestroy(s->ev_total[0]);
    if (s->ev_total[1]) cudaEventDestroy(s->ev_total[1]);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    if (s->ev_tab) cudaEventDestroy(s->ev_tab);
    if (s->h_tab) cudaFreeHost(s->h_tab);
    delete s;
    return 0;
}

extern "C" int nmpc_set_weights(nmpc_solver* s, const double* W_diag, const double* We_diag)
{
    if (!s) return set_err(NMPC_E_ARG, "null solver");
    if (W_diag) s->W.assign(W_diag, W_diag + s->W.size());
    if (We_diag) s->We.assign(We_diag, We_diag + s->We.size());
    s->tab_dirty = true;
    return 0;
}
extern "C" int nmpc_set_bounds(nmpc_solver* s, const double* lbx, const double* ubx, const double* lbu, const double* ubu)
{
    if (!s) return set_err(NMPC_E_ARG, "null solver");
    if (lbx) s->lbx.assign(lbx, lbx + s->lbx.size());
    if (ubx) s->ubx.assign(ubx, ubx + s->ubx.size());
    if (lbu) s->lbu.assign(lbu, lbu + s->lbu.size());
    if (ubu) s->ubu.assign(ubu, ubu + s->ubu.size());
    s->tab_dirty = true;
    return 0;
}
extern "C" int nmpc_set_params(nmpc_solver* s, const double* p)
{
    if (!s || !p) return set_err(NMPC_E_ARG, "null argument");
    s->p.assign(p, p + s->p.size());
    s->tab_dirty = true; s->p_dirty = true;
    return 0;
}
extern "C" int nmpc_get_tables(const nmpc_solver* s, double* W_diag, double* We_diag, double* lbx, double* ubx, double* lbu,
                               double* ubu, double* p)
{
    if (!s) return set_err(NMPC_E_ARG, "null solver");
    if (W_diag) memcpy(W_diag, s->W.data(), s->W.size() * 8);
    if (We_diag) memcpy(We_diag, s->We.data(), s->We.size() * 8);
    if (lbx) memcpy(lbx, s->lbx.data(), s->lbx.size() * 8);
    if (ubx) memcpy(ubx, s->ubx.data(), s->ubx.size() * 8);
    if (lbu) memcpy(lbu, s->lbu.data(), s->lbu.size() * 8);
    if (ubu) memcpy(ubu, s->ubu.data(), s->ubu.size() * 8);
    if (p) memcpy(p, s->p.data(), s->p.size() * 8);
    return 0;
}
extern "C" int nmpc_set_opts(nmpc_solver* s, const nmpc_ipm_opts* o)
{
    if (!s || !o) return set_err(NMPC_E_ARG, "null argument");
    if (o->iter_max < 0 || o->iter_max > 1000) return set_err(NMPC_E_ARG, "iter_max out of range");
    s->opts = *o;
    return 0;
}
extern "C" int nmpc_get_opts(const nmpc_solver* s, nmpc_ipm_opts* o)
{
    if (!s || !o) return set_err(NMPC_E_ARG, "null argument");
    *o = s->opts;
    return 0;
}

template <class M>
static int launch_lti(nmpc_solver* s, cudaStream_t st)
{
    k_lti_setup<M><<<1, 128, 0, st>>>(s->d_tab + s->off_p, OCP_DT, s->d_tab + s->off_lti, s->d_tab + s->off_thr);
    CK(cudaGetLastError());
    return 0;
}

static int upload_tables(nmpc_solver* s, cudaStream_t st)
{
    if (!s->tab_dirty) return 0;
    // one asynchronous copy of the packed tables from a pinned mirror, on the stream of the kernels that read them;
    // the host only waits if the previous upload has not left the mirror yet
    if (!s->h_tab) {
        CK(cudaMallocHost(&s->h_tab, s->off_lti * sizeof(double)));
        CK(cudaEventCreateWithFlags(&s->ev_tab, cudaEventDisableTiming));
    } else CK(cudaEventSynchronize(s->ev_tab));
    double* h = s->h_tab;
    memcpy(&h[s->off_W], s->W.data(), s->W.size() * 8);
    memcpy(&h[s->off_We], s->We.data(), s->We.size() * 8);
    memcpy(&h[s->off_lbx], s->lbx.data(), s->lbx.size() * 8);
    memcpy(&h[s->off_ubx], s->ubx.data(), s->ubx.size() * 8);
    memcpy(&h[s->off_lbu], s->lbu.data(), s->lbu.size() * 8);
    memcpy(&h[s->off_ubu], s->ubu.data(), s->ubu.size() * 8);
    memcpy(&h[s->off_p], s->p.data(), s->p.size() * 8);
    CK(cudaMemcpyAsync(s->d_tab, h, s->off_lti * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(s->ev_tab, st));
    if (s->p_dirty) {
        int rc = s->model == 0 ? launch_lti<DiffModel>(s, st) : s->model == 1 ? launch_lti<Omni4Model>(s, st) : launch_lti<TricModel>(s, st);
        if (rc) return rc;
        s->p_dirty = false;
    }
    k_stage_table<<<NSTAGE + 1, 64, 0, st>>>(s->mi.nv, s->mi.nx + s->mi.nu, s->t_w, s->trow, s->d_tab + s->off_lti, s->d_tab + s->off_W,
                                             s->d_tab + s->off_stg);
    CK(cudaGetLastError());
    s->tab_dirty = false;
    return 0;
}

static Tables make_tables(const nmpc_solver* s)
{
    Tables tb;
    tb.W = s->d_tab + s->off_W; tb.We = s->d_tab + s->off_We;
    tb.lbx = s->d_tab + s->off_lbx; tb.ubx = s->d_tab + s->off_ubx;
    tb.lbu = s->d_tab + s->off_lbu; tb.ubu = s->d_tab + s->off_ubu;
    tb.p = s->d_tab + s->off_p; tb.lti = s->d_tab + s->off_lti; tb.stg = s->d_tab + s->off_stg; tb.thr = s->d_tab + s->off_thr;
    tb.dt = OCP_DT;
    return tb;
}

static IpmOpts to_core_opts(const nmpc_ipm_opts& a)
{
    IpmOpts o;
    o.mu0 = a.mu0; o.alpha_min = a.alpha_min; o.res_g_max = a.res_g_max; o.res_b_max = a.res_b_max;
    o.res_d_max = a.res_d_max; o.res_m_max = a.res_m_max; o.reg_prim = a.reg_prim; o.lam_min = a.lam_min;
    o.t_min = a.t_min; o.tau_min = a.tau_min; o.thr0 = a.thr0; o.iter_max = a.iter_max; o.cond_pred_corr = a.cond_pred_corr;
    return o;
}

static int ensure_events(nmpc_solver* s, int nchunks)
{
    while (s->n_ev_chunks < nchunks) {
        for (int q = 0; q < 4; q++) { cudaEvent_t e; CK(cudaEventCreate(&e)); s->ev.push_back(e); }
        s->n_ev_chunks++;
    }
    return 0;
}

// ---- group path: one persistent launch per chunk ------------------------------------------------
typedef void (*grp_kernel_t)(int, int, Tables, const double*, int, IpmOpts, double*, int*, GrpOut, GrpResume);
template <class GP>
static int launch_group_k(nmpc_solver* s, grp_kernel_t kern, int i0, int n, const Tables& tb, const double* d_We, int ldWe,
                          const IpmOpts& o, const GrpOut& out, const GrpResume& rs, cudaStream_t st)
{
    const size_t smem = (size_t)GRP_WARPS * GP::WARP_D * sizeof(double);
    int& blocks_per_sm = s->grp_blocks_per_sm;
    if (!blocks_per_sm) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, GRP_WARPS * 32, smem));
        if (nb < 1) return set_err(NMPC_E_CUDA, "k_ipm_coop does not fit on an SM");
        if (const char* e = getenv("NMPC_GRP_BPS")) { int v = atoi(e); if (v >= 1 && v < nb) nb = v; }
        blocks_per_sm = nb;
    }
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, s->device));
    const int per_block = GRP_WARPS * GP::NSLOT;
    int blocks = nsm * blocks_per_sm;
    const int need = (n + per_block - 1) / per_block;        // n: upper bound of the queue length (resume: the chunk)
    if (blocks > need) blocks = need;
    s->grp_blocks = blocks;
    int* queue = s->d_cnt + s->cnt_cap - 1;
    CK(cudaMemsetAsync(queue, 0, sizeof(int), st));
    kern<<<blocks, GRP_WARPS * 32, smem, st>>>(i0, n, tb, d_We, ldWe, o, s->d_ws_g, queue, out, rs);
    CK(cudaGetLastError());
    return 0;
}
template <class M>
static int launch_group_any(nmpc_solver* s, int i0, int n, const Tables& tb, const double* d_We, int ldWe, const IpmOpts& o,
                            const GrpOut& out, const GrpResume& rs, cudaStream_t st)
{
    using S = Rti<M>;
    // G = 4 nv lanes per instance; 12 warps / SM for nv = 2 (168 registers), 8 for nv = 4 (no spills at 244 registers)
    return launch_group_k<Coop<M, 4 * S::NV>>(s, k_ipm_coop<M, 4 * S::NV, (S::NV == 2 ? NMPC_COOP_MINB : 2)>, i0, n, tb, d_We, ldWe, o, out, rs, st);
}

template <class M>
static int solve_device_t(nmpc_solver* s, int B, const double* d_x0bar, const double* d_yref, int nyref, const double* d_We,
                          double* d_x, double* d_u, int ld, int* d_status, int* d_qp_iter, double* d_stats, cudaStream_t st)
{
    const Tables tb = make_tables(s);
    const IpmOpts o = to_core_opts(s->opts);
    const int nchunks = (B + s->chunk - 1) / s->chunk;
    int rc = ensure_events(s, nchunks);
    if (rc) return rc;
    s->last_chunks = nchunks; s->last_launches = 0;
    k_fill_int<<<(B + 255) / 256, 256, 0, st>>>(B, d_status, 0);
    s->last_launches++;
    for (int c = 0; c < nchunks; c++) {
        const int i0 = c * s->chunk;
        const int n = (B - i0) < s->chunk ? (B - i0) : s->chunk;
        cudaEvent_t* ev = &s->ev[(size_t)c * 4];
        CK(cudaEventRecord(ev[0], st));
        dim3 g1((n + LIN_BLOCK - 1) / LIN_BLOCK, NSTAGE + 1);
        k_linearize<M><<<g1, LIN_BLOCK, 0, st>>>(B, i0, n, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ld, tb, s->d_ws);
        CK(cudaEventRecord(ev[1], st));
        {
            using S = Rti<M>;
            const int ldc = s->chunk;
            const int nb = (n + LANES * SW_TILES - 1) / (LANES * SW_TILES), nt = LANES * SW_TILES;
            int* act = s->d_cnt;                 // act[it]: lanes entering iteration it
            const size_t smB = (size_t)(S::CarryB::SC_N + ((NMPC_B_STAGE_IMAGE && !S::LEAN) ? S::R::DZA : 0)) * NMPC_SCRATCH_STRIDE * sizeof(double);
            bool& attr_set = s->sweep_attr_set;
            if (!attr_set) {
                CK(cudaFuncSetAttribute(k_sweep<M, S::SW_B_FIRST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smB));
                CK(cudaFuncSetAttribute(k_sweep<M, S::SW_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smB));
                attr_set = true;
            }
            CK(cudaMemsetAsync(s->d_cnt, 0, (size_t)s->cnt_cap * sizeof(int), st));
            k_sweep<M, S::SW_B_FIRST><<<nb, nt, smB, st>>>(B, i0, n, ldc, tb, d_We, o, s->d_ws, s->d_ctl_d, s->d_ctl_i, act, act, 1, 0);
            s->last_launches++;
            // hybrid schedule: the lockstep sweeps run while at least gate_min instances iterate (a sweep costs the
            // same whether 100% or 10% of the lanes are active), at most kmax iterations; whatever is unfinished is
            // handed to the persistent lane-group kernel, which has no lockstep between instances
            const bool hybrid = s->k3_group == 2;
            const int kmax = hybrid ? (s->hyb_kmax < o.iter_max ? s->hyb_kmax : o.iter_max) : o.iter_max;
            int gate_min = 1;
            if (hybrid) { gate_min = (int)(s->hyb_frac * n); if (gate_min < 1) gate_min = 1; }
            for (int it = 0; it < kmax; it++) {
                k_sweep<M, S::SW_FDF><<<nb, nt, 0, st>>>(B, i0, n, ldc, tb, d_We, o, s->d_ws, s->d_ctl_d, s->d_ctl_i, act + it, nullptr, gate_min, hybrid ? 1 : 0);
                k_sweep<M, S::SW_B><<<nb, nt, smB, st>>>(B, i0, n, ldc, tb, d_We, o, s->d_ws, s->d_ctl_d, s->d_ctl_i, act + it, act + it + 1, gate_min, 0);
                s->last_launches += 2;
            }
            k_ipm_finish<<<(n + 255) / 256, 256, 0, st>>>(B, i0, n, ldc, s->d_ctl_d, s->d_ctl_i, s->d_qp_status, d_qp_iter, d_stats);
            s->last_launches++;
            if (hybrid) {
                int* nres = s->d_cnt + s->cnt_cap - 2;
                int* bcnt = s->d_cnt + s->cnt_cap - 2 - 2 * HB_NB;      // zeroed with the counters at the start of the chunk
                k_handover_count<<<(n + 255) / 256, 256, 0, st>>>(n, ldc, s->d_ctl_d, s->d_ctl_i, bcnt, s->d_map);
                k_handover_assign<M><<<(n + 255) / 256, 256, 0, st>>>(n, ldc, s->d_ctl_d, s->d_ctl_i, bcnt, nres, s->d_list, s->d_map, s->d_ctl_g);
                dim3 gc((n + CV_BLOCK - 1) / CV_BLOCK, NSTAGE + 1);
                k_handover_convert<M><<<gc, CV_BLOCK, 0, st>>>(n, s->d_map, s->d_ws, s->d_ws_g, tb.thr);
                const GrpOut out{s->d_qp_status, d_qp_iter, d_stats, B};
                rc = launch_group_any<M>(s, i0, n, tb, d_We, B, o, out, GrpResume{nres, s->d_list, s->d_ctl_g}, st);
                if (rc) return rc;
                s->last_launches += 4;
            }
        }
        CK(cudaEventRecord(ev[2], st));
        if (s->k3_group == 2)
            k_step_h<M><<<g1, LIN_BLOCK, 0, st>>>(B, i0, n, d_x0bar, d_x, d_u, ld, s->d_ws, s->d_ws_g, s->d_map, s->d_qp_status, d_status);
        else
            k_step<M><<<g1, LIN_BLOCK, 0, st>>>(B, i0, n, d_x0bar, d_x, d_u, ld, s->d_ws, s->d_qp_status, d_status);
        CK(cudaEventRecord(ev[3], st));
        s->last_launches += 3;
    }
    CK(cudaGetLastError());
    return 0;
}


#if defined(NMPC_SOLO_PROF)
// phase cycles of block 0 of the block-per-instance kernel since the last call (experiment builds only)
extern "C" int nmpc_solo_prof(unsigned long long* out)
{
    CK(cudaMemcpyFromSymbol(out, g_solo_prof, SOLO_NPROF * sizeof(unsigned long long)));
    unsigned long long z[SOLO_NPROF] = {0};
    CK(cudaMemcpyToSymbol(g_solo_prof, z, sizeof(z)));
    return 0;
}
#endif

// block-per-instance schedule (rti_solo.cuh): K1 / K2 into the tile layout, one block per instance, K4 from the tiles
template <class M>
static int solve_device_solo(nmpc_solver* s, int B, const double* d_x0bar, const double* d_yref, int nyref, const double* d_We,
                             double* d_x, double* d_u, int ld, int* d_status, int* d_qp_iter, double* d_stats, cudaStream_t st)
{
    using SO = Solo<M>;
    const Tables tb = make_tables(s);
    const IpmOpts o = to_core_opts(s->opts);
    const int nchunks = (B + s->chunk - 1) / s->chunk;
    int rc = ensure_events(s, nchunks);
    if (rc) return rc;
    s->last_chunks = nchunks; s->last_launches = 0;
    if (!s->solo_attr_set) {
        CK(cudaFuncSetAttribute(k_ipm_solo<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SO::SM_BYTES));
        s->solo_attr_set = true;
    }
    k_fill_int<<<(B + 255) / 256, 256, 0, st>>>(B, d_status, 0);
    s->last_launches++;
    for (int c = 0; c < nchunks; c++) {
        const int i0 = c * s->chunk;
        const int n = (B - i0) < s->chunk ? (B - i0) : s->chunk;
        cudaEvent_t* ev = &s->ev[(size_t)c * 4];
        CK(cudaEventRecord(ev[0], st));
        dim3 g1((n + LIN_BLOCK - 1) / LIN_BLOCK, NSTAGE + 1);
        k_linearize<M><<<g1, LIN_BLOCK, 0, st>>>(B, i0, n, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ld, tb, s->d_ws);
        CK(cudaEventRecord(ev[1], st));
        k_ipm_solo<M><<<n, SO::THREADS, SO::SM_BYTES, st>>>(B, i0, tb, d_We, o, s->d_ws, s->d_qp_status, d_qp_iter, d_stats, s->d_solo_scr);
        CK(cudaEventRecord(ev[2], st));
        k_step<M><<<g1, LIN_BLOCK, 0, st>>>(B, i0, n, d_x0bar, d_x, d_u, ld, s->d_ws, s->d_qp_status, d_status);
        CK(cudaEventRecord(ev[3], st));
        s->last_launches += 3;
    }
    CK(cudaGetLastError());
    return 0;
}

// d_active / d_stepn (SQP passes, else null): only the instances with a non-zero flag are linearised, solved (through a
// compacted queue) and stepped; the inf-norm of their step is accumulated into d_stepn
template <class M>
static int solve_device_group(nmpc_solver* s, int B, const double* d_x0bar, const double* d_yref, int nyref, const double* d_We,
                              double* d_x, double* d_u, int ld, int* d_status, int* d_qp_iter, double* d_stats, cudaStream_t st,
                              const int* d_active = nullptr, unsigned long long* d_stepn = nullptr)
{
    using S = Rti<M>;
    using GR = GRec<S::NV>;
    const Tables tb = make_tables(s);
    const IpmOpts o = to_core_opts(s->opts);
    const int nchunks = (B + s->chunk - 1) / s->chunk;
    int rc = ensure_events(s, nchunks);
    if (rc) return rc;
    s->last_chunks = nchunks; s->last_launches = 0;
    if (!d_active) {
        k_fill_int<<<(B + 255) / 256, 256, 0, st>>>(B, d_status, 0);
        s->last_launches++;
    }
    const size_t sm_lin = (size_t)LING_BLOCK * ((GR::LHD + GR::NREC - GR::MC) | 1) * sizeof(double);
    bool& attr_set = s->lin_attr_set;
    if (!attr_set) {
        CK(cudaFuncSetAttribute(k_linearize_g<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_lin));
        attr_set = true;
    }
    const GrpOut out{s->d_qp_status, d_qp_iter, d_stats, B};
    for (int c = 0; c < nchunks; c++) {
        const int i0 = c * s->chunk;
        const int n = (B - i0) < s->chunk ? (B - i0) : s->chunk;
        cudaEvent_t* ev = &s->ev[(size_t)c * 4];
        CK(cudaEventRecord(ev[0], st));
        dim3 g1((n + LIN_BLOCK - 1) / LIN_BLOCK, NSTAGE + 1), g0((n + LING_BLOCK - 1) / LING_BLOCK, NSTAGE + 1);
        k_linearize_g<M><<<g0, LING_BLOCK, sm_lin, st>>>(B, i0, n, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ld, tb, o, s->d_ws_g, d_active);
        CK(cudaEventRecord(ev[1], st));
        GrpResume rs{nullptr, nullptr, nullptr};
        if (d_active) {
            int* cnt = s->d_cnt + s->cnt_cap - 2;
            CK(cudaMemsetAsync(cnt, 0, sizeof(int), st));
            k_sqp_list<<<(n + 255) / 256, 256, 0, st>>>(i0, n, d_active, cnt, s->d_list);
            rs = GrpResume{cnt, s->d_list, nullptr};
            s->last_launches++;
        }
        bool solo = false;
        {
            // small batches (the SQP passes of a small fleet): one block per queued instance (rti_solo.cuh)
            if (s->k3_group == 2 && n <= s->solo_max) {
                using SO = Solo<M>;
                if (!s->solo_g_attr_set) {
                    CK(cudaFuncSetAttribute(k_ipm_solo_g<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SO::SM_BYTES));
                    s->solo_g_attr_set = true;
                }
                k_ipm_solo_g<M><<<n, SO::THREADS, SO::SM_BYTES, st>>>(i0, n, tb, d_We, B, o, s->d_ws_g, out, rs, s->d_solo_scr);
                solo = true;
            }
        }
        if (!solo) {
            rc = launch_group_any<M>(s, i0, n, tb, d_We, B, o, out, rs, st);
            if (rc) return rc;
        }
        CK(cudaEventRecord(ev[2], st));
        k_step_g<M><<<g1, LIN_BLOCK, 0, st>>>(B, i0, n, d_x0bar, d_x, d_u, ld, s->d_ws_g, s->d_qp_status, d_status, d_active, d_stepn);
        CK(cudaEventRecord(ev[3], st));
        s->last_launches += 3;
    }
    CK(cudaGetLastError());
    return 0;
}

extern "C" int nmpc_rti_solve_device(nmpc_solver* s, int B, const double* d_x0bar, const double* d_yref, int nyref,
                                     const double* d_We, double* d_x, double* d_u, int ldxu,
                                     int* d_status, int* d_qp_iter, double* d_stats, void* stream)
{
    if (!s || !d_x0bar || !d_yref || !d_status || !d_qp_iter) return set_err(NMPC_E_ARG, "nmpc_rti_solve_device: null argument");
    if (B < 1) return set_err(NMPC_E_ARG, "nmpc_rti_solve_device: B < 1");
    if (B > s->cap) return set_err(NMPC_E_CAPACITY, "nmpc_rti_solve_device: batch exceeds capacity");
    const int ny = s->mi.nx + s->mi.nu;
    if (nyref != 3 && nyref != ny) return set_err(NMPC_E_ARG, "nmpc_rti_solve_device: nyref must be 3 or ny");
    if ((d_x == nullptr) != (d_u == nullptr)) return set_err(NMPC_E_ARG, "nmpc_rti_solve_device: pass both d_x and d_u or neither");
    CK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream, as in CUDA
    int rc = upload_tables(s, st);
    if (rc) return rc;
    if (!d_x) { d_x = s->d_x; d_u = s->d_u; ldxu = s->cap; }
    if (ldxu < B) return set_err(NMPC_E_ARG, "nmpc_rti_solve_device: leading dimension < B");
    CK(cudaEventRecord(s->ev_total[0], st));
    if (s->k3_group == 2 && B <= s->solo_max && s->d_ws) {
        // latency path: one block per instance
        if (s->model == 0) rc = solve_device_solo<DiffModel>(s, B, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ldxu, d_status, d_qp_iter, d_stats, st);
        else if (s->model == 1) rc = solve_device_solo<Omni4Model>(s, B, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ldxu, d_status, d_qp_iter, d_stats, st);
        else rc = solve_device_solo<TricModel>(s, B, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ldxu, d_status, d_qp_iter, d_stats, st);
    } else if (s->k3_group == 1 || (s->k3_group == 2 && B < s->hyb_min)) {
        switch (s->model) {
            case 0: rc = solve_device_group<DiffModel>(s, B, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ldxu, d_status, d_qp_iter, d_stats, st); break;
            case 1: rc = solve_device_group<Omni4Model>(s, B, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ldxu, d_status, d_qp_iter, d_stats, st); break;
            default: rc = solve_device_group<TricModel>(s, B, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ldxu, d_status, d_qp_iter, d_stats, st); break;
        }
    } else switch (s->model) {
        case 0: rc = solve_device_t<DiffModel>(s, B, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ldxu, d_status, d_qp_iter, d_stats, st); break;
        case 1: rc = solve_device_t<Omni4Model>(s, B, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ldxu, d_status, d_qp_iter, d_stats, st); break;
        default: rc = solve_device_t<TricModel>(s, B, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ldxu, d_status, d_qp_iter, d_stats, st); break;
    }
    if (rc) return rc;
    CK(cudaEventRecord(s->ev_total[1], st));
    return 0;
}

// ---- BASELINE config 4: warm-start shift and SQP to convergence (declared in include/nmpc_b200.h) ------------------
extern "C" int nmpc_shift_device(nmpc_solver* s, int B, double* d_x, double* d_u, int ldxu, const int* d_mask, void* stream)
{
    if (!s) return set_err(NMPC_E_ARG, "nmpc_shift_device: null solver");
    if (B < 1) return set_err(NMPC_E_ARG, "nmpc_shift_device: B < 1");
    if ((d_x == nullptr) != (d_u == nullptr)) return set_err(NMPC_E_ARG, "nmpc_shift_device: pass both d_x and d_u or neither");
    if (!d_x) { if (B > s->cap) return set_err(NMPC_E_CAPACITY, "nmpc_shift_device: batch exceeds capacity"); d_x = s->d_x; d_u = s->d_u; ldxu = s->cap; }
    if (ldxu < B) return set_err(NMPC_E_ARG, "nmpc_shift_device: leading dimension < B");
    CK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    k_shift<<<dim3((B + 127) / 128, s->mi.nx), 128, 0, st>>>(B, ldxu, s->mi.nx, NSTAGE + 1, d_x, d_mask);
    k_shift<<<dim3((B + 127) / 128, s->mi.nu), 128, 0, st>>>(B, ldxu, s->mi.nu, NSTAGE, d_u, d_mask);
    CK(cudaGetLastError());
    return 0;
}

static int ensure_staging(nmpc_solver* s);
static int ensure_sqp(nmpc_solver* s)
{
    if (s->d_sqp_active) return 0;
    CK(cudaMalloc(&s->d_sqp_active, (size_t)s->cap * sizeof(int)));
    CK(cudaMalloc(&s->d_sqp_iter, (size_t)s->cap * sizeof(int)));
    CK(cudaMalloc(&s->d_sqp_qp, (size_t)s->cap * sizeof(int)));
    CK(cudaMalloc(&s->d_sqp_stepn, (size_t)s->cap * sizeof(unsigned long long)));
    return 0;
}

static int group_pass(nmpc_solver* s, int B, const double* d_x0bar, const double* d_yref, int nyref, const double* d_We,
                      double* d_x, double* d_u, int ld, int* d_status, int* d_qp_iter, cudaStream_t st, const int* d_active,
                      unsigned long long* d_stepn)
{
    switch (s->model) {
        case 0: return solve_device_group<DiffModel>(s, B, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ld, d_status, d_qp_iter, nullptr, st, d_active, d_stepn);
        case 1: return solve_device_group<Omni4Model>(s, B, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ld, d_status, d_qp_iter, nullptr, st, d_active, d_stepn);
        default: return solve_device_group<TricModel>(s, B, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ld, d_status, d_qp_iter, nullptr, st, d_active, d_stepn);
    }
}

extern "C" int nmpc_sqp_solve_device(nmpc_solver* s, int B, const double* d_x0bar, const double* d_yref, int nyref,
                                     const double* d_We, double* d_x, double* d_u, int ldxu, int max_iter, double tol,
                                     int* d_status, int* d_sqp_iter, int* d_qp_iter, void* stream)
{
    if (!s || !d_x0bar || !d_yref || !d_status) return set_err(NMPC_E_ARG, "nmpc_sqp_solve_device: null argument");
    if (B < 1) return set_err(NMPC_E_ARG, "nmpc_sqp_solve_device: B < 1");
    if (B > s->cap) return set_err(NMPC_E_CAPACITY, "nmpc_sqp_solve_device: batch exceeds capacity");
    if (max_iter < 1 || max_iter > 1000) return set_err(NMPC_E_ARG, "nmpc_sqp_solve_device: max_iter must be 1..1000");
    if (!(tol >= 0.0)) return set_err(NMPC_E_ARG, "nmpc_sqp_solve_device: tol must be >= 0");
    const int ny = s->mi.nx + s->mi.nu;
    if (nyref != 3 && nyref != ny) return set_err(NMPC_E_ARG, "nmpc_sqp_solve_device: nyref must be 3 or ny");
    if ((d_x == nullptr) != (d_u == nullptr)) return set_err(NMPC_E_ARG, "nmpc_sqp_solve_device: pass both d_x and d_u or neither");
    if (!s->d_ws_g) return set_err(NMPC_E_ARG, "nmpc_sqp_solve_device: needs the lane-group workspace (NMPC_K3=sweep excludes it)");
    CK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = upload_tables(s, st); if (rc) return rc;
    rc = ensure_sqp(s); if (rc) return rc;
    rc = ensure_staging(s); if (rc) return rc;
    if (!d_x) { d_x = s->d_x; d_u = s->d_u; ldxu = s->cap; }
    if (ldxu < B) return set_err(NMPC_E_ARG, "nmpc_sqp_solve_device: leading dimension < B");
    CK(cudaEventRecord(s->ev_total[0], st));
    const int nb = (B + 255) / 256;
    k_sqp_begin<<<nb, 256, 0, st>>>(B, s->d_sqp_active, s->d_sqp_iter, s->d_sqp_qp, s->d_sqp_stepn);
    k_fill_int<<<nb, 256, 0, st>>>(B, d_status, 0);
    int launches = 2;
    // every pass is enqueued; a pass only touches the instances still iterating (their queue is built on the device),
    // so the passes after the last instance stopped cost a handful of empty launches
    for (int it = 0; it < max_iter; it++) {
        rc = group_pass(s, B, d_x0bar, d_yref, nyref, d_We, d_x, d_u, ldxu, d_status, s->d_iter, st, s->d_sqp_active, s->d_sqp_stepn);
        if (rc) return rc;
        k_sqp_update<<<nb, 256, 0, st>>>(B, max_iter, tol, s->d_sqp_active, d_status, s->d_iter, s->d_sqp_iter, s->d_sqp_qp, s->d_sqp_stepn);
        launches += s->last_launches + 1;
    }
    if (d_sqp_iter) CK(cudaMemcpyAsync(d_sqp_iter, s->d_sqp_iter, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (d_qp_iter) CK(cudaMemcpyAsync(d_qp_iter, s->d_sqp_qp, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, st));
    CK(cudaGetLastError());
    s->last_launches = launches;
    CK(cudaEventRecord(s->ev_total[1], st));
    return 0;
}

extern "C" int nmpc_iterate_device(nmpc_solver* s, double** d_x, double** d_u, int* leading_dim)
{
    if (!s) return set_err(NMPC_E_ARG, "null solver");
    if (d_x) *d_x = s->d_x;
    if (d_u) *d_u = s->d_u;
    if (leading_dim) *leading_dim = s->cap;
    return 0;
}

extern "C" int nmpc_reset(nmpc_solver* s)
{
    if (!s) return set_err(NMPC_E_ARG, "null solver");
    CK(cudaSetDevice(s->device));
    CK(cudaMemsetAsync(s->d_x, 0, (size_t)s->cap * (NSTAGE + 1) * s->mi.nx * sizeof(double), s->own_stream));
    CK(cudaMemsetAsync(s->d_u, 0, (size_t)s->cap * NSTAGE * s->mi.nu * sizeof(double), s->own_stream));
    CK(cudaStreamSynchronize(s->own_stream));
    return 0;
}

extern "C" int nmpc_reset_async(nmpc_solver* s, void* stream)
{
    if (!s) return set_err(NMPC_E_ARG, "null solver");
    CK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the default stream
    CK(cudaMemsetAsync(s->d_x, 0, (size_t)s->cap * (NSTAGE + 1) * s->mi.nx * sizeof(double), st));
    CK(cudaMemsetAsync(s->d_u, 0, (size_t)s->cap * NSTAGE * s->mi.nu * sizeof(double), st));
    return 0;
}

static int ensure_staging(nmpc_solver* s)
{
    if (s->d_stage_in) return 0;
    const int nx = s->mi.nx, nu = s->mi.nu, ny = nx + nu;
    const size_t cap = s->cap;
    const size_t in_rows = (size_t)(NSTAGE + 1) * nx > (size_t)(NSTAGE + 1) * ny ? (size_t)(NSTAGE + 1) * nx : (size_t)(NSTAGE + 1) * ny;
    CK(cudaMalloc(&s->d_stage_in, cap * in_rows * sizeof(double)));
    CK(cudaMalloc(&s->d_x0bar, cap * nx * sizeof(double)));
    CK(cudaMalloc(&s->d_yref, cap * (NSTAGE + 1) * ny * sizeof(double)));
    CK(cudaMalloc(&s->d_We, cap * nx * sizeof(double)));
    CK(cudaMalloc(&s->d_out, cap * (nx + nu) * sizeof(double)));
    CK(cudaMalloc(&s->d_out_aos, cap * (nx + nu) * sizeof(double)));
    CK(cudaMalloc(&s->d_status, cap * sizeof(int)));
    CK(cudaMalloc(&s->d_iter, cap * sizeof(int)));
    CK(cudaMalloc(&s->d_stats, cap * 8 * sizeof(double)));
    return 0;
}

static int h2d_transpose(nmpc_solver* s, int B, int R, const double* h, double* d_soa, int ld, cudaStream_t st)
{
    CK(cudaMemcpyAsync(s->d_stage_in, h, (size_t)B * R * sizeof(double), cudaMemcpyHostToDevice, st));
    dim3 g((B + 31) / 32, (R + 31) / 32), b(32, 8);
    k_aos_to_soa_ld<<<g, b, 0, st>>>(B, R, ld, s->d_stage_in, d_soa);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int nmpc_set_iterate_host(nmpc_solver* s, int B, const double* x, const double* u)
{
    if (!s || !x || !u) return set_err(NMPC_E_ARG, "null argument");
    if (B < 1 || B > s->cap) return set_err(NMPC_E_CAPACITY, "batch exceeds capacity");
    CK(cudaSetDevice(s->device));
    int rc = ensure_staging(s); if (rc) return rc;
    cudaStream_t st = s->own_stream;
    rc = h2d_transpose(s, B, (NSTAGE + 1) * s->mi.nx, x, s->d_x, s->cap, st); if (rc) return rc;
    CK(cudaStreamSynchronize(st));
    rc = h2d_transpose(s, B, NSTAGE * s->mi.nu, u, s->d_u, s->cap, st); if (rc) return rc;
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int nmpc_get_iterate_host(nmpc_solver* s, int B, double* x, double* u)
{
    if (!s || !x || !u) return set_err(NMPC_E_ARG, "null argument");
    if (B < 1 || B > s->cap) return set_err(NMPC_E_CAPACITY, "batch exceeds capacity");
    CK(cudaSetDevice(s->device));
    int rc = ensure_staging(s); if (rc) return rc;
    cudaStream_t st = s->own_stream;
    dim3 b(32, 8);
    {
        const int R = (NSTAGE + 1) * s->mi.nx;
        dim3 g((B + 31) / 32, (R + 31) / 32);
        k_soa_to_aos<<<g, b, 0, st>>>(B, R, s->cap, s->d_x, s->d_stage_in);
        CK(cudaMemcpyAsync(x, s->d_stage_in, (size_t)B * R * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    {
        const int R = NSTAGE * s->mi.nu;
        dim3 g((B + 31) / 32, (R + 31) / 32);
        k_soa_to_aos<<<g, b, 0, st>>>(B, R, s->cap, s->d_u, s->d_stage_in);
        CK(cudaMemcpyAsync(u, s->d_stage_in, (size_t)B * R * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return 0;
}

// ---- small-batch host path: the ROS drop-in (one robot per call) and small fleets.  What the general path below does with one
// copy and one transposition kernel per argument (nine launches and six copies around the solve, the device-to-host ones from
// pageable memory and therefore each a round trip of its own) is done here with ONE pinned staging buffer each way: inputs
// packed [x0bar | yref | We] on the host, one H2D, one kernel that scatters them into the structure-of-arrays buffers, the solve,
// one kernel that gathers [u_0 | x_1 | status | qp_iter], one D2H.
constexpr int SMALL_HOST_B = 64;
__global__ void k_small_in(int B, int nx, int R, int hasWe, const double* __restrict__ in, double* __restrict__ x0bar,
                           double* __restrict__ yref, double* __restrict__ We)
{
    const int per = nx + R + (hasWe ? nx : 0);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < B * per; e += gridDim.x * blockDim.x) {
        const int i = e / per, r = e - i * per;                       // output index: instance fastest within a row
        // rows 0..nx-1: x0bar, nx..nx+R-1: yref, then We; the packed input holds the three arrays back to back, each instance-major
        if (r < nx) x0bar[(size_t)r * B + i] = in[(size_t)i * nx + r];
        else if (r < nx + R) yref[(size_t)(r - nx) * B + i] = in[(size_t)B * nx + (size_t)i * R + (r - nx)];
        else We[(size_t)(r - nx - R) * B + i] = in[(size_t)B * (nx + R) + (size_t)i * nx + (r - nx - R)];
    }
}
__global__ void k_small_out(int B, int nx, int nu, int ld, const double* __restrict__ x, const double* __restrict__ u,
                            const int* __restrict__ status, const int* __restrict__ iter, double* __restrict__ out)
{
    const int per = nu + nx;
    int* io = reinterpret_cast<int*>(out + (size_t)B * per);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < B * (per + 1); e += gridDim.x * blockDim.x) {
        if (e < B * per) {
            const int i = e / per, r = e - i * per;
            if (r < nu) out[(size_t)i * nu + r] = u[(size_t)r * ld + i];                               // u_0
            else out[(size_t)B * nu + (size_t)i * nx + (r - nu)] = x[(size_t)(nx + r - nu) * ld + i];  // x_1
        } else {
            const int i = e - B * per;
            io[i] = status[i]; io[B + i] = iter[i];
        }
    }
}
static int solve_host_small(nmpc_solver* s, int B, const double* x0bar, const double* yref, int nyref, const double* We,
                            double* u0, double* x1, int* status, int* qp_iter)
{
    const int nx = s->mi.nx, nu = s->mi.nu, ny = nx + nu;
    const int R = (NSTAGE + 1) * nyref;
    if (!s->h_small) {
        const size_t cap_in = (size_t)SMALL_HOST_B * (2 * nx + (size_t)(NSTAGE + 1) * ny);
        CK(cudaMallocHost(&s->h_small, cap_in * sizeof(double)));
        CK(cudaMalloc(&s->d_small, cap_in * sizeof(double)));
    }
    cudaStream_t st = s->own_stream;
    const size_t n_in = (size_t)B * (nx + R + (We ? nx : 0));
    memcpy(s->h_small, x0bar, (size_t)B * nx * sizeof(double));
    memcpy(s->h_small + (size_t)B * nx, yref, (size_t)B * R * sizeof(double));
    if (We) memcpy(s->h_small + (size_t)B * (nx + R), We, (size_t)B * nx * sizeof(double));
    CK(cudaMemcpyAsync(s->d_small, s->h_small, n_in * sizeof(double), cudaMemcpyHostToDevice, st));
    k_small_in<<<(int)((n_in + 255) / 256), 256, 0, st>>>(B, nx, R, We ? 1 : 0, s->d_small, s->d_x0bar, s->d_yref, s->d_We);
    int rc = nmpc_rti_solve_device(s, B, s->d_x0bar, s->d_yref, nyref, We ? s->d_We : nullptr, nullptr, nullptr, 0,
                                   s->d_status, s->d_iter, s->d_stats, st);
    if (rc) return rc;
    s->last_host_B = B;
    const size_t n_out = (size_t)B * (nu + nx);                       // doubles, followed by 2 B ints
    k_small_out<<<(int)((n_out + B + 255) / 256), 256, 0, st>>>(B, nx, nu, s->cap, s->d_x, s->d_u, s->d_status, s->d_iter, s->d_small);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(s->h_small, s->d_small, n_out * sizeof(double) + 2 * (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    s->last_launches += 2;
    CK(cudaStreamSynchronize(st));
    memcpy(u0, s->h_small, (size_t)B * nu * sizeof(double));
    memcpy(x1, s->h_small + (size_t)B * nu, (size_t)B * nx * sizeof(double));
    const int* io = reinterpret_cast<const int*>(s->h_small + n_out);
    memcpy(status, io, (size_t)B * sizeof(int));
    memcpy(qp_iter, io + B, (size_t)B * sizeof(int));
    return 0;
}

extern "C" int nmpc_rti_solve_host(nmpc_solver* s, int B, const double* x0bar, const double* yref, int nyref, const double* We,
                                   double* u0, double* x1, int* status, int* qp_iter)
{
    if (!s || !x0bar || !yref || !u0 || !x1 || !status || !qp_iter) return set_err(NMPC_E_ARG, "nmpc_rti_solve_host: null argument");
    if (B < 1) return set_err(NMPC_E_ARG, "B < 1");
    if (B > s->cap) return set_err(NMPC_E_CAPACITY, "batch exceeds capacity");
    const int nx = s->mi.nx, nu = s->mi.nu, ny = nx + nu;
    if (nyref != 3 && nyref != ny) return set_err(NMPC_E_ARG, "nyref must be 3 or ny");
    CK(cudaSetDevice(s->device));
    int rc = ensure_staging(s); if (rc) return rc;
    if (B <= SMALL_HOST_B) return solve_host_small(s, B, x0bar, yref, nyref, We, u0, x1, status, qp_iter);
    cudaStream_t st = s->own_stream;
    // H2D: x0bar and yref land instance-major and are transposed on the device
    {
        CK(cudaMemcpyAsync(s->d_out_aos, x0bar, (size_t)B * nx * sizeof(double), cudaMemcpyHostToDevice, st));
        dim3 g((B + 31) / 32, (nx + 31) / 32), b(32, 8);
        k_aos_to_soa<<<g, b, 0, st>>>(B, nx, s->d_out_aos, s->d_x0bar);
    }
    rc = h2d_transpose(s, B, (NSTAGE + 1) * nyref, yref, s->d_yref, B, st); if (rc) return rc;
    if (We) {
        CK(cudaMemcpyAsync(s->d_out_aos, We, (size_t)B * nx * sizeof(double), cudaMemcpyHostToDevice, st));
        dim3 g((B + 31) / 32, (nx + 31) / 32), b(32, 8);
        k_aos_to_soa<<<g, b, 0, st>>>(B, nx, s->d_out_aos, s->d_We);
    }
    rc = nmpc_rti_solve_device(s, B, s->d_x0bar, s->d_yref, nyref, We ? s->d_We : nullptr, nullptr, nullptr, 0,
                               s->d_status, s->d_iter, s->d_stats, st);
    if (rc) return rc;
    s->last_host_B = B;
    // D2H: u_0 and x_1 only (what the controller reads back)
    {
        dim3 b(32, 8);
        dim3 g1((B + 31) / 32, (nu + 31) / 32);
        k_soa_to_aos<<<g1, b, 0, st>>>(B, nu, s->cap, s->d_u, s->d_out_aos);
        dim3 g2((B + 31) / 32, (nx + 31) / 32);
        k_soa_to_aos<<<g2, b, 0, st>>>(B, nx, s->cap, s->d_x + (size_t)nx * s->cap, s->d_out_aos + (size_t)B * nu);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(u0, s->d_out_aos, (size_t)B * nu * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(x1, s->d_out_aos + (size_t)B * nu, (size_t)B * nx * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(status, s->d_status, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(qp_iter, s->d_iter, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    s->last_launches += We ? 5 : 4;
    CK(cudaStreamSynchronize(st));
    return 0;
}

// ---- batched controller tick (declared in include/nmpc_b200.h) ------------------
static int ensure_ctrl(nmpc_solver* s, cudaStream_t st)
{
    if (s->d_vref) return 0;
    CK(cudaMalloc(&s->d_vref, (size_t)s->cap * s->mi.nv * sizeof(double)));
    CK(cudaMemsetAsync(s->d_vref, 0, (size_t)s->cap * s->mi.nv * sizeof(double), st));   // ordered before the first tick on st
    return 0;
}

extern "C" int nmpc_ctrl_reset(nmpc_solver* s, void* stream)
{
    if (!s) return set_err(NMPC_E_ARG, "null solver");
    CK(cudaSetDevice(s->device));
    int rc = ensure_ctrl(s, (cudaStream_t)stream);
    if (rc) return rc;
    CK(cudaMemsetAsync(s->d_vref, 0, (size_t)s->cap * s->mi.nv * sizeof(double), (cudaStream_t)stream));
    if (s->d_roll_cmd) CK(cudaMemsetAsync(s->d_roll_cmd, 0, (size_t)s->cap * 3 * sizeof(double), (cudaStream_t)stream));   // the rollout engine's commands
    return 0;
}

extern "C" int nmpc_ctrl_state_device(nmpc_solver* s, double** d_vref, int* leading_dim)
{
    if (!s) return set_err(NMPC_E_ARG, "null solver");
    CK(cudaSetDevice(s->device));
    const bool fresh = s->d_vref == nullptr;
    int rc = ensure_ctrl(s, s->own_stream);
    if (rc) return rc;
    if (fresh) CK(cudaStreamSynchronize(s->own_stream));      // the caller may touch the buffer on any stream
    if (d_vref) *d_vref = s->d_vref;
    if (leading_dim) *leading_dim = s->cap;
    return 0;
}

// one tick for B robots on `st`: pre-processing kernel, RTI step (sqp_max_iter <= 1, the reference's behaviour) or SQP to
// convergence on the persisted iterate, post-processing kernel
static int ctrl_tick_core(nmpc_solver* s, int B, const double* d_pose, const double* d_vel, const double* d_steer,
                          const double* d_refs, const int* d_nref, int nref_max, double dt, double* d_cmd, int* d_status,
                          int* d_qp_iter, int sqp_max_iter, double sqp_tol, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_staging(s);
    if (rc) return rc;
    rc = ensure_ctrl(s, st);
    if (rc) return rc;
    rc = upload_tables(s, st);
    if (rc) return rc;
    if (!d_status) d_status = s->d_status;
    if (!d_qp_iter) d_qp_iter = s->d_iter;
    const double* p = s->d_tab + s->off_p;
    const double* W0 = s->d_tab + s->off_W;
    const double* Wt = s->d_tab + s->off_We;
    double* We = s->model == NMPC_MODEL_DIFF ? s->d_We : nullptr;      // only the diff wrapper switches W_e (Diff.cpp:126-139)
    const int nb = (B + 127) / 128;
    switch (s->model) {
        case 0: k_ctrl_pre<DiffModel><<<nb, 128, 0, st>>>(B, d_pose, d_vel, d_steer, d_refs, d_nref, nref_max, s->d_vref, s->cap, p, W0, Wt, s->d_x0bar, s->d_yref, We); break;
        case 1: k_ctrl_pre<Omni4Model><<<nb, 128, 0, st>>>(B, d_pose, d_vel, d_steer, d_refs, d_nref, nref_max, s->d_vref, s->cap, p, W0, Wt, s->d_x0bar, s->d_yref, We); break;
        default: k_ctrl_pre<TricModel><<<nb, 128, 0, st>>>(B, d_pose, d_vel, d_steer, d_refs, d_nref, nref_max, s->d_vref, s->cap, p, W0, Wt, s->d_x0bar, s->d_yref, We); break;
    }
    CK(cudaGetLastError());
    if (sqp_max_iter <= 1)
        rc = nmpc_rti_solve_device(s, B, s->d_x0bar, s->d_yref, 3, We, nullptr, nullptr, 0, d_status, d_qp_iter, nullptr, stream);
    else
        rc = nmpc_sqp_solve_device(s, B, s->d_x0bar, s->d_yref, 3, We, nullptr, nullptr, 0, sqp_max_iter, sqp_tol, d_status, nullptr,
                                   d_qp_iter, stream);
    if (rc) return rc;
    switch (s->model) {
        case 0: k_ctrl_post<DiffModel><<<nb, 128, 0, st>>>(B, d_status, s->d_x0bar, s->d_u, s->cap, dt, p, s->d_vref, s->cap, d_cmd); break;
        case 1: k_ctrl_post<Omni4Model><<<nb, 128, 0, st>>>(B, d_status, s->d_x0bar, s->d_u, s->cap, dt, p, s->d_vref, s->cap, d_cmd); break;
        default: k_ctrl_post<TricModel><<<nb, 128, 0, st>>>(B, d_status, s->d_x0bar, s->d_u, s->cap, dt, p, s->d_vref, s->cap, d_cmd); break;
    }
    CK(cudaGetLastError());
    s->last_launches += 2;
    return 0;
}

static int ctrl_tick_check(nmpc_solver* s, int B, const double* d_pose, const double* d_vel, const double* d_steer,
                           const double* d_refs, int nref_max, double dt, double* d_cmd, const char* who)
{
    static thread_local char msg[160];
    auto bad = [&](int code, const char* what) { snprintf(msg, sizeof(msg), "%s: %s", who, what); return set_err(code, msg); };
    if (!s || !d_pose || !d_vel || !d_refs || !d_cmd) return bad(NMPC_E_ARG, "null argument");
    if (B < 1) return bad(NMPC_E_ARG, "B < 1");
    if (B > s->cap) return bad(NMPC_E_CAPACITY, "batch exceeds capacity");
    if (nref_max < 1) return bad(NMPC_E_ARG, "nref_max < 1 (run() needs at least one reference pose)");
    if (s->model == NMPC_MODEL_TRIC && !d_steer) return bad(NMPC_E_ARG, "tric needs the measured steering angle");
    if (!(dt > 0.0)) return bad(NMPC_E_ARG, "dt must be positive");
    return 0;
}

extern "C" int nmpc_ctrl_tick_device(nmpc_solver* s, int B, const double* d_pose, const double* d_vel, const double* d_steer,
                                     const double* d_refs, const int* d_nref, int nref_max, double dt, double* d_cmd,
                                     int* d_status, int* d_qp_iter, void* stream)
{
    int rc = ctrl_tick_check(s, B, d_pose, d_vel, d_steer, d_refs, nref_max, dt, d_cmd, "nmpc_ctrl_tick_device");
    if (rc) return rc;
    CK(cudaSetDevice(s->device));
    return ctrl_tick_core(s, B, d_pose, d_vel, d_steer, d_refs, d_nref, nref_max, dt, d_cmd, d_status, d_qp_iter, 1, 0.0, stream);
}

extern "C" int nmpc_ctrl_tick_sqp_device(nmpc_solver* s, int B, const double* d_pose, const double* d_vel, const double* d_steer,
                                         const double* d_refs, const int* d_nref, int nref_max, double dt, int sqp_max_iter,
                                         double sqp_tol, double* d_cmd, int* d_status, int* d_qp_iter, void* stream)
{
    int rc = ctrl_tick_check(s, B, d_pose, d_vel, d_steer, d_refs, nref_max, dt, d_cmd, "nmpc_ctrl_tick_sqp_device");
    if (rc) return rc;
    if (sqp_max_iter < 1 || sqp_max_iter > 1000 || !(sqp_tol >= 0.0))
        return set_err(NMPC_E_ARG, "nmpc_ctrl_tick_sqp_device: sqp_max_iter must be 1..1000 and sqp_tol >= 0");
    CK(cudaSetDevice(s->device));
    return ctrl_tick_core(s, B, d_pose, d_vel, d_steer, d_refs, d_nref, nref_max, dt, d_cmd, d_status, d_qp_iter, sqp_max_iter, sqp_tol, stream);
}

// host-buffer form of the tick: instance-major arrays, copied and transposed on the device like nmpc_rti_solve_host
extern "C" int nmpc_ctrl_tick_host(nmpc_solver* s, int B, const double* pose, const double* vel, const double* steer,
                                   const double* refs, const int* nref, int nref_max, double dt, double* cmd, int* status,
                                   int* qp_iter)
{
    if (!s || !pose || !vel || !refs || !cmd || !status) return set_err(NMPC_E_ARG, "nmpc_ctrl_tick_host: null argument");
    if (B < 1) return set_err(NMPC_E_ARG, "nmpc_ctrl_tick_host: B < 1");
    if (B > s->cap) return set_err(NMPC_E_CAPACITY, "nmpc_ctrl_tick_host: batch exceeds capacity");
    if (nref_max < 1 || nref_max > NSTAGE + 1) return set_err(NMPC_E_ARG, "nmpc_ctrl_tick_host: nref_max must be 1..N+1");
    if (s->model == NMPC_MODEL_TRIC && !steer) return set_err(NMPC_E_ARG, "nmpc_ctrl_tick_host: tric needs the measured steering angle");
    CK(cudaSetDevice(s->device));
    int rc = ensure_staging(s); if (rc) return rc;
    const size_t cap = s->cap;
    if (!s->d_cin) {
        CK(cudaMalloc(&s->d_cin, cap * (10 + 3 * (NSTAGE + 1)) * sizeof(double)));
        CK(cudaMalloc(&s->d_cnref, cap * sizeof(int)));
    }
    cudaStream_t st = s->own_stream;
    double *d_pose = s->d_cin, *d_vel = d_pose + 3 * cap, *d_steer = d_vel + 3 * cap, *d_cmd = d_steer + cap, *d_refs = d_cmd + 3 * cap;
    rc = h2d_transpose(s, B, 3, pose, d_pose, B, st); if (rc) return rc;
    rc = h2d_transpose(s, B, 3, vel, d_vel, B, st); if (rc) return rc;
    if (steer) CK(cudaMemcpyAsync(d_steer, steer, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = h2d_transpose(s, B, 3 * nref_max, refs, d_refs, B, st); if (rc) return rc;
    if (nref) CK(cudaMemcpyAsync(s->d_cnref, nref, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, st));
    // the command of an instance whose solve fails is left as the caller passed it (run() throws before writing it)
    rc = h2d_transpose(s, B, 3, cmd, d_cmd, B, st); if (rc) return rc;
    rc = nmpc_ctrl_tick_device(s, B, d_pose, d_vel, steer ? d_steer : nullptr, d_refs, nref ? s->d_cnref : nullptr, nref_max, dt,
                               d_cmd, s->d_status, s->d_iter, st);
    if (rc) return rc;
    dim3 b(32, 8), g((B + 31) / 32, 1);
    k_soa_to_aos<<<g, b, 0, st>>>(B, 3, B, d_cmd, s->d_out_aos);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(cmd, s->d_out_aos, (size_t)B * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(status, s->d_status, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (qp_iter) CK(cudaMemcpyAsync(qp_iter, s->d_iter, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    s->last_launches += 5;
    CK(cudaStreamSynchronize(st));
    return 0;
}

// ---- batched path discretisation (declared in include/nmpc_b200.h) --------------
extern "C" int nmpc_path_discretize_device(int device, int B, const double* d_segments, const int* d_path_offsets, int n_paths,
                                           const int* d_path_id, const double* d_nearest_u, double sample_period, int num_poses,
                                           int is_holonomic, double* d_poses, void* stream)
{
    if (!d_segments || !d_path_offsets || !d_nearest_u || !d_poses) return set_err(NMPC_E_ARG, "nmpc_path_discretize_device: null argument");
    if (B < 1 || n_paths < 1 || num_poses < 1) return set_err(NMPC_E_ARG, "nmpc_path_discretize_device: B, n_paths and num_poses must be >= 1");
    if (!(sample_period > 0.0)) return set_err(NMPC_E_ARG, "nmpc_path_discretize_device: sample_period must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return set_err(NMPC_E_NODEVICE, "no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return set_err(NMPC_E_ARG, "nmpc_path_discretize_device: bad device");
    CK(cudaSetDevice(device));
    k_path_discretize<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(B, d_segments, d_path_offsets, n_paths, d_path_id, d_nearest_u,
                                                                     sample_period, num_poses, is_holonomic, d_poses);
    CK(cudaGetLastError());
    return 0;
}

// ---- closed-loop pieces (declared in include/nmpc_b200.h) ------------------------
extern "C" int nmpc_plant_step_device(nmpc_solver* s, int B, double dt, const double* d_noise, double* d_xplant, double* d_pose,
                                      double* d_vel, double* d_steer, void* stream)
{
    if (!s || !d_xplant || !d_pose || !d_vel) return set_err(NMPC_E_ARG, "nmpc_plant_step_device: null argument");
    if (B < 1) return set_err(NMPC_E_ARG, "nmpc_plant_step_device: B < 1");
    if (B > s->cap) return set_err(NMPC_E_CAPACITY, "nmpc_plant_step_device: batch exceeds capacity");
    if (!(dt > 0.0)) return set_err(NMPC_E_ARG, "nmpc_plant_step_device: dt must be positive");
    CK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = upload_tables(s, st);
    if (rc) return rc;
    const double* p = s->d_tab + s->off_p;
    const int nb = (B + 127) / 128;
    switch (s->model) {
        case 0: k_plant_step<DiffModel><<<nb, 128, 0, st>>>(B, d_xplant, s->d_u, s->cap, d_noise, p, dt, d_pose, d_vel, d_steer); break;
        case 1: k_plant_step<Omni4Model><<<nb, 128, 0, st>>>(B, d_xplant, s->d_u, s->cap, d_noise, p, dt, d_pose, d_vel, d_steer); break;
        default: k_plant_step<TricModel><<<nb, 128, 0, st>>>(B, d_xplant, s->d_u, s->cap, d_noise, p, dt, d_pose, d_vel, d_steer); break;
    }
    CK(cudaGetLastError());
    return 0;
}

extern "C" int nmpc_path_nearest_device(int device, int B, const double* d_segments, const int* d_path_offsets, int n_paths,
                                        const int* d_path_id, const double* d_pose, double back, double ahead, double* d_u,
                                        void* stream)
{
    if (!d_segments || !d_path_offsets || !d_pose || !d_u) return set_err(NMPC_E_ARG, "nmpc_path_nearest_device: null argument");
    if (B < 1 || n_paths < 1) return set_err(NMPC_E_ARG, "nmpc_path_nearest_device: B and n_paths must be >= 1");
    if (!(back >= 0.0) || !(ahead >= 0.0)) return set_err(NMPC_E_ARG, "nmpc_path_nearest_device: the search window must be non-negative");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return set_err(NMPC_E_NODEVICE, "no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return set_err(NMPC_E_ARG, "nmpc_path_nearest_device: bad device");
    CK(cudaSetDevice(device));
    k_path_nearest<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(B, d_segments, d_path_offsets, n_paths, d_path_id, d_pose, back, ahead, d_u);
    CK(cudaGetLastError());
    return 0;
}

// ---- the closed-loop rollout engine ------------------------------------------------------------------
__global__ void k_count_nonzero(int B, const int* __restrict__ a, int* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned m = __ballot_sync(0xffffffffu, i < B && a[i] != 0);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, __popc(m));
}

extern "C" int nmpc_rollout_device(nmpc_solver* s, int B, int ticks, const nmpc_rollout_opts* o, const double* d_segments,
                                   const int* d_path_offsets, int n_paths, const int* d_path_id, double* d_u, double* d_xplant,
                                   double* d_pose, double* d_vel, double* d_steer, const double* d_noise, double* d_traj,
                                   double* d_cmds, int* d_nfail, void* stream)
{
    if (!s || !o || !d_segments || !d_path_offsets || !d_u || !d_xplant || !d_pose || !d_vel)
        return set_err(NMPC_E_ARG, "nmpc_rollout_device: null argument");
    if (B < 1 || ticks < 1 || n_paths < 1) return set_err(NMPC_E_ARG, "nmpc_rollout_device: B, ticks and n_paths must be >= 1");
    if (B > s->cap) return set_err(NMPC_E_CAPACITY, "nmpc_rollout_device: batch exceeds capacity");
    if (!(o->dt > 0.0) || !(o->back >= 0.0) || !(o->ahead >= 0.0)) return set_err(NMPC_E_ARG, "nmpc_rollout_device: dt > 0 and a non-negative search window are required");
    if (o->sqp_max_iter < 1 || o->sqp_max_iter > 1000 || !(o->sqp_tol >= 0.0)) return set_err(NMPC_E_ARG, "nmpc_rollout_device: sqp_max_iter must be 1..1000 and sqp_tol >= 0");
    if (s->model == NMPC_MODEL_TRIC && !d_steer) return set_err(NMPC_E_ARG, "nmpc_rollout_device: tric needs the steering-angle buffer");
    CK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (!s->d_roll_refs) {
        CK(cudaMalloc(&s->d_roll_refs, (size_t)s->cap * 3 * (NSTAGE + 1) * sizeof(double)));
        CK(cudaMalloc(&s->d_roll_cmd, (size_t)s->cap * 3 * sizeof(double)));
        CK(cudaMemsetAsync(s->d_roll_cmd, 0, (size_t)s->cap * 3 * sizeof(double), st));
    }
    const size_t row3 = (size_t)3 * B * sizeof(double);
    if (d_traj) CK(cudaMemcpyAsync(d_traj, d_pose, row3, cudaMemcpyDeviceToDevice, st));
    if (d_nfail) CK(cudaMemsetAsync(d_nfail, 0, (size_t)ticks * sizeof(int), st));
    int launches = 0;
    for (int t = 0; t < ticks; t++) {
        // the order of NMPCNavControlROS::processFollowPath (NMPCNavControlROS.cpp:648-698): nearest point, N+1 reference
        // poses, run(); then the nominal plant (scripts/test_scripts/acados_sim_diff.py:136-160) and the optional shift
        int rc = nmpc_path_nearest_device(s->device, B, d_segments, d_path_offsets, n_paths, d_path_id, d_pose, o->back, o->ahead, d_u, stream);
        if (rc) return rc;
        rc = nmpc_path_discretize_device(s->device, B, d_segments, d_path_offsets, n_paths, d_path_id, d_u, o->dt, NSTAGE + 1,
                                         o->is_holonomic, s->d_roll_refs, stream);
        if (rc) return rc;
        rc = ctrl_tick_core(s, B, d_pose, d_vel, d_steer, s->d_roll_refs, nullptr, NSTAGE + 1, o->dt, s->d_roll_cmd, s->d_status, s->d_iter,
                            o->sqp_max_iter, o->sqp_tol, stream);
        if (rc) return rc;
        launches += s->last_launches + 3;
        if (d_cmds) CK(cudaMemcpyAsync(d_cmds + (size_t)t * 3 * B, s->d_roll_cmd, row3, cudaMemcpyDeviceToDevice, st));
        if (d_nfail) k_count_nonzero<<<(B + 255) / 256, 256, 0, st>>>(B, s->d_status, d_nfail + t);
        rc = nmpc_plant_step_device(s, B, o->dt, d_noise ? d_noise + (size_t)t * s->mi.nu * B : nullptr, d_xplant, d_pose, d_vel, d_steer, stream);
        if (rc) return rc;
        if (d_traj) CK(cudaMemcpyAsync(d_traj + (size_t)(t + 1) * 3 * B, d_pose, row3, cudaMemcpyDeviceToDevice, st));
        if (o->shift) {
            rc = nmpc_shift_device(s, B, nullptr, nullptr, 0, nullptr, stream);
            if (rc) return rc;
            launches += 2;
        }
    }
    CK(cudaGetLastError());
    s->last_launches = launches;
    return 0;
}

extern "C" int nmpc_last_stats_host(nmpc_solver* s, int B, double* stats)
{
    if (!s || !stats) return set_err(NMPC_E_ARG, "null argument");
    if (!s->d_stats || B < 1 || B != s->last_host_B) return set_err(NMPC_E_ARG, "nmpc_last_stats_host: no host solve of this batch size yet");
    CK(cudaSetDevice(s->device));
    CK(cudaMemcpy(stats, s->d_stats, (size_t)B * 8 * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int nmpc_last_timing(nmpc_solver* s, double* ms4)
{
    if (!s || !ms4) return set_err(NMPC_E_ARG, "null argument");
    CK(cudaSetDevice(s->device));
    CK(cudaEventSynchronize(s->ev_total[1]));
    double acc[3] = {0, 0, 0};
    for (int c = 0; c < s->last_chunks; c++) {
        cudaEvent_t* ev = &s->ev[(size_t)c * 4];
        for (int q = 0; q < 3; q++) { float ms = 0; CK(cudaEventElapsedTime(&ms, ev[q], ev[q + 1])); acc[q] += ms; }
    }
    float tot = 0;
    CK(cudaEventElapsedTime(&tot, s->ev_total[0], s->ev_total[1]));
    ms4[0] = acc[0]; ms4[1] = acc[1]; ms4[2] = acc[2]; ms4[3] = tot;
    return 0;
}

extern "C" int nmpc_last_launches(const nmpc_solver* s) { return s ? s->last_launches : 0; }

extern "C" double nmpc_dfma_peak_tflops(int device, int iters)
{
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    double* d = nullptr;
    if (cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return -1.0;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_dfma<<<blocks, threads>>>(iters / 10 + 1, d);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(a);
        k_dfma<<<blocks, threads>>>(iters, d);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    const double flops = (double)blocks * threads * (double)iters * 8.0 * 2.0;
    return flops / (best * 1e-3) / 1e12;
}
