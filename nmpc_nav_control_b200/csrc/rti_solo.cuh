// K3 "solo" mapping: ONE thread block = ONE instance, the whole interior-point state in shared memory.
//
// Third machine mapping of the same iteration path as rti_core.cuh (per-lane sweeps) and rti_coop.cuh (lane-cooperative
// persistent kernel); it replaces what the reference reaches through `{m}_acados_solve(capsule)`
// (src/nmpc_nav_control/NMPCNavControlDiff.cpp:142, Tric.cpp:146) for the case the reference actually runs: ONE robot,
// one solve per 25 ms tick, latency is all that matters.  The other two mappings walk the horizon stage by stage with a
// few lanes per instance: 2,300 dependent stage steps of ~400 instructions each, 2.4 ms for a lone warp whatever the
// clock.  Here the iteration is re-cut so that almost nothing is sequential:
//
//  * everything that is local to a stage (constraint updates, residuals, gradients, ratio test, complementarity sums,
//    feed-forward terms, the closed-loop matrix) runs with one THREAD PER STAGE, all 81 stages at once;
//  * what couples the stages is six recursions per iteration.  Five of them are LINEAR with matrices that are known before
//    the recursion starts - the multiplier step  r_k = c_k + A_k' r_{k+1},  the cost-to-go gradient and the delta adjoint
//    v_k = Phi_k' (v_{k+1} + a_k) + psi_k,  and the two forward passes  dx_{k+1} = Phi_k dx_k + phi_k  with the closed-loop
//    matrix Phi_k = A_k - B_k K_k formed per stage in parallel - so one stage of them is a 7 x 7 matrix-vector product by
//    seven lanes: ~100 cycles instead of ~1,500;
//  * only the Riccati matrix recursion P_k = Q_k + A' P A - K' K is sequential AND non-linear; one warp walks it with the
//    lanes spread over the entries of P [B A], [B A]' P [B A] and P.
//
// [B A] of a stage is kept as a DENSE NX x NZ matrix J (z = [u; x]), so every product is a plain dot product with no
// per-lane branching on the column type.  Numerics: same Newton systems as the other mappings, summation orders differ
// (~1e-16 relative); tests compare all three with the oracle.
//
// Host emulation: every phase is a loop over work items strided by the thread count; the host build runs them with one
// "thread" (SOLO_TID = 0, SOLO_NT = 1), which executes the items of a phase in increasing order.  A phase never reads
// what another item of the same phase writes, except where stated (neighbour reads are split off into a read block that
// ends with a barrier).
#pragma once
#include "rti_records.cuh"

#if defined(__CUDA_ARCH__)
#define SOLO_HOST 0
#define SOLO_TID ((int)threadIdx.x)
#define SOLO_NT ((int)blockDim.x)
#define SOLO_SYNC() __syncthreads()
#define SOLO_WSYNC() __syncwarp()
#else
#define SOLO_HOST 1
#define SOLO_TID 0
#define SOLO_NT 1
#define SOLO_SYNC() ((void)0)
#define SOLO_WSYNC() ((void)0)
#endif

// optional phase timing of block 0 (build with -DNMPC_SOLO_PROF; read with nmpc_solo_prof): cycles per phase, summed over the
// solve in shared-memory slots and written out at the end
#define SOLO_NPROF 24
#if defined(__CUDACC__) && defined(NMPC_SOLO_PROF)
__device__ unsigned long long g_solo_prof[SOLO_NPROF];
#endif
#if defined(__CUDA_ARCH__) && defined(NMPC_SOLO_PROF)
#define SOLO_T(id) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long c_ = clock64(); long long* l_ = reinterpret_cast<long long*>(sm + O_SC + 2); \
                        l_[1 + (id)] += c_ - l_[0]; l_[0] = c_; } } while (0)
#define SOLO_T0() do { if (threadIdx.x == 0 && blockIdx.x == 0) { long long* l_ = reinterpret_cast<long long*>(sm + O_SC + 2); l_[0] = clock64(); \
                       for (int q_ = 0; q_ < SOLO_NPROF; q_++) l_[1 + q_] = 0; } } while (0)
#define SOLO_TEND() do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long* l_ = reinterpret_cast<const long long*>(sm + O_SC + 2); \
                         for (int q_ = 0; q_ < SOLO_NPROF; q_++) g_solo_prof[q_] += (unsigned long long)l_[1 + q_]; } } while (0)
#else
#define SOLO_T(id) ((void)0)
#define SOLO_T0() ((void)0)
#define SOLO_TEND() ((void)0)
#endif

namespace nmpc {

template <class M>
struct Solo {
    using S = Rti<M>;
    using R = typename S::R;
    using LaneCtl = typename S::LaneCtl;
    static constexpr int NV = S::NV, NX = S::NX, NU = S::NU, NZ = S::NZ, NC = S::NC, NB2 = S::NB2, NLU = S::NLU, NY = S::NY;
    static constexpr int N = NSTAGE;
    static constexpr int NTRI = NZ * (NZ + 1) / 2, NXTRI = NX * (NX + 1) / 2;
    // one thread per stage in the stage-parallel phases; the models with four channels have 176 entries of P [B A] per stage:
    // one per thread of six warps instead of two per thread of three (Riccati sub-phases 964 -> ~600 cycles)
    static constexpr int THREADS = NV > 2 ? 192 : ((N + 1 + 31) / 32) * 32;

    // ---- shared-memory block of one stage (doubles) --------------------------------------------------------------------
    // Models with more than two channels (omni4: 81 stages x 559 doubles = 362 KB) do not fit the shared memory of an SM: their
    // two largest per-stage arrays, the dense [B A] and the closed-loop matrix with its transpose, live in a per-instance
    // scratch area in global memory instead (L2 resident, 278 KB); the sequential loops fetch the next stage's rows into
    // registers while they work on the current one.
    static constexpr bool BIG = NV > 2;
    static constexpr int O_J = 0;                      // NX*NZ  dense [B A], row-major, columns in z order [u; x]   (not BIG)
    static constexpr int O_DLB = O_J + (BIG ? 0 : NX * NZ);   // NB2    lb - z for [u; ref]
    static constexpr int O_DUB = O_DLB + NB2;          // NB2
    static constexpr int O_Q = O_DUB + NB2;            // NZ     QP gradient
    static constexpr int O_B0 = O_Q + NZ;              // NX     phi(x,u) - x_next
    static constexpr int O_T = O_B0 + NX;              // 2*NB2  slacks, lower then upper
    static constexpr int O_LAM = O_T + 2 * NB2;        // 2*NB2
    static constexpr int O_Z = O_LAM + 2 * NB2;        // NZ
    static constexpr int O_PI = O_Z + NZ;              // NX     multiplier of the dynamics that define x_k
    static constexpr int O_DZ = O_PI + NX;             // NZ     final step
    static constexpr int O_DZA = O_DZ + NZ;            // NZ     predictor step
    static constexpr int O_MC = O_DZA + NZ;            // 2*NB2  dt_aff * dlam_aff
    static constexpr int O_LUU = O_MC + 2 * NB2;       // NLU    Cholesky factor of the control block, diagonal inverted
    static constexpr int O_KF = O_LUU + NLU;           // NV*NX  K = Luu^-1 S after the factorisation, then the gain Luu^-T K
    static constexpr int O_UF = O_KF + NV * NX;        // NV     feed-forward control of the running solve
    static constexpr int O_RB = O_UF + NV;             // NX     dynamics residual
    static constexpr int O_G = O_RB + NX;              // NZ     stage gradient incl. barrier terms; controls part reused by the delta solve
    static constexpr int O_DG = O_G + NZ;              // NZ     diagonal of the stage Hessian incl. barrier terms and regularisation
    static constexpr int O_RED = O_DG + NZ;            // NRED   per-stage partial results of the reductions
    static constexpr int NRED = 6;
    static constexpr int PS = (O_RED + NRED) | 1;      // odd stride: the threads of a warp (consecutive stages) hit different banks
    // ---- what the linear recursions touch, per stage, in a block of its own: rows padded to 16 bytes (NXP doubles, padding
    //      zero) so that a lane fetches its matrix row and the running vector with 16-byte loads; Phi is kept both ways (rows
    //      for the forward passes, rows of the transpose for the backward ones) --------------------------------------------------
    static constexpr int NXP = (NX + 1) & ~1;
    static constexpr int R_PHI = 0;                    // NX*NXP closed-loop matrix A - B K, row-major
    static constexpr int R_PHT = R_PHI + (BIG ? 0 : NX * NXP);     // NX*NXP its transpose
    static constexpr int R_V = R_PHT + (BIG ? 0 : NX * NXP);       // NXP    running backward recursion (multiplier step / gradient / delta adjoint)
    static constexpr int R_X = R_V + NXP;              // NXP    running forward recursion (state step)
    static constexpr int R_C = R_X + NXP;              // NXP    constant term of the forward recursion
    static constexpr int RS0 = R_C + NXP;
    static constexpr int RS = RS0 % 4 == 2 ? RS0 : RS0 + 2;      // even (alignment), not a multiple of four (bank spread of the per-stage threads)
    // ---- common area ------------------------------------------------------------------------------------------------------
    static constexpr int O_P = ((N + 1) * PS + 1) & ~1;  // NX*NXP cost-to-go of the successor stage, rows 16-byte aligned
    static constexpr int O_GM = O_P + NX * NXP;        // NX*NZ  P [B A]
    static constexpr int O_MM = O_GM + NX * NZ;        // NZ*NZ  [B A]' P [B A] + D (lower triangle used)
    static constexpr int O_WV = O_MM + NZ * NZ;        // NX     P rb + p of the successor stage
    static constexpr int O_GP = O_WV + NX;             // NZ     stage gradient incl. the cost-to-go of the successor
    static constexpr int O_LH = O_GP + NZ;             // NV     Luu^-1 (gradient of the controls)
    static constexpr int O_JB = (O_LH + NV + 1) & ~1;  // 2*JSZ  BIG models: double buffer of the stage's [B A] in the Riccati recursion
    static constexpr int O_CTL = O_JB + (NV > 2 ? 2 * (((3 + 2 * NV) * (3 + 3 * NV) + 1) & ~1) : 0);            // LaneCtl
    static constexpr int CTL_D = (int)((sizeof(LaneCtl) + 7) / 8);
    static constexpr int O_SC = O_CTL + CTL_D;         // a_step
    static constexpr int O_REC = (O_SC + 3 + SOLO_NPROF + 2 + 1) & ~1;
    static constexpr int SM_DOUBLES = O_REC + (N + 1) * RS;
    static_assert(O_P % 2 == 0 && O_REC % 2 == 0 && RS % 2 == 0 && NXP % 2 == 0, "16-byte loads of padded rows");
    NMPC_HD static double* rec(double* sm, int k) { return sm + O_REC + (size_t)k * RS; }
    // per-instance global scratch of the BIG models: per stage [J | Phi | Phi']
    static constexpr int JSZ = (NX * NZ + 1) & ~1;                // rows of Phi stay 16-byte aligned
    static constexpr int GS = JSZ + 2 * NX * NXP;
    static constexpr size_t GSCR_DOUBLES = BIG ? (size_t)(N + 1) * GS : 0;
    NMPC_HD static double*& gscr(double* sm) { return *reinterpret_cast<double**>(sm + O_SC + 1); }
    NMPC_HD static double* Jp(double* sm, int k) { return BIG ? gscr(sm) + (size_t)k * GS : sm + (size_t)k * PS + O_J; }
    NMPC_HD static double* phi(double* sm, int k) { return BIG ? gscr(sm) + (size_t)k * GS + JSZ : rec(sm, k) + R_PHI; }
    NMPC_HD static double* pht(double* sm, int k) { return BIG ? gscr(sm) + (size_t)k * GS + JSZ + NX * NXP : rec(sm, k) + R_PHT; }
    // s0 + row . v with the row in shared memory (16-byte aligned, padded) and v in registers
    NMPC_HD static double dotp_r(const double* row, const double (&v)[NX], double s0)
    {
        double s1 = 0.0;
#if SOLO_HOST
        for (int m = 0; m < NX; m++) { if (m & 1) s1 += row[m] * v[m]; else s0 += row[m] * v[m]; }
#else
        const double2* A = reinterpret_cast<const double2*>(row);
#pragma unroll
        for (int q = 0; q < NXP / 2; q++) {
            const double2 x = A[q];
            s0 += x.x * v[2 * q];
            if (2 * q + 1 < NX) s1 += x.y * v[2 * q + 1];
        }
#endif
        return s0 + s1;
    }
    // s0 + a . b over a padded row (both 16-byte aligned); even terms in one accumulator, odd terms in the other
    NMPC_HD static double dotp(const double* a, const double* b, double s0)
    {
        double s1 = 0.0;
#if SOLO_HOST
        for (int m = 0; m < NXP; m++) { if (m & 1) s1 += a[m] * b[m]; else s0 += a[m] * b[m]; }
#else
        const double2* A = reinterpret_cast<const double2*>(a);
        const double2* B = reinterpret_cast<const double2*>(b);
#pragma unroll
        for (int q = 0; q < NXP / 2; q++) { const double2 x = A[q], y = B[q]; s0 += x.x * y.x; s1 += x.y * y.y; }
#endif
        return s0 + s1;
    }
    static constexpr size_t SM_BYTES = (size_t)SM_DOUBLES * sizeof(double);

    NMPC_HD static int zidx(int b) { return b < NV ? b : NU + 3 + b; }       // z index of bounded component b ([u; ref])
    NMPC_HD static double cold_z0(double dl, double du_, double thr0)
    {
        double z0 = 0.0;
        const double t_l = -dl, t_u = du_;
        if (t_l < thr0) { if (t_u < thr0) z0 = 0.5 * (dl + du_); else z0 = dl + thr0; }
        else if (t_u < thr0) z0 = du_ - thr0;
        return z0;
    }
    // u = -Muu^-1 v with Muu = Luu Luu' (row-packed lower, inverted diagonal)
    NMPC_HD static void neg_solve_uu(const double* Luu, const double* v, double* u)
    {
        double lh[NV];
#pragma unroll
        for (int a = 0; a < NV; a++) {
            double s = v[a];
#pragma unroll
            for (int c = 0; c < a; c++) s -= Luu[a * (a + 1) / 2 + c] * lh[c];
            lh[a] = s * Luu[a * (a + 1) / 2 + a];
        }
#pragma unroll
        for (int a = NV - 1; a >= 0; a--) {
            double s = -lh[a];
#pragma unroll
            for (int b = a + 1; b < NV; b++) s -= Luu[b * (b + 1) / 2 + a] * u[b];
            u[a] = s * Luu[a * (a + 1) / 2 + a];
        }
    }

    // ---- where the QP of the instance comes from and where its solution goes: the tile layout of the lockstep sweeps (K1 / K2 =
    //      k_linearize, K4 = k_step) or the per-instance records of the lane-cooperative kernel (k_linearize_g, k_step_g; the SQP
    //      passes with their active-instance queue) ---------------------------------------------------------------------------------
    struct TileIO {
        double* lane;                                   // lane-resolved base of the instance's tile
        static constexpr int ER = R::ER;
        NMPC_HD const double* lin(int k) const { return lane + R::OFF_LIN + (size_t)k * R::NF_LIN * LANES; }
        NMPC_HD double e(int k, int i, int c) const { return lin(k)[(R::E + i * NC + c) * LANES]; }
        NMPC_HD double b0(int k, int i) const { return lin(k)[(R::B0 + i) * LANES]; }
        NMPC_HD double dlb(int k, int b) const { return lin(k)[(R::DLB + b) * LANES]; }
        NMPC_HD double dub(int k, int b) const { return lin(k)[(R::DUB + b) * LANES]; }
        NMPC_HD double q(int k, int w) const { return lin(k)[(R::Q + w) * LANES]; }
        NMPC_HD double z0(int j) const { return lane[R::OFF_IT + (R::Z + NU + j) * LANES]; }
        NMPC_HD void put_z(int k, int w, double v) const { lane[R::OFF_IT + ((size_t)k * R::NF_IT + R::Z + w) * LANES] = v; }
    };
    struct RecIO {
        using GR = GRec<NV>;
        double* base;                                   // the instance's records, stage-major
        static constexpr int ER = 3;
        NMPC_HD double e(int k, int i, int c) const { return base[(size_t)k * GR::NREC + GR::E + i * NC + c]; }
        NMPC_HD double b0(int k, int i) const { return base[(size_t)k * GR::NREC + GR::B0 + i]; }
        NMPC_HD double dlb(int k, int b) const { return base[(size_t)k * GR::NREC + GR::DLB + b]; }
        NMPC_HD double dub(int k, int b) const { return base[(size_t)k * GR::NREC + GR::DUB + b]; }
        NMPC_HD double q(int k, int w) const { return base[(size_t)k * GR::NREC + GR::Q + w]; }
        NMPC_HD double z0(int j) const { return base[GR::Z + NU + j]; }
        NMPC_HD void put_z(int k, int w, double v) const { base[(size_t)k * GR::NREC + GR::Z + w] = v; }
    };

    // ---- set-up: the QP of the instance from its tile (written by K1 / K2) into the stage blocks -----------------------------
    template <class IO>
    NMPC_HD static void load_qp(double* sm, const IO& io, const Tables& tb)
    {
        for (int e = SOLO_TID; e < (N + 1) * RS; e += SOLO_NT) sm[O_REC + e] = 0.0;      // incl. the row padding
        if (BIG) {
            for (int e = SOLO_TID; e < (N + 1) * 2 * NX * NXP; e += SOLO_NT)
                gscr(sm)[(size_t)(e / (2 * NX * NXP)) * GS + JSZ + e % (2 * NX * NXP)] = 0.0;
        }
        for (int k = SOLO_TID; k <= N; k += SOLO_NT) {
            double* st = sm + (size_t)k * PS;
            if (k < N) {
                const double* lti = tb.lti + k * 4 * NV;
                const double* thr = tb.thr + k * NC;
                double* J = Jp(sm, k);
                for (int e = 0; e < NX * NZ; e++) J[e] = 0.0;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    auto E = [&](int c) { return i < IO::ER ? io.e(k, i, c) : thr[c]; };
                    if (i < 2) J[i * NZ + NU + i] = 1.0;
                    J[i * NZ + NU + 2] = E(0);
#pragma unroll
                    for (int c = 0; c < NV; c++) {
                        J[i * NZ + c] = E(1 + 2 * NV + c);
                        J[i * NZ + NU + 3 + c] = E(1 + c);
                        J[i * NZ + NU + 3 + NV + c] = E(1 + NV + c);
                    }
                }
#pragma unroll
                for (int c = 0; c < NV; c++) {
                    J[(3 + c) * NZ + c] = lti[2 * NV + c];                  // au
                    J[(3 + c) * NZ + NU + 3 + c] = lti[c];                  // av
                    J[(3 + c) * NZ + NU + 3 + NV + c] = lti[NV + c];        // ar
                    J[(3 + NV + c) * NZ + c] = lti[3 * NV + c];             // ru
                    J[(3 + NV + c) * NZ + NU + 3 + NV + c] = 1.0;
                }
#pragma unroll
                for (int i = 0; i < NX; i++) st[O_B0 + i] = io.b0(k, i);
            }
#pragma unroll
            for (int b = 0; b < NB2; b++) {
                const bool act = (b < NV) ? (k < N) : (k > 0);
                st[O_DLB + b] = act ? io.dlb(k, b) : 0.0;
                st[O_DUB + b] = act ? io.dub(k, b) : 0.0;
            }
#pragma unroll
            for (int w = 0; w < NZ; w++) st[O_Q + w] = io.q(k, w);
            if (k == 0) {
#pragma unroll
                for (int j = 0; j < NX; j++) st[O_Z + NU + j] = io.z0(j);    // x0 elimination: the constant stage-0 state
            }
        }
        SOLO_SYNC();
    }

    // ---- B, step 1: everything of the multiplier step r_k that does not come from r_{k+1} ---------------------------------------
    NMPC_HD static void phase_adjoint_const(double* sm, const Tables& tb, const double* We, size_t wst, const IpmOpts& o,
                                            double sigmu, double mcw)
    {
        for (int k = SOLO_TID; k <= N; k += SOLO_NT) {
            if (k == 0) continue;
            double* st = sm + (size_t)k * PS;
            const bool hasU = k < N;
            double c[NX];
#pragma unroll
            for (int j = 0; j < NX; j++) {
                const double H = hasU ? tb.dt * tb.W[k * NY + j] : We[(size_t)j * wst];
                c[j] = st[O_Q + NU + j] + H * st[O_Z + NU + j] - st[O_PI + j] + H * st[O_DZ + NU + j];
            }
            if (hasU) {
                const double* J = Jp(sm, k);
                const double* pio = st + PS + O_PI;
#pragma unroll
                for (int j = 0; j < NX; j++) {
                    double s = 0.0;
#pragma unroll
                    for (int i = 0; i < NX; i++) s += J[i * NZ + NU + j] * pio[i];
                    c[j] += s;
                }
            }
#pragma unroll
            for (int q = 0; q < NV; q++) {
                const int b = NV + q, w = zidx(b);
                const double dl = st[O_DLB + b], du_ = st[O_DUB + b], zb = st[O_Z + w], dzb = st[O_DZ + w];
                const double ll = st[O_LAM + b], lu = st[O_LAM + NB2 + b], tl = st[O_T + b], tu = st[O_T + NB2 + b];
                const double rd_l = dl - zb + tl, rd_u = -du_ + zb + tu;
                const double rm_l = ll * tl - o.tau_min + mcw * st[O_MC + b] - sigmu;
                const double rm_u = lu * tu - o.tau_min + mcw * st[O_MC + NB2 + b] - sigmu;
                const double dt_l = dzb - rd_l, dt_u = -dzb - rd_u;
                const double dl_l = -(ll * dt_l + rm_l) / tl, dl_u = -(lu * dt_u + rm_u) / tu;
                c[3 + NV + q] += (lu - ll) - (dl_l - dl_u);
            }
#pragma unroll
            for (int j = 0; j < NX; j++) rec(sm, k)[R_V + j] = c[j];
        }
        SOLO_SYNC();
    }

    // ---- the linear recursions, by one warp.  Backward, V_k += Mat_k' V_{k+1} for k = N-1 .. 1: with A_k (the multiplier step; columns
    //      of the dense [B A]) and with Phi_k (cost-to-go gradient of the delta solve; rows of the transpose).  Forward,
    //      X_{k+1} = Phi_k X_k + C_k with X_0 = 0.  On the device lane j keeps component j of the running vector in a register and
    //      the others' components arrive by shuffle: the chain of a stage is one shuffle round and four dependent fused
    //      multiply-adds (even and odd terms accumulate separately), the matrix row is fetched ahead of it.  The host emulation
    //      passes the vector through the stage blocks instead; the sums are formed in the same order. ---------------------------------
#if !SOLO_HOST
    // s0 + row . v, v spread over the lanes (component i on lane i); the row is fetched TWO stages ahead of its use (the
    // recursions below keep two row buffers in registers): shared memory for the small models, L2 for the BIG ones
    template <int STRIDE>
    __device__ __forceinline__ static void load_row(double (&r)[NX], const double* row)
    {
#pragma unroll
        for (int i = 0; i < NX; i++) r[i] = row[i * STRIDE];
    }
    __device__ __forceinline__ static double dot_shfl(const double (&r)[NX], double v, double s0)
    {
        double s1 = 0.0;
#pragma unroll
        for (int i = 0; i < NX; i++) {
            const double vi = __shfl_sync(0xffffffffu, v, i);
            if (i & 1) s1 += r[i] * vi; else s0 += r[i] * vi;
        }
        return s0 + s1;
    }
#endif
    NMPC_HD static void rec_backward_A(double* sm)
    {
#if SOLO_HOST
        for (int k = N - 1; k >= 1; k--) {
            const double* A = Jp(sm, k) + NU;
            double* rc = rec(sm, k);
            for (int j = 0; j < NX; j++) {
                const double* vn = rc + RS + R_V;
                double s0 = rc[R_V + j], s1 = 0.0;
                for (int i = 0; i < NX; i++) { if (i & 1) s1 += A[i * NZ + j] * vn[i]; else s0 += A[i * NZ + j] * vn[i]; }
                rc[R_V + j] = s0 + s1;
            }
        }
#else
        if (threadIdx.x >= 32) return;
        const int j = threadIdx.x < NX ? threadIdx.x : 0;
        double v = rec(sm, N)[R_V + j];
        double ra[NX], rb[NX];
        load_row<NZ>(ra, Jp(sm, N - 1) + NU + j);
        load_row<NZ>(rb, Jp(sm, N - 2 >= 1 ? N - 2 : 1) + NU + j);
        for (int k = N - 1; k >= 1; k -= 2) {
            v = dot_shfl(ra, v, rec(sm, k)[R_V + j]);
            if (threadIdx.x < NX) rec(sm, k)[R_V + j] = v;
            if (k - 2 >= 1) load_row<NZ>(ra, Jp(sm, k - 2) + NU + j);
            if (k - 1 >= 1) {
                v = dot_shfl(rb, v, rec(sm, k - 1)[R_V + j]);
                if (threadIdx.x < NX) rec(sm, k - 1)[R_V + j] = v;
                if (k - 3 >= 1) load_row<NZ>(rb, Jp(sm, k - 3) + NU + j);
            }
        }
#endif
    }
    NMPC_HD static void rec_backward_phi(double* sm)
    {
#if SOLO_HOST
        for (int k = N - 1; k >= 1; k--) {
            double* rc = rec(sm, k);
            for (int j = 0; j < NX; j++) rc[R_V + j] = dotp(pht(sm, k) + j * NXP, rc + RS + R_V, rc[R_V + j]);
        }
#else
        if (threadIdx.x >= 32) return;
        const int j = threadIdx.x < NX ? threadIdx.x : 0;
        double v = rec(sm, N)[R_V + j];
        double ra[NX], rb[NX];
        load_row<1>(ra, pht(sm, N - 1) + j * NXP);
        load_row<1>(rb, pht(sm, N - 2 >= 1 ? N - 2 : 1) + j * NXP);
        for (int k = N - 1; k >= 1; k -= 2) {
            v = dot_shfl(ra, v, rec(sm, k)[R_V + j]);
            if (threadIdx.x < NX) rec(sm, k)[R_V + j] = v;
            if (k - 2 >= 1) load_row<1>(ra, pht(sm, k - 2) + j * NXP);
            if (k - 1 >= 1) {
                v = dot_shfl(rb, v, rec(sm, k - 1)[R_V + j]);
                if (threadIdx.x < NX) rec(sm, k - 1)[R_V + j] = v;
                if (k - 3 >= 1) load_row<1>(rb, pht(sm, k - 3) + j * NXP);
            }
        }
#endif
    }
    NMPC_HD static void rec_forward(double* sm)
    {
#if SOLO_HOST
        for (int i = 0; i < NX; i++) { rec(sm, 0)[R_X + i] = 0.0; rec(sm, 1)[R_X + i] = rec(sm, 0)[R_C + i]; }
        for (int k = 1; k < N; k++) {
            double* rc = rec(sm, k);
            for (int i = 0; i < NX; i++) rc[RS + R_X + i] = dotp(phi(sm, k) + i * NXP, rc + R_X, rc[R_C + i]);
        }
#else
        if (threadIdx.x >= 32) return;
        const int i = threadIdx.x < NX ? threadIdx.x : 0;
        double x = rec(sm, 0)[R_C + i];
        if (threadIdx.x < NX) { rec(sm, 0)[R_X + i] = 0.0; rec(sm, 1)[R_X + i] = x; }
        double ra[NX], rb[NX];
        load_row<1>(ra, phi(sm, 1) + i * NXP);
        load_row<1>(rb, phi(sm, 2 < N ? 2 : 1) + i * NXP);
        for (int k = 1; k < N; k += 2) {
            x = dot_shfl(ra, x, rec(sm, k)[R_C + i]);
            if (threadIdx.x < NX) rec(sm, k + 1)[R_X + i] = x;
            if (k + 2 < N) load_row<1>(ra, phi(sm, k + 2) + i * NXP);
            if (k + 1 < N) {
                x = dot_shfl(rb, x, rec(sm, k + 1)[R_C + i]);
                if (threadIdx.x < NX) rec(sm, k + 2)[R_X + i] = x;
                if (k + 3 < N) load_row<1>(rb, phi(sm, k + 3) + i * NXP);
            }
        }
#endif
    }

    // ---- B, step 3: apply the step, residuals, gradients, barrier terms (the arithmetic of Rti::stage_B_update) ---------------------
    struct NbTmp { double zn1[NX], pio[NX], rr1[NX]; };
    NMPC_HD static void phase_update(double* sm, const Tables& tb, const double* We, size_t wst, const IpmOpts& o, bool first,
                                     double a_step, double sigmu, double mcw)
    {
        NbTmp tmp[SOLO_HOST ? N + 1 : 1];
        // read block: what stage k needs from stage k+1 as it is BEFORE this phase
        for (int k = SOLO_TID; k <= N; k += SOLO_NT) {
            NbTmp& t = tmp[SOLO_HOST ? k : 0];
            if (k < N) {
                const double* nb = sm + (size_t)(k + 1) * PS;
#pragma unroll
                for (int j = 0; j < NX; j++) {
                    t.pio[j] = first ? 0.0 : nb[O_PI + j];
                    t.rr1[j] = first ? 0.0 : rec(sm, k + 1)[R_V + j];
                    if (first) t.zn1[j] = j >= 3 + NV ? cold_z0(nb[O_DLB + j - 3], nb[O_DUB + j - 3], o.thr0) : 0.0;
                    else t.zn1[j] = nb[O_Z + NU + j] + a_step * nb[O_DZ + NU + j];
                }
            }
        }
        SOLO_SYNC();
        for (int k = SOLO_TID; k <= N; k += SOLO_NT) {
            const NbTmp& t = tmp[SOLO_HOST ? k : 0];
            double* st = sm + (size_t)k * PS;
            const bool hasU = k < N, hasX = k > 0;
            const double* J = Jp(sm, k);
            double ng = 0.0, nb_ = 0.0, nd = 0.0, nm = 0.0, lru = 0.0, musum = 0.0;
            // bounded components: step in the slacks and multipliers, residuals and barrier terms at the new iterate
            double b_ldo[NB2], b_dld[NB2], b_lnew[NB2], b_gam[NB2], b_Gam[NB2], b_zn[NB2];
#pragma unroll
            for (int b = 0; b < NB2; b++) {
                const bool act = (b < NV) ? hasU : hasX;
                const int w = zidx(b);
                double ll = 0.0, lu = 0.0, tl = 1.0, tu = 1.0;
                b_ldo[b] = b_dld[b] = b_lnew[b] = b_gam[b] = b_Gam[b] = 0.0;
                b_zn[b] = first ? 0.0 : st[O_Z + w];
                if (act) {
                    const double dl = st[O_DLB + b], du_ = st[O_DUB + b];
                    if (first) {
                        double z0 = 0.0, t_l = -dl, t_u = du_;
                        if (t_l < o.thr0) {
                            if (t_u < o.thr0) { z0 = 0.5 * (dl + du_); t_l = o.thr0; t_u = o.thr0; }
                            else { t_l = o.thr0; z0 = dl + o.thr0; }
                        } else if (t_u < o.thr0) { t_u = o.thr0; z0 = du_ - o.thr0; }
                        b_zn[b] = z0; tl = t_l; tu = t_u; ll = o.mu0 / t_l; lu = o.mu0 / t_u;
                    } else {
                        const double zb = st[O_Z + w], dzb = st[O_DZ + w];
                        ll = st[O_LAM + b]; lu = st[O_LAM + NB2 + b]; tl = st[O_T + b]; tu = st[O_T + NB2 + b];
                        const double rd_l = dl - zb + tl, rd_u = -du_ + zb + tu;
                        const double rm_l = ll * tl - o.tau_min + mcw * st[O_MC + b] - sigmu;
                        const double rm_u = lu * tu - o.tau_min + mcw * st[O_MC + NB2 + b] - sigmu;
                        const double dt_l = dzb - rd_l, dt_u = -dzb - rd_u;
                        const double dl_l = -(ll * dt_l + rm_l) / tl, dl_u = -(lu * dt_u + rm_u) / tu;
                        b_ldo[b] = lu - ll; b_dld[b] = dl_l - dl_u;
                        ll += a_step * dl_l; lu += a_step * dl_u; tl += a_step * dt_l; tu += a_step * dt_u;
                        b_zn[b] = zb + a_step * dzb;
                    }
                    const double zn = b_zn[b];
                    const double rd_l = dl - zn + tl, rd_u = -du_ + zn + tu;
                    const double pm_l = ll * tl, pm_u = lu * tu;
                    musum += pm_l + pm_u;
                    const double rm_l = pm_l - o.tau_min, rm_u = pm_u - o.tau_min;
                    nd = fmax(nd, fmax(fabs(rd_l), fabs(rd_u)));
                    nm = fmax(nm, fmax(fabs(rm_l), fabs(rm_u)));
                    const double ti_l = tl < o.t_min ? 1.0 / o.t_min : 1.0 / tl, ti_u = tu < o.t_min ? 1.0 / o.t_min : 1.0 / tu;
                    const double l_l = ll < o.lam_min ? o.lam_min : ll, l_u = lu < o.lam_min ? o.lam_min : lu;
                    b_Gam[b] = ti_l * l_l + ti_u * l_u;
                    b_gam[b] = ti_l * (rm_l - ll * rd_l) - ti_u * (rm_u - lu * rd_u);
                    b_lnew[b] = lu - ll;
                }
                st[O_LAM + b] = ll; st[O_LAM + NB2 + b] = lu; st[O_T + b] = tl; st[O_T + NB2 + b] = tu;
            }
            // stationarity rows, new iterate, gradient (incl. barrier terms) and Hessian diagonal
            double zn[NZ];
#pragma unroll
            for (int w = 0; w < NZ; w++) {
                const bool isu = w < NU;
                const int j = w - NU;
                const int b = isu ? w : (j >= 3 + NV ? j - 3 : -1);
                const bool on = isu ? hasU : hasX;
                const double H = isu ? (hasU ? tb.dt * tb.W[k * NY + NX + w] : 0.0) : (hasU ? tb.dt * tb.W[k * NY + j] : We[(size_t)j * wst]);
                double v1 = 0.0, v2 = 0.0;
                if (hasU) {
#pragma unroll
                    for (int i = 0; i < NX; i++) { v1 += J[i * NZ + w] * t.pio[i]; v2 += J[i * NZ + w] * t.rr1[i]; }
                }
                const double q = st[O_Q + w];
                const double zold = ((first && on) || (isu && !hasU)) ? 0.0 : st[O_Z + w];
                const double dz = (first || !on) ? 0.0 : st[O_DZ + w];
                double znew, pin = 0.0;
                if (b >= 0 && on) znew = b_zn[b];
                else znew = zold + a_step * dz;
                if (isu) {
                    if (!first && hasU) {
                        const double r = q + H * zold + b_ldo[b] + v1 + H * dz - b_dld[b] + v2;
                        lru = fmax(lru, fabs(r));
                    }
                } else if (hasX && !first) {
                    pin = st[O_PI + j] + a_step * rec(sm, k)[R_V + j];          // V_k: the multiplier step of this stage (adjoint recursion)
                }
                double g = q + H * znew - pin + (v1 + a_step * v2);
                if (b >= 0) g += b_lnew[b];
                if (on) ng = fmax(ng, fabs(g));
                if (b >= 0) g += b_gam[b];
                zn[w] = znew;
                st[O_G + w] = g;
                st[O_DG + w] = H + o.reg_prim + (b >= 0 ? b_Gam[b] : 0.0);
                if (on || isu) st[O_Z + w] = znew;
                if (!isu && hasX) st[O_PI + j] = pin;
            }
            if (hasU) {
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < NZ; w++) s += J[i * NZ + w] * zn[w];
                    const double rb = s + (st[O_B0 + i] - t.zn1[i]);
                    st[O_RB + i] = rb;
                    nb_ = fmax(nb_, fabs(rb));
                }
            }
            double* red = st + O_RED;
            red[0] = ng; red[1] = nb_; red[2] = nd; red[3] = nm; red[4] = lru; red[5] = musum;
        }
        SOLO_SYNC();
    }

    // ---- B, step 4 for two controls (diff, tric): the Riccati recursion with TWO barrier phases per stage.  What a stage costs
    //      here is the number of barrier-delimited phases (~300 cycles each for a lone block, whatever they compute), so
    //      (A) every entry of M = [B A]' P [B A] + D and of the gradient g + [B A]' (P rb + p) is formed straight from P by one
    //          thread (seven independent dot products, then one more), and
    //      (B) the Schur complement uses the closed form of the 2 x 2 block, x' Muu^-1 y = x0 y0 / d0 + (x1 - l x0)(y1 - l y0) / d1
    //          with d0 = m00, l = m10 / d0, d1 = m11 - l m10 - the same pivots as the Cholesky factor, two independent divisions,
    //          no square root on the sequential path.  The factor itself, K and the feed-forward control are formed afterwards
    //          for all stages at once (phase_gain) from Muu, S and g_u, which phase A leaves in the stage blocks. ----------------------
    static constexpr bool FAST = (NV == 2);
    NMPC_HD static void tri_decode(int e, int& v, int& w) { v = 0; while ((v + 1) * (v + 2) / 2 <= e) v++; w = e - v * (v + 1) / 2; }
    NMPC_HD static void riccati_fast(double* sm)
    {
        const int t = SOLO_TID, nt = SOLO_NT;
        double* P = sm + O_P; double* Mm = sm + O_MM; double* GP = sm + O_GP;
        {
            double* stN = sm + (size_t)N * PS;
            for (int e = t; e < NX * NXP; e += nt) P[e] = (e / NXP == e % NXP) ? stN[O_DG + NU + e / NXP] : 0.0;
            for (int j = t; j < NX; j += nt) rec(sm, N)[R_V + j] = stN[O_G + NU + j];
        }
        constexpr int Q2 = SOLO_HOST ? (NTRI + NZ) : (NTRI + NZ + THREADS - 1) / THREADS;
        constexpr int Q4 = SOLO_HOST ? (NXTRI + NX) : (NXTRI + NX + THREADS - 1) / THREADS;
        int v2[Q2], w2[Q2], i4[Q4], j4[Q4];
#pragma unroll
        for (int q = 0; q < Q2; q++) {
            const int e = t + q * nt;
            v2[q] = -1; w2[q] = -1;
            if (e < NTRI) tri_decode(e, v2[q], w2[q]); else if (e < NTRI + NZ) v2[q] = e - NTRI;
        }
#pragma unroll
        for (int q = 0; q < Q4; q++) {
            const int e = t + q * nt;
            i4[q] = -1; j4[q] = -1;
            if (e < NXTRI) tri_decode(e, i4[q], j4[q]); else if (e < NXTRI + NX) i4[q] = e - NXTRI;
        }
        SOLO_SYNC();
        for (int k = N - 1; k >= 0; k--) {
            double* st = sm + (size_t)k * PS;
            const double* J = Jp(sm, k);
            const bool hasX = k > 0;
            // ---- phase A
#pragma unroll
            for (int q = 0; q < Q2; q++) {
                const int v = v2[q], w = w2[q];
                if (v < 0) continue;
                if (!hasX && v >= NU) {                        // stage 0 has no state: S = 0, nothing else needed
                    if (w >= 0 && w < NU) st[O_KF + w * NX + (v - NU)] = 0.0;
                    continue;
                }
                const double* col = w >= 0 ? J + w : st + O_RB;            // the vector P is applied to: column w of [B A], or rb
                const int cs = w >= 0 ? NZ : 1;
                double cv[NX];
#pragma unroll
                for (int m = 0; m < NX; m++) cv[m] = col[m * cs];
                double gi[NX];
#pragma unroll
                for (int i = 0; i < NX; i++) gi[i] = dotp_r(P + i * NXP, cv, w >= 0 ? 0.0 : rec(sm, k + 1)[R_V + i]);
                double s0 = w >= 0 ? (v == w ? st[O_DG + v] : 0.0) : st[O_G + v], s1 = 0.0;
#pragma unroll
                for (int i = 0; i < NX; i++) { if (i & 1) s1 += J[i * NZ + v] * gi[i]; else s0 += J[i * NZ + v] * gi[i]; }
                const double r = s0 + s1;
                if (w >= 0) {
                    Mm[v * NZ + w] = r;
                    if (v < NU) st[O_LUU + v * (v + 1) / 2 + w] = r;                  // Muu, row-packed
                    else if (w < NU) st[O_KF + w * NX + (v - NU)] = r;               // S
                } else {
                    GP[v] = r;
                    if (v < NU) st[O_UF + v] = r;                                    // g_u incl. the cost-to-go
                }
            }
            SOLO_SYNC();
            SOLO_T(16);
            // ---- phase B
            if (hasX) {
#pragma unroll
                for (int q = 0; q < Q4; q++) {
                    const int i = i4[q], j = j4[q];
                    if (i < 0) continue;
                    const double m00 = Mm[0], m10 = Mm[NZ], m11 = Mm[NZ + 1];
                    const double det = m00 * m11 - m10 * m10;
#if SOLO_HOST
                    const double i0 = m00 > 0.0 ? 1.0 / m00 : 0.0;
                    const double i1 = (m00 > 0.0 && det > 0.0) ? m00 * (1.0 / det) : 0.0;
#else
                    const double i0 = m00 > 0.0 ? __drcp_rn(m00) : 0.0;
                    const double i1 = (m00 > 0.0 && det > 0.0) ? m00 * __drcp_rn(det) : 0.0;
#endif
                    const double l = m10 * i0;
                    const double x0 = Mm[(NU + i) * NZ], x1 = Mm[(NU + i) * NZ + 1] - l * x0;
                    if (j >= 0) {
                        const double y0 = Mm[(NU + j) * NZ], y1 = Mm[(NU + j) * NZ + 1] - l * y0;
                        const double r = Mm[(NU + i) * NZ + NU + j] - (x0 * y0 * i0 + x1 * y1 * i1);
                        P[i * NXP + j] = r; P[j * NXP + i] = r;
                    } else {
                        const double y0 = GP[0], y1 = GP[1] - l * y0;
                        rec(sm, k)[R_V + i] = GP[NU + i] - (x0 * y0 * i0 + x1 * y1 * i1);
                    }
                }
            }
            SOLO_SYNC();
            SOLO_T(19);
        }
    }

    // ---- B, step 4: the Riccati recursion, matrices and gradient together (the arithmetic of Rti::stage_B_riccati), by the whole
    //      block: per stage four sub-phases whose work items are spread over the threads, a barrier after each ------------------------
    NMPC_HD static void riccati(double* sm)
    {
        const int t = SOLO_TID, nt = SOLO_NT;
        double* P = sm + O_P; double* Gm = sm + O_GM; double* Mm = sm + O_MM;
        double* WV = sm + O_WV; double* GP = sm + O_GP; double* LH = sm + O_LH;
        {
            double* stN = sm + (size_t)N * PS;
            for (int e = t; e < NX * NX; e += nt) P[e] = (e / NX == e % NX) ? stN[O_DG + NU + e / NX] : 0.0;
            for (int j = t; j < NX; j += nt) rec(sm, N)[R_V + j] = stN[O_G + NU + j];
        }
        // item -> index maps of the triangular sub-phases (the same for every stage)
        constexpr int Q2 = SOLO_HOST ? (NTRI + NZ) : (NTRI + NZ + THREADS - 1) / THREADS;
        constexpr int Q4 = SOLO_HOST ? (NXTRI + NX) : (NXTRI + NX + THREADS - 1) / THREADS;
        int v2[Q2], w2[Q2], i4[Q4], j4[Q4];
#pragma unroll
        for (int q = 0; q < Q2; q++) {
            const int e = t + q * nt;
            v2[q] = -1; w2[q] = -1;
            if (e < NTRI) tri_decode(e, v2[q], w2[q]); else if (e < NTRI + NZ) v2[q] = e - NTRI;
        }
#pragma unroll
        for (int q = 0; q < Q4; q++) {
            const int e = t + q * nt;
            i4[q] = -1; j4[q] = -1;
            if (e < NXTRI) tri_decode(e, i4[q], j4[q]); else if (e < NXTRI + NX) i4[q] = e - NXTRI;
        }
        // BIG models: [B A] of the stage is staged from the global scratch into a double buffer in shared memory, one stage ahead
        // (16-byte asynchronous copies by the first threads of the block)
        double* Jb = sm + O_JB;
        auto stage_J = [&](int kk, int buf) {
            if (BIG) {
                for (int e = 2 * t; e < JSZ; e += 2 * nt) grp_cp16(Jb + buf * JSZ + e, Jp(sm, kk) + e);
                grp_cp_commit();
            }
        };
        stage_J(N - 1, 0);
        grp_cp_wait<0>();
        SOLO_SYNC();
        for (int k = N - 1; k >= 0; k--) {
            double* st = sm + (size_t)k * PS;
            const int cur = (N - 1 - k) & 1;
            const double* J = BIG ? Jb + cur * JSZ : Jp(sm, k);
            if (k >= 1) stage_J(k - 1, cur ^ 1);
            const bool hasX = k > 0;
            // G = P [B A];  w = P rb + p_{k+1}
            for (int e = t; e < NX * NZ + NX; e += nt) {
                if (e < NX * NZ) {
                    const int i = e / NZ, w = e - i * NZ;
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int m = 0; m < NX; m++) { if (m & 1) s1 += P[i * NX + m] * J[m * NZ + w]; else s0 += P[i * NX + m] * J[m * NZ + w]; }
                    Gm[e] = s0 + s1;
                } else {
                    const int i = e - NX * NZ;
                    double s0 = rec(sm, k + 1)[R_V + i], s1 = 0.0;
#pragma unroll
                    for (int m = 0; m < NX; m++) { if (m & 1) s1 += P[i * NX + m] * st[O_RB + m]; else s0 += P[i * NX + m] * st[O_RB + m]; }
                    WV[i] = s0 + s1;
                }
            }
            SOLO_SYNC();
            SOLO_T(16);
            // M = [B A]' G + D (lower triangle);  gradient incl. the cost-to-go: g + [B A]' w
#pragma unroll
            for (int q = 0; q < Q2; q++) {
                const int v = v2[q], w = w2[q];
                if (v < 0) continue;
                if (w >= 0) {
                    double s0 = (v == w) ? st[O_DG + v] : 0.0, s1 = 0.0;
#pragma unroll
                    for (int i = 0; i < NX; i++) { if (i & 1) s1 += J[i * NZ + v] * Gm[i * NZ + w]; else s0 += J[i * NZ + v] * Gm[i * NZ + w]; }
                    Mm[v * NZ + w] = s0 + s1;
                } else {
                    double s0 = st[O_G + v], s1 = 0.0;
#pragma unroll
                    for (int i = 0; i < NX; i++) { if (i & 1) s1 += J[i * NZ + v] * WV[i]; else s0 += J[i * NZ + v] * WV[i]; }
                    GP[v] = s0 + s1;
                }
            }
            SOLO_SYNC();
            SOLO_T(17);
            // Cholesky of the control block; item j < NX: column j of K = Luu^-1 S; item NX: Luu^-1 g_u, the feed-forward control
            for (int j = t; j <= NX; j += nt) {
                double Luu[NLU];
#pragma unroll
                for (int a = 0; a < NV; a++) {
                    double d = Mm[a * NZ + a];
#pragma unroll
                    for (int c = 0; c < a; c++) d -= Luu[a * (a + 1) / 2 + c] * Luu[a * (a + 1) / 2 + c];
#if SOLO_HOST
                    const double inv = d > 0.0 ? 1.0 / sqrt(d) : 0.0;
#else
                    const double inv = d > 0.0 ? rsqrt(d) : 0.0;
#endif
                    Luu[a * (a + 1) / 2 + a] = inv;
#pragma unroll
                    for (int b = a + 1; b < NV; b++) {
                        double s = Mm[b * NZ + a];
#pragma unroll
                        for (int c = 0; c < a; c++) s -= Luu[b * (b + 1) / 2 + c] * Luu[a * (a + 1) / 2 + c];
                        Luu[b * (b + 1) / 2 + a] = s * inv;
                    }
                }
                if (j < NX) {
                    double kh[NV];
#pragma unroll
                    for (int a = 0; a < NV; a++) {
                        double s = hasX ? Mm[(NU + j) * NZ + a] : 0.0;
#pragma unroll
                        for (int c = 0; c < a; c++) s -= Luu[a * (a + 1) / 2 + c] * kh[c];
                        kh[a] = s * Luu[a * (a + 1) / 2 + a];
                        st[O_KF + a * NX + j] = kh[a];
                    }
                } else {
                    double lh[NV], uf[NV];
#pragma unroll
                    for (int a = 0; a < NV; a++) {
                        double s = GP[a];
#pragma unroll
                        for (int c = 0; c < a; c++) s -= Luu[a * (a + 1) / 2 + c] * lh[c];
                        lh[a] = s * Luu[a * (a + 1) / 2 + a];
                        LH[a] = lh[a];
                    }
#pragma unroll
                    for (int a = NV - 1; a >= 0; a--) {
                        double s = -lh[a];
#pragma unroll
                        for (int b = a + 1; b < NV; b++) s -= Luu[b * (b + 1) / 2 + a] * uf[b];
                        uf[a] = s * Luu[a * (a + 1) / 2 + a];
                        st[O_UF + a] = uf[a];
                    }
#pragma unroll
                    for (int q = 0; q < NLU; q++) st[O_LUU + q] = Luu[q];
                }
            }
            SOLO_SYNC();
            SOLO_T(18);
            // Schur complement: this stage's cost-to-go and its gradient
            if (hasX) {
#pragma unroll
                for (int q = 0; q < Q4; q++) {
                    const int i = i4[q], j = j4[q];
                    if (i < 0) continue;
                    if (j >= 0) {
                        double s = Mm[(NU + i) * NZ + NU + j];
#pragma unroll
                        for (int a = 0; a < NV; a++) s -= st[O_KF + a * NX + i] * st[O_KF + a * NX + j];
                        P[i * NX + j] = s; P[j * NX + i] = s;
                    } else {
                        double s = GP[NU + i];
#pragma unroll
                        for (int a = 0; a < NV; a++) s -= st[O_KF + a * NX + i] * LH[a];
                        rec(sm, k)[R_V + i] = s;
                    }
                }
            }
            grp_cp_wait<0>();
            SOLO_SYNC();
            SOLO_T(19);
        }
    }

    // ---- B, step 5: the gain Luu^-T K, the closed-loop matrix, the constant of the predictor's forward recursion --------------------
    NMPC_HD static void phase_gain(double* sm)
    {
        for (int k = SOLO_TID; k < N; k += SOLO_NT) {
            double* st = sm + (size_t)k * PS;
            const double* J = Jp(sm, k); const double* Luu = st + O_LUU;
            if (FAST) {
                // the two-phase recursion left Muu (row-packed), S and g_u: factorise, K = Luu^-1 S, feed-forward control
                double L[NLU];
#pragma unroll
                for (int a = 0; a < NV; a++) {
                    double d = st[O_LUU + a * (a + 1) / 2 + a];
#pragma unroll
                    for (int c = 0; c < a; c++) d -= L[a * (a + 1) / 2 + c] * L[a * (a + 1) / 2 + c];
                    const double inv = d > 0.0 ? 1.0 / sqrt(d) : 0.0;
                    L[a * (a + 1) / 2 + a] = inv;
#pragma unroll
                    for (int b = a + 1; b < NV; b++) {
                        double sacc = st[O_LUU + b * (b + 1) / 2 + a];
#pragma unroll
                        for (int c = 0; c < a; c++) sacc -= L[b * (b + 1) / 2 + c] * L[a * (a + 1) / 2 + c];
                        L[b * (b + 1) / 2 + a] = sacc * inv;
                    }
                }
#pragma unroll
                for (int q = 0; q < NLU; q++) st[O_LUU + q] = L[q];
#pragma unroll
                for (int j = 0; j < NX; j++) {
                    double kh[NV];
#pragma unroll
                    for (int a = 0; a < NV; a++) {
                        double sacc = st[O_KF + a * NX + j];
#pragma unroll
                        for (int c = 0; c < a; c++) sacc -= L[a * (a + 1) / 2 + c] * kh[c];
                        kh[a] = sacc * L[a * (a + 1) / 2 + a];
                        st[O_KF + a * NX + j] = kh[a];
                    }
                }
                double gu[NV], uf[NV];
#pragma unroll
                for (int a = 0; a < NV; a++) gu[a] = st[O_UF + a];
                neg_solve_uu(L, gu, uf);
#pragma unroll
                for (int a = 0; a < NV; a++) st[O_UF + a] = uf[a];
            }
#pragma unroll
            for (int i = 0; i < NX; i++) {
                double s = st[O_RB + i];
#pragma unroll
                for (int a = 0; a < NV; a++) s += J[i * NZ + a] * st[O_UF + a];
                rec(sm, k)[R_C + i] = s;
            }
        }
        if (FAST) SOLO_SYNC();
        // column j of the gain and of the closed-loop matrix of stage k: one work item per (stage, column), all threads
        for (int e = SOLO_TID; e < N * NX; e += SOLO_NT) {
            const int k = e / NX, j = e - k * NX;
            double* st = sm + (size_t)k * PS;
            const double* J = Jp(sm, k); const double* Luu = st + O_LUU;
            double kf[NV];
#pragma unroll
            for (int a = NV - 1; a >= 0; a--) {
                double s = st[O_KF + a * NX + j];
#pragma unroll
                for (int b = a + 1; b < NV; b++) s -= Luu[b * (b + 1) / 2 + a] * kf[b];
                kf[a] = s * Luu[a * (a + 1) / 2 + a];
            }
#pragma unroll
            for (int a = 0; a < NV; a++) st[O_KF + a * NX + j] = kf[a];
#pragma unroll
            for (int i = 0; i < NX; i++) {
                double s = J[i * NZ + NU + j];
#pragma unroll
                for (int a = 0; a < NV; a++) s -= J[i * NZ + a] * kf[a];
                phi(sm, k)[i * NXP + j] = s; pht(sm, k)[j * NXP + i] = s;
            }
        }
        SOLO_SYNC();
    }

    // ---- delta solve: feed-forward control -Muu^-1 (g_u + B' d_{k+1}) and the constant of its forward recursion ------------------------
    NMPC_HD static void phase_feedforward(double* sm)
    {
        for (int k = SOLO_TID; k < N; k += SOLO_NT) {
            double* st = sm + (size_t)k * PS;
            const double* J = Jp(sm, k);
            const double* vn = rec(sm, k + 1) + R_V;
            double tv[NV], uf[NV];
#pragma unroll
            for (int a = 0; a < NV; a++) {
                double s = st[O_G + a];
#pragma unroll
                for (int i = 0; i < NX; i++) s += J[i * NZ + a] * vn[i];
                tv[a] = s;
            }
            neg_solve_uu(st + O_LUU, tv, uf);
#pragma unroll
            for (int a = 0; a < NV; a++) st[O_UF + a] = uf[a];
#pragma unroll
            for (int i = 0; i < NX; i++) {
                double s = 0.0;
#pragma unroll
                for (int a = 0; a < NV; a++) s += J[i * NZ + a] * uf[a];
                rec(sm, k)[R_C + i] = s;
            }
        }
        SOLO_SYNC();
    }

    // ---- after a forward recursion: the control steps, the step in slacks / multipliers, ratio test and complementarity sums
    //      (the arithmetic of Rti::stage_F).  delta == false: predictor into DZA, writes MC;  true: DZ = DZA + delta step ---------------
    NMPC_HD static void phase_step(double* sm, const IpmOpts& o, bool delta, double sigmu, double mcw)
    {
        for (int k = SOLO_TID; k <= N; k += SOLO_NT) {
            double* st = sm + (size_t)k * PS;
            const bool hasU = k < N, hasX = k > 0;
            double* dzo = st + (delta ? O_DZ : O_DZA);
#pragma unroll
            for (int j = 0; j < NX; j++) dzo[NU + j] = rec(sm, k)[R_X + j];
            if (hasU) {
#pragma unroll
                for (int a = 0; a < NV; a++) {
                    double s = st[O_UF + a];
                    if (hasX) for (int j = 0; j < NX; j++) s -= st[O_KF + a * NX + j] * dzo[NU + j];
                    dzo[a] = s;
                }
            } else {
#pragma unroll
                for (int a = 0; a < NV; a++) dzo[a] = 0.0;
            }
            if (delta) for (int w = 0; w < NZ; w++) dzo[w] += st[O_DZA + w];
            double alpha = -1.0, S0 = 0.0, S1 = 0.0, S2 = 0.0;
#pragma unroll
            for (int b = 0; b < NB2; b++) {
                const bool act = (b < NV) ? hasU : hasX;
                if (!act) continue;
                const int w = zidx(b);
                const double ll = st[O_LAM + b], lu = st[O_LAM + NB2 + b], tl = st[O_T + b], tu = st[O_T + NB2 + b];
                const double zb = st[O_Z + w], dzb = dzo[w];
                const double rd_l = st[O_DLB + b] - zb + tl, rd_u = -st[O_DUB + b] + zb + tu;
                double rm_l = ll * tl - o.tau_min, rm_u = lu * tu - o.tau_min;
                if (delta) { rm_l += mcw * st[O_MC + b] - sigmu; rm_u += mcw * st[O_MC + NB2 + b] - sigmu; }
                const double dt_l = dzb - rd_l, dt_u = -dzb - rd_u;
                const double dl_l = -(ll * dt_l + rm_l) / tl, dl_u = -(lu * dt_u + rm_u) / tu;
                if (!delta) { st[O_MC + b] = dt_l * dl_l; st[O_MC + NB2 + b] = dt_u * dl_u; }
                if (alpha * dl_l > ll) alpha = ll / dl_l;
                if (alpha * dt_l > tl) alpha = tl / dt_l;
                if (alpha * dl_u > lu) alpha = lu / dl_u;
                if (alpha * dt_u > tu) alpha = tu / dt_u;
                S0 += ll * tl + lu * tu;
                S1 += ll * dt_l + tl * dl_l + lu * dt_u + tu * dl_u;
                S2 += dl_l * dt_l + dl_u * dt_u;
            }
            double* red = st + O_RED;
            red[0] = alpha; red[1] = S0; red[2] = S1; red[3] = S2;
        }
        SOLO_SYNC();
    }

    // ---- delta solve, step 1: right-hand side in the complementarity rows only (Rti::stage_Bd) ------------------------------------------
    NMPC_HD static void phase_delta_rhs(double* sm, double sigmu, double mcw)
    {
        for (int k = SOLO_TID; k <= N; k += SOLO_NT) {
            double* st = sm + (size_t)k * PS;
            const bool hasU = k < N, hasX = k > 0;
            double gq[NB2];
#pragma unroll
            for (int b = 0; b < NB2; b++) {
                const bool act = (b < NV) ? hasU : hasX;
                gq[b] = act ? (mcw * st[O_MC + b] - sigmu) / st[O_T + b] - (mcw * st[O_MC + NB2 + b] - sigmu) / st[O_T + NB2 + b] : 0.0;
            }
#pragma unroll
            for (int a = 0; a < NV; a++) st[O_G + a] = gq[a];
#pragma unroll
            for (int j = 0; j < NX; j++) {
                double s = j >= 3 + NV ? gq[j - 3] : 0.0;
                if (hasU) for (int a = 0; a < NV; a++) s -= st[O_KF + a * NX + j] * gq[a];
                rec(sm, k)[R_V + j] = s;
            }
        }
        SOLO_SYNC();
    }

    // ---- reductions over the stages by thread 0 (81 x a few values) and the control logic between the sweeps -------------------------
    NMPC_HD static LaneCtl& ctl(double* sm) { return *reinterpret_cast<LaneCtl*>(sm + O_CTL); }
    // the first nmax of nval per-stage values by max, the rest by sum, over all stages; by warp 0 (three stages per lane, then a
    // shuffle butterfly), valid on thread 0
    NMPC_HD static void reduce_stages(const double* sm, int nmax, int nval, double* out)
    {
#if SOLO_HOST
        for (int q = 0; q < nval; q++) out[q] = sm[O_RED + q];
        for (int k = 1; k <= N; k++)
            for (int q = 0; q < nval; q++) {
                const double v = sm[(size_t)k * PS + O_RED + q];
                out[q] = q < nmax ? fmax(out[q], v) : out[q] + v;
            }
#else
        const int l = threadIdx.x;
#pragma unroll
        for (int q = 0; q < NRED; q++) {
            if (q >= nval) break;
            double acc = sm[(size_t)l * PS + O_RED + q];
            for (int k = l + 32; k <= N; k += 32) {
                const double v = sm[(size_t)k * PS + O_RED + q];
                acc = q < nmax ? fmax(acc, v) : acc + v;
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const double v = __shfl_xor_sync(0xffffffffu, acc, m);
                acc = q < nmax ? fmax(acc, v) : acc + v;
            }
            out[q] = acc;
        }
#endif
    }
    NMPC_HD static void reduce_B(double* sm, const IpmOpts& o, bool first)
    {
        static_assert(N >= 31, "every lane of the reducing warp owns a stage");
        if (SOLO_TID < 32) {
            double r[NRED];
            reduce_stages(sm, 5, 6, r);
            if (SOLO_TID == 0) {
                typename S::CarryB cy;
                cy.sc = nullptr; cy.cur = 0;
                cy.ng = r[0]; cy.nb = r[1]; cy.nd = r[2]; cy.nm = r[3]; cy.lru = r[4]; cy.musum = r[5];
                S::after_B(ctl(sm), cy, o, first);
            }
        }
        SOLO_SYNC();
    }
    NMPC_HD static void reduce_F(double* sm, const IpmOpts& o, int kind)      // 0: predictor, 1: corrector, 2: centering repeat
    {
        if (SOLO_TID < 32) {
            double r[NRED];
            reduce_stages(sm, 1, 4, r);
            if (SOLO_TID == 0) {
                typename S::CarryF cy;
                cy.init();
                cy.alpha = r[0]; cy.S0 = r[1]; cy.S1 = r[2]; cy.S2 = r[3];
                if (kind == 0) S::after_F(ctl(sm), cy, o);
                else if (kind == 1) S::after_Fd(ctl(sm), cy, o);
                else S::after_Fd_fallback(ctl(sm), cy);
            }
        }
        SOLO_SYNC();
    }

    // one factorising pass: (apply the step) + residuals + exit test + Riccati + cost-to-go gradient + predictor feed-forward
    NMPC_HD static void pass_B(double* sm, const Tables& tb, const double* We, size_t wst, const IpmOpts& o, bool first)
    {
        double a = 0.0, sigmu = 0.0, mcw = 0.0;
        if (!first) {
            if (SOLO_TID == 0) sm[O_SC] = S::before_B(ctl(sm));
            SOLO_SYNC();
            a = sm[O_SC]; sigmu = ctl(sm).sigmu; mcw = ctl(sm).mcw;
            phase_adjoint_const(sm, tb, We, wst, o, sigmu, mcw);
            SOLO_T(1);
            rec_backward_A(sm);
            SOLO_SYNC();
            SOLO_T(2);
        }
        phase_update(sm, tb, We, wst, o, first, a, sigmu, mcw);
        SOLO_T(3);
        reduce_B(sm, o, first);
        SOLO_T(4);
        if (ctl(sm).done) return;
        if (FAST) riccati_fast(sm); else riccati(sm);
        SOLO_T(5);
        phase_gain(sm);
        SOLO_T(6);
    }
    NMPC_HD static void pass_delta(double* sm, const IpmOpts& o, double mcw, int kind)
    {
        const double sigmu = ctl(sm).sigmu;
        phase_delta_rhs(sm, sigmu, mcw);
        SOLO_T(10);
        rec_backward_phi(sm);
        SOLO_SYNC();
        SOLO_T(11);
        phase_feedforward(sm);
        SOLO_T(12);
        rec_forward(sm);
        SOLO_SYNC();
        SOLO_T(13);
        phase_step(sm, o, true, sigmu, mcw);
        SOLO_T(14);
        reduce_F(sm, o, kind);
        SOLO_T(15);
    }

    // the whole interior-point solve of the instance behind `io`; every thread of the block calls it.
    // Leaves the QP solution where K4 reads it (io.put_z) and the statistics in `out` (thread 0).
    template <class IO>
    NMPC_HD static void run(double* sm, const IO& io, const Tables& tb, const double* We, size_t wst, const IpmOpts& o,
                            typename S::LaneStats* out, double* gscratch = nullptr)
    {
        if (SOLO_TID == 0) gscr(sm) = gscratch;          // BIG models: GSCR_DOUBLES of global memory of this instance's own
        SOLO_SYNC();
        SOLO_T0();
        load_qp(sm, io, tb);
        if (SOLO_TID == 0) ctl(sm).init(true);
        SOLO_SYNC();
        SOLO_T(0);
        pass_B(sm, tb, We, wst, o, true);
        while (!ctl(sm).done) {
            rec_forward(sm);
            SOLO_SYNC();
            SOLO_T(7);
            phase_step(sm, o, false, 0.0, 0.0);
            SOLO_T(8);
            reduce_F(sm, o, 0);
            SOLO_T(9);
            pass_delta(sm, o, 1.0, 1);
            if (ctl(sm).fb) pass_delta(sm, o, 0.0, 2);
            pass_B(sm, tb, We, wst, o, false);
        }
        for (int k = SOLO_TID; k <= N; k += SOLO_NT) {
            const double* st = sm + (size_t)k * PS;
#pragma unroll
            for (int w = 0; w < NZ; w++) io.put_z(k, w, st[O_Z + w]);
        }
        SOLO_TEND();
        if (SOLO_TID == 0 && out) {
            const LaneCtl& c = ctl(sm);
            out->status = c.status; out->iter = c.iter; out->mu = c.mu; out->lin_res = c.lin_res; out->cond_fallbacks = c.nfb;
            for (int q = 0; q < 4; q++) out->res[q] = c.nrm[q];
        }
    }
};

}  // namespace nmpc
