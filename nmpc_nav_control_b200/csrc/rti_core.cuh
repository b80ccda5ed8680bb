// Per-lane (= per OCP instance) SQP-RTI logic: K1 RK4+sensitivities, K2 Gauss-Newton cost
// linearisation, K3 Riccati / primal-dual interior point, K4 step.
//
// Replaces what the reference reaches through `{m}_acados_solve(capsule)`
// (src/nmpc_nav_control/NMPCNavControlDiff.cpp:142, Omni4.cpp:139, Tric.cpp:146), i.e. the
// un-vendored acados ocp_nlp_sqp_rti -> sim_erk -> ocp_nlp_cost_nls -> HPIPM chain, restated
// in SURVEY.md Appendix B.  It is NOT a translation of HPIPM: the iteration path (initial
// point, Mehrotra predictor/corrector, step rules, exit test) is the same so that results
// agree with the oracle to ~1e-12, but the linear algebra is re-derived for this problem family:
//
//  * structure: x = [pose(3) | actual(NV) | ref(NV)]; only the 3 pose rows of [A|B] vary per
//    stage/instance (3 x (1+3NV) numbers, "E"); the lag/integrator rows are 4 constants per
//    channel taken from a broadcast table.  Stage Hessians are diagonal (y = [x;u], J = I).
//  * the Riccati recursion keeps the cost-to-go P explicitly (packed symmetric, registers only,
//    never stored) and factorises only the NV x NV control block; what is stored per stage is
//    Luu, K = Luu^-1 S and the two reduced gradients.
//  * the corrector is solved as predictor + delta, the delta system having zero equality
//    residuals, so neither P nor the dynamics residual is needed outside the factorising sweep.
//  * the multiplier step d(pi) is recovered by the adjoint recursion of the stationarity rows
//    inside the next factorising sweep, which also applies the step and evaluates the new
//    residuals: one IPM iteration = 4 sweeps over the horizon (B: update+residuals+factorise,
//    F: predictor forward, B': delta backward, F': delta forward), each streaming one
//    contiguous record per stage per warp.
//
// Data layout ("tile"): 32 instances (one per lane), record[stage][field][lane] fp64, so every
// access of a warp is one fully coalesced 256-byte row and a warp's working set is contiguous.
#pragma once
#include "models.cuh"

namespace nmpc {

// Interior-point options (defaults in nmpc_default_opts())
struct IpmOpts {
    double mu0, alpha_min, res_g_max, res_b_max, res_d_max, res_m_max;
    double reg_prim, lam_min, t_min, tau_min, thr0;
    int iter_max, cond_pred_corr;
};

// Per-stage record layout of one tile.  Four field groups, each its own array
// [stage][field][lane] so that every sweep can fetch exactly the contiguous field ranges it
// needs with a few 1-D bulk copies (offsets in units of LANES doubles):
//   LIN : QP data written by K1/K2, read-only for K3
//   IT  : QP iterate (slacks, multipliers, primal, dynamics multipliers)
//   ST  : steps
//   FA  : factorisation of the last B sweep
template <int NV_, int ER_ = 3>
struct Rec {
    static constexpr int ER = ER_;             // stored pose rows of [A|B]: the theta row is a per-stage constant for diff / omni4 (Tables::thr)
    static constexpr int NV = NV_, NX = 3 + 2 * NV, NU = NV, NZ = NX + NU, NC = 1 + 3 * NV, NB2 = 2 * NV;
    // ---- LIN
    static constexpr int E = 0;                 // ER*NC  pose rows of [A|B], columns [theta | actual | ref | u]
    static constexpr int DLB = E + ER * NC;     // NB2    lb - z   for [u; ref]
    static constexpr int DUB = DLB + NB2;       // NB2    ub - z
    static constexpr int Q = DUB + NB2;         // NZ     QP gradient, order [u; x]
    static constexpr int B0 = Q + NZ;           // NX     b = phi(x,u) - x_next
    static constexpr int NF_LIN = B0 + NX;
    // ---- IT
    static constexpr int T = 0;                 // 2*NB2  slacks, lower then upper
    static constexpr int LAM = T + 2 * NB2;     // 2*NB2
    static constexpr int Z = LAM + 2 * NB2;     // NZ     [u; x]
    static constexpr int PI = Z + NZ;           // NX     multiplier of the dynamics that define x_k
    static constexpr int NF_IT = PI + NX;
    // ---- ST
    static constexpr int DZ = 0;                // NZ     final step
    static constexpr int MC = DZ + NZ;          // 2*NB2  dt_aff * dlam_aff
    static constexpr int DZA = MC + 2 * NB2;    // NZ     predictor (affine) step
    static constexpr int NF_ST = DZA + NZ;
    // ---- FA
    static constexpr int LUU = 0;               // NV(NV+1)/2 row-packed lower, diagonal stored inverted
    static constexpr int KH = LUU + NV * (NV + 1) / 2;   // NV*NX  K = Luu^-1 S
    static constexpr int LHD = KH + NV * NX;    // NV     Luu^-1 q_u (delta)
    static constexpr int LH = LHD + NV;         // NV     Luu^-1 q_u (predictor)
    static constexpr int RB = LH + NV;          // NX     dynamics residual
    static constexpr int NF_FA = RB + NX;

    static constexpr int NF = NF_LIN + NF_IT + NF_ST + NF_FA;
    // a tile = the four group arrays back to back
    static constexpr size_t OFF_LIN = 0;
    static constexpr size_t OFF_IT = OFF_LIN + (size_t)(NSTAGE + 1) * NF_LIN * LANES;
    static constexpr size_t OFF_ST = OFF_IT + (size_t)(NSTAGE + 1) * NF_IT * LANES;
    static constexpr size_t OFF_FA = OFF_ST + (size_t)(NSTAGE + 1) * NF_ST * LANES;
    static constexpr size_t tile_doubles = (size_t)(NSTAGE + 1) * NF * LANES;
};

// lane-resolved views of one stage: `in` pointers may point into a shared-memory copy (device
// pipeline) or straight into the tile (host emulation / K1 / K4); field stride is LANES
struct StageIn { const double *lin, *it, *st, *fa; };
struct StageOut { double *it, *st, *fa; };
#if defined(__CUDACC__)
#define NMPC_RESTRICT __restrict__
#else
#define NMPC_RESTRICT
#endif
struct StageInR { const double* NMPC_RESTRICT lin; const double* NMPC_RESTRICT it; const double* NMPC_RESTRICT st; };
struct StageOutR { double* NMPC_RESTRICT it; };

template <class R>
NMPC_HD StageOut tile_stage_out(double* tile_lane, int k)
{
    StageOut o;
    o.it = tile_lane + R::OFF_IT + (size_t)k * R::NF_IT * LANES;
    o.st = tile_lane + R::OFF_ST + (size_t)k * R::NF_ST * LANES;
    o.fa = tile_lane + R::OFF_FA + (size_t)k * R::NF_FA * LANES;
    return o;
}
template <class R>
NMPC_HD StageIn tile_stage_in(const double* tile_lane, int k)
{
    StageIn i;
    i.lin = tile_lane + R::OFF_LIN + (size_t)k * R::NF_LIN * LANES;
    i.it = tile_lane + R::OFF_IT + (size_t)k * R::NF_IT * LANES;
    i.st = tile_lane + R::OFF_ST + (size_t)k * R::NF_ST * LANES;
    i.fa = tile_lane + R::OFF_FA + (size_t)k * R::NF_FA * LANES;
    return i;
}

// stage-wise problem tables shared by all instances (what the C-ABI setters fill)
struct Tables {
    const double* W;      // [N][NY]   diagonal of W, order [x;u]
    const double* We;     // [NX]      diagonal of W_e
    const double* lbx;    // [N][NV]   row k -> stage k+1
    const double* ubx;
    const double* lbu;    // [N][NV]   stages 0..N-1
    const double* ubu;
    const double* p;      // [N][NP]
    const double* lti;    // [N][4*NV] av, ar, au, ru per channel (written by the LTI set-up kernel)
    double dt;
    const double* thr;    // [N][NC] theta row of [A|B] (columns theta | actual | ref | u) where it does not depend on the state (diff, omni4)
    const double* stg;    // [N+1][TROW] group path: per stage av|ar|au|ru, the constants 0 and 1, then the diagonal of W; rows padded to even (Grp::TROW)
};

template <int NV>
struct Lin {
    static constexpr int NC = 1 + 3 * NV;
    double E[3][NC];
    double av[NV], ar[NV], au[NV], ru[NV];
    NMPC_HD double e(int i, int c) const { return E[i][c]; }
    NMPC_HD double a_v(int c) const { return av[c]; }
    NMPC_HD double a_r(int c) const { return ar[c]; }
    NMPC_HD double a_u(int c) const { return au[c]; }
    NMPC_HD double r_u(int c) const { return ru[c]; }
};
// the same view with only the per-instance rows of E in registers: the theta row (where it is a per-stage constant) and
// the lag constants are read from the broadcast tables where they are used (one address per warp, L1 hits) - the
// factorising sweep of omni4 has no registers to hold them
template <int NV, int ER>
struct LinLean {
    static constexpr int NC = 1 + 3 * NV;
    double E[ER][NC];
    const double* thr;
    const double* lti;
    NMPC_HD double e(int i, int c) const { return i < ER ? E[i < ER ? i : 0][c] : thr[c]; }
    NMPC_HD double a_v(int c) const { return lti[c]; }
    NMPC_HD double a_r(int c) const { return lti[NV + c]; }
    NMPC_HD double a_u(int c) const { return lti[2 * NV + c]; }
    NMPC_HD double r_u(int c) const { return lti[3 * NV + c]; }
};

#if defined(__CUDA_ARCH__)
// 8-byte asynchronous global -> shared copy (lane-private image of the factorising sweep)
__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src) : "memory");
}
#endif

template <class M>
struct Rti {
    static constexpr int NV = M::NV, NP = M::NP, NX = 3 + 2 * NV, NU = NV, NZ = NX + NU, NY = NZ;
    static constexpr int NC = 1 + 3 * NV, NB2 = 2 * NV, NPK = NX * (NX + 1) / 2, NLU = NV * (NV + 1) / 2;
    static constexpr int NCON = 2 * (NV + (NSTAGE - 1) * NB2 + NV);   // one-sided constraints
    static constexpr int PSTRIDE = NMPC_SCRATCH_STRIDE;               // element stride of the per-lane scratch column
    static constexpr bool LEAN = NV >= NMPC_B_LEAN_MINNV;            // register economy of the factorising sweep (platform.cuh)
    static constexpr bool RICCATI_FENCES = LEAN;
    using R = Rec<NV, M::THETA_ROW_LTI ? 2 : 3>;
    using L = Lin<NV>;
    template <bool B_, class D_ = void> struct LinSel { using type = Lin<NV>; };
    template <class D_> struct LinSel<true, D_> { using type = LinLean<NV, Rec<NV, M::THETA_ROW_LTI ? 2 : 3>::ER>; };

    NMPC_HD static constexpr int pk(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

    // ------------------------------------------------------------------------------------
    // K1: classical RK4 over one shooting interval with forward sensitivities, exploiting the
    // cascade structure (acados sim_erk, ERK4, one step).
    // Ep: pose rows of d(phi)/d[theta | actual | ref | u]; lti: av,ar,au,ru per channel.
    // ------------------------------------------------------------------------------------
    NMPC_HD static void rk4_sens(const double* x, const double* u, const double* p, double h,
                                 double* xn, double (*Ep)[NC], double* lti)
    {
        const double ac[4] = {0.0, 0.5, 0.5, 1.0};
        const double bv[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
        double itau[NV];
#pragma unroll
        for (int c = 0; c < NV; c++) itau[c] = M::inv_tau(c, p);
        double Kx[NX], KSp[3][NC], KSv[NV][3];
        double Sv[NV][3], Sr[NV];
#pragma unroll
        for (int i = 0; i < NX; i++) { xn[i] = x[i]; Kx[i] = 0.0; }
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int c = 0; c < NC; c++) { Ep[i][c] = (i == 2 && c == 0) ? 1.0 : 0.0; KSp[i][c] = 0.0; }
#pragma unroll
        for (int c = 0; c < NV; c++) { Sv[c][0] = 1.0; Sv[c][1] = 0.0; Sv[c][2] = 0.0; Sr[c] = 0.0; KSv[c][0] = KSv[c][1] = KSv[c][2] = 0.0; }
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const double ha = h * ac[s];
            double xs[NX], Sps[3][NC], Svs[NV][3], Srs[NV];
#pragma unroll
            for (int i = 0; i < NX; i++) xs[i] = x[i] + (s ? ha * Kx[i] : 0.0);
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int c = 0; c < NC; c++) Sps[i][c] = ((i == 2 && c == 0) ? 1.0 : 0.0) + (s ? ha * KSp[i][c] : 0.0);
#pragma unroll
            for (int c = 0; c < NV; c++) {
                Svs[c][0] = 1.0 + (s ? ha * KSv[c][0] : 0.0);
                Svs[c][1] = 0.0 + (s ? ha * KSv[c][1] : 0.0);
                Svs[c][2] = 0.0 + (s ? ha * KSv[c][2] : 0.0);
                Srs[c] = 0.0 + (s ? ha * 1.0 : 0.0);
            }
            double g[3], Jth[3], Jv[3][NV];
            M::pose_rates(xs[2], &xs[3], p, g, Jth, Jv);
#pragma unroll
            for (int i = 0; i < 3; i++) Kx[i] = g[i];
#pragma unroll
            for (int c = 0; c < NV; c++) {
                Kx[3 + c] = -itau[c] * xs[3 + c] + itau[c] * xs[3 + NV + c];
                Kx[3 + NV + c] = u[c];
                KSv[c][0] = -itau[c] * Svs[c][0];
                KSv[c][1] = -itau[c] * Svs[c][1] + itau[c];
                KSv[c][2] = -itau[c] * Svs[c][2] + itau[c] * Srs[c];
            }
#pragma unroll
            for (int i = 0; i < 3; i++) {
                KSp[i][0] = Jth[i] * Sps[2][0];
#pragma unroll
                for (int c = 0; c < NV; c++) {
                    KSp[i][1 + c] = Jth[i] * Sps[2][1 + c] + Jv[i][c] * Svs[c][0];
                    KSp[i][1 + NV + c] = Jth[i] * Sps[2][1 + NV + c] + Jv[i][c] * Svs[c][1];
                    KSp[i][1 + 2 * NV + c] = Jth[i] * Sps[2][1 + 2 * NV + c] + Jv[i][c] * Svs[c][2];
                }
            }
            const double hb = h * bv[s];
#pragma unroll
            for (int i = 0; i < NX; i++) xn[i] += hb * Kx[i];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int c = 0; c < NC; c++) Ep[i][c] += hb * KSp[i][c];
#pragma unroll
            for (int c = 0; c < NV; c++) {
                Sv[c][0] += hb * KSv[c][0]; Sv[c][1] += hb * KSv[c][1]; Sv[c][2] += hb * KSv[c][2];
                Sr[c] += hb * 1.0;
            }
        }
        if (lti) {
#pragma unroll
            for (int c = 0; c < NV; c++) {
                lti[c] = Sv[c][0]; lti[NV + c] = Sv[c][1]; lti[2 * NV + c] = Sv[c][2]; lti[3 * NV + c] = Sr[c];
            }
        }
    }

    // ------------------------------------------------------------------------------------
    // K1+K2 for one (instance, stage): writes the LIN record of stage k (and, for stage 0, the
    // constant stage-0 state of the QP into IT.Z).
    //   xk, uk, xk1: iterate of this stage (uk, xk1 unused for k == N);  yref: nyref entries
    //   (3 = pose only, NY = full); We: terminal weight diagonal (broadcast or per instance).
    // K2 (ocp_nlp_cost_nls with y=[x;u], scripts/<m>/generate_c_code.py:30-39): gradient
    //   s_k W (y - yref), s_k = dt for k<N and 1 for k=N; the Hessian s_k W is never stored.
    // ------------------------------------------------------------------------------------
    template <class RR = R, int LS = LANES>
    NMPC_HD static void linearize_stage(int k, const double* xk, const double* uk, const double* xk1,
                                        const double* yref, int nyref, const double* x0bar,
                                        const Tables& tb, const double* We, double* lin, double* it)
    {
        if (k < NSTAGE) {
            double xn[NX], Ep[3][NC];
            rk4_sens(xk, uk, tb.p + k * NP, tb.dt, xn, Ep, nullptr);
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int c = 0; c < NC; c++) if (i < RR::ER) lin[(RR::E + i * NC + c) * LS] = Ep[i][c];
#pragma unroll
            for (int i = 0; i < NX; i++) lin[(RR::B0 + i) * LS] = xn[i] - xk1[i];
            const double* Wk = tb.W + k * NY;
#pragma unroll
            for (int c = 0; c < NU; c++) {
                const double yr = (NX + c < nyref) ? yref[NX + c] : 0.0;
                lin[(RR::Q + c) * LS] = (tb.dt * Wk[NX + c]) * (uk[c] - yr);
                lin[(RR::DLB + c) * LS] = tb.lbu[k * NV + c] - uk[c];
                lin[(RR::DUB + c) * LS] = tb.ubu[k * NV + c] - uk[c];
            }
#pragma unroll
            for (int j = 0; j < NX; j++) {
                const double yr = (j < nyref) ? yref[j] : 0.0;
                lin[(RR::Q + NU + j) * LS] = (tb.dt * Wk[j]) * (xk[j] - yr);
            }
        } else {
#pragma unroll
            for (int c = 0; c < NU; c++) lin[(RR::Q + c) * LS] = 0.0;
#pragma unroll
            for (int j = 0; j < NX; j++) {
                const double yr = (j < nyref) ? yref[j] : 0.0;
                lin[(RR::Q + NU + j) * LS] = We[j] * (xk[j] - yr);
            }
        }
        if (k >= 1) {
#pragma unroll
            for (int c = 0; c < NV; c++) {
                lin[(RR::DLB + NV + c) * LS] = tb.lbx[(k - 1) * NV + c] - xk[3 + NV + c];
                lin[(RR::DUB + NV + c) * LS] = tb.ubx[(k - 1) * NV + c] - xk[3 + NV + c];
            }
        } else {
            // x0 elimination (d_ocp_qp_reduce_eq_dof): the stage-0 state is the constant x0bar - x_0
#pragma unroll
            for (int j = 0; j < NX; j++) it[(RR::Z + NU + j) * LS] = x0bar[j] - xk[j];
        }
    }

    // ------------------------------------------------------------------------------------
    // Cold start of the interior point for one (instance, stage) (HPIPM INIT_VAR with
    // warm_start = 0; the `first` branch of stage_B_update), written straight into the record so
    // that the group path's first factorising sweep is an ordinary one with step length 0:
    // z = 0 (bounded components projected thr0 inside their bounds; the stage-0 state stays the
    // constant written by linearize_stage), pi = 0, slacks from the bounds, lam = mu0 / t, and
    // zero step / second-order terms.   lin: DLB, DUB of the stage;  it: T, LAM, Z, PI, DZ, MC.
    // ------------------------------------------------------------------------------------
    template <class RR, int LS>
    NMPC_HD static void coldstart_stage(int k, const IpmOpts& o, const double* lin, double* it)
    {
        const bool hasU = k < NSTAGE, hasX = k > 0;
#pragma unroll
        for (int c = 0; c < NU; c++) it[(RR::Z + c) * LS] = 0.0;
        if (hasX) {
#pragma unroll
            for (int j = 0; j < NX; j++) it[(RR::Z + NU + j) * LS] = 0.0;
        }
#pragma unroll
        for (int b = 0; b < NB2; b++) {
            const bool act = (b < NV) ? hasU : hasX;
            const double dl = lin[(RR::DLB + b) * LS], du_ = lin[(RR::DUB + b) * LS];
            double zb = 0.0, t_l = -dl, t_u = du_;
            if (t_l < o.thr0) {
                if (t_u < o.thr0) { zb = 0.5 * (dl + du_); t_l = o.thr0; t_u = o.thr0; }
                else { t_l = o.thr0; zb = dl + o.thr0; }
            } else if (t_u < o.thr0) { t_u = o.thr0; zb = du_ - o.thr0; }
            if (act) {
                it[(RR::Z + (b < NV ? b : NU + 3 + b)) * LS] = zb;
                it[(RR::T + b) * LS] = t_l; it[(RR::T + NB2 + b) * LS] = t_u;
                it[(RR::LAM + b) * LS] = o.mu0 / t_l; it[(RR::LAM + NB2 + b) * LS] = o.mu0 / t_u;
            } else {
                it[(RR::T + b) * LS] = 1.0; it[(RR::T + NB2 + b) * LS] = 1.0;
                it[(RR::LAM + b) * LS] = 0.0; it[(RR::LAM + NB2 + b) * LS] = 0.0;
            }
            it[(RR::MC + b) * LS] = 0.0; it[(RR::MC + NB2 + b) * LS] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < NX; j++) it[(RR::PI + j) * LS] = 0.0;
#pragma unroll
        for (int c = 0; c < NZ; c++) it[(RR::DZ + c) * LS] = 0.0;
    }

    // ------------------------------------------------------------------------------------
    // structured products with [B A] of one stage
    // ------------------------------------------------------------------------------------
    using LLean = LinLean<NV, R::ER>;
    NMPC_HD static void load_lin(const double* lin, const double* lti, const double* thr, LLean& l)
    {
#pragma unroll
        for (int i = 0; i < R::ER; i++)
#pragma unroll
            for (int c = 0; c < NC; c++) l.E[i][c] = lin[(R::E + i * NC + c) * LANES];
        l.thr = thr; l.lti = lti;
    }
    NMPC_HD static void load_lin(const double* lin, const double* lti, const double* thr, L& l)
    {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int c = 0; c < NC; c++) l.E[i][c] = i < R::ER ? lin[(R::E + i * NC + c) * LANES] : thr[c];
#pragma unroll
        for (int c = 0; c < NV; c++) { l.av[c] = lti[c]; l.ar[c] = lti[NV + c]; l.au[c] = lti[2 * NV + c]; l.ru[c] = lti[3 * NV + c]; }
    }
    // [B A]' v  -> ou (NV), ox (NX)
    template <class LL>
    NMPC_HD static void apply_T(const LL& l, const double* v, double* ou, double* ox)
    {
#pragma unroll
        for (int c = 0; c < NV; c++)
            ou[c] = l.e(0, 1 + 2 * NV + c) * v[0] + l.e(1, 1 + 2 * NV + c) * v[1] + l.e(2, 1 + 2 * NV + c) * v[2]
                  + l.a_u(c) * v[3 + c] + l.r_u(c) * v[3 + NV + c];
        ox[0] = v[0]; ox[1] = v[1];
        ox[2] = l.e(0, 0) * v[0] + l.e(1, 0) * v[1] + l.e(2, 0) * v[2];
#pragma unroll
        for (int c = 0; c < NV; c++) {
            ox[3 + c] = l.e(0, 1 + c) * v[0] + l.e(1, 1 + c) * v[1] + l.e(2, 1 + c) * v[2] + l.a_v(c) * v[3 + c];
            ox[3 + NV + c] = l.e(0, 1 + NV + c) * v[0] + l.e(1, 1 + NV + c) * v[1] + l.e(2, 1 + NV + c) * v[2]
                           + l.a_r(c) * v[3 + c] + v[3 + NV + c];
        }
    }
    // A x + B u -> xn
    template <class LL>
    NMPC_HD static void apply(const LL& l, const double* u, const double* x, double* xn)
    {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            double a = (i < 2 ? x[i] : 0.0) + l.e(i, 0) * x[2];
#pragma unroll
            for (int c = 0; c < NV; c++)
                a += l.e(i, 1 + c) * x[3 + c] + l.e(i, 1 + NV + c) * x[3 + NV + c] + l.e(i, 1 + 2 * NV + c) * u[c];
            xn[i] = a;
        }
#pragma unroll
        for (int c = 0; c < NV; c++) {
            xn[3 + c] = l.a_v(c) * x[3 + c] + l.a_r(c) * x[3 + NV + c] + l.a_u(c) * u[c];
            xn[3 + NV + c] = x[3 + NV + c] + l.r_u(c) * u[c];
        }
    }
    // g = P * (column j of [B A]), j in z order [u; x]
    template <class LL>
    NMPC_HD static void P_col(const double* P, const LL& l, int j, double* g)
    {
        if (j < NV) {
            const int c = j;
#pragma unroll
            for (int i = 0; i < NX; i++)
                g[i] = P[pk(i, 0)] * l.e(0, 1 + 2 * NV + c) + P[pk(i, 1)] * l.e(1, 1 + 2 * NV + c) + P[pk(i, 2)] * l.e(2, 1 + 2 * NV + c)
                     + P[pk(i, 3 + c)] * l.a_u(c) + P[pk(i, 3 + NV + c)] * l.r_u(c);
        } else if (j < NV + 2) {
#pragma unroll
            for (int i = 0; i < NX; i++) g[i] = P[pk(i, j - NV)];
        } else if (j == NV + 2) {
#pragma unroll
            for (int i = 0; i < NX; i++) g[i] = P[pk(i, 0)] * l.e(0, 0) + P[pk(i, 1)] * l.e(1, 0) + P[pk(i, 2)] * l.e(2, 0);
        } else if (j < NV + 3 + NV) {
            const int c = j - NV - 3;
#pragma unroll
            for (int i = 0; i < NX; i++)
                g[i] = P[pk(i, 0)] * l.e(0, 1 + c) + P[pk(i, 1)] * l.e(1, 1 + c) + P[pk(i, 2)] * l.e(2, 1 + c) + P[pk(i, 3 + c)] * l.a_v(c);
        } else {
            const int c = j - NV - 3 - NV;
#pragma unroll
            for (int i = 0; i < NX; i++)
                g[i] = P[pk(i, 0)] * l.e(0, 1 + NV + c) + P[pk(i, 1)] * l.e(1, 1 + NV + c) + P[pk(i, 2)] * l.e(2, 1 + NV + c)
                     + P[pk(i, 3 + c)] * l.a_r(c) + P[pk(i, 3 + NV + c)];
        }
    }

    // u = -Luu^-T (lh + K dx), Luu row-packed with inverted diagonal
    NMPC_HD static void solve_u(const double* fa, int lh_field, bool with_K, const double* dx, double* du)
    {
        double v[NV];
#pragma unroll
        for (int a = 0; a < NV; a++) {
            double s = fa[(lh_field + a) * LANES];
            if (with_K) {
#pragma unroll
                for (int j = 0; j < NX; j++) s += fa[(R::KH + a * NX + j) * LANES] * dx[j];
            }
            v[a] = -s;
        }
#pragma unroll
        for (int a = NV - 1; a >= 0; a--) {
            double s = v[a];
#pragma unroll
            for (int b = a + 1; b < NV; b++) s -= fa[(R::LUU + b * (b + 1) / 2 + a) * LANES] * du[b];
            du[a] = s * fa[(R::LUU + a * (a + 1) / 2 + a) * LANES];
        }
    }


    // ====================================================================================
    // K3 sweeps, written as per-stage functions with explicit carries so that the device
    // pipeline can stream one stage record at a time through shared memory.
    // ====================================================================================

    // ---- B sweep: (apply previous step) + residuals + Riccati factorisation, stage N..0 -----
    // All vectors and matrices carried from stage to stage live in a per-lane scratch column `sc`
    // (shared memory on the device, element stride PSTRIDE; a plain array in the host emulation):
    // the cost-to-go matrix P (packed symmetric, ONE buffer: [A B] is upper triangular in the state order
    // x, y, theta, actual, ref, so entry (i, j) of A'PA only reads P(k, l) with k <= i, l <= j and the
    // congruence can overwrite P column by column from the last one), its gradient pv, and the three vectors of the
    // update / adjoint recursion.  Registers only carry the norm accumulators.  The stage body is
    // cut in two phases (update + residuals, then Riccati) so that neither holds the other's data.
    struct CarryB {
        static constexpr int SC_P = 0, SC_PV = NPK, SC_PIO = SC_PV + NX, SC_DPI = SC_PIO + NX, SC_XN = SC_DPI + NX,
                             SC_N = SC_XN + NX;
        double* sc;
        int cur;          // which P buffer holds the successor's cost-to-go
        double ng, nb, nd, nm, musum, lru;
        NMPC_HD void init(double* scratch)
        {
            sc = scratch; cur = 0;
#pragma unroll
            for (int i = SC_PV; i < SC_N; i++) sc[i * PSTRIDE] = 0.0;
            ng = nb = nd = nm = musum = lru = 0.0;
        }
    };
    // g = P * (column j of [B A]), P read from the scratch column
    template <class LL>
    NMPC_HD static void P_col_s(const double* P, const LL& l, int j, double* g)
    {
#define PP(i, m) P[pk(i, m) * PSTRIDE]
        if (j < NV) {
            const int c = j;
#pragma unroll
            for (int i = 0; i < NX; i++)
                g[i] = PP(i, 0) * l.e(0, 1 + 2 * NV + c) + PP(i, 1) * l.e(1, 1 + 2 * NV + c) + PP(i, 2) * l.e(2, 1 + 2 * NV + c)
                     + PP(i, 3 + c) * l.a_u(c) + PP(i, 3 + NV + c) * l.r_u(c);
        } else if (j < NV + 2) {
#pragma unroll
            for (int i = 0; i < NX; i++) g[i] = PP(i, j - NV);
        } else if (j == NV + 2) {
#pragma unroll
            for (int i = 0; i < NX; i++) g[i] = PP(i, 0) * l.e(0, 0) + PP(i, 1) * l.e(1, 0) + PP(i, 2) * l.e(2, 0);
        } else if (j < NV + 3 + NV) {
            const int c = j - NV - 3;
#pragma unroll
            for (int i = 0; i < NX; i++)
                g[i] = PP(i, 0) * l.e(0, 1 + c) + PP(i, 1) * l.e(1, 1 + c) + PP(i, 2) * l.e(2, 1 + c) + PP(i, 3 + c) * l.a_v(c);
        } else {
            const int c = j - NV - 3 - NV;
#pragma unroll
            for (int i = 0; i < NX; i++)
                g[i] = PP(i, 0) * l.e(0, 1 + NV + c) + PP(i, 1) * l.e(1, 1 + NV + c) + PP(i, 2) * l.e(2, 1 + NV + c)
                     + PP(i, 3 + c) * l.a_r(c) + PP(i, 3 + NV + c);
        }
#undef PP
    }

    // phase 1 of a B stage: apply the previous step, evaluate the residuals, store the iterate.
    // Leaves the stage gradient (incl. barrier terms) in gu/gx, the dynamics residual in rb and the
    // barrier Hessian terms in Gam.
    // reads : LIN[all], IT[all], ST[DZ, MC] (not when first)      writes: IT[all]
    //
    // Two formulations of the same arithmetic.  Whole stage (this one; diff, tric): every input of the stage is loaded up
    // front, one memory round trip per stage, ~400 bytes of spills that stay in L1.  Streamed (below; omni4): channel by
    // channel, few values live, several dependent round trips per stage.  Measured, 65,536 instances, K3 per step: diff
    // 34.5 ms whole / 37.7 ms streamed, tric 23.6 / 27.1, omni4 139 / 119 (whole-stage omni4 spills 1.9 KB in this half alone).
    template <bool IMAGE = false>
    NMPC_HD static void stage_B_update_whole(int k, const StageIn& in, const StageOut& out, const Tables& tb, const double* We,
                                       const IpmOpts& o, bool first, double a_step, double sigmu, double mcw, CarryB& cy,
                                       double* gu, double* gx, double* rb, double* Gam, L* lin_keep = nullptr)
    {
        using C = CarryB;
        const bool hasU = k < NSTAGE, hasX = k > 0;
        double* sc = cy.sc;
        double Hu[NV], Hx[NX], qu[NV], qx[NX];
#pragma unroll
        for (int c = 0; c < NV; c++) { Hu[c] = hasU ? tb.dt * tb.W[k * NY + NX + c] : 0.0; qu[c] = in.lin[(R::Q + c) * LANES]; }
#pragma unroll
        for (int j = 0; j < NX; j++) { Hx[j] = hasU ? tb.dt * tb.W[k * NY + j] : We[j]; qx[j] = in.lin[(R::Q + NU + j) * LANES]; }

        double zu[NV], zx[NX], pin[NX];
        double ll[NB2], lu[NB2], tl[NB2], tu[NB2], dl[NB2], du_[NB2];
#pragma unroll
        for (int b = 0; b < NB2; b++) { dl[b] = in.lin[(R::DLB + b) * LANES]; du_[b] = in.lin[(R::DUB + b) * LANES]; }
        // Every global load of the stage is issued HERE, unconditionally, before anything is computed (the fence below keeps
        // ptxas from sinking them to their uses): one memory round trip per stage.  With the loads left inside the `act` /
        // hasU / hasX branches the compiler cannot move them above the branches, and the stage waits on eight dependent
        // round trips (ncu: 59 % of the sweep's stall samples on the long scoreboard, 19 % on the four divisions by t alone).
        // Rows that a stage does not have (controls of stage N, states of stage 0) exist in the tile; what is read there is
        // discarded by the selects below.
        const int kt = hasU ? k : 0;                                   // the tables have N rows
        L lin_local;
        L& lin = lin_keep ? *lin_keep : lin_local;
        load_lin(in.lin, tb.lti + kt * 4 * NV, tb.thr + kt * NC, lin);
        double b0[NX];
#pragma unroll
        for (int i = 0; i < NX; i++) b0[i] = in.lin[(R::B0 + i) * LANES];
        double dzu[NV], dzx[NX], mcl[NB2], mcu[NB2];
        if (!first) {
#pragma unroll
            for (int c = 0; c < NV; c++) zu[c] = in.it[(R::Z + c) * LANES];
#pragma unroll
            for (int j = 0; j < NX; j++) { zx[j] = in.it[(R::Z + NU + j) * LANES]; pin[j] = in.it[(R::PI + j) * LANES]; }
#pragma unroll
            for (int b = 0; b < NB2; b++) {
                ll[b] = in.it[(R::LAM + b) * LANES]; lu[b] = in.it[(R::LAM + NB2 + b) * LANES];
                tl[b] = in.it[(R::T + b) * LANES];   tu[b] = in.it[(R::T + NB2 + b) * LANES];
            }
#if defined(__CUDA_ARCH__)
            if (IMAGE) {
                // in.st points at the lane's shared-memory image of the step rows, copied while the previous stage was
                // factorised: the loads above are in flight, the reads below wait for nothing
                NMPC_PHASE_FENCE();
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            }
#endif
#pragma unroll
            for (int c = 0; c < NV; c++) dzu[c] = in.st[(R::DZ + c) * LANES];
#pragma unroll
            for (int j = 0; j < NX; j++) dzx[j] = in.st[(R::DZ + NU + j) * LANES];
#pragma unroll
            for (int b = 0; b < NB2; b++) { mcl[b] = in.st[(R::MC + b) * LANES]; mcu[b] = in.st[(R::MC + NB2 + b) * LANES]; }
        } else {
#pragma unroll
            for (int j = 0; j < NX; j++) zx[j] = in.it[(R::Z + NU + j) * LANES];
        }
        if (!IMAGE) NMPC_PHASE_FENCE();

        double v1u[NV], v1x[NX], v2u[NV], v2x[NX];
        if (hasU) {
            double pio[NX], dpi[NX];
#pragma unroll
            for (int j = 0; j < NX; j++) { pio[j] = sc[(C::SC_PIO + j) * PSTRIDE]; dpi[j] = sc[(C::SC_DPI + j) * PSTRIDE]; }
            apply_T(lin, pio, v1u, v1x); apply_T(lin, dpi, v2u, v2x);
        } else {
#pragma unroll
            for (int c = 0; c < NV; c++) { v1u[c] = 0.0; v2u[c] = 0.0; }
#pragma unroll
            for (int j = 0; j < NX; j++) { v1x[j] = 0.0; v2x[j] = 0.0; }
        }
        double pi_old[NX];

        if (first) {
            // cold start (HPIPM INIT_VAR with warm_start = 0): z = 0, pi = 0, slacks from the
            // bounds with the thr0 projection, lam = mu0 / t
#pragma unroll
            for (int c = 0; c < NV; c++) zu[c] = 0.0;
#pragma unroll
            for (int j = 0; j < NX; j++) { if (hasX) zx[j] = 0.0; pin[j] = 0.0; pi_old[j] = 0.0; }
#pragma unroll
            for (int b = 0; b < NB2; b++) {
                const bool act = (b < NV) ? hasU : hasX;
                double zb = 0.0, t_l = -dl[b], t_u = du_[b];
                if (t_l < o.thr0) {
                    if (t_u < o.thr0) { zb = 0.5 * (dl[b] + du_[b]); t_l = o.thr0; t_u = o.thr0; }
                    else { t_l = o.thr0; zb = dl[b] + o.thr0; }
                } else if (t_u < o.thr0) { t_u = o.thr0; zb = du_[b] - o.thr0; }
                if (act) {
                    if (b < NV) zu[b] = zb; else zx[3 + b] = zb;      // ref state index 3+NV+(b-NV)
                    tl[b] = t_l; tu[b] = t_u; ll[b] = o.mu0 / t_l; lu[b] = o.mu0 / t_u;
                } else { tl[b] = 1.0; tu[b] = 1.0; ll[b] = 0.0; lu[b] = 0.0; }
            }
#pragma unroll
            for (int j = 0; j < NX; j++) sc[(C::SC_DPI + j) * PSTRIDE] = 0.0;
        } else {
            if (!hasU) {
#pragma unroll
                for (int c = 0; c < NV; c++) dzu[c] = 0.0;
            }
            if (!hasX) {
#pragma unroll
                for (int j = 0; j < NX; j++) { dzx[j] = 0.0; pin[j] = 0.0; }
            }
#pragma unroll
            for (int j = 0; j < NX; j++) pi_old[j] = pin[j];
            double ldo[NB2], dld[NB2];   // (lam_u - lam_l) old, (dlam_l - dlam_u)
#pragma unroll
            for (int b = 0; b < NB2; b++) {
                const bool act = (b < NV) ? hasU : hasX;
                if (act) {
                    const double mc_l = mcl[b], mc_u = mcu[b];
                    const double zb = (b < NV) ? zu[b] : zx[3 + b];
                    const double dzb = (b < NV) ? dzu[b] : dzx[3 + b];
                    const double rd_l = dl[b] - zb + tl[b], rd_u = -du_[b] + zb + tu[b];
                    const double rm_l = ll[b] * tl[b] - o.tau_min + mcw * mc_l - sigmu;
                    const double rm_u = lu[b] * tu[b] - o.tau_min + mcw * mc_u - sigmu;
                    const double dt_l = dzb - rd_l, dt_u = -dzb - rd_u;
                    const double dl_l = -(ll[b] * dt_l + rm_l) / tl[b];
                    const double dl_u = -(lu[b] * dt_u + rm_u) / tu[b];
                    ldo[b] = lu[b] - ll[b];
                    dld[b] = dl_l - dl_u;
                    ll[b] += a_step * dl_l; lu[b] += a_step * dl_u;
                    tl[b] += a_step * dt_l; tu[b] += a_step * dt_u;
                } else { ll[b] = 0.0; lu[b] = 0.0; tl[b] = 1.0; tu[b] = 1.0; ldo[b] = 0.0; dld[b] = 0.0; }
            }
            // stationarity residual of the Newton system, control rows (diagnostic: the quantity
            // HPIPM's iterative refinement would test)
            if (hasU) {
#pragma unroll
                for (int c = 0; c < NV; c++) {
                    const double r = qu[c] + Hu[c] * zu[c] + ldo[c] + v1u[c] + Hu[c] * dzu[c] - dld[c] + v2u[c];
                    cy.lru = fmax(cy.lru, fabs(r));
                }
            }
            // adjoint recursion for the multiplier step of the dynamics that define x_k
            if (hasX) {
#pragma unroll
                for (int j = 0; j < NX; j++) {
                    double r = qx[j] + Hx[j] * zx[j] - pin[j] + v1x[j] + Hx[j] * dzx[j] + v2x[j];
                    if (j >= 3 + NV) r += ldo[j - 3] - dld[j - 3];
                    sc[(C::SC_DPI + j) * PSTRIDE] = r;
                    pin[j] += a_step * r;
                    zx[j] += a_step * dzx[j];
                }
            }
#pragma unroll
            for (int c = 0; c < NV; c++) zu[c] += a_step * dzu[c];
        }

        // ---- residuals at the (new) iterate -----------------------------------------------
#pragma unroll
        for (int c = 0; c < NV; c++) {
            gu[c] = qu[c] + Hu[c] * zu[c] + (lu[c] - ll[c]) + (v1u[c] + a_step * v2u[c]);
            if (hasU) cy.ng = fmax(cy.ng, fabs(gu[c]));
        }
#pragma unroll
        for (int j = 0; j < NX; j++) {
            double r = qx[j] + Hx[j] * zx[j] - pin[j] + (v1x[j] + a_step * v2x[j]);
            if (j >= 3 + NV) r += lu[j - 3] - ll[j - 3];
            gx[j] = r;
            if (hasX) cy.ng = fmax(cy.ng, fabs(r));
        }
        if (hasU) {
            apply(lin, zu, zx, rb);
#pragma unroll
            for (int i = 0; i < NX; i++) {
                rb[i] += b0[i] - sc[(C::SC_XN + i) * PSTRIDE];
                cy.nb = fmax(cy.nb, fabs(rb[i]));
            }
        }
#pragma unroll
        for (int b = 0; b < NB2; b++) {
            const bool act = (b < NV) ? hasU : hasX;
            if (act) {
                const double zb = (b < NV) ? zu[b] : zx[3 + b];
                const double rd_l = dl[b] - zb + tl[b], rd_u = -du_[b] + zb + tu[b];
                const double pm_l = ll[b] * tl[b], pm_u = lu[b] * tu[b];
                cy.musum += pm_l + pm_u;
                const double rm_l = pm_l - o.tau_min, rm_u = pm_u - o.tau_min;
                cy.nd = fmax(cy.nd, fmax(fabs(rd_l), fabs(rd_u)));
                cy.nm = fmax(cy.nm, fmax(fabs(rm_l), fabs(rm_u)));
                const double ti_l = tl[b] < o.t_min ? 1.0 / o.t_min : 1.0 / tl[b];
                const double ti_u = tu[b] < o.t_min ? 1.0 / o.t_min : 1.0 / tu[b];
                const double l_l = ll[b] < o.lam_min ? o.lam_min : ll[b];
                const double l_u = lu[b] < o.lam_min ? o.lam_min : lu[b];
                Gam[b] = ti_l * l_l + ti_u * l_u;
                const double gam = ti_l * (rm_l - ll[b] * rd_l) - ti_u * (rm_u - lu[b] * rd_u);
                if (b < NV) gu[b] += gam; else gx[3 + b] += gam;
            } else Gam[b] = 0.0;
        }

        // ---- store the iterate, hand the carries to stage k-1 --------------------------------
#pragma unroll
        for (int c = 0; c < NV; c++) out.it[(R::Z + c) * LANES] = zu[c];
        if (hasX) {
#pragma unroll
            for (int j = 0; j < NX; j++) { out.it[(R::Z + NU + j) * LANES] = zx[j]; out.it[(R::PI + j) * LANES] = pin[j]; }
        }
#pragma unroll
        for (int b = 0; b < NB2; b++) {
            out.it[(R::LAM + b) * LANES] = ll[b]; out.it[(R::LAM + NB2 + b) * LANES] = lu[b];
            out.it[(R::T + b) * LANES] = tl[b];   out.it[(R::T + NB2 + b) * LANES] = tu[b];
        }
#pragma unroll
        for (int j = 0; j < NX; j++) { sc[(C::SC_PIO + j) * PSTRIDE] = pi_old[j]; sc[(C::SC_XN + j) * PSTRIDE] = zx[j]; }
    }

    // The streamed formulation: channel by channel, to keep few values live.  Everything a channel c = {control c,
    // actual c, ref c} needs from the rest of the stage is the pose part of the successor's multipliers (p0, d0) and, for
    // the dynamics residual, three running sums of the pose rows of [A B] z.  One bounded component (control or reference state):
    struct Bnd { double ldo, dld, lnew, gam, Gam, znew; };
    // The rows of one bounded component / one state component, loaded unconditionally at the top of its channel (rows that a
    // stage does not have exist in the tile and are discarded): left inside the `act` / hasX branches, every component waited
    // for a memory round trip of its own.
    struct BndRows { double dl, du, ll, lu, tl, tu, mcl, mcu; };
    struct StRows { double q, z, dz, pi; };
    NMPC_HD static void load_bound_rows(int b, bool first, const StageInR& in, BndRows& w)
    {
        w.dl = in.lin[(R::DLB + b) * LANES]; w.du = in.lin[(R::DUB + b) * LANES];
        w.ll = w.lu = w.tl = w.tu = w.mcl = w.mcu = 0.0;
        if (!first) {
            w.ll = in.it[(R::LAM + b) * LANES]; w.lu = in.it[(R::LAM + NB2 + b) * LANES];
            w.tl = in.it[(R::T + b) * LANES];   w.tu = in.it[(R::T + NB2 + b) * LANES];
            w.mcl = in.st[(R::MC + b) * LANES]; w.mcu = in.st[(R::MC + NB2 + b) * LANES];
        }
    }
    NMPC_HD static void load_state_rows(int j, bool first, const StageInR& in, StRows& w)
    {
        w.q = in.lin[(R::Q + NU + j) * LANES];
        w.z = in.it[(R::Z + NU + j) * LANES];
        w.dz = 0.0; w.pi = 0.0;
        if (!first) { w.dz = in.st[(R::DZ + NU + j) * LANES]; w.pi = in.it[(R::PI + j) * LANES]; }
    }
    NMPC_HD static void bound_component(bool act, bool first, int b, double zb, double dzb, const BndRows& w, const StageOutR& out,
                                        const IpmOpts& o, double a_step, double sigmu, double mcw, CarryB& cy, Bnd& r)
    {
        double ll, lu, tl, tu;
        r.ldo = 0.0; r.dld = 0.0; r.znew = zb;
        if (!act) {
            ll = 0.0; lu = 0.0; tl = 1.0; tu = 1.0;
            r.lnew = 0.0; r.gam = 0.0; r.Gam = 0.0;
            if (first) r.znew = 0.0;
        } else {
            const double dl = w.dl, du_ = w.du;
            if (first) {
                // cold start (HPIPM INIT_VAR with warm_start = 0): slacks from the bounds with the thr0 projection, lam = mu0 / t
                double z0 = 0.0, t_l = -dl, t_u = du_;
                if (t_l < o.thr0) {
                    if (t_u < o.thr0) { z0 = 0.5 * (dl + du_); t_l = o.thr0; t_u = o.thr0; }
                    else { t_l = o.thr0; z0 = dl + o.thr0; }
                } else if (t_u < o.thr0) { t_u = o.thr0; z0 = du_ - o.thr0; }
                r.znew = z0; tl = t_l; tu = t_u; ll = o.mu0 / t_l; lu = o.mu0 / t_u;
            } else {
                ll = w.ll; lu = w.lu;
                tl = w.tl; tu = w.tu;
                const double mc_l = w.mcl, mc_u = w.mcu;
                const double rd_l = dl - zb + tl, rd_u = -du_ + zb + tu;
                const double rm_l = ll * tl - o.tau_min + mcw * mc_l - sigmu;
                const double rm_u = lu * tu - o.tau_min + mcw * mc_u - sigmu;
                const double dt_l = dzb - rd_l, dt_u = -dzb - rd_u;
                const double dl_l = -(ll * dt_l + rm_l) / tl;
                const double dl_u = -(lu * dt_u + rm_u) / tu;
                r.ldo = lu - ll;
                r.dld = dl_l - dl_u;
                ll += a_step * dl_l; lu += a_step * dl_u;
                tl += a_step * dt_l; tu += a_step * dt_u;
                r.znew = zb + a_step * dzb;
            }
            // residuals at the new iterate
            const double zn = r.znew;
            const double rd_l = dl - zn + tl, rd_u = -du_ + zn + tu;
            const double pm_l = ll * tl, pm_u = lu * tu;
            cy.musum += pm_l + pm_u;
            const double rm_l = pm_l - o.tau_min, rm_u = pm_u - o.tau_min;
            cy.nd = fmax(cy.nd, fmax(fabs(rd_l), fabs(rd_u)));
            cy.nm = fmax(cy.nm, fmax(fabs(rm_l), fabs(rm_u)));
            const double ti_l = tl < o.t_min ? 1.0 / o.t_min : 1.0 / tl;
            const double ti_u = tu < o.t_min ? 1.0 / o.t_min : 1.0 / tu;
            const double l_l = ll < o.lam_min ? o.lam_min : ll;
            const double l_u = lu < o.lam_min ? o.lam_min : lu;
            r.Gam = ti_l * l_l + ti_u * l_u;
            r.gam = ti_l * (rm_l - ll * rd_l) - ti_u * (rm_u - lu * rd_u);
            r.lnew = lu - ll;
        }
        out.it[(R::LAM + b) * LANES] = ll; out.it[(R::LAM + NB2 + b) * LANES] = lu;
        out.it[(R::T + b) * LANES] = tl;   out.it[(R::T + NB2 + b) * LANES] = tu;
    }
    // one state component j: stationarity residual of the Newton system (= multiplier step of the dynamics that define x_k,
    // the adjoint recursion), new iterate, gradient at the new iterate.  v1 / v2: component j of [B A]' (old multipliers /
    // multiplier step of the successor).  bn: the bounded-component terms (reference states), else nullptr.
    NMPC_HD static double state_component(int k, int j, bool first, double v1, double v2, double H, const Bnd* bn,
                                          const StRows& w, const StageOutR& out, double a_step, CarryB& cy, double& gxj)
    {
        using C = CarryB;
        const bool hasX = k > 0;
        double* sc = cy.sc;
        const double qx = w.q;
        double zx, pin = 0.0, pi_old = 0.0;
        if (first) {
            zx = hasX ? (bn ? bn->znew : 0.0) : w.z;
            sc[(C::SC_DPI + j) * PSTRIDE] = 0.0;
        } else {
            zx = w.z;
            if (hasX) {
                const double dzx = w.dz;
                pin = w.pi;
                pi_old = pin;
                double r = qx + H * zx - pin + v1 + H * dzx + v2;
                if (bn) r += bn->ldo - bn->dld;
                sc[(C::SC_DPI + j) * PSTRIDE] = r;
                pin += a_step * r;
                zx = bn ? bn->znew : zx + a_step * dzx;
            }
        }
        double g = qx + H * zx - pin + (v1 + a_step * v2);
        if (bn) g += bn->lnew;
        if (hasX) cy.ng = fmax(cy.ng, fabs(g));
        if (bn) g += bn->gam;
        gxj = g;
        if (hasX) { out.it[(R::Z + NU + j) * LANES] = zx; out.it[(R::PI + j) * LANES] = pin; }
        sc[(C::SC_PIO + j) * PSTRIDE] = pi_old;
        return zx;
    }
    NMPC_HD static void stage_B_update_streamed(int k, const StageIn& in_, const StageOut& out_, const Tables& tb, const double* We,
                                       const IpmOpts& o, bool first, double a_step, double sigmu, double mcw, CarryB& cy,
                                       double* gu, double* gx, double* rb, double* Gam)
    {
        using C = CarryB;
        const bool hasU = k < NSTAGE, hasX = k > 0;
        double* sc = cy.sc;
        // no field of the stage is read after it has been written: the loads may move above the stores
        const StageInR in = {in_.lin, in_.it, in_.st};
        const StageOutR out = {out_.it};
        const double* lti = tb.lti + (hasU ? k : 0) * 4 * NV;
        const double* thr = tb.thr + (hasU ? k : 0) * NC;
        const double* Wk = tb.W + (hasU ? k : 0) * NY;
        // entry (i, c) of the pose rows of [A | B] (columns theta | actual | ref | u), loaded whether or not the stage has it;
        // nothing is linearised at stage N
#define EL_(i, c) ((i) < R::ER ? in.lin[(R::E + (i) * NC + (c)) * LANES] : thr[c])
        // pose part of the successor's carries; the successor's state for the dynamics residual
        double p0[3], d0[3], xo[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            p0[i] = sc[(C::SC_PIO + i) * PSTRIDE]; d0[i] = sc[(C::SC_DPI + i) * PSTRIDE]; xo[i] = sc[(C::SC_XN + i) * PSTRIDE];
        }
        // ---- pose components x, y, theta: unit columns for x and y ---------------------------------------------------------
        double acc[3];                       // pose rows of [A B] z at the new iterate
        double b0p[3];
        {
            StRows w[3];
#pragma unroll
            for (int j = 0; j < 3; j++) { load_state_rows(j, first, in, w[j]); b0p[j] = in.lin[(R::B0 + j) * LANES]; }
            const double l0 = EL_(0, 0), l1 = EL_(1, 0), l2 = EL_(2, 0);
            NMPC_PHASE_FENCE();
            const double e0 = hasU ? l0 : 0.0, e1 = hasU ? l1 : 0.0, e2 = hasU ? l2 : 0.0;
            double zp[3];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double v1 = hasU ? (j < 2 ? p0[j] : e0 * p0[0] + e1 * p0[1] + e2 * p0[2]) : 0.0;
                const double v2 = hasU ? (j < 2 ? d0[j] : e0 * d0[0] + e1 * d0[1] + e2 * d0[2]) : 0.0;
                const double H = hasU ? tb.dt * Wk[j] : We[j];
                zp[j] = state_component(k, j, first, v1, v2, H, nullptr, w[j], out, a_step, cy, gx[j]);
            }
            acc[0] = zp[0] + e0 * zp[2]; acc[1] = zp[1] + e1 * zp[2]; acc[2] = e2 * zp[2];
#pragma unroll
            for (int i = 0; i < 3; i++) sc[(C::SC_XN + i) * PSTRIDE] = zp[i];
        }
        // ---- the channels ---------------------------------------------------------------------------------------------------
#pragma unroll
        for (int c = 0; c < NV; c++) {
            NMPC_PHASE_FENCE();
            const int ja = 3 + c, jr = 3 + NV + c;
            // every row of the channel, ahead of its arithmetic
            const int qc = 1 + 2 * NV + c, qa = 1 + c, qr = 1 + NV + c;
            const double lc0 = EL_(0, qc), lc1 = EL_(1, qc), lc2 = EL_(2, qc);
            const double la0 = EL_(0, qa), la1 = EL_(1, qa), la2 = EL_(2, qa);
            const double lr0 = EL_(0, qr), lr1 = EL_(1, qr), lr2 = EL_(2, qr);
            const double qu = in.lin[(R::Q + c) * LANES];
            const double zu0 = in.it[(R::Z + c) * LANES];
            const double dzu0 = first ? 0.0 : in.st[(R::DZ + c) * LANES];
            BndRows wu, wr;
            load_bound_rows(c, first, in, wu);
            load_bound_rows(NV + c, first, in, wr);
            StRows wa, ws;
            load_state_rows(ja, first, in, wa);
            load_state_rows(jr, first, in, ws);
            const double b0a = in.lin[(R::B0 + ja) * LANES], b0r = in.lin[(R::B0 + jr) * LANES];
            NMPC_PHASE_FENCE();
            const double pa = sc[(C::SC_PIO + ja) * PSTRIDE], pr = sc[(C::SC_PIO + jr) * PSTRIDE];
            const double da = sc[(C::SC_DPI + ja) * PSTRIDE], dr = sc[(C::SC_DPI + jr) * PSTRIDE];
            const double xa = sc[(C::SC_XN + ja) * PSTRIDE], xr = sc[(C::SC_XN + jr) * PSTRIDE];
            const double av = hasU ? lti[c] : 0.0, ar = hasU ? lti[NV + c] : 0.0, au = hasU ? lti[2 * NV + c] : 0.0, ru = hasU ? lti[3 * NV + c] : 0.0;
            // control c
            double zun = 0.0;
            {
                const double e0 = hasU ? lc0 : 0.0, e1 = hasU ? lc1 : 0.0, e2 = hasU ? lc2 : 0.0;
                const double v1 = e0 * p0[0] + e1 * p0[1] + e2 * p0[2] + au * pa + ru * pr;
                const double v2 = e0 * d0[0] + e1 * d0[1] + e2 * d0[2] + au * da + ru * dr;
                const double H = hasU ? tb.dt * Wk[NX + c] : 0.0;
                const double zu = first ? 0.0 : zu0;
                const double dzu = (!first && hasU) ? dzu0 : 0.0;
                Bnd bn;
                bound_component(hasU, first, c, zu, dzu, wu, out, o, a_step, sigmu, mcw, cy, bn);
                if (!first && hasU) {
                    // stationarity residual of the Newton system, control rows (diagnostic: the quantity HPIPM's
                    // iterative refinement would test)
                    const double r = qu + H * zu + bn.ldo + v1 + H * dzu - bn.dld + v2;
                    cy.lru = fmax(cy.lru, fabs(r));
                }
                zun = bn.znew;
                double g = qu + H * zun + bn.lnew + (v1 + a_step * v2);
                if (hasU) cy.ng = fmax(cy.ng, fabs(g));
                gu[c] = g + bn.gam;
                Gam[c] = bn.Gam;
                out.it[(R::Z + c) * LANES] = zun;
                acc[0] += e0 * zun; acc[1] += e1 * zun; acc[2] += e2 * zun;
            }
            // actual c
            double zan;
            {
                const double e0 = hasU ? la0 : 0.0, e1 = hasU ? la1 : 0.0, e2 = hasU ? la2 : 0.0;
                const double v1 = e0 * p0[0] + e1 * p0[1] + e2 * p0[2] + av * pa;
                const double v2 = e0 * d0[0] + e1 * d0[1] + e2 * d0[2] + av * da;
                const double H = hasU ? tb.dt * Wk[ja] : We[ja];
                zan = state_component(k, ja, first, v1, v2, H, nullptr, wa, out, a_step, cy, gx[ja]);
                acc[0] += e0 * zan; acc[1] += e1 * zan; acc[2] += e2 * zan;
            }
            // reference c (bounded)
            double zrn;
            {
                const double e0 = hasU ? lr0 : 0.0, e1 = hasU ? lr1 : 0.0, e2 = hasU ? lr2 : 0.0;
                const double v1 = e0 * p0[0] + e1 * p0[1] + e2 * p0[2] + ar * pa + pr;
                const double v2 = e0 * d0[0] + e1 * d0[1] + e2 * d0[2] + ar * da + dr;
                const double H = hasU ? tb.dt * Wk[jr] : We[jr];
                const double zr = (first || !hasX) ? 0.0 : ws.z;
                const double dzr = (first || !hasX) ? 0.0 : ws.dz;
                Bnd bn;
                bound_component(hasX, first, NV + c, zr, dzr, wr, out, o, a_step, sigmu, mcw, cy, bn);
                zrn = state_component(k, jr, first, v1, v2, H, hasX ? &bn : nullptr, ws, out, a_step, cy, gx[jr]);
                Gam[NV + c] = bn.Gam;
                acc[0] += e0 * zrn; acc[1] += e1 * zrn; acc[2] += e2 * zrn;
            }
            // the channel's two rows of the dynamics residual
            if (hasU) {
                rb[ja] = av * zan + ar * zrn + au * zun + (b0a - xa);
                rb[jr] = zrn + ru * zun + (b0r - xr);
                cy.nb = fmax(cy.nb, fmax(fabs(rb[ja]), fabs(rb[jr])));
            }
            sc[(C::SC_XN + ja) * PSTRIDE] = zan; sc[(C::SC_XN + jr) * PSTRIDE] = zrn;
        }
        if (hasU) {
#pragma unroll
            for (int i = 0; i < 3; i++) {
                rb[i] = acc[i] + (b0p[i] - xo[i]);
                cy.nb = fmax(cy.nb, fabs(rb[i]));
            }
        }
#undef EL_
    }

    NMPC_HD static void stage_B_update(int k, const StageIn& in, const StageOut& out, const Tables& tb, const double* We,
                                       const IpmOpts& o, bool first, double a_step, double sigmu, double mcw, CarryB& cy,
                                       double* gu, double* gx, double* rb, double* Gam)
    {
        if (LEAN) stage_B_update_streamed(k, in, out, tb, We, o, first, a_step, sigmu, mcw, cy, gu, gx, rb, Gam);
        else stage_B_update_whole<false>(k, in, out, tb, We, o, first, a_step, sigmu, mcw, cy, gu, gx, rb, Gam);
    }

    // phase 2 of a B stage: one step of the Riccati recursion.
    // reads : LIN[E]      writes: FA[LUU,KH,LH,RB]
    NMPC_HD static void stage_B_riccati(int k, const StageIn& in, const StageOut& out, const Tables& tb, const double* We,
                                        const IpmOpts& o, CarryB& cy, double* gu, double* gx, const double* rb, const double* Gam,
                                        const L* lin_kept = nullptr)
    {
        using C = CarryB;
        const bool hasU = k < NSTAGE, hasX = k > 0;
        double* sc = cy.sc;
        const double* P = sc + (size_t)C::SC_P * PSTRIDE;         // successor's cost-to-go ...
        double* Pn = sc + (size_t)C::SC_P * PSTRIDE;              // ... overwritten in place by this stage's
        if (hasU) {
            typename LinSel<LEAN>::type lin;
            if (!LEAN && lin_kept) lin = *reinterpret_cast<const typename LinSel<LEAN>::type*>(lin_kept);   // the update half's view, kept in registers
            else load_lin(in.lin, tb.lti + k * 4 * NV, tb.thr + k * NC, lin);
#pragma unroll
            for (int i = 0; i < NX; i++) out.fa[(R::RB + i) * LANES] = rb[i];
            {
                double Pb[NX];
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    double s = sc[(C::SC_PV + i) * PSTRIDE];
#pragma unroll
                    for (int j = 0; j < NX; j++) s += P[pk(i, j) * PSTRIDE] * rb[j];
                    Pb[i] = s;
                }
                double tu_[NV], tx_[NX];
                apply_T(lin, Pb, tu_, tx_);
#pragma unroll
                for (int c = 0; c < NV; c++) gu[c] += tu_[c];
#pragma unroll
                for (int j = 0; j < NX; j++) {
                    gx[j] += tx_[j];
                    if (LEAN) sc[(C::SC_PV + j) * PSTRIDE] = gx[j];      // the successor's gradient is consumed: park this stage's in its slot
                }
            }
            // M = [B A]' P [B A] + diag(H + Gamma + reg), built column by column in z order [u; x].
            // Control columns first: Muu (packed lower) and Mxu, then the Cholesky of Muu; the state
            // columns are folded straight into this stage's P (Schur complement), one at a time.
            double Muu[NLU];
            double Kh[NV][NX];
#pragma unroll
            for (int a = 0; a < NV; a++) {
                if (RICCATI_FENCES) NMPC_PHASE_FENCE();
                double g[NX], cu[NV], cx[NX];
                P_col_s(P, lin, a, g);
                apply_T(lin, g, cu, cx);
#pragma unroll
                for (int b = a; b < NV; b++) Muu[b * (b + 1) / 2 + a] = cu[b];
#pragma unroll
                for (int j = 0; j < NX; j++) Kh[a][j] = cx[j];
            }
#pragma unroll
            for (int c = 0; c < NV; c++) Muu[c * (c + 1) / 2 + c] += tb.dt * tb.W[k * NY + NX + c] + o.reg_prim + Gam[c];
            // Cholesky of the control block, diagonal kept inverted
            double Luu[NLU];
#pragma unroll
            for (int a = 0; a < NV; a++) {
                double d = Muu[a * (a + 1) / 2 + a];
#pragma unroll
                for (int c = 0; c < a; c++) d -= Luu[a * (a + 1) / 2 + c] * Luu[a * (a + 1) / 2 + c];
                const double inv = d > 0.0 ? 1.0 / sqrt(d) : 0.0;
                Luu[a * (a + 1) / 2 + a] = inv;
#pragma unroll
                for (int b = a + 1; b < NV; b++) {
                    double s = Muu[b * (b + 1) / 2 + a];
#pragma unroll
                    for (int c = 0; c < a; c++) s -= Luu[b * (b + 1) / 2 + c] * Luu[a * (a + 1) / 2 + c];
                    Luu[b * (b + 1) / 2 + a] = s * inv;
                }
            }
#pragma unroll
            for (int i = 0; i < NLU; i++) out.fa[(R::LUU + i) * LANES] = Luu[i];
            double lh[NV];
#pragma unroll
            for (int a = 0; a < NV; a++) {
                double s = gu[a];
#pragma unroll
                for (int c = 0; c < a; c++) s -= Luu[a * (a + 1) / 2 + c] * lh[c];
                lh[a] = s * Luu[a * (a + 1) / 2 + a];
                out.fa[(R::LH + a) * LANES] = lh[a];
            }
            if (hasX) {
#pragma unroll
                for (int j = 0; j < NX; j++) {
                    double kj[NV];
#pragma unroll
                    for (int a = 0; a < NV; a++) {
                        double s = Kh[a][j];
#pragma unroll
                        for (int c = 0; c < a; c++) s -= Luu[a * (a + 1) / 2 + c] * kj[c];
                        kj[a] = s * Luu[a * (a + 1) / 2 + a];
                        Kh[a][j] = kj[a];
                        out.fa[(R::KH + a * NX + j) * LANES] = kj[a];
                    }
                }
                // last column first: column j reads only columns <= j of the successor's P (see CarryB)
#pragma unroll
                for (int j = NX - 1; j >= 0; j--) {
                    if (RICCATI_FENCES) NMPC_PHASE_FENCE();
                    double g[NX], cu[NV], cx[NX];
                    P_col_s(P, lin, NV + j, g);
                    apply_T(lin, g, cu, cx);
                    double kcj[NV];
#pragma unroll
                    for (int a = 0; a < NV; a++) kcj[a] = Kh[a][j];
#pragma unroll
                    for (int i = j; i < NX; i++) {
                        double s = cx[i];
#pragma unroll
                        for (int a = 0; a < NV; a++) s -= Kh[a][i] * kcj[a];
                        if (i == j) s += tb.dt * tb.W[k * NY + i] + o.reg_prim + (i >= 3 + NV ? Gam[i - 3] : 0.0);
                        Pn[pk(i, j) * PSTRIDE] = s;
                    }
                }
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    double s = LEAN ? sc[(C::SC_PV + i) * PSTRIDE] : gx[i];
#pragma unroll
                    for (int a = 0; a < NV; a++) s -= Kh[a][i] * lh[a];
                    sc[(C::SC_PV + i) * PSTRIDE] = s;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < NX; i++) {
#pragma unroll
                for (int j = 0; j <= i; j++) Pn[pk(i, j) * PSTRIDE] = (i == j) ? We[i] + o.reg_prim + (i >= 3 + NV ? Gam[i - 3] : 0.0) : 0.0;
                sc[(C::SC_PV + i) * PSTRIDE] = gx[i];
            }
        }
    }

#if defined(__CUDA_ARCH__)
    // nf rows of the stage BELOW the one `rows_k` points at (stage k - 1: one record further down in its group) pulled into L2;
    // the active lanes of the warp split the 128-byte lines.  Issued between the two halves of a factorising stage, so the
    // next stage's rows cross HBM while this one's Riccati half computes.
    __device__ __forceinline__ static void prefetch_rows_prev(const double* rows_k, int nf)
    {
        const unsigned m = __activemask();
        const int lane = threadIdx.x & (LANES - 1);
        const int n = __popc(m), r = __popc(m & ((1u << lane) - 1u));
        const char* p = reinterpret_cast<const char*>(rows_k - lane - (size_t)nf * LANES);
        const int lines = nf * (LANES * (int)sizeof(double) / 128);
        for (int i = r; i < lines; i += n) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (size_t)i * 128));
    }
#endif
    NMPC_HD static void stage_B(int k, const StageIn& in, const StageOut& out, const Tables& tb, const double* We,
                                const IpmOpts& o, bool first, double a_step, double sigmu, double mcw, CarryB& cy)
    {
        double gu[NV], gx[NX], rb[NX], Gam[NB2];
#if defined(__CUDA_ARCH__) && NMPC_B_KEEP_LIN
        if (!LEAN) {
            L lin;                        // the rows of [A B] stay in registers from the update half to the Riccati half
            stage_B_update_whole<false>(k, in, out, tb, We, o, first, a_step, sigmu, mcw, cy, gu, gx, rb, Gam, &lin);
            NMPC_PHASE_FENCE();
#if NMPC_B_L2_PREFETCH
            if (k > 0) {
                prefetch_rows_prev(in.lin, R::NF_LIN);
                if (!first) { prefetch_rows_prev(in.it, R::NF_IT); prefetch_rows_prev(in.st, R::NF_ST); }
            }
#endif
            stage_B_riccati(k, in, out, tb, We, o, cy, gu, gx, rb, Gam, &lin);
            return;
        }
#endif
        stage_B_update(k, in, out, tb, We, o, first, a_step, sigmu, mcw, cy, gu, gx, rb, Gam);
        NMPC_PHASE_FENCE();
#if defined(__CUDA_ARCH__) && NMPC_B_L2_PREFETCH
        if (k > 0) {
            prefetch_rows_prev(in.lin, R::NF_LIN);
            if (!first) { prefetch_rows_prev(in.it, R::NF_IT); prefetch_rows_prev(in.st, R::NF_ST); }
        }
#endif
        stage_B_riccati(k, in, out, tb, We, o, cy, gu, gx, rb, Gam);
    }

    // ---- forward sweeps.  delta == false: predictor (writes ST.DZA, ST.MC); delta == true: adds
    // the delta step to the predictor (writes ST.DZ).  Accumulates the ratio-test step (negated,
    // as HPIPM keeps it) and the three sums of mu(alpha) = (S0 + alpha S1 + alpha^2 S2) / nc.
    struct CarryF {
        double dx[NX], alpha, S0, S1, S2;
        NMPC_HD void init()
        {
#pragma unroll
            for (int j = 0; j < NX; j++) dx[j] = 0.0;
            alpha = -1.0; S0 = S1 = S2 = 0.0;
        }
    };
    // predictor reads: FA[all], LIN[E,DLB,DUB], IT[T,LAM,Z]            writes ST[MC,DZA]
    // delta     reads: FA[LUU,KH,LHD], LIN[E,DLB,DUB], IT[T,LAM,Z], ST[MC,DZA]   writes ST[DZ]
    NMPC_HD static void stage_F(int k, const StageIn& in, const StageOut& out, const Tables& tb, const IpmOpts& o,
                                bool delta, double sigmu, double mcw, CarryF& cy)
    {
        const bool hasU = k < NSTAGE, hasX = k > 0;
        // the constraint rows of the stage are loaded unconditionally, ahead of the arithmetic (rows that the stage does not
        // have exist in the tile and are discarded): inside the `act` branches every component waited for its own round trip
        double ll_[NB2], lu_[NB2], tl_[NB2], tu_[NB2], zb_[NB2], dlb_[NB2], dub_[NB2], mcl_[NB2], mcu_[NB2];
        auto load_rows = [&](int b) {
            ll_[b] = in.it[(R::LAM + b) * LANES]; lu_[b] = in.it[(R::LAM + NB2 + b) * LANES];
            tl_[b] = in.it[(R::T + b) * LANES];   tu_[b] = in.it[(R::T + NB2 + b) * LANES];
            zb_[b] = (b < NV) ? in.it[(R::Z + b) * LANES] : in.it[(R::Z + NU + 3 + b) * LANES];
            dlb_[b] = in.lin[(R::DLB + b) * LANES]; dub_[b] = in.lin[(R::DUB + b) * LANES];
            if (delta) { mcl_[b] = in.st[(R::MC + b) * LANES]; mcu_[b] = in.st[(R::MC + NB2 + b) * LANES]; }
        };
        // two channels: all rows at the top of the stage; four channels (72 values against the 128 registers of the solve
        // sweeps): the control rows and the reference-state rows as two groups, each ahead of its own arithmetic
        constexpr bool HOIST = !LEAN;
        if (HOIST) {
#pragma unroll
            for (int b = 0; b < NB2; b++) load_rows(b);
        }
        double du[NV];
#pragma unroll
        for (int c = 0; c < NV; c++) du[c] = 0.0;
        if (hasU) solve_u(in.fa, delta ? R::LHD : R::LH, hasX, cy.dx, du);
        double dzu[NV], dzx[NX];
        if (!delta) {
#pragma unroll
            for (int c = 0; c < NV; c++) { dzu[c] = du[c]; out.st[(R::DZA + c) * LANES] = du[c]; }
#pragma unroll
            for (int j = 0; j < NX; j++) { dzx[j] = cy.dx[j]; out.st[(R::DZA + NU + j) * LANES] = cy.dx[j]; }
        } else {
#pragma unroll
            for (int c = 0; c < NV; c++) { dzu[c] = in.st[(R::DZA + c) * LANES] + du[c]; out.st[(R::DZ + c) * LANES] = dzu[c]; }
#pragma unroll
            for (int j = 0; j < NX; j++) { dzx[j] = in.st[(R::DZA + NU + j) * LANES] + cy.dx[j]; out.st[(R::DZ + NU + j) * LANES] = dzx[j]; }
        }
#pragma unroll
        for (int b = 0; b < NB2; b++) {
            const bool act = (b < NV) ? hasU : hasX;
            if (!HOIST && (b == 0 || b == NV)) {
#pragma unroll
                for (int q = 0; q < NV; q++) load_rows(b + q);
            }
            if (act) {
                const double ll = ll_[b], lu = lu_[b];
                const double tl = tl_[b], tu = tu_[b];
                const double zb = zb_[b];
                const double dzb = (b < NV) ? dzu[b] : dzx[3 + b];
                const double rd_l = dlb_[b] - zb + tl, rd_u = -dub_[b] + zb + tu;
                double rm_l = ll * tl - o.tau_min, rm_u = lu * tu - o.tau_min;
                if (delta) {
                    rm_l += mcw * mcl_[b] - sigmu;
                    rm_u += mcw * mcu_[b] - sigmu;
                }
                const double dt_l = dzb - rd_l, dt_u = -dzb - rd_u;
                const double dl_l = -(ll * dt_l + rm_l) / tl, dl_u = -(lu * dt_u + rm_u) / tu;
                if (!delta) { out.st[(R::MC + b) * LANES] = dt_l * dl_l; out.st[(R::MC + NB2 + b) * LANES] = dt_u * dl_u; }
                if (cy.alpha * dl_l > ll) cy.alpha = ll / dl_l;
                if (cy.alpha * dt_l > tl) cy.alpha = tl / dt_l;
                if (cy.alpha * dl_u > lu) cy.alpha = lu / dl_u;
                if (cy.alpha * dt_u > tu) cy.alpha = tu / dt_u;
                cy.S0 += ll * tl + lu * tu;
                cy.S1 += ll * dt_l + tl * dl_l + lu * dt_u + tu * dl_u;
                cy.S2 += dl_l * dt_l + dl_u * dt_u;
            }
        }
        if (hasU) {
            L lin;
            load_lin(in.lin, tb.lti + k * 4 * NV, tb.thr + k * NC, lin);
            double xnew[NX];
            apply(lin, du, cy.dx, xnew);
#pragma unroll
            for (int j = 0; j < NX; j++) cy.dx[j] = xnew[j] + (delta ? 0.0 : in.fa[(R::RB + j) * LANES]);
        }
    }

    // ---- delta backward sweep: rhs only in the complementarity rows ---------------------------
    struct CarryD {
        double dp[NX];
        NMPC_HD void init()
        {
#pragma unroll
            for (int j = 0; j < NX; j++) dp[j] = 0.0;
        }
    };
    // reads: ST[MC], IT[T], FA[LUU,KH], LIN[E]      writes FA[LHD]
    NMPC_HD static void stage_Bd(int k, const StageIn& in, const StageOut& out, const Tables& tb, double sigmu, double mcw, CarryD& cy)
    {
        const bool hasU = k < NSTAGE, hasX = k > 0;
        // the constraint rows are loaded unconditionally ahead of the arithmetic, as in stage_F
        double tl_[NB2], tu_[NB2], mcl_[NB2], mcu_[NB2];
        auto load_rows = [&](int b) {
            tl_[b] = in.it[(R::T + b) * LANES]; tu_[b] = in.it[(R::T + NB2 + b) * LANES];
            mcl_[b] = in.st[(R::MC + b) * LANES]; mcu_[b] = in.st[(R::MC + NB2 + b) * LANES];
        };
        constexpr bool HOIST = !LEAN;
        double luu_[NLU], kh_[NV * NX];
        if (HOIST) {
#pragma unroll
            for (int b = 0; b < NB2; b++) load_rows(b);
#pragma unroll
            for (int i = 0; i < NLU; i++) luu_[i] = in.fa[(R::LUU + i) * LANES];
#pragma unroll
            for (int i = 0; i < NV * NX; i++) kh_[i] = in.fa[(R::KH + i) * LANES];
        }
        double qu[NV], qx[NX];
#pragma unroll
        for (int c = 0; c < NV; c++) qu[c] = 0.0;
#pragma unroll
        for (int j = 0; j < NX; j++) qx[j] = 0.0;
        if (hasU) {
            L lin;
            load_lin(in.lin, tb.lti + k * 4 * NV, tb.thr + k * NC, lin);
            apply_T(lin, cy.dp, qu, qx);
        }
#pragma unroll
        for (int b = 0; b < NB2; b++) {
            const bool act = (b < NV) ? hasU : hasX;
            if (!HOIST && (b == 0 || b == NV)) {
#pragma unroll
                for (int q = 0; q < NV; q++) load_rows(b + q);
            }
            if (act) {
                const double tl = tl_[b], tu = tu_[b];
                const double g = (mcw * mcl_[b] - sigmu) / tl - (mcw * mcu_[b] - sigmu) / tu;
                if (b < NV) qu[b] += g; else qx[3 + b] += g;
            }
        }
        if (hasU) {
            double lh[NV];
#pragma unroll
            for (int a = 0; a < NV; a++) {
                double s = qu[a];
#pragma unroll
                for (int c = 0; c < a; c++) s -= (HOIST ? luu_[a * (a + 1) / 2 + c] : in.fa[(R::LUU + a * (a + 1) / 2 + c) * LANES]) * lh[c];
                lh[a] = s * (HOIST ? luu_[a * (a + 1) / 2 + a] : in.fa[(R::LUU + a * (a + 1) / 2 + a) * LANES]);
                out.fa[(R::LHD + a) * LANES] = lh[a];
            }
            if (hasX) {
#pragma unroll
                for (int j = 0; j < NX; j++) {
                    double s = qx[j];
#pragma unroll
                    for (int a = 0; a < NV; a++) s -= (HOIST ? kh_[a * NX + j] : in.fa[(R::KH + a * NX + j) * LANES]) * lh[a];
                    cy.dp[j] = s;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < NX; j++) cy.dp[j] = qx[j];
        }
    }

    // ------------------------------------------------------------------------------------
    // K3 control: the interior-point loop of one lane, cut into the phases that sit between the
    // horizon sweeps.  One IPM iteration =
    //     F  (predictor forward)  -> after_F   : affine step length, sigma*mu
    //     Bd (delta backward)
    //     Fd (delta forward)      -> after_Fd  : step length, conditional-centering test
    //     [Bd, Fd with mcw = 0 for the lanes that fell back -> after_Fd_fallback]
    //     B  (apply step + residuals + factorise; before_B gives the damped step) -> after_B : exit test
    // The device runs two kernels per iteration over all tiles (k_sweep in rti_kernels.cu): FDF =
    // the three solve sweeps (light on registers, high occupancy) and B (the factorising sweep);
    // qp_ipm_lane below strings the same phases together for one lane (host emulation).
    // ------------------------------------------------------------------------------------
    enum { SW_B_FIRST = 0, SW_B = 1, SW_F = 2, SW_BD = 3, SW_FD = 4, SW_FDF = 5 };

    struct LaneCtl {
        double nrm[4], mu, alpha, sigmu, mu_aff0, lin_res, mcw;
        int done, iter, status, fb, nfb;
        NMPC_HD void init(bool active)
        {
            nrm[0] = nrm[1] = nrm[2] = nrm[3] = 0.0; mu = 0.0; alpha = 1.0; sigmu = 0.0; mu_aff0 = 0.0; lin_res = 0.0; mcw = 1.0;
            done = active ? 0 : 1; iter = 0; status = 0; fb = 0; nfb = 0;
        }
    };

    // exit test of HPIPM's main loop, evaluated right after the residuals
    NMPC_HD static void after_B(LaneCtl& c, const CarryB& cy, const IpmOpts& o, bool first)
    {
        c.nrm[0] = cy.ng; c.nrm[1] = cy.nb; c.nrm[2] = cy.nd; c.nrm[3] = cy.nm; c.mu = cy.musum / (double)NCON;
        if (!first) c.lin_res = fmax(c.lin_res, cy.lru);
        const bool more = c.iter < o.iter_max && c.alpha > o.alpha_min &&
                          (c.nrm[0] > o.res_g_max || c.nrm[1] > o.res_b_max || c.nrm[2] > o.res_d_max ||
                           fabs(c.nrm[3] - o.tau_min) > o.res_m_max);
        if (!more || c.mu != c.mu) {
            c.done = 1;
            c.status = (c.mu != c.mu) ? 3 : (c.iter >= o.iter_max ? 1 : (c.alpha <= o.alpha_min ? 2 : 0));
        }
    }
    NMPC_HD static void after_F(LaneCtl& c, const CarryF& cy, const IpmOpts& o)
    {
        const double a_aff = -cy.alpha;
        c.mu_aff0 = (cy.S0 + a_aff * (cy.S1 + a_aff * cy.S2)) / (double)NCON;
        const double r = c.mu_aff0 / c.mu;
        const double sm = r * r * r * c.mu;
        c.sigmu = sm > o.tau_min ? sm : o.tau_min;
    }
    NMPC_HD static void after_Fd(LaneCtl& c, const CarryF& cy, const IpmOpts& o)
    {
        c.alpha = -cy.alpha;
        c.mcw = 1.0;
        c.fb = 0;
        if (o.cond_pred_corr) {
            const double mu_c = (cy.S0 + c.alpha * (cy.S1 + c.alpha * cy.S2)) / (double)NCON;
            c.fb = mu_c > 2.0 * c.mu_aff0 ? 1 : 0;
        }
    }
    NMPC_HD static void after_Fd_fallback(LaneCtl& c, const CarryF& cy)
    {
        c.alpha = -cy.alpha;
        c.mcw = 0.0;
        c.nfb++;
    }
    // damped step of HPIPM's UPDATE_VAR_QP; also counts the iteration
    NMPC_HD static double before_B(LaneCtl& c)
    {
        double a = c.alpha;
        if (a < 1.0) a = a * ((1.0 - a) * 0.99 + a * 0.9999999);
        c.iter++;
        return a;
    }

    struct LaneStats {
        int status;       // hpipm-style: 0 ok, 1 max iter, 2 min step, 3 NaN
        int iter;
        double res[4];    // final inf norms res_g, res_b, res_d, res_m
        double mu;
        double lin_res;   // max over iterations of the stationarity residual of the Newton solve
        int cond_fallbacks;
    };

    // one sweep of one lane straight over its tile (lane-resolved base pointer)
    template <int KIND, class F>
    NMPC_HD static void sweep_lane(double* tile_lane, F&& f)
    {
        constexpr bool backward = (KIND == SW_B_FIRST || KIND == SW_B || KIND == SW_BD);
#pragma unroll 1
        for (int s = 0; s <= NSTAGE; s++) {
            const int k = backward ? NSTAGE - s : s;
#if defined(__CUDA_ARCH__) && NMPC_FDF_PREFETCH
            if (KIND == SW_F || KIND == SW_BD || KIND == SW_FD) {
                if (s < NSTAGE) prefetch_stage<KIND>(tile_lane, backward ? k - 1 : k + 1);
            }
#endif
            f(k, tile_stage_in<R>(tile_lane, k), tile_stage_out<R>(tile_lane, k));
        }
    }
#if defined(__CUDA_ARCH__) && NMPC_FDF_PREFETCH
    // The solve sweeps wait on about seven dependent DRAM round trips per stage (78 % of their stall cycles are on the
    // long scoreboard at 14 warps/SM).  The fields the NEXT stage will read are pulled into L2 while this stage computes:
    // the active lanes of the warp split the 128-byte lines of the field ranges between them.
    template <int KIND>
    __device__ static void prefetch_stage(const double* tile_lane, int k)
    {
        const unsigned m = __activemask();
        const int lane = threadIdx.x & (LANES - 1);
        const int n = __popc(m), r = __popc(m & ((1u << lane) - 1u));
        const char* tile0 = reinterpret_cast<const char*>(tile_lane - lane);
        auto pf = [&](size_t off, int nf, int f0, int f1) {
            const char* p = tile0 + (off + ((size_t)k * nf + f0) * LANES) * sizeof(double);
            const int lines = (f1 - f0) * (LANES * (int)sizeof(double) / 128);
            for (int i = r; i < lines; i += n) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (size_t)i * 128));
        };
        if (KIND == SW_F) {                       // LIN[E, DLB, DUB], IT[T, LAM, Z], FA[all]
            pf(R::OFF_LIN, R::NF_LIN, R::E, R::Q); pf(R::OFF_IT, R::NF_IT, R::T, R::PI); pf(R::OFF_FA, R::NF_FA, 0, R::NF_FA);
        } else if (KIND == SW_BD) {               // ST[MC], IT[T], FA[LUU, KH], LIN[E]
            pf(R::OFF_ST, R::NF_ST, R::MC, R::DZA); pf(R::OFF_IT, R::NF_IT, R::T, R::LAM); pf(R::OFF_FA, R::NF_FA, R::LUU, R::LHD);
            pf(R::OFF_LIN, R::NF_LIN, R::E, R::DLB);
        } else {                                  // Fd: FA[LUU, KH, LHD], ST[MC, DZA], IT[T, LAM, Z], LIN[E, DLB, DUB]
            pf(R::OFF_FA, R::NF_FA, R::LUU, R::LH); pf(R::OFF_ST, R::NF_ST, R::MC, R::NF_ST); pf(R::OFF_IT, R::NF_IT, R::T, R::PI);
            pf(R::OFF_LIN, R::NF_LIN, R::E, R::Q);
        }
    }
#endif

    // the phases run by one sweep kernel for one lane; `fallback` selects the mcw = 0 pass of Bd/Fd
    template <int KIND>
    NMPC_HD static void run_phase(double* tile_lane, const Tables& tb, const double* We, const IpmOpts& o, bool fallback, LaneCtl& c,
                                  double* scratch)
    {
        if (KIND == SW_B_FIRST) {
            if (c.done) return;
            CarryB cy; cy.init(scratch);
            sweep_lane<SW_B_FIRST>(tile_lane, [&](int k, const StageIn& in, const StageOut& out) {
                stage_B(k, in, out, tb, We, o, true, 0.0, 0.0, 0.0, cy); });
            after_B(c, cy, o, true);
        } else if (KIND == SW_F) {
            if (c.done) return;
            CarryF cy; cy.init();
            sweep_lane<SW_F>(tile_lane, [&](int k, const StageIn& in, const StageOut& out) {
                stage_F(k, in, out, tb, o, false, 0.0, 0.0, cy); });
            after_F(c, cy, o);
        } else if (KIND == SW_BD) {
            if (fallback ? !c.fb : c.done) return;
            CarryD cy; cy.init();
            const double mcw = fallback ? 0.0 : 1.0;
            sweep_lane<SW_BD>(tile_lane, [&](int k, const StageIn& in, const StageOut& out) {
                stage_Bd(k, in, out, tb, c.sigmu, mcw, cy); });
        } else if (KIND == SW_FD) {
            if (fallback ? !c.fb : c.done) return;
            CarryF cy; cy.init();
            const double mcw = fallback ? 0.0 : 1.0;
            sweep_lane<SW_FD>(tile_lane, [&](int k, const StageIn& in, const StageOut& out) {
                stage_F(k, in, out, tb, o, true, c.sigmu, mcw, cy); });
            if (fallback) after_Fd_fallback(c, cy); else after_Fd(c, cy, o);
        } else if (KIND == SW_FDF) {
            // the three solve sweeps of one iteration back to back (no lane depends on another), with
            // the pure-centering repeat for the lanes whose corrector overshoots
            run_phase<SW_F>(tile_lane, tb, We, o, false, c, scratch);
            run_phase<SW_BD>(tile_lane, tb, We, o, false, c, scratch);
            run_phase<SW_FD>(tile_lane, tb, We, o, false, c, scratch);
            if (!c.done && c.fb) {
                if (fallback) {
                    // hybrid schedule: the rare lane whose corrector overshoots would make its whole warp (and with it
                    // the launch) wait for two more sweeps; it leaves the lockstep path here (done = 2) and the lane-group
                    // kernel redoes this iteration's solve phase, centering repeat included, from the factorisation
                    c.done = 2;
                } else {
                    run_phase<SW_BD>(tile_lane, tb, We, o, true, c, scratch);
                    run_phase<SW_FD>(tile_lane, tb, We, o, true, c, scratch);
                }
            }
        } else {
            if (c.done) return;
            const double a = before_B(c);
            CarryB cy; cy.init(scratch);
#if defined(__CUDA_ARCH__) && NMPC_B_STAGE_IMAGE
            if (!LEAN) {
                // The step rows of a stage (ST[DZ, MC]) travel through a lane-private image in shared memory behind the carry
                // columns: the 8-byte asynchronous copies of stage k-1 are issued between the two halves of stage k, so they are
                // in flight during its Riccati half and need no registers.
                double* img = scratch + (size_t)CarryB::SC_N * PSTRIDE;
                auto fetch = [&](int k) {
                    const double* st = tile_lane + R::OFF_ST + (size_t)k * R::NF_ST * LANES;
#pragma unroll
                    for (int f = 0; f < R::DZA; f++) cp_async8(img + f * PSTRIDE, st + f * LANES);
                    asm volatile("cp.async.commit_group;\n" ::: "memory");
                };
                fetch(NSTAGE);
#pragma unroll 1
                for (int k = NSTAGE; k >= 0; k--) {
                    StageIn in = tile_stage_in<R>(tile_lane, k);
                    const StageOut out = tile_stage_out<R>(tile_lane, k);
                    in.st = img;
                    double gu[NV], gx[NX], rb[NX], Gam[NB2];
                    L lin;
                    stage_B_update_whole<true>(k, in, out, tb, We, o, false, a, c.sigmu, c.mcw, cy, gu, gx, rb, Gam, NMPC_B_KEEP_LIN ? &lin : nullptr);
                    NMPC_PHASE_FENCE();
                    if (k > 0) fetch(k - 1);
#if NMPC_B_L2_PREFETCH
                    if (k > 0) { prefetch_rows_prev(in.lin, R::NF_LIN); prefetch_rows_prev(tile_stage_in<R>(tile_lane, k).it, R::NF_IT); }
#endif
                    stage_B_riccati(k, in, out, tb, We, o, cy, gu, gx, rb, Gam, NMPC_B_KEEP_LIN ? &lin : nullptr);
                }
            } else
#endif
            sweep_lane<SW_B>(tile_lane, [&](int k, const StageIn& in, const StageOut& out) {
                stage_B(k, in, out, tb, We, o, false, a, c.sigmu, c.mcw, cy); });
            after_B(c, cy, o, false);
        }
    }

    // whole IPM of one lane (host emulation of the kernel sequence)
    NMPC_HD static void qp_ipm_lane(double* tile_lane, const Tables& tb, const double* We, const IpmOpts& o, LaneStats& st)
    {
        LaneCtl c; c.init(true);
        double scratch[CarryB::SC_N * PSTRIDE];
        run_phase<SW_B_FIRST>(tile_lane, tb, We, o, false, c, scratch);
        while (!c.done) {
            run_phase<SW_FDF>(tile_lane, tb, We, o, false, c, scratch);
            run_phase<SW_B>(tile_lane, tb, We, o, false, c, scratch);
        }
        st.status = c.status; st.iter = c.iter; st.mu = c.mu; st.lin_res = c.lin_res; st.cond_fallbacks = c.nfb;
        for (int q = 0; q < 4; q++) st.res[q] = c.nrm[q];
    }

    // ------------------------------------------------------------------------------------
    // K4: full step x += dx, u += du for one (instance, stage); x_0 is restored to x0bar.
    // `it` = IT record of the stage.
    // ------------------------------------------------------------------------------------
    template <class RR = R, int LS = LANES>
    NMPC_HD static void step_stage(int k, const double* it, const double* x0bar, double* xk, double* uk)
    {
        if (k < NSTAGE) {
#pragma unroll
            for (int c = 0; c < NU; c++) uk[c] += it[(RR::Z + c) * LS];
        }
#pragma unroll
        for (int j = 0; j < NX; j++) xk[j] = (k == 0) ? x0bar[j] : xk[j] + it[(RR::Z + NU + j) * LS];
    }
};

}  // namespace nmpc
