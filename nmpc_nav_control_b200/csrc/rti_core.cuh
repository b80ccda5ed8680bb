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

// per-stage record layout of one tile (offsets in units of LANES doubles)
template <int NV_>
struct Rec {
    static constexpr int NV = NV_, NX = 3 + 2 * NV, NU = NV, NZ = NX + NU, NC = 1 + 3 * NV, NB2 = 2 * NV;
    // QP data, written by K1/K2
    static constexpr int E = 0;                 // 3*NC   pose rows of [A|B], columns [theta | actual | ref | u]
    static constexpr int B0 = E + 3 * NC;       // NX     b = phi(x,u) - x_next
    static constexpr int Q = B0 + NX;           // NZ     QP gradient, order [u; x]
    static constexpr int DLB = Q + NZ;          // NB2    lb - z   for [u; ref]
    static constexpr int DUB = DLB + NB2;       // NB2    ub - z
    // QP iterate
    static constexpr int Z = DUB + NB2;         // NZ     [u; x]
    static constexpr int PI = Z + NZ;           // NX     multiplier of the dynamics that define x_k
    static constexpr int LAM = PI + NX;         // 2*NB2  lower then upper
    static constexpr int T = LAM + 2 * NB2;     // 2*NB2
    // steps
    static constexpr int DZA = T + 2 * NB2;     // NZ     predictor (affine) step
    static constexpr int DZ = DZA + NZ;         // NZ     final step
    static constexpr int MC = DZ + NZ;          // 2*NB2  dt_aff * dlam_aff
    // factorisation
    static constexpr int LUU = MC + 2 * NB2;    // NV(NV+1)/2 row-packed lower, diagonal stored inverted
    static constexpr int KH = LUU + NV * (NV + 1) / 2;   // NV*NX  K = Luu^-1 S
    static constexpr int LH = KH + NV * NX;     // NV     Luu^-1 q_u (predictor)
    static constexpr int LHD = LH + NV;         // NV     Luu^-1 q_u (delta)
    static constexpr int RB = LHD + NV;         // NX     dynamics residual
    static constexpr int NF = RB + NX;
    static constexpr size_t tile_doubles = (size_t)(NSTAGE + 1) * NF * LANES;
};

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
};

template <int NV>
struct Lin {
    static constexpr int NC = 1 + 3 * NV;
    double E[3][NC];
    double av[NV], ar[NV], au[NV], ru[NV];
};

template <class M>
struct Rti {
    static constexpr int NV = M::NV, NP = M::NP, NX = 3 + 2 * NV, NU = NV, NZ = NX + NU, NY = NZ;
    static constexpr int NC = 1 + 3 * NV, NB2 = 2 * NV, NPK = NX * (NX + 1) / 2, NLU = NV * (NV + 1) / 2;
    static constexpr int NCON = 2 * (NV + (NSTAGE - 1) * NB2 + NV);   // one-sided constraints
    using R = Rec<NV>;
    using L = Lin<NV>;

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
    // K1+K2 for one (instance, stage): writes the QP record of stage k.
    //   xk, uk, xk1: iterate of this stage (uk, xk1 unused for k == N);  yref: nyref entries
    //   (3 = pose only, NY = full); We: terminal weight diagonal (broadcast or per instance).
    // K2 (ocp_nlp_cost_nls with y=[x;u], scripts/<m>/generate_c_code.py:30-39): gradient
    //   s_k W (y - yref), s_k = dt for k<N and 1 for k=N; the Hessian s_k W is never stored.
    // ------------------------------------------------------------------------------------
    NMPC_HD static void linearize_stage(int k, const double* xk, const double* uk, const double* xk1,
                                        const double* yref, int nyref, const double* x0bar,
                                        const Tables& tb, const double* We, double* rec)
    {
        if (k < NSTAGE) {
            double xn[NX], Ep[3][NC];
            rk4_sens(xk, uk, tb.p + k * NP, tb.dt, xn, Ep, nullptr);
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int c = 0; c < NC; c++) rec[(R::E + i * NC + c) * LANES] = Ep[i][c];
#pragma unroll
            for (int i = 0; i < NX; i++) rec[(R::B0 + i) * LANES] = xn[i] - xk1[i];
            const double* Wk = tb.W + k * NY;
#pragma unroll
            for (int c = 0; c < NU; c++) {
                const double yr = (NX + c < nyref) ? yref[NX + c] : 0.0;
                rec[(R::Q + c) * LANES] = (tb.dt * Wk[NX + c]) * (uk[c] - yr);
                rec[(R::DLB + c) * LANES] = tb.lbu[k * NV + c] - uk[c];
                rec[(R::DUB + c) * LANES] = tb.ubu[k * NV + c] - uk[c];
            }
#pragma unroll
            for (int j = 0; j < NX; j++) {
                const double yr = (j < nyref) ? yref[j] : 0.0;
                rec[(R::Q + NU + j) * LANES] = (tb.dt * Wk[j]) * (xk[j] - yr);
            }
        } else {
#pragma unroll
            for (int c = 0; c < NU; c++) rec[(R::Q + c) * LANES] = 0.0;
#pragma unroll
            for (int j = 0; j < NX; j++) {
                const double yr = (j < nyref) ? yref[j] : 0.0;
                rec[(R::Q + NU + j) * LANES] = We[j] * (xk[j] - yr);
            }
        }
        if (k >= 1) {
#pragma unroll
            for (int c = 0; c < NV; c++) {
                rec[(R::DLB + NV + c) * LANES] = tb.lbx[(k - 1) * NV + c] - xk[3 + NV + c];
                rec[(R::DUB + NV + c) * LANES] = tb.ubx[(k - 1) * NV + c] - xk[3 + NV + c];
            }
        } else {
            // x0 elimination (d_ocp_qp_reduce_eq_dof): the stage-0 state is the constant x0bar - x_0
#pragma unroll
            for (int j = 0; j < NX; j++) rec[(R::Z + NU + j) * LANES] = x0bar[j] - xk[j];
        }
    }

    // ------------------------------------------------------------------------------------
    // structured products with [B A] of one stage
    // ------------------------------------------------------------------------------------
    NMPC_HD static void load_lin(const double* rec, const double* lti, L& l)
    {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int c = 0; c < NC; c++) l.E[i][c] = rec[(R::E + i * NC + c) * LANES];
#pragma unroll
        for (int c = 0; c < NV; c++) { l.av[c] = lti[c]; l.ar[c] = lti[NV + c]; l.au[c] = lti[2 * NV + c]; l.ru[c] = lti[3 * NV + c]; }
    }
    // [B A]' v  -> ou (NV), ox (NX)
    NMPC_HD static void apply_T(const L& l, const double* v, double* ou, double* ox)
    {
#pragma unroll
        for (int c = 0; c < NV; c++)
            ou[c] = l.E[0][1 + 2 * NV + c] * v[0] + l.E[1][1 + 2 * NV + c] * v[1] + l.E[2][1 + 2 * NV + c] * v[2]
                  + l.au[c] * v[3 + c] + l.ru[c] * v[3 + NV + c];
        ox[0] = v[0]; ox[1] = v[1];
        ox[2] = l.E[0][0] * v[0] + l.E[1][0] * v[1] + l.E[2][0] * v[2];
#pragma unroll
        for (int c = 0; c < NV; c++) {
            ox[3 + c] = l.E[0][1 + c] * v[0] + l.E[1][1 + c] * v[1] + l.E[2][1 + c] * v[2] + l.av[c] * v[3 + c];
            ox[3 + NV + c] = l.E[0][1 + NV + c] * v[0] + l.E[1][1 + NV + c] * v[1] + l.E[2][1 + NV + c] * v[2]
                           + l.ar[c] * v[3 + c] + v[3 + NV + c];
        }
    }
    // A x + B u -> xn
    NMPC_HD static void apply(const L& l, const double* u, const double* x, double* xn)
    {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            double a = (i < 2 ? x[i] : 0.0) + l.E[i][0] * x[2];
#pragma unroll
            for (int c = 0; c < NV; c++)
                a += l.E[i][1 + c] * x[3 + c] + l.E[i][1 + NV + c] * x[3 + NV + c] + l.E[i][1 + 2 * NV + c] * u[c];
            xn[i] = a;
        }
#pragma unroll
        for (int c = 0; c < NV; c++) {
            xn[3 + c] = l.av[c] * x[3 + c] + l.ar[c] * x[3 + NV + c] + l.au[c] * u[c];
            xn[3 + NV + c] = x[3 + NV + c] + l.ru[c] * u[c];
        }
    }
    // g = P * (column j of [B A]), j in z order [u; x]
    NMPC_HD static void P_col(const double* P, const L& l, int j, double* g)
    {
        if (j < NV) {
            const int c = j;
#pragma unroll
            for (int i = 0; i < NX; i++)
                g[i] = P[pk(i, 0)] * l.E[0][1 + 2 * NV + c] + P[pk(i, 1)] * l.E[1][1 + 2 * NV + c] + P[pk(i, 2)] * l.E[2][1 + 2 * NV + c]
                     + P[pk(i, 3 + c)] * l.au[c] + P[pk(i, 3 + NV + c)] * l.ru[c];
        } else if (j < NV + 2) {
#pragma unroll
            for (int i = 0; i < NX; i++) g[i] = P[pk(i, j - NV)];
        } else if (j == NV + 2) {
#pragma unroll
            for (int i = 0; i < NX; i++) g[i] = P[pk(i, 0)] * l.E[0][0] + P[pk(i, 1)] * l.E[1][0] + P[pk(i, 2)] * l.E[2][0];
        } else if (j < NV + 3 + NV) {
            const int c = j - NV - 3;
#pragma unroll
            for (int i = 0; i < NX; i++)
                g[i] = P[pk(i, 0)] * l.E[0][1 + c] + P[pk(i, 1)] * l.E[1][1 + c] + P[pk(i, 2)] * l.E[2][1 + c] + P[pk(i, 3 + c)] * l.av[c];
        } else {
            const int c = j - NV - 3 - NV;
#pragma unroll
            for (int i = 0; i < NX; i++)
                g[i] = P[pk(i, 0)] * l.E[0][1 + NV + c] + P[pk(i, 1)] * l.E[1][1 + NV + c] + P[pk(i, 2)] * l.E[2][1 + NV + c]
                     + P[pk(i, 3 + c)] * l.ar[c] + P[pk(i, 3 + NV + c)];
        }
    }

    // u = -Luu^-T (lh + K dx), Luu row-packed with inverted diagonal
    NMPC_HD static void solve_u(const double* rec, int lh_field, bool with_K, const double* dx, double* du)
    {
        double v[NV];
#pragma unroll
        for (int a = 0; a < NV; a++) {
            double s = rec[(lh_field + a) * LANES];
            if (with_K) {
#pragma unroll
                for (int j = 0; j < NX; j++) s += rec[(R::KH + a * NX + j) * LANES] * dx[j];
            }
            v[a] = -s;
        }
#pragma unroll
        for (int a = NV - 1; a >= 0; a--) {
            double s = v[a];
#pragma unroll
            for (int b = a + 1; b < NV; b++) s -= rec[(R::LUU + b * (b + 1) / 2 + a) * LANES] * du[b];
            du[a] = s * rec[(R::LUU + a * (a + 1) / 2 + a) * LANES];
        }
    }

    struct LaneStats {
        int status;       // hpipm-style: 0 ok, 1 max iter, 2 min step, 3 NaN
        int iter;
        double res[4];    // final inf norms res_g, res_b, res_d, res_m
        double mu;
        double lin_res;   // max over iterations of the stationarity residual of the Newton solve
        int cond_fallbacks;
    };

    // ------------------------------------------------------------------------------------
    // B sweep: (apply previous step) + residuals + Riccati factorisation, stage N..0
    // ------------------------------------------------------------------------------------
    NMPC_HD static void sweep_B(double* base, const Tables& tb, const double* We, const IpmOpts& o, bool first,
                                double a_step, double sigmu, double mcw, double* nrm, double* mu_out, double* lru_out)
    {
        double P[NPK], pv[NX], pi_o[NX], dpi[NX], xn[NX];
        double ng = 0.0, nb = 0.0, nd = 0.0, nm = 0.0, musum = 0.0, lru = 0.0;
#pragma unroll
        for (int i = 0; i < NX; i++) { pv[i] = 0.0; pi_o[i] = 0.0; dpi[i] = 0.0; xn[i] = 0.0; }
#pragma unroll
        for (int i = 0; i < NPK; i++) P[i] = 0.0;

        for (int k = NSTAGE; k >= 0; k--) {
            double* rec = base + (size_t)k * R::NF * LANES;
            const bool hasU = k < NSTAGE, hasX = k > 0;
            L lin;
            if (hasU) load_lin(rec, tb.lti + k * 4 * NV, lin);
            double Hu[NV], Hx[NX], qu[NV], qx[NX];
#pragma unroll
            for (int c = 0; c < NV; c++) { Hu[c] = hasU ? tb.dt * tb.W[k * NY + NX + c] : 0.0; qu[c] = rec[(R::Q + c) * LANES]; }
#pragma unroll
            for (int j = 0; j < NX; j++) { Hx[j] = hasU ? tb.dt * tb.W[k * NY + j] : We[j]; qx[j] = rec[(R::Q + NU + j) * LANES]; }

            double zu[NV], zx[NX], pin[NX];
            double ll[NB2], lu[NB2], tl[NB2], tu[NB2], dl[NB2], du_[NB2];
#pragma unroll
            for (int b = 0; b < NB2; b++) { dl[b] = rec[(R::DLB + b) * LANES]; du_[b] = rec[(R::DUB + b) * LANES]; }

            double v1u[NV], v1x[NX], v2u[NV], v2x[NX];
            if (hasU) { apply_T(lin, pi_o, v1u, v1x); apply_T(lin, dpi, v2u, v2x); }
            else {
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
                for (int j = 0; j < NX; j++) { zx[j] = hasX ? 0.0 : rec[(R::Z + NU + j) * LANES]; pin[j] = 0.0; pi_old[j] = 0.0; }
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
            } else {
                double dzu[NV], dzx[NX];
#pragma unroll
                for (int c = 0; c < NV; c++) { zu[c] = rec[(R::Z + c) * LANES]; dzu[c] = hasU ? rec[(R::DZ + c) * LANES] : 0.0; }
#pragma unroll
                for (int j = 0; j < NX; j++) {
                    zx[j] = rec[(R::Z + NU + j) * LANES];
                    dzx[j] = hasX ? rec[(R::DZ + NU + j) * LANES] : 0.0;
                    pin[j] = hasX ? rec[(R::PI + j) * LANES] : 0.0;
                    pi_old[j] = pin[j];
                }
                double ldo[NB2], dld[NB2];   // (lam_u - lam_l) old, (dlam_l - dlam_u)
#pragma unroll
                for (int b = 0; b < NB2; b++) {
                    const bool act = (b < NV) ? hasU : hasX;
                    if (act) {
                        ll[b] = rec[(R::LAM + b) * LANES]; lu[b] = rec[(R::LAM + NB2 + b) * LANES];
                        tl[b] = rec[(R::T + b) * LANES];   tu[b] = rec[(R::T + NB2 + b) * LANES];
                        const double mc_l = rec[(R::MC + b) * LANES], mc_u = rec[(R::MC + NB2 + b) * LANES];
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
                // stationarity residual of the Newton system, control rows (diagnostic: the
                // quantity HPIPM's iterative refinement would test)
                if (hasU) {
#pragma unroll
                    for (int c = 0; c < NV; c++) {
                        const double r = qu[c] + Hu[c] * zu[c] + ldo[c] + v1u[c] + Hu[c] * dzu[c] - dld[c] + v2u[c];
                        lru = fmax(lru, fabs(r));
                    }
                }
                // adjoint recursion for the multiplier step of the dynamics that define x_k
                if (hasX) {
#pragma unroll
                    for (int j = 0; j < NX; j++) {
                        double r = qx[j] + Hx[j] * zx[j] - pin[j] + v1x[j] + Hx[j] * dzx[j] + v2x[j];
                        if (j >= 3 + NV) r += ldo[j - 3] - dld[j - 3];
                        dpi[j] = r;
                        pin[j] += a_step * r;
                        zx[j] += a_step * dzx[j];
                    }
                }
#pragma unroll
                for (int c = 0; c < NV; c++) zu[c] += a_step * dzu[c];
            }

            // ---- residuals at the (new) iterate -------------------------------------------
            double rgu[NV], rgx[NX], rb[NX];
#pragma unroll
            for (int c = 0; c < NV; c++) {
                rgu[c] = qu[c] + Hu[c] * zu[c] + (lu[c] - ll[c]) + (v1u[c] + a_step * v2u[c]);
                if (hasU) ng = fmax(ng, fabs(rgu[c]));
            }
#pragma unroll
            for (int j = 0; j < NX; j++) {
                double r = qx[j] + Hx[j] * zx[j] - pin[j] + (v1x[j] + a_step * v2x[j]);
                if (j >= 3 + NV) r += lu[j - 3] - ll[j - 3];
                rgx[j] = r;
                if (hasX) ng = fmax(ng, fabs(r));
            }
            if (hasU) {
                apply(lin, zu, zx, rb);
#pragma unroll
                for (int i = 0; i < NX; i++) { rb[i] += rec[(R::B0 + i) * LANES] - xn[i]; nb = fmax(nb, fabs(rb[i])); }
            }
            double Gam[NB2], gam[NB2];
#pragma unroll
            for (int b = 0; b < NB2; b++) {
                const bool act = (b < NV) ? hasU : hasX;
                if (act) {
                    const double zb = (b < NV) ? zu[b] : zx[3 + b];
                    const double rd_l = dl[b] - zb + tl[b], rd_u = -du_[b] + zb + tu[b];
                    const double pm_l = ll[b] * tl[b], pm_u = lu[b] * tu[b];
                    musum += pm_l + pm_u;
                    const double rm_l = pm_l - o.tau_min, rm_u = pm_u - o.tau_min;
                    nd = fmax(nd, fmax(fabs(rd_l), fabs(rd_u)));
                    nm = fmax(nm, fmax(fabs(rm_l), fabs(rm_u)));
                    const double ti_l = tl[b] < o.t_min ? 1.0 / o.t_min : 1.0 / tl[b];
                    const double ti_u = tu[b] < o.t_min ? 1.0 / o.t_min : 1.0 / tu[b];
                    const double l_l = ll[b] < o.lam_min ? o.lam_min : ll[b];
                    const double l_u = lu[b] < o.lam_min ? o.lam_min : lu[b];
                    Gam[b] = ti_l * l_l + ti_u * l_u;
                    gam[b] = ti_l * (rm_l - ll[b] * rd_l) - ti_u * (rm_u - lu[b] * rd_u);
                } else { Gam[b] = 0.0; gam[b] = 0.0; }
            }

            // ---- store the iterate ----------------------------------------------------------
#pragma unroll
            for (int c = 0; c < NV; c++) rec[(R::Z + c) * LANES] = zu[c];
            if (hasX) {
#pragma unroll
                for (int j = 0; j < NX; j++) { rec[(R::Z + NU + j) * LANES] = zx[j]; rec[(R::PI + j) * LANES] = pin[j]; }
            }
#pragma unroll
            for (int b = 0; b < NB2; b++) {
                rec[(R::LAM + b) * LANES] = ll[b]; rec[(R::LAM + NB2 + b) * LANES] = lu[b];
                rec[(R::T + b) * LANES] = tl[b];   rec[(R::T + NB2 + b) * LANES] = tu[b];
            }

            // ---- Riccati step ---------------------------------------------------------------
            // gradient of the stage incl. barrier terms and the cost-to-go of the successor
            double gu[NV], gx[NX];
#pragma unroll
            for (int c = 0; c < NV; c++) gu[c] = rgu[c] + gam[c];
#pragma unroll
            for (int j = 0; j < NX; j++) gx[j] = rgx[j] + (j >= 3 + NV ? gam[j - 3] : 0.0);
            if (hasU) {
#pragma unroll
                for (int i = 0; i < NX; i++) rec[(R::RB + i) * LANES] = rb[i];
                double Pb[NX];
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    double s = pv[i];
#pragma unroll
                    for (int j = 0; j < NX; j++) s += P[pk(i, j)] * rb[j];
                    Pb[i] = s;
                }
                double tu_[NV], tx_[NX];
                apply_T(lin, Pb, tu_, tx_);
#pragma unroll
                for (int c = 0; c < NV; c++) gu[c] += tu_[c];
#pragma unroll
                for (int j = 0; j < NX; j++) gx[j] += tx_[j];

                // M = [B A]' P [B A] + diag(H + Gamma + reg), packed lower in z order
                constexpr int NMK = NZ * (NZ + 1) / 2;
                double Mk[NMK];
#pragma unroll
                for (int j = 0; j < NZ; j++) {
                    if (!hasX && j >= NV) break;
                    double g[NX], cu[NV], cx[NX];
                    P_col(P, lin, j, g);
                    apply_T(lin, g, cu, cx);
#pragma unroll
                    for (int i = j; i < NZ; i++) Mk[i * (i + 1) / 2 + j] = (i < NV) ? cu[i] : cx[i - NV];
                }
#pragma unroll
                for (int c = 0; c < NV; c++) Mk[c * (c + 1) / 2 + c] += Hu[c] + o.reg_prim + Gam[c];
                // Cholesky of the control block, diagonal kept inverted
                double Luu[NLU];
#pragma unroll
                for (int a = 0; a < NV; a++) {
                    double d = Mk[a * (a + 1) / 2 + a];
#pragma unroll
                    for (int c = 0; c < a; c++) d -= Luu[a * (a + 1) / 2 + c] * Luu[a * (a + 1) / 2 + c];
                    const double inv = d > 0.0 ? 1.0 / sqrt(d) : 0.0;
                    Luu[a * (a + 1) / 2 + a] = inv;
#pragma unroll
                    for (int b = a + 1; b < NV; b++) {
                        double s = Mk[b * (b + 1) / 2 + a];
#pragma unroll
                        for (int c = 0; c < a; c++) s -= Luu[b * (b + 1) / 2 + c] * Luu[a * (a + 1) / 2 + c];
                        Luu[b * (b + 1) / 2 + a] = s * inv;
                    }
                }
#pragma unroll
                for (int i = 0; i < NLU; i++) rec[(R::LUU + i) * LANES] = Luu[i];
                double lh[NV];
#pragma unroll
                for (int a = 0; a < NV; a++) {
                    double s = gu[a];
#pragma unroll
                    for (int c = 0; c < a; c++) s -= Luu[a * (a + 1) / 2 + c] * lh[c];
                    lh[a] = s * Luu[a * (a + 1) / 2 + a];
                    rec[(R::LH + a) * LANES] = lh[a];
                }
                if (hasX) {
                    double Kh[NV][NX];
#pragma unroll
                    for (int j = 0; j < NX; j++) {
#pragma unroll
                        for (int a = 0; a < NV; a++) {
                            double s = Mk[(NV + j) * (NV + j + 1) / 2 + a];
#pragma unroll
                            for (int c = 0; c < a; c++) s -= Luu[a * (a + 1) / 2 + c] * Kh[c][j];
                            Kh[a][j] = s * Luu[a * (a + 1) / 2 + a];
                            rec[(R::KH + a * NX + j) * LANES] = Kh[a][j];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < NX; i++) {
#pragma unroll
                        for (int j = 0; j <= i; j++) {
                            double s = Mk[(NV + i) * (NV + i + 1) / 2 + NV + j];
#pragma unroll
                            for (int a = 0; a < NV; a++) s -= Kh[a][i] * Kh[a][j];
                            if (i == j) s += Hx[i] + o.reg_prim + (i >= 3 + NV ? Gam[i - 3] : 0.0);
                            P[pk(i, j)] = s;
                        }
                        double s = gx[i];
#pragma unroll
                        for (int a = 0; a < NV; a++) s -= Kh[a][i] * lh[a];
                        pv[i] = s;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < NX; i++) {
#pragma unroll
                    for (int j = 0; j <= i; j++) P[pk(i, j)] = (i == j) ? Hx[i] + o.reg_prim + (i >= 3 + NV ? Gam[i - 3] : 0.0) : 0.0;
                    pv[i] = gx[i];
                }
            }
            // carries for stage k-1
#pragma unroll
            for (int j = 0; j < NX; j++) { pi_o[j] = pi_old[j]; xn[j] = zx[j]; }
            if (first) {
#pragma unroll
                for (int j = 0; j < NX; j++) dpi[j] = 0.0;
            }
        }
        nrm[0] = ng; nrm[1] = nb; nrm[2] = nd; nrm[3] = nm;
        *mu_out = musum / (double)NCON;
        *lru_out = lru;
    }

    // ------------------------------------------------------------------------------------
    // forward sweeps. delta == false: predictor (writes DZA, MC); delta == true: adds the delta
    // step to the predictor (writes DZ).  Returns the ratio-test step and the three sums that
    // give mu(alpha) = (S0 + alpha S1 + alpha^2 S2) / nc.
    // ------------------------------------------------------------------------------------
    NMPC_HD static void sweep_F(double* base, const Tables& tb, const IpmOpts& o, bool delta, double sigmu, double mcw,
                                double* alpha_out, double* S)
    {
        double dx[NX];
#pragma unroll
        for (int j = 0; j < NX; j++) dx[j] = 0.0;
        double alpha = -1.0, S0 = 0.0, S1 = 0.0, S2 = 0.0;
        for (int k = 0; k <= NSTAGE; k++) {
            double* rec = base + (size_t)k * R::NF * LANES;
            const bool hasU = k < NSTAGE, hasX = k > 0;
            double du[NV];
#pragma unroll
            for (int c = 0; c < NV; c++) du[c] = 0.0;
            if (hasU) solve_u(rec, delta ? R::LHD : R::LH, hasX, dx, du);
            double dzu[NV], dzx[NX];
            if (!delta) {
#pragma unroll
                for (int c = 0; c < NV; c++) { dzu[c] = du[c]; rec[(R::DZA + c) * LANES] = du[c]; }
#pragma unroll
                for (int j = 0; j < NX; j++) { dzx[j] = dx[j]; rec[(R::DZA + NU + j) * LANES] = dx[j]; }
            } else {
#pragma unroll
                for (int c = 0; c < NV; c++) { dzu[c] = rec[(R::DZA + c) * LANES] + du[c]; rec[(R::DZ + c) * LANES] = dzu[c]; }
#pragma unroll
                for (int j = 0; j < NX; j++) { dzx[j] = rec[(R::DZA + NU + j) * LANES] + dx[j]; rec[(R::DZ + NU + j) * LANES] = dzx[j]; }
            }
#pragma unroll
            for (int b = 0; b < NB2; b++) {
                const bool act = (b < NV) ? hasU : hasX;
                if (act) {
                    const double ll = rec[(R::LAM + b) * LANES], lu = rec[(R::LAM + NB2 + b) * LANES];
                    const double tl = rec[(R::T + b) * LANES], tu = rec[(R::T + NB2 + b) * LANES];
                    const double zb = (b < NV) ? rec[(R::Z + b) * LANES] : rec[(R::Z + NU + 3 + b) * LANES];
                    const double dzb = (b < NV) ? dzu[b] : dzx[3 + b];
                    const double rd_l = rec[(R::DLB + b) * LANES] - zb + tl, rd_u = -rec[(R::DUB + b) * LANES] + zb + tu;
                    double rm_l = ll * tl - o.tau_min, rm_u = lu * tu - o.tau_min;
                    if (delta) {
                        rm_l += mcw * rec[(R::MC + b) * LANES] - sigmu;
                        rm_u += mcw * rec[(R::MC + NB2 + b) * LANES] - sigmu;
                    }
                    const double dt_l = dzb - rd_l, dt_u = -dzb - rd_u;
                    const double dl_l = -(ll * dt_l + rm_l) / tl, dl_u = -(lu * dt_u + rm_u) / tu;
                    if (!delta) { rec[(R::MC + b) * LANES] = dt_l * dl_l; rec[(R::MC + NB2 + b) * LANES] = dt_u * dl_u; }
                    if (alpha * dl_l > ll) alpha = ll / dl_l;
                    if (alpha * dt_l > tl) alpha = tl / dt_l;
                    if (alpha * dl_u > lu) alpha = lu / dl_u;
                    if (alpha * dt_u > tu) alpha = tu / dt_u;
                    S0 += ll * tl + lu * tu;
                    S1 += ll * dt_l + tl * dl_l + lu * dt_u + tu * dl_u;
                    S2 += dl_l * dt_l + dl_u * dt_u;
                }
            }
            if (hasU) {
                L lin;
                load_lin(rec, tb.lti + k * 4 * NV, lin);
                double xnew[NX];
                apply(lin, du, dx, xnew);
#pragma unroll
                for (int j = 0; j < NX; j++) dx[j] = xnew[j] + (delta ? 0.0 : rec[(R::RB + j) * LANES]);
            }
        }
        *alpha_out = -alpha;
        S[0] = S0; S[1] = S1; S[2] = S2;
    }

    // delta backward sweep: rhs only in the complementarity rows
    NMPC_HD static void sweep_Bd(double* base, const Tables& tb, double sigmu, double mcw)
    {
        double dp[NX];
#pragma unroll
        for (int j = 0; j < NX; j++) dp[j] = 0.0;
        for (int k = NSTAGE; k >= 0; k--) {
            double* rec = base + (size_t)k * R::NF * LANES;
            const bool hasU = k < NSTAGE, hasX = k > 0;
            double qu[NV], qx[NX];
#pragma unroll
            for (int c = 0; c < NV; c++) qu[c] = 0.0;
#pragma unroll
            for (int j = 0; j < NX; j++) qx[j] = 0.0;
            if (hasU) {
                L lin;
                load_lin(rec, tb.lti + k * 4 * NV, lin);
                apply_T(lin, dp, qu, qx);
            }
#pragma unroll
            for (int b = 0; b < NB2; b++) {
                const bool act = (b < NV) ? hasU : hasX;
                if (act) {
                    const double tl = rec[(R::T + b) * LANES], tu = rec[(R::T + NB2 + b) * LANES];
                    const double g = (mcw * rec[(R::MC + b) * LANES] - sigmu) / tl - (mcw * rec[(R::MC + NB2 + b) * LANES] - sigmu) / tu;
                    if (b < NV) qu[b] += g; else qx[3 + b] += g;
                }
            }
            if (hasU) {
                double lh[NV];
#pragma unroll
                for (int a = 0; a < NV; a++) {
                    double s = qu[a];
#pragma unroll
                    for (int c = 0; c < a; c++) s -= rec[(R::LUU + a * (a + 1) / 2 + c) * LANES] * lh[c];
                    lh[a] = s * rec[(R::LUU + a * (a + 1) / 2 + a) * LANES];
                    rec[(R::LHD + a) * LANES] = lh[a];
                }
                if (hasX) {
#pragma unroll
                    for (int j = 0; j < NX; j++) {
                        double s = qx[j];
#pragma unroll
                        for (int a = 0; a < NV; a++) s -= rec[(R::KH + a * NX + j) * LANES] * lh[a];
                        dp[j] = s;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < NX; j++) dp[j] = qx[j];
            }
        }
    }

    // ------------------------------------------------------------------------------------
    // K3: the interior-point loop of one lane.  `active` false = padding lane of the last tile.
    // ------------------------------------------------------------------------------------
    NMPC_HD static void qp_ipm_lane(double* base, const Tables& tb, const double* We, const IpmOpts& o, bool active, LaneStats& st)
    {
        bool done = !active;
        double nrm[4] = {0, 0, 0, 0}, mu = 0.0, alpha = 1.0, lru = 0.0;
        int iter = 0;
        st.lin_res = 0.0; st.cond_fallbacks = 0; st.status = 0;
        if (!done) sweep_B(base, tb, We, o, true, 0.0, 0.0, 0.0, nrm, &mu, &lru);
        while (true) {
            if (!done) {
                const bool more = iter < o.iter_max && alpha > o.alpha_min &&
                                  (nrm[0] > o.res_g_max || nrm[1] > o.res_b_max || nrm[2] > o.res_d_max ||
                                   fabs(nrm[3] - o.tau_min) > o.res_m_max);
                if (!more || mu != mu) {
                    done = true;
                    st.status = (mu != mu) ? 3 : (iter >= o.iter_max ? 1 : (alpha <= o.alpha_min ? 2 : 0));
                }
            }
            if (!NMPC_ANY(!done)) break;
            double S[3], a_aff = 1.0, sigmu = 0.0, mu_aff0 = 0.0;
            if (!done) {
                sweep_F(base, tb, o, false, 0.0, 0.0, &a_aff, S);
                mu_aff0 = (S[0] + a_aff * (S[1] + a_aff * S[2])) / (double)NCON;
                const double r = mu_aff0 / mu;
                sigmu = r * r * r * mu;
                sigmu = sigmu > o.tau_min ? sigmu : o.tau_min;
                sweep_Bd(base, tb, sigmu, 1.0);
                sweep_F(base, tb, o, true, sigmu, 1.0, &alpha, S);
            }
            double mcw = 1.0;
            bool fb = false;
            if (!done && o.cond_pred_corr) {
                const double mu_c = (S[0] + alpha * (S[1] + alpha * S[2])) / (double)NCON;
                fb = mu_c > 2.0 * mu_aff0;
            }
            if (NMPC_ANY(fb)) {
                if (fb) {   // pure centering direction
                    mcw = 0.0;
                    st.cond_fallbacks++;
                    sweep_Bd(base, tb, sigmu, 0.0);
                    sweep_F(base, tb, o, true, sigmu, 0.0, &alpha, S);
                }
            }
            double a = alpha;
            if (a < 1.0) a = a * ((1.0 - a) * 0.99 + a * 0.9999999);
            if (!done) {
                iter++;
                sweep_B(base, tb, We, o, false, a, sigmu, mcw, nrm, &mu, &lru);
                st.lin_res = fmax(st.lin_res, lru);
            }
        }
        st.iter = iter;
        st.mu = mu;
#pragma unroll
        for (int q = 0; q < 4; q++) st.res[q] = nrm[q];
    }

    // ------------------------------------------------------------------------------------
    // K4: full step x += dx, u += du for one (instance, stage); x_0 is restored to x0bar.
    // ------------------------------------------------------------------------------------
    NMPC_HD static void step_stage(int k, const double* rec, const double* x0bar, double* xk, double* uk)
    {
        if (k < NSTAGE) {
#pragma unroll
            for (int c = 0; c < NU; c++) uk[c] += rec[(R::Z + c) * LANES];
        }
#pragma unroll
        for (int j = 0; j < NX; j++) xk[j] = (k == 0) ? x0bar[j] : xk[j] + rec[(R::Z + NU + j) * LANES];
    }
};

}  // namespace nmpc
