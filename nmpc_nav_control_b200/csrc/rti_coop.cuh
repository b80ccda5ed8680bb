// Warp-cooperative K3 ("latency path"): ONE warp solves ONE interior-point QP.
//
// The throughput path (rti_core.cuh, k_sweep) gives every instance one thread; one IPM iteration
// of one instance is then ~1.4 ms of sequential work, which is what a batch of 1 (the ROS drop-in,
// NMPCNavControlDiff.cpp:142) and the last stragglers of a large batch would pay.  Here the 32
// lanes of a warp share one instance:
//   * everything that is local to a stage (slack / multiplier updates, residuals, barrier terms,
//     ratio tests) runs as PARALLEL passes over all (stage, component) items of the horizon;
//   * the three recursions that are inherently sequential over the horizon (Riccati with the
//     adjoint recursion folded in, forward substitution, delta backward substitution) run stage by
//     stage with the small dense products (P*[B A], [B A]'*G, Schur complement) spread over lanes
//     through a per-warp shared-memory scratch.
// The iteration path (initial point, predictor / corrector, conditional centering, step rule, exit
// test) and the state layout (the tile records of Rec<NV>) are those of the throughput path, so an
// instance can be handed from one path to the other at an iteration boundary, and K1/K2/K4 are
// shared.  The arithmetic is the same Newton systems in a different summation order; results agree
// with the oracle to rounding, not bit for bit.
//
// Written with work-item loops (COOP_FOR) so that the host emulation (tests/host_emul) executes
// the identical code with one "lane".
#pragma once
#include "rti_core.cuh"

#if defined(__CUDA_ARCH__)
#define COOP_FOR(idx, n) for (int idx = (int)(threadIdx.x & 31u); idx < (n); idx += 32)
#define COOP_SYNC() __syncwarp()
#else
#define COOP_FOR(idx, n) for (int idx = 0; idx < (n); idx++)
#define COOP_SYNC() ((void)0)
#endif

namespace nmpc {

NMPC_HD double coop_sum(double v)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
    return v;
}
NMPC_HD double coop_max(double v)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
#endif
    return v;
}
// the ratio test keeps the NEGATED step length (HPIPM): start -1, the binding one is the largest
NMPC_HD double coop_max_neg(double v) { return coop_max(v); }

template <class M>
struct Coop {
    using S = Rti<M>;
    using R = typename S::R;
    using LaneCtl = typename S::LaneCtl;
    static constexpr int NV = S::NV, NX = S::NX, NU = S::NU, NZ = S::NZ, NY = S::NY, NC = S::NC, NB2 = S::NB2, NLU = S::NLU;
    static constexpr int NCON = S::NCON;
    static constexpr int NK = NSTAGE + 1;

    // per-warp scratch (shared memory on the device)
    struct Scratch {
        double BA[NX * NZ];      // [i][c] dynamics of the current stage, columns in z order [u; x]
        double P[NX * NX];       // cost-to-go of the successor (full symmetric)
        double G[NX * NZ];       // P * BA
        double Mm[NZ * NZ];      // BA' * G + diag (lower triangle used)
        double K[NV * NX];       // Luu^-1 Mux
        double Luu[NLU];         // row-packed lower Cholesky factor of Muu, diagonal inverted
        double pv[NX], Pb[NX];   // gradient of the cost-to-go; P*rb + pv
        double pn[NX], pw[NX];   // full-step / damped-step multipliers of the successor's dynamics
        double t1[NZ], t2[NZ];   // BA' * pn, BA' * pw
        double g[NZ];            // stage gradient with all terms
        double lh[NV];
        double dz[NZ];           // [du; dx] of the current stage in the substitutions
        double xn[NX];           // next dx / previous dp
        double in[3 * NC + 4 * NZ + 2 * NB2 + NV * NX + NLU + 2 * NV + 2 * NX];   // staged record of the stage
    };

    // ---- addressing of the instance's records (lane-resolved tile pointer, field stride LANES) ----
    NMPC_HD static double& LINf(double* t, int k, int f) { return t[R::OFF_LIN + ((size_t)k * R::NF_LIN + f) * LANES]; }
    NMPC_HD static double& ITf(double* t, int k, int f) { return t[R::OFF_IT + ((size_t)k * R::NF_IT + f) * LANES]; }
    NMPC_HD static double& STf(double* t, int k, int f) { return t[R::OFF_ST + ((size_t)k * R::NF_ST + f) * LANES]; }
    NMPC_HD static double& FAf(double* t, int k, int f) { return t[R::OFF_FA + ((size_t)k * R::NF_FA + f) * LANES]; }

    // component c of z = [u; x] at stage k: bound slot index b (0..NB2-1) or -1, and whether the stage has it
    NMPC_HD static int bound_of(int c) { return c < NV ? c : (c >= NU + 3 + NV ? c - NU - 3 : -1); }
    NMPC_HD static int comp_of_bound(int b) { return b < NV ? b : NU + 3 + b; }
    NMPC_HD static bool has_comp(int k, int c) { return c < NU ? (k < NSTAGE) : (k > 0); }
    NMPC_HD static double hess(const Tables& tb, const double* We, int k, int c)
    {
        if (k < NSTAGE) return tb.dt * (c < NU ? tb.W[k * NY + NX + c] : tb.W[k * NY + (c - NU)]);
        return c < NU ? 0.0 : We[c - NU];
    }
    // entry (i, c) of [B A] of stage k (c in z order) from the pose rows E and the LTI table
    NMPC_HD static double ba_entry(const double* E /* 3*NC, stride es */, int es, const double* lti, int i, int c)
    {
        // E columns: [theta | actual(NV) | ref(NV) | u(NV)]
        if (i < 3) {
            if (c < NU) return E[(i * NC + 1 + 2 * NV + c) * es];
            const int j = c - NU;
            if (j < 2) return (i == j) ? 1.0 : 0.0;
            if (j == 2) return E[(i * NC) * es];
            return E[(i * NC + 1 + (j - 3)) * es];          // actual then ref, contiguous in E
        }
        const int r = i - 3;
        if (r < NV) {            // actual_r' = av actual_r + ar ref_r + au u_r
            if (c < NU) return c == r ? lti[2 * NV + r] : 0.0;
            const int j = c - NU;
            if (j == 3 + r) return lti[r];
            if (j == 3 + NV + r) return lti[NV + r];
            return 0.0;
        }
        const int q = r - NV;    // ref_q' = ref_q + ru u_q
        if (c < NU) return c == q ? lti[3 * NV + q] : 0.0;
        return (c - NU == 3 + NV + q) ? 1.0 : 0.0;
    }

    struct Sums { double alpha, S0, S1, S2; };

    // ====================================================================================
    // B sweep: (apply previous step) + residuals + factorisation.
    // ====================================================================================
    NMPC_HD static void sweep_B(double* t, const Tables& tb, const double* We, const IpmOpts& o, bool first, LaneCtl& c, Scratch& s)
    {
        const double a = first ? 0.0 : S::before_B(c);
        const double sigmu = c.sigmu, mcw = c.mcw;
        // ---- parallel pass over all (stage, component): step, slacks / multipliers, barrier terms ----
        // leaves: IT.Z/T/LAM updated; ST.DZ[c] = q + H z_new + (lam_u - lam_l)_new ("rgq");
        //         ST.DZA[c] = q + H (z_old + dz) + ((lam_u-lam_l) + (dlam_u-dlam_l)) ("cz", full-step);
        //         ST.MC[b] = Gamma, ST.MC[NB2+b] = gamma
        double musum = 0.0, nd = 0.0, nm = 0.0;
        COOP_FOR(it, NK * NZ)
        {
            const int k = it / NZ, cc = it % NZ;
            if (!has_comp(k, cc)) {
                if (first && k == NSTAGE && cc < NU) ITf(t, k, R::Z + cc) = 0.0;
                continue;
            }
            const double H = hess(tb, We, k, cc), q = LINf(t, k, R::Q + cc);
            const int b = bound_of(cc);
            double z_old, dz;
            if (first) { z_old = (k == 0 && cc >= NU) ? ITf(t, k, R::Z + cc) : 0.0; dz = 0.0; }
            else { z_old = ITf(t, k, R::Z + cc); dz = (k == 0 && cc >= NU) ? 0.0 : STf(t, k, R::DZ + cc); }
            double z_new = z_old + a * dz;
            double lamdiff = 0.0, w1 = 0.0;
            if (b >= 0) {
                const double dl = LINf(t, k, R::DLB + b), du_ = LINf(t, k, R::DUB + b);
                double ll, lu, tl, tu;
                if (first) {
                    double zb = 0.0, t_l = -dl, t_u = du_;
                    if (t_l < o.thr0) {
                        if (t_u < o.thr0) { zb = 0.5 * (dl + du_); t_l = o.thr0; t_u = o.thr0; }
                        else { t_l = o.thr0; zb = dl + o.thr0; }
                    } else if (t_u < o.thr0) { t_u = o.thr0; zb = du_ - o.thr0; }
                    z_new = zb; z_old = zb;
                    tl = t_l; tu = t_u; ll = o.mu0 / t_l; lu = o.mu0 / t_u;
                } else {
                    ll = ITf(t, k, R::LAM + b); lu = ITf(t, k, R::LAM + NB2 + b);
                    tl = ITf(t, k, R::T + b);   tu = ITf(t, k, R::T + NB2 + b);
                    const double mc_l = STf(t, k, R::MC + b), mc_u = STf(t, k, R::MC + NB2 + b);
                    const double rd_l = dl - z_old + tl, rd_u = -du_ + z_old + tu;
                    const double rm_l = ll * tl - o.tau_min + mcw * mc_l - sigmu;
                    const double rm_u = lu * tu - o.tau_min + mcw * mc_u - sigmu;
                    const double dt_l = dz - rd_l, dt_u = -dz - rd_u;
                    const double dl_l = -(ll * dt_l + rm_l) / tl;
                    const double dl_u = -(lu * dt_u + rm_u) / tu;
                    w1 = (lu - ll) - (dl_l - dl_u);
                    ll += a * dl_l; lu += a * dl_u;
                    tl += a * dt_l; tu += a * dt_u;
                }
                ITf(t, k, R::LAM + b) = ll; ITf(t, k, R::LAM + NB2 + b) = lu;
                ITf(t, k, R::T + b) = tl;   ITf(t, k, R::T + NB2 + b) = tu;
                lamdiff = lu - ll;
                const double rd_l = dl - z_new + tl, rd_u = -du_ + z_new + tu;
                const double pm_l = ll * tl, pm_u = lu * tu;
                musum += pm_l + pm_u;
                const double rm_l = pm_l - o.tau_min, rm_u = pm_u - o.tau_min;
                nd = fmax(nd, fmax(fabs(rd_l), fabs(rd_u)));
                nm = fmax(nm, fmax(fabs(rm_l), fabs(rm_u)));
                const double ti_l = tl < o.t_min ? 1.0 / o.t_min : 1.0 / tl;
                const double ti_u = tu < o.t_min ? 1.0 / o.t_min : 1.0 / tu;
                const double l_l = ll < o.lam_min ? o.lam_min : ll;
                const double l_u = lu < o.lam_min ? o.lam_min : lu;
                STf(t, k, R::MC + b) = ti_l * l_l + ti_u * l_u;
                STf(t, k, R::MC + NB2 + b) = ti_l * (rm_l - ll * rd_l) - ti_u * (rm_u - lu * rd_u);
            }
            ITf(t, k, R::Z + cc) = z_new;
            STf(t, k, R::DZA + cc) = q + H * (z_old + dz) + w1;
            STf(t, k, R::DZ + cc) = q + H * z_new + lamdiff;
        }
        // inactive bound slots carry neutral values, as in the throughput path
        COOP_FOR(b, NB2)
        {
            const int k = b < NV ? NSTAGE : 0;
            ITf(t, k, R::LAM + b) = 0.0; ITf(t, k, R::LAM + NB2 + b) = 0.0;
            ITf(t, k, R::T + b) = 1.0;   ITf(t, k, R::T + NB2 + b) = 1.0;
            STf(t, k, R::MC + b) = 0.0;  STf(t, k, R::MC + NB2 + b) = 0.0;
        }
        if (first) { COOP_FOR(it, NK * NX) ITf(t, it / NX, R::PI + it % NX) = 0.0; }
        COOP_SYNC();
        // ---- parallel pass: dynamics residual rb_k = BA_k z_k + b_k - x_{k+1} ----------------------
        double nb = 0.0;
        COOP_FOR(it, NSTAGE * NX)
        {
            const int k = it / NX, i = it % NX;
            const double* E = &LINf(t, k, R::E);
            const double* lti = tb.lti + k * 4 * NV;
            double r = LINf(t, k, R::B0 + i) - ITf(t, k + 1, R::Z + NU + i);
            for (int cc = 0; cc < NZ; cc++) {
                const double bav = ba_entry(E, LANES, lti, i, cc);
                if (bav != 0.0) r += bav * ITf(t, k, R::Z + cc);
            }
            FAf(t, k, R::RB + i) = r;
            nb = fmax(nb, fabs(r));
        }
        COOP_SYNC();
        // ---- sequential backward pass: adjoint recursion + stationarity residual + Riccati ---------
        double ng = 0.0, lru = 0.0;
        COOP_FOR(i, NX) { s.pn[i] = 0.0; s.pw[i] = 0.0; s.pv[i] = 0.0; }
        COOP_SYNC();
        for (int k = NSTAGE; k >= 0; k--) {
            const bool hasU = k < NSTAGE, hasX = k > 0;
            // staged inputs of the stage
            double* in_rgq = s.in;                 // NZ
            double* in_cz = in_rgq + NZ;           // NZ
            double* in_Gam = in_cz + NZ;           // NB2
            double* in_gam = in_Gam + NB2;         // NB2
            double* in_rb = in_gam + NB2;          // NX
            double* in_pi = in_rb + NX;            // NX
            COOP_FOR(j, NZ) { in_rgq[j] = STf(t, k, R::DZ + j); in_cz[j] = STf(t, k, R::DZA + j); }
            COOP_FOR(j, NB2) { in_Gam[j] = STf(t, k, R::MC + j); in_gam[j] = STf(t, k, R::MC + NB2 + j); }
            COOP_FOR(j, NX) { in_rb[j] = hasU ? FAf(t, k, R::RB + j) : 0.0; in_pi[j] = (hasX && !first) ? ITf(t, k, R::PI + j) : 0.0; }
            if (hasU) {
                const double* E = &LINf(t, k, R::E);
                const double* lti = tb.lti + k * 4 * NV;
                COOP_FOR(e, NX * NZ) s.BA[e] = ba_entry(E, LANES, lti, e / NZ, e % NZ);
            }
            COOP_SYNC();
            if (hasU) {
                // phase a: G = P BA, Pb = P rb + pv, t1 = BA' pn, t2 = BA' pw
                COOP_FOR(e, NX * NZ + NX + 2 * NZ)
                {
                    if (e < NX * NZ) {
                        const int i = e / NZ, cc = e % NZ;
                        double acc = 0.0;
                        if (k < NSTAGE - 0 && hasX)   // P is meaningful only when the successor produced it
                            for (int m = 0; m < NX; m++) acc += s.P[i * NX + m] * s.BA[m * NZ + cc];
                        else
                            for (int m = 0; m < NX; m++) acc += s.P[i * NX + m] * s.BA[m * NZ + cc];
                        s.G[e] = acc;
                    } else if (e < NX * NZ + NX) {
                        const int i = e - NX * NZ;
                        double acc = s.pv[i];
                        for (int m = 0; m < NX; m++) acc += s.P[i * NX + m] * in_rb[m];
                        s.Pb[i] = acc;
                    } else {
                        const int q = e - NX * NZ - NX;
                        const int cc = q % NZ;
                        const double* v = q < NZ ? s.pn : s.pw;
                        double acc = 0.0;
                        for (int i = 0; i < NX; i++) acc += s.BA[i * NZ + cc] * v[i];
                        (q < NZ ? s.t1 : s.t2)[cc] = acc;
                    }
                }
            } else {
                COOP_FOR(cc, NZ) { s.t1[cc] = 0.0; s.t2[cc] = 0.0; }
            }
            COOP_SYNC();
            // phase b: multipliers of the dynamics defining x_k, stationarity residual, gradient, M
            const int zlo = hasU ? 0 : NU, zhi = hasX ? NZ : NU;
            COOP_FOR(e, NZ + NZ * NZ)
            {
                if (e < NZ) {
                    const int cc = e;
                    if (cc < zlo || cc >= zhi) continue;
                    double pw_k = 0.0;
                    if (cc >= NU) {
                        const int j = cc - NU;
                        const double pn_k = first ? 0.0 : in_cz[cc] + s.t1[cc];      // pi_old + dpi
                        pw_k = first ? 0.0 : in_pi[j] + a * (pn_k - in_pi[j]);
                        // the successor stage (k-1) needs them: kept in registers until phase d
                        s.xn[j] = pn_k;            // parked; moved to pn/pw after the last reader of t1/t2
                        s.dz[cc] = pw_k;
                        if (hasX) ITf(t, k, R::PI + j) = pw_k;
                    } else if (!first) {
                        lru = fmax(lru, fabs(in_cz[cc] + s.t1[cc]));
                    }
                    const double rg = in_rgq[cc] + s.t2[cc] - (cc >= NU ? pw_k : 0.0);
                    ng = fmax(ng, fabs(rg));
                    const int b = bound_of(cc);
                    double gg = rg + (b >= 0 ? in_gam[b] : 0.0);
                    if (hasU) for (int i = 0; i < NX; i++) gg += s.BA[i * NZ + cc] * s.Pb[i];
                    s.g[cc] = gg;
                } else {
                    const int q = e - NZ;
                    const int r = q / NZ, cc = q % NZ;
                    if (cc > r || r < zlo || r >= zhi || cc < zlo) continue;
                    double acc = 0.0;
                    if (hasU) for (int i = 0; i < NX; i++) acc += s.BA[i * NZ + r] * s.G[i * NZ + cc];
                    if (r == cc) {
                        const int b = bound_of(r);
                        acc += hess(tb, We, k, r) + o.reg_prim + (b >= 0 ? in_Gam[b] : 0.0);
                    }
                    s.Mm[r * NZ + cc] = acc;
                }
            }
            COOP_SYNC();
            // phase c: Cholesky of Muu (every lane, redundantly), lh, K
            if (hasU) {
                double Luu[NLU], lh[NV];
#pragma unroll
                for (int aa = 0; aa < NV; aa++) {
                    double d = s.Mm[aa * NZ + aa];
#pragma unroll
                    for (int q = 0; q < aa; q++) d -= Luu[aa * (aa + 1) / 2 + q] * Luu[aa * (aa + 1) / 2 + q];
                    const double inv = d > 0.0 ? 1.0 / sqrt(d) : 0.0;
                    Luu[aa * (aa + 1) / 2 + aa] = inv;
#pragma unroll
                    for (int bb = aa + 1; bb < NV; bb++) {
                        double v = s.Mm[bb * NZ + aa];
#pragma unroll
                        for (int q = 0; q < aa; q++) v -= Luu[bb * (bb + 1) / 2 + q] * Luu[aa * (aa + 1) / 2 + q];
                        Luu[bb * (bb + 1) / 2 + aa] = v * inv;
                    }
                }
#pragma unroll
                for (int aa = 0; aa < NV; aa++) {
                    double v = s.g[aa];
#pragma unroll
                    for (int q = 0; q < aa; q++) v -= Luu[aa * (aa + 1) / 2 + q] * lh[q];
                    lh[aa] = v * Luu[aa * (aa + 1) / 2 + aa];
                }
                COOP_FOR(e, NLU) FAf(t, k, R::LUU + e) = Luu[e];
                COOP_FOR(e, NV) { FAf(t, k, R::LH + e) = lh[e]; s.lh[e] = lh[e]; }
                if (hasX) {
                    COOP_FOR(j, NX)
                    {
                        double kk[NV];
#pragma unroll
                        for (int aa = 0; aa < NV; aa++) {
                            double v = s.Mm[(NU + j) * NZ + aa];
#pragma unroll
                            for (int q = 0; q < aa; q++) v -= Luu[aa * (aa + 1) / 2 + q] * kk[q];
                            kk[aa] = v * Luu[aa * (aa + 1) / 2 + aa];
                            s.K[aa * NX + j] = kk[aa];
                            FAf(t, k, R::KH + aa * NX + j) = kk[aa];
                        }
                    }
                }
            }
            COOP_SYNC();
            // phase d: Schur complement -> this stage's P and pv; hand the multipliers to stage k-1
            if (hasX) {
                COOP_FOR(e, NX * NX + NX)
                {
                    if (e < NX * NX) {
                        const int i = e / NX, j = e % NX;
                        const int r = i >= j ? i : j, cc = i >= j ? j : i;
                        double v = s.Mm[(NU + r) * NZ + NU + cc];
                        if (hasU) for (int aa = 0; aa < NV; aa++) v -= s.K[aa * NX + r] * s.K[aa * NX + cc];
                        s.P[e] = v;
                    } else {
                        const int i = e - NX * NX;
                        double v = s.g[NU + i];
                        if (hasU) for (int aa = 0; aa < NV; aa++) v -= s.K[aa * NX + i] * s.lh[aa];
                        s.pv[i] = v;
                        s.pn[i] = s.xn[i];
                        s.pw[i] = s.dz[NU + i];
                    }
                }
            }
            COOP_SYNC();
        }
        S_after_B(c, coop_max(ng), coop_max(nb), coop_max(nd), coop_max(nm), coop_sum(musum), coop_max(lru), o, first);
    }

    NMPC_HD static void S_after_B(LaneCtl& c, double ng, double nb, double nd, double nm, double musum, double lru,
                                  const IpmOpts& o, bool first)
    {
        typename S::CarryB cy;
        cy.ng = ng; cy.nb = nb; cy.nd = nd; cy.nm = nm; cy.musum = musum; cy.lru = lru;
        S::after_B(c, cy, o, first);
    }

    // ====================================================================================
    // forward substitution.  delta == false: predictor (FA.LH, with the dynamics residual), writes
    // ST.DZA and ST.MC; delta == true: delta step (FA.LHD, zero equality residual), writes ST.DZ.
    // ====================================================================================
    NMPC_HD static Sums sweep_F(double* t, const Tables& tb, const IpmOpts& o, bool delta, double sigmu, double mcw, Scratch& s)
    {
        COOP_FOR(i, NX) s.xn[i] = 0.0;
        COOP_SYNC();
        for (int k = 0; k <= NSTAGE; k++) {
            const bool hasU = k < NSTAGE, hasX = k > 0;
            double* in_L = s.in;                   // NLU
            double* in_K = in_L + NLU;             // NV*NX
            double* in_lh = in_K + NV * NX;        // NV
            double* in_rb = in_lh + NV;            // NX
            if (hasU) {
                COOP_FOR(e, NLU) in_L[e] = FAf(t, k, R::LUU + e);
                COOP_FOR(e, NV * NX) in_K[e] = FAf(t, k, R::KH + e);
                COOP_FOR(e, NV) in_lh[e] = FAf(t, k, delta ? R::LHD + e : R::LH + e);
                COOP_FOR(e, NX) in_rb[e] = delta ? 0.0 : FAf(t, k, R::RB + e);
                const double* E = &LINf(t, k, R::E);
                const double* lti = tb.lti + k * 4 * NV;
                COOP_FOR(e, NX * NZ) s.BA[e] = ba_entry(E, LANES, lti, e / NZ, e % NZ);
            }
            COOP_SYNC();
            // du = -Luu^-T (lh + K dx): every lane redundantly (NV is 2 or 4)
            double du[NV];
#pragma unroll
            for (int aa = 0; aa < NV; aa++) du[aa] = 0.0;
            if (hasU) {
                double v[NV];
#pragma unroll
                for (int aa = 0; aa < NV; aa++) {
                    double acc = in_lh[aa];
                    if (hasX) for (int j = 0; j < NX; j++) acc += in_K[aa * NX + j] * s.xn[j];
                    v[aa] = -acc;
                }
#pragma unroll
                for (int aa = NV - 1; aa >= 0; aa--) {
                    double acc = v[aa];
#pragma unroll
                    for (int bb = aa + 1; bb < NV; bb++) acc -= in_L[bb * (bb + 1) / 2 + aa] * du[bb];
                    du[aa] = acc * in_L[aa * (aa + 1) / 2 + aa];
                }
            }
            // store the step of this stage; dx_{k+1} = BA [du; dx] + rb
            COOP_FOR(cc, NZ)
            {
                const double v = cc < NU ? du[cc] : s.xn[cc - NU];
                if (!delta) STf(t, k, R::DZA + cc) = v;
                else STf(t, k, R::DZ + cc) = STf(t, k, R::DZA + cc) + v;
            }
            double xnew = 0.0;
            int mine = -1;
            if (hasU) {
                COOP_FOR(i, NX)
                {
                    double acc = in_rb[i];
#pragma unroll
                    for (int cc = 0; cc < NU; cc++) acc += s.BA[i * NZ + cc] * du[cc];
                    for (int j = 0; j < NX; j++) acc += s.BA[i * NZ + NU + j] * s.xn[j];
                    s.Pb[i] = acc;          // staging: xn is still being read by other lanes
                    (void)xnew; (void)mine;
                }
            }
            COOP_SYNC();
            if (hasU) { COOP_FOR(i, NX) s.xn[i] = s.Pb[i]; }
            COOP_SYNC();
        }
        // ---- parallel pass over the bound slots: slack / multiplier steps, ratio test, mu(alpha) sums ----
        Sums r; r.alpha = -1.0; r.S0 = r.S1 = r.S2 = 0.0;
        COOP_FOR(it, NK * NB2)
        {
            const int k = it / NB2, b = it % NB2;
            const int cc = comp_of_bound(b);
            if (!has_comp(k, cc)) continue;
            const double ll = ITf(t, k, R::LAM + b), lu = ITf(t, k, R::LAM + NB2 + b);
            const double tl = ITf(t, k, R::T + b), tu = ITf(t, k, R::T + NB2 + b);
            const double zb = ITf(t, k, R::Z + cc);
            const double dzb = delta ? STf(t, k, R::DZ + cc) : STf(t, k, R::DZA + cc);
            const double rd_l = LINf(t, k, R::DLB + b) - zb + tl, rd_u = -LINf(t, k, R::DUB + b) + zb + tu;
            double rm_l = ll * tl - o.tau_min, rm_u = lu * tu - o.tau_min;
            if (delta) {
                rm_l += mcw * STf(t, k, R::MC + b) - sigmu;
                rm_u += mcw * STf(t, k, R::MC + NB2 + b) - sigmu;
            }
            const double dt_l = dzb - rd_l, dt_u = -dzb - rd_u;
            const double dl_l = -(ll * dt_l + rm_l) / tl, dl_u = -(lu * dt_u + rm_u) / tu;
            if (!delta) { STf(t, k, R::MC + b) = dt_l * dl_l; STf(t, k, R::MC + NB2 + b) = dt_u * dl_u; }
            if (r.alpha * dl_l > ll) r.alpha = ll / dl_l;
            if (r.alpha * dt_l > tl) r.alpha = tl / dt_l;
            if (r.alpha * dl_u > lu) r.alpha = lu / dl_u;
            if (r.alpha * dt_u > tu) r.alpha = tu / dt_u;
            r.S0 += ll * tl + lu * tu;
            r.S1 += ll * dt_l + tl * dl_l + lu * dt_u + tu * dl_u;
            r.S2 += dl_l * dt_l + dl_u * dt_u;
        }
        r.alpha = coop_max_neg(r.alpha);
        r.S0 = coop_sum(r.S0); r.S1 = coop_sum(r.S1); r.S2 = coop_sum(r.S2);
        COOP_SYNC();
        return r;
    }

    // ====================================================================================
    // delta backward substitution: right-hand side only in the complementarity rows; writes FA.LHD
    // ====================================================================================
    NMPC_HD static void sweep_Bd(double* t, const Tables& tb, double sigmu, double mcw, Scratch& s)
    {
        // parallel pre-pass: barrier right-hand sides into ST.DZ (free between B and Fd)
        COOP_FOR(it, NK * NB2)
        {
            const int k = it / NB2, b = it % NB2;
            const int cc = comp_of_bound(b);
            if (!has_comp(k, cc)) continue;
            const double tl = ITf(t, k, R::T + b), tu = ITf(t, k, R::T + NB2 + b);
            STf(t, k, R::DZ + cc) = (mcw * STf(t, k, R::MC + b) - sigmu) / tl - (mcw * STf(t, k, R::MC + NB2 + b) - sigmu) / tu;
        }
        COOP_FOR(i, NX) s.xn[i] = 0.0;
        COOP_SYNC();
        for (int k = NSTAGE; k >= 0; k--) {
            const bool hasU = k < NSTAGE, hasX = k > 0;
            double* in_L = s.in;                   // NLU
            double* in_K = in_L + NLU;             // NV*NX
            double* in_q = in_K + NV * NX;         // NZ: barrier rhs per component (0 where unbounded)
            COOP_FOR(cc, NZ) in_q[cc] = (bound_of(cc) >= 0 && has_comp(k, cc)) ? STf(t, k, R::DZ + cc) : 0.0;
            if (hasU) {
                COOP_FOR(e, NLU) in_L[e] = FAf(t, k, R::LUU + e);
                COOP_FOR(e, NV * NX) in_K[e] = FAf(t, k, R::KH + e);
                const double* E = &LINf(t, k, R::E);
                const double* lti = tb.lti + k * 4 * NV;
                COOP_FOR(e, NX * NZ) s.BA[e] = ba_entry(E, LANES, lti, e / NZ, e % NZ);
            }
            COOP_SYNC();
            COOP_FOR(cc, NZ)
            {
                double acc = in_q[cc];
                if (hasU) for (int i = 0; i < NX; i++) acc += s.BA[i * NZ + cc] * s.xn[i];
                s.g[cc] = acc;
            }
            COOP_SYNC();
            if (hasU) {
                double lh[NV];
#pragma unroll
                for (int aa = 0; aa < NV; aa++) {
                    double v = s.g[aa];
#pragma unroll
                    for (int q = 0; q < aa; q++) v -= in_L[aa * (aa + 1) / 2 + q] * lh[q];
                    lh[aa] = v * in_L[aa * (aa + 1) / 2 + aa];
                }
                COOP_FOR(e, NV) FAf(t, k, R::LHD + e) = lh[e];
                if (hasX) {
                    COOP_FOR(j, NX)
                    {
                        double v = s.g[NU + j];
#pragma unroll
                        for (int aa = 0; aa < NV; aa++) v -= in_K[aa * NX + j] * lh[aa];
                        s.Pb[j] = v;
                    }
                }
            } else {
                COOP_FOR(j, NX) s.Pb[j] = s.g[NU + j];
            }
            COOP_SYNC();
            COOP_FOR(j, NX) s.xn[j] = s.Pb[j];
            COOP_SYNC();
        }
    }

    // ====================================================================================
    // the whole interior-point loop of one instance (all 32 lanes call this with the same arguments)
    // ====================================================================================
    NMPC_HD static void ipm(double* t, const Tables& tb, const double* We, const IpmOpts& o, LaneCtl& c, Scratch& s, bool resume)
    {
        if (!resume) { c.init(true); sweep_B(t, tb, We, o, true, c, s); }
        while (!c.done) {
            {
                const Sums r = sweep_F(t, tb, o, false, 0.0, 0.0, s);
                typename S::CarryF cy; cy.alpha = r.alpha; cy.S0 = r.S0; cy.S1 = r.S1; cy.S2 = r.S2;
                S::after_F(c, cy, o);
            }
            sweep_Bd(t, tb, c.sigmu, 1.0, s);
            {
                const Sums r = sweep_F(t, tb, o, true, c.sigmu, 1.0, s);
                typename S::CarryF cy; cy.alpha = r.alpha; cy.S0 = r.S0; cy.S1 = r.S1; cy.S2 = r.S2;
                S::after_Fd(c, cy, o);
            }
            if (c.fb) {
                sweep_Bd(t, tb, c.sigmu, 0.0, s);
                const Sums r = sweep_F(t, tb, o, true, c.sigmu, 0.0, s);
                typename S::CarryF cy; cy.alpha = r.alpha; cy.S0 = r.S0; cy.S1 = r.S1; cy.S2 = r.S2;
                S::after_Fd_fallback(c, cy);
            }
            sweep_B(t, tb, We, o, false, c, s);
        }
    }
};

}  // namespace nmpc
