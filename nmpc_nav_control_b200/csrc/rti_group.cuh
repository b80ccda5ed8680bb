// K3 "group" path: a persistent, lane-group-cooperative primal-dual interior point.
//
// Replaces (like rti_core.cuh) what the reference reaches through `{m}_acados_solve(capsule)`
// (src/nmpc_nav_control/NMPCNavControlDiff.cpp:142, Omni4.cpp:139, Tric.cpp:146): the HPIPM
// interior-point solve of the SQP-RTI QP.  Same iteration path as
// rti_core.cuh (initial point, Mehrotra predictor / corrector as predictor + delta, conditional
// centering, step rule, exit test; the per-lane control logic Rti<M>::after_* is shared), but a
// different machine mapping:
//
//  * G lanes (8 for nx = 7, 16 for nx = 11) share ONE OCP instance: the components of a stage
//    (state columns of the cost-to-go, the controls, the bound pairs) are spread over the lanes,
//    so the sequential depth of a horizon sweep drops from ~2500 instructions per stage (one
//    thread per instance) to a few hundred, and nothing spills.
//  * a warp holds 32/G instances ("slots") and runs the WHOLE interior-point loop of each on its
//    own; when an instance converges its slot is refilled from a global work queue at the next
//    iteration boundary, so lanes never idle on the iteration-count spread (3..28 iterations) and
//    there is no tail of nearly empty launches.  One launch per chunk instead of two per iteration.
//  * the state of an instance is ONE contiguous record per stage (GRec, all fields of the stage
//    back to back); each sweep prefetches exactly the field ranges it needs for the next stage
//    with 16-byte cp.async copies into a double-buffered shared-memory image of the record while
//    the current stage is being computed.
//
// All cross-lane communication goes through the slot's shared-memory scratch, in bulk-synchronous
// phases (GRP_PHASE_BEGIN .. GRP_PHASE_END = the phase body for this lane, then __syncwarp()).
// That is what lets tests/host_emul run the identical code on the CPU: there a phase is a loop
// over the 32 lanes of an emulated warp.
#pragma once
#include "rti_core.cuh"

#if defined(__CUDA_ARCH__)
#define GRP_NL 1
#define GRP_SYNC() __syncwarp()
#else
#define GRP_NL 32
#define GRP_SYNC() ((void)0)
#endif
#define GRP_PHASE_BEGIN(lanes) { for (int ln_ = 0; ln_ < GRP_NL; ++ln_) { Lane& L = (lanes)[ln_];
#define GRP_PHASE_END } GRP_SYNC(); }

namespace nmpc {

// ---------------------------------------------------------------------------------------------
// record of one (instance, stage): every range a sweep copies starts and ends on 16 bytes
//   [Q B0 | DLB DUB E LHD LUU KH LH RB | DZA MC T LAM Z PI DZ]
//   B  reads [Q, LHD) and [MC, NREC)            writes [LUU, DZA) and [T, DZ)
//   F  reads [DLB, DZA) and [T, PI)             writes [DZA, T)
//   Bd reads [E, LH) and [MC, LAM)              writes LHD
//   Fd reads [DLB, LH) and [DZA, PI)            writes DZ
// component order inside Q, Z, DZ, DZA is z = [u; x] as in Rec<NV>
// ---------------------------------------------------------------------------------------------
template <int NV_>
struct GRec {
    static constexpr int NV = NV_, NX = 3 + 2 * NV, NU = NV, NZ = NX + NU, NC = 1 + 3 * NV, NB2 = 2 * NV,
                         NLU = NV * (NV + 1) / 2;
    static constexpr int ev(int n) { return (n + 1) & ~1; }
    static constexpr int Q = 0;                    // NZ    QP gradient
    static constexpr int B0 = Q + NZ;              // NX    b = phi(x,u) - x_next
    static constexpr int DLB = B0 + NX;            // NB2   lb - z for [u; ref]
    static constexpr int DUB = DLB + NB2;          // NB2   ub - z
    static constexpr int E = DUB + NB2;            // 3*NC  pose rows of [A|B], columns [theta | actual | ref | u]
    static constexpr int LHD = E + ev(3 * NC);     // NV    Luu^-1 q_u (delta)
    static constexpr int LUU = LHD + NV;           // NLU   row-packed lower, diagonal inverted
    static constexpr int KH = LUU + ev(NLU);       // NV*NX K = Luu^-1 S
    static constexpr int LH = KH + NV * NX;        // NV    Luu^-1 q_u (predictor)
    static constexpr int RB = LH + NV;             // NX    dynamics residual
    static constexpr int DZA = RB + ev(NX);        // NZ    predictor step
    static constexpr int MC = DZA + ev(NZ);        // 2*NB2 dt_aff * dlam_aff
    static constexpr int T = MC + 2 * NB2;         // 2*NB2 slacks, lower then upper
    static constexpr int LAM = T + 2 * NB2;        // 2*NB2
    static constexpr int Z = LAM + 2 * NB2;        // NZ    [u; x]
    static constexpr int PI = Z + ev(NZ);          // NX
    static constexpr int DZ = PI + ev(NX);         // NZ    final step
    static constexpr int NREC = DZ + ev(NZ);
    static_assert(NV % 2 == 0, "NV-sized fields must keep 16-byte alignment");
    static_assert(DLB % 2 == 0 && E % 2 == 0 && LHD % 2 == 0 && LUU % 2 == 0 && LH % 2 == 0 && DZA % 2 == 0 && MC % 2 == 0 &&
                  T % 2 == 0 && LAM % 2 == 0 && PI % 2 == 0 && DZ % 2 == 0 && NREC % 2 == 0, "copy ranges must be 16-byte aligned");
    static constexpr size_t inst_doubles = (size_t)(NSTAGE + 1) * NREC;
};

// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only); a plain copy in the host emulation
NMPC_HD void grp_cp16(double* dst, const double* src)
{
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
#else
    dst[0] = src[0]; dst[1] = src[1];
#endif
}
NMPC_HD void grp_cp_commit()
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
}
NMPC_HD void grp_cp_wait_all()
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
}
NMPC_HD int grp_fetch_add(int* ctr)
{
#if defined(__CUDA_ARCH__)
    return atomicAdd(ctr, 1);
#else
    return (*ctr)++;
#endif
}
// uniform (same address for all lanes) read-only table loads
NMPC_HD double grp_ldg(const double* p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

struct GrpOut {            // per-instance results of K3 (global, indexed by instance of the batch)
    int* qp_status;
    int* qp_iter;
    double* stats;         // [8][B] or null
    int B;                 // leading dimension of stats
};

template <class M, int G_>
struct Grp {
    using S = Rti<M>;
    using LaneCtl = typename S::LaneCtl;
    static constexpr int G = G_, NSLOT = 32 / G;
    static constexpr int NV = S::NV, NX = S::NX, NU = S::NU, NZ = S::NZ, NY = S::NY, NC = S::NC, NB2 = S::NB2, NLU = S::NLU;
    static constexpr int CW = (NZ + G - 1) / G;          // z components per lane
    static constexpr int NK = NSTAGE + 1;
    static constexpr int WL = G > NX ? NX : 0;           // lane that writes the per-stage scalars of the factorisation
    static_assert(G >= NX && 32 % G == 0, "one state column per lane");
    using R = GRec<NV>;

    // ---- shared-memory scratch of one slot (doubles) ----------------------------------------
    static constexpr int O_IN = 0;                        // [2][NREC] double-buffered record image
    static constexpr int O_CAR = O_IN + 2 * R::NREC;      // [2][3][NX] carries of the B sweep: pio, dpi, xn
                                                          //   (F: dx [2][NX]; Bd: dp [2][NX])
    static constexpr int O_ZB = O_CAR + 6 * NX;           // NZ   new z of the stage, z order
    static constexpr int O_RB = O_ZB + NZ;                // NX   dynamics residual
    static constexpr int O_PV = O_RB + NX;                // NX   gradient of the cost-to-go
    static constexpr int O_PBA = O_PV + NX;               // [NX][NZ]  P * [A B], columns in w order [x; u]
    static constexpr int O_MUU = O_PBA + NX * NZ;         // [NV][NV]
    static constexpr int O_GU = O_MUU + NV * NV;          // NV   control gradient / F: s_a / Bd: q_u
    static constexpr int O_KB = O_GU + NV;                // [NV][NX]
    static constexpr int O_RED = O_IN;                    // [G][8] reductions at the end of a sweep / queue hand-out (the record image is dead there)
    static constexpr int O_END = O_KB + NV * NX;
    static_assert(G * 8 <= 2 * R::NREC, "reduction buffer must fit into the record image");
    // slot stride: even (16-byte copies) and = 8 mod 16 so that equal offsets of neighbouring slots
    // fall into different bank groups
    static constexpr int SLOT_D = ((O_END + 15) / 16) * 16 + 8;
    static constexpr int WARP_D = SLOT_D * NSLOT;         // doubles of shared memory per warp

    // z-order index of w-order component (w order = [x; u], the order lanes own components in)
    NMPC_HD static constexpr int zc(int w) { return w < NX ? NU + w : w - NX; }
    // bound pair of w-order component: u_a -> a, ref state j >= 3+NV -> j-3, else -1
    NMPC_HD static constexpr int bnd(int w) { return w >= NX ? w - NX : (w >= 3 + NV ? w - 3 : -1); }

    // 1.0 / 0.0 indicator: selecting a register-array element by a lane-dependent index with
    // arithmetic keeps the array in registers (a select chain is turned into a local-memory index)
    NMPC_HD static double sel(bool b) { return b ? 1.0 : 0.0; }

    // (column w of J = [A B]) . v, v indexed by state (stride 1); E = pose rows [3][NC], lt = av|ar|au|ru
    NMPC_HD static double jcol_dot(int w, const double* v, const double* E, const double* lt)
    {
        if (w < NX) {
            const int j = w;
            if (j < 2) return v[j];
            if (j == 2) return E[0] * v[0] + E[NC] * v[1] + E[2 * NC] * v[2];
            if (j < 3 + NV) {
                const int c = j - 3;
                return E[1 + c] * v[0] + E[NC + 1 + c] * v[1] + E[2 * NC + 1 + c] * v[2] + grp_ldg(lt + c) * v[3 + c];
            }
            const int c = j - 3 - NV;
            return E[1 + NV + c] * v[0] + E[NC + 1 + NV + c] * v[1] + E[2 * NC + 1 + NV + c] * v[2] + grp_ldg(lt + NV + c) * v[3 + c] +
                   v[3 + NV + c];
        }
        const int a = w - NX, cu = 1 + 2 * NV + a;
        return E[cu] * v[0] + E[NC + cu] * v[1] + E[2 * NC + cu] * v[2] + grp_ldg(lt + 2 * NV + a) * v[3 + a] +
               grp_ldg(lt + 3 * NV + a) * v[3 + NV + a];
    }
    // same with everything in registers (static indices after unrolling)
    NMPC_HD static double jcol_dot_r(int w, const double* v, const double* E, const double* lt)
    {
        if (w < NX) {
            const int j = w;
            if (j < 2) return v[j];
            if (j == 2) return E[0] * v[0] + E[NC] * v[1] + E[2 * NC] * v[2];
            if (j < 3 + NV) {
                const int c = j - 3;
                return E[1 + c] * v[0] + E[NC + 1 + c] * v[1] + E[2 * NC + 1 + c] * v[2] + lt[c] * v[3 + c];
            }
            const int c = j - 3 - NV;
            return E[1 + NV + c] * v[0] + E[NC + 1 + NV + c] * v[1] + E[2 * NC + 1 + NV + c] * v[2] + lt[NV + c] * v[3 + c] + v[3 + NV + c];
        }
        const int a = w - NX, cu = 1 + 2 * NV + a;
        return E[cu] * v[0] + E[NC + cu] * v[1] + E[2 * NC + cu] * v[2] + lt[2 * NV + a] * v[3 + a] + lt[3 * NV + a] * v[3 + NV + a];
    }
    // (row i of J) . [zu; zx]
    NMPC_HD static double jrow_dot(int i, const double* zu, const double* zx, const double* E, const double* lt)
    {
        if (i < 3) {
            double s = (i < 2 ? zx[i] : 0.0) + E[i * NC] * zx[2];
#pragma unroll
            for (int c = 0; c < NV; c++)
                s += E[i * NC + 1 + c] * zx[3 + c] + E[i * NC + 1 + NV + c] * zx[3 + NV + c] + E[i * NC + 1 + 2 * NV + c] * zu[c];
            return s;
        }
        const int c = i < 3 + NV ? i - 3 : i - 3 - NV;
        double uc = 0.0;        // zu may live in registers: select, do not index dynamically
#pragma unroll
        for (int a = 0; a < NV; a++) uc += zu[a] * sel(a == c);
        if (i < 3 + NV) return grp_ldg(lt + c) * zx[3 + c] + grp_ldg(lt + NV + c) * zx[3 + NV + c] + grp_ldg(lt + 2 * NV + c) * uc;
        return zx[3 + NV + c] + grp_ldg(lt + 3 * NV + c) * uc;
    }

    // ---- state of one lane (registers on the device) ----------------------------------------
    struct Lane {
        int r, slot;              // role in the group, slot of the warp
        int li;                   // instance (index into the chunk) of the slot, -1 = none
        bool act, first, run;     // slot has an instance / its next B sweep is the cold start / takes part in the current sweep
        LaneCtl c;                // replicated over the lanes of the slot
        double astep;             // damped step applied by the running B sweep
        double We;                // terminal weight of this lane's state component
        // B sweep
        double Pc[NX];            // column r of the cost-to-go of the successor stage
        double g[CW], Gam[CW], H[CW];
        double Ef[3 * NC];        // pose rows of [A|B] of the stage
        double lt[4 * NV];
        double Mx[NZ];            // column r of M = J'PJ + D, rows in w order
        double Kc[NV];            // column r of K
        double lh[NV];
        double ng, nb, nd, nm, musum, lru;
        // F / Bd sweeps
        double dxr;               // dx (F) / dp (Bd) component r
        double du[NV];
        double alpha, S0, S1, S2;
    };

    NMPC_HD static double* rec_of(double* ws, int li, int k) { return ws + (size_t)li * R::inst_doubles + (size_t)k * R::NREC; }

    // copy doubles [d0, d1) of the record (16-byte chunks spread over the group)
    NMPC_HD static void copy_range(double* dst, const double* src, int d0, int d1, int r)
    {
        for (int c = d0 + 2 * r; c < d1; c += 2 * G) grp_cp16(dst + c, src + c);
    }

    enum { SW_B = 0, SW_F = 1, SW_BD = 2, SW_FD = 3 };

    template <int KIND>
    NMPC_HD static void issue(const Lane& L, double* ws, double* scr, int k, int buf)
    {
        if (!L.run) return;
        double* dst = scr + O_IN + buf * R::NREC;
        const double* src = rec_of(ws, L.li, k);
        if (KIND == SW_B) {
            copy_range(dst, src, R::Q, R::LHD, L.r);
            if (!L.first) copy_range(dst, src, R::MC, R::NREC, L.r);
        } else if (KIND == SW_F) {
            copy_range(dst, src, R::DLB, R::DZA, L.r);
            copy_range(dst, src, R::T, R::PI, L.r);
        } else if (KIND == SW_BD) {
            copy_range(dst, src, R::E, R::LH, L.r);
            copy_range(dst, src, R::MC, R::LAM, L.r);
        } else {
            copy_range(dst, src, R::DLB, R::LH, L.r);
            copy_range(dst, src, R::DZA, R::PI, L.r);
        }
    }

    // all-gather of up to 8 per-lane partials through the slot scratch: after the call every lane
    // of the slot has reduced q[0..nmax) with max and q[nmax..n) with +
    template <class GF, class PF>
    NMPC_HD static void reduce(Lane* lanes, double* sm, int n, int nmax, GF get, PF put)
    {
        GRP_PHASE_BEGIN(lanes)
            double* scr = sm + L.slot * SLOT_D;
            for (int q = 0; q < n; q++) scr[O_RED + L.r * 8 + q] = get(L, q);
        GRP_PHASE_END
        GRP_PHASE_BEGIN(lanes)
            const double* scr = sm + L.slot * SLOT_D;
            for (int q = 0; q < n; q++) {
                double v = scr[O_RED + q];
                for (int rr = 1; rr < G; rr++) {
                    const double o = scr[O_RED + rr * 8 + q];
                    v = q < nmax ? fmax(v, o) : v + o;
                }
                put(L, q, v);
            }
        GRP_PHASE_END
    }

    // =========================================================================================
    // B sweep: (apply the previous step) + residuals + Riccati factorisation, stages N..0
    // (same arithmetic as Rti::stage_B_update / stage_B_riccati, one component per lane)
    // =========================================================================================
    NMPC_HD static void sweep_B(Lane* lanes, double* sm, double* ws, const Tables& tb, const IpmOpts& o)
    {
        GRP_PHASE_BEGIN(lanes)
            L.ng = L.nb = L.nd = L.nm = L.musum = L.lru = 0.0;
            if (L.run) {
                double* scr = sm + L.slot * SLOT_D;
                if (L.r < NX) { scr[O_CAR + L.r] = 0.0; scr[O_CAR + NX + L.r] = 0.0; scr[O_CAR + 2 * NX + L.r] = 0.0; scr[O_PV + L.r] = 0.0; }
                issue<SW_B>(L, ws, scr, NSTAGE, 0);
            }
            grp_cp_commit();
        GRP_PHASE_END
#pragma unroll 1
        for (int s = 0; s <= NSTAGE; s++) {
            const int k = NSTAGE - s, buf = s & 1;
            const bool hasU = k < NSTAGE, hasX = k > 0;
            const double* ltk = tb.lti + (hasU ? k : 0) * 4 * NV;
            GRP_PHASE_BEGIN(lanes)
                grp_cp_wait_all();
            GRP_PHASE_END
            // ---- B1: prefetch the next stage; per component: step, residuals, barrier terms ----
            GRP_PHASE_BEGIN(lanes)
                if (!L.run) continue;
                double* scr = sm + L.slot * SLOT_D;
                if (s < NSTAGE) issue<SW_B>(L, ws, scr, k - 1, buf ^ 1);
                grp_cp_commit();
                const double* rec = scr + O_IN + buf * R::NREC;
                double* grec = rec_of(ws, L.li, k);
                const double* car = scr + O_CAR + (s & 1) * 3 * NX;          // from stage k+1: pio | dpi | xn
                double* carn = scr + O_CAR + ((s & 1) ^ 1) * 3 * NX;
                const double a_step = L.astep, sigmu = L.c.sigmu, mcw = L.c.mcw;
                const bool first = L.first;
#pragma unroll
                for (int t = 0; t < CW; t++) {
                    const int w = L.r + t * G;
                    if (w >= NZ) continue;
                    const bool isx = w < NX;
                    const int c = zc(w), b = bnd(w);
                    const bool has = isx ? hasX : hasU;
                    const double H = hasU ? tb.dt * grp_ldg(tb.W + k * NY + w)   /* w order = y order [x; u] */ : (isx ? L.We : 0.0);
                    const double q = rec[R::Q + c];
                    double v1 = 0.0, v2 = 0.0;
                    if (hasU) { v1 = jcol_dot(w, car, rec + R::E, ltk); v2 = jcol_dot(w, car + NX, rec + R::E, ltk); }
                    double z, pin = 0.0, pi_old = 0.0, dpi_new = 0.0;
                    double ll = 0.0, lu = 0.0, tl = 1.0, tu = 1.0, dl = 0.0, du_ = 0.0;
                    const bool act = b >= 0 && has;
                    if (b >= 0) { dl = rec[R::DLB + b]; du_ = rec[R::DUB + b]; }
                    if (first) {
                        // cold start (HPIPM INIT_VAR with warm_start = 0)
                        z = (isx && !hasX) ? grec[R::Z + c] : 0.0;
                        if (b >= 0) {
                            double zb = 0.0, t_l = -dl, t_u = du_;
                            if (t_l < o.thr0) {
                                if (t_u < o.thr0) { zb = 0.5 * (dl + du_); t_l = o.thr0; t_u = o.thr0; }
                                else { t_l = o.thr0; zb = dl + o.thr0; }
                            } else if (t_u < o.thr0) { t_u = o.thr0; zb = du_ - o.thr0; }
                            if (act) { z = zb; tl = t_l; tu = t_u; ll = o.mu0 / t_l; lu = o.mu0 / t_u; }
                        }
                    } else {
                        z = rec[R::Z + c];
                        const double dz = has ? rec[R::DZ + c] : 0.0;
                        if (isx && hasX) { pin = rec[R::PI + w]; pi_old = pin; }
                        double ldo = 0.0, dld = 0.0;
                        if (act) {
                            ll = rec[R::LAM + b]; lu = rec[R::LAM + NB2 + b];
                            tl = rec[R::T + b];   tu = rec[R::T + NB2 + b];
                            const double mc_l = rec[R::MC + b], mc_u = rec[R::MC + NB2 + b];
                            const double rd_l = dl - z + tl, rd_u = -du_ + z + tu;
                            const double rm_l = ll * tl - o.tau_min + mcw * mc_l - sigmu;
                            const double rm_u = lu * tu - o.tau_min + mcw * mc_u - sigmu;
                            const double dt_l = dz - rd_l, dt_u = -dz - rd_u;
                            const double dl_l = -(ll * dt_l + rm_l) / tl;
                            const double dl_u = -(lu * dt_u + rm_u) / tu;
                            ldo = lu - ll;
                            dld = dl_l - dl_u;
                            ll += a_step * dl_l; lu += a_step * dl_u;
                            tl += a_step * dt_l; tu += a_step * dt_u;
                        }
                        if (!isx) {
                            if (hasU) {
                                const double r = q + H * z + ldo + v1 + H * dz - dld + v2;
                                L.lru = fmax(L.lru, fabs(r));
                            }
                        } else if (hasX) {
                            // adjoint recursion for the multiplier step of the dynamics that define x_k
                            double r = q + H * z - pin + v1 + H * dz + v2;
                            if (b >= 0) r += ldo - dld;
                            dpi_new = r;
                            pin += a_step * r;
                        }
                        z += a_step * dz;
                    }
                    // residuals at the (new) iterate
                    double g = q + H * z + (lu - ll) + (v1 + a_step * v2);
                    if (isx) g -= pin;
                    if (has) L.ng = fmax(L.ng, fabs(g));
                    double Gam = 0.0;
                    if (act) {
                        const double rd_l = dl - z + tl, rd_u = -du_ + z + tu;
                        const double pm_l = ll * tl, pm_u = lu * tu;
                        L.musum += pm_l + pm_u;
                        const double rm_l = pm_l - o.tau_min, rm_u = pm_u - o.tau_min;
                        L.nd = fmax(L.nd, fmax(fabs(rd_l), fabs(rd_u)));
                        L.nm = fmax(L.nm, fmax(fabs(rm_l), fabs(rm_u)));
                        const double ti_l = tl < o.t_min ? 1.0 / o.t_min : 1.0 / tl;
                        const double ti_u = tu < o.t_min ? 1.0 / o.t_min : 1.0 / tu;
                        const double l_l = ll < o.lam_min ? o.lam_min : ll;
                        const double l_u = lu < o.lam_min ? o.lam_min : lu;
                        Gam = ti_l * l_l + ti_u * l_u;
                        g += ti_l * (rm_l - ll * rd_l) - ti_u * (rm_u - lu * rd_u);
                    }
                    L.g[t] = g; L.Gam[t] = Gam; L.H[t] = H;
                    // store the iterate, hand the carries to stage k-1
                    if (!isx || hasX) grec[R::Z + c] = z;
                    if (isx && hasX) grec[R::PI + w] = pin;
                    if (b >= 0) {
                        grec[R::LAM + b] = ll; grec[R::LAM + NB2 + b] = lu;
                        grec[R::T + b] = tl;   grec[R::T + NB2 + b] = tu;
                    }
                    scr[O_ZB + c] = z;
                    if (isx) { carn[w] = pi_old; carn[NX + w] = dpi_new; carn[2 * NX + w] = z; }
                }
            GRP_PHASE_END
            if (!hasU) {
                // terminal stage: P = diag(We + reg + Gamma), p = g
                GRP_PHASE_BEGIN(lanes)
                    if (!L.run || L.r >= NX) continue;
                    double* scr = sm + L.slot * SLOT_D;
                    const double pd = L.We + o.reg_prim + L.Gam[0];
#pragma unroll
                    for (int i = 0; i < NX; i++) L.Pc[i] = pd * sel(i == L.r);
                    scr[O_PV + L.r] = L.g[0];
                GRP_PHASE_END
                continue;
            }
            // ---- B2: dynamics residual (row r) and row r of P * [A B] ----------------------------
            GRP_PHASE_BEGIN(lanes)
                if (!L.run) continue;
                double* scr = sm + L.slot * SLOT_D;
                const double* rec = scr + O_IN + buf * R::NREC;
#pragma unroll
                for (int i = 0; i < 3 * NC; i++) L.Ef[i] = rec[R::E + i];
#pragma unroll
                for (int i = 0; i < 4 * NV; i++) L.lt[i] = grp_ldg(ltk + i);
                if (L.r >= NX) continue;
                const double* car = scr + O_CAR + (s & 1) * 3 * NX;
                const double rb = jrow_dot(L.r, scr + O_ZB, scr + O_ZB + NU, rec + R::E, ltk) + rec[R::B0 + L.r] - car[2 * NX + L.r];
                L.nb = fmax(L.nb, fabs(rb));
                scr[O_RB + L.r] = rb;
                rec_of(ws, L.li, k)[R::RB + L.r] = rb;
#pragma unroll
                for (int w = 0; w < NZ; w++)
                    if (hasX || w >= NX) scr[O_PBA + L.r * NZ + w] = jcol_dot_r(w, L.Pc, L.Ef, L.lt);
            GRP_PHASE_END
            // ---- B4: column of M = J'PJ + D per component; gradient += J'(P rb + p) --------------
            GRP_PHASE_BEGIN(lanes)
                if (!L.run) continue;
                double* scr = sm + L.slot * SLOT_D;
#pragma unroll
                for (int t = 0; t < CW; t++) {
                    const int w = L.r + t * G;
                    if (w >= NZ) continue;
                    const bool isx = w < NX;
                    if (isx && !hasX) continue;
                    double col[NX];
#pragma unroll
                    for (int i = 0; i < NX; i++) col[i] = scr[O_PBA + i * NZ + w];
                    double gg = L.g[t] + jcol_dot(w, scr + O_PV, scr + O_IN + buf * R::NREC + R::E, ltk);
#pragma unroll
                    for (int i = 0; i < NX; i++) gg += col[i] * scr[O_RB + i];
                    L.g[t] = gg;
                    if (isx) {
                        const double dg = L.H[t] + o.reg_prim + L.Gam[t];
#pragma unroll
                        for (int wp = 0; wp < NZ; wp++) L.Mx[wp] = jcol_dot_r(wp, col, L.Ef, L.lt) + (wp < NX ? dg * sel(wp == w) : 0.0);
                    } else {
                        const int a = w - NX;
#pragma unroll
                        for (int ap = 0; ap < NV; ap++) {
                            double m = jcol_dot_r(NX + ap, col, L.Ef, L.lt);
                            if (ap == a) m += L.H[t] + o.reg_prim + L.Gam[t];
                            scr[O_MUU + ap * NV + a] = m;
                        }
                        scr[O_GU + a] = gg;
                    }
                }
            GRP_PHASE_END
            // ---- B5: Cholesky of the control block (every lane), lh, column r of K ----------------
            GRP_PHASE_BEGIN(lanes)
                if (!L.run) continue;
                double* scr = sm + L.slot * SLOT_D;
                double* grec = rec_of(ws, L.li, k);
                double Luu[NLU];
#pragma unroll
                for (int a = 0; a < NV; a++) {
                    double d = scr[O_MUU + a * NV + a];
#pragma unroll
                    for (int c = 0; c < a; c++) d -= Luu[a * (a + 1) / 2 + c] * Luu[a * (a + 1) / 2 + c];
                    const double inv = d > 0.0 ? 1.0 / sqrt(d) : 0.0;
                    Luu[a * (a + 1) / 2 + a] = inv;
#pragma unroll
                    for (int b = a + 1; b < NV; b++) {
                        double sacc = scr[O_MUU + b * NV + a];
#pragma unroll
                        for (int c = 0; c < a; c++) sacc -= Luu[b * (b + 1) / 2 + c] * Luu[a * (a + 1) / 2 + c];
                        Luu[b * (b + 1) / 2 + a] = sacc * inv;
                    }
                }
#pragma unroll
                for (int a = 0; a < NV; a++) {
                    double sacc = scr[O_GU + a];
#pragma unroll
                    for (int c = 0; c < a; c++) sacc -= Luu[a * (a + 1) / 2 + c] * L.lh[c];
                    L.lh[a] = sacc * Luu[a * (a + 1) / 2 + a];
                }
                // the factor goes out through one lane that has no state column (static indices only:
                // spreading Luu[i] over lanes by role turns into a local-memory jump table)
                if (L.r == WL) {
#pragma unroll
                    for (int i = 0; i < NLU; i++) grec[R::LUU + i] = Luu[i];
#pragma unroll
                    for (int a = 0; a < NV; a++) grec[R::LH + a] = L.lh[a];
                }
                if (hasX && L.r < NX) {
#pragma unroll
                    for (int a = 0; a < NV; a++) {
                        double sacc = L.Mx[NX + a];
#pragma unroll
                        for (int c = 0; c < a; c++) sacc -= Luu[a * (a + 1) / 2 + c] * L.Kc[c];
                        L.Kc[a] = sacc * Luu[a * (a + 1) / 2 + a];
                        scr[O_KB + a * NX + L.r] = L.Kc[a];
                        grec[R::KH + a * NX + L.r] = L.Kc[a];
                    }
                }
            GRP_PHASE_END
            if (!hasX) continue;
            // ---- B6: Schur complement -> column r of this stage's cost-to-go, and its gradient -----
            GRP_PHASE_BEGIN(lanes)
                if (!L.run || L.r >= NX) continue;
                double* scr = sm + L.slot * SLOT_D;
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    double sacc = L.Mx[i];
#pragma unroll
                    for (int a = 0; a < NV; a++) sacc -= scr[O_KB + a * NX + i] * L.Kc[a];
                    L.Pc[i] = sacc;
                }
                double pvn = L.g[0];
#pragma unroll
                for (int a = 0; a < NV; a++) pvn -= L.Kc[a] * L.lh[a];
                scr[O_PV + L.r] = pvn;
            GRP_PHASE_END
        }
        reduce(lanes, sm, 6, 5,
               [](const Lane& L, int q) { return q == 0 ? L.ng : q == 1 ? L.nb : q == 2 ? L.nd : q == 3 ? L.nm : q == 4 ? L.lru : L.musum; },
               [](Lane& L, int q, double v) { if (q == 0) L.ng = v; else if (q == 1) L.nb = v; else if (q == 2) L.nd = v; else if (q == 3) L.nm = v; else if (q == 4) L.lru = v; else L.musum = v; });
    }

    // =========================================================================================
    // forward sweeps (Rti::stage_F).  delta == false: predictor, writes DZA and MC;
    // delta == true: predictor + delta step, writes DZ.  mcw = 0 is the pure-centering repeat.
    // =========================================================================================
    template <bool DELTA>
    NMPC_HD static void sweep_F(Lane* lanes, double* sm, double* ws, const Tables& tb, const IpmOpts& o, double mcw)
    {
        constexpr int KIND = DELTA ? SW_FD : SW_F;
        GRP_PHASE_BEGIN(lanes)
            L.alpha = -1.0; L.S0 = L.S1 = L.S2 = 0.0; L.dxr = 0.0;
            if (L.run) {
                double* scr = sm + L.slot * SLOT_D;
                if (L.r < NX) scr[O_CAR + L.r] = 0.0;
                issue<KIND>(L, ws, scr, 0, 0);
            }
            grp_cp_commit();
        GRP_PHASE_END
#pragma unroll 1
        for (int k = 0; k <= NSTAGE; k++) {
            const int buf = k & 1;
            const bool hasU = k < NSTAGE, hasX = k > 0;
            const double* ltk = tb.lti + (hasU ? k : 0) * 4 * NV;
            GRP_PHASE_BEGIN(lanes)
                grp_cp_wait_all();
            GRP_PHASE_END
            // ---- F1: s_a = lh_a + K_a . dx --------------------------------------------------------
            GRP_PHASE_BEGIN(lanes)
                if (!L.run) continue;
                double* scr = sm + L.slot * SLOT_D;
                if (k < NSTAGE) issue<KIND>(L, ws, scr, k + 1, buf ^ 1);
                grp_cp_commit();
                if (!hasU) continue;
                const double* rec = scr + O_IN + buf * R::NREC;
                const double* dx = scr + O_CAR + (k & 1) * NX;
#pragma unroll
                for (int t = 0; t < CW; t++) {
                    const int w = L.r + t * G;
                    if (w < NX || w >= NZ) continue;
                    const int a = w - NX;
                    double sacc = rec[(DELTA ? R::LHD : R::LH) + a];
                    if (hasX) {
#pragma unroll
                        for (int j = 0; j < NX; j++) sacc += rec[R::KH + a * NX + j] * dx[j];
                    }
                    scr[O_GU + a] = sacc;
                }
            GRP_PHASE_END
            // ---- F2: du (every lane), own step component, bound pair, next dx (row r) ----------------
            GRP_PHASE_BEGIN(lanes)
                if (!L.run) continue;
                double* scr = sm + L.slot * SLOT_D;
                const double* rec = scr + O_IN + buf * R::NREC;
                double* grec = rec_of(ws, L.li, k);
                const double sigmu = L.c.sigmu;
#pragma unroll
                for (int a = 0; a < NV; a++) L.du[a] = 0.0;
                if (hasU) {
#pragma unroll
                    for (int a = NV - 1; a >= 0; a--) {
                        double sacc = -scr[O_GU + a];
#pragma unroll
                        for (int b = a + 1; b < NV; b++) sacc -= rec[R::LUU + b * (b + 1) / 2 + a] * L.du[b];
                        L.du[a] = sacc * rec[R::LUU + a * (a + 1) / 2 + a];
                    }
                }
#pragma unroll
                for (int t = 0; t < CW; t++) {
                    const int w = L.r + t * G;
                    if (w >= NZ) continue;
                    const bool isx = w < NX;
                    const int c = zc(w), b = bnd(w);
                    double dzw = 0.0;
                    if (isx) dzw = L.dxr;
                    else {
#pragma unroll
                        for (int a = 0; a < NV; a++) dzw += L.du[a] * sel(a == w - NX);
                    }
                    if (!DELTA) grec[R::DZA + c] = dzw;
                    else { dzw += rec[R::DZA + c]; grec[R::DZ + c] = dzw; }
                    const bool act = b >= 0 && (isx ? hasX : hasU);
                    if (act) {
                        const double ll = rec[R::LAM + b], lu = rec[R::LAM + NB2 + b];
                        const double tl = rec[R::T + b], tu = rec[R::T + NB2 + b];
                        const double zb = rec[R::Z + c];
                        const double rd_l = rec[R::DLB + b] - zb + tl, rd_u = -rec[R::DUB + b] + zb + tu;
                        double rm_l = ll * tl - o.tau_min, rm_u = lu * tu - o.tau_min;
                        if (DELTA) {
                            rm_l += mcw * rec[R::MC + b] - sigmu;
                            rm_u += mcw * rec[R::MC + NB2 + b] - sigmu;
                        }
                        const double dt_l = dzw - rd_l, dt_u = -dzw - rd_u;
                        const double dl_l = -(ll * dt_l + rm_l) / tl, dl_u = -(lu * dt_u + rm_u) / tu;
                        if (!DELTA) { grec[R::MC + b] = dt_l * dl_l; grec[R::MC + NB2 + b] = dt_u * dl_u; }
                        if (L.alpha * dl_l > ll) L.alpha = ll / dl_l;
                        if (L.alpha * dt_l > tl) L.alpha = tl / dt_l;
                        if (L.alpha * dl_u > lu) L.alpha = lu / dl_u;
                        if (L.alpha * dt_u > tu) L.alpha = tu / dt_u;
                        L.S0 += ll * tl + lu * tu;
                        L.S1 += ll * dt_l + tl * dl_l + lu * dt_u + tu * dl_u;
                        L.S2 += dl_l * dt_l + dl_u * dt_u;
                    }
                }
                if (hasU && L.r < NX) {
                    const double* dx = scr + O_CAR + (k & 1) * NX;
                    double xn = jrow_dot(L.r, L.du, dx, rec + R::E, ltk);
                    if (!DELTA) xn += rec[R::RB + L.r];
                    L.dxr = xn;
                    scr[O_CAR + ((k & 1) ^ 1) * NX + L.r] = xn;
                }
            GRP_PHASE_END
        }
        reduce(lanes, sm, 4, 1,
               [](const Lane& L, int q) { return q == 0 ? L.alpha : q == 1 ? L.S0 : q == 2 ? L.S1 : L.S2; },
               [](Lane& L, int q, double v) { if (q == 0) L.alpha = v; else if (q == 1) L.S0 = v; else if (q == 2) L.S1 = v; else L.S2 = v; });
    }

    // =========================================================================================
    // delta backward sweep (Rti::stage_Bd): right-hand side only in the complementarity rows
    // =========================================================================================
    NMPC_HD static void sweep_Bd(Lane* lanes, double* sm, double* ws, const Tables& tb, double mcw)
    {
        GRP_PHASE_BEGIN(lanes)
            L.dxr = 0.0;
            if (L.run) {
                double* scr = sm + L.slot * SLOT_D;
                if (L.r < NX) scr[O_CAR + L.r] = 0.0;
                issue<SW_BD>(L, ws, scr, NSTAGE, 0);
            }
            grp_cp_commit();
        GRP_PHASE_END
#pragma unroll 1
        for (int s = 0; s <= NSTAGE; s++) {
            const int k = NSTAGE - s, buf = s & 1;
            const bool hasU = k < NSTAGE, hasX = k > 0;
            const double* ltk = tb.lti + (hasU ? k : 0) * 4 * NV;
            GRP_PHASE_BEGIN(lanes)
                grp_cp_wait_all();
            GRP_PHASE_END
            // ---- D1: q = J' dp + complementarity terms ---------------------------------------------
            GRP_PHASE_BEGIN(lanes)
                if (!L.run) continue;
                double* scr = sm + L.slot * SLOT_D;
                if (s < NSTAGE) issue<SW_BD>(L, ws, scr, k - 1, buf ^ 1);
                grp_cp_commit();
                const double* rec = scr + O_IN + buf * R::NREC;
                const double* dp = scr + O_CAR + (s & 1) * NX;
                const double sigmu = L.c.sigmu;
#pragma unroll
                for (int t = 0; t < CW; t++) {
                    const int w = L.r + t * G;
                    if (w >= NZ) continue;
                    const bool isx = w < NX;
                    const int b = bnd(w);
                    double qv = hasU ? jcol_dot(w, dp, rec + R::E, ltk) : 0.0;
                    if (b >= 0 && (isx ? hasX : hasU)) {
                        const double tl = rec[R::T + b], tu = rec[R::T + NB2 + b];
                        qv += (mcw * rec[R::MC + b] - sigmu) / tl - (mcw * rec[R::MC + NB2 + b] - sigmu) / tu;
                    }
                    if (isx) L.g[0] = qv; else scr[O_GU + (w - NX)] = qv;
                }
            GRP_PHASE_END
            // ---- D2: lh = L^-1 q_u (every lane), dp = q_x - K' lh -------------------------------------
            GRP_PHASE_BEGIN(lanes)
                if (!L.run) continue;
                double* scr = sm + L.slot * SLOT_D;
                const double* rec = scr + O_IN + buf * R::NREC;
                double* dpn = scr + O_CAR + ((s & 1) ^ 1) * NX;
                if (hasU) {
                    double* grec = rec_of(ws, L.li, k);
#pragma unroll
                    for (int a = 0; a < NV; a++) {
                        double sacc = scr[O_GU + a];
#pragma unroll
                        for (int c = 0; c < a; c++) sacc -= rec[R::LUU + a * (a + 1) / 2 + c] * L.lh[c];
                        L.lh[a] = sacc * rec[R::LUU + a * (a + 1) / 2 + a];
                    }
                    if (L.r == WL) {
#pragma unroll
                        for (int a = 0; a < NV; a++) grec[R::LHD + a] = L.lh[a];
                    }
                    if (hasX && L.r < NX) {
                        double sacc = L.g[0];
#pragma unroll
                        for (int a = 0; a < NV; a++) sacc -= rec[R::KH + a * NX + L.r] * L.lh[a];
                        dpn[L.r] = sacc;
                    }
                } else if (L.r < NX) dpn[L.r] = L.g[0];
            GRP_PHASE_END
        }
    }

    // =========================================================================================
    // the interior-point loop of one warp: NSLOT instances in lockstep, slots refilled from the
    // work queue `next` (instances [0, n) of the chunk) at every iteration boundary
    // =========================================================================================
    NMPC_HD static void run_warp(Lane* lanes, double* sm, double* ws, int i0, int n, int* next, const Tables& tb,
                                 const double* We_inst, int ldWe, const IpmOpts& o, const GrpOut& out)
    {
#pragma unroll 1
        for (;;) {
            // ---- refill free slots --------------------------------------------------------------
            GRP_PHASE_BEGIN(lanes)
                if (!L.act && L.r == 0) {
                    int* q = reinterpret_cast<int*>(sm + L.slot * SLOT_D + O_RED);
                    *q = grp_fetch_add(next);
                }
            GRP_PHASE_END
            GRP_PHASE_BEGIN(lanes)
                if (!L.act) {
                    const int idx = *reinterpret_cast<const int*>(sm + L.slot * SLOT_D + O_RED);
                    if (idx < n) {
                        L.act = true; L.first = true; L.li = idx; L.c.init(true);
                        L.We = L.r < NX ? (We_inst ? We_inst[(size_t)L.r * ldWe + i0 + idx] : grp_ldg(tb.We + L.r)) : 0.0;
                    }
                }
                L.run = L.act;
                L.astep = 0.0;
                if (L.act && !L.first) L.astep = S::before_B(L.c);
            GRP_PHASE_END
            if (!warp_any(lanes, [](const Lane& L) { return L.act; })) break;

            sweep_B(lanes, sm, ws, tb, o);
            GRP_PHASE_BEGIN(lanes)
                if (!L.act) continue;
                typename S::CarryB cy;
                cy.ng = L.ng; cy.nb = L.nb; cy.nd = L.nd; cy.nm = L.nm; cy.musum = L.musum; cy.lru = L.lru;
                S::after_B(L.c, cy, o, L.first);
                L.first = false;
                if (L.c.done) {
                    if (L.r == 0) {
                        const int i = i0 + L.li;
                        out.qp_status[i] = L.c.status;
                        out.qp_iter[i] = L.c.iter;
                        if (out.stats) {
                            for (int q = 0; q < 4; q++) out.stats[(size_t)q * out.B + i] = L.c.nrm[q];
                            out.stats[(size_t)4 * out.B + i] = L.c.mu;
                            out.stats[(size_t)5 * out.B + i] = L.c.lin_res;
                            out.stats[(size_t)6 * out.B + i] = (double)L.c.nfb;
                            out.stats[(size_t)7 * out.B + i] = (double)L.c.status;
                        }
                    }
                    L.act = false;
                }
                L.run = L.act;
            GRP_PHASE_END
            if (!warp_any(lanes, [](const Lane& L) { return L.act; })) continue;

            sweep_F<false>(lanes, sm, ws, tb, o, 1.0);
            GRP_PHASE_BEGIN(lanes)
                if (!L.act) continue;
                typename S::CarryF cy;
                cy.alpha = L.alpha; cy.S0 = L.S0; cy.S1 = L.S1; cy.S2 = L.S2;
                S::after_F(L.c, cy, o);
            GRP_PHASE_END
            sweep_Bd(lanes, sm, ws, tb, 1.0);
            sweep_F<true>(lanes, sm, ws, tb, o, 1.0);
            GRP_PHASE_BEGIN(lanes)
                if (!L.act) continue;
                typename S::CarryF cy;
                cy.alpha = L.alpha; cy.S0 = L.S0; cy.S1 = L.S1; cy.S2 = L.S2;
                S::after_Fd(L.c, cy, o);
                L.run = L.c.fb != 0;
            GRP_PHASE_END
            if (warp_any(lanes, [](const Lane& L) { return L.act && L.run; })) {
                // conditional centering (rare): repeat the delta solve without the second-order term
                GRP_PHASE_BEGIN(lanes)
                    L.run = L.act && L.c.fb != 0;
                GRP_PHASE_END
                sweep_Bd(lanes, sm, ws, tb, 0.0);
                sweep_F<true>(lanes, sm, ws, tb, o, 0.0);
                GRP_PHASE_BEGIN(lanes)
                    if (!L.run) continue;
                    typename S::CarryF cy;
                    cy.alpha = L.alpha; cy.S0 = L.S0; cy.S1 = L.S1; cy.S2 = L.S2;
                    S::after_Fd_fallback(L.c, cy);
                GRP_PHASE_END
            }
        }
    }

    template <class F>
    NMPC_HD static bool warp_any(Lane* lanes, F f)
    {
#if defined(__CUDA_ARCH__)
        return __any_sync(0xffffffffu, f(lanes[0]));
#else
        bool r = false;
        for (int ln = 0; ln < GRP_NL; ln++) r = r || f(lanes[ln]);
        return r;
#endif
    }

    NMPC_HD static void init_lane(Lane& L, int lane)
    {
        L.r = lane % G; L.slot = lane / G; L.li = -1;
        L.act = false; L.first = false; L.run = false;
        L.c.init(false);
        L.astep = 0.0; L.We = 0.0;
    }
};

}  // namespace nmpc
