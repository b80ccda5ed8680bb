// K3 "coop" path: the persistent lane-cooperative primal-dual interior point, second mapping.
//
// Replaces (like rti_core.cuh) what the reference reaches through
// `{m}_acados_solve(capsule)` (src/nmpc_nav_control/NMPCNavControlDiff.cpp:142, Omni4.cpp:139,
// Tric.cpp:146): the HPIPM interior-point solve of the SQP-RTI QP.
// Same iteration path as the per-lane sweeps of rti_core.cuh (initial point, Mehrotra predictor / corrector as predictor +
// delta, conditional centering, step rule, exit test; the control logic Rti<M>::after_* is shared), per-instance records
// (GRec, rti_records.cuh), a work queue that refills a slot as soon as its instance converges; the G = 4 NV lanes of an
// instance divide a stage like this:
//
//  * every lane runs the SAME straight-line code; a lane's role is data (a column of E, three
//    coefficients of the stage table, a shuffle source lane), never a branch.  Lane r < NX owns
//    state r (its step, its column of the cost-to-go, of M and of K, its row of the dynamics);
//    every lane owns exactly one one-sided constraint (G = 2 * 2 NV of them per stage), the two
//    sides of a bounded component sitting XM = G/2 lanes apart and the lower side of a bounded
//    state on that state's own lane; the NV controls are replicated on all lanes.
//  * reductions over the group (K dx, the pose rows of [A B] z, the ratio test, the norms, the
//    mu sums) are xor-shuffle butterflies over the group's own lane mask; single remote values
//    (the lag partner of a state, a bounded state's step) are indexed shuffles.  Shared memory
//    is used only where a lane needs a whole vector of the others (the adjoint carries, P*rb+p,
//    the transposed P*[A B], K), written once and read back with 16-byte loads.
//  * the forward sweeps need no shared-memory exchange at all.
//
// The host emulation (tests/host_emul) runs the identical code: a phase is a loop over the 32
// lanes of an emulated warp, a shuffle reads the neighbour's field as the previous phase left it.
// Hence the one rule of this file: a field read through CO_SHFL / CO_SHFLX in a phase is never
// written in that phase.
#pragma once
#include "rti_records.cuh"

#if defined(__CUDA_ARCH__)
#define CO_SHFL(field, srclane) __shfl_sync(0xffffffffu, L.field, (srclane))
#define CO_SHFLX(field, m) __shfl_xor_sync(0xffffffffu, L.field, (m))
#define CO_PHASE_END_NS } }                     // results leave the phase through registers / shuffles only
#else
#define CO_SHFL(field, srclane) (lanes[(srclane)].field)
#define CO_SHFLX(field, m) (lanes[ln_ ^ (m)].field)
#define CO_PHASE_END_NS } }
#endif
#define CO_UNROLL _Pragma("unroll")
// the pose rows of [A B] and the lag constants of a stage for the Riccati phases: register copies (16-byte loads, fewer
// shared-memory instructions, ~60 more live registers) or read where used
#ifndef NMPC_COOP_EREG
#define NMPC_COOP_EREG 0
#endif
#if NMPC_COOP_EREG
#define CO_LOAD_E(rec, ltk) double Ef[EP], lt[4 * NV]; ldv(Ef, (rec) + R::E); ldv(lt, (ltk));
#else
#define CO_LOAD_E(rec, ltk) const double* Ef = (rec) + R::E; const double* lt = (ltk);
#endif

// butterfly all-reduce of the array field A[0..n) over the G lanes of a group; T is a scratch field of the same shape.
// OP(x, y) combines; the result is left in A on every lane (bitwise identical on all lanes: the tree is symmetric).
#define CO_ALLRED(A, T, n, OP)                                                                                        \
    GRP_PHASE_BEGIN(lanes) CO_UNROLL for (int a_ = 0; a_ < (n); a_++) L.T[a_] = OP(L.A[a_], CO_SHFLX(A[a_], 1)); CO_PHASE_END_NS \
    GRP_PHASE_BEGIN(lanes) CO_UNROLL for (int a_ = 0; a_ < (n); a_++) L.A[a_] = OP(L.T[a_], CO_SHFLX(T[a_], 2)); CO_PHASE_END_NS \
    GRP_PHASE_BEGIN(lanes) CO_UNROLL for (int a_ = 0; a_ < (n); a_++) L.T[a_] = OP(L.A[a_], CO_SHFLX(A[a_], 4)); CO_PHASE_END_NS \
    if (G == 16) { GRP_PHASE_BEGIN(lanes) CO_UNROLL for (int a_ = 0; a_ < (n); a_++) L.A[a_] = OP(L.T[a_], CO_SHFLX(T[a_], 8)); CO_PHASE_END_NS } \
    else { GRP_PHASE_BEGIN(lanes) CO_UNROLL for (int a_ = 0; a_ < (n); a_++) L.A[a_] = L.T[a_]; CO_PHASE_END_NS }
#define CO_ADD(x, y) ((x) + (y))
#define CO_MAX(x, y) ((x) > (y) ? (x) : (y))

namespace nmpc {

template <class M, int G_>
struct Coop {
    using S = Rti<M>;
    using LaneCtl = typename S::LaneCtl;
    using GP = RecOps<M>;                        // record helpers and the stage-table geometry
    static constexpr int G = G_, NSLOT = 32 / G, XM = G / 2;
    static constexpr int NV = S::NV, NX = S::NX, NU = S::NU, NZ = S::NZ, NY = S::NY, NC = S::NC, NB2 = S::NB2, NLU = S::NLU;
    static constexpr int NCT = 2 * NB2;
    static_assert(G == NCT && G > NX && (G == 8 || G == 16), "one one-sided constraint per lane, one state per lane");
    using R = GRec<NV>;
    static constexpr int LT_ZERO = GP::LT_ZERO, LT_ONE = GP::LT_ONE, T_W = GP::T_W, TROW = GP::TROW;

    enum { SW_B = 0, SW_F = 1, SW_BD = 2, SW_FD = 3 };
    NMPC_HD static constexpr int imax(int a, int b) { return a > b ? a : b; }
    template <int KIND> struct Img {
        static constexpr int A0 = KIND == SW_B ? R::Q : KIND == SW_BD ? R::E : R::DLB;
        static constexpr int A1 = KIND == SW_B ? R::LHD : KIND == SW_F ? R::DZA : R::LH;
        static constexpr int B0 = KIND == SW_B ? R::MC : KIND == SW_F ? R::T : KIND == SW_BD ? R::MC : R::DZA;
        static constexpr int B1 = KIND == SW_B ? R::NREC : KIND == SW_BD ? R::LAM : R::PI;
        static constexpr int SIZE = (A1 - A0) + (B1 - B0);
        static constexpr int D = NMPC_GRP_DEPTH;
        NMPC_HD static const double* a(const double* img) { return img - A0; }
        NMPC_HD static const double* b(const double* img) { return img + (A1 - A0) - B0; }
        NMPC_HD static double* a(double* img) { return img - A0; }
        NMPC_HD static double* b(double* img) { return img + (A1 - A0) - B0; }
    };

    // ---- shared-memory scratch of one slot (doubles) ----------------------------------------
    static constexpr int NXP = (NX + 1) & ~1;           // vectors over the states, padded to 16 bytes
    static constexpr int EP = (3 * NC + 1) & ~1;        // the pose rows of [A B] of a record, padded
    static constexpr int YS = NXP + 2;                  // row stride of the transposed P*[A B]: 16-byte rows, bank-spread columns
    static constexpr int DMAX = imax(2, NMPC_GRP_DEPTH);
    static constexpr int O_IN = 0;
    static constexpr int RING = DMAX * imax(imax(Img<SW_B>::SIZE, Img<SW_F>::SIZE), imax(Img<SW_BD>::SIZE, Img<SW_FD>::SIZE));
    static constexpr int O_V1 = O_IN + RING;            // [2][NXP] old multipliers of the successor stage (B);  dp (Bd)
    static constexpr int O_V2 = O_V1 + 2 * NXP;         // [2][NXP] multiplier step of the successor stage
    static constexpr int O_RBV = O_V2 + 2 * NXP;        // [NXP]    dynamics residual
    static constexpr int O_TV = O_RBV + NXP;            // [NXP]    P rb + p
    static constexpr int O_CU = O_TV + NXP;             // [NV][4]  summed constraint terms of the controls
    static constexpr int O_Y = O_CU + 4 * NV;           // [NZ][YS] (P [A B])', column w of it in row w
    static constexpr int O_KB = O_Y + NZ * YS;          // [NV][NXP] K
    static constexpr int O_AST = O_KB + NV * NXP;       // damped step handed from lane 0 to the slot (+ queue hand-out word)
    static constexpr int O_CTL = O_AST + 2;
    static constexpr int CTL_D = (int)((sizeof(LaneCtl) + 7) / 8);
    static constexpr int O_END = O_CTL + CTL_D;
    static constexpr int SLOT_D = ((O_END + 15) / 16) * 16 + 8;
    static constexpr int O_TAB = SLOT_D * NSLOT;        // [DMAX][TROW] ring of stage-table rows (per warp)
    static constexpr int WARP_D = SLOT_D * NSLOT + DMAX * TROW;

    struct Lane {
        int r, so, li, gi;
        bool act, first, run, skipB;
        double sigmu, mcw, astep;
        int to, wl, gb;
        double We_j;
        double* grec;
        const double* tsrc;
        // state role
        bool isx, bnd;            // owns a state; that state is a bounded (reference) state
        int ecol;                 // column of E of the own state (-1: x / y, unit column; lanes without a state: -1 and u0 = u1 = 0)
        double un0, un1;          // unit column of x / y
        int pf, pt, ptx;          // lanes of the forward / transpose lag partner; state index of the transpose partner
        int ksf, kpf, kuf, kst, kpt;   // offsets into the stage-table row
        int cj;                   // channel of the own state (selects the control of the forward lag row)
        // constraint role
        int cidx, zoff, dboff, csrc, ua;
        bool is_u;
        double sg;
        int ul[NV];               // a lane that holds the summed constraint terms of control a
        // sweep state
        double Pc[NX], pv, xn, zn, g, dg, gxt, tt, e0, e1, e2;
        double zu[NV], gu[NV], dgu[NV];
        double Mxx[NX], Mux[NV], Kc[NV], lh[NV];
        double c4[4], cs[4];
        double pp[3], pq[3], rd[NV + 3], rq[NV + 3];
        double dx, dxs, dxp, du[NV];
        double dp, gt, gs2;
        double nr[6], nq[6];      // ng, nb, nd, nm, lru (max) and musum (sum) / alpha, S0, S1, S2
        double aN, aD;
    };

    NMPC_HD static void init_lane(Lane& L, int lane, int warp)
    {
        const int r = lane % G;
        L.r = grp_pin(r); L.gb = grp_pin(lane - r);
        L.so = grp_pin(warp * WARP_D + (lane / G) * SLOT_D); L.to = grp_pin(warp * WARP_D + O_TAB); L.wl = grp_pin(lane);
        L.li = -1; L.gi = -1; L.act = L.first = L.run = L.skipB = false;
        L.sigmu = 0.0; L.mcw = 1.0; L.astep = 0.0; L.We_j = 0.0; L.grec = nullptr; L.tsrc = nullptr;
        // state role
        L.isx = r < NX; L.bnd = r >= 3 + NV && r < NX;
        L.ecol = -1; L.un0 = r == 0 ? 1.0 : 0.0; L.un1 = r == 1 ? 1.0 : 0.0;
        L.pf = lane; L.pt = lane; L.ptx = r < NX ? r : 0; L.cj = 0;
        L.ksf = L.kpf = L.kuf = L.kst = L.kpt = LT_ZERO;
        if (r == 2) L.ecol = 0;
        else if (r >= 3 && r < 3 + NV) {
            const int c = r - 3;
            L.ecol = 1 + c; L.cj = c; L.ksf = c; L.kpf = NV + c; L.kuf = 2 * NV + c; L.kst = c; L.pf = lane + NV;
        } else if (r >= 3 + NV && r < NX) {
            const int c = r - 3 - NV;
            L.ecol = 1 + NV + c; L.cj = c; L.ksf = LT_ONE; L.kuf = 3 * NV + c; L.kst = LT_ONE; L.kpt = NV + c; L.pt = lane - NV; L.ptx = r - NV;
        }
        // constraint role: reference state c lives on lane 3 + NV + c and hosts its own lower side, the upper side sits
        // XM lanes away; the controls take the remaining lane pairs in order
        const int rl = r % XM;
        int b = -1; bool lower = false;
        for (int c = 0; c < NV; c++)
            if ((3 + NV + c) % XM == rl) { b = NV + c; lower = (r == 3 + NV + c); }
        if (b < 0) {
            int a = 0;
            for (int q = 0; q < rl; q++) {
                bool ref = false;
                for (int c = 0; c < NV; c++) ref = ref || ((3 + NV + c) % XM == q);
                if (!ref) a++;
            }
            b = a; lower = r < XM;
        }
        L.is_u = b < NV; L.ua = b < NV ? b : 0;
        L.sg = lower ? 1.0 : -1.0;
        L.cidx = lower ? b : NB2 + b;
        L.dboff = lower ? R::DLB + b : R::DUB + b;
        L.zoff = b < NV ? b : NU + 3 + b;
        L.csrc = b < NV ? lane : L.gb + 3 + b;
        for (int a = 0; a < NV; a++) {
            int cnt = 0, lane_a = 0;
            for (int q = 0; q < XM; q++) {
                bool ref = false;
                for (int c = 0; c < NV; c++) ref = ref || ((3 + NV + c) % XM == q);
                if (!ref) { if (cnt == a) lane_a = q; cnt++; }
            }
            L.ul[a] = grp_pin(L.gb + lane_a);
        }
        L.ecol = grp_pin(L.ecol); L.pf = grp_pin(L.pf); L.pt = grp_pin(L.pt); L.ptx = grp_pin(L.ptx); L.cj = grp_pin(L.cj);
        L.ksf = grp_pin(L.ksf); L.kpf = grp_pin(L.kpf); L.kuf = grp_pin(L.kuf); L.kst = grp_pin(L.kst); L.kpt = grp_pin(L.kpt);
        L.cidx = grp_pin(L.cidx); L.dboff = grp_pin(L.dboff); L.zoff = grp_pin(L.zoff); L.csrc = grp_pin(L.csrc); L.ua = grp_pin(L.ua);
        L.pv = L.xn = L.zn = L.g = L.dg = L.gxt = L.tt = L.e0 = L.e1 = L.e2 = 0.0;
        L.dx = L.dxs = L.dxp = L.dp = L.gt = L.gs2 = 0.0; L.aN = 1.0; L.aD = -1.0;
        for (int i = 0; i < NX; i++) { L.Pc[i] = 0.0; L.Mxx[i] = 0.0; }
        for (int a = 0; a < NV; a++) { L.zu[a] = L.gu[a] = L.dgu[a] = L.Mux[a] = L.Kc[a] = L.lh[a] = L.du[a] = 0.0; }
        for (int q = 0; q < 4; q++) { L.c4[q] = L.cs[q] = 0.0; }
        for (int q = 0; q < 3; q++) { L.pp[q] = L.pq[q] = 0.0; }
        for (int q = 0; q < NV + 3; q++) { L.rd[q] = L.rq[q] = 0.0; }
        for (int q = 0; q < 6; q++) { L.nr[q] = L.nq[q] = 0.0; }
    }

    NMPC_HD static double* rec_of(double* ws, int li, int k) { return ws + (size_t)li * R::inst_doubles + (size_t)k * R::NREC; }
    NMPC_HD static void tile_to_record(const double* tl, int k, double* rec, const double* thr_k) { GP::tile_to_record(tl, k, rec, thr_k); }

    NMPC_HD static void copy_range(double* dst, const double* src, int d0, int d1, int r)
    {
#pragma unroll
        for (int c = d0; c < d1; c += 2 * G)
            if (c + 2 * r < d1) grp_cp16(dst + c + 2 * r, src + c + 2 * r);
    }
    NMPC_HD static void issue_tab(const Lane& L, double* sm, const double* src, int slot)
    {
        if (2 * L.wl < TROW) grp_cp16(sm + L.to + slot * TROW + 2 * L.wl, src);
    }
    template <int KIND>
    NMPC_HD static void issue(const Lane& L, const double* src, double* img)
    {
        copy_range(Img<KIND>::a(img), src, Img<KIND>::A0, Img<KIND>::A1, L.r);
        copy_range(Img<KIND>::b(img), src, Img<KIND>::B0, Img<KIND>::B1, L.r);
    }
    template <int KIND, int DIR>
    NMPC_HD static void prefetch(const Lane& L, double* sm, int ahead, int slot, bool valid)
    {
        if (valid) {
            issue_tab(L, sm, L.tsrc + DIR * ahead * TROW, slot);
            if (L.run) issue<KIND>(L, L.grec + DIR * ahead * R::NREC, sm + L.so + O_IN + slot * Img<KIND>::SIZE);
        }
        grp_cp_commit();
    }
    template <int KIND, int DIR>
    NMPC_HD static void begin_sweep(Lane& L, double* sm, double* ws, const Tables& tb, int k0)
    {
        L.tsrc = grp_pin_ptr(tb.stg + (size_t)k0 * TROW + 2 * L.wl - DIR * TROW);
        if (L.run) L.grec = grp_pin_ptr(rec_of(ws, L.li, k0) - DIR * R::NREC);
#pragma unroll
        for (int j = 0; j < Img<KIND>::D - 1; j++) prefetch<KIND, DIR>(L, sm, j + 1, j, true);
    }
    template <int KIND, int DIR>
    NMPC_HD static void begin_stage(Lane& L)
    {
        L.grec += DIR * R::NREC; L.tsrc += DIR * TROW;
        grp_cp_wait<Img<KIND>::D - 2>();
    }

    // N doubles (N even) from a 16-byte aligned address with 16-byte loads
    template <int N>
    NMPC_HD static void ldv(double (&dst)[N], const double* p)
    {
        static_assert(N % 2 == 0, "pairs");
#if defined(__CUDA_ARCH__)
#pragma unroll
        for (int i = 0; i < N; i += 2) { const double2 v = *reinterpret_cast<const double2*>(p + i); dst[i] = v.x; dst[i + 1] = v.y; }
#else
        for (int i = 0; i < N; i++) dst[i] = p[i];
#endif
    }
    // selects v[idx] of a small register array with a static loop (keeps the array in registers)
    template <int N>
    NMPC_HD static double pick(const double (&v)[N], int idx)
    {
        double o = v[0];
#pragma unroll
        for (int a = 1; a < N; a++) o = (idx == a) ? v[a] : o;
        return o;
    }
    // (column of E of the own state) . v[0..3): x / y lanes carry their unit column
    NMPC_HD static void own_ecol(const Lane& L, const double* E, double& e0, double& e1, double& e2)
    {
        const bool h = L.ecol >= 0;
        const int q = h ? L.ecol : 0;
        e0 = h ? E[q] : L.un0; e1 = h ? E[NC + q] : L.un1; e2 = h ? E[2 * NC + q] : 0.0;
    }

    // =========================================================================================
    // B sweep: apply the previous step, residuals, Riccati factorisation, stages N..0
    // (the arithmetic of Rti::stage_B_update / stage_B_riccati)
    // =========================================================================================
    NMPC_HD static void sweep_B(Lane* lanes, double* sm, double* ws, const Tables& tb, const IpmOpts& o)
    {
        constexpr int ISZ = Img<SW_B>::SIZE, D = Img<SW_B>::D;
        GRP_PHASE_BEGIN(lanes)
#pragma unroll
            for (int q = 0; q < 6; q++) L.nr[q] = 0.0;
            L.pv = 0.0; L.xn = 0.0;
            if (L.run) {
                double* scr = sm + L.so;
                if (L.r < NXP) { scr[O_V1 + L.r] = 0.0; scr[O_V1 + NXP + L.r] = 0.0; scr[O_V2 + L.r] = 0.0; scr[O_V2 + NXP + L.r] = 0.0; }
            }
            begin_sweep<SW_B, -1>(L, sm, ws, tb, NSTAGE);
        GRP_PHASE_END
#pragma unroll 1
        for (int s = 0; s <= NSTAGE; s++) {
            const int k = NSTAGE - s, slot = s % D, pslot = (s + D - 1) % D, par = s & 1;
            const bool hasU = k < NSTAGE, hasX = k > 0, pvalid = s + D - 1 <= NSTAGE;
            GRP_PHASE_BEGIN(lanes)
                begin_stage<SW_B, -1>(L);
            GRP_PHASE_END
            // ---- U1: prefetch; one one-sided constraint per lane -------------------------------------------------
            GRP_PHASE_BEGIN(lanes)
                prefetch<SW_B, -1>(L, sm, D - 1, pslot, pvalid);
                const double* img = sm + L.so + O_IN + slot * ISZ;
                const double* rec = Img<SW_B>::a(img); const double* rec2 = Img<SW_B>::b(img);
                double rpart = 0.0, lpart = 0.0, gpart = 0.0, Gp = 0.0;
                if (L.is_u ? hasU : hasX) {
                    const double a_step = L.astep, sg = L.sg;
                    const int c = L.cidx;
                    const double dbd = rec[L.dboff], z = rec2[R::Z + L.zoff], dz = rec2[R::DZ + L.zoff];
                    const double lam = rec2[R::LAM + c], tt = rec2[R::T + c], mc = rec2[R::MC + c];
                    const double rd = sg * (dbd - z) + tt;
                    const double rm = lam * tt - o.tau_min + L.mcw * mc - L.sigmu;
                    const double dt = sg * dz - rd;
                    const double dlam = -(lam * dt + rm) / tt;
                    const double lam_n = lam + a_step * dlam, t_n = tt + a_step * dt, zn = z + a_step * dz;
                    rpart = -sg * (lam + dlam);
                    const double rd_n = sg * (dbd - zn) + t_n;
                    const double pm = lam_n * t_n;
                    L.nr[5] += pm;
                    const double rm_n = pm - o.tau_min;
                    L.nr[2] = grp_maxabs(L.nr[2], rd_n);
                    L.nr[3] = grp_maxabs(L.nr[3], rm_n);
                    const double ti = t_n < o.t_min ? 1.0 / o.t_min : 1.0 / t_n;
                    const double lc = lam_n < o.lam_min ? o.lam_min : lam_n;
                    Gp = ti * lc;
                    lpart = -sg * lam_n;
                    gpart = sg * (ti * (rm_n - lam_n * rd_n));
                    if (L.run) { L.grec[R::LAM + c] = lam_n; L.grec[R::T + c] = t_n; }
                }
                L.c4[0] = rpart; L.c4[1] = lpart; L.c4[2] = gpart; L.c4[3] = Gp;
            CO_PHASE_END_NS
            // ---- U2: the two sides of a bounded component meet; the controls' sums go to the slot -------------------
            GRP_PHASE_BEGIN(lanes)
#pragma unroll
                for (int q = 0; q < 4; q++) L.cs[q] = L.c4[q] + CO_SHFLX(c4[q], XM);
                if (L.is_u && L.r < XM) {
                    double* cu = sm + L.so + O_CU + 4 * L.ua;
                    cu[0] = L.cs[0]; cu[1] = L.cs[1]; cu[2] = L.cs[2]; cu[3] = L.cs[3];
                }
            GRP_PHASE_END
            // ---- U3: stationarity of the own state and of the (replicated) controls, new iterate, adjoint carries ----
            GRP_PHASE_BEGIN(lanes)
                double* scr = sm + L.so;
                const double* img = scr + O_IN + slot * ISZ;
                const double* rec = Img<SW_B>::a(img); const double* rec2 = Img<SW_B>::b(img);
                const double* ltk = sm + L.to + slot * TROW;
                const double* v1 = scr + O_V1 + par * NXP; const double* v2 = scr + O_V2 + par * NXP;
                double* v1n = scr + O_V1 + (par ^ 1) * NXP; double* v2n = scr + O_V2 + (par ^ 1) * NXP;
                const double a_step = L.astep;
                double p1[NXP], p2[NXP];
                ldv(p1, v1); ldv(p2, v2);
                double e0, e1, e2;
                own_ecol(L, rec + R::E, e0, e1, e2);
                L.e0 = e0; L.e1 = e1; L.e2 = e2;
                L.pp[0] = L.pp[1] = L.pp[2] = 0.0;
                if (L.isx) {
                    const double kst = ltk[L.kst], kpt = ltk[L.kpt];
                    const double v1j = e0 * p1[0] + e1 * p1[1] + e2 * p1[2] + kst * v1[L.r] + kpt * v1[L.ptx];
                    const double v2j = e0 * p2[0] + e1 * p2[1] + e2 * p2[2] + kst * v2[L.r] + kpt * v2[L.ptx];
                    const double H = hasU ? tb.dt * ltk[T_W + L.r] : L.We_j;
                    const double qv = rec[R::Q + NU + L.r], z = rec2[R::Z + NU + L.r], dz = rec2[R::DZ + NU + L.r];
                    const double pin = hasX ? rec2[R::PI + L.r] : 0.0;
                    double rr = qv + H * z - pin + v1j + H * dz + v2j;
                    if (L.bnd) rr += L.cs[0];
                    const double pin_n = hasX ? pin + a_step * rr : 0.0;
                    const double zn = z + a_step * dz;
                    double g = qv + H * zn - pin_n + (v1j + a_step * v2j);
                    if (L.bnd) g += L.cs[1];
                    if (hasX) L.nr[0] = grp_maxabs(L.nr[0], g);
                    if (L.bnd) g += L.cs[2];
                    L.g = g; L.dg = H + o.reg_prim + (L.bnd ? L.cs[3] : 0.0);
                    L.zn = zn;
                    if (L.run) { L.grec[R::Z + NU + L.r] = zn; L.grec[R::PI + L.r] = pin_n; }
                    v1n[L.r] = pin; v2n[L.r] = hasX ? rr : 0.0;
                    L.pp[0] = e0 * zn; L.pp[1] = e1 * zn; L.pp[2] = e2 * zn;
                }
#pragma unroll
                for (int a = 0; a < NV; a++) {
                    const double* cu = scr + O_CU + 4 * a;
                    const int q = 1 + 2 * NV + a;
                    const double u0 = rec[R::E + q], u1 = rec[R::E + NC + q], u2 = rec[R::E + 2 * NC + q];
                    const double au = ltk[2 * NV + a], ru = ltk[3 * NV + a];
                    const double v1u = u0 * p1[0] + u1 * p1[1] + u2 * p1[2] + au * p1[3 + a] + ru * p1[3 + NV + a];
                    const double v2u = u0 * p2[0] + u1 * p2[1] + u2 * p2[2] + au * p2[3 + a] + ru * p2[3 + NV + a];
                    const double H = hasU ? tb.dt * ltk[T_W + NX + a] : 0.0;
                    const double qv = rec[R::Q + a], z = rec2[R::Z + a], dz = rec2[R::DZ + a];
                    const double rr = qv + H * z + v1u + H * dz + v2u + cu[0];
                    const double zn = z + a_step * dz;
                    double g = qv + H * zn + (v1u + a_step * v2u) + cu[1];
                    if (hasU) { L.nr[0] = grp_maxabs(L.nr[0], g); L.nr[4] = grp_maxabs(L.nr[4], rr); }
                    g += cu[2];
                    L.zu[a] = zn; L.gu[a] = g; L.dgu[a] = H + o.reg_prim + cu[3];
                }
                if (L.run && L.r < NV) L.grec[R::Z + L.r] = pick(L.zu, L.r);
            GRP_PHASE_END
            if (!hasU) {
                // terminal stage: P = diag(We + reg + Gamma), p = g
                GRP_PHASE_BEGIN(lanes)
#pragma unroll
                    for (int i = 0; i < NX; i++) L.Pc[i] = (i == L.r) ? L.dg : 0.0;
                    L.pv = L.g; L.xn = L.zn;
                CO_PHASE_END_NS
                continue;
            }
            // ---- the pose rows of [A B] z over the group --------------------------------------------------------------
            CO_ALLRED(pp, pq, 3, CO_ADD)
            // ---- U4: dynamics residual, row r -----------------------------------------------------------------------
            GRP_PHASE_BEGIN(lanes)
                double* scr = sm + L.so;
                const double* img = scr + O_IN + slot * ISZ;
                const double* rec = Img<SW_B>::a(img);
                const double* ltk = sm + L.to + slot * TROW;
                const double znp = CO_SHFL(zn, L.pf);
                if (L.isx) {
                    double ps[3];
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        double sacc = L.pp[i];
#pragma unroll
                        for (int a = 0; a < NV; a++) sacc += rec[R::E + i * NC + 1 + 2 * NV + a] * L.zu[a];
                        ps[i] = sacc;
                    }
                    const double psel = L.r == 0 ? ps[0] : L.r == 1 ? ps[1] : L.r == 2 ? ps[2] : 0.0;
                    const double jz = psel + ltk[L.ksf] * L.zn + ltk[L.kpf] * znp + ltk[L.kuf] * pick(L.zu, L.cj);
                    const double rb = jz + rec[R::B0 + L.r] - L.xn;
                    L.nr[1] = grp_maxabs(L.nr[1], rb);
                    if (L.run) L.grec[R::RB + L.r] = rb;
                    scr[O_RBV + L.r] = rb;
                    L.xn = L.zn;
                }
            GRP_PHASE_END
            // ---- R1: t = P rb + p, row r of P [A B] ----------------------------------------------------------------
            GRP_PHASE_BEGIN(lanes)
                double* scr = sm + L.so;
                const double* img = scr + O_IN + slot * ISZ;
                const double* rec = Img<SW_B>::a(img);
                const double* ltk = sm + L.to + slot * TROW;
                if (L.isx) {
                    double rbv[NXP];
                    ldv(rbv, scr + O_RBV);
                    CO_LOAD_E(rec, ltk)
                    double t0 = L.pv, t1 = 0.0;
#pragma unroll
                    for (int m = 0; m < NX; m++) { if (m & 1) t1 += L.Pc[m] * rbv[m]; else t0 += L.Pc[m] * rbv[m]; }
                    L.tt = t0 + t1;
                    scr[O_TV + L.r] = L.tt;
#pragma unroll
                    for (int w = 0; w < NZ; w++)
                        if (hasX || w >= NX) scr[O_Y + w * YS + L.r] = GP::jcol_dot_r(w, L.Pc, Ef, lt);
                }
            GRP_PHASE_END
            // ---- R2: column r of M = [A B]' P [A B] + D, gradients, Cholesky of the control block, K ------------------
            GRP_PHASE_BEGIN(lanes)
                double* scr = sm + L.so;
                const double* img = scr + O_IN + slot * ISZ;
                const double* rec = Img<SW_B>::a(img);
                const double* ltk = sm + L.to + slot * TROW;
                double tv[NXP];
                ldv(tv, scr + O_TV);
                CO_LOAD_E(rec, ltk)
                if (hasX && L.isx) {
                    double col[NXP];
                    ldv(col, scr + O_Y + L.r * YS);
                    L.gxt = L.g + (L.e0 * tv[0] + L.e1 * tv[1] + L.e2 * tv[2] + ltk[L.kst] * L.tt + ltk[L.kpt] * scr[O_TV + L.ptx]);
#pragma unroll
                    for (int i = 0; i < NX; i++) L.Mxx[i] = GP::jcol_dot_r(i, col, Ef, lt) + ((i == L.r) ? L.dg : 0.0);
#pragma unroll
                    for (int a = 0; a < NV; a++) L.Mux[a] = GP::jcol_dot_r(NX + a, col, Ef, lt);
                }
                double Muu[NLU], gg[NV];
#pragma unroll
                for (int b = 0; b < NV; b++) {
                    double colu[NXP];
                    ldv(colu, scr + O_Y + (NX + b) * YS);
                    gg[b] = L.gu[b] + GP::jcol_dot_r(NX + b, tv, Ef, lt);
#pragma unroll
                    for (int a = b; a < NV; a++)
                        Muu[a * (a + 1) / 2 + b] = GP::jcol_dot_r(NX + a, colu, Ef, lt) + (a == b ? L.dgu[b] : 0.0);
                }
                double Luu[NLU];
#pragma unroll
                for (int a = 0; a < NV; a++) {
                    double d = Muu[a * (a + 1) / 2 + a];
#pragma unroll
                    for (int c = 0; c < a; c++) d -= Luu[a * (a + 1) / 2 + c] * Luu[a * (a + 1) / 2 + c];
                    const double inv = d > 0.0 ? grp_rsqrt(d) : 0.0;
                    Luu[a * (a + 1) / 2 + a] = inv;
#pragma unroll
                    for (int b = a + 1; b < NV; b++) {
                        double sacc = Muu[b * (b + 1) / 2 + a];
#pragma unroll
                        for (int c = 0; c < a; c++) sacc -= Luu[b * (b + 1) / 2 + c] * Luu[a * (a + 1) / 2 + c];
                        Luu[b * (b + 1) / 2 + a] = sacc * inv;
                    }
                }
#pragma unroll
                for (int a = 0; a < NV; a++) {
                    double sacc = gg[a];
#pragma unroll
                    for (int c = 0; c < a; c++) sacc -= Luu[a * (a + 1) / 2 + c] * L.lh[c];
                    L.lh[a] = sacc * Luu[a * (a + 1) / 2 + a];
                }
                if (L.run && L.r == NX) {                         // a lane without a state writes the per-stage scalars
#pragma unroll
                    for (int i = 0; i < NLU; i++) L.grec[R::LUU + i] = Luu[i];
#pragma unroll
                    for (int a = 0; a < NV; a++) L.grec[R::LH + a] = L.lh[a];
                }
                if (hasX && L.isx) {
#pragma unroll
                    for (int a = 0; a < NV; a++) {
                        double sacc = L.Mux[a];
#pragma unroll
                        for (int c = 0; c < a; c++) sacc -= Luu[a * (a + 1) / 2 + c] * L.Kc[c];
                        L.Kc[a] = sacc * Luu[a * (a + 1) / 2 + a];
                        scr[O_KB + a * NXP + L.r] = L.Kc[a];
                        if (L.run) L.grec[R::KH + a * NX + L.r] = L.Kc[a];
                    }
                }
            GRP_PHASE_END
            if (!hasX) continue;
            // ---- R3: Schur complement -> column r of this stage's cost-to-go, and its gradient -----------------------
            GRP_PHASE_BEGIN(lanes)
                if (!L.isx) continue;
                const double* scr = sm + L.so;
#pragma unroll
                for (int i = 0; i < NX; i++) L.Pc[i] = L.Mxx[i];
#pragma unroll
                for (int a = 0; a < NV; a++) {
                    double kr[NXP];
                    ldv(kr, scr + O_KB + a * NXP);
#pragma unroll
                    for (int i = 0; i < NX; i++) L.Pc[i] -= kr[i] * L.Kc[a];
                }
                double pvn = L.gxt;
#pragma unroll
                for (int a = 0; a < NV; a++) pvn -= L.Kc[a] * L.lh[a];
                L.pv = pvn;
            CO_PHASE_END_NS
        }
        // norms (max) and the complementarity sum over the group
        GRP_PHASE_BEGIN(lanes)
            L.nq[5] = L.nr[5] + CO_SHFLX(nr[5], 1);
#pragma unroll
            for (int q = 0; q < 5; q++) { const double ov = CO_SHFLX(nr[q], 1); L.nq[q] = L.nr[q] > ov ? L.nr[q] : ov; }
        CO_PHASE_END_NS
        GRP_PHASE_BEGIN(lanes)
            L.nr[5] = L.nq[5] + CO_SHFLX(nq[5], 2);
#pragma unroll
            for (int q = 0; q < 5; q++) { const double ov = CO_SHFLX(nq[q], 2); L.nr[q] = L.nq[q] > ov ? L.nq[q] : ov; }
        CO_PHASE_END_NS
        GRP_PHASE_BEGIN(lanes)
            L.nq[5] = L.nr[5] + CO_SHFLX(nr[5], 4);
#pragma unroll
            for (int q = 0; q < 5; q++) { const double ov = CO_SHFLX(nr[q], 4); L.nq[q] = L.nr[q] > ov ? L.nr[q] : ov; }
        CO_PHASE_END_NS
        if (G == 16) {
            GRP_PHASE_BEGIN(lanes)
                L.nr[5] = L.nq[5] + CO_SHFLX(nq[5], 8);
#pragma unroll
                for (int q = 0; q < 5; q++) { const double ov = CO_SHFLX(nq[q], 8); L.nr[q] = L.nq[q] > ov ? L.nq[q] : ov; }
            CO_PHASE_END_NS
        } else {
            GRP_PHASE_BEGIN(lanes)
#pragma unroll
                for (int q = 0; q < 6; q++) L.nr[q] = L.nq[q];
            CO_PHASE_END_NS
        }
    }

    // =========================================================================================
    // forward sweeps (Rti::stage_F).  DELTA == false: predictor, writes DZA and MC;
    // DELTA == true: predictor + delta step, writes DZ.  mcw = 0 is the pure-centering repeat.
    // No shared-memory exchange: group reductions and indexed shuffles only.
    // =========================================================================================
    template <bool DELTA>
    NMPC_HD static void sweep_F(Lane* lanes, double* sm, double* ws, const Tables& tb, const IpmOpts& o, double mcw)
    {
        constexpr int KIND = DELTA ? SW_FD : SW_F;
        constexpr int ISZ = Img<KIND>::SIZE, D = Img<KIND>::D;
        GRP_PHASE_BEGIN(lanes)
            L.aN = 1.0; L.aD = -1.0; L.nr[1] = L.nr[2] = L.nr[3] = 0.0; L.dx = 0.0;
            begin_sweep<KIND, 1>(L, sm, ws, tb, 0);
        GRP_PHASE_END
#pragma unroll 1
        for (int k = 0; k <= NSTAGE; k++) {
            const int slot = k % D, pslot = (k + D - 1) % D;
            const bool hasU = k < NSTAGE, hasX = k > 0, pvalid = k + D - 1 <= NSTAGE;
            GRP_PHASE_BEGIN(lanes)
                begin_stage<KIND, 1>(L);
            GRP_PHASE_END
            // ---- F1: prefetch; partial products of K dx and of the pose rows of [A B] [dx; .]; remote dx values ------
            GRP_PHASE_BEGIN(lanes)
                prefetch<KIND, 1>(L, sm, D - 1, pslot, pvalid);
                const double* rec = Img<KIND>::a(sm + L.so + O_IN + slot * ISZ);
                double e0, e1, e2;
                own_ecol(L, rec + R::E, e0, e1, e2);
                const double dxm = L.isx ? L.dx : 0.0;
#pragma unroll
                for (int a = 0; a < NV; a++) L.rd[a] = (hasU && hasX && L.isx) ? rec[R::KH + a * NX + (L.isx ? L.r : 0)] * dxm : 0.0;
                L.rd[NV] = e0 * dxm; L.rd[NV + 1] = e1 * dxm; L.rd[NV + 2] = e2 * dxm;
                L.dxs = CO_SHFL(dx, L.csrc);
                L.dxp = CO_SHFL(dx, L.pf);
            CO_PHASE_END_NS
            CO_ALLRED(rd, rq, NV + 3, CO_ADD)
            // ---- F2: du (every lane); one constraint per lane: ratio test, mu sums; step stores; next dx (row r) -------
            GRP_PHASE_BEGIN(lanes)
                const double* img = sm + L.so + O_IN + slot * ISZ;
                const double* rec = Img<KIND>::a(img); const double* rec2 = Img<KIND>::b(img);
                const double* ltk = sm + L.to + slot * TROW;
#pragma unroll
                for (int a = 0; a < NV; a++) L.du[a] = 0.0;
                if (hasU) {
#pragma unroll
                    for (int a = NV - 1; a >= 0; a--) {
                        double sacc = -(rec[(DELTA ? R::LHD : R::LH) + a] + L.rd[a]);
#pragma unroll
                        for (int b = a + 1; b < NV; b++) sacc -= rec[R::LUU + b * (b + 1) / 2 + a] * L.du[b];
                        L.du[a] = sacc * rec[R::LUU + a * (a + 1) / 2 + a];
                    }
                }
                {
                    const bool act = L.is_u ? hasU : hasX;
                    const int c = L.cidx;
                    double dzw = L.is_u ? pick(L.du, L.ua) : L.dxs;
                    if (DELTA) dzw += rec2[R::DZA + L.zoff];
                    const double sg = L.sg;
                    const double lam = rec2[R::LAM + c], t = rec2[R::T + c], zb = rec2[R::Z + L.zoff];
                    const double rd = sg * (rec[L.dboff] - zb) + t;
                    double rm = lam * t - o.tau_min;
                    if (DELTA) rm += mcw * rec2[R::MC + c] - L.sigmu;
                    const double dt = sg * dzw - rd;
                    const double dl = -(lam * dt + rm) / t;
                    if (!DELTA && act && L.run) L.grec[R::MC + c] = dt * dl;
                    if (act && L.aN * dl < lam * L.aD) { L.aN = lam; L.aD = dl; }
                    if (act && L.aN * dt < t * L.aD) { L.aN = t; L.aD = dt; }
                    L.nr[1] += act ? lam * t : 0.0;
                    L.nr[2] += act ? lam * dt + t * dl : 0.0;
                    L.nr[3] += act ? dl * dt : 0.0;
                }
                if (L.run && L.isx) {
                    if (!DELTA) L.grec[R::DZA + NU + L.r] = L.dx;
                    else L.grec[R::DZ + NU + L.r] = L.dx + rec2[R::DZA + NU + L.r];
                }
                if (L.run && L.r < NV) {
                    const double duo = pick(L.du, L.r);
                    if (!DELTA) L.grec[R::DZA + L.r] = duo;
                    else L.grec[R::DZ + L.r] = duo + rec2[R::DZA + L.r];
                }
                if (hasU) {
                    double ps[3];
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        double sacc = L.rd[NV + i];
#pragma unroll
                        for (int a = 0; a < NV; a++) sacc += rec[R::E + i * NC + 1 + 2 * NV + a] * L.du[a];
                        ps[i] = sacc;
                    }
                    const double psel = L.r == 0 ? ps[0] : L.r == 1 ? ps[1] : L.r == 2 ? ps[2] : 0.0;
                    double xn = psel + ltk[L.ksf] * L.dx + ltk[L.kpf] * L.dxp + ltk[L.kuf] * pick(L.du, L.cj);
                    if (!DELTA) xn += rec[R::RB + (L.isx ? L.r : 0)];
                    L.dx = L.isx ? xn : 0.0;
                }
            CO_PHASE_END_NS
        }
        // ratio test (max of the negated step) and the three mu sums over the group
        GRP_PHASE_BEGIN(lanes)
            L.nr[0] = L.aN / L.aD;
        CO_PHASE_END_NS
        GRP_PHASE_BEGIN(lanes)
            { const double ov = CO_SHFLX(nr[0], 1); L.nq[0] = L.nr[0] > ov ? L.nr[0] : ov; }
#pragma unroll
            for (int q = 1; q < 4; q++) L.nq[q] = L.nr[q] + CO_SHFLX(nr[q], 1);
        CO_PHASE_END_NS
        GRP_PHASE_BEGIN(lanes)
            { const double ov = CO_SHFLX(nq[0], 2); L.nr[0] = L.nq[0] > ov ? L.nq[0] : ov; }
#pragma unroll
            for (int q = 1; q < 4; q++) L.nr[q] = L.nq[q] + CO_SHFLX(nq[q], 2);
        CO_PHASE_END_NS
        GRP_PHASE_BEGIN(lanes)
            { const double ov = CO_SHFLX(nr[0], 4); L.nq[0] = L.nr[0] > ov ? L.nr[0] : ov; }
#pragma unroll
            for (int q = 1; q < 4; q++) L.nq[q] = L.nr[q] + CO_SHFLX(nr[q], 4);
        CO_PHASE_END_NS
        if (G == 16) {
            GRP_PHASE_BEGIN(lanes)
                { const double ov = CO_SHFLX(nq[0], 8); L.nr[0] = L.nq[0] > ov ? L.nq[0] : ov; }
#pragma unroll
                for (int q = 1; q < 4; q++) L.nr[q] = L.nq[q] + CO_SHFLX(nq[q], 8);
            CO_PHASE_END_NS
        } else {
            GRP_PHASE_BEGIN(lanes)
#pragma unroll
                for (int q = 0; q < 4; q++) L.nr[q] = L.nq[q];
            CO_PHASE_END_NS
        }
    }

    // =========================================================================================
    // delta backward sweep (Rti::stage_Bd): right-hand side only in the complementarity rows
    // =========================================================================================
    NMPC_HD static void sweep_Bd(Lane* lanes, double* sm, double* ws, const Tables& tb, double mcw)
    {
        constexpr int ISZ = Img<SW_BD>::SIZE, D = Img<SW_BD>::D;
        GRP_PHASE_BEGIN(lanes)
            L.dp = 0.0;
            if (L.run) {
                double* scr = sm + L.so;
                if (L.r < NXP) { scr[O_V1 + L.r] = 0.0; scr[O_V1 + NXP + L.r] = 0.0; }
            }
            begin_sweep<SW_BD, -1>(L, sm, ws, tb, NSTAGE);
        GRP_PHASE_END
#pragma unroll 1
        for (int s = 0; s <= NSTAGE; s++) {
            const int k = NSTAGE - s, slot = s % D, pslot = (s + D - 1) % D, par = s & 1;
            const bool hasU = k < NSTAGE, hasX = k > 0, pvalid = s + D - 1 <= NSTAGE;
            GRP_PHASE_BEGIN(lanes)
                begin_stage<SW_BD, -1>(L);
            GRP_PHASE_END
            // ---- D1: prefetch; the complementarity term of the lane's constraint ------------------------------------
            GRP_PHASE_BEGIN(lanes)
                prefetch<SW_BD, -1>(L, sm, D - 1, pslot, pvalid);
                const double* rec2 = Img<SW_BD>::b(sm + L.so + O_IN + slot * ISZ);
                L.gt = 0.0;
                if (L.is_u ? hasU : hasX) L.gt = L.sg * ((mcw * rec2[R::MC + L.cidx] - L.sigmu) / rec2[R::T + L.cidx]);
            CO_PHASE_END_NS
            GRP_PHASE_BEGIN(lanes)
                L.gs2 = L.gt + CO_SHFLX(gt, XM);
            CO_PHASE_END_NS
            // ---- D2: q = [A B]' dp + complementarity terms, lh = L^-1 q_u (every lane), dp = q_x - K' lh ---------------
            GRP_PHASE_BEGIN(lanes)
                double* scr = sm + L.so;
                const double* img = scr + O_IN + slot * ISZ;
                const double* rec = Img<SW_BD>::a(img);
                const double* ltk = sm + L.to + slot * TROW;
                const double* dpv = scr + O_V1 + par * NXP;
                double* dpn = scr + O_V1 + (par ^ 1) * NXP;
                double d[NXP];
                ldv(d, dpv);
                double gua[NV];
#pragma unroll
                for (int a = 0; a < NV; a++) gua[a] = CO_SHFL(gs2, L.ul[a]);
                double qx = 0.0;
                if (hasU) {
                    double e0, e1, e2;
                    own_ecol(L, rec + R::E, e0, e1, e2);
                    qx = e0 * d[0] + e1 * d[1] + e2 * d[2] + ltk[L.kst] * L.dp + ltk[L.kpt] * dpv[L.ptx];
                }
                if (L.bnd && hasX) qx += L.gs2;
                if (hasU) {
#pragma unroll
                    for (int a = 0; a < NV; a++) {
                        const int q = 1 + 2 * NV + a;
                        double sacc = rec[R::E + q] * d[0] + rec[R::E + NC + q] * d[1] + rec[R::E + 2 * NC + q] * d[2]
                                    + ltk[2 * NV + a] * d[3 + a] + ltk[3 * NV + a] * d[3 + NV + a] + gua[a];
#pragma unroll
                        for (int c = 0; c < a; c++) sacc -= rec[R::LUU + a * (a + 1) / 2 + c] * L.lh[c];
                        L.lh[a] = sacc * rec[R::LUU + a * (a + 1) / 2 + a];
                    }
                    if (L.run && L.r == NX) {
#pragma unroll
                        for (int a = 0; a < NV; a++) L.grec[R::LHD + a] = L.lh[a];
                    }
                    if (hasX && L.isx) {
#pragma unroll
                        for (int a = 0; a < NV; a++) qx -= rec[R::KH + a * NX + L.r] * L.lh[a];
                    }
                }
                if (L.isx) { L.dp = qx; dpn[L.r] = qx; }
            GRP_PHASE_END
        }
    }

    // =========================================================================================
    // the interior-point loop of one warp: NSLOT instances in lockstep, slots refilled from the
    // work queue `next` (instances [0, n) of the chunk) at every iteration boundary
    // =========================================================================================
    NMPC_HD static void run_warp(Lane* lanes, double* sm, double* ws, int i0, int n, int* next, const Tables& tb,
                                 const double* We_inst, int ldWe, const IpmOpts& o, const GrpOut& out, const GrpResume& rs)
    {
        if (rs.n_dev) n = *rs.n_dev;
#define CTL(L) (*reinterpret_cast<LaneCtl*>(sm + (L).so + O_CTL))
#pragma unroll 1
        for (;;) {
            GRP_PHASE_BEGIN(lanes)
                if (!L.act && L.r == 0) {
                    int* q = reinterpret_cast<int*>(sm + L.so + O_AST + 1);
                    *q = grp_fetch_add(next);
                }
            GRP_PHASE_END
            GRP_PHASE_BEGIN(lanes)
                if (!L.act) {
                    const int idx = *reinterpret_cast<const int*>(sm + L.so + O_AST + 1);
                    if (idx < n) {
                        L.act = true;
                        L.first = rs.ctl == nullptr; L.skipB = !L.first;
                        L.gi = rs.list ? rs.list[idx] : idx;
                        L.li = rs.ctl ? idx : L.gi;               // resumed instances live in compacted records, fresh ones in their own
                        const double* wp = We_inst ? We_inst + i0 + L.gi : tb.We;
                        const size_t wl = We_inst ? (size_t)ldWe : 1;
                        L.We_j = L.isx ? wp[(size_t)L.r * wl] : 0.0;
                        if (L.r == 0) { if (rs.ctl) CTL(L) = reinterpret_cast<const LaneCtl*>(rs.ctl)[idx]; else CTL(L).init(true); }
                    }
                } else if (L.r == 0) sm[L.so + O_AST] = S::before_B(CTL(L));
            GRP_PHASE_END
            GRP_PHASE_BEGIN(lanes)
                L.run = L.act && !L.skipB;
                L.astep = 0.0;
                if (L.act) {
                    if (!L.first && !L.skipB) L.astep = sm[L.so + O_AST];
                    L.sigmu = CTL(L).sigmu; L.mcw = CTL(L).mcw;
                }
            GRP_PHASE_END
            if (!warp_any(lanes, [](const Lane& L) { return L.act; })) break;

            if (warp_any(lanes, [](const Lane& L) { return L.run; })) sweep_B(lanes, sm, ws, tb, o);
            GRP_PHASE_BEGIN(lanes)
                if (!L.run || L.r != 0) continue;
                LaneCtl& c = CTL(L);
                typename S::CarryB cy;
                cy.ng = L.nr[0]; cy.nb = L.nr[1]; cy.nd = L.nr[2]; cy.nm = L.nr[3]; cy.lru = L.nr[4]; cy.musum = L.nr[5];
                S::after_B(c, cy, o, L.first);
                if (c.done) {
                    const int i = i0 + L.gi;
                    out.qp_status[i] = c.status;
                    out.qp_iter[i] = c.iter;
                    if (out.stats) {
                        out.stats[(size_t)0 * out.B + i] = c.nrm[0]; out.stats[(size_t)1 * out.B + i] = c.nrm[1];
                        out.stats[(size_t)2 * out.B + i] = c.nrm[2]; out.stats[(size_t)3 * out.B + i] = c.nrm[3];
                        out.stats[(size_t)4 * out.B + i] = c.mu;
                        out.stats[(size_t)5 * out.B + i] = c.lin_res;
                        out.stats[(size_t)6 * out.B + i] = (double)c.nfb;
                        out.stats[(size_t)7 * out.B + i] = (double)c.status;
                    }
                }
            GRP_PHASE_END
            GRP_PHASE_BEGIN(lanes)
                if (L.run) { L.first = false; if (CTL(L).done) L.act = false; }
                L.skipB = false;
                L.run = L.act;
            GRP_PHASE_END
            if (!warp_any(lanes, [](const Lane& L) { return L.act; })) continue;

            sweep_F<false>(lanes, sm, ws, tb, o, 1.0);
            GRP_PHASE_BEGIN(lanes)
                if (!L.act || L.r != 0) continue;
                typename S::CarryF cy;
                cy.alpha = L.nr[0]; cy.S0 = L.nr[1]; cy.S1 = L.nr[2]; cy.S2 = L.nr[3];
                S::after_F(CTL(L), cy, o);
            GRP_PHASE_END
            GRP_PHASE_BEGIN(lanes)
                if (L.act) L.sigmu = CTL(L).sigmu;
            GRP_PHASE_END
            sweep_Bd(lanes, sm, ws, tb, 1.0);
            sweep_F<true>(lanes, sm, ws, tb, o, 1.0);
            GRP_PHASE_BEGIN(lanes)
                if (!L.act || L.r != 0) continue;
                typename S::CarryF cy;
                cy.alpha = L.nr[0]; cy.S0 = L.nr[1]; cy.S1 = L.nr[2]; cy.S2 = L.nr[3];
                S::after_Fd(CTL(L), cy, o);
            GRP_PHASE_END
            GRP_PHASE_BEGIN(lanes)
                L.run = L.act && CTL(L).fb != 0;
            GRP_PHASE_END
            if (warp_any(lanes, [](const Lane& L) { return L.run; })) {
                sweep_Bd(lanes, sm, ws, tb, 0.0);
                sweep_F<true>(lanes, sm, ws, tb, o, 0.0);
                GRP_PHASE_BEGIN(lanes)
                    if (!L.run || L.r != 0) continue;
                    typename S::CarryF cy;
                    cy.alpha = L.nr[0]; cy.S0 = L.nr[1]; cy.S1 = L.nr[2]; cy.S2 = L.nr[3];
                    S::after_Fd_fallback(CTL(L), cy);
                GRP_PHASE_END
            }
        }
#undef CTL
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
};

}  // namespace nmpc
