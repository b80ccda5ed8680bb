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
#ifndef NMPC_GRP_DEPTH
#define NMPC_GRP_DEPTH 2
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
    static constexpr int ER = 3;                   // all three pose rows of [A|B] are stored
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
template <int PENDING>
NMPC_HD void grp_cp_wait()
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group %0;\n" :: "n"(PENDING) : "memory");
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

// pins a lane-constant integer in a register: without it ptxas rematerialises role indices from
// threadIdx (S2R + shifts + masks) at every use, ~20% of all issued instructions
NMPC_HD int grp_pin(int v)
{
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+r"(v));
#endif
    return v;
}

// 1/sqrt(d): the device intrinsic (MUFU.RSQ64H + Newton steps, <= 1 ulp) instead of sqrt + divide
NMPC_HD double grp_rsqrt(double d)
{
#if defined(__CUDA_ARCH__)
    return rsqrt(d);
#else
    return 1.0 / sqrt(d);
#endif
}

// running max of |v| (inf-norms): compare + select instead of fmax's NaN-aware sequence; a NaN is ignored here
// and caught by the mu != mu test of after_B
NMPC_HD double grp_maxabs(double acc, double v) { v = fabs(v); return v > acc ? v : acc; }
// pins a pointer in registers (see grp_pin): the running record pointer is advanced, not recomputed
template <class T>
NMPC_HD T* grp_pin_ptr(T* p)
{
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+l"(p));
#endif
    return p;
}

struct GrpOut {            // per-instance results of K3 (global, indexed by instance of the batch)
    int* qp_status;
    int* qp_iter;
    double* stats;         // [8][B] or null
    int B;                 // leading dimension of stats
};
// hand-over from the per-sweep kernels (hybrid schedule): the queue holds *n_dev records, record q belongs to
// instance list[q] of the chunk, is in the middle of an iteration (factorised, predictor not yet run) and
// continues with the control block ctl[q] (an array of Rti<M>::LaneCtl).  All null = every instance from the cold start.
// list without ctl (SQP passes): the queue holds *n_dev instances list[q], each from the cold start in its own record.
struct GrpResume {
    const int* n_dev;
    const int* list;
    const void* ctl;
};

template <class M, int G_>
struct Grp {
    using S = Rti<M>;
    using LaneCtl = typename S::LaneCtl;
    static constexpr int G = G_, NSLOT = 32 / G;
    static constexpr int NV = S::NV, NX = S::NX, NU = S::NU, NZ = S::NZ, NY = S::NY, NC = S::NC, NB2 = S::NB2, NLU = S::NLU;
    static constexpr int NCT = 2 * NB2;                  // one-sided constraints per stage
    static constexpr int NK = NSTAGE + 1;
    static constexpr int LTE = 4 * NV + 2, LT_ZERO = 4 * NV, LT_ONE = 4 * NV + 1;
    // lane roles (all table driven, see init_lane):
    //   constraint role : lane c < NCT owns one-sided constraint c (lower NB2, then upper NB2)
    //   component role  : lane q < NC owns the z component whose column of the pose rows E is q
    //                     (theta | actual | ref | u); lanes NC.. own the pose components x, y (unit columns)
    //   column/row role : lane j < NX owns state column j of the cost-to-go, of M and K, and row j of
    //                     the dynamics;  lanes UL0.. own the control columns
    static constexpr int XL = NC;                         // lane of the pose component x ...
    static constexpr int YL = (NC + 1 < G) ? NC + 1 : NC - 1;   // ... and of y (on top of that lane's generic component if G = NC + 1)
    static constexpr int ULN = (G - NX >= NV) ? NV : 1;   // lanes that own control columns (NV / ULN each) ...
    static constexpr int ULB = G > NX ? NX : 0;           // ... starting at this lane
    static constexpr int WL = G > NX ? NX : 0;            // lane that writes the per-stage scalars of the factorisation
    static_assert(G >= NX && G >= NCT && G > NC && 32 % G == 0, "one state column, one constraint, one component per lane");
    using R = GRec<NV>;

    enum { SW_B = 0, SW_F = 1, SW_BD = 2, SW_FD = 3 };
    NMPC_HD static constexpr int imax(int a, int b) { return a > b ? a : b; }
    // the two record ranges [A0, A1), [B0, B1) a sweep reads (GRec), its compact image [A | B] and prefetch depth
    template <int KIND> struct Img {
        static constexpr int A0 = KIND == SW_B ? R::Q : KIND == SW_BD ? R::E : R::DLB;
        static constexpr int A1 = KIND == SW_B ? R::LHD : KIND == SW_F ? R::DZA : R::LH;
        static constexpr int B0 = KIND == SW_B ? R::MC : KIND == SW_F ? R::T : KIND == SW_BD ? R::MC : R::DZA;
        static constexpr int B1 = KIND == SW_B ? R::NREC : KIND == SW_BD ? R::LAM : R::PI;
        static constexpr int SIZE = (A1 - A0) + (B1 - B0);
        static constexpr int D = NMPC_GRP_DEPTH;         // measured: 3-4 stages in flight change nothing (a lone warp is bound by its
                                                         // ~2,500 dependent instructions per stage set, not by DRAM latency), 2 costs the least shared memory
        // virtual record bases: a(img)[R::X] for a field X of range A, b(img)[R::X] for a field of range B
        NMPC_HD static double* a(double* img) { return img - A0; }
        NMPC_HD static double* b(double* img) { return img + (A1 - A0) - B0; }
        NMPC_HD static const double* a(const double* img) { return img - A0; }
        NMPC_HD static const double* b(const double* img) { return img + (A1 - A0) - B0; }
    };

    // ---- shared-memory scratch of one slot (doubles) ----------------------------------------
    // ring of compact stage images: a sweep of kind K keeps Img<K>::D stages in flight, each image holding only the
    // two record ranges the sweep reads (Img<K>::SIZE doubles).  Depth matters for a warp that runs alone (the
    // stragglers of a batch, batch 1): its stage takes 0.3-1 us of compute against > 1 us of DRAM latency.
    static constexpr int O_IN = 0;
    static constexpr int RING = imax(imax(imax(Img<SW_B>::D * Img<SW_B>::SIZE, Img<SW_F>::D * Img<SW_F>::SIZE),
                                          imax(Img<SW_BD>::D * Img<SW_BD>::SIZE, Img<SW_FD>::D * Img<SW_FD>::SIZE)),
                                     imax(G * 8 /* the end-of-sweep reductions borrow the ring */,
                                          3 * Img<SW_B>::SIZE /* the B sweep keeps the stages of its two halves and the prefetch */));
    static constexpr int DMAX = imax(3, imax(imax(Img<SW_B>::D, Img<SW_F>::D), imax(Img<SW_BD>::D, Img<SW_FD>::D)));
    static constexpr int O_CAR = O_IN + RING;             // [2][3][NX] carries of the B sweep: pio, dpi, xn
                                                          //   (F: dx [2][NX]; Bd: dp [2][NX])
    static constexpr int O_PV = O_CAR + 6 * NX;           // NX   gradient of the cost-to-go
    // hand-off of the B sweep's update half (stage k) to its Riccati half (one super-step later): two buffers each
    static constexpr int O_RB = O_PV + NX;                // [2][NX] dynamics residual
    static constexpr int O_GX = O_RB + 2 * NX;            // [2][NX] stage gradient, state rows (Bd: q_x)
    static constexpr int O_DGX = O_GX + 2 * NX;           // [2][NX] diagonal H + reg + Gamma, state rows
    static constexpr int O_GU = O_DGX + 2 * NX;           // [2][NV] control gradient (F: s_a, Bd: q_u)
    static constexpr int O_DGU = O_GU + 2 * NV;           // [2][NV]
    static constexpr int O_GG = O_DGU + 2 * NV;           // NV   control gradient with the cost-to-go terms (B4 -> B5)
    static constexpr int O_CB = O_GG + NV;                // [NCT+1][4] per-constraint terms (last entry: zeros)
    static constexpr int O_PBA = O_CB + 4 * (NCT + 1);    // [NX][NZ]  P * [A B], columns [x; u]
    static constexpr int O_MUU = O_PBA + NX * NZ;         // [NV][NV]
    static constexpr int O_KB = O_MUU + NV * NV;          // [NV][NX]
    static constexpr int O_DUMP = O_KB + NV * NX;         // 4    target of role-masked stores
    static constexpr int O_AST = O_DUMP + 2;              //      damped step handed from lane 0 to the slot
    static constexpr int O_CTL = O_DUMP + 4;              // the slot's LaneCtl (owned by lane 0 between the sweeps)
    static constexpr int CTL_D = (int)((sizeof(LaneCtl) + 7) / 8);
    static constexpr int O_RED = O_IN;                    // [G][8] reductions at the end of a sweep / queue hand-out (the record image is dead there)
    static constexpr int O_END = O_CTL + CTL_D;
    static_assert(G * 8 <= RING, "reduction buffer must fit into the image ring");
    // slot stride: even (16-byte copies) and = 8 mod 16 so that equal offsets of neighbouring slots
    // fall into different bank groups
    static constexpr int SLOT_D = ((O_END + 15) / 16) * 16 + 8;
    // per-warp image of the stage table row (double buffered like the records): [lte (4NV+2) | W (NY)], rows padded
    static constexpr int T_W = (LTE + 1) & ~1;            // offset of the weights in a row
    static constexpr int TROW = T_W + ((NY + 1) & ~1);
    static constexpr int O_TAB = SLOT_D * NSLOT;          // [DMAX][TROW] ring of stage-table rows
    static constexpr int WARP_D = SLOT_D * NSLOT + DMAX * TROW;   // doubles of shared memory per warp
    static_assert(TROW / 2 <= 32, "one 16-byte chunk of the stage table per lane");

    // static-index products with J = [A B] (registers only): (column w of J) . v, w in [x; u] order
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
    // 1.0 / 0.0 indicator: selecting a register-array element by a lane-dependent index with
    // arithmetic keeps the array in registers (a select chain is turned into a local-memory index)
    NMPC_HD static double sel(bool b) { return b ? 1.0 : 0.0; }

    // ---- state of one lane (registers on the device) ----------------------------------------
    struct Lane {
        int r, so;                // role index in the group; offset (doubles) of the slot's scratch in the shared array
        int li;                   // instance (index into the chunk) of the slot, -1 = none
        bool act, first, run;     // slot has an instance / its next B sweep is the first / takes part in the current sweep
        bool skipB;               // resumed instance: joins the loop at the predictor, skipping the factorising sweep once
        int gi;                   // instance of the chunk the slot's record belongs to (= li unless resumed)
        double sigmu, mcw;        // copies of the slot's LaneCtl fields the sweeps use (LaneCtl itself lives in the scratch)
        double astep;             // damped step applied by the running B sweep
        int to, wl;               // offset of the warp's stage-table image; lane of the warp
        double We_c, We_xy;       // terminal weight of the lane's generic state component / of its x or y component
        double* grec;             // record of the stage being processed (advanced stage by stage)
        const double* tsrc;       // this lane's 16-byte chunk of the stage-table row being prefetched
        // component role (generic components q = r < NC)
        int cq_z, cq_x, cq_y;     // z-order index; state index (-1: control); index into W / y
        int cq_bl, cq_bu;         // lower / upper constraint (NCT: none)
        int cq_i1, cq_i2, cq_k1, cq_k2;   // LTI part of the column of J: k1 * v[i1] + k2 * v[i2], k from the lte table
        // constraint role (c = r < NCT)
        int ct_z, ct_x, ct_u;     // z-order index of the bounded component; its state index or -1; its control index or -1
        // column role (state j = r < NX): column j of J = pose part (E column cs_e, -1: unit vector e_j) + LTI part
        int cs_e, cs_i1, cs_i2, cs_k1, cs_k2;
        double ct_s;              // +1 lower, -1 upper
        // B sweep
        double Pc[NX];            // column r of the cost-to-go of the successor stage
        double Ef[3 * NC];        // pose rows of [A|B] of the stage
        double lt[4 * NV];
        double Mx[NZ];            // column r of M = J'PJ + D, rows [x; u]
        double Kc[NV];            // column r of K
        double lh[NV];
        double gx;                // stage gradient of state r
        double ng, nb, nd, nm, musum, lru;
        // F / Bd sweeps
        double dxr;               // dx (F) / q_x (Bd) component r
        double du[NV];
        double alpha, S0, S1, S2;
        double aN, aD;            // ratio test: the binding candidate kept as a fraction aN / aD (aD < 0), divided once per sweep
    };

    // lane: lane of the warp; warp: index of the warp's scratch in the shared array
    NMPC_HD static void init_lane(Lane& L, int lane, int warp)
    {
        const int r = grp_pin(lane % G);
        L.r = r; L.so = grp_pin(warp * WARP_D + (lane / G) * SLOT_D); L.li = -1;
        L.to = grp_pin(warp * WARP_D + O_TAB); L.wl = grp_pin(lane);
        L.act = false; L.first = false; L.run = false; L.skipB = false; L.gi = -1;
        L.sigmu = 0.0; L.mcw = 1.0;
        L.astep = 0.0; L.We_c = 0.0; L.We_xy = 0.0; L.grec = nullptr; L.tsrc = nullptr;
        // component role: E column q = r -> theta | actual c | ref c | u a
        L.cq_z = 0; L.cq_x = -1; L.cq_y = 0; L.cq_bl = NCT; L.cq_bu = NCT;
        L.cq_i1 = 0; L.cq_i2 = 0; L.cq_k1 = LT_ZERO; L.cq_k2 = LT_ZERO;
        if (r == 0) { L.cq_x = 2; }
        else if (r <= NV) { const int c = r - 1; L.cq_x = 3 + c; L.cq_i1 = 3 + c; L.cq_k1 = c; }
        else if (r <= 2 * NV) {
            const int c = r - 1 - NV;
            L.cq_x = 3 + NV + c; L.cq_i1 = 3 + c; L.cq_k1 = NV + c; L.cq_i2 = 3 + NV + c; L.cq_k2 = LT_ONE;
            L.cq_bl = NV + c; L.cq_bu = NB2 + NV + c;
        } else if (r < NC) {
            const int a = r - 1 - 2 * NV;
            L.cq_i1 = 3 + a; L.cq_k1 = 2 * NV + a; L.cq_i2 = 3 + NV + a; L.cq_k2 = 3 * NV + a;
            L.cq_bl = a; L.cq_bu = NB2 + a;
            L.cq_z = a; L.cq_y = NX + a;
        }
        if (L.cq_x >= 0) { L.cq_z = NU + L.cq_x; L.cq_y = L.cq_x; }
        // constraint role
        const int b = r % NB2;
        L.ct_s = (r < NB2) ? 1.0 : -1.0;
        L.ct_u = b < NV ? b : -1;
        L.ct_x = b < NV ? -1 : 3 + b;
        L.ct_z = b < NV ? b : NU + 3 + b;
        L.cs_e = -1; L.cs_i1 = 0; L.cs_i2 = 0; L.cs_k1 = LT_ZERO; L.cs_k2 = LT_ZERO;
        if (r == 2) L.cs_e = 0;
        else if (r >= 3 && r < 3 + NV) { const int c = r - 3; L.cs_e = 1 + c; L.cs_i1 = 3 + c; L.cs_k1 = c; }
        else if (r >= 3 + NV && r < NX) { const int c = r - 3 - NV; L.cs_e = 1 + NV + c; L.cs_i1 = 3 + c; L.cs_k1 = NV + c; L.cs_i2 = 3 + NV + c; L.cs_k2 = LT_ONE; }
        L.cs_e = grp_pin(L.cs_e); L.cs_i1 = grp_pin(L.cs_i1); L.cs_i2 = grp_pin(L.cs_i2); L.cs_k1 = grp_pin(L.cs_k1); L.cs_k2 = grp_pin(L.cs_k2);
        L.cq_z = grp_pin(L.cq_z); L.cq_x = grp_pin(L.cq_x); L.cq_y = grp_pin(L.cq_y); L.cq_bl = grp_pin(L.cq_bl); L.cq_bu = grp_pin(L.cq_bu);
        L.cq_i1 = grp_pin(L.cq_i1); L.cq_i2 = grp_pin(L.cq_i2); L.cq_k1 = grp_pin(L.cq_k1); L.cq_k2 = grp_pin(L.cq_k2);
        L.ct_z = grp_pin(L.ct_z); L.ct_x = grp_pin(L.ct_x); L.ct_u = grp_pin(L.ct_u);
    }

    // hand-over from the per-sweep path: stage k of one instance from its tile (Rec<NV>, lane-resolved pointer,
    // field stride LANES) into a group record.  State right after a factorising sweep: QP data, iterate,
    // factorisation (the steps DZ / DZA / MC / LHD are rewritten before they are read again).
    NMPC_HD static void tile_to_record(const double* tl, int k, double* rec, const double* thr_k)
    {
        tile_to_record_range<0, R::NREC>(tl, k, rec, thr_k);
    }
    // entries [D0, D1) of the record into out[0, D1 - D0): the hand-over kernel converts a record in pieces so that its
    // shared-memory staging stays small (every loop is unrolled, the range tests fold at compile time)
    template <int D0, int D1>
    NMPC_HD static void tile_to_record_range(const double* tl, int k, double* out, const double* thr_k)
    {
        using T = typename S::R;
        const double* lin = tl + T::OFF_LIN + (size_t)k * T::NF_LIN * LANES;
        const double* it = tl + T::OFF_IT + (size_t)k * T::NF_IT * LANES;
        const double* fa = tl + T::OFF_FA + (size_t)k * T::NF_FA * LANES;
        const bool dyn = k < NSTAGE;
#define NMPC_PUT(off, val) do { if ((off) >= D0 && (off) < D1) out[(off) - D0] = (val); } while (0)
#pragma unroll
        for (int i = 0; i < NZ; i++) NMPC_PUT(R::Q + i, lin[(T::Q + i) * LANES]);
#pragma unroll
        for (int i = 0; i < NX; i++) NMPC_PUT(R::B0 + i, dyn ? lin[(T::B0 + i) * LANES] : 0.0);
#pragma unroll
        for (int i = 0; i < NB2; i++) { NMPC_PUT(R::DLB + i, lin[(T::DLB + i) * LANES]); NMPC_PUT(R::DUB + i, lin[(T::DUB + i) * LANES]); }
        // the terminal stage has no dynamics: its E is never written in the tile, and the group kernel multiplies it by
        // zero carries, so it must be finite
#pragma unroll
        for (int i = 0; i < T::ER * NC; i++) NMPC_PUT(R::E + i, dyn ? lin[(T::E + i) * LANES] : 0.0);
#pragma unroll
        for (int i = T::ER * NC; i < 3 * NC; i++) NMPC_PUT(R::E + i, dyn ? thr_k[i - 2 * NC] : 0.0);   // theta row from the stage table
#pragma unroll
        for (int i = 0; i < NLU; i++) NMPC_PUT(R::LUU + i, fa[(T::LUU + i) * LANES]);
#pragma unroll
        for (int i = 0; i < NV * NX; i++) NMPC_PUT(R::KH + i, fa[(T::KH + i) * LANES]);
#pragma unroll
        for (int i = 0; i < NV; i++) NMPC_PUT(R::LH + i, fa[(T::LH + i) * LANES]);
#pragma unroll
        for (int i = 0; i < NX; i++) NMPC_PUT(R::RB + i, fa[(T::RB + i) * LANES]);
#pragma unroll
        for (int i = 0; i < 2 * NB2; i++) { NMPC_PUT(R::T + i, it[(T::T + i) * LANES]); NMPC_PUT(R::LAM + i, it[(T::LAM + i) * LANES]); }
#pragma unroll
        for (int i = 0; i < NZ; i++) NMPC_PUT(R::Z + i, it[(T::Z + i) * LANES]);
#pragma unroll
        for (int i = 0; i < NX; i++) NMPC_PUT(R::PI + i, it[(T::PI + i) * LANES]);
        // padding and the step fields stay defined (the steps are rewritten before they are read)
        if ((3 * NC) & 1) NMPC_PUT(R::E + 3 * NC, 0.0);
        if (NLU & 1) NMPC_PUT(R::LUU + NLU, 0.0);
        if (NX & 1) { NMPC_PUT(R::RB + NX, 0.0); NMPC_PUT(R::PI + NX, 0.0); }
        if (NZ & 1) { NMPC_PUT(R::Z + NZ, 0.0); NMPC_PUT(R::DZ + NZ, 0.0); NMPC_PUT(R::DZA + NZ, 0.0); }
#pragma unroll
        for (int i = 0; i < NV; i++) NMPC_PUT(R::LHD + i, 0.0);
#pragma unroll
        for (int i = 0; i < NZ; i++) { NMPC_PUT(R::DZ + i, 0.0); NMPC_PUT(R::DZA + i, 0.0); }
#pragma unroll
        for (int i = 0; i < 2 * NB2; i++) NMPC_PUT(R::MC + i, 0.0);
#undef NMPC_PUT
    }

    NMPC_HD static double* rec_of(double* ws, int li, int k) { return ws + (size_t)li * R::inst_doubles + (size_t)k * R::NREC; }

    // copy doubles [d0, d1) of the record (16-byte chunks spread over the group)
    NMPC_HD static void copy_range(double* dst, const double* src, int d0, int d1, int r)
    {
#pragma unroll
        for (int c = d0; c < d1; c += 2 * G)
            if (c + 2 * r < d1) grp_cp16(dst + c + 2 * r, src + c + 2 * r);
    }


    // the stage-table row of stage k into image `buf` of the warp (every lane of the warp, running slot or not)
    NMPC_HD static void issue_tab(const Lane& L, double* sm, const double* src, int slot)
    {
        if (2 * L.wl < TROW) grp_cp16(sm + L.to + slot * TROW + 2 * L.wl, src);
    }
    // the two ranges of one record into an image
    template <int KIND>
    NMPC_HD static void issue(const Lane& L, const double* src, double* img)
    {
        copy_range(Img<KIND>::a(img), src, Img<KIND>::A0, Img<KIND>::A1, L.r);
        copy_range(Img<KIND>::b(img), src, Img<KIND>::B0, Img<KIND>::B1, L.r);
    }
    // prefetch of the stage `ahead` stages after the current one (pointers L.grec / L.tsrc) into ring slot `slot`;
    // one cp.async group per stage, committed by every lane whether it copied or not
    template <int KIND, int DIR>
    NMPC_HD static void prefetch(const Lane& L, double* sm, int ahead, int slot, bool valid)
    {
        if (valid) {
            issue_tab(L, sm, L.tsrc + DIR * ahead * TROW, slot);
            if (L.run) issue<KIND>(L, L.grec + DIR * ahead * R::NREC, sm + L.so + O_IN + slot * Img<KIND>::SIZE);
        }
        grp_cp_commit();
    }
    // start of a sweep at stage k0: record / stage-table pointers one stage before it, first D-1 stages in flight
    template <int KIND, int DIR>
    NMPC_HD static void begin_sweep(Lane& L, double* sm, double* ws, const Tables& tb, int k0)
    {
        L.tsrc = grp_pin_ptr(tb.stg + (size_t)k0 * TROW + 2 * L.wl - DIR * TROW);
        if (L.run) L.grec = grp_pin_ptr(rec_of(ws, L.li, k0) - DIR * R::NREC);
#pragma unroll
        for (int j = 0; j < Img<KIND>::D - 1; j++) prefetch<KIND, DIR>(L, sm, j + 1, j, true);
    }
    // top of a stage: advance to it, wait for its image (D-2 younger groups may still be in flight)
    template <int KIND, int DIR>
    NMPC_HD static void begin_stage(Lane& L)
    {
        L.grec += DIR * R::NREC; L.tsrc += DIR * TROW;
        grp_cp_wait<Img<KIND>::D - 2>();
    }

    // all-gather of up to 8 per-lane partials through the slot scratch: after the call every lane
    // of the slot has reduced q[0..nmax) with max and q[nmax..n) with +
    template <class GF, class PF>
    NMPC_HD static void reduce(Lane* lanes, double* sm, int n, int nmax, GF get, PF put)
    {
        GRP_PHASE_BEGIN(lanes)
            double* scr = sm + L.so;
            for (int q = 0; q < n; q++) scr[O_RED + L.r * 8 + q] = get(L, q);
        GRP_PHASE_END
        GRP_PHASE_BEGIN(lanes)
            const double* scr = sm + L.so;
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

    // J' v for the lane's generic component (E column q = r): v in shared memory, indexed by state
    NMPC_HD static double jt_comp(const Lane& L, double e0, double e1, double e2, double k1, double k2, const double* v)
    {
        return e0 * v[0] + e1 * v[1] + e2 * v[2] + k1 * v[L.cq_i1] + k2 * v[L.cq_i2];
    }
    // (row i of J) . [zu; zx] with the vectors given as functors (shared memory or registers)
    template <class FU, class FX>
    NMPC_HD static double jrow(int i, FU zu, FX zx, const double* E, const double* lt)
    {
        if (i < 3) {
            double s0 = (i < 2 ? zx(i) : 0.0) + E[i * NC] * zx(2), s1 = 0.0;
#pragma unroll
            for (int c = 0; c < NV; c++) {
                s0 += E[i * NC + 1 + c] * zx(3 + c) + E[i * NC + 1 + 2 * NV + c] * zu(c);
                s1 += E[i * NC + 1 + NV + c] * zx(3 + NV + c);
            }
            return s0 + s1;
        }
        const int c = i < 3 + NV ? i - 3 : i - 3 - NV;
        if (i < 3 + NV) return lt[c] * zx(3 + c) + lt[NV + c] * zx(3 + NV + c) + lt[2 * NV + c] * zu(c);
        return zx(3 + NV + c) + lt[3 * NV + c] * zu(c);
    }

    // =========================================================================================
    // B sweep: apply the previous step, residuals, Riccati factorisation, stages N..0
    // (the arithmetic of Rti::stage_B_update / stage_B_riccati; the cold start is an ordinary
    // sweep with step 0 over the records initialised by Rti::coldstart_stage)
    // =========================================================================================
    NMPC_HD static void sweep_B(Lane* lanes, double* sm, double* ws, const Tables& tb, const IpmOpts& o)
    {
        // The sweep is two recursions that only meet in the stage gradient: the UPDATE half of stage k (step, multipliers,
        // residuals: needs the update half of stage k+1) and the RICCATI half (needs the Riccati half of stage k+1 and the
        // update half of stage k).  They run skewed by one stage: super-step t does update(N - t) and riccati(N - t + 1) in
        // the same phases, which shortens the dependent chain of a stage by a third and saves one phase boundary.
        constexpr int ISZ = Img<SW_B>::SIZE;
        GRP_PHASE_BEGIN(lanes)
            L.ng = L.nb = L.nd = L.nm = L.musum = L.lru = 0.0;
            L.tsrc = grp_pin_ptr(tb.stg + (size_t)(NSTAGE + 1) * TROW + 2 * L.wl);      // one stage before the first (N)
            if (L.run) {
                double* scr = sm + L.so;
                if (L.r < NX) { scr[O_CAR + L.r] = 0.0; scr[O_CAR + NX + L.r] = 0.0; scr[O_CAR + 2 * NX + L.r] = 0.0; scr[O_PV + L.r] = 0.0; }
                if (L.r < 4) scr[O_CB + 4 * NCT + L.r] = 0.0;
                L.grec = grp_pin_ptr(rec_of(ws, L.li, NSTAGE) + R::NREC);
            }
            prefetch<SW_B, -1>(L, sm, 1, 0, true);
        GRP_PHASE_END
#pragma unroll 1
        for (int t = 0; t <= NSTAGE + 1; t++) {
            const int ku = NSTAGE - t, kr = ku + 1;
            const bool Uv = t <= NSTAGE, Rv = t >= 1;
            const bool hasU_u = ku < NSTAGE, hasX_u = ku > 0, hasU_r = kr < NSTAGE, hasX_r = kr > 0;
            const int slot_u = t % 3, slot_r = (t + 2) % 3, pslot = (t + 1) % 3, par_u = t & 1, par_r = par_u ^ 1;
            GRP_PHASE_BEGIN(lanes)
                if (Uv) { L.grec -= R::NREC; L.tsrc -= TROW; }
                grp_cp_wait<0>();
            GRP_PHASE_END
            // ---- P1: prefetch; update(ku): one constraint per lane;  riccati(kr): row r of P * [A B] ----------------
            GRP_PHASE_BEGIN(lanes)
                prefetch<SW_B, -1>(L, sm, 1, pslot, Uv && ku > 0);
                if (!L.run) continue;
                double* scr = sm + L.so;
                if (Uv && L.r < NCT) {
                    double* grec = L.grec;
                    const double* img = scr + O_IN + slot_u * ISZ;
                    const double* rec = Img<SW_B>::a(img); const double* rec2 = Img<SW_B>::b(img);
                    const double a_step = L.astep;
                    const int c = L.r;
                    const bool actc = (L.ct_u >= 0) ? hasU_u : hasX_u;
                    double rpart = 0.0, lpart = 0.0, gpart = 0.0, Gp = 0.0;
                    if (actc) {
                        const double sg = L.ct_s;
                        const double dbd = rec[R::DLB + c], z = rec2[R::Z + L.ct_z], dz = rec2[R::DZ + L.ct_z];
                        const double lam = rec2[R::LAM + c], tt = rec2[R::T + c], mc = rec2[R::MC + c];
                        const double rd = sg * (dbd - z) + tt;
                        const double rm = lam * tt - o.tau_min + L.mcw * mc - L.sigmu;
                        const double dt = sg * dz - rd;
                        const double dlam = -(lam * dt + rm) / tt;
                        const double lam_n = lam + a_step * dlam, t_n = tt + a_step * dt, zn = z + a_step * dz;
                        rpart = -sg * (lam + dlam);
                        const double rd_n = sg * (dbd - zn) + t_n;
                        const double pm = lam_n * t_n;
                        L.musum += pm;
                        const double rm_n = pm - o.tau_min;
                        L.nd = grp_maxabs(L.nd, rd_n);
                        L.nm = grp_maxabs(L.nm, rm_n);
                        const double ti = t_n < o.t_min ? 1.0 / o.t_min : 1.0 / t_n;
                        const double lc = lam_n < o.lam_min ? o.lam_min : lam_n;
                        Gp = ti * lc;
                        lpart = -sg * lam_n;
                        gpart = sg * (ti * (rm_n - lam_n * rd_n));
                        grec[R::LAM + c] = lam_n; grec[R::T + c] = t_n;
                    }
                    scr[O_CB + 4 * c] = rpart; scr[O_CB + 4 * c + 1] = lpart; scr[O_CB + 4 * c + 2] = gpart; scr[O_CB + 4 * c + 3] = Gp;
                }
                if (Rv && hasU_r) {
                    const double* recr = Img<SW_B>::a(scr + O_IN + slot_r * ISZ);
                    const double* ltr = sm + L.to + slot_r * TROW;
#pragma unroll
                    for (int i = 0; i < 3 * NC; i++) L.Ef[i] = recr[R::E + i];
#pragma unroll
                    for (int i = 0; i < 4 * NV; i++) L.lt[i] = ltr[i];
                    if (L.r < NX) {
#pragma unroll
                        for (int w = 0; w < NZ; w++)
                            if (hasX_r || w >= NX) scr[O_PBA + L.r * NZ + w] = jcol_dot_r(w, L.Pc, L.Ef, L.lt);
                    }
                }
            GRP_PHASE_END
            // ---- P2: update(ku): one component per lane, row r of the dynamics residual;
            //      riccati(kr): column of M = J'PJ + D, gradient += J'(P rb + p) ------------------------------------------
            GRP_PHASE_BEGIN(lanes)
                if (!L.run) continue;
                double* scr = sm + L.so;
                if (Uv) {
                    double* grec = L.grec;
                    const double* img = scr + O_IN + slot_u * ISZ;
                    const double* rec = Img<SW_B>::a(img); const double* rec2 = Img<SW_B>::b(img);
                    const double* car = scr + O_CAR + par_u * 3 * NX;           // from stage ku+1: pio | dpi | xn
                    double* carn = scr + O_CAR + par_r * 3 * NX;
                    const double* ltk = sm + L.to + slot_u * TROW;
                    const double a_step = L.astep;
                    if (L.r < NC) {
                        // generic component: column q = r of the pose rows E
                        const int q = L.r;
                        const bool isx = L.cq_x >= 0;
                        const bool has = isx ? hasX_u : hasU_u;
                        const double e0 = rec[R::E + q], e1 = rec[R::E + NC + q], e2 = rec[R::E + 2 * NC + q];
                        const double k1 = ltk[L.cq_k1], k2 = ltk[L.cq_k2];
                        const double v1 = jt_comp(L, e0, e1, e2, k1, k2, car);
                        const double v2 = jt_comp(L, e0, e1, e2, k1, k2, car + NX);
                        const double H = hasU_u ? tb.dt * ltk[T_W + L.cq_y] : L.We_c;
                        const double qv = rec[R::Q + L.cq_z], z = rec2[R::Z + L.cq_z], dz = rec2[R::DZ + L.cq_z];
                        const bool haspi = isx && hasX_u;
                        const double pin = haspi ? rec2[R::PI + (isx ? L.cq_x : 0)] : 0.0;
                        const double* cl = scr + O_CB + 4 * L.cq_bl;
                        const double* cu = scr + O_CB + 4 * L.cq_bu;
                        double r = qv + H * z - pin + v1 + H * dz + v2;
                        r += cl[0] + cu[0];
                        const double pin_n = haspi ? pin + a_step * r : 0.0;
                        const double zn = z + a_step * dz;
                        double g = qv + H * zn - pin_n + (v1 + a_step * v2);
                        g += cl[1] + cu[1];
                        if (has) L.ng = grp_maxabs(L.ng, g);
                        if (!isx && hasU_u) L.lru = grp_maxabs(L.lru, r);
                        g += cl[2] + cu[2];
                        const double dg = H + o.reg_prim + (cl[3] + cu[3]);
                        grec[R::Z + L.cq_z] = zn;
                        if (isx) {
                            const int j = L.cq_x;
                            grec[R::PI + j] = pin_n;
                            carn[j] = pin; carn[NX + j] = haspi ? r : 0.0; carn[2 * NX + j] = zn;
                            scr[O_GX + par_u * NX + j] = g; scr[O_DGX + par_u * NX + j] = dg;
                        } else {
                            scr[O_GU + par_u * NV + L.cq_z] = g; scr[O_DGU + par_u * NV + L.cq_z] = dg;
                        }
                    }
                    if (L.r == XL || L.r == YL) {
                        // pose components x, y: unit columns of J, no bounds
                        const int j = L.r == XL ? 0 : 1;
                        const double v1 = car[j], v2 = car[NX + j];              // zero at the terminal stage
                        const double H = hasU_u ? tb.dt * ltk[T_W + j] : L.We_xy;
                        const double qv = rec[R::Q + NU + j], z = rec2[R::Z + NU + j], dz = rec2[R::DZ + NU + j];
                        const double pin = hasX_u ? rec2[R::PI + j] : 0.0;
                        const double r = qv + H * z - pin + v1 + H * dz + v2;
                        const double pin_n = hasX_u ? pin + a_step * r : 0.0;
                        const double zn = z + a_step * dz;
                        const double g = qv + H * zn - pin_n + (v1 + a_step * v2);
                        if (hasX_u) L.ng = grp_maxabs(L.ng, g);
                        grec[R::Z + NU + j] = zn;
                        grec[R::PI + j] = pin_n;
                        carn[j] = pin; carn[NX + j] = hasX_u ? r : 0.0; carn[2 * NX + j] = zn;
                        scr[O_GX + par_u * NX + j] = g; scr[O_DGX + par_u * NX + j] = H + o.reg_prim;
                    }
                    if (hasU_u && L.r < NX) {
                        // dynamics residual, row r, at the new iterate (recomputed from the record: no exchange)
                        const double rb = jrow(L.r, [&](int c) { return rec2[R::Z + c] + a_step * rec2[R::DZ + c]; },
                                               [&](int j) { return rec2[R::Z + NU + j] + a_step * rec2[R::DZ + NU + j]; }, rec + R::E, ltk)
                                          + rec[R::B0 + L.r] - car[2 * NX + L.r];
                        L.nb = grp_maxabs(L.nb, rb);
                        scr[O_RB + par_u * NX + L.r] = rb;
                        grec[R::RB + L.r] = rb;
                    }
                }
                if (Rv && !hasU_r) {
                    // terminal stage: P = diag(We + reg + Gamma), p = g
                    if (L.r < NX) {
                        const double pd = scr[O_DGX + par_r * NX + L.r];
#pragma unroll
                        for (int i = 0; i < NX; i++) L.Pc[i] = pd * sel(i == L.r);
                        scr[O_PV + L.r] = scr[O_GX + par_r * NX + L.r];
                    }
                }
                if (Rv && hasU_r) {
                    const double* recr = Img<SW_B>::a(scr + O_IN + slot_r * ISZ);
                    const double* ltr = sm + L.to + slot_r * TROW;
                    double rbv[NX], pvv[NX];
#pragma unroll
                    for (int i = 0; i < NX; i++) { rbv[i] = scr[O_RB + par_r * NX + i]; pvv[i] = scr[O_PV + i]; }
                    if (hasX_r && L.r < NX) {
                        double col[NX];
#pragma unroll
                        for (int i = 0; i < NX; i++) col[i] = scr[O_PBA + i * NZ + L.r];
                        // (J' p)_r with the lane's column role
                        const bool unit = L.cs_e < 0;
                        const int ce = unit ? 0 : L.cs_e;
                        const double e0 = unit ? sel(L.r == 0) : recr[R::E + ce], e1 = unit ? sel(L.r == 1) : recr[R::E + NC + ce],
                                     e2 = unit ? 0.0 : recr[R::E + 2 * NC + ce];
                        double gg = scr[O_GX + par_r * NX + L.r] + (e0 * pvv[0] + e1 * pvv[1] + e2 * pvv[2]
                                                                   + ltr[L.cs_k1] * scr[O_PV + L.cs_i1] + ltr[L.cs_k2] * scr[O_PV + L.cs_i2]);
#pragma unroll
                        for (int i = 0; i < NX; i++) gg += col[i] * rbv[i];
                        L.gx = gg;
                        const double dg = scr[O_DGX + par_r * NX + L.r];
#pragma unroll
                        for (int wp = 0; wp < NZ; wp++) L.Mx[wp] = jcol_dot_r(wp, col, L.Ef, L.lt) + (wp < NX ? dg * sel(wp == L.r) : 0.0);
                    }
                    if (L.r >= ULB && L.r < ULB + ULN) {
#pragma unroll
                        for (int e = 0; e < NV / ULN; e++) {
                            const int a = (L.r - ULB) * (NV / ULN) + e;
                            double col[NX];
#pragma unroll
                            for (int i = 0; i < NX; i++) col[i] = scr[O_PBA + i * NZ + NX + a];
                            // NOTE: `a` is lane dependent, so this one product indexes the lane's register arrays dynamically and
                            // ptxas keeps the whole lane state addressable in (L1-resident) local memory, loading a field where a phase
                            // needs it, instead of holding ~100 registers of state across all phases and spilling the temporaries of
                            // the hot ones.  Measured against the statically indexed form (which spills 270 B at the 168-register
                            // cap): batch-1 latency 3.9 -> 3.0 ms, 4,096 instances 12.8 -> 10.6 ms, hybrid 46.9 -> 45.3 ms.
                            double gg = scr[O_GU + par_r * NV + a] + jcol_dot_r(NX + a, pvv, L.Ef, L.lt);
#pragma unroll
                            for (int i = 0; i < NX; i++) gg += col[i] * rbv[i];
                            const double dg = scr[O_DGU + par_r * NV + a];
#pragma unroll
                            for (int ap = 0; ap < NV; ap++)
                                scr[O_MUU + ap * NV + a] = jcol_dot_r(NX + ap, col, L.Ef, L.lt) + dg * sel(ap == a);
                            scr[O_GG + a] = gg;
                        }
                    }
                }
            GRP_PHASE_END
            if (!Rv) continue;
            if (!hasU_r) continue;
            // ---- P3: riccati(kr): Cholesky of the control block (every lane), lh, column r of K ----------------
            GRP_PHASE_BEGIN(lanes)
                if (!L.run) continue;
                double* scr = sm + L.so;
                double* grec = L.grec + (Uv ? R::NREC : 0);       // record of stage kr
                double Luu[NLU];
#pragma unroll
                for (int a = 0; a < NV; a++) {
                    double d = scr[O_MUU + a * NV + a];
#pragma unroll
                    for (int c = 0; c < a; c++) d -= Luu[a * (a + 1) / 2 + c] * Luu[a * (a + 1) / 2 + c];
                    const double inv = d > 0.0 ? grp_rsqrt(d) : 0.0;
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
                    double sacc = scr[O_GG + a];
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
                if (hasX_r && L.r < NX) {
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
            if (!hasX_r) continue;
            // ---- P4: riccati(kr): Schur complement -> column r of this stage's cost-to-go, and its gradient ----------
            GRP_PHASE_BEGIN(lanes)
                if (!L.run || L.r >= NX) continue;
                double* scr = sm + L.so;
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    double sacc = L.Mx[i];
#pragma unroll
                    for (int a = 0; a < NV; a++) sacc -= scr[O_KB + a * NX + i] * L.Kc[a];
                    L.Pc[i] = sacc;
                }
                double pvn = L.gx;
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
            L.aN = 1.0; L.aD = -1.0; L.S0 = L.S1 = L.S2 = 0.0; L.dxr = 0.0;
            if (L.run) {
                double* scr = sm + L.so;
                if (L.r < NX) scr[O_CAR + L.r] = 0.0;
            }
            begin_sweep<KIND, 1>(L, sm, ws, tb, 0);
        GRP_PHASE_END
#pragma unroll 1
        for (int k = 0; k <= NSTAGE; k++) {
            constexpr int D = Img<KIND>::D;
            const int slot = k % D, pslot = (k + D - 1) % D;
            const bool hasU = k < NSTAGE, hasX = k > 0, pvalid = k + D - 1 <= NSTAGE;
            GRP_PHASE_BEGIN(lanes)
                begin_stage<KIND, 1>(L);
            GRP_PHASE_END
            // ---- F1: prefetch; s_a = lh_a + K_a . dx on the first NV lanes ------------------------------
            GRP_PHASE_BEGIN(lanes)
                prefetch<KIND, 1>(L, sm, D - 1, pslot, pvalid);
                if (!L.run) continue;
                double* scr = sm + L.so;
                if (!hasU || L.r >= NV) continue;
                const double* img = scr + O_IN + slot * Img<KIND>::SIZE;
                const double* rec = Img<KIND>::a(img); const double* rec2 = Img<KIND>::b(img);
                const double* dx = scr + O_CAR + (k & 1) * NX;
                const int a = L.r;
                double s0 = rec[(DELTA ? R::LHD : R::LH) + a], s1 = 0.0;
                if (hasX) {
#pragma unroll
                    for (int j = 0; j < NX; j++) {
                        if (j & 1) s1 += rec[R::KH + a * NX + j] * dx[j];
                        else s0 += rec[R::KH + a * NX + j] * dx[j];
                    }
                }
                scr[O_GU + a] = s0 + s1;
            GRP_PHASE_END
            // ---- F2: du (every lane); one constraint per lane: ratio test, mu sums; step stores;
            //      next dx (row r) -----------------------------------------------------------------------
            GRP_PHASE_BEGIN(lanes)
                if (!L.run) continue;
                double* scr = sm + L.so;
                const double* img = scr + O_IN + slot * Img<KIND>::SIZE;
                const double* rec = Img<KIND>::a(img); const double* rec2 = Img<KIND>::b(img);
                double* grec = L.grec;
                const double* dx = scr + O_CAR + (k & 1) * NX;
                const double* ltk = sm + L.to + slot * TROW;
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
                auto du_of = [&](int c) { double v = 0.0;
#pragma unroll
                    for (int a = 0; a < NV; a++) v += L.du[a] * sel(a == c);
                    return v; };
                if (L.r < NCT) {
                    const int c = L.r;
                    const bool actc = (L.ct_u >= 0) ? hasU : hasX;
                    if (actc) {
                        double dzw = L.ct_x >= 0 ? dx[L.ct_x >= 0 ? L.ct_x : 0] : du_of(L.ct_u);
                        if (DELTA) dzw += rec2[R::DZA + L.ct_z];
                        const double sg = L.ct_s;
                        const double lam = rec2[R::LAM + c], t = rec2[R::T + c], zb = rec2[R::Z + L.ct_z];
                        const double rd = sg * (rec[R::DLB + c] - zb) + t;
                        double rm = lam * t - o.tau_min;
                        if (DELTA) rm += mcw * rec2[R::MC + c] - L.sigmu;
                        const double dt = sg * dzw - rd;
                        const double dl = -(lam * dt + rm) / t;
                        if (!DELTA) grec[R::MC + c] = dt * dl;
                        // ratio test (HPIPM keeps the negated step alpha = aN / aD, aD < 0): alpha * d > n  <=>  aN * d < n * aD
                        if (L.aN * dl < lam * L.aD) { L.aN = lam; L.aD = dl; }
                        if (L.aN * dt < t * L.aD) { L.aN = t; L.aD = dt; }
                        L.S0 += lam * t;
                        L.S1 += lam * dt + t * dl;
                        L.S2 += dl * dt;
                    }
                }
                // step components -> DZA (predictor) / DZ (final): lane j stores state j, lane a < NV control a too
                if (L.r < NX) {
                    double dzw = L.dxr;
                    if (!DELTA) grec[R::DZA + NU + L.r] = dzw;
                    else grec[R::DZ + NU + L.r] = dzw + rec2[R::DZA + NU + L.r];
                }
                if (L.r < NV) {
                    double dzw = du_of(L.r);
                    if (!DELTA) grec[R::DZA + L.r] = dzw;
                    else grec[R::DZ + L.r] = dzw + rec2[R::DZA + L.r];
                }
                if (hasU && L.r < NX) {
                    double xn = jrow(L.r, du_of, [&](int j) { return dx[j]; }, rec + R::E, ltk);
                    if (!DELTA) xn += rec[R::RB + L.r];
                    L.dxr = xn;
                    scr[O_CAR + ((k & 1) ^ 1) * NX + L.r] = xn;
                }
            GRP_PHASE_END
        }
        reduce(lanes, sm, 4, 1,
               [](const Lane& L, int q) { return q == 0 ? L.aN / L.aD : q == 1 ? L.S0 : q == 2 ? L.S1 : L.S2; },
               [](Lane& L, int q, double v) { if (q == 0) L.alpha = v; else if (q == 1) L.S0 = v; else if (q == 2) L.S1 = v; else L.S2 = v; });
    }

    // =========================================================================================
    // delta backward sweep (Rti::stage_Bd): right-hand side only in the complementarity rows
    // =========================================================================================
    NMPC_HD static void sweep_Bd(Lane* lanes, double* sm, double* ws, const Tables& tb, double mcw)
    {
        GRP_PHASE_BEGIN(lanes)
            if (L.run) {
                double* scr = sm + L.so;
                if (L.r < NX) scr[O_CAR + L.r] = 0.0;
            }
            begin_sweep<SW_BD, -1>(L, sm, ws, tb, NSTAGE);
        GRP_PHASE_END
#pragma unroll 1
        for (int s = 0; s <= NSTAGE; s++) {
            constexpr int D = Img<SW_BD>::D;
            const int k = NSTAGE - s, slot = s % D, pslot = (s + D - 1) % D;
            const bool hasU = k < NSTAGE, hasX = k > 0, pvalid = s + D - 1 <= NSTAGE;
            GRP_PHASE_BEGIN(lanes)
                begin_stage<SW_BD, -1>(L);
            GRP_PHASE_END
            // ---- D1: q = J' dp + complementarity terms, one component per lane -----------------------
            GRP_PHASE_BEGIN(lanes)
                prefetch<SW_BD, -1>(L, sm, D - 1, pslot, pvalid);
                if (!L.run) continue;
                const double* ltk = sm + L.to + slot * TROW;
                double* scr = sm + L.so;
                const double* img = scr + O_IN + slot * Img<SW_BD>::SIZE;
                const double* rec = Img<SW_BD>::a(img); const double* rec2 = Img<SW_BD>::b(img);
                const double* dp = scr + O_CAR + (s & 1) * NX;
                if (L.r < NC) {
                    const int q = L.r;
                    const bool isx = L.cq_x >= 0;
                    double qv = 0.0;
                    if (hasU) {
                        const double e0 = rec[R::E + q], e1 = rec[R::E + NC + q], e2 = rec[R::E + 2 * NC + q];
                        qv = jt_comp(L, e0, e1, e2, ltk[L.cq_k1], ltk[L.cq_k2], dp);
                    }
                    if (L.cq_bl < NCT && (isx ? hasX : hasU)) {
                        const double tl = rec2[R::T + L.cq_bl], tu = rec2[R::T + L.cq_bu];
                        qv += (mcw * rec2[R::MC + L.cq_bl] - L.sigmu) / tl - (mcw * rec2[R::MC + L.cq_bu] - L.sigmu) / tu;
                    }
                    if (isx) scr[O_GX + L.cq_x] = qv; else scr[O_GU + L.cq_z] = qv;
                }
                if (L.r == XL || L.r == YL) {
                    const int j = L.r == XL ? 0 : 1;
                    scr[O_GX + j] = dp[j];            // zero at the terminal stage
                }
            GRP_PHASE_END
            // ---- D2: lh = L^-1 q_u (every lane), dp = q_x - K' lh -------------------------------------
            GRP_PHASE_BEGIN(lanes)
                if (!L.run) continue;
                double* scr = sm + L.so;
                const double* img = scr + O_IN + slot * Img<SW_BD>::SIZE;
                const double* rec = Img<SW_BD>::a(img); const double* rec2 = Img<SW_BD>::b(img);
                double* dpn = scr + O_CAR + ((s & 1) ^ 1) * NX;
                if (hasU) {
#pragma unroll
                    for (int a = 0; a < NV; a++) {
                        double sacc = scr[O_GU + a];
#pragma unroll
                        for (int c = 0; c < a; c++) sacc -= rec[R::LUU + a * (a + 1) / 2 + c] * L.lh[c];
                        L.lh[a] = sacc * rec[R::LUU + a * (a + 1) / 2 + a];
                    }
                    if (L.r == WL) {
                        double* grec = L.grec;
#pragma unroll
                        for (int a = 0; a < NV; a++) grec[R::LHD + a] = L.lh[a];
                    }
                    if (hasX && L.r < NX) {
                        double sacc = scr[O_GX + L.r];
#pragma unroll
                        for (int a = 0; a < NV; a++) sacc -= rec[R::KH + a * NX + L.r] * L.lh[a];
                        dpn[L.r] = sacc;
                    }
                } else if (L.r < NX) dpn[L.r] = scr[O_GX + L.r];
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
            // ---- refill free slots; lane 0 of a slot owns its LaneCtl between the sweeps --------------
            GRP_PHASE_BEGIN(lanes)
                if (!L.act && L.r == 0) {
                    int* q = reinterpret_cast<int*>(sm + L.so + O_RED);
                    *q = grp_fetch_add(next);
                }
            GRP_PHASE_END
            GRP_PHASE_BEGIN(lanes)
                if (!L.act) {
                    const int idx = *reinterpret_cast<const int*>(sm + L.so + O_RED);
                    if (idx < n) {
                        L.act = true;
                        L.first = rs.ctl == nullptr; L.skipB = !L.first;
                        L.gi = rs.list ? rs.list[idx] : idx;
                        L.li = rs.ctl ? idx : L.gi;               // resumed instances live in compacted records, fresh ones in their own
                        const double* wp = We_inst ? We_inst + i0 + L.gi : tb.We;
                        const size_t wl = We_inst ? (size_t)ldWe : 1;
                        L.We_c = (L.r < NC && L.cq_x >= 0) ? wp[(size_t)L.cq_x * wl] : 0.0;
                        L.We_xy = L.r == XL ? wp[0] : (L.r == YL ? wp[wl] : 0.0);
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
                cy.ng = L.ng; cy.nb = L.nb; cy.nd = L.nd; cy.nm = L.nm; cy.musum = L.musum; cy.lru = L.lru;
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
                cy.alpha = L.alpha; cy.S0 = L.S0; cy.S1 = L.S1; cy.S2 = L.S2;
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
                cy.alpha = L.alpha; cy.S0 = L.S0; cy.S1 = L.S1; cy.S2 = L.S2;
                S::after_Fd(CTL(L), cy, o);
            GRP_PHASE_END
            GRP_PHASE_BEGIN(lanes)
                L.run = L.act && CTL(L).fb != 0;
            GRP_PHASE_END
            if (warp_any(lanes, [](const Lane& L) { return L.run; })) {
                // conditional centering (rare): repeat the delta solve without the second-order term
                sweep_Bd(lanes, sm, ws, tb, 0.0);
                sweep_F<true>(lanes, sm, ws, tb, o, 0.0);
                GRP_PHASE_BEGIN(lanes)
                    if (!L.run || L.r != 0) continue;
                    typename S::CarryF cy;
                    cy.alpha = L.alpha; cy.S0 = L.S0; cy.S1 = L.S1; cy.S2 = L.S2;
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
