// Per-instance records of the persistent K3 kernel (rti_coop.cuh) and what is shared around them: the record layout
// (GRec), the stage-table geometry, the hand-over conversion from the tile layout of the lockstep sweeps, the queue
// descriptors, and the small device / host-emulation portability helpers.
//
// (History: this file held the first lane-group mapping of K3 - lanes taking table-driven roles and exchanging
// everything through shared memory.  The lane-cooperative mapping of rti_coop.cuh replaced it in round 2: 1.5x faster at
// every batch size, see profiles/README_r02_notes.txt.)
//
// Record layout: ONE contiguous record per (instance, stage) (GRec, all fields of the stage back to back); each sweep
// prefetches exactly the field ranges it needs for the next stage with 16-byte cp.async copies into a double-buffered
// shared-memory image of the record while the current stage is being computed.
//
// The host emulation (tests/host_emul) runs the kernel's code on the CPU: there a phase (GRP_PHASE_BEGIN .. GRP_PHASE_END)
// is a loop over the 32 lanes of an emulated warp.
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
#define NMPC_GRP_DEPTH 3        // stage images in flight per slot of the lane-cooperative kernel (measured 2 / 3 / 4: diff 65,536 32.5 / 32.3 / 34.4 ms)
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

template <class M>
struct RecOps {
    using S = Rti<M>;
    static constexpr int NV = S::NV, NX = S::NX, NU = S::NU, NZ = S::NZ, NY = S::NY, NC = S::NC, NB2 = S::NB2, NLU = S::NLU;
    using R = GRec<NV>;
    // per-stage row of the stage table (Tables::stg): [av | ar | au | ru (4 NV) | 0 | 1 | pad || diagonal of W (NY) | pad]
    static constexpr int LTE = 4 * NV + 2, LT_ZERO = 4 * NV, LT_ONE = 4 * NV + 1;
    static constexpr int T_W = (LTE + 1) & ~1;            // offset of the weights in a row
    static constexpr int TROW = T_W + ((NY + 1) & ~1);
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

};

}  // namespace nmpc
