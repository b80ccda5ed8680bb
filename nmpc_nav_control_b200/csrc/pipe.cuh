// Device-only streaming driver for K3: one warp = one tile of 32 instances.  Each sweep walks
// the horizon; the record of the NEXT stage is fetched into shared memory by 1-D bulk async
// copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier while the current stage is
// being computed from the other shared-memory slot.  A warp never synchronises with another
// warp; results leave with plain coalesced 256-byte stores.
#pragma once
#include <cstdint>
#include "rti_core.cuh"

namespace nmpc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// which contiguous field range of each group a sweep kind needs (first, count); 0 = not loaded
template <int NV, int KIND>
struct Plan {
    using R = Rec<NV>;
    static constexpr int lin_first = R::E;
    static constexpr int lin_count = (KIND <= 1) ? R::NF_LIN : (KIND == 3 ? R::DLB - R::E : R::Q - R::E);
    static constexpr int it_first = (KIND == 0) ? R::Z : R::T;
    static constexpr int it_count = (KIND == 0) ? R::NZ : (KIND == 1 ? R::NF_IT : (KIND == 3 ? 2 * R::NB2 : R::PI - R::T));
    static constexpr int st_first = (KIND == 1) ? R::DZ : R::MC;
    static constexpr int st_count = (KIND == 1) ? R::NZ + 2 * R::NB2 : (KIND == 3 ? 2 * R::NB2 : (KIND == 4 ? R::NF_ST - R::MC : 0));
    static constexpr int fa_first = R::LUU;
    static constexpr int fa_count = (KIND == 2) ? R::NF_FA : (KIND == 3 ? R::LHD - R::LUU : (KIND == 4 ? R::LH - R::LUU : 0));
    static constexpr int total = lin_count + it_count + st_count + fa_count;
};

template <int NV>
struct TmaDriver {
    using R = Rec<NV>;
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int SLOT_FIELDS = cmax(cmax(Plan<NV, 0>::total, Plan<NV, 1>::total),
                                            cmax(Plan<NV, 2>::total, cmax(Plan<NV, 3>::total, Plan<NV, 4>::total)));
    static constexpr size_t SMEM_BYTES = 2 * (size_t)SLOT_FIELDS * LANES * sizeof(double) + 2 * sizeof(uint64_t);

    double* tile;        // global tile base (no lane offset)
    double* smem;        // two slots
    uint64_t* bars;      // two mbarriers
    uint32_t parity;     // bit s = parity to wait for on slot s
    int lane;

    template <int KIND>
    __device__ __forceinline__ void issue(int slot, int k)
    {
        using P = Plan<NV, KIND>;
        double* dst = smem + (size_t)slot * SLOT_FIELDS * LANES;
        uint64_t* bar = &bars[slot];
        mbar_expect_tx(bar, (uint32_t)(P::total * LANES * sizeof(double)));
        if (P::lin_count) {
            bulk_g2s(dst, tile + R::OFF_LIN + ((size_t)k * R::NF_LIN + P::lin_first) * LANES, P::lin_count * LANES * 8, bar);
            dst += P::lin_count * LANES;
        }
        if (P::it_count) {
            bulk_g2s(dst, tile + R::OFF_IT + ((size_t)k * R::NF_IT + P::it_first) * LANES, P::it_count * LANES * 8, bar);
            dst += P::it_count * LANES;
        }
        if (P::st_count) {
            bulk_g2s(dst, tile + R::OFF_ST + ((size_t)k * R::NF_ST + P::st_first) * LANES, P::st_count * LANES * 8, bar);
            dst += P::st_count * LANES;
        }
        if (P::fa_count) {
            bulk_g2s(dst, tile + R::OFF_FA + ((size_t)k * R::NF_FA + P::fa_first) * LANES, P::fa_count * LANES * 8, bar);
        }
    }

    template <int KIND>
    __device__ __forceinline__ StageIn view(int slot) const
    {
        using P = Plan<NV, KIND>;
        const double* b = smem + (size_t)slot * SLOT_FIELDS * LANES + lane;
        StageIn v;
        v.lin = b - (size_t)P::lin_first * LANES;
        b += P::lin_count * LANES;
        v.it = b - (size_t)P::it_first * LANES;
        b += P::it_count * LANES;
        v.st = b - (size_t)P::st_first * LANES;
        b += P::st_count * LANES;
        v.fa = b - (size_t)P::fa_first * LANES;
        return v;
    }

    // KIND: 0 = B first, 1 = B, 2 = F predictor, 3 = Bd, 4 = F delta.  Called by all 32 lanes.
    template <int KIND, class F>
    __device__ __forceinline__ void sweep(bool enabled, F&& f)
    {
        constexpr bool backward = (KIND == 0 || KIND == 1 || KIND == 3);
        __syncwarp();
        if (lane == 0) {
            fence_proxy_async();               // earlier generic-proxy stores of this warp -> async-proxy reads
            issue<KIND>(0, backward ? NSTAGE : 0);
        }
        for (int s = 0; s <= NSTAGE; s++) {
            const int k = backward ? NSTAGE - s : s;
            const int slot = s & 1;
            if (s < NSTAGE) {
                __syncwarp();                  // every lane is done reading the other slot
                if (lane == 0) {
                    fence_proxy_async();
                    issue<KIND>(slot ^ 1, backward ? k - 1 : k + 1);
                }
            }
            mbar_wait(&bars[slot], (parity >> slot) & 1u);
            parity ^= (1u << slot);
            if (enabled) f(k, view<KIND>(slot), tile_stage_out<NV>(tile + lane, k));
        }
    }
};

}  // namespace nmpc
