// Shared-memory staging helpers of the two-steps-per-pass kernels (cheb_pair.cu: one-dimensional x-planes,
// cheb_cube.cu: three-dimensional lattices): mbarriers, bulk async copies (TMA), 128-bit shared-memory accesses.
// Internal, not part of the ABI.
#pragma once

#include <stdint.h>

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
// One contiguous run global -> shared through the TMA unit (SASS UBLKCP); bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ double2 lds_rec(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_rec(uint32_t addr, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};\n" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

// T_{n-1}: read exactly once per launch -- do not let it displace anything in L1
__device__ __forceinline__ double2 ld_prev(const double2 *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
// ... and E_{j-1} of the T2 mode, which the same thread overwrites one iteration later (no .nc)
__device__ __forceinline__ double2 ld_prev_rw(const double2 *p) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ld_table_pred(double &v, const double *p, unsigned take) {
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p ld.global.nc.f64 %0, [%1];\n}\n" : "+d"(v) : "l"(p), "r"(take));
}

// Keep a value in its register: the compiler otherwise re-derives loop invariants (shared-memory
// base, lane offsets) inside the row loop, which is instruction-issue-bound.
__device__ __forceinline__ void pin(uint32_t &v) { asm volatile("" : "+r"(v)); }

}  // namespace
