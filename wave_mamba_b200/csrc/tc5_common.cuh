// Shared device helpers for the tcgen05 / TMEM / TMA kernels (conv3x3_tc5.cu, pw_dw_tc5.cu).
#pragma once

#include "common.cuh"

namespace wm {
namespace tc5 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: 8-row groups SBO bytes apart, the two 16-byte K chunks LBO bytes apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version for sm_100
    return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u),
          "r"(0u)
        : "memory");
}

// kind::f16 (here: bf16 x bf16 -> fp32), same operand forms; K = 16 per instruction
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u),
          "r"(0u)
        : "memory");
}
// two fp32 -> one register of two bf16 (round to nearest even), `lo` in the low half = lower address
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// Blocks until the phase with parity `parity` has completed.  try_wait suspends the thread in
// hardware for a bounded time; the retry loop is bounded too (a protocol bug traps instead of
// hanging the GPU).
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity)
{
    uint32_t ok = 0;
#pragma unroll 1
    for (int spin = 0; spin < (1 << 20); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(mbar), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}
// Debuggable variant for pipeline bring-up: instead of trapping, a wait that times out records
// `code` (role << 24 | barrier << 16 | iteration) in the global word `err` (first failure wins) and
// returns, so the kernel terminates and the host can report WHICH wait never completed
// (wm_debug_pipeline_error).  Once the word is set every later wait gives up after 1024 probes.
__device__ __forceinline__ void mbar_wait_flag(uint32_t mbar, uint32_t parity, unsigned int *err,
                                               unsigned int code)
{
    uint32_t ok = 0;
    // fast path: a non-blocking probe (the phase has usually completed long ago; try_wait's
    // suspend / wake-up costs a few hundred cycles even then -- measured ~360 per wait)
#pragma unroll 1
    for (int spin = 0; spin < 32; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(mbar), "r"(parity)
            : "memory");
        if (ok) return;
    }
#pragma unroll 1
    for (int spin = 0; spin < (1 << 20); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(mbar), "r"(parity)
            : "memory");
        if (ok) return;
        if ((spin & 1023) == 1023 && *reinterpret_cast<volatile unsigned int *>(err) != 0u) return;
    }
    atomicCAS(err, 0u, code | 0x80000000u);
}
// tcgen05.commit: the mbarrier gets one arrival when every MMA issued so far by this thread is done
__device__ __forceinline__ void mma_commit(uint32_t mbar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(mbar)
        : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// the registers of every tmem_ld32 issued so far may be read after this
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ float tf32_lo(float v)
{
    return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
}

}  // namespace tc5
}  // namespace wm
