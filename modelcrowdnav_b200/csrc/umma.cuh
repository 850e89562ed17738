// umma.cuh -- thin inline-PTX layer over the Blackwell tensor-core path used by lookahead_tc.cu:
// tcgen05.mma (kind::f16, cta_group::1, M=128) with operands in shared memory (SWIZZLE_NONE, K-major
// canonical layout), accumulators in TMEM, tcgen05.ld for the epilogue, mbarrier completion.
//
// Operand layout used everywhere in this library ("chunked K-major"):
//   element (row r, k) of an R-row operand lives at byte   (k/8) * (R*16) + r*16 + (k%8)*2
// i.e. 8-element (16-byte) K-chunks; within a chunk the rows are contiguous.  In UMMA terms: core matrix =
// 8 rows x 16 B contiguous (128 B), SBO (8-row group stride) = 128 B, LBO (K-chunk stride) = R*16 B.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// SM100 shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), SWIZZLE_NONE, version 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): F16 x F16 -> F32, both operands K-major, M = 128.
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(int N)
{
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// Same with the B operand MN-major (idesc bit 16): B is read from a 128-row activation-style image whose
// ROWS are the K index and whose columns are N, i.e. element (k, n) at (n/8)*2048 + k*16 + (n%8)*2 -- exactly
// what epilogue_to_smem writes.  In descriptor terms: SBO (8-wide N-group stride) = 2048, LBO (8-row K-group
// stride) = 128; each K=16 step advances the start address by 256 B.
__device__ __forceinline__ void mma_layer_bmn(uint32_t tmem_d, uint32_t a_base, uint32_t a_rows, uint32_t b_base,
                                              int K, int N, bool accumulate_first)
{
    const uint32_t idesc = make_idesc_f16(N) | (1u << 16);
    const uint32_t a_lbo = a_rows * 16;
    for (int s = 0; s < K / 16; ++s) {
        const uint64_t ad = make_desc(a_base + (uint32_t)s * 2 * a_lbo, a_lbo, 128);
        const uint64_t bd = make_desc(b_base + (uint32_t)s * 256, 128, 2048);
        mma_f16(tmem_d, ad, bd, idesc, (s > 0 || accumulate_first) ? 1u : 0u);
    }
}

// A operand from TMEM (lane = row, 32-bit column c holds k = 2c (low half) and 2c+1 (high half)), B from smem.
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// D[128 x N] (+)= A[128 x K](TMEM) * B[N x K]^T(smem, chunked K-major, b_rows rows)
__device__ __forceinline__ void mma_layer_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_base, uint32_t b_rows, int K, int N,
                                             bool accumulate_first)
{
    const uint32_t idesc = make_idesc_f16(N);
    const uint32_t b_lbo = b_rows * 16;
    for (int s = 0; s < K / 16; ++s) {
        const uint64_t bd = make_desc(b_base + (uint32_t)s * 2 * b_lbo, b_lbo, 128);
        mma_f16_ts(tmem_d, tmem_a + (uint32_t)s * 8, bd, idesc, (s > 0 || accumulate_first) ? 1u : 0u);
    }
}

// A from TMEM, B MN-major from an activation-style smem image (rows = K index): the two group reductions
// (mean of mlp1 outputs, attention-weighted feature sum) are products with a constant 0/1 matrix kept in TMEM.
__device__ __forceinline__ void mma_layer_ts_bmn(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_base, int K, int N, bool accumulate_first)
{
    const uint32_t idesc = make_idesc_f16(N) | (1u << 16);
    for (int s = 0; s < K / 16; ++s) {
        const uint64_t bd = make_desc(b_base + (uint32_t)s * 256, 128, 2048);
        mma_f16_ts(tmem_d, tmem_a + (uint32_t)s * 8, bd, idesc, (s > 0 || accumulate_first) ? 1u : 0u);
    }
}

__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint32_t mbar_saddr)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(mbar_saddr) : "memory");
}

// mbarrier wait that traps instead of hanging the GPU if the producer never arrives (protocol bug guard)
__device__ __forceinline__ void mbar_wait_guarded(uint32_t mbar_saddr, uint32_t parity)
{
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(ok) : "r"(mbar_saddr), "r"(parity) : "memory");
        if (ok) break;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

// D[128 x N] (+)= A[128 x K] * B[N x K]^T, K a multiple of 16; one elected thread calls this.
// a_base/b_base: shared addresses of chunked K-major operands with a_rows / b_rows rows.
__device__ __forceinline__ void mma_layer(uint32_t tmem_d, uint32_t a_base, uint32_t a_rows, uint32_t b_base,
                                          uint32_t b_rows, int K, int N, bool accumulate_first)
{
    const uint32_t idesc = make_idesc_f16(N);
    const uint32_t a_lbo = a_rows * 16, b_lbo = b_rows * 16;
    for (int s = 0; s < K / 16; ++s) {
        const uint64_t ad = make_desc(a_base + (uint32_t)s * 2 * a_lbo, a_lbo, 128);
        const uint64_t bd = make_desc(b_base + (uint32_t)s * 2 * b_lbo, b_lbo, 128);
        mma_f16(tmem_d, ad, bd, idesc, (s > 0 || accumulate_first) ? 1u : 0u);
    }
}

__device__ __forceinline__ void commit(uint32_t mbar_saddr)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar_saddr) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar_saddr, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar_saddr), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar_saddr, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}\n"
        :: "r"(mbar_saddr), "r"(parity) : "memory");
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// one warp allocates `ncols` TMEM columns (power of two >= 32) and publishes the base address in smem
__device__ __forceinline__ void tmem_alloc(uint32_t dst_saddr, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_saddr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start (prologue:
// weight image into shared memory, TMEM allocation, barrier set-up) while the kernel before it in the stream drains; pdl_wait()
// returns once that kernel has completed and its writes are visible (a no-op under a plain launch); pdl_launch_dependents() in the
// earlier kernel lets the next one be scheduled as soon as SM resources free up.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- CTA pair (cta_group::2): one UMMA spans the two SMs of a 2-CTA cluster -----------------------------
// M = 256: CTA r of the pair owns rows [128 r, 128 r + 128) -- its A tile in its own shared memory, its
// accumulator in its own TMEM -- and supplies rows [r N/2, (r + 1) N/2) of the K-major B operand (N split in
// halves), so each SM keeps only HALF of every weight matrix resident.  The shared-memory descriptors are CTA-local
// offsets that both CTAs interpret in their own shared memory.  Only the rank-0 CTA issues; completion is
// multicast to the same mbarrier offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// Remote arrival with the default .release.cta semantics.  A .release.cluster arrival compiles to
// MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR (measured ~1500 cycles per hand-over with global loads in flight).  The
// operands handed over here live in the ARRIVING CTA's own shared memory and are read by its own SM's tensor
// core; fence.proxy.async before the arrival is what orders them, as in 2-SM GEMM epilogue -> mainloop signalling.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t mbar_saddr, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(ok) : "r"(mbar_saddr), "r"(parity) : "memory");
    return ok;
}
// non-blocking poll (try_wait may suspend the thread for a while; the issuer polls two contexts)
__device__ __forceinline__ uint32_t mbar_test_wait_cluster(uint32_t mbar_saddr, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(ok) : "r"(mbar_saddr), "r"(parity) : "memory");
    return ok;
}
// guarded (traps instead of hanging) cluster-scope acquire wait
__device__ __forceinline__ void mbar_wait_cluster(uint32_t mbar_saddr, uint32_t parity)
{
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(mbar_saddr, parity))
        if (clock64() - t0 > 4000000000LL) __trap();
}
__host__ __device__ __forceinline__ uint32_t make_idesc_f16_m256(int N)
{
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[256 x N] (+)= A[256 x K] * B[N x K]^T over the CTA pair: a_base = this CTA's 128-row A tile (a_rows = 128),
// b_base = this CTA's N/2-row half of B.  N a multiple of 16.
__device__ __forceinline__ void mma_layer_2(uint32_t tmem_d, uint32_t a_base, uint32_t a_rows, uint32_t b_base, int K, int N,
                                            bool accumulate_first)
{
    const uint32_t idesc = make_idesc_f16_m256(N);
    const uint32_t a_lbo = a_rows * 16, b_lbo = (uint32_t)(N / 2) * 16;
    for (int s = 0; s < K / 16; ++s) {
        const uint64_t ad = make_desc(a_base + (uint32_t)s * 2 * a_lbo, a_lbo, 128);
        const uint64_t bd = make_desc(b_base + (uint32_t)s * 2 * b_lbo, b_lbo, 128);
        mma_f16_2(tmem_d, ad, bd, idesc, (s > 0 || accumulate_first) ? 1u : 0u);
    }
}
// A operand from TMEM (each CTA of the pair supplies its own 128 lanes; 32-bit column c holds k = 2c, 2c+1)
__device__ __forceinline__ void mma_f16_2_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-steps [s0, s1) of D[256 x N] (+)= A(TMEM, packed fp16 from column tmem_a) * B[N x K]^T; b_base = this CTA's N/2-row half
__device__ __forceinline__ void mma_steps_2_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_base, int s0, int s1, int N,
                                               bool accumulate_first)
{
    const uint32_t idesc = make_idesc_f16_m256(N);
    const uint32_t b_lbo = (uint32_t)(N / 2) * 16;
    for (int s = s0; s < s1; ++s) {
        const uint64_t bd = make_desc(b_base + (uint32_t)s * 2 * b_lbo, b_lbo, 128);
        mma_f16_2_ts(tmem_d, tmem_a + (uint32_t)(s - s0) * 8, bd, idesc, (s > s0 || accumulate_first) ? 1u : 0u);
    }
}
__device__ __forceinline__ void commit_2(uint32_t mbar_saddr, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(mbar_saddr), "h"(cta_mask) : "memory");
}
// executed by one warp of EACH CTA of the pair
__device__ __forceinline__ void tmem_alloc_2(uint32_t dst_saddr, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_saddr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}

// TMA bulk copy global -> this CTA's shared memory (no tensor map: a contiguous, 16-byte aligned span); the bytes
// are accounted on `mbar` (one arrival + expect_tx is posted here, so initialise the barrier with count 1)
__device__ __forceinline__ void bulk_load(uint32_t dst_saddr, const void *src, uint32_t bytes, uint32_t mbar_saddr)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar_saddr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_saddr), "l"(src), "r"(bytes), "r"(mbar_saddr) : "memory");
}

// 32 lanes x 32 columns of fp32: thread i of the warp receives columns [col, col+32) of TMEM lane (base lane + i)
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

__device__ __forceinline__ void ld8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr) : "memory");
}

// pack two fp32 into f16x2 (lo = a, hi = b), optional ReLU fused in the conversion
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi)
{
    uint32_t d;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint32_t pack_f16x2_relu(float lo, float hi)
{
    uint32_t d;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// byte offset of the 16-byte chunk holding (row r, k in [8c, 8c+8)) in an R-row chunked K-major operand
__host__ __device__ __forceinline__ uint32_t chunk_off(uint32_t R, uint32_t r, uint32_t c) { return c * (R * 16) + r * 16; }

// Epilogue: TMEM columns [col, col + ncols) of this thread's lane -> fp16 row `row` of a 128-row operand in
// shared memory starting at K-chunk kc0 (ncols a multiple of 16).  RELU selects cvt.rn.relu.
template <bool RELU>
__device__ __forceinline__ void cvt_store8(const uint32_t *v, uint8_t *dst)
{
    const float *f = reinterpret_cast<const float *>(v);
    uint4 o;
    if (RELU) {
        o.x = pack_f16x2_relu(f[0], f[1]); o.y = pack_f16x2_relu(f[2], f[3]);
        o.z = pack_f16x2_relu(f[4], f[5]); o.w = pack_f16x2_relu(f[6], f[7]);
    } else {
        o.x = pack_f16x2(f[0], f[1]); o.y = pack_f16x2(f[2], f[3]);
        o.z = pack_f16x2(f[4], f[5]); o.w = pack_f16x2(f[6], f[7]);
    }
    *reinterpret_cast<uint4 *>(dst) = o;
}

// scaled, no ReLU: used for the group mean (sum / H)
__device__ __forceinline__ void scale_store8(const uint32_t *v, float scale, uint8_t *dst)
{
    const float *f = reinterpret_cast<const float *>(v);
    uint4 o;
    o.x = pack_f16x2(f[0] * scale, f[1] * scale); o.y = pack_f16x2(f[2] * scale, f[3] * scale);
    o.z = pack_f16x2(f[4] * scale, f[5] * scale); o.w = pack_f16x2(f[6] * scale, f[7] * scale);
    *reinterpret_cast<uint4 *>(dst) = o;
}

// TMEM columns [col, col + ncols) * scale -> fp16 row of a 128-row smem operand (ncols a multiple of 16)
__device__ __forceinline__ void epilogue_scaled_to_smem(uint32_t taddr_lane, int col, int ncols, float scale, uint8_t *dst,
                                                        int row, int kc0)
{
    int done = 0;
    while (ncols - done >= 32) {
        uint32_t v[32];
        ld32(taddr_lane + col + done, v);
        wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) scale_store8(v + q * 8, scale, dst + chunk_off(128, row, kc0 + done / 8 + q));
        done += 32;
    }
    if (ncols - done >= 16) {
        uint32_t v[16];
        ld16(taddr_lane + col + done, v);
        wait_ld();
#pragma unroll
        for (int q = 0; q < 2; ++q) scale_store8(v + q * 8, scale, dst + chunk_off(128, row, kc0 + done / 8 + q));
        done += 16;
    }
}

template <bool RELU, bool NARROW = false>
__device__ __forceinline__ void epilogue_to_smem(uint32_t taddr_lane, int col, int ncols, uint8_t *dst, int row, int kc0)
{
    int done = 0;
    while (!NARROW && ncols - done >= 64) {          // two TMEM loads in flight per wait
        uint32_t v[32], u[32];
        ld32(taddr_lane + col + done, v);
        ld32(taddr_lane + col + done + 32, u);
        wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) cvt_store8<RELU>(v + q * 8, dst + chunk_off(128, row, kc0 + done / 8 + q));
#pragma unroll
        for (int q = 0; q < 4; ++q) cvt_store8<RELU>(u + q * 8, dst + chunk_off(128, row, kc0 + done / 8 + 4 + q));
        done += 64;
    }
    if (!NARROW && ncols - done >= 48) {
        uint32_t v[32], u[16];
        ld32(taddr_lane + col + done, v);
        ld16(taddr_lane + col + done + 32, u);
        wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) cvt_store8<RELU>(v + q * 8, dst + chunk_off(128, row, kc0 + done / 8 + q));
#pragma unroll
        for (int q = 0; q < 2; ++q) cvt_store8<RELU>(u + q * 8, dst + chunk_off(128, row, kc0 + done / 8 + 4 + q));
        done += 48;
    }
    while (ncols - done >= 32) {
        uint32_t v[32];
        ld32(taddr_lane + col + done, v);
        wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) cvt_store8<RELU>(v + q * 8, dst + chunk_off(128, row, kc0 + done / 8 + q));
        done += 32;
    }
    if (ncols - done >= 16) {
        uint32_t v[16];
        ld16(taddr_lane + col + done, v);
        wait_ld();
#pragma unroll
        for (int q = 0; q < 2; ++q) cvt_store8<RELU>(v + q * 8, dst + chunk_off(128, row, kc0 + done / 8 + q));
        done += 16;
    }
}

// TMEM[col_src, +ncols) fp32 of this thread's lane -> fp16 pairs packed into TMEM[col_dst, +ncols/2), ascending,
// safe in place (col_dst == col_src): every store lands behind the columns still to be read.
// SCALE: multiply by `scale` (no ReLU); otherwise ReLU fused in the conversion.
// NO64: at most one 32-column load per wait.  Measured in the CTA-pair kernels (88 registers): the "two loads per wait" form
// costs 2.5 % of the lookahead at EACH of the three places it was used (mlp1.0, attention.0, mlp3 hidden layers) although it
// saves a tcgen05.ld round trip; a 48-column tail in one trip was 15 % slower still.  16-column pieces: +0.6 %, within noise.
template <bool RELU, bool NO64 = false>
__device__ __forceinline__ void compact_to_tmem(uint32_t tlane, int col_src, int ncols, int col_dst, float scale)
{
    int done = 0;
    while (!NO64 && ncols - done >= 64) {          // two TMEM loads in flight per wait (a load + wait round trip is ~290 cycles)
        uint32_t v[32], u[32], w[16];
        ld32(tlane + col_src + done, v);
        ld32(tlane + col_src + done + 32, u);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);
            w[j] = RELU ? pack_f16x2_relu(a, b) : pack_f16x2(a * scale, b * scale);
        }
        st16(tlane + col_dst + done / 2, w);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float a = __uint_as_float(u[2 * j]), b = __uint_as_float(u[2 * j + 1]);
            w[j] = RELU ? pack_f16x2_relu(a, b) : pack_f16x2(a * scale, b * scale);
        }
        st16(tlane + col_dst + done / 2 + 16, w);
        done += 64;
    }
    while (ncols - done >= 32) {
        uint32_t v[32], w[16];
        ld32(tlane + col_src + done, v);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);
            w[j] = RELU ? pack_f16x2_relu(a, b) : pack_f16x2(a * scale, b * scale);
        }
        st16(tlane + col_dst + done / 2, w);
        done += 32;
    }
    while (ncols - done >= 16) {
        uint32_t v[16], w[8];
        ld16(tlane + col_src + done, v);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);
            w[j] = RELU ? pack_f16x2_relu(a, b) : pack_f16x2(a * scale, b * scale);
        }
        st8(tlane + col_dst + done / 2, w);
        done += 16;
    }
    wait_st();
}

// compact_to_tmem<RELU = true> with a per-ROW fp32 bias vector added to the accumulator first (bias = 16-byte aligned,
// ncols floats for this thread's row, or nullptr for a padding row): the occupancy-map half of mlp1.0, which does not depend
// on the action and is computed once per (env, human) by tc_om_bias_kernel.  16-column pieces keep the bias + accumulator
// registers under the kernel's 88-register cap; the bias loads are issued before the TMEM load is waited for.
__device__ __forceinline__ void compact_to_tmem_bias(uint32_t tlane, int col_src, int ncols, int col_dst, const float *__restrict__ bias)
{
    for (int done = 0; done + 16 <= ncols; done += 16) {
        uint32_t v[16], w[8];
        float4 b[4];
        ld16(tlane + col_src + done, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = bias ? __ldg(reinterpret_cast<const float4 *>(bias + done) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            w[2 * j] = pack_f16x2_relu(__uint_as_float(v[4 * j]) + b[j].x, __uint_as_float(v[4 * j + 1]) + b[j].y);
            w[2 * j + 1] = pack_f16x2_relu(__uint_as_float(v[4 * j + 2]) + b[j].z, __uint_as_float(v[4 * j + 3]) + b[j].w);
        }
        st8(tlane + col_dst + done / 2, w);
    }
    wait_st();
}

}  // namespace umma
