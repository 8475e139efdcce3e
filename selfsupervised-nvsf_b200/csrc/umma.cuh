// UMMA (tcgen05.mma) building blocks shared by the tensor-core kernels of sigma_tc.cu: shared-memory
// operand layout (K-major, 128-byte swizzle), descriptors, TMEM loads, mbarrier / proxy fences.
#pragma once
#include "field_common.cuh"

namespace umma {

constexpr int kRows = 128;                         // UMMA M
constexpr uint32_t kW1Bytes = kHidden * kFeat * 2; // 2 K blocks of [64 rows][128 B]
constexpr uint32_t kW2Bytes = kGeo * kHidden * 2;  // 1 K block of [16 rows][128 B]
constexpr uint32_t kXBytes = kRows * kFeat * 2;    // 2 K blocks of [128 rows][128 B]
constexpr uint32_t kOffW1 = 0, kOffW2 = kW1Bytes, kOffX = kOffW2 + kW2Bytes;
constexpr uint32_t kOffBar = kOffX + kXBytes;
constexpr size_t kTcSmem = kOffBar + 16 + 1024;    // + slack to align the base to 1024 B
constexpr uint32_t kTmemCols = 64;
static_assert(kOffX % 1024 == 0, "swizzled tiles start on 1024-byte boundaries");

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside one [rows][128 B] K block with the
// 128-byte swizzle (Swizzle<3,4,3>: chunk index xor row mod 8)
__host__ __device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t chunk) {
    return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t smem, const void* gmem, bool valid) {
    const int bytes = valid ? 16 : 0;  // src-size 0: zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem), "l"(gmem), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_128B: start address, LBO (unused for a
// swizzled K-major operand, 1), SBO = 1024 B between 8-row groups, version 1 (sm_100), layout 2.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor of kind::f16: fp16 A and B (K-major both), fp32 D, M x N
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar)
                 : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane in one instruction
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// ---- A operand from tensor memory (the .ts form of tcgen05.mma) ------------------------------------
// A [128 lanes][K] lives in TMEM as packed fp16 pairs: 32-bit column j of lane r = A[r][2j], A[r][2j+1]; one
// K = 16 step reads 8 columns.  The thread that owns accumulator row r (TMEM lane r) converts it and stores it
// back with tcgen05.st; the activations of a hidden layer never pass through shared memory.
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// this thread's accumulator row D[src .. src + 64) -> relu -> fp16 -> packed row H[dst .. dst + 32), dst == src allowed
__device__ __forceinline__ void hidden_to_tmem(uint32_t src, uint32_t dst) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t v[16], o[8];
        tmem_ld16(src + q * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = pack_half2_relu(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
        tmem_st8(dst + q * 8, o);
    }
    tmem_st_wait();
}

// named barrier of one 128-thread warpgroup (barrier 0 is __syncthreads)
__device__ __forceinline__ void wg_barrier(uint32_t wg) {
    asm volatile("bar.sync %0, 128;\n" ::"r"(wg + 1u) : "memory");
}
__device__ __forceinline__ void st_chunk(unsigned char* tile, uint32_t row, uint32_t chunk,
                                         const float (&v)[8]) {
    uint4 o;
    o.x = pack_half2(v[0], v[1]); o.y = pack_half2(v[2], v[3]);
    o.z = pack_half2(v[4], v[5]); o.w = pack_half2(v[6], v[7]);
    *reinterpret_cast<uint4*>(tile + swz(row, chunk)) = o;
}

// relu + fp16 of eight fp32 accumulator values -> one 16-byte chunk of the operand tile
__device__ __forceinline__ void st_chunk_relu(unsigned char* tile, uint32_t row, uint32_t chunk,
                                              const uint32_t* v) {
    uint4 o;
    o.x = pack_half2_relu(__uint_as_float(v[0]), __uint_as_float(v[1]));
    o.y = pack_half2_relu(__uint_as_float(v[2]), __uint_as_float(v[3]));
    o.z = pack_half2_relu(__uint_as_float(v[4]), __uint_as_float(v[5]));
    o.w = pack_half2_relu(__uint_as_float(v[6]), __uint_as_float(v[7]));
    *reinterpret_cast<uint4*>(tile + swz(row, chunk)) = o;
}

// warpgroup-wide OR of a per-thread flag (named barrier with reduction)
__device__ __forceinline__ bool wg_any(uint32_t wg, bool flag) {
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 q, %1, 0;\n"
        "bar.red.or.pred p, %2, 128, q;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(r)
        : "r"((uint32_t)flag), "r"(wg + 1u)
        : "memory");
    return r != 0;
}
}  // namespace umma
