// nvsf_b200 — sigma MLP (120 -> 64 -> 16, tcnn FullyFusedMLP of network_dynamic.py:125-135) on the
// 5th-generation tensor cores: tcgen05.mma with both operands in shared memory (K-major, 128-byte
// swizzle), accumulators in tensor memory, epilogues through tcgen05.ld.  sm_100a only.
//
// One CTA = 128 threads = one UMMA M tile of 128 samples; 4 CTAs per SM (50 KB of shared memory and
// 64 TMEM columns each) hide each other's issue -> commit -> epilogue chain.  Per tile:
//   X [128 x 128] fp16  (cp.async, written in the swizzled K-major layout)
//   D1 [128 x 64] = X W1^T        8 x tcgen05.mma (M128 N64 K16), fp32 in TMEM columns [0,64)
//   H  = relu(D1) -> fp16 -> shared memory (re-uses the first K block of X), one row per thread
//   D2 [128 x 16] = H W2^T        4 x tcgen05.mma (M128 N16 K16), TMEM columns [0,16)
//   sigma = exp(D2[:,0]) (trunc_exp forward, activation.py:10), geo = fp16(D2)
// The mma.sync version of this stage (field_split.cu k_sigma_stage) runs the tensor pipe at 57 % with
// math_pipe_throttle as its first stall (profiles/r01_stages_v2_ncu_full.txt); here one thread issues
// twelve instructions per 128 samples and the other 127 only move data.
#include <algorithm>

#include "field_common.cuh"

namespace {

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
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// packed fp16 MLP image (field_common.cuh: W1 at kSigW1 [64][136], W2 at kSigW2 [16][72]) -> the
// swizzled K-major operand images [W1 K block 0][W1 K block 1][W2], copied verbatim to shared memory
__global__ void k_pack_sigma_tc(const __half* __restrict__ mlp, unsigned char* __restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk each
    if (i < (uint32_t)kHidden * 16) {
        const uint32_t r = i >> 4, c16 = i & 15;
        const uint4 v = *reinterpret_cast<const uint4*>(mlp + kSigW1 + r * kLdK128 + c16 * 8);
        *reinterpret_cast<uint4*>(dst + kOffW1 + (c16 >> 3) * (kHidden * 128) + swz(r, c16 & 7)) = v;
    } else if (i < (uint32_t)kHidden * 16 + kGeo * 8) {
        const uint32_t j = i - kHidden * 16, r = j >> 3, c = j & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(mlp + kSigW2 + r * kLdK64 + c * 8);
        *reinterpret_cast<uint4*>(dst + kOffW2 + swz(r, c)) = v;
    }
}

__global__ void __launch_bounds__(kRows, 4)
k_sigma_stage_tc(const unsigned char* __restrict__ wimg, const __half* __restrict__ feat,
                 size_t count, float* __restrict__ sigma_out, __half* __restrict__ geo_out) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // swizzle atoms are 1024-byte aligned
    unsigned char* sm = smem_raw + (base - raw);
    const uint32_t bar = base + kOffBar;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBar + 8);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;

    for (uint32_t i = tid; i < (kW1Bytes + kW2Bytes) / 16; i += kRows)
        reinterpret_cast<uint4*>(sm)[i] = __ldg(reinterpret_cast<const uint4*>(wimg) + i);
    if (tid == 0) mbar_init(bar, 1);
    if (warp == 0) {
        __syncwarp();  // tcgen05.alloc is .sync.aligned: the whole warp, converged
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    fence_async_smem();  // the weight images were written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((warp * 32u) << 16);  // this warp's 32 TMEM lanes
    constexpr uint32_t kIdesc1 = umma_idesc(kRows, kHidden), kIdesc2 = umma_idesc(kRows, kGeo);

    uint32_t phase = 0;
    const size_t n_tiles = (count + kRows - 1) / kRows;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t row0 = tile * kRows;
        // 1. feature rows -> swizzled K-major tile (coalesced 16-byte pieces)
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
            const uint32_t p = k * kRows + tid, r = p >> 4, c16 = p & 15;
            const bool ok = row0 + r < count;
            cp_async16(base + kOffX + (c16 >> 3) * (kRows * 128) + swz(r, c16 & 7),
                       feat + (ok ? row0 + r : 0) * kFeat + c16 * 8, ok);
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
        fence_async_smem();
        __syncthreads();
        // 2. D1 = X W1^T
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (uint32_t k = 0; k < kFeat / 16; ++k)
                umma_f16(tmem, umma_desc(base + kOffX + (k >> 2) * (kRows * 128) + (k & 3) * 32),
                         umma_desc(base + kOffW1 + (k >> 2) * (kHidden * 128) + (k & 3) * 32), kIdesc1,
                         k);
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        // 3. H = relu(D1) as fp16, row `tid`, into the first K block of X (MMA 1 has finished reading it)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t v[16];
            tmem_ld16(tlane + q * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint4 o;
                o.x = pack_half2(fmaxf(__uint_as_float(v[8 * h + 0]), 0.f), fmaxf(__uint_as_float(v[8 * h + 1]), 0.f));
                o.y = pack_half2(fmaxf(__uint_as_float(v[8 * h + 2]), 0.f), fmaxf(__uint_as_float(v[8 * h + 3]), 0.f));
                o.z = pack_half2(fmaxf(__uint_as_float(v[8 * h + 4]), 0.f), fmaxf(__uint_as_float(v[8 * h + 5]), 0.f));
                o.w = pack_half2(fmaxf(__uint_as_float(v[8 * h + 6]), 0.f), fmaxf(__uint_as_float(v[8 * h + 7]), 0.f));
                *reinterpret_cast<uint4*>(sm + kOffX + swz(tid, 2 * q + h)) = o;
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // 4. D2 = H W2^T
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (uint32_t k = 0; k < kHidden / 16; ++k)
                umma_f16(tmem, umma_desc(base + kOffX + k * 32), umma_desc(base + kOffW2 + k * 32),
                         kIdesc2, k);
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        // 5. sigma, geo
        {
            uint32_t v[16];
            tmem_ld16(tlane, v);
            tmem_ld_wait();
            const size_t row = row0 + tid;
            if (row < count) {
                sigma_out[row] = expf(__uint_as_float(v[0]));
                uint4 a, b;
                a.x = pack_half2(__uint_as_float(v[0]), __uint_as_float(v[1]));
                a.y = pack_half2(__uint_as_float(v[2]), __uint_as_float(v[3]));
                a.z = pack_half2(__uint_as_float(v[4]), __uint_as_float(v[5]));
                a.w = pack_half2(__uint_as_float(v[6]), __uint_as_float(v[7]));
                b.x = pack_half2(__uint_as_float(v[8]), __uint_as_float(v[9]));
                b.y = pack_half2(__uint_as_float(v[10]), __uint_as_float(v[11]));
                b.z = pack_half2(__uint_as_float(v[12]), __uint_as_float(v[13]));
                b.w = pack_half2(__uint_as_float(v[14]), __uint_as_float(v[15]));
                uint4* g = reinterpret_cast<uint4*>(geo_out + row * kGeo);
                g[0] = a;
                g[1] = b;
            }
        }
        tc_fence_before();
        __syncthreads();  // TMEM columns and the X tile are free for the next tile
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(kTmemCols)
                     : "memory");
}

bool g_tc_attr = false;

}  // namespace

size_t nvsf_sigma_tc_image_bytes() { return kW1Bytes + kW2Bytes; }

void nvsf_pack_sigma_tc(const __half* mlp, void* dst, cudaStream_t stream) {
    k_pack_sigma_tc<<<nvsf_div_up(kHidden * 16 + kGeo * 8, 128), 128, 0, stream>>>(
        mlp, reinterpret_cast<unsigned char*>(dst));
}

int nvsf_launch_sigma_tc(const void* wimg, const __half* feat, size_t count, float* sigma,
                         __half* geo, int sms, cudaStream_t stream) {
    if (!g_tc_attr) {
        cudaError_t e = cudaFuncSetAttribute(k_sigma_stage_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kTcSmem);
        if (e != cudaSuccess) return (int)e;
        g_tc_attr = true;
    }
    const size_t tiles = (count + kRows - 1) / kRows;
    const int grid = (int)std::min<size_t>(tiles, (size_t)sms * 4);
    k_sigma_stage_tc<<<grid, kRows, kTcSmem, stream>>>(reinterpret_cast<const unsigned char*>(wimg), feat,
                                                        count, sigma, geo);
    return NVSF_OK;
}
