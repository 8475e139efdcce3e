// Hardware probes behind the tcgen05 MLP-backward kernel (csrc/mlp_bwd_tc.cu): facts the weight-gradient
// GEMMs dW = dY^T X rely on, checked against a host reference on the GPU box:
//   (1) MN-major shared-memory operands (both A and B stored [k = sample row][m or n contiguous], 128-byte
//       swizzled rows): instruction-descriptor major bits 15 / 16, the K = 16 step advancing 16 rows (2048 B);
//   (2) the M = 64 accumulator layout in tensor memory (row r -> lane 32 (r / 16) + r % 16);
//   (3) several threads of one CTA issuing accumulating tcgen05.mma into the SAME accumulator.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_probe tools/umma_probe.cu && /tmp/umma_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t chunk) { return row * 128u + ((chunk ^ (row & 7u)) << 4); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16, fp16 A / B, fp32 D; bit 15 = A is MN-major, bit 16 = B is MN-major
__host__ __device__ constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// X [128 rows][64] and Y [128 rows][64] fp16 row-major in global -> D[m][n] = sum_r X[r][m] * Y[r][n]  (M = 64, N = 64)
// issuers: how many threads (one per warpgroup) issue `reps` accumulating rounds each into the same D
__global__ void __launch_bounds__(512, 1)
k_probe(const __half* X, const __half* Y, float* D, int issuers, int reps) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    const uint32_t tid = threadIdx.x, wg = tid >> 7, t = tid & 127u;
    unsigned char* xs = sm;                 // [128][128 B]
    unsigned char* ys = sm + 16384;
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 32768 + 64);
    const uint32_t bar = base + 32768;      // 4 mbarriers
    if (tid < 128) {
        for (int c = 0; c < 8; ++c) {
            *reinterpret_cast<uint4*>(xs + swz(t, c)) = *reinterpret_cast<const uint4*>(X + t * 64 + 8 * c);
            *reinterpret_cast<uint4*>(ys + swz(t, c)) = *reinterpret_cast<const uint4*>(Y + t * 64 + 8 * c);
        }
    }
    if (tid < 4) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar + 8 * tid));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;\n" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *slot;
    constexpr uint32_t id = idesc(64, 64, 1, 1);
    if (t == 0 && (int)wg < issuers) {
        if (wg == 0) {   // the first round initialises D; the others wait for it
            for (uint32_t k = 0; k < 8; ++k) umma(tmem, desc_sw128(base + k * 2048), desc_sw128(base + 16384 + k * 2048), id, k);
            commit(bar);
        }
    }
    if (tid == 0) mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (t == 0 && (int)wg < issuers) {
        for (int r = 0; r < reps; ++r)
            for (uint32_t k = 0; k < 8; ++k) umma(tmem, desc_sw128(base + k * 2048), desc_sw128(base + 16384 + k * 2048), id, 1u);
        commit(bar + 8 * wg);
        mbar_wait(bar + 8 * wg, wg == 0 ? 1 : 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (tid < 128) {   // lane L of TMEM -> D row (L / 32) * 16 + L % 32 when L % 32 < 16
        const uint32_t q = tid >> 5, l = tid & 31;
        for (int c = 0; c < 4; ++c) {
            uint32_t v[16];
            ld16(tmem + ((q * 32u) << 16) + c * 16, v);
            for (int i = 0; i < 16; ++i) D[tid * 64 + c * 16 + i] = __uint_as_float(v[i]);
        }
        (void)l;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;\n" ::"r"(tmem) : "memory");
}

int main() {
    std::vector<__half> hx(128 * 64), hy(128 * 64);
    std::vector<float> fx(128 * 64), fy(128 * 64);
    srand(1);
    for (int i = 0; i < 128 * 64; ++i) {
        fx[i] = (float)(rand() % 17 - 8) / 8.f; fy[i] = (float)(rand() % 13 - 6) / 4.f;   // exact in fp16, sums exact in fp32
        hx[i] = __float2half(fx[i]); hy[i] = __float2half(fy[i]);
    }
    std::vector<float> ref(64 * 64, 0.f);
    for (int m = 0; m < 64; ++m)
        for (int n = 0; n < 64; ++n) {
            float s = 0.f;
            for (int r = 0; r < 128; ++r) s += fx[r * 64 + m] * fy[r * 64 + n];
            ref[m * 64 + n] = s;
        }
    __half *dx, *dy; float* dd;
    cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dy, hy.size() * 2); cudaMalloc(&dd, 128 * 64 * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dy, hy.data(), hy.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    for (int issuers : {1, 4}) {
        const int reps = issuers == 1 ? 0 : 16;
        cudaMemset(dd, 0, 128 * 64 * 4);
        k_probe<<<1, 512, 40000>>>(dx, dy, dd, issuers, reps);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> out(128 * 64);
        cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost);
        const float mult = 1.f + (float)(issuers * reps);
        // which TMEM lanes hold which rows?  try the documented M = 64 map and report mismatches
        int bad = 0, used_lanes = 0;
        double maxerr = 0;
        for (int m = 0; m < 64; ++m) {
            const int lane = (m / 16) * 32 + (m % 16);
            for (int n = 0; n < 64; ++n) {
                const double err = fabs(out[lane * 64 + n] - mult * ref[m * 64 + n]);
                if (err > 1e-3 * fabs(mult * ref[m * 64 + n]) + 1e-3) ++bad;
                if (err > maxerr) maxerr = err;
            }
        }
        for (int l = 0; l < 128; ++l) {
            bool nz = false;
            for (int n = 0; n < 64; ++n) nz |= out[l * 64 + n] != 0.f;
            used_lanes += nz;
        }
        printf("issuers %d reps %d: status %s, mismatches %d / 4096, max abs err %.4g, non-zero TMEM lanes %d (expect 64), D[0][0] %.3f ref %.3f\n",
               issuers, reps, cudaGetErrorString(e), bad, maxerr, used_lanes, out[0], mult * ref[0]);
    }
    return 0;
}
