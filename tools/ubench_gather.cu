// Micro-benchmark of the primitive the two gather kernels (k_flow_tc, k_encode_sigma_tc) are built from: random
// 4- / 8- / 16-byte loads from a table that lives in L2 (one 32-byte sector per lane and load: the four finest
// static-hash levels, the five finest flow-grid levels).  What is the B200's ceiling in sectors per second, and how
// does it depend on the resident threads per SM and on the loads a thread keeps in flight?  This is the on-chip
// roofline the gather stages are measured against in DESIGN.md (no HBM or tensor figure bounds them).
// Build + run (GPU box):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_gather tools/ubench_gather.cu && /tmp/ubench_gather
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// BYTES per load (4, 8, 16), U independent loads in flight per thread and iteration (their addresses do not depend
// on loaded values), SPREAD: 0 = every lane its own random entry (32 sectors per instruction); 1 = the lanes of a warp
// share 8 random sectors (4 lanes per sector: a mid level); 2 = the whole warp reads one random 128-byte line
template <int BYTES, int U, int SPREAD>
__global__ void k_gather(const unsigned char* __restrict__ tab, uint32_t mask, int iters, uint32_t* __restrict__ out) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    uint32_t s = mix(gid * 2654435761u + 12345u), acc = 0;
    for (int i = 0; i < iters; ++i) {
        uint32_t idx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = mix(s + u + 1);
            uint32_t e = s;
            if (SPREAD == 1) e = (__shfl_sync(0xffffffffu, s, lane & ~3u) & ~3u) | (lane & 3u);
            if (SPREAD == 2) e = (__shfl_sync(0xffffffffu, s, 0) & ~31u) | lane;
            idx[u] = (e & mask) * BYTES;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (BYTES == 4) acc ^= __ldg(reinterpret_cast<const uint32_t*>(tab + idx[u]));
            if (BYTES == 8) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(tab + idx[u])); acc ^= v.x ^ v.y; }
            if (BYTES == 16) { const uint4 v = __ldg(reinterpret_cast<const uint4*>(tab + idx[u])); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
        }
    }
    if (acc == 0x12345678u) out[gid & 1023] = acc;   // keeps the loads alive
}

template <int BYTES, int U, int SPREAD>
void run(const char* what, unsigned char* tab, size_t table_bytes, int threads, int ctas_per_sm, int sms, uint32_t* out) {
    const uint32_t entries = (uint32_t)(table_bytes / BYTES), mask = entries - 1;
    const int blocks = sms * ctas_per_sm, iters = 2048 / U;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_gather<BYTES, U, SPREAD><<<blocks, threads>>>(tab, mask, iters / 4, out);   // warm-up: the table reaches L2
    cudaEventRecord(e0);
    k_gather<BYTES, U, SPREAD><<<blocks, threads>>>(tab, mask, iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double loads = (double)blocks * threads * iters * U;
    const double sectors = SPREAD == 0 ? loads : (SPREAD == 1 ? loads / 4 : loads * BYTES / 32);
    printf("%-44s table %4zu MB  %4d thr x %d CTA/SM  U=%2d: %7.3f ms  %7.1f G loads/s  %7.1f G sectors/s  %6.2f sectors/clk/SM\n",
           what, table_bytes >> 20, threads, ctas_per_sm, U, ms, loads / ms * 1e-6, sectors / ms * 1e-6,
           sectors / ms * 1e-6 / sms / 1.965);
}

int main() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t max_bytes = (size_t)512 << 20;
    unsigned char* tab = nullptr;
    uint32_t* out = nullptr;
    cudaMalloc(&tab, max_bytes);
    cudaMalloc(&out, 4096);
    cudaMemset(tab, 1, max_bytes);
    // L1-resident table (64 KB): the same ceiling holds for hits — it is the L1 tag stage (one 128-byte line per clock),
    // not the L2 -> L1 fill path, that caps scattered loads
    run<8, 8, 0>("8 B, every lane its own sector (L1 hits)", tab, (size_t)64 << 10, 768, 1, sms, out);
    run<8, 16, 0>("8 B, every lane its own sector (L1 hits)", tab, (size_t)64 << 10, 1024, 2, sms, out);
    run<4, 8, 2>("4 B, one random 128-byte line per warp (L1 hits)", tab, (size_t)64 << 10, 1024, 2, sms, out);
    for (size_t mb : {32, 128, 512}) {
        const size_t b = mb << 20;
        run<8, 8, 0>("8 B, every lane its own sector", tab, b, 768, 1, sms, out);
        run<8, 8, 0>("8 B, every lane its own sector", tab, b, 1024, 1, sms, out);
        run<8, 8, 0>("8 B, every lane its own sector", tab, b, 1024, 2, sms, out);
        run<8, 16, 0>("8 B, every lane its own sector", tab, b, 768, 1, sms, out);
        run<8, 16, 0>("8 B, every lane its own sector", tab, b, 1024, 2, sms, out);
        run<8, 32, 0>("8 B, every lane its own sector", tab, b, 1024, 2, sms, out);
        run<4, 8, 0>("4 B, every lane its own sector", tab, b, 1024, 1, sms, out);
        run<4, 16, 0>("4 B, every lane its own sector", tab, b, 1024, 2, sms, out);
        run<16, 8, 0>("16 B, every lane its own sector", tab, b, 768, 1, sms, out);
        run<16, 16, 0>("16 B, every lane its own sector", tab, b, 1024, 2, sms, out);
        run<8, 8, 1>("8 B, 8 random sectors per warp (4 lanes each)", tab, b, 768, 1, sms, out);
        run<8, 16, 1>("8 B, 8 random sectors per warp (4 lanes each)", tab, b, 1024, 2, sms, out);
        run<4, 8, 2>("4 B, one random 128-byte line per warp", tab, b, 1024, 1, sms, out);
        run<4, 16, 2>("4 B, one random 128-byte line per warp", tab, b, 1024, 2, sms, out);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
