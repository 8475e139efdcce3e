// Micro-benchmark: 16-byte texel delivery to every lane of a warp when the warp shares the texel — through the L1 data
// pipe (LDG.128, same address in all lanes) or through the constant port (LDC with a warp-uniform dynamic index).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_ldc tools/ubench_ldc.cu && /tmp/ubench_ldc
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__constant__ uint4 c_tab[2048];   // 32 KB

template <int MODE>   // 0 = LDG.128 broadcast, 1 = LDC uniform index, 2 = LDG.128, two distinct texels per warp, 3 = LDC two distinct
__global__ void k(const uint4* __restrict__ g, int iters, uint32_t* out) {
    const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t s = warp * 2654435761u + 17u, acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            s = s * 1664525u + 1013904223u;
            uint32_t idx = (s >> 8) & 2047u;
            if (MODE == 2 || MODE == 3) idx = (idx & ~1u) | (lane >> 4);
            uint4 v;
            if (MODE == 0 || MODE == 2) v = __ldg(g + idx);
            else v = c_tab[idx];
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <int MODE>
void run(const char* what, const uint4* g, uint32_t* out, int sms) {
    const int threads = 768, iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, threads>>>(g, 64, out);
    cudaEventRecord(e0);
    k<MODE><<<sms, threads>>>(g, iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double wl = (double)sms * (threads / 32) * iters * 8;   // warp-level loads
    printf("%-52s %7.3f ms  %6.3f warp-loads/clk/SM  (%5.1f B/clk/SM delivered)\n", what, ms, wl / (ms * 1e-3) / sms / 1.965e9,
           wl * 512 / (ms * 1e-3) / sms / 1.965e9);
}
int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint4* g; uint32_t* out; cudaMalloc(&g, 2048 * 16); cudaMalloc(&out, 64); cudaMemset(g, 1, 2048 * 16);
    run<0>("LDG.128, all lanes one texel (L1 hit)", g, out, sms);
    run<1>("LDC.128, warp-uniform dynamic index", g, out, sms);
    run<2>("LDG.128, two texels per warp (L1 hit)", g, out, sms);
    run<3>("LDC.128, two texels per warp", g, out, sms);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
