// Micro-benchmark: do raw 16-byte texel fetches through the texture path (tex1Dfetch<uint4> on linear memory, point
// sampling, exact bits) run beside LDG.128 gathers, i.e. is TEX a second write-back port for the plane look-ups?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_tex tools/ubench_tex.cu && /tmp/ubench_tex
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

// per iteration and thread: NL LDG.128 + NT tex fetches; SPREAD 0: all lanes of a warp the same texel, 1: consecutive lanes
// step through ~3 adjacent texels (a plane look-up of a warp of samples along a ray)
template <int NL, int NT, int SPREAD>
__global__ void k(const uint4* __restrict__ g, cudaTextureObject_t tex, uint32_t mask, int iters, uint32_t* out) {
    const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t s = warp * 2654435761u + 17u, acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < NL + NT; ++u) {
            s = s * 1664525u + 1013904223u;
            uint32_t idx = ((s >> 8) + (SPREAD ? lane / 12 : 0)) & mask;
            uint4 v;
            if (u < NL) v = __ldg(g + idx);
            else v = tex1Dfetch<uint4>(tex, (int)idx);
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <int NL, int NT, int SPREAD>
void run(const char* what, const uint4* g, cudaTextureObject_t tex, uint32_t mask, uint32_t* out, int sms) {
    const int threads = 768, iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NL, NT, SPREAD><<<sms, threads>>>(g, tex, mask, 64, out);
    cudaEventRecord(e0);
    k<NL, NT, SPREAD><<<sms, threads>>>(g, tex, mask, iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double wl = (double)sms * (threads / 32) * iters * (NL + NT);
    printf("%-64s %7.3f ms  %6.3f warp-fetches/clk/SM  %6.2f clk per warp-fetch\n", what, ms, wl / (ms * 1e-3) / sms / 1.965e9,
           1.0 / (wl / (ms * 1e-3) / sms / 1.965e9));
}
int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const uint32_t n = 1u << 18;   // 4 MB of texels: L2-resident, mostly L1 hits within a warp
    uint4* g; uint32_t* out; cudaMalloc(&g, (size_t)n * 16); cudaMalloc(&out, 64); cudaMemset(g, 1, (size_t)n * 16);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = g;
    rd.res.linear.desc = cudaCreateChannelDesc<uint4>(); rd.res.linear.sizeInBytes = (size_t)n * 16;
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    printf("create: %s\n", cudaGetErrorString(cudaCreateTextureObject(&tex, &rd, &td, nullptr)));
    for (uint32_t mask : {2047u, n - 1}) {
        printf("-- table of %u texels\n", mask + 1);
        run<8, 0, 0>("8 x LDG.128, warp-shared texel", g, tex, mask, out, sms);
        run<0, 8, 0>("8 x tex1Dfetch<uint4>, warp-shared texel", g, tex, mask, out, sms);
        run<4, 4, 0>("4 x LDG.128 + 4 x tex1Dfetch, warp-shared texel", g, tex, mask, out, sms);
        run<6, 2, 0>("6 x LDG.128 + 2 x tex1Dfetch, warp-shared texel", g, tex, mask, out, sms);
        run<8, 0, 1>("8 x LDG.128, ~3 adjacent texels per warp", g, tex, mask, out, sms);
        run<0, 8, 1>("8 x tex1Dfetch<uint4>, ~3 adjacent texels per warp", g, tex, mask, out, sms);
        run<6, 2, 1>("6 x LDG.128 + 2 x tex1Dfetch, ~3 adjacent texels per warp", g, tex, mask, out, sms);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
