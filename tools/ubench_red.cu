// Micro-benchmark of the gradient-scatter primitives k_encode_bwd / k_flowgrid_bwd are built from:
// what bounds red.global.add (lane operations, 32-byte sectors or instructions?) and how fast
// shared-memory float atomics are.  Build + run (GPU box):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ubench_red tools/ubench_red.cu && gpurun_out/ubench_red
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ void red1(float* p, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void red2(float* p, float a) {
    asm volatile("red.global.add.v2.f32 [%0], {%1,%1};" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void red4(float* p, float a) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(a) : "memory");
}

// MODE 0: 32 lanes, each its own random float            (32 sectors / instr)
// MODE 1: 32 lanes, each its own random 16 B, red.v4      (32 sectors / instr)
// MODE 2: groups of 8 lanes add 32 contiguous bytes       (4 sectors / instr, 32 lane ops)
// MODE 3: same bytes as 2 by 2 lanes of a group, red.v4   (4 sectors / instr, 8 lane ops)
// MODE 4: 32 lanes add 128 contiguous bytes               (4 sectors, 1 line)
// MODE 5: same bytes as 4 by 8 lanes, red.v4
// MODE 6: 32 lanes, each its own random 8 B, red.v2
// MODE 7: same as 2 but one lane per group adds a whole sector with two red.v4 (8 lane ops, 2 instr)
template <int MODE>
__global__ void k_red(float* tab, uint32_t mask_f, int iters) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    uint32_t s = mix(gid * 2654435761u + 12345u);
    for (int i = 0; i < iters; ++i) {
        s = mix(s + i);
        if (MODE == 0) red1(tab + (s & mask_f), 1.0f);
        if (MODE == 1) red4(tab + ((s & mask_f) & ~3u), 1.0f);
        if (MODE == 6) red2(tab + ((s & mask_f) & ~1u), 1.0f);
        if (MODE == 2 || MODE == 3 || MODE == 7) {
            const uint32_t sg = __shfl_sync(0xffffffffu, s, lane & ~7u);
            float* p = tab + ((sg & mask_f) & ~7u);
            if (MODE == 2) red1(p + (lane & 7), 1.0f);
            if (MODE == 3 && (lane & 3) == 0) red4(p + (lane & 4), 1.0f);
            if (MODE == 7 && (lane & 7) == 0) { red4(p, 1.0f); red4(p + 4, 1.0f); }
        }
        if (MODE == 4 || MODE == 5) {
            const uint32_t sg = __shfl_sync(0xffffffffu, s, 0);
            float* p = tab + ((sg & mask_f) & ~31u);
            if (MODE == 4) red1(p + lane, 1.0f);
            if (MODE == 5 && (lane & 3) == 0) red4(p + lane, 1.0f);
        }
    }
}

// shared-memory float atomics: every lane adds to a random word of a 128 KB table
__global__ void k_red_shared(float* out, int iters, uint32_t words_mask) {
    extern __shared__ float sm[];
    for (uint32_t i = threadIdx.x; i <= words_mask; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    uint32_t s = mix((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 777u);
    for (int i = 0; i < iters; ++i) {
        s = mix(s + i);
        atomicAdd(sm + (s & words_mask), 1.0f);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

template <int MODE>
void run(const char* what, float* tab, size_t table_bytes, int lanes_per_instr, int sectors_per_instr) {
    const int blocks = 148 * 8, threads = 256, iters = 256;
    const uint32_t mask_f = (uint32_t)(table_bytes / 4 - 1);
    cudaMemset(tab, 0, table_bytes);
    k_red<MODE><<<blocks, threads>>>(tab, mask_f, 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k_red<MODE><<<blocks, threads>>>(tab, mask_f, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    const double warps = (double)blocks * threads / 32 * iters;
    printf("%-58s table %4zu MB: %7.3f ms  %7.1f G lane-ops/s  %7.1f G sectors/s  %6.1f G warp-instr/s  %7.1f GB/s\n",
           what, table_bytes >> 20, ms, warps * lanes_per_instr / ms * 1e-6, warps * sectors_per_instr / ms * 1e-6,
           warps / ms * 1e-6,
           warps * lanes_per_instr * (MODE == 1 || MODE == 3 || MODE == 5 || MODE == 7 ? 16 : (MODE == 6 ? 8 : 4)) / ms * 1e-6);
}

int main() {
    float* tab = nullptr;
    cudaMalloc(&tab, (size_t)512 << 20);
    for (size_t mb : {8, 128, 512}) {
        const size_t bytes = mb << 20;
        run<0>("red.f32, 32 random floats per instr", tab, bytes, 32, 32);
        run<6>("red.v2, 32 random 8-byte pairs per instr", tab, bytes, 32, 32);
        run<1>("red.v4, 32 random 16-byte quads per instr", tab, bytes, 32, 32);
        run<2>("red.f32, 4 random sectors x 8 contiguous lanes", tab, bytes, 32, 4);
        run<3>("red.v4, 4 random sectors x 2 lanes", tab, bytes, 8, 4);
        run<7>("2 x red.v4 by 1 lane per random sector (4 lanes)", tab, bytes, 8, 4);
        run<4>("red.f32, one random 128-byte line x 32 lanes", tab, bytes, 32, 4);
        run<5>("red.v4, one random 128-byte line x 8 lanes", tab, bytes, 8, 4);
    }
    {
        float* out = nullptr;
        cudaMalloc(&out, 148 * 4);
        cudaFuncSetAttribute(k_red_shared, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
        const int iters = 2048;
        k_red_shared<<<148, 1024, 128 * 1024>>>(out, 8, 32767);
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        k_red_shared<<<148, 1024, 128 * 1024>>>(out, iters, 32767);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        printf("shared atomicAdd(float), random word of 128 KB, 1024 thr x 148 CTAs: %.3f ms  %.1f G lane-ops/s (%.2f per clk per SM at 1.965 GHz)\n",
               ms, 148.0 * 1024 * iters / ms * 1e-6, 148.0 * 1024 * iters / ms * 1e-6 / 148 / 1.965);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
