// Warp-level fp16 tensor-core building blocks of the MLP backward kernels (train.cu).
//
// Every operand of the backward GEMMs is fp16 with fp32 accumulation, like tcnn's FullyFusedMLP
// backward.  Gradients arrive with whatever scale the caller's loss (or GradScaler) gives them,
// so each launch multiplies its output-gradient rows by a power of two chosen on the device from
// their largest magnitude (exact) and divides the results by it again — bf16 would need no scale
// but its 8-bit mantissa costs several percent on the cancelling sums of these layers (measured).
// The forward activations are re-computed per tile from the kept inputs instead of being stored.
//
// One CTA = 8 warps = a tile of 128 rows (samples); warp w owns rows [16w, 16w+16) for the
// per-row chain (hidden re-computation, activation gradients, input gradient) and a slice of
// the weight-gradient matrices for the reduction over all 128 rows of the tile.
#pragma once

#include "field_common.cuh"

typedef __half bf16;  // historical name of the GEMM operand type of this file: fp16

constexpr int kBwdRows = 128;  // rows per tile
constexpr int kBwdWarps = 8;
constexpr int kLdH = 72;       // row stride (bf16) of the 64-wide hidden tiles
constexpr int kLdD = 24;       // row stride of the 16-wide output-gradient tile

__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], const void* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
    mma16816(c, a, b0, b1);
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) { return pack_half2(lo, hi); }

template <int NT>
__device__ __forceinline__ void zero1(float (&acc)[NT][4]) {
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
}

// A fragments of this warp's 16 rows of a row-major bf16 tile: a[kk] covers columns [16kk,16kk+16)
template <int KT>
__device__ __forceinline__ void load_a16(const bf16* A, int lda, uint32_t (&a)[KT][4], int lane) {
#pragma unroll
    for (int kk = 0; kk < KT; ++kk)
        ldsm_x4(a[kk], A + (lane & 15) * lda + kk * 16 + (lane >> 4) * 8);
}

// acc[16 x 8NT] += A[16 x 16KT] * W^T,  W in smem as [n][ldw] (k contiguous): the forward product
template <int KT, int NT>
__device__ __forceinline__ void gemm_nt(const uint32_t (&a)[KT][4], const bf16* W, int ldw,
                                        float (&acc)[NT][4], int lane) {
    static_assert(NT % 2 == 0, "NT must be even");
#pragma unroll
    for (int kk = 0; kk < KT; ++kk)
#pragma unroll
        for (int j = 0; j < NT; j += 2) {
            uint32_t b[4];
            ldsm_x4(b, W + ((j + (lane >> 4)) * 8 + (lane & 7)) * ldw + kk * 16 +
                           ((lane >> 3) & 1) * 8);
            mma_bf16(acc[j], a[kk], b[0], b[1]);
            mma_bf16(acc[j + 1], a[kk], b[2], b[3]);
        }
}

// acc[16 x 8NT] += A[16 x 16KT] * W[:, n0 : n0+8NT],  W in smem as [k][ldw] (n contiguous): the
// activation-gradient product dX = dY * W with W stored [out][in]
template <int KT, int NT>
__device__ __forceinline__ void gemm_nn(const uint32_t (&a)[KT][4], const bf16* W, int ldw, int n0,
                                        float (&acc)[NT][4], int lane) {
    static_assert(NT % 2 == 0, "NT must be even");
#pragma unroll
    for (int kk = 0; kk < KT; ++kk)
#pragma unroll
        for (int j = 0; j < NT; j += 2) {
            uint32_t b[4];
            ldsm_x4_t(b, W + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ldw + n0 +
                             (j + (lane >> 4)) * 8);
            mma_bf16(acc[j], a[kk], b[0], b[1]);
            mma_bf16(acc[j + 1], a[kk], b[2], b[3]);
        }
}

// acc[16 x 8NT] += D[:, m0:m0+16]^T * X[:, n0:n0+8NT] over the 128 rows of the tile: the weight
// gradient dW[out][in] = sum_rows dY[row][out] * X[row][in]  (both tiles row-major by row)
template <int NT>
__device__ __forceinline__ void gemm_tn(const bf16* D, int ldd, int m0, const bf16* X, int ldx,
                                        int n0, float (&acc)[NT][4], int lane) {
#pragma unroll 2
    for (int k0 = 0; k0 < kBwdRows; k0 += 16) {
        uint32_t a[4];
        ldsm_x4_t(a, D + (k0 + (lane & 7) + ((lane >> 4) & 1) * 8) * ldd + m0 +
                         ((lane >> 3) & 1) * 8);
        if (NT == 1) {
            uint32_t b[2];
            ldsm_x2_t(b, X + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ldx + n0);
            mma_bf16(acc[0], a, b[0], b[1]);
        } else {
#pragma unroll
            for (int j = 0; j + 1 < NT; j += 2) {
                uint32_t b[4];
                ldsm_x4_t(b, X + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ldx + n0 +
                                 (j + (lane >> 4)) * 8);
                mma_bf16(acc[j], a, b[0], b[1]);
                mma_bf16(acc[j + 1], a, b[2], b[3]);
            }
        }
    }
}

// accumulators (16 rows x 64 cols) -> A fragments of the next product (K = 64), optional ReLU
template <bool RELU>
__device__ __forceinline__ void acc_to_a(const float (&acc)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const float(&c0)[4] = acc[2 * kk];
        const float(&c1)[4] = acc[2 * kk + 1];
        auto f = [](float v) { return RELU ? fmaxf(v, 0.f) : v; };
        a[kk][0] = pack_bf2(f(c0[0]), f(c0[1]));
        a[kk][1] = pack_bf2(f(c0[2]), f(c0[3]));
        a[kk][2] = pack_bf2(f(c1[0]), f(c1[1]));
        a[kk][3] = pack_bf2(f(c1[2]), f(c1[3]));
    }
}

// store this warp's 16 x 64 accumulators as bf16 into a row-major tile (row stride kLdH)
template <bool RELU>
__device__ __forceinline__ void store_acc64(const float (&acc)[8][4], bf16* T, int lane) {
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        auto f = [](float v) { return RELU ? fmaxf(v, 0.f) : v; };
        *reinterpret_cast<uint32_t*>(T + gq * kLdH + 8 * j + 2 * tq) =
            pack_bf2(f(acc[j][0]), f(acc[j][1]));
        *reinterpret_cast<uint32_t*>(T + (gq + 8) * kLdH + 8 * j + 2 * tq) =
            pack_bf2(f(acc[j][2]), f(acc[j][3]));
    }
}

// add this warp's weight-gradient accumulators to the fp32 gradient matrix G[rows][ldg]
template <int NT>
__device__ __forceinline__ void flush_dw(const float (&acc)[NT][4], float* __restrict__ G, int ldg,
                                         int m0, int n0, int max_rows, int max_cols, float inv_scale,
                                         int lane) {
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int c = n0 + 8 * j + 2 * tq;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = m0 + gq + 8 * h;
            if (r < max_rows) {
                if (c < max_cols) atomicAdd(G + r * ldg + c, acc[j][2 * h] * inv_scale);
                if (c + 1 < max_cols) atomicAdd(G + r * ldg + c + 1, acc[j][2 * h + 1] * inv_scale);
            }
        }
    }
}
