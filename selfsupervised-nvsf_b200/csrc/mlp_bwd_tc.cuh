// MLP backward on the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory) — the
// Blackwell form of k_mlp_bwd (mlp_bwd.cuh / train.cu), same mathematics: fp16 operands, fp32
// accumulation, a per-launch power-of-two scale on the output gradients, forward activations re-computed per tile.
//
// One CTA = W warpgroups, each an independent chain over 128-row tiles (thread = row = TMEM lane), persistent
// over the tiles of the launch.  Per tile (two hidden layers; one hidden layer drops the steps marked *):
//
//     D1  = X W1^T                         SS   A = X tile (K-major)            B = W1
//     H1  = relu(D1)   -> fp16: smem Ha, packed in place in TMEM, sign bits kept in two registers
//   * D2  = H1 W2^T                        TS   A = H1 from tensor memory       B = W2
//   * H2  = relu(D2)   -> smem Hb, sign bits
//     dWo^T += HL^T dO                     SS   A = HL (MN-major view of the tile), B = dO (MN-major view)   [shared]
//     E   = dO Wo                          SS   A = dO tile (K-major, K = 16)   B = Wo^T image
//     dHL = E * relu'  -> fp16: smem (over HL), packed in place in TMEM
//   * dW2 += dH2^T H1                      SS   A = Hb, B = Ha (both MN-major views)                       [shared]
//   * F   = dH2 W2                         TS   A = dH2 from tensor memory      B = W2^T image
//   * dH1 = F * relu'  -> smem Ha (over H1), packed in place
//     dW1 += dH1^T X                       SS   A = Ha, B = X (MN-major views)                              [shared]
//     dX  = dH1 W1[:, cols]                TS   A = dH1 from tensor memory      B = W1^T image
//
// Facts this rests on, probed on a B200 by tools/umma_probe.cu: (1) a [128 rows][64] fp16 tile with 128-byte
// swizzled rows is a K-major A operand [M = row][K] and, with the instruction descriptor's major bits set and the
// K = 16 step advancing 16 rows (2048 B), an MN-major operand [M or N = feature][K = row] — the weight-gradient
// GEMMs need no transposed copy; (2) an M = 64 accumulator keeps row r in TMEM lane 32 (r / 16) + r % 16;
// (3) the issuing threads of different warpgroups may accumulate into the SAME tensor-memory accumulator: the dW
// accumulators are shared by the CTA (zeroed with tcgen05.st at kernel start) and live there for the whole launch.
// Activations reach the next layer as the A operand from tensor memory (the .ts form); shared memory holds only
// what a weight-gradient GEMM must read by sample (X, H1 / dH1, H2 / dH2, dO) and the weight images.
#pragma once

#include "umma.cuh"

namespace mlptc {
using namespace umma;

constexpr uint32_t kTile64 = 128 * 128;          // one [128 rows][64 halves] swizzled tile
constexpr uint32_t kDoTile = 128 * 32;           // [128 rows][16 halves], un-swizzled core matrices (8 rows x 16 B)
constexpr uint32_t kTile32 = 128 * 64;           // a [128 rows][32 halves] block, un-swizzled core matrices
constexpr uint32_t kChain = 96;                  // TMEM columns of one warpgroup's chain
// The X tile = KIN / 64 swizzled blocks of 64 columns, then (KIN % 64 == 32) one un-swizzled block of 32 columns
// (8 KB instead of a half-empty 16 KB swizzled block: what lets a fourth chain fit for the 32-wide inputs).
__host__ __device__ constexpr uint32_t x_bytes(int kin) { return (uint32_t)(kin / 64) * kTile64 + ((kin % 64) ? kTile32 : 0u); }
// byte offset of 8-half chunk `chunk` of row `row` inside the X tile
__device__ __forceinline__ uint32_t x_off(int kin, uint32_t row, uint32_t chunk) {
    const uint32_t kb = chunk >> 3;
    if (kb < (uint32_t)(kin / 64)) return kb * kTile64 + swz(row, chunk & 7u);
    return (uint32_t)(kin / 64) * kTile64 + (row >> 3) * 512u + (chunk & 3u) * 128u + (row & 7u) * 16u;
}

__host__ __device__ constexpr uint32_t idesc_mn(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint32_t dotile_off(uint32_t row, uint32_t chunk) {   // 8-half chunk `chunk` of row `row`
    return (row >> 3) * 256u + chunk * 128u + (row & 7u) * 16u;
}

// byte sizes / offsets of the packed weight images of one MLP (global scratch and shared memory alike)
template <class T>
struct Img {
    static constexpr uint32_t KB1 = (T::KIN + 63) / 64;                // K blocks of layer 1
    static constexpr uint32_t DXR = (T::DXN + 7) / 8 * 8;              // rows of the W1^T image
    static constexpr uint32_t w1 = 0;                                  // [KB1][64][128 B]
    static constexpr uint32_t w1t = w1 + KB1 * 64 * 128;               // [DXR][128 B]: row r = W1[:, DX0 + r]
    static constexpr uint32_t w2 = w1t + (DXR * 128 + 1023) / 1024 * 1024;   // [64][128 B]  (NHID == 2)
    static constexpr uint32_t w2t = w2 + (T::NHID == 2 ? 64 * 128 : 0);      // [64][128 B]: row n = W2[:, n]
    static constexpr uint32_t wot = w2t + (T::NHID == 2 ? 64 * 128 : 0);     // [64][32 B] un-swizzled: row n = Wo[:, n]
    static constexpr uint32_t total = wot + 2048;
};
constexpr uint32_t kImgSlot = 64 * 1024;   // scratch bytes reserved per MLP

// fp32 master matrix -> fp16 operand image.  B[r][c] = transpose ? src[c * ld + (r0 + r)] : src[(r0 + r) * ld + c]
// for r < rows, c < cols (zero elsewhere: the caller clears the slot).  mode 0: K blocks of 64 columns, each
// [rows_pad][128 B] with the 128-byte swizzle; mode 1: un-swizzled [rows][16].
__global__ void k_pack_tc_image(const float* __restrict__ src, int ld, int rows, int cols, int r0, int transpose,
                                int rows_pad, int mode, unsigned char* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int r = i / cols, c = i - r * cols;
    const float v = transpose ? __ldg(src + (size_t)c * ld + r0 + r) : __ldg(src + (size_t)(r0 + r) * ld + c);
    uint32_t off;
    if (mode == 0) off = (uint32_t)(c >> 6) * (uint32_t)rows_pad * 128u + swz((uint32_t)r, (uint32_t)(c & 63) >> 3) + (c & 7) * 2;
    else off = dotile_off((uint32_t)r, (uint32_t)c >> 3) + (c & 7) * 2;
    *reinterpret_cast<__half*>(dst + off) = __float2half_rn(v);
}

template <class T>
void pack_images(const float* W1, const float* W2, const float* Wo, unsigned char* slot, cudaStream_t s) {
    using I = Img<T>;
    cudaMemsetAsync(slot, 0, I::total, s);
    auto go = [&](const float* src, int ld, int rows, int cols, int r0, int tr, int rows_pad, int mode, uint32_t off) {
        k_pack_tc_image<<<(rows * cols + 255) / 256, 256, 0, s>>>(src, ld, rows, cols, r0, tr, rows_pad, mode, slot + off);
    };
    go(W1, T::KIN, 64, T::KIN, 0, 0, 64, 0, I::w1);                       // W1 [64][KIN]
    go(W1, T::KIN, T::DXN, 64, T::DX0, 1, (int)I::DXR, 0, I::w1t);         // row r = column DX0 + r of W1
    if (T::NHID == 2) {
        go(W2, 64, 64, 64, 0, 0, 64, 0, I::w2);
        go(W2, 64, 64, 64, 0, 1, 64, 0, I::w2t);
    }
    go(Wo, 64, 64, T::OUT_ROWS, 0, 1, 64, 1, I::wot);                     // row n = Wo[0..OUT_ROWS)[n]
}

template <class T, int W>
__host__ __device__ constexpr size_t smem_bytes() {
    return Img<T>::total + (size_t)W * (x_bytes(T::KIN) + (size_t)T::NHID * kTile64 + kDoTile) + 64 + W * 128 + 1024;
}
template <class T, int W>
__host__ __device__ constexpr uint32_t tmem_cols_needed() {
    return W * kChain + T::KIN + (T::NHID == 2 ? 64 : 0) + 16;
}

struct Grads {
    float* w1;  // [64][LDG1]
    float* w2;  // [64][64]
    float* wo;  // [OUT_ROWS][64]
};

// accumulator row (64 columns at src) -> fp16 row: RELU (sets the sign bits) or masked by the sign bits;
// written to the smem tile and / or packed into TMEM at dst (dst == src allowed)
template <bool RELU, bool TO_TMEM>
__device__ __forceinline__ void row64(uint32_t src, uint32_t dst, unsigned char* tile, uint32_t t, uint32_t (&bits)[2]) {
    if (RELU) bits[0] = bits[1] = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t v[16], o[8];
        tmem_ld16(src + q * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float f = __uint_as_float(v[i]);
            if (RELU) {
                if (f > 0.f) bits[q >> 1] |= 1u << ((q & 1) * 16 + i);
                else v[i] = 0u;
            } else if (!((bits[q >> 1] >> ((q & 1) * 16 + i)) & 1u)) {
                v[i] = 0u;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = pack_half2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
        *reinterpret_cast<uint4*>(tile + swz(t, 2 * q)) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(tile + swz(t, 2 * q + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
        if (TO_TMEM) tmem_st8(dst + q * 8, o);
    }
    if (TO_TMEM) tmem_st_wait();
}

template <class T, int W>
__global__ void __launch_bounds__(W * 128, 1)
k_mlp_bwd_tc(const typename T::Args A, const unsigned char* __restrict__ wimg, size_t n, Grads G,
             const float* __restrict__ scale2) {
    using I = Img<T>;
    constexpr int KIN = T::KIN, NHID = T::NHID;
    constexpr uint32_t KB1 = I::KB1;
    constexpr uint32_t kXBytes = x_bytes(KIN), KB64 = KIN / 64;
    constexpr bool kStageInX = T::DXN > 32;                   // dX staging: the X tile if the chunk needs 32 KB, else Ha
    constexpr uint32_t kWgBytes = kXBytes + NHID * kTile64 + kDoTile;
    constexpr uint32_t kOffBar = I::total + W * kWgBytes;
    static_assert(I::total % 1024 == 0 && kWgBytes % 1024 == 0, "swizzled tiles start on 1024-byte boundaries");
    static_assert(tmem_cols_needed<T, W>() <= 512, "tensor memory budget");
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    const uint32_t tid = threadIdx.x, wg = tid >> 7, t = tid & 127u, wq = (tid >> 5) & 3u;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBar + 8 * W);
    unsigned char* flags = sm + kOffBar + 64 + wg * 128;      // per row of the tile: carries gradient

    for (uint32_t i = tid; i < I::total / 16; i += W * 128)
        reinterpret_cast<uint4*>(sm)[i] = __ldg(reinterpret_cast<const uint4*>(wimg) + i);
    if (tid < (uint32_t)W) mbar_init(base + kOffBar + 8 * tid, 1);
    if (tid < 32) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tc = tmem + wg * kChain;                   // this warpgroup's chain columns
    const uint32_t tl = tc + ((wq * 32u) << 16);              // ... of this warp's lanes
    // shared weight-gradient accumulators (M = 64: lanes 32 q + 0..15 hold rows 16 q + 0..15)
    const uint32_t acc_w1 = tmem + W * kChain, acc_w2 = acc_w1 + KIN, acc_wo = acc_w2 + (NHID == 2 ? 64 : 0);
    constexpr uint32_t kAccCols = KIN + (NHID == 2 ? 64 : 0) + 16;
    if (wg == 0) {   // zero them: every later MMA accumulates
        const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t c = 0; c < kAccCols; c += 8) tmem_st8(acc_w1 + ((wq * 32u) << 16) + c, z);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    unsigned char* xg = sm + I::total + wg * kWgBytes;        // X tile(s)
    unsigned char* hag = xg + kXBytes;                        // Ha: H1, later dH1
    unsigned char* hbg = hag + kTile64;                       // Hb: H2, later dH2 (NHID == 2)
    unsigned char* dog = hag + NHID * kTile64;                // dO
    const uint32_t xs = base + I::total + wg * kWgBytes, has = xs + kXBytes, hbs = has + kTile64,
                   dos = has + NHID * kTile64;
    const uint32_t bar = base + kOffBar + 8 * wg;
    const float scale = __ldg(scale2), inv_scale = __ldg(scale2 + 1);
    constexpr uint32_t kId64 = umma_idesc(128, 64);
    uint32_t phase = 0;
    const uint32_t hls = NHID == 2 ? hbs : has;

    const size_t n_tiles = (n + 127) / 128;
    for (size_t tile = (size_t)blockIdx.x * W + wg; tile < n_tiles; tile += (size_t)gridDim.x * W) {
        // ---- the scaled output gradients of the tile (loaded by the warpgroup, coalesced) --------------------
        const size_t row0 = tile * 128;
        T::load_do(A, row0, n, scale, dog, t);
        wg_barrier(wg);
        const uint4 o0 = *reinterpret_cast<const uint4*>(dog + dotile_off(t, 0));
        const uint4 o1 = *reinterpret_cast<const uint4*>(dog + dotile_off(t, 1));
        const bool live = ((o0.x | o0.y | o0.z | o0.w | o1.x | o1.y | o1.z | o1.w) & 0x7fff7fffu) != 0;
        flags[t] = live ? 1 : 0;
        tc_fence_before();
        if (!wg_any(wg, live)) {      // nothing flows back through this tile
            T::store_dead(A, row0, n, t);
            continue;
        }
        T::load_x(A, row0, n, xg, t, wg, hag);   // (heads: Ha + Hb, still unused, are scratch for the per-ray encodings:
                                                 //  up to 128 rays x 144 B when every row is its own ray)
        {   // the next tile's rows start travelling towards the L2 while this one is computed
            const size_t nrow0 = row0 + (size_t)gridDim.x * W * 128;
            if (nrow0 < n) T::prefetch(A, nrow0, n, t);
        }
        fence_async_smem();
        tc_fence_before();
        wg_barrier(wg);
        if (t == 0) {   // D1 [0,64) = X W1^T
            tc_fence_after();
#pragma unroll
            for (uint32_t k = 0; k < (uint32_t)KIN / 16; ++k) {
                const uint64_t a = (k >> 2) < KB64 ? umma_desc(xs + (k >> 2) * kTile64 + (k & 3) * 32)
                                                   : desc_noswz(xs + KB64 * kTile64 + (k & 3) * 256, 128u, 512u);
                umma_f16(tc, a, umma_desc(base + I::w1 + (k >> 2) * (64 * 128) + (k & 3) * 32), kId64, k);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1u;
        tc_fence_after();
        uint32_t m1[2], m2[2] = {0u, 0u};
        row64<true, NHID == 2>(tl, tl, hag, t, m1);          // H1 -> Ha (+ packed in place [0,32))
        if (NHID == 2) {
            tc_fence_before();
            wg_barrier(wg);
            if (t == 0) {   // D2 [32,96) = H1 W2^T, A from tensor memory
                tc_fence_after();
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                    umma_f16_ts(tc + 32, tc + 8 * k, umma_desc(base + I::w2 + k * 32), kId64, k);
                umma_commit(bar);
            }
            mbar_wait(bar, phase); phase ^= 1u;
            tc_fence_after();
            row64<true, false>(tl + 32, 0, hbg, t, m2);      // H2 -> Hb
        }
        fence_async_smem();
        tc_fence_before();
        wg_barrier(wg);
        if (t == 0) {
            tc_fence_after();
            // dWo^T [64 hidden][16 out] += HL^T dO : both operands read by sample (MN-major views), K = 128 rows
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k)
                umma_f16(acc_wo, umma_desc(hls + k * 2048), desc_noswz(dos + k * 512, 256u, 128u),
                         idesc_mn(64, 16, 1, 1), 1u);
            // E [0,64) = dO Wo : K = 16
            umma_f16(tc, desc_noswz(dos, 128u, 256u), desc_noswz(base + I::wot, 128u, 256u), kId64, 0u);
            umma_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1u;
        tc_fence_after();
        if (NHID == 2) {
            row64<false, true>(tl, tl, hbg, t, m2);          // dH2 -> Hb (over H2), packed in place [0,32)
            fence_async_smem();
            tc_fence_before();
            wg_barrier(wg);
            if (t == 0) {
                tc_fence_after();
                // dW2 [64][64] += dH2^T H1
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k)
                    umma_f16(acc_w2, umma_desc(hbs + k * 2048), umma_desc(has + k * 2048), idesc_mn(64, 64, 1, 1), 1u);
                // F [32,96) = dH2 W2 : A from tensor memory, B = W2^T image
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                    umma_f16_ts(tc + 32, tc + 8 * k, umma_desc(base + I::w2t + k * 32), kId64, k);
                umma_commit(bar);
            }
            mbar_wait(bar, phase); phase ^= 1u;
            tc_fence_after();
            row64<false, true>(tl + 32, tl + 32, hag, t, m1);   // dH1 -> Ha (over H1), packed in place [32,64)
        } else {
            row64<false, true>(tl, tl, hag, t, m1);          // dH1 -> Ha (over H1), packed in place [0,32)
        }
        fence_async_smem();
        tc_fence_before();
        wg_barrier(wg);
        constexpr uint32_t kDh1 = NHID == 2 ? 32 : 0;         // packed dH1
        constexpr uint32_t kDx = NHID == 2 ? 64 : 32;         // dX chunk
        constexpr uint32_t kDxChunk = T::DXN > 64 ? 64 : I::DXR;   // columns per dX MMA batch
        constexpr int kDxChunks = (T::DXN + 63) / 64;
#pragma unroll
        for (int c = 0; c < kDxChunks; ++c) {
            if (t == 0) {
                tc_fence_after();
                if (c == 0) {   // dW1 [64][KIN] += dH1^T X, 64 (or the remaining) columns of X at a time
#pragma unroll
                    for (uint32_t nb = 0; nb < KB1; ++nb) {
                        constexpr uint32_t last = KIN - (KB1 - 1) * 64;
                        const uint32_t nn = nb + 1 < KB1 ? 64u : last;
#pragma unroll
                        for (uint32_t k = 0; k < 8; ++k) {
                            const uint64_t b = nb < KB64 ? umma_desc(xs + nb * kTile64 + k * 2048)
                                                         : desc_noswz(xs + KB64 * kTile64 + k * 1024, 512u, 128u);
                            umma_f16(acc_w1 + nb * 64, umma_desc(has + k * 2048), b, idesc_mn(64, (int)nn, 1, 1), 1u);
                        }
                    }
                }
                // dX chunk = dH1 W1[:, DX0 + 64 c ...] : A from tensor memory, B = W1^T image rows
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                    umma_f16_ts(tc + kDx, tc + kDh1 + 8 * k, umma_desc(base + I::w1t + c * (64 * 128) + k * 32),
                                umma_idesc(128, (int)kDxChunk), k);
                umma_commit(bar);
            }
            mbar_wait(bar, phase); phase ^= 1u;
            tc_fence_after();
            // this thread's row of the chunk -> staging (the X or Ha tile: their last reader, the dW1 batch, has completed),
            // 16-byte pieces XOR-swizzled by row so that neither side has bank conflicts; the warpgroup then
            // writes the rows out coalesced
            constexpr uint32_t kPieces = kDxChunk / 4;
            unsigned char* stage = kStageInX ? xg : hag;
#pragma unroll
            for (uint32_t q = 0; q < kDxChunk / 16; ++q) {
                uint32_t v[16];
                tmem_ld16(tl + kDx + q * 16, v);
                tmem_ld_wait();
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j)
                    *reinterpret_cast<float4*>(stage + t * (kDxChunk * 4) + (((q * 4 + j) ^ (t & (kPieces - 1))) << 4)) =
                        make_float4(__uint_as_float(v[4 * j]) * inv_scale, __uint_as_float(v[4 * j + 1]) * inv_scale,
                                    __uint_as_float(v[4 * j + 2]) * inv_scale, __uint_as_float(v[4 * j + 3]) * inv_scale);
            }
            tc_fence_before();
            wg_barrier(wg);
            T::store_dx(A, row0, n, c, stage, flags, t);
            if (c + 1 < kDxChunks) wg_barrier(wg);   // staging and the chunk's TMEM columns are free again
        }
        // the next tile's first MMA batch is issued behind warpgroup barriers every thread reaches after these reads
    }
    // ---- weight gradients: tensor memory -> fp32 global (atomics: one partial per CTA) -------------------
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    {
        const uint32_t lane = tid & 31u, warp = tid >> 5;
        const uint32_t q = warp & 3u;                 // TMEM lane quadrant this warp may read
        const uint32_t part = warp >> 2, parts = W;   // the warpgroups split the columns
        const bool has_row = lane < 16;
        const uint32_t m = 16 * q + lane;             // hidden unit (row of dW1 / dW2, column of dWo)
        for (uint32_t c0 = part * 16; c0 < kAccCols; c0 += parts * 16) {
            uint32_t v[16];
            tmem_ld16(acc_w1 + ((q * 32u) << 16) + c0, v);
            tmem_ld_wait();
            if (!has_row) continue;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float g = __uint_as_float(v[i]) * inv_scale;
                const uint32_t c = c0 + i;
                if (c < (uint32_t)KIN) {
                    if (c < (uint32_t)T::LDG1) atomicAdd(G.w1 + m * T::LDG1 + c, g);
                } else if (NHID == 2 && c < (uint32_t)KIN + 64) {
                    atomicAdd(G.w2 + m * 64 + (c - KIN), g);
                } else {
                    const uint32_t o = c - KIN - (NHID == 2 ? 64 : 0);
                    if (o < (uint32_t)T::OUT_ROWS) atomicAdd(G.wo + o * 64 + m, g);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace mlptc
