// nvsf_b200 — training path of the uniform-sample renderer: forward that keeps its per-sample
// intermediates, and the backward pass down to the fp32 master parameters (sm_100a).
//
// Replaces what autograd does for the reference's train_step (reference nvsf/nerf/trainer.py:193-200,
// 491-499 -> renderer_dynamic.py:109-265 -> network_dynamic.py:213-332): the backward of the
// cumprod compositing, of the masked colour heads, of trunc_exp (activation.py:14-19), of the
// sigma / flow MLPs (tcnn FullyFusedMLP and nn.Linear backward), of grid_sample (planes, including
// the coordinate gradient that carries the only gradient of the flow field) and the atomicAdd
// scatter of the tcnn hash-grid backward.  Gradient semantics follow the reference exactly: the
// warped dynamic-HASH queries are evaluated under no_grad (network_dynamic.py:245-249, 261-265),
// the warped PLANE queries are differentiable, a missing neighbour frame falls back to the
// un-warped feature (:238-239), the colour mask w > 1e-4 is not differentiated.
//
// B200-first structure (one launch sequence per chunk of rays, all intermediates fp32 or bf16):
//   k_composite_bwd   one warp per ray: two sweeps over the ray (total of dL/dw * w, then the
//                     suffix form of the cumprod backward) -> dL/d(sigma logit)
//   k_mlp_bwd<Heads>  128-row tiles, fp16 mma.sync (power-of-two gradient scale per launch): re-computes the hidden activations from the
//                     kept geo features + direction encoding, back-propagates w * dL/dimage *
//                     c(1-c), accumulates dW in registers across the tiles of a persistent CTA
//   k_mlp_bwd<Sigma>  same kernel template on the kept 120 sigma-net inputs -> dL/dfeatures
//   k_encode_bwd      one thread per sample: re-gathers the plane texels (product rule), scatters
//                     the plane gradients with WARP-AGGREGATED vector reds (consecutive samples
//                     of a ray fall into runs of equal texels: segmented shuffle reduction, one
//                     red.v4 per run instead of one per lane), the hash gradients with red.v4 /
//                     red.f32 into the time-collapsed gradient tables, and emits dL/dflow
//   k_mlp_bwd<Flow>   flow MLP backward -> dL/d(flow-grid features)
//   k_flowgrid_bwd    red.v2 scatter into the time-collapsed flow-grid gradient
//   k_expand_*        collapsed gradient tables -> the reference parameter layouts (the transpose
//                     of the time collapse of field.cu; linear, so exact)
#include <algorithm>
#include <string>

#include "mlp_bwd.cuh"
#include "mlp_bwd_tc.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// fp16 weight images for the backward kernels (row strides of mlp_bwd.cuh)
// ------------------------------------------------------------------------------------------------
constexpr int kBLdK32 = 40, kBLdK96 = 104, kBLdK128 = 136;
constexpr int kB_FlowW1 = 0;                                  // [64][40]
constexpr int kB_FlowW2 = kB_FlowW1 + kHidden * kBLdK32;      // [64][72]
constexpr int kB_FlowW3 = kB_FlowW2 + kHidden * kLdH;         // [16][72] rows 6..15 zero
constexpr int kB_SigW1 = kB_FlowW3 + 16 * kLdH;               // [64][136]
constexpr int kB_SigW2 = kB_SigW1 + kHidden * kBLdK128;       // [16][72]
constexpr int kB_Head = kB_SigW2 + 16 * kLdH;                 // 2 slots
constexpr int kB_HeadW1 = 0;                                  // [64][104] (camera: [64][40])
constexpr int kB_HeadW2 = kB_HeadW1 + kHidden * kBLdK96;      // [64][72]
constexpr int kB_HeadW3 = kB_HeadW2 + kHidden * kLdH;         // [16][72]
constexpr int kB_HeadHalves = kB_HeadW3 + 16 * kLdH;
constexpr int kB_Total = kB_Head + 2 * kB_HeadHalves;

__global__ void k_pack_matrix_bf16(const float* __restrict__ src, int src_ld, int rows, int cols,
                                   bf16* __restrict__ dst, int dst_ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int r = i / cols, c = i - r * cols;
    dst[r * dst_ld + c] = __float2half_rn(__ldg(src + r * src_ld + c));
}

// ------------------------------------------------------------------------------------------------
// the generic MLP backward tile kernel
// ------------------------------------------------------------------------------------------------
struct MlpGrads {
    float* w1;  // [64][LDG1]
    float* w2;  // [64][64]   (NHID == 2)
    float* wo;  // [OUT_ROWS][64]
};

template <class T>
constexpr size_t mlp_bwd_smem() {
    constexpr int ldx = T::KIN + 8;
    size_t halves = (size_t)kHidden * ldx + (T::NHID == 2 ? (size_t)kHidden * kLdH : 0) + 16 * kLdH;
    halves += (size_t)kBwdRows * ldx + (size_t)kBwdRows * kLdH * (2 * T::NHID) + (size_t)kBwdRows * kLdD;
    return halves * sizeof(bf16);
}

template <class T>
__global__ void __launch_bounds__(kBwdWarps * 32)
k_mlp_bwd(const typename T::Args A, const bf16* __restrict__ w1g, const bf16* __restrict__ w2g,
          const bf16* __restrict__ wog, size_t n, MlpGrads G, const float* __restrict__ scale2) {
    constexpr int KIN = T::KIN, NHID = T::NHID, LDX = KIN + 8, KT = KIN / 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bf16* W1 = reinterpret_cast<bf16*>(smem_raw);
    bf16* W2 = W1 + kHidden * LDX;
    bf16* Wo = W2 + (NHID == 2 ? kHidden * kLdH : 0);
    bf16* Xs = Wo + 16 * kLdH;
    bf16* H1s = Xs + kBwdRows * LDX;
    bf16* H2s = H1s + kBwdRows * kLdH;                        // NHID == 2 only
    bf16* dH1s = H1s + kBwdRows * kLdH * NHID;
    bf16* dH2s = dH1s + kBwdRows * kLdH;                      // NHID == 2 only
    bf16* Ds = dH1s + kBwdRows * kLdH * NHID;
    bf16* HLs = NHID == 2 ? H2s : H1s;                        // last hidden activations
    bf16* dHLs = NHID == 2 ? dH2s : dH1s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const float scale = __ldg(scale2), inv_scale = __ldg(scale2 + 1);  // power of two and its inverse
    block_copy16(W1, w1g, kHidden * LDX * 2 / 16, tid, kBwdWarps * 32);
    if (NHID == 2) block_copy16(W2, w2g, kHidden * kLdH * 2 / 16, tid, kBwdWarps * 32);
    block_copy16(Wo, wog, 16 * kLdH * 2 / 16, tid, kBwdWarps * 32);
    // hidden tiles start finite (skipped warps leave their rows untouched; 0 * NaN would poison dW)
    for (int i = tid; i < kBwdRows * kLdH * NHID * 2 / 8; i += kBwdWarps * 32)
        reinterpret_cast<uint4*>(H1s)[i] = make_uint4(0, 0, 0, 0);

    // weight-gradient accumulators, live across all tiles of this CTA
    const int mt = warp & 3, nh = warp >> 2;
    float accW1[KT][4], accW2[4][4], accWo[1][4];
    zero1<KT>(accW1);
    zero1<4>(accW2);
    zero1<1>(accWo);
    __syncthreads();

    const size_t n_tiles = (n + kBwdRows - 1) / kBwdRows;
    typename T::Pref pf;
    if (blockIdx.x < n_tiles) T::load(A, (size_t)blockIdx.x * kBwdRows, n, pf, tid);
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t row0 = tile * kBwdRows;
        T::fill(A, row0, n, Xs, Ds, tid, scale, pf);
        __syncthreads();
        // the next tile's global rows travel while this one is computed (one CTA of 8 warps per SM
        // hides no latency on its own: ncu had 24 % of the head kernel's samples on these loads)
        if (tile + gridDim.x < n_tiles) T::load(A, (tile + gridDim.x) * kBwdRows, n, pf, tid);

        // ---- per-warp chain on rows [16 warp, 16 warp + 16) ----
        const int r0 = warp * 16;
        bool active;
        {
            const uint4* d = reinterpret_cast<const uint4*>(Ds + (r0 + (lane >> 1)) * kLdD + (lane & 1) * 8);
            const uint4 v = *d;
            // -0.0 (0x8000) also counts as zero
            active = __any_sync(0xffffffffu, ((v.x | v.y | v.z | v.w) & 0x7fff7fffu) != 0);
        }
        if (!active) {
            for (int i = lane; i < 16 * kLdH / 8; i += 32) {
                reinterpret_cast<uint4*>(dH1s + r0 * kLdH)[i] = make_uint4(0, 0, 0, 0);
                if (NHID == 2) reinterpret_cast<uint4*>(dH2s + r0 * kLdH)[i] = make_uint4(0, 0, 0, 0);
            }
            T::sink_zero(A, row0 + r0, n, lane);
        } else {
            float h1[8][4], h2[8][4], d1[8][4];
            {
                uint32_t ax[KT][4];
                load_a16<KT>(Xs + r0 * LDX, LDX, ax, lane);
                zero1<8>(h1);
                gemm_nt<KT, 8>(ax, W1, LDX, h1, lane);
            }
            store_acc64<true>(h1, H1s + r0 * kLdH, lane);
            if (NHID == 2) {
                uint32_t a1[4][4];
                acc_to_a<true>(h1, a1);
                zero1<8>(h2);
                gemm_nt<4, 8>(a1, W2, kLdH, h2, lane);
                store_acc64<true>(h2, H2s + r0 * kLdH, lane);
            }
            // dH_last = (dOut * Wo) * relu'
            {
                uint32_t ad[1][4];
                load_a16<1>(Ds + r0 * kLdD, kLdD, ad, lane);
                zero1<8>(d1);
                gemm_nn<1, 8>(ad, Wo, kLdH, 0, d1, lane);
            }
            if (NHID == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) d1[j][i] = h2[j][i] > 0.f ? d1[j][i] : 0.f;
                store_acc64<false>(d1, dH2s + r0 * kLdH, lane);
                uint32_t a2[4][4];
                acc_to_a<false>(d1, a2);
                zero1<8>(d1);
                gemm_nn<4, 8>(a2, W2, kLdH, 0, d1, lane);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) d1[j][i] = h1[j][i] > 0.f ? d1[j][i] : 0.f;
            store_acc64<false>(d1, dH1s + r0 * kLdH, lane);
            // dX = dH1 * W1 on the columns the caller needs
            uint32_t a1[4][4];
            acc_to_a<false>(d1, a1);
#pragma unroll
            for (int c = 0; c < T::DX_CHUNKS; ++c) {
                float dx[T::DX_NT][4];
                zero1<T::DX_NT>(dx);
                gemm_nn<4, T::DX_NT>(a1, W1, LDX, T::DX_N0 + c * T::DX_NT * 8, dx, lane);
#pragma unroll
                for (int j = 0; j < T::DX_NT; ++j) {
                    const int col = T::DX_N0 + (c * T::DX_NT + j) * 8 + 2 * tq;
                    T::sink(A, row0 + r0 + gq, n, col, dx[j][0] * inv_scale, dx[j][1] * inv_scale);
                    T::sink(A, row0 + r0 + gq + 8, n, col, dx[j][2] * inv_scale, dx[j][3] * inv_scale);
                }
            }
        }
        __syncthreads();

        // ---- weight gradients over the 128 rows of the tile ----
        gemm_tn<KT>(dH1s, kLdH, 16 * mt, Xs, LDX, nh * (KIN / 2), accW1, lane);
        if (NHID == 2) gemm_tn<4>(dH2s, kLdH, 16 * mt, H1s, kLdH, nh * 32, accW2, lane);
        gemm_tn<1>(Ds, kLdD, 0, HLs, kLdH, 8 * warp, accWo, lane);
        __syncthreads();
    }
    (void)dHLs;
    flush_dw<KT>(accW1, G.w1, T::LDG1, 16 * mt, nh * (KIN / 2), kHidden, T::LDG1, inv_scale, lane);
    if (NHID == 2) flush_dw<4>(accW2, G.w2, kHidden, 16 * mt, nh * 32, kHidden, kHidden, inv_scale, lane);
    flush_dw<1>(accWo, G.wo, kHidden, 0, 8 * warp, T::OUT_ROWS, kHidden, inv_scale, lane);
}

// ---- sigma net: X = kept features, dOut = (d logit, d geo), dX = d features ----------------------
struct SigmaT {
    static constexpr int KIN = 128, NHID = 1, LDG1 = 128, OUT_ROWS = 16;
    static constexpr int DX_N0 = 0, DX_NT = 8, DX_CHUNKS = 2;
    struct Args {
        const __half* feats;   // [n,128]
        const float* dgeo16;   // [n,16]
        float* dfeat;          // [n,128]
    };
    struct Pref {};
    static __device__ __forceinline__ void load(const Args&, size_t, size_t, Pref&, int) {}
    static __device__ __forceinline__ void fill(const Args& A, size_t row0, size_t n, bf16* Xs,
                                                bf16* Ds, int tid, float scale, const Pref&) {
        constexpr int LDX = KIN + 8;
#pragma unroll 4
        for (int i = tid; i < kBwdRows * 16; i += kBwdWarps * 32) {
            const int r = i >> 4, c = i & 15;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (row0 + r < n) v = __ldcs(reinterpret_cast<const uint4*>(A.feats + (row0 + r) * 128) + c);
            *reinterpret_cast<uint4*>(Xs + r * LDX + c * 8) = v;
        }
        for (int i = tid; i < kBwdRows * 4; i += kBwdWarps * 32) {
            const int r = i >> 2, c = i & 3;
            uint2 o = make_uint2(0, 0);
            if (row0 + r < n) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(A.dgeo16 + (row0 + r) * 16) + c);
                o.x = pack_bf2(scale * g.x, scale * g.y); o.y = pack_bf2(scale * g.z, scale * g.w);
            }
            *reinterpret_cast<uint2*>(Ds + r * kLdD + c * 4) = o;
        }
    }
    static __device__ __forceinline__ void sink(const Args& A, size_t row, size_t n, int col, float a,
                                                float b) {
        if (row < n && col < 120) *reinterpret_cast<float2*>(A.dfeat + row * 128 + col) = make_float2(a, b);
    }
    // rows without gradient are skipped by k_encode_bwd (it tests dgeo16), nothing to write
    static __device__ __forceinline__ void sink_zero(const Args&, size_t, size_t, int) {}
};

// ---- flow MLP: X = kept flow-grid features, dOut = d flow, dX = d flow-grid features ------------
struct FlowT {
    static constexpr int KIN = 32, NHID = 2, LDG1 = 32, OUT_ROWS = 6;
    static constexpr int DX_N0 = 0, DX_NT = 4, DX_CHUNKS = 1;
    struct Args {
        const __half* flowfeat;  // [n,32]
        const float* dflow;      // [n,8]  (6 used, 2 zero)
        float* dflowfeat;        // [n,32]
    };
    struct Pref {};
    static __device__ __forceinline__ void load(const Args&, size_t, size_t, Pref&, int) {}
    static __device__ __forceinline__ void fill(const Args& A, size_t row0, size_t n, bf16* Xs,
                                                bf16* Ds, int tid, float scale, const Pref&) {
        constexpr int LDX = KIN + 8;
        for (int i = tid; i < kBwdRows * 4; i += kBwdWarps * 32) {
            const int r = i >> 2, c = i & 3;
            uint4 v = make_uint4(0, 0, 0, 0);
            uint2 o = make_uint2(0, 0);
            if (row0 + r < n) {
                v = __ldcs(reinterpret_cast<const uint4*>(A.flowfeat + (row0 + r) * 32) + c);
                if (c < 2) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(A.dflow + (row0 + r) * 8) + c);
                    o.x = pack_bf2(scale * g.x, scale * g.y); o.y = pack_bf2(scale * g.z, scale * g.w);
                }
            }
            *reinterpret_cast<uint4*>(Xs + r * LDX + c * 8) = v;
            *reinterpret_cast<uint2*>(Ds + r * kLdD + c * 4) = o;
        }
    }
    static __device__ __forceinline__ void sink(const Args& A, size_t row, size_t n, int col, float a,
                                                float b) {
        if (row < n) *reinterpret_cast<float2*>(A.dflowfeat + row * 32 + col) = make_float2(a, b);
    }
    static __device__ __forceinline__ void sink_zero(const Args&, size_t, size_t, int) {}
};

// ---- colour heads: X = [direction encoding | geo features], dOut from the kept colours --------
__device__ __forceinline__ void sh4_eval(float dx, float dy, float dz, float (&v)[16]) {
    const float x = ((dx + 1.0f) * 0.5f) * 2.0f - 1.0f, y = ((dy + 1.0f) * 0.5f) * 2.0f - 1.0f,
                z = ((dz + 1.0f) * 0.5f) * 2.0f - 1.0f;
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    v[0] = 0.28209479177387814f;
    v[1] = -0.48860251190291987f * y;
    v[2] = 0.48860251190291987f * z;
    v[3] = -0.48860251190291987f * x;
    v[4] = 1.0925484305920792f * xy;
    v[5] = -1.0925484305920792f * yz;
    v[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    v[7] = -1.0925484305920792f * xz;
    v[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    v[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    v[10] = 2.8906114426405538f * xy * z;
    v[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    v[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    v[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    v[14] = 1.4453057213202769f * z * (x2 - y2);
    v[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

template <bool LIDAR>
struct HeadT {
    static constexpr int KIN = LIDAR ? 96 : 32, NHID = 2, LDG1 = KIN, OUT_ROWS = 16;
    static constexpr int NDIR = LIDAR ? 72 : 16;
    static constexpr int DX_N0 = NDIR, DX_NT = 2, DX_CHUNKS = 1;
    struct Args {
        const __half* geo;      // [n,16]  col 0 = sigma logit, 1..15 = geo features
        const float* rgbs;      // [n,4]   kept colours (0 where masked out)
        const float* weights;   // [n]
        const float* g_image;   // [N,NCH]
        const float* rays_d;    // [N,3]
        float* dgeo16;          // [n,16]  cols 1..15 written (or accumulated)
        uint32_t S;
        int net;                // lidar: 0 = intensity_net (image channel 1), 1 = raydrop_net (channel 0)
        int accumulate;
    };
    struct Pref {           // the global rows of one tile row, loaded one tile ahead
        uint4 g0, g1;       // geo [16] halves
        float w;            // weights[g]
        float4 cc;          // kept colours
        float gi[3];        // dL/dimage of the row's ray
        float d[3];         // ray direction (camera: SH input)
    };
    static __device__ __forceinline__ void load(const Args& A, size_t row0, size_t n, Pref& pf, int tid) {
        const int r = tid >> 1, half = tid & 1;
        const size_t g = row0 + r;
        if (g >= n) return;
        const size_t ray = g / A.S;
        pf.g0 = __ldg(reinterpret_cast<const uint4*>(A.geo + g * 16));
        pf.g1 = __ldg(reinterpret_cast<const uint4*>(A.geo + g * 16) + 1);
        if (!LIDAR) {
            pf.d[0] = __ldg(A.rays_d + ray * 3); pf.d[1] = __ldg(A.rays_d + ray * 3 + 1);
            pf.d[2] = __ldg(A.rays_d + ray * 3 + 2);
        }
        if (half == 0) {
            pf.w = __ldg(A.weights + g);
            pf.cc = __ldg(reinterpret_cast<const float4*>(A.rgbs + g * 4));
            if (LIDAR) {
                pf.gi[0] = __ldg(A.g_image + ray * 2 + (A.net == 0 ? 1 : 0));
            } else {
                pf.gi[0] = __ldg(A.g_image + ray * 3); pf.gi[1] = __ldg(A.g_image + ray * 3 + 1);
                pf.gi[2] = __ldg(A.g_image + ray * 3 + 2);
            }
        }
    }
    static __device__ __forceinline__ void fill(const Args& A, size_t row0, size_t n, bf16* Xs,
                                                bf16* Ds, int tid, float scale, const Pref& pf) {
        constexpr int LDX = KIN + 8;
        const int r = tid >> 1, half = tid & 1;
        const size_t g = row0 + r;
        bf16* xr = Xs + r * LDX;
        if (LIDAR) {
            // tcnn Frequency (12 octaves) of (d+1)/2 is constant along a ray: evaluate the 72 columns
            // once per ray of the tile (36 threads x one sin/cos pair) into the tile row of the ray's
            // first sample; the other rows copy it below.  (Per-row evaluation was 60 % of this
            // kernel's instructions, profiles/r01_train_bwd_v2_ncu_full.txt.)
            const size_t last = (row0 + kBwdRows < n ? row0 + kBwdRows : n) - 1;
            const size_t ray_a = row0 / A.S;
            const int nr = (int)(last / A.S - ray_a) + 1;
            for (int t = tid; t < nr * 36; t += kBwdWarps * 32) {
                const int rl = t / 36, j = 2 * (t - rl * 36);
                const size_t ry = ray_a + rl;
                const size_t first = ry * A.S > row0 ? ry * A.S : row0;
                const int dim = j / 24, oct = (j >> 1) % 12;
                const float v = (__ldg(A.rays_d + ry * 3 + dim) + 1.0f) * 0.5f;
                const float a = scalbnf(v, oct);
                *reinterpret_cast<uint32_t*>(Xs + (first - row0) * LDX + j) = pack_bf2(sinpif(a), sinpif(a + 0.5f));
            }
        }
        if (g >= n) {
            for (int c = half * (KIN / 2); c < (half + 1) * (KIN / 2); c += 8)
                *reinterpret_cast<uint4*>(xr + c) = make_uint4(0, 0, 0, 0);
            if (half == 0) {
                *reinterpret_cast<uint4*>(Ds + r * kLdD) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(Ds + r * kLdD + 8) = make_uint4(0, 0, 0, 0);
            }
            // (no early return on the LiDAR path: the barrier below must be reached by every thread through the
            //  SAME instruction — a tile whose last row falls inside a warp used to arrive at two different
            //  __syncthreads() from one divergent warp and never left)
            if (!LIDAR) return;
        }
        const bool inb = g < n;
        const size_t ray = inb ? g / A.S : 0;
        if (inb) {
        // geo features: 16 halves; cols NDIR + (0..14) = geo[1..15], col NDIR+15 = 1 (padding)
        const __half* gh0 = reinterpret_cast<const __half*>(&pf.g0);
        const __half* gh1 = reinterpret_cast<const __half*>(&pf.g1);
        if (!LIDAR) {
            float sh[16];
            sh4_eval(pf.d[0], pf.d[1], pf.d[2], sh);
#pragma unroll
            for (int j = 0; j < 8; j += 2)
                *reinterpret_cast<uint32_t*>(xr + half * 8 + j) =
                    pack_bf2(half ? sh[8 + j] : sh[j], half ? sh[9 + j] : sh[j + 1]);
        }
        if (half == 0) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                const float a = __half2float(j + 1 < 8 ? gh0[j + 1] : gh1[j + 1 - 8]);
                const float b = __half2float(j + 2 < 8 ? gh0[j + 2] : gh1[j + 2 - 8]);
                *reinterpret_cast<uint32_t*>(xr + NDIR + j) = pack_bf2(a, b);
            }
            // output gradient: w * dL/dimage[ch] * c (1 - c) on the samples that passed the mask
            const float w = pf.w, ws = scale * w;
            uint4 o0 = make_uint4(0, 0, 0, 0);
            if (w > 1e-4f) {
                const float4 cc = pf.cc;
                const float2 c01 = make_float2(cc.x, cc.y), c23 = make_float2(cc.z, cc.w);
                if (LIDAR) {
                    const int ch = A.net == 0 ? 1 : 0;
                    const float c = ch ? c01.y : c01.x;
                    o0.x = pack_bf2(ws * pf.gi[0] * c * (1.f - c), 0.f);
                } else {
                    o0.x = pack_bf2(ws * pf.gi[0] * c01.x * (1.f - c01.x),
                                    ws * pf.gi[1] * c01.y * (1.f - c01.y));
                    o0.y = pack_bf2(ws * pf.gi[2] * c23.x * (1.f - c23.x), 0.f);
                }
            }
            *reinterpret_cast<uint4*>(Ds + r * kLdD) = o0;
            *reinterpret_cast<uint4*>(Ds + r * kLdD + 8) = make_uint4(0, 0, 0, 0);
        } else {
            // geo[9..15] -> cols NDIR+8 .. NDIR+14, col NDIR+15 = 1; lidar: cols 88..95 = 1 (tcnn input padding)
#pragma unroll
            for (int j = 8; j < 16; j += 2) {
                const float a = __half2float(gh1[j + 1 - 8]);
                const float b = j + 2 < 16 ? __half2float(gh1[j + 2 - 8]) : 1.f;   // tcnn input padding = 1
                *reinterpret_cast<uint32_t*>(xr + NDIR + j) = pack_bf2(a, b);
            }
            if (LIDAR) *reinterpret_cast<uint4*>(xr + 88) = make_uint4(kOnesH2, kOnesH2, kOnesH2, kOnesH2);
        }
        }
        if (LIDAR) {
            __syncthreads();
            const size_t first = ray * A.S > row0 ? ray * A.S : row0;
            if (inb && first != g) {   // 36 halves = 9 x 8 bytes of the ray's encoding
                const uint2* src = reinterpret_cast<const uint2*>(Xs + (first - row0) * LDX + half * 36);
                uint2* dst = reinterpret_cast<uint2*>(xr + half * 36);
#pragma unroll
                for (int i = 0; i < 9; ++i) dst[i] = src[i];
            }
        }
    }
    static __device__ __forceinline__ void sink(const Args& A, size_t row, size_t n, int col, float a,
                                                float b) {
        if (row >= n) return;
        const int k = col - NDIR + 1;  // geo index of `a`; `b` is k + 1
        float* p = A.dgeo16 + row * 16;
        if (A.accumulate) {
            p[k] += a;
            if (k + 1 < 16) p[k + 1] += b;
        } else {
            p[k] = a;
            if (k + 1 < 16) p[k + 1] = b;
        }
    }
    static __device__ __forceinline__ void sink_zero(const Args& A, size_t row0, size_t n, int lane) {
        if (A.accumulate) return;
        // 16 rows x 15 floats (cols 1..15)
        for (int i = lane; i < 16 * 15; i += 32) {
            const int r = i / 15, c = i - r * 15 + 1;
            if (row0 + r < n) A.dgeo16[(row0 + r) * 16 + c] = 0.f;
        }
    }
};


// ------------------------------------------------------------------------------------------------
// row-wise views of the same three MLPs for the tcgen05 backward (mlp_bwd_tc.cuh): thread = row
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
template <int KIN>
__device__ __forceinline__ void tc_store_x(unsigned char* xg, uint32_t t, int chunk, uint4 v) {
    *reinterpret_cast<uint4*>(xg + mlptc::x_off(KIN, t, (uint32_t)chunk)) = v;
}
struct SigmaTc {
    static constexpr int KIN = 128, NHID = 1, LDG1 = 128, OUT_ROWS = 16, DX0 = 0, DXN = 128;
    using Args = SigmaT::Args;
    // the warpgroup loads whole rows with consecutive lanes on consecutive 16 bytes (thread-per-row loads touch 32
    // lines per instruction: ncu had this kernel at 79 % l1tex with 12 % of the issue slots used)
    static __device__ __forceinline__ void load_do(const Args& A, size_t row0, size_t n, float scale, unsigned char* dog,
                                                   uint32_t t) {
        float4 g[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const size_t row = row0 + (t >> 2) + 32 * i;
            g[i] = row < n ? __ldg(reinterpret_cast<const float4*>(A.dgeo16 + row * 16) + (t & 3)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t r = (t >> 2) + 32 * i, c4 = t & 3;
            *reinterpret_cast<uint2*>(dog + mlptc::dotile_off(r, c4 >> 1) + (c4 & 1) * 8) =
                make_uint2(pack_half2(scale * g[i].x, scale * g[i].y), pack_half2(scale * g[i].z, scale * g[i].w));
        }
    }
    static __device__ __forceinline__ void load_x(const Args& A, size_t row0, size_t n, unsigned char* xg, uint32_t t,
                                                  uint32_t, unsigned char*) {
        uint4 v[16];   // all sixteen loads in flight before the first store
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const size_t row = row0 + (t >> 4) + 8 * i;
            v[i] = row < n ? __ldcs(reinterpret_cast<const uint4*>(A.feats + row * 128) + (t & 15)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) tc_store_x<128>(xg, (t >> 4) + 8 * i, (int)(t & 15), v[i]);
    }
    static __device__ __forceinline__ void prefetch(const Args& A, size_t row0, size_t n, uint32_t t) {
        const size_t rows = n - row0 < 128 ? n - row0 : 128;
        if ((size_t)t * 64 < rows * 128) prefetch_l2(A.feats + row0 * 128 + (size_t)t * 64);
        if ((size_t)(t + 128) * 64 < rows * 128) prefetch_l2(A.feats + row0 * 128 + (size_t)(t + 128) * 64);
        if ((size_t)t * 32 < rows * 16) prefetch_l2(A.dgeo16 + row0 * 16 + (size_t)t * 32);
    }
    static __device__ __forceinline__ void store_dx(const Args& A, size_t row0, size_t n, int c, const unsigned char* stage,
                                                    const unsigned char* flags, uint32_t t) {
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const uint32_t r = (t >> 4) + 8 * i, j = t & 15;
            // rows without gradient: k_encode_bwd tests dgeo16 itself and never reads them
            if (row0 + r < n && flags[r])
                *reinterpret_cast<float4*>(A.dfeat + (row0 + r) * 128 + 64 * c + 4 * j) =
                    *reinterpret_cast<const float4*>(stage + r * 256 + ((j ^ (r & 15)) << 4));
        }
    }
    static __device__ __forceinline__ void store_dead(const Args&, size_t, size_t, uint32_t) {}
};
struct FlowTc {
    static constexpr int KIN = 32, NHID = 2, LDG1 = 32, OUT_ROWS = 6, DX0 = 0, DXN = 32;
    using Args = FlowT::Args;
    static __device__ __forceinline__ void load_do(const Args& A, size_t row0, size_t n, float scale, unsigned char* dog,
                                                   uint32_t t) {
        float4 g[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const size_t row = row0 + (t >> 1) + 64 * i;
            g[i] = row < n ? __ldg(reinterpret_cast<const float4*>(A.dflow + row * 8) + (t & 1)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
            *reinterpret_cast<uint2*>(dog + mlptc::dotile_off((t >> 1) + 64 * i, 0) + (t & 1) * 8) =
                make_uint2(pack_half2(scale * g[i].x, scale * g[i].y), pack_half2(scale * g[i].z, scale * g[i].w));
        *reinterpret_cast<uint4*>(dog + mlptc::dotile_off(t, 1)) = make_uint4(0, 0, 0, 0);
    }
    static __device__ __forceinline__ void load_x(const Args& A, size_t row0, size_t n, unsigned char* xg, uint32_t t,
                                                  uint32_t, unsigned char*) {
        uint4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const size_t row = row0 + (t >> 2) + 32 * i;
            v[i] = row < n ? __ldcs(reinterpret_cast<const uint4*>(A.flowfeat + row * 32) + (t & 3)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) tc_store_x<32>(xg, (t >> 2) + 32 * i, (int)(t & 3), v[i]);
    }
    static __device__ __forceinline__ void prefetch(const Args& A, size_t row0, size_t n, uint32_t t) {
        const size_t rows = n - row0 < 128 ? n - row0 : 128;
        if ((size_t)t * 64 < rows * 32) prefetch_l2(A.flowfeat + row0 * 32 + (size_t)t * 64);
        if ((size_t)t * 32 < rows * 8) prefetch_l2(A.dflow + row0 * 8 + (size_t)t * 32);
    }
    static __device__ __forceinline__ void store_dx(const Args& A, size_t row0, size_t n, int, const unsigned char* stage,
                                                    const unsigned char* flags, uint32_t t) {
#pragma unroll 4
        for (int i = 0; i < 8; ++i) {
            const uint32_t r = (t >> 3) + 16 * i, j = t & 7;
            // rows without gradient: k_flowgrid_bwd applies the same zero test to dflow
            if (row0 + r < n && flags[r])
                *reinterpret_cast<float4*>(A.dflowfeat + (row0 + r) * 32 + 4 * j) =
                    *reinterpret_cast<const float4*>(stage + r * 128 + ((j ^ (r & 7)) << 4));
        }
    }
    static __device__ __forceinline__ void store_dead(const Args&, size_t, size_t, uint32_t) {}
};
template <bool LIDAR>
struct HeadTc {
    static constexpr int KIN = LIDAR ? 96 : 32, NHID = 2, LDG1 = KIN, OUT_ROWS = 16;
    static constexpr int NDIR = LIDAR ? 72 : 16;
    static constexpr int DX0 = NDIR, DXN = 16;
    using Args = typename HeadT<LIDAR>::Args;
    // weights / kept colours are 4 / 16 bytes per row: the row-per-thread loads are already coalesced
    static __device__ __forceinline__ void load_do(const Args& A, size_t row0, size_t n, float scale, unsigned char* dog,
                                                   uint32_t t) {
        uint4 o0 = make_uint4(0, 0, 0, 0);
        const size_t row = row0 + t;
        const float w = row < n ? __ldg(A.weights + row) : 0.f;
        if (w > 1e-4f) {   // only the samples that passed the colour mask carry gradient
            const size_t ray = row / A.S;
            const float4 cc = __ldg(reinterpret_cast<const float4*>(A.rgbs + row * 4));
            const float ws = scale * w;
            if (LIDAR) {
                const int ch = A.net == 0 ? 1 : 0;
                const float c = ch ? cc.y : cc.x;
                o0.x = pack_half2(ws * __ldg(A.g_image + ray * 2 + ch) * c * (1.f - c), 0.f);
            } else {
                o0.x = pack_half2(ws * __ldg(A.g_image + ray * 3) * cc.x * (1.f - cc.x),
                                  ws * __ldg(A.g_image + ray * 3 + 1) * cc.y * (1.f - cc.y));
                o0.y = pack_half2(ws * __ldg(A.g_image + ray * 3 + 2) * cc.z * (1.f - cc.z), 0.f);
            }
        }
        *reinterpret_cast<uint4*>(dog + mlptc::dotile_off(t, 0)) = o0;
        *reinterpret_cast<uint4*>(dog + mlptc::dotile_off(t, 1)) = make_uint4(0, 0, 0, 0);
    }
    // X row = [direction encoding (NDIR) | geo 1..15 | 1 (tcnn input padding) | lidar: 8 more ones]
    static __device__ __forceinline__ void load_x(const Args& A, size_t row0, size_t n, unsigned char* xg, uint32_t t,
                                                  uint32_t wg, unsigned char* scratch) {
        const size_t row = row0 + t;
        const bool inb = row < n;
        uint4 g0 = make_uint4(0, 0, 0, 0), g1 = g0;
        if (inb) {
            g0 = __ldg(reinterpret_cast<const uint4*>(A.geo + row * 16));
            g1 = __ldg(reinterpret_cast<const uint4*>(A.geo + row * 16) + 1);
        }
        if (LIDAR) {
            // tcnn Frequency (12 octaves) of (d + 1) / 2 is constant along a ray: the rays of the tile are encoded
            // once, cooperatively, into scratch (the still unused Ha / Hb tiles), then every row copies its ray's 144 bytes
            const size_t last = (row0 + 128 < n ? row0 + 128 : n) - 1;
            const size_t ray_a = row0 / A.S;
            const int nr = (int)(last / A.S - ray_a) + 1;
            for (int i = (int)t; i < nr * 36; i += 128) {
                const int rl = i / 36, j = 2 * (i - rl * 36);
                const int dim = j / 24, oct = (j >> 1) % 12;
                const float v = (__ldg(A.rays_d + (ray_a + rl) * 3 + dim) + 1.0f) * 0.5f;
                const float a = scalbnf(v, oct);
                *reinterpret_cast<uint32_t*>(scratch + rl * 144 + j * 2) = pack_half2(sinpif(a), sinpif(a + 0.5f));
            }
            umma::wg_barrier(wg);
            if (inb) {
                const uint4* src = reinterpret_cast<const uint4*>(scratch + (row / A.S - ray_a) * 144);
#pragma unroll
                for (int c = 0; c < 9; ++c) tc_store_x<KIN>(xg, t, c, src[c]);
            } else {
#pragma unroll
                for (int c = 0; c < 9; ++c) tc_store_x<KIN>(xg, t, c, make_uint4(0, 0, 0, 0));
            }
        } else {
            uint4 s0 = make_uint4(0, 0, 0, 0), s1 = s0;
            if (inb) {
                const size_t ray = row / A.S;
                float sh[16];
                sh4_eval(__ldg(A.rays_d + ray * 3), __ldg(A.rays_d + ray * 3 + 1), __ldg(A.rays_d + ray * 3 + 2), sh);
                s0 = make_uint4(pack_half2(sh[0], sh[1]), pack_half2(sh[2], sh[3]), pack_half2(sh[4], sh[5]), pack_half2(sh[6], sh[7]));
                s1 = make_uint4(pack_half2(sh[8], sh[9]), pack_half2(sh[10], sh[11]), pack_half2(sh[12], sh[13]), pack_half2(sh[14], sh[15]));
            }
            tc_store_x<KIN>(xg, t, 0, s0);
            tc_store_x<KIN>(xg, t, 1, s1);
        }
        // geo halves 1..15 then 1.0: the 16-half row shifted down by one half
        const uint32_t w[9] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, inb ? kOneH : 0u};
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = __funnelshift_r(w[i], w[i + 1], 16);
        constexpr int c0 = NDIR / 8;
        tc_store_x<KIN>(xg, t, c0, make_uint4(o[0], o[1], o[2], o[3]));
        tc_store_x<KIN>(xg, t, c0 + 1, make_uint4(o[4], o[5], o[6], o[7]));
        if (LIDAR) {
            const uint32_t one2 = inb ? kOnesH2 : 0u;
            tc_store_x<KIN>(xg, t, 11, make_uint4(one2, one2, one2, one2));
        }
    }
    static __device__ __forceinline__ void prefetch(const Args& A, size_t row0, size_t n, uint32_t t) {
        const size_t rows = n - row0 < 128 ? n - row0 : 128;
        if ((size_t)t * 32 < rows) prefetch_l2(A.weights + row0 + (size_t)t * 32);
        if ((size_t)t * 32 < rows * 4) prefetch_l2(A.rgbs + row0 * 4 + (size_t)t * 32);
        if ((size_t)t * 64 < rows * 16) prefetch_l2(A.geo + row0 * 16 + (size_t)t * 64);
        if (A.accumulate && (size_t)t * 32 < rows * 16) prefetch_l2(A.dgeo16 + row0 * 16 + (size_t)t * 32);
    }
    // dgeo16[row][1..15] (=, +=) dX[row][0..14]; column 0 (the sigma-logit slot) is someone else's
    static __device__ __forceinline__ void store_dx(const Args& A, size_t row0, size_t n, int, const unsigned char* stage,
                                                    const unsigned char* flags, uint32_t t) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t r = (t >> 2) + 32 * i, j = t & 3;
            if (row0 + r >= n || (A.accumulate && !flags[r])) continue;
            auto dx = [&](int e) {   // staged element e of row r (pieces XOR-swizzled by row)
                return *reinterpret_cast<const float*>(stage + r * 64 + ((((uint32_t)e >> 2) ^ (r & 3)) << 4) + (e & 3) * 4);
            };
            float4* p = reinterpret_cast<float4*>(A.dgeo16 + (row0 + r) * 16) + j;
            float4 v = make_float4(j ? dx(4 * j - 1) : 0.f, dx(4 * j), dx(4 * j + 1), dx(4 * j + 2));
            if (A.accumulate) {
                const float4 o = *p;
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            } else if (j == 0) {
                v.x = reinterpret_cast<const float*>(p)[0];
            }
            *p = v;
        }
    }
    static __device__ __forceinline__ void store_dead(const Args& A, size_t row0, size_t n, uint32_t t) {
        if (A.accumulate) return;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t r = (t >> 2) + 32 * i, j = t & 3;
            if (row0 + r >= n) continue;
            float* p = A.dgeo16 + (row0 + r) * 16 + 4 * j;
            if (j) *reinterpret_cast<float4*>(p) = make_float4(0.f, 0.f, 0.f, 0.f);
            else { p[1] = 0.f; p[2] = 0.f; p[3] = 0.f; }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// power-of-two gradient scale of one k_mlp_bwd launch: scale * max|dOut| ~ 16 (fp16 keeps 12 binades
// of head-room for the growth through the layers and 14 + 10 below for the small rows)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_absmax(const float* __restrict__ x, size_t n, float mul, unsigned* __restrict__ out) {
    // a pure streaming read (200 MB of d sigma-net outputs per modality and step): 16-byte loads, four of them in
    // flight per thread; scalar when the slice does not start on a 16-byte boundary (g_image + r0 * nch)
    float m = 0.f;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    size_t done = 0;
    if ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        const size_t n4 = n / 4;
        auto fold = [&](const float4& v) {
            m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        };
        size_t i = tid;
        for (; i + 3 * nth < n4; i += 4 * nth) {
            const float4 a = __ldg(x4 + i), b = __ldg(x4 + i + nth), c = __ldg(x4 + i + 2 * nth),
                         d = __ldg(x4 + i + 3 * nth);
            fold(a); fold(b); fold(c); fold(d);
        }
        for (; i < n4; i += nth) fold(__ldg(x4 + i));
        done = n4 * 4;
    }
    for (size_t i = done + tid; i < n; i += nth) m = fmaxf(m, fabsf(__ldg(x + i)));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    m *= mul;
    if ((threadIdx.x & 31) == 0 && m > 0.f && m < INFINITY) atomicMax(out, __float_as_uint(m));
}
__global__ void k_make_scale(const unsigned* __restrict__ mx, float* __restrict__ scale2) {
    const float m = __uint_as_float(*mx);
    float s = 1.0f;
    if (m > 0.f) {
        int e = 0;
        frexpf(m, &e);                 // m = f * 2^e, f in [0.5, 1)
        e = min(max(5 - e, -100), 100);  // scale * m in [8, 16)
        s = scalbnf(1.0f, e);
    }
    scale2[0] = s;
    scale2[1] = 1.0f / s;
}

// ------------------------------------------------------------------------------------------------
// compositing backward: one warp per ray
// ------------------------------------------------------------------------------------------------
template <bool LIDAR>
__global__ void __launch_bounds__(128)
k_composite_bwd(const __grid_constant__ nvsf_field_config_t cfg, const float* __restrict__ nears,
                const float* __restrict__ fars, const float* __restrict__ noise,
                const float* __restrict__ sigma, const __half* __restrict__ geo,
                const float* __restrict__ rgbs, const float* __restrict__ weights, uint32_t N,
                uint32_t S, float bg_color, const float* __restrict__ g_depth,
                const float* __restrict__ g_image, const float* __restrict__ g_wsum,
                const float* __restrict__ g_weights, float* __restrict__ dgeo16) {
    constexpr int NCH = LIDAR ? 2 : 3;
    const int lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= N) return;
    const float near = __ldg(nears + r), far = __ldg(fars + r);
    const float kexp = (cfg.active_sensor ? 2.0f : 1.0f) * cfg.density_scale;
    const float gd = g_depth ? __ldg(g_depth + r) : 0.f;
    float gi[NCH], gsum = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        gi[c] = g_image ? __ldg(g_image + (size_t)r * NCH + c) : 0.f;
        gsum += gi[c];
    }
    // camera: image += (1 - weights_sum) * bg  (renderer_dynamic.py:236-237)
    const float gws = (g_wsum ? __ldg(g_wsum + r) : 0.f) - (LIDAR ? 0.f : bg_color * gsum);

    auto dl_dw = [&](size_t g, float z) {
        const float4 cc = __ldg(reinterpret_cast<const float4*>(rgbs + g * 4));
        float q = gi[0] * cc.x + gi[1] * cc.y;
        if (!LIDAR) q += gi[2] * cc.z;
        float v = gws + gd * z + q;
        if (g_weights) v += __ldg(g_weights + g);
        return v;
    };

    // sweep 1 (front to back): transmittance at the start of every 32-sample chunk
    extern __shared__ float chunk_T[];   // [warps][ceil(S/32)]
    const uint32_t n_chunks = (S + 31) / 32;
    float* myT = chunk_T + (threadIdx.x >> 5) * n_chunks;
    auto one_minus_alpha = [&](uint32_t i, size_t g, float& z, float& delta) {
        z = uniform_z(near, far, i, S, noise, g);
        if (i + 1 < S) delta = uniform_z(near, far, i + 1, S, noise, g + 1) - z;
        else delta = (far - near) / (float)S;
        return expf((-kexp * delta) * __ldg(sigma + g));
    };
    {
        float carry = 1.0f;
        for (uint32_t c = 0; c < n_chunks; ++c) {
            const uint32_t i = c * 32 + lane;
            float v = 1.0f;
            if (i < S) {
                float z, delta;
                // exactly the forward's factor: alpha is rounded to fp32 first (renderer_dynamic.py:185-191)
                v = (1.0f - (1.0f - one_minus_alpha(i, (size_t)r * S + i, z, delta))) + 1e-15f;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) myT[c] = carry;
            carry *= v;
        }
    }
    __syncwarp();
    // sweep 2 (back to front): dL/dalpha_i = dL/dw_i T_i - (sum_{j>i} dL/dw_j w_j) / v_i.  The suffix
    // sum is accumulated from the far end, so samples behind a saturated ray get EXACTLY zero
    // (they are skipped by every later stage of the backward).
    float suffix = 0.f;
    for (int c = (int)n_chunks - 1; c >= 0; --c) {
        const uint32_t i = (uint32_t)c * 32 + lane;
        const bool in = i < S;
        const size_t g = (size_t)r * S + (in ? i : S - 1);
        float z, delta;
        const float e = one_minus_alpha(in ? i : S - 1, g, z, delta);   // exp(-k delta sigma)
        const float v = in ? (1.0f - (1.0f - e)) + 1e-15f : 1.0f;       // the forward's cumprod factor
        float incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl *= o;
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float T = myT[c] * excl;
        const float w = in ? __ldg(weights + g) : 0.f;
        const float dw = in ? dl_dw(g, z) : 0.f;
        float sfx = dw * w;  // inclusive suffix of dL/dw_j w_j inside the chunk
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float o = __shfl_down_sync(0xffffffffu, sfx, d);
            if (lane + d < 32) sfx += o;
        }
        const float chunk_sum = __shfl_sync(0xffffffffu, sfx, 0);
        float after = __shfl_down_sync(0xffffffffu, sfx, 1);
        if (lane == 31) after = 0.f;
        after += suffix;   // sum over j > i
        suffix += chunk_sum;
        if (in) {
            // d alpha / d sigma = kexp * delta * exp(-k delta sigma)
            const float dsig = kexp * delta * e * (dw * T - after / v);
            const float h0 = __half2float(geo[g * kGeo]);
            dgeo16[g * 16] = dsig * expf(fminf(fmaxf(h0, -15.f), 15.f));  // activation.py:17-19
        }
    }
}

// ------------------------------------------------------------------------------------------------
// encoder backward
// ------------------------------------------------------------------------------------------------
int g_enc_bwd_ctas = 2;  // nvsf_set_option("enc_bwd_ctas", 2 | 3)
// Extra binades of the per-launch power-of-two gradient scale of the fp16 MLP backward kernels (the scaled
// max |dOut| lands in [2^(3+shift), 2^(4+shift))): the larger the scale, the fewer small rows fall into the fp16
// subnormals; the bound is overflow of dH = dOut W through the layers (65504).  Options "bwd_shift_flow" /
// "bwd_shift_sigma" / "bwd_shift_heads".
int g_shift_flow = 0, g_shift_sigma = 0, g_shift_heads = 2;
float shift_mul(int s) { return s >= 0 ? 1.0f / (float)(1 << s) : (float)(1 << -s); }

struct GradTables {      // time-collapsed / channel-last gradient tables (scratch, zeroed per call)
    float* pls;          // layout of WsLayout::pls
    float* pld;          // [3 queries][pld_per_q]
    float* dyn;          // [dyn_per_q]      (un-warped query only: the warped ones carry no gradient)
    float* flow;         // float2 [fl_entries]
    float* hs;           // fp32 [hs_entries][4]: the caller's hash_static gradient itself
};

__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red1(float* p, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

// Runs of equal keys over consecutive lanes (consecutive samples of a ray fall into the same
// cell): `heads` has a bit for every lane that starts a run; seg_sum leaves in each head lane the
// sum over its run.  `nsteps` (warp-uniform) is the number of doubling steps the longest run needs.
struct Runs {
    unsigned heads;
    bool head;
    int nsteps;
    bool take[5];  // lane adds the value of lane + (1 << k): no run starts in (lane, lane + (1 << k)]
};
__device__ __forceinline__ Runs runs_from_heads(bool head, int lane) {
    Runs r;
    r.head = head;
    r.heads = __ballot_sync(0xffffffffu, head);
    const unsigned above = (r.heads >> lane) >> 1;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int d = 1 << k;
        r.take[k] = (lane + d < 32) && ((above & ((1u << d) - 1u)) == 0u);
    }
    // step k is needed iff some run is longer than 2^k lanes, i.e. 2^k consecutive non-head lanes
    const unsigned z1 = ~r.heads, z2 = z1 & (z1 >> 1), z4 = z2 & (z2 >> 2), z8 = z4 & (z4 >> 4),
                   z16 = z8 & (z8 >> 8);
    r.nsteps = (z1 != 0u) + (z2 != 0u) + (z4 != 0u) + (z8 != 0u) + (z16 != 0u);
    return r;
}
__device__ __forceinline__ float seg_sum(float v, const Runs& r) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        if (k < r.nsteps) {
            const float o = __shfl_down_sync(0xffffffffu, v, 1 << k);
            if (r.take[k]) v += o;
        }
    }
    return v;
}

// Transposed run reduction of the plane scatters.  The lane that owns a sample stages its eight
// channel gradients and the texel (index, weights, "last sample of its run inside this group of 8")
// in shared memory; then lane (q, f) walks the 8 consecutive samples of group q for channel f,
// accumulates the corner sums in registers while the texel stays the same and issues one red per
// corner when it changes: the 8 lanes of a group add 32 contiguous bytes.  Replaces a segmented
// shuffle reduction of 32 (2-D) / 16 (1-D) values per lane, which was 2/3 of k_encode_bwd's
// instructions (profiles/r01_train_bwd_v2_ncu_full.txt: SHFL 23 %, predicated FADD 27 %).
struct __align__(16) WarpStage {
    float d[36 * 8];  // row(s) = s + (s >> 3): the four groups read different banks
    uint4 info[32];
};
__device__ __forceinline__ void stage_rows(WarpStage& st, int lane, const float (&g)[8], uint32_t idx0,
                                           float wx, float wy, uint32_t flags) {
    const uint32_t nxt = __shfl_down_sync(0xffffffffu, idx0, 1);
    if ((lane & 7) == 7 || nxt != idx0) flags |= 4u;
    __syncwarp();  // the previous scatter's readers are done
    float4* row = reinterpret_cast<float4*>(st.d + (lane + (lane >> 3)) * 8);
    row[0] = make_float4(g[0], g[1], g[2], g[3]);
    row[1] = make_float4(g[4], g[5], g[6], g[7]);
    st.info[lane] = make_uint4(idx0, __float_as_uint(wx), __float_as_uint(wy), flags);
    __syncwarp();
}

struct Bilin {
    uint32_t x0, x1, y0, y1;
    float wx, wy;
};
__device__ __forceinline__ void ld8t(const float* p, float (&v)[8]) { ld8(p, v); }
__device__ __forceinline__ void ld8t(const __half* p, float (&v)[8]) { ld8h(p, v); }
template <class TT>
__device__ __forceinline__ void sample2d(const TT* __restrict__ base, uint32_t R, float pa, float pb,
                                         Bilin& b, float (&out)[8]) {
    plane_coord(pa, R, b.x0, b.x1, b.wx);
    plane_coord(pb, R, b.y0, b.y1, b.wy);
    float a[8], bb[8], c[8], d[8];
    ld8t(base + ((size_t)b.y0 * R + b.x0) * 8, a);
    ld8t(base + ((size_t)b.y0 * R + b.x1) * 8, bb);
    ld8t(base + ((size_t)b.y1 * R + b.x0) * 8, c);
    ld8t(base + ((size_t)b.y1 * R + b.x1) * 8, d);
    const float w00 = (1.f - b.wx) * (1.f - b.wy), w01 = b.wx * (1.f - b.wy),
                w10 = (1.f - b.wx) * b.wy, w11 = b.wx * b.wy;
#pragma unroll
    for (int f = 0; f < 8; ++f) out[f] = w00 * a[f] + w01 * bb[f] + w10 * c[f] + w11 * d[f];
}
__device__ __forceinline__ void scatter2d(float* __restrict__ base, uint32_t R, const Bilin& b,
                                          const float (&g)[8], WarpStage& st, int lane) {
    stage_rows(st, lane, g, b.y0 * R + b.x0, b.wx, b.wy,
               (b.x1 != b.x0 ? 1u : 0u) | (b.y1 != b.y0 ? 2u : 0u));
    const int q = lane >> 3, f = lane & 7;
    float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint4 in = st.info[8 * q + i];
        const float v = st.d[(9 * q + i) * 8 + f];
        const float wx = __uint_as_float(in.y), wy = __uint_as_float(in.z);
        const float v0 = v * (1.f - wx), v1 = v * wx;
        a00 = fmaf(v0, 1.f - wy, a00);
        a01 = fmaf(v1, 1.f - wy, a01);
        a10 = fmaf(v0, wy, a10);
        a11 = fmaf(v1, wy, a11);
        if (in.w & 4u) {   // (predicated reds instead of this branch measured slower: 4.0 -> 4.3 ms)
            float* t = base + (size_t)in.x * 8 + f;
            const uint32_t ox = (in.w & 1u) ? 8u : 0u, oy = (in.w & 2u) ? R * 8u : 0u;
            red1(t, a00);
            red1(t + ox, a01);
            red1(t + oy, a10);
            red1(t + ox + oy, a11);
            a00 = a01 = a10 = a11 = 0.f;
        }
    }
}

struct Lin {
    uint32_t x0, x1;
    float wx;
    float dcoord;  // d(texel coordinate)/d(p): R-1 inside the grid, 0 where grid_sample clamps
};
template <class TT>
__device__ __forceinline__ void sample1d(const TT* __restrict__ base, uint32_t R, float pa, Lin& l,
                                         float (&out)[8], float (&slope)[8]) {
    plane_coord(pa, R, l.x0, l.x1, l.wx);
    const float f = ((pa * 2.0f - 1.0f) + 1.0f) * 0.5f * (float)(R - 1);
    l.dcoord = (f > 0.f && f < (float)(R - 1)) ? (float)(R - 1) : 0.f;
    float a[8], b[8];
    ld8t(base + (size_t)l.x0 * 8, a);
    ld8t(base + (size_t)l.x1 * 8, b);
#pragma unroll
    for (int f2 = 0; f2 < 8; ++f2) {
        out[f2] = (1.f - l.wx) * a[f2] + l.wx * b[f2];
        slope[f2] = b[f2] - a[f2];
    }
}
__device__ __forceinline__ void scatter1d(float* __restrict__ base, const Lin& l, const float (&g)[8],
                                          WarpStage& st, int lane) {
    stage_rows(st, lane, g, l.x0, l.wx, 0.f, l.x1 != l.x0 ? 1u : 0u);
    const int q = lane >> 3, f = lane & 7;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint4 in = st.info[8 * q + i];
        const float v = st.d[(9 * q + i) * 8 + f];
        const float wx = __uint_as_float(in.y);
        a0 = fmaf(v, 1.f - wx, a0);
        a1 = fmaf(v, wx, a1);
        if (in.w & 4u) {
            float* t = base + (size_t)in.x * 8 + f;
            red1(t, a0);
            red1(t + ((in.w & 1u) ? 8u : 0u), a1);
            a0 = a1 = 0.f;
        }
    }
}

// The per-sample inputs of the encoder backward (610 B: 128 fp32 feature gradients, dgeo, flow) are read exactly
// once: loaded evict-first (ld.global.cs) they do not push the gradient tables — 67 MB of fp32 static-hash
// gradient that the reds hit at random — out of the L2 (tools/ubench_red.cu: random reds run at 189 G sectors/s
// on an L2-resident table, 60 G/s on a 128 MB one, 25 G/s from DRAM).
// (measured: k_encode_bwd 7.88 -> 7.60 ms per training step).
// Where its time goes (training step of 2 x 3.1 M samples, each part left out in turn): time planes 2.46 ms, static
// hash 2.06, dynamic hashes 1.57, space planes 1.04, sample set-up 0.4 of 7.53 ms.
int g_enc_bwd_h16 = 1;   // option "enc_bwd_h16": plane texels of the product rule from the fp16 mirrors
template <bool STREAM>
__device__ __forceinline__ float4 lds4(const float4* p) { return STREAM ? __ldcs(p) : __ldg(p); }
template <bool STREAM>
__device__ __forceinline__ float lds1(const float* p) { return STREAM ? __ldcs(p) : __ldg(p); }
template <bool STREAM>
__device__ __forceinline__ void ld8s(const float* p, float (&v)[8]) {
    const float4 a = lds4<STREAM>(reinterpret_cast<const float4*>(p)), b = lds4<STREAM>(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
#define NVSF_LDS4(p) lds4<STREAM>(p)
#define NVSF_LDS1(p) lds1<STREAM>(p)
// MIN_CTAS 2 = 128 registers / 16 warps per SM; 3 = 80 registers (80 bytes of spills) / 24 warps
template <bool FROM_RAYS, int MIN_CTAS, bool H16>
__global__ void __launch_bounds__(256, MIN_CTAS)
k_encode_bwd(const __grid_constant__ nvsf_field_config_t cfg, const __grid_constant__ FieldPtrs P,
             const GradTables G, const float* __restrict__ xin, const float* __restrict__ rays_o,
             const float* __restrict__ rays_d, const float* __restrict__ nears,
             const float* __restrict__ fars, const float* __restrict__ noise, uint32_t S,
             size_t begin, size_t count, const float* __restrict__ flow_in /* [.,8] global index */,
             const float* __restrict__ dgeo16 /* chunk-local */, const float* __restrict__ dfeat,
             float* __restrict__ dflow_out, const float* __restrict__ scale2) {
    constexpr bool STREAM = true;
    const size_t li = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool inb = li < count;
    // Rows whose sigma-net output gradient is zero AS THE SIGMA BACKWARD SAW IT (scaled, fp16) have
    // zero feature gradient; k_mlp_bwd<SigmaT> does not even write dfeat for warps of such rows.
    bool live = false;
    if (inb) {
        const float sc = __ldg(scale2);
        const float4* d = reinterpret_cast<const float4*>(dgeo16 + li * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 v = NVSF_LDS4(d + i);
            live = live || ((pack_half2(sc * v.x, sc * v.y) | pack_half2(sc * v.z, sc * v.w)) & 0x7fff7fffu) != 0;
        }
    }
    if (inb && !live) {
        float4* o = reinterpret_cast<float4*>(dflow_out + li * 8);
        o[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (__ballot_sync(0xffffffffu, live) == 0) return;
    __shared__ WarpStage stage[8];
    const int lane = threadIdx.x & 31;
    WarpStage& st = stage[threadIdx.x >> 5];

    float x = 0.5f, y = 0.5f, z = 0.5f;
    float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
    if (inb) {
        // same arithmetic as the forward (field_split.cu sample_position)
        const size_t g = begin + li;
        float px, py, pz;
        if (FROM_RAYS) {
            const size_t r = g / S;
            const uint32_t k = (uint32_t)(g - r * S);
            const float zz = uniform_z(__ldg(nears + r), __ldg(fars + r), k, S, noise, g);
            px = __ldg(rays_o + r * 3 + 0) + __ldg(rays_d + r * 3 + 0) * zz;
            py = __ldg(rays_o + r * 3 + 1) + __ldg(rays_d + r * 3 + 1) * zz;
            pz = __ldg(rays_o + r * 3 + 2) + __ldg(rays_d + r * 3 + 2) * zz;
            px = fminf(fmaxf(px, -cfg.bound), cfg.bound);
            py = fminf(fmaxf(py, -cfg.bound), cfg.bound);
            pz = fminf(fmaxf(pz, -cfg.bound), cfg.bound);
        } else {
            px = __ldg(xin + g * 3 + 0); py = __ldg(xin + g * 3 + 1); pz = __ldg(xin + g * 3 + 2);
        }
        const float inv2b = 1.0f / (2.0f * cfg.bound);
        x = (px + cfg.bound) * inv2b; y = (py + cfg.bound) * inv2b; z = (pz + cfg.bound) * inv2b;
        f0 = NVSF_LDS4(reinterpret_cast<const float4*>(flow_in + g * 8));
        f1 = NVSF_LDS4(reinterpret_cast<const float4*>(flow_in + g * 8) + 1);
    }
    const int valid1 = P.ti->valid[1], valid2 = P.ti->valid[2];
    float qx[3], qy[3], qz[3];
    int qi[3];
    qx[0] = x; qy[0] = y; qz[0] = z; qi[0] = 0;
    qx[1] = valid1 ? x + f0.x : x; qy[1] = valid1 ? y + f0.y : y;
    qz[1] = valid1 ? z + f0.z : z; qi[1] = valid1 ? 1 : 0;
    qx[2] = valid2 ? x + f0.w : x; qy[2] = valid2 ? y + f1.x : y;
    qz[2] = valid2 ? z + f1.y : z; qi[2] = valid2 ? 2 : 0;
    const float* df = dfeat + li * 128;
    const float lv_ = live ? 1.f : 0.f;

    // (a) space planes: f = A(xy) * B(xz) * C(yz)
#pragma unroll 1
    for (int s = 0; s < kPlScales; ++s) {
        const uint32_t R = cfg.pl_res[s];
        // the texels of the product rule come from the fp16 mirrors the forward itself interpolates (one 16-byte
        // load per texel instead of two): the re-gathers are 240 of this kernel's loads per sample
        const float* base32 = P.pls + P.pls_scale[s];
        const __half* base16 = P.pls16 + P.pls_scale[s];
        float* gbase = G.pls + P.pls_scale[s];
        float g8[8], A[8], B[8], C[8];
        if (live) ld8s<STREAM>(df + 8 * s, g8);
        else {
#pragma unroll
            for (int f = 0; f < 8; ++f) g8[f] = 0.f;
        }
        Bilin ba, bb, bc;
        if (H16) {
            sample2d(base16, R, x, y, ba, A);
            sample2d(base16 + (size_t)R * R * 8, R, x, z, bb, B);
            sample2d(base16 + (size_t)2 * R * R * 8, R, y, z, bc, C);
        } else {
            sample2d(base32, R, x, y, ba, A);
            sample2d(base32 + (size_t)R * R * 8, R, x, z, bb, B);
            sample2d(base32 + (size_t)2 * R * R * 8, R, y, z, bc, C);
        }
        float d8[8];
#pragma unroll
        for (int f = 0; f < 8; ++f) d8[f] = g8[f] * B[f] * C[f];
        scatter2d(gbase, R, ba, d8, st, lane);
#pragma unroll
        for (int f = 0; f < 8; ++f) d8[f] = g8[f] * A[f] * C[f];
        scatter2d(gbase + (size_t)R * R * 8, R, bb, d8, st, lane);
#pragma unroll
        for (int f = 0; f < 8; ++f) d8[f] = g8[f] * A[f] * B[f];
        scatter2d(gbase + (size_t)2 * R * R * 8, R, bc, d8, st, lane);
    }

    // (b) time planes (collapsed rows): sum_q wq * A_q(x_q) * B_q(y_q) * C_q(z_q); the warped
    // queries also give dL/dflow through the coordinate gradient of grid_sample
    float dfl[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int s = 0; s < kPlScales; ++s) {
        const uint32_t R = cfg.pl_res[s];
        float g8[8];
        if (live) ld8s<STREAM>(df + 32 + 8 * s, g8);
        else {
#pragma unroll
            for (int f = 0; f < 8; ++f) g8[f] = 0.f;
        }
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {
            const size_t toff = (size_t)qi[q] * P.pld_per_q + P.pld_scale[s];
            const float* base32 = P.pld + toff;
            const __half* base16 = P.pld16 + toff;
            float* gbase = G.pld + toff;
            const float wq = q == 0 ? 0.5f : 0.25f;
            float A[8], B[8], C[8], sa[8], sb[8], sc[8];
            Lin la, lb, lc;
            if (H16) {
                sample1d(base16, R, qx[q], la, A, sa);
                sample1d(base16 + (size_t)R * 8, R, qy[q], lb, B, sb);
                sample1d(base16 + (size_t)2 * R * 8, R, qz[q], lc, C, sc);
            } else {
                sample1d(base32, R, qx[q], la, A, sa);
                sample1d(base32 + (size_t)R * 8, R, qy[q], lb, B, sb);
                sample1d(base32 + (size_t)2 * R * 8, R, qz[q], lc, C, sc);
            }
            float d8[8], gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
            for (int f = 0; f < 8; ++f) {
                const float gw = wq * g8[f];
                d8[f] = gw * B[f] * C[f];
                gx = fmaf(d8[f], sa[f], gx);
                gy = fmaf(gw * A[f] * C[f], sb[f], gy);
                gz = fmaf(gw * A[f] * B[f], sc[f], gz);
            }
            scatter1d(gbase, la, d8, st, lane);
#pragma unroll
            for (int f = 0; f < 8; ++f) d8[f] = wq * g8[f] * A[f] * C[f];
            scatter1d(gbase + (size_t)R * 8, lb, d8, st, lane);
#pragma unroll
            for (int f = 0; f < 8; ++f) d8[f] = wq * g8[f] * A[f] * B[f];
            scatter1d(gbase + (size_t)2 * R * 8, lc, d8, st, lane);
            if (q > 0 && qi[q] != 0) {
                dfl[3 * (q - 1) + 0] += gx * la.dcoord;
                dfl[3 * (q - 1) + 1] += gy * lb.dcoord;
                dfl[3 * (q - 1) + 2] += gz * lc.dcoord;
            }
        }
    }
    if (inb && live) {
        float4* o = reinterpret_cast<float4*>(dflow_out + li * 8);
        if (STREAM) {
            __stcs(o, make_float4(dfl[0], dfl[1], dfl[2], dfl[3]));
            __stcs(o + 1, make_float4(dfl[4], dfl[5], 0.f, 0.f));
        } else {
            o[0] = make_float4(dfl[0], dfl[1], dfl[2], dfl[3]);
            o[1] = make_float4(dfl[4], dfl[5], 0.f, 0.f);
        }
    }
    // (c) static 3-D hash: tcnn grid backward, fp32 vector reds into the caller's gradient.
    // The hash scatters are bound by atomic throughput (about one red lane-operation per two
    // clocks per SM, whatever its width): on the coarse levels, where consecutive samples of a
    // ray share a cell (a warp holds at most 16 runs), the corner products are summed over each
    // run with shuffles first and only the head lane of a run issues the reds.
    const bool plive = __shfl_up_sync(0xffffffffu, (int)live, 1) != 0;
    const bool edge = lane == 0 || !live || !plive;
    // (the per-level gradient rows are loaded one iteration ahead: a dependent global load per
    // level was the largest stall of this part)
    float4 g4n = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) g4n = NVSF_LDS4(reinterpret_cast<const float4*>(df + 64));
#pragma unroll 1
    for (int l = 0; l < kHsLevels; ++l) {
        const LevelArgs L = lv(cfg.hs[l]);
        const float4 g4 = g4n;
        if (live && l + 1 < kHsLevels) g4n = NVSF_LDS4(reinterpret_cast<const float4*>(df + 64 + 4 * (l + 1)));
        uint32_t cx, cy, cz;
        float wx, wy, wz;
        grid_pos(L.scale, x, cx, wx);
        grid_pos(L.scale, y, cy, wy);
        grid_pos(L.scale, z, cz, wz);
        const bool head = edge | (__shfl_up_sync(0xffffffffu, cx, 1) != cx) |
                          (__shfl_up_sync(0xffffffffu, cy, 1) != cy) |
                          (__shfl_up_sync(0xffffffffu, cz, 1) != cz);
        const Runs r = runs_from_heads(head, lane);
        if (__popc(r.heads) <= 16) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float w = ((c & 1) ? wx : 1.f - wx) * ((c & 2) ? wy : 1.f - wy) *
                                ((c & 4) ? wz : 1.f - wz);
                const float s0 = seg_sum(w * g4.x, r), s1 = seg_sum(w * g4.y, r),
                            s2 = seg_sum(w * g4.z, r), s3 = seg_sum(w * g4.w, r);
                if (head && live) {
                    const uint32_t idx = L.offset + idx3(L, cx + (c & 1), cy + ((c >> 1) & 1), cz + (c >> 2));
                    red4(G.hs + (size_t)idx * 4, s0, s1, s2, s3);
                }
            }
        } else if (live) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float w = ((c & 1) ? wx : 1.f - wx) * ((c & 2) ? wy : 1.f - wy) *
                                ((c & 4) ? wz : 1.f - wz);
                const uint32_t idx = L.offset + idx3(L, cx + (c & 1), cy + ((c >> 1) & 1), cz + (c >> 2));
                red4(G.hs + (size_t)idx * 4, w * g4.x, w * g4.y, w * g4.z, w * g4.w);
            }
        }
    }
    // (d) dynamic 2-D hashes: gradient through the un-warped query only; a missing neighbour
    // frame re-uses the un-warped feature (network_dynamic.py:238-239) -> factor 0.5 + 0.25 each
    const float fac = lv_ * (0.5f + (valid1 ? 0.f : 0.25f) + (valid2 ? 0.f : 0.25f));
    float gsn = live ? NVSF_LDS1(df + 96) : 0.f;
#pragma unroll 1
    for (int p = 0; p < 3; ++p) {
        const float u = p == 2 ? y : x, w2 = p == 0 ? y : z;
        float* tab = G.dyn + P.dyn_plane[p];
#pragma unroll 1
        for (int l = 0; l < kHdLevels; ++l) {
            const LevelArgs L = lv(cfg.hd[p][l]);
            const float gs = fac * gsn;
            if (live && 8 * p + l + 1 < 3 * kHdLevels) gsn = NVSF_LDS1(df + 96 + 8 * p + l + 1);
            uint32_t cu, cv;
            float wu, wv;
            grid_pos(L.scale, u, cu, wu);
            grid_pos(L.scale, w2, cv, wv);
            const bool head = edge | (__shfl_up_sync(0xffffffffu, cu, 1) != cu) |
                              (__shfl_up_sync(0xffffffffu, cv, 1) != cv);
            const Runs r = runs_from_heads(head, lane);
            float v00 = (1.f - wu) * (1.f - wv) * gs, v01 = wu * (1.f - wv) * gs,
                  v10 = (1.f - wu) * wv * gs, v11 = wu * wv * gs;
            bool emit = live;
            if (__popc(r.heads) <= 16) {
                v00 = seg_sum(v00, r); v01 = seg_sum(v01, r);
                v10 = seg_sum(v10, r); v11 = seg_sum(v11, r);
                emit = head && live;
            }
            if (emit) {
                red1(tab + L.offset + idx2(L, cu, cv), v00);
                red1(tab + L.offset + idx2(L, cu + 1, cv), v01);
                red1(tab + L.offset + idx2(L, cu, cv + 1), v10);
                red1(tab + L.offset + idx2(L, cu + 1, cv + 1), v11);
            }
        }
    }
}

// flow-grid scatter: dL/d(collapsed flow grid) from dL/d(flow-grid features)
template <bool FROM_RAYS>
__global__ void __launch_bounds__(256)
k_flowgrid_bwd(const __grid_constant__ nvsf_field_config_t cfg, float* __restrict__ gflow,
               const float* __restrict__ xin, const float* __restrict__ rays_o,
               const float* __restrict__ rays_d, const float* __restrict__ nears,
               const float* __restrict__ fars, const float* __restrict__ noise, uint32_t S,
               size_t begin, size_t count, const float* __restrict__ dflow,
               const float* __restrict__ dflowfeat, const float* __restrict__ scale2) {
    const size_t li = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool live = li < count;
    if (live) {   // same zero test as k_mlp_bwd<FlowT> (scaled fp16): it leaves dflowfeat unwritten for such rows
        const float sc = __ldg(scale2);
        const float4 a = __ldg(reinterpret_cast<const float4*>(dflow + li * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(dflow + li * 8) + 1);
        live = ((pack_half2(sc * a.x, sc * a.y) | pack_half2(sc * a.z, sc * a.w) | pack_half2(sc * b.x, sc * b.y)) &
                0x7fff7fffu) != 0;
    }
    if (__ballot_sync(0xffffffffu, live) == 0) return;
    float x = 0.5f, y = 0.5f, z = 0.5f;
    if (live) {
        const size_t g = begin + li;
        float px, py, pz;
        if (FROM_RAYS) {
            const size_t r = g / S;
            const uint32_t k = (uint32_t)(g - r * S);
            const float zz = uniform_z(__ldg(nears + r), __ldg(fars + r), k, S, noise, g);
            px = __ldg(rays_o + r * 3 + 0) + __ldg(rays_d + r * 3 + 0) * zz;
            py = __ldg(rays_o + r * 3 + 1) + __ldg(rays_d + r * 3 + 1) * zz;
            pz = __ldg(rays_o + r * 3 + 2) + __ldg(rays_d + r * 3 + 2) * zz;
            px = fminf(fmaxf(px, -cfg.bound), cfg.bound);
            py = fminf(fmaxf(py, -cfg.bound), cfg.bound);
            pz = fminf(fmaxf(pz, -cfg.bound), cfg.bound);
        } else {
            px = __ldg(xin + g * 3 + 0); py = __ldg(xin + g * 3 + 1); pz = __ldg(xin + g * 3 + 2);
        }
        const float inv2b = 1.0f / (2.0f * cfg.bound);
        x = (px + cfg.bound) * inv2b; y = (py + cfg.bound) * inv2b; z = (pz + cfg.bound) * inv2b;
    }
    // The kernel is bound by the L2 atomic units (ncu: issue slots 9 % busy, lg_throttle), and on
    // the coarse levels consecutive samples of a ray sit in the same cell (a LiDAR ray has 110
    // samples per cell on level 0, 4 on level 9).  Where a warp holds at most 16 runs of equal
    // cells, the 16 corner x feature products are summed over each run with shuffles and only the
    // first lane of a run issues the 8 reds.
    float2 g2n = make_float2(0.f, 0.f);
    if (live) g2n = __ldg(reinterpret_cast<const float2*>(dflowfeat + li * 32));
#pragma unroll 1
    for (int l = 0; l < kFlLevels; ++l) {
        const LevelArgs L = lv(cfg.fl[l]);
        const float2 g2 = g2n;
        if (live && l + 1 < kFlLevels) g2n = __ldg(reinterpret_cast<const float2*>(dflowfeat + li * 32 + 2 * (l + 1)));
        uint32_t cx, cy, cz;
        float wx, wy, wz;
        grid_pos(L.scale, x, cx, wx);
        grid_pos(L.scale, y, cy, wy);
        grid_pos(L.scale, z, cz, wz);
        const uint32_t px_ = __shfl_up_sync(0xffffffffu, cx, 1), py_ = __shfl_up_sync(0xffffffffu, cy, 1),
                       pz_ = __shfl_up_sync(0xffffffffu, cz, 1);
        const bool plive = __shfl_up_sync(0xffffffffu, (int)live, 1) != 0;
        const bool head = lane == 0 || !live || !plive || px_ != cx || py_ != cy || pz_ != cz;
        const Runs r = runs_from_heads(head, lane);
        if (__popc(r.heads) <= 16) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float w = ((c & 1) ? wx : 1.f - wx) * ((c & 2) ? wy : 1.f - wy) *
                                ((c & 4) ? wz : 1.f - wz);
                const float sx = seg_sum(w * g2.x, r), sy = seg_sum(w * g2.y, r);
                if (head && live) {
                    const uint32_t idx = L.offset + idx3(L, cx + (c & 1), cy + ((c >> 1) & 1), cz + (c >> 2));
                    red2(gflow + (size_t)idx * 2, sx, sy);
                }
            }
        } else if (live) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float w = ((c & 1) ? wx : 1.f - wx) * ((c & 2) ? wy : 1.f - wy) *
                                ((c & 4) ? wz : 1.f - wz);
                const uint32_t idx = L.offset + idx3(L, cx + (c & 1), cy + ((c >> 1) & 1), cz + (c >> 2));
                red2(gflow + (size_t)idx * 2, w * g2.x, w * g2.y);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// collapsed gradient tables -> reference parameter layouts (transpose of the packing in field.cu)
// ------------------------------------------------------------------------------------------------
struct PlaneDst {
    uint32_t off[kPlScales][6];
};

// space planes: channel-last [R][R][8] -> += [8][R][R]
__global__ void k_expand_planes_static(const float* __restrict__ g, PlaneDst dst, uint32_t res,
                                       uint32_t scale, float* __restrict__ planes_grad) {
    const uint32_t combo = blockIdx.y == 0 ? 0u : (blockIdx.y == 1 ? 1u : 3u);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= res * res) return;
    float v[8];
    ld8(g + ((size_t)blockIdx.y * res * res + i) * 8, v);
    float* d = planes_grad + dst.off[scale][combo];
#pragma unroll
    for (int f = 0; f < 8; ++f) d[(size_t)f * res * res + i] += v[f];
}

// time planes: per query [R][8] -> += rows y0 / y1 of [8][Tres][R]
__global__ void k_expand_planes_dyn(const float* __restrict__ g, size_t per_q, PlaneDst dst,
                                    uint32_t res, uint32_t tres, uint32_t scale,
                                    const TimeInfo* __restrict__ ti, float* __restrict__ planes_grad) {
    const uint32_t p = blockIdx.y;
    const uint32_t combo = p == 0 ? 2u : (p == 1 ? 4u : 5u);
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= res) return;
    float* d = planes_grad + dst.off[scale][combo];
    for (int q = 0; q < 3; ++q) {   // sequential per thread: queries may share a time row
        if (!ti->valid[q]) continue;
        float v[8];
        ld8(g + q * per_q + ((size_t)p * res + x) * 8, v);
        const int y0 = ti->y0[q], y1 = ti->y1[q];
        const float w = ti->wy[q];
#pragma unroll
        for (int f = 0; f < 8; ++f) {
            d[((size_t)f * tres + y0) * res + x] += (1.0f - w) * v[f];
            d[((size_t)f * tres + y1) * res + x] += w * v[f];
        }
    }
}

// dynamic hash: grad[k][e][i] += blend_k * lag[0][i] * g[e] for the two blended time slices
__global__ void k_expand_dyn(const float* __restrict__ g, uint32_t entries,
                             const TimeInfo* __restrict__ ti, float* __restrict__ slices_grad) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= entries) return;
    const float v = __ldg(g + e);
    if (v == 0.f) return;
    const int k1 = ti->k1[0], k2 = ti->k2[0];
    const float w = ti->wk[0];
    const float l0 = ti->lag[0][0], l1 = ti->lag[0][1], l2 = ti->lag[0][2], l3 = ti->lag[0][3];
    float4* t = reinterpret_cast<float4*>(slices_grad);
    const float wa = k1 == k2 ? 1.0f : 1.0f - w;
    float4 a = t[(size_t)k1 * entries + e];
    a.x += wa * l0 * v; a.y += wa * l1 * v; a.z += wa * l2 * v; a.w += wa * l3 * v;
    t[(size_t)k1 * entries + e] = a;
    if (k1 != k2) {
        float4 b = t[(size_t)k2 * entries + e];
        b.x += w * l0 * v; b.y += w * l1 * v; b.z += w * l2 * v; b.w += w * l3 * v;
        t[(size_t)k2 * entries + e] = b;
    }
}

// flow grid: grad[e][2i+c] += lag[0][i] * g[e][c]
__global__ void k_expand_flow(const float2* __restrict__ g, uint32_t entries,
                              const TimeInfo* __restrict__ ti, float* __restrict__ grid_grad) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= entries) return;
    const float2 v = __ldg(g + e);
    if (v.x == 0.f && v.y == 0.f) return;
    const float l0 = ti->lag[0][0], l1 = ti->lag[0][1], l2 = ti->lag[0][2], l3 = ti->lag[0][3];
    float4* t = reinterpret_cast<float4*>(grid_grad) + 2 * (size_t)e;
    float4 a = t[0], b = t[1];
    a.x += l0 * v.x; a.y += l0 * v.y; a.z += l1 * v.x; a.w += l1 * v.y;
    b.x += l2 * v.x; b.y += l2 * v.y; b.z += l3 * v.x; b.w += l3 * v.y;
    t[0] = a; t[1] = b;
}

PlaneDst make_plane_dst(const nvsf_field_config_t* c) {
    PlaneDst s;
    uint32_t off = 0;
    const int comb[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    for (int sc = 0; sc < kPlScales; ++sc) {
        const uint32_t r[4] = {c->pl_res[sc], c->pl_res[sc], c->pl_res[sc], c->time_resolution};
        for (int k = 0; k < 6; ++k) {
            s.off[sc][k] = off;
            off += kPlaneF * r[comb[k][0]] * r[comb[k][1]];
        }
    }
    return s;
}

// ---- buffers ---------------------------------------------------------------------------------------
struct SavedLayout {   // per-sample intermediates kept by the training forward
    size_t sigma, geo, flow, feats, flowfeat, rgbs, split, total;
};
SavedLayout make_saved(size_t n) {
    SavedLayout L;
    size_t off = 0;
    L.sigma = off; off = ws_align(off + n * sizeof(float));
    L.geo = off; off = ws_align(off + n * kGeo * sizeof(__half));   // sigma + geo: the renderer's scratch layout
    L.flow = off; off = ws_align(off + n * 8 * sizeof(float));
    L.feats = off; off = ws_align(off + n * kFeat * sizeof(__half));
    L.flowfeat = off; off = ws_align(off + n * kFlowIn * sizeof(__half));
    L.rgbs = off; off = ws_align(off + n * 4 * sizeof(float));
    L.split = off; off = ws_align(off + nvsf_density_keep_scratch_bytes(n));  // density_mode 2 intermediates
    L.total = off;
    return L;
}

constexpr size_t kBwdChunkSamples = (size_t)6 << 20;  // samples per backward chunk (whole rays)

struct BwdLayout {
    size_t scales;                     // 4 x {unsigned max bits, pad, float scale, float 1/scale}
    size_t wimg;                       // fp16 weight images
    size_t wimg_tc;                    // 4 slots of operand images for the tcgen05 backward (flow, sigma, head a, head b)
    size_t g_pls, g_pld, g_dyn, g_flow; // collapsed gradient tables
    size_t tables_end;
    size_t dgeo16, dfeat, dflow, dflowfeat;
    size_t total;
    uint32_t chunk_rays;
};
BwdLayout make_bwd(const nvsf_field_config_t* cfg, uint32_t N, uint32_t S) {
    const WsLayout W = make_ws_layout(cfg);
    BwdLayout L;
    size_t off = 0;
    L.scales = off; off = ws_align(off + 64);
    L.wimg = off; off = ws_align(off + (size_t)kB_Total * sizeof(bf16));
    L.wimg_tc = off; off = ws_align(off + 4 * (size_t)mlptc::kImgSlot);
    L.g_pls = off; off = ws_align(off + W.pls_floats * sizeof(float));
    L.g_pld = off; off = ws_align(off + 3 * W.pld_floats_per_q * sizeof(float));
    L.g_dyn = off; off = ws_align(off + W.dyn_per_q * sizeof(float));
    L.g_flow = off; off = ws_align(off + (size_t)cfg->fl_entries * sizeof(float2));
    L.tables_end = off;
    const uint32_t cr = (uint32_t)std::max<size_t>(1, kBwdChunkSamples / std::max<uint32_t>(S, 1));
    L.chunk_rays = std::min<uint32_t>(N, cr);
    const size_t cs = (size_t)L.chunk_rays * S;
    L.dgeo16 = off; off = ws_align(off + cs * 16 * sizeof(float));
    L.dfeat = off; off = ws_align(off + cs * 128 * sizeof(float));
    L.dflow = off; off = ws_align(off + cs * 8 * sizeof(float));
    L.dflowfeat = off; off = ws_align(off + cs * 32 * sizeof(float));
    L.total = off;
    return L;
}

template <class T>
int launch_mlp_bwd(const typename T::Args& A, const bf16* w1, const bf16* w2, const bf16* wo, size_t n,
                   MlpGrads G, const float* scale2, int sms, cudaStream_t s) {
    static bool attr = false;
    static int per_sm = 1;
    constexpr size_t smem = mlp_bwd_smem<T>();
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_mlp_bwd<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return (int)e;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mlp_bwd<T>, kBwdWarps * 32, smem);
        if (per_sm < 1) per_sm = 1;
        attr = true;
    }
    const size_t tiles = (n + kBwdRows - 1) / kBwdRows;
    const int grid = (int)std::min<size_t>(tiles, (size_t)sms * per_sm);
    k_mlp_bwd<T><<<grid, kBwdWarps * 32, smem, s>>>(A, w1, w2, wo, n, G, scale2);
    return NVSF_OK;
}


int g_mlp_bwd_tc = 1;   // option "mlp_bwd_tc": MLP backward on tcgen05 (mlp_bwd_tc.cuh) instead of mma.sync (k_mlp_bwd)

template <class T, int W>
int launch_mlp_bwd_tc(const typename T::Args& A, const unsigned char* img, size_t n, MlpGrads G, const float* scale2,
                      int sms, cudaStream_t s) {
    static bool attr = false;
    constexpr size_t smem = mlptc::smem_bytes<T, W>();
    static_assert(smem <= 227 * 1024, "shared memory budget");
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(mlptc::k_mlp_bwd_tc<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    const size_t tiles = (n + 127) / 128;
    const int grid = (int)std::min<size_t>((tiles + W - 1) / W, (size_t)sms);
    mlptc::Grads g{G.w1, G.w2, G.wo};
    mlptc::k_mlp_bwd_tc<T, W><<<grid, W * 128, smem, s>>>(A, img, n, g, scale2);
    return NVSF_OK;
}

}  // namespace

extern "C" {

size_t nvsf_render_uniform_saved_bytes(uint32_t N, uint32_t S) {
    return make_saved((size_t)N * S).total;
}

size_t nvsf_render_uniform_backward_scratch_bytes(const nvsf_field_config_t* cfg, uint32_t N,
                                                  uint32_t S) {
    if (!field_cfg_ok(cfg)) return 0;
    return make_bwd(cfg, N, S).total;
}

void nvsf_render_uniform_debug_layout(const nvsf_field_config_t* cfg, uint32_t N, uint32_t S,
                                      size_t* out) {
    const SavedLayout SL = make_saved((size_t)N * S);
    const BwdLayout BL = make_bwd(cfg, N, S);
    const size_t v[12] = {SL.sigma, SL.geo, SL.flow, SL.feats, SL.flowfeat, SL.rgbs, BL.dgeo16, BL.dfeat,
                          BL.dflow, BL.dflowfeat, (size_t)BL.chunk_rays, BL.total};
    for (int i = 0; i < 12; ++i) out[i] = v[i];
}

int nvsf_render_uniform_train_forward(const nvsf_field_config_t* cfg, const void* workspace,
                                      uint32_t lidar, const float* rays_o, const float* rays_d,
                                      const float* nears, const float* fars, const float* noise,
                                      uint32_t N, uint32_t S, float bg_color, void* saved,
                                      size_t saved_bytes, float* depth, float* image,
                                      float* weights_sum, float* weights, float* z_vals,
                                      void* stream) {
    if (N == 0) return NVSF_OK;
    if (!field_cfg_ok(cfg) || !workspace || !rays_o || !rays_d || !nears || !fars || !saved ||
        !depth || !image || !weights_sum || !weights || !z_vals || S == 0)
        return NVSF_E_INVALID;
    const size_t n = (size_t)N * S;
    const SavedLayout L = make_saved(n);
    if (saved_bytes < L.total) return NVSF_E_WORKSPACE;
    unsigned char* sv = reinterpret_cast<unsigned char*>(saved);
    DensityKeep keep;
    keep.flow = reinterpret_cast<float*>(sv + L.flow);
    keep.feats = reinterpret_cast<__half*>(sv + L.feats);
    keep.flowfeat = reinterpret_cast<__half*>(sv + L.flowfeat);
    int st = nvsf_launch_density_split(cfg, workspace, nullptr, rays_o, rays_d, nears, fars, noise, S,
                                       n, reinterpret_cast<float*>(sv + L.sigma), sv + L.geo, nullptr,
                                       nullptr, sv + L.split, (cudaStream_t)stream, &keep);
    if (st != NVSF_OK) return st;
    return nvsf_render_composite_launch(cfg, workspace, lidar, rays_d, nears, fars, noise, N, S,
                                        bg_color, sv, L.flow, depth, image, weights_sum, weights,
                                        z_vals, sv + L.rgbs, stream);
}

int nvsf_render_uniform_backward(const nvsf_field_config_t* cfg, const void* workspace,
                                 const nvsf_field_params_t* params, uint32_t lidar,
                                 const float* rays_o, const float* rays_d, const float* nears,
                                 const float* fars, const float* noise, uint32_t N, uint32_t S,
                                 float bg_color, const void* saved, size_t saved_bytes,
                                 const float* weights, const float* g_depth, const float* g_image,
                                 const float* g_weights_sum, const float* g_weights,
                                 const nvsf_field_grads_t* grads, void* scratch,
                                 size_t scratch_bytes, void* stream) {
    if (N == 0) return NVSF_OK;
    if (!field_cfg_ok(cfg) || !workspace || !params || !rays_o || !rays_d || !nears || !fars ||
        !saved || !weights || !grads || !scratch || S == 0)
        return NVSF_E_INVALID;
    if (!grads->hash_static || !grads->hash_dynamic || !grads->planes || !grads->flow_grid ||
        !grads->flow_mlp || !grads->sigma_net || !grads->head_a || (lidar && !grads->head_b))
        return NVSF_E_INVALID;
    if (!params->flow_mlp || !params->sigma_net || !params->head_a || (lidar && !params->head_b))
        return NVSF_E_INVALID;
    const size_t n = (size_t)N * S;
    const SavedLayout SL = make_saved(n);
    if (saved_bytes < SL.total) return NVSF_E_WORKSPACE;
    const BwdLayout BL = make_bwd(cfg, N, S);
    if (scratch_bytes < BL.total) return NVSF_E_WORKSPACE;
    if (S > 96 * 1024) return NVSF_E_INVALID;  // per-warp chunk table of k_composite_bwd (48 KB of smem)
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned char* sv = reinterpret_cast<const unsigned char*>(saved);
    unsigned char* sc = reinterpret_cast<unsigned char*>(scratch);
    const FieldPtrs P = nvsf_make_field_ptrs(cfg, workspace);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

    // bf16 weight images
    bf16* wimg = reinterpret_cast<bf16*>(sc + BL.wimg);
    cudaMemsetAsync(wimg, 0, (size_t)kB_Total * sizeof(bf16), s);
    auto pack = [&](const float* src, int src_ld, int rows, int cols, bf16* dst, int dst_ld) {
        k_pack_matrix_bf16<<<nvsf_div_up(rows * cols, 256), 256, 0, s>>>(src, src_ld, rows, cols, dst, dst_ld);
    };
    const int H = kHidden;
    pack(params->flow_mlp, kFlowIn, H, kFlowIn, wimg + kB_FlowW1, kBLdK32);
    pack(params->flow_mlp + H * kFlowIn, H, H, H, wimg + kB_FlowW2, kLdH);
    pack(params->flow_mlp + H * kFlowIn + H * H, H, 6, H, wimg + kB_FlowW3, kLdH);
    pack(params->sigma_net, kFeat, H, kFeat, wimg + kB_SigW1, kBLdK128);
    pack(params->sigma_net + H * kFeat, H, kGeo, H, wimg + kB_SigW2, kLdH);
    const int kin = lidar ? 96 : 32, n_out = lidar ? 1 : 3;
    const float* heads[2] = {params->head_a, lidar ? params->head_b : nullptr};
    float* head_grads[2] = {grads->head_a, lidar ? grads->head_b : nullptr};
    for (int h = 0; h < 2; ++h) {
        if (!heads[h]) continue;
        bf16* hm = wimg + kB_Head + h * kB_HeadHalves;
        pack(heads[h], kin, H, kin, hm + kB_HeadW1, kin + 8);
        pack(heads[h] + H * kin, H, H, H, hm + kB_HeadW2, kLdH);
        pack(heads[h] + H * kin + H * H, H, n_out, H, hm + kB_HeadW3, kLdH);
    }
    // operand images of the tcgen05 backward
    unsigned char* img_tc = sc + BL.wimg_tc;
    if (g_mlp_bwd_tc) {
        mlptc::pack_images<FlowTc>(params->flow_mlp, params->flow_mlp + H * kFlowIn,
                                   params->flow_mlp + H * kFlowIn + H * H, img_tc, s);
        mlptc::pack_images<SigmaTc>(params->sigma_net, nullptr, params->sigma_net + H * kFeat,
                                    img_tc + mlptc::kImgSlot, s);
        for (int h = 0; h < 2; ++h) {
            if (!heads[h]) continue;
            unsigned char* slot = img_tc + (2 + h) * (size_t)mlptc::kImgSlot;
            if (lidar) mlptc::pack_images<HeadTc<true>>(heads[h], heads[h] + H * kin, heads[h] + H * kin + H * H, slot, s);
            else mlptc::pack_images<HeadTc<false>>(heads[h], heads[h] + H * kin, heads[h] + H * kin + H * H, slot, s);
        }
    }
    // collapsed gradient tables
    cudaMemsetAsync(sc + BL.g_pls, 0, BL.tables_end - BL.g_pls, s);
    GradTables G;
    G.pls = reinterpret_cast<float*>(sc + BL.g_pls);
    G.pld = reinterpret_cast<float*>(sc + BL.g_pld);
    G.dyn = reinterpret_cast<float*>(sc + BL.g_dyn);
    G.flow = reinterpret_cast<float*>(sc + BL.g_flow);
    G.hs = grads->hash_static;

    const float* sigma = reinterpret_cast<const float*>(sv + SL.sigma);
    const __half* geo = reinterpret_cast<const __half*>(sv + SL.geo);
    const float* flow = reinterpret_cast<const float*>(sv + SL.flow);
    const __half* feats = reinterpret_cast<const __half*>(sv + SL.feats);
    const __half* flowfeat = reinterpret_cast<const __half*>(sv + SL.flowfeat);
    const float* rgbs = reinterpret_cast<const float*>(sv + SL.rgbs);
    float* dgeo16 = reinterpret_cast<float*>(sc + BL.dgeo16);
    float* dfeat = reinterpret_cast<float*>(sc + BL.dfeat);
    float* dflow = reinterpret_cast<float*>(sc + BL.dflow);
    float* dflowfeat = reinterpret_cast<float*>(sc + BL.dflowfeat);
    const int nch = lidar ? 2 : 3;

    for (uint32_t r0 = 0; r0 < N; r0 += BL.chunk_rays) {
        const uint32_t nr = std::min<uint32_t>(BL.chunk_rays, N - r0);
        const size_t begin = (size_t)r0 * S, count = (size_t)nr * S;
        const float* noise_c = noise ? noise + begin : nullptr;
        // 1. compositing backward -> d sigma-logit
        {
            const unsigned blocks = nvsf_div_up(nr, 4u);
            const size_t cb_smem = (size_t)4 * ((S + 31) / 32) * sizeof(float);
            if (lidar)
                k_composite_bwd<true><<<blocks, 128, cb_smem, s>>>(
                    *cfg, nears + r0, fars + r0, noise_c, sigma + begin, geo + begin * kGeo,
                    rgbs + begin * 4, weights + begin, nr, S, bg_color, g_depth ? g_depth + r0 : nullptr,
                    g_image ? g_image + (size_t)r0 * nch : nullptr,
                    g_weights_sum ? g_weights_sum + r0 : nullptr, g_weights ? g_weights + begin : nullptr,
                    dgeo16);
            else
                k_composite_bwd<false><<<blocks, 128, cb_smem, s>>>(
                    *cfg, nears + r0, fars + r0, noise_c, sigma + begin, geo + begin * kGeo,
                    rgbs + begin * 4, weights + begin, nr, S, bg_color, g_depth ? g_depth + r0 : nullptr,
                    g_image ? g_image + (size_t)r0 * nch : nullptr,
                    g_weights_sum ? g_weights_sum + r0 : nullptr, g_weights ? g_weights + begin : nullptr,
                    dgeo16);
        }
        // power-of-two scales of the three fp16 MLP backward stages of this chunk
        unsigned char* scl = sc + BL.scales;
        cudaMemsetAsync(scl, 0, 64, s);
        auto make_scale = [&](int slot, const float* x, size_t cnt, float mul) {
            unsigned* mx = reinterpret_cast<unsigned*>(scl + 16 * slot);
            const unsigned blocks = (unsigned)std::min<size_t>(nvsf_div_up(cnt, (size_t)4096), (size_t)sms * 8);
            k_absmax<<<blocks, 256, 0, s>>>(x, cnt, mul, mx);
            k_make_scale<<<1, 1, 0, s>>>(mx, reinterpret_cast<float*>(mx) + 2);
            return reinterpret_cast<const float*>(mx) + 2;
        };
        // 2. colour heads backward -> d geo (cols 1..15 of dgeo16), head weight gradients
        int st = NVSF_OK;
        if (g_image) {
            const float* scale_h = make_scale(0, g_image + (size_t)r0 * nch, (size_t)nr * nch, shift_mul(g_shift_heads));
            for (int h = 0; h < (lidar ? 2 : 1) && st == NVSF_OK; ++h) {
                const bf16* hm = wimg + kB_Head + h * kB_HeadHalves;
                MlpGrads MG;
                MG.w1 = head_grads[h];
                MG.w2 = head_grads[h] + H * kin;
                MG.wo = head_grads[h] + H * kin + H * H;
                if (lidar) {
                    HeadT<true>::Args A{geo + begin * kGeo, rgbs + begin * 4, weights + begin,
                                        g_image + (size_t)r0 * 2, rays_d + (size_t)r0 * 3, dgeo16, S, h, h};
                    if (g_mlp_bwd_tc)
                        st = launch_mlp_bwd_tc<HeadTc<true>, 3>(A, img_tc + (2 + h) * (size_t)mlptc::kImgSlot, count, MG,
                                                                scale_h, sms, s);
                    else
                        st = launch_mlp_bwd<HeadT<true>>(A, hm + kB_HeadW1, hm + kB_HeadW2, hm + kB_HeadW3,
                                                         count, MG, scale_h, sms, s);
                } else {
                    HeadT<false>::Args A{geo + begin * kGeo, rgbs + begin * 4, weights + begin,
                                         g_image + (size_t)r0 * 3, rays_d + (size_t)r0 * 3, dgeo16, S, 0, 0};
                    if (g_mlp_bwd_tc)
                        st = launch_mlp_bwd_tc<HeadTc<false>, 4>(A, img_tc + 2 * (size_t)mlptc::kImgSlot, count, MG, scale_h,
                                                                 sms, s);
                    else
                        st = launch_mlp_bwd<HeadT<false>>(A, hm + kB_HeadW1, hm + kB_HeadW2, hm + kB_HeadW3,
                                                          count, MG, scale_h, sms, s);
                }
            }
        } else {
            cudaMemset2DAsync(dgeo16 + 1, 16 * sizeof(float), 0, 15 * sizeof(float), count, s);
        }
        if (st != NVSF_OK) return st;
        // 3. sigma net backward -> d features
        const float* scale_s = nullptr;
        {
            SigmaT::Args A{feats + begin * kFeat, dgeo16, dfeat};
            MlpGrads MG{grads->sigma_net, nullptr, grads->sigma_net + H * kFeat};
            scale_s = make_scale(1, dgeo16, count * 16, shift_mul(g_shift_sigma));
            if (g_mlp_bwd_tc)
                st = launch_mlp_bwd_tc<SigmaTc, 3>(A, img_tc + mlptc::kImgSlot, count, MG, scale_s, sms, s);
            else
                st = launch_mlp_bwd<SigmaT>(A, wimg + kB_SigW1, nullptr, wimg + kB_SigW2, count, MG, scale_s, sms, s);
            if (st != NVSF_OK) return st;
        }
        // 4. encoders backward -> table gradients, d flow
        {
            const unsigned blocks = (unsigned)nvsf_div_up(count, (size_t)256);
#define NVSF_ENC_BWD(C, ST)                                                                                     \
            k_encode_bwd<true, C, ST><<<blocks, 256, 0, s>>>(*cfg, P, G, nullptr, rays_o, rays_d, nears, fars, noise, \
                                                             S, begin, count, flow, dgeo16, dfeat, dflow, scale_s)
            if (g_enc_bwd_ctas == 3) { if (g_enc_bwd_h16) NVSF_ENC_BWD(3, true); else NVSF_ENC_BWD(3, false); }
            else { if (g_enc_bwd_h16) NVSF_ENC_BWD(2, true); else NVSF_ENC_BWD(2, false); }
#undef NVSF_ENC_BWD
        }
        // 5. flow MLP backward -> d flow-grid features; 6. flow-grid scatter
        {
            FlowT::Args A{flowfeat + begin * kFlowIn, dflow, dflowfeat};
            MlpGrads MG{grads->flow_mlp, grads->flow_mlp + H * kFlowIn, grads->flow_mlp + H * kFlowIn + H * H};
            const float* scale_f = make_scale(2, dflow, count * 8, shift_mul(g_shift_flow));
            if (g_mlp_bwd_tc)
                st = launch_mlp_bwd_tc<FlowTc, 4>(A, img_tc, count, MG, scale_f, sms, s);
            else
                st = launch_mlp_bwd<FlowT>(A, wimg + kB_FlowW1, wimg + kB_FlowW2, wimg + kB_FlowW3, count, MG,
                                           scale_f, sms, s);
            if (st != NVSF_OK) return st;
            const unsigned blocks = (unsigned)nvsf_div_up(count, (size_t)256);
            k_flowgrid_bwd<true><<<blocks, 256, 0, s>>>(*cfg, G.flow, nullptr, rays_o, rays_d, nears, fars,
                                                        noise, S, begin, count, dflow, dflowfeat, scale_f);
        }
    }
    // 7. collapsed tables -> parameter gradients
    const PlaneDst dst = make_plane_dst(cfg);
    const WsLayout W = make_ws_layout(cfg);
    for (int scl = 0; scl < kPlScales; ++scl) {
        const uint32_t R = cfg->pl_res[scl];
        k_expand_planes_static<<<dim3(nvsf_div_up(R * R, 256u), 3), 256, 0, s>>>(
            G.pls + W.pls_scale[scl], dst, R, scl, grads->planes);
        k_expand_planes_dyn<<<dim3(nvsf_div_up(R, 128u), 3), 128, 0, s>>>(
            G.pld + W.pld_scale[scl], W.pld_floats_per_q, dst, R, cfg->time_resolution, scl, P.ti,
            grads->planes);
    }
    {
        float* slices = grads->hash_dynamic;
        for (int p = 0; p < 3; ++p) {
            const uint32_t e = cfg->hd_entries[p];
            k_expand_dyn<<<nvsf_div_up(e, 256u), 256, 0, s>>>(G.dyn + W.dyn_plane[p], e, P.ti, slices);
            slices += (size_t)cfg->time_resolution * e * kHashF;
        }
    }
    k_expand_flow<<<nvsf_div_up(cfg->fl_entries, 256u), 256, 0, s>>>(
        reinterpret_cast<const float2*>(G.flow), cfg->fl_entries, P.ti, grads->flow_grid);
    return nvsf_launch_status();
}

// ---- NeRFNetwork.flow with autograd (scene-flow loss, trainer.py:237-265) --------------------------
// scratch: [scales 64 B][fp16 weight images][collapsed flow-grid gradient][qpos f32 [9][n] (forward) |
// dflowfeat f32 [n,32] (backward)]
size_t nvsf_field_flow_scratch_bytes(const nvsf_field_config_t* cfg, uint32_t n) {
    if (!field_cfg_ok(cfg)) return 0;
    return ws_align(64) + ws_align((size_t)kB_Total * sizeof(bf16)) +
           ws_align((size_t)cfg->fl_entries * sizeof(float2)) + ws_align((size_t)n * 32 * sizeof(float)) +
           ws_align((size_t)n * 9 * sizeof(float));
}

int nvsf_field_flow_forward(const nvsf_field_config_t* cfg, const void* workspace, const float* x,
                            uint32_t n, float* flow, void* flowfeat, void* scratch,
                            size_t scratch_bytes, void* stream) {
    if (n == 0) return NVSF_OK;
    if (!field_cfg_ok(cfg) || !workspace || !x || !flow || !flowfeat || !scratch) return NVSF_E_INVALID;
    if (scratch_bytes < nvsf_field_flow_scratch_bytes(cfg, n)) return NVSF_E_WORKSPACE;
    const FieldPtrs P = nvsf_make_field_ptrs(cfg, workspace);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned char* sc = reinterpret_cast<unsigned char*>(scratch);
    float* qpos = reinterpret_cast<float*>(sc + nvsf_field_flow_scratch_bytes(cfg, n) -
                                           ws_align((size_t)n * 9 * sizeof(float)));
    int st = nvsf_launch_flow_tc(cfg, P, x, nullptr, nullptr, nullptr, nullptr, nullptr, 1, 0, n, flow, qpos, n,
                                 reinterpret_cast<__half*>(flowfeat), sms, (cudaStream_t)stream);
    if (st != NVSF_OK) return st;
    return nvsf_launch_status();
}

int nvsf_field_flow_backward(const nvsf_field_config_t* cfg, const void* workspace,
                             const float* flow_mlp, const float* x, uint32_t n,
                             const void* flowfeat, const float* dflow, float* grads_flow_grid,
                             float* grads_flow_mlp, void* scratch, size_t scratch_bytes,
                             void* stream) {
    if (n == 0) return NVSF_OK;
    if (!field_cfg_ok(cfg) || !workspace || !flow_mlp || !x || !flowfeat || !dflow || !grads_flow_grid ||
        !grads_flow_mlp || !scratch)
        return NVSF_E_INVALID;
    if (scratch_bytes < nvsf_field_flow_scratch_bytes(cfg, n)) return NVSF_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const FieldPtrs P = nvsf_make_field_ptrs(cfg, workspace);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned char* sc = reinterpret_cast<unsigned char*>(scratch);
    size_t off = 0;
    unsigned char* scl = sc + off; off += ws_align(64);
    bf16* wimg = reinterpret_cast<bf16*>(sc + off); off += ws_align((size_t)kB_Total * sizeof(bf16));
    float* gflow = reinterpret_cast<float*>(sc + off); off += ws_align((size_t)cfg->fl_entries * sizeof(float2));
    float* dflowfeat = reinterpret_cast<float*>(sc + off);
    const int H = kHidden;
    cudaMemsetAsync(scl, 0, 64, s);
    cudaMemsetAsync(wimg, 0, (size_t)kB_SigW1 * sizeof(bf16), s);
    cudaMemsetAsync(gflow, 0, (size_t)cfg->fl_entries * sizeof(float2), s);
    auto pack = [&](const float* src, int src_ld, int rows, int cols, bf16* dst, int dst_ld) {
        k_pack_matrix_bf16<<<nvsf_div_up(rows * cols, 256), 256, 0, s>>>(src, src_ld, rows, cols, dst, dst_ld);
    };
    pack(flow_mlp, kFlowIn, H, kFlowIn, wimg + kB_FlowW1, kBLdK32);
    pack(flow_mlp + H * kFlowIn, H, H, H, wimg + kB_FlowW2, kLdH);
    pack(flow_mlp + H * kFlowIn + H * H, H, 6, H, wimg + kB_FlowW3, kLdH);
    unsigned* mx = reinterpret_cast<unsigned*>(scl);
    k_absmax<<<(unsigned)std::min<size_t>(nvsf_div_up((size_t)n * 8, (size_t)1024), (size_t)sms * 4), 256, 0, s>>>(
        dflow, (size_t)n * 8, shift_mul(g_shift_flow), mx);
    k_make_scale<<<1, 1, 0, s>>>(mx, reinterpret_cast<float*>(mx) + 2);
    const float* scale_f = reinterpret_cast<const float*>(mx) + 2;
    FlowT::Args A{reinterpret_cast<const __half*>(flowfeat), dflow, dflowfeat};
    MlpGrads MG{grads_flow_mlp, grads_flow_mlp + H * kFlowIn, grads_flow_mlp + H * kFlowIn + H * H};
    int st = launch_mlp_bwd<FlowT>(A, wimg + kB_FlowW1, wimg + kB_FlowW2, wimg + kB_FlowW3, n, MG, scale_f, sms, s);
    if (st != NVSF_OK) return st;
    k_flowgrid_bwd<false><<<(unsigned)nvsf_div_up((size_t)n, (size_t)256), 256, 0, s>>>(
        *cfg, gflow, x, nullptr, nullptr, nullptr, nullptr, nullptr, 1, 0, n, dflow, dflowfeat, scale_f);
    k_expand_flow<<<nvsf_div_up(cfg->fl_entries, 256u), 256, 0, s>>>(reinterpret_cast<const float2*>(gflow),
                                                                   cfg->fl_entries, P.ti, grads_flow_grid);
    return nvsf_launch_status();
}

}  // extern "C"

// ---- tuning options of the backward kernels (nvsf_split_set_option falls through to here) --------
int nvsf_train_set_option(const char* name, int value) {
    if (std::string(name) == "enc_bwd_ctas") {
        if (value != 2 && value != 3) return NVSF_E_INVALID;
        g_enc_bwd_ctas = value;
        return NVSF_OK;
    }
    if (std::string(name) == "enc_bwd_h16") {
        if (value != 0 && value != 1) return NVSF_E_INVALID;
        g_enc_bwd_h16 = value;
        return NVSF_OK;
    }
    if (std::string(name) == "mlp_bwd_tc") {
        if (value != 0 && value != 1) return NVSF_E_INVALID;
        g_mlp_bwd_tc = value;
        return NVSF_OK;
    }
    for (auto kv : {std::make_pair("bwd_shift_flow", &g_shift_flow), std::make_pair("bwd_shift_sigma", &g_shift_sigma),
                    std::make_pair("bwd_shift_heads", &g_shift_heads)})
        if (std::string(name) == kv.first) {
            if (value < -4 || value > 10) return NVSF_E_INVALID;
            *kv.second = value;
            return NVSF_OK;
        }
    return nvsf_render_set_option(name, value);
}
int nvsf_train_get_option(const char* name) {
    if (std::string(name) == "enc_bwd_ctas") return g_enc_bwd_ctas;
    if (std::string(name) == "enc_bwd_h16") return g_enc_bwd_h16;
    if (std::string(name) == "mlp_bwd_tc") return g_mlp_bwd_tc;
    if (std::string(name) == "bwd_shift_flow") return g_shift_flow;
    if (std::string(name) == "bwd_shift_sigma") return g_shift_sigma;
    if (std::string(name) == "bwd_shift_heads") return g_shift_heads;
    return nvsf_render_get_option(name);
}
