// nvsf_b200 — NeRFNetwork.color as a standalone operator (sm_100a).
//
// Replaces reference nvsf/nerf/models/network_dynamic.py:290-332: per sample, the view direction
// is encoded (LiDAR: tcnn Frequency with 12 octaves of (d+1)/2 -> 72 values; camera: tcnn
// SphericalHarmonics degree 4 -> 16 values), concatenated with the 15 geometry features and run
// through intensity_net + raydrop_net (87->64->64->1 each, output order [raydrop, intensity]) or
// color_net (31->64->64->3), then a sigmoid.  Samples whose mask is false produce zeros.
//
// Unlike the uniform renderer (render.cu), where the direction is constant along a ray and its
// part of layer 1 is hoisted out of the sample loop, here every sample carries its own direction
// (the marched samples of march_rays_train / march_rays), so the whole 96- (32-) wide layer 1 runs
// on the tensor cores.  One warp owns a tile of 32 samples: each lane encodes its sample's
// direction into a shared-memory row, the warp runs the three layers with mma.sync m16n8k16
// (hidden activations never leave registers) and the colours go out through shared memory as one
// coalesced store.  Tiles without any masked-in sample only write zeros.
#include <algorithm>

#include "umma.cuh"

namespace {

constexpr int kCWarps = 4;

template <bool LIDAR>
struct ColorDims {
    static constexpr int kDirPad = LIDAR ? 80 : 16;         // 72 -> 80 (k-tiles of 16)
    static constexpr int kK = kDirPad + 16;                 // + geo16 (col 0: logit -> constant 1)
    static constexpr int kLd = kK + 8;                      // 104 / 40 halves: conflict-free ldmatrix
    static constexpr int kNets = LIDAR ? 2 : 1;
    static constexpr size_t kSmem =
        (size_t)kNets * kHeadHalves * sizeof(__half) + (size_t)kCWarps * 32 * kLd * sizeof(__half);
};

__device__ __forceinline__ float sigmoid_c(float h) { return 1.0f / (1.0f + expf(-h)); }

template <bool LIDAR>
__global__ void __launch_bounds__(kCWarps * 32)
k_field_color(const __half* __restrict__ mlp, const float* __restrict__ dirs,
              const __half* __restrict__ geo, uint32_t geo_ld, uint32_t geo_off,
              const uint8_t* __restrict__ mask, size_t n, float* __restrict__ out,
              uint32_t out_ld) {
    using D = ColorDims<LIDAR>;
    constexpr int NETS = D::kNets;
    constexpr int NCH = LIDAR ? 2 : 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* Wsm = reinterpret_cast<__half*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __half* At = reinterpret_cast<__half*>(smem_raw + (size_t)NETS * kHeadHalves * sizeof(__half)) +
                 (size_t)warp * 32 * D::kLd;
    block_copy16(Wsm, mlp + kHeadBase, NETS * kHeadHalves * (int)sizeof(__half) / 16, tid,
                 kCWarps * 32);
    __syncthreads();

    const int gq = lane >> 2, tq = lane & 3;
    const bool fast_geo = geo_ld == 16 && geo_off == 1;  // rows of the density kernel's geo16
    const size_t n_tiles = (n + 31) / 32;
    for (size_t tile = (size_t)blockIdx.x * kCWarps + warp; tile < n_tiles;
         tile += (size_t)gridDim.x * kCWarps) {
        const size_t g = tile * 32 + lane;
        const bool in = g < n;
        const bool m = in && (mask == nullptr || mask[g] != 0);
        if (__ballot_sync(0xffffffffu, m) == 0) {
            if (in)
                for (uint32_t c = 0; c < out_ld; ++c) out[g * out_ld + c] = 0.f;
            continue;
        }
        // ---- stage this lane's input row: [dir encoding | 0 pad | geo16 with col 0 = 1] ----
        __half* row = At + lane * D::kLd;
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (m) {
            dx = __ldg(dirs + g * 3); dy = __ldg(dirs + g * 3 + 1); dz = __ldg(dirs + g * 3 + 2);
        }
        if (LIDAR) {
            // tcnn Frequency: out[dim*24 + 2*oct + p] = sin(2^oct * pi * x + p*pi/2), x = (d+1)/2
#pragma unroll
            for (int dim = 0; dim < 3; ++dim) {
                const float v = ((dim == 0 ? dx : (dim == 1 ? dy : dz)) + 1.0f) * 0.5f;
#pragma unroll
                for (int oct = 0; oct < 12; oct += 4) {
                    float s[4], c[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) sincospif(scalbnf(v, oct + k), &s[k], &c[k]);
                    uint4 o;
                    o.x = pack_half2(s[0], c[0]); o.y = pack_half2(s[1], c[1]);
                    o.z = pack_half2(s[2], c[2]); o.w = pack_half2(s[3], c[3]);
                    *reinterpret_cast<uint4*>(row + dim * 24 + oct * 2) = o;
                }
            }
            *reinterpret_cast<uint4*>(row + 72) = make_uint4(0, 0, 0, 0);
        } else {
            // tcnn SphericalHarmonics degree 4 of 2*((d+1)/2) - 1
            const float x = ((dx + 1.0f) * 0.5f) * 2.0f - 1.0f, y = ((dy + 1.0f) * 0.5f) * 2.0f - 1.0f,
                        z = ((dz + 1.0f) * 0.5f) * 2.0f - 1.0f;
            const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
            float v[16];
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
            float lo[8], hi[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { lo[k] = v[k]; hi[k] = v[8 + k]; }
            st8(row, 0, lo);
            st8(row, 8, hi);
        }
        {
            uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
            if (m) {
                if (fast_geo) {
                    const uint4* src = reinterpret_cast<const uint4*>(geo + g * 16);
                    a0 = __ldg(src); a1 = __ldg(src + 1);
                    a0.x = (a0.x & 0xffff0000u) | kOneH;  // column 0 (the sigma logit) becomes the constant-1 padding input
                } else {
                    __align__(16) __half h[16];
                    h[0] = __float2half(1.f);
#pragma unroll
                    for (int k = 0; k < 15; ++k) h[1 + k] = geo[g * geo_ld + geo_off + k];
                    a0 = *reinterpret_cast<const uint4*>(h);
                    a1 = *reinterpret_cast<const uint4*>(h + 8);
                }
            }
            uint4* dst = reinterpret_cast<uint4*>(row + D::kDirPad);
            dst[0] = a0; dst[1] = a1;
        }
        __syncwarp();
        constexpr int KTD = D::kDirPad / 16;
        uint32_t ad[2][KTD][4], ag[2][1][4];
        load_a_frags<KTD>(At, D::kLd, ad, lane);
        load_a_frags<1>(At + D::kDirPad, D::kLd, ag, lane);
        __syncwarp();
        float* cs = reinterpret_cast<float*>(At);  // [32][4] colour staging (tile is in registers)
        reinterpret_cast<float4*>(cs)[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
#pragma unroll
        for (int net = 0; net < NETS; ++net) {
            const __half* Wn = Wsm + net * kHeadHalves;
            float acc[2][8][4];
            zero_acc<8>(acc);
            // direction columns: W1d rows are kHeadDirMax (72) halves apart; for LiDAR the fifth
            // k-tile reads 8 halves past the row end (finite weights) against zero A columns
            warp_gemm_regA<KTD, 8>(ad, Wn + kHeadW1d, kHeadDirMax, acc, lane);
            warp_gemm_regA<1, 8>(ag, Wn + kHeadW1g, kLdK16, acc, lane);
            uint32_t a2[2][4][4];
            relu_to_a<8>(acc, a2);
            zero_acc<8>(acc);
            warp_gemm_regA<4, 8>(a2, Wn + kHeadW2, kLdK64, acc, lane);
            relu_to_a<8>(acc, a2);
            float o[2][1][4];
            zero_acc<1>(o);
            warp_gemm_regA<4, 1>(a2, Wn + kHeadW3, kLdK64, o, lane);
            // rows gq, gq+8 (m-tile 0), gq+16, gq+24 (m-tile 1); output columns 2tq, 2tq+1
            if (LIDAR) {
                if (tq == 0) {
                    const int ch = net == 0 ? 1 : 0;  // [raydrop, intensity], network_dynamic.py:317
                    cs[gq * 4 + ch] = sigmoid_c(o[0][0][0]);
                    cs[(gq + 8) * 4 + ch] = sigmoid_c(o[0][0][2]);
                    cs[(gq + 16) * 4 + ch] = sigmoid_c(o[1][0][0]);
                    cs[(gq + 24) * 4 + ch] = sigmoid_c(o[1][0][2]);
                }
            } else if (tq < 2) {
                const int c0 = 2 * tq;
                cs[gq * 4 + c0] = sigmoid_c(o[0][0][0]);
                cs[(gq + 8) * 4 + c0] = sigmoid_c(o[0][0][2]);
                cs[(gq + 16) * 4 + c0] = sigmoid_c(o[1][0][0]);
                cs[(gq + 24) * 4 + c0] = sigmoid_c(o[1][0][2]);
                if (tq == 0) {
                    cs[gq * 4 + 1] = sigmoid_c(o[0][0][1]);
                    cs[(gq + 8) * 4 + 1] = sigmoid_c(o[0][0][3]);
                    cs[(gq + 16) * 4 + 1] = sigmoid_c(o[1][0][1]);
                    cs[(gq + 24) * 4 + 1] = sigmoid_c(o[1][0][3]);
                }
            }
        }
        __syncwarp();
        if (in) {
            const float4 c = reinterpret_cast<const float4*>(cs)[lane];
            const float v[4] = {c.x, c.y, c.z, 0.f};
            for (uint32_t k = 0; k < out_ld; ++k)
                out[g * out_ld + k] = (m && k < (uint32_t)NCH) ? v[k] : 0.f;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// color_net (camera: SH-4 | geo -> 64 -> 64 -> 3) on tcgen05.mma / TMEM.  One CTA per SM, eight
// independent warpgroups; a warpgroup owns 128-sample tiles (thread = sample = UMMA row = TMEM lane):
// the thread encodes its direction and writes [SH 16 | geo 16] as the first 64 bytes of its row of
// the swizzled K-major operand tile, then three MMA batches (K = 32, 64, 64) with the hidden
// activations going TMEM -> tcgen05.ld -> relu/fp16 -> the same tile.  Operand images: the heads
// image of render.cu (camera workspaces carry the whole first layer in its net-1 rows).
// The mma.sync kernel above spends its time in ldmatrix / HMMA issue at 5.7 G samples/s.
// ------------------------------------------------------------------------------------------------
using namespace umma;
constexpr int kKWG = 8;
constexpr int kKThreads = kKWG * kRows;
constexpr uint32_t kKImg = 128 * 128 + 2 * kHidden * 128 + 2 * 16 * 128;   // render.cu kHImgBytes
constexpr uint32_t kKOffW1 = kHidden * 128;                                // net-1 rows: full layer 1
constexpr uint32_t kKOffW2 = 128 * 128, kKOffW3 = kKOffW2 + 2 * kHidden * 128;
constexpr uint32_t kKOffTiles = kKImg;
constexpr uint32_t kKOffBar = kKOffTiles + kKWG * kRows * 128;
constexpr size_t kColorTcSmem = kKOffBar + 8 * kKWG + 16 + 1024;
static_assert(kKImg % 1024 == 0, "swizzled tiles start on 1024-byte boundaries");

__global__ void __launch_bounds__(kKThreads, 1)
k_color_tc(const unsigned char* __restrict__ wimg, const float* __restrict__ dirs,
           const __half* __restrict__ geo, uint32_t geo_ld, uint32_t geo_off,
           const uint8_t* __restrict__ mask, size_t n, float* __restrict__ out, uint32_t out_ld) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kKOffBar + 8 * kKWG);
    const uint32_t tid = threadIdx.x, wg = tid >> 7, t = tid & 127u;
    for (uint32_t i = tid; i < kKImg / 16; i += kKThreads)
        reinterpret_cast<uint4*>(sm)[i] = __ldg(reinterpret_cast<const uint4*>(wimg) + i);
    if (tid < (uint32_t)kKWG) mbar_init(base + kKOffBar + 8 * tid, 1);
    if (tid < 32) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tcol = tmem + wg * 64u;
    const uint32_t tlane = tcol + ((((tid >> 5) & 3u) * 32u) << 16);
    const uint32_t xs = base + kKOffTiles + wg * (kRows * 128);
    unsigned char* xg = sm + kKOffTiles + wg * (kRows * 128);
    const uint32_t bar = base + kKOffBar + 8 * wg;
    constexpr uint32_t kIdesc64 = umma_idesc(kRows, kHidden), kIdesc16 = umma_idesc(kRows, 16);
    const bool fast_geo = geo_ld == 16 && geo_off == 1;
    uint32_t phase = 0;
    const size_t n_tiles = (n + kRows - 1) / kRows;
    for (size_t tile = (size_t)blockIdx.x * kKWG + wg; tile < n_tiles; tile += (size_t)gridDim.x * kKWG) {
        const size_t g = tile * kRows + t;
        const bool in = g < n;
        const bool m = in && (mask == nullptr || mask[g] != 0);
        if (!wg_any(wg, m)) {
            if (in)
                for (uint32_t c = 0; c < out_ld; ++c) out[g * out_ld + c] = 0.f;
            continue;
        }
        // ---- this thread's input row: [SH-4 of the direction | geo16 with column 0 = 1] ----
        {
            float dx = 0.f, dy = 0.f, dz = 0.f;
            if (m) { dx = __ldg(dirs + g * 3); dy = __ldg(dirs + g * 3 + 1); dz = __ldg(dirs + g * 3 + 2); }
            // tcnn SphericalHarmonics degree 4 of 2*((d+1)/2) - 1 (same expressions as k_field_color)
            const float x = ((dx + 1.0f) * 0.5f) * 2.0f - 1.0f, y = ((dy + 1.0f) * 0.5f) * 2.0f - 1.0f,
                        z = ((dz + 1.0f) * 0.5f) * 2.0f - 1.0f;
            const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
            float v[8];
            v[0] = 0.28209479177387814f;
            v[1] = -0.48860251190291987f * y;
            v[2] = 0.48860251190291987f * z;
            v[3] = -0.48860251190291987f * x;
            v[4] = 1.0925484305920792f * xy;
            v[5] = -1.0925484305920792f * yz;
            v[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
            v[7] = -1.0925484305920792f * xz;
            st_chunk(xg, t, 0, v);
            v[0] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
            v[1] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
            v[2] = 2.8906114426405538f * xy * z;
            v[3] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
            v[4] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
            v[5] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
            v[6] = 1.4453057213202769f * z * (x2 - y2);
            v[7] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
            st_chunk(xg, t, 1, v);
            uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
            if (m) {
                if (fast_geo) {
                    const uint4* src = reinterpret_cast<const uint4*>(geo + g * 16);
                    a0 = __ldg(src); a1 = __ldg(src + 1);
                    a0.x = (a0.x & 0xffff0000u) | kOneH;  // column 0 (the sigma logit) becomes the constant-1 padding input
                } else {
                    __align__(16) __half h[16];
                    h[0] = __float2half(1.f);
#pragma unroll
                    for (int k = 0; k < 15; ++k) h[1 + k] = geo[g * geo_ld + geo_off + k];
                    a0 = *reinterpret_cast<const uint4*>(h);
                    a1 = *reinterpret_cast<const uint4*>(h + 8);
                }
            }
            *reinterpret_cast<uint4*>(xg + swz(t, 2)) = a0;
            *reinterpret_cast<uint4*>(xg + swz(t, 3)) = a1;
        }
        fence_async_smem();
        tc_fence_before();
        wg_barrier(wg);
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (uint32_t k = 0; k < 2; ++k)
                umma_f16(tcol, umma_desc(xs + k * 32), umma_desc(base + kKOffW1 + k * 32), kIdesc64, k);
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
            // H = relu(D) -> tile; next batch: layer 2 (64 -> 64) or the output layer (64 -> 16)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t v[16];
                tmem_ld16(tlane + q * 16, v);
                tmem_ld_wait();
                st_chunk_relu(xg, t, 2 * q, v);
                st_chunk_relu(xg, t, 2 * q + 1, v + 8);
            }
            fence_async_smem();
            tc_fence_before();
            wg_barrier(wg);
            if (t == 0) {
                tc_fence_after();
                const uint32_t w = base + (layer == 0 ? kKOffW2 : kKOffW3);
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                    umma_f16(tcol, umma_desc(xs + k * 32), umma_desc(w + k * 32), layer == 0 ? kIdesc64 : kIdesc16, k);
                umma_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            tc_fence_after();
        }
        {
            uint32_t o[16];
            tmem_ld16(tlane, o);
            tmem_ld_wait();
            if (in) {
                const float c[4] = {sigmoid_c(__uint_as_float(o[0])), sigmoid_c(__uint_as_float(o[1])),
                                    sigmoid_c(__uint_as_float(o[2])), 0.f};
                for (uint32_t k = 0; k < out_ld; ++k) out[g * out_ld + k] = (m && k < 3u) ? c[k] : 0.f;
            }
        }
        tc_fence_before();   // the next tile's first MMA is issued behind a warpgroup barrier
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u)
                     : "memory");
}

bool g_color_tc_attr = false;

bool g_color_attr = false;
int ensure_color_attrs() {
    if (g_color_attr) return NVSF_OK;
    cudaError_t e = cudaFuncSetAttribute(k_field_color<true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)ColorDims<true>::kSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_field_color<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)ColorDims<false>::kSmem);
    if (e != cudaSuccess) return (int)e;
    g_color_attr = true;
    return NVSF_OK;
}

}  // namespace

extern "C" int nvsf_field_color(const nvsf_field_config_t* cfg, const void* workspace,
                                uint32_t lidar, const float* dirs, const void* geo,
                                uint32_t geo_ld, uint32_t geo_off, const uint8_t* mask, uint32_t n,
                                float* out, uint32_t out_ld, void* stream) {
    if (n == 0) return NVSF_OK;
    const uint32_t nch = lidar ? 2u : 3u;
    if (!field_cfg_ok(cfg) || !workspace || !dirs || !geo || !out || out_ld < nch || out_ld > 4 ||
        geo_ld < geo_off + 15)
        return NVSF_E_INVALID;
    if (geo_ld == 16 && geo_off == 1 && (reinterpret_cast<uintptr_t>(geo) & 15) != 0)
        return NVSF_E_INVALID;
    int st = ensure_color_attrs();
    if (st != NVSF_OK) return st;
    const FieldPtrs P = nvsf_make_field_ptrs(cfg, workspace);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t tiles = ((size_t)n + 31) / 32;
    const uint32_t blocks =
        (uint32_t)std::min<size_t>((tiles + kCWarps - 1) / kCWarps, (size_t)sms * 4);
    cudaStream_t s = (cudaStream_t)stream;
    if (!lidar && nvsf_heads_tc()) {   // color_net on tcgen05 (option "heads_tc")
        if (!g_color_tc_attr) {
            cudaError_t e = cudaFuncSetAttribute(k_color_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)kColorTcSmem);
            if (e != cudaSuccess) return (int)e;
            g_color_tc_attr = true;
        }
        const size_t tiles128 = ((size_t)n + kRows - 1) / kRows;
        const uint32_t grid = (uint32_t)std::min<size_t>((tiles128 + kKWG - 1) / kKWG, (size_t)sms);
        k_color_tc<<<grid, kKThreads, kColorTcSmem, s>>>(reinterpret_cast<const unsigned char*>(P.heads_tc), dirs,
                                                        reinterpret_cast<const __half*>(geo), geo_ld, geo_off,
                                                        mask, n, out, out_ld);
        return nvsf_launch_status();
    }
    if (lidar)
        k_field_color<true><<<blocks, kCWarps * 32, ColorDims<true>::kSmem, s>>>(
            P.mlp, dirs, reinterpret_cast<const __half*>(geo), geo_ld, geo_off, mask, n, out, out_ld);
    else
        k_field_color<false><<<blocks, kCWarps * 32, ColorDims<false>::kSmem, s>>>(
            P.mlp, dirs, reinterpret_cast<const __half*>(geo), geo_ld, geo_off, mask, n, out, out_ld);
    return nvsf_launch_status();
}
