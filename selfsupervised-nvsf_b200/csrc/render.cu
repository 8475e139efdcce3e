// nvsf_b200 — uniform-sample renderer: alpha compositing fused with the colour heads (sm_100a).
//
// Replaces the tail of NeRFRenderer.run (reference nvsf/nerf/models/renderer_dynamic.py:181-237:
// deltas, alphas, cumprod weights, w > 1e-4 mask, masked colour query, weights_sum / depth / image
// accumulation, background) and NeRFNetwork.color (network_dynamic.py:290-332: Frequency / SH
// direction encoding, intensity + raydrop nets or colour net, sigmoid).
//
// One warp owns one ray.  The direction encoding is constant along a ray, so its contribution to
// the first layer of the head MLPs is computed once per ray (u = W1[:, dir] * enc(d)) and used to
// initialise the tensor-core accumulators; per sample only the 15 geometry features go through
// layer 1.  Transmittance is an exclusive product scan over 32-sample chunks (warp shuffles) with
// a running carry; chunks in which no sample passes the w > 1e-4 mask skip the heads entirely.
#include <algorithm>

#include "field_common.cuh"

namespace {

constexpr int kRWarps = 4;
// resident CTAs per SM the register allocation is tuned for: 4 -> <= 128 registers (16 warps/SM,
// 24 bytes of spills) instead of 150 registers / 12 warps; the kernel is latency bound (ncu: stall
// `wait` + long_scoreboard at 14 % occupancy, profiles/r01_composite_v1_ncu_full.txt)
#ifndef NVSF_RENDER_MIN_CTAS
#define NVSF_RENDER_MIN_CTAS 4
#endif
constexpr int kGeoLd = kLdK16;  // 24 halves per staged geo row
constexpr int kWarpScratchBytes = 32 * kGeoLd * 2 + 2 * kHidden * 4 + kHeadDirMax * 4;

template <bool LIDAR>
constexpr size_t render_smem() {
    return (size_t)(LIDAR ? 2 : 1) * kHeadHalves * sizeof(__half) +
           (size_t)kRWarps * kWarpScratchBytes;
}

__device__ __forceinline__ float uniform_z2(float near, float far, uint32_t k, uint32_t S,
                                            const float* __restrict__ noise, size_t g) {
    const float step = 1.0f / (float)(S > 1 ? S - 1 : 1);
    const float lin = (k < S / 2) ? step * (float)k : 1.0f - step * (float)(S - 1 - k);
    float z = near + (far - near) * lin;
    if (noise) z = z + (__ldg(noise + g) - 0.5f) * ((far - near) / (float)S);
    return z;
}

__device__ __forceinline__ float sigmoidf_(float h) { return 1.0f / (1.0f + expf(-h)); }

template <bool LIDAR>
__global__ void __launch_bounds__(kRWarps * 32, NVSF_RENDER_MIN_CTAS)
k_render_composite(const __grid_constant__ nvsf_field_config_t cfg, const __half* __restrict__ mlp,
                   const float* __restrict__ rays_d, const float* __restrict__ nears,
                   const float* __restrict__ fars, const float* __restrict__ noise,
                   const float* __restrict__ sigma, const __half* __restrict__ geo, uint32_t N,
                   uint32_t S, float bg_color, float* __restrict__ depth_out,
                   float* __restrict__ image_out, float* __restrict__ ws_out,
                   float* __restrict__ weights_out, float* __restrict__ z_out,
                   float* __restrict__ rgbs_out /* [N*S,4] kept for the backward pass, or NULL */) {
    constexpr int NETS = LIDAR ? 2 : 1;
    constexpr int NDIR = LIDAR ? 72 : 16;
    constexpr int NCH = LIDAR ? 2 : 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* Wsm = reinterpret_cast<__half*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* wscr = smem_raw + (size_t)NETS * kHeadHalves * sizeof(__half) +
                          (size_t)warp * kWarpScratchBytes;
    __half* geo_s = reinterpret_cast<__half*>(wscr);
    float* u_s = reinterpret_cast<float*>(wscr + 32 * kGeoLd * 2);
    float* enc_s = u_s + 2 * kHidden;

    block_copy16(Wsm, mlp + kHeadBase, NETS * kHeadHalves * (int)sizeof(__half) / 16, tid,
                 kRWarps * 32);
    __syncthreads();

    const int gq = lane >> 2, tq = lane & 3;
    const float kexp = cfg.active_sensor ? 2.0f : 1.0f;

    for (uint32_t r = blockIdx.x * kRWarps + warp; r < N; r += gridDim.x * kRWarps) {
        const float near = __ldg(nears + r), far = __ldg(fars + r);
        const float dx = __ldg(rays_d + (size_t)r * 3), dy = __ldg(rays_d + (size_t)r * 3 + 1),
                    dz = __ldg(rays_d + (size_t)r * 3 + 2);
        // ---- direction encoding (network_dynamic.py:310-311 / 319-320) ----
        __syncwarp();
        if (LIDAR) {
            // tcnn Frequency, 12 octaves: sin(2^k pi x + (j&1) pi/2), x = (d+1)/2
            for (int j = lane; j < NDIR; j += 32) {
                const int dim = j / 24, oct = (j >> 1) % 12;
                const float v = ((dim == 0 ? dx : (dim == 1 ? dy : dz)) + 1.0f) * 0.5f;
                enc_s[j] = sinpif(scalbnf(v, oct) + 0.5f * (float)(j & 1));
            }
        } else if (lane < 16) {
            // tcnn SphericalHarmonics degree 4 of 2*((d+1)/2)-1
            const float x = ((dx + 1.0f) * 0.5f) * 2.0f - 1.0f, y = ((dy + 1.0f) * 0.5f) * 2.0f - 1.0f,
                        z = ((dz + 1.0f) * 0.5f) * 2.0f - 1.0f;
            const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
            float v;
            switch (lane) {
                case 0: v = 0.28209479177387814f; break;
                case 1: v = -0.48860251190291987f * y; break;
                case 2: v = 0.48860251190291987f * z; break;
                case 3: v = -0.48860251190291987f * x; break;
                case 4: v = 1.0925484305920792f * xy; break;
                case 5: v = -1.0925484305920792f * yz; break;
                case 6: v = 0.94617469575755997f * z2 - 0.31539156525251999f; break;
                case 7: v = -1.0925484305920792f * xz; break;
                case 8: v = 0.54627421529603959f * x2 - 0.54627421529603959f * y2; break;
                case 9: v = 0.59004358992664352f * y * (-3.0f * x2 + y2); break;
                case 10: v = 2.8906114426405538f * xy * z; break;
                case 11: v = 0.45704579946446572f * y * (1.0f - 5.0f * z2); break;
                case 12: v = 0.3731763325901154f * z * (5.0f * z2 - 3.0f); break;
                case 13: v = 0.45704579946446572f * x * (1.0f - 5.0f * z2); break;
                case 14: v = 1.4453057213202769f * z * (x2 - y2); break;
                default: v = 0.59004358992664352f * x * (-x2 + 3.0f * y2); break;
            }
            enc_s[lane] = v;
        }
        __syncwarp();
        // ---- per-ray part of layer 1: u[net][n] = sum_j W1[n][j] * enc[j] ----
#pragma unroll
        for (int net = 0; net < NETS; ++net) {
            const __half* W1d = Wsm + net * kHeadHalves + kHeadW1d;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int nrow = lane + 32 * h;
                float acc = 0.f;
                for (int j = 0; j < NDIR; j += 2) {
                    const float2 w = __half22float2(
                        *reinterpret_cast<const __half2*>(W1d + nrow * kHeadDirMax + j));
                    acc = fmaf(w.x, enc_s[j], acc);
                    acc = fmaf(w.y, enc_s[j + 1], acc);
                }
                u_s[net * kHidden + nrow] = acc;
            }
        }
        __syncwarp();

        float carry = 1.0f, ws = 0.f, dep = 0.f, img[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) img[c] = 0.f;

        for (uint32_t c0 = 0; c0 < S; c0 += 32) {
            const uint32_t i = c0 + lane;
            const bool in = i < S;
            const size_t g = (size_t)r * S + (in ? i : S - 1);
            const float z = uniform_z2(near, far, in ? i : S - 1, S, noise, g);
            float delta;
            if (i + 1 < S) delta = uniform_z2(near, far, i + 1, S, noise, g + 1) - z;
            else delta = (far - near) / (float)S;  // renderer_dynamic.py:160,182
            const float sg = in ? __ldg(sigma + g) : 0.f;
            const float alpha = in ? 1.0f - expf(((-kexp * delta) * cfg.density_scale) * sg) : 0.f;
            const float v = (1.0f - alpha) + 1e-15f;
            float incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl *= o;
            }
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0f;
            const float T = carry * excl;
            carry *= __shfl_sync(0xffffffffu, incl, 31);
            const float w = in ? alpha * T : 0.f;
            ws += w;
            dep = fmaf(w, z, dep);
            if (weights_out && in) { weights_out[g] = w; z_out[g] = z; }
            const bool m = w > 1e-4f;  // renderer_dynamic.py:202
            if (__ballot_sync(0xffffffffu, m) == 0) {
                if (rgbs_out && in) *reinterpret_cast<float4*>(rgbs_out + g * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
                continue;
            }

            // ---- heads on this 32-sample chunk ----
            {
                uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
                if (in) {
                    const uint4* src = reinterpret_cast<const uint4*>(geo + g * kGeo);
                    a0 = __ldg(src); a1 = __ldg(src + 1);
                }
                uint4* dst = reinterpret_cast<uint4*>(geo_s + lane * kGeoLd);
                dst[0] = a0; dst[1] = a1;
            }
            __syncwarp();
            uint32_t a[2][1][4];
            load_a_frags<1>(geo_s, kGeoLd, a, lane);
            if (rgbs_out) {  // the staged geo rows are in registers now: reuse them to stage colours
                __syncwarp();
                reinterpret_cast<float4*>(geo_s)[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
                __syncwarp();
            }
            const float wm = m ? w : 0.f;
            const float w0 = __shfl_sync(0xffffffffu, wm, gq), w1 = __shfl_sync(0xffffffffu, wm, gq + 8),
                        w2 = __shfl_sync(0xffffffffu, wm, gq + 16), w3 = __shfl_sync(0xffffffffu, wm, gq + 24);
#pragma unroll
            for (int net = 0; net < NETS; ++net) {
                const __half* Wn = Wsm + net * kHeadHalves;
                float acc[2][8][4];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 u = *reinterpret_cast<const float2*>(u_s + net * kHidden + 8 * j + 2 * tq);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        acc[mt][j][0] = u.x; acc[mt][j][1] = u.y;
                        acc[mt][j][2] = u.x; acc[mt][j][3] = u.y;
                    }
                }
                warp_gemm_regA<1, 8>(a, Wn + kHeadW1g, kLdK16, acc, lane);
                uint32_t a2[2][4][4];
                relu_to_a<8>(acc, a2);
                zero_acc<8>(acc);
                warp_gemm_regA<4, 8>(a2, Wn + kHeadW2, kLdK64, acc, lane);
                relu_to_a<8>(acc, a2);
                float o[2][1][4];
                zero_acc<1>(o);
                warp_gemm_regA<4, 1>(a2, Wn + kHeadW3, kLdK64, o, lane);
                // colours of rows gq, gq+8, gq+16, gq+24; columns 2tq, 2tq+1 of the output layer
                const float s00 = sigmoidf_(o[0][0][0]), s01 = sigmoidf_(o[0][0][1]),
                            s10 = sigmoidf_(o[0][0][2]), s11 = sigmoidf_(o[0][0][3]),
                            s20 = sigmoidf_(o[1][0][0]), s21 = sigmoidf_(o[1][0][1]),
                            s30 = sigmoidf_(o[1][0][2]), s31 = sigmoidf_(o[1][0][3]);
                if (LIDAR) {
                    // h = [raydrop, intensity] (network_dynamic.py:317): slot 0 = intensity, 1 = raydrop
                    if (tq == 0) {
                        const float s = w0 * s00 + w1 * s10 + w2 * s20 + w3 * s30;
                        if (net == 0) img[1] += s; else img[0] += s;
                        if (rgbs_out) {
                            const int ch = net == 0 ? 1 : 0;
                            float* cs = reinterpret_cast<float*>(geo_s);  // [32][4] colour staging
                            cs[gq * 4 + ch] = w0 > 0.f ? s00 : 0.f;
                            cs[(gq + 8) * 4 + ch] = w1 > 0.f ? s10 : 0.f;
                            cs[(gq + 16) * 4 + ch] = w2 > 0.f ? s20 : 0.f;
                            cs[(gq + 24) * 4 + ch] = w3 > 0.f ? s30 : 0.f;
                        }
                    }
                } else {
                    if (tq == 0) {
                        img[0] += w0 * s00 + w1 * s10 + w2 * s20 + w3 * s30;
                        img[1] += w0 * s01 + w1 * s11 + w2 * s21 + w3 * s31;
                    } else if (tq == 1) {
                        img[2] += w0 * s00 + w1 * s10 + w2 * s20 + w3 * s30;
                    }
                    if (rgbs_out && tq < 2) {
                        const float z1 = tq == 0 ? 1.f : 0.f;  // output column 3 is padding
                        float2* cs = reinterpret_cast<float2*>(geo_s);  // [32][2 x float2] colour staging
                        cs[gq * 2 + tq] = make_float2(w0 > 0.f ? s00 : 0.f, w0 > 0.f ? s01 * z1 : 0.f);
                        cs[(gq + 8) * 2 + tq] = make_float2(w1 > 0.f ? s10 : 0.f, w1 > 0.f ? s11 * z1 : 0.f);
                        cs[(gq + 16) * 2 + tq] = make_float2(w2 > 0.f ? s20 : 0.f, w2 > 0.f ? s21 * z1 : 0.f);
                        cs[(gq + 24) * 2 + tq] = make_float2(w3 > 0.f ? s30 : 0.f, w3 > 0.f ? s31 * z1 : 0.f);
                    }
                }
            }
            __syncwarp();
            if (rgbs_out) {
                if (in) *reinterpret_cast<float4*>(rgbs_out + g * 4) = reinterpret_cast<const float4*>(geo_s)[lane];
                __syncwarp();
            }
        }
        // ---- reduce over the warp and write the ray ----
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            ws += __shfl_xor_sync(0xffffffffu, ws, d);
            dep += __shfl_xor_sync(0xffffffffu, dep, d);
#pragma unroll
            for (int c = 0; c < NCH; ++c) img[c] += __shfl_xor_sync(0xffffffffu, img[c], d);
        }
        if (lane == 0) {
            ws_out[r] = ws;
            depth_out[r] = dep;
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                image_out[(size_t)r * NCH + c] = LIDAR ? img[c] : img[c] + (1.0f - ws) * bg_color;
        }
    }
}

bool g_attr = false;
int ensure_attrs() {
    if (g_attr) return NVSF_OK;
    cudaError_t e = cudaFuncSetAttribute(k_render_composite<true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)render_smem<true>());
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_render_composite<false>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)render_smem<false>());
    if (e != cudaSuccess) return (int)e;
    g_attr = true;
    return NVSF_OK;
}

}  // namespace

extern "C" {

size_t nvsf_render_uniform_scratch_bytes(uint32_t N, uint32_t S) {
    const size_t n = (size_t)N * S;
    return ws_align(n * sizeof(float)) + ws_align(n * kGeo * sizeof(__half)) +
           nvsf_density_split_scratch_bytes(n);
}

/* phase 1: field evaluation of all N*S samples into scratch (sigma f32, geo f16[16]) */
int nvsf_render_uniform_density(const nvsf_field_config_t* cfg, const void* workspace,
                                const float* rays_o, const float* rays_d, const float* nears,
                                const float* fars, const float* noise, uint32_t N, uint32_t S,
                                void* scratch, size_t scratch_bytes, void* stream) {
    if (N == 0) return NVSF_OK;
    if (!field_cfg_ok(cfg) || !workspace || !rays_o || !rays_d || !nears || !fars || !scratch ||
        S == 0)
        return NVSF_E_INVALID;
    if (scratch_bytes < nvsf_render_uniform_scratch_bytes(N, S)) return NVSF_E_WORKSPACE;
    const size_t n = (size_t)N * S;
    float* sigma = reinterpret_cast<float*>(scratch);
    __half* geo = reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(scratch) +
                                            ws_align(n * sizeof(float)));
    void* split = reinterpret_cast<unsigned char*>(geo) + ws_align(n * kGeo * sizeof(__half));
    return nvsf_launch_density(cfg, workspace, nullptr, rays_o, rays_d, nears, fars, noise, S, n,
                               sigma, geo, nullptr, nullptr, split, (cudaStream_t)stream);
}

/* phase 2: compositing + colour heads from scratch */
int nvsf_render_uniform_composite(const nvsf_field_config_t* cfg, const void* workspace,
                                  uint32_t lidar, const float* rays_d, const float* nears,
                                  const float* fars, const float* noise, uint32_t N, uint32_t S,
                                  float bg_color, const void* scratch, size_t scratch_bytes,
                                  float* depth, float* image, float* weights_sum, float* weights,
                                  float* z_vals, void* stream) {
    return nvsf_render_composite_launch(cfg, workspace, lidar, rays_d, nears, fars, noise, N, S,
                                        bg_color, scratch, scratch_bytes, depth, image, weights_sum,
                                        weights, z_vals, nullptr, stream);
}

}  // extern "C"

int nvsf_render_composite_launch(const nvsf_field_config_t* cfg, const void* workspace,
                                 uint32_t lidar, const float* rays_d, const float* nears,
                                 const float* fars, const float* noise, uint32_t N, uint32_t S,
                                 float bg_color, const void* scratch, size_t scratch_bytes,
                                 float* depth, float* image, float* weights_sum, float* weights,
                                 float* z_vals, void* rgbs, void* stream) {
    if (N == 0) return NVSF_OK;
    if (!field_cfg_ok(cfg) || !workspace || !rays_d || !nears || !fars || !scratch || !depth ||
        !image || !weights_sum || S == 0)
        return NVSF_E_INVALID;
    if ((weights == nullptr) != (z_vals == nullptr)) return NVSF_E_INVALID;
    if (scratch_bytes < ws_align((size_t)N * S * sizeof(float)) + (size_t)N * S * kGeo * sizeof(__half))
        return NVSF_E_WORKSPACE;
    int st = ensure_attrs();
    if (st != NVSF_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)N * S;
    const float* sigma = reinterpret_cast<const float*>(scratch);
    const __half* geo = reinterpret_cast<const __half*>(
        reinterpret_cast<const unsigned char*>(scratch) + ws_align(n * sizeof(float)));
    const FieldPtrs P = nvsf_make_field_ptrs(cfg, workspace);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t blocks = std::min<uint32_t>(nvsf_div_up(N, (uint32_t)kRWarps), (uint32_t)sms * 4);
    if (lidar) {
        k_render_composite<true><<<blocks, kRWarps * 32, render_smem<true>(), s>>>(
            *cfg, P.mlp, rays_d, nears, fars, noise, sigma, geo, N, S, bg_color, depth, image,
            weights_sum, weights, z_vals, reinterpret_cast<float*>(rgbs));
    } else {
        k_render_composite<false><<<blocks, kRWarps * 32, render_smem<false>(), s>>>(
            *cfg, P.mlp, rays_d, nears, fars, noise, sigma, geo, N, S, bg_color, depth, image,
            weights_sum, weights, z_vals, reinterpret_cast<float*>(rgbs));
    }
    return nvsf_launch_status();
}

extern "C" {

int nvsf_render_uniform(const nvsf_field_config_t* cfg, const void* workspace, uint32_t lidar,
                        const float* rays_o, const float* rays_d, const float* nears,
                        const float* fars, const float* noise, uint32_t N, uint32_t S,
                        float bg_color, void* scratch, size_t scratch_bytes, float* depth,
                        float* image, float* weights_sum, float* weights, float* z_vals,
                        void* stream) {
    int st = nvsf_render_uniform_density(cfg, workspace, rays_o, rays_d, nears, fars, noise, N, S,
                                         scratch, scratch_bytes, stream);
    if (st != NVSF_OK) return st;
    return nvsf_render_uniform_composite(cfg, workspace, lidar, rays_d, nears, fars, noise, N, S,
                                         bg_color, scratch, scratch_bytes, depth, image,
                                         weights_sum, weights, z_vals, stream);
}

}  // extern "C"
