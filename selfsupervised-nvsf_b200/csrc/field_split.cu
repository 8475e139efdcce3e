// nvsf_b200 — staged variant of the density evaluation (same arithmetic as k_field_density).
//
// ncu on the fused kernel (profiles/r01_density_v2_ncu_full.txt) shows it bound by gather latency
// at 16 warps/SM: the tensor-core MLP phases need ~128 registers and a 70 KB feature tile per CTA,
// which caps the number of warps that can have gathers in flight.  This variant separates the
// stages so that the gather stage — 82 % of the loads — runs lean (no shared memory, <= 64
// registers, 32 warps/SM):
//     k_flow_stage    flow-grid features + flow MLP (tensor cores)   -> flow  [n,8]  f32
//     k_encode_stage  the three warped queries, all remaining gathers -> feats [n,128] f16
//     k_sigma_stage   sigma MLP (tensor cores) + trunc_exp            -> sigma, geo
// The price is 288 B/sample of intermediate traffic (on top of 6.5 KB gathered per sample); the
// frame is processed in chunks so the intermediates stay small.  Selected with
// nvsf_set_option("density_mode", 1); mode 0 is the fused kernel.
#include <algorithm>
#include <string>
#include <vector>

#include "field_common.cuh"

namespace {

#ifndef NVSF_ENC_UNROLL_D
#define NVSF_ENC_UNROLL_D 1
#endif
// Measured on B200 (LiDAR frame, encode stage): level pairs fully unrolled 31.1 ms, x2 29.2 ms, rolled
// (1 pair per iteration) 28.0 ms; rolling further (single levels, rolled queries) is slower again
// (29.3-29.7 ms): the kernel trades instruction-cache misses against loads in flight.
constexpr int kEncUnrollD = NVSF_ENC_UNROLL_D;  // level-pair unrolling of the 2-D hash loop of k_encode_stage
constexpr int kSTile = 256;
constexpr int kFld = 56;      // flow-stage tile row: 32 features + 8 fp32 flow slots (+pad), halves
constexpr int kFlowWHalves = kSigW1;  // the flow MLP part of the weight image
constexpr size_t kFlowStageSmem = (size_t)kFlowWHalves * 2 + (size_t)kSTile * kFld * 2;
constexpr int kSigWHalves = kDensityWHalves - kSigW1;
constexpr size_t kSigmaStageSmem = (size_t)kSigWHalves * 2 + (size_t)kSTile * kLdK128 * 2;

template <bool FROM_RAYS>
__device__ __forceinline__ void sample_position(const nvsf_field_config_t& cfg, size_t g,
                                                const float* __restrict__ xin,
                                                const float* __restrict__ rays_o,
                                                const float* __restrict__ rays_d,
                                                const float* __restrict__ nears,
                                                const float* __restrict__ fars,
                                                const float* __restrict__ noise, uint32_t S,
                                                float& x, float& y, float& z) {
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

// ---- stage 1: flow -----------------------------------------------------------------------------
template <bool FROM_RAYS>
__global__ void __launch_bounds__(kSTile, 2)
k_flow_stage(const __grid_constant__ nvsf_field_config_t cfg, const __grid_constant__ FieldPtrs P,
             const float* __restrict__ xin, const float* __restrict__ rays_o,
             const float* __restrict__ rays_d, const float* __restrict__ nears,
             const float* __restrict__ fars, const float* __restrict__ noise, uint32_t S,
             size_t begin, size_t count, float* __restrict__ flow_out,
             __half* __restrict__ flowfeat_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* Wsm = reinterpret_cast<__half*>(smem_raw);
    __half* Xs = Wsm + kFlowWHalves;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    block_copy16(Wsm, P.mlp, kFlowWHalves * 2 / 16, tid, kSTile);
    __syncthreads();
    __half* xrow = Xs + tid * kFld;
    const __half* Aw = Xs + warp * 32 * kFld;
    const size_t n_tiles = (count + kSTile - 1) / kSTile;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t li = tile * kSTile + tid;
        const bool live = li < count;
        float x = 0.5f, y = 0.5f, z = 0.5f;
        if (live)
            sample_position<FROM_RAYS>(cfg, begin + li, xin, rays_o, rays_d, nears, fars, noise, S,
                                       x, y, z);
#pragma unroll 2
        for (int l = 0; l < kFlLevels; ++l) {
            const float2 f = hash3_f2(P.flow, lv(cfg.fl[l]), x, y, z);
            *reinterpret_cast<uint32_t*>(xrow + 2 * l) = pack_half2(f.x, f.y);
        }
        if (flowfeat_out && live) {  // kept for the backward pass (input of the flow MLP)
            const uint4* s = reinterpret_cast<const uint4*>(xrow);
            uint4* d = reinterpret_cast<uint4*>(flowfeat_out + li * kFlowIn);
#pragma unroll
            for (int i = 0; i < kFlowIn / 8; ++i) d[i] = s[i];
        }
        __syncwarp();
        float acc[2][8][4];
        zero_acc<8>(acc);
        {
            uint32_t a[2][2][4];
            load_a_frags<2>(Aw, kFld, a, lane);
            warp_gemm_regA<2, 8>(a, Wsm + kFlowW1, kLdK32, acc, lane);
        }
        uint32_t a2[2][4][4];
        relu_to_a<8>(acc, a2);
        zero_acc<8>(acc);
        warp_gemm_regA<4, 8>(a2, Wsm + kFlowW2, kLdK64, acc, lane);
        relu_to_a<8>(acc, a2);
        float o[2][1][4];
        zero_acc<1>(o);
        warp_gemm_regA<4, 1>(a2, Wsm + kFlowW3, kLdK64, o, lane);
        const int gq = lane >> 2, tq = lane & 3;
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            float* r0 = reinterpret_cast<float*>(Xs + (warp * 32 + mt * 16 + gq) * kFld + 32);
            *reinterpret_cast<float2*>(r0 + 2 * tq) = make_float2(o[mt][0][0], o[mt][0][1]);
            float* r1 = reinterpret_cast<float*>(Xs + (warp * 32 + mt * 16 + gq + 8) * kFld + 32);
            *reinterpret_cast<float2*>(r1 + 2 * tq) = make_float2(o[mt][0][2], o[mt][0][3]);
        }
        __syncwarp();
        if (live) {
            const float4* s = reinterpret_cast<const float4*>(xrow + 32);
            float4* d = reinterpret_cast<float4*>(flow_out + li * 8);
            d[0] = s[0];
            d[1] = s[1];
        }
        __syncwarp();
    }
}

// ---- stage 2: gathers ---------------------------------------------------------------------------
__device__ __forceinline__ void st8g(__half* row, int col, const float (&v)[8]) {
    uint4 o;
    o.x = pack_half2(v[0], v[1]); o.y = pack_half2(v[2], v[3]);
    o.z = pack_half2(v[4], v[5]); o.w = pack_half2(v[6], v[7]);
    *reinterpret_cast<uint4*>(row + col) = o;
}

template <bool FROM_RAYS>
__global__ void __launch_bounds__(256, 4)
k_encode_stage(const __grid_constant__ nvsf_field_config_t cfg,
               const __grid_constant__ FieldPtrs P, const float* __restrict__ xin,
               const float* __restrict__ rays_o, const float* __restrict__ rays_d,
               const float* __restrict__ nears, const float* __restrict__ fars,
               const float* __restrict__ noise, uint32_t S, size_t begin, size_t count,
               const float* __restrict__ flow_in, __half* __restrict__ feat_out) {
    const size_t li = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= count) return;
    float x, y, z;
    sample_position<FROM_RAYS>(cfg, begin + li, xin, rays_o, rays_d, nears, fars, noise, S, x, y, z);
    const int valid1 = P.ti->valid[1], valid2 = P.ti->valid[2];
    const float4 f0 = __ldg(reinterpret_cast<const float4*>(flow_in + li * 8));
    const float4 f1 = __ldg(reinterpret_cast<const float4*>(flow_in + li * 8) + 1);
    float qx[3], qy[3], qz[3];
    int qi[3];
    qx[0] = x; qy[0] = y; qz[0] = z; qi[0] = 0;
    qx[1] = valid1 ? x + f0.x : x; qy[1] = valid1 ? y + f0.y : y;
    qz[1] = valid1 ? z + f0.z : z; qi[1] = valid1 ? 1 : 0;
    qx[2] = valid2 ? x + f0.w : x; qy[2] = valid2 ? y + f1.x : y;
    qz[2] = valid2 ? z + f1.y : z; qi[2] = valid2 ? 2 : 0;
    __half* row = feat_out + li * kFeat;

    // (a) space planes -> [0,32)
#pragma unroll 1
    for (int s = 0; s < kPlScales; ++s) {
        const uint32_t R = cfg.pl_res[s];
        const float* base = P.pls + P.pls_scale[s];
        float v[8];
        plane2d_mul(base, R, x, y, v, true);
        plane2d_mul(base + (size_t)R * R * 8, R, x, z, v, false);
        plane2d_mul(base + (size_t)2 * R * R * 8, R, y, z, v, false);
        st8g(row, 8 * s, v);
    }
    // (b) collapsed time planes -> [32,64)
#pragma unroll 1
    for (int s = 0; s < kPlScales; ++s) {
        const uint32_t R = cfg.pl_res[s];
        float acc8[8];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float* base = P.pld + (size_t)qi[q] * P.pld_per_q + P.pld_scale[s];
            float v[8];
            plane1d_mul(base, R, qx[q], v, true);
            plane1d_mul(base + (size_t)R * 8, R, qy[q], v, false);
            plane1d_mul(base + (size_t)2 * R * 8, R, qz[q], v, false);
            const float wq = q == 0 ? 0.5f : 0.25f;
#pragma unroll
            for (int f = 0; f < 8; ++f) acc8[f] = q == 0 ? wq * v[f] : fmaf(wq, v[f], acc8[f]);
        }
        st8g(row, 32 + 8 * s, acc8);
    }
    // (c) static 3-D hash -> [64,96)
#pragma unroll 1
    for (int l = 0; l < kHsLevels; l += 2) {
        float v[8];
        hash3_f4(P.hs16, lv(cfg.hs[l]), x, y, z, v);
        hash3_f4(P.hs16, lv(cfg.hs[l + 1]), x, y, z, v + 4);
        st8g(row, 64 + 4 * l, v);
    }
    // (d) collapsed 2-D hashes -> [96,120), zero pad [120,128)
#pragma unroll 1
#ifdef NVSF_EXP_SKIP
    for (int p = 0; p < NVSF_EXP_SKIP; ++p) {
#else
    for (int p = 0; p < 3; ++p) {
#endif
        float u[3], w[3];
        const float* tab[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            u[q] = p == 2 ? qy[q] : qx[q];
            w[q] = p == 0 ? qy[q] : qz[q];
            tab[q] = P.dyn + (size_t)qi[q] * P.dyn_per_q + P.dyn_plane[p];
        }
        uint32_t packed[4] = {0u, 0u, 0u, 0u};
#pragma unroll kEncUnrollD
        for (int l2 = 0; l2 < kHdLevels / 2; ++l2) {
            float r2[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const LevelArgs L = lv(cfg.hd[p][2 * l2 + j]);
                const float a = hash2_f1(tab[0], L, u[0], w[0]);
                const float b = hash2_f1(tab[1], L, u[1], w[1]);
                const float c = hash2_f1(tab[2], L, u[2], w[2]);
                r2[j] = 0.5f * a + 0.25f * (b + c);
            }
            const uint32_t pk = pack_half2(r2[0], r2[1]);
            // rolled loop: select the destination register without dynamic indexing
            packed[0] = l2 == 0 ? pk : packed[0];
            packed[1] = l2 == 1 ? pk : packed[1];
            packed[2] = l2 == 2 ? pk : packed[2];
            packed[3] = l2 == 3 ? pk : packed[3];
        }
        *reinterpret_cast<uint4*>(row + 96 + 8 * p) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
    *reinterpret_cast<uint4*>(row + 120) = make_uint4(0, 0, 0, 0);
}

// ---- stage 3: sigma MLP -------------------------------------------------------------------------
__global__ void __launch_bounds__(kSTile, 2)
k_sigma_stage(const __half* __restrict__ mlp, const __half* __restrict__ feat, size_t count,
              float* __restrict__ sigma_out, __half* __restrict__ geo_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* Wsm = reinterpret_cast<__half*>(smem_raw);   // image starting at kSigW1
    __half* Xs = Wsm + kSigWHalves;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    block_copy16(Wsm, mlp + kSigW1, kSigWHalves * 2 / 16, tid, kSTile);
    __syncthreads();
    const __half* W1 = Wsm;                          // [64][136]
    const __half* W2 = Wsm + (kSigW2 - kSigW1);      // [16][72]
    __half* Aw = Xs + warp * 32 * kLdK128;
    const size_t n_tiles = (count + kSTile - 1) / kSTile;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t row0 = tile * kSTile + warp * 32;
        // coalesced copy of this warp's 32 rows x 256 B into the padded tile
        __syncwarp();
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const int piece = i * 32 + lane;  // 16-byte piece index within the 8 KB block
            const int r = piece >> 4, c = piece & 15;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (row0 + r < count)
                v = __ldcs(reinterpret_cast<const uint4*>(feat + (row0 + r) * kFeat) + c);
            *reinterpret_cast<uint4*>(Aw + r * kLdK128 + c * 8) = v;
        }
        __syncwarp();
        float acc[2][8][4];
        zero_acc<8>(acc);
#pragma unroll
        for (int kk = 0; kk < kFeat / 16; ++kk) {
            uint32_t a[2][1][4];
            ldsm_x4(a[0][0], Aw + (lane & 15) * kLdK128 + kk * 16 + (lane >> 4) * 8);
            ldsm_x4(a[1][0], Aw + (16 + (lane & 15)) * kLdK128 + kk * 16 + (lane >> 4) * 8);
            warp_gemm_regA<1, 8>(a, W1 + kk * 16, kLdK128, acc, lane);
        }
        uint32_t a2[2][4][4];
        relu_to_a<8>(acc, a2);
        float o[2][2][4];
        zero_acc<2>(o);
        warp_gemm_regA<4, 2>(a2, W2, kLdK64, o, lane);
        const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const size_t row = row0 + mt * 16 + hrow * 8 + gq;
                if (row < count) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float c0 = o[mt][j][2 * hrow], c1 = o[mt][j][2 * hrow + 1];
                        *reinterpret_cast<uint32_t*>(geo_out + row * kGeo + 8 * j + 2 * tq) =
                            pack_half2(c0, c1);
                        if (j == 0 && tq == 0) sigma_out[row] = expf(c0);
                    }
                }
            }
    }
}

// Optional per-stage timing with CUDA events on the launching stream (bench.py's roofline): four
// events per chunk, read back (and reset) by nvsf_stage_timing_read.
struct StageProf {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    size_t used = 0;
    cudaEvent_t next(cudaStream_t s) {
        if (used == ev.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev.push_back(e);
        }
        cudaEvent_t e = ev[used++];
        cudaEventRecord(e, s);
        return e;
    }
} g_prof;

bool g_attr = false;
int ensure_attrs() {
    if (g_attr) return NVSF_OK;
    cudaError_t e;
    e = cudaFuncSetAttribute(k_flow_stage<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kFlowStageSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_flow_stage<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kFlowStageSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_sigma_stage, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kSigmaStageSmem);
    if (e != cudaSuccess) return (int)e;
    g_attr = true;
    return NVSF_OK;
}

}  // namespace

size_t nvsf_density_split_scratch_bytes(size_t n) {
    const size_t chunk = std::min<size_t>(n, kSplitChunk);
    return ws_align(chunk * 8 * sizeof(float)) + ws_align(chunk * kFeat * sizeof(__half));
}

int nvsf_launch_density_split(const nvsf_field_config_t* cfg, const void* workspace,
                              const float* x, const float* rays_o, const float* rays_d,
                              const float* nears, const float* fars, const float* noise,
                              uint32_t S, size_t n, float* sigma, void* geo, void* features,
                              float* flow, void* split_scratch, cudaStream_t stream,
                              const DensityKeep* keep) {
    int st = ensure_attrs();
    if (st != NVSF_OK) return st;
    const FieldPtrs P = nvsf_make_field_ptrs(cfg, workspace);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t chunk = std::min<size_t>(n, kSplitChunk);
    float* flow_buf = reinterpret_cast<float*>(split_scratch);
    __half* feat_buf = reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(split_scratch) +
                                                 ws_align(chunk * 8 * sizeof(float)));
    for (size_t begin = 0; begin < n; begin += chunk) {
        const size_t count = std::min(chunk, n - begin);
        __half* ff = nullptr;
        if (keep) {  // training: the intermediates of every chunk are kept for the backward pass
            flow_buf = keep->flow + begin * 8;
            feat_buf = keep->feats + begin * kFeat;
            ff = keep->flowfeat + begin * kFlowIn;
        }
        const size_t tiles = (count + kSTile - 1) / kSTile;
        const int grid_p = (int)std::min<size_t>(tiles, (size_t)sms * 2);
        if (g_prof.on) g_prof.next(stream);
        if (x) {
            k_flow_stage<false><<<grid_p, kSTile, kFlowStageSmem, stream>>>(
                *cfg, P, x, nullptr, nullptr, nullptr, nullptr, nullptr, 1, begin, count, flow_buf, ff);
            if (g_prof.on) g_prof.next(stream);
            k_encode_stage<false><<<(unsigned)tiles, 256, 0, stream>>>(
                *cfg, P, x, nullptr, nullptr, nullptr, nullptr, nullptr, 1, begin, count, flow_buf,
                feat_buf);
        } else {
            k_flow_stage<true><<<grid_p, kSTile, kFlowStageSmem, stream>>>(
                *cfg, P, nullptr, rays_o, rays_d, nears, fars, noise, S, begin, count, flow_buf, ff);
            if (g_prof.on) g_prof.next(stream);
            k_encode_stage<true><<<(unsigned)tiles, 256, 0, stream>>>(
                *cfg, P, nullptr, rays_o, rays_d, nears, fars, noise, S, begin, count, flow_buf,
                feat_buf);
        }
        if (g_prof.on) g_prof.next(stream);
        k_sigma_stage<<<grid_p, kSTile, kSigmaStageSmem, stream>>>(
            P.mlp, feat_buf, count, sigma + begin, reinterpret_cast<__half*>(geo) + begin * kGeo);
        if (g_prof.on) g_prof.next(stream);
        if (features)
            cudaMemcpyAsync(reinterpret_cast<__half*>(features) + begin * kFeat, feat_buf,
                            count * kFeat * sizeof(__half), cudaMemcpyDeviceToDevice, stream);
        if (flow) {
            cudaMemcpy2DAsync(flow + begin * 6, 6 * sizeof(float), flow_buf, 8 * sizeof(float),
                              6 * sizeof(float), count, cudaMemcpyDeviceToDevice, stream);
        }
    }
    return nvsf_launch_status();
}

// ---- stage timing (see StageProf) -------------------------------------------------------------------
void nvsf_stage_timing_enable(int on) {
    g_prof.on = on != 0;
    g_prof.used = 0;
}

extern "C" int nvsf_stage_timing_read(float* ms3, uint32_t* launches) {
    if (!ms3) return NVSF_E_INVALID;
    ms3[0] = ms3[1] = ms3[2] = 0.f;
    const size_t groups = g_prof.used / 4;
    for (size_t g = 0; g < groups; ++g) {
        cudaError_t e = cudaEventSynchronize(g_prof.ev[4 * g + 3]);
        if (e != cudaSuccess) return (int)e;
        for (int k = 0; k < 3; ++k) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, g_prof.ev[4 * g + k], g_prof.ev[4 * g + k + 1]);
            ms3[k] += ms;
        }
    }
    if (launches) *launches = (uint32_t)groups;
    g_prof.used = 0;
    return NVSF_OK;
}
