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

extern int g_flow_ts;   // sigma_tc.cu: option "flow_ts"

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
constexpr int kSigTile = 384;  // rows per CTA tile of the sigma stage: 12 warps x 32 rows, 1 CTA per SM
constexpr size_t kSigmaStageSmem = (size_t)kSigWHalves * 2 + (size_t)2 * kSigTile * kLdK128 * 2;  // two tile buffers per warp (223 KB)

// ---- stage 1: flow -----------------------------------------------------------------------------
template <bool FROM_RAYS, bool TAB16>
__global__ void __launch_bounds__(kSTile, 2)
k_flow_stage(const __grid_constant__ nvsf_field_config_t cfg, const __grid_constant__ FieldPtrs P,
             const float* __restrict__ xin, const float* __restrict__ rays_o,
             const float* __restrict__ rays_d, const float* __restrict__ nears,
             const float* __restrict__ fars, const float* __restrict__ noise, uint32_t S,
             size_t begin, size_t count, float* __restrict__ flow_out,
             __half* __restrict__ flowfeat_out, float* __restrict__ qpos, size_t qstride) {
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
            const float2 f = TAB16 ? hash3_h2(P.flow16, lv(cfg.fl[l]), x, y, z)
                                   : hash3_f2(P.flow, lv(cfg.fl[l]), x, y, z);
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
            *reinterpret_cast<float2*>(r0 + 2 * tq) = make_float2(round_f16(o[mt][0][0]), round_f16(o[mt][0][1]));
            float* r1 = reinterpret_cast<float*>(Xs + (warp * 32 + mt * 16 + gq + 8) * kFld + 32);
            *reinterpret_cast<float2*>(r1 + 2 * tq) = make_float2(round_f16(o[mt][0][2]), round_f16(o[mt][0][3]));
        }
        __syncwarp();
        if (live) {
            const float4* s = reinterpret_cast<const float4*>(xrow + 32);
            float4* d = reinterpret_cast<float4*>(flow_out + li * 8);
            const float4 f0 = s[0], f1 = s[1];
            d[0] = f0;
            d[1] = f1;
            if (qpos) {  // the three query positions, planar [q * 3 + axis][sample], for k_dyn_stage
                const int valid1 = P.ti->valid[1], valid2 = P.ti->valid[2];
                float* q = qpos + li;
                q[0] = x; q[qstride] = y; q[2 * qstride] = z;
                q[3 * qstride] = valid1 ? x + f0.x : x;
                q[4 * qstride] = valid1 ? y + f0.y : y;
                q[5 * qstride] = valid1 ? z + f0.z : z;
                q[6 * qstride] = valid2 ? x + f0.w : x;
                q[7 * qstride] = valid2 ? y + f1.x : y;
                q[8 * qstride] = valid2 ? z + f1.y : z;
            }
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

// DYN_PRE: the 24 dynamic-hash features were produced by k_dyn_stage (planar fp16 rows
// dyn_in[c * dyn_stride + sample], c = plane * 8 + level) and are only merged into the row here.
template <bool FROM_RAYS, bool DYN_PRE, bool PAIR>
__global__ void __launch_bounds__(256, 4)
k_encode_stage(const __grid_constant__ nvsf_field_config_t cfg,
               const __grid_constant__ FieldPtrs P, const float* __restrict__ xin,
               const float* __restrict__ rays_o, const float* __restrict__ rays_d,
               const float* __restrict__ nears, const float* __restrict__ fars,
               const float* __restrict__ noise, uint32_t S, size_t begin, size_t count,
               const float* __restrict__ flow_in, __half* __restrict__ feat_out,
               const unsigned short* __restrict__ dyn_in, size_t dyn_stride) {
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

    // (a) space planes -> [0,32), (b) collapsed time planes -> [32,64); with DYN_PRE (mode 2) from
    // the fp16 texel mirrors (one 16-byte load per texel)
    if constexpr (DYN_PRE) {
#pragma unroll 1
        for (int s = 0; s < kPlScales; ++s) {
            const uint32_t R = cfg.pl_res[s];
            const __half* base = P.pls16 + P.pls_scale[s];
            float v[8];
            plane2d_mul_h(base, R, x, y, v, true);
            plane2d_mul_h(base + (size_t)R * R * 8, R, x, z, v, false);
            plane2d_mul_h(base + (size_t)2 * R * R * 8, R, y, z, v, false);
            st8g(row, 8 * s, v);
        }
#pragma unroll 1
        for (int s = 0; s < kPlScales; ++s) {
            const uint32_t R = cfg.pl_res[s];
            float acc8[8];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const __half* base = P.pld16 + (size_t)qi[q] * P.pld_per_q + P.pld_scale[s];
                float v[8];
                plane1d_mul_h(base, R, qx[q], v, true);
                plane1d_mul_h(base + (size_t)R * 8, R, qy[q], v, false);
                plane1d_mul_h(base + (size_t)2 * R * 8, R, qz[q], v, false);
                const float wq = q == 0 ? 0.5f : 0.25f;
#pragma unroll
                for (int f = 0; f < 8; ++f) acc8[f] = q == 0 ? wq * v[f] : fmaf(wq, v[f], acc8[f]);
            }
            st8g(row, 32 + 8 * s, acc8);
        }
    } else {
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
    }
    // (c) static 3-D hash -> [64,96)
#pragma unroll 1
    for (int l = 0; l < kHsLevels; l += 2) {
        float v[8];
        if (PAIR && cfg.hs[l].hashed && cfg.hs[l + 1].hashed) {
            hash3_f4_pair(P.hs16, lv(cfg.hs[l]), x, y, z, v);
            hash3_f4_pair(P.hs16, lv(cfg.hs[l + 1]), x, y, z, v + 4);
        } else {
            hash3_f4(P.hs16, lv(cfg.hs[l]), x, y, z, v);
            hash3_f4(P.hs16, lv(cfg.hs[l + 1]), x, y, z, v + 4);
        }
        st8g(row, 64 + 4 * l, v);
    }
    // (d) collapsed 2-D hashes -> [96,120), padding [120,128) = 1
    if constexpr (DYN_PRE) {
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            uint32_t h[8];
#pragma unroll
            for (int l = 0; l < 8; ++l) h[l] = __ldcs(dyn_in + (size_t)(8 * p + l) * dyn_stride + li);
            *reinterpret_cast<uint4*>(row + 96 + 8 * p) =
                make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
        }
    } else {
#pragma unroll 1
    for (int p = 0; p < 3; ++p) {
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
    }
    *reinterpret_cast<uint4*>(row + 120) = make_uint4(kOnesH2, kOnesH2, kOnesH2, kOnesH2);  // tcnn input padding = 1
}


// ---- stage 2a: dynamic 2-D hashes, tables staged in shared memory ---------------------------------
// ncu on k_encode_stage (profiles/encode_stage_ncu.json): 288 of its ~590 loads per sample are the
// 4-byte corner gathers of the 72 time-collapsed 2-D hash tables (3 planes x 8 levels x 3 query
// times); on the fine levels every lane of a warp hits its own 32-byte sector, so the kernel is
// bound by L1TEX wavefronts.  Those tables are small — 2^13..2^15 entries each — so this stage
// turns the problem around: a CTA (1024 threads, one per SM) copies the fp16 tables of a "type"
// (one plane, up to 4 consecutive levels, the 3 query times: <= 192 KB) into shared memory with
// cp.async.bulk (TMA, completion on an mbarrier) and then streams tiles of samples through them:
// a scattered 2-byte gather costs a bank-conflict degree (~3 cycles per warp) instead of up to
// 32 L1 wavefronts.  Types are handed out dynamically (one atomic tile counter per type; a CTA
// keeps its tables until its type runs dry, then helps the next one).  Output: planar fp16 rows
// dyn_out[c * stride + sample], c = plane * 8 + level, merged into the feature rows by
// k_encode_stage<., true>.  Price: the sample positions (rays + 32 B of flow) are re-read once
// per type from L2, and the table values are rounded to fp16 (they are fp16 parameters blended in
// fp32; the extra rounding is below the fp16 resolution of the feature rows themselves).
constexpr int kDynThreads = 1024;
constexpr int kDynMaxCombo = 4;
constexpr int kDynMaxTypes = 24;
constexpr size_t kDynTableBytes = 192 * 1024;

struct DynPlan {
    uint32_t ntypes, tile, tiles, stride;
    uint32_t tab_off[kDynMaxTypes][kDynMaxCombo];  // halves; the 3 query tables follow each other
    uint16_t cta_first[kDynMaxTypes];              // first CTA that starts on this type
    uint8_t ncombo[kDynMaxTypes], plane[kDynMaxTypes], level0[kDynMaxTypes];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (TMA engine); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// one level of a time-collapsed 2-D hash grid, table in shared memory (fp16); same arithmetic as
// hash2_f1.  HASHED is hoisted out of the corner loop so that the four gathers issue back to back.
template <bool HASHED>
__device__ __forceinline__ float hash2_sm(const __half* __restrict__ tab, const LevelArgs& L,
                                          float u, float v) {
    uint32_t cu, cv;
    float wu, wv;
    grid_pos(L.scale, u, cu, wu);
    grid_pos(L.scale, v, cv, wv);
    uint32_t i00, i10, i01, i11;
    if (HASHED) {
        const uint32_t m = L.size - 1, h0 = cv * 2654435761u, h1 = h0 + 2654435761u;
        i00 = (cu ^ h0) & m; i10 = ((cu + 1) ^ h0) & m;
        i01 = (cu ^ h1) & m; i11 = ((cu + 1) ^ h1) & m;
    } else {
        i00 = idx2(L, cu, cv); i10 = idx2(L, cu + 1, cv);
        i01 = idx2(L, cu, cv + 1); i11 = idx2(L, cu + 1, cv + 1);
    }
    const float t00 = __half2float(tab[i00]), t10 = __half2float(tab[i10]);
    const float t01 = __half2float(tab[i01]), t11 = __half2float(tab[i11]);
    // nested-lerp form (6 instead of 11 floating-point instructions; this kernel is issue bound)
    const float a = fmaf(wu, t10 - t00, t00), b = fmaf(wu, t11 - t01, t01);
    return fmaf(wv, b - a, a);
}
template <bool HASHED>
__device__ __forceinline__ float dyn_combo(const __half* __restrict__ t0, const LevelArgs& L,
                                           const float (&u)[3], const float (&w)[3]) {
    const float a = hash2_sm<HASHED>(t0, L, u[0], w[0]);
    const float b = hash2_sm<HASHED>(t0 + L.size, L, u[1], w[1]);
    const float c = hash2_sm<HASHED>(t0 + 2 * L.size, L, u[2], w[2]);
    return 0.5f * a + 0.25f * (b + c);
}

template <bool FROM_RAYS>
__global__ void __launch_bounds__(kDynThreads, 1)
k_dyn_stage(const __grid_constant__ nvsf_field_config_t cfg, const __grid_constant__ FieldPtrs P,
            const __grid_constant__ DynPlan plan, const float* __restrict__ xin,
            const float* __restrict__ rays_o, const float* __restrict__ rays_d,
            const float* __restrict__ nears, const float* __restrict__ fars,
            const float* __restrict__ noise, uint32_t S, size_t begin, size_t count,
            const float* __restrict__ qpos, __half* __restrict__ dyn_out,
            uint32_t* __restrict__ next_tile) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __half* tabs = reinterpret_cast<__half*>(smem_raw);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x;
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    const int ntypes = (int)plan.ntypes;
    int type = 0;
    for (int t = 1; t < ntypes; ++t)
        if (blockIdx.x >= plan.cta_first[t]) type = t;
    int loaded = -1;
    uint32_t parity = 0;
    const int valid1 = P.ti->valid[1], valid2 = P.ti->valid[2];
    for (int tried = 0; tried < ntypes;) {
        if (tid == 0) s_tile = atomicAdd(next_tile + type, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        __syncthreads();  // s_tile may be rewritten; nobody still reads the previous tables
        if (tile >= plan.tiles) {
            type = type + 1 == ntypes ? 0 : type + 1;
            ++tried;
            continue;
        }
        tried = 0;
        const int p = plan.plane[type], l0 = plan.level0[type], nc = plan.ncombo[type];
        if (loaded != type) {
            if (tid == 0) {
                uint32_t bytes = 0;
                for (int j = 0; j < nc; ++j) bytes += 3u * cfg.hd[p][l0 + j].size * 2u;
                mbar_expect_tx(&bar, bytes);
                for (int j = 0; j < nc; ++j) {
                    const uint32_t size = cfg.hd[p][l0 + j].size;
                    for (int q = 0; q < 3; ++q) {
                        const int qq = q == 0 ? 0 : (q == 1 ? (valid1 ? 1 : 0) : (valid2 ? 2 : 0));
                        bulk_g2s(tabs + plan.tab_off[type][j] + (size_t)q * size,
                                 P.dyn16 + (size_t)qq * P.dyn_per_q + P.dyn_plane[p] +
                                     cfg.hd[p][l0 + j].offset,
                                 size * 2u, &bar);
                    }
                }
            }
            mbar_wait(&bar, parity);
            parity ^= 1u;
            loaded = type;
        }
        const size_t base = (size_t)tile * plan.tile;
        const uint32_t nloc = (uint32_t)min((size_t)plan.tile, count - base);
        // plane p reads axes (a0, a1) = xy / xz / yz of the three query positions k_flow_stage left
        // in qpos (planar, coalesced): no per-visit recomputation of the sample position
        const int a0 = p == 2 ? 1 : 0, a1 = p == 0 ? 1 : 2;
        for (uint32_t i = tid; i < nloc; i += kDynThreads) {
            const size_t li = base + i;
            float u[3], w[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                u[q] = __ldg(qpos + (size_t)(3 * q + a0) * plan.stride + li);
                w[q] = __ldg(qpos + (size_t)(3 * q + a1) * plan.stride + li);
            }
            __half* out = dyn_out + (size_t)(8 * p + l0) * plan.stride + li;
#pragma unroll 1
            for (int j = 0; j < nc; ++j) {
                const LevelArgs L = lv(cfg.hd[p][l0 + j]);
                const __half* t0 = tabs + plan.tab_off[type][j];
                const float r = L.hashed ? dyn_combo<true>(t0, L, u, w) : dyn_combo<false>(t0, L, u, w);
                __stcs(reinterpret_cast<unsigned short*>(out + (size_t)j * plan.stride),
                       __half_as_ushort(__float2half_rn(r)));
            }
        }
    }
}

// Host side: pack consecutive levels of one plane into types whose tables fit the budget, and hand
// the CTAs out in proportion to the per-sample cost of each type.  Returns false when a level does
// not fit / is not 16-byte granular (the caller then uses the plain gather stage).
int g_dyn_tile = 32768;      // samples per work item (option "dyn_tile"); halved for small inputs, see make_dyn_plan
                             // (whole-frame launches, 52 M samples: 8192 -> 7.3 ms, 32768 -> 6.8 ms, 262144 -> 7.2 ms)
int g_dyn_overhead = 12;     // per-sample fixed cost in gathers  (option "dyn_overhead")
size_t g_split_chunk = kSplitChunk;  // samples per chunk (option "split_chunk", units of 64 K): a whole
                                     // LiDAR frame (52 M samples) is one chunk, 29.4 -> 27.x ms against 4 M chunks

bool make_dyn_plan(const nvsf_field_config_t* cfg, const FieldPtrs& P, size_t count, size_t stride,
                   int ctas, DynPlan& plan) {
    int nt = 0;
    for (int p = 0; p < 3; ++p) {
        int l = 0;
        while (l < kHdLevels) {
            if (nt == kDynMaxTypes) return false;
            size_t used = 0;
            int nc = 0;
            while (l + nc < kHdLevels && nc < kDynMaxCombo) {
                const nvsf_grid_level_t& g = cfg->hd[p][l + nc];
                const size_t bytes = (size_t)3 * g.size * sizeof(__half);
                if ((g.size * sizeof(__half)) % 16 || (g.offset * sizeof(__half)) % 16) return false;
                if (used + bytes > kDynTableBytes) break;
                plan.tab_off[nt][nc] = (uint32_t)(used / sizeof(__half));
                used += bytes;
                ++nc;
            }
            if (nc == 0) return false;
            plan.plane[nt] = (uint8_t)p;
            plan.level0[nt] = (uint8_t)l;
            plan.ncombo[nt] = (uint8_t)nc;
            ++nt;
            l += nc;
        }
        if ((P.dyn_plane[p] * sizeof(__half)) % 16) return false;
    }
    if ((P.dyn_per_q * sizeof(__half)) % 16) return false;
    plan.ntypes = (uint32_t)nt;
    plan.tile = (uint32_t)g_dyn_tile;
    // small inputs: keep at least ~4 work items per CTA so that the dynamic hand-out can balance
    while (plan.tile > 2048 && (size_t)nt * ((count + plan.tile - 1) / plan.tile) < (size_t)ctas * 4) plan.tile >>= 1;
    plan.tiles = (uint32_t)((count + plan.tile - 1) / plan.tile);
    plan.stride = (uint32_t)stride;
    double total = 0.0, acc = 0.0;
    for (int t = 0; t < nt; ++t) total += 12.0 * plan.ncombo[t] + g_dyn_overhead;
    for (int t = 0; t < nt; ++t) {
        plan.cta_first[t] = (uint16_t)std::min<double>(ctas - 1, acc / total * ctas + 0.5);
        acc += 12.0 * plan.ncombo[t] + g_dyn_overhead;
    }
    plan.cta_first[0] = 0;
    return true;
}

// ---- stage 3: sigma MLP -------------------------------------------------------------------------
// DRAM-stream bound (256 B in, 36 B out per sample).  Every warp owns two 32-row tile buffers and
// prefetches its next 8 KB of feature rows with cp.async (16-byte LDGSTS, L1 bypass) while the
// tensor cores work on the current one; ncu before the double buffering: 54 % of the HBM peak,
// long_scoreboard on the tile loads (profiles/r01_stages_v2_ncu_full.txt).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void sigma_prefetch(__half* buf, const __half* __restrict__ feat,
                                               size_t row0, size_t count, int lane) {
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int piece = i * 32 + lane;  // 16-byte piece index within the 8 KB block
        const int r = piece >> 4, c = piece & 15;
        const bool ok = row0 + r < count;
        cp_async16(buf + r * kLdK128 + c * 8, feat + (ok ? row0 + r : 0) * kFeat + c * 8, ok);
    }
}

__global__ void __launch_bounds__(kSigTile, 1)
k_sigma_stage(const __half* __restrict__ mlp, const __half* __restrict__ feat, size_t count,
              float* __restrict__ sigma_out, __half* __restrict__ geo_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* Wsm = reinterpret_cast<__half*>(smem_raw);   // image starting at kSigW1
    __half* Xs = Wsm + kSigWHalves;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const __half* W1 = Wsm;                          // [64][136]
    const __half* W2 = Wsm + (kSigW2 - kSigW1);      // [16][72]
    __half* bufs[2] = {Xs + (size_t)(2 * warp) * 32 * kLdK128, Xs + (size_t)(2 * warp + 1) * 32 * kLdK128};
    const size_t n_tiles = (count + kSigTile - 1) / kSigTile;
    size_t tile = blockIdx.x;
    if (tile < n_tiles) sigma_prefetch(bufs[0], feat, tile * kSigTile + warp * 32, count, lane);
    cp_async_commit();
    block_copy16(Wsm, mlp + kSigW1, kSigWHalves * 2 / 16, tid, kSigTile);
    __syncthreads();
    int cur = 0;
    for (; tile < n_tiles; tile += gridDim.x, cur ^= 1) {
        const size_t row0 = tile * kSigTile + warp * 32;
        const size_t next = tile + gridDim.x;
        if (next < n_tiles) sigma_prefetch(bufs[cur ^ 1], feat, next * kSigTile + warp * 32, count, lane);
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const __half* Aw = bufs[cur];
        float acc[2][8][4];
        zero_acc<8>(acc);
#pragma unroll
        for (int kk = 0; kk < kFeat / 16; ++kk) {
            uint32_t a[2][1][4];
            ldsm_x4(a[0][0], Aw + (lane & 15) * kLdK128 + kk * 16 + (lane >> 4) * 8);
            ldsm_x4(a[1][0], Aw + (16 + (lane & 15)) * kLdK128 + kk * 16 + (lane >> 4) * 8);
            warp_gemm_regA<1, 8>(a, W1 + kk * 16, kLdK128, acc, lane);
        }
        __syncwarp();  // every lane has read its fragments before the buffer is refilled
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
                        if (j == 0 && tq == 0) sigma_out[row] = expf(round_f16(c0));
                    }
                }
            }
    }
    cp_async_wait<0>();
}

// Optional per-stage timing with CUDA events on the launching stream (bench.py's roofline): five
// events per chunk (flow | dyn | encode | sigma), read back (and reset) by nvsf_stage_timing_read.
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
constexpr int kProfEvents = 5;

bool g_attr = false;
int ensure_attrs() {
    if (g_attr) return NVSF_OK;
    cudaError_t e;
    e = cudaFuncSetAttribute(k_flow_stage<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kFlowStageSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_flow_stage<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kFlowStageSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_flow_stage<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kFlowStageSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_flow_stage<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kFlowStageSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_sigma_stage, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kSigmaStageSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_dyn_stage<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kDynTableBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_dyn_stage<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kDynTableBytes);
    if (e != cudaSuccess) return (int)e;
    g_attr = true;
    return NVSF_OK;
}

int g_fuse_keep = 0;   // option "fuse_keep": training forward through the fused gather + sigma kernel, which then
                       // also writes the kept feature rows; measured neutral (train step 21.04 vs 21.19 ms), so the
                       // default stays k_encode_stage -> k_sigma_stage_tc
int g_half_math = 0;   // option "half_math": packed-half interpolation in the fused gather + sigma stage
                       // (k_encode_sigma_tc<true>): 27 % fewer instructions, measured 12.53 -> 12.67 ms —
                       // the stage is bound by the L1 data pipe, not by issue slots; kept for A/B runs
int g_flow_tc = 1;     // mode 2: flow stage on tcgen05 (sigma_tc.cu k_flow_tc, option "flow_tc")
int g_fuse_sigma = 1;  // mode 2: gather stage fused with the sigma MLP on tcgen05 (option "fuse_sigma")
int g_sigma_tc = 1;  // sigma stage on tcgen05 / TMEM (sigma_tc.cu) instead of mma.sync (option "sigma_tc"):
                     // 3.33 -> 2.51 ms per LiDAR frame on B200, at the DRAM floor of the 292 B/sample it streams
int g_enc_pair = 0;  // paired x-corner loads of the static hash (option "enc_pair"); measured neutral on B200
                     // (13.10 vs 13.16 ms per frame in the gather stage), so off by default

template <bool FROM_RAYS, class... A>
void launch_encode(bool dyn_pre, bool pair, unsigned tiles, cudaStream_t stream, A... a) {
    if (dyn_pre) {
        if (pair) k_encode_stage<FROM_RAYS, true, true><<<tiles, 256, 0, stream>>>(a...);
        else k_encode_stage<FROM_RAYS, true, false><<<tiles, 256, 0, stream>>>(a...);
    } else {
        if (pair) k_encode_stage<FROM_RAYS, false, true><<<tiles, 256, 0, stream>>>(a...);
        else k_encode_stage<FROM_RAYS, false, false><<<tiles, 256, 0, stream>>>(a...);
    }
}

// split scratch: flow f32 [chunk,8] | feats f16 [chunk,128] | dyn f16 [24][chunk] | qpos f32 [9][chunk]
// | tile counters
struct SplitScratch {
    size_t flow, feat, dyn, qpos, counters, total;
};
SplitScratch split_layout(size_t n, bool keep = false) {
    const size_t chunk = std::min<size_t>(n, kSplitChunk);   // sized for the largest chunk option
    SplitScratch L;
    size_t off = 0;
    // the training forward keeps flow / feats of every sample itself: only the mode-2 intermediates
    L.flow = off; off += keep ? 0 : ws_align(chunk * 8 * sizeof(float));
    L.feat = off; off += keep ? 0 : ws_align(std::min(chunk, kFeatChunk) * kFeat * sizeof(__half));
    L.dyn = off; off += ws_align(chunk * 3 * kHdLevels * sizeof(__half));
    L.qpos = off; off += ws_align(chunk * 9 * sizeof(float));
    L.counters = off; off += ws_align(kDynMaxTypes * sizeof(uint32_t));
    L.total = off;
    return L;
}

}  // namespace

size_t nvsf_density_keep_scratch_bytes(size_t n) { return split_layout(n, true).total; }
size_t nvsf_density_split_scratch_bytes(size_t n) { return split_layout(n).total; }

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
    const SplitScratch SL = split_layout(n, keep != nullptr);
    size_t chunk = std::min<size_t>(n, std::min<size_t>(g_split_chunk, kSplitChunk));
    unsigned char* sc = reinterpret_cast<unsigned char*>(split_scratch);
    float* flow_buf = reinterpret_cast<float*>(sc + SL.flow);
    __half* feat_buf = reinterpret_cast<__half*>(sc + SL.feat);
    __half* dyn_buf = reinterpret_cast<__half*>(sc + SL.dyn);
    float* qpos_buf = reinterpret_cast<float*>(sc + SL.qpos);
    uint32_t* counters = reinterpret_cast<uint32_t*>(sc + SL.counters);
    // mode 2: fp16 table mirrors, dynamic hashes from shared-memory-staged tables (needs the scratch
    // buffer for the dyn rows and query positions; the training forward passes the lean layout)
    const bool want_dyn = split_scratch != nullptr && nvsf_density_mode() == 2;
    {   // only the fused gather + sigma kernel (and the training forward, which keeps its rows itself)
        // can take chunks beyond the [kFeatChunk,128] feature scratch of the un-fused path
        DynPlan probe;
        const bool may_fuse = want_dyn && !features && g_fuse_sigma != 0 &&
                              make_dyn_plan(cfg, P, chunk, chunk, sms, probe);
        if (!may_fuse && !keep) chunk = std::min(chunk, kFeatChunk);
    }
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
        DynPlan plan;
        const bool dyn_pre = want_dyn && make_dyn_plan(cfg, P, count, count, sms, plan);
        const int grid_d = dyn_pre ? (int)std::min<size_t>((size_t)plan.ntypes * plan.tiles, (size_t)sms) : 0;
        if (dyn_pre) cudaMemsetAsync(counters, 0, kDynMaxTypes * sizeof(uint32_t), stream);
        const unsigned short* dyn_in = reinterpret_cast<const unsigned short*>(dyn_buf);
        const bool fused = dyn_pre && !features && g_fuse_sigma != 0 && (!keep || g_fuse_keep != 0);
        const bool flow_tc = dyn_pre && g_flow_tc != 0;
        // the fused gather stage reads the (warped) query positions only: the 32 B/sample flow rows
        // are written only when somebody reads them (1.7 GB per LiDAR frame)
        float* flow_dst = (fused && !flow && !keep) ? nullptr : flow_buf;
        if (g_prof.on) g_prof.next(stream);
        if (x) {
            if (flow_tc) {
                st = nvsf_launch_flow_tc(cfg, P, x, nullptr, nullptr, nullptr, nullptr, nullptr, 1, begin,
                                         count, flow_dst, qpos_buf, count, ff, sms, stream);
                if (st != NVSF_OK) return st;
            } else if (dyn_pre)
                k_flow_stage<false, true><<<grid_p, kSTile, kFlowStageSmem, stream>>>(
                    *cfg, P, x, nullptr, nullptr, nullptr, nullptr, nullptr, 1, begin, count, flow_buf, ff,
                    qpos_buf, count);
            else
                k_flow_stage<false, false><<<grid_p, kSTile, kFlowStageSmem, stream>>>(
                    *cfg, P, x, nullptr, nullptr, nullptr, nullptr, nullptr, 1, begin, count, flow_buf, ff,
                    nullptr, 0);
            if (g_prof.on) g_prof.next(stream);
            if (dyn_pre)
                k_dyn_stage<false><<<grid_d, kDynThreads, kDynTableBytes, stream>>>(
                    *cfg, P, plan, x, nullptr, nullptr, nullptr, nullptr, nullptr, 1, begin, count,
                    qpos_buf, dyn_buf, counters);
            if (g_prof.on) g_prof.next(stream);
            if (!fused) {
                launch_encode<false>(dyn_pre, g_enc_pair != 0, (unsigned)tiles, stream, *cfg, P, x,
                                     nullptr, nullptr, nullptr, nullptr, nullptr, 1, begin, count,
                                     flow_buf, feat_buf, dyn_pre ? dyn_in : nullptr, dyn_pre ? count : 0);
            }
        } else {
            if (flow_tc) {
                st = nvsf_launch_flow_tc(cfg, P, nullptr, rays_o, rays_d, nears, fars, noise, S, begin,
                                         count, flow_dst, qpos_buf, count, ff, sms, stream);
                if (st != NVSF_OK) return st;
            } else if (dyn_pre)
                k_flow_stage<true, true><<<grid_p, kSTile, kFlowStageSmem, stream>>>(
                    *cfg, P, nullptr, rays_o, rays_d, nears, fars, noise, S, begin, count, flow_buf, ff,
                    qpos_buf, count);
            else
                k_flow_stage<true, false><<<grid_p, kSTile, kFlowStageSmem, stream>>>(
                    *cfg, P, nullptr, rays_o, rays_d, nears, fars, noise, S, begin, count, flow_buf, ff,
                    nullptr, 0);
            if (g_prof.on) g_prof.next(stream);
            if (dyn_pre)
                k_dyn_stage<true><<<grid_d, kDynThreads, kDynTableBytes, stream>>>(
                    *cfg, P, plan, nullptr, rays_o, rays_d, nears, fars, noise, S, begin, count,
                    qpos_buf, dyn_buf, counters);
            if (g_prof.on) g_prof.next(stream);
            if (!fused) {
                launch_encode<true>(dyn_pre, g_enc_pair != 0, (unsigned)tiles, stream, *cfg, P, nullptr,
                                    rays_o, rays_d, nears, fars, noise, S, begin, count, flow_buf,
                                    feat_buf, dyn_pre ? dyn_in : nullptr, dyn_pre ? count : 0);
            }
        }
        if (fused) {  // gather stage + sigma MLP in one tcgen05 kernel: the feature rows stay on the SM
            st = nvsf_launch_encode_sigma_tc(cfg, P, qpos_buf, dyn_buf, count, count, sigma + begin,
                                             reinterpret_cast<__half*>(geo) + begin * kGeo, sms, stream, g_half_math,
                                             keep ? feat_buf : nullptr);
            if (st != NVSF_OK) return st;
        }
        if (g_prof.on) g_prof.next(stream);
        if (fused) {
            // sigma and geo were produced by the fused kernel
        } else if (g_sigma_tc) {
            st = nvsf_launch_sigma_tc(P.mlp_tc, feat_buf, count, sigma + begin,
                                      reinterpret_cast<__half*>(geo) + begin * kGeo, sms, stream);
            if (st != NVSF_OK) return st;
        } else {
            k_sigma_stage<<<(int)std::min<size_t>((count + kSigTile - 1) / kSigTile, (size_t)sms), kSigTile,
                            kSigmaStageSmem, stream>>>(
                P.mlp, feat_buf, count, sigma + begin, reinterpret_cast<__half*>(geo) + begin * kGeo);
        }
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

// ---- tuning options of the staged evaluation (nvsf_set_option falls through to here) -------------
int nvsf_split_set_option(const char* name, int value) {
    const std::string k(name);
    if (k == "dyn_tile") {
        if (value < 1024 || value > (1 << 22)) return NVSF_E_INVALID;
        g_dyn_tile = value;
        return NVSF_OK;
    }
    if (k == "dyn_overhead") {
        if (value < 0 || value > 1000) return NVSF_E_INVALID;
        g_dyn_overhead = value;
        return NVSF_OK;
    }
    if (k == "flow_tc") {
        if (value != 0 && value != 1) return NVSF_E_INVALID;
        g_flow_tc = value;
        return NVSF_OK;
    }
    if (k == "flow_ts") {
        if (value != 0 && value != 1) return NVSF_E_INVALID;
        g_flow_ts = value;
        return NVSF_OK;
    }
    if (k == "half_math") {
        if (value != 0 && value != 1) return NVSF_E_INVALID;
        g_half_math = value;
        return NVSF_OK;
    }
    if (k == "fuse_keep") {
        if (value != 0 && value != 1) return NVSF_E_INVALID;
        g_fuse_keep = value;
        return NVSF_OK;
    }
    if (k == "fuse_sigma") {
        if (value != 0 && value != 1) return NVSF_E_INVALID;
        g_fuse_sigma = value;
        return NVSF_OK;
    }
    if (k == "sigma_tc") {
        if (value != 0 && value != 1) return NVSF_E_INVALID;
        g_sigma_tc = value;
        return NVSF_OK;
    }
    if (k == "enc_pair") {
        if (value != 0 && value != 1) return NVSF_E_INVALID;
        g_enc_pair = value;
        return NVSF_OK;
    }
    if (k == "split_chunk") {  // units of 64 K samples, at most kSplitChunk
        if (value < 1 || (size_t)value * 65536 > kSplitChunk) return NVSF_E_INVALID;
        g_split_chunk = (size_t)value * 65536;
        return NVSF_OK;
    }
    return nvsf_train_set_option(name, value);
}

int nvsf_split_get_option(const char* name) {
    const std::string k(name);
    if (k == "dyn_tile") return g_dyn_tile;
    if (k == "dyn_overhead") return g_dyn_overhead;
    if (k == "split_chunk") return (int)(g_split_chunk / 65536);
    if (k == "enc_pair") return g_enc_pair;
    if (k == "sigma_tc") return g_sigma_tc;
    if (k == "fuse_sigma") return g_fuse_sigma;
    if (k == "flow_tc") return g_flow_tc;
    if (k == "flow_ts") return g_flow_ts;
    if (k == "half_math") return g_half_math;
    if (k == "fuse_keep") return g_fuse_keep;
    return nvsf_train_get_option(name);
}

// ---- stage timing (see StageProf) -------------------------------------------------------------------
void nvsf_stage_timing_enable(int on) {
    g_prof.on = on != 0;
    g_prof.used = 0;
}

extern "C" int nvsf_stage_timing_read(float* ms4, uint32_t* launches) {
    if (!ms4) return NVSF_E_INVALID;
    ms4[0] = ms4[1] = ms4[2] = ms4[3] = 0.f;
    const size_t groups = g_prof.used / kProfEvents;
    for (size_t g = 0; g < groups; ++g) {
        cudaError_t e = cudaEventSynchronize(g_prof.ev[kProfEvents * g + kProfEvents - 1]);
        if (e != cudaSuccess) return (int)e;
        for (int k = 0; k < kProfEvents - 1; ++k) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, g_prof.ev[kProfEvents * g + k], g_prof.ev[kProfEvents * g + k + 1]);
            ms4[k] += ms;
        }
    }
    if (launches) *launches = (uint32_t)groups;
    g_prof.used = 0;
    return NVSF_OK;
}
