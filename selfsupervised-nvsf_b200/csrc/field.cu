// nvsf_b200 — the per-sample field (Part 2 of include/nvsf_b200.h), sm_100a.
//
// Replaces NeRFNetwork.density (reference nvsf/nerf/models/network_dynamic.py:213-287) and the
// tcnn / grid_sample pipeline under it (hash_field.py, planes_field.py, flow_field.py).
//
// B200-first design, not a translation of the reference's ~100 kernel launches per call:
//  * Everything that depends only on the frame time t — the time-slice blend and the cubic
//    temporal basis of the dynamic hash grids and of the flow grid, and the time rows of the
//    (x,t),(y,t),(z,t) planes — is LINEAR in the tables, so it is collapsed once per call into
//    small tables (k_collapse_*).  The per-sample work then gathers 4-byte scalars from a
//    1.5 MB table instead of two 8-byte vectors from 25 MB of time slices, and 1-D instead of
//    2-D plane lookups: 6.5 KB gathered per sample instead of 13.3 KB, all of it L2 resident.
//  * One persistent kernel evaluates, for tiles of 128 samples: flow-grid features -> flow MLP
//    (tensor cores) -> the three warped queries -> 120 features -> sigma MLP (tensor cores) ->
//    trunc_exp.  The [N,120] feature tensor never leaves shared memory.
//  * Tables are fp16 (static hash, like tcnn) or fp32 (collapsed tables, planes); MLPs run in
//    fp16 with fp32 accumulation on mma.sync m16n8k16.
#include <algorithm>
#include <string>

#include "field_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// time setup + packing kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_time_setup(const float* __restrict__ time, uint32_t num_frames, uint32_t tres,
                             TimeInfo* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    TimeInfo ti;
    const float t = *time;
    const int f = (int)(t * (float)(num_frames - 1));  // network_dynamic.py:218
    ti.t[0] = t;
    ti.t[1] = (float)((double)(f + 1) / (double)num_frames);
    ti.t[2] = (float)((double)(f - 1) / (double)num_frames);
    ti.valid[0] = 1;
    ti.valid[1] = f < (int)num_frames - 1;
    ti.valid[2] = f > 0;
    for (int q = 0; q < 3; ++q) {
        const float tq = ti.t[q];
        for (int j = 0; j < 4; ++j) {
            double w = 1.0;
            for (int m = 0; m < 4; ++m)
                if (m != j) w *= ((double)tq - m / 3.0) / (j / 3.0 - m / 3.0);
            ti.lag[q][j] = (float)w;
        }
        const float idx = tq * (float)(tres - 1);  // hash_field.py:79
        const float k1 = floorf(idx), k2 = ceilf(idx);
        const int i1 = min(max((int)k1, 0), (int)tres - 1);
        const int i2 = min(max((int)k2, 0), (int)tres - 1);
        ti.k1[q] = i1;
        ti.k2[q] = i2;
        ti.wk[q] = (i1 == i2) ? 0.f : idx - k1;
        // grid_sample(align_corners=True, padding_mode='border') along the time axis
        float iy = ((tq * 2.0f - 1.0f) + 1.0f) * 0.5f * (float)(tres - 1);
        iy = fminf(fmaxf(iy, 0.f), (float)(tres - 1));
        const float y0 = floorf(iy);
        ti.y0[q] = (int)y0;
        ti.y1[q] = min((int)y0 + 1, (int)tres - 1);
        ti.wy[q] = iy - y0;
    }
    *out = ti;
}

__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }

// static hash: fp32 master -> fp16 table (8 values per thread)
__global__ void k_pack_hash_static(const float* __restrict__ src, __half* __restrict__ dst,
                                   size_t n8) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const float4 a = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 o;
    o.x = pack_half2(a.x, a.y); o.y = pack_half2(a.z, a.w);
    o.z = pack_half2(b.x, b.y); o.w = pack_half2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = o;
}

struct PlaneSrc {  // offsets (floats) of the reference's [1,8,H,W] tensors inside `planes`
    uint32_t off[kPlScales][6];
};

// space planes (combinations xy, xz, yz) [8][R][R] -> channel-last [R][R][8]
__global__ void k_pack_planes_static(const float* __restrict__ planes, PlaneSrc src,
                                     uint32_t res, uint32_t scale, float* __restrict__ dst,
                                     __half* __restrict__ dst16) {
    const uint32_t combo = blockIdx.y == 0 ? 0u : (blockIdx.y == 1 ? 1u : 3u);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= res * res) return;
    const float* s = planes + src.off[scale][combo];
    float v[8];
#pragma unroll
    for (int f = 0; f < 8; ++f) v[f] = __ldg(s + (size_t)f * res * res + i);
    float4* d = reinterpret_cast<float4*>(dst + ((size_t)blockIdx.y * res * res + i) * 8);
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
    reinterpret_cast<uint4*>(dst16)[(size_t)blockIdx.y * res * res + i] =
        make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]),
                   pack_half2(v[6], v[7]));
}

// time planes (combinations xt, yt, zt) [8][Tres][R] -> per query: time-blended rows [R][8]
__global__ void k_collapse_planes_dyn(const float* __restrict__ planes, PlaneSrc src,
                                      uint32_t res, uint32_t tres, uint32_t scale,
                                      const TimeInfo* __restrict__ ti, float* __restrict__ dst,
                                      __half* __restrict__ dst16, size_t per_q) {
    const uint32_t p = blockIdx.y, q = blockIdx.z;
    const uint32_t combo = p == 0 ? 2u : (p == 1 ? 4u : 5u);
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= res) return;
    const float* s = planes + src.off[scale][combo];
    const int y0 = ti->y0[q], y1 = ti->y1[q];
    const float w = ti->wy[q];
    float v[8];
#pragma unroll
    for (int f = 0; f < 8; ++f) {
        const float a = __ldg(s + ((size_t)f * tres + y0) * res + x);
        const float b = __ldg(s + ((size_t)f * tres + y1) * res + x);
        v[f] = (1.0f - w) * a + w * b;
    }
    float4* d = reinterpret_cast<float4*>(dst + q * per_q + ((size_t)p * res + x) * 8);
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
    *reinterpret_cast<uint4*>(dst16 + q * per_q + ((size_t)p * res + x) * 8) =
        make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]),
                   pack_half2(v[6], v[7]));
}

// dynamic hash: out[q][e] = sum_i lag[q][i] * ((k2-idx) * G[k1][e][i] + (idx-k1) * G[k2][e][i])
// (HashGridT.forward + interpT, hash_field.py:65-88; both are linear in the table entries)
__global__ void k_collapse_dyn(const float* __restrict__ slices /* [Tres][entries][4] */,
                               uint32_t entries, const TimeInfo* __restrict__ ti,
                               float* __restrict__ dst /* + q*per_q */,
                               __half* __restrict__ dst16 /* fp16 mirror, same indexing */,
                               size_t per_q) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= entries) return;
    const float4* tab = reinterpret_cast<const float4*>(slices);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        if (!ti->valid[q]) continue;
        const int k1 = ti->k1[q], k2 = ti->k2[q];
        const float w = ti->wk[q];
        float4 a = __ldg(tab + (size_t)k1 * entries + e);
        a = make_float4(round_h(a.x), round_h(a.y), round_h(a.z), round_h(a.w));
        float4 v = a;
        if (k1 != k2) {
            float4 b = __ldg(tab + (size_t)k2 * entries + e);
            b = make_float4(round_h(b.x), round_h(b.y), round_h(b.z), round_h(b.w));
            const float wa = 1.0f - w;  // == k2 - idx
            v = make_float4(wa * a.x + w * b.x, wa * a.y + w * b.y, wa * a.z + w * b.z,
                            wa * a.w + w * b.w);
        }
        const float o = ti->lag[q][0] * v.x + ti->lag[q][1] * v.y + ti->lag[q][2] * v.z +
                        ti->lag[q][3] * v.w;
        dst[q * per_q + e] = o;
        dst16[q * per_q + e] = __float2half_rn(o);
    }
}

// flow grid: out[e][c] = sum_i lag[0][i] * F[e][2i+c]   (FlowField.interpT, flow_field.py:105-114)
__global__ void k_collapse_flow(const float* __restrict__ grid /* [entries][8] */,
                                uint32_t entries, const TimeInfo* __restrict__ ti,
                                float2* __restrict__ dst, __half2* __restrict__ dst16) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= entries) return;
    const float4 a = __ldcs(reinterpret_cast<const float4*>(grid) + 2 * (size_t)e);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(grid) + 2 * (size_t)e + 1);
    const float l0 = ti->lag[0][0], l1 = ti->lag[0][1], l2 = ti->lag[0][2], l3 = ti->lag[0][3];
    float2 o;
    o.x = l0 * round_h(a.x) + l1 * round_h(a.z) + l2 * round_h(b.x) + l3 * round_h(b.z);
    o.y = l0 * round_h(a.y) + l1 * round_h(a.w) + l2 * round_h(b.y) + l3 * round_h(b.w);
    dst[e] = o;
    dst16[e] = __floats2half2_rn(o.x, o.y);
}

// fp32 [rows][src_ld] sub-block -> fp16 [rows][dst_ld] sub-block
__global__ void k_pack_matrix(const float* __restrict__ src, int src_ld, int src_col0, int rows,
                              int cols, __half* __restrict__ dst, int dst_ld, int dst_col0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int r = i / cols, c = i - r * cols;
    dst[r * dst_ld + dst_col0 + c] = __float2half_rn(__ldg(src + r * src_ld + src_col0 + c));
}

// dst[r][dst_col] = fp16(sum_c fp16(src[r][col0 + c])): the first-layer weight columns behind tcnn's
// constant-1 input padding, folded into one column (see kOneH in field_common.cuh)
__global__ void k_pack_padsum(const float* __restrict__ src, int src_ld, int col0, int ncols, int rows,
                              __half* __restrict__ dst, int dst_ld, int dst_col) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float s = 0.f;
    for (int c = 0; c < ncols; ++c) s += round_h(__ldg(src + r * src_ld + col0 + c));
    dst[r * dst_ld + dst_col] = __float2half_rn(s);
}

// ------------------------------------------------------------------------------------------------
// per-sample encoders (device)
// ------------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------------
// the density kernel
// ------------------------------------------------------------------------------------------------
constexpr int kTile = 256;  // samples per CTA tile, one warp per 32 rows
constexpr int kXld = kLdK128;
constexpr int kFlowCol = 32;  // the 6 flow outputs (fp32) are staged in row bytes [64,96)
constexpr size_t kDensitySmem = (size_t)kDensityWHalves * sizeof(__half) +
                                (size_t)kTile * kXld * sizeof(__half);

template <bool FROM_RAYS>
__global__ void __launch_bounds__(kTile, 2)
k_field_density(const __grid_constant__ nvsf_field_config_t cfg, const __grid_constant__ FieldPtrs P, const float* __restrict__ xin,
                const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                const float* __restrict__ nears, const float* __restrict__ fars,
                const float* __restrict__ noise, uint32_t S, size_t n,
                float* __restrict__ sigma_out, __half* __restrict__ geo_out,
                __half* __restrict__ feat_out, float* __restrict__ flow_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* Wsm = reinterpret_cast<__half*>(smem_raw);
    __half* Xs = Wsm + kDensityWHalves;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    block_copy16(Wsm, P.mlp, kDensityWHalves * (int)sizeof(__half) / 16, tid, kTile);
    __syncthreads();

    const int valid1 = P.ti->valid[1], valid2 = P.ti->valid[2];
    __half* xrow = Xs + tid * kXld;
    const __half* Aw = Xs + warp * 32 * kXld;
    const float inv2b = 1.0f / (2.0f * cfg.bound);

    const size_t n_tiles = (n + kTile - 1) / kTile;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t g = tile * kTile + tid;
        const bool live = g < n;
        // ---- position, normalised to [0,1]  (network_dynamic.py:217) ----
        float px = 0.f, py = 0.f, pz = 0.f;
        if (live) {
            if (FROM_RAYS) {
                const size_t r = g / S;
                const uint32_t k = (uint32_t)(g - r * S);
                const float z = uniform_z(__ldg(nears + r), __ldg(fars + r), k, S, noise, g);
                px = __ldg(rays_o + r * 3 + 0) + __ldg(rays_d + r * 3 + 0) * z;
                py = __ldg(rays_o + r * 3 + 1) + __ldg(rays_d + r * 3 + 1) * z;
                pz = __ldg(rays_o + r * 3 + 2) + __ldg(rays_d + r * 3 + 2) * z;
                px = fminf(fmaxf(px, -cfg.bound), cfg.bound);  // renderer_dynamic.py:169
                py = fminf(fmaxf(py, -cfg.bound), cfg.bound);
                pz = fminf(fmaxf(pz, -cfg.bound), cfg.bound);
            } else {
                px = __ldg(xin + g * 3 + 0); py = __ldg(xin + g * 3 + 1); pz = __ldg(xin + g * 3 + 2);
            }
        }
        const float x = (px + cfg.bound) * inv2b, y = (py + cfg.bound) * inv2b,
                    z = (pz + cfg.bound) * inv2b;

        // ---- phase 1: flow-grid features -> Xs[:, 0:32]  (flow_field.py:124-128) ----
#pragma unroll 1
        for (int l = 0; l < kFlLevels; ++l) {
            const float2 f = hash3_f2(P.flow, lv(cfg.fl[l]), x, y, z);
            *reinterpret_cast<uint32_t*>(xrow + 2 * l) = pack_half2(f.x, f.y);
        }
        __syncwarp();

        // ---- phase 2: flow MLP 32 -> 64 -> 64 -> 6 on tensor cores (flow_field.py:87-103) ----
        {
            float acc[2][8][4];
            zero_acc<8>(acc);
            {
                uint32_t a[2][2][4];
                load_a_frags<2>(Aw, kXld, a, lane);
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
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                __half* r0 = Xs + (warp * 32 + mt * 16 + gq) * kXld + kFlowCol;
                *reinterpret_cast<float2*>(reinterpret_cast<float*>(r0) + 2 * tq) =
                    make_float2(o[mt][0][0], o[mt][0][1]);
                *reinterpret_cast<float2*>(reinterpret_cast<float*>(r0 + 8 * kXld) + 2 * tq) =
                    make_float2(o[mt][0][2], o[mt][0][3]);
            }
        }
        __syncwarp();

        // ---- phase 3: the 120 sigma-net inputs ----
        float fl[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) fl[i] = round_f16(reinterpret_cast<const float*>(xrow + kFlowCol)[i]);
        __syncwarp();
        if (flow_out && live) {
#pragma unroll
            for (int i = 0; i < 6; ++i) flow_out[g * 6 + i] = fl[i];
        }
        // query positions: q=0 (x,t), q=1 (x+flow_f, t1), q=2 (x+flow_b, t2); a missing neighbour
        // frame falls back to the un-warped query (network_dynamic.py:238-239)
        float qx[3], qy[3], qz[3];
        int qi[3];
        qx[0] = x; qy[0] = y; qz[0] = z; qi[0] = 0;
        qx[1] = valid1 ? x + fl[0] : x; qy[1] = valid1 ? y + fl[1] : y;
        qz[1] = valid1 ? z + fl[2] : z; qi[1] = valid1 ? 1 : 0;
        qx[2] = valid2 ? x + fl[3] : x; qy[2] = valid2 ? y + fl[4] : y;
        qz[2] = valid2 ? z + fl[5] : z; qi[2] = valid2 ? 2 : 0;

        // (a) static planes: product of xy, xz, yz per scale -> cols [0,32)
#pragma unroll 1
        for (int s = 0; s < kPlScales; ++s) {
            const uint32_t R = cfg.pl_res[s];
            const float* base = P.pls + P.pls_scale[s];
            float v[8];
            plane2d_mul(base, R, x, y, v, true);
            plane2d_mul(base + (size_t)R * R * 8, R, x, z, v, false);
            plane2d_mul(base + (size_t)2 * R * R * 8, R, y, z, v, false);
            st8(xrow, 8 * s, v);
        }
        // (b) dynamic planes: product of xt, yt, zt per scale, 0.5 d + 0.25 (d1 + d2) -> [32,64)
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
            st8(xrow, 32 + 8 * s, acc8);
        }
        // (c) static 3-D hash -> [64,96)
#pragma unroll 1
        for (int l = 0; l < kHsLevels; ++l) {
            float v[4];
            hash3_f4(P.hs16, lv(cfg.hs[l]), x, y, z, v);
            uint2 o2;
            o2.x = pack_half2(v[0], v[1]); o2.y = pack_half2(v[2], v[3]);
            *reinterpret_cast<uint2*>(xrow + 64 + 4 * l) = o2;
        }
        // (d) dynamic 2-D hashes xy, xz, yz -> [96,120): 0.5 f(x,t) + 0.25 (f(x1,t1) + f(x2,t2))
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
#pragma unroll 1
            for (int l = 0; l < kHdLevels; ++l) {
                const LevelArgs L = lv(cfg.hd[p][l]);
                const float f0 = hash2_f1(tab[0], L, u[0], w[0]);
                const float f1 = hash2_f1(tab[1], L, u[1], w[1]);
                const float f2 = hash2_f1(tab[2], L, u[2], w[2]);
                xrow[96 + 8 * p + l] = __float2half_rn(0.5f * f0 + 0.25f * (f1 + f2));
            }
        }
        {
            const float one8[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};   // tcnn input padding = 1
            st8(xrow, 120, one8);
        }
        __syncwarp();
        if (feat_out && live) {
            const uint4* s = reinterpret_cast<const uint4*>(xrow);
            uint4* d = reinterpret_cast<uint4*>(feat_out + g * kFeat);
#pragma unroll
            for (int i = 0; i < kFeat / 8; ++i) d[i] = s[i];
        }

        // ---- phase 4: sigma MLP 128 -> 64 -> 16 (network_dynamic.py:125-135,278-282) ----
        {
            float acc[2][8][4];
            zero_acc<8>(acc);
#pragma unroll
            for (int kk = 0; kk < kFeat / 16; ++kk) {
                uint32_t a[2][1][4];
                ldsm_x4(a[0][0], Aw + (lane & 15) * kXld + kk * 16 + (lane >> 4) * 8);
                ldsm_x4(a[1][0], Aw + (16 + (lane & 15)) * kXld + kk * 16 + (lane >> 4) * 8);
                warp_gemm_regA<1, 8>(a, Wsm + kSigW1 + kk * 16, kLdK128, acc, lane);
            }
            uint32_t a2[2][4][4];
            relu_to_a<8>(acc, a2);
            float o[2][2][4];
            zero_acc<2>(o);
            warp_gemm_regA<4, 2>(a2, Wsm + kSigW2, kLdK64, o, lane);
            const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const size_t row = tile * kTile + warp * 32 + mt * 16 + hrow * 8 + gq;
                    if (row < n) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const float c0 = o[mt][j][2 * hrow], c1 = o[mt][j][2 * hrow + 1];
                            *reinterpret_cast<uint32_t*>(geo_out + row * kGeo + 8 * j + 2 * tq) =
                                pack_half2(c0, c1);
                            if (j == 0 && tq == 0) sigma_out[row] = expf(round_f16(c0));  // trunc_exp fwd
                        }
                    }
                }
        }
        __syncwarp();
    }
}

bool g_density_attr = false;
int ensure_density_attr() {
    if (g_density_attr) return NVSF_OK;
    cudaError_t e = cudaFuncSetAttribute(k_field_density<true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kDensitySmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_field_density<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kDensitySmem);
    if (e != cudaSuccess) return (int)e;
    g_density_attr = true;
    return NVSF_OK;
}

int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

PlaneSrc make_plane_src(const nvsf_field_config_t* c) {
    PlaneSrc s;
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

}  // namespace

FieldPtrs nvsf_make_field_ptrs(const nvsf_field_config_t* cfg, const void* workspace) {
    const WsLayout L = make_ws_layout(cfg);
    const unsigned char* w = reinterpret_cast<const unsigned char*>(workspace);
    FieldPtrs P;
    P.ti = reinterpret_cast<const TimeInfo*>(w + L.time);
    P.hs16 = reinterpret_cast<const uint2*>(w + L.hs16);
    P.pls = reinterpret_cast<const float*>(w + L.pls);
    P.dyn = reinterpret_cast<const float*>(w + L.dyn);
    P.dyn16 = reinterpret_cast<const __half*>(w + L.dyn16);
    P.pls16 = reinterpret_cast<const __half*>(w + L.pls16);
    P.pld16 = reinterpret_cast<const __half*>(w + L.pld16);
    P.flow16 = reinterpret_cast<const __half2*>(w + L.flow16);
    P.flow = reinterpret_cast<const float2*>(w + L.flow);
    P.pld = reinterpret_cast<const float*>(w + L.pld);
    P.mlp = reinterpret_cast<const __half*>(w + L.mlp);
    P.mlp_tc = w + L.mlp_tc;
    P.heads_tc = w + L.heads_tc;
    for (int s = 0; s < kPlScales; ++s) {
        P.pls_scale[s] = (uint32_t)L.pls_scale[s];
        P.pld_scale[s] = (uint32_t)L.pld_scale[s];
    }
    P.pld_per_q = (uint32_t)L.pld_floats_per_q;
    P.dyn_per_q = (uint32_t)L.dyn_per_q;
    for (int p = 0; p < 3; ++p) P.dyn_plane[p] = (uint32_t)L.dyn_plane[p];
    return P;
}

// Launch the density kernel (shared with the renderer in render.cu).
int nvsf_launch_density(const nvsf_field_config_t* cfg, const void* workspace, const float* x,
                        const float* rays_o, const float* rays_d, const float* nears,
                        const float* fars, const float* noise, uint32_t S, size_t n, float* sigma,
                        void* geo, void* features, float* flow, void* split_scratch,
                        cudaStream_t stream) {
    if (split_scratch && nvsf_density_mode() >= 1)
        return nvsf_launch_density_split(cfg, workspace, x, rays_o, rays_d, nears, fars, noise, S, n,
                                         sigma, geo, features, flow, split_scratch, stream);
    int st = ensure_density_attr();
    if (st != NVSF_OK) return st;
    const FieldPtrs P = nvsf_make_field_ptrs(cfg, workspace);
    const size_t n_tiles = (n + kTile - 1) / kTile;
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_field_density<true>, kTile,
                                                      kDensitySmem);
        if (blocks_per_sm <= 0) blocks_per_sm = 1;
    }
    const int grid = (int)std::min<size_t>(n_tiles, (size_t)num_sms() * blocks_per_sm);
    if (x) {
        k_field_density<false><<<grid, kTile, kDensitySmem, stream>>>(
            *cfg, P, x, nullptr, nullptr, nullptr, nullptr, nullptr, 1, n, sigma,
            reinterpret_cast<__half*>(geo), reinterpret_cast<__half*>(features), flow);
    } else {
        k_field_density<true><<<grid, kTile, kDensitySmem, stream>>>(
            *cfg, P, nullptr, rays_o, rays_d, nears, fars, noise, S, n, sigma,
            reinterpret_cast<__half*>(geo), reinterpret_cast<__half*>(features), flow);
    }
    return nvsf_launch_status();
}

extern "C" {

size_t nvsf_field_workspace_bytes(const nvsf_field_config_t* cfg) {
    if (!field_cfg_ok(cfg)) return 0;
    return make_ws_layout(cfg).total;
}

int nvsf_field_pack_params(const nvsf_field_config_t* cfg, const nvsf_field_params_t* prm,
                           uint32_t lidar, void* workspace, size_t workspace_bytes,
                           void* stream) {
    if (!field_cfg_ok(cfg) || !prm || !workspace) return NVSF_E_INVALID;
    if (!prm->hash_static || !prm->planes || !prm->flow_mlp || !prm->sigma_net || !prm->head_a ||
        (lidar && !prm->head_b))
        return NVSF_E_INVALID;
    if ((cfg->hs_entries * kHashF) % 8 != 0) return NVSF_E_INVALID;
    const WsLayout L = make_ws_layout(cfg);
    if (workspace_bytes < L.total) return NVSF_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    // static hash table -> fp16
    {
        const size_t n8 = (size_t)cfg->hs_entries * kHashF / 8;
        k_pack_hash_static<<<(unsigned)nvsf_div_up(n8, (size_t)256), 256, 0, s>>>(
            prm->hash_static, reinterpret_cast<__half*>(w + L.hs16), n8);
    }
    // space planes -> channel-last
    const PlaneSrc src = make_plane_src(cfg);
    for (int sc = 0; sc < kPlScales; ++sc) {
        const uint32_t R = cfg->pl_res[sc];
        dim3 grid(nvsf_div_up(R * R, 256u), 3);
        k_pack_planes_static<<<grid, 256, 0, s>>>(
            prm->planes, src, R, sc, reinterpret_cast<float*>(w + L.pls) + L.pls_scale[sc],
            reinterpret_cast<__half*>(w + L.pls16) + L.pls_scale[sc]);
    }
    // MLP weights -> fp16 shared-memory images
    __half* m = reinterpret_cast<__half*>(w + L.mlp);
    cudaMemsetAsync(m, 0, (size_t)kMlpHalves * sizeof(__half), s);
    auto pack = [&](const float* srcp, int src_ld, int src_col0, int rows, int cols, __half* dst,
                    int dst_ld, int dst_col0) {
        k_pack_matrix<<<nvsf_div_up(rows * cols, 256), 256, 0, s>>>(srcp, src_ld, src_col0, rows,
                                                                    cols, dst, dst_ld, dst_col0);
    };
    const int H = kHidden;
    // flow MLP: [64x32], [64x64], [6x64]
    pack(prm->flow_mlp, kFlowIn, 0, H, kFlowIn, m + kFlowW1, kLdK32, 0);
    pack(prm->flow_mlp + H * kFlowIn, H, 0, H, H, m + kFlowW2, kLdK64, 0);
    pack(prm->flow_mlp + H * kFlowIn + H * H, H, 0, 6, H, m + kFlowW3, kLdK64, 0);
    // sigma net: [64x128], [16x64]
    pack(prm->sigma_net, kFeat, 0, H, kFeat, m + kSigW1, kLdK128, 0);
    pack(prm->sigma_net + H * kFeat, H, 0, kGeo, H, m + kSigW2, kLdK64, 0);
    // heads: layer 1 is split into its direction columns and its geo columns
    const int n_dir = lidar ? 72 : 16, in_pad = lidar ? 96 : 32, n_out = lidar ? 1 : 3;
    const float* heads[2] = {prm->head_a, lidar ? prm->head_b : nullptr};
    for (int h = 0; h < 2; ++h) {
        if (!heads[h]) continue;
        __half* hm = m + kHeadBase + h * kHeadHalves;
        pack(heads[h], in_pad, 0, H, n_dir, hm + kHeadW1d, kHeadDirMax, 0);
        pack(heads[h], in_pad, n_dir, H, 15, hm + kHeadW1g, kLdK16, 1);  // geo col 0 is the logit ...
        // ... which the head kernels replace by the constant 1 of tcnn's input padding (87 -> 96, 31 -> 32)
        k_pack_padsum<<<1, H, 0, s>>>(heads[h], in_pad, n_dir + 15, in_pad - n_dir - 15, H, hm + kHeadW1g, kLdK16, 0);
        pack(heads[h] + H * in_pad, H, 0, H, H, hm + kHeadW2, kLdK64, 0);
        pack(heads[h] + H * in_pad + H * H, H, 0, n_out, H, hm + kHeadW3, kLdK64, 0);
    }
    nvsf_pack_sigma_tc(m, w + L.mlp_tc, s);  // swizzled K-major operand images of the sigma net
    nvsf_pack_heads_tc(m, w + L.heads_tc, lidar ? 2 : 1, s);
    return nvsf_launch_status();
}

int nvsf_field_pack_time(const nvsf_field_config_t* cfg, const nvsf_field_params_t* prm,
                         const float* time, void* workspace, size_t workspace_bytes,
                         void* stream) {
    if (!field_cfg_ok(cfg) || !prm || !workspace || !time) return NVSF_E_INVALID;
    if (!prm->hash_dynamic || !prm->planes || !prm->flow_grid) return NVSF_E_INVALID;
    const WsLayout L = make_ws_layout(cfg);
    if (workspace_bytes < L.total) return NVSF_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    TimeInfo* ti = reinterpret_cast<TimeInfo*>(w + L.time);
    k_time_setup<<<1, 32, 0, s>>>(time, cfg->num_frames, cfg->time_resolution, ti);
    // dynamic hash grids
    const float* slices = prm->hash_dynamic;
    for (int p = 0; p < 3; ++p) {
        const uint32_t e = cfg->hd_entries[p];
        k_collapse_dyn<<<nvsf_div_up(e, 256u), 256, 0, s>>>(
            slices, e, ti, reinterpret_cast<float*>(w + L.dyn) + L.dyn_plane[p],
            reinterpret_cast<__half*>(w + L.dyn16) + L.dyn_plane[p], L.dyn_per_q);
        slices += (size_t)cfg->time_resolution * e * kHashF;
    }
    // flow grid
    k_collapse_flow<<<nvsf_div_up(cfg->fl_entries, 256u), 256, 0, s>>>(
        prm->flow_grid, cfg->fl_entries, ti, reinterpret_cast<float2*>(w + L.flow),
        reinterpret_cast<__half2*>(w + L.flow16));
    // time planes
    const PlaneSrc src = make_plane_src(cfg);
    for (int sc = 0; sc < kPlScales; ++sc) {
        const uint32_t R = cfg->pl_res[sc];
        dim3 grid(nvsf_div_up(R, 128u), 3, 3);
        k_collapse_planes_dyn<<<grid, 128, 0, s>>>(
            prm->planes, src, R, cfg->time_resolution, sc, ti,
            reinterpret_cast<float*>(w + L.pld) + L.pld_scale[sc],
            reinterpret_cast<__half*>(w + L.pld16) + L.pld_scale[sc], L.pld_floats_per_q);
    }
    return nvsf_launch_status();
}

size_t nvsf_field_density_scratch_bytes(uint32_t n) { return nvsf_density_split_scratch_bytes(n); }

int nvsf_field_density(const nvsf_field_config_t* cfg, const void* workspace, const float* x,
                       uint32_t n, float* sigma, void* geo, void* features, float* flow,
                       void* scratch, size_t scratch_bytes, void* stream) {
    if (n == 0) return NVSF_OK;
    if (!field_cfg_ok(cfg) || !workspace || !x || !sigma || !geo) return NVSF_E_INVALID;
    if (scratch && scratch_bytes < nvsf_density_split_scratch_bytes(n)) return NVSF_E_WORKSPACE;
    return nvsf_launch_density(cfg, workspace, x, nullptr, nullptr, nullptr, nullptr, nullptr, 1, n,
                               sigma, geo, features, flow, scratch, (cudaStream_t)stream);
}

static int g_density_mode_value = 2;
static int g_march_mode_value = 1;
static int g_composite_mode_value = 2, g_composite_bwd_mode_value = 2;
int nvsf_set_option(const char* name, int value) {
    if (!name) return NVSF_E_INVALID;
    if (std::string(name) == "density_mode") {
        if (value < 0 || value > 2) return NVSF_E_INVALID;
        g_density_mode_value = value;
        return NVSF_OK;
    }
    if (std::string(name) == "march_mode") {
        if (value != 0 && value != 1) return NVSF_E_INVALID;
        g_march_mode_value = value;
        return NVSF_OK;
    }
    if (std::string(name) == "composite_mode") {
        if (value < 0 || value > 2) return NVSF_E_INVALID;
        g_composite_mode_value = value;
        return NVSF_OK;
    }
    if (std::string(name) == "composite_bwd_mode") {
        if (value < 0 || value > 2) return NVSF_E_INVALID;
        g_composite_bwd_mode_value = value;
        return NVSF_OK;
    }
    if (std::string(name) == "stage_timing") {
        nvsf_stage_timing_enable(value);
        return NVSF_OK;
    }
    return nvsf_split_set_option(name, value);
}

int nvsf_density_mode_get(void) { return g_density_mode_value; }
int nvsf_get_option(const char* name) {
    if (!name) return NVSF_E_INVALID;
    if (std::string(name) == "density_mode") return g_density_mode_value;
    if (std::string(name) == "march_mode") return g_march_mode_value;
    if (std::string(name) == "composite_mode") return g_composite_mode_value;
    if (std::string(name) == "composite_bwd_mode") return g_composite_bwd_mode_value;
    return nvsf_split_get_option(name);
}

}  // extern "C"

int nvsf_density_mode() { return g_density_mode_value; }
int nvsf_march_mode() { return g_march_mode_value; }
int nvsf_composite_mode() { return g_composite_mode_value; }
int nvsf_composite_bwd_mode() { return g_composite_bwd_mode_value; }
