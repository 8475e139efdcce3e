// nvsf_b200 — sigma MLP (120 -> 64 -> 16, tcnn FullyFusedMLP of network_dynamic.py:125-135) on the
// 5th-generation tensor cores: tcgen05.mma with both operands in shared memory (K-major, 128-byte
// swizzle), accumulators in tensor memory, epilogues through tcgen05.ld.  sm_100a only.
//
// One CTA = 128 threads = one UMMA M tile of 128 samples; 4 CTAs per SM (50 KB of shared memory and
// 64 TMEM columns each) hide each other's issue -> commit -> epilogue chain.  Per tile:
//   X [128 x 128] fp16  (cp.async, written in the swizzled K-major layout)
//   D1 [128 x 64] = X W1^T        8 x tcgen05.mma (M128 N64 K16), fp32 in TMEM columns [0,64)
//   H  = relu(D1) -> fp16 -> shared memory (re-uses the first K block of X), one row per thread
//   D2 [128 x 16] = H W2^T        4 x tcgen05.mma (M128 N16 K16), TMEM columns [0,16)
//   sigma = exp(D2[:,0]) (trunc_exp forward, activation.py:10), geo = fp16(D2)
// The mma.sync version of this stage (field_split.cu k_sigma_stage) runs the tensor pipe at 57 % with
// math_pipe_throttle as its first stall (profiles/r01_stages_v2_ncu_full.txt); here one thread issues
// twelve instructions per 128 samples and the other 127 only move data.
#include <algorithm>
#include <cstdlib>

#include "umma.cuh"

#ifndef NVSF_EXP_SKIP_TP
#define NVSF_EXP_SKIP_TP 0
#endif
#ifndef NVSF_TILE_ORDER
#define NVSF_TILE_ORDER 0   // tile -> CTA map of the persistent density kernels: 0 interleaved, 1 contiguous per CTA
#endif

namespace {

using namespace umma;

// packed fp16 MLP image (field_common.cuh: W1 at kSigW1 [64][136], W2 at kSigW2 [16][72]) -> the
// swizzled K-major operand images [W1 K block 0][W1 K block 1][W2], copied verbatim to shared memory
__global__ void k_pack_sigma_tc(const __half* __restrict__ mlp, unsigned char* __restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk each
    if (i < (uint32_t)kHidden * 16) {
        const uint32_t r = i >> 4, c16 = i & 15;
        const uint4 v = *reinterpret_cast<const uint4*>(mlp + kSigW1 + r * kLdK128 + c16 * 8);
        *reinterpret_cast<uint4*>(dst + kOffW1 + (c16 >> 3) * (kHidden * 128) + swz(r, c16 & 7)) = v;
    } else if (i < (uint32_t)kHidden * 16 + kGeo * 8) {
        const uint32_t j = i - kHidden * 16, r = j >> 3, c = j & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(mlp + kSigW2 + r * kLdK64 + c * 8);
        *reinterpret_cast<uint4*>(dst + kOffW2 + swz(r, c)) = v;
    }
}


// flow MLP (flow_field.py:87-103: 32 -> 64 -> 64 -> 6, no bias) operand images, same sizes as the
// sigma images: [W1: 64 rows x 128 B, K = 32 uses the first 64 B][W2: 64 rows x 128 B][W3: 16 rows x
// 128 B, rows 6..15 zero].  `dst` is zero-filled by the caller.
__global__ void k_pack_flow_tc(const __half* __restrict__ mlp, unsigned char* __restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk each
    if (i < (uint32_t)kHidden * 4) {                           // W1 [64][32]
        const uint32_t r = i >> 2, c = i & 3;
        const uint4 v = *reinterpret_cast<const uint4*>(mlp + kFlowW1 + r * kLdK32 + c * 8);
        *reinterpret_cast<uint4*>(dst + swz(r, c)) = v;
    } else if (i < (uint32_t)kHidden * 12) {                   // W2 [64][64]
        const uint32_t j = i - kHidden * 4, r = j >> 3, c = j & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(mlp + kFlowW2 + r * kLdK64 + c * 8);
        *reinterpret_cast<uint4*>(dst + kHidden * 128 + swz(r, c)) = v;
    } else if (i < (uint32_t)kHidden * 12 + 8 * 8) {           // W3 [8][64] (rows 6, 7 are zero)
        const uint32_t j = i - kHidden * 12, r = j >> 3, c = j & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(mlp + kFlowW3 + r * kLdK64 + c * 8);
        *reinterpret_cast<uint4*>(dst + 2 * kHidden * 128 + swz(r, c)) = v;
    }
}

__global__ void __launch_bounds__(kRows, 4)
k_sigma_stage_tc(const unsigned char* __restrict__ wimg, const __half* __restrict__ feat,
                 size_t count, float* __restrict__ sigma_out, __half* __restrict__ geo_out) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // swizzle atoms are 1024-byte aligned
    unsigned char* sm = smem_raw + (base - raw);
    const uint32_t bar = base + kOffBar;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBar + 8);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;

    for (uint32_t i = tid; i < (kW1Bytes + kW2Bytes) / 16; i += kRows)
        reinterpret_cast<uint4*>(sm)[i] = __ldg(reinterpret_cast<const uint4*>(wimg) + i);
    if (tid == 0) mbar_init(bar, 1);
    if (warp == 0) {
        __syncwarp();  // tcgen05.alloc is .sync.aligned: the whole warp, converged
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    fence_async_smem();  // the weight images were written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((warp * 32u) << 16);  // this warp's 32 TMEM lanes
    constexpr uint32_t kIdesc1 = umma_idesc(kRows, kHidden), kIdesc2 = umma_idesc(kRows, kGeo);

    uint32_t phase = 0;
    const size_t n_tiles = (count + kRows - 1) / kRows;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t row0 = tile * kRows;
        // 1. feature rows -> swizzled K-major tile (coalesced 16-byte pieces)
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
            const uint32_t p = k * kRows + tid, r = p >> 4, c16 = p & 15;
            const bool ok = row0 + r < count;
            cp_async16(base + kOffX + (c16 >> 3) * (kRows * 128) + swz(r, c16 & 7),
                       feat + (ok ? row0 + r : 0) * kFeat + c16 * 8, ok);
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
        fence_async_smem();
        __syncthreads();
        // 2. D1 = X W1^T
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (uint32_t k = 0; k < kFeat / 16; ++k)
                umma_f16(tmem, umma_desc(base + kOffX + (k >> 2) * (kRows * 128) + (k & 3) * 32),
                         umma_desc(base + kOffW1 + (k >> 2) * (kHidden * 128) + (k & 3) * 32), kIdesc1,
                         k);
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        // 3. H = relu(D1) as fp16, row `tid`, into the first K block of X (MMA 1 has finished reading it)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t v[16];
            tmem_ld16(tlane + q * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                st_chunk_relu(sm + kOffX, tid, 2 * q + h, v + 8 * h);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // 4. D2 = H W2^T
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (uint32_t k = 0; k < kHidden / 16; ++k)
                umma_f16(tmem, umma_desc(base + kOffX + k * 32), umma_desc(base + kOffW2 + k * 32),
                         kIdesc2, k);
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        // 5. sigma, geo
        {
            uint32_t v[16];
            tmem_ld16(tlane, v);
            tmem_ld_wait();
            const size_t row = row0 + tid;
            if (row < count) {
                sigma_out[row] = expf(round_f16(__uint_as_float(v[0])));
                uint4 a, b;
                a.x = pack_half2(__uint_as_float(v[0]), __uint_as_float(v[1]));
                a.y = pack_half2(__uint_as_float(v[2]), __uint_as_float(v[3]));
                a.z = pack_half2(__uint_as_float(v[4]), __uint_as_float(v[5]));
                a.w = pack_half2(__uint_as_float(v[6]), __uint_as_float(v[7]));
                b.x = pack_half2(__uint_as_float(v[8]), __uint_as_float(v[9]));
                b.y = pack_half2(__uint_as_float(v[10]), __uint_as_float(v[11]));
                b.z = pack_half2(__uint_as_float(v[12]), __uint_as_float(v[13]));
                b.w = pack_half2(__uint_as_float(v[14]), __uint_as_float(v[15]));
                uint4* g = reinterpret_cast<uint4*>(geo_out + row * kGeo);
                g[0] = a;
                g[1] = b;
            }
        }
        tc_fence_before();
        __syncthreads();  // TMEM columns and the X tile are free for the next tile
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(kTmemCols)
                     : "memory");
}

bool g_tc_attr = false;

// ---- gather stage fused with the sigma MLP ---------------------------------------------------------
// The feature rows never leave the SM: a CTA of kFusedWG warpgroups (one CTA per SM) — every warpgroup is an independent 128-sample
// UMMA tile with its own [128 rows][128 B] shared-memory operand tile, mbarrier, named barrier and
// 64 TMEM columns.  A thread gathers the features of its own sample (its row); the sigma-net input
// is consumed in two K halves that accumulate in TMEM:
//   half 0: space planes (32) + collapsed time planes (32)   -> tile -> 4 x tcgen05.mma (D1  = ..)
//   half 1: static hash (32) + dyn rows (24) + ones pad (8)  -> tile -> 4 x tcgen05.mma (D1 += ..)
//   H = relu(D1) fp16 -> tile -> 4 x tcgen05.mma (D2 = H W2^T) -> sigma = exp(D2[:,0]), geo
// The tile is re-used three times per sample block; a warpgroup waits on its own mbarrier before each
// refill (4 of the 32 warps pause for the ~0.5 us of an MMA batch, the other 28 keep gathering).
// Against the staged pair k_encode_stage -> k_sigma_stage this removes the 256 B/sample feature row
// from DRAM (written with one L1 tag per lane, read back by the sigma stage) and one launch.
#ifndef NVSF_FUSED_WG
#define NVSF_FUSED_WG 6   // warpgroups per CTA of the fused gather + sigma stage: 6 x 128 threads at 80 registers
                          // (8 x 64 registers: 11.77 ms, 7: 11.52, 6: 11.47, 4: 14.4 ms per LiDAR frame)
#endif
constexpr int kFusedWG = NVSF_FUSED_WG;
constexpr int kFusedThreads = kFusedWG * kRows;
constexpr uint32_t kFTile = kRows * 128;                         // one K block of 128 rows
constexpr uint32_t kFOffX = kOffW2 + kW2Bytes;
constexpr uint32_t kFOffBar = kFOffX + kFusedWG * kFTile;
constexpr size_t kFusedSmem = kFOffBar + 8 * kFusedWG + 16 + 1024;
static_assert(kFOffX % 1024 == 0, "swizzled tiles start on 1024-byte boundaries");
// the flow stage keeps 8 warpgroups (64 registers per thread): with 6 it measured 5.70 -> 5.97 ms,
// while the gather + sigma stage gains from 6 x 85 registers (11.77 -> 11.50 ms)
#ifndef NVSF_FLOW_WG
#define NVSF_FLOW_WG 8
#endif
constexpr int kFlowWG = NVSF_FLOW_WG;
constexpr int kFlowThreads = kFlowWG * kRows;
constexpr uint32_t kLOffBar = kFOffX + kFlowWG * kFTile;
constexpr size_t kFlowSmem = kLOffBar + 8 * kFlowWG + 16 + 1024;

template <bool H2>
__global__ void __launch_bounds__(kFusedThreads, 1)
k_encode_sigma_tc(const __grid_constant__ nvsf_field_config_t cfg,
                  const __grid_constant__ FieldPtrs P, const float* __restrict__ qpos,
                  const unsigned short* __restrict__ dyn_in, size_t stride, size_t count,
                  float* __restrict__ sigma_out, __half* __restrict__ geo_out,
                  __half* __restrict__ feat_out /* [n,128] kept for the backward pass, or NULL */) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kFOffBar + 8 * kFusedWG);
    const uint32_t tid = threadIdx.x, wg = tid >> 7, t = tid & 127u;

    for (uint32_t i = tid; i < (kW1Bytes + kW2Bytes) / 16; i += kFusedThreads)
        reinterpret_cast<uint4*>(sm)[i] = __ldg(reinterpret_cast<const uint4*>(P.mlp_tc) + i);
    if (tid < (uint32_t)kFusedWG) mbar_init(base + kFOffBar + 8 * tid, 1);
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
    const uint32_t tcol = tmem + wg * 64u;                       // this warpgroup's 64 columns
    const uint32_t tlane = tcol + ((((tid >> 5) & 3u) * 32u) << 16);  // this warp's 32 lanes
    const uint32_t xs = base + kFOffX + wg * kFTile;
    unsigned char* xg = sm + kFOffX + wg * kFTile;
    const uint32_t bar = base + kFOffBar + 8 * wg;
    constexpr uint32_t kIdesc1 = umma_idesc(kRows, kHidden), kIdesc2 = umma_idesc(kRows, kGeo);
    const int qi1 = P.ti->valid[1] ? 1 : 0, qi2 = P.ti->valid[2] ? 2 : 0;

    uint32_t phase = 0;
    const size_t n_tiles = (count + kRows - 1) / kRows;
#if NVSF_TILE_ORDER == 1   // A/B build: every CTA walks ONE contiguous range of tiles (adjacent rays in turn on one SM)
    const size_t per_cta = ((n_tiles + gridDim.x - 1) / gridDim.x + kFusedWG - 1) / kFusedWG * kFusedWG;
    const size_t t_end = min(n_tiles, ((size_t)blockIdx.x + 1) * per_cta);
    for (size_t tile = (size_t)blockIdx.x * per_cta + wg; tile < t_end; tile += kFusedWG) {
#elif NVSF_TILE_ORDER == 2   // A/B build: the warpgroups of a CTA walk ADJACENT rays (6 tiles = 768 samples each) in step
    for (size_t j = 0;; ++j) {
        const size_t v = (j / 6) * ((size_t)gridDim.x * kFusedWG) + (size_t)blockIdx.x * kFusedWG + wg;
        if (v * 6 >= n_tiles) break;
        const size_t tile = v * 6 + j % 6;
        if (tile >= n_tiles) continue;
#else
    for (size_t tile = (size_t)blockIdx.x * kFusedWG + wg; tile < n_tiles;
         tile += (size_t)gridDim.x * kFusedWG) {
#endif
        const size_t li = tile * kRows + t;
        const bool live = li < count;
        const size_t lc = live ? li : count - 1;   // dead rows of the last tile repeat a valid sample
        float qx[3], qy[3], qz[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            qx[q] = __ldg(qpos + (size_t)(3 * q + 0) * stride + lc);
            qy[q] = __ldg(qpos + (size_t)(3 * q + 1) * stride + lc);
            qz[q] = __ldg(qpos + (size_t)(3 * q + 2) * stride + lc);
        }
        // ---- half 0: planes --------------------------------------------------------------------
        if (H2) {
#pragma unroll 1
            for (int s = 0; s < kPlScales; ++s) {
                const uint32_t R = cfg.pl_res[s];
                const __half* b0 = P.pls16 + P.pls_scale[s];
                __half2 v[4];
                plane2d_mul_h2(b0, R, qx[0], qy[0], v, true);
                plane2d_mul_h2(b0 + (size_t)R * R * 8, R, qx[0], qz[0], v, false);
                plane2d_mul_h2(b0 + (size_t)2 * R * R * 8, R, qy[0], qz[0], v, false);
                *reinterpret_cast<uint4*>(xg + swz(t, s)) = *reinterpret_cast<const uint4*>(v);
            }
#pragma unroll 1
            for (int s = 0; s < kPlScales; ++s) {
                const uint32_t R = cfg.pl_res[s];
                __half2 acc[4];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int qq = q == 0 ? 0 : (q == 1 ? qi1 : qi2);
                    const __half* b0 = P.pld16 + (size_t)qq * P.pld_per_q + P.pld_scale[s];
                    __half2 v[4];
                    plane1d_mul_h2(b0, R, qx[q], v, true);
                    plane1d_mul_h2(b0 + (size_t)R * 8, R, qy[q], v, false);
                    plane1d_mul_h2(b0 + (size_t)2 * R * 8, R, qz[q], v, false);
                    const __half2 wq = __float2half2_rn(q == 0 ? 0.5f : 0.25f);
#pragma unroll
                    for (int f = 0; f < 4; ++f) acc[f] = q == 0 ? __hmul2(wq, v[f]) : __hfma2(wq, v[f], acc[f]);
                }
                *reinterpret_cast<uint4*>(xg + swz(t, 4 + s)) = *reinterpret_cast<const uint4*>(acc);
            }
        } else {
#pragma unroll 1
        for (int s = 0; s < kPlScales; ++s) {
            const uint32_t R = cfg.pl_res[s];
            const __half* b0 = P.pls16 + P.pls_scale[s];
            float v[8];
            plane2d_mul_h(b0, R, qx[0], qy[0], v, true);
            plane2d_mul_h(b0 + (size_t)R * R * 8, R, qx[0], qz[0], v, false);
            plane2d_mul_h(b0 + (size_t)2 * R * R * 8, R, qy[0], qz[0], v, false);
            st_chunk(xg, t, s, v);
        }
#pragma unroll 1
        for (int s = 0; s < kPlScales; ++s) {
            const uint32_t R = cfg.pl_res[s];
            float acc8[8];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int qq = q == 0 ? 0 : (q == 1 ? qi1 : qi2);
                const __half* b0 = P.pld16 + (size_t)qq * P.pld_per_q + P.pld_scale[s];
                float v[8];
#if NVSF_EXP_SKIP_TP   // measurement only (wrong results): what the stage would gain if 48 of its 120 plane loads vanished
                if (q == 0) {
#endif
                plane1d_mul_h(b0, R, qx[q], v, true);
                plane1d_mul_h(b0 + (size_t)R * 8, R, qy[q], v, false);
                plane1d_mul_h(b0 + (size_t)2 * R * 8, R, qz[q], v, false);
#if NVSF_EXP_SKIP_TP == 1
                } else {
#pragma unroll
                    for (int f = 0; f < 8; ++f) v[f] = acc8[f] + qx[q];
                }
#elif NVSF_EXP_SKIP_TP == 2   // loads of query 0 (eliminated as common subexpressions), arithmetic of query q
                } else {
                    const float pq[3] = {qx[q], qy[q], qz[q]}, p0[3] = {qx[0], qy[0], qz[0]};
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) {
                        uint32_t x0, x1, y0, y1;
                        float w0, wq_;
                        plane_coord(p0[ax], R, x0, x1, w0);
                        plane_coord(pq[ax], R, y0, y1, wq_);
                        const __half* bq = P.pld16 + P.pld_scale[s] + (size_t)ax * R * 8;   // query 0's table
                        float a[8], b[8];
                        ld8h(bq + (size_t)x0 * 8, a);
                        ld8h(bq + (size_t)x1 * 8, b);
                        wq_ += (float)(y0 & 1u) * 1e-9f;
#pragma unroll
                        for (int f = 0; f < 8; ++f) {
                            const float sv = (1.f - wq_) * a[f] + wq_ * b[f];
                            v[f] = ax == 0 ? sv : v[f] * sv;
                        }
                    }
                }
#endif
                const float wq = q == 0 ? 0.5f : 0.25f;
#pragma unroll
                for (int f = 0; f < 8; ++f) acc8[f] = q == 0 ? wq * v[f] : fmaf(wq, v[f], acc8[f]);
            }
            st_chunk(xg, t, 4 + s, acc8);
        }
        }
        if (feat_out && live) {   // training forward: the thread's own row (K half 0) also goes to global
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<uint4*>(feat_out + li * kFeat + 8 * c) = *reinterpret_cast<const uint4*>(xg + swz(t, c));
        }
        fence_async_smem();
        tc_fence_before();
        wg_barrier(wg);
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
                umma_f16(tcol, umma_desc(xs + k * 32), umma_desc(base + kOffW1 + k * 32), kIdesc1, k);
            umma_commit(bar);
        }
        // ---- half 1: static hash, dyn rows -----------------------------------------------------
#pragma unroll 1
        for (int l = 0; l < kHsLevels; l += 2) {
            float v[8];
            __half2 vh[4];
            if (H2) {
                hash3_f4_h2(P.hs16, lv(cfg.hs[l]), qx[0], qy[0], qz[0], vh);
                hash3_f4_h2(P.hs16, lv(cfg.hs[l + 1]), qx[0], qy[0], qz[0], vh + 2);
            } else {
                hash3_f4(P.hs16, lv(cfg.hs[l]), qx[0], qy[0], qz[0], v);
                hash3_f4(P.hs16, lv(cfg.hs[l + 1]), qx[0], qy[0], qz[0], v + 4);
            }
            if (l == 0) {  // the first-half MMAs must have read the tile before it is refilled
                mbar_wait(bar, phase);
                phase ^= 1u;
            }
            if (H2) *reinterpret_cast<uint4*>(xg + swz(t, l >> 1)) = *reinterpret_cast<const uint4*>(vh);
            else st_chunk(xg, t, l >> 1, v);
        }
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            uint32_t h[8];
#pragma unroll
            for (int l = 0; l < 8; ++l) h[l] = __ldcs(dyn_in + (size_t)(8 * p + l) * stride + lc);
            *reinterpret_cast<uint4*>(xg + swz(t, 4 + p)) =
                make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
        }
        *reinterpret_cast<uint4*>(xg + swz(t, 7)) = make_uint4(kOnesH2, kOnesH2, kOnesH2, kOnesH2);  // tcnn input padding = 1
        if (feat_out && live) {   // K half 1
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<uint4*>(feat_out + li * kFeat + 64 + 8 * c) = *reinterpret_cast<const uint4*>(xg + swz(t, c));
        }
        fence_async_smem();
        tc_fence_before();
        wg_barrier(wg);
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
                umma_f16(tcol, umma_desc(xs + k * 32),
                         umma_desc(base + kOffW1 + kHidden * 128 + k * 32), kIdesc1, 1u);
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        // ---- H = relu(D1) -> fp16, packed IN PLACE into columns [0,32) of this thread's TMEM lane; the
        // second layer takes it as its A operand from tensor memory (D2 = H W2^T into columns [32,48)):
        // no shared-memory round trip of the activations (16 KB written + 16 KB read back per tile)
        hidden_to_tmem(tlane, tlane);
        tc_fence_before();
        wg_barrier(wg);
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
                umma_f16_ts(tcol + 32, tcol + 8 * k, umma_desc(base + kOffW2 + k * 32), kIdesc2, k);
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        {
            uint32_t v[16];
            tmem_ld16(tlane + 32, v);
            tmem_ld_wait();
            if (live) {
                sigma_out[li] = expf(round_f16(__uint_as_float(v[0])));
                uint4 a, b;
                a.x = pack_half2(__uint_as_float(v[0]), __uint_as_float(v[1]));
                a.y = pack_half2(__uint_as_float(v[2]), __uint_as_float(v[3]));
                a.z = pack_half2(__uint_as_float(v[4]), __uint_as_float(v[5]));
                a.w = pack_half2(__uint_as_float(v[6]), __uint_as_float(v[7]));
                b.x = pack_half2(__uint_as_float(v[8]), __uint_as_float(v[9]));
                b.y = pack_half2(__uint_as_float(v[10]), __uint_as_float(v[11]));
                b.z = pack_half2(__uint_as_float(v[12]), __uint_as_float(v[13]));
                b.w = pack_half2(__uint_as_float(v[14]), __uint_as_float(v[15]));
                uint4* g = reinterpret_cast<uint4*>(geo_out + li * kGeo);
                g[0] = a;
                g[1] = b;
            }
        }
        // the next tile's first MMA batch is issued behind a warpgroup barrier that every thread
        // reaches only after this tcgen05.ld has completed (tc_fence_before precedes that barrier)
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u)
                     : "memory");
}


// development knob: NVSF_CARVEOUT=<percent> in the environment sets cudaFuncAttributePreferredSharedMemoryCarveout of
// the persistent density kernels (the rest of the 228 KB per SM is L1): measures how much the gathers owe to L1 capacity
static void apply_carveout(const void* fn) {
    static const char* e = getenv("NVSF_CARVEOUT");
    if (e && *e) cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
}

bool g_fused_attr = false;

// ---- flow stage on tcgen05 -------------------------------------------------------------------------
// FlowField.forward (flow_field.py:116-133) for 128-sample warpgroup tiles: the 16-level collapsed
// flow-grid gather fills K = 32 of the operand tile, then three chained tcgen05.mma batches
// (32 -> 64 relu, 64 -> 64 relu, 64 -> 16 of which 6 are the flow) with the hidden activations going
// TMEM -> registers -> the same tile.  The mma.sync flow stage needs 128 registers for its
// accumulators and runs its gathers at 24 % occupancy (stall long_scoreboard); here the accumulators
// live in TMEM and the gather threads stay at 64 registers, 32 warps per SM.
// Outputs: flow [n,8] f32 (6 used) and the three query positions, planar qpos[9][stride].
// TS = true: the hidden activations stay in tensor memory (packed in place, the next layer's A operand) although a
// warpgroup owns only 64 columns: the second layer runs as two N = 32 halves through columns [32,64), the first
// half's packed result waiting in 16 registers, so that H1 [0,32) stays readable until both are done.
template <bool FROM_RAYS, bool TS>
__global__ void __launch_bounds__(kFlowThreads, 1)
k_flow_tc(const __grid_constant__ nvsf_field_config_t cfg, const __grid_constant__ FieldPtrs P,
          const float* __restrict__ xin, const float* __restrict__ rays_o,
          const float* __restrict__ rays_d, const float* __restrict__ nears,
          const float* __restrict__ fars, const float* __restrict__ noise, uint32_t S, size_t begin,
          size_t count, float* __restrict__ flow_out, float* __restrict__ qpos, size_t stride,
          __half* __restrict__ flowfeat_out /* [n,32] kept for the backward pass, or NULL */) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kLOffBar + 8 * kFlowWG);
    const uint32_t tid = threadIdx.x, wg = tid >> 7, t = tid & 127u;
    const unsigned char* wimg = reinterpret_cast<const unsigned char*>(P.mlp_tc) + (kW1Bytes + kW2Bytes);

    for (uint32_t i = tid; i < (kW1Bytes + kW2Bytes) / 16; i += kFlowThreads)
        reinterpret_cast<uint4*>(sm)[i] = __ldg(reinterpret_cast<const uint4*>(wimg) + i);
    if (tid < (uint32_t)kFlowWG) mbar_init(base + kLOffBar + 8 * tid, 1);
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
    const uint32_t xs = base + kFOffX + wg * kFTile;
    unsigned char* xg = sm + kFOffX + wg * kFTile;
    const uint32_t bar = base + kLOffBar + 8 * wg;
    const uint32_t w1 = base, w2 = base + kHidden * 128, w3 = base + 2 * kHidden * 128;
    constexpr uint32_t kIdesc64 = umma_idesc(kRows, kHidden), kIdesc16 = umma_idesc(kRows, 16);
    const int valid1 = P.ti->valid[1], valid2 = P.ti->valid[2];

    uint32_t phase = 0;
    const size_t n_tiles = (count + kRows - 1) / kRows;
#if NVSF_TILE_ORDER == 1
    const size_t per_cta = ((n_tiles + gridDim.x - 1) / gridDim.x + kFlowWG - 1) / kFlowWG * kFlowWG;
    const size_t t_end = min(n_tiles, ((size_t)blockIdx.x + 1) * per_cta);
    for (size_t tile = (size_t)blockIdx.x * per_cta + wg; tile < t_end; tile += kFlowWG) {
#elif NVSF_TILE_ORDER == 2   // A/B build: the warpgroups of a CTA walk ADJACENT rays (6 tiles = 768 samples each) in step
    for (size_t j = 0;; ++j) {
        const size_t v = (j / 6) * ((size_t)gridDim.x * kFlowWG) + (size_t)blockIdx.x * kFlowWG + wg;
        if (v * 6 >= n_tiles) break;
        const size_t tile = v * 6 + j % 6;
        if (tile >= n_tiles) continue;
#else
    for (size_t tile = (size_t)blockIdx.x * kFlowWG + wg; tile < n_tiles;
         tile += (size_t)gridDim.x * kFlowWG) {
#endif
        const size_t li = tile * kRows + t;
        const bool live = li < count;
        float x, y, z;
        sample_position<FROM_RAYS>(cfg, begin + (live ? li : count - 1), xin, rays_o, rays_d, nears, fars,
                                   noise, S, x, y, z);
        // flow-grid features: 4 levels -> 8 halves -> one 16-byte chunk
#pragma unroll 1
        for (int c = 0; c < kFlLevels / 4; ++c) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = hash3_h2(P.flow16, lv(cfg.fl[4 * c + j]), x, y, z);
                v[2 * j] = f.x;
                v[2 * j + 1] = f.y;
            }
            st_chunk(xg, t, c, v);
            if (flowfeat_out && live)
                *reinterpret_cast<uint4*>(flowfeat_out + li * kFlowIn + 8 * c) =
                    *reinterpret_cast<const uint4*>(xg + swz(t, c));
        }
        fence_async_smem();
        tc_fence_before();
        wg_barrier(wg);
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (uint32_t k = 0; k < kFlowIn / 16; ++k)
                umma_f16(tcol, umma_desc(xs + k * 32), umma_desc(w1 + k * 32), kIdesc64, k);
            umma_commit(bar);
        }
        if (TS) {
            constexpr uint32_t kIdesc32 = umma_idesc(kRows, 32);
            mbar_wait(bar, phase); phase ^= 1u;
            tc_fence_after();
            hidden_to_tmem(tlane, tlane);                 // H1 packed in place [0,32)
            tc_fence_before();
            wg_barrier(wg);
            uint32_t h2[2][16];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (t == 0) {   // D2 half [32,64) = H1 W2[32 half .. 32 half + 32)^T, A from tensor memory
                    tc_fence_after();
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k)
                        umma_f16_ts(tcol + 32, tcol + 8 * k, umma_desc(w2 + half * (32 * 128) + k * 32), kIdesc32, k);
                    umma_commit(bar);
                }
                mbar_wait(bar, phase); phase ^= 1u;
                tc_fence_after();
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    uint32_t v[16];
                    tmem_ld16(tlane + 32 + q * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        h2[half][8 * q + i] = pack_half2_relu(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                }
                tc_fence_before();
                wg_barrier(wg);   // every thread has read the half before the next MMA batch lands in [32,64)
            }
            {   // H2 = [half 0 | half 1] packed over H1 [0,32): both halves' MMAs have read it
                uint32_t o[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = h2[q >> 1][(q & 1) * 8 + i];
                    tmem_st8(tlane + q * 8, o);
                }
                tmem_st_wait();
            }
            tc_fence_before();
            wg_barrier(wg);
            if (t == 0) {   // D3 [32,48) = H2 W3^T
                tc_fence_after();
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                    umma_f16_ts(tcol + 32, tcol + 8 * k, umma_desc(w3 + k * 32), kIdesc16, k);
                umma_commit(bar);
            }
        } else {
        // two hidden layers: D -> relu -> fp16 tile -> next MMA batch
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
            mbar_wait(bar, phase);
            phase ^= 1u;
            tc_fence_after();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t v[16];
                tmem_ld16(tlane + q * 16, v);
                tmem_ld_wait();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    st_chunk_relu(xg, t, 2 * q + h, v + 8 * h);
                }
            }
            fence_async_smem();
            tc_fence_before();
            wg_barrier(wg);
            if (t == 0) {
                tc_fence_after();
                const uint32_t wb = layer == 0 ? w2 : w3;
                const uint32_t idesc = layer == 0 ? kIdesc64 : kIdesc16;
#pragma unroll
                for (uint32_t k = 0; k < kHidden / 16; ++k)
                    umma_f16(tcol, umma_desc(xs + k * 32), umma_desc(wb + k * 32), idesc, k);
                umma_commit(bar);
            }
        }
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        {
            uint32_t v[16];
            tmem_ld16(tlane + (TS ? 32u : 0u), v);
            tmem_ld_wait();
            if (live) {
                const float4 f0 = make_float4(round_f16(__uint_as_float(v[0])), round_f16(__uint_as_float(v[1])),
                                              round_f16(__uint_as_float(v[2])), round_f16(__uint_as_float(v[3])));
                const float4 f1 = make_float4(round_f16(__uint_as_float(v[4])), round_f16(__uint_as_float(v[5])),
                                              round_f16(__uint_as_float(v[6])), round_f16(__uint_as_float(v[7])));
                if (flow_out) {   // NULL when only the query positions are consumed (fused render path)
                    float4* d = reinterpret_cast<float4*>(flow_out + li * 8);
                    d[0] = f0;
                    d[1] = f1;
                }
                float* q = qpos + li;
                q[0] = x; q[stride] = y; q[2 * stride] = z;
                q[3 * stride] = valid1 ? x + f0.x : x;
                q[4 * stride] = valid1 ? y + f0.y : y;
                q[5 * stride] = valid1 ? z + f0.z : z;
                q[6 * stride] = valid2 ? x + f0.w : x;
                q[7 * stride] = valid2 ? y + f1.x : y;
                q[8 * stride] = valid2 ? z + f1.y : z;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u)
                     : "memory");
}

bool g_flow_attr = false;

}  // namespace

int g_flow_ts = 0;   // option "flow_ts": hidden activations of the flow MLP kept in tensor memory (k_flow_tc<., true>);
                     // measured neutral (5.73 vs 5.70 ms per LiDAR frame: the stage is bound by its 128 global gathers
                     // per sample — 55 % of the L1 data pipe against 9 % + 11 % for the operand tiles), off by default

size_t nvsf_sigma_tc_image_bytes() { return 2 * (size_t)(kW1Bytes + kW2Bytes); }

// [sigma-net images][flow-MLP images], 18 KB each
void nvsf_pack_sigma_tc(const __half* mlp, void* dst, cudaStream_t stream) {
    unsigned char* d = reinterpret_cast<unsigned char*>(dst);
    cudaMemsetAsync(d, 0, nvsf_sigma_tc_image_bytes(), stream);
    k_pack_sigma_tc<<<nvsf_div_up(kHidden * 16 + kGeo * 8, 128), 128, 0, stream>>>(mlp, d);
    k_pack_flow_tc<<<nvsf_div_up(kHidden * 12 + 64, 128), 128, 0, stream>>>(mlp, d + kW1Bytes + kW2Bytes);
}

int nvsf_launch_sigma_tc(const void* wimg, const __half* feat, size_t count, float* sigma,
                         __half* geo, int sms, cudaStream_t stream) {
    if (!g_tc_attr) {
        cudaError_t e = cudaFuncSetAttribute(k_sigma_stage_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kTcSmem);
        if (e != cudaSuccess) return (int)e;
        g_tc_attr = true;
    }
    const size_t tiles = (count + kRows - 1) / kRows;
    const int grid = (int)std::min<size_t>(tiles, (size_t)sms * 4);
    k_sigma_stage_tc<<<grid, kRows, kTcSmem, stream>>>(reinterpret_cast<const unsigned char*>(wimg), feat,
                                                        count, sigma, geo);
    return NVSF_OK;
}

int nvsf_launch_encode_sigma_tc(const nvsf_field_config_t* cfg, const FieldPtrs& P, const float* qpos,
                                const void* dyn_in, size_t stride, size_t count, float* sigma,
                                __half* geo, int sms, cudaStream_t stream, int half_math, __half* feat_out) {
    if (!g_fused_attr) {
        cudaError_t e = cudaFuncSetAttribute(k_encode_sigma_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kFusedSmem);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(k_encode_sigma_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)kFusedSmem);
        if (e != cudaSuccess) return (int)e;
        apply_carveout((const void*)k_encode_sigma_tc<false>);
        apply_carveout((const void*)k_encode_sigma_tc<true>);
        g_fused_attr = true;
    }
    const size_t tiles = (count + kRows - 1) / kRows;
    const int grid = (int)std::min<size_t>((tiles + kFusedWG - 1) / kFusedWG, (size_t)sms);
    if (half_math)
        k_encode_sigma_tc<true><<<grid, kFusedThreads, kFusedSmem, stream>>>(
            *cfg, P, qpos, reinterpret_cast<const unsigned short*>(dyn_in), stride, count, sigma, geo, feat_out);
    else
        k_encode_sigma_tc<false><<<grid, kFusedThreads, kFusedSmem, stream>>>(
            *cfg, P, qpos, reinterpret_cast<const unsigned short*>(dyn_in), stride, count, sigma, geo, feat_out);
    return NVSF_OK;
}

int nvsf_launch_flow_tc(const nvsf_field_config_t* cfg, const FieldPtrs& P, const float* x,
                        const float* rays_o, const float* rays_d, const float* nears, const float* fars,
                        const float* noise, uint32_t S, size_t begin, size_t count, float* flow_out,
                        float* qpos, size_t stride, __half* flowfeat, int sms, cudaStream_t stream) {
    if (!g_flow_attr) {
        cudaError_t e = cudaFuncSetAttribute(k_flow_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kFlowSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k_flow_tc<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFlowSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k_flow_tc<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFlowSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k_flow_tc<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFlowSmem);
        if (e != cudaSuccess) return (int)e;
        apply_carveout((const void*)k_flow_tc<true, false>);
        apply_carveout((const void*)k_flow_tc<false, false>);
        g_flow_attr = true;
    }
    const size_t tiles = (count + kRows - 1) / kRows;
    const int grid = (int)std::min<size_t>((tiles + kFlowWG - 1) / kFlowWG, (size_t)sms);
#define NVSF_FLOW_TC(FR, TSV, XP, ...)                                                                      \
    k_flow_tc<FR, TSV><<<grid, kFlowThreads, kFlowSmem, stream>>>(*cfg, P, XP, __VA_ARGS__, begin, count, flow_out, \
                                                                  qpos, stride, flowfeat)
    if (x) {
        if (g_flow_ts) NVSF_FLOW_TC(false, true, x, nullptr, nullptr, nullptr, nullptr, nullptr, 1);
        else NVSF_FLOW_TC(false, false, x, nullptr, nullptr, nullptr, nullptr, nullptr, 1);
    } else {
        if (g_flow_ts) NVSF_FLOW_TC(true, true, nullptr, rays_o, rays_d, nears, fars, noise, S);
        else NVSF_FLOW_TC(true, false, nullptr, rays_o, rays_d, nears, fars, noise, S);
    }
#undef NVSF_FLOW_TC
    return NVSF_OK;
}
