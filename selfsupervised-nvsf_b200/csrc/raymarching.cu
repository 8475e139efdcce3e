// nvsf_b200 — ray-marching operators for sm_100a (Part 1 of include/nvsf_b200.h).
//
// Replaces the reference's `_raymarching` extension
// (reference nvsf/nerf/raymarching/src/raymarching.cu).  Designed for B200, not
// translated: rays are staged through shared memory with 16-byte loads, sample
// offsets come from block/warp prefix sums instead of global atomics (so the
// output order is deterministic), compositing streams each ray's samples through
// per-warp shared-memory tiles so that global traffic is coalesced while the
// per-ray arithmetic keeps the reference's exact operation order.
//
// Bit-exactness: every floating-point operation that decides a sample position
// or count is written with explicit rounding intrinsics (__fmaf_rn, __fmul_rn,
// ...) in the order the reference kernel executes on sm_100a (read from its
// SASS), so results do not depend on this file's compiler flags.
#include <float.h>

#include "common.cuh"

int nvsf_march_mode();  // field.cu: option "march_mode" (1 = warp-cooperative emission, default)
int nvsf_composite_mode();      // field.cu: option "composite_mode"
int nvsf_composite_bwd_mode();  // field.cu: option "composite_bwd_mode"

namespace {

constexpr int kRayBlock = 128;   // rays per CTA for per-ray kernels
constexpr int kTileBlock = 256;  // rays per CTA for the light utility kernels

// ---------------------------------------------------------------------------
// near_far_from_aabb  (reference kernel raymarching.cu:105-157)
// ---------------------------------------------------------------------------
// slab test of one ray in the reference's operation order
__device__ __forceinline__ void near_far_one(float ox, float oy, float oz, float dx, float dy, float dz,
                                             const float (&a)[6], float min_near, float& near_out, float& far_out) {
    const float rdx = __frcp_rn(dx), rdy = __frcp_rn(dy), rdz = __frcp_rn(dz);
    const float a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3], a4 = a[4], a5 = a[5];

    float near = __fmul_rn(__fsub_rn(a0, ox), rdx);
    float far = __fmul_rn(__fsub_rn(a3, ox), rdx);
    if (near > far) { float c = near; near = far; far = c; }

    float near_y = __fmul_rn(__fsub_rn(a1, oy), rdy);
    float far_y = __fmul_rn(__fsub_rn(a4, oy), rdy);
    if (near_y > far_y) { float c = near_y; near_y = far_y; far_y = c; }

    bool miss = (near > far_y) || (near_y > far);
    if (!miss) {
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = __fmul_rn(__fsub_rn(a2, oz), rdz);
        float far_z = __fmul_rn(__fsub_rn(a5, oz), rdz);
        if (near_z > far_z) { float c = near_z; near_z = far_z; far_z = c; }
        miss = (near > far_z) || (near_z > far);
        if (!miss) {
            if (near_z > near) near = near_z;
            if (far_z < far) far = far_z;
            if (near < min_near) near = min_near;
        }
    }
    if (miss) near = far = FLT_MAX;
    near_out = near;
    far_out = far;
}

// Four consecutive rays per thread: their 12 + 12 input floats are three 16-byte loads each and the results
// one 16-byte store each, all coalesced, no shared memory and no barrier (the kernel moves 17 MB for a camera
// frame: 3 us of HBM time, so every dependent step on its critical path shows).  16-byte aligned base pointers;
// the last N % 4 rays and unaligned views take the scalar kernel.
__global__ void __launch_bounds__(128)
k_near_far_from_aabb4(const float4* __restrict__ rays_o, const float4* __restrict__ rays_d,
                      const float* __restrict__ aabb, uint32_t n4, float min_near,
                      float4* __restrict__ nears, float4* __restrict__ fars) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 o0 = __ldg(rays_o + 3 * (size_t)i), o1 = __ldg(rays_o + 3 * (size_t)i + 1), o2 = __ldg(rays_o + 3 * (size_t)i + 2);
    const float4 d0 = __ldg(rays_d + 3 * (size_t)i), d1 = __ldg(rays_d + 3 * (size_t)i + 1), d2 = __ldg(rays_d + 3 * (size_t)i + 2);
    float a[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] = __ldg(aabb + k);
    float4 nr, fr;
    near_far_one(o0.x, o0.y, o0.z, d0.x, d0.y, d0.z, a, min_near, nr.x, fr.x);
    near_far_one(o0.w, o1.x, o1.y, d0.w, d1.x, d1.y, a, min_near, nr.y, fr.y);
    near_far_one(o1.z, o1.w, o2.x, d1.z, d1.w, d2.x, a, min_near, nr.z, fr.z);
    near_far_one(o2.y, o2.z, o2.w, d2.y, d2.z, d2.w, a, min_near, nr.w, fr.w);
    nears[i] = nr;
    fars[i] = fr;
}

__global__ void __launch_bounds__(128)
k_near_far_from_aabb(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                     const float* __restrict__ aabb, uint32_t first, uint32_t N, float min_near,
                     float* __restrict__ nears, float* __restrict__ fars) {
    const uint32_t n = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float a[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] = __ldg(aabb + k);
    float near, far;
    near_far_one(__ldg(rays_o + 3 * (size_t)n), __ldg(rays_o + 3 * (size_t)n + 1), __ldg(rays_o + 3 * (size_t)n + 2),
                 __ldg(rays_d + 3 * (size_t)n), __ldg(rays_d + 3 * (size_t)n + 1), __ldg(rays_d + 3 * (size_t)n + 2), a,
                 min_near, near, far);
    nears[n] = near;
    fars[n] = far;
}

// ---------------------------------------------------------------------------
// sph_from_ray  (reference kernel raymarching.cu:183-217)
// ---------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_sph_from_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float radius,
               uint32_t N, float* __restrict__ coords) {
    __shared__ __align__(16) float s_o[BLOCK * 3];
    __shared__ __align__(16) float s_d[BLOCK * 3];
    const uint32_t base = blockIdx.x * BLOCK;
    const uint32_t cnt = min((uint32_t)BLOCK, N - base);
    block_load_floats<BLOCK>(rays_o + (size_t)base * 3, s_o, cnt * 3);
    block_load_floats<BLOCK>(rays_d + (size_t)base * 3, s_d, cnt * 3);
    __syncthreads();
    const uint32_t tid = threadIdx.x;
    if (tid >= cnt) return;
    const uint32_t n = base + tid;
    const float ox = s_o[3 * tid], oy = s_o[3 * tid + 1], oz = s_o[3 * tid + 2];
    const float dx = s_d[3 * tid], dy = s_d[3 * tid + 1], dz = s_d[3 * tid + 2];

    // || o + t d || = radius, larger root (operation order of the reference on sm_100a)
    const float A = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    const float B = __fmaf_rn(oz, dz, __fmaf_rn(ox, dx, __fmul_rn(oy, dy)));
    const float C = __fmaf_rn(-radius, radius,
                              __fmaf_rn(oz, oz, __fmaf_rn(ox, ox, __fmul_rn(oy, oy))));
    const float disc = __fmaf_rn(B, B, -__fmul_rn(A, C));
    const float t = __fdiv_rn(__fadd_rn(-B, __fsqrt_rn(disc)), A);

    const float x = __fmaf_rn(dx, t, ox), y = __fmaf_rn(dy, t, oy), z = __fmaf_rn(dz, t, oz);
    const float theta = atan2f(__fsqrt_rn(__fmaf_rn(x, x, __fmul_rn(z, z))), y);
    const float phi = atan2f(z, x);
    const float kRPI = 0.3183098861837907f;
    float2 out;
    out.x = __fmaf_rn(__fmul_rn(2.0f, theta), kRPI, -1.0f);
    out.y = __fmul_rn(phi, kRPI);
    reinterpret_cast<float2*>(coords)[n] = out;
}

// ---------------------------------------------------------------------------
// morton3D / morton3D_invert  (reference raymarching.cu:71-95, 237-280)
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {
    // 10 input bits -> every third bit.  Equal to the reference's multiply/mask
    // form for v < 1024 (the only range a 32-bit 3-D Morton code can hold).
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
// The reference's expansion, kept bit-identical for ANY 32-bit input: the public
// morton3D operator must agree with it even for out-of-range coordinates.
__host__ __device__ __forceinline__ uint32_t spread3_wide(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t compact3(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_morton3D(const int32_t* __restrict__ coords, uint32_t N, int32_t* __restrict__ indices) {
    __shared__ __align__(16) float s_c[BLOCK * 3];
    const uint32_t base = blockIdx.x * BLOCK;
    const uint32_t cnt = min((uint32_t)BLOCK, N - base);
    block_load_floats<BLOCK>(reinterpret_cast<const float*>(coords) + (size_t)base * 3, s_c,
                             cnt * 3);
    __syncthreads();
    const uint32_t tid = threadIdx.x;
    if (tid >= cnt) return;
    const uint32_t x = __float_as_uint(s_c[3 * tid]);
    const uint32_t y = __float_as_uint(s_c[3 * tid + 1]);
    const uint32_t z = __float_as_uint(s_c[3 * tid + 2]);
    indices[base + tid] =
        (int32_t)(spread3_wide(x) | (spread3_wide(y) << 1) | (spread3_wide(z) << 2));
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_morton3D_invert(const int32_t* __restrict__ indices, uint32_t N, int32_t* __restrict__ coords) {
    __shared__ __align__(16) int32_t s_c[BLOCK * 3];
    const uint32_t base = blockIdx.x * BLOCK;
    const uint32_t cnt = min((uint32_t)BLOCK, N - base);
    const uint32_t tid = threadIdx.x;
    if (tid < cnt) {
        // arithmetic shift of a signed int, as in the reference (ind >> k on int)
        const int32_t ind = __ldg(indices + base + tid);
        s_c[3 * tid] = (int32_t)compact3((uint32_t)(ind >> 0));
        s_c[3 * tid + 1] = (int32_t)compact3((uint32_t)(ind >> 1));
        s_c[3 * tid + 2] = (int32_t)compact3((uint32_t)(ind >> 2));
    }
    __syncthreads();
    int32_t* dst = coords + (size_t)base * 3;
    const uint32_t nw = cnt * 3;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        const uint32_t nv = nw >> 2;
        for (uint32_t i = tid; i < nv; i += BLOCK)
            reinterpret_cast<int4*>(dst)[i] = reinterpret_cast<const int4*>(s_c)[i];
        for (uint32_t i = (nv << 2) + tid; i < nw; i += BLOCK) dst[i] = s_c[i];
    } else {
        for (uint32_t i = tid; i < nw; i += BLOCK) dst[i] = s_c[i];
    }
}

// ---------------------------------------------------------------------------
// packbits  (reference kernel raymarching.cu:287-306)
// ---------------------------------------------------------------------------
// One warp turns 1024 consecutive densities (8 coalesced float4 loads per lane,
// all in flight together) into 32 words of bitfield written with one 128-byte
// store.  Nibbles are merged across lanes with three shuffles.
constexpr int kPackIters = 8;
__global__ void __launch_bounds__(256)
k_packbits_vec(const float4* __restrict__ grid4, uint32_t n_super, float thresh,
               uint32_t* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= n_super) return;
    const float4* src = grid4 + (size_t)warp * (kPackIters * 32) + lane;
    float4 v[kPackIters];
#pragma unroll
    for (int it = 0; it < kPackIters; ++it) v[it] = __ldcs(src + it * 32);
    uint32_t mine = 0;
#pragma unroll
    for (int it = 0; it < kPackIters; ++it) {
        uint32_t x = (v[it].x > thresh ? 1u : 0u) | (v[it].y > thresh ? 2u : 0u) |
                     (v[it].z > thresh ? 4u : 0u) | (v[it].w > thresh ? 8u : 0u);
        x |= __shfl_down_sync(0xffffffffu, x, 1) << 4;
        x |= __shfl_down_sync(0xffffffffu, x, 2) << 8;
        x |= __shfl_down_sync(0xffffffffu, x, 4) << 16;
        const uint32_t w = __shfl_sync(0xffffffffu, x, (lane & 3) * 8);
        if ((lane >> 2) == (uint32_t)it) mine = w;
    }
    out[(size_t)warp * 32 + lane] = mine;
}

__global__ void __launch_bounds__(256)
k_packbits_scalar(const float* __restrict__ grid, uint32_t first, uint32_t N, float thresh,
                  uint8_t* __restrict__ bitfield) {
    const uint32_t n = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* g = grid + (size_t)n * 8;
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) bits |= (__ldg(g + i) > thresh) ? (1u << i) : 0u;
    bitfield[n] = (uint8_t)bits;
}

// ---------------------------------------------------------------------------
// Occupancy-grid DDA shared by the train and inference marchers
// (reference raymarching.cu:359-439 / 463-533 / 840-927).
// ---------------------------------------------------------------------------
struct MarchParams {
    float bound, dt_gamma, dt_min, dt_max;
    float Hf, rH, H3f, Hm1f, halfH;
    int cm1;
    uint32_t H;
    int h_pow2;
};

__device__ __forceinline__ MarchParams make_march_params(float bound, float dt_gamma,
                                                         uint32_t max_steps, uint32_t C,
                                                         uint32_t H) {
    MarchParams p;
    const float k2sqrt3 = __uint_as_float(0x405DB3D7u);  // 2*SQRT3() as the reference folds it
    p.bound = bound;
    p.dt_gamma = dt_gamma;
    p.Hf = (float)H;
    p.dt_min = __fdiv_rn(k2sqrt3, (float)max_steps);
    p.dt_max = __fdiv_rn(__fmul_rn((float)(int)(1u << (C - 1)), k2sqrt3), p.Hf);
    p.rH = __frcp_rn(p.Hf);
    p.H3f = (float)(H * H * H);
    p.Hm1f = (float)(H - 1);
    p.halfH = 0.5f * p.Hf;
    p.cm1 = (int)C - 1;
    p.H = H;
    p.h_pow2 = (H & (H - 1)) == 0;
    return p;
}

struct RayGeom {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, sx, sy, sz;
};

__device__ __forceinline__ RayGeom make_ray(float ox, float oy, float oz, float dx, float dy,
                                            float dz) {
    RayGeom r;
    r.ox = ox; r.oy = oy; r.oz = oz;
    r.dx = dx; r.dy = dy; r.dz = dz;
    r.rdx = __frcp_rn(dx); r.rdy = __frcp_rn(dy); r.rdz = __frcp_rn(dz);
    r.sx = copysignf(1.0f, dx); r.sy = copysignf(1.0f, dy); r.sz = copysignf(1.0f, dz);
    return r;
}

__device__ __forceinline__ float clamp_dt(const MarchParams& p, float t) {
    return fminf(p.dt_max, fmaxf(p.dt_min, __fmul_rn(p.dt_gamma, t)));
}

// exponent of frexpf(v) for v >= 0, clamped to [0, cm1]  (mip_from_pos / mip_from_dt)
__device__ __forceinline__ int mip_level(float v, int cm1) {
    const int f = (__float_as_int(v) >> 23) & 0xff;
    const int e = (f == 255) ? 0 : f - 126;  // zero/denormals give e <= 0 -> clamped to 0
    return min(max(e, 0), cm1);
}

__device__ __forceinline__ int grid_cell(const MarchParams& p, float f) {
    // reference: (int) clamp(0.5 * (x * mip_rbound + 1) * H, 0, H-1) with the product in double
    float v;
    if (p.h_pow2) v = __fmul_rn(f, p.halfH);  // exact, identical to the double path
    else v = __double2float_rn(__dmul_rn(__dmul_rn((double)f, 0.5), (double)p.H));
    return (int)fminf(p.Hm1f, fmaxf(v, 0.0f));
}

struct Probe {
    float x, y, z, dt;
    bool occ;
};

// Evaluates the sample at ray parameter t.  When the cell is empty, *t_skip is
// the ray parameter of the next voxel boundary (`tt` in the reference).
__device__ __forceinline__ Probe probe_at(const MarchParams& p, const RayGeom& r,
                                          const uint8_t* __restrict__ grid, float t,
                                          float* t_skip) {
    Probe q;
    q.x = fminf(p.bound, fmaxf(-p.bound, __fmaf_rn(r.dx, t, r.ox)));
    q.y = fminf(p.bound, fmaxf(-p.bound, __fmaf_rn(r.dy, t, r.oy)));
    q.z = fminf(p.bound, fmaxf(-p.bound, __fmaf_rn(r.dz, t, r.oz)));
    q.dt = clamp_dt(p, t);

    const float mx = fmaxf(fabsf(q.x), fmaxf(fabsf(q.y), fabsf(q.z)));
    const int level = max(mip_level(mx, p.cm1),
                          mip_level(__fmul_rn(__fmul_rn(q.dt, p.Hf), 0.5f), p.cm1));
    const float mip_bound = fminf(__int_as_float((127 + level) << 23), p.bound);
    const float mip_rbound = __frcp_rn(mip_bound);

    const int nx = grid_cell(p, __fmaf_rn(q.x, mip_rbound, 1.0f));
    const int ny = grid_cell(p, __fmaf_rn(q.y, mip_rbound, 1.0f));
    const int nz = grid_cell(p, __fmaf_rn(q.z, mip_rbound, 1.0f));

    const uint32_t morton = spread3((uint32_t)nx) | (spread3((uint32_t)ny) << 1) |
                            (spread3((uint32_t)nz) << 2);
    // reference computes level*H3 + morton in float (raymarching.cu:363,407)
    const uint32_t index = (uint32_t)__fmaf_rn(p.H3f, (float)level, (float)morton);
    q.occ = (__ldg(grid + (index >> 3)) >> (index & 7u)) & 1u;

    if (!q.occ) {
        const float ax = __fmaf_rn(r.sx, 0.5f, __fadd_rn((float)nx, 0.5f));
        const float ay = __fmaf_rn(r.sy, 0.5f, __fadd_rn((float)ny, 0.5f));
        const float az = __fmaf_rn(r.sz, 0.5f, __fadd_rn((float)nz, 0.5f));
        const float tx = __fmul_rn(
            __fmaf_rn(mip_bound, __fmaf_rn(__fmul_rn(ax, p.rH), 2.0f, -1.0f), -q.x), r.rdx);
        const float ty = __fmul_rn(
            __fmaf_rn(mip_bound, __fmaf_rn(__fmul_rn(ay, p.rH), 2.0f, -1.0f), -q.y), r.rdy);
        const float tz = __fmul_rn(
            __fmaf_rn(mip_bound, __fmaf_rn(__fmul_rn(az, p.rH), 2.0f, -1.0f), -q.z), r.rdz);
        *t_skip = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
    }
    return q;
}

__device__ __forceinline__ float skip_to(const MarchParams& p, float t, float tt) {
    do {
        t = __fadd_rn(t, clamp_dt(p, t));
    } while (t < tt);
    return t;
}

__device__ __forceinline__ float first_t(const MarchParams& p, float near, float noise) {
    return __fmaf_rn(noise, clamp_dt(p, near), near);
}

// ---------------------------------------------------------------------------
// march_rays_train, phase 1: per-ray sample counts + per-CTA sums
// ---------------------------------------------------------------------------
// (t, dt) of the first kStash samples of every ray, [ray/32][kStash][ray%32] so that the 32
// lanes of a warp touch one 256-byte row per sample.  Sparse rays (<= kStash samples, the common
// case behind an occupancy grid) are then emitted by phase 2 without a second DDA walk, and
// longer rays resume after their kStash-th sample instead of re-crossing the empty space in
// front of the surface.
constexpr int kStash = 16;
__host__ __device__ __forceinline__ size_t stash_index(uint32_t n, uint32_t s) {
    return ((size_t)(n >> 5) * kStash + s) * 32 + (n & 31u);
}

// workspace layout (uint32 words): [0]=base point counter, [1]=base ray counter,
// [2..3] pad, [4 .. 4+nblk) CTA sums -> exclusive CTA offsets, then N counts.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_march_train_count(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                    const uint8_t* __restrict__ grid, float bound, float dt_gamma,
                    uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                    const float* __restrict__ nears, const float* __restrict__ fars,
                    const float* __restrict__ noises, const int32_t* __restrict__ counter,
                    uint32_t* __restrict__ ws, uint32_t nblk, float2* __restrict__ stash) {
    __shared__ __align__(16) float s_o[BLOCK * 3];
    __shared__ __align__(16) float s_d[BLOCK * 3];
    __shared__ uint32_t s_warp[BLOCK / 32];
    const uint32_t base = blockIdx.x * BLOCK;
    const uint32_t cnt = min((uint32_t)BLOCK, N - base);
    block_load_floats<BLOCK>(rays_o + (size_t)base * 3, s_o, cnt * 3);
    block_load_floats<BLOCK>(rays_d + (size_t)base * 3, s_d, cnt * 3);
    if (blockIdx.x == 0 && threadIdx.x < 2) ws[threadIdx.x] = (uint32_t)counter[threadIdx.x];
    __syncthreads();

    const uint32_t tid = threadIdx.x;
    uint32_t num_steps = 0;
    if (tid < cnt) {
        const uint32_t n = base + tid;
        const MarchParams p = make_march_params(bound, dt_gamma, max_steps, C, H);
        const RayGeom r = make_ray(s_o[3 * tid], s_o[3 * tid + 1], s_o[3 * tid + 2],
                                   s_d[3 * tid], s_d[3 * tid + 1], s_d[3 * tid + 2]);
        const float far = __ldg(fars + n);
        float t = first_t(p, __ldg(nears + n), __ldg(noises + n));
        while (t < far && num_steps < max_steps) {
            float tt;
            const Probe q = probe_at(p, r, grid, t, &tt);
            if (q.occ) {
                // the first kStash samples are kept so that phase 2 need not walk to them again
                if (num_steps < (uint32_t)kStash) stash[stash_index(n, num_steps)] = make_float2(t, q.dt);
                ++num_steps;
                t = __fadd_rn(t, q.dt);
            } else {
                t = skip_to(p, t, tt);
            }
        }
        ws[4 + nblk + n] = num_steps;
    }
    const uint32_t wsum = warp_reduce_sum(num_steps);
    if ((tid & 31) == 0) s_warp[tid >> 5] = wsum;
    __syncthreads();
    if (tid == 0) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < BLOCK / 32; ++w) s += s_warp[w];
        ws[4 + blockIdx.x] = s;
    }
}

// phase 1b: exclusive scan of the CTA sums (single CTA), counter update.
__global__ void __launch_bounds__(1024)
k_march_train_scan(uint32_t* __restrict__ ws, uint32_t nblk, uint32_t N,
                   int32_t* __restrict__ counter) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    uint32_t* sums = ws + 4;
    for (uint32_t start = 0; start < nblk; start += 1024) {
        const uint32_t i = start + tid;
        const uint32_t v = i < nblk ? sums[i] : 0u;
        const uint32_t inc = warp_inclusive_scan(v, lane);
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = s_warp[lane];
            const uint32_t winc = warp_inclusive_scan(w, lane);
            s_warp[lane] = winc - w;  // exclusive warp offsets
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t excl = carry + s_warp[warp] + inc - v;
        if (i < nblk) sums[i] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        counter[0] = (int32_t)(ws[0] + s_carry);
        counter[1] = (int32_t)(ws[1] + N);
    }
}

// phase 1c: intra-CTA prefix sums -> rays[N,3] = (ray id, offset, count).
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_march_train_rows(const uint32_t* __restrict__ ws, uint32_t nblk, uint32_t N,
                   int32_t* __restrict__ rays) {
    __shared__ uint32_t s_warp[BLOCK / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = blockIdx.x * BLOCK + tid;
    const uint32_t c = n < N ? ws[4 + nblk + n] : 0u;
    const uint32_t inc = warp_inclusive_scan(c, lane);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w)
        if (w < (int)warp) woff += s_warp[w];
    if (n < N) {
        const uint32_t row = ws[1] + n;
        if (row < N) {  // the reference would write out of bounds here; rows are dropped instead
            const uint32_t off = ws[0] + ws[4 + blockIdx.x] + woff + inc - c;
            rays[(size_t)row * 3 + 0] = (int32_t)n;
            rays[(size_t)row * 3 + 1] = (int32_t)off;
            rays[(size_t)row * 3 + 2] = (int32_t)c;
        }
    }
}

// ---------------------------------------------------------------------------
// march_rays_train, phase 2: emit samples (reference raymarching.cu:463-533)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void zero_rows(float* xyzs, float* dirs, float* deltas, uint32_t r0,
                                          uint32_t r1, uint32_t tid, uint32_t nthreads) {
    for (size_t i = (size_t)r0 * 3 + tid; i < (size_t)r1 * 3; i += nthreads) {
        xyzs[i] = 0.0f;
        dirs[i] = 0.0f;
    }
    for (size_t i = (size_t)r0 * 2 + tid; i < (size_t)r1 * 2; i += nthreads) deltas[i] = 0.0f;
}

__global__ void __launch_bounds__(kRayBlock)
k_march_train_write(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                    const uint8_t* __restrict__ grid, float bound, float dt_gamma,
                    uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                    const float* __restrict__ nears, const float* __restrict__ fars,
                    float* __restrict__ xyzs, float* __restrict__ dirs,
                    float* __restrict__ deltas, const int32_t* __restrict__ rays,
                    const int32_t* __restrict__ counter, const float* __restrict__ noises,
                    uint32_t zero_tail_end) {
    const uint32_t i = blockIdx.x * kRayBlock + threadIdx.x;
    if (zero_tail_end > 0) {
        const uint32_t z1 = min(M, zero_tail_end);
        const uint32_t z0 = min((uint32_t)counter[0], z1);
        zero_rows(xyzs, dirs, deltas, z0, z1, i, gridDim.x * kRayBlock);
    }
    if (i >= N) return;
    const uint32_t n = (uint32_t)rays[(size_t)i * 3];
    const uint32_t offset = (uint32_t)rays[(size_t)i * 3 + 1];
    const uint32_t num_steps = (uint32_t)rays[(size_t)i * 3 + 2];
    if (num_steps == 0) return;
    if (offset + num_steps > M) {
        // dropped ray (raymarching.cu:457).  Offsets are monotone, so only the first
        // dropped ray starts inside the buffer; it clears what would stay unwritten.
        if (zero_tail_end > 0 && offset < M)
            zero_rows(xyzs, dirs, deltas, offset, min(M, zero_tail_end), 0, 1);
        return;
    }
    const MarchParams p = make_march_params(bound, dt_gamma, max_steps, C, H);
    const float* o = rays_o + (size_t)n * 3;
    const float* d = rays_d + (size_t)n * 3;
    const RayGeom r = make_ray(__ldg(o), __ldg(o + 1), __ldg(o + 2), __ldg(d), __ldg(d + 1),
                               __ldg(d + 2));
    const float far = __ldg(fars + n);
    float t = first_t(p, __ldg(nears + n), __ldg(noises + n));
    float last_t = t;
    float* px = xyzs + (size_t)offset * 3;
    float* pd = dirs + (size_t)offset * 3;
    float2* pl = reinterpret_cast<float2*>(deltas) + offset;
    uint32_t step = 0;
    while (t < far && step < num_steps) {
        float tt;
        const Probe q = probe_at(p, r, grid, t, &tt);
        if (q.occ) {
            px[0] = q.x; px[1] = q.y; px[2] = q.z;
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            t = __fadd_rn(t, q.dt);
            *pl = make_float2(q.dt, __fsub_rn(t, last_t));
            last_t = t;
            px += 3; pd += 3; ++pl;
            ++step;
        } else {
            t = skip_to(p, t, tt);
        }
    }
}

// ---------------------------------------------------------------------------
// march_rays (inference)  (reference kernel raymarching.cu:809-928)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kRayBlock)
k_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* __restrict__ rays_alive,
             const float* __restrict__ rays_t, const float* __restrict__ rays_o,
             const float* __restrict__ rays_d, float bound, float dt_gamma, uint32_t max_steps,
             uint32_t C, uint32_t H, const uint8_t* __restrict__ grid,
             const float* __restrict__ nears, const float* __restrict__ fars,
             float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
             const float* __restrict__ noises, uint32_t M_padded) {
    const uint32_t n = blockIdx.x * kRayBlock + threadIdx.x;
    {
        const uint32_t used = n_alive * n_step;
        if (M_padded > used)
            zero_rows(xyzs, dirs, deltas, used, M_padded, n, gridDim.x * kRayBlock);
    }
    if (n >= n_alive) return;
    const int32_t index = rays_alive[n];
    float* px = xyzs + (size_t)n * n_step * 3;
    float* pd = dirs + (size_t)n * n_step * 3;
    float2* pl = reinterpret_cast<float2*>(deltas) + (size_t)n * n_step;
    uint32_t step = 0;
    if (index >= 0) {
        const MarchParams p = make_march_params(bound, dt_gamma, max_steps, C, H);
        const float* o = rays_o + (size_t)index * 3;
        const float* d = rays_d + (size_t)index * 3;
        const RayGeom r = make_ray(__ldg(o), __ldg(o + 1), __ldg(o + 2), __ldg(d), __ldg(d + 1),
                                   __ldg(d + 2));
        const float far = __ldg(fars + index);
        float t = __ldg(rays_t + index);
        t = __fmaf_rn(__ldg(noises + n), clamp_dt(p, t), t);
        float last_t = t;
        while (t < far && step < n_step) {
            float tt;
            const Probe q = probe_at(p, r, grid, t, &tt);
            if (q.occ) {
                px[0] = q.x; px[1] = q.y; px[2] = q.z;
                pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
                t = __fadd_rn(t, q.dt);
                *pl = make_float2(q.dt, __fsub_rn(t, last_t));
                last_t = t;
                px += 3; pd += 3; ++pl;
                ++step;
            } else {
                t = skip_to(p, t, tt);
            }
        }
    }
    for (; step < n_step; ++step) {  // unused slots read as "terminated" (delta == 0)
        px[0] = px[1] = px[2] = 0.0f;
        pd[0] = pd[1] = pd[2] = 0.0f;
        *pl = make_float2(0.0f, 0.0f);
        px += 3; pd += 3; ++pl;
    }
}

// ---------------------------------------------------------------------------
// Warp-cooperative sample emission (default path of both marchers).
//
// A thread-per-ray marcher stores 32 B per sample with eight 4-byte stores whose
// 32 lanes hit 32 different rays' segments: every warp store touches 32 sectors
// for 128 useful bytes, and the kernel runs at the L2's partial-sector write
// rate (6 % of HBM bandwidth, as the reference's).  Here a lane still walks its
// own ray through the occupancy grid — the DDA is inherently serial — but only
// stages (t, dt) of up to K accepted samples in shared memory.  The warp then
// writes each ray's K-sample run together: positions are re-evaluated from
// (o, d, t) with the same FMA + clamp, so every bit equals the serial kernel's,
// while global stores are contiguous 128-byte rows.
// ---------------------------------------------------------------------------
constexpr uint32_t kSmallMarch = 148u * 32u * 16u;  // below this, 1-warp CTAs
constexpr int kMK = 16;           // staged samples per ray per round
constexpr int kMS = kMK + 1;      // staging row stride (odd -> lanes of one ray hit distinct banks)

struct MarchStage {
    float t[32 * kMS];            // ray parameter of each staged sample   [lane][kMS]
    float dt[32 * kMS];           // its step
    float4 ray[32 * 2];           // (ox, oy, oz, dx), (dy, dz, -, -) of the lane's ray
    uint4 meta[32];               // x = exclusive prefix sum of the staged counts (train) or the
                                  //     staged count (inference), y = output row of the ray's
                                  //     first staged sample, z = bits of `last_t` at round start
    uint8_t owner[32 * kMK];      // flattened staged sample -> lane
};

struct LaneMarch {
    float t, far, last_t;
    uint32_t step, limit;
    bool active;
};

// Emits staged sample s of lane j's ray to output row `row`: one lane per sample, neighbouring
// lanes write neighbouring rows (streaming stores: the outputs are not read again by this
// kernel and must not evict the occupancy bitfield from L1).  Same operations as the serial
// kernels (reference raymarching.cu:498-507), so every bit is identical.
__device__ __forceinline__ void emit_sample(const MarchStage& st, uint32_t j, uint32_t s,
                                            size_t row, float last_t, float bound,
                                            float* __restrict__ xyzs, float* __restrict__ dirs,
                                            float* __restrict__ deltas) {
    const float4 a = st.ray[2 * j];
    const float4 b = st.ray[2 * j + 1];
    const float t = st.t[j * kMS + s], dt = st.dt[j * kMS + s];
    const float prev = s > 0 ? __fadd_rn(st.t[j * kMS + s - 1], st.dt[j * kMS + s - 1]) : last_t;
    float* px = xyzs + row * 3;
    float* pd = dirs + row * 3;
    __stcs(px + 0, fminf(bound, fmaxf(-bound, __fmaf_rn(a.w, t, a.x))));
    __stcs(px + 1, fminf(bound, fmaxf(-bound, __fmaf_rn(b.x, t, a.y))));
    __stcs(px + 2, fminf(bound, fmaxf(-bound, __fmaf_rn(b.y, t, a.z))));
    __stcs(pd + 0, a.w);
    __stcs(pd + 1, b.x);
    __stcs(pd + 2, b.y);
    __stcs(reinterpret_cast<float2*>(deltas) + row,
           make_float2(dt, __fsub_rn(__fadd_rn(t, dt), prev)));
}

__device__ __forceinline__ void zero_sample(size_t row, float* __restrict__ xyzs,
                                            float* __restrict__ dirs,
                                            float* __restrict__ deltas) {
    float* px = xyzs + row * 3;
    float* pd = dirs + row * 3;
    __stcs(px + 0, 0.0f); __stcs(px + 1, 0.0f); __stcs(px + 2, 0.0f);
    __stcs(pd + 0, 0.0f); __stcs(pd + 1, 0.0f); __stcs(pd + 2, 0.0f);
    __stcs(reinterpret_cast<float2*>(deltas) + row, make_float2(0.0f, 0.0f));
}

// One round of the per-lane DDA.  The warp iterates in lockstep (exactly how SIMT executes the
// serial per-ray loop) and the round ends as soon as ANY lane has staged kMK samples or every ray
// is finished — lanes in the middle of a long empty stretch simply continue in the next round,
// so the total number of DDA iterations equals the serial kernel's.  Must be called by all 32
// lanes.  Returns the number of samples this lane staged.
__device__ __forceinline__ uint32_t march_round(const MarchParams& p, const RayGeom& r,
                                                const uint8_t* __restrict__ grid, MarchStage& st,
                                                int lane, LaneMarch& m) {
    uint32_t cnt = 0;
    while (true) {
        if (m.active) {
            float tt;
            const Probe q = probe_at(p, r, grid, m.t, &tt);
            if (q.occ) {
                st.t[lane * kMS + cnt] = m.t;
                st.dt[lane * kMS + cnt] = q.dt;
                m.t = __fadd_rn(m.t, q.dt);
                m.last_t = m.t;
                ++cnt;
                ++m.step;
            } else {
                m.t = skip_to(p, m.t, tt);
            }
            m.active = m.t < m.far && m.step < m.limit;
        }
        if (__any_sync(0xffffffffu, cnt >= (uint32_t)kMK) || !__any_sync(0xffffffffu, m.active))
            break;
    }
    return cnt;
}

// Flattened emission of everything staged in this round: lane-per-sample, `owner` maps the
// flattened sample index to the staging lane, meta[j] = (first flattened index, first output row,
// last_t at round start) of lane j.  Must be called by all 32 lanes.
__device__ __forceinline__ void emit_round(MarchStage& st, int lane, uint32_t cnt, uint32_t row0,
                                           float last_t0, float bound, float* __restrict__ xyzs,
                                           float* __restrict__ dirs, float* __restrict__ deltas) {
    const uint32_t inc = warp_inclusive_scan(cnt, lane);
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    st.meta[lane] = make_uint4(inc - cnt, row0, __float_as_uint(last_t0), 0u);
    for (uint32_t s = 0; s < cnt; ++s) st.owner[inc - cnt + s] = (uint8_t)lane;
    __syncwarp();
    for (uint32_t q = lane; q < total; q += 32) {
        const uint32_t j = st.owner[q];
        const uint4 mj = st.meta[j];
        const uint32_t s = q - mj.x;
        emit_sample(st, j, s, (size_t)mj.y + s, __uint_as_float(mj.z), bound, xyzs, dirs, deltas);
    }
    __syncwarp();
}

// march_rays_train, phase 2, warp-cooperative (reference raymarching.cu:463-533).
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_march_train_write_coop(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                         const uint8_t* __restrict__ grid, float bound, float dt_gamma,
                         uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                         const float* __restrict__ nears, const float* __restrict__ fars,
                         float* __restrict__ xyzs, float* __restrict__ dirs,
                         float* __restrict__ deltas, const int32_t* __restrict__ rays,
                         const int32_t* __restrict__ counter, const float* __restrict__ noises,
                         uint32_t zero_tail_end, const float2* __restrict__ stash) {
    __shared__ MarchStage s_stage[BLOCK / 32];
    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31;
    MarchStage& st = s_stage[threadIdx.x >> 5];
    if (zero_tail_end > 0) {
        const uint32_t z1 = min(M, zero_tail_end);
        const uint32_t z0 = min((uint32_t)counter[0], z1);
        zero_rows(xyzs, dirs, deltas, z0, z1, i, gridDim.x * BLOCK);
    }
    const MarchParams p = make_march_params(bound, dt_gamma, max_steps, C, H);
    uint32_t offset = 0, n = 0;
    LaneMarch m = {0.0f, 0.0f, 0.0f, 0u, 0u, false};
    float ox = 0.f, oy = 0.f, oz = 0.f, dx = 1.f, dy = 1.f, dz = 1.f;
    if (i < N) {
        n = (uint32_t)rays[(size_t)i * 3];
        offset = (uint32_t)rays[(size_t)i * 3 + 1];
        m.limit = n < N ? (uint32_t)rays[(size_t)i * 3 + 2] : 0u;   // rows nobody wrote: skip
        if (m.limit > 0 && offset + m.limit > M) {
            // dropped ray (raymarching.cu:457).  Offsets are monotone, so only the first
            // dropped ray starts inside the buffer; it clears what would stay unwritten.
            if (zero_tail_end > 0 && offset < M)
                zero_rows(xyzs, dirs, deltas, offset, min(M, zero_tail_end), 0, 1);
            m.limit = 0;
        }
        if (m.limit > 0) {
            const float* o = rays_o + (size_t)n * 3;
            const float* d = rays_d + (size_t)n * 3;
            ox = __ldg(o); oy = __ldg(o + 1); oz = __ldg(o + 2);
            dx = __ldg(d); dy = __ldg(d + 1); dz = __ldg(d + 2);
            m.far = __ldg(fars + n);
            m.t = first_t(p, __ldg(nears + n), __ldg(noises + n));
            m.last_t = m.t;
            m.active = m.t < m.far;
        }
    }
    st.ray[2 * lane] = make_float4(ox, oy, oz, dx);
    st.ray[2 * lane + 1] = make_float4(dy, dz, 0.0f, 0.0f);
    const RayGeom r = make_ray(ox, oy, oz, dx, dy, dz);
    if (stash != nullptr) {
        // round 0: the samples phase 1 kept; the walk resumes behind the last of them
        static_assert(kStash <= kMK, "stash must fit one staging round");
        const uint32_t c0 = min(m.limit, (uint32_t)kStash);
        const float last_t0 = m.last_t;
        for (uint32_t s = 0; s < c0; ++s) {
            const float2 v = __ldcs(stash + stash_index(n, s));
            st.t[lane * kMS + s] = v.x;
            st.dt[lane * kMS + s] = v.y;
            m.t = __fadd_rn(v.x, v.y);
        }
        if (c0 > 0) {
            m.last_t = m.t;
            m.step = c0;
            m.active = m.t < m.far && m.step < m.limit;
        }
        emit_round(st, lane, c0, offset, last_t0, bound, xyzs, dirs, deltas);
    }
    while (__any_sync(0xffffffffu, m.active)) {
        const uint32_t row0 = offset + m.step;
        const float last_t0 = m.last_t;
        const uint32_t cnt = march_round(p, r, grid, st, lane, m);
        emit_round(st, lane, cnt, row0, last_t0, bound, xyzs, dirs, deltas);
    }
}

// march_rays (inference), warp-cooperative (reference kernel raymarching.cu:809-928).  The 32
// rays of a warp own the contiguous rows [n0*n_step, (n0+32)*n_step); every slot is written
// (zeros where a ray produced no sample, so the caller need not pre-zero), kMK slots per ray
// per round.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_march_rays_coop(uint32_t n_alive, uint32_t n_step, const int32_t* __restrict__ rays_alive,
                  const float* __restrict__ rays_t, const float* __restrict__ rays_o,
                  const float* __restrict__ rays_d, float bound, float dt_gamma,
                  uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* __restrict__ grid,
                  const float* __restrict__ nears, const float* __restrict__ fars,
                  float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
                  const float* __restrict__ noises, uint32_t M_padded) {
    __shared__ MarchStage s_stage[BLOCK / 32];
    const uint32_t n = blockIdx.x * BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31;
    MarchStage& st = s_stage[threadIdx.x >> 5];
    {
        const uint32_t used = n_alive * n_step;
        if (M_padded > used)
            zero_rows(xyzs, dirs, deltas, used, M_padded, n, gridDim.x * BLOCK);
    }
    const uint32_t n0 = n - lane;            // first ray of this warp
    if (n0 >= n_alive) return;               // warp-uniform
    const uint32_t nrays = min(32u, n_alive - n0);
    const MarchParams p = make_march_params(bound, dt_gamma, max_steps, C, H);
    LaneMarch m = {0.0f, 0.0f, 0.0f, 0u, n_step, false};
    float ox = 0.f, oy = 0.f, oz = 0.f, dx = 1.f, dy = 1.f, dz = 1.f;
    if (n < n_alive) {
        const int32_t index = rays_alive[n];
        if (index >= 0) {
            const float* o = rays_o + (size_t)index * 3;
            const float* d = rays_d + (size_t)index * 3;
            ox = __ldg(o); oy = __ldg(o + 1); oz = __ldg(o + 2);
            dx = __ldg(d); dy = __ldg(d + 1); dz = __ldg(d + 2);
            m.far = __ldg(fars + index);
            m.t = __ldg(rays_t + index);
            m.t = __fmaf_rn(__ldg(noises + n), clamp_dt(p, m.t), m.t);
            m.last_t = m.t;
            m.active = m.t < m.far;
        }
    }
    st.ray[2 * lane] = make_float4(ox, oy, oz, dx);
    st.ray[2 * lane + 1] = make_float4(dy, dz, 0.0f, 0.0f);
    const RayGeom r = make_ray(ox, oy, oz, dx, dy, dz);
    const uint32_t base_row = n * n_step;    // rows [n*n_step, (n+1)*n_step) belong to this ray
    while (__any_sync(0xffffffffu, m.active)) {
        const uint32_t row0 = base_row + m.step;
        const float last_t0 = m.last_t;
        const uint32_t cnt = march_round(p, r, grid, st, lane, m);
        emit_round(st, lane, cnt, row0, last_t0, bound, xyzs, dirs, deltas);
    }
    // unused slots read as "terminated" (delta == 0): the warp's rows are one contiguous block
    st.meta[lane] = make_uint4(m.step, 0u, 0u, 0u);
    __syncwarp();
    const uint32_t slots = nrays * n_step;
    for (uint32_t q = lane; q < slots; q += 32) {
        const uint32_t j = q / n_step, s = q - j * n_step;
        if (s >= st.meta[j].x) zero_sample((size_t)n0 * n_step + q, xyzs, dirs, deltas);
    }
}

// ---------------------------------------------------------------------------
// Compositing.  Each warp owns 32 rays; the rays' sample streams are staged
// through a per-warp shared-memory tile in chunks of kCK samples with coalesced
// loads (half a warp per ray), then every lane walks its own ray front to back in
// exactly the reference's operation order (raymarching.cu:618-645, 734-771).
// ---------------------------------------------------------------------------
constexpr int kCK = 16;                 // samples per ray per chunk
constexpr int kSS = kCK + 1;            // sigma row stride (odd -> conflict free)
constexpr int kSR = 3 * kCK + 1;        // rgb row stride
constexpr int kSD = 2 * kCK + 1;        // deltas row stride
constexpr int kWarpWords = 32 * (kSS + kSR + kSD);
constexpr int kCompBlock = 128;
constexpr size_t kCompSmem = (size_t)(kCompBlock / 32) * kWarpWords * sizeof(float);

// exp(-sigma*delta) exactly as the reference's `__expf(-s * d)` executes
__device__ __forceinline__ float alpha_of(float sigma, float delta) {
    return __fsub_rn(1.0f, __expf(-__fmul_rn(sigma, delta)));
}

// Stage samples [k0, k0+need_l) of every live lane's ray into the warp tile.
__device__ __forceinline__ void stage_chunk(const float* __restrict__ sigmas,
                                            const float* __restrict__ rgbs,
                                            const float* __restrict__ deltas, uint32_t first,
                                            uint32_t need, uint32_t mask, int lane, float* s_sig,
                                            float* s_rgb, float* s_del) {
    const int half = lane >> 4, sub = lane & 15;
#pragma unroll 1
    for (int p = 0; p < 16; ++p) {
        if (((mask >> (2 * p)) & 3u) == 0) continue;  // warp-uniform
        const int src = 2 * p + half;
        const uint32_t o = __shfl_sync(0xffffffffu, first, src);
        const uint32_t nd = __shfl_sync(0xffffffffu, need, src);
        if ((uint32_t)sub < nd) s_sig[src * kSS + sub] = __ldg(sigmas + o + sub);
        const float* gr = rgbs + (size_t)o * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t w = sub + 16 * c;
            if (w < 3 * nd) s_rgb[src * kSR + w] = __ldg(gr + w);
        }
        const float* gd = deltas + (size_t)o * 2;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint32_t w = sub + 16 * c;
            if (w < 2 * nd) s_del[src * kSD + w] = __ldg(gd + w);
        }
    }
}

// Two regimes in one kernel.  (1) DIRECT: the first kDirect samples of a ray are read by its own lane, kDB
// samples (20 independent loads) in flight at a time — with a trained field most rays terminate
// (T < T_thresh) within a few samples, and a 16-sample tile per ray would be mostly wasted traffic; the
// reference's one-load-round-trip-per-sample walk is latency bound there.  (2) TILE: whatever is still alive
// continues through the coalesced warp tile.  Every lane executes the reference's per-sample operation sequence
// in both regimes, so the results are bit-identical to either alone.
constexpr int kDB = 4;          // samples per direct batch
constexpr int kDirect = 16;     // samples walked directly before switching to the tile (multiple of kCK)
static_assert(kDirect % kCK == 0 && kDirect % kDB == 0, "the tile takes over on a chunk boundary");

// Pure direct walk: one lane per ray, one sample per iteration like the reference kernel, with the NEXT sample's six
// values loaded before the current one is composited (the loads do not depend on the arithmetic, only on the
// early-termination test, so the walk pays one memory latency per two samples instead of one per sample and reads at
// most one sample past the ray's last).
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_composite_train_fwd_direct(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                             const float* __restrict__ deltas, const int32_t* __restrict__ rays, uint32_t M, uint32_t N,
                             float T_thresh, float* __restrict__ weights_sum, float* __restrict__ depth,
                             float* __restrict__ image) {
    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= N) return;
    const int32_t* rr = rays + (size_t)i * 3;
    const uint32_t index = (uint32_t)__ldg(rr), offset = (uint32_t)__ldg(rr + 1), cnt = (uint32_t)__ldg(rr + 2);
    float T = 1.0f, r = 0.f, g = 0.f, b = 0.f, ws = 0.f, t = 0.f, d = 0.f;
    if (cnt != 0 && offset + cnt <= M) {
        const float* ps = sigmas + offset;
        const float2* pd = reinterpret_cast<const float2*>(deltas) + offset;
        const float* pr = rgbs + (size_t)offset * 3;
        float sg = __ldg(ps), c0 = __ldg(pr), c1 = __ldg(pr + 1), c2 = __ldg(pr + 2);
        float2 dl = __ldg(pd);
        for (uint32_t k = 0; k < cnt; ++k) {
            const uint32_t kn = min(k + 1, cnt - 1);
            const float sg_n = __ldg(ps + kn), c0_n = __ldg(pr + 3 * kn), c1_n = __ldg(pr + 3 * kn + 1),
                        c2_n = __ldg(pr + 3 * kn + 2);
            const float2 dl_n = __ldg(pd + kn);
            const float alpha = alpha_of(sg, dl.x);
            const float weight = __fmul_rn(alpha, T);
            r = __fmaf_rn(weight, c0, r);
            g = __fmaf_rn(weight, c1, g);
            b = __fmaf_rn(weight, c2, b);
            t = __fadd_rn(dl.y, t);
            d = __fmaf_rn(weight, t, d);
            ws = __fadd_rn(weight, ws);
            T = __fmul_rn(__fsub_rn(1.0f, alpha), T);
            if (T < T_thresh) break;
            sg = sg_n; c0 = c0_n; c1 = c1_n; c2 = c2_n; dl = dl_n;
        }
    }
    weights_sum[index] = ws;
    depth[index] = d;
    image[(size_t)index * 3] = r;
    image[(size_t)index * 3 + 1] = g;
    image[(size_t)index * 3 + 2] = b;
}

template <int BLOCK, int DIRECT>
__global__ void __launch_bounds__(BLOCK)
k_composite_train_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                      const float* __restrict__ deltas, const int32_t* __restrict__ rays,
                      uint32_t M, uint32_t N, float T_thresh, float* __restrict__ weights_sum,
                      float* __restrict__ depth, float* __restrict__ image) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* s_sig = smem + warp * kWarpWords;
    float* s_rgb = s_sig + 32 * kSS;
    float* s_del = s_rgb + 32 * kSR;

    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    uint32_t index = 0, offset = 0, cnt = 0;
    const bool valid = i < N;
    if (valid) {
        const int32_t* rr = rays + (size_t)i * 3;
        index = (uint32_t)__ldg(rr);
        offset = (uint32_t)__ldg(rr + 1);
        cnt = (uint32_t)__ldg(rr + 2);
    }
    bool live = valid && cnt != 0 && offset + cnt <= M;

    float T = 1.0f, r = 0.f, g = 0.f, b = 0.f, ws = 0.f, t = 0.f, d = 0.f;
    uint32_t k0 = 0;
    // ---- direct regime ----
    if (DIRECT > 0 && live) {
        const float* ps = sigmas + offset;
        const float2* pd = reinterpret_cast<const float2*>(deltas) + offset;
        const float* pr = rgbs + (size_t)offset * 3;
        const uint32_t lim = min(cnt, (uint32_t)DIRECT);
        for (uint32_t kb = 0; kb < lim && live; kb += kDB) {
            float sg[kDB], c0[kDB], c1[kDB], c2[kDB];
            float2 dl[kDB];
#pragma unroll
            for (int j = 0; j < kDB; ++j) {
                const uint32_t k = min(kb + j, cnt - 1);   // reads past the ray's end repeat its last sample
                sg[j] = __ldg(ps + k);
                dl[j] = __ldg(pd + k);
                c0[j] = __ldg(pr + 3 * k); c1[j] = __ldg(pr + 3 * k + 1); c2[j] = __ldg(pr + 3 * k + 2);
            }
#pragma unroll
            for (int j = 0; j < kDB; ++j) {
                if (kb + j >= cnt) { live = false; break; }
                const float alpha = alpha_of(sg[j], dl[j].x);
                const float weight = __fmul_rn(alpha, T);
                r = __fmaf_rn(weight, c0[j], r);
                g = __fmaf_rn(weight, c1[j], g);
                b = __fmaf_rn(weight, c2[j], b);
                t = __fadd_rn(dl[j].y, t);
                d = __fmaf_rn(weight, t, d);
                ws = __fadd_rn(weight, ws);
                T = __fmul_rn(__fsub_rn(1.0f, alpha), T);
                if (T < T_thresh) { live = false; break; }
            }
        }
        if (cnt <= (uint32_t)DIRECT) live = false;
    }
    k0 = DIRECT;
    // ---- tile regime ----
    for (;;) {
        const uint32_t need = live ? min(cnt - k0, (uint32_t)kCK) : 0u;
        const uint32_t mask = __ballot_sync(0xffffffffu, need != 0);
        if (mask == 0) break;
        stage_chunk(sigmas, rgbs, deltas, offset + k0, need, mask, lane, s_sig, s_rgb, s_del);
        __syncwarp();
        const float* ps = s_sig + lane * kSS;
        const float* pr = s_rgb + lane * kSR;
        const float* pd = s_del + lane * kSD;
        for (uint32_t k = 0; k < need; ++k) {
            const float alpha = alpha_of(ps[k], pd[2 * k]);
            const float weight = __fmul_rn(alpha, T);
            r = __fmaf_rn(weight, pr[3 * k], r);
            g = __fmaf_rn(weight, pr[3 * k + 1], g);
            b = __fmaf_rn(weight, pr[3 * k + 2], b);
            t = __fadd_rn(pd[2 * k + 1], t);
            d = __fmaf_rn(weight, t, d);
            ws = __fadd_rn(weight, ws);
            T = __fmul_rn(__fsub_rn(1.0f, alpha), T);
            if (T < T_thresh) { live = false; break; }
        }
        __syncwarp();
        k0 += kCK;
        if (k0 >= cnt) live = false;
    }
    if (valid) {
        weights_sum[index] = ws;
        depth[index] = d;
        image[(size_t)index * 3] = r;
        image[(size_t)index * 3 + 1] = g;
        image[(size_t)index * 3 + 2] = b;
    }
}

template <int BLOCK, int DIRECT>
__global__ void __launch_bounds__(BLOCK)
k_composite_train_bwd(const float* __restrict__ grad_weights_sum,
                      const float* __restrict__ grad_image, const float* __restrict__ sigmas,
                      const float* __restrict__ rgbs, const float* __restrict__ deltas,
                      const int32_t* __restrict__ rays, const float* __restrict__ weights_sum,
                      const float* __restrict__ image, uint32_t M, uint32_t N, float T_thresh,
                      float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* s_sig = smem + warp * kWarpWords;
    float* s_rgb = s_sig + 32 * kSS;
    float* s_del = s_rgb + 32 * kSR;

    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    uint32_t index = 0, offset = 0, cnt = 0;
    const bool valid = i < N;
    if (valid) {
        index = (uint32_t)__ldg(rays + (size_t)i * 3);
        offset = (uint32_t)__ldg(rays + (size_t)i * 3 + 1);
        cnt = (uint32_t)__ldg(rays + (size_t)i * 3 + 2);
    }
    bool live = valid && cnt != 0 && offset + cnt <= M;

    float gi0 = 0.f, gi1 = 0.f, gi2 = 0.f, gws1 = 0.f, rf = 0.f, gf = 0.f, bf = 0.f;
    if (live) {
        gi0 = __ldg(grad_image + (size_t)index * 3);
        gi1 = __ldg(grad_image + (size_t)index * 3 + 1);
        gi2 = __ldg(grad_image + (size_t)index * 3 + 2);
        rf = __ldg(image + (size_t)index * 3);
        gf = __ldg(image + (size_t)index * 3 + 1);
        bf = __ldg(image + (size_t)index * 3 + 2);
        gws1 = __fmul_rn(__ldg(grad_weights_sum + index),
                         __fsub_rn(1.0f, __ldg(weights_sum + index)));
    }
    float T = 1.0f, r = 0.f, g = 0.f, b = 0.f;
    uint32_t k0 = 0;
    const int half = lane >> 4, sub = lane & 15;
    // ---- direct regime (see k_composite_train_fwd): the lane walks and writes its own first samples ----
    if (DIRECT > 0 && live) {
        const float* ps = sigmas + offset;
        const float2* pd = reinterpret_cast<const float2*>(deltas) + offset;
        const float* pr = rgbs + (size_t)offset * 3;
        float* gs = grad_sigmas + offset;
        float* gr = grad_rgbs + (size_t)offset * 3;
        const uint32_t lim = min(cnt, (uint32_t)DIRECT);
        for (uint32_t kb = 0; kb < lim && live; kb += kDB) {
            float sg[kDB], c0[kDB], c1[kDB], c2[kDB], d0[kDB];
#pragma unroll
            for (int j = 0; j < kDB; ++j) {
                const uint32_t k = min(kb + j, cnt - 1);
                sg[j] = __ldg(ps + k);
                d0[j] = __ldg(pd + k).x;
                c0[j] = __ldg(pr + 3 * k); c1[j] = __ldg(pr + 3 * k + 1); c2[j] = __ldg(pr + 3 * k + 2);
            }
#pragma unroll
            for (int j = 0; j < kDB; ++j) {
                const uint32_t k = kb + j;
                if (k >= cnt) { live = false; break; }
                const float alpha = alpha_of(sg[j], d0[j]);
                const float weight = __fmul_rn(alpha, T);
                r = __fmaf_rn(weight, c0[j], r);
                g = __fmaf_rn(weight, c1[j], g);
                b = __fmaf_rn(weight, c2[j], b);
                T = __fmul_rn(__fsub_rn(1.0f, alpha), T);
                gr[3 * k] = __fmul_rn(gi0, weight);
                gr[3 * k + 1] = __fmul_rn(gi1, weight);
                gr[3 * k + 2] = __fmul_rn(gi2, weight);
                const float t0 = __fmaf_rn(c0[j], T, -__fsub_rn(rf, r));
                const float t1 = __fmaf_rn(c1[j], T, -__fsub_rn(gf, g));
                const float t2 = __fmaf_rn(c2[j], T, -__fsub_rn(bf, b));
                const float acc = __fmaf_rn(gi2, t2, __fmaf_rn(gi0, t0, __fmul_rn(gi1, t1)));
                gs[k] = __fmul_rn(d0[j], __fadd_rn(gws1, acc));
                if (T < T_thresh) { live = false; break; }
            }
        }
        if (cnt <= (uint32_t)DIRECT) live = false;
    }
    k0 = DIRECT;
    for (;;) {
        const uint32_t need = live ? min(cnt - k0, (uint32_t)kCK) : 0u;
        const uint32_t mask = __ballot_sync(0xffffffffu, need != 0);
        if (mask == 0) break;
        stage_chunk(sigmas, rgbs, deltas, offset + k0, need, mask, lane, s_sig, s_rgb, s_del);
        __syncwarp();
        float* ps = s_sig + lane * kSS;
        float* pr = s_rgb + lane * kSR;
        const float* pd = s_del + lane * kSD;
        uint32_t done = 0;
        for (uint32_t k = 0; k < need; ++k) {
            const float d0 = pd[2 * k];
            const float c0 = pr[3 * k], c1 = pr[3 * k + 1], c2 = pr[3 * k + 2];
            const float alpha = alpha_of(ps[k], d0);
            const float weight = __fmul_rn(alpha, T);
            r = __fmaf_rn(weight, c0, r);
            g = __fmaf_rn(weight, c1, g);
            b = __fmaf_rn(weight, c2, b);
            T = __fmul_rn(__fsub_rn(1.0f, alpha), T);
            pr[3 * k] = __fmul_rn(gi0, weight);
            pr[3 * k + 1] = __fmul_rn(gi1, weight);
            pr[3 * k + 2] = __fmul_rn(gi2, weight);
            const float t0 = __fmaf_rn(c0, T, -__fsub_rn(rf, r));
            const float t1 = __fmaf_rn(c1, T, -__fsub_rn(gf, g));
            const float t2 = __fmaf_rn(c2, T, -__fsub_rn(bf, b));
            const float acc = __fmaf_rn(gi2, t2, __fmaf_rn(gi0, t0, __fmul_rn(gi1, t1)));
            ps[k] = __fmul_rn(d0, __fadd_rn(gws1, acc));
            ++done;
            if (T < T_thresh) { live = false; break; }
        }
        __syncwarp();
        // coalesced write-back of the gradients produced in this chunk
        const uint32_t first = offset + k0;
#pragma unroll 1
        for (int p = 0; p < 16; ++p) {
            if (((mask >> (2 * p)) & 3u) == 0) continue;
            const int src = 2 * p + half;
            const uint32_t o = __shfl_sync(0xffffffffu, first, src);
            const uint32_t nd = __shfl_sync(0xffffffffu, done, src);
            if ((uint32_t)sub < nd) grad_sigmas[o + sub] = s_sig[src * kSS + sub];
            float* gr = grad_rgbs + (size_t)o * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const uint32_t w = sub + 16 * c;
                if (w < 3 * nd) gr[w] = s_rgb[src * kSR + w];
            }
        }
        __syncwarp();
        k0 += kCK;
        if (k0 >= cnt) live = false;
    }
}

// composite_rays (inference, in place)  (reference kernel raymarching.cu:967-1053)
__global__ void __launch_bounds__(kCompBlock)
k_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh,
                 int32_t* __restrict__ rays_alive, float* __restrict__ rays_t,
                 const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                 const float* __restrict__ deltas, float* __restrict__ weights_sum,
                 float* __restrict__ depth, float* __restrict__ image) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* s_sig = smem + warp * kWarpWords;
    float* s_rgb = s_sig + 32 * kSS;
    float* s_del = s_rgb + 32 * kSR;

    const uint32_t n = blockIdx.x * kCompBlock + threadIdx.x;
    const bool valid = n < n_alive;
    int32_t index = -1;
    if (valid) index = rays_alive[n];
    bool live = valid && index >= 0;

    float t = 0.f, ws = 0.f, d = 0.f, r = 0.f, g = 0.f, b = 0.f;
    if (live) {
        t = rays_t[index];
        ws = weights_sum[index];
        d = depth[index];
        r = image[(size_t)index * 3];
        g = image[(size_t)index * 3 + 1];
        b = image[(size_t)index * 3 + 2];
    }
    const bool owner = live;
    uint32_t step = 0;
    uint32_t k0 = 0;
    const uint32_t first = n * n_step;
    for (;;) {
        const uint32_t need = live ? min(n_step - k0, (uint32_t)kCK) : 0u;
        const uint32_t mask = __ballot_sync(0xffffffffu, need != 0);
        if (mask == 0) break;
        stage_chunk(sigmas, rgbs, deltas, first + k0, need, mask, lane, s_sig, s_rgb, s_del);
        __syncwarp();
        const float* ps = s_sig + lane * kSS;
        const float* pr = s_rgb + lane * kSR;
        const float* pd = s_del + lane * kSD;
        for (uint32_t k = 0; k < need; ++k) {
            const float d0 = pd[2 * k];
            if (d0 == 0.0f) { live = false; break; }
            const float alpha = alpha_of(ps[k], d0);
            const float T = __fsub_rn(1.0f, ws);
            const float weight = __fmul_rn(alpha, T);
            ws = __fadd_rn(ws, weight);
            t = __fadd_rn(pd[2 * k + 1], t);
            d = __fmaf_rn(weight, t, d);
            r = __fmaf_rn(weight, pr[3 * k], r);
            g = __fmaf_rn(weight, pr[3 * k + 1], g);
            b = __fmaf_rn(weight, pr[3 * k + 2], b);
            if (T < T_thresh) { live = false; break; }
            ++step;
        }
        __syncwarp();
        k0 += kCK;
        if (k0 >= n_step) live = false;
    }
    if (owner) {
        if (step < n_step) rays_alive[n] = -1;
        else rays_t[index] = t;
        weights_sum[index] = ws;
        depth[index] = d;
        image[(size_t)index * 3] = r;
        image[(size_t)index * 3 + 1] = g;
        image[(size_t)index * 3 + 2] = b;
    }
}

bool g_comp_attr_set = false;
int ensure_comp_attrs() {
    if (g_comp_attr_set) return NVSF_OK;
    cudaError_t e;
    e = cudaFuncSetAttribute(k_composite_train_fwd<kCompBlock, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kCompSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_composite_train_fwd<kCompBlock, kDirect>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kCompSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_composite_train_bwd<kCompBlock, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kCompSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_composite_train_bwd<kCompBlock, kDirect>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kCompSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_composite_rays, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kCompSmem);
    if (e != cudaSuccess) return (int)e;
    g_comp_attr_set = true;
    return NVSF_OK;
}

inline bool march_cfg_ok(uint32_t C, uint32_t H, uint32_t max_steps) {
    // 3x10-bit Morton code; C*H^3 must fit the uint32 index of the reference
    return C >= 1 && C <= 31 && H >= 1 && H <= 1024 && max_steps >= 1 &&
           (uint64_t)C * H * H * H <= 0xffffffffull;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int nvsf_abi_version(void) { return NVSF_B200_ABI_VERSION; }

const char* nvsf_status_string(int status) {
    if (status == NVSF_OK) return "ok";
    if (status == NVSF_E_INVALID) return "nvsf: invalid argument";
    if (status == NVSF_E_WORKSPACE) return "nvsf: workspace too small";
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "nvsf: unknown status";
}

int nvsf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb,
                            uint32_t N, float min_near, float* nears, float* fars,
                            void* stream) {
    if (N == 0) return NVSF_OK;
    if (!rays_o || !rays_d || !aabb || !nears || !fars) return NVSF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const bool aligned = ((reinterpret_cast<uintptr_t>(rays_o) | reinterpret_cast<uintptr_t>(rays_d) |
                           reinterpret_cast<uintptr_t>(nears) | reinterpret_cast<uintptr_t>(fars)) & 15u) == 0;
    const uint32_t n4 = aligned ? N / 4 : 0;
    if (n4)
        k_near_far_from_aabb4<<<nvsf_div_up(n4, 128u), 128, 0, s>>>(
            reinterpret_cast<const float4*>(rays_o), reinterpret_cast<const float4*>(rays_d), aabb, n4, min_near,
            reinterpret_cast<float4*>(nears), reinterpret_cast<float4*>(fars));
    if (4 * n4 < N)
        k_near_far_from_aabb<<<nvsf_div_up(N - 4 * n4, 128u), 128, 0, s>>>(rays_o, rays_d, aabb, 4 * n4, N, min_near,
                                                                       nears, fars);
    return nvsf_launch_status();
}

int nvsf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N,
                      float* coords, void* stream) {
    if (N == 0) return NVSF_OK;
    if (!rays_o || !rays_d || !coords) return NVSF_E_INVALID;
    k_sph_from_ray<kTileBlock>
        <<<nvsf_div_up(N, (uint32_t)kTileBlock), kTileBlock, 0, (cudaStream_t)stream>>>(
            rays_o, rays_d, radius, N, coords);
    return nvsf_launch_status();
}

int nvsf_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream) {
    if (N == 0) return NVSF_OK;
    if (!coords || !indices) return NVSF_E_INVALID;
    k_morton3D<kTileBlock>
        <<<nvsf_div_up(N, (uint32_t)kTileBlock), kTileBlock, 0, (cudaStream_t)stream>>>(
            coords, N, indices);
    return nvsf_launch_status();
}

int nvsf_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream) {
    if (N == 0) return NVSF_OK;
    if (!coords || !indices) return NVSF_E_INVALID;
    k_morton3D_invert<kTileBlock>
        <<<nvsf_div_up(N, (uint32_t)kTileBlock), kTileBlock, 0, (cudaStream_t)stream>>>(
            indices, N, coords);
    return nvsf_launch_status();
}

int nvsf_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield,
                  void* stream) {
    if (N == 0) return NVSF_OK;
    if (!grid || !bitfield) return NVSF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t done = 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(grid) & 15u) == 0) &&
                         ((reinterpret_cast<uintptr_t>(bitfield) & 3u) == 0);
    const uint32_t bytes_per_super = kPackIters * 32 * 4 / 8;  // 128 bytes out per warp pass
    if (aligned && N >= bytes_per_super) {
        const uint32_t n_super = N / bytes_per_super;
        const uint32_t warps_per_block = 256 / 32;
        k_packbits_vec<<<nvsf_div_up(n_super, warps_per_block), 256, 0, s>>>(
            reinterpret_cast<const float4*>(grid), n_super, density_thresh,
            reinterpret_cast<uint32_t*>(bitfield));
        done = n_super * bytes_per_super;
    }
    if (done < N) {
        k_packbits_scalar<<<nvsf_div_up(N - done, 256u), 256, 0, s>>>(grid, done, N,
                                                                      density_thresh, bitfield);
    }
    return nvsf_launch_status();
}

static size_t march_stash_offset(uint32_t N) {   // bytes, 16-byte aligned
    const size_t nblk = nvsf_div_up((size_t)N, (size_t)32);   // the count pass runs 32- or kRayBlock-ray CTAs
    return ((4 + nblk + (size_t)N) * sizeof(uint32_t) + 15) & ~(size_t)15;
}

size_t nvsf_march_rays_train_workspace_bytes(uint32_t N) {
    return march_stash_offset(N) +
           nvsf_div_up((size_t)N, (size_t)32) * 32 * kStash * sizeof(float2);
}

int nvsf_march_rays_train_count(const float* rays_o, const float* rays_d, const uint8_t* grid,
                                float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                                uint32_t C, uint32_t H, const float* nears, const float* fars,
                                int32_t* rays, int32_t* counter, const float* noises,
                                void* workspace, size_t workspace_bytes, void* stream) {
    if (N == 0) return NVSF_OK;
    if (!rays_o || !rays_d || !grid || !nears || !fars || !rays || !counter || !noises ||
        !workspace)
        return NVSF_E_INVALID;
    if (!march_cfg_ok(C, H, max_steps)) return NVSF_E_INVALID;
    if (workspace_bytes < nvsf_march_rays_train_workspace_bytes(N)) return NVSF_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t* ws = reinterpret_cast<uint32_t*>(workspace);
    float2* stash = reinterpret_cast<float2*>(reinterpret_cast<char*>(workspace) +
                                              march_stash_offset(N));
    if (N <= 16u * 1024u) {   // small batches (the trainer's 4096 rays): one warp per CTA uses every SM
        const uint32_t nblk = nvsf_div_up(N, 32u);
        k_march_train_count<32><<<nblk, 32, 0, s>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears,
                                                    fars, noises, counter, ws, nblk, stash);
        k_march_train_scan<<<1, 1024, 0, s>>>(ws, nblk, N, counter);
        k_march_train_rows<32><<<nblk, 32, 0, s>>>(ws, nblk, N, rays);
    } else {
        const uint32_t nblk = nvsf_div_up(N, (uint32_t)kRayBlock);
        k_march_train_count<kRayBlock><<<nblk, kRayBlock, 0, s>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C,
                                                                  H, nears, fars, noises, counter, ws, nblk, stash);
        k_march_train_scan<<<1, 1024, 0, s>>>(ws, nblk, N, counter);
        k_march_train_rows<kRayBlock><<<nblk, kRayBlock, 0, s>>>(ws, nblk, N, rays);
    }
    return nvsf_launch_status();
}

int nvsf_march_rays_train_write_ws(const float* rays_o, const float* rays_d, const uint8_t* grid,
                                   float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                                   uint32_t C, uint32_t H, uint32_t M, const float* nears,
                                   const float* fars, float* xyzs, float* dirs, float* deltas,
                                   const int32_t* rays, const int32_t* counter,
                                   const float* noises, uint32_t zero_tail_end,
                                   const void* workspace, size_t workspace_bytes, void* stream) {
    if (N == 0) return NVSF_OK;
    if (!rays_o || !rays_d || !grid || !nears || !fars || !rays || !counter || !noises)
        return NVSF_E_INVALID;
    if (M > 0 && (!xyzs || !dirs || !deltas)) return NVSF_E_INVALID;
    if (!march_cfg_ok(C, H, max_steps)) return NVSF_E_INVALID;
    if (workspace && workspace_bytes < nvsf_march_rays_train_workspace_bytes(N))
        return NVSF_E_WORKSPACE;
    const float2* stash =
        workspace ? reinterpret_cast<const float2*>(reinterpret_cast<const char*>(workspace) +
                                                    march_stash_offset(N))
                  : nullptr;
    const uint32_t nblk = nvsf_div_up(N, (uint32_t)kRayBlock);
    // few rays: one warp per CTA so that every SM gets work (the walk is latency bound)
    if (nvsf_march_mode() && N <= kSmallMarch)
        k_march_train_write_coop<32><<<nvsf_div_up(N, 32u), 32, 0, (cudaStream_t)stream>>>(
            rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs,
            deltas, rays, counter, noises, zero_tail_end, stash);
    else if (nvsf_march_mode())
        k_march_train_write_coop<kRayBlock><<<nblk, kRayBlock, 0, (cudaStream_t)stream>>>(
            rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs,
            deltas, rays, counter, noises, zero_tail_end, stash);
    else
        k_march_train_write<<<nblk, kRayBlock, 0, (cudaStream_t)stream>>>(
            rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs,
            deltas, rays, counter, noises, zero_tail_end);
    return nvsf_launch_status();
}

int nvsf_march_rays_train_write(const float* rays_o, const float* rays_d, const uint8_t* grid,
                                float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                                uint32_t C, uint32_t H, uint32_t M, const float* nears,
                                const float* fars, float* xyzs, float* dirs, float* deltas,
                                const int32_t* rays, const int32_t* counter,
                                const float* noises, uint32_t zero_tail_end, void* stream) {
    return nvsf_march_rays_train_write_ws(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C,
                                          H, M, nears, fars, xyzs, dirs, deltas, rays, counter,
                                          noises, zero_tail_end, nullptr, 0, stream);
}

int nvsf_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid,
                          float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                          uint32_t C, uint32_t H, uint32_t M, const float* nears,
                          const float* fars, float* xyzs, float* dirs, float* deltas,
                          int32_t* rays, int32_t* counter, const float* noises,
                          void* workspace, size_t workspace_bytes, void* stream) {
    int st = nvsf_march_rays_train_count(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C,
                                         H, nears, fars, rays, counter, noises, workspace,
                                         workspace_bytes, stream);
    if (st != NVSF_OK) return st;
    return nvsf_march_rays_train_write_ws(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C,
                                          H, M, nears, fars, xyzs, dirs, deltas, rays, counter,
                                          noises, 0u, workspace, workspace_bytes, stream);
}

int nvsf_composite_rays_train_forward(const float* sigmas, const float* rgbs,
                                      const float* deltas, const int32_t* rays, uint32_t M,
                                      uint32_t N, float T_thresh, float* weights_sum,
                                      float* depth, float* image, void* stream) {
    if (N == 0) return NVSF_OK;
    if (!rays || !weights_sum || !depth || !image) return NVSF_E_INVALID;
    if (M > 0 && (!sigmas || !rgbs || !deltas)) return NVSF_E_INVALID;
    int st = ensure_comp_attrs();
    if (st != NVSF_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    // option "composite_mode": 0 tile, 1 direct prologue + tile, 2 direct (default).  Measured on B200 against the
    // reference kernel (profiles/r02_bench_ops_*.jsonl): with early termination (trained fields) the direct walk
    // reads only what it composites and matches or beats the reference in every configuration, the tile over-reads
    // (camera frame 0.24 vs 0.18 ms); without any termination the tile is 3 % faster (1.71 vs 1.75 ms, reference 1.90).
    const int mode = nvsf_composite_mode();
    // small batches (the trainer's 4096 rays): one warp per CTA spreads the rays over all SMs
    const bool small = N <= 32u * 1024u;
    if (mode == 2) {
        if (small) k_composite_train_fwd_direct<32><<<nvsf_div_up(N, 32u), 32, 0, s>>>(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image);
        // 64-ray CTAs: rays of very different lengths share an SM, finer CTAs even the tail out
        else k_composite_train_fwd_direct<64><<<nvsf_div_up(N, 64u), 64, 0, s>>>(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image);
    } else if (mode == 1) {
        if (small) k_composite_train_fwd<32, kDirect><<<nvsf_div_up(N, 32u), 32, kCompSmem / (kCompBlock / 32), s>>>(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image);
        else k_composite_train_fwd<kCompBlock, kDirect><<<nvsf_div_up(N, (uint32_t)kCompBlock), kCompBlock, kCompSmem, s>>>(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image);
    } else {
        if (small) k_composite_train_fwd<32, 0><<<nvsf_div_up(N, 32u), 32, kCompSmem / (kCompBlock / 32), s>>>(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image);
        else k_composite_train_fwd<kCompBlock, 0><<<nvsf_div_up(N, (uint32_t)kCompBlock), kCompBlock, kCompSmem, s>>>(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image);
    }
    return nvsf_launch_status();
}

int nvsf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                       const float* sigmas, const float* rgbs,
                                       const float* deltas, const int32_t* rays,
                                       const float* weights_sum, const float* image,
                                       uint32_t M, uint32_t N, float T_thresh,
                                       float* grad_sigmas, float* grad_rgbs, void* stream) {
    if (N == 0 || M == 0) return NVSF_OK;
    if (!grad_weights_sum || !grad_image || !sigmas || !rgbs || !deltas || !rays ||
        !weights_sum || !image || !grad_sigmas || !grad_rgbs)
        return NVSF_E_INVALID;
    int st = ensure_comp_attrs();
    if (st != NVSF_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    // option "composite_bwd_mode": 0 tile, 1 direct prologue + tile, 2 (default) = 1 for the small batches of a training
    // step, 0 for frames (measured on B200, profiles/r02_bench_ops_*.jsonl: 4096 rays 0.038 vs 0.040 ms, 529 408
    // rays 0.73 vs 0.45 ms)
    const int bmode = nvsf_composite_bwd_mode();
    const bool hybrid = bmode == 1 || (bmode == 2 && N <= 32u * 1024u);
#define NVSF_CB_ARGS grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N, T_thresh, grad_sigmas, grad_rgbs
    if (N <= 32u * 1024u) {
        if (hybrid) k_composite_train_bwd<32, kDirect><<<nvsf_div_up(N, 32u), 32, kCompSmem / (kCompBlock / 32), s>>>(NVSF_CB_ARGS);
        else k_composite_train_bwd<32, 0><<<nvsf_div_up(N, 32u), 32, kCompSmem / (kCompBlock / 32), s>>>(NVSF_CB_ARGS);
    } else {
        if (hybrid) k_composite_train_bwd<kCompBlock, kDirect><<<nvsf_div_up(N, (uint32_t)kCompBlock), kCompBlock, kCompSmem, s>>>(NVSF_CB_ARGS);
        else k_composite_train_bwd<kCompBlock, 0><<<nvsf_div_up(N, (uint32_t)kCompBlock), kCompBlock, kCompSmem, s>>>(NVSF_CB_ARGS);
    }
#undef NVSF_CB_ARGS
    return nvsf_launch_status();
}

int nvsf_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive,
                    const float* rays_t, const float* rays_o, const float* rays_d, float bound,
                    float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                    const uint8_t* grid, const float* nears, const float* fars, float* xyzs,
                    float* dirs, float* deltas, const float* noises, uint32_t M_padded,
                    void* stream) {
    if (n_alive == 0 && M_padded == 0) return NVSF_OK;
    if (!xyzs || !dirs || !deltas) return NVSF_E_INVALID;
    if (n_alive > 0 && (!rays_alive || !rays_t || !rays_o || !rays_d || !grid || !nears ||
                        !fars || !noises || n_step == 0))
        return NVSF_E_INVALID;
    if (!march_cfg_ok(C, H, max_steps)) return NVSF_E_INVALID;
    const uint32_t nblk = n_alive > 0 ? nvsf_div_up(n_alive, (uint32_t)kRayBlock) : 1u;
    if (nvsf_march_mode() && n_alive <= kSmallMarch)
        k_march_rays_coop<32><<<n_alive > 0 ? nvsf_div_up(n_alive, 32u) : 1u, 32, 0,
                                (cudaStream_t)stream>>>(
            n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H,
            grid, nears, fars, xyzs, dirs, deltas, noises, M_padded);
    else if (nvsf_march_mode())
        k_march_rays_coop<kRayBlock><<<nblk, kRayBlock, 0, (cudaStream_t)stream>>>(
            n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H,
            grid, nears, fars, xyzs, dirs, deltas, noises, M_padded);
    else
        k_march_rays<<<nblk, kRayBlock, 0, (cudaStream_t)stream>>>(
            n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H,
            grid, nears, fars, xyzs, dirs, deltas, noises, M_padded);
    return nvsf_launch_status();
}

int nvsf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive,
                        float* rays_t, const float* sigmas, const float* rgbs,
                        const float* deltas, float* weights_sum, float* depth, float* image,
                        void* stream) {
    if (n_alive == 0 || n_step == 0) return NVSF_OK;
    if (!rays_alive || !rays_t || !sigmas || !rgbs || !deltas || !weights_sum || !depth ||
        !image)
        return NVSF_E_INVALID;
    int st = ensure_comp_attrs();
    if (st != NVSF_OK) return st;
    k_composite_rays<<<nvsf_div_up(n_alive, (uint32_t)kCompBlock), kCompBlock, kCompSmem,
                       (cudaStream_t)stream>>>(n_alive, n_step, T_thresh, rays_alive, rays_t,
                                               sigmas, rgbs, deltas, weights_sum, depth, image);
    return nvsf_launch_status();
}

}  // extern "C"
