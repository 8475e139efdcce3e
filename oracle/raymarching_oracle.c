/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (oracle) of the reference's ray-marching kernels,
 * reference nvsf/nerf/raymarching/src/raymarching.cu.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product path never does.
 *
 * Parity status: PINNED for Part 1.  On the GPU box the `-m gpu` tests compare
 * this oracle, the CUDA product and the reference's own extension rebuilt for
 * sm_100a (oracle/_ref/_raymarching_ref.so, see oracle/build_ref.sh) on the same
 * inputs; tests/golden/ holds vectors produced by that reference build.
 *
 * Arithmetic: one scalar loop per ray, same operation order as the reference.
 * Where nvcc/ptxas fuse a multiply-add in the reference build (default
 * -fmad=true; read from the sm_100a SASS) the oracle calls fmaf(); everything
 * else is plain IEEE single precision (compile with -ffp-contract=off).
 * `__expf` is restated as exp2f(x * log2e), so compositing agrees to rounding
 * (1e-5 relative), not bit for bit; sample counts / offsets / positions are
 * bit-exact.
 *
 * Ordering: the reference assigns `rays` rows and sample offsets with
 * atomicAdd (raymarching.cu:445-446), i.e. in scheduling order.  The oracle
 * processes rays in ray-id order, which is the canonical order the product
 * uses and the one every comparison sorts the reference output into.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

/* raymarching.cu:51-60 */
static inline int mip_from_pos(float x, float y, float z, float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int exponent;
    frexpf(mx, &exponent);
    if (!isfinite(mx)) exponent = 0; /* CUDA frexpf reports 0 for inf/nan */
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}

/* raymarching.cu:62-69 */
static inline int mip_from_dt(float dt, float H, float max_cascade) {
    const float mx = (float)((double)(dt * H) * 0.5);
    int exponent;
    frexpf(mx, &exponent);
    if (!isfinite(mx)) exponent = 0;
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}

/* raymarching.cu:71-77 */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
/* raymarching.cu:79-86 */
static inline uint32_t morton3D_1(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
/* raymarching.cu:88-95 */
static inline uint32_t morton3D_invert_1(uint32_t x) {
    x = x & 0x49249249;
    x = (x | (x >> 2)) & 0xc30c30c3;
    x = (x | (x >> 4)) & 0x0f00f00f;
    x = (x | (x >> 8)) & 0xff0000ff;
    x = (x | (x >> 16)) & 0x0000ffff;
    return x;
}

/* raymarching.cu:105-157 */
ORACLE_API void oracle_near_far_from_aabb(const float* rays_o, const float* rays_d,
                                          const float* aabb, uint32_t N, float min_near,
                                          float* nears, float* fars) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; ++n) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float rdx = 1.0f / rays_d[n * 3], rdy = 1.0f / rays_d[n * 3 + 1],
                    rdz = 1.0f / rays_d[n * 3 + 2];
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx;
        if (near > far) { float c = near; near = far; far = c; }
        float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
        if (near_y > far_y) { float c = near_y; near_y = far_y; far_y = c; }
        if (near > far_y || near_y > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
        if (near_z > far_z) { float c = near_z; near_z = far_z; far_z = c; }
        if (near > far_z || near_z > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_z > near) near = near_z;
        if (far_z < far) far = far_z;
        if (near < min_near) near = min_near;
        nears[n] = near;
        fars[n] = far;
    }
}

/* raymarching.cu:183-217 */
ORACLE_API void oracle_sph_from_ray(const float* rays_o, const float* rays_d, float radius,
                                    uint32_t N, float* coords) {
    const float RPI = 0.3183098861837907f;
    for (uint32_t n = 0; n < N; ++n) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float A = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
        const float B = fmaf(oz, dz, fmaf(ox, dx, oy * dy));
        const float C = fmaf(-radius, radius, fmaf(oz, oz, fmaf(ox, ox, oy * oy)));
        const float t = (-B + sqrtf(fmaf(B, B, -(A * C)))) / A;
        const float x = fmaf(dx, t, ox), y = fmaf(dy, t, oy), z = fmaf(dz, t, oz);
        const float theta = atan2f(sqrtf(fmaf(x, x, z * z)), y);
        const float phi = atan2f(z, x);
        coords[n * 2] = fmaf(2.0f * theta, RPI, -1.0f);
        coords[n * 2 + 1] = phi * RPI;
    }
}

/* raymarching.cu:237-247 */
ORACLE_API void oracle_morton3D(const int32_t* coords, uint32_t N, int32_t* indices) {
    for (uint32_t n = 0; n < N; ++n)
        indices[n] = (int32_t)morton3D_1((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1],
                                         (uint32_t)coords[n * 3 + 2]);
}

/* raymarching.cu:257-272 (note: `ind >> k` shifts a signed int) */
ORACLE_API void oracle_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords) {
    for (uint32_t n = 0; n < N; ++n) {
        const int32_t ind = indices[n];
        coords[n * 3] = (int32_t)morton3D_invert_1((uint32_t)(ind >> 0));
        coords[n * 3 + 1] = (int32_t)morton3D_invert_1((uint32_t)(ind >> 1));
        coords[n * 3 + 2] = (int32_t)morton3D_invert_1((uint32_t)(ind >> 2));
    }
}

/* raymarching.cu:287-306 */
ORACLE_API void oracle_packbits(const float* grid, uint32_t N, float density_thresh,
                                uint8_t* bitfield) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; ++n) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; ++i)
            bits |= (grid[n * 8 + i] > density_thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* ------------------------------------------------------------------------- */
/* the occupancy DDA, raymarching.cu:359-439                                   */
/* ------------------------------------------------------------------------- */
typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, dt_gamma, dt_min, dt_max, rH, H3, Hf, Cf;
    uint32_t H;
    const uint8_t* grid;
} march_t;

static inline void march_init(march_t* m, const float* o, const float* d, const uint8_t* grid,
                              float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                              uint32_t H) {
    m->ox = o[0]; m->oy = o[1]; m->oz = o[2];
    m->dx = d[0]; m->dy = d[1]; m->dz = d[2];
    m->rdx = 1.0f / m->dx; m->rdy = 1.0f / m->dy; m->rdz = 1.0f / m->dz;
    m->bound = bound; m->dt_gamma = dt_gamma;
    m->Hf = (float)H; m->Cf = (float)C; m->H = H;
    m->rH = 1.0f / (float)H;
    m->H3 = (float)(H * H * H);
    const float two_sqrt3 = 2 * 1.7320508075688772f;
    m->dt_min = two_sqrt3 / (float)max_steps;
    m->dt_max = two_sqrt3 * (float)(1 << (C - 1)) / (float)H;
    m->grid = grid;
}

typedef struct { float x, y, z, dt, tt; int occ; } probe_t;

static inline probe_t march_probe(const march_t* m, float t) {
    probe_t q;
    const float bound = m->bound;
    q.x = clampf(fmaf(m->dx, t, m->ox), -bound, bound);
    q.y = clampf(fmaf(m->dy, t, m->oy), -bound, bound);
    q.z = clampf(fmaf(m->dz, t, m->oz), -bound, bound);
    q.dt = clampf(t * m->dt_gamma, m->dt_min, m->dt_max);
    const int la = mip_from_pos(q.x, q.y, q.z, m->Cf);
    const int lb = mip_from_dt(q.dt, m->Hf, m->Cf);
    const int level = la > lb ? la : lb;
    const float mip_bound = fminf(scalbnf(1.0f, level), bound);
    const float mip_rbound = 1.0f / mip_bound;
    const float Hm1 = (float)(m->H - 1);
    /* 0.5 * (x * mip_rbound + 1) * H : fused multiply-add in float, product in double */
    const int nx = (int)clampf((float)(0.5 * (double)fmaf(q.x, mip_rbound, 1.0f) * (double)m->H), 0.0f, Hm1);
    const int ny = (int)clampf((float)(0.5 * (double)fmaf(q.y, mip_rbound, 1.0f) * (double)m->H), 0.0f, Hm1);
    const int nz = (int)clampf((float)(0.5 * (double)fmaf(q.z, mip_rbound, 1.0f) * (double)m->H), 0.0f, Hm1);
    /* index = level * H3 + morton, evaluated in float with one rounding (FFMA) */
    const uint32_t index =
        (uint32_t)fmaf(m->H3, (float)level, (float)morton3D_1((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    q.occ = (m->grid[index / 8] & (1 << (index % 8))) != 0;
    q.tt = t;
    if (!q.occ) {
        const float sx = copysignf(1.0f, m->dx), sy = copysignf(1.0f, m->dy), sz = copysignf(1.0f, m->dz);
        const float tx = fmaf(mip_bound, fmaf((fmaf(sx, 0.5f, (float)nx + 0.5f)) * m->rH, 2.0f, -1.0f), -q.x) * m->rdx;
        const float ty = fmaf(mip_bound, fmaf((fmaf(sy, 0.5f, (float)ny + 0.5f)) * m->rH, 2.0f, -1.0f), -q.y) * m->rdy;
        const float tz = fmaf(mip_bound, fmaf((fmaf(sz, 0.5f, (float)nz + 0.5f)) * m->rH, 2.0f, -1.0f), -q.z) * m->rdz;
        q.tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    }
    return q;
}

static inline float march_skip(const march_t* m, float t, float tt) {
    do {
        t += clampf(t * m->dt_gamma, m->dt_min, m->dt_max);
    } while (t < tt);
    return t;
}

/* raymarching.cu:332-534.  counter is accumulated, rays rows in ray-id order. */
ORACLE_API void oracle_march_rays_train(const float* rays_o, const float* rays_d,
                                        const uint8_t* grid, float bound, float dt_gamma,
                                        uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                        uint32_t M, const float* nears, const float* fars,
                                        float* xyzs, float* dirs, float* deltas, int32_t* rays,
                                        int32_t* counter, const float* noises) {
    /* pass 1 (parallel): counts */
    uint32_t* counts = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(N ? N : 1));
    uint32_t* offsets = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(N ? N : 1));
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)N; ++n) {
        march_t m;
        march_init(&m, rays_o + n * 3, rays_d + n * 3, grid, bound, dt_gamma, max_steps, C, H);
        const float near = nears[n], far = fars[n];
        float t = fmaf(noises[n], clampf(near * dt_gamma, m.dt_min, m.dt_max), near);
        uint32_t num_steps = 0;
        while (t < far && num_steps < max_steps) {
            const probe_t q = march_probe(&m, t);
            if (q.occ) { num_steps++; t += q.dt; }
            else t = march_skip(&m, t, q.tt);
        }
        counts[n] = num_steps;
    }
    /* atomicAdd(counter, num_steps) / atomicAdd(counter+1, 1), taken in ray-id order */
    const uint32_t ray_base = (uint32_t)counter[1];
    uint32_t point = (uint32_t)counter[0];
    for (uint32_t n = 0; n < N; ++n) {
        const uint32_t row = ray_base + n;
        offsets[n] = point;
        if (row < N) { /* the reference would write out of bounds otherwise */
            rays[row * 3] = (int32_t)n;
            rays[row * 3 + 1] = (int32_t)point;
            rays[row * 3 + 2] = (int32_t)counts[n];
        }
        point += counts[n];
    }
    counter[0] = (int32_t)point;
    counter[1] = (int32_t)(ray_base + N);

    /* pass 2 (parallel): emit */
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)N; ++n) {
        const uint32_t point_index = offsets[n];
        const uint32_t num_steps = counts[n];
        if (num_steps == 0) continue;
        if (point_index + num_steps > M) continue;
        march_t m;
        march_init(&m, rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, grid, bound, dt_gamma, max_steps, C, H);
        const float near = nears[n], far = fars[n];
        float t = fmaf(noises[n], clampf(near * dt_gamma, m.dt_min, m.dt_max), near);
        float last_t = t;
        float* px = xyzs + (size_t)point_index * 3;
        float* pd = dirs + (size_t)point_index * 3;
        float* pl = deltas + (size_t)point_index * 2;
        uint32_t step = 0;
        while (t < far && step < num_steps) {
            const probe_t q = march_probe(&m, t);
            if (q.occ) {
                px[0] = q.x; px[1] = q.y; px[2] = q.z;
                pd[0] = m.dx; pd[1] = m.dy; pd[2] = m.dz;
                t += q.dt;
                pl[0] = q.dt;
                pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2;
                step++;
            } else {
                t = march_skip(&m, t, q.tt);
            }
        }
    }
    free(counts);
    free(offsets);
}

static inline float fast_alpha(float sigma, float delta) {
    /* 1 - __expf(-sigma*delta);  __expf(x) = ex2.approx(x * log2e) */
    return 1.0f - exp2f((sigma * delta) * -1.4426950216293334961f);
}

/* raymarching.cu:578-655 */
ORACLE_API void oracle_composite_rays_train_forward(const float* sigmas, const float* rgbs,
                                                    const float* deltas, const int32_t* rays,
                                                    uint32_t M, uint32_t N, float T_thresh,
                                                    float* weights_sum, float* depth,
                                                    float* image) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)N; ++n) {
        const uint32_t index = (uint32_t)rays[n * 3];
        const uint32_t offset = (uint32_t)rays[n * 3 + 1];
        const uint32_t num_steps = (uint32_t)rays[n * 3 + 2];
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
        if (!(num_steps == 0 || offset + num_steps > M)) {
            const float* ps = sigmas + offset;
            const float* pr = rgbs + (size_t)offset * 3;
            const float* pd = deltas + (size_t)offset * 2;
            for (uint32_t step = 0; step < num_steps; ++step) {
                const float alpha = fast_alpha(ps[step], pd[2 * step]);
                const float weight = alpha * T;
                r = fmaf(weight, pr[3 * step], r);
                g = fmaf(weight, pr[3 * step + 1], g);
                b = fmaf(weight, pr[3 * step + 2], b);
                t += pd[2 * step + 1];
                d = fmaf(weight, t, d);
                ws += weight;
                T *= 1.0f - alpha;
                if (T < T_thresh) break;
            }
        }
        weights_sum[index] = ws;
        depth[index] = d;
        image[index * 3] = r;
        image[index * 3 + 1] = g;
        image[index * 3 + 2] = b;
    }
}

/* raymarching.cu:691-772 */
ORACLE_API void oracle_composite_rays_train_backward(
    const float* grad_weights_sum, const float* grad_image, const float* sigmas,
    const float* rgbs, const float* deltas, const int32_t* rays, const float* weights_sum,
    const float* image, uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas,
    float* grad_rgbs) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)N; ++n) {
        const uint32_t index = (uint32_t)rays[n * 3];
        const uint32_t offset = (uint32_t)rays[n * 3 + 1];
        const uint32_t num_steps = (uint32_t)rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps > M) continue;
        const float gi0 = grad_image[index * 3], gi1 = grad_image[index * 3 + 1],
                    gi2 = grad_image[index * 3 + 2];
        const float r_final = image[index * 3], g_final = image[index * 3 + 1],
                    b_final = image[index * 3 + 2];
        const float gws = grad_weights_sum[index] * (1 - weights_sum[index]);
        const float* ps = sigmas + offset;
        const float* pr = rgbs + (size_t)offset * 3;
        const float* pd = deltas + (size_t)offset * 2;
        float* gs = grad_sigmas + offset;
        float* gr = grad_rgbs + (size_t)offset * 3;
        float T = 1.0f, r = 0, g = 0, b = 0;
        for (uint32_t step = 0; step < num_steps; ++step) {
            const float alpha = fast_alpha(ps[step], pd[2 * step]);
            const float weight = alpha * T;
            r = fmaf(weight, pr[3 * step], r);
            g = fmaf(weight, pr[3 * step + 1], g);
            b = fmaf(weight, pr[3 * step + 2], b);
            T *= 1.0f - alpha;
            gr[3 * step] = gi0 * weight;
            gr[3 * step + 1] = gi1 * weight;
            gr[3 * step + 2] = gi2 * weight;
            const float t0 = fmaf(pr[3 * step], T, -(r_final - r));
            const float t1 = fmaf(pr[3 * step + 1], T, -(g_final - g));
            const float t2 = fmaf(pr[3 * step + 2], T, -(b_final - b));
            gs[step] = pd[2 * step] * (gws + fmaf(gi2, t2, fmaf(gi0, t0, gi1 * t1)));
            if (T < T_thresh) break;
        }
    }
}

/* raymarching.cu:809-928 */
ORACLE_API void oracle_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive,
                                  const float* rays_t, const float* rays_o, const float* rays_d,
                                  float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                                  uint32_t H, const uint8_t* grid, const float* nears,
                                  const float* fars, float* xyzs, float* dirs, float* deltas,
                                  const float* noises) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)n_alive; ++n) {
        const int32_t index = rays_alive[n];
        march_t m;
        march_init(&m, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, grid, bound,
                   dt_gamma, max_steps, C, H);
        float* px = xyzs + (size_t)n * n_step * 3;
        float* pd = dirs + (size_t)n * n_step * 3;
        float* pl = deltas + (size_t)n * n_step * 2;
        float t = rays_t[index];
        const float far = fars[index];
        t = fmaf(noises[n], clampf(t * dt_gamma, m.dt_min, m.dt_max), t);
        float last_t = t;
        uint32_t step = 0;
        while (t < far && step < n_step) {
            const probe_t q = march_probe(&m, t);
            if (q.occ) {
                px[0] = q.x; px[1] = q.y; px[2] = q.z;
                pd[0] = m.dx; pd[1] = m.dy; pd[2] = m.dz;
                t += q.dt;
                pl[0] = q.dt;
                pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2;
                step++;
            } else {
                t = march_skip(&m, t, q.tt);
            }
        }
    }
    (void)nears;
}

/* raymarching.cu:967-1053 */
ORACLE_API void oracle_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh,
                                      int32_t* rays_alive, float* rays_t, const float* sigmas,
                                      const float* rgbs, const float* deltas,
                                      float* weights_sum, float* depth, float* image) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)n_alive; ++n) {
        const int32_t index = rays_alive[n];
        const float* ps = sigmas + (size_t)n * n_step;
        const float* pr = rgbs + (size_t)n * n_step * 3;
        const float* pd = deltas + (size_t)n * n_step * 2;
        float t = rays_t[index];
        float weight_sum = weights_sum[index], d = depth[index];
        float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
        uint32_t step = 0;
        while (step < n_step) {
            if (pd[2 * step] == 0) break;
            const float alpha = fast_alpha(ps[step], pd[2 * step]);
            const float T = 1 - weight_sum;
            const float weight = alpha * T;
            weight_sum += weight;
            t += pd[2 * step + 1];
            d = fmaf(weight, t, d);
            r = fmaf(weight, pr[3 * step], r);
            g = fmaf(weight, pr[3 * step + 1], g);
            b = fmaf(weight, pr[3 * step + 2], b);
            if (T < T_thresh) break;
            step++;
        }
        if (step < n_step) rays_alive[n] = -1;
        else rays_t[index] = t;
        weights_sum[index] = weight_sum;
        depth[index] = d;
        image[index * 3] = r;
        image[index * 3 + 1] = g;
        image[index * 3 + 2] = b;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Chamfer distance — restatement of nvsf/nerf/chamfer3D/chamfer3D.cu (TEST INFRASTRUCTURE).
 *
 * NmDistanceKernel (:9-150): for every point j of cloud A the squared distance to, and the index
 * of, its nearest point in cloud B; targets are scanned in ascending k with a strict `<` inside a
 * 512-point block (:43,:123) and a strict `>` across blocks (:129), i.e. the FIRST minimum wins.
 * The distance is d = x2*x2 + y2*y2 + z2*z2 with x2 = b_k - a_j (:38-42), which the reference
 * build contracts to FFMA(z2, z2, FFMA(x2, x2, FMUL(y2, y2))) (sm_100a SASS of the unmodified
 * kernel, oracle/_ref/chamfer_3D_ref.so: the y term is the plain product; confirmed against the
 * extension's outputs, tests/golden/chamfer_ref_sm100a.npz) — spelled out with fmaf() here.
 * NmDistanceGradKernel (:151-178): g = 2 grad_dist[j]; grad_a[j] += g (a_j - b_idx),
 * grad_b[idx] -= g (a_j - b_idx); chamfer_cuda_backward (:183-230) runs it in both directions.
 * ---------------------------------------------------------------------------------------------- */
ORACLE_API void oracle_chamfer_nn(const float* a, uint32_t n, const float* b, uint32_t m,
                                  uint32_t batch, float* dist, int32_t* idx) {
    for (uint32_t i = 0; i < batch; ++i) {
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < (int64_t)n; ++j) {
            const float* p = a + ((size_t)i * n + j) * 3;
            float best = 0.f;
            int32_t best_i = 0;
            for (uint32_t k = 0; k < m; ++k) {
                const float* q = b + ((size_t)i * m + k) * 3;
                const float x2 = q[0] - p[0], y2 = q[1] - p[1], z2 = q[2] - p[2];
                const float d = fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
                if (k == 0 || d < best) {
                    best = d;
                    best_i = (int32_t)k;
                }
            }
            dist[(size_t)i * n + j] = best;
            idx[(size_t)i * n + j] = best_i;
        }
    }
}

ORACLE_API void oracle_chamfer_grad(const float* a, uint32_t n, const float* b, uint32_t m,
                                    uint32_t batch, const float* grad_dist, const int32_t* idx,
                                    float* grad_a, float* grad_b) {
    for (uint32_t i = 0; i < batch; ++i) {
        for (uint32_t j = 0; j < n; ++j) {
            const size_t ja = (size_t)i * n + j, jb = (size_t)i * m + (uint32_t)idx[ja];
            const float g = grad_dist[ja] * 2.f;
            for (int c = 0; c < 3; ++c) {
                const float v = g * (a[ja * 3 + c] - b[jb * 3 + c]);
                grad_a[ja * 3 + c] += v;
                grad_b[jb * 3 + c] += -v;
            }
        }
    }
}
