// nvsf_b200 — loss head on the composited outputs (sm_100a).
//
// Replaces the per-ray supervision of Trainer.train_step (reference nvsf/nerf/trainer.py:188-219:
// raydrop mask, label smoothing, depth / raydrop / intensity criteria weighted by alpha_d / alpha_r /
// alpha_i; :503: alpha_rgb * criterion["rgb"](pred_rgb, gt_rgb)) with the element-wise criteria
// of main_nvsf.py:205-212 (reduction="none"): one streaming kernel writes the per-ray (per-element)
// loss the trainer keeps for its error map AND the derivative of that loss with respect to the
// renderer's outputs, so the backward pass of the renderer starts from these buffers without the
// ~20 element-wise ATen launches autograd would replay.  48 B/ray LiDAR, 48 B/ray camera: pure
// HBM streams, launch-latency bound at 4096 rays.
#include <algorithm>

#include "common.cuh"

namespace {

// criteria of main_nvsf.py:205-212 (element-wise)
__device__ __forceinline__ void crit(int kind, float param, float p, float t, float& l, float& g) {
    const float d = p - t;
    if (kind == NVSF_LOSS_L1) {                       // torch.nn.L1Loss
        l = fabsf(d);
        g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    } else if (kind == NVSF_LOSS_MSE) {               // torch.nn.MSELoss
        l = d * d;
        g = 2.f * d;
    } else if (kind == NVSF_LOSS_SMOOTHL1) {          // SmoothL1Loss(beta = param)
        const float a = fabsf(d);
        if (a < param) { l = 0.5f * d * d / param; g = d / param; }
        else { l = a - 0.5f * param; g = d > 0.f ? 1.f : -1.f; }
    } else {                                          // HuberLoss(delta = param)
        const float a = fabsf(d);
        if (a <= param) { l = 0.5f * d * d; g = d; }
        else { l = param * (a - 0.5f * param); g = d > 0.f ? param : -param; }
    }
}

__global__ void __launch_bounds__(256)
k_loss_lidar(const float* __restrict__ depth, const float* __restrict__ image,
             const float* __restrict__ gt, uint32_t n, nvsf_lidar_loss_cfg_t c,
             float* __restrict__ loss, float* __restrict__ g_depth, float* __restrict__ g_image) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float m = __ldg(gt + 3 * (size_t)i);               // gt_raydrop (trainer.py:188)
    const float gi = __ldg(gt + 3 * (size_t)i + 1) * m;      // gt_intensity (:189)
    const float gd = __ldg(gt + 3 * (size_t)i + 2) * m;      // gt_depth (:190)
    const float2 im = __ldg(reinterpret_cast<const float2*>(image) + i);
    const float pr = im.x;                                   // pred_raydrop (:206)
    const float pi = im.y * m;                               // pred_intensity (:205)
    const float pd = __ldg(depth + i) * m;                   // pred_depth (:206)
    const float gs = fminf(fmaxf(m, c.smooth), 1.f - c.smooth);  // clamp(smooth, 1 - smooth) (:211-213)
    float ld, gdd, lr, gr, li, gii;
    crit(c.depth_kind, c.depth_param, pd, gd, ld, gdd);
    crit(c.raydrop_kind, c.raydrop_param, pr, gs, lr, gr);
    crit(c.intensity_kind, c.intensity_param, pi, gi, li, gii);
    loss[i] = (c.alpha_d * ld + c.alpha_r * lr) + c.alpha_i * li;   // lidar_loss (:216-219)
    g_depth[i] = c.alpha_d * gdd * m;
    reinterpret_cast<float2*>(g_image)[i] = make_float2(c.alpha_r * gr, c.alpha_i * gii * m);
}

__global__ void __launch_bounds__(256)
k_loss_elem(const float* __restrict__ pred, const float* __restrict__ gt, size_t n, int kind,
            float param, float alpha, float* __restrict__ loss, float* __restrict__ g_pred) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float l, g;
    crit(kind, param, __ldg(pred + i), __ldg(gt + i), l, g);
    loss[i] = alpha * l;
    g_pred[i] = alpha * g;
}

bool kind_ok(int k) { return k >= NVSF_LOSS_L1 && k <= NVSF_LOSS_HUBER; }

}  // namespace

extern "C" {

int nvsf_loss_lidar(const float* depth, const float* image, const float* gt, uint32_t n,
                    const nvsf_lidar_loss_cfg_t* cfg, float* loss, float* g_depth, float* g_image,
                    void* stream) {
    if (n == 0) return NVSF_OK;
    if (!depth || !image || !gt || !cfg || !loss || !g_depth || !g_image) return NVSF_E_INVALID;
    if (!kind_ok(cfg->depth_kind) || !kind_ok(cfg->raydrop_kind) || !kind_ok(cfg->intensity_kind))
        return NVSF_E_INVALID;
    k_loss_lidar<<<nvsf_div_up(n, 256u), 256, 0, (cudaStream_t)stream>>>(depth, image, gt, n, *cfg, loss,
                                                                       g_depth, g_image);
    return nvsf_launch_status();
}

int nvsf_loss_elementwise(const float* pred, const float* gt, size_t n, int kind, float param,
                          float alpha, float* loss, float* g_pred, void* stream) {
    if (n == 0) return NVSF_OK;
    if (!pred || !gt || !loss || !g_pred || !kind_ok(kind)) return NVSF_E_INVALID;
    k_loss_elem<<<(unsigned)nvsf_div_up(n, (size_t)256), 256, 0, (cudaStream_t)stream>>>(pred, gt, n, kind, param,
                                                                                          alpha, loss, g_pred);
    return nvsf_launch_status();
}

}  // extern "C"

// ================================================================================================
// The remaining loss terms of Trainer.train_step on the LiDAR outputs (reference nvsf/nerf/trainer.py):
//   :276-296  line-of-sight loss of Urban Radiance Fields on (weights, z_vals)      -> k_los_*
//   :297-462  structural regularisation of depth patches: Sobel / finite-difference gradients,
//             edge-aware / smoothness / TV terms and the masked gradient loss        -> k_patch_*
// Each kernel writes the loss AND its derivative with respect to the renderer's outputs (weights,
// depth), like k_loss_lidar: the renderer's backward starts from those buffers.
// ================================================================================================
namespace {

// ---- URF line-of-sight loss ------------------------------------------------------------------------
// pass 1: max over all elements of distr = N(distance; sigma) (trainer.py:291-292) and the number of
// rays with gt_depth > 0 (:283)
__global__ void __launch_bounds__(256)
k_los_stats(const float* __restrict__ z_vals, const float* __restrict__ gt_depth, uint32_t R, uint32_t T,
            float eps, unsigned* __restrict__ stats /* [0] = max bits, [1] = count */) {
    const size_t n = (size_t)R * T;
    const float sigma = eps / 3.f;
    const float peak = 1.0f / (sigma * 2.5066282746310002f);
    float mx = 0.f;
    unsigned cnt = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(i / T);
        const float gd = __ldg(gt_depth + r), z = __ldg(z_vals + i);
        const bool near = (z > gd - eps) && (z < gd + eps);
        const float dist = near ? z - gd : 0.f;
        mx = fmaxf(mx, peak * expf(-(dist * dist / (2.f * sigma * sigma))));
        if (i - (size_t)r * T == 0 && gd > 0.f) ++cnt;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(stats, __float_as_uint(mx));
        if (cnt) atomicAdd(stats + 1, cnt);
    }
}

// pass 2: los_loss = 0.1 (sum (mask_empty w)^2 + sum (mask_near w - distr)^2) / count  (:284-296)
__global__ void __launch_bounds__(256)
k_los_loss(const float* __restrict__ weights, const float* __restrict__ z_vals,
           const float* __restrict__ gt_depth, uint32_t R, uint32_t T, float eps,
           const unsigned* __restrict__ stats, double* __restrict__ loss_acc, float* __restrict__ g_weights) {
    const size_t n = (size_t)R * T;
    const float sigma = eps / 3.f;
    const float peak = 1.0f / (sigma * 2.5066282746310002f);
    const float inv_max = 1.0f / __uint_as_float(__ldg(stats));
    const float k = 0.1f / (float)__ldg(stats + 1);
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(i / T);
        const float gd = __ldg(gt_depth + r), z = __ldg(z_vals + i), w = __ldg(weights + i);
        const bool empty = (z < gd - eps) || (z > gd + eps);
        const bool near = (z > gd - eps) && (z < gd + eps);
        float l = 0.f, g = 0.f;
        if (empty) { l = w * w; g = 2.f * w; }
        const float dist = near ? z - gd : 0.f;
        const float distr = near ? (peak * expf(-(dist * dist / (2.f * sigma * sigma)))) * inv_max : 0.f;
        const float e = (near ? w : 0.f) - distr;
        l += e * e;
        if (near) g += 2.f * e;
        acc += (double)l;
        g_weights[i] = k * g;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss_acc, acc * (double)k);
}

__global__ void k_los_finish(const double* __restrict__ acc, float* __restrict__ loss) { *loss = (float)*acc; }

// ---- ground-truth side of the gradient loss: masks from the second differences of the range image --
// trainer.py:392-428: gx = (D[:, :-1] - D[:, 1:]) / scale (last column repeated), gxx = |gx[:, :-1]| -
// |gx[:, 1:]| (last repeated), same along rows; a patch pixel (i, j) reads gxx at (row of pixel (i, 0),
// column of pixel (i, j)) and gyy at (row of pixel (i, j), column of pixel (0, j)); mask = |.| < thresh.
__device__ __forceinline__ float pano_gx(const float* D, uint32_t W, uint32_t r, uint32_t c, float inv_scale) {
    const uint32_t cc = c < W - 1 ? c : W - 2;
    return (D[(size_t)r * W + cc] - D[(size_t)r * W + cc + 1]) * inv_scale;
}
__device__ __forceinline__ float pano_gy(const float* D, uint32_t H, uint32_t W, uint32_t r, uint32_t c,
                                         float inv_scale) {
    const uint32_t rr = r < H - 1 ? r : H - 2;
    return (D[(size_t)rr * W + c] - D[(size_t)(rr + 1) * W + c]) * inv_scale;
}
__global__ void __launch_bounds__(256)
k_patch_masks(const float* __restrict__ pano, uint32_t H, uint32_t W, const int64_t* __restrict__ inds,
              uint32_t P, uint32_t h, uint32_t w, float scale, float thresh, float* __restrict__ mask_x,
              float* __restrict__ mask_y) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * h * w) return;
    const uint32_t p = i / (h * w), rem = i - p * h * w, pi = rem / w, pj = rem - pi * w;
    const float inv = 1.0f / scale;
    const int64_t me = inds[i], row0 = inds[(size_t)p * h * w + pi * w], col0 = inds[(size_t)p * h * w + pj];
    {
        const uint32_t r = (uint32_t)(row0 / W), c = (uint32_t)(me % W);
        const uint32_t cc = c < W - 1 ? c : W - 2;
        const float v = fabsf(pano_gx(pano, W, r, cc, inv)) - fabsf(pano_gx(pano, W, r, cc + 1, inv));
        mask_x[i] = fabsf(v) < thresh ? 1.f : 0.f;
    }
    {
        const uint32_t r = (uint32_t)(me / W), c = (uint32_t)(col0 % W);
        const uint32_t rr = r < H - 1 ? r : H - 2;
        const float v = fabsf(pano_gy(pano, H, W, rr, c, inv)) - fabsf(pano_gy(pano, H, W, rr + 1, c, inv));
        mask_y[i] = fabsf(v) < thresh ? 1.f : 0.f;
    }
}

// ---- structural regularisation of one depth patch per CTA ------------------------------------------
__device__ __forceinline__ float patch_grad_x(const float* d, int h, int w, int y, int x, bool sobel) {
    if (sobel) {   // F.conv2d(d, [[-1,0,1],[-2,0,2],[-1,0,1]], padding=1)  (cross-correlation, zero padding)
        float s = 0.f;
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= h) continue;
            const float k = dy == 0 ? 2.f : 1.f;
            if (x + 1 < w) s += k * d[yy * w + x + 1];
            if (x - 1 >= 0) s -= k * d[yy * w + x - 1];
        }
        return s;
    }
    const int xx = x < w - 1 ? x : w - 2;   // d[:, :-1] - d[:, 1:], last column repeated (:333-334)
    return d[y * w + xx] - d[y * w + xx + 1];
}
__device__ __forceinline__ float patch_grad_y(const float* d, int h, int w, int y, int x, bool sobel) {
    if (sobel) {   // [[-1,-2,-1],[0,0,0],[1,2,1]]
        float s = 0.f;
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= w) continue;
            const float k = dx == 0 ? 2.f : 1.f;
            if (y + 1 < h) s += k * d[(y + 1) * w + xx];
            if (y - 1 >= 0) s -= k * d[(y - 1) * w + xx];
        }
        return s;
    }
    const int yy = y < h - 1 ? y : h - 2;
    return d[yy * w + x] - d[(yy + 1) * w + x];
}

__device__ __forceinline__ float block_sum(float v, float* red /* [32] */) {
    __syncthreads();
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    return s;
}

__global__ void __launch_bounds__(128)
k_patch_loss(const float* __restrict__ pred, const float* __restrict__ gt_depth, const float* __restrict__ gt_raydrop,
             const float* __restrict__ mask_x, const float* __restrict__ mask_y, uint32_t h, uint32_t w,
             nvsf_patch_loss_cfg_t c, float* __restrict__ loss_map, float* __restrict__ grad_loss,
             const float* __restrict__ g_map, const float* __restrict__ g_grad, float* __restrict__ g_pred) {
    // forward pass: loss_map / grad_loss (g_pred == NULL).  backward pass: g_pred = sum_i g_map[i] *
    // d loss_map[i] / d pred + g_grad[patch] * d grad_loss[patch] / d pred (loss_map / grad_loss == NULL).
    extern __shared__ float sm[];
    const int n = (int)(h * w), H = (int)h, W = (int)w;
    float* d = sm;            // pred / scale
    float* t = d + n;         // gt / scale
    float* Gx = t + n;        // dL/dgx
    float* Gy = Gx + n;       // dL/dgy
    float* red = Gy + n;      // [32]
    const size_t base = (size_t)blockIdx.x * n;
    const float inv = 1.0f / c.scale;
    const bool sobel = c.sobel != 0, gl = c.grad_loss != 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        d[i] = __ldg(pred + base + i) * inv;
        t[i] = gl ? __ldg(gt_depth + base + i) * inv : 0.f;
    }
    __syncthreads();
    // cosine criterion: per-patch dot products first
    float ab_x = 0.f, aa_x = 0.f, bb_x = 0.f, ab_y = 0.f, aa_y = 0.f, bb_y = 0.f;
    if (gl && c.grad_kind == NVSF_LOSS_COS) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int y = i / W, x = i - y * W;
            const float rd = __ldg(gt_raydrop + base + i);
            const float mx = rd * __ldg(mask_x + base + i), my = rd * __ldg(mask_y + base + i);
            const float ax = patch_grad_x(d, H, W, y, x, sobel) * mx, bx = patch_grad_x(t, H, W, y, x, sobel) * mx;
            const float ay = patch_grad_y(d, H, W, y, x, sobel) * my, by = patch_grad_y(t, H, W, y, x, sobel) * my;
            ab_x += ax * bx; aa_x += ax * ax; bb_x += bx * bx;
            ab_y += ay * by; aa_y += ay * ay; bb_y += by * by;
        }
        ab_x = block_sum(ab_x, red); aa_x = block_sum(aa_x, red); bb_x = block_sum(bb_x, red);
        ab_y = block_sum(ab_y, red); aa_y = block_sum(aa_y, red); bb_y = block_sum(bb_y, red);
    }
    // torch.nn.CosineSimilarity(dim=1, eps=1e-8): x1.x2 / max(|x1| |x2|, eps)
    const float nx = sqrtf(aa_x) * sqrtf(bb_x), ny = sqrtf(aa_y) * sqrtf(bb_y);
    const float cos_x = ab_x / fmaxf(nx, 1e-8f), cos_y = ab_y / fmaxf(ny, 1e-8f);
    float gsum = 0.f;
    const float gg = (g_pred && g_grad) ? __ldg(g_grad + blockIdx.x) : 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int y = i / W, x = i - y * W;
        const float gx = patch_grad_x(d, H, W, y, x, sobel), gy = patch_grad_y(d, H, W, y, x, sobel);
        const float sx = gx > 0.f ? 1.f : (gx < 0.f ? -1.f : 0.f), sy = gy > 0.f ? 1.f : (gy < 0.f ? -1.f : 0.f);
        float l = 0.f, dgx = 0.f, dgy = 0.f;   // element-wise terms: value and derivative
        float egx = 0.f, egy = 0.f;            // gradient-loss term: derivative
        if (c.grad_norm_smooth) {    // :338-341
            const float ex = expf(-fabsf(gx)), ey = expf(-fabsf(gy));
            l += c.alpha_grad_norm * (ex + ey);
            dgx -= c.alpha_grad_norm * sx * ex; dgy -= c.alpha_grad_norm * sy * ey;
        }
        if (c.spatial_smooth) {      // :343-346
            l += c.alpha_spatial * (gx * gx + gy * gy);
            dgx += 2.f * c.alpha_spatial * gx; dgy += 2.f * c.alpha_spatial * gy;
        }
        if (c.tv_loss) {             // :348-351
            l += c.alpha_tv * (fabsf(gx) + fabsf(gy));
            dgx += c.alpha_tv * sx; dgy += c.alpha_tv * sy;
        }
        if (gl) {                    // :354-462
            const float rd = __ldg(gt_raydrop + base + i);
            const float mx = rd * __ldg(mask_x + base + i), my = rd * __ldg(mask_y + base + i);
            const float tx = patch_grad_x(t, H, W, y, x, sobel), ty = patch_grad_y(t, H, W, y, x, sobel);
            if (c.grad_kind == NVSF_LOSS_COS) {
                // (1 - cos) expanded to every pixel of the patch, then .sum(): n * (1 - cos) per direction
                const float ax = gx * mx, bx = tx * mx, ay = gy * my, by = ty * my;
                if (nx > 1e-8f) egx -= c.alpha_grad * (float)n * (bx / nx - cos_x * ax / aa_x) * mx;
                if (ny > 1e-8f) egy -= c.alpha_grad * (float)n * (by / ny - cos_y * ay / aa_y) * my;
                gsum += c.alpha_grad * ((1.f - cos_x) + (1.f - cos_y));
            } else {
                float lx, ggx, ly, ggy;
                crit(c.grad_kind, c.grad_param, gx * mx, tx * mx, lx, ggx);
                crit(c.grad_kind, c.grad_param, gy * my, ty * my, ly, ggy);
                gsum += c.alpha_grad * (lx + ly);
                egx += c.alpha_grad * ggx * mx; egy += c.alpha_grad * ggy * my;
            }
        }
        if (loss_map) loss_map[base + i] = l;
        const float gm = (g_pred && g_map) ? __ldg(g_map + base + i) : 0.f;
        Gx[i] = gm * dgx + gg * egx; Gy[i] = gm * dgy + gg * egy;
    }
    gsum = block_sum(gsum, red);   // also orders the Gx / Gy writes before the gather below
    if (threadIdx.x == 0 && grad_loss) grad_loss[blockIdx.x] = gsum;
    if (!g_pred) return;
    // adjoint of the gradient stencils: dL/dd(y', x') = sum over the pixels whose gx / gy read d(y', x')
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int y = i / W, x = i - y * W;
        float g = 0.f;
        if (sobel) {
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = y - dy;   // gx(yy, .) reads d(yy + dy, .) = d(y, .)
                if (yy < 0 || yy >= H) continue;
                const float k = dy == 0 ? 2.f : 1.f;
                if (x - 1 >= 0) g += k * Gx[yy * W + x - 1];   // gx(yy, x-1) reads +k d(yy+dy, x)
                if (x + 1 < W) g -= k * Gx[yy * W + x + 1];    // gx(yy, x+1) reads -k d(yy+dy, x)
            }
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x - dx;
                if (xx < 0 || xx >= W) continue;
                const float k = dx == 0 ? 2.f : 1.f;
                if (y - 1 >= 0) g += k * Gy[(y - 1) * W + xx];
                if (y + 1 < H) g -= k * Gy[(y + 1) * W + xx];
            }
        } else {
            // gx(y, x) = d(y, x) - d(y, x+1) for x <= W-2; gx(y, W-1) = gx(y, W-2)
            auto gxe = [&](int xx) { return Gx[y * W + xx] + (xx == W - 2 ? Gx[y * W + W - 1] : 0.f); };
            if (x <= W - 2) g += gxe(x);
            if (x >= 1) g -= gxe(x - 1);
            auto gye = [&](int yy) { return Gy[yy * W + x] + (yy == H - 2 ? Gy[(H - 1) * W + x] : 0.f); };
            if (y <= H - 2) g += gye(y);
            if (y >= 1) g -= gye(y - 1);
        }
        g_pred[base + i] = g * inv;
    }
}

}  // namespace

extern "C" {

int nvsf_loss_los(const float* weights, const float* z_vals, const float* gt_depth, uint32_t R, uint32_t T,
                  float eps, float* loss, float* g_weights, void* workspace, size_t workspace_bytes,
                  void* stream) {
    if (R == 0 || T == 0) return NVSF_OK;
    if (!weights || !z_vals || !gt_depth || !loss || !g_weights || !workspace || !(eps > 0.f)) return NVSF_E_INVALID;
    if (workspace_bytes < 16) return NVSF_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(workspace, 0, 16, s);
    unsigned* stats = reinterpret_cast<unsigned*>(workspace);
    double* acc = reinterpret_cast<double*>(stats + 2);
    const size_t n = (size_t)R * T;
    const unsigned blocks = (unsigned)std::min<size_t>(nvsf_div_up(n, (size_t)256), (size_t)148 * 8);
    k_los_stats<<<blocks, 256, 0, s>>>(z_vals, gt_depth, R, T, eps, stats);
    k_los_loss<<<blocks, 256, 0, s>>>(weights, z_vals, gt_depth, R, T, eps, stats, acc, g_weights);
    k_los_finish<<<1, 1, 0, s>>>(acc, loss);
    return nvsf_launch_status();
}

int nvsf_patch_grad_masks(const float* pano_depth, uint32_t H, uint32_t W, const int64_t* rays_pano_inds,
                          uint32_t P, uint32_t h, uint32_t w, float scale, float thresh, float* mask_x,
                          float* mask_y, void* stream) {
    if (P == 0) return NVSF_OK;
    if (!pano_depth || !rays_pano_inds || !mask_x || !mask_y || H < 3 || W < 3 || h == 0 || w == 0 || !(scale > 0.f))
        return NVSF_E_INVALID;
    k_patch_masks<<<nvsf_div_up(P * h * w, 256u), 256, 0, (cudaStream_t)stream>>>(
        pano_depth, H, W, rays_pano_inds, P, h, w, scale, thresh, mask_x, mask_y);
    return nvsf_launch_status();
}

int nvsf_loss_patch(const float* pred_depth, const float* gt_depth, const float* gt_raydrop, const float* mask_x,
                    const float* mask_y, uint32_t P, uint32_t h, uint32_t w, const nvsf_patch_loss_cfg_t* cfg,
                    float* loss_map, float* grad_loss, const float* g_map, const float* g_grad, float* g_pred,
                    void* stream) {
    if (P == 0) return NVSF_OK;
    if (!pred_depth || !cfg || h < 2 || w < 2 || !(cfg->scale > 0.f)) return NVSF_E_INVALID;
    if (!g_pred && (!loss_map || !grad_loss)) return NVSF_E_INVALID;   /* forward pass: both outputs */
    if (g_pred && !g_map && !g_grad) return NVSF_E_INVALID;            /* backward pass: an incoming gradient */
    if (cfg->grad_loss && (!gt_depth || !gt_raydrop || !mask_x || !mask_y)) return NVSF_E_INVALID;
    if (cfg->grad_loss && !(kind_ok(cfg->grad_kind) || cfg->grad_kind == NVSF_LOSS_COS)) return NVSF_E_INVALID;
    const size_t smem = ((size_t)4 * h * w + 32) * sizeof(float);
    if (smem > 200 * 1024) return NVSF_E_INVALID;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_patch_loss, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    k_patch_loss<<<P, 128, smem, (cudaStream_t)stream>>>(pred_depth, gt_depth, gt_raydrop, mask_x, mask_y, h, w,
                                                        *cfg, loss_map, grad_loss, g_map, g_grad, g_pred);
    return nvsf_launch_status();
}

}  // extern "C"
