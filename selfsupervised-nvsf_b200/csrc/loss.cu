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
