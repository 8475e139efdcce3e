// nvsf_b200 — Adam step over a flat fp32 parameter segment (sm_100a).
//
// Replaces torch.optim.Adam as the reference configures it (reference nvsf/scripts/main_nvsf.py:350-352:
// betas (0.9, 0.99), eps 1e-15, no weight decay, no amsgrad) for the parameters of the field, which
// live in ONE flat buffer per model replica (dist.py GradSync / optim.py): one streaming pass,
// 16 B read (p, g, m, v) + 12 B written (p, m, v) per parameter, 128-bit accesses, grid-stride
// over a multiple of the SM count.  `grad_scale` folds the 1/world_size of the gradient average
// and the inverse loss scale of a GradScaler into the same pass.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_adam(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
       float4* __restrict__ v, size_t n4, float step_size, float beta1, float beta2, float eps,
       float inv_sqrt_bc2, float grad_scale) {
    const float ob1 = 1.0f - beta1, ob2 = 1.0f - beta2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (size_t)gridDim.x * blockDim.x) {
        float4 pi = p[i], mi = m[i], vi = v[i];
        float4 gi = __ldcs(g + i);
        gi.x *= grad_scale; gi.y *= grad_scale; gi.z *= grad_scale; gi.w *= grad_scale;
#define NVSF_ADAM1(c)                                                        \
        mi.c = beta1 * mi.c + ob1 * gi.c;                                    \
        vi.c = beta2 * vi.c + ob2 * gi.c * gi.c;                             \
        pi.c -= step_size * (mi.c / (sqrtf(vi.c) * inv_sqrt_bc2 + eps));
        NVSF_ADAM1(x) NVSF_ADAM1(y) NVSF_ADAM1(z) NVSF_ADAM1(w)
#undef NVSF_ADAM1
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

__global__ void k_adam_tail(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t begin, size_t n, float step_size, float beta1,
                            float beta2, float eps, float inv_sqrt_bc2, float grad_scale) {
    const size_t i = begin + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i] * grad_scale;
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    p[i] -= step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
    m[i] = mi; v[i] = vi;
}

}  // namespace

extern "C" int nvsf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                              size_t n, float lr, float beta1, float beta2, float eps, uint32_t step,
                              float grad_scale, void* stream) {
    if (n == 0) return NVSF_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq || step == 0) return NVSF_E_INVALID;
    if ((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0)
        return NVSF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    // torch.optim.Adam: step_size = lr / (1 - beta1^t), denom = sqrt(v) / sqrt(1 - beta2^t) + eps
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    const size_t n4 = n / 4;
    if (n4) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const unsigned blocks = (unsigned)std::min<size_t>(nvsf_div_up(n4, (size_t)256), (size_t)sms * 8);
        k_adam<<<blocks, 256, 0, s>>>(reinterpret_cast<float4*>(params), reinterpret_cast<const float4*>(grads),
                                      reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq),
                                      n4, step_size, beta1, beta2, eps, inv_sqrt_bc2, grad_scale);
    }
    if (n4 * 4 < n)
        k_adam_tail<<<1, 4, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n4 * 4, n, step_size, beta1, beta2,
                                    eps, inv_sqrt_bc2, grad_scale);
    return nvsf_launch_status();
}
