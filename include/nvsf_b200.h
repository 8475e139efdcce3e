/*
 * nvsf_b200.h — C-ABI of the B200-native NVSF ray-rendering hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain device pointers,
 * sizes and a CUDA stream handle (`void*` == cudaStream_t, NULL = legacy default
 * stream) and returns an int status (0 = ok, >0 = cudaError_t of the launch,
 * <0 = NVSF_E_*).  No torch types appear in any signature; the caller allocates
 * every buffer (the reference's ownership rule, raymarching.py:40-41,235-250).
 *
 * Part 1 replaces the ten functions of the reference's pybind module
 * `_raymarching` (reference nvsf/nerf/raymarching/src/bindings.cpp:7-20, C++
 * declarations nvsf/nerf/raymarching/src/raymarching.h:6-96).  Argument order
 * follows raymarching.h; only fp32 is supported ("scalar_t should always be
 * float in use", raymarching.cu:103).
 *
 * Part 2 (nvsf_field_*, nvsf_render_*) replaces the per-sample field
 * evaluation NeRFNetwork.density / .color (network_dynamic.py:213-332) with its
 * tcnn encodings / MLPs and the uniform-sample renderer NeRFRenderer.run
 * (renderer_dynamic.py:109-265).
 */
#ifndef NVSF_B200_H_
#define NVSF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NVSF_OK 0
#define NVSF_E_INVALID (-1)   /* bad argument (null pointer, size, unsupported config) */
#define NVSF_E_WORKSPACE (-2) /* workspace too small */

/* ABI version of this header; bumped on any signature change. */
#define NVSF_B200_ABI_VERSION 3
int nvsf_abi_version(void);
/* Human-readable text for a status returned by any nvsf_* call. */
const char* nvsf_status_string(int status);

/* ------------------------------------------------------------------------- */
/* Part 1 — raymarching operators                                             */
/* ------------------------------------------------------------------------- */

/* replaces near_far_from_aabb (raymarching.h:6-12, kernel raymarching.cu:105-157).
 * rays_o/rays_d [N,3], aabb [6] (device), nears/fars [N]. */
int nvsf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb,
                            uint32_t N, float min_near, float* nears, float* fars,
                            void* stream);

/* replaces sph_from_ray (raymarching.h:13-17, raymarching.cu:183-217). coords [N,2]. */
int nvsf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N,
                      float* coords, void* stream);

/* replaces morton3D (raymarching.h:18, raymarching.cu:237-253). coords [N,3] i32 -> indices [N] i32. */
int nvsf_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream);

/* replaces morton3D_invert (raymarching.h:19-21, raymarching.cu:257-280). */
int nvsf_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream);

/* replaces packbits (raymarching.h:22-25, raymarching.cu:287-320).
 * grid [N*8] f32 -> bitfield [N] u8, bit i of byte n = grid[8n+i] > density_thresh. */
int nvsf_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield,
                  void* stream);

/* Scratch bytes needed by nvsf_march_rays_train* for N rays. */
size_t nvsf_march_rays_train_workspace_bytes(uint32_t N);

/* replaces march_rays_train (raymarching.h:27-44, kernel raymarching.cu:332-534).
 *
 * Same inputs/outputs as the reference plus a caller-owned scratch buffer.
 * Differences, all inside what the reference leaves unspecified:
 *   - `rays` rows are written in ray-id order and sample offsets are the
 *     exclusive prefix sum of the per-ray counts in that order (the reference's
 *     order comes from atomicAdd and is scheduling dependent, raymarching.cu:445-454);
 *   - `counter` (int32[2]) is accumulated exactly like the reference: offsets
 *     start at counter[0], ray rows at counter[1]; on return
 *     counter[0] += sum(counts), counter[1] += N;
 *   - rays whose slice would end beyond M are dropped (row still written), as in
 *     raymarching.cu:457 — here deterministically the trailing ones.
 * xyzs/dirs [M,3], deltas [M,2] must be zero-filled by the caller where the
 * reference requires zeros (rows not covered by any ray). */
int nvsf_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid,
                          float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                          uint32_t C, uint32_t H, uint32_t M, const float* nears,
                          const float* fars, float* xyzs, float* dirs, float* deltas,
                          int32_t* rays, int32_t* counter, const float* noises,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Two-phase form of the same operator (lets the host size the sample buffers
 * from counter[0] instead of allocating N*max_steps rows):
 *   phase 1 `_count`: first pass of raymarching.cu:377-439 for every ray, prefix
 *            sum, writes `rays` [N,3] and accumulates `counter`;
 *   phase 2 `_write`: second pass of raymarching.cu:463-533 into xyzs/dirs/deltas
 *            (M rows available).  If zero_tail_end > 0, rows
 *            [min(M, counter[0]), min(M, zero_tail_end)) are zero-filled so the
 *            caller may pass uninitialised buffers. */
int nvsf_march_rays_train_count(const float* rays_o, const float* rays_d, const uint8_t* grid,
                                float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                                uint32_t C, uint32_t H, const float* nears, const float* fars,
                                int32_t* rays, int32_t* counter, const float* noises,
                                void* workspace, size_t workspace_bytes, void* stream);
int nvsf_march_rays_train_write(const float* rays_o, const float* rays_d, const uint8_t* grid,
                                float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                                uint32_t C, uint32_t H, uint32_t M, const float* nears,
                                const float* fars, float* xyzs, float* dirs, float* deltas,
                                const int32_t* rays, const int32_t* counter,
                                const float* noises, uint32_t zero_tail_end, void* stream);

/* replaces composite_rays_train_forward (raymarching.h:45-54, kernel raymarching.cu:578-655). */
int nvsf_composite_rays_train_forward(const float* sigmas, const float* rgbs,
                                      const float* deltas, const int32_t* rays, uint32_t M,
                                      uint32_t N, float T_thresh, float* weights_sum,
                                      float* depth, float* image, void* stream);

/* replaces composite_rays_train_backward (raymarching.h:55-67, kernel raymarching.cu:691-772).
 * grad_sigmas [M], grad_rgbs [M,3] must be zero-filled by the caller (raymarching.py:338-339). */
int nvsf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                       const float* sigmas, const float* rgbs,
                                       const float* deltas, const int32_t* rays,
                                       const float* weights_sum, const float* image,
                                       uint32_t M, uint32_t N, float T_thresh,
                                       float* grad_sigmas, float* grad_rgbs, void* stream);

/* replaces march_rays (raymarching.h:69-86, kernel raymarching.cu:809-928).
 * xyzs/dirs [>= n_alive*n_step, 3], deltas [.., 2].  Unlike the reference the
 * kernel writes EVERY slot of rows [0, n_alive*n_step) (zeros where the ray
 * produced no sample) and zero-fills rows [n_alive*n_step, M_padded), so the
 * caller does not have to pre-zero. */
int nvsf_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive,
                    const float* rays_t, const float* rays_o, const float* rays_d, float bound,
                    float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                    const uint8_t* grid, const float* nears, const float* fars, float* xyzs,
                    float* dirs, float* deltas, const float* noises, uint32_t M_padded,
                    void* stream);

/* replaces composite_rays (raymarching.h:87-96, kernel raymarching.cu:967-1053). In place. */
int nvsf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive,
                        float* rays_t, const float* sigmas, const float* rgbs,
                        const float* deltas, float* weights_sum, float* depth, float* image,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NVSF_B200_H_ */
