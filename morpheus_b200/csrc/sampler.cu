// Ray sampling: fixed-step marching through a binary occupancy grid (replacing
// nerfacc.OccGridEstimator.sampling as called at morpheus.py:629-638: sigma_fn=None, alpha_thre=0,
// stratified jitter, cone_angle=0) and the fixed-S synthetic lattice of the BASELINE configs.
// nerfacc's source is not part of the reference tree (parity unpinned); the rule
// implemented here is the one restated in oracle/render.py:sample_occgrid and tested bit-for-bit
// against it: lattice t_k = max(tmin,near) + (jitter + k) * step, keep [t_k, t_k+step) when its
// midpoint lies in an occupied cell and t_k + step/2 < min(tmax, far).
// Two passes (count -> caller's exclusive scan -> write): one thread per ray, the 128^3 bitfield
// (2 MB as bytes) stays L2/L1 resident.
#include "common.cuh"

namespace mb {

struct Aabb { float lo[3], hi[3]; };

__device__ __forceinline__ bool slab(const float o[3], const float d[3], const Aabb& bb, float near_p, float far_p,
                                     float& tmin, float& tmax) {
    tmin = -INFINITY;
    tmax = INFINITY;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float inv = __fdiv_rn(1.0f, d[a]);
        const float ta = __fmul_rn(__fsub_rn(bb.lo[a], o[a]), inv);
        const float tb = __fmul_rn(__fsub_rn(bb.hi[a], o[a]), inv);
        tmin = fmaxf(tmin, fminf(ta, tb));
        tmax = fminf(tmax, fmaxf(ta, tb));
    }
    tmin = fmaxf(fmaxf(tmin, 0.0f), near_p);
    tmax = fminf(tmax, far_p);
    return tmax > tmin;
}

template <bool WRITE>
__global__ void march_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, uint32_t N,
                             const uint8_t* __restrict__ binaries, uint32_t res, Aabb bb, float step, float near_p, float far_p,
                             const float* __restrict__ jitter, int32_t* __restrict__ counts, const int32_t* __restrict__ offsets,
                             int64_t* __restrict__ ray_indices, float* __restrict__ t_starts, float* __restrict__ t_ends) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    float o[3], d[3];
#pragma unroll
    for (int a = 0; a < 3; a++) { o[a] = rays_o[3 * r + a]; d[a] = rays_d[3 * r + a]; }
    float tmin, tmax;
    int n = 0;
    const int base = WRITE ? offsets[r] : 0;
    if (slab(o, d, bb, near_p, far_p, tmin, tmax)) {
        float t = __fadd_rn(tmin, __fmul_rn(jitter ? jitter[r] : 0.0f, step));
        const float half = __fmul_rn(0.5f, step);
        const float fres = (float)res;
        while (true) {
            const float mid = __fadd_rn(t, half);
            if (!(mid < tmax)) break;
            uint32_t c[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const float p = __fadd_rn(o[a], __fmul_rn(d[a], mid));
                const float u = __fmul_rn(__fdiv_rn(__fsub_rn(p, bb.lo[a]), __fsub_rn(bb.hi[a], bb.lo[a])), fres);
                const int ci = (int)floorf(u);
                c[a] = (uint32_t)min(max(ci, 0), (int)res - 1);
            }
            if (binaries[((size_t)c[0] * res + c[1]) * res + c[2]]) {
                if (WRITE) {
                    ray_indices[base + n] = (int64_t)r;
                    t_starts[base + n] = t;
                    t_ends[base + n] = __fadd_rn(t, step);
                }
                n++;
            }
            t = __fadd_rn(t, step);
        }
    }
    if (!WRITE) counts[r] = n;
}

__global__ void uniform_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, uint32_t N, uint32_t S,
                               Aabb bb, const float* __restrict__ jitter, int64_t* __restrict__ ray_indices,
                               float* __restrict__ t_starts, float* __restrict__ t_ends) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * S) return;
    const uint32_t r = i / S, k = i - r * S;
    float o[3], d[3];
#pragma unroll
    for (int a = 0; a < 3; a++) { o[a] = rays_o[3 * r + a]; d[a] = rays_d[3 * r + a]; }
    float tmin, tmax;
    if (!slab(o, d, bb, 0.0f, INFINITY, tmin, tmax)) tmax = __fadd_rn(tmin, 1.0f);
    const float delta = __fdiv_rn(__fsub_rn(tmax, tmin), (float)(S + 1));
    const float a0 = __fadd_rn(tmin, __fmul_rn(__fadd_rn((float)k, jitter ? jitter[r] : 0.0f), delta));
    ray_indices[i] = (int64_t)r;
    t_starts[i] = a0;
    t_ends[i] = __fadd_rn(a0, delta);
}

static Aabb make_aabb(const float* h) {
    Aabb b;
    for (int a = 0; a < 3; a++) { b.lo[a] = h[a]; b.hi[a] = h[3 + a]; }
    return b;
}

// ---- occupancy refresh (nerfacc update_every_n_steps body; morpheus.py:905-913) -----------------
__global__ void occ_update_kernel(float* __restrict__ occs, const int64_t* __restrict__ idx, const float* __restrict__ sigma,
                                  uint32_t n, float decay, float step) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = idx ? idx[i] : (int64_t)i;
    occs[c] = fmaxf(occs[c] * decay, sigma[i] * step);
}
__global__ void occ_binarize_kernel(const float* __restrict__ occs, uint32_t n, float thre, const float* __restrict__ thre_dev,
                                    uint8_t* __restrict__ bin) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (thre_dev) thre = __ldg(thre_dev);      // device-resident threshold: no device->host round trip in the refresh
    if (i < n) bin[i] = occs[i] > thre ? 1 : 0;
}

// ---- fused Adam (torch.optim.Adam semantics: bias-corrected, eps added after sqrt(v_hat)) -----------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            const uint8_t* __restrict__ gid, const float* __restrict__ glr, uint64_t n, float b1, float b2,
                            float eps, float bc1, float bc2_sqrt, const int32_t* __restrict__ step_dev) {
    if (step_dev) {   // device-resident step count: the launch is then replayable inside a CUDA graph
        const float t = (float)__ldg(step_dev);
        bc1 = 1.0f - powf(b1, t);
        bc2_sqrt = sqrtf(1.0f - powf(b2, t));
    }
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = m[i] + (1.0f - b1) * (gi - m[i]);            // lerp form used by torch
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float lr = glr[gid ? gid[i] : 0];
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - (lr / bc1) * (mi / denom);
    }
}

// Adam with torch.optim.Adam's PER-PARAMETER bookkeeping restated per group (morpheus.py:154-155 builds one Adam over the named groups of
// models/model.py:313-324; on torch >= 2.0 zero_grad() sets .grad = None and Adam SKIPS such parameters: no moment decay, no step
// increment, no move): group_active[g] == 0 -> the group's elements are left untouched; group_step[g] (incremented by the prologue
// launch for active groups only) feeds the bias corrections.  zero_after != 0 folds the next step's zero_grad() into this launch.
__global__ void adam_groups_prologue(int32_t* __restrict__ group_step, const uint8_t* __restrict__ group_active, int n_groups) {
    const int g = threadIdx.x;
    if (g < n_groups && group_active[g]) group_step[g] += 1;
}
__global__ void adam_groups_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                   const uint8_t* __restrict__ gid, const float* __restrict__ glr, const uint8_t* __restrict__ group_active,
                                   const int32_t* __restrict__ group_step, int n_groups, uint64_t n, float b1, float b2, float eps, int zero_after) {
    __shared__ float s_lr_bc1[32], s_bc2[32];
    __shared__ uint8_t s_act[32];
    if (threadIdx.x < 32) {
        const int q = threadIdx.x;
        float lr_bc1 = 0.f, bc2 = 1.f;
        uint8_t a = 0;
        if (q < n_groups) {
            a = group_active[q];
            const float t = (float)max(group_step[q], 1);
            lr_bc1 = glr[q] / (1.0f - powf(b1, t));
            bc2 = sqrtf(1.0f - powf(b2, t));
        }
        s_lr_bc1[q] = lr_bc1; s_bc2[q] = bc2; s_act[q] = a;
    }
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int q = gid[i];
        if (s_act[q]) {
            const float gi = g[i];
            const float mi = m[i] + (1.0f - b1) * (gi - m[i]);            // lerp form used by torch
            const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
            m[i] = mi;
            v[i] = vi;
            const float denom = sqrtf(vi) / s_bc2[q] + eps;
            p[i] = p[i] - s_lr_bc1[q] * (mi / denom);
        }
        if (zero_after) g[i] = 0.f;
    }
}

// ---- SDS scalar chain (zero123_utils.py:177-212) -----------------------------------------------------
__global__ void sds_grad_kernel(const float* __restrict__ eu, const float* __restrict__ ec, const float* __restrict__ noise,
                                float s, float wg, float* __restrict__ grad, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float pred = eu[i] + s * (ec[i] - eu[i]);
    float g = wg * (pred - noise[i]);
    if (isnan(g)) g = 0.f;                                 // torch.nan_to_num
    else if (isinf(g)) g = g > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    grad[i] = g;
}
__global__ void add_noise_kernel(const float* __restrict__ z, const float* __restrict__ e, float a, float b, float* __restrict__ out,
                                 uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a * z[i] + b * e[i];
}

// device-scalar variants of the SDS chain: t, abar_t and the per-step weights stay on the device, so Zero123.train_step has no
// device->host synchronisation and the whole chain (VAE encode, add-noise, UNet, gradient, VAE input-gradient) is ONE CUDA graph
__global__ void add_noise_dev_kernel(const float* __restrict__ z, const float* __restrict__ e, const float* __restrict__ alphas_cumprod,
                                     const int64_t* __restrict__ t, float* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float ab = alphas_cumprod[t[0]];
    out[i] = sqrtf(ab) * z[i] + sqrtf(1.0f - ab) * e[i];
}
__global__ void sds_grad_dev_kernel(const float* __restrict__ eu, const float* __restrict__ ec, const float* __restrict__ noise, float s,
                                    const float* __restrict__ grad_scale, const float* __restrict__ alphas_cumprod, const int64_t* __restrict__ t,
                                    float view_weight, float* __restrict__ grad, uint32_t n, int accumulate) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float wg = grad_scale[0] * (1.0f - alphas_cumprod[t[0]]);       // grad_scale * w(t), w = 1 - abar_t (zero123_utils.py:210)
    const float pred = eu[i] + s * (ec[i] - eu[i]);
    float g = view_weight * (wg * (pred - noise[i]));
    if (accumulate) g += grad[i];
    grad[i] = g;
}

}  // namespace mb

using namespace mb;

extern "C" int mb_sample_rays_count(const float* rays_o, const float* rays_d, uint32_t N, const uint8_t* binaries, uint32_t res,
                                    const float* aabb_host6, float step, float near_plane, float far_plane, const float* jitter,
                                    int32_t* counts, mb_stream_t stream) {
    if (N == 0) return MB_OK;
    if (!rays_o || !rays_d || !binaries || !aabb_host6 || !counts) { set_error("sample_rays_count: null pointer"); return MB_EINVAL; }
    if (!(step > 0)) { set_error("sample_rays_count: step must be > 0"); return MB_EINVAL; }
    march_kernel<false><<<div_up(N, 128), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, N, binaries, res, make_aabb(aabb_host6), step,
                                                                          near_plane, far_plane, jitter, counts, nullptr, nullptr, nullptr, nullptr);
    return check_launch("sample_rays_count");
}

extern "C" int mb_sample_rays_write(const float* rays_o, const float* rays_d, uint32_t N, const uint8_t* binaries, uint32_t res,
                                    const float* aabb_host6, float step, float near_plane, float far_plane, const float* jitter,
                                    const int32_t* offsets, int64_t* ray_indices, float* t_starts, float* t_ends, mb_stream_t stream) {
    if (N == 0) return MB_OK;
    if (!rays_o || !rays_d || !binaries || !aabb_host6 || !offsets || !ray_indices || !t_starts || !t_ends) {
        set_error("sample_rays_write: null pointer");
        return MB_EINVAL;
    }
    march_kernel<true><<<div_up(N, 128), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, N, binaries, res, make_aabb(aabb_host6), step,
                                                                         near_plane, far_plane, jitter, nullptr, offsets, ray_indices, t_starts, t_ends);
    return check_launch("sample_rays_write");
}

extern "C" int mb_sample_rays_uniform(const float* rays_o, const float* rays_d, uint32_t N, uint32_t S, const float* aabb_host6,
                                      const float* jitter, int64_t* ray_indices, float* t_starts, float* t_ends, mb_stream_t stream) {
    if (N == 0 || S == 0) return MB_OK;
    if (!rays_o || !rays_d || !aabb_host6 || !ray_indices || !t_starts || !t_ends) { set_error("sample_rays_uniform: null pointer"); return MB_EINVAL; }
    uniform_kernel<<<div_up(N * S, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, N, S, make_aabb(aabb_host6), jitter, ray_indices,
                                                                         t_starts, t_ends);
    return check_launch("sample_rays_uniform");
}

extern "C" int mb_occ_update(float* occs, const int64_t* cell_idx, const float* sigma, uint32_t n, float decay, float step,
                             mb_stream_t stream) {
    if (n == 0) return MB_OK;
    if (!occs || !sigma) { set_error("occ_update: null pointer"); return MB_EINVAL; }
    occ_update_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(occs, cell_idx, sigma, n, decay, step);
    return check_launch("occ_update");
}

extern "C" int mb_occ_binarize(const float* occs, uint32_t n, float thre, uint8_t* binaries, mb_stream_t stream) {
    if (n == 0) return MB_OK;
    if (!occs || !binaries) { set_error("occ_binarize: null pointer"); return MB_EINVAL; }
    occ_binarize_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(occs, n, thre, nullptr, binaries);
    return check_launch("occ_binarize");
}

extern "C" int mb_occ_binarize_dev(const float* occs, uint32_t n, const float* thre_dev, uint8_t* binaries, mb_stream_t stream) {
    if (n == 0) return MB_OK;
    if (!occs || !binaries || !thre_dev) { set_error("occ_binarize_dev: null pointer"); return MB_EINVAL; }
    occ_binarize_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(occs, n, 0.f, thre_dev, binaries);
    return check_launch("occ_binarize_dev");
}

extern "C" int mb_adam_step(float* p, const float* g, float* m, float* v, const uint8_t* group_id, const float* group_lr, uint64_t n,
                            float beta1, float beta2, float eps, int step, mb_stream_t stream) {
    if (n == 0) return MB_OK;
    if (!p || !g || !m || !v || !group_lr || step < 1) { set_error("adam_step: bad argument"); return MB_EINVAL; }
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
    int sms = mb_sm_count();
    uint64_t blocks = (n + 255) / 256;
    if (blocks > (uint64_t)sms * 8) blocks = (uint64_t)sms * 8;
    adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, group_id, group_lr, n, beta1, beta2, eps, bc1, bc2_sqrt, nullptr);
    return check_launch("adam_step");
}

extern "C" int mb_adam_step_dev(float* p, const float* g, float* m, float* v, const uint8_t* group_id, const float* group_lr, uint64_t n,
                                float beta1, float beta2, float eps, const int32_t* step_dev, mb_stream_t stream) {
    if (n == 0) return MB_OK;
    if (!p || !g || !m || !v || !group_lr || !step_dev) { set_error("adam_step_dev: bad argument"); return MB_EINVAL; }
    int sms = mb_sm_count();
    uint64_t blocks = (n + 255) / 256;
    if (blocks > (uint64_t)sms * 8) blocks = (uint64_t)sms * 8;
    adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, group_id, group_lr, n, beta1, beta2, eps, 1.f, 1.f, step_dev);
    return check_launch("adam_step_dev");
}

extern "C" int mb_adam_step_groups(float* p, float* g, float* m, float* v, const uint8_t* group_id, const float* group_lr,
                                   const uint8_t* group_active, int32_t* group_step, int n_groups, uint64_t n, float beta1, float beta2,
                                   float eps, int zero_after, mb_stream_t stream) {
    if (n == 0) return MB_OK;
    if (!p || !g || !m || !v || !group_id || !group_lr || !group_active || !group_step || n_groups < 1 || n_groups > 32) {
        set_error("adam_step_groups: bad argument (1 <= n_groups <= 32, no null pointers)");
        return MB_EINVAL;
    }
    adam_groups_prologue<<<1, 32, 0, (cudaStream_t)stream>>>(group_step, group_active, n_groups);
    int sms = mb_sm_count();
    uint64_t blocks = (n + 255) / 256;
    if (blocks > (uint64_t)sms * 8) blocks = (uint64_t)sms * 8;
    adam_groups_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, group_id, group_lr, group_active, group_step, n_groups, n,
                                                                          beta1, beta2, eps, zero_after);
    return check_launch("adam_step_groups");
}

extern "C" int mb_sds_grad(const float* eps_uncond, const float* eps_cond, const float* noise, float guidance_scale,
                           float w_t_times_grad_scale, float* grad, uint32_t n, mb_stream_t stream) {
    if (n == 0) return MB_OK;
    if (!eps_uncond || !eps_cond || !noise || !grad) { set_error("sds_grad: null pointer"); return MB_EINVAL; }
    sds_grad_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(eps_uncond, eps_cond, noise, guidance_scale, w_t_times_grad_scale, grad, n);
    return check_launch("sds_grad");
}

extern "C" int mb_add_noise_dev(const float* z, const float* eps, const float* alphas_cumprod, const int64_t* t_dev, float* out, uint32_t n,
                                mb_stream_t stream) {
    if (n == 0) return MB_OK;
    if (!z || !eps || !alphas_cumprod || !t_dev || !out) { set_error("add_noise_dev: null pointer"); return MB_EINVAL; }
    add_noise_dev_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(z, eps, alphas_cumprod, t_dev, out, n);
    return check_launch("add_noise_dev");
}

extern "C" int mb_sds_grad_dev(const float* eps_uncond, const float* eps_cond, const float* noise, float guidance_scale, const float* grad_scale_dev,
                               const float* alphas_cumprod, const int64_t* t_dev, float view_weight, float* grad, uint32_t n, int accumulate,
                               mb_stream_t stream) {
    if (n == 0) return MB_OK;
    if (!eps_uncond || !eps_cond || !noise || !grad_scale_dev || !alphas_cumprod || !t_dev || !grad) { set_error("sds_grad_dev: null pointer"); return MB_EINVAL; }
    sds_grad_dev_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(eps_uncond, eps_cond, noise, guidance_scale, grad_scale_dev, alphas_cumprod,
                                                                        t_dev, view_weight, grad, n, accumulate);
    return check_launch("sds_grad_dev");
}

extern "C" int mb_add_noise(const float* z, const float* eps, float sqrt_abar, float sqrt_one_minus_abar, float* out, uint32_t n,
                            mb_stream_t stream) {
    if (n == 0) return MB_OK;
    if (!z || !eps || !out) { set_error("add_noise: null pointer"); return MB_EINVAL; }
    add_noise_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(z, eps, sqrt_abar, sqrt_one_minus_abar, out, n);
    return check_launch("add_noise");
}
