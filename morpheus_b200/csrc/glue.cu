// Host-glue kernels of the render-and-loss step: each replaces a swarm of tiny eager launches of the reference
// (and of our own first torch-level host code) by one launch forward and one backward.
//
//   mb_ray_points_*   xyz = o[ray] + d[ray] * (t0+t1)/2                     morpheus.py:645-646   (+ segmented-sum backward
//                     instead of torch's index_put(accumulate) kernels, 0.5 ms each at M = 524 288)
//   mb_pack_arena_*   nn.utils.weight_norm (W = v * g/||v||_row, models/decoders.py:51-52) of the 18 dense layers on the
//                     hot path, transposed / padded into the parameter arena of include/morpheus_b200.h; backward maps
//                     the flat gradient arena back onto weight_v / weight_g / weight / bias
//   mb_sdf_loss_*     utils.get_sdf_loss (utils.py:91-113) on packed samples, forward sums + per-sample gradient
#include "common.cuh"

namespace mb {

// ---------------------------------------------------------------------------------------------------------------
__global__ void ray_points_fwd_kernel(const float* __restrict__ o, const float* __restrict__ d, const int64_t* __restrict__ ri,
                                      const float* __restrict__ t0, const float* __restrict__ t1, uint32_t M, float* __restrict__ xyz) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int64_t r = ri[i];
    const float tm = __fdiv_rn(__fadd_rn(t0[i], t1[i]), 2.0f);
#pragma unroll
    for (int a = 0; a < 3; a++) xyz[(size_t)i * 3 + a] = __fadd_rn(__ldg(o + r * 3 + a), __fmul_rn(__ldg(d + r * 3 + a), tm));
}

// one warp per ray: g_o[r] = sum g_xyz, g_d[r] = sum g_xyz * t_mid over the ray's packed samples
__global__ void ray_points_bwd_kernel(const int32_t* __restrict__ seg, uint32_t N, const float* __restrict__ t0, const float* __restrict__ t1,
                                      const float* __restrict__ g_xyz, float* __restrict__ g_o, float* __restrict__ g_d) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= N) return;
    const int s0 = seg[r], s1 = seg[r + 1];
    float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = s0 + lane; i < s1; i += 32) {
        const float tm = __fdiv_rn(__fadd_rn(t0[i], t1[i]), 2.0f);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float g = g_xyz[(size_t)i * 3 + c];
            a[c] += g;
            a[3 + c] += g * tm;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < 6; c++) a[c] += __shfl_xor_sync(0xffffffffu, a[c], o);
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (g_o) g_o[(size_t)r * 3 + c] = a[c];
            if (g_d) g_d[(size_t)r * 3 + c] = a[3 + c];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// layer table, 8 x int64 per layer: {v_ptr (weight_v or weight), g_ptr (weight_g or 0), b_ptr, K, N, K_pad, N_pad, wt_off}
// grad table, 4 x int64 per layer: {gv_off, gg_off (-1 if none), gb_off, 0} (float offsets into the flat gradient buffer)
__global__ void pack_arena_fwd_kernel(const int64_t* __restrict__ tab, float* __restrict__ arena) {
    const int64_t* t = tab + (size_t)blockIdx.y * 8;
    const float* v = reinterpret_cast<const float*>(t[0]);
    const float* g = reinterpret_cast<const float*>(t[1]);
    const float* b = reinterpret_cast<const float*>(t[2]);
    const int K = (int)t[3], N = (int)t[4], Kp = (int)t[5], Np = (int)t[6];
    const size_t wt_off = (size_t)t[7], w_off = wt_off + (size_t)Kp * Np, b_off = w_off + (size_t)Kp * Np;
    const int n = blockIdx.x;
    if (n >= Np) return;
    __shared__ float s_scale;
    if (threadIdx.x < 32) {
        float s = 1.0f;
        if (g && n < N) {
            double acc = 0.0;
            for (int k = threadIdx.x; k < K; k += 32) { const double x = v[(size_t)n * K + k]; acc += x * x; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            s = __fdiv_rn(g[n], (float)sqrt(acc));
        }
        if (threadIdx.x == 0) s_scale = s;
    }
    __syncthreads();
    const float s = s_scale;
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
        float w = 0.f;
        if (n < N && k < K) {
            w = v[(size_t)n * K + k];
            if (g) w = __fmul_rn(w, s);
        }
        arena[w_off + (size_t)n * Kp + k] = w;
        arena[wt_off + (size_t)k * Np + n] = w;
    }
    if (threadIdx.x == 0) arena[b_off + n] = (n < N) ? b[n] : 0.f;
}

// direct != 0: gtab holds absolute device addresses of the parameters' .grad tensors and the gradients are ACCUMULATED there
// (the caller's optimiser owns one flat, zero-filled gradient buffer: no per-tensor autograd accumulation launches)
__global__ void pack_arena_bwd_kernel(const int64_t* __restrict__ tab, const int64_t* __restrict__ gtab, const float* __restrict__ g_arena,
                                      float* __restrict__ flat, const int direct) {
    const int64_t* t = tab + (size_t)blockIdx.y * 8;
    const int64_t* gt = gtab + (size_t)blockIdx.y * 4;
    const float* v = reinterpret_cast<const float*>(t[0]);
    const float* g = reinterpret_cast<const float*>(t[1]);
    const int K = (int)t[3], N = (int)t[4], Np = (int)t[6];
    const size_t wt_off = (size_t)t[7], b_off = wt_off + 2 * (size_t)t[5] * Np;
    const int n = blockIdx.x;
    if (n >= N) return;
    float* gv = direct ? reinterpret_cast<float*>(gt[0]) : flat + gt[0];
    float* gg = direct ? reinterpret_cast<float*>(gt[1]) : flat + gt[1];
    float* gb = direct ? reinterpret_cast<float*>(gt[2]) : flat + gt[2];
    __shared__ float s_dot, s_norm;
    if (g) {
        if (threadIdx.x < 32) {
            double acc = 0.0, dot = 0.0;
            for (int k = threadIdx.x; k < K; k += 32) {
                const double x = v[(size_t)n * K + k];
                acc += x * x;
                dot += x * (double)g_arena[wt_off + (size_t)k * Np + n];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); dot += __shfl_xor_sync(0xffffffffu, dot, o); }
            if (threadIdx.x == 0) { s_dot = (float)dot; s_norm = (float)sqrt(acc); }
        }
        __syncthreads();
        const float norm = s_norm, dot = s_dot, gn = g[n];
        const float s = gn / norm, c = gn * dot / (norm * norm * norm);
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const float val = g_arena[wt_off + (size_t)k * Np + n] * s - c * v[(size_t)n * K + k];
            if (direct) gv[(size_t)n * K + k] += val; else gv[(size_t)n * K + k] = val;
        }
        if (threadIdx.x == 0) { if (direct) gg[n] += dot / norm; else gg[n] = dot / norm; }
    } else {
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const float val = g_arena[wt_off + (size_t)k * Np + n];
            if (direct) gv[(size_t)n * K + k] += val; else gv[(size_t)n * K + k] = val;
        }
    }
    if (threadIdx.x == 0) { if (direct) gb[n] += g_arena[b_off + n]; else gb[n] = g_arena[b_off + n]; }
}

// ---------------------------------------------------------------------------------------------------------------
// utils.get_sdf_loss (utils.py:91-113) on packed samples: z = (t0+t1)/2, d_gt = depth[ray], mask = mask[ray].
// The reference normalises per SAMPLE (sum(dim=-1) over a size-1 axis, utils.py:106-111): n_i = front_i + band_i + 1e-8.
// out[0] += sum_i fs_i / n_i ; out[1] += sum_i |s_i - bound_i| band_i / n_i  (caller divides by the number of rays with depth)
__device__ __forceinline__ void sdf_loss_terms(float z, float dgt, float msk, bool has_mask, float s, float trunc, float& fs, float& sl,
                                               float& dfs, float& dsl) {
    const bool depth_ok = dgt > 0.f;
    const bool front = (z < (dgt - trunc)) || ((dgt < 0.f) && (z < 3.5f));
    const float bound = (dgt < 0.f) ? 10.f : (dgt - z);
    bool band = (fabsf(bound) <= trunc) && depth_ok;
    if (has_mask) band = band && (msk > 0.5f);
    const float n = (front ? 1.f : 0.f) + (band ? 1.f : 0.f) + 1e-8f;
    fs = sl = dfs = dsl = 0.f;
    if (front) {
        const float e = expf(-5.f * s) - 1.f, lin = s - bound;
        const float mx = fmaxf(e, lin);
        if (mx >= 0.f) {      // clamp(min=0) passes the gradient at 0 (torch: self >= min)
            fs = mx / n;
            dfs = ((e > lin) ? -5.f * expf(-5.f * s) : 1.f) / n;      // torch.max backward: ties go to the first argument's... (measure-zero)
            if (e == lin) dfs = 0.5f * (-5.f * expf(-5.f * s) + 1.f) / n;
        }
    }
    if (band) {
        const float df = s - bound;
        sl = fabsf(df) / n;
        dsl = ((df > 0.f) ? 1.f : (df < 0.f ? -1.f : 0.f)) / n;
    }
}

__global__ void sdf_loss_fwd_kernel(const float* __restrict__ t0, const float* __restrict__ t1, const int64_t* __restrict__ ri,
                                    const float* __restrict__ depth, const float* __restrict__ mask, const float* __restrict__ sdf,
                                    uint32_t M, float trunc, float* __restrict__ out) {
    float a = 0.f, b = 0.f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const int64_t r = ri[i];
        float fs, sl, dfs, dsl;
        sdf_loss_terms(__fdiv_rn(__fadd_rn(t0[i], t1[i]), 2.0f), depth[r], mask ? mask[r] : 1.f, mask != nullptr, sdf[i], trunc, fs, sl, dfs, dsl);
        a += fs;
        b += sl;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    __shared__ float sa[32], sb[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sa[w] = a; sb[w] = b; }
    __syncthreads();
    if (w == 0) {
        a = lane < (int)(blockDim.x >> 5) ? sa[lane] : 0.f;
        b = lane < (int)(blockDim.x >> 5) ? sb[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (lane == 0) { atomicAdd(out, a); atomicAdd(out + 1, b); }
    }
}

// g_sdf[i] = g_out[0] * d(fs_i)/ds + g_out[1] * d(sl_i)/ds   (g_out: device scalars, already divided by the ray count)
__global__ void sdf_loss_bwd_kernel(const float* __restrict__ t0, const float* __restrict__ t1, const int64_t* __restrict__ ri,
                                    const float* __restrict__ depth, const float* __restrict__ mask, const float* __restrict__ sdf,
                                    uint32_t M, float trunc, const float* __restrict__ g_out, float* __restrict__ g_sdf) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int64_t r = ri[i];
    float fs, sl, dfs, dsl;
    sdf_loss_terms(__fdiv_rn(__fadd_rn(t0[i], t1[i]), 2.0f), depth[r], mask ? mask[r] : 1.f, mask != nullptr, sdf[i], trunc, fs, sl, dfs, dsl);
    g_sdf[i] = g_out[0] * dfs + g_out[1] * dsl;
}


// ---------------------------------------------------------------------------------------------------------------
// pose correction of a ray batch: scene_representation.pose_optimisation (models/model.py:335-346) with
// PoseArray.get_rotation_matrices (models/pose.py:35-58): R columns c1, c2, c3 from the Euler angles (a, b, g) of the ray's
// frame, o' = o + t_f, d'_i = sum_j d_j R[i][j].  Same operation order / roundings as the eager torch expression.
struct PoseR { float R[3][3]; float ca, cb, cg, sa, sb, sg; };
__device__ __forceinline__ PoseR pose_matrix(const float* __restrict__ pr) {
    PoseR P;
    P.ca = cosf(pr[0]); P.cb = cosf(pr[1]); P.cg = cosf(pr[2]);
    P.sa = sinf(pr[0]); P.sb = sinf(pr[1]); P.sg = sinf(pr[2]);
    const float ca = P.ca, cb = P.cb, cg = P.cg, sa = P.sa, sb = P.sb, sg = P.sg;
    P.R[0][0] = __fmul_rn(ca, cb); P.R[1][0] = __fmul_rn(sa, cb); P.R[2][0] = -sb;
    P.R[0][1] = __fsub_rn(__fmul_rn(__fmul_rn(ca, sb), sg), __fmul_rn(sa, cg));
    P.R[1][1] = __fadd_rn(__fmul_rn(__fmul_rn(sa, sb), sg), __fmul_rn(ca, cg));
    P.R[2][1] = __fmul_rn(cb, sg);
    P.R[0][2] = __fadd_rn(__fmul_rn(__fmul_rn(ca, sb), cg), __fmul_rn(sa, sg));
    P.R[1][2] = __fsub_rn(__fmul_rn(__fmul_rn(sa, sb), cg), __fmul_rn(ca, sg));
    P.R[2][2] = __fmul_rn(cb, cg);
    return P;
}

__global__ void pose_rays_fwd_kernel(const float* __restrict__ pose, const int64_t* __restrict__ ids, const float* __restrict__ o,
                                     const float* __restrict__ d, uint32_t N, float* __restrict__ o2, float* __restrict__ d2) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float* pr = pose + ids[i] * 6;
    const PoseR P = pose_matrix(pr);
    const float dv[3] = {d[(size_t)i * 3], d[(size_t)i * 3 + 1], d[(size_t)i * 3 + 2]};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        o2[(size_t)i * 3 + a] = __fadd_rn(o[(size_t)i * 3 + a], pr[3 + a]);
        d2[(size_t)i * 3 + a] = __fadd_rn(__fadd_rn(__fmul_rn(dv[0], P.R[a][0]), __fmul_rn(dv[1], P.R[a][1])), __fmul_rn(dv[2], P.R[a][2]));
    }
}

// g_pose[f][0..2] += sum_ij g_d2[i] d[j] dR[i][j]/d(angle), g_pose[f][3..5] += g_o2 ; g_o = g_o2, g_d[j] = sum_i g_d2[i] R[i][j]
__global__ void pose_rays_bwd_kernel(const float* __restrict__ pose, const int64_t* __restrict__ ids, const float* __restrict__ d,
                                     const float* __restrict__ g_o2, const float* __restrict__ g_d2, uint32_t N, float* __restrict__ g_pose,
                                     float* __restrict__ g_o, float* __restrict__ g_d) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    float gp[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int64_t f = -1;
    if (i < N) {
        f = ids[i];
        const PoseR P = pose_matrix(pose + f * 6);
        const float ca = P.ca, cb = P.cb, cg = P.cg, sa = P.sa, sb = P.sb, sg = P.sg;
        const float dv[3] = {d[(size_t)i * 3], d[(size_t)i * 3 + 1], d[(size_t)i * 3 + 2]};
        float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
        if (g_o2) { go[0] = g_o2[(size_t)i * 3]; go[1] = g_o2[(size_t)i * 3 + 1]; go[2] = g_o2[(size_t)i * 3 + 2]; }
        if (g_d2) { gd[0] = g_d2[(size_t)i * 3]; gd[1] = g_d2[(size_t)i * 3 + 1]; gd[2] = g_d2[(size_t)i * 3 + 2]; }
        // dR/da, dR/db, dR/dg (row i, column j)
        const float Ra[3][3] = {{-sa * cb, -sa * sb * sg - ca * cg, -sa * sb * cg + ca * sg},
                                {ca * cb, ca * sb * sg - sa * cg, ca * sb * cg + sa * sg},
                                {0.f, 0.f, 0.f}};
        const float Rb[3][3] = {{-ca * sb, ca * cb * sg, ca * cb * cg},
                                {-sa * sb, sa * cb * sg, sa * cb * cg},
                                {-cb, -sb * sg, -sb * cg}};
        const float Rg[3][3] = {{0.f, ca * sb * cg + sa * sg, -ca * sb * sg + sa * cg},
                                {0.f, sa * sb * cg - ca * sg, -sa * sb * sg - ca * cg},
                                {0.f, cb * cg, -cb * sg}};
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float gR = gd[r] * dv[c];
                gp[0] += gR * Ra[r][c];
                gp[1] += gR * Rb[r][c];
                gp[2] += gR * Rg[r][c];
            }
        gp[3] = go[0]; gp[4] = go[1]; gp[5] = go[2];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (g_o) g_o[(size_t)i * 3 + c] = go[c];
            if (g_d) g_d[(size_t)i * 3 + c] = gd[0] * P.R[0][c] + gd[1] * P.R[1][c] + gd[2] * P.R[2][c];
        }
    }
    // rays of a batch normally share one frame: reduce over the warp when they do
    const int64_t f0 = __shfl_sync(0xffffffffu, f, 0);
    const bool uniform = __all_sync(0xffffffffu, f == f0 || f < 0);
    if (uniform) {
#pragma unroll
        for (int o_ = 16; o_ > 0; o_ >>= 1)
#pragma unroll
            for (int c = 0; c < 6; c++) gp[c] += __shfl_xor_sync(0xffffffffu, gp[c], o_);
        if (lane == 0 && f0 >= 0)
#pragma unroll
            for (int c = 0; c < 6; c++) atomicAdd(g_pose + f0 * 6 + c, gp[c]);
    } else if (f >= 0) {
#pragma unroll
        for (int c = 0; c < 6; c++) atomicAdd(g_pose + f * 6 + c, gp[c]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// per-ray loss heads of a real view (morpheus.py:946-983): w_rgb * mse(image, gt) + w_mask * bce(clip(opacity), mask)
// + w_depth * mse(depth * dm, gt_depth * dm), dm = (gt_depth > 0) & (|o + gt_depth d| <= 1.1) & (mask > .5).
// One launch: out[0] += weighted loss; the per-ray gradients (already weighted, for d(loss) = 1) are stored for the backward.
__global__ void ray_loss_kernel(const float* __restrict__ image, const float* __restrict__ opacity, const float* __restrict__ depth,
                                const float* __restrict__ gt_rgb, const float* __restrict__ gt_depth, const float* __restrict__ gt_mask,
                                const float* __restrict__ rays_o, const float* __restrict__ rays_d, uint32_t N, float w_rgb, float w_mask,
                                float w_depth, float* __restrict__ out, float* __restrict__ g_image, float* __restrict__ g_opacity,
                                float* __restrict__ g_depth) {
    float acc = 0.f;
    const float inv_n = 1.0f / (float)N, inv_3n = 1.0f / (3.0f * (float)N);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        float l = 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float df = image[(size_t)i * 3 + c] - gt_rgb[(size_t)i * 3 + c];
            l += w_rgb * df * df * inv_3n;
            g_image[(size_t)i * 3 + c] = w_rgb * 2.f * df * inv_3n;
        }
        const float m = gt_mask[i], op = opacity[i];
        const float x = fminf(fmaxf(op, 1e-5f), 1.0f - 1e-5f);
        l += -w_mask * (m * fmaxf(logf(x), -100.f) + (1.f - m) * fmaxf(logf(1.f - x), -100.f)) * inv_n;
        g_opacity[i] = (op >= 1e-5f && op <= 1.0f - 1e-5f) ? w_mask * (x - m) / fmaxf((1.f - x) * x, 1e-12f) * inv_n : 0.f;
        const float gd = gt_depth[i];
        float px[3];
#pragma unroll
        for (int c = 0; c < 3; c++) px[c] = rays_o[(size_t)i * 3 + c] + gd * rays_d[(size_t)i * 3 + c];
        const float nrm = sqrtf(px[0] * px[0] + px[1] * px[1] + px[2] * px[2]);
        const float dm = (gd > 0.f && nrm <= 1.1f && m > 0.5f) ? 1.f : 0.f;
        const float dd = depth[i] * dm - gd * dm;
        l += w_depth * dd * dd * inv_n;
        g_depth[i] = w_depth * 2.f * dd * dm * inv_n;
        acc += l;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ float sa[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sa[w] = acc;
    __syncthreads();
    if (w == 0) {
        acc = lane < (int)(blockDim.x >> 5) ? sa[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) atomicAdd(out, acc);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// deformation-code regulariser (morpheus.py:762-771): mean_c (2 c(t) - c(t - 1/F) - c(t + 1/F))^2 over the 48 code channels,
// c(.) = MultiCode.sample (models/deform_code.py:20-40: align_corners linear interpolation on three lines [16, S_v]).
struct CodeTap { int i0, i1; float w0, w1; };
__device__ __forceinline__ CodeTap code_tap(float t, int S) {
    t = fminf(fmaxf(t, 0.f), 1.f);
    const float g = __fsub_rn(__fmul_rn(t, 2.f), 1.f);
    const float pos = __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), (float)(S - 1));
    CodeTap c;
    c.i0 = min(max((int)floorf(pos), 0), S - 1);
    c.i1 = min(c.i0 + 1, S - 1);
    c.w1 = pos - (float)c.i0;
    c.w0 = 1.f - c.w1;
    if (c.i0 + 1 > S - 1) c.w1 = 0.f;
    return c;
}
// one block of 64 threads; thread = channel (v = tid / 16, c = tid % 16) for tid < 48.  g_out == nullptr: forward (out[0] = loss);
// else backward: g_code[v][c][.] += g_out[0] * d loss / d line
__global__ void code_reg_kernel(const float* __restrict__ c0, const float* __restrict__ c1, const float* __restrict__ c2, int S0, int S1, int S2,
                                const float* __restrict__ t_ptr, float inv_frames, float* __restrict__ out, const float* __restrict__ g_out,
                                float* __restrict__ g0, float* __restrict__ g1, float* __restrict__ g2) {
    const int tid = threadIdx.x;
    float contrib = 0.f;
    if (tid < 48) {
        const int v = tid >> 4, c = tid & 15;
        const int S = v == 0 ? S0 : (v == 1 ? S1 : S2);
        const float* line = (v == 0 ? c0 : (v == 1 ? c1 : c2)) + (size_t)c * S;
        const float t = t_ptr[0];
        const CodeTap a = code_tap(t, S), b = code_tap(__fsub_rn(t, inv_frames), S), d = code_tap(__fadd_rn(t, inv_frames), S);
        const float va = line[a.i0] * a.w0 + line[a.i1] * a.w1, vb = line[b.i0] * b.w0 + line[b.i1] * b.w1, vd = line[d.i0] * d.w0 + line[d.i1] * d.w1;
        const float r = 2.f * va - vb - vd;
        contrib = r * r * (1.0f / 48.0f);
        if (g_out) {
            float* gl = (v == 0 ? g0 : (v == 1 ? g1 : g2)) + (size_t)c * S;
            const float k = g_out[0] * 2.f * r * (1.0f / 48.0f);
            atomicAdd(gl + a.i0, 2.f * k * a.w0); atomicAdd(gl + a.i1, 2.f * k * a.w1);
            atomicAdd(gl + b.i0, -k * b.w0); atomicAdd(gl + b.i1, -k * b.w1);
            atomicAdd(gl + d.i0, -k * d.w0); atomicAdd(gl + d.i1, -k * d.w1);
        }
    }
    if (!g_out) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        __shared__ float sacc[2];
        if ((tid & 31) == 0) sacc[tid >> 5] = contrib;
        __syncthreads();
        if (tid == 0) out[0] = sacc[0] + sacc[1];
    }
}

}  // namespace mb

extern "C" int mb_ray_points_forward(const float* rays_o, const float* rays_d, const int64_t* ray_indices, const float* t_starts,
                                     const float* t_ends, uint32_t M, float* xyz, mb_stream_t stream) {
    using namespace mb;
    if (M == 0) return MB_OK;
    if (!rays_o || !rays_d || !ray_indices || !t_starts || !t_ends || !xyz) { set_error("ray_points_forward: null argument"); return MB_EINVAL; }
    ray_points_fwd_kernel<<<div_up(M, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, ray_indices, t_starts, t_ends, M, xyz);
    return check_launch("ray_points_forward");
}

extern "C" int mb_ray_points_backward(const int32_t* seg, uint32_t N, const float* t_starts, const float* t_ends, const float* g_xyz,
                                      float* g_o, float* g_d, mb_stream_t stream) {
    using namespace mb;
    if (N == 0) return MB_OK;
    if (!seg || !t_starts || !t_ends || !g_xyz) { set_error("ray_points_backward: null argument"); return MB_EINVAL; }
    ray_points_bwd_kernel<<<div_up(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(seg, N, t_starts, t_ends, g_xyz, g_o, g_d);
    return check_launch("ray_points_backward");
}

extern "C" int mb_pack_arena_forward(const int64_t* layer_table, int n_layers, float* arena, mb_stream_t stream) {
    using namespace mb;
    if (!layer_table || !arena || n_layers <= 0) { set_error("pack_arena_forward: bad argument"); return MB_EINVAL; }
    pack_arena_fwd_kernel<<<dim3(128, n_layers), 128, 0, (cudaStream_t)stream>>>(layer_table, arena);
    return check_launch("pack_arena_forward");
}

extern "C" int mb_pack_arena_backward(const int64_t* layer_table, const int64_t* grad_table, int n_layers, const float* g_arena, float* flat_grads,
                                      mb_stream_t stream) {
    using namespace mb;
    if (!layer_table || !grad_table || !g_arena || n_layers <= 0) { set_error("pack_arena_backward: bad argument"); return MB_EINVAL; }
    pack_arena_bwd_kernel<<<dim3(128, n_layers), 128, 0, (cudaStream_t)stream>>>(layer_table, grad_table, g_arena, flat_grads, flat_grads == nullptr ? 1 : 0);
    return check_launch("pack_arena_backward");
}

extern "C" int mb_sdf_loss_forward(const float* t_starts, const float* t_ends, const int64_t* ray_indices, const float* depth, const float* mask,
                                   const float* sdf, uint32_t M, float truncation, float* out2, mb_stream_t stream) {
    using namespace mb;
    if (!t_starts || !t_ends || !ray_indices || !depth || !sdf || !out2) { set_error("sdf_loss_forward: null argument"); return MB_EINVAL; }
    if (M == 0) return MB_OK;
    const uint32_t grid = min(div_up(M, 256), (uint32_t)mb_sm_count() * 4u);
    sdf_loss_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t_starts, t_ends, ray_indices, depth, mask, sdf, M, truncation, out2);
    return check_launch("sdf_loss_forward");
}

extern "C" int mb_sdf_loss_backward(const float* t_starts, const float* t_ends, const int64_t* ray_indices, const float* depth, const float* mask,
                                    const float* sdf, uint32_t M, float truncation, const float* g_out2, float* g_sdf, mb_stream_t stream) {
    using namespace mb;
    if (!t_starts || !t_ends || !ray_indices || !depth || !sdf || !g_out2 || !g_sdf) { set_error("sdf_loss_backward: null argument"); return MB_EINVAL; }
    if (M == 0) return MB_OK;
    sdf_loss_bwd_kernel<<<div_up(M, 256), 256, 0, (cudaStream_t)stream>>>(t_starts, t_ends, ray_indices, depth, mask, sdf, M, truncation, g_out2, g_sdf);
    return check_launch("sdf_loss_backward");
}

extern "C" int mb_pose_rays_forward(const float* pose, const int64_t* frame_ids, const float* rays_o, const float* rays_d, uint32_t N,
                                    float* rays_o_out, float* rays_d_out, mb_stream_t stream) {
    using namespace mb;
    if (N == 0) return MB_OK;
    if (!pose || !frame_ids || !rays_o || !rays_d || !rays_o_out || !rays_d_out) { set_error("pose_rays_forward: null argument"); return MB_EINVAL; }
    pose_rays_fwd_kernel<<<div_up(N, 128), 128, 0, (cudaStream_t)stream>>>(pose, frame_ids, rays_o, rays_d, N, rays_o_out, rays_d_out);
    return check_launch("pose_rays_forward");
}

extern "C" int mb_pose_rays_backward(const float* pose, const int64_t* frame_ids, const float* rays_d, const float* g_o_out, const float* g_d_out,
                                     uint32_t N, float* g_pose, float* g_rays_o, float* g_rays_d, mb_stream_t stream) {
    using namespace mb;
    if (N == 0) return MB_OK;
    if (!pose || !frame_ids || !rays_d || !g_pose) { set_error("pose_rays_backward: null argument"); return MB_EINVAL; }
    pose_rays_bwd_kernel<<<div_up(N, 128), 128, 0, (cudaStream_t)stream>>>(pose, frame_ids, rays_d, g_o_out, g_d_out, N, g_pose, g_rays_o, g_rays_d);
    return check_launch("pose_rays_backward");
}

extern "C" int mb_ray_loss(const float* image, const float* opacity, const float* depth, const float* gt_rgb, const float* gt_depth,
                           const float* gt_mask, const float* rays_o, const float* rays_d, uint32_t N, float w_rgb, float w_mask, float w_depth,
                           float* out1, float* g_image, float* g_opacity, float* g_depth, mb_stream_t stream) {
    using namespace mb;
    if (!image || !opacity || !depth || !gt_rgb || !gt_depth || !gt_mask || !rays_o || !rays_d || !out1 || !g_image || !g_opacity || !g_depth) {
        set_error("ray_loss: null argument");
        return MB_EINVAL;
    }
    if (N == 0) return MB_OK;
    ray_loss_kernel<<<min(div_up(N, 256), (uint32_t)mb_sm_count()), 256, 0, (cudaStream_t)stream>>>(image, opacity, depth, gt_rgb, gt_depth, gt_mask, rays_o, rays_d,
                                                                                                     N, w_rgb, w_mask, w_depth, out1, g_image, g_opacity, g_depth);
    return check_launch("ray_loss");
}

extern "C" int mb_code_reg(const float* const code[3], const int code_len[3], const float* t_dev, float inv_frames, float* out1, const float* g_out1,
                           float* const g_code[3], mb_stream_t stream) {
    using namespace mb;
    if (!code || !code_len || !t_dev || (!out1 && !g_out1)) { set_error("code_reg: null argument"); return MB_EINVAL; }
    if (g_out1 && !g_code) { set_error("code_reg: backward needs g_code"); return MB_EINVAL; }
    code_reg_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(code[0], code[1], code[2], code_len[0], code_len[1], code_len[2], t_dev, inv_frames, out1, g_out1,
                                                         g_out1 ? g_code[0] : nullptr, g_out1 ? g_code[1] : nullptr, g_out1 ? g_code[2] : nullptr);
    return check_launch("code_reg");
}
