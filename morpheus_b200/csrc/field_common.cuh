// Tile-level building blocks of the fused scene-field kernels (fp32 SIMT engine).
//
// Layout convention: every activation lives in shared memory as [feature][sample] with a row
// pitch of P floats (P = TM in the forward kernel; P = TM + 4 in the backward kernel, where the
// skew makes the strided row reads of the weight-gradient outer products bank-conflict free).  256 threads = 16 (tx: samples) x 16
// (ty: outputs) register tiles; weights stream from L2 through a double-buffered cp.async ring.
#pragma once
#include "common.cuh"

namespace mb {

constexpr int FT = 256;             // threads per CTA
constexpr int WCHUNK = 8;           // k-rows per weight chunk
constexpr int WSTAGES = 3;          // cp.async ring depth (one __syncthreads per chunk)
constexpr int WBUF_FLOATS = WSTAGES * WCHUNK * 128;
constexpr float FD_EPS = 2e-3f;     // models/model.py:367

// out[NP][P] = act( bias + W^T in ),  Wg is [KP][NP] (inner-dim major), KP % 16 == 0, NP % 16 == 0.
// mask != nullptr: out = (mask > 0) ? out : 0   (ReLU derivative against a stored activation).
template <int TM, int P, int NP>
__device__ __forceinline__ void dense(const float* __restrict__ Wg, const float* __restrict__ bias, int KP,
                                      const float* __restrict__ in, float* __restrict__ out, float* __restrict__ wbuf,
                                      bool relu, const float* __restrict__ mask) {
    constexpr int SM_ = TM / 16;
    constexpr int NO = NP / 16;
    constexpr int CH = WCHUNK * NP;        // floats per weight chunk (contiguous in global)
    constexpr int STG = WCHUNK * 128;      // ring stage stride
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[SM_][NO];
#pragma unroll
    for (int j = 0; j < NO; j++) {
        const float b = bias ? __ldg(bias + ty * NO + j) : 0.f;
#pragma unroll
        for (int i = 0; i < SM_; i++) acc[i][j] = b;
    }
    const int nchunk = KP / WCHUNK;
#pragma unroll
    for (int pre = 0; pre < 2; pre++) {
        if (pre < nchunk)
            for (int i = tid; i < CH / 4; i += FT) cp_async16(wbuf + pre * STG + i * 4, Wg + (size_t)pre * CH + i * 4);
        cp_async_commit();
    }
    for (int c = 0; c < nchunk; c++) {
        cp_async_wait<1>();
        __syncthreads();
        if (c + 2 < nchunk) {
            float* dst = wbuf + ((c + 2) % WSTAGES) * STG;
            const float* src = Wg + (size_t)(c + 2) * CH;
            for (int i = tid; i < CH / 4; i += FT) cp_async16(dst + i * 4, src + i * 4);
        }
        cp_async_commit();
        const float* wb = wbuf + (c % WSTAGES) * STG;
#pragma unroll
        for (int kk = 0; kk < WCHUNK; kk++) {
            const float* ip = in + (c * WCHUNK + kk) * P + tx * SM_;
            float a[SM_];
            if (SM_ == 4) {
                const float4 v = *reinterpret_cast<const float4*>(ip);
                a[0] = v.x; a[1 % SM_] = v.y; a[2 % SM_] = v.z; a[3 % SM_] = v.w;
            } else if (SM_ == 2) {
                const float2 v = *reinterpret_cast<const float2*>(ip);
                a[0] = v.x; a[1 % SM_] = v.y;
            } else {
#pragma unroll
                for (int i = 0; i < SM_; i++) a[i] = ip[i];
            }
            const float* wp = wb + kk * NP + ty * NO;
            float w[NO];
            if (NO % 4 == 0) {
#pragma unroll
                for (int j = 0; j < NO; j += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(wp + j);
                    w[j] = v.x; w[(j + 1) % NO] = v.y; w[(j + 2) % NO] = v.z; w[(j + 3) % NO] = v.w;
                }
            } else if (NO % 2 == 0) {
#pragma unroll
                for (int j = 0; j < NO; j += 2) {
                    const float2 v = *reinterpret_cast<const float2*>(wp + j);
                    w[j] = v.x; w[(j + 1) % NO] = v.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < NO; j++) w[j] = wp[j];
            }
#pragma unroll
            for (int i = 0; i < SM_; i++)
#pragma unroll
                for (int j = 0; j < NO; j++) acc[i][j] = __fmaf_rn(a[i], w[j], acc[i][j]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int j = 0; j < NO; j++) {
        const int n = ty * NO + j;
#pragma unroll
        for (int i = 0; i < SM_; i++) {
            float v = acc[i][j];
            if (relu) v = fmaxf(v, 0.f);
            if (mask) v = mask[n * P + tx * SM_ + i] > 0.f ? v : 0.f;
            out[n * P + tx * SM_ + i] = v;
        }
    }
    __syncthreads();
}

// out[m] = bias0 + sum_k Wn0[k] * in[k][m]  -- only output row 0 of a layer (sdf of an FD query)
template <int TM, int P>
__device__ __forceinline__ void dense_row0(const float* __restrict__ Wn0, float bias0, int K, const float* __restrict__ in,
                                           float* __restrict__ out) {
    const int tid = threadIdx.x;
    if (tid < TM) {
        float acc = bias0;
        for (int k = 0; k < K; k++) acc = __fmaf_rn(__ldg(Wn0 + k), in[k * P + tid], acc);
        out[tid] = acc;
    }
    __syncthreads();
}

// gWt[k][n] += sum_m A[k][m] * dZ[n][m]   (k < K, n < N);  gb[n] += sum_m dZ[n][m]
template <int TM, int P, int KI, int NJ>
__device__ __forceinline__ void wgrad(const float* __restrict__ A, const float* __restrict__ dZ, int K, int N, int NPg,
                                      float* __restrict__ gWt, float* __restrict__ gb) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[KI][NJ];
#pragma unroll
    for (int i = 0; i < KI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) acc[i][j] = 0.f;
#pragma unroll 2
    for (int m = 0; m < TM; m += 4) {
        float4 a[KI], z[NJ];
#pragma unroll
        for (int i = 0; i < KI; i++) a[i] = *reinterpret_cast<const float4*>(A + (ty + 16 * i) * P + m);
#pragma unroll
        for (int j = 0; j < NJ; j++) z[j] = *reinterpret_cast<const float4*>(dZ + (tx + 16 * j) * P + m);
#pragma unroll
        for (int i = 0; i < KI; i++)
#pragma unroll
            for (int j = 0; j < NJ; j++) {
                acc[i][j] = __fmaf_rn(a[i].x, z[j].x, acc[i][j]);
                acc[i][j] = __fmaf_rn(a[i].y, z[j].y, acc[i][j]);
                acc[i][j] = __fmaf_rn(a[i].z, z[j].z, acc[i][j]);
                acc[i][j] = __fmaf_rn(a[i].w, z[j].w, acc[i][j]);
            }
    }
#pragma unroll
    for (int i = 0; i < KI; i++) {
        const int k = ty + 16 * i;
        if (k < K) {
#pragma unroll
            for (int j = 0; j < NJ; j++) {
                const int n = tx + 16 * j;
                if (n < N) red_add(gWt + (size_t)k * NPg + n, acc[i][j]);
            }
        }
    }
    if (gb && tid < N) {
        float s = 0.f;
        for (int m = 0; m < TM; m++) s += dZ[tid * P + m];
        red_add(gb + tid, s);
    }
    __syncthreads();
}

// ---- encodings ------------------------------------------------------------------------------------
// rows 0..38 of dst: [p, sin(2^k p), cos(2^k p)]_{k<6}; bands >= n_freq are zero (encodings.py:35-57)
template <int TM, int P>
__device__ __forceinline__ void build_freq(const float* __restrict__ sp, float* __restrict__ dst, int n_freq) {
    for (int idx = threadIdx.x; idx < 3 * TM; idx += FT) {
        const int a = idx / TM, m = idx - a * TM;
        const float v = sp[a * P + m];
        dst[a * P + m] = v;
        float f = 1.0f;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            float s = 0.f, c = 0.f;
            if (k < n_freq) sincosf(v * f, &s, &c);
            dst[(3 + 6 * k + a) * P + m] = s;
            dst[(6 + 6 * k + a) * P + m] = c;
            f *= 2.0f;
        }
    }
}

// gp[a][m] += g0 + sum_k f_k (g_sin cos(f p) - g_cos sin(f p))
template <int TM, int P>
__device__ __forceinline__ void freq_backward(const float* __restrict__ sp, const float* __restrict__ g, float* __restrict__ gp,
                                              int n_freq) {
    for (int idx = threadIdx.x; idx < 3 * TM; idx += FT) {
        const int a = idx / TM, m = idx - a * TM;
        const float v = sp[a * P + m];
        float acc = g[a * P + m];
        float f = 1.0f;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            if (k < n_freq) {
                float s, c;
                sincosf(v * f, &s, &c);
                acc += f * (g[(3 + 6 * k + a) * P + m] * c - g[(6 + 6 * k + a) * P + m] * s);
            }
            f *= 2.0f;
        }
        atomicAdd(gp + a * P + m, acc);
    }
}

// per-level constants, computed once per CTA (the reference kernel re-derives them per (sample, level): gridencoder.cu:132-133)
struct LevelInfo {
    uint32_t res;        // ceil(exp2f(l*S)*H) in float32
    uint32_t size;       // hashmap_size = offsets[l+1]-offsets[l]
    uint32_t off;        // offsets[l]
    uint32_t mask;       // size-1 when size is a power of two (index % size == index & mask), else 0
    uint32_t hashed;     // get_grid_index falls through to the hash (stride > hashmap_size after the dense walk)
};

struct GridCtx {
    const float* emb;
    const int32_t* offsets;
    float S;
    uint32_t H;
    uint32_t n_levels;
    float bound, two_bound;
    const LevelInfo* lv;   // [16] in shared memory
};

// threads 0..15 fill the table; caller synchronises
__device__ __forceinline__ void init_levels(LevelInfo* lv, const int32_t* __restrict__ offsets, float S, uint32_t H) {
    const int l = threadIdx.x;
    if (l < 16) {
        LevelInfo v;
        v.res = level_resolution(l, S, H);
        v.off = (uint32_t)offsets[l];
        v.size = (uint32_t)offsets[l + 1] - v.off;
        v.mask = (v.size & (v.size - 1)) == 0 ? v.size - 1 : 0;
        const unsigned long long r = v.res;
        // dense walk of get_grid_index (gridencoder.cu:66-70) for D=3: the three strides 1, res, res^2 must all be <= size,
        // then stride = res^3; hash iff that final stride exceeds the table (gridtype == hash)
        v.hashed = (r > v.size || r * r > v.size || r * r * r > v.size) ? 1u : 0u;
        lv[l] = v;
    }
}

// entry index of grid corner (x,y,z): identical values to get_grid_index<3,C>/C
__device__ __forceinline__ uint32_t corner_index(const LevelInfo& L, uint32_t x, uint32_t y, uint32_t z) {
    if (L.hashed) {
        const uint32_t h = x ^ (y * 2654435761u) ^ (z * 805459861u);
        return L.mask ? (h & L.mask) : (h % L.size);
    }
    return x + y * L.res + z * L.res * L.res;      // < res^3 <= size: the reference's % size is the identity here
}

// one (sample, level) evaluation; D=3, C=2, hash grid, align_corners=False, linear (the only
// configuration MorpheuS builds: models/model.py:144-157).  Writes feat[2]; optionally dfeat/du [3][2].
__device__ __forceinline__ void grid_eval(const GridCtx& g, uint32_t level, const float u[3], float feat[2], float (*dfdu)[2]) {
    feat[0] = feat[1] = 0.f;
    if (dfdu) {
#pragma unroll
        for (int d = 0; d < 3; d++) dfdu[d][0] = dfdu[d][1] = 0.f;
    }
    if (u[0] < 0 || u[0] > 1 || u[1] < 0 || u[1] > 1 || u[2] < 0 || u[2] > 1) return;
    const LevelInfo L = g.lv[level];
    const float2* tab = reinterpret_cast<const float2*>(g.emb) + L.off;
    const uint32_t res = L.res;
    float pos[3], dv;
    uint32_t pg[3], pg1[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        pos[d] = locate(u[d], res, false, 0, pg[d], dv);
        pg1[d] = min(pg[d] + 1, res - 1);
    }
    float2 corner[8];
#pragma unroll
    for (uint32_t idx = 0; idx < 8; idx++)
        corner[idx] = __ldg(tab + corner_index(L, (idx & 1) ? pg1[0] : pg[0], (idx & 2) ? pg1[1] : pg[1], (idx & 4) ? pg1[2] : pg[2]));
#pragma unroll
    for (uint32_t idx = 0; idx < 8; idx++) {
        float w = 1.0f;
#pragma unroll
        for (uint32_t d = 0; d < 3; d++) w = __fmul_rn(w, (idx & (1u << d)) ? pos[d] : __fsub_rn(1.0f, pos[d]));
        feat[0] = __fmaf_rn(w, corner[idx].x, feat[0]);
        feat[1] = __fmaf_rn(w, corner[idx].y, feat[1]);
    }
    if (dfdu) {
        const float scale = (float)res;
#pragma unroll
        for (uint32_t gd = 0; gd < 3; gd++) {
#pragma unroll
            for (uint32_t idx = 0; idx < 4; idx++) {
                float w = scale;
                uint32_t cl = 0;
#pragma unroll
                for (uint32_t nd = 0; nd < 2; nd++) {
                    const uint32_t d = (nd >= gd) ? nd + 1 : nd;
                    if (idx & (1u << nd)) { w = __fmul_rn(w, pos[d]); cl |= (1u << d); }
                    else w = __fmul_rn(w, __fsub_rn(1.0f, pos[d]));
                }
                const float2 l = corner[cl], r = corner[cl | (1u << gd)];
                dfdu[gd][0] = __fadd_rn(dfdu[gd][0], __fmul_rn(w, __fsub_rn(r.x, l.x)));
                dfdu[gd][1] = __fadd_rn(dfdu[gd][1], __fmul_rn(w, __fsub_rn(r.y, l.y)));
            }
        }
    }
}

// dst rows [2l + c] for l < 16 (levels >= n_levels are zero; grid.py:53).  sp: world-space points [3][P].
template <int TM, int P>
__device__ __forceinline__ void build_grid(const GridCtx& g, const float* __restrict__ sp, float* __restrict__ dst) {
    for (int idx = threadIdx.x; idx < 16 * TM; idx += FT) {
        const int l = idx / TM, m = idx - l * TM;
        float feat[2] = {0.f, 0.f};
        if ((uint32_t)l < g.n_levels) {
            float u[3];
#pragma unroll
            for (int d = 0; d < 3; d++) u[d] = __fdiv_rn(__fadd_rn(sp[d * P + m], g.bound), g.two_bound);  // grid.py:157
            grid_eval(g, l, u, feat, nullptr);
        }
        dst[(2 * l) * P + m] = feat[0];
        dst[(2 * l + 1) * P + m] = feat[1];
    }
}

// scatter d(feat) into the table gradient and accumulate d/dp into gp[3][P] (shared, atomics)
template <int TM, int P>
__device__ __forceinline__ void grid_backward(const GridCtx& g, const float* __restrict__ sp, const float* __restrict__ gfeat,
                                              float* __restrict__ gemb, float* __restrict__ gp) {
    for (int idx = threadIdx.x; idx < 16 * TM; idx += FT) {
        const int l = idx / TM, m = idx - l * TM;
        if ((uint32_t)l >= g.n_levels) continue;
        const float g0 = gfeat[(2 * l) * P + m], g1 = gfeat[(2 * l + 1) * P + m];
        if (g0 == 0.f && g1 == 0.f) continue;
        float u[3];
#pragma unroll
        for (int d = 0; d < 3; d++) u[d] = __fdiv_rn(__fadd_rn(sp[d * P + m], g.bound), g.two_bound);
        if (u[0] < 0 || u[0] > 1 || u[1] < 0 || u[1] > 1 || u[2] < 0 || u[2] > 1) continue;
        const LevelInfo L = g.lv[l];
        const uint32_t res = L.res;
        const float2* tab = reinterpret_cast<const float2*>(g.emb) + L.off;
        float* gt = gemb + 2 * (size_t)L.off;
        float pos[3], dv;
        uint32_t pg[3], pg1[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            pos[d] = locate(u[d], res, false, 0, pg[d], dv);
            pg1[d] = min(pg[d] + 1, res - 1);
        }
        uint32_t cidx[8];
#pragma unroll
        for (uint32_t c = 0; c < 8; c++) {
            float w = 1.0f;
#pragma unroll
            for (uint32_t d = 0; d < 3; d++) w = __fmul_rn(w, (c & (1u << d)) ? pos[d] : __fsub_rn(1.0f, pos[d]));
            cidx[c] = corner_index(L, (c & 1) ? pg1[0] : pg[0], (c & 2) ? pg1[1] : pg[1], (c & 4) ? pg1[2] : pg[2]);
            red_add2(gt + 2 * cidx[c], w * g0, w * g1);
        }
        if (gp) {
            const float scale = (float)res;
#pragma unroll
            for (uint32_t gd = 0; gd < 3; gd++) {
                float acc = 0.f;
#pragma unroll
                for (uint32_t i4 = 0; i4 < 4; i4++) {
                    float w = scale;
                    uint32_t cl = 0;
#pragma unroll
                    for (uint32_t nd = 0; nd < 2; nd++) {
                        const uint32_t d = (nd >= gd) ? nd + 1 : nd;
                        if (i4 & (1u << nd)) { w *= pos[d]; cl |= (1u << d); }
                        else w *= (1.0f - pos[d]);
                    }
                    const float2 lo = __ldg(tab + cidx[cl]), hi = __ldg(tab + cidx[cl | (1u << gd)]);
                    acc += w * ((hi.x - lo.x) * g0 + (hi.y - lo.y) * g1);
                }
                atomicAdd(gp + gd * P + m, acc / g.two_bound);
            }
        }
    }
}

__device__ __forceinline__ float laplace_sigma(float sdf, float beta) {
    // models/density.py:27   alpha * (0.5 + 0.5 * sign(sdf) * expm1(-|sdf| / beta))
    const float sg = (sdf > 0.f) ? 1.f : ((sdf < 0.f) ? -1.f : 0.f);
    return (1.0f / beta) * (0.5f + 0.5f * sg * expm1f(-fabsf(sdf) / beta));
}

}  // namespace mb
