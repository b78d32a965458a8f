// Fused scene-field forward: one launch evaluates, per tile of 64 samples held in shared memory,
//   freq-encode -> deform MLP + topology MLP (87->128x5->3|2) -> x' = x + deform
//   -> hash-grid gather + freq-encode -> SDF MLP (73->64->64->33) -> Laplace sigma
//   -> colour-grid gather -> colour MLP (64->64->64->3) -> sigmoid
//   -> optional 6-point finite-difference normal (6 more grid gathers + SDF MLPs) -> shading
// i.e. scene_representation.forward / density / normal / warp / get_sigma_albedo of
// /root/reference/models/model.py:273-307,367-398,412-437,439-533 without ever writing an [M, .]
// intermediate to HBM (the reference materialises every layer's activations, ~5 KB/sample).
// HBM traffic: 16-28 B/sample in, 4-64 B/sample out; weights (0.7 MB) and tables (6.7 MB) are L2 hits.
#include "field_common.cuh"

namespace mb {

constexpr int FWD_TM = 64;
constexpr int FWD_P = 64;

struct FwdSmem {
    static constexpr int P = FWD_P;
    static constexpr int IN0 = 0;                     // [96][P]  deform/topo input, later the SDF-net input [80][P]
    static constexpr int BUFA = IN0 + 96 * P;         // [128][P]
    static constexpr int BUFB = BUFA + 128 * P;       // [128][P]
    static constexpr int WBUF = BUFB + 128 * P;       // weight ring
    static constexpr int SX = WBUF + WBUF_FLOATS;     // [3][P] x
    static constexpr int SXW = SX + 3 * P;            // [3][P] x + deform
    static constexpr int SPT = SXW + 3 * P;           // [3][P] FD query point
    static constexpr int STOPO = SPT + 3 * P;         // [2][P]
    static constexpr int SDEF = STOPO + 2 * P;        // [3][P]
    static constexpr int SSDF = SDEF + 3 * P;         // [1][P]
    static constexpr int SQ = SSDF + P;               // [6][P] FD sdf values
    static constexpr int ST = SQ + 6 * P;             // [1][P] t
    static constexpr int SALB = ST + P;               // [3][P] albedo
    static constexpr int TOTAL = SALB + 3 * P;
};

// rows 39..86 of in0: MultiCode.sample (models/deform_code.py:20-40): align_corners=True linear
// interpolation at t*(S-1), zero padding for the tap past the end.
template <int TM, int P>
__device__ __forceinline__ void build_code(const mb_field_params& p, const float* __restrict__ st, float* __restrict__ dst) {
    for (int idx = threadIdx.x; idx < 48 * TM; idx += FT) {
        const int r = idx / TM, m = idx - r * TM;
        const int v = r >> 4, c = r & 15;
        const int S = (int)p.code_len[v];
        float t = fminf(fmaxf(st[m], 0.f), 1.f);
        const float g = __fsub_rn(__fmul_rn(t, 2.f), 1.f);
        const float pos = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(S - 1));
        int i0 = (int)floorf(pos);
        i0 = min(max(i0, 0), S - 1);
        const float w1 = pos - (float)i0, w0 = 1.f - w1;
        const float* line = p.code[v] + (size_t)c * S;
        float val = __ldg(line + i0) * w0;
        if (i0 + 1 <= S - 1) val += __ldg(line + i0 + 1) * w1;
        dst[r * P + m] = val;
    }
}

template <int TM, int P>
__device__ __forceinline__ void zero_rows(float* dst, int r0, int r1) {
    for (int idx = threadIdx.x; idx < (r1 - r0) * TM; idx += FT) {
        const int r = idx / TM, m = idx - r * TM;
        dst[(r0 + r) * P + m] = 0.f;
    }
}

// SDF-net input at point sp: rows 0..38 freq, 39..70 grid, 71..72 topo, 73..79 zero
template <int TM, int P>
__device__ __forceinline__ void build_sdf_input(const mb_field_params& p, const GridCtx& g, const float* sp, const float* stopo,
                                                float* dst) {
    build_freq<TM, P>(sp, dst, (int)p.n_freq);
    build_grid<TM, P>(g, sp, dst + 39 * P);
    for (int idx = threadIdx.x; idx < 9 * TM; idx += FT) {
        const int r = idx / TM, m = idx - r * TM;
        dst[(71 + r) * P + m] = (r < 2) ? stopo[r * P + m] : 0.f;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(FT, 2) field_fwd_kernel(const mb_field_params p, const mb_field_io io) {
    constexpr int TM = FWD_TM, P = FWD_P;
    extern __shared__ __align__(16) float sm[];
    float* in0 = sm + FwdSmem::IN0;
    float* bufA = sm + FwdSmem::BUFA;
    float* bufB = sm + FwdSmem::BUFB;
    float* wbuf = sm + FwdSmem::WBUF;
    float* sx = sm + FwdSmem::SX;
    float* sxw = sm + FwdSmem::SXW;
    float* spt = sm + FwdSmem::SPT;
    float* stopo = sm + FwdSmem::STOPO;
    float* sdef = sm + FwdSmem::SDEF;
    float* ssdf = sm + FwdSmem::SSDF;
    float* sq = sm + FwdSmem::SQ;
    float* st = sm + FwdSmem::ST;
    float* salb = sm + FwdSmem::SALB;
    const int tid = threadIdx.x;
    const float* A = p.arena;
    const uint32_t flags = io.flags;
    __shared__ LevelInfo s_levels[16];
    if (p.offsets) init_levels(s_levels, p.offsets, p.S, p.H);
    __syncthreads();
    const GridCtx gs{p.emb_sdf, p.offsets, p.S, p.H, p.n_levels, p.bound, p.two_bound, s_levels};
    const GridCtx gc{p.emb_col, p.offsets, p.S, p.H, p.n_levels, p.bound, p.two_bound, s_levels};

    const uint32_t n_tiles = div_up(io.M, TM);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t m0 = tile * TM;
        const int nv = (int)min((uint32_t)TM, io.M - m0);
        // ---- load ----
        for (int idx = tid; idx < 3 * TM; idx += FT) {
            const int m = idx / 3, a = idx - m * 3;
            sx[a * P + m] = (m < nv) ? io.x[(size_t)m0 * 3 + idx] : 0.f;
        }
        if (tid < TM) {
            st[tid] = (io.t && tid < nv) ? io.t[m0 + tid] : 0.f;
            stopo[tid] = ((flags & MB_F_TOPO_IN) && tid < nv) ? io.topo_in[(size_t)(m0 + tid) * 2] : 0.f;
            stopo[P + tid] = ((flags & MB_F_TOPO_IN) && tid < nv) ? io.topo_in[(size_t)(m0 + tid) * 2 + 1] : 0.f;
            sdef[tid] = sdef[P + tid] = sdef[2 * P + tid] = 0.f;
        }
        __syncthreads();
        // ---- warp: deformation + topology networks ----
        if (flags & MB_F_WARP) {
            build_freq<TM, P>(sx, in0, (int)p.n_freq);
            build_code<TM, P>(p, st, in0 + 39 * P);
            zero_rows<TM, P>(in0, 87, 96);
            __syncthreads();
            for (int net = 0; net < 2; net++) {
                const mb_layer_desc* L = net == 0 ? p.deform : p.topo;
                dense<TM, P, 128>(A + L[0].wt_off, A + L[0].b_off, 96, in0, bufA, wbuf, true, nullptr);
                dense<TM, P, 128>(A + L[1].wt_off, A + L[1].b_off, 128, bufA, bufB, wbuf, true, nullptr);
                dense<TM, P, 128>(A + L[2].wt_off, A + L[2].b_off, 128, bufB, bufA, wbuf, true, nullptr);
                dense<TM, P, 128>(A + L[3].wt_off, A + L[3].b_off, 128, bufA, bufB, wbuf, true, nullptr);
                dense<TM, P, 128>(A + L[4].wt_off, A + L[4].b_off, 128, bufB, bufA, wbuf, true, nullptr);
                dense<TM, P, 16>(A + L[5].wt_off, A + L[5].b_off, 128, bufA, bufB, wbuf, false, nullptr);
                if (tid < TM) {
                    if (net == 0) {
                        sdef[tid] = bufB[tid]; sdef[P + tid] = bufB[P + tid]; sdef[2 * P + tid] = bufB[2 * P + tid];
                    } else {
                        stopo[tid] = bufB[tid]; stopo[P + tid] = bufB[P + tid];
                    }
                }
                __syncthreads();
            }
        }
        for (int idx = tid; idx < 3 * TM; idx += FT) {
            const int a = idx / TM, m = idx - a * TM;
            sxw[a * P + m] = sx[a * P + m] + sdef[a * P + m];
        }
        __syncthreads();
        // ---- main query ----
        if (flags & MB_F_MAIN) {
            build_sdf_input<TM, P>(p, gs, sxw, stopo, in0);
            dense<TM, P, 64>(A + p.sdf[0].wt_off, A + p.sdf[0].b_off, 80, in0, bufA, wbuf, true, nullptr);
            dense<TM, P, 64>(A + p.sdf[1].wt_off, A + p.sdf[1].b_off, 64, bufA, bufB, wbuf, true, nullptr);
            dense<TM, P, 48>(A + p.sdf[2].wt_off, A + p.sdf[2].b_off, 64, bufB, bufA, wbuf, false, nullptr);
            if (tid < TM) ssdf[tid] = bufA[tid];
            if (flags & MB_F_COLOR) {
                // colour input: rows 0..31 colour-grid features, rows 32..63 = h[1..32]
                build_grid<TM, P>(gc, sxw, bufB);
                for (int idx = tid; idx < 32 * TM; idx += FT) {
                    const int r = idx / TM, m = idx - r * TM;
                    bufB[(32 + r) * P + m] = bufA[(1 + r) * P + m];
                }
                __syncthreads();
                dense<TM, P, 64>(A + p.color[0].wt_off, A + p.color[0].b_off, 64, bufB, bufA, wbuf, true, nullptr);
                dense<TM, P, 64>(A + p.color[1].wt_off, A + p.color[1].b_off, 64, bufA, bufB, wbuf, true, nullptr);
                dense<TM, P, 16>(A + p.color[2].wt_off, A + p.color[2].b_off, 64, bufB, bufA, wbuf, false, nullptr);
                for (int idx = tid; idx < 3 * TM; idx += FT) {
                    const int a = idx / TM, m = idx - a * TM;
                    salb[a * P + m] = 1.0f / (1.0f + expf(-bufA[a * P + m]));
                }
            }
            __syncthreads();
        }
        // ---- finite-difference normal ----
        if (flags & MB_F_FD) {
            const float* pt = (flags & MB_F_FD_WARPED) ? sxw : sx;
            for (int q = 0; q < 6; q++) {
                const int axis = q >> 1;
                const float e = (q & 1) ? -FD_EPS : FD_EPS;
                for (int idx = tid; idx < 3 * TM; idx += FT) {
                    const int a = idx / TM, m = idx - a * TM;
                    float v = pt[a * P + m];
                    if (a == axis) v = __fadd_rn(v, e);
                    spt[a * P + m] = fminf(fmaxf(v, -p.bound), p.bound);   // model.py:372 clamp
                }
                __syncthreads();
                build_sdf_input<TM, P>(p, gs, spt, stopo, in0);
                dense<TM, P, 64>(A + p.sdf[0].wt_off, A + p.sdf[0].b_off, 80, in0, bufA, wbuf, true, nullptr);
                dense<TM, P, 64>(A + p.sdf[1].wt_off, A + p.sdf[1].b_off, 64, bufA, bufB, wbuf, true, nullptr);
                dense_row0<TM, P>(A + p.sdf[2].w_off, __ldg(A + p.sdf[2].b_off), 64, bufB, sq + q * P);
            }
        }
        // ---- epilogue: per-sample outputs, coalesced ----
        if (tid < nv) {
            const int m = tid;
            const uint32_t gm = m0 + m;
            if (flags & MB_F_MAIN) {
                const float s = ssdf[m];
                if (io.sdf) io.sdf[gm] = s;
                if (io.sigma) io.sigma[gm] = laplace_sigma(s, __ldg(p.beta));
            }
            float n[3] = {0.f, 0.f, 0.f};
            if (flags & MB_F_FD) {
                float raw[3];
#pragma unroll
                for (int a = 0; a < 3; a++) raw[a] = __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(sq[(2 * a) * P + m], sq[(2 * a + 1) * P + m])), FD_EPS);
                const float d2 = raw[0] * raw[0] + raw[1] * raw[1] + raw[2] * raw[2];
                const float inv = 1.0f / sqrtf(fmaxf(d2, 1e-20f));   // utils.py:70-71
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    float v = raw[a] * inv;
                    if (isnan(v)) v = 0.f;                             // torch.nan_to_num (model.py:397)
                    else if (isinf(v)) v = v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
                    n[a] = v;
                    if (io.normal) io.normal[(size_t)gm * 3 + a] = v;
                    if (io.normal_raw) io.normal_raw[(size_t)gm * 3 + a] = raw[a];
                }
            }
            if (io.color) {
                float c[3] = {0.f, 0.f, 0.f};
                if (flags & MB_F_COLOR) { c[0] = salb[m]; c[1] = salb[P + m]; c[2] = salb[2 * P + m]; }
                if (io.shading != MB_SHADE_ALBEDO) {
                    float ndl = 0.f;
                    if (io.light) ndl = n[0] * io.light[(size_t)gm * 3] + n[1] * io.light[(size_t)gm * 3 + 1] + n[2] * io.light[(size_t)gm * 3 + 2];
                    const float lam = io.ratio + (1.0f - io.ratio) * fmaxf(ndl, 0.f);   // model.py:522
                    if (io.shading == MB_SHADE_TEXTURELESS) c[0] = c[1] = c[2] = lam;
                    else if (io.shading == MB_SHADE_NORMAL) { c[0] = (n[0] + 1.f) * 0.5f; c[1] = (n[1] + 1.f) * 0.5f; c[2] = (n[2] + 1.f) * 0.5f; }
                    else { c[0] *= lam; c[1] *= lam; c[2] *= lam; }
                }
                io.color[(size_t)gm * 3] = c[0]; io.color[(size_t)gm * 3 + 1] = c[1]; io.color[(size_t)gm * 3 + 2] = c[2];
            }
            if (io.deform) {
                io.deform[(size_t)gm * 3] = sdef[m]; io.deform[(size_t)gm * 3 + 1] = sdef[P + m]; io.deform[(size_t)gm * 3 + 2] = sdef[2 * P + m];
            }
            if (io.topo) { io.topo[(size_t)gm * 2] = stopo[m]; io.topo[(size_t)gm * 2 + 1] = stopo[P + m]; }
        }
        __syncthreads();
    }
}

}  // namespace mb

extern "C" int mb_field_forward(const mb_field_params* p, const mb_field_io* io, mb_stream_t stream) {
    using namespace mb;
    if (!p || !io) { set_error("field_forward: null argument"); return MB_EINVAL; }
    if (io->M == 0) return MB_OK;
    if (!io->x || !p->arena) { set_error("field_forward: x/arena is null"); return MB_EINVAL; }
    if ((io->flags & MB_F_WARP) && !io->t) { set_error("field_forward: WARP needs t"); return MB_EINVAL; }
    if ((io->flags & MB_F_TOPO_IN) && !io->topo_in) { set_error("field_forward: TOPO_IN needs topo_in"); return MB_EINVAL; }
    if ((io->flags & MB_F_COLOR) && !(io->flags & MB_F_MAIN)) { set_error("field_forward: COLOR needs MAIN"); return MB_EINVAL; }
    if ((io->flags & (MB_F_MAIN | MB_F_FD)) && (!p->emb_sdf || !p->offsets)) { set_error("field_forward: missing SDF grid"); return MB_EINVAL; }
    if (io->shading != MB_SHADE_ALBEDO && !(io->flags & MB_F_FD)) { set_error("field_forward: shading needs FD normals"); return MB_EINVAL; }
    constexpr size_t smem = (size_t)FwdSmem::TOTAL * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(field_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("field_forward: cannot reserve %zu B smem: %s", smem, cudaGetErrorString(e)); return MB_ECUDA; }
        attr_set = true;
    }
    const uint32_t n_tiles = div_up(io->M, FWD_TM);
    const uint32_t grid = min(n_tiles, (uint32_t)mb_sm_count() * 2u);   // persistent: 2 CTAs per SM
    field_fwd_kernel<<<grid, FT, smem, (cudaStream_t)stream>>>(*p, *io);
    return check_launch("field_forward");
}
