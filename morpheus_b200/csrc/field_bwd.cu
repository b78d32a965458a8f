// Fused scene-field backward (recompute): per tile of 32 samples the forward activations are rebuilt
// in shared memory, then every layer's data gradient (dgrad) and weight gradient (wgrad) is formed
// on chip.  Outputs: one flat weight-gradient arena (same offsets as the parameter arena;
// red.global.add), hash-table gradients (red.v2), deformation-code gradients, d/dbeta, d/dx.
// Mirrors torch.autograd through models/model.py:273-307,367-398,412-437,483-533 and
// _grid_encode.backward (grid.py:75-96); summation order of the atomics is unspecified exactly as
// in the reference's kernel_grid_backward (gridencoder.cu:345).
#include "field_common.cuh"

namespace mb {

constexpr int BWD_TM = 32;
constexpr int BWD_P = 36;

// SMALL = the deform/topology backward is not done here (no WARP, or MB_F_SKIP_WARP_BWD with the tensor-core kernel):
// only the SDF / colour activations (400 rows) and 80-row gradient buffers are needed -> 99 KB -> 2 CTAs per SM.
template <bool SMALL>
struct BwdSmemT {
    static constexpr int P = BWD_P;
    static constexpr int STORE_ROWS = SMALL ? 400 : 640;
    static constexpr int IN_ROWS = SMALL ? 0 : 96;
    static constexpr int DZ_ROWS = SMALL ? 80 : 128;
    static constexpr int STORE = 0;                    // stored activations
    static constexpr int IN0 = STORE + STORE_ROWS * P;
    static constexpr int GIN0 = IN0 + IN_ROWS * P;
    static constexpr int DZA = GIN0 + IN_ROWS * P;
    static constexpr int DZB = DZA + DZ_ROWS * P;
    static constexpr int WBUF = DZB + DZ_ROWS * P;
    static constexpr int SX = WBUF + WBUF_FLOATS;      // [3][P]
    static constexpr int SXW = SX + 3 * P;
    static constexpr int SPT = SXW + 3 * P;
    static constexpr int STOPO = SPT + 3 * P;          // [2][P]
    static constexpr int SDEF = STOPO + 2 * P;         // [3][P]
    static constexpr int ST = SDEF + 3 * P;            // [1][P]
    static constexpr int SSDF = ST + P;                // [1][P]
    static constexpr int SALB = SSDF + P;              // [3][P]
    static constexpr int GXW = SALB + 3 * P;           // [3][P] grad wrt x + deform
    static constexpr int GX = GXW + 3 * P;             // [3][P] grad wrt x
    static constexpr int GPT = GX + 3 * P;             // [3][P] grad wrt the current FD point
    static constexpr int GTOPO = GPT + 3 * P;          // [2][P]
    static constexpr int GSQ = GTOPO + 2 * P;          // [6][P] grad wrt the six FD sdf values
    static constexpr int GALB = GSQ + 6 * P;           // [3][P] grad wrt albedo (pre-sigmoid applied later)
    static constexpr int RED = GALB + 3 * P;           // [8] block-reduce scratch
    static constexpr int TOTAL = RED + 8;
};

// S0 offsets inside STORE for the SDF / colour phase
constexpr int R_S0 = 0, R_S1 = 80, R_S2 = 144, R_C0 = 208, R_C1 = 272, R_C2 = 336;

template <int TM, int P>
__device__ __forceinline__ void build_code_b(const mb_field_params& p, const float* __restrict__ st, float* __restrict__ dst) {
    for (int idx = threadIdx.x; idx < 48 * TM; idx += FT) {
        const int r = idx / TM, m = idx - r * TM;
        const int v = r >> 4, c = r & 15;
        const int S = (int)p.code_len[v];
        const float t = fminf(fmaxf(st[m], 0.f), 1.f);
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
__device__ __forceinline__ void build_sdf_input_b(const mb_field_params& p, const GridCtx& g, const float* sp, const float* stopo,
                                                  bool use_topo, float* dst) {
    build_freq<TM, P>(sp, dst, (int)p.n_freq);
    build_grid<TM, P>(g, sp, dst + 39 * P);
    for (int idx = threadIdx.x; idx < 9 * TM; idx += FT) {
        const int r = idx / TM, m = idx - r * TM;
        dst[(71 + r) * P + m] = (r < 2 && use_topo) ? stopo[r * P + m] : 0.f;
    }
    __syncthreads();
}

template <int TM, int P>
__device__ __forceinline__ void zero_rows_b(float* dst, int r0, int r1) {
    for (int idx = threadIdx.x; idx < (r1 - r0) * TM; idx += FT) {
        const int r = idx / TM, m = idx - r * TM;
        dst[(r0 + r) * P + m] = 0.f;
    }
}

// backward through the 3-layer SDF net whose stored activations are S0,S1,S2 and whose output
// gradient (N_pad rows: 48 for the main query, 16 for an FD query that only uses row 0) is in dza.
// Leaves dS0 [80][P] in dzb.
template <int TM, int P, int NOUT /*48 or 16*/>
__device__ __forceinline__ void sdf_net_backward(const mb_field_params& p, const float* A, float* GA, float* store, float* dza,
                                                 float* dzb, float* wbuf) {
    const mb_layer_desc& L0 = p.sdf[0];
    const mb_layer_desc& L1 = p.sdf[1];
    const mb_layer_desc& L2 = p.sdf[2];
    wgrad<TM, P, 4, NOUT / 16>(store + R_S2 * P, dza, 64, NOUT == 48 ? 33 : 1, (int)L2.N_pad, GA + L2.wt_off, GA + L2.b_off);
    dense<TM, P, 64>(A + L2.w_off, nullptr, NOUT, dza, dzb, wbuf, false, store + R_S2 * P);
    wgrad<TM, P, 4, 4>(store + R_S1 * P, dzb, 64, 64, (int)L1.N_pad, GA + L1.wt_off, GA + L1.b_off);
    dense<TM, P, 64>(A + L1.w_off, nullptr, 64, dzb, dza, wbuf, false, store + R_S1 * P);
    wgrad<TM, P, 5, 4>(store + R_S0 * P, dza, 73, 64, (int)L0.N_pad, GA + L0.wt_off, GA + L0.b_off);
    dense<TM, P, 80>(A + L0.w_off, nullptr, 64, dza, dzb, wbuf, false, nullptr);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    if (threadIdx.x < 8) s = red[threadIdx.x];
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    __syncthreads();
    return s;  // valid in thread 0
}

template <bool SMALL>
__global__ void __launch_bounds__(FT, SMALL ? 2 : 1) field_bwd_kernel(const mb_field_params p, const mb_field_io io, const mb_field_grads gr) {
    constexpr int TM = BWD_TM, P = BWD_P;
    using BwdSmem = BwdSmemT<SMALL>;
    extern __shared__ __align__(16) float sm[];
    float* store = sm + BwdSmem::STORE;
    float* in0 = sm + BwdSmem::IN0;
    float* gin0 = sm + BwdSmem::GIN0;
    float* dza = sm + BwdSmem::DZA;
    float* dzb = sm + BwdSmem::DZB;
    float* wbuf = sm + BwdSmem::WBUF;
    float* sx = sm + BwdSmem::SX;
    float* sxw = sm + BwdSmem::SXW;
    float* spt = sm + BwdSmem::SPT;
    float* stopo = sm + BwdSmem::STOPO;
    float* sdef = sm + BwdSmem::SDEF;
    float* st = sm + BwdSmem::ST;
    float* ssdf = sm + BwdSmem::SSDF;
    float* salb = sm + BwdSmem::SALB;
    float* gxw = sm + BwdSmem::GXW;
    float* gx = sm + BwdSmem::GX;
    float* gpt = sm + BwdSmem::GPT;
    float* gtopo = sm + BwdSmem::GTOPO;
    float* gsq = sm + BwdSmem::GSQ;
    float* galb = sm + BwdSmem::GALB;
    float* red = sm + BwdSmem::RED;
    const int tid = threadIdx.x;
    const float* A = p.arena;
    float* GA = gr.g_arena;
    const uint32_t flags = io.flags;
    const bool topo_live = (flags & (MB_F_WARP | MB_F_TOPO_IN)) != 0;
    __shared__ LevelInfo s_levels[16];
    if (p.offsets) init_levels(s_levels, p.offsets, p.S, p.H);
    __syncthreads();
    const GridCtx gs{p.emb_sdf, p.offsets, p.S, p.H, p.n_levels, p.bound, p.two_bound, s_levels};
    const GridCtx gc{p.emb_col, p.offsets, p.S, p.H, p.n_levels, p.bound, p.two_bound, s_levels};

    const uint32_t n_tiles = div_up(io.M, TM);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t m0 = tile * TM;
        const int nv = (int)min((uint32_t)TM, io.M - m0);
        // ---- load inputs, saved forward values; clear gradient accumulators ----
        for (int idx = tid; idx < 3 * TM; idx += FT) {
            const int m = idx / 3, a = idx - m * 3;
            const bool ok = m < nv;
            sx[a * P + m] = ok ? io.x[(size_t)m0 * 3 + idx] : 0.f;
            sdef[a * P + m] = (ok && (flags & MB_F_WARP)) ? gr.deform[(size_t)m0 * 3 + idx] : 0.f;
            gxw[a * P + m] = 0.f;
            gx[a * P + m] = 0.f;
            galb[a * P + m] = 0.f;
        }
        for (int idx = tid; idx < 2 * TM; idx += FT) {
            const int m = idx / 2, a = idx - m * 2;
            const bool ok = m < nv;
            float v = 0.f;
            if (ok && (flags & MB_F_WARP)) v = gr.topo[(size_t)m0 * 2 + idx];
            else if (ok && (flags & MB_F_TOPO_IN)) v = io.topo_in[(size_t)m0 * 2 + idx];
            stopo[a * P + m] = v;
            gtopo[a * P + m] = (ok && gr.g_topo) ? gr.g_topo[(size_t)m0 * 2 + idx] : 0.f;
        }
        for (int idx = tid; idx < 6 * TM; idx += FT) gsq[(idx / TM) * P + idx % TM] = 0.f;
        if (tid < TM) st[tid] = (io.t && tid < nv) ? io.t[m0 + tid] : 0.f;
        __syncthreads();
        for (int idx = tid; idx < 3 * TM; idx += FT) {
            const int a = idx / TM, m = idx - a * TM;
            sxw[a * P + m] = sx[a * P + m] + sdef[a * P + m];
        }
        __syncthreads();

        // ---- upstream -> local gradients (albedo, FD sdf values) ----
        const bool need_fd = (flags & MB_F_FD) && (gr.g_normal || gr.g_normal_raw || (io.shading != MB_SHADE_ALBEDO && gr.g_color));
        // main forward first when colour is needed for the shading product rule
        bool main_done = false;
        auto main_forward = [&]() {
            build_sdf_input_b<TM, P>(p, gs, sxw, stopo, true, store + R_S0 * P);
            dense<TM, P, 64>(A + p.sdf[0].wt_off, A + p.sdf[0].b_off, 80, store + R_S0 * P, store + R_S1 * P, wbuf, true, nullptr);
            dense<TM, P, 64>(A + p.sdf[1].wt_off, A + p.sdf[1].b_off, 64, store + R_S1 * P, store + R_S2 * P, wbuf, true, nullptr);
            dense<TM, P, 48>(A + p.sdf[2].wt_off, A + p.sdf[2].b_off, 64, store + R_S2 * P, dza, wbuf, false, nullptr);
            if (tid < TM) ssdf[tid] = dza[tid];
            if (flags & MB_F_COLOR) {
                build_grid<TM, P>(gc, sxw, store + R_C0 * P);
                for (int idx = tid; idx < 32 * TM; idx += FT) {
                    const int r = idx / TM, m = idx - r * TM;
                    store[(R_C0 + 32 + r) * P + m] = dza[(1 + r) * P + m];
                }
                __syncthreads();
                dense<TM, P, 64>(A + p.color[0].wt_off, A + p.color[0].b_off, 64, store + R_C0 * P, store + R_C1 * P, wbuf, true, nullptr);
                dense<TM, P, 64>(A + p.color[1].wt_off, A + p.color[1].b_off, 64, store + R_C1 * P, store + R_C2 * P, wbuf, true, nullptr);
                dense<TM, P, 16>(A + p.color[2].wt_off, A + p.color[2].b_off, 64, store + R_C2 * P, dzb, wbuf, false, nullptr);
                for (int idx = tid; idx < 3 * TM; idx += FT) {
                    const int a = idx / TM, m = idx - a * TM;
                    salb[a * P + m] = 1.0f / (1.0f + expf(-dzb[a * P + m]));
                }
            }
            __syncthreads();
            main_done = true;
        };
        if (flags & MB_F_MAIN) main_forward();

        if (tid < nv) {
            const int m = tid;
            const uint32_t gm = m0 + m;
            float gc3[3] = {0.f, 0.f, 0.f};
            if (gr.g_color) { gc3[0] = gr.g_color[(size_t)gm * 3]; gc3[1] = gr.g_color[(size_t)gm * 3 + 1]; gc3[2] = gr.g_color[(size_t)gm * 3 + 2]; }
            float gn[3] = {0.f, 0.f, 0.f};
            if (gr.g_normal) { gn[0] = gr.g_normal[(size_t)gm * 3]; gn[1] = gr.g_normal[(size_t)gm * 3 + 1]; gn[2] = gr.g_normal[(size_t)gm * 3 + 2]; }
            float ga[3] = {gc3[0], gc3[1], gc3[2]};
            float n[3] = {0.f, 0.f, 0.f}, inv = 0.f;
            bool clamped = false;
            if (flags & MB_F_FD) {
                const float r0 = gr.normal_raw[(size_t)gm * 3], r1 = gr.normal_raw[(size_t)gm * 3 + 1], r2 = gr.normal_raw[(size_t)gm * 3 + 2];
                const float d2 = r0 * r0 + r1 * r1 + r2 * r2;
                clamped = !(d2 > 1e-20f);
                inv = 1.0f / sqrtf(fmaxf(d2, 1e-20f));
                n[0] = r0 * inv; n[1] = r1 * inv; n[2] = r2 * inv;
            }
            if (io.shading != MB_SHADE_ALBEDO) {
                float l[3] = {0.f, 0.f, 0.f};
                if (io.light) { l[0] = io.light[(size_t)gm * 3]; l[1] = io.light[(size_t)gm * 3 + 1]; l[2] = io.light[(size_t)gm * 3 + 2]; }
                const float ndl = n[0] * l[0] + n[1] * l[1] + n[2] * l[2];
                const float lam = io.ratio + (1.0f - io.ratio) * fmaxf(ndl, 0.f);
                float glam = 0.f;
                if (io.shading == MB_SHADE_LAMBERTIAN) {
                    const float a0 = (flags & MB_F_COLOR) ? salb[m] : 0.f, a1 = (flags & MB_F_COLOR) ? salb[P + m] : 0.f, a2 = (flags & MB_F_COLOR) ? salb[2 * P + m] : 0.f;
                    glam = gc3[0] * a0 + gc3[1] * a1 + gc3[2] * a2;
                    ga[0] = gc3[0] * lam; ga[1] = gc3[1] * lam; ga[2] = gc3[2] * lam;
                } else if (io.shading == MB_SHADE_TEXTURELESS) {
                    glam = gc3[0] + gc3[1] + gc3[2];
                    ga[0] = ga[1] = ga[2] = 0.f;
                } else {  // normal
                    gn[0] += 0.5f * gc3[0]; gn[1] += 0.5f * gc3[1]; gn[2] += 0.5f * gc3[2];
                    ga[0] = ga[1] = ga[2] = 0.f;
                }
                const float k = (ndl > 0.f) ? glam * (1.0f - io.ratio) : 0.f;
                gn[0] += k * l[0]; gn[1] += k * l[1]; gn[2] += k * l[2];
            }
            galb[m] = ga[0]; galb[P + m] = ga[1]; galb[2 * P + m] = ga[2];
            if (flags & MB_F_FD) {
                float graw[3];
                if (clamped) {
                    graw[0] = gn[0] * inv; graw[1] = gn[1] * inv; graw[2] = gn[2] * inv;
                } else {
                    const float dot = n[0] * gn[0] + n[1] * gn[1] + n[2] * gn[2];
                    graw[0] = inv * (gn[0] - n[0] * dot); graw[1] = inv * (gn[1] - n[1] * dot); graw[2] = inv * (gn[2] - n[2] * dot);
                }
                if (gr.g_normal_raw) { graw[0] += gr.g_normal_raw[(size_t)gm * 3]; graw[1] += gr.g_normal_raw[(size_t)gm * 3 + 1]; graw[2] += gr.g_normal_raw[(size_t)gm * 3 + 2]; }
                const float h = 0.5f / FD_EPS;
#pragma unroll
                for (int a = 0; a < 3; a++) { gsq[(2 * a) * P + m] = graw[a] * h; gsq[(2 * a + 1) * P + m] = -graw[a] * h; }
            }
        }
        __syncthreads();

        // ---- colour + SDF networks at x' ----
        if (main_done) {
            const bool have_color_grad = (flags & MB_F_COLOR) && gr.g_color && io.shading != MB_SHADE_TEXTURELESS && io.shading != MB_SHADE_NORMAL;
            if (have_color_grad) {
                for (int idx = tid; idx < 16 * TM; idx += FT) {
                    const int r = idx / TM, m = idx - r * TM;
                    float v = 0.f;
                    if (r < 3) { const float s = salb[r * P + m]; v = galb[r * P + m] * s * (1.0f - s); }
                    dza[r * P + m] = v;
                }
                __syncthreads();
                const mb_layer_desc& C0 = p.color[0];
                const mb_layer_desc& C1 = p.color[1];
                const mb_layer_desc& C2 = p.color[2];
                wgrad<TM, P, 4, 1>(store + R_C2 * P, dza, 64, 3, (int)C2.N_pad, GA + C2.wt_off, GA + C2.b_off);
                dense<TM, P, 64>(A + C2.w_off, nullptr, 16, dza, dzb, wbuf, false, store + R_C2 * P);
                wgrad<TM, P, 4, 4>(store + R_C1 * P, dzb, 64, 64, (int)C1.N_pad, GA + C1.wt_off, GA + C1.b_off);
                dense<TM, P, 64>(A + C1.w_off, nullptr, 64, dzb, dza, wbuf, false, store + R_C1 * P);
                wgrad<TM, P, 4, 4>(store + R_C0 * P, dza, 64, 64, (int)C0.N_pad, GA + C0.wt_off, GA + C0.b_off);
                dense<TM, P, 64>(A + C0.w_off, nullptr, 64, dza, dzb, wbuf, false, nullptr);   // dC0 in dzb
                grid_backward<TM, P>(gc, sxw, dzb, gr.g_emb_col, gxw);
                __syncthreads();
            }
            // dH: row 0 from sdf/sigma, rows 1..32 from the colour net
            float gbeta_local = 0.f;
            for (int idx = tid; idx < 48 * TM; idx += FT) {
                const int r = idx / TM, m = idx - r * TM;
                float v = 0.f;
                if (r == 0) {
                    const uint32_t gm = m0 + m;
                    if (m < nv) {
                        if (gr.g_sdf) v = gr.g_sdf[gm];
                        if (gr.g_sigma) {
                            const float s = ssdf[m], b = __ldg(p.beta);
                            const float a = fabsf(s), e = expf(-a / b);
                            const float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
                            const float gsig = gr.g_sigma[gm];
                            v += gsig * (-0.5f * sg * sg * e / (b * b));
                            const float sigma = (1.0f / b) * (0.5f + 0.5f * sg * expm1f(-a / b));
                            gbeta_local += gsig * (-sigma / b + 0.5f * sg * e * a / (b * b * b));
                        }
                    }
                } else if (r <= 32 && have_color_grad) {
                    v = dzb[(31 + r) * P + m];
                }
                dza[r * P + m] = v;   // NB: dza != dzb, rows of dzb are only read
            }
            __syncthreads();
            if (gr.g_sigma && gr.g_beta) {
                const float tot = block_sum(gbeta_local, red);
                if (tid == 0 && tot != 0.f) atomicAdd(gr.g_beta, tot);
            }
            sdf_net_backward<TM, P, 48>(p, A, GA, store, dza, dzb, wbuf);
            freq_backward<TM, P>(sxw, dzb, gxw, (int)p.n_freq);
            grid_backward<TM, P>(gs, sxw, dzb + 39 * P, gr.g_emb_sdf, gxw);
            if (topo_live)
                for (int idx = tid; idx < 2 * TM; idx += FT) gtopo[(idx / TM) * P + idx % TM] += dzb[(71 + idx / TM) * P + idx % TM];
            __syncthreads();
        }

        // ---- finite-difference normal: six SDF queries, each recomputed then back-propagated ----
        if (need_fd) {
            const float* pt = (flags & MB_F_FD_WARPED) ? sxw : sx;
            float* gdst = (flags & MB_F_FD_WARPED) ? gxw : gx;
            for (int q = 0; q < 6; q++) {
                const int axis = q >> 1;
                const float e = (q & 1) ? -FD_EPS : FD_EPS;
                for (int idx = tid; idx < 3 * TM; idx += FT) {
                    const int a = idx / TM, m = idx - a * TM;
                    float v = pt[a * P + m];
                    if (a == axis) v = __fadd_rn(v, e);
                    spt[a * P + m] = fminf(fmaxf(v, -p.bound), p.bound);
                    gpt[a * P + m] = 0.f;
                }
                __syncthreads();
                build_sdf_input_b<TM, P>(p, gs, spt, stopo, true, store + R_S0 * P);
                dense<TM, P, 64>(A + p.sdf[0].wt_off, A + p.sdf[0].b_off, 80, store + R_S0 * P, store + R_S1 * P, wbuf, true, nullptr);
                dense<TM, P, 64>(A + p.sdf[1].wt_off, A + p.sdf[1].b_off, 64, store + R_S1 * P, store + R_S2 * P, wbuf, true, nullptr);
                for (int idx = tid; idx < 16 * TM; idx += FT) {
                    const int r = idx / TM, m = idx - r * TM;
                    dza[r * P + m] = (r == 0) ? gsq[q * P + m] : 0.f;
                }
                __syncthreads();
                sdf_net_backward<TM, P, 16>(p, A, GA, store, dza, dzb, wbuf);
                freq_backward<TM, P>(spt, dzb, gpt, (int)p.n_freq);
                grid_backward<TM, P>(gs, spt, dzb + 39 * P, gr.g_emb_sdf, gpt);
                if (topo_live)
                    for (int idx = tid; idx < 2 * TM; idx += FT) gtopo[(idx / TM) * P + idx % TM] += dzb[(71 + idx / TM) * P + idx % TM];
                __syncthreads();
                for (int idx = tid; idx < 3 * TM; idx += FT) {
                    const int a = idx / TM, m = idx - a * TM;
                    float v = pt[a * P + m];
                    if (a == axis) v = __fadd_rn(v, e);
                    if (v >= -p.bound && v <= p.bound) gdst[a * P + m] += gpt[a * P + m];   // clamp derivative
                }
                __syncthreads();
            }
        }

        // ---- deformation / topology networks ----
        if ((flags & MB_F_WARP) && (flags & MB_F_SKIP_WARP_BWD)) {
            // hand d/d(deform), d/d(topo) to the tensor-core backward of the two big networks
            for (int idx = tid; idx < 3 * TM; idx += FT) {
                const int m = idx / 3, a = idx - m * 3;
                if (m < nv) {
                    float v = gxw[a * P + m];
                    if (gr.g_deform) v += gr.g_deform[(size_t)m0 * 3 + idx];
                    gr.g_def_out[(size_t)m0 * 3 + idx] = v;
                }
            }
            for (int idx = tid; idx < 2 * TM; idx += FT) {
                const int m = idx / 2, a = idx - m * 2;
                if (m < nv) gr.g_topo_out[(size_t)m0 * 2 + idx] = gtopo[a * P + m];
            }
        } else if (!SMALL && (flags & MB_F_WARP)) {
          if constexpr (!SMALL) {
            build_freq<TM, P>(sx, in0, (int)p.n_freq);
            build_code_b<TM, P>(p, st, in0 + 39 * P);
            zero_rows_b<TM, P>(in0, 87, 96);
            zero_rows_b<TM, P>(gin0, 0, 96);
            __syncthreads();
            for (int net = 0; net < 2; net++) {
                const mb_layer_desc* L = net == 0 ? p.deform : p.topo;
                const int nout = net == 0 ? 3 : 2;
                dense<TM, P, 128>(A + L[0].wt_off, A + L[0].b_off, 96, in0, store, wbuf, true, nullptr);
                for (int l = 1; l < 5; l++)
                    dense<TM, P, 128>(A + L[l].wt_off, A + L[l].b_off, 128, store + (l - 1) * 128 * P, store + l * 128 * P, wbuf, true, nullptr);
                for (int idx = tid; idx < 16 * TM; idx += FT) {
                    const int r = idx / TM, m = idx - r * TM;
                    float v = 0.f;
                    if (r < nout) {
                        if (net == 0) {
                            v = gxw[r * P + m];
                            if (gr.g_deform && m < nv) v += gr.g_deform[(size_t)(m0 + m) * 3 + r];
                        } else {
                            v = gtopo[r * P + m];
                        }
                    }
                    dza[r * P + m] = v;
                }
                __syncthreads();
                wgrad<TM, P, 8, 1>(store + 4 * 128 * P, dza, 128, nout, (int)L[5].N_pad, GA + L[5].wt_off, GA + L[5].b_off);
                dense<TM, P, 128>(A + L[5].w_off, nullptr, 16, dza, dzb, wbuf, false, store + 4 * 128 * P);
                float* cur = dzb;
                float* nxt = dza;
                for (int l = 4; l >= 1; l--) {
                    wgrad<TM, P, 8, 8>(store + (l - 1) * 128 * P, cur, 128, 128, (int)L[l].N_pad, GA + L[l].wt_off, GA + L[l].b_off);
                    dense<TM, P, 128>(A + L[l].w_off, nullptr, 128, cur, nxt, wbuf, false, store + (l - 1) * 128 * P);
                    float* tmp = cur; cur = nxt; nxt = tmp;
                }
                wgrad<TM, P, 6, 8>(in0, cur, 87, 128, (int)L[0].N_pad, GA + L[0].wt_off, GA + L[0].b_off);
                dense<TM, P, 96>(A + L[0].w_off, nullptr, 128, cur, nxt, wbuf, false, nullptr);
                for (int idx = tid; idx < 96 * TM; idx += FT) {
                    const int r = idx / TM, m = idx - r * TM;
                    gin0[r * P + m] += nxt[r * P + m];
                }
                __syncthreads();
            }
            freq_backward<TM, P>(sx, gin0, gx, (int)p.n_freq);
            // deformation-code gradient: rows 39..86 of gin0
            {
                const unsigned t0b = __float_as_uint(st[0]);
                const bool mine = (tid >= nv) || (tid < TM && __float_as_uint(st[tid]) == t0b) || tid >= TM;
                const int uniform = __syncthreads_and(mine ? 1 : 0);
                if (uniform) {
                    const int warp = tid >> 5, lane = tid & 31;
                    for (int r = warp; r < 48; r += FT / 32) {
                        float v = (lane < TM) ? gin0[(39 + r) * P + lane] : 0.f;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        if (lane == 0 && v != 0.f) {
                            const int vv = r >> 4, c = r & 15;
                            const int S = (int)p.code_len[vv];
                            const float t = fminf(fmaxf(st[0], 0.f), 1.f);
                            const float g = __fsub_rn(__fmul_rn(t, 2.f), 1.f);
                            const float pos = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(S - 1));
                            int i0 = min(max((int)floorf(pos), 0), S - 1);
                            const float w1 = pos - (float)i0, w0 = 1.f - w1;
                            float* gl = gr.g_code[vv] + (size_t)c * S;
                            atomicAdd(gl + i0, v * w0);
                            if (i0 + 1 <= S - 1) atomicAdd(gl + i0 + 1, v * w1);
                        }
                    }
                } else {
                    for (int idx = tid; idx < 48 * TM; idx += FT) {
                        const int r = idx / TM, m = idx - r * TM;
                        const float v = gin0[(39 + r) * P + m];
                        if (m >= nv || v == 0.f) continue;
                        const int vv = r >> 4, c = r & 15;
                        const int S = (int)p.code_len[vv];
                        const float t = fminf(fmaxf(st[m], 0.f), 1.f);
                        const float g = __fsub_rn(__fmul_rn(t, 2.f), 1.f);
                        const float pos = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(S - 1));
                        int i0 = min(max((int)floorf(pos), 0), S - 1);
                        const float w1 = pos - (float)i0, w0 = 1.f - w1;
                        float* gl = gr.g_code[vv] + (size_t)c * S;
                        atomicAdd(gl + i0, v * w0);
                        if (i0 + 1 <= S - 1) atomicAdd(gl + i0 + 1, v * w1);
                    }
                }
            }
            __syncthreads();
          }   // if constexpr (!SMALL)
        }
        // ---- outputs ----
        for (int idx = tid; idx < 3 * TM; idx += FT) {
            const int m = idx / 3, a = idx - m * 3;
            if (m < nv && gr.g_x) gr.g_x[(size_t)m0 * 3 + idx] = gx[a * P + m] + gxw[a * P + m];
        }
        if ((flags & MB_F_TOPO_IN) && gr.g_topo_in)
            for (int idx = tid; idx < 2 * TM; idx += FT) {
                const int m = idx / 2, a = idx - m * 2;
                if (m < nv) gr.g_topo_in[(size_t)m0 * 2 + idx] = gtopo[a * P + m];
            }
        __syncthreads();
    }
}

}  // namespace mb

extern "C" int mb_field_backward(const mb_field_params* p, const mb_field_io* io, const mb_field_grads* g, mb_stream_t stream) {
    using namespace mb;
    if (!p || !io || !g) { set_error("field_backward: null argument"); return MB_EINVAL; }
    if (io->M == 0) return MB_OK;
    if (!io->x || !p->arena || !g->g_arena) { set_error("field_backward: x/arena/g_arena is null"); return MB_EINVAL; }
    if ((io->flags & MB_F_SKIP_WARP_BWD) && (!g->g_def_out || !g->g_topo_out)) { set_error("field_backward: SKIP_WARP_BWD needs g_def_out/g_topo_out"); return MB_EINVAL; }
    if ((io->flags & MB_F_WARP) && (!io->t || !g->deform || !g->topo || !g->g_code[0] || !g->g_code[1] || !g->g_code[2])) {
        set_error("field_backward: WARP needs t, saved deform/topo and g_code");
        return MB_EINVAL;
    }
    if ((io->flags & MB_F_FD) && !g->normal_raw) { set_error("field_backward: FD needs the saved normal_raw"); return MB_EINVAL; }
    if ((io->flags & (MB_F_MAIN | MB_F_FD)) && !g->g_emb_sdf) { set_error("field_backward: g_emb_sdf is null"); return MB_EINVAL; }
    if ((io->flags & MB_F_COLOR) && !g->g_emb_col) { set_error("field_backward: g_emb_col is null"); return MB_EINVAL; }
    if ((io->flags & MB_F_TOPO_IN) && !io->topo_in) { set_error("field_backward: TOPO_IN needs topo_in"); return MB_EINVAL; }
    const uint32_t n_tiles = div_up(io->M, BWD_TM);
    const bool small = !(io->flags & MB_F_WARP) || (io->flags & MB_F_SKIP_WARP_BWD);
    static bool attr_set[2] = {false, false};
    if (small) {
        constexpr size_t smem = (size_t)BwdSmemT<true>::TOTAL * sizeof(float);
        if (!attr_set[1]) {
            cudaError_t e = cudaFuncSetAttribute(field_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { set_error("field_backward: cannot reserve %zu B smem: %s", smem, cudaGetErrorString(e)); return MB_ECUDA; }
            attr_set[1] = true;
        }
        const uint32_t grid = min(n_tiles, (uint32_t)mb_sm_count() * 2u);
        field_bwd_kernel<true><<<grid, FT, smem, (cudaStream_t)stream>>>(*p, *io, *g);
    } else {
        constexpr size_t smem = (size_t)BwdSmemT<false>::TOTAL * sizeof(float);
        if (!attr_set[0]) {
            cudaError_t e = cudaFuncSetAttribute(field_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { set_error("field_backward: cannot reserve %zu B smem: %s", smem, cudaGetErrorString(e)); return MB_ECUDA; }
            attr_set[0] = true;
        }
        const uint32_t grid = min(n_tiles, (uint32_t)mb_sm_count());
        field_bwd_kernel<false><<<grid, FT, smem, (cudaStream_t)stream>>>(*p, *io, *g);
    }
    return check_launch("field_backward");
}
