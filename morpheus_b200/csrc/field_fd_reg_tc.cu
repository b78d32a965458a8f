// Fused finite-difference normal regulariser of a real-view training step (tcgen05 + TMEM), forward AND backward in ONE pass.
//
// What it replaces (reference): on real views `scene_representation.forward(.., shading='albedo_normal')` evaluates the 6-point FD
// normal at every sample (models/model.py:367-398, :521) and `render_rays` evaluates it again at the perturbed point with topo = 0
// (morpheus.py:714-741): loss_normal_perturb = mean |n(x, topo) - n(x + delta, 0)|.  With ratio = 1 the colour does not depend on the
// normal (model.py:523-526), so these 12 SDF queries per sample feed nothing but this loss (and the returned `normal` tensors).  The
// general path of this library spends 2 forward launches (FD chains of the main query, the perturbed query) and 2 backward launches
// (field_bwd_fd_tc.cu, which recomputes the forward) on them: 4 gathers and 12 MMA round trips per (sample, set).
//
// Here one CTA sub-tile holds BOTH sets of 8 samples (96 rows = 8 samples x 2 sets x 6 queries, query index fastest), so the loss
// gradient is known as soon as the forward of the sub-tile is: gather -> L0 -> L1 -> (row-0 dot product = sdf) -> normals, loss,
// d loss / d sdf -> dZ1 straight from the layer-1 accumulator that is still in TMEM -> [wgrad1, dgrad1] -> [wgrad0, dgrad0] -> d(S0) ->
// table scatter.  One gather, 4 MMA round trips, no recompute, no per-sample intermediates in HBM.
//
// Shared-corner gathers / scatters: the six +-eps points of a (sample, set) lie within eps * res / (2 bound) <= 0.127 cells of the base
// point, so per level they touch the base cell's 8 corners plus, per axis, at most ONE far plane of 4 corners when the +eps (or -eps)
// point crosses a cell face.  A thread owns one (sample, set, level): 8 (+4 per crossing axis) corner loads feed all six interpolations
// (48 loads in the row-wise form), and the backward merges the six rows' contributions per corner in registers before the red.v2.
// The interpolation arithmetic (weight products in axis order, FMA chain over corners 0..7) is the one of grid_eval / the reference
// kernel (gridencoder.cu:171-195), so features are bit-identical to the row-wise gather.
//
// Gradient scaling for the fp16 (hi, lo) split: d loss / d sdf is only known per sub-tile, while the weight-gradient accumulators
// live in TMEM for a whole tile.  The scale (a power of two) is chosen at the first sub-tile with a non-zero gradient; if a later
// sub-tile would exceed the head-room (x 8 .. 16), the accumulators are flushed with the old scale and restart with a new one.
#include "field_common.cuh"
#include "tc_common.cuh"
#include "tc_field.cuh"

namespace mb {
namespace tcr {

using namespace mb::tc;

constexpr int TM = 128;                 // max samples per tile (one accumulator flush)
constexpr int SS = 8;                   // samples per sub-tile
constexpr int RT = 96;                  // rows per sub-tile = SS x 2 sets x 6 queries
constexpr int NWORK = 256;
constexpr int NTHREADS = NWORK + 64;    // + MMA-issue warp + weight-loader warp
constexpr int NSTAGE = 3;
constexpr int STAGE_BYTES = 5120;
constexpr int PITCH = RT * 16;          // bytes between 8-column core groups of a 96-row operand tile
constexpr int S0_LO = 10 * PITCH;       // lo offset of the 80-column S0 tile
constexpr int X_LO = 8 * PITCH;         // lo offset of a 64-column tile (dZ)
constexpr int X0_LO = 9 * PITCH;        // lo offset of the A1 tile: 64 columns + one core whose first column is the constant 1
#ifndef MB_FDR_PHASE_TIMING
#define MB_FDR_PHASE_TIMING 0           // 1: worker thread 0 accumulates clock64 deltas per phase (mb_debug_fdr_phases; a few % slower)
#endif
constexpr bool PHASE_TIMING = MB_FDR_PHASE_TIMING != 0;
__device__ unsigned long long g_fdr_phase[16];

struct Smem {
    static constexpr int S0 = 0;                          // 30720
    static constexpr int X0 = S0 + 2 * S0_LO;             // 27648
    static constexpr int DZ = X0 + 2 * X0_LO;             // 24576; G (fp32 [32][96]) aliases it after the last MMA of a sub-tile
    static constexpr int W = DZ + 2 * X_LO;               // NSTAGE x 5120
    static constexpr int F = W + NSTAGE * STAGE_BYTES;
    static constexpr int SPA = F;                         // [3][128] sample points x
    static constexpr int SPB = SPA + 3 * 512;             // [3][128] perturbed points x + delta
    static constexpr int STOPO = SPB + 3 * 512;           // [2][128]
    static constexpr int GACC = STOPO + 2 * 512;          // [3][128] d/dx
    static constexpr int GTOPO = GACC + 3 * 512;          // [2][128]
    static constexpr int PSUM = GTOPO + 2 * 512;          // [2][96] partial row-0 dot products (column halves)
    static constexpr int G0R = PSUM + 2 * 384;            // [96] d loss / d sdf per row (unscaled)
    static constexpr int CW2 = G0R + 384;                 // [64]  dW2[0, :] accumulator (scaled)
    static constexpr int MISC = CW2 + 256;                // 32 floats / ints of control state
    static constexpr int BAR = MISC + 128;                // full[3], empty[3], acc_ready, z_ready
    static constexpr int TMEMH = BAR + 8 * (2 * NSTAGE + 2);
    static constexpr int TOTAL = TMEMH + 16;
};
static_assert(Smem::BAR % 8 == 0, "alignment");
static_assert(Smem::TOTAL <= 113 * 1024, "two CTAs per SM");

// control words in MISC (floats unless noted)
enum { M_SCALE = 0, M_INV = 1, M_FLUSH = 2 /*int*/, M_FIRST = 3 /*int: wgrad MMAs of this sub-tile overwrite*/, M_OLDINV = 4, M_HAVE = 5 /*int*/,
       M_DIRTY = 6 /*int*/, M_LOSS = 7, M_GSUM = 8, M_MX = 16 /*[8]*/ };

#ifdef MB_FDR_NO_RED
#define red_add2(a, b, c) do { } while (0)
#endif
__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; i++) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

// ---- shared-corner stencil of one (sample, set, level) ------------------------------------------------------------------------------
struct Stencil {
    float pos[3][3];        // [axis][0: base, 1: +eps row, 2: -eps row] fractional position inside the row's cell
    uint32_t pg[3], p1[3];  // base cell, its upper planes min(pg + 1, res - 1)
    uint32_t far[3];        // far plane along the axis on the crossing side
    int cross[3];           // 0: both +-eps rows of the axis stay in the base cell, +1: the +eps row crosses, -1: the -eps row crosses
};

__device__ __forceinline__ void stencil_setup(Stencil& st, const GridCtx& g, uint32_t res, const float cb[3], const float pp[3], const float pm[3]) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float ub = __fdiv_rn(__fadd_rn(cb[d], g.bound), g.two_bound);
        const float up = __fdiv_rn(__fadd_rn(pp[d], g.bound), g.two_bound);
        const float um = __fdiv_rn(__fadd_rn(pm[d], g.bound), g.two_bound);
        float dv;
        uint32_t pgp, pgm;
        st.pos[d][0] = locate(ub, res, false, 0, st.pg[d], dv);
        st.pos[d][1] = locate(up, res, false, 0, pgp, dv);
        st.pos[d][2] = locate(um, res, false, 0, pgm, dv);
        st.p1[d] = min(st.pg[d] + 1, res - 1);
        st.cross[d] = (pgp != st.pg[d]) ? 1 : ((pgm != st.pg[d]) ? -1 : 0);
        st.far[d] = (pgp != st.pg[d]) ? min(pgp + 1, res - 1) : pgm;
    }
}
__device__ __forceinline__ uint32_t base_index(const Stencil& st, const LevelInfo& L, uint32_t c) {
    return corner_index(L, (c & 1) ? st.p1[0] : st.pg[0], (c & 2) ? st.p1[1] : st.pg[1], (c & 4) ? st.p1[2] : st.pg[2]);
}
// far-plane corner j of axis A: bit 0 of j <-> the lower of the two other axes, bit 1 <-> the higher
template <int A>
__device__ __forceinline__ uint32_t far_index(const Stencil& st, const LevelInfo& L, uint32_t j) {
    constexpr int O1 = (A == 0) ? 1 : 0, O2 = (A == 2) ? 1 : 2;
    uint32_t X[3];
    X[A] = st.far[A];
    X[O1] = (j & 1) ? st.p1[O1] : st.pg[O1];
    X[O2] = (j & 2) ? st.p1[O2] : st.pg[O2];
    return corner_index(L, X[0], X[1], X[2]);
}
template <int A>
__device__ __forceinline__ constexpr uint32_t compress(uint32_t k) {      // corner index without bit A
    return A == 0 ? (k >> 1) : (A == 1 ? ((k & 1) | ((k >> 2) << 1)) : (k & 3));
}
// the 8 corners of the row (axis A, SGN 0: +eps, 1: -eps) selected from the base cell / the far plane
template <int A, int SGN>
__device__ __forceinline__ void row_corners(const Stencil& st, const float2 (&cv)[8], const float2 (&ev)[4], float2 (&c)[8]) {
    const bool crossed = SGN == 0 ? st.cross[A] > 0 : st.cross[A] < 0;
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) {
        const bool bit = (k >> A) & 1;
        const uint32_t j = compress<A>(k);
        float2 alt;
        if (SGN == 0) alt = bit ? ev[j] : cv[k | (1u << A)];          // cell shifted by +1: its lower plane is the base cell's upper plane
        else alt = bit ? cv[k & ~(1u << A)] : ev[j];                 // cell shifted by -1: its upper plane is the base cell's lower plane
        c[k] = crossed ? alt : cv[k];
    }
}
// trilinear blend with the arithmetic of grid_eval (gridencoder.cu:171-195)
__device__ __forceinline__ float2 interp8(const float (&P)[3], const float2 (&c)[8]) {
    float2 f = make_float2(0.f, 0.f);
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) {
        float w = 1.0f;
#pragma unroll
        for (uint32_t d = 0; d < 3; d++) w = __fmul_rn(w, (k & (1u << d)) ? P[d] : __fsub_rn(1.0f, P[d]));
        f.x = __fmaf_rn(w, c[k].x, f.x);
        f.y = __fmaf_rn(w, c[k].y, f.y);
    }
    return f;
}

template <int A>
__device__ __forceinline__ void gather_axis(const Stencil& st, const LevelInfo& L, const float2* __restrict__ tab, const float2 (&cv)[8],
                                            float2& fplus, float2& fminus) {
    float2 ev[4];
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) ev[j] = st.cross[A] ? __ldg(tab + far_index<A>(st, L, j)) : make_float2(0.f, 0.f);
    float P[3] = {st.pos[0][0], st.pos[1][0], st.pos[2][0]};
    float2 c[8];
    P[A] = st.pos[A][1];
    row_corners<A, 0>(st, cv, ev, c);
    fplus = interp8(P, c);
    P[A] = st.pos[A][2];
    row_corners<A, 1>(st, cv, ev, c);
    fminus = interp8(P, c);
}

// features of the six rows of one (sample, set) at one level -> S0 columns 40 + 2 l, 41 + 2 l of rows r0 .. r0 + 5
__device__ __forceinline__ void fd_gather(uint8_t* S0, int r0, int l, const GridCtx& g, const float cb[3], const float pp[3], const float pm[3]) {
    float2 f[6];
#pragma unroll
    for (int q = 0; q < 6; q++) f[q] = make_float2(0.f, 0.f);
    if ((uint32_t)l < g.n_levels) {
        const LevelInfo L = g.lv[l];
        const float2* tab = reinterpret_cast<const float2*>(g.emb) + L.off;
        Stencil st;
        stencil_setup(st, g, L.res, cb, pp, pm);
        float2 cv[8];
#pragma unroll
        for (uint32_t c = 0; c < 8; c++) cv[c] = __ldg(tab + base_index(st, L, c));
        gather_axis<0>(st, L, tab, cv, f[0], f[1]);
        gather_axis<1>(st, L, tab, cv, f[2], f[3]);
        gather_axis<2>(st, L, tab, cv, f[4], f[5]);
    }
#pragma unroll
    for (int q = 0; q < 6; q++) store_pair(S0, r0 + q, 40 + 2 * l, f[q].x, f[q].y, S0_LO, PITCH);
}

// backward of one row: blend weights -> corner accumulators (base / far plane); d/d(point) of the row, masked by the clamp derivative, is
// added to the thread's running sum over the six rows
template <int A, int SGN>
__device__ __forceinline__ void scatter_row(const Stencil& st, const float2 (&cv)[8], const float2 (&ev)[4], float2 (&accb)[8], float2 (&acce)[4],
                                            float g0, float g1, float res_f, float two_bound, const bool (&mk)[3], float (&dxsum)[3]) {
    float P[3] = {st.pos[0][0], st.pos[1][0], st.pos[2][0]};
    P[A] = st.pos[A][1 + SGN];
    const bool crossed = SGN == 0 ? st.cross[A] > 0 : st.cross[A] < 0;
    float2 c[8];
    row_corners<A, SGN>(st, cv, ev, c);
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) {
        float w = 1.0f;
#pragma unroll
        for (uint32_t d = 0; d < 3; d++) w = __fmul_rn(w, (k & (1u << d)) ? P[d] : __fsub_rn(1.0f, P[d]));
        const float a0 = w * g0, a1 = w * g1;
        const bool bit = (k >> A) & 1;
        const uint32_t j = compress<A>(k);
        const bool far_side = SGN == 0 ? bit : !bit;                          // this corner lies on the far plane when the row crossed
        const uint32_t kalt = SGN == 0 ? (k | (1u << A)) : (k & ~(1u << A));  // ... else it is this corner of the base cell
        // static register indices only (k, kalt, j are compile-time after unrolling): three predicated accumulations
        if (!crossed) { accb[k].x += a0; accb[k].y += a1; }
        else if (far_side) { acce[j].x += a0; acce[j].y += a1; }
        else { accb[kalt].x += a0; accb[kalt].y += a1; }
    }
    // h_k = <corner_k, g>: d f / d pos_d . g = sum over the 4 corner pairs along d of (bilinear weight of the other two axes) * (h_hi - h_lo)
    float hk[8];
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) hk[k] = c[k].x * g0 + c[k].y * g1;
#pragma unroll
    for (uint32_t gd = 0; gd < 3; gd++) {
        float acc = 0.f;
#pragma unroll
        for (uint32_t i4 = 0; i4 < 4; i4++) {
            float w = res_f;
            uint32_t cl = 0;
#pragma unroll
            for (uint32_t nd = 0; nd < 2; nd++) {
                const uint32_t d = (nd >= gd) ? nd + 1 : nd;
                if (i4 & (1u << nd)) { w *= P[d]; cl |= (1u << d); }
                else w *= (1.0f - P[d]);
            }
            acc += w * (hk[cl | (1u << gd)] - hk[cl]);
        }
        if (mk[gd]) dxsum[gd] += acc / two_bound;
    }
}

template <int A>
__device__ __forceinline__ void scatter_axis(const Stencil& st, const LevelInfo& L, const float2* __restrict__ tab, float* __restrict__ gt,
                                             const float2 (&cv)[8], float2 (&accb)[8], const float* __restrict__ G, int r0, int l, float inv_scale,
                                             float two_bound, bool live, const bool (&mb_)[3], bool mplus, bool mminus, float (&dxsum)[3]) {
    float2 ev[4], acce[4];
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
        ev[j] = (live && st.cross[A]) ? __ldg(tab + far_index<A>(st, L, j)) : make_float2(0.f, 0.f);
        acce[j] = make_float2(0.f, 0.f);
    }
    const float res_f = (float)L.res;
    bool mk[3] = {mb_[0], mb_[1], mb_[2]};
    {
        const int r = r0 + 2 * A;
        const float g0 = live ? G[(2 * l) * RT + r] * inv_scale : 0.f, g1 = live ? G[(2 * l + 1) * RT + r] * inv_scale : 0.f;
        mk[A] = mplus;
        scatter_row<A, 0>(st, cv, ev, accb, acce, g0, g1, res_f, two_bound, mk, dxsum);
    }
    {
        const int r = r0 + 2 * A + 1;
        const float g0 = live ? G[(2 * l) * RT + r] * inv_scale : 0.f, g1 = live ? G[(2 * l + 1) * RT + r] * inv_scale : 0.f;
        mk[A] = mminus;
        scatter_row<A, 1>(st, cv, ev, accb, acce, g0, g1, res_f, two_bound, mk, dxsum);
    }
    if (live && st.cross[A]) {
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) red_add2(gt + 2 * far_index<A>(st, L, j), acce[j].x, acce[j].y);
    }
}

// the (clamped) query points of one (sample, set): base, +eps and -eps per axis; inb* = clamp derivative (1 where the clamp is inactive)
struct PairPoints { float cb[3], pp[3], pm[3]; bool inb[3], inp[3], inm[3]; };
__device__ __forceinline__ void pair_points(PairPoints& q, const float* __restrict__ base, int s, float bound) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float v = base[d * TM + s], vp = __fadd_rn(v, FD_EPS), vm = __fadd_rn(v, -FD_EPS);
        q.cb[d] = fminf(fmaxf(v, -bound), bound);
        q.pp[d] = fminf(fmaxf(vp, -bound), bound);
        q.pm[d] = fminf(fmaxf(vm, -bound), bound);
        q.inb[d] = v >= -bound && v <= bound;
        q.inp[d] = vp >= -bound && vp <= bound;
        q.inm[d] = vm >= -bound && vm <= bound;
    }
}

// hash-grid backward of one sub-tile.  Thread = (level, pair) with the 16 (sample, set) pairs of one level in the 16 lanes of a half-warp:
// neighbouring pairs (the two sets of a sample, consecutive samples of a ray) that share the base cell are merged by a segmented
// shuffle reduction before the red.v2 (the table scatter is bound by the LSU's atomic issue rate); d/d(point), summed over the six rows
// and the 16 levels, goes to gacc [3][TM] (shared-memory atomics).  G[32][RT] = feature gradients of the 96 rows (scaled).
__device__ __noinline__ void fd_scatter(const GridCtx g, const float* __restrict__ base, int s, const float* __restrict__ G, float* __restrict__ gemb,
                                        float* __restrict__ gacc, float inv_scale, int tid) {
    const int l = (tid >> 4) & 15, pr = tid & 15, lane = tid & 31, l16 = lane & 15;
    const int r0 = pr * 6;
    const bool live = (uint32_t)l < g.n_levels;
    const LevelInfo L = g.lv[live ? l : 0];
    const float2* tab = reinterpret_cast<const float2*>(g.emb) + L.off;
    float* gt = gemb + 2 * (size_t)L.off;
    PairPoints q;
    pair_points(q, base, s, g.bound);
    Stencil st;
    stencil_setup(st, g, L.res, q.cb, q.pp, q.pm);
    float2 cv[8], accb[8];
#pragma unroll
    for (uint32_t c = 0; c < 8; c++) {
        cv[c] = live ? __ldg(tab + base_index(st, L, c)) : make_float2(0.f, 0.f);
        accb[c] = make_float2(0.f, 0.f);
    }
    float dxsum[3] = {0.f, 0.f, 0.f};
    scatter_axis<0>(st, L, tab, gt, cv, accb, G, r0, l, inv_scale, g.two_bound, live, q.inb, q.inp[0], q.inm[0], dxsum);
    scatter_axis<1>(st, L, tab, gt, cv, accb, G, r0, l, inv_scale, g.two_bound, live, q.inb, q.inp[1], q.inm[1], dxsum);
    scatter_axis<2>(st, L, tab, gt, cv, accb, G, r0, l, inv_scale, g.two_bound, live, q.inb, q.inp[2], q.inm[2], dxsum);
    // ---- merge runs of equal base cells among the 16 pairs of this level (segmented reduction towards the run head) ----
    const uint32_t key = live ? (st.pg[0] | (st.pg[1] << 8) | (st.pg[2] << 16)) : (0x80000000u | (uint32_t)lane);
    const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1, 16);
    const bool head = l16 == 0 || key != prev;
    const uint32_t heads = (__ballot_sync(0xffffffffu, head) >> (lane & 16)) & 0xFFFFu;
    const uint32_t after = heads >> (l16 + 1);
    const int rem = after ? __ffs(after) : 16 - l16;          // lanes l16 .. l16 + rem - 1 belong to this lane's run
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
        const bool take = d < rem;
#pragma unroll
        for (uint32_t c = 0; c < 8; c++) {
            const float vx = __shfl_down_sync(0xffffffffu, accb[c].x, d, 16), vy = __shfl_down_sync(0xffffffffu, accb[c].y, d, 16);
            if (take) { accb[c].x += vx; accb[c].y += vy; }
        }
    }
    if (live && head) {
#pragma unroll
        for (uint32_t c = 0; c < 8; c++) {
            if (accb[c].x != 0.f || accb[c].y != 0.f) red_add2(gt + 2 * base_index(st, L, c), accb[c].x, accb[c].y);
        }
    }
    // ---- d/d(sample point): sum over the two sets (lane bit 0) and the two levels of the warp (lane bit 4), then shared-memory atomics ----
#pragma unroll
    for (int d = 0; d < 3; d++) {
        dxsum[d] += __shfl_xor_sync(0xffffffffu, dxsum[d], 1);
        dxsum[d] += __shfl_xor_sync(0xffffffffu, dxsum[d], 16);
    }
    if ((lane & 17) == 0) {
#pragma unroll
        for (int d = 0; d < 3; d++) atomicAdd(gacc + d * TM + s, dxsum[d]);
    }
}

struct Args {
    const float* x;          // [M,3]
    const float* topo;       // [M,2] or NULL (zeros)
    const float* noise;      // [M,3] or NULL (zeros)
    float noise_std;
    uint32_t M;
    float gmul;              // d(out) / d(sum |n - n_p|), e.g. 1 / (3 M)
    float* normal;           // [M,3] or NULL
    float* normal_raw;       // [M,3] or NULL
    float* loss;             // [1], accumulated: gmul * sum |n - n_p|
    float* g_x;              // [M,3] written
    float* g_topo;           // [M,2] written, or NULL
    float* g_emb;            // table gradient, accumulated
    float* g_arena;          // arena gradient, accumulated
};

__global__ void __launch_bounds__(NTHREADS, 2) fd_reg_tc_kernel(const mb_field_params p, const Args a, const uint8_t* __restrict__ tcw_f,
                                                               const uint32_t* __restrict__ off_f, const uint8_t* __restrict__ tcw_d,
                                                               const uint32_t* __restrict__ off_d, const int nsub) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* spa = reinterpret_cast<float*>(smem + Smem::SPA);
    float* spb = reinterpret_cast<float*>(smem + Smem::SPB);
    float* stopo = reinterpret_cast<float*>(smem + Smem::STOPO);
    float* gacc = reinterpret_cast<float*>(smem + Smem::GACC);
    float* gtopo = reinterpret_cast<float*>(smem + Smem::GTOPO);
    float* psum = reinterpret_cast<float*>(smem + Smem::PSUM);
    float* g0r = reinterpret_cast<float*>(smem + Smem::G0R);
    float* cw2 = reinterpret_cast<float*>(smem + Smem::CW2);
    float* misc = reinterpret_cast<float*>(smem + Smem::MISC);
    volatile int* misci = reinterpret_cast<volatile int*>(smem + Smem::MISC);
    float* G = reinterpret_cast<float*>(smem + Smem::DZ);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::BAR);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + Smem::TMEMH);
    uint64_t* full = bars;
    uint64_t* empty = bars + NSTAGE;
    uint64_t* acc_ready = bars + 2 * NSTAGE;
    uint64_t* z_ready = bars + 2 * NSTAGE + 1;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* AR = p.arena;
    float* GA = a.g_arena;

    __shared__ LevelInfo s_levels[16];
    if (p.offsets) init_levels(s_levels, p.offsets, p.S, p.H);
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; i++) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        mbar_init(acc_ready, 1);
        mbar_init(z_ready, NWORK / 32);
        mbar_fence_init();
    }
    if (warp == NWORK / 32) tmem_alloc<256>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;
    const uint32_t TMt = (uint32_t)SS * (uint32_t)nsub;       // samples per tile
    const uint32_t n_tiles = div_up(a.M, TMt);
    const uint32_t my_tiles = (blockIdx.x < n_tiles) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // weight slabs: forward table rows 12, 13 (sdf0: 5 x 4096 B, sdf1: 4 x 4096 B); dgrad table rows 1, 0 (sdf1: 4 x 4096 B, sdf0: 4 x 5120 B)
    const uint32_t f0_off = off_f[3 * 12], f1_off = off_f[3 * 13], d1_off = off_d[3 * 1], d0_off = off_d[3 * 0];

    if (warp == NWORK / 32 + 1) {
        // ================================ weight loader thread ================================
        if (lane == 0 && my_tiles > 0) {
            uint32_t loads = 0;
            auto load = [&](const uint8_t* src, uint32_t bytes) {
                const uint32_t stg = loads % NSTAGE;
                if (loads >= NSTAGE) mbar_wait(empty + stg, ((loads / NSTAGE) - 1) & 1);
                mbar_arrive_expect_tx(full + stg, bytes);
                bulk_g2s(smem + Smem::W + stg * STAGE_BYTES, src, bytes, full + stg);
                loads++;
            };
            for (uint64_t u = 0; u < (uint64_t)my_tiles * (uint64_t)nsub; u++) {
                for (uint32_t st = 0; st < 5; st++) load(tcw_f + f0_off + (size_t)st * 4096, 4096);
                for (uint32_t st = 0; st < 4; st++) load(tcw_f + f1_off + (size_t)st * 4096, 4096);
                for (uint32_t st = 0; st < 4; st++) load(tcw_d + d1_off + (size_t)st * 4096, 4096);
                for (uint32_t st = 0; st < 4; st++) load(tcw_d + d0_off + (size_t)st * 5120, 5120);
            }
        }
    } else if (warp == NWORK / 32) {
        // ================================ MMA-issue thread ================================
        if (lane == 0 && my_tiles > 0) {
            uint32_t uses = 0, z_count = 0;
            const uint32_t sm_base = smem_u32(smem);
            const uint32_t s0_base = sm_base + Smem::S0, x0_base = sm_base + Smem::X0, dz_base = sm_base + Smem::DZ, w_base = sm_base + Smem::W;
            auto wait_z = [&]() { mbar_wait(z_ready, z_count & 1); z_count++; tc_fence_after(); };
            auto gemm_ring = [&](uint32_t a_base, uint32_t a_lo, uint32_t nk, uint32_t rows, uint32_t n) {
                const uint32_t idesc = make_idesc_f16(n);
                const uint64_t a_hi0 = make_smem_desc(a_base, PITCH, 128), a_lo0 = make_smem_desc(a_base + a_lo, PITCH, 128);
                const uint64_t b_op = make_smem_desc(w_base, 16u * rows, 128);
                const uint64_t b_lo_add = (32u * rows) >> 4;
                for (uint32_t s = 0; s < nk; s++) {
                    const uint32_t stg = uses % NSTAGE;
                    mbar_wait(full + stg, (uses / NSTAGE) & 1);
                    tc_fence_after();
                    const uint64_t a_hi = a_hi0 + (uint64_t)s * (2 * PITCH >> 4), a_lod = a_lo0 + (uint64_t)s * (2 * PITCH >> 4);
                    const uint64_t b_hi = b_op + (uint64_t)stg * (STAGE_BYTES >> 4), b_lo = b_hi + b_lo_add;
                    umma_f16(tmem, a_hi, b_hi, idesc, s > 0 ? 1u : 0u);
                    umma_f16(tmem, a_hi, b_lo, idesc, 1u);
                    umma_f16(tmem, a_lod, b_hi, idesc, 1u);
                    umma_commit(empty + stg);
                    uses++;
                }
            };
            // acc[wcol][k][n] (+)= A^T dZ : both operands MN-major (features x rows), K = 96 rows in 6 steps
            auto wgrad = [&](uint32_t a_base, uint32_t a_lo, uint32_t wcol, bool first) {
                const uint32_t idesc = make_idesc_f16(64) | (1u << 15) | (1u << 16);
                const uint64_t a_hi0 = make_smem_desc(a_base, 128, PITCH), a_lo0 = make_smem_desc(a_base + a_lo, 128, PITCH);
                const uint64_t b_hi0 = make_smem_desc(dz_base, 128, PITCH), b_lo0 = make_smem_desc(dz_base + X_LO, 128, PITCH);
#pragma unroll
                for (uint32_t s = 0; s < RT / 16; s++) {
                    umma_f16(tmem + wcol, a_hi0 + s * 16, b_hi0 + s * 16, idesc, (first && s == 0) ? 0u : 1u);
                    umma_f16(tmem + wcol, a_hi0 + s * 16, b_lo0 + s * 16, idesc, 1u);
                    umma_f16(tmem + wcol, a_lo0 + s * 16, b_hi0 + s * 16, idesc, 1u);
                }
            };
            for (uint32_t it = 0; it < my_tiles; it++) {
                for (int j = 0; j < nsub; j++) {
                    wait_z(); gemm_ring(s0_base, S0_LO, 5, 64, 64); umma_commit(acc_ready);                                   // Z1 = S0 W0^T
                    wait_z(); gemm_ring(x0_base, X0_LO, 4, 64, 64); umma_commit(acc_ready);                                    // Z2 = A1 W1^T (stays in TMEM)
                    wait_z();
                    const bool first = misci[M_FIRST] != 0;
                    wgrad(x0_base, X0_LO, 192, first); gemm_ring(dz_base, X_LO, 4, 64, 64); umma_commit(acc_ready);            // dW1, dA1
                    wait_z(); wgrad(s0_base, S0_LO, 128, first); gemm_ring(dz_base, X_LO, 4, 80, 80); umma_commit(acc_ready);  // dW0, dS0
                }
            }
        }
    } else {
        // ================================ workers (8 warps) ================================
        const int q4 = warp & 3, h = warp >> 2;
        const int m = q4 * 32 + lane;                    // operand-tile row of this thread in the epilogues (valid if q4 < 3)
        const bool erow = q4 < 3;
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        uint32_t acc_count = 0;
        const GridCtx gs{p.emb_sdf, p.offsets, p.S, p.H, p.n_levels, p.bound, p.two_bound, s_levels};
        uint8_t* S0 = smem + Smem::S0;
        uint8_t* X0 = smem + Smem::X0;
        uint8_t* DZ = smem + Smem::DZ;
        auto bar_workers = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory"); };
        auto signal_z = [&]() { fence_proxy_async(); tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(z_ready); };
        auto wait_acc = [&]() { mbar_wait(acc_ready, acc_count & 1); acc_count++; tc_fence_after(); };
        // flush of the TMEM weight-gradient accumulators (+ dW2 row 0) into the gradient arena, descaled
        auto flush = [&](float inv_scale) {
            const int krow = q4 * 32 + lane;
            if (q4 < 3) {
                {   // sdf layer 0: 80 rows (kind 2 order) x 64 columns at TMEM column 128
                    int korig = (krow < 80) ? tc_korig(2, krow) : -1;
                    if (korig >= (int)p.sdf[0].K) korig = -1;
                    float v[32];
                    tmem_ld32(tmem + lane_base + 128 + h * 32, v);
                    float* dst = nullptr;
                    if (korig >= 0) dst = GA + p.sdf[0].wt_off + (size_t)korig * p.sdf[0].N_pad + h * 32;
                    else if (krow == 39) dst = GA + p.sdf[0].b_off + h * 32;       // constant-1 feature: column sums of dZ0 = bias gradient
                    if (dst) {
#pragma unroll
                        for (int c = 0; c < 8; c++) red_add4(dst + 4 * c, v[4 * c] * inv_scale, v[4 * c + 1] * inv_scale, v[4 * c + 2] * inv_scale, v[4 * c + 3] * inv_scale);
                    }
                }
                {   // sdf layer 1: 64 rows (+ row 64 = bias gradient) x 64 columns at TMEM column 192
                    float v[32];
                    tmem_ld32(tmem + lane_base + 192 + h * 32, v);
                    float* dst = (krow < 64) ? GA + p.sdf[1].wt_off + (size_t)krow * p.sdf[1].N_pad + h * 32 : GA + p.sdf[1].b_off + h * 32;
                    if (krow <= 64) {
#pragma unroll
                        for (int c = 0; c < 8; c++) red_add4(dst + 4 * c, v[4 * c] * inv_scale, v[4 * c + 1] * inv_scale, v[4 * c + 2] * inv_scale, v[4 * c + 3] * inv_scale);
                    }
                }
            }
            if (tid >= 128 && tid < 192) {
                const float sv = cw2[tid - 128];       // dW2[0][k]: Wt slot [k][n = 0]
                if (sv != 0.f) red_add(GA + p.sdf[2].wt_off + (size_t)(tid - 128) * p.sdf[2].N_pad, sv * inv_scale);
                cw2[tid - 128] = 0.f;
            }
            tc_fence_before();
        };

        if (tid < RT) {      // core 8 of the A1 tile: column 64 = 1 (never overwritten: the A1 epilogue writes cores 0..7 only)
            const float one8[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            store_core(X0, tid, 8, one8, X0_LO, PITCH);
        }
        const float b2 = __ldg(AR + p.sdf[2].b_off);
        unsigned long long ph[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        long long tprev = clock64();
#define FDR_PHASE(i) do { if (PHASE_TIMING && tid == 0) { const long long tn = clock64(); ph[i] += (unsigned long long)(tn - tprev); tprev = tn; } } while (0)
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint32_t m0 = tile * TMt;
            const int nv = (int)min(TMt, a.M - m0);
            // ---- per-sample inputs ----
            for (int idx = tid; idx < 3 * TM; idx += NWORK) {
                const int mm = idx / 3, ax = idx - mm * 3;
                float v = 0.f, w = 0.f;
                if (mm < nv) {
                    v = a.x[(size_t)m0 * 3 + idx];
                    w = a.noise ? __fadd_rn(v, __fmul_rn(a.noise[(size_t)m0 * 3 + idx], a.noise_std)) : v;      // xyzs + randn * std (morpheus.py:724)
                }
                spa[ax * TM + mm] = v;
                spb[ax * TM + mm] = w;
                gacc[ax * TM + mm] = 0.f;
            }
            for (int idx = tid; idx < 2 * TM; idx += NWORK) {
                const int mm = idx / 2, ax = idx - mm * 2;
                stopo[ax * TM + mm] = (mm < nv && a.topo) ? a.topo[(size_t)m0 * 2 + idx] : 0.f;
                gtopo[ax * TM + mm] = 0.f;
            }
            if (tid < 64) cw2[tid] = 0.f;
            if (tid == 0) {
                misc[M_SCALE] = 1.f; misc[M_INV] = 1.f; misci[M_HAVE] = 0; misci[M_DIRTY] = 0; misc[M_LOSS] = 0.f; misc[M_GSUM] = 0.f;
            }
            bar_workers();
            FDR_PHASE(0);

#pragma unroll 1
            for (int j = 0; j < nsub; j++) {
                // ---- rows of the sub-tile: r = 6 pr + q, pair pr = 2 sl + set, sample s = 8 j + sl.  No staging of the row points: every
                //      consumer derives them from the sample points (a few FADD / FMNMX), which saves a barrier per sub-tile ----
                FDR_PHASE(1);
                // ---- S0: hash-grid features (thread = level x pair, shared corners; the 16 pairs of a level sit in one half-warp, so
                //      pairs in the same cell coalesce into the same sectors) ----
                {
                    const int l = tid >> 4, pr = tid & 15;
                    const int s = SS * j + (pr >> 1);
                    PairPoints q;
                    pair_points(q, (pr & 1) ? spb : spa, s, p.bound);
                    fd_gather(S0, pr * 6, l, gs, q.cb, q.pp, q.pm);
                }
                FDR_PHASE(2);
                // ---- S0: frequency features (3 axes x 96 rows) + pads.  (Sharing the sin / cos of the coordinates the six rows of a pair have in
                //      common was tried: the fp16 (hi, lo) conversions and 2-byte stores dominate this phase, not the sincosf -- slower.) ----
                for (int it = tid; it < 4 * RT; it += NWORK) {
                    const int r = it % RT, ax = it / RT;
                    const int pr = r / 6, qq = r - pr * 6;
                    const int s = SS * j + (pr >> 1);
                    if (ax < 3) {
                        float v = ((pr & 1) ? spb : spa)[ax * TM + s];
                        if (ax == (qq >> 1)) v = __fadd_rn(v, (qq & 1) ? -FD_EPS : FD_EPS);
                        freq_axis_tc(S0, r, ax, fminf(fmaxf(v, -p.bound), p.bound), (int)p.n_freq, S0_LO, PITCH);
                    } else {
                        store_one(S0, r, 39, 1.0f, S0_LO, PITCH);      // constant-1 pad feature (its W0 column is zero): db0 from the wgrad MMA
                        const float v[8] = {(pr & 1) ? 0.f : stopo[s], (pr & 1) ? 0.f : stopo[TM + s], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        store_core(S0, r, 9, v, S0_LO, PITCH);
                    }
                }
                FDR_PHASE(3);
                // ---- A1 = relu(S0 W0^T + b0) -> X0 ----
                signal_z(); wait_acc();
                FDR_PHASE(4);
                if (erow) {
                    float v[32];
                    tmem_ld32(tmem + lane_base + h * 32, v);
                    const float* b0 = AR + p.sdf[0].b_off + h * 32;
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        float o[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) o[i] = fmaxf(v[c * 8 + i] + __ldg(b0 + c * 8 + i), 0.f);
                        store_core(X0, m, h * 4 + c, o, X0_LO, PITCH);
                    }
                }
                FDR_PHASE(5);
                // ---- Z2 = A1 W1^T (+ b1): pass 1 = sdf of the row (dot product with W2[0, :]); the accumulator stays in TMEM ----
                signal_z(); wait_acc();
                FDR_PHASE(6);
                if (erow) {
                    float v[32];
                    tmem_ld32(tmem + lane_base + h * 32, v);
                    const float* b1 = AR + p.sdf[1].b_off + h * 32;
                    const float* w2 = AR + p.sdf[2].w_off + h * 32;          // W[n = 0][k]
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; i++) acc = __fmaf_rn(fmaxf(v[i] + __ldg(b1 + i), 0.f), __ldg(w2 + i), acc);
                    psum[h * RT + m] = acc;
                }
                bar_workers();
                FDR_PHASE(7);
                // ---- normals of both sets, loss, d loss / d sdf of the 12 rows of a sample (8 threads), scale control (thread 0) ----
                if (warp == 0) {
                    float mx = 0.f;
                    if (lane < SS) {
                        const int s = SS * j + lane;
                        const bool valid = s < nv;
                        float n[2][3], raw[2][3], inv[2];
                        bool clampd[2];
#pragma unroll
                        for (int set = 0; set < 2; set++) {
                            const int r0 = lane * 12 + set * 6;
                            float sq[6];
#pragma unroll
                            for (int qq = 0; qq < 6; qq++) sq[qq] = __fadd_rn(__fadd_rn(psum[r0 + qq], psum[RT + r0 + qq]), b2);
#pragma unroll
                            for (int ax = 0; ax < 3; ax++) raw[set][ax] = __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(sq[2 * ax], sq[2 * ax + 1])), FD_EPS);
                            const float d2 = raw[set][0] * raw[set][0] + raw[set][1] * raw[set][1] + raw[set][2] * raw[set][2];
                            clampd[set] = !(d2 > 1e-20f);
                            inv[set] = 1.0f / sqrtf(fmaxf(d2, 1e-20f));
#pragma unroll
                            for (int ax = 0; ax < 3; ax++) {
                                float v = raw[set][ax] * inv[set];
                                if (isnan(v)) v = 0.f;
                                else if (isinf(v)) v = v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
                                n[set][ax] = v;
                            }
                        }
                        float lsum = 0.f, gn[3];
#pragma unroll
                        for (int ax = 0; ax < 3; ax++) {
                            const float d = n[0][ax] - n[1][ax];
                            lsum += fabsf(d);
                            gn[ax] = (d > 0.f) ? a.gmul : ((d < 0.f) ? -a.gmul : 0.f);          // d |d| / d d = sign(d), sign(0) = 0 (torch)
                        }
#pragma unroll
                        for (int set = 0; set < 2; set++) {
                            const float sg = set ? -1.f : 1.f;
                            const float g3[3] = {sg * gn[0], sg * gn[1], sg * gn[2]};
                            float graw[3];
                            if (clampd[set]) { graw[0] = g3[0] * inv[set]; graw[1] = g3[1] * inv[set]; graw[2] = g3[2] * inv[set]; }
                            else {
                                const float dot = n[set][0] * g3[0] + n[set][1] * g3[1] + n[set][2] * g3[2];
#pragma unroll
                                for (int ax = 0; ax < 3; ax++) graw[ax] = inv[set] * (g3[ax] - n[set][ax] * dot);
                            }
                            const float hh = 0.5f / FD_EPS;
                            const int r0 = lane * 12 + set * 6;
#pragma unroll
                            for (int ax = 0; ax < 3; ax++) {
                                const float gp_ = valid ? graw[ax] * hh : 0.f;
                                g0r[r0 + 2 * ax] = gp_;
                                g0r[r0 + 2 * ax + 1] = -gp_;
                                mx = fmaxf(mx, fabsf(gp_));
                            }
                        }
                        // (the +-eps pairs cancel exactly: the bias of the sdf output gets no gradient from an FD normal)
                        if (valid) {
                            atomicAdd(misc + M_LOSS, lsum);
                            const size_t gm = (size_t)(m0 + s);
                            if (a.normal) { a.normal[gm * 3] = n[0][0]; a.normal[gm * 3 + 1] = n[0][1]; a.normal[gm * 3 + 2] = n[0][2]; }
                            if (a.normal_raw) { a.normal_raw[gm * 3] = raw[0][0]; a.normal_raw[gm * 3 + 1] = raw[0][1]; a.normal_raw[gm * 3 + 2] = raw[0][2]; }
                        }
                    }
#pragma unroll
                    for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    if (lane == 0) {
                        int flushf = 0;
                        const bool dirty = misci[M_DIRTY] != 0;
                        int firstf = (j == 0) ? 1 : 0;
                        if (mx > 0.f && isfinite(mx)) {
                            const float cur = misc[M_SCALE];
                            const bool have = misci[M_HAVE] != 0;
                            if (!have || !dirty || mx * cur >= 2048.f) {
                                if (have && dirty) { flushf = 1; firstf = 1; misc[M_OLDINV] = misc[M_INV]; }
                                int e = 0;
                                frexpf(mx, &e);
                                e = max(-100, min(100, 9 - e));          // mx * 2^e in [256, 512)
                                misc[M_SCALE] = ldexpf(1.0f, e);
                                misc[M_INV] = ldexpf(1.0f, -e);
                                misci[M_HAVE] = 1;
                            }
                            misci[M_DIRTY] = 1;
                        }
                        misci[M_FLUSH] = flushf;
                        misci[M_FIRST] = firstf;
                    }
                }
                bar_workers();
                if (misci[M_FLUSH]) {       // (uniform across the workers) rescale: retire what the accumulators hold at the old scale
                    flush(misc[M_OLDINV]);
                    bar_workers();
                }
                const float scale = misc[M_SCALE], inv_scale = misc[M_INV];
                FDR_PHASE(8);
                // ---- pass 2: dZ1 = g0 W2[0,:] (A2 > 0) -> DZ ; dW2[0,:] += g0 A2 ----
                {
                    float z[32], ga[32];
                    if (erow) {
                        float v[32];
                        tmem_ld32(tmem + lane_base + h * 32, v);
                        const float g0s = g0r[m] * scale;
                        const float* b1 = AR + p.sdf[1].b_off + h * 32;
                        const float* w2 = AR + p.sdf[2].w_off + h * 32;
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            const float a2 = fmaxf(v[i] + __ldg(b1 + i), 0.f);
                            z[i] = (a2 > 0.f) ? g0s * __ldg(w2 + i) : 0.f;
                            ga[i] = g0s * a2;
                        }
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            float o[8];
#pragma unroll
                            for (int i = 0; i < 8; i++) o[i] = z[c * 8 + i];
                            store_core(DZ, m, h * 4 + c, o, X_LO, PITCH);
                        }
                        const float cg = warp_colsum32(ga, lane);
                        atomicAdd(cw2 + h * 32 + lane, cg);
                    }
                }
                FDR_PHASE(9);
                // ---- dZ0 = (dZ1 W1) (A1 > 0) -> DZ ----
                signal_z(); wait_acc();
                FDR_PHASE(10);
                if (erow) {
                    float v[32];
                    tmem_ld32(tmem + lane_base + h * 32, v);
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const int kc = h * 4 + c;
                        const uint4 aa = *reinterpret_cast<const uint4*>(X0 + kc * PITCH + (m >> 3) * 128 + (m & 7) * 16);
                        const uint32_t aw[4] = {aa.x, aa.y, aa.z, aa.w};
                        float o[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const uint16_t hb = (uint16_t)(aw[i >> 1] >> ((i & 1) * 16));
                            const bool pos = (hb & 0x7FFF) != 0 && !(hb & 0x8000);
                            o[i] = pos ? v[c * 8 + i] : 0.f;
                        }
                        store_core(DZ, m, kc, o, X_LO, PITCH);
                    }
                }
                FDR_PHASE(11);
                // ---- d(S0) (80 columns): h = 0: frequency columns 0..38 -> sin/cos backward ; h = 1: grid columns 40..71 -> G, topo 72..73 ----
                signal_z(); wait_acc();
                FDR_PHASE(12);
                if (erow) {
                    const int pr = m / 6, qq = m - pr * 6;
                    const int s = SS * j + (pr >> 1);
                    if (h == 0) {
                        // frequency path: d/d(point) = g_x + sum_k 2^k (g_sin cos - g_cos sin).  The sin / cos VALUES are the forward features,
                        // still in the S0 operand tile as fp16 (hi, lo) pairs (exact to 2^-22): no second sincosf
                        float v[32], w[8];
                        tmem_ld32(tmem + lane_base, v);
                        tmem_ld8(tmem + lane_base + 32, w);
                        float acc3[3] = {v[0], v[1], v[2]};
                        const uint8_t* srow = S0 + (m >> 3) * 128 + (m & 7) * 16;
#pragma unroll
                        for (int c = 0; c < 5; c++) {
                            const uint4 hi = *reinterpret_cast<const uint4*>(srow + c * PITCH);
                            const uint4 lo = *reinterpret_cast<const uint4*>(srow + c * PITCH + S0_LO);
                            const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                const int fi = c * 8 + i;
                                if (fi >= 3 && fi < 39) {
                                    const int kk = (fi - 3) / 6, t = (fi - 3) - kk * 6;       // t < 3: sin of axis t ; else cos of axis t - 3
                                    const __half hv = __ushort_as_half((unsigned short)(hw[i >> 1] >> ((i & 1) * 16)));
                                    const __half lv = __ushort_as_half((unsigned short)(lw[i >> 1] >> ((i & 1) * 16)));
                                    const float val = __half2float(hv) + __half2float(lv);
                                    const float fr = (float)(1 << kk);
                                    if (kk < (int)p.n_freq) {
                                        if (t < 3) { const int ic = fi + 3; acc3[t] -= fr * (ic < 32 ? v[ic] : w[ic - 32]) * val; }
                                        else { const int is = fi - 3; acc3[t - 3] += fr * (is < 32 ? v[is] : w[is - 32]) * val; }
                                    }
                                }
                            }
                        }
                        const float* base = (pr & 1) ? spb : spa;
#pragma unroll
                        for (int ax = 0; ax < 3; ax++) {
                            float pv = base[ax * TM + s];
                            if (ax == (qq >> 1)) pv = __fadd_rn(pv, (qq & 1) ? -FD_EPS : FD_EPS);
                            if (pv >= -p.bound && pv <= p.bound) atomicAdd(gacc + ax * TM + s, acc3[ax] * inv_scale);      // clamp derivative
                        }
                    } else {
                        // G aliases DZ: the MMAs that read DZ completed before acc_ready fired
                        float v[32], t4[4];
                        tmem_ld32(tmem + lane_base + 40, v);
                        tmem_ld4(tmem + lane_base + 72, t4);
#pragma unroll
                        for (int i = 0; i < 32; i++) G[i * RT + m] = v[i];
                        if (a.topo && (pr & 1) == 0) {
                            atomicAdd(gtopo + s, t4[0] * inv_scale);
                            atomicAdd(gtopo + TM + s, t4[1] * inv_scale);
                        }
                    }
                }
                tc_fence_before();
                bar_workers();
                FDR_PHASE(13);
                {
                    const int pr = tid & 15;
                    fd_scatter(gs, (pr & 1) ? spb : spa, SS * j + (pr >> 1), G, a.g_emb, gacc, inv_scale, tid);
                }
                bar_workers();          // G (aliasing DZ) and S0 are rewritten by the next sub-tile
                FDR_PHASE(14);
            }

            FDR_PHASE(15);
            // ---- end of tile: flush the accumulators, loss, per-sample gradients ----
            bar_workers();
            flush(misc[M_INV]);
            if (tid == 0 && misc[M_LOSS] != 0.f) red_add(a.loss, misc[M_LOSS] * a.gmul);
            bar_workers();
            for (int idx = tid; idx < 3 * TM; idx += NWORK) {
                const int mm = idx / 3, ax = idx - mm * 3;
                if (mm < nv) a.g_x[(size_t)m0 * 3 + idx] = gacc[ax * TM + mm];
            }
            if (a.g_topo) {
                for (int idx = tid; idx < 2 * TM; idx += NWORK) {
                    const int mm = idx / 2, ax = idx - mm * 2;
                    if (mm < nv) a.g_topo[(size_t)m0 * 2 + idx] = gtopo[ax * TM + mm];
                }
            }
            bar_workers();
        }
        if (PHASE_TIMING && tid == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) atomicAdd(&g_fdr_phase[i], ph[i]);
        }
#undef FDR_PHASE
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NWORK / 32) tmem_dealloc<256>(tmem);
}

}  // namespace tcr
}  // namespace mb

extern "C" int mb_fd_regulariser_tc(const mb_field_params* p, const float* x, const float* topo, const float* noise, float noise_std, uint32_t M,
                                    float gmul, float* normal, float* normal_raw, float* loss, float* g_x, float* g_topo, float* g_emb_sdf,
                                    float* g_arena, const void* tc_weights, const uint32_t* tc_off, const void* tc_weights_t, const uint32_t* tc_off_t,
                                    mb_stream_t stream) {
    using namespace mb;
    if (!p || !tc_weights || !tc_off || !tc_weights_t || !tc_off_t) { set_error("fd_regulariser_tc: null argument"); return MB_EINVAL; }
    if (M == 0) return MB_OK;
    if (!x || !p->arena || !p->emb_sdf || !p->offsets || !loss || !g_x || !g_emb_sdf || !g_arena) {
        set_error("fd_regulariser_tc: x / arena / emb_sdf / offsets / loss / g_x / g_emb_sdf / g_arena is null");
        return MB_EINVAL;
    }
    if (topo && !g_topo) { set_error("fd_regulariser_tc: topo needs g_topo"); return MB_EINVAL; }
    // shared-corner stencil: a +-eps point may cross at most one cell face per axis and never both sides of the base cell
    if (!(mb::FD_EPS * 128.0f / p->two_bound < 0.5f) || p->H != 16) { set_error("fd_regulariser_tc: eps * max_resolution / (2 bound) must stay below 0.5 cells"); return MB_EINVAL; }
    constexpr size_t smem = (size_t)tcr::Smem::TOTAL;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tcr::fd_reg_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("fd_regulariser_tc: cannot reserve %zu B smem: %s", smem, cudaGetErrorString(e)); return MB_ECUDA; }
        attr_set = true;
    }
    const uint32_t slots = (uint32_t)mb_sm_count() * 2u;
    const uint32_t n128 = div_up(M, tcr::TM);
    const int nsub = n128 >= 4 * slots ? 16 : (n128 >= 2 * slots ? 8 : 4);       // samples per tile: 128 / 64 / 32
    const uint32_t n_tiles = div_up(M, (uint32_t)tcr::SS * (uint32_t)nsub);
    const uint32_t grid = min(n_tiles, slots);
    tcr::Args a{x, topo, noise, noise_std, M, gmul, normal, normal_raw, loss, g_x, g_topo, g_emb_sdf, g_arena};
    tcr::fd_reg_tc_kernel<<<grid, tcr::NTHREADS, smem, (cudaStream_t)stream>>>(*p, a, (const uint8_t*)tc_weights, tc_off, (const uint8_t*)tc_weights_t,
                                                                              tc_off_t, nsub);
    return check_launch("fd_regulariser_tc");
}

/* debug: cumulative clock64 cycles per phase of mb_fd_regulariser_tc (worker thread 0 of every CTA; needs -DMB_FDR_PHASE_TIMING=1) */
extern "C" int mb_debug_fdr_phases(unsigned long long* host_out16, int reset) {
    using namespace mb;
    unsigned long long z[16] = {0};
    if (host_out16 && cudaMemcpyFromSymbol(host_out16, tcr::g_fdr_phase, sizeof(z)) != cudaSuccess) { set_error("debug_fdr_phases: copy failed"); return MB_ECUDA; }
    if (reset && cudaMemcpyToSymbol(tcr::g_fdr_phase, z, sizeof(z)) != cudaSuccess) { set_error("debug_fdr_phases: reset failed"); return MB_ECUDA; }
    return MB_OK;
}
