// Tensor-core (tcgen05 + TMEM) backward of everything that sits on the hash grids: the SDF network (73->64->64->33),
// the colour network (64->64->64->3), Laplace density, shading and the finite-difference normal queries
// (models/model.py:273-307,367-398,483-533 under autograd).  Replaces csrc/field_bwd.cu when the tensor-core engine is on;
// the deformation / topology networks are handled by csrc/field_bwd_tc.cu, which consumes g_def_out / g_topo_out.
//
// Per tile of 128 rows (samples, or (sample, +-eps query) pairs for the FD normals) the forward is recomputed on the
// tensor cores (activations kept in shared memory as fp16 hi/lo operand tiles), then for every layer
//     wgrad  dW[k][n] += A[m][k] dZ[m][n]   (MN-major operands, accumulator = a dedicated TMEM column range that lives for
//                                            the whole tile: flushed once per tile with red.global.add.v4)
//     dgrad  dA[m][k]  = dZ[m][n] W[n][k]   (K-major operands, weights streamed through a bulk-copy ring)
// and the input gradient d(S0) is pushed through the frequency encoding (d/dx) and the hash grids (table scatter with
// red.v2 + d/dx by corner differencing, as kernel_grid_backward / kernel_input_backward of gridencoder.cu:253-378).
// dZ is scaled per tile by a power of two so that the fp16 (hi, lo) split keeps ~22 bits; 3 MMAs per product.
// TMEM map (512 columns): [0,128) work accumulator, [128,192) dW sdf0, [192,256) sdf1, [256,320) sdf2, [320,384) col0,
// [384,448) col1, [448,464) col2.
#include "field_common.cuh"
#include "tc_common.cuh"
#include "tc_field.cuh"

namespace mb {
namespace tcs {

using namespace mb::tc;

constexpr int TM = 128;
constexpr int NWORK = 512;              // 16 worker warps
constexpr int NTHREADS = NWORK + 64;    // + MMA-issue warp + weight-loader warp
constexpr int NSTAGE = 3;
constexpr int STAGE_BYTES = 5120;
constexpr int S0_LO = 20480;        // lo offset of the 80-column S0 tile
constexpr int X_LO = 16384;         // lo offset of a 64-column tile
constexpr int MAX_OPS = 64;

struct OpS {
    uint8_t kind;     // 0 forward GEMM, 1 backward (wgrad + dgrad)
    uint8_t a_tile;   // forward: A operand tile; backward: stored activation (wgrad A operand).  0 = S0, 1..3 = X0..X2
    uint8_t nk;       // forward: K/16;  backward: dgrad K steps (used dZ columns / 16)
    uint8_t n;        // forward: MMA N; backward: dgrad MMA N = rows of the weight slab (64 or 80)
    uint8_t wlayer;   // weight-table row (forward table: 0..5 = sdf0..2, col0..2; dgrad table: same order)
    uint8_t wn;       // backward: wgrad N (dZ columns used: 64, 48 or 16)
    uint16_t wcol;    // backward: TMEM column of the wgrad accumulator
    uint32_t src_off; // byte offset of the layer's K=16 weight slabs in the forward (kind 0) / dgrad (kind 1) table
    uint32_t rows;    // rows of one slab (N_pad forward, R dgrad): slab bytes = 64 * rows
};

struct Smem {
    static constexpr int S0 = 0;                         // 40960
    static constexpr int X = 40960;                      // 3 x 32768
    static constexpr int DZ = X + 3 * 32768;             // 32768
    static constexpr int G = DZ + 32768;                 // fp32 [32][128]
    static constexpr int W = G + 16384;                  // NSTAGE x 5120
    static constexpr int F = W + NSTAGE * STAGE_BYTES;   // fp32 rows of 128
    static constexpr int SX = F;                         // [3]
    static constexpr int SXW = SX + 3 * 512;             // [3]
    static constexpr int SPT = SXW + 3 * 512;            // [3]
    static constexpr int STOPO = SPT + 3 * 512;          // [2]
    static constexpr int STQ = STOPO + 2 * 512;          // [2]
    static constexpr int SSDF = STQ + 2 * 512;           // [1]
    static constexpr int SALB = SSDF + 512;              // [3]
    static constexpr int GXW = SALB + 3 * 512;           // [3]
    static constexpr int GX = GXW + 3 * 512;             // [3]
    static constexpr int GPT = GX + 3 * 512;             // [3]
    static constexpr int GTOPO = GPT + 3 * 512;          // [2]
    static constexpr int GSQ = GTOPO + 2 * 512;          // [6]
    static constexpr int GALB = GSQ + 6 * 512;           // [3]
    static constexpr int GSD = GALB + 3 * 512;           // [1]  d/d(sdf) of the main query
    static constexpr int CS = GSD + 512;                 // [1]  column sums (sdf L2 feature columns, colour path)
    static constexpr int CSB = CS + 512;                 // [2]  per-tile bias-gradient accumulators: sdf1 | sdf0 | col1 | col0 (64 each)
    static constexpr int MISC = CSB + 2 * 512;           // 16 floats
    static constexpr int OPS = MISC + 64;                // OpS[MAX_OPS]
    static constexpr int BAR = OPS + MAX_OPS * 16;       // full[3], empty[3], acc_ready, z_ready
    static constexpr int TMEMH = BAR + 8 * (2 * NSTAGE + 2);
    static constexpr int TOTAL = TMEMH + 16;
};
static_assert(Smem::BAR % 8 == 0 && Smem::OPS % 8 == 0, "alignment");
static_assert(Smem::TOTAL <= 232448, "shared memory budget");
static_assert(sizeof(OpS) == 16, "OpS layout");

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
// column sums of a [32 lanes][16] register tile: afterwards lane L holds the sum of column (L & 15) over the 16 lanes that
// share its bit 4 (the two half-warps hold partial sums of the same column)
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
    for (int half = 8; half >= 1; half >>= 1) {
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
__device__ __forceinline__ uint32_t tile_base(int t) { return t == 0 ? Smem::S0 : Smem::X + (t - 1) * 32768; }
__device__ __forceinline__ uint32_t tile_lo(int t) { return t == 0 ? S0_LO : X_LO; }

// grid backward from G[32][128] (feature gradients of the row-wise points pt3, row r <-> query index qbase + r):
// work item = (level, run of <= 6 consecutive rows aligned to multiples of 6 in query space, i.e. the +-eps queries of
// ONE sample in an FD sub-tile, 6 neighbouring samples of a ray in the main sub-tile).  Rows of a run that fall into
// the same grid cell are merged in registers, so the table scatter costs one red.v2 per corner per CELL instead of
// per row (the scatter is LSU-throughput bound).  d/d(point) by corner differencing as kernel_input_backward.
__device__ __noinline__ void grid_bwd_runs(const GridCtx g, const float* __restrict__ pt3, const float* __restrict__ G, float* __restrict__ gemb,
                                            float* __restrict__ gp, float inv_scale, int qbase, int tid) {
    // grid backward from G[32][128] (feature gradients of the row-wise points pt3, row r <-> query index qbase + r):
    // work item = (run of <= 6 consecutive rows aligned to multiples of 6 in query space, level).  A run is the +-eps
    // queries of ONE sample in an FD sub-tile (6 neighbouring samples of a ray in the main sub-tile).  Rows of a run that
    // fall into the same grid cell are merged in registers, so the table scatter costs one red.v2 per corner per CELL
    // instead of per row (the scatter is LSU-throughput bound).  The 16 levels of a run sit in the 16 lanes of a half-warp:
    // d/d(point) (corner differencing, as kernel_input_backward) is summed over levels with shuffles and added to gp by
    // one lane without atomics (a row belongs to exactly one run).
    const int lane = tid & 31;
    const int run0 = qbase / 6;
    const int nruns = (qbase + TM - 1) / 6 - run0 + 1;
    const int nl = (int)min(g.n_levels, 16u);
    const int items = nruns * 16;
    for (int it0 = (tid & ~31); it0 < items; it0 += NWORK) {       // warp-uniform trip count
        const int it = it0 + lane;
        const int l = it & 15, rr = it >> 4;
        const bool live = rr < nruns && l < nl;
        const LevelInfo L = g.lv[live ? l : 0];
        const uint32_t res = L.res;
        const float scale = (float)res;
        const float2* tab = reinterpret_cast<const float2*>(g.emb) + L.off;
        float* gt = gemb + 2 * (size_t)L.off;
        bool have = false;
        uint32_t c0 = 0, c1 = 0, c2 = 0;
        uint32_t cidx[8];
        float2 cv[8];
        float acc[16];
#pragma unroll 1
        for (int i = 0; i < 6; i++) {
            const int r = (run0 + rr) * 6 + i - qbase;
            const bool rowok = live && r >= 0 && r < TM;
            float dx[3] = {0.f, 0.f, 0.f};
            if (rowok) {
                const float g0 = G[(2 * l) * TM + r] * inv_scale, g1 = G[(2 * l + 1) * TM + r] * inv_scale;
                float u[3];
#pragma unroll
                for (int d = 0; d < 3; d++) u[d] = __fdiv_rn(__fadd_rn(pt3[d * TM + r], g.bound), g.two_bound);
                const bool inb = !(u[0] < 0 || u[0] > 1 || u[1] < 0 || u[1] > 1 || u[2] < 0 || u[2] > 1);
                if (inb && !(g0 == 0.f && g1 == 0.f)) {
                    float pos[3], dv;
                    uint32_t pg[3];
#pragma unroll
                    for (int d = 0; d < 3; d++) pos[d] = locate(u[d], res, false, 0, pg[d], dv);
                    if (!have || pg[0] != c0 || pg[1] != c1 || pg[2] != c2) {
                        if (have) {
#pragma unroll
                            for (int c = 0; c < 8; c++) red_add2(gt + 2 * cidx[c], acc[2 * c], acc[2 * c + 1]);
                        }
                        have = true;
                        c0 = pg[0]; c1 = pg[1]; c2 = pg[2];
                        const uint32_t p1[3] = {min(pg[0] + 1, res - 1), min(pg[1] + 1, res - 1), min(pg[2] + 1, res - 1)};
#pragma unroll
                        for (uint32_t c = 0; c < 8; c++) {
                            cidx[c] = corner_index(L, (c & 1) ? p1[0] : pg[0], (c & 2) ? p1[1] : pg[1], (c & 4) ? p1[2] : pg[2]);
                            cv[c] = __ldg(tab + cidx[c]);
                            acc[2 * c] = acc[2 * c + 1] = 0.f;
                        }
                    }
#pragma unroll
                    for (uint32_t c = 0; c < 8; c++) {
                        float w = 1.0f;
#pragma unroll
                        for (uint32_t d = 0; d < 3; d++) w = __fmul_rn(w, (c & (1u << d)) ? pos[d] : __fsub_rn(1.0f, pos[d]));
                        acc[2 * c] = __fmaf_rn(w, g0, acc[2 * c]);
                        acc[2 * c + 1] = __fmaf_rn(w, g1, acc[2 * c + 1]);
                    }
#pragma unroll
                    for (uint32_t gd = 0; gd < 3; gd++) {
                        float a = 0.f;
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
                            const float2 lo = cv[cl], hi = cv[cl | (1u << gd)];
                            a += w * ((hi.x - lo.x) * g0 + (hi.y - lo.y) * g1);
                        }
                        dx[gd] = a / g.two_bound;
                    }
                }
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
                for (int d = 0; d < 3; d++) dx[d] += __shfl_xor_sync(0xffffffffu, dx[d], o);
            }
            if (l == 0 && rowok) {
#pragma unroll
                for (int d = 0; d < 3; d++) gp[d * TM + r] += dx[d];
            }
        }
        if (have) {
#pragma unroll
            for (int c = 0; c < 8; c++) red_add2(gt + 2 * cidx[c], acc[2 * c], acc[2 * c + 1]);
        }
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) field_bwd_sdf_tc_kernel(const mb_field_params p, const mb_field_io io, const mb_field_grads gr,
                                                                      const uint8_t* __restrict__ tcw_f, const uint32_t* __restrict__ off_f,
                                                                      const uint8_t* __restrict__ tcw_d, const uint32_t* __restrict__ off_d) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* sx = reinterpret_cast<float*>(smem + Smem::SX);
    float* sxw = reinterpret_cast<float*>(smem + Smem::SXW);
    float* spt = reinterpret_cast<float*>(smem + Smem::SPT);
    float* stopo = reinterpret_cast<float*>(smem + Smem::STOPO);
    float* stq = reinterpret_cast<float*>(smem + Smem::STQ);
    float* ssdf = reinterpret_cast<float*>(smem + Smem::SSDF);
    float* salb = reinterpret_cast<float*>(smem + Smem::SALB);
    float* gxw = reinterpret_cast<float*>(smem + Smem::GXW);
    float* gx = reinterpret_cast<float*>(smem + Smem::GX);
    float* gpt = reinterpret_cast<float*>(smem + Smem::GPT);
    float* gtopo = reinterpret_cast<float*>(smem + Smem::GTOPO);
    float* gsq = reinterpret_cast<float*>(smem + Smem::GSQ);
    float* galb = reinterpret_cast<float*>(smem + Smem::GALB);
    float* gsd = reinterpret_cast<float*>(smem + Smem::GSD);
    float* cs = reinterpret_cast<float*>(smem + Smem::CS);
    float* csb = reinterpret_cast<float*>(smem + Smem::CSB);
    float* misc = reinterpret_cast<float*>(smem + Smem::MISC);
    float* G = reinterpret_cast<float*>(smem + Smem::G);
    OpS* ops = reinterpret_cast<OpS*>(smem + Smem::OPS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::BAR);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + Smem::TMEMH);
    uint64_t* full = bars;
    uint64_t* empty = bars + NSTAGE;
    uint64_t* acc_ready = bars + 2 * NSTAGE;
    uint64_t* z_ready = bars + 2 * NSTAGE + 1;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t flags = io.flags;
    const float* AR = p.arena;
    float* GA = gr.g_arena;
    const bool topo_live = (flags & (MB_F_WARP | MB_F_TOPO_IN)) != 0;
    const bool do_main = (flags & MB_F_MAIN) != 0;
    const bool do_color = do_main && (flags & MB_F_COLOR);
    const bool color_grad = do_color && gr.g_color && io.shading != MB_SHADE_TEXTURELESS && io.shading != MB_SHADE_NORMAL;
    const bool fd_grads = (flags & MB_F_FD) && (gr.g_normal || gr.g_normal_raw || (io.shading != MB_SHADE_ALBEDO && gr.g_color));
    const bool delegate = (flags & MB_F_FD_DELEGATE) && gr.g_fd;       // FD chains run in mb_field_backward_fd_tc
    const bool need_fd = fd_grads && !delegate;

    __shared__ int n_ops_s;
    __shared__ LevelInfo s_levels[16];
    if (p.offsets) init_levels(s_levels, p.offsets, p.S, p.H);
    if (tid == 0) {
        int n = 0;
        auto fwd = [&](int a_tile, int nk, int nn, int wl) { ops[n] = OpS{0, (uint8_t)a_tile, (uint8_t)nk, (uint8_t)nn, (uint8_t)wl, 0, 0, off_f[3 * (12 + wl)], off_f[3 * (12 + wl) + 2]}; n++; };
        auto bwd = [&](int act_tile, int nk, int rows, int wl, int wn, int wcol) { ops[n] = OpS{1, (uint8_t)act_tile, (uint8_t)nk, (uint8_t)rows, (uint8_t)wl, (uint8_t)wn, (uint16_t)wcol, off_d[3 * wl], (uint32_t)rows}; n++; };
        if (do_main) {
            fwd(0, 5, 64, 0); fwd(1, 4, 64, 1); fwd(2, 4, 48, 2);             // S0 -> X0 -> X1 -> h
            if (do_color) { fwd(1, 4, 64, 3); fwd(2, 4, 64, 4); fwd(3, 4, 16, 5); }   // C0 = X0 -> X1 -> X2 -> rgb
            if (color_grad) { bwd(3, 1, 64, 5, 16, 448); bwd(2, 4, 64, 4, 64, 384); bwd(1, 4, 64, 3, 64, 320); }
            if (do_color) { fwd(0, 5, 64, 0); fwd(1, 4, 64, 1); }             // recompute A1, A2 (the colour tiles reused X0, X1)
            bwd(2, 3, 64, 2, 48, 256); bwd(1, 4, 64, 1, 64, 192); bwd(0, 4, 80, 0, 64, 128);
        }
        if (need_fd)
            for (int j = 0; j < 6; j++) {
                fwd(0, 5, 64, 0); fwd(1, 4, 64, 1);
                bwd(1, 4, 64, 1, 64, 192); ops[n - 1].kind = 2;      // + layer-2 weight gradient from the 16-column dZ2 side tile
                bwd(0, 4, 80, 0, 64, 128);
            }
        n_ops_s = n;
        for (int i = 0; i < NSTAGE; i++) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        mbar_init(acc_ready, 1);
        mbar_init(z_ready, NWORK / 32);        // one elected arrival per worker warp
        mbar_fence_init();
    }
    if (warp == NWORK / 32) tmem_alloc<512>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;
    const int n_ops = n_ops_s;
    const uint32_t n_tiles = div_up(io.M, TM);
    const uint32_t my_tiles = (blockIdx.x < n_tiles) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == NWORK / 32 + 1) {
        // ================================ weight loader thread ================================
        if (lane == 0 && my_tiles > 0 && n_ops > 0) {
            const uint64_t total_ops = (uint64_t)my_tiles * n_ops;
            uint32_t loads = 0;
            for (uint64_t u = 0; u < total_ops; u++) {
                const OpS o = ops[u % n_ops];
                const uint8_t* src = (o.kind == 0 ? tcw_f : tcw_d) + o.src_off;
                const uint32_t bytes = 64u * o.rows;
                for (uint32_t st = 0; st < o.nk; st++) {
                    const uint32_t stg = loads % NSTAGE;
                    if (loads >= NSTAGE) mbar_wait(empty + stg, ((loads / NSTAGE) - 1) & 1);
                    mbar_arrive_expect_tx(full + stg, bytes);
                    bulk_g2s(smem + Smem::W + stg * STAGE_BYTES, src + (size_t)st * bytes, bytes, full + stg);
                    loads++;
                }
            }
        }
    } else if (warp == NWORK / 32) {
        // ================================ MMA-issue thread ================================
        if (lane == 0 && my_tiles > 0 && n_ops > 0) {
            const uint64_t total_ops = (uint64_t)my_tiles * n_ops;
            uint32_t uses = 0, z_count = 0;
            const uint32_t sm_base = smem_u32(smem);
            const uint32_t dz_base = sm_base + Smem::DZ;
            const uint64_t zw_hi0 = make_smem_desc(dz_base, 128, 2048), zw_lo0 = make_smem_desc(dz_base + X_LO, 128, 2048);   // dZ as MN-major wgrad B
            uint32_t used_mask = 0;          // wgrad accumulators already written in this tile (bit = (wcol-128)/64)
            for (uint64_t u = 0; u < total_ops; u++) {
                const uint32_t oi = (uint32_t)(u % n_ops);
                if (oi == 0) used_mask = 0;
                const OpS o = ops[oi];
                mbar_wait(z_ready, z_count & 1);
                z_count++;
                tc_fence_after();
                if (o.kind == 2) {
                    // ---- FD chain: layer-2 weight gradient acc[256][k][0:16] (+)= A2^T dZ2 (dZ2 = 16-column side tile in G) ----
                    const uint32_t a_base = sm_base + tile_base(2), g_base = sm_base + Smem::G;
                    const uint32_t idesc = make_idesc_f16(16) | (1u << 15) | (1u << 16);
                    const bool first = !(used_mask & 4u);
                    used_mask |= 4u;
                    const uint64_t a_hi0 = make_smem_desc(a_base, 128, 2048), a_lo0 = make_smem_desc(a_base + X_LO, 128, 2048);
                    const uint64_t b_hi0 = make_smem_desc(g_base, 128, 2048), b_lo0 = make_smem_desc(g_base + 4096, 128, 2048);
#pragma unroll
                    for (uint32_t s = 0; s < 8; s++) {
                        umma_f16(tmem + 256, a_hi0 + s * 16, b_hi0 + s * 16, idesc, (first && s == 0) ? 0u : 1u);
                        umma_f16(tmem + 256, a_hi0 + s * 16, b_lo0 + s * 16, idesc, 1u);
                        umma_f16(tmem + 256, a_lo0 + s * 16, b_hi0 + s * 16, idesc, 1u);
                    }
                }
                if (o.kind >= 1) {
                    // ---- wgrad: acc[wcol][k][n] (+)= A^T dZ ; both operands MN-major, K = 128 rows in 8 steps ----
                    const uint32_t a_base = sm_base + tile_base(o.a_tile), a_lo = tile_lo(o.a_tile);
                    const uint32_t idesc = make_idesc_f16(o.wn) | (1u << 15) | (1u << 16);
                    const uint32_t bit = 1u << ((o.wcol - 128) / 64);
                    const bool first = !(used_mask & bit);
                    used_mask |= bit;
                    const uint64_t a_hi0 = make_smem_desc(a_base, 128, 2048), a_lo0 = make_smem_desc(a_base + a_lo, 128, 2048);
#pragma unroll
                    for (uint32_t s = 0; s < 8; s++) {
                        const uint64_t a_hi = a_hi0 + s * 16, a_lod = a_lo0 + s * 16;        // + 256 B per K step (MN-major)
                        const uint64_t b_hi = zw_hi0 + s * 16, b_lo = zw_lo0 + s * 16;
                        umma_f16(tmem + o.wcol, a_hi, b_hi, idesc, (first && s == 0) ? 0u : 1u);
                        umma_f16(tmem + o.wcol, a_hi, b_lo, idesc, 1u);
                        umma_f16(tmem + o.wcol, a_lod, b_hi, idesc, 1u);
                    }
                }
                {
                    // ---- forward GEMM (A = activation tile) or dgrad (A = dZ tile); B = weight slabs from the ring ----
                    const uint32_t a_base = (o.kind == 0) ? sm_base + tile_base(o.a_tile) : dz_base;
                    const uint32_t a_lo = (o.kind == 0) ? tile_lo(o.a_tile) : (uint32_t)X_LO;
                    const uint32_t rows = o.rows;                                            // rows of the slab (N_pad or R)
                    const uint32_t idesc = make_idesc_f16(o.n);
                    const uint64_t a_hi0 = make_smem_desc(a_base, 2048, 128), a_lo0 = make_smem_desc(a_base + a_lo, 2048, 128);
                    const uint64_t b_op = make_smem_desc(sm_base + Smem::W, 16u * rows, 128);
                    const uint64_t b_lo_add = (32u * rows) >> 4;
                    for (uint32_t s = 0; s < o.nk; s++) {
                        const uint32_t stg = uses % NSTAGE;
                        mbar_wait(full + stg, (uses / NSTAGE) & 1);
                        tc_fence_after();
                        const uint64_t a_hi = a_hi0 + (uint64_t)s * 256, a_lod = a_lo0 + (uint64_t)s * 256;    // + 4096 B per K step
                        const uint64_t b_hi = b_op + (uint64_t)stg * (STAGE_BYTES >> 4), b_lo = b_hi + b_lo_add;
                        umma_f16(tmem, a_hi, b_hi, idesc, s > 0 ? 1u : 0u);
                        umma_f16(tmem, a_hi, b_lo, idesc, 1u);
                        umma_f16(tmem, a_lod, b_hi, idesc, 1u);
                        umma_commit(empty + stg);
                        uses++;
                    }
                }
                umma_commit(acc_ready);
            }
        }
    } else {
        // ================================ workers ================================
        // 512 threads: row m = tid & 127 of the 128-row tile, part = tid >> 7 in 0..3 (column / level quarter of every phase)
        const int m = tid & (TM - 1);
        const int part = tid >> 7;
        const int warp_q = warp & 3;
        const uint32_t lane_base = (uint32_t)(warp_q * 32) << 16;
        uint32_t acc_count = 0;
        const GridCtx gs{p.emb_sdf, p.offsets, p.S, p.H, p.n_levels, p.bound, p.two_bound, s_levels};
        const GridCtx gc{p.emb_col, p.offsets, p.S, p.H, p.n_levels, p.bound, p.two_bound, s_levels};
        uint8_t* S0 = smem + Smem::S0;
        uint8_t* X0 = smem + Smem::X;
        uint8_t* X1 = X0 + 32768;
        uint8_t* X2 = X0 + 65536;
        uint8_t* DZ = smem + Smem::DZ;
        auto bar_workers = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory"); };
        auto signal_z = [&]() { fence_proxy_async(); tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(z_ready); };
        auto wait_acc = [&]() { mbar_wait(acc_ready, acc_count & 1); acc_count++; tc_fence_after(); };

        // S0 operand (80 columns) at the row-wise points pt3 with row-wise topo: every part gathers 4 grid levels (one core),
        // parts 0..2 build the frequency features of one axis, part 3 the pad / topo columns
        auto build_s0 = [&](const float* pt3, const float* topo2) {
            const float pnt[3] = {pt3[m], pt3[TM + m], pt3[2 * TM + m]};
            gather_levels_tc(S0, m, 40, gs, 4 * part, 4, pnt, S0_LO);
            if (part < 3) {
                freq_axis_tc(S0, m, part, pt3[part * TM + m], (int)p.n_freq, S0_LO);
            } else {
                store_one(S0, m, 39, 0.f, S0_LO);
                const float v[8] = {topo2[m], topo2[TM + m], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                store_core(S0, m, 9, v, S0_LO);
            }
        };
        // hidden forward epilogue: relu(acc + bias) -> 64-column tile dst (16 columns per part)
        auto ep_hidden = [&](const float* bias, uint8_t* dst) {
            float v[16];
            const int col0 = part * 16;
            tmem_ld16(tmem + lane_base + col0, v);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; i++) o[i] = fmaxf(v[j * 8 + i] + __ldg(bias + col0 + j * 8 + i), 0.f);
                store_core(dst, m, col0 / 8 + j, o, X_LO);
            }
        };
        // dgrad epilogue: dZ_prev = acc * (act > 0) -> DZ tile (64 columns); bias gradient of the previous layer
        auto ep_mask = [&](const uint8_t* act, float* cacc) {
            float v[16];
            const int col0 = part * 16;
            tmem_ld16(tmem + lane_base + col0, v);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int kc = col0 / 8 + j;
                const uint4 a = *reinterpret_cast<const uint4*>(act + kc * 2048 + (m >> 3) * 128 + (m & 7) * 16);
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint16_t hb = (uint16_t)(aw[i >> 1] >> ((i & 1) * 16));
                    const bool pos = (hb & 0x7FFF) != 0 && !(hb & 0x8000);
                    o[i] = pos ? v[j * 8 + i] : 0.f;
                    v[j * 8 + i] = o[i];
                }
                store_core(DZ, m, kc, o, X_LO);
            }
            const float csum = warp_colsum16(v, lane);
            atomicAdd(cacc + col0 + (lane & 15), csum);      // bias gradient: flushed once per tile
        };
        // FD query, layer-1 forward epilogue fused with the layer-2 backward: A2 = relu(acc + b1) -> X1 (wgrad operand) and,
        // because the query's only output is sdf = row 0 of layer 2, dZ1[m][k] = g0[m] * W2[0][k] * (A2 > 0) directly (rank-1:
        // no dgrad MMA, no extra round trip); dZ2 = {g0, 0...} goes to a 16-column side tile for the layer-2 weight gradient
        auto ep_hidden_fd = [&](float g0s) {
            float v[16];
            const int col0 = part * 16;
            tmem_ld16(tmem + lane_base + col0, v);
            const float* b1 = AR + p.sdf[1].b_off + col0;
            const float* w2 = AR + p.sdf[2].w_off + col0;          // W[n = 0][k]
#pragma unroll
            for (int j = 0; j < 2; j++) {
                float o[8], z[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    o[i] = fmaxf(v[j * 8 + i] + __ldg(b1 + j * 8 + i), 0.f);
                    z[i] = (o[i] > 0.f) ? g0s * __ldg(w2 + j * 8 + i) : 0.f;
                    v[j * 8 + i] = z[i];
                }
                store_core(X1, m, col0 / 8 + j, o, X_LO);
                store_core(DZ, m, col0 / 8 + j, z, X_LO);
            }
            const float csum = warp_colsum16(v, lane);
            atomicAdd(csb + col0 + (lane & 15), csum);
        };
        // d(S0) epilogue (80 columns in the work accumulator): every part moves 8 grid columns to G; parts 0..2 push two
        // frequency bands each (columns 3+12*part .. 14+12*part) through sin/cos -> gp; part 0 adds the raw-point columns;
        // part 3 returns the topo columns
        auto ep_ds0 = [&](const float* pt3, float* gp, float inv_scale, float& gt0, float& gt1) {
            gt0 = gt1 = 0.f;
            {
                float w[8];
                tmem_ld8(tmem + lane_base + 40 + 8 * part, w);
#pragma unroll
                for (int i = 0; i < 8; i++) G[(8 * part + i) * TM + m] = w[i];
            }
            if (part < 3) {
                float v[16];
                tmem_ld16(tmem + lane_base + 12 * part, v);      // columns 12*part .. 12*part+15
                float acc[3] = {0.f, 0.f, 0.f};
                if (part == 0) { acc[0] = v[0]; acc[1] = v[1]; acc[2] = v[2]; }
                float f = (part == 0) ? 1.0f : (part == 1 ? 4.0f : 16.0f);
#pragma unroll
                for (int kk = 0; kk < 2; kk++) {
                    if (2 * part + kk < (int)p.n_freq) {
#pragma unroll
                        for (int a = 0; a < 3; a++) {
                            float sn, cn;
                            sincosf(pt3[a * TM + m] * f, &sn, &cn);
                            acc[a] += f * (v[3 + 6 * kk + a] * cn - v[6 + 6 * kk + a] * sn);
                        }
                    }
                    f *= 2.0f;
                }
#pragma unroll
                for (int a = 0; a < 3; a++) atomicAdd(gp + a * TM + m, acc[a] * inv_scale);
            } else {
                float t4[4];
                tmem_ld4(tmem + lane_base + 72, t4);             // columns 72, 73: topo
                gt0 = t4[0] * inv_scale;
                gt1 = t4[1] * inv_scale;
            }
        };
        // flush one wgrad accumulator: rows = input features (tc order), 16 columns per part
        auto flush_acc = [&](int wcol, const mb_layer_desc& L, int kind, int krows, int ncols, float inv_scale) {
            const int krow = warp_q * 32 + lane;
            int korig = (krow < krows) ? tc_korig(kind, krow) : -1;
            if (korig >= (int)L.K) korig = -1;
            const int col0 = part * 16;
            if (col0 >= ncols) return;                 // warp-uniform
            float v[16];
            tmem_ld16(tmem + lane_base + wcol + col0, v);
            if (korig < 0) return;
            float* dst = GA + L.wt_off + (size_t)korig * L.N_pad + col0;
            if (ncols - col0 >= 16) {
#pragma unroll
                for (int j = 0; j < 4; j++) red_add4(dst + 4 * j, v[4 * j] * inv_scale, v[4 * j + 1] * inv_scale, v[4 * j + 2] * inv_scale, v[4 * j + 3] * inv_scale);
            } else {
                for (int j = 0; j < ncols - col0; j++) red_add(dst + j, v[j] * inv_scale);
            }
        };
        const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint32_t m0 = tile * TM;
            const int nv = (int)min((uint32_t)TM, io.M - m0);
            // ---- load inputs / saved values, clear accumulators ----
            for (int idx = tid; idx < 3 * TM; idx += NWORK) {
                const int mm = idx / 3, a = idx - mm * 3;
                const bool ok = mm < nv;
                const float xv = ok ? io.x[(size_t)m0 * 3 + idx] : 0.f;
                const float dv = (ok && (flags & MB_F_WARP)) ? gr.deform[(size_t)m0 * 3 + idx] : 0.f;
                sx[a * TM + mm] = xv;
                sxw[a * TM + mm] = xv + dv;
                gxw[a * TM + mm] = 0.f;
                gx[a * TM + mm] = 0.f;
                galb[a * TM + mm] = 0.f;
            }
            for (int idx = tid; idx < 2 * TM; idx += NWORK) {
                const int mm = idx / 2, a = idx - mm * 2;
                const bool ok = mm < nv;
                float v = 0.f;
                if (ok && (flags & MB_F_WARP)) v = gr.topo[(size_t)m0 * 2 + idx];
                else if (ok && (flags & MB_F_TOPO_IN)) v = io.topo_in[(size_t)m0 * 2 + idx];
                stopo[a * TM + mm] = v;
                gtopo[a * TM + mm] = (ok && gr.g_topo) ? gr.g_topo[(size_t)m0 * 2 + idx] : 0.f;
            }
            for (int idx = tid; idx < 6 * TM; idx += NWORK) gsq[idx] = 0.f;
            if (tid < TM) { gsd[tid] = 0.f; cs[tid] = 0.f; }
            if (tid < 256) csb[tid] = 0.f;
            bar_workers();

            // ---- forward of the main query (needed for albedo / sdf / the colour-net input) ----
            if (do_main) {
                build_s0(sxw, stopo);
                signal_z(); wait_acc(); ep_hidden(AR + p.sdf[0].b_off, X0);
                signal_z(); wait_acc(); ep_hidden(AR + p.sdf[1].b_off, X1);
                signal_z(); wait_acc();
                {
                    // h = acc[:, 0:33] + b : col 0 = sdf, cols 1..32 = geometric feature -> colour-net cores 4..7 (one per part)
                    float v[16];
                    tmem_ld16(tmem + lane_base + 8 * part, v);
                    const float* b2 = AR + p.sdf[2].b_off;
                    if (part == 0) ssdf[m] = v[0] + __ldg(b2);
                    if (do_color) {
                        float o[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) o[i] = v[1 + i] + __ldg(b2 + 8 * part + 1 + i);
                        store_core(X0, m, 4 + part, o, X_LO);
                    }
                }
                if (do_color) {
                    const float pnt[3] = {sxw[m], sxw[TM + m], sxw[2 * TM + m]};
                    gather_levels_tc(X0, m, 0, gc, 4 * part, 4, pnt, X_LO);
                    signal_z(); wait_acc(); ep_hidden(AR + p.color[0].b_off, X1);
                    signal_z(); wait_acc(); ep_hidden(AR + p.color[1].b_off, X2);
                    signal_z(); wait_acc();
                    if (part == 0) {
                        float v[4];
                        tmem_ld4(tmem + lane_base, v);
#pragma unroll
                        for (int a = 0; a < 3; a++) salb[a * TM + m] = 1.0f / (1.0f + expf(-(v[a] + __ldg(AR + p.color[2].b_off + a))));
                    }
                }
                tc_fence_before();
                bar_workers();
            }

            // ---- upstream -> local gradients (mirrors csrc/field_bwd.cu) ----
            float gbeta_local = 0.f;
            if (tid < nv) {
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
                        const float a0 = do_color ? salb[m] : 0.f, a1 = do_color ? salb[TM + m] : 0.f, a2 = do_color ? salb[2 * TM + m] : 0.f;
                        glam = gc3[0] * a0 + gc3[1] * a1 + gc3[2] * a2;
                        ga[0] = gc3[0] * lam; ga[1] = gc3[1] * lam; ga[2] = gc3[2] * lam;
                    } else if (io.shading == MB_SHADE_TEXTURELESS) {
                        glam = gc3[0] + gc3[1] + gc3[2];
                        ga[0] = ga[1] = ga[2] = 0.f;
                    } else {
                        gn[0] += 0.5f * gc3[0]; gn[1] += 0.5f * gc3[1]; gn[2] += 0.5f * gc3[2];
                        ga[0] = ga[1] = ga[2] = 0.f;
                    }
                    const float k = (ndl > 0.f) ? glam * (1.0f - io.ratio) : 0.f;
                    gn[0] += k * l[0]; gn[1] += k * l[1]; gn[2] += k * l[2];
                }
                if (color_grad) {
#pragma unroll
                    for (int a = 0; a < 3; a++) { const float s = salb[a * TM + m]; galb[a * TM + m] = ga[a] * s * (1.0f - s); }
                }
                if (fd_grads) {
                    float graw[3];
                    if (clamped) { graw[0] = gn[0] * inv; graw[1] = gn[1] * inv; graw[2] = gn[2] * inv; }
                    else {
                        const float dot = n[0] * gn[0] + n[1] * gn[1] + n[2] * gn[2];
                        graw[0] = inv * (gn[0] - n[0] * dot); graw[1] = inv * (gn[1] - n[1] * dot); graw[2] = inv * (gn[2] - n[2] * dot);
                    }
                    if (gr.g_normal_raw) { graw[0] += gr.g_normal_raw[(size_t)gm * 3]; graw[1] += gr.g_normal_raw[(size_t)gm * 3 + 1]; graw[2] += gr.g_normal_raw[(size_t)gm * 3 + 2]; }
                    const float h = 0.5f / FD_EPS;
#pragma unroll
                    for (int a = 0; a < 3; a++) { gsq[(2 * a) * TM + m] = graw[a] * h; gsq[(2 * a + 1) * TM + m] = -graw[a] * h; }
                }
                if (delegate) {
#pragma unroll
                    for (int q = 0; q < 6; q++) gr.g_fd[(size_t)gm * 6 + q] = gsq[q * TM + m];
                }
                if (do_main) {
                    float v = gr.g_sdf ? gr.g_sdf[gm] : 0.f;
                    if (gr.g_sigma) {
                        const float s = ssdf[m], b = __ldg(p.beta);
                        const float a = fabsf(s), e = expf(-a / b);
                        const float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
                        const float gsig = gr.g_sigma[gm];
                        v += gsig * (-0.5f * sg * sg * e / (b * b));
                        const float sigma = (1.0f / b) * (0.5f + 0.5f * sg * expm1f(-a / b));
                        gbeta_local += gsig * (-sigma / b + 0.5f * sg * e * a / (b * b * b));
                    }
                    gsd[m] = v;
                }
            }
            // per-tile scale: max |start-of-chain gradient|  (rows live in threads tid < 128 = warps 0..3)
            {
                if (tid < TM) {
                    float mx = fabsf(gsd[tid]);
#pragma unroll
                    for (int a = 0; a < 3; a++) mx = fmaxf(mx, fabsf(galb[a * TM + tid]));
#pragma unroll
                    for (int q = 0; q < 6; q++) mx = fmaxf(mx, fabsf(gsq[q * TM + tid]));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                        gbeta_local += __shfl_xor_sync(0xffffffffu, gbeta_local, o);
                    }
                    if (lane == 0) { misc[warp] = mx; misc[8 + warp] = gbeta_local; }
                }
                bar_workers();
                if (tid == 0) {
                    float mm = 0.f, gb = 0.f;
                    for (int w4 = 0; w4 < 4; w4++) { mm = fmaxf(mm, misc[w4]); gb += misc[8 + w4]; }
                    int e = 0;
                    if (mm > 0.f && isfinite(mm)) { frexpf(mm, &e); e = 10 - e; }
                    e = max(-100, min(100, e));
                    if (gb != 0.f && gr.g_beta) atomicAdd(gr.g_beta, gb);
                    misc[4] = ldexpf(1.0f, e);
                    misc[5] = ldexpf(1.0f, -e);
                }
                bar_workers();
            }
            const float scale = misc[4], inv_scale = misc[5];
            bar_workers();

            // ---- main query backward ----
            if (do_main) {
                if (color_grad) {
                    // dZ of the colour output layer: 16 columns
                    if (part == 0) {
                        const float v[8] = {galb[m] * scale, galb[TM + m] * scale, galb[2 * TM + m] * scale, 0.f, 0.f, 0.f, 0.f, 0.f};
                        store_core(DZ, m, 0, v, X_LO);
                        for (int a = 0; a < 3; a++) {
                            float s = galb[a * TM + m];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                            if (lane == 0 && s != 0.f) red_add(GA + p.color[2].b_off + a, s);
                        }
                    } else if (part == 1) {
                        store_core(DZ, m, 1, z8, X_LO);
                    }
                    signal_z(); wait_acc(); ep_mask(X2, csb + 128);      // -> dZ colour L1
                    signal_z(); wait_acc(); ep_mask(X1, csb + 192);      // -> dZ colour L0
                    signal_z(); wait_acc();
                    // d(C0): columns 0..31 colour-grid features -> G ; columns 32..63 = d(feat) = dZ2 columns 1..32 of the SDF net
                    if (part < 2) {
                        float v[16];
                        tmem_ld16(tmem + lane_base + 16 * part, v);
#pragma unroll
                        for (int i = 0; i < 16; i++) G[(16 * part + i) * TM + m] = v[i];
                    } else {
                        float v[32];
                        tmem_ld32(tmem + lane_base + 32, v);          // v[i] = dZ2 column 1 + i
                        if (part == 2) {
                            const float g0 = gsd[m] * scale;
                            const float o0[8] = {g0, v[0], v[1], v[2], v[3], v[4], v[5], v[6]};
                            store_core(DZ, m, 0, o0, X_LO);
                            float o[8];
#pragma unroll
                            for (int i = 0; i < 8; i++) o[i] = v[7 + i];
                            store_core(DZ, m, 1, o, X_LO);
#pragma unroll
                            for (int i = 0; i < 8; i++) o[i] = v[15 + i];
                            store_core(DZ, m, 2, o, X_LO);
                            float h16[16];
#pragma unroll
                            for (int i = 0; i < 16; i++) h16[i] = v[i];
                            const float csum = warp_colsum16(h16, lane);
                            atomicAdd(cs + (lane & 15), csum);
                        } else {
                            float o[8];
#pragma unroll
                            for (int i = 0; i < 8; i++) o[i] = v[23 + i];
                            store_core(DZ, m, 3, o, X_LO);
                            const float o4[8] = {v[31], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                            store_core(DZ, m, 4, o4, X_LO);
                            store_core(DZ, m, 5, z8, X_LO);
                            float h16[16];
#pragma unroll
                            for (int i = 0; i < 16; i++) h16[i] = v[16 + i];
                            const float csum = warp_colsum16(h16, lane);
                            atomicAdd(cs + 16 + (lane & 15), csum);
                        }
                    }
                    tc_fence_before();
                    bar_workers();
                    if (tid < 32) { const float s = cs[tid]; if (s != 0.f) red_add(GA + p.sdf[2].b_off + 1 + tid, s * inv_scale); cs[tid] = 0.f; }
                    grid_bwd_runs(gc, sxw, G, gr.g_emb_col, gxw, inv_scale, 0, tid);
                    bar_workers();
                } else {
                    if (part == 0) {
                        const float v[8] = {gsd[m] * scale, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        store_core(DZ, m, 0, v, X_LO);
                        store_core(DZ, m, 4, z8, X_LO);
                    } else if (part == 1) {
                        store_core(DZ, m, 1, z8, X_LO);
                        store_core(DZ, m, 5, z8, X_LO);
                    } else {
                        store_core(DZ, m, part, z8, X_LO);
                    }
                }
                if (tid < TM) {      // bias gradient of sdf L2, column 0
                    float s = gsd[tid];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if (lane == 0 && s != 0.f) red_add(GA + p.sdf[2].b_off, s);
                }
                if (do_color) {      // A1, A2 were overwritten by the colour tiles: recompute from S0
                    signal_z(); wait_acc(); ep_hidden(AR + p.sdf[0].b_off, X0);
                    signal_z(); wait_acc(); ep_hidden(AR + p.sdf[1].b_off, X1);
                }
                signal_z(); wait_acc(); ep_mask(X1, csb);
                signal_z(); wait_acc(); ep_mask(X0, csb + 64);
                signal_z(); wait_acc();
                float gt0, gt1;
                ep_ds0(sxw, gxw, inv_scale, gt0, gt1);
                if (part == 3 && topo_live) { gtopo[m] += gt0; gtopo[TM + m] += gt1; }
                tc_fence_before();
                bar_workers();
                grid_bwd_runs(gs, sxw, G, gr.g_emb_sdf, gxw, inv_scale, 0, tid);
                bar_workers();
            }

            // ---- finite-difference normal queries: 6 sub-tiles of (sample, query) rows, query index fastest ----
            // per sub-tile: S0 -> [MMA] -> A1 -> [MMA] -> A2 & dZ1 (fused epilogue) -> [MMA: wgrad2, wgrad1, dgrad1] -> dZ0
            //               -> [MMA: wgrad0, dgrad0] -> d(S0) -> frequency / grid backward           (4 round trips)
            if (need_fd) {
                const float* pt = (flags & MB_F_FD_WARPED) ? sxw : sx;
                float* gdst = (flags & MB_F_FD_WARPED) ? gxw : gx;
                uint8_t* DZ2 = smem + Smem::G;            // 16-column dZ2 tile (hi 4 KB | lo 4 KB) in the idle G scratch
#pragma unroll 1
                for (int j = 0; j <= 6; j++) {
                    if (tid < TM) {
                        if (j > 0) {      // fold the point gradients of sub-tile j-1 into the sample gradients (clamp derivative)
                            const int Q = (j - 1) * TM + tid, s = Q / 6, q = Q - s * 6;
                            const int axis = q >> 1;
                            const float e = (q & 1) ? -FD_EPS : FD_EPS;
#pragma unroll
                            for (int a = 0; a < 3; a++) {
                                float v = pt[a * TM + s];
                                if (a == axis) v = __fadd_rn(v, e);
                                if (v >= -p.bound && v <= p.bound) atomicAdd(gdst + a * TM + s, gpt[a * TM + tid]);
                            }
                        }
                        if (j < 6) {
                            const int Q = j * TM + tid, s = Q / 6, q = Q - s * 6;
                            const int axis = q >> 1;
                            const float e = (q & 1) ? -FD_EPS : FD_EPS;
#pragma unroll
                            for (int a = 0; a < 3; a++) {
                                float v = pt[a * TM + s];
                                if (a == axis) v = __fadd_rn(v, e);
                                spt[a * TM + tid] = fminf(fmaxf(v, -p.bound), p.bound);
                                gpt[a * TM + tid] = 0.f;
                            }
                            stq[tid] = stopo[s];
                            stq[TM + tid] = stopo[TM + s];
                        }
                    }
                    bar_workers();
                    if (j == 6) break;
                    build_s0(spt, stq);
                    signal_z(); wait_acc(); ep_hidden(AR + p.sdf[0].b_off, X0);
                    signal_z(); wait_acc();
                    {
                        const int Q = j * TM + m, s = Q / 6, q = Q - s * 6;
                        const float g0 = gsq[q * TM + s];
                        ep_hidden_fd(g0 * scale);
                        if (part == 0) {
                            const float v[8] = {g0 * scale, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                            store_core(DZ2, m, 0, v, 4096);
                            float sred = g0;
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) sred += __shfl_xor_sync(0xffffffffu, sred, o);
                            if (lane == 0 && sred != 0.f) red_add(GA + p.sdf[2].b_off, sred);
                        } else if (part == 1) {
                            store_core(DZ2, m, 1, z8, 4096);
                        }
                    }
                    signal_z(); wait_acc(); ep_mask(X0, csb + 64);
                    signal_z(); wait_acc();
                    float gt0, gt1;
                    ep_ds0(spt, gpt, inv_scale, gt0, gt1);
                    if (part == 3 && topo_live) {
                        const int s = (j * TM + m) / 6;
                        atomicAdd(gtopo + s, gt0);
                        atomicAdd(gtopo + TM + s, gt1);
                    }
                    tc_fence_before();
                    bar_workers();
                    grid_bwd_runs(gs, spt, G, gr.g_emb_sdf, gpt, inv_scale, j * TM, tid);
                    bar_workers();
                }
            }

            // ---- flush the per-tile weight-gradient accumulators ----
            if (do_main || need_fd) {
                flush_acc(128, p.sdf[0], 2, 80, 64, inv_scale);
                flush_acc(192, p.sdf[1], 0, 64, 64, inv_scale);
                flush_acc(256, p.sdf[2], 0, 64, do_main ? 33 : 1, inv_scale);
            }
            if (color_grad) {
                flush_acc(320, p.color[0], 0, 64, 64, inv_scale);
                flush_acc(384, p.color[1], 0, 64, 64, inv_scale);
                flush_acc(448, p.color[2], 0, 64, 3, inv_scale);
            }
            if (tid < 256) {      // bias gradients accumulated over the tile: sdf1 | sdf0 | col1 | col0
                const float sv = csb[tid];
                if (sv != 0.f) {
                    const uint32_t boff = (tid < 64) ? p.sdf[1].b_off : (tid < 128 ? p.sdf[0].b_off : (tid < 192 ? p.color[1].b_off : p.color[0].b_off));
                    red_add(GA + boff + (tid & 63), sv * inv_scale);
                }
            }
            tc_fence_before();
            bar_workers();

            // ---- outputs ----
            const bool skip_warp = (flags & MB_F_WARP) && (flags & MB_F_SKIP_WARP_BWD);
            for (int idx = tid; idx < 3 * TM; idx += NWORK) {
                const int mm = idx / 3, a = idx - mm * 3;
                if (mm < nv) {
                    if (gr.g_x) gr.g_x[(size_t)m0 * 3 + idx] = gx[a * TM + mm] + gxw[a * TM + mm];
                    if (skip_warp) {
                        float v = gxw[a * TM + mm];
                        if (gr.g_deform) v += gr.g_deform[(size_t)m0 * 3 + idx];
                        gr.g_def_out[(size_t)m0 * 3 + idx] = v;
                    }
                }
            }
            for (int idx = tid; idx < 2 * TM; idx += NWORK) {
                const int mm = idx / 2, a = idx - mm * 2;
                if (mm < nv) {
                    if (skip_warp) gr.g_topo_out[(size_t)m0 * 2 + idx] = gtopo[a * TM + mm];
                    if ((flags & MB_F_TOPO_IN) && gr.g_topo_in) gr.g_topo_in[(size_t)m0 * 2 + idx] = gtopo[a * TM + mm];
                }
            }
            bar_workers();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NWORK / 32) tmem_dealloc<512>(tmem);
}

}  // namespace tcs
}  // namespace mb

extern "C" int mb_field_backward_sdf_tc(const mb_field_params* p, const mb_field_io* io, const mb_field_grads* g, const void* tc_weights,
                                        const uint32_t* tc_off, const void* tc_weights_t, const uint32_t* tc_off_t, mb_stream_t stream) {
    using namespace mb;
    if (!p || !io || !g || !tc_weights || !tc_off || !tc_weights_t || !tc_off_t) { set_error("field_backward_sdf_tc: null argument"); return MB_EINVAL; }
    if (io->M == 0) return MB_OK;
    if (!io->x || !p->arena || !g->g_arena || !g->g_emb_sdf) { set_error("field_backward_sdf_tc: x/arena/g_arena/g_emb_sdf is null"); return MB_EINVAL; }
    if ((io->flags & MB_F_WARP) && (!(io->flags & MB_F_SKIP_WARP_BWD) || !g->deform || !g->topo || !g->g_def_out || !g->g_topo_out)) {
        set_error("field_backward_sdf_tc: WARP requires MB_F_SKIP_WARP_BWD with saved deform/topo and g_def_out/g_topo_out");
        return MB_EINVAL;
    }
    if ((io->flags & MB_F_FD) && !g->normal_raw) { set_error("field_backward_sdf_tc: FD needs the saved normal_raw"); return MB_EINVAL; }
    if ((io->flags & MB_F_COLOR) && !g->g_emb_col) { set_error("field_backward_sdf_tc: g_emb_col is null"); return MB_EINVAL; }
    if ((io->flags & MB_F_TOPO_IN) && !io->topo_in) { set_error("field_backward_sdf_tc: TOPO_IN needs topo_in"); return MB_EINVAL; }
    constexpr size_t smem = (size_t)tcs::Smem::TOTAL;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tcs::field_bwd_sdf_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("field_backward_sdf_tc: cannot reserve %zu B smem: %s", smem, cudaGetErrorString(e)); return MB_ECUDA; }
        attr_set = true;
    }
    const uint32_t n_tiles = div_up(io->M, tcs::TM);
    const uint32_t grid = min(n_tiles, (uint32_t)mb_sm_count());
    tcs::field_bwd_sdf_tc_kernel<<<grid, tcs::NTHREADS, smem, (cudaStream_t)stream>>>(*p, *io, *g, (const uint8_t*)tc_weights, tc_off,
                                                                                      (const uint8_t*)tc_weights_t, tc_off_t);
    return check_launch("field_backward_sdf_tc");
}
