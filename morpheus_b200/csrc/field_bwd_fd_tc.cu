// Tensor-core backward of the finite-difference normal queries (models/model.py:367-398 under autograd), specialised so
// that TWO CTAs fit on one SM: the general kernel (field_bwd_sdf_tc.cu) needs 221 KB of shared memory and therefore runs
// a single, latency-bound chain per SM; the FD queries are 12 of the 13 SDF queries of a real-view training sample.
//
//   * sub-tile = 96 rows = 16 samples x 6 (+-eps) queries (query index fastest): sample-aligned, so the per-cell merge of the
//     table scatter works on whole samples; operand tiles use a 96-row pitch (30 KB + 24 KB + 24 KB per CTA);
//   * an FD query only feeds output row 0 of the last SDF layer, so layer 2 never touches the tensor cores here: the layer-1
//     epilogue turns A2 = relu(acc + b1) directly into dZ1 = g0 * W2[0,:] * (A2 > 0) and into the weight-gradient row
//     dW2[0,:] += g0 * A2 (column sums by warp shuffles); A2 itself is never stored;
//   * per sub-tile: S0 -> [MMA] -> A1 -> [MMA] -> dZ1 -> [MMA: wgrad1, dgrad1] -> dZ0 -> [MMA: wgrad0, dgrad0] -> d(S0)
//     -> sin/cos backward + hash-grid backward (per-cell merged red.v2 scatter, corner-difference d/dx);
//   * bias gradients ride on the weight-gradient MMAs: pad column 39 of S0 and an extra column 64 of A1 hold the constant 1, so
//     row 39 / row 64 of the dW accumulators are the column sums of dZ0 / dZ1 (no shuffle reductions);
//   * weight-gradient accumulators (sdf0: 80 x 64, sdf1: 64 x 64) live in TMEM for a whole 128-sample tile (8 sub-tiles)
//     and are flushed once per tile; 256 TMEM columns per CTA.
// Gradients are scaled per tile by a power of two before the fp16 (hi, lo) split; 3 MMAs per product (tc_common.cuh).
#include "field_common.cuh"
#include "tc_common.cuh"
#include "tc_field.cuh"

namespace mb {
namespace tcf {

using namespace mb::tc;

constexpr int TM = 128;                 // samples per tile (one gradient scale, one accumulator flush)
constexpr int RT = 96;                  // rows per sub-tile
constexpr int NSUB = TM * 6 / RT;       // 8
constexpr int NWORK = 256;
constexpr int NTHREADS = NWORK + 64;    // + MMA-issue warp + weight-loader warp
constexpr int NSTAGE = 3;
constexpr int STAGE_BYTES = 5120;
constexpr int PITCH = RT * 16;          // bytes between 8-column core groups of a 96-row operand tile
constexpr int S0_LO = 10 * PITCH;       // lo offset of the 80-column S0 tile
constexpr int X_LO = 8 * PITCH;         // lo offset of a 64-column tile (dZ)
constexpr int X0_LO = 9 * PITCH;        // lo offset of the A1 tile: 64 columns + one core whose first column is the constant 1
constexpr bool PHASE_TIMING = false;    // true: worker thread 0 accumulates clock64 deltas per phase (mb_debug_fd_phases; ~10 % slower)

struct Smem {
    static constexpr int S0 = 0;                          // 30720
    static constexpr int X0 = S0 + 2 * S0_LO;             // 27648
    static constexpr int DZ = X0 + 2 * X0_LO;             // 24576; G (fp32 [32][96]) aliases it after the last MMA of a sub-tile
    static constexpr int W = DZ + 2 * X_LO;               // NSTAGE x 5120
    static constexpr int F = W + NSTAGE * STAGE_BYTES;
    static constexpr int SP = F;                          // [3][128] sample points (x or x + deform)
    static constexpr int STOPO = SP + 3 * 512;            // [2][128]
    static constexpr int GSQ = STOPO + 2 * 512;           // [6][128] d/d(sdf) of the six queries of a sample
    static constexpr int GACC = GSQ + 6 * 512;            // [3][128] d/d(sample point)
    static constexpr int GTOPO = GACC + 3 * 512;          // [2][128]
    static constexpr int SPT = GTOPO + 2 * 512;           // [3][96] row-wise query points
    static constexpr int GPT = SPT + 3 * 384;             // [3][96]
    static constexpr int STQ = GPT + 3 * 384;             // [2][96]
    static constexpr int CSB = STQ + 2 * 384;             // [128] bias-gradient accumulators sdf1 | sdf0
    static constexpr int CW2 = CSB + 512;                 // [64]  dW2[0, :] accumulator
    static constexpr int MISC = CW2 + 256;                // 16 floats
    static constexpr int BAR = MISC + 64;                 // full[2], empty[2], acc_ready, z_ready
    static constexpr int TMEMH = BAR + 8 * (2 * NSTAGE + 2);
    static constexpr int TOTAL = TMEMH + 16;
};
static_assert(Smem::BAR % 8 == 0, "alignment");
static_assert(Smem::TOTAL <= 113 * 1024, "two CTAs per SM");

// phase timing (clock64 deltas of worker thread 0 of every CTA, summed): read back with mb_debug_fd_phases()
__device__ unsigned long long g_fd_phase[16];

__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// column sums of a [32 lanes][32] register tile -> lane L holds the sum of column L
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

// hash-grid backward of one sub-tile: G[32][RT] = feature gradients of the 96 row-wise points; work item = (sample, level): the
// six +-eps rows of a sample that fall into the same cell are merged in registers (one red.v2 per corner per cell); the 16
// levels of a sample sit in the 16 lanes of a half-warp, d/d(point) is summed over levels by shuffles and added to gp.
__device__ __noinline__ void grid_bwd_samples(const GridCtx g, const float* __restrict__ pt3, const float* __restrict__ G, float* __restrict__ gemb,
                                               float* __restrict__ gp, float inv_scale, int tid) {
    const int l = tid & 15, s = tid >> 4;              // 256 threads = 16 samples x 16 levels
    const int nl = (int)min(g.n_levels, 16u);
    const bool live = l < nl;
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
    float dxr[6][3];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const int r = s * 6 + i;
        float dx[3] = {0.f, 0.f, 0.f};
        if (live) {
            const float g0 = G[(2 * l) * RT + r] * inv_scale, g1 = G[(2 * l + 1) * RT + r] * inv_scale;
            float u[3];
#pragma unroll
            for (int d = 0; d < 3; d++) u[d] = __fdiv_rn(__fadd_rn(pt3[d * RT + r], g.bound), g.two_bound);
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
        dxr[i][0] = dx[0]; dxr[i][1] = dx[1]; dxr[i][2] = dx[2];
    }
    if (have) {
#pragma unroll
        for (int c = 0; c < 8; c++) red_add2(gt + 2 * cidx[c], acc[2 * c], acc[2 * c + 1]);
    }
    // d/d(point) summed over the 16 levels (lanes of a half-warp) once, after the row loop (no warp-synchronous step per row)
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
            for (int d = 0; d < 3; d++) dxr[i][d] += __shfl_xor_sync(0xffffffffu, dxr[i][d], o);
        }
        if (l == 0) {
#pragma unroll
            for (int d = 0; d < 3; d++) gp[d * RT + s * 6 + i] += dxr[i][d];
        }
    }
}

__global__ void __launch_bounds__(NTHREADS, 2) field_bwd_fd_tc_kernel(const mb_field_params p, const mb_field_io io, const mb_field_grads gr,
                                                                     const uint8_t* __restrict__ tcw_f, const uint32_t* __restrict__ off_f,
                                                                     const uint8_t* __restrict__ tcw_d, const uint32_t* __restrict__ off_d,
                                                                     const int accumulate, const int nsub) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* sp = reinterpret_cast<float*>(smem + Smem::SP);
    float* stopo = reinterpret_cast<float*>(smem + Smem::STOPO);
    float* gsq = reinterpret_cast<float*>(smem + Smem::GSQ);
    float* gacc = reinterpret_cast<float*>(smem + Smem::GACC);
    float* gtopo = reinterpret_cast<float*>(smem + Smem::GTOPO);
    float* spt = reinterpret_cast<float*>(smem + Smem::SPT);
    float* gpt = reinterpret_cast<float*>(smem + Smem::GPT);
    float* stq = reinterpret_cast<float*>(smem + Smem::STQ);
    float* csb = reinterpret_cast<float*>(smem + Smem::CSB);
    float* cw2 = reinterpret_cast<float*>(smem + Smem::CW2);
    float* misc = reinterpret_cast<float*>(smem + Smem::MISC);
    float* G = reinterpret_cast<float*>(smem + Smem::DZ);
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
    const bool warped = (flags & MB_F_WARP) && (flags & MB_F_FD_WARPED);
    const bool topo_live = (flags & (MB_F_WARP | MB_F_TOPO_IN)) != 0;

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
    // a tile = nsub sub-tiles = 16 * nsub samples (<= 128): one gradient scale and one accumulator flush per tile; the host picks a
    // smaller tile for small M so that the persistent grid stays balanced (ray-sharded runs: 512 rays per GPU at 8 GPUs)
    const uint32_t TMt = 16u * (uint32_t)nsub;
    const uint32_t n_tiles = div_up(io.M, TMt);
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
            // D[work] (=) A[a_base : K-major, nk K-steps] * slabs(rows)^T
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
                    wait_z(); gemm_ring(s0_base, S0_LO, 5, 64, 64); umma_commit(acc_ready);                                   // A1 = S0 W0^T
                    wait_z(); gemm_ring(x0_base, X0_LO, 4, 64, 64); umma_commit(acc_ready);                                    // A2 = A1 W1^T
                    wait_z(); wgrad(x0_base, X0_LO, 192, j == 0); gemm_ring(dz_base, X_LO, 4, 64, 64); umma_commit(acc_ready);  // dW1, dA1
                    wait_z(); wgrad(s0_base, S0_LO, 128, j == 0); gemm_ring(dz_base, X_LO, 4, 80, 80); umma_commit(acc_ready); // dW0, dS0
                }
            }
        }
    } else {
        // ================================ workers (8 warps) ================================
        // TMEM epilogues: warp w reads lane quarter q = w & 3 (rows 32q .. 32q+31 < 96 for q < 3) and column half h = w >> 2
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
        unsigned long long ph[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        long long tprev = clock64();
#define FD_PHASE(i) do { if (PHASE_TIMING && tid == 0) { const long long tn = clock64(); ph[i] += (unsigned long long)(tn - tprev); tprev = tn; } } while (0)

        if (tid < RT) {      // core 8 of the A1 tile: column 64 = 1 (never overwritten: the A1 epilogue writes cores 0..7 only)
            const float one8[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            store_core(X0, tid, 8, one8, X0_LO, PITCH);
        }
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint32_t m0 = tile * TMt;
            const int nv = (int)min(TMt, io.M - m0);
            // ---- per-sample inputs, upstream -> d/d(sdf) of the six queries ----
            for (int idx = tid; idx < 3 * TM; idx += NWORK) {
                const int mm = idx / 3, a = idx - mm * 3;
                float v = 0.f;
                if (mm < nv) {
                    v = io.x[(size_t)m0 * 3 + idx];
                    if (warped) v += gr.deform[(size_t)m0 * 3 + idx];
                }
                sp[a * TM + mm] = v;
                gacc[a * TM + mm] = 0.f;
            }
            for (int idx = tid; idx < 2 * TM; idx += NWORK) {
                const int mm = idx / 2, a = idx - mm * 2;
                float v = 0.f;
                if (mm < nv) {
                    if (flags & MB_F_WARP) v = gr.topo[(size_t)m0 * 2 + idx];
                    else if (flags & MB_F_TOPO_IN) v = io.topo_in[(size_t)m0 * 2 + idx];
                }
                stopo[a * TM + mm] = v;
                gtopo[a * TM + mm] = 0.f;
            }
            if (tid < TM) {
                float g6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (tid < nv) {
                    const uint32_t gm = m0 + tid;
                    if (gr.g_fd) {
#pragma unroll
                        for (int qq = 0; qq < 6; qq++) g6[qq] = gr.g_fd[(size_t)gm * 6 + qq];
                    } else {
                        float gn[3] = {0.f, 0.f, 0.f};
                        if (gr.g_normal) { gn[0] = gr.g_normal[(size_t)gm * 3]; gn[1] = gr.g_normal[(size_t)gm * 3 + 1]; gn[2] = gr.g_normal[(size_t)gm * 3 + 2]; }
                        const float r0 = gr.normal_raw[(size_t)gm * 3], r1 = gr.normal_raw[(size_t)gm * 3 + 1], r2 = gr.normal_raw[(size_t)gm * 3 + 2];
                        const float d2 = r0 * r0 + r1 * r1 + r2 * r2;
                        const bool clamped = !(d2 > 1e-20f);
                        const float inv = 1.0f / sqrtf(fmaxf(d2, 1e-20f));
                        const float n[3] = {r0 * inv, r1 * inv, r2 * inv};
                        float graw[3];
                        if (clamped) { graw[0] = gn[0] * inv; graw[1] = gn[1] * inv; graw[2] = gn[2] * inv; }
                        else {
                            const float dot = n[0] * gn[0] + n[1] * gn[1] + n[2] * gn[2];
                            graw[0] = inv * (gn[0] - n[0] * dot); graw[1] = inv * (gn[1] - n[1] * dot); graw[2] = inv * (gn[2] - n[2] * dot);
                        }
                        if (gr.g_normal_raw) { graw[0] += gr.g_normal_raw[(size_t)gm * 3]; graw[1] += gr.g_normal_raw[(size_t)gm * 3 + 1]; graw[2] += gr.g_normal_raw[(size_t)gm * 3 + 2]; }
                        const float hh = 0.5f / FD_EPS;
#pragma unroll
                        for (int a = 0; a < 3; a++) { g6[2 * a] = graw[a] * hh; g6[2 * a + 1] = -graw[a] * hh; }
                    }
                }
                float mx = 0.f, gsum = 0.f;
#pragma unroll
                for (int qq = 0; qq < 6; qq++) { gsq[qq * TM + tid] = g6[qq]; mx = fmaxf(mx, fabsf(g6[qq])); gsum += g6[qq]; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
                }
                if (lane == 0) {
                    misc[warp] = mx;
                    if (gsum != 0.f) red_add(GA + p.sdf[2].b_off, gsum);       // bias gradient of the sdf output (row 0 of layer 2)
                }
                csb[tid] = 0.f;
                if (tid < 64) cw2[tid] = 0.f;
            }
            bar_workers();
            if (tid == 0) {
                const float mm = fmaxf(fmaxf(misc[0], misc[1]), fmaxf(misc[2], misc[3]));
                int e = 0;
                if (mm > 0.f && isfinite(mm)) { frexpf(mm, &e); e = 10 - e; }
                e = max(-100, min(100, e));
                misc[4] = ldexpf(1.0f, e);
                misc[5] = ldexpf(1.0f, -e);
            }
            bar_workers();
            const float scale = misc[4], inv_scale = misc[5];
            FD_PHASE(0);      // tile prologue

#pragma unroll 1
            for (int j = 0; j < nsub; j++) {
                // ---- rows of the sub-tile: sample s = 16 j + r / 6, query r % 6 ----
                if (tid < RT) {
                    const int s = 16 * j + tid / 6, qq = tid % 6;
                    const int axis = qq >> 1;
                    const float e = (qq & 1) ? -FD_EPS : FD_EPS;
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        float v = sp[a * TM + s];
                        if (a == axis) v = __fadd_rn(v, e);
                        spt[a * RT + tid] = fminf(fmaxf(v, -p.bound), p.bound);
                        gpt[a * RT + tid] = 0.f;
                    }
                    stq[tid] = stopo[s];
                    stq[RT + tid] = stopo[TM + s];
                }
                bar_workers();
                FD_PHASE(1);      // row setup + barrier
                // ---- S0: 16 levels x 96 rows in pairs of levels (768 items) + 3 axes x 96 rows of frequency features + pads ----
                for (int it = tid; it < 8 * RT; it += NWORK) {
                    const int r = it % RT, lp = it / RT;
                    const float pnt[3] = {spt[r], spt[RT + r], spt[2 * RT + r]};
                    gather_levels_tc(S0, r, 40, gs, 2 * lp, 2, pnt, S0_LO, PITCH);
                }
                for (int it = tid; it < 4 * RT; it += NWORK) {
                    const int r = it % RT, a = it / RT;
                    if (a < 3) {
                        freq_axis_tc(S0, r, a, spt[a * RT + r], (int)p.n_freq, S0_LO, PITCH);
                    } else {
                        store_one(S0, r, 39, 1.0f, S0_LO, PITCH);      // constant-1 pad feature (its W0 column is zero): db0 from the wgrad MMA
                        const float v[8] = {stq[r], stq[RT + r], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        store_core(S0, r, 9, v, S0_LO, PITCH);
                    }
                }
                FD_PHASE(2);      // gather + encodings
                // ---- A1 = relu(S0 W0^T + b0) -> X0 ----
                signal_z(); wait_acc();
                FD_PHASE(3);      // wait MMA fwd0
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
                FD_PHASE(4);      // epilogue A1
                // ---- A2 = relu(A1 W1^T + b1): dZ1 = g0 W2[0,:] (A2 > 0) -> DZ ; dW2[0,:] += g0 A2 ; db1 += dZ1 ----
                signal_z(); wait_acc();
                FD_PHASE(5);      // wait MMA fwd1
                {
                    float z[32], ga[32];
                    if (erow) {
                        float v[32];
                        tmem_ld32(tmem + lane_base + h * 32, v);
                        const int s = 16 * j + m / 6, qq = m % 6;
                        const float g0s = gsq[qq * TM + s] * scale;
                        const float* b1 = AR + p.sdf[1].b_off + h * 32;
                        const float* w2 = AR + p.sdf[2].w_off + h * 32;          // W[n = 0][k]
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
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i++) z[i] = ga[i] = 0.f;
                    }
                    if (q4 < 3) {       // warp-uniform
                        const float cg = warp_colsum32(ga, lane);
                        atomicAdd(cw2 + h * 32 + lane, cg);
                    }
                }
                FD_PHASE(6);      // fused epilogue A2 -> dZ1
                // ---- dZ0 = (dZ1 W1) (A1 > 0) -> DZ ; db0 += dZ0 ----
                signal_z(); wait_acc();
                FD_PHASE(7);      // wait MMA bwd1
                if (q4 < 3) {
                    float v[32];
                    tmem_ld32(tmem + lane_base + h * 32, v);
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const int kc = h * 4 + c;
                        const uint4 a = *reinterpret_cast<const uint4*>(X0 + kc * PITCH + (m >> 3) * 128 + (m & 7) * 16);
                        const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
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
                FD_PHASE(8);      // epilogue dZ0
                // ---- d(S0) (80 columns): h = 0: frequency columns 0..38 -> sin/cos backward ; h = 1: grid columns 40..71 -> G, topo 72..73 ----
                signal_z(); wait_acc();
                FD_PHASE(9);      // wait MMA bwd0
                if (erow) {
                    if (h == 0) {
                        float v[32], w[8];
                        tmem_ld32(tmem + lane_base, v);
                        tmem_ld8(tmem + lane_base + 32, w);
                        float f = 1.0f;
                        float acc3[3] = {v[0], v[1], v[2]};
#pragma unroll
                        for (int k = 0; k < 6; k++) {
                            if (k < (int)p.n_freq) {
#pragma unroll
                                for (int a = 0; a < 3; a++) {
                                    float sn, cn;
                                    sincosf(spt[a * RT + m] * f, &sn, &cn);
                                    const int is = 3 + 6 * k + a, ic = 6 + 6 * k + a;
                                    const float gs_ = is < 32 ? v[is] : w[is - 32];
                                    const float gc_ = ic < 32 ? v[ic] : w[ic - 32];
                                    acc3[a] += f * (gs_ * cn - gc_ * sn);
                                }
                            }
                            f *= 2.0f;
                        }
#pragma unroll
                        for (int a = 0; a < 3; a++) atomicAdd(gpt + a * RT + m, acc3[a] * inv_scale);
                    } else {
                        // G aliases DZ: the MMAs that read DZ completed before acc_ready fired
                        float v[32], t4[4];
                        tmem_ld32(tmem + lane_base + 40, v);
                        tmem_ld4(tmem + lane_base + 72, t4);
#pragma unroll
                        for (int i = 0; i < 32; i++) G[i * RT + m] = v[i];
                        if (topo_live) {
                            const int s = 16 * j + m / 6;
                            atomicAdd(gtopo + s, t4[0] * inv_scale);
                            atomicAdd(gtopo + TM + s, t4[1] * inv_scale);
                        }
                    }
                }
                tc_fence_before();
                bar_workers();
                FD_PHASE(10);     // epilogue d(S0) + barrier
                grid_bwd_samples(gs, spt, G, gr.g_emb_sdf, gpt, inv_scale, tid);
                bar_workers();
                FD_PHASE(11);     // table scatter + barrier
                // ---- fold the row gradients into the sample gradients (clamp derivative) ----
                if (tid < RT) {
                    const int s = 16 * j + tid / 6, qq = tid % 6;
                    const int axis = qq >> 1;
                    const float e = (qq & 1) ? -FD_EPS : FD_EPS;
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        float v = sp[a * TM + s];
                        if (a == axis) v = __fadd_rn(v, e);
                        if (v >= -p.bound && v <= p.bound) atomicAdd(gacc + a * TM + s, gpt[a * RT + tid]);
                    }
                }
            }

            FD_PHASE(12);         // fold (last sub-tile)
            // ---- flush the weight-gradient accumulators of the tile: rows = input features (tc order), 32 columns per warp half ----
            {
                const int krow = q4 * 32 + lane;
                {   // sdf layer 0: 80 rows (kind 2 order) x 64 columns at TMEM column 128
                    int korig = (krow < 80) ? tc_korig(2, krow) : -1;
                    if (korig >= (int)p.sdf[0].K) korig = -1;
                    if (q4 < 3) {
                        float v[32];
                        tmem_ld32(tmem + lane_base + 128 + h * 32, v);
                        if (korig >= 0) {
                            float* dst = GA + p.sdf[0].wt_off + (size_t)korig * p.sdf[0].N_pad + h * 32;
#pragma unroll
                            for (int c = 0; c < 8; c++) red_add4(dst + 4 * c, v[4 * c] * inv_scale, v[4 * c + 1] * inv_scale, v[4 * c + 2] * inv_scale, v[4 * c + 3] * inv_scale);
                        } else if (krow == 39) {       // the constant-1 feature: column sums of dZ0 = bias gradient of layer 0
                            float* dst = GA + p.sdf[0].b_off + h * 32;
#pragma unroll
                            for (int c = 0; c < 8; c++) red_add4(dst + 4 * c, v[4 * c] * inv_scale, v[4 * c + 1] * inv_scale, v[4 * c + 2] * inv_scale, v[4 * c + 3] * inv_scale);
                        }
                    }
                }
                {   // sdf layer 1: 64 rows x 64 columns at TMEM column 192
                    if (q4 < 3) {
                        float v[32];
                        tmem_ld32(tmem + lane_base + 192 + h * 32, v);
                        float* dst = (krow < 64) ? GA + p.sdf[1].wt_off + (size_t)krow * p.sdf[1].N_pad + h * 32
                                                 : GA + p.sdf[1].b_off + h * 32;      // row 64 = constant-1 feature: bias gradient of layer 1
                        if (krow <= 64) {
#pragma unroll
                            for (int c = 0; c < 8; c++) red_add4(dst + 4 * c, v[4 * c] * inv_scale, v[4 * c + 1] * inv_scale, v[4 * c + 2] * inv_scale, v[4 * c + 3] * inv_scale);
                        }
                    }
                }
            }
            if (tid >= 128 && tid < 192) {
                const float sv = cw2[tid - 128];       // dW2[0][k]: Wt slot [k][n = 0]
                if (sv != 0.f) red_add(GA + p.sdf[2].wt_off + (size_t)(tid - 128) * p.sdf[2].N_pad, sv * inv_scale);
            }
            tc_fence_before();
            bar_workers();

            // ---- outputs: d/dx (and d/d(deform) when the queries sit at the warped point), d/d(topo) ----
            const bool skip_warp = (flags & MB_F_WARP) && (flags & MB_F_SKIP_WARP_BWD);
            for (int idx = tid; idx < 3 * TM; idx += NWORK) {
                const int mm = idx / 3, a = idx - mm * 3;
                if (mm < nv) {
                    const float gv = gacc[a * TM + mm];
                    const size_t o = (size_t)m0 * 3 + idx;
                    if (gr.g_x) gr.g_x[o] = accumulate ? gr.g_x[o] + gv : gv;
                    if (skip_warp) {
                        if (accumulate) { if (warped) gr.g_def_out[o] += gv; }
                        else gr.g_def_out[o] = (warped ? gv : 0.f) + (gr.g_deform ? gr.g_deform[o] : 0.f);
                    }
                }
            }
            for (int idx = tid; idx < 2 * TM; idx += NWORK) {
                const int mm = idx / 2, a = idx - mm * 2;
                if (mm < nv) {
                    const float gv = gtopo[a * TM + mm];
                    const size_t o = (size_t)m0 * 2 + idx;
                    if (skip_warp) gr.g_topo_out[o] = accumulate ? gr.g_topo_out[o] + gv : gv + (gr.g_topo ? gr.g_topo[o] : 0.f);
                    if ((flags & MB_F_TOPO_IN) && gr.g_topo_in) gr.g_topo_in[o] = accumulate ? gr.g_topo_in[o] + gv : gv;
                }
            }
            bar_workers();
            FD_PHASE(13);         // flush + outputs
        }
        if (PHASE_TIMING && tid == 0) {
#pragma unroll
            for (int i = 0; i < 14; i++) atomicAdd(&g_fd_phase[i], ph[i]);
        }
#undef FD_PHASE
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NWORK / 32) tmem_dealloc<256>(tmem);
}

}  // namespace tcf
}  // namespace mb

extern "C" int mb_field_backward_fd_tc(const mb_field_params* p, const mb_field_io* io, const mb_field_grads* g, const void* tc_weights,
                                       const uint32_t* tc_off, const void* tc_weights_t, const uint32_t* tc_off_t, int accumulate, mb_stream_t stream) {
    // (declared with C linkage in include/morpheus_b200.h)
    using namespace mb;
    if (!p || !io || !g || !tc_weights || !tc_off || !tc_weights_t || !tc_off_t) { set_error("field_backward_fd_tc: null argument"); return MB_EINVAL; }
    if (io->M == 0) return MB_OK;
    if (!(io->flags & MB_F_FD)) { set_error("field_backward_fd_tc: flags lack MB_F_FD"); return MB_EINVAL; }
    if (!io->x || !p->arena || !g->g_arena || !g->g_emb_sdf) { set_error("field_backward_fd_tc: x/arena/g_arena/g_emb_sdf is null"); return MB_EINVAL; }
    if (!g->g_fd && !g->normal_raw) { set_error("field_backward_fd_tc: needs g_fd or the saved normal_raw"); return MB_EINVAL; }
    if (!g->g_fd && io->shading != MB_SHADE_ALBEDO && g->g_color) {
        set_error("field_backward_fd_tc: shading gradients reach the normals through the colour path: run mb_field_backward_sdf_tc with MB_F_FD_DELEGATE first");
        return MB_EINVAL;
    }
    if ((io->flags & MB_F_WARP) && (!(io->flags & MB_F_SKIP_WARP_BWD) || !g->deform || !g->topo || !g->g_def_out || !g->g_topo_out)) {
        set_error("field_backward_fd_tc: WARP requires MB_F_SKIP_WARP_BWD with saved deform/topo and g_def_out/g_topo_out");
        return MB_EINVAL;
    }
    if ((io->flags & MB_F_TOPO_IN) && !io->topo_in) { set_error("field_backward_fd_tc: TOPO_IN needs topo_in"); return MB_EINVAL; }
    constexpr size_t smem = (size_t)tcf::Smem::TOTAL;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tcf::field_bwd_fd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("field_backward_fd_tc: cannot reserve %zu B smem: %s", smem, cudaGetErrorString(e)); return MB_ECUDA; }
        attr_set = true;
    }
    const uint32_t slots = (uint32_t)mb_sm_count() * 2u;
    const uint32_t n128 = div_up(io->M, tcf::TM);
    const int nsub = n128 >= 4 * slots ? 8 : (n128 >= 2 * slots ? 4 : 2);       // samples per tile: 128 / 64 / 32
    const uint32_t n_tiles = div_up(io->M, 16u * (uint32_t)nsub);
    const uint32_t grid = min(n_tiles, slots);
    tcf::field_bwd_fd_tc_kernel<<<grid, tcf::NTHREADS, smem, (cudaStream_t)stream>>>(*p, *io, *g, (const uint8_t*)tc_weights, tc_off,
                                                                                     (const uint8_t*)tc_weights_t, tc_off_t, accumulate, nsub);
    return check_launch("field_backward_fd_tc");
}

/* debug: cumulative clock64 cycles per phase of mb_field_backward_fd_tc (worker thread 0 of every CTA); reset != 0 clears them */
extern "C" int mb_debug_fd_phases(unsigned long long* host_out16, int reset) {
    using namespace mb;
    unsigned long long z[16] = {0};
    if (host_out16 && cudaMemcpyFromSymbol(host_out16, tcf::g_fd_phase, sizeof(z)) != cudaSuccess) { set_error("debug_fd_phases: copy failed"); return MB_ECUDA; }
    if (reset && cudaMemcpyToSymbol(tcf::g_fd_phase, z, sizeof(z)) != cudaSuccess) { set_error("debug_fd_phases: reset failed"); return MB_ECUDA; }
    return MB_OK;
}
