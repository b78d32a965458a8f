// Tensor-core (tcgen05 + TMEM) backward of the deformation and topology networks -- the two 87->128x5->{3,2} MLPs of
// models/model.py:138-139 that carry ~2/3 of the per-sample FLOPs.  Per tile of 128 samples and per layer l = 5..0:
//     wgrad:  dW_l[k][n]  = sum_m A_l[m][k] * dZ_l[m][n]     (M = k, N = n, K = samples; both operands MN-major)
//     dgrad:  dA_l[m][k]  = sum_n dZ_l[m][n] * W_l[n][k]      (M = samples, N = k, K = n; K-major operands)
//     dZ_{l-1} = dA_l * (A_l > 0)
// A_l (l = 1..5) comes from the activation stash written by the forward kernel (64 KB fp16 hi/lo operand tile per
// layer, double-buffered bulk loads), A_0 is rebuilt (frequency encoding + deformation code).  dZ is scaled per tile
// by a power of two so the fp16 (hi, lo) split keeps ~22 bits; 3 MMAs per K step as in the forward engine.
// Accumulators: TMEM columns 0..127 = dgrad, 128..255 = wgrad.  Epilogues (8 worker warps): ReLU mask + hi/lo split
// of dZ (in place), bias gradients by an in-register 32x32 transpose-reduce, weight gradients by red.global.add.v4
// into the flat gradient arena, d(in0) -> frequency-encoding and deformation-code gradients.
#include "field_common.cuh"
#include "tc_common.cuh"
#include "tc_field.cuh"

namespace mb {
namespace tcb {

using namespace mb::tc;

constexpr int TM = 128;
constexpr int NWORK = 256;
constexpr int NTHREADS = NWORK + 64;    // + MMA-issue warp + weight-loader warp
constexpr int NSTAGE = 3;
constexpr int STAGE_BYTES = 8192;
constexpr int TILE_BYTES = 65536;

struct Smem {
    static constexpr int SA0 = 0;
    static constexpr int SA1 = TILE_BYTES;
    static constexpr int SZ = 2 * TILE_BYTES;
    static constexpr int W = 3 * TILE_BYTES;
    static constexpr int F = W + NSTAGE * STAGE_BYTES;     // fp32 scratch rows of 128
    static constexpr int SX = F;                           // [3][128]
    static constexpr int ST = SX + 3 * 512;                // [128]
    static constexpr int SG = ST + 512;                    // [3][128] upstream gradient of the net output
    static constexpr int CS = SG + 3 * 512;                // [128] column sums
    static constexpr int MISC = CS + 512;                  // [8] floats: warp maxima; [8] = scale
    static constexpr int BAR = MISC + 64;                  // full[3], empty[3], acc_ready, z_ready, sa_full[2]
    static constexpr int TMEMH = BAR + 8 * (2 * NSTAGE + 4);
    static constexpr int TOTAL = TMEMH + 16;
};
static_assert(Smem::BAR % 8 == 0, "alignment");
static_assert(Smem::TOTAL <= 232448, "shared memory budget");

// 32x32 transpose-reduce: every lane holds v[0..31] (one row, 32 columns); returns in lane j the sum over the 32 lanes of column j
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

__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ uint32_t make_idesc_f16_mn(uint32_t n) {   // both operands MN-major
    return make_idesc_f16(n) | (1u << 15) | (1u << 16);
}

__global__ void __launch_bounds__(NTHREADS, 1) field_bwd_warp_tc_kernel(const mb_field_params p, const float* __restrict__ gx_in,
                                                                       const float* __restrict__ gt_in, uint32_t M,
                                                                       const float* __restrict__ g_def, const float* __restrict__ g_topo,
                                                                       const uint8_t* __restrict__ stash, const uint8_t* __restrict__ tcw,
                                                                       const uint32_t* __restrict__ tc_off, float* __restrict__ g_arena,
                                                                       float* g_code0, float* g_code1, float* g_code2, float* __restrict__ g_x) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* sx = reinterpret_cast<float*>(smem + Smem::SX);
    float* st = reinterpret_cast<float*>(smem + Smem::ST);
    float* sg = reinterpret_cast<float*>(smem + Smem::SG);
    float* cs = reinterpret_cast<float*>(smem + Smem::CS);
    float* misc = reinterpret_cast<float*>(smem + Smem::MISC);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::BAR);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + Smem::TMEMH);
    uint64_t* full = bars;
    uint64_t* empty = bars + NSTAGE;
    uint64_t* acc_ready = bars + 2 * NSTAGE;
    uint64_t* z_ready = bars + 2 * NSTAGE + 1;
    uint64_t* sa_full = bars + 2 * NSTAGE + 2;   // [2]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; i++) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        mbar_init(acc_ready, 1);
        mbar_init(z_ready, NWORK / 32);
        mbar_init(sa_full, 1);
        mbar_init(sa_full + 1, 1);
        mbar_fence_init();
    }
    if (warp == NWORK / 32) tmem_alloc<256>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;
    const uint32_t n_tiles = div_up(M, TM);
    const uint32_t my_tiles = (blockIdx.x < n_tiles) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    // per layer (index net*6 + l) of the dgrad weight table: tc_off[3*i] = byte offset, [3*i+1] = number of K=16 slabs (N_pad/16), [3*i+2] = rows R (128 or 96)
    if (warp == NWORK / 32 + 1) {
        // ======================================= weight loader thread =======================================
        if (lane == 0 && my_tiles > 0) {
            uint32_t loads = 0;
            for (uint32_t it = 0; it < my_tiles; it++)
                for (uint32_t r = 0; r < 12; r++) {                   // (net, layer) units in processing order
                    const uint32_t li = (r / 6) * 6 + (5 - r % 6);
                    const uint32_t nk = tc_off[3 * li + 1], R = tc_off[3 * li + 2];
                    const uint32_t bytes = 64u * R;
                    const uint8_t* src = tcw + tc_off[3 * li];
                    for (uint32_t st = 0; st < nk; st++) {
                        const uint32_t stg = loads % NSTAGE;
                        if (loads >= NSTAGE) mbar_wait(empty + stg, ((loads / NSTAGE) - 1) & 1);
                        mbar_arrive_expect_tx(full + stg, bytes);
                        bulk_g2s(smem + Smem::W + stg * STAGE_BYTES, src + (size_t)st * bytes, bytes, full + stg);
                        loads++;
                    }
                }
        }
    } else if (warp == NWORK / 32) {
        // ======================================= MMA-issue thread =======================================
        if (lane == 0 && my_tiles > 0) {
            uint32_t uses = 0, z_count = 0, acc_count = 0, sa_count[2] = {0, 0};
            auto load_tile = [&](uint64_t tile, uint32_t net, uint32_t l, uint32_t buf) {   // A_l (l >= 1) = stash slot l-1
                mbar_arrive_expect_tx(sa_full + buf, TILE_BYTES);
                bulk_g2s(smem + (buf ? Smem::SA1 : Smem::SA0), stash + (tile * 10 + net * 5 + (l - 1)) * (uint64_t)TILE_BYTES, TILE_BYTES, sa_full + buf);
            };
            const uint32_t sz_base = smem_u32(smem + Smem::SZ);
            const uint32_t w_base = smem_u32(smem + Smem::W);
            // dZ tile as the MN-major B operand of wgrad (zw) and as the K-major A operand of dgrad (zd)
            const uint64_t zw_hi0 = make_smem_desc(sz_base, 128, 2048), zw_lo0 = make_smem_desc(sz_base + A_LO_OFF, 128, 2048);
            const uint64_t zd_hi0 = make_smem_desc(sz_base, 2048, 128), zd_lo0 = make_smem_desc(sz_base + A_LO_OFF, 2048, 128);
            for (uint32_t it = 0; it < my_tiles; it++) {
                const uint64_t tile = blockIdx.x + (uint64_t)it * gridDim.x;
                for (uint32_t net = 0; net < 2; net++) {
                    // SA0 is free (layer 1 of the previous net finished its epilogue before z_ready of its layer 0);
                    // SA1 held in0 of the previous net: free once that net's last MMAs completed
                    load_tile(tile, net, 5, 0);
                    if (acc_count > 0) mbar_wait(acc_ready, (acc_count - 1) & 1);
                    load_tile(tile, net, 4, 1);
                    for (int l = 5; l >= 0; l--) {
                        const uint32_t buf = (5 - l) & 1;
                        const uint32_t li = net * 6 + l;
                        mbar_wait(z_ready, z_count & 1);
                        z_count++;
                        tc_fence_after();
                        // the epilogue of layer l+1 is complete -> its A buffer (same parity as l-1) is free
                        if (l <= 4 && l - 1 >= 1) load_tile(tile, net, l - 1, (5 - (l - 1)) & 1);
                        if (l >= 1) { mbar_wait(sa_full + buf, sa_count[buf] & 1); sa_count[buf]++; }
                        const uint32_t sa_base = smem_u32(smem + (buf ? Smem::SA1 : Smem::SA0));
                        const uint32_t n_l = (l == 5) ? 16u : 128u;                 // N of layer l (padded)
                        // ---- wgrad: D_w[k][n] (TMEM cols 128..) ----
                        {
                            const uint32_t idesc = make_idesc_f16_mn(n_l);
                            const uint64_t a_hi0 = make_smem_desc(sa_base, 128, 2048), a_lo0 = make_smem_desc(sa_base + A_LO_OFF, 128, 2048);
#pragma unroll
                            for (uint32_t s = 0; s < 8; s++) {
                                const uint64_t a_hi = a_hi0 + s * 16, a_lo = a_lo0 + s * 16;          // + 256 B per K step (MN-major)
                                const uint64_t b_hi = zw_hi0 + s * 16, b_lo = zw_lo0 + s * 16;
                                umma_f16(tmem + 128, a_hi, b_hi, idesc, s > 0 ? 1u : 0u);
                                umma_f16(tmem + 128, a_hi, b_lo, idesc, 1u);
                                umma_f16(tmem + 128, a_lo, b_hi, idesc, 1u);
                            }
                        }
                        // ---- dgrad: D_a[m][k] (TMEM cols 0..) ----
                        {
                            const uint32_t nk = tc_off[3 * li + 1], R = tc_off[3 * li + 2];
                            const uint32_t idesc = make_idesc_f16(R);
                            const uint64_t b_op = make_smem_desc(w_base, 16u * R, 128);
                            const uint64_t b_lo_add = (32u * R) >> 4;
                            for (uint32_t s = 0; s < nk; s++) {
                                const uint32_t stg = uses % NSTAGE;
                                mbar_wait(full + stg, (uses / NSTAGE) & 1);
                                tc_fence_after();
                                const uint64_t a_hi = zd_hi0 + (uint64_t)s * 256, a_lo = zd_lo0 + (uint64_t)s * 256;   // + 4096 B per K step
                                const uint64_t b_hi = b_op + (uint64_t)stg * (STAGE_BYTES >> 4), b_lo = b_hi + b_lo_add;
                                umma_f16(tmem, a_hi, b_hi, idesc, s > 0 ? 1u : 0u);
                                umma_f16(tmem, a_hi, b_lo, idesc, 1u);
                                umma_f16(tmem, a_lo, b_hi, idesc, 1u);
                                umma_commit(empty + stg);
                                uses++;
                            }
                        }
                        umma_commit(acc_ready);
                        acc_count++;
                    }
                }
            }
        }
    } else {
        // ======================================= workers =======================================
        const int m = tid & (TM - 1);
        const int wg = tid >> 7;
        const int warp_q = warp & 3;
        const uint32_t lane_base = (uint32_t)(warp_q * 32) << 16;
        uint32_t acc_count = 0;
        auto bar_workers = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory"); };
        auto signal_z = [&]() { fence_proxy_async(); tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(z_ready); };
        uint8_t* SZ = smem + Smem::SZ;

        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint32_t m0 = tile * TM;
            const int nv = (int)min((uint32_t)TM, M - m0);
            for (int idx = tid; idx < 3 * TM; idx += NWORK) {
                const int mm = idx / 3, a = idx - mm * 3;
                sx[a * TM + mm] = (mm < nv) ? gx_in[(size_t)m0 * 3 + idx] : 0.f;
            }
            if (tid < TM) st[tid] = (tid < nv) ? gt_in[m0 + tid] : 0.f;
            bar_workers();
            // all samples of the tile at the same time step (the normal case: one frame per batch)?
            if (tid == 0) misc[10] = 0.f;
            bar_workers();
            if (tid < nv && __float_as_uint(st[tid]) != __float_as_uint(st[0])) misc[10] = 1.f;
            bar_workers();
            const bool t_uniform = misc[10] == 0.f;

            for (int net = 0; net < 2; net++) {
                const mb_layer_desc* L = net == 0 ? p.deform : p.topo;
                const int nout = net == 0 ? 3 : 2;
                // ---- upstream gradient of the net output, per-tile power-of-two scale ----
                if (tid < TM) {
                    float mx = 0.f;
                    for (int a = 0; a < 3; a++) {
                        float v = 0.f;
                        if (a < nout && tid < nv) v = net == 0 ? g_def[(size_t)(m0 + tid) * 3 + a] : g_topo[(size_t)(m0 + tid) * 2 + a];
                        sg[a * TM + tid] = v;
                        mx = fmaxf(mx, fabsf(v));
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    if (lane == 0) misc[warp] = mx;
                }
                cs[tid & 127] = 0.f;
                bar_workers();
                if (tid == 0) {
                    const float mx = fmaxf(fmaxf(misc[0], misc[1]), fmaxf(misc[2], misc[3]));
                    int e = 0;
                    if (mx > 0.f && isfinite(mx)) { frexpf(mx, &e); e = 10 - e; }     // mx * 2^e in [512, 1024)
                    e = max(-100, min(100, e));
                    misc[8] = ldexpf(1.0f, e);
                    misc[9] = ldexpf(1.0f, -e);
                }
                bar_workers();
                const float scale = misc[8], inv_scale = misc[9];
                // dZ_5 tile: 16 columns (2 cores); bias gradient of layer 5
                if (wg == 0) {
                    float v[8] = {sg[m] * scale, sg[TM + m] * scale, (nout == 3 ? sg[2 * TM + m] * scale : 0.f), 0.f, 0.f, 0.f, 0.f, 0.f};
                    store_core(SZ, m, 0, v);
                } else {
                    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    store_core(SZ, m, 1, z);
                }
                if (tid < TM) {
                    for (int a = 0; a < nout; a++) {
                        float s = sg[a * TM + tid];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                        if (lane == 0 && s != 0.f) red_add(g_arena + L[5].b_off + a, s);
                    }
                }
                signal_z();

                for (int l = 5; l >= 0; l--) {
                    const uint32_t buf = (5 - l) & 1;
                    const uint8_t* SA = smem + (buf ? Smem::SA1 : Smem::SA0);
                    mbar_wait(acc_ready, acc_count & 1);
                    acc_count++;
                    tc_fence_after();
                    // ---------------- wgrad epilogue: D_w[k = this thread's row][n] -> gradient arena ----------------
                    {
                        const int krow = warp_q * 32 + lane;            // row of D_w = input feature (tc order for l == 0)
                        int korig = krow;
                        if (l == 0) korig = (krow < 96) ? tc_korig(1, krow) : -1;
                        float* gw = g_arena + L[l].wt_off;
                        const int npad = (int)L[l].N_pad;
                        if (l == 5) {
                            if (wg == 0) {
                                float v[16];
                                tmem_ld16(tmem + lane_base + 128, v);
                                for (int a = 0; a < nout; a++) red_add(gw + (size_t)korig * npad + a, v[a] * inv_scale);
                            }
                        } else {
#pragma unroll 1
                            for (int cb = 0; cb < 2; cb++) {
                                float v[32];
                                const int col0 = wg * 64 + cb * 32;
                                tmem_ld32(tmem + lane_base + 128 + col0, v);
                                if (korig >= 0) {
                                    float* dst = gw + (size_t)korig * npad + col0;
#pragma unroll
                                    for (int j = 0; j < 8; j++)
                                        red_add4(dst + 4 * j, v[4 * j] * inv_scale, v[4 * j + 1] * inv_scale, v[4 * j + 2] * inv_scale, v[4 * j + 3] * inv_scale);
                                }
                            }
                        }
                    }
                    // ---------------- dgrad epilogue ----------------
                    if (l >= 1) {
                        // dZ_{l-1}[m][k] = D_a[m][k] * (A_l[m][k] > 0); bias gradient of layer l-1 = column sums
#pragma unroll 1
                        for (int cb = 0; cb < 2; cb++) {
                            float v[32];
                            const int col0 = wg * 64 + cb * 32;
                            tmem_ld32(tmem + lane_base + col0, v);
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const int kc = col0 / 8 + j;
                                const uint4 a = *reinterpret_cast<const uint4*>(SA + kc * 2048 + (m >> 3) * 128 + (m & 7) * 16);
                                const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
                                float o[8];
#pragma unroll
                                for (int i = 0; i < 8; i++) {
                                    const uint16_t hbits = (uint16_t)(aw[i >> 1] >> ((i & 1) * 16));
                                    const bool pos = (hbits & 0x7FFF) != 0 && !(hbits & 0x8000);
                                    o[i] = pos ? v[j * 8 + i] : 0.f;
                                    v[j * 8 + i] = o[i];
                                }
                                store_core(SZ, m, kc, o);
                            }
                            const float csum = warp_colsum32(v, lane);
                            atomicAdd(cs + col0 + lane, csum);
                        }
                        bar_workers();
                        if (tid < TM) {
                            const float s = cs[tid];
                            if (s != 0.f) red_add(g_arena + L[l - 1].b_off + tid, s * inv_scale);
                            cs[tid] = 0.f;
                        }
                        if (l == 1) {
                            // rebuild in0 (A_0) in SA1: cores 0-4 freq(x), 5-10 code(t), 11-15 zero
                            uint8_t* S1 = smem + Smem::SA1;
                            if (wg == 0) {
                                const float pnt[3] = {sx[m], sx[TM + m], sx[2 * TM + m]};
                                build_freq_tc(S1, m, pnt, (int)p.n_freq);
                                const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                                store_core(S1, m, 11, z); store_core(S1, m, 12, z); store_core(S1, m, 13, z);
                            } else {
                                const float tt = st[m];
#pragma unroll 1
                                for (int cc = 0; cc < 6; cc++) {
                                    float v[8];
#pragma unroll
                                    for (int i = 0; i < 8; i++) { const int r = cc * 8 + i; v[i] = code_value(p, r >> 4, r & 15, tt); }
                                    store_core(S1, m, 5 + cc, v);
                                }
                                const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                                store_core(S1, m, 14, z); store_core(S1, m, 15, z);
                            }
                        }
                        signal_z();
                    } else {
                        // l == 0: D_a = d(in0) [m][k_tc < 96]: wg0 has columns 0..47, wg1 48..95
                        if (wg == 0) {
                            float v[32], w[16];
                            tmem_ld32(tmem + lane_base, v);
                            tmem_ld16(tmem + lane_base + 32, w);
                            // frequency-encoding backward (columns 0..38): g_x[a] += g0 + sum_k f_k (g_sin cos - g_cos sin)
                            if (m < nv) {
                                float f = 1.0f;
                                float acc[3] = {v[0], v[1], v[2]};
#pragma unroll
                                for (int k = 0; k < 6; k++) {
                                    if (k < (int)p.n_freq) {
#pragma unroll
                                        for (int a = 0; a < 3; a++) {
                                            float sn, cn;
                                            sincosf(sx[a * TM + m] * f, &sn, &cn);
                                            const int is = 3 + 6 * k + a, ic = 6 + 6 * k + a;
                                            const float gs_ = is < 32 ? v[is] : w[is - 32];
                                            const float gc_ = ic < 32 ? v[ic] : w[ic - 32];
                                            acc[a] += f * (gs_ * cn - gc_ * sn);
                                        }
                                    }
                                    f *= 2.0f;
                                }
#pragma unroll
                                for (int a = 0; a < 3; a++) g_x[(size_t)(m0 + m) * 3 + a] += acc[a] * inv_scale;
                            }
                            // code columns 40..47 -> code rows 0..7 (w[8..15])
                            float cv[32];
#pragma unroll
                            for (int i = 0; i < 32; i++) cv[i] = (i < 8) ? w[8 + i] : 0.f;
                            if (t_uniform) {
                                const float s = warp_colsum32(cv, lane);
                                if (lane < 8) atomicAdd(cs + lane, s);
                            } else if (m < nv) {
#pragma unroll 1
                                for (int i = 0; i < 8; i++) {
                                    const int r = i, vv = r >> 4, c = r & 15;
                                    const int S = (int)p.code_len[vv];
                                    const float t = fminf(fmaxf(st[m], 0.f), 1.f);
                                    const float pos = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(t, 2.f), 1.f), 1.f), 0.5f), (float)(S - 1));
                                    const int i0 = min(max((int)floorf(pos), 0), S - 1);
                                    const float w1 = pos - (float)i0, w0 = 1.f - w1;
                                    float* gl = (vv == 0 ? g_code0 : (vv == 1 ? g_code1 : g_code2)) + (size_t)c * S;
                                    const float gv = w[8 + i] * inv_scale;
                                    atomicAdd(gl + i0, gv * w0);
                                    if (i0 + 1 <= S - 1) atomicAdd(gl + i0 + 1, gv * w1);
                                }
                            }
                        } else {
                            float v[32], w[16];
                            tmem_ld32(tmem + lane_base + 48, v);     // columns 48..79 -> code rows 8..39
                            tmem_ld16(tmem + lane_base + 80, w);     // columns 80..95 -> code rows 40..47 (+ 8 pads)
                            if (t_uniform) {
                                float tmp[32];
#pragma unroll
                                for (int i = 0; i < 32; i++) tmp[i] = v[i];
                                const float s1 = warp_colsum32(tmp, lane);
                                atomicAdd(cs + 8 + lane, s1);
#pragma unroll
                                for (int i = 0; i < 32; i++) tmp[i] = (i < 8) ? w[i] : 0.f;
                                const float s2 = warp_colsum32(tmp, lane);
                                if (lane < 8) atomicAdd(cs + 40 + lane, s2);
                            } else if (m < nv) {
#pragma unroll 1
                                for (int i = 0; i < 40; i++) {
                                    const int r = 8 + i, vv = r >> 4, c = r & 15;
                                    const int S = (int)p.code_len[vv];
                                    const float t = fminf(fmaxf(st[m], 0.f), 1.f);
                                    const float pos = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(t, 2.f), 1.f), 1.f), 0.5f), (float)(S - 1));
                                    const int i0 = min(max((int)floorf(pos), 0), S - 1);
                                    const float w1 = pos - (float)i0, w0 = 1.f - w1;
                                    float* gl = (vv == 0 ? g_code0 : (vv == 1 ? g_code1 : g_code2)) + (size_t)c * S;
                                    const float gv = (i < 32 ? v[i] : w[i - 32]) * inv_scale;
                                    atomicAdd(gl + i0, gv * w0);
                                    if (i0 + 1 <= S - 1) atomicAdd(gl + i0 + 1, gv * w1);
                                }
                            }
                        }
                        tc_fence_before();
                        bar_workers();
                        if (t_uniform && tid < 48) {
                            const float s = cs[tid] * inv_scale;
                            if (s != 0.f) {
                                const int vv = tid >> 4, c = tid & 15;
                                const int S = (int)p.code_len[vv];
                                const float t = fminf(fmaxf(st[0], 0.f), 1.f);
                                const float pos = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(t, 2.f), 1.f), 1.f), 0.5f), (float)(S - 1));
                                const int i0 = min(max((int)floorf(pos), 0), S - 1);
                                const float w1 = pos - (float)i0, w0 = 1.f - w1;
                                float* gl = (vv == 0 ? g_code0 : (vv == 1 ? g_code1 : g_code2)) + (size_t)c * S;
                                atomicAdd(gl + i0, s * w0);
                                if (i0 + 1 <= S - 1) atomicAdd(gl + i0 + 1, s * w1);
                            }
                        }
                        bar_workers();
                        if (tid < TM) cs[tid] = 0.f;
                    }
                }
            }
            bar_workers();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NWORK / 32) tmem_dealloc<256>(tmem);
}

}  // namespace tcb
}  // namespace mb

extern "C" int mb_field_backward_warp_tc(const mb_field_params* p, const float* x, const float* t, uint32_t M, const float* g_def,
                                         const float* g_topo, const void* stash, const void* tc_weights_t, const uint32_t* tc_off_t,
                                         float* g_arena, float* const g_code[3], float* g_x, mb_stream_t stream) {
    using namespace mb;
    if (!p || !x || !t || !g_def || !g_topo || !stash || !tc_weights_t || !tc_off_t || !g_arena || !g_code || !g_x) {
        set_error("field_backward_warp_tc: null argument");
        return MB_EINVAL;
    }
    if (M == 0) return MB_OK;
    constexpr size_t smem = (size_t)tcb::Smem::TOTAL;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tcb::field_bwd_warp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("field_backward_warp_tc: cannot reserve %zu B smem: %s", smem, cudaGetErrorString(e)); return MB_ECUDA; }
        attr_set = true;
    }
    const uint32_t n_tiles = div_up(M, tcb::TM);
    const uint32_t grid = min(n_tiles, (uint32_t)mb_sm_count());
    tcb::field_bwd_warp_tc_kernel<<<grid, tcb::NTHREADS, smem, (cudaStream_t)stream>>>(
        *p, x, t, M, g_def, g_topo, (const uint8_t*)stash, (const uint8_t*)tc_weights_t, tc_off_t, g_arena, g_code[0], g_code[1], g_code[2], g_x);
    return check_launch("field_backward_warp_tc");
}
