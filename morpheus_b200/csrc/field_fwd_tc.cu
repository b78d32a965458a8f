// Tensor-core (tcgen05 + TMEM) fused scene-field forward.
//
// Same contract as field_fwd.cu (mb_field_forward) -- scene_representation.forward / density / normal / warp of
// /root/reference/models/model.py:273-307,367-398,412-437,439-533 in ONE launch -- but every dense layer runs on the
// 5th-generation tensor cores:
//   * tile = 128 samples = the M dimension of one tcgen05.mma (cta_group::1, M=128, N = layer width, K=16 per step);
//   * accumulators live in TMEM (128 fp32 columns per CTA, 2 CTAs per SM);
//   * activations stay in shared memory between layers as fp16 (hi, lo) pairs in the UMMA canonical K-major layout;
//     the epilogue (tcgen05.ld -> bias -> ReLU -> hi/lo split) writes the next layer's A operand in place;
//   * weights are pre-packed (mb_pack_tc) as fp16 (hi, lo) K=16 slabs in exactly the shared-memory byte order and
//     stream from L2 through a 4-stage ring of cp.async.bulk copies completing on mbarriers;
//   * 3 MMAs per K step (hi*hi + hi*lo + lo*hi) keep ~2^-22 relative accuracy (fp32 parity, see tc_common.cuh);
//   * warp roles: warps 0-7 build inputs (hash-grid gathers, encodings) and run epilogues, warp 8 lane 0 issues the
//     bulk copies and the MMAs; hand-offs are mbarriers (a_ready: 256 arrivals, acc_ready / empty[]: tcgen05.commit).
#include <stdlib.h>

#include "field_common.cuh"
#include "tc_common.cuh"
#include "tc_field.cuh"

namespace mb {
namespace tc {

constexpr int TM = 128;                 // samples per tile
constexpr int NWORK = 256;              // worker threads (8 warps)
constexpr int NTHREADS = NWORK + 64;    // + MMA-issue warp + weight-loader warp
constexpr int NSTAGE = 4;
constexpr int STAGE_BYTES = 8192;       // one K=16 slab of a 128-wide layer: (hi + lo) * 2 cores * 128 rows * 16 B
constexpr int MAX_OPS = 40;

struct Op { uint32_t src_off; uint16_t nk, n, n_pad, pad; };   // src_off: bytes into the tc weight arena

struct Smem {
    static constexpr int A = 0;                              // 65536: A_hi | A_lo
    static constexpr int W = 65536;                          // NSTAGE * 8192
    static constexpr int F = W + NSTAGE * STAGE_BYTES;       // fp32 per-sample scratch: 26 rows x 128
    static constexpr int SX = F;                             // [3][128]
    static constexpr int SXW = SX + 3 * 512;
    static constexpr int SPT = SXW + 3 * 512;
    static constexpr int STOPO = SPT + 3 * 512;              // [2][128]
    static constexpr int SDEF = STOPO + 2 * 512;             // [3][128]
    static constexpr int SSDF = SDEF + 3 * 512;
    static constexpr int SQ = SSDF + 512;                    // [6][128]
    static constexpr int ST = SQ + 6 * 512;
    static constexpr int SALB = ST + 512;                    // [3][128]
    static constexpr int STQ = SALB + 3 * 512;               // [2][128] per-row topo of an FD sub-tile
    static constexpr int PSUM = STQ + 2 * 512;               // [2][128] partial dot products of the FD row-0 shortcut
    static constexpr int OPS = PSUM + 2 * 512;               // Op[MAX_OPS]
    static constexpr int BAR = OPS + MAX_OPS * 12;           // mbarriers (8-byte aligned)
    static constexpr int TMEMH = BAR + 8 * (2 * NSTAGE + 2);
    static constexpr int TOTAL = TMEMH + 16;
};
static_assert(Smem::OPS % 4 == 0 && Smem::BAR % 8 == 0, "alignment");

struct Ctx {
    uint8_t* smem;
    uint64_t *full, *empty, *acc_ready, *a_ready;
    uint32_t tmem;
    uint32_t acc_count;   // worker side: ops completed (acc_ready phase)
};

// worker: wait for the accumulator of the current op
__device__ __forceinline__ void wait_acc(Ctx& c) {
    mbar_wait(c.acc_ready, c.acc_count & 1);
    c.acc_count++;
    tc_fence_after();
}
// worker: A operand for the next op is complete and the previous accumulator has been read
__device__ __forceinline__ void signal_a(Ctx& c) {
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(c.a_ready);     // one elected arrival per worker warp
}

// epilogue of a hidden layer: acc[:, 0:N] + bias -> ReLU -> next A operand (cores 0..N/8-1), columns split over the 2 warpgroups
template <int N>
__device__ __forceinline__ void epilogue_hidden(Ctx& c, const float* __restrict__ bias, int m, int wg, int warp_q) {
    uint8_t* A = c.smem + Smem::A;
    constexpr int HALF = N / 2;                 // columns per warpgroup (64 or 32)
    const int col0 = wg * HALF;
#pragma unroll
    for (int cb = 0; cb < HALF / 32; cb++) {
        float v[32];
        tmem_ld32(c.tmem + ((uint32_t)(warp_q * 32) << 16) + col0 + cb * 32, v);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; i++) o[i] = fmaxf(v[j * 8 + i] + __ldg(bias + col0 + cb * 32 + j * 8 + i), 0.f);
            store_core(A, m, (col0 + cb * 32) / 8 + j, o);
        }
    }
}

__global__ void __launch_bounds__(NTHREADS, 2) field_fwd_tc_kernel(const mb_field_params p, const mb_field_io io,
                                                                  const uint8_t* __restrict__ tcw, const uint32_t* __restrict__ tc_off,
                                                                  uint8_t* __restrict__ stash) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* sx = reinterpret_cast<float*>(smem + Smem::SX);
    float* sxw = reinterpret_cast<float*>(smem + Smem::SXW);
    float* spt = reinterpret_cast<float*>(smem + Smem::SPT);
    float* stopo = reinterpret_cast<float*>(smem + Smem::STOPO);
    float* sdef = reinterpret_cast<float*>(smem + Smem::SDEF);
    float* ssdf = reinterpret_cast<float*>(smem + Smem::SSDF);
    float* sq = reinterpret_cast<float*>(smem + Smem::SQ);
    float* st = reinterpret_cast<float*>(smem + Smem::ST);
    float* salb = reinterpret_cast<float*>(smem + Smem::SALB);
    float* stq = reinterpret_cast<float*>(smem + Smem::STQ);
    float* psum = reinterpret_cast<float*>(smem + Smem::PSUM);
    Op* ops = reinterpret_cast<Op*>(smem + Smem::OPS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::BAR);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + Smem::TMEMH);
    uint8_t* A = smem + Smem::A;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t flags = io.flags;
    const float* AR = p.arena;

    Ctx c;
    c.smem = smem;
    c.full = bars;
    c.empty = bars + NSTAGE;
    c.acc_ready = bars + 2 * NSTAGE;
    c.a_ready = bars + 2 * NSTAGE + 1;
    c.acc_count = 0;

    // ---- op list (identical for every tile): layer index into tc_off[] = order deform[6], topo[6], sdf[3], color[3], sdf2_row0 ----
    // tc_off[3*i+0] = byte offset, [3*i+1] = nk, [3*i+2] = n_pad ; entry 18 = sdf layer 2 restricted to N=16 (FD queries)
    __shared__ int n_ops_s;
    __shared__ LevelInfo s_levels[16];
    if (p.offsets) init_levels(s_levels, p.offsets, p.S, p.H);
    if (tid == 0) {
        int n = 0;
        auto push = [&](int layer, int nmma) {
            ops[n].src_off = tc_off[3 * layer];
            ops[n].nk = (uint16_t)tc_off[3 * layer + 1];
            ops[n].n_pad = (uint16_t)tc_off[3 * layer + 2];
            ops[n].n = (uint16_t)nmma;
            n++;
        };
        if (flags & MB_F_WARP)
            for (int net = 0; net < 2; net++) {
                for (int l = 0; l < 5; l++) push(net * 6 + l, 128);
                push(net * 6 + 5, 16);
            }
        if (flags & MB_F_MAIN) {
            push(12, 64); push(13, 64); push(14, 48);
            if (flags & MB_F_COLOR) { push(15, 64); push(16, 64); push(17, 16); }
        }
        if (flags & MB_F_FD)
            for (int q = 0; q < 6; q++) { push(12, 64); push(13, 64); }      // layer 2 (output row 0 only) is a dot product in the epilogue
        n_ops_s = n;
        for (int i = 0; i < NSTAGE; i++) { mbar_init(c.full + i, 1); mbar_init(c.empty + i, 1); }
        mbar_init(c.acc_ready, 1);
        mbar_init(c.a_ready, NWORK / 32);
        mbar_fence_init();
    }
    if (warp == NWORK / 32) tmem_alloc<128>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    c.tmem = *tmem_holder;
    const int n_ops = n_ops_s;
    const uint32_t n_tiles = div_up(io.M, TM);
    const uint32_t my_tiles = (blockIdx.x < n_tiles) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == NWORK / 32 + 1) {
        // =============================== loader warp: weight slabs -> ring (one bulk copy per K=16 slab) ===============================
        if (lane == 0 && my_tiles > 0 && n_ops > 0) {
            const uint64_t total_ops = (uint64_t)my_tiles * n_ops;
            uint32_t loads = 0;
            for (uint64_t u = 0; u < total_ops; u++) {
                const Op o = ops[u % n_ops];
                const uint32_t bytes = 64u * o.n_pad;     // (hi + lo) * 2 cores * n_pad rows * 16 B
                for (uint32_t st = 0; st < o.nk; st++) {
                    const uint32_t stg = loads % NSTAGE;
                    if (loads >= NSTAGE) mbar_wait(c.empty + stg, ((loads / NSTAGE) - 1) & 1);
                    mbar_arrive_expect_tx(c.full + stg, bytes);
                    bulk_g2s(smem + Smem::W + stg * STAGE_BYTES, tcw + o.src_off + (size_t)st * bytes, bytes, c.full + stg);
                    loads++;
                }
            }
        }
    } else if (warp == NWORK / 32) {
        // =============================== MMA-issue warp ===============================
        if (lane == 0 && my_tiles > 0 && n_ops > 0) {
            const uint64_t total_ops = (uint64_t)my_tiles * n_ops;
            uint32_t uses = 0, a_count = 0;
            const uint32_t a_base = smem_u32(A);
            const uint32_t w_base = smem_u32(smem + Smem::W);
            const uint64_t a_hi0 = make_smem_desc(a_base, 2048, 128);
            const uint64_t a_lo0 = make_smem_desc(a_base + A_LO_OFF, 2048, 128);
            for (uint64_t u_op = 0; u_op < total_ops; u_op++) {
                const Op o = ops[u_op % n_ops];
                // descriptors: per op the ring-stage bases (LBO depends on the layer width); per K step only the start address moves
                const uint64_t b_op = make_smem_desc(w_base, 16u * o.n_pad, 128);
                const uint64_t b_lo_add = (32u * o.n_pad) >> 4;
                const uint32_t idesc = make_idesc_f16(o.n);
                mbar_wait(c.a_ready, a_count & 1);
                a_count++;
                tc_fence_after();
                // training: stash the hidden activations A_1..A_5 of the deform / topology nets (the A operand of ops
                // 1..5 and 7..11 of a tile) for the tensor-core backward: one 64 KB bulk store per layer
                const uint32_t op_in_tile = (uint32_t)(u_op % n_ops);
                const bool do_stash = stash && (flags & MB_F_WARP) && op_in_tile < 12 && (op_in_tile % 6) >= 1;
                if (do_stash) {
                    const uint64_t tile = blockIdx.x + (u_op / n_ops) * (uint64_t)gridDim.x;
                    const uint32_t slot = (op_in_tile / 6) * 5 + (op_in_tile % 6) - 1;
                    bulk_s2g(stash + (tile * 10 + slot) * 65536ull, A, 65536);
                }
                for (uint32_t s = 0; s < o.nk; s++) {
                    const uint32_t stg = uses % NSTAGE;
                    mbar_wait(c.full + stg, (uses / NSTAGE) & 1);
                    tc_fence_after();
                    const uint64_t a_hi = a_hi0 + (uint64_t)s * 256, a_lo = a_lo0 + (uint64_t)s * 256;     // + 4096 B per K step
                    const uint64_t b_hi = b_op + (uint64_t)stg * (STAGE_BYTES >> 4), b_lo = b_hi + b_lo_add;
                    umma_f16(c.tmem, a_hi, b_hi, idesc, s > 0 ? 1u : 0u);
                    umma_f16(c.tmem, a_hi, b_lo, idesc, 1u);
                    umma_f16(c.tmem, a_lo, b_hi, idesc, 1u);
                    umma_commit(c.empty + stg);
                    uses++;
                }
                if (do_stash) bulk_store_wait_read();     // the epilogue of this op overwrites A
                umma_commit(c.acc_ready);
            }
        }
    } else {
        // =============================== worker warps ===============================
        const int m = tid & (TM - 1);
        const int wg = tid >> 7;            // warpgroup 0/1
        const int warp_q = warp & 3;        // TMEM lane quarter
        const GridCtx gs{p.emb_sdf, p.offsets, p.S, p.H, p.n_levels, p.bound, p.two_bound, s_levels};
        const GridCtx gc{p.emb_col, p.offsets, p.S, p.H, p.n_levels, p.bound, p.two_bound, s_levels};
        auto bar_workers = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory"); };

        // SDF-net input at point pt[3]: cores 0-4 freq, 5-8 grid, 9 topo
        // SDF-net input: the two warpgroups split the 16 grid levels 8 / 8 (each level is one dependent L2 round trip) and the
        // frequency features 2 axes / 1 axis + pad + topo
        auto build_sdf_input = [&](const float* pt3, const float* topo2) {
            const float pnt[3] = {pt3[m], pt3[TM + m], pt3[2 * TM + m]};
            if (wg == 0) {
#pragma unroll 1
                for (int a = 0; a < 2; a++) freq_axis_tc(A, m, a, pt3[a * TM + m], (int)p.n_freq);
                gather_levels_tc(A, m, 40, gs, 0, 8, pnt);
            } else {
                freq_axis_tc(A, m, 2, pt3[2 * TM + m], (int)p.n_freq);
                store_one(A, m, 39, 0.f);
                float v[8] = {topo2[m], topo2[TM + m], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                store_core(A, m, 9, v);
                gather_levels_tc(A, m, 40, gs, 8, 8, pnt);
            }
        };

        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint32_t m0 = tile * TM;
            const int nv = (int)min((uint32_t)TM, io.M - m0);
            // ---- load per-sample inputs ----
            for (int idx = tid; idx < 3 * TM; idx += NWORK) {
                const int mm = idx / 3, a = idx - mm * 3;
                sx[a * TM + mm] = (mm < nv) ? io.x[(size_t)m0 * 3 + idx] : 0.f;
            }
            if (tid < TM) {
                st[tid] = (io.t && tid < nv) ? io.t[m0 + tid] : 0.f;
                stopo[tid] = ((flags & MB_F_TOPO_IN) && tid < nv) ? io.topo_in[(size_t)(m0 + tid) * 2] : 0.f;
                stopo[TM + tid] = ((flags & MB_F_TOPO_IN) && tid < nv) ? io.topo_in[(size_t)(m0 + tid) * 2 + 1] : 0.f;
                sdef[tid] = sdef[TM + tid] = sdef[2 * TM + tid] = 0.f;
            }
            bar_workers();
            // ---- deformation + topology networks ----
            if (flags & MB_F_WARP) {
                for (int net = 0; net < 2; net++) {
                    const mb_layer_desc* L = net == 0 ? p.deform : p.topo;
                    if (wg == 0) {
#pragma unroll 1
                        for (int a = 0; a < 3; a++) freq_axis_tc(A, m, a, sx[a * TM + m], (int)p.n_freq);
                        store_one(A, m, 39, 0.f);
                    } else {
                        const float tt = st[m];
#pragma unroll 1
                        for (int cc = 0; cc < 6; cc++) {
                            float v[8];
#pragma unroll
                            for (int i = 0; i < 8; i++) { const int r = cc * 8 + i; v[i] = code_value(p, r >> 4, r & 15, tt); }
                            store_core(A, m, 5 + cc, v);
                        }
                        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        store_core(A, m, 11, z);
                    }
                    signal_a(c);
                    for (int l = 0; l < 5; l++) {
                        wait_acc(c);
                        epilogue_hidden<128>(c, AR + L[l].b_off, m, wg, warp_q);
                        signal_a(c);
                    }
                    wait_acc(c);
                    if (wg == 0) {
                        float v[16];
                        tmem_ld16(c.tmem + ((uint32_t)(warp_q * 32) << 16), v);
                        if (net == 0) {
#pragma unroll
                            for (int a = 0; a < 3; a++) sdef[a * TM + m] = v[a] + __ldg(AR + L[5].b_off + a);
                        } else {
                            stopo[m] = v[0] + __ldg(AR + L[5].b_off);
                            stopo[TM + m] = v[1] + __ldg(AR + L[5].b_off + 1);
                        }
                    }
                    tc_fence_before();
                    bar_workers();
                }
            }
            if (tid < TM) {
#pragma unroll
                for (int a = 0; a < 3; a++) sxw[a * TM + tid] = sx[a * TM + tid] + sdef[a * TM + tid];
            }
            bar_workers();
            // ---- main query ----
            if (flags & MB_F_MAIN) {
                build_sdf_input(sxw, stopo);
                signal_a(c);
                wait_acc(c);
                epilogue_hidden<64>(c, AR + p.sdf[0].b_off, m, wg, warp_q);
                signal_a(c);
                wait_acc(c);
                epilogue_hidden<64>(c, AR + p.sdf[1].b_off, m, wg, warp_q);
                signal_a(c);
                wait_acc(c);
                if (wg == 0) {
                    // h = acc[:, 0:33] + b : col 0 = sdf, cols 1..32 = geometric feature -> colour-net cores 4..7
                    float v[32], w[16];
                    tmem_ld32(c.tmem + ((uint32_t)(warp_q * 32) << 16), v);
                    tmem_ld16(c.tmem + ((uint32_t)(warp_q * 32) << 16) + 32, w);
                    const float* b2 = AR + p.sdf[2].b_off;
                    ssdf[m] = v[0] + __ldg(b2);
                    if (flags & MB_F_COLOR) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            float o[8];
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                const int col = 1 + j * 8 + i;
                                o[i] = (col < 32 ? v[col] : w[col - 32]) + __ldg(b2 + col);
                            }
                            store_core(A, m, 4 + j, o);
                        }
                    }
                }
                if (flags & MB_F_COLOR) {
                    const float pnt[3] = {sxw[m], sxw[TM + m], sxw[2 * TM + m]};
                    gather_levels_tc(A, m, 0, gc, wg * 8, 8, pnt);
                    signal_a(c);
                    wait_acc(c);
                    epilogue_hidden<64>(c, AR + p.color[0].b_off, m, wg, warp_q);
                    signal_a(c);
                    wait_acc(c);
                    epilogue_hidden<64>(c, AR + p.color[1].b_off, m, wg, warp_q);
                    signal_a(c);
                    wait_acc(c);
                    if (wg == 0) {
                        float v[16];
                        tmem_ld16(c.tmem + ((uint32_t)(warp_q * 32) << 16), v);
#pragma unroll
                        for (int a = 0; a < 3; a++) salb[a * TM + m] = 1.0f / (1.0f + expf(-(v[a] + __ldg(AR + p.color[2].b_off + a))));
                    }
                }
                tc_fence_before();
                bar_workers();
            }
            // ---- finite-difference normal: 6 SDF queries per sample, processed as 6 sub-tiles of 128 (sample, query) rows with
            //      the query index fastest, so that the six +-eps points of a sample sit in neighbouring lanes and their
            //      hash-grid gathers hit the same L1 sectors (same cell at all but the finest levels) ----
            if (flags & MB_F_FD) {
                const float* pt = (flags & MB_F_FD_WARPED) ? sxw : sx;
                for (int j = 0; j < 6; j++) {
                    if (tid < TM) {
                        const int Q = j * TM + tid, s = Q / 6, q = Q - s * 6;
                        const int axis = q >> 1;
                        const float e = (q & 1) ? -FD_EPS : FD_EPS;
#pragma unroll
                        for (int a = 0; a < 3; a++) {
                            float v = pt[a * TM + s];
                            if (a == axis) v = __fadd_rn(v, e);
                            spt[a * TM + tid] = fminf(fmaxf(v, -p.bound), p.bound);
                        }
                        stq[tid] = stopo[s];
                        stq[TM + tid] = stopo[TM + s];
                    }
                    bar_workers();
                    build_sdf_input(spt, stq);
                    signal_a(c);
                    wait_acc(c);
                    epilogue_hidden<64>(c, AR + p.sdf[0].b_off, m, wg, warp_q);
                    signal_a(c);
                    wait_acc(c);
                    {
                        // layer-1 epilogue fused with layer 2: an FD query only needs output row 0 (the sdf), i.e. the dot product
                        // of relu(acc + b1) with W2[0, :] -- exact fp32 FMAs, no third MMA round trip, no operand re-split
                        float v[32];
                        tmem_ld32(c.tmem + ((uint32_t)(warp_q * 32) << 16) + wg * 32, v);
                        const float* b1 = AR + p.sdf[1].b_off + wg * 32;
                        const float* w2 = AR + p.sdf[2].w_off + wg * 32;          // W[n = 0][k], n-major slot
                        float acc = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; i++) acc = __fmaf_rn(fmaxf(v[i] + __ldg(b1 + i), 0.f), __ldg(w2 + i), acc);
                        psum[wg * TM + m] = acc;
                    }
                    tc_fence_before();
                    bar_workers();
                    if (tid < TM) {
                        const int Q = j * TM + tid, s = Q / 6, q = Q - s * 6;
                        sq[q * TM + s] = __fadd_rn(__fadd_rn(psum[tid], psum[TM + tid]), __ldg(AR + p.sdf[2].b_off));
                    }
                }
                bar_workers();
            }
            // ---- per-sample outputs ----
            if (tid < nv) {
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
                    for (int a = 0; a < 3; a++) raw[a] = __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(sq[(2 * a) * TM + m], sq[(2 * a + 1) * TM + m])), FD_EPS);
                    const float d2 = raw[0] * raw[0] + raw[1] * raw[1] + raw[2] * raw[2];
                    const float inv = 1.0f / sqrtf(fmaxf(d2, 1e-20f));
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        float v = raw[a] * inv;
                        if (isnan(v)) v = 0.f;
                        else if (isinf(v)) v = v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
                        n[a] = v;
                        if (io.normal) io.normal[(size_t)gm * 3 + a] = v;
                        if (io.normal_raw) io.normal_raw[(size_t)gm * 3 + a] = raw[a];
                    }
                }
                if (io.color) {
                    float col[3] = {0.f, 0.f, 0.f};
                    if (flags & MB_F_COLOR) { col[0] = salb[m]; col[1] = salb[TM + m]; col[2] = salb[2 * TM + m]; }
                    if (io.shading != MB_SHADE_ALBEDO) {
                        float ndl = 0.f;
                        if (io.light) ndl = n[0] * io.light[(size_t)gm * 3] + n[1] * io.light[(size_t)gm * 3 + 1] + n[2] * io.light[(size_t)gm * 3 + 2];
                        const float lam = io.ratio + (1.0f - io.ratio) * fmaxf(ndl, 0.f);
                        if (io.shading == MB_SHADE_TEXTURELESS) col[0] = col[1] = col[2] = lam;
                        else if (io.shading == MB_SHADE_NORMAL) { col[0] = (n[0] + 1.f) * 0.5f; col[1] = (n[1] + 1.f) * 0.5f; col[2] = (n[2] + 1.f) * 0.5f; }
                        else { col[0] *= lam; col[1] *= lam; col[2] *= lam; }
                    }
                    io.color[(size_t)gm * 3] = col[0]; io.color[(size_t)gm * 3 + 1] = col[1]; io.color[(size_t)gm * 3 + 2] = col[2];
                }
                if (io.deform) { io.deform[(size_t)gm * 3] = sdef[m]; io.deform[(size_t)gm * 3 + 1] = sdef[TM + m]; io.deform[(size_t)gm * 3 + 2] = sdef[2 * TM + m]; }
                if (io.topo) { io.topo[(size_t)gm * 2] = stopo[m]; io.topo[(size_t)gm * 2 + 1] = stopo[TM + m]; }
            }
            bar_workers();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NWORK / 32) tmem_dealloc<128>(c.tmem);
}

// ---- weight packing: fp32 arena (W n-major [N_pad][K_pad]) -> fp16 (hi, lo) K=16 slabs in UMMA canonical order ----------
// slab s of a layer (bytes 64*N_pad): hi part [kcore 2][N_pad/8][8][8 halves], then lo part (same shape).
// K permutation of the first layers (inputs are built core-aligned by the kernel):
//   deform/topo L0: tc k 0..38 -> k, 39 -> 0-pad, 40..87 -> 39 + (k-40) (code), 88..95 -> pad
//   sdf L0        : tc k 0..38 -> k, 39 -> pad, 40..71 -> 39 + (k-40) (grid), 72,73 -> 71,72 (topo), 74..79 -> pad
__global__ void pack_tc_kernel(const float* __restrict__ arena, const uint32_t* __restrict__ desc /* per layer: w_off, K, K_pad, N_pad, kind, dst_off, K_tc */,
                               int n_layers, uint8_t* __restrict__ out) {
    const int layer = blockIdx.y;
    if (layer >= n_layers) return;
    const uint32_t* d = desc + layer * 8;
    // d = {src_off, K_valid, pitch, rows, kind, dst_off, inner, mode}
    //   mode 0 (forward B operand):  B[r = n][kk = k_tc] = W[n][korig(k_tc)]      src = W  n-major, pitch K_pad
    //   mode 1 (dgrad   B operand):  B[r = k_tc][kk = n] = Wt[korig(k_tc)][n]     src = Wt k-major, pitch N_pad
    const uint32_t w_off = d[0], K = d[1], K_pad = d[2], N_pad = d[3], kind = d[4], dst = d[5], K_tc = d[6], mode = d[7];
    const uint32_t total = N_pad * K_tc;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t n = i / K_tc, k = i - n * K_tc;
        const int ko = tc_korig((int)kind, (int)(mode == 0 ? k : n));
        float w = 0.f;
        if (ko >= 0 && (uint32_t)ko < K) w = (mode == 0) ? arena[w_off + (size_t)n * K_pad + ko] : arena[w_off + (size_t)ko * K_pad + k];
        const __half h = __float2half_rn(w);
        const __half l = __float2half_rn(w - __half2float(h));
        const uint32_t s = k >> 4, kc = (k >> 3) & 1, ki = k & 7;
        const size_t slab = (size_t)dst + (size_t)s * 64 * N_pad;
        const size_t off = (size_t)kc * (N_pad * 16) + (n >> 3) * 128 + (n & 7) * 16 + ki * 2;
        *reinterpret_cast<__half*>(out + slab + off) = h;
        *reinterpret_cast<__half*>(out + slab + 32 * N_pad + off) = l;
    }
}

}  // namespace tc
}  // namespace mb

extern "C" int mb_pack_tc(const float* arena, const uint32_t* layer_desc, int n_layers, void* out, mb_stream_t stream) {
    using namespace mb;
    if (!arena || !layer_desc || !out || n_layers <= 0) { set_error("pack_tc: bad argument"); return MB_EINVAL; }
    dim3 grid(8, n_layers);
    tc::pack_tc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(arena, layer_desc, n_layers, (uint8_t*)out);
    return check_launch("pack_tc");
}

extern "C" int mb_field_forward_tc(const mb_field_params* p, const mb_field_io* io, const void* tc_weights, const uint32_t* tc_off,
                                   void* stash, mb_stream_t stream) {
    using namespace mb;
    if (!p || !io || !tc_weights || !tc_off) { set_error("field_forward_tc: null argument"); return MB_EINVAL; }
    if (io->M == 0) return MB_OK;
    if (!io->x || !p->arena) { set_error("field_forward_tc: x/arena is null"); return MB_EINVAL; }
    if ((io->flags & MB_F_WARP) && !io->t) { set_error("field_forward_tc: WARP needs t"); return MB_EINVAL; }
    if ((io->flags & MB_F_TOPO_IN) && !io->topo_in) { set_error("field_forward_tc: TOPO_IN needs topo_in"); return MB_EINVAL; }
    if ((io->flags & MB_F_COLOR) && !(io->flags & MB_F_MAIN)) { set_error("field_forward_tc: COLOR needs MAIN"); return MB_EINVAL; }
    if (io->shading != MB_SHADE_ALBEDO && !(io->flags & MB_F_FD)) { set_error("field_forward_tc: shading needs FD normals"); return MB_EINVAL; }
    constexpr size_t smem = (size_t)tc::Smem::TOTAL + 128;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc::field_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("field_forward_tc: cannot reserve %zu B smem: %s", smem, cudaGetErrorString(e)); return MB_ECUDA; }
        attr_set = true;
    }
    const uint32_t n_tiles = div_up(io->M, tc::TM);
    static int ctas_per_sm = 0;      // MB_FWD_CTAS_PER_SM=1: occupancy experiment (default 2 resident CTAs per SM)
    if (!ctas_per_sm) { const char* e = getenv("MB_FWD_CTAS_PER_SM"); ctas_per_sm = (e && atoi(e) == 1) ? 1 : 2; }
    const uint32_t grid = min(n_tiles, (uint32_t)mb_sm_count() * (uint32_t)ctas_per_sm);
    tc::field_fwd_tc_kernel<<<grid, tc::NTHREADS, smem, (cudaStream_t)stream>>>(*p, *io, (const uint8_t*)tc_weights, tc_off, (uint8_t*)stash);
    return check_launch("field_forward_tc");
}
