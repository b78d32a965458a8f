// Implicit-GEMM 3x3 (stride 1, pad 1) / 1x1 convolution on the 5th-generation tensor cores (tcgen05 + TMEM) for the frozen networks of
// the SDS step: the SD-1.x UNet (ldm/modules/diffusionmodules/openaimodel.py:745-777, ResBlock convs :256-276, proj_in/out of
// ldm/modules/attention.py:239-253) and the KL-VAE encoder (ldm/modules/diffusionmodules/model.py:434-459, ResnetBlock :82-140) --
// forward, and the input-gradient ("dgrad") backward of the VAE, which is the same kernel on transposed / flipped weights.  In strict fp32
// the reference path runs these on cuDNN's fp32 SIMT kernels (~10 TFLOP/s: profiles/r02_sds_launches_summary.md); here
//     out[p, n] = sum_{tap, c} x[p + offset(tap), c] * W[n, c, tap]            p = (b, y, x) pixel, M = B H W rows, N = C_out, K = 9 C_in
// is evaluated as THREE kind::f16 MMAs per K step on fp16 (hi, lo) splits (x_hi w_hi + x_hi w_lo + x_lo w_hi, fp32 accumulation in TMEM,
// ~2^-22 per product: the same numerics contract as the field kernels, tc_common.cuh).
//
//   * activations are pre-split once per layer into NHWC fp16 planes (hi, lo) by nchw_split_kernel (optionally fused with the SiLU that
//     precedes every such convolution in both networks): one pixel's 8 consecutive channels are 16 contiguous bytes = exactly one row of a
//     UMMA core matrix, so the im2col gather is a cp.async of 16-byte chunks with zero-fill at the image border -- no im2col buffer;
//   * weights are packed ONCE (frozen) per (n-tile, tap, 64-channel block) into the UMMA canonical K-major byte order (hi | lo), so a
//     pipeline stage is one contiguous cp.async.bulk of 256 * N_TILE bytes;
//   * tile = 128 pixels x N_TILE (128 or 160) output channels, K stage = 64 channels of one tap (4 K=16 steps), 3-stage ring (MB_CONV_KC /
//     NSTG / LAG; measured alternatives on the SDS chain: 32 channels x 6 stages x lag 4 -> 19.05 ms, 64 x 3 x lag 2 -> 18.99 ms, this shape 18.4 ms);
//     warps 0-3 gather A and later run the epilogue (TMEM -> + bias -> coalesced NCHW stores, 32 consecutive pixels per store),
//     warp 4 streams B, warp 5 issues the MMAs; split-K over the (tap, channel-block) stages (gridDim.z) with red.global.add fills
//     the machine for the small-M layers of the UNet (M = 2 x 32 x 32 ... 2 x 4 x 4).
#include "tc_common.cuh"

namespace mb {
namespace conv {

using namespace mb::tc;

constexpr int TM = 128;             // pixels per tile
#ifndef MB_CONV_KC
#define MB_CONV_KC 64
#endif
#ifndef MB_CONV_NSTG
#define MB_CONV_NSTG 3
#endif
#ifndef MB_CONV_LAG
#define MB_CONV_LAG 1
#endif
constexpr int KC = MB_CONV_KC;      // channels per pipeline stage (one tap); weights are packed in 64-channel blocks = 64 / KC stages
constexpr int KB = 64;              // channel block of the weight pack / of the host-side divisibility rule
constexpr int NSTG = MB_CONV_NSTG;  // ring depth
constexpr int LAG = MB_CONV_LAG;    // a producer publishes stage i - LAG after issuing stage i (LAG + 1 cp.async groups in flight per thread)
static_assert(LAG >= 1 && LAG < NSTG && KB % KC == 0 && KC % 16 == 0, "pipeline shape");
constexpr int A_STAGE = 2 * (KC / 8) * TM * 16;      // hi + lo: 8 K-cores x 128 rows x 16 B each = 32768
constexpr int A_LO = A_STAGE / 2;
constexpr int NTHREADS = 192;       // 4 producer / epilogue warps + B loader warp + MMA warp

template <int N_TILE>
struct Smem {
    static constexpr int B_STAGE = (KC / 16) * 64 * N_TILE;            // KC / 16 slabs x (hi + lo) x 2 K-cores x N_TILE rows x 16 B
    static constexpr int A = 0;
    static constexpr int B = NSTG * A_STAGE;
    static constexpr int BAR = B + NSTG * B_STAGE;                     // full_a[NSTG], full_b[NSTG], empty[NSTG], acc_ready
    static constexpr int TMEMH = BAR + 8 * (3 * NSTG + 1);
    static constexpr int TOTAL = TMEMH + 16;
};

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes) : "memory");
}

struct Args {
    const __half* xh;        // [B, H, W, Cin] hi plane
    const __half* xl;        // lo plane
    const uint8_t* wpk;      // packed weights (pack_conv_kernel)
    const float* bias;       // [Cout] or NULL
    float* out;              // [B, Cout, H, W] fp32 (NCHW)
    int B, H, W, Cin, Cout, ntaps;      // ntaps = 9 (3x3, pad 1) or 1 (1x1)
    float out_mul;           // epilogue factor: 1 / (weight scale chosen at pack time)
    const float* out_mul_dev;   // optional device factor: 1 / (activation scale applied by nchw_split_kernel), or NULL
    int out_rows;            // != 0: out is row-major [B H W, Cout] (token-major linear layer) instead of NCHW
};

template <int N_TILE>
__global__ void __launch_bounds__(NTHREADS, 1) conv_tc_kernel(const Args a) {
    using S = Smem<N_TILE>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR);
    uint64_t* full_a = bars;
    uint64_t* full_b = bars + NSTG;
    uint64_t* empty = bars + 2 * NSTG;
    uint64_t* acc_ready = bars + 3 * NSTG;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + S::TMEMH);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < NSTG; i++) { mbar_init(full_a + i, TM); mbar_init(full_b + i, 1); mbar_init(empty + i, 1); }
        mbar_init(acc_ready, 1);
        mbar_fence_init();
    }
    if (warp == 5) tmem_alloc<256>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    const int cblocks = a.Cin / KC;
    const int n_stages_total = a.ntaps * cblocks;
    // split-K: this CTA owns stages [s_begin, s_end)
    const int nsplit = gridDim.z, z = blockIdx.z;
    const int s_begin = (int)(((long long)n_stages_total * z) / nsplit), s_end = (int)(((long long)n_stages_total * (z + 1)) / nsplit);
    const int n_st = s_end - s_begin;
    const int m0 = blockIdx.x * TM, nt = blockIdx.y;
    const int HW = a.H * a.W, Mtot = a.B * HW;

    if (warp < 4) {
        // ======================= A producers: im2col gather, 16 x 16-byte cp.async per thread and stage =======================
        const int r = tid;                               // operand-tile row = pixel m0 + r
        const int p = m0 + r;
        const bool pvalid = p < Mtot;
        const int b = pvalid ? p / HW : 0, rem = pvalid ? p - b * HW : 0, y = rem / a.W, x = rem - y * a.W;
        uint8_t* arow = smem + S::A + (r >> 3) * 128 + (r & 7) * 16;
        for (int i = 0; i < n_st; i++) {
            const int s = s_begin + i, stg = i % NSTG;
            if (i >= NSTG) mbar_wait(empty + stg, ((i / NSTG) - 1) & 1);
            const int tap = s / cblocks, cb = s - tap * cblocks;
            const int dy = a.ntaps == 9 ? tap / 3 - 1 : 0, dx = a.ntaps == 9 ? tap % 3 - 1 : 0;
            const int yy = y + dy, xx = x + dx;
            const bool ok = pvalid && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
            const size_t off = ok ? ((size_t)(b * a.H + yy) * a.W + xx) * a.Cin + (size_t)cb * KC : 0;
            const __half* sh = a.xh + off;
            const __half* sl = a.xl + off;
            uint8_t* dst = arow + stg * A_STAGE;
            const uint32_t nb = ok ? 16u : 0u;           // 0 source bytes -> the 16 destination bytes are zero-filled (padding)
#pragma unroll
            for (int kc = 0; kc < KC / 8; kc++) {
                cp_async16_zfill(dst + kc * (TM * 16), sh + kc * 8, nb);
                cp_async16_zfill(dst + A_LO + kc * (TM * 16), sl + kc * 8, nb);
            }
            cp_async_commit();
            if (i >= LAG) {                              // stage i-LAG has landed: publish it to the tensor core (async proxy)
                cp_async_wait<LAG>();
                fence_proxy_async();
                mbar_arrive(full_a + (i - LAG) % NSTG);
            }
        }
        {   // drain: the last min(LAG, n_st) stages
            cp_async_wait<0>();
            fence_proxy_async();
            for (int i = (n_st > LAG ? n_st - LAG : 0); i < n_st; i++) mbar_arrive(full_a + i % NSTG);
        }
        // ======================= epilogue: TMEM -> (+ bias) -> NCHW fp32, 32 consecutive pixels per store instruction =======================
        mbar_wait(acc_ready, 0);
        tc_fence_after();
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        float* obase = a.out + (size_t)b * a.Cout * HW + rem;
        const float omul = a.out_mul * (a.out_mul_dev ? __ldg(a.out_mul_dev) : 1.0f);
#pragma unroll 1
        for (int cbk = 0; cbk < N_TILE / 32; cbk++) {
            float v[32];
            tmem_ld32(tmem + lane_base + cbk * 32, v);
            if (pvalid && n_st > 0) {
                const int n0 = nt * N_TILE + cbk * 32;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    v[j] *= omul;
                    if (a.bias && z == 0) v[j] += __ldg(a.bias + n0 + j);
                }
                if (a.out_rows) {
                    // token-major output: this thread owns 32 consecutive floats of its row
                    float* dst = a.out + (size_t)p * a.Cout + n0;
                    if (nsplit > 1) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        float* dst = obase + (size_t)(n0 + j) * HW;
                        if (nsplit > 1) atomicAdd(dst, v[j]);
                        else *dst = v[j];
                    }
                }
            }
        }
    } else if (warp == 4) {
        // ======================= B loader: one bulk copy per stage =======================
        if (lane == 0) {
            const uint8_t* wsrc = a.wpk + ((size_t)nt * n_stages_total + s_begin) * S::B_STAGE;
            for (int i = 0; i < n_st; i++) {
                const int stg = i % NSTG;
                if (i >= NSTG) mbar_wait(empty + stg, ((i / NSTG) - 1) & 1);
                mbar_arrive_expect_tx(full_b + stg, S::B_STAGE);
                bulk_g2s(smem + S::B + stg * S::B_STAGE, wsrc + (size_t)i * S::B_STAGE, S::B_STAGE, full_b + stg);
            }
        }
    } else {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16(N_TILE);
            const uint32_t a_base = smem_u32(smem + S::A), b_base = smem_u32(smem + S::B);
            for (int i = 0; i < n_st; i++) {
                const int stg = i % NSTG;
                mbar_wait(full_a + stg, (i / NSTG) & 1);
                mbar_wait(full_b + stg, (i / NSTG) & 1);
                tc_fence_after();
                const uint64_t a_hi0 = make_smem_desc(a_base + stg * A_STAGE, TM * 16, 128);
                const uint64_t a_lo0 = make_smem_desc(a_base + stg * A_STAGE + A_LO, TM * 16, 128);
                const uint64_t b_hi0 = make_smem_desc(b_base + stg * S::B_STAGE, 16u * N_TILE, 128);
#pragma unroll
                for (int k = 0; k < KC / 16; k++) {
                    const uint64_t a_hi = a_hi0 + (uint64_t)k * ((2 * TM * 16) >> 4), a_lo = a_lo0 + (uint64_t)k * ((2 * TM * 16) >> 4);
                    const uint64_t b_hi = b_hi0 + (uint64_t)k * ((64 * N_TILE) >> 4), b_lo = b_hi + ((32 * N_TILE) >> 4);
                    umma_f16(tmem, a_hi, b_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);
                    umma_f16(tmem, a_hi, b_lo, idesc, 1u);
                    umma_f16(tmem, a_lo, b_hi, idesc, 1u);
                }
                umma_commit(empty + stg);
            }
            umma_commit(acc_ready);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc<256>(tmem);
}

// ---- weight packing: fp32 [Cout, Cin, kh, kw] -> per (n-tile, tap, 64-channel block) stage blocks in UMMA canonical K-major order (hi | lo) ----
// transposed != 0 packs the input-gradient operator: rows = C_in of the forward layer, K = C_out, taps flipped (dx = conv(dy, W^T flipped)).
__global__ void pack_conv_kernel(const float* __restrict__ w, int Cout_w, int Cin_w, int ntaps, int n_tile, int transposed, float wscale,
                                 uint8_t* __restrict__ out) {
    const int N = transposed ? Cin_w : Cout_w;        // rows of the packed operator
    const int K = transposed ? Cout_w : Cin_w;        // channels contracted per tap
    const int cblocks = K / KC, n_tiles = N / n_tile;
    const size_t total = (size_t)N * K * ntaps;
    const size_t stage_bytes = (size_t)(KC / 16) * 64 * n_tile;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % K);
        const int tap = (int)((i / K) % ntaps);
        const int n = (int)(i / ((size_t)K * ntaps));
        float v;
        if (!transposed) v = w[((size_t)n * Cin_w + c) * ntaps + tap];
        else v = w[((size_t)c * Cin_w + n) * ntaps + (ntaps - 1 - tap)];
        v *= wscale;      // power of two chosen by the host so that max |w| ~ 2^8: hi AND lo parts stay normal fp16 numbers (full 22 bits)
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn(v - __half2float(h));
        const int nt = n / n_tile, nn = n - nt * n_tile;
        const int cb = c / KC, cc = c - cb * KC, slab = cc >> 4, kk = cc & 15, kcore = kk >> 3, ki = kk & 7;
        const size_t stage = (size_t)nt * (ntaps * cblocks) + (size_t)tap * cblocks + cb;
        const size_t off = stage * stage_bytes + (size_t)slab * (64 * n_tile) + (size_t)kcore * (16 * n_tile) + (nn >> 3) * 128 + (nn & 7) * 16 + ki * 2;
        *reinterpret_cast<__half*>(out + off) = h;
        *reinterpret_cast<__half*>(out + off + 32 * n_tile) = l;
    }
    (void)n_tiles;
}

// ---- activation split: fp32 NCHW -> fp16 (hi, lo) NHWC planes, optional SiLU (x * sigmoid(x)) applied first ----
__global__ void nchw_split_kernel(const float* __restrict__ x, int C, int HW, int act, const float* __restrict__ scale_dev, __half* __restrict__ hi,
                                  __half* __restrict__ lo) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 256 threads: 8 rows per pass
    // dynamic power-of-two scale (gradients of the VAE backward are ~1e-6: below fp16's normal range without it)
    const float sc = scale_dev ? __ldg(scale_dev) : 1.0f;
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, p = p0 + tx;
        float v = 0.f;
        if (c < C && p < HW) {
            v = x[((size_t)b * C + c) * HW + p];
            if (act == 1) v = v / (1.0f + expf(-v));
            v *= sc;
        }
        tile[j][tx] = v;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int p = p0 + j, c = c0 + tx;
        if (p < HW && c < C) {
            const float v = tile[tx][j];
            const __half h = __float2half_rn(v);
            const size_t o = ((size_t)b * HW + p) * C + c;
            hi[o] = h;
            lo[o] = __float2half_rn(v - __half2float(h));
        }
    }
}

// ---- row-major activations [rows, C] fp32 (tokens of a linear layer) -> fp16 (hi, lo) planes of the same layout ----
__global__ void rows_split_kernel(const float* __restrict__ x, size_t n, const float* __restrict__ scale_dev, __half* __restrict__ hi, __half* __restrict__ lo) {
    const float sc = scale_dev ? __ldg(scale_dev) : 1.0f;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += (size_t)gridDim.x * blockDim.x * 2) {
        const float2 v = *reinterpret_cast<const float2*>(x + i);
        const __half2 h = __floats2half2_rn(v.x * sc, v.y * sc);
        const float2 hf = __half22float2(h);
        *reinterpret_cast<__half2*>(hi + i) = h;
        *reinterpret_cast<__half2*>(lo + i) = __floats2half2_rn(v.x * sc - hf.x, v.y * sc - hf.y);
    }
}

template <int N_TILE>
int launch(const Args& a, int nsplit, cudaStream_t stream) {
    using S = Smem<N_TILE>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL);
        if (e != cudaSuccess) { set_error("conv_tc: cannot reserve %d B smem: %s", (int)S::TOTAL, cudaGetErrorString(e)); return MB_ECUDA; }
        attr_set = true;
    }
    const int Mtot = a.B * a.H * a.W;
    dim3 grid((Mtot + TM - 1) / TM, a.Cout / N_TILE, nsplit);
    conv_tc_kernel<N_TILE><<<grid, NTHREADS, S::TOTAL, stream>>>(a);
    return check_launch("conv_tc");
}

}  // namespace conv
}  // namespace mb

extern "C" int mb_conv_pack_weights(const float* w, int Cout, int Cin, int ntaps, int n_tile, int transposed, float wscale, void* out, mb_stream_t stream) {
    using namespace mb;
    const int N = transposed ? Cin : Cout, K = transposed ? Cout : Cin;
    if (!w || !out || (ntaps != 9 && ntaps != 1) || (n_tile != 128 && n_tile != 160) || N % n_tile || K % conv::KC) {
        set_error("conv_pack_weights: need ntaps in {1, 9}, n_tile in {128, 160}, rows %% n_tile == 0, channels %% 64 == 0");
        return MB_EINVAL;
    }
    conv::pack_conv_kernel<<<1024, 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, ntaps, n_tile, transposed, wscale, (uint8_t*)out);
    return check_launch("conv_pack_weights");
}

extern "C" int mb_nchw_split(const float* x, int B, int C, int HW, int act, const float* scale_dev, void* hi, void* lo, mb_stream_t stream) {
    using namespace mb;
    if (!x || !hi || !lo || B <= 0 || C <= 0 || HW <= 0) { set_error("nchw_split: bad argument"); return MB_EINVAL; }
    dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
    conv::nchw_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, C, HW, act, scale_dev, (__half*)hi, (__half*)lo);
    return check_launch("nchw_split");
}

extern "C" int mb_rows_split(const float* x, uint64_t n_elems, const float* scale_dev, void* hi, void* lo, mb_stream_t stream) {
    using namespace mb;
    if (!x || !hi || !lo || (n_elems & 1)) { set_error("rows_split: bad argument (even element count)"); return MB_EINVAL; }
    if (n_elems == 0) return MB_OK;
    uint64_t nb = (n_elems / 2 + 255) / 256;
    if (nb > 148 * 16) nb = 148 * 16;
    const unsigned blocks = (unsigned)nb;
    conv::rows_split_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (size_t)n_elems, scale_dev, (__half*)hi, (__half*)lo);
    return check_launch("rows_split");
}

extern "C" int mb_conv_tc(const void* x_hi, const void* x_lo, const void* w_packed, const float* bias, float* out, int B, int H, int W, int Cin, int Cout,
                          int ntaps, int n_tile, int nsplit, float out_mul, const float* out_mul_dev, int out_rows, mb_stream_t stream) {
    using namespace mb;
    if (!x_hi || !x_lo || !w_packed || !out) { set_error("conv_tc: null pointer"); return MB_EINVAL; }
    if ((ntaps != 9 && ntaps != 1) || (n_tile != 128 && n_tile != 160) || Cout % n_tile || Cin % conv::KC || B <= 0 || H <= 0 || W <= 0) {
        set_error("conv_tc: need ntaps in {1, 9}, n_tile in {128, 160}, Cout %% n_tile == 0, Cin %% 64 == 0");
        return MB_EINVAL;
    }
    const int n_stages = ntaps * (Cin / conv::KC);
    if (nsplit < 1) nsplit = 1;
    if (nsplit > n_stages) nsplit = n_stages;
    conv::Args a{(const __half*)x_hi, (const __half*)x_lo, (const uint8_t*)w_packed, bias, out, B, H, W, Cin, Cout, ntaps, out_mul, out_mul_dev, out_rows};
    return n_tile == 128 ? conv::launch<128>(a, nsplit, (cudaStream_t)stream) : conv::launch<160>(a, nsplit, (cudaStream_t)stream);
}
