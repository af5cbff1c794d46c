// Stem layers 3 and 4 (conv3x3 stride 2 pad 1 + folded BN [+ Hardswish], lib/models/vit_dist/vit_dist.py:36-54)
// on the tcgen05 tensor cores, for the search branch (S = 256).
//
// A stride-2 convolution reads in[2oy+ky-1][2ox+kx-1].  The fp32 NCHW input of a band of 8 output rows is
// split by (row parity, column parity) into four planes P[rp][cp][r][c] = in[2r+rp][2c+cp], stored in shared
// memory as fp16 hi/lo in 8-channel chunks [chunk][row][c][8] with pitch exactly W_out, so that the 128
// output pixels of an M tile are 128 consecutive 16-byte rows of any plane: a no-swizzle K-major UMMA
// descriptor addresses tap (ky, kx) by picking the plane (rp = ky odd ? 0 : 1, cp = kx odd ? 0 : 1) and a start
// row.  The kx = 0 taps need column index ox-1: they accumulate UNSHIFTED into a second TMEM accumulator T_B
// that the epilogue shifts by one TMEM lane with a warp shuffle (a warp's 32 lanes are whole image rows; the
// value entering at ox = 0 is the zero padding).  Products are hi*hi + lo*hi + hi*lo, fp32 accumulation.
#include "vt_internal.h"
#include "vt_tc.cuh"

namespace vt {

using namespace tc;

namespace {

template <int CIN, int CCH, int COUT, int NPAD, int WOUT>
struct TcConv {
    static constexpr int kBR = 8;                               // output rows per CTA
    static constexpr int kRowsPerTile = 128 / WOUT;
    static constexpr int kTiles = kBR / kRowsPerTile;            // M tiles per CTA
    static constexpr int kPlaneRows = kBR + 1;                   // local row 0 <-> parity-plane row oy0 - 1
    static constexpr int kChunkBytes = kPlaneRows * WOUT * 16;
    static constexpr int kPlaneBytes = CCH * kChunkBytes;
    static constexpr int kABytes = 4 * kPlaneBytes;              // one precision
    static constexpr int kKA = 6 * CCH * 8, kKB = 3 * CCH * 8;   // K of the two accumulators
    static constexpr int kWPrecBytes = (kKA + kKB) * NPAD * 2;   // one precision: T_A chunks then T_B chunks
    static constexpr int kWBytes = 2 * kWPrecBytes;
    static constexpr int kOffA = 0;
    static constexpr int kOffW = 2 * kABytes;
    static constexpr int kOffBias = kOffW + kWBytes;
    static constexpr int kOffBar = kOffBias + NPAD * 4;
    static constexpr int kSmemBytes = kOffBar + 32;
    static constexpr int kTmemCols = (kTiles * 2 * NPAD <= 32) ? 32 : (kTiles * 2 * NPAD <= 64) ? 64 : (kTiles * 2 * NPAD <= 128) ? 128 : 256;
    static constexpr int kThreads = 256;
    static_assert(CCH % 2 == 0 && CCH * 8 >= CIN && NPAD % 16 == 0 && NPAD >= COUT && 128 % WOUT == 0 && kBR % kRowsPerTile == 0, "shape");
    static_assert(kOffBar % 8 == 0, "barrier alignment");
};

}  // namespace

// grid: (Hout / 8 bands, n tracks).  in: [n][CIN][2*Hout][2*WOUT] fp32; wt: packed fp16 hi|lo weight blob; bias fp32 [NPAD].
template <int CIN, int CCH, int COUT, int NPAD, int WOUT, bool HSWISH, bool TOKENS>
__global__ void __launch_bounds__(256)
conv_s2_tc_kernel(const float* __restrict__ in, const uint8_t* __restrict__ wt, const float* __restrict__ bias,
                  float* __restrict__ out, const float* __restrict__ pos, int tok_stride_rows, int tok_off) {
    using K = TcConv<CIN, CCH, COUT, NPAD, WOUT>;
    constexpr int Hout = WOUT, Hin = 2 * WOUT, Win = 2 * WOUT;
    extern __shared__ __align__(128) uint8_t smem_tc[];
    uint8_t* sA = smem_tc + K::kOffA;
    uint8_t* sW = smem_tc + K::kOffW;
    float* sBias = reinterpret_cast<float*>(smem_tc + K::kOffBias);
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem_tc + K::kOffBar);
    uint64_t* bar_d = bar_w + 1;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_w + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int oy0 = blockIdx.x * K::kBR;
    const int b = blockIdx.y;

    if (warp == 0) tmem_alloc(s_tmem, K::kTmemCols);
    if (tid == 32) { mbar_init(bar_w, 1); mbar_init(bar_d, 1); mbar_fence_init(); }
    if (tid < NPAD) sBias[tid] = tid < COUT ? __ldg(bias + tid) : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) bulk_g2s_elect(sW, wt, K::kWBytes, bar_w);

    // ---- stage the band's input as parity planes, fp16 hi / lo ----------------------------------------------
    {
        const float* inb = in + (size_t)b * CIN * Hin * Win;
        constexpr int kQuads = Win / 4;
        constexpr int kItems = K::kPlaneRows * 2 * CCH * kQuads;            // (lr, rp, chunk, quad of 4 input columns)
        for (int i = tid; i < kItems; i += K::kThreads) {
            const int quad = i % kQuads;
            const int chunk = (i / kQuads) % CCH;
            const int rp = (i / (kQuads * CCH)) & 1;
            const int lr = i / (kQuads * CCH * 2);
            const int gy = 2 * (oy0 - 1 + lr) + rp;
            float v[8][4];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int ci = chunk * 8 + c;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ci < CIN && gy >= 0 && gy < Hin) t = __ldg(reinterpret_cast<const float4*>(inb + ((size_t)ci * Hin + gy) * Win + 4 * quad));
                v[c][0] = t.x; v[c][1] = t.y; v[c][2] = t.z; v[c][3] = t.w;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {                                   // input column 4*quad + e -> parity e & 1, index 2*quad + e/2
                const int cp = e & 1, c = 2 * quad + (e >> 1);
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) split_pack2(v[2 * j][e], v[2 * j + 1][e], hi[j], lo[j]);
                const int off = (rp * 2 + cp) * K::kPlaneBytes + chunk * K::kChunkBytes + (lr * WOUT + c) * 16;
                *reinterpret_cast<uint4*>(sA + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(sA + K::kABytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        fence_async_smem();
    }
    __syncthreads();

    const uint32_t tbase = __shfl_sync(0xffffffffu, *s_tmem, 0);
    if (warp == 0) {                 // convergent issue, elected lane performs
        const uint32_t sbase = smem_u32(smem_tc);
        const uint32_t idesc = instr_desc_f16(128, NPAD, false);
        mbar_wait(bar_w, 0);
        tc_fence_after();
#pragma unroll
        for (int tile = 0; tile < K::kTiles; ++tile) {
#pragma unroll
            for (int acc = 0; acc < 2; ++acc) {                             // 0: T_A (kx = 1, 2)   1: T_B (kx = 0, shifted later)
                const uint32_t d = tbase + (tile * 2 + acc) * NPAD;
                const int ntaps = acc == 0 ? 6 : 3;
                bool first = true;
#pragma unroll
                for (int tap = 0; tap < ntaps; ++tap) {
                    const int ky = acc == 0 ? tap / 2 : tap;
                    const int kx = acc == 0 ? 1 + (tap & 1) : 0;
                    const int rp = (ky == 1) ? 0 : 1, cp = (kx == 1) ? 0 : 1;
                    const int row0 = tile * K::kRowsPerTile + (ky == 0 ? 0 : 1);
#pragma unroll
                    for (int kp = 0; kp < CCH / 2; ++kp) {
                        const uint32_t aoff = (rp * 2 + cp) * K::kPlaneBytes + (2 * kp) * K::kChunkBytes + row0 * WOUT * 16;
                        const uint64_t ah = smem_desc(sbase + K::kOffA + aoff, K::kChunkBytes, 128);
                        const uint64_t al = smem_desc(sbase + K::kOffA + K::kABytes + aoff, K::kChunkBytes, 128);
                        const uint32_t boff = ((acc == 0 ? 0 : K::kKA / 8) + tap * CCH + 2 * kp) * (NPAD * 16);
                        const uint64_t bh = smem_desc(sbase + K::kOffW + boff, NPAD * 16, 128);
                        const uint64_t bl = smem_desc(sbase + K::kOffW + K::kWPrecBytes + boff, NPAD * 16, 128);
                        mma_ss_elect(d, ah, bh, idesc, first ? 0u : 1u);
                        mma_ss_elect(d, al, bh, idesc, 1u);
                        mma_ss_elect(d, ah, bl, idesc, 1u);
                        first = false;
                    }
                }
            }
        }
        mma_commit_elect(bar_d);
    }
    mbar_wait(bar_d, 0);
    tc_fence_after();

    // ---- epilogue: thread -> (tile, pixel) [and a channel half when the CTA has one tile] -----------------------
    {
        constexpr int kChPerThread = (K::kTiles == 2) ? NPAD : NPAD / 2;
        const int tile = (K::kTiles == 2) ? (warp >> 2) : 0;
        const int ch0 = (K::kTiles == 2) ? 0 : (warp >> 2) * kChPerThread;
        const int r = 32 * (warp & 3) + lane;                               // row of the M tile = TMEM lane
        const int oy = oy0 + tile * K::kRowsPerTile + r / WOUT, ox = r % WOUT;
        const uint32_t ta = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + tile * 2 * NPAD + ch0;
        static_assert(kChPerThread % 8 == 0, "channel split");
#pragma unroll 1
        for (int c0 = 0; c0 < kChPerThread; c0 += 8) {
            uint32_t ra[8], rb[8];
            tmem_ld8(ta + c0, ra);
            tmem_ld8(ta + NPAD + c0, rb);
            tc_wait_ld();
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float tb = __shfl_up_sync(0xffffffffu, __uint_as_float(rb[j]), 1);     // T_B[oy][ox-1]
                if (ox == 0) tb = 0.f;
                float t = __uint_as_float(ra[j]) + tb + sBias[ch0 + c0 + j];
                if (HSWISH) t = t * fminf(fmaxf(t + 3.f, 0.f), 6.f) / 6.f;
                v[j] = t;
            }
            if (TOKENS) {
                const int tok = oy * WOUT + ox;
                float* o = out + ((size_t)b * tok_stride_rows + tok_off + tok) * COUT + ch0 + c0;
                const float* pe = pos + (size_t)tok * COUT + ch0 + c0;
                if (ch0 + c0 < COUT) {
                    const float4 e0 = __ldg(reinterpret_cast<const float4*>(pe)), e1 = __ldg(reinterpret_cast<const float4*>(pe + 4));
                    *reinterpret_cast<float4*>(o) = make_float4(v[0] + e0.x, v[1] + e0.y, v[2] + e0.z, v[3] + e0.w);
                    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4] + e1.x, v[5] + e1.y, v[6] + e1.z, v[7] + e1.w);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (ch0 + c0 + j < COUT) out[(((size_t)b * COUT + ch0 + c0 + j) * Hout + oy) * WOUT + ox] = v[j];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, K::kTmemCols);
}

template <int CIN, int CCH, int COUT, int NPAD, int WOUT, bool HSWISH, bool TOKENS>
static int run_tc_conv(const float* in, int n, const uint8_t* wt, const float* bias, float* out, const float* pos,
                       int tok_stride_rows, int tok_off, cudaStream_t st) {
    using K = TcConv<CIN, CCH, COUT, NPAD, WOUT>;
    auto kern = conv_s2_tc_kernel<CIN, CCH, COUT, NPAD, WOUT, HSWISH, TOKENS>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::kSmemBytes) != cudaSuccess) return -1;
        configured = true;
    }
    int launched = 0;
    for (int first = 0; first < n; first += 32768) {
        const int m = n - first < 32768 ? n - first : 32768;
        kern<<<dim3(WOUT / K::kBR, m), 256, K::kSmemBytes, st>>>(in + (size_t)first * CIN * 4 * WOUT * WOUT, wt, bias,
                                                                 TOKENS ? out + (size_t)first * tok_stride_rows * COUT
                                                                        : out + (size_t)first * COUT * WOUT * WOUT,
                                                                 pos, tok_stride_rows, tok_off);
        ++launched;
    }
    return cudaGetLastError() == cudaSuccess ? launched : -1;
}

// conv3 (12 -> 24, 64x64 -> 32x32, Hardswish) and conv4 (24 -> 48, 32x32 -> 16x16, tokens + pos-embed) of the search branch
int launch_stem34_tc(const float* a2, int n, const ModelW& w, float* a3, float* tokens, int tok_stride_rows, int tok_off,
                     cudaStream_t st) {
    int total = 0, r;
    if ((r = run_tc_conv<12, 2, 24, 32, 32, true, false>(a2, n, w.stem_tc_w[0], w.stem_tc_b[0], a3, nullptr, 0, 0, st)) < 0) return r;
    total += r;
    if ((r = run_tc_conv<24, 4, 48, 48, 16, false, true>(a3, n, w.stem_tc_w[1], w.stem_tc_b[1], tokens, w.pos_x, tok_stride_rows, tok_off, st)) < 0) return r;
    total += r;
    return total;
}

size_t stem_tc_weight_bytes(int layer) {      // layer 0: conv3, 1: conv4
    return layer == 0 ? (size_t)TcConv<12, 2, 24, 32, 32>::kWBytes : (size_t)TcConv<24, 4, 48, 48, 16>::kWBytes;
}

}  // namespace vt
