// Micro-benchmark (development aid): throughput of the fused stem's per-pixel bilinear arithmetic in four arrangements, on register data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_arith gather_arith.cu && ./gather_arith
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned mad_hi(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
constexpr unsigned kC = 2u + (0x6400u << 2);

template <int V>
__device__ __forceinline__ void px3(const uint32_t (&wd)[6], unsigned sh0, unsigned sh1, unsigned wx, unsigned bz, int (&px)[3]) {
    const uint32_t u0 = __funnelshift_r(wd[0], wd[1], sh0), u1 = __funnelshift_r(wd[1], wd[2], sh0);
    const uint32_t t0 = __funnelshift_r(wd[3], wd[4], sh1), t1 = __funnelshift_r(wd[4], wd[5], sh1);
    if (V == 0) {
        const int b0 = bz & 0xffff, b1 = bz >> 16;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const unsigned sel = ch == 0 ? 0x0030u : ch == 1 ? 0x0041u : 0x0052u;
            const int h0 = (int)__dp2a_lo(wx, __byte_perm(u0, u1, sel), 0u);
            const int h1 = (int)__dp2a_lo(wx, __byte_perm(t0, t1, sel), 0u);
            px[ch] = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + (int)kC) >> 2;
        }
        return;
    }
    const uint32_t urg = __byte_perm(u0, u1, 0x4130), ubb = __byte_perm(u0, u1, 0x0052);
    const uint32_t trg = __byte_perm(t0, t1, 0x4130), tbb = __byte_perm(t0, t1, 0x0052);
    const unsigned h0[3] = {__dp2a_lo(wx, urg, 0u), __dp2a_hi(wx, urg, 0u), __dp2a_lo(wx, ubb, 0u)};
    const unsigned h1[3] = {__dp2a_lo(wx, trg, 0u), __dp2a_hi(wx, trg, 0u), __dp2a_lo(wx, tbb, 0u)};
    if (V == 1) {
        const unsigned b0s = (bz & 0xffffu) << 12, b1s = (bz >> 16) << 12;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) px[ch] = (int)(mad_hi(h1[ch] & ~15u, b1s, mad_hi(h0[ch] & ~15u, b0s, kC)) >> 2);
    } else if (V == 2) {
        const unsigned b0s = (bz & 0xffffu) << 12, b1s = (bz >> 16) << 12;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) px[ch] = (int)((__umulhi(h0[ch] & ~15u, b0s) + __umulhi(h1[ch] & ~15u, b1s) + kC) >> 2);
    } else if (V == 3) {
        const unsigned b0s = (bz & 0xffffu) << 16, b1s = (bz >> 16) << 16;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
            px[ch] = (int)((__umulhi(__umulhi(h0[ch], 1u << 28), b0s) + __umulhi(__umulhi(h1[ch], 1u << 28), b1s) + kC) >> 2);
    } else {            // 4: the original multiplies, two permutes
        const int b0 = bz & 0xffff, b1 = bz >> 16;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) px[ch] = (((b0 * ((int)h0[ch] >> 4)) >> 16) + ((b1 * ((int)h1[ch] >> 4)) >> 16) + (int)kC) >> 2;
    }
}

template <int V>
__global__ void __launch_bounds__(512) k(const uint32_t* in, uint32_t* out, int iters) {
    uint32_t wd[6];
    for (int i = 0; i < 6; ++i) wd[i] = in[threadIdx.x * 6 + i];
    unsigned sh0 = in[3000 + threadIdx.x] & 24, sh1 = in[3600 + threadIdx.x] & 24;
    unsigned wx = (in[4200 + threadIdx.x] & 2047) | ((2048 - (in[4200 + threadIdx.x] & 2047)) << 16);
    unsigned bz = (in[4800 + threadIdx.x] & 2047) | ((2048 - (in[4800 + threadIdx.x] & 2047)) << 16);
    unsigned acc = 0;
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        int px[3];
        px3<V>(wd, sh0, sh1, wx, bz, px);
        acc += px[0] ^ px[1] ^ px[2];
        wd[0] += acc; wd[3] ^= acc;              // data dependence between iterations (2 extra ALU ops, the same in every variant)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int V>
float run(const uint32_t* in, uint32_t* out, int iters, unsigned* check) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<V><<<148 * 2, 512>>>(in, out, 16);
    cudaEventRecord(a);
    k<V><<<148 * 2, 512>>>(in, out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    cudaMemcpy(check, out, 4, cudaMemcpyDeviceToHost);
    return ms;
}

int main() {
    uint32_t *in, *out;
    cudaMalloc(&in, 8192 * 4); cudaMalloc(&out, 148 * 2 * 512 * 4);
    uint32_t h[8192];
    uint32_t s = 12345;
    for (int i = 0; i < 8192; ++i) { s = s * 1664525u + 1013904223u; h[i] = s; }
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const int iters = 4096;
    unsigned c[5];
    float t[5] = {run<0>(in, out, iters, c + 0), run<1>(in, out, iters, c + 1), run<2>(in, out, iters, c + 2), run<3>(in, out, iters, c + 3), run<4>(in, out, iters, c + 4)};
    const double px = 148.0 * 2 * 512 * iters;
    const char* names[5] = {"original (5 SHF + 2 IMAD + IADD3, 3 PRMT)", "LOP + IMAD.HI chained through the addend", "LOP + IMAD.HI + IADD3", "IMAD.HI for the shifts too", "original multiplies, 2 PRMT"};
    for (int v = 0; v < 5; ++v)
        printf("variant %d  %-48s %.3f ms  %.1f cycles/pixel/SMSP-lane-group @1.965GHz  check %08x\n", v, names[v], t[v],
               t[v] * 1e-3 * 1.965e9 / (px / (148.0 * 4 * 32)), c[v]);
    return cudaGetLastError() != cudaSuccess;
}
