// Standalone GPU probe: pins down the tcgen05 operand layouts the block kernel relies on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I vittracker_b200/csrc tools/tc_probe.cu -o /tmp/tc_probe && /tmp/tc_probe
// For each case it computes D = A (128 x K) * B^T (N x K) with fp16 inputs / fp32 accumulate through one
// chain of tcgen05.mma and compares with a host double-precision product.
//   mode 0: A from shared memory (K-major, no swizzle), B from shared memory K-major
//   mode 1: A from TMEM (written with tcgen05.st, element 2c in the low half of column c), B K-major
//   mode 2: A from TMEM, B MN-major ([k][n] storage, n contiguous)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "vt_tc.cuh"

using namespace vt::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// A: [128][K] fp16 row-major in global; B: [N][K] fp16 row-major (logical); D: [128][N] fp32
__global__ void __launch_bounds__(128) probe_kernel(const __half* A, const __half* B, float* D, int N, int K, int mode, int swap_halves) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t* sA = smem;                                  // K-major: [k/8][row 0..127][8 elems] -> LBO = 128*16, SBO = 128
    uint8_t* sB = smem + 128 * K * 2;                    // K-major: [k/8][n][8]  |  MN-major: [n/8][k/8][k%8][n%8] (SBO = (K/8)*128, LBO = 128)
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 0) { mbar_init(&s_bar, 1); mbar_fence_init(); }
    // stage A (only used in mode 0) and B
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<__half*>(sA + (k / 8) * (128 * 16) + r * 16 + (k % 8) * 2) = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        if (mode == 2) *reinterpret_cast<__half*>(sB + (n / 8) * ((K / 8) * 128) + (k / 8) * 128 + (k % 8) * 16 + (n % 8) * 2) = B[i];
        else *reinterpret_cast<__half*>(sB + (k / 8) * (N * 16) + n * 16 + (k % 8) * 2) = B[i];
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = s_tmem;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t a_col = 256;                          // TMEM columns [256, 256 + K/2) hold A in modes 1, 2
    if (mode != 0) {
        // thread = row: pack K fp16 values into K/2 columns
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t r[8];
            for (int j = 0; j < 8; ++j) {
                const __half e0 = A[tid * K + 2 * (c0 + j)], e1 = A[tid * K + 2 * (c0 + j) + 1];
                const uint16_t u0 = *reinterpret_cast<const uint16_t*>(&e0), u1 = *reinterpret_cast<const uint16_t*>(&e1);
                r[j] = swap_halves ? ((uint32_t)u0 << 16 | u1) : ((uint32_t)u1 << 16 | u0);
            }
            tmem_st8(tbase + lane_base + a_col + c0, r);
        }
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = instr_desc_f16(128, N, mode == 2);
        for (int ks = 0; ks < K / 16; ++ks) {
            uint64_t bd;
            if (mode == 2) bd = smem_desc(smem_u32(sB) + ks * 2 * 128, 128, (K / 8) * 128);
            else bd = smem_desc(smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
            if (mode == 0) {
                const uint64_t ad = smem_desc(smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
                mma_ss(tbase, ad, bd, idesc, ks > 0);
            } else {
                mma_ts(tbase, tbase + a_col + ks * 8, bd, idesc, ks > 0);
            }
        }
        mma_commit(&s_bar);
    }
    mbar_wait(&s_bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tbase + lane_base + c0, r);
        tc_wait_ld();
        for (int j = 0; j < 16; ++j) D[tid * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

static double run_case(int N, int K, int mode, int swap_halves) {
    std::vector<__half> hA(128 * K), hB(N * K);
    std::vector<float> fA(128 * K), fB(N * K);
    srand(1234 + N * 7 + K * 13 + mode);
    for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 2001 - 1000) / 500.0f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 2001 - 1000) / 500.0f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
    __half *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, 128 * N * 4));
    const size_t smem = 128 * K * 2 + N * K * 2 + 256;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, N, K, mode, swap_halves);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); exit(2); }
    std::vector<float> hD(128 * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += (double)fA[r * K + k] * (double)fB[n * K + k];
            double d = fabs(acc - (double)hD[r * N + n]);
            if (!(d <= maxerr)) maxerr = d;            // NaN-propagating
        }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return maxerr;
}

int main() {
    const int cases[][2] = {{48, 48}, {144, 48}, {160, 48}, {192, 48}, {48, 192}, {48, 320}, {256, 16}, {16, 16}};
    for (auto& c : cases) {
        printf("N=%3d K=%3d | SS K-major: %.3e | TS (lo half = even k): %.3e | TS swapped halves: %.3e | TS + B MN-major: %.3e\n", c[0], c[1],
               run_case(c[0], c[1], 0, 0), run_case(c[0], c[1], 1, 0), run_case(c[0], c[1], 1, 1), run_case(c[0], c[1], 2, 0));
        fflush(stdout);
    }
    return 0;
}
