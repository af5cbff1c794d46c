// Development aid: exhaustive check (all 2^32 bit patterns) that  q = p * r;  e = fma(-6, q, p);  q' = fma(e, r, q)  with r = fl(1/6)
// equals the IEEE-754 correctly rounded p / 6.f.  Prints the number of mismatches and the magnitude range in which they occur.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ unsigned long long g_bad;
__device__ unsigned int g_min_bad = 0x7f800000u, g_max_bad = 0u;
__global__ void check() {
    const float r = 1.f / 6.f;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32); i += (unsigned long long)gridDim.x * blockDim.x) {
        const float p = __uint_as_float((unsigned)i);
        if (!(fabsf(p) <= 3.4e38f)) continue;                       // NaN / inf
        const float want = __fdiv_rn(p, 6.f);
        const float q = __fmul_rn(p, r);
        const float e = __fmaf_rn(-6.f, q, p);
        const float got = __fmaf_rn(e, r, q);
        if (__float_as_uint(want) != __float_as_uint(got)) {
            atomicAdd(&g_bad, 1ull);
            const unsigned m = (unsigned)i & 0x7fffffffu;
            atomicMin(&g_min_bad, m); atomicMax(&g_max_bad, m);
        }
    }
}
int main() {
    check<<<148 * 16, 256>>>();
    cudaDeviceSynchronize();
    unsigned long long bad; unsigned lo, hi;
    cudaMemcpyFromSymbol(&bad, g_bad, 8); cudaMemcpyFromSymbol(&lo, g_min_bad, 4); cudaMemcpyFromSymbol(&hi, g_max_bad, 4);
    float flo, fhi; memcpy(&flo, &lo, 4); memcpy(&fhi, &hi, 4);
    printf("mismatches %llu  |p| range of mismatches [%g, %g] (bits %08x..%08x)  err %s\n", bad, flo, fhi, lo, hi, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
