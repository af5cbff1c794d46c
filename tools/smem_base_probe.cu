// Development aid: where does the dynamic shared-memory window of a kernel start (shared-space address), and where does a 512-column TMEM allocation?
#include <cstdio>
#include <cstdint>
extern __shared__ __align__(128) uint8_t sm[];
__global__ void probe(uint32_t* out) {
    __shared__ uint32_t s_t;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_t)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) { out[0] = (uint32_t)__cvta_generic_to_shared(sm); out[1] = s_t; }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(s_t) : "memory");
}
__global__ void probe2(uint32_t* out) { if (threadIdx.x == 0) out[0] = (uint32_t)__cvta_generic_to_shared(sm); }
int main() {
    uint32_t* d; cudaMalloc(&d, 16); uint32_t h[4] = {0, 0, 0, 0};
    cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    probe<<<1, 128, 4096>>>(d); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("static 4 B + dynamic: smem base 0x%x tmem base 0x%x (%s)\n", h[0], h[1], cudaGetErrorString(cudaGetLastError()));
    probe2<<<1, 128, 200 * 1024>>>(d); cudaMemcpy(h, d, 4, cudaMemcpyDeviceToHost);
    printf("no static, 200 KB dynamic: smem base 0x%x (%s)\n", h[0], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
