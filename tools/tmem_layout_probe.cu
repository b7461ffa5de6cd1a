// Probe of the tcgen05.ld.16x256b register mapping: fill TMEM with value = lane*1000 + column (tcgen05.st 32x32b, where
// the mapping is trivially thread = lane), read back with 16x256b.x2 at lane offsets 0 and 16, print what thread t got.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I xna_basecaller_b200/csrc tools/tmem_layout_probe.cu -o tools/tmem_probe.bin
#include <cstdio>
#include <cuda_runtime.h>
#include "xb_ptx.cuh"
using namespace xbptx;

__global__ void __launch_bounds__(128, 1) probe(uint32_t *out) {
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tmem_alloc(&holder, 32); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = holder;
    uint32_t v[32];
    for (int c = 0; c < 32; c++) v[c] = (warp * 32 + lane) * 1000 + c;
    tmem_st_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16), v);
    tmem_st_wait();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 1) {        // lanes 32..63
        uint32_t r[8], s[8];
        const uint32_t base = tm + ((uint32_t)32 << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(base));
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7])
                     : "r"(base + ((uint32_t)16 << 16)));
        tmem_ld_wait();
        for (int i = 0; i < 8; i++) { out[lane * 16 + i] = r[i]; out[lane * 16 + 8 + i] = s[i]; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 32); }
}

int main() {
    uint32_t *d, h[512];
    cudaMalloc(&d, sizeof h);
    probe<<<1, 128>>>(d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    for (int t = 0; t < 32; t++) {
        printf("thread %2d:", t);
        for (int i = 0; i < 16; i++) printf(" (%d,%d)", h[t * 16 + i] / 1000 - 32, h[t * 16 + i] % 1000);
        printf("\n");
    }
    return 0;
}
