// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) as a function of N and of where A lives
// (shared memory descriptor vs tensor memory), issued back to back by one thread.  Operand contents are junk.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I xna_basecaller_b200/csrc tools/mma_microbench.cu -o /tmp/mma_mb
#include <cstdio>
#include <cuda_runtime.h>
#include "xb_ptx.cuh"
using namespace xbptx;

template <int N, bool A_TMEM, bool SAME_A>
__global__ void __launch_bounds__(128, 1) bench(long long *out, int iters) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((uint32_t *)smem)[i] = 0x3c003c00u;   // fp16 ones
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 1) { tmem_alloc(&holder, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = holder;
    if (warp == 0 && elect_one()) {
        constexpr uint32_t idesc = umma_idesc_f16(0, 128, N);
        const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 16384;
        long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int kk = SAME_A ? 0 : ((it & 7) * 4 + k);
                if (A_TMEM) mma_f16_ts(tm + 384, tm + kk * 8, umma_desc_sw128(b_addr + k * 32), idesc, 1);
                else mma_f16_ss(tm + 384, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, 1);
            }
        }
        long long t1 = clock64();
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tm, 512);
}

template <int N, bool A_TMEM, bool SAME_A> void run(const char *name) {
    long long *d; cudaMalloc(&d, 16);
    auto k = bench<N, A_TMEM, SAME_A>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 2000;
    for (int rep = 0; rep < 2; rep++) k<<<1, 128, 64 * 1024>>>(d, iters);
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-28s N=%3d  issue %.1f cyc/mma   complete %.1f cyc/mma   (%s)\n", name, N, h[0] / (4.0 * iters), h[1] / (4.0 * iters),
           cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<32, true, false>("A in TMEM");  run<96, true, false>("A in TMEM");  run<128, true, false>("A in TMEM");  run<256, true, false>("A in TMEM");
    run<32, true, true>("A in TMEM (same A cols)");  run<96, true, true>("A in TMEM (same A cols)");
    run<32, false, false>("A in SMEM"); run<96, false, false>("A in SMEM"); run<128, false, false>("A in SMEM"); run<256, false, false>("A in SMEM");
    return 0;
}
