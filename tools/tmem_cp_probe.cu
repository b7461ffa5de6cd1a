// Probe of tcgen05.cp.128x256b: what lands where in tensor memory when the source is one K = 16 slice of a
// [128 rows x 64 fp16] K-major tile in the 128B-swizzled layout TMA produces (16-byte chunk index XOR row % 8)?
// Element (row, k) holds the fp16 pair code (row * 64 + k) as raw 16-bit integers; after the copy every thread reads its
// lane's 32 columns back with tcgen05.ld.32x32b and the host prints (row, k) per (lane, column, half).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I xna_basecaller_b200/csrc tools/tmem_cp_probe.cu -o tools/tmem_cp_probe.bin
#include <cstdio>
#include <cuda_runtime.h>
#include "xb_ptx.cuh"
using namespace xbptx;

__global__ void __launch_bounds__(128, 1) probe(uint32_t *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t holder;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // [128 rows][128 B], 16-byte chunk c of row r stored at chunk position c ^ (r % 8)
    for (int i = threadIdx.x; i < 128 * 64; i += 128) {
        const int r = i / 64, k = i % 64, chunk = k / 8, pos = chunk ^ (r & 7);
        ((uint16_t *)smem)[r * 64 + pos * 8 + (k % 8)] = (uint16_t)(r * 64 + k);
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&holder, 32); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = holder;
    if (warp == 1 && elect_one()) {
        const uint64_t d = umma_desc_sw128(smem_u32(smem));
        for (int k = 0; k < 4; k++) tmem_cp_128x256b(tm + k * 8, d + (uint64_t)((k * 32) >> 4));
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int c = 0; c < 32; c++) out[(warp * 32 + lane) * 32 + c] = v[c];
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 32); }
}

int main() {
    uint32_t *d; static uint32_t h[128 * 32];
    cudaMalloc(&d, sizeof h);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024);
    probe<<<1, 128, 20 * 1024>>>(d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int lane = 0; lane < 128; lane++)
        for (int c = 0; c < 32; c++) {
            const int lo = h[lane * 32 + c] & 0xffff, hi = h[lane * 32 + c] >> 16;
            if (lo != lane * 64 + 2 * c || hi != lane * 64 + 2 * c + 1) bad++;
        }
    printf("expected layout (lane = row, column c = elements 2c, 2c+1): %d mismatches of 4096\n", bad);
    for (int lane : {0, 1, 2, 9, 33, 127}) {
        printf("lane %3d:", lane);
        for (int c = 0; c < 12; c++) {
            const int lo = h[lane * 32 + c] & 0xffff, hi = h[lane * 32 + c] >> 16;
            printf(" (%d,%d|%d,%d)", lo / 64, lo % 64, hi / 64, hi % 64);
        }
        printf("\n");
    }
    return 0;
}
