// Latency of ONE isolated 48 KB load from L2 into shared memory -- the h_{t-1} fetch on the LSTM's serial chain
// (lstm_persistent.cu), where nothing else is in flight to hide the first-byte latency:
//   box3d      one 3-D tensor box {64 k, 32 rows, 12 k-blocks}, 128B swizzle (what the kernel does)
//   box3d x2/4 the same split along k-blocks, each part on its own mbarrier (first / last completion)
//   bulk       one 1-D cp.async.bulk of 48 KB contiguous bytes (the exchange buffer would be written in the smem image)
//   bulk x2/4  the same in 2 / 4 pieces
// Every iteration reads a different, L2-resident region.  grid = 1 (idle chip) and 144 (all SMs doing the same).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I xna_basecaller_b200/csrc tools/tma_latency_microbench.cu -o tools/tma_lat.bin -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "xb_ptx.cuh"
using namespace xbptx;

constexpr int BYTES = 49152, REGIONS = 8, ROWS = 32;

__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// MODE 0: tensor boxes, 1: bulk copies.  PARTS pieces.
template <int MODE, int PARTS>
__global__ void __launch_bounds__(32, 1) lat_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t *src, long long *out, int iters) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = (uint64_t *)(smem + BYTES);
    if (threadIdx.x == 0) {
        for (int i = 0; i < PARTS; i++) mbar_init(&bar[i], 1);
        fence_barrier_init();
    }
    __syncwarp();
    long long first = 0, last = 0;
    if (elect_one()) {
        for (int it = 0; it < iters; it++) {
            const int region = (blockIdx.x * REGIONS + it % REGIONS);
            const long long t0 = clock64();
#pragma unroll
            for (int p = 0; p < PARTS; p++) {
                mbar_expect_tx(&bar[p], BYTES / PARTS);
                if (MODE == 0) tma_load_3d(smem + p * (BYTES / PARTS), &tm, &bar[p], 0, region * ROWS, p * (12 / PARTS));
                else bulk_load(smem + p * (BYTES / PARTS), src + (size_t)region * BYTES + p * (BYTES / PARTS), BYTES / PARTS, &bar[p]);
            }
            mbar_wait(&bar[0], it & 1);
            const long long t1 = clock64();
#pragma unroll
            for (int p = 1; p < PARTS; p++) mbar_wait(&bar[p], it & 1);
            const long long t2 = clock64();
            if (it >= REGIONS) { first += t1 - t0; last += t2 - t0; }      // first pass warms L2
        }
        out[blockIdx.x * 2] = first;
        out[blockIdx.x * 2 + 1] = last;
    }
}

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE, int PARTS> void run(encode_fn enc, void *buf, long long *out, int grid, const char *name) {
    CUtensorMap m;
    cuuint64_t dims[3] = {64, (cuuint64_t)148 * REGIONS * ROWS, 12};
    cuuint64_t strides[2] = {768 * 2, 128};
    cuuint32_t box[3] = {64, ROWS, 12 / PARTS};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    auto k = lat_kernel<MODE, PARTS>;
    const int smem = BYTES + 1024 + 256, iters = 8 * REGIONS + REGIONS;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; rep++) {
        k<<<grid, 32, smem>>>(m, (const uint8_t *)buf, out, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
    }
    long long h[296];
    cudaMemcpy(h, out, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
    double f = 0, l = 0;
    for (int i = 0; i < grid; i++) { f += h[2 * i]; l += h[2 * i + 1]; }
    const double n = (double)grid * (iters - REGIONS);
    printf("%-12s grid %3d: first part complete after %6.0f cycles, all 48 KB after %6.0f cycles\n", name, grid, f / n, l / n);
}

int main() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    encode_fn enc = (encode_fn)fn;
    void *buf; long long *out;
    const size_t bytes = (size_t)148 * REGIONS * BYTES;
    cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes); cudaMalloc(&out, 296 * 8);
    for (int grid : {1, 144}) {
        run<0, 1>(enc, buf, out, grid, "box3d");
        run<0, 2>(enc, buf, out, grid, "box3d x2");
        run<0, 4>(enc, buf, out, grid, "box3d x4");
        run<1, 1>(enc, buf, out, grid, "bulk");
        run<1, 2>(enc, buf, out, grid, "bulk x2");
        run<1, 4>(enc, buf, out, grid, "bulk x4");
    }
    return 0;
}
