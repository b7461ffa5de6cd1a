// Is the ~32 B/clk/SM that TMA boxes of 128-byte row segments deliver a limit of the ISSUING SM's TMA unit or of the
// receiving SM's port?  Every CTA streams 32 KB stages of a 4.7 MB L2-resident matrix ([64 rows x 64 K] x 4 K-blocks,
// 128B swizzle, the W_ih access pattern of inproj_gemm.cu) through a 6-stage ring:
//   unicast          : each CTA loads its own stages
//   multicast (CS)   : clusters of CS CTAs; each CTA issues 1/CS of every stage and multicasts it to all CS CTAs
// Reported: bytes received per SM per SM-clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I xna_basecaller_b200/csrc tools/tma_multicast_microbench.cu -o tools/tma_mc.bin -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "xb_ptx.cuh"
using namespace xbptx;
namespace cg = cooperative_groups;

constexpr int STAGES = 6, STAGE_BYTES = 32768, KB_BYTES = 8192;

__device__ __forceinline__ void tma_load_3d_mc(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t cta) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}

template <int CS, bool MC>
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tmFull, const __grid_constant__ CUtensorMap tmPart,
                                                     long long *out, int iters) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full = (uint64_t *)(smem + STAGES * STAGE_BYTES), *empty = full + STAGES;
    const int warp = threadIdx.x >> 5;
    uint32_t rank = 0;
    if (CS > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], MC ? CS : 1); }
        fence_barrier_init();
    }
    if (CS > 1) cg::this_cluster().sync(); else __syncthreads();
    if (warp == 0) {
        if (elect_one()) {
            for (int it = 0; it < iters; it++) {
                const int s = it % STAGES, nt = it % 48, ks = (it / 48) % 3;
                mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                mbar_expect_tx(&full[s], STAGE_BYTES);
                if (MC) {
                    constexpr int PART_KB = 4 / CS;                       // K-blocks per part
                    tma_load_3d_mc(smem + s * STAGE_BYTES + rank * PART_KB * KB_BYTES, &tmPart, &full[s], 0, nt * 64,
                                   ks * 4 + rank * PART_KB, (uint16_t)((1 << CS) - 1));
                } else {
                    tma_load_3d(smem + s * STAGE_BYTES, &tmFull, &full[s], 0, nt * 64, ks * 4);
                }
            }
        }
    } else {
        long long t0 = 0;
        for (int it = 0; it < iters; it++) {
            const int s = it % STAGES;
            mbar_wait(&full[s], (it / STAGES) & 1);
            if (it == STAGES) t0 = clock64();
            if (threadIdx.x == 32) {
                if (MC) { for (int c = 0; c < CS; c++) mbar_arrive_remote(&empty[s], c); }
                else mbar_arrive(&empty[s]);
            }
            __syncwarp();
        }
        long long t1 = clock64();
        if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
    }
    if (CS > 1) cg::this_cluster().sync(); else __syncthreads();
}

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(encode_fn enc, void *base, int kblocks) {
    CUtensorMap m;
    cuuint64_t dims[3] = {64, 3072, 12};
    cuuint64_t strides[2] = {768 * 2, 128};
    cuuint32_t box[3] = {64, 64, (cuuint32_t)kblocks};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

template <int CS, bool MC> void run(encode_fn enc, void *w, long long *out, const char *name) {
    CUtensorMap full = make_map(enc, w, 4), part = make_map(enc, w, MC ? 4 / CS : 4);
    auto k = stream_kernel<CS, MC>;
    const int smem = STAGES * STAGE_BYTES + 1024 + 256, iters = 4000;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    int grid = 148 / CS * CS;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; rep++) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, k, full, part, out, iters);
        if (e != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(e)); return; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
    }
    long long h[148];
    cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double mx = 0, sum = 0;
    for (int i = 0; i < grid; i++) { sum += h[i]; if (h[i] > mx) mx = h[i]; }
    const double bytes = (double)(iters - STAGES) * STAGE_BYTES;
    printf("%-28s grid %3d: %.1f B/clk/SM received (mean), %.1f (slowest SM); %.0f cycles per 32 KB stage\n", name, grid,
           bytes / (sum / grid), bytes / mx, (sum / grid) / (iters - STAGES));
}

int main() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    encode_fn enc = (encode_fn)fn;
    void *w; long long *out;
    cudaMalloc(&w, 3072 * 768 * 2); cudaMemset(w, 0, 3072 * 768 * 2); cudaMalloc(&out, 148 * 8);
    run<1, false>(enc, w, out, "unicast");
    run<2, false>(enc, w, out, "unicast, clusters of 2");
    run<2, true>(enc, w, out, "multicast, clusters of 2");
    run<4, true>(enc, w, out, "multicast, clusters of 4");
    return 0;
}
