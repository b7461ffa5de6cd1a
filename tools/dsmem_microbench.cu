// Can the h exchange of the LSTM recurrence go through distributed shared memory?  A cluster of CS CTAs (one per SM)
// all-gathers: every round each CTA pushes its `bytes`-sized slice into the shared memory of every CTA of the cluster
// (cp.async.bulk shared::cta -> shared::cluster, complete_tx on the receiver's mbarrier) and waits until all CS slices
// of the round have landed at home.  Reported per piece size: cycles per round, bytes received per SM-clock, and the
// one-way latency of a single push (ping-pong between rank 0 and rank CS-1).  Also prints how many clusters of CS CTAs
// with 200 KB of shared memory can be resident (cudaOccupancyMaxActiveClusters).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I xna_basecaller_b200/csrc tools/dsmem_microbench.cu -o tools/dsmem.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "xb_ptx.cuh"
using namespace xbptx;

__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
    return r;
}
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ bool mbar_try_wait_cluster_acq(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// MODE 0: one thread issues all CS bulk copies; 1: CS lanes of warp 0 issue one each; 2: 128 threads write 16-byte
// st.shared::cluster pieces and one thread per destination arrives remotely (release.cluster)
template <int MODE>
__global__ void __launch_bounds__(128, 1) allgather_kernel(int CS, int bytes, int rounds, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    // [2 buffers][CS slices][bytes] receive area, then the send slice, then barriers
    uint8_t *recv = smem;
    uint8_t *send = smem + 2 * 16 * 6144;
    uint64_t *full = (uint64_t *)(send + 8192);
    const uint32_t rank = cluster_rank();
    if (threadIdx.x == 0) {
        mbar_init(&full[0], MODE == 2 ? CS : 1);
        mbar_init(&full[1], MODE == 2 ? CS : 1);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 8192 / 4; i += blockDim.x) ((uint32_t *)send)[i] = rank * 1000 + i;
    fence_proxy_async();
    __syncthreads();
    cluster_sync_all();
    long long t0 = clock64();
    for (int r = 0; r < rounds; r++) {
        const int b = r & 1;
        if (MODE == 0) {
            if (threadIdx.x == 0) {
                mbar_expect_tx(&full[b], (uint32_t)(CS * bytes));
                for (int d = 0; d < CS; d++) {
                    const uint32_t dst = (rank + d) % CS;
                    bulk_s2c(mapa(smem_u32(recv + (b * 16 + rank) * 6144), dst), smem_u32(send), bytes, mapa(smem_u32(&full[b]), dst));
                }
            }
        } else if (MODE == 1) {
            if (threadIdx.x == 0) mbar_expect_tx(&full[b], (uint32_t)(CS * bytes));
            if (threadIdx.x < CS) {
                const uint32_t dst = (rank + threadIdx.x) % CS;
                bulk_s2c(mapa(smem_u32(recv + (b * 16 + rank) * 6144), dst), smem_u32(send), bytes, mapa(smem_u32(&full[b]), dst));
            }
        } else {
            const int pieces = bytes / 16;
            for (int i = threadIdx.x; i < pieces * CS; i += blockDim.x) {
                const uint32_t dst = (rank + i / pieces) % CS, pc = i % pieces;
                st_cluster_v4(mapa(smem_u32(recv + (b * 16 + rank) * 6144 + pc * 16), dst), ((const uint4 *)send)[pc]);
            }
            __syncthreads();
            if (threadIdx.x < CS) mbar_arrive_cluster(mapa(smem_u32(&full[b]), (rank + threadIdx.x) % CS));
        }
        if (MODE == 2) {
            while (!mbar_try_wait_cluster_acq(&full[b], (r >> 1) & 1)) {
            }
        } else {
            mbar_wait(&full[b], (r >> 1) & 1);
        }
        __syncthreads();
    }
    long long t1 = clock64();
    cluster_sync_all();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    // keep the received bytes alive
    if (recv[(threadIdx.x * 16) % bytes] == 255 && rounds < 0) out[0] = 0;
}

// ping-pong between rank 0 and rank CS-1 with one `bytes`-sized push each way
__global__ void __launch_bounds__(32, 1) pingpong_kernel(int CS, int bytes, int rounds, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t *recv = smem, *send = smem + 8192;
    uint64_t *full = (uint64_t *)(send + 8192);
    const uint32_t rank = cluster_rank();
    if (threadIdx.x == 0) { mbar_init(full, 1); fence_barrier_init(); }
    fence_proxy_async();
    cluster_sync_all();
    const uint32_t peer = rank == 0 ? CS - 1 : 0;
    long long t0 = clock64();
    if (threadIdx.x == 0 && (rank == 0 || rank == (uint32_t)CS - 1)) {
        for (int r = 0; r < rounds; r++) {
            mbar_expect_tx(full, bytes);
            if (rank == 0) bulk_s2c(mapa(smem_u32(recv), peer), smem_u32(send), bytes, mapa(smem_u32(full), peer));
            mbar_wait(full, r & 1);
            if (rank != 0) bulk_s2c(mapa(smem_u32(recv), peer), smem_u32(send), bytes, mapa(smem_u32(full), peer));
        }
    }
    long long t1 = clock64();
    cluster_sync_all();
    if (threadIdx.x == 0 && rank == 0) out[blockIdx.x / CS] = t1 - t0;
}

template <typename K>
static int launch(K kern, int grid, int threads, int CS, size_t smem, int bytes, int rounds, long long *out) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = -1;
    cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, CS, bytes, rounds, out);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("launch failed (CS %d): %s\n", CS, cudaGetErrorString(e)); cudaGetLastError(); return -1; }
    return ncl;
}

int main() {
    long long *out;
    cudaMallocManaged(&out, 1024 * sizeof(long long));
    const int rounds = 400;
    for (int CS : {8, 16}) {
        for (int nclusters : {1, 148 / CS}) {
            for (int bytes : {768, 1536, 3072, 6144}) {
                for (int mode = 0; mode < 3; mode++) {
                    const size_t sm = 2 * 16 * 6144 + 8192 + 2048;   // receive area: 6 KB pitch
                    if (sm > 227 * 1024) { printf("smem too large\n"); return 1; }
                    int ncl = mode == 0 ? launch(allgather_kernel<0>, nclusters * CS, 128, CS, sm, bytes, rounds, out)
                            : mode == 1 ? launch(allgather_kernel<1>, nclusters * CS, 128, CS, sm, bytes, rounds, out)
                                        : launch(allgather_kernel<2>, nclusters * CS, 128, CS, sm, bytes, rounds, out);
                    if (ncl < 0) continue;
                    long long mx = 0; double mean = 0;
                    for (int i = 0; i < nclusters * CS; i++) { mx = out[i] > mx ? out[i] : mx; mean += out[i]; }
                    mean /= nclusters * CS;
                    printf("CS %2d clusters %2d (max resident %2d) mode %d piece %5d B: %7.0f cycles/round (slowest %7.0f), %5.1f B/clk/SM received\n",
                           CS, nclusters, ncl, mode, bytes, mean / rounds, (double)mx / rounds, (double)CS * bytes / (mean / rounds));
                }
            }
        }
        for (int bytes : {16, 1536, 3072}) {
            int ncl = launch(pingpong_kernel, CS, 32, CS, 8192 * 2 + 2048, bytes, rounds, out);
            if (ncl >= 0) printf("CS %2d ping-pong %5d B: one way %6.0f cycles\n", CS, bytes, (double)out[0] / rounds / 2);
        }
    }
    return 0;
}
