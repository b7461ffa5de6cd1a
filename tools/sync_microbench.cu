// Microbenchmark of the inter-SM signalling primitives the persistent LSTM depends on (B200, sm_100a):
//   1. cost of MEMBAR.GPU / red.release with and without a store in front of it;
//   2. ping-pong between two CTAs on different SMs through L2: cycles per one-way hand-off for
//      relaxed flag only | 16-byte payload + release/acquire flag | payload with the tag inside (no fence).
// One cooperative launch with N CTAs (all co-resident); CTA 0 and CTA `peer` play, the rest idle or make noise.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I xna_basecaller_b200/csrc tools/sync_microbench.cu -o tools/sync_mb.bin
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "xb_ptx.cuh"
using namespace xbptx;

__device__ __forceinline__ void st_relaxed_u32(uint32_t *p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) { uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_relaxed_v4(uint4 *p, uint4 v) { asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4 *p) { uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// mode 3: payload by a 256-byte bulk async store (smem -> global), wait_group, then a RELAXED flag; reader polls the
// flag relaxed and checks the payload with relaxed loads (counts stale reads)
// mode 0: relaxed flag ping-pong; 1: payload (plain st) + st.release flag / ld.acquire; 2: tagged 16-byte payload, relaxed
__global__ void pingpong(uint32_t *flags, uint4 *payload, long long *out, int iters, int mode, int peer) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x == 0 ? 0 : (blockIdx.x == peer ? 1 : -1);
    if (me < 0) return;
    uint32_t *mine = flags + me * 64, *other = flags + (1 - me) * 64;
    uint4 *pm = payload + me * 64, *po = payload + (1 - me) * 64;
    __shared__ __align__(128) uint4 sbuf[16];
    long long t0 = clock64();
    for (int i = 1; i <= iters; i++) {
        if (me == 0) {
            if (mode == 0) { st_relaxed_u32(mine, i); while (ld_relaxed_u32(other) != (uint32_t)i) {} }
            else if (mode == 1) { pm[1] = make_uint4(i, i, i, i); st_release_u32(mine, i); while (ld_acquire_u32(other) != (uint32_t)i) {} if (ld_relaxed_v4(po + 1).x != (uint32_t)i) out[7]++; }
            else if (mode == 2) { st_relaxed_v4(pm, make_uint4(i, 1, 2, 3)); while (ld_relaxed_v4(po).x != (uint32_t)i) {} }
            else {
                for (int k = 0; k < 16; k++) sbuf[k] = make_uint4(i, k, i, k);
                fence_proxy_async();
                bulk_store(pm + 16, sbuf, 256);
                st_relaxed_u32(mine, i);
                while (ld_relaxed_u32(other) != (uint32_t)i) {}
                if (ld_relaxed_v4(po + 16).x != (uint32_t)i || ld_relaxed_v4(po + 31).z != (uint32_t)i) out[7]++;
            }
        } else {
            if (mode == 0) { while (ld_relaxed_u32(other) != (uint32_t)i) {} st_relaxed_u32(mine, i); }
            else if (mode == 1) { while (ld_acquire_u32(other) != (uint32_t)i) {} if (ld_relaxed_v4(po + 1).x != (uint32_t)i) out[7]++; pm[1] = make_uint4(i, i, i, i); st_release_u32(mine, i); }
            else if (mode == 2) { while (ld_relaxed_v4(po).x != (uint32_t)i) {} st_relaxed_v4(pm, make_uint4(i, 1, 2, 3)); }
            else {
                while (ld_relaxed_u32(other) != (uint32_t)i) {}
                if (ld_relaxed_v4(po + 16).x != (uint32_t)i || ld_relaxed_v4(po + 31).z != (uint32_t)i) out[7]++;
                for (int k = 0; k < 16; k++) sbuf[k] = make_uint4(i, k, i, k);
                fence_proxy_async();
                bulk_store(pm + 16, sbuf, 256);
                st_relaxed_u32(mine, i);
            }
        }
    }
    long long t1 = clock64();
    if (me == 0) out[mode] = (t1 - t0) / (2 * iters);
}

// cost of a fence for one thread: 0 nothing, 1 MEMBAR.GPU alone, 2 16-byte store + MEMBAR.GPU, 3 store + red.release, 4 store only
__global__ void fence_cost(uint4 *buf, int *ctr, long long *out, int iters, int mode) {
    if (threadIdx.x != 0) return;
    uint4 *mine = buf + (size_t)blockIdx.x * 4096;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (mode == 2 || mode == 3 || mode == 4) mine[i & 255] = make_uint4(i, i, i, i);
        if (mode == 1 || mode == 2) __threadfence();
        if (mode == 3) red_release_gpu_add(ctr + blockIdx.x * 32, 1);
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[8 + mode] = (t1 - t0) / iters;
}

int main(int argc, char **argv) {
    int nsm = 148;
    uint32_t *flags; uint4 *payload, *buf; long long *out; int *ctr;
    cudaMalloc(&flags, 4096); cudaMalloc(&payload, 8192); cudaMalloc(&out, 256); cudaMalloc(&ctr, 148 * 128);
    cudaMalloc(&buf, (size_t)148 * 4096 * 16);
    long long h[32];
    for (int peer : {1, 2, 37, 74, 111, 147}) {
        cudaMemset(out, 0, 256);
        for (int mode = 0; mode < 4; mode++) {
            cudaMemset(flags, 0, 4096); cudaMemset(payload, 0, 8192);
            int iters = 2000;
            void *args[] = {&flags, &payload, &out, &iters, &mode, &peer};
            cudaLaunchCooperativeKernel((void *)pingpong, dim3(nsm), dim3(32), args, 0, 0);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, out, 256, cudaMemcpyDeviceToHost);
        printf("peer CTA %3d: one-way hand-off  relaxed flag %5lld cyc | payload + release/acquire %5lld cyc | tagged 16 B relaxed %5lld cyc | bulk store + wait_group + relaxed flag %5lld cyc (stale payload reads: %lld)\n",
               peer, h[0], h[1], h[2], h[3], h[7]);
    }
    for (int grid : {1, 148}) {
        cudaMemset(out, 0, 256);
        for (int mode = 0; mode < 5; mode++) {
            fence_cost<<<grid, 32>>>(buf, ctr, out, 2000, mode);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, out, 256, cudaMemcpyDeviceToHost);
        printf("grid %3d: loop %lld | MEMBAR.GPU %lld | st+MEMBAR.GPU %lld | st+red.release %lld | st only %lld cycles per iteration\n",
               grid, h[8], h[9], h[10], h[11], h[12]);
    }
    return 0;
}
