// How long does an un-fenced store take to become visible to another SM, depending on what the writer does next?
// Writer (CTA 0) stores a tagged 16-byte unit holding its globaltimer; reader (CTA peer) polls and reports
// globaltimer_now - timestamp (ns).  after: 0 = writer spins on ALU only, 1 = writer issues an unrelated relaxed
// load, 2 = MEMBAR.GPU, 3 = writer waits on a shared-memory mbarrier-like spin (ld.shared), 4 = st.cg-style store.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/visibility_microbench.cu -o tools/vis_mb.bin
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void st_relaxed_v4(uint4 *p, uint4 v) { asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void st_weak_v4(uint4 *p, uint4 v) { asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void st_cg_v4(uint4 *p, uint4 v) { asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4 *p) { uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) { uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

__global__ void vis(uint4 *slot, uint32_t *other, unsigned long long *out, int iters, int store_kind, int after, int peer) {
    __shared__ volatile int sflag;
    if (threadIdx.x != 0) return;
    if (blockIdx.x == 0) {
        for (int i = 1; i <= iters; i++) {
            unsigned long long ts = gtime();
            uint4 v = make_uint4(i, (uint32_t)ts, (uint32_t)(ts >> 32), 0);
            if (store_kind == 0) st_relaxed_v4(slot, v); else if (store_kind == 1) st_weak_v4(slot, v); else st_cg_v4(slot, v);
            if (after == 1) { uint32_t d = ld_relaxed_u32(other + 64); if (d == 0x12345) out[20] = d; }
            if (after == 2) __threadfence();
            long long c0 = clock64();
            if (after == 3) { while (clock64() - c0 < 20000) { if (sflag == 12345) out[21] = 1; } }
            else { while (clock64() - c0 < 20000) {} }
        }
    } else if (blockIdx.x == peer) {
        unsigned long long sum = 0, mx = 0;
        for (int i = 1; i <= iters; i++) {
            uint4 v;
            do { v = ld_relaxed_v4(slot); } while (v.x != (uint32_t)i);
            unsigned long long now = gtime();
            unsigned long long ts = ((unsigned long long)v.z << 32) | v.y;
            unsigned long long d = now - ts;
            sum += d; if (d > mx) mx = d;
        }
        out[0] = sum / iters; out[1] = mx;
    }
}

int main() {
    uint4 *slot; uint32_t *other; unsigned long long *out;
    cudaMalloc(&slot, 4096); cudaMalloc(&other, 4096); cudaMalloc(&out, 256);
    const char *sk[] = {"st.relaxed.gpu", "st (weak)", "st.cg"};
    const char *af[] = {"ALU spin", "unrelated ld.relaxed", "MEMBAR.GPU", "ld.shared spin"};
    for (int peer : {1, 147})
        for (int store_kind = 0; store_kind < 3; store_kind++)
            for (int after = 0; after < 4; after++) {
                cudaMemset(slot, 0, 4096); cudaMemset(out, 0, 256);
                int iters = 300;
                void *args[] = {&slot, &other, &out, &iters, &store_kind, &after, &peer};
                cudaLaunchCooperativeKernel((void *)vis, dim3(148), dim3(32), args, 0, 0);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
                unsigned long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                printf("peer %3d  %-15s then %-22s: visible after mean %5llu ns, max %5llu ns\n", peer, sk[store_kind], af[after], h[0], h[1]);
            }
    return 0;
}
