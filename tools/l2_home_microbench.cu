// Which L2 partition (die) is a line's home, as seen from one SM?  One thread times dependent ld.global.cg loads (L2
// hits after a warm-up pass) of one word per 128-byte line over a buffer and prints the latency histogram and the
// run lengths of "near" / "far" lines along the address space -- i.e. the granularity at which addresses alternate
// between the two dies of a B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/l2_home_microbench.cu -o tools/l2_home.bin
#include <cstdio>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__global__ void probe(const int *buf, int lines, int stride_ints, unsigned short *lat, int *smid_out) {
    if (threadIdx.x != 0) return;
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    *smid_out = smid;
    int sink = 0;
    for (int pass = 0; pass < 2; pass++)
        for (int i = 0; i < lines; i++) {
            const int *p = buf + (size_t)i * stride_ints + (sink & 1);      // dependent on the previous load
            long long t0 = clock64();
            int v;
            asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
            sink += v;
            long long t1 = clock64();
            if (pass == 1) lat[i] = (unsigned short)min(65535ll, t1 - t0 + (sink & 1));
        }
}

int main() {
    const int lines = 1 << 17;                       // 16 MB of 128-byte lines: stays in L2
    int *buf; unsigned short *lat; int *smid;
    cudaMalloc(&buf, (size_t)lines * 128); cudaMemset(buf, 0, (size_t)lines * 128);
    cudaMalloc(&lat, lines * 2); cudaMalloc(&smid, 4);
    for (int rep = 0; rep < 2; rep++) {
        probe<<<1, 32>>>(buf, lines, 32, lat, smid);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed\n"); return 1; }
    }
    std::vector<unsigned short> h(lines);
    int hs; cudaMemcpy(h.data(), lat, lines * 2, cudaMemcpyDeviceToHost); cudaMemcpy(&hs, smid, 4, cudaMemcpyDeviceToHost);
    std::vector<unsigned short> s(h); std::sort(s.begin(), s.end());
    printf("SM %d: L2-hit latency percentiles (cycles): p1 %d p10 %d p25 %d p50 %d p75 %d p90 %d p99 %d\n", hs, s[lines / 100],
           s[lines / 10], s[lines / 4], s[lines / 2], s[3 * lines / 4], s[9 * lines / 10], s[99 * lines / 100]);
    int hist[40] = {0};
    for (int i = 0; i < lines; i++) hist[std::min(39, h[i] / 25)]++;
    for (int b = 0; b < 40; b++) if (hist[b] > lines / 500) printf("  %4d-%4d cycles: %6.2f %%\n", b * 25, b * 25 + 24, 100.0 * hist[b] / lines);
    // near/far split at the median of p10 and p90
    const int thr = (s[lines / 10] + s[9 * lines / 10]) / 2;
    int runs[16] = {0}, cur = 1;
    for (int i = 1; i < lines; i++) {
        if ((h[i] > thr) == (h[i - 1] > thr)) cur++;
        else { int b = 0; while ((1 << (b + 1)) <= cur && b < 15) b++; runs[b]++; cur = 1; }
    }
    printf("threshold %d cycles; run lengths of same-side lines (in 128 B lines):\n", thr);
    for (int b = 0; b < 16; b++) if (runs[b]) printf("  %5d..%5d lines: %d runs\n", 1 << b, (2 << b) - 1, runs[b]);
    printf("first 64 lines: ");
    for (int i = 0; i < 64; i++) printf("%c", h[i] > thr ? 'F' : 'n');
    printf("\nlines 0,32,64.. (4 KB apart): ");
    for (int i = 0; i < 64; i++) printf("%c", h[i * 32] > thr ? 'F' : 'n');
    printf("\n");
    return 0;
}
