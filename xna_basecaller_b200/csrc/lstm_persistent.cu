// Persistent LSTM recurrence for sm_100a: one cooperative launch runs all T steps of one layer with W_hh
// resident on chip (bonito/nn.py:176-235, RNNWrapper.forward :189-193; torch.nn.LSTM gate order i,f,g,o, zero
// initial state; a reversed layer walks time backwards by indexing instead of the reference's two flips).
//
// Decomposition.  W_hh (3072 x 768, 4.7 MB 16-bit) does not fit one SM, so it is cut into 24 tiles of 128 gate
// rows = [i|f|g|o] x 32 hidden units (rows are permuted at weight-load time).  The batch is cut into G <= 6
// groups of <= 96 chunks; CTA (g, j) owns tile j for group g, so 24*G <= 144 CTAs are resident, one per SM.
//   - the W_hh tile lives in TENSOR MEMORY for the whole kernel (128 lanes x 384 columns, the A operand of
//     tcgen05.mma), loaded once with tcgen05.st;
//   - per step the CTA loads h_{t-1} of its group (96 x 768, 147 KB) into shared memory with 12 TMA boxes
//     (128B swizzle; the B operand), issues 48 tcgen05.mma (M=128 gate rows, N=96 chunks, K=16) into a
//     96-column fp32 accumulator in TMEM, adds the hoisted input projection G[t] (TMA-prefetched one step
//     ahead), applies the cell update with the fp32 cell state held in registers, and writes its 32-unit slice
//     of h_t to HBM;
//   - the 24 CTAs of a group exchange h_t through L2: release-add on a per-group counter, acquire-poll by
//     the TMA producer of each CTA, then a proxy fence before the next TMA load.
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..11
// epilogue (phase 1: TMEM -> +G -> fp32 scratch in smem, re-using the dead h buffer; phase 2: one thread
// per (unit, chunk) cell with all four gates, no cross-lane exchange, 12 cells per thread).
#include <stdlib.h>

#include "xb_common.cuh"
#include "xb_ptx.cuh"
#include "xb_gemm.cuh"

using namespace xbptx;

namespace {

constexpr int NB = 96;                     // chunks (MMA N) per group
constexpr int TILES = 24;                  // gate tiles per group
constexpr int KCH = XB_FEATURES / 64;      // 12 K chunks of 64
constexpr int H_CHUNK_BYTES = NB * 64 * 2; // 12288
constexpr int H_BYTES = KCH * H_CHUNK_BYTES;        // 147456
constexpr int G_BYTES = NB * 128 * 2;               // 24576 per buffer
constexpr int P_STRIDE = NB + 1;                    // fp32 scratch row stride (bank-conflict free)
constexpr int P_BYTES = 128 * P_STRIDE * 4;         // 49664
constexpr int HS_STRIDE_W = 17;                     // staging row stride in 32-bit words (16 data + 1 pad)
constexpr int HS_OFFSET = ((P_BYTES + 127) / 128) * 128;
constexpr int SMEM_BYTES = H_BYTES + 2 * G_BYTES + 512 + 1024;
constexpr int THREADS = 384;
constexpr int EPI_THREADS = 256;
constexpr int D_COL = 384;                          // accumulator columns start after the 384 columns of W_hh
static_assert(HS_OFFSET + NB * HS_STRIDE_W * 4 <= H_BYTES, "epilogue scratch must fit in the h buffer");

struct PLParams {
    int T, N, reverse;
    int batch0, nbatch, G;      // batch rows [batch0, batch0 + nbatch) are split into G groups
    const uint16_t *w_hh;       // (3072, 768) 16-bit, tile-permuted rows
    uint16_t *y;                // (T, N, 768) 16-bit output = hidden states
    int *counters;              // (G) zeroed before launch
    long long *dbg;             // optional timeline (XB_LSTM_DEBUG=1): clock64 stamps of CTA 0, steps 64..71
};

#define DBG(ev) do { if (p.dbg && blockIdx.x == 0 && s >= 64 && s < 72) p.dbg[(s - 64) * 16 + (ev)] = clock64(); } while (0)

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

template <bool BF16>
__global__ void __launch_bounds__(THREADS, 1)
lstm_persistent_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmG, const PLParams p) {
    using X = xb16<BF16>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *hbuf = smem;
    uint8_t *gbuf = smem + H_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + H_BYTES + 2 * G_BYTES);
    uint64_t *h_full = bars;                 // [12]
    uint64_t *d_full = bars + 12;
    uint64_t *d_empty = bars + 13;
    uint64_t *g_full = bars + 14;            // [2]
    uint64_t *g_empty = bars + 16;           // [2]
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(bars + 18);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x / TILES, j = blockIdx.x % TILES;
    const int T = p.T, N = p.N;
    const int b0 = p.batch0 + (int)(((long long)g * p.nbatch) / p.G);
    const int b1 = p.batch0 + (int)(((long long)(g + 1) * p.nbatch) / p.G);
    const int count = b1 - b0;               // <= NB valid chunks in this group

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmY);
        prefetch_tmap(&tmG);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < KCH; i++) mbar_init(&h_full[i], 1);
        mbar_init(d_full, 1);
        mbar_init(d_empty, EPI_THREADS / 32);
        for (int i = 0; i < 2; i++) {
            mbar_init(&g_full[i], 1);
            mbar_init(&g_empty[i], EPI_THREADS / 32);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_holder, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    // W_hh tile -> tensor memory: lane = gate row of the tile, column c holds elements k = 2c, 2c+1
    if (warp >= 4 && warp < 8) {
        const int q = warp & 3, r = q * 32 + lane;
        const uint4 *src = reinterpret_cast<const uint4 *>(p.w_hh + ((size_t)j * 128 + r) * XB_FEATURES);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < XB_FEATURES / 2; c0 += 32) {
            uint32_t v[32];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint4 t4 = __ldg(src + c0 / 4 + i);
                v[4 * i] = t4.x; v[4 * i + 1] = t4.y; v[4 * i + 2] = t4.z; v[4 * i + 3] = t4.w;
            }
            tmem_st_32x32b_x32(taddr + c0, v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            const int t0 = p.reverse ? T - 1 : 0;
            mbar_expect_tx(&g_full[0], G_BYTES);
            tma_load_2d(gbuf, &tmG, &g_full[0], j * 128, t0 * N + b0);
            for (int s = 0; s < T; s++) {
                const int t = p.reverse ? T - 1 - s : s;
                if (s > 0) {
                    const int tp = p.reverse ? t + 1 : t - 1;
                    const int need = TILES * s;
                    DBG(0);
                    while (ld_acquire_gpu(p.counters + g) < need) {
                    }
                    DBG(1);
                    fence_proxy_async_all();
#pragma unroll 1
                    for (int kc = 0; kc < KCH; kc++) {
                        mbar_expect_tx(&h_full[kc], H_CHUNK_BYTES);
                        tma_load_2d(hbuf + kc * H_CHUNK_BYTES, &tmY, &h_full[kc], kc * 64, tp * N + b0);
                    }
                    DBG(2);
                }
                if (s + 1 < T) {      // input projection of the next step, one step ahead
                    const int sn = s + 1, q = sn & 1, u = sn >> 1;
                    const int tn = p.reverse ? T - 1 - sn : sn;
                    if (u >= 1) mbar_wait(&g_empty[q], (u - 1) & 1);
                    mbar_expect_tx(&g_full[q], G_BYTES);
                    tma_load_2d(gbuf + q * G_BYTES, &tmG, &g_full[q], j * 128, tn * N + b0);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(BF16 ? 1u : 0u, 128, NB);
        const uint32_t hb = smem_u32(hbuf);
        for (int s = 1; s < T; s++) {
            mbar_wait(d_empty, (s - 1) & 1);          // epilogue of step s-1 has drained the accumulator
            tc_fence_after();
            if (lane == 0) DBG(3);
#pragma unroll 1
            for (int kc = 0; kc < KCH; kc++) {
                mbar_wait(&h_full[kc], (s - 1) & 1);
                tc_fence_after();
                if (lane == 0 && kc == 0) DBG(4);
                if (lane == 0 && kc == KCH - 1) DBG(5);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        mma_f16_ts(tmem_base + D_COL, tmem_base + kc * 32 + k * 8,
                                   umma_desc_sw128(hb + kc * H_CHUNK_BYTES + k * 32), idesc, (kc | k) != 0);
                }
                __syncwarp();
            }
            if (elect_one()) mma_commit(d_full);
            __syncwarp();
            if (lane == 0) DBG(6);
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (256 threads)
        const int e = warp - 4, q = warp & 3, half = e >> 2;
        const int et = threadIdx.x - 128;                    // 0..255
        const int r = q * 32 + lane;                         // gate row of the tile == TMEM lane
        float *P = reinterpret_cast<float *>(hbuf);
        uint32_t *HS = reinterpret_cast<uint32_t *>(hbuf + HS_OFFSET);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + D_COL + half * 48;
        float cst[12];
#pragma unroll
        for (int i = 0; i < 12; i++) cst[i] = 0.0f;

        for (int s = 0; s < T; s++) {
            const int t = p.reverse ? T - 1 - s : s;
            const uint16_t *Gs = reinterpret_cast<const uint16_t *>(gbuf + (s & 1) * G_BYTES);
            mbar_wait(&g_full[s & 1], (s >> 1) & 1);
            if (s > 0) {
                mbar_wait(d_full, (s - 1) & 1);
                tc_fence_after();
            }
            if (et == 0) DBG(7);
            // phase 1: accumulator (+ G) -> fp32 scratch P[row][chunk]
#pragma unroll
            for (int cc = 0; cc < 3; cc++) {
                uint32_t acc[16];
                if (s > 0) {
                    tmem_ld_32x32b_x16(taddr + cc * 16, acc);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) acc[i] = 0u;
                }
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int b = half * 48 + cc * 16 + i;
                    typename X::T gv = *reinterpret_cast<const typename X::T *>(Gs + b * 128 + r);
                    P[r * P_STRIDE + b] = __uint_as_float(acc[i]) + X::to(gv);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(d_empty);
                mbar_arrive(&g_empty[s & 1]);
            }
            named_bar_sync(1, EPI_THREADS);
            if (et == 0) DBG(8);
            // phase 2: cells (unit ul = 4e + ui, chunk b = lane + 32m)
#pragma unroll
            for (int ui = 0; ui < 4; ui++) {
                const int ul = 4 * e + ui;
#pragma unroll
                for (int m = 0; m < 3; m++) {
                    const int b = lane + 32 * m;
                    const float ig = sigmoid_f(P[(ul)*P_STRIDE + b]);
                    const float fg = sigmoid_f(P[(32 + ul) * P_STRIDE + b]);
                    const float gg = tanh_f(P[(64 + ul) * P_STRIDE + b]);
                    const float og = sigmoid_f(P[(96 + ul) * P_STRIDE + b]);
                    const float cn = fg * cst[ui * 3 + m] + ig * gg;
                    cst[ui * 3 + m] = cn;
                    const float hn = og * tanh_f(cn);
                    typename X::T hv = X::from(hn);
                    reinterpret_cast<uint16_t *>(HS + b * HS_STRIDE_W)[ul] = *reinterpret_cast<uint16_t *>(&hv);
                }
            }
            named_bar_sync(1, EPI_THREADS);
            if (et == 0) DBG(9);
            // coalesced store of the (chunks x 32 units) slice: 16 words per chunk row
            uint32_t *yrow = reinterpret_cast<uint32_t *>(p.y + ((size_t)t * N + b0) * XB_FEATURES + j * 32);
            for (int v = et; v < NB * 16; v += EPI_THREADS) {
                const int b = v >> 4, wv = v & 15;
                if (b < count) yrow[(size_t)b * (XB_FEATURES / 2) + wv] = HS[b * HS_STRIDE_W + wv];
            }
            if (et == 0) DBG(10);
            __threadfence();
            fence_proxy_async_all();
            named_bar_sync(1, EPI_THREADS);
            if (et == 0) DBG(11);
            if (et == 0) red_release_gpu_add(p.counters + g, 1);
            if (et == 0) DBG(12);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

// debug: copy the timeline of the last launch (8 steps x 16 stamps) to the host
extern "C" int xb_debug_lstm_timeline(xb_handle *h, long long *out_host) {
    if (!h || !h->lstm_counters) return XB_ERR_STATE;
    XB_CUDA(h, cudaDeviceSynchronize());
    XB_CUDA(h, cudaMemcpy(out_host, h->lstm_counters + 16, 8 * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
    return XB_OK;
}

// Recurrent part of one LSTM layer.  h->gates must already hold the input projection (T*N, 3072).
int xb_lstm_recurrence_persistent(xb_handle *h, int layer, void *y_tnc, int T, int N, int reverse, cudaStream_t s) {
    if (!h->lstm_counters) {
        void *q = nullptr;
        XB_CUDA(h, cudaMalloc(&q, 64 * sizeof(int) + 8 * 16 * sizeof(long long)));
        h->lstm_counters = reinterpret_cast<int *>(q);
    }
    CUtensorMap tmY, tmG;
    if (int rc = xb_make_tmap_2d_box(h, &tmY, y_tnc, (uint64_t)T * N, XB_FEATURES, XB_FEATURES, 64, NB, 1)) return rc;
    if (int rc = xb_make_tmap_2d_box(h, &tmG, h->gates, (uint64_t)T * N, XB_GATES, XB_GATES, 128, NB, 0)) return rc;
    const int max_groups = h->num_sms / TILES;                 // 6 on a 148-SM B200
    const int block_cap = max_groups * NB;
    for (int batch0 = 0; batch0 < N; batch0 += block_cap) {
        PLParams p;
        p.T = T; p.N = N; p.reverse = reverse;
        p.batch0 = batch0;
        p.nbatch = (N - batch0 < block_cap) ? N - batch0 : block_cap;
        p.G = (p.nbatch + NB - 1) / NB;
        p.w_hh = reinterpret_cast<const uint16_t *>(h->lstm[layer].w_hh);
        p.y = reinterpret_cast<uint16_t *>(y_tnc);
        p.counters = h->lstm_counters;
        p.dbg = getenv("XB_LSTM_DEBUG") ? reinterpret_cast<long long *>(h->lstm_counters + 16) : nullptr;
        XB_CUDA(h, cudaMemsetAsync(h->lstm_counters, 0, 16 * sizeof(int), s));
        void *args[] = {(void *)&tmY, (void *)&tmG, (void *)&p};
        const void *fn = h->bf16 ? (const void *)lstm_persistent_kernel<true> : (const void *)lstm_persistent_kernel<false>;
        static bool configured[2] = {false, false};
        if (!configured[h->bf16]) {
            XB_CUDA(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            configured[h->bf16] = true;
        }
        XB_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3(p.G * TILES), dim3(THREADS), args, SMEM_BYTES, s));
        h->launches++;
    }
    return XB_OK;
}
