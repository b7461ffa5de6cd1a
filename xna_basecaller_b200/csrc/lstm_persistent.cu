// Persistent LSTM recurrence for sm_100a: one cooperative launch runs all T steps of one layer with W_hh
// resident on chip (bonito/nn.py:176-235, RNNWrapper.forward :189-193; torch.nn.LSTM gate order i,f,g,o, zero
// initial state; a reversed layer walks time backwards by indexing instead of the reference's two flips).
//
// Decomposition.  W_hh (3072 x 768, 4.7 MB 16-bit) does not fit one SM, so it is cut into 24 tiles of 128 gate
// rows = [i|f|g|o] x 32 hidden units (rows are permuted at weight-load time).  The batch is cut into G <= 6
// groups of <= 96 chunks; CTA (g, j) owns tile j for group g, so 24*G <= 144 CTAs are resident, one per SM.
//   - the W_hh tile lives in TENSOR MEMORY for the whole kernel (128 lanes x 384 columns, the A operand of
//     tcgen05.mma), loaded once with tcgen05.st;
//   - a group is cut again into SUB = 3 sub-batches of <= 32 chunks that run the recurrence independently and
//     out of phase: while one sub-batch waits for its h exchange, the tensor pipe and the epilogue warps work on
//     the others (the serial chain of one step -- exchange through L2, TMA, MMA, cell update -- is ~3x longer
//     than its tensor-pipe time);
//   - per step and sub-batch the CTA loads h_{t-1} (32 x 768) with 12 TMA boxes (128B swizzle; the B operand),
//     issues 48 tcgen05.mma (M=128 gate rows, N=32 chunks, K=16) into a 32-column fp32 accumulator in TMEM, adds
//     the hoisted input projection G[t] (TMA-prefetched one step ahead), applies the cell update with the fp32
//     cell state held in registers, and writes its 32-unit slice of h_t to HBM;
//   - the 24 CTAs of a group exchange h_t through L2: release-add on a per-(group, sub-batch) counter,
//     acquire-poll by the TMA producer, proxy fence, next TMA load.
//
// Warp roles (512 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..15 epilogue,
// four per sub-batch (phase 1: TMEM -> fp32 scratch in smem, re-using the dead h buffer; phase 2: one thread per
// (unit, chunk) cell with all four gates: lanes = 32 units, 8 chunks per thread).
#include <stdlib.h>

#include "xb_common.cuh"
#include "xb_ptx.cuh"
#include "xb_gemm.cuh"

using namespace xbptx;

namespace {

constexpr int SUB = 3;                     // sub-batches per group
constexpr int NS = 32;                     // chunks (MMA N) per sub-batch
constexpr int NB = SUB * NS;               // chunks per group
constexpr int TILES = 24;                  // gate tiles per group
constexpr int KCH = XB_FEATURES / 64;      // 12 K chunks of 64
constexpr int H_CHUNK_BYTES = NS * 64 * 2; // 4096
constexpr int H_BYTES = KCH * H_CHUNK_BYTES;        // 49152 per sub-batch
constexpr int G_BYTES = NS * 128 * 2;               // 8192 per buffer
constexpr int P_STRIDE = NS + 1;                    // fp32 scratch row stride (bank-conflict free)
constexpr int HS_OFFSET = ((128 * P_STRIDE * 4 + 127) / 128) * 128;
constexpr int NBAR = 8;                             // mbarriers per sub-batch
constexpr int SMEM_BYTES = SUB * (H_BYTES + 2 * G_BYTES) + 1024 + 1024;
constexpr int THREADS = 128 + SUB * 128;
constexpr int D_COL = 384;                          // accumulators start after the 384 columns of W_hh
static_assert(HS_OFFSET + NS * 64 <= H_BYTES, "epilogue scratch must fit in the h buffer");
static_assert(D_COL + SUB * NS <= 512, "tensor memory columns");

struct PLParams {
    int T, N, reverse;
    int batch0, nbatch, G;      // batch rows [batch0, batch0 + nbatch) are split into G groups
    const uint16_t *w_hh;       // (3072, 768) 16-bit, tile-permuted rows (gate-major within a tile)
    uint16_t *y;                // (T, N, 768) 16-bit output = hidden states
    int *counters;              // (G * SUB) zeroed before launch
    int use3d;                  // tmY is the 3-D {k, chunk, k-block} view (one TMA per sub-batch step)
    long long *dbg;             // optional timeline (XB_LSTM_DEBUG=1): clock64 stamps of CTA 0, steps 64..71
};

#define DBG(ev) do { if (p.dbg && blockIdx.x == 0 && s >= 64 && s < 72) p.dbg[(s - 64) * 16 + (ev)] = clock64(); } while (0)

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcpf(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float clampf(float x, float lim) { return fminf(fmaxf(x, -lim), lim); }

// c' = sig(f) c + sig(i) tanh(g);  h = sig(o) tanh(c').  Seven MUFU ops per cell: the four gate activations share
// one reciprocal (1/(d_i d_f d_g d_o) times the complementary products); inputs are clamped so the product of the
// four denominators stays far from fp32 overflow (sigmoid(-15) = 3e-7, tanh(7.5) = 1 - 6e-7).
__device__ __forceinline__ float lstm_cell(float pi, float pf, float pg, float po, float &c) {
    constexpr float L2E = 1.4426950408889634f;
    const float ei = ex2f(-L2E * clampf(pi, 15.f)), ef = ex2f(-L2E * clampf(pf, 15.f));
    const float eg = ex2f(-2.f * L2E * clampf(pg, 7.5f)), eo = ex2f(-L2E * clampf(po, 15.f));
    const float di = 1.f + ei, df = 1.f + ef, dg = 1.f + eg, dq = 1.f + eo;
    const float p1 = di * df, p2 = dg * dq;
    const float r = rcpf(p1 * p2);
    const float si = r * df * p2, sf = r * di * p2, so = r * p1 * dg, tg = (1.f - eg) * (r * p1 * dq);
    const float cn = sf * c + si * tg;
    c = cn;
    const float ec = ex2f(-2.f * L2E * clampf(cn, 7.5f));
    return so * (1.f - ec) * rcpf(1.f + ec);
}

template <bool BF16>
__global__ void __launch_bounds__(THREADS, 1)
lstm_persistent_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmG, const PLParams p) {
    using X = xb16<BF16>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *hbuf = smem;                                   // [SUB][H_BYTES]
    uint8_t *gbuf = smem + SUB * H_BYTES;                   // [SUB][2][G_BYTES]
    uint64_t *bars = reinterpret_cast<uint64_t *>(gbuf + SUB * 2 * G_BYTES);
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(bars + SUB * NBAR);
    // per sub-batch: h_full, d_full, d_empty, g_full[2], g_empty[2]
    auto h_full = [&](int sub) { return bars + sub * NBAR; };
    auto d_full = [&](int sub) { return bars + sub * NBAR + 1; };
    auto d_empty = [&](int sub) { return bars + sub * NBAR + 2; };
    auto g_full = [&](int sub, int q) { return bars + sub * NBAR + 3 + q; };
    auto g_empty = [&](int sub, int q) { return bars + sub * NBAR + 5 + q; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x / TILES, j = blockIdx.x % TILES;
    const int T = p.T, N = p.N;
    const int b0 = p.batch0 + (int)(((long long)g * p.nbatch) / p.G);
    const int b1 = p.batch0 + (int)(((long long)(g + 1) * p.nbatch) / p.G);
    const int count = b1 - b0;               // <= NB valid chunks in this group
    auto sub_row0 = [&](int sub) { return b0 + (sub * count) / SUB; };

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmY);
        prefetch_tmap(&tmG);
    }
    if (warp == 1 && lane == 0) {
        for (int sub = 0; sub < SUB; sub++) {
            mbar_init(h_full(sub), 1);
            mbar_init(d_full(sub), 1);
            mbar_init(d_empty(sub), 4);
            for (int q = 0; q < 2; q++) {
                mbar_init(g_full(sub, q), 1);
                mbar_init(g_empty(sub, q), 4);
            }
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
            for (int sub = 0; sub < SUB; sub++) {
                mbar_expect_tx(g_full(sub, 0), G_BYTES);
                tma_load_2d(gbuf + (sub * 2) * G_BYTES, &tmG, g_full(sub, 0), j * 128, t0 * N + sub_row0(sub));
            }
            for (int s = 0; s < T; s++) {
                const int t = p.reverse ? T - 1 - s : s;
                for (int sub = 0; sub < SUB; sub++) {
                    const int row0 = sub_row0(sub);
                    if (s > 0) {
                        const int tp = p.reverse ? t + 1 : t - 1;
                        const int need = TILES * s;
                        if (sub == 0) DBG(0);
                        while (ld_acquire_gpu(p.counters + g * SUB + sub) < need) {
                        }
                        if (sub == 0) DBG(1);
                        fence_proxy_async_global();
                        // h_{t-1} of the sub-batch: ONE 3-D box {64 k, 32 chunks, 12 k-blocks} lands as 12 swizzled
                        // [32 x 128 B] K-major tiles (per-box TMA latency, not bytes, dominated with 12 boxes)
                        mbar_expect_tx(h_full(sub), H_BYTES);
                        if (p.use3d) {
                            tma_load_3d(hbuf + sub * H_BYTES, &tmY, h_full(sub), 0, tp * N + row0, 0);
                        } else {
#pragma unroll 1
                            for (int kc = 0; kc < KCH; kc++)
                                tma_load_2d(hbuf + sub * H_BYTES + kc * H_CHUNK_BYTES, &tmY, h_full(sub), kc * 64, tp * N + row0);
                        }
                        if (sub == 0) DBG(2);
                    }
                    if (s + 1 < T) {      // input projection of the next step, one step ahead
                        const int sn = s + 1, q = sn & 1, u = sn >> 1;
                        const int tn = p.reverse ? T - 1 - sn : sn;
                        if (u >= 1) mbar_wait(g_empty(sub, q), (u - 1) & 1);
                        mbar_expect_tx(g_full(sub, q), G_BYTES);
                        tma_load_2d(gbuf + (sub * 2 + q) * G_BYTES, &tmG, g_full(sub, q), j * 128, tn * N + row0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(BF16 ? 1u : 0u, 128, NS);
        const uint32_t hb = smem_u32(hbuf);
        for (int s = 1; s < T; s++) {
            for (int sub = 0; sub < SUB; sub++) {
                mbar_wait(d_empty(sub), (s - 1) & 1);     // epilogue of step s-1 has drained this accumulator
                tc_fence_after();
                if (lane == 0 && sub == 0) DBG(3);
                mbar_wait(h_full(sub), (s - 1) & 1);
                tc_fence_after();
                if (lane == 0 && sub == 0) DBG(4);
                if (elect_one()) {
                    const uint64_t bdesc0 = umma_desc_sw128(hb + sub * H_BYTES);
#pragma unroll 4
                    for (int kk = 0; kk < KCH * 4; kk++) {
                        const int kc = kk >> 2, k = kk & 3;
                        mma_f16_ts(tmem_base + D_COL + sub * NS, tmem_base + kk * 8,
                                   bdesc0 + (uint64_t)((kc * H_CHUNK_BYTES + k * 32) >> 4), idesc, kk != 0);
                    }
                    mma_commit(d_full(sub));
                }
                __syncwarp();
                if (lane == 0 && sub == 0) DBG(5);
                if (lane == 0 && sub == 0) DBG(6);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (128 threads per sub-batch)
        const int sub = (warp - 4) >> 2, q = warp & 3;
        const int et = threadIdx.x - 128 - sub * 128;        // 0..127 within the sub-batch
        const int r = q * 32 + lane;                         // gate row of the tile == TMEM lane
        const int row0 = sub_row0(sub);
        const int cnt = b0 + ((sub + 1) * count) / SUB - row0;   // valid chunks of this sub-batch (<= NS)
        float *P = reinterpret_cast<float *>(hbuf + sub * H_BYTES);
        uint16_t *HS = reinterpret_cast<uint16_t *>(hbuf + sub * H_BYTES + HS_OFFSET);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + D_COL + sub * NS;
        const int bar_id = 1 + sub;
        float cst[8];
#pragma unroll
        for (int i = 0; i < 8; i++) cst[i] = 0.0f;

        for (int s = 0; s < T; s++) {
            const int t = p.reverse ? T - 1 - s : s;
            const uint16_t *Gs = reinterpret_cast<const uint16_t *>(gbuf + (sub * 2 + (s & 1)) * G_BYTES);
            if (s > 0) {
                mbar_wait(d_full(sub), (s - 1) & 1);
                tc_fence_after();
            }
            if (et == 0 && sub == 0) DBG(7);
            // phase 1: accumulator -> fp32 scratch P[gate row][chunk]
#pragma unroll
            for (int cc = 0; cc < NS / 16; cc++) {
                uint32_t acc[16];
                if (s > 0) {
                    tmem_ld_32x32b_x16(taddr + cc * 16, acc);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) acc[i] = 0u;
                }
#pragma unroll
                for (int i = 0; i < 16; i++) P[r * P_STRIDE + cc * 16 + i] = __uint_as_float(acc[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d_empty(sub));
            mbar_wait(g_full(sub, s & 1), (s >> 1) & 1);
            named_bar_sync(bar_id, 128);
            if (et == 0 && sub == 0) DBG(8);
            // phase 2: cells (unit = lane, chunk b = q + 4m); G row b holds [unit][i,f,g,o] 16-bit
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const int b = q + 4 * m;
                const uint2 graw = *reinterpret_cast<const uint2 *>(Gs + b * 128 + lane * 4);
                const float2 g01 = X::unpack(graw.x), g23 = X::unpack(graw.y);
                const float pi = P[(lane)*P_STRIDE + b] + g01.x;
                const float pf = P[(32 + lane) * P_STRIDE + b] + g01.y;
                const float pg = P[(64 + lane) * P_STRIDE + b] + g23.x;
                const float po = P[(96 + lane) * P_STRIDE + b] + g23.y;
                const float hn = lstm_cell(pi, pf, pg, po, cst[m]);
                typename X::T hv = X::from(hn);
                HS[b * 32 + lane] = *reinterpret_cast<uint16_t *>(&hv);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(g_empty(sub, s & 1));
            named_bar_sync(bar_id, 128);
            if (et == 0 && sub == 0) DBG(9);
            // coalesced store of the (chunks x 32 units) slice: 64 bytes per chunk row
            {
                const int b = et >> 2, part = et & 3;
                if (b < cnt)
                    *reinterpret_cast<uint4 *>(p.y + ((size_t)t * N + row0 + b) * XB_FEATURES + j * 32 + part * 8) =
                        reinterpret_cast<const uint4 *>(HS)[et];
            }
            fence_proxy_async();          // scratch (generic proxy) before the next TMA write into the same smem
            named_bar_sync(bar_id, 128);
            if (et == 0) {
                if (sub == 0) DBG(10);
                __threadfence();
                red_release_gpu_add(p.counters + g * SUB + sub, 1);
                if (sub == 0) DBG(11);
            }
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
    XB_CUDA(h, cudaMemcpy(out_host, h->lstm_counters + 32, 8 * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
    return XB_OK;
}

// Recurrent part of one LSTM layer.  h->gates must already hold the input projection (T*N, 3072) with the
// columns of every 128-wide tile ordered [unit][i,f,g,o] (weight repack mode 3 in xb_api.cu).
int xb_lstm_recurrence_persistent(xb_handle *h, int layer, void *y_tnc, int T, int N, int reverse, cudaStream_t s) {
    if (!h->lstm_counters) {
        void *q = nullptr;
        XB_CUDA(h, cudaMalloc(&q, 32 * sizeof(int) + 8 * 16 * sizeof(long long)));
        h->lstm_counters = reinterpret_cast<int *>(q);
    }
    CUtensorMap tmY, tmG;
    int use3d = getenv("XB_LSTM_NO3D") ? 0 : 1;
    if (use3d && xb_make_tmap_hview(h, &tmY, y_tnc, (uint64_t)T * N, NS) != XB_OK) use3d = 0;
    if (!use3d)
        if (int rc = xb_make_tmap_2d_box(h, &tmY, y_tnc, (uint64_t)T * N, XB_FEATURES, XB_FEATURES, 64, NS, 1)) return rc;
    if (int rc = xb_make_tmap_2d_box(h, &tmG, h->gates, (uint64_t)T * N, XB_GATES, XB_GATES, 128, NS, 0)) return rc;
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
        p.use3d = use3d;
        p.dbg = getenv("XB_LSTM_DEBUG") ? reinterpret_cast<long long *>(h->lstm_counters + 32) : nullptr;
        XB_CUDA(h, cudaMemsetAsync(h->lstm_counters, 0, 32 * sizeof(int), s));
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
