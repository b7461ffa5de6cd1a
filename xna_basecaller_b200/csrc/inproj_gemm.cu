// LSTM input projection G = x W_ih^T + (b_ih + b_hh) for all T*N rows at once (the hoisted half of
// torch.nn.LSTM, bonito/nn.py:189-193), as an A-STATIONARY persistent tcgen05 GEMM.
//
// Why not the generic tile kernel (gemm_tc.cu).  A B200 SM ingests ~64 B/clk from L2 while its tensor core
// retires 8192 dense 16-bit FLOP/clk.  A 128x128 output tile with K = 768 needs 2 x 196 KB of operands for 3072
// MMA cycles: 6144 cycles of ingest, so that kernel runs ingest-bound at <= 50% of the tensor peak (measured
// 37%).  Here the 128 x 768 block of x rows is loaded ONCE per M tile into TENSOR MEMORY (384 of the 512
// columns; the A operand of tcgen05.mma) and only W_ih streams through shared memory: 96 KB per 64-column
// N tile for 1536 MMA cycles -- ingest and tensor time balance, and x is read from HBM exactly once.
//
// One CTA per SM, persistent over M tiles (tile = blockIdx.x, += gridDim.x).  Per M tile: 48 N tiles of 64
// gate columns; two 64-column fp32 accumulators in the remaining 128 TMEM columns, so the epilogue of N tile i
// (TMEM -> +bias -> 16-bit -> staged rows -> 128-byte row segments of G) overlaps the MMAs of N tile i+1.
// W_ih arrives as 3-D TMA boxes of four [64 rows x 64 K] 128B-swizzled blocks (32 KB) through a 4-stage ring
// (128 KB in flight); one mbarrier wait and one commit per 16 MMAs keep the single issuing thread ahead of
// the tensor pipe (with one wait per 4 MMAs the issue loop, not the pipe, set the pace).
//
// Warp roles (256 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..7 load the
// x block into TMEM (thread = row) and run the epilogues (warp % 4 = TMEM lane quarter).
#include <stdlib.h>

#include "xb_common.cuh"
#include "xb_ptx.cuh"
#include "xb_gemm.cuh"

using namespace xbptx;

namespace {

constexpr int BM = 128, BNI = 64, BK = 64;
constexpr int KB = XB_FEATURES / BK;                 // 12 K blocks
constexpr int NT = XB_GATES / BNI;                   // 48 N tiles
constexpr int KPS = 4;                               // K blocks per pipeline stage (one 3-D TMA box, 16 MMAs per wait)
constexpr int SPT = KB / KPS;                        // stages per N tile
constexpr int STAGES = 6;
constexpr int KBLOCK_BYTES = BNI * BK * 2;           // 8 KB: one [64 rows x 128 B] swizzle atom column
constexpr int STAGE_BYTES = KPS * KBLOCK_BYTES;      // 32 KB
constexpr int A_COLS = XB_FEATURES / 2;              // 384 TMEM columns hold the x block
constexpr int ROW_PITCH = BNI * 2 + 16;              // staged epilogue row pitch (bytes)
constexpr int STG_BYTES = 32 * ROW_PITCH;            // per epilogue warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 4 * STG_BYTES + XB_GATES * 4 + 1024 /*align*/ + 512 /*barriers*/;

struct IPParams {
    const uint16_t *x;        // (M, 768) 16-bit
    const float *bias;        // (3072) fp32, same column order as the rows of W_ih
    uint16_t *out;            // (M, 3072) 16-bit
    int M;
    int no_prefetch;
    long long *dbg;           // optional (XB_INPROJ_DEBUG): stall cycles of the MMA thread of CTA 0: [acc_empty, full, issue, a_ready]
};

template <bool BF16>
__global__ void __launch_bounds__(256, 1)
inproj_kernel(const __grid_constant__ CUtensorMap tmW, const IPParams p) {
    using X = xb16<BF16>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *stgbuf = smem + STAGES * STAGE_BYTES;
    float *sbias = reinterpret_cast<float *>(stgbuf + 4 * STG_BYTES);
    uint64_t *full = reinterpret_cast<uint64_t *>(sbias + XB_GATES);
    uint64_t *empty = full + STAGES;
    uint64_t *acc_full = empty + STAGES;      // [2]
    uint64_t *acc_empty = acc_full + 2;       // [2]
    uint64_t *a_ready = acc_empty + 2;        // [SPT]: K blocks ks*KPS .. ks*KPS+KPS-1 of the x block are in tensor memory
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(a_ready + SPT);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int ntiles = (p.M + BM - 1) / BM;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmW);
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);
        }
        for (int ks = 0; ks < SPT; ks++) mbar_init(&a_ready[ks], 4);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_holder, 512);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < XB_GATES; i += 256) sbias[i] = p.bias[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer: W_ih boxes, all tiles
        if (elect_one()) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int nt = 0; nt < NT; nt++)
                    for (int ks = 0; ks < SPT; ks++, it++) {
                        const int s = it % STAGES;
                        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                        mbar_expect_tx(&full[s], STAGE_BYTES);
                        tma_load_3d(smem + s * STAGE_BYTES, &tmW, &full[s], 0, nt * BNI, ks * KPS);
                    }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(BF16 ? 1u : 0u, BM, BNI);
        uint32_t it = 0, nit = 0, tit = 0;
        long long st_acc = 0, st_full = 0, st_issue = 0, st_a = 0, tt;
        const bool dbg = p.dbg && blockIdx.x == 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tit++) {
            for (int nt = 0; nt < NT; nt++, nit++) {
                const int buf = nit & 1;
                tt = clock64();
                mbar_wait(&acc_empty[buf], ((nit >> 1) & 1) ^ 1);
                tc_fence_after();
                st_acc += clock64() - tt;
                const uint32_t d = tmem_base + A_COLS + buf * BNI;
                for (int ks = 0; ks < SPT; ks++, it++) {
                    const int s = it % STAGES;
                    tt = clock64();
                    mbar_wait(&full[s], (it / STAGES) & 1);
                    tc_fence_after();
                    st_full += clock64() - tt;
                    if (nt == 0) {                             // first N tile: the x block arrives K-block group by group
                        tt = clock64();
                        mbar_wait(&a_ready[ks], tit & 1);
                        tc_fence_after();
                        st_a += clock64() - tt;
                    }
                    tt = clock64();
                    if (elect_one()) {
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + s * STAGE_BYTES));
                        const uint32_t a0 = tmem_base + ks * KPS * (BK / 2);
#pragma unroll
                        for (int kk = 0; kk < KPS * 4; kk++) {
                            const int kc = kk >> 2, k = kk & 3;
                            mma_f16_ts(d, a0 + kk * 8, bdesc + (uint64_t)((kc * KBLOCK_BYTES + k * 32) >> 4), idesc,
                                       (ks | kk) != 0);
                        }
                        mma_commit(&empty[s]);
                        if (ks == SPT - 1) mma_commit(&acc_full[buf]);
                    }
                    __syncwarp();
                    st_issue += clock64() - tt;
                }
            }
        }
        if (dbg && lane == 0) { p.dbg[0] = st_acc; p.dbg[1] = st_full; p.dbg[2] = st_issue; p.dbg[3] = st_a; p.dbg[4] = tit; }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ x block loader + epilogue
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        uint8_t *stg = stgbuf + q * STG_BYTES;
        uint32_t nit = 0;
        bool pending = false;                                  // epilogue of the previous tile's last N tile
        int pend_m0 = 0;

        auto epilogue = [&](int m0, int nt, uint32_t ni) {
            const int buf = ni & 1;
            uint32_t acc[BNI];
            {
                uint32_t a0[32], a1[32];
                tmem_ld_32x32b_x32(lane_base + A_COLS + buf * BNI, a0);
                tmem_ld_32x32b_x32(lane_base + A_COLS + buf * BNI + 32, a1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i++) { acc[i] = a0[i]; acc[32 + i] = a1[i]; }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);       // the MMAs of N tile ni+2 may overwrite it now
            const float *bs = sbias + nt * BNI;
            uint4 *d = reinterpret_cast<uint4 *>(stg + lane * ROW_PITCH);
#pragma unroll
            for (int j = 0; j < BNI / 8; j++) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; e++)
                    pk[e] = X::pack(__uint_as_float(acc[8 * j + 2 * e]) + bs[8 * j + 2 * e],
                                    __uint_as_float(acc[8 * j + 2 * e + 1]) + bs[8 * j + 2 * e + 1]);
                d[j] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            __syncwarp();
            const int sub = lane >> 3, l8 = lane & 7;          // 8 lanes x 16 B = one 128-byte row segment
#pragma unroll
            for (int rr = 0; rr < 32; rr += 4) {
                const int mm = m0 + q * 32 + rr + sub;
                if (mm < p.M)
                    reinterpret_cast<uint4 *>(p.out + (size_t)mm * XB_GATES + nt * BNI)[l8] =
                        *reinterpret_cast<const uint4 *>(stg + (rr + sub) * ROW_PITCH + l8 * 16);
            }
            __syncwarp();
        };

        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int m0 = tile * BM;
            if (pending) {                                     // all MMAs of the previous tile have completed:
                mbar_wait(&acc_full[(nit - 1) & 1], ((nit - 1) >> 1) & 1);   // its x block may be replaced
                tc_fence_after();
            }
            const long long tl0 = clock64();
            {   // x block -> tensor memory (lane = row, column c holds elements k = 2c, 2c+1).  The tensor pipe idles
                // during this, so it has to be quick: a direct row-per-thread read costs one L1 wavefront per lane
                // (12 k wavefronts per tile, ~20 k cycles).  Instead a warp reads its 32 rows coalesced, 128 bytes
                // (= one K block) of four rows per instruction, four K blocks in flight, transposes through its
                // staging rows and hands each thread its own row for tcgen05.st.
                const int sub = lane >> 3, l8 = lane & 7;
                const int mrow0 = m0 + q * 32;
                constexpr int NBK = KPS;                                 // K blocks in flight per round = one MMA stage
#pragma unroll 1
                for (int kb0 = 0; kb0 < KB; kb0 += NBK) {
                    uint4 ld[NBK][8];
#pragma unroll
                    for (int b = 0; b < NBK; b++)
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int mm = mrow0 + 4 * i + sub;
                            ld[b][i] = (mm < p.M) ? __ldg(reinterpret_cast<const uint4 *>(p.x + (size_t)mm * XB_FEATURES + (kb0 + b) * BK) + l8)
                                                  : make_uint4(0, 0, 0, 0);
                        }
#pragma unroll
                    for (int b = 0; b < NBK; b++) {
#pragma unroll
                        for (int i = 0; i < 8; i++)
                            *reinterpret_cast<uint4 *>(stg + (4 * i + sub) * ROW_PITCH + l8 * 16) = ld[b][i];
                        __syncwarp();
                        uint32_t v[32];
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const uint4 t4 = *reinterpret_cast<const uint4 *>(stg + lane * ROW_PITCH + i * 16);
                            v[4 * i] = t4.x; v[4 * i + 1] = t4.y; v[4 * i + 2] = t4.z; v[4 * i + 3] = t4.w;
                        }
                        __syncwarp();
                        tmem_st_32x32b_x32(lane_base + (kb0 + b) * (BK / 2), v);
                    }
                    tmem_st_wait();                            // the MMAs of the first N tile start on this group at once
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_ready[kb0 / KPS]);
                }
            }
            if (p.dbg && blockIdx.x == 0 && warp == 4 && lane == 0) p.dbg[5] += clock64() - tl0;
            if (pending) epilogue(pend_m0, NT - 1, nit - 1);
            for (int nt = 0; nt < NT - 1; nt++, nit++) {
                if (nt == NT - 8 && !p.no_prefetch) {      // next tile's x rows into L2 shortly before they are needed (G streams through L2)
                    const int mn = m0 + (int)gridDim.x * BM + r;
                    if (mn < p.M) {
                        const char *nx = reinterpret_cast<const char *>(p.x + (size_t)mn * XB_FEATURES);
#pragma unroll
                        for (int i = 0; i < 12; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + i * 128));
                    }
                }
                mbar_wait(&acc_full[nit & 1], (nit >> 1) & 1);
                tc_fence_after();
                epilogue(m0, nt, nit);
            }
            nit++;                                             // the last N tile is finished after the next x load
            pending = true;
            pend_m0 = m0;
        }
        if (pending) {
            mbar_wait(&acc_full[(nit - 1) & 1], ((nit - 1) >> 1) & 1);
            tc_fence_after();
            epilogue(pend_m0, NT - 1, nit - 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

// gates (M, 3072) 16-bit = x (M, 768) . w_ih^T + bias
int xb_inproj_launch(xb_handle *h, const void *x, const void *w_ih, const float *bias, void *gates, int M, cudaStream_t s) {
    CUtensorMap tmW;
    if (int rc = xb_make_tmap_hview(h, &tmW, w_ih, XB_GATES, BNI, KPS)) return rc;
    IPParams p;
    p.x = reinterpret_cast<const uint16_t *>(x);
    p.bias = bias;
    p.out = reinterpret_cast<uint16_t *>(gates);
    p.M = M;
    p.dbg = nullptr;
    p.no_prefetch = getenv("XB_INPROJ_NOPF") ? 1 : 0;
    if (getenv("XB_INPROJ_DEBUG")) {
        static long long *d = nullptr;
        if (!d) cudaMalloc(&d, 64);
        p.dbg = d;
    }
    const int ntiles = (M + BM - 1) / BM;
    const int grid = ntiles < h->num_sms ? ntiles : h->num_sms;
    static bool configured[2][64] = {};   // per device: function attributes live in the device's context
    const int dv = h->device & 63;
    if (h->bf16) {
        if (!configured[1][dv]) {
            XB_CUDA(h, cudaFuncSetAttribute(inproj_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            configured[1][dv] = true;
        }
        inproj_kernel<true><<<grid, 256, SMEM_BYTES, s>>>(tmW, p);
    } else {
        if (!configured[0][dv]) {
            XB_CUDA(h, cudaFuncSetAttribute(inproj_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            configured[0][dv] = true;
        }
        inproj_kernel<false><<<grid, 256, SMEM_BYTES, s>>>(tmW, p);
    }
    XB_LAUNCH_CHECK(h);
    if (p.dbg) {
        long long v[6];
        cudaDeviceSynchronize();
        cudaMemcpy(v, p.dbg, sizeof v, cudaMemcpyDeviceToHost);
        fprintf(stderr, "inproj MMA thread of CTA 0: %lld tiles; stall cycles: acc_empty %lld, full %lld, a_ready %lld; issue %lld; x-block load (warp 4) %lld\n", v[4], v[0], v[1], v[3], v[2], v[5]);
        cudaMemset(p.dbg, 0, 64);
    }
    return XB_OK;
}
