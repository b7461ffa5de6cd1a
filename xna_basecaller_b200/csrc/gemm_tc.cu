// tcgen05 + TMA GEMM for sm_100a with fused epilogues:  D(M,N) = A(M,K) . B(N,K)^T, 16-bit K-major operands,
// fp32 accumulation in tensor memory.  One 128x128 output tile per CTA, 64-wide K blocks through a TMA ->
// mbarrier -> tcgen05.mma ring, two CTAs per SM so that one CTA's epilogue overlaps the other's main loop.
//
// Warp roles (256 threads): warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4..7 = epilogue (warp q reads TMEM lanes 32q..32q+31 = tile rows).
//
// Epilogues
//   EPI_F32    plain fp32 store (self test)
//   EPI_CONV3  third Convolution of the stem as an implicit GEMM over im2col rows: + bias, swish, 16-bit store
//              in (T, N, 768) order, which absorbs nn.Permute([2,0,1])       (bonito/nn.py:57-68,156-167); the product
//              runs conv3_gemm.cu instead, this form remains in -DXB_EXPERIMENTS builds
//   EPI_INPROJ LSTM input projection x_t W_ih^T + (b_ih + b_hh) for all t at once, 16-bit store (nn.py:189-193)
//   EPI_HEAD   LinearCRFEncoder: scale * tanh(x W^T + b), blank score inserted in front of every group of
//              n_base columns -> fp32 (T, N, C*NZ)                            (bonito/nn.py:112-133)
//   EPI_LSTM   one LSTM time step: gates = acc (h_{t-1} W_hh^T) + G[t]; c, h update in fp32; h stored 16-bit.
//              Columns of a tile are [i | f | g | o] x 32 hidden units (weights are row-permuted at load time).
#include "xb_common.cuh"
#include "xb_ptx.cuh"
#include "xb_gemm.cuh"
#include "xb_head_epilogue.cuh"

using namespace xbptx;

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int TILE_A_BYTES = BM * BK * 2, TILE_B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = TILE_A_BYTES + TILE_B_BYTES;
// Pipeline depth.  Three stages (96 KB) leave room for two CTAs per SM, whose epilogues and main loops overlap.  The
// one-launch-per-step recurrence GEMM of the cross-check builds (a 24-tile grid: one CTA per SM, latency-bound on the TMA
// round trip) runs six stages instead: 192 KB in flight per SM.
template <int EPI> struct Pipe {
    static constexpr int STAGES = (EPI == EPI_LSTM) ? 6 : 3;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool BF16, int EPI>
__global__ void __launch_bounds__(256, (EPI == EPI_LSTM) ? 1 : 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    constexpr int STAGES = Pipe<EPI>::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *tmem_full = empty + STAGES;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    // K need not be a multiple of 64: the tensor maps are built over exactly K columns and TMA zero-fills the rest
    // split-K (EPI_F32 only, gridDim.z slices): slice z covers K blocks [kb0, kb0 + kblocks) and adds its partial tile to the
    // zeroed output -- for the weight-gradient contractions whose output has fewer tiles than the GPU has SMs
    const int kb_all = (p.K + BK - 1) / BK;
    const int kb_per = (kb_all + gridDim.z - 1) / gridDim.z;
    const int kb0 = blockIdx.z * kb_per;
    const int kblocks = (EPI == EPI_LSTM && p.first) ? 0 : max(0, min(kb_all, kb0 + kb_per) - kb0);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_holder, BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        if (elect_one()) {
            for (int kb = 0; kb < kblocks; kb++) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], STAGE_BYTES);
                uint8_t *sa = smem + s * STAGE_BYTES;
                tma_load_2d(sa, &tmA, &full[s], (kb0 + kb) * BK, m0 + p.a_row_offset);
                tma_load_2d(sa + TILE_A_BYTES, &tmB, &full[s], (kb0 + kb) * BK, n0);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_f16(BF16 ? 1u : 0u, BM, BN);
        for (int kb = 0; kb < kblocks; kb++) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
                const uint32_t b_addr = a_addr + TILE_A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 16; k++) {
                    mma_f16_ss(tmem_base, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                               (kb | k) != 0);
                }
                mma_commit(&empty[s]);
                if (kb == kblocks - 1) mma_commit(tmem_full);
            }
            __syncwarp();
        }
    }
    // Epilogue: warp w reads TMEM lanes 32*(w%4)..+31 (= tile rows).  The fused LSTM step keeps its four dedicated
    // warps; every other epilogue runs on all eight warps (the producer / MMA warps are idle once the main loop has
    // been issued): warps 0-3 take tile columns 0..63, warps 4-7 columns 64..127.  (EPI_F32, the self test, also keeps four.)
    __syncwarp();
    if ((EPI == EPI_LSTM || EPI == EPI_F32 || EPI == EPI_BF16OUT || EPI == EPI_CONV3_BWD) ? warp >= 4 : true) {
        const int q = warp & 3;
        const int r = q * 32 + lane;             // tile row == TMEM lane
        const int m = m0 + r;
        const bool row_ok = m < p.M;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        if (kblocks > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
        }
        using X = xb16<BF16>;
        if constexpr (EPI == EPI_LSTM) {
            // row = batch element; columns [g*32 + u], g in (i,f,g,o), u = unit within the 32-unit slice
            const int unit0 = blockIdx.x * 32;
            const size_t grow = ((size_t)p.t_cur * p.NB + m) * XB_GATES + n0;
            const uint16_t *G = reinterpret_cast<const uint16_t *>(p.gates) + grow;
            float *cst = p.cstate + (size_t)m * XB_FEATURES + unit0;
            uint16_t *hout = reinterpret_cast<uint16_t *>(p.out) + ((size_t)p.t_cur * p.NB + m) * XB_FEATURES + unit0;
#pragma unroll 1
            for (int ub = 0; ub < 32; ub += 8) {
                uint32_t acc[4][8];
                if (kblocks > 0) {
#pragma unroll
                    for (int g = 0; g < 4; g++) tmem_ld_32x32b_x8(taddr + g * 32 + ub, acc[g]);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int g = 0; g < 4; g++)
#pragma unroll
                        for (int j = 0; j < 8; j++) acc[g][j] = 0u;
                }
                if (row_ok) {
                    float gv[4][8];
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        uint4 raw = *reinterpret_cast<const uint4 *>(G + g * 32 + ub);
                        const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            float2 f = X::unpack(rw[j]);
                            gv[g][2 * j] = __uint_as_float(acc[g][2 * j]) + f.x;
                            gv[g][2 * j + 1] = __uint_as_float(acc[g][2 * j + 1]) + f.y;
                        }
                    }
                    float cprev[8];
                    if (p.first) {
#pragma unroll
                        for (int j = 0; j < 8; j++) cprev[j] = 0.0f;
                    } else {
                        float4 c0 = *reinterpret_cast<const float4 *>(cst + ub);
                        float4 c1 = *reinterpret_cast<const float4 *>(cst + ub + 4);
                        cprev[0] = c0.x; cprev[1] = c0.y; cprev[2] = c0.z; cprev[3] = c0.w;
                        cprev[4] = c1.x; cprev[5] = c1.y; cprev[6] = c1.z; cprev[7] = c1.w;
                    }
                    float cn[8], hn[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        float ig = fast_sigmoid(gv[0][j]), fg = fast_sigmoid(gv[1][j]);
                        float gg = fast_tanh(gv[2][j]), og = fast_sigmoid(gv[3][j]);
                        cn[j] = fg * cprev[j] + ig * gg;
                        hn[j] = og * fast_tanh(cn[j]);
                    }
                    *reinterpret_cast<float4 *>(cst + ub) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                    *reinterpret_cast<float4 *>(cst + ub + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
                    uint4 hv;
                    hv.x = X::pack(hn[0], hn[1]); hv.y = X::pack(hn[2], hn[3]);
                    hv.z = X::pack(hn[4], hn[5]); hv.w = X::pack(hn[6], hn[7]);
                    *reinterpret_cast<uint4 *>(hout + ub) = hv;
                }
            }
        } else if constexpr (EPI == EPI_BF16OUT) {
            // plain bf16 store of the accumulator (gradients flowing to the layer below)
#pragma unroll 1
            for (int cb = 0; cb < BN; cb += 32) {
                uint32_t acc[32];
                tmem_ld_32x32b_x32(taddr + cb, acc);
                tmem_ld_wait();
                const int nb = n0 + cb;
                if (row_ok) {
                    __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(p.out) + (size_t)m * p.ldo + nb;
                    if (nb + 32 <= p.N && (p.ldo & 7) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 raw;
                            __nv_bfloat162 *b2 = reinterpret_cast<__nv_bfloat162 *>(&raw);
#pragma unroll
                            for (int e = 0; e < 4; e++)
                                b2[e] = __floats2bfloat162_rn(__uint_as_float(acc[j + 2 * e]), __uint_as_float(acc[j + 2 * e + 1]));
                            *reinterpret_cast<uint4 *>(o + j) = raw;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (nb + j < p.N) o[j] = __float2bfloat16_rn(__uint_as_float(acc[j]));
                    }
                }
            }
        } else if constexpr (EPI == EPI_CONV3_BWD) {
            // conv3's pre-activation is recomputed (im2col rows x W3^T + b3) instead of stored; with the incoming gradient
            // dy (T, NB, 768) bf16 in time-major order this writes d pre = dy * swish'(pre) in im2col ROW order (chunk, t)
#pragma unroll 1
            for (int cb = 0; cb < BN; cb += 32) {
                uint32_t acc[32];
                tmem_ld_32x32b_x32(taddr + cb, acc);
                tmem_ld_wait();
                const int nb = n0 + cb;
                if (row_ok) {
                    const int b = m / p.T, t = m - b * p.T;
                    const __nv_bfloat16 *dyp = reinterpret_cast<const __nv_bfloat16 *>(p.dy) + ((size_t)t * p.NB + b) * XB_FEATURES + nb;
                    __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(p.out) + (size_t)m * p.ldo + nb;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float pre = __uint_as_float(acc[j]) + __ldg(p.bias + nb + j);
                        const float sg = 1.0f / (1.0f + __expf(-pre));
                        o[j] = __float2bfloat16_rn(__bfloat162float(dyp[j]) * (sg * (1.0f + pre * (1.0f - sg))));
                    }
                }
            }
        } else if constexpr (EPI == EPI_F32) {
#pragma unroll 1
            for (int cb = 0; cb < (kblocks > 0 ? BN : 0); cb += 32) {       // an empty K slice contributes nothing
                uint32_t acc[32];
                tmem_ld_32x32b_x32(taddr + cb, acc);
                tmem_ld_wait();
                const int nb = n0 + cb;
                if (row_ok) {
                    float *o = reinterpret_cast<float *>(p.out) + (size_t)m * p.ldo + nb;
                    if (gridDim.z > 1) {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (nb + j < p.N) atomicAdd(o + j, __uint_as_float(acc[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (nb + j < p.N) o[j] = __uint_as_float(acc[j]);
                    }
                }
            }
        } else {
            // Coalesced epilogues.  The accumulator arrives one tile ROW per thread; a direct store would touch 32
            // different output rows per instruction.  Each epilogue warp therefore stages its 32 x 128 block in
            // the (now idle: every MMA has completed) pipeline buffers and writes whole row segments.
            uint8_t *stg = smem + warp * (STAGES * STAGE_BYTES / 8);
            constexpr int HN = BN / 2;                           // columns per warp
            const int c0 = (warp >> 2) * HN;                     // first tile column of this warp
            if constexpr (EPI == EPI_CONV3 || EPI == EPI_INPROJ) {
                constexpr int RSB = HN * 2 + 16;                 // staged row pitch in bytes (16 B pad: conflict free)
#pragma unroll 1
                for (int cb = c0; cb < c0 + HN; cb += 32) {
                    uint32_t acc[32];
                    tmem_ld_32x32b_x32(taddr + cb, acc);
                    tmem_ld_wait();
                    const int nb = n0 + cb;
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        float v0 = __uint_as_float(acc[2 * j]) + __ldg(p.bias + nb + 2 * j);
                        float v1 = __uint_as_float(acc[2 * j + 1]) + __ldg(p.bias + nb + 2 * j + 1);
                        if constexpr (EPI == EPI_CONV3) {      // swish with ONE MUFU per element: sigmoid(x) = 0.5 tanh(0.5 x) + 0.5
                            v0 = v0 * fmaf(tanh_approx(0.5f * v0), 0.5f, 0.5f);      // (the exp + reciprocal form kept this epilogue,
                            v1 = v1 * fmaf(tanh_approx(0.5f * v1), 0.5f, 0.5f);      // and with it the conv3 GEMM, MUFU-bound)
                        }
                        pk[j] = X::pack(v0, v1);
                    }
                    uint4 *d = reinterpret_cast<uint4 *>(stg + lane * RSB + (cb - c0) * 2);
#pragma unroll
                    for (int j = 0; j < 4; j++) d[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
                __syncwarp();
                // 8 lanes x 16 B = one 128-byte row segment; four rows per instruction
                const int half = lane >> 3, l16 = lane & 7;
#pragma unroll 4
                for (int rr = 0; rr < 32; rr += 4) {
                    const int mm = m0 + q * 32 + rr + half;
                    if (mm < p.M) {
                        size_t orow;
                        if constexpr (EPI == EPI_CONV3) {
                            const int b = mm / p.T, t = mm - b * p.T;        // im2col rows are (chunk, t)
                            orow = (size_t)t * p.NB + b;
                        } else {
                            orow = (size_t)mm;
                        }
                        uint16_t *o = reinterpret_cast<uint16_t *>(p.out) + orow * p.ldo + n0 + c0;
                        reinterpret_cast<uint4 *>(o)[l16] = *reinterpret_cast<const uint4 *>(stg + (rr + half) * RSB + l16 * 16);
                    }
                }
            } else {   // EPI_HEAD
                constexpr int RS = 81;                           // staged row pitch in floats (odd: conflict free)
                float *S = reinterpret_cast<float *>(stg);
                const int col0 = n0 + c0;                        // head columns of this warp: [col0, col_end)
                const int col_end = min(col0 + HN, p.head_rows);
                int seg_start, seg_len;
                xbhead::segment(col0, col_end, p.n_base, p.expand, seg_start, seg_len);
#pragma unroll 1
                for (int cb = c0; cb < c0 + HN; cb += 32) {
                    if (n0 + cb >= col_end) break;
                    uint32_t acc[32];
                    tmem_ld_32x32b_x32(taddr + cb, acc);
                    tmem_ld_wait();
                    xbhead::stage32(acc, n0 + cb, col_end, p.n_base, p.expand, p.bias, p.scale, p.blank, S + lane * RS, seg_start);
                }
                __syncwarp();
                float *obase = reinterpret_cast<float *>(p.out) + seg_start;
#pragma unroll 1
                for (int rr = 0; rr < 32; rr++) {
                    const int mm = m0 + q * 32 + rr;
                    if (mm >= p.M) break;
                    float *o = obase + (size_t)mm * p.ldo;
                    const float *Sr = S + rr * RS;
                    for (int idx = lane; idx < seg_len; idx += 32) o[idx] = Sr[idx];
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// 2-D K-major tensor map: rows x K 16-bit elements, row pitch ld elements, box 64 x 128, 128B swizzle.
int xb_make_tmap_2d(xb_handle *h, CUtensorMap *out, const void *base, uint64_t rows, uint64_t K, uint64_t ld) {
    return xb_make_tmap_2d_box(h, out, base, rows, K, ld, BK, BM, 1);
}

// general form: box_inner x box_rows elements; swizzle128 = 1 -> CU_TENSOR_MAP_SWIZZLE_128B (box_inner must be 64)
int xb_make_tmap_2d_box(xb_handle *h, CUtensorMap *out, const void *base, uint64_t rows, uint64_t K, uint64_t ld,
                        uint32_t box_inner, uint32_t box_rows, int swizzle128) {
    if (!h->encode_tiled) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return xb_fail(h, XB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        h->encode_tiled = fn;
    }
    cuuint64_t dims[2] = {K, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<encode_tiled_fn>(h->encode_tiled)(
        out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16 /* 16-bit payload; the element type only matters for OOB fill */, 2, const_cast<void *>(base), dims,
        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return xb_fail(h, XB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return XB_OK;
}

// rank-n form (n <= 5): dims / box in elements (innermost first), strides in BYTES for dimensions 1 .. n-1, 128B swizzle
// (the innermost box extent must span 128 bytes); elem_bytes 2 (16-bit payload) or 4 (fp32)
int xb_make_tmap_nd(xb_handle *h, CUtensorMap *out, const void *base, int rank, int elem_bytes, const uint64_t *dims,
                    const uint64_t *strides_bytes, const uint32_t *box) {
    if (!h->encode_tiled) {
        CUtensorMap tmp;
        if (int rc = xb_make_tmap_2d(h, &tmp, base, 128, 64, 64)) return rc;        // resolves the entry point
    }
    if (rank < 2 || rank > 5 || (elem_bytes != 2 && elem_bytes != 4)) return xb_fail(h, XB_ERR_ARG, "tensor map rank %d / element size %d", rank, elem_bytes);
    cuuint64_t d[5], st[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; i++) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; i++) st[i] = strides_bytes[i];
    CUresult r = reinterpret_cast<encode_tiled_fn>(h->encode_tiled)(
        out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
        const_cast<void *>(base), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return xb_fail(h, XB_ERR_CUDA, "cuTensorMapEncodeTiled (rank %d) failed with CUresult %d", rank, (int)r);
    return XB_OK;
}

// 3-D view of a (rows, 768) 16-bit activation matrix for the persistent LSTM: dims {64 k, rows, 12 k-blocks},
// strides {1536 B, 128 B}, box {64, box_rows, box_kblocks}, 128B swizzle: one TMA lands box_kblocks K-major
// [box_rows x 128 B] tiles.
int xb_make_tmap_hview(xb_handle *h, CUtensorMap *out, const void *base, uint64_t rows, uint32_t box_rows,
                       uint32_t box_kblocks) {
    if (!h->encode_tiled) {
        CUtensorMap tmp;
        if (int rc = xb_make_tmap_2d(h, &tmp, base, rows, XB_FEATURES, XB_FEATURES)) return rc;   // resolves the entry point
    }
    cuuint64_t dims[3] = {64, rows, XB_FEATURES / 64};
    cuuint64_t strides[2] = {XB_FEATURES * 2, 128};
    cuuint32_t box[3] = {64, box_rows, box_kblocks};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = reinterpret_cast<encode_tiled_fn>(h->encode_tiled)(
        out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims,
        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return xb_fail(h, XB_ERR_CUDA, "cuTensorMapEncodeTiled (3-D h view) failed with CUresult %d", (int)r);
    return XB_OK;
}

template <bool BF16, int EPI>
static int launch_one(xb_handle *h, const CUtensorMap &tmA, const CUtensorMap &tmB, const GemmParams &p, int rows_a,
                      cudaStream_t s) {
    auto k = gemm_tc_kernel<BF16, EPI>;
    constexpr int SMEM_BYTES = Pipe<EPI>::SMEM_BYTES;
    static bool configured[64] = {};   // per instantiation and device (function attributes live in the context)
    if (!configured[h->device & 63]) {
        XB_CUDA(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured[h->device & 63] = true;
    }
    dim3 grid((p.N + BN - 1) / BN, (rows_a + BM - 1) / BM, (EPI == EPI_F32 && p.split_k > 1) ? p.split_k : 1);
    k<<<grid, 256, SMEM_BYTES, s>>>(tmA, tmB, p);
    XB_LAUNCH_CHECK(h);
    return XB_OK;
}

int xb_gemm_launch(xb_handle *h, int epi, const CUtensorMap &tmA, const CUtensorMap &tmB, const GemmParams &p,
                   cudaStream_t s, bool bf16_operands) {
    if (p.K <= 0) return xb_fail(h, XB_ERR_ARG, "GEMM K=%d", p.K);
    if (bf16_operands) {      // the training backward: gradients travel as bf16 (range), accumulated in fp32
        switch (epi) {
            case EPI_F32: return launch_one<true, EPI_F32>(h, tmA, tmB, p, p.M, s);
            case EPI_BF16OUT: return launch_one<true, EPI_BF16OUT>(h, tmA, tmB, p, p.M, s);
        }
        return xb_fail(h, XB_ERR_ARG, "no bf16 GEMM with epilogue %d", epi);
    }
#define XB_EPI_CASE(E)                                                                       \
    case E:                                                                                  \
        return launch_one<false, E>(h, tmA, tmB, p, p.M, s);      /* fp16 operands in both weight modes (xb_api.cu repack) */
    switch (epi) {
        XB_EPI_CASE(EPI_F32)
        XB_EPI_CASE(EPI_CONV3_BWD)
#ifdef XB_EXPERIMENTS      // the tile-GEMM forms of conv3, the input projection, the head and the step-wise LSTM (cross-check builds only)
        XB_EPI_CASE(EPI_CONV3)
        XB_EPI_CASE(EPI_INPROJ)
        XB_EPI_CASE(EPI_HEAD)
        XB_EPI_CASE(EPI_LSTM)
#endif
    }
#undef XB_EPI_CASE
    return xb_fail(h, XB_ERR_ARG, "unknown epilogue %d", epi);
}
