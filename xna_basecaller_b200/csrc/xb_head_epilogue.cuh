// LinearCRFEncoder epilogue helpers shared by the tile GEMM (gemm_tc.cu) and the A-stationary GEMM (inproj_gemm.cu):
// scale * tanh(acc + bias) with the blank score written in front of every group of n_base columns (nn.py:117-129).
#pragma once
#include <stdint.h>

namespace xbhead {

__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

// 32 consecutive head columns whose first column sits at position E0 inside its group of NB.  NB and E0 are
// compile-time, so every staged position is an immediate offset (the generic loop spends more instructions on index
// bookkeeping than on the tanh).  Sg points at the staged position of the blank of the group that holds column 0.
template <int NB, int E0>
__device__ __forceinline__ void expand32(const uint32_t *acc, const float *__restrict__ bias, float scale, float blank, float *Sg) {
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const int g = (E0 + j) / NB, e = (E0 + j) % NB;
        const float v = scale * fast_tanh(__uint_as_float(acc[j]) + __ldg(bias + j));
        if (e == 0) Sg[g * (NB + 1)] = blank;
        Sg[g * (NB + 1) + 1 + e] = v;
    }
}
template <int NB>
__device__ __forceinline__ void expand32_nb(int e0, const uint32_t *acc, const float *__restrict__ bias, float scale, float blank,
                                            float *Sg) {
    switch (e0) {
        case 0: expand32<NB, 0>(acc, bias, scale, blank, Sg); break;
        case 1: expand32<NB, 1 % NB>(acc, bias, scale, blank, Sg); break;
        case 2: expand32<NB, 2 % NB>(acc, bias, scale, blank, Sg); break;
        case 3: expand32<NB, 3 % NB>(acc, bias, scale, blank, Sg); break;
        case 4: expand32<NB, 4 % NB>(acc, bias, scale, blank, Sg); break;
        default: expand32<NB, 5 % NB>(acc, bias, scale, blank, Sg); break;
    }
}

// One thread's 32 columns [col, col + 32) of a row into its staged row Sr (which starts at output position seg_start).
__device__ __forceinline__ void stage32(const uint32_t *acc, int col, int col_end, int nbs, int expand, const float *__restrict__ bias,
                                        float scale, float blank, float *Sr, int seg_start) {
    int c = col / nbs, e = col - c * nbs;
    if (expand && col + 32 <= col_end && nbs >= 4 && nbs <= 6) {      // all 32 columns valid: static path
        float *Sg = Sr + c * (nbs + 1) - seg_start;
        if (nbs == 5) expand32_nb<5>(e, acc, bias + col, scale, blank, Sg);
        else if (nbs == 4) expand32_nb<4>(e, acc, bias + col, scale, blank, Sg);
        else expand32_nb<6>(e, acc, bias + col, scale, blank, Sg);
        return;
    }
#pragma unroll
    for (int j = 0; j < 32; j++, col++) {
        if (col < col_end) {
            const float v = scale * fast_tanh(__uint_as_float(acc[j]) + __ldg(bias + col));
            if (expand) {
                const int o = col + c + 1 - seg_start;
                if (e == 0) Sr[o - 1] = blank;
                Sr[o] = v;
                if (++e == nbs) { e = 0; c++; }
            } else {
                Sr[col - seg_start] = v;
            }
        }
    }
}

// Output positions [seg_start, seg_start + seg_len) covered by head columns [col0, col_end).
__device__ __forceinline__ void segment(int col0, int col_end, int nbs, int expand, int &seg_start, int &seg_len) {
    seg_start = 0; seg_len = 0;
    if (col_end <= col0) return;
    if (expand) {
        seg_start = col0 + col0 / nbs + ((col0 % nbs) ? 1 : 0);
        seg_len = (col_end - 1) + (col_end - 1) / nbs + 2 - seg_start;
    } else {
        seg_start = col0;
        seg_len = col_end - col0;
    }
}

}  // namespace xbhead
