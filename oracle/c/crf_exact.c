/* Plain-C restatement of the CTC-CRF decode arithmetic -- the bit-exact checker for the CUDA decode.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): built by oracle/c/Makefile into
 * oracle/_build/libxb_oracle.so and loaded by tests/, smoke() and bench.py's cpu_baseline leg.
 *
 * Follows (paths relative to /root/reference/ub-bonito/):
 *   bonito/crf/model.py:26-36    CTC_CRF.idx                       -> src_state()
 *   bonito/crf/model.py:41-46    CTC_CRF.logZ (alpha_0 = beta_T = 0) -> xbo_crf_alpha / xbo_crf_logz
 *   bonito/crf/model.py:92-95    CTC_CRF.viterbi (posteriors(Max).argmax % NZ) -> xbo_crf_viterbi
 *   bonito/crf/model.py:215-218  SeqdistModel.decode_batch (posteriors + 1e-8, log, viterbi, int16, T)
 *   bonito/crf/model.py:97-100   CTC_CRF.path_to_str                -> xbo_pack
 *   bonito/crf/basecall.py:56-76 left-packed (N,T) int8 sequence / qstring
 * and the seqdist semiring algebra restated in oracle/seqdist_restated.py (Log: logsumexp / softmax,
 * Max: max / one-hot at first arg-max; edge marginal = (M + alpha[idx]) + beta).
 *
 * Arithmetic contract shared with xna_basecaller_b200/csrc/crf_decode.cu (all fp32, round to nearest,
 * no contraction; exp/log from xb_exact_math.h):
 *   logsumexp over the NZ edges of a state:  m = max_k x_k;  s = (..(e_0 + e_1) + ..) + e_{NZ-1},
 *       e_k = exp(x_k - m);  result = m + log(s).   Backward (beta) edge order: stay, then moves j = 0..n-1.
 *   softmax over all C*NZ edges of a step, factored per SOURCE state s: z = M + beta_{t+1}[dst] over the edges leaving
 *       s (stay, then moves), m_s = max z, w = exp(z - m_s) (shared with Log-beta_t[s] = m_s + log(sum w));
 *       g_s = m_s + alpha_t[s], gmax = max_s g_s, f_s = exp(g_s - gmax); numerator e = w * f_src; per destination
 *       state s_c = sum_k e (k order); S = tree_sum(s_c) (below); p = e * (1 / S);  lp = log(p + 1e-8).
 *   tree_sum over states: state c sits in lane c%32 of warp c/32 (missing lanes contribute +0); each warp
 *       does the xor-butterfly v_i += v_{i^off}, off = 16,8,4,2,1; warp totals are added in warp order.
 *   arg-max over the flat edge index c*NZ+k: strictly greater wins, i.e. first index on ties.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../xna_basecaller_b200/csrc/xb_exact_math.h"

#define MAXC 1296
#define MAXNZ 8

typedef struct {
    int n, sl, C, NZ, n_pow;   /* n_pow = n^(sl-1) */
} lattice;

static int lattice_init(lattice *L, int n_base, int state_len) {
    L->n = n_base; L->sl = state_len; L->NZ = n_base + 1;
    int C = 1, p = 1;
    for (int i = 0; i < state_len; i++) { C *= n_base; if (i < state_len - 1) p *= n_base; }
    L->C = C; L->n_pow = p;
    return (C <= MAXC && L->NZ <= MAXNZ && n_base >= 2) ? 0 : -1;
}

/* source state of edge k entering state c (crf/model.py:31-36) */
static inline int src_state(const lattice *L, int c, int k) {
    return k == 0 ? c : (k - 1) * L->n_pow + c / L->n;
}

static float tree_sum(const float *v, int C) {
    float total = 0.0f;
    int W = (C + 31) / 32;
    for (int w = 0; w < W; w++) {
        float a[32], b[32];
        for (int i = 0; i < 32; i++) { int c = w * 32 + i; a[i] = c < C ? v[c] : 0.0f; }
        for (int off = 16; off >= 1; off >>= 1) {
            for (int i = 0; i < 32; i++) b[i] = XB_ADD(a[i], a[i ^ off]);
            memcpy(a, b, sizeof a);
        }
        total = (w == 0) ? a[0] : XB_ADD(total, a[0]);
    }
    return total;
}

static inline float lse_finish(float m, float s) { return XB_ADD(m, xb_logf(s)); }

/* alpha_{t+1}[c] from alpha_t and the scores of step t */
static void alpha_step(const lattice *L, const float *M, const float *a, float *out, int use_max) {
    for (int c = 0; c < L->C; c++) {
        float x[MAXNZ], m = 0.0f;
        for (int k = 0; k < L->NZ; k++) {
            x[k] = XB_ADD(M[c * L->NZ + k], a[src_state(L, c, k)]);
            m = (k == 0 || x[k] > m) ? x[k] : m;
        }
        if (use_max) { out[c] = m; continue; }
        float s = 0.0f;
        for (int k = 0; k < L->NZ; k++) {
            float e = xb_expf(XB_SUB(x[k], m));
            s = (k == 0) ? e : XB_ADD(s, e);
        }
        out[c] = lse_finish(m, s);
    }
}

/* beta_t[c'] from beta_{t+1} and the scores of step t; edges leaving c': stay (c',0) then the moves
 * into c = (c' % n_pow)*n + j through edge k = 1 + c'/n_pow, j = 0..n-1 */
static void beta_step(const lattice *L, const float *M, const float *b, float *out, int use_max) {
    for (int cp = 0; cp < L->C; cp++) {
        float x[MAXNZ], m;
        x[0] = XB_ADD(M[cp * L->NZ], b[cp]);
        m = x[0];
        int k = 1 + cp / L->n_pow, base = (cp % L->n_pow) * L->n;
        for (int j = 0; j < L->n; j++) {
            int c = base + j;
            x[1 + j] = XB_ADD(M[c * L->NZ + k], b[c]);
            m = x[1 + j] > m ? x[1 + j] : m;
        }
        if (use_max) { out[cp] = m; continue; }
        float s = 0.0f;
        for (int j = 0; j < L->NZ; j++) {
            float e = xb_expf(XB_SUB(x[j], m));
            s = (j == 0) ? e : XB_ADD(s, e);
        }
        out[cp] = lse_finish(m, s);
    }
}

int xbo_crf_alpha(const float *scores, int T, int N, int n_base, int state_len, float *alpha) {
    lattice L; if (lattice_init(&L, n_base, state_len)) return -1;
    size_t S = (size_t)L.C * L.NZ;
    for (int n = 0; n < N; n++) {
        for (int c = 0; c < L.C; c++) alpha[(size_t)n * L.C + c] = 0.0f;
        for (int t = 0; t < T; t++)
            alpha_step(&L, scores + ((size_t)t * N + n) * S, alpha + ((size_t)t * N + n) * L.C,
                       alpha + ((size_t)(t + 1) * N + n) * L.C, 0);
    }
    return 0;
}

int xbo_crf_logz(const float *scores, int T, int N, int n_base, int state_len, float *logz) {
    lattice L; if (lattice_init(&L, n_base, state_len)) return -1;
    size_t S = (size_t)L.C * L.NZ;
    for (int n = 0; n < N; n++) {
        float a[2][MAXC];
        for (int c = 0; c < L.C; c++) a[0][c] = 0.0f;
        for (int t = 0; t < T; t++) alpha_step(&L, scores + ((size_t)t * N + n) * S, a[t & 1], a[(t + 1) & 1], 0);
        const float *aT = a[T & 1];
        float m = aT[0], e[MAXC];
        for (int c = 1; c < L.C; c++) m = aT[c] > m ? aT[c] : m;
        for (int c = 0; c < L.C; c++) e[c] = xb_expf(XB_SUB(aT[c], m));
        logz[n] = lse_finish(m, tree_sum(e, L.C));
    }
    return 0;
}

/* Max-semiring forward sweep + arg-max against stored bmax (T+1, C) of ONE sequence (stride C). */
static void viterbi_forward(const lattice *L, const float *lp, size_t lp_stride, const float *bmax, int T,
                            int8_t *labels) {
    float am[2][MAXC];
    for (int c = 0; c < L->C; c++) am[0][c] = 0.0f;
    for (int t = 0; t < T; t++) {
        const float *M = lp + (size_t)t * lp_stride, *a = am[t & 1], *b = bmax + (size_t)(t + 1) * L->C;
        float *an = am[(t + 1) & 1];
        float best = 0.0f; int besti = -1;
        for (int c = 0; c < L->C; c++) {
            float m = 0.0f;
            for (int k = 0; k < L->NZ; k++) {
                float v = XB_ADD(M[c * L->NZ + k], a[src_state(L, c, k)]);
                m = (k == 0 || v > m) ? v : m;
                float sc = XB_ADD(v, b[c]);
                if (besti < 0 || sc > best) { best = sc; besti = c * L->NZ + k; }
            }
            an[c] = m;
        }
        labels[t] = (int8_t)(besti % L->NZ);
    }
}

/* decode_batch: labels (N,T) int8; optional post / lp outputs (T,N,C*NZ). */
/* sequences [n_begin, n_end) only: lets the caller spread a batch over host threads (bench.py's CPU baseline) */
int xbo_crf_decode_range(const float *scores, int T, int N, int n_begin, int n_end, int n_base, int state_len,
                         float *post, float *lp_out, int8_t *labels) {
    lattice L; if (lattice_init(&L, n_base, state_len)) return -1;
    size_t S = (size_t)L.C * L.NZ;
    float *alpha = (float *)malloc((size_t)(T + 1) * L.C * sizeof(float));
    float *bmax = (float *)malloc((size_t)(T + 1) * L.C * sizeof(float));
    float *lp = (float *)malloc((size_t)T * S * sizeof(float));
    float *x = (float *)malloc(S * sizeof(float));
    float *sc = (float *)malloc(L.C * sizeof(float));
    float *gs = (float *)malloc(L.C * sizeof(float));
    if (!alpha || !bmax || !lp || !x || !sc || !gs) return -2;
    for (int n = n_begin; n < n_end; n++) {
        for (int c = 0; c < L.C; c++) alpha[c] = 0.0f;
        for (int t = 0; t < T; t++)
            alpha_step(&L, scores + ((size_t)t * N + n) * S, alpha + (size_t)t * L.C, alpha + (size_t)(t + 1) * L.C, 0);
        float beta[2][MAXC];
        for (int c = 0; c < L.C; c++) { beta[T & 1][c] = 0.0f; bmax[(size_t)T * L.C + c] = 0.0f; }
        for (int t = T - 1; t >= 0; t--) {
            const float *M = scores + ((size_t)t * N + n) * S, *a = alpha + (size_t)t * L.C;
            const float *b1 = beta[(t + 1) & 1];
            /* Edges grouped by SOURCE state s (stay, then the moves j = 0..n-1): z = M + beta_{t+1}[dst],
             * m_s = max z, w = exp(z - m_s): sum_w gives Log-beta_t[s], and the softmax numerator of an edge factors
             * as exp(x - gmax) = w * f_s with x = z + alpha_t[s], f_s = exp((m_s + alpha_t[s]) - gmax): one exp per
             * state instead of one per edge for the posterior path. */
            float gmax = 0.0f;
            float *bt = beta[t & 1];
            for (int sp = 0; sp < L.C; sp++) {
                float z[MAXNZ], m;
                int kk = 1 + sp / L.n_pow, base = (sp % L.n_pow) * L.n;
                z[0] = XB_ADD(M[sp * L.NZ], b1[sp]);
                m = z[0];
                for (int j = 0; j < L.n; j++) {
                    z[1 + j] = XB_ADD(M[(base + j) * L.NZ + kk], b1[base + j]);
                    m = z[1 + j] > m ? z[1 + j] : m;
                }
                float sy = 0.0f;
                for (int j = 0; j < L.NZ; j++) {
                    float w = xb_expf(XB_SUB(z[j], m));
                    x[j == 0 ? sp * L.NZ : (base + j - 1) * L.NZ + kk] = w;
                    sy = (j == 0) ? w : XB_ADD(sy, w);
                }
                bt[sp] = lse_finish(m, sy);
                gs[sp] = XB_ADD(m, a[sp]);
                gmax = (sp == 0 || gs[sp] > gmax) ? gs[sp] : gmax;
            }
            for (int sp = 0; sp < L.C; sp++) gs[sp] = xb_expf(XB_SUB(gs[sp], gmax));      /* f_s */
            for (int c = 0; c < L.C; c++) {
                float s = 0.0f;
                for (int k = 0; k < L.NZ; k++) {
                    float e = XB_MUL(x[c * L.NZ + k], gs[src_state(&L, c, k)]);
                    x[c * L.NZ + k] = e;
                    s = (k == 0) ? e : XB_ADD(s, e);
                }
                sc[c] = s;
            }
            float inv = XB_RCP(tree_sum(sc, L.C));
            float *lpt = lp + (size_t)t * S;
            for (size_t i = 0; i < S; i++) {
                float p = XB_MUL(x[i], inv);
                if (post) post[((size_t)t * N + n) * S + i] = p;
                lpt[i] = xb_logf(XB_ADD(p, XB_POST_EPS));
            }
            if (lp_out) memcpy(lp_out + ((size_t)t * N + n) * S, lpt, S * sizeof(float));
            beta_step(&L, lpt, bmax + (size_t)(t + 1) * L.C, bmax + (size_t)t * L.C, 1);
        }
        viterbi_forward(&L, lp, S, bmax, T, labels + (size_t)n * T);
    }
    free(alpha); free(bmax); free(lp); free(x); free(sc); free(gs);
    return 0;
}

int xbo_crf_decode(const float *scores, int T, int N, int n_base, int state_len, float *post, float *lp_out,
                   int8_t *labels) {
    return xbo_crf_decode_range(scores, T, N, 0, N, n_base, state_len, post, lp_out, labels);
}

/* ---------------------------------------------------------------------------------------------------------------
 * Linear-domain (scaled) decode: the arithmetic contract of the CUDA kernels crf_lin_alpha / crf_lin_backward /
 * crf_lin_viterbi in xna_basecaller_b200/csrc/crf_decode_lin.cu.  Same reference semantics as xbo_crf_decode
 * (bonito/crf/model.py:41-46, 92-95, 215-218: Log-semiring posteriors, + 1e-8, Max-semiring marginals of their logs,
 * arg-max % NZ), restated without per-state logarithms:
 *   E = exp(M) per edge (xb_expf; scores clamped to [-80, 80]);  Log semiring -> sums of products;  the Max semiring
 *   over log(p + 1e-8) -> max of products of (p + 1e-8) (log is monotone, so the arg-max is the same in exact arithmetic).
 *   Every state vector is rescaled by a power of two taken from its own maximum (exact), so nothing over/underflows:
 *     scale(mx) = 2^(127 - biased_exponent(mx))  (1 if mx is zero, subnormal, inf or nan);  v_hat = v * scale(max v).
 *   forward:   u_{t+1}[c] = fma-chain over k (k order) of E_t[c,k] * a_hat_t[src(c,k)],  a_hat_0 = 1
 *   backward:  w_e = E_t[e] * b_hat_{t+1}[dst(e)] over the edges leaving s (stay, then moves j = 0..n-1);
 *              u_t[s] = ((w_0 + w_1) + ...);  b_hat_T = 1
 *   posterior: x_e = a_hat_t[s] * w_e;  xs_s = ((x_0 + x_1) + ...);  tot_t = tree_sum over states (as above);
 *              p_e = x_e * (1 / tot_t);  P_e = p_e + 1e-8
 *   Max-beta:  um_t[s] = max_e P_e * bm_hat_{t+1}[dst(e)],  bm_hat_T = 1
 *   Max-alpha: v_k = P_t[c,k] * am_hat_t[src(c,k)];  um_{t+1}[c] = max_k v_k,  am_hat_0 = 1
 *   label_t  = (arg-max over the flat edge index c*NZ+k of v_k * bm_hat_{t+1}[c], first index on ties) % NZ
 * lin_input != 0: `scores` already holds E (what the fused head writes). */
#define pow2_scale xb_pow2_scale
static inline float vec_max(const float *v, int C) {
    float m = v[0];
    for (int c = 1; c < C; c++) m = v[c] > m ? v[c] : m;
    return m;
}
static inline float edge_E(const float *row, size_t i, int lin_input) {
    return lin_input ? row[i] : xb_score_exp(row[i]);
}

int xbo_crf_decode_lin_range(const float *scores, int lin_input, int T, int N, int n_begin, int n_end, int n_base,
                             int state_len, float *post, int8_t *labels) {
    lattice L; if (lattice_init(&L, n_base, state_len)) return -1;
    const int C = L.C, NZ = L.NZ;
    size_t S = (size_t)C * NZ;
    float *ah = (float *)malloc((size_t)(T + 1) * C * sizeof(float));     /* a_hat */
    float *bh = (float *)malloc((size_t)(T + 1) * C * sizeof(float));     /* b_hat */
    float *bm = (float *)malloc((size_t)(T + 1) * C * sizeof(float));     /* bm_hat */
    float *tot = (float *)malloc((size_t)T * sizeof(float));
    float *E = (float *)malloc(S * sizeof(float));
    if (!ah || !bh || !bm || !tot || !E) return -2;
    for (int n = n_begin; n < n_end; n++) {
        float u[MAXC], xs[MAXC];
        /* forward */
        for (int c = 0; c < C; c++) ah[c] = 1.0f;
        for (int t = 0; t < T; t++) {
            const float *row = scores + ((size_t)t * N + n) * S, *a = ah + (size_t)t * C;
            for (int c = 0; c < C; c++) {
                float acc = XB_MUL(edge_E(row, (size_t)c * NZ, lin_input), a[c]);
                for (int k = 1; k < NZ; k++)
                    acc = XB_FMA(edge_E(row, (size_t)c * NZ + k, lin_input), a[src_state(&L, c, k)], acc);
                u[c] = acc;
            }
            float sc = pow2_scale(vec_max(u, C));
            for (int c = 0; c < C; c++) ah[(size_t)(t + 1) * C + c] = XB_MUL(u[c], sc);
        }
        /* backward: b_hat, posteriors' normaliser, Max-beta over P */
        for (int c = 0; c < C; c++) { bh[(size_t)T * C + c] = 1.0f; bm[(size_t)T * C + c] = 1.0f; }
        for (int t = T - 1; t >= 0; t--) {
            const float *row = scores + ((size_t)t * N + n) * S;
            const float *a = ah + (size_t)t * C, *b1 = bh + (size_t)(t + 1) * C, *m1 = bm + (size_t)(t + 1) * C;
            for (size_t i = 0; i < S; i++) E[i] = edge_E(row, i, lin_input);
            float um[MAXC];
            /* first the sums (they define tot), then the Max recursion which needs 1 / tot */
            for (int s = 0; s < C; s++) {
                int kk = 1 + s / L.n_pow, base = (s % L.n_pow) * L.n;
                float w = XB_MUL(E[(size_t)s * NZ], b1[s]);
                float x = XB_MUL(a[s], w);
                float su = w, sx = x;
                for (int j = 0; j < L.n; j++) {
                    w = XB_MUL(E[(size_t)(base + j) * NZ + kk], b1[base + j]);
                    x = XB_MUL(a[s], w);
                    su = XB_ADD(su, w);
                    sx = XB_ADD(sx, x);
                }
                u[s] = su;
                xs[s] = sx;
            }
            tot[t] = tree_sum(xs, C);
            float inv = XB_RCP(tot[t]);
            for (int s = 0; s < C; s++) {
                int kk = 1 + s / L.n_pow, base = (s % L.n_pow) * L.n;
                float x = XB_MUL(a[s], XB_MUL(E[(size_t)s * NZ], b1[s]));
                float p = XB_MUL(x, inv);
                if (post) post[((size_t)t * N + n) * S + (size_t)s * NZ] = p;
                float best = XB_MUL(XB_ADD(p, XB_POST_EPS), m1[s]);
                for (int j = 0; j < L.n; j++) {
                    size_t e = (size_t)(base + j) * NZ + kk;
                    x = XB_MUL(a[s], XB_MUL(E[e], b1[base + j]));
                    p = XB_MUL(x, inv);
                    if (post) post[((size_t)t * N + n) * S + e] = p;
                    float v = XB_MUL(XB_ADD(p, XB_POST_EPS), m1[base + j]);
                    best = v > best ? v : best;
                }
                um[s] = best;
            }
            float sb = pow2_scale(vec_max(u, C)), sm = pow2_scale(vec_max(um, C));
            for (int s = 0; s < C; s++) {
                bh[(size_t)t * C + s] = XB_MUL(u[s], sb);
                bm[(size_t)t * C + s] = XB_MUL(um[s], sm);
            }
        }
        /* forward, Max semiring over P, arg-max */
        float am[2][MAXC];
        for (int c = 0; c < C; c++) am[0][c] = 1.0f;
        for (int t = 0; t < T; t++) {
            const float *row = scores + ((size_t)t * N + n) * S;
            const float *a = ah + (size_t)t * C, *b1 = bh + (size_t)(t + 1) * C, *m1 = bm + (size_t)(t + 1) * C;
            const float *ac = am[t & 1];
            float inv = XB_RCP(tot[t]);
            float best = 0.0f; int besti = -1;
            for (int c = 0; c < C; c++) {
                float m = 0.0f;
                for (int k = 0; k < NZ; k++) {
                    int sidx = src_state(&L, c, k);
                    float w = XB_MUL(edge_E(row, (size_t)c * NZ + k, lin_input), b1[c]);
                    float x = XB_MUL(a[sidx], w);
                    float P = XB_ADD(XB_MUL(x, inv), XB_POST_EPS);
                    float v = XB_MUL(P, ac[sidx]);
                    m = (k == 0 || v > m) ? v : m;
                    float cand = XB_MUL(v, m1[c]);
                    if (besti < 0 || cand > best) { best = cand; besti = c * NZ + k; }
                }
                u[c] = m;
            }
            float sc = pow2_scale(vec_max(u, C));
            for (int c = 0; c < C; c++) am[(t + 1) & 1][c] = XB_MUL(u[c], sc);
            labels[(size_t)n * T + t] = (int8_t)(besti % NZ);
        }
    }
    free(ah); free(bh); free(bm); free(tot); free(E);
    return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Beam search over the CRF lattice: the plain-C statement of xna_basecaller_b200/csrc/beam_search.cu (see the algorithm in
 * that file's header; koi.decode.beam_search, which the reference calls at crf/basecall.py:33-46, is third-party and
 * ACGT-only, so this decoder is pinned by brute force in the tests, not against koi).
 * labels (N,T) int8 (0 = stay, k = move through edge k), quals (N,T) uint8 (phred+33 at emitting steps, else 0). */
static int beta_rank(const float *b0, int C, int c) {
    int r = 0;
    for (int d = 0; d < C; d++) r += (b0[d] > b0[c]) || (b0[d] == b0[c] && d < c);
    return r;
}

int xbo_crf_beam_search(const float *scores, int T, int N, int n_base, int state_len, int beam_width, float beam_cut,
                        int8_t *labels, uint8_t *quals) {
    lattice L; if (lattice_init(&L, n_base, state_len)) return -1;
    const int C = L.C, NZ = L.NZ, W = beam_width;
    if (W < 1 || W > 32 || W * NZ > 256) return -3;
    size_t S = (size_t)C * NZ;
    float *beta = (float *)malloc((size_t)(T + 1) * C * sizeof(float));
    unsigned char *back = (unsigned char *)malloc((size_t)T * 32 + 1);
    if (!beta || !back) return -2;
    const unsigned long long FNV = 1099511628211ULL;
    for (int n = 0; n < N; n++) {
        for (int c = 0; c < C; c++) beta[(size_t)T * C + c] = 0.0f;
        for (int t = T - 1; t >= 0; t--)
            beta_step(&L, scores + ((size_t)t * N + n) * S, beta + (size_t)(t + 1) * C, beta + (size_t)t * C, 0);
        float bscore[2][32]; int bstate[2][32], istate[32]; unsigned long long bhash[2][32];
        int nb = W < C ? W : C;
        for (int c = 0; c < C; c++) {
            int r = beta_rank(beta, C, c);
            if (r < W) { bscore[0][r] = 0.0f; bstate[0][r] = c; istate[r] = c; bhash[0][r] = (unsigned long long)c + 1ULL; }
        }
        for (int t = 0; t < T; t++) {
            const int cur = t & 1, nxt = cur ^ 1, ncand = nb * NZ;
            const float *M = scores + ((size_t)t * N + n) * S, *b1 = beta + (size_t)(t + 1) * C;
            float cscore[256], ckey[256], lsc[256]; int cstate[256], cfirst[256], alive[256]; unsigned long long chash[256];
            for (int i = 0; i < ncand; i++) {
                int e = i / NZ, k = i - e * NZ, s = bstate[cur][e], s2 = s, edge = s * NZ;
                unsigned long long h = bhash[cur][e];
                if (k > 0) { s2 = (s % L.n_pow) * L.n + (k - 1); edge = s2 * NZ + 1 + s / L.n_pow; h = h * FNV + (unsigned long long)k; }
                cscore[i] = XB_ADD(bscore[cur][e], M[edge]); cstate[i] = s2; chash[i] = h;
            }
            for (int i = 0; i < ncand; i++) {
                cfirst[i] = i;
                for (int j = 0; j < i; j++) if (chash[j] == chash[i] && cstate[j] == cstate[i]) { cfirst[i] = j; break; }
            }
            for (int i = 0; i < ncand; i++) {
                alive[i] = cfirst[i] == i;
                lsc[i] = XB_NEG_BIG; ckey[i] = XB_NEG_BIG;
                if (!alive[i]) continue;
                float m = cscore[i];
                for (int j = i + 1; j < ncand; j++) if (cfirst[j] == i && cscore[j] > m) m = cscore[j];
                float ssum = 0.0f; int any = 0;
                for (int j = i; j < ncand; j++) if (cfirst[j] == i) {
                    float ex = xb_expf(XB_SUB(cscore[j], m));
                    ssum = any ? XB_ADD(ssum, ex) : ex; any = 1;
                }
                lsc[i] = XB_ADD(m, xb_logf(ssum));
                ckey[i] = XB_ADD(lsc[i], b1[cstate[i]]);
            }
            float best = XB_NEG_BIG;
            for (int j = 0; j < ncand; j++) if (alive[j] && ckey[j] > best) best = ckey[j];
            int cnt = 0;
            for (int i = 0; i < ncand; i++) {
                if (!alive[i]) continue;
                int rank = 0;
                for (int j = 0; j < ncand; j++) if (alive[j]) rank += (ckey[j] > ckey[i]) || (ckey[j] == ckey[i] && j < i);
                if (rank < W && ckey[i] >= XB_SUB(best, beam_cut)) {
                    bscore[nxt][rank] = lsc[i]; bstate[nxt][rank] = cstate[i]; bhash[nxt][rank] = chash[i];
                    back[(size_t)t * 32 + rank] = (unsigned char)((i / NZ) | ((i % NZ) << 5));
                    cnt++;
                }
            }
            nb = cnt;
        }
        int slot = 0;
        for (int t = T - 1; t >= 0; t--) {
            unsigned char bk = back[(size_t)t * 32 + slot];
            labels[(size_t)n * T + t] = (int8_t)(bk >> 5);
            slot = bk & 31;
        }
        int s = istate[slot];          /* forward along the path: new base (b + 1) -> label = edge index 1 + dropped base */
        for (int t = 0; t < T; t++) {
            int b1 = labels[(size_t)n * T + t];
            uint8_t q = 0;
            if (b1 > 0) {
                int s2 = (s % L.n_pow) * L.n + (b1 - 1), k = 1 + s / L.n_pow;
                const float *M = scores + ((size_t)t * N + n) * S;
                float lp = XB_SUB(XB_ADD(M[s2 * NZ + k], beta[(size_t)(t + 1) * C + s2]), beta[(size_t)t * C + s]);
                float err = XB_SUB(1.0f, xb_expf(lp));
                err = err < 1e-5f ? 1e-5f : err;
                float qf = XB_MUL(-4.34294481903251828f, xb_logf(err));
                int qi = (int)(XB_ADD(qf, 0.5f));
                qi = qi < 1 ? 1 : (qi > 50 ? 50 : qi);
                q = (uint8_t)(33 + qi);
                labels[(size_t)n * T + t] = (int8_t)k;
                s = s2;
            }
            if (quals) quals[(size_t)n * T + t] = q;
        }
    }
    free(beta); free(back);
    return 0;
}

int xbo_crf_posteriors(const float *scores, int T, int N, int n_base, int state_len, float *post) {
    int8_t *lab = (int8_t *)malloc((size_t)N * T);
    if (!lab) return -2;
    int rc = xbo_crf_decode(scores, T, N, n_base, state_len, post, NULL, lab);
    free(lab);
    return rc;
}

/* CTC_CRF.viterbi on arbitrary scores: labels (N,T) int8 */
int xbo_crf_viterbi(const float *scores, int T, int N, int n_base, int state_len, int8_t *labels) {
    lattice L; if (lattice_init(&L, n_base, state_len)) return -1;
    size_t S = (size_t)L.C * L.NZ;
    float *bmax = (float *)malloc((size_t)(T + 1) * L.C * sizeof(float));
    if (!bmax) return -2;
    for (int n = 0; n < N; n++) {
        for (int c = 0; c < L.C; c++) bmax[(size_t)T * L.C + c] = 0.0f;
        for (int t = T - 1; t >= 0; t--)
            beta_step(&L, scores + ((size_t)t * N + n) * S, bmax + (size_t)(t + 1) * L.C, bmax + (size_t)t * L.C, 1);
        viterbi_forward(&L, scores + (size_t)n * S, (size_t)N * S, bmax, T, labels + (size_t)n * T);
    }
    free(bmax);
    return 0;
}

/* path_to_str + left-pack: seq/qstring (N,T) int8 zero padded, lens (N) */
int xbo_pack(const int8_t *labels, int N, int T, const char *alphabet, int8_t *seq, int8_t *qstring, int32_t *lens) {
    for (int n = 0; n < N; n++) {
        int len = 0;
        memset(seq + (size_t)n * T, 0, T);
        memset(qstring + (size_t)n * T, 0, T);
        for (int t = 0; t < T; t++) {
            int l = labels[(size_t)n * T + t];
            if (l != 0) { seq[(size_t)n * T + len] = (int8_t)alphabet[l]; qstring[(size_t)n * T + len] = 'O'; len++; }
        }
        lens[n] = len;
    }
    return 0;
}

void xbo_score_exp_array(const float *x, float *y, long n) { for (long i = 0; i < n; i++) y[i] = xb_score_exp(x[i]); }
void xbo_expf_array(const float *x, float *y, long n) { for (long i = 0; i < n; i++) y[i] = xb_expf(x[i]); }
void xbo_logf_array(const float *x, float *y, long n) { for (long i = 0; i < n; i++) y[i] = xb_logf(x[i]); }
void xbo_expf_le0_array(const float *x, float *y, long n) { for (long i = 0; i < n; i++) y[i] = xb_expf_le0(x[i]); }
void xbo_logf_norm_array(const float *x, float *y, long n) { for (long i = 0; i < n; i++) y[i] = xb_logf_norm(x[i]); }
