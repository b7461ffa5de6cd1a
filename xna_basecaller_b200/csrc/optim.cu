// Optimiser step of the training loop: global gradient-norm clipping + AdamW over all parameter tensors in two launches.
//
// Reference: Trainer.train_one_step (bonito/training.py:112-115): scaler.unscale_ -> torch.nn.utils.clip_grad_norm_(params,
// max_norm=2.0) -> scaler.step(optimizer) with torch.optim.AdamW (training.py:183-184).  The arithmetic follows torch's
// single-tensor AdamW (decoupled weight decay, bias-corrected moments, eps added after the bias-corrected square root):
//   total_norm = || (||g_1||, ..., ||g_n||) ||_2;  clip = min(1, max_norm / (total_norm + 1e-6));  g <- g * clip
//   p <- p (1 - lr wd);  m <- m + (1 - b1)(g - m);  v <- b2 v + (1 - b2) g^2
//   p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The gradients arrive unscaled (the backward transports them as bf16, no GradScaler), so unscale_ has no counterpart.
#include "xb_common.cuh"

namespace {

constexpr int MAX_TENSORS = 32;          // per launch (the kernel parameter block holds the pointer table)
constexpr int CHUNK = 65536;             // elements per block

struct TensorTable {
    float *p[MAX_TENSORS], *g[MAX_TENSORS], *m[MAX_TENSORS], *v[MAX_TENSORS];
    long long numel[MAX_TENSORS];
    int first_block[MAX_TENSORS + 1];    // prefix sum of ceil(numel / CHUNK)
    int n;
};

__device__ __forceinline__ int find_tensor(const TensorTable &tt, int block) {
    int i = 0;
    while (i + 1 < tt.n && tt.first_block[i + 1] <= block) i++;
    return i;
}

// partial[b] = sum of squares of the block's chunk (fixed reduction tree: the norm is reproducible run to run)
__global__ void __launch_bounds__(256) sumsq_kernel(const TensorTable tt, float *__restrict__ partial) {
    __shared__ float red[8];
    const int i = find_tensor(tt, blockIdx.x);
    const long long lo = (long long)(blockIdx.x - tt.first_block[i]) * CHUNK;
    const long long hi = lo + CHUNK < tt.numel[i] ? lo + CHUNK : tt.numel[i];
    const float *g = tt.g[i];
    float s = 0.0f;
    for (long long k = lo + threadIdx.x; k < hi; k += 256) { const float x = g[k]; s = fmaf(x, x, s); }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = red[0];
        for (int w = 1; w < 8; w++) t += red[w];
        partial[blockIdx.x] = t;
    }
}

// one block: total norm from the partial sums in order, accumulated onto *norm_sq (so that several launches can chain)
__global__ void __launch_bounds__(256) norm_finish_kernel(const float *__restrict__ partial, int n, float *__restrict__ norm_sq) {
    __shared__ float red[256];
    float s = 0.0f;
    for (int k = threadIdx.x; k < n; k += 256) s += partial[k];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off >= 1; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *norm_sq += red[0];
}

__global__ void __launch_bounds__(256)
adamw_kernel(const TensorTable tt, const float *__restrict__ norm_sq, float lr, float beta1, float beta2, float eps, float wd,
             float max_norm, float bc1, float bc2_sqrt) {
    const int i = find_tensor(tt, blockIdx.x);
    const long long lo = (long long)(blockIdx.x - tt.first_block[i]) * CHUNK;
    const long long hi = lo + CHUNK < tt.numel[i] ? lo + CHUNK : tt.numel[i];
    float clip = 1.0f;
    if (max_norm > 0.0f) {
        const float c = max_norm / (sqrtf(*norm_sq) + 1e-6f);
        clip = c < 1.0f ? c : 1.0f;
    }
    float *p = tt.p[i], *m = tt.m[i], *v = tt.v[i];
    const float *g = tt.g[i];
    const float step_size = lr / bc1;
    for (long long k = lo + threadIdx.x; k < hi; k += 256) {
        const float gk = g[k] * clip;
        float pk = p[k] * (1.0f - lr * wd);
        const float mk = m[k] + (1.0f - beta1) * (gk - m[k]);
        const float vk = beta2 * v[k] + (1.0f - beta2) * gk * gk;
        const float denom = sqrtf(vk) / bc2_sqrt + eps;
        pk -= step_size * (mk / denom);
        p[k] = pk; m[k] = mk; v[k] = vk;
    }
}

__global__ void sqrt_scalar_kernel(const float *a, float *o) { *o = sqrtf(*a); }

}  // namespace

extern "C" {

// One optimiser step over n_tensors fp32 tensors (host arrays of device pointers).  max_norm <= 0: no clipping.
// step >= 1 is the AdamW step count after this update.  grad_norm (device, 1 float) receives the total norm BEFORE
// clipping (what clip_grad_norm_ returns); scratch is device memory of at least 4 * (sum ceil(numel/65536) + 1) bytes.
int xb_adamw_step(float *const *params, float *const *grads, float *const *exp_avg, float *const *exp_avg_sq,
                  const int64_t *numel, int n_tensors, float lr, float beta1, float beta2, float eps, float weight_decay,
                  float max_norm, int64_t step, float *grad_norm, float *scratch, void *stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !numel || n_tensors <= 0 || step < 1 || !grad_norm || !scratch)
        return xb_fail(nullptr, XB_ERR_ARG, "xb_adamw_step: bad arguments");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    float *norm_sq = scratch, *partial = scratch + 1;
    if (cudaMemsetAsync(norm_sq, 0, sizeof(float), s) != cudaSuccess) return xb_fail(nullptr, XB_ERR_CUDA, "xb_adamw_step: memset failed");
    const float bc1 = 1.0f - powf(beta1, (float)step), bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
    std::vector<TensorTable> tables;
    for (int base = 0; base < n_tensors; base += MAX_TENSORS) {
        TensorTable tt;
        tt.n = n_tensors - base < MAX_TENSORS ? n_tensors - base : MAX_TENSORS;
        tt.first_block[0] = 0;
        for (int i = 0; i < tt.n; i++) {
            tt.p[i] = params[base + i]; tt.g[i] = grads[base + i]; tt.m[i] = exp_avg[base + i]; tt.v[i] = exp_avg_sq[base + i];
            tt.numel[i] = numel[base + i];
            tt.first_block[i + 1] = tt.first_block[i] + (int)((numel[base + i] + CHUNK - 1) / CHUNK);
        }
        tables.push_back(tt);
    }
    for (auto &tt : tables) {
        const int blocks = tt.first_block[tt.n];
        if (blocks == 0) continue;
        sumsq_kernel<<<blocks, 256, 0, s>>>(tt, partial);
        norm_finish_kernel<<<1, 256, 0, s>>>(partial, blocks, norm_sq);
    }
    for (auto &tt : tables) {
        const int blocks = tt.first_block[tt.n];
        if (blocks == 0) continue;
        adamw_kernel<<<blocks, 256, 0, s>>>(tt, norm_sq, lr, beta1, beta2, eps, weight_decay, max_norm, bc1, bc2_sqrt);
    }
    sqrt_scalar_kernel<<<1, 1, 0, s>>>(norm_sq, grad_norm);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return xb_fail(nullptr, XB_ERR_CUDA, "xb_adamw_step: launch failed: %s", cudaGetErrorString(e));
    return XB_OK;
}

}  // extern "C"
