// Losses of the adversarial training step, forward and backward in ONE pass over the data
// (cm/train_trgb_segnet_conf.py:237-245 criteria; :437-446 critic terms; :452 CrossEntropyLoss on the day logits;
// scripts/main.py:223 CrossEntropyLoss(ignore_index=13)).  HBM-bound: the cross entropy reads the K logits of a
// pixel once (NCHW: class planes, so consecutive threads read consecutive pixels of every plane) and writes the K
// gradient values; the constant-target critic losses read the map once and write its gradient.
//
// Reductions are deterministic: FP64 per-block partials in a caller-provided scratch, folded by a second single-block
// kernel in block order (no floating-point atomics).
#include "hn_common.cuh"

namespace hn {

constexpr int kLossThreads = 256;

__device__ __forceinline__ double block_sum(double v, double *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;  // valid in thread 0
}

// number of labels != ignore_index (and inside [0,K)); out-of-range labels raise bit 0 of *flags like torch's device assert
__global__ void __launch_bounds__(kLossThreads) ce_count_kernel(const long long *__restrict__ labels, long long n, int K, long long ignore,
                                                                double *partial, int *flags)
{
    __shared__ double sh[32];
    double cnt = 0.0;
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long t = __ldg(labels + i);
        if (t == ignore) continue;
        if (t < 0 || t >= K) { bad = 1; continue; }
        cnt += 1.0;
    }
    if (bad) atomicOr(flags, 1);
    double s = block_sum(cnt, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out = (sum of `nblocks` partials, in block order) / denominator; denominator 0 -> NaN (torch's 0/0 for an all-ignored batch)
__global__ void loss_fold_kernel(const double *__restrict__ partial, int nblocks, const double *denom_ptr, double denom, double *out_f64,
                                 float *out_f32)
{
    __shared__ double sh[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) v += partial[i];
    double s = block_sum(v, sh);
    if (threadIdx.x == 0) {
        double d = denom_ptr ? *denom_ptr : denom;
        double r = d != 0.0 ? s / d : nan("");
        if (out_f64) out_f64[0] = r;
        if (out_f32) out_f32[0] = (float)r;
    }
}

// logits: NCHW FP32 [N][K][HW]; labels int64 [N][HW]; dlogits (optional) same layout as logits.
// partial[block] = sum over the block's pixels of -log softmax(logits)[label]
template <int KMAX>
__global__ void __launch_bounds__(kLossThreads) ce_fwd_bwd_kernel(const float *__restrict__ logits, const long long *__restrict__ labels, long long n_img,
                                                                  long long hw, int K, long long ignore, const double *count_ptr, double count,
                                                                  float grad_scale, float *__restrict__ dlogits, double *partial)
{
    __shared__ double sh[32];
    const long long npix = n_img * hw;
    double cnt = count_ptr ? *count_ptr : count;
    const float gs = cnt > 0.0 ? (float)((double)grad_scale / cnt) : 0.f;
    double acc = 0.0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        const long long n = p / hw, q = p - n * hw;
        const float *src = logits + (n * K) * hw + q;
        float v[KMAX];
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) {
                v[k] = __ldg(src + (long long)k * hw);
                m = fmaxf(m, v[k]);
            }
        const long long t = __ldg(labels + p);
        const bool live = t != ignore && t >= 0 && t < K;
        float xt = 0.f, s = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) {
                if (k == (int)t) xt = v[k];
                v[k] = expf(v[k] - m);
                s += v[k];
            }
        const float inv = 1.f / s;
        if (live) acc += (double)(logf(s) - (xt - m));      // -log softmax(x)[t]
        if (dlogits) {
            float *dst = dlogits + (n * K) * hw + q;
#pragma unroll
            for (int k = 0; k < KMAX; ++k)
                if (k < K) dst[(long long)k * hw] = live ? (v[k] * inv - (k == (int)t ? 1.f : 0.f)) * gs : 0.f;
        }
    }
    double s = block_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// critic-map losses against a constant (tgt == nullptr) or elementwise target: kind 0 = MSELoss, 1 = BCEWithLogitsLoss (mean)
__global__ void __launch_bounds__(kLossThreads) critic_loss_kernel(const float *__restrict__ x, const float *__restrict__ tgt, long long n,
                                                                   float target_const, int kind, float grad_scale, float *__restrict__ dx,
                                                                   double *partial)
{
    __shared__ double sh[32];
    double acc = 0.0;
    const float gs = grad_scale / (float)n;
    const uintptr_t align = reinterpret_cast<uintptr_t>(x) | (dx ? reinterpret_cast<uintptr_t>(dx) : 0) | (tgt ? reinterpret_cast<uintptr_t>(tgt) : 0);
    const long long n4 = (align & 15) == 0 ? n >> 2 : 0;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    const float4 *t4 = reinterpret_cast<const float4 *>(tgt);
    float4 *d4 = reinterpret_cast<float4 *>(dx);
    auto one = [&](float v, float target, float &g) {
        if (kind == 0) {
            float d = v - target;
            acc += (double)(d * d);
            g = 2.f * d * gs;
        } else {
            float e = expf(-fabsf(v));
            acc += (double)(fmaxf(v, 0.f) - v * target + log1pf(e));
            float sig = v >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
            g = (sig - target) * gs;
        }
    };
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldg(x4 + i), g;
        float4 t = tgt ? __ldg(t4 + i) : make_float4(target_const, target_const, target_const, target_const);
        one(v.x, t.x, g.x);
        one(v.y, t.y, g.y);
        one(v.z, t.z, g.z);
        one(v.w, t.w, g.w);
        if (dx) d4[i] = g;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float g;
        one(__ldg(x + i), tgt ? __ldg(tgt + i) : target_const, g);
        if (dx) dx[i] = g;
    }
    double s = block_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// y[i] *= *s  (gradient of a loss term scaled by the upstream gradient, a device scalar)
__global__ void __launch_bounds__(kLossThreads) scale_by_scalar_kernel(float *__restrict__ y, long long n, const float *__restrict__ s)
{
    const float f = __ldg(s);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] *= f;
}

static int loss_grid(long long work_items)
{
    long long b = cdiv(work_items, kLossThreads);
    long long cap = (long long)num_sms() * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace hn

using namespace hn;

extern "C" int64_t hn_loss_scratch_bytes(void) { return (int64_t)sizeof(double) * (2 * (num_sms() * 8 + 64) + 8); }

extern "C" int hn_ce_loss_fwd_bwd(const float *logits_nchw, const int64_t *labels, int64_t n_images, int32_t k, int64_t hw, int64_t ignore_index,
                                  float grad_scale, float *loss_out, float *dlogits_nchw, int32_t *flags, void *scratch, int64_t scratch_bytes,
                                  void *stream)
{
    HN_CHECK_ARG(logits_nchw && labels && loss_out && scratch && flags, "hn_ce_loss_fwd_bwd: null pointer");
    HN_CHECK_ARG(k >= 1 && k <= 32, "hn_ce_loss_fwd_bwd: 1 <= K <= 32 classes supported (got %d)", k);
    HN_CHECK_ARG(n_images >= 0 && hw >= 0, "hn_ce_loss_fwd_bwd: negative size");
    HN_CHECK_ARG(scratch_bytes >= hn_loss_scratch_bytes(), "hn_ce_loss_fwd_bwd: scratch too small (%lld < %lld)", (long long)scratch_bytes,
                 (long long)hn_loss_scratch_bytes());
    cudaStream_t st = (cudaStream_t)stream;
    const long long npix = n_images * hw;
    double *partial = reinterpret_cast<double *>(scratch);
    const int cap = num_sms() * 8 + 64;
    double *count = partial + 2 * cap;        // [0] = number of live labels, [1] = loss (FP64)
    const int g = loss_grid(npix);
    ce_count_kernel<<<g, kLossThreads, 0, st>>>(reinterpret_cast<const long long *>(labels), npix, k, ignore_index, partial, flags);
    HN_LAUNCH_CHECK();
    loss_fold_kernel<<<1, 256, 0, st>>>(partial, g, nullptr, 1.0, count, nullptr);
    HN_LAUNCH_CHECK();
    if (k <= 16)
        ce_fwd_bwd_kernel<16><<<g, kLossThreads, 0, st>>>(logits_nchw, reinterpret_cast<const long long *>(labels), n_images, hw, k, ignore_index,
                                                            count, 0.0, grad_scale, dlogits_nchw, partial + cap);
    else
        ce_fwd_bwd_kernel<32><<<g, kLossThreads, 0, st>>>(logits_nchw, reinterpret_cast<const long long *>(labels), n_images, hw, k, ignore_index,
                                                            count, 0.0, grad_scale, dlogits_nchw, partial + cap);
    HN_LAUNCH_CHECK();
    loss_fold_kernel<<<1, 256, 0, st>>>(partial + cap, g, count, 0.0, count + 1, loss_out);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_critic_loss_fwd_bwd(const float *x, const float *target, float target_const, int64_t n, int32_t kind, float grad_scale,
                                      float *loss_out, float *dx, void *scratch, int64_t scratch_bytes, void *stream)
{
    HN_CHECK_ARG(x && loss_out && scratch, "hn_critic_loss_fwd_bwd: null pointer");
    HN_CHECK_ARG(kind == 0 || kind == 1, "hn_critic_loss_fwd_bwd: kind must be 0 (MSE) or 1 (BCE with logits)");
    HN_CHECK_ARG(n >= 1, "hn_critic_loss_fwd_bwd: empty input");
    HN_CHECK_ARG(scratch_bytes >= hn_loss_scratch_bytes(), "hn_critic_loss_fwd_bwd: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    double *partial = reinterpret_cast<double *>(scratch);
    const int g = loss_grid(cdiv(n, 4));
    critic_loss_kernel<<<g, kLossThreads, 0, st>>>(x, target, n, target_const, kind, grad_scale, dx, partial);
    HN_LAUNCH_CHECK();
    loss_fold_kernel<<<1, 256, 0, st>>>(partial, g, nullptr, (double)n, nullptr, loss_out);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_scale_by_scalar(float *y, int64_t n, const float *scalar_dev, void *stream)
{
    HN_CHECK_ARG(y && scalar_dev && n >= 0, "hn_scale_by_scalar: bad arguments");
    if (n == 0) return HN_OK;
    scale_by_scalar_kernel<<<loss_grid(n), kLossThreads, 0, (cudaStream_t)stream>>>(y, n, scalar_dev);
    HN_LAUNCH_CHECK();
    return HN_OK;
}
