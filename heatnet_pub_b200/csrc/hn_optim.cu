// Multi-tensor optimiser step of the trainers on this path: RMSprop (cm/train_trgb_segnet_conf.py:270) and Adam
// (scripts/main.py:159), with the gradient-norm clip of scripts/main.py:256-257 and the 1/world gradient average of the
// data-parallel all-reduce folded into the same pass.  ONE launch updates every parameter tensor of the model (548 tensors
// for conv_segnet): a device table of slots {param, grad, state...} plus a block -> (slot, chunk) map replaces ~10 small
// torch kernels per tensor.  HBM-bound: RMSprop reads p, g, v and writes p, v (20 B/parameter); Adam 28 B/parameter.
#include "hn_common.cuh"

namespace hn {

constexpr int kOptThreads = 256;
constexpr int kOptChunk = 16384;  // elements per block: 64 per thread, 16 float4 iterations

struct OptHyper {
    float lr, a, b, eps, weight_decay, momentum;   // RMSprop: a = alpha.  Adam: a = beta1, b = beta2
    float bias_c1, bias_c2_sqrt;                   // Adam: 1 - beta1^t, sqrt(1 - beta2^t)
    float max_norm;                                // > 0: clip by total gradient norm read from *sqnorm
    const double *sqnorm;                          // device: sum of squares of all gradients (before grad_scale)
    float grad_scale;                              // constant factor on every gradient (1/world after an all-reduce SUM)
};

__device__ __forceinline__ float clip_coef(const OptHyper &h)
{
    float s = h.grad_scale;
    if (h.max_norm > 0.f && h.sqnorm) {
        // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), applied when < 1
        float total = (float)sqrt(*h.sqnorm) * fabsf(h.grad_scale);
        float coef = h.max_norm / (total + 1e-6f);
        if (coef < 1.f) s *= coef;
    }
    return s;
}

template <int ALGO>  // 0 = RMSprop, 1 = Adam
__device__ __forceinline__ void update_one(float &p, float g, float &s1, float &s2, const OptHyper &h)
{
    if (h.weight_decay != 0.f) g = fmaf(h.weight_decay, p, g);
    if (ALGO == 0) {
        // torch.optim.RMSprop (centered=False): v = alpha v + (1-alpha) g^2; p -= lr * g / (sqrt(v) + eps)   [momentum: buf = m buf + g/avg]
        s1 = h.a * s1 + (1.f - h.a) * g * g;
        float avg = sqrtf(s1) + h.eps;
        if (h.momentum > 0.f) {
            s2 = h.momentum * s2 + g / avg;
            p -= h.lr * s2;
        } else {
            p -= h.lr * (g / avg);
        }
    } else {
        // torch.optim.Adam: m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
        s1 = s1 + (1.f - h.a) * (g - s1);
        s2 = h.b * s2 + (1.f - h.b) * g * g;
        float denom = sqrtf(s2) / h.bias_c2_sqrt + h.eps;
        p -= (h.lr / h.bias_c1) * (s1 / denom);
    }
}

template <int ALGO>
__global__ void __launch_bounds__(kOptThreads) optim_step_kernel(const hn_param_slot *__restrict__ slots, const int2 *__restrict__ block_map, OptHyper h)
{
    const int2 bm = block_map[blockIdx.x];
    const hn_param_slot s = slots[bm.x];
    const long long begin = (long long)bm.y * kOptChunk;
    const long long end = begin + kOptChunk < s.numel ? begin + kOptChunk : s.numel;
    const float gscale = clip_coef(h);
    const bool two = ALGO == 1 || h.momentum > 0.f;
    float *p = s.param, *s1 = s.state1, *s2 = s.state2;
    const float *g = s.grad;
    const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(s1) |
                           (two ? reinterpret_cast<uintptr_t>(s2) : 0)) & 15) == 0;
    long long i = begin;
    if (aligned) {
        const long long vend = begin + ((end - begin) & ~3LL);
        for (i = begin + 4 * threadIdx.x; i < vend; i += 4 * kOptThreads) {
            float4 pv = *reinterpret_cast<float4 *>(p + i), gv = *reinterpret_cast<const float4 *>(g + i);
            float4 av = *reinterpret_cast<float4 *>(s1 + i), bv = two ? *reinterpret_cast<float4 *>(s2 + i) : make_float4(0, 0, 0, 0);
            update_one<ALGO>(pv.x, gv.x * gscale, av.x, bv.x, h);
            update_one<ALGO>(pv.y, gv.y * gscale, av.y, bv.y, h);
            update_one<ALGO>(pv.z, gv.z * gscale, av.z, bv.z, h);
            update_one<ALGO>(pv.w, gv.w * gscale, av.w, bv.w, h);
            *reinterpret_cast<float4 *>(p + i) = pv;
            *reinterpret_cast<float4 *>(s1 + i) = av;
            if (two) *reinterpret_cast<float4 *>(s2 + i) = bv;
        }
        i = vend;
    }
    for (i += threadIdx.x; i < end; i += kOptThreads) {
        float pv = p[i], av = s1[i], bv = two ? s2[i] : 0.f;
        update_one<ALGO>(pv, g[i] * gscale, av, bv, h);
        p[i] = pv;
        s1[i] = av;
        if (two) s2[i] = bv;
    }
}

// partial[block] = sum of squares of the block's gradient chunk (FP64); folded in block order afterwards
__global__ void __launch_bounds__(kOptThreads) grad_sqnorm_kernel(const hn_param_slot *__restrict__ slots, const int2 *__restrict__ block_map,
                                                                  double *__restrict__ partial)
{
    __shared__ double sh[kOptThreads / 32];
    const int2 bm = block_map[blockIdx.x];
    const hn_param_slot s = slots[bm.x];
    const long long begin = (long long)bm.y * kOptChunk;
    const long long end = begin + kOptChunk < s.numel ? begin + kOptChunk : s.numel;
    double acc = 0.0;
    for (long long i = begin + threadIdx.x; i < end; i += kOptThreads) {
        float v = s.grad[i];
        acc += (double)v * (double)v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kOptThreads / 32; ++w) t += sh[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void sqnorm_fold_kernel(const double *__restrict__ partial, int n, double *out)
{
    __shared__ double sh[256];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += partial[i];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

}  // namespace hn

using namespace hn;

extern "C" int32_t hn_optim_chunk(void) { return kOptChunk; }

extern "C" int hn_grad_sqnorm(const hn_param_slot *slots_dev, const int32_t *block_map_dev, int32_t n_blocks, double *partial_dev,
                              double *sqnorm_dev, void *stream)
{
    HN_CHECK_ARG(slots_dev && block_map_dev && partial_dev && sqnorm_dev && n_blocks >= 1, "hn_grad_sqnorm: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    grad_sqnorm_kernel<<<n_blocks, kOptThreads, 0, st>>>(slots_dev, reinterpret_cast<const int2 *>(block_map_dev), partial_dev);
    HN_LAUNCH_CHECK();
    sqnorm_fold_kernel<<<1, 256, 0, st>>>(partial_dev, n_blocks, sqnorm_dev);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

static int fill_common(OptHyper &h, float lr, float eps, float weight_decay, float max_norm, const double *sqnorm_dev, float grad_scale)
{
    HN_CHECK_ARG(lr >= 0.f && eps >= 0.f && weight_decay >= 0.f, "optimizer: negative lr / eps / weight_decay");
    HN_CHECK_ARG(!(max_norm > 0.f) || sqnorm_dev, "optimizer: max_norm needs the squared gradient norm (hn_grad_sqnorm)");
    h.lr = lr;
    h.eps = eps;
    h.weight_decay = weight_decay;
    h.max_norm = max_norm;
    h.sqnorm = sqnorm_dev;
    h.grad_scale = grad_scale;
    return HN_OK;
}

extern "C" int hn_rmsprop_step(const hn_param_slot *slots_dev, const int32_t *block_map_dev, int32_t n_blocks, float lr, float alpha, float eps,
                               float weight_decay, float momentum, float grad_scale, float max_norm, const double *sqnorm_dev, void *stream)
{
    HN_CHECK_ARG(slots_dev && block_map_dev && n_blocks >= 1, "hn_rmsprop_step: bad arguments");
    HN_CHECK_ARG(alpha >= 0.f && momentum >= 0.f, "hn_rmsprop_step: negative alpha / momentum");
    OptHyper h = {};
    int rc = fill_common(h, lr, eps, weight_decay, max_norm, sqnorm_dev, grad_scale);
    if (rc) return rc;
    h.a = alpha;
    h.momentum = momentum;
    optim_step_kernel<0><<<n_blocks, kOptThreads, 0, (cudaStream_t)stream>>>(slots_dev, reinterpret_cast<const int2 *>(block_map_dev), h);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_adam_step(const hn_param_slot *slots_dev, const int32_t *block_map_dev, int32_t n_blocks, float lr, float beta1, float beta2,
                            float eps, float weight_decay, int64_t step, float grad_scale, float max_norm, const double *sqnorm_dev, void *stream)
{
    HN_CHECK_ARG(slots_dev && block_map_dev && n_blocks >= 1 && step >= 1, "hn_adam_step: bad arguments");
    HN_CHECK_ARG(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, "hn_adam_step: betas must be in [0,1)");
    OptHyper h = {};
    int rc = fill_common(h, lr, eps, weight_decay, max_norm, sqnorm_dev, grad_scale);
    if (rc) return rc;
    h.a = beta1;
    h.b = beta2;
    h.bias_c1 = (float)(1.0 - pow((double)beta1, (double)step));
    h.bias_c2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    optim_step_kernel<1><<<n_blocks, kOptThreads, 0, (cudaStream_t)stream>>>(slots_dev, reinterpret_cast<const int2 *>(block_map_dev), h);
    HN_LAUNCH_CHECK();
    return HN_OK;
}
