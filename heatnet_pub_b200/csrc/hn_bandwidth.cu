// HBM-bound kernels of the HeatNet hot path on NHWC views: layout conversion, weight packing, BN fold /
// statistics / apply, max-pool, pyramid adaptive average pool, bilinear resize.  Every kernel moves
// 8 channels (16 B of BF16 / 32 B of FP32) per thread so that a warp touches contiguous 512 B / 1 KB
// segments; grids are sized in whole waves of the SM count with grid-stride loops.
#include <type_traits>

#include <stdlib.h>

#include "hn_common.cuh"

namespace hn {

// (rows, chunks-per-row) grid for the row-decomposed kernels: rows in grid.x (up to 2^31-1), at most 4 items per thread
static inline dim3 row_grid(int64_t rows, int64_t items_per_row, int threads = 256)
{
    int64_t chunks = cdiv(items_per_row, (int64_t)threads * 4);
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) chunks = 65535;
    return dim3((unsigned)rows, (unsigned)chunks);
}

static inline int wave_grid(int64_t work_items, int threads, int waves_per_sm = 8)
{
    int64_t want = cdiv(work_items, threads);
    int64_t cap = (int64_t)num_sms() * waves_per_sm;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

// ------------------------------------------------------------------------------------------------
// layout conversion
// ------------------------------------------------------------------------------------------------
// small-C path (network inputs: C = 3 / 1 / 4): one thread per pixel
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_smallc(const float *__restrict__ src, T *__restrict__ dst, int64_t npix_per_img,
                                                           int64_t n_img, int C, int ld)
{
    const int64_t total = npix_per_img * n_img;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t n = i / npix_per_img, p = i - n * npix_per_img;
        const float *s = src + n * C * npix_per_img + p;
        T *d = dst + i * ld;
        for (int c = 0; c < C; ++c) d[c] = from_f32<T>(__ldg(s + (int64_t)c * npix_per_img));
    }
}

// C <= 16 (the 13-class logits / their gradient on the way back into the NHWC engine): all plane loads of a pixel are issued before
// the first store, and the stores are the widest aligned vectors that stay INSIDE the C valid channels (the pad channels of a
// channel-slice view may belong to a neighbour).  `vec` = the destination's pixel stride and base allow 8-byte (BF16) / 16-byte
// (FP32) stores.  Replaces 32x32 tile transposes that ran with 13 of 32 channel rows populated (0.38 -> ~0.15 ms per pass).
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_c16(const float *__restrict__ src, T *__restrict__ dst, int64_t npix_per_img, int64_t n_img,
                                                        int C, int ld, int vec)
{
    const int64_t total = npix_per_img * n_img;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t n = i / npix_per_img, p = i - n * npix_per_img;
        const float *s = src + n * C * npix_per_img + p;
        float v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = c < C ? __ldg(s + (int64_t)c * npix_per_img) : 0.f;
        T *d = dst + i * ld;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
            if (c + 4 <= C && vec) {
                if constexpr (sizeof(T) == 4) {
                    *reinterpret_cast<float4 *>(reinterpret_cast<float *>(d) + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                } else {
                    uint2 q;
                    __nv_bfloat162 a = __floats2bfloat162_rn(v[c], v[c + 1]), b = __floats2bfloat162_rn(v[c + 2], v[c + 3]);
                    q.x = *reinterpret_cast<uint32_t *>(&a);
                    q.y = *reinterpret_cast<uint32_t *>(&b);
                    *reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(d) + c) = q;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j < C) d[c + j] = from_f32<T>(v[c + j]);
            }
        }
    }
}

// small-C NHWC -> NCHW (the logits: C = 13 in a 16-wide buffer): one thread per pixel reads its channel vector
// (contiguous), writes C coalesced planes
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_smallc(const T *__restrict__ src, float *__restrict__ dst, int64_t npix_per_img,
                                                           int64_t n_img, int C, int ld)
{
    const int64_t total = npix_per_img * n_img;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t n = i / npix_per_img, p = i - n * npix_per_img;
        const T *s = src + i * ld;
        float v[16];
        if (sizeof(T) == 4 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
            for (int c = 0; c < 16; c += 4)
                if (c < C) {
                    float4 q = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(s) + c);
                    v[c] = q.x; v[c + 1] = q.y; v[c + 2] = q.z; v[c + 3] = q.w;
                }
        } else {
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (c < C) v[c] = to_f32<T>(s[c]);
        }
        float *d = dst + n * C * npix_per_img + p;
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (c < C) d[(int64_t)c * npix_per_img] = v[c];
    }
}

// general path: 32x32 tile transpose through shared memory.  grid = (pixel tiles, channel tiles, N)
template <typename T, bool kToNHWC>
__global__ void __launch_bounds__(256) transpose_tiles(const void *__restrict__ src_, void *__restrict__ dst_, int64_t HW, int C,
                                                       int ld)
{
    __shared__ float tile[32][33];
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int64_t n = blockIdx.z;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    if (kToNHWC) {
        const float *src = (const float *)src_ + n * C * HW;
        T *dst = (T *)dst_ + n * HW * ld;
        for (int j = ty; j < 32; j += 8)
            if (c0 + j < C && p0 + tx < HW) tile[j][tx] = src[(int64_t)(c0 + j) * HW + p0 + tx];
        __syncthreads();
        for (int j = ty; j < 32; j += 8)
            if (p0 + j < HW && c0 + tx < C) dst[(p0 + j) * ld + c0 + tx] = from_f32<T>(tile[tx][j]);
    } else {
        const T *src = (const T *)src_ + n * HW * ld;
        float *dst = (float *)dst_ + n * C * HW;
        for (int j = ty; j < 32; j += 8)
            if (p0 + j < HW && c0 + tx < C) tile[j][tx] = to_f32<T>(src[(p0 + j) * ld + c0 + tx]);
        __syncthreads();
        for (int j = ty; j < 32; j += 8)
            if (c0 + j < C && p0 + tx < HW) dst[(int64_t)(c0 + j) * HW + p0 + tx] = tile[tx][j];
    }
}

// OIHW FP32 -> [cout_pad][kpad], k = (r*S+s)*Cin + c
template <typename T>
__global__ void __launch_bounds__(256) pack_weight_kernel(const float *__restrict__ w, const float *__restrict__ row_scale,
                                                          T *__restrict__ dst, int cout, int cin, int R, int S, int cout_pad, int kpad)
{
    const int64_t total = (int64_t)cout_pad * kpad;
    const int K = R * S * cin;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int o = (int)(i / kpad), k = (int)(i - (int64_t)o * kpad);
        float v = 0.f;
        if (o < cout && k < K) {
            int tap = k / cin, c = k - tap * cin;
            int r = tap / S, s = tap - r * S;
            v = __ldg(w + (((int64_t)o * cin + c) * R + r) * S + s);
            if (row_scale) v *= __ldg(row_scale + o);      // BatchNorm2d(eval) scale folded into the filter
        }
        dst[i] = from_f32<T>(v);
    }
}

__global__ void bn_fold_kernel(const float *gamma, const float *beta, const float *mean, const float *var, const float *bias,
                               float eps, float *scale, float *shift, int C)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float b = bias ? bias[c] : 0.f;
    if (gamma) {
        float sc = gamma[c] / sqrtf(var[c] + eps);   // matches ATen: weight * rsqrt(var+eps) within 1 ulp
        scale[c] = sc;
        shift[c] = beta[c] + (b - mean[c]) * sc;
    } else {
        scale[c] = 1.f;
        shift[c] = b;
    }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm2d train-mode statistics / finalize, and the fused apply pass
// ------------------------------------------------------------------------------------------------
// blockDim = (CVB, PL): CVB 8-channel vectors x PL pixel lanes.  grid = (pixel chunks, cvec blocks)
// Optional fused finalize (hn_bn_batch_stats): the LAST CTA to finish (ticket counter) turns the sums into the BatchNorm2d
// scale/shift, the saved mean/invstd and the running-statistic update, so a train-mode BN forward is memset + ONE kernel.
struct BnFinalize {
    unsigned int *ticket;      // nullptr: plain statistics
    double count;
    const float *gamma, *beta;
    float eps, momentum;
    float *running_mean, *running_var;
    long long *num_batches_tracked;
    float *scale, *shift, *save_mean, *save_invstd;
};

__device__ __forceinline__ void bn_finalize_channel(int c, double s, double q, const BnFinalize &f)
{
    double mean = s / f.count;
    double var = q / f.count - mean * mean;
    if (var < 0) var = 0;
    double invstd = 1.0 / sqrt(var + (double)f.eps);
    float g = f.gamma ? f.gamma[c] : 1.f, b = f.beta ? f.beta[c] : 0.f;
    f.scale[c] = (float)(g * invstd);
    f.shift[c] = (float)(b - mean * g * invstd);
    if (f.save_mean) f.save_mean[c] = (float)mean;
    if (f.save_invstd) f.save_invstd[c] = (float)invstd;
    if (f.running_mean) {
        double unbiased = f.count > 1 ? var * f.count / (f.count - 1.0) : var;
        f.running_mean[c] = (float)((1.0 - f.momentum) * f.running_mean[c] + f.momentum * mean);
        f.running_var[c] = (float)((1.0 - f.momentum) * f.running_var[c] + f.momentum * unbiased);
    }
}

template <typename T>
__global__ void channel_stats_kernel(const T *__restrict__ x, int64_t npix, int C, int ld, int64_t pix_per_cta, double *sum,
                                     double *sqsum, BnFinalize fin)
{
    extern __shared__ float red[];  // [2][PL][CVB*8]
    const int cv = blockIdx.y * blockDim.x + threadIdx.x;
    const int ncv = C / 8;
    const int PL = blockDim.y, CVB = blockDim.x;
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    const int64_t p_begin = (int64_t)blockIdx.x * pix_per_cta;
    const int64_t p_end = min(p_begin + pix_per_cta, npix);
    if (cv < ncv) {
        int64_t p = p_begin + threadIdx.y;
        for (; p + 3 * PL < p_end; p += 4 * PL) {          // four pixels (up to 128 B per thread) in flight
            float v[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) Vec8<T>::load(x + (p + u * PL) * ld + cv * 8, v[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    s[i] += v[u][i];
                    q[i] = fmaf(v[u][i], v[u][i], q[i]);
                }
        }
        for (; p < p_end; p += PL) {
            float v[8];
            Vec8<T>::load(x + p * ld + cv * 8, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s[i] += v[i];
                q[i] = fmaf(v[i], v[i], q[i]);
            }
        }
    }
    float *rs = red, *rq = red + PL * CVB * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        rs[(threadIdx.y * CVB + threadIdx.x) * 8 + i] = s[i];
        rq[(threadIdx.y * CVB + threadIdx.x) * 8 + i] = q[i];
    }
    __syncthreads();
    // threads of pixel-lane 0..: each (threadIdx.x, i) column reduced by one thread
    const int tid = threadIdx.y * CVB + threadIdx.x;
    for (int col = tid; col < CVB * 8; col += CVB * PL) {
        double a = 0.0, b = 0.0;
        for (int l = 0; l < PL; ++l) {
            a += (double)rs[l * CVB * 8 + col];
            b += (double)rq[l * CVB * 8 + col];
        }
        int ch = blockIdx.y * CVB * 8 + col;
        if (ch < C) {
            atomicAdd(sum + ch, a);
            atomicAdd(sqsum + ch, b);
        }
    }
    if (fin.ticket) {
        __shared__ bool last;
        __threadfence();
        __syncthreads();
        if (tid == 0) last = atomicAdd(fin.ticket, 1u) == gridDim.x * gridDim.y - 1;
        __syncthreads();
        if (last) {
            __threadfence();
            for (int c = tid; c < C; c += CVB * PL) bn_finalize_channel(c, __ldcg(sum + c), __ldcg(sqsum + c), fin);
            if (tid == 0 && fin.num_batches_tracked) *fin.num_batches_tracked += 1;
        }
    }
}

// Thread-fixed-channel variant of the normalise pass (per-channel scale AND shift, not per image): blockDim = (CVB, PL) as in
// channel_stats_kernel; a thread keeps one 8-channel group, so scale / shift live in registers and the pixel loop only
// streams x (+ residual) in and y out, two pixels in flight.  Same fmaf(x, scale, shift) + residual + activation expression.
template <typename TI, typename T>
__global__ void __launch_bounds__(256) affine_act_fixed_kernel(const TI *__restrict__ x, int ldx, const float *__restrict__ scale,
                                                               const float *__restrict__ shift, const T *__restrict__ res, int ldr, int act,
                                                               float slope, const float *slope_ptr, T *__restrict__ y, int ldy, int64_t npix,
                                                               int C, int64_t pix_per_cta)
{
    const int cv = blockIdx.y * blockDim.x + threadIdx.x;
    if (cv >= C / 8) return;
    const int PL = blockDim.y, c = cv * 8;
    if (slope_ptr) slope = __ldg(slope_ptr);
    float sc[8], sh[8];
    Vec8<float>::load(scale + c, sc);
    Vec8<float>::load(shift + c, sh);
    const int64_t p_begin = (int64_t)blockIdx.x * pix_per_cta;
    const int64_t p_end = min(p_begin + pix_per_cta, npix);
    auto one = [&](int64_t p, float (&v)[8], const float (&r)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[j] = fmaf(v[j], sc[j], sh[j]);
            if (res) v[j] += r[j];
            v[j] = apply_act(v[j], act, slope);
        }
        Vec8<T>::store(y + p * ldy + c, v);
    };
    int64_t p = p_begin + threadIdx.y;
    for (; p + PL < p_end; p += 2 * PL) {
        float v0[8], v1[8], r0[8], r1[8];
        Vec8<TI>::load(x + p * ldx + c, v0);
        Vec8<TI>::load(x + (p + PL) * ldx + c, v1);
        if (res) {
            Vec8<T>::load(res + p * ldr + c, r0);
            Vec8<T>::load(res + (p + PL) * ldr + c, r1);
        }
        one(p, v0, r0);
        one(p + PL, v1, r1);
    }
    if (p < p_end) {
        float v0[8], r0[8];
        Vec8<TI>::load(x + p * ldx + c, v0);
        if (res) Vec8<T>::load(res + p * ldr + c, r0);
        one(p, v0, r0);
    }
}

// Train-mode BatchNorm2d finalize + normalise in ONE launch: the per-channel scale / shift are formed in every thread's prologue
// from the FP64 sums the conv epilogue accumulated (same arithmetic as bn_finalize_kernel), the threads of the first pixel chunk
// also write the vectors the backward needs (scale, shift, mean, invstd) and update the running statistics.  Removes one tiny
// launch per BN layer (158 per training step).
struct BnTrainFin {
    const double *sum, *sqsum;      // [groups][C] each
    int groups;                     // image groups with separate statistics (blockIdx.z = group); scale .. save_invstd are [groups][C]
    int64_t group_pix;              // pixels per group
    double count;
    const float *gamma, *beta;
    float eps, momentum;
    float *running_mean, *running_var;
    long long *num_batches_tracked;
    float *scale, *shift, *save_mean, *save_invstd;
};

// blockDim = (CVB, PL): a thread keeps ONE 4-channel group (scale / shift in 8 registers) and streams its pixels four at a
// time (all loads of a batch issued before the first use); 64 registers -> 4 CTAs per SM, i.e. >= 64 KB of loads in flight per SM.
template <typename TI, typename T>
__global__ void __launch_bounds__(256, 4) bn_apply_train_kernel(const TI *__restrict__ x, int ldx, const BnTrainFin fin, const T *__restrict__ res,
                                                                int ldr, int act, float slope, const float *slope_ptr, T *__restrict__ y, int ldy,
                                                                int64_t npix, int C, int64_t pix_per_cta)
{
    // per-channel scale / shift of this CTA's channel block: one thread per channel (FP64), shared with the other pixel lanes
    __shared__ float s_sc[128], s_sh[128];
    const int PL = blockDim.y, CVB = blockDim.x;
    const int tid = threadIdx.y * CVB + threadIdx.x;
    const int grp = blockIdx.z;
    if (blockIdx.x == 0 && blockIdx.y == 0 && grp == 0 && tid == 0 && fin.num_batches_tracked) *fin.num_batches_tracked += fin.groups;
    if (tid < CVB * 4) {
        const int ch = blockIdx.y * CVB * 4 + tid;
        if (ch < C) {
            const double mean = fin.sum[grp * C + ch] / fin.count;
            double var = fin.sqsum[grp * C + ch] / fin.count - mean * mean;
            if (var < 0) var = 0;
            const double invstd = 1.0 / sqrt(var + (double)fin.eps);
            const float g = fin.gamma ? fin.gamma[ch] : 1.f, b = fin.beta ? fin.beta[ch] : 0.f;
            const float scv = (float)(g * invstd), shv = (float)(b - mean * g * invstd);
            s_sc[tid] = scv;
            s_sh[tid] = shv;
            if (blockIdx.x == 0) {
                fin.scale[grp * C + ch] = scv;
                fin.shift[grp * C + ch] = shv;
                fin.save_mean[grp * C + ch] = (float)mean;
                fin.save_invstd[grp * C + ch] = (float)invstd;
                if (fin.running_mean && grp == 0) {
                    // the groups are consecutive forward calls of the reference (day batch, then night batch): one momentum
                    // update per group, in order, by the same thread
                    float rm = fin.running_mean[ch], rv = fin.running_var[ch];
                    for (int q = 0; q < fin.groups; ++q) {
                        const double mq = fin.sum[q * C + ch] / fin.count;
                        double vq = fin.sqsum[q * C + ch] / fin.count - mq * mq;
                        if (vq < 0) vq = 0;
                        const double unbiased = fin.count > 1 ? vq * fin.count / (fin.count - 1.0) : vq;
                        rm = (float)((1.0 - fin.momentum) * rm + fin.momentum * mq);
                        rv = (float)((1.0 - fin.momentum) * rv + fin.momentum * unbiased);
                    }
                    fin.running_mean[ch] = rm;
                    fin.running_var[ch] = rv;
                }
            }
        }
    }
    __syncthreads();
    const int cv = blockIdx.y * CVB + threadIdx.x;
    if (cv >= C / 4) return;
    const int c = cv * 4;
    if (slope_ptr) slope = __ldg(slope_ptr);
    float sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sc[j] = s_sc[threadIdx.x * 4 + j];
        sh[j] = s_sh[threadIdx.x * 4 + j];
    }
    const int64_t g_begin = (int64_t)grp * fin.group_pix;                   // this group's pixel range (npix = pixels per group)
    const int64_t p_begin = g_begin + (int64_t)blockIdx.x * pix_per_cta;
    const int64_t p_end = min(p_begin + pix_per_cta, g_begin + npix);
    auto one = [&](int64_t p, float (&v)[4], const float (&r)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = fmaf(v[j], sc[j], sh[j]);
            if (res) v[j] += r[j];
            v[j] = apply_act(v[j], act, slope);
        }
        Vec4<T>::store(y + p * ldy + c, v);
    };
    int64_t p = p_begin + threadIdx.y;
    for (; p + 3 * PL < p_end; p += 4 * PL) {
        typename Vec4<TI>::raw_t xv[4];
        typename Vec4<T>::raw_t rv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) xv[u] = Vec4<TI>::ldraw(x + (p + u * PL) * ldx + c);
        if (res) {
#pragma unroll
            for (int u = 0; u < 4; ++u) rv[u] = Vec4<T>::ldraw(res + (p + u * PL) * ldr + c);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float v[4], r[4];
            Vec4<TI>::unpack(xv[u], v);
            if (res) Vec4<T>::unpack(rv[u], r);
            one(p + u * PL, v, r);
        }
    }
    for (; p < p_end; p += PL) {
        float v0[4], r0[4];
        Vec4<TI>::load(x + p * ldx + c, v0);
        if (res) Vec4<T>::load(res + p * ldr + c, r0);
        one(p, v0, r0);
    }
}

// narrow / unaligned views (C = 13 logits, C = 1 critic maps): grid = (pixel chunks, C), block-wide reduction per channel
template <typename T>
__global__ void __launch_bounds__(256) channel_stats_scalar_kernel(const T *__restrict__ x, int64_t npix, int ld, int64_t pix_per_cta,
                                                                   double *sum, double *sqsum)
{
    __shared__ double sa[8], sb[8];
    const int c = blockIdx.y;
    const int64_t p0 = (int64_t)blockIdx.x * pix_per_cta;
    const int64_t p1 = min(p0 + pix_per_cta, npix);
    double a = 0.0, b = 0.0;
    for (int64_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
        float v = to_f32<T>(x[p * ld + c]);
        a += v;
        b += (double)v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) { a += sa[i]; b += sb[i]; }
        atomicAdd(sum + c, a);
        atomicAdd(sqsum + c, b);
    }
}

__global__ void bn_finalize_kernel(const double *sum, const double *sqsum, double count, const float *gamma, const float *beta,
                                   float eps, float momentum, float *running_mean, float *running_var, float *scale, float *shift,
                                   float *save_mean, float *save_invstd, int C, long long *num_batches_tracked = nullptr)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
    if (c >= C) return;
    double mean = sum[c] / count;
    double var = sqsum[c] / count - mean * mean;
    if (var < 0) var = 0;
    double invstd = 1.0 / sqrt(var + (double)eps);
    float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    scale[c] = (float)(g * invstd);
    shift[c] = (float)(b - mean * g * invstd);
    if (save_mean) save_mean[c] = (float)mean;
    if (save_invstd) save_invstd[c] = (float)invstd;
    if (running_mean) {
        double unbiased = count > 1 ? var * count / (count - 1.0) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
}

__global__ void bn_count_kernel(long long *num_batches_tracked) { *num_batches_tracked += 1; }

template <typename TI, typename T>
__global__ void __launch_bounds__(256) affine_act_kernel(const TI *__restrict__ x, int ldx, const float *__restrict__ scale,
                                                         const float *__restrict__ shift, const T *__restrict__ res, int ldr, int act,
                                                         float slope, const float *slope_ptr, T *__restrict__ y, int ldy, int64_t npix,
                                                         int C, int64_t pix_per_image)
{
    const int ncv = C / 8;
    if (slope_ptr) slope = __ldg(slope_ptr);
    // a CTA walks blocks of kPixPerBlock pixels; inside a block the (pixel, 8-channel vector) split is 32-bit math
    constexpr int kPixPerBlock = 64;
    const int64_t nblocks = (npix + kPixPerBlock - 1) / kPixPerBlock;
    for (int64_t pb = blockIdx.x; pb < nblocks; pb += gridDim.x)
    for (int i = threadIdx.x; i < kPixPerBlock * ncv; i += blockDim.x) {
        const int pl = i / ncv;
        const int64_t p = pb * kPixPerBlock + pl;
        if (p >= npix) break;
        const int c = (i - pl * ncv) * 8;
        float v[8];
        Vec8<TI>::load(x + p * ldx + c, v);
        if (scale) {
            // pix_per_image > 0: per-(image, channel) parameters (Dropout2d masks)
            const int64_t poff = pix_per_image > 0 ? (p / pix_per_image) * C : 0;
            float sc[8], sh[8];
            Vec8<float>::load(scale + poff + c, sc);
            if (shift) {
                Vec8<float>::load(shift + poff + c, sh);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] *= sc[j];
            }
        }
        if (res) {
            float r[8];
            Vec8<T>::load(res + p * ldr + c, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += r[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], act, slope);
        Vec8<T>::store(y + p * ldy + c, v);
    }
}

// ------------------------------------------------------------------------------------------------
// nn.MaxPool2d(3, 2, 1) (pads with -inf)
// ------------------------------------------------------------------------------------------------
// grid.x = output rows (n*Ho + ho), grid.y = chunks of the row's Wo * C/8 work items: no 64-bit div/mod per element
template <typename T>
__global__ void __launch_bounds__(256) maxpool_kernel(const T *__restrict__ x, int ldx, T *__restrict__ y, int ldy, int N, int H, int W,
                                                      int Ho, int Wo, int C)
{
    const int ncv = C / 8;
    const int row = blockIdx.x;
    const int n = row / Ho, ho = row - n * Ho;
    const int items = Wo * ncv;
    const T *xin = x + (int64_t)n * H * W * ldx;
    T *yrow = y + (int64_t)row * Wo * ldy;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < items; i += gridDim.y * blockDim.x) {
        const int wo = i / ncv, c = (i - wo * ncv) * 8;
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            int hi = 2 * ho - 1 + r;
            if (hi < 0 || hi >= H) continue;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                int wi = 2 * wo - 1 + s;
                if (wi < 0 || wi >= W) continue;
                float v[8];
                Vec8<T>::load(xin + ((int64_t)hi * W + wi) * ldx + c, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
            }
        }
        Vec8<T>::store(yrow + (int64_t)wo * ldy + c, m);
    }
}

// ------------------------------------------------------------------------------------------------
// pyramid adaptive average pool: all sizes in one pass over x.
// stage 1: CTA = (n, input row); thread = one 8-channel vector; accumulates the row into every column bin
//          of every size (registers), writes row partials [N][H][ncols][C] FP32.
// stage 2: thread = (n, bin, 8-channel vector): adds the rows of the bin, divides by the bin area.
// Deterministic (no atomics).  Bin i of size s covers [floor(i*H/s), ceil((i+1)*H/s)).
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPoolCols = 12;  // sum of sizes, (1,2,3,6) -> 12
struct PoolSizes {
    int n, s[4];
    int ncols, w0[kMaxPoolCols], w1[kMaxPoolCols];  // column-bin bounds of every size, concatenated
    int nseg, segb[2 * kMaxPoolCols + 2];           // sorted union of all bounds: column segments, each inside or outside every bin
};

template <typename T>
__global__ void __launch_bounds__(256) pyramid_rows_kernel(const T *__restrict__ x, int ldx, int H, int W, int C,
                                                           const __grid_constant__ PoolSizes ps, float *__restrict__ rowpart)
{
    const int cv = blockIdx.y * blockDim.x + threadIdx.x;
    if (cv * 8 >= C) return;
    const int64_t nh = blockIdx.x;  // n*H + h
    const int ncols = ps.ncols;
    float acc[kMaxPoolCols][8];
#pragma unroll
    for (int b = 0; b < kMaxPoolCols; ++b)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[b][j] = 0.f;
    const T *row = x + nh * W * ldx + cv * 8;
    // walk the row segment by segment (warp-uniform bounds): 8 adds per load inside a segment, 4 loads in flight; a finished
    // segment is added to every bin that contains it (bins of one size overlap when W % s != 0)
    for (int sg = 0; sg < ps.nseg; ++sg) {
        const int s0 = ps.segb[sg], s1 = ps.segb[sg + 1];
        float seg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) seg[j] = 0.f;
        int w = s0;
        for (; w + 8 <= s1; w += 8) {                      // 8 loads in flight: the kernel is latency-bound (ncu: 17 % issue active)
            float v[8][8];
#pragma unroll
            for (int u = 0; u < 8; ++u) Vec8<T>::load(row + (int64_t)(w + u) * ldx, v[u]);
#pragma unroll
            for (int j = 0; j < 8; ++j) seg[j] += ((v[0][j] + v[1][j]) + (v[2][j] + v[3][j])) + ((v[4][j] + v[5][j]) + (v[6][j] + v[7][j]));
        }
        for (; w + 4 <= s1; w += 4) {
            float v[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) Vec8<T>::load(row + (int64_t)(w + u) * ldx, v[u]);
#pragma unroll
            for (int j = 0; j < 8; ++j) seg[j] += (v[0][j] + v[1][j]) + (v[2][j] + v[3][j]);
        }
        for (; w < s1; ++w) {
            float v[8];
            Vec8<T>::load(row + (int64_t)w * ldx, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) seg[j] += v[j];
        }
#pragma unroll
        for (int b = 0; b < kMaxPoolCols; ++b) {
            if (b < ncols && s0 >= ps.w0[b] && s1 <= ps.w1[b]) {
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[b][j] += seg[j];
            }
        }
    }
    float *out = rowpart + nh * ncols * C + cv * 8;
#pragma unroll
    for (int b = 0; b < kMaxPoolCols; ++b)
        if (b < ncols) Vec8<float>::store(out + (int64_t)b * C, acc[b]);
}

template <typename T>
__global__ void __launch_bounds__(256) pyramid_bins_kernel(const float *__restrict__ rowpart, int N, int H, int W, int C, PoolSizes ps,
                                                           T *__restrict__ out)
{
    int nbins = 0, ncols = 0;
    for (int a = 0; a < ps.n; ++a) {
        nbins += ps.s[a] * ps.s[a];
        ncols += ps.s[a];
    }
    const int ncv = C / 8;
    const int64_t total = (int64_t)N * nbins * ncv;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % ncv) * 8;
        int bin = (int)((i / ncv) % nbins);
        int n = (int)(i / ((int64_t)ncv * nbins));
        int a = 0, base = 0, colbase = 0;
        while (bin - base >= ps.s[a] * ps.s[a]) {
            base += ps.s[a] * ps.s[a];
            colbase += ps.s[a];
            ++a;
        }
        int s = ps.s[a], bi = (bin - base) / s, bj = (bin - base) % s;
        int h0 = (bi * H) / s, h1 = ((bi + 1) * H + s - 1) / s;
        int w0 = (bj * W) / s, w1 = ((bj + 1) * W + s - 1) / s;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int h = h0; h < h1; ++h) {
            float v[8];
            Vec8<float>::load(rowpart + (((int64_t)n * H + h) * ncols + colbase + bj) * C + c, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += v[j];
        }
        float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= inv;
        // per-size dense NHWC blocks: size a occupies [N][s*s][C] starting at pixel row N*base
        Vec8<T>::store(out + ((int64_t)base * N + (int64_t)n * s * s + (bin - base)) * C + c, acc);
    }
}

// ------------------------------------------------------------------------------------------------
// bilinear resize, align_corners = False (ATen upsample_bilinear2d index rule, computed in FP32)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int &i0, int &i1, float &l1)
{
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

// grid.x = output rows (n*Ho + ho), grid.y = chunks of the row's Wo * C/8 work items
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) bilinear_vec_kernel(const TI *__restrict__ x, int ldx, TO *__restrict__ y, int ldy, int N, int H,
                                                           int W, int Ho, int Wo, int C, float sh, float sw)
{
    const int ncv = C / 8;
    const int row = blockIdx.x;
    const int n = row / Ho, ho = row - n * Ho;
    int y0, y1;
    float ly;
    bilinear_src(ho, sh, H, y0, y1, ly);
    const float hy = 1.f - ly;
    const TI *r0 = x + ((int64_t)n * H + y0) * W * ldx;
    const TI *r1 = x + ((int64_t)n * H + y1) * W * ldx;
    TO *yrow = y + (int64_t)row * Wo * ldy;
    const int items = Wo * ncv;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < items; i += gridDim.y * blockDim.x) {
        const int wo = i / ncv, c = (i - wo * ncv) * 8;
        int x0, x1;
        float lx;
        bilinear_src(wo, sw, W, x0, x1, lx);
        float a00[8], a01[8], a10[8], a11[8], o[8];
        Vec8<TI>::load(r0 + (int64_t)x0 * ldx + c, a00);
        Vec8<TI>::load(r0 + (int64_t)x1 * ldx + c, a01);
        Vec8<TI>::load(r1 + (int64_t)x0 * ldx + c, a10);
        Vec8<TI>::load(r1 + (int64_t)x1 * ldx + c, a11);
        const float hx = 1.f - lx;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = hy * (hx * a00[j] + lx * a01[j]) + ly * (hx * a10[j] + lx * a11[j]);
        Vec8<TO>::store(yrow + (int64_t)wo * ldy + c, o);
    }
}

// exact 2x upsample (the three PSPUpsample stages): one thread produces the 2x2 output block of one input pixel from its
// 3x3 neighbourhood (9 loads for 4 outputs instead of 16) with the same FP32 expression as the generic kernel, so
// the results are bit-identical to it.  grid.x = input rows (n*H + y), grid.y = chunks of the row's W * C/8 items.
template <typename T>
__global__ void __launch_bounds__(256) bilinear_up2_kernel(const T *__restrict__ x, int ldx, T *__restrict__ y, int ldy, int N, int H, int W,
                                                           int C)
{
    const int ncv = C / 8;
    const int row = blockIdx.x;
    const int n = row / H, yi = row - n * H;
    const int ym = yi > 0 ? yi - 1 : 0, yp = yi < H - 1 ? yi + 1 : H - 1;
    const T *rm = x + ((int64_t)n * H + ym) * W * ldx;
    const T *r0 = x + ((int64_t)n * H + yi) * W * ldx;
    const T *rp = x + ((int64_t)n * H + yp) * W * ldx;
    const int Wo = 2 * W;
    T *o0 = y + ((int64_t)n * 2 * H + 2 * yi) * Wo * ldy;
    T *o1 = o0 + (int64_t)Wo * ldy;
    // output row 2y   : src = y - 0.25 -> rows (y-1, y), ly = 0.75 (clamped to row 0: ly = 0)
    // output row 2y+1 : src = y + 0.25 -> rows (y, y+1), ly = 0.25
    const float ly0 = yi > 0 ? 0.75f : 1.f, ly1 = 0.25f;   // yi == 0: all weight on tap 1 (= row 0 itself; tap 0 is the same clamped row)
    const int items = W * ncv;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < items; i += gridDim.y * blockDim.x) {
        const int xi = i / ncv, c = (i - xi * ncv) * 8;
        const int xm = xi > 0 ? xi - 1 : 0, xp = xi < W - 1 ? xi + 1 : W - 1;
        const float lx0 = xi > 0 ? 0.75f : 1.f, lx1 = 0.25f;
        float a[3][3][8];
        Vec8<T>::load(rm + (int64_t)xm * ldx + c, a[0][0]); Vec8<T>::load(rm + (int64_t)xi * ldx + c, a[0][1]); Vec8<T>::load(rm + (int64_t)xp * ldx + c, a[0][2]);
        Vec8<T>::load(r0 + (int64_t)xm * ldx + c, a[1][0]); Vec8<T>::load(r0 + (int64_t)xi * ldx + c, a[1][1]); Vec8<T>::load(r0 + (int64_t)xp * ldx + c, a[1][2]);
        Vec8<T>::load(rp + (int64_t)xm * ldx + c, a[2][0]); Vec8<T>::load(rp + (int64_t)xi * ldx + c, a[2][1]); Vec8<T>::load(rp + (int64_t)xp * ldx + c, a[2][2]);
        float o[8];
        // rows/cols are already clamped (ym/yp, xm/xp), so the tap pairs are static: (0,1) for the even output and (1,2) for
        // the odd one; at the low border the weight of the second tap is 0 and both taps are the same pixel
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const float ly = dy == 0 ? ly0 : ly1, hy = 1.f - ly;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const float lx = dx == 0 ? lx0 : lx1, hx = 1.f - lx;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    o[j] = hy * (hx * a[dy][dx][j] + lx * a[dy][dx + 1][j]) + ly * (hx * a[dy + 1][dx][j] + lx * a[dy + 1][dx + 1][j]);
                Vec8<T>::store((dy == 0 ? o0 : o1) + (int64_t)(2 * xi + dx) * ldy + c, o);
            }
        }
    }
}

// Column-walking variant of the exact 2x upsample: a thread owns one 8-channel vector of one input COLUMN position and walks
// RY consecutive input rows, keeping the 3x3 neighbourhood in registers as a sliding window -- 3 new loads per 4 stores instead
// of 9 (the L1/L2 load traffic, not HBM, was the bound of the 9-load version).  Same FP32 expression per output as
// bilinear_up2_kernel, so the results are bit-identical to it.  grid = (W*C/8 item chunks, row chunks, N).
template <typename T, int RY>
__global__ void __launch_bounds__(128) bilinear_up2_walk_kernel(const T *__restrict__ x, int ldx, T *__restrict__ y, int ldy, int N, int H,
                                                                int W, int C)
{
    const int ncv = C / 8;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= W * ncv) return;
    const int xi = item / ncv, c = (item - xi * ncv) * 8;
    const int n = blockIdx.z;
    const int y_begin = blockIdx.y * RY, y_end = y_begin + RY < H ? y_begin + RY : H;
    const int xm = xi > 0 ? xi - 1 : 0, xp = xi < W - 1 ? xi + 1 : W - 1;
    const float lx0 = xi > 0 ? 0.75f : 1.f, lx1 = 0.25f;
    const T *img = x + (int64_t)n * H * W * ldx + c;
    const int Wo = 2 * W;
    float a[3][3][8];
    auto load_row = [&](int row, float (&dst)[3][8]) {
        const T *r = img + (int64_t)row * W * ldx;
        Vec8<T>::load(r + (int64_t)xm * ldx, dst[0]);
        Vec8<T>::load(r + (int64_t)xi * ldx, dst[1]);
        Vec8<T>::load(r + (int64_t)xp * ldx, dst[2]);
    };
    load_row(y_begin > 0 ? y_begin - 1 : 0, a[0]);
    load_row(y_begin, a[1]);
    for (int yi = y_begin; yi < y_end; ++yi) {
        load_row(yi < H - 1 ? yi + 1 : H - 1, a[2]);
        const float ly0 = yi > 0 ? 0.75f : 1.f, ly1 = 0.25f;
        T *o0 = y + ((int64_t)n * 2 * H + 2 * yi) * Wo * ldy + c;
        T *o1 = o0 + (int64_t)Wo * ldy;
        float o[8];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const float ly = dy == 0 ? ly0 : ly1, hy = 1.f - ly;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const float lx = dx == 0 ? lx0 : lx1, hx = 1.f - lx;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    o[j] = hy * (hx * a[dy][dx][j] + lx * a[dy][dx + 1][j]) + ly * (hx * a[dy + 1][dx][j] + lx * a[dy + 1][dx + 1][j]);
                Vec8<T>::store((dy == 0 ? o0 : o1) + (int64_t)(2 * xi + dx) * ldy, o);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a[0][k][j] = a[1][k][j];
                a[1][k][j] = a[2][k][j];
            }
    }
}

// y = sum_i upsample(x_i): up to 4 low-resolution sources accumulated in FP32 and written once (the PSP priors after
// the bottleneck projection).  Same index rule as bilinear_vec_kernel.
struct BilinearSrcs {
    const void *ptr[4];
    int h[4], w[4], ld[4];
    int n;
};
template <typename T>
__global__ void __launch_bounds__(256) bilinear_sum_kernel(const __grid_constant__ BilinearSrcs src, T *__restrict__ y, int ldy, int N, int Ho,
                                                           int Wo, int C)
{
    const int ncv = C / 8;
    const int row = blockIdx.x;
    const int n = row / Ho, ho = row - n * Ho;
    T *yrow = y + (int64_t)row * Wo * ldy;
    const int items = Wo * ncv;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < items; i += gridDim.y * blockDim.x) {
        const int wo = i / ncv, c = (i - wo * ncv) * 8;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (s >= src.n) break;
            const int H = src.h[s], W = src.w[s], ld = src.ld[s];
            int y0, y1, x0, x1;
            float ly, lx;
            bilinear_src(ho, (float)H / (float)Ho, H, y0, y1, ly);
            bilinear_src(wo, (float)W / (float)Wo, W, x0, x1, lx);
            const T *b = (const T *)src.ptr[s] + (int64_t)n * H * W * ld + c;
            float a00[8], a01[8], a10[8], a11[8];
            Vec8<T>::load(b + ((int64_t)y0 * W + x0) * ld, a00);
            Vec8<T>::load(b + ((int64_t)y0 * W + x1) * ld, a01);
            Vec8<T>::load(b + ((int64_t)y1 * W + x0) * ld, a10);
            Vec8<T>::load(b + ((int64_t)y1 * W + x1) * ld, a11);
            const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += hy * (hx * a00[j] + lx * a01[j]) + ly * (hx * a10[j] + lx * a11[j]);
        }
        Vec8<T>::store(yrow + (int64_t)wo * ldy + c, acc);
    }
}

// Same sum, restructured for the PSP head (tiny sources, 16x..80x magnification): a thread owns one 8-channel vector and walks
// SEG consecutive output pixels of one row.  The source taps of an output pixel change only every Wo/W_s pixels, so the
// vertically blended tap pair of every source lives in registers and is reloaded on a change of x0 -- ~1 load per output
// vector instead of 16; the stores of a warp cover 512 contiguous bytes.  (The blend is re-associated -- vertical first --
// which differs from bilinear_vec_kernel by FP32 rounding only.)
template <typename T, int SEG>
__global__ void __launch_bounds__(128) bilinear_sum_seg_kernel(const __grid_constant__ BilinearSrcs src, T *__restrict__ y, int ldy, int N,
                                                               int Ho, int Wo, int C)
{
    const int cv = blockIdx.z * blockDim.x + threadIdx.x;
    if (cv >= C / 8) return;
    const int c = cv * 8;
    const int row = blockIdx.x;
    const int n = row / Ho, ho = row - n * Ho;
    const int w_begin = blockIdx.y * SEG, w_end = w_begin + SEG < Wo ? w_begin + SEG : Wo;
    T *yrow = y + (int64_t)row * Wo * ldy + c;
    float a0[4][8], a1[4][8];
    int cur[4] = {-1, -1, -1, -1};
    float sw[4];                                           // W_s / Wo once per thread, not once per output pixel (a division each)
#pragma unroll
    for (int s = 0; s < 4; ++s) sw[s] = s < src.n ? (float)src.w[s] / (float)Wo : 0.f;
    for (int wo = w_begin; wo < w_end; ++wo) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (s >= src.n) break;
            const int H = src.h[s], W = src.w[s], ld = src.ld[s];
            int x0, x1;
            float lx;
            bilinear_src(wo, sw[s], W, x0, x1, lx);
            if (x0 != cur[s]) {                                   // warp-uniform
                int y0, y1;
                float ly;
                bilinear_src(ho, (float)H / (float)Ho, H, y0, y1, ly);
                const float hy = 1.f - ly;
                const T *b = (const T *)src.ptr[s] + (int64_t)n * H * W * ld + c;
                float t0[8], t1[8];
                Vec8<T>::load(b + ((int64_t)y0 * W + x0) * ld, t0);
                Vec8<T>::load(b + ((int64_t)y1 * W + x0) * ld, t1);
#pragma unroll
                for (int j = 0; j < 8; ++j) a0[s][j] = hy * t0[j] + ly * t1[j];
                Vec8<T>::load(b + ((int64_t)y0 * W + x1) * ld, t0);
                Vec8<T>::load(b + ((int64_t)y1 * W + x1) * ld, t1);
#pragma unroll
                for (int j = 0; j < 8; ++j) a1[s][j] = hy * t0[j] + ly * t1[j];
                cur[s] = x0;
            }
            const float hx = 1.f - lx;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += hx * a0[s][j] + lx * a1[s][j];
        }
        Vec8<T>::store(yrow + (int64_t)wo * ldy, acc);
    }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) bilinear_scalar_kernel(const TI *__restrict__ x, int ldx, TO *__restrict__ y, int ldy, int N,
                                                              int H, int W, int Ho, int Wo, int C, float sh, float sw)
{
    const int64_t total = (int64_t)N * Ho * Wo * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        int64_t p = i / C;
        int wo = (int)(p % Wo);
        int ho = (int)((p / Wo) % Ho);
        int n = (int)(p / ((int64_t)Wo * Ho));
        int y0, y1, x0, x1;
        float ly, lx;
        bilinear_src(ho, sh, H, y0, y1, ly);
        bilinear_src(wo, sw, W, x0, x1, lx);
        const TI *b = x + (int64_t)n * H * W * ldx + c;
        float a00 = to_f32<TI>(b[((int64_t)y0 * W + x0) * ldx]), a01 = to_f32<TI>(b[((int64_t)y0 * W + x1) * ldx]);
        float a10 = to_f32<TI>(b[((int64_t)y1 * W + x0) * ldx]), a11 = to_f32<TI>(b[((int64_t)y1 * W + x1) * ldx]);
        const float hy = 1.f - ly, hx = 1.f - lx;
        y[p * ldy + c] = from_f32<TO>(hy * (hx * a00 + lx * a01) + ly * (hx * a10 + lx * a11));
    }
}

template <typename TI, typename TO>
static int launch_bilinear(const hn_tensor *x, const hn_tensor *y, cudaStream_t st)
{
    const float sh = (float)x->h / (float)y->h, sw = (float)x->w / (float)y->w;
    if constexpr (std::is_same<TI, TO>::value) {
        if (vec8_ok(x) && vec8_ok(y) && y->h == 2 * x->h && y->w == 2 * x->w) {
            static const bool no_walk = getenv("HN_UP2_NO_WALK") != nullptr;
            constexpr int RY = 8;
            const int64_t items = (int64_t)x->w * (x->c / 8);
            if (!no_walk && x->n <= 65535 && cdiv(x->h, RY) <= 65535) {
                dim3 grid((unsigned)cdiv(items, 128), (unsigned)cdiv(x->h, RY), (unsigned)x->n);
                bilinear_up2_walk_kernel<TI, RY><<<grid, 128, 0, st>>>((const TI *)x->ptr, x->ld, (TO *)y->ptr, y->ld, x->n, x->h, x->w, x->c);
            } else
                bilinear_up2_kernel<TI><<<row_grid((int64_t)x->n * x->h, (int64_t)x->w * (x->c / 8)), 256, 0, st>>>(
                    (const TI *)x->ptr, x->ld, (TO *)y->ptr, y->ld, x->n, x->h, x->w, x->c);
            HN_LAUNCH_CHECK();
            return HN_OK;
        }
    }
    if (vec8_ok(x) && vec8_ok(y)) {
        bilinear_vec_kernel<TI, TO><<<row_grid((int64_t)y->n * y->h, (int64_t)y->w * (y->c / 8)), 256, 0, st>>>((const TI *)x->ptr, x->ld, (TO *)y->ptr, y->ld, x->n, x->h,
                                                                         x->w, y->h, y->w, x->c, sh, sw);
    } else {
        int64_t total = (int64_t)y->n * y->h * y->w * y->c;
        bilinear_scalar_kernel<TI, TO><<<wave_grid(total, 256), 256, 0, st>>>((const TI *)x->ptr, x->ld, (TO *)y->ptr, y->ld, x->n,
                                                                            x->h, x->w, y->h, y->w, x->c, sh, sw);
    }
    HN_LAUNCH_CHECK();
    return HN_OK;
}

}  // namespace hn

using namespace hn;

extern "C" int hn_nchw_to_nhwc(const float *src, const hn_tensor *dst, void *stream)
{
    HN_CHECK_ARG(src && dst && dst->ptr, "hn_nchw_to_nhwc: null pointer");
    HN_CHECK_ARG(dst->ld >= dst->c && dst->c > 0, "hn_nchw_to_nhwc: bad channel stride");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = (int64_t)dst->h * dst->w;
    if (HW * dst->n == 0) return HN_OK;
    if (dst->c <= 8) {
        int grid = wave_grid(HW * dst->n, 256);
        if (dst->dtype == HN_BF16)
            nchw_to_nhwc_smallc<__nv_bfloat16><<<grid, 256, 0, st>>>(src, (__nv_bfloat16 *)dst->ptr, HW, dst->n, dst->c, dst->ld);
        else
            nchw_to_nhwc_smallc<float><<<grid, 256, 0, st>>>(src, (float *)dst->ptr, HW, dst->n, dst->c, dst->ld);
    } else if (dst->c <= 16) {
        int grid = wave_grid(HW * dst->n, 256);
        const size_t esz = elsize(dst->dtype);
        const int vec = ((dst->ld * esz) % (4 * esz) == 0 && (reinterpret_cast<uintptr_t>(dst->ptr) % (4 * esz)) == 0) ? 1 : 0;
        if (dst->dtype == HN_BF16)
            nchw_to_nhwc_c16<__nv_bfloat16><<<grid, 256, 0, st>>>(src, (__nv_bfloat16 *)dst->ptr, HW, dst->n, dst->c, dst->ld, vec);
        else
            nchw_to_nhwc_c16<float><<<grid, 256, 0, st>>>(src, (float *)dst->ptr, HW, dst->n, dst->c, dst->ld, vec);
    } else {
        dim3 grid((unsigned)cdiv(HW, 32), (unsigned)cdiv(dst->c, 32), (unsigned)dst->n);
        if (dst->dtype == HN_BF16)
            transpose_tiles<__nv_bfloat16, true><<<grid, 256, 0, st>>>(src, dst->ptr, HW, dst->c, dst->ld);
        else
            transpose_tiles<float, true><<<grid, 256, 0, st>>>(src, dst->ptr, HW, dst->c, dst->ld);
    }
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_nhwc_to_nchw(const hn_tensor *src, float *dst, void *stream)
{
    HN_CHECK_ARG(src && src->ptr && dst, "hn_nhwc_to_nchw: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = (int64_t)src->h * src->w;
    if (HW * src->n == 0) return HN_OK;
    if (src->c <= 16 && src->ld >= ((src->c + 3) / 4) * 4) {      // the float4 path may read up to 3 pad channels
        int grid = wave_grid(HW * src->n, 256);
        if (src->dtype == HN_BF16)
            nhwc_to_nchw_smallc<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)src->ptr, dst, HW, src->n, src->c, src->ld);
        else
            nhwc_to_nchw_smallc<float><<<grid, 256, 0, st>>>((const float *)src->ptr, dst, HW, src->n, src->c, src->ld);
        HN_LAUNCH_CHECK();
        return HN_OK;
    }
    dim3 grid((unsigned)cdiv(HW, 32), (unsigned)cdiv(src->c, 32), (unsigned)src->n);
    if (src->dtype == HN_BF16)
        transpose_tiles<__nv_bfloat16, false><<<grid, 256, 0, st>>>(src->ptr, dst, HW, src->c, src->ld);
    else
        transpose_tiles<float, false><<<grid, 256, 0, st>>>(src->ptr, dst, HW, src->c, src->ld);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_pack_weight_scaled(const float *w_oihw, const float *row_scale, void *dst, int32_t dtype, int32_t cout, int32_t cin,
                                     int32_t r, int32_t s, int32_t cout_pad, int32_t kpad, void *stream);
extern "C" int hn_pack_weight(const float *w_oihw, void *dst, int32_t dtype, int32_t cout, int32_t cin, int32_t r, int32_t s,
                              int32_t cout_pad, int32_t kpad, void *stream)
{
    return hn_pack_weight_scaled(w_oihw, nullptr, dst, dtype, cout, cin, r, s, cout_pad, kpad, stream);
}

extern "C" int hn_pack_weight_scaled(const float *w_oihw, const float *row_scale, void *dst, int32_t dtype, int32_t cout, int32_t cin,
                                     int32_t r, int32_t s, int32_t cout_pad, int32_t kpad, void *stream)
{
    HN_CHECK_ARG(w_oihw && dst, "hn_pack_weight: null pointer");
    HN_CHECK_ARG(cout_pad >= cout && kpad >= r * s * cin, "hn_pack_weight: padded sizes too small");
    int64_t total = (int64_t)cout_pad * kpad;
    int grid = wave_grid(total, 256);
    if (dtype == HN_BF16)
        pack_weight_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(w_oihw, row_scale, (__nv_bfloat16 *)dst, cout, cin, r, s, cout_pad, kpad);
    else
        pack_weight_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(w_oihw, row_scale, (float *)dst, cout, cin, r, s, cout_pad, kpad);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_bn_fold(const float *gamma, const float *beta, const float *mean, const float *var, const float *bias, float eps,
                          float *scale, float *shift, int32_t c, void *stream)
{
    HN_CHECK_ARG(scale && shift && c > 0, "hn_bn_fold: bad arguments");
    HN_CHECK_ARG(!gamma || (beta && mean && var), "hn_bn_fold: gamma given without beta/mean/var");
    bn_fold_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, bias, eps, scale, shift, c);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_channel_stats(const hn_tensor *x, double *sum, double *sqsum, void *stream)
{
    HN_CHECK_ARG(x && x->ptr && sum && sqsum, "hn_channel_stats: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    HN_CUDA(cudaMemsetAsync(sum, 0, sizeof(double) * x->c, st));
    HN_CUDA(cudaMemsetAsync(sqsum, 0, sizeof(double) * x->c, st));
    const int64_t npix = (int64_t)x->n * x->h * x->w;
    if (npix == 0) return HN_OK;
    if (!vec8_ok(x)) {
        int64_t nch = cdiv((int64_t)num_sms() * 4, x->c);
        int64_t ppc = cdiv(npix, nch);
        if (ppc < 2048) ppc = 2048;
        nch = cdiv(npix, ppc);
        dim3 g((unsigned)nch, (unsigned)x->c);
        if (x->dtype == HN_BF16) channel_stats_scalar_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16 *)x->ptr, npix, x->ld, ppc, sum, sqsum);
        else channel_stats_scalar_kernel<float><<<g, 256, 0, st>>>((const float *)x->ptr, npix, x->ld, ppc, sum, sqsum);
        HN_LAUNCH_CHECK();
        return HN_OK;
    }
    const int ncv = x->c / 8;
    const int CVB = ncv < 32 ? ncv : 32;
    const int PL = 256 / CVB;
    const int cvblocks = (int)cdiv(ncv, CVB);
    // ~4 waves of CTAs, at least PL*8 pixels each
    int64_t chunks = cdiv((int64_t)num_sms() * 4, cvblocks);
    int64_t pix_per_cta = cdiv(npix, chunks);
    if (pix_per_cta < (int64_t)PL * 8) pix_per_cta = (int64_t)PL * 8;
    chunks = cdiv(npix, pix_per_cta);
    dim3 grid((unsigned)chunks, (unsigned)cvblocks), block(CVB, PL);
    size_t smem = (size_t)2 * PL * CVB * 8 * sizeof(float);
    if (x->dtype == HN_BF16)
        channel_stats_kernel<__nv_bfloat16><<<grid, block, smem, st>>>((const __nv_bfloat16 *)x->ptr, npix, x->c, x->ld, pix_per_cta, sum, sqsum, BnFinalize{});
    else
        channel_stats_kernel<float><<<grid, block, smem, st>>>((const float *)x->ptr, npix, x->c, x->ld, pix_per_cta, sum, sqsum, BnFinalize{});
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_bn_finalize_tracked(const double *sum, const double *sqsum, int64_t count, const float *gamma, const float *beta, float eps,
                                      float momentum, float *running_mean, float *running_var, int64_t *num_batches_tracked, float *scale,
                                      float *shift, float *save_mean, float *save_invstd, int32_t c, void *stream)
{
    HN_CHECK_ARG(sum && sqsum && scale && shift && c > 0 && count > 0, "hn_bn_finalize_tracked: bad arguments");
    bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum, sqsum, (double)count, gamma, beta, eps, momentum,
                                                                         running_mean, running_var, scale, shift, save_mean,
                                                                         save_invstd, c, reinterpret_cast<long long *>(num_batches_tracked));
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int64_t hn_bn_batch_stats_scratch_bytes(int32_t c) { return (int64_t)sizeof(double) * (2 * (int64_t)c + 1); }

extern "C" int hn_bn_batch_stats(const hn_tensor *x, void *scratch, const float *gamma, const float *beta, float eps, float momentum,
                                 float *running_mean, float *running_var, int64_t *num_batches_tracked, float *scale, float *shift,
                                 float *save_mean, float *save_invstd, void *stream)
{
    HN_CHECK_ARG(x && x->ptr && scratch && scale && shift, "hn_bn_batch_stats: null pointer");
    HN_CHECK_ARG(x->c >= 1 && x->ld >= x->c, "hn_bn_batch_stats: bad channel strides");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t npix = (int64_t)x->n * x->h * x->w;
    HN_CHECK_ARG(npix > 0, "hn_bn_batch_stats: empty input");
    double *sum = reinterpret_cast<double *>(scratch), *sqsum = sum + x->c;
    if (!vec8_ok(x)) {      // narrow / unaligned views: the three-launch sequence
        int rc = hn_channel_stats(x, sum, sqsum, stream);
        if (rc) return rc;
        rc = hn_bn_finalize(sum, sqsum, npix, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, save_mean, save_invstd, x->c,
                            stream);
        if (rc) return rc;
        if (num_batches_tracked) {
            // rare path (C not a multiple of 8): a one-thread finalize of the counter
            bn_count_kernel<<<1, 1, 0, st>>>(reinterpret_cast<long long *>(num_batches_tracked));
            HN_LAUNCH_CHECK();
        }
        return HN_OK;
    }
    HN_CUDA(cudaMemsetAsync(scratch, 0, (size_t)hn_bn_batch_stats_scratch_bytes(x->c), st));
    const int ncv = x->c / 8;
    const int CVB = ncv < 32 ? ncv : 32;
    const int PL = 256 / CVB;
    const int cvblocks = (int)cdiv(ncv, CVB);
    int64_t chunks = cdiv((int64_t)num_sms() * 4, cvblocks);
    int64_t pix_per_cta = cdiv(npix, chunks);
    if (pix_per_cta < (int64_t)PL * 8) pix_per_cta = (int64_t)PL * 8;
    chunks = cdiv(npix, pix_per_cta);
    dim3 grid((unsigned)chunks, (unsigned)cvblocks), block(CVB, PL);
    size_t smem = (size_t)2 * PL * CVB * 8 * sizeof(float);
    BnFinalize f{};
    f.ticket = reinterpret_cast<unsigned int *>(sum + 2 * x->c);
    f.count = (double)npix;
    f.gamma = gamma;
    f.beta = beta;
    f.eps = eps;
    f.momentum = momentum;
    f.running_mean = running_mean;
    f.running_var = running_var;
    f.num_batches_tracked = reinterpret_cast<long long *>(num_batches_tracked);
    f.scale = scale;
    f.shift = shift;
    f.save_mean = save_mean;
    f.save_invstd = save_invstd;
    if (x->dtype == HN_BF16)
        channel_stats_kernel<__nv_bfloat16><<<grid, block, smem, st>>>((const __nv_bfloat16 *)x->ptr, npix, x->c, x->ld, pix_per_cta, sum, sqsum, f);
    else
        channel_stats_kernel<float><<<grid, block, smem, st>>>((const float *)x->ptr, npix, x->c, x->ld, pix_per_cta, sum, sqsum, f);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_bn_finalize(const double *sum, const double *sqsum, int64_t count, const float *gamma, const float *beta, float eps,
                              float momentum, float *running_mean, float *running_var, float *scale, float *shift, float *save_mean,
                              float *save_invstd, int32_t c, void *stream)
{
    HN_CHECK_ARG(sum && sqsum && scale && shift && c > 0 && count > 0, "hn_bn_finalize: bad arguments");
    bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum, sqsum, (double)count, gamma, beta, eps, momentum,
                                                                         running_mean, running_var, scale, shift, save_mean,
                                                                         save_invstd, c);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_affine_act(const hn_tensor *x, const hn_epilogue *ep, const hn_tensor *y, void *stream)
{
    HN_CHECK_ARG(x && y && ep && x->ptr && y->ptr, "hn_affine_act: null pointer");
    HN_CHECK_ARG(x->dtype == y->dtype || (x->dtype == HN_F32 && y->dtype == HN_BF16),
                 "hn_affine_act: x/y dtypes must match, or FP32 -> BF16");
    HN_CHECK_ARG(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c, "hn_affine_act: shape mismatch");
    HN_CHECK_ARG(vec8_ok(x) && vec8_ok(y), "hn_affine_act: views must be 8-channel aligned");
    HN_CHECK_ARG(ep->scale != nullptr || ep->shift == nullptr, "hn_affine_act: shift given without scale");
    const int64_t npix = (int64_t)x->n * x->h * x->w;
    const int64_t ppi = ep->per_image ? (int64_t)x->h * x->w : 0;
    if (npix == 0) return HN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    using bf16 = __nv_bfloat16;
    if (ep->scale && ep->shift && !ep->per_image && npix >= 4096) {      // the BatchNorm2d normalise pass: thread-fixed channels
        const int ncv = x->c / 8;
        const int CVB = ncv < 32 ? ncv : 32;
        const int PL = 256 / CVB;
        const int cvblocks = (int)cdiv(ncv, CVB);
        int64_t chunks = cdiv((int64_t)num_sms() * 8, cvblocks);
        int64_t pix_per_cta = cdiv(npix, chunks);
        if (pix_per_cta < (int64_t)PL * 4) pix_per_cta = (int64_t)PL * 4;
        chunks = cdiv(npix, pix_per_cta);
        dim3 g((unsigned)chunks, (unsigned)cvblocks), b(CVB, PL);
        if (y->dtype == HN_BF16 && x->dtype == HN_BF16)
            affine_act_fixed_kernel<bf16, bf16><<<g, b, 0, st>>>((const bf16 *)x->ptr, x->ld, ep->scale, ep->shift, (const bf16 *)ep->residual,
                                                                 ep->residual_ld, ep->act, ep->slope, ep->slope_ptr, (bf16 *)y->ptr, y->ld, npix, x->c, pix_per_cta);
        else if (y->dtype == HN_BF16)
            affine_act_fixed_kernel<float, bf16><<<g, b, 0, st>>>((const float *)x->ptr, x->ld, ep->scale, ep->shift, (const bf16 *)ep->residual,
                                                                  ep->residual_ld, ep->act, ep->slope, ep->slope_ptr, (bf16 *)y->ptr, y->ld, npix, x->c, pix_per_cta);
        else
            affine_act_fixed_kernel<float, float><<<g, b, 0, st>>>((const float *)x->ptr, x->ld, ep->scale, ep->shift, (const float *)ep->residual,
                                                                   ep->residual_ld, ep->act, ep->slope, ep->slope_ptr, (float *)y->ptr, y->ld, npix, x->c, pix_per_cta);
        HN_LAUNCH_CHECK();
        return HN_OK;
    }
    int grid = wave_grid(cdiv(npix, 64) * 256, 256, 16);
    if (y->dtype == HN_BF16 && x->dtype == HN_BF16)
        affine_act_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16 *)x->ptr, x->ld, ep->scale, ep->shift, (const bf16 *)ep->residual,
                                                          ep->residual_ld, ep->act, ep->slope, ep->slope_ptr, (bf16 *)y->ptr, y->ld, npix,
                                                          x->c, ppi);
    else if (y->dtype == HN_BF16)   // FP32 pre-normalisation values -> BF16 activations (train-mode BatchNorm2d)
        affine_act_kernel<float, bf16><<<grid, 256, 0, st>>>((const float *)x->ptr, x->ld, ep->scale, ep->shift, (const bf16 *)ep->residual,
                                                           ep->residual_ld, ep->act, ep->slope, ep->slope_ptr, (bf16 *)y->ptr, y->ld, npix,
                                                           x->c, ppi);
    else
        affine_act_kernel<float, float><<<grid, 256, 0, st>>>((const float *)x->ptr, x->ld, ep->scale, ep->shift, (const float *)ep->residual,
                                                            ep->residual_ld, ep->act, ep->slope, ep->slope_ptr, (float *)y->ptr, y->ld, npix,
                                                            x->c, ppi);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_bn_apply_train(const hn_tensor *x, const double *sum, const double *sqsum, int64_t count, const float *gamma,
                                 const float *beta, float eps, float momentum, float *running_mean, float *running_var,
                                 int64_t *num_batches_tracked, const hn_epilogue *ep, const hn_tensor *y, float *scale, float *shift,
                                 float *save_mean, float *save_invstd, void *stream)
{
    const int groups = ep && ep->stat_groups > 1 ? ep->stat_groups : 1;
    HN_CHECK_ARG(x && y && ep && sum && sqsum && scale && shift && save_mean && save_invstd && x->ptr && y->ptr, "hn_bn_apply_train: null pointer");
    HN_CHECK_ARG(x->dtype == y->dtype || (x->dtype == HN_F32 && y->dtype == HN_BF16), "hn_bn_apply_train: x/y dtypes must match, or FP32 -> BF16");
    HN_CHECK_ARG(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c, "hn_bn_apply_train: shape mismatch");
    HN_CHECK_ARG(vec8_ok(x) && vec8_ok(y), "hn_bn_apply_train: views must be 8-channel aligned");
    HN_CHECK_ARG(count > 0 && (running_mean != nullptr) == (running_var != nullptr), "hn_bn_apply_train: bad count / running statistics");
    HN_CHECK_ARG(x->n % groups == 0, "hn_bn_apply_train: %d images do not split into %d statistics groups", x->n, groups);
    const int64_t npix = (int64_t)(x->n / groups) * x->h * x->w;          // per group; `count` is the per-group pixel count too
    if (npix == 0) return HN_OK;
    HN_CHECK_ARG(count == npix, "hn_bn_apply_train: count must be the number of pixels per statistics group");
    cudaStream_t st = (cudaStream_t)stream;
    using bf16 = __nv_bfloat16;
    BnTrainFin fin{sum, sqsum, groups, npix, (double)count, gamma, beta, eps, momentum, running_mean, running_var, (long long *)num_batches_tracked,
                   scale, shift, save_mean, save_invstd};
    const int ncv = x->c / 4;
    const int CVB = ncv < 32 ? ncv : 32;
    const int PL = 256 / CVB;
    const int cvblocks = (int)cdiv(ncv, CVB);
    int64_t chunks = cdiv((int64_t)num_sms() * 8, cvblocks);       // two waves of 4 CTAs per SM
    int64_t pix_per_cta = cdiv(npix, chunks);
    if (pix_per_cta < (int64_t)PL * 8) pix_per_cta = (int64_t)PL * 8;
    chunks = cdiv(npix, pix_per_cta);
    dim3 g((unsigned)chunks, (unsigned)cvblocks, (unsigned)groups), b(CVB, PL);
    if (y->dtype == HN_BF16 && x->dtype == HN_BF16)
        bn_apply_train_kernel<bf16, bf16><<<g, b, 0, st>>>((const bf16 *)x->ptr, x->ld, fin, (const bf16 *)ep->residual, ep->residual_ld, ep->act,
                                                           ep->slope, ep->slope_ptr, (bf16 *)y->ptr, y->ld, npix, x->c, pix_per_cta);
    else if (y->dtype == HN_BF16)
        bn_apply_train_kernel<float, bf16><<<g, b, 0, st>>>((const float *)x->ptr, x->ld, fin, (const bf16 *)ep->residual, ep->residual_ld, ep->act,
                                                            ep->slope, ep->slope_ptr, (bf16 *)y->ptr, y->ld, npix, x->c, pix_per_cta);
    else
        bn_apply_train_kernel<float, float><<<g, b, 0, st>>>((const float *)x->ptr, x->ld, fin, (const float *)ep->residual, ep->residual_ld, ep->act,
                                                             ep->slope, ep->slope_ptr, (float *)y->ptr, y->ld, npix, x->c, pix_per_cta);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_maxpool3x3s2_fwd(const hn_tensor *x, const hn_tensor *y, void *stream)
{
    HN_CHECK_ARG(x && y && x->ptr && y->ptr, "hn_maxpool3x3s2_fwd: null pointer");
    HN_CHECK_ARG(x->dtype == y->dtype && x->c == y->c && x->n == y->n, "hn_maxpool3x3s2_fwd: dtype/shape mismatch");
    HN_CHECK_ARG(y->h == (x->h - 1) / 2 + 1 && y->w == (x->w - 1) / 2 + 1, "hn_maxpool3x3s2_fwd: output must be %dx%d",
                 (x->h - 1) / 2 + 1, (x->w - 1) / 2 + 1);
    HN_CHECK_ARG(vec8_ok(x) && vec8_ok(y), "hn_maxpool3x3s2_fwd: views must be 8-channel aligned");
    int64_t total = (int64_t)y->n * y->h * y->w * (y->c / 8);
    if (total == 0) return HN_OK;
    dim3 grid = row_grid((int64_t)y->n * y->h, (int64_t)y->w * (y->c / 8));
    cudaStream_t st = (cudaStream_t)stream;
    if (x->dtype == HN_BF16)
        maxpool_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x->ptr, x->ld, (__nv_bfloat16 *)y->ptr, y->ld, x->n, x->h,
                                                          x->w, y->h, y->w, x->c);
    else
        maxpool_kernel<float><<<grid, 256, 0, st>>>((const float *)x->ptr, x->ld, (float *)y->ptr, y->ld, x->n, x->h, x->w, y->h, y->w, x->c);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int64_t hn_pyramid_pool_workspace_bytes(const hn_tensor *x, const int32_t *sizes, int32_t nsizes)
{
    int ncols = 0;
    for (int i = 0; i < nsizes; ++i) ncols += sizes[i];
    return (int64_t)x->n * x->h * ncols * x->c * (int64_t)sizeof(float);
}

extern "C" int hn_pyramid_pool_fwd(const hn_tensor *x, const int32_t *sizes, int32_t nsizes, void *out, void *workspace,
                                      int64_t workspace_bytes, void *stream)
{
    HN_CHECK_ARG(x && x->ptr && sizes && out && workspace, "hn_pyramid_pool_fwd: null pointer");
    HN_CHECK_ARG(nsizes >= 1 && nsizes <= 4, "hn_pyramid_pool_fwd: 1..4 pyramid sizes supported");
    HN_CHECK_ARG(vec8_ok(x), "hn_pyramid_pool_fwd: view must be 8-channel aligned");
    PoolSizes ps{};
    ps.n = nsizes;
    int ncols = 0;
    for (int i = 0; i < nsizes; ++i) {
        HN_CHECK_ARG(sizes[i] >= 1, "hn_pyramid_pool_fwd: bad size");
        ps.s[i] = sizes[i];
        ncols += sizes[i];
    }
    HN_CHECK_ARG(ncols <= kMaxPoolCols, "hn_pyramid_pool_fwd: sum of sizes %d > %d", ncols, kMaxPoolCols);
    ps.ncols = 0;
    for (int i = 0; i < nsizes; ++i)
        for (int j = 0; j < sizes[i]; ++j) {
            ps.w0[ps.ncols] = (j * x->w) / sizes[i];
            ps.w1[ps.ncols] = ((j + 1) * x->w + sizes[i] - 1) / sizes[i];
            ++ps.ncols;
        }
    {   // segment bounds = sorted union of all bin bounds
        int nb = 0, b[2 * kMaxPoolCols + 2];
        for (int i = 0; i < ps.ncols; ++i) { b[nb++] = ps.w0[i]; b[nb++] = ps.w1[i]; }
        for (int i = 1; i < nb; ++i)
            for (int j = i; j > 0 && b[j - 1] > b[j]; --j) { int t = b[j]; b[j] = b[j - 1]; b[j - 1] = t; }
        int m = 0;
        for (int i = 0; i < nb; ++i)
            if (m == 0 || ps.segb[m - 1] != b[i]) ps.segb[m++] = b[i];
        ps.nseg = m - 1;
    }
    if (workspace_bytes < hn_pyramid_pool_workspace_bytes(x, sizes, nsizes)) {
        set_error("hn_pyramid_pool_fwd: workspace too small");
        return HN_ERR_WORKSPACE;
    }
    if ((int64_t)x->n * x->h * x->w == 0) return HN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int ncv = x->c / 8;
    const int threads = ncv < 256 ? ((ncv + 31) / 32) * 32 : 256;
    dim3 grid1((unsigned)(x->n * x->h), (unsigned)cdiv(ncv, threads));
    int nbins = 0;
    for (int i = 0; i < nsizes; ++i) nbins += sizes[i] * sizes[i];
    int grid2 = wave_grid((int64_t)x->n * nbins * ncv, 256);
    if (x->dtype == HN_BF16) {
        pyramid_rows_kernel<__nv_bfloat16><<<grid1, threads, 0, st>>>((const __nv_bfloat16 *)x->ptr, x->ld, x->h, x->w, x->c, ps, (float *)workspace);
        pyramid_bins_kernel<__nv_bfloat16><<<grid2, 256, 0, st>>>((const float *)workspace, x->n, x->h, x->w, x->c, ps, (__nv_bfloat16 *)out);
    } else {
        pyramid_rows_kernel<float><<<grid1, threads, 0, st>>>((const float *)x->ptr, x->ld, x->h, x->w, x->c, ps, (float *)workspace);
        pyramid_bins_kernel<float><<<grid2, 256, 0, st>>>((const float *)workspace, x->n, x->h, x->w, x->c, ps, (float *)out);
    }
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_bilinear_fwd(const hn_tensor *x, const hn_tensor *y, void *stream)
{
    HN_CHECK_ARG(x && y && x->ptr && y->ptr, "hn_bilinear_fwd: null pointer");
    HN_CHECK_ARG(x->n == y->n && x->c == y->c, "hn_bilinear_fwd: batch/channel mismatch");
    HN_CHECK_ARG(x->h > 0 && x->w > 0, "hn_bilinear_fwd: empty input");
    if ((int64_t)y->n * y->h * y->w == 0) return HN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (x->dtype == HN_BF16 && y->dtype == HN_BF16) return launch_bilinear<__nv_bfloat16, __nv_bfloat16>(x, y, st);
    if (x->dtype == HN_F32 && y->dtype == HN_F32) return launch_bilinear<float, float>(x, y, st);
    if (x->dtype == HN_BF16 && y->dtype == HN_F32) return launch_bilinear<__nv_bfloat16, float>(x, y, st);
    return launch_bilinear<float, __nv_bfloat16>(x, y, st);
}

extern "C" int hn_bilinear_sum_fwd(const hn_tensor *xs, int32_t nsrc, const hn_tensor *y, void *stream)
{
    HN_CHECK_ARG(xs && y && y->ptr && nsrc >= 1 && nsrc <= 4, "hn_bilinear_sum_fwd: 1..4 sources");
    HN_CHECK_ARG(vec8_ok(y), "hn_bilinear_sum_fwd: output view must be 8-channel aligned");
    BilinearSrcs src{};
    src.n = nsrc;
    for (int i = 0; i < nsrc; ++i) {
        HN_CHECK_ARG(xs[i].ptr && xs[i].dtype == y->dtype && xs[i].n == y->n && xs[i].c == y->c && vec8_ok(&xs[i]) && xs[i].h > 0 && xs[i].w > 0,
                     "hn_bilinear_sum_fwd: source %d does not match the output", i);
        src.ptr[i] = xs[i].ptr; src.h[i] = xs[i].h; src.w[i] = xs[i].w; src.ld[i] = xs[i].ld;
    }
    if ((int64_t)y->n * y->h * y->w == 0) return HN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    bool magnify = true;          // every source at least 4x smaller along W: the register-resident tap pairs pay off
    for (int i = 0; i < nsrc; ++i) magnify = magnify && (4 * xs[i].w <= y->w);
    if (magnify && (int64_t)y->n * y->h < 65535) {
        constexpr int SEG = 16;
        dim3 g((unsigned)(y->n * y->h), (unsigned)cdiv(y->w, SEG), (unsigned)cdiv(y->c / 8, 128));
        if (y->dtype == HN_BF16) bilinear_sum_seg_kernel<__nv_bfloat16, SEG><<<g, 128, 0, st>>>(src, (__nv_bfloat16 *)y->ptr, y->ld, y->n, y->h, y->w, y->c);
        else bilinear_sum_seg_kernel<float, SEG><<<g, 128, 0, st>>>(src, (float *)y->ptr, y->ld, y->n, y->h, y->w, y->c);
        HN_LAUNCH_CHECK();
        return HN_OK;
    }
    dim3 grid = row_grid((int64_t)y->n * y->h, (int64_t)y->w * (y->c / 8));
    if (y->dtype == HN_BF16) bilinear_sum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(src, (__nv_bfloat16 *)y->ptr, y->ld, y->n, y->h, y->w, y->c);
    else bilinear_sum_kernel<float><<<grid, 256, 0, st>>>(src, (float *)y->ptr, y->ld, y->n, y->h, y->w, y->c);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------ stem preparation
namespace hn {
// NHWC (C <= 4, any channel stride) -> zero-bordered 4-channel image [N][Hp][Wp][4] BF16 (border 3 = conv padding)
template <typename T>
__global__ void __launch_bounds__(256) stem_pad_kernel(const T *__restrict__ x, int ldx, int N, int H, int W, int C, __nv_bfloat16 *__restrict__ dst,
                                                       int Hp, int Wp)
{
    const int row = blockIdx.x;          // n*Hp + hp
    const int n = row / Hp, hp = row - n * Hp;
    const int hi = hp - 3;
    for (int wp = blockIdx.y * blockDim.x + threadIdx.x; wp < Wp; wp += gridDim.y * blockDim.x) {
        const int wi = wp - 3;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
            const T *s = x + (((int64_t)n * H + hi) * W + wi) * ldx;
            for (int c = 0; c < C; ++c) v[c] = to_f32<T>(s[c]);
        }
        uint2 o;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&o);
        h[0] = __floats2bfloat162_rn(v[0], v[1]);
        h[1] = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2 *>(dst + ((int64_t)row * Wp + wp) * 4) = o;
    }
}
// OIHW FP32 [cout][cin<=4][7][7] -> [cout_pad=64][7][8][4] BF16 (zero for s = 7 and c >= cin), optional per-row scale
__global__ void pack_stem_weight_kernel(const float *__restrict__ w, const float *__restrict__ row_scale, __nv_bfloat16 *__restrict__ dst, int cout,
                                        int cin)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 64 * 224) return;
    const int o = i / 224, k = i - o * 224;
    const int r = k / 32, s = (k % 32) / 4, c = k % 4;
    float v = 0.f;
    if (o < cout && s < 7 && c < cin) {
        v = w[(((int64_t)o * cin + c) * 7 + r) * 7 + s];
        if (row_scale) v *= row_scale[o];
    }
    dst[i] = __float2bfloat16_rn(v);
}
int conv_stem_tc(const hn_tensor *xpad, const void *w, int cout, const hn_epilogue *ep, const hn_tensor *y, cudaStream_t st);
int conv_stem_dense(const hn_tensor *xpad, const void *w, int cout, const hn_epilogue *ep, const hn_tensor *y, cudaStream_t st);
}  // namespace hn

extern "C" int hn_stem_pad(const hn_tensor *x, const hn_tensor *xpad, void *stream)
{
    HN_CHECK_ARG(x && xpad && x->ptr && xpad->ptr, "hn_stem_pad: null pointer");
    HN_CHECK_ARG(x->c >= 1 && x->c <= 4 && xpad->c == 4 && xpad->ld == 4 && xpad->dtype == HN_BF16 && xpad->n == x->n, "hn_stem_pad: bad views");
    HN_CHECK_ARG(xpad->h >= x->h + 6 && xpad->w >= x->w + 6, "hn_stem_pad: padded image must be at least (H+6)x(W+6)");
    dim3 grid((unsigned)(xpad->n * xpad->h), (unsigned)hn::cdiv(xpad->w, 1024));
    cudaStream_t st = (cudaStream_t)stream;
    if (x->dtype == HN_BF16)
        hn::stem_pad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x->ptr, x->ld, x->n, x->h, x->w, x->c, (__nv_bfloat16 *)xpad->ptr, xpad->h, xpad->w);
    else
        hn::stem_pad_kernel<float><<<grid, 256, 0, st>>>((const float *)x->ptr, x->ld, x->n, x->h, x->w, x->c, (__nv_bfloat16 *)xpad->ptr, xpad->h, xpad->w);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_pack_stem_weight(const float *w_oihw, const float *row_scale, void *dst, int32_t cout, int32_t cin, void *stream)
{
    HN_CHECK_ARG(w_oihw && dst && cout >= 1 && cout <= 64 && cin >= 1 && cin <= 4, "hn_pack_stem_weight: 7x7 filter with Cout <= 64, Cin <= 4");
    hn::pack_stem_weight_kernel<<<(64 * 224 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w_oihw, row_scale, (__nv_bfloat16 *)dst, cout, cin);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_stem7x7s2_fwd(const hn_tensor *xpad, const void *w_packed, int32_t cout, const hn_epilogue *ep, const hn_tensor *y, void *stream)
{
    HN_CHECK_ARG(xpad && w_packed && ep && y && xpad->ptr && y->ptr, "hn_stem7x7s2_fwd: null pointer");
    HN_CHECK_ARG(!ep->stat_sum || (ep->stat_sqsum && !ep->residual && ep->act == HN_ACT_NONE && (y->dtype == HN_F32 || !ep->scale) &&
                                   (reinterpret_cast<uintptr_t>(y->ptr) & 15) == 0 && (y->ld * hn::elsize(y->dtype)) % 16 == 0),
                 "hn_stem7x7s2_fwd: fused statistics need a 16-byte aligned output view, no activation (no explicit scale with a BF16 output)");
    HN_CHECK_ARG(y->n == xpad->n && y->c == cout, "hn_stem7x7s2_fwd: output view mismatch");
    // default: the dense-row kernel (hn_conv_stem.cu; needs hn_stem_pad_slack_bytes() readable bytes behind the image);
    // HN_STEM_WINDOW_TMA=1 selects the overlapping-window tensor-map variant (hn_conv_tc.cu)
    static const bool window_tma = getenv("HN_STEM_WINDOW_TMA") != nullptr;
    if (window_tma || xpad->w % 2 != 0) return hn::conv_stem_tc(xpad, w_packed, cout, ep, y, (cudaStream_t)stream);
    return hn::conv_stem_dense(xpad, w_packed, cout, ep, y, (cudaStream_t)stream);
}
