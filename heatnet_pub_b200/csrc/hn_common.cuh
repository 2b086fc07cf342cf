// Shared helpers for libheatnet_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "heatnet_b200.h"

namespace hn {

void set_error(const char *fmt, ...);

#define HN_CHECK_ARG(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            hn::set_error(__VA_ARGS__);    \
            return HN_ERR_ARG;             \
        }                                  \
    } while (0)

#define HN_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            hn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return HN_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define HN_LAUNCH_CHECK() HN_CUDA(cudaGetLastError())

// Per-device host state (cudaFuncSetAttribute is per device, and so are SM counts): one process may drive several GPUs
// (torch.nn.DataParallel-style callers), so nothing device-dependent is cached in a plain static.
constexpr int HN_MAX_DEVICES = 64;
inline int current_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= HN_MAX_DEVICES) dev = 0;
    return dev;
}

inline int num_sms()
{
    static int n[HN_MAX_DEVICES] = {};
    const int dev = current_device();
    if (n[dev] == 0) {
        if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n[dev] = 148;
    }
    return n[dev];
}

inline size_t elsize(int dtype) { return dtype == HN_BF16 ? 2 : 4; }
inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device-side dtype helpers ----
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float apply_act(float v, int act, float slope)
{
    if (act == HN_ACT_RELU) return fmaxf(v, 0.f);
    if (act == HN_ACT_LEAKY) return v >= 0.f ? v : v * slope;
    return v;
}

// 8 channels (16 B of BF16 / 32 B of FP32) <-> 8 floats
template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[8])
    {
        uint4 raw = *reinterpret_cast<const uint4 *>(p);
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[8])
    {
        uint4 raw;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4 *>(p) = raw;
    }
};
template <> struct Vec8<float> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[8])
    {
        float4 a = *reinterpret_cast<const float4 *>(p);
        float4 b = *reinterpret_cast<const float4 *>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[8])
    {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};

// 4 channels (8 B of BF16 / 16 B of FP32) <-> 4 floats: the streaming BatchNorm kernels keep per-channel coefficients in registers,
// and 4 channels per thread (instead of 8) halves that footprint, which buys the occupancy a bandwidth-bound kernel needs
template <typename T> struct Vec4;
template <> struct Vec4<__nv_bfloat16> {
    typedef uint2 raw_t;          // packed form: batches of loads are held packed and widened one pixel at a time
    static __device__ __forceinline__ raw_t ldraw(const __nv_bfloat16 *p) { return *reinterpret_cast<const uint2 *>(p); }
    static __device__ __forceinline__ void unpack(const raw_t &raw, float (&v)[4])
    {
        v[0] = __uint_as_float(raw.x << 16); v[1] = __uint_as_float(raw.x & 0xffff0000u);
        v[2] = __uint_as_float(raw.y << 16); v[3] = __uint_as_float(raw.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[4]) { unpack(ldraw(p), v); }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[4])
    {
        uint2 raw;
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        raw.x = *reinterpret_cast<uint32_t *>(&a);
        raw.y = *reinterpret_cast<uint32_t *>(&b);
        *reinterpret_cast<uint2 *>(p) = raw;
    }
};
template <> struct Vec4<float> {
    typedef float4 raw_t;
    static __device__ __forceinline__ raw_t ldraw(const float *p) { return *reinterpret_cast<const float4 *>(p); }
    static __device__ __forceinline__ void unpack(const raw_t &a, float (&v)[4]) { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
    static __device__ __forceinline__ void load(const float *p, float (&v)[4]) { unpack(ldraw(p), v); }
    static __device__ __forceinline__ void store(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};

// true when an NHWC view can be accessed in 8-channel vectors
inline bool vec8_ok(const hn_tensor *t)
{
    size_t es = elsize(t->dtype);
    return t->c % 8 == 0 && t->ld % 8 == 0 && (reinterpret_cast<uintptr_t>(t->ptr) % (8 * es)) == 0;
}

// conv launchers implemented in the per-engine translation units
int conv2d_fwd_f32(const hn_tensor *x, const void *w, const hn_conv *cv, const hn_epilogue *ep, const hn_tensor *y,
                   cudaStream_t st);
int conv2d_fwd_tc(const hn_tensor *x, const void *w, const hn_conv *cv, const hn_epilogue *ep, const hn_tensor *y,
                  void *ws, int64_t ws_bytes, cudaStream_t st);
int64_t conv2d_tc_workspace(const hn_tensor *x, const hn_conv *cv);
bool conv_implicit_ok(const hn_tensor *x, const hn_conv *cv);     // BF16 implicit GEMM without an im2col workspace (stride 1 or 2)
struct TcSubConv {              // one parity phase of a stride-2 dgrad (hn_conv2d_dgrad_s2): stride-1 correlation onto a sub-lattice
    int R, S, pad_h, pad_w;     // taps and per-axis padding of the phase's sub-filter
    int out_h, out_w;           // extent of the phase's output lattice
    void *y;                    // BF16 element (n=0, m=0, n=0, c=0) of the lattice
    int64_t pix_stride, row_stride, img_stride;   // lattice strides in elements
    int accumulate;             // add to the lattice's current contents
};
int conv2d_fwd_tc_sub(const hn_tensor *x, const void *w, int cout, const TcSubConv *sc, cudaStream_t st);
bool conv_halo_ok(const hn_tensor *x, const hn_conv *cv, const hn_tensor *y, bool upsample);
bool conv_pair_ok(int bn, int num_m_tiles, int num_kb, bool halo);                     // run this Cout tile on CTA pairs (cta_group::2)? (hn_conv_tc.cu)
struct HeadArgs {               // fused 1x1 classifier head of the halo kernel (hn_conv3x3_head_fwd)
    const float *w, *b;         // HOST pointers: [n][64] FP32, [n] FP32 or NULL (they become kernel parameters)
    float *out;                 // NCHW FP32 logits
    int n;
};
int conv2d_fwd_halo(const hn_tensor *x, const void *w, const hn_conv *cv, const hn_epilogue *ep, const hn_tensor *y, bool upsample,
                    cudaStream_t st, const HeadArgs *head = nullptr);

}  // namespace hn
