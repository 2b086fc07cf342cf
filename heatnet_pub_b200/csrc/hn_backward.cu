// Bandwidth-bound backward kernels on NHWC views: activation / BatchNorm2d(train) backward, max-pool backward
// (saved window indices), bilinear and pyramid-pool adjoints (gather form: deterministic, no atomics),
// zero-insertion for strided-conv dgrad, gradient unpack/accumulate into the OIHW FP32 masters.
// Same conventions as hn_bandwidth.cu: 8 channels per thread, row-decomposed grids, no 64-bit div/mod per element.
#include <stdlib.h>
#include <string.h>

#include "hn_common.cuh"

namespace hn {
using bf16_t = __nv_bfloat16;

static inline dim3 row_grid_b(int64_t rows, int64_t items_per_row, int threads = 256)
{
    int64_t chunks = cdiv(items_per_row, (int64_t)threads * 4);
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) chunks = 65535;
    return dim3((unsigned)rows, (unsigned)chunks);
}
static inline int wave_grid_b(int64_t work_items, int threads, int waves_per_sm = 8)
{
    int64_t want = cdiv(work_items, threads);
    int64_t cap = (int64_t)num_sms() * waves_per_sm;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

__device__ __forceinline__ float act_grad(float out, int act, float slope)
{
    // derivative of y = act(z) expressed through the saved OUTPUT (sign(out) == sign(z) for slope > 0)
    if (act == HN_ACT_RELU) return out > 0.f ? 1.f : 0.f;
    if (act == HN_ACT_LEAKY) return out >= 0.f ? 1.f : slope;
    return 1.f;
}

// ------------------------------------------------------------------------------------------------
// dz = dout * act'(out)   (conv + bias + activation layers: PSP bottleneck, critics)
// ------------------------------------------------------------------------------------------------
template <typename TG, typename TO>
__global__ void __launch_bounds__(256) act_bwd_kernel(const TG *__restrict__ dout, int ldg, const TO *__restrict__ out, int ldo, int act,
                                                      float slope, TG *__restrict__ dz, int ldz, int64_t npix, int C)
{
    const int ncv = C / 8;
    constexpr int kPixPerBlock = 64;
    const int64_t nblocks = (npix + kPixPerBlock - 1) / kPixPerBlock;
    for (int64_t pb = blockIdx.x; pb < nblocks; pb += gridDim.x)
        for (int i = threadIdx.x; i < kPixPerBlock * ncv; i += blockDim.x) {
            const int pl = i / ncv;
            const int64_t p = pb * kPixPerBlock + pl;
            if (p >= npix) break;
            const int c = (i - pl * ncv) * 8;
            float g[8], o[8];
            Vec8<TG>::load(dout + p * ldg + c, g);
            Vec8<TO>::load(out + p * ldo + c, o);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] *= act_grad(o[j], act, slope);
            Vec8<TG>::store(dz + p * ldz + c, g);
        }
}

// scalar variant for C % 8 != 0 (the critics' 1-channel classifier map)
template <typename TG, typename TO>
__global__ void __launch_bounds__(256) act_bwd_scalar_kernel(const TG *__restrict__ dout, int ldg, const TO *__restrict__ out, int ldo,
                                                             int act, float slope, TG *__restrict__ dz, int ldz, int64_t npix, int C)
{
    const int64_t total = npix * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t p = i / C;
        int c = (int)(i - p * C);
        float g = to_f32<TG>(dout[p * ldg + c]) * act_grad(to_f32<TO>(out[p * ldo + c]), act, slope);
        dz[p * ldz + c] = from_f32<TG>(g);
    }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm2d (train) backward.  z = raw*scale + shift (+ res), out = act(z).
//   reduce: s1[c] = sum dz, s2[c] = sum dz * xhat,  dz = dout*act'(out), xhat = (raw - mean)*invstd; for PReLU also
//           sp = sum over z<0 of dout * z (gradient of the shared slope)
//   apply : draw = gamma*invstd*(dz - s1/m - xhat*s2/m)   and   dres (+)= dz
// blockDim = (CVB, PL) as in channel_stats_kernel.
// ------------------------------------------------------------------------------------------------
// ZOUT = false: `out` is not read; the pre-activation z = fmaf(raw, fscale, fshift) is recomputed with the forward pass's own
// scale/shift vectors (bit-identical to what the forward activated; only valid for layers without a residual input) -- one
// tensor less to stream in both passes.
// blockDim = (CVB, PL): a thread owns ONE 4-channel group and streams its pixels four at a time (every load of a batch is
// issued before the first use, batches are held packed); 64 registers -> 4 CTAs per SM.
template <typename TG, typename TO, typename TR, bool ZOUT>
__global__ void __launch_bounds__(256, 4) bn_bwd_reduce_kernel(const TG *__restrict__ dout, int ldg, const TO *__restrict__ out, int ldo,
                                                               const TR *__restrict__ raw, int ldr, const float *__restrict__ mean,
                                                               const float *__restrict__ invstd, int act, float slope, const float *slope_ptr,
                                                               int64_t npix, int C, int64_t pix_per_cta, double *s1, double *s2, double *sprelu,
                                                               const float *__restrict__ fscale, const float *__restrict__ fshift)
{
    extern __shared__ float red[];  // [2][PL][CVB*4]
    const int cv = blockIdx.y * blockDim.x + threadIdx.x;
    const int ncv = C / 4;
    const int PL = blockDim.y, CVB = blockDim.x;
    const int grp = blockIdx.z;       // statistics group: pixels [grp*npix, (grp+1)*npix), per-group vectors [grp][C], sums [grp][2C+1]
    s1 += (int64_t)grp * (2 * C + 1);
    s2 += (int64_t)grp * (2 * C + 1);
    if (sprelu) sprelu += (int64_t)grp * (2 * C + 1);
    if (slope_ptr) slope = __ldg(slope_ptr);
    float a[4], b[4], mu[4], is[4], fs[4], fh[4], sp = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = b[i] = 0.f;
    const int64_t p_begin = (int64_t)grp * npix + (int64_t)blockIdx.x * pix_per_cta;
    const int64_t p_end = min(p_begin + pix_per_cta, (int64_t)(grp + 1) * npix);
    if (cv < ncv) {
        const int c = cv * 4;
        Vec4<float>::load(mean + grp * C + c, mu);
        Vec4<float>::load(invstd + grp * C + c, is);
        if (!ZOUT) {
            Vec4<float>::load(fscale + grp * C + c, fs);
            Vec4<float>::load(fshift + grp * C + c, fh);
        }
        auto one = [&](const float (&g)[4], float (&o)[4], const float (&r)[4]) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!ZOUT) o[i] = fmaf(r[i], fs[i], fh[i]);                   // z itself: act'(z) == act'(act(z)) for (leaky) ReLU
                if (sprelu && o[i] < 0.f) sp += g[i] * (ZOUT ? o[i] / slope : o[i]);      // z = out / slope on the negative side
                const float dz = g[i] * act_grad(o[i], act, slope);
                a[i] += dz;
                b[i] = fmaf(dz, (r[i] - mu[i]) * is[i], b[i]);
            }
        };
        int64_t p = p_begin + threadIdx.y;
        for (; p + 3 * PL < p_end; p += 4 * PL) {
            typename Vec4<TG>::raw_t gv[4];
            typename Vec4<TO>::raw_t ov[4];
            typename Vec4<TR>::raw_t rv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) gv[u] = Vec4<TG>::ldraw(dout + (p + u * PL) * ldg + c);
            if (ZOUT) {
#pragma unroll
                for (int u = 0; u < 4; ++u) ov[u] = Vec4<TO>::ldraw(out + (p + u * PL) * ldo + c);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) rv[u] = Vec4<TR>::ldraw(raw + (p + u * PL) * ldr + c);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float g[4], o[4], r[4];
                Vec4<TG>::unpack(gv[u], g);
                if (ZOUT) Vec4<TO>::unpack(ov[u], o);
                Vec4<TR>::unpack(rv[u], r);
                one(g, o, r);
            }
        }
        for (; p < p_end; p += PL) {
            float g[4], o[4], r[4];
            Vec4<TG>::load(dout + p * ldg + c, g);
            if (ZOUT) Vec4<TO>::load(out + p * ldo + c, o);
            Vec4<TR>::load(raw + p * ldr + c, r);
            one(g, o, r);
        }
    }
    float *ra = red, *rb = red + PL * CVB * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ra[(threadIdx.y * CVB + threadIdx.x) * 4 + i] = a[i];
        rb[(threadIdx.y * CVB + threadIdx.x) * 4 + i] = b[i];
    }
    __syncthreads();
    const int tid = threadIdx.y * CVB + threadIdx.x;
    for (int col = tid; col < CVB * 4; col += CVB * PL) {
        double x = 0.0, y = 0.0;
        for (int l = 0; l < PL; ++l) {
            x += (double)ra[l * CVB * 4 + col];
            y += (double)rb[l * CVB * 4 + col];
        }
        int ch = blockIdx.y * CVB * 4 + col;
        if (ch < C) {
            atomicAdd(s1 + ch, x);
            atomicAdd(s2 + ch, y);
        }
    }
    if (sprelu) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sp += __shfl_xor_sync(0xffffffffu, sp, o);
        if (((threadIdx.y * CVB + threadIdx.x) & 31) == 0 && sp != 0.f) atomicAdd(sprelu, (double)sp);
    }
}

// Same thread layout.  The per-channel coefficients (FP64 sums -> FP32) are formed once per CTA in shared memory and RE-READ from
// there for every batch of four pixels (volatile loads: 5 LDS.128 per 8-12 global loads), so they occupy registers only while a
// batch is being computed -- with them resident the kernel needed 124 registers (2 CTAs per SM) or spilled its addresses:
//   draw = ka*dz + kb*raw + kd,   ka = gamma*invstd, kb = -ka*invstd*s2/m, kd = -ka*s1/m - kb*mean;   z = raw*fs + fh
// DRES: 0 = no residual-branch gradient, 1 = dres = dz, 2 = dres += dz.
__device__ __forceinline__ void lds_f4_volatile(const float *p, float (&v)[4])
{
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}

template <typename TG, typename TO, typename TR, bool ZOUT, int DRES>
__global__ void __launch_bounds__(256, 4) bn_bwd_apply_kernel(const TG *__restrict__ dout, int ldg, const TO *__restrict__ out, int ldo,
                                                              const TR *__restrict__ raw, int ldr, const float *__restrict__ mean,
                                                              const float *__restrict__ invstd, const float *__restrict__ gamma,
                                                              const double *__restrict__ s1, const double *__restrict__ s2, double inv_count,
                                                              int act, float slope, const float *slope_ptr, TG *__restrict__ draw, int ldd,
                                                              TG *__restrict__ dres, int ldres, int npix, int C, int pix_per_cta,
                                                              float *__restrict__ dbeta, float *__restrict__ dgamma, float *__restrict__ dslope,
                                                              const double *__restrict__ sprelu, int param_accumulate,
                                                              const float *__restrict__ fscale, const float *__restrict__ fshift)
{
    __shared__ __align__(16) float s_k[5][128];
    const int PL = blockDim.y, CVB = blockDim.x;
    const int tid = threadIdx.y * CVB + threadIdx.x;
    const int grp = blockIdx.z, groups = gridDim.z;     // statistics group (see the reduce kernel)
    const int gstride = 2 * C + 1;
    if (tid < CVB * 4) {
        const int ch = blockIdx.y * CVB * 4 + tid;
        if (ch < C) {
            const double d1 = s1[grp * gstride + ch], d2 = s2[grp * gstride + ch];
            if (blockIdx.x == 0 && grp == 0) {  // parameter gradients straight from the FP64 sums (no extra launches), summed over the groups
                double t1 = 0.0, t2 = 0.0, tp = 0.0;
                for (int q = 0; q < groups; ++q) {
                    t1 += s1[q * gstride + ch];
                    t2 += s2[q * gstride + ch];
                    if (sprelu) tp += sprelu[q * gstride];
                }
                if (dbeta) dbeta[ch] = (param_accumulate ? dbeta[ch] : 0.f) + (float)t1;
                if (dgamma) dgamma[ch] = (param_accumulate ? dgamma[ch] : 0.f) + (float)t2;
                if (dslope && sprelu && ch == 0) dslope[0] = (param_accumulate ? dslope[0] : 0.f) + (float)tp;
            }
            const float m1 = (float)(d1 * inv_count), m2 = (float)(d2 * inv_count);
            const float is = invstd[grp * C + ch], mu = mean[grp * C + ch];
            const float ka = (gamma ? gamma[ch] : 1.f) * is;
            const float kb = -ka * is * m2;
            s_k[0][tid] = ka;
            s_k[1][tid] = kb;
            s_k[2][tid] = -ka * m1 - kb * mu;
            s_k[3][tid] = ZOUT ? 0.f : fscale[grp * C + ch];
            s_k[4][tid] = ZOUT ? 0.f : fshift[grp * C + ch];
        }
    }
    __syncthreads();
    const int cv = blockIdx.y * CVB + threadIdx.x;
    if (cv >= C / 4) return;
    const int c = cv * 4;
    if (slope_ptr) slope = __ldg(slope_ptr);
    dout += c; out += c; raw += c; draw += c;
    if (DRES) dres += c;
    const float *kp = &s_k[0][threadIdx.x * 4];
    const int p_begin = grp * npix + blockIdx.x * pix_per_cta;
    const int p_end = min(p_begin + pix_per_cta, (grp + 1) * npix);
    auto one = [&](int p, const float (&g)[4], const float (&o)[4], const float (&r)[4], const float (&e)[4], const float (&ka)[4],
                   const float (&kb)[4], const float (&kd)[4], const float (&fs)[4], const float (&fh)[4]) {
        float dx[4], dz[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dz[j] = g[j] * act_grad(ZOUT ? o[j] : fmaf(r[j], fs[j], fh[j]), act, slope);
            dx[j] = fmaf(ka[j], dz[j], fmaf(kb[j], r[j], kd[j]));
        }
        Vec4<TG>::store(draw + (int64_t)p * ldd, dx);
        if (DRES) {
            if (DRES == 2) {
#pragma unroll
                for (int j = 0; j < 4; ++j) dz[j] += e[j];
            }
            Vec4<TG>::store(dres + (int64_t)p * ldres, dz);
        }
    };
    int p = p_begin + threadIdx.y;
    for (; p + 3 * PL < p_end; p += 4 * PL) {
        typename Vec4<TG>::raw_t gv[4], ev[4];
        typename Vec4<TO>::raw_t ov[4];
        typename Vec4<TR>::raw_t rv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) gv[u] = Vec4<TG>::ldraw(dout + (int64_t)(p + u * PL) * ldg);
        if (ZOUT) {
#pragma unroll
            for (int u = 0; u < 4; ++u) ov[u] = Vec4<TO>::ldraw(out + (int64_t)(p + u * PL) * ldo);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) rv[u] = Vec4<TR>::ldraw(raw + (int64_t)(p + u * PL) * ldr);
        if (DRES == 2) {
#pragma unroll
            for (int u = 0; u < 4; ++u) ev[u] = Vec4<TG>::ldraw(dres + (int64_t)(p + u * PL) * ldres);
        }
        float ka[4], kb[4], kd[4], fs[4], fh[4];
        lds_f4_volatile(kp, ka);
        lds_f4_volatile(kp + 128, kb);
        lds_f4_volatile(kp + 256, kd);
        if (!ZOUT) {
            lds_f4_volatile(kp + 384, fs);
            lds_f4_volatile(kp + 512, fh);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float g[4], o[4], r[4], e[4];
            Vec4<TG>::unpack(gv[u], g);
            if (ZOUT) Vec4<TO>::unpack(ov[u], o);
            Vec4<TR>::unpack(rv[u], r);
            if (DRES == 2) Vec4<TG>::unpack(ev[u], e);
            one(p + u * PL, g, o, r, e, ka, kb, kd, fs, fh);
        }
    }
    if (p < p_end) {
        float ka[4], kb[4], kd[4], fs[4], fh[4];
        lds_f4_volatile(kp, ka);
        lds_f4_volatile(kp + 128, kb);
        lds_f4_volatile(kp + 256, kd);
        if (!ZOUT) {
            lds_f4_volatile(kp + 384, fs);
            lds_f4_volatile(kp + 512, fh);
        }
        for (; p < p_end; p += PL) {
            float g[4], o[4], r[4], e[4];
            Vec4<TG>::load(dout + (int64_t)p * ldg, g);
            if (ZOUT) Vec4<TO>::load(out + (int64_t)p * ldo, o);
            Vec4<TR>::load(raw + (int64_t)p * ldr, r);
            if (DRES == 2) Vec4<TG>::load(dres + (int64_t)p * ldres, e);
            one(p, g, o, r, e, ka, kb, kd, fs, fh);
        }
    }
}

// y (+)= x   (gradient accumulation between NHWC views; x may be FP32 or the view's dtype)
template <typename TX, typename TY>
__global__ void __launch_bounds__(256) accumulate_kernel(const TX *__restrict__ x, int ldx, TY *__restrict__ y, int ldy, int accumulate,
                                                         int64_t npix, int C)
{
    const int ncv = C / 8;
    constexpr int kPixPerBlock = 64;
    const int64_t nblocks = (npix + kPixPerBlock - 1) / kPixPerBlock;
    for (int64_t pb = blockIdx.x; pb < nblocks; pb += gridDim.x)
        for (int i = threadIdx.x; i < kPixPerBlock * ncv; i += blockDim.x) {
            const int pl = i / ncv;
            const int64_t p = pb * kPixPerBlock + pl;
            if (p >= npix) break;
            const int c = (i - pl * ncv) * 8;
            float a[8];
            Vec8<TX>::load(x + p * ldx + c, a);
            if (accumulate) {
                float b[8];
                Vec8<TY>::load(y + p * ldy + c, b);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] += b[j];
            }
            Vec8<TY>::store(y + p * ldy + c, a);
        }
}

template <typename TX, typename TY>
__global__ void __launch_bounds__(256) accumulate_scalar_kernel(const TX *__restrict__ x, int ldx, TY *__restrict__ y, int ldy, int accumulate,
                                                                int64_t npix, int C)
{
    const int64_t total = npix * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t p = i / C;
        int c = (int)(i - p * C);
        float v = to_f32<TX>(x[p * ldx + c]);
        if (accumulate) v += to_f32<TY>(y[p * ldy + c]);
        y[p * ldy + c] = from_f32<TY>(v);
    }
}

// ------------------------------------------------------------------------------------------------
// max-pool 3x3 s2 p1: forward that also records the winning tap (0..8, first max in window order like ATen),
// and the backward gather: every input pixel looks at the <= 4 windows that contain it.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) maxpool_idx_kernel(const T *__restrict__ x, int ldx, T *__restrict__ y, int ldy, uint8_t *__restrict__ idx,
                                                          int N, int H, int W, int Ho, int Wo, int C)
{
    const int ncv = C / 8;
    const int row = blockIdx.x;
    const int n = row / Ho, ho = row - n * Ho;
    const int items = Wo * ncv;
    const T *xin = x + (int64_t)n * H * W * ldx;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < items; i += gridDim.y * blockDim.x) {
        const int wo = i / ncv, c = (i - wo * ncv) * 8;
        float m[8];
        uint8_t w8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; w8[j] = 0; }
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
                for (int j = 0; j < 8; ++j)
                    if (v[j] > m[j]) { m[j] = v[j]; w8[j] = (uint8_t)(r * 3 + s); }   // strictly greater: first max wins (ATen)
            }
        }
        const int64_t po = (int64_t)row * Wo + wo;
        Vec8<T>::store(y + po * ldy + c, m);
        uint2 packed;
        uint8_t *pb = reinterpret_cast<uint8_t *>(&packed);
#pragma unroll
        for (int j = 0; j < 8; ++j) pb[j] = w8[j];
        *reinterpret_cast<uint2 *>(idx + po * C + c) = packed;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const T *__restrict__ dy, int ldy, const uint8_t *__restrict__ idx, T *__restrict__ dx,
                                                          int ldx, int accumulate, int N, int H, int W, int Ho, int Wo, int C)
{
    const int ncv = C / 8;
    const int row = blockIdx.x;                    // n*H + hi
    const int n = row / H, hi = row - n * H;
    const int items = W * ncv;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < items; i += gridDim.y * blockDim.x) {
        const int wi = i / ncv, c = (i - wi * ncv) * 8;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        // windows ho with 2*ho-1 <= hi <= 2*ho+1
        for (int ho = hi / 2; ho <= (hi + 1) / 2; ++ho) {
            if (ho >= Ho) continue;
            const int r = hi - (2 * ho - 1);
            for (int wo = wi / 2; wo <= (wi + 1) / 2; ++wo) {
                if (wo >= Wo) continue;
                const int s = wi - (2 * wo - 1);
                const int64_t po = ((int64_t)n * Ho + ho) * Wo + wo;
                const uint2 packed = *reinterpret_cast<const uint2 *>(idx + po * C + c);
                const uint8_t *pb = reinterpret_cast<const uint8_t *>(&packed);
                float g[8];
                Vec8<T>::load(dy + po * ldy + c, g);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (pb[j] == r * 3 + s) acc[j] += g[j];
            }
        }
        T *d = dx + ((int64_t)row * W + wi) * ldx + c;
        if (accumulate) {
            float e[8];
            Vec8<T>::load(d, e);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += e[j];
        }
        Vec8<T>::store(d, acc);
    }
}

// ------------------------------------------------------------------------------------------------
// bilinear adjoint (gather): dx[n, yi, xi, c] = sum over outputs (yo, xo) whose source taps include (yi, xi)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src_b(int dst, float scale, int in_size, int &i0, int &i1, float &l1)
{
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - (float)i0;
}
// candidate output range [lo, hi] that can touch input index i (conservative; exact membership is re-checked)
__device__ __forceinline__ void out_range(int i, float scale, int out_size, int &lo, int &hi)
{
    float a = ((float)i - 1.0f + 0.5f) / scale - 0.5f;
    float b = ((float)i + 1.0f + 0.5f) / scale - 0.5f;
    lo = (int)floorf(a) - 1;
    hi = (int)ceilf(b) + 1;
    if (lo < 0) lo = 0;
    if (hi > out_size - 1) hi = out_size - 1;
}

template <typename TG, typename TX, bool kVec>
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const TG *__restrict__ dy, int ldy, TX *__restrict__ dx, int ldx, int accumulate,
                                                           int N, int H, int W, int Ho, int Wo, int C, float sh, float sw)
{
    constexpr int V = kVec ? 8 : 1;
    const int ncv = C / V;
    const int row = blockIdx.x;               // n*H + yi
    const int n = row / H, yi = row - n * H;
    int ylo, yhi;
    out_range(yi, sh, Ho, ylo, yhi);
    const int items = W * ncv;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < items; i += gridDim.y * blockDim.x) {
        const int xi = i / ncv, c = (i - xi * ncv) * V;
        int xlo, xhi;
        out_range(xi, sw, Wo, xlo, xhi);
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0.f;
        for (int yo = ylo; yo <= yhi; ++yo) {
            int y0, y1;
            float ly;
            bilinear_src_b(yo, sh, H, y0, y1, ly);
            const float wy = (y0 == yi ? 1.f - ly : 0.f) + (y1 == yi ? ly : 0.f);
            if (wy == 0.f) continue;
            const TG *grow = dy + ((int64_t)n * Ho + yo) * Wo * ldy + c;
            for (int xo = xlo; xo <= xhi; ++xo) {
                int x0, x1;
                float lx;
                bilinear_src_b(xo, sw, W, x0, x1, lx);
                const float wx = (x0 == xi ? 1.f - lx : 0.f) + (x1 == xi ? lx : 0.f);
                if (wx == 0.f) continue;
                const float wgt = wy * wx;
                if constexpr (kVec) {
                    float g[8];
                    Vec8<TG>::load(grow + (int64_t)xo * ldy, g);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, g[j], acc[j]);
                } else {
                    acc[0] = fmaf(wgt, to_f32<TG>(grow[(int64_t)xo * ldy]), acc[0]);
                }
            }
        }
        TX *d = dx + ((int64_t)row * W + xi) * ldx + c;
        if constexpr (kVec) {
            if (accumulate) {
                float e[8];
                Vec8<TX>::load(d, e);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += e[j];
            }
            Vec8<TX>::store(d, acc);
        } else {
            float v = acc[0] + (accumulate ? to_f32<TX>(d[0]) : 0.f);
            d[0] = from_f32<TX>(v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// exact 2x adjoint (PSPUpsample): input index i receives  .25*dy[2i-1] + .75*dy[2i] + .75*dy[2i+1] + .25*dy[2i+2]  along each
// axis; at the borders the clamped taps fold in (weight 1 for dy[0] -> 0 and dy[2H-1] -> H-1).  A thread produces a 2x2 block
// of dx for 8 channels from the 6x6 block of dy around it: 9 loads per output instead of 16.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void up2_adjoint_weights(int i, int size, float (&w)[4])
{
    // weights of dy[2i-1], dy[2i], dy[2i+1], dy[2i+2] for input index i (0 when the output index does not exist / does not tap i)
    w[0] = i >= 1 ? 0.25f : 0.f;
    w[1] = i >= 1 ? 0.75f : 1.f;
    w[2] = i + 1 <= size - 1 ? 0.75f : 1.f;
    w[3] = i + 1 <= size - 1 ? 0.25f : 0.f;
}

template <typename TG, typename TX>
__global__ void __launch_bounds__(256) bilinear_up2_bwd_kernel(const TG *__restrict__ dy, int ldy, TX *__restrict__ dx, int ldx, int accumulate,
                                                               int N, int H, int W, int C)
{
    const int ncv = C / 8;
    const int Ho = 2 * H, Wo = 2 * W;
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    const int row = blockIdx.x;                 // n*H2 + tile row
    const int n = row / H2, yi = (row - n * H2) * 2;
    float wyA[4], wyB[4];
    up2_adjoint_weights(yi, H, wyA);
    up2_adjoint_weights(yi + 1, H, wyB);
    const bool rowB = yi + 1 < H;
    const int items = W2 * ncv;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < items; i += gridDim.y * blockDim.x) {
        const int xt = i / ncv, c = (i - xt * ncv) * 8;
        const int xi = xt * 2;
        const bool colB = xi + 1 < W;
        float wxA[4], wxB[4];
        up2_adjoint_weights(xi, W, wxA);
        up2_adjoint_weights(xi + 1, W, wxB);
        float a00[8], a01[8], a10[8], a11[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a00[j] = a01[j] = a10[j] = a11[j] = 0.f;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const int yo = 2 * yi - 1 + r;
            if (yo < 0 || yo >= Ho) continue;
            const TG *grow = dy + (((int64_t)n * Ho + yo) * Wo) * ldy + c;
            float ra[8], rb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) ra[j] = rb[j] = 0.f;
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const int xo = 2 * xi - 1 + q;
                if (xo < 0 || xo >= Wo) continue;
                float g[8];
                Vec8<TG>::load(grow + (int64_t)xo * ldy, g);
                if (q < 4) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) ra[j] = fmaf(wxA[q], g[j], ra[j]);
                }
                if (q >= 2) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) rb[j] = fmaf(wxB[q - 2], g[j], rb[j]);
                }
            }
            if (r < 4) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    a00[j] = fmaf(wyA[r], ra[j], a00[j]);
                    a01[j] = fmaf(wyA[r], rb[j], a01[j]);
                }
            }
            if (r >= 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    a10[j] = fmaf(wyB[r - 2], ra[j], a10[j]);
                    a11[j] = fmaf(wyB[r - 2], rb[j], a11[j]);
                }
            }
        }
        auto put = [&](int y, int x, float (&v)[8]) {
            TX *d = dx + (((int64_t)n * H + y) * W + x) * ldx + c;
            if (accumulate) {
                float e[8];
                Vec8<TX>::load(d, e);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += e[j];
            }
            Vec8<TX>::store(d, v);
        };
        put(yi, xi, a00);
        if (colB) put(yi, xi + 1, a01);
        if (rowB) put(yi + 1, xi, a10);
        if (rowB && colB) put(yi + 1, xi + 1, a11);
    }
}

// Walk form of the exact-2x adjoint for the large decoder maps: a thread owns one dx row segment (UP2_SEG input columns) of one
// 8-channel group and walks it left to right.  Per input column it loads the two NEW dy columns of the four dy rows above it
// (8 independent 16-byte loads), folds the rows with the row weights into two column sums, and combines them with the two
// column sums kept from the previous step: the 4-tap column window slides through registers.  Every dy element is loaded by
// exactly two threads (vertically adjacent rows share two dy rows; the second read hits L1/L2) instead of 2.25 times in 6x6 blocks
// with ~100 live registers; measured on the three PSPUpsample adjoints of a training step: 252 us -> see profiles/.
constexpr int UP2_SEG = 16;
template <typename TG, typename TX>
__global__ void __launch_bounds__(128) bilinear_up2_bwd_walk_kernel(const TG *__restrict__ dy, int ldy, TX *__restrict__ dx, int ldx, int accumulate,
                                                                    int N, int H, int W, int C)
{
    const int ncv = C / 8;
    const int Ho = 2 * H, Wo = 2 * W;
    const int nseg = (W + UP2_SEG - 1) / UP2_SEG;
    const int row = blockIdx.x;                 // n*H + yi
    const int n = row / H, yi = row - n * H;
    const int item = blockIdx.y * blockDim.x + threadIdx.x;
    if (item >= nseg * ncv) return;
    const int seg = item / ncv, c = (item - seg * ncv) * 8;
    float wy[4];
    up2_adjoint_weights(yi, H, wy);
    const TG *gbase = dy + ((int64_t)n * Ho * Wo) * ldy + c;
    // column sum over the (up to) four dy rows 2yi-1 .. 2yi+2 at dy column xo (zero outside the map)
    auto colsum = [&](int xo, float (&t)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = 0.f;
        if (xo < 0 || xo >= Wo) return;
        float g[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int yo = 2 * yi - 1 + r;
            if (yo >= 0 && yo < Ho) Vec8<TG>::load(gbase + ((int64_t)yo * Wo + xo) * ldy, g[r]);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) g[r][j] = 0.f;
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = fmaf(wy[r], g[r][j], t[j]);
    };
    const int x0 = seg * UP2_SEG, x1 = min(x0 + UP2_SEG, W);
    float t0[8], t1[8], t2[8], t3[8];
    colsum(2 * x0 - 1, t0);
    colsum(2 * x0, t1);
    TX *drow = dx + ((int64_t)row * W) * ldx + c;
    for (int xi = x0; xi < x1; ++xi) {
        colsum(2 * xi + 1, t2);
        colsum(2 * xi + 2, t3);
        float wx[4], v[8];
        up2_adjoint_weights(xi, W, wx);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(wx[0], t0[j], fmaf(wx[1], t1[j], fmaf(wx[2], t2[j], wx[3] * t3[j])));
        TX *d = drow + (int64_t)xi * ldx;
        if (accumulate) {
            float e[8];
            Vec8<TX>::load(d, e);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += e[j];
        }
        Vec8<TX>::store(d, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { t0[j] = t2[j]; t1[j] = t3[j]; }
    }
}

// ------------------------------------------------------------------------------------------------
// large magnification (PSP priors s x s -> h x w, critic maps x32): the gather form leaves only N*H*W*C/8 threads with a long
// serial loop each.  Separable two-pass adjoint instead, every element of dy read once by a fully parallel pass:
//   pass A  tmp[n, yo, xi, c] = sum_xo wx(xo, xi) dy[n, yo, xo, c]      (FP32 workspace, N*Ho*W*C)
//   pass B  dx[n, yi, xi, c] (+)= sum_yo wy(yo, yi) tmp[n, yo, xi, c]
// ------------------------------------------------------------------------------------------------
constexpr int kSepMaxW = 8;
template <typename TG, int MAXW>
__global__ void __launch_bounds__(256) bilinear_bwd_rows_vec_kernel(const TG *__restrict__ dy, int ldy, float *__restrict__ tmp, int64_t nrows,
                                                                    int W, int Wo, int C, float sw)
{
    constexpr int kSepMaxW = MAXW;          // accumulator columns held in registers: W <= MAXW (2 / 4 / 8: fewer registers, more loads in flight)
    // one thread: one (n, yo) row x one 8-channel group, all W <= kSepMaxW input columns in registers
    const int ncv = C / 8;
    const int64_t total = nrows * ncv;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / ncv;
        const int c = (int)(t - row * ncv) * 8;
        float acc[kSepMaxW][8];
#pragma unroll
        for (int x = 0; x < kSepMaxW; ++x)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[x][j] = 0.f;
        const TG *g0 = dy + row * Wo * ldy + c;
        auto fold = [&](int xo, const float (&g)[8]) {
            int x0, x1;
            float lx;
            bilinear_src_b(xo, sw, W, x0, x1, lx);
#pragma unroll
            for (int x = 0; x < kSepMaxW; ++x) {
                const float w = (x == x0 ? 1.f - lx : 0.f) + (x == x1 ? lx : 0.f);
                if (x < W) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[x][j] = fmaf(w, g[j], acc[x][j]);
                }
            }
        };
        int xo = 0;
        for (; xo + 3 < Wo; xo += 4) {              // four output columns (64 B) in flight per thread
            float g[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) Vec8<TG>::load(g0 + (int64_t)(xo + u) * ldy, g[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) fold(xo + u, g[u]);
        }
        for (; xo < Wo; ++xo) {
            float g[8];
            Vec8<TG>::load(g0 + (int64_t)xo * ldy, g);
            fold(xo, g);
        }
#pragma unroll
        for (int x = 0; x < kSepMaxW; ++x)
            if (x < W) Vec8<float>::store(tmp + (row * W + x) * C + c, acc[x]);
    }
}

// scalar-channel variant (critic maps, C == 1 typically): one thread per (row, xi, c), gathers its output columns
template <typename TG>
__global__ void __launch_bounds__(256) bilinear_bwd_rows_scalar_kernel(const TG *__restrict__ dy, int ldy, float *__restrict__ tmp, int64_t nrows,
                                                                       int W, int Wo, int C, float sw)
{
    const int64_t total = nrows * W * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        const int64_t rx = t / C;
        const int xi = (int)(rx % W);
        const int64_t row = rx / W;
        int xlo, xhi;
        out_range(xi, sw, Wo, xlo, xhi);
        float acc = 0.f;
        const TG *g0 = dy + row * Wo * ldy + c;
        for (int xo = xlo; xo <= xhi; ++xo) {
            int x0, x1;
            float lx;
            bilinear_src_b(xo, sw, W, x0, x1, lx);
            const float w = (x0 == xi ? 1.f - lx : 0.f) + (x1 == xi ? lx : 0.f);
            if (w != 0.f) acc = fmaf(w, to_f32<TG>(g0[(int64_t)xo * ldy]), acc);
        }
        tmp[t] = acc;
    }
}

// pass B: one thread per element of dx (V channels), sums its output rows from tmp
template <typename TX, bool kVec>
__global__ void __launch_bounds__(256) bilinear_bwd_cols_kernel(const float *__restrict__ tmp, TX *__restrict__ dx, int ldx, int accumulate, int N,
                                                                int H, int W, int Ho, int C, float sh)
{
    constexpr int V = kVec ? 8 : 1;
    const int ncv = C / V;
    const int64_t total = (int64_t)N * H * W * ncv;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % ncv) * V;
        int64_t r = t / ncv;
        const int xi = (int)(r % W);
        r /= W;
        const int yi = (int)(r % H);
        const int n = (int)(r / H);
        int ylo, yhi;
        out_range(yi, sh, Ho, ylo, yhi);
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0.f;
        for (int yo = ylo; yo <= yhi; ++yo) {
            int y0, y1;
            float ly;
            bilinear_src_b(yo, sh, H, y0, y1, ly);
            const float w = (y0 == yi ? 1.f - ly : 0.f) + (y1 == yi ? ly : 0.f);
            if (w == 0.f) continue;
            const float *src = tmp + (((int64_t)n * Ho + yo) * W + xi) * C + c;
            if constexpr (kVec) {
                float g[8];
                Vec8<float>::load(src, g);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, g[j], acc[j]);
            } else {
                acc[0] = fmaf(w, src[0], acc[0]);
            }
        }
        TX *d = dx + (((int64_t)n * H + yi) * W + xi) * ldx + c;
        if constexpr (kVec) {
            if (accumulate) {
                float e[8];
                Vec8<TX>::load(d, e);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += e[j];
            }
            Vec8<TX>::store(d, acc);
        } else {
            d[0] = from_f32<TX>(acc[0] + (accumulate ? to_f32<TX>(d[0]) : 0.f));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pyramid pool adjoint: dx[n,h,w,c] (+)= sum over sizes s and bins (i,j) containing (h,w) of dbin / area
// dpool: the forward's per-size dense blocks [N][s][s][C]
// ------------------------------------------------------------------------------------------------
struct PoolSizesB {
    int n, s[4];
};
template <typename T>
__global__ void __launch_bounds__(256) pyramid_bwd_kernel(const T *__restrict__ dpool, T *__restrict__ dx, int ldx, int accumulate, int N, int H,
                                                          int W, int C, PoolSizesB ps)
{
    const int ncv = C / 8;
    const int row = blockIdx.x;       // n*H + h
    const int n = row / H, h = row - n * H;
    const int items = W * ncv;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < items; i += gridDim.y * blockDim.x) {
        const int w = i / ncv, c = (i - w * ncv) * 8;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        int64_t base = 0;
        for (int a = 0; a < ps.n; ++a) {
            const int s = ps.s[a];
            // bins bi with floor(bi*H/s) <= h < ceil((bi+1)*H/s): at most two, around h*s/H
            int b0 = (h * s) / H;
            for (int bi = max(b0 - 1, 0); bi <= min(b0 + 1, s - 1); ++bi) {
                const int h0 = (bi * H) / s, h1 = ((bi + 1) * H + s - 1) / s;
                if (h < h0 || h >= h1) continue;
                int c0 = (w * s) / W;
                for (int bj = max(c0 - 1, 0); bj <= min(c0 + 1, s - 1); ++bj) {
                    const int w0 = (bj * W) / s, w1 = ((bj + 1) * W + s - 1) / s;
                    if (w < w0 || w >= w1) continue;
                    float g[8];
                    Vec8<T>::load(dpool + (base * N + ((int64_t)n * s + bi) * s + bj) * C + c, g);
                    const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(g[j], inv, acc[j]);
                }
            }
            base += (int64_t)s * s;
        }
        T *d = dx + ((int64_t)row * W + w) * ldx + c;
        if (accumulate) {
            float e[8];
            Vec8<T>::load(d, e);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += e[j];
        }
        Vec8<T>::store(d, acc);
    }
}

// Row form of the same adjoint: a CTA owns one image row (n, h) and a block of 8-channel groups.  Which bins contain h is
// row-constant; which bins contain each column w (<= 2 per size: adaptive bins overlap by at most one pixel) and their 1/width
// are tabulated once per CTA in shared memory, so the inner loop is loads and FMAs only (the element-wise kernel above spent
// ~100 integer instructions on divisions per 16 bytes written: 690 us for a 210 MB gradient).
template <typename T>
__global__ void __launch_bounds__(256) pyramid_bwd_rows_kernel(const T *__restrict__ dpool, T *__restrict__ dx, int ldx, int accumulate, int N,
                                                               int H, int W, int C, PoolSizesB ps)
{
    extern __shared__ float tabf[];                       // [4][W][2] 1/width (0 = no such bin), then int [4][W] first bin
    int *tabi = reinterpret_cast<int *>(tabf + 4 * W * 2);
    for (int i = threadIdx.x; i < ps.n * W; i += blockDim.x) {
        const int a = i / W, w = i - a * W, s = ps.s[a];
        const int c0 = (w * s) / W;
        int first = -1, k = 0;
        float inv[2] = {0.f, 0.f};
        for (int bj = max(c0 - 1, 0); bj <= min(c0 + 1, s - 1) && k < 2; ++bj) {
            const int w0 = (bj * W) / s, w1 = ((bj + 1) * W + s - 1) / s;
            if (w < w0 || w >= w1) continue;
            if (first < 0) first = bj;
            inv[k++] = 1.f / (float)(w1 - w0);
        }
        tabi[a * W + w] = first;
        tabf[(a * W + w) * 2] = inv[0];
        tabf[(a * W + w) * 2 + 1] = inv[1];
    }
    __syncthreads();
    const int row = blockIdx.x;       // n*H + h
    const int n = row / H, h = row - n * H;
    const int cv = blockIdx.y * blockDim.x + threadIdx.x;
    if (cv >= C / 8) return;
    const int c = cv * 8;
    // row part: per size up to two bins containing h, with 1/height
    int rb[4][2];
    float rinv[4][2];
    int64_t sbase[4];
    int64_t base = 0;
    for (int a = 0; a < 4; ++a) {
        rb[a][0] = rb[a][1] = -1;
        rinv[a][0] = rinv[a][1] = 0.f;
        sbase[a] = base;
        if (a >= ps.n) continue;
        const int s = ps.s[a];
        const int b0 = (h * s) / H;
        int k = 0;
        for (int bi = max(b0 - 1, 0); bi <= min(b0 + 1, s - 1) && k < 2; ++bi) {
            const int h0 = (bi * H) / s, h1 = ((bi + 1) * H + s - 1) / s;
            if (h < h0 || h >= h1) continue;
            rb[a][k] = bi;
            rinv[a][k] = 1.f / (float)(h1 - h0);
            ++k;
        }
        base += (int64_t)s * s * N;
    }
    T *drow = dx + (int64_t)row * W * ldx + c;
#pragma unroll 2
    for (int w = 0; w < W; ++w) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        if (accumulate) Vec8<T>::load(drow + (int64_t)w * ldx, acc);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            if (a >= ps.n) break;
            const int s = ps.s[a];
            const int bj0 = tabi[a * W + w];
#pragma unroll
            for (int kh = 0; kh < 2; ++kh) {
                if (rb[a][kh] < 0) continue;
#pragma unroll
                for (int kw = 0; kw < 2; ++kw) {
                    const float iw = tabf[(a * W + w) * 2 + kw];
                    if (iw == 0.f) continue;
                    float g[8];
                    Vec8<T>::load(dpool + (sbase[a] + ((int64_t)n * s + rb[a][kh]) * s + bj0 + kw) * C + c, g);
                    const float inv = rinv[a][kh] * iw;
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(g[j], inv, acc[j]);
                }
            }
        }
        Vec8<T>::store(drow + (int64_t)w * ldx, acc);
    }
}

// ------------------------------------------------------------------------------------------------
// zero insertion: up[n, h*stride, w*stride, c] = x[n,h,w,c], zeros elsewhere (strided-conv dgrad as a stride-1 conv)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) dilate_scalar_kernel(const T *__restrict__ x, int ldx, T *__restrict__ up, int ldu, int N, int H, int W,
                                                            int Hu, int Wu, int C, int stride)
{
    const int64_t total = (int64_t)N * Hu * Wu * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        int64_t p = i / C;
        int wu = (int)(p % Wu), hu = (int)((p / Wu) % Hu), n = (int)(p / ((int64_t)Wu * Hu));
        T v = from_f32<T>(0.f);
        if (hu % stride == 0 && wu % stride == 0 && hu / stride < H && wu / stride < W)
            v = x[(((int64_t)n * H + hu / stride) * W + wu / stride) * ldx + c];
        up[p * ldu + c] = v;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) dilate_kernel(const T *__restrict__ x, int ldx, T *__restrict__ up, int ldu, int N, int H, int W, int Hu,
                                                     int Wu, int C, int stride)
{
    const int ncv = C / 8;
    const int row = blockIdx.x;   // n*Hu + hu
    const int n = row / Hu, hu = row - n * Hu;
    const bool rowlive = (hu % stride) == 0 && hu / stride < H;
    const int items = Wu * ncv;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < items; i += gridDim.y * blockDim.x) {
        const int wu = i / ncv, c = (i - wu * ncv) * 8;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
        if (rowlive && (wu % stride) == 0 && wu / stride < W)
            Vec8<T>::load(x + (((int64_t)n * H + hu / stride) * W + wu / stride) * ldx + c, v);
        Vec8<T>::store(up + ((int64_t)row * Wu + wu) * ldu + c, v);
    }
}

// packed FP32 weight gradient [cout_pad][kpad] (k = (r*S+s)*Cin + c) -> OIHW FP32, optionally accumulating
__global__ void __launch_bounds__(256) unpack_wgrad_kernel(const float *__restrict__ packed, float *__restrict__ grad, int cout, int cin, int R,
                                                           int S, int kpad, int accumulate)
{
    const int64_t total = (int64_t)cout * cin * R * S;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int s = (int)(i % S);
        int r = (int)((i / S) % R);
        int c = (int)((i / ((int64_t)S * R)) % cin);
        int o = (int)(i / ((int64_t)S * R * cin));
        float v = packed[(int64_t)o * kpad + (r * S + s) * cin + c];
        grad[i] = accumulate ? grad[i] + v : v;
    }
}

// dgrad weight pack: OIHW FP32 -> [cin_pad][kpad'] with k' = ((R-1-r)*S + (S-1-s))*Cout + o  (flipped taps, in/out swapped)
template <typename T>
__global__ void __launch_bounds__(256) pack_weight_dgrad_kernel(const float *__restrict__ w, T *__restrict__ dst, int cout, int cin, int R, int S,
                                                                int cin_pad, int kpad)
{
    const int64_t total = (int64_t)cin_pad * kpad;
    const int K = R * S * cout;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int ci = (int)(i / kpad), k = (int)(i - (int64_t)ci * kpad);
        float v = 0.f;
        if (ci < cin && k < K) {
            int tap = k / cout, o = k - tap * cout;
            int rf = tap / S, sf = tap - rf * S;
            int r = R - 1 - rf, s = S - 1 - sf;
            v = __ldg(w + (((int64_t)o * cin + ci) * R + r) * S + s);
        }
        dst[i] = from_f32<T>(v);
    }
}

// ------------------------------------------------------------------------------------------------
// Stride-2 dgrad by parity phases.  Forward: y[o] = sum_r x[2o - pad + r] w[r].  For input position i = 2m + rho only the taps
// r = phi, phi + 2, ... (phi = (rho + pad) mod 2) can reach it, from o = m + delta - t with r = phi + 2t, delta = (rho + pad - phi)/2:
//     dx[2m + rho] = sum_{u=0}^{T-1} dy[m - pad' + u] * w[phi + 2(T-1-u)],   T = ceil((k - phi)/2),  pad' = T - 1 - delta
// i.e. a stride-1 correlation of dY with a T-tap sub-filter, per axis.  The four (rho_y, rho_x) phases together touch every
// filter tap exactly once: k*k tap evaluations per output pixel QUAD instead of 4*k*k on the zero-inserted gradient.
// ------------------------------------------------------------------------------------------------
struct PhaseAxis { int T, padp, phi, count; };     // taps, padding, first filter tap, lattice extent
static PhaseAxis phase_axis(int rho, int k, int pad, int in_size)
{
    PhaseAxis a;
    a.phi = (rho + pad) & 1;
    a.T = a.phi < k ? (k - a.phi + 1) / 2 : 0;
    const int delta = (rho + pad - a.phi) / 2;
    a.padp = a.T - 1 - delta;
    a.count = in_size > rho ? (in_size - rho + 1) / 2 : 0;
    return a;
}

// sub-filter pack of one phase: dst[c][k], k = (uy*Tx + ux)*Cout + o  <-  w[o][c][phi_y + 2(Ty-1-uy)][phi_x + 2(Tx-1-ux)]
template <typename T>
__global__ void __launch_bounds__(256) pack_weight_dgrad_phase_kernel(const float *__restrict__ w, T *__restrict__ dst, int cout, int cin, int R,
                                                                      int S, int Ty, int Tx, int phiy, int phix, int cin_pad, int kpad)
{
    const int64_t total = (int64_t)cin_pad * kpad;
    const int K = Ty * Tx * cout;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int ci = (int)(i / kpad), k = (int)(i - (int64_t)ci * kpad);
        float v = 0.f;
        if (ci < cin && k < K) {
            int tap = k / cout, o = k - tap * cout;
            int uy = tap / Tx, ux = tap - uy * Tx;
            int r = phiy + 2 * (Ty - 1 - uy), s2 = phix + 2 * (Tx - 1 - ux);
            v = __ldg(w + (((int64_t)o * cin + ci) * R + r) * S + s2);
        }
        dst[i] = from_f32<T>(v);
    }
}

// zeros on one parity lattice of an NHWC view (phases no filter tap reaches, e.g. three of the four of a 1x1 stride-2 conv)
template <typename T>
__global__ void __launch_bounds__(256) zero_lattice_kernel(T *__restrict__ y, int64_t pix_stride, int64_t row_stride, int64_t img_stride, int N,
                                                           int Hl, int Wl, int C)
{
    const int64_t total = (int64_t)N * Hl * Wl * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        int64_t q = i / C;
        int wl = (int)(q % Wl), hl = (int)((q / Wl) % Hl), n = (int)(q / ((int64_t)Wl * Hl));
        y[n * img_stride + hl * row_stride + wl * pix_stride + c] = from_f32<T>(0.f);
    }
}

// ------------------------------------------------------------------------------------------------
// All weight packs of a training step in ONE launch (hn_pack_weights_multi): after every optimizer step each live convolution
// needs its forward pack, its dgrad pack and (stride 2) its phase packs rebuilt from the FP32 master -- ~170 tiny launches per
// step otherwise.  A block handles PACK_CHUNK consecutive destination elements of one job (block_map: job, chunk).
// ------------------------------------------------------------------------------------------------
constexpr int PACK_CHUNK = 4096;

__device__ __forceinline__ float pack_source(const hn_pack_job &j, int row, int k)
{
    // row = destination row, k = destination column; returns the master-weight element that belongs there (0 for padding)
    if (j.kind == 0) {                                   // forward: [cout_pad][kpad], k = (r*S+s)*Cin + c
        if (row >= j.cout || k >= j.r * j.s * j.cin) return 0.f;
        const int tap = k / j.cin, c = k - tap * j.cin;
        const int r = tap / j.s, s2 = tap - r * j.s;
        return __ldg(j.src + (((int64_t)row * j.cin + c) * j.r + r) * j.s + s2);
    }
    if (j.kind == 1) {                                   // dgrad: [cin_pad][kpad'], k' = ((R-1-r)*S + (S-1-s))*Cout + o
        if (row >= j.cin || k >= j.r * j.s * j.cout) return 0.f;
        const int tap = k / j.cout, o = k - tap * j.cout;
        const int rf = tap / j.s, sf = tap - rf * j.s;
        return __ldg(j.src + (((int64_t)o * j.cin + row) * j.r + (j.r - 1 - rf)) * j.s + (j.s - 1 - sf));
    }
    // kind 2: stride-2 dgrad phase: [cin_pad][kpad''], k'' = (uy*Tx + ux)*Cout + o  (ty, tx, phiy, phix precomputed by the host)
    if (row >= j.cin || k >= j.ty * j.tx * j.cout) return 0.f;
    const int tap = k / j.cout, o = k - tap * j.cout;
    const int uy = tap / j.tx, ux = tap - uy * j.tx;
    return __ldg(j.src + (((int64_t)o * j.cin + row) * j.r + (j.phiy + 2 * (j.ty - 1 - uy))) * j.s + (j.phix + 2 * (j.tx - 1 - ux)));
}

__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const hn_pack_job *__restrict__ jobs, const int32_t *__restrict__ block_map)
{
    const hn_pack_job j = jobs[block_map[2 * blockIdx.x]];
    const int64_t total = (int64_t)j.rows_pad * j.kpad;
    const int64_t i0 = (int64_t)block_map[2 * blockIdx.x + 1] * PACK_CHUNK;
    for (int e = threadIdx.x; e < PACK_CHUNK; e += 256) {
        const int64_t i = i0 + e;
        if (i >= total) break;
        const int row = (int)(i / j.kpad), k = (int)(i - (int64_t)row * j.kpad);
        const float v = pack_source(j, row, k);
        if (j.dtype == HN_BF16) ((bf16_t *)j.dst)[i] = __float2bfloat16_rn(v);
        else ((float *)j.dst)[i] = v;
    }
}

// f64 -> f32 vector add into a parameter gradient: grad (+)= alpha * src
__global__ void vec_f64_to_grad_kernel(const double *__restrict__ src, float *__restrict__ grad, int n, int accumulate)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) grad[i] = (accumulate ? grad[i] : 0.f) + (float)src[i];
}

}  // namespace hn

using namespace hn;
using bf16 = __nv_bfloat16;

static bool same_shape(const hn_tensor *a, const hn_tensor *b) { return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c; }

extern "C" int hn_act_bwd(const hn_tensor *dout, const hn_tensor *out, int32_t act, float slope, const hn_tensor *dz, void *stream)
{
    HN_CHECK_ARG(dout && out && dz && dout->ptr && out->ptr && dz->ptr, "hn_act_bwd: null pointer");
    HN_CHECK_ARG(same_shape(dout, out) && same_shape(dout, dz) && dout->dtype == dz->dtype, "hn_act_bwd: shape/dtype mismatch");
    const int64_t npix = (int64_t)dout->n * dout->h * dout->w;
    if (npix == 0) return HN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = vec8_ok(dout) && vec8_ok(out) && vec8_ok(dz);
#define HN_ACT_BWD(TG, TO)                                                                                                               \
    do {                                                                                                                                 \
        if (vec)                                                                                                                         \
            act_bwd_kernel<TG, TO><<<wave_grid_b(cdiv(npix, 64) * 256, 256, 16), 256, 0, st>>>((const TG *)dout->ptr, dout->ld, (const TO *)out->ptr, out->ld, act, slope, (TG *)dz->ptr, dz->ld, npix, dout->c); \
        else                                                                                                                             \
            act_bwd_scalar_kernel<TG, TO><<<wave_grid_b(npix * dout->c, 256), 256, 0, st>>>((const TG *)dout->ptr, dout->ld, (const TO *)out->ptr, out->ld, act, slope, (TG *)dz->ptr, dz->ld, npix, dout->c); \
    } while (0)
    if (dout->dtype == HN_BF16 && out->dtype == HN_BF16) HN_ACT_BWD(bf16, bf16);
    else if (dout->dtype == HN_F32 && out->dtype == HN_F32) HN_ACT_BWD(float, float);
    else if (dout->dtype == HN_BF16 && out->dtype == HN_F32) HN_ACT_BWD(bf16, float);
    else HN_ACT_BWD(float, bf16);
#undef HN_ACT_BWD
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_bn_bwd(const hn_tensor *dout, const hn_tensor *out, const hn_tensor *raw, const float *mean, const float *invstd,
                         const float *gamma, int32_t act, float slope, const float *slope_ptr, double *sums /* [2*C + 1] scratch+result */,
                         const hn_tensor *draw, const hn_tensor *dres, int32_t dres_accumulate, int32_t want_prelu_grad, float *dbeta, float *dgamma,
                         float *dslope, int32_t param_accumulate, const float *fwd_scale, const float *fwd_shift, int32_t groups, void *stream)
{
    if (groups < 1) groups = 1;
    HN_CHECK_ARG(dout && out && raw && mean && invstd && sums && draw, "hn_bn_bwd: null pointer");
    HN_CHECK_ARG((fwd_scale != nullptr) == (fwd_shift != nullptr), "hn_bn_bwd: give both forward vectors or neither");
    const bool zout = fwd_scale == nullptr;      // read the saved output; else recompute z = raw*scale + shift
    HN_CHECK_ARG(same_shape(dout, out) && same_shape(dout, raw) && same_shape(dout, draw), "hn_bn_bwd: shape mismatch");
    HN_CHECK_ARG(raw->dtype == HN_F32 || (raw->dtype == HN_BF16 && dout->dtype == HN_BF16), "hn_bn_bwd: the pre-normalisation tensor is FP32 (or BF16 on the BF16 path)");
    HN_CHECK_ARG(dout->dtype == draw->dtype && (!dres || dres->dtype == dout->dtype), "hn_bn_bwd: gradient dtypes must match");
    HN_CHECK_ARG(vec8_ok(dout) && vec8_ok(out) && vec8_ok(raw) && vec8_ok(draw) && (!dres || vec8_ok(dres)), "hn_bn_bwd: views must be 8-channel aligned");
    const int C = dout->c;
    HN_CHECK_ARG(dout->n % groups == 0, "hn_bn_bwd: %d images do not split into %d statistics groups", dout->n, groups);
    HN_CHECK_ARG((int64_t)dout->n * dout->h * dout->w < ((int64_t)1 << 31), "hn_bn_bwd: too many pixels");
    const int64_t npix = (int64_t)(dout->n / groups) * dout->h * dout->w;      // per statistics group
    cudaStream_t st = (cudaStream_t)stream;
    HN_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (2 * C + 1) * groups, st));
    if (npix == 0) return HN_OK;
    const int ncv = C / 4;
    const int CVB = ncv < 32 ? ncv : 32;
    const int PL = 256 / CVB;
    const int cvblocks = (int)cdiv(ncv, CVB);
    int64_t chunks = cdiv((int64_t)num_sms() * 4, cvblocks);         // reduce pass: one wave of 4 CTAs per SM
    int64_t pix_per_cta = cdiv(npix, chunks);
    if (pix_per_cta < (int64_t)PL * 16) pix_per_cta = (int64_t)PL * 16;
    chunks = cdiv(npix, pix_per_cta);
    dim3 grid((unsigned)chunks, (unsigned)cvblocks, (unsigned)groups), block(CVB, PL);
    size_t smem = (size_t)2 * PL * CVB * 4 * sizeof(float);
    double *s1 = sums, *s2 = sums + C, *sp = want_prelu_grad ? sums + 2 * C : nullptr;
    // apply pass: no reduction, so finer pixel chunks (two waves)
    int64_t chunks2 = cdiv((int64_t)num_sms() * 8, cvblocks);
    int64_t pix_per_cta2 = cdiv(npix, chunks2);
    if (pix_per_cta2 < (int64_t)PL * 8) pix_per_cta2 = (int64_t)PL * 8;
    chunks2 = cdiv(npix, pix_per_cta2);
    dim3 grid2((unsigned)chunks2, (unsigned)cvblocks, (unsigned)groups);
    const double inv_count = 1.0 / (double)npix;
#define HN_BN_BWD_ZD(TG, TO, TR, Z, D)                                                                                                    \
        bn_bwd_apply_kernel<TG, TO, TR, Z, D><<<grid2, block, 0, st>>>((const TG *)dout->ptr, dout->ld, (const TO *)out->ptr, out->ld, (const TR *)raw->ptr, raw->ld, mean, invstd, gamma, s1, s2, inv_count, act, slope, slope_ptr, (TG *)draw->ptr, draw->ld, dres ? (TG *)dres->ptr : nullptr, dres ? dres->ld : 0, (int)npix, C, (int)pix_per_cta2, dbeta, dgamma, want_prelu_grad ? dslope : nullptr, sp, param_accumulate, fwd_scale, fwd_shift)
#define HN_BN_BWD_Z(TG, TO, TR, Z)                                                                                                       \
    do {                                                                                                                                  \
        bn_bwd_reduce_kernel<TG, TO, TR, Z><<<grid, block, smem, st>>>((const TG *)dout->ptr, dout->ld, (const TO *)out->ptr, out->ld, (const TR *)raw->ptr, raw->ld, mean, invstd, act, slope, slope_ptr, npix, C, pix_per_cta, s1, s2, sp, fwd_scale, fwd_shift); \
        if (!dres) HN_BN_BWD_ZD(TG, TO, TR, Z, 0);                                                                                        \
        else if (!dres_accumulate) HN_BN_BWD_ZD(TG, TO, TR, Z, 1);                                                                        \
        else HN_BN_BWD_ZD(TG, TO, TR, Z, 2);                                                                                              \
    } while (0)
#define HN_BN_BWD(TG, TO, TR)                                                                                                             \
    do {                                                                                                                                  \
        if (zout) HN_BN_BWD_Z(TG, TO, TR, true);                                                                                          \
        else HN_BN_BWD_Z(TG, TO, TR, false);                                                                                              \
    } while (0)
    if (dout->dtype == HN_BF16 && out->dtype == HN_BF16 && raw->dtype == HN_BF16) HN_BN_BWD(bf16, bf16, bf16);
    else if (dout->dtype == HN_BF16 && out->dtype == HN_BF16) HN_BN_BWD(bf16, bf16, float);
    else if (dout->dtype == HN_F32 && out->dtype == HN_F32) HN_BN_BWD(float, float, float);
    else {
        set_error("hn_bn_bwd: unsupported dtype combination");
        return HN_ERR_ARG;
    }
#undef HN_BN_BWD_ZD
#undef HN_BN_BWD
#undef HN_BN_BWD_Z
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_accumulate(const hn_tensor *x, const hn_tensor *y, int32_t accumulate, void *stream)
{
    HN_CHECK_ARG(x && y && x->ptr && y->ptr && same_shape(x, y), "hn_accumulate: bad arguments");
    const int64_t npix = (int64_t)x->n * x->h * x->w;
    if (npix == 0) return HN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = vec8_ok(x) && vec8_ok(y);
    int grid = vec ? wave_grid_b(cdiv(npix, 64) * 256, 256, 16) : wave_grid_b(npix * x->c, 256);
#define HN_ACC(TX, TY)                                                                                                       \
    do {                                                                                                                     \
        if (vec) accumulate_kernel<TX, TY><<<grid, 256, 0, st>>>((const TX *)x->ptr, x->ld, (TY *)y->ptr, y->ld, accumulate, npix, x->c); \
        else accumulate_scalar_kernel<TX, TY><<<grid, 256, 0, st>>>((const TX *)x->ptr, x->ld, (TY *)y->ptr, y->ld, accumulate, npix, x->c); \
    } while (0)
    if (x->dtype == HN_BF16 && y->dtype == HN_BF16) HN_ACC(bf16, bf16);
    else if (x->dtype == HN_F32 && y->dtype == HN_F32) HN_ACC(float, float);
    else if (x->dtype == HN_F32 && y->dtype == HN_BF16) HN_ACC(float, bf16);
    else HN_ACC(bf16, float);
#undef HN_ACC
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_maxpool3x3s2_fwd_idx(const hn_tensor *x, const hn_tensor *y, uint8_t *idx, void *stream)
{
    HN_CHECK_ARG(x && y && idx && x->ptr && y->ptr, "hn_maxpool3x3s2_fwd_idx: null pointer");
    HN_CHECK_ARG(x->dtype == y->dtype && x->c == y->c && x->n == y->n, "hn_maxpool3x3s2_fwd_idx: dtype/shape mismatch");
    HN_CHECK_ARG(y->h == (x->h - 1) / 2 + 1 && y->w == (x->w - 1) / 2 + 1, "hn_maxpool3x3s2_fwd_idx: bad output size");
    HN_CHECK_ARG(vec8_ok(x) && vec8_ok(y), "hn_maxpool3x3s2_fwd_idx: views must be 8-channel aligned");
    if ((int64_t)y->n * y->h * y->w == 0) return HN_OK;
    dim3 grid = row_grid_b((int64_t)y->n * y->h, (int64_t)y->w * (y->c / 8));
    cudaStream_t st = (cudaStream_t)stream;
    if (x->dtype == HN_BF16) maxpool_idx_kernel<bf16><<<grid, 256, 0, st>>>((const bf16 *)x->ptr, x->ld, (bf16 *)y->ptr, y->ld, idx, x->n, x->h, x->w, y->h, y->w, x->c);
    else maxpool_idx_kernel<float><<<grid, 256, 0, st>>>((const float *)x->ptr, x->ld, (float *)y->ptr, y->ld, idx, x->n, x->h, x->w, y->h, y->w, x->c);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_maxpool3x3s2_bwd(const hn_tensor *dy, const uint8_t *idx, const hn_tensor *dx, int32_t accumulate, void *stream)
{
    HN_CHECK_ARG(dy && dx && idx && dy->ptr && dx->ptr, "hn_maxpool3x3s2_bwd: null pointer");
    HN_CHECK_ARG(dy->dtype == dx->dtype && dy->c == dx->c && dy->n == dx->n, "hn_maxpool3x3s2_bwd: dtype/shape mismatch");
    HN_CHECK_ARG(dy->h == (dx->h - 1) / 2 + 1 && dy->w == (dx->w - 1) / 2 + 1, "hn_maxpool3x3s2_bwd: bad sizes");
    HN_CHECK_ARG(vec8_ok(dy) && vec8_ok(dx), "hn_maxpool3x3s2_bwd: views must be 8-channel aligned");
    if ((int64_t)dx->n * dx->h * dx->w == 0) return HN_OK;
    dim3 grid = row_grid_b((int64_t)dx->n * dx->h, (int64_t)dx->w * (dx->c / 8));
    cudaStream_t st = (cudaStream_t)stream;
    if (dy->dtype == HN_BF16) maxpool_bwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16 *)dy->ptr, dy->ld, idx, (bf16 *)dx->ptr, dx->ld, accumulate, dx->n, dx->h, dx->w, dy->h, dy->w, dx->c);
    else maxpool_bwd_kernel<float><<<grid, 256, 0, st>>>((const float *)dy->ptr, dy->ld, idx, (float *)dx->ptr, dx->ld, accumulate, dx->n, dx->h, dx->w, dy->h, dy->w, dx->c);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

// which algorithm the adjoint uses: 0 = generic gather, 1 = exact 2x micro-tiles, 2 = separable two-pass (needs workspace)
static int bilinear_bwd_mode(const hn_tensor *dy, const hn_tensor *dx)
{
    const bool vec = vec8_ok(dy) && vec8_ok(dx);
    if (vec && dy->h == 2 * dx->h && dy->w == 2 * dx->w) return 1;
    const int64_t mag = ((int64_t)dy->h * dy->w) / ((int64_t)dx->h * dx->w > 0 ? (int64_t)dx->h * dx->w : 1);
    if (mag >= 16 && (vec ? dx->w <= kSepMaxW : true)) return 2;
    return 0;
}

extern "C" int64_t hn_bilinear_bwd_workspace_bytes(const hn_tensor *dy, const hn_tensor *dx)
{
    if (!dy || !dx || dx->h <= 0 || dx->w <= 0) return 0;
    if (bilinear_bwd_mode(dy, dx) != 2) return 0;
    return (int64_t)sizeof(float) * dy->n * dy->h * dx->w * dx->c;
}

extern "C" int hn_bilinear_bwd(const hn_tensor *dy, const hn_tensor *dx, int32_t accumulate, void *workspace, int64_t workspace_bytes,
                               void *stream)
{
    HN_CHECK_ARG(dy && dx && dy->ptr && dx->ptr, "hn_bilinear_bwd: null pointer");
    HN_CHECK_ARG(dy->n == dx->n && dy->c == dx->c && dx->h > 0 && dx->w > 0, "hn_bilinear_bwd: shape mismatch");
    if ((int64_t)dx->n * dx->h * dx->w == 0) return HN_OK;
    const float sh = (float)dx->h / (float)dy->h, sw = (float)dx->w / (float)dy->w;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = vec8_ok(dy) && vec8_ok(dx);
    int mode = bilinear_bwd_mode(dy, dx);
    if (mode == 2 && (!workspace || workspace_bytes < hn_bilinear_bwd_workspace_bytes(dy, dx))) mode = 0;   // no scratch: gather form
    static const bool no_walk = getenv("HN_NO_UP2_BWD_WALK") != nullptr;
    if (mode == 1 && !no_walk && (int64_t)dx->n * dx->h <= 0x7fffffff && (int64_t)dx->n * dx->h * dx->w * dx->c >= (1 << 18)) {
        const int nseg = (int)cdiv(dx->w, UP2_SEG);
        dim3 grid((unsigned)(dx->n * dx->h), (unsigned)cdiv((int64_t)nseg * (dx->c / 8), 128));
#define HN_UP2_WALK(TG, TX) bilinear_up2_bwd_walk_kernel<TG, TX><<<grid, 128, 0, st>>>((const TG *)dy->ptr, dy->ld, (TX *)dx->ptr, dx->ld, accumulate, dx->n, dx->h, dx->w, dx->c)
        if (dy->dtype == HN_BF16 && dx->dtype == HN_BF16) HN_UP2_WALK(bf16, bf16);
        else if (dy->dtype == HN_F32 && dx->dtype == HN_F32) HN_UP2_WALK(float, float);
        else if (dy->dtype == HN_F32 && dx->dtype == HN_BF16) HN_UP2_WALK(float, bf16);
        else HN_UP2_WALK(bf16, float);
#undef HN_UP2_WALK
        HN_LAUNCH_CHECK();
        return HN_OK;
    }
    if (mode == 1) {
        dim3 grid = row_grid_b((int64_t)dx->n * ((dx->h + 1) / 2), (int64_t)((dx->w + 1) / 2) * (dx->c / 8));
#define HN_UP2_BWD(TG, TX) bilinear_up2_bwd_kernel<TG, TX><<<grid, 256, 0, st>>>((const TG *)dy->ptr, dy->ld, (TX *)dx->ptr, dx->ld, accumulate, dx->n, dx->h, dx->w, dx->c)
        if (dy->dtype == HN_BF16 && dx->dtype == HN_BF16) HN_UP2_BWD(bf16, bf16);
        else if (dy->dtype == HN_F32 && dx->dtype == HN_F32) HN_UP2_BWD(float, float);
        else if (dy->dtype == HN_F32 && dx->dtype == HN_BF16) HN_UP2_BWD(float, bf16);
        else HN_UP2_BWD(bf16, float);
#undef HN_UP2_BWD
        HN_LAUNCH_CHECK();
        return HN_OK;
    }
    if (mode == 2) {
        float *tmp = reinterpret_cast<float *>(workspace);
        const int64_t nrows = (int64_t)dy->n * dy->h;
        if (vec) {
            const int g1 = wave_grid_b(nrows * (dx->c / 8), 256, 8);
#define HN_ROWS_VEC(TG, MW) bilinear_bwd_rows_vec_kernel<TG, MW><<<g1, 256, 0, st>>>((const TG *)dy->ptr, dy->ld, tmp, nrows, dx->w, dy->w, dx->c, sw)
            if (dy->dtype == HN_BF16) {
                if (dx->w <= 2) HN_ROWS_VEC(bf16, 2);
                else if (dx->w <= 4) HN_ROWS_VEC(bf16, 4);
                else HN_ROWS_VEC(bf16, 8);
            } else {
                if (dx->w <= 2) HN_ROWS_VEC(float, 2);
                else if (dx->w <= 4) HN_ROWS_VEC(float, 4);
                else HN_ROWS_VEC(float, 8);
            }
#undef HN_ROWS_VEC
        } else {
            const int g1 = wave_grid_b(nrows * dx->w * dx->c, 256, 8);
            if (dy->dtype == HN_BF16) bilinear_bwd_rows_scalar_kernel<bf16><<<g1, 256, 0, st>>>((const bf16 *)dy->ptr, dy->ld, tmp, nrows, dx->w, dy->w, dx->c, sw);
            else bilinear_bwd_rows_scalar_kernel<float><<<g1, 256, 0, st>>>((const float *)dy->ptr, dy->ld, tmp, nrows, dx->w, dy->w, dx->c, sw);
        }
        HN_LAUNCH_CHECK();
        const int g2 = wave_grid_b((int64_t)dx->n * dx->h * dx->w * (vec ? dx->c / 8 : dx->c), 256, 8);
        if (vec) {
            if (dx->dtype == HN_BF16) bilinear_bwd_cols_kernel<bf16, true><<<g2, 256, 0, st>>>(tmp, (bf16 *)dx->ptr, dx->ld, accumulate, dx->n, dx->h, dx->w, dy->h, dx->c, sh);
            else bilinear_bwd_cols_kernel<float, true><<<g2, 256, 0, st>>>(tmp, (float *)dx->ptr, dx->ld, accumulate, dx->n, dx->h, dx->w, dy->h, dx->c, sh);
        } else {
            if (dx->dtype == HN_BF16) bilinear_bwd_cols_kernel<bf16, false><<<g2, 256, 0, st>>>(tmp, (bf16 *)dx->ptr, dx->ld, accumulate, dx->n, dx->h, dx->w, dy->h, dx->c, sh);
            else bilinear_bwd_cols_kernel<float, false><<<g2, 256, 0, st>>>(tmp, (float *)dx->ptr, dx->ld, accumulate, dx->n, dx->h, dx->w, dy->h, dx->c, sh);
        }
        HN_LAUNCH_CHECK();
        return HN_OK;
    }
    dim3 grid = row_grid_b((int64_t)dx->n * dx->h, (int64_t)dx->w * (vec ? dx->c / 8 : dx->c));
#define HN_BIL_BWD(TG, TX)                                                                                                            \
    do {                                                                                                                              \
        if (vec) bilinear_bwd_kernel<TG, TX, true><<<grid, 256, 0, st>>>((const TG *)dy->ptr, dy->ld, (TX *)dx->ptr, dx->ld, accumulate, dx->n, dx->h, dx->w, dy->h, dy->w, dx->c, sh, sw); \
        else bilinear_bwd_kernel<TG, TX, false><<<grid, 256, 0, st>>>((const TG *)dy->ptr, dy->ld, (TX *)dx->ptr, dx->ld, accumulate, dx->n, dx->h, dx->w, dy->h, dy->w, dx->c, sh, sw); \
    } while (0)
    if (dy->dtype == HN_BF16 && dx->dtype == HN_BF16) HN_BIL_BWD(bf16, bf16);
    else if (dy->dtype == HN_F32 && dx->dtype == HN_F32) HN_BIL_BWD(float, float);
    else if (dy->dtype == HN_F32 && dx->dtype == HN_BF16) HN_BIL_BWD(float, bf16);
    else HN_BIL_BWD(bf16, float);
#undef HN_BIL_BWD
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_pyramid_pool_bwd(const void *dpool, const int32_t *sizes, int32_t nsizes, const hn_tensor *dx, int32_t accumulate, void *stream)
{
    HN_CHECK_ARG(dpool && sizes && dx && dx->ptr, "hn_pyramid_pool_bwd: null pointer");
    HN_CHECK_ARG(nsizes >= 1 && nsizes <= 4 && vec8_ok(dx), "hn_pyramid_pool_bwd: bad arguments");
    if ((int64_t)dx->n * dx->h * dx->w == 0) return HN_OK;
    PoolSizesB ps{};
    ps.n = nsizes;
    for (int i = 0; i < nsizes; ++i) ps.s[i] = sizes[i];
    cudaStream_t st = (cudaStream_t)stream;
    const size_t tab_bytes = (size_t)4 * dx->w * 2 * sizeof(float) + (size_t)4 * dx->w * sizeof(int);
    if (tab_bytes <= 40 * 1024 && (int64_t)dx->n * dx->h <= 0x7fffffff) {
        const int ncv = dx->c / 8;
        const int threads = ncv >= 256 ? 256 : (ncv >= 128 ? 128 : (ncv >= 64 ? 64 : 32));
        dim3 g((unsigned)(dx->n * dx->h), (unsigned)cdiv(ncv, threads));
        if (dx->dtype == HN_BF16)
            pyramid_bwd_rows_kernel<bf16><<<g, threads, tab_bytes, st>>>((const bf16 *)dpool, (bf16 *)dx->ptr, dx->ld, accumulate, dx->n, dx->h, dx->w, dx->c, ps);
        else
            pyramid_bwd_rows_kernel<float><<<g, threads, tab_bytes, st>>>((const float *)dpool, (float *)dx->ptr, dx->ld, accumulate, dx->n, dx->h, dx->w, dx->c, ps);
        HN_LAUNCH_CHECK();
        return HN_OK;
    }
    dim3 grid = row_grid_b((int64_t)dx->n * dx->h, (int64_t)dx->w * (dx->c / 8));
    if (dx->dtype == HN_BF16) pyramid_bwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16 *)dpool, (bf16 *)dx->ptr, dx->ld, accumulate, dx->n, dx->h, dx->w, dx->c, ps);
    else pyramid_bwd_kernel<float><<<grid, 256, 0, st>>>((const float *)dpool, (float *)dx->ptr, dx->ld, accumulate, dx->n, dx->h, dx->w, dx->c, ps);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_dilate(const hn_tensor *x, int32_t stride, const hn_tensor *up, void *stream)
{
    HN_CHECK_ARG(x && up && x->ptr && up->ptr && stride >= 1, "hn_dilate: bad arguments");
    HN_CHECK_ARG(x->dtype == up->dtype && x->n == up->n && x->c == up->c, "hn_dilate: dtype/shape mismatch");
    HN_CHECK_ARG(up->h >= (x->h - 1) * stride + 1 && up->w >= (x->w - 1) * stride + 1, "hn_dilate: output too small");
    if ((int64_t)up->n * up->h * up->w == 0) return HN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (!(vec8_ok(x) && vec8_ok(up))) {
        int g = wave_grid_b((int64_t)up->n * up->h * up->w * up->c, 256);
        if (x->dtype == HN_BF16) dilate_scalar_kernel<bf16><<<g, 256, 0, st>>>((const bf16 *)x->ptr, x->ld, (bf16 *)up->ptr, up->ld, x->n, x->h, x->w, up->h, up->w, x->c, stride);
        else dilate_scalar_kernel<float><<<g, 256, 0, st>>>((const float *)x->ptr, x->ld, (float *)up->ptr, up->ld, x->n, x->h, x->w, up->h, up->w, x->c, stride);
        HN_LAUNCH_CHECK();
        return HN_OK;
    }
    dim3 grid = row_grid_b((int64_t)up->n * up->h, (int64_t)up->w * (up->c / 8));
    if (x->dtype == HN_BF16) dilate_kernel<bf16><<<grid, 256, 0, st>>>((const bf16 *)x->ptr, x->ld, (bf16 *)up->ptr, up->ld, x->n, x->h, x->w, up->h, up->w, x->c, stride);
    else dilate_kernel<float><<<grid, 256, 0, st>>>((const float *)x->ptr, x->ld, (float *)up->ptr, up->ld, x->n, x->h, x->w, up->h, up->w, x->c, stride);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_unpack_wgrad(const float *packed, float *grad_oihw, int32_t cout, int32_t cin, int32_t r, int32_t s, int32_t kpad,
                               int32_t accumulate, void *stream)
{
    HN_CHECK_ARG(packed && grad_oihw && kpad >= cin * r * s, "hn_unpack_wgrad: bad arguments");
    int64_t total = (int64_t)cout * cin * r * s;
    unpack_wgrad_kernel<<<wave_grid_b(total, 256), 256, 0, (cudaStream_t)stream>>>(packed, grad_oihw, cout, cin, r, s, kpad, accumulate);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_pack_weight_dgrad(const float *w_oihw, void *dst, int32_t dtype, int32_t cout, int32_t cin, int32_t r, int32_t s,
                                    int32_t cin_pad, int32_t kpad, void *stream)
{
    HN_CHECK_ARG(w_oihw && dst && cin_pad >= cin && kpad >= r * s * cout, "hn_pack_weight_dgrad: bad arguments");
    int64_t total = (int64_t)cin_pad * kpad;
    int grid = wave_grid_b(total, 256);
    if (dtype == HN_BF16) pack_weight_dgrad_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(w_oihw, (bf16 *)dst, cout, cin, r, s, cin_pad, kpad);
    else pack_weight_dgrad_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(w_oihw, (float *)dst, cout, cin, r, s, cin_pad, kpad);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int32_t hn_dgrad_s2_phase_kpad(int32_t cout, int32_t r, int32_t s, int32_t pad, int32_t phase)
{
    const PhaseAxis ay = phase_axis(phase >> 1, r, pad, 2), ax = phase_axis(phase & 1, s, pad, 2);
    return ay.T * ax.T == 0 ? 0 : hn_conv_kpad(cout, ay.T, ax.T);
}

extern "C" int hn_pack_weight_dgrad_phase(const float *w_oihw, void *dst, int32_t dtype, int32_t cout, int32_t cin, int32_t r, int32_t s,
                                          int32_t pad, int32_t phase, int32_t cin_pad, void *stream)
{
    HN_CHECK_ARG(w_oihw && dst && cin_pad >= cin && phase >= 0 && phase < 4, "hn_pack_weight_dgrad_phase: bad arguments");
    const PhaseAxis ay = phase_axis(phase >> 1, r, pad, 2), ax = phase_axis(phase & 1, s, pad, 2);
    HN_CHECK_ARG(ay.T > 0 && ax.T > 0, "hn_pack_weight_dgrad_phase: phase %d has no taps", phase);
    const int kpad = hn_conv_kpad(cout, ay.T, ax.T);
    int64_t total = (int64_t)cin_pad * kpad;
    int grid = wave_grid_b(total, 256);
    if (dtype == HN_BF16)
        pack_weight_dgrad_phase_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(w_oihw, (bf16 *)dst, cout, cin, r, s, ay.T, ax.T, ay.phi, ax.phi, cin_pad, kpad);
    else
        pack_weight_dgrad_phase_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(w_oihw, (float *)dst, cout, cin, r, s, ay.T, ax.T, ay.phi, ax.phi, cin_pad, kpad);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_conv2d_dgrad_s2_ok(const hn_tensor *dy, const hn_conv *cv, const hn_tensor *dx)
{
    if (!dy || !cv || !dx) return 0;
    return dy->dtype == HN_BF16 && dx->dtype == HN_BF16 && cv->stride == 2 && cv->dil == 1 && dy->c % 64 == 0 && dy->ld % 8 == 0 &&
           (reinterpret_cast<uintptr_t>(dy->ptr) & 15) == 0 && dx->ld % 8 == 0 && (reinterpret_cast<uintptr_t>(dx->ptr) & 15) == 0 && cv->pad <= cv->r - 1 &&
           cv->pad <= cv->s - 1;
}

extern "C" int hn_conv2d_dgrad_s2(const hn_tensor *dy, const void *const *w_phase, const hn_conv *cv, const hn_tensor *dx, int32_t accumulate,
                                  void *stream)
{
    HN_CHECK_ARG(dy && w_phase && cv && dx && dy->ptr && dx->ptr, "hn_conv2d_dgrad_s2: null pointer");
    HN_CHECK_ARG(hn_conv2d_dgrad_s2_ok(dy, cv, dx), "hn_conv2d_dgrad_s2: needs BF16, stride 2, dilation 1, dY channels %% 64 == 0, aligned views");
    HN_CHECK_ARG(dy->n == dx->n && dy->c == cv->cout, "hn_conv2d_dgrad_s2: shape mismatch");
    HN_CHECK_ARG(dy->h == (dx->h + 2 * cv->pad - cv->r) / 2 + 1 && dy->w == (dx->w + 2 * cv->pad - cv->s) / 2 + 1,
                 "hn_conv2d_dgrad_s2: dY must be %dx%d for a %dx%d input", (dx->h + 2 * cv->pad - cv->r) / 2 + 1, (dx->w + 2 * cv->pad - cv->s) / 2 + 1,
                 dx->h, dx->w);
    cudaStream_t st = (cudaStream_t)stream;
    for (int phase = 0; phase < 4; ++phase) {
        const int ry = phase >> 1, rx = phase & 1;
        const PhaseAxis ay = phase_axis(ry, cv->r, cv->pad, dx->h), ax = phase_axis(rx, cv->s, cv->pad, dx->w);
        if (ay.count == 0 || ax.count == 0) continue;
        bf16 *base = (bf16 *)dx->ptr + ((int64_t)ry * dx->w + rx) * dx->ld;
        const int64_t pix = 2 * (int64_t)dx->ld, row = 2 * (int64_t)dx->w * dx->ld, img = (int64_t)dx->h * dx->w * dx->ld;
        if (ay.T == 0 || ax.T == 0) {
            if (!accumulate) {
                const int64_t total = (int64_t)dx->n * ay.count * ax.count * dx->c;
                zero_lattice_kernel<bf16><<<wave_grid_b(total, 256), 256, 0, st>>>(base, pix, row, img, dx->n, ay.count, ax.count, dx->c);
                HN_LAUNCH_CHECK();
            }
            continue;
        }
        HN_CHECK_ARG(w_phase[phase] != nullptr, "hn_conv2d_dgrad_s2: missing sub-filter pack of phase %d", phase);
        TcSubConv sc{ay.T, ax.T, ay.padp, ax.padp, ay.count, ax.count, base, pix, row, img, accumulate};
        int rc = conv2d_fwd_tc_sub(dy, w_phase[phase], dx->c, &sc, st);
        if (rc) return rc;
    }
    return HN_OK;
}

extern "C" int32_t hn_pack_chunk(void) { return PACK_CHUNK; }

extern "C" int hn_pack_job_init(hn_pack_job *job, const float *w_oihw, void *dst, int32_t dtype, int32_t kind, int32_t cout, int32_t cin, int32_t r,
                                int32_t s, int32_t pad, int32_t phase)
{
    HN_CHECK_ARG(job && w_oihw && dst && kind >= 0 && kind <= 2, "hn_pack_job_init: bad arguments");
    memset(job, 0, sizeof(*job));
    job->src = w_oihw; job->dst = dst; job->dtype = dtype; job->kind = kind;
    job->cout = cout; job->cin = cin; job->r = r; job->s = s;
    if (kind == 0) {
        job->rows_pad = hn_conv_cout_pad(cout, dtype);
        job->kpad = hn_conv_kpad(cin, r, s);
    } else if (kind == 1) {
        job->rows_pad = hn_conv_cout_pad(cin, dtype);
        job->kpad = hn_conv_kpad(cout, r, s);
    } else {
        const PhaseAxis ay = phase_axis(phase >> 1, r, pad, 2), ax = phase_axis(phase & 1, s, pad, 2);
        HN_CHECK_ARG(ay.T > 0 && ax.T > 0, "hn_pack_job_init: phase %d has no taps", phase);
        job->ty = ay.T; job->tx = ax.T; job->phiy = ay.phi; job->phix = ax.phi;
        job->rows_pad = hn_conv_cout_pad(cin, dtype);
        job->kpad = hn_conv_kpad(cout, ay.T, ax.T);
    }
    return HN_OK;
}

extern "C" int hn_pack_weights_multi(const hn_pack_job *jobs_dev, const int32_t *block_map_dev, int32_t n_blocks, void *stream)
{
    HN_CHECK_ARG(jobs_dev && block_map_dev && n_blocks >= 0, "hn_pack_weights_multi: bad arguments");
    if (n_blocks == 0) return HN_OK;
    pack_weights_multi_kernel<<<n_blocks, 256, 0, (cudaStream_t)stream>>>(jobs_dev, block_map_dev);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_vec_to_grad(const double *src, float *grad, int32_t n, int32_t accumulate, void *stream)
{
    HN_CHECK_ARG(src && grad && n > 0, "hn_vec_to_grad: bad arguments");
    vec_f64_to_grad_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(src, grad, n, accumulate);
    HN_LAUNCH_CHECK();
    return HN_OK;
}
