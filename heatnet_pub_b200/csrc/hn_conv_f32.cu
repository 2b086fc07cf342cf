// FP32 CUDA-core implicit-GEMM convolution: the parity path that meets the north star's 1e-4 relative
// bound on logits (tcgen05 kind::tf32 operands cannot -- SURVEY.md appendix D).  Generic in kernel size,
// stride, dilation and padding; NHWC FP32 activations, weights packed [cout_pad][kpad] FP32 with
// k = (r*S+s)*Cin + c.  64x64 output tile per CTA, 4x4 per thread, K chunks of 16 through shared memory.
// Same fused epilogue as the tensor-core engine.
#include "hn_common.cuh"

namespace hn {

constexpr int F_BM = 64, F_BN = 64, F_BK = 16;

struct ConvGeom {
    int N, H, W, C, ldx;
    int Ho, Wo, Cout, ldy;
    int R, S, stride, pad, dil;
    int K, kpad;
};

__global__ void __launch_bounds__(256) conv_f32_kernel(const float *__restrict__ x, const float *__restrict__ w, float *__restrict__ y,
                                                       ConvGeom g, const float *__restrict__ scale, const float *__restrict__ shift,
                                                       const float *__restrict__ res, int ldr, int act, float slope,
                                                       const float *slope_ptr)
{
    __shared__ float As[F_BK][F_BM + 4];
    __shared__ float Bs[F_BK][F_BN + 4];
    const int tid = threadIdx.x;
    const int64_t M = (int64_t)g.N * g.Ho * g.Wo;
    const int64_t m0 = (int64_t)blockIdx.x * F_BM;
    const int n0 = blockIdx.y * F_BN;

    // loader role: row = tid/4 (pixel for A, cout for B), kq = (tid%4)*4
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const int64_t am = m0 + lrow;
    const bool a_valid = am < M;
    int an = 0, aho = 0, awo = 0;
    if (a_valid) {
        awo = (int)(am % g.Wo);
        aho = (int)((am / g.Wo) % g.Ho);
        an = (int)(am / ((int64_t)g.Wo * g.Ho));
    }
    const float *wrow = w + (int64_t)(n0 + lrow) * g.kpad;  // cout_pad rows exist, zero padded

    const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads, 4x4 outputs each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < g.K; k0 += F_BK) {
        // ---- A: gather 4 consecutive k of one pixel ----
        float av[4] = {0.f, 0.f, 0.f, 0.f};
        if (a_valid) {
            int k = k0 + lk;
            if ((g.C & 3) == 0) {
                if (k < g.K) {   // 4 consecutive k share one tap when Cin % 4 == 0
                    int tap = k / g.C, c = k - tap * g.C;
                    int r = tap / g.S, s = tap - r * g.S;
                    int hi = aho * g.stride - g.pad + r * g.dil, wi = awo * g.stride - g.pad + s * g.dil;
                    if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.W) {
                        const float *p = x + (((int64_t)an * g.H + hi) * g.W + wi) * g.ldx + c;
                        if ((g.ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
                            float4 v = *reinterpret_cast<const float4 *>(p);
                            av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
                        } else {
                            av[0] = p[0]; av[1] = p[1]; av[2] = p[2]; av[3] = p[3];
                        }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int kk = k + j;
                    if (kk < g.K) {
                        int tap = kk / g.C, c = kk - tap * g.C;
                        int r = tap / g.S, s = tap - r * g.S;
                        int hi = aho * g.stride - g.pad + r * g.dil, wi = awo * g.stride - g.pad + s * g.dil;
                        if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.W)
                            av[j] = x[(((int64_t)an * g.H + hi) * g.W + wi) * g.ldx + c];
                    }
                }
            }
        }
        float4 bv = *reinterpret_cast<const float4 *>(wrow + k0 + lk);  // kpad % 64 == 0: always in bounds
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) As[lk + j][lrow] = av[j];
        Bs[lk + 0][lrow] = bv.x; Bs[lk + 1][lrow] = bv.y; Bs[lk + 2][lrow] = bv.z; Bs[lk + 3][lrow] = bv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < F_BK; ++kk) {
            float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }
    if (slope_ptr) slope = __ldg(slope_ptr);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = n0 + tx * 4 + j;
            if (co >= g.Cout) continue;
            float v = acc[i][j];
            if (scale) v *= __ldg(scale + co);
            if (shift) v += __ldg(shift + co);
            if (res) v += res[m * ldr + co];
            y[m * g.ldy + co] = apply_act(v, act, slope);
        }
    }
}

int conv2d_fwd_f32(const hn_tensor *x, const void *w, const hn_conv *cv, const hn_epilogue *ep, const hn_tensor *y, cudaStream_t st)
{
    ConvGeom g;
    g.N = x->n; g.H = x->h; g.W = x->w; g.C = x->c; g.ldx = x->ld;
    g.Ho = y->h; g.Wo = y->w; g.Cout = cv->cout; g.ldy = y->ld;
    g.R = cv->r; g.S = cv->s; g.stride = cv->stride; g.pad = cv->pad; g.dil = cv->dil;
    g.K = cv->r * cv->s * x->c;
    g.kpad = hn_conv_kpad(x->c, cv->r, cv->s);
    const int64_t M = (int64_t)g.N * g.Ho * g.Wo;
    if (M == 0) return HN_OK;
    const int cout_pad = hn_conv_cout_pad(cv->cout, HN_F32);   // multiple of F_BN: every weight row the loader reads exists
    dim3 grid((unsigned)cdiv(M, F_BM), (unsigned)(cout_pad / F_BN));
    conv_f32_kernel<<<grid, 256, 0, st>>>((const float *)x->ptr, (const float *)w, (float *)y->ptr, g, ep->scale, ep->shift,
                                         (const float *)ep->residual, ep->residual_ld, ep->act, ep->slope, ep->slope_ptr);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

}  // namespace hn
