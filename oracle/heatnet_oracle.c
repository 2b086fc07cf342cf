/* Plain-C oracle for the primitives of the HeatNet hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates, in scalar C with double accumulation, the numerical semantics the reference obtains from
 * PyTorch / numpy calls.  Nothing in the product path links or loads this file;
 * tests/ use it through ctypes to pin oracle/heatnet_oracle.py and to check the CUDA kernels.
 * Layout: NCHW FP32, as in the reference.  Citations are relative to /root/reference
 * ("cm/" = models/confusion_maximization/).
 *
 * Build: gcc -O2 -shared -fPIC oracle/heatnet_oracle.c -o oracle/libheatnet_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX4(n, c, h, w, C, H, W) ((((size_t)(n) * (C) + (c)) * (H) + (h)) * (W) + (w))

/* nn.Conv2d (zero padding, square kernel, groups = 1): cm/models/extractors.py:71-76,111-123;
 * cm/models/pspnet.py:13,18,32,57; cm/discriminator_model.py:40-44.  w is OIHW, bias may be NULL. */
void hno_conv2d(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W,
                int K, int R, int S, int stride, int pad, int dil)
{
    int Ho = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1;
    int Wo = (W + 2 * pad - dil * (S - 1) - 1) / stride + 1;

    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k)
            for (int ho = 0; ho < Ho; ++ho)
                for (int wo = 0; wo < Wo; ++wo) {
                    double acc = bias ? (double)bias[k] : 0.0;
                    for (int c = 0; c < C; ++c)
                        for (int r = 0; r < R; ++r) {
                            int hi = ho * stride - pad + r * dil;
                            if (hi < 0 || hi >= H) continue;
                            for (int s = 0; s < S; ++s) {
                                int wi = wo * stride - pad + s * dil;
                                if (wi < 0 || wi >= W) continue;
                                acc += (double)x[IDX4(n, c, hi, wi, C, H, W)] *
                                       (double)w[IDX4(k, c, r, s, C, R, S)];
                            }
                        }
                    y[IDX4(n, k, ho, wo, K, Ho, Wo)] = (float)acc;
                }
}

/* nn.BatchNorm2d.  training != 0: normalise with the biased batch variance, update running_mean /
 * running_var (unbiased) with `momentum`; else use the running statistics.  (appendix B.3) */
void hno_batchnorm(const float *x, float *y, const float *gamma, const float *beta, float *running_mean,
                   float *running_var, int N, int C, int H, int W, int training, float momentum, float eps)
{
    size_t hw = (size_t)H * W;
    double cnt = (double)N * hw;
    for (int c = 0; c < C; ++c) {
        double mean, var;
        if (training) {
            double s = 0.0, ss = 0.0;
            for (int n = 0; n < N; ++n)
                for (size_t i = 0; i < hw; ++i) s += x[((size_t)n * C + c) * hw + i];
            mean = s / cnt;
            for (int n = 0; n < N; ++n)
                for (size_t i = 0; i < hw; ++i) {
                    double d = x[((size_t)n * C + c) * hw + i] - mean;
                    ss += d * d;
                }
            var = ss / cnt;
            running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
            running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * (ss / (cnt - 1.0)));
        } else {
            mean = running_mean[c];
            var = running_var[c];
        }
        double inv = 1.0 / sqrt(var + (double)eps);
        for (int n = 0; n < N; ++n)
            for (size_t i = 0; i < hw; ++i) {
                size_t j = ((size_t)n * C + c) * hw + i;
                y[j] = (float)(((double)x[j] - mean) * inv * gamma[c] + beta[c]);
            }
    }
}

/* nn.MaxPool2d(3, 2, 1): pads with -inf, ceil_mode False (cm/models/extractors.py:128; appendix B.5). */
void hno_maxpool3x3s2(const float *x, float *y, int N, int C, int H, int W)
{
    int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    for (int nc = 0; nc < N * C; ++nc)
        for (int ho = 0; ho < Ho; ++ho)
            for (int wo = 0; wo < Wo; ++wo) {
                float m = -INFINITY;
                for (int r = 0; r < 3; ++r)
                    for (int s = 0; s < 3; ++s) {
                        int hi = 2 * ho - 1 + r, wi = 2 * wo - 1 + s;
                        if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
                        float v = x[((size_t)nc * H + hi) * W + wi];
                        if (v > m) m = v;
                    }
                y[((size_t)nc * Ho + ho) * Wo + wo] = m;
            }
}

/* nn.AdaptiveAvgPool2d((s, s)): bin i covers [floor(i*H/s), ceil((i+1)*H/s)) (cm/models/pspnet.py:17; B.2). */
void hno_adaptive_avgpool(const float *x, float *y, int N, int C, int H, int W, int s)
{
    for (int nc = 0; nc < N * C; ++nc)
        for (int i = 0; i < s; ++i) {
            int h0 = (i * H) / s, h1 = ((i + 1) * H + s - 1) / s;
            for (int j = 0; j < s; ++j) {
                int w0 = (j * W) / s, w1 = ((j + 1) * W + s - 1) / s;
                double acc = 0.0;
                for (int h = h0; h < h1; ++h)
                    for (int w = w0; w < w1; ++w) acc += x[((size_t)nc * H + h) * W + w];
                y[((size_t)nc * s + i) * s + j] = (float)(acc / ((double)(h1 - h0) * (w1 - w0)));
            }
        }
}

/* F.upsample(mode='bilinear') == F.interpolate(align_corners=False) (cm/models/pspnet.py:23,39;
 * cm/discriminator_model.py:47): src = (dst + 0.5) * in/out - 0.5 clamped at 0; neighbour clamped at in-1. */
void hno_bilinear(const float *x, float *y, int N, int C, int H, int W, int Ho, int Wo)
{
    double sh = (double)H / Ho, sw = (double)W / Wo;
    for (int nc = 0; nc < N * C; ++nc)
        for (int ho = 0; ho < Ho; ++ho) {
            double fy = ((double)ho + 0.5) * sh - 0.5;
            if (fy < 0) fy = 0;
            int y0 = (int)fy, y1 = y0 + (y0 < H - 1 ? 1 : 0);
            double ly = fy - y0;
            for (int wo = 0; wo < Wo; ++wo) {
                double fx = ((double)wo + 0.5) * sw - 0.5;
                if (fx < 0) fx = 0;
                int x0 = (int)fx, x1 = x0 + (x0 < W - 1 ? 1 : 0);
                double lx = fx - x0;
                const float *p = x + (size_t)nc * H * W;
                double v = (1 - ly) * ((1 - lx) * p[(size_t)y0 * W + x0] + lx * p[(size_t)y0 * W + x1]) +
                           ly * ((1 - lx) * p[(size_t)y1 * W + x0] + lx * p[(size_t)y1 * W + x1]);
                y[((size_t)nc * Ho + ho) * Wo + wo] = (float)v;
            }
        }
}

/* y = x >= 0 ? x : slope * x : ReLU (slope 0), PReLU (one shared slope, cm/models/pspnet.py:34),
 * LeakyReLU(0.2) (cm/discriminator_model.py:46). */
void hno_leaky(const float *x, float *y, size_t n, float slope)
{
    for (size_t i = 0; i < n; ++i) y[i] = x[i] >= 0 ? x[i] : slope * x[i];
}

/* First-max argmax over the class dimension of (N, K, H, W) scores (scripts/iou_eval.py:154-157). */
void hno_argmax(const float *scores, int64_t *out, int N, int K, int H, int W)
{
    size_t hw = (size_t)H * W;
    for (int n = 0; n < N; ++n)
        for (size_t i = 0; i < hw; ++i) {
            int best = 0;
            float bv = scores[((size_t)n * K) * hw + i];
            for (int k = 1; k < K; ++k) {
                float v = scores[((size_t)n * K + k) * hw + i];
                if (v > bv) { bv = v; best = k; }
            }
            out[(size_t)n * hw + i] = best;
        }
}

/* ConfusionMatrix.add (scripts/iou_eval.py:82-88): conf[t*K + p] += 1 for every pixel, int32 accumulator
 * with wrap-around.  Returns -1 when a value is outside [0, K) (the reference's range asserts :58-79). */
int hno_confusion(const int64_t *pred, const int64_t *target, size_t n, int K, int32_t *conf)
{
    for (size_t i = 0; i < n; ++i)
        if (pred[i] < 0 || pred[i] >= K || target[i] < 0 || target[i] >= K) return -1;
    for (size_t i = 0; i < n; ++i) {
        uint32_t *cell = (uint32_t *)&conf[target[i] * K + pred[i]];
        *cell += 1u;
    }
    return 0;
}

/* CrossEntropyLoss (mean over non-ignored pixels; cm/train_trgb_segnet_conf.py:244, scripts/main.py:223). */
double hno_cross_entropy(const float *logits, const int64_t *labels, int N, int K, int H, int W, int ignore_index)
{
    size_t hw = (size_t)H * W, cnt = 0;
    double total = 0.0;
    for (int n = 0; n < N; ++n)
        for (size_t i = 0; i < hw; ++i) {
            int64_t t = labels[(size_t)n * hw + i];
            if (t == ignore_index) continue;
            double m = -INFINITY;
            for (int k = 0; k < K; ++k) {
                double v = logits[((size_t)n * K + k) * hw + i];
                if (v > m) m = v;
            }
            double se = 0.0;
            for (int k = 0; k < K; ++k) se += exp((double)logits[((size_t)n * K + k) * hw + i] - m);
            total += m + log(se) - (double)logits[((size_t)n * K + t) * hw + i];
            ++cnt;
        }
    return cnt ? total / (double)cnt : NAN;
}
