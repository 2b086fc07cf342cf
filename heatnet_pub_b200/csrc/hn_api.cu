// C-ABI glue of libheatnet_b200.so: error reporting, device check, convolution dispatch.
#include <stdarg.h>
#include <string.h>

#include "hn_common.cuh"

namespace hn {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace hn

using namespace hn;

extern "C" const char *hn_last_error(void) { return g_err; }
extern "C" int hn_version(void) { return 100; }

extern "C" int hn_device_check(void)
{
    int dev = 0, major = 0, minor = 0;
    HN_CUDA(cudaGetDevice(&dev));
    HN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    HN_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (major != 10) {
        set_error("libheatnet_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
        return HN_ERR_CUDA;
    }
    return num_sms();
}

extern "C" int32_t hn_conv_cout_pad(int32_t cout, int32_t dtype)
{
    if (dtype == HN_F32) return (int32_t)(cdiv(cout, 64) * 64);
    if (cout <= 16) return 16;
    if (cout <= 32) return 32;
    return (int32_t)(cdiv(cout, 64) * 64);
}

extern "C" int32_t hn_conv_kpad(int32_t cin, int32_t r, int32_t s) { return (int32_t)(cdiv((int64_t)cin * r * s, 64) * 64); }

static int conv_out_dim(int in, int k, int stride, int pad, int dil) { return (in + 2 * pad - dil * (k - 1) - 1) / stride + 1; }

extern "C" int64_t hn_conv2d_workspace_bytes(const hn_tensor *x, const hn_conv *cv)
{
    if (!x || !cv || x->dtype != HN_BF16) return 0;
    return conv2d_tc_workspace(x, cv);
}

extern "C" int hn_conv2d_fwd(const hn_tensor *x, const void *w_packed, const hn_conv *cv, const hn_epilogue *ep, const hn_tensor *y,
                             void *workspace, int64_t workspace_bytes, void *stream)
{
    HN_CHECK_ARG(x && w_packed && cv && ep && y && x->ptr && y->ptr, "hn_conv2d_fwd: null pointer");
    HN_CHECK_ARG(cv->r >= 1 && cv->s >= 1 && cv->stride >= 1 && cv->dil >= 1 && cv->pad >= 0 && cv->cout >= 1,
                 "hn_conv2d_fwd: bad conv geometry");
    HN_CHECK_ARG(x->c >= 1 && x->ld >= x->c && y->ld >= cv->cout, "hn_conv2d_fwd: bad channel strides");
    const int Ho = conv_out_dim(x->h, cv->r, cv->stride, cv->pad, cv->dil), Wo = conv_out_dim(x->w, cv->s, cv->stride, cv->pad, cv->dil);
    HN_CHECK_ARG(Ho >= 1 && Wo >= 1, "hn_conv2d_fwd: kernel size can't be greater than actual input size (%dx%d, k=%d)", x->h, x->w, cv->r);
    HN_CHECK_ARG(y->n == x->n && y->h == Ho && y->w == Wo && y->c == cv->cout, "hn_conv2d_fwd: output view must be N=%d %dx%d C=%d", x->n,
                 Ho, Wo, cv->cout);
    // y = act(acc * scale + shift + residual): either vector may be absent (scale: 1, shift: 0)
    HN_CHECK_ARG(!ep->out_nchw, "hn_conv2d_fwd: out_nchw is not implemented (see hn_conv3x3_head_fwd for NCHW logits)");
    HN_CHECK_ARG((ep->stat_sum != nullptr) == (ep->stat_sqsum != nullptr), "hn_conv2d_fwd: give both statistics accumulators or neither");
    if (ep->stat_sum) {
        const int cp = hn_conv_cout_pad(cv->cout, x->dtype);
        const size_t esz = elsize(y->dtype);
        HN_CHECK_ARG(x->dtype == HN_BF16 && cp >= 32 && (reinterpret_cast<uintptr_t>(y->ptr) & 15) == 0 && (y->ld * esz) % 16 == 0 &&
                         !ep->residual && ep->act == HN_ACT_NONE && (y->dtype == HN_F32 || !ep->scale),
                     "hn_conv2d_fwd: fused statistics need the BF16 engine, a 16-byte aligned output view, Cout >= 17, no residual / activation "
                     "(and no explicit scale with a BF16 output)");
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (x->dtype == HN_F32) {
        HN_CHECK_ARG(y->dtype == HN_F32, "hn_conv2d_fwd: FP32 path writes FP32");
        return conv2d_fwd_f32(x, w_packed, cv, ep, y, st);
    }
    return conv2d_fwd_tc(x, w_packed, cv, ep, y, workspace, workspace_bytes, st);
}

// ---- role-level cycle accounting (only meaningful in the -DHN_PROFILE_ROLES build; zeros otherwise)
namespace hn {
#ifdef HN_PROFILE_ROLES
__device__ unsigned long long g_role_cycles[16];
#endif
}
extern "C" int hn_prof_read(unsigned long long *out16, int reset)
{
#ifdef HN_PROFILE_ROLES
    HN_CUDA(cudaMemcpyFromSymbol(out16, hn::g_role_cycles, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {0};
        HN_CUDA(cudaMemcpyToSymbol(hn::g_role_cycles, z, sizeof(z)));
    }
    return 1;
#else
    for (int i = 0; i < 16; ++i) out16[i] = 0;
    (void)reset;
    return 0;
#endif
}
