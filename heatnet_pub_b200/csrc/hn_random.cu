// Keyed counter-based random numbers for nn.Dropout2d (cm/models/pspnet.py:49,55,64-73).
// The reference draws its channel masks from torch's CUDA generator; no other kernel can reproduce that stream, so parity runs
// inject masks.  What the product needs instead is (a) reproducibility from a seed and (b) a NEW mask on every replay of a
// captured CUDA graph.  Both follow from keeping the generator state -- {seed, calls so far} -- in device memory: the kernel
// reads the call counter, draws Philox4x32-10(key = seed, counter = (call, element)) and bumps the counter itself, so a graph
// that replays the same launch still advances the stream.
#include "hn_common.cuh"

namespace hn {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// Philox4x32-10 (Salmon et al., SC'11): word 0 of the output block
__device__ __forceinline__ uint32_t philox4x32_10_w0(uint64_t seed, uint64_t call, uint64_t elem)
{
    uint32_t c[4] = {(uint32_t)elem, (uint32_t)(elem >> 32), (uint32_t)call, (uint32_t)(call >> 32)};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c[0];
}

// one block: every thread reads the call counter before thread 0 advances it
__global__ void __launch_bounds__(1024) dropout2d_scale_kernel(unsigned long long *__restrict__ state, int64_t n, float p, float *__restrict__ out)
{
    const unsigned long long seed = state[0], call = state[1];
    __syncthreads();
    const float keep_scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float u = (float)(philox4x32_10_w0(seed, call, (uint64_t)i) >> 8) * (1.0f / 16777216.0f);      // [0, 1), 24 bits
        out[i] = u >= p ? keep_scale : 0.f;
    }
    if (threadIdx.x == 0) state[1] = call + 1;
}

}  // namespace hn

using namespace hn;

extern "C" int hn_dropout2d_scale(uint64_t *state_dev, int64_t n, float p, float *scale_out, void *stream)
{
    HN_CHECK_ARG(state_dev && scale_out && n >= 0, "hn_dropout2d_scale: null pointer");
    HN_CHECK_ARG(p >= 0.f && p <= 1.f, "dropout probability has to be between 0 and 1, but got %g", (double)p);
    if (n == 0) return HN_OK;
    dropout2d_scale_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long *>(state_dev), n, p, scale_out);
    HN_LAUNCH_CHECK();
    return HN_OK;
}
