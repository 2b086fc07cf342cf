// Probe (dev tool, not part of the library): semantics of cuTensorMapEncodeTiled's elementStrides on sm_100a.
// Question: to fetch TW x TH pixels at a spatial stride of 2 (the A operand of a stride-2 convolution as an implicit GEMM) is the
// box given as the TRAVERSED extent (boxDim = 2*TW, the loaded element count being ceil(boxDim / elementStride)) and how many
// bytes does the transaction deposit?  The kernel waits on the mbarrier with a bounded spin so a wrong guess cannot hang the GPU.
//   build+run:  nvcc -gencode arch=compute_100a,code=sm_100a -I heatnet_pub_b200/csrc -I include scripts/probe_tma_stride.cu -o /tmp/probe_tma -lcuda && /tmp/probe_tma
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hn_tc_ptx.cuh"
using namespace hn;

constexpr int H = 24, W = 40, C = 64, TH = 4, TW = 8;

__global__ void probe(const __grid_constant__ CUtensorMap tmap, __nv_bfloat16 *out /*[TH*TW][64]*/, int *status, int x0, int y0, int expect_bytes)
{
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * TH * TW * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0xdeadbeefu;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(smem_u32(&bar), expect_bytes);
        tma_load_4d(smem_u32(smem), &tmap, smem_u32(&bar), 0, x0, y0, 0);
    }
    // bounded wait
    int ok = 0;
    for (int it = 0; it < 2000000 && !ok; ++it) {
        uint32_t done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
        ok = done;
    }
    if (tid == 0) status[0] = ok;
    __syncthreads();
    // un-swizzle (128B swizzle: chunk j of row p sits at chunk j ^ (p & 7))
    for (int i = tid; i < 2 * TH * TW * 8; i += blockDim.x) {
        int p = i / 8, j = i % 8;
        uint4 v = *reinterpret_cast<const uint4 *>(smem + p * 128 + ((j ^ (p & 7)) << 4));
        *reinterpret_cast<uint4 *>(out + p * 64 + j * 8) = v;
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    std::vector<__nv_bfloat16> x((size_t)H * W * C);
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w)
            for (int c = 0; c < C; ++c) x[((size_t)h * W + w) * C + c] = __float2bfloat16((float)(h * 64 + w + (c == 0 ? 0 : 0)));   // value = 64h + w
    __nv_bfloat16 *dx, *dout;
    int *dstatus;
    cudaMalloc(&dx, x.size() * 2);
    cudaMalloc(&dout, 2 * TH * TW * 64 * 2);
    cudaMalloc(&dstatus, 4);
    cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    PFN_encodeTiled enc = (PFN_encodeTiled)fn;
    for (int variant = 0; variant < 2; ++variant) {
        // variant 0: boxDim = traversed extent (2*TW, 2*TH); variant 1: boxDim = loaded element count (TW, TH)
        CUtensorMap tm;
        cuuint64_t gd[4] = {C, W, H, 1};
        cuuint64_t gs[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t bx[4] = {64, (cuuint32_t)(variant == 0 ? 2 * TW : TW), (cuuint32_t)(variant == 0 ? 2 * TH : TH), 1};
        cuuint32_t es[4] = {1, 2, 2, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("variant %d (box %u x %u): encode rc=%d\n", variant, bx[1], bx[2], (int)r);
        if (r != CUDA_SUCCESS) continue;
        for (int expect : {TH * TW * 128, TH * TW * 128 / 4, 4 * TH * TW * 128}) {
            const int x0 = -1, y0 = 3;       // negative start: zero fill on the left, stride-2 samples at w = -1, 1, 3, ...
            cudaMemset(dout, 0, 2 * TH * TW * 64 * 2);
            probe<<<1, 128, 2 * TH * TW * 128 + 2048>>>(tm, dout, dstatus, x0, y0, expect);
            cudaError_t e = cudaDeviceSynchronize();
            int status = -1;
            cudaMemcpy(&status, dstatus, 4, cudaMemcpyDeviceToHost);
            std::vector<__nv_bfloat16> o(2 * TH * TW * 64);
            cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost);
            int good = 0, touched = 0;
            for (int p = 0; p < 2 * TH * TW; ++p) {
                float v = __bfloat162float(o[p * 64 + 5]);
                uint16_t bits;
                memcpy(&bits, &o[p * 64 + 5], 2);
                if (bits != 0xbeef && bits != 0xdead) ++touched;
                if (p < TH * TW) {
                    int ty = p / TW, tx = p % TW;
                    int h = y0 + 2 * ty, w = x0 + 2 * tx;
                    float want = (h >= 0 && h < H && w >= 0 && w < W) ? (float)(h * 64 + w) : 0.f;
                    if (v == want) ++good;
                }
            }
            printf("  expect_tx %5d: cuda=%s barrier_completed=%d rows_written=%d rows_matching_stride2_pattern=%d/%d  first rows:", expect,
                   cudaGetErrorString(e), status, touched, good, TH * TW);
            for (int p = 0; p < 10; ++p) printf(" %g", __bfloat162float(o[p * 64 + 5]));
            printf("\n");
        }
    }
    return 0;
}
