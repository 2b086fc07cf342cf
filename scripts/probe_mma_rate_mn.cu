// Probe (dev tool): sustained cycles per tcgen05.mma when the operands are MN-major (the weight-gradient kernel's layout: NHWC tiles
// whose GEMM M / N index is the contiguous one), single CTA (M = 128) and CTA pair (M = 256), N = 256 / 128.
// The descriptors mirror hn_conv_wgrad.cu: [64 px rows][128 B] blocks, LBO = 8 KB between 64-wide MN blocks, K step = 16 rows = 2 KB.
//   build+run:  nvcc -gencode arch=compute_100a,code=sm_100a -I heatnet_pub_b200/csrc -I include scripts/probe_mma_rate_mn.cu -o /tmp/probe_mn && /tmp/probe_mn
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>

#include "hn_tc_ptx.cuh"
using namespace hn;
namespace cg = cooperative_groups;

__device__ __forceinline__ bool bounded_wait(uint32_t bar, uint32_t parity)
{
    for (int it = 0; it < 4000000; ++it) {
        uint32_t done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity));
        if (done) return true;
    }
    return false;
}

// mode: bit 0 = A MN-major, bit 1 = B MN-major, bit 2 = the descriptors change from one group of 4 MMAs to the next (as they do in the
// convolution kernels: a new stage / filter tap per group), bit 3 = every MMA gets its own recomputed descriptor
template <bool PAIR>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int mode, int iters, long long *out, int *status)
{
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const unsigned rank = PAIR ? cluster_ctarank() : 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) {
        if (PAIR) { tmem_alloc_pair(smem_u32(&tmem_slot), 512); tmem_relinquish_pair(); }
        else { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
    }
    tcgen05_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    int ok = 1;
    if (warp == 1 && rank == 0) {
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 64 * 1024);
        const bool amn = mode & 1, bmn = mode & 2;
        const uint64_t adesc = amn ? make_mnmajor_sw128_desc(sa, 8192) : make_kmajor_sw128_desc(sa);
        const uint64_t bdesc = bmn ? make_mnmajor_sw128_desc(sb, 8192) : make_kmajor_sw128_desc(sb);
        const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, n) | (amn ? (1u << 15) : 0u) | (bmn ? (1u << 16) : 0u);
        const int astep = amn ? 128 : 2, bstep = bmn ? 128 : 2;          // K += 16: 2 KB for MN-major rows, 32 B inside a K-major row
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2 && ok; ++rep) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                // (offsets of 0 / 1 KB stay inside the operand buffers and keep the 1024-byte swizzle phase)
                const uint64_t ai = (mode & 4) ? adesc + (uint64_t)((i & 3) << 6) : adesc;
                const uint64_t bi = (mode & 4) ? bdesc + (uint64_t)((i & 1) << 6) : bdesc;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t ak = (mode & 8) ? adesc + (uint64_t)(((i * 4 + k) % 5) << 6) : ai;
                        umma_bf16_t<PAIR>(tmem, ak + astep * k, bi + bstep * k, idesc, 1);
                    }
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit_t<PAIR>(smem_u32(&bar));
            __syncwarp();
            ok = bounded_wait(smem_u32(&bar), rep & 1) ? 1 : 0;
            t1 = clock64();
        }
        if ((threadIdx.x & 31) == 0) { out[blockIdx.x >> (PAIR ? 1 : 0)] = t1 - t0; status[blockIdx.x >> (PAIR ? 1 : 0)] = ok; }
    }
    tcgen05_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512);
    }
}

template <bool PAIR>
static void run(int grid, int n, int mode, long long *d, int *st)
{
    const int iters = 2000;
    cudaMemset(st, 0, 148 * sizeof(int));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, rate_kernel<PAIR>, n, mode, iters, d, st);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
    const int units = PAIR ? grid / 2 : grid;
    long long h[148];
    int hs[148];
    cudaMemcpy(h, d, units * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(hs, st, units * sizeof(int), cudaMemcpyDeviceToHost);
    long long mx = 0;
    int all_ok = 1;
    for (int i = 0; i < units; ++i) { mx = h[i] > mx ? h[i] : mx; all_ok &= hs[i]; }
    const double per = (double)mx / (iters * 4.0);
    const char *names[4] = {"A K-major,  B K-major ", "A MN-major, B K-major ", "A K-major,  B MN-major", "A MN-major, B MN-major"};
    const char *var = (mode & 8) ? "descriptor recomputed per MMA  " : ((mode & 4) ? "descriptors change per 4 MMAs " : "loop-invariant descriptors     ");
    printf("grid %3d %s N=%3d %s %s: %7.1f cycles/MMA  (%.0f MAC/cycle/SM of 4096; %s)\n", grid, PAIR ? "pair  M=256" : "single M=128", n, names[mode & 3], var, per,
           128.0 * n * 16 / per, all_ok ? "completed" : "TIMED OUT");
}

int main()
{
    long long *d;
    int *st;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaMalloc(&st, 148 * sizeof(int));
    cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int n : {256, 128})
        for (int mode = 0; mode < 4; ++mode) {
            run<false>(148, n, mode, d, st);
            run<true>(148, n, mode, d, st);
        }
    for (int n : {128, 64, 32})
        for (int mode : {0, 4, 8}) {
            run<false>(148, n, mode, d, st);
            run<true>(148, n, mode, d, st);
        }
    return 0;
}
