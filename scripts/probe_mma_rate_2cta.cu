// Probe (dev tool, not part of the library): sustained cycles per tcgen05.mma with cta_group::2 (a CTA pair, M = 256 across the two
// SMs, each CTA holding its 128 A rows and half of the B rows in its own shared memory), as a function of N.  Question asked by
// the round-1 review: the single-CTA SS-mode instruction never retires faster than ~82 cycles whatever N <= 128
// (profiles/r1_mma_rate_probe.txt), which caps the Cout = 64 layers at 39 % of the tensor pipe -- does the pair form lift that?
// Every wait is bounded so that a wrong assumption cannot hang the GPU.
//   build+run:  nvcc -gencode arch=compute_100a,code=sm_100a -I heatnet_pub_b200/csrc -I include scripts/probe_mma_rate_2cta.cu -o /tmp/probe2 && /tmp/probe2
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>

#include "hn_tc_ptx.cuh"
using namespace hn;
namespace cg = cooperative_groups;

__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool bounded_wait(uint32_t bar, uint32_t parity)
{
    for (int it = 0; it < 4000000; ++it) {
        uint32_t done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity));
        if (done) return true;
    }
    return false;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate2_kernel(int n, int iters, long long *out, int *status)
{
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc2(smem_u32(&tmem_slot), 512); tmem_relinquish2(); }
    tcgen05_fence_before();
    cluster.sync();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    int ok = 1;
    if (warp == 1 && rank == 0) {          // the leader CTA issues for the pair
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 64 * 1024);
        const uint64_t adesc = make_kmajor_sw128_desc(sa);
        const uint64_t bdesc = make_kmajor_sw128_desc(sb);
        const uint32_t idesc = make_idesc_bf16(256, n);
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2 && ok; ++rep) {          // rep 0 warms up
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma2_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
                }
                __syncwarp();
            }
            if (elect_one()) umma2_commit(smem_u32(&bar));
            __syncwarp();
            ok = bounded_wait(smem_u32(&bar), rep & 1) ? 1 : 0;
            t1 = clock64();
        }
        if ((threadIdx.x & 31) == 0) { out[blockIdx.x / 2] = t1 - t0; status[blockIdx.x / 2] = ok; }
    }
    tcgen05_fence_before();
    cluster.sync();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc2(tmem, 512); }
}

int main()
{
    long long *d;
    int *st;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaMalloc(&st, 148 * sizeof(int));
    cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    for (int grid : {2, 148}) {
        for (int n : {256, 128, 64, 32}) {
            cudaMemset(st, 0, 148 * sizeof(int));
            rate2_kernel<<<grid, 128, 200 * 1024>>>(n, iters, d, st);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("N=%d grid %d: error: %s\n", n, grid, cudaGetErrorString(e)); return 1; }
            long long h[148];
            int hs[148];
            cudaMemcpy(h, d, (grid / 2) * sizeof(long long), cudaMemcpyDeviceToHost);
            cudaMemcpy(hs, st, (grid / 2) * sizeof(int), cudaMemcpyDeviceToHost);
            long long mx = 0;
            int all_ok = 1;
            for (int i = 0; i < grid / 2; ++i) { mx = h[i] > mx ? h[i] : mx; all_ok &= hs[i]; }
            const double per = (double)mx / (iters * 4.0);
            const double macs_per_sm = 128.0 * n * 16;          // each SM of the pair computes its 128 rows x N x 16
            printf("grid %3d (pairs %2d)  cta_group::2 M=256 N=%3d : %7.1f cycles/MMA  (%.0f MAC/cycle/SM; pipe floor %d cycles; barrier %s)\n", grid,
                   grid / 2, n, per, macs_per_sm / per, 128 * n / 256, all_ok ? "completed" : "TIMED OUT");
        }
    }
    return 0;
}
