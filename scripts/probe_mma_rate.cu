// Probe (dev tool, not part of the library): sustained cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) as a
// function of N, of the A-operand descriptor (canonical SWIZZLE_128B tile vs. the halo kernel's sub-window with a 1280-byte
// row-group pitch), of where consecutive MMAs read A (same 32-byte K slice vs. walking the 4 K slices of a 64-channel block)
// and of the number of independent accumulators.  Operand contents are irrelevant (shared memory is left uninitialised).
//   build+run:  nvcc -gencode arch=compute_100a,code=sm_100a -I heatnet_pub_b200/csrc -I include scripts/probe_mma_rate.cu -o /tmp/probe_rate && /tmp/probe_rate
#include <cstdio>
#include <cstdlib>

#include "hn_tc_ptx.cuh"
using namespace hn;

struct Cfg {
    int n;          // MMA N
    int sbo;        // A stride-byte-offset (1024 canonical, 1280 = halo patch with PW = 10)
    int a_step;     // 1: consecutive MMAs walk the four 32-byte K slices; 0: always slice 0
    int nacc;       // independent accumulators used round-robin
    int m64;        // 1: M = 64 instruction
};

__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg c, int iters, long long *out)
{
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 64 * 1024);
        uint64_t adesc = 0;
        adesc |= (uint64_t)((sa & 0x3FFFFu) >> 4);
        adesc |= (uint64_t)1 << 16;
        adesc |= (uint64_t)((c.sbo >> 4) & 0x3FFF) << 32;
        adesc |= (uint64_t)1 << 46;
        adesc |= (uint64_t)2 << 61;
        const uint64_t bdesc = make_kmajor_sw128_desc(sb);
        const uint32_t idesc = make_idesc_bf16(c.m64 ? 64 : 128, c.n);
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2; ++rep) {          // rep 0 warms up
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int idx = i * 4 + k;
                        umma_bf16(tmem + (idx % c.nacc) * c.n, adesc + 2 * (c.a_step ? k : 0),
                                  bdesc + 2 * (c.a_step ? k : 0), idesc, 1);
                    }
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(smem_u32(&bar));
            __syncwarp();
            mbar_wait(smem_u32(&bar), rep & 1);
            t1 = clock64();
        }
        if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tmem, 512); }
}

int main()
{
    long long *d;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    Cfg cfgs[] = {
        {256, 1024, 1, 1, 0}, {128, 1024, 1, 1, 0}, {64, 1024, 1, 1, 0}, {32, 1024, 1, 1, 0}, {16, 1024, 1, 1, 0},
        {64, 1280, 1, 1, 0},  {64, 1024, 0, 1, 0},  {64, 1024, 1, 2, 0}, {64, 1024, 1, 4, 0}, {128, 1024, 1, 2, 0},
        {64, 1024, 1, 1, 1},  {128, 1024, 1, 1, 1}, {256, 1024, 1, 1, 1},
    };
    for (int grid : {1, 148}) {
        for (const Cfg &c : cfgs) {
            rate_kernel<<<grid, 128, 200 * 1024>>>(c, iters, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            long long h[148];
            cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            const double per = (double)mx / (iters * 4.0);
            const double macs = (c.m64 ? 64.0 : 128.0) * c.n * 16;
            printf("grid %3d  M=%3d N=%3d sbo=%4d a_step=%d nacc=%d : %7.1f cycles/MMA  (%.0f MAC/cycle/SM; floor %d)\n", grid, c.m64 ? 64 : 128,
                   c.n, c.sbo, c.a_step, c.nacc, per, macs / per, 128 * c.n / 256);
        }
    }
    return 0;
}
