// Probe (dev tool, not part of the library): can a K-major SWIZZLE_128B UMMA operand start at a row that is NOT a
// multiple of 8 (i.e. not 1024-byte aligned), with 8-row groups SBO bytes apart?  This is what a convolution needs
// to feed all 9 taps of a 3x3 filter from ONE halo patch in shared memory.
//   build+run:  nvcc -gencode arch=compute_100a,code=sm_100a -I heatnet_pub_b200/csrc -I include scripts/probe_umma_offset.cu -o /tmp/probe && /tmp/probe
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "hn_tc_ptx.cuh"
using namespace hn;

constexpr int PATCH_ROWS = 192, N = 64, K = 64;

__global__ void probe(const __nv_bfloat16 *patch /*[PATCH_ROWS][64]*/, const __nv_bfloat16 *b /*[N][64]*/, float *d /*[128][N]*/, int r0,
                      int sbo_bytes, int base_off_mode)
{
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sa = smem, *sb = smem + PATCH_ROWS * 128;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    // absolute-address swizzle: 16-byte chunk j of the 128-byte row at byte address A goes to chunk j ^ ((A >> 7) & 7)
    for (int i = tid; i < PATCH_ROWS * 8; i += blockDim.x) {
        int p = i / 8, j = i % 8;
        uint32_t rowaddr = smem_u32(sa + p * 128);
        uint32_t dst = rowaddr + ((j ^ ((rowaddr >> 7) & 7)) << 4);
        uint4 v = *reinterpret_cast<const uint4 *>(patch + p * 64 + j * 8);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
    }
    for (int i = tid; i < N * 8; i += blockDim.x) {
        int p = i / 8, j = i % 8;
        uint32_t rowaddr = smem_u32(sb + p * 128);
        uint32_t dst = rowaddr + ((j ^ ((rowaddr >> 7) & 7)) << 4);
        uint4 v = *reinterpret_cast<const uint4 *>(b + p * 64 + j * 8);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
    }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 64); tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t a0 = smem_u32(sa) + r0 * 128;
        uint64_t adesc = 0;
        adesc |= (uint64_t)((a0 & 0x3FFFFu) >> 4);
        adesc |= (uint64_t)1 << 16;
        adesc |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
        adesc |= (uint64_t)1 << 46;
        if (base_off_mode == 1) adesc |= (uint64_t)((a0 >> 7) & 7) << 49;
        adesc |= (uint64_t)2 << 61;
        const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(sb));
        const uint32_t idesc = make_idesc_bf16(128, N);
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0);
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tcgen05_fence_after();
    if (warp < 4) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) d[(warp * 32 + (tid & 31)) * N + c0 + j] = __uint_as_float(r[j]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

int main()
{
    std::vector<__nv_bfloat16> patch(PATCH_ROWS * 64), b(N * 64);
    std::vector<float> pf(PATCH_ROWS * 64), bf(N * 64);
    srand(1);
    for (size_t i = 0; i < patch.size(); ++i) { pf[i] = (float)(rand() % 7 - 3); patch[i] = __float2bfloat16(pf[i]); }
    for (size_t i = 0; i < b.size(); ++i) { bf[i] = (float)(rand() % 5 - 2); b[i] = __float2bfloat16(bf[i]); }
    __nv_bfloat16 *dp, *db;
    float *dd;
    cudaMalloc(&dp, patch.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dd, 128 * N * 4);
    cudaMemcpy(dp, patch.data(), patch.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    size_t smem = PATCH_ROWS * 128 + N * 128 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<float> got(128 * N);
    int r0s[] = {0, 8, 3, 10, 11, 21};
    int sbos[] = {1024, 1280, 2304};
    for (int sbo : sbos)
        for (int r0 : r0s)
            for (int mode = 0; mode < 2; ++mode) {
                cudaMemset(dd, 0, 128 * N * 4);
                probe<<<1, 128, smem>>>(dp, db, dd, r0, sbo, mode);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("sbo %d r0 %d mode %d: CUDA error %s\n", sbo, r0, mode, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(got.data(), dd, got.size() * 4, cudaMemcpyDeviceToHost);
                int bad = 0;
                for (int i = 0; i < 128; ++i) {
                    int prow = r0 + (i / 8) * (sbo / 128) + (i % 8);
                    for (int n = 0; n < N; ++n) {
                        float ref = 0;
                        for (int k = 0; k < K; ++k) ref += pf[prow * 64 + k] * bf[n * 64 + k];
                        if (ref != got[i * N + n]) ++bad;
                    }
                }
                printf("sbo %4d  r0 %2d  base_offset %s : %s (%d mismatches)\n", sbo, r0, mode ? "(addr>>7)&7" : "0", bad ? "WRONG" : "ok", bad);
            }
    return 0;
}
