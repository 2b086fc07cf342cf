// iou_eval confusion matrix on the device (replaces the D2H + np.bincount of scripts/iou_eval.py:53-88
// and the `max(1)` of :154-157).  HBM-bound integer work: 16 B/pixel on the label path, (4K+8) B/pixel on
// the scores path.
//
// Counting scheme: no atomics in the hot loop.  Every warp owns a table  word[ceil(K*K/2)][32 lanes]  in
// shared memory; lane L only ever touches column L (bank L), so the read-modify-write is conflict-free and
// race-free.  One 32-bit word packs the 16-bit counters of bins 2j and 2j+1; a warp flushes its table to the
// global int64 matrix (warp-shuffle reduction + one atomicAdd per bin) before any counter can overflow.
#include <stdlib.h>

#include "hn_common.cuh"
#include "hn_tc_ptx.cuh"

namespace hn {

constexpr int kLabelUnroll = 4;
constexpr int kMaxIterBeforeFlush = 8000;  // x2 pixels per lane per load x4 unroll = 64000 < 65535 (+ the tail loop's < 8)

// streaming 16-byte load: read-only path, no L1 allocation (every label is touched once)
__device__ __forceinline__ longlong2 ld_stream(const longlong2 *p)
{
    longlong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ void count_pixel(uint32_t *table, int lane, long long p, long long t, int K, int &flags)
{
    bool pbad = (unsigned long long)p >= (unsigned long long)K;
    bool tbad = (unsigned long long)t >= (unsigned long long)K;
    if (pbad | tbad) {
        flags |= (pbad ? 1 : 0) | (tbad ? 2 : 0);
        return;
    }
    int bin = (int)t * K + (int)p;
    // lane-private word: never contended, but the shared-memory atomic (result unused) is fire-and-forget, whereas a
    // load-add-store chain on possibly-aliasing addresses serialises the pixels of a thread on the LDS latency
    atomicAdd(table + (bin >> 1) * 32 + lane, (bin & 1) ? 0x10000u : 1u);
}

__device__ __forceinline__ void flush_table(uint32_t *table, int lane, int nwords, int nbins,
                                            unsigned long long *conf)
{
    __syncwarp();
    for (int j = 0; j < nwords; ++j) {
        uint32_t v = table[j * 32 + lane];
        table[j * 32 + lane] = 0;
        uint32_t lo = v & 0xFFFFu, hi = v >> 16;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo += __shfl_xor_sync(0xffffffffu, lo, o);
            hi += __shfl_xor_sync(0xffffffffu, hi, o);
        }
        if (lane == 0) {
            if (lo) atomicAdd(conf + 2 * j, (unsigned long long)lo);
            if (hi && 2 * j + 1 < nbins) atomicAdd(conf + 2 * j + 1, (unsigned long long)hi);
        }
    }
    __syncwarp();
}

// label path: pred and target are int64 [n].  Latency-bound unless enough bytes are in flight (Little: ~35 KB per SM at
// 6.5 TB/s): each lane issues 2*U independent 16-byte loads before the first use, 16 warps per SM -> 64 KB in flight.
template <int U>
__global__ void __launch_bounds__(512) confusion_labels_kernel(const long long *__restrict__ pred,
                                                               const long long *__restrict__ target, long long n, int K,
                                                               unsigned long long *conf, int *flags_out)
{
    extern __shared__ uint32_t smem[];
    const int nbins = K * K, nwords = (nbins + 1) / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *table = smem + warp * nwords * 32;
    for (int j = 0; j < nwords; ++j) table[j * 32 + lane] = 0;
    __syncwarp();

    int flags = 0, iters = 0;
    const long long nvec = n >> 1;  // pairs of pixels (16-byte loads)
    const longlong2 *p2 = reinterpret_cast<const longlong2 *>(pred);
    const longlong2 *t2 = reinterpret_cast<const longlong2 *>(target);
    const long long gthreads = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * gthreads < nvec; i += U * gthreads) {
        longlong2 pv[U], tv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            pv[u] = ld_stream(p2 + i + u * gthreads);
            tv[u] = ld_stream(t2 + i + u * gthreads);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            count_pixel(table, lane, pv[u].x, tv[u].x, K, flags);
            count_pixel(table, lane, pv[u].y, tv[u].y, K, flags);
        }
        if (++iters == kMaxIterBeforeFlush) {
            flush_table(table, lane, nwords, nbins, conf);
            iters = 0;
        }
    }
    for (; i < nvec; i += gthreads) {
        longlong2 pa = __ldg(p2 + i), ta = __ldg(t2 + i);
        count_pixel(table, lane, pa.x, ta.x, K, flags);
        count_pixel(table, lane, pa.y, ta.y, K, flags);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) count_pixel(table, lane, pred[n - 1], target[n - 1], K, flags);
    flush_table(table, lane, nwords, nbins, conf);
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0 && flags) atomicOr(flags_out, flags);
}

// label path, TMA-staged (K <= 15): a producer warp streams 1024-pixel chunks of pred and target into a 6-stage shared-memory
// ring with cp.async.bulk (96 KB in flight per SM, independent of the consumers' register budget); 8 consumer warps count
// from shared memory with 16-byte loads.  Same per-lane private counters and flush rule as above.
constexpr int CL_WARPS = 8;
constexpr int CL_CHUNK = 1024;                 // pixels per stage: 8 KB of pred + 8 KB of target
constexpr int CL_STAGES = 6;
constexpr int CL_STAGE_BYTES = 2 * CL_CHUNK * 8;

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

__global__ void __launch_bounds__((CL_WARPS + 1) * 32, 1) confusion_labels_tma_kernel(const long long *__restrict__ pred,
                                                                                      const long long *__restrict__ target, long long n, int K,
                                                                                      unsigned long long *conf, int *flags_out)
{
    extern __shared__ __align__(128) uint8_t cl_smem[];
    const int nbins = K * K, nwords = (nbins + 1) / 2;
    uint8_t *ring = cl_smem;
    uint32_t *tables = reinterpret_cast<uint32_t *>(ring + CL_STAGES * CL_STAGE_BYTES);
    uint64_t *full = reinterpret_cast<uint64_t *>(tables + CL_WARPS * nwords * 32), *empty = full + CL_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_even = n & ~1LL;
    const long long nchunks = (n_even + CL_CHUNK - 1) / CL_CHUNK;
    if (threadIdx.x == 0) {
        for (int i = 0; i < CL_STAGES; ++i) {
            mbar_init(smem_u32(full + i), 1);
            mbar_init(smem_u32(empty + i), CL_WARPS);
        }
        fence_barrier_init();
    }
    uint32_t *table = tables + (warp < CL_WARPS ? warp : 0) * nwords * 32;
    if (warp < CL_WARPS)
        for (int j = 0; j < nwords; ++j) table[j * 32 + lane] = 0;
    __syncthreads();

    if (warp == CL_WARPS) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
                mbar_wait(smem_u32(empty + stage), phase ^ 1);
                const long long left = n_even - c * CL_CHUNK;
                const uint32_t bytes = (uint32_t)(left < CL_CHUNK ? left : CL_CHUNK) * 8u;
                const uint32_t fb = smem_u32(full + stage);
                mbar_expect_tx(fb, 2 * bytes);
                const uint32_t dst = smem_u32(ring + stage * CL_STAGE_BYTES);
                bulk_load_1d(dst, pred + c * CL_CHUNK, bytes, fb);
                bulk_load_1d(dst + CL_CHUNK * 8, target + c * CL_CHUNK, bytes, fb);
                if (++stage == CL_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }
    int flags = 0, iters = 0, stage = 0;
    uint32_t phase = 0;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        mbar_wait(smem_u32(full + stage), phase);
        const long long left = n_even - c * CL_CHUNK;
        const int pairs = (int)(left < CL_CHUNK ? left : CL_CHUNK) >> 1;
        const longlong2 *ps = reinterpret_cast<const longlong2 *>(ring + stage * CL_STAGE_BYTES);
        const longlong2 *ts = ps + CL_CHUNK / 2;
#pragma unroll
        for (int u = 0; u < CL_CHUNK / 2 / (CL_WARPS * 32); ++u) {
            const int i = u * CL_WARPS * 32 + threadIdx.x;
            if (i < pairs) {
                const longlong2 pv = ps[i], tv = ts[i];
                count_pixel(table, lane, pv.x, tv.x, K, flags);
                count_pixel(table, lane, pv.y, tv.y, K, flags);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(empty + stage));
        if (++stage == CL_STAGES) { stage = 0; phase ^= 1; }
        if (++iters == 2 * kMaxIterBeforeFlush) {        // <= 4 pixels per lane per chunk
            flush_table(table, lane, nwords, nbins, conf);
            iters = 0;
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) count_pixel(table, lane, pred[n - 1], target[n - 1], K, flags);
    flush_table(table, lane, nwords, nbins, conf);
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0 && flags) atomicOr(flags_out, flags);
}

// scores path: NCHW FP32 [N][K][hw]; first-max argmax over K fused (torch max(1) tie rule)
__global__ void __launch_bounds__(256) confusion_scores_kernel(const float *__restrict__ scores,
                                                               const long long *__restrict__ target, long long n_images,
                                                               long long hw, int K, unsigned long long *conf,
                                                               int *flags_out)
{
    extern __shared__ uint32_t smem[];
    const int nbins = K * K, nwords = (nbins + 1) / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *table = smem + warp * nwords * 32;
    for (int j = 0; j < nwords; ++j) table[j * 32 + lane] = 0;
    __syncwarp();
    int flags = 0, iters = 0;
    const long long total = n_images * hw;
    const long long gthreads = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gthreads) {
        long long img = i / hw, px = i - img * hw;
        const float *s = scores + img * K * hw + px;
        float best = __ldg(s);
        int arg = 0;
        for (int k = 1; k < K; ++k) {
            float v = __ldg(s + (long long)k * hw);
            if (v > best) { best = v; arg = k; }
        }
        count_pixel(table, lane, arg, __ldg(target + i), K, flags);
        if (++iters == 4 * kMaxIterBeforeFlush) {
            flush_table(table, lane, nwords, nbins, conf);
            iters = 0;
        }
    }
    flush_table(table, lane, nwords, nbins, conf);
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0 && flags) atomicOr(flags_out, flags);
}

__global__ void __launch_bounds__(256) argmax_labels_kernel(const float *__restrict__ scores, long long n_images,
                                                            long long hw, int K, uint8_t *out_u8, long long *out_i64)
{
    const long long total = n_images * hw;
    const long long gthreads = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gthreads) {
        long long img = i / hw, px = i - img * hw;
        const float *s = scores + img * K * hw + px;
        float best = __ldg(s);
        int arg = 0;
        for (int k = 1; k < K; ++k) {
            float v = __ldg(s + (long long)k * hw);
            if (v > best) { best = v; arg = k; }
        }
        if (out_u8) out_u8[i] = (uint8_t)arg;
        if (out_i64) out_i64[i] = arg;
    }
}

}  // namespace hn

extern "C" int hn_confusion(const int64_t *pred_labels, const float *scores, int64_t n_images, int64_t hw,
                            const int64_t *target, int32_t k, int64_t *conf, int32_t *flags, void *stream)
{
    using namespace hn;
    HN_CHECK_ARG((pred_labels != nullptr) != (scores != nullptr), "hn_confusion: give exactly one of pred_labels / scores");
    HN_CHECK_ARG(target && conf && flags, "hn_confusion: null pointer");
    HN_CHECK_ARG(k >= 1 && k <= 32, "hn_confusion: K=%d outside [1,32]", k);
    HN_CHECK_ARG(n_images >= 0 && hw >= 0, "hn_confusion: negative size");
    const int64_t n = n_images * hw;
    if (n == 0) return HN_OK;
    const int nwords = (k * k + 1) / 2;
    // one private table per warp: as many warps (<= 8) as fit in ~200 KB of shared memory
    const int max_warps = pred_labels ? 16 : 8;
    int warps = (int)((200 * 1024) / ((size_t)nwords * 32 * sizeof(uint32_t)));
    warps = warps < 1 ? 1 : (warps > max_warps ? max_warps : warps);
    const int threads = warps * 32;
    const size_t smem = (size_t)warps * nwords * 32 * sizeof(uint32_t);
    cudaStream_t st = (cudaStream_t)stream;
    // one resident wave: grid = SMs x CTAs/SM that fit in shared memory
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int64_t want = cdiv(scores ? n : cdiv(n, 2), threads);
    int grid = (int)((want < (int64_t)num_sms() * per_sm) ? want : (int64_t)num_sms() * per_sm);
    if (pred_labels) {
        HN_CHECK_ARG((reinterpret_cast<uintptr_t>(pred_labels) | reinterpret_cast<uintptr_t>(target)) % 16 == 0,
                     "hn_confusion: label pointers must be 16-byte aligned");
        static const bool no_tma = getenv("HN_CONFUSION_NO_TMA") != nullptr;
        const size_t tsmem = (size_t)CL_STAGES * CL_STAGE_BYTES + (size_t)CL_WARPS * nwords * 32 * sizeof(uint32_t) + 2 * CL_STAGES * 8;
        if (!no_tma && tsmem <= 226 * 1024 && n >= 4 * CL_CHUNK) {
            HN_CUDA(cudaFuncSetAttribute(confusion_labels_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
            const int64_t nch = cdiv(n & ~(int64_t)1, CL_CHUNK);
            const int g = (int)(nch < num_sms() ? nch : num_sms());
            confusion_labels_tma_kernel<<<g, (CL_WARPS + 1) * 32, tsmem, st>>>((const long long *)pred_labels, (const long long *)target, n, k,
                                                                               (unsigned long long *)conf, flags);
            HN_LAUNCH_CHECK();
            return HN_OK;
        }
        HN_CUDA(cudaFuncSetAttribute(confusion_labels_kernel<kLabelUnroll>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        confusion_labels_kernel<kLabelUnroll><<<grid, threads, smem, st>>>((const long long *)pred_labels, (const long long *)target, n, k,
                                                         (unsigned long long *)conf, flags);
    } else {
        HN_CUDA(cudaFuncSetAttribute(confusion_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        confusion_scores_kernel<<<grid, threads, smem, st>>>(scores, (const long long *)target, n_images, hw, k,
                                                         (unsigned long long *)conf, flags);
    }
    HN_LAUNCH_CHECK();
    return HN_OK;
}

extern "C" int hn_argmax_labels(const float *scores, int64_t n_images, int64_t hw, int32_t k, uint8_t *out_u8,
                                int64_t *out_i64, void *stream)
{
    using namespace hn;
    HN_CHECK_ARG(scores && (out_u8 || out_i64), "hn_argmax_labels: null pointer");
    HN_CHECK_ARG(k >= 1 && k <= 255, "hn_argmax_labels: K=%d", k);
    const int64_t n = n_images * hw;
    if (n == 0) return HN_OK;
    int64_t want = cdiv(n, 256);
    int grid = (int)((want < (int64_t)num_sms() * 8) ? want : (int64_t)num_sms() * 8);
    argmax_labels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(scores, n_images, hw, k, out_u8, (long long *)out_i64);
    HN_LAUNCH_CHECK();
    return HN_OK;
}
