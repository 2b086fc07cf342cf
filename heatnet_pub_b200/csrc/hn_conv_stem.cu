// 7x7 stride-2 stems (cm/models/extractors.py:111-123; Cin = 3 / 1 / 4) as a tcgen05 GEMM whose im2col matrix is never
// built -- not even in shared memory.
//
// The input is the zero-bordered 4-channel BF16 image of hn_stem_pad: one pixel = 8 bytes, so the 8 pixels x 4 channels a
// filter row needs for output pixel wo are the 64 contiguous bytes that start at byte 16*wo of the padded input row, and
// the windows of consecutive output pixels overlap by 48 bytes.  A NO-SWIZZLE K-major UMMA operand is a grid of core
// matrices (8 rows x 16 bytes, rows 16 bytes apart) placed by two free strides: SBO between 8-row groups and LBO between
// 16-byte K chunks.  With SBO = 128 B and LBO = 16 B element (row m, chunk j) is read from byte 16*(m + j): the dense
// input row ITSELF is the A operand of 128 consecutive output pixels, overlaps included.  A tile therefore needs just the 7
// input rows of its output row segment (7 plain bulk copies of 2112 bytes) instead of 7 x 128 separate 64-byte windows;
// the 28 KB filter stays resident in shared memory in core-matrix order.  14 MMAs (7 filter rows x K = 32) per 128 pixels.
#include <stdlib.h>
#include <string.h>

#include "hn_common.cuh"
#include "hn_tc_epilogue.cuh"
#include "hn_tc_ptx.cuh"

namespace hn {

constexpr int ST_THREADS = 384;                 // warp 0 producer, warp 1 MMA, warp 2 TMEM alloc, warps 4-11 epilogue
constexpr int ST_ROW_BYTES = 2112;              // (2*127 + 8) pixels * 8 B = 2096, rounded up to a multiple of 64
constexpr int ST_STAGE_BYTES = 7 * ST_ROW_BYTES + 64;   // 14848: keeps every stage 128-byte aligned
constexpr int ST_STAGES = 6;
constexpr int ST_B_BYTES = 7 * 4096;            // [7 filter rows][8 cout groups][4 K chunks][8 couts][16 B]

__device__ __forceinline__ void bulk_load_1d_stem(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

// no-swizzle K-major descriptor: start address, LBO (K direction), SBO (M/N direction), version 1, layout type 0
__device__ __forceinline__ uint64_t make_kmajor_noswizzle_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

struct StemParams {
    TcParams t;
    const uint8_t *xpad;      // [N][Hp][Wp][4] BF16
    int Hp, Wp;
    const uint8_t *w;         // [64][7][8][4] BF16 (hn_pack_stem_weight)
};

__global__ void __launch_bounds__(ST_THREADS, 1)
conv_stem_dense_kernel(const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_r, const StemParams sp)
{
    constexpr int BLOCK_N = 64;
    constexpr int TMEM_COLS = 2 * BLOCK_N;
    constexpr uint32_t IDESC = make_idesc_bf16(128, BLOCK_N);
    const TcParams &p = sp.t;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *b_tile = smem;                                             // 28 KB, resident
    uint8_t *a_ring = b_tile + ST_B_BYTES;                              // ST_STAGES x ST_STAGE_BYTES
    uint8_t *epi_stage = a_ring + ((ST_STAGES * ST_STAGE_BYTES + 1023) / 1024) * 1024;
    uint64_t *bars = reinterpret_cast<uint64_t *>(epi_stage + NUM_EPI_WARPS * EPI_STAGE_BYTES);
    uint64_t *full_bar = bars, *empty_bar = bars + ST_STAGES, *tfull_bar = bars + 2 * ST_STAGES, *tempty_bar = tfull_bar + 2;
    uint64_t *res_bar = tempty_bar + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(res_bar + NUM_EPI_WARPS);
    float *s_shift = reinterpret_cast<float *>(tmem_slot + 4);         // 2 x 64 floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.n_img * p.tiles_h * p.tiles_w;

    if (warp == 0 && lane == 0 && p.tma_out) prefetch_tmap(&tmap_y);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < ST_STAGES; ++i) {
            mbar_init(smem_u32(full_bar + i), 1);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(tfull_bar + i), 1);
            mbar_init(smem_u32(tempty_bar + i), 4);
        }
        for (int i = 0; i < NUM_EPI_WARPS; ++i) mbar_init(smem_u32(res_bar + i), 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
        tmem_relinquish();
    }
    // filter -> core-matrix order: chunk (cout n, filter row r, K chunk j) = 16 bytes
    for (int i = threadIdx.x; i < 64 * 7 * 4; i += ST_THREADS) {
        const int n = i / 28, rj = i - n * 28, r = rj >> 2, j = rj & 3;
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(sp.w + ((size_t)n * 224 + r * 32 + j * 8) * 2));
        sts128(smem_u32(b_tile + r * 4096 + (n >> 3) * 512 + j * 128 + (n & 7) * 16), v);
    }
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== producer: 7 dense input-row segments per tile =====================
        int stage = 0;
        uint32_t phase = 0;
        const size_t row_pitch = (size_t)sp.Wp * 8;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int tw = tile % p.tiles_w, ho = (tile / p.tiles_w) % p.tiles_h, img = tile / (p.tiles_w * p.tiles_h);
            mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
            const uint32_t fb = smem_u32(full_bar + stage);
            const uint32_t sa = smem_u32(a_ring + stage * ST_STAGE_BYTES);
            const uint8_t *src = sp.xpad + ((size_t)img * sp.Hp + 2 * ho) * row_pitch + (size_t)tw * 128 * 16;
            if (elect_one()) {
                mbar_expect_tx(fb, 7 * ST_ROW_BYTES);
#pragma unroll
                for (int r = 0; r < 7; ++r) bulk_load_1d_stem(sa + r * ST_ROW_BYTES, src + r * row_pitch, ST_ROW_BYTES, fb);
            }
            __syncwarp();
            if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);
            mbar_wait(smem_u32(full_bar + stage), phase);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            const uint32_t sa = smem_u32(a_ring + stage * ST_STAGE_BYTES);
            if (elect_one()) {
#pragma unroll
                for (int r = 0; r < 7; ++r)
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        // A: row m, chunk j at 16*(m + j) of input row r; this MMA covers chunks 2kk, 2kk+1
                        const uint64_t adesc = make_kmajor_noswizzle_desc(sa + r * ST_ROW_BYTES + kk * 32, 16, 128);
                        const uint64_t bdesc = make_kmajor_noswizzle_desc(smem_u32(b_tile) + r * 4096 + kk * 256, 128, 512);
                        umma_bf16(d_tmem, adesc, bdesc, IDESC, (r | kk) != 0);
                    }
                umma_commit(smem_u32(empty_bar + stage));
                umma_commit(smem_u32(tfull_bar + acc));
            }
            __syncwarp();
            if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else if (warp >= EPI_WARP0) {
        conv_epilogue<BLOCK_N>(p, &tmap_y, &tmap_r, tmem_base, tfull_bar, tempty_bar, res_bar, epi_stage, s_shift, num_tiles, warp, lane);
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

int conv_stem_dense(const hn_tensor *xpad, const void *w, int cout, const hn_epilogue *ep, const hn_tensor *y, cudaStream_t st)
{
    const int Ho = y->h, Wo = y->w;
    if ((int64_t)y->n * Ho * Wo == 0) return HN_OK;
    HN_CHECK_ARG(xpad->c == 4 && xpad->ld == 4 && xpad->dtype == HN_BF16, "conv_stem: input must be the padded 4-channel BF16 image");
    HN_CHECK_ARG(xpad->w >= 2 * Wo + 6 && xpad->h >= 2 * Ho + 5 && xpad->w % 2 == 0,
                 "conv_stem: padded image too small or odd width (%dx%d for output %dx%d)", xpad->h, xpad->w, Ho, Wo);
    HN_CHECK_ARG((reinterpret_cast<uintptr_t>(xpad->ptr) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0, "conv_stem: 16-byte alignment");
    HN_CHECK_ARG(hn_conv_cout_pad(cout, HN_BF16) == 64, "conv_stem: Cout must be 33..64 (got %d)", cout);
    HN_CHECK_ARG(!ep->residual, "conv_stem: no residual input");
    StemParams sp{};
    TcParams &p = sp.t;
    sp.xpad = (const uint8_t *)xpad->ptr; sp.Hp = xpad->h; sp.Wp = xpad->w; sp.w = (const uint8_t *)w;
    p.hmul = 2;
    p.TH = 1; p.TW = 128;
    p.tiles_w = (int)cdiv(Wo, 128); p.tiles_h = Ho; p.n_img = y->n;
    p.Ho = Ho; p.Wo = Wo;
    p.R = 7; p.S = 1; p.pad = p.pad_w = 0; p.dil = 1; p.cblocks = 1;
    p.n_tiles = 1; p.Cout = cout;
    p.y = y->ptr; p.ldy = y->ld; p.y_f32 = (y->dtype == HN_F32);
    p.scale = ep->scale; p.shift = ep->shift; p.res = nullptr; p.ldr = 0;
    p.act = ep->act; p.slope = ep->slope; p.slope_ptr = ep->slope_ptr;
    p.stat_sum = ep->stat_sum; p.stat_sqsum = ep->stat_sqsum;
    if (int rcg = set_stat_groups(p, ep, y)) return rcg;
    p.ebw = 32;
    CUtensorMap ty, tr;
    memset(&ty, 0, sizeof(ty));
    memset(&tr, 0, sizeof(tr));
    {
        const uint64_t esz = p.y_f32 ? 4 : 2;
        const bool ok = (reinterpret_cast<uintptr_t>(y->ptr) & 15) == 0 && ((uint64_t)y->ld * esz) % 16 == 0;
        if (ok) {
            uint64_t dims[4] = {(uint64_t)cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)p.n_img};
            uint64_t strides[4] = {esz, (uint64_t)y->ld * esz, (uint64_t)y->ld * esz * Wo, (uint64_t)y->ld * esz * Wo * Ho};
            uint32_t box[4] = {(uint32_t)(128 / esz), 32, 1, 1};
            int rc = make_tmap(&ty, y->ptr, 4, dims, strides, box, p.y_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
            if (rc) return rc;
            p.tma_out = 1;
        }
    }
    constexpr size_t smem = ST_B_BYTES + ((ST_STAGES * ST_STAGE_BYTES + 1023) / 1024) * 1024 + NUM_EPI_WARPS * EPI_STAGE_BYTES +
                            (2 * ST_STAGES + 4 + NUM_EPI_WARPS) * 8 + 16 + 2 * 64 * 4 + 1024;
    static_assert(smem <= 227 * 1024, "stem kernel shared memory");
    static bool configured_dev[HN_MAX_DEVICES] = {};
    bool &configured = configured_dev[current_device()];
    if (!configured) {
        HN_CUDA(cudaFuncSetAttribute(conv_stem_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int num_tiles = p.n_img * p.tiles_h * p.tiles_w;
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    conv_stem_dense_kernel<<<grid, ST_THREADS, smem, st>>>(ty, tr, sp);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

}  // namespace hn

// bytes the buffer behind the padded image must extend past its last row: right-edge tiles read whole 2112-byte row segments
extern "C" int64_t hn_stem_pad_slack_bytes(void) { return 4096; }
