// Convolution weight gradient:  dW[co][(r*S+s)*Cin + ci] = sum over output pixels of dY[pix][co] * X[pix shifted by tap][ci].
//
// BF16 path (tcgen05): a GEMM whose reduction axis is the PIXEL axis.  Both operands are NHWC, i.e. their GEMM
// M / N index (channels) is the contiguous one, so the shared-memory tiles are MN-major: a TMA box
// {64 channels, TW, TH, 1} of a TH x TW = 64-pixel patch lands as [64 pixel rows][128 B], which is exactly the
// 128B-swizzled MN-major UMMA layout with K = pixel rows.  For filter tap (r,s) the X patch is the dY patch
// shifted by (r*dil - pad, s*dil - pad); TMA zero-fills padding and ragged edges for both operands.
// Work item = (128-wide Cout tile, <=256-wide Cin tile, tap, pixel split); FP32 accumulators in TMEM; the
// epilogue adds the partial tile into the packed FP32 gradient [Cout_pad][Kpad] with red.global.add.f32
// (split-K over pixels keeps all 148 SMs busy even for layers with few channel tiles).
// Strided / small-Cin convolutions run the same kernel in flat mode on the im2col workspace (K columns = taps*Cin).
//
// FP32 path: CUDA-core implicit GEMM with the same decomposition (parity mode).
#include <stdlib.h>
#include <string.h>

#include "hn_common.cuh"
#include "hn_tc_ptx.cuh"

namespace hn {

constexpr int WG_THREADS = 256;     // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4-7 epilogue
constexpr int WG_KPIX = 64;         // pixels per k-block
constexpr int WG_BLOCK_BYTES = WG_KPIX * 128;   // one [64 px][64 ch] block = 8 KB
constexpr int WG_M = 128;

struct WgParams {
    int tiles_w, tiles_h, n_img;    // pixel patches: n_img * tiles_h * tiles_w k-blocks in total
    int TH, TW;                     // TH*TW = 64
    int taps, S, pad, dil;          // taps = R*S (1 in flat mode)
    int stride;                     // conv stride (1 or 2): X patch of tap (r,s) starts at (dY patch origin) * stride + tap * dil - pad
    int m_tiles;                    // Cout / 128 tiles
    int cblocks;                    // 64-column blocks per tap (Cin / 64, or Kpad / 64 in flat mode)
    int blocks_total;               // taps * cblocks: the packed gradient row is blocks_total * 64 columns, block = tap * cblocks + cb
    int nb_item, n_col_items;       // an item covers nb_item (<= BN / 64) consecutive blocks -- possibly of DIFFERENT taps
    int splits;                     // pixel splits per (m, column item)
    int Cout;                       // valid rows
    int kpad;                       // row stride of the packed gradient
    float *dw;                      // [Cout_pad][kpad] FP32, accumulated into
};

// PAIR: two CTAs of a cluster take two adjacent Cout tiles of the same (column item, pixel split) and run ONE M = 256 instruction
// stream (hn_tc_ptx.cuh, "CTA pairs"): each stages its own dY tile and HALF of the X blocks (every item has an even block count),
// so an SM moves 32 KB instead of 48 KB per k-block from L2 into shared memory and the ring is 6 deep instead of 4.
template <int BN, int STAGES, bool PAIR = false>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WgParams p)
{
    constexpr int A_BYTES = 2 * WG_BLOCK_BYTES;              // 128 couts = 2 blocks
    constexpr int B_BLOCKS = PAIR ? BN / 128 : BN / 64;       // X blocks staged by this CTA
    constexpr int B_BYTES = B_BLOCKS * WG_BLOCK_BYTES;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int ACC_COLS = BN < 32 ? 32 : BN;
    constexpr int TMEM_COLS = 2 * ACC_COLS;          // two accumulators: the epilogue (red.global.add) of item i overlaps the main loop of item i+1

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *full_bar = bars, *empty_bar = bars + STAGES, *tfull_bar = bars + 2 * STAGES, *tempty_bar = bars + 2 * STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);
    float *tr_buf = reinterpret_cast<float *>(tmem_slot + 4);          // 4 epilogue warps x [32][36] floats (16-byte aligned)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = p.n_img * p.tiles_h * p.tiles_w;
    const int m_units = PAIR ? (p.m_tiles + 1) / 2 : p.m_tiles;          // Cout tiles, or pairs of them (an odd count leaves a filler
                                                                         // tile beyond Cout: its dY loads are zero-filled, its rows masked)
    const int num_items = m_units * p.n_col_items * p.splits;
    const int item0 = PAIR ? blockIdx.x >> 1 : blockIdx.x, item_step = PAIR ? gridDim.x >> 1 : gridDim.x;   // both CTAs of a pair walk the same items

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_dy);
        prefetch_tmap(&tmap_x);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(smem_u32(full_bar + i), 1);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(tfull_bar + i), 1);
            mbar_init(smem_u32(tempty_bar + i), PAIR ? 8 : 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) { tmem_alloc_pair(smem_u32(tmem_slot), TMEM_COLS); tmem_relinquish_pair(); }
        else { tmem_alloc(smem_u32(tmem_slot), TMEM_COLS); tmem_relinquish(); }
    }
    tcgen05_fence_before();
    if constexpr (PAIR) cluster_sync_all();
    else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work item -> (split, column item, mt), split fastest so concurrently running CTAs stream disjoint pixel ranges.
    // Column blocks are 64 gradient columns = (tap, 64 input channels).  Because the N index of the MMA only has to be a stack of
    // 64-wide MN-major blocks LBO bytes apart, ONE MMA can span blocks of different taps: for Cin = 64 a 3x3 filter is 9 blocks
    // and an item takes 3 of them (N = 192) instead of running 9 separate N = 64 passes at 39 % of the tensor pipe, and the
    // dY tile is fetched once per 3 taps.
    auto decode = [&](int item, int &mt, int &blk0, int &nb, int &kb0, int &kb1) {
        const int sp = item % p.splits;
        const int rest = item / p.splits;
        const int ci = rest % p.n_col_items;
        mt = rest / p.n_col_items;
        if (PAIR) mt = 2 * mt + (int)rank;
        blk0 = ci * p.nb_item;
        nb = p.blocks_total - blk0 < p.nb_item ? p.blocks_total - blk0 : p.nb_item;
        const int per = (num_kb + p.splits - 1) / p.splits;
        kb0 = sp * per;
        kb1 = kb0 + per < num_kb ? kb0 + per : num_kb;
    };

    if (warp == 0) {
        {   // whole warp runs the loop, the elected lane issues (see elect_one)
            int stage = 0;
            uint32_t phase = 0;
            for (int item = item0; item < num_items; item += item_step) {
                int mt, blk0, nb, kb0, kb1;
                decode(item, mt, blk0, nb, kb0, kb1);
                int bcol[BN / 64], bdx[BN / 64], bdy[BN / 64];      // per block: channel offset and tap shift
#pragma unroll
                for (int b = 0; b < BN / 64; ++b) {
                    const int blk = blk0 + (b < nb ? b : 0);
                    const int tap = blk / p.cblocks, cb = blk - tap * p.cblocks;
                    const int r = tap / p.S, sx = tap - r * p.S;
                    bcol[b] = cb * 64;
                    bdx[b] = sx * p.dil - p.pad;
                    bdy[b] = r * p.dil - p.pad;
                }
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int tw = kb % p.tiles_w, th = (kb / p.tiles_w) % p.tiles_h, img = kb / (p.tiles_w * p.tiles_h);
                    mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
                    const uint32_t fb = smem_u32(full_bar + stage);
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    if (elect_one()) {
                        if constexpr (PAIR) {
                            // this CTA's dY tile and its half of the X blocks [rank * nb/2, +nb/2), counted on the leader's barrier
                            const uint32_t lfb = mapa_u32(fb, 0);
                            const int hb = nb >> 1;
                            if (rank == 0) mbar_expect_tx(fb, 2 * (2 + hb) * WG_BLOCK_BYTES);
#pragma unroll
                            for (int b = 0; b < 2; ++b)
                                tma_load_4d_pair(sa + b * WG_BLOCK_BYTES, &tmap_dy, lfb, mt * WG_M + b * 64, tw * p.TW, th * p.TH, img);
#pragma unroll
                            for (int b = 0; b < B_BLOCKS; ++b)
                                if (b < hb) {
                                    const int g = (int)rank * hb + b;
                                    tma_load_4d_pair(sa + A_BYTES + b * WG_BLOCK_BYTES, &tmap_x, lfb, bcol[g], tw * p.TW * p.stride + bdx[g],
                                                     th * p.TH * p.stride + bdy[g], img);
                                }
                        } else {
                            mbar_expect_tx(fb, (2 + nb) * WG_BLOCK_BYTES);
#pragma unroll
                            for (int b = 0; b < 2; ++b)
                                tma_load_4d(sa + b * WG_BLOCK_BYTES, &tmap_dy, fb, mt * WG_M + b * 64, tw * p.TW, th * p.TH, img);
#pragma unroll
                            for (int b = 0; b < BN / 64; ++b)
                                if (b < nb) tma_load_4d(sa + A_BYTES + b * WG_BLOCK_BYTES, &tmap_x, fb, bcol[b], tw * p.TW * p.stride + bdx[b], th * p.TH * p.stride + bdy[b], img);
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        {
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            int acc = 0;
            for (int item = item0; item < num_items; item += item_step) {
                int mt, blk0, nb, kb0, kb1;
                decode(item, mt, blk0, nb, kb0, kb1);
                const uint32_t idesc = make_idesc_bf16_mn(PAIR ? 2 * WG_M : WG_M, 64 * nb);
                mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(smem_u32(full_bar + stage), phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t adesc = make_mnmajor_sw128_desc(sa, WG_BLOCK_BYTES);
                    const uint64_t bdesc = make_mnmajor_sw128_desc(sa + A_BYTES, WG_BLOCK_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < WG_KPIX / 16; ++k) {
                            // 16 pixel rows = 2048 B further along K: +128 in the (addr >> 4) field
                            umma_bf16_t<PAIR>(d_tmem, adesc + 128 * k, bdesc + 128 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit_t<PAIR>(smem_u32(empty_bar + stage));
                        if (kb == kb1 - 1) umma_commit_t<PAIR>(smem_u32(tfull_bar + acc));
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (kb1 <= kb0 && elect_one()) umma_commit_t<PAIR>(smem_u32(tfull_bar + acc));     // empty split: nothing was issued
                __syncwarp();
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        uint32_t acc_phase = 0;
        int acc = 0;
        for (int item = item0; item < num_items; item += item_step) {
            int mt, blk0, nb, kb0, kb1;
            decode(item, mt, blk0, nb, kb0, kb1);
            mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS;
            // Thread = one Cout row of the accumulator.  Adding it straight to the packed gradient would touch 32 rows (32
            // sectors) per warp instruction; instead every 32x32 block goes through a padded shared-memory tile so that one
            // red.global.add.v4.f32 per lane covers 4 rows x 128 contiguous bytes (8x fewer L2 sectors, 4x fewer instructions).
            const uint32_t tbuf = smem_u32(tr_buf + q * (32 * 36));
            const int64_t colbase = (int64_t)blk0 * 64;
            constexpr int CH = 32;
#pragma unroll 1
            for (int c0 = 0; c0 < 64 * nb; c0 += CH) {
                uint32_t raw[CH];
                tmem_ld_32x32(taddr + c0, raw);
                tmem_ld_wait();
                if (kb1 <= kb0) continue;                                    // warp-uniform
#pragma unroll
                for (int j = 0; j < CH / 4; ++j)
                    sts128(tbuf + lane * 144 + j * 16, make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]));
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = 4 * i + (lane >> 3), c4 = (lane & 7) * 4;
                    if (c4 < CH) {
                        const uint4 v = lds128(tbuf + r * 144 + c4 * 4);
                        const int cor = mt * WG_M + q * 32 + r;
                        if (cor < p.Cout) {
                            float *dst = p.dw + (int64_t)cor * p.kpad + colbase + c0 + c4;
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(v.x)),
                                         "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                                         : "memory");
                        }
                    }
                }
                __syncwarp();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(tempty_bar + acc), 0));
                else mbar_arrive(smem_u32(tempty_bar + acc));
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    tcgen05_fence_before();
    if constexpr (PAIR) cluster_sync_all();
    else __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int BN, int STAGES, bool PAIR = false>
static int launch_wgrad(const CUtensorMap &tdy, const CUtensorMap &tx, const WgParams &p, cudaStream_t st)
{
    constexpr size_t smem = (size_t)STAGES * (2 + (PAIR ? BN / 128 : BN / 64)) * WG_BLOCK_BYTES + (2 * STAGES + 4) * 8 + 16 + 4 * 32 * 36 * 4 + 1024;
    static_assert(smem <= 227 * 1024, "wgrad shared memory");
    static bool configured_dev[HN_MAX_DEVICES] = {};
    static int pairs_dev[HN_MAX_DEVICES] = {};
    bool &configured = configured_dev[current_device()];
    int &pairs = pairs_dev[current_device()];
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(WG_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (!configured) {
        HN_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<BN, STAGES, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (PAIR) {
            cfg.gridDim = dim3(2, 1, 1);
            if (cudaOccupancyMaxActiveClusters(&pairs, wgrad_tc_kernel<BN, STAGES, PAIR>, &cfg) != cudaSuccess || pairs < 1) {
                cudaGetLastError();
                pairs = num_sms() / 2;
            }
        }
        configured = true;
    }
    if constexpr (PAIR) {
        const int items = ((p.m_tiles + 1) / 2) * p.n_col_items * p.splits;
        cfg.gridDim = dim3(2 * (unsigned)(items < pairs ? items : pairs), 1, 1);
        cfg.stream = st;
        HN_CUDA(cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<BN, STAGES, PAIR>, tdy, tx, p));
    } else {
        const int items = p.m_tiles * p.n_col_items * p.splits;
        int grid = items < num_sms() ? items : num_sms();
        wgrad_tc_kernel<BN, STAGES, PAIR><<<grid, WG_THREADS, smem, st>>>(tdy, tx, p);
    }
    HN_LAUNCH_CHECK();
    return HN_OK;
}

// im2col gather shared with the forward engine (hn_conv_tc.cu)
int im2col_bf16(const hn_tensor *x, const hn_conv *cv, int Ho, int Wo, int kpad, void *ws, cudaStream_t st);

int conv2d_wgrad_tc(const hn_tensor *x, const hn_tensor *dy, const hn_conv *cv, float *dw_packed, void *ws, int64_t ws_bytes,
                    cudaStream_t st)
{
    const int kpad = hn_conv_kpad(x->c, cv->r, cv->s);
    const int64_t M = (int64_t)dy->n * dy->h * dy->w;
    if (M == 0) return HN_OK;
    HN_CHECK_ARG((reinterpret_cast<uintptr_t>(dy->ptr) & 15) == 0 && dy->ld % 8 == 0, "conv_wgrad: dY view must be 16-byte aligned");
    const bool implicit = conv_implicit_ok(x, cv);      // stride 2: the X map traverses the pixel axes with element stride 2
    WgParams p{};
    p.stride = 1;
    CUtensorMap tdy, tx;
    p.Cout = cv->cout;
    p.kpad = kpad;
    p.dw = dw_packed;
    p.m_tiles = (int)cdiv(cv->cout, WG_M);
    int ncols;
    if (implicit) {
        const int menu[4][2] = {{8, 8}, {4, 16}, {2, 32}, {1, 64}};
        int best = 0;
        int64_t best_area = -1;
        for (int i = 0; i < 4; ++i) {
            int64_t area = cdiv(dy->h, menu[i][0]) * menu[i][0] * cdiv(dy->w, menu[i][1]) * menu[i][1];
            if (best_area < 0 || area < best_area) { best_area = area; best = i; }
        }
        p.TH = menu[best][0]; p.TW = menu[best][1];
        p.tiles_h = (int)cdiv(dy->h, p.TH); p.tiles_w = (int)cdiv(dy->w, p.TW); p.n_img = dy->n;
        p.taps = cv->r * cv->s; p.S = cv->s; p.pad = cv->pad; p.dil = cv->dil; p.stride = cv->stride;
        ncols = x->c;
        uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
        uint64_t dd[4] = {(uint64_t)dy->c, (uint64_t)dy->w, (uint64_t)dy->h, (uint64_t)dy->n};
        uint64_t ds[4] = {2, (uint64_t)dy->ld * 2, (uint64_t)dy->ld * 2 * dy->w, (uint64_t)dy->ld * 2 * dy->w * dy->h};
        int rc = make_tmap(&tdy, dy->ptr, 4, dd, ds, box);
        if (rc) return rc;
        uint64_t xd[4] = {(uint64_t)x->c, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
        uint64_t xs[4] = {2, (uint64_t)x->ld * 2, (uint64_t)x->ld * 2 * x->w, (uint64_t)x->ld * 2 * x->w * x->h};
        const uint32_t st_ = (uint32_t)cv->stride;
        uint32_t xbox[4] = {64, (uint32_t)p.TW * st_, (uint32_t)p.TH * st_, 1};
        uint32_t es[4] = {1, st_, st_, 1};
        rc = make_tmap(&tx, x->ptr, 4, xd, xs, xbox, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B, es);
        if (rc) return rc;
    } else {
        const int64_t need = M * kpad * 2;
        if (!ws || ws_bytes < need) {
            set_error("conv_wgrad: workspace of %lld bytes required, %lld given", (long long)need, (long long)ws_bytes);
            return HN_ERR_WORKSPACE;
        }
        int rc = im2col_bf16(x, cv, dy->h, dy->w, kpad, ws, st);
        if (rc) return rc;
        HN_CHECK_ARG(M < ((int64_t)1 << 31), "conv_wgrad: too many pixels");
        p.TH = 1; p.TW = 64;
        p.tiles_h = 1; p.tiles_w = (int)cdiv(M, 64); p.n_img = 1;
        p.taps = 1; p.S = 1; p.pad = 0; p.dil = 1;
        ncols = kpad;
        uint32_t box[4] = {64, 64, 1, 1};
        uint64_t dd[4] = {(uint64_t)dy->c, (uint64_t)M, 1, 1};
        uint64_t ds[4] = {2, (uint64_t)dy->ld * 2, (uint64_t)dy->ld * 2, (uint64_t)dy->ld * 2};
        rc = make_tmap(&tdy, dy->ptr, 4, dd, ds, box);
        if (rc) return rc;
        uint64_t xd[4] = {(uint64_t)kpad, (uint64_t)M, 1, 1};
        uint64_t xs[4] = {2, (uint64_t)kpad * 2, (uint64_t)kpad * 2, (uint64_t)kpad * 2};
        rc = make_tmap(&tx, ws, 4, xd, xs, box);
        if (rc) return rc;
    }
    HN_CHECK_ARG(ncols % 64 == 0, "conv_wgrad: column count must be a multiple of 64");
    p.cblocks = ncols / 64;
    p.blocks_total = p.taps * p.cblocks;
    p.n_col_items = (int)cdiv(p.blocks_total, 4);
    p.nb_item = (int)cdiv(p.blocks_total, p.n_col_items);            // 1..4 blocks (N = 64..256) per item, evenly spread
    const int bn = p.nb_item >= 3 ? 256 : (p.nb_item == 2 ? 128 : 64);
    const int num_kb = p.n_img * p.tiles_h * p.tiles_w;
    // CTA pairs: wide items only (N = 256: four X blocks), every item with an even block count, at least two Cout tiles
    static const bool no_pair = getenv("HN_NO_PAIR") != nullptr || getenv("HN_NO_PAIR_WGRAD") != nullptr;
    const int last_nb = p.blocks_total - (p.n_col_items - 1) * p.nb_item;
    const bool pair = !no_pair && bn == 256 && p.nb_item % 2 == 0 && last_nb % 2 == 0 && p.m_tiles >= 2;
    const int base_items = (pair ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_col_items;
    int splits = (int)cdiv((pair ? 1 : 2) * (int64_t)num_sms(), base_items);          // ~2 items per SM (a pair item occupies two)
    if (splits > num_kb) splits = num_kb;
    if (splits < 1) splits = 1;
    // make every split non-empty
    const int per = (int)cdiv(num_kb, splits);
    splits = (int)cdiv(num_kb, per);
    p.splits = splits;
    if (pair) return launch_wgrad<256, 6, true>(tdy, tx, p, st);
    switch (bn) {
        case 256: return launch_wgrad<256, 4>(tdy, tx, p, st);
        case 128: return launch_wgrad<128, 6>(tdy, tx, p, st);
        default: return launch_wgrad<64, 8>(tdy, tx, p, st);
    }
}

// ------------------------------------------------------------------------------------------------ FP32 CUDA-core wgrad
constexpr int WF_BM = 64, WF_BN = 64, WF_BK = 16;   // cout tile, k-column tile, pixels per step

__global__ void __launch_bounds__(256) wgrad_f32_kernel(const float *__restrict__ x, int ldx, const float *__restrict__ dy, int ldy, int N,
                                                        int H, int W, int C, int Ho, int Wo, int Cout, int R, int S, int stride, int pad,
                                                        int dil, int kpad, int64_t pix_per_split, float *__restrict__ dw)
{
    __shared__ float As[WF_BK][WF_BM + 4];   // dY[pix][co]
    __shared__ float Bs[WF_BK][WF_BN + 4];   // im2col(X)[pix][k]
    const int tid = threadIdx.x;
    const int co0 = blockIdx.x * WF_BM, k0 = blockIdx.y * WF_BN;
    const int K = R * S * C;
    const int64_t M = (int64_t)N * Ho * Wo;
    const int64_t p0 = (int64_t)blockIdx.z * pix_per_split;
    const int64_t p1 = p0 + pix_per_split < M ? p0 + pix_per_split : M;
    // loader: pixel lp = tid / 16 (0..15), 4 consecutive columns lc = (tid % 16) * 4
    const int lp = tid >> 4, lc = (tid & 15) * 4;
    // the 4 k-columns this thread gathers are fixed for the whole kernel: decode taps once
    int kr[4], ks[4], kc[4];
    bool kv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int k = k0 + lc + j;
        kv[j] = k < K;
        int tap = kv[j] ? k / C : 0;
        kc[j] = kv[j] ? k - tap * C : 0;
        kr[j] = tap / S;
        ks[j] = tap - kr[j] * S;
    }
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int64_t pb = p0; pb < p1; pb += WF_BK) {
        const int64_t pix = pb + lp;
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (pix < p1) {
            const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((int64_t)Wo * Ho));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (co0 + lc + j < Cout) av[j] = dy[pix * ldy + co0 + lc + j];
                if (kv[j]) {
                    int hi = ho * stride - pad + kr[j] * dil, wi = wo * stride - pad + ks[j] * dil;
                    if (hi >= 0 && hi < H && wi >= 0 && wi < W) bv[j] = x[(((int64_t)n * H + hi) * W + wi) * ldx + kc[j]];
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            As[lp][lc + j] = av[j];
            Bs[lp][lc + j] = bv[j];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < WF_BK; ++kk) {
            float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int co = co0 + ty * 4 + i;
        if (co >= Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + tx * 4 + j;
            if (k < K) atomicAdd(dw + (int64_t)co * kpad + k, acc[i][j]);
        }
    }
}

int conv2d_wgrad_f32(const hn_tensor *x, const hn_tensor *dy, const hn_conv *cv, float *dw_packed, cudaStream_t st)
{
    const int kpad = hn_conv_kpad(x->c, cv->r, cv->s);
    const int K = cv->r * cv->s * x->c;
    const int64_t M = (int64_t)dy->n * dy->h * dy->w;
    if (M == 0) return HN_OK;
    const int gx = (int)cdiv(cv->cout, WF_BM), gy = (int)cdiv(K, WF_BN);
    int64_t splits = cdiv(4 * (int64_t)num_sms(), (int64_t)gx * gy);
    int64_t pps = cdiv(M, splits);
    pps = cdiv(pps, WF_BK) * WF_BK;
    if (pps < 256) pps = 256;
    splits = cdiv(M, pps);
    if (splits > 65535) { pps = cdiv(cdiv(M, 65535), WF_BK) * WF_BK; splits = cdiv(M, pps); }
    dim3 grid(gx, gy, (unsigned)splits);
    wgrad_f32_kernel<<<grid, 256, 0, st>>>((const float *)x->ptr, x->ld, (const float *)dy->ptr, dy->ld, x->n, x->h, x->w, x->c, dy->h, dy->w,
                                          cv->cout, cv->r, cv->s, cv->stride, cv->pad, cv->dil, kpad, pps, dw_packed);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

}  // namespace hn

using namespace hn;

static int conv_out_dim_w(int in, int k, int stride, int pad, int dil) { return (in + 2 * pad - dil * (k - 1) - 1) / stride + 1; }

extern "C" int64_t hn_conv2d_wgrad_workspace_bytes(const hn_tensor *x, const hn_conv *cv)
{
    if (!x || !cv || x->dtype != HN_BF16) return 0;
    if (conv_implicit_ok(x, cv)) return 0;
    const int Ho = conv_out_dim_w(x->h, cv->r, cv->stride, cv->pad, cv->dil), Wo = conv_out_dim_w(x->w, cv->s, cv->stride, cv->pad, cv->dil);
    return (int64_t)x->n * Ho * Wo * hn_conv_kpad(x->c, cv->r, cv->s) * 2;
}

extern "C" int hn_conv2d_wgrad(const hn_tensor *x, const hn_tensor *dy, const hn_conv *cv, float *dw_packed, int32_t zero_init,
                               void *workspace, int64_t workspace_bytes, void *stream)
{
    HN_CHECK_ARG(x && dy && cv && dw_packed && x->ptr && dy->ptr, "hn_conv2d_wgrad: null pointer");
    HN_CHECK_ARG(x->dtype == dy->dtype, "hn_conv2d_wgrad: x and dY must share a dtype");
    const int Ho = conv_out_dim_w(x->h, cv->r, cv->stride, cv->pad, cv->dil), Wo = conv_out_dim_w(x->w, cv->s, cv->stride, cv->pad, cv->dil);
    HN_CHECK_ARG(dy->n == x->n && dy->h == Ho && dy->w == Wo && dy->c == cv->cout, "hn_conv2d_wgrad: dY must be N=%d %dx%d C=%d", x->n, Ho, Wo,
                 cv->cout);
    cudaStream_t st = (cudaStream_t)stream;
    const int kpad = hn_conv_kpad(x->c, cv->r, cv->s);
    const int cout_pad = hn_conv_cout_pad(cv->cout, x->dtype);
    if (zero_init) HN_CUDA(cudaMemsetAsync(dw_packed, 0, sizeof(float) * (size_t)cout_pad * kpad, st));
    if (x->dtype == HN_F32) return conv2d_wgrad_f32(x, dy, cv, dw_packed, st);
    return conv2d_wgrad_tc(x, dy, cv, dw_packed, workspace, workspace_bytes, st);
}
