// Shared epilogue of the tcgen05 convolution kernels (hn_conv_tc.cu, hn_conv_halo.cu).
#pragma once
#include "hn_common.cuh"
#include "hn_tc_ptx.cuh"

namespace hn {

constexpr int EPI_WARP0 = 4;

// Classifier weights travel as KERNEL PARAMETERS (constant bank): every FMA of the head takes its weight as a constant-bank
// operand (LDC / uniform registers) -- no shared-memory or global loads in the 13 x 64 FMA loop of each pixel.
constexpr int HEAD_MAX = 16;
struct HeadConst {
    float w[HEAD_MAX][64];
    float b[HEAD_MAX];
};

struct TcParams {
    // tiling of the output pixel space
    int tiles_w, tiles_h, n_img;   // M tiles = n_img * tiles_h * tiles_w
    int TH, TW;                    // TH*TW = 128
    int Ho, Wo;                    // output spatial size (flat mode: Ho = 1, Wo = total pixels)
    int n_tiles;                   // Cout tiles
    int R, S, pad, dil;            // pad = rows; pad_w = columns (equal except for the phases of a strided dgrad)
    int pad_w;
    int hmul, wmul;                // input row / column of tap (r, s) = (output row / column) * hmul / wmul - pad + tap * dil: the conv stride
                                   // (hmul = 2, wmul = 1 for the stem's window map, whose column stride is folded into the map)
    int cblocks;                   // Cin / 64 (flat-from-workspace: kpad / 64 with R=S=1)
    int Cout;
    // epilogue
    void *y;
    int ldy, y_f32;
    const float *scale, *shift;
    const void *res;
    int ldr;
    int act;
    float slope;
    const float *slope_ptr;
    // TMA epilogue: output (and residual) go through a swizzled shared-memory staging tile per epilogue warp
    int tma_out, tma_res;
    int ebw;                       // staging box = {128 B of channels, ebw pixels, 32/ebw rows, 1}: ebw = min(TW, 32)
    // train-mode BatchNorm2d statistics fused into the FP32-output epilogue: per-channel sum / sum of squares of the STORED
    // pre-normalisation values over the valid pixels, FP32 per 32-row block, FP64 atomics into [Cout] accumulators
    double *stat_sum, *stat_sqsum;
    // image groups with SEPARATE statistics (the day and the night batch of a training step run through the convolutions as
    // one batch, BatchNorm2d keeps normalising each domain by itself): accumulators are [groups][Cout]; a 32-pixel block belongs
    // to group (image / stat_group_img) -- flat mode: (first pixel / stat_group_pix), stat_group_pix % 32 == 0.  0 = one group.
    int stat_group_img, stat_group_pix;
    // fused 1x1 classifier head (halo kernel, BLOCK_N = 64 = Cout): logits written as NCHW FP32, the activation is not stored
    float *head_out;               // [n_img][head_n][Ho][Wo]
    int head_n;                    // <= HEAD_MAX; 0 = no head
    HeadConst head;                // classifier weights (rows >= head_n are zero)
};

// host: statistics groups of a launch whose tiling is already in p (n_img / Ho / Wo; flat mode has n_img == 1)
inline int set_stat_groups(TcParams &p, const hn_epilogue *ep, const hn_tensor *y)
{
    p.stat_group_img = p.stat_group_pix = 0;
    if (!ep->stat_sum || ep->stat_groups <= 1) return HN_OK;
    const int g = ep->stat_groups;
    if (y->n % g != 0) {
        set_error("conv: %d images do not split into %d statistics groups", y->n, g);
        return HN_ERR_ARG;
    }
    if (p.n_img == y->n) p.stat_group_img = y->n / g;
    else {
        const int64_t gp = (int64_t)(y->n / g) * y->h * y->w;
        if (gp % 32 != 0 || gp >= ((int64_t)1 << 31)) {
            set_error("conv: flattened statistics groups need (N / G) * H * W to be a multiple of 32 (got %lld)", (long long)gp);
            return HN_ERR_ARG;
        }
        p.stat_group_pix = (int)gp;
    }
    return HN_OK;
}

constexpr int EPI_STAGE_BYTES = 32 * 128;   // 32 pixels x 128 B per epilogue warp
constexpr int NUM_EPI_WARPS = 8;


__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// phase B of the fused classifier: wait for the head MMA of a tile, read this thread's pixel (TMEM lane) back and write its
// logits as NCHW FP32 (streaming stores: the logits are not re-read by this kernel)
__device__ __forceinline__ void head_emit(const TcParams &p, uint32_t taddr, uint32_t bar, uint32_t phase, int img, int ho, int wo, bool valid)
{
    mbar_wait(bar, phase);
    tcgen05_fence_after();
    uint32_t raw[16];
    tmem_ld_32x16(taddr, raw);
    tmem_ld_wait();
    tcgen05_fence_before();
    if (valid) {
        const int64_t plane = (int64_t)p.Ho * p.Wo;
        float *o = p.head_out + (int64_t)img * p.head_n * plane + (int64_t)ho * p.Wo + wo;
#pragma unroll
        for (int k = 0; k < HEAD_MAX; ++k)
            if (k < p.head_n) __stcs(o + k * plane, __uint_as_float(raw[k]) + p.head.b[k]);
    }
}

// Runs on warps EPI_WARP0 .. EPI_WARP0+7 of a CTA.  This CTA walks tile = blockIdx.x, += gridDim.x, ... < num_tiles.
//
// Warp w may only touch TMEM lanes [32*(w%4), +32): one accumulator row (= output pixel) per thread.  Wide tiles
// (>= 128 columns) are split between two groups of 4 warps.  Per 128-byte chunk of a pixel's channels the warp
// (optionally) TMA-loads the residual chunk into its private staging tile, tcgen05.ld's the accumulator, adds the
// per-channel shift, packs to BF16, adds the residual and applies ReLU on packed BF16x2 (the same two roundings a BF16
// torch pipeline performs), writes the row back into the (128B-swizzled, bank-conflict-free) staging tile and one lane
// TMA-stores the 32-pixel x 128-byte box.  Global traffic is full-line bulk transfers; the LSU only sees shared memory.
//
// The epilogue is INSTRUCTION-ISSUE bound on the output-heavy layers (ncu: ~340 SASS instructions per 32 columns in the
// first version, 60 % of all instructions of the kernel), so the hot path is kept minimal: the BN scale is folded into
// the weights on the host (p.scale == NULL), the shift comes from a shared-memory table via 128-bit broadcast loads (with
// ~224 KB of the SM carved out as shared memory there is next to no L1: per-element __ldg costs an L2 round trip),
// residual add / ReLU run on packed pairs, swizzled staging offsets are precomputed.  Anything else (FP32 outputs,
// unaligned views, explicit scale, LeakyReLU + residual) takes the general path below.
// CTA pairs (PAIR): `tile` walks an extended index space -- tile >> 1 is the (M-tile pair, Cout tile) the pair works on, tile & 1
// (= this CTA's cluster rank) the M tile of the pair; an odd M-tile count leaves the last pair's second CTA a tile beyond the last
// image, whose loads are zero-filled and whose stores are clipped by the TMA unit.
template <bool PAIR>
__device__ __forceinline__ void decode_tile(const TcParams &p, int tile, int &nt, int &mt)
{
    if constexpr (PAIR) {
        const int pt = tile >> 1;
        nt = pt % p.n_tiles;
        mt = (pt / p.n_tiles) * 2 + (tile & 1);
    } else {
        nt = tile % p.n_tiles;
        mt = tile / p.n_tiles;
    }
}
// size of that index space
__host__ __device__ inline int pair_num_tiles(int num_m_tiles, int n_tiles) { return 2 * ((num_m_tiles + 1) / 2) * n_tiles; }

// NBUF: accumulator buffers of ACC_COLS TMEM columns each; this CTA's it-th tile uses buffer it % NBUF.
// RES2: every epilogue warp owns TWO staging tiles (the second set starts NUM_EPI_WARPS * EPI_STAGE_BYTES behind the first) and two
// residual barriers (res_bar[NUM_EPI_WARPS + ew]): chunks alternate between them, the residual of chunk i+1 is fetched while chunk i
// is computed (with one tile it could only be requested after chunk i's store had drained the tile: ~2300 cycles of load latency per
// 128 x 256 tile sat exposed in the residual 1x1 layers), and a store may still be reading one tile while the next chunk fills the other.
template <int BLOCK_N, bool ONE_GROUP = false, int NACC = 1, bool PAIR = false, int NBUF = 2, bool RES2 = false>
__device__ __forceinline__ void conv_epilogue(const TcParams &p, const CUtensorMap *tmap_y_p, const CUtensorMap *tmap_r_p, uint32_t tmem_base,
                                              uint64_t *tfull_bar, uint64_t *tempty_bar, uint64_t *res_bar, uint8_t *epi_stage,
                                              float *s_shift, int num_tiles, int warp, int lane)
{
    // NACC = 2: the MMA warp spreads a tile's MMAs over two accumulators (independent accumulation chains: back-to-back MMAs
    // into ONE accumulator serialise on the tensor pipe's latency, ~100 cycles each, which bounds narrow N = 64 tiles); the
    // epilogue adds the two halves.  ACC_COLS = TMEM columns per tile buffer.
    constexpr int SUB_ACC = BLOCK_N < 32 ? 32 : BLOCK_N;
    constexpr int ACC_COLS = NACC * SUB_ACC;
    const CUtensorMap &tmap_y = *tmap_y_p;
    const CUtensorMap &tmap_r = *tmap_r_p;
    constexpr int GROUPS = (BLOCK_N >= 128 && !ONE_GROUP) ? 2 : 1;   // ONE_GROUP: warps 8-11 have another job
    constexpr int COLS = BLOCK_N / GROUPS;                 // columns per warp
    constexpr int SUB = COLS < 32 ? COLS : 32;             // columns per tcgen05.ld
    const int ew = warp - EPI_WARP0;
    const int q = ew & 3, grp = ew >> 2;
    static_assert(NBUF % 2 == 0, "the alternating epilogue groups need an even number of accumulator buffers");
    static_assert(!RES2 || (BLOCK_N >= 128 && !ONE_GROUP), "double-buffered staging is only wired up for the wide-tile epilogue");
    // narrow tiles (one column group): the two groups of 4 warps ALTERNATE tiles -- group g owns the accumulator buffers = g (mod 2), i.e.
    // every other tile of this CTA -- so two tiles are in the epilogue at once (measured: the N = 64 layers were epilogue-bound)
    constexpr bool ALT = GROUPS == 1 && !ONE_GROUP;
    if (!ALT && grp >= GROUPS) return;
    const int row = q * 32 + lane;                     // accumulator row = pixel within the tile
    const int col0 = ALT ? 0 : grp * COLS;
    const int esz = p.y_f32 ? 4 : 2;
    const int tch = 128 / esz;                         // columns per staging chunk (128 B per pixel)
    const uint32_t stage = smem_u32(epi_stage + ew * EPI_STAGE_BYTES);
    const uint32_t srow = stage + lane * 128;          // this thread's pixel row in the staging tile
    uint32_t off[8];                                   // swizzled address of each 16-byte chunk of that row
#pragma unroll
    for (int c = 0; c < 8; ++c) off[c] = srow + ((c ^ (lane & 7)) << 4);
    const uint32_t rbar = smem_u32(res_bar + ew);
    uint32_t rphase = 0;
    constexpr uint32_t STAGE2_DELTA = NUM_EPI_WARPS * EPI_STAGE_BYTES;      // second staging tile of this warp (RES2)
    const uint32_t rbar2 = smem_u32(res_bar + NUM_EPI_WARPS + ew);
    uint32_t rphase2 = 0, cc = 0;                                           // cc: chunks this warp has processed (its parity picks the tile)
    bool primed = false;                                                    // RES2: the current chunk's residual has been requested
    (void)rbar2; (void)rphase2; (void)cc; (void)primed;
    float slope = p.slope;
    if (p.slope_ptr) slope = __ldg(p.slope_ptr);
    const int act = p.act;
    const bool res_vec = p.res != nullptr && (p.ldr & 7) == 0 && (reinterpret_cast<uintptr_t>(p.res) & 15) == 0;
    // packed-BF16 fast path: TMA staging, BF16 output, no explicit scale, residual (if any) staged by TMA
    const bool head = BLOCK_N == 64 && p.head_n > 0;   // host guarantees the fast-path preconditions when a head is given
    const bool tma_out = p.tma_out && !head;
    const bool fast = SUB == 32 && (tma_out || head) && !p.y_f32 && p.scale == nullptr && (p.res == nullptr || p.tma_res) &&
                      !(act == HN_ACT_LEAKY && p.res != nullptr);
    float *tab = s_shift + (ALT ? grp * BLOCK_N : col0);   // this group's per-channel shift table (ALT: one full table per group)
    const uint32_t tab_addr = smem_u32(tab);
    int loaded_ctile = -1;
    long long ew_tfull = 0, ew_bulk = 0, ew_res = 0, ew_ld = 0, ew_st = 0, ew_arr = 0;
    (void)ew_tfull; (void)ew_bulk; (void)ew_res; (void)ew_ld; (void)ew_st; (void)ew_arr;
#ifdef HN_PROFILE_ROLES
    const long long epi_t0 = clock64();
#endif
    int prev_img = 0, prev_ho = 0, prev_wo = 0;
    bool have_prev = false, prev_valid = false;
    uint32_t hph = 0;                                   // phase of this group's head barrier
    const uint32_t head_bar = smem_u32(res_bar + 4 * grp);            // no residual with a head: this group's first res_bar is free
    const uint32_t head_ready = smem_u32(res_bar + 4 * grp + 1);      // pairs: "both CTAs have staged their tile" (initialised with count 2)
    uint32_t rdy_phase = 0;
    (void)head_ready; (void)rdy_phase;
    const uint32_t head_d2 = tmem_base + NBUF * ACC_COLS + grp * HEAD_MAX;   // this group's logits accumulator (16 TMEM columns)
    if constexpr (BLOCK_N == 64) {
        if (head) {
            // classifier weights -> BF16 [16 classes][64 ch] K-major SWIZZLE_128B tile behind the staging buffers (rows >= head_n: 0)
            // (a CTA of a pair holds classes [8 * rank, +8): its half of the N = 16 rows of the pair's B operand)
            const int t = q * 32 + lane, kl = PAIR ? (t >> 3) & 7 : t >> 3, cidx = t & 7;
            const int k = PAIR ? 8 * (int)cluster_ctarank() + kl : kl;
            __nv_bfloat162 h4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) h4[e] = __floats2bfloat162_rn(p.head.w[k][cidx * 8 + 2 * e], p.head.w[k][cidx * 8 + 2 * e + 1]);
            if (grp == 0) {
                if (!PAIR || t < 64)
                    sts128(smem_u32(epi_stage + NUM_EPI_WARPS * EPI_STAGE_BYTES) + kl * 128 + ((cidx ^ (kl & 7)) << 4), *reinterpret_cast<const uint4 *>(h4));
                fence_proxy_async();
            }
            named_bar_sync(3, ALT ? 256 : 128);
        }
    }
    // fused BatchNorm2d statistics (p.stat_sum): per-warp FP32 partial sums of this warp's COLS columns, one per lane and 32-column chunk
    // Two lane <-> channel mappings share the accumulators: FP32 output -- slot i, lane l = channel 32*i + l of the warp's columns;
    // BF16 output (64 channels per 128-byte staging row, a lane reads one BF16 PAIR per row) -- slot 2*j + e, lane l = channel
    // 64*j + 2*l + e.  BF16 statistics are taken from the ROUNDED values, i.e. they are exactly the statistics of the stored tensor.
    constexpr int STAT_CHUNKS = COLS >= 64 ? COLS / 32 : 2;
    float st_s[STAT_CHUNKS], st_q[STAT_CHUNKS];
#pragma unroll
    for (int i = 0; i < STAT_CHUNKS; ++i) st_s[i] = st_q[i] = 0.f;
    int stat_ctile = -1, stat_grp = 0;
    const bool stat_bf16 = p.stat_sum != nullptr && !p.y_f32;
    auto stat_flush = [&]() {
        if (stat_ctile < 0) return;
#pragma unroll
        for (int i = 0; i < STAT_CHUNKS; ++i) {
            const int ch = stat_bf16 ? stat_ctile + 64 * (i >> 1) + 2 * lane + (i & 1) : stat_ctile + 32 * i + lane;
            const bool mine = stat_bf16 ? (64 * (i >> 1) + 2 * lane + (i & 1) < COLS) : (32 * i + lane < COLS);
            if (mine && ch < p.Cout && (st_s[i] != 0.f || st_q[i] != 0.f)) {
                atomicAdd(p.stat_sum + stat_grp * p.Cout + ch, (double)st_s[i]);
                atomicAdd(p.stat_sqsum + stat_grp * p.Cout + ch, (double)st_q[i]);
            }
            st_s[i] = st_q[i] = 0.f;
        }
    };
    // "this accumulator has been read": the MMA issuer lives in the leader CTA of a pair
    const uint32_t pair_rank = PAIR ? cluster_ctarank() : 0;
    const uint32_t tempty_leader = PAIR ? mapa_u32(smem_u32(tempty_bar), 0) : 0;
    (void)pair_rank; (void)tempty_leader;
    auto tempty_arrive = [&](int a) {
        if constexpr (PAIR) {
            if (pair_rank == 0) mbar_arrive(smem_u32(tempty_bar + a));
            else mbar_arrive_cluster(tempty_leader + a * 8);
        } else mbar_arrive(smem_u32(tempty_bar + a));
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        if (ALT && (it & 1) != grp) continue;
        const int acc = it % NBUF;                              // (ALT: the groups own the even / the odd buffers)
        const uint32_t acc_phase = (uint32_t)(it / NBUF) & 1u;
        int nt, mt;
        decode_tile<PAIR>(p, tile, nt, mt);
        const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, img = mt / (p.tiles_w * p.tiles_h);
        const int ho = th * p.TH + row / p.TW, wo = tw * p.TW + row % p.TW;
        const bool valid = ho < p.Ho && wo < p.Wo && (!PAIR || img < p.n_img);
        const int64_t pix = ((int64_t)img * p.Ho + ho) * p.Wo + wo;
        const int ctile = nt * BLOCK_N + col0;         // first output channel of this warp's columns
        if (p.stat_sum) {
            int grp = 0;                             // warp-uniform: the statistics group of this warp's 32-pixel block
            if (p.stat_group_img) grp = img / p.stat_group_img;
            else if (p.stat_group_pix) grp = (tw * p.TW + q * 32) / p.stat_group_pix;
            if (PAIR && img >= p.n_img) grp = stat_grp;          // the filler tile of an odd pair contributes nothing
            if (ctile != stat_ctile || grp != stat_grp) {
                stat_flush();
                stat_ctile = ctile;
                stat_grp = grp;
            }
        }
        // coordinates of this warp's 32-pixel box (first pixel = row q*32 of the tile)
        const int bx = tw * p.TW + (q * 32) % p.TW, by = th * p.TH + (q * 32) / p.TW;
        const __nv_bfloat16 *rrow = (const __nv_bfloat16 *)p.res + pix * p.ldr + ctile;
        if (p.shift && ctile != loaded_ctile) {        // same decision in all 4 warps of the group
            if (loaded_ctile >= 0) named_bar_sync(1 + grp, 128);          // everybody is done with the old table
            if (q == 0)
                for (int e = lane; e < COLS; e += 32) tab[e] = ctile + e < p.Cout ? __ldg(p.shift + ctile + e) : 0.f;
            named_bar_sync(1 + grp, 128);
            loaded_ctile = ctile;
        }
        bool waited = false;
        if constexpr (BLOCK_N == 64) {
            if (head) {
                // ---------------- conv + activation + 1x1 classifier on the tensor core ----------------
                // The activation tile -> BF16 -> this group's swizzled staging buffer, which IS a K-major SWIZZLE_128B
                // [128 pixels][64 ch] UMMA operand; one elected thread issues logits[128][16] = tile x head^T into this group's
                // spare TMEM columns.  The logits of the group's PREVIOUS tile (its MMA had two tile periods to retire) are read
                // back first -- that wait also proves the staging buffer and the logits columns are free again.
                if (have_prev) {
                    head_emit(p, head_d2 + ((uint32_t)(q * 32) << 16), head_bar, hph, prev_img, prev_ho, prev_wo, prev_valid);
                    hph ^= 1;
                }
                { HN_PROF_T0(); mbar_wait(smem_u32(tfull_bar + acc), acc_phase); HN_PROF_ADD(ew_tfull); }
                tcgen05_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS;
#pragma unroll
                for (int si = 0; si < 2; ++si) {
                    uint32_t raw[32];
                    tmem_ld_32x32(taddr + si * 32, raw);
                    float v[32];
                    if constexpr (NACC == 2) {
                        uint32_t raw2[32];
                        tmem_ld_32x32(taddr + SUB_ACC + si * 32, raw2);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) + __uint_as_float(raw2[j]);
                    } else {
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
                    }
                    if (p.shift) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint4 t4 = lds128(tab_addr + (si * 32 + 4 * j) * 4);
                            v[4 * j] += __uint_as_float(t4.x);
                            v[4 * j + 1] += __uint_as_float(t4.y);
                            v[4 * j + 2] += __uint_as_float(t4.z);
                            v[4 * j + 3] += __uint_as_float(t4.w);
                        }
                    }
                    if (act == HN_ACT_LEAKY) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f) + slope * fminf(v[j], 0.f);
                    } else if (act == HN_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                    }
                    __nv_bfloat162 h[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) sts128(off[si * 4 + j], *reinterpret_cast<const uint4 *>(&h[4 * j]));
                }
                tcgen05_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) tempty_arrive(acc);
                named_bar_sync(1 + grp, 128);           // all four staging quarters written, previous logits read
                if (q == 0 && lane == 0) {
                    bool issue = true;
                    if constexpr (PAIR) {
                        // the pair's classifier MMA reads the staging tiles of BOTH CTAs: this group of either CTA reports "staged" on the
                        // leader's barrier, the leader's thread issues for both and commits to the head barrier of both
                        mbar_arrive_cluster(mapa_u32(head_ready, 0));
                        issue = cluster_ctarank() == 0;
                        if (issue) mbar_wait_cluster(head_ready, rdy_phase);
                        rdy_phase ^= 1;
                    }
                    if (issue) {
                        tcgen05_fence_after();
                        const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(epi_stage + (ew & 4) * EPI_STAGE_BYTES));
                        const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(epi_stage + NUM_EPI_WARPS * EPI_STAGE_BYTES));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_t<PAIR>(head_d2, adesc + 2 * k, bdesc + 2 * k, make_idesc_bf16(PAIR ? 256 : 128, HEAD_MAX), k != 0);
                        umma_commit_t<PAIR>(head_bar);
                    }
                }
                have_prev = true;
                prev_img = img; prev_ho = ho; prev_wo = wo; prev_valid = valid;
                continue;
            }
        }
        for (int ck = 0; ck < COLS; ck += tch) {
            const bool chunk_on = ctile + ck < p.Cout;          // warp-uniform
            const uint32_t sdelta = RES2 ? (cc & 1u) * STAGE2_DELTA : 0u;   // this chunk's staging tile
            if constexpr (RES2) {
                if (tma_out && lane == 0) {
                    if (p.tma_res) {
                        // all earlier stores have drained: both tiles are reusable; request this chunk's residual if nobody has yet
                        // (the warp's first chunk), then the NEXT chunk's into the other tile
                        { HN_PROF_T0(); bulk_wait_read0(); HN_PROF_ADD(ew_bulk); }
                        if (!primed && chunk_on) {
                            const uint32_t rb = (cc & 1u) ? rbar2 : rbar;
                            mbar_expect_tx(rb, EPI_STAGE_BYTES);
                            tma_load_4d(stage + sdelta, &tmap_r, rb, ctile + ck, bx, by, img);
                        }
                        int ntile = tile, nck = ck + tch;
                        if (nck >= COLS) { nck = 0; ntile = tile + (int)gridDim.x; }
                        if (ntile < num_tiles) {
                            int nnt, nmt;
                            decode_tile<PAIR>(p, ntile, nnt, nmt);
                            const int ntw = nmt % p.tiles_w, nth = (nmt / p.tiles_w) % p.tiles_h, nimg = nmt / (p.tiles_w * p.tiles_h);
                            const int nctile = nnt * BLOCK_N + col0 + nck;
                            if (nctile < p.Cout) {
                                const uint32_t rb = (cc & 1u) ? rbar : rbar2;           // the other tile's barrier
                                mbar_expect_tx(rb, EPI_STAGE_BYTES);
                                tma_load_4d(stage + (STAGE2_DELTA - sdelta), &tmap_r, rb, nctile, ntw * p.TW + (q * 32) % p.TW,
                                            nth * p.TH + (q * 32) / p.TW, nimg);
                            }
                        }
                    } else {
                        // no residual: only the store that read THIS tile (two chunks ago) must have drained
                        HN_PROF_T0(); bulk_wait_read1(); HN_PROF_ADD(ew_bulk);
                    }
                }
                primed = true;
            } else if (tma_out && lane == 0) {
                { HN_PROF_T0(); bulk_wait_read0(); HN_PROF_ADD(ew_bulk); }   // previous store has drained the staging tile
                if (p.tma_res && chunk_on) {
                    mbar_expect_tx(rbar, EPI_STAGE_BYTES);
                    tma_load_4d(stage, &tmap_r, rbar, ctile + ck, bx, by, img);
                }
            }
            if (tma_out) __syncwarp();
            if (!waited) {
                { HN_PROF_T0(); mbar_wait(smem_u32(tfull_bar + acc), acc_phase); HN_PROF_ADD(ew_tfull); }
                tcgen05_fence_after();
                waited = true;
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS + col0 + ck;
            if (p.tma_res && chunk_on) {
                if (RES2 && (cc & 1u)) {
                    { HN_PROF_T0(); mbar_wait(rbar2, rphase2); HN_PROF_ADD(ew_res); }
                    rphase2 ^= 1;
                } else {
                    { HN_PROF_T0(); mbar_wait(rbar, rphase); HN_PROF_ADD(ew_res); }
                    rphase ^= 1;
                }
            }
            const int nsub = (tch < COLS - ck ? tch : COLS - ck) / SUB;
            for (int si = 0; si < nsub; ++si) {
                uint32_t raw[SUB];
                HN_PROF_T0();
                if constexpr (SUB == 32) tmem_ld_32x32(taddr + si * SUB, raw);
                else if constexpr (SUB == 16) tmem_ld_32x16(taddr + si * SUB, raw);
                if constexpr (NACC == 2 && SUB == 32) {
                    uint32_t raw2[SUB];
                    tmem_ld_32x32(taddr + SUB_ACC + si * SUB, raw2);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < SUB; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(raw2[j]));
                } else {
                    tmem_ld_wait();
                }
                HN_PROF_ADD(ew_ld);
                const int cbase = ctile + ck + si * SUB;
                if (chunk_on) {
                    float v[SUB];
#pragma unroll
                    for (int j = 0; j < SUB; ++j) v[j] = __uint_as_float(raw[j]);
                    if (fast) {
                        if constexpr (SUB == 32) {
                            if (p.shift) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const uint4 t4 = lds128(tab_addr + (ck + si * SUB + 4 * j) * 4);
                                    v[4 * j] += __uint_as_float(t4.x);
                                    v[4 * j + 1] += __uint_as_float(t4.y);
                                    v[4 * j + 2] += __uint_as_float(t4.z);
                                    v[4 * j + 3] += __uint_as_float(t4.w);
                                }
                            }
                            if (act == HN_ACT_LEAKY) {
#pragma unroll
                                for (int j = 0; j < SUB; ++j) v[j] = fmaxf(v[j], 0.f) + slope * fminf(v[j], 0.f);
                            }
                            __nv_bfloat162 h[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                            if (p.res) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const uint4 rv = lds128((off[si * 4 + j] + sdelta));
                                    const __nv_bfloat162 *r2 = reinterpret_cast<const __nv_bfloat162 *>(&rv);
#pragma unroll
                                    for (int i = 0; i < 4; ++i) h[4 * j + i] = __hadd2(h[4 * j + i], r2[i]);
                                }
                            }
                            if (act == HN_ACT_RELU) {
                                const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
                                for (int j = 0; j < 16; ++j) h[j] = __hmax2(h[j], z);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) sts128((off[si * 4 + j] + sdelta), *reinterpret_cast<const uint4 *>(&h[4 * j]));
                        }
                    } else {
                        // ---------------- general path ----------------
                        const bool full = cbase + SUB <= p.Cout;
                        if (p.scale) {
#pragma unroll
                            for (int j = 0; j < SUB; ++j)
                                if (cbase + j < p.Cout) v[j] *= __ldg(p.scale + cbase + j);
                        }
                        if (p.shift) {
#pragma unroll
                            for (int j = 0; j < SUB; ++j) v[j] += tab[ck + si * SUB + j];
                        }
                        if (p.res) {
                            if (p.tma_res) {          // residual chunk sits in the staging tile (BF16, swizzled)
#pragma unroll
                                for (int j = 0; j < SUB / 8; ++j) {
                                    const uint4 rv = lds128((off[((si * SUB) >> 3) + j] + sdelta));
                                    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&rv);
#pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        const float2 f = __bfloat1622float2(h[i]);
                                        v[8 * j + 2 * i] += f.x;
                                        v[8 * j + 2 * i + 1] += f.y;
                                    }
                                }
                            } else if (valid) {
                                if (res_vec && full) {
#pragma unroll
                                    for (int j = 0; j < SUB / 8; ++j) {
                                        float r8[8];
                                        Vec8<__nv_bfloat16>::load(rrow + ck + si * SUB + 8 * j, r8);
#pragma unroll
                                        for (int i = 0; i < 8; ++i) v[8 * j + i] += r8[i];
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < SUB; ++j)
                                        if (cbase + j < p.Cout) v[j] += __bfloat162float(rrow[ck + si * SUB + j]);
                                }
                            }
                        }
#pragma unroll
                        for (int j = 0; j < SUB; ++j) v[j] = apply_act(v[j], act, slope);
                        if (tma_out) {
                            if (p.y_f32) {
#pragma unroll
                                for (int j = 0; j < SUB / 4; ++j) {
                                    uint4 o = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                                                         __float_as_uint(v[4 * j + 3]));
                                    sts128((off[((si * SUB) >> 2) + j] + sdelta), o);
                                }
                                if constexpr (SUB == 32) {
                                    if (p.stat_sum) {
                                        // the staging tile now holds this warp's 32 pixels x 32 channels in FP32 (16-byte chunks swizzled by
                                        // row): lane c sums channel c over the VALID rows (ragged tiles hold bias-only garbage beyond the
                                        // image) into per-warp register accumulators, flushed (FP64 atomics) when the channel tile changes
                                        __syncwarp();
                                        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
                                        float cs = 0.f, cq = 0.f;
                                        if (vmask == 0xffffffffu) {
#pragma unroll 8
                                            for (int r = 0; r < 32; ++r) {
                                                const float x = lds32((stage + sdelta) + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
                                                cs += x;
                                                cq = fmaf(x, x, cq);
                                            }
                                        } else {
                                            for (int r = 0; r < 32; ++r) {
                                                if ((vmask >> r) & 1u) {
                                                    const float x = lds32((stage + sdelta) + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
                                                    cs += x;
                                                    cq = fmaf(x, x, cq);
                                                }
                                            }
                                        }
#pragma unroll
                                        for (int i = 0; i < STAT_CHUNKS; ++i)
                                            if (i == (ck >> 5)) { st_s[i] += cs; st_q[i] += cq; }
                                    }
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < SUB / 8; ++j) {
                                    uint4 o;
                                    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
                                    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]);
                                    sts128((off[((si * SUB) >> 3) + j] + sdelta), o);
                                }
                            }
                        } else if (valid) {
                            if (p.y_f32) {
                                float *yp = (float *)p.y + pix * p.ldy + cbase;
                                if (full && (p.ldy & 3) == 0) {
#pragma unroll
                                    for (int j = 0; j < SUB; j += 4) *reinterpret_cast<float4 *>(yp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                                } else {
#pragma unroll
                                    for (int j = 0; j < SUB; ++j)
                                        if (cbase + j < p.Cout) yp[j] = v[j];
                                }
                            } else {
                                __nv_bfloat16 *yp = (__nv_bfloat16 *)p.y + pix * p.ldy + cbase;
                                if (full && (p.ldy & 7) == 0) {
#pragma unroll
                                    for (int j = 0; j < SUB; j += 8) {
                                        float o8[8];
#pragma unroll
                                        for (int i = 0; i < 8; ++i) o8[i] = v[j + i];
                                        Vec8<__nv_bfloat16>::store(yp + j, o8);
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < SUB; ++j)
                                        if (cbase + j < p.Cout) yp[j] = __float2bfloat16_rn(v[j]);
                                }
                            }
                        }
                    }
                }
            }
            if constexpr (SUB == 32) {
                if (stat_bf16 && fast && chunk_on) {
                    // the staging tile holds this warp's 32 pixels x (up to) 64 channels as BF16 (16-byte chunks swizzled by row): lane l
                    // sums the channel pair (2l, 2l+1) over the VALID rows
                    __syncwarp();
                    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
                    const int ncol = (tch < COLS - ck ? tch : COLS - ck);
                    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
                    if (2 * lane < ncol) {
#pragma unroll 8
                        for (int r = 0; r < 32; ++r) {
                            if (vmask == 0xffffffffu || ((vmask >> r) & 1u)) {
                                const uint32_t w2 = lds32u((stage + sdelta) + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
                                const float x0 = __uint_as_float(w2 << 16), x1 = __uint_as_float(w2 & 0xffff0000u);
                                s0 += x0; q0 = fmaf(x0, x0, q0);
                                s1 += x1; q1 = fmaf(x1, x1, q1);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < STAT_CHUNKS / 2; ++i)
                        if (i == (ck >> 6)) { st_s[2 * i] += s0; st_q[2 * i] += q0; st_s[2 * i + 1] += s1; st_q[2 * i + 1] += q1; }
                }
            }
            if (tma_out) {
                HN_PROF_T0();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0 && chunk_on) {
                    tma_store_4d(&tmap_y, (stage + sdelta), ctile + ck, bx, by, img);
                    bulk_commit();
                }
                HN_PROF_ADD(ew_st);
            }
            ++cc;
        }
        {
            HN_PROF_T0();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tempty_arrive(acc);
            HN_PROF_ADD(ew_arr);
        }
    }
    if (p.stat_sum) stat_flush();
    if constexpr (BLOCK_N == 64) {
        if (head && have_prev) head_emit(p, head_d2 + ((uint32_t)(q * 32) << 16), head_bar, hph, prev_img, prev_ho, prev_wo, prev_valid);
    }
    if (tma_out && lane == 0) bulk_wait_read0();    // staging tile must outlive the last bulk store
#ifdef HN_PROFILE_ROLES
    if (ew == 0 && lane == 0) {
        HN_PROF_FLUSH(5, ew_tfull); HN_PROF_FLUSH(6, ew_bulk); HN_PROF_FLUSH(7, ew_res); HN_PROF_FLUSH(8, clock64() - epi_t0);
        HN_PROF_FLUSH(12, ew_ld); HN_PROF_FLUSH(13, ew_st); HN_PROF_FLUSH(14, ew_arr);
    }
#endif
}

}  // namespace hn
