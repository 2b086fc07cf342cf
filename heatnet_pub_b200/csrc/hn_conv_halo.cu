// 3x3 stride-1 convolution from ONE shared-memory halo patch per k-block, optionally fused with the 2x bilinear
// upsample that precedes it in PSPUpsample (cm/models/pspnet.py:37-40).
//
// The generic implicit-GEMM kernel (hn_conv_tc.cu) fetches the A tile of every filter tap separately: 9 TMA boxes per
// 64-channel k-block, 9x the L2->SM traffic of the input patch.  For layers with few output channels (Cout = 64: 64 MACs
// per fetched element) that traffic, not the tensor pipe, is the bound.  Here an M tile is a 16x8 patch of output
// pixels and ONE (16+2d)x(8+2d)-pixel halo patch of the input is staged per k-block; the A operand of tap (r,s) is the
// sub-window that starts (r*d) patch rows and (s*d) pixels into it.  That works because the 128B swizzle of both TMA and
// tcgen05 is a pure function of the shared-memory ADDRESS bits (chunk ^= (addr >> 7) & 7): a K-major SWIZZLE_128B
// descriptor may start at any 128-byte row of the patch, and its stride-byte-offset (distance between 8-row groups =
// one 8-pixel tile row) is simply the patch row pitch (8+2d)*128 B.  (Verified on B200 by scripts/probe_umma_offset.cu.)
// Weights stream through their own ring, one [BLOCK_N][64] tile per (k-block, tap); when the whole filter fits in the
// ring (Cin = 64, Cout = 64: 72 KB) it is loaded once per CTA and stays resident across all tiles.
//
// UPSAMPLE variant: the halo patch is not fetched but COMPUTED by producer warps from the low-resolution input with the
// exact ATen bilinear rule (align_corners = False, FP32 weights), written with the same address-based swizzle.  The 2x
// upsampled activation (the three largest tensors of the network) never exists in HBM.
#include <stdlib.h>
#include <string.h>

#include "hn_common.cuh"
#include "hn_tc_epilogue.cuh"
#include "hn_tc_ptx.cuh"

namespace hn {

constexpr int HALO_THREADS = 384;
constexpr int HALO_TH = 16, HALO_TW = 8;
constexpr int HALO_NA_MAX = 6;    // A patch slots (upper bound; HaloParams::na are used)
constexpr int HALO_NB_MAX = 16;   // B ring stages (upper bound)

struct HaloParams {
    TcParams t;
    int PW, PH;              // patch extent in pixels
    int a_slot_bytes;        // patch bytes rounded up to 1024
    int na;                  // A patch slots: 2, or up to 4 for CTA pairs (a tile lasts half as long, the patch load does not)
    int nb;                  // B ring depth (stages of TPS filter taps each)
    int b_resident;          // ring holds every (k-block, tap) tile: load once per CTA
    // upsample variant: low-resolution input [N][Hl][Wl][C]
    const __nv_bfloat16 *xlow;
    int ldx, Hl, Wl;
};

__device__ __forceinline__ void bilinear_src_h(int dst, int in_size, int &i0, int &i1, float &l1)
{
    float src = 0.5f * ((float)dst + 0.5f) - 0.5f;     // scale = in/out = 0.5 exactly
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

// PAIR: clusters of two CTAs work on two horizontally adjacent tiles with ONE M = 256 instruction stream (hn_tc_ptx.cuh, "CTA
// pairs"): each CTA stages its own halo patch and half of the weight rows, the leader issues and commits for both.  These layers
// are bound by the per-instruction floor of the tensor pipe, which the pair form halves.
template <int BLOCK_N, bool UPSAMPLE, bool PAIR = false>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_r, const HaloParams hp)
{
    constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;      // weight rows staged by this CTA
    // a B ring stage holds TPS filter taps (TPS boxes on one barrier); pairs take a whole filter row per stage: at 43 cycles
    // per MMA the issuing warp cannot afford a barrier wait + commit every 4 instructions
    constexpr int TPS = PAIR ? 3 : 1;
    constexpr int B_TAP_BYTES = B_ROWS * 128;
    constexpr int B_STAGE_BYTES = TPS * B_TAP_BYTES;
    // BLOCK_N = 64: the 36+ MMAs of a tile alternate between TWO accumulators (independent dependency chains on the tensor
    // pipe), summed by the epilogue; + 2 x 16 columns for the fused classifier head
    // (pairs: one chain already retires an instruction every 43 cycles; the TMEM columns go into FOUR tile buffers instead, because
    // a tile lasts half as long while the epilogue's latency and the cross-CTA hand-over do not shrink)
    constexpr int NACC = (BLOCK_N == 64 && !PAIR) ? 2 : 1;
    constexpr int NBUF = (BLOCK_N == 64 && PAIR) ? 4 : 2;
    constexpr int SUB_ACC = BLOCK_N < 32 ? 32 : BLOCK_N;
    constexpr int ACC_COLS = NACC * SUB_ACC;
    constexpr int TMEM_COLS = BLOCK_N == 64 ? 512 : 2 * ACC_COLS;
    static_assert(NBUF * ACC_COLS + 2 * HEAD_MAX <= TMEM_COLS || BLOCK_N != 64, "accumulators + classifier columns exceed the TMEM allocation");
    constexpr uint32_t IDESC = make_idesc_bf16(PAIR ? 256 : 128, BLOCK_N);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    // A-producer warps of the UPSAMPLE variant: warps 2,3 always; warps 8-11 too when the epilogue only needs 4 warps
    constexpr int NPROD = 6;   // warps 2, 3, 8-11 (the UPSAMPLE variant's epilogue runs on warps 4-7 only)
    const TcParams &p = hp.t;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *a_ring = smem;
    uint8_t *b_ring = a_ring + hp.na * hp.a_slot_bytes;
    uint8_t *epi_stage = b_ring + hp.nb * B_STAGE_BYTES;
    // [16][64] BF16 classifier tile (2 KB, 1024-aligned) sits right behind the staging buffers: hn_tc_epilogue.cuh relies on it
    uint64_t *bars = reinterpret_cast<uint64_t *>(epi_stage + NUM_EPI_WARPS * EPI_STAGE_BYTES + 2048);
    uint64_t *a_full = bars, *a_empty = bars + HALO_NA_MAX, *b_full = bars + 2 * HALO_NA_MAX, *b_empty = b_full + HALO_NB_MAX;
    uint64_t *tfull_bar = b_empty + HALO_NB_MAX, *tempty_bar = tfull_bar + 4, *res_bar = tempty_bar + 4;   // up to 4 tile buffers
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(res_bar + NUM_EPI_WARPS);
    float *s_shift = reinterpret_cast<float *>(tmem_slot + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m_tiles = p.n_img * p.tiles_h * p.tiles_w;
    const int num_tiles = PAIR ? pair_num_tiles(num_m_tiles, p.n_tiles) : num_m_tiles * p.n_tiles;
    const int num_kb = p.cblocks;
    const int d = p.dil;

    if (warp == 0 && lane == 0) {
        if (!UPSAMPLE) prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        if (p.tma_out) prefetch_tmap(&tmap_y);
        if (p.tma_res) prefetch_tmap(&tmap_r);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < HALO_NA_MAX; ++i) {
            mbar_init(smem_u32(a_full + i), UPSAMPLE ? NPROD * (PAIR ? 2 : 1) : 1);   // computed patches: the producer warps of both CTAs arrive
            mbar_init(smem_u32(a_empty + i), 1);
        }
        for (int i = 0; i < HALO_NB_MAX; ++i) {
            mbar_init(smem_u32(b_full + i), 1);
            mbar_init(smem_u32(b_empty + i), 1);
        }
        for (int i = 0; i < NBUF; ++i) {
            mbar_init(smem_u32(tfull_bar + i), 1);
            mbar_init(smem_u32(tempty_bar + i), ((BLOCK_N >= 128 && !UPSAMPLE) ? 8 : 4) * (PAIR ? 2 : 1));
        }
        // (with a classifier head, pairs use the second res_bar of each epilogue group as the two-CTA "tile staged" barrier)
        for (int i = 0; i < NUM_EPI_WARPS; ++i) mbar_init(smem_u32(res_bar + i), (PAIR && p.head_n > 0 && (i & 3) == 1) ? 2 : 1);
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

    const bool is_prod = UPSAMPLE && (warp == 2 || warp == 3 || warp >= 8);
    const bool is_epi = warp >= EPI_WARP0 && !(UPSAMPLE && warp >= 8);

    if (warp == 0) {
        // ===================== TMA producer: weight tiles (and the halo patches when they are fetched) =====================
        // whole warp runs the loop (uniform values stay in uniform registers), the elected lane issues
        {
            int aslot = 0, bst = 0;
            uint32_t aphase = 0, bphase = 0;
            bool first = true;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int nt, mt;
                decode_tile<PAIR>(p, tile, nt, mt);
                const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, img = mt / (p.tiles_w * p.tiles_h);
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (!UPSAMPLE) {
                        mbar_wait(smem_u32(a_empty + aslot), aphase ^ 1);
                        const uint32_t fb = smem_u32(a_full + aslot);
                        if (elect_one()) {
                            if constexpr (PAIR) {          // the patches of both CTAs are counted on the leader's barrier
                                if (rank == 0) mbar_expect_tx(fb, 2 * hp.PW * hp.PH * 128);
                                tma_load_4d_pair(smem_u32(a_ring + aslot * hp.a_slot_bytes), &tmap_a, mapa_u32(fb, 0), kb * 64, tw * HALO_TW - d,
                                                 th * HALO_TH - d, img);
                            } else {
                                mbar_expect_tx(fb, hp.PW * hp.PH * 128);
                                tma_load_4d(smem_u32(a_ring + aslot * hp.a_slot_bytes), &tmap_a, fb, kb * 64, tw * HALO_TW - d, th * HALO_TH - d, img);
                            }
                        }
                        __syncwarp();
                        if (++aslot == hp.na) { aslot = 0; aphase ^= 1; }
                    }
                    if (hp.b_resident && !first) continue;
                    for (int tap = 0; tap < 9; tap += TPS) {
                        if (!hp.b_resident) mbar_wait(smem_u32(b_empty + bst), bphase ^ 1);
                        const uint32_t fb = smem_u32(b_full + bst);
                        if (elect_one()) {
                            if constexpr (PAIR) {
                                const uint32_t lfb = mapa_u32(fb, 0);
                                if (rank == 0) mbar_expect_tx(fb, 2 * B_STAGE_BYTES);
#pragma unroll
                                for (int t = 0; t < TPS; ++t)
                                    tma_load_2d_pair(smem_u32(b_ring + bst * B_STAGE_BYTES + t * B_TAP_BYTES), &tmap_b, lfb, ((tap + t) * num_kb + kb) * 64,
                                                     nt * BLOCK_N + (int)rank * B_ROWS);
                            } else {
                                mbar_expect_tx(fb, B_STAGE_BYTES);
#pragma unroll
                                for (int t = 0; t < TPS; ++t)
                                    tma_load_2d(smem_u32(b_ring + bst * B_STAGE_BYTES + t * B_TAP_BYTES), &tmap_b, fb, ((tap + t) * num_kb + kb) * 64, nt * BLOCK_N);
                            }
                        }
                        __syncwarp();
                        if (++bst == hp.nb) { bst = 0; bphase ^= 1; }
                    }
                }
                first = false;
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (whole warp, elected lane issues; the leader CTA of a pair) =====================
        {
            int aslot = 0, bst = 0, acc = 0;
            uint32_t aphase = 0, bphase = 0, acc_phase = 0;
            bool first = true;
            long long mw_afull = 0, mw_bfull = 0, mw_tempty = 0, ntl = 0;
            (void)mw_afull; (void)mw_bfull; (void)mw_tempty; (void)ntl;
#ifdef HN_PROFILE_ROLES
            const long long mma_t0 = clock64();
#endif
            const uint32_t sbo = (uint32_t)(hp.PW * 128) >> 4;
            // everything of the A descriptor but the start address: LBO = 1, SBO = patch row pitch, version 1, SWIZZLE_128B
            const uint64_t adesc_hi = ((uint64_t)1 << 16) | ((uint64_t)(sbo & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            // B: canonical K-major SWIZZLE_128B tile (make_kmajor_sw128_desc) without the start address
            const uint64_t bdesc_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            // start-address fields (address >> 4; the whole dynamic shared memory lies below 256 KB, so sums never leave the 14-bit field)
            const uint32_t a_lo0 = (smem_u32(a_ring) & 0x3FFFFu) >> 4, a_slot16 = (uint32_t)hp.a_slot_bytes >> 4;
            const uint32_t b_lo0 = (smem_u32(b_ring) & 0x3FFFFu) >> 4;
            uint32_t tap_off[9];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) tap_off[tap] = (uint32_t)(((tap / 3) * d * hp.PW + (tap % 3) * d) * 128) >> 4;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                {
                    HN_PROF_T0();
                    if constexpr (PAIR) mbar_wait_cluster(smem_u32(tempty_bar + acc), acc_phase ^ 1);
                    else mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);
                    HN_PROF_ADD(mw_tempty);
                }
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                if (hp.b_resident) bst = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    {
                        HN_PROF_T0();
                        if constexpr (PAIR && UPSAMPLE) mbar_wait_cluster(smem_u32(a_full + aslot), aphase);
                        else mbar_wait(smem_u32(a_full + aslot), aphase);
                        HN_PROF_ADD(mw_afull);
                    }
                    tcgen05_fence_after();
                    // Descriptors are assembled from low words prepared outside the loops (slot / stage base >> 4 plus per-tap offsets):
                    // every instruction between two MMAs costs issue time once a tile instruction retires every ~43 cycles
                    // (profiles/r2_mma_rate_probe_mn_major_and_issue_loop.txt), and ONE elected region covers all taps of a weight stage.
                    const uint32_t a_lo = a_lo0 + (uint32_t)aslot * a_slot16;
#pragma unroll
                    for (int tg = 0; tg < 9 / TPS; ++tg) {
                        if (!hp.b_resident || first) {
                            HN_PROF_T0();
                            mbar_wait(smem_u32(b_full + bst), bphase);
                            HN_PROF_ADD(mw_bfull);
                            tcgen05_fence_after();
                        }
                        const uint32_t b_lo = b_lo0 + (uint32_t)bst * (uint32_t)(B_STAGE_BYTES >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int t = 0; t < TPS; ++t) {
                                const int tap = tg * TPS + t;
                                // sub-window of the patch: starts (r*d) patch rows and (s*d) pixels in; 8-row groups are one patch row apart
                                const uint64_t adesc = adesc_hi | (uint64_t)(a_lo + tap_off[tap]);
                                const uint64_t bdesc = bdesc_hi | (uint64_t)(b_lo + (uint32_t)t * (uint32_t)(B_TAP_BYTES >> 4));
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_bf16_t<PAIR>(d_tmem + (NACC == 2 ? (tap & 1) * SUB_ACC : 0), adesc + 2 * k, bdesc + 2 * k, IDESC,
                                                      NACC == 2 ? ((kb | (tap >> 1) | k) != 0) : ((kb | tap | k) != 0));
                            }
                            if (!hp.b_resident) umma_commit_t<PAIR>(smem_u32(b_empty + bst));
                            if (tg == 9 / TPS - 1) {
                                umma_commit_t<PAIR>(smem_u32(a_empty + aslot));
                                if (kb == num_kb - 1) umma_commit_t<PAIR>(smem_u32(tfull_bar + acc));
                            }
                        }
                        __syncwarp();
                        if (++bst == hp.nb) { bst = 0; bphase ^= 1; }
                    }
                    if (++aslot == hp.na) { aslot = 0; aphase ^= 1; }
                }
                if (++acc == NBUF) { acc = 0; acc_phase ^= 1; }
                first = false;
                ++ntl;
            }
#ifdef HN_PROFILE_ROLES
            if (lane == 0) {
                HN_PROF_FLUSH(2, mw_afull); HN_PROF_FLUSH(11, mw_bfull); HN_PROF_FLUSH(3, mw_tempty); HN_PROF_FLUSH(4, clock64() - mma_t0);
                HN_PROF_FLUSH(9, ntl); HN_PROF_FLUSH(10, 1);
            }
#endif
        }
    } else if (is_prod) {
        // ===================== halo patch producers (UPSAMPLE): bilinear 2x of the low-res input on the fly =====================
        const int pi = warp == 2 ? 0 : (warp == 3 ? 1 : warp - 8 + 2);
        int aslot = 0;
        uint32_t aphase = 0;
        const int Hu = 2 * hp.Hl, Wu = 2 * hp.Wl;
        const int ntask = hp.PH * hp.PW * 8;
        const uint32_t a_full0 = PAIR ? mapa_u32(smem_u32(a_full), 0) : smem_u32(a_full);   // the leader's "patch ready" barriers
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int nt, mt;
            decode_tile<PAIR>(p, tile, nt, mt);
            (void)nt;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, img = mt / (p.tiles_w * p.tiles_h);
            const int uy0 = th * HALO_TH - 1, ux0 = tw * HALO_TW - 1;
            const bool img_ok = !PAIR || img < p.n_img;          // the filler tile of an odd pair: an all-zero patch
            const __nv_bfloat16 *ximg = hp.xlow + (int64_t)(img_ok ? img : 0) * hp.Hl * hp.Wl * hp.ldx;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(smem_u32(a_empty + aslot), aphase ^ 1);
                const uint32_t a_base = smem_u32(a_ring + aslot * hp.a_slot_bytes);
                // 4 tasks (one 16-byte channel chunk of one patch pixel each) per iteration: all 16 loads are issued before
                // the first is consumed, so the L1/L2 latency of the gathers overlaps
                constexpr int UNR = 4, STRIDE = NPROD * 32;
                for (int t0 = pi * 32 + lane; t0 < ntask; t0 += UNR * STRIDE) {
                    uint4 v[UNR][4];
                    float ly[UNR], lx[UNR];
                    uint32_t dst[UNR];
                    bool live[UNR], inb[UNR];
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        const int t = t0 + u * STRIDE;
                        live[u] = t < ntask;
                        inb[u] = false;
                        if (!live[u]) continue;
                        const int pp = t >> 3, j = t & 7;
                        const int py = pp / hp.PW, px = pp - py * hp.PW;
                        const int uy = uy0 + py, ux = ux0 + px;
                        const uint32_t rowaddr = a_base + pp * 128;
                        dst[u] = rowaddr + ((j ^ ((rowaddr >> 7) & 7)) << 4);
                        if (img_ok && uy >= 0 && uy < Hu && ux >= 0 && ux < Wu) {
                            inb[u] = true;
                            int y0, y1, x0, x1;
                            bilinear_src_h(uy, hp.Hl, y0, y1, ly[u]);
                            bilinear_src_h(ux, hp.Wl, x0, x1, lx[u]);
                            const __nv_bfloat16 *b = ximg + kb * 64 + j * 8;
                            v[u][0] = __ldg(reinterpret_cast<const uint4 *>(b + ((int64_t)y0 * hp.Wl + x0) * hp.ldx));
                            v[u][1] = __ldg(reinterpret_cast<const uint4 *>(b + ((int64_t)y0 * hp.Wl + x1) * hp.ldx));
                            v[u][2] = __ldg(reinterpret_cast<const uint4 *>(b + ((int64_t)y1 * hp.Wl + x0) * hp.ldx));
                            v[u][3] = __ldg(reinterpret_cast<const uint4 *>(b + ((int64_t)y1 * hp.Wl + x1) * hp.ldx));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        if (!live[u]) continue;
                        uint4 outv = make_uint4(0, 0, 0, 0);
                        if (inb[u]) {
                            const float hy = 1.f - ly[u], hx = 1.f - lx[u];
                            const __nv_bfloat162 *h00 = reinterpret_cast<const __nv_bfloat162 *>(&v[u][0]);
                            const __nv_bfloat162 *h01 = reinterpret_cast<const __nv_bfloat162 *>(&v[u][1]);
                            const __nv_bfloat162 *h10 = reinterpret_cast<const __nv_bfloat162 *>(&v[u][2]);
                            const __nv_bfloat162 *h11 = reinterpret_cast<const __nv_bfloat162 *>(&v[u][3]);
                            __nv_bfloat162 *ho = reinterpret_cast<__nv_bfloat162 *>(&outv);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float2 a00 = __bfloat1622float2(h00[q]), a01 = __bfloat1622float2(h01[q]);
                                const float2 a10 = __bfloat1622float2(h10[q]), a11 = __bfloat1622float2(h11[q]);
                                const float ox = hy * (hx * a00.x + lx[u] * a01.x) + ly[u] * (hx * a10.x + lx[u] * a11.x);
                                const float oy = hy * (hx * a00.y + lx[u] * a01.y) + ly[u] * (hx * a10.y + lx[u] * a11.y);
                                ho[q] = __floats2bfloat162_rn(ox, oy);
                            }
                        }
                        sts128(dst[u], outv);
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (PAIR) mbar_arrive_cluster(a_full0 + aslot * 8);
                    else mbar_arrive(smem_u32(a_full + aslot));
                }
                if (++aslot == hp.na) { aslot = 0; aphase ^= 1; }
            }
        }
    } else if (is_epi) {
        conv_epilogue<BLOCK_N, UPSAMPLE, NACC, PAIR, NBUF>(p, &tmap_y, &tmap_r, tmem_base, tfull_bar, tempty_bar, res_bar, epi_stage, s_shift, num_tiles, warp, lane);
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

template <int BN, bool UP, bool PAIR = false>
static int launch_halo(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &ty, const CUtensorMap &tr, const HaloParams &hp,
                       size_t smem, int num_m_tiles, cudaStream_t st)
{
    static bool configured_dev[HN_MAX_DEVICES] = {};
    static int pairs_dev[HN_MAX_DEVICES] = {};
    bool &configured = configured_dev[current_device()];
    int &pairs = pairs_dev[current_device()];
    if (!configured) {
        HN_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN, UP, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        if (PAIR) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2, 1, 1);
            cfg.blockDim = dim3(HALO_THREADS, 1, 1);
            cfg.dynamicSmemBytes = 227 * 1024;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&pairs, conv_halo_kernel<BN, UP, PAIR>, &cfg) != cudaSuccess || pairs < 1) {
                cudaGetLastError();
                pairs = num_sms() / 2;
            }
        }
        configured = true;
    }
    if constexpr (PAIR) {
        const int want = ((num_m_tiles + 1) / 2) * hp.t.n_tiles;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (unsigned)(want < pairs ? want : pairs), 1, 1);
        cfg.blockDim = dim3(HALO_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        HN_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<BN, UP, PAIR>, ta, tb, ty, tr, hp));
    } else {
        const int num_tiles = num_m_tiles * hp.t.n_tiles;
        int grid = num_tiles < num_sms() ? num_tiles : num_sms();
        conv_halo_kernel<BN, UP, PAIR><<<grid, HALO_THREADS, smem, st>>>(ta, tb, ty, tr, hp);
    }
    HN_LAUNCH_CHECK();
    return HN_OK;
}

// Is the halo kernel applicable / profitable for this 3x3 convolution of x (or of the 2x upsample of x)?
bool conv_halo_ok(const hn_tensor *x, const hn_conv *cv, const hn_tensor *y, bool upsample)
{
    static const bool disabled = getenv("HN_NO_HALO") != nullptr;
    if (disabled && !upsample) return false;
    if (x->dtype != HN_BF16 || cv->r != 3 || cv->s != 3 || cv->stride != 1 || cv->pad != cv->dil) return false;
    if (cv->dil != 1 && (upsample || cv->dil != 2)) return false;
    if (x->c % 64 != 0 || x->ld % 8 != 0 || (reinterpret_cast<uintptr_t>(x->ptr) & 15) != 0) return false;
    if (hn_conv_cout_pad(cv->cout, HN_BF16) < 64) return false;
    if (upsample) return true;
    // ragged-tile waste of the fixed 16x8 tile against the best tile of the generic kernel's menu
    const int64_t halo_area = cdiv(y->h, HALO_TH) * HALO_TH * cdiv(y->w, HALO_TW) * HALO_TW;
    const int menu[5][2] = {{8, 16}, {4, 32}, {16, 8}, {2, 64}, {1, 128}};
    int64_t best = -1;
    for (int i = 0; i < 5; ++i) {
        int64_t a = cdiv(y->h, menu[i][0]) * menu[i][0] * cdiv(y->w, menu[i][1]) * menu[i][1];
        if (best < 0 || a < best) best = a;
    }
    // more padded work than the generic kernel is acceptable up to a point: 4 %, or 30 % for Cout <= 64, where the generic kernel is
    // bound by its nine tap fetches per k-block (~570 TFLOP/s) and this one, on CTA pairs, is not (~1200)
    const int waste_pct = hn_conv_cout_pad(cv->cout, HN_BF16) <= 64 ? 130 : 104;
    if (halo_area * 100 > best * waste_pct) return false;
    // the gain is the 9x smaller A traffic: decisive for narrow outputs, irrelevant for wide compute-bound layers
    return hn_conv_cout_pad(cv->cout, HN_BF16) <= 128;
}

int conv2d_fwd_halo(const hn_tensor *x, const void *w, const hn_conv *cv, const hn_epilogue *ep, const hn_tensor *y, bool upsample,
                    cudaStream_t st, const HeadArgs *head)
{
    // with a classifier head the activation is not stored: y may be NULL, the output extent is that of a pad = dil 3x3
    const int Ho = y ? y->h : (upsample ? 2 * x->h : x->h), Wo = y ? y->w : (upsample ? 2 * x->w : x->w);
    if ((int64_t)x->n * Ho * Wo == 0) return HN_OK;
    const int d = cv->dil;
    const int kpad = hn_conv_kpad(x->c, 3, 3);
    const int cout_pad = hn_conv_cout_pad(cv->cout, HN_BF16);
    HaloParams hp{};
    TcParams &p = hp.t;
    hp.PW = HALO_TW + 2 * d;
    hp.PH = HALO_TH + 2 * d;
    hp.a_slot_bytes = (int)(cdiv((int64_t)hp.PW * hp.PH * 128, 1024) * 1024);
    p.TH = HALO_TH; p.TW = HALO_TW;
    p.tiles_h = (int)cdiv(Ho, HALO_TH); p.tiles_w = (int)cdiv(Wo, HALO_TW); p.n_img = x->n;
    p.Ho = Ho; p.Wo = Wo;
    p.R = 3; p.S = 3; p.pad = p.pad_w = cv->pad; p.dil = d; p.cblocks = x->c / 64;
    p.Cout = cv->cout;
    const int num_m_tiles = p.n_img * p.tiles_h * p.tiles_w;
    int bn = 64;
    for (int cand : {256, 128})
        if (cout_pad % cand == 0 && (int64_t)num_m_tiles * (cout_pad / cand) >= 2 * (int64_t)num_sms()) { bn = cand; break; }
    if (head) bn = 64;
    p.n_tiles = cout_pad / bn;
    const bool pair = bn <= 128 && conv_pair_ok(bn, num_m_tiles, 9 * p.cblocks, true);          // (the 256-wide tile only exists for upconv3x3)
    const int b_rows = pair ? bn / 2 : bn;          // weight rows each CTA stages per (k-block, tap)
    const int tail = 2048 /*classifier tile*/ + 2048 /*barriers + shift table*/;
    // shared-memory plan: A slots + B ring + epilogue staging + barriers
    const int tps = pair ? 3 : 1;                   // filter taps per B ring stage (the kernel's TPS)
    const int stage_bytes = tps * b_rows * 128;
    const int stages_per_kb = 9 / tps;
    static const int na_env = getenv("HN_HALO_NA") ? atoi(getenv("HN_HALO_NA")) : 0;          // tuning knob (2..4)
    hp.na = pair ? 4 : 2;
    if (na_env >= 2 && na_env <= HALO_NA_MAX) hp.na = na_env;
    int64_t fixed = 0;
    int nb = 0;
    for (;; --hp.na) {          // pairs: as many patch slots as leave room for a useful weight ring
        fixed = (int64_t)hp.na * hp.a_slot_bytes + NUM_EPI_WARPS * EPI_STAGE_BYTES + 1024 /*align*/ + tail;
        nb = (int)((227 * 1024 - fixed) / stage_bytes);
        if (nb >= (pair ? 4 : 2) || hp.na == 2) break;
    }
    if (nb > HALO_NB_MAX) nb = HALO_NB_MAX;
    HN_CHECK_ARG(nb >= 2, "conv_halo: shared memory too small for this shape");
    hp.nb = nb;
    hp.b_resident = (p.cblocks * stages_per_kb <= nb && p.n_tiles == 1) ? 1 : 0;   // one Cout tile: the same weights for every tile
    if (hp.b_resident) hp.nb = p.cblocks * stages_per_kb;
    const size_t smem = (size_t)hp.na * hp.a_slot_bytes + (size_t)hp.nb * stage_bytes + NUM_EPI_WARPS * EPI_STAGE_BYTES + 1024 + tail;

    CUtensorMap ta, tb, ty, tr;
    memset(&ta, 0, sizeof(ta));
    memset(&ty, 0, sizeof(ty));
    memset(&tr, 0, sizeof(tr));
    if (upsample) {
        hp.xlow = (const __nv_bfloat16 *)x->ptr;
        hp.ldx = x->ld; hp.Hl = x->h; hp.Wl = x->w;
    } else {
        uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
        uint64_t strides[4] = {2, (uint64_t)x->ld * 2, (uint64_t)x->ld * 2 * x->w, (uint64_t)x->ld * 2 * x->w * x->h};
        uint32_t box[4] = {64, (uint32_t)hp.PW, (uint32_t)hp.PH, 1};
        int rc = make_tmap(&ta, x->ptr, 4, dims, strides, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)kpad, (uint64_t)cout_pad};
        uint64_t strides[2] = {2, (uint64_t)kpad * 2};
        uint32_t box[2] = {64, (uint32_t)b_rows};
        int rc = make_tmap(&tb, w, 2, dims, strides, box);
        if (rc) return rc;
    }
    p.scale = ep->scale; p.shift = ep->shift; p.res = ep->residual; p.ldr = ep->residual_ld;
    p.act = ep->act; p.slope = ep->slope; p.slope_ptr = ep->slope_ptr;
    p.stat_sum = ep->stat_sum; p.stat_sqsum = ep->stat_sqsum;
    if (int rcg = set_stat_groups(p, ep, y)) return rcg;
    p.ebw = HALO_TW;
    if (head) {
        HN_CHECK_ARG(cv->cout == 64 && !ep->scale && !ep->residual, "conv3x3_head: needs Cout == 64, no explicit scale, no residual");
        HN_CHECK_ARG(head->n >= 1 && head->n <= HEAD_MAX && head->w && head->out, "conv3x3_head: 1..%d classes", HEAD_MAX);
        p.head_out = head->out; p.head_n = head->n;
        memcpy(p.head.w, head->w, sizeof(float) * 64 * head->n);          // hp{} zero-initialised the rest
        if (head->b) memcpy(p.head.b, head->b, sizeof(float) * head->n);
    } else {
        p.y = y->ptr; p.ldy = y->ld; p.y_f32 = (y->dtype == HN_F32);
        const uint64_t esz = p.y_f32 ? 4 : 2;
        const bool ok = (reinterpret_cast<uintptr_t>(y->ptr) & 15) == 0 && ((uint64_t)y->ld * esz) % 16 == 0;
        if (ok) {
            uint64_t dims[4] = {(uint64_t)cv->cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)p.n_img};
            uint64_t strides[4] = {esz, (uint64_t)y->ld * esz, (uint64_t)y->ld * esz * Wo, (uint64_t)y->ld * esz * Wo * Ho};
            uint32_t box[4] = {(uint32_t)(128 / esz), (uint32_t)p.ebw, (uint32_t)(32 / p.ebw), 1};
            int rc = make_tmap(&ty, y->ptr, 4, dims, strides, box, p.y_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
            if (rc) return rc;
            p.tma_out = 1;
            if (ep->residual && !p.y_f32 && (reinterpret_cast<uintptr_t>(ep->residual) & 15) == 0 && ((uint64_t)ep->residual_ld * 2) % 16 == 0) {
                uint64_t rs[4] = {2, (uint64_t)ep->residual_ld * 2, (uint64_t)ep->residual_ld * 2 * Wo, (uint64_t)ep->residual_ld * 2 * Wo * Ho};
                rc = make_tmap(&tr, ep->residual, 4, dims, rs, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
                if (rc) return rc;
                p.tma_res = 1;
            }
        }
    }
    if (pair) {
        if (upsample) return bn == 128 ? launch_halo<128, true, true>(ta, tb, ty, tr, hp, smem, num_m_tiles, st)
                                       : launch_halo<64, true, true>(ta, tb, ty, tr, hp, smem, num_m_tiles, st);
        return bn == 128 ? launch_halo<128, false, true>(ta, tb, ty, tr, hp, smem, num_m_tiles, st)
                         : launch_halo<64, false, true>(ta, tb, ty, tr, hp, smem, num_m_tiles, st);
    }
    if (upsample) {
        switch (bn) {
            case 256: return launch_halo<256, true>(ta, tb, ty, tr, hp, smem, num_m_tiles, st);
            case 128: return launch_halo<128, true>(ta, tb, ty, tr, hp, smem, num_m_tiles, st);
            default: return launch_halo<64, true>(ta, tb, ty, tr, hp, smem, num_m_tiles, st);
        }
    }
    switch (bn) {
        case 256: return launch_halo<256, false>(ta, tb, ty, tr, hp, smem, num_m_tiles, st);
        case 128: return launch_halo<128, false>(ta, tb, ty, tr, hp, smem, num_m_tiles, st);
        default: return launch_halo<64, false>(ta, tb, ty, tr, hp, smem, num_m_tiles, st);
    }
}

}  // namespace hn

using namespace hn;

// PSPUpsample fused: y = epilogue(conv3x3_pad1(bilinear_2x(x)))  (cm/models/pspnet.py:37-40 without the upsampled tensor)
extern "C" int hn_upconv3x3_fwd(const hn_tensor *x, const void *w_packed, const hn_conv *cv, const hn_epilogue *ep, const hn_tensor *y,
                                void *stream)
{
    HN_CHECK_ARG(x && w_packed && cv && ep && y && x->ptr && y->ptr, "hn_upconv3x3_fwd: null pointer");
    HN_CHECK_ARG(y->n == x->n && y->h == 2 * x->h && y->w == 2 * x->w && y->c == cv->cout, "hn_upconv3x3_fwd: output must be N=%d %dx%d C=%d",
                 x->n, 2 * x->h, 2 * x->w, cv->cout);
    HN_CHECK_ARG(conv_halo_ok(x, cv, y, true), "hn_upconv3x3_fwd: needs BF16 NHWC input with Cin %% 64 == 0, a 3x3 stride-1 pad-1 filter and Cout >= 33");
    HN_CHECK_ARG(!ep->out_nchw && !ep->stat_sum && !ep->stat_sqsum, "hn_upconv3x3_fwd: out_nchw / fused statistics are not implemented");
    return conv2d_fwd_halo(x, w_packed, cv, ep, y, true, (cudaStream_t)stream);
}

// up_3 + final of PSPNet in one kernel (cm/models/pspnet.py:72-75 in eval mode: conv3x3 -> BN -> PReLU -> Dropout2d (identity)
// -> 1x1 classifier): the 64-channel full-resolution activation never leaves the SM; logits are written NCHW FP32.
extern "C" int hn_conv3x3_head_fwd(const hn_tensor *x, const void *w_packed, const hn_conv *cv, const hn_epilogue *ep, const float *head_w_host,
                                   const float *head_b_host, int32_t head_classes, float *logits_nchw, void *stream)
{
    HN_CHECK_ARG(x && w_packed && cv && ep && head_w_host && logits_nchw && x->ptr, "hn_conv3x3_head_fwd: null pointer");
    HN_CHECK_ARG(x->dtype == HN_BF16 && cv->r == 3 && cv->s == 3 && cv->stride == 1 && cv->pad == 1 && cv->dil == 1 && cv->cout == 64 &&
                     x->c % 64 == 0 && x->ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x->ptr) & 15) == 0,
                 "hn_conv3x3_head_fwd: needs BF16 NHWC input with Cin %% 64 == 0 and a 3x3 stride-1 pad-1 filter with 64 outputs");
    HN_CHECK_ARG(!ep->out_nchw && !ep->stat_sum && !ep->stat_sqsum && !ep->scale && !ep->residual,
                 "hn_conv3x3_head_fwd: epilogue = shift + activation only (fold the BN scale into the filter)");
    HeadArgs h{head_w_host, head_b_host, logits_nchw, head_classes};
    return conv2d_fwd_halo(x, w_packed, cv, ep, nullptr, false, (cudaStream_t)stream, &h);
}
