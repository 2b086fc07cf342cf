// BF16 implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a):
//   D[pixels, Cout] = im2col(X)[pixels, R*S*Cin] * W[Cout, R*S*Cin]^T,   FP32 accumulators in TMEM.
//
// * Operands are staged by TMA into 128B-swizzled, K-major shared-memory tiles; one elected thread issues
//   tcgen05.mma (cta_group::1, M=128, N=BLOCK_N, K=16 per instruction, 4 per 64-channel k-block).
// * The im2col matrix is never materialised for stride-1 convolutions: an M tile is a TH x TW patch of output
//   pixels of one image (TH*TW = 128) and the A tile of filter tap (r,s) is the same patch of the NHWC input
//   shifted by (r*dil - pad, s*dil - pad), fetched as one 4-D TMA box {64 ch, TW, TH, 1}.  Out-of-bounds
//   coordinates (zero padding, ragged right/bottom tiles) are zero-filled by the TMA unit.  1x1 convolutions
//   use the same path with the pixel axis flattened (TH=1, TW=128).
// * Strided / small-Cin convolutions (7x7 s2 stems, 3x3 s2, 1x1 s2, the critics' 4x4 s2) first run a
//   bandwidth-bound im2col gather into a workspace and then the flattened path with K = kpad.
// * Warp-specialised persistent kernel: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
//   warps 4-11 = epilogue (tcgen05.ld -> scale/shift/residual/activation -> global; the residual is
//   prefetched into registers while the warp waits for the accumulator).  Two TMEM accumulator
//   buffers let the epilogue of tile i overlap the main loop of tile i+1.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "hn_common.cuh"
#include "hn_tc_ptx.cuh"
#include "hn_tc_epilogue.cuh"

namespace hn {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 BF16 = 128 B = one swizzle-128B row (the stem variant uses 32 = one swizzle-64B row)
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 384;   // 4 control warps + 8 epilogue warps

// ------------------------------------------------------------------------------------------------ kernel
// PAIR: launched as clusters of two CTAs (hn_tc_ptx.cuh, "CTA pairs"): each CTA stages its own M tile and half of the Cout tile's
// weight rows, the leader's MMA warp issues one M = 256 instruction for both and commits to the barriers of both.
// RES2: a second set of epilogue staging tiles (hn_tc_epilogue.cuh), paid for with one operand stage in the single-CTA 256-wide form.
template <int BLOCK_N, int STAGES, int BK = BLOCK_K, bool PAIR = false, bool RES2 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_r, const TcParams p)
{
    constexpr int A_STAGE_BYTES = BLOCK_M * BK * 2;
    constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;      // weight rows staged by this CTA
    constexpr int B_STAGE_BYTES = B_ROWS * BK * 2;
    constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static_assert(A_STAGE_BYTES % 1024 == 0 && B_STAGE_BYTES % 1024 == 0, "stage tiles must keep 1024-byte alignment");
    constexpr int ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;   // columns per accumulator buffer
    constexpr int TMEM_COLS = 2 * ACC_COLS;                   // power of two >= 32
    constexpr uint32_t IDESC = make_idesc_bf16(PAIR ? 2 * BLOCK_M : BLOCK_M, BLOCK_N);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int EPI_SETS = RES2 ? 2 : 1;                                  // staging tiles + residual barriers per epilogue warp
    uint8_t *epi_stage = smem + STAGES * STAGE_BYTES;                       // EPI_SETS x NUM_EPI_WARPS x 4 KB, 1024-aligned
    uint64_t *bars = reinterpret_cast<uint64_t *>(epi_stage + EPI_SETS * NUM_EPI_WARPS * EPI_STAGE_BYTES);
    uint64_t *full_bar = bars, *empty_bar = bars + STAGES, *tfull_bar = bars + 2 * STAGES, *tempty_bar = bars + 2 * STAGES + 2;
    uint64_t *res_bar = bars + 2 * STAGES + 4;                              // EPI_SETS per epilogue warp
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4 + EPI_SETS * NUM_EPI_WARPS);
    float *s_shift = reinterpret_cast<float *>(tmem_slot + 4);                // epilogue shift table(s): BLOCK_N floats (x2 for BLOCK_N < 128)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m_tiles = p.n_img * p.tiles_h * p.tiles_w;
    const int num_tiles = PAIR ? pair_num_tiles(num_m_tiles, p.n_tiles) : num_m_tiles * p.n_tiles;
    const int num_kb = p.R * p.S * p.cblocks;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        if (p.tma_out) prefetch_tmap(&tmap_y);
        if (p.tma_res) prefetch_tmap(&tmap_r);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(smem_u32(full_bar + i), 1);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(tfull_bar + i), 1);
            mbar_init(smem_u32(tempty_bar + i), (BLOCK_N >= 128 ? 8 : 4) * (PAIR ? 2 : 1));   // one arrive per working epilogue warp (of both CTAs)
        }
        for (int i = 0; i < EPI_SETS * NUM_EPI_WARPS; ++i) mbar_init(smem_u32(res_bar + i), 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) { tmem_alloc_pair(smem_u32(tmem_slot), TMEM_COLS); tmem_relinquish_pair(); }
        else { tmem_alloc(smem_u32(tmem_slot), TMEM_COLS); tmem_relinquish(); }
    }
    tcgen05_fence_before();
    if constexpr (PAIR) cluster_sync_all();          // the peer's barriers are initialised before anything is signalled on them
    else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (whole warp, elected lane issues) =====================
        {
            int stage = 0;
            uint32_t phase = 0;
            long long pw_empty = 0, pw_total = 0, ntl = 0;
            (void)pw_empty; (void)pw_total; (void)ntl;
#ifdef HN_PROFILE_ROLES
            const long long prod_t0 = clock64();
#endif
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                ++ntl;
                int nt, mt;
                decode_tile<PAIR>(p, tile, nt, mt);
                const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, img = mt / (p.tiles_w * p.tiles_h);
                const int w_base = tw * p.TW * p.wmul - p.pad_w, h_base = th * p.TH * p.hmul - p.pad;
                int kb = 0;
                for (int r = 0; r < p.R; ++r)
                    for (int s = 0; s < p.S; ++s)
                        for (int cb = 0; cb < p.cblocks; ++cb, ++kb) {
                            { HN_PROF_T0(); mbar_wait(smem_u32(empty_bar + stage), phase ^ 1); HN_PROF_ADD(pw_empty); }
                            const uint32_t fb = smem_u32(full_bar + stage);
                            const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                            if (elect_one()) {
                                if constexpr (PAIR) {
                                    // both CTAs' bytes are counted on the LEADER's barrier (its MMA warp consumes the stage of both)
                                    const uint32_t lfb = mapa_u32(fb, 0);
                                    if (rank == 0) mbar_expect_tx(fb, 2 * STAGE_BYTES);
                                    tma_load_4d_pair(sa, &tmap_a, lfb, cb * BK, w_base + s * p.dil, h_base + r * p.dil, img);
                                    tma_load_2d_pair(sa + A_STAGE_BYTES, &tmap_b, lfb, kb * BK, nt * BLOCK_N + (int)rank * B_ROWS);
                                } else {
                                    mbar_expect_tx(fb, STAGE_BYTES);
                                    tma_load_4d(sa, &tmap_a, fb, cb * BK, w_base + s * p.dil, h_base + r * p.dil, img);
                                    tma_load_2d(sa + A_STAGE_BYTES, &tmap_b, fb, kb * BK, nt * BLOCK_N);
                                }
                            }
                            __syncwarp();
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
            }
#ifdef HN_PROFILE_ROLES
            pw_total = clock64() - prod_t0;
            if (lane == 0) { HN_PROF_FLUSH(0, pw_empty); HN_PROF_FLUSH(1, pw_total); HN_PROF_FLUSH(9, ntl); HN_PROF_FLUSH(10, 1); }
#endif
        }
    } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (whole warp, elected lane issues; the leader CTA of a pair) =====================
        {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            long long mw_full = 0, mw_tempty = 0;
            (void)mw_full; (void)mw_tempty;
            const uint64_t desc_hi = make_kmajor_desc(0, BK * 2);                     // everything but the start address
            const uint32_t lo0 = (smem_u32(smem) & 0x3FFFFu) >> 4;                    // (all of the dynamic shared memory lies below 256 KB)
#ifdef HN_PROFILE_ROLES
            const long long mma_t0 = clock64();
#endif
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                {
                    HN_PROF_T0();
                    if constexpr (PAIR) mbar_wait_cluster(smem_u32(tempty_bar + acc), acc_phase ^ 1);
                    else mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);
                    HN_PROF_ADD(mw_tempty);
                }
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                for (int kb = 0; kb < num_kb; ++kb) {
                    { HN_PROF_T0(); mbar_wait(smem_u32(full_bar + stage), phase); HN_PROF_ADD(mw_full); }
                    tcgen05_fence_after();
                    // start-address fields from a base prepared outside the loops: nothing but adds between two groups of MMAs
                    const uint32_t a_lo = lo0 + (uint32_t)stage * (uint32_t)(STAGE_BYTES >> 4);
                    const uint64_t adesc = desc_hi | (uint64_t)a_lo;
                    const uint64_t bdesc = desc_hi | (uint64_t)(a_lo + (uint32_t)(A_STAGE_BYTES >> 4));
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            // advance 16 BF16 = 32 B along K inside the swizzle atom: +2 in the (addr >> 4) field
                            umma_bf16_t<PAIR>(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0);
                        }
                        umma_commit_t<PAIR>(smem_u32(empty_bar + stage));   // frees the smem slot (of both CTAs) when these MMAs retire
                        if (kb == num_kb - 1) umma_commit_t<PAIR>(smem_u32(tfull_bar + acc));   // accumulator ready for the epilogue(s)
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
#ifdef HN_PROFILE_ROLES
            if (lane == 0) { HN_PROF_FLUSH(2, mw_full); HN_PROF_FLUSH(3, mw_tempty); HN_PROF_FLUSH(4, clock64() - mma_t0); }
#endif
        }
    } else if (warp >= EPI_WARP0) {
        conv_epilogue<BLOCK_N, false, 1, PAIR, 2, RES2>(p, &tmap_y, &tmap_r, tmem_base, tfull_bar, tempty_bar, res_bar, epi_stage, s_shift, num_tiles, warp, lane);
    }

    tcgen05_fence_before();
    if constexpr (PAIR) cluster_sync_all();          // no CTA leaves while its peer can still signal its barriers / write its TMEM
    else __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ im2col gather
// A[m][k], k = (r*S+s)*Cin + c, zero for padding and for k >= R*S*Cin.
// grid.x = output rows (n*Ho + ho), grid.y = chunks of the row's Wo * kpad/8 work items; thread = 8 consecutive k.
__global__ void __launch_bounds__(256) im2col_bf16_kernel(const __nv_bfloat16 *__restrict__ x, int ldx, int N, int H, int W, int C,
                                                          int Ho, int Wo, int R, int S, int stride, int pad, int dil, int kpad,
                                                          __nv_bfloat16 *__restrict__ a)
{
    const int k8n = kpad / 8;
    const int K = R * S * C;
    const bool vec = (C % 8 == 0) && (ldx % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const int row = blockIdx.x;
    const int n = row / Ho, ho = row - n * Ho;
    const __nv_bfloat16 *xin = x + (int64_t)n * H * W * ldx;
    __nv_bfloat16 *arow = a + (int64_t)row * Wo * kpad;
    const int items = Wo * k8n;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < items; i += gridDim.y * blockDim.x) {
        const int wo = i / k8n;
        const int k0 = (i - wo * k8n) * 8;
        uint4 out = make_uint4(0, 0, 0, 0);
        if (vec) {
            if (k0 < K) {
                int tap = k0 / C, c = k0 - tap * C;
                int r = tap / S, s = tap - r * S;
                int hi = ho * stride - pad + r * dil, wi = wo * stride - pad + s * dil;
                if (hi >= 0 && hi < H && wi >= 0 && wi < W)
                    out = *reinterpret_cast<const uint4 *>(xin + ((int64_t)hi * W + wi) * ldx + c);
            }
        } else {
            __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(&out);
            int tap = k0 / C, c = k0 - tap * C;
            int r = tap / S, s = tap - r * S;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (k0 + j < K) {
                    int hi = ho * stride - pad + r * dil, wi = wo * stride - pad + s * dil;
                    if (hi >= 0 && hi < H && wi >= 0 && wi < W) o[j] = xin[((int64_t)hi * W + wi) * ldx + c];
                }
                if (++c == C) {      // next filter tap
                    c = 0;
                    if (++s == S) { s = 0; ++r; }
                }
            }
        }
        *reinterpret_cast<uint4 *>(arow + (int64_t)wo * kpad + k0) = out;
    }
}

// Small-Cin gather (the 7x7 stride-2 stems, Cin = 3 / 1 / 4; the critic on 13-channel logits): a CTA stages the R input rows
// that 64 consecutive output pixels of one output row need in shared memory (coalesced, zero padded), then assembles the
// K-major rows from shared memory and writes them as 16-byte chunks.  grid = (ceil(Wo/64), N*Ho).
constexpr int IM2COL_TW = 64;
// `pitch` = elements per staged pixel: ldx when the input's pixel stride is C itself or a power of two (the run of pixels a patch row
// needs is then copied without a division per element), else C (compacted, general strides).
__global__ void __launch_bounds__(256) im2col_smallc_kernel(const __nv_bfloat16 *__restrict__ x, int ldx, int N, int H, int W, int C, int Ho,
                                                            int Wo, int R, int S, int stride, int pad, int dil, int kpad, int pitch,
                                                            __nv_bfloat16 *__restrict__ a)
{
    extern __shared__ __nv_bfloat16 patch[];   // [R][ncols][pitch]
    const int row = blockIdx.y;
    const int n = row / Ho, ho = row - n * Ho;
    const int wo0 = blockIdx.x * IM2COL_TW;
    const int ncols = (IM2COL_TW - 1) * stride + (S - 1) * dil + 1;
    const int wi0 = wo0 * stride - pad;
    const __nv_bfloat16 *xin = x + (int64_t)n * H * W * ldx;
    const int per_row = ncols * pitch;
    const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
    if (pitch == ldx) {
        // the pixels a patch row needs are one contiguous run of the input row: only its valid range [e_lo, e_hi) (columns left /
        // right of the image are zero padding) and, for padded pixels, the channel index (a mask: pitch is a power of two) matter
        const int e_lo = (wi0 < 0 ? -wi0 : 0) * pitch;
        const int e_hi = (W - wi0 < ncols ? (W - wi0 > 0 ? W - wi0 : 0) : ncols) * pitch;
        const bool dense = pitch == C;
        for (int r = 0; r < R; ++r) {
            const int hi = ho * stride - pad + r * dil;
            const bool row_ok = hi >= 0 && hi < H;
            const __nv_bfloat16 *src = xin + ((int64_t)hi * W + wi0) * ldx;
            __nv_bfloat16 *dst = patch + r * per_row;
            for (int e = threadIdx.x; e < per_row; e += blockDim.x)
                dst[e] = (row_ok && e >= e_lo && e < e_hi && (dense || (e & (pitch - 1)) < C)) ? src[e] : zero;
        }
    } else {
        for (int i = threadIdx.x; i < R * per_row; i += blockDim.x) {
            const int r = i / per_row, rem = i - r * per_row;
            const int col = rem / C, c = rem - col * C;
            const int hi = ho * stride - pad + r * dil, wi = wi0 + col;
            __nv_bfloat16 v = zero;
            if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = xin[((int64_t)hi * W + wi) * ldx + c];
            patch[i] = v;
        }
    }
    __syncthreads();
    const int k8n = kpad / 8;
    const int K = R * S * C;
    const int npx = min(IM2COL_TW, Wo - wo0);
    __nv_bfloat16 *arow = a + ((int64_t)row * Wo + wo0) * kpad;
    for (int i = threadIdx.x; i < npx * k8n; i += blockDim.x) {
        const int pl = i / k8n, k0 = (i - pl * k8n) * 8;
        uint4 out = make_uint4(0, 0, 0, 0);
        __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(&out);
        int tap = k0 / C, c = k0 - tap * C;
        int r = tap / S, s = tap - r * S;
        int idx = r * per_row + (pl * stride + s * dil) * pitch + c;     // walks (c, s, r) with increments only
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (k0 + j < K) o[j] = patch[idx];
            ++idx;
            if (++c == C) {
                c = 0;
                idx += dil * pitch - C;
                if (++s == S) {
                    s = 0;
                    idx += per_row - S * dil * pitch;
                }
            }
        }
        *reinterpret_cast<uint4 *>(arow + (int64_t)pl * kpad + k0) = out;
    }
}

int im2col_bf16(const hn_tensor *x, const hn_conv *cv, int Ho, int Wo, int kpad, void *ws, cudaStream_t st)
{
    HN_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 127) == 0, "im2col: workspace must be 128-byte aligned");
    if (x->c < 8 || x->c % 8 != 0) {
        const int ncols = (IM2COL_TW - 1) * cv->stride + (cv->s - 1) * cv->dil + 1;
        const bool pow2 = (x->ld & (x->ld - 1)) == 0 && x->ld <= 64;
        const int pitch = (x->ld == x->c || pow2) ? x->ld : x->c;
        const size_t smem = (size_t)cv->r * ncols * pitch * 2;
        if (smem <= 48 * 1024 && (int64_t)x->n * Ho <= 65535) {
            dim3 grid((unsigned)cdiv(Wo, IM2COL_TW), (unsigned)(x->n * Ho));
            im2col_smallc_kernel<<<grid, 256, smem, st>>>((const __nv_bfloat16 *)x->ptr, x->ld, x->n, x->h, x->w, x->c, Ho, Wo, cv->r, cv->s,
                                                       cv->stride, cv->pad, cv->dil, kpad, pitch, (__nv_bfloat16 *)ws);
            HN_LAUNCH_CHECK();
            return HN_OK;
        }
    }
    int64_t chunks = cdiv((int64_t)Wo * (kpad / 8), 1024);
    dim3 grid((unsigned)(x->n * Ho), (unsigned)(chunks < 1 ? 1 : (chunks > 65535 ? 65535 : chunks)));
    im2col_bf16_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x->ptr, x->ld, x->n, x->h, x->w, x->c, Ho, Wo, cv->r, cv->s, cv->stride,
                                            cv->pad, cv->dil, kpad, (__nv_bfloat16 *)ws);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// BF16 tensor map over up to 4 dims (innermost first), 128B swizzle, zero OOB fill
int make_tmap(CUtensorMap *m, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box,
              CUtensorMapDataType dtype, CUtensorMapSwizzle swizzle, const uint32_t *elem_strides)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return HN_ERR_CUDA;
    }
    cuuint64_t gd[4];
    cuuint64_t gs[3];
    cuuint32_t bx[4], es[4];
    for (int i = 0; i < rank; ++i) {
        gd[i] = dims[i];
        bx[i] = box[i];
        es[i] = elem_strides ? elem_strides[i] : 1;
        if (i > 0) gs[i - 1] = strides_bytes[i];
    }
    CUresult r = enc(m, dtype, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return HN_ERR_CUDA;
    }
    return HN_OK;
}

// Implicit GEMM straight from the NHWC input: stride 1, and stride 2 through a tensor map whose box traverses the pixel axes with
// element stride 2 (the TMA unit then delivers every other pixel of every other row: no im2col workspace for the critics' 4x4 s2
// convolutions, layer2.0's 3x3 s2 / 1x1 s2).  HN_NO_STRIDED_TMA=1 falls back to the im2col gather.
static bool implicit_ok(const hn_tensor *x, const hn_conv *cv)
{
    static const bool no_strided = getenv("HN_NO_STRIDED_TMA") != nullptr;
    return (cv->stride == 1 || (cv->stride == 2 && !no_strided)) && x->c % 64 == 0 && x->ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x->ptr) & 15) == 0;
}
bool conv_implicit_ok(const hn_tensor *x, const hn_conv *cv) { return implicit_ok(x, cv); }

int64_t conv2d_tc_workspace(const hn_tensor *x, const hn_conv *cv)
{
    if (implicit_ok(x, cv)) return 0;
    const int Ho = (x->h + 2 * cv->pad - cv->dil * (cv->r - 1) - 1) / cv->stride + 1;
    const int Wo = (x->w + 2 * cv->pad - cv->dil * (cv->s - 1) - 1) / cv->stride + 1;
    return (int64_t)x->n * Ho * Wo * hn_conv_kpad(x->c, cv->r, cv->s) * 2;
}

// CTA pairs pay off twice: a tile too narrow for one instruction to keep a tensor pipe busy (N <= 128) issues at up to twice the rate,
// and every tile width stages only half of the weight rows per SM (N = 256: a third less L2 -> shared-memory traffic per k-block, one
// more pipeline stage; measured +9 % on the 3x3 / deep 1x1 layers).  Needs two M tiles.  HN_NO_PAIR=1 keeps every launch on single
// CTAs (A/B measurements, bisecting), HN_PAIR_MAX_BN / HN_PAIR_MIN_KB move the thresholds.
// Not for short reductions (1x1 layers: a tile is a handful of instructions and the kernel is bound by HBM and by the tile hand-over
// between the roles, which the pair's cross-CTA signalling lengthens: measured 15-50 % slower) -- num_kb >= 9 k-blocks per tile.
bool conv_pair_ok(int bn, int num_m_tiles, int num_kb, bool halo)
{
    static const bool disabled = getenv("HN_NO_PAIR") != nullptr;
    static const bool tc_disabled = getenv("HN_NO_PAIR_TC") != nullptr;          // generic kernel only
    static const int min_kb = getenv("HN_PAIR_MIN_KB") ? atoi(getenv("HN_PAIR_MIN_KB")) : 9;
    static const int max_bn = getenv("HN_PAIR_MAX_BN") ? atoi(getenv("HN_PAIR_MAX_BN")) : 256;
    if (disabled || (tc_disabled && !halo)) return false;
    return bn >= 32 && bn <= max_bn && num_m_tiles >= 2 && (halo || num_kb >= min_kb);
}

// clusters of two CTAs that can be resident at once (one CTA per SM: normally num_sms / 2)
template <typename K>
static int max_pairs(K kernel, size_t smem, int threads)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2, 1, 1);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = num_sms() / 2;
    }
    return n;
}

template <int BN, int STAGES, int BK = BLOCK_K, bool PAIR = false, bool RES2 = false>
static int launch_tc(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &ty, const CUtensorMap &tr, const TcParams &p,
                     int num_m_tiles, cudaStream_t st)
{
    constexpr int EPI_SETS = RES2 ? 2 : 1;
    constexpr size_t smem = (size_t)STAGES * (BLOCK_M * BK * 2 + (PAIR ? BN / 2 : BN) * BK * 2) + EPI_SETS * NUM_EPI_WARPS * EPI_STAGE_BYTES +
                            (2 * STAGES + 4 + EPI_SETS * NUM_EPI_WARPS) * 8 + 16 + (BN < 128 ? 2 : 1) * BN * 4 + 1024;   // narrow tiles: one shift table per epilogue group
    static_assert(smem <= 227 * 1024, "shared memory budget exceeded");
    static bool configured_dev[HN_MAX_DEVICES] = {};
    static int pairs_dev[HN_MAX_DEVICES] = {};
    bool &configured = configured_dev[current_device()];
    int &pairs = pairs_dev[current_device()];
    if (!configured) {
        HN_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, BK, PAIR, RES2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (PAIR) pairs = max_pairs(conv_tc_kernel<BN, STAGES, BK, PAIR, RES2>, smem, NUM_THREADS);
        configured = true;
    }
    if constexpr (PAIR) {
        const int want = ((num_m_tiles + 1) / 2) * p.n_tiles;          // (M-tile pair, Cout tile) work items
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (unsigned)(want < pairs ? want : pairs), 1, 1);
        cfg.blockDim = dim3(NUM_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        HN_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, STAGES, BK, PAIR, RES2>, ta, tb, ty, tr, p));
    } else {
        const int num_tiles = num_m_tiles * p.n_tiles;
        int grid = num_tiles < num_sms() ? num_tiles : num_sms();
        conv_tc_kernel<BN, STAGES, BK, PAIR, RES2><<<grid, NUM_THREADS, smem, st>>>(ta, tb, ty, tr, p);
    }
    HN_LAUNCH_CHECK();
    return HN_OK;
}

// the Cout-tile menu of the kernel; `pair` selects the two-CTA form (the weight map's box then holds bn / 2 rows)
static int launch_tc_bn(int bn, bool pair, const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &ty, const CUtensorMap &tr,
                        const TcParams &p, int num_m_tiles, cudaStream_t st)
{
    if (pair) {
        switch (bn) {
            case 256: return launch_tc<256, 5, BLOCK_K, true>(ta, tb, ty, tr, p, num_m_tiles, st);
            case 128: return launch_tc<128, 6, BLOCK_K, true>(ta, tb, ty, tr, p, num_m_tiles, st);
            case 64: return launch_tc<64, 8, BLOCK_K, true>(ta, tb, ty, tr, p, num_m_tiles, st);
            case 32: return launch_tc<32, 8, BLOCK_K, true>(ta, tb, ty, tr, p, num_m_tiles, st);
        }
    } else {
        switch (bn) {
            case 256: {
                // single-CTA 256-wide tiles are the short-reduction 1x1 layers (bound by HBM and the epilogue): one operand stage less,
                // two staging tiles per epilogue warp (RES2).  HN_NO_RES2=1 keeps the 4-stage / one-tile form for A/B runs.
                static const bool no_res2 = getenv("HN_NO_RES2") != nullptr;
                if (no_res2) return launch_tc<256, 4>(ta, tb, ty, tr, p, num_m_tiles, st);
                return launch_tc<256, 3, BLOCK_K, false, true>(ta, tb, ty, tr, p, num_m_tiles, st);
            }
            case 128: return launch_tc<128, 6>(ta, tb, ty, tr, p, num_m_tiles, st);
            case 64: return launch_tc<64, 7>(ta, tb, ty, tr, p, num_m_tiles, st);
            case 32: return launch_tc<32, 8>(ta, tb, ty, tr, p, num_m_tiles, st);
            case 16: return launch_tc<16, 8>(ta, tb, ty, tr, p, num_m_tiles, st);
        }
    }
    set_error("conv_tc: unsupported Cout tile %d", bn);
    return HN_ERR_ARG;
}

// ------------------------------------------------------------------------------------------------ 1x1 with a handful of input channels
// y[p][co] = act(sum_c x[p][c] * w[co][c] * scale[co] + shift[co] + res[p][co]) for Cin <= 16 (the dgrad of the 13-class classifier:
// 13 -> 64 channels at full resolution).  Padding K to a 64-wide k-block costs an im2col pass that writes and re-reads 5x the
// output; at 13 MACs per output this is bandwidth-bound CUDA-core work instead: a thread owns 8 output channels of one pixel,
// the (scaled, transposed) FP32 filter sits in shared memory.
template <typename OutT>
__global__ void __launch_bounds__(256) conv1x1_smallk_kernel(const __nv_bfloat16 *__restrict__ x, int ldx, int C, int64_t M,
                                                             const __nv_bfloat16 *__restrict__ w, int kpad, int cout,
                                                             const float *__restrict__ scale, const float *__restrict__ shift,
                                                             const __nv_bfloat16 *__restrict__ res, int ldr, int act, float slope,
                                                             const float *__restrict__ slope_ptr, OutT *__restrict__ y, int ldy)
{
    extern __shared__ float wsm[];                     // [C][cout8] filter, then [cout8] shift
    const int cout8 = (cout + 7) & ~7, chunks = cout8 >> 3;
    float *shs = wsm + C * cout8;
    for (int i = threadIdx.x; i < C * cout8; i += blockDim.x) {
        const int c = i / cout8, co = i - c * cout8;
        wsm[i] = co < cout ? __bfloat162float(w[(int64_t)co * kpad + c]) * (scale ? __ldg(scale + co) : 1.f) : 0.f;
    }
    for (int i = threadIdx.x; i < cout8; i += blockDim.x) shs[i] = (shift && i < cout) ? __ldg(shift + i) : 0.f;
    __syncthreads();
    if (slope_ptr) slope = __ldg(slope_ptr);
    const bool xvec = (ldx & 7) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    const bool yvec = (ldy * sizeof(OutT)) % 16 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
    const bool rvec = res && (ldr & 7) == 0 && (reinterpret_cast<uintptr_t>(res) & 15) == 0;
    // a WARP owns 32 consecutive pixels x one chunk of 8 output channels: every filter read is a shared-memory broadcast (with the
    // chunk index varying inside the warp the 128-bit reads were 2-way bank-conflicted and the kernel LSU-bound: 1.37 ms instead
    // of 0.3 for the classifier's dgrad)
    const int64_t pix_groups = (M + 31) >> 5;
    const int64_t items = pix_groups * chunks * 32;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < items; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t wg = idx >> 5;
        const int64_t pix = (wg / chunks) * 32 + (idx & 31);
        const int co0 = (int)(wg % chunks) * 8;
        if (pix >= M) continue;
        float xv[16];
        const __nv_bfloat16 *xp = x + pix * ldx;
        if (xvec) {           // 16-byte loads may run past C inside the pixel's own stride (ldx >= 8 * ceil(C / 8)); the tail is not used
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                if (v * 8 < C) {
                    float f8[8];
                    Vec8<__nv_bfloat16>::load(xp + v * 8, f8);
#pragma unroll
                    for (int j = 0; j < 8; ++j) xv[v * 8 + j] = f8[j];
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) xv[c] = c < C ? __bfloat162float(xp[c]) : 0.f;
        }
        float acc[8];
        {
            const float4 s0 = *reinterpret_cast<const float4 *>(shs + co0), s1 = *reinterpret_cast<const float4 *>(shs + co0 + 4);
            acc[0] = s0.x; acc[1] = s0.y; acc[2] = s0.z; acc[3] = s0.w; acc[4] = s1.x; acc[5] = s1.y; acc[6] = s1.z; acc[7] = s1.w;
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            if (c < C) {
                const float4 w0 = *reinterpret_cast<const float4 *>(wsm + c * cout8 + co0), w1 = *reinterpret_cast<const float4 *>(wsm + c * cout8 + co0 + 4);
                acc[0] = fmaf(xv[c], w0.x, acc[0]); acc[1] = fmaf(xv[c], w0.y, acc[1]); acc[2] = fmaf(xv[c], w0.z, acc[2]); acc[3] = fmaf(xv[c], w0.w, acc[3]);
                acc[4] = fmaf(xv[c], w1.x, acc[4]); acc[5] = fmaf(xv[c], w1.y, acc[5]); acc[6] = fmaf(xv[c], w1.z, acc[6]); acc[7] = fmaf(xv[c], w1.w, acc[7]);
            }
        }
        const bool full = co0 + 8 <= cout;
        if (res) {
            const __nv_bfloat16 *rp = res + pix * ldr + co0;
            if (rvec && full) {
                float r8[8];
                Vec8<__nv_bfloat16>::load(rp, r8);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += r8[j];
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (co0 + j < cout) acc[j] += __bfloat162float(rp[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = apply_act(acc[j], act, slope);
        OutT *yp = y + pix * ldy + co0;
        if (yvec && full) Vec8<OutT>::store(yp, acc);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (co0 + j < cout) yp[j] = from_f32<OutT>(acc[j]);
        }
    }
}

static bool smallk_ok(const hn_tensor *x, const hn_conv *cv, const hn_epilogue *ep)
{
    static const bool disabled = getenv("HN_NO_SMALLK") != nullptr;
    return !disabled && cv->r == 1 && cv->s == 1 && cv->stride == 1 && cv->pad == 0 && x->c <= 16 && cv->cout <= 1024 && !ep->stat_sum;
}

static int conv1x1_smallk(const hn_tensor *x, const void *w, const hn_conv *cv, const hn_epilogue *ep, const hn_tensor *y, cudaStream_t st)
{
    const int64_t M = (int64_t)x->n * x->h * x->w;
    const int kpad = hn_conv_kpad(x->c, 1, 1);
    const int cout8 = (cv->cout + 7) & ~7;
    const size_t smem = (size_t)(x->c + 1) * cout8 * 4;
    const int64_t items = cdiv(M, 32) * 32 * (cout8 / 8);
    int64_t blocks = cdiv(items, 256);
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    const __nv_bfloat16 *xp = (const __nv_bfloat16 *)x->ptr, *wp = (const __nv_bfloat16 *)w, *rp = (const __nv_bfloat16 *)ep->residual;
    if (y->dtype == HN_F32)
        conv1x1_smallk_kernel<float><<<(unsigned)blocks, 256, smem, st>>>(xp, x->ld, x->c, M, wp, kpad, cv->cout, ep->scale, ep->shift, rp, ep->residual_ld,
                                                                          ep->act, ep->slope, ep->slope_ptr, (float *)y->ptr, y->ld);
    else
        conv1x1_smallk_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, smem, st>>>(xp, x->ld, x->c, M, wp, kpad, cv->cout, ep->scale, ep->shift, rp,
                                                                                  ep->residual_ld, ep->act, ep->slope, ep->slope_ptr,
                                                                                  (__nv_bfloat16 *)y->ptr, y->ld);
    HN_LAUNCH_CHECK();
    return HN_OK;
}

int conv2d_fwd_tc(const hn_tensor *x, const void *w, const hn_conv *cv, const hn_epilogue *ep, const hn_tensor *y, void *ws,
                  int64_t ws_bytes, cudaStream_t st)
{
    const int Ho = y->h, Wo = y->w;
    const int64_t M = (int64_t)x->n * Ho * Wo;
    if (M == 0) return HN_OK;
    if (smallk_ok(x, cv, ep)) return conv1x1_smallk(x, w, cv, ep, y, st);
    if (conv_halo_ok(x, cv, y, false)) return conv2d_fwd_halo(x, w, cv, ep, y, false, st);   // 3x3 with narrow outputs: one halo patch per k-block
    const int kpad = hn_conv_kpad(x->c, cv->r, cv->s);
    const int cout_pad = hn_conv_cout_pad(cv->cout, HN_BF16);
    HN_CHECK_ARG((reinterpret_cast<uintptr_t>(w) & 15) == 0, "conv_tc: packed weights must be 16-byte aligned");

    TcParams p{};
    p.hmul = p.wmul = 1;
    CUtensorMap ta, tb;
    const bool implicit = implicit_ok(x, cv);
    const bool flat = !implicit || (cv->r == 1 && cv->s == 1 && cv->pad == 0 && cv->stride == 1);
    if (flat) {
        // A is a dense-or-strided [M][K] matrix: the NHWC view itself (1x1 s1) or the im2col workspace
        const void *abase = x->ptr;
        uint64_t lda = (uint64_t)x->ld;
        int cblocks = x->c / 64;
        if (!implicit) {
            const int64_t need = conv2d_tc_workspace(x, cv);
            if (!ws || ws_bytes < need) {
                set_error("conv_tc: workspace of %lld bytes required, %lld given", (long long)need, (long long)ws_bytes);
                return HN_ERR_WORKSPACE;
            }
            int rc0 = im2col_bf16(x, cv, Ho, Wo, kpad, ws, st);
            if (rc0) return rc0;
            abase = ws;
            lda = (uint64_t)kpad;
            cblocks = kpad / 64;
        }
        uint64_t dims[4] = {(uint64_t)cblocks * 64, (uint64_t)M, 1, 1};
        uint64_t strides[4] = {2, lda * 2, lda * 2, lda * 2};   // dims 2,3 have extent 1
        uint32_t box[4] = {64, 128, 1, 1};
        int rc = make_tmap(&ta, abase, 4, dims, strides, box);
        if (rc) return rc;
        p.tiles_w = (int)cdiv(M, 128); p.tiles_h = 1; p.n_img = 1;
        p.TH = 1; p.TW = 128; p.Ho = 1; p.Wo = (int)M;
        HN_CHECK_ARG(M < (int64_t)1 << 31, "conv_tc: too many pixels");
        p.R = 1; p.S = 1; p.pad = p.pad_w = 0; p.dil = 1; p.cblocks = cblocks;
    } else {
        // spatial tile menu: minimise padded area, prefer squarer tiles (halo reuse in L2)
        const int menu[5][2] = {{8, 16}, {4, 32}, {16, 8}, {2, 64}, {1, 128}};
        int best = 0;
        int64_t best_area = -1;
        for (int i = 0; i < 5; ++i) {
            int64_t area = cdiv(Ho, menu[i][0]) * menu[i][0] * cdiv(Wo, menu[i][1]) * menu[i][1];
            if (best_area < 0 || area < best_area) { best_area = area; best = i; }
        }
        p.TH = menu[best][0]; p.TW = menu[best][1];
        uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
        uint64_t strides[4] = {2, (uint64_t)x->ld * 2, (uint64_t)x->ld * 2 * x->w, (uint64_t)x->ld * 2 * x->w * x->h};
        const uint32_t st_ = (uint32_t)cv->stride;
        uint32_t box[4] = {64, (uint32_t)p.TW * st_, (uint32_t)p.TH * st_, 1};      // traversed extent: TW x TH pixels at stride st_
        uint32_t es[4] = {1, st_, st_, 1};
        int rc = make_tmap(&ta, x->ptr, 4, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B, es);
        if (rc) return rc;
        p.tiles_w = (int)cdiv(Wo, p.TW); p.tiles_h = (int)cdiv(Ho, p.TH); p.n_img = x->n;
        p.Ho = Ho; p.Wo = Wo;
        p.hmul = p.wmul = cv->stride;
        p.R = cv->r; p.S = cv->s; p.pad = p.pad_w = cv->pad; p.dil = cv->dil; p.cblocks = x->c / 64;
    }
    const int num_m_tiles = p.n_img * p.tiles_h * p.tiles_w;
    // Cout tile: the widest of 256/128/64 that divides cout_pad and still leaves >= 2 tiles per SM
    int bn;
    if (cout_pad < 64) bn = cout_pad;   // 16 or 32
    else {
        bn = 64;
        for (int cand : {256, 128}) {
            if (cout_pad % cand == 0 && (int64_t)num_m_tiles * (cout_pad / cand) >= 2 * (int64_t)num_sms()) { bn = cand; break; }
        }
    }
    {   // tuning knobs for experiments: HN_BN_RES=<bn> caps the Cout tile of residual layers, HN_BN_MAX=<bn> of all layers
        static const int bn_res = getenv("HN_BN_RES") ? atoi(getenv("HN_BN_RES")) : 0;
        static const int bn_max = getenv("HN_BN_MAX") ? atoi(getenv("HN_BN_MAX")) : 0;
        if (bn_res && ep->residual && bn > bn_res && cout_pad % bn_res == 0) bn = bn_res;
        if (bn_max && bn > bn_max && cout_pad % bn_max == 0) bn = bn_max;
    }
    p.n_tiles = cout_pad / bn;
    p.Cout = cv->cout;
    const bool pair = conv_pair_ok(bn, num_m_tiles, p.R * p.S * p.cblocks, false);
    {
        uint64_t dims[2] = {(uint64_t)kpad, (uint64_t)cout_pad};
        uint64_t strides[2] = {2, (uint64_t)kpad * 2};
        uint32_t box[2] = {64, (uint32_t)(pair ? bn / 2 : bn)};
        int rc = make_tmap(&tb, w, 2, dims, strides, box);
        if (rc) return rc;
    }
    p.y = y->ptr; p.ldy = y->ld; p.y_f32 = (y->dtype == HN_F32);
    p.scale = ep->scale; p.shift = ep->shift; p.res = ep->residual; p.ldr = ep->residual_ld;
    p.act = ep->act; p.slope = ep->slope; p.slope_ptr = ep->slope_ptr;
    p.stat_sum = ep->stat_sum; p.stat_sqsum = ep->stat_sqsum;
    if (int rcg = set_stat_groups(p, ep, y)) return rcg;
    // epilogue tensor maps: the output view (and the residual view) as {C, W, H, N} with a {128 B, ebw, 32/ebw, 1} box
    CUtensorMap ty, tr;
    memset(&ty, 0, sizeof(ty));
    memset(&tr, 0, sizeof(tr));
    p.ebw = p.TW < 32 ? p.TW : 32;
    {
        const uint64_t esz = p.y_f32 ? 4 : 2;
        const bool ok = (reinterpret_cast<uintptr_t>(y->ptr) & 15) == 0 && ((uint64_t)y->ld * esz) % 16 == 0;
        if (ok) {
            uint64_t dims[4] = {(uint64_t)cv->cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.n_img};
            uint64_t strides[4] = {esz, (uint64_t)y->ld * esz, (uint64_t)y->ld * esz * p.Wo, (uint64_t)y->ld * esz * p.Wo * p.Ho};
            uint32_t box[4] = {(uint32_t)(128 / esz), (uint32_t)p.ebw, (uint32_t)(32 / p.ebw), 1};
            int rc = make_tmap(&ty, y->ptr, 4, dims, strides, box, p.y_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
            if (rc) return rc;
            p.tma_out = 1;
            if (ep->residual && !p.y_f32 && (reinterpret_cast<uintptr_t>(ep->residual) & 15) == 0 && ((uint64_t)ep->residual_ld * 2) % 16 == 0) {
                uint64_t rs[4] = {2, (uint64_t)ep->residual_ld * 2, (uint64_t)ep->residual_ld * 2 * p.Wo, (uint64_t)ep->residual_ld * 2 * p.Wo * p.Ho};
                rc = make_tmap(&tr, ep->residual, 4, dims, rs, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
                if (rc) return rc;
                p.tma_res = 1;
            }
        }
    }
    return launch_tc_bn(bn, pair, ta, tb, ty, tr, p, num_m_tiles, st);
}


// ------------------------------------------------------------------------------------------------ phases of a strided dgrad
// One parity class (rho_y, rho_x) of the input gradient of a stride-2 convolution is a stride-1 correlation of dY with the
// sub-filter of the taps that can reach that parity (hn_backward.cu: hn_conv2d_dgrad_s2).  It runs on the implicit-GEMM path
// above with two generalisations: per-axis padding, and an output (and accumulate-residual) tensor map whose pixel / row strides
// address the phase's sub-lattice of dX directly, so no zero-inserted gradient and no scatter pass exist.
int conv2d_fwd_tc_sub(const hn_tensor *x, const void *w, int cout, const TcSubConv *sc, cudaStream_t st)
{
    const int Ho = sc->out_h, Wo = sc->out_w;
    const int64_t M = (int64_t)x->n * Ho * Wo;
    if (M == 0) return HN_OK;
    HN_CHECK_ARG(x->dtype == HN_BF16 && x->c % 64 == 0 && x->ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x->ptr) & 15) == 0,
                 "conv_tc_sub: input must be a 16-byte aligned BF16 view with C %% 64 == 0");
    const uint64_t esz = 2;
    HN_CHECK_ARG((reinterpret_cast<uintptr_t>(sc->y) & 15) == 0 && (sc->pix_stride * esz) % 16 == 0 && (sc->row_stride * esz) % 16 == 0 &&
                     (sc->img_stride * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                 "conv_tc_sub: output lattice must be 16-byte aligned");
    const int kpad = hn_conv_kpad(x->c, sc->R, sc->S);
    const int cout_pad = hn_conv_cout_pad(cout, HN_BF16);
    TcParams p{};
    p.hmul = p.wmul = 1;
    const int menu[5][2] = {{8, 16}, {4, 32}, {16, 8}, {2, 64}, {1, 128}};
    int best = 0;
    int64_t best_area = -1;
    for (int i = 0; i < 5; ++i) {
        int64_t area = cdiv(Ho, menu[i][0]) * menu[i][0] * cdiv(Wo, menu[i][1]) * menu[i][1];
        if (best_area < 0 || area < best_area) { best_area = area; best = i; }
    }
    p.TH = menu[best][0]; p.TW = menu[best][1];
    CUtensorMap ta, tb, ty, tr;
    memset(&tr, 0, sizeof(tr));
    {
        uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
        uint64_t strides[4] = {2, (uint64_t)x->ld * 2, (uint64_t)x->ld * 2 * x->w, (uint64_t)x->ld * 2 * x->w * x->h};
        uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
        int rc = make_tmap(&ta, x->ptr, 4, dims, strides, box);
        if (rc) return rc;
    }
    p.tiles_w = (int)cdiv(Wo, p.TW); p.tiles_h = (int)cdiv(Ho, p.TH); p.n_img = x->n;
    p.Ho = Ho; p.Wo = Wo;
    p.R = sc->R; p.S = sc->S; p.pad = sc->pad_h; p.pad_w = sc->pad_w; p.dil = 1; p.cblocks = x->c / 64;
    const int num_m_tiles = p.n_img * p.tiles_h * p.tiles_w;
    int bn;
    if (cout_pad < 64) bn = cout_pad;
    else {
        bn = 64;
        for (int cand : {256, 128}) {
            if (cout_pad % cand == 0 && (int64_t)num_m_tiles * (cout_pad / cand) >= 2 * (int64_t)num_sms()) { bn = cand; break; }
        }
    }
    p.n_tiles = cout_pad / bn;
    p.Cout = cout;
    const bool pair = conv_pair_ok(bn, num_m_tiles, p.R * p.S * p.cblocks, false);
    {
        uint64_t dims[2] = {(uint64_t)kpad, (uint64_t)cout_pad};
        uint64_t strides[2] = {2, (uint64_t)kpad * 2};
        uint32_t box[2] = {64, (uint32_t)(pair ? bn / 2 : bn)};
        int rc = make_tmap(&tb, w, 2, dims, strides, box);
        if (rc) return rc;
    }
    p.y = sc->y; p.ldy = 0; p.y_f32 = 0;
    p.act = HN_ACT_NONE;
    p.ebw = p.TW < 32 ? p.TW : 32;
    {
        uint64_t dims[4] = {(uint64_t)cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)p.n_img};
        uint64_t strides[4] = {esz, (uint64_t)sc->pix_stride * esz, (uint64_t)sc->row_stride * esz, (uint64_t)sc->img_stride * esz};
        uint32_t box[4] = {64, (uint32_t)p.ebw, (uint32_t)(32 / p.ebw), 1};
        int rc = make_tmap(&ty, sc->y, 4, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
        if (rc) return rc;
        p.tma_out = 1;
        if (sc->accumulate) {          // dX += ...: the lattice itself is the epilogue's residual input
            rc = make_tmap(&tr, sc->y, 4, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
            if (rc) return rc;
            p.res = sc->y; p.ldr = 0; p.tma_res = 1;
        }
    }
    return launch_tc_bn(bn, pair, ta, tb, ty, tr, p, num_m_tiles, st);
}


// ------------------------------------------------------------------------------------------------ 7x7 stride-2 stems
// The stems (Cin = 3 / 1 / 4) without an im2col pass: the input is kept as a zero-bordered, 4-channel NHWC image
// xpad[N][Hp][Wp][4] (border 3 = the conv padding).  The 8 pixels x 4 channels = 32 BF16 that one filter row touches for
// output pixel wo are CONTIGUOUS in it (they start at padded pixel 2*wo), and the windows of consecutive output pixels
// overlap by 6 pixels -- which a tiled tensor map expresses directly: dims {32, Wo, Hp, N} with a 16-byte stride between
// windows.  One TMA box {32, 128, 1, 1} is the A tile of filter row r for 128 output pixels (64-byte rows, SWIZZLE_64B);
// K = 7 rows x 32 (28 real taps + 4 zero) per output.  The weights are packed [Cout][7][8][4] to match.
int conv_stem_tc(const hn_tensor *xpad, const void *w, int cout, const hn_epilogue *ep, const hn_tensor *y, cudaStream_t st)
{
    const int Ho = y->h, Wo = y->w;
    if ((int64_t)y->n * Ho * Wo == 0) return HN_OK;
    HN_CHECK_ARG(xpad->c == 4 && xpad->ld == 4 && xpad->dtype == HN_BF16, "conv_stem: input must be the padded 4-channel BF16 image");
    HN_CHECK_ARG(xpad->w >= 2 * Wo + 6 && xpad->h >= 2 * Ho + 5, "conv_stem: padded image too small (%dx%d for output %dx%d)", xpad->h, xpad->w, Ho, Wo);
    HN_CHECK_ARG((reinterpret_cast<uintptr_t>(xpad->ptr) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0, "conv_stem: 16-byte alignment");
    const int cout_pad = hn_conv_cout_pad(cout, HN_BF16);
    HN_CHECK_ARG(cout_pad == 64, "conv_stem: Cout must be 64 (got %d)", cout);
    constexpr int BK = 32, KTOT = 7 * BK;
    TcParams p{};
    p.hmul = 2; p.wmul = 1;
    p.TH = 1; p.TW = 128;
    p.tiles_w = (int)cdiv(Wo, 128); p.tiles_h = Ho; p.n_img = y->n;
    p.Ho = Ho; p.Wo = Wo;
    p.R = 7; p.S = 1; p.pad = p.pad_w = 0; p.dil = 1; p.cblocks = 1;
    p.n_tiles = 1; p.Cout = cout;
    CUtensorMap ta, tb, ty, tr;
    memset(&ty, 0, sizeof(ty));
    memset(&tr, 0, sizeof(tr));
    {
        const uint64_t rowb = (uint64_t)xpad->w * 8;
        uint64_t dims[4] = {32, (uint64_t)Wo, (uint64_t)xpad->h, (uint64_t)xpad->n};
        uint64_t strides[4] = {2, 16, rowb, rowb * xpad->h};
        uint32_t box[4] = {32, 128, 1, 1};
        int rc = make_tmap(&ta, xpad->ptr, 4, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)KTOT, (uint64_t)cout_pad};
        uint64_t strides[2] = {2, (uint64_t)KTOT * 2};
        uint32_t box[2] = {32, 64};
        int rc = make_tmap(&tb, w, 2, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
    }
    p.y = y->ptr; p.ldy = y->ld; p.y_f32 = (y->dtype == HN_F32);
    p.scale = ep->scale; p.shift = ep->shift; p.res = ep->residual; p.ldr = ep->residual_ld;
    p.act = ep->act; p.slope = ep->slope; p.slope_ptr = ep->slope_ptr;
    p.stat_sum = ep->stat_sum; p.stat_sqsum = ep->stat_sqsum;
    if (int rcg = set_stat_groups(p, ep, y)) return rcg;
    p.ebw = 32;
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
    HN_CHECK_ARG(!ep->residual, "conv_stem: no residual input");
    return launch_tc<64, 12, BK>(ta, tb, ty, tr, p, p.n_img * p.tiles_h * p.tiles_w, st);
}

}  // namespace hn
