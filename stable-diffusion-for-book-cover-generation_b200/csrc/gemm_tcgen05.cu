// b200sd -- GEMM / implicit-GEMM 3x3 convolution on the sm_100a tensor cores.
//
//   out[M, N] = A[M, K] * W[N, K]^T (+ bias) (+ rowbias) (+ residual)     bf16 x bf16 -> fp32 -> bf16
//
// Warp-specialised, one 128 x BLOCK_N output tile (x one K split) per CTA:
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier full/empty)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (accumulator in TMEM)
//   warps 2..5  : epilogue (tcgen05.ld TMEM -> registers -> fused bias/temb/residual/GEGLU -> global)
// 3x3 convolution is the same pipeline with the A operand gathered by 4-D TMA boxes over the NHWC
// activation: one (64 ch, W, rows, images) box per (tap, 64-channel block); the +-1 halo of the
// padding comes from TMA out-of-bounds zero fill, so no im2col buffer ever exists.  A second
// activation source (a1) extends the channel axis, which fuses the up-block torch.cat.
// With <= 113 KB of smem and <= 256 TMEM columns per CTA two CTAs share an SM, so one CTA's
// epilogue overlaps the other's main loop.  Small-M / deep-K layers are split along K with a
// fp32 partial tiles in an L2-resident workspace; the last CTA of a tile sums them in a fixed order
// (deterministic) and runs the epilogue.
#include <atomic>
#include <cstring>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

extern std::atomic<long long> g_b200sd_launches;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kABytes = BLOCK_M * BLOCK_K * 2;  // 16 KB (always reserved in full)

constexpr size_t kSplitWsBytes = 64ull << 20;  // fp32 partial tiles [tile][split][128][block_n]
constexpr int kMaxSplitTiles = 4096;           // per-tile arrival counters live after the partials

struct KParams {
    CUtensorMap tmA0, tmA1, tmB;
    const float* bias;
    const float* rowbias;
    const void* residual;
    void* out;
    float* ws_partials;
    unsigned int* ws_counters;
    int M, N;
    int num_k_blocks;    // total K / 64
    int kb_per_split;
    int split_k;
    int block_n;
    int stages;
    int conv;            // 0 plain, 1 conv3x3
    int cblocks0;        // C0 / 64
    int cblocks;         // (C0 + C1) / 64
    int tile_w, tile_h, tile_n;   // conv box geometry
    int tiles_y;         // H / tile_h (conv, tile_n == 1)
    int rows_valid;      // valid rows in a tile (<= 128)
    int a_bytes;         // bytes TMA delivers for A per stage
    int ldc, ldr, ldrb;
    int rows_per_image;
    int epilogue;
    int out_f32;
    int res_f32;
    uint32_t tmem_cols;
};

// Final epilogue for 8 consecutive output columns of one row: v = accumulators (fp32).
__device__ __forceinline__ void epilogue_store8(const KParams& p, int row, int col, float (&v)[8]) {
    if (p.bias) {
        const float4* b = reinterpret_cast<const float4*>(p.bias + col);
        const float4 t0 = __ldg(b), t1 = __ldg(b + 1);
        v[0] += t0.x; v[1] += t0.y; v[2] += t0.z; v[3] += t0.w;
        v[4] += t1.x; v[5] += t1.y; v[6] += t1.z; v[7] += t1.w;
    }
    if (p.rowbias) {
        const float4* b = reinterpret_cast<const float4*>(p.rowbias + (size_t)(row / p.rows_per_image) * p.ldrb + col);
        const float4 t0 = __ldg(b), t1 = __ldg(b + 1);
        v[0] += t0.x; v[1] += t0.y; v[2] += t0.z; v[3] += t0.w;
        v[4] += t1.x; v[5] += t1.y; v[6] += t1.z; v[7] += t1.w;
    }
    if (p.residual) {
        float r[8];
        if (p.res_f32) ld8<B200SD_F32>(p.residual, (size_t)row * p.ldr + col, r);
        else ld8<B200SD_BF16>(p.residual, (size_t)row * p.ldr + col, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += r[i];
    }
    if (p.out_f32) {
        float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) + (size_t)row * p.ldc + col);
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        uint4 u;
        u.x = pack_bf16x2(v[0], v[1]);
        u.y = pack_bf16x2(v[2], v[3]);
        u.z = pack_bf16x2(v[4], v[5]);
        u.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(static_cast<bf16*>(p.out) + (size_t)row * p.ldc + col) = u;
    }
}

__global__ void __launch_bounds__(kNumThreads, 1) gemm_tcgen05_kernel(const __grid_constant__ KParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_bytes = p.block_n * BLOCK_K * 2;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)p.stages * kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * b_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kMaxStages;
    uint64_t* tmem_full_bar = bars + 2 * kMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);
    uint32_t* last_flag = tmem_slot + 1;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x;
    const int n_tile = blockIdx.y;
    const int split = blockIdx.z;
    const int kb_begin = split * p.kb_per_split;
    const int kb_end = min(kb_begin + p.kb_per_split, p.num_k_blocks);
    const int n0 = n_tile * p.block_n;
    const int m0 = m_tile * p.rows_valid;

    ptx::pdl_trigger();
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&p.tmA0);
        ptx::prefetch_tmap(&p.tmB);
        if (p.cblocks0 != p.cblocks) ptx::prefetch_tmap(&p.tmA1);
        for (int s = 0; s < p.stages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_init(tmem_full_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, p.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::pdl_wait();  // everything above overlapped the previous kernel's tail

    if (warp == 0) {
        // ================= TMA producer =================
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int img = 0, y0 = 0;
            if (p.conv) {
                if (p.tile_n > 1) { img = m_tile * p.tile_n; y0 = 0; }
                else { img = m_tile / p.tiles_y; y0 = (m_tile % p.tiles_y) * p.tile_h; }
            }
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                ptx::mbar_expect_tx(&full_bar[stage], (uint32_t)(p.a_bytes + b_bytes));
                uint8_t* dst_a = smem_a + (size_t)stage * kABytes;
                uint8_t* dst_b = smem_b + (size_t)stage * b_bytes;
                if (!p.conv) {
                    if (kb < p.cblocks0) ptx::tma_load_2d(dst_a, &p.tmA0, &full_bar[stage], kb * BLOCK_K, m0);
                    else ptx::tma_load_2d(dst_a, &p.tmA1, &full_bar[stage], (kb - p.cblocks0) * BLOCK_K, m0);
                } else {
                    const int tap = kb / p.cblocks;
                    const int cb = kb - tap * p.cblocks;
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    if (cb < p.cblocks0)
                        ptx::tma_load_4d(dst_a, &p.tmA0, &full_bar[stage], cb * BLOCK_K, dx, y0 + dy, img);
                    else
                        ptx::tma_load_4d(dst_a, &p.tmA1, &full_bar[stage], (cb - p.cblocks0) * BLOCK_K, dx, y0 + dy, img);
                }
                ptx::tma_load_2d(dst_b, &p.tmB, &full_bar[stage], kb * BLOCK_K, n0);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one elected thread) =================
        if (ptx::elect_one()) {
            const uint32_t idesc = ptx::umma_idesc_bf16(BLOCK_M, (uint32_t)p.block_n);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                ptx::mbar_wait(&full_bar[stage], phase);
                ptx::tc_fence_after();
                const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(smem_a + (size_t)stage * kABytes));
                const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(smem_b + (size_t)stage * b_bytes));
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    // advance 32 B (= 16 bf16) along K inside the swizzle atom: +2 in 16-byte units
                    ptx::umma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
                }
                ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            ptx::umma_commit(tmem_full_bar);
        }
    } else {
        // ================= epilogue warps =================
        // Phase A: TMEM -> registers -> fp32 staging tile in the (now idle) smem ring, one row per thread.
        // Phase B: each warp re-reads ITS 32 rows with lanes running along the columns, so every global
        //          access (residual load, output store, split-K reduction) is a coalesced 16/32-byte vector.
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int r_in_tile = q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int epi_tid = threadIdx.x - 64;  // 0..127
        const int pitch_f = p.block_n + 4;     // +16 B: conflict-free 16-byte row-strided stores
        float* stage = reinterpret_cast<float*>(smem);
        const int tile_id = m_tile * gridDim.y + n_tile;

        ptx::mbar_wait(tmem_full_bar, 0);
        ptx::tc_fence_after();

        {
            float4* my_row = reinterpret_cast<float4*>(stage + (size_t)r_in_tile * pitch_f);
            for (int c = 0; c < p.block_n; c += 32) {
                uint32_t r0[16], r1[16];
                const bool two = (c + 16 < p.block_n);
                ptx::tmem_ld_32x32b_x16(taddr + c, r0);
                if (two) ptx::tmem_ld_32x32b_x16(taddr + c + 16, r1);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    my_row[c / 4 + i] = make_float4(__uint_as_float(r0[4 * i]), __uint_as_float(r0[4 * i + 1]),
                                                    __uint_as_float(r0[4 * i + 2]), __uint_as_float(r0[4 * i + 3]));
                if (two) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        my_row[c / 4 + 4 + i] = make_float4(__uint_as_float(r1[4 * i]), __uint_as_float(r1[4 * i + 1]),
                                                            __uint_as_float(r1[4 * i + 2]), __uint_as_float(r1[4 * i + 3]));
                }
            }
        }
        ptx::tc_fence_before();
        __syncwarp();

        const float* wstage = stage + (size_t)(q * 32) * pitch_f;
        const int rows_here = min(32, min(p.rows_valid, p.M - m0) - q * 32);  // valid rows of this warp (may be <= 0)
        bool do_final = true;
        const float* src = wstage;
        int src_pitch = pitch_f;

        if (p.split_k > 1) {
            // every split publishes its fp32 partial tile (coalesced, L2-resident); the last-arriving CTA
            // of the tile sums the splits in a FIXED order (deterministic) and runs the epilogue.
            float* mine = p.ws_partials + (((size_t)tile_id * p.split_k + split) * BLOCK_M + q * 32) * p.block_n;
            const int groups4 = p.block_n / 4;
            const int items = rows_here * groups4;
#pragma unroll 4
            for (int idx = lane; idx < items; idx += 32) {
                const int rl = idx / groups4, c4 = idx - rl * groups4;
                const float4 v = *reinterpret_cast<const float4*>(wstage + (size_t)rl * pitch_f + c4 * 4);
                __stcg(reinterpret_cast<float4*>(mine + (size_t)rl * p.block_n + c4 * 4), v);
            }
            __threadfence();
            ptx::named_bar_sync(1, 128);
            if (epi_tid == 0) {
                const unsigned int prev = atomicAdd(p.ws_counters + tile_id, 1u);
                const bool last = (prev == (unsigned int)p.split_k - 1);
                if (last) p.ws_counters[tile_id] = 0;  // self-cleaning
                __threadfence();
                *last_flag = last ? 1u : 0u;
            }
            ptx::named_bar_sync(1, 128);
            do_final = (*last_flag != 0);
            src = p.ws_partials + ((size_t)tile_id * p.split_k * BLOCK_M + q * 32) * p.block_n;
            src_pitch = p.block_n;
        }

        if (do_final && rows_here > 0) {
            const bool from_ws = p.split_k > 1;
            if (p.epilogue == B200SD_EPI_GEGLU) {
                const int half = p.block_n / 2;
                const int groups = half / 8;
                const int items = rows_here * groups;
#pragma unroll 2
                for (int idx = lane; idx < items; idx += 32) {
                    const int rl = idx / groups, c8 = idx - rl * groups;
                    const float* sp = src + (size_t)rl * src_pitch + c8 * 8;
                    const float4 a0 = *reinterpret_cast<const float4*>(sp), a1 = *reinterpret_cast<const float4*>(sp + 4);
                    const float4 g0 = *reinterpret_cast<const float4*>(sp + half), g1 = *reinterpret_cast<const float4*>(sp + half + 4);
                    const float4* bv = reinterpret_cast<const float4*>(p.bias + n0 + c8 * 8);
                    const float4* bg = reinterpret_cast<const float4*>(p.bias + n0 + half + c8 * 8);
                    const float4 bv0 = __ldg(bv), bv1 = __ldg(bv + 1), bg0 = __ldg(bg), bg1 = __ldg(bg + 1);
                    uint4 u;
                    u.x = pack_bf16x2((a0.x + bv0.x) * gelu_erf_f(g0.x + bg0.x), (a0.y + bv0.y) * gelu_erf_f(g0.y + bg0.y));
                    u.y = pack_bf16x2((a0.z + bv0.z) * gelu_erf_f(g0.z + bg0.z), (a0.w + bv0.w) * gelu_erf_f(g0.w + bg0.w));
                    u.z = pack_bf16x2((a1.x + bv1.x) * gelu_erf_f(g1.x + bg1.x), (a1.y + bv1.y) * gelu_erf_f(g1.y + bg1.y));
                    u.w = pack_bf16x2((a1.z + bv1.z) * gelu_erf_f(g1.z + bg1.z), (a1.w + bv1.w) * gelu_erf_f(g1.w + bg1.w));
                    const int row = m0 + q * 32 + rl;
                    *reinterpret_cast<uint4*>(static_cast<bf16*>(p.out) + (size_t)row * p.ldc + n_tile * half + c8 * 8) = u;
                }
            } else {
                const int groups = p.block_n / 8;
                const int items = rows_here * groups;
#pragma unroll 4
                for (int idx = lane; idx < items; idx += 32) {
                    const int rl = idx / groups, c8 = idx - rl * groups;
                    float v[8];
                    if (from_ws) {
                        const size_t split_stride = (size_t)BLOCK_M * p.block_n;
                        const float* ap = src + (size_t)rl * src_pitch + c8 * 8;
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = 0.f;
#pragma unroll 4
                        for (int sp_i = 0; sp_i < p.split_k; ++sp_i) {
                            const float4 a0 = __ldcg(reinterpret_cast<const float4*>(ap + sp_i * split_stride));
                            const float4 a1 = __ldcg(reinterpret_cast<const float4*>(ap + sp_i * split_stride) + 1);
                            v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
                            v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
                        }
                    } else {
                        const float* sp = src + (size_t)rl * src_pitch + c8 * 8;
                        const float4 a0 = *reinterpret_cast<const float4*>(sp), a1 = *reinterpret_cast<const float4*>(sp + 4);
                        v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
                    }
                    epilogue_store8(p, m0 + q * 32 + rl, n0 + c8 * 8, v);
                }
            }
        }
        ptx::tc_fence_before();
    }

    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

uint32_t pow2_cols(int n) {
    uint32_t c = 32;
    while ((int)c < n) c <<= 1;
    return c;
}

int pick_block_n(int N, int m_tiles, int epilogue) {
    // Legal tile widths: multiple of 16 (32 for GEGLU), <= 256, dividing N.  Prefer 160 (two CTAs per SM,
    // 90 % of the smem read bandwidth the MMA needs); when that leaves the machine more than half empty
    // (small-M layers) fall to a narrower tile so more CTAs stream operands concurrently.
    const int step = (epilogue == B200SD_EPI_GEGLU) ? 32 : 16;
    const int sms = b200sd_num_sms();
    int best = 0;
    for (int bn = 160; bn >= 64; bn -= step)
        if (N % bn == 0) { best = bn; break; }
    if (best == 0) {
        for (int bn = 256; bn >= step; bn -= step)
            if (N % bn == 0) { best = bn; break; }
        return best;
    }
    if (epilogue == B200SD_EPI_LINEAR && (long)m_tiles * (N / best) * 2 <= sms) {
        for (int bn = 80; bn < best; bn += step)  // narrowest tile (>= 80) that still fits one wave
            if (N % bn == 0 && (long)m_tiles * (N / bn) <= sms) { best = bn; break; }
    }
    return best;
}

}  // namespace

extern "C" size_t b200sd_gemm_workspace_bytes(void) { return kSplitWsBytes + kMaxSplitTiles * sizeof(unsigned int); }

extern "C" int b200sd_geglu_tile(int N) { return pick_block_n(N, 1, B200SD_EPI_GEGLU); }

extern "C" int b200sd_gemm(const b200sd_gemm_args* a, b200sd_stream_t stream) {
    B200SD_REQUIRE(a != nullptr, "gemm: null args");
    B200SD_REQUIRE(a->a0 && a->w && a->out, "gemm: null operand pointer");
    B200SD_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "gemm: non-positive dims M=%d N=%d K=%d", a->M, a->N, a->K);
    B200SD_REQUIRE(a->conv_taps == 1 || a->conv_taps == 9, "gemm: conv_taps must be 1 or 9");
    const int C0 = a->C0, C1 = a->a1 ? a->C1 : 0, C = C0 + C1;
    B200SD_REQUIRE(C0 > 0 && C0 % BLOCK_K == 0 && C1 % BLOCK_K == 0, "gemm: channel counts must be multiples of 64 (C0=%d C1=%d)", C0, C1);
    B200SD_REQUIRE(a->K == a->conv_taps * C, "gemm: K=%d != taps*C=%d", a->K, a->conv_taps * C);
    B200SD_REQUIRE(a->N % 16 == 0, "gemm: N=%d must be a multiple of 16", a->N);
    B200SD_REQUIRE(a->epilogue == B200SD_EPI_LINEAR || a->epilogue == B200SD_EPI_GEGLU, "gemm: bad epilogue");
    B200SD_REQUIRE(a->epilogue != B200SD_EPI_GEGLU || (a->bias && !a->residual && !a->rowbias && a->out_dtype == B200SD_BF16),
                   "gemm: GEGLU epilogue needs bias, bf16 out and no residual/rowbias");
    B200SD_REQUIRE(a->out_dtype == B200SD_BF16 || a->out_dtype == B200SD_F32, "gemm: bad out dtype");
    B200SD_REQUIRE(a->residual_dtype == B200SD_BF16 || a->residual_dtype == B200SD_F32, "gemm: bad residual dtype");
    B200SD_REQUIRE(a->ldc % 8 == 0 && (!a->residual || a->ldr % 8 == 0), "gemm: ldc/ldr must be multiples of 8");
    B200SD_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->a0) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(a->w) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->a1) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(a->rowbias) & 15) == 0,
                   "gemm: pointers must be 16-byte aligned");
    B200SD_REQUIRE(!a->rowbias || (a->rows_per_image > 0 && a->ldrb % 4 == 0), "gemm: rowbias needs rows_per_image and ldrb %% 4 == 0");

    KParams p;
    memset(&p, 0, sizeof(p));
    p.M = a->M;
    p.N = a->N;
    p.num_k_blocks = a->K / BLOCK_K;
    p.conv = a->conv_taps == 9;
    p.cblocks0 = C0 / BLOCK_K;
    p.cblocks = C / BLOCK_K;
    p.bias = a->bias;
    p.rowbias = a->rowbias;
    p.residual = a->residual;
    p.res_f32 = a->residual_dtype == B200SD_F32;
    p.out = a->out;
    p.ldc = a->ldc;
    p.ldr = a->ldr;
    p.ldrb = a->ldrb > 0 ? a->ldrb : a->N;
    p.rows_per_image = a->rows_per_image > 0 ? a->rows_per_image : 1;
    p.epilogue = a->epilogue;
    p.out_f32 = a->out_dtype == B200SD_F32;

    // ---- M tiling ----
    int m_tiles;
    if (!p.conv) {
        p.rows_valid = BLOCK_M;
        p.a_bytes = kABytes;
        p.tile_w = p.tile_h = p.tile_n = 1;
        p.tiles_y = 1;
        m_tiles = ceil_div(a->M, BLOCK_M);
    } else {
        const int H = a->H, W = a->W, NB = a->batch;
        B200SD_REQUIRE(NB > 0 && H > 0 && W > 0 && a->M == NB * H * W, "gemm(conv): M=%d != batch*H*W", a->M);
        B200SD_REQUIRE(W <= BLOCK_M, "gemm(conv): W=%d > 128 unsupported", W);
        const int max_h = BLOCK_M / W;
        if (H <= max_h) {
            p.tile_h = H;
            p.tile_n = max_h / H;
            if (p.tile_n < 1) p.tile_n = 1;
            if (p.tile_n > NB) p.tile_n = NB;
        } else {
            p.tile_n = 1;
            p.tile_h = 1;
            for (int h = max_h; h >= 1; --h)
                if (H % h == 0) { p.tile_h = h; break; }
        }
        p.tile_w = W;
        p.tiles_y = H / p.tile_h;
        p.rows_valid = p.tile_w * p.tile_h * p.tile_n;
        p.a_bytes = p.rows_valid * BLOCK_K * 2;
        m_tiles = (p.tile_n > 1) ? ceil_div(NB, p.tile_n) : NB * p.tiles_y;
    }

    // ---- N tiling ----
    int bn = a->block_n > 0 ? a->block_n : pick_block_n(a->N, m_tiles, a->epilogue);
    B200SD_REQUIRE(bn >= 16 && bn <= 256 && bn % 16 == 0 && a->N % bn == 0, "gemm: bad block_n %d for N=%d", bn, a->N);
    B200SD_REQUIRE(a->epilogue != B200SD_EPI_GEGLU || bn % 32 == 0, "gemm: GEGLU needs block_n %% 32 == 0");
    p.block_n = bn;
    const int n_tiles = a->N / bn;
    p.tmem_cols = pow2_cols(bn);

    // ---- split-K ----
    const int sms = b200sd_num_sms();
    int split = a->split_k;
    static const bool no_split = getenv("B200SD_SPLITK") && getenv("B200SD_SPLITK")[0] == '0';
    if (split <= 0) {
        split = 1;
        const int tiles = m_tiles * n_tiles;
        if (!no_split && a->epilogue == B200SD_EPI_LINEAR && tiles * 2 <= sms && p.num_k_blocks >= 32) {
            split = sms / tiles;
            const int max_by_k = p.num_k_blocks / 16;  // keep >= 16 k-blocks per split
            if (split > max_by_k) split = max_by_k;
            if (split > 8) split = 8;
            if (split < 1) split = 1;
        }
    }
    B200SD_REQUIRE(split == 1 || a->epilogue == B200SD_EPI_LINEAR, "gemm: split-K only with the linear epilogue");
    p.kb_per_split = ceil_div(p.num_k_blocks, split);
    split = ceil_div(p.num_k_blocks, p.kb_per_split);  // no empty splits
    p.split_k = split;
    if (split > 1) {
        const size_t need = (size_t)m_tiles * n_tiles * split * BLOCK_M * bn * sizeof(float);
        B200SD_REQUIRE(a->workspace != nullptr && a->workspace_bytes >= b200sd_gemm_workspace_bytes(), "gemm: split-K needs the workspace");
        B200SD_REQUIRE(need <= kSplitWsBytes && m_tiles * n_tiles <= kMaxSplitTiles, "gemm: split-K workspace too small (%zu B)", need);
        p.ws_partials = static_cast<float*>(a->workspace);
        p.ws_counters = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(a->workspace) + kSplitWsBytes);
    }

    // ---- pipeline depth ----
    // Grids that fit in one wave keep every k-block of a short K loop in flight (deep ring, 1 CTA/SM);
    // larger grids stay <= ~110 KB so two CTAs share an SM and overlap epilogue with main loop.
    const int stage_bytes = kABytes + bn * BLOCK_K * 2;
    const long total_ctas = (long)m_tiles * n_tiles * split;
    int stages;
    if (total_ctas <= sms || bn > 160) stages = (200 * 1024) / stage_bytes;
    else stages = (110 * 1024) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages > p.kb_per_split && total_ctas <= sms) stages = p.kb_per_split;
    const int min_stages = ceil_div(BLOCK_M * (bn + 4) * (int)sizeof(float), stage_bytes);  // epilogue staging tile
    if (stages < min_stages) stages = min_stages;
    if (stages < 2) stages = 2;
    p.stages = stages;
    B200SD_REQUIRE((size_t)stages * stage_bytes >= (size_t)BLOCK_M * (bn + 4) * sizeof(float), "gemm: smem ring smaller than the epilogue staging tile");
    const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/;
    B200SD_REQUIRE(smem_bytes <= 227 * 1024, "gemm: smem budget exceeded (%zu B)", smem_bytes);

    // ---- tensor maps ----
    if (!p.conv) {
        const uint64_t dimsA[2] = {(uint64_t)C0, (uint64_t)a->M};
        const uint64_t strA[2] = {0, (uint64_t)(a->lda0 > 0 ? a->lda0 : C0) * 2};
        const uint32_t boxA[2] = {BLOCK_K, BLOCK_M};
        int rc = b200sd_make_tmap(&p.tmA0, a->a0, 2, dimsA, strA, boxA, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        if (C1 > 0) {
            const uint64_t dimsA1[2] = {(uint64_t)C1, (uint64_t)a->M};
            const uint64_t strA1[2] = {0, (uint64_t)(a->lda1 > 0 ? a->lda1 : C1) * 2};
            rc = b200sd_make_tmap(&p.tmA1, a->a1, 2, dimsA1, strA1, boxA, CU_TENSOR_MAP_SWIZZLE_128B);
            if (rc) return rc;
        }
    } else {
        const uint32_t boxA[4] = {BLOCK_K, (uint32_t)p.tile_w, (uint32_t)p.tile_h, (uint32_t)p.tile_n};
        const uint64_t dims0[4] = {(uint64_t)C0, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->batch};
        const uint64_t str0[4] = {0, (uint64_t)C0 * 2, (uint64_t)a->W * C0 * 2, (uint64_t)a->H * a->W * C0 * 2};
        int rc = b200sd_make_tmap(&p.tmA0, a->a0, 4, dims0, str0, boxA, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        if (C1 > 0) {
            const uint64_t dims1[4] = {(uint64_t)C1, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->batch};
            const uint64_t str1[4] = {0, (uint64_t)C1 * 2, (uint64_t)a->W * C1 * 2, (uint64_t)a->H * a->W * C1 * 2};
            rc = b200sd_make_tmap(&p.tmA1, a->a1, 4, dims1, str1, boxA, CU_TENSOR_MAP_SWIZZLE_128B);
            if (rc) return rc;
        }
    }
    {
        const uint64_t dimsB[2] = {(uint64_t)a->K, (uint64_t)a->N};
        const uint64_t strB[2] = {0, (uint64_t)a->K * 2};
        const uint32_t boxB[2] = {BLOCK_K, (uint32_t)bn};
        int rc = b200sd_make_tmap(&p.tmB, a->w, 2, dimsB, strB, boxB, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }

    static size_t configured_smem = 0;
    if (smem_bytes > configured_smem) {
        B200SD_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        configured_smem = 227 * 1024;
    }
    dim3 grid(m_tiles, n_tiles, split);
    B200SD_CUDA(b200sd_launch(gemm_tcgen05_kernel, dim3(grid), dim3(kNumThreads), smem_bytes, static_cast<cudaStream_t>(stream), p));
    g_b200sd_launches.fetch_add(1, std::memory_order_relaxed);
    B200SD_LAUNCH_CHECK();
    return B200SD_OK;
}
