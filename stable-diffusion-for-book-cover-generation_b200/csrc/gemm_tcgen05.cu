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
// deterministic last-CTA reduction through an L2-resident fp32 workspace.
#include <atomic>
#include <cstring>
#include <cstdio>

#include "common.cuh"

extern std::atomic<long long> g_b200sd_launches;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kABytes = BLOCK_M * BLOCK_K * 2;  // 16 KB (always reserved in full)

constexpr size_t kSplitWsBytes = 64ull << 20;  // fp32 partial tiles
constexpr int kMaxSplitTiles = 4096;           // per-tile arrival counters live after the partials

struct KParams {
    CUtensorMap tmA0, tmA1, tmB;
    const float* bias;
    const float* rowbias;
    const bf16* residual;
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
    uint32_t tmem_cols;
};

__device__ __forceinline__ void epilogue_store16(const KParams& p, int row, int col, float (&v)[16]) {
    // v holds out[row, col .. col+15] before bias / residual
    if (p.bias) {
        const float4* b = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 t = __ldg(b + i);
            v[4 * i + 0] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
        }
    }
    if (p.rowbias) {
        const float4* b = reinterpret_cast<const float4*>(p.rowbias + (size_t)(row / p.rows_per_image) * p.ldrb + col);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 t = __ldg(b + i);
            v[4 * i + 0] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
        }
    }
    if (p.residual) {
        const uint4* r = reinterpret_cast<const uint4*>(p.residual + (size_t)row * p.ldr + col);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            uint4 u = __ldg(r + i);
            float2 f;
            f = unpack_bf16x2(u.x); v[8 * i + 0] += f.x; v[8 * i + 1] += f.y;
            f = unpack_bf16x2(u.y); v[8 * i + 2] += f.x; v[8 * i + 3] += f.y;
            f = unpack_bf16x2(u.z); v[8 * i + 4] += f.x; v[8 * i + 5] += f.y;
            f = unpack_bf16x2(u.w); v[8 * i + 6] += f.x; v[8 * i + 7] += f.y;
        }
    }
    if (p.out_f32) {
        float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) + (size_t)row * p.ldc + col);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
        uint4* o = reinterpret_cast<uint4*>(static_cast<bf16*>(p.out) + (size_t)row * p.ldc + col);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            uint4 u;
            u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
            u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
            u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
            u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
            o[i] = u;
        }
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
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int r_in_tile = q * 32 + lane;
        const int row = m0 + r_in_tile;
        const bool row_ok = (r_in_tile < p.rows_valid) && (row < p.M);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int epi_tid = threadIdx.x - 64;  // 0..127

        ptx::mbar_wait(tmem_full_bar, 0);
        ptx::tc_fence_after();

        bool do_final = true;
        if (p.split_k > 1) {
            // write fp32 partial tile, then the last-arriving CTA of this tile reduces all splits
            const int tile_id = m_tile * gridDim.y + n_tile;
            float* my = p.ws_partials + ((size_t)(tile_id * p.split_k + split) * BLOCK_M + r_in_tile) * p.block_n;
            for (int c = 0; c < p.block_n; c += 16) {
                uint32_t r[16];
                ptx::tmem_ld_32x32b_x16(taddr + c, r);
                ptx::tmem_ld_wait();
                float4* o = reinterpret_cast<float4*>(my + c);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    __stcg(o + i, make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                              __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
            }
            __threadfence();
            ptx::named_bar_sync(1, 128);
            if (epi_tid == 0) {
                unsigned int prev = atomicAdd(p.ws_counters + tile_id, 1u);
                const bool last = (prev == (unsigned int)p.split_k - 1);
                if (last) p.ws_counters[tile_id] = 0;  // self-cleaning
                __threadfence();
                *last_flag = last ? 1u : 0u;
            }
            ptx::named_bar_sync(1, 128);
            do_final = (*last_flag != 0);
        }

        if (do_final) {
            const int tile_id = m_tile * gridDim.y + n_tile;
            const int ncols = (p.epilogue == B200SD_EPI_GEGLU) ? p.block_n / 2 : p.block_n;
            for (int c = 0; c < ncols; c += 16) {
                float v[16];
                if (p.split_k > 1) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                    for (int s = 0; s < p.split_k; ++s) {
                        const float4* src = reinterpret_cast<const float4*>(
                            p.ws_partials + ((size_t)(tile_id * p.split_k + s) * BLOCK_M + r_in_tile) * p.block_n + c);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float4 t = __ldcg(src + i);
                            v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
                        }
                    }
                } else {
                    uint32_t r[16];
                    ptx::tmem_ld_32x32b_x16(taddr + c, r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
                }
                if (p.epilogue == B200SD_EPI_GEGLU) {
                    // columns [0, bn/2) = value, [bn/2, bn) = gate (weights interleaved per tile on the host)
                    uint32_t g[16];
                    ptx::tmem_ld_32x32b_x16(taddr + p.block_n / 2 + c, g);
                    ptx::tmem_ld_wait();
                    const float* bv = p.bias + n0 + c;
                    const float* bg = p.bias + n0 + p.block_n / 2 + c;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float val = v[i] + __ldg(bv + i);
                        float gate = __uint_as_float(g[i]) + __ldg(bg + i);
                        v[i] = val * gelu_erf_f(gate);
                    }
                    if (row_ok) {
                        const int col = n_tile * (p.block_n / 2) + c;
                        uint4* o = reinterpret_cast<uint4*>(static_cast<bf16*>(p.out) + (size_t)row * p.ldc + col);
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            uint4 u;
                            u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
                            u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
                            u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
                            u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
                            o[i] = u;
                        }
                    }
                } else if (row_ok) {
                    epilogue_store16(p, row, n0 + c, v);
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
    // Largest legal tile width (multiple of 16, <= 256, divides N; multiple of 32 for GEGLU).
    const int step = (epilogue == B200SD_EPI_GEGLU) ? 32 : 16;
    int best = 0;
    for (int bn = 256; bn >= step; bn -= step)
        if (N % bn == 0) { best = bn; break; }
    if (best == 0) return 0;
    // prefer <= 160 columns when it lets two CTAs share an SM and still fills the machine
    if (best > 160) {
        for (int bn = 160; bn >= 64; bn -= step)
            if (N % bn == 0) {
                (void)m_tiles;
                return bn;
            }
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
    p.residual = static_cast<const bf16*>(a->residual);
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
    if (split <= 0) {
        split = 1;
        const int tiles = m_tiles * n_tiles;
        if (a->epilogue == B200SD_EPI_LINEAR && tiles < sms && p.num_k_blocks >= 16) {
            split = sms / tiles;
            const int max_by_k = p.num_k_blocks / 8;  // keep >= 8 k-blocks per split
            if (split > max_by_k) split = max_by_k;
            if (split > 16) split = 16;
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

    // ---- pipeline depth: keep <= ~110 KB so two CTAs fit per SM when the tile allows it ----
    const int stage_bytes = kABytes + bn * BLOCK_K * 2;
    int stages = (bn <= 160) ? (110 * 1024) / stage_bytes : (200 * 1024) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) stages = 2;
    p.stages = stages;
    const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/;

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
    gemm_tcgen05_kernel<<<grid, kNumThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(p);
    g_b200sd_launches.fetch_add(1, std::memory_order_relaxed);
    B200SD_LAUNCH_CHECK();
    return B200SD_OK;
}
