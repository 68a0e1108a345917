// b200sd -- flash attention on the sm_100a tensor cores (tcgen05 + TMEM), self-attention sized:
// head dim 40 / 80, S_q and S_kv multiples of 128 (the 64x64 and 32x32 levels of the UNet: 98 % of the
// attention FLOPs).  Other shapes keep the register-resident kernel in attention.cu.
//
// One CTA = 128 query rows of one (batch, head); it streams 128-key tiles:
//   warp 0    : TMA producer   Q once; K_j, V^T_j through a 2-stage smem ring (128B-swizzled boxes).  The
//               q/k/v buffers are addressed as 3-D tensors (d, head, row), so a 64-wide box over a 40-wide
//               head is zero-filled past the head by TMA -- no padding kernels, no masking.
//   warp 1    : TMEM allocator + single-thread MMA issuer:
//                   S_j   = Q K_j^T          (M128 x N128, K = d)      -> TMEM columns [0,128)
//                   O_j   = P_j V_j          (M128 x N = d, K = 128)   -> TMEM columns [128, 128+d)
//   warps 2-5 : softmax, one thread per query row (its TMEM lane): tcgen05.ld S, online max / exp2 / sum in
//               fp32, P_j written as bf16 into a swizzled smem tile (the A operand of the second MMA), then
//               O_j is read back and folded into register accumulators with the running-max correction.
// V is consumed K-major (keys contiguous), produced by a small transpose kernel into caller scratch.
// With 100 KB smem and 256 TMEM columns two CTAs share an SM and fill each other's MMA/softmax bubbles.
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

extern std::atomic<long long> g_b200sd_launches;

namespace {

constexpr int kQ = 128;     // query rows per CTA
constexpr int kKV = 128;    // keys per tile
constexpr int kBlk = kQ * 128;  // bytes of one [128 rows x 64 bf16] swizzled block

struct AttnParams {
    CUtensorMap tmQ, tmK, tmVt;
    bf16* out;
    float* lse;   // optional [batch][heads][Sq]: log2-domain log-sum-exp of the scaled scores (for the backward)
    int Sq, Skv, heads, ldo;
    float scale_log2;
};

// V [B*Skv, ldv] (head h at columns h*D..) -> Vt [B*H][D][Skv]
__global__ void transpose_v_kernel(const bf16* __restrict__ v, bf16* __restrict__ vt, int Skv, int heads, int D, int ldv) {
    __shared__ bf16 tile[64][72];  // [key][d] (+pad)
    ptx::pdl_trigger();
    ptx::pdl_wait();
    const int bh = blockIdx.z, b = bh / heads, h = bh % heads;
    const int s0 = blockIdx.x * 64, d0 = blockIdx.y * 64;
    const int dcount = min(64, D - d0);
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {   // 64 keys x 8 16-byte chunks
        const int r = i >> 3, c = (i & 7) * 8;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (c < dcount && s0 + r < Skv) u = __ldg(reinterpret_cast<const uint4*>(v + (size_t)(b * Skv + s0 + r) * ldv + h * D + d0 + c));
        *reinterpret_cast<uint4*>(&tile[r][c]) = u;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < dcount * 8; i += blockDim.x) {  // dcount rows x 8 chunks of 8 keys
        const int dd = i >> 3, sc = (i & 7) * 8;
        bf16 tmp[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) tmp[j] = tile[sc + j][dd];
        if (s0 + sc < Skv)
            *reinterpret_cast<uint4*>(vt + ((size_t)bh * D + d0 + dd) * Skv + s0 + sc) = *reinterpret_cast<uint4*>(tmp);
    }
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Blackwell packed fp32x2 math and 3-input max: the softmax warps are instruction-issue bound, these halve the
// non-MUFU instructions per score element (FFMA2 / FADD2 / FMNMX3 in SASS)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// D = head dim (40 or 80); DKB = number of 64-wide blocks covering it (1 or 2)
// HALVES = softmax threads per query row (1: warps 2-5 own whole rows; 2: warps 2-9, each thread owns 64 of the 128
//          score columns and half of the O columns -- twice the warps to hide TMEM-load / MUFU latency behind)
// LAZY   = single-pass softmax against a lagging reference max: tile j is exponentiated against the reference of
//          tile j-1 while its own max is collected in the same pass; only when some row of the warp outgrew the
//          reference by more than 2^kLazyThr is the tile redone against the true max (S is still in TMEM).
constexpr float kLazyThr = 6.0f;

template <int D, int DKB, int HALVES, bool LAZY>
__global__ void __launch_bounds__(64 + 128 * HALVES, (D <= 40) ? 2 : 1) attention_tc_kernel(const __grid_constant__ AttnParams p) {
    constexpr int DN = (D + 15) / 16 * 16;    // MMA N of the PV product (48 / 80)
    constexpr int KSTEPS = (D + 15) / 16;     // UMMA K steps of the QK^T product
    constexpr int VBLK = D * 128;             // bytes of one V^T block [D rows x 64 keys]
    constexpr int VBLK_PAD = ((DN * 128 + 1023) / 1024) * 1024;  // padded so the MMA may read DN rows
    constexpr uint32_t kTmemCols = 256;
    constexpr int CW = 128 / HALVES;          // score columns per softmax thread
    constexpr int DNH = DN / HALVES;          // O columns per softmax thread (48 / 24 / 80 / 40)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                               // DKB blocks
    uint8_t* sK = sQ + DKB * kBlk;                    // 2 stages x DKB blocks
    uint8_t* sV = sK + 2 * DKB * kBlk;                // 2 stages x 2 key-blocks x VBLK_PAD
    uint8_t* sP = sV + 2 * 2 * VBLK_PAD;              // 2 key-blocks
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kBlk);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;   // [2]
    uint64_t* k_empty = bars + 3;  // [2]
    uint64_t* v_full = bars + 5;   // [2]
    uint64_t* v_empty = bars + 7;  // [2]
    uint64_t* s_full = bars + 9;
    uint64_t* p_full = bars + 10;
    uint64_t* o_full = bars + 11;
    uint64_t* o_empty = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
    float* xch = reinterpret_cast<float*>(bars + 32);   // [2 parities][2 halves][128 rows] row exchange between half-threads

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kQ, h = blockIdx.y, b = blockIdx.z;
    const int num_tiles = p.Skv / kKV;

    ptx::pdl_trigger();
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&p.tmQ);
        ptx::prefetch_tmap(&p.tmK);
        ptx::prefetch_tmap(&p.tmVt);
        ptx::mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&k_full[i], 1);
            ptx::mbar_init(&k_empty[i], 1);
            ptx::mbar_init(&v_full[i], 1);
            ptx::mbar_init(&v_empty[i], 1);
        }
        ptx::mbar_init(s_full, 1);
        ptx::mbar_init(p_full, 4 * HALVES);    // one (warp-aggregated) arrival per softmax warp
        ptx::mbar_init(o_full, 1);
        ptx::mbar_init(o_empty, 4 * HALVES);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;
    ptx::pdl_wait();

    if (warp == 0) {
        // ================= TMA producer =================
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(q_full, DKB * kBlk);
            for (int kb = 0; kb < DKB; ++kb) tma_load_3d(sQ + kb * kBlk, &p.tmQ, q_full, kb * 64, h, b * p.Sq + q0);
            for (int j = 0; j < num_tiles; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                ptx::mbar_wait(&k_empty[st], ph ^ 1);
                ptx::mbar_expect_tx(&k_full[st], DKB * kBlk);
                for (int kb = 0; kb < DKB; ++kb)
                    tma_load_3d(sK + (st * DKB + kb) * kBlk, &p.tmK, &k_full[st], kb * 64, h, b * p.Skv + j * kKV);
                ptx::mbar_wait(&v_empty[st], ph ^ 1);
                ptx::mbar_expect_tx(&v_full[st], 2 * VBLK);
                for (int kk = 0; kk < 2; ++kk)
                    tma_load_3d(sV + (st * 2 + kk) * VBLK_PAD, &p.tmVt, &v_full[st], j * kKV + kk * 64, 0, b * p.heads + h);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (ptx::elect_one()) {
            const uint32_t idesc_s = ptx::umma_idesc_bf16(128, 128);
            const uint32_t idesc_o = ptx::umma_idesc_bf16(128, DN);
            ptx::mbar_wait(q_full, 0);
            for (int j = 0; j < num_tiles; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                // ---- S_j = Q K_j^T ----
                ptx::mbar_wait(&k_full[st], ph);
                ptx::tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks) {
                    const int kb = ks / 4, kin = ks % 4;
                    const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(sQ + kb * kBlk)) + 2 * kin;
                    const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(sK + (st * DKB + kb) * kBlk)) + 2 * kin;
                    ptx::umma_bf16_ss(tmem_S, da, db, idesc_s, ks > 0 ? 1u : 0u);
                }
                ptx::umma_commit(&k_empty[st]);
                ptx::umma_commit(s_full);
                // ---- O_j = P_j V_j ----
                ptx::mbar_wait(p_full, j & 1);
                ptx::mbar_wait(&v_full[st], ph);
                ptx::mbar_wait(o_empty, (j & 1) ^ 1);
                ptx::tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {   // 128 keys = 8 x 16
                    const int kk = ks / 4, kin = ks % 4;
                    const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(sP + kk * kBlk)) + 2 * kin;
                    const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(sV + (st * 2 + kk) * VBLK_PAD)) + 2 * kin;
                    ptx::umma_bf16_ss(tmem_O, da, db, idesc_o, ks > 0 ? 1u : 0u);
                }
                ptx::umma_commit(&v_empty[st]);
                ptx::umma_commit(o_full);
            }
        }
    } else {
        // ================= softmax warps: HALVES threads per query row =================
        // Software-pipelined: the fold of O_{j-1} into the register accumulator happens AFTER P_j has been
        // handed to the tensor core, so the P_{j-1} V_{j-1} product overlaps the softmax of tile j.
        const int qd = warp & 3;                              // TMEM lane quadrant this warp may address
        const int hf = (HALVES == 2) ? ((warp - 2) >> 2) : 0; // which half of the columns
        const int row = qd * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        const uint32_t tS = tmem_S + lane_addr + hf * CW;
        const uint32_t tO = tmem_O + lane_addr + hf * DNH;
        float m = -INFINITY, l = 0.f, corr_prev = 1.f;
        float o[DNH];
#pragma unroll
        for (int i = 0; i < DNH; ++i) o[i] = 0.f;

        // value of the thread owning the other half of this row (double-buffered by exchange parity)
        auto partner = [&](float v, int n) -> float {
            if constexpr (HALVES == 1) {
                return v;
            } else {
                float* x = xch + (n & 1) * 256;
                x[hf * 128 + row] = v;
                ptx::named_bar_sync(1 + qd, 64);
                return x[(hf ^ 1) * 128 + row];
            }
        };
        auto fold_o = [&](int jj, float corr) {
            ptx::mbar_wait(o_full, jj & 1);
            ptx::tc_fence_after();
            uint32_t r[DNH];
#pragma unroll
            for (int c = 0; c + 16 <= DNH; c += 16) ptx::tmem_ld_32x32b_x16(tO + c, *reinterpret_cast<uint32_t(*)[16]>(&r[c]));
            if constexpr (DNH % 16 == 8) ptx::tmem_ld_32x32b_x8(tO + DNH - 8, *reinterpret_cast<uint32_t(*)[8]>(&r[DNH - 8]));
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < DNH; ++i) o[i] = o[i] * corr + __uint_as_float(r[i]);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(o_empty);
        };
        auto max_pass = [&]() -> float {
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < CW; c += 64) {
                uint32_t r0[32], r1[32];
                tmem_ld_32x32b_x32(tS + c, r0);
                tmem_ld_32x32b_x32(tS + c + 32, r1);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmax3(mx, __uint_as_float(r0[i]), __uint_as_float(r1[i]));
            }
            return mx;
        };
        // P = exp2(S * scale - off) -> bf16 -> swizzled smem tile; returns the fp32 row sum (of this thread's columns)
        auto exp_pass = [&](float off, bool track, float& mx) -> float {
            float2 rs2 = make_float2(0.f, 0.f);
            const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), noff2 = make_float2(-off, -off);
#pragma unroll
            for (int c = 0; c < CW; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tS + c, r);
                ptx::tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float2 t = ffma2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sc2, noff2);
                    const float2 e = make_float2(fast_exp2(t.x), fast_exp2(t.y));
                    rs2 = fadd2(rs2, e);
                    pk[i] = pack_bf16x2(e.x, e.y);
                    if (track) mx = fmax3(mx, __uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
                }
                const int col = hf * CW + c;
                uint8_t* blk = sP + (col / 64) * kBlk + row * 128;
#pragma unroll
                for (int g = 0; g < 4; ++g) {   // four 16-byte chunks (8 keys each)
                    const int chunk = ((col % 64) / 8 + g) ^ (row & 7);
                    *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                }
            }
            return rs2.x + rs2.y;
        };

        for (int j = 0; j < num_tiles; ++j) {
            ptx::mbar_wait(s_full, j & 1);
            ptx::tc_fence_after();
            float mx = -INFINITY, rs = 0.f, corr = 1.f;
            bool redo = true;
            if (LAZY && j > 0) {
                rs = exp_pass(m * p.scale_log2, true, mx);
                mx = fmaxf(mx, partner(mx, j));
                redo = __any_sync(0xffffffffu, (mx - m) * p.scale_log2 > kLazyThr);
            } else {
                mx = max_pass();
                mx = fmaxf(mx, partner(mx, j));
            }
            if (redo) {   // warp-uniform, and identical in the two warps sharing these rows
                const float mn = fmaxf(m, mx);
                corr = fast_exp2((m - mn) * p.scale_log2);
                m = mn;
                float unused = 0.f;
                rs = exp_pass(mn * p.scale_log2, false, unused);
            }
            l = l * corr + rs;
            ptx::tc_fence_before();
            ptx::fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(p_full);
            if (j > 0) fold_o(j - 1, corr_prev);   // P_{j-1} V_{j-1} ran while this tile's softmax was computed
            corr_prev = corr;
        }
        fold_o(num_tiles - 1, corr_prev);
        // ---- normalise and store ----
        l += (HALVES == 2) ? partner(l, num_tiles) : 0.f;
        const float inv = 1.0f / l;
        if (p.lse != nullptr && hf == 0) p.lse[((size_t)b * p.heads + h) * p.Sq + q0 + row] = m * p.scale_log2 + log2f(l);
        bf16* dst = p.out + ((size_t)b * p.Sq + q0 + row) * p.ldo + h * D + hf * DNH;
#pragma unroll
        for (int c = 0; c < DNH; c += 8) {
            if (hf * DNH + c < D) {
                uint4 u;
                u.x = pack_bf16x2(o[c] * inv, o[c + 1] * inv);
                u.y = pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv);
                u.z = pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv);
                u.w = pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv);
                *reinterpret_cast<uint4*>(dst + c) = u;
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int D, int DKB, int HALVES, bool LAZY>
int launch_tc(const bf16* q, const bf16* k, const bf16* vt, bf16* out, float* lse, int batch, int heads, int Sq, int Skv, int ldq,
              int ldk, int ldo, float scale, cudaStream_t s) {
    constexpr int DN = (D + 15) / 16 * 16;
    constexpr int VBLK_PAD = ((DN * 128 + 1023) / 1024) * 1024;
    AttnParams p;
    memset(&p, 0, sizeof(p));
    {
        const uint64_t dims[3] = {(uint64_t)D, (uint64_t)heads, (uint64_t)batch * Sq};
        const uint64_t str[3] = {0, (uint64_t)D * 2, (uint64_t)ldq * 2};
        const uint32_t box[3] = {64, 1, 128};
        int rc = b200sd_make_tmap(&p.tmQ, q, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {(uint64_t)D, (uint64_t)heads, (uint64_t)batch * Skv};
        const uint64_t str[3] = {0, (uint64_t)D * 2, (uint64_t)ldk * 2};
        const uint32_t box[3] = {64, 1, 128};
        int rc = b200sd_make_tmap(&p.tmK, k, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {(uint64_t)Skv, (uint64_t)D, (uint64_t)batch * heads};
        const uint64_t str[3] = {0, (uint64_t)Skv * 2, (uint64_t)D * Skv * 2};
        const uint32_t box[3] = {64, (uint32_t)D, 1};
        int rc = b200sd_make_tmap(&p.tmVt, vt, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    p.out = out;
    p.lse = lse;
    p.Sq = Sq;
    p.Skv = Skv;
    p.heads = heads;
    p.ldo = ldo;
    p.scale_log2 = scale * 1.4426950408889634f;
    const size_t smem = (size_t)DKB * kBlk + 2 * DKB * kBlk + 4 * VBLK_PAD + 2 * kBlk + 256 + 2048 + 1024;
    static bool configured = false;
    if (!configured) {
        B200SD_CUDA(cudaFuncSetAttribute(attention_tc_kernel<D, DKB, HALVES, LAZY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        B200SD_CUDA(cudaFuncSetAttribute(attention_tc_kernel<D, DKB, HALVES, LAZY>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = true;
    }
    B200SD_CUDA(b200sd_launch(attention_tc_kernel<D, DKB, HALVES, LAZY>, dim3(Sq / kQ, heads, batch), dim3(64 + 128 * HALVES), smem, s, p));
    g_b200sd_launches.fetch_add(1, std::memory_order_relaxed);
    B200SD_LAUNCH_CHECK();
    return B200SD_OK;
}

}  // namespace

// Returns B200SD_ERR_UNSUPPORTED when the shape is not covered (the caller falls back to attention.cu's kernel).
int b200sd_attention_tc(const void* q, const void* k, const void* v, void* out, float* lse, int batch, int heads, int Sq, int Skv,
                        int d, int ldq, int ldk, int ldv, int ldo, float scale, void* workspace, size_t ws_bytes,
                        cudaStream_t s) {
    if (!(d == 40 || d == 80) || Sq % kQ != 0 || Skv % kKV != 0) return B200SD_ERR_UNSUPPORTED;
    const size_t need = (size_t)batch * heads * d * Skv * sizeof(bf16);
    if (workspace == nullptr || ws_bytes < need) return B200SD_ERR_UNSUPPORTED;
    if ((ldq * 2) % 16 != 0 || (ldk * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(workspace) & 127) != 0) return B200SD_ERR_UNSUPPORTED;
    bf16* vt = static_cast<bf16*>(workspace);
    B200SD_CUDA(b200sd_launch(transpose_v_kernel, dim3(ceil_div(Skv, 64), ceil_div(d, 64), batch * heads), dim3(256), 0, s,
                              static_cast<const bf16*>(v), vt, Skv, heads, d, ldv));
    g_b200sd_launches.fetch_add(1, std::memory_order_relaxed);
    // softmax organisation: B200SD_ATTN_FWD = 0 (4 warps, two passes), 1 (4 warps, lazy), 2 (8 warps, two passes), 3 (8 warps, lazy)
    static const int variant = [] { const char* e = getenv("B200SD_ATTN_FWD"); return e ? atoi(e) : 3; }();
#define B200SD_ATTN_GO(DD, KB, HV, LZ)                                                                                          \
    return launch_tc<DD, KB, HV, LZ>(static_cast<const bf16*>(q), static_cast<const bf16*>(k), vt, static_cast<bf16*>(out), lse, \
                                     batch, heads, Sq, Skv, ldq, ldk, ldo, scale, s)
    if (d == 40) {
        switch (variant) {
            case 0: B200SD_ATTN_GO(40, 1, 1, false);
            case 1: B200SD_ATTN_GO(40, 1, 1, true);
            case 2: B200SD_ATTN_GO(40, 1, 2, false);
            default: B200SD_ATTN_GO(40, 1, 2, true);
        }
    }
    switch (variant) {
        case 0: B200SD_ATTN_GO(80, 2, 1, false);
        case 1: B200SD_ATTN_GO(80, 2, 1, true);
        case 2: B200SD_ATTN_GO(80, 2, 2, false);
        default: B200SD_ATTN_GO(80, 2, 2, true);
    }
#undef B200SD_ATTN_GO
}
