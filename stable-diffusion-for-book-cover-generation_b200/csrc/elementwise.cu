// b200sd -- scheduler / loss elementwise kernels.  HBM-bound: 128-bit
// vectorised, coalesced, grid sized in multiples of the SM count.
#include <atomic>

#include "common.cuh"

extern std::atomic<long long> g_b200sd_launches;
#define COUNT_LAUNCH() g_b200sd_launches.fetch_add(1, std::memory_order_relaxed)

namespace {

constexpr int kThreads = 256;

template <int DT>
struct Elem;
template <>
struct Elem<B200SD_F32> {
    using T = float;
};
template <>
struct Elem<B200SD_BF16> {
    using T = bf16;
};

// 8 consecutive elements -> float[8]
template <int DT>
__device__ __forceinline__ void load8(const void* p, int64_t i, float (&v)[8]) {
    if constexpr (DT == B200SD_F32) {
        const float4* q = reinterpret_cast<const float4*>(static_cast<const float*>(p) + i);
        float4 a = __ldg(q), b = __ldg(q + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const bf16*>(p) + i));
        float2 f;
        f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
        f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
        f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
        f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
    }
}
template <int DT>
__device__ __forceinline__ void store8(void* p, int64_t i, const float (&v)[8]) {
    if constexpr (DT == B200SD_F32) {
        float4* q = reinterpret_cast<float4*>(static_cast<float*>(p) + i);
        q[0] = make_float4(v[0], v[1], v[2], v[3]);
        q[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        uint4 u;
        u.x = pack_bf16x2(v[0], v[1]);
        u.y = pack_bf16x2(v[2], v[3]);
        u.z = pack_bf16x2(v[4], v[5]);
        u.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(static_cast<bf16*>(p) + i) = u;
    }
}
template <int DT>
__device__ __forceinline__ float load1(const void* p, int64_t i) {
    if constexpr (DT == B200SD_F32) return static_cast<const float*>(p)[i];
    else return __bfloat162float(static_cast<const bf16*>(p)[i]);
}
template <int DT>
__device__ __forceinline__ void store1(void* p, int64_t i, float v) {
    if constexpr (DT == B200SD_F32) static_cast<float*>(p)[i] = v;
    else static_cast<bf16*>(p)[i] = __float2bfloat16_rn(v);
}

inline int grid_for(int64_t nvec) {
    int64_t blocks = (nvec + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)b200sd_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ---------------------------------------------------------------------------------------------
// CFG + DDIM
// ---------------------------------------------------------------------------------------------
struct DdimCoef {
    float g, sa_t, sb_t, sa_p, sb_p;
};

__device__ __forceinline__ float ddim_one(float eu, float ec, float x, bool cfg, const DdimCoef& c, float& eps) {
    eps = cfg ? (eu + c.g * (ec - eu)) : eu;
    float x0 = (x - c.sb_t * eps) / c.sa_t;
    return c.sa_p * x0 + c.sb_p * eps;
}

template <int EDT, int XDT>
__global__ void __launch_bounds__(kThreads) cfg_ddim_kernel(const void* __restrict__ eps_u,
                                                            const void* __restrict__ eps_c,
                                                            const void* __restrict__ x, void* __restrict__ out,
                                                            void* __restrict__ eps_out, int64_t n, int64_t nvec, DdimCoef c,
                                                            const float4* __restrict__ coef_table, const int* __restrict__ cursor) {
    ptx::pdl_trigger();
    ptx::pdl_wait();
    if (coef_table != nullptr) {     // captured sampler: this step's coefficients come from a device table (b200sd_sampler_advance)
        const float4 v = coef_table[cursor[1]];
        c.sa_t = v.x, c.sb_t = v.y, c.sa_p = v.z, c.sb_p = v.w;
    }
    const bool cfg = eps_c != nullptr;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float eu[8], ec[8], xv[8], o[8], e[8];
        load8<EDT>(eps_u, i * 8, eu);
        if (cfg) load8<EDT>(eps_c, i * 8, ec);
        load8<XDT>(x, i * 8, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = ddim_one(eu[j], ec[j], xv[j], cfg, c, e[j]);
        store8<XDT>(out, i * 8, o);
        if (eps_out) store8<EDT>(eps_out, i * 8, e);
    }
    // scalar tail (n % 8)
    {
        for (int64_t i = nvec * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            float e;
            float eu = load1<EDT>(eps_u, i);
            float ec = cfg ? load1<EDT>(eps_c, i) : 0.f;
            float o = ddim_one(eu, ec, load1<XDT>(x, i), cfg, c, e);
            store1<XDT>(out, i, o);
            if (eps_out) store1<EDT>(eps_out, i, e);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// CFG + PLMS
// ---------------------------------------------------------------------------------------------
struct PlmsArgs {
    const void* hist[4];
    float w[5];
    int nhist;
    float g, cx, ce;
};

template <int EDT, int XDT>
__global__ void __launch_bounds__(kThreads) cfg_plms_kernel(const void* __restrict__ eps_u,
                                                            const void* __restrict__ eps_c,
                                                            const void* __restrict__ x, void* __restrict__ out,
                                                            void* __restrict__ eps_out, int64_t n, int64_t nvec, PlmsArgs a) {
    ptx::pdl_trigger();
    ptx::pdl_wait();
    const bool cfg = eps_c != nullptr;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float eu[8], ec[8], xv[8], e[8], acc[8], o[8];
        load8<EDT>(eps_u, i * 8, eu);
        if (cfg) load8<EDT>(eps_c, i * 8, ec);
        load8<XDT>(x, i * 8, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            e[j] = cfg ? (eu[j] + a.g * (ec[j] - eu[j])) : eu[j];
            acc[j] = a.w[0] * e[j];
        }
        for (int h = 0; h < a.nhist; ++h) {
            float hv[8];
            load8<EDT>(a.hist[h], i * 8, hv);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += a.w[1 + h] * hv[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = a.cx * xv[j] - a.ce * acc[j];
        store8<XDT>(out, i * 8, o);
        if (eps_out) store8<EDT>(eps_out, i * 8, e);
    }
    {
        for (int64_t i = nvec * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            float eu = load1<EDT>(eps_u, i);
            float e = cfg ? (eu + a.g * (load1<EDT>(eps_c, i) - eu)) : eu;
            float acc = a.w[0] * e;
            for (int h = 0; h < a.nhist; ++h) acc += a.w[1 + h] * load1<EDT>(a.hist[h], i);
            store1<XDT>(out, i, a.cx * load1<XDT>(x, i) - a.ce * acc);
            if (eps_out) store1<EDT>(eps_out, i, e);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// add_noise
// ---------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(kThreads) add_noise_kernel(const void* __restrict__ x0,
                                                             const void* __restrict__ noise,
                                                             const int64_t* __restrict__ timesteps,
                                                             const float* __restrict__ sa_table,
                                                             const float* __restrict__ sb_table,
                                                             void* __restrict__ out, int64_t per_sample, int64_t nvec, int T) {
    ptx::pdl_trigger();
    ptx::pdl_wait();
    const int b = blockIdx.y;
    int64_t t = timesteps[b];
    t = t < 0 ? 0 : (t >= T ? T - 1 : t);
    const float sa = __ldg(sa_table + t), sb = __ldg(sb_table + t);
    const int64_t base = (int64_t)b * per_sample;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float a[8], e[8], o[8];
        load8<DT>(x0, base + i * 8, a);
        load8<DT>(noise, base + i * 8, e);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = sa * a[j] + sb * e[j];
        store8<DT>(out, base + i * 8, o);
    }
    {
        for (int64_t i = nvec * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += stride)
            store1<DT>(out, base + i, sa * load1<DT>(x0, base + i) + sb * load1<DT>(noise, base + i));
    }
}

// ---------------------------------------------------------------------------------------------
// MSE
// ---------------------------------------------------------------------------------------------
constexpr int kMsePartials = 1024;

__device__ __forceinline__ float block_sum(float v, float* sm) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    float r = 0.f;
    if (warp == 0) {
        r = lane < (blockDim.x >> 5) ? sm[lane] : 0.f;
        r = warp_sum(r);
    }
    return r;  // valid in warp 0
}

template <int PDT, int TDT>
__global__ void __launch_bounds__(kThreads) mse_fwd_kernel(const void* __restrict__ pred,
                                                           const void* __restrict__ target,
                                                           float* __restrict__ loss_out, float* __restrict__ ws,
                                                           int64_t n, int64_t nvec) {
    __shared__ float sm[32];
    __shared__ bool is_last;
    ptx::pdl_trigger();
    ptx::pdl_wait();
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float p[8], t[8];
        load8<PDT>(pred, i * 8, p);
        load8<TDT>(target, i * 8, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float d = p[j] - t[j];
            acc += d * d;
        }
    }
    {
        for (int64_t i = nvec * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            float d = load1<PDT>(pred, i) - load1<TDT>(target, i);
            acc += d * d;
        }
    }
    float bs = block_sum(acc, sm);
    unsigned int* counter = reinterpret_cast<unsigned int*>(ws + kMsePartials);
    if (threadIdx.x == 0) {
        ws[blockIdx.x] = bs;
        __threadfence();
        unsigned int prev = atomicAdd(counter, 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // deterministic: fixed-order tree over the per-block partials, accumulated in double
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += (double)__ldcg(ws + i);
        __shared__ double dsm[kThreads];
        dsm[threadIdx.x] = s;
        __syncthreads();
        for (int o = kThreads / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) dsm[threadIdx.x] += dsm[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            loss_out[0] = (float)(dsm[0] / (double)n);
            *counter = 0;  // self-cleaning for the next launch
        }
    }
}

template <int PDT, int TDT>
__global__ void __launch_bounds__(kThreads) mse_bwd_kernel(const void* __restrict__ pred,
                                                           const void* __restrict__ target,
                                                           const float* __restrict__ grad_loss,
                                                           void* __restrict__ grad_pred, int64_t n, int64_t nvec) {
    ptx::pdl_trigger();
    ptx::pdl_wait();
    const float s = __ldg(grad_loss) * 2.0f / (float)n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float p[8], t[8], g[8];
        load8<PDT>(pred, i * 8, p);
        load8<TDT>(target, i * 8, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = s * (p[j] - t[j]);
        store8<PDT>(grad_pred, i * 8, g);
    }
    {
        for (int64_t i = nvec * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            store1<PDT>(grad_pred, i, s * (load1<PDT>(pred, i) - load1<TDT>(target, i)));
    }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

#define DISPATCH2(E, X, KERNEL, GRID, STREAM, ...)                                                                   \
    do {                                                                                                              \
        if ((E) == B200SD_F32 && (X) == B200SD_F32)                                                                   \
            B200SD_CUDA(b200sd_launch(KERNEL<B200SD_F32, B200SD_F32>, dim3(GRID), dim3(kThreads), 0, STREAM, __VA_ARGS__));  \
        else if ((E) == B200SD_BF16 && (X) == B200SD_F32)                                                             \
            B200SD_CUDA(b200sd_launch(KERNEL<B200SD_BF16, B200SD_F32>, dim3(GRID), dim3(kThreads), 0, STREAM, __VA_ARGS__)); \
        else if ((E) == B200SD_F32 && (X) == B200SD_BF16)                                                             \
            B200SD_CUDA(b200sd_launch(KERNEL<B200SD_F32, B200SD_BF16>, dim3(GRID), dim3(kThreads), 0, STREAM, __VA_ARGS__)); \
        else                                                                                                          \
            B200SD_CUDA(b200sd_launch(KERNEL<B200SD_BF16, B200SD_BF16>, dim3(GRID), dim3(kThreads), 0, STREAM, __VA_ARGS__)); \
    } while (0)

static bool dtype_ok(int d) { return d == B200SD_F32 || d == B200SD_BF16; }

extern "C" int b200sd_cfg_ddim_step(const void* eps_u, const void* eps_c, const void* x, void* out, void* eps_out,
                                    int64_t n, float guidance, float sa_t, float sb_t, float sa_p, float sb_p,
                                    int eps_dtype, int x_dtype, b200sd_stream_t stream) {
    B200SD_REQUIRE(eps_u && x && out, "cfg_ddim_step: null pointer");
    B200SD_REQUIRE(n >= 0, "cfg_ddim_step: negative n");
    B200SD_REQUIRE(dtype_ok(eps_dtype) && dtype_ok(x_dtype), "cfg_ddim_step: bad dtype");
    const bool vec = aligned16(eps_u) && aligned16(eps_c) && aligned16(x) && aligned16(out) && aligned16(eps_out);
    const int64_t nvec = vec ? n / 8 : 0;  // unaligned views (odd slices) take the scalar path
    B200SD_REQUIRE(sa_t != 0.f, "cfg_ddim_step: sa_t == 0");
    if (n == 0) return B200SD_OK;
    DdimCoef c{guidance, sa_t, sb_t, sa_p, sb_p};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DISPATCH2(eps_dtype, x_dtype, cfg_ddim_kernel, grid_for(vec ? n / 8 : n), s, eps_u, eps_c, x, out, eps_out, n, nvec, c,
              (const float4*)nullptr, (const int*)nullptr);
    COUNT_LAUNCH();
    B200SD_LAUNCH_CHECK();
    return B200SD_OK;
}

// Captured sampler (one CUDA graph per denoising step, replayed with no host-side arguments): the step index lives on the
// device.  cursor[0] = next step, cursor[1] = the step being executed.
namespace {
__global__ void sampler_advance_kernel(const float* __restrict__ timesteps, int n_steps, int* __restrict__ cursor,
                                       float* __restrict__ in_t, int n_t) {
    const int c = cursor[0];
    const float t = timesteps[c];
    for (int i = threadIdx.x; i < n_t; i += blockDim.x) in_t[i] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        cursor[1] = c;
        cursor[0] = (c + 1 == n_steps) ? 0 : c + 1;
    }
}
}  // namespace

extern "C" int b200sd_sampler_advance(const float* timesteps, int n_steps, int* cursor, float* in_t, int n_t,
                                      b200sd_stream_t stream) {
    B200SD_REQUIRE(timesteps && cursor && in_t, "sampler_advance: null pointer");
    B200SD_REQUIRE(n_steps > 0 && n_t > 0, "sampler_advance: empty table");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    B200SD_CUDA(b200sd_launch(sampler_advance_kernel, dim3(1), dim3(64), 0, s, timesteps, n_steps, cursor, in_t, n_t));
    COUNT_LAUNCH();
    B200SD_LAUNCH_CHECK();
    return B200SD_OK;
}

// Captured PLMS step: everything that changes from call to call -- the linear-multistep weights, which slots of the 4-deep eps
// ring hold the history, where this call's eps goes, whether x is the running sample or the one saved by the first call
// (PNDM's second call re-does the first timestep), the two scalars of _get_prev_sample -- is one row of a device table indexed
// by the cursor.  Row layout (12 floats): w0 w1 w2 w3 | cx ce | h1 h2 h3 (ring slots, -1 = unused) | out_slot (-1 = not
// kept) | x_from_saved | save_x.
namespace {
constexpr int kPlmsRow = 12;
__global__ void __launch_bounds__(kThreads) cfg_plms_table_kernel(const float* __restrict__ eps_u, const float* __restrict__ eps_c,
                                                                  float* __restrict__ x, float* __restrict__ saved,
                                                                  float* __restrict__ ring, int64_t n, float g,
                                                                  const float* __restrict__ table, const int* __restrict__ cursor) {
    ptx::pdl_trigger();
    ptx::pdl_wait();
    const float* row = table + (size_t)cursor[1] * kPlmsRow;
    const float w0 = row[0], w1 = row[1], w2 = row[2], w3 = row[3], cx = row[4], ce = row[5];
    const int h1 = (int)row[6], h2 = (int)row[7], h3 = (int)row[8], slot = (int)row[9];
    const bool from_saved = row[10] != 0.f, save_x = row[11] != 0.f;
    const bool cfg = eps_c != nullptr;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float eu = eps_u[i];
        const float e = cfg ? (eu + g * (eps_c[i] - eu)) : eu;
        float acc = w0 * e;
        if (h1 >= 0) acc = fmaf(w1, ring[(size_t)h1 * n + i], acc);
        if (h2 >= 0) acc = fmaf(w2, ring[(size_t)h2 * n + i], acc);
        if (h3 >= 0) acc = fmaf(w3, ring[(size_t)h3 * n + i], acc);
        const float cur = x[i];
        const float xv = from_saved ? saved[i] : cur;
        if (save_x) saved[i] = cur;
        x[i] = cx * xv - ce * acc;
        if (slot >= 0) ring[(size_t)slot * n + i] = e;
    }
}
}  // namespace

extern "C" int b200sd_cfg_plms_step_table(const float* eps_u, const float* eps_c, float* x, float* saved, float* ring, int64_t n,
                                          float guidance, const float* table, const int* cursor, b200sd_stream_t stream) {
    B200SD_REQUIRE(eps_u && x && saved && ring && table && cursor, "cfg_plms_step_table: null pointer");
    B200SD_REQUIRE(n >= 0, "cfg_plms_step_table: negative n");
    if (n == 0) return B200SD_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    B200SD_CUDA(b200sd_launch(cfg_plms_table_kernel, dim3(grid_for(n)), dim3(kThreads), 0, s, eps_u, eps_c, x, saved, ring, n, guidance,
                              table, cursor));
    COUNT_LAUNCH();
    B200SD_LAUNCH_CHECK();
    return B200SD_OK;
}

extern "C" int b200sd_cfg_ddim_step_table(const void* eps_u, const void* eps_c, const void* x, void* out, void* eps_out,
                                          int64_t n, float guidance, const float* coef_table, const int* cursor,
                                          int eps_dtype, int x_dtype, b200sd_stream_t stream) {
    B200SD_REQUIRE(eps_u && x && out && coef_table && cursor, "cfg_ddim_step_table: null pointer");
    B200SD_REQUIRE(n >= 0, "cfg_ddim_step_table: negative n");
    B200SD_REQUIRE(dtype_ok(eps_dtype) && dtype_ok(x_dtype), "cfg_ddim_step_table: bad dtype");
    B200SD_REQUIRE(aligned16(coef_table), "cfg_ddim_step_table: coef_table must be 16-byte aligned");
    const bool vec = aligned16(eps_u) && aligned16(eps_c) && aligned16(x) && aligned16(out) && aligned16(eps_out);
    const int64_t nvec = vec ? n / 8 : 0;
    if (n == 0) return B200SD_OK;
    DdimCoef c{guidance, 1.f, 0.f, 1.f, 0.f};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DISPATCH2(eps_dtype, x_dtype, cfg_ddim_kernel, grid_for(vec ? n / 8 : n), s, eps_u, eps_c, x, out, eps_out, n, nvec, c,
              reinterpret_cast<const float4*>(coef_table), cursor);
    COUNT_LAUNCH();
    B200SD_LAUNCH_CHECK();
    return B200SD_OK;
}

extern "C" int b200sd_cfg_plms_step(const void* eps_u, const void* eps_c, const void* x, void* out, void* eps_out,
                                    const void* hist0, const void* hist1, const void* hist2, const void* hist3,
                                    int nhist, const float* w_host5, int64_t n, float guidance, float cx, float ce,
                                    int eps_dtype, int x_dtype, b200sd_stream_t stream) {
    B200SD_REQUIRE(eps_u && x && out && w_host5, "cfg_plms_step: null pointer");
    B200SD_REQUIRE(nhist >= 0 && nhist <= 4, "cfg_plms_step: nhist out of range");
    B200SD_REQUIRE(dtype_ok(eps_dtype) && dtype_ok(x_dtype), "cfg_plms_step: bad dtype");
    PlmsArgs a;
    a.hist[0] = hist0; a.hist[1] = hist1; a.hist[2] = hist2; a.hist[3] = hist3;
    bool vec = aligned16(eps_u) && aligned16(eps_c) && aligned16(x) && aligned16(out) && aligned16(eps_out);
    for (int i = 0; i < nhist; ++i) {
        B200SD_REQUIRE(a.hist[i] != nullptr, "cfg_plms_step: null history pointer");
        vec = vec && aligned16(a.hist[i]);
    }
    const int64_t nvec = vec ? n / 8 : 0;
    for (int i = 0; i < 5; ++i) a.w[i] = w_host5[i];
    a.nhist = nhist; a.g = guidance; a.cx = cx; a.ce = ce;
    if (n == 0) return B200SD_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DISPATCH2(eps_dtype, x_dtype, cfg_plms_kernel, grid_for(vec ? n / 8 : n), s, eps_u, eps_c, x, out, eps_out, n, nvec, a);
    COUNT_LAUNCH();
    B200SD_LAUNCH_CHECK();
    return B200SD_OK;
}

extern "C" int b200sd_add_noise(const void* x0, const void* noise, const int64_t* timesteps, const float* sa_table,
                                const float* sb_table, void* out, int batch, int64_t per_sample,
                                int num_train_timesteps, int dtype, b200sd_stream_t stream) {
    B200SD_REQUIRE(x0 && noise && timesteps && sa_table && sb_table && out, "add_noise: null pointer");
    B200SD_REQUIRE(batch >= 0 && per_sample >= 0 && batch <= 65535, "add_noise: bad sizes");
    B200SD_REQUIRE(dtype_ok(dtype), "add_noise: bad dtype");
    const bool vec = aligned16(x0) && aligned16(noise) && aligned16(out) && (per_sample % 8 == 0 || batch <= 1);
    const int64_t nvec = vec ? per_sample / 8 : 0;
    if (batch == 0 || per_sample == 0) return B200SD_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int gx = grid_for(vec ? per_sample / 8 : per_sample);
    int cap = b200sd_num_sms() * 8 / batch;
    if (cap < 1) cap = 1;
    if (gx > cap) gx = cap;
    dim3 grid(gx, batch);
    if (dtype == B200SD_F32)
        B200SD_CUDA(b200sd_launch(add_noise_kernel<B200SD_F32>, dim3(grid), dim3(kThreads), 0, s, x0, noise, timesteps, sa_table, sb_table, out, per_sample, nvec, num_train_timesteps));
    else
        B200SD_CUDA(b200sd_launch(add_noise_kernel<B200SD_BF16>, dim3(grid), dim3(kThreads), 0, s, x0, noise, timesteps, sa_table, sb_table, out, per_sample, nvec, num_train_timesteps));
    COUNT_LAUNCH();
    B200SD_LAUNCH_CHECK();
    return B200SD_OK;
}

extern "C" int b200sd_mse_workspace_floats(void) { return kMsePartials + 4; }

extern "C" int b200sd_mse_loss_fwd(const void* pred, const void* target, float* loss_out, float* workspace,
                                   int64_t n, int pred_dtype, int target_dtype, b200sd_stream_t stream) {
    B200SD_REQUIRE(pred && target && loss_out && workspace, "mse_loss_fwd: null pointer");
    B200SD_REQUIRE(n > 0, "mse_loss_fwd: n must be positive");
    B200SD_REQUIRE(dtype_ok(pred_dtype) && dtype_ok(target_dtype), "mse_loss_fwd: bad dtype");
    const bool vec = aligned16(pred) && aligned16(target);
    const int64_t nvec = vec ? n / 8 : 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int grid = grid_for(vec ? n / 8 : n);
    if (grid > kMsePartials) grid = kMsePartials;
    DISPATCH2(pred_dtype, target_dtype, mse_fwd_kernel, grid, s, pred, target, loss_out, workspace, n, nvec);
    COUNT_LAUNCH();
    B200SD_LAUNCH_CHECK();
    return B200SD_OK;
}

extern "C" int b200sd_mse_loss_bwd(const void* pred, const void* target, const float* grad_loss, void* grad_pred,
                                   int64_t n, int pred_dtype, int target_dtype, b200sd_stream_t stream) {
    B200SD_REQUIRE(pred && target && grad_loss && grad_pred, "mse_loss_bwd: null pointer");
    B200SD_REQUIRE(n > 0, "mse_loss_bwd: n must be positive");
    B200SD_REQUIRE(dtype_ok(pred_dtype) && dtype_ok(target_dtype), "mse_loss_bwd: bad dtype");
    const bool vec = aligned16(pred) && aligned16(target) && aligned16(grad_pred);
    const int64_t nvec = vec ? n / 8 : 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DISPATCH2(pred_dtype, target_dtype, mse_bwd_kernel, grid_for(vec ? n / 8 : n), s, pred, target, grad_loss, grad_pred, n, nvec);
    COUNT_LAUNCH();
    B200SD_LAUNCH_CHECK();
    return B200SD_OK;
}
