// GroupNorm (+SiLU, +dropout) forward/backward, nearest-x2 up-sampling and its adjoint, channel-slice
// add and per-channel pixel reductions -- the HBM-bound glue between the implicit-GEMM convolutions.
// Replaces attention.py:38-39 / unet_backbone.py:20-21,29-31,63,98-99 (native_group_norm, silu,
// dropout, upsample_nearest2d; SURVEY.md §2.2 rows 4,5,6,9).
//
// Layout: NHWC with pitch (channel slices of concat buffers) and an optional zero halo that these
// kernels never write.  Thread mapping for every kernel here: a thread owns one 16-byte channel
// vector position `cv` and walks pixels, so global accesses are 16 B per lane and contiguous along C.
//
// HBM roofline (bf16 activations): gn_stats 2 B/elem, gn_apply 4 B/elem, gn_bwd 2x(4)+2 = 10 B/elem.
#include "common.cuh"

#define NT 256

struct PixMap {
    int cvs;    // channel vectors per pixel
    int ppi;    // pixels processed per block iteration
    int cv;     // this thread's channel vector
    int prow;   // this thread's pixel lane
    bool active;
};
template <int VEC> __device__ __forceinline__ PixMap make_map(int C) {
    PixMap m;
    m.cvs = C / VEC;
    m.ppi = NT / m.cvs; if (m.ppi < 1) m.ppi = 1;
    m.cv = threadIdx.x % m.cvs;
    m.prow = threadIdx.x / m.cvs;
    m.active = m.prow < m.ppi;
    return m;
}

template <typename T, int VEC> __device__ __forceinline__ void ldv(const T* p, float* v) {
    if (VEC == 1) v[0] = ldf<T>(p);
    else { Vec16<T> t; t.load(p);
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = t.v[i]; }
}
template <typename T, int VEC> __device__ __forceinline__ void stv(T* p, const float* v) {
    if (VEC == 1) stf<T>(p, v[0]);
    else { Vec16<T> t;
#pragma unroll
        for (int i = 0; i < VEC; ++i) t.v[i] = v[i];
        t.store(p); }
}

static inline bool vec_ok(const ddpm_tensor* t, int vec, int esz) {
    return (t->C % vec) == 0 && (t->pitch % vec) == 0 && ((((uintptr_t)t->ptr) % (vec * esz)) == 0);
}
static inline int blocks_per_image(int N, int HW, int ppi) {
    int want = (148 * 16 + N - 1) / N;
    int maxb = (HW + ppi - 1) / ppi;
    if (want < 1) want = 1;
    return want < maxb ? want : maxb;
}

// ------------------------------------------------------------------------------------ gn_stats
template <typename T, int VEC>
__global__ void __launch_bounds__(NT, 4) gn_stats_kernel(TV x, int G, double* stats) {
    extern __shared__ double sg[];          // [G][2]
    const int n = blockIdx.y, cpg = x.C / G;
    for (int i = threadIdx.x; i < 2 * G; i += NT) sg[i] = 0.0;
    __syncthreads();
    PixMap m = make_map<VEC>(x.C);
    if (m.cvs > NT) {                        // very wide tensors: loop channel vectors too
        // (not reachable for C <= 2048 with VEC >= 4; scalar path handles C <= 256)
    }
    float s[VEC], q[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { s[i] = 0.f; q[i] = 0.f; }
    const int HW = x.H * x.W;
    const int per = (HW + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
    if (m.active) {
        for (int p = p0 + m.prow; p < p1; p += m.ppi) {
            int y = p / x.W, xx = p - y * x.W;
            float v[VEC];
            ldv<T, VEC>(x.at<T>(n, y, xx, m.cv * VEC), v);
#pragma unroll
            for (int i = 0; i < VEC; ++i) { s[i] += v[i]; q[i] += v[i] * v[i]; }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            int g = (m.cv * VEC + i) / cpg;
            atomicAdd(&sg[2 * g], (double)s[i]);
            atomicAdd(&sg[2 * g + 1], (double)q[i]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += NT) atomicAdd(&stats[(size_t)n * 2 * G + i], sg[i]);
}

extern "C" int ddpm_gn_stats(const ddpm_tensor* x, int dtype, int groups, double* stats, void* stream) {
    if (!tensor_ok(x) || !stats || groups <= 0 || x->C % groups) return DDPM_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * groups * x->N, st));
    TV v(*x);
    int HW = x->H * x->W;
    size_t sm = sizeof(double) * 2 * groups;
#define GO(T, VEC) { int cvs = x->C / VEC; if (cvs > NT) return DDPM_E_ARG; int ppi = NT / cvs; \
        dim3 grid(blocks_per_image(x->N, HW, ppi * 8), x->N); \
        gn_stats_kernel<T, VEC><<<grid, NT, sm, st>>>(v, groups, stats); }
    if (dtype == DDPM_BF16) { if (vec_ok(x, 8, 2)) GO(bf16, 8) else GO(bf16, 1) }
    else if (dtype == DDPM_F32) { if (vec_ok(x, 4, 4)) GO(float, 4) else GO(float, 1) }
    else return DDPM_E_ARG;
#undef GO
    LAUNCH_OK();
    return 0;
}

// per-channel affine tables for image n:  y = x*scale + shift ; xhat = x*rstd - mean*rstd
__device__ __forceinline__ void group_moments(const double* stats, int n, int G, int g, double cnt, float eps,
                                              float* mean, float* rstd) {
    double s = stats[((size_t)n * G + g) * 2], q = stats[((size_t)n * G + g) * 2 + 1];
    double mu = s / cnt;
    double var = q / cnt - mu * mu;
    if (var < 0.0) var = 0.0;
    *mean = (float)mu;
    *rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// ------------------------------------------------------------------------------------ gn_apply
template <typename T, int VEC>
__global__ void __launch_bounds__(NT, 4) gn_apply_kernel(TV x, TV o, int G, const double* stats, const float* gamma,
                                                      const float* beta, float eps, int act, float p_drop,
                                                      const uint64_t* rng, uint32_t layer) {
    extern __shared__ float tb[];            // scale[C], shift[C]
    const int n = blockIdx.y, C = x.C, cpg = C / G, HW = x.H * x.W;
    float* scale = tb; float* shift = tb + C;
    for (int c = threadIdx.x; c < C; c += NT) {
        float mu, rs;
        group_moments(stats, n, G, c / cpg, (double)cpg * HW, eps, &mu, &rs);
        float sc = rs * gamma[c];
        scale[c] = sc; shift[c] = beta[c] - mu * sc;
    }
    __syncthreads();
    PixMap m = make_map<VEC>(C);
    if (!m.active) return;
    float sc[VEC], sh[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { sc[i] = scale[m.cv * VEC + i]; sh[i] = shift[m.cv * VEC + i]; }
    const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    const int per = (HW + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
    constexpr int U = 2;
    for (int pb = p0 + m.prow; pb < p1; pb += U * m.ppi) {
        float v[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int p = pb + u * m.ppi;
            if (p < p1) { int y = p / x.W, xx = p - y * x.W; ldv<T, VEC>(x.at<T>(n, y, xx, m.cv * VEC), v[u]); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int p = pb + u * m.ppi;
            if (p >= p1) break;
            uint32_t keep = 0xffffffffu;
            if (p_drop > 0.f) keep = dropout_mask<VEC>(rng, layer, ((uint64_t)n * HW + p) * C + m.cv * VEC, p_drop);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float z = fmaf(v[u][i], sc[i], sh[i]);
                if (act) z = silu_f(z);
                v[u][i] = ((keep >> i) & 1u) ? z * keep_scale : 0.f;
            }
            int y = p / x.W, xx = p - y * x.W;
            stv<T, VEC>(o.at<T>(n, y, xx, m.cv * VEC), v[u]);
        }
    }
}

extern "C" int ddpm_gn_apply(const ddpm_tensor* x, int dtype, int groups, const double* stats, const float* gamma,
                             const float* beta, float eps, int act, float p_drop, const uint64_t* rng,
                             uint32_t layer_id, const ddpm_tensor* out, void* stream) {
    if (!tensor_ok(x) || !tensor_ok(out) || !stats || !gamma || !beta || groups <= 0 || x->C % groups) return DDPM_E_ARG;
    if (out->N != x->N || out->H != x->H || out->W != x->W || out->C != x->C) return DDPM_E_ARG;
    if (p_drop > 0.f && !rng) return DDPM_E_ARG;
    if (p_drop < 0.f || p_drop >= 1.f) return DDPM_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    TV v(*x), o(*out);
    int HW = x->H * x->W;
    size_t sm = sizeof(float) * 2 * x->C;
#define GO(T, VEC) { int cvs = x->C / VEC; if (cvs > NT) return DDPM_E_ARG; int ppi = NT / cvs; \
        dim3 grid(blocks_per_image(x->N, HW, ppi * 4), x->N); \
        gn_apply_kernel<T, VEC><<<grid, NT, sm, st>>>(v, o, groups, stats, gamma, beta, eps, act, p_drop, rng, layer_id); }
    if (dtype == DDPM_BF16) { if (vec_ok(x, 8, 2) && vec_ok(out, 8, 2)) GO(bf16, 8) else GO(bf16, 1) }
    else if (dtype == DDPM_F32) { if (vec_ok(x, 4, 4) && vec_ok(out, 4, 4)) GO(float, 4) else GO(float, 1) }
    else return DDPM_E_ARG;
#undef GO
    LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------ gn_bwd
// z = x*scale+shift, y = drop(act(z)).  dz = dy * mask/(1-p) * act'(z).
// pass 1: ws[n][c] = (sum_p dz, sum_p dz*xhat).
template <typename T, int VEC>
__global__ void __launch_bounds__(NT, 4) gn_bwd_reduce_kernel(TV x, TV dy, int G, const double* stats, const float* gamma,
                                                              const float* beta, float eps, int act, float p_drop,
                                                              const uint64_t* rng, uint32_t layer, float* ws) {
    extern __shared__ float tb[];            // rs[C], mr[C], ga[C], be[C], s1[C], s2[C]
    const int n = blockIdx.y, C = x.C, cpg = C / G, HW = x.H * x.W;
    float* rsv = tb; float* mrv = tb + C; float* gav = tb + 2 * C; float* bev = tb + 3 * C;
    float* a1 = tb + 4 * C; float* a2 = tb + 5 * C;
    for (int c = threadIdx.x; c < C; c += NT) {
        float mu, rs;
        group_moments(stats, n, G, c / cpg, (double)cpg * HW, eps, &mu, &rs);
        rsv[c] = rs; mrv[c] = mu * rs; gav[c] = gamma[c]; bev[c] = beta[c]; a1[c] = 0.f; a2[c] = 0.f;
    }
    __syncthreads();
    PixMap m = make_map<VEC>(C);
    if (m.active) {
        float s1[VEC], s2[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
        const int c0 = m.cv * VEC;
        const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
        const int per = (HW + gridDim.x - 1) / gridDim.x;
        const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
        for (int p = p0 + m.prow; p < p1; p += m.ppi) {
            int y = p / x.W, xx = p - y * x.W;
            float v[VEC], d[VEC];
            ldv<T, VEC>(x.at<T>(n, y, xx, c0), v);
            ldv<T, VEC>(dy.at<T>(n, y, xx, c0), d);
            uint32_t keep = 0xffffffffu;
            if (p_drop > 0.f) keep = dropout_mask<VEC>(rng, layer, ((uint64_t)n * HW + p) * C + c0, p_drop);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float xh = fmaf(v[i], rsv[c0 + i], -mrv[c0 + i]);
                float dz = ((keep >> i) & 1u) ? d[i] * keep_scale : 0.f;
                if (act) dz *= dsilu_f(fmaf(xh, gav[c0 + i], bev[c0 + i]));
                s1[i] += dz; s2[i] = fmaf(dz, xh, s2[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            atomicAdd(&a1[c0 + i], s1[i]);
            atomicAdd(&a2[c0 + i], s2[i]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += NT) {
        atomicAdd(&ws[((size_t)n * C + c) * 2], a1[c]);
        atomicAdd(&ws[((size_t)n * C + c) * 2 + 1], a2[c]);
    }
}

// pass 2: dx = rstd * (dz*gamma - A_g - xhat*B_g),  A_g = mean_g(gamma*S1), B_g = mean_g(gamma*S2)
template <typename T, int VEC>
__global__ void __launch_bounds__(NT, 4) gn_bwd_apply_kernel(TV x, TV dy, TV dx, int G, const double* stats,
                                                             const float* gamma, const float* beta, float eps, int act,
                                                             float p_drop, const uint64_t* rng, uint32_t layer,
                                                             const float* ws, int accumulate) {
    extern __shared__ float tb[];            // rs[C], mr[C], ga[C], be[C], ra[C] (= rs*A_g), rb[C] (= rs*B_g), ag[G], bg[G]
    const int n = blockIdx.y, C = x.C, cpg = C / G, HW = x.H * x.W;
    float* rsv = tb; float* mrv = tb + C; float* gav = tb + 2 * C; float* bev = tb + 3 * C;
    float* rav = tb + 4 * C; float* rbv = tb + 5 * C; float* ag = tb + 6 * C; float* bg = ag + G;
    for (int g = threadIdx.x; g < G; g += NT) {
        float a = 0.f, b = 0.f;
        for (int j = 0; j < cpg; ++j) {
            int c = g * cpg + j;
            a += gamma[c] * ws[((size_t)n * C + c) * 2];
            b += gamma[c] * ws[((size_t)n * C + c) * 2 + 1];
        }
        float inv = 1.0f / ((float)cpg * (float)HW);
        ag[g] = a * inv; bg[g] = b * inv;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += NT) {
        float mu, rs;
        group_moments(stats, n, G, c / cpg, (double)cpg * HW, eps, &mu, &rs);
        rsv[c] = rs; mrv[c] = mu * rs; gav[c] = gamma[c]; bev[c] = beta[c];
        rav[c] = rs * ag[c / cpg]; rbv[c] = rs * bg[c / cpg];
    }
    __syncthreads();
    PixMap m = make_map<VEC>(C);
    if (!m.active) return;
    const int c0 = m.cv * VEC;
    const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    const int per = (HW + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
    for (int p = p0 + m.prow; p < p1; p += m.ppi) {
        int y = p / x.W, xx = p - y * x.W;
        float v[VEC], d[VEC], r[VEC];
        ldv<T, VEC>(x.at<T>(n, y, xx, c0), v);
        ldv<T, VEC>(dy.at<T>(n, y, xx, c0), d);
        if (accumulate) ldv<T, VEC>(dx.at<T>(n, y, xx, c0), r);
        uint32_t keep = 0xffffffffu;
        if (p_drop > 0.f) keep = dropout_mask<VEC>(rng, layer, ((uint64_t)n * HW + p) * C + c0, p_drop);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float rs = rsv[c0 + i], ga = gav[c0 + i];
            float xh = fmaf(v[i], rs, -mrv[c0 + i]);
            float dz = ((keep >> i) & 1u) ? d[i] * keep_scale : 0.f;
            if (act) dz *= dsilu_f(fmaf(xh, ga, bev[c0 + i]));
            float g = fmaf(dz * ga, rs, -rav[c0 + i]) - xh * rbv[c0 + i];     // rs*(dz*ga - A - xh*B)
            r[i] = accumulate ? r[i] + g : g;
        }
        stv<T, VEC>(dx.at<T>(n, y, xx, c0), r);
    }
}

__global__ void gn_param_grad_kernel(const float* ws, int N, int C, float* dgamma, float* dbeta) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f, b = 0.f;
    for (int n = 0; n < N; ++n) { a += ws[((size_t)n * C + c) * 2]; b += ws[((size_t)n * C + c) * 2 + 1]; }
    if (dbeta) dbeta[c] += a;
    if (dgamma) dgamma[c] += b;
}

extern "C" int ddpm_gn_bwd(const ddpm_tensor* x, int dtype, int groups, const double* stats, const float* gamma,
                           const float* beta, float eps, int act, float p_drop, const uint64_t* rng,
                           uint32_t layer_id, const ddpm_tensor* dy, const ddpm_tensor* dx, int accumulate,
                           float* dgamma, float* dbeta, float* ws, void* stream) {
    if (!tensor_ok(x) || !tensor_ok(dy) || !tensor_ok(dx) || !stats || !gamma || !beta || !ws) return DDPM_E_ARG;
    if (groups <= 0 || x->C % groups || dy->C != x->C || dx->C != x->C) return DDPM_E_ARG;
    if (p_drop > 0.f && !rng) return DDPM_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(ws, 0, sizeof(float) * 2 * x->C * x->N, st));
    TV v(*x), d(*dy), o(*dx);
    int HW = x->H * x->W, C = x->C;
#define GO(T, VEC) { int cvs = C / VEC; if (cvs > NT) return DDPM_E_ARG; int ppi = NT / cvs; \
        dim3 grid(blocks_per_image(x->N, HW, ppi * 4), x->N); \
        gn_bwd_reduce_kernel<T, VEC><<<grid, NT, sizeof(float) * 6 * C, st>>>(v, d, groups, stats, gamma, beta, eps, act, p_drop, rng, layer_id, ws); \
        LAUNCH_OK(); \
        gn_bwd_apply_kernel<T, VEC><<<grid, NT, sizeof(float) * (6 * C + 2 * groups), st>>>(v, d, o, groups, stats, gamma, beta, eps, act, p_drop, rng, layer_id, ws, accumulate); \
        LAUNCH_OK(); }
    if (dtype == DDPM_BF16) { if (vec_ok(x, 8, 2) && vec_ok(dy, 8, 2) && vec_ok(dx, 8, 2)) GO(bf16, 8) else GO(bf16, 1) }
    else if (dtype == DDPM_F32) { if (vec_ok(x, 4, 4) && vec_ok(dy, 4, 4) && vec_ok(dx, 4, 4)) GO(float, 4) else GO(float, 1) }
    else return DDPM_E_ARG;
#undef GO
    if (dgamma || dbeta) {
        gn_param_grad_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, x->N, C, dgamma, dbeta);
        LAUNCH_OK();
    }
    return 0;
}

// ------------------------------------------------------------------------------------ pixel maps
// generic walker over (n, y, x, cv) of the OUTPUT view
template <typename T, int VEC, typename F>
__global__ void __launch_bounds__(NT) pix_kernel(int N, int H, int W, int C, F f) {
    const int cvs = C / VEC;
    const int64_t total = (int64_t)N * H * W * cvs;
    for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
        int cv = (int)(i % cvs); int64_t r = i / cvs;
        int x = (int)(r % W); r /= W;
        int y = (int)(r % H); int n = (int)(r / H);
        f(n, y, x, cv * VEC);
    }
}
static inline int pix_grid(int64_t total) {
    int64_t g = (total + NT - 1) / NT;
    return (int)(g < 148 * 8 ? (g < 1 ? 1 : g) : 148 * 8);
}

template <typename T, int VEC> struct UpFwd {
    TV in, out;
    __device__ void operator()(int n, int y, int x, int c) const {
        float v[VEC];
        ldv<T, VEC>(in.at<T>(n, y >> 1, x >> 1, c), v);
        stv<T, VEC>(out.at<T>(n, y, x, c), v);
    }
};
extern "C" int ddpm_upsample2x(const ddpm_tensor* x, const ddpm_tensor* out, int dtype, void* stream) {
    if (!tensor_ok(x) || !tensor_ok(out) || out->H != 2 * x->H || out->W != 2 * x->W || out->C != x->C || out->N != x->N) return DDPM_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
#define GO(T, VEC) { UpFwd<T, VEC> f{TV(*x), TV(*out)}; int64_t tot = (int64_t)out->N * out->H * out->W * (out->C / VEC); \
        pix_kernel<T, VEC><<<pix_grid(tot), NT, 0, st>>>(out->N, out->H, out->W, out->C, f); }
    if (dtype == DDPM_BF16) { if (vec_ok(x, 8, 2) && vec_ok(out, 8, 2)) GO(bf16, 8) else GO(bf16, 1) }
    else if (dtype == DDPM_F32) { if (vec_ok(x, 4, 4) && vec_ok(out, 4, 4)) GO(float, 4) else GO(float, 1) }
    else return DDPM_E_ARG;
#undef GO
    LAUNCH_OK();
    return 0;
}

template <typename T, int VEC> struct UpBwd {
    TV dy, dx; int acc;
    __device__ void operator()(int n, int y, int x, int c) const {
        float s[VEC], v[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) s[i] = 0.f;
        if (acc) ldv<T, VEC>(dx.at<T>(n, y, x, c), s);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ldv<T, VEC>(dy.at<T>(n, 2 * y + (j >> 1), 2 * x + (j & 1), c), v);
#pragma unroll
            for (int i = 0; i < VEC; ++i) s[i] += v[i];
        }
        stv<T, VEC>(dx.at<T>(n, y, x, c), s);
    }
};
extern "C" int ddpm_upsample2x_bwd(const ddpm_tensor* dy, const ddpm_tensor* dx, int dtype, int accumulate, void* stream) {
    if (!tensor_ok(dy) || !tensor_ok(dx) || dy->H != 2 * dx->H || dy->W != 2 * dx->W || dy->C != dx->C || dy->N != dx->N) return DDPM_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
#define GO(T, VEC) { UpBwd<T, VEC> f{TV(*dy), TV(*dx), accumulate}; int64_t tot = (int64_t)dx->N * dx->H * dx->W * (dx->C / VEC); \
        pix_kernel<T, VEC><<<pix_grid(tot), NT, 0, st>>>(dx->N, dx->H, dx->W, dx->C, f); }
    if (dtype == DDPM_BF16) { if (vec_ok(dy, 8, 2) && vec_ok(dx, 8, 2)) GO(bf16, 8) else GO(bf16, 1) }
    else if (dtype == DDPM_F32) { if (vec_ok(dy, 4, 4) && vec_ok(dx, 4, 4)) GO(float, 4) else GO(float, 1) }
    else return DDPM_E_ARG;
#undef GO
    LAUNCH_OK();
    return 0;
}

template <typename T, int VEC> struct ZeroUp {
    TV in, out;
    __device__ void operator()(int n, int y, int x, int c) const {
        float v[VEC];
        if (((y | x) & 1) == 0) ldv<T, VEC>(in.at<T>(n, y >> 1, x >> 1, c), v);
        else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) v[i] = 0.f;
        }
        stv<T, VEC>(out.at<T>(n, y, x, c), v);
    }
};
extern "C" int ddpm_zero_upsample2x(const ddpm_tensor* x, const ddpm_tensor* out, int dtype, void* stream) {
    if (!tensor_ok(x) || !tensor_ok(out) || out->H != 2 * x->H || out->W != 2 * x->W || out->C != x->C || out->N != x->N) return DDPM_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
#define GO(T, VEC) { ZeroUp<T, VEC> f{TV(*x), TV(*out)}; int64_t tot = (int64_t)out->N * out->H * out->W * (out->C / VEC); \
        pix_kernel<T, VEC><<<pix_grid(tot), NT, 0, st>>>(out->N, out->H, out->W, out->C, f); }
    if (dtype == DDPM_BF16) { if (vec_ok(x, 8, 2) && vec_ok(out, 8, 2)) GO(bf16, 8) else GO(bf16, 1) }
    else if (dtype == DDPM_F32) { if (vec_ok(x, 4, 4) && vec_ok(out, 4, 4)) GO(float, 4) else GO(float, 1) }
    else return DDPM_E_ARG;
#undef GO
    LAUNCH_OK();
    return 0;
}

template <typename T, int VEC> struct AddOp {
    TV a, b, o;
    __device__ void operator()(int n, int y, int x, int c) const {
        float u[VEC], v[VEC];
        ldv<T, VEC>(a.at<T>(n, y, x, c), u);
        ldv<T, VEC>(b.at<T>(n, y, x, c), v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) u[i] += v[i];
        stv<T, VEC>(o.at<T>(n, y, x, c), u);
    }
};
extern "C" int ddpm_add(const ddpm_tensor* a, const ddpm_tensor* b, const ddpm_tensor* out, int dtype, void* stream) {
    if (!tensor_ok(a) || !tensor_ok(b) || !tensor_ok(out)) return DDPM_E_ARG;
    if (a->N != out->N || a->H != out->H || a->W != out->W || a->C != out->C) return DDPM_E_ARG;
    if (b->N != out->N || b->H != out->H || b->W != out->W || b->C != out->C) return DDPM_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
#define GO(T, VEC) { AddOp<T, VEC> f{TV(*a), TV(*b), TV(*out)}; int64_t tot = (int64_t)out->N * out->H * out->W * (out->C / VEC); \
        pix_kernel<T, VEC><<<pix_grid(tot), NT, 0, st>>>(out->N, out->H, out->W, out->C, f); }
    if (dtype == DDPM_BF16) { if (vec_ok(a, 8, 2) && vec_ok(b, 8, 2) && vec_ok(out, 8, 2)) GO(bf16, 8) else GO(bf16, 1) }
    else if (dtype == DDPM_F32) { if (vec_ok(a, 4, 4) && vec_ok(b, 4, 4) && vec_ok(out, 4, 4)) GO(float, 4) else GO(float, 1) }
    else return DDPM_E_ARG;
#undef GO
    LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------ colsum
template <typename T, int VEC>
__global__ void __launch_bounds__(NT) colsum_kernel(TV dy, float* out_nc, float* dbias) {
    extern __shared__ float acc[];           // [C]
    const int n = blockIdx.y, C = dy.C, HW = dy.H * dy.W;
    for (int c = threadIdx.x; c < C; c += NT) acc[c] = 0.f;
    __syncthreads();
    PixMap m = make_map<VEC>(C);
    if (m.active) {
        float s[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) s[i] = 0.f;
        const int per = (HW + gridDim.x - 1) / gridDim.x;
        const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
        for (int p = p0 + m.prow; p < p1; p += m.ppi) {
            int y = p / dy.W, xx = p - y * dy.W;
            float v[VEC];
            ldv<T, VEC>(dy.at<T>(n, y, xx, m.cv * VEC), v);
#pragma unroll
            for (int i = 0; i < VEC; ++i) s[i] += v[i];
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) atomicAdd(&acc[m.cv * VEC + i], s[i]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += NT) {
        if (out_nc) atomicAdd(&out_nc[(size_t)n * C + c], acc[c]);
        if (dbias) atomicAdd(&dbias[c], acc[c]);
    }
}
extern "C" int ddpm_colsum(const ddpm_tensor* dy, int dtype, float* out_nc, float* dbias, void* stream) {
    if (!tensor_ok(dy) || (!out_nc && !dbias)) return DDPM_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (out_nc) CUDA_TRY(cudaMemsetAsync(out_nc, 0, sizeof(float) * dy->N * dy->C, st));
    TV v(*dy);
    int HW = dy->H * dy->W;
#define GO(T, VEC) { int cvs = dy->C / VEC; if (cvs > NT) return DDPM_E_ARG; int ppi = NT / cvs; \
        dim3 grid(blocks_per_image(dy->N, HW, ppi * 8), dy->N); \
        colsum_kernel<T, VEC><<<grid, NT, sizeof(float) * dy->C, st>>>(v, out_nc, dbias); }
    if (dtype == DDPM_BF16) { if (vec_ok(dy, 8, 2)) GO(bf16, 8) else GO(bf16, 1) }
    else if (dtype == DDPM_F32) { if (vec_ok(dy, 4, 4)) GO(float, 4) else GO(float, 1) }
    else return DDPM_E_ARG;
#undef GO
    LAUNCH_OK();
    return 0;
}
