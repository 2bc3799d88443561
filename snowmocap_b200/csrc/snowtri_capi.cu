// C ABI of the triangulation engine (declared in include/snowtri.h).  Host-side glue only:
// argument checks, shared-memory layout, launches.  No exception leaves this file.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "snowtri.h"
#include "snowtri_internal.h"
#include "snowtri_kernels.cuh"

using namespace snowtri;

static char g_err[512] = "";
char* snowtri_global_error() { return g_err; }

static void inv3(const double* m, double* o) {
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    o[0] = (e * i - f * h) / det; o[1] = (c * h - b * i) / det; o[2] = (b * f - c * e) / det;
    o[3] = (f * g - d * i) / det; o[4] = (a * i - c * g) / det; o[5] = (c * d - a * f) / det;
    o[6] = (d * h - e * g) / det; o[7] = (b * g - a * h) / det; o[8] = (a * e - b * d) / det;
}

static void default_params(Params* p) {
    // defaults of the reference signatures (triangulation.py:50, 95-100)
    p->kst = 0.5; p->ast = 0.0; p->dthr = 0.05; p->cond_tol = 0.1; p->score_tol = 0.0;
    p->kst_f = 0.5f; p->num_tol = 0; p->center = 18;
}

extern "C" int snowtri_version(void) { return 1; }

extern "C" const char* snowtri_last_error(snowtri_t* h) { return h ? h->err : g_err; }

extern "C" int snowtri_create(snowtri_t** out, int device, int C, const double* K, const double* R,
                              const double* t) {
    if (!out || !K || !R || !t) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_create: NULL argument");
    if (C < 1 || C > 255) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_create: C=%d outside [1,255]", C);
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, SNOWTRI_E_CUDA, "snowtri_create: no CUDA device (%s); this library has no CPU path",
                    cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_create: device %d of %d", device, ndev);
    snowtri_t* h = (snowtri_t*)calloc(1, sizeof(snowtri_t));
    if (!h) return fail(nullptr, SNOWTRI_E_NOMEM, "snowtri_create: out of host memory");
    h->device = device;
    h->C = C;
    default_params(&h->prm);
    h->jit_mode = 1;
    snprintf(h->jit_status, sizeof(h->jit_status), "not used");
    h->precision = SNOWTRI_PREC_F64;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        free(h);
        return fail(nullptr, SNOWTRI_E_CUDA, "snowtri_create: %s", cudaGetErrorString(e));
    }
    if (prop.major < 10) {
        free(h);
        return fail(nullptr, SNOWTRI_E_UNSUPPORTED, "snowtri_create: device sm_%d%d, this build is sm_100a only",
                    prop.major, prop.minor);
    }
    h->sm_count = prop.multiProcessorCount;
    h->total_mem = (size_t)prop.totalGlobalMem;
    h->max_smem = (int)prop.sharedMemPerBlockOptin;
    h->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
    double* cam = (double*)malloc(sizeof(double) * 12 * C);
    h->kinv_host = (double*)malloc(sizeof(double) * 9 * C);
    h->r_host = (double*)malloc(sizeof(double) * 9 * C);
    memcpy(h->r_host, R, sizeof(double) * 9 * C);
    for (int c = 0; c < C; ++c) {
        double* Kinv = h->kinv_host + 9 * c;
        inv3(K + 9 * c, Kinv);
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k) {
                double s = 0;
                for (int m = 0; m < 3; ++m) s += R[9 * c + 3 * r + m] * Kinv[3 * m + k];
                cam[12 * c + 3 * r + k] = s;
            }
        for (int k = 0; k < 3; ++k) cam[12 * c + 9 + k] = t[3 * c + k];
    }
    e = cudaMalloc(&h->d_cam, sizeof(double) * 12 * C);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_cam, cam, sizeof(double) * 12 * C, cudaMemcpyHostToDevice);
    h->cam_host = cam;
    if (e != cudaSuccess) {
        if (h->d_cam) cudaFree(h->d_cam);
        free(cam);
        free(h->kinv_host);
        free(h->r_host);
        free(h);
        return fail(nullptr, SNOWTRI_E_CUDA, "snowtri_create: %s", cudaGetErrorString(e));
    }
    cudaFuncSetAttribute(condense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem);
    *out = h;
    return SNOWTRI_OK;
}

extern "C" int snowtri_destroy(snowtri_t* h) {
    if (!h) return SNOWTRI_OK;
    cudaSetDevice(h->device);
    for (int i = 0; i < 6; ++i)
        if (h->stage[i]) cudaFree(h->stage[i]);
    if (h->gen_scratch) cudaFree(h->gen_scratch);
    snowtri_jit_free(h);
    snowtri_comm_destroy(h);
    free(h->p1_args);
    free(h->mf_args);
    if (h->pipe_in) {
        cudaStreamDestroy(h->pipe_in);
        cudaStreamDestroy(h->pipe_k);
        cudaStreamDestroy(h->pipe_out);
        for (int i = 0; i < SNOWTRI_PIPE_EVENTS; ++i) {
            cudaEventDestroy(h->pipe_ev[i]);
            cudaEventDestroy(h->pipe_evh[i]);
        }
        cudaEventDestroy(h->pipe_start);
    }
    if (h->d_cam) cudaFree(h->d_cam);
    free(h->cam_host);
    free(h->kinv_host);
    free(h->r_host);
    free(h);
    return SNOWTRI_OK;
}

extern "C" int snowtri_set_params(snowtri_t* h, double kst, double ast, double dthr, double cond_tol, int num_tol,
                                  double score_tol, int center) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_set_params: NULL handle");
    if (center < 0) return fail(h, SNOWTRI_E_ARG, "snowtri_set_params: center_point_index %d < 0", center);
    h->prm.kst = kst; h->prm.ast = ast; h->prm.dthr = dthr; h->prm.cond_tol = cond_tol;
    h->prm.num_tol = num_tol; h->prm.score_tol = score_tol; h->prm.center = center;
    float kf = (float)kst;
    if ((double)kf < kst) kf = nextafterf(kf, INFINITY);
    h->prm.kst_f = kf;
    return SNOWTRI_OK;
}

extern "C" int snowtri_set_precision(snowtri_t* h, int precision) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_set_precision: NULL handle");
    if (precision != SNOWTRI_PREC_F64 && precision != SNOWTRI_PREC_F32 && precision != SNOWTRI_PREC_MIXED &&
        precision != SNOWTRI_PREC_F32_EXPERIMENTAL)
        return fail(h, SNOWTRI_E_ARG, "snowtri_set_precision: unknown precision %d", precision);
    h->allow_f32_multi = precision == SNOWTRI_PREC_F32_EXPERIMENTAL ? 1 : 0;
    h->precision = precision == SNOWTRI_PREC_F32_EXPERIMENTAL ? SNOWTRI_PREC_F32 : precision;
    return SNOWTRI_OK;
}

extern "C" int snowtri_set_tuning(snowtri_t* h, int frames_per_group, int max_ctas, int threads) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_set_tuning: NULL handle");
    if (threads != 0 && threads != 256 && threads != 512 && threads != -256 && threads != -1)
        return fail(h, SNOWTRI_E_ARG, "snowtri_set_tuning: threads must be 0 (auto), 256, 512, -256 or -1");
    h->no_p1 = threads != 0 ? 1 : 0;         /* any explicit block size selects the general fused kernel */
    if (threads == -1) threads = 0;
    h->tune_G = frames_per_group > 0 ? frames_per_group : 0;
    h->tune_ctas = max_ctas > 0 ? max_ctas : 0;
    h->no_fly = threads == -256 ? 1 : 0;   /* -256: 256 threads with stored rays even when P == 1 */
    h->tune_threads = threads == -256 ? 256 : threads;
    return SNOWTRI_OK;
}

extern "C" int snowtri_set_general_kernels(snowtri_t* h, int generation) {
    if (!h || (generation != 1 && generation != 2))
        return fail(h, SNOWTRI_E_ARG, "snowtri_set_general_kernels: generation must be 1 or 2");
    h->gen1_only = generation == 1 ? 1 : 0;
    return SNOWTRI_OK;
}

extern "C" int snowtri_set_jit(snowtri_t* h, int mode) {
    if (!h || mode < 0 || mode > 2) return fail(h, SNOWTRI_E_ARG, "snowtri_set_jit: mode must be 0 (off), 1 (auto) or 2 (always)");
    h->jit_mode = mode;
    return SNOWTRI_OK;
}

extern "C" const char* snowtri_jit_status(snowtri_t* h) { return h ? h->jit_status : ""; }

extern "C" int snowtri_set_pipeline(snowtri_t* h, int frames_per_chunk) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_set_pipeline: NULL handle");
    h->tune_chunk = frames_per_chunk > 0 ? frames_per_chunk : 0;
    return SNOWTRI_OK;
}

extern "C" long long snowtri_launch_count(snowtri_t* h) { return h ? h->launches : 0; }

extern "C" const char* snowtri_last_kernel(snowtri_t* h) {
    if (!h || h->launches == 0) return "";
    if (h->last_fly == 4) return "p1-jit";
    if (h->last_fly == 3) return h->last_gen2 == 3 ? "general2" : (h->last_gen2 == 1 ? "general2m" : "general");
    return h->last_fly == 2 ? "p1" : (h->last_fly == 1 ? "fused-fly" : "fused");
}

extern "C" int snowtri_last_launch_info(snowtri_t* h, int* grid, int* block, int* smem_bytes, int* frames_per_group) {
    if (!h) return SNOWTRI_E_ARG;
    if (grid) *grid = h->last_grid;
    if (block) *block = h->last_block;
    if (smem_bytes) *smem_bytes = h->last_smem;
    if (frames_per_group) *frames_per_group = h->last_G;
    return SNOWTRI_OK;
}

// Shared-memory layout of fused_kernel for G frames per group; returns total bytes.
// fly: rays are not stored, the (u,v,score) staging is double-buffered instead.
static size_t fused_layout(int C, int P, int J, int Jout, int Pout, int G, size_t tsz, bool fly, int nthreads,
                           FusedSmem* L) {
    const size_t npairs = (size_t)C * (C - 1) / 2, ncand = npairs * P * P, R = (size_t)C * P * J;
    size_t o = 16;  // two mbarriers
    auto take = [&](size_t bytes, size_t align) -> int {
        o = (o + align - 1) / align * align;
        const size_t r = o;
        o += bytes;
        return (int)r;
    };
    memset(L, 0, sizeof(*L));
    L->cam = take((size_t)C * 9 * tsz, 16);
    L->pairs = take(npairs * 2, 4);
    L->pd = take(npairs * 6 * tsz, 16);
    const size_t uvb = (G * R * 8 + 127) / 128 * 128, sb = (G * R * 4 + 127) / 128 * 128;
    L->stage_uv = take(uvb * (fly ? 2 : 1), 128);
    L->stage_s = take(sb * (fly ? 2 : 1), 128);
    L->stage_stride_uv = fly ? (int)uvb : 0;
    L->stage_stride_s = fly ? (int)sb : 0;
    if (!fly) {
        L->hx = take(G * R * tsz, 16);
        L->hy = take(G * R * tsz, 16);
        L->hz = take(G * R * tsz, 16);
        L->sc = take(G * R * 4, 16);
    }
    L->cnt = take((size_t)G * C * 4, 4);
    L->cen = take(G * ncand * 24, 8);
    L->keep = take(G * ncand, 4);
    L->ab = take(G * ncand, 4);
    L->klist = take(G * ncand * 4, 4);
    L->memb = take(G * ncand * 4, 4);
    L->membp = take(G * ncand * 4, 4);
    L->cstart = take(G * ncand * 4, 4);
    L->cn = take(G * ncand * 4, 4);
    L->ksum = take(G * ncand * 8, 8);
    L->slot = take(G * ncand * 4, 4);
    L->kcount = take((size_t)G * 2 * 4, 4);
    L->ks = take((size_t)G * Pout * Jout * tsz, 16);
    L->cobs = take((size_t)G * Pout * kCliqueMax, 4);
    L->clq = take((size_t)G * Pout, 4);
    L->wtmp = take((size_t)(2 * (nthreads / 32) + 4) * 4, 4);
    L->total = (int)o;
    return o;
}

template <typename T, int NT, int CM, int NCH, bool FLY>
static cudaError_t launch_fused(const FusedArgs<T>& a, int grid, cudaStream_t st) {
    auto kern = fused_kernel<T, NT, CM, NCH, FLY>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a.sm.total);
    if (e != cudaSuccess) return e;
    kern<<<grid, NT, a.sm.total, st>>>(a);
    return cudaGetLastError();
}

template <typename T, int NT, int CM, bool FLY>
static cudaError_t launch_fused_nch(const FusedArgs<T>& a, int nch, int grid, cudaStream_t st) {
    return nch == 5 ? launch_fused<T, NT, CM, 5, FLY>(a, grid, st) : launch_fused<T, NT, CM, 0, FLY>(a, grid, st);
}

template <typename T>
static cudaError_t launch_fused_any(const FusedArgs<T>& a, int nt, int cm, int nch, bool fly, int grid, cudaStream_t st) {
    if (fly) {  // fly mode exists for 256 threads and the <= 8-camera tables only
        return cm == 4 ? launch_fused_nch<T, 256, 4, true>(a, nch, grid, st)
                       : launch_fused_nch<T, 256, 8, true>(a, nch, grid, st);
    }
    if (nt == 256) {
        switch (cm) {
            case 4: return launch_fused_nch<T, 256, 4, false>(a, nch, grid, st);
            case 8: return launch_fused_nch<T, 256, 8, false>(a, nch, grid, st);
            default: return launch_fused_nch<T, 256, 0, false>(a, nch, grid, st);
        }
    }
    switch (cm) {
        case 4: return launch_fused_nch<T, 512, 4, false>(a, nch, grid, st);
        case 8: return launch_fused_nch<T, 512, 8, false>(a, nch, grid, st);
        default: return launch_fused_nch<T, 512, 0, false>(a, nch, grid, st);
    }
}

template <typename T>
static int run_fused(snowtri_t* h, const float* d_kpts, const float* d_scores, const int* d_counts, int F, int P,
                     int J, int keypoint_num, int Pout, float* d_out, float* d_pscores, int* d_nout, void* stream) {
    const int C = h->C;
    const size_t tsz = sizeof(T);
    FusedArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.kpts = d_kpts; a.scores = d_scores; a.counts = d_counts;
    a.out = d_out; a.pscores = d_pscores; a.nout = d_nout; a.cam = h->d_cam;
    a.F = F; a.C = C; a.P = P; a.J = J; a.Jout = keypoint_num; a.Pout = Pout;
    a.npairs = C * (C - 1) / 2;
    a.ncand = a.npairs * P * P;
    a.R = C * P * J;
    a.prm = h->prm;
    a.all_kept = (h->prm.ast <= 0.0 && h->prm.kst >= 0.0) ? 1 : 0;
    a.never_filter = (h->prm.score_tol <= 0.0 && h->prm.kst >= 0.0) ? 1 : 0;
    a.inv_dthr = h->prm.dthr > 0.0 ? (T)(1.0 / h->prm.dthr) : (T)INFINITY;
    a.tol2 = h->prm.cond_tol >= 0.0 ? h->prm.cond_tol * h->prm.cond_tol : -1.0;

    // template selection
    const int cm = C <= 4 ? 4 : (C <= kCliqueMax ? 8 : 0);
    const int nch = (J > 32 && J <= 160) ? 5 : 0;
    for (int x = 0; x < cm - 1 && x < C; ++x)
        for (int y = x + 1; y < cm && y < C; ++y) {
            const int e = (x * cm - x * (x + 1) / 2 + y - x - 1) * 6;
            for (int k = 0; k < 3; ++k) {
                a.pdc[e + k] = (T)(h->cam_host[12 * y + 9 + k] - h->cam_host[12 * x + 9 + k]);
                a.pdc[e + 3 + k] = (T)((h->cam_host[12 * x + 9 + k] + h->cam_host[12 * y + 9 + k]) / 2);
            }
        }
    for (int c = 0; c < C && c < kCliqueMax; ++c)
        for (int k = 0; k < 9; ++k) a.camc[9 * c + k] = (T)h->cam_host[12 * c + k];

    // Block size / frames per group.  256 threads x 2 CTAs per SM when two groups fit in shared
    // memory side by side (one CTA's barriers overlap the other's math), else 512 x 1.  With one
    // person per camera (and <= 8 cameras) rays are recomputed on the fly instead of stored.
    const size_t full = (size_t)h->max_smem, half = ((size_t)h->smem_per_sm - 2048) / 2 - 1024;
    int cap = h->tune_G > 0 ? h->tune_G : 32;
    if (cap > F) cap = F;
    auto largest_G = [&](size_t budget, int cap_, bool fly, int nt_) -> int {
        for (int g = cap_; g >= 1; --g)
            if ((size_t)g * a.R <= 65535 && fused_layout(C, P, J, a.Jout, Pout, g, tsz, fly, nt_, &a.sm) <= budget) return g;
        return 0;
    };
    bool fly = (P == 1 && cm > 0 && h->tune_threads != 512 && !h->no_fly);
    int nt = h->tune_threads, G = 0;
    if (fly) {
        G = largest_G(half, cap, true, 256);
        if (G == 0) fly = false;
        else nt = 256;
    }
    if (!fly) {
        if (nt == 0) {
            const int g2 = largest_G(half, cap, false, 256);
            if (g2 >= 1 && (size_t)g2 * a.R >= 2048) { nt = 256; G = g2; }
            else { nt = 512; G = largest_G(full, cap, false, 512); }
        } else {
            G = largest_G(nt == 256 ? half : full, cap, false, nt);
            if (G == 0 && nt == 256) G = largest_G(full, cap, false, nt);
        }
    }
    if (G == 0)
        return fail(h, SNOWTRI_E_UNSUPPORTED,
                    "snowtri_run: one frame (C=%d P=%d J=%d Pout=%d) needs %zu B of shared memory, device allows %d",
                    C, P, J, Pout, fused_layout(C, P, J, a.Jout, Pout, 1, tsz, false, 512, &a.sm), h->max_smem);
    const int ctas_per_sm = (nt == 256 && fused_layout(C, P, J, a.Jout, Pout, G, tsz, fly, nt, &a.sm) <= half) ? 2 : 1;
    if (h->tune_G == 0) {
        // keep a few groups per CTA when the batch is large enough; warp-per-frame phases like
        // G to be a multiple of the warp count; TMA staging needs (G*R) % 4 == 0
        while (G > 1 && (F + G - 1) / G < 4 * h->sm_count * ctas_per_sm) --G;
        const int nw = nt / 32;
        if (a.ncand <= 64 && G > nw) G = G / nw * nw;
        for (int g = G; g >= 1 && g > G - 4; --g)
            if (((size_t)g * a.R) % 4 == 0) { G = g; break; }
    }
    fused_layout(C, P, J, a.Jout, Pout, G, tsz, fly, nt, &a.sm);
    a.G = G;
    const bool aligned = (((uintptr_t)d_kpts & 15u) == 0) && (((uintptr_t)d_scores & 15u) == 0);
    a.use_tma = (aligned && ((size_t)G * a.R) % 4 == 0) ? 1 : 0;

    const int ngroups = (F + G - 1) / G;
    int grid = h->sm_count * ctas_per_sm;
    if (grid > ngroups) grid = ngroups;
    if (h->tune_ctas > 0 && grid > h->tune_ctas) grid = h->tune_ctas;
    const cudaError_t e = launch_fused_any<T>(a, nt, cm, nch, fly, grid, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(h, SNOWTRI_E_CUDA, "fused_kernel launch failed: %s", cudaGetErrorString(e));
    h->launches += 1;
    h->last_grid = grid; h->last_block = nt; h->last_smem = a.sm.total; h->last_G = G;
    h->last_fly = fly ? 1 : 0;
    return SNOWTRI_OK;
}

extern "C" int snowtri_run(snowtri_t* h, const float* d_kpts, const float* d_scores, const int* d_counts, int F,
                           int P, int J, int keypoint_num, int Pout, float* d_out, float* d_pscores, int* d_nout,
                           void* stream) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_run: NULL handle");
    if (F == 0) return SNOWTRI_OK;
    if (!d_kpts || !d_scores || !d_out || !d_pscores || !d_nout) return fail(h, SNOWTRI_E_ARG, "snowtri_run: NULL buffer");
    if (F < 0 || P < 1 || P > 255 || J < 1 || Pout < 1)
        return fail(h, SNOWTRI_E_ARG, "snowtri_run: bad sizes F=%d P=%d J=%d Pout=%d", F, P, J, Pout);
    if (keypoint_num < 1 || keypoint_num > J)
        return fail(h, SNOWTRI_E_ARG, "snowtri_run: keypoint_num=%d must be in [1, J=%d] (the reference raises IndexError above J)",
                    keypoint_num, J);
    if (h->prm.center >= J)
        return fail(h, SNOWTRI_E_ARG, "snowtri_run: center_point_index=%d >= J=%d (the reference raises IndexError)",
                    h->prm.center, J);
    if (((uintptr_t)d_out & 15u) != 0) return fail(h, SNOWTRI_E_ARG, "snowtri_run: d_out must be 16-byte aligned");
    if (((uintptr_t)d_kpts & 7u) != 0) return fail(h, SNOWTRI_E_ARG, "snowtri_run: d_kpts must be 8-byte aligned");
    CUDA_TRY(h, cudaSetDevice(h->device));

    if (snowtri_p1_eligible(h, P, Pout, keypoint_num))
        return snowtri_p1_run(h, d_kpts, d_scores, d_counts, F, J, keypoint_num, Pout, d_out, d_pscores, d_nout, stream);
    if (snowtri_general_eligible(h))
        return snowtri_general_run(h, d_kpts, d_scores, d_counts, F, P, J, keypoint_num, Pout, d_out, d_pscores, d_nout, stream);
    // (fused_kernel: explicit block size requested, or a positive condense_score_tol / kst < 0)
    // The general kernel computes in float32 only on request AND for one person per camera: with several
    // persons, wrongly matched ("ghost") clusters fuse midpoints that lie metres apart, so the float32
    // error of the 1/distance weights (1e-4..1e-3) moves their joints beyond the 1e-4 parity bound.
    // (Its float32 keep decision is guarded by a float64 re-evaluation, see fused_kernel phase 1a.)
    // A positive condense_score_tol or kst < 0 also computes in float64: that filter has no guard.
    const bool never_filter = h->prm.score_tol <= 0.0 && h->prm.kst >= 0.0;
    const bool f32_ok = h->precision == SNOWTRI_PREC_F32 && never_filter && (P == 1 || h->allow_f32_multi);
    const int rc = f32_ok
                       ? run_fused<float>(h, d_kpts, d_scores, d_counts, F, P, J, keypoint_num, Pout, d_out, d_pscores, d_nout, stream)
                       : run_fused<double>(h, d_kpts, d_scores, d_counts, F, P, J, keypoint_num, Pout, d_out, d_pscores, d_nout, stream);
    return rc;
}

static int ensure_stage(snowtri_t* h, int i, size_t bytes) {
    if (h->stage_cap[i] >= bytes) return SNOWTRI_OK;
    if (h->stage[i]) cudaFree(h->stage[i]);
    h->stage[i] = nullptr;
    h->stage_cap[i] = 0;
    CUDA_TRY(h, cudaMalloc(&h->stage[i], bytes));
    h->stage_cap[i] = bytes;
    return SNOWTRI_OK;
}

extern "C" int snowtri_run_host(snowtri_t* h, const float* h_kpts, const float* h_scores, const int* h_counts, int F,
                                int P, int J, int keypoint_num, int Pout, float* h_out, float* h_pscores, int* h_nout,
                                void* stream) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_run_host: NULL handle");
    if (F == 0) return SNOWTRI_OK;
    if (!h_kpts || !h_scores || !h_out || !h_pscores || !h_nout || F < 0 || P < 1 || J < 1 || Pout < 1 ||
        keypoint_num < 1)
        return fail(h, SNOWTRI_E_ARG, "snowtri_run_host: bad argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t rays = (size_t)F * h->C * P * J, outs = (size_t)F * Pout * keypoint_num;
    const size_t need[6] = {rays * 8, rays * 4, (size_t)F * h->C * 4, outs * 16, (size_t)F * Pout * 4, (size_t)F * 4};
    for (int i = 0; i < 6; ++i) {
        int rc = ensure_stage(h, i, need[i]);
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // Chunked pipeline over three internal streams: chunk i's kernels (stream `k`) and its device->host copy (stream
    // `out`) overlap the host->device copy of the chunks behind it (stream `in`), so the PCIe link carries both
    // directions at once and never waits for a kernel.  Small batches go through in one chunk.
    if (!h->pipe_in) {
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->pipe_in, cudaStreamNonBlocking));
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->pipe_k, cudaStreamNonBlocking));
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->pipe_out, cudaStreamNonBlocking));
        for (int i = 0; i < SNOWTRI_PIPE_EVENTS; ++i) {
            CUDA_TRY(h, cudaEventCreateWithFlags(&h->pipe_ev[i], cudaEventDisableTiming));
            CUDA_TRY(h, cudaEventCreateWithFlags(&h->pipe_evh[i], cudaEventDisableTiming));
        }
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->pipe_start, cudaEventDisableTiming));
    }
    const size_t in_per_frame = (size_t)h->C * P * J * 12;
    int chunk = h->tune_chunk > 0 ? h->tune_chunk : (int)((size_t)(24u << 20) / in_per_frame);  // ~24 MB of input per chunk
    // several persons per camera: the matching kernel runs one CTA per frame, two per SM -- a chunk should be a few full
    // waves of them (at 24 MB, BASELINE configs[2] had 492-frame chunks = 1.7 waves)
    if (h->tune_chunk <= 0 && P > 1 && chunk < 8 * h->sm_count) chunk = 8 * h->sm_count;
    if (chunk < 1) chunk = 1;
    if (chunk > F) chunk = F;
    const int nchunks = (F + chunk - 1) / chunk;
    CUDA_TRY(h, cudaEventRecord(h->pipe_start, st));            // order the pipeline after the caller's stream
    CUDA_TRY(h, cudaStreamWaitEvent(h->pipe_in, h->pipe_start, 0));
    CUDA_TRY(h, cudaStreamWaitEvent(h->pipe_k, h->pipe_start, 0));
    CUDA_TRY(h, cudaStreamWaitEvent(h->pipe_out, h->pipe_start, 0));
    const size_t rpf = (size_t)h->C * P * J, opf = (size_t)Pout * keypoint_num;
    for (int ci = 0; ci < nchunks; ++ci) {
        const int f0 = ci * chunk, fc = F - f0 < chunk ? F - f0 : chunk;
        float* dk = (float*)h->stage[0] + (size_t)f0 * rpf * 2;
        float* ds = (float*)h->stage[1] + (size_t)f0 * rpf;
        int* dc = (int*)h->stage[2] + (size_t)f0 * h->C;
        float* dout = (float*)h->stage[3] + (size_t)f0 * opf * 4;
        float* dps = (float*)h->stage[4] + (size_t)f0 * Pout;
        int* dn = (int*)h->stage[5] + f0;
        CUDA_TRY(h, cudaMemcpyAsync(dk, h_kpts + (size_t)f0 * rpf * 2, (size_t)fc * rpf * 8, cudaMemcpyHostToDevice, h->pipe_in));
        CUDA_TRY(h, cudaMemcpyAsync(ds, h_scores + (size_t)f0 * rpf, (size_t)fc * rpf * 4, cudaMemcpyHostToDevice, h->pipe_in));
        if (h_counts)
            CUDA_TRY(h, cudaMemcpyAsync(dc, h_counts + (size_t)f0 * h->C, (size_t)fc * h->C * 4, cudaMemcpyHostToDevice, h->pipe_in));
        cudaEvent_t evh = h->pipe_evh[ci % SNOWTRI_PIPE_EVENTS];
        CUDA_TRY(h, cudaEventRecord(evh, h->pipe_in));
        CUDA_TRY(h, cudaStreamWaitEvent(h->pipe_k, evh, 0));
        const int rc = snowtri_run(h, dk, ds, h_counts ? dc : nullptr, fc, P, J, keypoint_num, Pout, dout, dps, dn, h->pipe_k);
        if (rc) return rc;
        cudaEvent_t ev = h->pipe_ev[ci % SNOWTRI_PIPE_EVENTS];
        CUDA_TRY(h, cudaEventRecord(ev, h->pipe_k));
        CUDA_TRY(h, cudaStreamWaitEvent(h->pipe_out, ev, 0));
        CUDA_TRY(h, cudaMemcpyAsync(h_out + (size_t)f0 * opf * 4, dout, (size_t)fc * opf * 16, cudaMemcpyDeviceToHost, h->pipe_out));
        CUDA_TRY(h, cudaMemcpyAsync(h_pscores + (size_t)f0 * Pout, dps, (size_t)fc * Pout * 4, cudaMemcpyDeviceToHost, h->pipe_out));
        CUDA_TRY(h, cudaMemcpyAsync(h_nout + f0, dn, (size_t)fc * 4, cudaMemcpyDeviceToHost, h->pipe_out));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->pipe_out));
    CUDA_TRY(h, cudaStreamSynchronize(h->pipe_k));
    CUDA_TRY(h, cudaStreamSynchronize(h->pipe_in));
    return SNOWTRI_OK;
}

extern "C" int snowtri_candidates(snowtri_t* h, const float* d_kpts, const float* d_scores, const int* d_counts, int F,
                                  int P, int J, double* d_cand, double* d_avg, int* d_keep, void* stream) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_candidates: NULL handle");
    if (F == 0) return SNOWTRI_OK;
    if (!d_kpts || !d_scores || !d_cand || !d_avg || !d_keep || F < 0 || P < 1 || P > 255 || J < 1)
        return fail(h, SNOWTRI_E_ARG, "snowtri_candidates: bad argument");
    if (((uintptr_t)d_cand & 15u) != 0 || ((uintptr_t)d_kpts & 7u) != 0)
        return fail(h, SNOWTRI_E_ARG, "snowtri_candidates: misaligned buffer");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CandArgs a;
    memset(&a, 0, sizeof(a));
    a.cam = h->d_cam;
    a.C = h->C; a.P = P; a.J = J;
    a.npairs = h->C * (h->C - 1) / 2;
    a.ncand = a.npairs * P * P;
    a.prm = h->prm;
    if (a.ncand == 0) return SNOWTRI_OK;
    const int chunk = 32768;
    for (int f0 = 0; f0 < F; f0 += chunk) {
        const int fc = F - f0 < chunk ? F - f0 : chunk;
        const size_t ro = (size_t)f0 * h->C * P * J, co = (size_t)f0 * a.ncand;
        a.kpts = d_kpts + ro * 2;
        a.scores = d_scores + ro;
        a.counts = d_counts ? d_counts + (size_t)f0 * h->C : nullptr;
        a.cand = d_cand + co * J * 4;
        a.avg = d_avg + co;
        a.keep = d_keep + co;
        a.F = fc;
        dim3 grid((a.ncand + kWarps - 1) / kWarps, fc);
        candidates_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(a);
        CUDA_TRY(h, cudaGetLastError());
        h->launches += 1;
    }
    return SNOWTRI_OK;
}

extern "C" int snowtri_condense(snowtri_t* h, const double* d_cand, const int* d_ncand, int F, int N, int J,
                                int keypoint_num, int Pout, double* d_out, double* d_pscores, int* d_nout,
                                void* stream) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_condense: NULL handle");
    if (F == 0) return SNOWTRI_OK;
    if (!d_ncand || !d_out || !d_pscores || !d_nout || F < 0 || N < 0 || J < 1 || Pout < 1 || (N > 0 && !d_cand))
        return fail(h, SNOWTRI_E_ARG, "snowtri_condense: bad argument");
    if (keypoint_num < 1 || keypoint_num > J)
        return fail(h, SNOWTRI_E_ARG, "snowtri_condense: keypoint_num=%d must be in [1, J=%d] (the reference raises IndexError above J)",
                    keypoint_num, J);
    if (N > 1 && h->prm.center >= J)
        return fail(h, SNOWTRI_E_ARG, "snowtri_condense: center_point_index=%d >= J=%d (the reference raises IndexError)",
                    h->prm.center, J);
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t smem = (size_t)N * 29 + 64;
    if (smem > (size_t)h->max_smem)
        return fail(h, SNOWTRI_E_UNSUPPORTED, "snowtri_condense: N=%d candidates need %zu B of shared memory (max %d)", N,
                    smem, h->max_smem);
    CondArgs a;
    a.cand = d_cand; a.ncand = d_ncand; a.out = d_out; a.pscores = d_pscores; a.nout = d_nout;
    a.F = F; a.N = N; a.J = J; a.Jout = keypoint_num; a.Pout = Pout;
    a.prm = h->prm;
    condense_kernel<<<F, kThreads, smem, (cudaStream_t)stream>>>(a);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return SNOWTRI_OK;
}

extern "C" int snowtri_skew_ray(snowtri_t* h, int n, const double* d_hm, const double* d_hs, const double* d_tm,
                                const double* d_ts, double* d_dist, double* d_mid, void* stream) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_skew_ray: NULL handle");
    if (n == 0) return SNOWTRI_OK;
    if (n < 0 || !d_hm || !d_hs || !d_tm || !d_ts || !d_dist || !d_mid) return fail(h, SNOWTRI_E_ARG, "snowtri_skew_ray: bad argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    skew_ray_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, d_hm, d_hs, d_tm, d_ts, d_dist, d_mid);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return SNOWTRI_OK;
}
