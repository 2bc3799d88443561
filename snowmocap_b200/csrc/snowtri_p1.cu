// Host side of the single-person kernel (snowtri_p1.cuh): constant tables, tile size, launch.
#include <math.h>
#include <stddef.h>
#include <string.h>

#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "snowtri_internal.h"
#include "snowtri_p1.cuh"

// defaults of the rig-specialised all-float32 build for up to 4 cameras (see the measurements in DESIGN.md)
#ifndef P1_JIT_DEFAULT_NI
#define P1_JIT_DEFAULT_NI 2
#endif
#ifndef P1_JIT_DEFAULT_MINB
#define P1_JIT_DEFAULT_MINB 2
#endif

using namespace snowtri;

// ---- single-person path (snowtri_p1.cuh) -------------------------------------------------------
template <typename T, typename TD, int C>
static int run_p1_c(snowtri_t* h, const float* d_kpts, const float* d_scores, const int* d_counts, int F, int J,
                    int keypoint_num, int Pout, float* d_out, float* d_pscores, int* d_nout, void* stream) {
    constexpr int NP = C * (C - 1) / 2;
#ifdef P1_NT
    constexpr int NT = P1_NT, NW = NT / 32;
#else
    constexpr int NT = 256, NW = NT / 32;
#endif
    // the argument block is a few KB: it lives in the handle (one handle per thread by contract), not in a static
    if (h->p1_args_bytes < sizeof(P1Args<T, C>)) {
        free(h->p1_args);
        h->p1_args = aligned_alloc(64, (sizeof(P1Args<T, C>) + 63) & ~(size_t)63);  // the block has 32-byte aligned members
        h->p1_args_bytes = h->p1_args ? sizeof(P1Args<T, C>) : 0;
        if (!h->p1_args) return fail(h, SNOWTRI_E_NOMEM, "snowtri_run: out of host memory");
    }
    typedef P1Args<T, C> P1ArgsT;
    P1Args<T, C>& a = *reinterpret_cast<P1Args<T, C>*>(h->p1_args);
    memset(&a, 0, sizeof(a));
    a.kpts = d_kpts; a.scores = d_scores; a.counts = d_counts;
    a.out = d_out; a.pscores = d_pscores; a.nout = d_nout;
    a.F = F; a.J = J; a.Jout = keypoint_num; a.Pout = Pout;
    a.jmagic = p1_div_magic(keypoint_num);
    a.center = h->prm.center; a.num_tol = h->prm.num_tol; a.kst_f = h->prm.kst_f;
    const double inv = h->prm.dthr > 0.0 ? 1.0 / h->prm.dthr : (double)INFINITY;
    a.inv_dthr = (T)inv;
    a.inv_dthr64 = inv;
    const double band = sizeof(TD) == 8 ? kGuardBandMixed : kGuardBandF32;
    a.guard_lo = isinf(inv) ? (T)INFINITY : (T)(inv * (1.0 - band));
    a.guard_hi = isinf(inv) ? (T)INFINITY : (T)(inv * (1.0 + band));
    a.guard_w = (isinf(inv) || sizeof(T) == 8) ? 0.f : (float)(inv * band);
    a.tol2 = h->prm.cond_tol >= 0.0 ? h->prm.cond_tol * h->prm.cond_tol : -1.0;
    a.kscale[0] = (T)0;
    for (int n = 1; n <= NP; ++n) a.kscale[n] = (T)(0.0005 / (double)n);
    for (int c = 0; c < C; ++c)
        for (int k = 0; k < 9; ++k) {
            a.cam64[12 * c + (k / 3) * 4 + k % 3] = h->cam_host[12 * c + k];
            a.camc[12 * c + (k / 3) * 4 + k % 3] = (T)h->cam_host[12 * c + k];
        }
    for (int x = 0; x < C - 1; ++x)
        for (int y = x + 1; y < C; ++y) {
            const int e = pair_index(C, x, y);
            a.px[e] = (unsigned char)x;
            a.py[e] = (unsigned char)y;
            for (int k = 0; k < 3; ++k) {
                const double tm = h->cam_host[12 * x + 9 + k], ts = h->cam_host[12 * y + 9 + k];
                a.pd64[e * 8 + k] = ts - tm;
                a.pd64[e * 8 + 4 + k] = (tm + ts) / 2;
                a.pdc[e * 8 + k] = (T)(ts - tm);
                a.pdc[e * 8 + 4 + k] = (T)((tm + ts) / 2);
            }
            // d.(hm x hs) = -hm^T [d]x hs  with hm = Mx [u v 1]^T, hs = My [u v 1]^T   =>   E = -Mx^T [d]x My
            const double* cam = h->cam_host;
            long double d[3];
            for (int k = 0; k < 3; ++k) d[k] = (long double)cam[12 * y + 9 + k] - (long double)cam[12 * x + 9 + k];
            const long double dx[9] = {0, -d[2], d[1], d[2], 0, -d[0], -d[1], d[0], 0};
            for (int i = 0; i < 3; ++i)
                for (int j2 = 0; j2 < 3; ++j2) {
                    long double v = 0;
                    for (int r = 0; r < 3; ++r)
                        for (int q = 0; q < 3; ++q) v += (long double)cam[12 * x + 3 * r + i] * dx[3 * r + q] * (long double)cam[12 * y + 3 * q + j2];
                    a.E64[10 * e + 3 * i + j2] = (double)(-v);
                }
        }
    auto kern = p1_kernel<T, TD, C, NT>;
    int ni = P1_NI;  // items per lane of the kernel being launched (the rig-specialised build may use two)
    auto smem_of = [&](int gw) -> size_t { return (size_t)NW * (size_t)p1_warp_bytes(C, ni, gw, Pout) + (sizeof(TD) != sizeof(T) ? NP * 80 : 0); };
    // frames per warp tile for `occ` resident CTAs per SM: few wasted lanes in the last step and in the centre step,
    // not longer than the contiguous frame range a warp owns, and the shared memory of all resident CTAs must fit
    auto choose_gw = [&](int occ) -> int {
        const long long slots = (long long)h->sm_count * occ * NW;
        const long long range = ((long long)F + slots - 1) / slots;
        int Gw = 1;
        double best = -1.0;
        for (int gw = 32; gw >= 1; gw >>= 1) {
            if (gw > 1 && smem_of(gw) * occ > (size_t)h->smem_per_sm - 1024 * (size_t)occ) continue;
            if (gw > 1 && gw > range) continue;
            const long long items = (long long)gw * keypoint_num;
            const double eff = (double)items / (double)((items + 31) / 32 * 32);
            const double centre = (double)(gw * NP) / (double)((gw * NP + 31) / 32 * 32);  // lanes busy in the centre step
            const double score = eff * (0.9 + 0.1 * centre);
            if (score > best + 1e-9) { best = score; Gw = gw; }
        }
        if (h->tune_G > 0) {   // tests: a forced tile size, as far as one CTA's shared memory goes
            Gw = h->tune_G > 32 ? 32 : h->tune_G;
            while (Gw > 1 && smem_of(Gw) > (size_t)h->max_smem) Gw >>= 1;
        }
        return Gw;
    };
    auto grid_of = [&](int occ) -> int {
        int grid = h->sm_count * occ;               // one frame range per warp; a short batch uses fewer CTAs
        if ((long long)grid * NW > F) grid = (F + NW - 1) / NW;
        if (h->tune_ctas > 0 && grid > h->tune_ctas) grid = h->tune_ctas;
        return grid < 1 ? 1 : grid;
    };
    // Rig-specialised kernel (float modes): constants and batch shape baked in by NVRTC.  Soft failure.
    if (sizeof(T) == 4 && h->jit_mode > 0 && NT == 256) {
        // Same argument block (rig, thresholds, batch shape; the data pointers aside) and same knobs as the last
        // specialised launch: launch that kernel again without rebuilding and hashing its source.
        unsigned long long key = 1469598103934665603ull;
        {
            auto mix = [&](const void* ptr, size_t n) {
                const unsigned char* b = (const unsigned char*)ptr;
                for (size_t i = 0; i < n; ++i) key = (key ^ b[i]) * 1099511628211ull;
            };
            const size_t head = offsetof(P1ArgsT, F);   // the six data pointers come first
            mix((const unsigned char*)&a + head, sizeof(a) - head);
            const int knobs[6] = {h->jit_mode, h->tune_G, h->tune_ctas, (int)sizeof(TD), C, h->sm_count};
            mix(knobs, sizeof(knobs));
            for (const char* name : {"SNOWTRI_JIT_MINB", "SNOWTRI_JIT_DEFINES"})
                if (const char* e = getenv(name)) mix(e, strlen(e) + 1);
        }
        if (h->p1_fast_fn && h->p1_fast_key == key) {
            a.Gw = h->p1_fast_gw;
            if (snowtri_jit_launch(h->p1_fast_fn, h->p1_fast_grid, NT, h->p1_fast_smem, stream, &a) == 0) {
                h->launches += 1;
                h->last_grid = h->p1_fast_grid; h->last_block = NT; h->last_smem = (int)h->p1_fast_smem; h->last_G = h->p1_fast_gw;
                h->last_fly = 4;
                return SNOWTRI_OK;
            }
        }
        h->p1_fast_fn = nullptr;   // the cache below may evict (unload) the module the shortcut points at
        char buf[256];
        std::string src = "#define P1_JIT 1\n";
        // Items per lane and resident CTAs per SM of the specialised build.  With the constants out of the register
        // file and the next step's inputs staged through shared memory (no prefetch registers), two items per lane
        // side by side (twice the independent dependency chains, constants materialised once per pair) fit in 128
        // registers in both float modes.  Measured at BASELINE configs[1], ms per 131 072 frames (profiles/r2l, r2m):
        //   all-float32: 2 items x 2 CTAs 0.249 | 1 x 2: 0.274 | 1 x 3 (80 registers): 0.280 | 1 x 4: 0.304 | 2 x 3: 0.415 (spills)
        //   mixed:       2 x 2: 0.305 | 1 x 2: 0.315 | 1 x 3: 0.402 | 1 x 1: 0.394
        // More resident warps do not help: the issue rate saturates near 0.75-0.8 per scheduler either way.
        ni = C <= 4 ? P1_JIT_DEFAULT_NI : 1;
        int minb = C <= 4 ? P1_JIT_DEFAULT_MINB : p1_min_blocks<T, TD, C>();
        if (const char* e = getenv("SNOWTRI_JIT_MINB")) minb = atoi(e) >= 1 && atoi(e) <= 4 ? atoi(e) : minb;   // experiments
        if (const char* extra = getenv("SNOWTRI_JIT_DEFINES")) {  // experiments: "P1_NI=2 OTHER=1"
            std::string e(extra);
            size_t pos = 0;
            while (pos < e.size()) {
                size_t sp = e.find(' ', pos);
                if (sp == std::string::npos) sp = e.size();
                std::string tok = e.substr(pos, sp - pos);
                const size_t eq = tok.find('=');
                if (tok.compare(0, 6, "P1_NI=") == 0) ni = atoi(tok.c_str() + 6) == 2 ? 2 : 1;
                else if (!tok.empty()) src += "#define " + (eq == std::string::npos ? tok : tok.substr(0, eq) + " " + tok.substr(eq + 1)) + "\n";
                pos = sp + 1;
            }
        }
        snprintf(buf, sizeof(buf), "#define P1_NI %d\n", ni);
        src += buf;
        const int Gw = choose_gw(minb);
        const size_t smem = smem_of(Gw);
        a.Gw = Gw;
        auto def_i = [&](const char* n, int v) { snprintf(buf, sizeof(buf), "#define %s %d\n", n, v); src += buf; };
        auto def_f = [&](const char* n, float v) {
            if (isinf(v)) snprintf(buf, sizeof(buf), "#define %s __int_as_float(0x7f800000)\n", n);
            else snprintf(buf, sizeof(buf), "#define %s %.9ef\n", n, (double)v);
            src += buf;
        };
        def_i("P1_JIT_J", J); def_i("P1_JIT_JOUT", keypoint_num); def_i("P1_JIT_POUT", Pout); def_i("P1_JIT_GW", Gw);
        def_f("P1_JIT_KST", a.kst_f); def_f("P1_JIT_INV_DTHR", (float)a.inv_dthr); def_f("P1_JIT_GUARD_W", a.guard_w);
        def_f("P1_JIT_KSCALE_FULL", (float)a.kscale[NP]);
        auto def_arr = [&](const char* n, const T* v, int cnt) {
            src += std::string("#define ") + n + " {";
            for (int i = 0; i < cnt; ++i) { snprintf(buf, sizeof(buf), "%s%.9ef", i ? "," : "", (double)v[i]); src += buf; }
            src += "}\n";
        };
        def_arr("P1_JIT_CAMC", a.camc, C * 12);
        def_arr("P1_JIT_PDC", a.pdc, NP * 8);
        snprintf(buf, sizeof(buf),
                 "#include \"snowtri_p1.cuh\"\nextern \"C\" __global__ void __launch_bounds__(256, %d) p1_jit("
                 "const __grid_constant__ snowtri::P1Args<float, %d> a) { snowtri::p1_body<float, %s, %d, 256>(a); }\n",
                 minb, C, sizeof(TD) == 8 ? "double" : "float", C);
        src += buf;
        if (smem <= (size_t)h->max_smem && (h->jit_mode == 2 || F >= 65536 || snowtri_jit_cached(h, src))) {
            if (void* fn = snowtri_jit_get(h, src, "p1_jit", smem)) {
                const int grid = grid_of(minb);
                const int rc = snowtri_jit_launch(fn, grid, NT, smem, stream, &a);
                if (rc == 0) {
                    h->launches += 1;
                    h->last_grid = grid; h->last_block = NT; h->last_smem = (int)smem; h->last_G = Gw;
                    h->last_fly = 4;
                    h->p1_fast_key = key; h->p1_fast_fn = fn; h->p1_fast_grid = grid; h->p1_fast_gw = Gw; h->p1_fast_smem = smem;
                    return SNOWTRI_OK;
                }
                snprintf(h->jit_status, sizeof(h->jit_status), "failed: cuLaunchKernel, CUresult %d", rc);
            }
        }
    }
    // precompiled kernel
    ni = P1_NI;
    int occ = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(8) > 49152 ? (int)smem_of(8) : 49152) != cudaSuccess)
        cudaGetLastError();
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem_of(8));
    if (occ < 1) occ = 1;
    const int Gw = choose_gw(occ);
    a.Gw = Gw;
    const size_t smem = smem_of(Gw);
    if (smem > (size_t)h->max_smem)
        return fail(h, SNOWTRI_E_UNSUPPORTED, "snowtri_run: single-person path needs %zu B of shared memory (Pout=%d)", smem, Pout);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 49152 ? smem : 49152));
    if (e != cudaSuccess) return fail(h, SNOWTRI_E_CUDA, "p1_kernel attribute: %s", cudaGetErrorString(e));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
    if (occ < 1) occ = 1;
    const int grid = grid_of(occ);
    kern<<<grid, NT, smem, (cudaStream_t)stream>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, SNOWTRI_E_CUDA, "p1_kernel launch failed: %s", cudaGetErrorString(e));
    h->launches += 1;
    h->last_grid = grid; h->last_block = NT; h->last_smem = (int)smem; h->last_G = Gw;
    h->last_fly = 2;
    return SNOWTRI_OK;
}

template <typename T, typename TD>
static int run_p1(snowtri_t* h, const float* d_kpts, const float* d_scores, const int* d_counts, int F, int J,
                  int keypoint_num, int Pout, float* d_out, float* d_pscores, int* d_nout, void* stream) {
#define P1_CASE(CC) \
    case CC: return run_p1_c<T, TD, CC>(h, d_kpts, d_scores, d_counts, F, J, keypoint_num, Pout, d_out, d_pscores, d_nout, stream)
    switch (h->C) {
#ifdef P1_ONLY_C
        P1_CASE(P1_ONLY_C);
#else
        P1_CASE(2); P1_CASE(3); P1_CASE(4); P1_CASE(5); P1_CASE(6); P1_CASE(7); P1_CASE(8);
#endif
    }
#undef P1_CASE
    return fail(h, SNOWTRI_E_UNSUPPORTED, "single-person path supports 2..8 cameras");
}


bool snowtri_p1_eligible(const snowtri_t* h, int P, int Pout, int keypoint_num) {
    const bool all_kept = h->prm.ast <= 0.0 && h->prm.kst >= 0.0;
    const bool never_filter = h->prm.score_tol <= 0.0 && h->prm.kst >= 0.0;
    return !h->no_p1 && P == 1 && h->C >= 2 && h->C <= 8 && all_kept && never_filter && Pout <= 64 && keypoint_num >= 2 && keypoint_num <= kP1MaxJout;   // (ceil(2^32 / 1) does not fit the magic constant)
}

int snowtri_p1_run(snowtri_t* h, const float* d_kpts, const float* d_scores, const int* d_counts, int F, int J,
                   int keypoint_num, int Pout, float* d_out, float* d_pscores, int* d_nout, void* stream) {
    switch (h->precision) {
        case SNOWTRI_PREC_F32:
            return run_p1<float, float>(h, d_kpts, d_scores, d_counts, F, J, keypoint_num, Pout, d_out, d_pscores, d_nout, stream);
        case SNOWTRI_PREC_MIXED:
            return run_p1<float, double>(h, d_kpts, d_scores, d_counts, F, J, keypoint_num, Pout, d_out, d_pscores, d_nout, stream);
        default:
            return run_p1<double, double>(h, d_kpts, d_scores, d_counts, F, J, keypoint_num, Pout, d_out, d_pscores, d_nout, stream);
    }
}
