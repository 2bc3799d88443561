// Host side of the streaming general path (snowtri_general.cuh): scratch management, frame chunking, launches.
#include <math.h>
#include <string.h>

#include "snowtri_internal.h"
#include "snowtri_general.cuh"

using namespace snowtri;

static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

template <typename T, typename TD>
static int general_run(snowtri_t* h, const float* d_kpts, const float* d_scores, const int* d_counts, int F, int P,
                       int J, int keypoint_num, int Pout, float* d_out, float* d_pscores, int* d_nout, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int C = h->C;
    GenArgs a;
    memset(&a, 0, sizeof(a));
    a.cam = h->d_cam;
    a.C = C; a.P = P; a.J = J; a.Jout = keypoint_num; a.Pout = Pout;
    a.npairs = C * (C - 1) / 2;
    const long long ncand_ll = (long long)a.npairs * P * P;
    if (ncand_ll > 0x7fffffffLL / 4) return fail(h, SNOWTRI_E_UNSUPPORTED, "snowtri_run: %lld candidates per frame", ncand_ll);
    a.ncand = (int)ncand_ll;
    a.prm = h->prm;
    a.all_kept = (h->prm.ast <= 0.0 && h->prm.kst >= 0.0) ? 1 : 0;
    a.tol2 = h->prm.cond_tol >= 0.0 ? h->prm.cond_tol * h->prm.cond_tol : -1.0;

    // scratch per frame: keep, ab (1 B), cen (24 B), klist, memb, cstart, cn (4 B each) per candidate, + kcount
    const size_t nc = (size_t)(a.ncand > 0 ? a.ncand : 1);
    const size_t per_frame = nc * 50 + 4;
    const size_t budget = (size_t)384 << 20;
    long long fc_max = (long long)(budget / per_frame);
    if (fc_max < 1) fc_max = 1;
    if (fc_max > F) fc_max = F;
    if (h->tune_G > 0 && fc_max > h->tune_G) fc_max = h->tune_G;  // tests: force several chunks
    const size_t need = align16(fc_max * nc) * 2 + align16(fc_max * nc * 24) + align16(fc_max * nc * 4) * 4 +
                        align16(fc_max * nc * 8) +
                        align16(fc_max * 4) + 256;
    if (h->gen_scratch_bytes < need) {
        if (h->gen_scratch) cudaFree(h->gen_scratch);
        h->gen_scratch = nullptr;
        h->gen_scratch_bytes = 0;
        CUDA_TRY(h, cudaMalloc(&h->gen_scratch, need));
        h->gen_scratch_bytes = need;
    }
    unsigned char* p = (unsigned char*)h->gen_scratch;
    auto take = [&](size_t bytes) { unsigned char* r = p; p += align16(bytes); return r; };
    a.cen = (double*)take(fc_max * nc * 24);
    uint2* memb2 = (uint2*)take(fc_max * nc * 8);
    a.klist = (uint32_t*)take(fc_max * nc * 4);
    a.memb = (uint32_t*)take(fc_max * nc * 4);
    a.cstart = (int*)take(fc_max * nc * 4);
    a.cn = (int*)take(fc_max * nc * 4);
    a.kcount = (int*)take(fc_max * 4);
    a.keep = take(fc_max * nc);
    a.ab = take(fc_max * nc);

    const size_t tab = GenTables<T>::bytes(C, a.npairs);
    const int pb = P <= 4 ? 4 : 8;   // secondary persons scored side by side in the keep kernel
    const int nchunk = (keypoint_num + 31) / 32;
    const size_t R = (size_t)C * P * J;
    int last_grid = 0;
    for (int f0 = 0; f0 < F; f0 += (int)fc_max) {
        const int fc = F - f0 < fc_max ? F - f0 : (int)fc_max;
        a.F = fc;
        a.kpts = d_kpts + (size_t)f0 * R * 2;
        a.scores = d_scores + (size_t)f0 * R;
        a.counts = d_counts ? d_counts + (size_t)f0 * C : nullptr;
        a.out = d_out + (size_t)f0 * Pout * keypoint_num * 4;
        a.pscores = d_pscores + (size_t)f0 * Pout;
        a.nout = d_nout + f0;
        if (a.ncand > 0) {
            const long long items = (long long)fc * a.npairs * P;
            const long long blocks = (items + kGenWarps - 1) / kGenWarps;
            if (blocks > 0x7fffffffLL) return fail(h, SNOWTRI_E_UNSUPPORTED, "snowtri_run: batch too large");
            // frames whose raw (u,v,score) fit in shared memory are staged once per CTA (12 bytes per ray)
            const size_t staged = ((tab + 15) & ~(size_t)15) + R * 12;
            const bool use_smem = !h->no_fly && staged <= (size_t)h->max_smem;
            const bool small = staged <= (size_t)100 * 1024;   // two or more CTAs per SM
            cudaError_t e = cudaSuccess;
#define KEEP_LAUNCH(PB_)                                                                                          \
    do {                                                                                                          \
        if (!use_smem) {                                                                                          \
            gen_keep_kernel<T, PB_><<<(unsigned)blocks, kGenWarps * 32, tab, st>>>(a);                           \
        } else if (small) {                                                                                       \
            e = cudaFuncSetAttribute(gen_keep_smem_kernel<T, PB_, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)staged);                                                                \
            gen_keep_smem_kernel<T, PB_, 256><<<fc, 256, staged, st>>>(a);                                        \
        } else {                                                                                                  \
            e = cudaFuncSetAttribute(gen_keep_smem_kernel<T, PB_, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)staged);                                                                \
            gen_keep_smem_kernel<T, PB_, 512><<<fc, 512, staged, st>>>(a);                                        \
        }                                                                                                         \
    } while (0)
            if (pb == 4) KEEP_LAUNCH(4);
            else KEEP_LAUNCH(8);
#undef KEEP_LAUNCH
            if (e != cudaSuccess) return fail(h, SNOWTRI_E_CUDA, "gen_keep_smem_kernel attribute: %s", cudaGetErrorString(e));
            h->launches += 1;
            last_grid = (int)blocks;
        }
        const long long ncells = (long long)fc * a.ncand;
        if (ncells > 0) gen_centre_kernel<<<(unsigned)((ncells + 255) / 256), 256, 0, st>>>(a);
        if (a.ncand > 64) gen_cluster_block_kernel<<<fc, 256, 0, st>>>(a);
        else gen_cluster_warp_kernel<<<(fc + kGenWarps - 1) / kGenWarps, kGenWarps * 32, 0, st>>>(a);
        if (ncells > 0) gen_members_kernel<<<(unsigned)((ncells + 255) / 256), 256, 0, st>>>(a, memb2);
        const long long fitems = (long long)fc * Pout * nchunk;
        gen_fuse_kernel<T, TD><<<(unsigned)((fitems + kGenWarps - 1) / kGenWarps), kGenWarps * 32, tab, st>>>(a, memb2);
        gen_pscore_kernel<<<(unsigned)(((long long)fc * Pout + kGenWarps - 1) / kGenWarps), kGenWarps * 32, 0, st>>>(a);
        CUDA_TRY(h, cudaGetLastError());
        h->launches += 3 + (ncells > 0 ? 2 : 0);
    }
    h->last_grid = last_grid; h->last_block = kGenWarps * 32; h->last_smem = (int)tab; h->last_G = (int)fc_max;
    h->last_fly = 3;
    return SNOWTRI_OK;
}

bool snowtri_general_eligible(const snowtri_t* h) {
    const bool never_filter = h->prm.score_tol <= 0.0 && h->prm.kst >= 0.0;
    return never_filter && !h->no_p1 && h->C <= 255;
}

int snowtri_general_run(snowtri_t* h, const float* d_kpts, const float* d_scores, const int* d_counts, int F, int P,
                        int J, int keypoint_num, int Pout, float* d_out, float* d_pscores, int* d_nout, void* stream) {
    if (h->precision == SNOWTRI_PREC_F64)
        return general_run<double, double>(h, d_kpts, d_scores, d_counts, F, P, J, keypoint_num, Pout, d_out, d_pscores,
                                           d_nout, stream);
    return general_run<float, double>(h, d_kpts, d_scores, d_counts, F, P, J, keypoint_num, Pout, d_out, d_pscores,
                                      d_nout, stream);
}
