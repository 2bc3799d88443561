// Host side of the streaming general path (snowtri_general.cuh): scratch management, frame chunking, launches.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "snowtri_internal.h"
#include "snowtri_general.cuh"
#include "snowtri_match.cuh"
#include "snowtri_mfuse.cuh"

using namespace snowtri;

static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// Second-generation fuse (snowtri_mfuse.cuh): argument block with the rig's constants, launch.
template <int C>
static int mfuse_launch(snowtri_t* h, const GenArgs& g, const uint2* memb2, const GenDesc* desc, cudaStream_t st) {
    constexpr int NP = C * (C - 1) / 2;
    constexpr int NT = C <= 6 ? 256 : 192, MINB = 2;   // registers: 128 per thread up to 6 cameras, 170 beyond
    if (h->mf_args_bytes < sizeof(MFArgs<C>)) {
        free(h->mf_args);
        h->mf_args = aligned_alloc(64, (sizeof(MFArgs<C>) + 63) & ~(size_t)63);
        h->mf_args_bytes = h->mf_args ? sizeof(MFArgs<C>) : 0;
        if (!h->mf_args) return fail(h, SNOWTRI_E_NOMEM, "snowtri_run: out of host memory");
    }
    MFArgs<C>& a = *reinterpret_cast<MFArgs<C>*>(h->mf_args);
    memset(&a, 0, sizeof(a));
    a.kpts = g.kpts; a.scores = g.scores; a.out = g.out; a.pscores = g.pscores; a.nout = g.nout;
    a.kcount = g.kcount; a.desc = desc; a.memb2 = memb2;
    a.F = g.F; a.P = g.P; a.J = g.J; a.Jout = g.Jout; a.Pout = g.Pout; a.ncand = g.ncand;
    // frames per tile: about five hundred (row, joint) items, at most 32 rows.  Measured at BASELINE configs[2] (4 of 8
    // rows filled, 133 joints): 1 frame per tile 0.754 ms, 2 frames 0.840 ms, 4 frames 0.952 ms (profiles/r2g) -- small
    // tiles balance better over the persistent warps than the few idle lanes of a tile's last step cost.
    {
        const int rows = g.Pout / 2 > 0 ? g.Pout / 2 : 1;
        int gw = (512 + rows * g.Jout - 1) / (rows * g.Jout);
        const int cap = 32 / g.Pout < 1 ? 1 : 32 / g.Pout;
        a.Gw = gw < 1 ? 1 : (gw > cap ? cap : gw);
    }
    a.tile_counter = g.tile_counter;
    a.kst_f = h->prm.kst_f;
    const double inv = h->prm.dthr > 0.0 ? 1.0 / h->prm.dthr : (double)INFINITY;
    a.inv_dthr = (float)inv;
    a.inv_dthr64 = inv;
    a.guard_w = isinf(inv) ? 0.f : (float)(inv * kGuardBandMixed);
    const double* cam = h->cam_host;
    for (int c = 0; c < C; ++c) {
        for (int k = 0; k < 9; ++k) {
            a.cam64[12 * c + (k / 3) * 4 + k % 3] = cam[12 * c + k];
            a.camc[12 * c + (k / 3) * 4 + k % 3] = (float)cam[12 * c + k];
        }
        for (int k = 0; k < 3; ++k) a.cam64t[4 * c + k] = cam[12 * c + 9 + k];
    }
    for (int x = 0; x < C - 1; ++x)
        for (int y = x + 1; y < C; ++y) {
            const int e = pair_index(C, x, y);
            long double d[3];
            for (int k = 0; k < 3; ++k) {
                const double tm = cam[12 * x + 9 + k], ts = cam[12 * y + 9 + k];
                d[k] = (long double)ts - (long double)tm;
                a.pdc[e * 8 + k] = (float)(ts - tm);
                a.pdc[e * 8 + 4 + k] = (float)((tm + ts) / 2);
            }
            // d.(hm x hs) = -hm^T [d]x hs  with hm = Mx [u v 1]^T, hs = My [u v 1]^T   =>   E = -Mx^T [d]x My
            const long double dx[9] = {0, -d[2], d[1], d[2], 0, -d[0], -d[1], d[0], 0};
            for (int i = 0; i < 3; ++i)
                for (int j2 = 0; j2 < 3; ++j2) {
                    long double v = 0;
                    for (int r = 0; r < 3; ++r)
                        for (int q = 0; q < 3; ++q) v += (long double)cam[12 * x + 3 * r + i] * dx[3 * r + q] * (long double)cam[12 * y + 3 * q + j2];
                    a.E[10 * e + 3 * i + j2] = (double)(-v);
                }
        }
    // pair loop fully unrolled (default) or with the first camera of a pair in a rolled loop (SNOWTRI_MF_ROLLED=1)
    const char* env = getenv("SNOWTRI_MF_ROLLED");
    const bool rolled = env ? atoi(env) != 0 : false;   // measured at 8 cameras: unrolled 0.82 ms, rolled 1.15 ms (profiles/r2b)
    auto kern = rolled ? mfuse_kernel<C, NT, MINB, true> : mfuse_kernel<C, NT, MINB, false>;
    constexpr int nt = NT;   // 128-thread CTAs, three per SM: 0.820 vs 0.840 ms at 8 cameras (profiles/r2g) -- not kept
    if (const char* e2 = getenv("SNOWTRI_MF_GW")) a.Gw = atoi(e2) > 0 && atoi(e2) * g.Pout <= 32 ? atoi(e2) : a.Gw;   // experiments
    const size_t smem = (size_t)(nt / 32) * mfuse_warp_bytes<C>();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 49152 ? smem : 49152));
    if (e != cudaSuccess) return fail(h, SNOWTRI_E_CUDA, "mfuse_kernel attribute: %s", cudaGetErrorString(e));
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, smem);
    if (occ < 1) occ = 1;
    int grid = h->sm_count * occ;  // persistent warps fetch tiles from a counter; a short batch uses fewer CTAs
    const long long tiles = ((long long)g.F + a.Gw - 1) / a.Gw;
    if ((long long)grid * (nt / 32) > tiles) grid = (int)((tiles + nt / 32 - 1) / (nt / 32));
    if (h->tune_ctas > 0 && grid > h->tune_ctas) grid = h->tune_ctas;
    kern<<<grid, nt, smem, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, SNOWTRI_E_CUDA, "mfuse_kernel launch failed: %s", cudaGetErrorString(e));
    return SNOWTRI_OK;
}

static int mfuse_run(snowtri_t* h, const GenArgs& g, const uint2* memb2, const GenDesc* desc, cudaStream_t st) {
    switch (h->C) {
        case 2: return mfuse_launch<2>(h, g, memb2, desc, st);
        case 3: return mfuse_launch<3>(h, g, memb2, desc, st);
        case 4: return mfuse_launch<4>(h, g, memb2, desc, st);
        case 5: return mfuse_launch<5>(h, g, memb2, desc, st);
        case 6: return mfuse_launch<6>(h, g, memb2, desc, st);
        case 7: return mfuse_launch<7>(h, g, memb2, desc, st);
        case 8: return mfuse_launch<8>(h, g, memb2, desc, st);
    }
    return fail(h, SNOWTRI_E_UNSUPPORTED, "mfuse: 2..8 cameras");
}

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

    // Second-generation kernels (snowtri_match.cuh, snowtri_mfuse.cuh) exist for the float32-bulk mode:
    //   match (keep + centre in one launch): any rig, needs a positive distance threshold;
    //   fuse  (members + fuse + person score in one launch): up to 8 cameras, up to 32 output slots.
    const bool gen2 = sizeof(T) == 4 && !h->gen1_only;
    const bool match2 = gen2 && h->prm.dthr > 0.0 && a.ncand > 0;
    const bool fuse2 = gen2 && C >= 2 && C <= 8 && Pout <= 32 && P <= 255;
    const size_t R = (size_t)C * P * J;
    const size_t mtab = MatchTables::bytes(C, a.npairs);
    const size_t mstaged = mtab + match_desc_bytes(a.npairs * ((P + kTile - 1) / kTile) * ((P + kTile - 1) / kTile)) + match_camf_bytes(C) +
                           (MATCH_JM ? match_jm_bytes(C, P, J) : R * 20);   // item descriptors, rays (16 B) + scores (4 B) of one frame
    const bool match_smem = match2 && !h->no_fly && mstaged <= ((size_t)h->smem_per_sm - 2048) / 2 - 1024;  // two CTAs per SM
    // a camera pair's 2 P rows staged per CTA (two CTAs per SM), when the whole frame does not fit
    const size_t mpair = mtab + 80 + (MATCH_JM ? match_jm_bytes(2, P, J) : (size_t)2 * P * J * 20);
    const char* env_pairk = getenv("SNOWTRI_MATCH_PAIR");   // experiments: 0 = rays from the scratch array
    const bool match_pair = match2 && !match_smem && !h->no_fly && mpair <= ((size_t)h->smem_per_sm - 2048) / 2 - 1024 &&
                            !(env_pairk && atoi(env_pairk) == 0);
    const bool match_glob = match2 && !match_smem && !match_pair;

    // scratch per frame: keep, ab (1 B), cen (24 B), klist, memb, cstart, cn (4 B each), memb2 (8 B) per candidate,
    // + kcount, + row descriptors (16 B per output slot), + rays (16 B each) for the large-rig match kernel
    const size_t nc = (size_t)(a.ncand > 0 ? a.ncand : 1);
    const size_t per_frame = nc * 50 + 4 + (fuse2 ? (size_t)Pout * 16 : 0) + (match_glob ? R * 16 : 0);
    // Scratch budget: a chunk should hold enough frames to fill the device with frame-per-CTA kernels.  At 384 MB,
    // BASELINE configs[4] (7.4 MB of scratch per frame) ran 50-frame chunks: the clustering kernel had 50 CTAs on 148 SMs
    // and took 48 % of the step (profiles/r2w).  A thirty-second of the device memory, between 384 MB and 4 GB.
    size_t budget = h->total_mem / 32;
    if (budget < ((size_t)384 << 20)) budget = (size_t)384 << 20;
    if (budget > ((size_t)4 << 30)) budget = (size_t)4 << 30;
    long long fc_max = (long long)(budget / per_frame);
    if (fc_max < 1) fc_max = 1;
    if (fc_max > F) fc_max = F;
    if (h->tune_G > 0 && fc_max > h->tune_G) fc_max = h->tune_G;  // tests: force several chunks
    const size_t need = align16(fc_max * nc) * 2 + align16(fc_max * nc * 24) + align16(fc_max * nc * 4) * 4 +
                        align16(fc_max * nc * 8) + align16(fc_max * 4) + align16((size_t)fc_max * Pout * 16) + 16 +
                        (match_glob ? align16((size_t)fc_max * R * 16) : 0) + 256;
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
    GenDesc* desc = (GenDesc*)take((size_t)fc_max * Pout * 16);
    float4* rays = match_glob ? (float4*)take((size_t)fc_max * R * 16) : nullptr;
    a.klist = (uint32_t*)take(fc_max * nc * 4);
    a.memb = (uint32_t*)take(fc_max * nc * 4);
    a.cstart = (int*)take(fc_max * nc * 4);
    a.cn = (int*)take(fc_max * nc * 4);
    a.kcount = (int*)take(fc_max * 4);
    a.tile_counter = fuse2 ? (int*)take(16) : nullptr;
    a.keep = take(fc_max * nc);
    a.ab = take(fc_max * nc);
    a.memb2 = fuse2 ? memb2 : nullptr;   // the clustering kernels then also decode the members and describe the rows
    a.desc = fuse2 ? desc : nullptr;

    const size_t tab = GenTables<T>::bytes(C, a.npairs);
    if (tab > (size_t)h->max_smem)
        return fail(h, SNOWTRI_E_UNSUPPORTED, "snowtri_run: the camera tables of %d cameras need %zu B of shared memory (max %d)",
                    C, tab, h->max_smem);
    const int pb = P <= 4 ? 4 : 8;   // secondary persons scored side by side in the first-generation keep kernel
    const int nchunk = (keypoint_num + 31) / 32;
    const int tpp = (P + kTile - 1) / kTile;
    // frames with at most this many candidates are clustered by one warp (no CTA barriers in the greedy chain), larger
    // ones by one CTA: BASELINE configs[2] (448 candidates) 0.126 -> 0.089 ms per 10 000 frames (profiles/r3d)
    int cluster_warp_max = 512;
    if (const char* e5 = getenv("SNOWTRI_CLUSTER_WARP_MAX")) cluster_warp_max = atoi(e5);   // experiments
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
        const long long ncells = (long long)fc * a.ncand;
        if (match2) {
            // ---- matching: keep decision + float64 centre of every candidate, one launch (+ the ray pre-pass of
            // frames that do not fit in shared memory)
            const int mitems = a.npairs * tpp * tpp;  // (camera pair, 4 x 4 person tile) per frame
            if (match_smem) {
                // eight warps per CTA, two CTAs per SM (128 registers).  Measured at BASELINE configs[2] (28 items per
                // frame): 8 warps 1.10 ms, 7 warps (divides the items evenly) 1.20 ms, 6 warps 1.22 ms, 4 warps 1.41 ms
                // (profiles/r2g) -- the ray build and the decisions scale with the warps, the idle half round does not hurt.
                int nw = mitems < 8 ? (mitems < 1 ? 1 : mitems) : 8;
                if (const char* e4 = getenv("SNOWTRI_MATCH_NW")) nw = atoi(e4) >= 1 && atoi(e4) <= 8 ? atoi(e4) : nw;   // experiments
                cudaError_t e = cudaFuncSetAttribute(gen_match_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)(mstaged > 49152 ? mstaged : 49152));
                if (e != cudaSuccess) return fail(h, SNOWTRI_E_CUDA, "gen_match_smem_kernel attribute: %s", cudaGetErrorString(e));
                gen_match_smem_kernel<<<fc, nw * 32, mstaged, st>>>(a);
                last_grid = fc;
                h->launches += 1;
            } else if (match_pair) {
                const int tiles = tpp * tpp;
                const int nw = tiles >= 8 ? 8 : (tiles < 4 ? 4 : tiles);   // at least four warps for the ray build
                const long long blocks = (long long)fc * a.npairs;
                if (blocks > 0x7fffffffLL) return fail(h, SNOWTRI_E_UNSUPPORTED, "snowtri_run: batch too large");
                cudaError_t e = cudaFuncSetAttribute(gen_match_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)(mpair > 49152 ? mpair : 49152));
                if (e != cudaSuccess) return fail(h, SNOWTRI_E_CUDA, "gen_match_pair_kernel attribute: %s", cudaGetErrorString(e));
                gen_match_pair_kernel<<<(unsigned)blocks, nw * 32, mpair, st>>>(a);
                last_grid = (int)blocks;
                h->launches += 1;
            } else {
                const size_t rtab = (size_t)C * 9 * 4;
                if (rtab > 49152 || mtab > 49152)
                    return fail(h, SNOWTRI_E_UNSUPPORTED, "snowtri_run: %d cameras exceed the shared-memory tables", C);
                gen_rays_kernel<<<h->sm_count * 8, 256, rtab, st>>>(a, rays);
                const long long items = (long long)fc * mitems;
                const long long blocks = (items + kGenWarps - 1) / kGenWarps;
                if (blocks > 0x7fffffffLL) return fail(h, SNOWTRI_E_UNSUPPORTED, "snowtri_run: batch too large");
                gen_match_global_kernel<<<(unsigned)blocks, kGenWarps * 32, mtab, st>>>(a, rays);
                last_grid = (int)blocks;
                h->launches += 2;
            }
        } else if (a.ncand > 0) {
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
            if (tab > 49152) e = cudaFuncSetAttribute(gen_keep_kernel<T, PB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab); \
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
            last_grid = (int)blocks;
            gen_centre_kernel<<<(unsigned)((ncells + 255) / 256), 256, 0, st>>>(a);
            h->launches += 2;
        }
        // ---- ordered compaction + greedy clustering (+ member decode and row descriptors for the new fuse)
        if (a.ncand > 16384) gen_cluster_block_kernel<1024><<<fc, 1024, 0, st>>>(a);   // (7 680 candidates: 256 threads 0.43 ms, 1 024 threads 0.68 ms per 2 000 frames; 126 976: 1 024 threads 2.4x faster)
        else if (a.ncand > cluster_warp_max) gen_cluster_block_kernel<256><<<fc, 256, 0, st>>>(a);
        else gen_cluster_warp_kernel<<<(fc + kGenWarps - 1) / kGenWarps, kGenWarps * 32, 0, st>>>(a);
        h->launches += 1;
        // ---- fuse + person score + persons per frame
        if (fuse2) {
            const int rc = mfuse_run(h, a, memb2, desc, st);
            if (rc) return rc;
            h->launches += 1;
        } else {
            if (ncells > 0) {
                gen_members_kernel<<<(unsigned)((ncells + 255) / 256), 256, 0, st>>>(a, memb2);
                h->launches += 1;
            }
            const long long fitems = (long long)fc * Pout * nchunk;
            if (tab > 49152) {
                cudaError_t e = cudaFuncSetAttribute(gen_fuse_kernel<T, TD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab);
                if (e != cudaSuccess) return fail(h, SNOWTRI_E_CUDA, "gen_fuse_kernel attribute: %s", cudaGetErrorString(e));
            }
            gen_fuse_kernel<T, TD><<<(unsigned)((fitems + kGenWarps - 1) / kGenWarps), kGenWarps * 32, tab, st>>>(a, memb2);
            gen_pscore_kernel<<<(unsigned)(((long long)fc * Pout + kGenWarps - 1) / kGenWarps), kGenWarps * 32, 0, st>>>(a);
            h->launches += 2;
        }
        CUDA_TRY(h, cudaGetLastError());
    }
    h->last_grid = last_grid; h->last_block = kGenWarps * 32; h->last_smem = (int)tab; h->last_G = (int)fc_max;
    h->last_fly = 3;
    h->last_gen2 = (match2 ? 1 : 0) | (fuse2 ? 2 : 0);
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
