// General fused path, streaming design: any number of cameras and persons per camera.
//
// fused_kernel (snowtri_fused.cuh) stages every ray of a frame in shared memory, which caps the problem at
// C*P*J*(12..24) bytes <= 227 KB (8 cameras x 4 persons x 133 joints in float64) -- BASELINE configs[3] and
// [4] do not fit.  This path keeps nothing per frame on chip: the (u,v,score) of a frame are 51 KB .. 817 KB and
// stay L2-resident while the frame is worked on, rays are rebuilt in registers where they are used (6 FMAs), and
// the candidate list lives in a scratch area of global memory that only ever holds one byte, one centre and one
// index per candidate -- candidates themselves (J x 4 values each) are never materialised.
//
//   K1 gen_keep_kernel     one warp per (frame, camera pair, main person): for every secondary person the
//                          mean gated score over the joints (distance-only form, keep_score) -> keep flag;
//                          float64 centre-joint midpoint of every kept candidate          (reference :56-87)
//   K2 gen_cluster_kernel  ordered compaction + greedy clustering per frame (one warp or one CTA per frame)
//                          -> member lists, cluster count                                  (reference :101-134)
//   K3 gen_fuse_kernel     one warp per (frame, output slot, 32 joints): re-solves the member pairs of the
//                          cluster and accumulates the score-weighted mean                 (reference :138-148)
//   K4 gen_pscore_kernel   person score = mean keypoint score, persons per frame           (reference :150-156)
//
// Precision: T = double, or T = float with the ray-distance numerator in double ("mixed", as in p1_kernel) --
// with several persons in view wrongly matched clusters fuse midpoints that lie metres apart, so the weights must
// be good to ~1e-7 for the joints to hold the 1e-4 bound; all-float32 is not offered here.  Discrete decisions
// (keep, clustering, distance gate near the threshold) are float64 in both modes.
// Requires never_filter (score_tol <= 0, kst >= 0): the output slot of a cluster is its index.
#pragma once
#include "snowtri_fused.cuh"
#include "snowtri_p1.cuh"

namespace snowtri {

struct GenDesc {               // one output row (frame, slot) of the second-generation fuse (snowtri_mfuse.cuh)
    unsigned long long obs;    // byte c: person index of camera c in this cluster, 0xff = camera not in the cluster
    int n;                     // members (reference :148 divides by it)
    int start;                 // offset of the member list in the frame's memb2 block; bit 31 set: clique
};

struct GenArgs {
    const float* kpts;    // (F,C,P,J,2)
    const float* scores;  // (F,C,P,J)
    const int* counts;    // (F,C) or null
    float* out;           // (F,Pout,Jout,4)
    float* pscores;       // (F,Pout)
    int* nout;            // (F)
    const double* cam;    // (C,12): M = R*inv(K) row-major (9), t (3)
    int F, C, P, J, Jout, Pout, npairs, ncand;
    int all_kept;
    Params prm;
    double tol2;
    // scratch, per frame of the chunk
    unsigned char* keep;  // (F,ncand)
    unsigned char* ab;    // (F,ncand)
    double* cen;          // (F,ncand,3)
    uint32_t* klist;      // (F,ncand)
    uint32_t* memb;       // (F,ncand)
    int* cstart;          // (F,ncand)
    int* cn;              // (F,ncand)
    int* kcount;          // (F)
    uint2* memb2;         // (F,ncand) decoded members, written by the clustering kernels when non-null
    GenDesc* desc;        // (F,Pout) row descriptors, written by the clustering kernels when non-null
    int* tile_counter;    // work counter of the fuse kernel that follows, zeroed by the clustering kernels when non-null
};

constexpr int kGenWarps = 8;  // warps per CTA of K1/K3/K4

// Shared-memory camera tables of K1/K3: M as T (C*9), M and t as double (C*12), pair table (npairs uchar2).
template <typename T>
struct GenTables {
    T* camM;
    const double* camD;  // (C,12) copy of a.cam
    uchar2* pairs;
    __device__ GenTables(unsigned char* smem, const GenArgs& a) {
        double* cd = reinterpret_cast<double*>(smem);
        camD = cd;
        camM = reinterpret_cast<T*>(cd + a.C * 12);
        pairs = reinterpret_cast<uchar2*>(camM + a.C * 9);
        for (int i = threadIdx.x; i < a.C * 12; i += blockDim.x) cd[i] = a.cam[i];
        for (int i = threadIdx.x; i < a.C * 9; i += blockDim.x) camM[i] = (T)a.cam[(i / 9) * 12 + (i % 9)];
        for (int p = threadIdx.x; p < a.npairs; p += blockDim.x) {
            int mc, sc;
            decode_pair(p, a.C, mc, sc);
            pairs[p] = make_uchar2((unsigned char)mc, (unsigned char)sc);
        }
        __syncthreads();
    }
    __host__ __device__ static size_t bytes(int C, int npairs) { return (size_t)C * 12 * 8 + (size_t)C * 9 * sizeof(T) + (size_t)npairs * 2 + 16; }
};

// Mean candidate score in float64 (guard path of the float32 keep decision), one warp, lanes over joints.
__device__ __noinline__ double gen_candidate_mean_f64(const GenArgs& a, const double* camD, const float2* kf,
                                                      const float* sf, int mc, int pm, int sc, int ps, int lane) {
    const int rm = (mc * a.P + pm) * a.J, rs = (sc * a.P + ps) * a.J;
    V3<double> d;
    d.x = camD[12 * sc + 9] - camD[12 * mc + 9];
    d.y = camD[12 * sc + 10] - camD[12 * mc + 10];
    d.z = camD[12 * sc + 11] - camD[12 * mc + 11];
    double sum = 0.0;
    for (int j = lane; j < a.J; j += 32) {
        const float2 qm = kf[rm + j], qs = kf[rs + j];
        const V3<double> hm = back_project<double>(camD + 12 * mc, (double)qm.x, (double)qm.y);
        const V3<double> hs = back_project<double>(camD + 12 * sc, (double)qs.x, (double)qs.y);
        const PairSol<double> s = pair_solve(hm, hs, d);
        const double gg = gated_g(s, sf[rm + j], sf[rs + j], a.prm.kst_f, a.prm.dthr);
        sum += (gg + gg) * s.det;
    }
    sum = warp_sum(sum);
    return sum / (double)a.J;
}

// ---- K1 ------------------------------------------------------------------------------------------------------
// One (frame, camera pair, main person) item by one warp.  kf/sf point at the frame's (u,v) and scores -- in
// global memory (gen_keep_kernel) or staged in shared memory (gen_keep_smem_kernel).
template <typename T, int PB>  // PB: secondary persons handled side by side (running sums in registers)
__device__ __forceinline__ void gen_keep_item(const GenArgs& a, const GenTables<T>& tb, const float2* kf, const float* sf,
                                              int f, int pair, int pm, int lane) {
    const int mc = tb.pairs[pair].x, sc = tb.pairs[pair].y;
    const int C = a.C, P = a.P, J = a.J;
    const int cm = a.counts ? max(0, min(P, a.counts[(size_t)f * C + mc])) : P;
    const int cs = a.counts ? max(0, min(P, a.counts[(size_t)f * C + sc])) : P;
    const size_t nbase = (size_t)f * a.ncand + (size_t)(pair * P + pm) * P;
    if (pm >= cm) {
        for (int ps = lane; ps < P; ps += 32) a.keep[nbase + ps] = 0;
        return;
    }
    const int rm = (mc * P + pm) * J;
    const T* Mm = tb.camM + 9 * mc;
    T Ms[9];  // secondary camera's matrix in registers: used once per joint of every secondary person
#pragma unroll
    for (int i = 0; i < 9; ++i) Ms[i] = tb.camM[9 * sc + i];
    V3<T> d;
    d.x = (T)(tb.camD[12 * sc + 9] - tb.camD[12 * mc + 9]);
    d.y = (T)(tb.camD[12 * sc + 10] - tb.camD[12 * mc + 10]);
    d.z = (T)(tb.camD[12 * sc + 11] - tb.camD[12 * mc + 11]);
    const T dthr2 = a.prm.dthr < 0.0 ? (T)-1 : (T)(a.prm.dthr * a.prm.dthr);
    const float kst_f = a.prm.kst_f;
    if (a.all_kept) {  // ast <= 0 with kst >= 0 can never reject: no sums needed
        for (int ps = lane; ps < P; ps += 32) a.keep[nbase + ps] = ps < cs ? 1 : 0;
        return;
    }
    for (int ps = cs + lane; ps < P; ps += 32) a.keep[nbase + ps] = 0;
    T Mmr[9];  // both camera matrices in registers
#pragma unroll
    for (int i = 0; i < 9; ++i) Mmr[i] = Mm[i];
    // Joints outer, secondary persons inner: the main ray of a joint is built once and meets PB secondary
    // persons, each with its own running sum (few registers, so several CTAs stay resident).
    for (int ps0 = 0; ps0 < cs; ps0 += PB) {
        T sum[PB];
        float err[PB];
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            sum[u] = (T)0;
            err[u] = 0.f;
        }
        for (int j0 = 0; j0 < J; j0 += 32) {  // every lane stays in the loop: keep_score_warp votes
            const bool valid = j0 + lane < J;
            const int j = valid ? j0 + lane : J - 1;
            const float2 q0 = kf[rm + j];
            const float s0 = sf[rm + j];
            const V3<T> h0 = back_project<T>(Mmr, (T)q0.x, (T)q0.y);
            const bool low0 = !valid || s0 < kst_f;
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                if (ps0 + u < cs) {
                    const int rs = (sc * P + ps0 + u) * J;
                    const float2 q = kf[rs + j];
                    const float ss = sf[rs + j];
                    const V3<T> hs = back_project<T>(Ms, (T)q.x, (T)q.y);
                    keep_score_warp(h0, hs, d, (T)s0 + (T)ss, low0 || ss < kst_f, dthr2, sum[u], err[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int ps = ps0 + u;
            if (ps < cs) {
                const T tot = warp_sum(sum[u]);
                // mean < ast  <=>  sum < ast*J, decided without the division unless the sum sits on the threshold
                const double thrJ = a.prm.ast * (double)J, gap = fabs((double)tot - thrJ);
                bool kept = !((double)tot < thrJ);  // NaN mean is kept (Q9)
                if constexpr (sizeof(T) == 4) {
                    // float32 sum closer to the threshold than its own error bound: redo in float64 (discrete decision)
                    const float e = warp_sum(err[u]);
                    const double slack = (double)kDistDelta * (double)e + 1e-5 * fabs((double)tot);
                    if (!(gap > slack)) kept = !(gen_candidate_mean_f64(a, tb.camD, kf, sf, mc, pm, sc, ps, lane) < a.prm.ast);
                } else {
                    if (!(gap > 1e-9 * fabs(thrJ))) kept = !((double)tot / (double)J < a.prm.ast);
                }
                if (lane == 0) a.keep[nbase + ps] = kept ? 1 : 0;
            }
        }
    }
}

// Any size: inputs read straight from global memory (L2-resident while the frame is worked on).
template <typename T, int PB>
__global__ void __launch_bounds__(kGenWarps * 32, (sizeof(T) == 8 ? 2 : 3)) gen_keep_kernel(const __grid_constant__ GenArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    GenTables<T> tb(smem, a);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long item = (long long)blockIdx.x * kGenWarps + warp, per_frame = (long long)a.npairs * a.P;
    if (item >= (long long)a.F * per_frame) return;
    const int f = (int)(item / per_frame), r = (int)(item - (long long)f * per_frame);
    const size_t R = (size_t)a.C * a.P * a.J;
    gen_keep_item<T, PB>(a, tb, reinterpret_cast<const float2*>(a.kpts) + (size_t)f * R, a.scores + (size_t)f * R, f,
                          r / a.P, r % a.P, lane);
}

// Frames whose raw (u,v,score) fit in shared memory (12 bytes per ray): one CTA per frame stages them once and
// its warps walk the (pair, main person) items out of shared memory -- each ray is read C-1 times P times.
template <typename T, int PB, int NT>
__global__ void __launch_bounds__(NT, (sizeof(T) == 8 ? 512 : 768) / NT) gen_keep_smem_kernel(const __grid_constant__ GenArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    GenTables<T> tb(smem, a);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x;
    const int R = a.C * a.P * a.J;
    float2* uv = reinterpret_cast<float2*>(smem + ((GenTables<T>::bytes(a.C, a.npairs) + 15) & ~(size_t)15));
    float* sv = reinterpret_cast<float*>(uv + R);
    const float2* kf = reinterpret_cast<const float2*>(a.kpts) + (size_t)f * R;
    const float* sf = a.scores + (size_t)f * R;
    for (int i = threadIdx.x; i < R; i += NT) {
        uv[i] = __ldg(kf + i);
        sv[i] = __ldg(sf + i);
    }
    __syncthreads();
    const int items = a.npairs * a.P;
    for (int it = warp; it < items; it += NT / 32) gen_keep_item<T, PB>(a, tb, uv, sv, f, it / a.P, it % a.P, lane);
}

// ---- K1b: centre-joint midpoint of every kept candidate, float64, one thread per candidate (reference :112,124)
__global__ void __launch_bounds__(256) gen_centre_kernel(const __grid_constant__ GenArgs a) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long long)a.F * a.ncand || !a.keep[n]) return;
    const int f = (int)(n / a.ncand), c = (int)(n - (long long)f * a.ncand);
    const int P = a.P, PP = P * P, J = a.J;
    const int pair = c / PP, pm = (c / P) % P, ps = c % P;
    int mc, sc;
    decode_pair(pair, a.C, mc, sc);
    const float2* kf = reinterpret_cast<const float2*>(a.kpts) + (size_t)f * a.C * P * J;
    const float2 qm = kf[(mc * P + pm) * J + a.prm.center], qs = kf[(sc * P + ps) * J + a.prm.center];
    const double* cm = a.cam + 12 * mc;
    const double* cs = a.cam + 12 * sc;
    const V3<double> h0 = back_project<double>(cm, (double)qm.x, (double)qm.y);
    const V3<double> h1 = back_project<double>(cs, (double)qs.x, (double)qs.y);
    V3<double> dd, mid;
    dd.x = cs[9] - cm[9]; dd.y = cs[10] - cm[10]; dd.z = cs[11] - cm[11];
    mid.x = (cm[9] + cs[9]) / 2; mid.y = (cm[10] + cs[10]) / 2; mid.z = (cm[11] + cs[11]) / 2;
    const PairSol<double> s = pair_solve(h0, h1, dd);
    const V3<double> w = pair_midpoint(s, h0, h1, mid);
    double* c3 = a.cen + n * 3;
    c3[0] = w.x;
    c3[1] = w.y;
    c3[2] = w.z;
}

// ---- tail of the clustering kernels (second generation): decode every cluster member once and recognise CLIQUE
// clusters -- one observation per camera and every pair of the observed cameras present, i.e. what a correctly matched
// person is.  Called by all `nthreads` threads of the frame's CTA (or the 32 lanes of its warp) after the frame's
// member lists are complete and visible.
//   memb2[i].x = row of the main ray      (mc*P + pm) | mc << 24
//   memb2[i].y = row of the secondary ray (sc*P + ps) | sc << 24        (rows < 2^24, cameras < 2^8)
__device__ __forceinline__ void gen_describe(const GenArgs& a, int f, int K, int tid, int nthreads) {
    const size_t o = (size_t)f * a.ncand;
    const int P = a.P, PP = P * P, C = a.C;
    const int total = K > 0 ? a.cstart[o + K - 1] + a.cn[o + K - 1] : 0;
    for (int i = tid; i < total; i += nthreads) {
        const int c = (int)a.memb[o + i];
        const int pair = c / PP, pm = (c / P) % P, ps = c % P;
        int mc, sc;
        decode_pair(pair, C, mc, sc);
        a.memb2[o + i] = make_uint2((uint32_t)(mc * P + pm) | ((uint32_t)mc << 24), (uint32_t)(sc * P + ps) | ((uint32_t)sc << 24));
    }
    if (!a.desc) return;
    const int lane = tid & 31, warp = tid >> 5, nw = nthreads >> 5;
    const int rows = min(K, a.Pout);
    for (int k = warp; k < rows; k += nw) {
        const int n = a.cn[o + k], st = a.cstart[o + k];
        bool clique = C <= 8 && n <= 28 && P <= 255;
        unsigned long long obs = ~0ull;
        if (clique) {  // every member sits in one lane
            int mc = -1, sc = -1, pm = 0, ps = 0;
            if (lane < n) {
                const int c = (int)a.memb[o + st + lane];
                pm = (c / P) % P;
                ps = c % P;
                decode_pair(c / PP, C, mc, sc);
            }
            bool bad = false;
            int ncam = 0;
            for (int c = 0; c < C; ++c) {
                const bool hit_m = mc == c, hit_s = sc == c;
                const int person = hit_m ? pm : ps;
                const unsigned hit = __ballot_sync(kFull, hit_m || hit_s);
                if (hit) {
                    const int ob = __shfl_sync(kFull, person, __ffs(hit) - 1);
                    bad |= __ballot_sync(kFull, (hit_m || hit_s) && person != ob) != 0u;
                    ++ncam;
                    obs = (obs & ~(0xffull << (8 * c))) | ((unsigned long long)(ob & 0xff) << (8 * c));
                }
            }
            clique = !bad && n == ncam * (ncam - 1) / 2;
        }
        if (lane == 0) {
            GenDesc d;
            d.obs = obs;
            d.n = n;
            d.start = st | (clique ? (int)0x80000000 : 0);
            a.desc[(size_t)f * a.Pout + k] = d;
        }
    }
}

// ---- K2 ------------------------------------------------------------------------------------------------------
// Large candidate counts: one CTA per frame.
template <int NT>  // 256 threads, or 1024 for frames with thousands of candidates: the greedy loop is a chain of sweeps over
                   // the remaining candidates, two CTA barriers per NT of them
__global__ void __launch_bounds__(NT) gen_cluster_block_kernel(const __grid_constant__ GenArgs a) {
    __shared__ int wtmp[2 * (NT / 32) + 4];
    const int f = blockIdx.x;
    const size_t o = (size_t)f * a.ncand;
    const int K = cluster_block<NT>(a.ncand, a.keep + o, a.klist + o, a.cen + 3 * o, a.ab + o, a.memb + o, a.cstart + o,
                                    a.cn + o, wtmp, a.tol2, a.prm.num_tol);
    if (threadIdx.x == 0) a.kcount[f] = K;
    if (a.tile_counter && blockIdx.x == 0 && threadIdx.x == 0) *a.tile_counter = 0;
    if (a.memb2) {  // cluster_block ends with a barrier: the lists are visible to the whole CTA
        __syncthreads();
        gen_describe(a, f, K, threadIdx.x, NT);
    }
}

// Small candidate counts: one warp per frame.
__global__ void __launch_bounds__(kGenWarps * 32) gen_cluster_warp_kernel(const __grid_constant__ GenArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x * kGenWarps + warp;
    if (f >= a.F) return;
    const size_t o = (size_t)f * a.ncand;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t* kl = a.klist + o;
    int nk = 0;
    // ordered compaction of the kept candidates, the keep bytes fetched 16 blocks at a time (one by one every block is
    // an L2 round trip)
    for (int base0 = 0; base0 < a.ncand; base0 += 16 * 32) {
        unsigned char kb[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int i = base0 + u * 32 + lane;
            kb[u] = i < a.ncand ? a.keep[o + i] : (unsigned char)0;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (base0 + u * 32 >= a.ncand) break;
            const bool k = kb[u] != 0;
            const unsigned b = __ballot_sync(kFull, k);
            if (k) kl[nk + __popc(b & lt)] = (uint32_t)(base0 + u * 32 + lane);
            nk += __popc(b);
        }
    }
    __syncwarp();
    const int K = nk <= 32 * kClusterRegs
                      ? cluster_warp_regs(nk, kl, a.cen + 3 * o, a.memb + o, a.cstart + o, a.cn + o, a.tol2, a.prm.num_tol, lane)
                      : cluster_warp(nk, kl, a.cen + 3 * o, a.ab + o, a.memb + o, a.cstart + o, a.cn + o, a.tol2,
                                     a.prm.num_tol, lane);
    if (lane == 0) a.kcount[f] = K;
    if (a.tile_counter && f == 0 && lane == 0) *a.tile_counter = 0;
    if (a.memb2) {
        __syncwarp();
        gen_describe(a, f, K, lane, 32);
    }
}

// ---- K2b: decode every cluster member once: dense candidate index -> (main row, secondary row, pair) -------
// memb2[i] = (main ray row start | pair << 24 ... ) does not fit 32 bits for large rigs, so two words are stored:
//   .x = row of the main ray      (mc*P + pm)   | mc << 24
//   .y = row of the secondary ray (sc*P + ps)   | sc << 24        (rows < 2^24, cameras < 2^8)
__global__ void __launch_bounds__(256) gen_members_kernel(const __grid_constant__ GenArgs a, uint2* memb2) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long long)a.F * a.ncand) return;
    const int f = (int)(n / a.ncand), i = (int)(n - (long long)f * a.ncand);
    const size_t o = (size_t)f * a.ncand;
    const int K = a.kcount[f];
    const int total = K > 0 ? a.cstart[o + K - 1] + a.cn[o + K - 1] : 0;
    if (i >= total) return;
    const int c = (int)a.memb[o + i];
    const int P = a.P, PP = P * P;
    const int pair = c / PP, pm = (c / P) % P, ps = c % P;
    int mc, sc;
    decode_pair(pair, a.C, mc, sc);
    memb2[o + i] = make_uint2((uint32_t)(mc * P + pm) | ((uint32_t)mc << 24), (uint32_t)(sc * P + ps) | ((uint32_t)sc << 24));
}

// ---- K3 ------------------------------------------------------------------------------------------------------
template <typename T, typename TD>
__global__ void __launch_bounds__(kGenWarps * 32) gen_fuse_kernel(const __grid_constant__ GenArgs a, const uint2* __restrict__ memb2) {
    constexpr bool MIXED = sizeof(TD) != sizeof(T);
    extern __shared__ __align__(16) unsigned char smem[];
    GenTables<T> tb(smem, a);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nchunk = (a.Jout + 31) / 32;
    const long long item = (long long)blockIdx.x * kGenWarps + warp;
    if (item >= (long long)a.F * a.Pout * nchunk) return;
    const int f = (int)(item / ((long long)a.Pout * nchunk));
    const int r = (int)(item - (long long)f * a.Pout * nchunk);
    const int k = r / nchunk, chunk = r - k * nchunk;
    const int j = chunk * 32 + lane;
    const bool active = j < a.Jout;
    const int jj = active ? j : a.Jout - 1;
    float4* orow = reinterpret_cast<float4*>(a.out) + ((size_t)f * a.Pout + k) * a.Jout;
    const size_t o = (size_t)f * a.ncand;
    if (k >= a.kcount[f]) {
        if (active) orow[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const int C = a.C, P = a.P, J = a.J, PP = P * P;
    const int n = a.cn[o + k];
    const uint2* mb = memb2 + o + a.cstart[o + k];
    const size_t R = (size_t)C * P * J;
    const float2* kf = reinterpret_cast<const float2*>(a.kpts) + (size_t)f * R;
    const float* sf = a.scores + (size_t)f * R;
    const float kst_f = a.prm.kst_f;
    const double inv64 = a.prm.dthr > 0.0 ? 1.0 / a.prm.dthr : (double)INFINITY;
    const T inv_dthr = (T)inv64;
    const T band = isinf(inv64) ? (T)0 : (T)(inv64 * kGuardBandMixed);

    T S = (T)0, X = (T)0, Y = (T)0, Z = (T)0;
    uint32_t prev = 0xffffffffu;
    V3<T> hm;
    V3<TD> hmD;
    V3<TD> tmD;  // centre of the main camera
    T Am = (T)0, smT = (T)0;
    bool lowm = false;
    for (int m = 0; m < n; ++m) {
        const uint2 mm = mb[m];
        const int mc = mm.x >> 24, sc = mm.y >> 24;
        if (mm.x != prev) {  // list order groups the candidates of one (pair, main person)
            const int rmain = (int)(mm.x & 0xffffffu) * J;
            const float2 q = kf[rmain + jj];
            const float s = sf[rmain + jj];
            hm = back_project<T>(tb.camM + 9 * mc, (T)q.x, (T)q.y);
            if constexpr (MIXED) hmD = back_project<TD>(tb.camD + 12 * mc, (TD)q.x, (TD)q.y);
            tmD.x = (TD)tb.camD[12 * mc + 9]; tmD.y = (TD)tb.camD[12 * mc + 10]; tmD.z = (TD)tb.camD[12 * mc + 11];
            Am = dot3(hm, hm);
            smT = (T)s;
            lowm = s < kst_f;
            prev = mm.x;
        }
        const int rs = (int)(mm.y & 0xffffffu) * J;
        const float2 q = kf[rs + jj];
        const float ss = sf[rs + jj];
        const V3<T> hs = back_project<T>(tb.camM + 9 * sc, (T)q.x, (T)q.y);
        const TD tsx = (TD)tb.camD[12 * sc + 9], tsy = (TD)tb.camD[12 * sc + 10], tsz = (TD)tb.camD[12 * sc + 11];
        V3<TD> dD;
        dD.x = tsx - tmD.x; dD.y = tsy - tmD.y; dD.z = tsz - tmD.z;
        V3<T> d, mid;
        d.x = (T)dD.x; d.y = (T)dD.y; d.z = (T)dD.z;
        mid.x = (T)((tsx + tmD.x) * (TD)0.5); mid.y = (T)((tsy + tmD.y) * (TD)0.5); mid.z = (T)((tsz + tmD.z) * (TD)0.5);
        const PairSolN<T> s = pair_solve_n(hm, Am, hs, dot3(hs, hs), d);
        T dn;
        V3<TD> hsD;
        if constexpr (MIXED) {
            hsD = back_project<TD>(tb.camD + 12 * sc, (TD)q.x, (TD)q.y);
            dn = (T)cross_dot(hmD, hsD, dD);
        } else {
            dn = cross_dot(hm, hs, d);
        }
        const T rr = rsqrt_fast(s.det * dn * dn);
        const T rd = rr * s.det;  // 1/dist
        bool far = rd < inv_dthr;  // dist > dthr (strict); NaN is not gated (Q8/Q9)
        if constexpr (MIXED) {
            if (fabsf(rd - inv_dthr) < band) {  // decide in float64
                const PairSol<double> s64 = pair_solve(hmD, hsD, dD);
                far = rsqrt_fast(s64.qq) * s64.det < inv64;
            }
        }
        if (far || lowm || ss < kst_f) continue;  // zero score: contributes nothing (its point is not needed)
        const T gq = (smT + (T)ss) * rr;
        const T w = gq * s.det;
        const V3<T> v = pair_v(s.n0, s.n1, hm, hs);
        S += w;
        X = fma(w, mid.x, fma((T)0.5 * gq, v.x, X));
        Y = fma(w, mid.y, fma((T)0.5 * gq, v.y, Y));
        Z = fma(w, mid.z, fma((T)0.5 * gq, v.z, Z));
    }
    float4 out4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (S != (T)0) {  // S == 0 leaves (0,0,0) with score 0 (Q7)
        const T rS = rcp_t(S);
        out4 = make_float4((float)(X * rS), (float)(Y * rS), (float)(Z * rS),
                           (float)(S * (T)0.0005 * rcp_t((T)n)));  // S/n with zero-score members counted (Q6)
    }
    if (active) orow[j] = out4;
}

// ---- K4 ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGenWarps * 32) gen_pscore_kernel(const __grid_constant__ GenArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * kGenWarps + warp;
    if (row >= (long long)a.F * a.Pout) return;
    const int f = (int)(row / a.Pout), k = (int)(row - (long long)f * a.Pout);
    const float4* orow = reinterpret_cast<const float4*>(a.out) + (size_t)row * a.Jout;
    double s = 0.0;
    for (int j = lane; j < a.Jout; j += 32) s += (double)orow[j].w;
    s = warp_sum(s);
    if (lane == 0) {
        a.pscores[row] = (float)(s / (double)a.Jout);
        if (k == 0) a.nout[f] = a.kcount[f];
    }
}

}  // namespace snowtri
