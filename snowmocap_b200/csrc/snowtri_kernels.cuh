// sm_100a kernels of the triangulation path.  See DESIGN.md for the data layout and the
// per-kernel rooflines.  Reference behaviour being reproduced:
//   rays        snowvision/camera.py:234-253
//   pair solve  snowvision/triangulation.py:24-31
//   candidates  snowvision/triangulation.py:50-93
//   condense    snowvision/triangulation.py:95-162   (quirks Q1-Q12 of SURVEY.md 8a)
#pragma once
#include "snowtri_fused.cuh"
#include "snowtri_math.cuh"

namespace snowtri {

constexpr int kThreads = 256;  // block size of the candidate / condense kernels
constexpr int kWarps = kThreads / 32;

// Single-warp greedy clustering with the exact sqrt-distance comparison of the reference, used by
// the float64 condense kernel (reference triangulation.py:107-134).
template <typename CenFn>
__device__ int cluster_warp_exact(int N, const uint32_t* klist, CenFn cen, unsigned char* ab, uint32_t* memb,
                                  int* cstart, int* cn, double tol, int num_tol, int lane) {
    for (int i = lane; i < N; i += 32) ab[i] = 0;
    __syncwarp();
    int K = 0, mpos = 0, mc = 0;
    const unsigned lt = (1u << lane) - 1u;
    while (mc < N - 1) {  // the last candidate is never a main (Q1/Q2)
        double mx, my, mz;
        cen(klist[mc], mx, my, mz);
        const int start = mpos;
        if (lane == 0) memb[mpos] = klist[mc];
        mpos += 1;
        int next = N;
        for (int base = (mc + 1) & ~31; base < N; base += 32) {
            const int i = base + lane;
            const bool live = (i > mc) && (i < N) && !ab[i];
            bool take = false;
            if (live) {
                double sx, sy, sz;
                cen(klist[i], sx, sy, sz);
                const double dx = mx - sx, dy = my - sy, dz = mz - sz;
                const double dist = sqrt(dx * dx + dy * dy + dz * dz);
                take = !(dist > tol);  // distance to the MAIN (Q3); NaN distance is absorbed
            }
            const unsigned bt = __ballot_sync(kFull, take);
            if (take) {
                memb[mpos + __popc(bt & lt)] = klist[i];
                ab[i] = 1;
            }
            mpos += __popc(bt);
            const unsigned bl = __ballot_sync(kFull, live && !take);
            if (bl != 0u && next == N) next = base + __ffs(bl) - 1;
        }
        const int n = mpos - start;
        if (n >= num_tol) {
            if (lane == 0) {
                cstart[K] = start;
                cn[K] = n;
            }
            ++K;
        } else {
            mpos = start;  // members stay absorbed (Q5)
        }
        mc = next;
        __syncwarp();
    }
    return K;
}

// ------------------------------------------------------------------------------------------
// Human_Triangulation alone: one warp per dense candidate, lanes over joints, inputs read
// straight from global memory (no size limit).  Always float64 (reference triangulation.py:50-93).
struct CandArgs {
    const float* kpts;
    const float* scores;
    const int* counts;
    const double* cam;
    double* cand;  // (F,Nc,J,4)
    double* avg;   // (F,Nc)
    int* keep;     // (F,Nc)
    int F, C, P, J, npairs, ncand;
    Params prm;
};

__global__ void __launch_bounds__(kThreads) candidates_kernel(const CandArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.y;
    const int n = blockIdx.x * kWarps + warp;
    if (n >= a.ncand) return;
    const int C = a.C, P = a.P, J = a.J;
    const int pair = n / (P * P), pm = (n / P) % P, ps = n % P;
    int mc, sc;
    decode_pair(pair, C, mc, sc);
    const int cm = a.counts ? a.counts[(size_t)f * C + mc] : P;
    const int cs = a.counts ? a.counts[(size_t)f * C + sc] : P;
    const size_t o = (size_t)f * a.ncand + n;
    if (!(pm < cm && ps < cs)) {
        if (lane == 0) {
            a.keep[o] = 0;
            a.avg[o] = 0.0;
        }
        return;
    }
    double Mm[9], Ms[9];
    V3<double> d, mid;
    {
        const double* qm = a.cam + mc * 12;
        const double* qs = a.cam + sc * 12;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            Mm[i] = qm[i];
            Ms[i] = qs[i];
        }
        d.x = qs[9] - qm[9]; d.y = qs[10] - qm[10]; d.z = qs[11] - qm[11];
        mid.x = (qm[9] + qs[9]) / 2; mid.y = (qm[10] + qs[10]) / 2; mid.z = (qm[11] + qs[11]) / 2;
    }
    const size_t rm = (((size_t)f * C + mc) * P + pm) * J, rs = (((size_t)f * C + sc) * P + ps) * J;
    const float2* uv = reinterpret_cast<const float2*>(a.kpts);
    double sum = 0.0;
    for (int j = lane; j < J; j += 32) {
        const float2 pmj = uv[rm + j], psj = uv[rs + j];
        const V3<double> hm = back_project<double>(Mm, (double)pmj.x, (double)pmj.y);
        const V3<double> hs = back_project<double>(Ms, (double)psj.x, (double)psj.y);
        const PairSol<double> s = pair_solve(hm, hs, d);
        const double gg = gated_g(s, a.scores[rm + j], a.scores[rs + j], a.prm.kst_f, a.prm.dthr);
        const double score = (gg + gg) * s.det;
        const V3<double> w = pair_midpoint(s, hm, hs, mid);
        double2* dst = reinterpret_cast<double2*>(a.cand + (o * J + j) * 4);
        dst[0] = make_double2(w.x, w.y);
        dst[1] = make_double2(w.z, score);
        sum += score;
    }
    sum = warp_sum(sum);
    if (lane == 0) {
        const double avg = sum / (double)J;
        a.avg[o] = avg;
        a.keep[o] = (avg < a.prm.ast) ? 0 : 1;
    }
}

// ------------------------------------------------------------------------------------------
// Human_Triangulation_Condense alone on materialised candidates (float64), one CTA per frame.
struct CondArgs {
    const double* cand;  // (F,N,J,4)
    const int* ncand;    // (F)
    double* out;         // (F,Pout,Jout,4)
    double* pscores;     // (F,Pout)
    int* nout;           // (F)
    int F, N, J, Jout, Pout;
    Params prm;
};

__global__ void __launch_bounds__(kThreads) condense_kernel(const CondArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.x, N = a.N, J = a.J, Jout = a.Jout, Pout = a.Pout;
    const int n_in = max(0, min(N, a.ncand[f]));
    // smem: klist u32[N], memb u32[N], cstart i32[N], cn i32[N], slot i32[N], ksum f64[N], ab u8[N], misc
    double* ksum = reinterpret_cast<double*>(smem);
    uint32_t* klist = reinterpret_cast<uint32_t*>(ksum + N);
    uint32_t* memb = klist + N;
    int* cstart = reinterpret_cast<int*>(memb + N);
    int* cn = cstart + N;
    int* slot = cn + N;
    int* misc = slot + N;  // [0] clusters, [1] emitted
    unsigned char* ab = reinterpret_cast<unsigned char*>(misc + 2);
    const double* cf = a.cand + (size_t)f * N * J * 4;

    for (int i = tid; i < n_in; i += kThreads) klist[i] = (uint32_t)i;
    __syncthreads();
    if (warp == 0) {
        const int ctr = a.prm.center;
        auto cen_fn = [cf, J, ctr](uint32_t n, double& x, double& y, double& z) {
            const double* q = cf + ((size_t)n * J + ctr) * 4;
            x = q[0];
            y = q[1];
            z = q[2];
        };
        const int K = cluster_warp_exact(n_in, klist, cen_fn, ab, memb, cstart, cn, a.prm.cond_tol, a.prm.num_tol, lane);
        if (lane == 0) misc[0] = K;
    }
    __syncthreads();
    const int K = misc[0];
    // reference triangulation.py:138-148 on stored candidates
    auto fuse_one = [&](int k, int j, double& X, double& Y, double& Z) -> double {
        const int n = cn[k];
        const uint32_t* mb = memb + cstart[k];
        double S = 0.0;
        for (int m = 0; m < n; ++m) S += cf[((size_t)mb[m] * J + j) * 4 + 3];
        X = Y = Z = 0.0;
        if (S == 0.0) return 0.0;
        for (int m = 0; m < n; ++m) {
            const double* q = cf + ((size_t)mb[m] * J + j) * 4;
            const double w = q[3] / S;
            X += q[0] * w;
            Y += q[1] * w;
            Z += q[2] * w;
        }
        return S / (double)n;
    };
    for (int k = warp; k < K; k += kWarps) {
        double s = 0.0;
        for (int j = lane; j < Jout; j += 32) {
            double X, Y, Z;
            s += fuse_one(k, j, X, Y, Z);
        }
        s = warp_sum(s);
        if (lane == 0) ksum[k] = s / (double)Jout;
    }
    __syncthreads();
    if (warp == 0) {
        const unsigned lt = (1u << lane) - 1u;
        int emitted = 0;
        for (int base = 0; base < K; base += 32) {
            const int k = base + lane;
            const bool pass = (k < K) && !(ksum[k] < a.prm.score_tol);
            const unsigned b = __ballot_sync(kFull, pass);
            if (k < K) slot[k] = pass ? emitted + __popc(b & lt) : -1;
            emitted += __popc(b);
        }
        if (lane == 0) {
            misc[1] = emitted;
            a.nout[f] = emitted;
        }
    }
    __syncthreads();
    for (int k = warp; k < K; k += kWarps) {
        const int s = slot[k];
        if (s < 0 || s >= Pout) continue;
        double* o = a.out + ((size_t)f * Pout + s) * Jout * 4;
        for (int j = lane; j < Jout; j += 32) {
            double X, Y, Z;
            const double ks = fuse_one(k, j, X, Y, Z);
            o[4 * j] = X;
            o[4 * j + 1] = Y;
            o[4 * j + 2] = Z;
            o[4 * j + 3] = ks;
        }
        if (lane == 0) a.pscores[(size_t)f * Pout + s] = ksum[k];
    }
    for (int s = misc[1] + warp; s < Pout; s += kWarps) {
        double* o = a.out + ((size_t)f * Pout + s) * Jout * 4;
        for (int j = lane; j < Jout * 4; j += 32) o[j] = 0.0;
        if (lane == 0) a.pscores[(size_t)f * Pout + s] = 0.0;
    }
}

// Batched Skew_Ray_Solver (reference triangulation.py:24-31), float64, one pair per thread.
__global__ void skew_ray_kernel(int n, const double* __restrict__ hm, const double* __restrict__ hs,
                                const double* __restrict__ tm, const double* __restrict__ ts,
                                double* __restrict__ dist, double* __restrict__ mid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V3<double> a{hm[3 * i], hm[3 * i + 1], hm[3 * i + 2]}, b{hs[3 * i], hs[3 * i + 1], hs[3 * i + 2]};
    V3<double> d{ts[3 * i] - tm[3 * i], ts[3 * i + 1] - tm[3 * i + 1], ts[3 * i + 2] - tm[3 * i + 2]};
    V3<double> c{(ts[3 * i] + tm[3 * i]) / 2, (ts[3 * i + 1] + tm[3 * i + 1]) / 2, (ts[3 * i + 2] + tm[3 * i + 2]) / 2};
    const PairSol<double> s = pair_solve(a, b, d);
    const V3<double> w = pair_midpoint(s, a, b, c);
    dist[i] = sqrt(s.qq) / s.det;
    mid[3 * i] = w.x;
    mid[3 * i + 1] = w.y;
    mid[3 * i + 2] = w.z;
}

}  // namespace snowtri
