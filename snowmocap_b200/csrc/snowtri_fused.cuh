// Fused sm_100a kernel: rays -> candidates -> gating -> clustering -> score-weighted fuse for a
// batch of frames (reference main.py:55-71 per frame; camera.py:234-253, triangulation.py:24-162).
// See DESIGN.md ("fused kernel") for the shared-memory layout and the phase structure.
#pragma once
#include "snowtri_internal.h"
#include "snowtri_math.cuh"

namespace snowtri {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kCliqueMax = 8;  // the register-resident fuse path handles up to 8 cameras

// Byte offsets of the shared-memory regions of the fused kernel (computed on the host).
struct FusedSmem {
    int cam, pairs, pd, stage_uv, stage_s, hx, hy, hz, sc, cnt, cen, keep, ab, klist, memb, membp, cstart, cn,
        ksum, slot, kcount, ks, cobs, clq, wtmp, total;
    int stage_stride_uv, stage_stride_s;  // bytes between the two staging buffers (fly mode), else 0
};

template <typename T>
struct FusedArgs {
    const float* kpts;    // (F,C,P,J,2)
    const float* scores;  // (F,C,P,J)
    const int* counts;    // (F,C) or null
    float* out;           // (F,Pout,Jout,4)
    float* pscores;       // (F,Pout)
    int* nout;            // (F)
    const double* cam;    // (C,12): M = R*inv(K) row-major (9), t (3)
    int F, C, P, J, Jout, Pout;
    int npairs, ncand;    // ncand = npairs*P*P dense candidates per frame
    int G;                // frames per group
    int R;                // rays per frame = C*P*J
    int use_tma, all_kept, never_filter;
    FusedSmem sm;
    Params prm;
    T inv_dthr;           // 1/dthr (+inf when dthr <= 0): dist > dthr <=> det*rsqrt(q.q) < inv_dthr
    double tol2;          // cond_tol^2 (-1 when cond_tol < 0): dist > tol <=> dist^2 > tol2
    // constant-bank tables of the <= 8-camera paths: pair constants (d = ts - tm, mid = (tm+ts)/2)
    // in the triangular order of CM cameras, entry (a*CM - a*(a+1)/2 + b - a - 1)*6; M = R*inv(K)
    T pdc[kCliqueMax * (kCliqueMax - 1) / 2 * 6];
    T camc[kCliqueMax * 9];
};

__device__ __forceinline__ void decode_pair(int p, int C, int& mc, int& sc) {
    int m = 0, rem = p;
    while (rem >= C - 1 - m) {
        rem -= C - 1 - m;
        ++m;
    }
    mc = m;
    sc = m + 1 + rem;
}

// dist > tol on squared distances; NaN is never "greater" (absorbed, like the reference).
__device__ __forceinline__ bool centre_far(double dx, double dy, double dz, double tol2) {
    return (dx * dx + dy * dy + dz * dz) > tol2;
}

// Greedy clustering of one frame by ONE WARP (reference triangulation.py:107-134); used when the
// candidate count is small and several frames share a CTA.
//   N      kept candidates, klist[i] = dense index of the i-th kept candidate (reference list order)
//   cen    centre-joint midpoints (3 doubles per dense candidate)
//   out    memb (dense indices grouped by cluster, in list order), cstart/cn per emitted cluster
// Returns the number of clusters that pass num_tol.  `ab` is N bytes of scratch.
__device__ __forceinline__ int cluster_warp(int N, const uint32_t* klist, const double* cen, unsigned char* ab,
                                            uint32_t* memb, int* cstart, int* cn, double tol2, int num_tol,
                                            int lane) {
    for (int i = lane; i < N; i += 32) ab[i] = 0;
    __syncwarp();
    int K = 0, mpos = 0, mc = 0;
    const unsigned lt = (1u << lane) - 1u;
    while (mc < N - 1) {  // the last candidate is never a main (Q1/Q2)
        const double* cm = cen + 3 * klist[mc];
        const double mx = cm[0], my = cm[1], mz = cm[2];
        const int start = mpos;
        if (lane == 0) memb[mpos] = klist[mc];
        mpos += 1;
        int next = N;
        for (int base = (mc + 1) & ~31; base < N; base += 32) {
            const int i = base + lane;
            const bool live = (i > mc) && (i < N) && !ab[i];
            bool take = false;
            if (live) {
                const double* cs = cen + 3 * klist[i];
                take = !centre_far(mx - cs[0], my - cs[1], mz - cs[2], tol2);  // distance to the MAIN (Q3)
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

// Same greedy clustering by the WHOLE CTA (large candidate counts, one frame at a time).
// wtmp: 2*NW+4 ints of scratch.  Must be called by all NT threads; returns K to every thread.
// The per-warp counts of a CTA pass (wcnt[NW], NW <= 32): `off` = sum over the warps before `warp`, `tot` = sum over all.
// Every warp scans the NW counts with its own lanes (5 shuffle steps) instead of every thread walking the array.
// The same for up to 32 * kClusterRegs kept candidates, from registers: lane l holds candidates l, l + 32, ... (index and
// centre), the main's centre travels by shuffles, the absorbed flags are a bit mask per lane.  cluster_warp reads
// klist -> centre from global memory for every main and every block of candidates -- three dependent L2 round trips per
// block, ~80 us per frame of BASELINE configs[2] (140 kept candidates, 8 mains); here the frame's centres are read once.
// Same sweeps, same order, same comparisons (centre_far on the same differences): identical clusters.
constexpr int kClusterRegs = 8;
__device__ __forceinline__ int cluster_warp_regs(int N, const uint32_t* klist, const double* cen, uint32_t* memb, int* cstart,
                                                 int* cn, double tol2, int num_tol, int lane) {
    double cx[kClusterRegs], cy[kClusterRegs], cz[kClusterRegs];
    uint32_t id[kClusterRegs];
#pragma unroll
    for (int r = 0; r < kClusterRegs; ++r) {
        const int i = r * 32 + lane;
        id[r] = i < N ? klist[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < kClusterRegs; ++r) {
        const double* c = cen + 3 * (size_t)id[r];
        const bool in = r * 32 + lane < N;
        cx[r] = in ? c[0] : 0.0;
        cy[r] = in ? c[1] : 0.0;
        cz[r] = in ? c[2] : 0.0;
    }
    unsigned absorbed = 0u;  // bit r: candidate r * 32 + lane
    int K = 0, mpos = 0, mc = 0;
    const unsigned lt = (1u << lane) - 1u;
    while (mc < N - 1) {  // the last candidate is never a main (Q1/Q2)
        const int rm = mc >> 5;
        double mx = cx[0], my = cy[0], mz = cz[0];
        uint32_t mid = id[0];
#pragma unroll
        for (int r = 1; r < kClusterRegs; ++r)
            if (r == rm) {
                mx = cx[r]; my = cy[r]; mz = cz[r];
                mid = id[r];
            }
        mx = __shfl_sync(kFull, mx, mc & 31);
        my = __shfl_sync(kFull, my, mc & 31);
        mz = __shfl_sync(kFull, mz, mc & 31);
        mid = __shfl_sync(kFull, mid, mc & 31);
        const int start = mpos;
        if (lane == 0) memb[mpos] = mid;
        mpos += 1;
        int next = N;
#pragma unroll
        for (int r = 0; r < kClusterRegs; ++r) {
            const int base = r * 32;
            if (base + 31 <= mc || base >= N) continue;  // warp-uniform: blocks before the main / beyond the list
            const int i = base + lane;
            const bool live = (i > mc) && (i < N) && !((absorbed >> r) & 1u);
            const bool take = live && !centre_far(mx - cx[r], my - cy[r], mz - cz[r], tol2);  // distance to the MAIN (Q3)
            const unsigned bt = __ballot_sync(kFull, take);
            if (take) {
                memb[mpos + __popc(bt & lt)] = id[r];
                absorbed |= 1u << r;
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
    }
    __syncwarp();
    return K;
}

template <int NW>
__device__ __forceinline__ void cluster_warp_prefix(const int* wcnt, int warp, int lane, int& off, int& tot) {
    static_assert(NW <= 32, "one lane per warp of the CTA");
    const int c = lane < NW ? wcnt[lane] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += up;
    }
    off = __shfl_sync(kFull, incl - c, warp);
    tot = __shfl_sync(kFull, incl, 31);
}

template <int NT>
__device__ int cluster_block(int ncand, const unsigned char* keep, uint32_t* klist, const double* cen,
                             unsigned char* ab, uint32_t* memb, int* cstart, int* cn, int* wtmp, double tol2,
                             int num_tol) {
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    int* wcnt = wtmp;            // [NW] per-warp counts
    int* wmin = wtmp + NW;       // [NW] per-warp first rejected index
    int* state = wtmp + 2 * NW;  // [0] main, [1] mpos, [2] K, [3] nk
    // ---- ordered compaction of the keep flags -> klist ----
    int base = 0;
    for (int r0 = 0; r0 < ncand; r0 += NT) {
        const int i = r0 + tid;
        const bool k = (i < ncand) && keep[i];
        const unsigned b = __ballot_sync(kFull, k);
        if (lane == 0) wcnt[warp] = __popc(b);
        __syncthreads();
        int off, tot;
        cluster_warp_prefix<NW>(wcnt, warp, lane, off, tot);
        off += base;
        if (k) {
            klist[off + __popc(b & lt)] = (uint32_t)i;
            ab[off + __popc(b & lt)] = 0;
        }
        base += tot;
        __syncthreads();
    }
    const int N = base;
    if (tid == 0) {
        state[0] = 0;
        state[1] = 0;
        state[2] = 0;
    }
    __syncthreads();
    while (true) {
        const int mc = state[0], mpos0 = state[1];
        if (mc >= N - 1) break;
        const double* cm = cen + 3 * klist[mc];
        const double mx = cm[0], my = cm[1], mz = cm[2];
        int taken = 0, next = N;
        for (int r0 = mc + 1; r0 < N; r0 += NT) {
            const int i = r0 + tid;
            const bool live = (i < N) && !ab[i];
            bool take = false;
            if (live) {
                const double* cs = cen + 3 * klist[i];
                take = !centre_far(mx - cs[0], my - cs[1], mz - cs[2], tol2);
            }
            const unsigned bt = __ballot_sync(kFull, take);
            const unsigned bl = __ballot_sync(kFull, live && !take);
            if (lane == 0) {
                wcnt[warp] = __popc(bt);
                wmin[warp] = bl ? (r0 + warp * 32 + __ffs(bl) - 1) : N;
            }
            __syncthreads();
            int off, tot;
            cluster_warp_prefix<NW>(wcnt, warp, lane, off, tot);
            int mn = lane < NW ? wmin[lane] : N;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(kFull, mn, o));
            if (take) {
                memb[mpos0 + 1 + taken + off + __popc(bt & lt)] = klist[i];
                ab[i] = 1;
            }
            taken += tot;
            next = min(next, mn);
            __syncthreads();
        }
        if (tid == 0) {
            memb[mpos0] = klist[mc];
            const int n = 1 + taken;
            if (n >= num_tol) {
                cstart[state[2]] = mpos0;
                cn[state[2]] = n;
                state[2] += 1;
                state[1] = mpos0 + n;
            }  // else: members stay absorbed, list position is reused (Q5)
            state[0] = next;
        }
        __syncthreads();
    }
    const int K = state[2];
    __syncthreads();
    return K;
}

// g-weights of one solved pair: returns g = (sm+ss)*0.00025*rsqrt(q.q) (0 when gated) and the
// reference score w = 2*g*det through `w`.  `sa` = sm+ss in T, `low` = (sm<kst || ss<kst).
template <typename T>
__device__ __forceinline__ T pair_weights(const PairSol<T>& s, T sa, bool low, T inv_dthr, T& w) {
    const T r = rsqrt_fast(s.qq);
    const T rd = r * s.det;            // = 1/dist
    const T c1 = sa * (T)0.00025;
    T g = c1 * r;
    w = (c1 + c1) * rd;
    if (low || rd < inv_dthr) {        // dist > dthr (strict); NaN distance is not gated, like the reference
        g = (T)0;
        w = (T)0;
    }
    return g;
}

// Keep-phase variant for float32: the score w = c/dist has absolute error ~ w * (delta/dist) where
// delta ~ 1e-5 m bounds the float32 error of the ray distance itself, so `err` accumulates w/dist and
// the caller compares |sum - threshold| against delta*err.  A distance inside the guard band of dthr
// (the gate could flip) poisons the bound with +inf.
__device__ __forceinline__ float pair_weight_err(const PairSol<float>& s, float sa, bool low, float inv_dthr,
                                                 float guard_w, float& err) {
    const float r = rsqrt_fast(s.qq);
    const float rd = r * s.det;  // = 1/dist
    float w = sa * 0.0005f * rd;
    if (!low && fabsf(rd - inv_dthr) < guard_w) err = INFINITY;
    if (low || rd < inv_dthr) w = 0.f;
    err = fmaf(w, rd, err);
    return w;
}
constexpr float kDistDelta = 1e-5f;   // metres; bound on the float32 error of a ray-to-ray distance
constexpr float kGateGuard = 4e-3f;   // relative half-width of the guard band around 1/dthr

// Keep phase (reference triangulation.py:70-79 needs only the SCORE of every joint of every candidate, not the
// midpoint): with n = hm x hs,  dist = |d.n| / |n|, so
//     dist > dthr  <=>  (d.n)^2 > dthr^2 * n.n            (no reciprocal, no square root)
//     score        =   (sm+ss) * 0.0005 * sqrt(n.n) * rsqrt((d.n)^2)    only for the joints that pass the gate
// 14 FMA-class operations for a gated joint instead of the ~43 of the full pair solve.  Wrongly matched
// candidates -- the large majority when several persons are in view -- are gated at nearly every joint.
// `dthr2` = dthr^2, or -1 when dthr < 0 (every finite distance is then "greater").  NaN is not gated (Q8/Q9).
// float32: `err` accumulates score/dist (see kDistDelta); a gate within its guard band adds the whole score it
// could contribute.
template <typename T>
__device__ __forceinline__ T keep_score(const V3<T>& hm, const V3<T>& hs, const V3<T>& d, T sa, bool low, T dthr2,
                                        float& err) {
    V3<T> n;
    n.x = fma(hm.y, hs.z, -(hm.z * hs.y));
    n.y = fma(hm.z, hs.x, -(hm.x * hs.z));
    n.z = fma(hm.x, hs.y, -(hm.y * hs.x));
    const T nn = dot3(n, n), dn = dot3(n, d), dn2 = dn * dn, lim = dthr2 * nn;
    if (low) return (T)0;
    bool gated = dn2 > lim;
    if constexpr (sizeof(T) == 4) {
        // gate inside its guard band: either outcome is possible, so the score it could contribute goes into the
        // error bound (in units of kDistDelta) and the joint is scored as the float32 comparison says
        if (fabsf(dn2 - lim) < (2.f * kGateGuard) * lim) {
            const float rdu = nn * rsqrt_fast(nn) * rsqrt_fast(dn2);
            err += sa * 0.0005f * rdu * (1.0f / kDistDelta);
        }
    }
    if (gated) return (T)0;
    const T rd = nn * rsqrt_fast(nn) * rsqrt_fast(dn2);  // sqrt(n.n)/|d.n| = 1/dist
    const T w = sa * (T)0.0005 * rd;
    if constexpr (sizeof(T) == 4) err = fmaf(w, rd, err);
    return w;
}

// Warp-uniform variant (all 32 lanes must call it; lanes without a joint pass low = true): the reciprocal
// square roots are behind a vote, so a warp whose 32 joints are all gated -- the normal case for a wrongly
// matched candidate -- never issues them.
template <typename T>
__device__ __forceinline__ void keep_score_warp(const V3<T>& hm, const V3<T>& hs, const V3<T>& d, T sa, bool low,
                                                T dthr2, T& sum, float& err) {
    V3<T> n;
    n.x = fma(hm.y, hs.z, -(hm.z * hs.y));
    n.y = fma(hm.z, hs.x, -(hm.x * hs.z));
    n.z = fma(hm.x, hs.y, -(hm.y * hs.x));
    const T nn = dot3(n, n), dn = dot3(n, d), dn2 = dn * dn, lim = dthr2 * nn;
    const bool pass = !low && !(dn2 > lim);  // NaN is not gated (Q8/Q9)
    bool near = false;
    if constexpr (sizeof(T) == 4) near = !low && fabsf(dn2 - lim) < (2.f * kGateGuard) * lim;
    if (__any_sync(kFull, pass || near)) {
        const T rd = nn * rsqrt_fast(nn) * rsqrt_fast(dn2);  // sqrt(n.n)/|d.n| = 1/dist
        const T w = sa * (T)0.0005 * rd;
        if (pass) sum += w;
        if constexpr (sizeof(T) == 4) {
            // error bound in units of kDistDelta: score/dist, plus the whole score of a joint whose gate could flip
            if (pass) err = fmaf((float)w, (float)rd, err);
            if (near) err += (float)w * (1.0f / kDistDelta);
        }
    }
}

// Mean candidate score in float64 from the raw inputs in global memory, one warp, lanes over joints
// (reference triangulation.py:70-79).  Cold path of the float32 kernel's keep decision.
template <typename T>
__device__ __noinline__ double candidate_mean_f64(const FusedArgs<T>& a, int f, int mc, int pm, int sc, int ps, int lane) {
    const float2* kg = reinterpret_cast<const float2*>(a.kpts) + (size_t)f * a.R;
    const float* sg = a.scores + (size_t)f * a.R;
    const int rm = (mc * a.P + pm) * a.J, rs = (sc * a.P + ps) * a.J;
    V3<double> d;
    d.x = a.cam[12 * sc + 9] - a.cam[12 * mc + 9];
    d.y = a.cam[12 * sc + 10] - a.cam[12 * mc + 10];
    d.z = a.cam[12 * sc + 11] - a.cam[12 * mc + 11];
    double sum = 0.0;
    for (int j = lane; j < a.J; j += 32) {
        const float2 qm = kg[rm + j], qs = kg[rs + j];
        const V3<double> hm = back_project<double>(a.cam + 12 * mc, (double)qm.x, (double)qm.y);
        const V3<double> hs = back_project<double>(a.cam + 12 * sc, (double)qs.x, (double)qs.y);
        const PairSol<double> s = pair_solve(hm, hs, d);
        const double gg = gated_g(s, sg[rm + j], sg[rs + j], a.prm.kst_f, a.prm.dthr);
        sum += (gg + gg) * s.det;
    }
    sum = warp_sum(sum);
    return sum / (double)a.J;
}

// ------------------------------------------------------------------------------------------
// Template: T compute type, NT threads, CM cameras of the unrolled clique path (0 = off),
//           NCH joint chunks held in registers by phase 1a (0 = generic loop, any J),
//           FLY: rays are recomputed from the staged (u,v) instead of being stored (P == 1, C <= CM).
template <typename T, int NT, int CM, int NCH, bool FLY>
__global__ void __launch_bounds__(NT, 512 / NT) fused_kernel(const __grid_constant__ FusedArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = a.C, P = a.P, J = a.J, Jout = a.Jout, Pout = a.Pout, R = a.R, G = a.G;
    const int ncand = a.ncand, PP = P * P;

    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);  // two mbarriers
    T* camM = reinterpret_cast<T*>(smem + a.sm.cam);
    uchar2* pairs = reinterpret_cast<uchar2*>(smem + a.sm.pairs);
    T* pd = reinterpret_cast<T*>(smem + a.sm.pd);  // per pair: d (3), mid (3)
    T* hx = reinterpret_cast<T*>(smem + a.sm.hx);
    T* hy = reinterpret_cast<T*>(smem + a.sm.hy);
    T* hz = reinterpret_cast<T*>(smem + a.sm.hz);
    float* sc_ = reinterpret_cast<float*>(smem + a.sm.sc);
    int* cnt = reinterpret_cast<int*>(smem + a.sm.cnt);
    double* cen = reinterpret_cast<double*>(smem + a.sm.cen);
    unsigned char* keep = smem + a.sm.keep;
    unsigned char* ab = smem + a.sm.ab;
    uint32_t* klist = reinterpret_cast<uint32_t*>(smem + a.sm.klist);
    uint32_t* memb = reinterpret_cast<uint32_t*>(smem + a.sm.memb);    // dense idx, then ray bases
    uint32_t* membp = reinterpret_cast<uint32_t*>(smem + a.sm.membp);  // pair | mc<<16 | sc<<24
    int* cstart = reinterpret_cast<int*>(smem + a.sm.cstart);
    int* cn = reinterpret_cast<int*>(smem + a.sm.cn);
    double* ksum = reinterpret_cast<double*>(smem + a.sm.ksum);
    int* slot = reinterpret_cast<int*>(smem + a.sm.slot);
    int* kcount = reinterpret_cast<int*>(smem + a.sm.kcount);  // [G] clusters, [G..2G) emitted persons
    T* ksbuf = reinterpret_cast<T*>(smem + a.sm.ks);           // [G*Pout*Jout] keypoint scores
    signed char* cobs = reinterpret_cast<signed char*>(smem + a.sm.cobs);  // [G*Pout][8] person per camera
    unsigned char* clq = smem + a.sm.clq;                                   // [G*Pout] clique flag
    int* wtmp = reinterpret_cast<int*>(smem + a.sm.wtmp);

    // ---- one-time tables ---------------------------------------------------------------
    for (int i = tid; i < C * 9; i += NT) camM[i] = (T)a.cam[(i / 9) * 12 + (i % 9)];
    for (int p = tid; p < a.npairs; p += NT) {
        int mc, sc;
        decode_pair(p, C, mc, sc);
        pairs[p] = make_uchar2((unsigned char)mc, (unsigned char)sc);
        for (int k = 0; k < 3; ++k) {
            const double tm = a.cam[mc * 12 + 9 + k], ts = a.cam[sc * 12 + 9 + k];
            pd[p * 6 + k] = (T)(ts - tm);
            pd[p * 6 + 3 + k] = (T)((tm + ts) / 2);
        }
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
    }
    __syncthreads();

    const int ngroups = (a.F + G - 1) / G;
    auto group_uses_tma = [&](int gr) -> bool {
        const int gc = min(G, a.F - gr * G);
        return a.use_tma && (((gc * R) & 3) == 0);
    };
    // staging buffer `b` (fly mode alternates between two, otherwise always 0)
    auto stage_uv_ptr = [&](int b) { return reinterpret_cast<float2*>(smem + a.sm.stage_uv + b * a.sm.stage_stride_uv); };
    auto stage_s_ptr = [&](int b) { return reinterpret_cast<float*>(smem + a.sm.stage_s + b * a.sm.stage_stride_s); };
    auto issue_load = [&](int gr, int b) {
        const int gc = min(G, a.F - gr * G);
        const uint32_t n = (uint32_t)(gc * R);
        mbar_expect_tx(bar + b, n * 12u);
        bulk_g2s(stage_uv_ptr(b), a.kpts + (size_t)gr * G * R * 2, n * 8u, bar + b);
        bulk_g2s(stage_s_ptr(b), a.scores + (size_t)gr * G * R, n * 4u, bar + b);
    };
    if (tid == 0 && (int)blockIdx.x < ngroups && group_uses_tma(blockIdx.x)) issue_load(blockIdx.x, 0);

    const T inv_dthr = a.inv_dthr;
    const float kst_f = a.prm.kst_f;
    const T dthr2 = a.prm.dthr < 0.0 ? (T)-1 : (T)(a.prm.dthr * a.prm.dthr);
    const float2* uv = nullptr;  // current group's (u,v) and scores (smem staging or global)
    const float* sv = nullptr;

    auto ray_index = [&](int g, int c, int p, int j) -> int { return ((g * C + c) * P + p) * J + j; };
    // ray + score of flat index i (camera c): stored arrays, or recomputed from (u,v) in fly mode
    auto get_ray = [&](int i, int c, V3<T>& h, float& s) {
        if constexpr (FLY) {
            const float2 q = uv[i];
            h = back_project<T>(a.camc + 9 * c, (T)q.x, (T)q.y);
            s = sv[i];
        } else {
            h.x = hx[i];
            h.y = hy[i];
            h.z = hz[i];
            s = sc_[i];
        }
    };
    auto load_pd = [&](int pair, V3<T>& d, V3<T>& mid) {
        const T* q = pd + pair * 6;
        d.x = q[0]; d.y = q[1]; d.z = q[2];
        mid.x = q[3]; mid.y = q[4]; mid.z = q[5];
    };
    // deferred tail of a group: per-person mean score from ksbuf, person counts (never_filter path)
    auto write_pscores = [&](int f0p, int Gp) {
        for (int it = warp; it < Gp * Pout; it += NW) {
            const int g = it / Pout, k = it - g * Pout;
            T s = (T)0;
            for (int j = lane; j < Jout; j += 32) s += ksbuf[it * Jout + j];
            s = warp_sum(s);
            if (lane == 0) {
                const bool has = k < kcount[g];
                a.pscores[(size_t)(f0p + g) * Pout + k] = has ? (float)((double)s / (double)Jout) : 0.f;
                if (k == 0) a.nout[f0p + g] = kcount[g];
            }
        }
    };

    uint32_t phase0 = 0, phase1 = 0;
    int buf = 0;
    int prev_f0 = -1, prev_G = 0;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int f0 = grp * G;
        const int Gc = min(G, a.F - f0);
        const int nray = Gc * R;
        if (FLY && tid == 0) {  // the other staging buffer was last read before the previous group's barriers
            const int next = grp + gridDim.x;
            if (next < ngroups && group_uses_tma(next)) issue_load(next, buf ^ 1);
        }
        // deferred per-person scores of the previous group (its ksbuf/kcount are still intact)
        if (a.never_filter && prev_f0 >= 0) write_pscores(prev_f0, prev_G);
        if (group_uses_tma(grp)) {
            if (buf == 0) {
                mbar_wait(bar, phase0);
                phase0 ^= 1u;
            } else {
                mbar_wait(bar + 1, phase1);
                phase1 ^= 1u;
            }
            uv = stage_uv_ptr(buf);
            sv = stage_s_ptr(buf);
        } else {
            uv = reinterpret_cast<const float2*>(a.kpts) + (size_t)f0 * R;
            sv = a.scores + (size_t)f0 * R;
        }
        // ---- phase 0: counts (+ rays into shared memory unless fly mode) ------------------
        for (int i = tid; i < Gc * C; i += NT) {
            int v = a.counts ? a.counts[(size_t)f0 * C + i] : P;
            cnt[i] = max(0, min(P, v));
        }
        if constexpr (!FLY) {
            for (int row = warp; row < Gc * C * P; row += NW) {  // one (frame, camera, person) row per warp
                const int c = (row / P) % C;
                const T* M = camM + 9 * c;
                for (int j = lane; j < J; j += 32) {
                    const int i = row * J + j;
                    const float2 p2 = uv[i];
                    const V3<T> h = back_project<T>(M, (T)p2.x, (T)p2.y);
                    hx[i] = h.x;
                    hy[i] = h.y;
                    hz[i] = h.z;
                    sc_[i] = sv[i];
                }
            }
        }
        __syncthreads();
        if (!FLY && tid == 0) {  // the single staging buffer is free again: prefetch the next group
            const int next = grp + gridDim.x;
            if (next < ngroups && group_uses_tma(next)) issue_load(next, 0);
        }

        // ---- phase 1a: keep flags --------------------------------------------------------
        if (a.all_kept) {
            // ast <= 0 and kst >= 0: the mean-score gate can never reject a valid candidate
            for (int n = tid; n < Gc * ncand; n += NT) {
                const int g = n / ncand, c = n - g * ncand;
                const int pair = c / PP, pm = (c / P) % P, ps = c % P;
                keep[n] = (pm < cnt[g * C + pairs[pair].x] && ps < cnt[g * C + pairs[pair].y]) ? 1 : 0;
            }
        } else {
            const int items = Gc * a.npairs * P;  // (frame, camera pair, main person)
            for (int it = warp; it < items; it += NW) {
                const int g = it / (a.npairs * P), r = it - g * (a.npairs * P);
                const int pair = r / P, pm = r - pair * P;
                const int mc = pairs[pair].x, sc = pairs[pair].y;
                const int nbase = g * ncand + (pair * P + pm) * P;
                const int ncs = cnt[g * C + sc];
                if (pm >= cnt[g * C + mc]) {
                    for (int ps = lane; ps < P; ps += 32) keep[nbase + ps] = 0;
                    continue;
                }
                V3<T> d, mid;
                load_pd(pair, d, mid);
                const int rm = ray_index(g, mc, pm, 0);
                constexpr int NC = NCH > 0 ? NCH : 1;
                V3<T> hm[NC];
                T Am[NC], smT[NC];
                bool lowm[NC];
                if constexpr (NCH > 0) {
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) {
                        const int j = min(ch * 32 + lane, J - 1);
                        float s;
                        get_ray(rm + j, mc, hm[ch], s);
                        Am[ch] = dot3(hm[ch], hm[ch]);
                        smT[ch] = (T)s;
                        lowm[ch] = s < kst_f;
                    }
                }
                for (int ps = 0; ps < P; ++ps) {
                    if (ps >= ncs) {
                        if (lane == 0) keep[nbase + ps] = 0;
                        continue;
                    }
                    const int rs = ray_index(g, sc, ps, 0);
                    T sum = (T)0;
                    float err = 0.f;  // float32 only: error bound of `sum` in units of kDistDelta
                    if constexpr (NCH > 0) {
#pragma unroll
                        for (int ch = 0; ch < NCH; ++ch) {
                            const int j = ch * 32 + lane;
                            if (j < J) {
                                V3<T> hs;
                                float ss;
                                get_ray(rs + j, sc, hs, ss);
                                sum += keep_score(hm[ch], hs, d, smT[ch] + (T)ss, lowm[ch] || ss < kst_f, dthr2, err);
                            }
                        }
                    } else {
                        for (int j = lane; j < J; j += 32) {
                            V3<T> h0, hs;
                            float s0, ss;
                            get_ray(rm + j, mc, h0, s0);
                            get_ray(rs + j, sc, hs, ss);
                            sum += keep_score(h0, hs, d, (T)s0 + (T)ss, s0 < kst_f || ss < kst_f, dthr2, err);
                        }
                    }
                    sum = warp_sum(sum);
                    double avg = (double)sum / (double)J;
                    if constexpr (sizeof(T) == 4) {
                        // float32 sum closer to the threshold than its own error bound: the whole warp
                        // redoes this candidate in float64 from the raw inputs (discrete decision)
                        err = warp_sum(err);
                        const double slack = (double)kDistDelta * (double)err + 1e-5 * fabs((double)sum);
                        if (!(fabs((double)sum - a.prm.ast * (double)J) > slack))
                            avg = candidate_mean_f64(a, f0 + g, mc, pm, sc, ps, lane);
                    }
                    if (lane == 0) keep[nbase + ps] = (avg < a.prm.ast) ? 0 : 1;  // NaN mean is kept (Q9)
                }
            }
            __syncthreads();
        }
        // ---- phase 1b: centre-joint midpoint of every kept candidate ------------------------
        // Always float64 from the raw pixel coordinates (global memory, L2-resident after the staging
        // copy): the clustering test |centre - main| > tol is a discrete decision and must not depend on
        // the compute precision of the fuse.
        for (int n = tid; n < Gc * ncand; n += NT) {
            if (!keep[n]) continue;
            const int g = n / ncand, c = n - g * ncand;
            const int pair = c / PP, pm = (c / P) % P, ps = c % P;
            const int mc = pairs[pair].x, sc = pairs[pair].y;
            const float2* kg = reinterpret_cast<const float2*>(a.kpts) + (size_t)(f0 + g) * R;
            const float2 qm = kg[(mc * P + pm) * J + a.prm.center], qs = kg[(sc * P + ps) * J + a.prm.center];
            const V3<double> hm = back_project<double>(a.cam + 12 * mc, (double)qm.x, (double)qm.y);
            const V3<double> hs = back_project<double>(a.cam + 12 * sc, (double)qs.x, (double)qs.y);
            V3<double> d, mid;
            d.x = a.cam[12 * sc + 9] - a.cam[12 * mc + 9];
            d.y = a.cam[12 * sc + 10] - a.cam[12 * mc + 10];
            d.z = a.cam[12 * sc + 11] - a.cam[12 * mc + 11];
            mid.x = (a.cam[12 * mc + 9] + a.cam[12 * sc + 9]) / 2;
            mid.y = (a.cam[12 * mc + 10] + a.cam[12 * sc + 10]) / 2;
            mid.z = (a.cam[12 * mc + 11] + a.cam[12 * sc + 11]) / 2;
            const PairSol<double> s = pair_solve(hm, hs, d);
            const V3<double> w = pair_midpoint(s, hm, hs, mid);
            cen[3 * n] = w.x;
            cen[3 * n + 1] = w.y;
            cen[3 * n + 2] = w.z;
        }
        __syncthreads();

        // ---- phase 2a: ordered compaction + greedy clustering --------------------------------
        if (ncand > 64) {
            for (int g = 0; g < Gc; ++g) {  // whole CTA, one frame at a time
                const int K = cluster_block<NT>(ncand, keep + g * ncand, klist + g * ncand, cen + (size_t)3 * g * ncand,
                                                ab + g * ncand, memb + g * ncand, cstart + g * ncand, cn + g * ncand,
                                                wtmp, a.tol2, a.prm.num_tol);
                if (tid == 0) kcount[g] = K;
            }
        } else {
            for (int g = warp; g < Gc; g += NW) {  // one warp per frame
                const unsigned lt = (1u << lane) - 1u;
                uint32_t* kl = klist + g * ncand;
                int nk = 0;
                for (int base = 0; base < ncand; base += 32) {
                    const int i = base + lane;
                    const bool k = (i < ncand) && keep[g * ncand + i];
                    const unsigned b = __ballot_sync(kFull, k);
                    if (k) kl[nk + __popc(b & lt)] = (uint32_t)i;
                    nk += __popc(b);
                }
                __syncwarp();
                const int K = cluster_warp(nk, kl, cen + (size_t)3 * g * ncand, ab + g * ncand, memb + g * ncand,
                                           cstart + g * ncand, cn + g * ncand, a.tol2, a.prm.num_tol, lane);
                if (lane == 0) kcount[g] = K;
            }
        }
        __syncthreads();

        // ---- phase 2b: member tables + clique detection, one warp per cluster ------------------
        for (int g = 0; g < Gc; ++g) {
            const int K = kcount[g];
            for (int k = warp; k < K; k += NW) {
                const int n = cn[g * ncand + k];
                const int st = g * ncand + cstart[g * ncand + k];
                const bool try_clique = CM > 0 && a.never_filter && k < Pout && n <= CM * (CM - 1) / 2;
                int mc = -1, sc = -1, pm = 0, ps = 0;
                for (int m0 = 0; m0 < n; m0 += 32) {
                    const int m = m0 + lane;
                    if (m < n) {
                        const int c = (int)memb[st + m];
                        const int pair = c / PP;
                        pm = (c / P) % P;
                        ps = c % P;
                        mc = pairs[pair].x;
                        sc = pairs[pair].y;
                        memb[st + m] = (uint32_t)ray_index(g, mc, pm, 0) | ((uint32_t)ray_index(g, sc, ps, 0) << 16);
                        membp[st + m] = (uint32_t)pair | ((uint32_t)mc << 16) | ((uint32_t)sc << 24);
                    }
                }
                if (try_clique) {  // n <= 28: every member sits in one lane
                    bool bad = false;
                    int ncam = 0;
                    for (int c = 0; c < C; ++c) {
                        const bool hit_m = (mc == c), hit_s = (sc == c);
                        const int person = hit_m ? pm : ps;
                        const unsigned hit = __ballot_sync(kFull, hit_m || hit_s);
                        int obs = -1;
                        if (hit) {
                            obs = __shfl_sync(kFull, person, __ffs(hit) - 1);
                            bad |= __ballot_sync(kFull, (hit_m || hit_s) && person != obs) != 0u;
                            ++ncam;
                        }
                        if (lane == 0) cobs[(g * Pout + k) * kCliqueMax + c] = (signed char)obs;
                    }
                    if (lane == 0) clq[g * Pout + k] = (!bad && n == ncam * (ncam - 1) / 2) ? 1 : 0;
                } else if (lane == 0 && k < Pout) {
                    clq[g * Pout + k] = 0;
                }
            }
        }
        __syncthreads();

        // ---- phase 3: fuse ----------------------------------------------------------------
        // Generic member loop: re-solve each member pair and accumulate sum(w), sum(w*W) in list
        // order (reference triangulation.py:138-148).  The main ray is reused while it repeats.
        auto fuse_members = [&](int g, int k, int j, T& X, T& Y, T& Z) -> T {
            const int n = cn[g * ncand + k];
            const int st = g * ncand + cstart[g * ncand + k];
            T S = (T)0;
            X = Y = Z = (T)0;
            uint32_t prev = 0xffffffffu;
            V3<T> hm;
            T Am = (T)0, smT = (T)0;
            bool lowm = false;
            for (int m = 0; m < n; ++m) {
                const uint32_t rb = memb[st + m], pp = membp[st + m];
                const int pair = pp & 0xffff;
                const uint32_t rmb = rb & 0xffffu;
                if (rmb != prev) {
                    float s;
                    get_ray((int)rmb + j, (pp >> 16) & 0xff, hm, s);
                    Am = dot3(hm, hm);
                    smT = (T)s;
                    lowm = s < kst_f;
                    prev = rmb;
                }
                V3<T> hs, d, mid;
                float ss;
                get_ray((int)(rb >> 16) + j, pp >> 24, hs, ss);
                load_pd(pair, d, mid);
                const PairSol<T> s = pair_solve_a(hm, Am, hs, dot3(hs, hs), d);
                T w;
                const T gg = pair_weights(s, smT + (T)ss, lowm || ss < kst_f, inv_dthr, w);
                const V3<T> v = pair_v(s, hm, hs);
                S += w;
                X = fma(gg, v.x, fma(w, mid.x, X));
                Y = fma(gg, v.y, fma(w, mid.y, Y));
                Z = fma(gg, v.z, fma(w, mid.z, Z));
            }
            return finish_joint(S, n, X, Y, Z);
        };
        // Clique cluster: rays of the <= CM observations in registers, all pairs unrolled, pair and
        // camera constants from the kernel-parameter constant bank.
        auto fuse_clique = [&](int g, int k, int j, T& X, T& Y, T& Z) -> T {
            constexpr int CMX = CM > 0 ? CM : 1;
            const signed char* ob = cobs + (g * Pout + k) * kCliqueMax;
            V3<T> h[CMX];
            T A[CMX], sT[CMX];
            bool on[CMX], low[CMX];
#pragma unroll
            for (int c = 0; c < CMX; ++c) {
                const int p = (c < C) ? (int)ob[c] : -1;
                on[c] = p >= 0;
                float s;
                get_ray(ray_index(g, c < C ? c : 0, on[c] ? p : 0, j), c, h[c], s);
                A[c] = dot3(h[c], h[c]);
                sT[c] = (T)s;
                low[c] = s < kst_f;
            }
            T S = (T)0;
            X = Y = Z = (T)0;
#pragma unroll
            for (int x = 0; x < CMX - 1; ++x) {
#pragma unroll
                for (int y = x + 1; y < CMX; ++y) {
                    if (on[x] && on[y]) {
                        const int e = (x * CMX - x * (x + 1) / 2 + y - x - 1) * 6;
                        V3<T> d, mid;
                        d.x = a.pdc[e]; d.y = a.pdc[e + 1]; d.z = a.pdc[e + 2];
                        mid.x = a.pdc[e + 3]; mid.y = a.pdc[e + 4]; mid.z = a.pdc[e + 5];
                        const PairSol<T> sol = pair_solve_a(h[x], A[x], h[y], A[y], d);
                        T w;
                        const T gg = pair_weights(sol, sT[x] + sT[y], low[x] || low[y], inv_dthr, w);
                        const V3<T> v = pair_v(sol, h[x], h[y]);
                        S += w;
                        X = fma(gg, v.x, fma(w, mid.x, X));
                        Y = fma(gg, v.y, fma(w, mid.y, Y));
                        Z = fma(gg, v.z, fma(w, mid.z, Z));
                    }
                }
            }
            return finish_joint(S, cn[g * ncand + k], X, Y, Z);
        };

        const int PJo = Pout * Jout;
        if (a.never_filter) {
            // score_tol <= 0 and kst >= 0: no person can be rejected, output slot == cluster index.
            // Flattened (frame, slot, joint) lanes: no idle lanes when Jout is not a multiple of 32.
            const int Q = Gc * PJo;
            const float invJ = 1.0f / (float)Jout;
            for (int q = tid; q < Q; q += NT) {
                const int row = (int)(((float)q + 0.5f) * invJ);  // exact for q < 2^22
                const int j = q - row * Jout;
                const int g = row / Pout, k = row - g * Pout;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                T ks = (T)0;
                if (k < kcount[g]) {
                    T X, Y, Z;
                    if (CM > 0 && clq[row])
                        ks = fuse_clique(g, k, j, X, Y, Z);
                    else
                        ks = fuse_members(g, k, j, X, Y, Z);
                    o = make_float4((float)X, (float)Y, (float)Z, (float)ks);
                }
                ksbuf[q] = ks;
                reinterpret_cast<float4*>(a.out)[(size_t)f0 * PJo + q] = o;
            }
            prev_f0 = f0;
            prev_G = Gc;
        } else {
            // general case: a person may be rejected by condense_score_tol, which shifts the output
            // slots of later persons.  Pass A: per-cluster mean score; slots; pass B: write.
            for (int g = 0; g < Gc; ++g) {
                const int K = kcount[g];
                for (int k = warp; k < K; k += NW) {
                    T s = (T)0;
                    for (int j = lane; j < Jout; j += 32) {
                        T X, Y, Z;
                        s += fuse_members(g, k, j, X, Y, Z);
                    }
                    s = warp_sum(s);
                    if (lane == 0) ksum[g * ncand + k] = (double)s / (double)Jout;
                }
            }
            __syncthreads();
            for (int g = warp; g < Gc; g += NW) {
                const unsigned lt = (1u << lane) - 1u;
                const int K = kcount[g];
                int emitted = 0;
                for (int base = 0; base < K; base += 32) {
                    const int k = base + lane;
                    const bool pass = (k < K) && !(ksum[g * ncand + k] < a.prm.score_tol);
                    const unsigned b = __ballot_sync(kFull, pass);
                    if (k < K) slot[g * ncand + k] = pass ? emitted + __popc(b & lt) : -1;
                    emitted += __popc(b);
                }
                if (lane == 0) {
                    kcount[G + g] = emitted;
                    a.nout[f0 + g] = emitted;
                }
            }
            __syncthreads();
            for (int g = 0; g < Gc; ++g) {
                const int K = kcount[g];
                for (int k = warp; k < K; k += NW) {
                    const int s = slot[g * ncand + k];
                    if (s < 0 || s >= Pout) continue;
                    float4* o = reinterpret_cast<float4*>(a.out) + ((size_t)(f0 + g) * Pout + s) * Jout;
                    for (int j = lane; j < Jout; j += 32) {
                        T X, Y, Z;
                        const T ks = fuse_members(g, k, j, X, Y, Z);
                        o[j] = make_float4((float)X, (float)Y, (float)Z, (float)ks);
                    }
                    if (lane == 0) a.pscores[(size_t)(f0 + g) * Pout + s] = (float)ksum[g * ncand + k];
                }
                const int emitted = kcount[G + g];
                for (int s = emitted + warp; s < Pout; s += NW) {
                    float4* o = reinterpret_cast<float4*>(a.out) + ((size_t)(f0 + g) * Pout + s) * Jout;
                    for (int j = lane; j < Jout; j += 32) o[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (lane == 0) a.pscores[(size_t)(f0 + g) * Pout + s] = 0.f;
                }
            }
        }
        if constexpr (FLY) buf ^= 1;
        __syncthreads();  // rays, tables and the staging buffer are rewritten by the next group
    }
    if (a.never_filter && prev_f0 >= 0) write_pscores(prev_f0, prev_G);
}

}  // namespace snowtri
