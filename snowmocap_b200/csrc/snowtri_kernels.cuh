// sm_100a kernels of the triangulation path.  See DESIGN.md for the data layout and the
// per-kernel rooflines.  Reference behaviour being reproduced:
//   rays        snowvision/camera.py:234-253
//   pair solve  snowvision/triangulation.py:24-31
//   candidates  snowvision/triangulation.py:50-93
//   condense    snowvision/triangulation.py:95-162   (quirks Q1-Q12 of SURVEY.md 8a)
#pragma once
#include "snowtri_math.cuh"

namespace snowtri {

constexpr int kThreads = 256;          // block size of the candidate / condense kernels
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kCliqueMax = 8;          // the register-resident fuse path handles up to 8 cameras

struct Params {
    double kst, ast, dthr, cond_tol, score_tol;
    float kst_f;  // smallest float >= kst: (float s < kst_f) <=> ((double)s < kst)
    int num_tol, center;
};

// Byte offsets of the shared-memory regions of the fused kernel (computed on the host).
struct FusedSmem {
    int cam, pairs, pd, stage_uv, stage_s, hx, hy, hz, sc, cnt, cen, keep, ab, klist, memb, membp, cstart, cn,
        ksum, slot, kcount, ks, cobs, clq, total;
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
    // pair constants (d = ts - tm, mid = (tm + ts)/2) in the triangular order of the CM-camera
    // clique path: entry (a*CM - a*(a+1)/2 + b - a - 1)*6; read as constant-bank operands
    T pdc[kCliqueMax * (kCliqueMax - 1) / 2 * 6];
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

// Greedy clustering of one frame by one warp (reference triangulation.py:107-134).
//   N      kept candidates, klist[i] = dense index of the i-th kept candidate (reference list order)
//   cen    centre-joint midpoint of a dense candidate
//   out    memb (dense indices grouped by cluster, in list order), cstart/cn per emitted cluster
// Returns the number of clusters that pass num_tol.  `ab` is N bytes of scratch.
template <typename CenFn>
__device__ int cluster_warp(int N, const uint32_t* klist, CenFn cen, unsigned char* ab, uint32_t* memb,
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
// Fused kernel: persistent CTAs, each iteration handles a group of G consecutive frames whose
// rays fit in shared memory.  Phases per group (separated by __syncthreads):
//   0   TMA-staged (u,v,score) -> world rays (SoA hx/hy/hz + score) in smem; next group prefetched
//   1a  every dense candidate: mean gated score -> keep flag.  One warp per (camera pair, main
//       person): the main ray of every joint chunk stays in registers while the sub persons loop
//   1b  centre-joint midpoint of every kept candidate (one thread each, float64)
//   2   one warp per frame: ordered compaction, greedy clustering, member tables, clique detection
//   3   one (cluster, joint) per lane: score-weighted fuse, float4 store.  Clique clusters (one
//       person per camera, all pairs present) keep their <= CM rays in registers and unroll the pairs
//   4   per-person mean score, unused slots zeroed
// Template: T compute type, NT threads, CM cameras of the unrolled clique path (0 = off),
//           NCH joint chunks held in registers by phase 1a (0 = generic loop, any J).
template <typename T, int NT, int CM, int NCH>
__global__ void __launch_bounds__(NT, 512 / NT) fused_kernel(const __grid_constant__ FusedArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = a.C, P = a.P, J = a.J, Jout = a.Jout, Pout = a.Pout, R = a.R, G = a.G;
    const int ncand = a.ncand, PP = P * P, PJ = P * J;

    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    T* camM = reinterpret_cast<T*>(smem + a.sm.cam);
    uchar2* pairs = reinterpret_cast<uchar2*>(smem + a.sm.pairs);
    T* pd = reinterpret_cast<T*>(smem + a.sm.pd);  // per pair: d (3), mid (3)
    float2* stage_uv = reinterpret_cast<float2*>(smem + a.sm.stage_uv);
    float* stage_s = reinterpret_cast<float*>(smem + a.sm.stage_s);
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
    uint32_t* membp = reinterpret_cast<uint32_t*>(smem + a.sm.membp);  // pair | pm<<16 | ps<<24
    int* cstart = reinterpret_cast<int*>(smem + a.sm.cstart);
    int* cn = reinterpret_cast<int*>(smem + a.sm.cn);
    double* ksum = reinterpret_cast<double*>(smem + a.sm.ksum);
    int* slot = reinterpret_cast<int*>(smem + a.sm.slot);
    int* kcount = reinterpret_cast<int*>(smem + a.sm.kcount);  // [G] clusters, [G..2G) emitted persons
    T* ksbuf = reinterpret_cast<T*>(smem + a.sm.ks);           // [G*Pout*Jout] keypoint scores
    signed char* cobs = reinterpret_cast<signed char*>(smem + a.sm.cobs);  // [G*Pout][8] person per camera
    unsigned char* clq = smem + a.sm.clq;                                   // [G*Pout] clique flag

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
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();

    const int ngroups = (a.F + G - 1) / G;
    uint32_t phase = 0;
    auto group_uses_tma = [&](int gr) -> bool {
        const int gc = min(G, a.F - gr * G);
        return a.use_tma && (((gc * R) & 3) == 0);
    };
    auto issue_load = [&](int gr) {
        const int gc = min(G, a.F - gr * G);
        const uint32_t n = (uint32_t)(gc * R);
        mbar_expect_tx(bar, n * 12u);
        bulk_g2s(stage_uv, a.kpts + (size_t)gr * G * R * 2, n * 8u, bar);
        bulk_g2s(stage_s, a.scores + (size_t)gr * G * R, n * 4u, bar);
    };
    if (tid == 0 && (int)blockIdx.x < ngroups && group_uses_tma(blockIdx.x)) issue_load(blockIdx.x);

    const T dthr = (T)a.prm.dthr;
    const float kst_f = a.prm.kst_f;

    auto ray_index = [&](int g, int c, int p, int j) -> int { return ((g * C + c) * P + p) * J + j; };
    auto load_ray = [&](int i) -> V3<T> {
        V3<T> h;
        h.x = hx[i];
        h.y = hy[i];
        h.z = hz[i];
        return h;
    };
    auto load_pd = [&](int pair, V3<T>& d, V3<T>& mid) {
        const T* q = pd + pair * 6;
        d.x = q[0]; d.y = q[1]; d.z = q[2];
        mid.x = q[3]; mid.y = q[4]; mid.z = q[5];
    };

    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int f0 = grp * G;
        const int Gc = min(G, a.F - f0);
        const int nray = Gc * R;
        const float2* uv;
        const float* sv;
        if (group_uses_tma(grp)) {
            mbar_wait(bar, phase);
            phase ^= 1u;
            uv = stage_uv;
            sv = stage_s;
        } else {
            uv = reinterpret_cast<const float2*>(a.kpts) + (size_t)f0 * R;
            sv = a.scores + (size_t)f0 * R;
        }
        // ---- phase 0: counts + rays ------------------------------------------------------
        for (int i = tid; i < Gc * C; i += NT) {
            int v = a.counts ? a.counts[(size_t)f0 * C + i] : P;
            cnt[i] = max(0, min(P, v));
        }
        for (int i = tid; i < nray; i += NT) {
            const int c = (i / PJ) % C;
            const float2 p2 = uv[i];
            const V3<T> h = back_project<T>(camM + 9 * c, (T)p2.x, (T)p2.y);
            hx[i] = h.x;
            hy[i] = h.y;
            hz[i] = h.z;
            sc_[i] = sv[i];
        }
        __syncthreads();
        if (tid == 0) {
            const int next = grp + gridDim.x;
            if (next < ngroups && group_uses_tma(next)) issue_load(next);
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
                if constexpr (NCH > 0) {
                    V3<T> hm[NCH];
                    T Am[NCH];
                    float sm[NCH];
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) {
                        const int j = min(ch * 32 + lane, J - 1);
                        hm[ch] = load_ray(rm + j);
                        Am[ch] = dot3(hm[ch], hm[ch]);
                        sm[ch] = sc_[rm + j];
                    }
                    for (int ps = 0; ps < P; ++ps) {
                        if (ps >= ncs) {
                            if (lane == 0) keep[nbase + ps] = 0;
                            continue;
                        }
                        const int rs = ray_index(g, sc, ps, 0);
                        T sum = (T)0;
#pragma unroll
                        for (int ch = 0; ch < NCH; ++ch) {
                            const int j = ch * 32 + lane;
                            if (j < J) {
                                const V3<T> hs = load_ray(rs + j);
                                const PairSol<T> s = pair_solve_a(hm[ch], Am[ch], hs, dot3(hs, hs), d);
                                const T gg = gated_g(s, sm[ch], sc_[rs + j], kst_f, dthr);
                                sum += (gg + gg) * s.det;
                            }
                        }
                        sum = warp_sum(sum);
                        const double avg = (double)sum / (double)J;
                        if (lane == 0) keep[nbase + ps] = (avg < a.prm.ast) ? 0 : 1;  // NaN mean is kept (Q9)
                    }
                } else {
                    for (int ps = 0; ps < P; ++ps) {
                        if (ps >= ncs) {
                            if (lane == 0) keep[nbase + ps] = 0;
                            continue;
                        }
                        const int rs = ray_index(g, sc, ps, 0);
                        T sum = (T)0;
                        for (int j = lane; j < J; j += 32) {
                            const V3<T> hm = load_ray(rm + j), hs = load_ray(rs + j);
                            const PairSol<T> s = pair_solve(hm, hs, d);
                            const T gg = gated_g(s, sc_[rm + j], sc_[rs + j], kst_f, dthr);
                            sum += (gg + gg) * s.det;
                        }
                        sum = warp_sum(sum);
                        const double avg = (double)sum / (double)J;
                        if (lane == 0) keep[nbase + ps] = (avg < a.prm.ast) ? 0 : 1;
                    }
                }
            }
            __syncthreads();
        }
        // ---- phase 1b: centre-joint midpoint of every kept candidate ------------------------
        for (int n = tid; n < Gc * ncand; n += NT) {
            if (!keep[n]) continue;
            const int g = n / ncand, c = n - g * ncand;
            const int pair = c / PP, pm = (c / P) % P, ps = c % P;
            const int mc = pairs[pair].x, sc = pairs[pair].y;
            const V3<T> hm = load_ray(ray_index(g, mc, pm, a.prm.center));
            const V3<T> hs = load_ray(ray_index(g, sc, ps, a.prm.center));
            V3<T> d, mid;
            load_pd(pair, d, mid);
            const PairSol<T> s = pair_solve(hm, hs, d);
            const V3<T> w = pair_midpoint(s, hm, hs, mid);
            cen[3 * n] = (double)w.x;
            cen[3 * n + 1] = (double)w.y;
            cen[3 * n + 2] = (double)w.z;
        }
        __syncthreads();

        // ---- phase 2: ordered compaction + greedy clustering, one warp per frame ----------
        for (int g = warp; g < Gc; g += NW) {
            const unsigned lt = (1u << lane) - 1u;
            uint32_t* kl = klist + g * ncand;
            uint32_t* mb = memb + g * ncand;
            uint32_t* mp = membp + g * ncand;
            int nk = 0;
            for (int base = 0; base < ncand; base += 32) {
                const int i = base + lane;
                const bool k = (i < ncand) && keep[g * ncand + i];
                const unsigned b = __ballot_sync(kFull, k);
                if (k) kl[nk + __popc(b & lt)] = (uint32_t)i;
                nk += __popc(b);
            }
            __syncwarp();
            const double* cg = cen + (size_t)3 * g * ncand;
            auto cen_fn = [cg](uint32_t n, double& x, double& y, double& z) {
                x = cg[3 * n];
                y = cg[3 * n + 1];
                z = cg[3 * n + 2];
            };
            const int K = cluster_warp(nk, kl, cen_fn, ab + g * ncand, mb, cstart + g * ncand, cn + g * ncand,
                                       a.prm.cond_tol, a.prm.num_tol, lane);
            if (lane == 0) kcount[g] = K;
            __syncwarp();
            // member tables: ray base offsets (main | sub << 16) and (pair | pm << 16 | ps << 24)
            const int nmemb = K > 0 ? cstart[g * ncand + K - 1] + cn[g * ncand + K - 1] : 0;
            for (int i = lane; i < nmemb; i += 32) {
                const int c = (int)mb[i];
                const int pair = c / PP, pm = (c / P) % P, ps = c % P;
                const uint32_t rmb = (uint32_t)ray_index(g, pairs[pair].x, pm, 0);
                const uint32_t rsb = (uint32_t)ray_index(g, pairs[pair].y, ps, 0);
                mb[i] = rmb | (rsb << 16);
                mp[i] = (uint32_t)pair | ((uint32_t)pm << 16) | ((uint32_t)ps << 24);
            }
            __syncwarp();
            if (CM > 0 && a.never_filter) {
                // clique detection: one person per camera and all pairs of the present cameras
                const int kmax = min(K, Pout);
                for (int k = 0; k < kmax; ++k) {
                    const int n = cn[g * ncand + k];
                    const uint32_t* q = mp + cstart[g * ncand + k];
                    int obs = -1;
                    bool conflict = false;
                    if (n <= CM * (CM - 1) / 2 && lane < C) {
                        for (int m = 0; m < n; ++m) {
                            const uint32_t e = q[m];
                            const int pair = e & 0xffff, pm = (e >> 16) & 0xff, ps = e >> 24;
                            if (pairs[pair].x == lane) {
                                conflict |= (obs >= 0 && obs != pm);
                                obs = pm;
                            }
                            if (pairs[pair].y == lane) {
                                conflict |= (obs >= 0 && obs != ps);
                                obs = ps;
                            }
                        }
                    }
                    const int m = __popc(__ballot_sync(kFull, obs >= 0));
                    const bool bad = __ballot_sync(kFull, conflict) != 0u;
                    if (lane < kCliqueMax) cobs[(g * Pout + k) * kCliqueMax + lane] = (signed char)obs;
                    if (lane == 0) clq[g * Pout + k] = (!bad && m >= 2 && n == m * (m - 1) / 2) ? 1 : 0;
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
            T Am = (T)0;
            float sm = 0.f;
            for (int m = 0; m < n; ++m) {
                const uint32_t rb = memb[st + m];
                const int pair = membp[st + m] & 0xffff;
                const uint32_t rmb = rb & 0xffffu;
                if (rmb != prev) {
                    hm = load_ray(rmb + j);
                    Am = dot3(hm, hm);
                    sm = sc_[rmb + j];
                    prev = rmb;
                }
                const int rs = (int)(rb >> 16) + j;
                const V3<T> hs = load_ray(rs);
                V3<T> d, mid;
                load_pd(pair, d, mid);
                const PairSol<T> s = pair_solve_a(hm, Am, hs, dot3(hs, hs), d);
                const T gg = gated_g(s, sm, sc_[rs], kst_f, dthr);
                const V3<T> v = pair_v(s, hm, hs);
                const T w = (gg + gg) * s.det;
                S += w;
                X = fma(gg, v.x, fma(w, mid.x, X));
                Y = fma(gg, v.y, fma(w, mid.y, Y));
                Z = fma(gg, v.z, fma(w, mid.z, Z));
            }
            return finish_joint(S, n, X, Y, Z);
        };
        // Clique cluster: rays of the <= CM observations in registers, all pairs unrolled, pair
        // constants from the kernel-parameter constant bank.
        auto fuse_clique = [&](int g, int k, int j, T& X, T& Y, T& Z) -> T {
            constexpr int CMX = CM > 0 ? CM : 1;
            const signed char* ob = cobs + (g * Pout + k) * kCliqueMax;
            V3<T> h[CMX];
            T A[CMX];
            float s[CMX];
            bool on[CMX];
#pragma unroll
            for (int c = 0; c < CMX; ++c) {
                const int p = (c < C) ? (int)ob[c] : -1;
                on[c] = p >= 0;
                const int r = ray_index(g, c < C ? c : 0, on[c] ? p : 0, j);
                h[c] = load_ray(r);
                A[c] = dot3(h[c], h[c]);
                s[c] = sc_[r];
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
                        const T gg = gated_g(sol, s[x], s[y], kst_f, dthr);
                        const V3<T> v = pair_v(sol, h[x], h[y]);
                        const T w = (gg + gg) * sol.det;
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
            // score_tol <= 0 and kst >= 0: no person can be rejected, output slot == cluster index
            const int Q = Gc * PJo;
            for (int q = tid; q < Q; q += NT) {
                const int g = q / PJo, r = q - g * PJo;
                const int k = r / Jout, j = r - k * Jout;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                T ks = (T)0;
                if (k < kcount[g]) {
                    T X, Y, Z;
                    if (CM > 0 && clq[g * Pout + k])
                        ks = fuse_clique(g, k, j, X, Y, Z);
                    else
                        ks = fuse_members(g, k, j, X, Y, Z);
                    o = make_float4((float)X, (float)Y, (float)Z, (float)ks);
                }
                ksbuf[q] = ks;
                reinterpret_cast<float4*>(a.out)[(size_t)f0 * PJo + q] = o;
            }
            __syncthreads();
            for (int it = warp; it < Gc * Pout; it += NW) {
                const int g = it / Pout, k = it - g * Pout;
                T s = (T)0;
                for (int j = lane; j < Jout; j += 32) s += ksbuf[it * Jout + j];
                s = warp_sum(s);
                if (lane == 0) {
                    const bool has = k < kcount[g];
                    a.pscores[(size_t)(f0 + g) * Pout + k] = has ? (float)((double)s / (double)Jout) : 0.f;
                    if (k == 0) a.nout[f0 + g] = kcount[g];
                }
            }
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
        __syncthreads();  // rays and tables are rewritten by the next group
    }
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
        const int K = cluster_warp(n_in, klist, cen_fn, ab, memb, cstart, cn, a.prm.cond_tol, a.prm.num_tol, lane);
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
