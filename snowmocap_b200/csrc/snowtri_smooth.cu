// Human_Triangulation_Smooth on the device (reference snowvision/triangulation.py:4-22, 164-186, called
// frame after frame by main.py:72-78): one critically damped second-order follower per (person, joint, axis).
//
// The recurrence is sequential in time, so a thread owns one (person, joint) -- three independent axes --
// and walks the frames of the batch in order; the inputs of later frames do not depend on the state and are
// prefetched several frames ahead.  State (xp, y, yd per axis, the person count of the first frame) stays
// on the device between calls, so a clip can be streamed through in batches.
//
// Semantics kept from the reference: the first frame passes through unchanged and seeds the followers with
// x0 = the point; later frames zip the persons of the frame with the followers of the FIRST frame by list
// position, so persons beyond that count are dropped (nsm = min(nout, n0)) and followers of absent persons
// are not advanced; scores pass through untouched.
//
// Long batches are cut into chunks of kChunk frames that run in parallel (the update of a present frame is
// affine in the state (xp, y, yd): s' = A s + b x):
//   pass A  every (chunk, person, joint) runs the recurrence over its chunk from a ZERO state -> b_c, and
//           counts the frames n_c in which its person was present;
//   pass B  start_c+1 = A^n_c start_c + b_c (A^n from a table computed on the host) for every chunk: a prefix scan
//           of the chunk maps under composition, one launch (smooth_carry_scan_kernel);
//   pass C  every (chunk, person, joint) runs the recurrence again from its true start state and writes the
//           smoothed points.
// Inside a chunk the arithmetic is the reference's sequential recurrence; only the hand-over between chunks
// is evaluated differently (agreement with the sequential kernel ~1e-15 relative).
//
// Arithmetic: float64 like the reference.  (x - xp)/T and (...)/k2 are evaluated as multiplications by the
// reciprocals computed once on the host (1 ulp per step away from a true division, damped by the filter).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <cooperative_groups.h>

#include "snowtri_internal.h"

struct snowtri_smooth_state {
    int device, P, J;
    double f, z, r;
    double* d_state;  // [0] initialised, [1] n0, then xp, y, yd as (P, J, 3) each
    // chunk-parallel path
    double* d_work;   // per chunk: b (P,J,9) then start (P,J,9)
    int* d_cnt;       // per chunk: present-frame count per person (P), then a "seeded here" flag
    double* d_apow;   // (kChunk+1, 9) powers of the state matrix for the current delta_time
    size_t work_chunks;
    double apow_T;
    int sequential;   // 1 = always the single-launch sequential kernel
    int force_scan;   // 1 = long batches always take the chunk scan (pass A / B / C)
    int coop_blocks;  // CTAs of smooth_overlap_kernel a cooperative launch can hold (0: not asked yet)
    int forget;       // frames after which the follower's state no longer reaches the output in float64 (0: never within kMaxForget)
    double forget_T;  // delta_time `forget` was computed for
    double apow_host[(128 + 1) * 9];  // (kChunk + 1) matrices
};

namespace snowtri {

struct SmoothArgs {
    void* pts;        // (F, Pout, J, 4) float32 or float64: x, y, z, score (score untouched)
    const int* nout;  // (F)
    int* nsm;         // (F) persons in the smoothed list
    double* state;
    int F, Pout, P, J;
    double T, invT, k1, inv_k2, k3;
};

#ifndef SMOOTH_BATCH_BYTES
#define SMOOTH_BATCH_BYTES 256
#endif
constexpr int kSmoothBatchBytes = SMOOTH_BATCH_BYTES;  // input bytes a thread holds in registers per batch (16 float4 / 8 double4)  -- measured 128: 0.247 ms, 256: 0.240 ms, 512: 0.265 ms per 131 072 frames

template <bool B>
struct SBool {
    static constexpr bool value = B;
};

template <typename V>  // float4 or double4
__global__ void __launch_bounds__(128) smooth_kernel(const SmoothArgs a) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = tid / a.J, j = tid - k * a.J;
    const bool lead = tid == 0;
    if (k >= a.P) return;
    const size_t N = (size_t)a.P * a.J * 3;
    double* sxp = a.state + 2 + ((size_t)k * a.J + j) * 3;
    double* sy = sxp + N;
    double* syd = sy + N;
    bool init = a.state[0] != 0.0;
    int n0 = (int)a.state[1];
    double xp[3], y[3], yd[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        xp[c] = sxp[c];
        y[c] = sy[c];
        yd[c] = syd[c];
    }
    V* pts = reinterpret_cast<V*>(a.pts);
    const size_t stride = (size_t)a.Pout * a.J;  // points per frame
    const bool inrange = k < a.Pout;
    const size_t base = (size_t)(inrange ? k : 0) * a.J + j;
    // inputs a batch of frames at a time (see smooth_chunk_kernel): loads back to back, one wait, then the steps
    constexpr int NB = kSmoothBatchBytes / (int)sizeof(V);
    for (int t0 = 0; t0 < a.F; t0 += NB) {
        V buf[NB];
        int nb[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int t = min(t0 + u, a.F - 1);
            buf[u] = pts[(size_t)t * stride + base];
            nb[u] = a.nout[t];
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int t = t0 + u;
            if (t >= a.F) break;
            const V p = buf[u];
            const int n = min(max(nb[u], 0), a.Pout);
            const double x[3] = {(double)p.x, (double)p.y, (double)p.z};
            if (!init) {  // first frame of the clip: seed the followers, points pass through (reference :177-184)
                init = true;
                n0 = min(n, a.P);
                if (k < n0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        xp[c] = y[c] = x[c];
                        yd[c] = 0.0;
                    }
                }
                if (lead) a.nsm[t] = n;
                continue;
            }
            const int m = min(n, n0);
            if (lead) a.nsm[t] = m;
            if (k < m) {  // reference :15-22
                V o = p;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double xd = (x[c] - xp[c]) * a.invT;
                    xp[c] = x[c];
                    y[c] = y[c] + a.T * yd[c];
                    yd[c] = yd[c] + a.T * (x[c] + a.k3 * xd - y[c] - a.k1 * yd[c]) * a.inv_k2;
                }
                o.x = (decltype(o.x))y[0];
                o.y = (decltype(o.y))y[1];
                o.z = (decltype(o.z))y[2];
                pts[(size_t)t * stride + base] = o;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        sxp[c] = xp[c];
        sy[c] = y[c];
        syd[c] = yd[c];
    }
    // The clip header (state[0] = seeded, state[1] = persons of the first frame) is read by every block at entry, so
    // it is NOT rewritten here: a block that starts after block 0 has finished would see the new header on the
    // clip's first frame.  smooth_finish_kernel, launched behind this kernel on the same stream, writes it.
}


constexpr int kChunk = 128;  // frames per chunk of the parallel path

// One frame of one (person, joint): the reference's update (triangulation.py:15-22) on the three axes, in the bulk
// kernels with the constants folded:  y' = y + T yd;  yd' = cd yd + cx x - cp xp - g y'  with g = T/k2, cd = 1 - g k1,
// cx = g (1 + k3/T), cp = g k3/T -- 5 instead of 7 float64 operations per axis and a carried chain of two (the chunk
// scan of the control points 0.116 -> 0.104 ms; the one-pass kernels are bound by DRAM and did not move).  Same
// algebra, different rounding (~1e-16 per step, damped by the filter); the sequential kernel of the per-frame drop-in
// keeps the reference's operation order.
struct FollowerCoef {
    double T, g, cd, cx, cp;
};
__device__ __forceinline__ FollowerCoef follower_coef(double T, double invT, double k1, double inv_k2, double k3) {
    FollowerCoef q;
    q.T = T;
    q.g = T * inv_k2;
    q.cd = 1.0 - q.g * k1;
    q.cx = q.g * (1.0 + k3 * invT);
    q.cp = q.g * (k3 * invT);
    return q;
}
__device__ __forceinline__ void follower_step1(const FollowerCoef& q, double x, double& xp, double& y, double& yd) {
    const double t1 = fma(q.cx, x, -(q.cp * xp));
    xp = x;
    y = fma(q.T, yd, y);
    yd = fma(-q.g, y, fma(q.cd, yd, t1));
}
__device__ __forceinline__ void follower_step(const FollowerCoef& q, const double* x, double* xp, double* y, double* yd) {
#pragma unroll
    for (int c = 0; c < 3; ++c) follower_step1(q, x[c], xp[c], y[c], yd[c]);
}

struct ChunkArgs {
    SmoothArgs s;
    double* work;  // per chunk: b (P*J*9), start (P*J*9)
    int* cnt;      // per chunk: P counts + 1 seeded flag
    const double* apow;
    int nchunks;
};

// PASS 0: pass A (from a zero state, no output)   PASS 2: pass C (from the true start state, writes)
template <typename V, int PASS>
__global__ void __launch_bounds__(128) smooth_chunk_kernel(const ChunkArgs ca) {
    const SmoothArgs& a = ca.s;
    // dense 1-D grid over (chunk, person, joint): no idle lanes when P*J is not a multiple of the block size
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nthr = a.P * a.J;
    if (gid >= (long long)ca.nchunks * nthr) return;
    const int chunk = (int)(gid / nthr), tid = (int)(gid - (long long)chunk * nthr);
    const int k = tid / a.J, j = tid - k * a.J;
    const int t_begin = chunk * kChunk, t_end = min(a.F, t_begin + kChunk);
    const FollowerCoef coef = follower_coef(a.T, a.invT, a.k1, a.inv_k2, a.k3);
    const bool was_init = a.state[0] != 0.0;
    const int n0 = was_init ? (int)a.state[1] : min(min(max(a.nout[0], 0), a.Pout), a.P);
    const size_t ch = (size_t)k * a.J + j, nch = (size_t)a.P * a.J;
    double* wb = ca.work + (size_t)chunk * nch * 18 + ch * 9;
    double xp[3] = {0, 0, 0}, y[3] = {0, 0, 0}, yd[3] = {0, 0, 0};
    if (PASS == 2) {
        const double* ws = wb + nch * 9;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            xp[c] = ws[c];
            y[c] = ws[3 + c];
            yd[c] = ws[6 + c];
        }
    }
    V* pts = reinterpret_cast<V*>(a.pts);
    const size_t stride = (size_t)a.Pout * a.J;
    const bool inrange = k < a.Pout;
    const size_t base = (size_t)(inrange ? k : 0) * a.J + j;
    int present = 0;
    // Inputs are fetched a batch of frames at a time -- all loads of a batch issued back to back (clamped
    // addresses, no branches), one wait, then the batch's recurrence steps from registers.  (A ring that refills one
    // slot per step makes every step wait for the load it has just issued: a scoreboard wait covers every load
    // assigned to it.)
    constexpr int NB = kSmoothBatchBytes / (int)sizeof(V);
    for (int t0 = t_begin; t0 < t_end; t0 += NB) {
        V buf[NB];
        int nb[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int t = min(t0 + u, t_end - 1);
            buf[u] = pts[(size_t)t * stride + base];
            nb[u] = a.nout[t];
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int t = t0 + u;
            if (t >= t_end) break;
            const V p = buf[u];
            const int n = min(max(nb[u], 0), a.Pout);
            const double x[3] = {(double)p.x, (double)p.y, (double)p.z};
            if (!was_init && t == 0) {  // first frame of the clip: seed, pass through (reference :177-184)
                if (k < n0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        xp[c] = y[c] = x[c];
                        yd[c] = 0.0;
                    }
                }
                if (PASS == 2 && tid == 0) a.nsm[t] = n;
                continue;
            }
            const int m = min(n, n0);
            if (PASS == 2 && tid == 0) a.nsm[t] = m;
            if (k < m) {
                follower_step(coef, x, xp, y, yd);
                ++present;
                if (PASS == 2) {
                    V o = p;
                    o.x = (decltype(o.x))y[0];
                    o.y = (decltype(o.y))y[1];
                    o.z = (decltype(o.z))y[2];
                    pts[(size_t)t * stride + base] = o;
                }
            }
        }
    }
    if (PASS == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            wb[c] = xp[c];
            wb[3 + c] = y[c];
            wb[6 + c] = yd[c];
        }
        if (j == 0) ca.cnt[(size_t)chunk * (a.P + 1) + k] = present;
        if (tid == 0) ca.cnt[(size_t)chunk * (a.P + 1) + a.P] = (!was_init && chunk == 0) ? 1 : 0;
    }
}


// ---- one-pass path: chunks with a warm-up instead of a hand-over ------------------------------------------------
// The follower is a stable filter: W present frames after ANY start state the state is the same to the last bit of a
// float64 (|A^W| < 1e-19 for the state matrix A of one present frame; the host finds W from the powers of A -- 80 frames
// for the reference's shipped f = 2.5, z = 0.75, r = 0 at 30 fps).  So a (chunk, person, joint) thread does not need the state its
// predecessor ends with: it starts from a ZERO state W present frames before its chunk, walks those frames without
// writing, and is exact from its first own frame on.  One launch, the batch is read 1 + W/L times and written once
// (pass A / B / C: read twice, written once, four launches and a work array).  A thread that runs out of history
// before it has seen W present frames (first chunk, a person that was absent) starts from the carried state at frame 0 of
// the batch, i.e. it is the sequential recurrence.
// The points are smoothed IN PLACE, and a warm-up reads frames that belong to earlier chunks: the launch is cooperative
// (every CTA resident) and a grid-wide barrier separates the warm-ups -- which read only frames before the thread's own
// chunk, the carried state and the clip header -- from the chunk walks, which read and write only the thread's own
// frames, the state after the batch and the header.
constexpr int kMaxForget = 256;   // slower filters take the chunk scan

template <typename V>
__global__ void __launch_bounds__(128) smooth_overlap_kernel(const SmoothArgs a, int L, int W, int nchunks) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nthr = a.P * a.J;
    const bool active = gid < (long long)nchunks * nthr;
    const int chunk = active ? (int)(gid / nthr) : 0, tid = active ? (int)(gid - (long long)chunk * nthr) : 0;
    const int k = tid / a.J, j = tid - k * a.J;
    const int t_begin = chunk * L, t_end = min(a.F, t_begin + L);
    const FollowerCoef coef = follower_coef(a.T, a.invT, a.k1, a.inv_k2, a.k3);
    const bool was_init = a.state[0] != 0.0;
    const int n0 = was_init ? (int)a.state[1] : min(min(max(a.nout[0], 0), a.Pout), a.P);
    const size_t N = (size_t)a.P * a.J * 3, ch = ((size_t)k * a.J + j) * 3;
    double xp[3] = {0, 0, 0}, y[3] = {0, 0, 0}, yd[3] = {0, 0, 0};
    V* pts = reinterpret_cast<V*>(a.pts);
    const size_t stride = (size_t)a.Pout * a.J;
    const bool inrange = k < a.Pout;
    const size_t base = (size_t)(inrange ? k : 0) * a.J + j;
    constexpr int NB = kSmoothBatchBytes / (int)sizeof(V);
    // frames [t_from, t_to) in order; OWN: the thread's own frames.  Inputs travel in two half-batches: the loads of the
    // next half are issued before the steps of the current one, so a thread always has loads in flight.  (Measured
    // against one batch of 16 per thread: 0.139 vs 0.140 ms at 256-frame chunks, 0.172 vs 0.168 at 512 -- no gain; nor
    // from 15 instead of 21 float64 operations per step.  At 256-frame chunks the kernel moves 645 MB in 130 us, 76 % of
    // the measured copy bandwidth with reads and writes of 512-byte pieces interleaved; profiles/r3h, r3i.  Also tried:
    // the next half staged in shared memory by cp.async, which holds no scoreboard -- 0.167 ms at 256-frame chunks,
    // 0.164 at 512: slower, profiles/r3j.)
    auto walk = [&](int t_from, int t_to, auto ownc) {
        constexpr bool OWN = decltype(ownc)::value;
        constexpr int NH = NB / 2;
        V bufA[NH], bufB[NH];
        int nbA[NH], nbB[NH];
        auto fetch = [&](V (&buf)[NH], int (&nb)[NH], int t0) {
#pragma unroll
            for (int u = 0; u < NH; ++u) {
                const int t = min(t0 + u, t_to - 1);
                buf[u] = pts[(size_t)t * stride + base];
                nb[u] = a.nout[t];
            }
        };
        auto steps = [&](const V (&buf)[NH], const int (&nb)[NH], int t0) {
#pragma unroll
            for (int u = 0; u < NH; ++u) {
                const int t = t0 + u;
                if (t >= t_to) break;
                const V p = buf[u];
                const int n = min(max(nb[u], 0), a.Pout);
                const double x[3] = {(double)p.x, (double)p.y, (double)p.z};
                if (!was_init && t == 0) {  // first frame of the clip: seed, pass through (reference :177-184)
                    if (k < n0) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            xp[c] = y[c] = x[c];
                            yd[c] = 0.0;
                        }
                    }
                    if (OWN && tid == 0) a.nsm[t] = n;
                    continue;
                }
                const int m = min(n, n0);
                if (OWN && tid == 0) a.nsm[t] = m;
                if (k < m) {
                    follower_step(coef, x, xp, y, yd);
                    if (OWN) {
                        V o = p;
                        o.x = (decltype(o.x))y[0];
                        o.y = (decltype(o.y))y[1];
                        o.z = (decltype(o.z))y[2];
                        pts[(size_t)t * stride + base] = o;
                    }
                }
            }
        };
        if (t_from >= t_to) return;
        fetch(bufA, nbA, t_from);
        for (int t0 = t_from; t0 < t_to; t0 += 2 * NH) {
            if (t0 + NH < t_to) fetch(bufB, nbB, t0 + NH);
            steps(bufA, nbA, t0);
            if (t0 + 2 * NH < t_to) fetch(bufA, nbA, t0 + 2 * NH);
            if (t0 + NH < t_to) steps(bufB, nbB, t0 + NH);
        }
    };
    if (active) {
        // warm-up start: W present frames of this person before the chunk (the seed frame of a clip is not a step)
        // (person counts fetched eight frames at a time: one by one this walk is ~W dependent L2 round trips)
        int tw = t_begin, need = W;
        if (k >= min(n0, a.Pout)) tw = 0;   // a follower this batch never advances: no history to look for, its state is carried over
        while (tw > 0 && need > 0) {
            int nb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) nb[u] = a.nout[max(tw - 1 - u, 0)];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (tw > 0 && need > 0) {
                    --tw;
                    const int m = min(min(max(nb[u], 0), a.Pout), n0);
                    if (k < m && (was_init || tw > 0)) --need;
                }
            }
        }
        if (need > 0) {  // out of history: the walk starts at frame 0 of the batch, from the state the last batch left
            tw = k >= min(n0, a.Pout) ? t_begin : 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                xp[c] = a.state[2 + ch + c];
                y[c] = a.state[2 + N + ch + c];
                yd[c] = a.state[2 + 2 * N + ch + c];
            }
        }
        walk(tw, t_begin, SBool<false>{});
    }
    cooperative_groups::this_grid().sync();
    if (!active) return;
    walk(t_begin, t_end, SBool<true>{});
    if (chunk == nchunks - 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            a.state[2 + ch + c] = xp[c];
            a.state[2 + N + ch + c] = y[c];
            a.state[2 + 2 * N + ch + c] = yd[c];
        }
        if (tid == 0 && !was_init) {
            a.state[1] = (double)n0;
            a.state[0] = 1.0;
        }
    }
}

// Pass B, one launch: the chunk maps of a (person, joint, axis) channel are affine, so their hand-over is a prefix
// scan under composition.  One CTA per channel, a thread per chunk (blocks of kScanThreads chunks, the state carried
// from block to block): warp-level Hillis-Steele scan by shuffles (5 levels), the 16 warp totals scanned by warp 0,
// every thread then applies the prefix of the chunks before it to the clip state.  Two memory round trips and ~10
// compositions deep.  (Round 1 walked the chunks in three launches over groups of 32 -- compose the group maps, walk
// the groups, expand to chunk starts: 96 dependent round trips, 24.8 + 26.1 + 26.8 us per 1024 chunks, profiles/r2n;
// this kernel takes 39 us, profiles/r2p.)  Composition order differs from a sequential walk by rounding only (~1e-15).
constexpr int kScanThreads = 512;
struct AMap {  // s -> M s + o on (xp, y, yd)
    double M[9], o[3];
};
__device__ __forceinline__ AMap amap_compose(const AMap& later, const AMap& earlier) {  // later(earlier(s))
    AMap r;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int q = 0; q < 3; ++q)
            r.M[3 * i + q] = later.M[3 * i] * earlier.M[q] + later.M[3 * i + 1] * earlier.M[3 + q] + later.M[3 * i + 2] * earlier.M[6 + q];
        r.o[i] = later.o[i] + later.M[3 * i] * earlier.o[0] + later.M[3 * i + 1] * earlier.o[1] + later.M[3 * i + 2] * earlier.o[2];
    }
    return r;
}
__device__ __forceinline__ AMap amap_shfl_up(const AMap& m, int d) {
    AMap r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.M[i] = __shfl_up_sync(0xffffffffu, m.M[i], d);
#pragma unroll
    for (int i = 0; i < 3; ++i) r.o[i] = __shfl_up_sync(0xffffffffu, m.o[i], d);
    return r;
}
__device__ __forceinline__ AMap amap_identity() {
    AMap r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.M[i] = (i % 4 == 0) ? 1.0 : 0.0;
    r.o[0] = r.o[1] = r.o[2] = 0.0;
    return r;
}

__global__ void __launch_bounds__(kScanThreads) smooth_carry_scan_kernel(const ChunkArgs ca) {
    __shared__ double apow_s[(kChunk + 1) * 9];
    __shared__ double wtot[kScanThreads / 32][12];
    __shared__ double carry[3];
    const SmoothArgs& a = ca.s;
    const int tid = blockIdx.x;  // channel = (person, joint, axis)
    const int chn = tid / 3, c = tid - chn * 3, k = chn / a.J;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (kChunk + 1) * 9; i += blockDim.x) apow_s[i] = ca.apow[i];
    const size_t nch = (size_t)a.P * a.J, N = nch * 3;
    const bool was_init = a.state[0] != 0.0;
    const int n0 = was_init ? (int)a.state[1] : min(min(max(a.nout[0], 0), a.Pout), a.P);
    double* sxp = a.state + 2 + (size_t)chn * 3 + c;
    double* sy = sxp + N;
    double* syd = sy + N;
    if (threadIdx.x == 0) {
        carry[0] = *sxp;
        carry[1] = *sy;
        carry[2] = *syd;
    }
    __syncthreads();
    for (int base = 0; base < ca.nchunks; base += kScanThreads) {
        const int chunk = base + threadIdx.x;
        AMap m = amap_identity();
        double* wb = nullptr;
        if (chunk < ca.nchunks) {
            wb = ca.work + (size_t)chunk * nch * 18 + (size_t)chn * 9 + c;
            const int n = ca.cnt[(size_t)chunk * (a.P + 1) + k];
            const bool seeded = ca.cnt[(size_t)chunk * (a.P + 1) + a.P] != 0 && k < n0;
            m.o[0] = wb[0]; m.o[1] = wb[3]; m.o[2] = wb[6];
            const double* A = apow_s + n * 9;
#pragma unroll
            for (int i = 0; i < 9; ++i) m.M[i] = seeded ? 0.0 : A[i];  // a chunk that seeded the followers is the constant map
        }
        // inclusive scan inside the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const AMap up = amap_shfl_up(m, d);
            if (lane >= d) m = amap_compose(m, up);
        }
        if (lane == 31) {
#pragma unroll
            for (int i = 0; i < 9; ++i) wtot[warp][i] = m.M[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) wtot[warp][9 + i] = m.o[i];
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the warp totals
            AMap t = amap_identity();
            if (lane < kScanThreads / 32) {
#pragma unroll
                for (int i = 0; i < 9; ++i) t.M[i] = wtot[lane][i];
#pragma unroll
                for (int i = 0; i < 3; ++i) t.o[i] = wtot[lane][9 + i];
            }
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const AMap up = amap_shfl_up(t, d);
                if (lane >= d) t = amap_compose(t, up);
            }
            AMap ex = amap_shfl_up(t, 1);
            if (lane == 0) ex = amap_identity();
            if (lane < kScanThreads / 32) {
#pragma unroll
                for (int i = 0; i < 9; ++i) wtot[lane][i] = ex.M[i];
#pragma unroll
                for (int i = 0; i < 3; ++i) wtot[lane][9 + i] = ex.o[i];
            }
        }
        __syncthreads();
        AMap pre;  // all chunks of the preceding warps of this block
#pragma unroll
        for (int i = 0; i < 9; ++i) pre.M[i] = wtot[warp][i];
#pragma unroll
        for (int i = 0; i < 3; ++i) pre.o[i] = wtot[warp][9 + i];
        const AMap incl = amap_compose(m, pre);      // chunks base .. chunk
        AMap excl = amap_shfl_up(incl, 1);           // chunks base .. chunk - 1
        if (lane == 0) excl = pre;
        const double s0 = carry[0], s1 = carry[1], s2 = carry[2];
        if (chunk < ca.nchunks) {
            double* ws = wb + nch * 9;
            ws[0] = excl.o[0] + excl.M[0] * s0 + excl.M[1] * s1 + excl.M[2] * s2;
            ws[3] = excl.o[1] + excl.M[3] * s0 + excl.M[4] * s1 + excl.M[5] * s2;
            ws[6] = excl.o[2] + excl.M[6] * s0 + excl.M[7] * s1 + excl.M[8] * s2;
        }
        __syncthreads();  // everyone has read the carried state
        if (threadIdx.x == kScanThreads - 1) {  // identity maps beyond the last chunk: this is the state after the block
            carry[0] = incl.o[0] + incl.M[0] * s0 + incl.M[1] * s1 + incl.M[2] * s2;
            carry[1] = incl.o[1] + incl.M[3] * s0 + incl.M[4] * s1 + incl.M[5] * s2;
            carry[2] = incl.o[2] + incl.M[6] * s0 + incl.M[7] * s1 + incl.M[8] * s2;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *sxp = carry[0];
        *sy = carry[1];
        *syd = carry[2];
    }
}

__global__ void smooth_finish_kernel(double* state, const int* nout, int Pout, int P) {
    if (state[0] == 0.0) {
        state[1] = (double)min(min(max(nout[0], 0), Pout), P);
        state[0] = 1.0;
    }
}

}  // namespace snowtri

using namespace snowtri;

extern "C" int snowtri_smooth_create(snowtri_t* h, snowtri_smooth_t** out, int max_persons, int J, double f,
                                     double z, double r) {
    if (!h || !out) return fail(h, SNOWTRI_E_ARG, "snowtri_smooth_create: NULL argument");
    *out = nullptr;
    if (max_persons < 1 || J < 1 || !(f > 0.0))
        return fail(h, SNOWTRI_E_ARG, "snowtri_smooth_create: bad sizes (max_persons=%d J=%d f=%g)", max_persons, J, f);
    CUDA_TRY(h, cudaSetDevice(h->device));
    snowtri_smooth_t* s = (snowtri_smooth_t*)calloc(1, sizeof(snowtri_smooth_t));
    if (!s) return fail(h, SNOWTRI_E_NOMEM, "snowtri_smooth_create: out of host memory");
    s->device = h->device; s->P = max_persons; s->J = J; s->f = f; s->z = z; s->r = r;
    const size_t bytes = (2 + (size_t)3 * max_persons * J * 3) * sizeof(double);
    cudaError_t e = cudaMalloc(&s->d_state, bytes);
    if (e == cudaSuccess) e = cudaMemset(s->d_state, 0, bytes);
    if (e != cudaSuccess) {
        if (s->d_state) cudaFree(s->d_state);
        free(s);
        return fail(h, SNOWTRI_E_CUDA, "snowtri_smooth_create: %s", cudaGetErrorString(e));
    }
    *out = s;
    return SNOWTRI_OK;
}

extern "C" int snowtri_smooth_destroy(snowtri_smooth_t* s) {
    if (!s) return SNOWTRI_OK;
    cudaSetDevice(s->device);
    if (s->d_state) cudaFree(s->d_state);
    if (s->d_work) cudaFree(s->d_work);
    if (s->d_cnt) cudaFree(s->d_cnt);
    if (s->d_apow) cudaFree(s->d_apow);
    free(s);
    return SNOWTRI_OK;
}

extern "C" int snowtri_smooth_set_chunked(snowtri_smooth_t* s, int enabled) {
    if (!s) return SNOWTRI_E_ARG;
    s->sequential = enabled ? 0 : 1;
    s->force_scan = enabled == 2 ? 1 : 0;
    return SNOWTRI_OK;
}

extern "C" int snowtri_smooth_reset(snowtri_t* h, snowtri_smooth_t* s, void* stream) {
    if (!h || !s) return fail(h, SNOWTRI_E_ARG, "snowtri_smooth_reset: NULL argument");
    CUDA_TRY(h, cudaSetDevice(s->device));
    CUDA_TRY(h, cudaMemsetAsync(s->d_state, 0, 2 * sizeof(double), (cudaStream_t)stream));
    return SNOWTRI_OK;
}

static int smooth_run(snowtri_t* h, snowtri_smooth_t* s, void* d_pts, bool f64, const int* d_nout, int* d_nsm, int F,
                      int Pout, int J, double delta_time, void* stream) {
    if (!h || !s) return fail(h, SNOWTRI_E_ARG, "snowtri_smooth_run: NULL handle");
    if (F == 0) return SNOWTRI_OK;
    if (!d_pts || !d_nout || !d_nsm || F < 0 || Pout < 1)
        return fail(h, SNOWTRI_E_ARG, "snowtri_smooth_run: bad argument");
    if (J != s->J) return fail(h, SNOWTRI_E_ARG, "snowtri_smooth_run: J=%d, state was created for %d joints", J, s->J);
    if (!(delta_time > 0.0)) return fail(h, SNOWTRI_E_ARG, "snowtri_smooth_run: delta_time must be positive");
    if (((uintptr_t)d_pts & (f64 ? 31u : 15u)) != 0) return fail(h, SNOWTRI_E_ARG, "snowtri_smooth_run: misaligned points");
    CUDA_TRY(h, cudaSetDevice(s->device));
    const double pi = 3.14159265358979323846;
    SmoothArgs a;
    a.pts = d_pts; a.nout = d_nout; a.nsm = d_nsm; a.state = s->d_state;
    a.F = F; a.Pout = Pout; a.P = s->P; a.J = J;
    a.T = delta_time; a.invT = 1.0 / delta_time;
    a.k1 = s->z / (pi * s->f);                                         // reference triangulation.py:7
    const double k2 = 1 / ((2 * pi * s->f) * (2 * pi * s->f));          // :8
    a.inv_k2 = 1.0 / k2;
    a.k3 = s->r * s->z / (2 * pi * s->f);                               // :9
    const int threads = s->P * J, grid = (threads + 127) / 128;
    cudaStream_t st = (cudaStream_t)stream;
    if (F <= 2 * kChunk || s->sequential) {  // short batch: one sequential launch
        if (f64) smooth_kernel<double4><<<grid, 128, 0, st>>>(a);
        else smooth_kernel<float4><<<grid, 128, 0, st>>>(a);
        smooth_finish_kernel<<<1, 1, 0, st>>>(s->d_state, d_nout, Pout, s->P);
        CUDA_TRY(h, cudaGetLastError());
        h->launches += 2;
        return SNOWTRI_OK;
    }
    // state matrix of one present frame, (xp, y, yd)' = A (xp, y, yd) + b x; its powers say after how many frames the
    // follower has forgotten its start state (smooth_overlap_kernel) and feed the chunk scan
    if (s->forget_T != delta_time) {
        const double T = delta_time, g = T / k2;
        const double A[9] = {0, 0, 0, 0, 1, T, -a.k3 / k2, -g, 1 - g * (T + a.k1)};
        double pw[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, nx[9];
        s->forget = 0;
        for (int n = 1; n <= kMaxForget && !s->forget; ++n) {
            double big = 0;
            for (int r_ = 0; r_ < 3; ++r_)
                for (int c_ = 0; c_ < 3; ++c_) {
                    double v = 0;
                    for (int m_ = 0; m_ < 3; ++m_) v += A[r_ * 3 + m_] * pw[m_ * 3 + c_];
                    nx[r_ * 3 + c_] = v;
                    big = fmax(big, fabs(v));
                }
            memcpy(pw, nx, sizeof(pw));
            if (big != big) break;            // absurd parameters: never
            if (big < 1e-19) s->forget = n;
        }
        s->forget_T = delta_time;
    }
    if (s->forget > 0 && !s->force_scan) {
        // one-pass path: a cooperative launch, so the chunk length is what lets every CTA be resident
        if (!s->coop_blocks) {
            int dev_coop = 0, per_sm_f = 0, per_sm_d = 0, sms = 0;
            CUDA_TRY(h, cudaDeviceGetAttribute(&dev_coop, cudaDevAttrCooperativeLaunch, s->device));
            CUDA_TRY(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
            CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_f, smooth_overlap_kernel<float4>, 128, 0));
            CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_d, smooth_overlap_kernel<double4>, 128, 0));
            const int per_sm = per_sm_f < per_sm_d ? per_sm_f : per_sm_d;
            s->coop_blocks = dev_coop && per_sm > 0 ? per_sm * sms : -1;
        }
        if (s->coop_blocks > 0) {
            int L = 256;   // frames per chunk: the batch is read 1 + forget/L times; shorter chunks = more threads in flight (131 072 frames x 133 joints: 256 or 320 frames 0.140 ms, 384 0.153, 512 0.168, 1024 0.232; profiles/r3e)
            if (const char* e = getenv("SNOWTRI_SMOOTH_CHUNK")) L = atoi(e) >= 32 ? atoi(e) : L;   // experiments
            while (L < 2 * s->forget) L *= 2;
            const long long max_threads = (long long)s->coop_blocks * 128;
            while (L < F && ((long long)(F + L - 1) / L) * threads > max_threads) L *= 2;   // (if even one chunk does not fit, the test below fails)
            int nchunks = (F + L - 1) / L;
            const long long g2 = ((long long)nchunks * threads + 127) / 128;
            if (g2 <= s->coop_blocks) {
                int W = s->forget;
                void* args[] = {(void*)&a, (void*)&L, (void*)&W, (void*)&nchunks};
                const void* fn = f64 ? (const void*)smooth_overlap_kernel<double4> : (const void*)smooth_overlap_kernel<float4>;
                CUDA_TRY(h, cudaLaunchCooperativeKernel(fn, dim3((unsigned)g2), dim3(128), args, 0, st));
                h->launches += 1;
                return SNOWTRI_OK;
            }
        }
    }
    // chunk-parallel path
    const int nchunks = (F + kChunk - 1) / kChunk;
    const size_t nch = (size_t)s->P * J;
    if (s->work_chunks < (size_t)nchunks) {
        if (s->d_work) cudaFree(s->d_work);
        if (s->d_cnt) cudaFree(s->d_cnt);
            s->d_work = nullptr; s->d_cnt = nullptr; s->work_chunks = 0;
        CUDA_TRY(h, cudaMalloc(&s->d_work, (size_t)nchunks * nch * 18 * sizeof(double)));
        CUDA_TRY(h, cudaMalloc(&s->d_cnt, (size_t)nchunks * (s->P + 1) * sizeof(int)));
        s->work_chunks = (size_t)nchunks;
    }
    if (!s->d_apow || s->apow_T != delta_time) {
        // state matrix of one present frame, (xp, y, yd)' = A (xp, y, yd) + b x, and its powers
        if (!s->d_apow) CUDA_TRY(h, cudaMalloc(&s->d_apow, (size_t)(kChunk + 1) * 9 * sizeof(double)));
        const double T = delta_time, g = T / k2;
        const double A[9] = {0, 0, 0, 0, 1, T, -a.k3 / k2, -g, 1 - g * (T + a.k1)};
        double* pw = s->apow_host;
        const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        memcpy(pw, I, sizeof(I));
        for (int n = 1; n <= kChunk; ++n)
            for (int r_ = 0; r_ < 3; ++r_)
                for (int c_ = 0; c_ < 3; ++c_) {
                    double v = 0;
                    for (int m_ = 0; m_ < 3; ++m_) v += A[r_ * 3 + m_] * pw[(n - 1) * 9 + m_ * 3 + c_];
                    pw[n * 9 + r_ * 3 + c_] = v;
                }
        CUDA_TRY(h, cudaMemcpyAsync(s->d_apow, pw, sizeof(s->apow_host), cudaMemcpyHostToDevice, st));
        CUDA_TRY(h, cudaStreamSynchronize(st));  // pageable source: make sure the copy has read it before it can change
        s->apow_T = delta_time;
    }
    ChunkArgs ca;
    ca.s = a; ca.work = s->d_work; ca.cnt = s->d_cnt; ca.apow = s->d_apow; ca.nchunks = nchunks;
    const unsigned g2 = (unsigned)(((long long)nchunks * threads + 127) / 128);
    if (f64) smooth_chunk_kernel<double4, 0><<<g2, 128, 0, st>>>(ca);
    else smooth_chunk_kernel<float4, 0><<<g2, 128, 0, st>>>(ca);
    smooth_carry_scan_kernel<<<threads * 3, kScanThreads, 0, st>>>(ca);
    if (f64) smooth_chunk_kernel<double4, 2><<<g2, 128, 0, st>>>(ca);
    else smooth_chunk_kernel<float4, 2><<<g2, 128, 0, st>>>(ca);
    smooth_finish_kernel<<<1, 1, 0, st>>>(s->d_state, d_nout, Pout, s->P);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 4;
    return SNOWTRI_OK;
}

extern "C" int snowtri_smooth_run(snowtri_t* h, snowtri_smooth_t* s, float* d_out, const int* d_nout, int* d_nsmooth,
                                  int F, int Pout, int J, double delta_time, void* stream) {
    return smooth_run(h, s, d_out, false, d_nout, d_nsmooth, F, Pout, J, delta_time, stream);
}

extern "C" int snowtri_smooth_run_f64(snowtri_t* h, snowtri_smooth_t* s, double* d_out, const int* d_nout,
                                      int* d_nsmooth, int F, int Pout, int J, double delta_time, void* stream) {
    return smooth_run(h, s, d_out, true, d_nout, d_nsmooth, F, Pout, J, delta_time, stream);
}


// ---- ragged ingestion ---------------------------------------------------------------------
// What main.py:50-55 does per frame and camera -- `for person, score in zip(keypoints, scores):
// add_human_2D_points(person, score, camera_index)` -- for a whole clip: the detector's outputs
// (N_fc, J, 2) / (N_fc, J) of every (frame, camera) are concatenated in (frame, camera, detector order) and
// described by CSR offsets; this kernel pads them into the dense batch layout of snowtri_run.
namespace snowtri {
__global__ void __launch_bounds__(256) pack_ragged_kernel(const float2* __restrict__ det_kpts, const float* __restrict__ det_scores,
                                                          const long long* __restrict__ offsets, int FC, int P, int J,
                                                          float2* __restrict__ kpts, float* __restrict__ scores,
                                                          int* __restrict__ counts) {
    // one warp per (frame, camera, person slot) row of J joints
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (long long)FC * P) return;
    const int fc = (int)(row / P), p = (int)(row - (long long)fc * P);
    const long long o0 = offsets[fc], n = offsets[fc + 1] - o0;
    if (p == 0 && lane == 0) counts[fc] = (int)(n < 0 ? 0 : (n > P ? P : n));  // persons beyond P slots are dropped
    const bool have = p < n;
    const float2* src = det_kpts + (o0 + p) * J;
    const float* ssrc = det_scores + (o0 + p) * J;
    float2* dst = kpts + row * J;
    float* sdst = scores + row * J;
    for (int j = lane; j < J; j += 32) {
        dst[j] = have ? src[j] : make_float2(0.f, 0.f);
        sdst[j] = have ? ssrc[j] : 0.f;
    }
}
}  // namespace snowtri

extern "C" int snowtri_pack_ragged(snowtri_t* h, const float* d_det_kpts, const float* d_det_scores,
                                   const long long* d_offsets, int F, int P, int J, float* d_kpts, float* d_scores,
                                   int* d_counts, void* stream) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_pack_ragged: NULL handle");
    if (F == 0) return SNOWTRI_OK;
    if (!d_offsets || !d_kpts || !d_scores || !d_counts || F < 0 || P < 1 || J < 1)
        return fail(h, SNOWTRI_E_ARG, "snowtri_pack_ragged: bad argument");
    if ((((uintptr_t)d_det_kpts | (uintptr_t)d_kpts) & 7u) != 0) return fail(h, SNOWTRI_E_ARG, "snowtri_pack_ragged: misaligned keypoints");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const long long rows = (long long)F * h->C * P;
    const long long blocks = (rows + 7) / 8;
    if (blocks > 2147483647LL) return fail(h, SNOWTRI_E_UNSUPPORTED, "snowtri_pack_ragged: batch too large");
    pack_ragged_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(d_det_kpts), d_det_scores, d_offsets, F * h->C, P, J,
        reinterpret_cast<float2*>(d_kpts), d_scores, d_counts);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return SNOWTRI_OK;
}
