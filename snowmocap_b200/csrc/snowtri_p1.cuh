// Single-person fused kernel (P == 1, C <= 8 cameras): the shipped configuration of the reference
// (configs/snowmocap_default_config.json: one performer, average_score_threshold 0, condense_score_tol 0)
// and BASELINE configs[0]/[1].  Same result as fused_kernel, different mapping onto the machine:
//
//   * warp-autonomous: every warp owns tiles of `Gw` consecutive frames and walks them alone; the
//     only synchronisation is __syncwarp(), so no CTA barrier ever stalls the FP pipes;
//   * lanes run over the flattened (frame, joint) index of the tile: 32 consecutive joints of one
//     camera row are one coalesced 256-byte read, and a tile of Gw*J items wastes < 32 lanes;
//     the (u,v,score) of the NEXT step are copied to shared memory with cp.async before the current
//     step is solved, so HBM latency is hidden by the ~700 instructions of a step and not by occupancy;
//   * with one person per camera a candidate IS a camera pair, so a cluster is a bit mask over the
//     C(C-1)/2 pairs.  The greedy clustering (reference triangulation.py:107-134) runs per frame in
//     one lane on bit masks; the fuse (reference :138-148) is a fully unrolled loop over the pairs
//     with camera and pair constants in the kernel-parameter constant bank.  When every lane of the
//     warp has the full clique (the normal case) the loop is branch-free;
//   * the rays of a joint are built once in registers and shared by all pairs;
//   * sum(score * point) is accumulated as  sum(w*mid) + 1/2 * sum_c alpha_c * h_c  with one scalar
//     alpha per camera instead of a 3-vector per pair (see snowtri_math.cuh for the algebra);
//   * the item loop carries no (frame, joint) bookkeeping: an item's frame is one multiply-high of its flat index,
//     and a tile whose frames all hold the full clique (one vote per tile) never looks at the masks;
//   * the person score (mean keypoint score, reference :150) is accumulated in registers per lane; whenever a step
//     contains the end of a frame -- a warp-uniform test, so a step without a frame end spends one add per item on
//     it -- the lanes' sums are added up by a fixed-order warp shuffle: deterministic, no atomics, no shared memory,
//     no re-read of the output.  (Reading the scores back at the end of a tile was tried: a warp's tile lives for
//     the whole launch, the rows are long gone from L2, +25 % DRAM reads.)
//
// Precision (template T = bulk arithmetic, TD = arithmetic of the ray-distance numerator d.(hm x hs)):
//   <double,double>  everything in float64, like the reference.
//   <float,double>   "mixed": rays, normal equations and the fuse in float32; the ill-conditioned
//                    distance numerator (two nearly intersecting rays) in float64.
//   <float,float>    everything in float32.
//   In both float modes every DISCRETE decision is float64: the clustering centres always, and the
//   distance gate (dist > dthr) is re-evaluated in float64 whenever the float32 value lies within
//   the guard band of the threshold (cold path p1_item_exact).
// Requirements (checked by the host): P == 1, all_kept (ast <= 0, kst >= 0) and never_filter
// (score_tol <= 0, kst >= 0); anything else takes fused_kernel.
#pragma once
#ifndef __CUDACC_RTC__
#include <math.h>
#endif

#include "snowtri_math.cuh"

namespace snowtri {

// relative half-width of the guard band around 1/dthr inside which the gate is re-decided in float64
constexpr double kGuardBandF32 = 4e-3;
constexpr double kGuardBandMixed = 2e-5;
constexpr unsigned kFullMask = 0xffffffffu;

// pair index of cameras x < y in the reference's (mc, sc) enumeration order
__host__ __device__ constexpr int pair_index(int C, int x, int y) { return x * C - x * (x + 1) / 2 + y - x - 1; }

// compile-time loop over the camera pairs x < y in the reference's order: f(IntC<x>, IntC<y>)
template <int V>
struct IntC {
    static constexpr int value = V;
};
template <bool V>
struct BoolC {
    static constexpr bool value = V;
};
template <int C, int X = 0, int Y = 1, typename F>
__device__ __forceinline__ void static_for_pairs(F&& f) {
    if constexpr (X < C - 1) {
        f(IntC<X>{}, IntC<Y>{});
        if constexpr (Y + 1 < C) static_for_pairs<C, X, Y + 1>(f);
        else static_for_pairs<C, X + 1, X + 2>(f);
    }
}

template <typename T, int C>
struct P1Args {
    static constexpr int NP = C * (C - 1) / 2;
    const float* kpts;    // (F,C,1,J,2)
    const float* scores;  // (F,C,1,J)
    const int* counts;    // (F,C) or null
    float* out;           // (F,Pout,Jout,4)
    float* pscores;       // (F,Pout)
    int* nout;            // (F)
    int F, J, Jout, Pout, Gw, center, num_tol;
    uint32_t jmagic;       // p1_div_magic(Jout)
    float kst_f;
    T inv_dthr;
    float guard_w;         // float modes: re-decide in float64 when |1/dist - 1/dthr| < guard_w (0 = never)
    T guard_lo, guard_hi;  // same band as an interval (cold path)
    double inv_dthr64, tol2;
    T kscale[NP + 1];                  // 0.0005 / n  (score unit and the division by the cluster size, Q6/Q10)
    alignas(16) T camc[C * 12];        // M = R*inv(K), row-major, rows padded to 4
    alignas(16) T pdc[NP * 8];         // per pair: d = ts - tm (3), pad, mid = (tm + ts)/2 (3), pad
    alignas(16) double cam64[C * 12];  // float64 copies: clustering centres, guard band, mixed mode
    alignas(16) double pd64[NP * 8];
    // mixed mode: d.(hm x hs) = [um vm 1] E [us vs 1]^T with E = -Mm^T [d]x Ms per pair (rows of 3, stride 10), composed
    // on the host in extended precision: the float64 distance numerator straight from the pixel coordinates (8 DFMA
    // per pair, no float64 rays)
    alignas(16) double E64[NP * 10];
    unsigned char px[NP], py[NP];
};

// Cold path: one (frame, slot, joint) item pair by pair in a rolled loop, with the distance gate of
// every pair inside the guard band decided in float64 from the raw pixel coordinates.  Also used
// for the second and later clusters of a frame.  `kp`/`sp` point at the item's camera-0 entry.
template <typename T, typename TD, int C>
__device__ __noinline__ float4 p1_item_exact(const P1Args<T, C>& a, const float2* __restrict__ kp,
                                             const float* __restrict__ sp, unsigned mask) {
    constexpr int NP = C * (C - 1) / 2;
    T S = (T)0, X = (T)0, Y = (T)0, Z = (T)0;
    for (int e = 0; e < NP; ++e) {
        if (!((mask >> e) & 1u)) continue;
        const int x = a.px[e], y = a.py[e];
        const float2 pm = kp[(size_t)x * a.J], ps = kp[(size_t)y * a.J];
        const float sm = sp[(size_t)x * a.J], ss = sp[(size_t)y * a.J];
        if (sm < a.kst_f || ss < a.kst_f) continue;
        const V3<T> hm = back_project4<T>(a.camc + 12 * x, (T)pm.x, (T)pm.y);
        const V3<T> hs = back_project4<T>(a.camc + 12 * y, (T)ps.x, (T)ps.y);
        const V3<double> hm64 = back_project4<double>(a.cam64 + 12 * x, (double)pm.x, (double)pm.y);
        const V3<double> hs64 = back_project4<double>(a.cam64 + 12 * y, (double)ps.x, (double)ps.y);
        V3<T> d;
        d.x = a.pdc[e * 8]; d.y = a.pdc[e * 8 + 1]; d.z = a.pdc[e * 8 + 2];
        V3<double> d64;
        d64.x = a.pd64[e * 8]; d64.y = a.pd64[e * 8 + 1]; d64.z = a.pd64[e * 8 + 2];
        const PairSolN<T> s = pair_solve_n(hm, dot3(hm, hm), hs, dot3(hs, hs), d);
        T dn;
        if constexpr (sizeof(TD) == 8) dn = (T)cross_dot(hm64, hs64, d64);
        else dn = cross_dot(hm, hs, d);
        const T r = rsqrt_fast(s.det * dn * dn);
        const T rd = r * s.det;
        bool far = rd < a.inv_dthr;
        if (sizeof(T) == 4 && rd > a.guard_lo && rd < a.guard_hi) {
            const PairSol<double> s64 = pair_solve(hm64, hs64, d64);
            far = rsqrt_fast(s64.qq) * s64.det < a.inv_dthr64;
        }
        if (far) continue;
        const T gq = ((T)sm + (T)ss) * r;
        const T w = gq * s.det;
        const V3<T> vv = pair_v(s.n0, s.n1, hm, hs);
        S += w;
        X = fma(w, a.pdc[e * 8 + 4], fma((T)0.5 * gq, vv.x, X));
        Y = fma(w, a.pdc[e * 8 + 5], fma((T)0.5 * gq, vv.y, Y));
        Z = fma(w, a.pdc[e * 8 + 6], fma((T)0.5 * gq, vv.z, Z));
    }
    if (S == (T)0) return make_float4(0.f, 0.f, 0.f, 0.f);
    const T rS = rcp_t(S);
    return make_float4((float)(X * rS), (float)(Y * rS), (float)(Z * rS), (float)(S * a.kscale[__popc(mask)]));
}

#ifndef P1_KEEP
#define P1_KEEP 1
#endif
#ifndef P1_E64_SMEM
#define P1_E64_SMEM 1   // mixed mode: float64 pair constants read from shared memory inside the loop
#endif
#ifndef P1_NI
#define P1_NI 1   // items a lane solves side by side per step (1 or 2)
#endif
// Camera and pair constants come from the kernel-parameter constant bank in the precompiled kernels.
#ifdef P1_JIT
// Rig-specialised build (NVRTC, see snowtri_jit.cu): the generated translation unit defines P1_JIT_CAMC / _PDC
// (brace lists of the camera and pair constants) and the scalar P1_JIT_* values before including this header, so
// they reach the instruction stream as immediates and the index arithmetic on J, Jout, Pout folds.
__device__ constexpr float kJ_camc[] = P1_JIT_CAMC;
__device__ constexpr float kJ_pdc[] = P1_JIT_PDC;
#define P1_CAMC(T, i) ((T)kJ_camc[i])
#define P1_PDC(T, i) ((T)kJ_pdc[i])
#else
#define P1_CAMC(T, i) (a.camc[i])
#define P1_PDC(T, i) (a.pdc[i])
#endif
// resident CTAs per SM the register allocation is held to (256-thread CTAs)
template <typename T, typename TD, int C>
constexpr int p1_min_blocks() {
#ifdef P1_MINB
    return P1_MINB;
#else
    return sizeof(T) == 4 ? (C <= 4 ? 2 : 1) : (C <= 3 ? 2 : 1);
#endif
}

// floor(q / d) for 0 <= q < 2^32 / d with magic = ceil(2^32 / d): one IMAD.HI instead of a running (frame, joint) pair
__host__ __device__ constexpr uint32_t p1_div_magic(int d) { return (uint32_t)((0x100000000ull + (uint32_t)d - 1u) / (uint32_t)d); }
constexpr int kP1MaxJout = 8192;  // 32 * Jout * Jout < 2^32 keeps the magic division exact over a tile

// Staging of the next step's inputs: every lane copies the (u, v) and the score of its own next items from global to
// shared memory with cp.async (LDGSTS: no destination registers, so nothing tempts the compiler to sink the loads
// behind the pair solves -- with register loads ptxas did exactly that and the step no longer covered the memory
// latency, profiles/r2h) and reads them back one step later.  A lane's record holds its C float2 and C scores; the
// record stride is 8 bytes times an odd number, which makes the 64-bit copies and the 64-bit reads conflict-free (a
// 48-byte stride with 128-bit reads made the copies 4-way conflicted: 15 M conflict cycles per launch, profiles/r2j).
__host__ __device__ constexpr int p1_rec_bytes(int C) {
    const int r = (C * 12 + 7) / 8 * 8;
    return (r / 8) % 2 ? r : r + 8;
}
__host__ __device__ constexpr int p1_stage_bytes(int C, int NI) { return NI * 32 * p1_rec_bytes(C); }
// per-warp shared memory: [centres | stage 1] [stage 0] [cluster masks].  Kept small on purpose: what the CTAs of an SM
// do not claim stays L1, and the copies in flight need their lines there -- 2 x 94 KB per SM left 60 KB of L1 and ran at
// 0.247 ms per 131 072 frames, 2 x 98 KB tipped the carve-out to 228 KB (28 KB of L1): 0.338 ms (profiles/r2j, r2k).
__host__ __device__ constexpr int p1_warp_bytes(int C, int NI, int Gw, int Pout) {
    const int NP = C * (C - 1) / 2;
    const int cen = Gw * NP * 24, stg = p1_stage_bytes(C, NI);
    return (((cen > stg ? cen : stg) + 15) / 16 * 16 + stg + ((Gw * Pout * 4 + 7) & ~7) + 15) / 16 * 16;
}
template <typename T, typename TD, int C, int NT>
__device__ __forceinline__ void p1_body(const P1Args<T, C>& a) {
    constexpr int NP = C * (C - 1) / 2;
    constexpr int NW = NT / 32;
    constexpr unsigned ALL = (NP >= 32) ? 0xffffffffu : ((1u << NP) - 1u);
    constexpr bool MIXED = sizeof(TD) != sizeof(T);
    constexpr int NI = P1_NI;
    constexpr int STEP = 32 * NI;
    constexpr int REC = p1_rec_bytes(C), STG = p1_stage_bytes(C, NI);
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef P1_JIT
    constexpr int J = P1_JIT_J, Jout = P1_JIT_JOUT, Pout = P1_JIT_POUT, Gw = P1_JIT_GW;
    constexpr float kst_f = P1_JIT_KST, guard_w = P1_JIT_GUARD_W;
    constexpr T inv_dthr = (T)P1_JIT_INV_DTHR, kscale_full = (T)P1_JIT_KSCALE_FULL;
    constexpr uint32_t jmagic = p1_div_magic(P1_JIT_JOUT);
#else
    const int J = a.J, Jout = a.Jout, Pout = a.Pout, Gw = a.Gw;
    const float kst_f = a.kst_f, guard_w = a.guard_w;
    const T inv_dthr = a.inv_dthr, kscale_full = a.kscale[NP];
    const uint32_t jmagic = a.jmagic;
#endif
    // per-warp scratch (p1_warp_bytes): the centres (Gw*NP*3 doubles) are dead once the tile is clustered, so the
    // second staging buffer lives on top of them
    unsigned char* wbase = smem + (size_t)warp * p1_warp_bytes(C, NI, Gw, Pout);
    const int cen_bytes = (max(Gw * NP * 24, STG) + 15) / 16 * 16;
    double* cen = reinterpret_cast<double*>(wbase);
    unsigned char* stage1 = wbase;
    unsigned char* stage0 = wbase + cen_bytes;
    uint32_t* meta = reinterpret_cast<uint32_t*>(stage0 + STG);
#if P1_E64_SMEM
    uint32_t e64s = 0;  // mixed mode: the pairs' float64 epipolar forms, one copy per CTA behind the warps' scratch
    if constexpr (MIXED) {
        double* e64 = reinterpret_cast<double*>(smem + (size_t)NW * p1_warp_bytes(C, NI, Gw, Pout));
        for (int i = threadIdx.x; i < NP * 10; i += NT) e64[i] = a.E64[i];
        __syncthreads();
        e64s = smem_u32(e64);
    }
#endif

    const float2* kp2 = reinterpret_cast<const float2*>(a.kpts);
    const int CJ = C * J;
    const int in_skip = CJ - Jout;            // item q = g*Jout + j reads input element g*CJ + j = q + g*in_skip
    const int out_skip = (Pout - 1) * Jout;   // ... and writes output row (g*Pout)*Jout + j = q + g*out_skip
    // Each warp owns one contiguous range of frames and walks it tile by tile: its input is two
    // sequential streams (kpts, scores).  (An explicit L2 look-ahead with cp.async.bulk.prefetch.L2 or per-lane
    // prefetch.global.L2 was measured and made the kernel 4-10 % slower; one step of look-ahead is enough.)
    const int gwarp = blockIdx.x * NW + warp, nwarps = gridDim.x * NW;
    const int fa = (int)((long long)a.F * gwarp / nwarps), fb = (int)((long long)a.F * (gwarp + 1) / nwarps);
    T m2[C][3];  // constant addends of the rays (third column of M): an FFMA takes one constant-bank operand only
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            m2[c][k] = P1_CAMC(T, 12 * c + 4 * k + 2);
#if P1_KEEP && !defined(P1_JIT)
            keep_in_register(m2[c][k]);
#endif
        }

    for (int f0 = fa; f0 < fb; f0 += Gw) {
        const int Gc = min(Gw, fb - f0);
        const int nitems = Gc * Jout;
        const int nsteps = (nitems + STEP - 1) / STEP;  // in the last one the lanes past the end redo the last item and store nothing
        const float2* kpt = kp2 + (size_t)f0 * CJ;
        const float* sct = a.scores + (size_t)f0 * CJ;
        const float2* kpc[C];  // per-camera rows of the tile: one 64-bit add per load in the item loop
        const float* scc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            kpc[c] = kpt + c * J;
            scc[c] = sct + c * J;
#if P1_KEEP && !defined(P1_JIT)
            keep_in_register(kpc[c]);
            keep_in_register(scc[c]);
#endif
        }
        float4* outt = reinterpret_cast<float4*>(a.out) + (size_t)f0 * Pout * Jout;

        // the inputs of the NI items of this lane in the step that starts at item q0: global -> this lane's records
        auto stage_in = [&](unsigned char* stage, uint32_t q0) {
            const uint32_t rec = smem_u32(stage) + lane * REC;
#pragma unroll
            for (int u = 0; u < NI; ++u) {
                const uint32_t q = min(q0 + 32 * u + lane, (uint32_t)nitems - 1u);
                const uint32_t off = q + __umulhi(q, jmagic) * (uint32_t)in_skip;  // unsigned: stays 32-bit arithmetic
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    cp_async8(rec + u * 32 * REC + c * 8, kpc[c] + off);
                    cp_async4(rec + u * 32 * REC + C * 8 + c * 4, scc[c] + off);
                }
            }
        };
        // ---- first step's inputs: in flight while the tile's clustering runs ---------------------------
        stage_in(stage0, 0u);

        // ---- centre-joint midpoint of every candidate, float64 (reference triangulation.py:112,124) ----
        for (int idx = lane; idx < Gc * NP; idx += 32) {
            const int gg = idx / NP, e = idx - gg * NP;
            const int x = a.px[e], y = a.py[e];
            const size_t fb = (size_t)gg * CJ + a.center;
            const float2 pm = __ldg(kpt + fb + (size_t)x * J), ps = __ldg(kpt + fb + (size_t)y * J);
            const V3<double> hm = back_project4<double>(a.cam64 + 12 * x, (double)pm.x, (double)pm.y);
            const V3<double> hs = back_project4<double>(a.cam64 + 12 * y, (double)ps.x, (double)ps.y);
            V3<double> d, mid;
            d.x = a.pd64[e * 8]; d.y = a.pd64[e * 8 + 1]; d.z = a.pd64[e * 8 + 2];
            mid.x = a.pd64[e * 8 + 4]; mid.y = a.pd64[e * 8 + 5]; mid.z = a.pd64[e * 8 + 6];
            const PairSol<double> s = pair_solve(hm, hs, d);
            const V3<double> w = pair_midpoint(s, hm, hs, mid);
            cen[3 * idx] = w.x;
            cen[3 * idx + 1] = w.y;
            cen[3 * idx + 2] = w.z;
        }
        __syncwarp();

        // ---- greedy clustering on pair bit masks, one lane per frame (reference :107-134) ----------
        int K = 0;
        unsigned mask0 = ALL;  // slot 0 of this lane's frame (lanes beyond the tile: neutral)
        if (lane < Gc) {
            unsigned present = (1u << C) - 1u;
            if (a.counts) {
                present = 0;
#pragma unroll
                for (int c = 0; c < C; ++c) present |= (a.counts[(size_t)(f0 + lane) * C + c] > 0 ? 1u : 0u) << c;
            }
            unsigned valid = 0;  // the reference's candidate list, in list order (bit e = pair e)
#pragma unroll
            for (int x = 0; x < C - 1; ++x)
#pragma unroll
                for (int y = x + 1; y < C; ++y)
                    if ((present >> x) & (present >> y) & 1u) valid |= 1u << pair_index(C, x, y);
            mask0 = 0u;
            if (valid) {
                const int last = 31 - __clz(valid);
                unsigned mains = valid & ~(1u << last);  // the last candidate is never a main (Q1/Q2)
                unsigned absorbed = 0;
                const double* cg = cen + 3 * lane * NP;
                while (mains) {
                    const int m = __ffs(mains) - 1;
                    mains &= mains - 1;
                    if ((absorbed >> m) & 1u) continue;
                    const double mx = cg[3 * m], my = cg[3 * m + 1], mz = cg[3 * m + 2];
                    unsigned mask = 1u << m;
                    unsigned rest = valid & ~absorbed & ~((2u << m) - 1u);
                    while (rest) {
                        const int i = __ffs(rest) - 1;
                        rest &= rest - 1;
                        const double dx = mx - cg[3 * i], dy = my - cg[3 * i + 1], dz = mz - cg[3 * i + 2];
                        if (!((dx * dx + dy * dy + dz * dz) > a.tol2)) {  // distance to the MAIN (Q3); NaN absorbs
                            mask |= 1u << i;
                            absorbed |= 1u << i;
                        }
                    }
                    if (__popc(mask) >= a.num_tol) {  // otherwise its members stay absorbed (Q5)
                        if (K < Pout) meta[lane * Pout + K] = mask;
                        if (K == 0) mask0 = mask;
                        ++K;
                    }
                }
            }
            for (int k = K; k < Pout; ++k) meta[lane * Pout + k] = 0u;
            a.nout[f0 + lane] = K;
        }
        const bool multi = __any_sync(kFullMask, K > 1) && Pout > 1;  // some frame has a second cluster
        const bool tile_full = __all_sync(kFullMask, mask0 == ALL);    // every frame's first cluster is the full clique
        __syncwarp();  // the centres are dead from here on: stage 1 may overwrite them

        // ---- fuse: lanes over the flattened (frame, joint) index of the tile ---------------------------
        // A step covers 32*NI consecutive items; a lane solves NI of them side by side (independent
        // dependency chains, constants fetched once per pair).  TF: the whole tile is known to be full cliques,
        // so the loop neither looks at the masks nor votes.
        // Person score (mean keypoint score, reference :150): a lane adds up the scores of its items; when a step
        // contains the end of a frame (a warp-uniform test, at most one per step when Jout >= 32*NI) the warp adds up
        // the lanes' sums in a fixed order and lane 0 stores the frame's score.  No shared memory, no atomics.
        float acc = 0.f;
        auto run_steps = [&](auto tfc) {
            constexpr bool TF = decltype(tfc)::value;
            for (int st = 0; st < nsteps; ++st) {
                const int qb = st * STEP + lane;  // this lane's first item of the step
                V3<T> h[NI][C];
                TD ud[NI][MIXED ? C : 1], vd[NI][MIXED ? C : 1];  // mixed: pixel coordinates in float64
                T A[NI][C], sc[NI][C];
                cp_async_wait_all();  // this lane's records of the step (written by its own copies)
                // next step's items of this lane: copies issued first thing (into the buffer that was last read one
                // step ago), read one step later -- a whole step of arithmetic covers the memory latency
                if (st + 1 < nsteps) stage_in((st & 1) ? stage0 : stage1, (uint32_t)(st + 1) * STEP);
                {
                    const unsigned char* rec = ((st & 1) ? stage1 : stage0) + lane * REC;
#pragma unroll
                    for (int u = 0; u < NI; ++u) {
                        float w[REC / 4];
#pragma unroll
                        for (int i = 0; i < (C * 12 + 7) / 8; ++i) {
                            const float2 v = *reinterpret_cast<const float2*>(rec + u * 32 * REC + i * 8);
                            w[2 * i] = v.x; w[2 * i + 1] = v.y;
                        }
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const T pu = (T)w[2 * c], pv = (T)w[2 * c + 1];
                            h[u][c].x = fma(P1_CAMC(T, 12 * c + 0), pu, fma(P1_CAMC(T, 12 * c + 1), pv, m2[c][0]));
                            h[u][c].y = fma(P1_CAMC(T, 12 * c + 4), pu, fma(P1_CAMC(T, 12 * c + 5), pv, m2[c][1]));
                            h[u][c].z = fma(P1_CAMC(T, 12 * c + 8), pu, fma(P1_CAMC(T, 12 * c + 9), pv, m2[c][2]));
                            if constexpr (MIXED) {
                                ud[u][c] = (TD)w[2 * c];
                                vd[u][c] = (TD)w[2 * c + 1];
                            }
                            A[u][c] = dot3(h[u][c], h[u][c]);
                            // a score below the keypoint threshold kills every pair of its camera: poison it so that
                            // max(sm + ss, 0) is 0 (scores that pass are >= kst >= 0 on this path)
                            sc[u][c] = w[2 * C + c] < kst_f ? (T)-1e30 : (T)w[2 * C + c];
                        }
                    }
                }
                // output slot 0: the unrolled path
                int gq_[NI];          // frame of the item inside the tile (only looked at when the masks are)
                unsigned mask[NI];
                bool full = true;
                if constexpr (!TF) {
                    bool allfull = true;
#pragma unroll
                    for (int u = 0; u < NI; ++u) {
                        gq_[u] = (int)__umulhi((uint32_t)min(qb + 32 * u, nitems - 1), jmagic);
                        mask[u] = meta[gq_[u] * Pout];
                        allfull = allfull && mask[u] == ALL;
                    }
                    full = __all_sync(kFullMask, allfull);
                } else {
#pragma unroll
                    for (int u = 0; u < NI; ++u) {
                        gq_[u] = 0;
                        mask[u] = ALL;
                    }
                }
                T S[NI], Xm[NI], Ym[NI], Zm[NI], al[NI][C];
                float margin[NI];  // float modes: smallest |1/dist - 1/dthr| over the pairs
#pragma unroll
                for (int u = 0; u < NI; ++u) {
                    S[u] = Xm[u] = Ym[u] = Zm[u] = (T)0;
                    margin[u] = INFINITY;
#pragma unroll
                    for (int c = 0; c < C; ++c) al[u][c] = (T)0;
                }
                auto pair = [&](auto xc, auto yc, auto uc) {
                    constexpr int x = decltype(xc)::value, y = decltype(yc)::value, u = decltype(uc)::value;
                    constexpr int e = pair_index(C, x, y);
                    V3<T> d;
                    d.x = P1_PDC(T, e * 8); d.y = P1_PDC(T, e * 8 + 1); d.z = P1_PDC(T, e * 8 + 2);
                    const PairSolN<T> s = pair_solve_n(h[u][x], A[u][x], h[u][y], A[u][y], d);
                    T dn;
                    if constexpr (MIXED) {  // l = E [uy vy 1]^T, d.n = [ux vx 1] l
#if P1_E64_SMEM
                        // the pair's nine float64 constants from the CTA's shared-memory copy, read where they are
                        // used: DFMA takes no constant-bank operand on this part, and left to itself the compiler hoists
                        // the 54 loads out of the loop and shuffles them between register files (profiles/r2l)
                        double E[10];
#pragma unroll
                        for (int i = 0; i < 5; ++i)
                            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(E[2 * i]), "=d"(E[2 * i + 1]) : "r"(e64s + (10 * e + 2 * i) * 8));
#else
                        const double* E = a.E64 + 10 * e;
#endif
                        const TD l0 = fma(E[0], ud[u][y], fma(E[1], vd[u][y], E[2]));
                        const TD l1 = fma(E[3], ud[u][y], fma(E[4], vd[u][y], E[5]));
                        const TD l2 = fma(E[6], ud[u][y], fma(E[7], vd[u][y], E[8]));
                        dn = (T)fma(l0, ud[u][x], fma(l1, vd[u][x], l2));
                    } else {
                        dn = cross_dot(h[u][x], h[u][y], d);
                    }
                    const T r = rsqrt_fast(s.det * dn * dn);  // q.q = det * (d.n)^2
                    const T rd = r * s.det;                   // 1/dist
                    if constexpr (sizeof(T) == 4) margin[u] = fminf(margin[u], fabsf(rd - inv_dthr));
                    T gq = fmax(sc[u][x] + sc[u][y], (T)0) * r;
                    if (rd < inv_dthr) gq = (T)0;  // dist > dthr (strict); NaN is not gated (Q8/Q9)
                    const T w = gq * s.det;
                    S[u] += w;
                    al[u][x] = fma(gq, s.n0, al[u][x]);
                    al[u][y] = fma(-gq, s.n1, al[u][y]);
                    Xm[u] = fma(w, P1_PDC(T, e * 8 + 4), Xm[u]);
                    Ym[u] = fma(w, P1_PDC(T, e * 8 + 5), Ym[u]);
                    Zm[u] = fma(w, P1_PDC(T, e * 8 + 6), Zm[u]);
                };
                if (full) {
                    static_for_pairs<C>([&](auto xc, auto yc) {
                        pair(xc, yc, IntC<0>{});
                        if constexpr (NI > 1) pair(xc, yc, IntC<NI - 1>{});
                    });
                } else {
                    static_for_pairs<C>([&](auto xc, auto yc) {
                        constexpr int e = pair_index(C, decltype(xc)::value, decltype(yc)::value);
                        if ((mask[0] >> e) & 1u) pair(xc, yc, IntC<0>{});
                        if constexpr (NI > 1)
                            if ((mask[NI - 1] >> e) & 1u) pair(xc, yc, IntC<NI - 1>{});
                    });
                }
                float ks[NI];  // keypoint scores of slot 0, for the person score
#pragma unroll
                for (int u = 0; u < NI; ++u) {
                    const bool live = qb + 32 * u < nitems;
                    const int q = min(qb + 32 * u, nitems - 1);
                    T X = (T)0, Y = (T)0, Z = (T)0;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        X = fma(al[u][c], h[u][c].x, X);
                        Y = fma(al[u][c], h[u][c].y, Y);
                        Z = fma(al[u][c], h[u][c].z, Z);
                    }
                    const bool some = S[u] != (T)0;  // S == 0 leaves (0,0,0) with score 0 (Q7)
                    const T rS = rcp_fast(some ? S[u] : (T)1);
                    float4 o;
                    o.x = some ? (float)(fma((T)0.5, X, Xm[u]) * rS) : 0.f;
                    o.y = some ? (float)(fma((T)0.5, Y, Ym[u]) * rS) : 0.f;
                    o.z = some ? (float)(fma((T)0.5, Z, Zm[u]) * rS) : 0.f;
                    o.w = (float)(S[u] * (full ? kscale_full : a.kscale[__popc(mask[u])]));
                    float4* orow;
                    if constexpr (TF) {
                        if (Pout == 1) orow = outt + q;
                        else orow = outt + (q + (int)__umulhi((uint32_t)q, jmagic) * out_skip);
                    } else {
                        orow = outt + (q + gq_[u] * out_skip);
                    }
                    if (live) *orow = o;
                    ks[u] = live ? o.w : 0.f;
                    if (sizeof(T) == 4 && margin[u] < guard_w && live) {
                        // a float32 distance within the guard band of dthr: decide in float64 and store again (rare)
                        const int off = q + (int)__umulhi((uint32_t)q, jmagic) * in_skip;
                        const float4 ex = p1_item_exact<T, TD, C>(a, kpt + off, sct + off, mask[u]);
                        *orow = ex;
                        ks[u] = ex.w;
                    }
                    if (Pout > 1 && live) {
                        // further clusters of the same frame are rare with one person per camera: rolled cold path
                        const int gg = (int)__umulhi((uint32_t)q, jmagic);
                        const int off = q + gg * in_skip;
                        for (int k = 1; k < Pout; ++k) {
                            const unsigned mk = multi ? meta[gg * Pout + k] : 0u;
                            orow[k * Jout] = mk ? p1_item_exact<T, TD, C>(a, kpt + off, sct + off, mk) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
                if (Jout >= STEP) {
                    // items [Q, Q + STEP) of the tile: does frame gA end inside?  (warp-uniform)
                    const uint32_t Q = (uint32_t)st * STEP, gA = __umulhi(Q, jmagic);
                    const int bnd = (int)((gA + 1u) * (uint32_t)Jout - Q);  // items of the step that still belong to frame gA
                    if (bnd > STEP) {
#pragma unroll
                        for (int u = 0; u < NI; ++u) acc += ks[u];
                    } else {
                        float next = 0.f;
#pragma unroll
                        for (int u = 0; u < NI; ++u) {
                            const bool old = lane + 32 * u < bnd;
                            acc += old ? ks[u] : 0.f;
                            next += old ? 0.f : ks[u];
                        }
                        const float tot = warp_sum(acc);
                        if (lane == 0) a.pscores[(size_t)(f0 + (int)gA) * Pout] = tot / (float)Jout;
                        acc = next;
                    }
                }
            }
        };
        if (tile_full) run_steps(BoolC<true>{});
        else run_steps(BoolC<false>{});
        __syncwarp();

        // ---- person scores not produced in the loop ------------------------------------------------------
        // Slot 0 when a step can hold several frame ends (Jout < 32*NI: short skeletons), and the rows of later
        // clusters when some frame of the tile has one: the keypoint scores were just written by this warp and a
        // tile of such rows is small, so they are read back from L2.  Fixed summation order.
        if (Jout < STEP || multi) {
            for (int row = 0; row < Gc * Pout; ++row) {
                if (!(row % Pout == 0 ? Jout < STEP : multi)) continue;
                const float4* o = outt + (size_t)row * Jout;
                float sum = 0.f;
                for (int jj = lane; jj < Jout; jj += 32) sum += __ldcg(&o[jj].w);
                sum = warp_sum(sum);
                if (lane == 0) a.pscores[(size_t)f0 * Pout + row] = sum / (float)Jout;
            }
        }
        if (!multi) {
            for (int row = lane; row < Gc * Pout; row += 32)
                if (row % Pout) a.pscores[(size_t)f0 * Pout + row] = 0.f;
        }
        __syncwarp();
    }
}

template <typename T, typename TD, int C, int NT>
__global__ void __launch_bounds__(NT, (p1_min_blocks<T, TD, C>())) p1_kernel(const __grid_constant__ P1Args<T, C> a) {
    p1_body<T, TD, C, NT>(a);
}

}  // namespace snowtri
