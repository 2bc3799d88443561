// Streaming general path, second generation of its two hot kernels (float32 bulk / float64 decisions, "mixed"):
//
//   gen_match_*_kernel   cross-camera person matching = the keep decision of every (camera pair, person pair)
//                        candidate (reference triangulation.py:56-87) AND the float64 centre-joint midpoint of the
//                        kept ones (:112,124) in one launch (was gen_keep_kernel + gen_centre_kernel).
//
// Why it is faster than gen_keep_kernel (profiles/r1e/gen_keep_cfg3_mixed_ncu_summary.txt: 93 warp-instructions per
// 32 joint evaluations, 5 % of them local-memory traffic):
//   * rays are built ONCE per frame -- (x, y, z, |h|^2) as one float4, the sign of |h|^2 carrying "score below the
//     keypoint threshold" -- into shared memory (frames that fit) or a scratch array (large rigs), instead of 6 FMAs
//     per ray per use;
//   * a warp owns a 4 x 4 tile of (main person, secondary person) candidates of one camera pair; a lane holds the
//     4 + 4 rays of its joint in registers and evaluates the 16 candidates from them: 0.5 ray loads per evaluation,
//     the per-main-ray part of the triple product (d x hm) hoisted out of the 4 secondary persons;
//   * the gate needs no cross product:  d.(hm x hs) = hs.(d x hm)  and  |hm x hs|^2 = |hm|^2 |hs|^2 - (hm.hs)^2,
//     10 FMA-class operations per evaluation;
//   * 2 x 16 running sums per lane in registers (static indices only: nothing in local memory), two butterfly
//     reductions per tile (2 x 16 shuffles) instead of two 5-step reductions per candidate;
//   * the decision, its float64 re-evaluation when the float32 sum is closer to the threshold than its error bound,
//     and the float64 centre of the kept candidates happen in the same warp, so the keep byte and the centre are the
//     only things written.
// Discrete decisions stay exact: same error-bound logic as gen_keep_item (kDistDelta, guard band on the gate), and
// anything that turns the float32 sum into NaN (non-finite inputs, parallel rays) falls into the float64 path.
// Requires dthr > 0 (the gate is tested on 1/dist against 1/dthr); the host keeps the first-generation kernels for
// dthr <= 0 and for the all-float64 mode.
#pragma once
#include "snowtri_general.cuh"

namespace snowtri {

constexpr int kTile = 4;  // persons per side of a register tile

// (x, y, z, +-|h|^2) of every ray of a chunk of frames: pre-pass of the large-rig variant
__global__ void __launch_bounds__(256) gen_rays_kernel(const __grid_constant__ GenArgs a, float4* __restrict__ rays) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* camM = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < a.C * 9; i += blockDim.x) camM[i] = (float)a.cam[(i / 9) * 12 + (i % 9)];
    __syncthreads();
    const size_t R = (size_t)a.C * a.P * a.J, n = (size_t)a.F * R;
    const float2* kf = reinterpret_cast<const float2*>(a.kpts);
    const int PJ = a.P * a.J;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)((i % R) / PJ);
        const float2 q = __ldg(kf + i);
        const float s = __ldg(a.scores + i);
        const V3<float> h = back_project<float>(camM + 9 * c, q.x, q.y);
        const float cc = dot3(h, h);
        rays[i] = make_float4(h.x, h.y, h.z, s < a.prm.kst_f ? -cc : cc);
    }
}

// After the call lane l holds the warp-wide sum of v[(l >> 1) & 15] (both lanes of a pair hold the same value).
// Halving butterfly: 8 + 4 + 2 + 1 + 1 shuffles instead of 16 x 5; fixed summation order.
__device__ __forceinline__ float reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1) {
        const bool hi = (lane & (2 * w)) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const float send = hi ? v[i] : v[i + w];
            const float keep = hi ? v[i + w] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, 2 * w);
        }
    }
    return v[0] + __shfl_xor_sync(kFull, v[0], 1);
}

// One (frame, camera pair, 4 x 4 person tile) item as a lane sees it (warp-uniform in the main loop; in the tail
// pass every group of lanes looks at another item).
// Joint-major staging of the rays of a CTA (MATCH_JM): element (joint j, camera slot b, person p) of `rays` lives at
// j * rw_r + b * ppad + p and of `scs` at j * rw_s + b * ppad + p, ppad = P rounded up to 4.  The 4 + 4 rays a lane needs
// for its joint are then two runs of consecutive float4 (constant offsets from one address per side instead of eight
// row addresses that depend on J: 38 -> 12 address instructions per pass) and the 4 scores of a side are one 128-bit
// load.  rw_r is odd and rw_s is 4 x odd, so that the 128-bit loads of consecutive joints (and the stores of the ray
// build, a thread per ray along j) fall into different banks.  rw_r == 0: row-major ([row][j]; the scratch-array kernel).
struct MatchLayout {
    int rw_r, rw_s, ppad;
};
__host__ __device__ inline MatchLayout match_layout_jm(int slots, int P) {
    MatchLayout l;
    l.ppad = (P + 3) & ~3;
    const int rw = slots * l.ppad;
    l.rw_r = rw | 1;
    l.rw_s = ((rw / 4) | 1) * 4;
    return l;
}
__host__ __device__ inline size_t match_jm_bytes(int slots, int P, int J) {
    const MatchLayout l = match_layout_jm(slots, P);
    return (size_t)J * ((size_t)l.rw_r * 16 + (size_t)l.rw_s * 4) + 16;
}

struct MatchItem {
    int pair, mc, sc, pm0, ps0, nm, ns;  // nm / ns: persons of the tile that are present
    int rw_r, rw_s;                      // joint-major strides (0: row-major)
    V3<float> d;                         // ts - tm
    const float4 *rm, *rs;               // rays of the tile's first main / secondary person
    const float *qm, *qs;                // their scores
};

// What an item is, independent of the frame: built once per CTA (MatchTables) instead of per (frame, item) -- the index
// divisions and the float64 centre differences were 6 % of the kernel's instructions (profiles/r2o).
struct MatchItemDesc {
    float dx, dy, dz;            // ts - tm
    unsigned short pair;
    unsigned char mc, sc, pm0, ps0;
    unsigned char pad[2];
};
__host__ __device__ inline size_t match_camf_bytes(int C) { return ((size_t)C * 36 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t match_desc_bytes(int items) { return ((size_t)items * sizeof(MatchItemDesc) + 15) & ~(size_t)15; }
__device__ __forceinline__ MatchItemDesc match_item_desc(const GenArgs& a, const double* camD, const uchar2* pairs, int it) {
    const int tpp = (a.P + kTile - 1) / kTile;
    MatchItemDesc q;
    const int pair = it / (tpp * tpp), tt = it - pair * tpp * tpp;
    q.pair = (unsigned short)pair;
    q.pm0 = (unsigned char)((tt / tpp) * kTile);
    q.ps0 = (unsigned char)((tt % tpp) * kTile);
    q.mc = pairs[pair].x;
    q.sc = pairs[pair].y;
    q.dx = (float)(camD[12 * q.sc + 9] - camD[12 * q.mc + 9]);
    q.dy = (float)(camD[12 * q.sc + 10] - camD[12 * q.mc + 10]);
    q.dz = (float)(camD[12 * q.sc + 11] - camD[12 * q.mc + 11]);
    q.pad[0] = q.pad[1] = 0;
    return q;
}

// `slot_m` / `slot_s`: which block of P rows of `rays` / `scs` holds the main / secondary camera (the camera index when the
// arrays hold the whole frame, 0 / 1 when a CTA staged just its camera pair)
__device__ __forceinline__ MatchItem match_item_of(const GenArgs& a, const MatchItemDesc& q, const float4* rays, const float* scs, int f,
                                                   int slot_m = -1, int slot_s = -1, MatchLayout lay = MatchLayout{0, 0, 0}) {
    const int C = a.C, P = a.P, J = a.J;
    MatchItem t;
    t.pair = q.pair;
    t.pm0 = q.pm0;
    t.ps0 = q.ps0;
    t.mc = q.mc;
    t.sc = q.sc;
    const int cm = a.counts ? max(0, min(P, a.counts[(size_t)f * C + t.mc])) : P;
    const int cs = a.counts ? max(0, min(P, a.counts[(size_t)f * C + t.sc])) : P;
    t.nm = max(0, min(kTile, cm - t.pm0));
    t.ns = max(0, min(kTile, cs - t.ps0));
    t.d.x = q.dx;
    t.d.y = q.dy;
    t.d.z = q.dz;
    const int bm = slot_m < 0 ? t.mc : slot_m, bs = slot_s < 0 ? t.sc : slot_s;
    t.rw_r = lay.rw_r;
    t.rw_s = lay.rw_s;
    if (lay.rw_r) {
        t.rm = rays + bm * lay.ppad + t.pm0;
        t.rs = rays + bs * lay.ppad + t.ps0;
        t.qm = scs + bm * lay.ppad + t.pm0;
        t.qs = scs + bs * lay.ppad + t.ps0;
    } else {
        t.rm = rays + (size_t)(bm * P + t.pm0) * J;
        t.rs = rays + (size_t)(bs * P + t.ps0) * J;
        t.qm = scs + (size_t)(bm * P + t.pm0) * J;
        t.qs = scs + (size_t)(bs * P + t.ps0) * J;
    }
    return t;
}

// Gate on 1/dist, pulled in / pushed out by the guard band: a joint surely passes at 1/dist >= r_sure, may pass at
// 1/dist >= r_maybe.  The band is the float32 error of the distance (kGateGuard, relative) or the distance error bound
// relative to a small threshold, whichever is wider.
struct MatchGate {
    float r_sure, r_maybe;
    __device__ explicit MatchGate(double dthr) {
        const float guard = fmaxf(kGateGuard, 4.f * kDistDelta / (float)dthr);
        r_sure = (float)(1.0 / dthr) * (1.f + guard);
        r_maybe = (float)(1.0 / dthr) * (1.f - guard);
    }
};

// The 16 candidates of the tile at joint j of this lane: two running sums per candidate, a LOWER and an UPPER bound
// of its float64 score sum (in units of 0.0005).  A joint's score is c/dist and the float32 ray distance is good to
// kDistDelta metres, so with t = kDistDelta/dist the true score lies in [w/(1+t), w/(1-t)], which contains
// [w(1-t), w(1+2t)] for t <= 1/2 (beyond that the joint adds nothing to the lower bound and +inf to the upper).  A
// joint whose gate sits inside its guard band counts in the upper bound only.  One-sided bounds matter: scores are
// heavy-tailed (rays that happen to pass within 0.01 mm score 80 +- 80), which blows up a symmetric error bar -- with
// it every fourth correctly matched candidate went to the float64 path (profiles/r2a) -- but not the lower bound, and
// "kept" only needs the lower bound above the threshold.
// Straight-line and predicated: a vote around the score arithmetic does not pay -- with 32 joints per warp some lane
// of a wrongly matched pair passes the 5 cm gate two times out of three (the voted branch ran for 65 % of the
// evaluations and serialised their dependency chains, profiles/r2b).
__device__ __forceinline__ void match_eval(const MatchItem& t, int J, int j, bool valid, const MatchGate& g,
                                           float (&lo)[kTile * kTile], float (&hi)[kTile * kTile]) {
    float4 m[kTile], s[kTile];
    float sm[kTile], ss[kTile], lim_sure[kTile], lim_maybe[kTile];
#pragma unroll
    for (int i = 0; i < kTile; ++i) {  // persons beyond the count alias person 0 of the tile and are masked
        const int r = i < t.nm ? i : 0;
        m[i] = t.rm[(size_t)r * J + j];
        sm[i] = t.qm[(size_t)r * J + j];
    }
#pragma unroll
    for (int k = 0; k < kTile; ++k) {
        const int r = k < t.ns ? k : 0;
        s[k] = t.rs[(size_t)r * J + j];
        ss[k] = t.qs[(size_t)r * J + j];
        // a secondary ray with a low score (it carries -|hs|^2) or beyond the count is gated through its limits
        const bool oks = k < t.ns && !(s[k].w < 0.f);
        lim_sure[k] = oks ? g.r_sure : INFINITY;
        lim_maybe[k] = oks ? g.r_maybe : INFINITY;
        s[k].w = fabsf(s[k].w);
    }
#pragma unroll
    for (int i = 0; i < kTile; ++i) {
        const bool okm = valid && i < t.nm && !(m[i].w < 0.f);
        V3<float> e;  // d x hm:  d.(hm x hs) = hs.(d x hm)
        e.x = fmaf(t.d.y, m[i].z, -(t.d.z * m[i].y));
        e.y = fmaf(t.d.z, m[i].x, -(t.d.x * m[i].z));
        e.z = fmaf(t.d.x, m[i].y, -(t.d.y * m[i].x));
#pragma unroll
        for (int k = 0; k < kTile; ++k) {
            const float B = fmaf(m[i].x, s[k].x, fmaf(m[i].y, s[k].y, m[i].z * s[k].z));
            const float dn = fmaf(e.x, s[k].x, fmaf(e.y, s[k].y, e.z * s[k].z));
            const float nn = fmaf(m[i].w, s[k].w, -(B * B));   // |hm x hs|^2
            const float rd = nn * rsqrt_fast(nn * (dn * dn));  // sqrt(n.n)/|d.n| = 1/dist
            const bool sure = okm && !(rd < lim_sure[k]);      // dist > dthr is gated (strict); NaN is not (Q8/Q9)
            const bool maybe = okm && !(rd < lim_maybe[k]);
            const float w = (sm[i] + ss[k]) * rd;              // score / 0.0005
            const float tt = kDistDelta * rd;                  // relative half-width of the distance error
            const bool tight = rd <= 0.5f / kDistDelta;
            if (sure && tight) lo[i * kTile + k] += fmaf(-w, tt, w);                  // w (1 - t)
            if (maybe) hi[i * kTile + k] += tight ? fmaf(w + w, tt, w) : INFINITY;    // w (1 + 2t)
        }
    }
}

// Decision of the tile's 16 candidates from the reduced bounds (lanes 2c and 2c+1 hold candidate c = (i, k) of the
// tile), float64 re-evaluation where the bounds straddle the threshold, keep byte of every candidate slot of the tile
// and float64 centre-joint midpoint of the kept ones (reference :112,124).  `t` is warp-uniform here.
__device__ __forceinline__ void match_decide(const GenArgs& a, const double* camD, const float2* kf, const float* sf, int f,
                                             const MatchItem& t, bool have_sums, float lo_tot, float hi_tot, int lane) {
    const int P = a.P, J = a.J;
    const int ci = (lane >> 3) & 3, ck = (lane >> 1) & 3;
    const bool present = ci < t.nm && ck < t.ns;
    bool kept = present;  // ast <= 0 with kst >= 0 can never reject (all_kept): no sums were needed
    if (have_sums) {
        // mean < ast  <=>  sum < ast*J.  4e-5: float32 rounding of the weights (|hm x hs|^2 from the Gram form) and
        // of the 133-term sums, relative.  Anything else (also NaN) is decided in float64 from the raw inputs.
        const double thrJ = a.prm.ast * (double)J;
        const bool surely_kept = (double)(0.0005f * lo_tot) * (1.0 - 4e-5) >= thrJ;
        const bool surely_not = (double)(0.0005f * hi_tot) * (1.0 + 4e-5) < thrJ;
        kept = present && surely_kept;
        unsigned redo = __ballot_sync(kFull, present && !(lane & 1) && !surely_kept && !surely_not);
        while (redo) {  // discrete decision: the whole warp redoes this candidate in float64
            const int l = __ffs(redo) - 1;
            redo &= redo - 1;
            const int cc = l >> 1;
            const double mean = gen_candidate_mean_f64(a, camD, kf, sf, t.mc, t.pm0 + (cc >> 2), t.sc, t.ps0 + (cc & 3), lane);
            if ((lane >> 1) == cc) kept = !(mean < a.prm.ast);  // NaN mean is kept (Q9)
        }
    }
    const int pm = t.pm0 + ci, ps = t.ps0 + ck;
    if (!(lane & 1) && pm < P && ps < P) {
        const size_t n = (size_t)f * a.ncand + ((size_t)t.pair * P + pm) * P + ps;
        a.keep[n] = kept ? 1 : 0;
        if (kept) {
            const float2 q0 = kf[(size_t)(t.mc * P + pm) * J + a.prm.center], q1 = kf[(size_t)(t.sc * P + ps) * J + a.prm.center];
            const double* c0 = camD + 12 * t.mc;
            const double* c1 = camD + 12 * t.sc;
            const V3<double> h0 = back_project<double>(c0, (double)q0.x, (double)q0.y);
            const V3<double> h1 = back_project<double>(c1, (double)q1.x, (double)q1.y);
            V3<double> dd, mid;
            dd.x = c1[9] - c0[9]; dd.y = c1[10] - c0[10]; dd.z = c1[11] - c0[11];
            mid.x = (c0[9] + c1[9]) / 2; mid.y = (c0[10] + c1[10]) / 2; mid.z = (c0[11] + c1[11]) / 2;
            const PairSol<double> sol = pair_solve(h0, h1, dd);
            const V3<double> w = pair_midpoint(sol, h0, h1, mid);
            double* c3 = a.cen + n * 3;
            c3[0] = w.x;
            c3[1] = w.y;
            c3[2] = w.z;
        }
    }
}

// Second form of the same bounds (MATCH_SUMS, the default): with t = kDistDelta/dist the two bounds are
//   lower = sum w (1 - t) = S1 - kDistDelta S2,   upper = sum w (1 + 2t) = S1 + 2 kDistDelta S2 + band,
//   S1 = sum w,  S2 = sum w/dist   over the joints that surely pass the gate with t <= 1/2,
// so a candidate needs two running sums and no per-joint selects (27 -> 23 instructions per evaluation).  What the
// sums leave out:
//   * a joint whose gate sits inside its guard band adds at most kappa (sm + ss) to the upper bound, kappa =
//     r_sure (1 + 2 kDistDelta r_sure): the lane adds up sm + ss of such joints (`bandq`, all 16 candidates of the tile
//     together; scores that pass the keypoint threshold are >= 0 on this path) and the tile's upper bounds all get
//     kappa * sum over lanes of bandq -- a fraction of a joint's score, against thresholds of J joint scores;
//   * a joint with t > 1/2 (rays that pass within 0.02 mm) has no upper bound: it sets the candidate's bit in `wild`
//     (one register per lane for the 16 candidates) and the candidate's upper bound is +inf.
// FULL: all 4 + 4 persons of the tile are present (warp-uniform; the usual case) -- no row selects, no count tests.
template <bool FULL, bool JM>
__device__ __forceinline__ void match_eval_sums(const MatchItem& t, int J, int j, bool valid, const MatchGate& g,
                                                float (&s1)[kTile * kTile], float (&s2)[kTile * kTile], unsigned& wild,
                                                float& bandq) {
    float4 m[kTile], s[kTile];
    float sm[kTile], ss[kTile], lim_sure[kTile], lim_maybe[kTile];
    const float4 *pm = t.rm, *ps = t.rs;
    const float *qm = t.qm, *qs = t.qs;
    if constexpr (JM) {
        pm += j * t.rw_r; ps += j * t.rw_r;
        qm += j * t.rw_s; qs += j * t.rw_s;
        if constexpr (FULL) {
            const float4 a4 = *reinterpret_cast<const float4*>(qm), b4 = *reinterpret_cast<const float4*>(qs);
            sm[0] = a4.x; sm[1] = a4.y; sm[2] = a4.z; sm[3] = a4.w;
            ss[0] = b4.x; ss[1] = b4.y; ss[2] = b4.z; ss[3] = b4.w;
        }
    } else {
        pm += j; ps += j;
        qm += j; qs += j;
    }
    const int rstride = JM ? 1 : J;
#pragma unroll
    for (int i = 0; i < kTile; ++i) {  // persons beyond the count alias person 0 of the tile and are masked
        const int r = (FULL || i < t.nm) ? i : 0;
        m[i] = pm[(size_t)r * rstride];
        if constexpr (!(JM && FULL)) sm[i] = qm[(size_t)r * rstride];
    }
#pragma unroll
    for (int k = 0; k < kTile; ++k) {
        const int r = (FULL || k < t.ns) ? k : 0;
        s[k] = ps[(size_t)r * rstride];
        if constexpr (!(JM && FULL)) ss[k] = qs[(size_t)r * rstride];
        // a secondary ray with a low score (it carries -|hs|^2) or beyond the count is gated through its limits
        const bool oks = (FULL || k < t.ns) && !(s[k].w < 0.f);
        lim_sure[k] = oks ? g.r_sure : INFINITY;
        lim_maybe[k] = oks ? g.r_maybe : INFINITY;
        s[k].w = fabsf(s[k].w);
    }
    constexpr float kTight = 0.5f / kDistDelta;
#pragma unroll
    for (int i = 0; i < kTile; ++i) {
        // a main ray that does not count (idle lane, person beyond the count, score below the keypoint threshold) gets
        // 1/dist = -inf: it fails every test below without a predicate of its own
        const float pen = (valid && (FULL || i < t.nm) && !(m[i].w < 0.f)) ? 0.f : INFINITY;
        m[i].w = fabsf(m[i].w);
        V3<float> e;  // d x hm:  d.(hm x hs) = hs.(d x hm)
        e.x = fmaf(t.d.y, m[i].z, -(t.d.z * m[i].y));
        e.y = fmaf(t.d.z, m[i].x, -(t.d.x * m[i].z));
        e.z = fmaf(t.d.x, m[i].y, -(t.d.y * m[i].x));
#pragma unroll
        for (int k = 0; k < kTile; ++k) {
            const float B = fmaf(m[i].x, s[k].x, fmaf(m[i].y, s[k].y, m[i].z * s[k].z));
            const float dn = fmaf(e.x, s[k].x, fmaf(e.y, s[k].y, e.z * s[k].z));
            const float nn = fmaf(m[i].w, s[k].w, -(B * B));   // |hm x hs|^2
            const float rd = fmaf(nn, rsqrt_fast(nn * (dn * dn)), -pen);  // sqrt(n.n)/|d.n| = 1/dist
            const float q = sm[i] + ss[k];
            const float w = q * rd;                            // score / 0.0005
            // dist > dthr is gated (strict); a NaN distance is not (Q8/Q9): it passes the tests and poisons the sums.
            //   pass = !(rd < lim_sure);  sure = pass && !(rd > kTight);  wild = pass && rd > kTight;
            //   band = !pass && !(rd < lim_maybe)
            // Predicated straight-line code: written in C++ the compiler branches around the tests and every evaluation
            // becomes its own basic block.
            asm("{\n\t.reg .pred pp, pq, ps, pw, pb;\n\t"
                "setp.geu.f32 pp|pq, %4, %5;\n\t"
                "setp.leu.and.f32 ps|pw, %4, %7, pp;\n\t"
                "setp.geu.and.f32 pb, %4, %6, pq;\n\t"
                "@ps add.f32 %0, %0, %8;\n\t"
                "@ps fma.rn.f32 %1, %8, %4, %1;\n\t"
                "@pw or.b32 %2, %2, %9;\n\t"
                "@pb add.f32 %3, %3, %10;\n\t}"
                : "+f"(s1[i * kTile + k]), "+f"(s2[i * kTile + k]), "+r"(wild), "+f"(bandq)
                : "f"(rd), "f"(lim_sure[k]), "f"(lim_maybe[k]), "f"(kTight), "f"(w), "r"(1u << (i * kTile + k)), "f"(q));
        }
    }
}

#ifndef MATCH_SUMS
#define MATCH_SUMS 1
#endif
#ifndef MATCH_JM
#define MATCH_JM MATCH_SUMS   // joint-major staging needs the second form of the evaluation
#endif
template <bool V>
struct MBool {
    static constexpr bool value = V;
};

// One item by one warp, joints in chunks of 32 (the last one partly idle: 5 of 32 lanes at 133 joints).  Tried and
// dropped: the leftover joints of four items sharing one pass -- 11 % fewer instructions, but either the unrolled form
// outgrows the instruction cache (1.33 -> 1.47 ms at BASELINE configs[2]) or the rolled form pays the saving back in
// index arithmetic and needs more than the 128 registers two 7-warp CTAs per SM leave (2.03 ms); profiles/r2d, r2e.
// Also dropped: the leftover joints split over the lanes (32 / tail lanes per joint, 3 evaluations per lane): 4.8 %
// fewer instructions, but the short dependent tail issues at 70.6 % instead of 74.8 % -- 1.089 vs 1.079 ms, profiles/r2t.
template <bool JM>
__device__ __forceinline__ void gen_match_item(const GenArgs& a, const double* camD, const MatchItemDesc& q, const float4* rays,
                                               const float* scs, const float2* kf, const float* sf, int f, int lane,
                                               int slot_m = -1, int slot_s = -1, MatchLayout lay = MatchLayout{0, 0, 0}) {
    const MatchItem t = match_item_of(a, q, rays, scs, f, slot_m, slot_s, lay);
    const bool sums = !a.all_kept && t.nm > 0 && t.ns > 0;
    float lo_tot = 0.f, hi_tot = 0.f;
    if (sums) {
        const MatchGate g(a.prm.dthr);
        float lo[kTile * kTile], hi[kTile * kTile];
#pragma unroll
        for (int i = 0; i < kTile * kTile; ++i) lo[i] = hi[i] = 0.f;
#if MATCH_SUMS
        unsigned wild = 0u;
        float bandq = 0.f;
        auto passes = [&](auto fullc, auto jmc) {
            for (int j0 = 0; j0 < a.J; j0 += 32) {
                const bool valid = j0 + lane < a.J;
                match_eval_sums<decltype(fullc)::value, decltype(jmc)::value>(t, a.J, valid ? j0 + lane : a.J - 1, valid, g, lo, hi, wild, bandq);
            }
        };
        const bool full = t.nm == kTile && t.ns == kTile;
        if (full) passes(MBool<true>{}, MBool<JM>{});
        else passes(MBool<false>{}, MBool<JM>{});
        const float S1 = reduce16(lo, lane), S2 = reduce16(hi, lane);
        wild = __reduce_or_sync(kFull, wild);
        // guard-band joints of the whole tile, charged to every candidate; 1.001: float32 rounding of this small term
        // (a threshold below 2 kDistDelta puts the band itself beyond t = 1/2: no bound)
        const float bsum = warp_sum(bandq);
        const float kappa = g.r_sure <= 0.5f / kDistDelta ? g.r_sure * fmaf(2.f * kDistDelta, g.r_sure, 1.f) * 1.001f : INFINITY;
        const float band = bsum > 0.f ? bsum * kappa : 0.f;
        lo_tot = fmaf(-kDistDelta, S2, S1);
        hi_tot = ((wild >> ((lane >> 1) & 15)) & 1u) ? INFINITY : fmaf(2.f * kDistDelta, S2, S1) + band;
#else
        for (int j0 = 0; j0 < a.J; j0 += 32) {
            const bool valid = j0 + lane < a.J;
            match_eval(t, a.J, valid ? j0 + lane : a.J - 1, valid, g, lo, hi);
        }
        lo_tot = reduce16(lo, lane);
        hi_tot = reduce16(hi, lane);
#endif
    }
    match_decide(a, camD, kf, sf, f, t, sums, lo_tot, hi_tot, lane);
}

// Shared-memory tables of the match kernels: camera matrices and centres in float64 (C*12), the pair table.
struct MatchTables {
    const double* camD;
    const uchar2* pairs;
    __device__ MatchTables(unsigned char* smem, const GenArgs& a) {
        double* cd = reinterpret_cast<double*>(smem);
        uchar2* pr = reinterpret_cast<uchar2*>(cd + a.C * 12);
        camD = cd;
        pairs = pr;
        for (int i = threadIdx.x; i < a.C * 12; i += blockDim.x) cd[i] = a.cam[i];
        for (int p = threadIdx.x; p < a.npairs; p += blockDim.x) {
            int mc, sc;
            decode_pair(p, a.C, mc, sc);
            pr[p] = make_uchar2((unsigned char)mc, (unsigned char)sc);
        }
    }
    __host__ __device__ static size_t bytes(int C, int npairs) { return (((size_t)C * 96 + (size_t)npairs * 2) + 15) & ~(size_t)15; }
};

// Frames whose rays fit in shared memory (20 bytes per ray): one CTA per frame builds them once and its warps walk the
// (camera pair, tile) items out of shared memory.  Two CTAs per SM: one CTA's ray build (global loads) and decisions
// overlap the other's arithmetic.
__global__ void __launch_bounds__(256, 2) gen_match_smem_kernel(const __grid_constant__ GenArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    MatchTables tb(smem, a);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
    const int f = blockIdx.x, C = a.C, P = a.P, J = a.J;
    const int R = C * P * J;
    const int tpp = (P + kTile - 1) / kTile;
    const int items = a.npairs * tpp * tpp;
    MatchItemDesc* idesc = reinterpret_cast<MatchItemDesc*>(smem + MatchTables::bytes(C, a.npairs));
    float* camF = reinterpret_cast<float*>(smem + MatchTables::bytes(C, a.npairs) + match_desc_bytes(items));  // M of every camera, float32
    float4* rays = reinterpret_cast<float4*>(smem + MatchTables::bytes(C, a.npairs) + match_desc_bytes(items) + match_camf_bytes(C));
#if MATCH_JM
    const MatchLayout lay = match_layout_jm(C, P);
    float* scs = reinterpret_cast<float*>(rays + (size_t)J * lay.rw_r);
#else
    const MatchLayout lay{0, 0, 0};
    float* scs = reinterpret_cast<float*>(rays + R);
#endif
    const float2* kf = reinterpret_cast<const float2*>(a.kpts) + (size_t)f * R;
    const float* sf = a.scores + (size_t)f * R;
    for (int i = threadIdx.x; i < C * 9; i += blockDim.x) camF[i] = (float)a.cam[(i / 9) * 12 + (i % 9)];
    __syncthreads();  // the tables
    for (int it = threadIdx.x; it < items; it += blockDim.x) idesc[it] = match_item_desc(a, tb.camD, tb.pairs, it);
    // Ray build, a thread per ray: the loads of a batch of rays are issued back to back before any of them is used --
    // a warp per (camera, person) row with a load-use-store loop over the joints was ~20 dependent memory round trips per
    // warp and frame, during which the CTA did nothing else (15 % of the stall samples on 2 % of the instructions,
    // profiles/r2r).
    {
        constexpr int NB = 8;
        const uint32_t pj_magic = (uint32_t)((0x100000000ull + (uint32_t)(P * J) - 1u) / (uint32_t)(P * J));  // exact: C (P J)^2 < 2^32 here
        const uint32_t j_magic = (uint32_t)((0x100000000ull + (uint32_t)J - 1u) / (uint32_t)J);
        for (int i0 = threadIdx.x; i0 < R; i0 += NB * blockDim.x) {
            float2 q[NB];
            float sv[NB];
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int i = min(i0 + u * (int)blockDim.x, R - 1);
                q[u] = __ldg(kf + i);
                sv[u] = __ldg(sf + i);
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int i = i0 + u * (int)blockDim.x;
                if (i < R) {
                    const int c = P * J > 1 ? (int)__umulhi((uint32_t)i, pj_magic) : i;   // (ceil(2^32 / 1) does not fit)
                    const V3<float> h = back_project<float>(camF + 9 * c, q[u].x, q[u].y);
                    const float cc = dot3(h, h);
                    int dr = i, ds = i;
                    if (lay.rw_r) {
                        const int rem = i - c * P * J, pp = J > 1 ? (int)__umulhi((uint32_t)rem, j_magic) : rem, jj = rem - pp * J;
                        dr = jj * lay.rw_r + c * lay.ppad + pp;
                        ds = jj * lay.rw_s + c * lay.ppad + pp;
                    }
                    rays[dr] = make_float4(h.x, h.y, h.z, sv[u] < a.prm.kst_f ? -cc : cc);
                    scs[ds] = sv[u];
                }
            }
        }
    }
    __syncthreads();
    for (int it = warp; it < items; it += NW) gen_match_item<(MATCH_JM != 0)>(a, tb.camD, idesc[it], rays, scs, kf, sf, f, lane, -1, -1, lay);
}

// Large rigs, where a frame's rays do not fit in shared memory but the 2 P rows of ONE camera pair do (20 bytes per ray:
// 43 KB at 8 persons x 133 joints, 85 KB at 16): one CTA per (frame, camera pair) builds the rays of its two cameras from
// the raw inputs (L2-resident while the frame's C(C-1)/2 CTAs run: consecutive block indices) and its warps walk the
// pair's person tiles out of shared memory.  Against gen_match_global_kernel (rays from a scratch array in L2 at every
// pass): 4.87 -> see DESIGN.md ns per item at BASELINE configs[3].
__global__ void __launch_bounds__(256, 2) gen_match_pair_kernel(const __grid_constant__ GenArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    MatchTables tb(smem, a);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
    const int C = a.C, P = a.P, J = a.J, PJ = P * J;
    const int f = blockIdx.x / a.npairs, pair = blockIdx.x - f * a.npairs;
    const size_t R = (size_t)C * PJ;
    float* camF = reinterpret_cast<float*>(smem + MatchTables::bytes(C, a.npairs));  // M of the two cameras, float32
    float4* rays = reinterpret_cast<float4*>(smem + MatchTables::bytes(C, a.npairs) + 80);
#if MATCH_JM
    const MatchLayout lay = match_layout_jm(2, P);
    float* scs = reinterpret_cast<float*>(rays + (size_t)J * lay.rw_r);
#else
    const MatchLayout lay{0, 0, 0};
    float* scs = reinterpret_cast<float*>(rays + 2 * PJ);
#endif
    const uint32_t j_magic = (uint32_t)((0x100000000ull + (uint32_t)J - 1u) / (uint32_t)J);
    const float2* kf = reinterpret_cast<const float2*>(a.kpts) + (size_t)f * R;
    const float* sf = a.scores + (size_t)f * R;
    int mc, sc;
    decode_pair(pair, C, mc, sc);
    if (threadIdx.x < 18) {
        const int c = threadIdx.x < 9 ? mc : sc, i = threadIdx.x < 9 ? threadIdx.x : threadIdx.x - 9;
        camF[threadIdx.x] = (float)a.cam[12 * c + i];
    }
    __syncthreads();  // tables, camF
    {
        constexpr int NB = 8;
        for (int i0 = threadIdx.x; i0 < 2 * PJ; i0 += NB * blockDim.x) {
            float2 q[NB];
            float sv[NB];
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int i = min(i0 + u * (int)blockDim.x, 2 * PJ - 1);
                const int src = i < PJ ? mc * PJ + i : sc * PJ + (i - PJ);
                q[u] = __ldg(kf + src);
                sv[u] = __ldg(sf + src);
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int i = i0 + u * (int)blockDim.x;
                if (i < 2 * PJ) {
                    const V3<float> h = back_project<float>(camF + (i < PJ ? 0 : 9), q[u].x, q[u].y);
                    const float cc = dot3(h, h);
                    int dr = i, ds = i;
                    if (lay.rw_r) {
                        const int slot = i < PJ ? 0 : 1, rem = i - slot * PJ;
                        const int pp = J > 1 ? (int)__umulhi((uint32_t)rem, j_magic) : rem, jj = rem - pp * J;
                        dr = jj * lay.rw_r + slot * lay.ppad + pp;
                        ds = jj * lay.rw_s + slot * lay.ppad + pp;
                    }
                    rays[dr] = make_float4(h.x, h.y, h.z, sv[u] < a.prm.kst_f ? -cc : cc);
                    scs[ds] = sv[u];
                }
            }
        }
    }
    __syncthreads();
    const int tpp = (P + kTile - 1) / kTile;
    for (int tt = warp; tt < tpp * tpp; tt += NW)
        gen_match_item<(MATCH_JM != 0)>(a, tb.camD, match_item_desc(a, tb.camD, tb.pairs, pair * tpp * tpp + tt), rays, scs, kf, sf, f, lane, 0, 1, lay);
}

// Any size: rays from the scratch array written by gen_rays_kernel, scores straight from the input (both L2-resident
// while the frame is worked on); one warp per (frame, camera pair, tile).
__global__ void __launch_bounds__(kGenWarps * 32, 2) gen_match_global_kernel(const __grid_constant__ GenArgs a, const float4* __restrict__ rays) {
    extern __shared__ __align__(16) unsigned char smem[];
    MatchTables tb(smem, a);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tpp = (a.P + kTile - 1) / kTile;
    const long long per_frame = (long long)a.npairs * tpp * tpp;
    const long long item = (long long)blockIdx.x * kGenWarps + warp;
    if (item >= (long long)a.F * per_frame) return;
    const int f = (int)(item / per_frame), it = (int)(item - (long long)f * per_frame);
    const size_t R = (size_t)a.C * a.P * a.J;
    gen_match_item<false>(a, tb.camD, match_item_desc(a, tb.camD, tb.pairs, it), rays + (size_t)f * R, a.scores + (size_t)f * R,
                   reinterpret_cast<const float2*>(a.kpts) + (size_t)f * R, a.scores + (size_t)f * R, f, lane);
}

}  // namespace snowtri
