// Streaming general path, second generation of the fuse (reference triangulation.py:138-156) for rigs of up to 8
// cameras, "mixed" precision (float32 bulk, float64 ray-distance numerator and decisions).
//
//   gen_describe     (device function, tail of the clustering kernels) decodes every cluster member once and
//                    recognises CLIQUE clusters: one observation per camera and every pair of the observed
//                    cameras present -- what a correctly matched person is.  Descriptor = person index per camera.
//   mfuse_kernel<C>  warp-autonomous like p1_kernel: a warp owns a contiguous range of frames and walks it in tiles
//                    of <= 32 output rows; lanes run over the flattened (row, joint) index (no idle lanes), the rays
//                    of a joint are built once in registers and shared by all C(C-1)/2 pairs of the clique (fully
//                    unrolled, camera/pair constants in the kernel-parameter constant bank), sum(score*point) is
//                    accumulated with one scalar per camera (alpha_c, see snowtri_math.cuh), the next item's inputs
//                    are in flight while the current one is solved, the person score and the persons-per-frame
//                    count come out of the same launch (was gen_members + gen_fuse + gen_pscore).
//                    Non-clique clusters (ghosts absorbed into a person, rows of a guard-band joint) take the rolled
//                    member loop mf_item_members, the float64-guarded equivalent of gen_fuse_kernel.
//   float64 only where float32 cannot hold the result: the distance numerator d.(hm x hs) -- the epipolar form
//   x_m^T E x_s with E = -M_m^T [d]x M_s (3x3 per camera pair, composed on the host in float64) straight from the pixel
//   coordinates: 8 DFMA per pair, no float64 rays.
#pragma once
#include "snowtri_general.cuh"

namespace snowtri {

template <int C>
struct MFArgs {
    static constexpr int NP = C * (C - 1) / 2;
    const float* kpts;     // (F,C,P,J,2)
    const float* scores;   // (F,C,P,J)
    float* out;            // (F,Pout,Jout,4)
    float* pscores;        // (F,Pout)
    int* nout;             // (F)
    const int* kcount;     // (F) clusters per frame
    const GenDesc* desc;   // (F,Pout)
    const uint2* memb2;    // (F,ncand)
    int* tile_counter;     // next tile of Gw frames (zero at launch)
    int F, P, J, Jout, Pout, Gw, ncand;
    float kst_f, inv_dthr, guard_w;
    double inv_dthr64;
    alignas(16) float camc[C * 12];   // M = R*inv(K), rows padded to 4
    alignas(16) float pdc[NP * 8];    // per pair: d = ts - tm (3), pad, mid = (tm + ts)/2 (3), pad
    alignas(16) double E[NP * 10];    // per pair (stride 10: 16-byte aligned rows for 128-bit constant loads): d.(hm x hs) = [um vm 1] E [us vs 1]^T
    alignas(16) double cam64[C * 12]; // float64 M (rows padded to 4) for the rolled path; t in cam64t
    alignas(16) double cam64t[C * 4];
};

// Rolled path: the members of a cluster one by one, gate decided in float64 inside its guard band.  `kf`/`sf` point at
// the frame's inputs at joint j.
template <int C>
__device__ __noinline__ float4 mf_item_members(const MFArgs<C>& a, const float2* __restrict__ kf, const float* __restrict__ sf,
                                               const uint2* __restrict__ mb, int n) {
    const int J = a.J;
    float S = 0.f, X = 0.f, Y = 0.f, Z = 0.f;
    for (int m = 0; m < n; ++m) {
        const uint2 mm = mb[m];
        const int mc = mm.x >> 24, sc = mm.y >> 24;
        const size_t rm = (size_t)(mm.x & 0xffffffu) * J, rs = (size_t)(mm.y & 0xffffffu) * J;
        const float2 q0 = kf[rm], q1 = kf[rs];
        const float s0 = sf[rm], s1 = sf[rs];
        const V3<float> hm = back_project4<float>(a.camc + 12 * mc, q0.x, q0.y);
        const V3<float> hs = back_project4<float>(a.camc + 12 * sc, q1.x, q1.y);
        const V3<double> hmD = back_project4<double>(a.cam64 + 12 * mc, (double)q0.x, (double)q0.y);
        const V3<double> hsD = back_project4<double>(a.cam64 + 12 * sc, (double)q1.x, (double)q1.y);
        V3<double> dD;
        dD.x = a.cam64t[4 * sc] - a.cam64t[4 * mc];
        dD.y = a.cam64t[4 * sc + 1] - a.cam64t[4 * mc + 1];
        dD.z = a.cam64t[4 * sc + 2] - a.cam64t[4 * mc + 2];
        V3<float> d, mid;
        d.x = (float)dD.x; d.y = (float)dD.y; d.z = (float)dD.z;
        mid.x = (float)((a.cam64t[4 * sc] + a.cam64t[4 * mc]) * 0.5);
        mid.y = (float)((a.cam64t[4 * sc + 1] + a.cam64t[4 * mc + 1]) * 0.5);
        mid.z = (float)((a.cam64t[4 * sc + 2] + a.cam64t[4 * mc + 2]) * 0.5);
        const PairSolN<float> s = pair_solve_n(hm, dot3(hm, hm), hs, dot3(hs, hs), d);
        const float dn = (float)cross_dot(hmD, hsD, dD);
        const float rr = rsqrt_fast(s.det * dn * dn);
        const float rd = rr * s.det;  // 1/dist
        bool far = rd < a.inv_dthr;   // dist > dthr (strict); NaN is not gated (Q8/Q9)
        if (fabsf(rd - a.inv_dthr) < a.guard_w) {  // decide in float64
            const PairSol<double> s64 = pair_solve(hmD, hsD, dD);
            far = rsqrt_fast(s64.qq) * s64.det < a.inv_dthr64;
        }
        if (far || s0 < a.kst_f || s1 < a.kst_f) continue;  // zero score: contributes nothing
        const float gq = (s0 + s1) * rr;
        const float w = gq * s.det;
        const V3<float> v = pair_v(s.n0, s.n1, hm, hs);
        S += w;
        X = fmaf(w, mid.x, fmaf(0.5f * gq, v.x, X));
        Y = fmaf(w, mid.y, fmaf(0.5f * gq, v.y, Y));
        Z = fmaf(w, mid.z, fmaf(0.5f * gq, v.z, Z));
    }
    if (S == 0.f) return make_float4(0.f, 0.f, 0.f, 0.f);  // S == 0 leaves (0,0,0) with score 0 (Q7)
    const float rS = rcp_t(S);
    return make_float4(X * rS, Y * rS, Z * rS, S * 0.0005f * rcp_t((float)n));  // S/n, zero-score members counted (Q6)
}

// per-warp shared memory of mfuse_kernel: row table (32 x 24 B), person-score columns (32 x 33 floats), input stage
template <int C>
__host__ __device__ constexpr size_t mfuse_warp_bytes() { return 32 * 24 + 32 * 33 * 4 + (size_t)C * 32 * 12; }

template <int C, int NT, int MINB, bool ROLLED>
__global__ void __launch_bounds__(NT, MINB) mfuse_kernel(const __grid_constant__ MFArgs<C> a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int J = a.J, Jout = a.Jout, Pout = a.Pout, P = a.P, Gw = a.Gw;
    unsigned char* ws = smem + (size_t)warp * mfuse_warp_bytes<C>();
    unsigned long long* r_obs = reinterpret_cast<unsigned long long*>(ws);  // row table: person per camera,
    int* r_n = reinterpret_cast<int*>(ws + 32 * 8);                         // members,
    int* r_start = r_n + 32;                                                // member list | clique flag,
    int* r_gk = r_start + 32;                                               // frame of the tile << 8 | output slot
    float* part = reinterpret_cast<float*>(ws + 32 * 24);                   // person-score columns [row][33]
    float2* st_uv = reinterpret_cast<float2*>(ws + 32 * 24 + 32 * 33 * 4);  // staged inputs of the next item [C][32]
    float* st_s = reinterpret_cast<float*>(st_uv + C * 32);

    const float2* kp2 = reinterpret_cast<const float2*>(a.kpts);
    const size_t R = (size_t)C * P * J;
    const unsigned lt = (1u << lane) - 1u;

    // Tiles of Gw frames are handed out through a counter (zeroed by the clustering kernel that runs before this one):
    // rows per frame vary, and with static frame ranges the average SM was idle for the last fifth of the launch.
    for (;;) {
        int tile = 0;
        if (lane == 0) tile = atomicAdd(a.tile_counter, 1);
        tile = __shfl_sync(kFull, tile, 0);
        if ((long long)tile * Gw >= a.F) break;
        const int f0 = tile * Gw;
        const int Gc = min(Gw, a.F - f0);
        // ---- rows of the tile: the clusters that have an output slot, compacted in (frame, slot) order ----------
        int nrows;
        {
            const int g = lane / Pout, k = lane - g * Pout;
            const bool inr = lane < Gc * Pout;
            const int K = inr ? a.kcount[f0 + g] : 0;
            const bool active = inr && k < K;
            if (inr && k == 0) a.nout[f0 + g] = K;
            const unsigned act = __ballot_sync(kFull, active);
            nrows = __popc(act);
            if (active) {
                const GenDesc d = a.desc[(size_t)(f0 + g) * Pout + k];
                const int r = __popc(act & lt);
                r_obs[r] = d.obs;
                r_n[r] = d.n;
                r_start[r] = d.start;
                r_gk[r] = (g << 8) | k;
            }
            unsigned idle = __ballot_sync(kFull, inr && !active);  // empty slots: zeros
            while (idle) {
                const int l = __ffs(idle) - 1;
                idle &= idle - 1;
                const int gg = l / Pout, kk = l - gg * Pout;
                float4* o = reinterpret_cast<float4*>(a.out) + ((size_t)(f0 + gg) * Pout + kk) * Jout;
                for (int jj = lane; jj < Jout; jj += 32) o[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane == 0) a.pscores[(size_t)(f0 + gg) * Pout + kk] = 0.f;
            }
        }
        for (int r = 0; r < nrows; ++r) part[r * 33 + lane] = 0.f;
        __syncwarp();
        const int nitems = nrows * Jout;
        if (nitems == 0) continue;

        // item = (row, joint); lanes past the end redo the last item and store nothing
        auto locate = [&](int q, int& row, int& j) {
            q = min(q, nitems - 1);
            row = q / Jout;
            j = q - row * Jout;
        };
        // issue the copies of an item's C input rows into this lane's stage slots (absent camera: its person 0)
        auto stage_inputs = [&](int row, int j) {
            const unsigned long long obs = r_obs[row];
            const size_t fo = (size_t)(f0 + (r_gk[row] >> 8)) * R + j;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const unsigned p = (unsigned)(obs >> (8 * c)) & 0xffu;
                const size_t off = fo + (size_t)(c * P + (p == 0xffu ? 0u : p)) * J;
                cp_async8(st_uv + c * 32 + lane, kp2 + off);
                cp_async4(st_s + c * 32 + lane, a.scores + off);
            }
        };
        int row, j;
        locate(lane, row, j);
        stage_inputs(row, j);
        // this lane's share of the person score of row `racc`: the items of one row a lane meets are 32 joints apart, so
        // they all belong to column j & 31 of the row -- the sum does not depend on where the row sits in the tile
        float acc = 0.f;
        int racc = row, cacc = j & 31;

        for (int q0 = 0; q0 < nitems; q0 += 32) {
            const bool live = q0 + lane < nitems;
            const unsigned long long obs = r_obs[row];
            const int n = r_n[row], start = r_start[row], gk = r_gk[row];
            const bool clique = start < 0;
            V3<float> h[C];
            float A[C], sc[C];
            double ud[C], vd[C];
            cp_async_wait_all();
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float2 p2 = st_uv[c * 32 + lane];
                const float s1 = st_s[c * 32 + lane];
                h[c] = back_project4<float>(a.camc + 12 * c, p2.x, p2.y);
                A[c] = dot3(h[c], h[c]);
                ud[c] = (double)p2.x;
                vd[c] = (double)p2.y;
                const bool here = ((unsigned)(obs >> (8 * c)) & 0xffu) != 0xffu;
                // a score below the keypoint threshold (or an absent camera) kills every pair of the camera: poison it
                // so that max(sm + ss, 0) is 0 (scores that pass are >= kst >= 0 on this path)
                sc[c] = (!here || s1 < a.kst_f) ? -1e30f : s1;
            }
            // next item of this lane: its copies travel while this one is solved (a lane reads only its own slots)
            int rown, jn;
            locate(q0 + 32 + lane, rown, jn);
            if (q0 + 32 < nitems) stage_inputs(rown, jn);

            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            const size_t fo = (size_t)(f0 + (gk >> 8));
            const float2* kfj = kp2 + fo * R + j;
            const float* sfj = a.scores + fo * R + j;
            const uint2* mb = a.memb2 + fo * a.ncand + (start & 0x7fffffff);
            if (clique) {  // per lane: a row's bits must not depend on which rows share its warp
                // The pairs (x, y), x < y.  ROLLED: x in a rolled loop, y unrolled -- all 28 pairs of 8 cameras unrolled
                // are 25 KB of straight-line code per item and 40 % of the warp stalls were instruction fetches
                // (profiles/r2a); rolled, the body is 7 pair solves, camera x's registers are picked by a switch on
                // the warp-uniform x and the pair constants are indexed at run time.  An absent camera (sub-clique)
                // or a low score contributes exactly zero through the score test.
                float S = 0.f, Xm = 0.f, Ym = 0.f, Zm = 0.f, X = 0.f, Y = 0.f, Z = 0.f, al[C];
                float margin = INFINITY;  // smallest |1/dist - 1/dthr| over the pairs
#pragma unroll
                for (int c = 0; c < C; ++c) al[c] = 0.f;
                // one pair: camera x's values in scalars, camera y's by static index, pair constants at index e
                auto pair = [&](const V3<float>& hx, float Ax, float sx, double udx, double vdx, float& alx, int y, int e) {
                    V3<float> d;
                    d.x = a.pdc[e * 8]; d.y = a.pdc[e * 8 + 1]; d.z = a.pdc[e * 8 + 2];
                    const PairSolN<float> s = pair_solve_n(hx, Ax, h[y], A[y], d);
                    // d.(hm x hs) in float64 from the pixel coordinates: l = E [uy vy 1]^T, dn = [ux vx 1] l
                    const double* E = a.E + 10 * e;
                    const double l0 = fma(E[0], ud[y], fma(E[1], vd[y], E[2]));
                    const double l1 = fma(E[3], ud[y], fma(E[4], vd[y], E[5]));
                    const double l2 = fma(E[6], ud[y], fma(E[7], vd[y], E[8]));
                    const float dn = (float)fma(l0, udx, fma(l1, vdx, l2));
                    const float r = rsqrt_fast(s.det * dn * dn);  // q.q = det * (d.n)^2
                    const float rd = r * s.det;                   // 1/dist
                    margin = fminf(margin, fabsf(rd - a.inv_dthr));
                    // score: zero when a score is below kst (poisoned, also absent cameras) or dist > dthr
                    // (strict; a NaN distance is not gated, Q8/Q9)
                    const float ssum = sx + sc[y];
                    const float gq = (ssum > 0.f && !(rd < a.inv_dthr)) ? ssum * r : 0.f;
                    const float w = gq * s.det;
                    S += w;
                    alx = fmaf(gq, s.n0, alx);
                    al[y] = fmaf(-gq, s.n1, al[y]);
                    Xm = fmaf(w, a.pdc[e * 8 + 4], Xm);
                    Ym = fmaf(w, a.pdc[e * 8 + 5], Ym);
                    Zm = fmaf(w, a.pdc[e * 8 + 6], Zm);
                };
                if constexpr (ROLLED) {
#pragma unroll 1
                    for (int x = 0; x < C - 1; ++x) {
                        V3<float> hx = h[0];
                        float Ax = A[0], sx = sc[0], alx = al[0];
                        double udx = ud[0], vdx = vd[0];
                        switch (x) {
#define MF_PICK(c_)                                                                                     \
    case c_:                                                                                            \
        if constexpr (c_ < C) {                                                                         \
            hx = h[c_]; Ax = A[c_]; sx = sc[c_]; alx = al[c_]; udx = ud[c_]; vdx = vd[c_];              \
        }                                                                                               \
        break;
                            MF_PICK(1) MF_PICK(2) MF_PICK(3) MF_PICK(4) MF_PICK(5) MF_PICK(6) MF_PICK(7)
#undef MF_PICK
                            default: break;
                        }
                        const int ebase = x * C - x * (x + 1) / 2 - x - 1;  // pair_index(C, x, y) = ebase + y
#pragma unroll
                        for (int y = 1; y < C; ++y)
                            if (y > x) pair(hx, Ax, sx, udx, vdx, alx, y, ebase + y);
                        X = fmaf(alx, hx.x, X);  // camera x has met every partner: its alpha is complete
                        Y = fmaf(alx, hx.y, Y);
                        Z = fmaf(alx, hx.z, Z);
                    }
                    X = fmaf(al[C - 1], h[C - 1].x, X);
                    Y = fmaf(al[C - 1], h[C - 1].y, Y);
                    Z = fmaf(al[C - 1], h[C - 1].z, Z);
                } else {
                    static_for_pairs<C>([&](auto xc, auto yc) {
                        constexpr int x = decltype(xc)::value, y = decltype(yc)::value;
                        pair(h[x], A[x], sc[x], ud[x], vd[x], al[x], y, pair_index(C, x, y));
                    });
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        X = fmaf(al[c], h[c].x, X);
                        Y = fmaf(al[c], h[c].y, Y);
                        Z = fmaf(al[c], h[c].z, Z);
                    }
                }
                if (margin < a.guard_w) {  // a float32 distance within the guard band of dthr: decide in float64
                    o = mf_item_members<C>(a, kfj, sfj, mb, n);
                } else if (S != 0.f) {  // S == 0 leaves (0,0,0) with score 0 (Q7)
                    const float rS = rcp_fast(S);
                    o.x = fmaf(0.5f, X, Xm) * rS;
                    o.y = fmaf(0.5f, Y, Ym) * rS;
                    o.z = fmaf(0.5f, Z, Zm) * rS;
                    o.w = S * 0.0005f * rcp_t((float)n);  // S/n, zero-score members counted (Q6)
                }
            } else {
                o = mf_item_members<C>(a, kfj, sfj, mb, n);
            }
            if (live) {
                reinterpret_cast<float4*>(a.out)[(fo * Pout + (gk & 0xff)) * Jout + j] = o;
                if (row != racc) {
                    part[racc * 33 + cacc] = acc;
                    acc = 0.f;
                    racc = row;
                    cacc = j & 31;
                }
                acc += o.w;
            }
            row = rown;
            j = jn;
        }
        if (lane < nitems) part[racc * 33 + cacc] = acc;
        __syncwarp();
        // ---- person score = mean keypoint score (reference :150): lane r adds up the 32 columns of row r ----------
        if (lane < nrows) {
            const float* prow = part + lane * 33;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                s0 += prow[i];
                s1 += prow[i + 1];
                s2 += prow[i + 2];
                s3 += prow[i + 3];
            }
            const int gk = r_gk[lane];
            a.pscores[(size_t)(f0 + (gk >> 8)) * Pout + (gk & 0xff)] = ((s0 + s1) + (s2 + s3)) / (float)Jout;
        }
        __syncwarp();
    }
}

}  // namespace snowtri
