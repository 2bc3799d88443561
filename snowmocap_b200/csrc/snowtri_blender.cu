// Blender control points on the device: Human_Triangulation_Blender (reference
// snowvision/blender.py:93-143 with its helpers :11-96) and Human_Triangulation_Blender_Smooth (:145-178),
// the two steps main.py:80-87 runs on every frame after Human_Triangulation_Smooth.
//
// blender_kernel: one lane per person row of the snowtri_run / snowtri_condense output (F*Pout rows of J joints,
// x y z score each).  A warp owns a tile of 32 rows.  Only 28 of the 133 joints are used (3..22 and eight hand
// joints); the warp fetches them row by row -- lane l loads joint slot l, one 16/32-byte access each, the 20 body
// joints of a row being contiguous -- and parks them transposed in shared memory ([slot][axis][row], odd pitch,
// no bank conflicts either way), then every lane derives the 24 control points of its own row.  The arithmetic
// is float64 for both layouts (the reference's, and cheap next to the memory traffic: ~0.9 KB moved per row).
//
// Root rotation (blender.py:15-36): the reference stacks the unit pelvis axis x, the unit spine axis y and
// z = x X y / |x X y| as COLUMNS and hands that non-orthogonal matrix to SciPy, which replaces it by the
// nearest rotation U V^T of its SVD (orthogonal Procrustes) before Markley's quaternion extraction.  For this
// matrix the polar factor has a closed form: M^T M = [[1,c,0],[c,1,0],[0,0,1]] with c = x.y, so
// U V^T = M (M^T M)^(-1/2) = [a x + b y | b x + a y | z],  a,b = (1/sqrt(1+c) +- 1/sqrt(1-c)) / 2
// (symmetric orthogonalisation of x and y; z is already orthogonal to both).  No iteration, no SVD.
// SciPy raises on a NaN matrix; the batch kernel writes NaN and clears the control point's valid bit instead
// (the drop-in function raises like the reference, snowmocap_b200/blender.py).
//
// blender_smooth_kernel: one thread per (person, control point), four channels each, walking the frames of the
// batch in order (the recurrence is sequential); per-control-point constants from the smooth profile; state
// (xp, y, yd per channel, first-frame person count) stays on the device between calls.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <cooperative_groups.h>

#include "snowtri_internal.h"

#define SNOWTRI_NCTRL 24
#ifndef BLENDER_MINB
#define BLENDER_MINB 6   // CTAs per SM the register budget is capped for (measured: 6 x 16 best, tools/gpu_blender.sh)
#endif
#ifndef BLENDER_UNROLL
#define BLENDER_UNROLL 16   // row loads in flight per lane while staging a tile
#endif

struct snowtri_blender_smooth_state {
    int device, P;
    double fzr[SNOWTRI_NCTRL * 3];
    double* d_state;  // [0] initialised, [1] n0, then (P, 24, 12): xp[4], y[4], yd[4]
    double* d_work;   // chunk-parallel path: per chunk 33 values per (person, control point), see kBsWork
    size_t work_chunks;
    int sequential;   // 1 = always the single-launch sequential kernel
    int force_scan;   // 1 = long batches always take the chunk scan (pass A / B / C)
    int coop_blocks;  // CTAs of blender_smooth_overlap_kernel a cooperative launch can hold (0: not asked yet, -1: none)
    int forget[SNOWTRI_NCTRL];   // present frames after which a follower has forgotten its state (0: not within kBsMaxForget)
    double forget_T;  // delta_time `forget` was computed for
};

namespace snowtri {

constexpr int kSlots = 28;   // joints a person row contributes
constexpr int kPitch = 33;   // rows per tile + 1
constexpr int kBlenderWarps = 2;
constexpr int kStageUnroll = BLENDER_UNROLL;

// slot -> joint: slots 0..19 are joints 3..22, then the eight hand joints
__device__ __forceinline__ int slot_joint(int s) {
    // 91 palm_l, 96 index_l, 100 hand_l_ik, 108 pinky_l, 112 palm_r, 117 index_r, 121 hand_r_ik, 129 pinky_r
    return s < 20 ? s + 3 : (s == 20 ? 91 : s == 21 ? 96 : s == 22 ? 100 : s == 23 ? 108 : s == 24 ? 112
                             : s == 25 ? 117 : s == 26 ? 121 : 129);
}
__host__ __device__ constexpr int joint_slot(int j) {
    return j <= 22 ? j - 3 : (j == 91 ? 20 : j == 96 ? 21 : j == 100 ? 22 : j == 108 ? 23 : j == 112 ? 24
                              : j == 117 ? 25 : j == 121 ? 26 : 27);
}

struct BlenderArgs {
    const void* pts;   // (rows, J, 4)
    const int* nout;   // (F) or null: every row holds a person
    void* ctrl;        // (rows, 24, 4)
    unsigned* valid;   // (rows)
    long long rows;
    int Pout, J;
};

struct D3 {
    double x, y, z;
};
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ D3 cross(D3 a, D3 b) {   // numpy.cross component order
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 unit(D3 a) {          // v / numpy.linalg.norm(v): 0/0 -> NaN like the reference
    const double n = sqrt(dot(a, a));
    return {a.x / n, a.y / n, a.z / n};
}
__device__ __forceinline__ D3 mid(D3 a, D3 b) { return (a + b) * 0.5; }
__device__ __forceinline__ bool has_nan(D3 a) { return isnan(a.x) || isnan(a.y) || isnan(a.z); }

// Get_Joint_Pole, blender.py:88-96
__device__ __forceinline__ D3 joint_pole(D3 joint, D3 upper, D3 lower) {
    const D3 a = upper - joint, b = lower - joint, c = upper - lower;
    return joint + unit(cross(cross(b, a), c));
}
// Get_Hand_Pole / Get_Foot_Pole, blender.py:65-73, 78-86 (left: cross(second, first); right: cross(first, second))
__device__ __forceinline__ D3 side_pole(D3 first, D3 second, D3 root, bool is_left) {
    const D3 a = first - root, b = second - root;
    return root + unit(is_left ? cross(b, a) : cross(a, b));
}

template <typename V>
__device__ __forceinline__ void put(V* dst, int k, D3 v, unsigned& mask) {
    V o;
    o.x = (decltype(o.x))v.x;
    o.y = (decltype(o.x))v.y;
    o.z = (decltype(o.x))v.z;
    o.w = 0;
    dst[k] = o;
    if (!has_nan(v)) mask |= 1u << k;
}

template <typename V>  // float4 (snowtri_run layout) or double4 (snowtri_condense layout)
__global__ void __launch_bounds__(kBlenderWarps * 32, BLENDER_MINB) blender_kernel(const BlenderArgs a) {
    using T = decltype(V().x);
    __shared__ T sm[kBlenderWarps][kSlots * 3 * kPitch];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* S = sm[warp];
    const long long tiles = (a.rows + 31) / 32;
    const V* __restrict__ pts = reinterpret_cast<const V*>(a.pts);
    V* ctrl = reinterpret_cast<V*>(a.ctrl);
    const int jl = slot_joint(lane < kSlots ? lane : 0);
    for (long long tile = (long long)blockIdx.x * kBlenderWarps + warp; tile < tiles;
         tile += (long long)gridDim.x * kBlenderWarps) {
        const long long r0 = tile * 32;
        const int nrows = (int)min(32LL, a.rows - r0);
        // stage: lane = joint slot, loop over the rows of the tile (independent loads, all in flight together)
        if (lane < kSlots) {
#pragma unroll kStageUnroll
            for (int r = 0; r < nrows; ++r) {
                const V p = pts[(size_t)(r0 + r) * a.J + jl];
                S[(lane * 3 + 0) * kPitch + r] = p.x;
                S[(lane * 3 + 1) * kPitch + r] = p.y;
                S[(lane * 3 + 2) * kPitch + r] = p.z;
            }
        }
        __syncwarp();
        if (lane < nrows) {
            const long long row = r0 + lane;
            bool present = true;
            if (a.nout) {
                const long long f = row / a.Pout;
                present = (int)(row - f * a.Pout) < a.nout[f];
            }
            V* dst = ctrl + (size_t)row * SNOWTRI_NCTRL;
            unsigned mask = 0;
            if (!present) {   // empty slot of the padded output: zeros, nothing valid
                V zero;
                zero.x = zero.y = zero.z = zero.w = 0;
#pragma unroll
                for (int k = 0; k < SNOWTRI_NCTRL; ++k) dst[k] = zero;
            } else {
                auto P = [&](int j) -> D3 {
                    const int s = joint_slot(j) * 3;
                    return {(double)S[(s + 0) * kPitch + lane], (double)S[(s + 1) * kPitch + lane],
                            (double)S[(s + 2) * kPitch + lane]};
                };
                const D3 p3 = P(3), p4 = P(4), p5 = P(5), p6 = P(6), p11 = P(11), p12 = P(12);
                const D3 pelvis_mid = mid(p11, p12), shoulder_mid = mid(p5, p6), ear_mid = mid(p3, p4);
                const D3 spine = shoulder_mid - pelvis_mid, neck = ear_mid - shoulder_mid;
                put(dst, 0, pelvis_mid, mask);                                  // root_position   :11-13
                {                                                               // root_rotation   :15-36
                    const D3 x = unit(p11 - p12), y = unit(spine), z = unit(cross(x, y));
                    const double c = dot(x, y);
                    const double ip = 1.0 / sqrt(1.0 + c), im = 1.0 / sqrt(1.0 - c);
                    const double ca = 0.5 * (ip + im), cb = 0.5 * (ip - im);
                    const D3 xo = x * ca + y * cb, yo = x * cb + y * ca;
                    // rotation matrix m[i][j]: columns xo, yo, z
                    const double m00 = xo.x, m10 = xo.y, m20 = xo.z, m01 = yo.x, m11 = yo.y, m21 = yo.z;
                    const double m02 = z.x, m12 = z.y, m22 = z.z;
                    const double tr = m00 + m11 + m22;
                    double qx, qy, qz, qw;   // Markley: largest of (m00, m11, m22, trace), first wins ties
                    int choice = 0;
                    double best = m00;
                    if (m11 > best) { best = m11; choice = 1; }
                    if (m22 > best) { best = m22; choice = 2; }
                    if (tr > best) { choice = 3; }
                    if (choice == 0) {
                        qx = 1.0 - tr + 2.0 * m00; qy = m10 + m01; qz = m20 + m02; qw = m21 - m12;
                    } else if (choice == 1) {
                        qx = m10 + m01; qy = 1.0 - tr + 2.0 * m11; qz = m21 + m12; qw = m02 - m20;
                    } else if (choice == 2) {
                        qx = m20 + m02; qy = m21 + m12; qz = 1.0 - tr + 2.0 * m22; qw = m10 - m01;
                    } else {
                        qx = m21 - m12; qy = m02 - m20; qz = m10 - m01; qw = 1.0 + tr;
                    }
                    const double qn = sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
                    V o;   // (w, x, y, z), blender.py:27
                    o.x = (T)(qw / qn);
                    o.y = (T)(qx / qn);
                    o.z = (T)(qy / qn);
                    o.w = (T)(qz / qn);
                    dst[1] = o;
                    if (!(isnan(qw / qn) || isnan(qx / qn) || isnan(qy / qn) || isnan(qz / qn))) mask |= 1u << 1;
                }
                put(dst, 2, p6, mask);                                          // clavicle_r_ik
                put(dst, 3, p5, mask);                                          // clavicle_l_ik
                const D3 p10 = P(10), p9 = P(9), p16 = P(16), p15 = P(15);
                put(dst, 4, p10, mask);                                         // arm_r_ik
                put(dst, 5, joint_pole(P(8), p6, p10), mask);                   // arm_r_pole
                put(dst, 6, p9, mask);                                          // arm_l_ik
                put(dst, 7, joint_pole(P(7), p5, p9), mask);                    // arm_l_pole
                put(dst, 8, p16, mask);                                         // leg_r_ik
                put(dst, 9, joint_pole(P(14), p12, p16), mask);                 // leg_r_pole
                put(dst, 10, p15, mask);                                        // leg_l_ik
                put(dst, 11, joint_pole(P(13), p11, p15), mask);                // leg_l_pole
                put(dst, 12, P(121), mask);                                     // hand_r_ik
                put(dst, 13, side_pole(P(117), P(129), P(112), false), mask);   // hand_r_pole
                put(dst, 14, P(100), mask);                                     // hand_l_ik
                put(dst, 15, side_pole(P(96), P(108), P(91), true), mask);      // hand_l_pole
                const D3 p20 = P(20), p21 = P(21), p17 = P(17), p18 = P(18);
                put(dst, 16, mid(p20, p21), mask);                              // foot_r_ik
                put(dst, 17, side_pole(p20, p21, P(22), false), mask);          // foot_r_pole
                put(dst, 18, mid(p17, p18), mask);                              // foot_l_ik
                put(dst, 19, side_pole(p17, p18, P(19), true), mask);           // foot_l_pole
                put(dst, 20, shoulder_mid, mask);                               // chest_ik        :38-40
                put(dst, 21, shoulder_mid + unit(cross(p5 - p6, spine)), mask); // chest_pole      :42-48
                put(dst, 22, shoulder_mid + unit(neck), mask);                  // head_ik         :50-55
                put(dst, 23, ear_mid + unit(cross(p3 - p4, neck)), mask);       // head_pole       :57-63
            }
            a.valid[row] = mask;
        }
        __syncwarp();
    }
}

struct BlenderSmoothArgs {
    void* ctrl;             // (F, Pout, 24, 4), smoothed in place
    const unsigned* valid;  // (F, Pout)
    const int* nout;        // (F)
    int* nsm;               // (F)
    double* state;
    int F, Pout, P;
    double T, invT;
    double k1[SNOWTRI_NCTRL], inv_k2[SNOWTRI_NCTRL], k3[SNOWTRI_NCTRL];
    int forget[SNOWTRI_NCTRL];   // one-pass path only
};

#ifndef BS_BATCH_BYTES
#define BS_BATCH_BYTES 256
#endif
constexpr int kBsBatchBytes = BS_BATCH_BYTES;   // input bytes a thread holds in registers per batch (16 float4 / 8 double4)

template <typename V>
__global__ void __launch_bounds__(96) blender_smooth_kernel(const BlenderSmoothArgs a) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = tid / SNOWTRI_NCTRL, c = tid - k * SNOWTRI_NCTRL;
    if (k >= a.P) return;
    const bool lead = tid == 0;
    double* st = a.state + 2 + (size_t)tid * 12;
    bool init = a.state[0] != 0.0;
    int n0 = (int)a.state[1];
    double xp[4], y[4], yd[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        xp[i] = st[i];
        y[i] = st[4 + i];
        yd[i] = st[8 + i];
    }
    const double k1 = a.k1[c], inv_k2 = a.inv_k2[c], k3 = a.k3[c];
    V* ctrl = reinterpret_cast<V*>(a.ctrl);
    const bool inrange = k < a.Pout;
    const size_t stride = (size_t)a.Pout * SNOWTRI_NCTRL;
    const size_t base = (size_t)(inrange ? k : 0) * SNOWTRI_NCTRL + c;
    const size_t vbase = inrange ? k : 0;
    // inputs a batch of frames at a time: loads back to back (clamped addresses), one wait, then the steps
    constexpr int NB = kBsBatchBytes / (int)sizeof(V);
    for (int t0 = 0; t0 < a.F; t0 += NB) {
        V buf[NB];
        unsigned vb[NB];
        int nb[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int t = min(t0 + u, a.F - 1);
            buf[u] = ctrl[(size_t)t * stride + base];
            vb[u] = a.valid[(size_t)t * a.Pout + vbase];
            nb[u] = a.nout[t];
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int t = t0 + u;
            if (t >= a.F) break;
            const V p = buf[u];
            const bool ok = (vb[u] >> c) & 1u;
            const int n = min(max(nb[u], 0), a.Pout);
            const double x[4] = {(double)p.x, (double)p.y, (double)p.z, (double)p.w};
            if (!init) {   // first frame of the clip (blender.py:165-176): followers start at the control point,
                init = true;   // or at zero where it is invalid; the frame passes through
                n0 = min(n, a.P);
                if (k < n0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        xp[i] = y[i] = ok ? x[i] : 0.0;
                        yd[i] = 0.0;
                    }
                }
                if (lead) a.nsm[t] = n;
                continue;
            }
            const int m = min(n, n0);
            if (lead) a.nsm[t] = m;
            if (k < m) {   // blender.py:155-160 + triangulation.py:15-22; an invalid control point re-feeds xp
                V o;
                double out[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double xi = ok ? x[i] : xp[i];
                    const double xd = (xi - xp[i]) * a.invT;
                    xp[i] = xi;
                    y[i] = y[i] + a.T * yd[i];
                    yd[i] = yd[i] + a.T * (xi + k3 * xd - y[i] - k1 * yd[i]) * inv_k2;
                    out[i] = y[i];
                }
                o.x = (decltype(o.x))out[0];
                o.y = (decltype(o.x))out[1];
                o.z = (decltype(o.x))out[2];
                o.w = (decltype(o.x))out[3];
                ctrl[(size_t)t * stride + base] = o;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        st[i] = xp[i];
        st[4 + i] = y[i];
        st[8 + i] = yd[i];
    }
    // every thread has read state[0..1] before any thread can get here only within a CTA; the flags are
    // therefore rewritten by a separate tiny launch (blender_smooth_finish_kernel)
}

// ---- chunk-parallel path for long batches --------------------------------------------------------------------
// One step of a follower is affine in its state s = (xp, y, yd): with a valid control point s' = A1 s + b1 x, with
// an invalid one (the follower is fed its own xp) s' = A0 s, for an absent person s' = s.  So a chunk of frames
// acts on its start state as s_end = M s_start + b, where M (3x3, shared by the four channels of a control point)
// depends only on the valid/absent pattern of the chunk and b on the inputs:
//   pass A  every (chunk, person, control point) thread runs the chunk from a ZERO state with the real inputs (-> b,
//           four channels) and from the three unit states with zero inputs (-> the columns of M).  Chunk 0 knows
//           its true start (the stored state, or the seeding first frame of a clip), so it runs for real, writes
//           its outputs and leaves its end state in b;
//   pass B  start_c = M_(c-1) start_(c-1) + b_(c-1) along the chunks, as a two-level chain (groups of 32 chunks);
//   pass C  every chunk >= 1 runs again from its true start state and writes the smoothed control points.
// Inside a chunk the arithmetic is the reference's recurrence; only the hand-over between chunks is evaluated
// differently (agreement with the sequential kernel ~1e-15 relative).
constexpr int kBsChunk = 128;
constexpr int kBsWork = 33;   // per (chunk, thread): M (9), b (4 channels x 3), start (4 channels x 3)

struct BsChunkArgs {
    BlenderSmoothArgs s;
    double* work;   // [(chunk * kBsWork + i) * NT + tid]
    int nchunks, NT;
};

// (bulk kernels: the update with its constants folded, 5 instead of 7 float64 operations per channel -- see follower_step
// in snowtri_smooth.cu; the sequential kernel keeps the reference's operation order)
struct BsCoef {
    double T, g, cd, cx, cp;
};
__device__ __forceinline__ BsCoef bs_coef(const BlenderSmoothArgs& a, double k1, double inv_k2, double k3) {
    BsCoef q;
    q.T = a.T;
    q.g = a.T * inv_k2;
    q.cd = 1.0 - q.g * k1;
    q.cx = q.g * (1.0 + k3 * a.invT);
    q.cp = q.g * (k3 * a.invT);
    return q;
}
__device__ __forceinline__ void bs_step(const BsCoef& q, bool ok, double x, double& xp, double& y, double& yd) {
    const double xi = ok ? x : xp;   // blender.py:157-160
    const double t1 = fma(q.cx, xi, -(q.cp * xp));   // triangulation.py:15-22
    xp = xi;
    y = fma(q.T, yd, y);
    yd = fma(-q.g, y, fma(q.cd, yd, t1));
}

template <typename V, int PASS>   // PASS 0 = A, 2 = C
__global__ void __launch_bounds__(96) blender_smooth_chunk_kernel(const BsChunkArgs ca) {
    const BlenderSmoothArgs& a = ca.s;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = gtid / ca.NT + (PASS == 2 ? 1 : 0), tid = gtid % ca.NT;
    if (chunk >= ca.nchunks) return;
    const int k = tid / SNOWTRI_NCTRL, c = tid - k * SNOWTRI_NCTRL;
    const bool was_init = a.state[0] != 0.0;
    const int n0 = was_init ? (int)a.state[1] : min(min(max(a.nout[0], 0), a.Pout), a.P);
    const double k1 = a.k1[c], inv_k2 = a.inv_k2[c], k3 = a.k3[c];
    const BsCoef coef = bs_coef(a, k1, inv_k2, k3);
    V* ctrl = reinterpret_cast<V*>(a.ctrl);
    const bool inrange = k < a.Pout;
    const size_t stride = (size_t)a.Pout * SNOWTRI_NCTRL;
    const size_t base = (size_t)(inrange ? k : 0) * SNOWTRI_NCTRL + c;
    const size_t vbase = inrange ? k : 0;
    const int t_begin = chunk * kBsChunk, t_end = min(a.F, t_begin + kBsChunk);
    double* w = ca.work + (size_t)chunk * kBsWork * ca.NT + tid;
    const bool real = PASS == 2 || chunk == 0;   // runs from the true start state and writes outputs
    double xp[4], y[4], yd[4];
    double bx[3] = {1.0, 0.0, 0.0}, by[3] = {0.0, 1.0, 0.0}, bd[3] = {0.0, 0.0, 1.0};   // unit states (pass A)
    if (PASS == 2) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            xp[i] = w[(size_t)(21 + 3 * i + 0) * ca.NT];
            y[i] = w[(size_t)(21 + 3 * i + 1) * ca.NT];
            yd[i] = w[(size_t)(21 + 3 * i + 2) * ca.NT];
        }
    } else if (chunk == 0) {
        const double* st = a.state + 2 + (size_t)tid * 12;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            xp[i] = st[i];
            y[i] = st[4 + i];
            yd[i] = st[8 + i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) xp[i] = y[i] = yd[i] = 0.0;
    }
    bool init = was_init || chunk > 0;
    // a batch of frames at a time: all loads issued back to back (clamped addresses), one wait, then the steps
    constexpr int NB = kBsBatchBytes / (int)sizeof(V);
    for (int t0 = t_begin; t0 < t_end; t0 += NB) {
        V buf[NB];
        unsigned vb[NB];
        int nb[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int t = min(t0 + u, t_end - 1);
            buf[u] = ctrl[(size_t)t * stride + base];
            vb[u] = a.valid[(size_t)t * a.Pout + vbase];
            nb[u] = a.nout[t];
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int t = t0 + u;
            if (t >= t_end) break;
            const V p = buf[u];
            const bool ok = (vb[u] >> c) & 1u;
            const int n = min(max(nb[u], 0), a.Pout);
            const double x[4] = {(double)p.x, (double)p.y, (double)p.z, (double)p.w};
            if (!init) {   // only chunk 0 of a new clip: the seeding frame (blender.py:165-176)
                init = true;
                if (k < n0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        xp[i] = y[i] = ok ? x[i] : 0.0;
                        yd[i] = 0.0;
                    }
                }
                if (PASS == 0 && tid == 0) a.nsm[t] = n;
                continue;
            }
            const int m = min(n, n0);
            if (PASS == 0 && tid == 0) a.nsm[t] = m;
            if (k < m) {
#pragma unroll
                for (int i = 0; i < 4; ++i) bs_step(coef, ok, x[i], xp[i], y[i], yd[i]);
                if (real) {
                    V o;
                    o.x = (decltype(o.x))y[0];
                    o.y = (decltype(o.x))y[1];
                    o.z = (decltype(o.x))y[2];
                    o.w = (decltype(o.x))y[3];
                    ctrl[(size_t)t * stride + base] = o;
                } else {
#pragma unroll
                    for (int j = 0; j < 3; ++j) bs_step(coef, ok, 0.0, bx[j], by[j], bd[j]);
                }
            }
        }
    }
    if (PASS == 0) {
        // M column j = image of unit state j: rows (xp, y, yd)
#pragma unroll
        for (int j = 0; j < 3; ++j) {   // chunk 0 ran from its true start: its end state does not depend on start_0
            w[(size_t)(0 + j) * ca.NT] = chunk == 0 ? 0.0 : bx[j];
            w[(size_t)(3 + j) * ca.NT] = chunk == 0 ? 0.0 : by[j];
            w[(size_t)(6 + j) * ca.NT] = chunk == 0 ? 0.0 : bd[j];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            w[(size_t)(9 + 3 * i + 0) * ca.NT] = xp[i];
            w[(size_t)(9 + 3 * i + 1) * ca.NT] = y[i];
            w[(size_t)(9 + 3 * i + 2) * ca.NT] = yd[i];
        }
    }
}


// ---- one-pass path: chunks with a warm-up instead of a hand-over (see smooth_overlap_kernel in snowtri_smooth.cu) -------
// The (y, yd) part of a follower's state decays by the same 2x2 map on every present frame, valid or not; xp is
// replaced by the input on a valid frame and kept on an invalid one.  So a (chunk, person, control point) thread that
// starts from a ZERO state on a present, valid frame and then walks W present frames (|A^W| < 1e-19, W per control
// point from its f, z, r) holds the exact state to the last bit of a float64 when it reaches its own chunk.  A thread
// that runs out of history first (first chunk, absent person, a control point that stayed invalid) starts at frame 0
// of the batch from the carried state: the sequential recurrence.  Cooperative launch: a grid-wide barrier separates
// the warm-ups (they read frames of earlier chunks, which are smoothed in place) from the chunk walks.
constexpr int kBsMaxForget = 256;

template <bool B>
struct BsBool {
    static constexpr bool value = B;
};

template <typename V>
__global__ void __launch_bounds__(96) blender_smooth_overlap_kernel(const BlenderSmoothArgs a, int L, int nchunks) {
    const int NT = a.P * SNOWTRI_NCTRL;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = gid < (long long)nchunks * NT;
    const int chunk = active ? (int)(gid / NT) : 0, tid = active ? (int)(gid - (long long)chunk * NT) : 0;
    const int k = tid / SNOWTRI_NCTRL, c = tid - k * SNOWTRI_NCTRL;
    const bool was_init = a.state[0] != 0.0;
    const int n0 = was_init ? (int)a.state[1] : min(min(max(a.nout[0], 0), a.Pout), a.P);
    const double k1 = a.k1[c], inv_k2 = a.inv_k2[c], k3 = a.k3[c];
    const BsCoef coef = bs_coef(a, k1, inv_k2, k3);
    V* ctrl = reinterpret_cast<V*>(a.ctrl);
    const bool inrange = k < a.Pout;
    const size_t stride = (size_t)a.Pout * SNOWTRI_NCTRL;
    const size_t base = (size_t)(inrange ? k : 0) * SNOWTRI_NCTRL + c;
    const size_t vbase = inrange ? k : 0;
    const int t_begin = chunk * L, t_end = min(a.F, t_begin + L);
    double xp[4] = {0, 0, 0, 0}, y[4] = {0, 0, 0, 0}, yd[4] = {0, 0, 0, 0};
    constexpr int NB = kBsBatchBytes / (int)sizeof(V);
    auto walk = [&](int t_from, int t_to, auto ownc) {
        constexpr bool OWN = decltype(ownc)::value;
        for (int t0 = t_from; t0 < t_to; t0 += NB) {
            V buf[NB];
            unsigned vb[NB];
            int nb[NB];
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int t = min(t0 + u, t_to - 1);
                buf[u] = ctrl[(size_t)t * stride + base];
                vb[u] = a.valid[(size_t)t * a.Pout + vbase];
                nb[u] = a.nout[t];
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int t = t0 + u;
                if (t >= t_to) break;
                const V p = buf[u];
                const bool ok = (vb[u] >> c) & 1u;
                const int n = min(max(nb[u], 0), a.Pout);
                const double x[4] = {(double)p.x, (double)p.y, (double)p.z, (double)p.w};
                if (!was_init && t == 0) {   // the seeding frame of a clip (blender.py:165-176)
                    if (k < n0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            xp[i] = y[i] = ok ? x[i] : 0.0;
                            yd[i] = 0.0;
                        }
                    }
                    if (OWN && tid == 0) a.nsm[t] = n;
                    continue;
                }
                const int m = min(n, n0);
                if (OWN && tid == 0) a.nsm[t] = m;
                if (k < m) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) bs_step(coef, ok, x[i], xp[i], y[i], yd[i]);
                    if (OWN) {
                        V o;
                        o.x = (decltype(o.x))y[0];
                        o.y = (decltype(o.x))y[1];
                        o.z = (decltype(o.x))y[2];
                        o.w = (decltype(o.x))y[3];
                        ctrl[(size_t)t * stride + base] = o;
                    }
                }
            }
        }
    };
    if (active) {
        // warm-up start: walking back, W present frames, then the next present frame whose control point is valid
        int tw = t_begin, need = a.forget[c];
        bool found = false;
        const bool idle = k >= min(n0, a.Pout);   // a follower this batch never advances: its state is carried over
        if (idle) tw = 0;
        while (tw > 0 && !found) {
            int nb[8];
            unsigned vb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = max(tw - 1 - u, 0);
                nb[u] = a.nout[t];
                vb[u] = a.valid[(size_t)t * a.Pout + vbase];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (tw > 0 && !found) {
                    --tw;
                    const int m = min(min(max(nb[u], 0), a.Pout), n0);
                    if (k < m && (was_init || tw > 0)) {
                        if (need > 0) --need;
                        else if ((vb[u] >> c) & 1u) found = true;
                    }
                }
            }
        }
        if (!found) {  // out of history: from frame 0 of the batch and the state the last batch left
            tw = idle ? t_begin : 0;
            const double* st = a.state + 2 + (size_t)tid * 12;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                xp[i] = st[i];
                y[i] = st[4 + i];
                yd[i] = st[8 + i];
            }
        }
        walk(tw, t_begin, BsBool<false>{});
    }
    cooperative_groups::this_grid().sync();
    if (!active) return;
    walk(t_begin, t_end, BsBool<true>{});
    if (chunk == nchunks - 1) {
        double* st = a.state + 2 + (size_t)tid * 12;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            st[i] = xp[i];
            st[4 + i] = y[i];
            st[8 + i] = yd[i];
        }
        if (tid == 0 && !was_init) {
            a.state[1] = (double)n0;
            a.state[0] = 1.0;
        }
    }
}

// pass B (chunk 0 carries M = 0 and b = its true end state, so the chain start_(c+1) = M_c start_c + b_c holds for
// every chunk from an arbitrary start_0): see blender_smooth_carry_scan_kernel below.

struct Affine {   // s -> M s + b for the four channels of a control point
    double M[9], b[12];
};
__device__ __forceinline__ void affine_load(Affine& f, const double* w, int NT) {
#pragma unroll
    for (int i = 0; i < 9; ++i) f.M[i] = w[(size_t)i * NT];
#pragma unroll
    for (int i = 0; i < 12; ++i) f.b[i] = w[(size_t)(9 + i) * NT];
}
__device__ __forceinline__ void affine_apply(const Affine& f, double* s) {   // s (4 channels x 3) <- M s + b
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double s0 = s[3 * i], s1 = s[3 * i + 1], s2 = s[3 * i + 2];
        s[3 * i + 0] = f.b[3 * i + 0] + (f.M[0] * s0 + f.M[1] * s1 + f.M[2] * s2);
        s[3 * i + 1] = f.b[3 * i + 1] + (f.M[3] * s0 + f.M[4] * s1 + f.M[5] * s2);
        s[3 * i + 2] = f.b[3 * i + 2] + (f.M[6] * s0 + f.M[7] * s1 + f.M[8] * s2);
    }
}

// Pass B, one launch: the hand-over is a prefix scan of the chunk maps under composition.  One CTA per
// (person, control point), a thread per chunk (blocks of kBsScanThreads chunks, the state carried from block to block):
// Hillis-Steele scan by warp shuffles, the warp totals scanned by warp 0, every thread applies the prefix of the chunks
// before it.  (Round 1 walked the chunks in three launches over groups of 32: 96 dependent memory round trips,
// 26 + 24 + 30 us per 1024 chunks; this kernel takes 36.5 us, profiles/r2p.)
constexpr int kBsScanThreads = 256;
__device__ __forceinline__ Affine affine_compose(const Affine& later, const Affine& earlier) {  // later(earlier(s))
    Affine r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int q = 0; q < 3; ++q)
            r.M[3 * i + q] = later.M[3 * i] * earlier.M[q] + later.M[3 * i + 1] * earlier.M[3 + q] + later.M[3 * i + 2] * earlier.M[6 + q];
#pragma unroll
    for (int i = 0; i < 12; ++i) r.b[i] = earlier.b[i];
    affine_apply(later, r.b);
    return r;
}
__device__ __forceinline__ Affine affine_shfl_up(const Affine& m, int d) {
    Affine r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.M[i] = __shfl_up_sync(0xffffffffu, m.M[i], d);
#pragma unroll
    for (int i = 0; i < 12; ++i) r.b[i] = __shfl_up_sync(0xffffffffu, m.b[i], d);
    return r;
}
__device__ __forceinline__ Affine affine_identity() {
    Affine r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.M[i] = (i % 4 == 0) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < 12; ++i) r.b[i] = 0.0;
    return r;
}
__device__ __forceinline__ void affine_to_smem(const Affine& m, double* w) {
#pragma unroll
    for (int i = 0; i < 9; ++i) w[i] = m.M[i];
#pragma unroll
    for (int i = 0; i < 12; ++i) w[9 + i] = m.b[i];
}
__device__ __forceinline__ Affine affine_from_smem(const double* w) {
    Affine m;
#pragma unroll
    for (int i = 0; i < 9; ++i) m.M[i] = w[i];
#pragma unroll
    for (int i = 0; i < 12; ++i) m.b[i] = w[9 + i];
    return m;
}

__global__ void __launch_bounds__(kBsScanThreads) blender_smooth_carry_scan_kernel(const BsChunkArgs ca) {
    __shared__ double wtot[kBsScanThreads / 32][21];
    __shared__ double carry[12];
    const int tid = blockIdx.x;  // (person, control point)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 12) carry[threadIdx.x] = 0.0;  // arbitrary: chunk 0 has M = 0
    __syncthreads();
    for (int base = 0; base < ca.nchunks; base += kBsScanThreads) {
        const int chunk = base + threadIdx.x;
        Affine m = affine_identity();
        if (chunk < ca.nchunks) affine_load(m, ca.work + (size_t)chunk * kBsWork * ca.NT + tid, ca.NT);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const Affine up = affine_shfl_up(m, d);
            if (lane >= d) m = affine_compose(m, up);
        }
        if (lane == 31) affine_to_smem(m, wtot[warp]);
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the warp totals
            Affine t = lane < kBsScanThreads / 32 ? affine_from_smem(wtot[lane]) : affine_identity();
#pragma unroll
            for (int d = 1; d < kBsScanThreads / 32; d <<= 1) {
                const Affine up = affine_shfl_up(t, d);
                if (lane >= d) t = affine_compose(t, up);
            }
            Affine ex = affine_shfl_up(t, 1);
            if (lane == 0) ex = affine_identity();
            if (lane < kBsScanThreads / 32) affine_to_smem(ex, wtot[lane]);
        }
        __syncthreads();
        const Affine pre = affine_from_smem(wtot[warp]);
        const Affine incl = affine_compose(m, pre);
        Affine excl = affine_shfl_up(incl, 1);
        if (lane == 0) excl = pre;
        double st[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) st[i] = carry[i];
        if (chunk < ca.nchunks) {
            double v[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) v[i] = st[i];
            affine_apply(excl, v);
            double* w = ca.work + (size_t)chunk * kBsWork * ca.NT + tid;
#pragma unroll
            for (int i = 0; i < 12; ++i) w[(size_t)(21 + i) * ca.NT] = v[i];
        }
        __syncthreads();
        if (threadIdx.x == kBsScanThreads - 1) {
            affine_apply(incl, st);
#pragma unroll
            for (int i = 0; i < 12; ++i) carry[i] = st[i];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double* st = ca.s.state + 2 + (size_t)tid * 12;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            st[i] = carry[3 * i];
            st[4 + i] = carry[3 * i + 1];
            st[8 + i] = carry[3 * i + 2];
        }
    }
}

__global__ void blender_smooth_finish_kernel(double* state, const int* nout, int F, int Pout, int P) {
    // replay the person-count bookkeeping of the batch: first frame of a clip fixes n0
    if (threadIdx.x || blockIdx.x) return;
    if (state[0] == 0.0 && F > 0) {
        state[0] = 1.0;
        state[1] = (double)min(min(max(nout[0], 0), Pout), P);
    }
}

}  // namespace snowtri

using namespace snowtri;

static int blender_run(snowtri_t* h, const void* d_points, bool f64, const int* d_nout, int F, int Pout, int J,
                       void* d_ctrl, unsigned* d_valid, void* stream) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_blender_run: NULL handle");
    if (F < 0 || Pout < 0) return fail(h, SNOWTRI_E_ARG, "snowtri_blender_run: bad argument");
    if (F == 0 || Pout == 0) return SNOWTRI_OK;
    if (!d_points || !d_ctrl || !d_valid) return fail(h, SNOWTRI_E_ARG, "snowtri_blender_run: NULL buffer");
    // the reference indexes person[129] (blender.py:103): fewer joints is its IndexError
    if (J < 130) return fail(h, SNOWTRI_E_ARG, "snowtri_blender_run: needs the Wholebody keypoints (J >= 130), got J=%d", J);
    if ((((uintptr_t)d_points) & 15u) || (((uintptr_t)d_ctrl) & 15u))
        return fail(h, SNOWTRI_E_ARG, "snowtri_blender_run: misaligned buffer");
    CUDA_TRY(h, cudaSetDevice(h->device));
    BlenderArgs a;
    a.pts = d_points; a.nout = d_nout; a.ctrl = d_ctrl; a.valid = d_valid;
    a.rows = (long long)F * Pout; a.Pout = Pout; a.J = J;
    const long long tiles = (a.rows + 31) / 32;
    const long long want = (tiles + kBlenderWarps - 1) / kBlenderWarps;
    const long long cap = (long long)h->sm_count * 16;   // persistent over tiles beyond 16 CTAs per SM
    const unsigned blocks = (unsigned)(want < cap ? want : cap);
    cudaStream_t st = (cudaStream_t)stream;
    if (f64) blender_kernel<double4><<<blocks, kBlenderWarps * 32, 0, st>>>(a);
    else blender_kernel<float4><<<blocks, kBlenderWarps * 32, 0, st>>>(a);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return SNOWTRI_OK;
}

extern "C" int snowtri_blender_run(snowtri_t* h, const float* d_points, const int* d_nout, int F, int Pout, int J,
                                   float* d_ctrl, unsigned* d_valid, void* stream) {
    return blender_run(h, d_points, false, d_nout, F, Pout, J, d_ctrl, d_valid, stream);
}

extern "C" int snowtri_blender_run_f64(snowtri_t* h, const double* d_points, const int* d_nout, int F, int Pout, int J,
                                       double* d_ctrl, unsigned* d_valid, void* stream) {
    return blender_run(h, d_points, true, d_nout, F, Pout, J, d_ctrl, d_valid, stream);
}

extern "C" int snowtri_blender_smooth_create(snowtri_t* h, snowtri_blender_smooth_t** out, int max_persons,
                                             const double* fzr) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_blender_smooth_create: NULL handle");
    if (!out || !fzr || max_persons < 1) return fail(h, SNOWTRI_E_ARG, "snowtri_blender_smooth_create: bad argument");
    for (int c = 0; c < SNOWTRI_NCTRL; ++c)
        if (!(fzr[3 * c] > 0.0))   // the reference divides by f (triangulation.py:7-9)
            return fail(h, SNOWTRI_E_ARG, "snowtri_blender_smooth_create: f must be positive (control point %d)", c);
    CUDA_TRY(h, cudaSetDevice(h->device));
    snowtri_blender_smooth_t* s = (snowtri_blender_smooth_t*)calloc(1, sizeof(*s));
    if (!s) return fail(h, SNOWTRI_E_NOMEM, "snowtri_blender_smooth_create: out of host memory");
    s->device = h->device;
    s->P = max_persons;
    memcpy(s->fzr, fzr, sizeof(s->fzr));
    const size_t bytes = (2 + (size_t)max_persons * SNOWTRI_NCTRL * 12) * sizeof(double);
    cudaError_t e = cudaMalloc(&s->d_state, bytes);
    if (e == cudaSuccess) e = cudaMemset(s->d_state, 0, bytes);
    if (e != cudaSuccess) {
        if (s->d_state) cudaFree(s->d_state);
        free(s);
        return fail(h, SNOWTRI_E_CUDA, "snowtri_blender_smooth_create: %s", cudaGetErrorString(e));
    }
    *out = s;
    return SNOWTRI_OK;
}

extern "C" int snowtri_blender_smooth_destroy(snowtri_blender_smooth_t* s) {
    if (!s) return SNOWTRI_OK;
    cudaSetDevice(s->device);
    cudaFree(s->d_state);
    if (s->d_work) cudaFree(s->d_work);
    free(s);
    return SNOWTRI_OK;
}

extern "C" int snowtri_blender_smooth_set_chunked(snowtri_blender_smooth_t* s, int enabled) {
    if (!s) return SNOWTRI_E_ARG;
    s->sequential = enabled ? 0 : 1;
    s->force_scan = enabled == 2 ? 1 : 0;
    return SNOWTRI_OK;
}

extern "C" int snowtri_blender_smooth_reset(snowtri_t* h, snowtri_blender_smooth_t* s, void* stream) {
    if (!h || !s) return fail(h, SNOWTRI_E_ARG, "snowtri_blender_smooth_reset: NULL argument");
    CUDA_TRY(h, cudaSetDevice(s->device));
    CUDA_TRY(h, cudaMemsetAsync(s->d_state, 0, (2 + (size_t)s->P * SNOWTRI_NCTRL * 12) * sizeof(double),
                                (cudaStream_t)stream));
    return SNOWTRI_OK;
}

static int blender_smooth_run(snowtri_t* h, snowtri_blender_smooth_t* s, void* d_ctrl, bool f64, const unsigned* d_valid,
                              const int* d_nout, int* d_nsm, int F, int Pout, double delta_time, void* stream) {
    if (!h || !s) return fail(h, SNOWTRI_E_ARG, "snowtri_blender_smooth_run: NULL argument");
    if (F < 0 || Pout < 1 || !(delta_time > 0.0)) return fail(h, SNOWTRI_E_ARG, "snowtri_blender_smooth_run: bad argument");
    if (F == 0) return SNOWTRI_OK;
    if (!d_ctrl || !d_valid || !d_nout || !d_nsm) return fail(h, SNOWTRI_E_ARG, "snowtri_blender_smooth_run: NULL buffer");
    CUDA_TRY(h, cudaSetDevice(s->device));
    BlenderSmoothArgs a;
    memset(&a, 0, sizeof(a));
    a.ctrl = d_ctrl; a.valid = d_valid; a.nout = d_nout; a.nsm = d_nsm; a.state = s->d_state;
    a.F = F; a.Pout = Pout; a.P = s->P;
    a.T = delta_time; a.invT = 1.0 / delta_time;
    const double pi = 3.141592653589793;   // math.pi
    for (int c = 0; c < SNOWTRI_NCTRL; ++c) {   // triangulation.py:7-9
        const double f = s->fzr[3 * c], z = s->fzr[3 * c + 1], r = s->fzr[3 * c + 2];
        a.k1[c] = z / (pi * f);
        a.inv_k2[c] = 1.0 / (1.0 / ((2 * pi * f) * (2 * pi * f)));
        a.k3[c] = r * z / (2 * pi * f);
    }
    const int threads = s->P * SNOWTRI_NCTRL;
    cudaStream_t st = (cudaStream_t)stream;
    bool one_pass = false;
    if (!s->sequential && !s->force_scan && F > 2 * kBsChunk) {
        if (s->forget_T != delta_time) {   // powers of the state matrix of one present, valid frame (see snowtri_smooth.cu)
            for (int c = 0; c < SNOWTRI_NCTRL; ++c) {
                const double T = delta_time, k2 = 1.0 / a.inv_k2[c], g = T / k2;
                const double A[9] = {0, 0, 0, 0, 1, T, -a.k3[c] / k2, -g, 1 - g * (T + a.k1[c])};
                double pw[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, nx[9];
                s->forget[c] = 0;
                for (int n = 1; n <= kBsMaxForget && !s->forget[c]; ++n) {
                    double big = 0;
                    for (int r_ = 0; r_ < 3; ++r_)
                        for (int c_ = 0; c_ < 3; ++c_) {
                            double v = 0;
                            for (int m_ = 0; m_ < 3; ++m_) v += A[r_ * 3 + m_] * pw[m_ * 3 + c_];
                            nx[r_ * 3 + c_] = v;
                            big = fmax(big, fabs(v));
                        }
                    memcpy(pw, nx, sizeof(pw));
                    if (big != big) break;
                    if (big < 1e-19) s->forget[c] = n;
                }
            }
            s->forget_T = delta_time;
        }
        int wmax = 0;
        bool all = true;
        for (int c = 0; c < SNOWTRI_NCTRL; ++c) {
            all = all && s->forget[c] > 0;
            wmax = s->forget[c] > wmax ? s->forget[c] : wmax;
            a.forget[c] = s->forget[c];
        }
        if (all && !s->coop_blocks) {
            int dev_coop = 0, per_sm_f = 0, per_sm_d = 0;
            CUDA_TRY(h, cudaDeviceGetAttribute(&dev_coop, cudaDevAttrCooperativeLaunch, s->device));
            CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_f, blender_smooth_overlap_kernel<float4>, 96, 0));
            CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_d, blender_smooth_overlap_kernel<double4>, 96, 0));
            const int per_sm = per_sm_f < per_sm_d ? per_sm_f : per_sm_d;
            s->coop_blocks = dev_coop && per_sm > 0 ? per_sm * h->sm_count : -1;
        }
        if (all && s->coop_blocks > 0) {
            int L = 64;   // frames per chunk (131 072 frames x 24 control points: 32 or 64 frames 0.058 ms, 128 0.061, 256 0.071; profiles/r3g)
            if (const char* e = getenv("SNOWTRI_BS_CHUNK")) L = atoi(e) >= 16 ? atoi(e) : L;   // experiments
            const long long max_threads = (long long)s->coop_blocks * 96;
            while (L < F && ((long long)(F + L - 1) / L) * threads > max_threads) L *= 2;   // (if even one chunk does not fit, the test below fails)
            int nchunks = (F + L - 1) / L;
            const long long g2 = ((long long)nchunks * threads + 95) / 96;
            if (g2 <= s->coop_blocks) {
                void* args[] = {(void*)&a, (void*)&L, (void*)&nchunks};
                const void* fn = f64 ? (const void*)blender_smooth_overlap_kernel<double4> : (const void*)blender_smooth_overlap_kernel<float4>;
                CUDA_TRY(h, cudaLaunchCooperativeKernel(fn, dim3((unsigned)g2), dim3(96), args, 0, st));
                h->launches += 1;
                one_pass = true;
            }
        }
    }
    if (one_pass) return SNOWTRI_OK;
    if (s->sequential || F <= 2 * kBsChunk) {
        if (f64) blender_smooth_kernel<double4><<<(threads + 95) / 96, 96, 0, st>>>(a);
        else blender_smooth_kernel<float4><<<(threads + 95) / 96, 96, 0, st>>>(a);
        CUDA_TRY(h, cudaGetLastError());
        h->launches += 1;
    } else {
        BsChunkArgs ca;
        ca.s = a;
        ca.nchunks = (F + kBsChunk - 1) / kBsChunk;
        ca.NT = threads;
        if (s->work_chunks < (size_t)ca.nchunks) {
            CUDA_TRY(h, cudaStreamSynchronize(st));   // an earlier batch may still be using the old buffer
            if (s->d_work) cudaFree(s->d_work);
            s->d_work = nullptr;
            s->work_chunks = 0;
            const size_t maps = (size_t)ca.nchunks;
            CUDA_TRY(h, cudaMalloc(&s->d_work, maps * kBsWork * threads * sizeof(double)));
            s->work_chunks = (size_t)ca.nchunks;
        }
        ca.work = s->d_work;
        const long long ta = (long long)ca.nchunks * threads, tc = (long long)(ca.nchunks - 1) * threads;
        if (f64) blender_smooth_chunk_kernel<double4, 0><<<(unsigned)((ta + 95) / 96), 96, 0, st>>>(ca);
        else blender_smooth_chunk_kernel<float4, 0><<<(unsigned)((ta + 95) / 96), 96, 0, st>>>(ca);
        blender_smooth_carry_scan_kernel<<<threads, kBsScanThreads, 0, st>>>(ca);
        if (f64) blender_smooth_chunk_kernel<double4, 2><<<(unsigned)((tc + 95) / 96), 96, 0, st>>>(ca);
        else blender_smooth_chunk_kernel<float4, 2><<<(unsigned)((tc + 95) / 96), 96, 0, st>>>(ca);
        CUDA_TRY(h, cudaGetLastError());
        h->launches += 3;
    }
    blender_smooth_finish_kernel<<<1, 32, 0, st>>>(s->d_state, d_nout, F, Pout, s->P);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return SNOWTRI_OK;
}

extern "C" int snowtri_blender_smooth_run(snowtri_t* h, snowtri_blender_smooth_t* s, float* d_ctrl,
                                          const unsigned* d_valid, const int* d_nout, int* d_nsmooth, int F, int Pout,
                                          double delta_time, void* stream) {
    return blender_smooth_run(h, s, d_ctrl, false, d_valid, d_nout, d_nsmooth, F, Pout, delta_time, stream);
}

extern "C" int snowtri_blender_smooth_run_f64(snowtri_t* h, snowtri_blender_smooth_t* s, double* d_ctrl,
                                              const unsigned* d_valid, const int* d_nout, int* d_nsmooth, int F,
                                              int Pout, double delta_time, void* stream) {
    return blender_smooth_run(h, s, d_ctrl, true, d_valid, d_nout, d_nsmooth, F, Pout, delta_time, stream);
}

// main.py:55-87 for a clip in one call: the four stages back to back on one stream, nothing returns to the host in
// between.  Each stage is the entry point of the same name; this only removes the host work between the launches.
extern "C" int snowtri_clip_run(snowtri_t* h, snowtri_smooth_t* sm, snowtri_blender_smooth_t* bs, const float* d_kpts,
                                const float* d_scores, const int* d_counts, int F, int P, int J, int Pout, float* d_out,
                                float* d_pscores, int* d_nout, int* d_nsmooth, float* d_ctrl, unsigned* d_valid,
                                int* d_nfinal, double delta_time, void* stream) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_clip_run: NULL handle");
    if (F == 0) return SNOWTRI_OK;
    if ((sm && !d_nsmooth) || (bs && !d_nfinal)) return fail(h, SNOWTRI_E_ARG, "snowtri_clip_run: NULL count buffer");
    int rc = snowtri_run(h, d_kpts, d_scores, d_counts, F, P, J, J, Pout, d_out, d_pscores, d_nout, stream);
    if (rc != SNOWTRI_OK) return rc;
    const int* counts = d_nout;
    if (sm) {
        rc = snowtri_smooth_run(h, sm, d_out, d_nout, d_nsmooth, F, Pout, J, delta_time, stream);
        if (rc != SNOWTRI_OK) return rc;
        counts = d_nsmooth;
    }
    rc = snowtri_blender_run(h, d_out, counts, F, Pout, J, d_ctrl, d_valid, stream);
    if (rc != SNOWTRI_OK) return rc;
    if (bs) rc = snowtri_blender_smooth_run(h, bs, d_ctrl, d_valid, counts, d_nfinal, F, Pout, delta_time, stream);
    return rc;
}
