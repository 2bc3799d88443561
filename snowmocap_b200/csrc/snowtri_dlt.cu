// Optional DLT mode: the estimator BASELINE.json's north_star describes -- a 2C x 4
// homogeneous system per joint, solved through the 4 x 4 normal matrix.  The reference (snowvision) does NOT
// triangulate this way (it fuses pairwise skew-ray midpoints), so this mode is not parity-graded against it; its
// oracle is NumPy's SVD of the same matrix (oracle/dlt_oracle.py).
//
// Formulation (one person per camera, every camera sees the same person):
//   Q_c = [R_c^T | -R_c^T t_c]  (3 x 4, world -> camera), (xn, yn) = normalised image point from K_c^-1 [u v 1]
//   rows  xn*Q_c[2] - Q_c[0],  yn*Q_c[2] - Q_c[1]   for every camera whose score passes kst  ->  A (2V x 4)
//   X = eigenvector of N = A^T A with the smallest eigenvalue, dehomogenised.
// One thread per (frame, joint): N is accumulated in the bulk type T (float: 10 FMAs per row), the eigenvector is
// found in float64 by inverse iteration from the inhomogeneous least-squares solution (two LDL^T solves; the
// spectrum is lambda_min ~ noise^2 << lambda_2, so two steps reach 1e-12).  4 flop/byte: HBM-bound.
#include <math.h>
#include <string.h>

#include "snowtri_internal.h"

namespace snowtri {

constexpr int kDltMaxC = 32;

struct DltArgs {
    const float* kpts;    // (F,C,1,J,2)
    const float* scores;  // (F,C,1,J)
    float* out;           // (F,J,4): x, y, z, number of views used (0 = fewer than two views, point zeroed)
    long long n;          // F*J
    int C, J;
    float kst_f;
    float Q[kDltMaxC * 12];    // world -> camera [R^T | -R^T t], row-major 3 x 4
    float Kinv[kDltMaxC * 9];  // inverse intrinsics
};

// LDL^T factor of the symmetric 4 x 4 normal matrix without pivoting (N: 10 unique entries, row-major upper:
// 0:(0,0) 1:(0,1) 2:(0,2) 3:(0,3) 4:(1,1) 5:(1,2) 6:(1,3) 7:(2,2) 8:(2,3) 9:(3,3)).  The reciprocals of the
// pivots are formed once (MUFU.RCP64H + two Newton steps) and shared by the three triangular solves below; the
// leading 3 x 3 block is the factor of the inhomogeneous system.
struct Ldl4 {
    double l10, l20, l30, l21, l31, l32, r0, r1, r2, r3;  // unit lower factor and 1/d
};
__device__ __forceinline__ Ldl4 ldl4_factor(const double* N) {
    Ldl4 f;
    f.r0 = rcp_t(N[0]);
    f.l10 = N[1] * f.r0; f.l20 = N[2] * f.r0; f.l30 = N[3] * f.r0;
    const double d1 = fma(-f.l10, N[1], N[4]);
    f.r1 = rcp_t(d1);
    f.l21 = fma(-f.l20, N[1], N[5]) * f.r1;
    f.l31 = fma(-f.l30, N[1], N[6]) * f.r1;
    const double d2 = fma(-f.l21 * f.l21, d1, fma(-f.l20, N[2], N[7]));
    f.r2 = rcp_t(d2);
    f.l32 = fma(-f.l31 * f.l21, d1, fma(-f.l30, N[2], N[8])) * f.r2;
    const double d3 = fma(-f.l32 * f.l32, d2, fma(-f.l31 * f.l31, d1, fma(-f.l30, N[3], N[9])));
    f.r3 = rcp_t(d3);
    return f;
}
__device__ __forceinline__ void ldl4_solve(const Ldl4& f, const double* b, double* y) {
    const double z0 = b[0];
    const double z1 = fma(-f.l10, z0, b[1]);
    const double z2 = fma(-f.l21, z1, fma(-f.l20, z0, b[2]));
    const double z3 = fma(-f.l32, z2, fma(-f.l31, z1, fma(-f.l30, z0, b[3])));
    const double w3 = z3 * f.r3;
    const double w2 = fma(z2, f.r2, -f.l32 * w3);
    const double w1 = fma(z1, f.r1, fma(-f.l21, w2, -f.l31 * w3));
    const double w0 = fma(z0, f.r0, fma(-f.l10, w1, fma(-f.l20, w2, -f.l30 * w3)));
    y[0] = w0; y[1] = w1; y[2] = w2; y[3] = w3;
}

template <typename T, int CT>  // CT > 0: camera count known at compile time (loads issued up front, loop unrolled)
__global__ void __launch_bounds__(256) dlt_kernel(const __grid_constant__ DltArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const long long f = i / a.J;
    const int j = (int)(i - f * a.J);
    const float2* kp = reinterpret_cast<const float2*>(a.kpts) + (size_t)f * a.C * a.J + j;
    const float* sp = a.scores + (size_t)f * a.C * a.J + j;
    T N[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) N[k] = (T)0;
    int views = 0;
    auto add_view = [&](int c, float2 q) {
        const float* Ki = a.Kinv + 9 * c;
        const float iw = __frcp_rn(fmaf(Ki[6], q.x, fmaf(Ki[7], q.y, Ki[8])));
        const T xn = (T)(fmaf(Ki[0], q.x, fmaf(Ki[1], q.y, Ki[2])) * iw);
        const T yn = (T)(fmaf(Ki[3], q.x, fmaf(Ki[4], q.y, Ki[5])) * iw);
        const float* Q = a.Q + 12 * c;
        T r[4], t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            r[k] = fma(xn, (T)Q[8 + k], -(T)Q[k]);
            t[k] = fma(yn, (T)Q[8 + k], -(T)Q[4 + k]);
        }
        int e = 0;
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int qq = p; qq < 4; ++qq) {
                N[e] = fma(r[p], r[qq], fma(t[p], t[qq], N[e]));
                ++e;
            }
    };
    if constexpr (CT > 0) {
        float2 q[CT];
        float sc[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            q[c] = __ldg(kp + (size_t)c * a.J);
            sc[c] = __ldg(sp + (size_t)c * a.J);
        }
#pragma unroll
        for (int c = 0; c < CT; ++c)
            if (!(sc[c] < a.kst_f)) {
                ++views;
                add_view(c, q[c]);
            }
    } else {
        for (int c = 0; c < a.C; ++c) {
            const float2 q = __ldg(kp + (size_t)c * a.J);
            const float s = __ldg(sp + (size_t)c * a.J);
            if (s < a.kst_f) continue;
            ++views;
            add_view(c, q);
        }
    }
    float4 o = make_float4(0.f, 0.f, 0.f, (float)views);
    if (views >= 2) {
        double Nd[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) Nd[k] = (double)N[k];
        const Ldl4 f = ldl4_factor(Nd);
        // inhomogeneous start: N[:3,:3] p = -N[:3,3] with the leading 3 x 3 block of the factor, x0 = (p, 1)
        double x[4];
        {
            const double z0 = -Nd[3], z1 = fma(-f.l10, z0, -Nd[6]), z2 = fma(-f.l21, z1, fma(-f.l20, z0, -Nd[8]));
            const double w2 = z2 * f.r2, w1 = fma(z1, f.r1, -f.l21 * w2), w0 = fma(z0, f.r0, fma(-f.l10, w1, -f.l20 * w2));
            x[0] = w0; x[1] = w1; x[2] = w2; x[3] = 1.0;
        }
#pragma unroll
        for (int it = 0; it < 2; ++it) {  // inverse iteration on N; rescaled by 1/y3 (only the direction matters)
            double y[4];
            ldl4_solve(f, x, y);
            const double inv = rcp_t(y[3]);
            x[0] = y[0] * inv; x[1] = y[1] * inv; x[2] = y[2] * inv; x[3] = 1.0;
        }
        o.x = (float)x[0];
        o.y = (float)x[1];
        o.z = (float)x[2];
    }
    reinterpret_cast<float4*>(a.out)[i] = o;
}

}  // namespace snowtri

using namespace snowtri;

extern "C" int snowtri_dlt_run(snowtri_t* h, const float* d_kpts, const float* d_scores, int F, int J, float* d_out,
                               int accumulate_f64, void* stream) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_dlt_run: NULL handle");
    if (F == 0) return SNOWTRI_OK;
    if (!d_kpts || !d_scores || !d_out || F < 0 || J < 1) return fail(h, SNOWTRI_E_ARG, "snowtri_dlt_run: bad argument");
    if (h->C > kDltMaxC) return fail(h, SNOWTRI_E_UNSUPPORTED, "snowtri_dlt_run: at most %d cameras", kDltMaxC);
    if ((((uintptr_t)d_kpts) & 7u) || (((uintptr_t)d_out) & 15u)) return fail(h, SNOWTRI_E_ARG, "snowtri_dlt_run: misaligned buffer");
    CUDA_TRY(h, cudaSetDevice(h->device));
    DltArgs a;
    memset(&a, 0, sizeof(a));
    a.kpts = d_kpts; a.scores = d_scores; a.out = d_out;
    a.n = (long long)F * J; a.C = h->C; a.J = J; a.kst_f = h->prm.kst_f;
    for (int c = 0; c < h->C; ++c) {
        // the handle keeps K^-1 and R (camera -> world) next to M = R*K^-1 and t for this mode
        const double* Kinv = h->kinv_host + 9 * c;
        const double* R = h->r_host + 9 * c;
        const double* t = h->cam_host + 12 * c + 9;
        for (int k = 0; k < 9; ++k) a.Kinv[9 * c + k] = (float)Kinv[k];
        for (int r = 0; r < 3; ++r) {
            double rt = 0.0;
            for (int k = 0; k < 3; ++k) {
                a.Q[12 * c + 4 * r + k] = (float)R[3 * k + r];  // R^T
                rt += R[3 * k + r] * t[k];
            }
            a.Q[12 * c + 4 * r + 3] = (float)(-rt);
        }
    }
    const unsigned blocks = (unsigned)((a.n + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
#define DLT_LAUNCH(CT_)                                                        \
    do {                                                                       \
        if (accumulate_f64) dlt_kernel<double, CT_><<<blocks, 256, 0, st>>>(a); \
        else dlt_kernel<float, CT_><<<blocks, 256, 0, st>>>(a);                 \
    } while (0)
    switch (h->C) {
        case 2: DLT_LAUNCH(2); break;
        case 3: DLT_LAUNCH(3); break;
        case 4: DLT_LAUNCH(4); break;
        case 6: DLT_LAUNCH(6); break;
        case 8: DLT_LAUNCH(8); break;
        default: DLT_LAUNCH(0); break;
    }
#undef DLT_LAUNCH
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return SNOWTRI_OK;
}
