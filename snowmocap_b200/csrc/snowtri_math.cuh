// Device math of the triangulation path, templated on the compute type T (double or float).
//
// Reference algebra being evaluated (reference snowvision/triangulation.py:24-31, 72-74):
//   H = [hm hs],  S = (H^T H)^-1 H^T d,  d = ts - tm
//   Wm = tm + hm*S0,  Ws = ts - hs*S1,  dist = |Wm - Ws|,  W = (Wm + Ws)/2
//   score = ((sm + ss)/2) / (dist*1000), zeroed when sm<kst or ss<kst or dist>dthr
// Closed form used here (no division by the 2x2 determinant on the hot path):
//   A=hm.hm  Cc=hs.hs  B=hm.hs  D=hm.d  E=hs.d
//   det = A*Cc - B^2 (>0)   n0 = Cc*D - B*E = det*S0   n1 = A*E - B*D = det*S1
//   q   = hm*n0 + hs*n1 - d*det = det*(Wm - Ws)            => dist = |q|/det
//   v   = hm*n0 - hs*n1                                     => W = mid + v/(2*det), mid=(tm+ts)/2
//   score = (sm+ss)*0.0005*det*rsqrt(q.q)
//   score*(W - mid) = (sm+ss)*0.00025*rsqrt(q.q) * v        (what the fuse step accumulates)
//   dist > dthr  <=>  q.q > (dthr*det)^2
#pragma once
#ifdef __CUDACC_RTC__   // runtime compilation (NVRTC) of a rig-specialised kernel: no host headers
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
#ifndef INFINITY
#define INFINITY __int_as_float(0x7f800000)
#endif
#else
#include <cuda_runtime.h>
#include <stdint.h>
#endif

namespace snowtri {

template <typename T>
struct V3 {
    T x, y, z;
};

template <typename T>
__device__ __forceinline__ T dot3(const V3<T>& a, const V3<T>& b) {
    return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z));
}

// 1/sqrt(x).  double: MUFU.RSQ64H seed (~2^-22) + two Newton steps (~1 ulp); float: MUFU.RSQ.
__device__ __forceinline__ float rsqrt_t(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rsqrt_t(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double t = x * r;
    double e = fma(-t, r, 1.0);
    r = fma(0.5 * r, e, r);
    t = x * r;
    e = fma(-t, r, 1.0);
    r = fma(0.5 * r, e, r);
    return r;
}

// Fused-path variant: one Newton step (relative error ~2e-14 in double), MUFU.RSQ in float.
__device__ __forceinline__ float rsqrt_fast(float x) {
    float r;  // flush-to-zero form: no denormal rescaling around the MUFU (q.q is never denormal for real rays)
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ double rsqrt_fast(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double t = x * r;
    const double e = fma(-t, r, 1.0);
    return fma(0.5 * r, e, r);
}

// 1/x for positive finite x.  double: MUFU.RCP64H seed + two Newton steps; float: MUFU.RCP + one step.
__device__ __forceinline__ float rcp_t(float x) {
    float r = __frcp_rn(x);
    return r;
}
__device__ __forceinline__ double rcp_t(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// 1/x for positive finite normal x without the range checks of __frcp_rn: MUFU.RCP + one Newton step
// (float, <= 1 ulp); the double version is rcp_t.
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}
__device__ __forceinline__ double rcp_fast(double x) { return rcp_t(x); }

template <typename T>
struct PairSol {
    T det, n0, n1, qq;
};

template <typename T>
__device__ __forceinline__ PairSol<T> pair_solve(const V3<T>& hm, const V3<T>& hs, const V3<T>& d) {
    const T A = dot3(hm, hm), Cc = dot3(hs, hs), B = dot3(hm, hs);
    const T D = dot3(hm, d), E = dot3(hs, d);
    PairSol<T> s;
    s.det = fma(A, Cc, -(B * B));
    s.n0 = fma(Cc, D, -(B * E));
    s.n1 = fma(A, E, -(B * D));
    const T qx = fma(hm.x, s.n0, fma(hs.x, s.n1, -(d.x * s.det)));
    const T qy = fma(hm.y, s.n0, fma(hs.y, s.n1, -(d.y * s.det)));
    const T qz = fma(hm.z, s.n0, fma(hs.z, s.n1, -(d.z * s.det)));
    s.qq = fma(qx, qx, fma(qy, qy, qz * qz));
    return s;
}

// Same with the squared norms A = hm.hm and Cc = hs.hs supplied by the caller (reused rays).
template <typename T>
__device__ __forceinline__ PairSol<T> pair_solve_a(const V3<T>& hm, T A, const V3<T>& hs, T Cc, const V3<T>& d) {
    const T B = dot3(hm, hs), D = dot3(hm, d), E = dot3(hs, d);
    PairSol<T> s;
    s.det = fma(A, Cc, -(B * B));
    s.n0 = fma(Cc, D, -(B * E));
    s.n1 = fma(A, E, -(B * D));
    const T qx = fma(hm.x, s.n0, fma(hs.x, s.n1, -(d.x * s.det)));
    const T qy = fma(hm.y, s.n0, fma(hs.y, s.n1, -(d.y * s.det)));
    const T qz = fma(hm.z, s.n0, fma(hs.z, s.n1, -(d.z * s.det)));
    s.qq = fma(qx, qx, fma(qy, qy, qz * qz));
    return s;
}

// Cross-product form used by the single-person kernel: with n = hm x hs, |n|^2 = det and
// d.n = +-dist*|n|, so q.q = det*(d.n)^2 without forming q.  The normal-equation part (det, n0, n1)
// is well conditioned; d.n is the difference of nearly equal terms when the rays nearly intersect,
// which is why the mixed mode evaluates cross_dot in float64.
template <typename T>
struct PairSolN {
    T det, n0, n1;
};
template <typename T>
__device__ __forceinline__ PairSolN<T> pair_solve_n(const V3<T>& hm, T A, const V3<T>& hs, T Cc, const V3<T>& d) {
    const T B = dot3(hm, hs), D = dot3(hm, d), E = dot3(hs, d);
    PairSolN<T> s;
    s.det = fma(A, Cc, -(B * B));
    s.n0 = fma(Cc, D, -(B * E));
    s.n1 = fma(A, E, -(B * D));
    return s;
}
template <typename T>
__device__ __forceinline__ T cross_dot(const V3<T>& hm, const V3<T>& hs, const V3<T>& d) {
    V3<T> n;
    n.x = fma(hm.y, hs.z, -(hm.z * hs.y));
    n.y = fma(hm.z, hs.x, -(hm.x * hs.z));
    n.z = fma(hm.x, hs.y, -(hm.y * hs.x));
    return dot3(n, d);
}

// g = gated (sm+ss)*0.00025*rsqrt(q.q); the reference's score is 2*g*det.
template <typename T>
__device__ __forceinline__ T gated_g(const PairSol<T>& s, float sm, float ss, float kst_f, T dthr) {
    const T thr = dthr * s.det;
    const T r = rsqrt_t(s.qq);
    T g = ((T)sm + (T)ss) * (r * (T)0.00025);
    // same comparisons as the reference (strict; NaN distance is not gated)
    if (sm < kst_f || ss < kst_f || s.qq > thr * thr) g = (T)0;
    return g;
}

template <typename T>
__device__ __forceinline__ V3<T> pair_v(const PairSol<T>& s, const V3<T>& hm, const V3<T>& hs) {
    V3<T> v;
    v.x = fma(hm.x, s.n0, -(hs.x * s.n1));
    v.y = fma(hm.y, s.n0, -(hs.y * s.n1));
    v.z = fma(hm.z, s.n0, -(hs.z * s.n1));
    return v;
}

template <typename T>
__device__ __forceinline__ V3<T> pair_v(T n0, T n1, const V3<T>& hm, const V3<T>& hs) {
    V3<T> v;
    v.x = fma(hm.x, n0, -(hs.x * n1));
    v.y = fma(hm.y, n0, -(hs.y * n1));
    v.z = fma(hm.z, n0, -(hs.z * n1));
    return v;
}

// Explicit midpoint W = mid + v/(2 det) (needed for clustering centres and candidate output).
template <typename T>
__device__ __forceinline__ V3<T> pair_midpoint(const PairSol<T>& s, const V3<T>& hm, const V3<T>& hs,
                                               const V3<T>& mid) {
    const V3<T> v = pair_v(s, hm, hs);
    const T h = (T)0.5 * rcp_t(s.det);
    V3<T> w;
    w.x = fma(v.x, h, mid.x);
    w.y = fma(v.y, h, mid.y);
    w.z = fma(v.z, h, mid.z);
    return w;
}

// End of the per-joint fuse (reference triangulation.py:142-148): S = sum of member scores,
// (X,Y,Z) = sum of score*point.  Returns the keypoint score S/n; a zero S leaves (0,0,0), 0 (Q6/Q7).
template <typename T>
__device__ __forceinline__ T finish_joint(T S, int n, T& X, T& Y, T& Z) {
    if (S == (T)0) {
        X = Y = Z = (T)0;
        return (T)0;
    }
    const T rS = rcp_t(S);
    X *= rS;
    Y *= rS;
    Z *= rS;
    return S * rcp_t((T)n);
}

// Back-projection f = (R K^-1) [u v 1]^T (reference snowvision/camera.py:240-244); M row-major 3x3.
template <typename T>
__device__ __forceinline__ V3<T> back_project(const T* __restrict__ M, T u, T v) {
    V3<T> h;
    h.x = fma(M[0], u, fma(M[1], v, M[2]));
    h.y = fma(M[3], u, fma(M[4], v, M[5]));
    h.z = fma(M[6], u, fma(M[7], v, M[8]));
    return h;
}

// Same with the rows of M padded to 4 entries (16-byte aligned rows in the constant bank).
template <typename T>
__device__ __forceinline__ V3<T> back_project4(const T* __restrict__ M, T u, T v) {
    V3<T> h;
    h.x = fma(M[0], u, fma(M[1], v, M[2]));
    h.y = fma(M[4], u, fma(M[5], v, M[6]));
    h.z = fma(M[8], u, fma(M[9], v, M[10]));
    return h;
}

// Hide how a value was derived so that the compiler keeps it in a register instead of
// re-deriving it inside a hot loop (per-tile base pointers, constant addends).
__device__ __forceinline__ void keep_in_register(float& x) { asm volatile("" : "+f"(x)); }
__device__ __forceinline__ void keep_in_register(double& x) { asm volatile("" : "+d"(x)); }
template <typename P>
__device__ __forceinline__ void keep_in_register(P*& p) { asm volatile("" : "+l"(p)); }

// h = M[:,0]*u + M[:,1]*v + addend, the addend M[:,2] passed in registers (rows of M padded to 4)
template <typename T>
__device__ __forceinline__ V3<T> back_project4r(const T* __restrict__ M, const T* addend, T u, T v) {
    V3<T> h;
    h.x = fma(M[0], u, fma(M[1], v, addend[0]));
    h.y = fma(M[4], u, fma(M[5], v, addend[1]));
    h.z = fma(M[8], u, fma(M[9], v, addend[2]));
    return h;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- mbarrier + 1-D bulk async copy (TMA) -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 4- and 8-byte asynchronous global->shared copies (LDGSTS): a lane's next inputs travel while the current ones are
// solved, without holding registers; wait_all makes the thread's own copies visible to itself
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) { cp_async4(smem_u32(dst), src); }
__device__ __forceinline__ void cp_async8(void* dst, const void* src) { cp_async8(smem_u32(dst), src); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

}  // namespace snowtri
