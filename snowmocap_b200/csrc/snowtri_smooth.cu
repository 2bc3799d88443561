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
// Arithmetic: float64 like the reference.  (x - xp)/T and (...)/k2 are evaluated as multiplications by the
// reciprocals computed once on the host (1 ulp per step away from a true division, damped by the filter).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "snowtri_internal.h"

struct snowtri_smooth_state {
    int device, P, J;
    double f, z, r;
    double* d_state;  // [0] initialised, [1] n0, then xp, y, yd as (P, J, 3) each
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

constexpr int kSmoothAhead = 4;  // frames of input prefetched ahead of the recurrence

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
    V buf[kSmoothAhead];
    int nb[kSmoothAhead];
#pragma unroll
    for (int u = 0; u < kSmoothAhead; ++u)
        if (u < a.F) {
            buf[u] = pts[(size_t)u * stride + base];
            nb[u] = a.nout[u];
        }
    for (int t0 = 0; t0 < a.F; t0 += kSmoothAhead) {
#pragma unroll
        for (int u = 0; u < kSmoothAhead; ++u) {
            const int t = t0 + u;
            if (t >= a.F) break;
            const V p = buf[u];
            const int n = min(max(nb[u], 0), a.Pout);
            if (t + kSmoothAhead < a.F) {  // refill this slot for frame t + kSmoothAhead
                buf[u] = pts[(size_t)(t + kSmoothAhead) * stride + base];
                nb[u] = a.nout[t + kSmoothAhead];
            }
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
    __syncthreads();
    if (lead) {
        a.state[0] = init ? 1.0 : 0.0;
        a.state[1] = (double)n0;
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
    free(s);
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
    if (f64) smooth_kernel<double4><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    else smooth_kernel<float4><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
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
