/* TEST INFRASTRUCTURE ONLY -- plain-C FP64 restatement of the SnowMocap triangulation path.
 *
 * Used as (1) the bulk parity checker for sizes the Python loop oracle cannot finish and
 * (2) the "best-effort CPU" baseline line of bench.py.  Never linked into or called by the
 * product package.  Parity status: PINNED -- tests/test_oracle.py checks every entry point
 * against the golden vectors produced by the real reference (tests/golden/make_golden.py).
 *
 * Restated reference lines (paths relative to /root/reference):
 *   rays        f = R * (inv(K) * [u v 1]^T)                  snowvision/camera.py:240-244
 *   pair solve  S = inv(H^T H) H^T (ts - tm), H = [hm hs]     snowvision/triangulation.py:24-31
 *   candidates  all (mc<sc, pm, ps), gated scores, mean gate  snowvision/triangulation.py:50-93
 *   condense    greedy centre-joint clustering + fuse         snowvision/triangulation.py:95-162
 *
 * The 2x2 / 3x3 inverses use the adjugate formula instead of LAPACK getrf/getri; the
 * difference is O(1e-16 * cond) and is covered by the golden-vector tolerance (1e-10).
 *
 * Layouts: kpts (F,C,P,J,2) f32, scores (F,C,P,J) f32, counts (F,C) i32 (NULL = all P),
 *          K,R (C,3,3) f64 row-major, t (C,3) f64.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    double kst, ast, dthr, cond_tol, score_tol;
    int num_tol, center;
} oracle_params;

static void inv3(const double* m, double* o) {
    double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    o[0] = (e * i - f * h) / det; o[1] = (c * h - b * i) / det; o[2] = (b * f - c * e) / det;
    o[3] = (f * g - d * i) / det; o[4] = (a * i - c * g) / det; o[5] = (c * d - a * f) / det;
    o[6] = (d * h - e * g) / det; o[7] = (b * g - a * h) / det; o[8] = (a * e - b * d) / det;
}

/* camera.py:240-244 */
static void back_project(const double* Kinv, const double* R, double u, double v, double* f) {
    double w[3] = {u, v, 1.0}, q[3];
    for (int r = 0; r < 3; ++r) q[r] = Kinv[3 * r] * w[0] + Kinv[3 * r + 1] * w[1] + Kinv[3 * r + 2] * w[2];
    for (int r = 0; r < 3; ++r) f[r] = R[3 * r] * q[0] + R[3 * r + 1] * q[1] + R[3 * r + 2] * q[2];
}

/* triangulation.py:24-31 */
static double skew_ray_solve(const double* hm, const double* hs, const double* tm, const double* ts,
                             double* W) {
    double g00 = 0, g01 = 0, g11 = 0, r0 = 0, r1 = 0, d[3];
    for (int i = 0; i < 3; ++i) {
        d[i] = ts[i] - tm[i];
        g00 += hm[i] * hm[i]; g01 += hm[i] * hs[i]; g11 += hs[i] * hs[i];
    }
    for (int i = 0; i < 3; ++i) { r0 += hm[i] * d[i]; r1 += hs[i] * d[i]; }
    double det = g00 * g11 - g01 * g01;
    double i00 = g11 / det, i01 = -g01 / det, i11 = g00 / det;
    double S0 = i00 * r0 + i01 * r1, S1 = i01 * r0 + i11 * r1;
    double n2 = 0;
    for (int i = 0; i < 3; ++i) {
        double wm = hm[i] * S0 + tm[i], ws = -hs[i] * S1 + ts[i], e = wm - ws;
        n2 += e * e;
        W[i] = (wm + ws) / 2;
    }
    return sqrt(n2);
}

/* Candidate list of one frame (kept candidates only, materialised like the reference's lists). */
typedef struct {
    int n, cap, J;
    double* pts;   /* n x J x 3 */
    double* ks;    /* n x J */
    double* ps;    /* n */
    int* idx;      /* n x 4 : mc, sc, pm, ps */
} cand_list;

static int cand_reserve(cand_list* L, int need) {
    if (need <= L->cap) return 0;
    int cap = L->cap ? L->cap : 16;
    while (cap < need) cap *= 2;
    double* p = (double*)realloc(L->pts, sizeof(double) * (size_t)cap * L->J * 3);
    if (!p) return -1;
    L->pts = p;
    double* k = (double*)realloc(L->ks, sizeof(double) * (size_t)cap * L->J);
    if (!k) return -1;
    L->ks = k;
    double* s = (double*)realloc(L->ps, sizeof(double) * (size_t)cap);
    if (!s) return -1;
    L->ps = s;
    int* x = (int*)realloc(L->idx, sizeof(int) * (size_t)cap * 4);
    if (!x) return -1;
    L->idx = x;
    L->cap = cap;
    return 0;
}

static void cand_free(cand_list* L) { free(L->pts); free(L->ks); free(L->ps); free(L->idx); memset(L, 0, sizeof(*L)); }

/* triangulation.py:50-93 for one frame; rays is scratch of C*P*J*3 doubles. */
static int triangulate_frame(int C, int P, int J, const float* kpts, const float* scores, const int* counts,
                             const double* Kinv, const double* R, const double* t, const oracle_params* prm,
                             double* rays, cand_list* L) {
    L->n = 0;
    L->J = J;
    for (int c = 0; c < C; ++c) {
        int pc = counts ? counts[c] : P;
        for (int p = 0; p < pc; ++p)
            for (int j = 0; j < J; ++j) {
                const float* uv = kpts + (((size_t)c * P + p) * J + j) * 2;
                back_project(Kinv + 9 * c, R + 9 * c, (double)uv[0], (double)uv[1],
                             rays + (((size_t)c * P + p) * J + j) * 3);
            }
    }
    for (int mc = 0; mc < C - 1; ++mc)
        for (int sc = mc + 1; sc < C; ++sc) {
            int nm = counts ? counts[mc] : P, ns = counts ? counts[sc] : P;
            for (int pm = 0; pm < nm; ++pm)
                for (int ps = 0; ps < ns; ++ps) {
                    if (cand_reserve(L, L->n + 1)) return -1;
                    double* pts = L->pts + (size_t)L->n * J * 3;
                    double* ks = L->ks + (size_t)L->n * J;
                    double sum = 0;
                    for (int j = 0; j < J; ++j) {
                        size_t im = ((size_t)mc * P + pm) * J + j, is = ((size_t)sc * P + ps) * J + j;
                        double dist = skew_ray_solve(rays + im * 3, rays + is * 3, t + 3 * mc, t + 3 * sc, pts + 3 * j);
                        double sm = (double)scores[im], ss = (double)scores[is];
                        double score = ((sm + ss) / 2) / (dist * 1000);
                        if (sm < prm->kst || ss < prm->kst || dist > prm->dthr) score = 0;
                        ks[j] = score;
                        sum += score;
                    }
                    double avg = sum / J;
                    if (avg < prm->ast) continue;          /* NaN is kept (Q9) */
                    L->ps[L->n] = avg;
                    int* x = L->idx + 4 * L->n;
                    x[0] = mc; x[1] = sc; x[2] = pm; x[3] = ps;
                    L->n++;
                }
        }
    return 0;
}

/* triangulation.py:95-162 on materialised candidates.  Returns the number of emitted persons;
 * only the first Pout are written. */
static int condense_cands(int N, int J, const double* pts, const double* ks, const oracle_params* prm,
                          int Jout, int Pout, double* o_pts, double* o_ks, double* o_ps) {
    int nout = 0;
    if (N <= 1) return 0;
    char* absorbed = (char*)calloc((size_t)N, 1);
    int* members = (int*)malloc(sizeof(int) * (size_t)N);
    double* tmp_p = (double*)malloc(sizeof(double) * (size_t)Jout * 3);
    double* tmp_k = (double*)malloc(sizeof(double) * (size_t)Jout);
    for (int mc = 0; mc < N - 1; ++mc) {
        if (absorbed[mc]) continue;
        const double* cm = pts + ((size_t)mc * J + prm->center) * 3;
        int n = 0;
        members[n++] = mc;
        for (int sc = mc + 1; sc < N; ++sc) {
            if (absorbed[sc]) continue;
            const double* cs = pts + ((size_t)sc * J + prm->center) * 3;
            double dx = cm[0] - cs[0], dy = cm[1] - cs[1], dz = cm[2] - cs[2];
            if (sqrt(dx * dx + dy * dy + dz * dz) > prm->cond_tol) continue;
            absorbed[sc] = 1;
            members[n++] = sc;
        }
        if (n < prm->num_tol) continue;
        double ksum = 0;
        for (int j = 0; j < Jout; ++j) {
            double total = 0;
            for (int k = 0; k < n; ++k) total += ks[(size_t)members[k] * J + j];
            tmp_p[3 * j] = tmp_p[3 * j + 1] = tmp_p[3 * j + 2] = 0;
            tmp_k[j] = 0;
            if (total == 0.0) continue;
            double x = 0, y = 0, z = 0;
            for (int k = 0; k < n; ++k) {
                double w = ks[(size_t)members[k] * J + j] / total;
                const double* q = pts + ((size_t)members[k] * J + j) * 3;
                x += q[0] * w; y += q[1] * w; z += q[2] * w;
            }
            tmp_p[3 * j] = x; tmp_p[3 * j + 1] = y; tmp_p[3 * j + 2] = z;
            tmp_k[j] = total / n;
            ksum += tmp_k[j];
        }
        double avg = ksum / Jout;
        if (avg < prm->score_tol) continue;
        if (nout < Pout) {
            memcpy(o_pts + (size_t)nout * Jout * 3, tmp_p, sizeof(double) * (size_t)Jout * 3);
            memcpy(o_ks + (size_t)nout * Jout, tmp_k, sizeof(double) * (size_t)Jout);
            o_ps[nout] = avg;
        }
        nout++;
    }
    free(absorbed); free(members); free(tmp_p); free(tmp_k);
    return nout;
}

static void fill_params(oracle_params* p, double kst, double ast, double dthr, double cond_tol, int num_tol,
                        double score_tol, int center) {
    p->kst = kst; p->ast = ast; p->dthr = dthr; p->cond_tol = cond_tol;
    p->num_tol = num_tol; p->score_tol = score_tol; p->center = center;
}

/* Batch of frames: rays -> candidates -> condense.  Outputs are dense FP64:
 * out_pts (F,Pout,Jout,3), out_ks (F,Pout,Jout), out_ps (F,Pout), nout (F) [true count, may exceed Pout],
 * ncand (F) [kept candidates] (may be NULL).  Returns 0, or -1 on allocation failure / bad args. */
int snow_oracle_fused(int F, int C, int P, int J, const float* kpts, const float* scores, const int* counts,
                      const double* K, const double* R, const double* t,
                      double kst, double ast, double dthr, double cond_tol, int num_tol, double score_tol,
                      int center, int Jout, int Pout,
                      double* out_pts, double* out_ks, double* out_ps, int* nout, int* ncand, int nthreads) {
    if (center < 0 || center >= J || Jout > J || Jout < 1) return -1;
    oracle_params prm;
    fill_params(&prm, kst, ast, dthr, cond_tol, num_tol, score_tol, center);
    double* Kinv = (double*)malloc(sizeof(double) * 9 * (size_t)C);
    for (int c = 0; c < C; ++c) inv3(K + 9 * c, Kinv + 9 * c);
    int fail = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        cand_list L;
        memset(&L, 0, sizeof(L));
        double* rays = (double*)malloc(sizeof(double) * (size_t)C * P * J * 3);
#pragma omp for schedule(dynamic, 1)
        for (int f = 0; f < F; ++f) {
            size_t fo = (size_t)f * C * P * J;
            if (triangulate_frame(C, P, J, kpts + fo * 2, scores + fo, counts ? counts + (size_t)f * C : NULL,
                                  Kinv, R, t, &prm, rays, &L)) { fail = 1; continue; }
            memset(out_pts + (size_t)f * Pout * Jout * 3, 0, sizeof(double) * (size_t)Pout * Jout * 3);
            memset(out_ks + (size_t)f * Pout * Jout, 0, sizeof(double) * (size_t)Pout * Jout);
            memset(out_ps + (size_t)f * Pout, 0, sizeof(double) * (size_t)Pout);
            nout[f] = condense_cands(L.n, J, L.pts, L.ks, &prm, Jout, Pout,
                                     out_pts + (size_t)f * Pout * Jout * 3, out_ks + (size_t)f * Pout * Jout,
                                     out_ps + (size_t)f * Pout);
            if (ncand) ncand[f] = L.n;
        }
        free(rays);
        cand_free(&L);
    }
    free(Kinv);
    return fail ? -1 : 0;
}

/* One frame, Human_Triangulation only.  cand_* have room for Nmax candidates; returns the number of
 * kept candidates (may exceed Nmax, in which case only Nmax were written), or -1 on error. */
int snow_oracle_candidates(int C, int P, int J, const float* kpts, const float* scores, const int* counts,
                           const double* K, const double* R, const double* t,
                           double kst, double ast, double dthr, int Nmax,
                           double* cand_pts, double* cand_ks, double* cand_ps, int* cand_idx) {
    oracle_params prm;
    fill_params(&prm, kst, ast, dthr, 0, 0, 0, 0);
    double* Kinv = (double*)malloc(sizeof(double) * 9 * (size_t)C);
    for (int c = 0; c < C; ++c) inv3(K + 9 * c, Kinv + 9 * c);
    double* rays = (double*)malloc(sizeof(double) * (size_t)C * P * J * 3);
    cand_list L;
    memset(&L, 0, sizeof(L));
    int rc = triangulate_frame(C, P, J, kpts, scores, counts, Kinv, R, t, &prm, rays, &L);
    int n = L.n;
    if (!rc) {
        int m = n < Nmax ? n : Nmax;
        memcpy(cand_pts, L.pts, sizeof(double) * (size_t)m * J * 3);
        memcpy(cand_ks, L.ks, sizeof(double) * (size_t)m * J);
        memcpy(cand_ps, L.ps, sizeof(double) * (size_t)m);
        memcpy(cand_idx, L.idx, sizeof(int) * (size_t)m * 4);
    }
    free(rays); free(Kinv); cand_free(&L);
    return rc ? -1 : n;
}

/* Human_Triangulation_Condense on caller-supplied candidates (N,J,3)/(N,J). Returns emitted count. */
int snow_oracle_condense(int N, int J, const double* pts, const double* ks,
                         double cond_tol, int num_tol, double score_tol, int center, int Jout, int Pout,
                         double* out_pts, double* out_ks, double* out_ps) {
    if (Jout > J || Jout < 1) return -1;
    if (N > 0 && (center < 0 || center >= J)) return -1;
    oracle_params prm;
    fill_params(&prm, 0, 0, 0, cond_tol, num_tol, score_tol, center);
    return condense_cands(N, J, pts, ks, &prm, Jout, Pout, out_pts, out_ks, out_ps);
}

/* Batched Skew_Ray_Solver: n independent pairs. hm,hs,tm,ts (n,3); dist (n), W (n,3). */
void snow_oracle_skew_ray(int n, const double* hm, const double* hs, const double* tm, const double* ts,
                          double* dist, double* W) {
    for (int i = 0; i < n; ++i) dist[i] = skew_ray_solve(hm + 3 * i, hs + 3 * i, tm + 3 * i, ts + 3 * i, W + 3 * i);
}

int snow_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}


/* Human_Triangulation_Smooth over consecutive frames (reference snowvision/triangulation.py:4-22,164-186,
 * called frame after frame by main.py:72-78).  Dense layout: pts (F,P,J,3) in/out, nout (F) persons present,
 * nsm (F) persons in the smoothed list.  state: [0]=initialised, [1]=n0 (persons of the first frame), then
 * xp, y, yd as (P,J,3) each.  The first frame passes through and seeds the followers; later frames update
 * persons k < min(nout, n0) only. */
int snow_oracle_smooth(int F, int P, int J, double* pts, const int* nout, int* nsm, double f, double z,
                       double r, double T, double* state) {
    const double pi = 3.14159265358979323846;
    const double k1 = z / (pi * f), k2 = 1 / ((2 * pi * f) * (2 * pi * f)), k3 = r * z / (2 * pi * f);
    const size_t N = (size_t)P * J * 3;
    double *xp = state + 2, *y = xp + N, *yd = y + N;
    for (int t = 0; t < F; ++t) {
        int n = nout[t] < 0 ? 0 : (nout[t] > P ? P : nout[t]);
        double* x = pts + (size_t)t * N;
        if (state[0] == 0.0) {
            for (size_t i = 0; i < (size_t)n * J * 3; ++i) { xp[i] = x[i]; y[i] = x[i]; yd[i] = 0.0; }
            state[0] = 1.0;
            state[1] = (double)n;
            nsm[t] = n;
            continue;
        }
        const int n0 = (int)state[1], m = n < n0 ? n : n0;
        for (size_t i = 0; i < (size_t)m * J * 3; ++i) {
            const double xd = (x[i] - xp[i]) / T;
            xp[i] = x[i];
            y[i] = y[i] + T * yd[i];
            yd[i] = yd[i] + T * (x[i] + k3 * xd - y[i] - k1 * yd[i]) / k2;
            x[i] = y[i];
        }
        nsm[t] = m;
    }
    return 0;
}
