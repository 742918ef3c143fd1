/*
 * CPU oracle, C restatement (float64, OpenMP) of the force path for sizes the NumPy oracle cannot reach in seconds.
 *
 * *** TEST INFRASTRUCTURE, NOT PRODUCT CODE *** - built into oracle/_build/libtm_oracle.so by oracle/build_oracle.py,
 * loaded only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * Same mathematics as oracle/tm_oracle.py (which is pinned to the reference's Python functions through
 * tests/golden/): nonbonded energy per timemachine/potentials/nonbonded.py:221-339 with the switch of :23-39 and the
 * minimum image of jax_utils.py:37-44 (floor(d/b + 0.5)); analytic gradients per the reference kernels
 * (timemachine/cpp/src/kernels/k_nonbonded_common.cuh:72-94,184-246); harmonic bond / angle per
 * timemachine/potentials/bonded.py:34-138; BAOAB per timemachine/integrator.py:137-144.
 * tests/test_oracle_c.py checks this file against the NumPy oracle to 1e-10.
 *
 * The all-pairs loop is O(N^2) like the reference's dense JAX path (pairwise_distances, jax_utils.py:144-181), with
 * an early distance rejection; rows are distributed over OpenMP threads and reduced per thread.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SWITCH_CUTOFF 1.2

static inline double min_image(double d, double b) { return d - b * floor(d / b + 0.5); }

static inline void pair_terms(
    double d, double qi, double qj, double si, double sj, double ei, double ej, double beta, double qs, double ls,
    double *u, double *du_dd, double *dq_i, double *dq_j, double *dsig, double *deps_i, double *deps_j) {
    const double inv_d = 1.0 / d;
    const double bd = beta * d;
    const double e = erfc(bd);
    const double de = -2.0 * beta / sqrt(M_PI) * exp(-bd * bd);
    double s = 0.0, ds = 0.0;
    if (d < SWITCH_CUTOFF) {
        const double r = d / SWITCH_CUTOFF;
        const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
        const double arg = 0.5 * M_PI * r8;
        const double c = cos(arg), sn = sin(arg);
        s = c * c * c;
        ds = -12.0 * M_PI * (r8 / d) * sn * c * c; /* d/dd cos^3(pi/2 (d/c)^8) */
    }
    const double damp = e * s;
    const double ddamp = e * ds + de * s;
    const double qij = qs * qi * qj;
    *u = qij * damp * inv_d;
    *du_dd = qij * (ddamp * inv_d - damp * inv_d * inv_d);
    *dq_i = qs * qj * damp * inv_d;
    *dq_j = qs * qi * damp * inv_d;
    *dsig = 0.0;
    *deps_i = 0.0;
    *deps_j = 0.0;
    if (ei != 0.0 && ej != 0.0) {
        const double eps = ei * ej;
        const double sig = si + sj;
        const double s1 = sig * inv_d;
        const double s2 = s1 * s1;
        const double s6 = s2 * s2 * s2;
        const double s12 = s6 * s6;
        *u += ls * 4.0 * eps * (s12 - s6);
        *du_dd += -ls * 24.0 * eps * inv_d * (2.0 * s12 - s6);
        *dsig = ls * 24.0 * eps * (2.0 * s12 - s6) / sig;
        *deps_i = ls * 4.0 * (s12 - s6) * ej;
        *deps_j = ls * 4.0 * (s12 - s6) * ei;
    }
}

/* rows x cols block.  triangular != 0: rows == cols == idxs and only pairs with row position < col position count.
 * du_dx [N,3], du_dp [N,4] are ACCUMULATED into (may be NULL); returns the energy. */
double tmo_nonbonded_block(
    int N, const double *x, const double *params, const double *box, const int *rows, int n_rows, const int *cols,
    int n_cols, int triangular, double beta, double cutoff, double *du_dx, double *du_dp) {
    const double bx = box[0], by = box[4], bz = box[8];
    const double c2 = cutoff * cutoff;
    double u_total = 0.0;
    int n_threads = 1;
#ifdef _OPENMP
    n_threads = omp_get_max_threads();
#endif
    double *fx = du_dx ? (double *)calloc((size_t)n_threads * N * 3, sizeof(double)) : NULL;
    double *fp = du_dp ? (double *)calloc((size_t)n_threads * N * 4, sizeof(double)) : NULL;
#pragma omp parallel reduction(+ : u_total)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        double *mx = fx ? fx + (size_t)tid * N * 3 : NULL;
        double *mp = fp ? fp + (size_t)tid * N * 4 : NULL;
#pragma omp for schedule(dynamic, 8)
        for (int a = 0; a < n_rows; a++) {
            const int i = rows[a];
            const double xi = x[i * 3], yi = x[i * 3 + 1], zi = x[i * 3 + 2];
            const double qi = params[i * 4], si = params[i * 4 + 1], ei = params[i * 4 + 2], wi = params[i * 4 + 3];
            for (int b = triangular ? a + 1 : 0; b < n_cols; b++) {
                const int j = cols[b];
                if (j == i) {
                    continue;
                }
                const double dx = min_image(xi - x[j * 3], bx);
                const double dy = min_image(yi - x[j * 3 + 1], by);
                const double dz = min_image(zi - x[j * 3 + 2], bz);
                const double dw = wi - params[j * 4 + 3];
                const double d2 = dx * dx + dy * dy + dz * dz + dw * dw;
                if (!(d2 < c2)) {
                    continue;
                }
                const double d = sqrt(d2);
                double u, du_dd, dq_i, dq_j, dsig, deps_i, deps_j;
                pair_terms(d, qi, params[j * 4], si, params[j * 4 + 1], ei, params[j * 4 + 2], beta, 1.0, 1.0, &u, &du_dd,
                           &dq_i, &dq_j, &dsig, &deps_i, &deps_j);
                u_total += u;
                const double pref = du_dd / d;
                if (mx) {
                    mx[i * 3] += pref * dx;
                    mx[i * 3 + 1] += pref * dy;
                    mx[i * 3 + 2] += pref * dz;
                    mx[j * 3] -= pref * dx;
                    mx[j * 3 + 1] -= pref * dy;
                    mx[j * 3 + 2] -= pref * dz;
                }
                if (mp) {
                    mp[i * 4] += dq_i;
                    mp[j * 4] += dq_j;
                    mp[i * 4 + 1] += dsig;
                    mp[j * 4 + 1] += dsig;
                    mp[i * 4 + 2] += deps_i;
                    mp[j * 4 + 2] += deps_j;
                    mp[i * 4 + 3] += pref * dw;
                    mp[j * 4 + 3] -= pref * dw;
                }
            }
        }
    }
    if (fx) {
        for (int t = 0; t < n_threads; t++) {
            for (size_t k = 0; k < (size_t)N * 3; k++) {
                du_dx[k] += fx[(size_t)t * N * 3 + k];
            }
        }
        free(fx);
    }
    if (fp) {
        for (int t = 0; t < n_threads; t++) {
            for (size_t k = 0; k < (size_t)N * 4; k++) {
                du_dp[k] += fp[(size_t)t * N * 4 + k];
            }
        }
        free(fp);
    }
    return u_total;
}

/* explicit pairs with (charge, lj) scales; sign = +1 (pair list) or -1 (exclusions); accumulates, returns energy */
double tmo_nonbonded_pairs(
    int N, const double *x, const double *params, const double *box, const int *pairs, const double *scales, int M,
    double sign, double beta, double cutoff, double *du_dx, double *du_dp) {
    (void)N;
    const double bx = box[0], by = box[4], bz = box[8];
    const double c2 = cutoff * cutoff;
    double u_total = 0.0;
    for (int m = 0; m < M; m++) {
        const int i = pairs[m * 2], j = pairs[m * 2 + 1];
        const double dx = min_image(x[i * 3] - x[j * 3], bx);
        const double dy = min_image(x[i * 3 + 1] - x[j * 3 + 1], by);
        const double dz = min_image(x[i * 3 + 2] - x[j * 3 + 2], bz);
        const double dw = params[i * 4 + 3] - params[j * 4 + 3];
        const double d2 = dx * dx + dy * dy + dz * dz + dw * dw;
        if (!(d2 < c2)) {
            continue;
        }
        const double d = sqrt(d2);
        double u, du_dd, dq_i, dq_j, dsig, deps_i, deps_j;
        pair_terms(d, params[i * 4], params[j * 4], params[i * 4 + 1], params[j * 4 + 1], params[i * 4 + 2],
                   params[j * 4 + 2], beta, scales[m * 2], scales[m * 2 + 1], &u, &du_dd, &dq_i, &dq_j, &dsig, &deps_i, &deps_j);
        u_total += sign * u;
        const double pref = sign * du_dd / d;
        if (du_dx) {
            du_dx[i * 3] += pref * dx;
            du_dx[i * 3 + 1] += pref * dy;
            du_dx[i * 3 + 2] += pref * dz;
            du_dx[j * 3] -= pref * dx;
            du_dx[j * 3 + 1] -= pref * dy;
            du_dx[j * 3 + 2] -= pref * dz;
        }
        if (du_dp) {
            du_dp[i * 4] += sign * dq_i;
            du_dp[j * 4] += sign * dq_j;
            du_dp[i * 4 + 1] += sign * dsig;
            du_dp[j * 4 + 1] += sign * dsig;
            du_dp[i * 4 + 2] += sign * deps_i;
            du_dp[j * 4 + 2] += sign * deps_j;
            du_dp[i * 4 + 3] += pref * dw;
            du_dp[j * 4 + 3] -= pref * dw;
        }
    }
    return u_total;
}

double tmo_harmonic_bond(int B, const double *x, const double *p, const int *idxs, double *du_dx) {
    double u = 0.0;
    for (int b = 0; b < B; b++) {
        const int i = idxs[b * 2], j = idxs[b * 2 + 1];
        const double dx = x[i * 3] - x[j * 3], dy = x[i * 3 + 1] - x[j * 3 + 1], dz = x[i * 3 + 2] - x[j * 3 + 2];
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        const double kb = p[b * 2], b0 = p[b * 2 + 1];
        const double db = r - b0;
        u += kb / 2 * db * db;
        if (du_dx) {
            const double g = b0 != 0.0 ? kb * db / r : kb;
            du_dx[i * 3] += g * dx;
            du_dx[i * 3 + 1] += g * dy;
            du_dx[i * 3 + 2] += g * dz;
            du_dx[j * 3] -= g * dx;
            du_dx[j * 3 + 1] -= g * dy;
            du_dx[j * 3 + 2] -= g * dz;
        }
    }
    return u;
}

double tmo_harmonic_angle(int A, const double *x, const double *p, const int *idxs, double *du_dx) {
    double u = 0.0;
    for (int t = 0; t < A; t++) {
        const int i = idxs[t * 3], j = idxs[t * 3 + 1], k = idxs[t * 3 + 2];
        const double ka = p[t * 3], a0 = p[t * 3 + 1], eps = p[t * 3 + 2];
        double a[4], b[4];
        for (int d = 0; d < 3; d++) {
            a[d] = x[i * 3 + d] - x[j * 3 + d];
            b[d] = x[k * 3 + d] - x[j * 3 + d];
        }
        a[3] = eps;
        b[3] = eps;
        double aa = 0, bb = 0, ab = 0;
        for (int d = 0; d < 4; d++) {
            aa += a[d] * a[d];
            bb += b[d] * b[d];
            ab += a[d] * b[d];
        }
        const double na = sqrt(aa), nb = sqrt(bb);
        double hi = 0, lo = 0;
        for (int d = 0; d < 4; d++) {
            const double m = nb * a[d] - na * b[d];
            const double s = nb * a[d] + na * b[d];
            hi += m * m;
            lo += s * s;
        }
        const double theta = 2 * atan2(sqrt(hi), sqrt(lo));
        const double delta = theta - a0;
        u += ka / 2 * delta * delta;
        if (du_dx) {
            double aab[4], bba[4], n1 = 0, n2 = 0;
            for (int d = 0; d < 4; d++) {
                aab[d] = a[d] * ab - b[d] * aa;
                bba[d] = b[d] * ab - a[d] * bb;
                n1 += aab[d] * aab[d];
                n2 += bba[d] * bba[d];
            }
            n1 = sqrt(n1);
            n2 = sqrt(n2);
            const double pref = ka * delta;
            for (int d = 0; d < 3; d++) {
                const double gi = n1 == 0 ? 0 : pref / na * aab[d] / n1;
                const double gk = n2 == 0 ? 0 : pref / nb * bba[d] / n2;
                du_dx[i * 3 + d] += gi;
                du_dx[k * 3 + d] += gk;
                du_dx[j * 3 + d] -= gi + gk;
            }
        }
    }
    return u;
}

/* one BAOAB step in float64 (integrator.py:137-144); force = -du_dx */
void tmo_baoab(int N, double *x, double *v, const double *du_dx, double ca, const double *cb, const double *cc,
               double dt, const double *noise) {
    for (int i = 0; i < N; i++) {
        for (int d = 0; d < 3; d++) {
            const int q = i * 3 + d;
            const double v_mid = v[q] - cb[i] * du_dx[q];
            const double v_new = ca * v_mid + cc[i] * noise[q];
            x[q] += 0.5 * dt * (v_mid + v_new);
            v[q] = v_new;
        }
    }
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline asks for all host cores explicitly. */
void tmo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) {
        omp_set_num_threads(n);
    }
#else
    (void)n;
#endif
}

int tmo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
