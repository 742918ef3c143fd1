// rmsd_align(x1, x2): x2 moved rigidly onto x1 (reference rmsd_align.cpp:11-61, a host function there too: Eigen's
// JacobiSVD of the 3 x 3 correlation matrix).  Here the SVD is a one-sided Jacobi iteration written out for 3 x 3, so the
// library needs no linear-algebra dependency.  Host code only; f64.
#include "potential.hpp"

#include <cmath>
#include <utility>

namespace tmb {

namespace {

struct M3 {
    double m[3][3];
};

double det3(const M3 &a) {
    return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) - a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) +
           a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}

// c = u diag(s) v^T with s[0] >= s[1] >= s[2] >= 0, u and v orthogonal.  One-sided Jacobi: rotate pairs of columns of
// a = c v until they are mutually orthogonal; their norms are the singular values, the normalised columns are u.
void svd3(const M3 &c, M3 &u, double s[3], M3 &v) {
    M3 a = c;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) {
            v.m[i][j] = i == j ? 1.0 : 0.0;
        }
    }
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < 2; p++) {
            for (int q = p + 1; q < 3; q++) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int k = 0; k < 3; k++) {
                    alpha += a.m[k][p] * a.m[k][p];
                    beta += a.m[k][q] * a.m[k][q];
                    gamma += a.m[k][p] * a.m[k][q];
                }
                if (gamma == 0 || std::fabs(gamma) <= 1e-300) {
                    continue;
                }
                off = std::max(off, std::fabs(gamma) / std::sqrt(alpha * beta));
                const double zeta = (beta - alpha) / (2 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                const double cs = 1 / std::sqrt(1 + t * t);
                const double sn = cs * t;
                for (int k = 0; k < 3; k++) {
                    const double ap = a.m[k][p], aq = a.m[k][q];
                    a.m[k][p] = cs * ap - sn * aq;
                    a.m[k][q] = sn * ap + cs * aq;
                    const double vp = v.m[k][p], vq = v.m[k][q];
                    v.m[k][p] = cs * vp - sn * vq;
                    v.m[k][q] = sn * vp + cs * vq;
                }
            }
        }
        if (off < 1e-15) {
            break;
        }
    }
    int order[3] = {0, 1, 2};
    double norms[3];
    for (int j = 0; j < 3; j++) {
        norms[j] = std::sqrt(a.m[0][j] * a.m[0][j] + a.m[1][j] * a.m[1][j] + a.m[2][j] * a.m[2][j]);
    }
    for (int i = 0; i < 2; i++) {
        for (int j = 0; j < 2 - i; j++) {
            if (norms[order[j]] < norms[order[j + 1]]) {
                std::swap(order[j], order[j + 1]);
            }
        }
    }
    M3 vs;
    const double tiny = 1e-13 * std::max(norms[order[0]], 1e-300);
    int rank = 0;
    for (int j = 0; j < 3; j++) {
        const int src = order[j];
        s[j] = norms[src];
        for (int k = 0; k < 3; k++) {
            vs.m[k][j] = v.m[k][src];
            u.m[k][j] = norms[src] > tiny ? a.m[k][src] / norms[src] : 0.0;
        }
        if (norms[src] > tiny) {
            rank = j + 1;
        }
    }
    v = vs;
    // complete u to an orthonormal basis where c is rank deficient (planar / collinear / single-point inputs)
    if (rank == 0) {
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) {
                u.m[i][j] = i == j ? 1.0 : 0.0;
            }
        }
    } else {
        if (rank == 1) {
            // any unit vector orthogonal to column 0
            const int smallest = std::fabs(u.m[0][0]) <= std::fabs(u.m[1][0]) && std::fabs(u.m[0][0]) <= std::fabs(u.m[2][0]) ? 0 : (std::fabs(u.m[1][0]) <= std::fabs(u.m[2][0]) ? 1 : 2);
            double e[3] = {0, 0, 0};
            e[smallest] = 1;
            const double d = e[0] * u.m[0][0] + e[1] * u.m[1][0] + e[2] * u.m[2][0];
            double w[3], n = 0;
            for (int k = 0; k < 3; k++) {
                w[k] = e[k] - d * u.m[k][0];
                n += w[k] * w[k];
            }
            n = std::sqrt(n);
            for (int k = 0; k < 3; k++) {
                u.m[k][1] = w[k] / n;
            }
        }
        if (rank <= 2) {
            u.m[0][2] = u.m[1][0] * u.m[2][1] - u.m[2][0] * u.m[1][1];
            u.m[1][2] = u.m[2][0] * u.m[0][1] - u.m[0][0] * u.m[2][1];
            u.m[2][2] = u.m[0][0] * u.m[1][1] - u.m[1][0] * u.m[0][1];
        }
    }
}

} // namespace

void rmsd_align_host(int N, const double *x1, const double *x2, double *x2_aligned) {
    if (N <= 0) {
        return;
    }
    double c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
    for (int i = 0; i < N; i++) {
        for (int d = 0; d < 3; d++) {
            c1[d] += x1[i * 3 + d];
            c2[d] += x2[i * 3 + d];
        }
    }
    for (int d = 0; d < 3; d++) {
        c1[d] /= N;
        c2[d] /= N;
    }
    // correlation of the centred sets: c = x2c^T x1c
    M3 c = {};
    for (int i = 0; i < N; i++) {
        for (int r = 0; r < 3; r++) {
            const double a = x2[i * 3 + r] - c2[r];
            for (int k = 0; k < 3; k++) {
                c.m[r][k] += a * (x1[i * 3 + k] - c1[k]);
            }
        }
    }
    M3 u, v;
    double s[3];
    svd3(c, u, s, v);
    // a reflection is undone on the axis of the smallest singular value (rmsd_align.cpp:43-48)
    M3 vt;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) {
            vt.m[i][j] = v.m[j][i];
        }
    }
    if (det3(u) * det3(vt) < 0.0) {
        for (int i = 0; i < 3; i++) {
            u.m[i][2] = -u.m[i][2];
        }
    }
    M3 rot = {};
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) {
            for (int k = 0; k < 3; k++) {
                rot.m[i][j] += u.m[i][k] * vt.m[k][j];
            }
        }
    }
    for (int i = 0; i < N; i++) {
        const double a[3] = {x2[i * 3 + 0] - c2[0], x2[i * 3 + 1] - c2[1], x2[i * 3 + 2] - c2[2]};
        for (int j = 0; j < 3; j++) {
            x2_aligned[i * 3 + j] = a[0] * rot.m[0][j] + a[1] * rot.m[1][j] + a[2] * rot.m[2][j] + c1[j];
        }
    }
}

} // namespace tmb
