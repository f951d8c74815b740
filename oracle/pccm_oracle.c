/*
 * Plain-C CPU oracle for the open-pcc-metric hot path.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never linked into the product.
 * Parity status: unpinned at the Open3D boundary (Open3D 0.18.0 is a third-party
 * dependency absent from /root/reference); anchored on the reference's call sites.
 *
 *   oracle_knn            exact brute-force k-NN, float64, squared distance
 *                         accumulated as (dx*dx + dy*dy) + dz*dz (nanoflann L2 order);
 *                         rows sorted by (d2, index)  <- canonical tie rule.
 *                         Restates cloud_pair.py:22 (search_knn_vector_3d) and the
 *                         neighbour queries inside estimate_normals() /
 *                         compute_nearest_neighbor_distance() (cloud_pair.py:61-64,109).
 *   oracle_normals        Open3D EstimateNormals(KNN, fast_normal_computation=true):
 *                         ComputeCovariance (nine cumulants / count) + FastEigen3x3.
 *
 * Build:  make -C oracle        (gcc -O2 -ffp-contract=off; this image has no libgomp,
 *                               so threads come from the caller: every entry point
 *                               takes a [begin, end) query range -- oracle/cnn.py
 *                               fans ranges out over a thread pool, ctypes drops the GIL)
 * -ffp-contract=off keeps every multiply and add separately rounded, as the x86-64
 * Open3D wheel (no FMA contraction in its baseline build) and numpy do.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* k-NN of every query in `pts`; out arrays are [nq][k]; entries beyond n are -1/inf. */
int oracle_knn(const double *pts, int64_t n, const double *q, int64_t begin, int64_t end, int k,
               int64_t *idx_out, double *d2_out) {
    if (k <= 0) return -1;
    for (int64_t i = begin; i < end; ++i) {
        double *bd = d2_out + i * (int64_t)k;
        int64_t *bi = idx_out + i * (int64_t)k;
        int cnt = 0;
        const double qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
        for (int64_t j = 0; j < n; ++j) {
            const double dx = qx - pts[3 * j];
            const double dy = qy - pts[3 * j + 1];
            const double dz = qz - pts[3 * j + 2];
            const double d = (dx * dx + dy * dy) + dz * dz;
            /* ascending j: an equal distance never displaces an earlier index */
            if (cnt == k && !(d < bd[k - 1])) continue;
            int p = cnt < k ? cnt : k - 1;
            while (p > 0 && bd[p - 1] > d) {
                bd[p] = bd[p - 1];
                bi[p] = bi[p - 1];
                --p;
            }
            bd[p] = d;
            bi[p] = j;
            if (cnt < k) ++cnt;
        }
        for (int p = cnt; p < k; ++p) {
            bd[p] = INFINITY;
            bi[p] = -1;
        }
    }
    return 0;
}

/* ---- Open3D FastEigen3x3 ------------------------------------------------ */
static void cross3(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static void eigenvector0(const double A[3][3], double eval0, double *out) {
    double row0[3] = {A[0][0] - eval0, A[0][1], A[0][2]};
    double row1[3] = {A[0][1], A[1][1] - eval0, A[1][2]};
    double row2[3] = {A[0][2], A[1][2], A[2][2] - eval0};
    double r0xr1[3], r0xr2[3], r1xr2[3];
    cross3(row0, row1, r0xr1);
    cross3(row0, row2, r0xr2);
    cross3(row1, row2, r1xr2);
    double d0 = dot3(r0xr1, r0xr1), d1 = dot3(r0xr2, r0xr2), d2 = dot3(r1xr2, r1xr2);
    double dmax = d0;
    int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    const double *v = imax == 0 ? r0xr1 : (imax == 1 ? r0xr2 : r1xr2);
    double s = sqrt(imax == 0 ? d0 : (imax == 1 ? d1 : d2));
    out[0] = v[0] / s; out[1] = v[1] / s; out[2] = v[2] / s;
}

static void eigenvector1(const double A[3][3], const double *evec0, double eval1, double *out) {
    double U[3], V[3];
    if (fabs(evec0[0]) > fabs(evec0[1])) {
        double inv_length = 1 / sqrt(evec0[0] * evec0[0] + evec0[2] * evec0[2]);
        U[0] = -evec0[2] * inv_length; U[1] = 0; U[2] = evec0[0] * inv_length;
    } else {
        double inv_length = 1 / sqrt(evec0[1] * evec0[1] + evec0[2] * evec0[2]);
        U[0] = 0; U[1] = evec0[2] * inv_length; U[2] = -evec0[1] * inv_length;
    }
    cross3(evec0, U, V);
    double AU[3] = {A[0][0] * U[0] + A[0][1] * U[1] + A[0][2] * U[2],
                    A[0][1] * U[0] + A[1][1] * U[1] + A[1][2] * U[2],
                    A[0][2] * U[0] + A[1][2] * U[1] + A[2][2] * U[2]};
    double AV[3] = {A[0][0] * V[0] + A[0][1] * V[1] + A[0][2] * V[2],
                    A[0][1] * V[0] + A[1][1] * V[1] + A[1][2] * V[2],
                    A[0][2] * V[0] + A[1][2] * V[1] + A[2][2] * V[2]};
    double m00 = U[0] * AU[0] + U[1] * AU[1] + U[2] * AU[2] - eval1;
    double m01 = U[0] * AV[0] + U[1] * AV[1] + U[2] * AV[2];
    double m11 = V[0] * AV[0] + V[1] * AV[1] + V[2] * AV[2] - eval1;
    double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
    if (a00 >= a11) {
        double mx = a00 > a01 ? a00 : a01;
        if (mx > 0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1 / sqrt(1 + m01 * m01); m01 *= m00; }
            else            { m00 /= m01; m01 = 1 / sqrt(1 + m00 * m00); m00 *= m01; }
            for (int i = 0; i < 3; ++i) out[i] = m01 * U[i] - m00 * V[i];
        } else { out[0] = U[0]; out[1] = U[1]; out[2] = U[2]; }
    } else {
        double mx = a11 > a01 ? a11 : a01;
        if (mx > 0) {
            if (a11 >= a01) { m01 /= m11; m11 = 1 / sqrt(1 + m01 * m01); m01 *= m11; }
            else            { m11 /= m01; m01 = 1 / sqrt(1 + m11 * m11); m11 *= m01; }
            for (int i = 0; i < 3; ++i) out[i] = m11 * U[i] - m01 * V[i];
        } else { out[0] = U[0]; out[1] = U[1]; out[2] = U[2]; }
    }
}

static void fast_eigen_3x3(const double cov[3][3], double *out) {
    double max_coeff = cov[0][0];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            if (cov[i][j] > max_coeff) max_coeff = cov[i][j];
    if (max_coeff == 0) { out[0] = out[1] = out[2] = 0; return; }
    double A[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = cov[i][j] / max_coeff;
    double norm = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    if (norm > 0) {
        double q = (A[0][0] + A[1][1] + A[2][2]) / 3;
        double b00 = A[0][0] - q, b11 = A[1][1] - q, b22 = A[2][2] - q;
        double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2) / 6);
        double c00 = b11 * b22 - A[1][2] * A[1][2];
        double c01 = A[0][1] * b22 - A[1][2] * A[0][2];
        double c02 = A[0][1] * A[1][2] - b11 * A[0][2];
        double det = (b00 * c00 - A[0][1] * c01 + A[0][2] * c02) / (p * p * p);
        double half_det = det * 0.5;
        half_det = fmin(fmax(half_det, -1.0), 1.0);
        double angle = acos(half_det) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        double beta2 = cos(angle) * 2;
        double beta0 = cos(angle + two_thirds_pi) * 2;
        double beta1 = -(beta0 + beta2);
        double e0 = q + p * beta0, e1 = q + p * beta1, e2 = q + p * beta2;
        double ev0[3], ev1[3], ev2[3];
        if (half_det >= 0) {
            eigenvector0(A, e2, ev2);
            if (e2 < e0 && e2 < e1) { memcpy(out, ev2, sizeof ev2); return; }
            eigenvector1(A, ev2, e1, ev1);
            if (e1 < e0 && e1 < e2) { memcpy(out, ev1, sizeof ev1); return; }
            cross3(ev1, ev2, out);
            return;
        }
        eigenvector0(A, e0, ev0);
        if (e0 < e1 && e0 < e2) { memcpy(out, ev0, sizeof ev0); return; }
        eigenvector1(A, ev0, e1, ev1);
        if (e1 < e0 && e1 < e2) { memcpy(out, ev1, sizeof ev1); return; }
        cross3(ev0, ev1, out);
        return;
    }
    double a00 = A[0][0] * max_coeff, a11 = A[1][1] * max_coeff, a22 = A[2][2] * max_coeff;
    out[0] = out[1] = out[2] = 0;
    if (a00 < a11 && a00 < a22) out[0] = 1;
    else if (a11 < a00 && a11 < a22) out[1] = 1;
    else out[2] = 1;
}

/* normals from a canonical k-NN index table nn[n][k] (entries < 0 = absent). */
int oracle_normals(const double *pts, int64_t begin, int64_t end, const int64_t *nn, int k, double *normals) {
    for (int64_t i = begin; i < end; ++i) {
        double c[9] = {0};
        int cnt = 0;
        for (int j = 0; j < k; ++j) {
            int64_t id = nn[i * (int64_t)k + j];
            if (id < 0) break;
            const double x = pts[3 * id], y = pts[3 * id + 1], z = pts[3 * id + 2];
            c[0] += x; c[1] += y; c[2] += z;
            c[3] += x * x; c[4] += x * y; c[5] += x * z;
            c[6] += y * y; c[7] += y * z; c[8] += z * z;
            ++cnt;
        }
        double cov[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        if (cnt >= 3) {
            for (int j = 0; j < 9; ++j) c[j] /= (double)cnt;
            cov[0][0] = c[3] - c[0] * c[0];
            cov[1][1] = c[6] - c[1] * c[1];
            cov[2][2] = c[8] - c[2] * c[2];
            cov[0][1] = cov[1][0] = c[4] - c[0] * c[1];
            cov[0][2] = cov[2][0] = c[5] - c[0] * c[2];
            cov[1][2] = cov[2][1] = c[7] - c[1] * c[2];
        }
        double nv[3];
        fast_eigen_3x3(cov, nv);
        if (sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]) == 0.0) { nv[0] = 0; nv[1] = 0; nv[2] = 1; }
        normals[3 * i] = nv[0]; normals[3 * i + 1] = nv[1]; normals[3 * i + 2] = nv[2];
    }
    return 0;
}
