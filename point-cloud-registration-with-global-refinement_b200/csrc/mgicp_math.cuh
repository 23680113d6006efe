// mgicp_math.cuh -- per-point / per-correspondence fp64 math of the GICP path, usable from device
// code and (for the CPU-side unit probes in tests/) from host code.  Everything here is plain fp64 with
// FMA contraction disabled at compile time (-fmad=false) so decisions match the Open3D-CPU semantics.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MG_HD __host__ __device__ __forceinline__
#else
#define MG_HD inline
#endif

namespace mg {

struct V3 { double x, y, z; };

MG_HD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
MG_HD V3 cross(const V3 &a, const V3 &b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
MG_HD double dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// squared distance exactly as nanoflann's L2 metric evaluates it for dim 3: ((dx*dx)+dy*dy)+dz*dz
MG_HD double dist2(double ax, double ay, double az, double bx, double by, double bz) {
    double dx = ax - bx, dy = ay - by, dz = az - bz;
    return dx * dx + dy * dy + dz * dz;
}


// ---------------------------------------------------------------------------------------------
// Deterministic sin / cos / acos (fdlibm kernels, plain fp64 operations, <= 1 ulp from libm).
// glibc and CUDA's libdevice round these functions differently in the last bit; under the reference's
// L1 kernel the ICP iteration amplifies a 1-ulp difference to 1e-5..1e-3 m, so the engine and its CPU
// oracle both evaluate exactly this code (compiled without FMA contraction on both sides).
// ---------------------------------------------------------------------------------------------
MG_HD double det_ksin(double x, double y) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = x * x, v = z * x;
    double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
MG_HD double det_kcos(double x, double y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = x * x;
    double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    double ax = fabs(x);
    if (ax < 0.3) return 1.0 - (0.5 * z - (z * r - x * y));
    double qx = ax > 0.78125 ? 0.28125 : (double)(float)(0.25 * ax);
    double hz = 0.5 * z - qx, a = 1.0 - qx;
    return a - (hz - (z * r - x * y));
}
// argument reduction by multiples of pi/2 (two-part Cody-Waite; |x| stays below a few pi on this path)
MG_HD int det_reduce(double x, double &y0, double &y1) {
    const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11;
    double n = rint(x * invpio2);
    double r0 = x - n * pio2_1, w = n * pio2_1t;
    y0 = r0 - w;
    y1 = (r0 - y0) - w;
    return (int)n & 3;
}
MG_HD double det_sin(double x) {
    double y0, y1;
    int k = det_reduce(x, y0, y1);
    switch (k) {
        case 0: return det_ksin(y0, y1);
        case 1: return det_kcos(y0, y1);
        case 2: return -det_ksin(y0, y1);
        default: return -det_kcos(y0, y1);
    }
}
MG_HD double det_cos(double x) {
    double y0, y1;
    int k = det_reduce(x, y0, y1);
    switch (k) {
        case 0: return det_kcos(y0, y1);
        case 1: return -det_ksin(y0, y1);
        case 2: return -det_kcos(y0, y1);
        default: return det_ksin(y0, y1);
    }
}
MG_HD double det_acos_pq(double z) {
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01, pS2 = 2.01212532134862925881e-01,
                 pS3 = -4.00555345006794114027e-02, pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00, qS3 = -6.88283971605453293030e-01,
                 qS4 = 7.70381505559019352791e-02;
    double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    return p / q;
}
MG_HD double det_acos(double x) {   // |x| <= 1 (callers clamp)
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17, pi = 3.14159265358979311600e+00;
    if (x >= 1.0) return 0.0;
    if (x <= -1.0) return pi + 2.0 * pio2_lo;
    if (fabs(x) < 0.5) {
        double r = det_acos_pq(x * x);
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    if (x < 0) {
        double z = (1.0 + x) * 0.5, s = sqrt(z), r = det_acos_pq(z);
        double w = r * s - pio2_lo;
        return pi - 2.0 * (s + w);
    }
    double z = (1.0 - x) * 0.5, s = sqrt(z);
    double df = (double)(float)s;
    double c = (z - df * df) / (s + df);
    double r = det_acos_pq(z);
    double w = r * s + c;
    return 2.0 * (df + w);
}

// ---------------------------------------------------------------------------------------------
// Open3D utility::FastEigen3x3 (robust closed-form symmetric 3x3 eigen solver, Geometric Tools),
// as used by EstimateNormals(fast_normal_computation=true) behind estimate_normals
// (/root/reference/ALL_FUNCTIONS.py:301-302).  Sign conventions of the cross products are kept: the
// sign decides Open3D's GetRotationFromE1ToX branch (c < -0.99).  A = (a00,a01,a02,a11,a12,a22).
// ---------------------------------------------------------------------------------------------
MG_HD V3 eigvec0(const double A[6], double ev) {
    V3 r0 = v3(A[0] - ev, A[1], A[2]);
    V3 r1 = v3(A[1], A[3] - ev, A[4]);
    V3 r2 = v3(A[2], A[4], A[5] - ev);
    V3 c01 = cross(r0, r1), c02 = cross(r0, r2), c12 = cross(r1, r2);
    double d0 = dot(c01, c01), d1 = dot(c02, c02), d2 = dot(c12, c12);
    double dmax = d0;
    int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    V3 c = imax == 0 ? c01 : (imax == 1 ? c02 : c12);
    double s = sqrt(imax == 0 ? d0 : (imax == 1 ? d1 : d2));
    return v3(c.x / s, c.y / s, c.z / s);
}

MG_HD V3 eigvec1(const double A[6], const V3 &e0, double ev1) {
    V3 U, V;
    if (fabs(e0.x) > fabs(e0.y)) {
        double inv = 1.0 / sqrt(e0.x * e0.x + e0.z * e0.z);
        U = v3(-e0.z * inv, 0.0, e0.x * inv);
    } else {
        double inv = 1.0 / sqrt(e0.y * e0.y + e0.z * e0.z);
        U = v3(0.0, e0.z * inv, -e0.y * inv);
    }
    V = cross(e0, U);
    V3 AU = v3(A[0] * U.x + A[1] * U.y + A[2] * U.z, A[1] * U.x + A[3] * U.y + A[4] * U.z, A[2] * U.x + A[4] * U.y + A[5] * U.z);
    V3 AV = v3(A[0] * V.x + A[1] * V.y + A[2] * V.z, A[1] * V.x + A[3] * V.y + A[4] * V.z, A[2] * V.x + A[4] * V.y + A[5] * V.z);
    double m00 = U.x * AU.x + U.y * AU.y + U.z * AU.z - ev1;
    double m01 = U.x * AV.x + U.y * AV.y + U.z * AV.z;
    double m11 = V.x * AV.x + V.y * AV.y + V.z * AV.z - ev1;
    double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
    if (a00 >= a11) {
        double mx = a00 > a01 ? a00 : a01;
        if (mx > 0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1.0 / sqrt(1.0 + m01 * m01); m01 *= m00; }
            else            { m00 /= m01; m01 = 1.0 / sqrt(1.0 + m00 * m00); m00 *= m01; }
            return v3(m01 * U.x - m00 * V.x, m01 * U.y - m00 * V.y, m01 * U.z - m00 * V.z);
        }
        return U;
    } else {
        double mx = a11 > a01 ? a11 : a01;
        if (mx > 0) {
            if (a11 >= a01) { m01 /= m11; m11 = 1.0 / sqrt(1.0 + m01 * m01); m01 *= m11; }
            else            { m11 /= m01; m01 = 1.0 / sqrt(1.0 + m11 * m11); m11 *= m01; }
            return v3(m11 * U.x - m01 * V.x, m11 * U.y - m01 * V.y, m11 * U.z - m01 * V.z);
        }
        return U;
    }
}

MG_HD V3 fast_eigen3x3(const double cov[6]) {
    double mc = cov[0];
#pragma unroll
    for (int i = 1; i < 6; ++i) if (cov[i] > mc) mc = cov[i];
    if (mc == 0) return v3(0, 0, 0);
    double A[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) A[i] = cov[i] / mc;
    double norm = A[1] * A[1] + A[2] * A[2] + A[4] * A[4];
    if (norm > 0) {
        double q = (A[0] + A[3] + A[5]) / 3.0;
        double b00 = A[0] - q, b11 = A[3] - q, b22 = A[5] - q;
        double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2.0) / 6.0);
        double c00 = b11 * b22 - A[4] * A[4];
        double c01 = A[1] * b22 - A[4] * A[2];
        double c02 = A[1] * A[4] - b11 * A[2];
        double det = (b00 * c00 - A[1] * c01 + A[2] * c02) / (p * p * p);
        double half_det = det * 0.5;
        half_det = fmin(fmax(half_det, -1.0), 1.0);
        double angle = det_acos(half_det) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        double beta2 = det_cos(angle) * 2.0;
        double beta0 = det_cos(angle + two_thirds_pi) * 2.0;
        double beta1 = -(beta0 + beta2);
        double e0 = q + p * beta0, e1 = q + p * beta1, e2 = q + p * beta2;
        if (half_det >= 0) {
            V3 v2 = eigvec0(A, e2);
            if (e2 < e0 && e2 < e1) return v2;
            V3 v1 = eigvec1(A, v2, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return cross(v1, v2);
        } else {
            V3 v0 = eigvec0(A, e0);
            if (e0 < e1 && e0 < e2) return v0;
            V3 v1 = eigvec1(A, v0, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return cross(v0, v1);
        }
    }
    if (cov[0] < cov[3] && cov[0] < cov[5]) return v3(1, 0, 0);
    if (cov[3] < cov[0] && cov[3] < cov[5]) return v3(0, 1, 0);
    return v3(0, 0, 1);
}

// cumulant accumulator of Open3D's ComputeCovariance; neighbours must be added in ascending
// distance order (the order nanoflann returns them in) for bit-parity of the sums.
struct Cumulants {
    double c[9];
    MG_HD void clear() {
#pragma unroll
        for (int i = 0; i < 9; ++i) c[i] = 0.0;
    }
    MG_HD void add(double x, double y, double z) {
        c[0] += x; c[1] += y; c[2] += z;
        c[3] += x * x; c[4] += x * y; c[5] += x * z;
        c[6] += y * y; c[7] += y * z; c[8] += z * z;
    }
    MG_HD void covariance(int n, double cov[6]) const {
        double m[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) m[i] = c[i] / (double)n;
        cov[0] = m[3] - m[0] * m[0];
        cov[1] = m[4] - m[0] * m[1];
        cov[2] = m[5] - m[0] * m[2];
        cov[3] = m[6] - m[1] * m[1];
        cov[4] = m[7] - m[1] * m[2];
        cov[5] = m[8] - m[2] * m[2];
    }
};

// normal from neighbour covariance with Open3D's post-processing (zero norm -> (0,0,1))
MG_HD V3 normal_from_cov(const double cov[6]) {
    V3 n = fast_eigen3x3(cov);
    double nn = sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
    if (nn == 0.0) return v3(0, 0, 1);
    return n;
}

// Effective plane normal of the GICP covariance C = R diag(eps,1,1) R^T built by Open3D's
// InitializePointCloudForGeneralizedICP from a unit normal n: R = GetRotationFromE1ToX(n) maps e1
// onto n, EXCEPT that it returns Identity when n.x < -0.99.  Hence C = I - (1-eps) m m^T with
// m = n, or m = e1 in the exceptional branch.
MG_HD V3 effective_normal(const V3 &n) { return (n.x < -0.99) ? v3(1.0, 0.0, 0.0) : n; }

// robust kernel weights (Open3D RobustKernel.cpp)
MG_HD double kernel_weight(int loss, double k, double r) {
    switch (loss) {
        case 0: return 1.0;
        case 1: return 1.0 / fabs(r);
        case 2: { double e = fabs(r); return k / (e > k ? e : k); }
        case 3: return 1.0 / (1.0 + (r / k) * (r / k));
        case 4: return k / ((k + r * r) * (k + r * r));
        case 5: { double e = fabs(r) / k; if (e > 1.0) e = 1.0; double q = 1.0 - e * e; return q * q; }
        default: return 1.0;
    }
}

// ---------------------------------------------------------------------------------------------
// One correspondence of TransformationEstimationForGeneralizedICP::ComputeTransformation:
//   M = C_t + C_s,  W = M^{-1/2} (principal root),  J = W [ -[p]x | I ],  r = W (p - q),
//   three rows, each weighted by the robust kernel; acc[0..20] += upper(J^T w J), acc[21..26] += J^T w r.
// Both covariances are I - k m m^T (k = 1 - eps, |m| = 1), so with u = a + b, v = a - b:
//   M = 2I - k (a a^T + b b^T) has eigenpairs (2 - k|u|^2/2, u), (2 - k|v|^2/2, v), (2, u x v) and
//   W = I/sqrt2 + g(|u|^2/2) u u^T / 2 + g(|v|^2/2) v v^T / 2,
//   g(x) = ((2-kx)^{-1/2} - 2^{-1/2}) / x = k / ( sqrt(2-kx) sqrt2 (sqrt(2-kx) + sqrt2) )   (no cancellation).
// ---------------------------------------------------------------------------------------------
MG_HD void gicp_weight_matrix(const V3 &a, const V3 &b, double k, double W[6]) {
    const double s2 = 1.4142135623730951;   // sqrt(2)
    const double is2 = 0.7071067811865476;  // 1/sqrt(2)
    V3 u = v3(a.x + b.x, a.y + b.y, a.z + b.z);
    V3 v = v3(a.x - b.x, a.y - b.y, a.z - b.z);
    double x1 = 0.5 * dot(u, u), x2 = 0.5 * dot(v, v);
    double l1 = sqrt(2.0 - k * x1), l2 = sqrt(2.0 - k * x2);
    double h1 = 0.5 * k / (l1 * s2 * (l1 + s2));
    double h2 = 0.5 * k / (l2 * s2 * (l2 + s2));
    W[0] = is2 + h1 * u.x * u.x + h2 * v.x * v.x;
    W[1] = h1 * u.x * u.y + h2 * v.x * v.y;
    W[2] = h1 * u.x * u.z + h2 * v.x * v.z;
    W[3] = is2 + h1 * u.y * u.y + h2 * v.y * v.y;
    W[4] = h1 * u.y * u.z + h2 * v.y * v.z;
    W[5] = is2 + h1 * u.z * u.z + h2 * v.z * v.z;
}

// `acc` is anything indexable that yields the 27 running sums (a plain array on the host, the kernel's per-thread column
// of shared memory on the device).  The three row contributions of a sum are combined in registers and added to the running
// sum once: 27 read-modify-write round trips to the accumulator per correspondence instead of 81 (ICP kernel: 54.3 -> 50.2 ms
// at 296 pairs).
template <class Acc>
MG_HD void gicp_accumulate(const V3 &p, const V3 &q, const V3 &ms, const V3 &mt, double k, int loss, double loss_k,
                           Acc &&acc) {
    double W[6];
    gicp_weight_matrix(mt, ms, k, W);
    const V3 d = v3(p.x - q.x, p.y - q.y, p.z - q.z);
    const V3 w0 = v3(W[0], W[1], W[2]), w1 = v3(W[1], W[3], W[4]), w2 = v3(W[2], W[4], W[5]);
    const V3 c0 = cross(p, w0), c1 = cross(p, w1), c2 = cross(p, w2);     // rows of W * (-[p]x)
    const double J0[6] = {c0.x, c0.y, c0.z, w0.x, w0.y, w0.z};
    const double J1[6] = {c1.x, c1.y, c1.z, w1.x, w1.y, w1.z};
    const double J2[6] = {c2.x, c2.y, c2.z, w2.x, w2.y, w2.z};
    const double r0 = dot(w0, d), r1 = dot(w1, d), r2 = dot(w2, d);
    double wt0, wt1, wt2;
    if (loss == 1) {
        // L1 (the reference's kernel, ALL_FUNCTIONS.py:284), tested first: the kernel is a run-time argument.
        // (Sharing one division among the three weights, 1 / (|r0| |r1| |r2|), and another between h1 and h2 was measured:
        // 1.5 % per iteration, not worth weights that are no longer the correctly rounded 1 / |r|.)
        wt0 = 1.0 / fabs(r0); wt1 = 1.0 / fabs(r1); wt2 = 1.0 / fabs(r2);
    } else {
        wt0 = kernel_weight(loss, loss_k, r0); wt1 = kernel_weight(loss, loss_k, r1); wt2 = kernel_weight(loss, loss_k, r2);
    }
    int a = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double jw0 = J0[i] * wt0, jw1 = J1[i] * wt1, jw2 = J2[i] * wt2;
#pragma unroll
        for (int j = i; j < 6; ++j) {
            const double t = fma(jw2, J2[j], fma(jw1, J1[j], jw0 * J0[j]));   // explicit FMA: same rounding on GPU and host
            acc[a] = acc[a] + t;
            ++a;
        }
        acc[21 + i] = acc[21 + i] + fma(jw2, r2, fma(jw1, r1, jw0 * r0));
    }
}

// ---------------------------------------------------------------------------------------------
// SolveJacobianSystemAndObtainExtrinsicMatrix: x = ldlt(JTJ).solve(-JTr) (no definiteness checks,
// Open3D defaults), then TransformVector6dToMatrix4d: R = Rz(x2) Ry(x1) Rx(x0), t = x3..5.
// sums: 21 upper-triangular JTJ terms (row-major) then 6 JTr terms.  U is row-major 4x4.
// ---------------------------------------------------------------------------------------------
MG_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // lower triangle, i >= j

MG_HD void ldlt_solve6(const double sums[27], double x[6]) {
    // Only the lower triangle is stored (21 doubles); every loop has constant bounds and is fully unrolled and every index
    // is a compile-time constant (the symmetric row/column swaps are selects over the candidate pivots), so the whole
    // factorisation stays in registers on the GPU.  The arithmetic is that of the textbook full-matrix form: pivot = first
    // largest |diagonal|, A[i][j] -= A[i][k] * (A[j][k] / d) for k < j <= i, L[i][k] = A[i][k] / d.
    double L[21], b[6];
    int perm[6];
    {
        int a = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) { L[tri(j, i)] = sums[a]; ++a; }      // sums: upper triangle row-major = lower column-major
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) perm[i] = i;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        int piv = k;
        double best = fabs(L[tri(k, k)]);
#pragma unroll
        for (int i = k + 1; i < 6; ++i) if (fabs(L[tri(i, i)]) > best) { best = fabs(L[tri(i, i)]); piv = i; }
#pragma unroll
        for (int c = k + 1; c < 6; ++c) {
            if (piv == c) {     // symmetric swap of rows/columns k and c
                { double t = L[tri(k, k)]; L[tri(k, k)] = L[tri(c, c)]; L[tri(c, c)] = t; }
#pragma unroll
                for (int j = 0; j < k; ++j) { double t = L[tri(k, j)]; L[tri(k, j)] = L[tri(c, j)]; L[tri(c, j)] = t; }
#pragma unroll
                for (int i = k + 1; i < c; ++i) { double t = L[tri(i, k)]; L[tri(i, k)] = L[tri(c, i)]; L[tri(c, i)] = t; }
#pragma unroll
                for (int i = c + 1; i < 6; ++i) { double t = L[tri(i, k)]; L[tri(i, k)] = L[tri(i, c)]; L[tri(i, c)] = t; }
                int t = perm[k]; perm[k] = perm[c]; perm[c] = t;
            }
        }
        const double d = L[tri(k, k)];
        double col[6];
#pragma unroll
        for (int i = k + 1; i < 6; ++i) col[i] = L[tri(i, k)];
#pragma unroll
        for (int i = k + 1; i < 6; ++i)
#pragma unroll
            for (int j = k + 1; j <= i; ++j) L[tri(i, j)] -= col[i] * (col[j] / d);
#pragma unroll
        for (int i = k + 1; i < 6; ++i) L[tri(i, k)] = col[i] / d;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) if (perm[i] == j) v = -sums[21 + j];
        b[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) b[i] -= L[tri(i, j)] * b[j];
#pragma unroll
    for (int i = 0; i < 6; ++i) b[i] /= L[tri(i, i)];
#pragma unroll
    for (int i = 5; i >= 0; --i)
#pragma unroll
        for (int j = i + 1; j < 6; ++j) b[i] -= L[tri(j, i)] * b[j];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) if (perm[i] == j) v = b[i];
        x[j] = v;
    }
}

MG_HD void vec6_to_mat4(const double x[6], double T[16]) {
    double ca = det_cos(x[0]), sa = det_sin(x[0]), cb = det_cos(x[1]), sb = det_sin(x[1]), cg = det_cos(x[2]), sg = det_sin(x[2]);
    // Rz*Ry then *Rx, spelled out in the same association order as the oracle ((Rz Ry) Rx)
    double zy[9] = {cg * cb, -sg, cg * sb, sg * cb, cg, sg * sb, -sb, 0.0, cb};
    T[0] = zy[0]; T[1] = zy[1] * ca + zy[2] * sa; T[2] = zy[1] * (-sa) + zy[2] * ca; T[3] = x[3];
    T[4] = zy[3]; T[5] = zy[4] * ca + zy[5] * sa; T[6] = zy[4] * (-sa) + zy[5] * ca; T[7] = x[4];
    T[8] = zy[6]; T[9] = zy[7] * ca + zy[8] * sa; T[10] = zy[7] * (-sa) + zy[8] * ca; T[11] = x[5];
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

MG_HD void mat4_mul(const double A[16], const double B[16], double C[16]) {
    double T[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
            T[4 * i + j] = s;
        }
#pragma unroll
    for (int i = 0; i < 16; ++i) C[i] = T[i];
}

// PointCloud::Transform on a point: homogeneous multiply and divide by w
MG_HD V3 transform_point(const double T[16], const V3 &p) {
    double nx = T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3] * 1.0;
    double ny = T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7] * 1.0;
    double nz = T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11] * 1.0;
    // affine T (last row exactly 0 0 0 1): w is exactly 1 and x / 1.0 == x, so the division is skipped (bit-identical)
    if (T[12] == 0.0 && T[13] == 0.0 && T[14] == 0.0 && T[15] == 1.0) return v3(nx, ny, nz);
    double nw = T[12] * p.x + T[13] * p.y + T[14] * p.z + T[15] * 1.0;
    return v3(nx / nw, ny / nw, nz / nw);
}
MG_HD V3 rotate_vec(const double T[16], const V3 &m) {
    return v3(T[0] * m.x + T[1] * m.y + T[2] * m.z, T[4] * m.x + T[5] * m.y + T[6] * m.z, T[8] * m.x + T[9] * m.y + T[10] * m.z);
}

}  // namespace mg
