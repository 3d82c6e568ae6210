// icp_math.cuh -- small dense maths of the ICP path, usable on host and device.
//
// SE(3)/SO(3) helpers in the MRPT conventions the reference relies on
// (R = Rz(yaw) Ry(pitch) Rx(roll); se(3) vectors ordered (v, omega);
// LidarOdometry.cpp:321-327 uses Lie::SE<3>::log, cpp:272-275 TPose3D),
// the cyclic-Jacobi 3x3 symmetric eigen solver of the plane fit
//  and a column-pivoting Householder QR
// for the 6x6 Gauss-Newton system (row K / A.6).
//
// Everything here is compiled with -fmad=false: the plane-fit decisions must
// be bit-identical to a non-contracted IEEE evaluation.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b2
{
struct Pose
{
    double R[9];  // row-major
    double t[3];
};

B2_HD void mat3_mul(const double* A, const double* B, double* C)
{
    double T[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            T[i * 3 + j] = (A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j]) + A[i * 3 + 2] * B[6 + j];
#pragma unroll
    for (int i = 0; i < 9; i++) C[i] = T[i];
}

B2_HD void mat3_vec(const double* A, const double* v, double* o)
{
    const double a = (A[0] * v[0] + A[1] * v[1]) + A[2] * v[2];
    const double b = (A[3] * v[0] + A[4] * v[1]) + A[5] * v[2];
    const double c = (A[6] * v[0] + A[7] * v[1]) + A[8] * v[2];
    o[0] = a, o[1] = b, o[2] = c;
}

// T = Ta * Tb
B2_HD void pose_compose(const Pose& a, const Pose& b, Pose& o)
{
    double tt[3];
    mat3_vec(a.R, b.t, tt);
    tt[0] += a.t[0], tt[1] += a.t[1], tt[2] += a.t[2];
    mat3_mul(a.R, b.R, o.R);
    o.t[0] = tt[0], o.t[1] = tt[1], o.t[2] = tt[2];
}

// T = Ta^-1 * Tb
B2_HD void pose_inverse_compose(const Pose& a, const Pose& b, Pose& o)
{
    const double RaT[9] = {a.R[0], a.R[3], a.R[6], a.R[1], a.R[4], a.R[7], a.R[2], a.R[5], a.R[8]};
    const double d[3] = {b.t[0] - a.t[0], b.t[1] - a.t[1], b.t[2] - a.t[2]};
    double tt[3];
    mat3_vec(RaT, d, tt);
    mat3_mul(RaT, b.R, o.R);
    o.t[0] = tt[0], o.t[1] = tt[1], o.t[2] = tt[2];
}

B2_HD void pose_from_ypr(const double* p6, Pose& o)
{
    const double cy = cos(p6[3]), sy = sin(p6[3]);
    const double cp = cos(p6[4]), sp = sin(p6[4]);
    const double cr = cos(p6[5]), sr = sin(p6[5]);
    o.R[0] = cy * cp, o.R[1] = cy * sp * sr - sy * cr, o.R[2] = cy * sp * cr + sy * sr;
    o.R[3] = sy * cp, o.R[4] = sy * sp * sr + cy * cr, o.R[5] = sy * sp * cr - cy * sr;
    o.R[6] = -sp, o.R[7] = cp * sr, o.R[8] = cp * cr;
    o.t[0] = p6[0], o.t[1] = p6[1], o.t[2] = p6[2];
}

B2_HD void pose_to_ypr(const Pose& T, double* p6)
{
    p6[0] = T.t[0], p6[1] = T.t[1], p6[2] = T.t[2];
    const double cpitch = sqrt(T.R[0] * T.R[0] + T.R[3] * T.R[3]);
    const double pitch = atan2(-T.R[6], cpitch);
    double yaw, roll;
    if (cpitch < 1e-12)
    {
        roll = 0.0;
        yaw = atan2(-T.R[1], T.R[4]);
    }
    else
    {
        yaw = atan2(T.R[3], T.R[0]);
        roll = atan2(T.R[7], T.R[8]);
    }
    p6[3] = yaw, p6[4] = pitch, p6[5] = roll;
}

B2_HD void so3_coeffs(double th2, double& A, double& B, double& C)
{
    if (th2 < 1e-8)
    {
        A = 1.0 - th2 / 6.0;
        B = 0.5 - th2 / 24.0;
        C = 1.0 / 6.0 - th2 / 120.0;
    }
    else
    {
        const double th = sqrt(th2);
        const double s = sin(th), c = cos(th);
        A = s / th;
        B = (1.0 - c) / th2;
        C = (th - s) / (th2 * th);
    }
}

// exp: se(3) (v, w) -> SE(3)
B2_HD void se3_exp(const double* eps, Pose& o)
{
    const double wx = eps[3], wy = eps[4], wz = eps[5];
    const double th2 = (wx * wx + wy * wy) + wz * wz;
    double A, B, C;
    so3_coeffs(th2, A, B, C);
    const double W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    const double W2[9] = {wx * wx - th2, wx * wy, wx * wz, wx * wy, wy * wy - th2,
                          wy * wz,       wx * wz, wy * wz, wz * wz - th2};
    double V[9];
#pragma unroll
    for (int i = 0; i < 9; i++)
    {
        const double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        o.R[i] = (I + A * W[i]) + B * W2[i];
        V[i] = (I + B * W[i]) + C * W2[i];
    }
    mat3_vec(V, eps, o.t);
}

// log: SE(3) -> se(3) (v, w)
B2_HD void se3_log(const Pose& T, double* eps)
{
    const double* R = T.R;
    const double w[3] = {0.5 * (R[7] - R[5]), 0.5 * (R[2] - R[6]), 0.5 * (R[3] - R[1])};
    const double s = sqrt((w[0] * w[0] + w[1] * w[1]) + w[2] * w[2]);
    const double c = 0.5 * (((R[0] + R[4]) + R[8]) - 1.0);
    const double th = atan2(s, c);
    double om[3];
    if (s < 1e-8 && c > 0)
    {
        const double k = 1.0 + th * th / 6.0;
        om[0] = k * w[0], om[1] = k * w[1], om[2] = k * w[2];
    }
    else if (s < 1e-8)
    {  // rotation by ~pi: axis from the diagonal of (R + I)/2
        double a[3] = {sqrt(fmax(0.0, 0.5 * (R[0] + 1.0))), sqrt(fmax(0.0, 0.5 * (R[4] + 1.0))),
                       sqrt(fmax(0.0, 0.5 * (R[8] + 1.0)))};
        int m = 0;
        if (a[1] > a[m]) m = 1;
        if (a[2] > a[m]) m = 2;
        for (int i = 0; i < 3; i++)
            if (i != m && (R[m * 3 + i] + R[i * 3 + m]) < 0) a[i] = -a[i];
        om[0] = th * a[0], om[1] = th * a[1], om[2] = th * a[2];
    }
    else
    {
        const double k = th / s;
        om[0] = k * w[0], om[1] = k * w[1], om[2] = k * w[2];
    }
    const double th2 = (om[0] * om[0] + om[1] * om[1]) + om[2] * om[2];
    double D;
    if (th2 < 1e-8)
        D = 1.0 / 12.0 + th2 / 720.0;
    else
    {
        const double thn = sqrt(th2);
        D = (1.0 - (thn * sin(thn)) / (2.0 * (1.0 - cos(thn)))) / th2;
    }
    const double wx = om[0], wy = om[1], wz = om[2];
    const double W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    const double W2[9] = {wx * wx - th2, wx * wy, wx * wz, wx * wy, wy * wy - th2,
                          wy * wz,       wx * wz, wy * wz, wz * wz - th2};
    double Vi[9];
#pragma unroll
    for (int i = 0; i < 9; i++)
    {
        const double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        Vi[i] = (I - 0.5 * W[i]) + D * W2[i];
    }
    mat3_vec(Vi, T.t, eps);
    eps[3] = om[0], eps[4] = om[1], eps[5] = om[2];
}

// Cyclic Jacobi, symmetric 3x3 (full storage, destroyed). Eigenvalues
// ascending, eigenvectors in the columns of V. Fixed operation order: the
// plane-fit gates computed from it are compared bit-for-bit with the oracle.
B2_HD void jacobi3(double* A, double* ev, double* V)
{
#pragma unroll
    for (int i = 0; i < 9; i++) V[i] = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    double frob = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) frob += A[i] * A[i];
    const double tol = 1e-30 * frob;
    for (int sweep = 0; sweep < 30; sweep++)
    {
        double off = 0;
        off += A[1] * A[1];
        off += A[2] * A[2];
        off += A[5] * A[5];
        if (!(off > tol)) break;
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int q = p + 1; q < 3; q++)
            {
                const double apq = A[p * 3 + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
                const double tt =
                    (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0);
                const double s = tt * c;
                A[p * 3 + p] = A[p * 3 + p] - tt * apq;
                A[q * 3 + q] = A[q * 3 + q] + tt * apq;
                A[p * 3 + q] = A[q * 3 + p] = 0.0;
#pragma unroll
                for (int r = 0; r < 3; r++)
                {
                    if (r != p && r != q)
                    {
                        const double arp = A[r * 3 + p], arq = A[r * 3 + q];
                        const double nrp = c * arp - s * arq;
                        const double nrq = s * arp + c * arq;
                        A[r * 3 + p] = A[p * 3 + r] = nrp;
                        A[r * 3 + q] = A[q * 3 + r] = nrq;
                    }
                    const double vrp = V[r * 3 + p], vrq = V[r * 3 + q];
                    V[r * 3 + p] = c * vrp - s * vrq;
                    V[r * 3 + q] = s * vrp + c * vrq;
                }
            }
    }
    ev[0] = A[0], ev[1] = A[4], ev[2] = A[8];
    // ascending, strict '>' exchange network (0,1),(1,2),(0,1)
#pragma unroll
    for (int pass = 0; pass < 3; pass++)
    {
        const int j = (pass == 1) ? 1 : 0;
        if (ev[j] > ev[j + 1])
        {
            const double te = ev[j];
            ev[j] = ev[j + 1], ev[j + 1] = te;
#pragma unroll
            for (int r = 0; r < 3; r++)
            {
                const double tv = V[r * 3 + j];
                V[r * 3 + j] = V[r * 3 + j + 1], V[r * 3 + j + 1] = tv;
            }
        }
    }
}

// Cyclic Jacobi for a symmetric 4x4 (Horn's N matrix, row N / A.10):
// eigenvector of the LARGEST eigenvalue -> unit quaternion (w, x, y, z), w >= 0.
B2_HD void horn_quaternion(const double* S /* 3x3 cross-covariance, row-major */, double* q)
{
    const double Sxx = S[0], Sxy = S[1], Sxz = S[2], Syx = S[3], Syy = S[4], Syz = S[5],
                 Szx = S[6], Szy = S[7], Szz = S[8];
    double A[16] = {Sxx + Syy + Szz, Syz - Szy,       Szx - Sxz,        Sxy - Syx,
                    Syz - Szy,       Sxx - Syy - Szz, Sxy + Syx,        Szx + Sxz,
                    Szx - Sxz,       Sxy + Syx,       -Sxx + Syy - Szz, Syz + Szy,
                    Sxy - Syx,       Szx + Sxz,       Syz + Szy,        -Sxx - Syy + Szz};
    double V[16];
    for (int i = 0; i < 16; i++) V[i] = (i % 5 == 0) ? 1.0 : 0.0;
    double frob = 0;
    for (int i = 0; i < 16; i++) frob += A[i] * A[i];
    const double tol = 1e-30 * frob;
    for (int sweep = 0; sweep < 30; sweep++)
    {
        double off = 0;
        for (int p = 0; p < 4; p++)
            for (int r = p + 1; r < 4; r++) off += A[p * 4 + r] * A[p * 4 + r];
        if (!(off > tol)) break;
        for (int p = 0; p < 4; p++)
            for (int r = p + 1; r < 4; r++)
            {
                const double apq = A[p * 4 + r];
                if (apq == 0.0) continue;
                const double theta = (A[r * 4 + r] - A[p * 4 + p]) / (2.0 * apq);
                const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0);
                const double s = tt * c;
                A[p * 4 + p] = A[p * 4 + p] - tt * apq;
                A[r * 4 + r] = A[r * 4 + r] + tt * apq;
                A[p * 4 + r] = A[r * 4 + p] = 0.0;
                for (int k = 0; k < 4; k++)
                {
                    if (k != p && k != r)
                    {
                        const double akp = A[k * 4 + p], akr = A[k * 4 + r];
                        const double np_ = c * akp - s * akr;
                        const double nr_ = s * akp + c * akr;
                        A[k * 4 + p] = A[p * 4 + k] = np_;
                        A[k * 4 + r] = A[r * 4 + k] = nr_;
                    }
                    const double vkp = V[k * 4 + p], vkr = V[k * 4 + r];
                    V[k * 4 + p] = c * vkp - s * vkr;
                    V[k * 4 + r] = s * vkp + c * vkr;
                }
            }
    }
    // the oracle sorts ascending (stable, strict '>') and takes the last
    // column: that is the LAST index among equal maxima
    int best = 0;
    for (int i = 1; i < 4; i++)
        if (!(A[best * 5] > A[i * 5])) best = i;
    double qw = V[0 * 4 + best], qx = V[1 * 4 + best], qy = V[2 * 4 + best], qz = V[3 * 4 + best];
    const double qn = sqrt(((qw * qw + qx * qx) + qy * qy) + qz * qz);
    qw /= qn, qx /= qn, qy /= qn, qz /= qn;
    if (qw < 0) qw = -qw, qx = -qx, qy = -qy, qz = -qz;
    q[0] = qw, q[1] = qx, q[2] = qy, q[3] = qz;
}

B2_HD void quaternion_to_R(const double* q, double* R)
{
    const double qw = q[0], qx = q[1], qy = q[2], qz = q[3];
    R[0] = 1 - 2 * (qy * qy + qz * qz), R[1] = 2 * (qx * qy - qw * qz), R[2] = 2 * (qx * qz + qw * qy);
    R[3] = 2 * (qx * qy + qw * qz), R[4] = 1 - 2 * (qx * qx + qz * qz), R[5] = 2 * (qy * qz - qw * qx);
    R[6] = 2 * (qx * qz - qw * qy), R[7] = 2 * (qy * qz + qw * qx), R[8] = 1 - 2 * (qx * qx + qy * qy);
}

// Column-pivoting Householder QR solve, 6x6; rank-revealing like Eigen's
// default threshold. Returns the rank; rank-deficient systems get the basic
// solution (zeros in the dependent unknowns).
B2_HD int qr_solve6(const double* Ain, const double* bin, double* x)
{
    const int N = 6;
    double A[36], b[6];
    int perm[6];
    for (int i = 0; i < 36; i++) A[i] = Ain[i];
    for (int i = 0; i < 6; i++) b[i] = bin[i], perm[i] = i;
    int rank = N;
    double r00 = 0;
    for (int k = 0; k < N; k++)
    {
        int best = k;
        double bestn = -1;
        for (int j = k; j < N; j++)
        {
            double s = 0;
            for (int i = k; i < N; i++) s += A[i * N + j] * A[i * N + j];
            if (s > bestn) bestn = s, best = j;
        }
        if (best != k)
        {
            for (int i = 0; i < N; i++)
            {
                const double tmp = A[i * N + k];
                A[i * N + k] = A[i * N + best], A[i * N + best] = tmp;
            }
            const int tp = perm[k];
            perm[k] = perm[best], perm[best] = tp;
        }
        const double normx = sqrt(bestn);
        if (k == 0) r00 = normx;
        if (!(normx > 2.220446049250313e-16 * N * r00) || normx == 0.0)
        {
            rank = k;
            break;
        }
        double v[6];
        const double x0 = A[k * N + k];
        const double alpha = (x0 >= 0) ? -normx : normx;
        for (int i = 0; i < N; i++) v[i] = (i < k) ? 0.0 : A[i * N + k];
        v[k] = x0 - alpha;
        double vnorm2 = 0;
        for (int i = k; i < N; i++) vnorm2 += v[i] * v[i];
        if (vnorm2 > 0)
        {
            for (int j = k; j < N; j++)
            {
                double dot = 0;
                for (int i = k; i < N; i++) dot += v[i] * A[i * N + j];
                const double f = 2.0 * dot / vnorm2;
                for (int i = k; i < N; i++) A[i * N + j] -= f * v[i];
            }
            double dot = 0;
            for (int i = k; i < N; i++) dot += v[i] * b[i];
            const double f = 2.0 * dot / vnorm2;
            for (int i = k; i < N; i++) b[i] -= f * v[i];
        }
    }
    double y[6] = {0, 0, 0, 0, 0, 0};
    for (int i = rank - 1; i >= 0; i--)
    {
        double s = b[i];
        for (int j = i + 1; j < rank; j++) s -= A[i * N + j] * y[j];
        y[i] = s / A[i * N + i];
    }
    for (int i = 0; i < N; i++) x[perm[i]] = y[i];
    return rank;
}

B2_HD int inverse6(const double* A, double* Ainv);

// Cholesky solve of a symmetric positive definite 6x6 system. Every index is
// a compile-time constant after unrolling, so on the device the factor lives
// in registers (the pivoted QR above indexes dynamically and goes through
// local memory). Returns false when a pivot is not safely positive: the
// caller then falls back to the rank-revealing QR.
// Factor: L (lower, unit-free) with the reciprocals of its diagonal, so that the
// two triangular solves multiply instead of dividing (FP64 division and square
// root are long dependent instruction sequences on the device).
struct Chol6
{
    double L[6][6];
    double inv[6];
};

B2_HD bool chol_factor6(const double* H, Chol6& F)
{
    double dmax = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) dmax = fmax(dmax, H[i * 6 + i]);
    const double tiny = 1e-13 * dmax;
#pragma unroll
    for (int j = 0; j < 6; j++)
    {
        double s = H[j * 6 + j];
#pragma unroll
        for (int k = 0; k < j; k++) s -= F.L[j][k] * F.L[j][k];
        if (!(s > tiny)) return false;
        const double d = sqrt(s);
        const double inv = 1.0 / d;
        F.L[j][j] = d;
        F.inv[j] = inv;
#pragma unroll
        for (int i = j + 1; i < 6; i++)
        {
            double t = H[i * 6 + j];
#pragma unroll
            for (int k = 0; k < j; k++) t -= F.L[i][k] * F.L[j][k];
            F.L[i][j] = t * inv;
        }
    }
    return true;
}

B2_HD void chol_apply6(const Chol6& F, const double* b, double* x)
{
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++)
    {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= F.L[i][k] * y[k];
        y[i] = s * F.inv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--)
    {
        double s = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) s -= F.L[k][i] * x[k];
        x[i] = s * F.inv[i];
    }
}

B2_HD bool chol_solve6(const double* H, const double* b, double* x)
{
    Chol6 F;
    if (!chol_factor6(H, F)) return false;
    chol_apply6(F, b, x);
    return true;
}

// 6x6 solve used by the Gauss-Newton step: Cholesky when H is safely positive
// definite, else the rank-revealing QR (degenerate geometry)
B2_HD void solve6_spd(const double* H, const double* b, double* x)
{
    if (!chol_solve6(H, b, x)) qr_solve6(H, b, x);
}

// inverse of a symmetric 6x6: Cholesky column by column, QR on failure
B2_HD int inverse6_spd(const double* A, double* Ainv)
{
    Chol6 F;
    if (!chol_factor6(A, F)) return inverse6(A, Ainv);
#pragma unroll 1
    for (int c = 0; c < 6; c++)
    {
        double e[6] = {0, 0, 0, 0, 0, 0}, x[6];
        e[c] = 1.0;
        chol_apply6(F, e, x);
        for (int i = 0; i < 6; i++) Ainv[i * 6 + c] = x[i];
    }
    return 6;
}

B2_HD int inverse6(const double* A, double* Ainv)
{
    int rank = 6;
    for (int c = 0; c < 6; c++)
    {
        double e[6] = {0, 0, 0, 0, 0, 0}, x[6];
        e[c] = 1.0;
        const int r = qr_solve6(A, e, x);
        if (r < rank) rank = r;
        for (int i = 0; i < 6; i++) Ainv[i * 6 + c] = x[i];
    }
    return rank;
}

}  // namespace b2
