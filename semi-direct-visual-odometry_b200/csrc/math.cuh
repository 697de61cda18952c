// math.cuh -- small double-precision device helpers: SE3 (Sophus convention), pivoted LDLT.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace svo {

struct Pose {      // world -> camera, quaternion x y z w + translation
    double q[4];
    double t[3];
};

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3])
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// v + w*uv + qv x uv, uv = 2 qv x v
__device__ __forceinline__ void quat_rotate(const double q[4], const double v[3], double out[3])
{
    double uv[3], c2[3];
    cross3(q, v, uv);
    uv[0] *= 2.0;
    uv[1] *= 2.0;
    uv[2] *= 2.0;
    cross3(q, uv, c2);
#pragma unroll
    for (int i = 0; i < 3; i++) out[i] = v[i] + q[3] * uv[i] + c2[i];
}

__device__ __forceinline__ void quat_to_R(const double q[4], double R[9])
{
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z);
    R[1] = 2 * (x * y - z * w);
    R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);
    R[4] = 1 - 2 * (x * x + z * z);
    R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);
    R[7] = 2 * (y * z + x * w);
    R[8] = 1 - 2 * (x * x + y * y);
}

// pose <- pose * exp(-dx);  dx = (upsilon, omega)   [ImageAlignment::update, src/image_alignment.cpp:372-380]
static __device__ __noinline__ void pose_update_right_exp_neg(Pose& p, const double dx[6])
{
    double ups[3] = {-dx[0], -dx[1], -dx[2]};
    double om[3]  = {-dx[3], -dx[4], -dx[5]};
    const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
    const double th  = sqrt(th2);
    double imag, real, a, b;
    // ONE sincos and ONE reciprocal on the serial path of the solver (every other thread of the CTA waits for this):
    // sin th = 2 s c, cos th = 1 - 2 s^2 from the half angle; 1/th feeds sin(th/2)/th, (1-cos th)/th^2, (th-sin th)/th^3
    if (th2 < 1e-20) {
        const double th4 = th2 * th2;
        imag = 0.5 - th2 / 48.0 + th4 / 3840.0;
        real = 1.0 - th2 / 8.0 + th4 / 384.0;
        a    = 0.5;
        b    = 1.0 / 6.0;
    } else {
        double sh, ch;
        sincos(0.5 * th, &sh, &ch);
        const double r  = 1.0 / th;
        const double r2 = r * r;
        imag            = sh * r;
        real            = ch;
        const double s  = 2.0 * sh * ch;   // sin th
        a               = 2.0 * sh * sh * r2;  // (1 - cos th) / th^2
        b               = (th - s) * r2 * r;
    }
    double eq[4] = {imag * om[0], imag * om[1], imag * om[2], real};
    double c1[3], c2[3], et[3];
    cross3(om, ups, c1);
    cross3(om, c1, c2);
#pragma unroll
    for (int i = 0; i < 3; i++) et[i] = ups[i] + a * c1[i] + b * c2[i];
    // translation: t + R(q) * et
    double rt[3];
    quat_rotate(p.q, et, rt);
#pragma unroll
    for (int i = 0; i < 3; i++) p.t[i] += rt[i];
    // rotation: q * eq, renormalised
    const double ax = p.q[0], ay = p.q[1], az = p.q[2], aw = p.q[3];
    double r[4];
    r[0] = aw * eq[0] + ax * eq[3] + ay * eq[2] - az * eq[1];
    r[1] = aw * eq[1] + ay * eq[3] + az * eq[0] - ax * eq[2];
    r[2] = aw * eq[2] + az * eq[3] + ax * eq[1] - ay * eq[0];
    r[3] = aw * eq[3] - ax * eq[0] - ay * eq[1] - az * eq[2];
    const double ninv = 1.0 / sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
#pragma unroll
    for (int i = 0; i < 4; i++) p.q[i] = r[i] * ninv;
}

// inverse transform of a point: R^T (p - t)
__device__ __forceinline__ void pose_inv_act(const Pose& T, const double p[3], double out[3])
{
    double qi[4] = {-T.q[0], -T.q[1], -T.q[2], T.q[3]};
    double d[3]  = {p[0] - T.t[0], p[1] - T.t[1], p[2] - T.t[2]};
    quat_rotate(qi, d, out);
}

// camera centre in world: -R^T t  (Frame::cameraInWorld, src/frame.cpp:116-120)
__device__ __forceinline__ void pose_camera_in_world(const Pose& T, double C[3])
{
    double qi[4] = {-T.q[0], -T.q[1], -T.q[2], T.q[3]};
    double r[3];
    quat_rotate(qi, T.t, r);
    C[0] = -r[0];
    C[1] = -r[1];
    C[2] = -r[2];
}

// LDLT with symmetric pivoting on the largest |diagonal| and the D^-1 rule of Eigen's solve
// (pivots not above the smallest normal double give 0).  A: n x n row-major (n <= 6), destroyed.
template <int N>
__device__ __noinline__ void ldlt_solve(double* A, const double* b, double* x)
{
    int perm[N];
#pragma unroll
    for (int i = 0; i < N; i++) perm[i] = i;
    for (int k = 0; k < N; k++) {
        int piv    = k;
        double big = fabs(A[k * N + k]);
        for (int i = k + 1; i < N; i++) {
            const double v = fabs(A[i * N + i]);
            if (v > big) {
                big = v;
                piv = i;
            }
        }
        if (piv != k) {
            for (int j = 0; j < N; j++) {
                const double t = A[k * N + j];
                A[k * N + j]   = A[piv * N + j];
                A[piv * N + j] = t;
            }
            for (int j = 0; j < N; j++) {
                const double t = A[j * N + k];
                A[j * N + k]   = A[j * N + piv];
                A[j * N + piv] = t;
            }
            const int t = perm[k];
            perm[k]     = perm[piv];
            perm[piv]   = t;
        }
        double dk = A[k * N + k];
        for (int j = 0; j < k; j++) dk -= A[k * N + j] * A[k * N + j] * A[j * N + j];
        A[k * N + k] = dk;
        for (int i = k + 1; i < N; i++) {
            double s = A[i * N + k];
            for (int j = 0; j < k; j++) s -= A[i * N + j] * A[k * N + j] * A[j * N + j];
            A[i * N + k] = (fabs(dk) > 0.0) ? s / dk : s;
        }
    }
    double y[N];
    for (int i = 0; i < N; i++) y[i] = b[perm[i]];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < i; j++) y[i] -= A[i * N + j] * y[j];
    for (int i = 0; i < N; i++) y[i] = (fabs(A[i * N + i]) > 2.2250738585072014e-308) ? y[i] / A[i * N + i] : 0.0;
    for (int i = N - 1; i >= 0; i--)
        for (int j = i + 1; j < N; j++) y[i] -= A[j * N + i] * y[j];
    for (int i = 0; i < N; i++) x[perm[i]] = y[i];
}

// 1 / d for a normal, finite d: hardware seed (about 20 bits) and three Newton steps, no special-case branch; the result is
// within an ulp or two of the correctly rounded quotient.  On the serial path of the solver every cycle is paid by the whole
// CTA waiting at a barrier.
__device__ __forceinline__ double rcp_newton(double d)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x        = fma(x, e, x);
    e        = fma(-d, x, 1.0);
    x        = fma(x, e, x);
    e        = fma(-d, x, 1.0);
    x        = fma(x, e, x);
    return x;
}

// 1 / sqrt(x) for a normal, finite x > 0: hardware seed and three Newton steps (see rcp_newton)
__device__ __forceinline__ double rsqrt_newton(double x)
{
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double e = fma(-x * r, r, 1.0);
        r              = fma(0.5 * r, e, r);
    }
    return r;
}

// sin and cos of |x| <= 0.5 by their Taylor polynomials (next terms 2e-20 / 7e-19): the half angle of a Gauss-Newton step
// never leaves that range; sincos() beyond it
__device__ __forceinline__ void sincos_small(double x, double* s, double* c)
{
    if (fabs(x) > 0.5) {
        sincos(x, s, c);
        return;
    }
    const double z = x * x;
    double ps = -1.0 / 1307674368000.0;
    ps        = fma(ps, z, 1.0 / 6227020800.0);
    ps        = fma(ps, z, -1.0 / 39916800.0);
    ps        = fma(ps, z, 1.0 / 362880.0);
    ps        = fma(ps, z, -1.0 / 5040.0);
    ps        = fma(ps, z, 1.0 / 120.0);
    ps        = fma(ps, z, -1.0 / 6.0);
    *s        = fma(x * z, ps, x);
    double pc = -1.0 / 87178291200.0;
    pc        = fma(pc, z, 1.0 / 479001600.0);
    pc        = fma(pc, z, -1.0 / 3628800.0);
    pc        = fma(pc, z, 1.0 / 40320.0);
    pc        = fma(pc, z, -1.0 / 720.0);
    pc        = fma(pc, z, 1.0 / 24.0);
    pc        = fma(pc, z, -0.5);
    *c        = fma(z, pc, 1.0);
}

// pose_update_right_exp_neg for the single-CTA kernel, where every cycle of it is paid by 511 waiting threads: inlined (no
// stack traffic), reciprocal square roots by Newton steps, polynomial sin / cos.  Same formulas, results within a few ulp.
__device__ __forceinline__ void pose_update_right_exp_neg_fast(Pose& p, const double dx[6])
{
    const double ups[3] = {-dx[0], -dx[1], -dx[2]};
    const double om[3]  = {-dx[3], -dx[4], -dx[5]};
    const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
    double imag, real, a, b;
    if (th2 < 1e-20) {
        const double th4 = th2 * th2;
        imag = 0.5 - th2 / 48.0 + th4 / 3840.0;
        real = 1.0 - th2 / 8.0 + th4 / 384.0;
        a    = 0.5;
        b    = 1.0 / 6.0;
    } else {
        const double r  = rsqrt_newton(th2), th = th2 * r, r2 = r * r;
        double sh, ch;
        sincos_small(0.5 * th, &sh, &ch);
        imag           = sh * r;
        real           = ch;
        const double s = 2.0 * sh * ch;       // sin th
        a              = 2.0 * sh * sh * r2;  // (1 - cos th) / th^2
        b              = (th - s) * r2 * r;
    }
    const double eq[4] = {imag * om[0], imag * om[1], imag * om[2], real};
    double c1[3], c2[3], et[3];
    cross3(om, ups, c1);
    cross3(om, c1, c2);
#pragma unroll
    for (int i = 0; i < 3; i++) et[i] = ups[i] + a * c1[i] + b * c2[i];
    double rt[3];
    quat_rotate(p.q, et, rt);
#pragma unroll
    for (int i = 0; i < 3; i++) p.t[i] += rt[i];
    const double ax = p.q[0], ay = p.q[1], az = p.q[2], aw = p.q[3];
    double q[4];
    q[0] = aw * eq[0] + ax * eq[3] + ay * eq[2] - az * eq[1];
    q[1] = aw * eq[1] + ay * eq[3] + az * eq[0] - ax * eq[2];
    q[2] = aw * eq[2] + az * eq[3] + ax * eq[1] - ay * eq[0];
    q[3] = aw * eq[3] - ax * eq[0] - ay * eq[1] - az * eq[2];
    const double n2   = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const double ninv = fabs(n2 - 1.0) < 1e-9 ? fma(-0.5, n2, 1.5) : rsqrt_newton(n2);  // (first-order: its error is 3/8 (n2 - 1)^2)
#pragma unroll
    for (int i = 0; i < 4; i++) p.q[i] = q[i] * ninv;
}

// Fast path of the 6x6 solve: LDL^T without pivoting, fully unrolled so the whole factorisation lives in registers.
// E: upper triangle of the symmetric matrix, row-major (21 values); diag_add is added to the diagonal (LM damping).
// Returns false when a pivot is not safely positive (semi-definite / degenerate systems): the caller then falls back
// to ldlt_solve, which reproduces Eigen's pivoted LDLT including its zero-pivot rule.  For the well-conditioned SPD
// systems of the alignment both agree to rounding.
__device__ __forceinline__ bool ldlt6_nopivot(const double* E, double diag_add, const double* b, double* x)
{
    double A[6][6];
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = i; j < 6; j++, k++) A[j][i] = E[k];  // lower triangle
    }
    double scale = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        A[i][i] += diag_add;
        scale = fmax(scale, fabs(A[i][i]));
    }
    const double tiny = scale * 1e-13;
    double D[6], invD[6];
    bool ok = scale > 0.0 && isfinite(scale);
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double v[6];
        double d = A[j][j];
#pragma unroll
        for (int i = 0; i < j; i++) {
            v[i] = A[j][i] * D[i];
            d -= A[j][i] * v[i];
        }
        D[j]    = d;
        ok      = ok && (d > tiny);
        invD[j] = rcp_newton(d);  // (d > tiny > 0 is checked below; a failed check discards everything computed here)
#pragma unroll
        for (int r = j + 1; r < 6; r++) {
            double sacc = A[r][j];
#pragma unroll
            for (int i = 0; i < j; i++) sacc -= A[r][i] * v[i];
            A[r][j] = sacc * invD[j];
        }
    }
    if (!ok) return false;
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double sacc = b[i];
#pragma unroll
        for (int j = 0; j < i; j++) sacc -= A[i][j] * y[j];
        y[i] = sacc;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) y[i] *= invD[i];
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double sacc = y[i];
#pragma unroll
        for (int j = i + 1; j < 6; j++) sacc -= A[j][i] * x[j];
        x[i] = sacc;
    }
    return true;
}

// Optimizer::updateParameters, Nielsen branch (src/optimizer.cpp:449-466)
__device__ __forceinline__ bool nielsen_update(double pre, double cur, double& lambda, double& nu)
{
    const double rho = pre - cur;
    if (rho > 0.0) {
        const double t = 2.0 * rho - 1.0;
        lambda *= fmax(1.0 / 3.0, 1.0 - t * t * t);
        nu = 2.0;
        return true;
    }
    lambda *= nu;
    nu *= 2.0;
    return false;
}

}  // namespace svo
