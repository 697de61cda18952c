// svo_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code; see svo_oracle.h).
//
// Dependency-free double-precision restatement of the reference's photometric-alignment path.
// Citations are relative to the reference tree (amin-abouee/semi-direct-visual-odometry).
// Third-party arithmetic that is NOT in the reference tree is restated from the library's
// documented behaviour: cv::pyrDown (OpenCV, unpinned), Simd::AbsGradientSaturatedSum (Simd,
// unpinned), Eigen::LDLT (unpinned), Sophus::SE3d (unpinned), std::nth_element (libstdc++ 13).
#include "svo_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <thread>
#include <vector>
#include <atomic>

namespace {

// ------------------------------------------------------------------------------------------------
// pyramid geometry
// ------------------------------------------------------------------------------------------------
inline int levelDim(int v, int level)
{
    for (int i = 0; i < level; i++) v = (v + 1) / 2;  // cv::pyrDown default dsize, src/image_pyramid.cpp:49-50
    return v;
}
inline int64_t levelOffset(int w, int h, int level)
{
    int64_t off = 0;
    for (int l = 0; l < level; l++) off += (int64_t)levelDim(w, l) * levelDim(h, l);
    return off;
}
inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// ------------------------------------------------------------------------------------------------
// SE3 (Sophus convention): params qx qy qz qw tx ty tz; tangent (upsilon, omega)
// ------------------------------------------------------------------------------------------------
struct SE3 {
    double q[4];  // x y z w
    double t[3];
};
inline SE3 fromParams(const double p[7])
{
    SE3 T;
    for (int i = 0; i < 4; i++) T.q[i] = p[i];
    for (int i = 0; i < 3; i++) T.t[i] = p[4 + i];
    return T;
}
inline void toParams(const SE3& T, double p[7])
{
    for (int i = 0; i < 4; i++) p[i] = T.q[i];
    for (int i = 0; i < 3; i++) p[4 + i] = T.t[i];
}
inline void cross(const double a[3], const double b[3], double c[3])
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
// Eigen::Quaternion::_transformVector: v + w*uv + qv x uv with uv = 2 qv x v
inline void rotate(const double q[4], const double v[3], double out[3])
{
    double uv[3], c2[3];
    cross(q, v, uv);
    uv[0] *= 2;
    uv[1] *= 2;
    uv[2] *= 2;
    cross(q, uv, c2);
    for (int i = 0; i < 3; i++) out[i] = v[i] + q[3] * uv[i] + c2[i];
}
inline void quatMul(const double a[4], const double b[4], double o[4])
{
    const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
    const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
    double r[4];
    r[0] = aw * bx + ax * bw + ay * bz - az * by;
    r[1] = aw * by + ay * bw + az * bx - ax * bz;
    r[2] = aw * bz + az * bw + ax * by - ay * bx;
    r[3] = aw * bw - ax * bx - ay * by - az * bz;
    const double n = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
    for (int i = 0; i < 4; i++) o[i] = r[i] / n;
}
inline SE3 mul(const SE3& a, const SE3& b)
{
    SE3 o;
    quatMul(a.q, b.q, o.q);
    double rt[3];
    rotate(a.q, b.t, rt);
    for (int i = 0; i < 3; i++) o.t[i] = a.t[i] + rt[i];
    return o;
}
inline void act(const SE3& T, const double p[3], double out[3])
{
    rotate(T.q, p, out);
    for (int i = 0; i < 3; i++) out[i] += T.t[i];
}
inline SE3 inverse(const SE3& T)
{
    SE3 o;
    o.q[0] = -T.q[0];
    o.q[1] = -T.q[1];
    o.q[2] = -T.q[2];
    o.q[3] = T.q[3];
    double rt[3];
    rotate(o.q, T.t, rt);
    for (int i = 0; i < 3; i++) o.t[i] = -rt[i];
    return o;
}
// Sophus::SE3d::exp: rotation by SO3::exp (unit quaternion, Taylor below eps), t = V * upsilon
inline SE3 expSE3(const double xi[6])
{
    const double* ups = xi;
    const double* om  = xi + 3;
    const double th2  = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
    const double th   = std::sqrt(th2);
    double imag, real;
    if (th2 < 1e-10 * 1e-10) {
        const double th4 = th2 * th2;
        imag             = 0.5 - th2 / 48.0 + th4 / 3840.0;
        real             = 1.0 - th2 / 8.0 + th4 / 384.0;
    } else {
        imag = std::sin(0.5 * th) / th;
        real = std::cos(0.5 * th);
    }
    SE3 T;
    T.q[0] = imag * om[0];
    T.q[1] = imag * om[1];
    T.q[2] = imag * om[2];
    T.q[3] = real;
    // V = I + (1-cos)/th^2 [om]x + (th - sin)/th^3 [om]x^2 ;  th -> 0: V = I + 1/2 [om]x
    double a, b;
    if (th < 1e-10) {
        a = 0.5;
        b = 1.0 / 6.0;
    } else {
        a = (1.0 - std::cos(th)) / th2;
        b = (th - std::sin(th)) / (th2 * th);
    }
    double c1[3], c2[3];
    cross(om, ups, c1);
    cross(om, c1, c2);
    for (int i = 0; i < 3; i++) T.t[i] = ups[i] + a * c1[i] + b * c2[i];
    return T;
}
inline void rotationMatrix(const double q[4], double R[9])
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
// Frame::cameraInWorld, src/frame.cpp:116-120: -R^T t
inline void cameraInWorld(const SE3& T, double C[3])
{
    double R[9];
    rotationMatrix(T.q, R);
    for (int i = 0; i < 3; i++) C[i] = -(R[0 + i] * T.t[0] + R[3 + i] * T.t[1] + R[6 + i] * T.t[2]);
}

// ------------------------------------------------------------------------------------------------
// bilinear interpolation, src/algorithm.cpp:885-905 (truncation, no bounds check)
// ------------------------------------------------------------------------------------------------
inline double bilinearDouble(const uint8_t* img, int pitch, double x, double y)
{
    const int32_t x1 = static_cast<int32_t>(x);
    const int32_t y1 = static_cast<int32_t>(y);
    const int32_t x2 = x1 + 1;
    const int32_t y2 = y1 + 1;
    const double a   = (x2 - x) * img[y1 * pitch + x1] + (x - x1) * img[y1 * pitch + x2];
    const double b   = (x2 - x) * img[y2 * pitch + x1] + (x - x1) * img[y2 * pitch + x2];
    return (y2 - y) * a + (y - y1) * b;
}
inline float bilinearFloat(const uint8_t* img, int pitch, double x, double y)
{
    const int x1  = static_cast<int>(x);
    const int y1  = static_cast<int>(y);
    const int x2  = x1 + 1;
    const int y2  = y1 + 1;
    const float a = (x2 - x) * img[y1 * pitch + x1] + (x - x1) * img[y1 * pitch + x2];  // rounded to float
    const float b = (x2 - x) * img[y2 * pitch + x1] + (x - x1) * img[y2 * pitch + x2];
    return ((y2 - y) * a + (y - y1) * b);
}

// ------------------------------------------------------------------------------------------------
// median / MAD / sigma, src/algorithm.cpp:834-872
// ------------------------------------------------------------------------------------------------
double medianImpl(std::vector<double>& vec, uint32_t numValid, int mode)
{
    const size_t mid = numValid / 2;
    if (vec.empty()) return std::numeric_limits<double>::quiet_NaN();
    std::nth_element(vec.begin(), vec.begin() + mid, vec.end());
    if (vec.size() % 2 != 0) return vec[mid];
    if (mid == 0) return vec[0];  // the reference reads vec[-1] here (UB); defined as vec[0]
    double lower;
    if (mode == ORC_MEDIAN_LIBSTDCXX)
        lower = vec[mid - 1];  // literal: whatever nth_element left there
    else
        lower = *std::max_element(vec.begin(), vec.begin() + mid);  // the true (mid-1)-th order statistic
    return (lower + vec[mid]) / 2.0;
}
double medianOf(const double* v, size_t n, uint32_t numValid, int mode)
{
    std::vector<double> vec(v, v + n);
    return medianImpl(vec, numValid, mode);
}
double sigmaOf(const double* v, size_t n, uint32_t numValid, int mode)
{
    const double med = medianOf(v, n, numValid, mode);
    std::vector<double> diff(n);
    for (size_t i = 0; i < n; i++) diff[i] = std::abs(v[i] - med);
    const double mad = medianImpl(diff, numValid, mode);
    return 1.482602218505602 * mad;  // include/algorithm.hpp:139
}

// ------------------------------------------------------------------------------------------------
// Eigen::LDLT::solve restated: symmetric pivoting on the largest |diagonal|, D^-1 with zero for
// pivots below the smallest normal double (call sites src/optimizer.cpp:96,306)
// ------------------------------------------------------------------------------------------------
bool ldltSolve(const double* Ain, const double* b, int n, double* x)
{
    double A[36];
    int perm[6];
    for (int i = 0; i < n * n; i++) A[i] = Ain[i];
    for (int i = 0; i < n; i++) perm[i] = i;
    bool ok = true;
    for (int k = 0; k < n; k++) {
        int piv    = k;
        double big = std::abs(A[k * n + k]);
        for (int i = k + 1; i < n; i++)
            if (std::abs(A[i * n + i]) > big) {
                big = std::abs(A[i * n + i]);
                piv = i;
            }
        if (piv != k) {  // symmetric row/column swap
            for (int j = 0; j < n; j++) std::swap(A[k * n + j], A[piv * n + j]);
            for (int j = 0; j < n; j++) std::swap(A[j * n + k], A[j * n + piv]);
            std::swap(perm[k], perm[piv]);
        }
        // A(k,k) -= sum_j L(k,j)^2 D(j);  A(i,k) = (A(i,k) - sum_j L(i,j) L(k,j) D(j)) / D(k)
        double dk = A[k * n + k];
        for (int j = 0; j < k; j++) dk -= A[k * n + j] * A[k * n + j] * A[j * n + j];
        A[k * n + k] = dk;
        for (int i = k + 1; i < n; i++) {
            double s = A[i * n + k];
            for (int j = 0; j < k; j++) s -= A[i * n + j] * A[k * n + j] * A[j * n + j];
            if (std::abs(dk) > 0)
                A[i * n + k] = s / dk;
            else {
                A[i * n + k] = s;
                if (s != 0.0) ok = false;
            }
        }
    }
    double y[6];
    for (int i = 0; i < n; i++) y[i] = b[perm[i]];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++) y[i] -= A[i * n + j] * y[j];
    const double tol = std::numeric_limits<double>::min();
    for (int i = 0; i < n; i++) y[i] = std::abs(A[i * n + i]) > tol ? y[i] / A[i * n + i] : 0.0;
    for (int i = n - 1; i >= 0; i--)
        for (int j = i + 1; j < n; j++) y[i] -= A[j * n + i] * y[j];
    for (int i = 0; i < n; i++) x[perm[i]] = y[i];
    return ok;
}

// ------------------------------------------------------------------------------------------------
// Optimizer, src/optimizer.cpp (LM :161-370, GN :41-159, Tukey :485-507, chi2 :470-483)
// ------------------------------------------------------------------------------------------------
struct FirstIter {
    bool filled = false;
    double H[36], g[6], dx[6], chi2, sigma, lambda;
    int n_px;
};

struct Optimizer {
    int P;
    size_t N = 0;
    std::vector<double> J, r, w;
    std::vector<uint8_t> vis;
    double H[36], g[6], dx[6];
    uint32_t maxIteration = 20;      // src/optimizer.cpp:18
    double minChiSquaredError = 1e-1;  // :20
    double stepSize           = 1e-16; // :21
    double maxCoffDx          = 1e+3;  // :23
    int medianMode            = ORC_MEDIAN_EXACT;
    double lastSigma          = 0.0;
    int evaluations           = 0;
    int iterations            = 0;

    explicit Optimizer(int p) : P(p) {}

    void initParameters(size_t n)  // :378-385
    {
        N = n;
        J.assign(n * P, 0.0);
        r.resize(n);
        w.resize(n);
        vis.resize(n);
    }
    void resetResidualParameters()  // :398-403
    {
        std::fill(r.begin(), r.end(), std::numeric_limits<double>::max());
        std::fill(w.begin(), w.end(), 0.0);
        std::fill(vis.begin(), vis.end(), 0);
    }
    void tukeyWeighting(uint32_t numValid)  // :485-507
    {
        double sigma = sigmaOf(r.data(), N, numValid, medianMode);
        if (sigma <= std::numeric_limits<double>::epsilon()) sigma = std::numeric_limits<double>::epsilon();
        lastSigma       = sigma;
        const double c  = 4.6851 * sigma;
        const double c2 = c * c;
        for (size_t i = 0; i < N; i++) {
            if (vis[i]) {
                const double a = std::abs(r[i]);
                if (a <= c) {
                    const double t = 1.0 - (r[i] * r[i]) / c2;
                    w[i]           = t * t;
                } else
                    w[i] = 0;
            }
        }
    }
    double chiSquared() const  // :470-483
    {
        double s = 0.0;
        for (size_t i = 0; i < N; i++)
            if (vis[i]) s += r[i] * r[i] * w[i];
        return s;
    }
    void normalEquations()  // :279-280 (H = J^T W J, g = J^T W r over all rows)
    {
        for (int i = 0; i < P * P; i++) H[i] = 0;
        for (int i = 0; i < P; i++) g[i] = 0;
        for (size_t i = 0; i < N; i++) {
            const double wi = w[i];
            if (wi == 0.0) continue;  // zero-weight rows contribute exactly 0
            const double* Ji = &J[i * P];
            const double wr  = wi * r[i];
            for (int a = 0; a < P; a++) {
                const double wa = wi * Ji[a];
                g[a] += Ji[a] * wr;
                for (int b = a; b < P; b++) H[a * P + b] += wa * Ji[b];
            }
        }
        for (int a = 0; a < P; a++)
            for (int b = 0; b < a; b++) H[a * P + b] = H[b * P + a];
    }
    static bool updateParameters(double pre, double cur, double& lambda, double& nu)  // Nielsen, :449-466
    {
        const double rho = pre - cur;
        if (rho > 0.0) {
            lambda *= std::max<double>(1.0 / 3.0, 1.0 - std::pow((2 * rho - 1), 3));
            nu = 2.0;
            return true;
        }
        lambda *= nu;
        nu *= 2;
        return false;
    }
    double maxCoeff() const
    {
        double m = dx[0];
        for (int i = 1; i < P; i++) m = std::max(m, dx[i]);
        return m;
    }
    bool anyNan() const
    {
        for (int i = 0; i < P; i++)
            if (std::isnan(dx[i])) return true;
        return false;
    }
    void record(FirstIter* fi, double chi2, double lambda, int n)
    {
        if (!fi || fi->filled) return;
        fi->filled = true;
        for (int i = 0; i < 36; i++) fi->H[i] = 0;
        for (int a = 0; a < P; a++)
            for (int b = 0; b < P; b++) fi->H[a * 6 + b] = H[a * P + b];
        for (int a = 0; a < 6; a++) fi->g[a] = a < P ? g[a] : 0;
        fi->chi2   = chi2;
        fi->sigma  = lastSigma;
        fi->lambda = lambda;
        fi->n_px   = n;
    }

    template <typename T>
    std::pair<int, double> optimizeLM(T& params, const std::function<uint32_t(T&)>& residual,
                                      const std::function<void(T&, const double*)>& update, bool faithful, FirstIter* fi)
    {
        int status = ORC_ST_FAILED;
        if (N < (size_t)P) return {ORC_ST_NON_SUFF_POINTS, -1.0};
        uint32_t cur = 0;
        double step = 0, chi2 = 0, preChi2 = 0, lambda = 1e-2, nu = 2.0;
        uint32_t cnt = 0, preCnt = 0;
        resetResidualParameters();
        cnt = residual(params);
        evaluations++;
        tukeyWeighting(cnt);
        chi2 = chiSquared();
        T preParams = params;
        std::vector<double> preR, preW;
        std::vector<uint8_t> preVis;
        bool success = true;
        while (cur < maxIteration) {
            if (success) {
                preParams = params;
                preChi2   = chi2;
                preR      = r;
                preW      = w;
                preVis    = vis;
                preCnt    = cnt;
                status    = ORC_ST_SUCCESS;
            }
            normalEquations();
            if (cur == 0) {
                double mx = H[0];
                for (int i = 1; i < P; i++) mx = std::max(mx, H[i * P + i]);
                lambda *= mx;  // :296-299
            }
            const bool first = fi && !fi->filled;
            record(fi, chi2, lambda, (int)cnt);  // undamped H
            for (int i = 0; i < P; i++) H[i * P + i] += lambda;
            ldltSolve(H, g, P, dx);
            if (first)
                for (int a = 0; a < 6; a++) fi->dx[a] = a < P ? dx[a] : 0;
            update(params, dx);  // :310 -- applied before any check
            iterations++;
            if (maxCoeff() > maxCoffDx) {
                status = ORC_ST_MAX_COFF_DX;
                break;
            }
            if (anyNan()) {
                status = ORC_ST_NAN_IN_DX;
                break;
            }
            step = 0;
            for (int i = 0; i < P; i++) step += dx[i] * dx[i];
            // :328 -- `normDiffPose < m_normInfDiff` is always true in the reference
            if (step < stepSize || lambda >= 1e14 || lambda <= 1e-14 || faithful) {
                status = step < stepSize ? ORC_ST_SMALL_STEP : status;
                status = std::abs(lambda) >= 1e14 ? ORC_ST_LAMBDA : status;
                break;
            }
            resetResidualParameters();
            cnt = residual(params);
            evaluations++;
            tukeyWeighting(cnt);
            chi2    = chiSquared();
            success = updateParameters(preChi2, chi2, lambda, nu);
            if (!success) {
                chi2   = preChi2;
                params = preParams;
                r      = preR;
                w      = preW;
                vis    = preVis;
                cnt    = preCnt;
            }
            ++cur;
        }
        return {status, std::sqrt(chi2 / cnt)};
    }

    template <typename T>
    std::pair<int, double> optimizeGN(T& params, const std::function<uint32_t(T&)>& residual,
                                      const std::function<void(T&, const double*)>& update, FirstIter* fi)
    {
        int status = ORC_ST_FAILED;
        if (N < (size_t)P) return {ORC_ST_NON_SUFF_POINTS, -1.0};
        uint32_t cur = 0, cnt = 0;
        double chi2 = 0, step = 0;
        double preChi2 = std::numeric_limits<double>::max();
        T preParams    = params;
        while (cur < maxIteration) {
            resetResidualParameters();
            cnt = residual(params);
            evaluations++;
            tukeyWeighting(cnt);
            chi2 = chiSquared();
            normalEquations();
            ldltSolve(H, g, P, dx);
            if (fi && !fi->filled) {
                record(fi, chi2, 0.0, (int)cnt);
                for (int a = 0; a < 6; a++) fi->dx[a] = a < P ? dx[a] : 0;
            }
            iterations++;
            if (maxCoeff() > maxCoffDx) {
                status = ORC_ST_MAX_COFF_DX;
                break;
            }
            if (anyNan()) {
                status = ORC_ST_NAN_IN_DX;
                break;
            }
            if (chi2 > preChi2) {
                status = ORC_ST_INCREASE_CHI2;
                params = preParams;  // rollback, :113-118
                break;
            }
            preParams = params;
            preChi2   = chi2;
            step      = 0;
            for (int i = 0; i < P; i++) step += dx[i] * dx[i];
            if (step < stepSize || chi2 < minChiSquaredError) {
                update(params, dx);
                status = step < stepSize ? ORC_ST_SMALL_STEP : status;
                status = chi2 < minChiSquaredError ? ORC_ST_SMALL_CHI2 : status;
                break;
            } else {
                update(params, dx);
                status = ORC_ST_SUCCESS;
            }
            ++cur;
        }
        return {status, std::sqrt(chi2 / cnt)};
    }
};

// patch offsets: odd P follows the reference (-half..half); even P is the "even-patch extension"
// (-P/2 .. P/2-1), SURVEY 9.2
inline void patchRange(int P, int& begin, int& end)
{
    const int half = P / 2;
    begin          = -half;
    end            = (P % 2) ? half : half - 1;
}

// ImageAlignment::computeImageJac, src/image_alignment.cpp:194-248
inline void imageJac(const double p[3], double fx, double fy, double Jm[12])
{
    const double x = p[0], y = p[1], z = p[2];
    const double x2 = x * x, y2 = y * y, z2 = z * z;
    Jm[0]  = fx / z;
    Jm[1]  = 0.0;
    Jm[2]  = -(fx * x) / z2;
    Jm[3]  = -(fx * x * y) / z2;
    Jm[4]  = (fx * x2) / z2 + fx;
    Jm[5]  = -(fx * y) / z;
    Jm[6]  = 0.0;
    Jm[7]  = fy / z;
    Jm[8]  = -(fy * y) / z2;
    Jm[9]  = -(fy * y2) / z2 - fy;
    Jm[10] = (fy * x * y) / z2;
    Jm[11] = (fy * x) / z;
}

struct AlignCtx {
    const uint8_t *refPyr, *kfPyr, *curPyr;
    int w, h;
    const orc_feature* feats;
    int nRef, nKf;
    SE3 Tref, Tkf;
    double K[4];
    int P, half, area, pb, pe;
    std::vector<double> refPatches;  // F x area
    std::vector<uint8_t> refVis;
    Optimizer opt{6};
};

// ImageAlignment::computeJacobian, src/image_alignment.cpp:69-126 (+ SingleFeature :128-192)
void computeJacobian(AlignCtx& c, int level)
{
    std::fill(c.refVis.begin(), c.refVis.end(), 0);
    std::fill(c.refPatches.begin(), c.refPatches.end(), 0.0);
    const int border    = c.half + 2;
    const int lw        = levelDim(c.w, level);
    const int lh        = levelDim(c.h, level);
    const double denom  = 1 << level;
    const double scale  = 1.0 / denom;
    const double fx     = c.K[0] / denom;
    const double fy     = c.K[1] / denom;
    const int64_t off   = levelOffset(c.w, c.h, level);
    for (int which = 0; which < 2; which++) {
        const SE3& Tf      = which == 0 ? c.Tref : c.Tkf;
        const uint8_t* img = (which == 0 ? c.refPyr : c.kfPyr);
        const int begin    = which == 0 ? 0 : c.nRef;
        const int end      = which == 0 ? c.nRef : c.nRef + c.nKf;
        if (begin == end) continue;
        img += off;
        double C[3];
        cameraInWorld(Tf, C);
        const SE3 Tinv = inverse(Tf);
        for (int f = begin; f < end; f++) {
            const orc_feature& ft = c.feats[f];
            if (!ft.has_point) continue;
            const double u = ft.px[0] * scale;
            const double v = ft.px[1] * scale;
            const int uI   = (int)std::floor(u);
            const int vI   = (int)std::floor(v);
            if ((uI - border) < 0 || (vI - border) < 0 || (uI + border) >= lw || (vI + border) >= lh) continue;
            c.refVis[f] = 1;
            const double d0 = ft.point[0] - C[0], d1 = ft.point[1] - C[1], d2 = ft.point[2] - C[2];
            const double depthNorm = std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            const double pC[3]     = {ft.bearing[0] * depthNorm, ft.bearing[1] * depthNorm, ft.bearing[2] * depthNorm};
            double pW[3];
            act(Tinv, pC, pW);  // Frame::camera2world, src/frame.cpp:94-97
            double Jm[12];
            imageJac(pW, fx, fy, Jm);
            int cnt = 0;
            for (int y = c.pb; y <= c.pe; y++) {
                for (int x = c.pb; x <= c.pe; x++, cnt++) {
                    const double row = v + y, col = u + x;
                    c.refPatches[(size_t)f * c.area + cnt] = bilinearDouble(img, lw, col, row);
                    const double dx = 0.5 * (bilinearDouble(img, lw, col + 1, row) - bilinearDouble(img, lw, col - 1, row));
                    const double dy = 0.5 * (bilinearDouble(img, lw, col, row + 1) - bilinearDouble(img, lw, col, row - 1));
                    double* Jr = &c.opt.J[((size_t)f * c.area + cnt) * 6];
                    for (int k = 0; k < 6; k++) Jr[k] = dx * Jm[k] + dy * Jm[6 + k];
                }
            }
        }
    }
}

// ImageAlignment::computeResiduals, src/image_alignment.cpp:251-308 (+ SingleFeature :310-370)
uint32_t computeResiduals(AlignCtx& c, int level, const SE3& pose)
{
    const int border   = c.half + 2;
    const int lw       = levelDim(c.w, level);
    const int lh       = levelDim(c.h, level);
    const double scale = 1.0 / (double)(1 << level);
    const uint8_t* img = c.curPyr + levelOffset(c.w, c.h, level);
    uint32_t total     = 0;
    for (int which = 0; which < 2; which++) {
        const SE3& Tf   = which == 0 ? c.Tref : c.Tkf;
        const int begin = which == 0 ? 0 : c.nRef;
        const int end   = which == 0 ? c.nRef : c.nRef + c.nKf;
        if (begin == end) continue;
        double C[3];
        cameraInWorld(Tf, C);
        const SE3 Tinv = inverse(Tf);
        for (int f = begin; f < end; f++) {
            if (!c.refVis[f]) continue;
            const orc_feature& ft = c.feats[f];
            const double d0 = ft.point[0] - C[0], d1 = ft.point[1] - C[1], d2 = ft.point[2] - C[2];
            const double depthNorm = std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            const double pC[3]     = {ft.bearing[0] * depthNorm, ft.bearing[1] * depthNorm, ft.bearing[2] * depthNorm};
            double pW[3], pCur[3];
            act(Tinv, pC, pW);
            act(pose, pW, pCur);
            // PinholeCamera::project2d, src/pinhole_camera.cpp:55-56 (no z > 0 test)
            const double uu = c.K[0] * (pCur[0] / pCur[2]) + c.K[2];
            const double vv = c.K[1] * (pCur[1] / pCur[2]) + c.K[3];
            const double u = uu * scale, v = vv * scale;
            if (!(std::isfinite(u) && std::isfinite(v))) continue;  // the reference would index out of bounds
            const int uI = (int)std::floor(u);
            const int vI = (int)std::floor(v);
            if ((uI - border) < 0 || (vI - border) < 0 || (uI + border) >= lw || (vI + border) >= lh) continue;
            int cnt = 0;
            for (int y = c.pb; y <= c.pe; y++) {
                for (int x = c.pb; x <= c.pe; x++, cnt++, total++) {
                    const double val   = bilinearDouble(img, lw, u + x, v + y);
                    const size_t idx   = (size_t)f * c.area + cnt;
                    c.opt.r[idx]       = val - c.refPatches[idx];
                    c.opt.vis[idx]     = 1;
                }
            }
        }
    }
    return total;
}

double sparseAlign(const uint8_t* refPyr, const uint8_t* kfPyr, const uint8_t* curPyr, int w, int h,
                   const orc_feature* feats, int nRef, int nKf, const double Tref[7], const double Tkf[7],
                   const double K[4], const orc_align_params* prm, double Tcur[7], orc_level_stats* stats,
                   int32_t* statusOut, int32_t* evalsOut)
{
    if (statusOut) *statusOut = ORC_ST_SUCCESS;
    if (evalsOut) *evalsOut = 0;
    if (nRef == 0) return 0;  // src/image_alignment.cpp:27-28
    AlignCtx c;
    c.refPyr = refPyr;
    c.kfPyr  = kfPyr;
    c.curPyr = curPyr;
    c.w      = w;
    c.h      = h;
    c.feats  = feats;
    c.nRef   = nRef;
    c.nKf    = nKf;
    c.Tref   = fromParams(Tref);
    c.Tkf    = fromParams(Tkf);
    for (int i = 0; i < 4; i++) c.K[i] = K[i];
    c.P    = prm->patch_size;
    c.half = c.P / 2;
    c.area = c.P * c.P;
    patchRange(c.P, c.pb, c.pe);
    const int F = nRef + nKf;
    c.refPatches.assign((size_t)F * c.area, 0.0);
    c.refVis.assign(F, 0);
    c.opt.initParameters((size_t)F * c.area);
    c.opt.maxIteration = prm->max_iter > 0 ? prm->max_iter : 20;
    c.opt.medianMode   = prm->median_mode;
    SE3 pose           = fromParams(Tcur);
    double error       = 0.0;
    int status         = ORC_ST_FAILED;
    int si             = 0;
    for (int level = prm->max_level; level >= prm->min_level; level--, si++) {
        computeJacobian(c, level);
        std::function<uint32_t(SE3&)> res = [&c, level](SE3& p) -> uint32_t { return computeResiduals(c, level, p); };
        // ImageAlignment::update, src/image_alignment.cpp:372-380: pose = pose * exp(-dx)
        std::function<void(SE3&, const double*)> upd = [](SE3& p, const double* dx) {
            double m[6];
            for (int i = 0; i < 6; i++) m[i] = -dx[i];
            p = mul(p, expSE3(m));
        };
        FirstIter fi;
        const int it0 = c.opt.iterations, ev0 = c.opt.evaluations;
        std::pair<int, double> out;
        if (prm->mode == ORC_GN)
            out = c.opt.optimizeGN<SE3>(pose, res, upd, &fi);
        else
            out = c.opt.optimizeLM<SE3>(pose, res, upd, prm->mode == ORC_LM_FAITHFUL, &fi);
        status = out.first;
        error  = out.second;
        if (stats) {
            orc_level_stats& s = stats[si];
            std::memset(&s, 0, sizeof(s));
            if (fi.filled) {
                std::memcpy(s.H, fi.H, sizeof(s.H));
                std::memcpy(s.g, fi.g, sizeof(s.g));
                std::memcpy(s.dx, fi.dx, sizeof(s.dx));
                s.chi2   = fi.chi2;
                s.sigma  = fi.sigma;
                s.lambda = fi.lambda;
                s.n_px   = fi.n_px;
            }
            toParams(pose, s.pose_after);
            s.rmse        = error;
            s.status      = status;
            s.iterations  = c.opt.iterations - it0;
            s.evaluations = c.opt.evaluations - ev0;
        }
    }
    toParams(pose, Tcur);
    if (statusOut) *statusOut = status;
    if (evalsOut) *evalsOut = c.opt.evaluations;
    return error;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

void orc_abs_gradient(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch)
{
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            if (x == 0 || y == 0 || x == w - 1 || y == h - 1) {
                dst[y * dpitch + x] = 0;
                continue;
            }
            const int dx = std::abs((int)src[y * spitch + x + 1] - (int)src[y * spitch + x - 1]);
            const int dy = std::abs((int)src[(y + 1) * spitch + x] - (int)src[(y - 1) * spitch + x]);
            dst[y * dpitch + x] = (uint8_t)std::min(dx + dy, 255);
        }
    }
}

void orc_pyrdown(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch)
{
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    static const int k[5] = {1, 4, 6, 4, 1};
    std::vector<int> rowbuf((size_t)5 * dw);
    for (int dy = 0; dy < dh; dy++) {
        // horizontal pass for the five source rows of this output row (integer, no rounding)
        for (int j = 0; j < 5; j++) {
            const int sy       = reflect101(2 * dy - 2 + j, h);
            const uint8_t* row = src + (size_t)sy * spitch;
            for (int dx = 0; dx < dw; dx++) {
                int s = 0;
                for (int i = 0; i < 5; i++) s += k[i] * row[reflect101(2 * dx - 2 + i, w)];
                rowbuf[(size_t)j * dw + dx] = s;
            }
        }
        for (int dx = 0; dx < dw; dx++) {
            int s = 0;
            for (int j = 0; j < 5; j++) s += k[j] * rowbuf[(size_t)j * dw + dx];
            dst[(size_t)dy * dpitch + dx] = (uint8_t)((s + 128) >> 8);
        }
    }
}

int64_t orc_pyramid_bytes(int w, int h, int levels) { return levelOffset(w, h, levels); }

void orc_build_pyramid(const uint8_t* img, int w, int h, int pitch, int levels, uint8_t* imgPyr, uint8_t* gradPyr)
{
    // src/image_pyramid.cpp:36-52: gradient of the base image, then both stacks by repeated pyrDown
    for (int y = 0; y < h; y++) std::memcpy(imgPyr + (size_t)y * w, img + (size_t)y * pitch, w);
    orc_abs_gradient(imgPyr, w, h, w, gradPyr, w);
    for (int l = 1; l < levels; l++) {
        const int sw = levelDim(w, l - 1), sh = levelDim(h, l - 1), dw = levelDim(w, l);
        orc_pyrdown(imgPyr + levelOffset(w, h, l - 1), sw, sh, sw, imgPyr + levelOffset(w, h, l), dw);
        orc_pyrdown(gradPyr + levelOffset(w, h, l - 1), sw, sh, sw, gradPyr + levelOffset(w, h, l), dw);
    }
}

int orc_grid_select(const uint8_t* grad, int w, int h, int pitch, int cell, uint32_t thr, const uint8_t* occ,
                    int32_t* out, int maxOut)
{
    const int rows = h / cell + 1, cols = w / cell + 1;  // src/feature_selection.cpp:19-25
    int n = 0;
    for (int r = 0; r < rows; r++) {
        for (int c = 0; c < cols; c++) {
            if (occ && occ[r * cols + c]) continue;
            const int maxCol = (c + 1) * cell < w ? cell : w - (c * cell);  // :114
            const int maxRow = (r + 1) * cell < h ? cell : h - (r * cell);  // :115
            uint32_t mx = 0;
            int rowIdx = 0, colIdx = 0;
            for (int i = 0; i < maxRow; i++)
                for (int j = 0; j < maxCol; j++) {
                    const uint8_t v = grad[(size_t)(r * cell + i) * pitch + c * cell + j];
                    if (v > mx) {  // strict: first maximal pixel in raster order, :122-133
                        rowIdx = r * cell + i;
                        colIdx = c * cell + j;
                        mx     = v;
                    }
                }
            if (mx > thr) {
                if (n < maxOut) {
                    out[3 * n + 0] = colIdx;
                    out[3 * n + 1] = rowIdx;
                    out[3 * n + 2] = (int)mx;
                }
                n++;
            }
        }
    }
    return n;
}

// FeatureSelection::gradientMagnitudeWithSSC + SSC, src/feature_selection.cpp:27-89, :165-248
int orc_select_ssc(const uint8_t* grad, int w, int h, int pitch, uint32_t thr, int numRetPoints, int cell, const uint8_t* occ,
                   int useBucketing, int32_t* out, int maxOut, int32_t* info)
{
    struct KP {
        float x, y, response;  // cv::KeyPoint: pt is Point2f, response float
    };
    std::vector<KP> kps;
    for (int i = 0; i < h; i++)  // :41-51, raster order
        for (int j = 0; j < w; j++) {
            const uint8_t v = grad[(size_t)i * pitch + j];
            if (v > thr) kps.push_back({(float)j, (float)i, (float)v});
        }
    // :54-55 std::sort by response, descending; equal responses: stable (raster order) -- see the header
    std::stable_sort(kps.begin(), kps.end(), [](const KP& a, const KP& b) { return a.response > b.response; });

    // SSC, :165-248
    const int cols = w, rows = h;
    std::vector<int32_t> resultVec, result;
    int iterations = 0, widthUsed = -1;
    {
        const int32_t exp1   = rows + cols + 2 * numRetPoints;
        const long long exp2 = ((long long)4 * cols + (long long)4 * numRetPoints + (long long)4 * rows * numRetPoints +
                                (long long)rows * rows + (long long)cols * cols - (long long)2 * rows * cols +
                                (long long)4 * rows * cols * numRetPoints);
        const double exp3 = std::sqrt((double)exp2);
        const double exp4 = (2 * (numRetPoints - 1));
        const double sol1 = -std::round((exp1 + exp3) / exp4);
        const double sol2 = -std::round((exp1 - exp3) / exp4);
        int high = (sol1 > sol2) ? (int)sol1 : (int)sol2;
        int low  = (int)std::sqrt((double)kps.size() / numRetPoints);
        int width, prevWidth = -1;
        bool complete        = false;
        const float K        = (float)numRetPoints;
        const float tolerance = 0.1f;
        const uint32_t Kmin  = (uint32_t)std::round(K - (K * tolerance));
        const uint32_t Kmax  = (uint32_t)std::round(K + (K * tolerance));
        while (!complete) {
            width = low + (high - low) / 2;
            // width <= 0 divides by zero in the reference (c = 0): defined here as "stop with the previous result"
            if (width == prevWidth || low > high || width <= 0) {
                resultVec = result;
                break;
            }
            iterations++;
            widthUsed = width;
            result.clear();
            const double c            = width / 2.0;
            const int32_t numCellCols = (int32_t)(cols / c);
            const int32_t numCellRows = (int32_t)(rows / c);
            std::vector<uint8_t> covered((size_t)(numCellRows + 1) * (numCellCols + 1), 0);
            const int32_t reach = (int32_t)(width / c);
            for (size_t i = 0; i < kps.size(); ++i) {
                const int32_t row = (int32_t)(kps[i].y / c);
                const int32_t col = (int32_t)(kps[i].x / c);
                if (!covered[(size_t)row * (numCellCols + 1) + col]) {
                    result.push_back((int32_t)i);
                    const int32_t rowMin = row >= reach ? row - reach : 0;
                    const int32_t rowMax = (row + reach <= numCellRows) ? row + reach : numCellRows;
                    const int32_t colMin = col >= reach ? col - reach : 0;
                    const int32_t colMax = (col + reach <= numCellCols) ? col + reach : numCellCols;
                    for (int32_t r = rowMin; r <= rowMax; ++r)
                        for (int32_t cc = colMin; cc <= colMax; ++cc) covered[(size_t)r * (numCellCols + 1) + cc] = 1;
                }
            }
            if (result.size() >= Kmin && result.size() <= Kmax) {
                resultVec = result;
                complete  = true;
            } else if (result.size() < Kmin)
                high = width - 1;
            else
                low = width + 1;
            prevWidth = width;
        }
    }
    if (info) {
        info[0] = (int32_t)kps.size();
        info[1] = widthUsed;
        info[2] = iterations;
        info[3] = (int32_t)resultVec.size();
    }
    int n = 0;
    const int gridCols = w / cell + 1, gridRows = h / cell + 1;
    std::vector<uint8_t> grid((size_t)gridRows * gridCols, 0);
    if (occ) std::copy(occ, occ + grid.size(), grid.begin());
    for (size_t i = 0; i < resultVec.size(); i++) {  // :62-89
        const KP& kp = kps[resultVec[i]];
        if (useBucketing) {
            const int32_t idx = (int32_t)kp.x / cell, idy = (int32_t)kp.y / cell;
            if (grid[(size_t)idy * gridCols + idx]) continue;
            grid[(size_t)idy * gridCols + idx] = 1;
        }
        if (n < maxOut) {
            out[3 * n + 0] = (int32_t)kp.x;
            out[3 * n + 1] = (int32_t)kp.y;
            out[3 * n + 2] = (int32_t)kp.response;
        }
        n++;
    }
    return n;
}

double orc_bilinear_double(const uint8_t* img, int pitch, double x, double y) { return bilinearDouble(img, pitch, x, y); }
float orc_bilinear_float(const uint8_t* img, int pitch, double x, double y) { return bilinearFloat(img, pitch, x, y); }
double orc_median(const double* v, int n, int numValid, int mode) { return medianOf(v, n, numValid, mode); }
double orc_sigma(const double* v, int n, int numValid, int mode) { return sigmaOf(v, n, numValid, mode); }
void orc_project2d(const double K[4], const double p[3], double uv[2])
{
    uv[0] = K[0] * (p[0] / p[2]) + K[2];
    uv[1] = K[1] * (p[1] / p[2]) + K[3];
}
void orc_image_jac(const double p[3], double fx, double fy, double J[12]) { imageJac(p, fx, fy, J); }
void orc_se3_exp(const double xi[6], double qt[7]) { toParams(expSE3(xi), qt); }
void orc_se3_mul(const double a[7], const double b[7], double out[7]) { toParams(mul(fromParams(a), fromParams(b)), out); }
void orc_se3_act(const double T[7], const double p[3], double out[3]) { act(fromParams(T), p, out); }
void orc_se3_inv(const double T[7], double out[7]) { toParams(inverse(fromParams(T)), out); }
int orc_ldlt_solve(const double* A, const double* b, int n, double* x) { return ldltSolve(A, b, n, x) ? 0 : 1; }

double orc_sparse_align(const uint8_t* refPyr, const uint8_t* kfPyr, const uint8_t* curPyr, int w, int h,
                        const orc_feature* feats, int nRef, int nKf, const double Tref[7], const double Tkf[7],
                        const double K[4], const orc_align_params* prm, double Tcur[7], orc_level_stats* stats,
                        int32_t* statusOut)
{
    return sparseAlign(refPyr, kfPyr, curPyr, w, h, feats, nRef, nKf, Tref, Tkf, K, prm, Tcur, stats, statusOut, nullptr);
}

void orc_sparse_align_batch(orc_align_job* jobs, int nJobs, int w, int h, const double K[4], const orc_align_params* prm,
                            int nThreads)
{
    if (nThreads < 1) nThreads = 1;
    std::atomic<int> next{0};
    auto worker = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= nJobs) break;
            orc_align_job& j = jobs[i];
            j.rmse = sparseAlign(j.ref_pyr, j.kf_pyr, j.cur_pyr, w, h, j.feats, j.n_ref, j.n_kf, j.T_ref, j.T_kf, K, prm,
                                 j.T_cur, nullptr, &j.status, &j.evaluations);
        }
    };
    if (nThreads == 1) {
        worker();
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < nThreads; t++) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
}

// FeatureAlignment::align, src/feature_alignment.cpp:25-62
double orc_feature_align(const uint8_t* refGrad, const uint8_t* curGrad, int w, int h, const double refPx[2],
                         const double* A, double pxInOut[2], const orc_fa_params* prm, int32_t* statusOut,
                         int32_t* iterationsOut)
{
    const int P    = prm->patch_size;
    const int half = P / 2;
    const int area = P * P;
    int pb, pe;
    patchRange(P, pb, pe);
    Optimizer opt(3);
    opt.initParameters(area);
    opt.maxIteration = prm->max_iter > 0 ? prm->max_iter : 20;
    opt.medianMode   = prm->median_mode;
    std::vector<double> refPatch(area, 0.0);
    const double I2[4] = {1, 0, 0, 1};
    const double* Aw   = A ? A : I2;
    // border: half + 2 for the identity warp (:67); for an affine template the farthest central-difference tap
    double border = half + 2;
    if (A) {
        double m = 0;
        for (int sx = -1; sx <= 1; sx += 2)
            for (int sy = -1; sy <= 1; sy += 2) {
                const double cx = sx * (half + 1), cy = sy * (half + 1);
                m = std::max(m, std::abs(Aw[0] * cx + Aw[1] * cy));
                m = std::max(m, std::abs(Aw[2] * cx + Aw[3] * cy));
            }
        border = std::ceil(m) + 1;
    }
    auto inFrame = [w, h](double x, double y, double b) {  // PinholeCamera::isInFrame, src/pinhole_camera.cpp:163-169
        return x >= b && y >= b && x < w - b && y < h - b;
    };
    // computeJacobian, :64-110
    if (inFrame(refPx[0], refPx[1], border)) {
        int cnt = 0;
        for (int y = pb; y <= pe; y++)
            for (int x = pb; x <= pe; x++, cnt++) {
                auto sample = [&](double ox, double oy) {
                    const double col = refPx[0] + Aw[0] * ox + Aw[1] * oy;
                    const double row = refPx[1] + Aw[2] * ox + Aw[3] * oy;
                    return bilinearFloat(refGrad, w, col, row);
                };
                refPatch[cnt]  = sample(x, y);
                const double dx = 0.5 * (sample(x + 1, y) - sample(x - 1, y));
                const double dy = 0.5 * (sample(x, y + 1) - sample(x, y - 1));
                opt.J[cnt * 3 + 0] = dx;
                opt.J[cnt * 3 + 1] = dy;
                opt.J[cnt * 3 + 2] = 1.0;
            }
    }
    struct V3 {
        double v[3];
    };
    V3 flow{{pxInOut[0], pxInOut[1], 0.0}};
    const double curBorder = half + 2;
    // computeResiduals, :113-168
    std::function<uint32_t(V3&)> res = [&](V3& p) -> uint32_t {
        if (!inFrame(p.v[0], p.v[1], curBorder)) return 0;
        uint32_t total = 0;
        int cnt        = 0;
        for (int y = pb; y <= pe; y++)
            for (int x = pb; x <= pe; x++, cnt++, total++) {
                const double val = bilinearFloat(curGrad, w, p.v[0] + x, p.v[1] + y);
                opt.r[cnt]       = -(val - refPatch[cnt] + p.v[2]);  // :152
                opt.vis[cnt]     = 1;
            }
        return total;
    };
    std::function<void(V3&, const double*)> upd = [](V3& p, const double* dx) {  // :200-205
        p.v[0] += dx[0];
        p.v[1] += dx[1];
        p.v[2] += dx[2];
    };
    std::pair<int, double> out;
    if (prm->mode == ORC_GN)
        out = opt.optimizeGN<V3>(flow, res, upd, nullptr);
    else
        out = opt.optimizeLM<V3>(flow, res, upd, prm->mode == ORC_LM_FAITHFUL, nullptr);
    pxInOut[0] = flow.v[0];
    pxInOut[1] = flow.v[1];
    if (statusOut) *statusOut = out.first;
    if (iterationsOut) *iterationsOut = opt.iterations;
    return out.second;
}

// ------------------------------------------------------------------------------------------------
// epipolar search, src/algorithm.cpp:335-551, :682-703
// ------------------------------------------------------------------------------------------------
namespace {
struct Cam {
    double fx, fy, cx, cy;
    int w, h;
    void project(const double p[3], double uv[2]) const  // PinholeCamera::project2d, src/pinhole_camera.cpp:55-56
    {
        uv[0] = fx * (p[0] / p[2]) + cx;
        uv[1] = fy * (p[1] / p[2]) + cy;
    }
    void bearing(double x, double y, double b[3]) const  // inverseProject2d, :81-101 (normalised)
    {
        b[0] = (x - cx) / fx;
        b[1] = (y - cy) / fy;
        b[2] = 1.0;
        const double n = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
        for (int i = 0; i < 3; i++) b[i] /= n;
    }
    bool inFrame(double x, double y, double bd) const { return x >= bd && y >= bd && x < w - bd && y < h - bd; }  // :163-169
};

// curFrame->camera2image(relativePose * refFrame->image2camera(px, depth))
void projectAtDepth(const Cam& cam, const SE3& Trel, double x, double y, double depth, double uv[2])
{
    double b[3], pc[3], pr[3];
    cam.bearing(x, y, b);
    for (int i = 0; i < 3; i++) pr[i] = b[i] * depth;
    act(Trel, pr, pc);
    cam.project(pc, uv);
}

// algorithm::applyAffineWarp, :369-394: leaves `data` untouched when the location is out of frame
void applyAffineWarp(const uint8_t* img, const Cam& cam, const double loc[2], int half, const double A[4], uint8_t* data)
{
    const double bx = A[0] * half + A[1] * half, by = A[2] * half + A[3] * half;
    const double maxBoundary = std::ceil(std::max(std::fabs(bx), std::fabs(by))) + 2;
    if (!cam.inFrame(loc[0], loc[1], maxBoundary)) return;
    int idx = 0;
    for (int i = -half; i <= half; i++)
        for (int j = -half; j <= half; j++) {
            const double x = loc[0] + (A[0] * j + A[1] * i), y = loc[1] + (A[2] * j + A[3] * i);
            data[idx++]    = (uint8_t)bilinearFloat(img, cam.w, x, y);  // float -> uint8: truncation
        }
}

// algorithm::computeScore, :396-410 (mean-removed sum of absolute differences, despite the name)
double computeScore(const uint8_t* ref, const uint8_t* cur, int n, int meanMode)
{
    double refMean, curMean;
    if (meanMode == ORC_MEAN_EIGEN_U8) {  // Eigen: Scalar(redux(sum)) / Scalar(size()) with Scalar = uint8_t
        uint8_t sr = 0, sc = 0;
        for (int i = 0; i < n; i++) {
            sr = (uint8_t)(sr + ref[i]);
            sc = (uint8_t)(sc + cur[i]);
        }
        refMean = (double)(uint8_t)(sr / (uint8_t)n);
        curMean = (double)(uint8_t)(sc / (uint8_t)n);
    } else {
        double sr = 0, sc = 0;
        for (int i = 0; i < n; i++) {
            sr += ref[i];
            sc += cur[i];
        }
        refMean = sr / n;
        curMean = sc / n;
    }
    double sum = 0.0;
    for (int i = 0; i < n; i++) sum += std::fabs((ref[i] - refMean) - (cur[i] - curMean));
    return sum;
}

// algorithm::depthFromTriangulation, :682-703
bool depthFromTriangulation(const SE3& Trel, const double bref[3], const double bcur[3], double* depth)
{
    double a0[3];
    rotate(Trel.q, bref, a0);  // R * bea_ref
    const double a1[3] = {-bcur[0], -bcur[1], -bcur[2]};
    const double m00 = a0[0] * a0[0] + a0[1] * a0[1] + a0[2] * a0[2];
    const double m01 = a0[0] * a1[0] + a0[1] * a1[1] + a0[2] * a1[2];
    const double m11 = a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2];
    const double det = m00 * m11 - m01 * m01;
    if (det < 0.000001) return false;
    const double r0 = a0[0] * Trel.t[0] + a0[1] * Trel.t[1] + a0[2] * Trel.t[2];  // A^T t
    const double r1 = a1[0] * Trel.t[0] + a1[1] * Trel.t[1] + a1[2] * Trel.t[2];
    // depths = -AtA^-1 A^T t, closed-form 2x2 inverse (Eigen's inverse() for fixed 2x2)
    const double d0 = -((m11 * r0 - m01 * r1) / det);
    *depth          = std::fabs(d0);
    return true;
}
}  // namespace

void orc_epipolar_match(const uint8_t* refImg, const uint8_t* curImg, int w, int h, const double K[4], const double Trel7[7],
                        const double refPx[2], const double refBearing[3], double depth, double minDepth, double maxDepth,
                        const orc_epi_params* prm, orc_epi_result* out)
{
    const Cam cam{K[0], K[1], K[2], K[3], w, h};
    const SE3 Trel = fromParams(Trel7);
    const int P = prm->patch_size, half = P / 2, area = P * P;
    const uint32_t thresholdZSSD = (uint32_t)area * 128;  // :427
    out->depth = 0, out->px[0] = out->px[1] = 0, out->score = DBL_MAX, out->found = 0, out->steps = 0;
    double locMin[2], locMax[2];
    projectAtDepth(cam, Trel, refPx[0], refPx[1], minDepth, locMin);
    projectAtDepth(cam, Trel, refPx[0], refPx[1], maxDepth, locMax);
    auto clampLoc = [&](double* l) {  // :435-451
        l[0] = l[0] >= 0 ? l[0] : 0.0;
        l[0] = l[0] < w ? l[0] : w - 1;
        l[1] = l[1] >= 0 ? l[1] : 0.0;
        l[1] = l[1] < h ? l[1] : h - 1;
    };
    clampLoc(locMin);
    clampLoc(locMax);
    const double epi[2] = {locMax[0] - locMin[0], locMax[1] - locMin[1]};
    // getAffineWarp at the initial depth, :335-367 (halfPatchSize is unsigned there)
    double A[4];
    {
        double c[2], du[2], dv[2];
        projectAtDepth(cam, Trel, refPx[0], refPx[1], depth, c);
        projectAtDepth(cam, Trel, refPx[0] + half, refPx[1], depth, du);
        projectAtDepth(cam, Trel, refPx[0], refPx[1] + half, depth, dv);
        A[0] = (du[0] - c[0]) / half;  // column 0 = duDiff / half
        A[2] = (du[1] - c[1]) / half;
        A[1] = (dv[0] - c[0]) / half;  // column 1 = dvDiff / half
        A[3] = (dv[1] - c[1]) / half;
    }
    const double norm = std::sqrt(epi[0] * epi[0] + epi[1] * epi[1]);
    std::vector<uint8_t> refPatch(area, 0), curPatch(area, 0);
    const double I2[4] = {1, 0, 0, 1};
    applyAffineWarp(refImg, cam, refPx, half, I2, refPatch.data());  // :464-467
    if (norm < 2.0) {  // :469-483
        const double c[2] = {(locMax[0] + locMin[0]) / 2.0, (locMax[1] + locMin[1]) / 2.0};
        double bc[3];
        cam.bearing(c[0], c[1], bc);
        out->px[0] = c[0], out->px[1] = c[1];
        out->found = depthFromTriangulation(Trel, refBearing, bc, &out->depth) ? 1 : 0;
        return;
    }
    const uint32_t pixelStep = (uint32_t)std::ceil(norm);  // :498
    const double step[2]     = {epi[0] / norm, epi[1] / norm};
    double minScore = DBL_MAX, best[2] = {0, 0};
    for (uint32_t i = 0; i < pixelStep; i++) {  // :509-523; an out-of-frame step scores the PREVIOUS patch again
        const double loc[2] = {locMin[0] + i * step[0], locMin[1] + i * step[1]};
        applyAffineWarp(curImg, cam, loc, half, A, curPatch.data());
        const double z = computeScore(refPatch.data(), curPatch.data(), area, prm->mean_mode);
        if (z < minScore) {
            minScore = z;
            best[0] = loc[0], best[1] = loc[1];
        }
    }
    out->steps = (int32_t)pixelStep;
    out->score = minScore;
    out->px[0] = best[0], out->px[1] = best[1];
    if (minScore < thresholdZSSD) {  // :526-548
        double bc[3];
        cam.bearing(best[0], best[1], bc);
        out->found = depthFromTriangulation(Trel, refBearing, bc, &out->depth) ? 1 : 0;
    }
}

// Map::reprojectMap, src/map.cpp:260-489
int orc_reproject_map(const uint8_t* const* grads, const uint8_t* curGrad, int w, int h, const double K[4], const double Tcur7[7],
                      const orc_reproj_candidate* cands, int n, int cell, const int32_t* cellOrder, int nCells, int maxMatches,
                      const orc_fa_params* fa, double* matches, uint8_t* projected)
{
    const SE3 T        = fromParams(Tcur7);
    const int gridCols = (int)std::ceil((double)w / cell);  // Map::initializeGrid, :226-227
    std::vector<std::vector<int>> cells(nCells);
    std::vector<double> px(2 * (size_t)std::max(n, 1));
    for (int i = 0; i < n; i++) {  // reprojectPoint, :492-504, in insertion order
        double pc[3];
        act(T, cands[i].point, pc);
        const double u = K[0] * (pc[0] / pc[2]) + K[2], v = K[1] * (pc[1] / pc[2]) + K[3];
        px[2 * i] = u, px[2 * i + 1] = v;
        const bool in = u >= 3 && v >= 3 && u < w - 3 && v < h - 3;
        if (projected) projected[i] = in ? 1 : 0;
        if (in) {
            const int k = (int)v / cell * gridCols + (int)u / cell;
            if (k >= 0 && k < nCells) cells[k].push_back(i);
        }
    }
    int m = 0;
    for (int i = 0; i < nCells; i++) {  // :475-488
        const int idx = cellOrder[i];
        if (idx < 0 || idx >= nCells) continue;
        std::vector<int>& cl = cells[idx];
        if (cl.empty()) continue;
        // reprojectCell, :506-579: sort by type, descending (stable here); the first non-DELETED candidate is aligned and kept
        std::stable_sort(cl.begin(), cl.end(), [&](int a, int b) { return cands[a].type > cands[b].type; });
        bool matched = false;
        for (int ci : cl) {
            if (cands[ci].type == 1) continue;  // Point::PointType::DELETED
            double p[2] = {px[2 * ci], px[2 * ci + 1]};
            int32_t st = 0, it = 0;
            const double err = orc_feature_align(grads[cands[ci].ref_slot], curGrad, w, h, cands[ci].ref_px, nullptr, p, fa, &st, &it);
            double* o = matches + 6 * (size_t)m;
            o[0] = idx, o[1] = ci, o[2] = p[0], o[3] = p[1], o[4] = err, o[5] = st;
            matched = true;
            break;
        }
        if (matched) m++;
        if (m > maxMatches) break;  // :484-487
    }
    return m;
}

// ---------------------------------------------------------------------------------------------------------------------
// cv::calcOpticalFlowPyrLK as the reference calls it (src/algorithm.cpp:60-62), restated from OpenCV 4.x lkpyramid.cpp.
namespace {
struct KltLevel {
    int w, h, pad;
    std::vector<uint8_t> img;    // (h + 2 pad) x (w + 2 pad), BORDER_REFLECT_101
    std::vector<int16_t> deriv;  // same extent, 2 channels (dI/dx, dI/dy), zero outside the image
    int pitch() const { return w + 2 * pad; }
    const uint8_t* I(int x, int y) const { return &img[(size_t)(y + pad) * pitch() + x + pad]; }
    const int16_t* D(int x, int y) const { return &deriv[((size_t)(y + pad) * pitch() + x + pad) * 2]; }
};
void kltPad(const std::vector<uint8_t>& src, int w, int h, int pad, KltLevel& L, bool withDeriv)
{
    L.w = w, L.h = h, L.pad = pad;
    const int P = L.pitch();
    L.img.assign((size_t)(h + 2 * pad) * P, 0);
    for (int y = -pad; y < h + pad; y++)
        for (int x = -pad; x < w + pad; x++) L.img[(size_t)(y + pad) * P + x + pad] = src[(size_t)reflect101(y, h) * w + reflect101(x, w)];
    if (!withDeriv) return;
    // calcSharrDeriv: t0 = 3 (row-1 + row+1) + 10 row, t1 = row+1 - row-1; dx = t0[x+1] - t0[x-1], dy = 3 (t1[x-1] + t1[x+1]) + 10 t1[x];
    // rows and columns beyond the image are reflected (101)
    L.deriv.assign((size_t)(h + 2 * pad) * P * 2, 0);
    std::vector<int> t0(w + 2), t1(w + 2);
    for (int y = 0; y < h; y++) {
        const uint8_t* r0 = &src[(size_t)reflect101(y - 1, h) * w];
        const uint8_t* r1 = &src[(size_t)y * w];
        const uint8_t* r2 = &src[(size_t)reflect101(y + 1, h) * w];
        for (int x = -1; x <= w; x++) {
            const int xx = reflect101(x, w);
            t0[x + 1]    = (r0[xx] + r2[xx]) * 3 + r1[xx] * 10;
            t1[x + 1]    = r2[xx] - r0[xx];
        }
        for (int x = 0; x < w; x++) {
            int16_t* d = &L.deriv[((size_t)(y + pad) * P + x + pad) * 2];
            d[0]       = (int16_t)(t0[x + 2] - t0[x]);
            d[1]       = (int16_t)((t1[x + 2] + t1[x]) * 3 + t1[x + 1] * 10);
        }
    }
}
inline int kltDescale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
inline int kltRound(float v) { return (int)lrintf(v); }  // cvRound: nearest, ties to even
}  // namespace

int orc_klt_track(const uint8_t* refImg, const uint8_t* curImg, int w, int h, const float* prevPts, float* nextPts, int n,
                  const orc_klt_params* prm, uint8_t* status, float* err)
{
    const int win = prm->win;
    // TermCriteria clamps of calcOpticalFlowPyrLK; epsilon is compared against |delta|^2
    const int maxCount = std::min(std::max(prm->max_count, 0), 100);
    double eps         = std::min(std::max(prm->epsilon, 0.0), 10.0);
    eps *= eps;
    // buildOpticalFlowPyramid: stop before a level that is not larger than the window
    std::vector<std::vector<uint8_t>> pr(1), pc(1);
    std::vector<int> lw{w}, lh{h};
    pr[0].assign((size_t)w * h, 0), pc[0].assign((size_t)w * h, 0);
    std::memcpy(pr[0].data(), refImg, (size_t)w * h), std::memcpy(pc[0].data(), curImg, (size_t)w * h);
    int maxLevel = 0;
    for (int l = 1; l <= prm->max_level; l++) {
        const int sw = lw[l - 1], sh = lh[l - 1], dw = (sw + 1) / 2, dh = (sh + 1) / 2;
        if (dw <= win || dh <= win) break;
        pr.emplace_back((size_t)dw * dh), pc.emplace_back((size_t)dw * dh);
        orc_pyrdown(pr[l - 1].data(), sw, sh, sw, pr[l].data(), dw);
        orc_pyrdown(pc[l - 1].data(), sw, sh, sw, pc[l].data(), dw);
        lw.push_back(dw), lh.push_back(dh);
        maxLevel = l;
    }
    for (int i = 0; i < n; i++) {
        status[i] = 1;
        if (err) err[i] = 0.f;
    }
    const float halfWin  = (win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    std::vector<int16_t> Iw((size_t)win * win), dIw((size_t)win * win * 2);
    for (int level = maxLevel; level >= 0; level--) {
        KltLevel I, J;
        kltPad(pr[level], lw[level], lh[level], win, I, true);
        kltPad(pc[level], lw[level], lh[level], win, J, false);
        const int cols = lw[level], rows = lh[level];
        for (int pt = 0; pt < n; pt++) {
            const float sc = (float)(1. / (1 << level));
            float px = prevPts[2 * pt] * sc, py = prevPts[2 * pt + 1] * sc;
            float nx, ny;
            if (level == maxLevel) {
                if (prm->use_initial_flow)
                    nx = nextPts[2 * pt] * sc, ny = nextPts[2 * pt + 1] * sc;
                else
                    nx = px, ny = py;
            } else
                nx = nextPts[2 * pt] * 2.f, ny = nextPts[2 * pt + 1] * 2.f;
            nextPts[2 * pt] = nx, nextPts[2 * pt + 1] = ny;
            px -= halfWin, py -= halfWin;
            const int ipx = (int)std::floor(px), ipy = (int)std::floor(py);
            if (ipx < -win || ipx >= cols || ipy < -win || ipy >= rows) {
                if (level == 0) {
                    status[pt] = 0;
                    if (err) err[pt] = 0.f;
                }
                continue;
            }
            float a = px - ipx, b = py - ipy;
            int iw00 = kltRound((1.f - a) * (1.f - b) * (1 << 14)), iw01 = kltRound(a * (1.f - b) * (1 << 14));
            int iw10 = kltRound((1.f - a) * b * (1 << 14)), iw11 = (1 << 14) - iw00 - iw01 - iw10;
            float A11 = 0, A12 = 0, A22 = 0;
            for (int y = 0; y < win; y++)
                for (int x = 0; x < win; x++) {
                    const uint8_t* s0 = I.I(ipx + x, ipy + y);
                    const uint8_t* s1 = I.I(ipx + x, ipy + y + 1);
                    const int16_t* d0 = I.D(ipx + x, ipy + y);
                    const int16_t* d1 = I.D(ipx + x, ipy + y + 1);
                    const int ival  = kltDescale(s0[0] * iw00 + s0[1] * iw01 + s1[0] * iw10 + s1[1] * iw11, 14 - 5);
                    const int ixval = kltDescale(d0[0] * iw00 + d0[2] * iw01 + d1[0] * iw10 + d1[2] * iw11, 14);
                    const int iyval = kltDescale(d0[1] * iw00 + d0[3] * iw01 + d1[1] * iw10 + d1[3] * iw11, 14);
                    Iw[y * win + x]            = (int16_t)ival;
                    dIw[(y * win + x) * 2]     = (int16_t)ixval;
                    dIw[(y * win + x) * 2 + 1] = (int16_t)iyval;
                    A11 += (float)(ixval * ixval);
                    A12 += (float)(ixval * iyval);
                    A22 += (float)(iyval * iyval);
                }
            A11 *= FLT_SCALE, A12 *= FLT_SCALE, A22 *= FLT_SCALE;
            float D            = A11 * A22 - A12 * A12;
            const float minEig = (A22 + A11 - std::sqrt((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
            if (minEig < prm->min_eig_threshold || D < FLT_EPSILON) {
                if (level == 0) status[pt] = 0;
                continue;
            }
            D = 1.f / D;
            nx -= halfWin, ny -= halfWin;
            float pdx = 0.f, pdy = 0.f;
            for (int j = 0; j < maxCount; j++) {
                const int inx = (int)std::floor(nx), iny = (int)std::floor(ny);
                if (inx < -win || inx >= cols || iny < -win || iny >= rows) {
                    if (level == 0) status[pt] = 0;
                    break;
                }
                a = nx - inx, b = ny - iny;
                iw00 = kltRound((1.f - a) * (1.f - b) * (1 << 14)), iw01 = kltRound(a * (1.f - b) * (1 << 14));
                iw10 = kltRound((1.f - a) * b * (1 << 14)), iw11 = (1 << 14) - iw00 - iw01 - iw10;
                float b1 = 0, b2 = 0;
                for (int y = 0; y < win; y++)
                    for (int x = 0; x < win; x++) {
                        const uint8_t* s0 = J.I(inx + x, iny + y);
                        const uint8_t* s1 = J.I(inx + x, iny + y + 1);
                        const int diff = kltDescale(s0[0] * iw00 + s0[1] * iw01 + s1[0] * iw10 + s1[1] * iw11, 14 - 5) - Iw[y * win + x];
                        b1 += (float)(diff * dIw[(y * win + x) * 2]);
                        b2 += (float)(diff * dIw[(y * win + x) * 2 + 1]);
                    }
                b1 *= FLT_SCALE, b2 *= FLT_SCALE;
                const float dx = (float)((A12 * b2 - A22 * b1) * D), dy = (float)((A12 * b1 - A11 * b2) * D);
                nx += dx, ny += dy;
                nextPts[2 * pt] = nx + halfWin, nextPts[2 * pt + 1] = ny + halfWin;
                if ((double)dx * dx + (double)dy * dy <= eps) break;
                if (j > 0 && std::abs(dx + pdx) < 0.01 && std::abs(dy + pdy) < 0.01) {
                    nextPts[2 * pt] -= dx * 0.5f, nextPts[2 * pt + 1] -= dy * 0.5f;
                    break;
                }
                pdx = dx, pdy = dy;
            }
            if (status[pt] && err && level == 0) {
                const float ex = nextPts[2 * pt] - halfWin, ey = nextPts[2 * pt + 1] - halfWin;
                const int iex = (int)std::floor(ex), iey = (int)std::floor(ey);
                if (iex < -win || iex >= cols || iey < -win || iey >= rows) {
                    status[pt] = 0;
                    continue;
                }
                const float aa = ex - iex, bb = ey - iey;
                iw00 = kltRound((1.f - aa) * (1.f - bb) * (1 << 14)), iw01 = kltRound(aa * (1.f - bb) * (1 << 14));
                iw10 = kltRound((1.f - aa) * bb * (1 << 14)), iw11 = (1 << 14) - iw00 - iw01 - iw10;
                float errval = 0.f;
                for (int y = 0; y < win; y++)
                    for (int x = 0; x < win; x++) {
                        const uint8_t* s0 = J.I(iex + x, iey + y);
                        const uint8_t* s1 = J.I(iex + x, iey + y + 1);
                        const int diff = kltDescale(s0[0] * iw00 + s0[1] * iw01 + s1[0] * iw10 + s1[1] * iw11, 14 - 5) - Iw[y * win + x];
                        errval += std::abs((float)diff);
                    }
                err[pt] = errval * 1.f / (32 * win * win);
            }
        }
    }
    return maxLevel;
}

int orc_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
