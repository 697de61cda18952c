// klt.cu -- the tracker of algorithm::computeOpticalFlowSparse (src/algorithm.cpp:29-107; the same call at src/map.cpp:322,403):
//   cv::calcOpticalFlowPyrLK(refImg, curImg, refPoints, curPoints, status, errors, Size(win, win), 3,
//                            TermCriteria(COUNT + EPS, 30, 1e-4), OPTFLOW_USE_INITIAL_FLOW)
// OpenCV's pyramidal Lucas-Kanade (4.x modules/video/src/lkpyramid.cpp), one warp per point and ONE launch for all levels:
// the image pyramids OpenCV rebuilds on every call (buildOpticalFlowPyramid = cv::pyrDown per level) are the ones the frame
// slots already hold.  Per level and point:
//   stage    the Scharr derivatives (calcSharrDeriv: 3-10-3, int16, reflected at the image border, zero outside the image)
//            of the (win+1)^2 neighbourhood of the reference point into shared memory -- OpenCV computes and pads a full
//            derivative image per level; only ~n (win+1)^2 of its pixels are ever read;
//   template window values (x32) and derivatives with 14-bit fixed-point bilinear weights, exactly OpenCV's integers;
//            A = sum of derivative products, summed exactly in int64 (OpenCV sums float lanes: same value up to rounding);
//   iterate  b = sum (J - I) dI over the window of the current image (reflect-101 padding outside), delta = A^-1 b in FP32
//            without contraction, the |delta|^2 <= eps^2 exit and the oscillation exit with its half step back;
//   level 0  status and the mean absolute window difference.
// Lanes stride over the window pixels; every lane carries the (uniform) point state.
#include <float.h>

#include "ctx.h"

namespace {

struct KltArgs {
    ArenaView view;
    int refSlot, curSlot;
    const float2* prev;
    float2* next;
    uint8_t* status;
    float* err;
    int n;
    int win, topLevel, maxCount, useInitialFlow;
    double epsSq;   // |delta|^2 is formed and compared in double, as OpenCV does (Point2f::ddot)
    double minEig;  // minEigThreshold, compared against the float eigenvalue in double
};

__device__ __forceinline__ int klt_reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
__device__ __forceinline__ int klt_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

struct KltImg {
    const uint8_t* p;
    int w, h, pitch;
    // BORDER_REFLECT_101 padding of buildOpticalFlowPyramid
    __device__ __forceinline__ int at(int x, int y) const { return __ldg(p + (long long)klt_reflect(y, h) * pitch + klt_reflect(x, w)); }
};

__device__ __forceinline__ void klt_weights(float a, float b, int& w00, int& w01, int& w10, int& w11)
{
    const float a1 = __fsub_rn(1.f, a), b1 = __fsub_rn(1.f, b);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(a1, b1), 16384.f));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, b1), 16384.f));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(a1, b), 16384.f));
    w11 = 16384 - w00 - w01 - w10;
}

// exact warp total of 32-bit lane values (the total may need up to 37 bits): two hardware reductions, of the low
// 16 bits and of the (signed) rest, instead of five 64-bit shuffle steps on the serial path of every iteration
__device__ __forceinline__ long long warp_sum_i32(int v)
{
    const int lo = __reduce_add_sync(0xffffffffu, v & 0xffff);
    const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
    return (long long)hi * 65536 + (long long)lo;
}

// window value of image J at integer origin (ox, oy) + (x, y), x32.  INTERIOR (uniform over the warp): the whole
// (win+1)^2 footprint lies inside the image, no reflection
template <bool INTERIOR>
__device__ __forceinline__ int klt_sample(const KltImg& J, int ox, int oy, int x, int y, int w00, int w01, int w10, int w11)
{
    const int X = ox + x, Y = oy + y;
    int v;
    if (INTERIOR) {
        const uint8_t* p = J.p + Y * J.pitch + X;
        v = (int)__ldg(p) * w00 + (int)__ldg(p + 1) * w01 + (int)__ldg(p + J.pitch) * w10 + (int)__ldg(p + J.pitch + 1) * w11;
    } else {
        v = J.at(X, Y) * w00 + J.at(X + 1, Y) * w01 + J.at(X, Y + 1) * w10 + J.at(X + 1, Y + 1) * w11;
    }
    return klt_descale(v, 14 - 5);
}
__device__ __forceinline__ bool klt_interior(const KltImg& J, int ox, int oy, int win)
{
    return ox >= 0 && oy >= 0 && ox + win < J.w && oy + win < J.h;
}

__global__ void __launch_bounds__(128) k_klt_track(const KltArgs a)
{
    extern __shared__ __align__(16) unsigned char klt_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pt = blockIdx.x * 4 + warp;
    if (pt >= a.n) return;
    const int win = a.win, W1 = win + 1, area = win * win;
    const int rcpWin = (65536 + win - 1) / win;  // i / win = (i rcpWin) >> 16 exactly for i < 441, win <= 21
    // per warp: derivative pairs of the (win+1)^2 neighbourhood, then the template (value, dx, dy)
    const size_t perWarp = ((size_t)W1 * W1 * 4 + (size_t)area * 6 + 31) & ~size_t(15);
    short2* dS  = reinterpret_cast<short2*>(klt_smem + perWarp * warp);
    short* Iw   = reinterpret_cast<short*>(reinterpret_cast<unsigned char*>(dS) + (size_t)W1 * W1 * 4);
    short2* dIw = reinterpret_cast<short2*>(Iw + ((area + 1) & ~1));

    const float2 prev0 = a.prev[pt];
    float2 nxt         = a.next[pt];  // the caller's initial guess (OPTFLOW_USE_INITIAL_FLOW)
    const float halfWin = (float)(win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    bool status = true;
    float err   = 0.f;

    for (int level = a.topLevel; level >= 0; level--) {
        KltImg I, J;
        I.w = J.w = a.view.w[level];
        I.h = J.h = a.view.h[level];
        I.pitch = J.pitch = a.view.pitch[level];
        I.p = a.view.img[level] + (long long)a.refSlot * a.view.plane_stride[level];
        J.p = a.view.img[level] + (long long)a.curSlot * a.view.plane_stride[level];
        const int cols = I.w, rows = I.h;
        const float sc = (float)(1. / (double)(1 << level));
        float px = __fmul_rn(prev0.x, sc), py = __fmul_rn(prev0.y, sc);
        float nx, ny;
        if (level == a.topLevel) {
            if (a.useInitialFlow)
                nx = __fmul_rn(nxt.x, sc), ny = __fmul_rn(nxt.y, sc);
            else
                nx = px, ny = py;
        } else {
            nx = __fmul_rn(nxt.x, 2.f), ny = __fmul_rn(nxt.y, 2.f);
        }
        nxt = make_float2(nx, ny);
        px  = __fsub_rn(px, halfWin);
        py  = __fsub_rn(py, halfWin);
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -win || ipx >= cols || ipy < -win || ipy >= rows) {
            if (level == 0) {
                status = false;
                err    = 0.f;
            }
            continue;
        }
        __syncwarp();
        // --- Scharr derivatives of the neighbourhood (zero outside the image, reflected rows / columns at its border) ---
        for (int i = lane; i < W1 * W1; i += 32) {
            const int yy = i / W1, xx = i - yy * W1;
            const int X = ipx + xx, Y = ipy + yy;
            short2 d = make_short2(0, 0);
            if (X >= 0 && X < cols && Y >= 0 && Y < rows) {
                const int xm = klt_reflect(X - 1, cols), xp = klt_reflect(X + 1, cols);
                const uint8_t* r0 = I.p + (long long)klt_reflect(Y - 1, rows) * I.pitch;
                const uint8_t* r1 = I.p + (long long)Y * I.pitch;
                const uint8_t* r2 = I.p + (long long)klt_reflect(Y + 1, rows) * I.pitch;
                const int a0 = __ldg(r0 + xm), a1 = __ldg(r0 + X), a2 = __ldg(r0 + xp);
                const int b0 = __ldg(r1 + xm), b2 = __ldg(r1 + xp);
                const int c0 = __ldg(r2 + xm), c1 = __ldg(r2 + X), c2 = __ldg(r2 + xp);
                d.x = (short)(((a2 + c2) * 3 + b2 * 10) - ((a0 + c0) * 3 + b0 * 10));
                d.y = (short)(((c0 - a0) + (c2 - a2)) * 3 + (c1 - a1) * 10);
            }
            dS[i] = d;
        }
        __syncwarp();
        // --- template and the matrix of derivative products ---
        int w00, w01, w10, w11;
        klt_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);
        long long s11 = 0, s12 = 0, s22 = 0;
        const bool inI = klt_interior(I, ipx, ipy, win);
        for (int i = lane; i < area; i += 32) {
            const int y = (i * rcpWin) >> 16, x = i - y * win;
            const int ival = inI ? klt_sample<true>(I, ipx, ipy, x, y, w00, w01, w10, w11) : klt_sample<false>(I, ipx, ipy, x, y, w00, w01, w10, w11);
            const short2 d00 = dS[y * W1 + x], d01 = dS[y * W1 + x + 1], d10 = dS[(y + 1) * W1 + x], d11 = dS[(y + 1) * W1 + x + 1];
            const int ix = klt_descale(d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11, 14);
            const int iy = klt_descale(d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11, 14);
            Iw[i]  = (short)ival;
            dIw[i] = make_short2((short)ix, (short)iy);
            s11 += (long long)(ix * ix);
            s12 += (long long)(ix * iy);
            s22 += (long long)(iy * iy);
        }
        // per lane <= ceil(441 / 32) = 14 products of <= 4,080^2: 2.3e8, inside 32 bits
        s11 = warp_sum_i32((int)s11), s12 = warp_sum_i32((int)s12), s22 = warp_sum_i32((int)s22);
        __syncwarp();
        const float A11 = __fmul_rn((float)s11, FLT_SCALE), A12 = __fmul_rn((float)s12, FLT_SCALE), A22 = __fmul_rn((float)s22, FLT_SCALE);
        float D         = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA  = __fsub_rn(A11, A22);
        const float rt  = __fsqrt_rn(__fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12)));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), rt), (float)(2 * win * win));
        if ((double)minEig < a.minEig || D < FLT_EPSILON) {
            if (level == 0) status = false;
            continue;
        }
        D  = __fdiv_rn(1.f, D);
        nx = __fsub_rn(nx, halfWin);
        ny = __fsub_rn(ny, halfWin);
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < a.maxCount; j++) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -win || inx >= cols || iny < -win || iny >= rows) {
                if (level == 0) status = false;
                break;
            }
            klt_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            // per lane <= 14 products of <= 8,160 x 4,080 = 3.3e7: 4.7e8, inside 32 bits
            int pb1 = 0, pb2 = 0;
            if (klt_interior(J, inx, iny, win)) {
                // four window pixels per lane in flight, branch-free (a lane past the end repeats the last pixel with
                // weight 0): the iteration is a chain of dependent latencies, not of instructions
                for (int i0 = lane; i0 < area; i0 += 128) {
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int i  = i0 + 32 * u;
                        const int ii = min(i, area - 1);
                        const int y = (ii * rcpWin) >> 16, x = ii - y * win;
                        int diff = klt_sample<true>(J, inx, iny, x, y, w00, w01, w10, w11) - (int)Iw[ii];
                        diff     = i < area ? diff : 0;
                        const short2 d = dIw[ii];
                        pb1 += diff * (int)d.x;
                        pb2 += diff * (int)d.y;
                    }
                }
            } else {
                for (int i = lane; i < area; i += 32) {
                    const int y = (i * rcpWin) >> 16, x = i - y * win;
                    const int diff = klt_sample<false>(J, inx, iny, x, y, w00, w01, w10, w11) - (int)Iw[i];
                    const short2 d = dIw[i];
                    pb1 += diff * (int)d.x;
                    pb2 += diff * (int)d.y;
                }
            }
            const long long sb1 = warp_sum_i32(pb1), sb2 = warp_sum_i32(pb2);
            const float b1 = __fmul_rn((float)sb1, FLT_SCALE), b2 = __fmul_rn((float)sb2, FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx  = __fadd_rn(nx, dx);
            ny  = __fadd_rn(ny, dy);
            nxt = make_float2(__fadd_rn(nx, halfWin), __fadd_rn(ny, halfWin));
            if ((double)dx * (double)dx + (double)dy * (double)dy <= a.epsSq) break;
            if (j > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
                nxt.x = __fsub_rn(nxt.x, __fmul_rn(dx, 0.5f));
                nxt.y = __fsub_rn(nxt.y, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx, pdy = dy;
        }
        if (status && level == 0) {
            const float ex = __fsub_rn(nxt.x, halfWin), ey = __fsub_rn(nxt.y, halfWin);
            const int iex = (int)floorf(ex), iey = (int)floorf(ey);
            if (iex < -win || iex >= cols || iey < -win || iey >= rows) {
                status = false;
                continue;
            }
            klt_weights(__fsub_rn(ex, (float)iex), __fsub_rn(ey, (float)iey), w00, w01, w10, w11);
            int pe = 0;
            for (int i = lane; i < area; i += 32) {
                const int y = (i * rcpWin) >> 16, x = i - y * win;
                pe += abs(klt_sample<false>(J, iex, iey, x, y, w00, w01, w10, w11) - (int)Iw[i]);
            }
            const long long se = warp_sum_i32(pe);
            err = __fdiv_rn((float)se, (float)(32 * win * win));
        }
    }
    if (lane == 0) {
        a.next[pt]   = nxt;
        a.status[pt] = status ? 1 : 0;
        a.err[pt]    = err;
    }
}

}  // namespace

size_t klt_smem_bytes(int win)
{
    const int W1 = win + 1, area = win * win;
    const size_t perWarp = ((size_t)W1 * W1 * 4 + (size_t)area * 6 + 31) & ~size_t(15);
    return perWarp * 4;
}

svo_status launch_klt_track(svo_ctx* ctx, int refSlot, int curSlot, int n, const svo_klt_params& prm, int topLevel)
{
    if (n == 0) return SVO_OK;
    KltArgs args;
    args.view     = make_view(ctx->arena);
    args.refSlot  = refSlot;
    args.curSlot  = curSlot;
    args.prev     = ctx->d_klt_prev;
    args.next     = ctx->d_klt_next;
    args.status   = ctx->d_klt_status;
    args.err      = ctx->d_klt_err;
    args.n        = n;
    args.win      = prm.win;
    args.topLevel = topLevel;
    // TermCriteria handling of calcOpticalFlowPyrLK: count in 0..100, epsilon in 0..10, squared
    args.maxCount = std::min(std::max(prm.max_count, 0), 100);
    const double eps = std::min(std::max(prm.epsilon, 0.0), 10.0);
    args.epsSq    = eps * eps;
    args.useInitialFlow = prm.use_initial_flow ? 1 : 0;
    args.minEig   = prm.min_eig_threshold;
    k_klt_track<<<(n + 3) / 4, 128, klt_smem_bytes(prm.win), ctx->stream>>>(args);
    ctx->launches++;
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}
