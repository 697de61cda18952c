// pyramid.cu -- image + gradient pyramids, bit-exact to the reference's ImagePyramid::createImagePyramid
// (src/image_pyramid.cpp:36-52): Simd::AbsGradientSaturatedSum on the base image
// (3rd_party/simd/include/Simd/SimdLib.h:856-884), then cv::pyrDown on both stacks
// (5x5 [1 4 6 4 1]^2, (sum+128)>>8, BORDER_REFLECT_101, dst = (src+1)/2).
//
// Both kernels are HBM-bound integer stencils: every input byte is read once from DRAM (row reuse
// is served by L1/L2), every output byte written once.  Rows are padded to 16 B so 32-bit and
// 128-bit accesses are aligned.
#include "ctx.h"

namespace {

// dst = min(255, |s[y][x+1]-s[y][x-1]| + |s[y+1][x]-s[y-1][x]|), border 0.  One thread = 4 pixels,
// SIMD-in-word (__vabsdiffu4 / __vaddus4).  grid: (ceil(w/4/128), h, n_frames)
__global__ void __launch_bounds__(128) k_abs_gradient(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int w,
                                                      int h, int pitch, long long plane_stride)
{
    const int xw = blockIdx.x * blockDim.x + threadIdx.x;  // word index in the row
    const int y  = blockIdx.y;
    const int x0 = xw * 4;
    if (x0 >= w) return;
    const uint8_t* s = src + (long long)blockIdx.z * plane_stride;
    uint8_t* d       = dst + (long long)blockIdx.z * plane_stride;
    const int pw     = pitch >> 2;
    uint32_t out     = 0;
    if (y > 0 && y < h - 1) {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(s + (long long)y * pitch);
        const uint32_t* up  = reinterpret_cast<const uint32_t*>(s + (long long)(y - 1) * pitch);
        const uint32_t* dn  = reinterpret_cast<const uint32_t*>(s + (long long)(y + 1) * pitch);
        const uint32_t c    = __ldg(row + xw);
        const uint32_t l    = xw > 0 ? __ldg(row + xw - 1) : 0u;
        const uint32_t r    = xw + 1 < pw ? __ldg(row + xw + 1) : 0u;
        const uint32_t left  = __funnelshift_r(l, c, 24);  // bytes x-1 .. x+2
        const uint32_t right = __funnelshift_r(c, r, 8);   // bytes x+1 .. x+4
        const uint32_t gx    = __vabsdiffu4(right, left);
        const uint32_t gy    = __vabsdiffu4(__ldg(dn + xw), __ldg(up + xw));
        out                  = __vaddus4(gx, gy);
        // zero the first / last column and anything beyond the image width
        uint32_t mask = 0xffffffffu;
        if (x0 == 0) mask &= 0xffffff00u;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (x0 + k >= w - 1) mask &= ~(0xffu << (8 * k));
        out &= mask;
    }
    *reinterpret_cast<uint32_t*>(d + (long long)y * pitch + x0) = out;
}

// Dense (arbitrary pitch / alignment) source frames -> level 0 of the arena (16-B aligned pitched rows).
// One thread = 16 output bytes: five aligned 32-bit source words, funnel-shifted into place, one 128-bit store.
// grid: (ceil(pitch/16/128), h, n_frames)
__global__ void __launch_bounds__(128) k_repack(const uint8_t* __restrict__ src, long long src_pitch, long long src_frame_stride,
                                                uint8_t* __restrict__ dst, int w, int h, int pitch, long long plane_stride)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int y  = blockIdx.y;
    if (x0 >= pitch) return;
    const uint8_t* s    = src + (long long)blockIdx.z * src_frame_stride + (long long)y * src_pitch + x0;
    const uintptr_t ad  = reinterpret_cast<uintptr_t>(s);
    const uint32_t* s4  = reinterpret_cast<const uint32_t*>(ad & ~uintptr_t(3));
    const int off       = (int)(ad & 3u);
    const uint32_t sh   = (uint32_t)off * 8u;
    const int valid     = min(16, w - x0);  // bytes of this segment inside the image (<= 0: pure padding)
    const int need      = valid > 0 ? off + valid : 0;
    uint32_t v[5];
#pragma unroll
    for (int i = 0; i < 5; i++) v[i] = (i * 4 < need) ? __ldg(s4 + i) : 0u;  // never read a word without a needed byte
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        o[i] = __funnelshift_r(v[i], v[i + 1], sh);
        const int rem = valid - 4 * i;  // valid bytes in this word
        if (rem <= 0)
            o[i] = 0u;
        else if (rem < 4)
            o[i] &= (1u << (8 * rem)) - 1u;
    }
    *reinterpret_cast<uint4*>(dst + (long long)blockIdx.z * plane_stride + (long long)y * pitch + x0) = make_uint4(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// cv::pyrDown, one level, for both stacks of n frames.  A block produces a TX x TY output tile:
// the (2TX+3) x (2TY+3) source tile is staged in shared memory (reflect-101 applied while loading),
// the horizontal [1 4 6 4 1] pass runs into a second shared buffer (16-bit), the vertical pass
// writes (sum + 128) >> 8.  grid: (ceil(dw/TX), ceil(dh/TY), 2 * n_frames)
constexpr int TX = 64, TY = 8;
constexpr int SW = 2 * TX + 3, SH = 2 * TY + 3;

__global__ void __launch_bounds__(256) k_pyrdown(const uint8_t* __restrict__ src_img, uint8_t* __restrict__ dst_img,
                                                 const uint8_t* __restrict__ src_grad, uint8_t* __restrict__ dst_grad,
                                                 int sw, int sh, int spitch, long long sstride, int dw, int dh,
                                                 int dpitch, long long dstride)
{
    __shared__ uint8_t tile[SH][SW + 1];
    __shared__ uint16_t hsum[SH][TX];
    const int frame     = blockIdx.z >> 1;
    const bool grad     = blockIdx.z & 1;
    const uint8_t* src  = (grad ? src_grad : src_img) + (long long)frame * sstride;
    uint8_t* dst        = (grad ? dst_grad : dst_img) + (long long)frame * dstride;
    const int ox        = blockIdx.x * TX;
    const int oy        = blockIdx.y * TY;
    const int sx0       = 2 * ox - 2;
    const int sy0       = 2 * oy - 2;
    for (int i = threadIdx.x; i < SH * SW; i += blockDim.x) {
        const int ty = i / SW, tx = i - ty * SW;
        const int gx = reflect101(sx0 + tx, sw);
        const int gy = reflect101(sy0 + ty, sh);
        tile[ty][tx] = __ldg(src + (long long)gy * spitch + gx);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SH * TX; i += blockDim.x) {
        const int ty = i / TX, tx = i - ty * TX;
        const uint8_t* p = &tile[ty][2 * tx];
        hsum[ty][tx]     = (uint16_t)(p[0] + 4 * p[1] + 6 * p[2] + 4 * p[3] + p[4]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TY * TX; i += blockDim.x) {
        const int ty = i / TX, tx = i - ty * TX;
        const int x = ox + tx, y = oy + ty;
        if (x < dw && y < dh) {
            const int s = hsum[2 * ty][tx] + 4 * hsum[2 * ty + 1][tx] + 6 * hsum[2 * ty + 2][tx] +
                          4 * hsum[2 * ty + 3][tx] + hsum[2 * ty + 4][tx];
            dst[(long long)y * dpitch + x] = (uint8_t)((s + 128) >> 8);
        }
    }
}

}  // namespace

svo_status launch_pyramid_build(svo_ctx* ctx, int first_slot, int n)
{
    const PyramidArena& a = ctx->arena;
    const LevelGeom& g0   = a.geom[0];
    {
        dim3 grid((g0.pitch / 4 + 127) / 128, g0.h, n);
        k_abs_gradient<<<grid, 128, 0, ctx->stream>>>(a.img[0] + first_slot * g0.plane_stride,
                                                      a.grad[0] + first_slot * g0.plane_stride, g0.w, g0.h, g0.pitch,
                                                      g0.plane_stride);
        ctx->launches++;
    }
    for (int l = 1; l < a.levels; l++) {
        const LevelGeom& s = a.geom[l - 1];
        const LevelGeom& d = a.geom[l];
        dim3 grid((d.w + TX - 1) / TX, (d.h + TY - 1) / TY, 2 * n);
        k_pyrdown<<<grid, 256, 0, ctx->stream>>>(a.img[l - 1] + first_slot * s.plane_stride,
                                                 a.img[l] + first_slot * d.plane_stride,
                                                 a.grad[l - 1] + first_slot * s.plane_stride,
                                                 a.grad[l] + first_slot * d.plane_stride, s.w, s.h, s.pitch,
                                                 s.plane_stride, d.w, d.h, d.pitch, d.plane_stride);
        ctx->launches++;
    }
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}

svo_status launch_repack(svo_ctx* ctx, const uint8_t* dsrc, long long src_pitch, long long src_frame_stride, int first_slot, int n)
{
    const LevelGeom& g = ctx->arena.geom[0];
    dim3 grid((g.pitch / 16 + 127) / 128, g.h, n);
    k_repack<<<grid, 128, 0, ctx->stream>>>(dsrc, src_pitch, src_frame_stride, ctx->arena.img[0] + (int64_t)first_slot * g.plane_stride,
                                            g.w, g.h, g.pitch, g.plane_stride);
    ctx->launches++;
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}
