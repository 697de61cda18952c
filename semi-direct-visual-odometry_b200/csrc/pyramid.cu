// pyramid.cu -- image + gradient pyramids, bit-exact to the reference's ImagePyramid::createImagePyramid
// (src/image_pyramid.cpp:36-52): Simd::AbsGradientSaturatedSum on the base image
// (3rd_party/simd/include/Simd/SimdLib.h:856-884), then cv::pyrDown on both stacks
// (5x5 [1 4 6 4 1]^2, (sum+128)>>8, BORDER_REFLECT_101, dst = (src+1)/2).
//
// Both kernels are HBM-bound integer stencils: every input byte is read once from DRAM (row reuse
// is served by L1/L2), every output byte written once.  Rows are padded to 16 B so 32-bit and
// 128-bit accesses are aligned.
#include <algorithm>
#include <type_traits>
#include <stdlib.h>

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

// ------------------------------------------------------------------------------------------------------------
// cv::pyrDown of both stacks, one level per launch, word-vectorised.  A CTA (256 threads) produces a 64 x 16 output
// tile of the image AND of the gradient stack.  BASE (level 0 -> 1) reads ONLY the level-0 image: the gradient
// (Simd::AbsGradientSaturatedSum) of the tile is computed in shared memory with SIMD-in-word byte arithmetic, its
// exclusive 128 x 32 area is written out as gradient level 0 with 16-byte stores, and both tiles are decimated -- so a
// frame's level-0 image is read from HBM once and the level-0 gradient never re-read.
//   stage 1  aligned 128-bit loads of the source tile(s) (rows/columns outside the image are skipped)
//   stage 2  BASE: gradient words from three image rows (funnel shifts + __vabsdiffu4 + __vaddus4), border pixels 0
//   stage 3  BORDER_REFLECT_101: the <= 2 rows / columns a 5-tap kernel reaches outside the image are mirrored inside
//            shared memory (border CTAs only)
//   stage 4  horizontal [1 4 6 4 1] into 16-bit sums, two outputs per thread from three words
//   stage 5  vertical [1 4 6 4 1], (sum + 128) >> 8, four outputs per thread, one 32-bit store
// Tile columns are words; word 4 of a tile row holds source columns 2 ox .. 2 ox + 3 (so the exclusive area starts
// 16-byte aligned in shared memory), word 3 the four columns before, word 36 the four after.
// ------------------------------------------------------------------------------------------------------------
constexpr int PT_X = 64, PT_Y = 16;      // output tile
constexpr int PT_ROWS = 2 * PT_Y + 3;    // 35 source rows feed the vertical taps
constexpr int PT_PW = 40;                // tile row pitch in words (multiple of 4: 16-byte aligned rows)
constexpr int PT_W0 = 3, PT_W1 = 37;     // loaded word columns [PT_W0, PT_W1)

template <bool BASE>
__global__ void __launch_bounds__(256) k_pyr_level(const uint8_t* __restrict__ src_img, const uint8_t* __restrict__ src_grad,
                                                   uint8_t* __restrict__ dst_img, uint8_t* __restrict__ dst_grad,
                                                   uint8_t* __restrict__ grad0,  // BASE: gradient level 0 (written)
                                                   int sw, int sh, int spitch, long long sstride, int dw, int dh, int dpitch,
                                                   long long dstride)
{
    constexpr int IROWS = BASE ? PT_ROWS + 2 : PT_ROWS;  // BASE: one more image row above and below for the gradient
    __shared__ __align__(16) uint32_t tI[IROWS][PT_PW];
    __shared__ __align__(16) uint32_t tG[PT_ROWS][PT_PW];
    __shared__ __align__(16) uint16_t hI[PT_ROWS][PT_X];
    __shared__ __align__(16) uint16_t hG[PT_ROWS][PT_X];
    const int tid   = threadIdx.x;
    const int frame = blockIdx.z;
    const int ox = blockIdx.x * PT_X, oy = blockIdx.y * PT_Y;
    const int x00 = 2 * ox - 16;                  // source column of tile word 0, byte 0
    const int y00 = 2 * oy - 2;                   // source row of tG / hI row 0
    const int iy0 = BASE ? y00 - 1 : y00;         // source row of tI row 0
    const uint8_t* sI = src_img + (long long)frame * sstride;
    const uint8_t* sG = BASE ? nullptr : src_grad + (long long)frame * sstride;
    const int spw = spitch >> 2;

    // ---- stage 1: loads, 16 bytes each (10 per tile row; the first and last reach past the words the taps need and
    // are dropped where they would leave the pitched row) ----
    const int spq = spitch >> 4;  // 16-byte groups per source row
    for (int i = tid; i < IROWS * (PT_PW / 4); i += 256) {
        const int r = i / (PT_PW / 4), qc = i - r * (PT_PW / 4);
        const int y = iy0 + r, xq = (x00 >> 4) + qc;  // source 16-byte column
        uint4 v     = make_uint4(0u, 0u, 0u, 0u);
        if (y >= 0 && y < sh && xq >= 0 && xq < spq) v = __ldg(reinterpret_cast<const uint4*>(sI + (long long)y * spitch) + xq);
        *reinterpret_cast<uint4*>(&tI[r][4 * qc]) = v;
    }
    if (!BASE) {
        for (int i = tid; i < PT_ROWS * (PT_PW / 4); i += 256) {
            const int r = i / (PT_PW / 4), qc = i - r * (PT_PW / 4);
            const int y = y00 + r, xq = (x00 >> 4) + qc;
            uint4 v     = make_uint4(0u, 0u, 0u, 0u);
            if (y >= 0 && y < sh && xq >= 0 && xq < spq) v = __ldg(reinterpret_cast<const uint4*>(sG + (long long)y * spitch) + xq);
            *reinterpret_cast<uint4*>(&tG[r][4 * qc]) = v;
        }
    }
    __syncthreads();

    // ---- stage 2 (BASE): gradient of the tile.  Warp w takes tile rows w, w + 8, ...; lane l the words 3 + l (and,
    // in a second sweep, 35 + l for l < 2): no index arithmetic per item, conflict-free rows ----
    if (BASE) {
        const int lane = tid & 31, wrp = tid >> 5;
        const bool rowBorder = y00 <= 0 || y00 + PT_ROWS >= sh - 1;  // the tile touches the first / last image row
#pragma unroll
        for (int sweep = 0; sweep < 2; sweep++) {
            const int wc = PT_W0 + 32 * sweep + lane;
            if (wc < PT_W1) {
                const int x0       = x00 + 4 * wc;
                const bool inX     = x0 + 3 >= 0 && x0 < sw;
                const bool colEdge = x0 <= 0 || x0 + 3 >= sw - 1;  // first / last image column inside this word
                uint32_t mask      = 0xffffffffu;
                if (colEdge) {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (x0 + k <= 0 || x0 + k >= sw - 1) mask &= ~(0xffu << (8 * k));
                }
                for (int r = wrp; r < PT_ROWS; r += 8) {
                    const uint32_t c = tI[r + 1][wc];
                    const uint32_t l = tI[r + 1][wc - 1];  // words PT_W0 - 1 and PT_W1 exist in the tile (16-byte loads)
                    const uint32_t q = tI[r + 1][wc + 1];
                    const uint32_t left  = __funnelshift_r(l, c, 24);  // columns x-1 .. x+2
                    const uint32_t right = __funnelshift_r(c, q, 8);   // columns x+1 .. x+4
                    uint32_t out = __vaddus4(__vabsdiffu4(right, left), __vabsdiffu4(tI[r + 2][wc], tI[r][wc])) & mask;
                    // first / last row and column of the image and everything beyond: 0 (SimdLib.h:856-884)
                    if (!inX) out = 0u;
                    if (rowBorder) {
                        const int y = y00 + r;
                        if (!(y > 0 && y < sh - 1)) out = 0u;
                    }
                    tG[r][wc] = out;
                }
            }
        }
        __syncthreads();
    }

    // ---- stage 3: BORDER_REFLECT_101 inside shared memory (border CTAs only) ----
    const int xlo = 2 * ox - 2, xhi = 2 * ox + 2 * PT_X;  // source columns / rows the taps of this tile reach
    const int ylo = y00, yhi = y00 + PT_ROWS - 1;
    if (xlo < 0 || xhi >= sw || ylo < 0 || yhi >= sh) {
        uint8_t* bI = reinterpret_cast<uint8_t*>(&tI[BASE ? 1 : 0][0]);  // row 0 = source row y00 in both tiles
        uint8_t* bG = reinterpret_cast<uint8_t*>(&tG[0][0]);
        // columns first (in-image rows), then whole rows (which then carry the mirrored columns)
        for (int i = tid; i < PT_ROWS * 4; i += 256) {
            const int r = i >> 2, k = i & 3;  // k: 0,1 -> columns -2,-1;  2,3 -> columns sw, sw+1
            const int x = k < 2 ? k - 2 : sw + (k - 2);
            const int y = y00 + r;
            if (x >= xlo && x <= xhi && (x < 0 || x >= sw) && y >= 0 && y < sh) {
                const int xs = reflect101(x, sw);
                bI[r * PT_PW * 4 + (x - x00)] = bI[r * PT_PW * 4 + (xs - x00)];
                bG[r * PT_PW * 4 + (x - x00)] = bG[r * PT_PW * 4 + (xs - x00)];
            }
        }
        __syncthreads();
        for (int i = tid; i < 4 * (PT_W1 - PT_W0); i += 256) {
            const int k = i / (PT_W1 - PT_W0), wc = PT_W0 + i - k * (PT_W1 - PT_W0);
            const int y = k < 2 ? k - 2 : sh + (k - 2);
            if (y >= ylo && y <= yhi && (y < 0 || y >= sh)) {
                const int ys = reflect101(y, sh);
                tI[(BASE ? 1 : 0) + y - y00][wc] = tI[(BASE ? 1 : 0) + ys - y00][wc];
                tG[y - y00][wc]                  = tG[ys - y00][wc];
            }
        }
        __syncthreads();
    }

    // ---- BASE: gradient level 0, the exclusive 128 x 32 area, 16-byte stores ----
    if (BASE) {
        const int j = tid >> 3, q = tid & 7;  // 32 rows x 8 segments of 16 bytes
        const int y = 2 * oy + j, x = 2 * ox + 16 * q;
        if (y < sh && x < spitch) {
            uint4 v = *reinterpret_cast<const uint4*>(&tG[2 + j][4 + 4 * q]);
            // nothing but zeros beyond the image width (columns inside the row padding)
            uint32_t* vv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int rem = sw - (x + 4 * k);
                if (rem <= 0)
                    vv[k] = 0u;
                else if (rem < 4)
                    vv[k] &= (1u << (8 * rem)) - 1u;
            }
            *reinterpret_cast<uint4*>(grad0 + (long long)frame * sstride + (long long)y * spitch + x) = v;
        }
    }

    // ---- stage 4: horizontal pass, outputs (tx, tx + 1) from tile bytes 2 tx + 14 .. 2 tx + 20.  Lane = output pair
    // tp (tx = 2 tp), warp w takes rows w, w + 8, ... of both stacks ----
    {
        const int tp = tid & 31, wrp = tid >> 5;
        for (int r = wrp; r < PT_ROWS; r += 8) {
#pragma unroll
            for (int st = 0; st < 2; st++) {
                const uint32_t* row = st ? &tG[r][0] : &tI[(BASE ? 1 : 0) + r][0];
                const uint32_t w0 = row[3 + tp], w1 = row[4 + tp], w2 = row[5 + tp];  // tile bytes 4 tp + 12 .. 4 tp + 23
                // taps of output tx: bytes 2, 3 of w0 and 0, 1, 2 of w1; of tx + 1: bytes 0 .. 3 of w1 and 0 of w2
                // (dp4a: four byte products per instruction)
                const uint32_t h0 = __dp4a(w0, 0x04010000u, __dp4a(w1, 0x00010406u, 0u));
                const uint32_t h1 = __dp4a(w1, 0x04060401u, w2 & 0xffu);
                uint16_t* hrow    = st ? &hG[r][0] : &hI[r][0];
                *reinterpret_cast<uint32_t*>(hrow + 2 * tp) = h0 | (h1 << 16);
            }
        }
    }
    __syncthreads();

    // ---- stage 5: vertical pass, four outputs per thread ----
    for (int i = tid; i < 2 * PT_Y * (PT_X / 4); i += 256) {
        const int st = i / (PT_Y * (PT_X / 4));
        const int j  = i - st * (PT_Y * (PT_X / 4));
        const int ty = j / (PT_X / 4), tq = j - ty * (PT_X / 4);
        const int x = ox + 4 * tq, y = oy + ty;
        if (y >= dh || x >= dpitch) continue;
        const uint16_t(*h)[PT_X] = st ? hG : hI;
        // two 16-bit lanes per register: the horizontal sums are <= 4,080, the weighted column sum + 128 <= 65,408
        uint32_t a0 = 0x00800080u, a1 = 0x00800080u;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const uint2 v    = *reinterpret_cast<const uint2*>(&h[2 * ty + k][4 * tq]);
            const uint32_t c = k == 0 || k == 4 ? 1u : (k == 2 ? 6u : 4u);
            a0 += c * v.x;
            a1 += c * v.y;
        }
        uint32_t out = __byte_perm(a0, a1, 0x7531);  // the high byte of each lane = (sum + 128) >> 8
        if (x + 3 >= dw) {                            // zeros in the row padding
            const int rem = dw - x;
            out           = rem <= 0 ? 0u : (out & ((1u << (8 * rem)) - 1u));
        }
        uint8_t* d = (st ? dst_grad : dst_img) + (long long)frame * dstride + (long long)y * dpitch + x;
        *reinterpret_cast<uint32_t*>(d) = out;
    }
}

// ------------------------------------------------------------------------------------------------------------
// cv::pyrDown of both stacks for BATCHES of frames, register-marching: a warp (= one CTA) owns a strip of 256 source columns
// (a lane holds eight: two words; one lane of overlap on either side, the inner 30 lanes give 120 output columns) and walks
// down the rows of its row segment.  Per source row and lane: one coalesced 64-bit load per stack (issued two output rows
// ahead), the two neighbour words by shuffle, the horizontal [1 4 6 4 1] of the lane's four outputs as six dp4a; the last
// four rows of horizontal sums stay in registers (two 16-bit lanes per register), every second row emits one output row:
// (sum + 128) >> 8, four bytes per lane, one aligned 32-bit store.  No shared memory, no barrier, no index arithmetic beyond
// row pointers.  BORDER_REFLECT_101: rows by folding the row index (border segments only), columns by one byte permute per
// word whose selector depends on the lane only (border strips only).  The tile kernel above needs ~190 instructions per
// source word of both stacks; this one ~40.  grid: (strips, row segments, frames), block 32.
// ------------------------------------------------------------------------------------------------------------
constexpr int PM_INNER = 30;   // inner lanes of a strip (8 source columns, 4 output columns each)

__device__ __forceinline__ uint32_t pm_hsum(uint32_t wl, uint32_t w, uint32_t wr)
{
    const uint32_t h0 = __dp4a(wl, 0x04010000u, __dp4a(w, 0x00010406u, 0u));  // centre = byte 0 of w
    const uint32_t h1 = __dp4a(w, 0x04060401u, wr & 0xffu);                    // centre = byte 2 of w
    return h0 | (h1 << 16);
}

// BASE (level 0 -> 1): reads ONLY the level-0 image; the gradient (Simd::AbsGradientSaturatedSum) of every row is computed
// in registers from three image rows, written out as gradient level 0 by the lane that owns the word, and decimated with
// the image.  Rows and columns beyond the image follow BORDER_REFLECT_101 of the GRADIENT image (whose first / last row and
// column are 0): walking virtual rows over the reflected image reproduces it, except for those zeros, which are forced.
// EDGE: the strips that hold the first columns (strip 0) or columns beyond the image (from strip `edge0` on) apply the column
// reflection; the strips in between (launched separately: nothing but the blockIdx differs) carry none of that code.  A CTA is
// ONE warp and every loop bound derives from blockIdx: the compiler sees warp-uniform control flow around the shuffles.
template <bool BASE, bool EDGE>
__device__ __forceinline__ void pyr_march_strip(const uint8_t* __restrict__ src_img, const uint8_t* __restrict__ src_grad,
                                                uint8_t* __restrict__ dst_img, uint8_t* __restrict__ dst_grad,
                                                uint8_t* __restrict__ grad0, int sw, int sh, int spitch, long long sstride, int dw,
                                                int dh, int dpitch, long long dstride, int rows)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x, frame = blockIdx.z;
    const int strip = blockIdx.x;
    // a lane holds EIGHT source columns (two words): columns 8 L .. 8 L + 7 with L = strip * PM_INNER - 1 + lane; lanes 1 .. 30
    // are the inner ones: four outputs each (columns 4 L .. 4 L + 3 of the next level), one aligned 32-bit store
    const int L   = strip * PM_INNER - 1 + lane;
    const int oy0 = blockIdx.y * rows;
    const int oy1 = min(oy0 + rows, dh);
    const uint8_t* sI = src_img + (long long)frame * sstride;
    const uint8_t* sG = BASE ? nullptr : src_grad + (long long)frame * sstride;
    const bool ld     = L >= 0 && 8 * L < spitch;  // (the pitch is a multiple of 16)
    const bool inner  = lane >= 1 && lane <= PM_INNER;
    // column reflection, per word (k = 0 .. 3: the low word, 4 .. 7: the high word): column c = 8 L + k.  Columns sw, sw + 1 (the
    // only ones beyond the image a tap can reach) are columns sw - 2, sw - 3, which live in the same word or the one to its left;
    // columns -2, -1 (high word of L = -1) are columns 2, 1 (low word of L = 0).  gmask (BASE): the columns where the
    // gradient is defined as 0 (first, last, beyond).
    uint32_t selLo = 0x7654u, selHi = 0x7654u, gmLo = 0xffffffffu, gmHi = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = 8 * L + k, word = 2 * L + (k >> 2), kk = k & 3;
        if (c >= sw && c <= sw + 1) {
            const int sc = 2 * (sw - 1) - c;
            const int nb = (sc >> 2) == word ? 4 + (sc & 3) : (sc & 3);
            if (k < 4)
                selLo = (selLo & ~(0xfu << (4 * kk))) | ((uint32_t)nb << (4 * kk));
            else
                selHi = (selHi & ~(0xfu << (4 * kk))) | ((uint32_t)nb << (4 * kk));
        }
        if (c <= 0 || c >= sw - 1) {
            if (k < 4)
                gmLo &= ~(0xffu << (8 * kk));
            else
                gmHi &= ~(0xffu << (8 * kk));
        }
    }
    const bool fixL = strip == 0;  // lane 0 holds L = -1
    // rows: a segment away from the first / last rows of the image walks a pointer; the others fold the row index
    const int y0        = 2 * oy0 - 2;  // first virtual source row of this segment
    const bool interior = y0 - 1 >= 0 && 2 * oy1 + 6 < sh;  // (rows y0 - 1 .. 2 oy1 + 6 are touched: taps, gradient halo, loads ahead)
    auto run = [&](auto interiorTag) {
    constexpr bool INTERIOR = decltype(interiorTag)::value;
    auto rowOf = [&](int y) {  // BORDER_REFLECT_101 of a row index (|y| small: one fold)
        if (INTERIOR) return y;
        int yy = y < 0 ? -y : y;
        return yy >= sh ? 2 * sh - 2 - yy : yy;
    };
    const long long wofs = ld ? 8ll * L : 0ll;  // (lanes outside the pitched row read the first 8 bytes and discard them)
    auto ldraw = [&](const uint8_t* base, int y) -> uint2 {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(base + (long long)rowOf(y) * spitch + wofs));
        return ld ? v : make_uint2(0u, 0u);
    };
    auto fixcols = [&](uint2 w) -> uint2 {  // (all lanes; the selector of a word that needs no fix is the identity)
        if (EDGE) {
            const uint32_t ph = __shfl_up_sync(FULL, w.y, 1);
            const uint32_t lo = __byte_perm(ph, w.x, selLo);
            w.y               = __byte_perm(w.x, w.y, selHi);  // (its left word is the RAW low word: columns below sw)
            w.x               = lo;
            const uint32_t n0 = __shfl_down_sync(FULL, w.x, 1);
            if (fixL && lane == 0) w.y = __byte_perm(n0, 0u, 0x1200u);
        }
        return w;
    };
    // horizontal sums of the lane's four outputs (two registers of two 16-bit lanes) and the neighbour words they need
    auto hsumOf = [&](uint2 w, uint32_t& h01, uint32_t& h23) {
        const uint32_t ph = __shfl_up_sync(FULL, w.y, 1), nl = __shfl_down_sync(FULL, w.x, 1);
        h01 = pm_hsum(ph, w.x, w.y);
        h23 = pm_hsum(w.x, w.y, nl);
    };
    const int ox        = 4 * L;
    const bool ostore   = inner && ox >= 0 && ox < dpitch;
    const int orem      = dw - ox;
    const uint32_t omsk = orem <= 0 ? 0u : (orem < 4 ? (1u << (8 * orem)) - 1u : 0xffffffffu);
    uint8_t* oI = dst_img + (long long)frame * dstride + (long long)oy0 * dpitch + (ostore ? ox : 0);
    uint8_t* oG = dst_grad + (long long)frame * dstride + (long long)oy0 * dpitch + (ostore ? ox : 0);
    auto emit = [&](uint8_t* dst, const uint32_t (&a)[2], const uint32_t (&b)[2], const uint32_t (&c)[2], const uint32_t (&d)[2], uint32_t e0,
                    uint32_t e1) {
        const uint32_t acc0 = a[0] + 4u * b[0] + 6u * c[0] + 4u * d[0] + e0 + 0x00800080u;  // two 16-bit lanes, <= 65,408 each
        const uint32_t acc1 = a[1] + 4u * b[1] + 6u * c[1] + 4u * d[1] + e1 + 0x00800080u;
        if (ostore) *reinterpret_cast<uint32_t*>(dst) = __byte_perm(acc0, acc1, 0x7531u) & omsk;  // the high byte of each lane
    };
    if (!BASE) {
        uint32_t i1[2], i2[2], i3[2], i4[2], g1[2], g2[2], g3[2], g4[2];
        hsumOf(fixcols(ldraw(sI, y0)), i1[0], i1[1]);
        hsumOf(fixcols(ldraw(sI, y0 + 1)), i2[0], i2[1]);
        hsumOf(fixcols(ldraw(sI, y0 + 2)), i3[0], i3[1]);
        hsumOf(fixcols(ldraw(sI, y0 + 3)), i4[0], i4[1]);
        hsumOf(fixcols(ldraw(sG, y0)), g1[0], g1[1]);
        hsumOf(fixcols(ldraw(sG, y0 + 1)), g2[0], g2[1]);
        hsumOf(fixcols(ldraw(sG, y0 + 2)), g3[0], g3[1]);
        hsumOf(fixcols(ldraw(sG, y0 + 3)), g4[0], g4[1]);
        // (loads run TWO output rows ahead of the arithmetic: with 32 warps per SM that is what it takes to keep HBM busy)
        uint2 ra = ldraw(sI, y0 + 4), rb = ldraw(sI, y0 + 5), rc = ldraw(sG, y0 + 4), rd = ldraw(sG, y0 + 5);
        uint2 pa = ldraw(sI, y0 + 6), pb = ldraw(sI, y0 + 7), pc = ldraw(sG, y0 + 6), pd = ldraw(sG, y0 + 7);
#pragma unroll 1
        for (int oy = oy0; oy < oy1; oy++) {
            const uint2 na = pa, nb = pb, nc = pc, nd = pd;
            pa = ldraw(sI, 2 * oy + 6), pb = ldraw(sI, 2 * oy + 7), pc = ldraw(sG, 2 * oy + 6), pd = ldraw(sG, 2 * oy + 7);
            uint32_t in1[2], gn1[2], in2[2], gn2[2];
            hsumOf(fixcols(ra), in1[0], in1[1]);
            hsumOf(fixcols(rc), gn1[0], gn1[1]);
            emit(oI, i1, i2, i3, i4, in1[0], in1[1]);
            emit(oG, g1, g2, g3, g4, gn1[0], gn1[1]);
            oI += dpitch, oG += dpitch;
            hsumOf(fixcols(rb), in2[0], in2[1]);
            hsumOf(fixcols(rd), gn2[0], gn2[1]);
#pragma unroll
            for (int q = 0; q < 2; q++) {
                i1[q] = i3[q], i2[q] = i4[q], i3[q] = in1[q], i4[q] = in2[q];
                g1[q] = g3[q], g2[q] = g4[q], g3[q] = gn1[q], g4[q] = gn2[q];
            }
            ra = na, rb = nb, rc = nc, rd = nd;
        }
    } else {
        // image rows y - 1, y, y + 1 in registers: up, cur, dn.  Gradient rows [2 oy0, 2 oy1) of the image belong to this
        // segment: its first two and last two virtual rows are the halo of the taps.
        const bool gstore = inner && ld;
        uint8_t* pG0      = grad0 + (long long)frame * sstride + (long long)(2 * oy0) * spitch + wofs;
        uint2 up = fixcols(ldraw(sI, y0 - 1)), cur = fixcols(ldraw(sI, y0));
        uint2 r1 = ldraw(sI, y0 + 1), r2 = ldraw(sI, y0 + 2), r3 = ldraw(sI, y0 + 3), r4 = ldraw(sI, y0 + 4);
        auto step = [&](int y, bool own, uint2 rawDn, uint32_t (&hi)[2], uint32_t (&hg)[2]) {  // virtual row y: its horizontal sums
            const uint2 dn    = fixcols(rawDn);
            const uint32_t ph = __shfl_up_sync(FULL, cur.y, 1), nl = __shfl_down_sync(FULL, cur.x, 1);
            hi[0]             = pm_hsum(ph, cur.x, cur.y);
            hi[1]             = pm_hsum(cur.x, cur.y, nl);
            // columns x - 1 / x + 1 of the two words
            const uint32_t l0 = __funnelshift_r(ph, cur.x, 24), q0 = __funnelshift_r(cur.x, cur.y, 8);
            const uint32_t l1 = __funnelshift_r(cur.x, cur.y, 24), q1 = __funnelshift_r(cur.y, nl, 8);
            uint2 g;
            g.x = __vaddus4(__vabsdiffu4(q0, l0), __vabsdiffu4(dn.x, up.x)) & gmLo;
            g.y = __vaddus4(__vabsdiffu4(q1, l1), __vabsdiffu4(dn.y, up.y)) & gmHi;
            if (!INTERIOR) {
                const int yr = rowOf(y);
                if (yr == 0 || yr == sh - 1) g = make_uint2(0u, 0u);  // SimdLib.h:856-884: first / last row
                own = own && y >= 0 && y < sh;
            }
            if (own) {
                if (gstore) *reinterpret_cast<uint2*>(pG0) = g;
                pG0 += spitch;
            }
            hsumOf(fixcols(g), hg[0], hg[1]);
            up  = cur;
            cur = dn;
        };
        uint32_t i1[2], i2[2], i3[2], i4[2], g1[2], g2[2], g3[2], g4[2];
        step(y0, false, r1, i1, g1);
        step(y0 + 1, false, r2, i2, g2);
        step(y0 + 2, true, r3, i3, g3);
        step(y0 + 3, true, r4, i4, g4);
        uint2 ra = ldraw(sI, y0 + 5), rb = ldraw(sI, y0 + 6), pa = ldraw(sI, y0 + 7), pb = ldraw(sI, y0 + 8);
#pragma unroll 1
        for (int oy = oy0; oy < oy1; oy++) {
            const uint2 na = pa, nb = pb;  // (loads run two output rows ahead)
            pa = ldraw(sI, 2 * oy + 7), pb = ldraw(sI, 2 * oy + 8);  // (rows y0 + 9, y0 + 10 at the first pass)
            const bool own = oy + 1 < oy1;  // (the last output row's two new rows are the halo below the segment)
            uint32_t in1[2], gn1[2], in2[2], gn2[2];
            step(2 * oy + 2, own, ra, in1, gn1);
            emit(oI, i1, i2, i3, i4, in1[0], in1[1]);
            emit(oG, g1, g2, g3, g4, gn1[0], gn1[1]);
            oI += dpitch, oG += dpitch;
            step(2 * oy + 3, own, rb, in2, gn2);
#pragma unroll
            for (int q = 0; q < 2; q++) {
                i1[q] = i3[q], i2[q] = i4[q], i3[q] = in1[q], i4[q] = in2[q];
                g1[q] = g3[q], g2[q] = g4[q], g3[q] = gn1[q], g4[q] = gn2[q];
            }
            ra = na, rb = nb;
        }
    }
    };  // run
    if (interior)
        run(std::true_type{});
    else
        run(std::false_type{});
}

// strips 1 .. edge0 - 1 hold no border column (blockIdx-uniform choice of the instantiation)
template <bool BASE>
__global__ void __launch_bounds__(32) k_pyr_march(const uint8_t* __restrict__ src_img, const uint8_t* __restrict__ src_grad,
                                                  uint8_t* __restrict__ dst_img, uint8_t* __restrict__ dst_grad,
                                                  uint8_t* __restrict__ grad0, int sw, int sh, int spitch, long long sstride,
                                                  int dw, int dh, int dpitch, long long dstride, int edge0, int rows)
{
    if (blockIdx.x == 0 || (int)blockIdx.x >= edge0)
        pyr_march_strip<BASE, true>(src_img, src_grad, dst_img, dst_grad, grad0, sw, sh, spitch, sstride, dw, dh, dpitch, dstride, rows);
    else
        pyr_march_strip<BASE, false>(src_img, src_grad, dst_img, dst_grad, grad0, sw, sh, spitch, sstride, dw, dh, dpitch, dstride, rows);
}

// cv::pyrDown, one level, both stacks, byte-wise: only for levels smaller than 8 x 8, where the reflection can wrap
// more than once.  grid: (ceil(dw/TX), ceil(dh/TY), 2 * n_frames)
constexpr int TX = 64, TY = 8;
constexpr int SW = 2 * TX + 3, SH = 2 * TY + 3;

__global__ void __launch_bounds__(256) k_pyrdown_small(const uint8_t* __restrict__ src_img, uint8_t* __restrict__ dst_img,
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
    bool grad0_done       = false;
    // Batches: the register-marching kernels (level 0 -> 1 with the gradient fused).  One launch pair per level over the whole
    // batch: chunking the batch so that L2 holds what a level hands to the next saves 19 % of the DRAM reads and loses more than
    // that to the extra launches (measured, tests/cuda/pyr_probe.py).  A single frame (the per-frame front end) keeps the fused
    // tile kernel: a handful of CTAs either way, and one launch instead of two.
    const char* me   = getenv("SVO_PYR_MARCH_MIN");  // (measurements) smallest batch that takes the marching kernels
    const int nMarch = me ? atoi(me) : 4;
    if (n >= nMarch && a.levels >= 2 && g0.w >= 64 && g0.h >= 8) {
        const char* ce  = getenv("SVO_PYR_CHUNK");  // (measurements) frames per chunk
        const int chunk = ce ? std::max(1, atoi(ce)) : 1 << 30;
        for (int c0 = 0; c0 < n; c0 += chunk) {
            const int m = std::min(chunk, n - c0), fs = first_slot + c0;
            for (int l = 1; l < a.levels; l++) {
                const LevelGeom& s = a.geom[l - 1];
                const LevelGeom& d = a.geom[l];
                if (s.w >= 64 && s.h >= 8) {
                    // strips 1 .. edge0 - 1 need no column reflection; strip 0 and the strips from edge0 on (the first whose
                    // words reach column sw) do.  Row segments: long (less halo) when the batch fills the device anyway.
                    const int strips = (d.w + 4 * PM_INNER - 1) / (4 * PM_INNER);
                    int edge0        = 1;
                    while (edge0 < strips && 8 * (PM_INNER * edge0 + 31) <= s.w) edge0++;  // (the strip's last column is below sw)
                    // row segments: as long as possible (less halo) while the batch still fills the device
                    int rows = 48;
                    while (rows > 6 && (long long)strips * ((d.h + rows - 1) / rows) * m < 148 * 32) rows /= 2;
                    const int segs = (d.h + rows - 1) / rows;
                    const uint8_t* sg = l == 1 ? nullptr : a.grad[l - 1] + fs * s.plane_stride;
                    uint8_t* g0w      = l == 1 ? a.grad[0] + fs * s.plane_stride : nullptr;
                    const dim3 grid(strips, segs, m);
                    if (l == 1)
                        k_pyr_march<true><<<grid, 32, 0, ctx->pyr_stream>>>(a.img[0] + fs * s.plane_stride, sg, a.img[1] + fs * d.plane_stride,
                                                                           a.grad[1] + fs * d.plane_stride, g0w, s.w, s.h, s.pitch, s.plane_stride,
                                                                           d.w, d.h, d.pitch, d.plane_stride, edge0, rows);
                    else
                        k_pyr_march<false><<<grid, 32, 0, ctx->pyr_stream>>>(a.img[l - 1] + fs * s.plane_stride, sg, a.img[l] + fs * d.plane_stride,
                                                                            a.grad[l] + fs * d.plane_stride, g0w, s.w, s.h, s.pitch, s.plane_stride,
                                                                            d.w, d.h, d.pitch, d.plane_stride, edge0, rows);
                } else if (s.w >= 8 && s.h >= 8) {
                    dim3 grid((d.w + PT_X - 1) / PT_X, (d.h + PT_Y - 1) / PT_Y, m);
                    k_pyr_level<false><<<grid, 256, 0, ctx->pyr_stream>>>(
                        a.img[l - 1] + fs * s.plane_stride, a.grad[l - 1] + fs * s.plane_stride, a.img[l] + fs * d.plane_stride,
                        a.grad[l] + fs * d.plane_stride, nullptr, s.w, s.h, s.pitch, s.plane_stride, d.w, d.h, d.pitch, d.plane_stride);
                } else {
                    dim3 grid((d.w + TX - 1) / TX, (d.h + TY - 1) / TY, 2 * m);
                    k_pyrdown_small<<<grid, 256, 0, ctx->pyr_stream>>>(a.img[l - 1] + fs * s.plane_stride, a.img[l] + fs * d.plane_stride,
                                                                   a.grad[l - 1] + fs * s.plane_stride, a.grad[l] + fs * d.plane_stride,
                                                                   s.w, s.h, s.pitch, s.plane_stride, d.w, d.h, d.pitch, d.plane_stride);
                }
                ctx->launches++;
            }
        }
        SVO_CUDA(cudaGetLastError());
        return SVO_OK;
    }
    for (int l = 1; l < a.levels; l++) {
        const LevelGeom& s = a.geom[l - 1];
        const LevelGeom& d = a.geom[l];
        if (s.w >= 8 && s.h >= 8) {
            dim3 grid((d.w + PT_X - 1) / PT_X, (d.h + PT_Y - 1) / PT_Y, n);
            if (l == 1) {  // fused: gradient level 0 + level 1 of both stacks from one read of the level-0 image
                k_pyr_level<true><<<grid, 256, 0, ctx->pyr_stream>>>(
                    a.img[0] + first_slot * s.plane_stride, nullptr, a.img[1] + first_slot * d.plane_stride,
                    a.grad[1] + first_slot * d.plane_stride, a.grad[0] + first_slot * s.plane_stride, s.w, s.h, s.pitch,
                    s.plane_stride, d.w, d.h, d.pitch, d.plane_stride);
                grad0_done = true;
            } else {
                k_pyr_level<false><<<grid, 256, 0, ctx->pyr_stream>>>(
                    a.img[l - 1] + first_slot * s.plane_stride, a.grad[l - 1] + first_slot * s.plane_stride,
                    a.img[l] + first_slot * d.plane_stride, a.grad[l] + first_slot * d.plane_stride, nullptr, s.w, s.h,
                    s.pitch, s.plane_stride, d.w, d.h, d.pitch, d.plane_stride);
            }
        } else {
            if (l == 1) {
                dim3 g((g0.pitch / 4 + 127) / 128, g0.h, n);
                k_abs_gradient<<<g, 128, 0, ctx->pyr_stream>>>(a.img[0] + first_slot * g0.plane_stride,
                                                             a.grad[0] + first_slot * g0.plane_stride, g0.w, g0.h, g0.pitch,
                                                             g0.plane_stride);
                ctx->launches++;
                grad0_done = true;
            }
            dim3 grid((d.w + TX - 1) / TX, (d.h + TY - 1) / TY, 2 * n);
            k_pyrdown_small<<<grid, 256, 0, ctx->pyr_stream>>>(a.img[l - 1] + first_slot * s.plane_stride,
                                                           a.img[l] + first_slot * d.plane_stride,
                                                           a.grad[l - 1] + first_slot * s.plane_stride,
                                                           a.grad[l] + first_slot * d.plane_stride, s.w, s.h, s.pitch,
                                                           s.plane_stride, d.w, d.h, d.pitch, d.plane_stride);
        }
        ctx->launches++;
    }
    if (!grad0_done) {  // a single-level pyramid still has its gradient image
        dim3 g((g0.pitch / 4 + 127) / 128, g0.h, n);
        k_abs_gradient<<<g, 128, 0, ctx->pyr_stream>>>(a.img[0] + first_slot * g0.plane_stride,
                                                     a.grad[0] + first_slot * g0.plane_stride, g0.w, g0.h, g0.pitch, g0.plane_stride);
        ctx->launches++;
    }
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}

svo_status launch_repack(svo_ctx* ctx, const uint8_t* dsrc, long long src_pitch, long long src_frame_stride, int first_slot, int n)
{
    const LevelGeom& g = ctx->arena.geom[0];
    dim3 grid((g.pitch / 16 + 127) / 128, g.h, n);
    k_repack<<<grid, 128, 0, ctx->pyr_stream>>>(dsrc, src_pitch, src_frame_stride, ctx->arena.img[0] + (int64_t)first_slot * g.plane_stride,
                                            g.w, g.h, g.pitch, g.plane_stride);
    ctx->launches++;
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}
