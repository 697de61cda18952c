// select_bench.cu -- micro-timings of the block_select.cuh building blocks on one CTA (512 threads x 25 keys).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I semi-direct-visual-odometry_b200/csrc -o select_bench select_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "block_select.cuh"

constexpr int AREA = 25;
constexpr int REP  = 50;

__global__ void __launch_bounds__(512, 1) bench(const int* qin, long long* out, uint32_t* res, int variant)
{
    extern __shared__ __align__(16) unsigned char smem[];
    SelCtx sc;
    sc.pp     = 0;
    sc.s.priv = reinterpret_cast<uint32_t*>(smem);
    sc.s.bins = sc.s.priv + 16 * 512;
    sc.s.tot  = sc.s.bins + 1024;
    sc.s.wtot = sc.s.tot + 64;
    const int tid = threadIdx.x;
    for (int i = tid; i < 16 * 512 + 1024 + 64 + 64; i += 512) sc.s.priv[i] = 0;
    uint32_t key[AREA];
    for (int i = 0; i < AREA; i++) key[i] = (uint32_t)(qin[i * 512 + tid] + (1 << 25));
    __syncthreads();
    const int k = 512 * AREA / 2;
    uint32_t acc = 0;
    // warm
    Bracket br{0u, 4, false};
    uint32_t pred;
    int tier;
    uint32_t truth = tiered_select<AREA>(key, true, 512 - 32, k, true, br, sc, &pred, &tier);
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < REP; r++) {
        if (variant == 0) {  // full hot select (bracket valid, centred on truth, shift 4)
            Bracket b2{truth + 100u * r, 4, true};
            acc += tiered_select<AREA>(key, true, 512 - 32, k, true, b2, sc, &pred, &tier) + tier;
        } else if (variant == 1) {  // cold select
            Bracket b2{0u, 4, false};
            acc += tiered_select<AREA>(key, true, 512 - 32, k, true, b2, sc, &pred, &tier) + tier;
        } else if (variant == 9) {  // warm cold select: adaptive window around a stale result 1.5 units away
            Bracket b2{truth + 98304u, 4, false};
            acc += tiered_select<AREA>(key, true, 512 - 32, k, true, b2, sc, &pred, &tier) + tier;
        } else if (variant == 2) {  // sweep A only
            uint32_t bin, rank, pb;
            acc += bracket_sweep_a<AREA>(key, true, truth - 4096u + r, 4, k, sc, &bin, &rank, &pb) + bin;
        } else if (variant == 3) {  // sweep B only
            uint32_t ko, ro, po;
            bool hp;
            bracket_sweep_b<AREA>(key, true, truth - 8u, 8u, sc, &ko, &ro, &po, &hp);
            acc += ko;
        } else if (variant == 4) {  // private coarse pass
            uint32_t below;
            acc += private_pass<AREA>(key, true, DigitCoarse{512 - 32}, (uint32_t)k, sc, &below) + below;
        } else if (variant == 5) {  // private prefix pass (few match)
            uint32_t below;
            acc += private_pass<AREA>(key, true, DigitRange{truth & 0xffff0000u, 10}, 10u, sc, &below) + below;
        } else if (variant == 6) {  // element loop of sweep A without atomics / scan: pure ALU
            uint32_t below = 0, inr = 0;
            const uint32_t lo = truth - 4096u + r;
#pragma unroll
            for (int i = 0; i < AREA; i++) {
                const uint32_t t = key[i] - lo;
                below += t >> 31;
                inr += ((t >> 4) < 512u) ? 1u : 0u;
            }
            acc += below + inr;
        } else if (variant == 10 || variant == 11) {  // FMA-heavy producer loop (12 FFMA per key), 11: + fused sweep-A ops
            float f0 = __uint_as_float(key[0]) * 1e-30f + r, g = 1.0001f;
            uint32_t below = 0, cnt = 0, k1 = 0, k2 = 0;
            const uint32_t lo = truth - 4096u + r, width = 8192u;
#pragma unroll
            for (int i = 0; i < AREA; i++) {
                float x = __uint_as_float(key[i] | 0x3f000000u);
#pragma unroll
                for (int j = 0; j < 12; j++) x = fmaf(x, g, f0);
                const uint32_t kk = (uint32_t)(__float2int_rn(x * 1e-3f)) + key[i];
                if (variant == 11) {
                    const uint32_t t = kk - lo;
                    below += t >> 31;
                    const bool in = t < width;
                    k2  = (in && cnt == 1) ? kk : k2;
                    k1  = (in && cnt == 0) ? kk : k1;
                    cnt += in ? 1u : 0u;
                } else {
                    acc += kk;
                }
            }
            acc += below + cnt + k1 + k2;
        } else if (variant == 7) {  // two barriers only
            __syncthreads();
            acc += sc.s.wtot[tid & 15];
            __syncthreads();
        } else if (variant == 8) {  // locate_in_bins only (one barrier inside)
            uint32_t total, bin = 0, rank = 0, pb;
            bool found;
            locate_in_bins(sc.s.bins, sc.s.bins + 512, sc.s, 0u, -1, false, &total, &bin, &rank, &pb, &found);
            acc += bin + total;
        }
        key[r % AREA == 0 ? 0 : 1] ^= (acc & 0);  // keep the loop from being hoisted
    }
    long long t1 = clock64();
    if (tid == 0) {
        out[variant] = (t1 - t0) / REP;
        res[0]       = truth;
        res[1]       = acc;
    }
}

int main()
{
    std::vector<int> q(512 * AREA);
    srand(7);
    std::vector<int> sorted;
    for (auto& v : q) {
        // roughly Gaussian residuals, sigma ~ 4 intensity units, fixed point 2^16
        double s = 0;
        for (int j = 0; j < 12; j++) s += rand() / (double)RAND_MAX;
        v = (int)((s - 6.0) * 4.0 * 65536.0);
        sorted.push_back(v);
    }
    std::sort(sorted.begin(), sorted.end());
    int* dq;
    long long* dout;
    uint32_t* dres;
    cudaMalloc(&dq, q.size() * 4);
    cudaMalloc(&dout, 16 * 8);
    cudaMalloc(&dres, 16);
    cudaMemcpy(dq, q.data(), q.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (16 * 512 + 1024 + 64 + 64) * 4;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char* names[] = {"hot select (A+B)", "cold select (2 private + A + B)", "sweep A", "sweep B", "private coarse pass",
                           "private prefix pass", "sweep A element loop only (ALU)", "two barriers", "locate_in_bins", "warm-cold select (window + A + B)", "producer loop 12 FFMA/key", "producer loop + fused sweep-A ops"};
    for (int v = 0; v < 12; v++) {
        bench<<<1, 512, smem>>>(dq, dout, dres, v);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("variant %d: %s\n", v, cudaGetErrorString(e));
            return 1;
        }
        long long c;
        uint32_t r[2];
        cudaMemcpy(&c, dout + v, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(r, dres, 8, cudaMemcpyDeviceToHost);
        printf("%-36s %7lld cycles   (median key %u, expected %u)\n", names[v], c, r[0], (uint32_t)(sorted[512 * AREA / 2] + (1 << 25)));
    }
    return 0;
}
