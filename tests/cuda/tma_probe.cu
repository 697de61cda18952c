// tma_probe.cu -- standalone probe of the TMA behaviour k_sparse_align relies on (run on a B200):
//   * per-feature boxes of 16 B x 8 rows at arbitrary (unaligned) x, issued by many threads onto one mbarrier
//   * the shared-memory layout such a box gets under CU_TENSOR_MAP_SWIZZLE_128B vs SWIZZLE_NONE
//   * zero fill of out-of-bounds coordinates
//   * issue -> completion latency for 512 boxes
// build: nvcc -gencode arch=compute_100a,code=sm_100a --cudart=shared -lcuda -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess) {                                                            \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            std::exit(1);                                                                  \
        }                                                                                  \
    } while (0)

constexpr int W = 150, H = 64, PITCH = 160, SLOTS = 4, NT = 512;

__host__ __device__ inline uint8_t pat(int x, int y, int s) { return (uint8_t)((x * 31 + y * 17 + s * 101 + (x * y) % 7) & 0xff); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(NT) probe(const __grid_constant__ CUtensorMap tmap, const int* coords, uint8_t* out,
                                            long long* cycles, int variant)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    // (the runtime only guarantees 16 bytes for the dynamic part: align by hand, the launch reserves 1 KB of slack)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(NT));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int cx = coords[3 * tid], cy = coords[3 * tid + 1], cs = coords[3 * tid + 2];
    const long long t0 = clock64();
    const bool issue = variant >= 2 || (variant == 1 && tid == 0);
    if (issue) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(128) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                smem_u32(smem + tid * 128)),
            "l"(&tmap), "r"(cx), "r"(cy), "r"(cs), "r"(smem_u32(&bar))
            : "memory");
    } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(&bar)), "r"(0)
            : "memory");
    }
    const long long t1 = clock64();
    for (int i = 0; i < 128; i++) out[tid * 128 + i] = smem[tid * 128 + i];
    if (tid == 0) cycles[0] = t1 - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv)
{
    const int variant = argc > 1 ? std::atoi(argv[1]) : 4;
    std::printf("variant %d\n", variant);
    std::vector<uint8_t> h((size_t)SLOTS * H * PITCH);
    for (int s = 0; s < SLOTS; s++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < PITCH; x++) h[((size_t)s * H + y) * PITCH + x] = pat(x, y, s);
    uint8_t* d;
    CK(cudaMalloc(&d, h.size()));
    CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    std::vector<int> coords(3 * NT);
    for (int t = 0; t < NT; t++) {
        coords[3 * t]     = (t * 7) % (W - 10) - 5;  // some negative, all unaligned
        coords[3 * t + 1] = (t * 5) % (H + 4) - 4;   // some rows out of bounds (top and bottom)
        if (variant <= 3) {  // in bounds
            coords[3 * t]     = variant == 3 ? (t * 7) % (W - 20) : ((t * 16) % 128);
            coords[3 * t + 1] = (t * 5) % (H - 8);
        }
        coords[3 * t + 2] = t % SLOTS;
    }
    int* dc;
    CK(cudaMalloc(&dc, coords.size() * 4));
    CK(cudaMemcpy(dc, coords.data(), coords.size() * 4, cudaMemcpyHostToDevice));
    uint8_t* dout;
    CK(cudaMalloc(&dout, NT * 128));
    long long* dcyc;
    CK(cudaMalloc(&dcyc, 8));

    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode || qres != cudaDriverEntryPointSuccess) {
        std::printf("cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, NT * 128 + 1024));
    int rc = 0;
    for (int mode = 0; mode < 2; mode++) {
        CUtensorMap tmap;
        // logical width W (< PITCH): columns >= W must read as zero even though memory holds the pattern
        cuuint64_t dims[3]    = {W, H, SLOTS};
        cuuint64_t strides[2] = {PITCH, (cuuint64_t)PITCH * H};
        cuuint32_t box[3]     = {16, 8, 1};
        cuuint32_t estr[3]    = {1, 1, 1};
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            mode ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            std::printf("mode %d: encode failed %d\n", mode, (int)r);
            rc = 1;
            continue;
        }
        long long cyc = 0;
        for (int rep = 0; rep < 3; rep++) {
            probe<<<1, NT, NT * 128 + 1024>>>(tmap, dc, dout, dcyc, variant);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
            std::printf("mode %s rep %d: 512 boxes issue->complete %lld cycles\n", mode ? "SWIZZLE_128B" : "SWIZZLE_NONE", rep, cyc);
        }
        std::vector<uint8_t> out(NT * 128);
        CK(cudaMemcpy(out.data(), dout, out.size(), cudaMemcpyDeviceToHost));
        long bad_dense = 0, bad_xor = 0;
        for (int t = 0; t < (variant >= 2 ? NT : variant); t++)
            for (int r8 = 0; r8 < 8; r8++)
                for (int j = 0; j < 16; j++) {
                    const int x = coords[3 * t] + j, y = coords[3 * t + 1] + r8, s = coords[3 * t + 2];
                    const uint8_t want = (x < 0 || x >= W || y < 0 || y >= H) ? 0 : pat(x, y, s);
                    bad_dense += out[t * 128 + r8 * 16 + j] != want;
                    bad_xor += out[t * 128 + ((r8 ^ (t & 7)) * 16) + j] != want;
                }
        std::printf("mode %s: mismatches dense-layout %ld, xor(row, tile%%8)-layout %ld\n", mode ? "SWIZZLE_128B" : "SWIZZLE_NONE",
                    bad_dense, bad_xor);
        if (mode == 0 && bad_dense) rc = 1;
        if (mode == 1 && bad_xor) rc = 1;
    }
    std::printf(rc ? "PROBE FAILED\n" : "PROBE OK\n");
    return rc;
}
