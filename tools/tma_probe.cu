// Probe: cp.async.bulk.tensor (TMA tiled mode) semantics needed by the patch producers of ss_conv_i8.cu:
//  (1) is the 32B/64B shared-memory swizzle a function of the ABSOLUTE shared-memory address (like the UMMA descriptor's) or
//      relative to the box's destination?  -> load the same box at destination offsets 0 / 128 / 384 and dump shared memory
//  (2) elementStrides = 2 along W (parity-split rows of a stride-2 conv): which pixels arrive, how many
//  (3) negative / out-of-range coordinates: zero fill, and the full box counts towards the mbarrier's transaction bytes
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probes/tma_probe tools/tma_probe.cu   (driver entry point fetched at run time)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int x0, int y0, int n0, uint32_t dst_off, uint32_t bytes, uint8_t* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u;
    uint8_t* b = sm + (base - smem_u32(sm));
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) b[i] = 0xEE;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(base + dst_off), "l"(&tm), "r"(c0), "r"(x0), "r"(y0), "r"(n0), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) out[i] = b[i];
}

int main() {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q) != cudaSuccess || encode == nullptr) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    // tensor u8 [N=2][H=6][W=20][C=32]; byte value = (x * 8 + y) & 0xff in byte 0, channel index in the others' low bits
    const int N = 2, H = 6, W = 20, C = 32;
    std::vector<uint8_t> h((size_t)N * H * W * C);
    for (int n = 0; n < N; ++n) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < C; ++c)
        h[(((size_t)n * H + y) * W + x) * C + c] = c == 0 ? (uint8_t)(100 * n + 20 * y + x + 1) : (uint8_t)(c == 16 ? 0xA0 + x : c);
    uint8_t *d, *out;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 8192);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    std::vector<uint8_t> o(8192);
    auto run = [&](const char* what, int RB, CUtensorMapSwizzle sw, int bx, int by, int sx, int c0, int x0, int y0, int n0, uint32_t dst_off) {
        CUtensorMap tm;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)C, (cuuint64_t)W * C, (cuuint64_t)H * W * C};
        cuuint32_t box[4] = {(cuuint32_t)RB, (cuuint32_t)bx, (cuuint32_t)by, 1};
        cuuint32_t es[4] = {1, (cuuint32_t)sx, 1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", what, (int)r); return; }
        const int npx = (bx + sx - 1) / sx;
        const uint32_t bytes = (uint32_t)(RB * npx * by);
        probe<<<1, 128, 16384>>>(tm, c0, x0, y0, n0, dst_off, bytes, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(o.data(), out, 8192, cudaMemcpyDeviceToHost);
        printf("== %s: box {%d,%d,%d} elemstride %d at (c %d, x %d, y %d, n %d), dst offset %u, expect_tx %u\n", what, RB, bx, by, sx, c0, x0, y0, n0, dst_off, bytes);
        // for each 16-byte chunk of the first rows: which pixel id (byte 0 of a pixel) or 0xA0+x marker (byte 16) it holds
        for (uint32_t row = 0; row < (uint32_t)(npx * by) + 2 && row < 40; ++row) {
            const uint32_t a = dst_off + row * RB;
            printf("  smem+%4u:", a);
            for (int ch = 0; ch < RB / 16; ++ch) printf("  [%3u %3u]", o[a + ch * 16], o[a + ch * 16 + 1]);
            printf("\n");
        }
    };
    run("swizzle32 dst+0", 32, CU_TENSOR_MAP_SWIZZLE_32B, 12, 2, 1, 0, 0, 0, 0, 0);
    run("swizzle32 dst+128", 32, CU_TENSOR_MAP_SWIZZLE_32B, 12, 2, 1, 0, 0, 0, 0, 128);
    run("swizzle32 dst+384", 32, CU_TENSOR_MAP_SWIZZLE_32B, 12, 2, 1, 0, 0, 0, 0, 384);
    run("elemstride 2, x0 = -3 (odd, negative), y0 = -1", 32, CU_TENSOR_MAP_SWIZZLE_32B, 20, 2, 2, 0, -3, -1, 1, 0);
    run("OOB in n", 32, CU_TENSOR_MAP_SWIZZLE_32B, 12, 1, 1, 0, 0, 0, 5, 0);
    return 0;
}
