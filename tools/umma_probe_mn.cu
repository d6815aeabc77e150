// Probe of tcgen05.mma kind::f16 (bf16 x bf16 -> f32) with MN-MAJOR operands on sm_100a (bring-up tool, not product code).
// The weight-gradient kernel contracts over PIXELS, and pixels are the slow dimension of NHWC activations, so both
// operands are "MN-major": a shared-memory row = one pixel (K index), its bytes = channels (M or N index).  Checks
//   (1) the MN-major descriptor fields (LBO = stride between channel blocks, SBO = stride between 8-pixel groups),
//   (2) that the swizzle is applied on absolute address bits for MN-major as well (start address shifted by whole rows),
//   (3) the issue rate of 128 x N x 16 and 64 x N x 16 MMAs for N = 16 / 32 / 64.
// nvcc -gencode arch=compute_100a,code=sm_100a -o build/umma_probe_mn tools/umma_probe_mn.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline uint32_t layout_code(int rb) { return rb == 128 ? 2u : (rb == 64 ? 4u : 6u); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__host__ __device__ inline uint32_t swz(uint32_t off, int rb) {
    const uint32_t mask = rb == 128 ? 7u : (rb == 64 ? 3u : 1u);
    return off ^ (((off >> 7) & mask) << 4);
}
// kind::f16, D = f32, A = B = bf16, both MN-major
__host__ __device__ constexpr uint32_t idesc_bf16_mn(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
        "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (clock64() - t0 > 2000000000LL) __trap();
    }
}

struct Cfg {
    int m;             // 128 or 64
    int n;             // 16 .. 64
    int a_rb, a_lbo, a_sbo, a_shift;   // A patch: row bytes (= swizzle), bytes between channel blocks, bytes between 8-row groups, start row
    int b_rb, b_lbo, b_sbo, b_shift;
    int nk, a_kadv, b_kadv;            // number of K=16 MMAs and the start-address advance (bytes) between them
};

constexpr int A_REGION = 65536, B_REGION = 32768;

// a_g / b_g: logical (unswizzled) images of the two regions; the kernel copies them into shared memory with the swizzle
__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* a_g, const uint8_t* b_g, Cfg c, float* d_out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < A_REGION / 16; i += 128)
        *reinterpret_cast<uint4*>(sm + swz((uint32_t)i * 16, c.a_rb)) = *reinterpret_cast<const uint4*>(a_g + (size_t)i * 16);
    for (int i = tid; i < B_REGION / 16; i += 128)
        *reinterpret_cast<uint4*>(sm + A_REGION + swz((uint32_t)i * 16, c.b_rb)) = *reinterpret_cast<const uint4*>(b_g + (size_t)i * 16);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint64_t ad = make_desc(base + (uint32_t)c.a_shift * c.a_rb, (uint32_t)c.a_lbo, (uint32_t)c.a_sbo, layout_code(c.a_rb));
        const uint64_t bd = make_desc(base + A_REGION + (uint32_t)c.b_shift * c.b_rb, (uint32_t)c.b_lbo, (uint32_t)c.b_sbo, layout_code(c.b_rb));
        const uint32_t idesc = idesc_bf16_mn(c.m, c.n);
        for (int k = 0; k < c.nk; ++k) mma_f16(tmem, ad + (uint64_t)((k * c.a_kadv) >> 4), bd + (uint64_t)((k * c.b_kadv) >> 4), idesc, k > 0);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    wait_bar(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = tid >> 5, lane = tid & 31;
    for (int j0 = 0; j0 < c.n; j0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)j0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) d_out[(warp * 32 + lane) * 64 + j0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

__global__ void __launch_bounds__(128) rate_kernel(Cfg c, int iters, long long* cycles_out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < (A_REGION + B_REGION) / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0x3c003c00u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint64_t ad = make_desc(base, (uint32_t)c.a_lbo, (uint32_t)c.a_sbo, layout_code(c.a_rb));
        const uint64_t bd = make_desc(base + A_REGION, (uint32_t)c.b_lbo, (uint32_t)c.b_sbo, layout_code(c.b_rb));
        const uint32_t idesc = idesc_bf16_mn(c.m, c.n);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            // 25 taps: same A, B start shifted by one row per tap, a different accumulator per tap
#pragma unroll 5
            for (int tap = 0; tap < 25; ++tap)
                mma_f16(tmem + (uint32_t)((tap * c.n) % 448), ad, bd + (uint64_t)((tap * c.b_rb) >> 4), idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        wait_bar(smem_u32(&bar), 0);
        cycles_out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static float bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

int main() {
    std::vector<uint16_t> a(A_REGION / 2), b(B_REGION / 2);
    srand(7);
    // small integers: every product and partial sum is exact in fp32, so the comparison is exact
    auto rnd = []() {
        const int v = rand() % 9 - 4;
        __nv_bfloat16 h = __float2bfloat16((float)v);
        uint16_t u;
        memcpy(&u, &h, 2);
        return u;
    };
    for (auto& v : a) v = rnd();
    for (auto& v : b) v = rnd();
    uint8_t *da, *db;
    float* dd;
    cudaMalloc(&da, A_REGION);
    cudaMalloc(&db, B_REGION);
    cudaMalloc(&dd, 128 * 64 * 4);
    cudaMemcpy(da, a.data(), A_REGION, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), B_REGION, cudaMemcpyHostToDevice);
    const int smem = 1024 + A_REGION + B_REGION;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);

    struct Named { const char* what; Cfg c; };
    const Named tests[] = {
        // m   n   a_rb a_lbo  a_sbo a_sh  b_rb b_lbo b_sbo b_sh nk a_kadv b_kadv
        {"canonical A rb128 x2 blocks, B rb32 N16", {128, 16, 128, 16384, 1024, 0, 32, 8192, 256, 0, 1, 0, 0}},
        {"same, 2 K-steps (advance 16 rows)", {128, 16, 128, 16384, 1024, 0, 32, 8192, 256, 0, 2, 2048, 512}},
        {"A start shifted 3 rows", {128, 16, 128, 16384, 1024, 3, 32, 8192, 256, 0, 1, 0, 0}},
        {"B start shifted 5 rows", {128, 16, 128, 16384, 1024, 0, 32, 8192, 256, 5, 1, 0, 0}},
        {"A,B shifted 13 / 27 rows", {128, 16, 128, 16384, 1024, 13, 32, 8192, 256, 27, 1, 0, 0}},
        {"patch rows: A sbo 12 rows, B sbo 12 rows, B shift 14", {128, 16, 128, 16384, 12 * 128, 0, 32, 8192, 12 * 32, 14, 1, 0, 0}},
        {"patch rows, 2 K-steps advancing 2 patch rows", {128, 16, 128, 16384, 12 * 128, 5, 32, 8192, 12 * 32, 31, 2, 24 * 128, 24 * 32}},
        {"B rb64 N32 shifted 9", {128, 32, 128, 16384, 1024, 0, 64, 8192, 12 * 64, 9, 1, 0, 0}},
        {"B rb128 N64 shifted 9", {128, 64, 128, 16384, 1024, 0, 128, 8192, 12 * 128, 9, 1, 0, 0}},
        {"A rb64 x4 blocks (lbo 8192), B rb32", {128, 16, 64, 8192, 512, 2, 32, 8192, 256, 3, 1, 0, 0}},
        {"M64: A rb128 one block shifted 7, B rb32 N16", {64, 16, 128, 16384, 12 * 128, 7, 32, 8192, 12 * 32, 20, 1, 0, 0}},
        {"M64: A rb64 x2 blocks, B rb64 N32", {64, 32, 64, 8192, 12 * 64, 7, 64, 8192, 12 * 64, 20, 1, 0, 0}},
        {"B rb32 N32 via 2 blocks (lbo 8192)", {128, 32, 128, 16384, 1024, 0, 32, 8192, 256, 5, 1, 0, 0}},
    };
    int all_ok = 1;
    for (const Named& t : tests) {
        const Cfg& c = t.c;
        cudaMemset(dd, 0, 128 * 64 * 4);
        probe_kernel<<<1, 128, smem>>>(da, db, c, dd);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("%-60s CUDA error %s\n", t.what, cudaGetErrorString(e));
            return 1;
        }
        std::vector<float> d(128 * 64);
        cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
        const int a_cpb = c.a_rb / 2, b_cpb = c.b_rb / 2;   // channels per block
        int bad = 0, first_m = -1, first_n = -1;
        float got0 = 0, exp0 = 0;
        for (int m = 0; m < c.m; ++m)
            for (int n = 0; n < c.n; ++n) {
                float acc = 0.f;
                for (int kk = 0; kk < c.nk; ++kk)
                    for (int k = 0; k < 16; ++k) {
                        const size_t ao = (size_t)(m / a_cpb) * c.a_lbo + (size_t)kk * c.a_kadv + (size_t)(k / 8) * c.a_sbo +
                                          (size_t)(c.a_shift + k % 8) * c.a_rb + (size_t)(m % a_cpb) * 2;
                        const size_t bo = (size_t)(n / b_cpb) * c.b_lbo + (size_t)kk * c.b_kadv + (size_t)(k / 8) * c.b_sbo +
                                          (size_t)(c.b_shift + k % 8) * c.b_rb + (size_t)(n % b_cpb) * 2;
                        acc += bf2f(a[ao / 2]) * bf2f(b[bo / 2]);
                    }
                const float got = d[m * 64 + n];
                if (got != acc) {
                    if (!bad) { first_m = m; first_n = n; got0 = got; exp0 = acc; }
                    ++bad;
                }
            }
        printf("%-60s %s", t.what, bad ? "MISMATCH" : "ok");
        if (bad) printf("  (%d of %d wrong, first at m=%d n=%d: got %g expected %g)", bad, c.m * c.n, first_m, first_n, got0, exp0);
        printf("\n");
        if (bad) all_ok = 0;
    }
    printf("descriptor semantics: %s\n", all_ok ? "ALL OK" : "SOME FAILED");

    long long* dc;
    cudaMalloc(&dc, 148 * sizeof(long long));
    const Named rates[] = {
        {"M128 N16 (A rb128 x2, B rb32)", {128, 16, 128, 16384, 12 * 128, 0, 32, 8192, 12 * 32, 0, 1, 0, 0}},
        {"M128 N32 (B rb64)", {128, 32, 128, 16384, 12 * 128, 0, 64, 8192, 12 * 64, 0, 1, 0, 0}},
        {"M128 N64 (B rb128)", {128, 64, 128, 16384, 12 * 128, 0, 128, 8192, 12 * 128, 0, 1, 0, 0}},
        {"M64  N16 (A rb128, B rb32)", {64, 16, 128, 16384, 12 * 128, 0, 32, 8192, 12 * 32, 0, 1, 0, 0}},
        {"M64  N32 (A rb128, B rb64)", {64, 32, 128, 16384, 12 * 128, 0, 64, 8192, 12 * 64, 0, 1, 0, 0}},
        {"M64  N64 (A rb128, B rb128)", {64, 64, 128, 16384, 12 * 128, 0, 128, 8192, 12 * 128, 0, 1, 0, 0}},
    };
    for (const Named& t : rates) {
        const int iters = 400;
        rate_kernel<<<148, 128, smem>>>(t.c, iters, dc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("rate %s: CUDA error %s\n", t.what, cudaGetErrorString(e));
            return 1;
        }
        std::vector<long long> cc(148);
        cudaMemcpy(cc.data(), dc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (long long v : cc) mx = v > mx ? v : mx;
        printf("rate bf16 MN-major %-32s %.1f cycles per MMA (K = 16)\n", t.what, (double)mx / (iters * 25.0));
    }
    return 0;
}
