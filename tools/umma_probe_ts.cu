// Probe: tcgen05.mma kind::f16 with the A operand in TENSOR MEMORY (written with tcgen05.st) and B MN-major in shared memory.
// Checks the TMEM layout of A (lane = M row, 32-bit column j = K elements 2j, 2j+1) and the issue rate against the
// shared-memory-A form.  nvcc -gencode arch=compute_100a,code=sm_100a -o build/umma_probe_ts tools/umma_probe_ts.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__host__ __device__ inline uint32_t swz(uint32_t off, uint32_t mask) { return off ^ (((off >> 7) & mask) << 4); }
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (clock64() - t0 > 2000000000LL) __trap();
    }
}
// D f32, A bf16 (TMEM), B bf16 MN-major (b_major bit 16), M = 128
__host__ __device__ constexpr uint32_t idesc_ts(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
        "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc)
        : "memory");
}

constexpr int N = 32;       // B: [16 pixels][32 channels] bf16, 64-byte rows, SWIZZLE_64B

// mode 0: correctness (a_g: [128][16] bf16 row-major = A, b_g: logical [16][32]); mode 1: rate
__global__ void __launch_bounds__(128) probe(const uint16_t* a_g, const uint16_t* b_g, float* d_out, int iters, long long* cyc) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // B: row k (pixel) = 64 bytes; 16 rows, swizzled on absolute address
    for (int i = tid; i < 16 * 4; i += 128) {
        const int k = i / 4, ch = i % 4;
        *reinterpret_cast<uint4*>(sm + swz((uint32_t)(k * 64 + ch * 16), 3u)) = *reinterpret_cast<const uint4*>(b_g + k * 32 + ch * 8);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    // A -> TMEM columns 256..263: thread m (lane m of TMEM) writes its 16 K elements as 8 packed 32-bit columns
    {
        uint32_t r[8];
        for (int j = 0; j < 8; ++j) r[j] = (uint32_t)a_g[tid * 16 + 2 * j] | ((uint32_t)a_g[tid * 16 + 2 * j + 1] << 16);
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 256u;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                     "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                     : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint64_t bd = make_desc(base, 16u, 512u, 4u);     // SW64, 8-row groups 512 B apart
        const long long t0 = clock64();
        if (iters == 0) {
            mma_ts(tmem, tmem + 256u, bd, idesc_ts(N), 0u);
        } else {
            for (int it = 0; it < iters; ++it)
                for (int tap = 0; tap < 15; ++tap) mma_ts(tmem + (uint32_t)(tap * N) % 224u, tmem + 256u, bd, idesc_ts(N), 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        wait_bar(smem_u32(&bar), 0);
        if (cyc) cyc[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    wait_bar(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (iters == 0) {
        for (int j0 = 0; j0 < N; j0 += 16) {
            uint32_t r[16];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)j0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; ++j) d_out[(warp * 32 + lane) * N + j0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static float bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

int main() {
    std::vector<uint16_t> a(128 * 16), b(16 * 32);
    srand(3);
    auto rnd = []() {
        __nv_bfloat16 h = __float2bfloat16((float)(rand() % 9 - 4));
        uint16_t u;
        memcpy(&u, &h, 2);
        return u;
    };
    for (auto& v : a) v = rnd();
    for (auto& v : b) v = rnd();
    uint16_t *da, *db;
    float* dd;
    long long* dc;
    cudaMalloc(&da, a.size() * 2);
    cudaMalloc(&db, b.size() * 2);
    cudaMalloc(&dd, 128 * N * 4);
    cudaMalloc(&dc, 148 * 8);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, 128 * N * 4);
    probe<<<1, 128, 4096>>>(da, db, dd, 0, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("A-in-TMEM MMA: CUDA error %s\n", cudaGetErrorString(e));
        return 1;
    }
    std::vector<float> d(128 * N);
    cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            float acc = 0.f;
            for (int k = 0; k < 16; ++k) acc += bf2f(a[m * 16 + k]) * bf2f(b[k * 32 + n]);
            if (d[m * N + n] != acc) {
                if (bad < 4) printf("  mismatch m=%d n=%d got %g expected %g\n", m, n, d[m * N + n], acc);
                ++bad;
            }
        }
    printf("A in TMEM (tcgen05.st, lane = row, column j = K 2j,2j+1), B MN-major smem: %s (%d wrong of %d)\n", bad ? "MISMATCH" : "ok", bad,
           128 * N);
    probe<<<148, 128, 4096>>>(da, db, dd, 400, dc);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("rate: CUDA error %s\n", cudaGetErrorString(e));
        return 1;
    }
    std::vector<long long> cc(148);
    cudaMemcpy(cc.data(), dc, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (long long v : cc) mx = v > mx ? v : mx;
    printf("rate: %.1f cycles per 128x32x16 MMA with A in TMEM (shared-memory A, MN-major: 48-71)\n", (double)mx / (400 * 15.0));
    return 0;
}
