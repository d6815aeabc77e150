// Probe of tcgen05 shared-memory descriptor semantics on sm_100a (bring-up tool, not product code).
//
// Question: can the A operand of an implicit-GEMM conv be read straight out of a shared-memory halo patch
// ([patch pixel][64 channels], 128-byte rows, SWIZZLE_128B written by ABSOLUTE address bits) by shifting the
// descriptor start by (ky*PW + kx) rows, with stride-byte-offset = PW*128 between 8-row groups?
// For every (PW, shift, base_offset mode) the probe runs D[128x32] = A[128x64] * B[32x64]^T and compares with
// the host result.   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe tools/umma_probe.cu && ./umma_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

constexpr int N = 32;
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// patch: rows of 128 bytes (64 bf16).  A row r (0..127) = patch row (r/8)*PW + (r%8) + shift.
__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* patch_g, int patch_rows, const __nv_bfloat16* b_g,
                                                     int PW, int shift, int bo_mode, float* d_out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    uint8_t* sB = sm;                 // 32 rows x 128 B = 4 KB, canonical
    uint8_t* sP = sm + 4096;          // patch
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    // write with absolute-address swizzle: 16-byte chunk index ^= (byte address >> 7) & 7
    for (int i = tid; i < patch_rows * 8; i += 128) {
        const int row = i >> 3, ch = i & 7;
        const uint32_t rowaddr = (uint32_t)(sP - sm) + row * 128;
        const uint32_t dst = rowaddr + ((ch ^ ((rowaddr >> 7) & 7)) << 4);
        *reinterpret_cast<uint4*>(sm + dst) = *reinterpret_cast<const uint4*>(patch_g + row * 64 + ch * 8);
    }
    for (int i = tid; i < N * 8; i += 128) {
        const int row = i >> 3, ch = i & 7;
        *reinterpret_cast<uint4*>(sB + row * 128 + ((ch ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(b_g + row * 64 + ch * 8);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t a_addr = base + 4096 + (uint32_t)shift * 128;
        const uint32_t bo = (bo_mode == 1) ? ((a_addr >> 7) & 7) : 0;
        const uint64_t ad = make_desc(a_addr, (uint32_t)PW * 128, bo);
        const uint64_t bd = make_desc(base, 1024, 0);
        for (int k = 0; k < 4; ++k) {
            const uint32_t acc = k > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(IDESC), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait
    {
        uint32_t ok = 0;
        const long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            if (clock64() - t0 > 2000000000LL) __trap();
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = tid >> 5, lane = tid & 31;
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) d_out[(warp * 32 + lane) * N + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}


// ------------------------------------------------------------------------------------------------------------
// MMA issue-rate probe: one CTA per SM issues `iters` x 4 MMAs (128 x N x 16, SS operands, no loads) and reports
// cycles per MMA.  Tells whether small-N MMAs are bound by the shared-memory operand reads.
__global__ void __launch_bounds__(128) rate_kernel(int n, int iters, long long* cycles_out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t ad = make_desc(base, 1024, 0);
        const uint64_t bd = make_desc(base + 16384, 1024, 0);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                    "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(1u)
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            if (clock64() - t0 > 4000000000LL) __trap();
        }
        cycles_out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

static void run_rate() {
    long long* dc;
    cudaMalloc(&dc, 148 * sizeof(long long));
    const int smem = 1024 + 16384 + 32768;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int Ns[] = {16, 32, 64, 96, 128, 192, 256};
    for (int grid : {1, 148})
        for (int n : Ns) {
            const int iters = 2000;
            rate_kernel<<<grid, 128, smem>>>(n, iters, dc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("rate N=%d: CUDA error %s\n", n, cudaGetErrorString(e));
                return;
            }
            std::vector<long long> c(grid);
            cudaMemcpy(c.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (long long v : c) mx = v > mx ? v : mx;
            const double per = (double)mx / (iters * 4.0);
            printf("rate grid=%3d N=%3d: %.1f cycles per 128xNx16 MMA  (ideal %.1f; %.0f%% of tensor peak)\n", grid, n, per,
                   128.0 * n / 256.0, 100.0 * (128.0 * n / 256.0) / per);
        }
}

int main() {
    run_rate();
    const int MAXROWS = 16 * 24 + 16;
    std::vector<__nv_bfloat16> patch(MAXROWS * 64), b(N * 64);
    std::vector<float> pf(MAXROWS * 64), bf(N * 64);
    srand(1);
    for (size_t i = 0; i < patch.size(); ++i) {
        pf[i] = (float)(rand() % 7 - 3);
        patch[i] = __float2bfloat16(pf[i]);
    }
    for (size_t i = 0; i < b.size(); ++i) {
        bf[i] = (float)(rand() % 5 - 2);
        b[i] = __float2bfloat16(bf[i]);
    }
    __nv_bfloat16 *dp, *db;
    float* dd;
    cudaMalloc(&dp, patch.size() * 2);
    cudaMalloc(&db, b.size() * 2);
    cudaMalloc(&dd, 128 * N * 4);
    cudaMemcpy(dp, patch.data(), patch.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    const int smem = 1024 + 4096 + MAXROWS * 128;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> got(128 * N);
    const int PWs[] = {8, 16, 24, 12, 20};
    for (int PW : PWs)
        for (int shift = 0; shift < 10; ++shift)
            for (int bo = 0; bo < 2; ++bo) {
                cudaMemset(dd, 0, 128 * N * 4);
                probe_kernel<<<1, 128, smem>>>(dp, MAXROWS, db, PW, shift, bo, dd);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("PW=%d shift=%d bo=%d CUDA error %s\n", PW, shift, bo, cudaGetErrorString(e));
                    return 1;
                }
                cudaMemcpy(got.data(), dd, 128 * N * 4, cudaMemcpyDeviceToHost);
                int bad = 0;
                for (int r = 0; r < 128; ++r) {
                    const int prow = (r / 8) * PW + (r % 8) + shift;
                    for (int n = 0; n < N; ++n) {
                        float ref = 0;
                        for (int k = 0; k < 64; ++k) ref += pf[prow * 64 + k] * bf[n * 64 + k];
                        if (ref != got[r * N + n]) ++bad;
                    }
                }
                printf("PW=%2d shift=%d base_offset_mode=%d : %s (%d wrong)\n", PW, shift, bo, bad ? "MISMATCH" : "ok", bad);
            }
    return 0;
}
