// Probe of tcgen05.mma kind::i8 on sm_100a (bring-up tool, not product code):
//   (1) correctness of u8 x s8 -> s32 with the A operand read out of a shared-memory halo patch whose rows are
//       32 / 64 / 128 bytes (SWIZZLE_32B / 64B / 128B, swizzle applied on absolute address bits), shifted start,
//       arbitrary 8-row-group pitch;
//   (2) issue rate of 128 x N x 32 i8 MMAs versus N.
// nvcc -gencode arch=compute_100a,code=sm_100a -o build/umma_probe_i8 tools/umma_probe_i8.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__host__ __device__ inline uint32_t layout_code(int rb) { return rb == 128 ? 2u : (rb == 64 ? 4u : 6u); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
// absolute-address swizzle of a byte offset (offset relative to a 1024-aligned base)
__host__ __device__ inline uint32_t swz(uint32_t off, int rb) {
    const uint32_t mask = rb == 128 ? 7u : (rb == 64 ? 3u : 1u);
    return off ^ (((off >> 7) & mask) << 4);
}
__host__ __device__ constexpr uint32_t idesc_i8(int n, int a_signed, int b_signed) {
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
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

constexpr int N = 32;

// patch rows of rb bytes; A row r = patch row (r/8)*PW + (r%8) + shift; B canonical [N][rb].
__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* patch_g, int patch_rows, const int8_t* b_g, int rb, int PW,
                                                     int shift, int a_signed, int* d_out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    const int cpr = rb / 16;  // 16-byte chunks per row
    for (int i = tid; i < patch_rows * cpr; i += 128) {
        const int row = i / cpr, ch = i % cpr;
        const uint32_t off = 4096u + (uint32_t)row * rb + ch * 16;
        *reinterpret_cast<uint4*>(sm + swz(off, rb)) = *reinterpret_cast<const uint4*>(patch_g + (size_t)row * rb + ch * 16);
    }
    for (int i = tid; i < N * cpr; i += 128) {
        const int row = i / cpr, ch = i % cpr;
        const uint32_t off = (uint32_t)row * rb + ch * 16;
        *reinterpret_cast<uint4*>(sm + swz(off, rb)) = *reinterpret_cast<const uint4*>(b_g + (size_t)row * rb + ch * 16);
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
        const uint32_t lay = layout_code(rb);
        const uint64_t ad = make_desc(base + 4096 + (uint32_t)shift * rb, (uint32_t)PW * rb, lay);
        const uint64_t bd = make_desc(base, 8u * rb, lay);
        const uint32_t idesc = idesc_i8(N, a_signed, 1);
        for (int k = 0; k < rb / 32; ++k) mma_i8(tmem, ad + 2 * k, bd + 2 * k, idesc, k > 0);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    wait_bar(smem_u32(&bar), 0);
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
    for (int j = 0; j < 32; ++j) d_out[(warp * 32 + lane) * N + j] = (int)r[j];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

__global__ void __launch_bounds__(128) rate_kernel(int n, int iters, int rb, int a_sbo, int a_shift, long long* cycles_out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < (65536 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0x01010101u;
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
        const uint32_t idesc = idesc_i8(n, 0, 1);
        const uint32_t lay = layout_code(rb);
        const uint64_t ad = make_desc(base + a_shift, (uint32_t)a_sbo, lay);
        const uint64_t bd = make_desc(base + 65536, 8u * rb, lay);
        const int ks = rb / 32;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int k = 0; k < 4; ++k) mma_i8(tmem, ad + 2 * (k % ks), bd + 2 * (k % ks), idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        wait_bar(smem_u32(&bar), 0);
        cycles_out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
    // ---- rate
    long long* dc;
    cudaMalloc(&dc, 148 * sizeof(long long));
    const int smem_r = 1024 + 65536 + 32768;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r);
    struct Cfg { int rb, sbo, shift; const char* what; };
    const Cfg cfgs[] = {{32, 256, 0, "canonical rb32"},   {32, 384, 0, "5x5 s1 patch (12 px rows)"}, {32, 384, 13 * 32, "5x5 s1 patch, tap shift"},
                        {32, 1280, 0, "5x5 s2 patch (2 x 20 px rows)"}, {32, 1280, 33 * 32, "5x5 s2 patch, tap shift"},
                        {64, 512, 0, "canonical rb64"},   {64, 640, 0, "3x3 patch rb64 (10 px rows)"}, {64, 640, 11 * 64, "3x3 patch rb64, tap shift"},
                        {32, 320, 0, "3x3 patch rb32"},   {128, 1024, 0, "canonical rb128 (first layer)"}};
    for (const Cfg& c : cfgs)
        for (int n : {64, 96, 128}) {
            const int iters = 2000;
            rate_kernel<<<148, 128, smem_r>>>(n, iters, c.rb, c.sbo, c.shift, dc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("rate i8 N=%d: CUDA error %s\n", n, cudaGetErrorString(e));
                return 1;
            }
            std::vector<long long> cc(148);
            cudaMemcpy(cc.data(), dc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (long long v : cc) mx = v > mx ? v : mx;
            const double per = (double)mx / (iters * 4.0);
            printf("rate i8 %-34s rowbytes=%3d sbo=%4d N=%3d: %.1f cycles per 128xNx32 MMA (N/2 = %.1f)\n", c.what, c.rb, c.sbo, n, per, n / 2.0);
        }
    return 0;
    // ---- correctness
    const int MAXROWS = 16 * 24 + 16;
    std::vector<uint8_t> patch(MAXROWS * 128);
    std::vector<int8_t> b(N * 128);
    srand(1);
    for (auto& v : patch) v = (uint8_t)(rand() % 256);
    for (auto& v : b) v = (int8_t)(rand() % 256 - 128);
    uint8_t* dp;
    int8_t* db;
    int* dd;
    cudaMalloc(&dp, patch.size());
    cudaMalloc(&db, b.size());
    cudaMalloc(&dd, 128 * N * 4);
    cudaMemcpy(dp, patch.data(), patch.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice);
    const int smem = 1024 + 4096 + MAXROWS * 128;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<int> got(128 * N);
    for (int a_signed : {0, 1})
        for (int rb : {128, 64, 32})
            for (int PW : {8, 12, 16})
                for (int shift : {0, 1, 5, 13}) {
                    cudaMemset(dd, 0, 128 * N * 4);
                    probe_kernel<<<1, 128, smem>>>(dp, MAXROWS, db, rb, PW, shift, a_signed, dd);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) {
                        printf("a_signed=%d rb=%d PW=%d shift=%d CUDA error %s\n", a_signed, rb, PW, shift, cudaGetErrorString(e));
                        return 1;
                    }
                    cudaMemcpy(got.data(), dd, 128 * N * 4, cudaMemcpyDeviceToHost);
                    int bad = 0;
                    for (int r = 0; r < 128; ++r) {
                        const int prow = (r / 8) * PW + (r % 8) + shift;
                        for (int n = 0; n < N; ++n) {
                            int ref = 0;
                            for (int k = 0; k < rb; ++k) {
                                const int av = a_signed ? (int)(int8_t)patch[prow * rb + k] : (int)patch[prow * rb + k];
                                ref += av * (int)b[n * rb + k];
                            }
                            if (ref != got[r * N + n]) ++bad;
                        }
                    }
                    printf("i8 a_signed=%d rowbytes=%3d PW=%2d shift=%2d : %s (%d wrong)\n", a_signed, rb, PW, shift, bad ? "MISMATCH" : "ok", bad);
                }
    return 0;
}
