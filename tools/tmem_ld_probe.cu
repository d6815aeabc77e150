// Probe: tcgen05.ld (TMEM -> registers) throughput per SM as a function of the number of reading warps and of the loads in
// flight per warp.  Decides whether the fused blocks' epilogue (3 s32 digit planes = 48 KB of TMEM per 128x32 tile-step) is
// bound by the TMEM read port or by latency.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_probe tmem_ld_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int INFLIGHT>
__global__ void __launch_bounds__(512, 1) probe(int iters, int nwarps, long long* out, int* sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_slot;
    int acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        const uint32_t lane_base = base + ((uint32_t)((warp & 3) * 32) << 16);
        __syncwarp();
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            int r[INFLIGHT][16];
#pragma unroll
            for (int k = 0; k < INFLIGHT; ++k) {
                const uint32_t a = lane_base + (uint32_t)(((it * INFLIGHT + k) * 16) & 511);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[k][0]), "=r"(r[k][1]), "=r"(r[k][2]), "=r"(r[k][3]), "=r"(r[k][4]), "=r"(r[k][5]), "=r"(r[k][6]), "=r"(r[k][7]),
                      "=r"(r[k][8]), "=r"(r[k][9]), "=r"(r[k][10]), "=r"(r[k][11]), "=r"(r[k][12]), "=r"(r[k][13]), "=r"(r[k][14]), "=r"(r[k][15])
                    : "r"(a)
                    : "memory");
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int k = 0; k < INFLIGHT; ++k)
#pragma unroll
                for (int i = 0; i < 16; ++i) acc ^= r[k][i];
        }
        t1 = clock64();
    }
    if ((threadIdx.x & 31) == 0 && warp < nwarps) out[blockIdx.x * 16 + warp] = t1 - t0;
    if (acc == 0x12345678) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
    }
}

int main() {
    long long* out;
    int* sink;
    cudaMalloc(&out, 148 * 16 * sizeof(long long));
    cudaMalloc(&sink, 4);
    const int iters = 2000;
    printf("tcgen05.ld 32x32b.x16 (2 KB per warp-instruction), %d iterations, grid 148\n", iters);
    for (int inflight : {1, 3, 6}) {
        for (int nw : {1, 4, 8, 16}) {
            cudaMemset(out, 0, 148 * 16 * sizeof(long long));
            if (inflight == 1) probe<1><<<148, 512>>>(iters, nw, out, sink);
            else if (inflight == 3) probe<3><<<148, 512>>>(iters, nw, out, sink);
            else probe<6><<<148, 512>>>(iters, nw, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[148 * 16];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
            const double bytes = (double)iters * inflight * 2048.0 * nw;
            printf("in flight %d  warps %2d : %8lld cycles  -> %7.1f cycles per warp-load, %7.1f B/clk/SM\n", inflight, nw, mx,
                   (double)mx / (iters * inflight), bytes / (double)mx);
        }
    }
    return 0;
}
