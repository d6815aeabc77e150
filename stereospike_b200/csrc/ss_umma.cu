// tcgen05 (5th-gen tensor core) implementation of the fused spiking block for sm_100a.
//
// One CTA owns a tile of 128 output pixels x BN output channels and runs ALL T timesteps for it:
//   per timestep   D[128 x BN] (fp32, TMEM) = sum over K-blocks of  A[128 x 64] (bf16 spikes, smem)  x  W[BN x 64]^T
//   then the epilogue warps read D with tcgen05.ld, apply gain -> charge -> fire -> hard reset with the membrane
//   potential held in REGISTERS across the T loop (it never touches HBM between timesteps), add the skip /
//   SEW residual and write bf16 NHWC spikes.  Two TMEM accumulators ping-pong so the MMAs of timestep t+1 overlap
//   the neuron epilogue of timestep t.
//
// Warp roles (warp-specialised, mbarrier pipelines, no __syncthreads in the steady state):
//   warps 0-3   A producers: thread r gathers output pixel r's 64-element K slice (one 128-byte row) with
//               8 x cp.async(16 B) into a SWIZZLE_128B K-major tile.  The gather goes through a per-tile table of
//               source-pixel offsets built from (ymap, xmap), so strided zero-padded convs and the
//               nearest-neighbour-upsampled decoder convs are the same code (zero-fill for padded taps).
//   warp 4      MMA issuer (one elected lane issues tcgen05.mma, commits to mbarriers); owns the TMEM allocation.
//   warp 5      weight producer: TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) of `planes` bf16 weight planes.
//   warps 8..   epilogue (4 or 8 warps; TMEM lane quarter = warp % 4).
//
// fp32 parity: spikes are exact in bf16; fp32 weights are split into up to three bf16 planes (hi + mid + lo
// reproduce the fp32 value exactly) and each plane is one more MMA on the same A tile, accumulated in fp32 in TMEM.
//
// Replaces: Conv2d/NNConvUpsampling -> MultiplyBy -> IF/LIF/PLIF per timestep + skip adds
// (reference network/SNN_models.py:75-129,171-186; network/blocks.py:110-132,145-171).
#include <cuda.h>

#include "ss_common.cuh"

namespace ss {
namespace {

constexpr int TILE_M = 128;
constexpr int TILE_K = 64;                       // bf16 elements = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = TILE_M * TILE_K * 2;  // 16 KB
constexpr int MAX_TAPS = 25;
constexpr int LAG = 2;                           // cp.async groups kept in flight per producer thread

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (surfacing as a CUDA error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) __trap();
    }
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"): 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address, 16-byte units
    d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M x N.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct UmmaExtra {
    int planes;
    int stages;
    int nkb;  // K blocks of 64
};

template <int BN>
struct Cfg {
    static constexpr int NSPLIT = BN > 64 ? 2 : 1;      // column halves handled by separate epilogue warps
    static constexpr int CW = BN / NSPLIT;              // columns per epilogue thread
    static constexpr int EPI_WARPS = 4 * NSPLIT;
    static constexpr int THREADS = 32 * (8 + EPI_WARPS);
    // two ping-pong buffers x two accumulators (plane 0 | planes 1..) x BN fp32 columns
    static constexpr int TMEM_COLS = 4 * BN;
};

template <int BN>
__global__ void __launch_bounds__(Cfg<BN>::THREADS, 1)
conv_neuron_umma_kernel(const ConvParams p, const UmmaExtra e, const __grid_constant__ CUtensorMap wmap) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);

    const int b_stage_bytes = e.planes * BN * TILE_K * 2;
    const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;  // multiple of 1024
    uint8_t* tail = sm + (size_t)e.stages * stage_bytes;
    int* srcoff = reinterpret_cast<int*>(tail);                           // [ks*ks][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail + MAX_TAPS * TILE_M * 4);
    // bars: full[stages], empty[stages], acc_full[2], acc_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * 8 + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TILE_M;
    const int n0 = blockIdx.y * BN;
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = smem_u32(bars + 8);
    const uint32_t bar_accf = smem_u32(bars + 16);
    const uint32_t bar_acce = smem_u32(bars + 18);

    if (threadIdx.x == 0) {
        for (int s = 0; s < e.stages; ++s) {
            mbar_init(bar_full + 8 * s, TILE_M + 1);  // 128 A-producer threads + 1 expect_tx arrival
            mbar_init(bar_empty + 8 * s, 1);          // one tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_accf + 8 * i, 1);
            mbar_init(bar_acce + 8 * i, 32 * C::EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp < 4) {
        // source-pixel table of this tile: srcoff[tap][r] = pixel index inside one timestep, or -1
        const int r = threadIdx.x;
        const int m = m0 + r;
        const int HW = p.Hout * p.Wout;
        int b = 0, oy = 0, ox = 0;
        const bool live = m < p.M;
        if (live) {
            b = m / HW;
            const int q = m - b * HW;
            oy = q / p.Wout;
            ox = q - oy * p.Wout;
        }
        for (int ky = 0; ky < p.ks; ++ky) {
            const int sy = live ? __ldg(p.ymap + oy * p.ks + ky) : -1;
            for (int kx = 0; kx < p.ks; ++kx) {
                const int sx = live ? __ldg(p.xmap + ox * p.ks + kx) : -1;
                srcoff[(ky * p.ks + kx) * TILE_M + r] = (sy >= 0 && sx >= 0) ? (b * p.Hin + sy) * p.Win + sx : -1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ================================================================== A producers (cp.async gather)
        const int r = threadIdx.x;
        const uint32_t row_off = (uint32_t)r * 128u;
        const uint32_t xr = (uint32_t)(r & 7);
        const __nv_bfloat16* xin = reinterpret_cast<const __nv_bfloat16*>(p.x);
        const size_t t_stride = (size_t)p.B * p.Hin * p.Win * p.Cin;
        const int ntaps = p.ks * p.ks;
        const bool one_tap = (p.Cin % TILE_K) == 0;
        int stage = 0;
        uint32_t phase = 0;
        int issued = 0;       // K blocks issued so far (global counter over t and kb)
        int arrive_stage = 0; // stage of the oldest un-signalled block
        for (int t = 0; t < p.T; ++t) {
            const __nv_bfloat16* xt = xin + (size_t)t * t_stride;
            int tap = 0, c0 = 0;
            for (int kb = 0; kb < e.nkb; ++kb) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                const uint32_t dst = base + (uint32_t)stage * stage_bytes + row_off;
                if (one_tap) {
                    const int off = srcoff[tap * TILE_M + r];
                    const __nv_bfloat16* src = (off >= 0) ? xt + (size_t)off * p.Cin + c0 : xin;
                    const uint32_t nbytes = (off >= 0) ? 16u : 0u;
#pragma unroll
                    for (int j = 0; j < 8; ++j) cp_async_16(dst + (((uint32_t)j ^ xr) << 4), src + (off >= 0 ? j * 8 : 0), nbytes);
                    c0 += TILE_K;
                    if (c0 >= p.Cin) {
                        c0 = 0;
                        ++tap;
                    }
                } else {
                    // Cin in {8,16,32}: a 64-element K block spans several taps
                    int tp = tap, cc = c0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int off = (tp < ntaps) ? srcoff[tp * TILE_M + r] : -1;
                        const __nv_bfloat16* src = (off >= 0) ? xt + (size_t)off * p.Cin + cc : xin;
                        cp_async_16(dst + (((uint32_t)j ^ xr) << 4), src, (off >= 0) ? 16u : 0u);
                        cc += 8;
                        if (cc >= p.Cin) {
                            cc = 0;
                            ++tp;
                        }
                    }
                    tap = tp;
                    c0 = cc;
                }
                cp_async_commit();
                ++issued;
                if (issued > LAG) {
                    cp_async_wait<LAG>();
                    fence_proxy_async();
                    mbar_arrive(bar_full + 8 * arrive_stage);
                    if (++arrive_stage == e.stages) arrive_stage = 0;
                }
                if (++stage == e.stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
        // drain
        cp_async_wait<0>();
        fence_proxy_async();
        const int pending = issued < LAG ? issued : LAG;
        for (int i = 0; i < pending; ++i) {
            mbar_arrive(bar_full + 8 * arrive_stage);
            if (++arrive_stage == e.stages) arrive_stage = 0;
        }
    } else if (warp == 4) {
        // ================================================================== MMA issuer
        const uint32_t idesc = make_idesc(TILE_M, BN);
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < p.T; ++t) {
            const int buf = t & 1;
            mbar_wait(bar_acce + 8 * buf, (((uint32_t)t >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 2 * BN);
            for (int kb = 0; kb < e.nkb; ++kb) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = base + (uint32_t)stage * stage_bytes;
                    const uint64_t adesc = make_sw128_desc(a_addr);
                    for (int pl = 0; pl < e.planes; ++pl) {
                        const uint64_t bdesc = make_sw128_desc(a_addr + A_STAGE_BYTES + pl * BN * TILE_K * 2);
                        // Plane 0 (the bf16 head of every weight) accumulates alone: spike x bf16 products summed in an
                        // fp32 accumulator are then essentially error-free.  The residual planes (2^-9, 2^-17 of the
                        // weight) share a second accumulator whose own truncation is far below one fp32 ulp of the sum.
                        const uint32_t d = d_tmem + (pl > 0 ? (uint32_t)BN : 0u);
                        const int first = (pl <= 1) ? kb : 1;
#pragma unroll
                        for (int k = 0; k < TILE_K / 16; ++k) {
                            // +32 bytes per 16-element K step inside the 128-byte swizzle row (encoded >> 4)
                            umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (first | k) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(bar_empty + 8 * stage);
                    if (kb == e.nkb - 1) umma_commit(bar_accf + 8 * buf);
                }
                __syncwarp();
                if (++stage == e.stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 5) {
        // ================================================================== weight producer (TMA)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = 0; t < p.T; ++t) {
                for (int kb = 0; kb < e.nkb; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                    const uint32_t full = bar_full + 8 * stage;
                    mbar_arrive_expect_tx(full, (uint32_t)b_stage_bytes);
                    const uint32_t dstb = base + (uint32_t)stage * stage_bytes + A_STAGE_BYTES;
                    for (int pl = 0; pl < e.planes; ++pl)
                        tma_load_2d(dstb + pl * BN * TILE_K * 2, &wmap, full, kb * TILE_K, pl * p.Cout + n0);
                    if (++stage == e.stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp >= 8) {
        // ================================================================== epilogue: TMEM -> neuron -> HBM
        const int ew = warp - 8;
        const int quarter = ew & 3;          // == warp % 4: the TMEM lanes this warp may touch
        const int chalf = ew >> 2;
        const int row = quarter * 32 + lane;
        const int m = m0 + row;
        const bool live = m < p.M;
        const int col0 = chalf * C::CW;
        float decay = 0.0f;
        if (p.neuron == SS_NEURON_PLIF) decay = __ldg(p.decay);
        float v[C::CW];
        if (p.v_in != nullptr && live) {
            const float4* vi = reinterpret_cast<const float4*>(p.v_in + (size_t)m * p.Cout + n0 + col0);
#pragma unroll
            for (int j = 0; j < C::CW / 4; ++j) {
                const float4 q = __ldg(vi + j);
                v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < C::CW; ++j) v[j] = p.v_reset;
        }
        for (int t = 0; t < p.T; ++t) {
            const int buf = t & 1;
            mbar_wait(bar_accf + 8 * buf, ((uint32_t)t >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 2 * BN + col0);
            const size_t o = ((size_t)t * p.M + m) * p.Cout + n0 + col0;
#pragma unroll
            for (int ch = 0; ch < C::CW / 16; ++ch) {
                uint32_t acc[16];
                tmem_ld16(taddr + ch * 16, acc);
                if (e.planes > 1) {
                    uint32_t rest[16];
                    tmem_ld16(taddr + BN + ch * 16, rest);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        acc[j] = __float_as_uint(__fadd_rn(__uint_as_float(acc[j]), __uint_as_float(rest[j])));
                } else {
                    tmem_ld_wait();
                }
                float s[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float h;
                    s[j] = neuron_step(p.neuron, __fmul_rn(__uint_as_float(acc[j]), p.gain), v[ch * 16 + j], p.v_th,
                                       p.v_reset, p.tau, decay, h);
                    acc[j] = __float_as_uint(h);
                }
                if (live) {
                    if (p.h_seq != nullptr) {
                        float4* hp = reinterpret_cast<float4*>(p.h_seq + o + ch * 16);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            hp[j] = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                                                __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
                    }
                    if (p.resid != nullptr) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.resid + o + ch * 16);
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint4 raw = __ldg(rp + j);
                            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float2 f = __bfloat1622float2(h2[q]);
                                s[8 * j + 2 * q] += f.x;
                                s[8 * j + 2 * q + 1] += f.y;
                            }
                        }
                    }
                    uint4* op = reinterpret_cast<uint4*>(p.out + o + ch * 16);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        uint4 pk;
                        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                        for (int q = 0; q < 4; ++q) h2[q] = __floats2bfloat162_rn(s[8 * j + 2 * q], s[8 * j + 2 * q + 1]);
                        op[j] = pk;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_acce + 8 * buf);
        }
        if (p.v_out != nullptr && live) {
            float4* vo = reinterpret_cast<float4*>(p.v_out + (size_t)m * p.Cout + n0 + col0);
#pragma unroll
            for (int j = 0; j < C::CW / 4; ++j) vo[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

template <int BN>
int launch_bn(const ConvParams& p, const void* w_umma, int planes, cudaStream_t st) {
    using C = Cfg<BN>;
    const int Kpad = (p.K + TILE_K - 1) / TILE_K * TILE_K;
    EncodeTiledFn enc = get_encode_fn();
    if (enc == nullptr) {
        set_error("umma: cuTensorMapEncodeTiled not available from the driver");
        return SS_ECUDA;
    }
    CUtensorMap wmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)Kpad, (cuuint64_t)planes * p.Cout};
    const cuuint64_t gstride[1] = {(cuuint64_t)Kpad * 2};
    const cuuint32_t box[2] = {(cuuint32_t)TILE_K, (cuuint32_t)BN};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(&wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_umma), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("umma: cuTensorMapEncodeTiled failed (%d)", (int)cr);
        return SS_ECUDA;
    }
    UmmaExtra e;
    e.planes = planes;
    e.nkb = Kpad / TILE_K;
    const int stage_bytes = A_STAGE_BYTES + planes * BN * TILE_K * 2;
    const int tail_bytes = MAX_TAPS * TILE_M * 4 + 256;
    const int budget = 220 * 1024 - 1024 - tail_bytes;
    int stages = budget / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < LAG + 1) {
        set_error("umma: not enough shared memory for %d planes at BN=%d", planes, BN);
        return SS_EUNSUPPORTED;
    }
    e.stages = stages;
    const size_t smem = 1024 + (size_t)stages * stage_bytes + tail_bytes;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv_neuron_umma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess)
            return check_launch("umma: cudaFuncSetAttribute");
        attr_set = true;
    }
    dim3 grid((p.M + TILE_M - 1) / TILE_M, p.Cout / BN);
    conv_neuron_umma_kernel<BN><<<grid, C::THREADS, smem, st>>>(p, e, wmap);
    count_launch();
    return check_launch("conv_neuron_umma");
}

}  // namespace

int launch_conv_neuron_umma(const ConvParams& p, const void* w_umma, int planes, cudaStream_t st) {
    if (p.ks * p.ks > MAX_TAPS) {
        set_error("umma: at most %d taps", MAX_TAPS);
        return SS_EUNSUPPORTED;
    }
    if (p.Cin % 8 != 0 || (p.Cin < TILE_K && (TILE_K % p.Cin) != 0) || (p.Cin > TILE_K && (p.Cin % TILE_K) != 0)) {
        set_error("umma: unsupported Cin %d", p.Cin);
        return SS_EUNSUPPORTED;
    }
    if (p.Cout % 128 == 0) return launch_bn<128>(p, w_umma, planes, st);
    if (p.Cout % 64 == 0) return launch_bn<64>(p, w_umma, planes, st);
    if (p.Cout % 32 == 0) return launch_bn<32>(p, w_umma, planes, st);
    set_error("umma: Cout %% 32 != 0");
    return SS_EUNSUPPORTED;
}

}  // namespace ss
