// Weight gradient of the fused spiking block on the sm_100a tensor cores (tcgen05, bf16 x bf16 -> fp32).
//
// Replaces: autograd's cuDNN wgrad of every Conv2d on the path, including the convs behind UpsamplingNearest2d
// (reference train.py:239 -> network/blocks.py:110-132, network/SNN_models.py:75-129).
//
//   g_w[ky][kx][c][n] = sum over (t, b, oy, ox) of  g[t][b][oy][ox][n] * x[t][b][src(oy, ky)][src(ox, kx)][c]
//
// The contraction runs over PIXELS, the slow dimension of NHWC, so both MMA operands are MN-major: a shared-memory row is
// one pixel (a K index), its bytes are channels (the M / N index).  That is exactly the layout of the forward kernel's halo
// patch, so the same trick applies: the input patch of a 16 x 8 output tile is staged ONCE per (tile, timestep, channel
// chunk) and every filter tap is a different start address of the B descriptor (probe: tools/umma_probe_mn.cu -- the
// hardware swizzles on absolute address bits for MN-major operands too).  Stride 2 = parity-split patch rows, upsampling =
// gather in the producer, exactly as in ss_conv_i8.cu.
//
//   A = g tile      [K = 16 pixels (2 tile rows x 8)][M = 128 output channels]   bf16, SWIZZLE_128B, 2 channel blocks (LBO)
//   B = x patch     [K = 16 pixels, shifted by tap  ][N = 16 / 32 input channels] u8 spikes converted to bf16 by the producer
//   D = one fp32 accumulator [128][N] PER TAP in TMEM, alive for the whole CTA.  512 columns hold 16 accumulators of N = 32:
//       a 5x5 filter is split into two groups of filter rows (3 + 2) handled by different CTAs
//
// A CTA owns one (128-channel block of Cout, N-channel chunk of Cin) pair and a contiguous range of (tile, timestep) units;
// the accumulators never leave TMEM until the end, when four warps add them into g_w with coalesced atomics.
//
// Few output channels (Cout <= 64: the full-resolution blocks, where most pixels are).  M = 128 lanes would be mostly
// padding, so the A operand is made of SH = 2 / 4 copies of the SAME g patch, copy i starting one pixel after copy i-1
// (leading-byte-offset = one patch pixel): accumulator rows [i*CH, (i+1)*CH) then hold the tap whose column offset is
// (SH-1-i)*stride further right, and one MMA produces SH taps -- 15 or 10 MMAs per K step instead of 25.  The g patch gets
// SH-1 halo columns on the left so that every copy still sees every pixel exactly once across the tiles of a row.
//
// N = 64 input channels per accumulator (stride-1 blocks with Cin % 64 == 0).  Both operands are MN-major, and the tensor core
// reads MN-major shared memory at ~85 B/clk: an MMA costs ~57 cycles at N = 32 and ~70 at N = 64 (tools/wgrad_probe.py: the
// issuing thread sits in the issue queue 87 % of the time), so twice the channels per MMA is 1.6x the throughput.  512 TMEM
// columns then hold 8 accumulators: the KS x (column shifts) accumulators of a filter are cut into balanced groups of <= 8
// (5x5: 6+6+6+7, 3x3: 4+5, 2 copies: 5+5+5, 4 copies: 5+5), one CTA family per group.  The groups of one unit range sit on
// neighbouring CTAs and advance in lock step, so the g tiles / x patches they all read come from HBM once.
//
// Warp roles: 0-3 x-patch producers (LDG u8 -> bf16 -> swizzled STS), 4-7 g-tile producers (cp.async), 8 MMA issuer.
#include <cuda_bf16.h>
#include <cstdlib>

#include "ss_common.cuh"
#include "ss_umma.cuh"

namespace ss {
namespace {

constexpr int WG_THREADS = 288;
constexpr int WG_MAX_STAGES = 6;
constexpr int G_TILE_BYTES = 32768;          // 2 channel blocks x 128 pixels x 128 B

struct WgParams {
    int T, B, Hin, Win, Cin, Hout, Wout, Cout;
    int pad, upsample;
    int HsO, Hup, Wup;
    int tiles_x, mtiles;
    int nblk, nchunk, nsplit;    // grid = nblk * nchunk * nsplit * groups
    int NPS;
    int cin_real;                // channels of x that exist in g_w's K index (Cin, or <= 4 for the packed first layer)
    int x_small;                 // 1 = every x value is < 128 (spikes / spike sums): fast u8 -> bf16 conversion
    float yscale, xscale;
    const uint8_t* x;
    const __nv_bfloat16* g;
    float* g_w;
    int dbg;                     // instrumentation build only (-DSS_ROLE_TIMING, SS_WG_DBG): 1 = no MMA issue, 2 = no x staging, 4 = no g staging
};
#ifdef SS_ROLE_TIMING
#define WG_DBG(bit) ((p.dbg & (bit)) != 0)
// per-role cycle counters of CTA 0 (tools/wgrad_probe.py): [role][slot], role 0 = x producer, 1 = g producer, 2 = MMA thread
__device__ unsigned long long ss_wg_dbg[3 * 8];
#define WG_DECL() unsigned long long wg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long wg_start = clock64()
#define WG_T0() const long long wg_t0 = clock64()
#define WG_ACC(slot) wg_acc[slot] += (unsigned long long)(clock64() - wg_t0)
#define WG_DUMP(role)                                                                              \
    do {                                                                                           \
        wg_acc[7] = (unsigned long long)(clock64() - wg_start);                                    \
        if (blockIdx.x == 0)                                                                       \
            for (int _i = 0; _i < 8; ++_i) ss_wg_dbg[(role) * 8 + _i] = wg_acc[_i];                 \
    } while (0)
#else
#define WG_DBG(bit) false
#define WG_DECL()
#define WG_T0()
#define WG_ACC(slot)
#define WG_DUMP(role)
#endif

// MN-major descriptor: lbo = bytes between channel blocks, sbo = bytes between 8-pixel groups
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// same MMA with the descriptors given as (low, high) words and the accumulate flag as an immediate
template <int ACC>
__device__ __forceinline__ void umma_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, %6;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACC)
        : "memory");
}

template <int STRIDE>
__device__ __forceinline__ int wg_row_source(const WgParams& p, int ty, int pr) {
    const int gi = ty * 16 * STRIDE + pr;
    const int per = STRIDE * p.HsO;
    const int b = gi / per;
    const int local = gi - b * per;
    if (b >= p.B) return -1;
    if (p.upsample) {
        if (local >= p.Hup) return -1;
        const int iy = min((int)floorf((float)local * p.yscale), p.Hin - 1);   // ATen upsample_nearest index rule
        return b * p.Hin + iy;
    }
    const int iy = local - p.pad;
    return (iy >= 0 && iy < p.Hin) ? b * p.Hin + iy : -1;
}
template <int STRIDE, int PWHALF>
__device__ __forceinline__ int wg_col_source(const WgParams& p, int tx, int pc) {
    if (p.upsample) {
        const int u = tx * 8 + pc;
        if (u >= p.Wup) return -1;
        return min((int)floorf((float)u * p.xscale), p.Win - 1);
    }
    int ix;
    if (STRIDE == 1) {
        ix = tx * 8 + pc - p.pad;
    } else {
        const int plane = pc / PWHALF;
        const int idx = pc - plane * PWHALF;
        ix = 2 * (tx * 8 - p.pad / 2 + idx) + plane;
    }
    return (ix >= 0 && ix < p.Win) ? ix : -1;
}

// 16 u8 -> 16 bf16 (exact: values <= 255 have <= 8 significant bits)
__device__ __forceinline__ void cvt16(const uint4 v, uint4& lo, uint4& hi) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 a = __floats2bfloat162_rn((float)(w[i] & 0xFFu), (float)((w[i] >> 8) & 0xFFu));
        const __nv_bfloat162 b = __floats2bfloat162_rn((float)((w[i] >> 16) & 0xFFu), (float)(w[i] >> 24));
        o[2 * i] = *reinterpret_cast<const uint32_t*>(&a);
        o[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&b);
    }
    lo = make_uint4(o[0], o[1], o[2], o[3]);
    hi = make_uint4(o[4], o[5], o[6], o[7]);
}

// 16 u8 -> 16 bf16 for values < 128 (spikes {0,1} and spike sums {0..3}: every activation behind the first layer): the byte is
// dropped into the mantissa of bf16 128.0 (0x4300 | b == 128 + b exactly, 7 mantissa bits) and 128 is subtracted two lanes at a
// time -- 2 PRMT + 2 HSUB2 per four values instead of 4 I2F (quarter-rate pipe) + 2 F2FP, which made the x producers the
// bottleneck of the N = 64 kernels (tools/wgrad_probe.py).
__device__ __forceinline__ void cvt16_small(const uint4 v, uint4& lo, uint4& hi) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[8];
    const uint32_t k128 = 0x43004300u;
    const __nv_bfloat162 m = *reinterpret_cast<const __nv_bfloat162*>(&k128);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t p01 = __byte_perm(w[i], 0x43434343u, 0x4140);   // bytes {b0, 0x43, b1, 0x43}
        const uint32_t p23 = __byte_perm(w[i], 0x43434343u, 0x4342);   // bytes {b2, 0x43, b3, 0x43}
        const __nv_bfloat162 a = __hsub2(*reinterpret_cast<const __nv_bfloat162*>(&p01), m);
        const __nv_bfloat162 b = __hsub2(*reinterpret_cast<const __nv_bfloat162*>(&p23), m);
        o[2 * i] = *reinterpret_cast<const uint32_t*>(&a);
        o[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&b);
    }
    lo = make_uint4(o[0], o[1], o[2], o[3]);
    hi = make_uint4(o[4], o[5], o[6], o[7]);
}

// NB: input channels per CTA (16 -> 32-byte patch rows, 32 -> 64-byte rows).  FIRST4: x is the packed first-layer input
// u8 [..][4]; its 4 channels are widened to one 16-channel chunk (channels 4..15 zero).
// column shifts of the B operand issued per filter row, and how many of them
template <int KS, int STRIDE, int SH>
__host__ __device__ constexpr int wg_nshift() {
    return SH == 1 ? KS : (STRIDE == 1 ? (KS + SH - 1) / SH : 3);
}
template <int KS, int STRIDE, int SH>
__host__ __device__ constexpr int wg_shift(int si) {
    return SH == 1 ? si : (STRIDE == 1 ? si * SH : si);     // stride 1: {0, SH, 2*SH..};  stride 2, SH 2: {0, 1, 2}
}

// filter rows per CTA: as many as fit 512 TMEM columns, then balanced over the resulting number of groups (5 rows, 4 fit -> 3 + 2)
template <int KS, int NB, int NS>
__host__ __device__ constexpr int wg_rows_per_group() {
    constexpr int fit = (512 / NB) / NS < KS ? (512 / NB) / NS : KS;
    constexpr int ngrp = (KS + fit - 1) / fit;
    return (KS + ngrp - 1) / ngrp;
}
// accumulator groups of the N = 64 kernels: <= 8 accumulators each; group g owns accumulators [NALL*g/G, NALL*(g+1)/G)
template <int NALL>
__host__ __device__ constexpr int wg_tap_groups() {
    return (NALL + 7) / 8;
}

template <int KS, int STRIDE, int NB, bool FIRST4, int SH>
__global__ void __launch_bounds__(WG_THREADS, 1) conv_wgrad_umma_kernel(const WgParams p) {
    constexpr int RBX = 2 * NB;
    constexpr int CH = 128 / SH;                        // output channels per copy of the g patch
    constexpr int RBG = SH == 1 ? 128 : CH * 2;         // bytes per g-patch pixel (= its swizzle width)
    constexpr int GPW = 8 + SH - 1;                     // g-patch columns: SH-1 halo columns on the left
    constexpr int NPIXG = 16 * GPW;
    constexpr int G_BYTES = SH == 1 ? G_TILE_BYTES : (NPIXG * RBG + 1023) / 1024 * 1024;
    constexpr int NS = wg_nshift<KS, STRIDE, SH>();
    // TAPMODE (N = 64): a CTA owns a contiguous, balanced range of the KS * NS accumulators (filter row x column shift, row-major)
    // instead of whole filter rows, so that every group issues the same number of MMAs (+-1)
    constexpr bool TAPMODE = NB == 64;
    constexpr int NALL = KS * NS;                        // accumulators of the whole filter
    constexpr int GK = wg_rows_per_group<KS, NB, NS>();  // filter rows per CTA (TMEM: 512 columns)
    constexpr int NGRP = TAPMODE ? wg_tap_groups<NALL>() : (KS + GK - 1) / GK;          // groups of accumulators / of filter rows
    constexpr int NACC = TAPMODE ? (NALL + NGRP - 1) / NGRP : GK * NS;   // accumulators ([128][NB] each) of one CTA
    static_assert(!TAPMODE || STRIDE == 1, "accumulator groups: stride-1 blocks");
    static_assert(SH == 1 || (KS == 5 && (STRIDE == 1 || SH == 2)), "shifted copies: 5x5 only; stride 2 with SH = 2");
    constexpr int cPWhalf = 8 + (KS - 1) / 2;
    constexpr int cPWp = STRIDE == 1 ? 8 + KS - 1 : 2 * cPWhalf;
    constexpr int cPH = 15 * STRIDE + KS;
    constexpr int cPPIX = cPH * cPWp;
    constexpr int cPBX = (cPPIX * RBX + 1023) / 1024 * 1024;
    constexpr int cSTAGE = G_BYTES + cPBX;
    constexpr int cNPIX = (cPPIX + 127) / 128;
    constexpr uint32_t swz_mask = (uint32_t)(RBX >> 4) - 1u;
    static_assert(NACC * NB <= 512, "accumulators do not fit TMEM");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);
    uint8_t* tail = sm + (size_t)p.NPS * cSTAGE;
    int* rowsrc = reinterpret_cast<int*>(tail);              // [40]
    int* colsrc = rowsrc + 40;                               // [24]
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 256);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * WG_MAX_STAGES + 1);
    const uint32_t bar_full_x = smem_u32(bars);
    const uint32_t bar_full_g = smem_u32(bars + WG_MAX_STAGES);
    const uint32_t bar_empty = smem_u32(bars + 2 * WG_MAX_STAGES);
    const uint32_t bar_done = smem_u32(bars + 3 * WG_MAX_STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < WG_MAX_STAGES; ++s) {
            mbar_init(bar_full_x + 8 * s, 128);
            mbar_init(bar_full_g + 8 * s, 128);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this CTA's (output-channel block, input-channel chunk, unit range)
    // The groups of one (block, chunk, unit range) sit on neighbouring CTAs and walk the same units at the same pace (equal MMA
    // counts), so the g tiles and x patches they all read come from HBM once and from L2 afterwards: handing the groups different
    // numbers of CTAs de-synchronised them and cost 15-30 % on the full-resolution blocks (0.2-0.5 GB of g per pass).
    int item = blockIdx.x;
    const int grp = item % NGRP; item /= NGRP;
    const int nsplit = p.nsplit;
    const int ky0 = TAPMODE ? 0 : grp * GK;
    const int nky = min(GK, KS - ky0);
    const int tap0 = NALL * grp / NGRP;                       // TAPMODE: this CTA's accumulators [tap0, tap0 + ntap) of the filter
    const int ntap = NALL * (grp + 1) / NGRP - tap0;
    const int split = item % nsplit; item /= nsplit;
    const int chunk = item % p.nchunk;
    const int nblk = item / p.nchunk;
    const long long U = (long long)p.mtiles * p.T;
    const int u0 = (int)(U * split / nsplit);
    const int u1 = (int)(U * (split + 1) / nsplit);
    const int n0 = nblk * 128;
    const int c0 = chunk * NB;

    if (warp < 4) {
        // ================================================================== x-patch producers
        const int tid = threadIdx.x;
        const int xpix_bytes = FIRST4 ? 4 : p.Cin;
        const size_t t_stride = (size_t)p.B * p.Hin * p.Win * xpix_bytes;
        int stage = 0;
        uint32_t phase = 0;
        int cur_mt = -1;
        int goff[cNPIX];
        // Software pipeline: the global loads of unit u + 1 are issued BEFORE unit u is converted and stored, so their L2 round
        // trip overlaps the conversion instead of adding to every stage (the loads / wait / convert / store chain was serial).
        using Raw = uint4[cNPIX][NB / 16];
        auto issue_loads = [&](int u, Raw& raw) {
            const int mt = u / p.T;
            const int t = u - mt * p.T;
            if (mt != cur_mt) {
                cur_mt = mt;
                const int ty = mt / p.tiles_x, tx = mt - ty * p.tiles_x;
                named_sync(1, 128);     // everyone is done with the previous tables
                if (tid < cPH) rowsrc[tid] = wg_row_source<STRIDE>(p, ty, tid);
                if (tid >= 64 && tid - 64 < cPWp) colsrc[tid - 64] = wg_col_source<STRIDE, cPWhalf>(p, tx, tid - 64);
                named_sync(1, 128);
#pragma unroll
                for (int i = 0; i < cNPIX; ++i) {
                    const int pix = tid + i * 128;
                    goff[i] = -2;
                    if (pix < cPPIX) {
                        const int pr = pix / cPWp;
                        const int pc = pix - pr * cPWp;
                        const int r = rowsrc[pr], c = colsrc[pc];
                        goff[i] = (r >= 0 && c >= 0) ? (r * p.Win + c) * xpix_bytes : -1;
                    }
                }
            }
            const uint8_t* xt = p.x + (size_t)t * t_stride + (FIRST4 ? 0 : c0);
#pragma unroll
            for (int i = 0; i < cNPIX; ++i) {
#pragma unroll
                for (int h = 0; h < NB / 16; ++h) raw[i][h] = make_uint4(0u, 0u, 0u, 0u);
                if (goff[i] >= 0 && !WG_DBG(2)) {
                    if (FIRST4) {
                        raw[i][0].x = __ldg(reinterpret_cast<const uint32_t*>(xt + goff[i]));
                    } else {
#pragma unroll
                        for (int h = 0; h < NB / 16; ++h) raw[i][h] = __ldg(reinterpret_cast<const uint4*>(xt + goff[i]) + h);
                    }
                }
            }
        };
        WG_DECL();
        auto convert_store = [&](const Raw& raw) {
            {
                WG_T0();
                mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                WG_ACC(0);      // wait for a free stage
            }
            WG_T0();
            uint8_t* dst0 = sm + (size_t)stage * cSTAGE + G_BYTES;
#pragma unroll
            for (int i = 0; i < cNPIX; ++i) {
                if ((int)(tid + i * 128) < cPPIX && !WG_DBG(2)) {
                    const uint32_t off = (uint32_t)(tid + i * 128) * RBX;
#pragma unroll
                    for (int h = 0; h < NB / 16; ++h) {
                        uint4 lo, hi;
                        if (FIRST4 || !p.x_small) cvt16(raw[i][h], lo, hi);   // event counts: the full 0..255 range
                        else cvt16_small(raw[i][h], lo, hi);                  // spikes and spike sums (< 128): two values per instruction
                        *reinterpret_cast<uint4*>(dst0 + swizzle_off(off + h * 32, swz_mask)) = lo;
                        *reinterpret_cast<uint4*>(dst0 + swizzle_off(off + h * 32 + 16, swz_mask)) = hi;
                    }
                }
            }
            fence_proxy_async();        // generic-proxy stores -> visible to the tensor core's async-proxy reads
            mbar_arrive(bar_full_x + 8 * stage);
            WG_ACC(1);          // convert + store + fence
            if (++stage == p.NPS) {
                stage = 0;
                phase ^= 1u;
            }
        };
        Raw raw_a, raw_b;
        if (u0 < u1) issue_loads(u0, raw_a);
        for (int u = u0; u < u1; u += 2) {
            if (u + 1 < u1) issue_loads(u + 1, raw_b);
            convert_store(raw_a);
            if (u + 1 < u1) {
                if (u + 2 < u1) issue_loads(u + 2, raw_a);
                convert_store(raw_b);
            }
        }
        // ================================================================== epilogue: accumulators -> g_w (atomics)
        if (u1 > u0) {
            WG_T0();
            mbar_wait(bar_done, 0);
            WG_ACC(2);          // wait for the last MMA
            tc_fence_after();
            const int L = warp * 32 + lane;               // TMEM lane = (copy, output channel)
            const int copy = L / CH;
            const int n = n0 + (L - copy * CH);
            const bool n_ok = n < p.Cout;
            const int Kc = p.cin_real;                    // g_w row = tap * Kc + channel
            for (int a = 0; a < (TAPMODE ? ntap : nky * NS); ++a) {
                const int ky = TAPMODE ? (tap0 + a) / NS : ky0 + a / NS;
                const int shift = wg_shift<KS, STRIDE, SH>(TAPMODE ? (tap0 + a) - ky * NS : a % NS);
                const int kx = shift + STRIDE * (SH - 1 - copy);
                const int tap = ky * KS + kx;
                // stride 2: shift 2 of the un-shifted copy is the same tap as shift 0 of the shifted one -- count it once
                const bool dup = STRIDE == 2 && SH == 2 && copy == 1 && shift == 2;
#pragma unroll
                for (int h = 0; h < NB / 16; ++h) {
                    int d[16];
                    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * NB + h * 16), d);
                    tmem_ld_wait();
                    if (n_ok && kx < KS && !dup) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int c = c0 + h * 16 + i;
                            const float v = __int_as_float(d[i]);
                            if (c < Kc && v != 0.0f) atomicAdd(p.g_w + ((size_t)tap * Kc + c) * p.Cout + n, v);
                        }
                    }
                }
            }
            tc_fence_before();
        }
        if (threadIdx.x == 0) WG_DUMP(0);
    } else if (warp < 8) {
        // ================================================================== g-tile producers (cp.async, zero fill)
        const int m = threadIdx.x - 128;
        const size_t t_stride = (size_t)p.B * p.Hout * p.Wout * p.Cout;
        int stage = 0;
        uint32_t phase = 0;
        int cur_mt = -1;
        if constexpr (SH == 1) {
            // tile pixel m: row m >> 3, column m & 7; 128 channels = 2 blocks of 64 (16 KB apart)
            long long pix_off = -1;
            for (int u = u0; u < u1; ++u) {
                const int mt = u / p.T;
                const int t = u - mt * p.T;
                if (mt != cur_mt) {
                    cur_mt = mt;
                    const int ty = mt / p.tiles_x, tx = mt - ty * p.tiles_x;
                    const int so = ty * 16 + (m >> 3);
                    const int b = so / p.HsO;
                    const int oy = so - b * p.HsO;
                    const int ox = tx * 8 + (m & 7);
                    pix_off = (b < p.B && oy < p.Hout && ox < p.Wout) ? ((long long)(b * p.Hout + oy) * p.Wout + ox) * p.Cout : -1;
                }
                mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                const uint32_t dst0 = base + (uint32_t)stage * cSTAGE + (uint32_t)m * 128u;
                const __nv_bfloat16* src = p.g + (size_t)t * t_stride + (pix_off >= 0 ? pix_off : 0) + n0;
#pragma unroll
                for (int c = 0; c < (WG_DBG(4) ? 0 : 16); ++c) {
                    const bool ok = pix_off >= 0 && n0 + c * 8 < p.Cout;
                    const uint32_t dst = dst0 + (uint32_t)(c >> 3) * 16384u + ((uint32_t)((c & 7) ^ (m & 7)) << 4);
                    cp_async_16(dst, ok ? (const void*)(src + c * 8) : (const void*)p.g, ok ? 16u : 0u);
                }
                cp_async_arrive_noinc(bar_full_g + 8 * stage);
                if (++stage == p.NPS) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        } else {
            // g patch: 16 rows x GPW columns (output columns tx*8 - (SH-1) .. tx*8 + 7), CH channels per pixel
            constexpr int NPT = (NPIXG + 127) / 128;
            constexpr uint32_t gmask = (uint32_t)(RBG >> 4) - 1u;
            long long pix_off[NPT];
            for (int u = u0; u < u1; ++u) {
                const int mt = u / p.T;
                const int t = u - mt * p.T;
                if (mt != cur_mt) {
                    cur_mt = mt;
                    const int ty = mt / p.tiles_x, tx = mt - ty * p.tiles_x;
#pragma unroll
                    for (int i = 0; i < NPT; ++i) {
                        const int pix = m + i * 128;
                        const int r = pix / GPW, cc = pix - r * GPW;
                        const int so = ty * 16 + r;
                        const int b = so / p.HsO;
                        const int oy = so - b * p.HsO;
                        const int ox = tx * 8 - (SH - 1) + cc;
                        pix_off[i] = (pix < NPIXG && b < p.B && oy < p.Hout && ox >= 0 && ox < p.Wout)
                                         ? ((long long)(b * p.Hout + oy) * p.Wout + ox) * p.Cout : -1;
                    }
                }
                mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                const uint32_t dst0 = base + (uint32_t)stage * cSTAGE;
#pragma unroll
                for (int i = 0; i < NPT; ++i) {
                    const int pix = m + i * 128;
                    if (pix < NPIXG && !WG_DBG(4)) {
                        const __nv_bfloat16* src = p.g + (size_t)t * t_stride + (pix_off[i] >= 0 ? pix_off[i] : 0);
#pragma unroll
                        for (int c = 0; c < RBG / 16; ++c) {
                            const bool ok = pix_off[i] >= 0 && c * 8 < p.Cout;
                            cp_async_16(dst0 + swizzle_off((uint32_t)pix * RBG + c * 16, gmask), ok ? (const void*)(src + c * 8) : (const void*)p.g,
                                        ok ? 16u : 0u);
                        }
                    }
                }
                cp_async_arrive_noinc(bar_full_g + 8 * stage);
                if (++stage == p.NPS) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 8) {
        // ================================================================== MMA issuer
        if (elect_one()) {
            // kind::f16: D f32, A and B bf16, both MN-major, M = 128, N = NB
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NB >> 3) << 17) |
                                       ((uint32_t)(128 >> 4) << 24);
            constexpr uint32_t b_layout = RBX == 128 ? 2u : (RBX == 64 ? 4u : 6u);
            constexpr uint32_t b_sbo = (uint32_t)(STRIDE * cPWp * RBX);
            int stage = 0;
            uint32_t phase = 0;
            // per-accumulator constants of this CTA: descriptor offset (16-byte units) of the tap's patch shift, and whether it exists
            uint32_t boff[NACC];
            uint32_t live = 0u;
#pragma unroll
            for (int ai = 0; ai < NACC; ++ai) {
                int ky, kx;
                bool ok;
                if constexpr (TAPMODE) {
                    const int acc = tap0 + ai;
                    ky = acc / NS;
                    kx = wg_shift<KS, STRIDE, SH>(acc - ky * NS);
                    ok = ai < ntap;
                } else {
                    ky = ai / NS;             // relative to this CTA's first filter row
                    kx = wg_shift<KS, STRIDE, SH>(ai - ky * NS);
                    ok = ky < nky;
                }
                const int toff = STRIDE == 1 ? ky * cPWp + kx : ky * cPWp + (kx & 1) * cPWhalf + (kx >> 1);
                boff[ai] = (uint32_t)(toff * RBX) >> 4;
                if (ok) live |= 1u << ai;
            }
            WG_DECL();
            for (int u = u0; u < u1; ++u) {
                {
                    WG_T0();
                    mbar_wait(bar_full_x + 8 * stage, phase);
                    WG_ACC(0);      // wait for the x patch
                }
                {
                    WG_T0();
                    mbar_wait(bar_full_g + 8 * stage, phase);
                    WG_ACC(1);      // wait for the g tile
                }
                WG_T0();
                fence_proxy_async();      // the g tile arrived through cp.async (generic proxy)
                tc_fence_after();
                const uint32_t sbase = base + (uint32_t)stage * cSTAGE;
                // A: SH == 1: two 64-channel blocks 16 KB apart, 8-pixel groups 1 KB apart;
                //    SH  > 1: SH copies one pixel apart, 8-pixel groups one patch row apart
                constexpr uint32_t a_layout = RBG == 128 ? 2u : 4u;
                constexpr uint32_t a_kstep = SH == 1 ? 2048u : (uint32_t)(2 * GPW * RBG);
                const uint64_t a0 = SH == 1 ? make_desc_mn(sbase, 16384u, 1024u, 2u)
                                            : make_desc_mn(sbase, (uint32_t)RBG, (uint32_t)(GPW * RBG), a_layout);
                const uint64_t b0 = make_desc_mn(sbase + G_BYTES + (uint32_t)(ky0 * cPWp * RBX), 16u, b_sbo, b_layout);
                // The issuing thread is a single instruction stream: at ~25 scalar instructions per MMA (tap -> (ky, kx) division,
                // 64-bit descriptor adds, predicate set-up) it spent ~70 cycles per MMA and WAS the bound of the N = 64 kernels
                // (tools/wgrad_probe.py).  Everything tap-dependent is hoisted into boff[] / live, the K loop is unrolled, and
                // the descriptors advance by 32-bit adds on their low words (the 14-bit address field cannot carry out of it).
                const uint32_t a_hi = (uint32_t)(a0 >> 32), b_hi = (uint32_t)(b0 >> 32);
                const uint32_t a_lo0 = (uint32_t)a0, b_lo0 = (uint32_t)b0;
                const bool first_unit = u == u0;
                if constexpr (TAPMODE) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        // K = 16 pixels = tile rows 2*ks, 2*ks + 1
                        const uint32_t a_lo = a_lo0 + (((uint32_t)ks * a_kstep) >> 4);
                        const uint32_t bk_lo = b_lo0 + ((uint32_t)(ks * 2 * STRIDE * cPWp * RBX) >> 4);
#pragma unroll
                        for (int ai = 0; ai < NACC; ++ai) {
                            if (WG_DBG(1)) continue;
                            if (live & (1u << ai)) {
                                if (ks == 0 && first_unit) umma_f16_lohi<0>(tmem_base + (uint32_t)(ai * NB), a_lo, a_hi, bk_lo + boff[ai], b_hi, idesc);
                                else umma_f16_lohi<1>(tmem_base + (uint32_t)(ai * NB), a_lo, a_hi, bk_lo + boff[ai], b_hi, idesc);
                            }
                        }
                    }
                } else {
                    // 10 - 15 accumulators of N <= 32: the tap offsets are compile-time constants; unrolling the K loop as well
                    // thrashes the instruction cache (+5-10 %), and the register-array form above is slower than this one
                    const uint32_t first = first_unit ? 0u : 1u;
#pragma unroll 1
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint64_t a = a0 + (uint64_t)(((uint32_t)ks * a_kstep) >> 4);
                        const uint64_t bk = b0 + (uint64_t)((uint32_t)(ks * 2 * STRIDE * cPWp * RBX) >> 4);
                        const uint32_t acc = (ks == 0) ? first : 1u;
#pragma unroll
                        for (int ai = 0; ai < NACC; ++ai) {
                            if (WG_DBG(1)) continue;
                            const int ky = ai / NS;       // relative to this CTA's first filter row
                            const int kx = wg_shift<KS, STRIDE, SH>(ai - ky * NS);
                            const int toff = STRIDE == 1 ? ky * cPWp + kx : ky * cPWp + (kx & 1) * cPWhalf + (kx >> 1);
                            if (ky < nky)
                                umma_f16(tmem_base + (uint32_t)(ai * NB), a, bk + (uint64_t)((uint32_t)(toff * RBX) >> 4), idesc, acc);
                        }
                    }
                }
                umma_commit(bar_empty + 8 * stage);
                WG_ACC(2);          // issue
                if (++stage == p.NPS) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            if (u1 > u0) umma_commit(bar_done);
            WG_DUMP(2);
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace
}  // namespace ss

using namespace ss;

extern "C" int ss_conv_wgrad_bf16(const ss_block_desc* g, const void* x, const void* g_bf16, float* g_w, void* stream) {
    if (g == nullptr) {
        set_error("ss_conv_wgrad_bf16: null descriptor");
        return SS_EINVAL;
    }
    if (g->T == 0 || g->B == 0) return SS_OK;
    if (x == nullptr || g_bf16 == nullptr || g_w == nullptr) {
        set_error("ss_conv_wgrad_bf16: null argument");
        return SS_EINVAL;
    }
    const bool first = g->Cin >= 1 && g->Cin <= 4;
    const bool up = g->upsample != 0;
    if (g->T < 0 || g->B < 0 || g->Hin <= 0 || g->Win <= 0 || g->Hout <= 0 || g->Wout <= 0 || g->Cout <= 0 || g->Cout % 16 != 0 ||
        (!first && g->Cin % 16 != 0)) {
        set_error("ss_conv_wgrad_bf16: bad geometry (Cout %% 16, Cin %% 16 or Cin <= 4)");
        return SS_EINVAL;
    }
    if (!(g->ks == 3 || g->ks == 5) || !(g->stride == 1 || g->stride == 2) || (up && g->stride != 1) ||
        (g->stride == 2 && (g->pad % 2 != 0 || g->ks != 5)) || (first && (g->ks != 5 || g->stride != 1 || up))) {
        set_error("ss_conv_wgrad_bf16: unsupported conv shape (ks %d stride %d pad %d upsample %d)", g->ks, g->stride, g->pad,
                  g->upsample);
        return SS_EUNSUPPORTED;
    }
    WgParams p{};
    p.T = g->T; p.B = g->B; p.Hin = g->Hin; p.Win = g->Win; p.Cin = g->Cin; p.Hout = g->Hout; p.Wout = g->Wout; p.Cout = g->Cout;
    p.pad = up ? 0 : g->pad;
    p.upsample = up ? 1 : 0;
    p.Hup = g->Hout + g->ks - 1;
    p.Wup = g->Wout + g->ks - 1;
    if (up) {
        p.HsO = p.Hup;
    } else {
        const int need = g->Hin + p.pad;
        const int reach = g->stride * (g->Hout - 1) + g->ks - p.pad;
        const int span = need > reach ? need : reach;
        p.HsO = (span + g->stride - 1) / g->stride;
        if (p.HsO < g->Hout) p.HsO = g->Hout;
    }
    const long long rows = (long long)p.HsO * g->B;
    // copies of the g patch per MMA (see the kernel header): only where 128 accumulator rows would be mostly padding
    const int SH = (g->ks == 5 && g->Cout <= 32 && g->stride == 1) ? 4 : ((g->ks == 5 && g->Cout <= 64) ? 2 : 1);
    p.tiles_x = (g->Wout + (SH - 1) + 7) / 8;
    p.mtiles = (int)((rows + 15) / 16) * p.tiles_x;
    // input channels per CTA
    static int nb64_env = -1;
    if (nb64_env < 0) {
        const char* e = getenv("SS_WGRAD_N64");
        nb64_env = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    const int NB = (!first && nb64_env && g->stride == 1 && g->Cin % 64 == 0) ? 64 : ((!first && g->Cin % 32 == 0) ? 32 : 16);
    const int NS = SH == 1 ? g->ks : (g->stride == 1 ? (g->ks + SH - 1) / SH : 3);
    int NGRP;
    if (NB == 64) {
        NGRP = (g->ks * NS + 7) / 8;                       // accumulator groups (wg_tap_groups)
    } else {
        int GK = (512 / NB) / NS;                          // wg_rows_per_group
        if (GK > g->ks) GK = g->ks;
        NGRP = (g->ks + GK - 1) / GK;
    }
    p.nblk = SH > 1 ? 1 : (g->Cout + 127) / 128;
    p.nchunk = first ? 1 : g->Cin / NB;
    p.cin_real = g->Cin;
    p.x_small = g->planes != 0 ? 1 : 0;
    const long long U = (long long)p.mtiles * g->T;
    if (U > 0x7fffffffLL || (long long)g->B * g->Hin * g->Win * g->Cin > 0x7fffffffLL ||
        (long long)g->B * g->Hout * g->Wout * g->Cout > 0x7fffffffLL) {
        set_error("ss_conv_wgrad_bf16: problem too large for 32-bit indexing");
        return SS_EINVAL;
    }
    const int dev = current_device();
    const int num_sms = device_sm_count(dev);
    const int pairs = p.nblk * p.nchunk * NGRP;
    long long ns = num_sms / pairs;                        // one wave: every CTA keeps its accumulators for its whole life,
                                                           // so a second, partial wave would double the kernel's duration
    if (ns > U) ns = U;
    if (ns < 1) ns = 1;
    p.nsplit = (int)ns;
    p.yscale = up ? (float)g->Hin / (float)p.Hup : 1.0f;
    p.xscale = up ? (float)g->Win / (float)p.Wup : 1.0f;
    p.x = reinterpret_cast<const uint8_t*>(x);
    p.g = reinterpret_cast<const __nv_bfloat16*>(g_bf16);
    p.g_w = g_w;
#ifdef SS_ROLE_TIMING
    {
        const char* e = getenv("SS_WG_DBG");
        p.dbg = e != nullptr ? atoi(e) : 0;
    }
#endif
    const int PH = 15 * g->stride + g->ks;
    const int PWp = g->stride == 1 ? 8 + g->ks - 1 : 2 * (8 + (g->ks - 1) / 2);
    const int PBX = (PH * PWp * 2 * NB + 1023) / 1024 * 1024;
    const int g_bytes = SH == 1 ? G_TILE_BYTES : (16 * (8 + SH - 1) * (256 / SH) + 1023) / 1024 * 1024;
    const int stage_bytes = g_bytes + PBX;
    const int tail_bytes = 256 + (3 * WG_MAX_STAGES + 1) * 8 + 64;
    int nps = (227 * 1024 - 1024 - tail_bytes) / stage_bytes;
    if (nps > WG_MAX_STAGES) nps = WG_MAX_STAGES;
    if (nps < 2) {
        set_error("ss_conv_wgrad_bf16: not enough shared memory");
        return SS_EUNSUPPORTED;
    }
    p.NPS = nps;
    const size_t smem = 1024 + (size_t)nps * stage_bytes + tail_bytes;
    const unsigned grid = (unsigned)(pairs * p.nsplit);
    cudaStream_t st = (cudaStream_t)stream;
    bool launched = false;
#define SS_TRY_WG(KS_, ST_, NB_, F4_, SH_)                                                                                       \
    if (!launched && g->ks == KS_ && g->stride == ST_ && NB == NB_ && first == F4_ && SH == SH_) {                               \
        SS_ENSURE_SMEM((conv_wgrad_umma_kernel<KS_, ST_, NB_, F4_, SH_>), dev, 227 * 1024);                                      \
        conv_wgrad_umma_kernel<KS_, ST_, NB_, F4_, SH_><<<grid, WG_THREADS, smem, st>>>(p);                                      \
        launched = true;                                                                                                         \
    }
    SS_TRY_WG(5, 1, 64, false, 1)
    SS_TRY_WG(5, 1, 64, false, 2)
    SS_TRY_WG(5, 1, 64, false, 4)
    SS_TRY_WG(3, 1, 64, false, 1)
    SS_TRY_WG(5, 1, 32, false, 1)
    SS_TRY_WG(5, 1, 32, false, 2)
    SS_TRY_WG(5, 1, 32, false, 4)
    SS_TRY_WG(5, 1, 16, false, 1)
    SS_TRY_WG(5, 1, 16, false, 2)
    SS_TRY_WG(5, 1, 16, false, 4)
    SS_TRY_WG(5, 2, 32, false, 1)
    SS_TRY_WG(5, 2, 32, false, 2)
    SS_TRY_WG(5, 2, 16, false, 1)
    SS_TRY_WG(5, 2, 16, false, 2)
    SS_TRY_WG(3, 1, 32, false, 1)
    SS_TRY_WG(3, 1, 16, false, 1)
    SS_TRY_WG(5, 1, 16, true, 1)
    SS_TRY_WG(5, 1, 16, true, 2)
    SS_TRY_WG(5, 1, 16, true, 4)
#undef SS_TRY_WG
    if (!launched) {
        set_error("ss_conv_wgrad_bf16: no kernel instance for ks %d stride %d", g->ks, g->stride);
        return SS_EUNSUPPORTED;
    }
    count_launch();
    return check_launch("conv_wgrad_umma");
}

#ifdef SS_ROLE_TIMING
extern "C" int ss_wg_debug_read(unsigned long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, ss::ss_wg_dbg, sizeof(unsigned long long) * 3 * 8) == cudaSuccess ? 0 : -1;
}
#endif
