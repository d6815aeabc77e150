// Prediction heads + I-neuron readout (reference network/SNN_models.py:133-150,172-188).
//
// Each head is NNConvUpsampling(C -> 1, k=3, up to 260x346, bias) -> MultiplyBy -> accumulate into the
// potential of one non-firing IF pool ("I-neurons").  Cout = 1 gives the tensor cores nothing to do and the
// reference materialises a [B,C,262,348] fp32 upsampled tensor per head.  Here the work is restructured
// (SURVEY.md section 8(a) row 9): because nearest-neighbour upsampling only REPLICATES source pixels, the nine
// per-tap channel dots  P[tap][sy][sx] = sum_c act[sy][sx][c] * w[tap][c]  are taken once per SOURCE pixel
// (9*Hs*Ws*C MACs instead of 9*H*W*C: 8x fewer over the four heads), and every output pixel then just gathers
// 9 of those scalars per head.  Same math up to fp32 reassociation (tap-major instead of interleaved).
// The readout is LINEAR and never fires, so the time loop can be folded as well: sum_t conv(act_t) = conv(sum_t act_t).
// When the producing blocks hand over  acts_sum = sum_{t < T-1} act_t  (u8, accumulated for free in their epilogues) the
// heads are evaluated on two pseudo-timesteps only -- the summed past (bias counted T-1 times) and the last step, which
// the four running-sum outputs need separately -- instead of T.  Again the same math up to fp32 reassociation.
//   kernel 1  head_taps_mma_kernel: u8 NHWC spikes -> fp32 taps[e][b][tap][sy][sx]     (one launch, all heads) as warp-level
//                                  bf16 MMAs with the fp32 weights split exactly into three bf16 pieces; head_taps_kernel is
//                                  the CUDA-core version for channel counts the MMA tiling does not cover
//   kernel 2  heads_gather_kernel: one thread per output pixel, potential in a register across heads and
//                                  timesteps, 36 coalesced gathers per timestep (all issued before the first add), depth
//                                  planes written at the last timestep.  HBM/L2-bound.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "ss_common.cuh"

namespace ss {
namespace {

struct HeadsParams {
    int T, B, H, W;
    float gain;
    int C[4], Hs[4], Ws[4], woff[4];
    int NE;                   // evaluated (pseudo-)timesteps: T, or 2 when acts_sum is given and T > 1
    int summed;               // 1: step 0 = acts_sum (first T-1 steps), step 1 = last timestep of acts
    float bias_mul[2];        // summed mode: {T-1, 1}
    long long pix_begin[5];   // prefix sums of NE*B*Hs*Ws over the heads (kernel 1 work partition)
    int blk_begin[5];         // head_taps_mma_kernel: prefix sums of the blocks (TM_PIX pixels each) of the heads
    int bt[4];                // head_taps_mma_kernel: bytes a thread loads per pixel and channel group (4 / 8 / 16)
    const uint8_t* acts[4];
    const uint8_t* acts_sum[4];
    const float* w[4];
    const float* bias[4];
    const int* ymap[4];
    const int* xmap[4];
    float* taps[4];
    float* v_io;
    float* depths;
};

constexpr int TAPS_NT = 128;

// one thread per (head, t, b, sy, sx): nine dots over the channels
__global__ void __launch_bounds__(TAPS_NT) head_taps_kernel(const HeadsParams p) {
    extern __shared__ float wsm[];  // concatenated [9][C_i] weights of the four heads
    // only the heads this block's pixels belong to (pixels are ordered by head, so that is one head for all but three blocks):
    // loading all four weight sets (17 KB) per 128 pixels cost as much as the dots themselves
    const long long g_first = (long long)blockIdx.x * TAPS_NT;
    const long long g_last = min(g_first + TAPS_NT, p.pix_begin[4]) - 1;
    for (int i = 0; i < 4; ++i) {
        if (g_last < p.pix_begin[i] || g_first >= p.pix_begin[i + 1]) continue;
        for (int j = threadIdx.x; j < 9 * p.C[i]; j += TAPS_NT) wsm[p.woff[i] + j] = __ldg(p.w[i] + j);
    }
    __syncthreads();
    const long long gid = g_first + threadIdx.x;
    if (gid >= p.pix_begin[4]) return;
    int hd = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (gid >= p.pix_begin[i]) hd = i;
    const long long q = gid - p.pix_begin[hd];          // (e*B + b)*Hs*Ws + s
    const int C = p.C[hd];
    const int S = p.Hs[hd] * p.Ws[hd];
    const long long tb = q / S;
    const int s = (int)(q - tb * S);
    const uint8_t* a0 = p.acts[hd] + (size_t)q * C;
    if (p.summed) {
        const long long e = tb / p.B, b = tb - e * p.B;
        a0 = (e == 0) ? p.acts_sum[hd] + ((size_t)b * S + s) * C
                      : p.acts[hd] + (((size_t)(p.T - 1) * p.B + b) * S + s) * C;
    }
    const uint4* src = reinterpret_cast<const uint4*>(a0);
    const float* wt = wsm + p.woff[hd];
    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0f;
    for (int c16 = 0; c16 < C / 16; ++c16) {
        const uint4 raw = __ldg(src + c16);
        const uint32_t wds[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float a0 = (float)(wds[q] & 0xFFu), a1 = (float)((wds[q] >> 8) & 0xFFu);
            const float a2 = (float)((wds[q] >> 16) & 0xFFu), a3 = (float)(wds[q] >> 24);
            const float* wc = wt + c16 * 16 + q * 4;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float4 w4 = *reinterpret_cast<const float4*>(wc + k * C);
                acc[k] = fmaf(a3, w4.w, fmaf(a2, w4.z, fmaf(a1, w4.y, fmaf(a0, w4.x, acc[k]))));
            }
        }
    }
    float* dst = p.taps[hd] + (size_t)tb * 9 * S + s;
#pragma unroll
    for (int k = 0; k < 9; ++k) dst[(size_t)k * S] = acc[k];
}

// ---- the same nine dots on the tensor cores (warp-level MMA: the tile is 16 pixels x 8 taps, far too small for tcgen05) ----
// The CUDA-core kernel above is bound by instruction issue (14 instructions per pixel and 4 channels, 0.08 ms per forward).
// Here a warp takes 32 source pixels: A = their u8 activations widened to bf16 (integers up to 255 are exact), B = the fp32
// head weights split into three bf16 pieces hi + mid + lo (8 + 8 + 8 significand bits: the split is EXACT), all products are
// exact in fp32 and accumulate in the MMA's fp32 accumulators -- the fp32 dot product of the kernel above up to the order of
// the additions.  The k index of an MMA is a dummy, so each thread's four k slots of a 16-channel step are mapped to four
// CONSECUTIVE channel bytes (one 32-bit word of the pixel), and a thread loads 8 or 16 contiguous bytes per pixel:
//   k-step s of a 64-channel group, thread tg (lane & 3), slot j  <->  channel  64*group + BT*tg + 4*s + j     (BT = bytes/thread)
// The B fragments are built once per block in shared memory, already in per-lane register order:
//   wfrag[k-step][2][lane][4 words]: (lo, mid) and (hi, tap 8) -- two conflict-free 16-byte loads per lane and k-step.  Taps 0..7
// are the columns of n-tile 0, accumulated lo, mid, hi; tap 8 has an n-tile of its own whose columns 0..2 are its three pieces
// (one MMA instead of three), summed at the end.
constexpr int TM_NT = 128;          // 4 warps
constexpr int TM_ITERS = 4;         // 32-pixel tiles per warp: amortises the weight-fragment build
constexpr int TM_PIX = TM_NT * TM_ITERS;

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// BT: bytes of one pixel a thread loads per channel group (4: C = 16, 8: C = 32, 16: C a multiple of 64)
template <int BT>
__device__ __forceinline__ void head_taps_mma_body(const HeadsParams& p, const int hd, uint32_t* wfrag) {
    const int C = p.C[hd];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, tg = lane & 3;
    constexpr int CG = BT * 4;                       // channels per group
    constexpr int KS = CG / 16;                      // k-steps per group
    const int ngroups = C / CG;
    // ---- weight fragments: entry (k-step, lane) = the lane's 4 channels x {tap g: lo, mid, hi; n-tile 1: piece g of tap 8}
    const int nks = ngroups * KS;
    for (int ent = threadIdx.x; ent < nks * 32; ent += TM_NT) {
        const int el = ent & 31, ks = ent >> 5;
        const int eg = el >> 2, etg = el & 3;
        const int ch = (ks / KS) * CG + BT * etg + 4 * (ks % KS);
        auto split = [](const float4 w4, float (&pc)[3][4]) {          // pc[lo, mid, hi][j]
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float hi = __bfloat162float(__float2bfloat16_rn(wv[j]));
                const float r1 = wv[j] - hi;                                   // exact
                const float mid = __bfloat162float(__float2bfloat16_rn(r1));
                pc[2][j] = hi; pc[1][j] = mid; pc[0][j] = r1 - mid;           // exact, and fits 8 bits
            }
        };
        float pc[3][4], p8[3][4];
        split(__ldg(reinterpret_cast<const float4*>(p.w[hd] + (size_t)eg * C + ch)), pc);
        split(__ldg(reinterpret_cast<const float4*>(p.w[hd] + (size_t)8 * C + ch)), p8);
        const int q8 = eg < 3 ? eg : 0;
        const uint32_t t0 = eg < 3 ? pack_bf16(p8[q8][0], p8[q8][1]) : 0u, t1 = eg < 3 ? pack_bf16(p8[q8][2], p8[q8][3]) : 0u;
        uint4* dst = reinterpret_cast<uint4*>(wfrag);
        dst[(size_t)(ks * 2 + 0) * 32 + el] = make_uint4(pack_bf16(pc[0][0], pc[0][1]), pack_bf16(pc[0][2], pc[0][3]),
                                                         pack_bf16(pc[1][0], pc[1][1]), pack_bf16(pc[1][2], pc[1][3]));
        dst[(size_t)(ks * 2 + 1) * 32 + el] = make_uint4(pack_bf16(pc[2][0], pc[2][1]), pack_bf16(pc[2][2], pc[2][3]), t0, t1);
    }
    __syncthreads();
    const int S = p.Hs[hd] * p.Ws[hd];
    const int Q = p.NE * p.B * S;                    // pixels of this head (< 2^31, checked by the launcher): q = (e*B + b)*S + s
    const int qblk = ((int)blockIdx.x - p.blk_begin[hd]) * TM_PIX;
    // source bytes of pixel q: linear in q -- acts + q*C, or in the summed mode acts_sum + q*C for the first B*S pixels (e = 0)
    // and the last timestep of acts for the rest (e = 1)
    const int BS = p.B * S;
    const uint8_t* base0 = p.summed ? p.acts_sum[hd] : p.acts[hd];
    const uint8_t* base1 = p.summed ? p.acts[hd] + (size_t)(p.T - 2) * BS * C : p.acts[hd];
    const int split = p.summed ? BS : 0;
    // (a software pipeline over tiles and channel groups -- next loads issued before the current MMAs -- was measured: no gain,
    // 128 instead of 89 registers)
    for (int it = 0; it < TM_ITERS; ++it) {
        const int q0 = qblk + (it * (TM_NT / 32) + warp) * 32;
        if (q0 >= Q) break;
        const int tb0 = q0 / S, s0 = q0 - tb0 * S;       // one division per tile: the rows follow with carries
        // rows of this thread: m-tile mt, half h -> pixel q = q0 + 16*mt + 8*h + g (addresses are recomputed from q where they are
        // needed: registers decide how many warps hide the load latency here)
        int tbr[2][2];                                   // q / S of the row
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int sp = s0 + 16 * mt + 8 * h + g, tb = tb0;
                while (sp >= S) {
                    sp -= S;
                    ++tb;
                }
                tbr[mt][h] = tb;
            }
        float acc[2][2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.0f;
        for (int grp = 0; grp < ngroups; ++grp) {
            uint32_t raw[2][2][KS];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int k = 0; k < KS; ++k) raw[mt][h][k] = 0u;
                    const int q = q0 + 16 * mt + 8 * h + g;
                    if (q < Q) {
                        const uint8_t* a = (q < split ? base0 : base1) + (size_t)q * C + (grp * CG + BT * tg);
                        if constexpr (BT == 16) {
                            const uint4 v = __ldg(reinterpret_cast<const uint4*>(a));
                            raw[mt][h][0] = v.x; raw[mt][h][1] = v.y; raw[mt][h][2] = v.z; raw[mt][h][KS - 1] = v.w;
                        } else if constexpr (BT == 8) {
                            const uint2 v = __ldg(reinterpret_cast<const uint2*>(a));
                            raw[mt][h][0] = v.x; raw[mt][h][KS - 1] = v.y;
                        } else {
                            raw[mt][h][0] = __ldg(reinterpret_cast<const uint32_t*>(a));
                        }
                    }
                }
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const uint4* wf = reinterpret_cast<const uint4*>(wfrag) + (size_t)(grp * KS + k) * 64 + lane;
                const uint4 b_lm = wf[0], b_h8 = wf[32];        // (lo.r0, lo.r1, mid.r0, mid.r1), (hi.r0, hi.r1, tap8.r0, tap8.r1)
                uint32_t afr[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t wd = raw[mt][h][k];
                        afr[mt][h] = pack_bf16((float)(wd & 0xFFu), (float)((wd >> 8) & 0xFFu));          // a0 / a1: slots j = 0, 1
                        afr[mt][2 + h] = pack_bf16((float)((wd >> 16) & 0xFFu), (float)(wd >> 24));        // a2 / a3: slots j = 2, 3
                    }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma_bf16_16816(acc[mt][0], afr[mt], b_lm.x, b_lm.y);      // smallest pieces first
                    mma_bf16_16816(acc[mt][0], afr[mt], b_lm.z, b_lm.w);
                    mma_bf16_16816(acc[mt][0], afr[mt], b_h8.x, b_h8.y);
                    mma_bf16_16816(acc[mt][1], afr[mt], b_h8.z, b_h8.w);      // tap 8: its three pieces are columns 0..2
                }
            }
        }
        // accumulator (mt, 0): c0, c1 = (pixel row g, taps 2*tg, 2*tg + 1), c2, c3 = (pixel row g + 8, same taps);
        // accumulator (mt, 1): columns 0, 1 (thread tg = 0) and 2 (tg = 1: c0 / c2) = the lo, mid and hi sums of tap 8
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float hi8 = __shfl_down_sync(0xffffffffu, acc[mt][1][2 * h], 1);
                const int q = q0 + 16 * mt + 8 * h + g;
                if (q >= Q) continue;
                // tap 0 of the pixel sits at tb*9*S + s = q + tb*8*S
                float* dst = p.taps[hd] + ((long long)tbr[mt][h] * 8 * S + q) + (size_t)(2 * tg) * S;
                dst[0] = acc[mt][0][2 * h];
                dst[S] = acc[mt][0][2 * h + 1];
                if (tg == 0) dst[(size_t)8 * S] = (acc[mt][1][2 * h] + acc[mt][1][2 * h + 1]) + hi8;
            }
    }
}

__global__ void __launch_bounds__(TM_NT) head_taps_mma_kernel(const HeadsParams p) {
    extern __shared__ __align__(16) uint32_t wfrag_sm[];
    int hd = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if ((int)blockIdx.x >= p.blk_begin[i]) hd = i;
    if (p.bt[hd] == 16) head_taps_mma_body<16>(p, hd, wfrag_sm);
    else if (p.bt[hd] == 8) head_taps_mma_body<8>(p, hd, wfrag_sm);
    else head_taps_mma_body<4>(p, hd, wfrag_sm);
}

constexpr int GATHER_NT = 128;

// NE_T = 2: the summed mode (two pseudo-timesteps), fully unrolled; NE_T = 0: any number of timesteps, one at a time.
// All tap loads of a timestep (of both pseudo-timesteps when NE_T = 2) are issued before the first add: written as
// `if (off >= 0) acc += load` every add waited for its own load and the kernel ran at one L2 round trip per tap.
template <int NE_T>
__global__ void __launch_bounds__(GATHER_NT) heads_gather_kernel(const HeadsParams p) {
    const int HW = p.H * p.W;
    const long long pix = (long long)blockIdx.x * GATHER_NT + threadIdx.x;
    if (pix >= (long long)p.B * HW) return;
    const int b = (int)(pix / HW);
    const int r = (int)(pix - (long long)b * HW);
    const int y = r / p.W, x = r - (r / p.W) * p.W;

    float bias[4];
    int off[4][9];   // offset of each tap inside one [9][Hs][Ws] block, or -1
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bias[i] = __ldg(p.bias[i]);
        const int S = p.Hs[i] * p.Ws[i];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int sy = __ldg(p.ymap[i] + y * 3 + ky);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int sx = __ldg(p.xmap[i] + x * 3 + kx);
                off[i][ky * 3 + kx] = (sy >= 0 && sx >= 0) ? (ky * 3 + kx) * S + sy * p.Ws[i] + sx : -1;
            }
        }
    }
    float v = p.v_io[pix];
    constexpr int EU = NE_T > 0 ? NE_T : 1;          // timesteps whose loads are in flight together
    for (int e0 = 0; e0 < p.NE; e0 += EU) {
        float tv[EU][4][9];
#pragma unroll
        for (int eu = 0; eu < EU; ++eu)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float* base = p.taps[i] + ((size_t)(e0 + eu) * p.B + b) * 9 * p.Hs[i] * p.Ws[i];
#pragma unroll
                for (int k = 0; k < 9; ++k) tv[eu][i][k] = off[i][k] >= 0 ? __ldg(base + off[i][k]) : 0.0f;
            }
#pragma unroll
        for (int eu = 0; eu < EU; ++eu) {
            const int e = e0 + eu;
            const float bm = p.summed ? p.bias_mul[e] : 1.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < 9; ++k) acc += tv[eu][i][k];      // a padding tap adds +0: same sum as skipping it
                // (conv + bias) * gain, then IF charge with v_threshold = inf: v = v + x, never fires
                v = __fadd_rn(v, __fmul_rn(fmaf(bm, bias[i], acc), p.gain));
                if (e == p.NE - 1) p.depths[(size_t)i * p.B * HW + pix] = v;
            }
        }
    }
    p.v_io[pix] = v;
}

}  // namespace
}  // namespace ss

extern "C" int ss_heads_fwd(const ss_heads_args* a, float* v_io, float* depths, void* stream) {
    using namespace ss;
    if (a == nullptr || v_io == nullptr || depths == nullptr) {
        set_error("ss_heads_fwd: null argument");
        return SS_EINVAL;
    }
    HeadsParams p;
    p.T = a->T; p.B = a->B; p.H = a->H; p.W = a->W; p.gain = a->gain;
    int off = 0;
    p.pix_begin[0] = 0;
    p.summed = (a->T > 1 && a->acts_sum[0] != nullptr) ? 1 : 0;
    p.NE = p.summed ? 2 : a->T;
    p.bias_mul[0] = (float)(a->T - 1);
    p.bias_mul[1] = 1.0f;
    for (int i = 0; i < 4; ++i) {
        if (p.summed && a->acts_sum[i] == nullptr) {
            set_error("ss_heads_fwd: acts_sum must be given for all four heads or for none");
            return SS_EINVAL;
        }
        p.acts_sum[i] = reinterpret_cast<const uint8_t*>(a->acts_sum[i]);
        if (a->C[i] % 16 != 0 || a->C[i] <= 0) {
            set_error("ss_heads_fwd: head %d channels %d not a multiple of 16", i, a->C[i]);
            return SS_EINVAL;
        }
        if (a->taps[i] == nullptr && a->T > 0 && a->B > 0) {
            set_error("ss_heads_fwd: head %d needs its taps workspace", i);
            return SS_EINVAL;
        }
        p.C[i] = a->C[i]; p.Hs[i] = a->Hs[i]; p.Ws[i] = a->Ws[i];
        p.woff[i] = off;
        off += 9 * a->C[i];
        p.pix_begin[i + 1] = p.pix_begin[i] + (long long)p.NE * a->B * a->Hs[i] * a->Ws[i];
        p.acts[i] = reinterpret_cast<const uint8_t*>(a->acts[i]);
        p.w[i] = a->w[i]; p.bias[i] = a->bias[i]; p.ymap[i] = a->ymap[i]; p.xmap[i] = a->xmap[i];
        p.taps[i] = a->taps[i];
    }
    p.v_io = v_io; p.depths = depths;
    const long long npix = (long long)p.B * p.H * p.W;
    if (npix == 0 || p.T == 0) return SS_OK;
    const size_t smem = (size_t)off * sizeof(float);
    if (smem > 200 * 1024) {
        set_error("ss_heads_fwd: head weights do not fit in shared memory");
        return SS_EUNSUPPORTED;
    }
    if (smem > 48 * 1024) cudaFuncSetAttribute(head_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaStream_t st = (cudaStream_t)stream;
    // tensor-core taps when every head has 16, 32 or a multiple of 64 channels (SS_HEAD_TAPS_MMA=0: the CUDA-core kernel)
    static const bool mma_enabled = [] {
        const char* e = getenv("SS_HEAD_TAPS_MMA");
        return e == nullptr || atoi(e) != 0;
    }();
    bool use_mma = mma_enabled;
    size_t smem_mma = 0;
    p.blk_begin[0] = 0;
    for (int i = 0; i < 4; ++i) {
        p.bt[i] = p.C[i] == 16 ? 4 : p.C[i] == 32 ? 8 : p.C[i] % 64 == 0 ? 16 : 0;
        const long long q = p.pix_begin[i + 1] - p.pix_begin[i];
        if (p.bt[i] == 0 || q + TM_PIX >= (1LL << 31)) use_mma = false;
        p.blk_begin[i + 1] = p.blk_begin[i] + (int)((q + TM_PIX - 1) / TM_PIX);
        smem_mma = std::max(smem_mma, (size_t)p.C[i] * 64);       // C/16 k-steps x 32 lanes x 32 bytes
    }
    if (use_mma && smem_mma <= 200 * 1024) {
        if (smem_mma > 48 * 1024)
            cudaFuncSetAttribute(head_taps_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma);
        head_taps_mma_kernel<<<(unsigned)p.blk_begin[4], TM_NT, smem_mma, st>>>(p);
    } else {
        head_taps_kernel<<<(unsigned)((p.pix_begin[4] + TAPS_NT - 1) / TAPS_NT), TAPS_NT, smem, st>>>(p);
    }
    count_launch();
    if (check_launch("head_taps") != SS_OK) return SS_ECUDA;
    const unsigned gblocks = (unsigned)((npix + GATHER_NT - 1) / GATHER_NT);
    if (p.NE == 2) heads_gather_kernel<2><<<gblocks, GATHER_NT, 0, st>>>(p);
    else heads_gather_kernel<0><<<gblocks, GATHER_NT, 0, st>>>(p);
    count_launch();
    return check_launch("heads_gather");
}
