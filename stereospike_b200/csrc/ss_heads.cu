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
//   kernel 1  head_taps_kernel   : u8 NHWC spikes -> fp32 taps[e][b][tap][sy][sx]      (one launch, all heads)
//   kernel 2  heads_gather_kernel: one thread per output pixel, potential in a register across heads and
//                                  timesteps, 36 coalesced gathers per timestep, depth planes written at the
//                                  last timestep.  HBM/L2-bound.
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
    head_taps_kernel<<<(unsigned)((p.pix_begin[4] + TAPS_NT - 1) / TAPS_NT), TAPS_NT, smem, st>>>(p);
    count_launch();
    if (check_launch("head_taps") != SS_OK) return SS_ECUDA;
    const unsigned gblocks = (unsigned)((npix + GATHER_NT - 1) / GATHER_NT);
    if (p.NE == 2) heads_gather_kernel<2><<<gblocks, GATHER_NT, 0, st>>>(p);
    else heads_gather_kernel<0><<<gblocks, GATHER_NT, 0, st>>>(p);
    count_launch();
    return check_launch("heads_gather");
}
