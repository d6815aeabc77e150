// Prediction heads + I-neuron readout (reference network/SNN_models.py:133-150,172-188).
//
// Each head is NNConvUpsampling(C -> 1, k=3, up to 260x346, bias) -> MultiplyBy -> accumulate into the
// potential of one non-firing IF pool ("I-neurons").  Cout = 1 gives the tensor cores nothing to do and the
// reference materialises a [B,C,262,348] fp32 upsampled tensor per head; here one thread owns one output
// pixel, gathers the 9 source pixels through the (ymap,xmap) tables straight from the bf16 NHWC spike
// tensors and keeps the potential in a register across heads and timesteps.  HBM-bound: the only
// compulsory traffic is the low-resolution spike tensors (L1/L2-resident across the 2x..8x replicated
// readers) plus one fp32 depth plane per head at the last timestep.
#include "ss_common.cuh"

namespace ss {
namespace {

struct HeadsParams {
    int T, B, H, W;
    float gain;
    int C[4], Hs[4], Ws[4], woff[4];
    const __nv_bfloat16* acts[4];
    const float* w[4];
    const float* bias[4];
    const int* ymap[4];
    const int* xmap[4];
    float* v_io;
    float* depths;
};

constexpr int HEADS_NT = 128;

__global__ void __launch_bounds__(HEADS_NT) heads_ineuron_kernel(const HeadsParams p) {
    extern __shared__ float wsm[];  // concatenated [9][C_i] weights of the four heads
    {
        for (int i = 0; i < 4; ++i)
            for (int j = threadIdx.x; j < 9 * p.C[i]; j += HEADS_NT) wsm[p.woff[i] + j] = __ldg(p.w[i] + j);
    }
    __syncthreads();
    const int HW = p.H * p.W;
    const long long pix = (long long)blockIdx.x * HEADS_NT + threadIdx.x;
    if (pix >= (long long)p.B * HW) return;
    const int b = (int)(pix / HW);
    const int r = (int)(pix - (long long)b * HW);
    const int y = r / p.W, x = r - (r / p.W) * p.W;

    float bias[4];
    int sy[4][3], sx[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bias[i] = __ldg(p.bias[i]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            sy[i][k] = __ldg(p.ymap[i] + y * 3 + k);
            sx[i][k] = __ldg(p.xmap[i] + x * 3 + k);
        }
    }
    float v = p.v_io[pix];
    for (int t = 0; t < p.T; ++t) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int C = p.C[i];
            const __nv_bfloat16* base = p.acts[i] + ((size_t)t * p.B + b) * p.Hs[i] * p.Ws[i] * C;
            float acc = 0.0f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    if (sy[i][ky] < 0 || sx[i][kx] < 0) continue;
                    const uint4* src = reinterpret_cast<const uint4*>(base + ((size_t)sy[i][ky] * p.Ws[i] + sx[i][kx]) * C);
                    const float* wt = wsm + p.woff[i] + (ky * 3 + kx) * C;
                    for (int c8 = 0; c8 < C / 8; ++c8) {
                        const uint4 raw = __ldg(src + c8);
                        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
                        const float4 w0 = *reinterpret_cast<const float4*>(wt + c8 * 8);
                        const float4 w1 = *reinterpret_cast<const float4*>(wt + c8 * 8 + 4);
                        const float2 f0 = __bfloat1622float2(h2[0]), f1 = __bfloat1622float2(h2[1]);
                        const float2 f2 = __bfloat1622float2(h2[2]), f3 = __bfloat1622float2(h2[3]);
                        acc = fmaf(f0.x, w0.x, acc);
                        acc = fmaf(f0.y, w0.y, acc);
                        acc = fmaf(f1.x, w0.z, acc);
                        acc = fmaf(f1.y, w0.w, acc);
                        acc = fmaf(f2.x, w1.x, acc);
                        acc = fmaf(f2.y, w1.y, acc);
                        acc = fmaf(f3.x, w1.z, acc);
                        acc = fmaf(f3.y, w1.w, acc);
                    }
                }
            }
            // (conv + bias) * gain, then IF charge with v_threshold = inf: v = v + x, never fires
            v = __fadd_rn(v, __fmul_rn(__fadd_rn(acc, bias[i]), p.gain));
            if (t == p.T - 1) p.depths[(size_t)i * p.B * HW + pix] = v;
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
    for (int i = 0; i < 4; ++i) {
        if (a->C[i] % 8 != 0 || a->C[i] <= 0) {
            set_error("ss_heads_fwd: head %d channels %d not a multiple of 8", i, a->C[i]);
            return SS_EINVAL;
        }
        p.C[i] = a->C[i]; p.Hs[i] = a->Hs[i]; p.Ws[i] = a->Ws[i];
        p.woff[i] = off;
        off += 9 * a->C[i];
        p.acts[i] = reinterpret_cast<const __nv_bfloat16*>(a->acts[i]);
        p.w[i] = a->w[i]; p.bias[i] = a->bias[i]; p.ymap[i] = a->ymap[i]; p.xmap[i] = a->xmap[i];
    }
    p.v_io = v_io; p.depths = depths;
    const long long npix = (long long)p.B * p.H * p.W;
    if (npix == 0 || p.T == 0) return SS_OK;
    const size_t smem = (size_t)off * sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(heads_ineuron_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    heads_ineuron_kernel<<<(unsigned)((npix + HEADS_NT - 1) / HEADS_NT), HEADS_NT, smem, (cudaStream_t)stream>>>(p);
    count_launch();
    return check_launch("heads_ineuron");
}
