// fp32 CUDA-core implementation of the fused spiking block (implicit GEMM + neuron over T).
//
// This is the exact-fp32 path: activations are small integers (exact), weights stay fp32, products are
// accumulated in fp32 in ascending k order.  It serves the first layer (Cin = 2 or 4, K = 50 / 100, HBM
// bound, reads the reference's fp32 NCHW frames directly) and is the on-device cross-check for the tcgen05
// path.  Replaces: Conv2d/NNConvUpsampling -> MultiplyBy -> IF/LIF/PLIF node per timestep
// (reference network/SNN_models.py:75-129, network/blocks.py:110-132,145-171).
#include "ss_common.cuh"

namespace ss {

namespace {

constexpr int BM = 64;   // output pixels per CTA
constexpr int BK = 32;   // k elements per smem stage
constexpr int NT = 256;  // threads per CTA
constexpr int APAD = 4;

template <int BN, int IN_LAYOUT>
__global__ void __launch_bounds__(NT) conv_neuron_simt_kernel(const ConvParams p) {
    constexpr int TN = BN / 16;  // output channels per thread
    __shared__ __align__(16) float As[BK][BM + APAD];
    __shared__ __align__(16) float Bs[BK][BN];
    __shared__ int row_img[BM];  // b index, or -1 for rows past M
    __shared__ int row_oy[BM];
    __shared__ int row_ox[BM];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int HW = p.Hout * p.Wout;

    if (tid < BM) {
        const int m = m0 + tid;
        if (m < p.M) {
            const int b = m / HW;
            const int r = m - b * HW;
            row_img[tid] = b;
            row_oy[tid] = r / p.Wout;
            row_ox[tid] = r - (r / p.Wout) * p.Wout;
        } else {
            row_img[tid] = -1;
            row_oy[tid] = 0;
            row_ox[tid] = 0;
        }
    }
    __syncthreads();

    const int tx = tid & 15;   // n group
    const int ty = tid >> 4;   // m group
    const int lrow = tid & 63; // loader row
    const int lchk = tid >> 6; // loader chunk (8 k elements)
    const int l_b = row_img[lrow], l_oy = row_oy[lrow], l_ox = row_ox[lrow];

    float decay = 0.0f;
    if (p.neuron == SS_NEURON_PLIF) decay = __ldg(p.decay);

    float v[4][TN];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            v[i][j] = (p.v_in != nullptr && m < p.M) ? p.v_in[(size_t)m * p.Cout + n0 + tx * TN + j] : p.v_reset;
        }
    }

    const int nkb = (p.K + BK - 1) / BK;
    for (int t = 0; t < p.T; ++t) {
        float acc[4][TN];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

        for (int kb = 0; kb < nkb; ++kb) {
            // ---- stage A: gather 8 k elements of one output pixel
            {
                float a[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] = 0.0f;
                const int k0 = kb * BK + lchk * 8;
                if (l_b >= 0) {
                    if (IN_LAYOUT == SS_IN_U8_TBHWC) {
                        // Cin % 8 == 0: the 8 elements share one tap
                        const int tap = k0 / p.Cin;
                        const int c = k0 - tap * p.Cin;
                        if (k0 < p.K) {
                            const int ky = tap / p.ks, kx = tap - ky * p.ks;
                            const int sy = __ldg(p.ymap + l_oy * p.ks + ky);
                            const int sx = __ldg(p.xmap + l_ox * p.ks + kx);
                            if (sy >= 0 && sx >= 0) {
                                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.x) +
                                    ((((size_t)t * p.B + l_b) * p.Hin + sy) * p.Win + sx) * p.Cin + c;
                                const uint2 raw = __ldg(reinterpret_cast<const uint2*>(src));
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    a[e] = (float)((raw.x >> (8 * e)) & 0xFFu);
                                    a[4 + e] = (float)((raw.y >> (8 * e)) & 0xFFu);
                                }
                            }
                        }
                    } else {
                        const float* xin = reinterpret_cast<const float*>(p.x);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int k = k0 + e;
                            if (k < p.K) {
                                const int tap = k / p.Cin;
                                const int c = k - tap * p.Cin;
                                const int ky = tap / p.ks, kx = tap - ky * p.ks;
                                const int sy = __ldg(p.ymap + l_oy * p.ks + ky);
                                const int sx = __ldg(p.xmap + l_ox * p.ks + kx);
                                if (sy >= 0 && sx >= 0)
                                    a[e] = __ldg(xin + ((((size_t)l_b * p.T + t) * p.Cin + c) * p.Hin + sy) * p.Win + sx);
                            }
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) As[lchk * 8 + e][lrow] = a[e];
            }
            // ---- stage B: weights [K][Cout]
            {
                constexpr int V4 = BK * BN / 4;
                for (int idx = tid; idx < V4; idx += NT) {
                    const int kk = idx / (BN / 4);
                    const int n4 = idx - kk * (BN / 4);
                    const int k = kb * BK + kk;
                    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (k < p.K) w4 = __ldg(reinterpret_cast<const float4*>(p.w_kn + (size_t)k * p.Cout + n0 + n4 * 4));
                    *reinterpret_cast<float4*>(&Bs[kk][n4 * 4]) = w4;
                }
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < BK; ++kk) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                const float av[4] = {a4.x, a4.y, a4.z, a4.w};
                float bv[TN];
#pragma unroll
                for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }

        // ---- epilogue: gain -> charge -> fire -> reset -> (+ residual) -> bf16 NHWC
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty * 4 + i;
            if (m >= p.M) continue;
            const size_t o = ((size_t)t * p.M + m) * p.Cout + n0 + tx * TN;
            float s[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                float h;
                s[j] = neuron_step(p.neuron, __fmul_rn(acc[i][j], p.gain), v[i][j], p.v_th, p.v_reset, p.tau, decay, h);
                if (p.h_seq != nullptr) p.h_seq[o + j] = h;
            }
            if (p.resid != nullptr) {
#pragma unroll
                for (int j = 0; j < TN; ++j) s[j] += (float)p.resid[o + j];
            }
#pragma unroll
            for (int j = 0; j < TN; j += 2)
                *reinterpret_cast<uchar2*>(p.out + o + j) = make_uchar2((unsigned char)s[j], (unsigned char)s[j + 1]);
        }
    }

    if (p.v_out != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty * 4 + i;
            if (m >= p.M) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) p.v_out[(size_t)m * p.Cout + n0 + tx * TN + j] = v[i][j];
        }
    }
}

template <int BN>
int launch_bn(const ConvParams& p, int in_layout, cudaStream_t st) {
    dim3 grid((p.M + BM - 1) / BM, p.Cout / BN);
    if (in_layout == SS_IN_U8_TBHWC)
        conv_neuron_simt_kernel<BN, SS_IN_U8_TBHWC><<<grid, NT, 0, st>>>(p);
    else
        conv_neuron_simt_kernel<BN, SS_IN_F32_BTCHW><<<grid, NT, 0, st>>>(p);
    count_launch();
    return check_launch("conv_neuron_simt");
}

}  // namespace

int launch_conv_neuron_simt(const ConvParams& p, int in_layout, cudaStream_t st) {
    if (in_layout == SS_IN_U8_TBHWC && (p.Cin % 8) != 0) {
        set_error("simt: u8 input needs Cin %% 8 == 0 (got %d)", p.Cin);
        return SS_EINVAL;
    }
    if (p.Cout % 32 != 0) {
        set_error("simt: Cout %% 32 != 0 (got %d)", p.Cout);
        return SS_EINVAL;
    }
    if (p.Cout % 64 == 0) return launch_bn<64>(p, in_layout, st);
    return launch_bn<32>(p, in_layout, st);
}

}  // namespace ss
