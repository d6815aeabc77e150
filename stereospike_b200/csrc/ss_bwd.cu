// Backward of the fused spiking block and of the prediction heads (fp32 CUDA-core kernels).
//
// Replaces PyTorch autograd through the reference path: the SpikingJelly surrogate
// autograd.Function.backward (ATan / Sigmoid evaluated at h - v_th, detach_reset=True), cuDNN dgrad / wgrad of
// every Conv2d, and upsample_nearest2d_backward -- the (ymap, xmap) tables fold the upsampling into the conv
// gradient, so no upsampled gradient tensor is ever materialised.
#include <cuda_bf16.h>

#include "ss_common.cuh"

namespace ss {
namespace {

// ------------------------------------------------------------------------------------------ neuron BPTT scan
__device__ __forceinline__ float surrogate_grad(int kind, float alpha, float u) {
    if (kind == SS_SURR_ATAN) {
        const float q = 1.5707963267948966f * alpha * u;
        return alpha * 0.5f / (1.0f + q * q);
    }
    const float sg = 1.0f / (1.0f + __expf(-alpha * u));
    return alpha * sg * (1.0f - sg);
}

// One thread scans VEC consecutive neurons (VEC = 4: 16-byte loads of h and g_s, 8-byte bf16 stores) backwards in time.
constexpr int SCAN_TC = 5;   // timesteps of the reverse scan whose loads are in flight together
template <int VEC>
__global__ void __launch_bounds__(256, 3) neuron_bwd_kernel(int T, long long N, int neuron, int surrogate, float alpha,
                                                         float gain, float v_th, float v_reset, float tau,
                                                         const float* __restrict__ decay_p, const float* __restrict__ h_seq,
                                                         const float* __restrict__ v_init, const float* g_s,
                                                         const float* __restrict__ g_v_last, float* g_acc,
                                                         __nv_bfloat16* __restrict__ g_acc_bf16, float* __restrict__ g_v_init,
                                                         float* __restrict__ g_decay) {
    const long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    float gd_local = 0.0f;
    if (n < N) {
        float r = 1.0f;
        if (neuron == SS_NEURON_LIF) r = 1.0f / tau;
        if (neuron == SS_NEURON_PLIF) r = __ldg(decay_p);
        const float keep = (neuron == SS_NEURON_IF) ? 1.0f : 1.0f - r;
        float g_v[VEC], h[VEC], h_prev[VEC];
        auto load = [&](const float* src, float (&dst)[VEC]) {
            if constexpr (VEC == 4) {
                const float4 q = *reinterpret_cast<const float4*>(src);
                dst[0] = q.x; dst[1] = q.y; dst[2] = q.z; dst[VEC - 1] = q.w;
            } else {
                dst[0] = *src;
            }
        };
#pragma unroll
        for (int i = 0; i < VEC; ++i) g_v[i] = 0.0f;
        if (g_v_last != nullptr) load(g_v_last + n, g_v);
        load(h_seq + (size_t)(T - 1) * N + n, h);
        // The recurrence runs backwards in time, but its inputs do not depend on it: the loads of SCAN_TC timesteps are issued
        // together (one load pair per step left the kernel waiting on HBM latency with 32 bytes in flight per thread).
        for (int t1 = T - 1; t1 >= 0; t1 -= SCAN_TC) {
            float hp[SCAN_TC][VEC], gsv[SCAN_TC][VEC];
#pragma unroll
            for (int j = 0; j < SCAN_TC; ++j) {
                const int t = t1 - j;
                if (t > 0) load(h_seq + (size_t)(t - 1) * N + n, hp[j]);
                if (t >= 0) load(g_s + (size_t)t * N + n, gsv[j]);
            }
#pragma unroll
            for (int j = 0; j < SCAN_TC; ++j) {
                const int t = t1 - j;
                if (t < 0) break;
                // potential before this step: reset(h_{t-1}) or the initial state
                float v_prev[VEC];
                if (t > 0) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        h_prev[i] = hp[j][i];
                        v_prev[i] = (h_prev[i] - v_th >= 0.0f) ? v_reset : h_prev[i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        h_prev[i] = 0.0f;
                        v_prev[i] = v_reset;
                    }
                    if (v_init != nullptr) load(v_init + n, v_prev);
                }
                float gx[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float u = h[i] - v_th;
                    const float sp = (u >= 0.0f) ? 1.0f : 0.0f;
                    const float g_h = gsv[j][i] * surrogate_grad(surrogate, alpha, u) + g_v[i] * (1.0f - sp);
                    gx[i] = ((neuron == SS_NEURON_IF) ? g_h : g_h * r) * gain;
                    if (neuron == SS_NEURON_PLIF) gd_local += g_h * ((h[i] - v_prev[i]) / r);  // d h / d r = x - (v - v_reset)
                    g_v[i] = g_h * keep;
                    h[i] = h_prev[i];
                }
                if constexpr (VEC == 4) {
                    if (g_acc != nullptr) *reinterpret_cast<float4*>(g_acc + (size_t)t * N + n) = make_float4(gx[0], gx[1], gx[2], gx[VEC - 1]);
                    if (g_acc_bf16 != nullptr) {
                        const __nv_bfloat162 lo = __floats2bfloat162_rn(gx[0], gx[1]);
                        const __nv_bfloat162 hi = __floats2bfloat162_rn(gx[2], gx[VEC - 1]);
                        uint2 pk;
                        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                        *reinterpret_cast<uint2*>(g_acc_bf16 + (size_t)t * N + n) = pk;
                    }
                } else {
                    if (g_acc != nullptr) g_acc[(size_t)t * N + n] = gx[0];
                    if (g_acc_bf16 != nullptr) g_acc_bf16[(size_t)t * N + n] = __float2bfloat16_rn(gx[0]);
                }
            }
        }
        if (g_v_init != nullptr) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) g_v_init[n + i] = g_v[i];
        }
    }
    if (neuron == SS_NEURON_PLIF && g_decay != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gd_local += __shfl_xor_sync(0xffffffffu, gd_local, o);
        __shared__ float red[8];
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gd_local;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.0f;
            for (int i = 0; i < 8; ++i) tot += red[i];
            atomicAdd(g_decay, tot);
        }
    }
}

// ------------------------------------------------------------------------------------------ conv dgrad
// G[m][k] = sum_n g_acc[m][n] * W[k][n], scattered to g_x[t][src(m, tap)][c] with atomics (k = tap*Cin + c).
constexpr int DG_BM = 64, DG_BK = 64, DG_BN = 32;

__global__ void __launch_bounds__(256) conv_dgrad_kernel(const ConvParams p, const float* __restrict__ g_acc, float* g_x) {
    __shared__ __align__(16) float Gs[DG_BN][DG_BM + 4];  // [n][m]
    __shared__ __align__(16) float Ws[DG_BN][DG_BK + 4];  // [n][k]
    __shared__ int row_b[DG_BM], row_oy[DG_BM], row_ox[DG_BM];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * DG_BM;
    const int k0 = blockIdx.y * DG_BK;
    const int t = blockIdx.z;
    const int HW = p.Hout * p.Wout;
    if (tid < DG_BM) {
        const int m = m0 + tid;
        if (m < p.M) {
            const int b = m / HW, q = m - b * HW;
            row_b[tid] = b;
            row_oy[tid] = q / p.Wout;
            row_ox[tid] = q - (q / p.Wout) * p.Wout;
        } else {
            row_b[tid] = -1;
            row_oy[tid] = row_ox[tid] = 0;
        }
    }
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int nb = 0; nb < p.Cout; nb += DG_BN) {
        __syncthreads();
        for (int idx = tid; idx < DG_BM * DG_BN; idx += 256) {
            const int mm = idx / DG_BN, nn = idx - mm * DG_BN;
            const int m = m0 + mm;
            Gs[nn][mm] = (m < p.M) ? __ldg(g_acc + ((size_t)t * p.M + m) * p.Cout + nb + nn) : 0.0f;
        }
        for (int idx = tid; idx < DG_BK * DG_BN; idx += 256) {
            const int kk = idx / DG_BN, nn = idx - kk * DG_BN;
            const int k = k0 + kk;
            Ws[nn][kk] = (k < p.K) ? __ldg(p.w_kn + (size_t)k * p.Cout + nb + nn) : 0.0f;
        }
        __syncthreads();
#pragma unroll 8
        for (int nn = 0; nn < DG_BN; ++nn) {
            const float4 g4 = *reinterpret_cast<const float4*>(&Gs[nn][ty * 4]);
            const float4 w4 = *reinterpret_cast<const float4*>(&Ws[nn][tx * 4]);
            const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], wv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int mm = ty * 4 + i;
        const int b = row_b[mm];
        if (b < 0) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k >= p.K) continue;
            const int tap = k / p.Cin, c = k - tap * p.Cin;
            const int ky = tap / p.ks, kx = tap - ky * p.ks;
            const int sy = __ldg(p.ymap + row_oy[mm] * p.ks + ky);
            const int sx = __ldg(p.xmap + row_ox[mm] * p.ks + kx);
            if (sy < 0 || sx < 0) continue;
            atomicAdd(g_x + ((((size_t)t * p.B + b) * p.Hin + sy) * p.Win + sx) * p.Cin + c, acc[i][j]);
        }
    }
}

// ------------------------------------------------------------------------------------------ conv wgrad
// g_w[k][n] += sum over (t, m) of A[t][m][k] * g_acc[t][m][n]; the (t, m) range is split across blockIdx.z.
constexpr int WG_BK = 64, WG_BN = 64, WG_BM = 32;

template <int IN_LAYOUT>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const ConvParams p, const float* __restrict__ g_acc, float* g_w,
                                                         int rows_per_split) {
    __shared__ __align__(16) float As[WG_BM][WG_BK + 4];  // [m][k]
    __shared__ __align__(16) float Gs[WG_BM][WG_BN + 4];  // [m][n]
    const int tid = threadIdx.x;
    const int k0 = blockIdx.x * WG_BK;
    const int n0 = blockIdx.y * WG_BN;
    const long long TM = (long long)p.T * p.M;
    const long long r_begin = (long long)blockIdx.z * rows_per_split;
    long long r_end = r_begin + rows_per_split;
    if (r_end > TM) r_end = TM;
    const int HW = p.Hout * p.Wout;
    const int tx = tid & 15, ty = tid >> 4;  // thread owns k = ty*4.., n = tx*4..
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (long long r0 = r_begin; r0 < r_end; r0 += WG_BM) {
        __syncthreads();
        // gather A: 32 rows x 64 k
        for (int idx = tid; idx < WG_BM * WG_BK; idx += 256) {
            const int mm = idx / WG_BK, kk = idx - mm * WG_BK;
            const long long r = r0 + mm;
            const int k = k0 + kk;
            float a = 0.0f;
            if (r < r_end && k < p.K) {
                const int t = (int)(r / p.M);
                const int m = (int)(r - (long long)t * p.M);
                const int b = m / HW, q = m - b * HW;
                const int oy = q / p.Wout, ox = q - oy * p.Wout;
                const int tap = k / p.Cin, c = k - tap * p.Cin;
                const int ky = tap / p.ks, kx = tap - ky * p.ks;
                const int sy = __ldg(p.ymap + oy * p.ks + ky);
                const int sx = __ldg(p.xmap + ox * p.ks + kx);
                if (sy >= 0 && sx >= 0) {
                    if (IN_LAYOUT == SS_IN_U8_TBHWC)
                        a = (float)reinterpret_cast<const uint8_t*>(
                            p.x)[((((size_t)t * p.B + b) * p.Hin + sy) * p.Win + sx) * p.Cin + c];
                    else
                        a = __ldg(reinterpret_cast<const float*>(p.x) +
                                  ((((size_t)b * p.T + t) * p.Cin + c) * p.Hin + sy) * p.Win + sx);
                }
            }
            As[mm][kk] = a;
        }
        for (int idx = tid; idx < WG_BM * WG_BN; idx += 256) {
            const int mm = idx / WG_BN, nn = idx - mm * WG_BN;
            const long long r = r0 + mm;
            Gs[mm][nn] = (r < r_end && n0 + nn < p.Cout) ? __ldg(g_acc + (size_t)r * p.Cout + n0 + nn) : 0.0f;
        }
        __syncthreads();
#pragma unroll 8
        for (int mm = 0; mm < WG_BM; ++mm) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[mm][ty * 4]);
            const float4 g4 = *reinterpret_cast<const float4*>(&Gs[mm][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], gv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + ty * 4 + i;
        if (k >= p.K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < p.Cout && acc[i][j] != 0.0f) atomicAdd(g_w + (size_t)k * p.Cout + n, acc[i][j]);
        }
    }
}

// ------------------------------------------------------------------------------------------ heads backward
struct HeadsBwdParams {
    int T, B, H, W;
    float gain;
    int C[4], Hs[4], Ws[4];
    const uint8_t* acts[4];
    const float* w[4];
    const int* ymap[4];
    const int* xmap[4];
    const float* g_depths;
    float* g_acts[4];
    float* g_w[4];
    float* g_bias[4];
    float* bins[4];
    int store;          // 1: g_acts is written (first writer of a fresh buffer), 0: accumulated into
};

// step 1: bin the per-pixel head gradients onto (source pixel, tap).  Class 0 = timesteps before the last
// (every head sees the sum of all four depth gradients), class 1 = last timestep (suffix sums).
// Gather form: one thread per (head, sample, source pixel).  Nearest-neighbour upsampling maps a contiguous run of virtual
// rows / columns onto each source row / column, so the output pixels that read source pixel s through tap (ky, kx) are a
// rectangle; the thread walks the union of the nine rectangles once and adds every depth gradient to the taps it belongs to.
// (The scatter form -- one thread per output pixel, 72 atomics each, most of them colliding -- took 0.58 ms of the 24 ms step.)
__device__ __forceinline__ int head_virtual_source(const int* __restrict__ map, int n_out, int v) {
    // source index of virtual coordinate v in [0, n_out + 2): the forward table is indexed by (output, tap) = (v - tap, tap)
    const int o = min(v, n_out - 1);
    return __ldg(map + o * 3 + (v - o));
}
// first virtual coordinate in [0, n_virtual] whose source index is >= s (the map is non-decreasing)
__device__ __forceinline__ int head_lower_bound(const int* __restrict__ map, int n_out, int s) {
    int lo = 0, hi = n_out + 2;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (head_virtual_source(map, n_out, mid) < s) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(128) heads_bin_kernel(const HeadsBwdParams p, long long s_begin1, long long s_begin2, long long s_begin3,
                                                        long long s_end) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= s_end) return;
    const int i = gid >= s_begin3 ? 3 : (gid >= s_begin2 ? 2 : (gid >= s_begin1 ? 1 : 0));
    const long long base = i == 3 ? s_begin3 : (i == 2 ? s_begin2 : (i == 1 ? s_begin1 : 0));
    const int Hs = p.Hs[i], Ws = p.Ws[i];
    const int S = Hs * Ws;
    const long long q = gid - base;                 // b * S + sy * Ws + sx
    const int b = (int)(q / S);
    const int r = (int)(q - (long long)b * S);
    const int sy = r / Ws, sx = r - (r / Ws) * Ws;
    const int HW = p.H * p.W;
    // virtual rows / columns [v0, v1) that replicate this source pixel
    const int vy0 = head_lower_bound(p.ymap[i], p.H, sy), vy1 = head_lower_bound(p.ymap[i], p.H, sy + 1);
    const int vx0 = head_lower_bound(p.xmap[i], p.W, sx), vx1 = head_lower_bound(p.xmap[i], p.W, sx + 1);
    float acc0[9], acc1[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc0[k] = acc1[k] = 0.0f;
    const int y_lo = max(vy0 - 2, 0), y_hi = min(vy1 - 1, p.H - 1);
    const int x_lo = max(vx0 - 2, 0), x_hi = min(vx1 - 1, p.W - 1);
    const float* gd = p.g_depths + (size_t)b * HW;
    const size_t plane_d = (size_t)p.B * HW;
    for (int y = y_lo; y <= y_hi; ++y) {
        for (int x = x_lo; x <= x_hi; ++x) {
            const size_t o = (size_t)y * p.W + x;
            const float g0 = __ldg(gd + o), g1 = __ldg(gd + plane_d + o), g2 = __ldg(gd + 2 * plane_d + o), g3 = __ldg(gd + 3 * plane_d + o);
            const float g_all = ((g0 + g1) + g2) + g3;
            // suffix sums: head i sees the depth maps i..3 at the last timestep
            const float g_last = i == 3 ? g3 : (i == 2 ? g3 + g2 : (i == 1 ? (g3 + g2) + g1 : ((g3 + g2) + g1) + g0));
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const bool in_y = y + ky >= vy0 && y + ky < vy1;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    if (in_y && x + kx >= vx0 && x + kx < vx1) {
                        acc0[ky * 3 + kx] += g_all;
                        acc1[ky * 3 + kx] += g_last;
                    }
                }
            }
        }
    }
    const size_t plane = (size_t)p.B * S * 9;
    float* dst = p.bins[i] + (size_t)q * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        dst[k] = p.T > 1 ? acc0[k] : 0.0f;
        dst[plane + k] = acc1[k];
    }
}

// bias gradient: gain * sum over pixels and timesteps of the head gradient (class-0 gradient T-1 times + the last step's)
__global__ void __launch_bounds__(256) heads_bias_kernel(const HeadsBwdParams p) {
    const int HW = p.H * p.W;
    const long long n = (long long)p.B * HW;
    float bs[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < n; pix += (long long)gridDim.x * blockDim.x) {
        float gd[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) gd[i] = __ldg(p.g_depths + (size_t)i * n + pix);
        const float g_all = ((gd[0] + gd[1]) + gd[2]) + gd[3];
        float suffix = 0.0f;
#pragma unroll
        for (int i = 3; i >= 0; --i) {
            suffix += gd[i];
            bs[i] += g_all * (float)(p.T - 1) + suffix;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float vsum = bs[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
        if ((threadIdx.x & 31) == 0 && vsum != 0.0f) atomicAdd(p.g_bias[i], vsum * p.gain);
    }
}

// step 2: per source pixel s and 8-channel group:
//   GACT: g_act[t][s][c] (+)= gain * sum_tap bin_cls(t)[s][tap] * w[tap][c]          (cls = 1 for the last timestep, else 0)
//   GW  : g_w[tap][c]    += gain * (bin_0[s][tap] * sum_{t<T-1} act[t][s][c] + bin_1[s][tap] * act[T-1][s][c])
// Two instantiations: the activation-gradient writer is a pure streaming kernel that wants many resident warps; the
// weight-gradient one keeps 9 x 8 partials per thread in registers (they meet in shared memory once per block and reach
// HBM with 9*C atomics per block) and only reads the u8 activations.
template <bool GACT, bool GW>
__global__ void __launch_bounds__(256, GW ? 2 : 4) heads_src_kernel(const HeadsBwdParams p, int head) {
    extern __shared__ float sh[];  // w [9][C] then g_w accumulators [9][C]
    const int C = p.C[head];
    float* wsm = sh;
    float* gws = sh + 9 * C;
    for (int j = threadIdx.x; j < 9 * C; j += blockDim.x) {
        wsm[j] = __ldg(p.w[head] + j);
        gws[j] = 0.0f;
    }
    __syncthreads();
    const int S = p.B * p.Hs[head] * p.Ws[head];
    const size_t plane = (size_t)S * 9;
    const int c8n = C / 8;
    const int cg = threadIdx.x % c8n;            // channel group of this thread
    const int c0 = cg * 8;
    const int lanes = blockDim.x / c8n;          // pixels handled concurrently by one block
    const int pl = threadIdx.x / c8n;
    const int T = p.T;
    float acc[GW ? 9 : 1][8];
    if (GW) {
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[GW ? k : 0][e] = 0.0f;
    }
    if (pl < lanes) {
        for (long long s = (long long)blockIdx.x * lanes + pl; s < S; s += (long long)gridDim.x * lanes) {
            float b0[9], b1[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                b0[k] = (T > 1) ? __ldg(p.bins[head] + (size_t)s * 9 + k) * p.gain : 0.0f;
                b1[k] = __ldg(p.bins[head] + plane + (size_t)s * 9 + k) * p.gain;
            }
            if (GACT) {
                float ga0[8], ga1[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) ga0[e] = ga1[e] = 0.0f;
#pragma unroll
                for (int k = 0; k < 9; ++k)
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float wv = wsm[k * C + c0 + e];
                        ga0[e] = fmaf(b0[k], wv, ga0[e]);
                        ga1[e] = fmaf(b1[k], wv, ga1[e]);
                    }
                for (int t = 0; t < T; ++t) {
                    float4* gp = reinterpret_cast<float4*>(p.g_acts[head] + ((size_t)t * S + s) * C + c0);
                    float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
                    if (!p.store) {
                        q0 = gp[0];
                        q1 = gp[1];
                    }
                    const bool last = t == T - 1;
                    q0.x += last ? ga1[0] : ga0[0]; q0.y += last ? ga1[1] : ga0[1];
                    q0.z += last ? ga1[2] : ga0[2]; q0.w += last ? ga1[3] : ga0[3];
                    q1.x += last ? ga1[4] : ga0[4]; q1.y += last ? ga1[5] : ga0[5];
                    q1.z += last ? ga1[6] : ga0[6]; q1.w += last ? ga1[7] : ga0[7];
                    gp[0] = q0;
                    gp[1] = q1;
                }
            }
            if (GW) {
                // sum of the first T-1 steps (small integers: exact in fp32) and the last step; TU timesteps are loaded together: the
                // block count is capped for the reduction below, so one 8-byte load in flight per thread left the kernel latency-bound
                float asum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, alast[8];
                constexpr int TU = 5;
                for (int t0 = 0; t0 < T; t0 += TU) {
                    uint2 raws[TU];
#pragma unroll
                    for (int j = 0; j < TU; ++j)
                        if (t0 + j < T)
                            raws[j] = __ldg(reinterpret_cast<const uint2*>(p.acts[head] + ((size_t)(t0 + j) * S + s) * C + c0));
#pragma unroll
                    for (int j = 0; j < TU; ++j) {
                        if (t0 + j >= T) break;
                        const bool last = t0 + j == T - 1;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float lo = (float)((raws[j].x >> (8 * e)) & 0xFFu), hi = (float)((raws[j].y >> (8 * e)) & 0xFFu);
                            if (last) {
                                alast[e] = lo;
                                alast[4 + e] = hi;
                            } else {
                                asum[e] += lo;
                                asum[4 + e] += hi;
                            }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 9; ++k)
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        acc[GW ? k : 0][e] = fmaf(b0[k], asum[e], fmaf(b1[k], alast[e], acc[GW ? k : 0][e]));
            }
        }
    }
    if (GW) {
        // block reduction of the 9 x 8 partials: lanes that share a channel group (lane % c8n) are folded with shuffles first, so
        // that one lane per group and warp reaches shared memory -- 256 threads x 72 atomics on 72 * c8n addresses (64-way
        // conflicts at C = 32) were 0.1 ms per block, most of this kernel
        const bool pow2 = (c8n & (c8n - 1)) == 0 && c8n <= 32;
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float v = acc[GW ? k : 0][e];
                if (pow2) {
                    for (int o = c8n; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane < c8n && v != 0.0f) atomicAdd(&gws[k * C + c0 + e], v);
                } else if (pl < lanes && v != 0.0f) {
                    atomicAdd(&gws[k * C + c0 + e], v);
                }
            }
        __syncthreads();
        for (int j = threadIdx.x; j < 9 * C; j += blockDim.x)
            if (gws[j] != 0.0f) atomicAdd(p.g_w[head] + j, gws[j]);
    }
}

int fill_params(const ss_conv_geom* g, ConvParams& p) {
    if (g == nullptr || g->Cin <= 0 || g->Cout <= 0 || g->ks <= 0) {
        set_error("bad geometry");
        return SS_EINVAL;
    }
    p = ConvParams();
    p.T = g->T; p.B = g->B; p.Hin = g->Hin; p.Win = g->Win; p.Cin = g->Cin;
    p.Hout = g->Hout; p.Wout = g->Wout; p.Cout = g->Cout; p.ks = g->ks;
    p.K = g->ks * g->ks * g->Cin;
    p.M = g->B * g->Hout * g->Wout;
    p.neuron = g->neuron; p.gain = g->gain; p.v_th = g->v_th; p.v_reset = g->v_reset; p.tau = g->tau;
    return SS_OK;
}

}  // namespace
}  // namespace ss

using namespace ss;

extern "C" int ss_neuron_bwd_ex(int32_t T, int64_t N, int32_t neuron, int32_t surrogate, float alpha, float gain, float v_th,
                                float v_reset, float tau, const float* decay, const float* h_seq, const float* v_init,
                                const float* g_s, const float* g_v_last, float* g_acc, void* g_acc_bf16, float* g_v_init,
                                float* g_decay, void* stream) {
    if (h_seq == nullptr || g_s == nullptr || (g_acc == nullptr && g_acc_bf16 == nullptr) || T < 0 || N < 0) {
        set_error("ss_neuron_bwd: bad argument");
        return SS_EINVAL;
    }
    if (neuron == SS_NEURON_PLIF && decay == nullptr) {
        set_error("ss_neuron_bwd: PLIF needs decay");
        return SS_EINVAL;
    }
    if (T == 0 || N == 0) return SS_OK;
    auto aligned16 = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if (N % 4 == 0 && aligned16(h_seq) && aligned16(g_s) && aligned16(g_acc) && aligned16(g_acc_bf16) && aligned16(v_init) &&
        aligned16(g_v_last) && aligned16(g_v_init)) {
        const long long nt = N / 4;
        neuron_bwd_kernel<4><<<(unsigned)((nt + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            T, N, neuron, surrogate, alpha, gain, v_th, v_reset, tau, decay, h_seq, v_init, g_s, g_v_last, g_acc,
            reinterpret_cast<__nv_bfloat16*>(g_acc_bf16), g_v_init, g_decay);
    } else {
        neuron_bwd_kernel<1><<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            T, N, neuron, surrogate, alpha, gain, v_th, v_reset, tau, decay, h_seq, v_init, g_s, g_v_last, g_acc,
            reinterpret_cast<__nv_bfloat16*>(g_acc_bf16), g_v_init, g_decay);
    }
    count_launch();
    return check_launch("neuron_bwd");
}

extern "C" int ss_neuron_bwd(int32_t T, int64_t N, int32_t neuron, int32_t surrogate, float alpha, float gain, float v_th,
                             float v_reset, float tau, const float* decay, const float* h_seq, const float* v_init,
                             const float* g_s, const float* g_v_last, float* g_acc, float* g_v_init, float* g_decay,
                             void* stream) {
    return ss_neuron_bwd_ex(T, N, neuron, surrogate, alpha, gain, v_th, v_reset, tau, decay, h_seq, v_init, g_s, g_v_last, g_acc,
                            nullptr, g_v_init, g_decay, stream);
}

extern "C" int ss_conv_dgrad(const ss_conv_geom* g, const int32_t* ymap, const int32_t* xmap, const float* w_kn,
                             const float* g_acc, float* g_x, void* stream) {
    ConvParams p;
    if (fill_params(g, p) != SS_OK) return SS_EINVAL;
    if (ymap == nullptr || xmap == nullptr || w_kn == nullptr || g_acc == nullptr || g_x == nullptr) {
        set_error("ss_conv_dgrad: null argument");
        return SS_EINVAL;
    }
    if (p.Cout % DG_BN != 0) {
        set_error("ss_conv_dgrad: Cout %% 32 != 0");
        return SS_EINVAL;
    }
    if (p.T == 0 || p.M == 0) return SS_OK;
    p.ymap = ymap; p.xmap = xmap; p.w_kn = w_kn;
    dim3 grid((p.M + DG_BM - 1) / DG_BM, (p.K + DG_BK - 1) / DG_BK, p.T);
    conv_dgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g_acc, g_x);
    count_launch();
    return check_launch("conv_dgrad");
}

extern "C" int ss_conv_wgrad(const ss_conv_geom* g, const void* x, const int32_t* ymap, const int32_t* xmap,
                             const float* g_acc, float* g_w, void* stream) {
    ConvParams p;
    if (fill_params(g, p) != SS_OK) return SS_EINVAL;
    if (x == nullptr || ymap == nullptr || xmap == nullptr || g_acc == nullptr || g_w == nullptr) {
        set_error("ss_conv_wgrad: null argument");
        return SS_EINVAL;
    }
    if (p.T == 0 || p.M == 0) return SS_OK;
    p.x = x; p.ymap = ymap; p.xmap = xmap;
    const long long TM = (long long)p.T * p.M;
    const int tiles = ((p.K + WG_BK - 1) / WG_BK) * ((p.Cout + WG_BN - 1) / WG_BN);
    // enough splits to fill the machine ~4x, each at least 256 rows
    long long splits = (148LL * 4 + tiles - 1) / tiles;
    if (splits < 1) splits = 1;
    long long rows = (TM + splits - 1) / splits;
    if (rows < 256) rows = 256;
    rows = (rows + WG_BM - 1) / WG_BM * WG_BM;
    splits = (TM + rows - 1) / rows;
    if (splits > 65535) {
        set_error("ss_conv_wgrad: too many splits");
        return SS_EINVAL;
    }
    dim3 grid((p.K + WG_BK - 1) / WG_BK, (p.Cout + WG_BN - 1) / WG_BN, (unsigned)splits);
    if (g->in_layout == SS_IN_U8_TBHWC)
        conv_wgrad_kernel<SS_IN_U8_TBHWC><<<grid, 256, 0, (cudaStream_t)stream>>>(p, g_acc, g_w, (int)rows);
    else
        conv_wgrad_kernel<SS_IN_F32_BTCHW><<<grid, 256, 0, (cudaStream_t)stream>>>(p, g_acc, g_w, (int)rows);
    count_launch();
    return check_launch("conv_wgrad");
}

extern "C" int ss_heads_bwd(const ss_heads_args* a, const float* g_depths, float* const* g_acts, float* const* g_w,
                            float* const* g_bias, float* const* bins, int32_t store_g_acts, void* stream) {
    if (a == nullptr || g_depths == nullptr || g_acts == nullptr || g_w == nullptr || g_bias == nullptr || bins == nullptr) {
        set_error("ss_heads_bwd: null argument");
        return SS_EINVAL;
    }
    HeadsBwdParams p;
    p.T = a->T; p.B = a->B; p.H = a->H; p.W = a->W; p.gain = a->gain;
    for (int i = 0; i < 4; ++i) {
        if (a->C[i] % 8 != 0) {
            set_error("ss_heads_bwd: channels not a multiple of 8");
            return SS_EINVAL;
        }
        p.C[i] = a->C[i]; p.Hs[i] = a->Hs[i]; p.Ws[i] = a->Ws[i];
        p.acts[i] = reinterpret_cast<const uint8_t*>(a->acts[i]);
        p.w[i] = a->w[i]; p.ymap[i] = a->ymap[i]; p.xmap[i] = a->xmap[i];
        p.g_acts[i] = g_acts[i]; p.g_w[i] = g_w[i]; p.g_bias[i] = g_bias[i]; p.bins[i] = bins[i];
    }
    p.g_depths = g_depths;
    p.store = store_g_acts ? 1 : 0;
    const long long npix = (long long)p.B * p.H * p.W;
    if (npix == 0 || p.T == 0) return SS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    {
        long long sb[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) sb[i + 1] = sb[i] + (long long)p.B * p.Hs[i] * p.Ws[i];
        heads_bin_kernel<<<(unsigned)((sb[4] + 127) / 128), 128, 0, st>>>(p, sb[1], sb[2], sb[3], sb[4]);
        count_launch();
        if (check_launch("heads_bin") != SS_OK) return SS_ECUDA;
        heads_bias_kernel<<<296, 256, 0, st>>>(p);
    }
    count_launch();
    if (check_launch("heads_bin") != SS_OK) return SS_ECUDA;
    for (int i = 0; i < 4; ++i) {
        const int lanes = 256 / (p.C[i] / 8);
        if (lanes < 1) {
            set_error("ss_heads_bwd: more than 2048 channels");
            return SS_EUNSUPPORTED;
        }
        const long long S = (long long)p.B * p.Hs[i] * p.Ws[i];
        // weight-gradient pass: one resident wave (128 registers x 256 threads = 2 blocks per SM) -- every block ends with 9*C
        // global atomics on the same addresses, so more blocks only add contention (0.1 ms per head at 8 blocks per SM)
        long long blocks = (S + lanes - 1) / lanes;
        if (blocks > 148 * 2) blocks = 148 * 2;
        const size_t smem = (size_t)18 * p.C[i] * sizeof(float);
        long long blocks_a = (S + lanes - 1) / lanes;
        if (blocks_a > 148 * 32) blocks_a = 148 * 32;
        heads_src_kernel<true, false><<<(unsigned)blocks_a, 256, smem, st>>>(p, i);
        heads_src_kernel<false, true><<<(unsigned)blocks, 256, smem, st>>>(p, i);
        count_launch(2);
        if (check_launch("heads_src") != SS_OK) return SS_ECUDA;
    }
    return SS_OK;
}
