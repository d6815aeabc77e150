// Training loss and depth metric of the reference in two fused passes.
//
// Replaces network/loss.py:7-135 (ScaleInvariant_Loss, GradientMatching_Loss and their multi-scale wrappers, Total_Loss)
// and network/metrics.py:83-95 (MeanDepthError) -- about forty small PyTorch kernels with boolean-mask gathers per call
// (train.py:238,257) -- by
//   ss_loss_fwd: per valid pixel and scale, the residual statistics  n, sum r, sum r^2, sum |r|  and the Sobel term
//                sum (|gx| + |gy|)  (zero-padded 3x3 cross-correlations of the NaN-masked residual), reduced in fp64;
//                the signs of gx / gy are kept (2 x 2 bits per pixel and scale) for the backward pass;
//   ss_loss_bwd: d loss / d pred_k = a_k * (2 r / n - 2 S1 / n^2) + b_k / n * sum_q (sgn gx_q * SX[p-q] + sgn gy_q * SY[p-q])
//                on valid pixels, 0 elsewhere (the reference zeroes the residual there in place).
// HBM-bound: (1 + nscale) * 4 B per pixel in, nscale B out (forward); + nscale * 4 B out (backward).
#include "ss_common.cuh"

namespace ss {
namespace {

struct LossParams {
    int nscale, B, H, W;
    const float* pred[4];
    const float* gt;
    double* sums;        // [nscale][5]: n, sum r, sum r^2, sum (|gx| + |gy|), sum |r|
    uint8_t* signs;      // [nscale][B*H*W]: (sgn gx + 1) | (sgn gy + 1) << 2, 0x5 (= both zero) on invalid pixels
    float* g_pred[4];
    const float* coef_si;   // backward: device float[nscale] = scale-invariant weight * upstream gradient
    const float* coef_gm;   // ... gradient-matching weight * upstream gradient
};

__device__ __forceinline__ int sgn(float v) { return (v > 0.0f) - (v < 0.0f); }

__global__ void __launch_bounds__(256) loss_fwd_kernel(const LossParams p) {
    const long long HW = (long long)p.H * p.W;
    const long long N = (long long)p.B * HW;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float acc[4][5];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[k][j] = 0.0f;
    if (pix < N) {
        const long long b = pix / HW;
        const int q = (int)(pix - b * HW);
        const int y = q / p.W, x = q - y * p.W;
        float g[9];
        long long off[9];
        bool ok[9];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int yy = y + dy - 1, xx = x + dx - 1;
                const bool in = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
                const long long o = b * HW + (long long)yy * p.W + xx;
                off[dy * 3 + dx] = o;
                g[dy * 3 + dx] = in ? __ldg(p.gt + o) : 0.0f;
                ok[dy * 3 + dx] = in && !isnan(g[dy * 3 + dx]);
            }
        const bool valid = ok[4];
        for (int k = 0; k < p.nscale; ++k) {
            float r[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) r[j] = ok[j] ? __ldg(p.pred[k] + off[j]) - g[j] : 0.0f;
            // conv2d = cross-correlation with [[1,0,-1],[2,0,-2],[1,0,-1]] and [[1,2,1],[0,0,0],[-1,-2,-1]], zero padding 1
            const float gx = (r[0] - r[2]) + 2.0f * (r[3] - r[5]) + (r[6] - r[8]);
            const float gy = (r[0] + 2.0f * r[1] + r[2]) - (r[6] + 2.0f * r[7] + r[8]);
            uint8_t s = 0x5;
            if (valid) {
                acc[k][0] = 1.0f;
                acc[k][1] = r[4];
                acc[k][2] = r[4] * r[4];
                acc[k][3] = fabsf(gx) + fabsf(gy);
                acc[k][4] = fabsf(r[4]);
                s = (uint8_t)((sgn(gx) + 1) | ((sgn(gy) + 1) << 2));
            }
            if (p.signs != nullptr) p.signs[(size_t)k * N + pix] = s;
        }
    }
    __shared__ double red[8][20];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            double v = (double)acc[k][j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][k * 5 + j] = v;
        }
    __syncthreads();
    if (threadIdx.x < p.nscale * 5) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        if (v != 0.0) atomicAdd(p.sums + threadIdx.x, v);
    }
}

__global__ void __launch_bounds__(256) loss_bwd_kernel(const LossParams p) {
    const long long HW = (long long)p.H * p.W;
    const long long N = (long long)p.B * HW;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= N) return;
    const long long b = pix / HW;
    const int q = (int)(pix - b * HW);
    const int y = q / p.W, x = q - y * p.W;
    const float gtv = __ldg(p.gt + pix);
    const bool valid = !isnan(gtv);
    for (int k = 0; k < p.nscale; ++k) {
        float gout = 0.0f;
        if (valid) {
            const double n = p.sums[k * 5 + 0], s1 = p.sums[k * 5 + 1];
            const float r = __ldg(p.pred[k] + pix) - gtv;
            // Sobel term: neighbour q = p + (ey, ex) saw this pixel through kernel entry [1 - ey][1 - ex]
            int sob = 0;
#pragma unroll
            for (int ey = -1; ey <= 1; ++ey)
#pragma unroll
                for (int ex = -1; ex <= 1; ++ex) {
                    const int yy = y + ey, xx = x + ex;
                    if (yy < 0 || yy >= p.H || xx < 0 || xx >= p.W) continue;
                    const uint8_t s = p.signs[(size_t)k * N + b * HW + (long long)yy * p.W + xx];
                    const int sx = (int)(s & 3u) - 1, sy = (int)((s >> 2) & 3u) - 1;
                    const int ky = 1 - ey, kx = 1 - ex;
                    const int wx = (kx == 0 ? 1 : (kx == 2 ? -1 : 0)) * (ky == 1 ? 2 : 1);      // SX[ky][kx]
                    const int wy = (ky == 0 ? 1 : (ky == 2 ? -1 : 0)) * (kx == 1 ? 2 : 1);      // SY[ky][kx]
                    sob += sx * wx + sy * wy;
                }
            const double gsi = 2.0 * (double)r / n - 2.0 * s1 / (n * n);
            gout = (float)((double)__ldg(p.coef_si + k) * gsi + (double)__ldg(p.coef_gm + k) * (double)sob / n);
        }
        p.g_pred[k][pix] = gout;
    }
}

int fill(LossParams& p, int nscale, int B, int H, int W, const float* const* pred, const float* gt) {
    if (nscale < 1 || nscale > 4 || B < 0 || H <= 0 || W <= 0 || pred == nullptr || gt == nullptr) {
        set_error("ss_loss: bad argument (1 <= nscale <= 4)");
        return SS_EINVAL;
    }
    p = LossParams();
    p.nscale = nscale; p.B = B; p.H = H; p.W = W; p.gt = gt;
    for (int k = 0; k < nscale; ++k) {
        if (pred[k] == nullptr) {
            set_error("ss_loss: null prediction");
            return SS_EINVAL;
        }
        p.pred[k] = pred[k];
    }
    return SS_OK;
}

}  // namespace
}  // namespace ss

using namespace ss;

extern "C" int ss_loss_fwd(int32_t nscale, int32_t B, int32_t H, int32_t W, const float* const* pred, const float* gt, double* sums,
                           void* signs, void* stream) {
    LossParams p;
    if (fill(p, nscale, B, H, W, pred, gt) != SS_OK) return SS_EINVAL;
    if (sums == nullptr) {
        set_error("ss_loss_fwd: null sums");
        return SS_EINVAL;
    }
    const long long N = (long long)B * H * W;
    if (N == 0) return SS_OK;
    p.sums = sums;
    p.signs = reinterpret_cast<uint8_t*>(signs);
    loss_fwd_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    return check_launch("loss_fwd");
}

extern "C" int ss_loss_bwd(int32_t nscale, int32_t B, int32_t H, int32_t W, const float* const* pred, const float* gt,
                           const double* sums, const void* signs, const float* coef_si, const float* coef_gm, float* const* g_pred,
                           void* stream) {
    LossParams p;
    if (fill(p, nscale, B, H, W, pred, gt) != SS_OK) return SS_EINVAL;
    if (sums == nullptr || signs == nullptr || coef_si == nullptr || coef_gm == nullptr || g_pred == nullptr) {
        set_error("ss_loss_bwd: null argument");
        return SS_EINVAL;
    }
    const long long N = (long long)B * H * W;
    if (N == 0) return SS_OK;
    p.sums = const_cast<double*>(sums);
    p.signs = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(signs));
    p.coef_si = coef_si;
    p.coef_gm = coef_gm;
    for (int k = 0; k < nscale; ++k) {
        if (g_pred[k] == nullptr) {
            set_error("ss_loss_bwd: null gradient buffer");
            return SS_EINVAL;
        }
        p.g_pred[k] = g_pred[k];
    }
    loss_bwd_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    return check_launch("loss_bwd");
}
