// C-ABI entry points (include/stereospike_b200.h): argument validation, dispatch, error reporting.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "ss_common.cuh"

namespace ss {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return SS_ECUDA;
    }
    return SS_OK;
}

namespace {

// Stand-alone neuron layer (the reference calls IFNode directly for its I-neuron pool, SNN_models.py:150,172).
__global__ void __launch_bounds__(256) neuron_fwd_kernel(int T, long long N, int neuron, float v_th, float v_reset, float tau,
                                                         const float* __restrict__ decay_p, const float* __restrict__ x,
                                                         float* __restrict__ v_io, float* __restrict__ s_out,
                                                         float* __restrict__ h_seq) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float decay = (neuron == SS_NEURON_PLIF) ? __ldg(decay_p) : 0.0f;
    float v = v_io[n];
    for (int t = 0; t < T; ++t) {
        float h;
        const float s = neuron_step(neuron, x[(size_t)t * N + n], v, v_th, v_reset, tau, decay, h);
        s_out[(size_t)t * N + n] = s;
        if (h_seq != nullptr) h_seq[(size_t)t * N + n] = h;
    }
    v_io[n] = v;
}

}  // namespace
}  // namespace ss

using namespace ss;

extern "C" int ss_neuron_fwd(int32_t T, int64_t N, int32_t neuron, float v_th, float v_reset, float tau, const float* decay,
                             const float* x, float* v_io, float* s_out, float* h_seq, void* stream) {
    if (x == nullptr || v_io == nullptr || s_out == nullptr || T < 0 || N < 0 || neuron < SS_NEURON_IF ||
        neuron > SS_NEURON_PLIF || (neuron == SS_NEURON_PLIF && decay == nullptr)) {
        set_error("ss_neuron_fwd: bad argument");
        return SS_EINVAL;
    }
    if (T == 0 || N == 0) return SS_OK;
    neuron_fwd_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, N, neuron, v_th, v_reset, tau, decay,
                                                                                     x, v_io, s_out, h_seq);
    count_launch();
    return check_launch("neuron_fwd");
}

extern "C" int ss_abi_version(void) { return SS_ABI_VERSION; }
extern "C" const char* ss_last_error(void) { return g_err; }
extern "C" int64_t ss_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int ss_conv_neuron_fwd(const ss_conv_geom* g, const void* x, const int32_t* ymap, const int32_t* xmap,
                                  const float* w_kn, const float* decay, const float* v_in, float* v_out,
                                  const void* resid, void* out, float* h_seq, void* stream) {
    if (g == nullptr) {
        set_error("ss_conv_neuron_fwd: null geometry");
        return SS_EINVAL;
    }
    if (g->T == 0 || g->B == 0) return SS_OK;  // empty batch / sequence: nothing to do (empty tensors have null pointers)
    if (x == nullptr || ymap == nullptr || xmap == nullptr || out == nullptr) {
        set_error("ss_conv_neuron_fwd: null argument");
        return SS_EINVAL;
    }
    if (g->T < 0 || g->B < 0 || g->Hin <= 0 || g->Win <= 0 || g->Cin <= 0 || g->Hout <= 0 || g->Wout <= 0 ||
        g->Cout <= 0 || g->ks <= 0) {
        set_error("ss_conv_neuron_fwd: bad geometry");
        return SS_EINVAL;
    }
    if (g->neuron < SS_NEURON_IF || g->neuron > SS_NEURON_PLIF) {
        set_error("ss_conv_neuron_fwd: unknown neuron kind %d", g->neuron);
        return SS_EINVAL;
    }
    if (g->neuron == SS_NEURON_PLIF && decay == nullptr) {
        set_error("ss_conv_neuron_fwd: PLIF needs the decay scalar");
        return SS_EINVAL;
    }
    if (g->neuron == SS_NEURON_LIF && !(g->tau > 1.0f)) {
        set_error("ss_conv_neuron_fwd: LIF needs tau > 1");
        return SS_EINVAL;
    }
    ConvParams p;
    p.T = g->T; p.B = g->B; p.Hin = g->Hin; p.Win = g->Win; p.Cin = g->Cin;
    p.Hout = g->Hout; p.Wout = g->Wout; p.Cout = g->Cout; p.ks = g->ks;
    p.K = g->ks * g->ks * g->Cin;
    const long long M = (long long)g->B * g->Hout * g->Wout;
    if (M > 0x7fffffffLL / 2) {
        set_error("ss_conv_neuron_fwd: B*Hout*Wout too large");
        return SS_EINVAL;
    }
    p.M = (int)M;
    p.neuron = g->neuron; p.gain = g->gain; p.v_th = g->v_th; p.v_reset = g->v_reset; p.tau = g->tau;
    p.x = x; p.ymap = ymap; p.xmap = xmap; p.w_kn = w_kn; p.decay = decay; p.v_in = v_in; p.v_out = v_out;
    p.resid = reinterpret_cast<const uint8_t*>(resid);
    p.out = reinterpret_cast<uint8_t*>(out);
    p.h_seq = h_seq;

    if (w_kn == nullptr) {
        set_error("ss_conv_neuron_fwd: SIMT path needs w_kn");
        return SS_EINVAL;
    }
    return launch_conv_neuron_simt(p, g->in_layout, (cudaStream_t)stream);
}
