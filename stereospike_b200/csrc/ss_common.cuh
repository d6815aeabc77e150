// Shared declarations for the stereospike_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "stereospike_b200.h"

namespace ss {

// Kernel-side view of one fused spiking block (ss_conv_geom + the device pointers of the call).
struct ConvParams {
    int T, B, Hin, Win, Cin, Hout, Wout, Cout, ks, K;
    int M;                      // B*Hout*Wout output pixels per timestep
    int neuron;
    float gain, v_th, v_reset, tau;
    const void* x;
    const int* ymap;
    const int* xmap;
    const float* w_kn;
    const float* decay;
    const float* v_in;
    float* v_out;
    const uint8_t* resid;
    uint8_t* out;
    float* h_seq;
};

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);

// Per-DEVICE launch state.  cudaFuncSetAttribute and the SM count belong to a device, not to the process: a caller that drives two
// GPUs from one process must get MaxDynamicSharedMemorySize set on each of them (`static bool attr` cached it once).
constexpr int SS_MAX_DEVICES = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < SS_MAX_DEVICES) ? dev : 0;
}
inline int device_sm_count(int dev) {
    static int sms[SS_MAX_DEVICES] = {};
    if (sms[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = n > 0 ? n : 148;
    }
    return sms[dev];
}
// once per (kernel instance = call site, device): opt in to the large dynamic shared memory carve-out
#define SS_ENSURE_SMEM(kernel, dev, bytes)                                                       \
    do {                                                                                         \
        static bool ss_attr_done[ss::SS_MAX_DEVICES] = {};                                       \
        if (!ss_attr_done[dev]) {                                                                \
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes));  \
            ss_attr_done[dev] = true;                                                            \
        }                                                                                        \
    } while (0)

int launch_conv_neuron_simt(const ConvParams& p, int in_layout, cudaStream_t st);

// Correctly rounded a / b for a loop-invariant divisor b (Markstein): y = RN(1 / b) hoisted out of the loop, then
//   q0 = RN(a * y);  r = RN(a - b * q0) (exact, one FMA);  q = RN(q0 + r * y).
// Bit-identical to IEEE division for EVERY fp32 a with |a| >= 1e-30 and a normal quotient -- checked exhaustively over all 2^32
// operands for 17 divisors, all-ones mantissas included (tools/div_probe.cu, profiles/r2h_div_probe.log); below that (and for the
// sign of a zero quotient) it may differ in the last bits, which no membrane potential can observe.  3 instructions per division
// instead of the 5 of the two-correction form used in round 1.
__device__ __forceinline__ float div_const_prepare(float b) { return __frcp_rn(b); }
__device__ __forceinline__ float div_const(float a, float b, float y) {
    const float q0 = __fmul_rn(a, y);
    const float r = fmaf(-b, q0, a);
    return fmaf(r, y, q0);
}

// Loop-invariant neuron parameters (rtau = div_const_prepare(tau)).
struct NeuronConst {
    float gain, v_th, v_reset, tau, rtau, decay;
};

// One neuron step (SpikingJelly BaseNode.forward: charge -> fire -> hard reset), fp32.
// Returns the spike (0/1); v is updated in place; h_out receives the pre-reset potential.
// VR0: v_reset == 0 known at compile time (every call site of the reference): x - (v - v_reset) is x - v, bit for bit
template <int KIND, bool VR0 = false>
__device__ __forceinline__ float neuron_step_t(float x, float& v, const NeuronConst& c, float& h_out) {
    float h;
    if (KIND == SS_NEURON_IF) {
        h = __fadd_rn(v, x);
    } else if (KIND == SS_NEURON_LIF) {
        // true division: (x - v) / tau is not (x - v) * (1 / tau) in fp32
        const float dv = (VR0 || c.v_reset == 0.0f) ? __fsub_rn(x, v) : __fsub_rn(x, __fsub_rn(v, c.v_reset));
        h = __fadd_rn(v, div_const(dv, c.tau, c.rtau));
    } else {
        const float dv = (VR0 || c.v_reset == 0.0f) ? __fsub_rn(x, v) : __fsub_rn(x, __fsub_rn(v, c.v_reset));
        h = __fadd_rn(v, __fmul_rn(dv, c.decay));
    }
    h_out = h;
    // heaviside(h - v_th): with gradual underflow h - v_th is zero only for h == v_th and has the sign of the exact
    // difference otherwise, so the comparison needs no subtraction
    const bool fire = h >= c.v_th;
    v = fire ? (VR0 ? 0.0f : c.v_reset) : h;
    return fire ? 1.0f : 0.0f;
}

__device__ __forceinline__ float neuron_step(int kind, float x, float& v, float v_th, float v_reset, float tau, float decay,
                                             float& h_out) {
    NeuronConst c;
    c.gain = 1.0f; c.v_th = v_th; c.v_reset = v_reset; c.tau = tau; c.rtau = 0.0f; c.decay = decay;
    if (kind == SS_NEURON_IF) return neuron_step_t<SS_NEURON_IF>(x, v, c, h_out);
    if (kind == SS_NEURON_LIF) {
        c.rtau = div_const_prepare(tau);
        return neuron_step_t<SS_NEURON_LIF>(x, v, c, h_out);
    }
    return neuron_step_t<SS_NEURON_PLIF>(x, v, c, h_out);
}

}  // namespace ss
