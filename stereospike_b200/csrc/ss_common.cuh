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

int launch_conv_neuron_simt(const ConvParams& p, int in_layout, cudaStream_t st);

// One neuron step (SpikingJelly BaseNode.forward: charge -> fire -> hard reset), fp32.
// Returns the spike (0/1); v is updated in place; h_out receives the pre-reset potential.
__device__ __forceinline__ float neuron_step(int kind, float x, float& v, float v_th, float v_reset,
                                             float tau, float decay, float& h_out) {
    float h;
    if (kind == SS_NEURON_IF) {
        h = v + x;
    } else if (kind == SS_NEURON_LIF) {
        // true division: (x - v) / tau is not (x - v) * (1 / tau) in fp32
        h = (v_reset == 0.0f) ? v + __fdiv_rn(x - v, tau) : v + __fdiv_rn(x - (v - v_reset), tau);
    } else {
        h = (v_reset == 0.0f) ? __fadd_rn(v, __fmul_rn(x - v, decay))
                              : __fadd_rn(v, __fmul_rn(x - (v - v_reset), decay));
    }
    h_out = h;
    const float s = (h - v_th >= 0.0f) ? 1.0f : 0.0f;
    v = (s != 0.0f) ? v_reset : h;
    return s;
}

}  // namespace ss
