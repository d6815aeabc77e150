/* stereospike_b200 -- C ABI of the B200-native StereoSpike hot path.
 *
 * The reference (urancon/StereoSpike) is pure Python and has no FFI: its "operator interface" for this
 * path is the nn.Module surface of network/blocks.py and network/SNN_models.py plus the SpikingJelly
 * neuron modules they call.  Each entry point below names the reference code it replaces; the Python
 * host side (stereospike_b200/*.py) mirrors the reference classes and reaches these symbols through
 * ctypes.
 *
 * Conventions: plain pointers and sizes only.  Every pointer is a DEVICE pointer owned by the caller
 * (the library never allocates or frees device memory and never synchronises the host); `stream` is a
 * cudaStream_t passed as void*.  Functions return 0 on success or a negative SS_E* code, and
 * ss_last_error() returns a thread-local description of the last failure.
 *
 * Data layout in HBM
 *   spikes / activations : u8,   [T][B][H][W][C]  (timestep-major NHWC; values are small non-negative
 *                          integers: spikes {0,1}, spike sums {0..3}, event counts {0..255} -- exact)
 *   first-layer input    : fp32, [B][T][C][H][W]  (the reference's own NCHW event-count frames,
 *                          train.py:201-218); ss_pack_events turns it into u8 [T][B][H][W][4] for the
 *                          tensor-core path, the SIMT path reads it directly
 *   weights              : fp32 OIHW as in the reference's state dict; ss_pack_weights_i8 derives the
 *                          int8 digit planes + per-channel power-of-two scale the tensor-core path uses
 *   membrane potentials  : fp32, [B][H][W][C]
 *   pre-reset potentials : fp32, [T][B][H][W][C]  (optional, saved for the surrogate backward)
 *   depth (I-neurons)    : fp32, [B][H][W]
 */
#ifndef STEREOSPIKE_B200_H
#define STEREOSPIKE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SS_ABI_VERSION 3

/* error codes */
#define SS_OK 0
#define SS_EINVAL (-1)      /* bad argument / unsupported geometry */
#define SS_ECUDA (-2)       /* a CUDA runtime / driver call failed (launch errors included) */
#define SS_EUNSUPPORTED (-3)

/* neuron kinds -- spikingjelly.clock_driven.neuron.{IFNode,LIFNode,ParametricLIFNode}
 * (call sites network/SNN_models.py:78,266 and network/blocks.py:150,157) */
#define SS_NEURON_IF 0
#define SS_NEURON_LIF 1     /* h = v + (x - (v - v_reset)) / tau                */
#define SS_NEURON_PLIF 2    /* h = v + (x - (v - v_reset)) * (*decay)  (decay = sigmoid(w), device scalar) */

/* surrogate kinds -- spikingjelly.clock_driven.surrogate.{ATan(alpha=2),Sigmoid(alpha=4)}
 * (train.py:118, network/blocks.py:142) */
#define SS_SURR_ATAN 0
#define SS_SURR_SIGMOID 1

/* input layouts of ss_conv_neuron_fwd */
#define SS_IN_U8_TBHWC 0
#define SS_IN_F32_BTCHW 1

/* implementations */
#define SS_IMPL_AUTO 0
#define SS_IMPL_SIMT 1      /* fp32 CUDA-core implicit GEMM (exact fp32 weights) */
#define SS_IMPL_UMMA 2      /* tcgen05 tensor-core implicit GEMM (ss_conv_i8_fwd) */

/* ------------------------------------------------------------------------------------------------
 * Fused spiking block on the tensor cores (the hot path).
 * Replaces  nn.Sequential(Conv2d(bias=False) | NNConvUpsampling, MultiplyBy, neuron)  called once per
 * timestep (network/SNN_models.py:75-129, network/blocks.py:110-132,145-157) and, through `resid`, the
 * skip additions (SNN_models.py:171,176,181,186) and the SEW 'ADD' connect (blocks.py:170-171).
 * All T timesteps run in ONE launch; the membrane potential stays in registers across the time loop. */
typedef struct ss_block_desc {
    int32_t T, B;
    int32_t Hin, Win, Cin;       /* source activation, u8 [T][B][Hin][Win][Cin], Cin % 32 == 0 -- or Cin == 4:
                                    first-layer mode (packed event frames, 5x5 stride-1 conv) */
    int32_t Hout, Wout, Cout;    /* block output,     u8 [T][B][Hout][Wout][Cout], Cout % 32 == 0 */
    int32_t ks;                  /* 3 or 5 */
    int32_t stride;              /* 1 or 2 (2: ks 5, even pad) */
    int32_t pad;                 /* zero padding of the plain conv; ignored when upsample != 0 */
    int32_t upsample;            /* != 0: NNConvUpsampling -- nearest-neighbour upsample of the source to
                                    (Hout+ks-1, Wout+ks-1) (ATen index rule) followed by a valid conv */
    int32_t neuron;              /* SS_NEURON_* */
    int32_t planes;              /* int8 digit planes per weight: 2 (16-bit), 3 (24-bit, fp32-class), 4 (32-bit) */
    float gain;                  /* MultiplyBy scalar, applied to the conv result (blocks.py:106-107) */
    float v_th, v_reset, tau;
} ss_block_desc;

/* channel-block bytes the kernel will use for (Cin, ks); ss_pack_weights_i8 uses the same rule */
int ss_conv_i8_rowbytes(int32_t Cin, int32_t ks);

/* fp32 OIHW weights -> `planes` balanced base-256 digit planes of a per-output-channel power-of-two fixed
 * point, laid out [Cout/32][Cin/RB][ks*ks][planes*32][RB] and pre-swizzled as the kernel wants them in shared
 * memory (Cout*Cin*ks*ks*planes bytes; first-layer mode Cin <= 4, ks 5: [Cout/32][planes*32][128],
 * Cout*128*planes bytes, input channels zero-padded to 4).  wscale[n] = 2^wexp[n]:  w[n][...] ~= wscale[n] * sum_p digit_p * 256^(planes-1-p),
 * absolute error <= wscale[n] / 2. */
int ss_pack_weights_i8(const float* w_oihw, int32_t Cout, int32_t Cin, int32_t ks, int32_t planes, void* w_i8,
                       float* wscale, int32_t* wexp, void* stream);

/* fp32 [B][T][C][H][W] event-count frames (C <= 4) -> u8 [T][B][H][W][4] (channels >= C zero).  Counts are
 * rounded and clamped to 0..255; if any input is not already such an integer, bit 0 of *status (device int,
 * may be NULL) is set. */
int ss_pack_events(const float* x_btchw, int32_t B, int32_t T, int32_t C, int32_t H, int32_t W, void* out_tbhw4,
                   int32_t* status, void* stream);
/* Same into u8 [T][B][H][W][Cpad], C <= Cpad, Cpad = 4 or a multiple of 32 (<= 256): the channel-concatenated temporal mode of
 * the reference (train.py:206-218: the first conv gets 2 * nfpdm * cameras channels), whose first block then runs as an ordinary
 * Cin = Cpad block of ss_conv_i8_fwd with zero weights for the padding channels. */
int ss_pack_events_c(const float* x_btchw, int32_t B, int32_t T, int32_t C, int32_t Cpad, int32_t H, int32_t W, void* out,
                     int32_t* status, void* stream);

/* Event stream -> event-count frames (the step in front of the path; replaces mvsecRectifyEvents and
 * mvsecCumulateSpikesIntoFrames, datasets/MVSEC/utils.py:31-56,215-281).
 *   events_xytp : fp64 [n][4] = (x, y, t, polarity) as the reference stores them
 *   xmap, ymap  : fp64 [H][W] rectification look-up tables, or both NULL for already rectified events
 *   t0          : subtracted from every timestamp (utils.py:251-253)
 *   frame_start/_end : fp64 [n_frames]; an event is counted in frame f iff start[f] < t - t0 < end[f] (both strict, as upstream)
 *   camera      : 0 = left (channels 0-1), 1 = right (channels 2-3); channel = 2*camera + (polarity == 1 ? 0 : 1)
 *   counts_fhw4 : u32 [n_frames][H][W][4], ACCUMULATED into (caller zero-fills once, then calls once per camera)
 * ss_events_pack turns the counts (frame f = b*T + t) into the first block's input u8 [T][B][H][W][4], saturating at 255
 * (bit 1 of *status is set if it had to). */
int ss_events_accumulate(const double* events_xytp, int64_t n_events, const double* xmap, const double* ymap, double t0,
                         const double* frame_start, const double* frame_end, int32_t n_frames, int32_t H, int32_t W,
                         int32_t camera, uint32_t* counts_fhw4, void* stream);
int ss_events_pack(const uint32_t* counts_fhw4, int32_t B, int32_t T, int32_t H, int32_t W, void* out_tbhw4, int32_t* status,
                   void* stream);

/*   x      : u8 [T][B][Hin][Win][Cin]
 *   w_i8, wscale : from ss_pack_weights_i8 (same Cout, Cin, ks, planes)
 *   decay  : device scalar, PLIF only (sigmoid(w))
 *   v_in   : fp32 [B][Hout][Wout][Cout] initial membrane potential or NULL (= v_reset); v_out: final potential or NULL
 *   resid  : u8 [T][B][Hout][Wout][Cout] added to the spikes before they are written, or NULL
 *   out    : u8 [T][B][Hout][Wout][Cout]
 *   h_seq  : fp32 [T][B][Hout][Wout][Cout] pre-reset potential h_t (for the backward), or NULL
 *   tsum   : u8 [B][Hout][Wout][Cout] = sum over the first T-1 timesteps of `out` (needs 3*(T-1) <= 255), or NULL */
int ss_conv_i8_fwd(const ss_block_desc* g, const void* x, const void* w_i8, const float* wscale, const float* decay,
                   const float* v_in, float* v_out, const void* resid, void* out, float* h_seq, void* tsum, void* stream);

/* Tile iteration of ss_conv_i8_fwd_ex.  NNConvUpsampling (blocks.py:110-132) by ~2x replicates every source pixel 2 (rarely
 * 3) times, so the 5 taps of an axis of the 5x5 conv over the upsampled image read only 3 (or 2) DISTINCT source pixels per
 * output.  Per axis every output o has a first source pixel s0(o) and one of five replication patterns (the source index of
 * its 5 taps relative to s0): L (0,1,1,2,2), M (0,0,1,1,2) -- the two regular ones -- and next to a 3-fold replication
 * A (0,0,0,1,1), B (0,1,1,1,2), C (0,0,1,1,1).  With the integer weights of the replicated taps summed (exactly) the block
 * becomes, bit for bit:
 *   SS_TILES_FOLDED   : four 3x3 convs on the SOURCE image, one per (row class, column class) in {L,M}^2 -- 9 taps instead of
 *                       25 for every output whose row AND column are regular (93 % at 260x346).  g = the virtual conv (ks 3,
 *                       stride 1, pad 0, upsample 0, Hin/Win = source, Hout/Wout = REAL output); w_i8 holds 4 weight sets per
 *                       output-channel tile (packed with Cout' = 4*Cout, set = tile*4 + class, class = 2*row_class +
 *                       col_class); ymap_out [2][Hin-2], xmap_out [2][Win-2]: real output row / column of virtual position s
 *                       for class L / M, or -1.
 *   SS_TILES_ROW_LIST : the irregular output rows (classes A, B, C) over all columns: rows folded to 3 taps, columns still
 *                       the 5 taps over the upsampled row -- 15 taps.  g = the real upsampled conv (ks 5, upsample 1).  A
 *                       tile row is an entry of the class's list: rl_src / rl_out [nclass][rl_n] = pixel offset (inside one
 *                       timestep, i.e. (b*H + row)*W) of the entry's first source row / of its output row, -1 = padding entry;
 *                       w_i8 = nclass weight sets per output-channel tile (ss_pack_digits_i8_rect with Cout' = nclass*Cout,
 *                       ksy 3, ksx 5).  transposed = 1: the same pass for the irregular output COLUMNS -- the list entries
 *                       are (sample, output column) (offsets b*H*W + column), the tile columns walk the rows, the weight sets
 *                       are those of the transposed filter, and rl_collive [Hout] masks the rows the row pass already wrote. */
#define SS_TILES_PLAIN 0
#define SS_TILES_FOLDED 1
#define SS_TILES_ROW_LIST 2
typedef struct ss_tile_maps {
    int32_t mode;
    int32_t nclass;        /* ROW_LIST: weight sets per output-channel tile */
    int32_t rl_n;          /* ROW_LIST: entries per class (lists padded with -1 to a common length) */
    int32_t transposed;    /* ROW_LIST: 0 = list of output rows, 1 = list of output columns, 2 = list of output rows with FOLDED
                              columns: the dense 3x3 sets (w_i8 = w_dense, nclass = 4 = (row class, column class)) evaluated only
                              on the regular rows of each row class -- rl_src / rl_out [2][rl_n], output column through xmap_out;
                              replaces the SS_TILES_FOLDED pass where many source rows are irregular (Cin % 64 == 0) */
    const int32_t* ymap_out;
    const int32_t* xmap_out;
    const int32_t* rl_src;
    const int32_t* rl_out;
    const uint8_t* rl_collive;
    uint64_t* stats;       /* any mode, optional: [6] counters this launch ADDS to -- {spikes fired, nonzero outputs (spikes + skip),
                              sum of outputs^2} over all T steps, then the same three over the last step only.  The firing rates of
                              SNN_models.py:194-245 and the spike penalty of loss.py:96-107 without a pass over the spike maps. */
    const int32_t* item_tab;   /* ROW_LIST, optional: n_items pairs {weight set = output-channel tile * nclass + class, m-tile = tile row *
                              tiles_x + tile column}, in execution order (items of one weight set contiguous).  The class lists are then
                              CONCATENATED in rl_src / rl_out (each class padded to whole 16-row tiles on its own, rl_n = total number of
                              entries, tile row = index / 16) instead of [nclass][rl_n] padded to a common length, so a class with few
                              rows costs only the tiles it needs. */
    int32_t n_items;
    int32_t defer_wait;    /* any mode: != 0 declares that this launch reads nothing the previous launch in the stream writes and writes
                              nothing it reads or writes (a later pass of the same folded block: same inputs, other output pixels).  Its
                              CTAs then start on the SMs the previous grid's last round leaves idle (programmatic dependent launch
                              without the wait at the top) and wait for the previous grid only before exiting. */
    int32_t independent_steps; /* any mode: != 0 declares that the T steps of this call are INDEPENDENT samples, not a time sequence: the
                              membrane potential restarts from rest at every step (stateless inference only: h_seq, v_in, v_out, tsum
                              and stats must be NULL, 3 planes, v_reset 0).  A single-step call on a batch of k*B' samples -- the
                              reference's calling convention, one forward(x) per frame (SNN_models.py:152-192) -- is then issued as
                              T = k, B = B' on the same [k*B'][H][W][C] tensors: identical results, but a tile streams its weights
                              once for k patches instead of once per patch. */
} ss_tile_maps;
int ss_conv_i8_fwd_ex(const ss_block_desc* g, const ss_tile_maps* tm, const void* x, const void* w_i8, const float* wscale,
                      const float* decay, const float* v_in, float* v_out, const void* resid, void* out, float* h_seq,
                      void* tsum, void* stream);

/* Digit planes of ALREADY QUANTISED integer weights (fp32 holding exact integers, |q| < 2^(8*planes-1)), same layout as
 * ss_pack_weights_i8; zero_exp: device int32 [Cout] of zeros.  Used for the folded weight sets (sums of quantised taps).
 * More generally the values are quantised as round(q * 2^-zero_exp[n]): with fp32 weights and the exponents ss_pack_weights_folded
 * returns, this writes the plain 25-tap image of exactly the taps the folded sets are sums of (small calls of a folded block).
 * _rect: filters of ksy rows x ksx columns (OIHW [Cout][Cin][ksy][ksx]); ksy != ksx selects the 32-byte-row image of the
 * row-list pass. */
int ss_pack_digits_i8(const float* q_oihw, int32_t Cout, int32_t Cin, int32_t ks, int32_t planes, const int32_t* zero_exp,
                      void* w_i8, void* stream);
int ss_pack_digits_i8_rect(const float* q_oihw, int32_t Cout, int32_t Cin, int32_t ksy, int32_t ksx, int32_t planes,
                           const int32_t* zero_exp, void* w_i8, void* stream);

/* The three weight images of a folded NNConvUpsampling block (ss_conv_i8_fwd_ex: SS_TILES_FOLDED dense pass, the two
 * SS_TILES_ROW_LIST passes) straight from its fp32 5x5 OIHW weight, one launch: per output channel a power-of-two fixed point
 * with head-room for the tap sums (1 bit, one more while any folded sum overflows the top balanced digit), the quantised taps
 * summed per replication pattern (exact integers), digit planes written in the kernels' shared-memory layout.
 *   w_dense: 4 sets of 3x3 per output-channel tile (class = 2*row_class + col_class)  4*Cout*Cin*9*planes bytes
 *   w_rows / w_cols: 3 sets of 3x5 each (irregular rows; irregular columns, transposed frame)  3*Cout*Cin*15*planes bytes each
 *   wscale fp32 [Cout] = 2^e, wexp int32 [Cout] = e (workspace / by-product).   planes 2 or 3. */
int ss_pack_weights_folded(const float* w_oihw, int32_t Cout, int32_t Cin, int32_t planes, void* w_dense, void* w_rows,
                           void* w_cols, float* wscale, int32_t* wexp, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Same block on the CUDA cores in plain fp32 (exact fp32 weights, ascending-k accumulation): the first-layer
 * path for non-integer inputs and the on-device cross-check of the tensor-core path. */
typedef struct ss_conv_geom {
    int32_t T, B;
    int32_t Hin, Win, Cin;       /* source activation */
    int32_t Hout, Wout, Cout;    /* block output */
    int32_t ks;                  /* taps per axis; GEMM K = ks*ks*Cin, k = (ky*ks+kx)*Cin + c */
    int32_t in_layout;           /* SS_IN_* */
    int32_t neuron;              /* SS_NEURON_* */
    int32_t reserved0;
    float gain;                  /* MultiplyBy scalar */
    float v_th, v_reset, tau;
    int32_t reserved1, reserved2;
} ss_conv_geom;

/*   ymap [Hout*ks], xmap [Wout*ks] : source row / column read by output row oy at tap ky (resp. ox, kx),
 *                                    or -1 for a zero-padded tap.  One table pair expresses strided
 *                                    zero-padded convs and the nearest-neighbour-upsampled valid convs.
 *   x      : u8 [T][B][Hin][Win][Cin] (SS_IN_U8_TBHWC, Cin % 8 == 0) or fp32 [B][T][Cin][Hin][Win] (SS_IN_F32_BTCHW)
 *   w_kn   : fp32 [K][Cout], k ordered (ky,kx,c)
 *   other arguments as ss_conv_i8_fwd */
int ss_conv_neuron_fwd(const ss_conv_geom* g, const void* x, const int32_t* ymap, const int32_t* xmap,
                       const float* w_kn, const float* decay, const float* v_in, float* v_out, const void* resid,
                       void* out, float* h_seq, void* stream);

/* Prediction heads + I-neuron readout, forward, all T timesteps.
 * Replaces  Ineurons(predict_depthK(out_addK))  for K = 4,3,2,1 (SNN_models.py:133-150,172-188):
 * v += gain * (conv3x3(NNupsample(out_addK)) + bias_K), in that order, every timestep.
 *   acts[i] (u8 [T][B][Hs][Ws][C]), C[i], Hs[i], Ws[i], w[i] (fp32 [9][C]), bias[i] (device scalar),
 *   ymap[i] [H*3], xmap[i] [W*3] for i = 0..3 in execution order (head 4 first).
 *   taps[i]: fp32 workspace [T][B][9][Hs][Ws] ([2][B][9][Hs][Ws] suffices with acts_sum) -- the 9 per-tap channel dots are taken at SOURCE resolution
 *   (9*Hs*Ws*C MACs instead of 9*H*W*C) and gathered per output pixel; same math up to fp32 reassociation.
 *   v_io   : fp32 [B][H][W] I-neuron potential, updated in place (caller zero-fills / sets the prior)
 *   depths : fp32 [4][B][H][W]; depths[i] = potential right after head i of the LAST timestep
 */
typedef struct ss_heads_args {
    int32_t T, B, H, W;
    float gain;
    int32_t C[4], Hs[4], Ws[4];
    const void* acts[4];
    const float* w[4];
    const float* bias[4];
    const int32_t* ymap[4];
    const int32_t* xmap[4];
    float* taps[4];
    const void* acts_sum[4];     /* optional (all four or none): u8 [B][Hs][Ws][C] = sum of acts over the first T-1 timesteps
                                    (ss_conv_i8_fwd's `tsum` output).  The readout is linear and never fires, so the heads are
                                    then evaluated on that sum and on the last timestep only (2 passes instead of T). */
} ss_heads_args;
int ss_heads_fwd(const ss_heads_args* a, float* v_io, float* depths, void* stream);

/* Stand-alone neuron layer over T steps (spikingjelly neuron.*Node.forward on an arbitrary tensor, e.g. the
 * I-neuron pool of SNN_models.py:150).  x, s_out fp32 [T][N]; v_io fp32 [N] in/out; h_seq fp32 [T][N] or NULL. */
int ss_neuron_fwd(int32_t T, int64_t N, int32_t neuron, float v_th, float v_reset, float tau, const float* decay,
                  const float* x, float* v_io, float* s_out, float* h_seq, void* stream);

/* Surrogate-gradient scan (BPTT through the neuron), reverse time.  Replaces autograd through
 * BaseNode.forward with surrogate.{ATan,Sigmoid}.backward and detach_reset=True:
 *   g_h(t) = g_s(t) * ds/du(h_t - v_th) + g_v(t) * (1 - s_t);  g_x(t) = g_h(t) * r;  g_v(t-1) = g_h(t) * (1 - r)
 * with r = 1 (IF: g_v(t-1) = g_h(t)), 1/tau (LIF), *decay (PLIF).
 *   h_seq fp32 [T][N], g_s fp32 [T][N] (gradient w.r.t. the block's spikes), g_v_last fp32 [N] or NULL
 *   g_acc  fp32 [T][N]: gradient w.r.t. the conv result (gain folded in)
 *   g_decay: device scalar accumulated with atomicAdd (PLIF; gradient w.r.t. *decay), may be NULL
 */
int ss_neuron_bwd(int32_t T, int64_t N, int32_t neuron, int32_t surrogate, float alpha, float gain,
                  float v_th, float v_reset, float tau, const float* decay, const float* h_seq,
                  const float* v_init, const float* g_s, const float* g_v_last, float* g_acc,
                  float* g_v_init, float* g_decay, void* stream);
/* Same scan; additionally (or instead: g_acc may be NULL) writes g_acc rounded to bf16 [T][N], the operand of the
 * tensor-core gradient kernels below. */
int ss_neuron_bwd_ex(int32_t T, int64_t N, int32_t neuron, int32_t surrogate, float alpha, float gain,
                     float v_th, float v_reset, float tau, const float* decay, const float* h_seq,
                     const float* v_init, const float* g_s, const float* g_v_last, float* g_acc, void* g_acc_bf16,
                     float* g_v_init, float* g_decay, void* stream);

/* Convolution gradients of the fused block (replaces cuDNN dgrad / wgrad and upsample_nearest2d_backward
 * reached through autograd; the (ymap, xmap) tables fold the upsampling into the gather / scatter).
 *   g_acc fp32 [T][B][Hout][Wout][Cout] (output of ss_neuron_bwd)
 *   g_x   fp32 [T][B][Hin][Win][Cin], ACCUMULATED into (caller zero-fills or passes the skip-path gradient)
 *   g_w   fp32 [K][Cout], ACCUMULATED into
 */
int ss_conv_dgrad(const ss_conv_geom* g, const int32_t* ymap, const int32_t* xmap, const float* w_kn,
                  const float* g_acc, float* g_x, void* stream);
int ss_conv_wgrad(const ss_conv_geom* g, const void* x, const int32_t* ymap, const int32_t* xmap,
                  const float* g_acc, float* g_w, void* stream);

/* ---- tensor-core gradients (bf16 operands, fp32 accumulation; what train.py:239 reaches through cuDNN under autograd) ----
 *
 * ss_corr_bf16: stride-1 correlation of a zero-padded bf16 NHWC tensor with `nclass` weight sets, evaluated on a virtual
 * grid Hv x Wv and routed to an fp32 NHWC destination through per-axis output maps.  Every data gradient of the path is
 * one call:  same-padded 3x3 conv -> flipped/transposed weights, pad 1;  stride-2 5x5 conv -> four (row parity, column
 * parity) classes of 3x3 weights on the output-resolution gradient, destination (2i+py, 2j+px);  NNConvUpsampling ->
 * flipped 5x5 weights, pad 4, virtual grid = the upsampled image, destination = its nearest-neighbour source pixel
 * (several virtual pixels share one: SS_CORR_ATOMIC).  Same kernel as ss_conv_i8_fwd (halo patch, taps by descriptor
 * shift, T accumulator slots) with tcgen05 kind::f16 and a store epilogue.
 *   src_bf16 bf16 [T][B][Hg][Wg][Cg];  w_img from ss_pack_weights_bf16 (OIHW fp32 [nsets*ntile][Cg][ks][ks], set index =
 *   out-channel tile * nclass + class, class = 2*row_class + col_class);  ymap_out int32 [nclass > 1 ? 2 : 1][Hv], xmap_out
 *   likewise [..][Wv]: destination row / column or -1 (NULL, NULL = identity, needs nclass 1 and Hv x Wv == Hdst x Wdst);
 *   dst fp32 [T][B][Hdst][Wdst][Cdst]. */
#define SS_CORR_STORE 0       /* dst  = result (every destination element written exactly once)   */
#define SS_CORR_ACCUMULATE 1  /* dst += result, plain read-modify-write (destinations are unique) */
#define SS_CORR_ATOMIC 2      /* dst += result with red.global.add (destinations may repeat)       */
typedef struct ss_corr_desc {
    int32_t T, B;
    int32_t Hg, Wg, Cg;
    int32_t Hv, Wv;
    int32_t Hdst, Wdst, Cdst;
    int32_t ks, pad;
    int32_t nclass;      /* 1 or 4 */
    int32_t out_mode;    /* SS_CORR_* */
    int32_t ntile;       /* destination channels per weight set: 32 or 64 */
    int32_t reserved;
} ss_corr_desc;
int ss_pack_weights_bf16(const float* w_oihw, int32_t Cout, int32_t Cin, int32_t ks, int32_t ntile, void* w_img, void* stream);
int ss_corr_bf16(const ss_corr_desc* d, const void* src_bf16, const void* w_img, const int32_t* ymap_out,
                 const int32_t* xmap_out, float* dst, void* stream);

/* ss_conv_wgrad_bf16: weight gradient of one fused block (geometry as ss_conv_i8_fwd's descriptor; gain / neuron fields
 * ignored; planes == 0 declares that x may hold any u8 value -- event-count frames -- otherwise every x value must be < 128,
 * which holds for spikes and spike sums and selects the two-lane u8 -> bf16 conversion of the patch producers).  The contraction runs over pixels, the slow dimension of NHWC, so both MMA operands are MN-major: A = a tile of
 * g (pixels x 128 output channels), B = the halo patch of x converted to bf16 (pixels x 16, 32 or 64 input channels), one
 * accumulator per filter tap in TMEM (the taps are cut into groups of <= 512 / N handled by different CTAs); a tap is again a
 * start-address shift of the patch.
 *   x u8 [T][B][Hin][Win][Cin] (Cin % 16 == 0; Cin == 4: packed event frames);  g_bf16 bf16 [T][B][Hout][Wout][Cout];
 *   g_w fp32 [ks*ks*Cin][Cout] accumulated with atomics. */
int ss_conv_wgrad_bf16(const ss_block_desc* g, const void* x, const void* g_bf16, float* g_w, void* stream);

/* Backward of ss_heads_fwd (autograd through predict_depthK + the I-neuron running sum).
 *   g_depths fp32 [4][B][H][W]: gradient w.r.t. the four returned depth maps (execution order)
 *   g_acts[i] fp32 [T][B][Hs][Ws][C]: accumulated into, or (store_g_acts != 0) every element written -- the heads are the first
 *   consumer visited in the backward pass, so the caller can hand in uninitialised buffers (acts u8 as in ss_heads_fwd; taps unused);
 *   g_w[i] fp32 [9][C] accumulated; g_bias[i] scalar accumulated
 *   bins[i]   fp32 [2][B][Hs][Ws][9] zero-filled workspace
 */
int ss_heads_bwd(const ss_heads_args* a, const float* g_depths, float* const* g_acts, float* const* g_w,
                 float* const* g_bias, float* const* bins, int32_t store_g_acts, void* stream);

/* ---- loss + metric (first "next" row of the scope table: the step right after the path in the training loop) ----
 * Replaces network/loss.py:7-135 (Total_Loss = multi-scale scale-invariant loss + alpha * multi-scale Sobel gradient-matching
 * loss on the NaN-masked residual) and network/metrics.py:83-95 (MeanDepthError), train.py:238,257.
 *   pred[k] fp32 [B][H][W], k < nscale <= 4 (the four depth maps, all at full resolution); gt fp32 [B][H][W], NaN = invalid
 *   sums   device fp64 [nscale][5], zero-filled by the caller: n, sum r, sum r^2, sum(|gx| + |gy|), sum |r|   (r = pred - gt)
 *          => SI_k = S2/n - (S1/n)^2,  GM_k = G/n,  MDE_k = A/n
 *   signs  device u8 [nscale][B*H*W] written by the forward pass (signs of the Sobel responses), read by the backward pass
 *   coef_si / coef_gm  device fp32 [nscale] = weight of SI_k / GM_k in the total * upstream gradient;
 *   g_pred[k] fp32 [B][H][W] = d loss / d pred[k] */
int ss_loss_fwd(int32_t nscale, int32_t B, int32_t H, int32_t W, const float* const* pred, const float* gt, double* sums,
                void* signs, void* stream);
int ss_loss_bwd(int32_t nscale, int32_t B, int32_t H, int32_t W, const float* const* pred, const float* gt, const double* sums,
                const void* signs, const float* coef_si, const float* coef_gm, float* const* g_pred, void* stream);

int ss_abi_version(void);
const char* ss_last_error(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
int64_t ss_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
