// tcgen05 (5th-gen tensor core) implementation of the fused spiking block for sm_100a -- integer edition.
//
// Replaces: Conv2d / NNConvUpsampling -> MultiplyBy -> IF/LIF/PLIF node, once per timestep, plus the skip / SEW
// additions (reference network/SNN_models.py:75-129,171-186; network/blocks.py:110-132,145-171).
//
// Arithmetic.  Every conv input on this path is a small non-negative integer (spikes {0,1}, spike sums {0..3},
// event counts), so activations live in HBM as u8 NHWC and are EXACT.  fp32 weights are converted once to
// `planes` signed base-256 digits of a per-output-channel power-of-two fixed point (3 planes = 24 bits).  The
// contraction runs on the tensor cores as  u8 x s8 -> s32  (tcgen05.mma kind::i8, twice the bf16 MAC rate), one
// accumulator column block per digit plane stacked along the MMA N dimension; integer accumulation is exact and
// order independent, the epilogue recombines the planes in 64-bit arithmetic and rounds ONCE to fp32.  The result
// is the correctly rounded dot product of the quantised weights -- closer to the exact value than any fp32
// summation order, deterministic, and independent of the tiling.
//
// Data movement (implicit GEMM without im2col re-reads).  A CTA owns an output tile of 16 rows x 8 columns
// (M = 128) x 32 output channels.  For each 32/64-channel block of the input the producer warps gather the tile's
// receptive field ONCE into shared memory as a swizzled "halo patch" ([patch pixel][channel block], one swizzle
// row per pixel); every tap (ky,kx) of the filter is then just a different START ADDRESS of the same patch in the
// MMA's shared-memory descriptor (8-row groups = 8 consecutive pixels of one output row, stride-byte-offset = one
// patch row), so each input byte is fetched from L2 once per tile instead of ks*ks times.  Zero padding is
// zero-filled cp.async; stride 2 is a parity-split patch row (even | odd columns) with a 2-row group stride; the
// nearest-neighbour upsampling of the decoder is a gather table (the upsampled tensor is never materialised);
// the batch is stacked vertically so tiles straddle images and the 17x22 / 33x44 layers waste < 15 % of a tile.
//
// Time loop.  Weights dominate the traffic of the deep layers, so the T timesteps of a tile are processed against
// each weight block while it is resident: TMEM holds up to 512 / (planes*32) accumulator slots, one per timestep.
// The epilogue warps then run the neuron recurrence over those slots with the membrane potential in REGISTERS
// (it never touches HBM between timesteps), add the residual and store u8 spikes.  Layers whose weights fit in
// shared memory (<= 2 channel blocks) keep them resident across tiles and pipeline MMA(t+1) with epilogue(t).
//
// Warp roles (mbarrier pipelines, no __syncthreads in the steady state):
//   warps 0-3  patch producers (cp.async gather, zero fill)     warp 4  MMA issuer (one lane), owns TMEM
//   warp 5     weight producer (cp.async.bulk of pre-swizzled images)      warps 8-15  epilogue (TMEM -> neuron -> HBM)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "ss_common.cuh"
#include "ss_umma.cuh"

namespace ss {
namespace {

constexpr int NWB = 2;        // weight buffers

// tile iteration modes: SS_TILES_PLAIN / FOLDED / ROW_BANDS / COL_BANDS (include/stereospike_b200.h)
constexpr int MAX_STAGES = 8;
constexpr int MAX_SLOTS = 8;
constexpr int THREADS = 512;

// box heights of the TMA tensor maps: a run of patch rows inside one image is loaded as a binary decomposition of its length
constexpr int TMA_NMAPS = 6;    // 1, 2, 4, 8, 16, 32 rows

struct I8Params {
    // tensor maps of the activation tensor u8 [T*B][Hin][Win][Cin] (box = RB bytes x one patch row x 2^k rows), see the TMA producer
    alignas(64) CUtensorMap tmap[TMA_NMAPS];
    int tma;               // 1 = the halo patches are staged by cp.async.bulk.tensor (non-upsampled blocks), 0 = cp.async gathers
    int T, B, Hin, Win, Cin, Hout, Wout, Cout;
    int ks, stride, pad, upsample;
    int N;                 // planes * 32
    int RB, ncb, ntaps;
    int PH, PWp, PWhalf, ppix;
    int HsO, Hup, Wup;
    int tiles_x, mtiles, nitems;
    int mtiles2, nitems2;                          // CTA pairs: pairs of m-tiles per weight set, items per pair walk
    uint32_t m_mtiles2;
    uint32_t m_mtiles, m_tiles_x, m_per, m_hso;   // ceil(2^32 / d): exact unsigned division by multiply-high while n * d < 2^32 (0 = use '/')
    int Hv, Wv;            // iteration space of the tiles (== Hout, Wout except for the folded / band passes of an upsampled conv)
    int mode;              // SS_TILES_*
    int nclass;            // weight sets per output-channel tile (folded pass: 4 = {L,M} x {L,M}), else 1
    const int* ymap_out;   // folded: [2][Hv] real output row of virtual row s for class L, M, or -1
    const int* xmap_out;   // folded: [2][Wv]
    // row-list pass (SS_TILES_ROW_LIST): tile row r of m-tile row ty of class c is entry e = ty*16 + r of the class's list
    const int* rl_src;     // [nclass][rl_n] pixel offset (inside one timestep) of the entry's first source row at column 0, or -1
    const int* rl_out;     // [nclass][rl_n] pixel offset of the entry's output row at column 0, or -1
    const uint8_t* rl_collive;   // [c_nout] 1 = this output column belongs to the pass (NULL = all)
    int rl_n;              // entries per class (padded with -1 to a common length)
    const int2* item_tab;  // row-list passes, optional: item -> {weight set, m-tile}.  The class lists are then CONCATENATED (each padded
                           // to whole 16-row tiles on its own, rl_n = total entries) instead of padded to a common length, and a class
                           // gets exactly the tiles it needs
    int defer_wait;        // this launch does not read what the previous launch in the stream writes (2nd / 3rd pass of a folded block):
                           // it starts as soon as the previous grid leaves SMs free and only waits for it before exiting
    int rl_fold;           // row-list pass with FOLDED columns: tile columns = source positions, classes = (row class, column class),
                           // output column through xmap_out -- the dense folded pass restricted to the regular rows of each class
    int in_rowstep;        // pixels between the ROWSTEP source rows of an entry (Win; 1 for the transposed pass)
    int in_colpitch;       // pixels between neighbouring source columns (1; Win for the transposed pass)
    int out_colpitch;      // pixels between neighbouring output columns (1; Wout for the transposed pass)
    int c_in, c_up, c_nout;   // column axis: source size, upsampled size, outputs
    float c_scale;            // float(c_in) / c_up
    int TC, NPS, WB, PB;
    int resident;
    int nwb;               // weight buffers allocated in shared memory (1 when a single channel block is resident)
    float yscale, xscale;
    int neuron;
    float gain, v_th, v_reset, tau;
    const uint8_t* x;
    const int8_t* w;
    const float* wscale;
    const float* decay;
    const float* v_in;
    float* v_out;
    const uint8_t* resid;
    uint8_t* out;
    float* h_seq;
    uint8_t* tsum;
    unsigned long long* stats;   // optional [6]: {spikes, nonzero outputs, sum of out^2} over all steps, then over the last step
    float* g_dst;          // MODE_BF16 (gradient-side correlation): fp32 NHWC destination [T][B][Hout][Wout][Cout]
    int g_mode;            // SS_CORR_STORE / SS_CORR_ACCUMULATE / SS_CORR_ATOMIC
};

constexpr int MODE_I8 = 0;     // forward block: u8 activations x int8 weight digit planes -> s32, neuron epilogue
constexpr int MODE_BF16 = 1;   // gradient-side correlation: bf16 gradients x bf16 weights -> f32, store / accumulate / atomic epilogue


// Optional instrumentation build (-DSS_ROLE_TIMING, see tools/role_timing.py): per CTA and role, cycles spent waiting on
// each barrier class vs working, accumulated with clock64() and dumped through ss_debug_read().  Not part of the product .so.
#ifdef SS_ROLE_TIMING
__device__ unsigned long long ss_dbg[148 * 4 * 8];
#define SS_T0() const long long _t0 = clock64()
#define SS_ACC(role, slot) ss_acc[slot] += (unsigned long long)(clock64() - _t0)
#define SS_DECL() unsigned long long ss_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long _tstart = clock64()
#define SS_DUMP(role)                                                                              \
    do {                                                                                           \
        ss_acc[7] = (unsigned long long)(clock64() - _tstart);                                     \
        if (blockIdx.x < 148)                                                                      \
            for (int _i = 0; _i < 8; ++_i) ss_dbg[(blockIdx.x * 4 + (role)) * 8 + _i] = ss_acc[_i]; \
    } while (0)
#else
#define SS_T0()
#define SS_ACC(role, slot)
#define SS_DECL()
#define SS_DUMP(role)
#endif

// n / d for a loop-invariant divisor: one multiply-high instead of the ~35-instruction software division (the tile
// decoding sits on the producers' and the epilogue's per-item critical path, which is what bounds single-step calls)
__device__ __forceinline__ int fast_div(int n, int d, uint32_t magic) {
    return magic != 0u ? (int)__umulhi((uint32_t)n, magic) : n / d;
}
__host__ inline uint32_t div_magic(long long d, long long n_max) {
    // valid while n_max * d < 2^32 (error term n * (magic * d - 2^32) < 2^32)
    if (d <= 1 || n_max * d >= (1LL << 32)) return 0u;
    return (uint32_t)(((1ULL << 32) + (unsigned long long)d - 1ULL) / (unsigned long long)d);
}

// item -> (weight set, m-tile): round-robin division, or the table of a row-list pass with per-class tile counts
template <int ROWSTEP>
__device__ __forceinline__ void item_decode(const I8Params& p, int it, int mt_per, uint32_t m_mt_per, int& ntile, int& mt) {
    if constexpr (ROWSTEP > 1) {
        if (p.item_tab != nullptr) {
            const int2 e = __ldg(p.item_tab + it);
            ntile = e.x;
            mt = e.y;
            return;
        }
    }
    ntile = fast_div(it, mt_per, m_mt_per);
    mt = it - ntile * mt_per;
}

// ------------------------------------------------------------------------------------------------ geometry
// Pixel offset (inside one timestep) of the source row read by patch row `pr` of tile row `ty`, at column 0, or -1 (zero
// padding / gap between stacked images / dead list entry).
template <int STRIDE, int ROWSTEP>
__device__ __forceinline__ int row_source(const I8Params& p, int ty, int pr, int cls) {
    if constexpr (ROWSTEP > 1) {
        // row list: every tile row has its own ROWSTEP source rows (no sliding window between tile rows)
        const int e = ty * 16 + pr / ROWSTEP;
        if (e >= p.rl_n) return -1;
        const int s0 = __ldg(p.rl_src + (p.item_tab != nullptr ? 0 : (p.rl_fold ? cls >> 1 : cls) * p.rl_n) + e);
        return s0 < 0 ? -1 : s0 + (pr % ROWSTEP) * p.in_rowstep;
    } else {
        const int gi = ty * 16 * STRIDE + pr;
        const int per = STRIDE * p.HsO;
        const int b = fast_div(gi, per, p.m_per);
        const int local = gi - b * per;
        if (b >= p.B) return -1;
        if (p.upsample) {
            if (local >= p.Hup) return -1;
            // ATen upsample_nearest: min(int(floorf(dst * scale)), in - 1), scale = float(in) / out
            const int iy = min((int)floorf((float)local * p.yscale), p.Hin - 1);
            return (b * p.Hin + iy) * p.Win;
        }
        const int iy = local - p.pad;
        return (iy >= 0 && iy < p.Hin) ? (b * p.Hin + iy) * p.Win : -1;
    }
}
// Pixel offset along the row of patch column `pc` of tile column `tx`, or -1.
template <int STRIDE, int PWHALF, int ROWSTEP>
__device__ __forceinline__ int col_source(const I8Params& p, int tx, int pc) {
    if constexpr (ROWSTEP > 1) {
        const int u = tx * 8 + pc;             // column of the (virtual) upsampled image
        if (p.rl_fold) return u < p.c_in ? u * p.in_colpitch : -1;      // folded columns: the source column itself
        if (u >= p.c_up) return -1;
        return min((int)floorf((float)u * p.c_scale), p.c_in - 1) * p.in_colpitch;
    } else {
        if (p.upsample) {
            const int u = tx * 8 + pc;
            if (u >= p.Wup) return -1;
            return min((int)floorf((float)u * p.xscale), p.Win - 1);
        }
        int ix;
        if (STRIDE == 1) {
            ix = tx * 8 + pc - p.pad;
        } else {
            // parity-split row: [even-type columns | odd-type columns]; input col = 2*(ox0 - pad/2 + idx) + plane
            const int plane = pc / PWHALF;
            const int idx = pc - plane * PWHALF;
            ix = 2 * (tx * 8 - p.pad / 2 + idx) + plane;
        }
        return (ix >= 0 && ix < p.Win) ? ix : -1;
    }
}
// Stride-2 patches are stored PLANE-major: [even-type columns: PH rows x PWHALF pixels][odd-type columns: same], each plane padded
// to a multiple of 4 pixels (128 bytes at 32-byte rows) so that both planes are legal TMA destinations; a plane row is then a
// dense run of PWHALF pixels, which is exactly what one row of a tensor-map box with element stride 2 along W delivers.
__host__ __device__ constexpr int plane_pixels(int PH, int PWHALF) { return (PH * PWHALF + 3) / 4 * 4; }
// patch pixel at which the A operand of tap (ky,kx) starts
template <int STRIDE, int PWP, int PWHALF, int PH>
__device__ __forceinline__ constexpr int tap_offset(int ky, int kx) {
    return STRIDE == 1 ? ky * PWP + kx : (kx & 1) * plane_pixels(PH, PWHALF) + ky * PWHALF + (kx >> 1);
}

// Issues every MMA of one (patch stage, weight buffer) pair except the very first one (tap 0, k-step 0), which the
// caller issues itself because it carries the run-time accumulate flag.
template <int MODE, int KS, int KSX, int STRIDE, int RB, int CN, int PWP, int PWHALF, int PH, int TAP = 0, int K = 1>
__device__ __forceinline__ void issue_taps(uint32_t d, uint64_t a0, uint64_t b0, uint32_t idesc) {
    constexpr int KSTEPS = RB / 32;
    if constexpr (TAP < KS * KSX) {
        if constexpr (K < KSTEPS) {
            constexpr int ky = TAP / KSX, kx = TAP % KSX;
            constexpr uint32_t aoff = (uint32_t)(tap_offset<STRIDE, PWP, PWHALF, PH>(ky, kx) * RB + K * 32) >> 4;
            constexpr uint32_t boff = (uint32_t)(TAP * CN * RB + K * 32) >> 4;
            umma_i8_off<aoff, boff, MODE>(d, a0, b0, idesc);
            issue_taps<MODE, KS, KSX, STRIDE, RB, CN, PWP, PWHALF, PH, TAP, K + 1>(d, a0, b0, idesc);
        } else {
            issue_taps<MODE, KS, KSX, STRIDE, RB, CN, PWP, PWHALF, PH, TAP + 1, 0>(d, a0, b0, idesc);
        }
    }
}

// FIRST: the first layer (Cin <= 4, event-count frames u8 [T][B][H][W][4]).  Its K = ks*ks*4 <= 128 is one swizzle row,
// so the producers assemble an explicit im2col tile (KS = 1 "tap", RB = 128: row = output pixel, byte = tap*4 + c) with
// L1-cached 4-byte loads instead of a halo patch -- 4 MMAs per tile instead of ks*ks*(32-channel padded blocks).
// PAIR: two CTAs of a 2-cluster process two m-tiles against the same weight set with one tcgen05.mma.cta_group::2 (M = 256):
// each CTA stages its own patch and HALF of the weight rows, so the shared-memory operand feed per MMA drops from 7 KB to
// 5.5 KB per SM (N = 96) and the MMA leaves the feed-bound regime.  Rank 0 issues; rank 1 relays its producers' barriers.
// KSX / ROWSTEP: rectangular filters (KS rows x KSX columns) and the row-list pass of a folded NNConvUpsampling block, whose
// tile rows are arbitrary (sample, output row) entries with ROWSTEP private source rows each (patch row = ROWSTEP * tile row + ky).
// NK: neuron kind fixed at compile time (and v_reset == 0), -1 = run-time switch.  The epilogue is instruction-issue bound on the
// full-resolution blocks; the generic version executes the other kinds' arithmetic predicated off (~25 % of its instructions).
// LEAN: stateless inference instantiation -- no saved potentials, no carried / returned membrane state, no firing statistics (the
// launcher picks it when none of those pointers is given).  The epilogue-bound blocks are limited by dependent-issue latency under
// a 126-register allocation; without the h_seq / v_out / statistics paths the compiler has 16+ fewer live values to keep.
// LEAN = 2: the T "timesteps" of an item are INDEPENDENT samples (ss_tile_maps.independent_steps): the potential restarts from
// rest at every step.  A single-step call on a batch -- the reference's calling convention, forward(x) once per frame -- then
// runs as T' = k steps of B / k samples: one weight stream per tile serves k patches instead of one (the deep blocks of a T = 1
// call are bound by re-streaming their weights from L2 for every tile).
template <int PLANES, int KS, int STRIDE, int RB, bool FIRST = false, int MODE = MODE_I8, bool PAIR = false, int KSX = KS, int ROWSTEP = 1,
          int NK = -1, int LEAN = 0>
__global__ void __launch_bounds__(THREADS, 1) conv_i8_kernel(const __grid_constant__ I8Params p) {
    // compile-time geometry: every descriptor offset of the MMA issue loop folds to an immediate
    constexpr int cN = PLANES * 32;
    constexpr int cNTAPS = KS * KSX;
    constexpr int cPWhalf = 8 + (KSX - 1) / 2;
    constexpr int cPWp = STRIDE == 1 ? 8 + KSX - 1 : 2 * cPWhalf;
    constexpr int cPH = ROWSTEP > 1 ? 16 * ROWSTEP : 15 * STRIDE + KS;
    static_assert(ROWSTEP == 1 || (STRIDE == 1 && ROWSTEP >= KS && !FIRST && !PAIR && MODE == MODE_I8), "row-list pass: stride-1 int8 blocks");
    static_assert(cPH <= 48 && cPWp <= 24, "geometry tables");
    constexpr int cPPIX = cPH * cPWp;                        // pixels of the patch that carry data
    constexpr int cPLANE = plane_pixels(cPH, cPWhalf);       // stride 2: pixels per parity plane (padded)
    constexpr int cPALLOC = STRIDE == 1 ? cPPIX : 2 * cPLANE;
    constexpr int cNB = PAIR ? cN / 2 : cN;                 // weight rows (of the MMA's N) held by this CTA
    constexpr int cWB = cNTAPS * cNB * RB;
    constexpr int MMA_MODE = PAIR ? 2 : MODE;
    static_assert(!PAIR || (MODE == MODE_I8 && !FIRST && cN % 16 == 0), "CTA pairs: int8 forward blocks only");
    constexpr int cPB = (cPALLOC * RB + 1023) / 1024 * 1024;
    constexpr int cTC = (512 / cN) < MAX_SLOTS ? (512 / cN) : MAX_SLOTS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);

    const uint32_t w_base = base;
    const uint32_t patch_base = base + (uint32_t)p.nwb * cWB;
    uint8_t* tail = sm + (size_t)p.nwb * cWB + (size_t)p.NPS * cPB;
    int* rowsrc = reinterpret_cast<int*>(tail);              // [2][48]
    int* colsrc = rowsrc + 96;                               // [2][24]
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 640);
    // bars: full_p[8], empty_p[8], full_w[2], empty_w[2], full_a[8], empty_a[8], tok[2]
    // (+ raw_full[2], raw_empty[2] of the first-layer producers at bars + 38)
    // (+ CTA pairs: peer_full_p[8] at bars + 42, peer_full_w[2] at bars + 50: the peer CTA's producers, relayed to rank 0)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 52);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t bar_full_p = smem_u32(bars);
    const uint32_t bar_empty_p = smem_u32(bars + 8);
    const uint32_t bar_full_w = smem_u32(bars + 16);
    const uint32_t bar_empty_w = smem_u32(bars + 18);
    const uint32_t bar_full_a = smem_u32(bars + 20);
    const uint32_t bar_empty_a = smem_u32(bars + 28);
    const uint32_t bar_tok = smem_u32(bars + 36);
    const uint32_t bar_peer_p = smem_u32(bars + 42);
    const uint32_t bar_peer_w = smem_u32(bars + 50);
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    // item walk: a pair shares the weight set and takes m-tiles 2q, 2q+1
    const int it0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int its = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int mt_per = PAIR ? p.mtiles2 : p.mtiles;            // items per weight set
    const uint32_t m_mt_per = PAIR ? p.m_mtiles2 : p.m_mtiles;
    const int nit = PAIR ? p.nitems2 : p.nitems;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) {
            mbar_init(bar_full_p + 8 * s, FIRST ? 64 : (p.tma ? 1 : 128));   // TMA: one arrive.expect_tx, the bytes do the rest
            mbar_init(bar_empty_p + 8 * s, 1);
        }
        for (int s = 0; s < NWB; ++s) {
            mbar_init(bar_full_w + 8 * s, 1);
            mbar_init(bar_empty_w + 8 * s, 1);
        }
        for (int s = 0; s < MAX_SLOTS; ++s) {
            mbar_init(bar_full_a + 8 * s, 1);
            mbar_init(bar_empty_a + 8 * s, PAIR ? 16 : 8);      // one arrival per epilogue warp (of both CTAs of a pair)
        }
        mbar_init(bar_tok, 1);
        mbar_init(bar_tok + 8, 1);
        for (int s = 0; s < 4; ++s) mbar_init(smem_u32(bars + 38 + s), 64);
        for (int s = 0; s < 10; ++s) mbar_init(bar_peer_p + 8 * s, 1);
        fence_barrier_init();
    }
    if (warp == 4) {
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();      // the peer's barriers are initialised before anybody arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation) overlapped the tail of the previous
    // kernel in the stream; from here on we read what it wrote.  Let our own dependents start their prologue early as well.
    // A later pass of the same folded block reads only what the first pass read (the previous layer's output, complete before the
    // first pass got past its own wait) and writes other pixels of the same tensors, so it does not wait here: its CTAs take the SMs
    // the previous pass's last round leaves idle (a pass of 192 items on 148 SMs is two rounds with the second two-thirds empty).
    // It waits before exiting instead, so that whoever waits for THIS grid still waits for everything before it.
    if (!p.defer_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    constexpr uint32_t swz_mask = (uint32_t)(RB >> 4) - 1u;  // 32 -> 1, 64 -> 3, 128 -> 7

    if (warp < 4 && FIRST) {
        // ================================================================== first-layer im2col producers (two steps)
        // The im2col row of a pixel is 25 four-byte taps; copying them one by one (25 cp.async per pixel and timestep) made
        // this layer producer-bound.  Instead warps 0-1 stage the tile's raw 20 x 12-pixel halo patch (960 B) with 8-byte
        // copies into a small double buffer, and warps 2-3 expand it into the swizzled 128-byte im2col rows with ordinary
        // shared-memory loads / 16-byte stores (bytes 100..127 = 0).  Only the expanding warps fence towards the async proxy,
        // and they never have global loads in flight, so the fence does not wait for memory.
        constexpr int PH5 = 20, PW5 = 12;            // halo patch of the real 5x5 filter (the template's KS is the single im2col "tap")
        uint8_t* raw = tail + 1152;                  // [2][1024]
        const uint32_t bar_raw_full = smem_u32(bars + 38);
        const uint32_t bar_raw_empty = smem_u32(bars + 40);
        const size_t t_stride = (size_t)p.B * p.Hin * p.Win * 4;
        if (warp < 2) {
            const int tid = threadIdx.x;             // 0..63
            // two pixels per copy when every pair stays 8-byte aligned and on one side of the image border
            const bool pair = (p.Win & 1) == 0 && (p.pad & 1) == 0;
            const int upr = pair ? PW5 / 2 : PW5;    // copies per patch row
            const int nunits = PH5 * upr;
            int urow[4], useg[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int u = tid + i * 64;
                urow[i] = u < nunits ? u / upr : -1;
                useg[i] = u - (u / upr) * upr;
            }
            int rb = 0, itcount = 0;
            uint32_t rphase = 0;
            for (int it = it0; it < nit; it += its, ++itcount) {
                const int mt = it % p.mtiles;
                const int ty = fast_div(mt, p.tiles_x, p.m_tiles_x), tx = mt - ty * p.tiles_x;
                int* rs = rowsrc + (itcount & 1) * 48;
                int* cs = colsrc + (itcount & 1) * 24;
                if (tid < PH5) rs[tid] = row_source<1, 1>(p, ty, tid, 0);
                if (tid >= 32 && tid - 32 < PW5) cs[tid - 32] = col_source<1, 10, 1>(p, tx, tid - 32);
                named_sync(1, 64);
                int goff[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    goff[i] = -2;
                    if (urow[i] >= 0) {
                        const int r = rs[urow[i]], c = cs[pair ? 2 * useg[i] : useg[i]];
                        goff[i] = (r >= 0 && c >= 0) ? (r + c) * 4 : -1;
                    }
                }
                for (int t = 0; t < p.T; ++t) {
                    mbar_wait(bar_raw_empty + 8 * rb, rphase ^ 1u);
                    const uint8_t* xt = p.x + (size_t)t * t_stride;
                    const uint32_t dst = smem_u32(raw) + (uint32_t)rb * 1024u;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (goff[i] != -2) {
                            const bool ok = goff[i] >= 0;
                            const uint32_t d = dst + (uint32_t)(urow[i] * (PW5 * 4) + useg[i] * (pair ? 8 : 4));
                            if (pair) cp_async_8(d, ok ? xt + goff[i] : p.x, ok ? 8u : 0u);
                            else cp_async_4(d, ok ? xt + goff[i] : p.x, ok ? 4u : 0u);
                        }
                    }
                    cp_async_arrive_noinc(bar_raw_full + 8 * rb);
                    rb ^= 1;
                    if (rb == 0) rphase ^= 1u;
                }
            }
        } else {
            const int bt = threadIdx.x - 64;         // 0..63: expands tile pixels bt and bt + 64
            int stage = 0, rb = 0;
            uint32_t phase = 0, rphase = 0;
            for (int it = it0; it < nit; it += its) {
                for (int t = 0; t < p.T; ++t) {
                    mbar_wait(bar_raw_full + 8 * rb, rphase);
                    mbar_wait(bar_empty_p + 8 * stage, phase ^ 1u);
                    const uint8_t* rw = raw + rb * 1024;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = bt + h * 64;
                        const int g = r >> 3, j = r & 7;
                        uint32_t w[32];
#pragma unroll
                        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                            for (int kx = 0; kx < 5; ++kx)
                                w[ky * 5 + kx] = *reinterpret_cast<const uint32_t*>(rw + ((g + ky) * PW5 + j + kx) * 4);
#pragma unroll
                        for (int k = 25; k < 32; ++k) w[k] = 0u;
                        uint8_t* row = sm + (size_t)p.nwb * cWB + (size_t)stage * cPB + (size_t)r * 128;
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            *reinterpret_cast<uint4*>(row + ((c ^ (r & 7)) << 4)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
                    }
                    fence_proxy_async();      // generic-proxy stores -> visible to the MMA's async-proxy reads
                    mbar_arrive(bar_full_p + 8 * stage);
                    mbar_arrive(bar_raw_empty + 8 * rb);
                    if (++stage == p.NPS) {
                        stage = 0;
                        phase ^= 1u;
                    }
                    rb ^= 1;
                    if (rb == 0) rphase ^= 1u;
                }
            }
        }
    } else if (warp < 4 && p.tma) {
        // ================================================================== patch producer, TMA edition (non-upsampled blocks)
        // The activation tensor is a 4-d tensor map u8 [T*B][Hin][Win][Cin]; a halo patch is a box of it: RB channel bytes x one
        // patch row of pixels (element stride 2 along W for the parity planes of a stride-2 conv) x a run of rows, written by the
        // TMA unit straight into the swizzled layout the MMA descriptors read (the hardware swizzle of cp.async.bulk.tensor is a
        // function of the absolute shared-memory address, like the UMMA one: tools/tma_probe.cu).  Zero padding, the right / bottom
        // image borders and the batch tail are the tensor map's out-of-bounds zero fill.  The vertically stacked batch makes a
        // patch straddle images, and the box height is a property of the map, so the rows of one image are loaded as the binary
        // decomposition of their count (maps with 1, 2, 4 .. 32 rows).  One elected thread issues everything: ~2-8 instructions
        // per stage instead of 128 threads x 6-18 cp.async.
        if constexpr (ROWSTEP > 1) {
            // Row list with folded columns (rl_fold): every tile row owns ROWSTEP consecutive source rows of one image, i.e. one box
            // of tmap[0] (ROWSTEP rows high); a dead list entry is loaded from an out-of-range image = zeros.
            // One elected lane per producer warp issues the boxes of 4 tile rows (a single thread issuing all 16 small boxes was
            // slower than the gather producers); warp 0's lane also posts the stage's transaction count.
            if (elect_one()) {
                constexpr uint32_t ROWBYTES = (uint32_t)(cPWp * RB);
                constexpr uint32_t STAGE_TX = (uint32_t)cPH * ROWBYTES;
                const int n_oob = p.T * p.B;
                int stage = 0;
                uint32_t phase = 0;
                for (int it = it0; it < nit; it += its) {
                    int ntile, mt;
                    item_decode<ROWSTEP>(p, it, mt_per, m_mt_per, ntile, mt);
                    const int ty = fast_div(mt, p.tiles_x, p.m_tiles_x), tx = mt - ty * p.tiles_x;
                    const int rcls = p.rl_fold ? (ntile % p.nclass) >> 1 : ntile % p.nclass;
                    const int rl_base = p.item_tab != nullptr ? 0 : rcls * p.rl_n;
                    int ent[4];                       // (image << 16 | source row) of this thread's tile rows, -1 = dead
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int e = ty * 16 + warp * 4 + k;
                        const int s0 = e < p.rl_n ? __ldg(p.rl_src + rl_base + e) : -1;
                        ent[k] = -1;
                        if (s0 >= 0) {
                            const int row = s0 / p.Win;           // b * Hin + y
                            const int b = row / p.Hin;
                            ent[k] = (b << 16) | (row - b * p.Hin);
                        }
                    }
                    const int x0 = tx * 8;
                    for (int t0 = 0; t0 < p.T; t0 += cTC) {
                        const int tc = min(cTC, p.T - t0);
                        const int n_outer = p.resident ? tc : p.ncb;
                        const int n_inner = p.resident ? p.ncb : tc;
                        for (int o = 0; o < n_outer; ++o) {
                            for (int in = 0; in < n_inner; ++in) {
                                const int cb = p.resident ? in : o;
                                const int t = t0 + (p.resident ? o : in);
                                mbar_wait(bar_empty_p + 8 * stage, phase ^ 1u);
                                const uint32_t full = bar_full_p + 8 * stage;
                                if (warp == 0) mbar_arrive_expect_tx(full, STAGE_TX);
                                const uint32_t dst0 = patch_base + (uint32_t)stage * cPB + (uint32_t)(warp * 4 * ROWSTEP) * ROWBYTES;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const int v = ent[k];
                                    const int n = v >= 0 ? t * p.B + (v >> 16) : n_oob;
                                    tma_load_4d(dst0 + (uint32_t)(k * ROWSTEP) * ROWBYTES, &p.tmap[0], cb * RB, x0, v >= 0 ? (v & 0xFFFF) : 0, n, full);
                                }
                                if (++stage == p.NPS) {
                                    stage = 0;
                                    phase ^= 1u;
                                }
                            }
                        }
                    }
                }
            }
        } else
        if (warp == 0 && elect_one()) {
            constexpr int ROWPIX = STRIDE == 1 ? cPWp : cPWhalf;              // pixels per patch row (of one plane)
            constexpr uint32_t ROWBYTES = (uint32_t)(ROWPIX * RB);
            constexpr bool EVEN = ROWBYTES % 128u != 0u;                        // a box must start on 128 bytes: even rows only
            constexpr uint32_t STAGE_TX = (uint32_t)((STRIDE == 1 ? 1 : 2) * cPH) * ROWBYTES;
            const int per = STRIDE * p.HsO;
            const int n_oob = p.T * p.B;
            int stage = 0;
            uint32_t phase = 0;
            SS_DECL();
            for (int it = it0; it < nit; it += its) {
                const int ntile = fast_div(it, mt_per, m_mt_per);
                const int mt = PAIR ? 2 * (it - ntile * mt_per) + (int)crank : it - ntile * mt_per;
                const int ty = fast_div(mt, p.tiles_x, p.m_tiles_x), tx = mt - ty * p.tiles_x;
                const int x0 = STRIDE == 1 ? tx * 8 - p.pad : 2 * (tx * 8 - p.pad / 2);
                const int gi0 = ty * 16 * STRIDE;
                const int b0 = fast_div(gi0, per, p.m_per);
                const int local0 = gi0 - b0 * per;
                for (int t0 = 0; t0 < p.T; t0 += cTC) {
                    const int tc = min(cTC, p.T - t0);
                    const int n_outer = p.resident ? tc : p.ncb;
                    const int n_inner = p.resident ? p.ncb : tc;
                    for (int o = 0; o < n_outer; ++o) {
                        for (int in = 0; in < n_inner; ++in) {
                            const int cb = p.resident ? in : o;
                            const int t = t0 + (p.resident ? o : in);
                            {
                                SS_T0();
                                mbar_wait(bar_empty_p + 8 * stage, phase ^ 1u);
                                SS_ACC(0, 1);     // wait for a free patch stage
                            }
                            SS_T0();
                            const uint32_t full = bar_full_p + 8 * stage;
                            mbar_arrive_expect_tx(full, STAGE_TX);
                            const uint32_t dst0 = patch_base + (uint32_t)stage * cPB;
                            const int c0 = cb * RB;
                            int r = 0, b = b0, local = local0;
                            while (r < cPH) {
                                // rows [r, r + L) of the patch = rows local .. of image b (its zero padding included)
                                int L = min(per - local, cPH - r);
                                // the next image's first row is zero padding from either side: split one row later when that makes
                                // the next run start on an even row
                                if (EVEN && ((r + L) & 1) && r + L < cPH) ++L;
                                const int n = b < p.B ? t * p.B + b : n_oob;
                                int rr = r, yy = local - p.pad;
#pragma unroll
                                for (int lg = TMA_NMAPS - 1; lg >= 0; --lg) {
                                    if (L & (1 << lg)) {
                                        const uint32_t dst = dst0 + (uint32_t)rr * ROWBYTES;
                                        tma_load_4d(dst, &p.tmap[lg], c0, x0, yy, n, full);
                                        if constexpr (STRIDE == 2) tma_load_4d(dst + (uint32_t)(cPLANE * RB), &p.tmap[lg], c0, x0 + 1, yy, n, full);
                                        rr += 1 << lg;
                                        yy += 1 << lg;
                                    }
                                }
                                r += L;
                                local += L;
                                if (local >= per) {
                                    local -= per;
                                    ++b;
                                }
                            }
                            SS_ACC(0, 2);         // copy issue
                            if (++stage == p.NPS) {
                                stage = 0;
                                phase ^= 1u;
                            }
                        }
                    }
                }
            }
            SS_DUMP(0);
        }
        __syncwarp();
    } else if (warp < 4) {
        // ================================================================== patch producers
        // Work unit = one 16-byte chunk of one patch pixel.  Consecutive lanes take consecutive chunks of consecutive INPUT
        // pixels (for the parity-split stride-2 rows: alternating planes), so one warp-level cp.async reads a contiguous
        // 512-byte run of the NHWC tensor instead of 32 separate half-sectors.
        const int tid = threadIdx.x;
        const size_t t_stride = (size_t)p.B * p.Hin * p.Win * p.Cin;
        constexpr int chunks = RB >> 4;
        constexpr int cUNITS = cPPIX * chunks;
        constexpr int cNU = (cUNITS + 127) / 128;     // units per producer thread
        int stage = 0;
        uint32_t phase = 0;
        int itcount = 0;
        // loop-invariant part of every unit: its patch row / column and its swizzled shared-memory offset
        int urc[cNU];            // patch row * 32 + patch column, or -1 past the end of the patch
        uint32_t soff[cNU];
#pragma unroll
        for (int i = 0; i < cNU; ++i) {
            const int u = tid + i * 128;
            const int pixl = u / chunks, ch = u - pixl * chunks;
            int pr = pixl / cPWp;
            int pc = pixl - pr * cPWp;
            int spix = pr * cPWp + pc;                                 // pixel of the patch in shared memory
            if (STRIDE == 2) {
                // neighbours in memory = alternating parity planes; the planes are stored one after the other
                spix = (pc & 1) * cPLANE + pr * cPWhalf + (pc >> 1);
                pc = (pc & 1) * cPWhalf + (pc >> 1);
            }
            soff[i] = swizzle_off((uint32_t)(spix * RB + ch * 16), swz_mask);
            urc[i] = u < cUNITS ? pr * 32 + pc : -1;
        }
        SS_DECL();
        for (int it = it0; it < nit; it += its, ++itcount) {
            SS_T0();
            int ntile, mt;
            item_decode<ROWSTEP>(p, it, mt_per, m_mt_per, ntile, mt);
            if constexpr (PAIR) mt = 2 * mt + (int)crank;   // (an odd tail tile is all padding)
            const int ty = fast_div(mt, p.tiles_x, p.m_tiles_x), tx = mt - ty * p.tiles_x;
            int* rs = rowsrc + (itcount & 1) * 48;
            int* cs = colsrc + (itcount & 1) * 24;
            if (tid < cPH) rs[tid] = row_source<STRIDE, ROWSTEP>(p, ty, tid, ntile % p.nclass);
            if (tid >= 64 && tid - 64 < cPWp) cs[tid - 64] = col_source<STRIDE, cPWhalf, ROWSTEP>(p, tx, tid - 64);
            named_sync(1, 128);
            int goff[cNU];
#pragma unroll
            for (int i = 0; i < cNU; ++i) {
                goff[i] = -2;
                if (urc[i] >= 0) {
                    const int r = rs[urc[i] >> 5], c = cs[urc[i] & 31];
                    const int ch = (tid + i * 128) % chunks;
                    goff[i] = (r >= 0 && c >= 0) ? (r + c) * p.Cin + ch * 16 : -1;
                }
            }
            SS_ACC(0, 0);     // geometry
            for (int t0 = 0; t0 < p.T; t0 += cTC) {
                const int tc = min(cTC, p.T - t0);
                const int n_outer = p.resident ? tc : p.ncb;
                const int n_inner = p.resident ? p.ncb : tc;
                for (int o = 0; o < n_outer; ++o) {
                    for (int in = 0; in < n_inner; ++in) {
                        const int cb = p.resident ? in : o;
                        const int t = t0 + (p.resident ? o : in);
                        {
                            SS_T0();
                            mbar_wait(bar_empty_p + 8 * stage, phase ^ 1u);
                            SS_ACC(0, 1);     // wait for a free patch stage
                        }
                        SS_T0();
                        const uint8_t* xt = p.x + (size_t)t * t_stride + cb * RB;
                        const uint32_t dst0 = patch_base + (uint32_t)stage * cPB;
#pragma unroll
                        for (int i = 0; i < cNU; ++i) {
                            if (goff[i] != -2) {
                                const bool ok = goff[i] >= 0;
                                cp_async_16(dst0 + soff[i], ok ? xt + goff[i] : p.x, ok ? 16u : 0u);
                            }
                        }
                        // the barrier arrival fires when this thread's copies have landed; nobody blocks here
                        cp_async_arrive_noinc(bar_full_p + 8 * stage);
                        SS_ACC(0, 2);         // copy issue
                        if (++stage == p.NPS) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
        if (threadIdx.x == 0) SS_DUMP(0);
    } else if (warp == 4 || warp == 6) {
        // ================================================================== MMA issuers: two elected threads (warps 4 and 6) ping-pong
        // The tensor pipe's issue queue is only ~3 MMAs deep and the bookkeeping between two stages (barrier waits, proxy
        // fence, descriptor set-up) costs ~500 cycles, which left the pipe idle ~330 cycles per 1400-cycle stage with a single
        // issuing thread (measured with the SS_MMA_TIMING build).  So stage g is issued by thread g & 1: while one thread's
        // 25 MMAs drain, the other has already waited for its patch, fenced and built its descriptors, and only waits for the
        // "issued" token of its predecessor.  Both threads walk the same loop nest and keep identical phase bookkeeping.
        if (PAIR && crank != 0u) {
            // ---------------------------------------------------------- peer CTA of a pair: no MMA issue here.  One thread of
            // warp 4 forwards "patch stage filled", one of warp 6 "weight half loaded" to rank 0, in the producers' order.
            if (elect_one()) {
                if (warp == 4) {
                    int stage = 0;
                    uint32_t phase = 0;
                    for (int it = it0; it < nit; it += its)
                        for (int k = 0; k < p.T * p.ncb; ++k) {
                            mbar_wait(bar_full_p + 8 * stage, phase);
                            fence_proxy_async();     // our cp.async writes -> the tensor core's async-proxy reads of OUR shared memory
                            mbar_arrive_cluster(mapa_u32(bar_peer_p + 8 * stage, 0u));
                            if (++stage == p.NPS) {
                                stage = 0;
                                phase ^= 1u;
                            }
                        }
                } else {
                    uint32_t wu = 0;
                    int loaded_ntile = -1;
                    for (int it = it0; it < nit; it += its) {
                        const int ntile = fast_div(it, mt_per, m_mt_per);
                        if (p.resident) {
                            if (ntile != loaded_ntile) {
                                loaded_ntile = ntile;
                                ++wu;
                                for (int cb = 0; cb < p.ncb; ++cb) {
                                    mbar_wait(bar_full_w + 8 * cb, (wu - 1u) & 1u);
                                    mbar_arrive_cluster(mapa_u32(bar_peer_w + 8 * cb, 0u));
                                }
                            }
                        } else {
                            for (int t0 = 0; t0 < p.T; t0 += cTC)
                                for (int cb = 0; cb < p.ncb; ++cb) {
                                    const int buf = (int)(wu % NWB);
                                    mbar_wait(bar_full_w + 8 * buf, (wu / NWB) & 1u);
                                    mbar_arrive_cluster(mapa_u32(bar_peer_w + 8 * buf, 0u));
                                    ++wu;
                                }
                        }
                    }
                }
            }
        } else if (elect_one()) {
            const uint32_t role = warp == 4 ? 0u : 1u;
            // instruction descriptor: i8 = s32 accumulate, A u8, B s8;  bf16 = f32 accumulate, A and B bf16; both operands K-major;
            // a CTA pair runs M = 256 (128 rows per CTA)
            constexpr uint32_t idesc = (MODE == MODE_I8 ? ((2u << 4) | (0u << 7) | (1u << 10)) : ((1u << 4) | (1u << 7) | (1u << 10))) |
                                       ((uint32_t)(cN >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
            constexpr uint32_t layout = RB == 128 ? 2u : (RB == 64 ? 4u : 6u);
            // 8-row groups = 8 pixels of one output row; the next output row is ROWSTEP / STRIDE patch rows further (of one plane)
            constexpr uint32_t a_sbo = (uint32_t)((ROWSTEP > 1 ? ROWSTEP * cPWp : (STRIDE == 1 ? cPWp : 2 * cPWhalf)) * RB);
            constexpr uint32_t b_sbo = (uint32_t)(8 * RB);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t slot_phase = 0;   // bit s: parity to wait on empty_a[s] (starts "free")
            uint32_t wu = 0;           // streaming: weight-use counter;  resident: number of loads done
            int loaded_ntile = -1;
            uint32_t w_pending = 0;    // resident: bit cb set while full_w[cb] has not been observed for the current load
            uint32_t g = 0;            // global stage counter
            uint32_t sbase = 0;        // running accumulator-slot count
            SS_DECL();
            uint32_t mine = 0;         // stages issued by this thread so far
            const uint32_t tok_wait = bar_tok + 8 * (role ^ 1u);   // predecessor's "issued" token
            const uint32_t tok_post = bar_tok + 8 * role;
            const uint64_t a_stage0 = make_desc(patch_base, a_sbo, layout);
            const uint64_t b_buf0 = make_desc(w_base, b_sbo, layout);

            // completion of everything issued so far -> a barrier (of both CTAs of a pair, at the same shared-memory offset)
            auto commit = [&](uint32_t bar) {
                if constexpr (PAIR) umma_commit_pair(bar);
                else umma_commit(bar);
            };
            // hand-backs from the epilogue warps arrive from both CTAs of a pair
            auto wait_slot = [&](uint32_t bar, uint32_t parity) {
                if constexpr (PAIR) mbar_wait_cluster(bar, parity);
                else mbar_wait(bar, parity);
            };
            // returns true when this thread issued the stage
            auto do_stage = [&](int wbuf, int slot, bool first) -> bool {
                const bool own = (g & 1u) == role;
                ++g;
                if (own) {
                    {
                        SS_T0();
                        mbar_wait(bar_full_p + 8 * stage, phase);
                        if constexpr (PAIR) mbar_wait_cluster(bar_peer_p + 8 * stage, phase);   // ... and for the peer's
                        SS_ACC(1, 0);         // wait for the patch
                    }
                    fence_proxy_async();   // cp.async wrote the patch through the generic proxy; the MMA reads it through the async proxy
                    const uint64_t a0 = a_stage0 + (uint64_t)((uint32_t)stage * (uint32_t)(cPB >> 4));
                    const uint64_t b0 = b_buf0 + (uint64_t)((uint32_t)wbuf * (uint32_t)(cWB >> 4));
                    const uint32_t d = tmem_base + (uint32_t)(slot * cN);
                    // everything is ready: wait until the other thread has issued the previous stage
                    {
                        SS_T0();
                        mbar_wait(tok_wait, role == 0 ? ((mine & 1u) ^ 1u) : (mine & 1u));
                        SS_ACC(1, 1);         // wait for the other issuer's token
                    }
                    ++mine;
                    SS_T0();
                    tc_fence_after();
                    umma_i8<MMA_MODE>(d, a0, b0, idesc, first ? 0u : 1u);
                    issue_taps<MMA_MODE, KS, KSX, STRIDE, RB, cNB, cPWp, cPWhalf, cPH>(d, a0, b0, idesc);
                    tc_fence_before();
                    mbar_arrive(tok_post);
                    commit(bar_empty_p + 8 * stage);
                    SS_ACC(1, 2);             // issue
                }
                if (++stage == p.NPS) {
                    stage = 0;
                    phase ^= 1u;
                }
                return own;
            };

            for (int it = it0; it < nit; it += its) {
                int ntile, mt_unused;                               // weight-set index: (output-channel tile, class)
                item_decode<ROWSTEP>(p, it, mt_per, m_mt_per, ntile, mt_unused);
                if (p.resident && ntile != loaded_ntile) {
                    loaded_ntile = ntile;
                    w_pending = (1u << p.ncb) - 1u;
                    ++wu;
                }
                for (int t0 = 0; t0 < p.T; t0 += cTC) {
                    const int tc = min(cTC, p.T - t0);
                    // Accumulator slots rotate across items (slot = running count mod cTC) instead of restarting at 0: with a
                    // short sequence (T = 1: the reference's forward() contract) the next tile's MMAs then run into a free slot
                    // while the epilogue is still busy with the previous one.  T == cTC: every item uses slots 0..T-1 as before.
                    if (p.resident) {
                        for (int s0 = 0; s0 < tc; ++s0) {
                            const int s = (int)((sbase + (uint32_t)s0) % (uint32_t)cTC);
                            // BOTH threads wait for every slot hand-back (and every weight fill below), not only the thread that
                            // issues into it: a parity wait is only meaningful if the waiter is at most one phase behind.
                            {
                                SS_T0();
                                wait_slot(bar_empty_a + 8 * s, ((slot_phase >> s) & 1u) ^ 1u);
                                SS_ACC(1, 3);     // wait for a free accumulator slot
                            }
                            slot_phase ^= 1u << s;
                            bool last_own = false;
                            for (int cb = 0; cb < p.ncb; ++cb) {
                                if (w_pending & (1u << cb)) {
                                    mbar_wait(bar_full_w + 8 * cb, (wu - 1u) & 1u);
                                    if constexpr (PAIR) mbar_wait_cluster(bar_peer_w + 8 * cb, (wu - 1u) & 1u);
                                }
                                w_pending &= ~(1u << cb);
                                last_own = do_stage(cb, s, cb == 0);
                            }
                            if (last_own) commit(bar_full_a + 8 * s);
                        }
                    } else {
                        for (int cb = 0; cb < p.ncb; ++cb) {
                            const int buf = (int)(wu % NWB);
                            {
                                SS_T0();
                                mbar_wait(bar_full_w + 8 * buf, (wu / NWB) & 1u);
                                if constexpr (PAIR) mbar_wait_cluster(bar_peer_w + 8 * buf, (wu / NWB) & 1u);
                                SS_ACC(1, 4);     // wait for streamed weights
                            }
                            bool last_own = false;
                            for (int s0 = 0; s0 < tc; ++s0) {
                                const int s = (int)((sbase + (uint32_t)s0) % (uint32_t)cTC);
                                if (cb == 0) {
                                    SS_T0();
                                    wait_slot(bar_empty_a + 8 * s, ((slot_phase >> s) & 1u) ^ 1u);
                                    SS_ACC(1, 3);
                                    slot_phase ^= 1u << s;
                                }
                                last_own = do_stage(buf, s, cb == 0);
                                if (cb == p.ncb - 1 && last_own) commit(bar_full_a + 8 * s);
                            }
                            if (last_own) commit(bar_empty_w + 8 * buf);
                            ++wu;
                        }
                    }
                    sbase += (uint32_t)tc;
                }
                if (p.resident) {
                    const int nxt = it + its;
                    // the thread that issued the last stage releases the resident weight buffers before a reload
                    int nxt_ntile = ntile, nxt_mt;
                    if (nxt < nit) item_decode<ROWSTEP>(p, nxt, mt_per, m_mt_per, nxt_ntile, nxt_mt);
                    if (nxt < nit && nxt_ntile != ntile && ((g - 1u) & 1u) == role)
                        for (int cb = 0; cb < p.ncb; ++cb) commit(bar_empty_w + 8 * cb);
                }
            }
            SS_DUMP(1 + (int)role);
        }
        __syncwarp();
    } else if (warp == 5) {
        // ================================================================== weight producer (bulk copies)
        if (lane == 0) {
            uint32_t wu = 0;
            int loaded_ntile = -1;
            auto load = [&](int buf, int ntile, int cb) {
                const uint32_t full = bar_full_w + 8 * buf;
                mbar_arrive_expect_tx(full, (uint32_t)cWB);
                const uint32_t dst = w_base + (uint32_t)buf * cWB;
                if constexpr (PAIR) {
                    // this CTA's half of the N rows of every tap (the packed image is [tap][cN rows][RB]; the swizzle of a row
                    // only depends on row mod 8 groups, which the 48-row split preserves)
                    const int8_t* src = p.w + (size_t)(ntile * p.ncb + cb) * (size_t)(cNTAPS * cN * RB) + (size_t)crank * (cNB * RB);
                    for (int tap = 0; tap < cNTAPS; ++tap)
                        bulk_load(dst + (uint32_t)(tap * cNB * RB), src + (size_t)tap * (cN * RB), (uint32_t)(cNB * RB), full);
                } else {
                    const int8_t* src = p.w + (size_t)(ntile * p.ncb + cb) * cWB;
                    for (int o = 0; o < cWB; o += 16384) bulk_load(dst + o, src + o, (uint32_t)min(16384, cWB - o), full);
                }
            };
            for (int it = it0; it < nit; it += its) {
                int ntile, mt_unused;
                item_decode<ROWSTEP>(p, it, mt_per, m_mt_per, ntile, mt_unused);
                if (p.resident) {
                    if (ntile != loaded_ntile) {
                        for (int cb = 0; cb < p.ncb; ++cb) {
                            if (wu > 0) mbar_wait(bar_empty_w + 8 * cb, (wu - 1u) & 1u);
                            load(cb, ntile, cb);
                        }
                        loaded_ntile = ntile;
                        ++wu;
                    }
                } else {
                    for (int t0 = 0; t0 < p.T; t0 += cTC) {
                        for (int cb = 0; cb < p.ncb; ++cb) {
                            const int buf = (int)(wu % NWB);
                            if (wu >= NWB) mbar_wait(bar_empty_w + 8 * buf, ((wu / NWB) - 1u) & 1u);
                            load(buf, ntile, cb);
                            ++wu;
                        }
                    }
                }
            }
        }
    } else if (warp >= 8) {
        // ================================================================== epilogue: TMEM -> neuron -> HBM
        const int ew = warp - 8;
        const int quarter = ew & 3;   // == warp % 4: the TMEM lanes this warp may touch
        const int hf = ew >> 2;       // which 16 of the tile's 32 output channels
        const int m = quarter * 32 + lane;
        const int g = m >> 3, j = m & 7;
        const int M = p.B * p.Hout * p.Wout;
        if constexpr (MODE == MODE_BF16) {
            // ---------------------------------------------------------- gradient-side epilogue: fp32 accumulators -> HBM
            // Thread = one virtual output pixel (TMEM lane) x half of the tile's cN output channels.  The pixel is routed
            // through the (class, virtual position) -> destination maps (stride-2 parity classes, nearest-neighbour
            // upsampling); several virtual pixels may share a destination (upsampling), hence the atomic mode.
            constexpr int HALF = cN / 2;
            constexpr int NCH = HALF / 16;
            const size_t t_out = (size_t)M * p.Cout;
            uint32_t slot_phase = 0;
            uint32_t sbase = 0;
            for (int it = it0; it < nit; it += its) {
                const int wset = fast_div(it, p.mtiles, p.m_mtiles);
                const int mt = it - wset * p.mtiles;
                const int ty = fast_div(mt, p.tiles_x, p.m_tiles_x), tx = mt - ty * p.tiles_x;
                const int cls = wset % p.nclass;
                const int so = ty * 16 + g;
                const int b = fast_div(so, p.HsO, p.m_hso);
                int oy = so - b * p.HsO;
                int ox = tx * 8 + j;
                bool live = oy < p.Hv && ox < p.Wv && b < p.B;
                if (p.mode == SS_TILES_FOLDED) {
                    oy = live ? __ldg(p.ymap_out + (cls >> 1) * p.Hv + oy) : -1;
                    ox = live ? __ldg(p.xmap_out + (cls & 1) * p.Wv + ox) : -1;
                    live = oy >= 0 && ox >= 0;
                }
                live = live && oy < p.Hout && ox < p.Wout;
                const int nb = (wset / p.nclass) * cN + hf * HALF;
                const size_t o0 = live ? ((size_t)(b * p.Hout + oy) * p.Wout + ox) * p.Cout + nb : 0;
                for (int t0 = 0; t0 < p.T; t0 += cTC) {
                    const int tc = min(cTC, p.T - t0);
                    for (int s0 = 0; s0 < tc; ++s0) {
                        const int s = (int)((sbase + (uint32_t)s0) % (uint32_t)cTC);      // rotating slot, as in the MMA role
                        float* dst = p.g_dst + (size_t)(t0 + s0) * t_out + o0;
                        // accumulate mode: fetch the destination BEFORE blocking on the accumulator, so the round trip to
                        // HBM overlaps the MMAs of this slot instead of stalling the adds behind it
                        float4 oldv[NCH * 4];
                        if (live && p.g_mode == SS_CORR_ACCUMULATE) {
#pragma unroll
                            for (int q = 0; q < NCH * 4; ++q) oldv[q] = *reinterpret_cast<const float4*>(dst + q * 4);
                        } else {
#pragma unroll
                            for (int q = 0; q < NCH * 4; ++q) oldv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        mbar_wait(bar_full_a + 8 * s, (slot_phase >> s) & 1u);
                        slot_phase ^= 1u << s;
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * cN + hf * HALF);
                        int d[NCH][16];
#pragma unroll
                        for (int c = 0; c < NCH; ++c) tmem_ld16(taddr + c * 16, d[c]);
                        tmem_ld_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_empty_a + 8 * s);
                        if (!live) continue;
#pragma unroll
                        for (int c = 0; c < NCH; ++c)
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float4 v = make_float4(__int_as_float(d[c][4 * q]), __int_as_float(d[c][4 * q + 1]),
                                                       __int_as_float(d[c][4 * q + 2]), __int_as_float(d[c][4 * q + 3]));
                                float4* dp = reinterpret_cast<float4*>(dst + c * 16 + q * 4);
                                if (p.g_mode == SS_CORR_ATOMIC) {
                                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                                                 : "memory");
                                } else {
                                    const float4 old = oldv[c * 4 + q];       // zeros in store mode
                                    v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
                                    *dp = v;
                                }
                            }
                    }
                    sbase += (uint32_t)tc;
                }
            }
        } else {
        NeuronConst nc;
        nc.gain = p.gain; nc.v_th = p.v_th; nc.v_reset = p.v_reset; nc.tau = p.tau;
        nc.rtau = div_const_prepare(p.tau);
        nc.decay = (p.neuron == SS_NEURON_PLIF) ? __ldg(p.decay) : 0.0f;
        const size_t t_out = (size_t)M * p.Cout;
        uint32_t slot_phase = 0;
        uint32_t sbase = 0;           // running accumulator-slot count (same sequence as in the MMA role)
        const uint32_t empty_a_remote = PAIR ? mapa_u32(bar_empty_a, 0u) : 0u;
        SS_DECL();
        int sc_ntile = -1;
        float sc[16];                 // wscale * gain (wscale is a power of two, so this product is exact)
        // firing statistics (SNN_models.py:194-245, loss.py:96-107) straight from the registers that hold the spikes: per thread
        // {spikes, nonzero outputs, sum out^2} over all steps [0..2] and over the last step [3..5]; one atomic per warp at the end
        uint32_t st[6] = {0u, 0u, 0u, 0u, 0u, 0u};
        for (int it = it0; it < nit; it += its) {
            int ntile, mt;
            item_decode<ROWSTEP>(p, it, mt_per, m_mt_per, ntile, mt);
            if constexpr (PAIR) mt = 2 * mt + (int)crank;
            const int ty = fast_div(mt, p.tiles_x, p.m_tiles_x), tx = mt - ty * p.tiles_x;
            const int wset = ntile;                       // (output-channel tile, class)
            const int cls = wset % p.nclass;
            size_t pix = 0;
            bool live;
            if constexpr (ROWSTEP > 1) {
                // row list: the tile row is an entry of this class's list; columns are plain output columns
                const int e = ty * 16 + g;
                const int orow = e < p.rl_n ? __ldg(p.rl_out + (p.item_tab != nullptr ? 0 : (p.rl_fold ? cls >> 1 : cls) * p.rl_n) + e) : -1;
                int oc = tx * 8 + j;
                if (p.rl_fold) oc = oc < p.Wv ? __ldg(p.xmap_out + (cls & 1) * p.Wv + oc) : -1;   // source position -> output column of the class
                live = orow >= 0 && oc >= 0 && oc < p.c_nout && (p.rl_collive == nullptr || __ldg(p.rl_collive + oc) != 0);
                if (live) pix = (size_t)orow + (size_t)oc * p.out_colpitch;
            } else {
                const int so = ty * 16 + g;
                const int b = fast_div(so, p.HsO, p.m_hso);
                int oy = so - b * p.HsO;
                int ox = tx * 8 + j;
                live = oy < p.Hv && ox < p.Wv;
                if (p.mode == SS_TILES_FOLDED) {
                    // virtual (class, source position) -> real output pixel; -1 = not a regular position of this class
                    oy = live ? __ldg(p.ymap_out + (cls >> 1) * p.Hv + oy) : -1;
                    ox = live ? __ldg(p.xmap_out + (cls & 1) * p.Wv + ox) : -1;
                    live = oy >= 0 && ox >= 0;
                }
                live = live && b < p.B && oy < p.Hout && ox < p.Wout;
                if (live) pix = (size_t)(b * p.Hout + oy) * p.Wout + ox;
            }
            const int nb = (wset / p.nclass) * 32 + hf * 16;
            const size_t o0 = pix * p.Cout + nb;                  // element offset inside one timestep
            const bool use_resid = p.resid != nullptr && live;
            uint4 rs_next = make_uint4(0u, 0u, 0u, 0u);
            if (use_resid) rs_next = __ldg(reinterpret_cast<const uint4*>(p.resid + o0));
            uint32_t ts[4] = {0u, 0u, 0u, 0u};   // running byte-wise sum of the first T-1 output steps (feeds the linear heads)
            float v[16];
            if (ntile != sc_ntile) {
                sc_ntile = ntile;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(p.wscale + nb) + i);
                    sc[4 * i] = q.x * nc.gain; sc[4 * i + 1] = q.y * nc.gain; sc[4 * i + 2] = q.z * nc.gain; sc[4 * i + 3] = q.w * nc.gain;
                }
            }
            if (!LEAN && p.v_in != nullptr && live) {
                const float4* vi = reinterpret_cast<const float4*>(p.v_in + o0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 q = __ldg(vi + i);
                    v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = NK >= 0 ? 0.0f : p.v_reset;
            }
            for (int t0 = 0; t0 < p.T; t0 += cTC) {
                const int tc = min(cTC, p.T - t0);
                for (int s0 = 0; s0 < tc; ++s0) {
                    const int t = t0 + s0;
                    const int s = (int)((sbase + (uint32_t)s0) % (uint32_t)cTC);
                    // residual of this step was requested one step ago; request the next one before blocking
                    const uint4 rs_cur = rs_next;
                    if (use_resid && t + 1 < p.T) rs_next = __ldg(reinterpret_cast<const uint4*>(p.resid + (size_t)(t + 1) * t_out + o0));
                    {
                        SS_T0();
                        mbar_wait(bar_full_a + 8 * s, (slot_phase >> s) & 1u);
                        SS_ACC(3, 0);         // wait for the accumulator
                    }
                    slot_phase ^= 1u << s;
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * cN + hf * 16);
                    int d[PLANES][16];
#pragma unroll
                    for (int pl = 0; pl < PLANES; ++pl) tmem_ld16(taddr + pl * 32, d[pl]);
                    tmem_ld_wait();
                    tc_fence_before();
                    // the slot is in every lane's registers now: one lane hands it back to the MMA threads (rank 0's barrier for a
                    // CTA pair -- a remote arrival per thread made the paired epilogue slower than the single-CTA one)
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (PAIR) mbar_arrive_cluster(empty_a_remote + 8 * s);
                        else mbar_arrive(bar_empty_a + 8 * s);
                    }
                    if (!live) continue;
                    float x[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        // recombine the base-256 digit planes exactly, round once to fp32, then scale and gain
                        float conv;
                        if (PLANES == 2) {
                            conv = __int2float_rn(d[0][i] * 256 + d[1][i]);
                        } else if (PLANES == 3) {
                            // (splitting S = (hi >> 11) * 2^19 + ((hi & 2047) * 256 + d2) into two exact fp32 conversions + one FMA is
                            //  bit-identical and avoids the 64-bit conversion, but measured 0-10 % slower: I2F.S64 is not the bound)
                            conv = __ll2float_rn((long long)(d[0][i] * 256 + d[1][i]) * 256LL + (long long)d[2][i]);
                        } else {
                            conv = __ll2float_rn((long long)(d[0][i] * 256 + d[1][i]) * 65536LL +
                                                 (long long)(d[PLANES - 2][i] * 256 + d[PLANES - 1][i]));
                        }
                        x[i] = __fmul_rn(conv, sc[i]);   // == (conv * wscale) * gain: the first product is exact
                    }
                    float hbuf[16];
                    uint32_t sb[16];      // spike of each channel as 0 / 1
                    if constexpr (LEAN == 2) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = 0.0f;      // independent steps: every step starts from rest (v_reset == 0)
                    }
                    if constexpr (NK >= 0) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) sb[i] = neuron_step_t<NK, true>(x[i], v[i], nc, hbuf[i]) != 0.0f ? 1u : 0u;
                    } else if (p.neuron == SS_NEURON_IF) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) sb[i] = neuron_step_t<SS_NEURON_IF>(x[i], v[i], nc, hbuf[i]) != 0.0f ? 1u : 0u;
                    } else if (p.neuron == SS_NEURON_LIF) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) sb[i] = neuron_step_t<SS_NEURON_LIF>(x[i], v[i], nc, hbuf[i]) != 0.0f ? 1u : 0u;
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) sb[i] = neuron_step_t<SS_NEURON_PLIF>(x[i], v[i], nc, hbuf[i]) != 0.0f ? 1u : 0u;
                    }
                    // 16 spikes -> 16 bytes (+ residual bytes; sums stay <= 3, no carry between bytes)
                    uint32_t pk[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        pk[q] = sb[4 * q] | (sb[4 * q + 1] << 8) | (sb[4 * q + 2] << 16) | (sb[4 * q + 3] << 24);
                    const size_t o = (size_t)t * t_out + o0;
                    uint32_t n_spk = 0u;
                    if (!LEAN && p.stats != nullptr) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) n_spk = __dp4a(pk[q], 0x01010101u, n_spk);
                    }
                    pk[0] += rs_cur.x; pk[1] += rs_cur.y; pk[2] += rs_cur.z; pk[3] += rs_cur.w;
                    if (!LEAN && p.stats != nullptr) {
                        uint32_t n_nz = 0u, n_sq = 0u;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            n_nz = __dp4a((pk[q] | (pk[q] >> 1)) & 0x01010101u, 0x01010101u, n_nz);   // bytes are 0..3
                            n_sq = __dp4a(pk[q], pk[q], n_sq);
                        }
                        st[0] += n_spk; st[1] += n_nz; st[2] += n_sq;
                        if (t + 1 == p.T) { st[3] += n_spk; st[4] += n_nz; st[5] += n_sq; }
                    }
                    *reinterpret_cast<uint4*>(p.out + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    if (t + 1 < p.T) {
                        ts[0] += pk[0]; ts[1] += pk[1]; ts[2] += pk[2]; ts[3] += pk[3];
                    }
                    if (!LEAN && p.h_seq != nullptr) {
                        float4* hp = reinterpret_cast<float4*>(p.h_seq + o);
#pragma unroll
                        for (int i = 0; i < 4; ++i) hp[i] = make_float4(hbuf[4 * i], hbuf[4 * i + 1], hbuf[4 * i + 2], hbuf[4 * i + 3]);
                    }
                }
                sbase += (uint32_t)tc;
            }
            if (p.tsum != nullptr && live) *reinterpret_cast<uint4*>(p.tsum + o0) = make_uint4(ts[0], ts[1], ts[2], ts[3]);
            if (!LEAN && p.v_out != nullptr && live) {
                float4* vo = reinterpret_cast<float4*>(p.v_out + o0);
#pragma unroll
                for (int i = 0; i < 4; ++i) vo[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
        }
        if (!LEAN && p.stats != nullptr) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                uint32_t a = st[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0 && a != 0u) atomicAdd(p.stats + k, (unsigned long long)a);
            }
        }
        if (threadIdx.x == 256) SS_DUMP(3);
        }   // MODE_I8
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();      // rank 0's MMAs read the peer's shared memory and write its TMEM until the very end
    if (warp == 4) {
        tc_fence_after();
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
    if (p.defer_wait && threadIdx.x == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ weight packing
// per output channel: exponent e such that every |w| * 2^-e fits `planes` balanced base-256 digits
__global__ void __launch_bounds__(256) weight_exponent_kernel(const float* __restrict__ w, int Cout, int per_out, int planes,
                                                              float* __restrict__ wscale, int* __restrict__ wexp) {
    const int n = blockIdx.x;
    float m = 0.0f;
    for (int i = threadIdx.x; i < per_out; i += blockDim.x) m = fmaxf(m, fabsf(w[(size_t)n * per_out + i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) m = fmaxf(m, red[i]);
        int e = 0;
        if (m > 0.0f) {
            int ex;
            frexpf(m, &ex);                    // m = f * 2^ex, f in [0.5, 1)
            e = ex - (8 * planes - 1);         // |q| < 2^(8*planes-1)
            // balanced digits carry upwards: the top digit must stay <= 127 for the largest magnitude
            const long long q = llrint(ldexp((double)m, -e));
            long long rest = q;
            for (int pl = planes - 1; pl > 0; --pl) {
                const long long dgt = ((rest + 128) & 255) - 128;
                rest = (rest - dgt) >> 8;
            }
            if (rest > 127) ++e;
        }
        wexp[n] = e;
        wscale[n] = ldexpf(1.0f, e);
    }
}

// [ntile][cb][tap][N = planes x 32 rows][RB bytes], swizzled exactly as it must sit in shared memory
__global__ void __launch_bounds__(256) weight_pack_kernel(const float* __restrict__ w, int Cout, int Cin, int ks, int ksx, int planes, int RB,
                                                          const int* __restrict__ wexp, int8_t* __restrict__ out) {
    const int ntaps = ks * ksx;
    const int ncb = Cin / RB;
    const long long total = (long long)Cout * Cin * ntaps;   // one thread per weight
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    // decompose idx as (ntile, cb, tap, r32, c)
    long long r = idx;
    const int c = (int)(r % RB); r /= RB;
    const int r32 = (int)(r % 32); r /= 32;
    const int tap = (int)(r % ntaps); r /= ntaps;
    const int cb = (int)(r % ncb); r /= ncb;
    const int ntile = (int)r;
    const int n = ntile * 32 + r32;
    const int ch = cb * RB + c;
    const int ky = tap / ksx, kx = tap - ky * ksx;
    const float wv = w[(((size_t)n * Cin + ch) * ks + ky) * ksx + kx];   // OIHW (ks rows x ksx columns)
    long long q = llrint(ldexp((double)wv, -wexp[n]));
    const int N = planes * 32;
    const size_t buf = (size_t)(ntile * ncb + cb) * ((size_t)ntaps * N * RB);
    const uint32_t mask = (uint32_t)(RB >> 4) - 1u;
    for (int pl = planes - 1; pl >= 0; --pl) {
        long long dgt;
        if (pl > 0) {
            dgt = ((q + 128) & 255) - 128;
            q = (q - dgt) >> 8;
        } else {
            dgt = q;
        }
        const uint32_t off = (uint32_t)((tap * N + pl * 32 + r32) * RB + c);
        out[buf + (off ^ (((off >> 7) & mask) << 4))] = (int8_t)dgt;
    }
}

// first layer (Cin <= 4): [ntile][planes x 32 rows][128 B], byte k = (ky*5+kx)*4 + c, bytes 100..127 zero
__global__ void __launch_bounds__(256) weight_pack_first_kernel(const float* __restrict__ w, int Cout, int Cin, int planes,
                                                                const int* __restrict__ wexp, int8_t* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Cout * 128) return;
    const int k = idx & 127;
    const int n = idx >> 7;
    const int ntile = n >> 5, r32 = n & 31;
    const int tap = k >> 2, c = k & 3;
    long long q = 0;
    if (tap < 25 && c < Cin) q = llrint(ldexp((double)w[((size_t)n * Cin + c) * 25 + tap], -wexp[n]));
    const int N = planes * 32;
    const size_t buf = (size_t)ntile * N * 128;
    for (int pl = planes - 1; pl >= 0; --pl) {
        long long dgt;
        if (pl > 0) {
            dgt = ((q + 128) & 255) - 128;
            q = (q - dgt) >> 8;
        } else {
            dgt = q;
        }
        const uint32_t off = (uint32_t)((pl * 32 + r32) * 128 + k);
        out[buf + (off ^ (((off >> 7) & 7u) << 4))] = (int8_t)dgt;
    }
}

// fp32 NCHW event-count frames [B][T][C][H][W] (reference layout, train.py:201-218) -> u8 NHWC [T][B][H][W][Cpad], channels >= C zero.
// One thread = PX consecutive pixels x one group of 4 channels: PX = 4 reads each channel plane with 16-byte loads (the scalar
// version moved 72 MB at 1.9 TB/s).  Cpad = 4: the first-layer im2col mode; Cpad = 32, 64: the channel-concatenated temporal mode
// (train.py:206-218: first conv with 2 * nfpdm * cameras input channels), which runs as an ordinary 32-byte-row block.
template <int PX>
__global__ void __launch_bounds__(256) pack_events_kernel(const float* __restrict__ x, int B, int T, int C, int Cpad, int H, int W,
                                                          uint32_t* __restrict__ out, int* __restrict__ status) {
    const long long HW = (long long)H * W;
    const int groups = Cpad >> 2;
    const long long npg = HW / PX;                        // pixel groups per frame
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)T * B * npg * groups) return;
    const int grp = (int)(idx % groups);
    const long long pg = (idx / groups) % npg;
    const int b = (int)((idx / (groups * npg)) % B);
    const int t = (int)(idx / (groups * npg * B));
    uint32_t word[PX];
#pragma unroll
    for (int i = 0; i < PX; ++i) word[i] = 0u;
    bool bad = false;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        const int c = grp * 4 + cc;
        if (c >= C) break;
        const float* src = x + (((size_t)b * T + t) * C + c) * HW + pg * PX;
        float f[PX];
        if constexpr (PX == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(src));
            f[0] = q.x; f[1] = q.y; f[2] = q.z; f[3] = q.w;
        } else {
            f[0] = __ldg(src);
        }
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            const float r = fminf(fmaxf(rintf(f[i]), 0.0f), 255.0f);
            bad |= (r != f[i]);
            word[i] |= (uint32_t)r << (8 * cc);
        }
    }
    const size_t o = ((((size_t)t * B + b) * HW + pg * PX) * groups) + grp;     // in 4-byte words
#pragma unroll
    for (int i = 0; i < PX; ++i) out[o + (size_t)i * groups] = word[i];
    if (bad && status != nullptr) atomicOr(status, 1);
}

// bf16 weight image of the gradient-side correlation: [wset = Cout / NT][cb][tap][NT rows][RB bytes = RB/2 input channels],
// swizzled exactly as it must sit in shared memory (same image layout as the int8 one, 2-byte elements)
__global__ void __launch_bounds__(256) weight_pack_bf16_kernel(const float* __restrict__ w, int Cout, int Cin, int ks, int NT, int RB,
                                                               uint16_t* __restrict__ out) {
    const int ntaps = ks * ks;
    const int cpr = RB / 2;                                   // input channels per swizzle row
    const int ncb = Cin / cpr;
    const long long total = (long long)Cout * Cin * ntaps;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    long long r = idx;
    const int c = (int)(r % cpr); r /= cpr;
    const int rn = (int)(r % NT); r /= NT;
    const int tap = (int)(r % ntaps); r /= ntaps;
    const int cb = (int)(r % ncb); r /= ncb;
    const int wset = (int)r;
    const int n = wset * NT + rn;
    const int ch = cb * cpr + c;
    const int ky = tap / ks, kx = tap - ky * ks;
    const float wv = w[(((size_t)n * Cin + ch) * ks + ky) * ks + kx];   // OIHW
    const size_t buf = (size_t)(wset * ncb + cb) * ((size_t)ntaps * NT * RB);
    const uint32_t mask = (uint32_t)(RB >> 4) - 1u;
    const uint32_t off = (uint32_t)((tap * NT + rn) * RB + c * 2);
    out[(buf + (off ^ (((off >> 7) & mask) << 4))) >> 1] = __bfloat16_as_ushort(__float2bfloat16_rn(wv));
}

// ------------------------------------------------------------------------------------------------ TMA tensor maps (host)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// the driver API entry point is fetched through the runtime (the library does not link libcuda)
TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TensorMapEncodeFn>(f);
        cudaGetLastError();
    }
    return fn;
}

struct TmapKey {
    const void* x;
    long long nimg;
    int Hin, Win, Cin, RB, stride, ks, fixed_rows;
    bool operator==(const TmapKey& o) const {
        return x == o.x && nimg == o.nimg && Hin == o.Hin && Win == o.Win && Cin == o.Cin && RB == o.RB && stride == o.stride && ks == o.ks &&
               fixed_rows == o.fixed_rows;
    }
};
struct TmapEntry {
    TmapKey key;
    CUtensorMap maps[TMA_NMAPS];
    bool valid;
};

// Fills p.tmap / p.tma for a non-upsampled block whose patches can be staged by TMA (see the producer): x = u8 [nimg][Hin][Win][Cin]
// (Cin in BYTES), RB-byte channel blocks, patch rows of PWp pixels (stride 1) or PWhalf pixels per parity plane (stride 2).
// Encoding a map costs a few microseconds on the host, so the maps of the last calls are kept (the activation buffers of a model
// come back at the same addresses from the caching allocator; a graph capture bakes them into the launch anyway).
void setup_tma(I8Params& p, const void* x, long long nimg, int Hin, int Win, int Cin, int RB, int stride, int ks, int pad,
               int fixed_rows = 0) {
    // fixed_rows > 0: a single map (tmap[0]) whose box is that many rows high (row-list pass: one box per tile row)
    p.tma = 0;
    static int tma_env = -1;
    if (tma_env < 0) {
        const char* e = getenv("SS_TMA");
        tma_env = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (tma_env == 0) return;
    const int rowpix = stride == 1 ? 8 + ks - 1 : 8 + (ks - 1) / 2;
    const int rowbytes = rowpix * RB;
    // every box lands on a 128-byte boundary: any row when the row pitch allows it, else even rows only, which needs a zero row
    // between stacked images to move the split (pad >= 1) -- and stride 2 needs an even pad for the plane origin
    const bool even_only = rowbytes % 128 != 0;
    if (even_only && ((2 * rowbytes) % 128 != 0 || pad < 1)) return;
    if (stride == 2 && (pad & 1)) return;
    if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0 || Cin % 16 != 0 || !(RB == 32 || RB == 64 || RB == 128)) return;
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (enc == nullptr) return;
    static TmapEntry cache[64];
    static int next = 0;
    static std::mutex mu;
    const TmapKey key{x, nimg, Hin, Win, Cin, RB, stride, ks, fixed_rows};
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < 64; ++i)
        if (cache[i].valid && cache[i].key == key) {
            memcpy(p.tmap, cache[i].maps, sizeof(p.tmap));
            p.tma = 1;
            return;
        }
    TmapEntry& e = cache[next];
    e.valid = false;
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)nimg};
    const cuuint64_t strides[3] = {(cuuint64_t)Cin, (cuuint64_t)Win * Cin, (cuuint64_t)Hin * Win * Cin};
    const cuuint32_t es[4] = {1u, (cuuint32_t)stride, 1u, 1u};
    const CUtensorMapSwizzle sw = RB == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (RB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    for (int lg = 0; lg < TMA_NMAPS; ++lg) {
        // box: RB channel bytes x one patch row (stride 2: every other pixel of 2 * rowpix) x 2^lg rows x one image
        const cuuint32_t box[4] = {(cuuint32_t)RB, (cuuint32_t)(stride * rowpix), fixed_rows > 0 ? (cuuint32_t)fixed_rows : 1u << lg, 1u};
        if (enc(&e.maps[lg], CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return;
    }
    e.key = key;
    e.valid = true;
    next = (next + 1) % 64;
    memcpy(p.tmap, e.maps, sizeof(p.tmap));
    p.tma = 1;
}

int rowbytes_for(int Cin, int ks) { return (ks <= 3 && Cin % 64 == 0) ? 64 : 32; }
// ------------------------------------------------------------------------------------------------ folded weight sets
// Digit-plane images of a folded NNConvUpsampling block straight from its fp32 5x5 weight (one launch; the same derivation as
// stereospike_b200.ops.fold_weight_sets, which needed ~250 small torch launches per block and therefore kept the fold out of the
// training step, where the weights change every iteration).
//   patterns of the 5 taps of one axis (source offset of tap k): L 0,1,1,2,2   M 0,0,1,1,2   A 0,0,0,1,1   B 0,1,1,1,2   C 0,0,1,1,1
//   dense[cy][cx] (cy, cx in {L, M}): 3x3, f[dy][dx] = sum of the taps (ky, kx) with pat_cy[ky] == dy and pat_cx[kx] == dx
//   rows[c] (c in {A, B, C}): 3x5, f[dy][kx] = sum over ky with pat_c[ky] == dy;   cols[c]: 3x5 in the transposed frame, f[dx][ky]
// Quantise with 1 bit of head-room for the tap sums, add a bit while any folded sum overflows the top balanced digit (<= 3 more bits
// can ever be needed: a sum has at most 4 taps), then write the three images.
// (a constexpr function, not a __constant__ table: with the loops unrolled every accumulator index below is a compile-time constant
//  and the accumulators stay in registers; indexed through constant memory they lived in local memory and the kernel took 0.2 ms)
__host__ __device__ constexpr int fold_pat(int p, int k) {
    constexpr int T[5][5] = {{0, 1, 1, 2, 2}, {0, 0, 1, 1, 2}, {0, 0, 0, 1, 1}, {0, 1, 1, 1, 2}, {0, 0, 1, 1, 1}};
    return T[p][k];
}

// folded sums of one (output channel, input channel) filter q[5][5]; calls f(set kind, set index, tap, value)
template <typename F>
__device__ __forceinline__ void fold_sums(const int (&q)[25], F&& f) {
#pragma unroll
    for (int cy = 0; cy < 2; ++cy)
#pragma unroll
        for (int cx = 0; cx < 2; ++cx) {
            int acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) acc[fold_pat(cy, ky) * 3 + fold_pat(cx, kx)] += q[ky * 5 + kx];
#pragma unroll
            for (int t = 0; t < 9; ++t) f(0, cy * 2 + cx, t, acc[t]);
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int ar[15], ac[15];
#pragma unroll
        for (int t = 0; t < 15; ++t) ar[t] = ac[t] = 0;
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
                ar[fold_pat(2 + c, ky) * 5 + kx] += q[ky * 5 + kx];      // rows: [dy][kx]
                ac[fold_pat(2 + c, kx) * 5 + ky] += q[ky * 5 + kx];      // cols: [dx][ky]
            }
#pragma unroll
        for (int t = 0; t < 15; ++t) {
            f(1, c, t, ar[t]);
            f(2, c, t, ac[t]);
        }
    }
}

// Kernel A, one block per output channel: the exponent (1 bit of head-room, one more while any folded sum overflows).
__global__ void __launch_bounds__(256) weight_fold_exponent_kernel(const float* __restrict__ w, int Cin, int planes, float* __restrict__ wscale,
                                                                   int* __restrict__ wexp) {
    const int n = blockIdx.x;
    const float* wn = w + (size_t)n * Cin * 25;
    __shared__ float redf[8];
    __shared__ int redi[8];
    __shared__ int s_e;
    float m = 0.0f;
    for (int i = threadIdx.x; i < Cin * 25; i += blockDim.x) m = fmaxf(m, fabsf(wn[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) m = fmaxf(m, redf[i]);
        int ex = 0;
        if (m > 0.0f) frexpf(m, &ex);                 // m = f * 2^ex, f in [0.5, 1)  ==  floor(log2 m) + 1
        s_e = ex - (8 * planes - 1) + 1;
    }
    __syncthreads();
    int ilimit = 127;
    for (int pl = 1; pl < planes; ++pl) ilimit *= 256;      // planes <= 3 (the host refuses 4: 127 * 2^24 does not fit)
    // w * 2^-e is exact in fp32 (a power-of-two scaling of an fp32 number, |result| < 2^24), so rounding it to the nearest-even
    // integer in fp32 gives the same q as the float64 host derivation
    for (int it = 0; it < 4; ++it) {
        const int e = s_e;
        const float sc = ldexpf(1.0f, -e);
        int worst = 0;
        for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) {
            int q[25];
#pragma unroll
            for (int k = 0; k < 25; ++k) q[k] = __float2int_rn(wn[ci * 25 + k] * sc);
            fold_sums(q, [&](int, int, int, int v) { worst = max(worst, abs(v)); });
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) worst = max(worst, __shfl_xor_sync(0xffffffffu, worst, o));
        __syncthreads();                              // everybody has read s_e
        if ((threadIdx.x & 31) == 0) redi[threadIdx.x >> 5] = worst;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 0; i < 8; ++i) worst = max(worst, redi[i]);
            if (worst > ilimit) s_e = e + 1;
        }
        __syncthreads();
        if (s_e == e) break;
    }
    if (threadIdx.x == 0) {
        wexp[n] = s_e;
        wscale[n] = ldexpf(1.0f, s_e);
    }
}

// Kernel B, one block per (output channel, 64 input channels): the quantised taps go through shared memory, then every thread
// owns (folded tap, 16 consecutive input channels) items: 16 sums, their digits, one 16-byte store per plane -- 16 consecutive
// channels of one (tap, plane, output channel) row are one 16-byte chunk of the image.  (One thread per input channel writing its
// 378 bytes one by one took 0.09-0.16 ms per block: ~100 dependent instructions per byte-sized store.)
constexpr int FOLD_ITEMS = 36 + 45 + 45;      // dense 4 x 9, rows 3 x 15, cols 3 x 15
__constant__ uint32_t c_fold_mask[FOLD_ITEMS];  // bit (ky * 5 + kx) set: the tap belongs to the item's sum

__global__ void __launch_bounds__(256) weight_fold_write_kernel(const float* __restrict__ w, int Cin, int planes, int RBd,
                                                                const int* __restrict__ wexp, int8_t* __restrict__ w_dense,
                                                                int8_t* __restrict__ w_rows, int8_t* __restrict__ w_cols) {
    const int n = blockIdx.x;
    const int ci0 = blockIdx.y * 64;
    __shared__ int q[25][64 + 1];
    const float sc = ldexpf(1.0f, -wexp[n]);
    const float* wn = w + ((size_t)n * Cin + ci0) * 25;
    for (int i = threadIdx.x; i < 64 * 25; i += blockDim.x) {
        const int ci = i / 25, k = i - ci * 25;
        q[k][ci] = (ci0 + ci < Cin) ? __float2int_rn(wn[i] * sc) : 0;
    }
    __syncthreads();
    const int tile = n >> 5, r32 = n & 31;
    const int N = planes * 32;
    for (int idx = threadIdx.x; idx < FOLD_ITEMS * 4; idx += blockDim.x) {
        const int item = idx >> 2, cg = idx & 3;          // 16-channel group of the 64
        const int ch0 = ci0 + cg * 16;
        if (ch0 >= Cin) continue;
        int v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0;
        uint32_t mask = c_fold_mask[item];
        while (mask != 0u) {
            const int k = __ffs(mask) - 1;
            mask &= mask - 1u;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += q[k][cg * 16 + j];
        }
        // destination image and coordinates of the item
        int8_t* out;
        int set, tap, ntaps, nsets, RB;
        if (item < 36) { out = w_dense; set = item / 9; tap = item - set * 9; ntaps = 9; nsets = 4; RB = RBd; }
        else if (item < 81) { out = w_rows; set = (item - 36) / 15; tap = (item - 36) - set * 15; ntaps = 15; nsets = 3; RB = 32; }
        else { out = w_cols; set = (item - 81) / 15; tap = (item - 81) - set * 15; ntaps = 15; nsets = 3; RB = 32; }
        const int ntile = tile * nsets + set;             // packed "output channel" = (tile * nsets + set) * 32 + r32
        const int ncb = Cin / RB;
        const int cb = ch0 / RB, c = ch0 - cb * RB;
        const size_t buf = (size_t)(ntile * ncb + cb) * ((size_t)ntaps * N * RB);
        const uint32_t smask = (uint32_t)(RB >> 4) - 1u;
#pragma unroll
        for (int pl = planes - 1; pl >= 0; --pl) {
            uint32_t wd[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                int dgt;
                if (pl > 0) {
                    dgt = ((v[j] + 128) & 255) - 128;
                    v[j] = (v[j] - dgt) >> 8;
                } else {
                    dgt = v[j];
                }
                wd[j >> 2] |= (uint32_t)(dgt & 255) << (8 * (j & 3));
            }
            const uint32_t off = (uint32_t)((tap * N + pl * 32 + r32) * RB + c);
            *reinterpret_cast<uint4*>(out + buf + (off ^ (((off >> 7) & smask) << 4))) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        }
    }
}

// gradient-side correlation: row bytes of the bf16 source patch (Cg channels = 2*Cg bytes per pixel)
int corr_rowbytes_for(int Cg, int ks) { return (ks <= 3 && (2 * Cg) % 64 == 0) ? 64 : 32; }

}  // namespace
}  // namespace ss

using namespace ss;

extern "C" int ss_conv_i8_rowbytes(int32_t Cin, int32_t ks) { return rowbytes_for(Cin, ks); }

extern "C" int ss_pack_weights_i8(const float* w_oihw, int32_t Cout, int32_t Cin, int32_t ks, int32_t planes, void* w_i8,
                                  float* wscale, int32_t* wexp, void* stream) {
    const bool first = Cin >= 1 && Cin <= 4;
    if (w_oihw == nullptr || w_i8 == nullptr || wscale == nullptr || wexp == nullptr || Cout <= 0 || Cin <= 0 || ks <= 0 ||
        planes < 2 || planes > 4 || Cout % 32 != 0 || (!first && Cin % 32 != 0) || (first && ks != 5)) {
        set_error("ss_pack_weights_i8: bad argument (Cout %% 32; Cin %% 32, or Cin <= 4 with ks 5; planes 2..4)");
        return SS_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int RB = rowbytes_for(Cin, ks);
    weight_exponent_kernel<<<Cout, 256, 0, st>>>(w_oihw, Cout, Cin * ks * ks, planes, wscale, wexp);
    count_launch();
    if (check_launch("weight_exponent") != SS_OK) return SS_ECUDA;
    if (first) {
        weight_pack_first_kernel<<<(Cout * 128 + 255) / 256, 256, 0, st>>>(w_oihw, Cout, Cin, planes, wexp,
                                                                            reinterpret_cast<int8_t*>(w_i8));
        count_launch();
        return check_launch("weight_pack_first");
    }
    const long long total = (long long)Cout * Cin * ks * ks;
    weight_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w_oihw, Cout, Cin, ks, ks, planes, RB, wexp,
                                                                       reinterpret_cast<int8_t*>(w_i8));
    count_launch();
    return check_launch("weight_pack");
}

extern "C" int ss_pack_events_c(const float* x_btchw, int32_t B, int32_t T, int32_t C, int32_t Cpad, int32_t H, int32_t W, void* out,
                                int32_t* status, void* stream) {
    if (x_btchw == nullptr || out == nullptr || C <= 0 || C > Cpad || !(Cpad == 4 || (Cpad % 32 == 0 && Cpad <= 256)) || B < 0 || T < 0 ||
        H <= 0 || W <= 0) {
        set_error("ss_pack_events: bad argument (1 <= C <= Cpad, Cpad = 4 or a multiple of 32)");
        return SS_EINVAL;
    }
    const long long HW = (long long)H * W;
    if ((long long)T * B * HW == 0) return SS_OK;
    const bool vec = HW % 4 == 0 && (reinterpret_cast<uintptr_t>(x_btchw) & 15u) == 0;
    const long long n = (long long)T * B * (HW / (vec ? 4 : 1)) * (Cpad / 4);
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (vec) pack_events_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(x_btchw, B, T, C, Cpad, H, W, reinterpret_cast<uint32_t*>(out), status);
    else pack_events_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x_btchw, B, T, C, Cpad, H, W, reinterpret_cast<uint32_t*>(out), status);
    count_launch();
    return check_launch("pack_events");
}

extern "C" int ss_pack_events(const float* x_btchw, int32_t B, int32_t T, int32_t C, int32_t H, int32_t W, void* out_tbhw4,
                              int32_t* status, void* stream) {
    if (C > 4) {
        set_error("ss_pack_events: bad argument (1 <= C <= 4); use ss_pack_events_c for more channels");
        return SS_EINVAL;
    }
    return ss_pack_events_c(x_btchw, B, T, C, 4, H, W, out_tbhw4, status, stream);
}

static int conv_i8_launch(const ss_block_desc* g, const ss_tile_maps* tm, const void* x, const void* w_i8, const float* wscale,
                          const float* decay, const float* v_in, float* v_out, const void* resid, void* out, float* h_seq,
                          void* tsum, void* stream) {
    if (g == nullptr) {
        set_error("ss_conv_i8_fwd: null descriptor");
        return SS_EINVAL;
    }
    if (g->T == 0 || g->B == 0) return SS_OK;
    if (x == nullptr || w_i8 == nullptr || wscale == nullptr || out == nullptr) {
        set_error("ss_conv_i8_fwd: null argument");
        return SS_EINVAL;
    }
    if (g->T < 0 || g->B < 0 || g->Hin <= 0 || g->Win <= 0 || g->Hout <= 0 || g->Wout <= 0 || g->Cin <= 0 || g->Cout <= 0) {
        set_error("ss_conv_i8_fwd: bad geometry");
        return SS_EINVAL;
    }
    const bool first = g->Cin == 4;
    if ((!first && g->Cin % 32 != 0) || g->Cout % 32 != 0) {
        set_error("ss_conv_i8_fwd: Cin and Cout must be multiples of 32 (got %d, %d); Cin == 4 selects the first-layer mode", g->Cin,
                  g->Cout);
        return SS_EUNSUPPORTED;
    }
    if (first && (g->ks != 5 || g->stride != 1 || g->upsample != 0)) {
        set_error("ss_conv_i8_fwd: the first-layer mode (Cin == 4) is a 5x5 stride-1 conv");
        return SS_EUNSUPPORTED;
    }
    if (g->planes < 2 || g->planes > 4) {
        set_error("ss_conv_i8_fwd: planes must be 2, 3 or 4");
        return SS_EINVAL;
    }
    if (g->neuron < SS_NEURON_IF || g->neuron > SS_NEURON_PLIF) {
        set_error("ss_conv_i8_fwd: unknown neuron kind %d", g->neuron);
        return SS_EINVAL;
    }
    if (g->neuron == SS_NEURON_PLIF && decay == nullptr) {
        set_error("ss_conv_i8_fwd: PLIF needs the decay scalar");
        return SS_EINVAL;
    }
    if (g->neuron == SS_NEURON_LIF && !(g->tau > 1.0f)) {
        set_error("ss_conv_i8_fwd: LIF needs tau > 1");
        return SS_EINVAL;
    }
    if (tsum != nullptr && 3 * (g->T - 1) > 255) {
        set_error("ss_conv_i8_fwd: tsum needs 3*(T-1) <= 255 (got T = %d)", g->T);
        return SS_EUNSUPPORTED;
    }
    const bool up = g->upsample != 0;
    const int mode = tm != nullptr ? tm->mode : SS_TILES_PLAIN;
    if (mode < SS_TILES_PLAIN || mode > SS_TILES_ROW_LIST) {
        set_error("ss_conv_i8_fwd_ex: unknown tile mode %d", mode);
        return SS_EINVAL;
    }
    const bool rowlist = mode == SS_TILES_ROW_LIST;
    const bool indep = tm != nullptr && tm->independent_steps != 0;
    if (indep && (h_seq != nullptr || v_in != nullptr || v_out != nullptr || tsum != nullptr || tm->stats != nullptr || g->planes != 3 ||
                  g->v_reset != 0.0f)) {
        set_error("ss_conv_i8_fwd_ex: independent steps are for stateless inference (no h_seq / v_in / v_out / tsum / stats, 3 planes, v_reset 0)");
        return SS_EINVAL;
    }
    if (mode == SS_TILES_FOLDED && (tm->ymap_out == nullptr || tm->xmap_out == nullptr || g->ks != 3 || g->stride != 1 || g->pad != 0 ||
                                    up || g->Hin < 3 || g->Win < 3)) {
        set_error("ss_conv_i8_fwd_ex: the folded pass is a 3x3 stride-1 pad-0 conv on the source with output maps");
        return SS_EINVAL;
    }
    if (rowlist && tm->transposed == 2 && (tm->xmap_out == nullptr || tm->nclass != 4 || g->Cin % 64 != 0 || g->Win < 3)) {
        set_error("ss_conv_i8_fwd_ex: the row-list pass with folded columns needs xmap_out, 4 classes and Cin %% 64 == 0");
        return SS_EINVAL;
    }
    if (rowlist && (!up || g->ks != 5 || first || tm->rl_src == nullptr || tm->rl_out == nullptr || tm->rl_n <= 0 || tm->nclass <= 0 ||
                    tm->nclass > 8)) {
        set_error("ss_conv_i8_fwd_ex: the row-list pass needs a 5x5 upsampled conv, the entry tables and 1..8 classes");
        return SS_EINVAL;
    }
    if (!(g->ks == 3 || g->ks == 5) || !(g->stride == 1 || g->stride == 2) || (up && g->stride != 1) ||
        (g->stride == 2 && (g->pad % 2 != 0 || g->ks != 5))) {
        set_error("ss_conv_i8_fwd: unsupported conv shape (ks %d stride %d pad %d upsample %d)", g->ks, g->stride, g->pad, g->upsample);
        return SS_EUNSUPPORTED;
    }
    I8Params p{};
    p.T = g->T; p.B = g->B; p.Hin = g->Hin; p.Win = g->Win; p.Cin = g->Cin;
    p.Hout = g->Hout; p.Wout = g->Wout; p.Cout = g->Cout;
    p.ks = g->ks; p.stride = g->stride; p.pad = up ? 0 : g->pad; p.upsample = up ? 1 : 0;
    if (!up && mode != SS_TILES_FOLDED) {
        const int ho = (g->Hin + 2 * g->pad - g->ks) / g->stride + 1, wo = (g->Win + 2 * g->pad - g->ks) / g->stride + 1;
        if (ho != g->Hout || wo != g->Wout) {
            set_error("ss_conv_i8_fwd: Hout/Wout (%d,%d) do not match the conv geometry (%d,%d)", g->Hout, g->Wout, ho, wo);
            return SS_EINVAL;
        }
    }
    p.N = g->planes * 32;
    if (first) {
        p.RB = 128; p.ncb = 1; p.ntaps = 1; p.PH = 16; p.PWhalf = 8; p.PWp = 8;
    } else if (rowlist && tm->transposed == 2) {
        // rows AND columns folded (the dense 3x3 sets), tile rows from the list of regular rows of each row class
        p.RB = rowbytes_for(g->Cin, 3);
        p.ncb = g->Cin / p.RB;
        p.ntaps = 9;
        p.PH = 48; p.PWhalf = 9; p.PWp = 10;
    } else if (rowlist) {
        // rows folded to 3 taps (class-specific sums of the 5 filter rows), columns still the 5 taps over the upsampled row
        p.RB = 32;
        p.ncb = g->Cin / 32;
        p.ntaps = 15;
        p.PH = 48; p.PWhalf = 10; p.PWp = 12;
    } else {
        p.RB = rowbytes_for(g->Cin, g->ks);
        p.ncb = g->Cin / p.RB;
        p.ntaps = g->ks * g->ks;
        p.PH = 15 * g->stride + g->ks;
        p.PWhalf = 8 + (g->ks - 1) / 2;
        p.PWp = (g->stride == 1) ? 8 + g->ks - 1 : 2 * p.PWhalf;
    }
    p.ppix = p.PH * p.PWp;
    p.Hup = g->Hout + g->ks - 1;
    p.Wup = g->Wout + g->ks - 1;
    p.mode = mode;
    p.nclass = mode == SS_TILES_FOLDED ? 4 : (rowlist ? tm->nclass : 1);
    p.Hv = g->Hout; p.Wv = g->Wout;
    p.ymap_out = p.xmap_out = p.rl_src = p.rl_out = nullptr;
    p.rl_collive = nullptr;
    if (mode == SS_TILES_FOLDED) {
        p.Hv = g->Hin - 2; p.Wv = g->Win - 2;
        p.ymap_out = tm->ymap_out; p.xmap_out = tm->xmap_out;
    } else if (rowlist) {
        const bool tr = tm->transposed == 1;      // the list holds output COLUMNS (and the tile columns walk the rows)
        p.rl_fold = tm->transposed == 2 ? 1 : 0;
        p.rl_src = tm->rl_src; p.rl_out = tm->rl_out; p.rl_collive = tm->rl_collive; p.rl_n = tm->rl_n;
        p.in_rowstep = tr ? 1 : g->Win;
        p.in_colpitch = tr ? g->Win : 1;
        p.out_colpitch = tr ? g->Wout : 1;
        p.c_in = tr ? g->Hin : g->Win;
        p.c_nout = tr ? g->Hout : g->Wout;
        p.c_up = p.c_nout + g->ks - 1;
        p.c_scale = (float)p.c_in / (float)p.c_up;
        p.Wv = p.c_nout;
        if (p.rl_fold) {
            p.Wv = g->Win - 2;                // tile columns walk the source positions of the 3x3 folded conv
            p.xmap_out = tm->xmap_out;
            p.nclass = 4;
        }
    }
    if (mode == SS_TILES_FOLDED) {
        p.HsO = g->Hin;                       // virtual 3x3 pad-0 conv: image b's rows are Hin apart, no shared padding
    } else if (rowlist) {
        p.HsO = 16;                           // unused: the tile rows come from the list
    } else if (up) {
        p.HsO = p.Hup;
    } else {
        // Stacked output rows per image.  Rows of the next image start stride*HsO input rows later; that must be
        // past this image's real rows (Hin + pad) and far enough that a tap reaching below the last output row
        // lands in the next image's top padding (local index < pad), which is zero-filled like ours.
        const int need = g->Hin + p.pad;
        const int reach = g->stride * (g->Hout - 1) + g->ks - p.pad;
        const int span = need > reach ? need : reach;
        p.HsO = (span + g->stride - 1) / g->stride;
        if (p.HsO < g->Hout) p.HsO = g->Hout;
    }
    const long long rows = rowlist ? (long long)tm->rl_n : (long long)p.HsO * g->B;
    const int tiles_y = (int)((rows + 15) / 16);
    p.tiles_x = (p.Wv + 7) / 8;
    p.mtiles = tiles_y * p.tiles_x;
    const int ntiles = g->Cout / 32;
    const long long nitems = (long long)p.mtiles * ntiles * p.nclass;
    if (nitems > 0x7fffffffLL || (long long)g->B * g->Hin * g->Win * g->Cin > 0x7fffffffLL) {
        set_error("ss_conv_i8_fwd: problem too large for 32-bit indexing");
        return SS_EINVAL;
    }
    p.nitems = (int)nitems;
    p.defer_wait = (tm != nullptr && tm->defer_wait != 0) ? 1 : 0;
    p.item_tab = nullptr;
    if (rowlist && tm->item_tab != nullptr) {
        // per-class tile counts: the caller enumerates the items (weight set, m-tile) itself; lists concatenated, rl_n = total
        if (tm->n_items <= 0) {
            set_error("ss_conv_i8_fwd_ex: item table without items");
            return SS_EINVAL;
        }
        p.item_tab = reinterpret_cast<const int2*>(tm->item_tab);
        p.nitems = tm->n_items;
    }
    p.m_mtiles = div_magic(p.mtiles, nitems);
    p.m_tiles_x = div_magic(p.tiles_x, p.mtiles);
    p.m_per = div_magic((long long)p.stride * p.HsO, (long long)p.stride * (rows + 64));
    p.m_hso = div_magic(p.HsO, rows + 64);
    // CTA pairs (cta_group::2): two m-tiles per MMA, each CTA holds half of the weight rows.  Measured (r1h): the MMA-bound
    // deep decoder blocks gain ~5 %, everything else loses to the coarser item granularity (a pair walks items in lock
    // step), so pairs are used where the weights are streamed in >= 8 channel blocks and the number of rounds over the
    // machine does not grow.  SS_PAIR=0 never, SS_PAIR=2 wherever possible (tests), default = the rule above.
    static int pair_env = -1;
    if (pair_env < 0) {
        const char* e = getenv("SS_PAIR");
        pair_env = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
    }
    p.mtiles2 = (p.mtiles + 1) / 2;
    p.nitems2 = p.mtiles2 * ntiles * p.nclass;
    p.m_mtiles2 = div_magic(p.mtiles2, p.nitems2);
    const int dev = current_device();
    const int sms = device_sm_count(dev);
    bool pair = pair_env != 0 && !indep && !first && mode == SS_TILES_PLAIN && (g->planes * 32) % 16 == 0 && p.mtiles >= 2 && sms >= 2;
    if (pair && pair_env == 1) {
        const long long rounds1 = (nitems + sms - 1) / sms;
        const long long rounds2 = ((long long)p.nitems2 + sms / 2 - 1) / (sms / 2);
        pair = g->Cin / rowbytes_for(g->Cin, g->ks) >= 8 && rounds2 <= rounds1;
    }
    p.TC = 512 / p.N;
    if (p.TC > MAX_SLOTS) p.TC = MAX_SLOTS;
    p.WB = p.ntaps * (pair ? p.N / 2 : p.N) * p.RB;       // per CTA
    // stride 2: two parity planes, each padded to a multiple of 4 pixels (see plane_pixels)
    p.PB = ((g->stride == 2 && !first && !rowlist ? 2 * plane_pixels(p.PH, p.PWhalf) : p.ppix) * p.RB + 1023) / 1024 * 1024;
    p.resident = p.ncb <= NWB ? 1 : 0;
    p.nwb = p.ncb < NWB ? p.ncb : NWB;
    if (p.PH > 48 || p.PWp > 24) {
        set_error("ss_conv_i8_fwd: patch too large");
        return SS_EUNSUPPORTED;
    }
    const int tail_bytes = 1152 + 2048 + 64;   // tables + barriers, first-layer raw patch double buffer
    const int budget = 227 * 1024 - 1024 - tail_bytes - p.nwb * p.WB;
    int nps = budget / p.PB;
    if (nps > MAX_STAGES) nps = MAX_STAGES;
    if (nps < 2) {
        set_error("ss_conv_i8_fwd: not enough shared memory (weights %d B x %d, patch %d B)", p.WB, p.nwb, p.PB);
        return SS_EUNSUPPORTED;
    }
    p.NPS = nps;
    p.yscale = up ? (float)g->Hin / (float)p.Hup : 1.0f;
    p.xscale = up ? (float)g->Win / (float)p.Wup : 1.0f;
    p.neuron = g->neuron; p.gain = g->gain; p.v_th = g->v_th; p.v_reset = g->v_reset; p.tau = g->tau;
    p.x = reinterpret_cast<const uint8_t*>(x);
    p.w = reinterpret_cast<const int8_t*>(w_i8);
    p.wscale = wscale; p.decay = decay; p.v_in = v_in; p.v_out = v_out;
    p.resid = reinterpret_cast<const uint8_t*>(resid);
    p.out = reinterpret_cast<uint8_t*>(out);
    p.h_seq = h_seq;
    p.tsum = reinterpret_cast<uint8_t*>(tsum);
    p.stats = tm != nullptr ? reinterpret_cast<unsigned long long*>(tm->stats) : nullptr;
    p.tma = 0;
    if (!first && !up && !rowlist) setup_tma(p, x, (long long)g->T * g->B, g->Hin, g->Win, g->Cin, p.RB, g->stride, g->ks, p.pad);
    // row list with folded columns: 3 source rows x 10 source columns per tile row, straight from the source tensor
    if (rowlist && p.rl_fold && g->Hin < 65536) setup_tma(p, x, (long long)g->T * g->B, g->Hin, g->Win, g->Cin, p.RB, 1, 3, 0, 3);

    const size_t smem = 1024 + (size_t)p.nwb * p.WB + (size_t)p.NPS * p.PB + tail_bytes;
    const int num_sms = sms;
    int grid = p.nitems < num_sms ? p.nitems : num_sms;
    if (pair) {
        grid = 2 * p.nitems2 < num_sms ? 2 * p.nitems2 : (num_sms & ~1);
    }
    cudaStream_t st = (cudaStream_t)stream;
    bool launched = false;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: our prologue may overlap the previous kernel's tail
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    attrs[1].id = cudaLaunchAttributeClusterDimension;                  // CTA pairs
    attrs[1].val.clusterDim.x = 2;
    attrs[1].val.clusterDim.y = 1;
    attrs[1].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = pair ? 2 : 1;
#define SS_TRY_PAIR(PL, KS_, ST_, RB_)                                                                                     \
    if (!launched && pair && g->planes == PL && g->ks == KS_ && g->stride == ST_ && p.RB == RB_) {                         \
        SS_ENSURE_SMEM((conv_i8_kernel<PL, KS_, ST_, RB_, false, MODE_I8, true>), dev, 227 * 1024);                        \
        cudaLaunchKernelEx(&cfg, conv_i8_kernel<PL, KS_, ST_, RB_, false, MODE_I8, true>, p);                              \
        launched = true;                                                                                                   \
    }
#define SS_TRY_NK(PL, KS_, ST_, RB_, NK_)                                                                                   \
    if (!launched && spec && g->neuron == NK_ && g->planes == PL && g->ks == KS_ && g->stride == ST_ && p.RB == RB_) {     \
        SS_ENSURE_SMEM((conv_i8_kernel<PL, KS_, ST_, RB_, false, MODE_I8, false, KS_, 1, NK_>), dev, 227 * 1024);          \
        cudaLaunchKernelEx(&cfg, conv_i8_kernel<PL, KS_, ST_, RB_, false, MODE_I8, false, KS_, 1, NK_>, p);                \
        launched = true;                                                                                                   \
    }
#define SS_TRY(PL, KS_, ST_, RB_)                                                                                          \
    SS_TRY_PAIR(PL, KS_, ST_, RB_)                                                                                         \
    if (!launched && g->planes == PL && g->ks == KS_ && g->stride == ST_ && p.RB == RB_) {                                 \
        SS_ENSURE_SMEM((conv_i8_kernel<PL, KS_, ST_, RB_>), dev, 227 * 1024);                                              \
        cudaLaunchKernelEx(&cfg, conv_i8_kernel<PL, KS_, ST_, RB_>, p);                                                    \
        launched = true;                                                                                                   \
    }
#define SS_TRY_FIRST(PL)                                                                                                   \
    if (!launched && first && g->planes == PL) {                                                                           \
        SS_ENSURE_SMEM((conv_i8_kernel<PL, 1, 1, 128, true>), dev, 227 * 1024);                                            \
        cudaLaunchKernelEx(&cfg, conv_i8_kernel<PL, 1, 1, 128, true>, p);                                                  \
        launched = true;                                                                                                   \
    }
#define SS_TRY_ROWLIST(PL)                                                                                                 \
    if (!launched && rowlist && p.rl_fold && g->planes == PL) {                                                            \
        SS_ENSURE_SMEM((conv_i8_kernel<PL, 3, 1, 64, false, MODE_I8, false, 3, 3>), dev, 227 * 1024);                      \
        cudaLaunchKernelEx(&cfg, conv_i8_kernel<PL, 3, 1, 64, false, MODE_I8, false, 3, 3>, p);                            \
        launched = true;                                                                                                   \
    }                                                                                                                      \
    if (!launched && rowlist && g->planes == PL) {                                                                         \
        SS_ENSURE_SMEM((conv_i8_kernel<PL, 3, 1, 32, false, MODE_I8, false, 5, 3>), dev, 227 * 1024);                      \
        cudaLaunchKernelEx(&cfg, conv_i8_kernel<PL, 3, 1, 32, false, MODE_I8, false, 5, 3>, p);                            \
        launched = true;                                                                                                   \
    }
    // independent steps (LEAN = 2): one instance per shape and neuron kind
#define SS_TRY_INDEP(COND, KS_, ST_, RB_, FIRST_, KSX_, RS_, NK_)                                                            \
    if (!launched && indep && g->neuron == NK_ && (COND)) {                                                                 \
        SS_ENSURE_SMEM((conv_i8_kernel<3, KS_, ST_, RB_, FIRST_, MODE_I8, false, KSX_, RS_, NK_, 2>), dev, 227 * 1024);      \
        cudaLaunchKernelEx(&cfg, conv_i8_kernel<3, KS_, ST_, RB_, FIRST_, MODE_I8, false, KSX_, RS_, NK_, 2>, p);            \
        launched = true;                                                                                                    \
    }
#define SS_PLAIN_IS(KS_, ST_, RB_) (!first && !rowlist && g->ks == KS_ && g->stride == ST_ && p.RB == RB_)
#define SS_TRY_INDEP_NK(NK_)                                                                                                \
    SS_TRY_INDEP(first, 1, 1, 128, true, 1, 1, NK_)                                                                         \
    SS_TRY_INDEP(rowlist && p.rl_fold, 3, 1, 64, false, 3, 3, NK_)                                                          \
    SS_TRY_INDEP(rowlist && !p.rl_fold, 3, 1, 32, false, 5, 3, NK_)                                                         \
    SS_TRY_INDEP(SS_PLAIN_IS(5, 1, 32), 5, 1, 32, false, 5, 1, NK_)                                                         \
    SS_TRY_INDEP(SS_PLAIN_IS(5, 2, 32), 5, 2, 32, false, 5, 1, NK_)                                                         \
    SS_TRY_INDEP(SS_PLAIN_IS(3, 1, 64), 3, 1, 64, false, 3, 1, NK_)                                                         \
    SS_TRY_INDEP(SS_PLAIN_IS(3, 1, 32), 3, 1, 32, false, 3, 1, NK_)
    SS_TRY_INDEP_NK(SS_NEURON_IF)
    SS_TRY_INDEP_NK(SS_NEURON_LIF)
    SS_TRY_INDEP_NK(SS_NEURON_PLIF)
#undef SS_TRY_INDEP_NK
#undef SS_PLAIN_IS
#undef SS_TRY_INDEP
    if (indep && !launched) {
        set_error("ss_conv_i8_fwd_ex: no independent-steps instance for ks %d stride %d rowbytes %d", g->ks, g->stride, p.RB);
        return SS_EUNSUPPORTED;
    }
    // stateless inference (no h_seq, no membrane state in or out, no statistics): the LEAN instances of the epilogue-bound shapes
    static int lean_env = -1;
    if (lean_env < 0) {
        const char* e = getenv("SS_LEAN");
        lean_env = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    // (measured, alternating runs on one box: first layer -5 %, deconv1's dense pass -2.5 %; the wider blocks +1-2 %, so only the
    //  32-output-channel ones use it)
    const bool lean = lean_env != 0 && h_seq == nullptr && v_in == nullptr && v_out == nullptr && p.stats == nullptr && !pair && !rowlist &&
                      g->planes == 3 && g->v_reset == 0.0f && g->Cout <= 32;
#define SS_TRY_LEAN(KS_, ST_, RB_, FIRST_, NK_)                                                                             \
    if (!launched && lean && g->neuron == NK_ && first == FIRST_ && (FIRST_ || (g->ks == KS_ && g->stride == ST_ && p.RB == RB_))) { \
        SS_ENSURE_SMEM((conv_i8_kernel<3, KS_, ST_, RB_, FIRST_, MODE_I8, false, KS_, 1, NK_, true>), dev, 227 * 1024);    \
        cudaLaunchKernelEx(&cfg, conv_i8_kernel<3, KS_, ST_, RB_, FIRST_, MODE_I8, false, KS_, 1, NK_, true>, p);          \
        launched = true;                                                                                                    \
    }
#define SS_TRY_LEAN_NK(NK_) SS_TRY_LEAN(1, 1, 128, true, NK_) SS_TRY_LEAN(3, 1, 64, false, NK_)
    SS_TRY_LEAN_NK(SS_NEURON_IF)
    SS_TRY_LEAN_NK(SS_NEURON_LIF)
    SS_TRY_LEAN_NK(SS_NEURON_PLIF)
#undef SS_TRY_LEAN_NK
#undef SS_TRY_LEAN
#define SS_TRY_PL(PL) SS_TRY_ROWLIST(PL) SS_TRY_FIRST(PL) SS_TRY(PL, 5, 1, 32) SS_TRY(PL, 5, 2, 32) SS_TRY(PL, 3, 1, 64) SS_TRY(PL, 3, 1, 32)
    // default precision (3 planes), v_reset == 0 (every call site of the reference), no CTA pair: neuron kind compiled in
    const bool spec = !pair && !rowlist && !first && g->planes == 3 && g->v_reset == 0.0f;
#define SS_TRY_SPEC(NK_)                                                                                                   \
    SS_TRY_NK(3, 5, 1, 32, NK_) SS_TRY_NK(3, 5, 2, 32, NK_) SS_TRY_NK(3, 3, 1, 64, NK_) SS_TRY_NK(3, 3, 1, 32, NK_)        \
    if (!launched && first && g->planes == 3 && g->v_reset == 0.0f && g->neuron == NK_) {                                  \
        SS_ENSURE_SMEM((conv_i8_kernel<3, 1, 1, 128, true, MODE_I8, false, 1, 1, NK_>), dev, 227 * 1024);                  \
        cudaLaunchKernelEx(&cfg, conv_i8_kernel<3, 1, 1, 128, true, MODE_I8, false, 1, 1, NK_>, p);                        \
        launched = true;                                                                                                   \
    }                                                                                                                      \
    if (!launched && rowlist && p.rl_fold && g->planes == 3 && g->v_reset == 0.0f && g->neuron == NK_) {                   \
        SS_ENSURE_SMEM((conv_i8_kernel<3, 3, 1, 64, false, MODE_I8, false, 3, 3, NK_>), dev, 227 * 1024);                  \
        cudaLaunchKernelEx(&cfg, conv_i8_kernel<3, 3, 1, 64, false, MODE_I8, false, 3, 3, NK_>, p);                        \
        launched = true;                                                                                                   \
    }                                                                                                                      \
    if (!launched && rowlist && g->planes == 3 && g->v_reset == 0.0f && g->neuron == NK_) {                                \
        SS_ENSURE_SMEM((conv_i8_kernel<3, 3, 1, 32, false, MODE_I8, false, 5, 3, NK_>), dev, 227 * 1024);                  \
        cudaLaunchKernelEx(&cfg, conv_i8_kernel<3, 3, 1, 32, false, MODE_I8, false, 5, 3, NK_>, p);                        \
        launched = true;                                                                                                   \
    }
    SS_TRY_SPEC(SS_NEURON_IF)
    SS_TRY_SPEC(SS_NEURON_LIF)
    SS_TRY_SPEC(SS_NEURON_PLIF)
#undef SS_TRY_SPEC
    SS_TRY_PL(2)
    SS_TRY_PL(3)
    SS_TRY_PL(4)
#undef SS_TRY_PL
#undef SS_TRY_ROWLIST
#undef SS_TRY_FIRST
#undef SS_TRY
#undef SS_TRY_NK
#undef SS_TRY_PAIR
    if (!launched) {
        set_error("ss_conv_i8_fwd: no kernel instance for planes %d ks %d stride %d rowbytes %d", g->planes, g->ks, g->stride, p.RB);
        return SS_EUNSUPPORTED;
    }
    count_launch();
    return check_launch("conv_i8");
}

extern "C" int ss_conv_i8_fwd(const ss_block_desc* g, const void* x, const void* w_i8, const float* wscale, const float* decay,
                              const float* v_in, float* v_out, const void* resid, void* out, float* h_seq, void* tsum, void* stream) {
    return conv_i8_launch(g, nullptr, x, w_i8, wscale, decay, v_in, v_out, resid, out, h_seq, tsum, stream);
}

extern "C" int ss_conv_i8_fwd_ex(const ss_block_desc* g, const ss_tile_maps* tm, const void* x, const void* w_i8, const float* wscale,
                                 const float* decay, const float* v_in, float* v_out, const void* resid, void* out, float* h_seq,
                                 void* tsum, void* stream) {
    return conv_i8_launch(g, tm, x, w_i8, wscale, decay, v_in, v_out, resid, out, h_seq, tsum, stream);
}

extern "C" int ss_pack_digits_i8_rect(const float* q_oihw, int32_t Cout, int32_t Cin, int32_t ksy, int32_t ksx, int32_t planes,
                                      const int32_t* zero_exp, void* w_i8, void* stream) {
    if (q_oihw == nullptr || w_i8 == nullptr || zero_exp == nullptr || Cout <= 0 || Cin <= 0 || ksy <= 0 || ksx <= 0 || planes < 2 ||
        planes > 4 || Cout % 32 != 0 || Cin % 32 != 0) {
        set_error("ss_pack_digits_i8: bad argument (Cout %% 32, Cin %% 32, planes 2..4)");
        return SS_EINVAL;
    }
    // square filters keep the row bytes of ss_conv_i8_fwd; the rectangular 3x5 sets of the row-list pass use 32-byte rows
    const int RB = ksy == ksx ? rowbytes_for(Cin, ksy) : 32;
    const long long total = (long long)Cout * Cin * ksy * ksx;
    weight_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(q_oihw, Cout, Cin, ksy, ksx, planes, RB,
                                                                                           zero_exp, reinterpret_cast<int8_t*>(w_i8));
    count_launch();
    return check_launch("pack_digits");
}

extern "C" int ss_pack_digits_i8(const float* q_oihw, int32_t Cout, int32_t Cin, int32_t ks, int32_t planes, const int32_t* zero_exp,
                                 void* w_i8, void* stream) {
    return ss_pack_digits_i8_rect(q_oihw, Cout, Cin, ks, ks, planes, zero_exp, w_i8, stream);
}

extern "C" int ss_pack_weights_folded(const float* w_oihw, int32_t Cout, int32_t Cin, int32_t planes, void* w_dense, void* w_rows,
                                      void* w_cols, float* wscale, int32_t* wexp, void* stream) {
    if (w_oihw == nullptr || w_dense == nullptr || w_rows == nullptr || w_cols == nullptr || wscale == nullptr || Cout <= 0 || Cin <= 0 ||
        planes < 2 || planes > 3 || Cout % 32 != 0 || Cin % 32 != 0) {
        set_error("ss_pack_weights_folded: bad argument (5x5 weight, Cout %% 32, Cin %% 32, planes 2 or 3)");
        return SS_EINVAL;
    }
    // item -> taps table of the write kernel (per device, once)
    static bool table_done[SS_MAX_DEVICES] = {};
    const int dev = current_device();
    if (!table_done[dev]) {
        uint32_t mask[FOLD_ITEMS];
        for (int i = 0; i < FOLD_ITEMS; ++i) mask[i] = 0u;
        for (int ky = 0; ky < 5; ++ky)
            for (int kx = 0; kx < 5; ++kx) {
                const uint32_t bit = 1u << (ky * 5 + kx);
                for (int cy = 0; cy < 2; ++cy)
                    for (int cx = 0; cx < 2; ++cx) mask[(cy * 2 + cx) * 9 + fold_pat(cy, ky) * 3 + fold_pat(cx, kx)] |= bit;
                for (int c = 0; c < 3; ++c) {
                    mask[36 + c * 15 + fold_pat(2 + c, ky) * 5 + kx] |= bit;      // rows: [dy][kx]
                    mask[81 + c * 15 + fold_pat(2 + c, kx) * 5 + ky] |= bit;      // cols: [dx][ky]
                }
            }
        if (cudaMemcpyToSymbol(c_fold_mask, mask, sizeof(mask)) != cudaSuccess) {
            set_error("ss_pack_weights_folded: %s", cudaGetErrorString(cudaGetLastError()));
            return SS_ECUDA;
        }
        table_done[dev] = true;
    }
    if (wexp == nullptr) {
        set_error("ss_pack_weights_folded: null wexp workspace");
        return SS_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    weight_fold_exponent_kernel<<<Cout, 256, 0, st>>>(w_oihw, Cin, planes, wscale, wexp);
    count_launch();
    if (check_launch("weight_fold_exponent") != SS_OK) return SS_ECUDA;
    weight_fold_write_kernel<<<dim3((unsigned)Cout, (unsigned)((Cin + 63) / 64)), 256, 0, st>>>(
        w_oihw, Cin, planes, rowbytes_for(Cin, 3), wexp, reinterpret_cast<int8_t*>(w_dense), reinterpret_cast<int8_t*>(w_rows),
        reinterpret_cast<int8_t*>(w_cols));
    count_launch();
    return check_launch("weight_fold_pack");
}

// ------------------------------------------------------------------------------------------------ gradient-side correlation
extern "C" int ss_pack_weights_bf16(const float* w_oihw, int32_t Cout, int32_t Cin, int32_t ks, int32_t ntile, void* w_img,
                                    void* stream) {
    if (w_oihw == nullptr || w_img == nullptr || Cout <= 0 || Cin <= 0 || !(ks == 3 || ks == 5) || !(ntile == 32 || ntile == 64) ||
        Cout % ntile != 0 || Cin % 16 != 0) {
        set_error("ss_pack_weights_bf16: bad argument (ks 3 or 5, ntile 32 or 64, Cout %% ntile, Cin %% 16)");
        return SS_EINVAL;
    }
    const int RB = corr_rowbytes_for(Cin, ks);
    const long long total = (long long)Cout * Cin * ks * ks;
    weight_pack_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, ks, ntile, RB,
                                                                                                 reinterpret_cast<uint16_t*>(w_img));
    count_launch();
    return check_launch("weight_pack_bf16");
}

extern "C" int ss_corr_bf16(const ss_corr_desc* d, const void* src_bf16, const void* w_img, const int32_t* ymap_out,
                            const int32_t* xmap_out, float* dst, void* stream) {
    if (d == nullptr) {
        set_error("ss_corr_bf16: null descriptor");
        return SS_EINVAL;
    }
    if (d->T == 0 || d->B == 0) return SS_OK;
    if (src_bf16 == nullptr || w_img == nullptr || dst == nullptr) {
        set_error("ss_corr_bf16: null argument");
        return SS_EINVAL;
    }
    if (d->T < 0 || d->B < 0 || d->Hg <= 0 || d->Wg <= 0 || d->Cg <= 0 || d->Hv <= 0 || d->Wv <= 0 || d->Hdst <= 0 || d->Wdst <= 0 ||
        d->Cdst <= 0 || !(d->ks == 3 || d->ks == 5) || d->pad < 0 || d->pad >= d->ks || !(d->nclass == 1 || d->nclass == 4) ||
        !(d->ntile == 32 || d->ntile == 64) || d->Cdst % d->ntile != 0 || d->Cg % 16 != 0 || d->out_mode < SS_CORR_STORE ||
        d->out_mode > SS_CORR_ATOMIC) {
        set_error("ss_corr_bf16: bad descriptor");
        return SS_EINVAL;
    }
    const bool mapped = ymap_out != nullptr && xmap_out != nullptr;
    if (!mapped && (d->nclass != 1 || d->Hv != d->Hdst || d->Wv != d->Wdst)) {
        set_error("ss_corr_bf16: classes / a virtual grid different from the destination need the output maps");
        return SS_EINVAL;
    }
    I8Params p{};
    p.T = d->T; p.B = d->B; p.Hin = d->Hg; p.Win = d->Wg; p.Cin = 2 * d->Cg;   // the producers count bytes
    p.Hout = d->Hdst; p.Wout = d->Wdst; p.Cout = d->Cdst;
    p.ks = d->ks; p.stride = 1; p.pad = d->pad; p.upsample = 0;
    p.N = d->ntile;
    p.RB = corr_rowbytes_for(d->Cg, d->ks);
    p.ncb = p.Cin / p.RB;
    p.ntaps = d->ks * d->ks;
    p.PH = 15 + d->ks;
    p.PWhalf = 8 + (d->ks - 1) / 2;
    p.PWp = 8 + d->ks - 1;
    p.ppix = p.PH * p.PWp;
    p.Hup = p.Wup = 0;
    p.mode = mapped ? SS_TILES_FOLDED : SS_TILES_PLAIN;
    p.nclass = d->nclass;
    p.Hv = d->Hv; p.Wv = d->Wv;
    p.ymap_out = ymap_out; p.xmap_out = xmap_out;
    {
        // stacked virtual rows per image: past the source rows + padding, and far enough that taps reaching below the last
        // virtual row land in the next image's (zero) top padding
        const int need = d->Hg + d->pad;
        const int reach = (d->Hv - 1) + d->ks - d->pad;
        int span = need > reach ? need : reach;
        if (span < d->Hv) span = d->Hv;
        p.HsO = span;
    }
    const long long rows = (long long)p.HsO * d->B;
    const int tiles_y = (int)((rows + 15) / 16);
    p.tiles_x = (p.Wv + 7) / 8;
    p.mtiles = tiles_y * p.tiles_x;
    const long long nitems = (long long)p.mtiles * (d->Cdst / d->ntile) * p.nclass;
    if (nitems > 0x7fffffffLL || (long long)d->B * d->Hg * d->Wg * p.Cin > 0x7fffffffLL) {
        set_error("ss_corr_bf16: problem too large for 32-bit indexing");
        return SS_EINVAL;
    }
    p.nitems = (int)nitems;
    p.m_mtiles = div_magic(p.mtiles, nitems);
    p.m_tiles_x = div_magic(p.tiles_x, p.mtiles);
    p.m_per = div_magic((long long)p.stride * p.HsO, (long long)p.stride * (rows + 64));
    p.m_hso = div_magic(p.HsO, rows + 64);
    p.TC = 512 / p.N;
    if (p.TC > MAX_SLOTS) p.TC = MAX_SLOTS;
    p.WB = p.ntaps * p.N * p.RB;
    p.PB = (p.ppix * p.RB + 1023) / 1024 * 1024;
    p.resident = p.ncb <= NWB ? 1 : 0;
    p.nwb = p.ncb < NWB ? p.ncb : NWB;
    const int tail_bytes = 1152 + 2048 + 64;   // tables + barriers, first-layer raw patch double buffer
    const int budget = 227 * 1024 - 1024 - tail_bytes - p.nwb * p.WB;
    int nps = budget / p.PB;
    if (nps > MAX_STAGES) nps = MAX_STAGES;
    if (nps < 2) {
        set_error("ss_corr_bf16: not enough shared memory");
        return SS_EUNSUPPORTED;
    }
    p.NPS = nps;
    p.yscale = p.xscale = 1.0f;
    p.neuron = SS_NEURON_IF; p.gain = 1.0f; p.v_th = 1.0f; p.v_reset = 0.0f; p.tau = 2.0f;
    p.x = reinterpret_cast<const uint8_t*>(src_bf16);
    p.w = reinterpret_cast<const int8_t*>(w_img);
    p.g_dst = dst;
    p.g_mode = d->out_mode;
    setup_tma(p, src_bf16, (long long)d->T * d->B, d->Hg, d->Wg, p.Cin, p.RB, 1, d->ks, d->pad);

    const size_t smem = 1024 + (size_t)p.nwb * p.WB + (size_t)p.NPS * p.PB + tail_bytes;
    const int dev = current_device();
    const int num_sms = device_sm_count(dev);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(p.nitems < num_sms ? p.nitems : num_sms));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    bool launched = false;
#define SS_TRY_CORR(PL, KS_, RB_)                                                                                                  \
    if (!launched && d->ntile == PL * 32 && d->ks == KS_ && p.RB == RB_) {                                                         \
        SS_ENSURE_SMEM((conv_i8_kernel<PL, KS_, 1, RB_, false, MODE_BF16>), dev, 227 * 1024);                                      \
        cudaLaunchKernelEx(&cfg, conv_i8_kernel<PL, KS_, 1, RB_, false, MODE_BF16>, p);                                            \
        launched = true;                                                                                                           \
    }
    SS_TRY_CORR(2, 5, 32)
    SS_TRY_CORR(2, 3, 64)
    SS_TRY_CORR(2, 3, 32)
    SS_TRY_CORR(1, 5, 32)
    SS_TRY_CORR(1, 3, 64)
    SS_TRY_CORR(1, 3, 32)
#undef SS_TRY_CORR
    if (!launched) {
        set_error("ss_corr_bf16: no kernel instance for ntile %d ks %d rowbytes %d", d->ntile, d->ks, p.RB);
        return SS_EUNSUPPORTED;
    }
    count_launch();
    return check_launch("corr_bf16");
}

#ifdef SS_ROLE_TIMING
extern "C" int ss_debug_read(unsigned long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, ss::ss_dbg, sizeof(unsigned long long) * 148 * 4 * 8) == cudaSuccess ? 0 : -1;
}
#endif
