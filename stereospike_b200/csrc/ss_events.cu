// Event stream -> event-count frames on the device (the step in front of the hot path).
//
// Replaces the per-event Python loops of the reference's data layer:
//   mvsecRectifyEvents            datasets/MVSEC/utils.py:31-56    x,y -> rectification LUTs, drop events outside the field of view
//   mvsecCumulateSpikesIntoFrames datasets/MVSEC/utils.py:215-281  events with start < t < end (both strict) are counted per
//                                                                  polarity: channel 0 = ON (p == 1), channel 1 = everything else
// and produces the network's input directly in the layout the first block reads (u8 [T][B][H][W][4], left camera in
// channels 0-1, right camera in 2-3; train.py:201-218), so the dense fp64/fp32 frames never exist and never cross PCIe.
// One thread per event (binary search of its frame), 32-bit atomics into a [F][H][W][4] count image, then a saturating pack.
// HBM-bound: 32 B read per event + one count image.
#include "ss_common.cuh"

namespace ss {
namespace {

__global__ void __launch_bounds__(256) events_accumulate_kernel(const double* __restrict__ ev, long long n, const double* __restrict__ xmap,
                                                                const double* __restrict__ ymap, double t0,
                                                                const double* __restrict__ starts, const double* __restrict__ ends,
                                                                int F, int H, int W, int cam, unsigned int* __restrict__ counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = ev[4 * i], y = ev[4 * i + 1];
    const double t = ev[4 * i + 2] - t0;          // utils.py:251-253: timestamps are shifted by the first spike time
    const double pol = ev[4 * i + 3];
    if (xmap != nullptr) {
        // utils.py:43-48: the raw coordinates are truncated to integers and looked up in the rectification maps
        const int xi = (int)x, yi = (int)y;
        if (xi < 0 || xi >= W || yi < 0 || yi >= H) return;
        x = xmap[(size_t)yi * W + xi];
        y = ymap[(size_t)yi * W + xi];
        // utils.py:52-55 keeps 0 <= x <= 346, 0 <= y <= 260; x == 346 / y == 260 would then index out of the frame
        // (IndexError upstream) -- such events are dropped here
        if (!(x >= 0.0 && x <= (double)W && y >= 0.0 && y <= (double)H)) return;
    }
    const int px = (int)x, py = (int)y;           // utils.py:262-263: int() truncation
    if (px < 0 || px >= W || py < 0 || py >= H) return;
    // frame f with starts[f] < t < ends[f]  (utils.py:259: both comparisons strict; the boundaries are the caller's float64 values)
    int lo = 0, hi = F;                            // first f with ends[f] > t
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ends[mid] > t) hi = mid; else lo = mid + 1;
    }
    if (lo >= F || !(starts[lo] < t)) return;
    const int ch = cam * 2 + (pol == 1.0 ? 0 : 1);   // utils.py:266-269
    atomicAdd(counts + (((size_t)lo * H + py) * W + px) * 4 + ch, 1u);
}

// counts u32 [F = B*T][H][W][4] (frame f = b*T + t)  ->  u8 [T][B][H][W][4], saturating at 255
__global__ void __launch_bounds__(256) counts_pack_kernel(const uint4* __restrict__ counts, int B, int T, long long HW,
                                                          uint32_t* __restrict__ out, int* __restrict__ status) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * T * HW) return;
    const long long hw = i % HW;
    const int b = (int)((i / HW) % B);
    const int t = (int)(i / (HW * B));
    const uint4 c = counts[((size_t)b * T + t) * HW + hw];
    if ((c.x | c.y | c.z | c.w) > 255u && status != nullptr) atomicOr(status, 2);
    out[i] = min(c.x, 255u) | (min(c.y, 255u) << 8) | (min(c.z, 255u) << 16) | (min(c.w, 255u) << 24);
}

}  // namespace
}  // namespace ss

using namespace ss;

extern "C" int ss_events_accumulate(const double* events_xytp, int64_t n_events, const double* xmap, const double* ymap, double t0,
                                    const double* frame_start, const double* frame_end, int32_t n_frames, int32_t H, int32_t W,
                                    int32_t camera, uint32_t* counts_fhw4, void* stream) {
    if (n_events < 0 || n_frames <= 0 || H <= 0 || W <= 0 || camera < 0 || camera > 1 || frame_start == nullptr || frame_end == nullptr ||
        counts_fhw4 == nullptr || (n_events > 0 && events_xytp == nullptr) || ((xmap == nullptr) != (ymap == nullptr))) {
        set_error("ss_events_accumulate: bad argument");
        return SS_EINVAL;
    }
    if (n_events == 0) return SS_OK;
    events_accumulate_kernel<<<(unsigned)((n_events + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        events_xytp, n_events, xmap, ymap, t0, frame_start, frame_end, n_frames, H, W, camera, counts_fhw4);
    count_launch();
    return check_launch("events_accumulate");
}

extern "C" int ss_events_pack(const uint32_t* counts_fhw4, int32_t B, int32_t T, int32_t H, int32_t W, void* out_tbhw4, int32_t* status,
                              void* stream) {
    if (counts_fhw4 == nullptr || out_tbhw4 == nullptr || B < 0 || T < 0 || H <= 0 || W <= 0) {
        set_error("ss_events_pack: bad argument");
        return SS_EINVAL;
    }
    const long long n = (long long)B * T * H * W;
    if (n == 0) return SS_OK;
    counts_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(counts_fhw4), B, T,
                                                                                      (long long)H * W,
                                                                                      reinterpret_cast<uint32_t*>(out_tbhw4), status);
    count_launch();
    return check_launch("events_pack");
}
