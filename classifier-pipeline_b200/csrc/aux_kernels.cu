// aux_kernels.cu -- small per-frame kernels beside the persistent extraction kernel:
//   background_step_kernel   WeightedBackground.process_frame on an arbitrary frame  (K7, motiondetector.py:197-244)
//   frame_median_kernel      np.median of a uint16 frame                             (K8, clip.py:474-487; interpreter.py:389)
#include "background_step.cuh"
#include "median.cuh"

namespace cpt {

// ------------------------------------------------------------------------------------------------
// WeightedBackground.process_frame(frame) for one state record per CTA.  `frames` holds one int32
// frame (already truncated, np.int32(frame)) per record.  Same state layout as the extraction kernel,
// so a record can move between the two (streaming extractor sharing the motion detector's background).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) background_step_kernel(Geometry g, uint8_t *state, const int32_t *frames,
                                                                   const int *record_index, WeightTable wt) {
    __shared__ unsigned long long red_sum[32];
    __shared__ int red_changed;
    const int rec = record_index ? record_index[blockIdx.x] : (int)blockIdx.x;
    const int32_t *A = frames + (size_t)blockIdx.x * g.npx;
    background_step(g, state + (size_t)rec * state_bytes(g.npx), [A](int p) { return A[p]; }, wt, red_sum, &red_changed);
}

// ------------------------------------------------------------------------------------------------
// Median of each uint16 frame (mean of the two middle values for an even pixel count), one CTA per
// frame: the frame is staged in shared memory once, then median.cuh selects the ranks (n-1)/2 and n/2.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) frame_median_kernel(const uint16_t *frames, int npx, float *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t *px = reinterpret_cast<uint16_t *>(smem_raw);
    uint32_t *bins = reinterpret_cast<uint32_t *>(smem_raw + (((size_t)npx * sizeof(uint16_t) + 15) & ~(size_t)15));
    __shared__ int red[80];
    const uint16_t *src = frames + (size_t)blockIdx.x * npx;
    for (int i = threadIdx.x; i < npx / 8; i += blockDim.x) *reinterpret_cast<uint4 *>(px + i * 8) = ldg16(src + i * 8);
    for (int i = (npx / 8) * 8 + threadIdx.x; i < npx; i += blockDim.x) px[i] = src[i];
    __syncthreads();
    const int twice = rect_median_sum(px, npx, 0, 0, npx, 1, bins, red);  // the frame as one long row
    if (threadIdx.x == 0) out[blockIdx.x] = 0.5f * (float)twice;  // exact: at most 17 significant bits
}

}  // namespace cpt
