// aux_kernels.cu -- small per-frame kernels beside the persistent extraction kernel:
//   background_step_kernel   WeightedBackground.process_frame on an arbitrary frame  (K7, motiondetector.py:197-244)
//   frame_median_kernel      np.median of a uint16 frame                             (K8, clip.py:474-487; interpreter.py:389)
#include "background_step.cuh"

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
// frame: the frame is staged in shared memory once, then two 256-bin histogram passes (high byte,
// then low byte inside the selected bucket) select the ranks (n-1)/2 and n/2.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void select_bucket(const uint32_t *hist, int rank, int &bucket, int &rank_in_bucket) {
    // called by one thread: 256 bins
    int acc = 0;
    for (int b = 0; b < 256; ++b) {
        int c = (int)hist[b];
        if (rank < acc + c) {
            bucket = b;
            rank_in_bucket = rank - acc;
            return;
        }
        acc += c;
    }
    bucket = 255;
    rank_in_bucket = 0;
}

__global__ void __launch_bounds__(256) frame_median_kernel(const uint16_t *frames, int npx, float *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t *px = reinterpret_cast<uint16_t *>(smem_raw);
    __shared__ uint32_t hist[2][256];
    __shared__ int sel[2][2];  // [which rank][bucket, rank in bucket]
    const uint16_t *src = frames + (size_t)blockIdx.x * npx;
    const int tid = threadIdx.x;
    hist[0][tid] = 0;
    hist[1][tid] = 0;
    __syncthreads();
    for (int i = tid; i < npx / 8; i += blockDim.x) {
        uint4 v = ldg16(src + i * 8);
        *reinterpret_cast<uint4 *>(px + i * 8) = v;
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            atomicAdd(&hist[0][(w[q] >> 8) & 0xff], 1u);
            atomicAdd(&hist[0][w[q] >> 24], 1u);
        }
    }
    for (int i = (npx / 8) * 8 + tid; i < npx; i += blockDim.x) {
        uint16_t v = src[i];
        px[i] = v;
        atomicAdd(&hist[0][v >> 8], 1u);
    }
    __syncthreads();
    if (tid < 2) select_bucket(hist[0], tid == 0 ? (npx - 1) / 2 : npx / 2, sel[tid][0], sel[tid][1]);
    __syncthreads();
    const int b_lo = sel[0][0], b_hi = sel[1][0];
    hist[0][tid] = 0;  // reuse: low-byte histogram of bucket b_lo; hist[1]: of bucket b_hi
    __syncthreads();
    for (int i = tid; i < npx; i += blockDim.x) {
        const int v = px[i], hb = v >> 8;
        if (hb == b_lo) atomicAdd(&hist[0][v & 0xff], 1u);
        if (hb == b_hi) atomicAdd(&hist[1][v & 0xff], 1u);
    }
    __syncthreads();
    if (tid == 0) {
        int lo_b, hi_b, r;
        select_bucket(hist[0], sel[0][1], lo_b, r);
        select_bucket(hist[1], sel[1][1], hi_b, r);
        const int v_lo = (b_lo << 8) | lo_b, v_hi = (b_hi << 8) | hi_b;
        out[blockIdx.x] = 0.5f * (float)(v_lo + v_hi);  // exact: at most 17 significant bits
    }
}

}  // namespace cpt
