// motion_kernels.cu -- CPTVMotionDetector.process_frame (piclassifier/cptvmotiondetector.py:74-205) as ONE launch
// per frame (M1 of SURVEY.md section 8a): batch-1 latency is the metric, so the running mean, the weighted
// background update and the motion count are fused into a single 1024-thread CTA and the host reads back one
// 32-byte result record.
//   phase 1  RunningMean.add (motiondetector.py:160-175): uint32 sum += new - oldest-in-ring (modular, as numpy)
//   phase 2  WeightedBackground.process_frame(sum / n) (motiondetector.py:197-244) unless FFC affected
//   phase 3  detect(): count(clip(cur, T) - clip(oldest_nonffc, T) > delta_thresh), T = background average
//            (cptvmotiondetector.py:74-120), with the optional second delta ring (one_diff_only False)
// The frame ring, the sum and the delta ring live in HBM for the life of the detector.
#include <cstring>

#include "background_step.cuh"
#include "cptrack_internal.cuh"

struct cpt_motion {
    cpt_ctx *ctx = nullptr;
    int ring_frames = 0, mean_frames = 0, diff_frames = 0, weight_slot = 0;
    uint16_t *d_ring = nullptr;     // [ring_frames][H][W]
    uint32_t *d_sum = nullptr;      // [H][W]
    double *d_diff = nullptr;       // [diff_frames][crop_h][crop_w]
    int32_t *d_mean_count = nullptr;
    cpt_motion_result *d_result = nullptr;
    cpt_motion_result *h_result = nullptr;  // pinned
    uint16_t *h_stage = nullptr;            // pinned frame
    bool mean_started = false;
};

namespace cpt {

struct MotionArgs {
    Geometry g;
    uint16_t *ring;
    uint32_t *sum;
    double *diff;
    int32_t *mean_count;
    uint8_t *bg_state;
    WeightTable wt;
    cpt_motion_result *result;
    int slot_new, slot_oldest, slot_nonffc;
    int diff_slot_new, diff_slot_old;
    int mean_frames;
    uint32_t flags;
    int delta_thresh;
    double init_average;  // WeightedBackground(init_average=...) until the background has seen a frame
};

// RunningMean(frames, window) constructor (motiondetector.py:161-164): sum of the listed ring slots
__global__ void __launch_bounds__(256) motion_mean_init_kernel(const uint16_t *ring, int npx, const int *slots, int n,
                                                               uint32_t *sum, int32_t *mean_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npx) {
        uint32_t s = 0;
        for (int k = 0; k < n; ++k) s += ring[(size_t)slots[k] * npx + i];
        sum[i] = s;
    }
    if (i == 0) *mean_count = n;
}

__global__ void __launch_bounds__(1024, 1) motion_step_kernel(const MotionArgs a) {
    __shared__ unsigned long long red_sum[32];
    __shared__ int red_changed;
    __shared__ int red_count[32];
    __shared__ int s_error;
    const Geometry &g = a.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint16_t *cur = a.ring + (size_t)a.slot_new * g.npx;
    if (tid == 0) s_error = 0;
    __syncthreads();

    // ---- phase 1: running mean
    int n = *a.mean_count;
    if (a.flags & CPT_MOTION_MEAN) {
        const bool restart = (a.flags & CPT_MOTION_MEAN_RESTART) || n == 0;
        const bool full = !restart && n == a.mean_frames;
        const uint16_t *old = a.ring + (size_t)(a.slot_oldest < 0 ? a.slot_new : a.slot_oldest) * g.npx;
        for (int i = tid; i < g.npx / 2; i += blockDim.x) {
            const uint32_t c2 = reinterpret_cast<const uint32_t *>(cur)[i];
            uint2 s = reinterpret_cast<uint2 *>(a.sum)[i];
            if (restart) {
                s.x = c2 & 0xffffu;
                s.y = c2 >> 16;
            } else {
                if (full) {  // `running_mean -= oldest; running_mean += new`, uint32 modular
                    const uint32_t o2 = reinterpret_cast<const uint32_t *>(old)[i];
                    s.x -= o2 & 0xffffu;
                    s.y -= o2 >> 16;
                }
                s.x += c2 & 0xffffu;
                s.y += c2 >> 16;
            }
            reinterpret_cast<uint2 *>(a.sum)[i] = s;
        }
        n = restart ? 1 : (full ? n : n + 1);
    }
    __syncthreads();
    if (tid == 0) *a.mean_count = n;

    // ---- phase 2: background from the running mean (np.int32(sum / n): truncation == floor division here)
    if ((a.flags & CPT_MOTION_BACKGROUND) && n > 0) {
        const uint32_t *S = a.sum;
        const uint32_t un = (uint32_t)n;
        int *err = &s_error;
        background_step(g, a.bg_state, [S, un, err](int p) {
            const uint32_t v = S[p] / un;
            if (v > 65535u) { *err = 1; return 65535; }
            return (int)v;
        }, a.wt, red_sum, &red_changed);
    }

    // ---- phase 3: motion count against the oldest non-FFC frame
    int count = 0;
    const StateHeader *bg_hdr = reinterpret_cast<const StateHeader *>(a.bg_state);
    const double T = bg_hdr->initialised ? bg_hdr->average : a.init_average;
    if (a.flags & CPT_MOTION_DETECT) {
        const uint16_t *old = a.ring + (size_t)a.slot_nonffc * g.npx;
        const double dt = (double)a.delta_thresh;
        const bool warmer = a.flags & CPT_MOTION_WARMER_ONLY, one = a.flags & CPT_MOTION_ONE_DIFF;
        double *dnew = a.diff ? a.diff + (size_t)a.diff_slot_new * g.ncrop : nullptr;
        const double *dold = (a.diff && a.diff_slot_old >= 0) ? a.diff + (size_t)a.diff_slot_old * g.ncrop : nullptr;
        for (int i = tid; i < g.ncrop; i += blockDim.x) {
            const int y = i / g.crop_w + g.edge, x = i - (y - g.edge) * g.crop_w + g.edge;
            const int p = y * g.W + x;
            double d = fmax((double)cur[p], T) - fmax((double)old[p], T);
            if (!warmer) d = fabs(d);
            if (one) {
                count += d > dt;
            } else {
                if (d >= dt) d = dt;
                if (dold) count += (dold[i] + d == 2.0 * dt);
                dnew[i] = d;
            }
        }
    }
    count = __reduce_add_sync(0xffffffffu, count);
    if (lane == 0) red_count[warp] = count;
    __syncthreads();
    if (tid == 0) {
        int total = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += red_count[w];
        cpt_motion_result r;
        r.average = T;
        r.diff = total;
        r.error = s_error;
        r.mean_frames = n;
        r.reserved = 0;
        *a.result = r;
    }
}

}  // namespace cpt

using cpt::fail;

extern "C" {

cpt_motion *cpt_motion_open(cpt_ctx *c, int ring_frames, int mean_frames, int diff_frames, int weight_slot) {
    if (!c || ring_frames < 1 || mean_frames < 1 || diff_frames < 0 || weight_slot < 0 || weight_slot >= 4) {
        fail(CPT_ERR_INVALID, "cpt_motion_open: bad argument");
        return nullptr;
    }
    if (!c->tables[weight_slot].d_thr) {
        fail(CPT_ERR_INVALID, "weight table slot %d is not set (cpt_set_weight_table)", weight_slot);
        return nullptr;
    }
    if (cudaSetDevice(c->device) != cudaSuccess) {
        fail(CPT_ERR_CUDA, "cudaSetDevice failed");
        return nullptr;
    }
    cpt_motion *m = new cpt_motion();
    m->ctx = c;
    m->ring_frames = ring_frames; m->mean_frames = mean_frames; m->diff_frames = diff_frames; m->weight_slot = weight_slot;
    const size_t npx = c->g.npx;
    bool ok = cudaMalloc(&m->d_ring, (size_t)ring_frames * npx * sizeof(uint16_t)) == cudaSuccess &&
              cudaMalloc(&m->d_sum, npx * sizeof(uint32_t)) == cudaSuccess &&
              cudaMalloc(&m->d_mean_count, sizeof(int32_t)) == cudaSuccess &&
              cudaMalloc(&m->d_result, sizeof(cpt_motion_result)) == cudaSuccess &&
              cudaHostAlloc(&m->h_result, sizeof(cpt_motion_result), cudaHostAllocDefault) == cudaSuccess &&
              cudaHostAlloc(&m->h_stage, npx * sizeof(uint16_t), cudaHostAllocDefault) == cudaSuccess;
    if (ok && diff_frames > 0) ok = cudaMalloc(&m->d_diff, (size_t)diff_frames * c->g.ncrop * sizeof(double)) == cudaSuccess;
    if (ok) {
        ok = cudaMemset(m->d_ring, 0, (size_t)ring_frames * npx * sizeof(uint16_t)) == cudaSuccess &&
             cudaMemset(m->d_sum, 0, npx * sizeof(uint32_t)) == cudaSuccess &&
             cudaMemset(m->d_mean_count, 0, sizeof(int32_t)) == cudaSuccess;
        if (ok && m->d_diff) ok = cudaMemset(m->d_diff, 0, (size_t)diff_frames * c->g.ncrop * sizeof(double)) == cudaSuccess;
    }
    if (!ok) {
        fail(CPT_ERR_NOMEM, "cpt_motion_open: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        cpt_motion_close(m);
        return nullptr;
    }
    return m;
}

void cpt_motion_close(cpt_motion *m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(m->d_ring);
    cudaFree(m->d_sum);
    cudaFree(m->d_diff);
    cudaFree(m->d_mean_count);
    cudaFree(m->d_result);
    cudaFreeHost(m->h_result);
    cudaFreeHost(m->h_stage);
    delete m;
}

int cpt_motion_store(cpt_motion *m, const uint16_t *h_pix, int slot) {
    if (!m || !h_pix) return fail(CPT_ERR_INVALID, "null argument");
    if (slot < 0 || slot >= m->ring_frames) return fail(CPT_ERR_INVALID, "ring slot %d out of range", slot);
    cpt_ctx *c = m->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t bytes = (size_t)c->g.npx * sizeof(uint16_t);
    CUDA_TRY(cudaStreamSynchronize(c->stream));  // the staging buffer is free again
    memcpy(m->h_stage, h_pix, bytes);
    CUDA_TRY(cudaMemcpyAsync(m->d_ring + (size_t)slot * c->g.npx, m->h_stage, bytes, cudaMemcpyHostToDevice, c->stream));
    return CPT_OK;
}

int cpt_motion_mean_init(cpt_motion *m, const int32_t *h_slots, int n) {
    if (!m || !h_slots) return fail(CPT_ERR_INVALID, "null argument");
    if (n < 1 || n > m->mean_frames || n > 64) return fail(CPT_ERR_INVALID, "bad frame count %d", n);
    for (int k = 0; k < n; ++k)
        if (h_slots[k] < 0 || h_slots[k] >= m->ring_frames) return fail(CPT_ERR_INVALID, "ring slot out of range");
    cpt_ctx *c = m->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    int *d_slots = nullptr;
    CUDA_TRY(cudaMalloc(&d_slots, sizeof(int) * n));
    cudaError_t e = cudaMemcpyAsync(d_slots, h_slots, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        cpt::motion_mean_init_kernel<<<(c->g.npx + 255) / 256, 256, 0, c->stream>>>(m->d_ring, c->g.npx, d_slots, n, m->d_sum, m->d_mean_count);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_slots);
    if (e != cudaSuccess) return fail(CPT_ERR_CUDA, "cpt_motion_mean_init: %s", cudaGetErrorString(e));
    return CPT_OK;
}

int cpt_motion_step(cpt_motion *m, const uint16_t *h_pix, void *d_background_state, int slot_new, int slot_oldest,
                    int slot_nonffc, int diff_slot_new, int diff_slot_old, uint32_t flags, int delta_thresh,
                    double init_average, cpt_motion_result *h_result) {
    if (!m || !h_pix || !d_background_state || !h_result) return fail(CPT_ERR_INVALID, "null argument");
    if (slot_new < 0 || slot_new >= m->ring_frames || slot_oldest >= m->ring_frames) return fail(CPT_ERR_INVALID, "ring slot out of range");
    if (flags & CPT_MOTION_DETECT) {
        if (slot_nonffc < 0 || slot_nonffc >= m->ring_frames) return fail(CPT_ERR_INVALID, "oldest non-FFC slot out of range");
        if (!(flags & CPT_MOTION_ONE_DIFF)) {
            if (!m->d_diff) return fail(CPT_ERR_INVALID, "the detector was opened without a delta ring (one_diff_only)");
            if (diff_slot_new < 0 || diff_slot_new >= m->diff_frames || diff_slot_old >= m->diff_frames)
                return fail(CPT_ERR_INVALID, "delta ring slot out of range");
        }
    }
    int rc = cpt_motion_store(m, h_pix, slot_new);
    if (rc) return rc;
    cpt_ctx *c = m->ctx;
    cpt::MotionArgs a{};
    a.g = c->g;
    a.ring = m->d_ring; a.sum = m->d_sum; a.diff = m->d_diff; a.mean_count = m->d_mean_count;
    a.bg_state = (uint8_t *)d_background_state;
    a.wt = c->tables[m->weight_slot].device();
    a.result = m->d_result;
    a.slot_new = slot_new; a.slot_oldest = slot_oldest; a.slot_nonffc = slot_nonffc;
    a.diff_slot_new = diff_slot_new; a.diff_slot_old = diff_slot_old;
    a.mean_frames = m->mean_frames;
    a.flags = flags;
    a.delta_thresh = delta_thresh;
    a.init_average = init_average;
    cpt::motion_step_kernel<<<1, 1024, 0, c->stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(m->h_result, m->d_result, sizeof(cpt_motion_result), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *h_result = *m->h_result;
    if (h_result->error) return fail(CPT_ERR_INVALID, "running mean left the uint16 range (inconsistent ring / mean window)");
    return CPT_OK;
}

}  // extern "C"
