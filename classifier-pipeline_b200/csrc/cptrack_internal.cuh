// cptrack_internal.cuh -- what the translation units behind the C ABI share: the context record,
// error reporting and the CUDA_TRY macro.  Not part of the public interface (include/cptrack.h).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <vector>

#include "cptrack_kernels.cuh"

namespace cpt {

char *error_buffer();  // thread local, 512 bytes (cptrack.cu)

inline int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

struct HostWeightTable {
    std::vector<double> w;
    uint32_t *d_thr = nullptr;
    int max_count = 0;
    int has_bounds = 0;
    int max_bound = 0;
    int linear_upto = 0;
    cpt::WeightTable device() const { return cpt::WeightTable{d_thr, max_count, has_bounds, max_bound, linear_upto}; }
};

}  // namespace cpt

#define CUDA_TRY(expr)                                                                               \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess) return cpt::fail(CPT_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

struct cpt_ctx {
    int device = 0;
    cpt::Geometry g{};
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
    int num_sms = 0;
    cpt::HostWeightTable tables[4];
    float *scratch = nullptr;
    size_t scratch_ctas = 0;
    // scratch of the split extraction path (grown on demand): quad bytes, strip pass records, per-frame headers and the
    // thresholded masks between frame_regions_kernel and frame_components_kernel
    int8_t *qbytes = nullptr;
    cpt::StripRec *prec = nullptr;
    cpt::FrameHdr *fhdr = nullptr;
    uint32_t *maskbits = nullptr;
    int *fallback = nullptr;   // [1 + split_frames]: frames left to frame_components_kernel (count first)
    size_t split_frames = 0, split_clips = 0;
    bool force_single = false;  // cpt_debug_force_single_kernel
    bool time_kernels = false;  // cpt_debug_kernel_times
    bool timed_valid = false;
    cudaEvent_t ev_k[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int *work_counter = nullptr;
    uint16_t *zero_frame = nullptr;
    long long *debug = nullptr;
    // staging buffers of cpt_extract_batch_host
    void *stage_frames[2] = {nullptr, nullptr};
    size_t stage_frames_bytes = 0;
    cpt_region *stage_regions[2] = {nullptr, nullptr};
    cpt_frame_info *stage_info[2] = {nullptr, nullptr};
    float *stage_filtered[2] = {nullptr, nullptr};
    uint8_t *stage_labels[2] = {nullptr, nullptr};
    size_t stage_out_frames = 0;
    bool stage_has_filtered = false, stage_has_labels = false;
    cpt_clip *d_clips = nullptr;
    size_t d_clips_cap = 0;
    cudaEvent_t ev_h2d[2], ev_compute[2], ev_d2h[2];
    bool events = false;
    // scratch of cpt_detect_objects_u8 (grown on demand)
    void *detect_scratch = nullptr;
    size_t detect_scratch_bytes = 0;
    uint8_t *u8_frames[2] = {nullptr, nullptr};  // normalised / denoised images of the denoise pipeline
    size_t u8_bytes = 0;
    void *cptv_scratch = nullptr;  // per-frame change images of cpt_cptv_decode
    size_t cptv_scratch_bytes = 0;
    // staging of cpt_extract_batch_cptv_host: packed stream bytes, frame table rows and clip row offsets per buffer
    uint8_t *pk_stream[2] = {nullptr, nullptr};
    cpt_cptv_frame *pk_table[2] = {nullptr, nullptr}, *pk_h_table[2] = {nullptr, nullptr};
    int32_t *pk_first[2] = {nullptr, nullptr}, *pk_h_first[2] = {nullptr, nullptr};
    int32_t *pk_change = nullptr;
    size_t pk_bytes = 0, pk_rows = 0, pk_clips = 0;
    float *scratch_filtered = nullptr;  // filtered images of host-staged calls that do not return them
    size_t scratch_filtered_frames = 0;
    bool nlm_table_ready = false;  // cpt_nlm_denoise_u8 uploaded its weight table (constant memory of this device)
};

