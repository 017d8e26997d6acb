// nlm_kernel.cu -- cv2.fastNlMeansDenoising(uint8, None) (K3 of SURVEY.md section 8a / Appendix B; the call site is
// track/cliptracker.py:116-117, active when TrackingConfig.denoise is set): h = 3, 7x7 template, 21x21 search.
// OpenCV's integer algorithm, bit for bit: for every pixel p and search offset o in [-10, 10]^2
//     dist = sum over the 7x7 template of (ext[p + o + t] - ext[p + t])^2        (ext: BORDER_REFLECT_101 by 13)
//     w = table[dist >> 6],  est += w * ext[p + o],  wsum += w;      out = (est + wsum / 2) / wsum
// with table[a] = rint(19096 * exp(-(a * 64 / 49) / 9)), zero below 0.001 * 19096 -- only a <= 47 is non-zero.
// One CTA per 16x16 tile and frame: the 42x42 neighbourhood sits in shared memory; per offset the squared
// differences of the tile + template border are formed once and box-summed separably (4 ops per pixel and offset
// instead of 49).  Compute bound by design (441 offsets): this is the stand-alone primitive; fusing it into the
// persistent extraction kernel is the next step.
#include <algorithm>
#include <cmath>

#include "cptrack_internal.cuh"

namespace cpt {

constexpr int kNlmTile = 16, kNlmT = 3, kNlmS = 10, kNlmB = kNlmT + kNlmS;  // template / search / border radii
constexpr int kNlmExt = kNlmTile + 2 * kNlmB;                               // 42
constexpr int kNlmSq = kNlmTile + 2 * kNlmT;                                // 22
constexpr int kNlmWeights = 64;

__constant__ int c_nlm_weights[kNlmWeights];

__device__ __forceinline__ int nlm_reflect(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * n - 2 - p;
    return p;
}

// info != nullptr: only the frames the extraction kernel marked for denoising (reserved[1] != 0) are processed
__global__ void __launch_bounds__(256) nlm_denoise_kernel(const uint8_t *src, int W, int H, uint8_t *dst, const cpt_frame_info *info) {
    if (info && info[blockIdx.z].reserved[1] == 0) return;
    __shared__ uint8_t ext[kNlmExt][kNlmExt + 2];
    __shared__ uint32_t sq[kNlmSq][kNlmSq + 1];
    __shared__ uint32_t hs[kNlmSq][kNlmTile + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int x0 = blockIdx.x * kNlmTile, y0 = blockIdx.y * kNlmTile;
    const uint8_t *img = src + (size_t)blockIdx.z * W * H;
    for (int i = tid; i < kNlmExt * kNlmExt; i += 256) {
        const int ey = i / kNlmExt, ex = i - ey * kNlmExt;
        ext[ey][ex] = img[nlm_reflect(y0 + ey - kNlmB, H) * W + nlm_reflect(x0 + ex - kNlmB, W)];
    }
    __syncthreads();
    uint32_t est = 0, wsum = 0;
    for (int oy = -kNlmS; oy <= kNlmS; ++oy)
        for (int ox = -kNlmS; ox <= kNlmS; ++ox) {
            // squared differences over the tile grown by the template radius
            for (int i = tid; i < kNlmSq * kNlmSq; i += 256) {
                const int sy = i / kNlmSq, sx = i - sy * kNlmSq;
                const int d = (int)ext[sy + kNlmS + oy][sx + kNlmS + ox] - (int)ext[sy + kNlmS][sx + kNlmS];
                sq[sy][sx] = (uint32_t)(d * d);
            }
            __syncthreads();
            // horizontal 7-sums
            for (int i = tid; i < kNlmSq * kNlmTile; i += 256) {
                const int sy = i >> 4, sx = i & 15;
                uint32_t acc = 0;
#pragma unroll
                for (int j = 0; j < 2 * kNlmT + 1; ++j) acc += sq[sy][sx + j];
                hs[sy][sx] = acc;
            }
            __syncthreads();
            uint32_t dist = 0;
#pragma unroll
            for (int j = 0; j < 2 * kNlmT + 1; ++j) dist += hs[ty + j][tx];
            const uint32_t a = dist >> 6;
            if (a < (uint32_t)kNlmWeights) {
                const uint32_t w = (uint32_t)c_nlm_weights[a];
                est += w * ext[ty + kNlmB + oy][tx + kNlmB + ox];
                wsum += w;
            }
            // (hs is rewritten only after the next offset's first barrier, sq only after this offset's second)
        }
    const int x = x0 + tx, y = y0 + ty;
    if (x < W && y < H) dst[(size_t)blockIdx.z * W * H + y * W + x] = (uint8_t)((est + wsum / 2) / wsum);
}

}  // namespace cpt

using cpt::fail;

namespace cpt {

int nlm_prepare(cpt_ctx *c) {
    if (!c->nlm_table_ready) {
        // OpenCV FastNlMeansDenoisingInvoker: fixed_point_mult = INT_MAX / (21 * 21 * 255); template 49 -> shift 6
        const int fpm = 2147483647 / (21 * 21 * 255);
        const double mult = 64.0 / 49.0;
        int table[cpt::kNlmWeights];
        for (int a = 0; a < cpt::kNlmWeights; ++a) {
            const double w = std::exp(-(a * mult) / 9.0);
            int wi = (int)std::rint(fpm * w);
            if (wi < 0.001 * fpm) wi = 0;
            table[a] = wi;
        }
        if (table[cpt::kNlmWeights - 1] != 0) return fail(CPT_ERR_UNSUPPORTED, "internal: NLM weight table is longer than expected");
        CUDA_TRY(cudaMemcpyToSymbol(cpt::c_nlm_weights, table, sizeof(table)));
        c->nlm_table_ready = true;
    }
    return CPT_OK;
}

int nlm_launch(cpt_ctx *c, const uint8_t *d_src, int width, int height, long long n_frames, uint8_t *d_dst, const cpt_frame_info *info,
               cudaStream_t stream) {
    int rc = nlm_prepare(c);
    if (rc) return rc;
    for (long long f0 = 0; f0 < n_frames; f0 += 65535) {  // gridDim.z limit
        const int nz = (int)std::min<long long>(65535, n_frames - f0);
        dim3 grid((width + kNlmTile - 1) / kNlmTile, (height + kNlmTile - 1) / kNlmTile, nz);
        nlm_denoise_kernel<<<grid, 256, 0, stream>>>(d_src + (size_t)f0 * width * height, width, height, d_dst + (size_t)f0 * width * height,
                                                       info ? info + f0 : nullptr);
    }
    CUDA_TRY(cudaGetLastError());
    return CPT_OK;
}

}  // namespace cpt

extern "C" {

int cpt_nlm_denoise_u8(cpt_ctx *c, const uint8_t *d_src, int width, int height, int n_frames, uint8_t *d_dst) {
    if (!c || !d_src || !d_dst) return fail(CPT_ERR_INVALID, "null argument");
    if (width < 1 || height < 1 || n_frames < 0) return fail(CPT_ERR_INVALID, "bad size");
    if (n_frames == 0) return CPT_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    return cpt::nlm_launch(c, d_src, width, height, n_frames, d_dst, nullptr, c->stream);
}

}  // extern "C"
