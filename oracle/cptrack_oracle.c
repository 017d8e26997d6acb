/*
 * cptrack_oracle.c -- CPU restatement of classifier-pipeline's track-extraction and
 * classifier-input preprocessing arithmetic.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is the checker the CUDA path is compared with
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference leg).  Nothing
 * in the product package imports, links or executes it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against
 * fixtures produced by running the unmodified Python reference in the build container
 * (tests/golden/make_golden.py): per-frame background, normalised uint8 image (CRC),
 * threshold, label image, component stats/centroids, region lists and final
 * WeightedBackground state, for the reference's two test clips and seeded synthetic clips,
 * with and without NLM denoising.
 *
 * Every function cites the reference file:line (relative to the reference's src/) it
 * follows.  Scalar, single-threaded per clip; orc_extract_batch runs clips in parallel with
 * OpenMP (mirrors the reference's multiprocessing.Pool over files,
 * track/trackextractor.py:80-85).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MEAN_FRAMES 45 /* track/cliptrackextractor.py:173 get_last_x(x=45) */

/* ------------------------------------------------------------------------------------------
 * small helpers
 * ---------------------------------------------------------------------------------------- */
static inline int reflect101(int p, int n) {
    /* cv2 BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba */
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * n - 2 - p;
    }
    return p;
}

/* Python round(): half to even on a double -> nearest integer (cliptracker.py:103-105,
 * motiondetector.py:232).  rint() under the default rounding mode is exactly that. */
static inline long py_round(double v) { return (long)rint(v); }

/* ------------------------------------------------------------------------------------------
 * K2  ClipTracker._get_filtered_frame (track/cliptracker.py:93-122) with
 *     imageprocessing.normalize (ml_tools/imageprocessing.py:151-169)
 *
 *   avg_change = int(round(np.average(thermal) - background_alg.get_average()))
 *   G = clip(float32(thermal) - background - avg_change, 0, None)           (stored as fp32)
 *   N = 255 * (G - min) / (max - min)    in fp32: multiply first, then IEEE divide
 *   threshold = background_thresh / (max - min) * 255   (fp32 under numpy 2 promotion)
 *   degenerate max == min: N = 0 (max == 0) or G / max; threshold = background_thresh
 * U = uint8(N) (truncation) is what detect_objects then sees (imageprocessing.py:241).
 * ---------------------------------------------------------------------------------------- */
void orc_normalise_frame(const uint16_t *pix, const int32_t *bg, double bg_average, int n_px,
                         int background_thresh, uint8_t *u8_out, float *thresh_out,
                         float *max_out, float *min_out, int32_t *avg_change_out) {
    uint64_t sum = 0;
    for (int i = 0; i < n_px; i++) sum += pix[i];
    double mean = (double)sum / (double)n_px;
    int32_t avg_change = (int32_t)py_round(mean - bg_average);
    float mx = -INFINITY, mn = INFINITY;
    for (int i = 0; i < n_px; i++) {
        double g = (double)pix[i] - (double)bg[i] - (double)avg_change;
        float gf = (float)(g < 0 ? 0 : g);
        if (gf > mx) mx = gf;
        if (gf < mn) mn = gf;
    }
    float thresh;
    if (mx == mn) {
        for (int i = 0; i < n_px; i++) u8_out[i] = (mx == 0.0f) ? 0 : 1; /* G / max == 1 */
        thresh = (float)background_thresh;
    } else {
        float range = mx - mn;
        for (int i = 0; i < n_px; i++) {
            double g = (double)pix[i] - (double)bg[i] - (double)avg_change;
            float gf = (float)(g < 0 ? 0 : g);
            volatile float num = 255.0f * (gf - mn); /* volatile: no fma/contraction */
            float nrm = num / range;
            u8_out[i] = (uint8_t)nrm;
        }
        volatile float q = (float)background_thresh / range;
        thresh = q * 255.0f;
    }
    *thresh_out = thresh;
    *max_out = mx;
    *min_out = mn;
    *avg_change_out = avg_change;
}

/* ------------------------------------------------------------------------------------------
 * K4  imageprocessing.detect_objects (ml_tools/imageprocessing.py:240-247), the three OpenCV
 *     calls before labelling.  OpenCV 4.x semantics (third-party, pinned ~=4.12; verified
 *     against 4.13 through the golden label images):
 *   GaussianBlur(u8,(5,5),0): taps [1,4,6,4,1]/16 in fixed point, one rounding:
 *                             (sum_ij k_i k_j U + 128) >> 8, BORDER_REFLECT_101
 *   threshold(THRESH_BINARY): U > floor(thresh)
 *   morphologyEx(MORPH_CLOSE, (5,5)): the *tuple* (5,5) is taken as a 2x1 structuring
 *       element anchored at row 1: dilate D[y]=max(M[y],M[y-1]); erode C[y]=min(D[y],D[y-1]);
 *       out-of-image rows do not contribute.
 * ---------------------------------------------------------------------------------------- */
void orc_blur5(const uint8_t *src, int W, int H, uint8_t *dst) {
    static const int k[5] = {1, 4, 6, 4, 1};
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int acc = 0;
            for (int j = -2; j <= 2; j++) {
                const uint8_t *row = src + (size_t)reflect101(y + j, H) * W;
                int racc = 0;
                for (int i = -2; i <= 2; i++) racc += k[i + 2] * row[reflect101(x + i, W)];
                acc += k[j + 2] * racc;
            }
            dst[(size_t)y * W + x] = (uint8_t)((acc + 128) >> 8);
        }
}

void orc_threshold_close(const uint8_t *blurred, int W, int H, float thresh, uint8_t *mask) {
    int ith = (int)floorf(thresh);
    uint8_t *m = (uint8_t *)malloc((size_t)W * H);
    for (int i = 0; i < W * H; i++) m[i] = (ith < 0) ? 1 : (ith >= 255 ? 0 : (blurred[i] > ith));
    uint8_t *d = (uint8_t *)malloc((size_t)W * H);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            uint8_t v = m[y * W + x];
            if (y > 0 && m[(y - 1) * W + x] > v) v = m[(y - 1) * W + x];
            d[y * W + x] = v;
        }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            uint8_t v = d[y * W + x];
            if (y > 0 && d[(y - 1) * W + x] < v) v = d[(y - 1) * W + x];
            mask[y * W + x] = v;
        }
    free(m);
    free(d);
}

/* ------------------------------------------------------------------------------------------
 * K5  cv2.connectedComponentsWithStats(mask) (imageprocessing.py:248): 8-connectivity.
 *     OpenCV numbers labels in the order its 2x2-block raster scan first meets each
 *     component, i.e. by increasing  key = (y/2) * ceil(W/2) + (x/2)  of the component's
 *     first block.  stats row = [left, top, width, height, area]; centroid = (sum x / area,
 *     sum y / area) in double.  Row 0 (background) is not produced here.
 *  comp_out rows: left, top, width, height, area, sum_x, sum_y, key   (int32 x 8)
 *  labels_out: int32 label image (0 = background).  Returns the component count.
 * ---------------------------------------------------------------------------------------- */
static int uf_find(int32_t *p, int a) {
    while (p[a] != a) {
        p[a] = p[p[a]];
        a = p[a];
    }
    return a;
}

typedef struct {
    int32_t key, root;
} orc_keyroot;

static int cmp_keyroot(const void *a, const void *b) {
    const orc_keyroot *x = (const orc_keyroot *)a, *y = (const orc_keyroot *)b;
    return (x->key > y->key) - (x->key < y->key);
}

int orc_cc8(const uint8_t *mask, int W, int H, int32_t *labels_out, int32_t *comp_out,
            int max_comp) {
    int n = W * H;
    int32_t *parent = (int32_t *)malloc(sizeof(int32_t) * n);
    for (int i = 0; i < n; i++) parent[i] = i;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            if (!mask[y * W + x]) continue;
            int me = y * W + x;
            static const int dx[4] = {-1, -1, 0, 1}, dy[4] = {0, -1, -1, -1};
            for (int q = 0; q < 4; q++) {
                int xx = x + dx[q], yy = y + dy[q];
                if (xx < 0 || xx >= W || yy < 0) continue;
                if (!mask[yy * W + xx]) continue;
                int a = uf_find(parent, me), b = uf_find(parent, yy * W + xx);
                if (a != b) {
                    if (a < b) parent[b] = a;
                    else parent[a] = b;
                }
            }
        }
    int bw = (W + 1) / 2;
    int32_t *minkey = (int32_t *)malloc(sizeof(int32_t) * n);
    for (int i = 0; i < n; i++) minkey[i] = INT32_MAX;
    int ncomp = 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            if (!mask[y * W + x]) continue;
            int r = uf_find(parent, y * W + x);
            int key = (y / 2) * bw + (x / 2);
            if (minkey[r] == INT32_MAX) ncomp++;
            if (key < minkey[r]) minkey[r] = key;
        }
    orc_keyroot *order = (orc_keyroot *)malloc(sizeof(orc_keyroot) * (ncomp > 0 ? ncomp : 1));
    int c = 0;
    for (int i = 0; i < n; i++)
        if (minkey[i] != INT32_MAX) {
            order[c].key = minkey[i];
            order[c].root = i;
            c++;
        }
    qsort(order, ncomp, sizeof(orc_keyroot), cmp_keyroot);
    int32_t *root_label = minkey; /* reuse: root index -> 1-based label */
    for (int i = 0; i < ncomp; i++) root_label[order[i].root] = i + 1;
    int stored = ncomp < max_comp ? ncomp : max_comp;
    for (int i = 0; i < stored; i++) {
        int32_t *row = comp_out + i * 8;
        row[0] = INT32_MAX; row[1] = INT32_MAX; row[2] = -1; row[3] = -1;
        row[4] = 0; row[5] = 0; row[6] = 0; row[7] = order[i].key;
    }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int lab = 0;
            if (mask[y * W + x]) {
                lab = root_label[uf_find(parent, y * W + x)];
                if (lab <= stored) {
                    int32_t *row = comp_out + (lab - 1) * 8;
                    if (x < row[0]) row[0] = x;
                    if (y < row[1]) row[1] = y;
                    if (x > row[2]) row[2] = x; /* right (inclusive) for now */
                    if (y > row[3]) row[3] = y;
                    row[4] += 1; row[5] += x; row[6] += y;
                }
            }
            if (labels_out) labels_out[y * W + x] = lab;
        }
    for (int i = 0; i < stored; i++) {
        int32_t *row = comp_out + i * 8;
        row[2] = row[2] - row[0] + 1;
        row[3] = row[3] - row[1] + 1;
    }
    free(order);
    free(minkey);
    free(parent);
    return ncomp;
}

/* ------------------------------------------------------------------------------------------
 * K3  cv2.fastNlMeansDenoising(u8, None) (track/cliptracker.py:116-117), defaults h=3,
 *     templateWindowSize=7, searchWindowSize=21.  OpenCV (third-party) integer algorithm,
 *     restated:
 *       ext = src padded by 13 with BORDER_REFLECT_101
 *       fixed_point_mult = INT_MAX / (21*21*255); almost_template_window_size_sq_bin_shift = 6
 *       weight table W[a] = round(fpm * exp(-(a * 64/49) / (h*h))), zero below 0.001*fpm
 *       for each pixel, each of the 441 search offsets: dist = sum over 7x7 of squared
 *       differences; w = W[dist >> 6]; est += w * ext[p+off]; wsum += w
 *       out = (est + wsum/2) / wsum
 * ---------------------------------------------------------------------------------------- */
void orc_nlm_denoise(const uint8_t *src, int W, int H, uint8_t *dst) {
    const int tr = 3, sr = 10, br = tr + sr; /* template, search, border radii */
    const int tw = 7, sw = 21;
    int EW = W + 2 * br, EH = H + 2 * br;
    uint8_t *ext = (uint8_t *)malloc((size_t)EW * EH);
    for (int y = 0; y < EH; y++)
        for (int x = 0; x < EW; x++)
            ext[y * EW + x] = src[reflect101(y - br, H) * W + reflect101(x - br, W)];
    const int fpm = INT32_MAX / (sw * sw * 255);
    int tws_sq = tw * tw, shift = 0;
    while ((1 << shift) < tws_sq) shift++;
    double mult = (double)(1 << shift) / tws_sq;
    int max_dist = 255 * 255;
    int almost_max = (int)(max_dist / mult + 1);
    int *wt = (int *)malloc(sizeof(int) * almost_max);
    const double h = 3.0, thr = 0.001;
    for (int a = 0; a < almost_max; a++) {
        double dist = a * mult;
        double w = exp(-dist / (h * h));
        if (isnan(w)) w = 1.0;
        int wi = (int)rint(fpm * w);
        if (wi < thr * fpm) wi = 0;
        wt[a] = wi;
    }
    /* per search offset: squared-difference image then a 7x7 box sum via integral image */
    size_t npx = (size_t)W * H;
    int64_t *est = (int64_t *)calloc(npx, sizeof(int64_t));
    int64_t *wsum = (int64_t *)calloc(npx, sizeof(int64_t));
    int DW = W + 2 * tr, DH = H + 2 * tr;
    int32_t *integ = (int32_t *)malloc(sizeof(int32_t) * (size_t)(DW + 1) * (DH + 1));
    for (int oy = -sr; oy <= sr; oy++)
        for (int ox = -sr; ox <= sr; ox++) {
            for (int x = 0; x <= DW; x++) integ[x] = 0;
            for (int y = 0; y < DH; y++) {
                int32_t rowacc = 0;
                integ[(y + 1) * (DW + 1)] = 0;
                for (int x = 0; x < DW; x++) {
                    int ey = y + br - tr, ex = x + br - tr;
                    int d = (int)ext[(ey + oy) * EW + ex + ox] - (int)ext[ey * EW + ex];
                    rowacc += d * d;
                    integ[(y + 1) * (DW + 1) + x + 1] = integ[y * (DW + 1) + x + 1] + rowacc;
                }
            }
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) {
                    int32_t dist = integ[(y + tw) * (DW + 1) + x + tw] - integ[y * (DW + 1) + x + tw] -
                                   integ[(y + tw) * (DW + 1) + x] + integ[y * (DW + 1) + x];
                    int w = wt[dist >> shift];
                    est[y * W + x] += (int64_t)w * ext[(y + br + oy) * EW + x + br + ox];
                    wsum[y * W + x] += w;
                }
        }
    for (size_t i = 0; i < npx; i++) {
        /* OpenCV divByWeightsSum: (estimation + weights_sum/2) / weights_sum, unsigned */
        uint32_t e = (uint32_t)est[i], w = (uint32_t)wsum[i];
        dst[i] = (uint8_t)((e + w / 2) / w);
    }
    free(integ); free(wsum); free(est); free(wt); free(ext);
}

/* ------------------------------------------------------------------------------------------
 * K7  WeightedBackground (piclassifier/motiondetector.py:178-248)
 *   first call (197-212): background[crop] = frame[crop]; average = np.average(frame[crop])
 *                         (unrounded double); edges replicated
 *   later calls (213-236): keep where background < frame - weight  (fp64, weight accumulated by
 *                          repeated += weight_add); else background = frame, weight = 0;
 *                          if any pixel changed: average = int(round(mean(background[crop])))
 *                          and edges are replicated again (239-244: rows first, then columns).
 * State: bg int32 [H][W] (integer valued), weight double [(H-2e)][(W-2e)].
 * `frame` is the already-truncated int32 full frame (np.int32(...) at motiondetector.py:198).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int W, H, edge;
    double weight_add;
    int initialised;
    double average;
    int32_t *bg;    /* H*W */
    double *weight; /* (H-2e)*(W-2e) */
} orc_background;

static void bg_set_edges(orc_background *s) {
    int W = s->W, H = s->H, e = s->edge;
    for (int i = 0; i < e; i++) {
        memcpy(s->bg + (size_t)i * W, s->bg + (size_t)e * W, sizeof(int32_t) * W);
        memcpy(s->bg + (size_t)(H - 1 - i) * W, s->bg + (size_t)(H - 1 - e) * W, sizeof(int32_t) * W);
        for (int y = 0; y < H; y++) {
            s->bg[y * W + i] = s->bg[y * W + e];
            s->bg[y * W + W - 1 - i] = s->bg[y * W + W - 1 - e];
        }
    }
}

orc_background *orc_background_create(int W, int H, int edge, double weight_add) {
    orc_background *s = (orc_background *)calloc(1, sizeof(orc_background));
    s->W = W; s->H = H; s->edge = edge; s->weight_add = weight_add;
    s->bg = (int32_t *)calloc((size_t)W * H, sizeof(int32_t));
    s->weight = (double *)calloc((size_t)(W - 2 * edge) * (H - 2 * edge), sizeof(double));
    return s;
}

void orc_background_destroy(orc_background *s) {
    if (!s) return;
    free(s->bg); free(s->weight); free(s);
}

void orc_background_process(orc_background *s, const int32_t *frame) {
    int W = s->W, H = s->H, e = s->edge, cw = W - 2 * e, ch = H - 2 * e;
    if (!s->initialised) {
        double sum = 0;
        for (int y = 0; y < ch; y++)
            for (int x = 0; x < cw; x++) {
                int32_t v = frame[(y + e) * W + x + e];
                s->bg[(y + e) * W + x + e] = v;
                sum += v;
            }
        s->average = sum / ((double)cw * ch);
        bg_set_edges(s);
        s->initialised = 1;
        return;
    }
    int changed = 0;
    for (int y = 0; y < ch; y++)
        for (int x = 0; x < cw; x++) {
            int32_t *b = &s->bg[(y + e) * W + x + e];
            double *w = &s->weight[y * cw + x];
            int32_t a = frame[(y + e) * W + x + e];
            volatile double rhs = (double)a - *w;
            if ((double)*b < rhs) {
                *w = *w + s->weight_add;
            } else {
                if (*b != a) changed = 1;
                *b = a;
                *w = 0;
            }
        }
    if (changed) {
        double sum = 0;
        for (int y = 0; y < ch; y++)
            for (int x = 0; x < cw; x++) sum += s->bg[(y + e) * W + x + e];
        s->average = (double)py_round(sum / ((double)cw * ch));
        bg_set_edges(s);
    }
}

void orc_background_get(const orc_background *s, int32_t *bg_out, double *weight_out, double *avg_out) {
    if (bg_out) memcpy(bg_out, s->bg, sizeof(int32_t) * (size_t)s->W * s->H);
    if (weight_out)
        memcpy(weight_out, s->weight, sizeof(double) * (size_t)(s->W - 2 * s->edge) * (s->H - 2 * s->edge));
    if (avg_out) *avg_out = s->average;
}

void orc_background_set(orc_background *s, const int32_t *bg, const double *weight, double avg) {
    memcpy(s->bg, bg, sizeof(int32_t) * (size_t)s->W * s->H);
    if (weight)
        memcpy(s->weight, weight, sizeof(double) * (size_t)(s->W - 2 * s->edge) * (s->H - 2 * s->edge));
    s->average = avg;
    s->initialised = 1;
}

/* ------------------------------------------------------------------------------------------
 * K6  ClipTracker.get_delta_frame (track/cliptracker.py:249-261) + per-region np.var over the
 *     raw component bounding box (:316-318).
 *   norm255(F) = normalize(F, new_max=255) on the fp64 filtered frame: because min/max are
 *   np.float64 scalars the arithmetic is fp64:  255 * (float32(F) - min) / (max - min);
 *   degenerate: zeros when max == min == 0, F / max when max == min != 0.
 *   delta = | float32(norm255(F_t)) - float32(norm255(F_{t-1})) |  (fp32)
 *   variance = population variance of delta over the box.  (numpy reduces in fp32 with pairwise
 *   sums; this restatement accumulates in fp64 -- tolerance class, see tests.)
 * ---------------------------------------------------------------------------------------- */
static inline float norm255_f64(double f, double mn, double mx) {
    if (mx == mn) return (mx == 0.0) ? 0.0f : (float)(f / mx);
    return (float)(255.0 * (f - mn) / (mx - mn));
}

double orc_box_variance(const float *filt, const float *prev, int W, double mn, double mx,
                        double pmn, double pmx, int left, int top, int w, int h) {
    double s = 0, s2 = 0;
    for (int y = top; y < top + h; y++)
        for (int x = left; x < left + w; x++) {
            float a = norm255_f64(filt[y * W + x], mn, mx);
            float b = norm255_f64(prev[y * W + x], pmn, pmx);
            float d = fabsf(a - b);
            s += d;
        }
    double n = (double)w * h, mean = s / n;
    for (int y = top; y < top + h; y++)
        for (int x = left; x < left + w; x++) {
            float a = norm255_f64(filt[y * W + x], mn, mx);
            float b = norm255_f64(prev[y * W + x], pmn, pmx);
            double d = (double)fabsf(a - b) - mean;
            s2 += d * d;
        }
    return s2 / n;
}

/* ------------------------------------------------------------------------------------------
 * K8  ClipStats.add_frame (track/clip.py:474-487): per frame min, max, median, mean of the
 *     thermal frame and sum |filtered|.  np.median of an even count = mean of the two middle
 *     order statistics.
 * ---------------------------------------------------------------------------------------- */
static int cmp_u16(const void *a, const void *b) {
    return (int)*(const uint16_t *)a - (int)*(const uint16_t *)b;
}

void orc_frame_stats(const uint16_t *pix, const float *filtered, int n_px, double *out5) {
    uint16_t *tmp = (uint16_t *)malloc(sizeof(uint16_t) * n_px);
    memcpy(tmp, pix, sizeof(uint16_t) * n_px);
    qsort(tmp, n_px, sizeof(uint16_t), cmp_u16);
    double sum = 0, fs = 0;
    for (int i = 0; i < n_px; i++) sum += pix[i];
    if (filtered)
        for (int i = 0; i < n_px; i++) fs += fabs((double)filtered[i]);
    out5[0] = tmp[0];
    out5[1] = tmp[n_px - 1];
    out5[2] = (n_px & 1) ? tmp[n_px / 2] : 0.5 * ((double)tmp[n_px / 2 - 1] + (double)tmp[n_px / 2]);
    out5[3] = sum / n_px;
    out5[4] = fs;
    free(tmp);
}

/* ------------------------------------------------------------------------------------------
 * Whole-clip extraction: ClipTrackExtractor.init_clip / _track_clip / process_frame
 * (track/cliptrackextractor.py:98-247), device-able part only (everything up to the
 * component table and per-component variance; matching/Kalman stay with the host shims).
 *
 *   init:   WeightedBackground.process_frame(init_frame)                          (:131-139)
 *   frame t: filtered = float32(pix) - background                                 (:212)
 *            U, thresh = _get_filtered_frame ; [NLM] ; blur/threshold/close ; CC  (:214-219)
 *            per-component variance of the delta frame (cliptracker.py:263-318)
 *            if update_background: background.process_frame(mean of last <=45 frames) (:168-176)
 * Outputs (any pointer may be NULL): see parameter names; comp rows as in orc_cc8.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t W, H, edge, background_thresh;
    double weight_add;
    int32_t denoise, update_background, max_comp, calc_stats;
} orc_params;

int orc_extract_clip(const uint16_t *frames, int n_frames, const uint16_t *init_frame,
                     const orc_params *p, float *out_filtered, uint8_t *out_labels,
                     int32_t *out_ncomp, int32_t *out_comp, double *out_var, uint8_t *out_u,
                     float *out_thresh, float *out_norm, int32_t *out_bg, double *out_avg,
                     double *out_fstats, int32_t *final_bg, double *final_weight,
                     double *final_avg) {
    int W = p->W, H = p->H, n = W * H;
    orc_background *bgs = orc_background_create(W, H, p->edge, p->weight_add);
    int32_t *tmp32 = (int32_t *)malloc(sizeof(int32_t) * n);
    for (int i = 0; i < n; i++) tmp32[i] = init_frame[i];
    orc_background_process(bgs, tmp32);
    uint32_t *ssum = (uint32_t *)calloc(n, sizeof(uint32_t));
    float *filt = (float *)malloc(sizeof(float) * n), *prev = (float *)malloc(sizeof(float) * n);
    uint8_t *u = (uint8_t *)malloc(n), *u2 = (uint8_t *)malloc(n), *bl = (uint8_t *)malloc(n),
            *mask = (uint8_t *)malloc(n);
    int32_t *labels = (int32_t *)malloc(sizeof(int32_t) * n);
    int32_t *comp = (int32_t *)malloc(sizeof(int32_t) * 8 * (p->max_comp > 0 ? p->max_comp : 1));
    double pmn = 0, pmx = 0;
    for (int t = 0; t < n_frames; t++) {
        const uint16_t *pix = frames + (size_t)t * n;
        if (out_bg) memcpy(out_bg + (size_t)t * n, bgs->bg, sizeof(int32_t) * n);
        if (out_avg) out_avg[t] = bgs->average;
        double fmn = INFINITY, fmx = -INFINITY;
        for (int i = 0; i < n; i++) {
            filt[i] = (float)((double)pix[i] - (double)bgs->bg[i]);
            if (filt[i] < fmn) fmn = filt[i];
            if (filt[i] > fmx) fmx = filt[i];
        }
        if (out_filtered) memcpy(out_filtered + (size_t)t * n, filt, sizeof(float) * n);
        float thresh, mx, mn;
        int32_t avg_change;
        orc_normalise_frame(pix, bgs->bg, bgs->average, n, p->background_thresh, u, &thresh, &mx,
                            &mn, &avg_change);
        const uint8_t *det_in = u;
        if (p->denoise) {
            orc_nlm_denoise(u, W, H, u2);
            det_in = u2;
        }
        if (out_u) memcpy(out_u + (size_t)t * n, det_in, n);
        if (out_thresh) out_thresh[t] = thresh;
        if (out_norm) { out_norm[2 * t] = mx; out_norm[2 * t + 1] = mn; }
        orc_blur5(det_in, W, H, bl);
        orc_threshold_close(bl, W, H, thresh, mask);
        int nc = orc_cc8(mask, W, H, labels, comp, p->max_comp);
        if (out_ncomp) out_ncomp[t] = nc;
        int stored = nc < p->max_comp ? nc : p->max_comp;
        if (out_comp) memcpy(out_comp + (size_t)t * p->max_comp * 8, comp, sizeof(int32_t) * 8 * stored);
        if (out_var)
            for (int c = 0; c < stored; c++)
                out_var[(size_t)t * p->max_comp + c] =
                    (t == 0) ? 0.0
                             : orc_box_variance(filt, prev, W, fmn, fmx, pmn, pmx, comp[c * 8],
                                                comp[c * 8 + 1], comp[c * 8 + 2], comp[c * 8 + 3]);
        if (out_labels)
            for (int i = 0; i < n; i++) out_labels[(size_t)t * n + i] = labels[i] > 255 ? 255 : (uint8_t)labels[i];
        if (out_fstats && p->calc_stats) orc_frame_stats(pix, filt, n, out_fstats + 5 * (size_t)t);
        /* sliding sum == np.mean of the last <=45 thermal frames (exact integers) */
        for (int i = 0; i < n; i++) ssum[i] += pix[i];
        if (t >= ORC_MEAN_FRAMES) {
            const uint16_t *old = frames + (size_t)(t - ORC_MEAN_FRAMES) * n;
            for (int i = 0; i < n; i++) ssum[i] -= old[i];
        }
        if (p->update_background) {
            int cnt = t + 1 < ORC_MEAN_FRAMES ? t + 1 : ORC_MEAN_FRAMES;
            for (int i = 0; i < n; i++) tmp32[i] = (int32_t)((double)ssum[i] / (double)cnt);
            orc_background_process(bgs, tmp32);
        }
        float *sw = prev; prev = filt; filt = sw;
        pmn = fmn; pmx = fmx;
    }
    orc_background_get(bgs, final_bg, final_weight, final_avg);
    free(comp); free(labels); free(mask); free(bl); free(u2); free(u); free(prev); free(filt);
    free(ssum); free(tmp32);
    orc_background_destroy(bgs);
    return 0;
}

/* Batch driver for the CPU baseline: clips are independent (track/trackextractor.py:80-85
 * runs them in a process pool); here one OpenMP thread per clip, regions-only outputs. */
int orc_extract_batch(const uint16_t *frames, int n_clips, int n_frames, const orc_params *params,
                      int32_t *out_ncomp, int32_t *out_comp, double *out_var, int n_threads) {
    int n = params[0].W * params[0].H;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
    for (int c = 0; c < n_clips; c++) {
        const uint16_t *clip = frames + (size_t)c * n_frames * n;
        const orc_params *p = &params[c];
        orc_extract_clip(clip, n_frames, clip, p, NULL, NULL, out_ncomp + (size_t)c * n_frames,
                         out_comp + (size_t)c * n_frames * p->max_comp * 8,
                         out_var + (size_t)c * n_frames * p->max_comp, NULL, NULL, NULL, NULL, NULL,
                         NULL, NULL, NULL, NULL);
    }
    return 0;
}
