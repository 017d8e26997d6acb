// extract_kernel.cu -- the persistent per-clip extraction kernel (see cptrack_kernels.cuh).
#include "cptrack_kernels.cuh"

namespace cpt {

namespace {

__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * n - 2 - p;
    return p;
}

__device__ __forceinline__ const uint16_t *frame_ptr(const KernelArgs &a, const cpt_clip &c, int t) {
    int64_t idx = c.ring_frames ? (int64_t)((c.first_frame + t) % c.ring_frames) : (int64_t)t;
    return a.frames + (size_t)(c.frame_offset + idx) * a.g.npx;
}

// replicate the edge_pixels border of B from the crop interior (motiondetector.py:239-244:
// rows first, then columns).  Two barriers inside.
__device__ void replicate_edges(Smem &s, const Geometry &g) {
    const int tid = threadIdx.x, W = g.W, H = g.H, e = g.edge;
    if (e > 0) {
        for (int i = tid; i < e * W; i += kThreads) {
            int r = i / W, x = i - r * W;
            s.B[r * W + x] = s.B[e * W + x];
            s.B[(H - 1 - r) * W + x] = s.B[(H - 1 - e) * W + x];
        }
        __syncthreads();
        for (int i = tid; i < e * H; i += kThreads) {
            int r = i / H, y = i - r * H;
            s.B[y * W + r] = s.B[y * W + e];
            s.B[y * W + W - 1 - r] = s.B[y * W + W - 1 - e];
        }
        __syncthreads();
    }
}

// normalize(F, new_max=255) of get_delta_frame (track/cliptracker.py:249-261), cast to fp32.
// exact_f32: 255*(max-min) < 2^24, so numerator and denominator are exact fp32 integers and the
// fp32 IEEE quotient is the correctly rounded value the reference's fp64-then-cast produces.
__device__ __forceinline__ float norm255(int f, int mn, int mx, bool exact_f32) {
    if (mx == mn) return (mx == 0) ? 0.0f : (exact_f32 ? __fdiv_rn((float)f, (float)mx) : (float)((double)f / (double)mx));
    if (exact_f32) return __fdiv_rn((float)(255 * (f - mn)), (float)(mx - mn));
    return (float)(255.0 * ((double)f - (double)mn) / ((double)mx - (double)mn));
}

// ------------------------------------------------------------------------------------------------
// K4 second half + K5: closed mask -> components.  On entry s.M holds the thresholded mask.
// On exit: s.c_* hold per-slot statistics, s.c_rank the OpenCV label order, s.parent the
// run -> slot map, and (if want_labels) s.U the uint8 label image.  Returns the component count.
// All threads of the CTA must call this.
// ------------------------------------------------------------------------------------------------
__device__ int label_components(Smem &s, const Geometry &g, bool want_labels) {
    const int tid = threadIdx.x;
    const int W = g.W;
    uint32_t c = 0;
    int y = 0, wi = 0;
    if (tid < g.words) {
        y = (int)(((uint32_t)tid * g.rw_magic) >> 13);
        wi = tid - y * g.row_words;
        uint32_t m0 = s.M[tid];
        if (y == 0) c = m0;
        else {
            uint32_t m1 = s.M[tid - g.row_words];
            uint32_t m2 = (y >= 2) ? s.M[tid - 2 * g.row_words] : 0u;
            c = m1 | (m0 & m2);
        }
        s.C[tid] = c;
    }
    if (tid < kCompSlots) {
        s.c_key[tid] = INT32_MAX; s.c_area[tid] = 0; s.c_sx[tid] = 0; s.c_sy[tid] = 0;
        s.c_l[tid] = INT32_MAX; s.c_t[tid] = INT32_MAX; s.c_r[tid] = -1; s.c_b[tid] = -1;
        s.acc_s[tid] = 0.0; s.acc_s2[tid] = 0.0;
    }
    if (tid == 0) s.ncomp = 0;
    if (tid < kMaxH) { s.need_u[tid] = 0; s.need_b[tid] = 0; }
    if (want_labels) {
        uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < g.npx / 16; i += kThreads) reinterpret_cast<uint4 *>(s.U)[i] = z;
    }
    if (!__syncthreads_or(c != 0)) return 0;

    // ---- run starts and ids
    uint32_t stw = 0;
    int base = 0;
    if (tid < g.words) {
        uint32_t carry = 0;
        for (int q = 0; q < wi; ++q) {
            uint32_t cq = s.C[tid - wi + q];
            base += __popc(cq & ~((cq << 1) | carry));
            carry = cq >> 31;
        }
        stw = c & ~((c << 1) | carry);
        s.ST[tid] = stw;
        s.base[tid] = (uint8_t)base;
        uint32_t bitsleft = stw;
        int n = 0;
        while (bitsleft) {
            bitsleft &= bitsleft - 1;
            int id = y * kRunsPerRow + base + n;
            s.parent[id] = (uint16_t)id;
            ++n;
        }
    }
    __syncthreads();
    // ---- unions with the row above (8-connectivity)
    if (tid < g.words && y > 0 && c != 0) {
        const int up = tid - g.row_words;
        uint32_t u = s.C[up];
        uint32_t u_l = (wi > 0) ? (s.C[up - 1] >> 31) : 0u;
        uint32_t u_r = (wi + 1 < g.row_words) ? (s.C[up + 1] & 1u) : 0u;
        uint32_t c_l = (wi > 0) ? (s.C[tid - 1] >> 31) : 0u;
        uint32_t c_r = (wi + 1 < g.row_words) ? (s.C[tid + 1] & 1u) : 0u;
        uint32_t ul = (u << 1) | u_l, ur = (u >> 1) | (u_r << 31);
        uint32_t cl = (c << 1) | c_l, cr = (c >> 1) | (c_r << 31);
        uint32_t needA = c & u & ~(cl & ul);   // pixel above, unless the left neighbour already links to it
        uint32_t needB = c & ul & ~u & ~cl;    // upper-left only
        uint32_t needC = c & ur & ~u & ~cr;    // upper-right only
        const int xb = wi * 32;
        while (needA) {
            int b = __ffs(needA) - 1;
            needA &= needA - 1;
            uf_union(s.parent, run_id(s, g, xb + b, y), run_id(s, g, xb + b, y - 1));
        }
        while (needB) {
            int b = __ffs(needB) - 1;
            needB &= needB - 1;
            uf_union(s.parent, run_id(s, g, xb + b, y), run_id(s, g, xb + b - 1, y - 1));
        }
        while (needC) {
            int b = __ffs(needC) - 1;
            needC &= needC - 1;
            uf_union(s.parent, run_id(s, g, xb + b, y), run_id(s, g, xb + b + 1, y - 1));
        }
    }
    __syncthreads();
    // ---- roots -> component slots
    if (stw) {
        uint32_t bitsleft = stw;
        int n = 0;
        while (bitsleft) {
            bitsleft &= bitsleft - 1;
            int id = y * kRunsPerRow + base + n;
            if (s.parent[id] == id) {
                int slot = atomicAdd(&s.ncomp, 1);
                s.parent[id] = (uint16_t)(kSlotFlag | (slot < CPT_MAX_COMPONENTS ? slot : CPT_MAX_COMPONENTS));
            }
            ++n;
        }
    }
    __syncthreads();
    const int ncomp = s.ncomp;
    // ---- per-run statistics into the slot tables
    if (stw) {
        uint32_t bitsleft = stw;
        int n = 0;
        while (bitsleft) {
            int b = __ffs(bitsleft) - 1;
            bitsleft &= bitsleft - 1;
            int id = y * kRunsPerRow + base + n;
            ++n;
            int slot = uf_slot(s.parent, id);
            s.parent[id] = (uint16_t)(kSlotFlag | slot);
            int xs = wi * 32 + b;
            uint32_t inv = ~(c >> b);
            int len = (inv == 0) ? 32 : (__ffs(inv) - 1);
            if (b + len >= 32) {  // run continues into the following words
                len = 32 - b;
                for (int q = wi + 1; q < g.row_words; ++q) {
                    uint32_t cn = ~s.C[tid - wi + q];
                    if (cn == 0) { len += 32; continue; }
                    len += __ffs(cn) - 1;
                    break;
                }
            }
            atomicMin(&s.c_key[slot], (y >> 1) * g.block_w + (xs >> 1));
            atomicAdd(&s.c_area[slot], len);
            atomicAdd(&s.c_sx[slot], len * (2 * xs + len - 1) / 2);
            atomicAdd(&s.c_sy[slot], len * y);
            atomicMin(&s.c_l[slot], xs);
            atomicMax(&s.c_r[slot], xs + len - 1);
            atomicMin(&s.c_t[slot], y);
            atomicMax(&s.c_b[slot], y);
        }
    }
    __syncthreads();
    // ---- OpenCV label order: rank by the key of the component's first 2x2 block
    const int nslots = min(ncomp, CPT_MAX_COMPONENTS);
    if (tid < nslots) {
        int key = s.c_key[tid], rank = 0;
        for (int q = 0; q < nslots; ++q) rank += (s.c_key[q] < key);
        s.c_rank[tid] = (uint8_t)rank;
    }
    __syncthreads();
    // ---- label image
    if (want_labels && stw) {
        uint32_t bitsleft = stw;
        int n = 0;
        while (bitsleft) {
            int b = __ffs(bitsleft) - 1;
            bitsleft &= bitsleft - 1;
            int id = y * kRunsPerRow + base + n;
            ++n;
            int slot = s.parent[id] & 0xff;
            uint8_t lab = (slot < CPT_MAX_COMPONENTS) ? (uint8_t)(s.c_rank[slot] + 1) : (uint8_t)255;
            uint8_t *row = s.U + y * W;
            int x = wi * 32 + b;
            while (x < W && ((s.C[y * g.row_words + (x >> 5)] >> (x & 31)) & 1u)) row[x++] = lab;
        }
    }
    return ncomp;
}

// ------------------------------------------------------------------------------------------------
// K4 first half: 5x5 binomial blur of s.U (fixed point, one rounding, BORDER_REFLECT_101) and
// threshold `> ith`, 8 pixels per thread in packed 16-bit lanes -> bit rows in s.M.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void blur_threshold(Smem &s, const Geometry &g, int ith) {
    const int tid = threadIdx.x, W = g.W, H = g.H;
    uint8_t *M8 = reinterpret_cast<uint8_t *>(s.M);
    const uint32_t T = (ith >= 0 && ith < 255) ? (uint32_t)(((ith + 1) << 8) - 128) : 0u;
    const bool fast_rows = H >= 4;
#pragma unroll 1
    for (int j = 0; j < 3; ++j) {
        int grp = tid + j * kThreads;
        if (grp >= g.groups) break;
        int y = (int)(((uint32_t)grp * g.gpr_magic) >> 17), gx = grp - y * g.gpr, x0 = gx * 8;
        uint32_t bits = 0;
        if (ith < 0) {
            bits = 0xffu;
        } else if (ith < 255 && ((s.need_b[y] >> gx) & 1u)) {
            // (groups outside need_b cannot exceed the threshold: every input of their window is <= ith)
            uint32_t V[6] = {0, 0, 0, 0, 0, 0};
            const bool left_edge = (x0 == 0), right_edge = (x0 + 8 == W);
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                const uint32_t wgt = (r == 0 || r == 4) ? 1u : ((r == 2) ? 6u : 4u);
                int yy = y + r - 2;
                if (fast_rows) {
                    yy = (yy < 0) ? -yy : yy;
                    yy = (yy >= H) ? 2 * H - 2 - yy : yy;
                } else {
                    yy = reflect101(yy, H);
                }
                const uint8_t *row = s.U + yy * W + x0;
                uint2 m = *reinterpret_cast<const uint2 *>(row);
                uint32_t lw = left_edge ? 0u : *reinterpret_cast<const uint32_t *>(row - 4);
                uint32_t rw = right_edge ? 0u : *reinterpret_cast<const uint32_t *>(row + 8);
                uint32_t lp = left_edge ? __byte_perm(m.x, 0, 0x4142) : __byte_perm(lw, 0, 0x4342);
                uint32_t rp = right_edge ? __byte_perm(m.y, 0, 0x4142) : __byte_perm(rw, 0, 0x4140);
                V[0] += wgt * lp;
                V[1] += wgt * __byte_perm(m.x, 0, 0x4140);
                V[2] += wgt * __byte_perm(m.x, 0, 0x4342);
                V[3] += wgt * __byte_perm(m.y, 0, 0x4140);
                V[4] += wgt * __byte_perm(m.y, 0, 0x4342);
                V[5] += wgt * rp;
            }
            uint32_t odd[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) odd[q] = __byte_perm(V[q], V[q + 1], 0x5432);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t sum = V[q] + V[q + 2] + 6u * V[q + 1] + 4u * (odd[q] + odd[q + 1]);
                bits |= ((sum & 0xffffu) >= T ? 1u : 0u) << (2 * q);
                bits |= ((sum >> 16) >= T ? 1u : 0u) << (2 * q + 1);
            }
        }
        M8[y * g.row_words * 4 + gx] = (uint8_t)bits;
    }
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1) extract_clips_kernel(const KernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &s = *reinterpret_cast<Smem *>(smem_raw);
    const Geometry &g = a.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = g.W, npx = g.npx;

    // clips are handed out dynamically so that ragged batches stay balanced
    while (true) {
        __syncthreads();
        if (tid == 0) s.bcast_i[15] = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int ci = s.bcast_i[15];
        if (ci >= a.n_clips) break;
        const cpt_clip clip = a.clips[ci];
        const WeightTable wt = a.tables[clip.weight_table & 3];
        uint8_t *st_raw = a.state ? a.state + (size_t)ci * state_bytes(npx) : nullptr;
        StateHeader *st_hdr = reinterpret_cast<StateHeader *>(st_raw);
        uint16_t *st_B = reinterpret_cast<uint16_t *>(st_raw + sizeof(StateHeader));
        uint16_t *st_K = st_B + npx;
        uint32_t *st_S = reinterpret_cast<uint32_t *>(st_K + npx);
        float *st_F = reinterpret_cast<float *>(st_S + npx);
        float *scratch = a.scratch ? a.scratch + (size_t)blockIdx.x * 2 * npx : nullptr;
        const bool want_stats = clip.flags & CPT_CLIP_FRAME_STATS;

        double average;
        int prev_fmin = 0, prev_fmax = 0, have_prev = 0;
        int frames_seen = 0;

        // ---------------------------------------------------------------- init / resume
        for (int i = tid; i < kMaxWords; i += kThreads) s.M[i] = 0;
        for (int i = tid; i < kSmemWeights; i += kThreads) s.wthr[i] = (i <= wt.max_count) ? __ldg(wt.thr + i) : 65536u;
        if (tid < kMaxH) { s.need_u[tid] = 0; s.need_b[tid] = 0; }
        if (clip.flags & CPT_CLIP_RESUME) {
            for (int i = tid; i < npx; i += kThreads) {
                s.B[i] = st_B[i];
                s.K[i] = st_K[i];
                s.S[i] = st_S[i];
            }
            average = st_hdr->average;
            prev_fmin = st_hdr->prev_fmin;
            prev_fmax = st_hdr->prev_fmax;
            have_prev = st_hdr->have_prev;
            frames_seen = st_hdr->frames_seen;
            __syncthreads();
        } else {
            // WeightedBackground first call: motiondetector.py:199-212
            const uint16_t *init = a.frames + (size_t)clip.init_offset * npx;
            uint32_t csum = 0;
            for (int i = tid; i < npx; i += kThreads) {
                int y = i / W, x = i - y * W;
                uint16_t v = __ldg(init + i);
                s.B[i] = v;
                s.K[i] = 0;
                s.S[i] = 0;
                if (x >= g.edge && x < W - g.edge && y >= g.edge && y < g.H - g.edge) csum += v;
            }
            csum = __reduce_add_sync(0xffffffffu, csum);
            if (lane == 0) s.red_u[warp] = csum;
            __syncthreads();
            if (warp == 0) {
                uint32_t v = s.red_u[lane];
                v = __reduce_add_sync(0xffffffffu, v);
                if (lane == 0) s.bcast_d[0] = (double)v / (double)g.ncrop;
            }
            __syncthreads();
            average = s.bcast_d[0];
            replicate_edges(s, g);
        }

        for (int t = 0; t < clip.n_frames; ++t) {
            const int t_abs = clip.first_frame + t;
            const size_t o = (size_t)(clip.out_offset + t);
            const uint16_t *P = frame_ptr(a, clip, t);
            const uint16_t *Pold = (t_abs >= kMeanFrames) ? frame_ptr(a, clip, t - kMeanFrames) : nullptr;
            float *fcur = a.filtered ? a.filtered + o * npx : scratch + (size_t)(t & 1) * npx;
            const float *fprev = (t == 0) ? st_F : (a.filtered ? a.filtered + (o - 1) * npx : scratch + (size_t)((t - 1) & 1) * npx);

            // ------------------------------------------------------------ sweep 1 (K1, K7 sum, K8)
            uint4 pv[3];
            int gmaxf[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
            uint32_t psum = 0, fabs_sum = 0;
            int fmin = INT32_MAX, fmax = INT32_MIN, pmin = INT32_MAX, pmax = INT32_MIN;
            const bool more = (t + 1 < clip.n_frames);
            const uint16_t *Pnext = more ? frame_ptr(a, clip, t + 1) : nullptr;
            const uint16_t *Pold_next = (more && t_abs + 1 >= kMeanFrames) ? frame_ptr(a, clip, t + 1 - kMeanFrames) : nullptr;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                int grp = tid + j * kThreads;
                if (grp < g.groups) {
                    pv[j] = ldg16(P + grp * 8);
                    if ((grp & 7) == 0) {  // one 128-byte line per 8 groups: pull the next frame into L2
                        if (Pnext) asm volatile("prefetch.global.L2 [%0];" ::"l"(Pnext + grp * 8));
                        if (Pold_next) asm volatile("prefetch.global.L2 [%0];" ::"l"(Pold_next + grp * 8));
                    }
                    uint4 qv = make_uint4(0, 0, 0, 0);
                    if (Pold) qv = ldg16(Pold + grp * 8);
                    int p[8], b[8], q[8];
                    unpack8(pv[j], p);
                    unpack8(qv, q);
                    unpack8(*reinterpret_cast<const uint4 *>(s.B + grp * 8), b);
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        int d = p[i] - b[i];
                        psum += p[i];
                        fmin = min(fmin, d);
                        gmaxf[j] = max(gmaxf[j], d);
                        f[i] = (float)d;
                    }
                    fmax = max(fmax, gmaxf[j]);
                    if (want_stats) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            pmin = min(pmin, p[i]);
                            pmax = max(pmax, p[i]);
                            fabs_sum += abs(p[i] - b[i]);
                        }
                    }
                    float4 *dst = reinterpret_cast<float4 *>(fcur + grp * 8);
                    dst[0] = make_float4(f[0], f[1], f[2], f[3]);
                    dst[1] = make_float4(f[4], f[5], f[6], f[7]);
                    uint4 *sp = reinterpret_cast<uint4 *>(s.S + grp * 8);
                    uint4 s0 = sp[0], s1 = sp[1];
                    s0.x += p[0] - q[0]; s0.y += p[1] - q[1]; s0.z += p[2] - q[2]; s0.w += p[3] - q[3];
                    s1.x += p[4] - q[4]; s1.y += p[5] - q[5]; s1.z += p[6] - q[6]; s1.w += p[7] - q[7];
                    sp[0] = s0;
                    sp[1] = s1;
                }
            }
            psum = __reduce_add_sync(0xffffffffu, psum);
            fmin = __reduce_min_sync(0xffffffffu, fmin);
            fmax = __reduce_max_sync(0xffffffffu, fmax);
            if (want_stats) {
                pmin = __reduce_min_sync(0xffffffffu, pmin);
                pmax = __reduce_max_sync(0xffffffffu, pmax);
                fabs_sum = __reduce_add_sync(0xffffffffu, fabs_sum);
            }
            if (lane == 0) {
                s.red_u[warp * 6 + 0] = psum;
                s.red_u[warp * 6 + 1] = (uint32_t)fmin;
                s.red_u[warp * 6 + 2] = (uint32_t)fmax;
                s.red_u[warp * 6 + 3] = (uint32_t)pmin;
                s.red_u[warp * 6 + 4] = (uint32_t)pmax;
                s.red_u[warp * 6 + 5] = fabs_sum;
            }
            __syncthreads();
            // ------------------------------------------------------------ scalars (K2)
            if (warp == 0) {
                uint32_t v0 = __reduce_add_sync(0xffffffffu, s.red_u[lane * 6 + 0]);
                int v1 = __reduce_min_sync(0xffffffffu, (int)s.red_u[lane * 6 + 1]);
                int v2 = __reduce_max_sync(0xffffffffu, (int)s.red_u[lane * 6 + 2]);
                int v3 = __reduce_min_sync(0xffffffffu, (int)s.red_u[lane * 6 + 3]);
                int v4 = __reduce_max_sync(0xffffffffu, (int)s.red_u[lane * 6 + 4]);
                uint32_t v5 = __reduce_add_sync(0xffffffffu, s.red_u[lane * 6 + 5]);
                if (lane == 0) {
                    // avg_change = int(round(np.average(thermal) - background average)), cliptracker.py:103-105
                    double mean = (double)v0 / (double)npx;
                    int ac = (int)rint(mean - average);
                    int gmx = max(v2 - ac, 0), gmn = max(v1 - ac, 0);
                    float thr;
                    uint32_t magic = 0;
                    int shift = 0;
                    if (gmx == gmn) {
                        thr = (float)clip.background_thresh;  // cliptracker.py:118-119
                    } else {
                        float range = (float)gmx - (float)gmn;
                        thr = __fmul_rn(__fdiv_rn((float)clip.background_thresh, range), 255.0f);
                        unsigned r = (unsigned)(gmx - gmn);
                        if (255ull * r < (1ull << 24)) {
                            int l = 32 - __clz(r - 1);  // ceil(log2 r), r >= 1
                            if (r == 1) l = 0;
                            shift = 24 + l;
                            magic = (uint32_t)(((1ull << shift) + r - 1) / r);
                        }
                    }
                    s.bcast_i[0] = ac; s.bcast_i[1] = gmn; s.bcast_i[2] = gmx;
                    s.bcast_i[3] = v1; s.bcast_i[4] = v2;
                    s.bcast_i[5] = __float_as_int(thr);
                    s.bcast_i[6] = (int)magic; s.bcast_i[7] = shift;
                    // a group can only produce foreground if one of its pixels has U > floor(thr):
                    // U = floor(255 v / r) >= ith + 1  <=>  v >= ceil((ith + 1) r / 255), v = max(F - ac, 0) - gmn,
                    // i.e. F >= fth (vth >= 1 so the clamp never matters).  INT32_MIN: every group is hot.
                    int fth = INT32_MIN;
                    if (magic) {
                        int it = (int)floorf(thr);
                        if (it >= 0 && it < 255) {
                            long long r = gmx - gmn;
                            fth = (int)(((it + 1) * r + 254) / 255) + ac + gmn;
                        }
                    }
                    s.bcast_i[8] = fth;
                    cpt_frame_info fi;
                    fi.threshold = thr; fi.norm_min = gmn; fi.norm_max = gmx; fi.avg_change = ac;
                    fi.filtered_min = v1; fi.filtered_max = v2; fi.n_components = 0;
                    fi.thermal_min = v3; fi.thermal_max = v4; fi.thermal_sum = v0;
                    fi.abs_filtered_sum = v5; fi.thermal_median = 0.f;
                    fi.background_average = average; fi.reserved[0] = 0; fi.reserved[1] = 0;
                    a.info[o] = fi;
                }
            }
            __syncthreads();
            const int ac = s.bcast_i[0], gmn = s.bcast_i[1], gmx = s.bcast_i[2];
            const int cur_fmin = s.bcast_i[3], cur_fmax = s.bcast_i[4];
            const float thr = __int_as_float(s.bcast_i[5]);
            const uint32_t umagic = (uint32_t)s.bcast_i[6];
            const int ushift = s.bcast_i[7];
            const int ith = (int)floorf(thr);

            // ------------------------------------------------------------ sweep 2a: hot groups
            // Blur weights sum to 256, so an output can exceed ith only if some input of its 5x5 window
            // does.  Hot groups mark the outputs that may fire (rows +-2, neighbouring groups) and the
            // inputs those outputs read (rows +-4, groups +-2); everything else skips K2/K4 arithmetic.
            const int fth = s.bcast_i[8];
            const bool degenerate = (gmx == gmn);
            const bool no_fg = ith >= 255;               // also covers degenerate frames when bt >= 1... see below
            const bool all_hot = (fth == INT32_MIN);
            if (!all_hot) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    int grp = tid + j * kThreads;
                    if (grp < g.groups && gmaxf[j] >= fth) {
                        int y = (int)(((uint32_t)grp * g.gpr_magic) >> 17), gx = grp - y * g.gpr;
                        uint32_t mu = (0x1fu << gx) >> 2, mb = (0x7u << gx) >> 1;
                        for (int yy = max(y - 4, 0); yy <= min(y + 4, g.H - 1); ++yy) atomicOr(&s.need_u[yy], mu);
                        for (int yy = max(y - 2, 0); yy <= min(y + 2, g.H - 1); ++yy) atomicOr(&s.need_b[yy], mb);
                    }
                }
            } else if (tid < g.H) {
                s.need_u[tid] = 0xffffffffu;
                s.need_b[tid] = 0xffffffffu;
            }
            __syncthreads();
            // ------------------------------------------------------------ sweep 2b: U (K2)
            if (!no_fg) {
                const float range_f = (float)gmx - (float)gmn;
                const uint32_t degen_val = (gmx == 0) ? 0u : 1u;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    int grp = tid + j * kThreads;
                    if (grp < g.groups) {
                        int y = (int)(((uint32_t)grp * g.gpr_magic) >> 17), gx = grp - y * g.gpr;
                        if (!((s.need_u[y] >> gx) & 1u)) continue;
                        int p[8], b[8];
                        unpack8(pv[j], p);
                        unpack8(*reinterpret_cast<const uint4 *>(s.B + grp * 8), b);
                        uint32_t u[8];
                        if (degenerate) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) u[i] = degen_val;
                        } else if (umagic) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) u[i] = norm_u8_int(max(p[i] - b[i] - ac, 0) - gmn, umagic, ushift);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) u[i] = norm_u8(max(p[i] - b[i] - ac, 0) - gmn, range_f);
                        }
                        uint2 w;
                        w.x = u[0] | (u[1] << 8) | (u[2] << 16) | (u[3] << 24);
                        w.y = u[4] | (u[5] << 8) | (u[6] << 16) | (u[7] << 24);
                        *reinterpret_cast<uint2 *>(s.U + grp * 8) = w;
                    }
                }
            }
            __syncthreads();

            // ------------------------------------------------------------ blur + threshold (K4)
            blur_threshold(s, g, ith);
            __syncthreads();

            // ------------------------------------------------------------ close + components (K4, K5)
            const int ncomp = label_components(s, g, a.labels != nullptr);
            const int nslots = min(ncomp, CPT_MAX_COMPONENTS);
            const int nout = min(nslots, g.max_regions);

            // ------------------------------------------------------------ delta-frame variance (K6)
            if (nslots > 0 && have_prev) {
                const bool exact = 255ll * max(cur_fmax - cur_fmin, prev_fmax - prev_fmin) < (1ll << 24) &&
                                   max(max(abs(cur_fmax), abs(cur_fmin)), max(abs(prev_fmax), abs(prev_fmin))) < (1 << 24);
                for (int slot = 0; slot < nslots; ++slot) {
                    if (s.c_rank[slot] >= nout) continue;
                    const int l = s.c_l[slot], tp = s.c_t[slot];
                    const int bw = s.c_r[slot] - l + 1, bh = s.c_b[slot] - tp + 1;
                    if (warp >= bh) continue;
                    double s1 = 0.0, s2 = 0.0;
                    for (int yy = tp + warp; yy < tp + bh; yy += kWarps)
                        for (int xx = l + lane; xx < l + bw; xx += 32) {
                            int fc = (int)fcur[yy * W + xx], fp = (int)fprev[yy * W + xx];
                            float d = fabsf(norm255(fc, cur_fmin, cur_fmax, exact) - norm255(fp, prev_fmin, prev_fmax, exact));
                            s1 += (double)d;
                            s2 += (double)d * (double)d;
                        }
                    for (int off = 16; off; off >>= 1) {
                        s1 += __shfl_xor_sync(0xffffffffu, s1, off);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, off);
                    }
                    if (lane == 0) {
                        atomicAdd(&s.acc_s[slot], s1);
                        atomicAdd(&s.acc_s2[slot], s2);
                    }
                }
            }
            if (ncomp > 0) __syncthreads();
            if (tid < nslots) {
                const int rank = s.c_rank[tid];
                if (rank < nout) {
                    cpt_region r;
                    r.x = s.c_l[tid]; r.y = s.c_t[tid];
                    r.width = s.c_r[tid] - r.x + 1; r.height = s.c_b[tid] - r.y + 1;
                    r.area = s.c_area[tid]; r.sum_x = s.c_sx[tid]; r.sum_y = s.c_sy[tid];
                    r.key = s.c_key[tid];
                    double n = (double)r.width * (double)r.height;
                    double mean = s.acc_s[tid] / n;
                    double var = s.acc_s2[tid] / n - mean * mean;
                    r.pixel_variance = (have_prev && var > 0.0) ? var : 0.0;
                    a.regions[o * g.max_regions + rank] = r;
                }
            }
            if (tid == 0 && ncomp > 0) a.info[o].n_components = ncomp;
            // ------------------------------------------------------------ label image out
            if (a.labels) {
                uint4 *dst = reinterpret_cast<uint4 *>(a.labels + o * npx);
                for (int i = tid; i < npx / 16; i += kThreads) dst[i] = reinterpret_cast<const uint4 *>(s.U)[i];
            }

            // ------------------------------------------------------------ sweep 3: background (K7)
            if (clip.flags & CPT_CLIP_UPDATE_BACKGROUND) {
                const uint32_t cnt = (uint32_t)min(t_abs + 1, kMeanFrames);
                // A = floor(S / cnt) == umulhi(2S, 2^31/cnt + 1): exact for S < 2^22, cnt <= 45
                const uint32_t magic = (0x80000000u / cnt) + 1u;
                const bool table_in_smem = (t_abs + 1 < kSmemWeights);  // k never exceeds the frames seen
                uint32_t bsum = 0;
                int changed = 0;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    int grp = tid + j * kThreads;
                    if (grp < g.groups) {
                        int yy = (int)(((uint32_t)grp * g.gpr_magic) >> 17), x0 = (grp - yy * g.gpr) * 8;
                        if (yy >= g.edge && yy < g.H - g.edge) {
                            const int lo = max(g.edge - x0, 0), hi = min(W - g.edge - x0, 8);
                            const uint32_t inc = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
                            int b[8], k[8];
                            unpack8(*reinterpret_cast<const uint4 *>(s.B + grp * 8), b);
                            unpack8(*reinterpret_cast<const uint4 *>(s.K + grp * 8), k);
                            const uint4 *sp = reinterpret_cast<const uint4 *>(s.S + grp * 8);
                            uint4 s0 = sp[0], s1 = sp[1];
                            uint32_t sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int A = (int)__umulhi(sv[i] << 1, magic);
                                const int d = A - b[i];
                                const int kk = k[i];
                                const uint32_t e = table_in_smem ? s.wthr[kk] : __ldg(wt.thr + kk);
                                const int thr_d = (int)(e & kThrMask);
                                const int bound = (int)((1u << (e >> 17)) >> 1);
                                const bool keep = (d >= thr_d) || (d == thr_d - 1 && b[i] < bound);
                                const bool on = (inc >> i) & 1u;
                                const int nb = (on && !keep) ? A : b[i];
                                changed |= nb ^ b[i];
                                k[i] = on ? (keep ? min(kk + 1, wt.max_count) : 0) : kk;
                                b[i] = nb;
                                bsum += on ? (uint32_t)nb : 0u;
                            }
                            *reinterpret_cast<uint4 *>(s.B + grp * 8) = pack8(b);
                            *reinterpret_cast<uint4 *>(s.K + grp * 8) = pack8(k);
                        }
                    }
                }
                bsum = __reduce_add_sync(0xffffffffu, bsum);
                if (lane == 0) s.red_u[warp * 6] = bsum;
                int any_changed = __syncthreads_or(changed);
                if (any_changed) {
                    if (warp == 0) {
                        uint32_t v = __reduce_add_sync(0xffffffffu, s.red_u[lane * 6]);
                        if (lane == 0) s.bcast_d[0] = rint((double)v / (double)g.ncrop);
                    }
                    __syncthreads();
                    average = s.bcast_d[0];
                    replicate_edges(s, g);
                }
            }
            prev_fmin = cur_fmin;
            prev_fmax = cur_fmax;
            have_prev = 1;
            ++frames_seen;
            __syncthreads();
        }

        // ---------------------------------------------------------------- save state
        if (st_raw) {
            for (int i = tid; i < npx; i += kThreads) {
                st_B[i] = s.B[i];
                st_K[i] = s.K[i];
                st_S[i] = s.S[i];
            }
            if (clip.n_frames > 0) {
                int t = clip.n_frames - 1;
                const float *flast = a.filtered ? a.filtered + (size_t)(clip.out_offset + t) * npx : scratch + (size_t)(t & 1) * npx;
                for (int i = tid; i < npx; i += kThreads) st_F[i] = flast[i];
            }
            if (tid == 0) {
                st_hdr->average = average;
                st_hdr->frames_seen = frames_seen;
                st_hdr->initialised = 1;
                st_hdr->prev_fmin = prev_fmin;
                st_hdr->prev_fmax = prev_fmax;
                st_hdr->have_prev = have_prev;
            }
        }
    }
}

}  // namespace cpt
