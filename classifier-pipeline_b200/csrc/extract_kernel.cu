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

        double average;
        int prev_fmin = 0, prev_fmax = 0, have_prev = 0;
        int frames_seen = 0;

        // ---------------------------------------------------------------- init / resume
        for (int i = tid; i < kMaxWords; i += kThreads) s.M[i] = 0;
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
            const bool want_stats = clip.flags & CPT_CLIP_FRAME_STATS;

            // ------------------------------------------------------------ sweep 1 (K1, K7 sum, K8)
            uint4 pv[3];
            uint32_t psum = 0, fabs_sum = 0;
            int fmin = INT32_MAX, fmax = INT32_MIN, pmin = INT32_MAX, pmax = INT32_MIN;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                int grp = tid + j * kThreads;
                if (grp < g.groups) {
                    pv[j] = ldg16(P + grp * 8);
                    int p[8], b[8];
                    unpack8(pv[j], p);
                    unpack8(*reinterpret_cast<const uint4 *>(s.B + grp * 8), b);
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        int d = p[i] - b[i];
                        psum += p[i];
                        fmin = min(fmin, d);
                        fmax = max(fmax, d);
                        pmin = min(pmin, p[i]);
                        pmax = max(pmax, p[i]);
                        fabs_sum += abs(d);
                        f[i] = (float)d;
                    }
                    float4 *dst = reinterpret_cast<float4 *>(fcur + grp * 8);
                    dst[0] = make_float4(f[0], f[1], f[2], f[3]);
                    dst[1] = make_float4(f[4], f[5], f[6], f[7]);
                    uint4 *sp = reinterpret_cast<uint4 *>(s.S + grp * 8);
                    uint4 s0 = sp[0], s1 = sp[1];
                    s0.x += p[0]; s0.y += p[1]; s0.z += p[2]; s0.w += p[3];
                    s1.x += p[4]; s1.y += p[5]; s1.z += p[6]; s1.w += p[7];
                    if (Pold) {
                        int q[8];
                        unpack8(ldg16(Pold + grp * 8), q);
                        s0.x -= q[0]; s0.y -= q[1]; s0.z -= q[2]; s0.w -= q[3];
                        s1.x -= q[4]; s1.y -= q[5]; s1.z -= q[6]; s1.w -= q[7];
                    }
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
                    if (gmx == gmn) {
                        thr = (float)clip.background_thresh;  // cliptracker.py:118-119
                    } else {
                        float range = (float)gmx - (float)gmn;
                        thr = __fmul_rn(__fdiv_rn((float)clip.background_thresh, range), 255.0f);
                    }
                    s.bcast_i[0] = ac; s.bcast_i[1] = gmn; s.bcast_i[2] = gmx;
                    s.bcast_i[3] = v1; s.bcast_i[4] = v2;
                    s.bcast_i[5] = __float_as_int(thr);
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
            const int ith = (int)floorf(thr);

            // ------------------------------------------------------------ sweep 2: U (K2)
            {
                const float range_f = (float)gmx - (float)gmn;
                const bool degenerate = (gmx == gmn);
                const uint32_t degen_val = (gmx == 0) ? 0u : 1u;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    int grp = tid + j * kThreads;
                    if (grp < g.groups) {
                        int p[8], b[8];
                        unpack8(pv[j], p);
                        unpack8(*reinterpret_cast<const uint4 *>(s.B + grp * 8), b);
                        uint32_t u[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            int gg = max(p[i] - b[i] - ac, 0);
                            u[i] = degenerate ? degen_val : norm_u8(gg - gmn, range_f);
                        }
                        uint2 w;
                        w.x = u[0] | (u[1] << 8) | (u[2] << 16) | (u[3] << 24);
                        w.y = u[4] | (u[5] << 8) | (u[6] << 16) | (u[7] << 24);
                        *reinterpret_cast<uint2 *>(s.U + grp * 8) = w;
                    }
                }
            }
            __syncthreads();

            // ------------------------------------------------------------ blur 5x5 + threshold (K4)
            {
                uint8_t *M8 = reinterpret_cast<uint8_t *>(s.M);
                const uint32_t T = (ith >= 0 && ith < 255) ? (uint32_t)(((ith + 1) << 8) - 128) : 0u;
#pragma unroll 1
                for (int j = 0; j < 3; ++j) {
                    int grp = tid + j * kThreads;
                    if (grp >= g.groups) break;
                    int y = grp / g.gpr, gx = grp - y * g.gpr, x0 = gx * 8;
                    uint32_t bits = 0;
                    if (ith < 0) {
                        bits = 0xffu;
                    } else if (ith < 255) {
                        uint32_t V[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
                        for (int r = 0; r < 5; ++r) {
                            const int wgt = (r == 0 || r == 4) ? 1 : ((r == 2) ? 6 : 4);
                            const uint8_t *row = s.U + reflect101(y + r - 2, g.H) * W + x0;
                            uint2 m = *reinterpret_cast<const uint2 *>(row);
                            uint32_t lp, rp;
                            if (x0 == 0) lp = __byte_perm(m.x, 0, 0x4142);
                            else lp = __byte_perm(*reinterpret_cast<const uint32_t *>(row - 4), 0, 0x4342);
                            if (x0 + 8 == W) rp = __byte_perm(m.y, 0, 0x4142);
                            else rp = __byte_perm(*reinterpret_cast<const uint32_t *>(row + 8), 0, 0x4140);
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
            __syncthreads();

            // ------------------------------------------------------------ close (K4) + clear label image
            int any_fg;
            {
                uint32_t c = 0;
                if (tid < g.words) {
                    int y = tid / g.row_words;
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
                }
                if (tid == 0) s.ncomp = 0;
                if (a.labels) {
                    uint4 z = make_uint4(0, 0, 0, 0);
                    for (int i = tid; i < npx / 16; i += kThreads) reinterpret_cast<uint4 *>(s.U)[i] = z;
                }
                any_fg = __syncthreads_or(c != 0);
            }

            int ncomp = 0;
            if (any_fg) {
                // ---------------------------------------------------- run starts + ids (K5)
                uint32_t c = 0, stw = 0;
                int y = 0, wi = 0;
                if (tid < g.words) {
                    y = tid / g.row_words;
                    wi = tid - y * g.row_words;
                    c = s.C[tid];
                    int base = 0;
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
                // ---------------------------------------------------- unions with the row above
                if (tid < g.words && y > 0 && c != 0) {
                    const int up = tid - g.row_words;
                    uint32_t u = s.C[up];
                    uint32_t u_l = (wi > 0) ? (s.C[up - 1] >> 31) : 0u;
                    uint32_t u_r = (wi + 1 < g.row_words) ? (s.C[up + 1] & 1u) : 0u;
                    uint32_t c_l = (wi > 0) ? (s.C[tid - 1] >> 31) : 0u;
                    uint32_t c_r = (wi + 1 < g.row_words) ? (s.C[tid + 1] & 1u) : 0u;
                    uint32_t ul = (u << 1) | u_l, ur = (u >> 1) | (u_r << 31);
                    uint32_t cl = (c << 1) | c_l, cr = (c >> 1) | (c_r << 31);
                    uint32_t needA = c & u & ~(cl & ul);
                    uint32_t needB = c & ul & ~u & ~cl;
                    uint32_t needC = c & ur & ~u & ~cr;
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
                // ---------------------------------------------------- roots -> component slots
                if (tid < g.words && stw) {
                    int base = s.base[tid];
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
                ncomp = s.ncomp;
                // ---------------------------------------------------- per-run statistics
                if (tid < g.words && stw) {
                    int base = s.base[tid];
                    uint32_t bitsleft = stw;
                    int n = 0;
                    while (bitsleft) {
                        int b = __ffs(bitsleft) - 1;
                        bitsleft &= bitsleft - 1;
                        int id = y * kRunsPerRow + base + n;
                        ++n;
                        int slot = uf_slot(s.parent, id);
                        s.parent[id] = (uint16_t)(kSlotFlag | slot);
                        // run length: ones from bit b upward, continuing into following words
                        int xs = wi * 32 + b;
                        uint32_t inv = ~(c >> b);
                        int len = (inv == 0) ? 32 : (__ffs(inv) - 1);
                        if (b + len >= 32) {
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
                // ---------------------------------------------------- OpenCV label order
                const int nslots = min(ncomp, CPT_MAX_COMPONENTS);
                if (tid < nslots) {
                    int key = s.c_key[tid], rank = 0;
                    for (int q = 0; q < nslots; ++q) rank += (s.c_key[q] < key);
                    s.c_rank[tid] = (uint8_t)rank;
                }
                __syncthreads();
                // ---------------------------------------------------- label image fill (runs)
                if (a.labels && tid < g.words && stw) {
                    int base = s.base[tid];
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
                // ---------------------------------------------------- regions + variance (K5, K6)
                {
                    const double mn = (double)cur_fmin, mx = (double)cur_fmax;
                    const double pmn = (double)prev_fmin, pmx = (double)prev_fmax;
                    const int nout = min(nslots, g.max_regions);
                    for (int slot = warp; slot < nslots; slot += kWarps) {
                        int rank = s.c_rank[slot];
                        if (rank >= nout) continue;
                        int l = s.c_l[slot], tp = s.c_t[slot];
                        int bw = s.c_r[slot] - l + 1, bh = s.c_b[slot] - tp + 1;
                        double var = 0.0;
                        if (have_prev) {
                            int n = bw * bh;
                            double sum = 0.0;
                            for (int i = lane; i < n; i += 32) {
                                int yy = tp + i / bw, xx = l + i % bw;
                                float d = fabsf(norm255_f64(fcur[yy * W + xx], mn, mx) - norm255_f64(fprev[yy * W + xx], pmn, pmx));
                                sum += (double)d;
                            }
                            for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                            double mean = sum / (double)n, s2 = 0.0;
                            for (int i = lane; i < n; i += 32) {
                                int yy = tp + i / bw, xx = l + i % bw;
                                float d = fabsf(norm255_f64(fcur[yy * W + xx], mn, mx) - norm255_f64(fprev[yy * W + xx], pmn, pmx));
                                double dd = (double)d - mean;
                                s2 += dd * dd;
                            }
                            for (int off = 16; off; off >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
                            var = s2 / (double)n;
                        }
                        if (lane == 0) {
                            cpt_region r;
                            r.x = l; r.y = tp; r.width = bw; r.height = bh;
                            r.area = s.c_area[slot]; r.sum_x = s.c_sx[slot]; r.sum_y = s.c_sy[slot];
                            r.key = s.c_key[slot]; r.pixel_variance = var;
                            a.regions[o * g.max_regions + rank] = r;
                        }
                    }
                    if (tid == 0) a.info[o].n_components = ncomp;
                }
                __syncthreads();
            }
            // ------------------------------------------------------------ label image out
            if (a.labels) {
                uint4 *dst = reinterpret_cast<uint4 *>(a.labels + o * npx);
                for (int i = tid; i < npx / 16; i += kThreads) dst[i] = reinterpret_cast<const uint4 *>(s.U)[i];
            }

            // ------------------------------------------------------------ sweep 3: background (K7)
            if (clip.flags & CPT_CLIP_UPDATE_BACKGROUND) {
                const uint32_t cnt = (uint32_t)min(t_abs + 1, kMeanFrames);
                const uint32_t magic = (uint32_t)(0x100000000ull / cnt) + 1u;  // exact for S < 2^22, cnt <= 45
                uint32_t bsum = 0;
                int changed = 0;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    int grp = tid + j * kThreads;
                    if (grp < g.groups) {
                        int yy = grp / g.gpr, x0 = (grp - yy * g.gpr) * 8;
                        bool row_in = (yy >= g.edge) && (yy < g.H - g.edge);
                        if (row_in) {
                            int b[8], k[8];
                            unpack8(*reinterpret_cast<const uint4 *>(s.B + grp * 8), b);
                            unpack8(*reinterpret_cast<const uint4 *>(s.K + grp * 8), k);
                            const uint4 *sp = reinterpret_cast<const uint4 *>(s.S + grp * 8);
                            uint4 s0 = sp[0], s1 = sp[1];
                            uint32_t sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                int x = x0 + i;
                                if (x >= g.edge && x < W - g.edge) {
                                    int A = (int)__umulhi(sv[i], magic);
                                    int d = A - b[i];
                                    int kk = k[i];
                                    int ck = (int)__ldg(wt.ceil_w + kk);
                                    bool keep = d > ck;
                                    if (d == ck) {
                                        // literal fp64 test of motiondetector.py:214-218
                                        double rhs = __dsub_rn((double)A, __ldg(wt.w + kk));
                                        keep = (double)b[i] < rhs;
                                    }
                                    if (keep) {
                                        k[i] = min(kk + 1, wt.max_count);
                                    } else {
                                        changed |= (A != b[i]);
                                        b[i] = A;
                                        k[i] = 0;
                                    }
                                    bsum += (uint32_t)b[i];
                                }
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
        __syncthreads();
    }
}

}  // namespace cpt
