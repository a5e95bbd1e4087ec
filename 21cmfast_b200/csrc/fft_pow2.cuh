/*
 * fft_pow2.cuh -- power-of-two fast path of the batched 1-D FFTs (included by fft.cu, CUDA only).
 *
 * Each line of N points is owned by N/R threads; a thread keeps R = 8 points in registers for the
 * whole transform (positions t + m N/R) and runs radix-8/4/2 Stockham stages on them.  Between
 * stages the CTA's TL lines are transposed through ONE shared-memory tile (no ping-pong buffer):
 * write outputs at their Stockham positions, barrier, read back positions t + m N/R.  The last
 * stage's outputs already sit at t + m N/R, so the first load and the final store go straight
 * between registers and global memory.
 *
 * Tile layout: point i of line c at [i * TL + (c ^ (i & (TL-1)))].  With TL = 16 a half-warp of
 * the butterfly phase touches exactly one 128-byte row (conflict-free for every stage pattern),
 * and the XOR keeps the column-wise accesses of the contiguous-axis (z) kernels conflict-free too.
 *
 * z axis: a length-NR real transform runs as a length-NR/2 complex FFT on (even, odd) pairs with
 * the standard Hermitian split/merge step, so the z pass costs half the butterflies and half the
 * shared memory of the generic kernel.
 */
#pragma once

template <int TL> DEV int sidx(int i, int c) { return i * TL + (c ^ (i & (TL - 1))); }

/* one Stockham stage on the registers of a thread; RAD <= R, NS = product of earlier radices */
template <int N, int R, int RAD, int NS, int SIGN>
DEV void stage_regs(float2 (&v)[R], int t, const float2 *__restrict__ tw, int tw_stride) {
    constexpr int NB = R / RAD;   /* butterflies per thread */
    constexpr int STEP = N / R;   /* distance between the positions a thread holds */
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int j = t + b * STEP;
        float2 u[RAD];
#pragma unroll
        for (int q = 0; q < RAD; q++) u[q] = v[b + NB * q];
        if (NS > 1) {
            const int k = j & (NS - 1);
            const int base = k * (N / (NS * RAD)) * tw_stride;
#pragma unroll
            for (int q = 1; q < RAD; q++) {
                float2 w = ldg(&tw[q * base]);
                if (SIGN > 0) w.y = -w.y;
                u[q] = cmul(u[q], w);
            }
        }
        small_dft<RAD>(u, SIGN);
#pragma unroll
        for (int q = 0; q < RAD; q++) v[b + NB * q] = u[q];
    }
}

/* where a line's point p lives in shared memory and how the owners of a line synchronise */
template <int TL> struct TilePolicy { /* TL lines interleaved, whole CTA works on the tile */
    float2 *S;
    int c;
    DEV float2 &at(int p) const { return S[sidx<TL>(p, c)]; }
    DEV void sync() const { __syncthreads(); }
};
template <int LT> struct LinePolicy { /* one line per LT threads, padded by one slot every 8 points */
    float2 *L;
    int bar_id;
    DEV float2 &at(int p) const { return L[p + (p >> 3)]; }
    DEV void sync() const {
        if (LT <= 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(LT) : "memory");
    }
};

/* transpose between stages: outputs of stage (RAD, NS) -> inputs t + m STEP of the next stage */
template <int N, int R, int RAD, int NS, class Pol>
DEV void exchange(float2 (&v)[R], int t, const Pol &pol) {
    constexpr int NB = R / RAD;
    constexpr int STEP = N / R;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int j = t + b * STEP;
        const int k = j & (NS - 1);
        const int j0 = (j - k) * RAD + k;
#pragma unroll
        for (int q = 0; q < RAD; q++) pol.at(j0 + q * NS) = v[b + NB * q];
    }
    pol.sync();
#pragma unroll
    for (int m = 0; m < R; m++) v[m] = pol.at(t + m * STEP);
    pol.sync();
}

/* full length-N transform of the line whose points t + m N/8 live in v */
template <int N, int SIGN, class Pol>
DEV void fft_line_regs(float2 (&v)[8], int t, const Pol &pol, const float2 *__restrict__ tw, int tws) {
    static_assert(N >= 16 && N <= 2048 && (N & (N - 1)) == 0, "power of two 16..2048");
    stage_regs<N, 8, 8, 1, SIGN>(v, t, tw, tws);
    exchange<N, 8, 8, 1>(v, t, pol);
    if constexpr (N == 16) {
        stage_regs<N, 8, 2, 8, SIGN>(v, t, tw, tws);
    } else if constexpr (N == 32) {
        stage_regs<N, 8, 4, 8, SIGN>(v, t, tw, tws);
    } else if constexpr (N == 64) {
        stage_regs<N, 8, 8, 8, SIGN>(v, t, tw, tws);
    } else {
        stage_regs<N, 8, 8, 8, SIGN>(v, t, tw, tws);
        exchange<N, 8, 8, 8>(v, t, pol);
        if constexpr (N == 128) {
            stage_regs<N, 8, 2, 64, SIGN>(v, t, tw, tws);
        } else if constexpr (N == 256) {
            stage_regs<N, 8, 4, 64, SIGN>(v, t, tw, tws);
        } else if constexpr (N == 512) {
            stage_regs<N, 8, 8, 64, SIGN>(v, t, tw, tws);
        } else {
            stage_regs<N, 8, 8, 64, SIGN>(v, t, tw, tws);
            exchange<N, 8, 8, 64>(v, t, pol);
            if constexpr (N == 1024) stage_regs<N, 8, 2, 512, SIGN>(v, t, tw, tws);
            else stage_regs<N, 8, 4, 512, SIGN>(v, t, tw, tws);
        }
    }
}

/* lines per CTA: 16 while the CTA stays within 1024 threads */
template <int N> struct Pow2Cfg {
    static constexpr int TL = (N <= 512) ? 16 : (N == 1024 ? 8 : 4);
    static constexpr int THREADS = TL * (N / 8);
    static constexpr int MIN_CTAS = THREADS <= 256 ? 4 : (THREADS <= 512 ? 2 : 1);
};

/* ------------------------------------------------------------------ strided axes (x, y)
 * Persistent CTAs: each loops over tiles of TL adjacent lines and issues the global loads of its
 * NEXT tile into a second register set before it transforms the current one, so the HBM latency
 * of tile k+1 hides behind the butterflies, barriers and stores of tile k even with one CTA per
 * SM. */
template <int N, int SIGN>
__global__ void __launch_bounds__(Pow2Cfg<N>::THREADS, Pow2Cfg<N>::MIN_CTAS) fft_strided_pow2_kernel(const float2 *__restrict__ src,
                                                                               float2 *__restrict__ dst,
                                                                               StridedArgs a, int tiles_per_group,
                                                                               int ntiles) {
    constexpr int TL = Pow2Cfg<N>::TL;
    constexpr int STEP = N / 8;
    DYN_SMEM(float2, S);
    const int c = threadIdx.x & (TL - 1), t = threadIdx.x / TL;
    const bool has_kmul = a.kmul != KMUL_NONE || a.op != KOP_NONE;
    float2 v[8], vn[8];
    int tile = blockIdx.x;
    /* prologue: first tile */
    {
        const int g = tile / tiles_per_group, col = (tile - g * tiles_per_group) * TL + c;
        const long long base = (long long)g * a.group_stride + col;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            vn[m] = make_float2(0.f, 0.f);
            if (tile < ntiles && col < a.ncols) vn[m] = src[base + (long long)(t + m * STEP) * a.line_stride];
        }
    }
    for (; tile < ntiles; tile += gridDim.x) {
        const int g = tile / tiles_per_group, col = (tile - g * tiles_per_group) * TL + c;
        const bool live = col < a.ncols;
        const long long base = (long long)g * a.group_stride + col;
#pragma unroll
        for (int m = 0; m < 8; m++) v[m] = vn[m];
        /* prefetch the next tile of this CTA */
        {
            const int nt = tile + gridDim.x;
            const int gn = nt / tiles_per_group, coln = (nt - gn * tiles_per_group) * TL + c;
            const long long basen = (long long)gn * a.group_stride + coln;
            const bool liven = nt < ntiles && coln < a.ncols;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                vn[m] = make_float2(0.f, 0.f);
                if (liven) vn[m] = src[basen + (long long)(t + m * STEP) * a.line_stride];
            }
        }
        if (has_kmul && live) {
            if (a.wtab && a.op == KOP_NONE) {
                /* window from the |n|^2 table: the (y, kz) part is a per-thread constant of the tile */
                const int iy = col / a.pitch, iz = col - iy * a.pitch;
                if (iz < a.nzc) {
                    const int sy = (iy > a.ny / 2) ? iy - a.ny : iy;
                    const int syz = sy * sy + iz * iz;
#pragma unroll
                    for (int m = 0; m < 8; m++) {
                        const int i = t + m * STEP;
                        const int sx = (i > N / 2) ? i - N : i;
                        const float W = ldg(&a.wtab[sx * sx + syz]);
                        v[m].x *= W;
                        v[m].y *= W;
                    }
                }
            } else {
#pragma unroll
                for (int m = 0; m < 8; m++) v[m] = apply_kmul(v[m], t + m * STEP, col, a);
            }
        }
        fft_line_regs<N, SIGN>(v, t, TilePolicy<TL>{S, c}, a.tw, 1);
        if (live) {
#pragma unroll
            for (int m = 0; m < 8; m++) {
                float2 x = v[m];
                if (a.scale != 1.f) { x.x *= a.scale; x.y *= a.scale; }
                dst[base + (long long)(t + m * STEP) * a.line_stride] = x;
            }
        }
        __syncthreads(); /* the tile buffer is reused by the next iteration */
    }
}

template <int N> static void launch_strided_pow2(const float2 *src, float2 *dst, const StridedArgs &a, int ngroups) {
    constexpr int TL = Pow2Cfg<N>::TL;
    const size_t smem = (size_t)N * TL * sizeof(float2);
    const int tiles_per_group = (a.ncols + TL - 1) / TL;
    const long long ntiles = (long long)tiles_per_group * ngroups;
    auto kf = &fft_strided_pow2_kernel<N, -1>;
    auto ki = &fft_strided_pow2_kernel<N, 1>;
    /* persistent grid: exactly the CTAs that are resident at once */
    allow_smem(kf, smem);
    allow_smem(ki, smem);
    static int per_sm = 0;
    if (per_sm == 0) {
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ki, Pow2Cfg<N>::THREADS, smem));
        if (per_sm < 1) per_sm = 1;
    }
    long long grid = (long long)dev_num_sms() * per_sm;
    if (grid > ntiles) grid = ntiles;
    if (a.sign < 0) {
        allow_smem(kf, smem);
        B200_LAUNCH_T("fft_strided_pow2_kernel", kf, dim3((unsigned)grid), Pow2Cfg<N>::THREADS, smem, src, dst, a,
                      tiles_per_group, (int)ntiles);
    } else {
        allow_smem(ki, smem);
        B200_LAUNCH_T("fft_strided_pow2_kernel", ki, dim3((unsigned)grid), Pow2Cfg<N>::THREADS, smem, src, dst, a,
                      tiles_per_group, (int)ntiles);
    }
}

static bool pow2_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("B200_FFT_GENERIC");
        on = (e && e[0] == '1') ? 0 : 1;
    }
    return on == 1;
}

static bool pow2_strided(const float2 *src, float2 *dst, const StridedArgs &a, int ngroups) {
    if (!pow2_enabled()) return false;
    switch (a.n) {
        case 16: launch_strided_pow2<16>(src, dst, a, ngroups); return true;
        case 32: launch_strided_pow2<32>(src, dst, a, ngroups); return true;
        case 64: launch_strided_pow2<64>(src, dst, a, ngroups); return true;
        case 128: launch_strided_pow2<128>(src, dst, a, ngroups); return true;
        case 256: launch_strided_pow2<256>(src, dst, a, ngroups); return true;
        case 512: launch_strided_pow2<512>(src, dst, a, ngroups); return true;
        case 1024: launch_strided_pow2<1024>(src, dst, a, ngroups); return true;
        case 2048: launch_strided_pow2<2048>(src, dst, a, ngroups); return true;
        default: return false;
    }
}

/* ------------------------------------------------------------------ contiguous axis (z) */
/* complex rows (NH + 1 values) -> NR = 2 NH reals per row.
   The z axis is contiguous, so a line can be owned by NH/8 consecutive lanes (one warp at
   nz = 512): spectrum loads and real stores are coalesced straight from/to registers, the
   inter-stage transposes use a private padded strip of shared memory, and lines synchronise with
   __syncwarp / a per-line named barrier -- no CTA-wide barrier anywhere. */
template <int NH> struct ZLineCfg {
    static constexpr int LT = NH / 8;                       /* threads per line */
    static constexpr int THREADS = LT > 256 ? LT : 256;
    static constexpr int LINES = THREADS / LT;              /* lines per CTA */
    static constexpr int STRIP = NH + NH / 8;               /* padded points per line */
};
template <int NH>
__global__ void __launch_bounds__(ZLineCfg<NH>::THREADS)
fft_c2r_z_pow2_kernel(const float2 *__restrict__ src, float *__restrict__ dst, ZArgs a) {
    using Cfg = ZLineCfg<NH>;
    constexpr int LT = Cfg::LT, STEP = NH / 8;
    DYN_SMEM(float2, S);
    __shared__ float red_min[32], red_max[32];
    const int line = threadIdx.x / LT, t = threadIdx.x - line * LT;
    const LinePolicy<LT> pol{S + line * Cfg::STRIP, 1 + line};
    float2 *dst2 = reinterpret_cast<float2 *>(dst);
    const long long row_stride2 = a.real_row_stride / 2;
    float lmin = 3.0e38f, lmax = -3.0e38f;
    /* uniform trip count for every thread of the CTA (the line barriers need all owners) */
    for (long long row0 = (long long)blockIdx.x * Cfg::LINES; row0 < a.nrows;
         row0 += (long long)gridDim.x * Cfg::LINES) {
        const long long row = row0 + line;
        const bool live = row < a.nrows;
        const float2 *X = src + row * a.pitch;
        float2 xa[8], xb[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int n = t + m * STEP;
            xa[m] = live ? X[n] : make_float2(0.f, 0.f);
            xb[m] = live ? X[NH - n] : make_float2(0.f, 0.f);
        }
        if (t == 0) { xa[0].y = 0.f; xb[0].y = 0.f; } /* DC and Nyquist are real */
        /* Z[n] = (X[n] + conj X[NH-n]) + i e^{+2 pi i n / NR} (X[n] - conj X[NH-n]) */
        float2 v[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int n = t + m * STEP;
            float2 B = xb[m];
            B.y = -B.y;
            const float2 sum = cadd(xa[m], B), dif = csub(xa[m], B);
            float2 w = ldg(&a.tw[n]); /* e^{-2 pi i n / NR} */
            w.y = -w.y;
            const float2 p = cmul(w, dif);
            v[m] = make_float2(sum.x - p.y, sum.y + p.x);
        }
        fft_line_regs<NH, 1>(v, t, pol, a.tw, 2);
        /* z[n] = (x[2n], x[2n+1]) */
#pragma unroll
        for (int m = 0; m < 8; m++) {
            float2 x = v[m];
            x.x *= a.scale; x.y *= a.scale;
            if (live) {
                lmin = fminf(lmin, fminf(x.x, x.y));
                lmax = fmaxf(lmax, fmaxf(x.x, x.y));
            }
            if (a.clip) {
                x.x = fmaxf(fminf(x.x, a.clip_hi), a.clip_lo);
                x.y = fmaxf(fminf(x.y, a.clip_hi), a.clip_lo);
            }
            if (live) dst2[row * row_stride2 + t + m * STEP] = x;
        }
    }
    if (a.minmax_keys) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
            lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        }
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) { red_min[warp] = lmin; red_max[warp] = lmax; }
        __syncthreads();
        if (warp == 0) {
            const int nw = Cfg::THREADS / 32;
            lmin = lane < nw ? red_min[lane] : 3.0e38f;
            lmax = lane < nw ? red_max[lane] : -3.0e38f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
                lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
            }
            if (lane == 0) {
                atomic_min_i32(&a.minmax_keys[0], float_order_key(float_as_int_bits(lmin)));
                atomic_max_i32(&a.minmax_keys[1], float_order_key(float_as_int_bits(lmax)));
            }
        }
    }
}

/* NR = 2 NH reals per row -> NH + 1 complex values */
template <int NH>
__global__ void __launch_bounds__(Pow2Cfg<NH>::THREADS) fft_r2c_z_pow2_kernel(const float *__restrict__ src,
                                                                             float2 *__restrict__ dst, ZArgs a) {
    constexpr int TL = Pow2Cfg<NH>::TL;
    constexpr int STEP = NH / 8;
    constexpr int NT = Pow2Cfg<NH>::THREADS;
    DYN_SMEM(float2, S);
    const long long row0 = (long long)blockIdx.x * TL;
    const float2 *src2 = reinterpret_cast<const float2 *>(src);
    const long long row_stride2 = a.real_row_stride / 2;
    for (int e = threadIdx.x; e < TL * NH; e += NT) {
        const int l = e / NH, n = e - l * NH;
        float2 x = make_float2(0.f, 0.f);
        if (row0 + l < a.nrows) {
            x = src2[(row0 + l) * row_stride2 + n];
            if (a.premul != 1.f || a.clip) {
                double c0 = (double)x.x * (double)a.premul, c1 = (double)x.y * (double)a.premul;
                if (a.clip) {
                    c0 = fmax(fmin(c0, (double)a.clip_hi), (double)a.clip_lo);
                    c1 = fmax(fmin(c1, (double)a.clip_hi), (double)a.clip_lo);
                }
                x = make_float2((float)c0, (float)c1);
            }
        }
        S[sidx<TL>(n, l)] = x;
    }
    __syncthreads();
    const int c = threadIdx.x & (TL - 1), t = threadIdx.x / TL;
    float2 v[8];
#pragma unroll
    for (int m = 0; m < 8; m++) v[m] = S[sidx<TL>(t + m * STEP, c)];
    __syncthreads();
    fft_line_regs<NH, -1>(v, t, TilePolicy<TL>{S, c}, a.tw, 2);
#pragma unroll
    for (int m = 0; m < 8; m++) S[sidx<TL>(t + m * STEP, c)] = v[m];
    __syncthreads();
    /* X[k] = (Z[k] + conj Z[NH-k]) / 2 - (i/2) e^{-2 pi i k / NR} (Z[k] - conj Z[NH-k]) */
    for (int e = threadIdx.x; e < TL * NH; e += NT) {
        const int l = e / NH, k = e - l * NH;
        if (row0 + l < a.nrows) {
            const float2 Zk = S[sidx<TL>(k, l)];
            float2 Zc = S[sidx<TL>((NH - k) & (NH - 1), l)];
            Zc.y = -Zc.y;
            const float2 s = cadd(Zk, Zc), d = csub(Zk, Zc);
            const float2 p = cmul(ldg(&a.tw[k]), d);
            float2 X = make_float2(0.5f * (s.x + p.y), 0.5f * (s.y - p.x));
            if (a.scale != 1.f) { X.x *= a.scale; X.y *= a.scale; }
            dst[(row0 + l) * a.pitch + k] = X;
        }
    }
    if (threadIdx.x < TL) {
        const int l = threadIdx.x;
        if (row0 + l < a.nrows) {
            const float2 Z0 = S[sidx<TL>(0, l)];
            float2 X = make_float2(Z0.x - Z0.y, 0.f);
            if (a.scale != 1.f) X.x *= a.scale;
            dst[(row0 + l) * a.pitch + NH] = X;
        }
    }
}

template <int NH> static void launch_c2r_z_pow2(const float2 *src, float *dst, const ZArgs &a) {
    using Cfg = ZLineCfg<NH>;
    const size_t smem = (size_t)Cfg::LINES * Cfg::STRIP * sizeof(float2);
    auto k = &fft_c2r_z_pow2_kernel<NH>;
    allow_smem(k, smem);
    static int per_sm = 0;
    if (per_sm == 0) {
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, Cfg::THREADS, smem));
        if (per_sm < 1) per_sm = 1;
    }
    const long long groups = ((long long)a.nrows + Cfg::LINES - 1) / Cfg::LINES;
    long long grid = (long long)dev_num_sms() * per_sm;
    if (grid > groups) grid = groups;
    B200_LAUNCH_T("fft_c2r_z_pow2_kernel", k, dim3((unsigned)grid), Cfg::THREADS, smem, src, dst, a);
}
template <int NH> static void launch_r2c_z_pow2(const float *src, float2 *dst, const ZArgs &a) {
    constexpr int TL = Pow2Cfg<NH>::TL;
    const size_t smem = (size_t)(NH + 1) * TL * sizeof(float2);
    auto k = &fft_r2c_z_pow2_kernel<NH>;
    allow_smem(k, smem);
    B200_LAUNCH_T("fft_r2c_z_pow2_kernel", k, dim3((a.nrows + TL - 1) / TL), Pow2Cfg<NH>::THREADS, smem, src, dst, a);
}

static bool z_ok(const ZArgs &a, const void *real_ptr) {
    return pow2_enabled() && (a.real_row_stride % 2 == 0) && (((uintptr_t)real_ptr) % 8 == 0);
}
static bool pow2_c2r_z(const float2 *src, float *dst, const ZArgs &a) {
    if (!z_ok(a, dst)) return false;
    switch (a.n) {
        case 32: launch_c2r_z_pow2<16>(src, dst, a); return true;
        case 64: launch_c2r_z_pow2<32>(src, dst, a); return true;
        case 128: launch_c2r_z_pow2<64>(src, dst, a); return true;
        case 256: launch_c2r_z_pow2<128>(src, dst, a); return true;
        case 512: launch_c2r_z_pow2<256>(src, dst, a); return true;
        case 1024: launch_c2r_z_pow2<512>(src, dst, a); return true;
        case 2048: launch_c2r_z_pow2<1024>(src, dst, a); return true;
        default: return false;
    }
}
static bool pow2_r2c_z(const float *src, float2 *dst, const ZArgs &a) {
    if (!z_ok(a, src)) return false;
    switch (a.n) {
        case 32: launch_r2c_z_pow2<16>(src, dst, a); return true;
        case 64: launch_r2c_z_pow2<32>(src, dst, a); return true;
        case 128: launch_r2c_z_pow2<64>(src, dst, a); return true;
        case 256: launch_r2c_z_pow2<128>(src, dst, a); return true;
        case 512: launch_r2c_z_pow2<256>(src, dst, a); return true;
        case 1024: launch_r2c_z_pow2<512>(src, dst, a); return true;
        case 2048: launch_r2c_z_pow2<1024>(src, dst, a); return true;
        default: return false;
    }
}
