/*
 * fft_pow2.cuh -- power-of-two fast path of the batched 1-D FFTs (included by fft.cu, CUDA only).
 *
 * Each line of N points is owned by N/8 threads; a thread keeps 8 points in registers for the
 * whole transform (positions t + m N/8) and runs radix-8/4/2 Stockham stages on them.  Between
 * stages the lines are transposed through shared memory: write outputs at their Stockham
 * positions, barrier, read back positions t + m N/8.  Two tile buffers alternate between the
 * exchanges, so every exchange costs ONE barrier.  The last stage's outputs already sit at
 * t + m N/8, so the first load and the final store go straight between registers and global memory.
 *
 * Instruction diet (the strided passes are issue-bound before they are HBM-bound):
 *   - complex arithmetic on Blackwell's packed FADD2 / FMUL2 / FFMA2 (two floats per instruction;
 *     the +-i rotations ride on the operand swizzles),
 *   - the stage twiddles live in shared memory as (w, i w) float4 rows indexed by the thread's
 *     butterfly phase: one LDS.128 + FMUL2 + FFMA2 per complex multiply, no index arithmetic,
 *   - tile rows are unswizzled [point][line]: a half-warp owns one 128-byte row, every exchange
 *     access is a per-thread base plus a compile-time offset,
 *   - the k-space multipliers are a template parameter, so the hot instantiations carry no
 *     double-precision window code.
 *
 * z axis: a length-NR real transform runs as a length-NR/2 complex FFT on (even, odd) pairs with
 * the standard Hermitian split/merge step, so the z pass costs half the butterflies and half the
 * shared memory of the generic kernel.
 */
#pragma once

/* ------------------------------------------------------------------ packed complex helpers */
DEV float2 pk_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
DEV float2 pk_sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
template <int S> DEV float2 pk_rot(float2 a) { return S > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }
DEV float2 pk_scale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
/* a * w with w given as (w.x, w.y, -w.y, w.x) */
DEV float2 pk_cmul4(float2 a, float4 w) {
    const float2 t = __fmul2_rn(make_float2(a.y, a.y), make_float2(w.z, w.w));
    return __ffma2_rn(make_float2(a.x, a.x), make_float2(w.x, w.y), t);
}
DEV float2 pk_cmul(float2 a, float2 w) { return pk_cmul4(a, make_float4(w.x, w.y, -w.y, w.x)); }

template <int S> DEV void pk_bfly2(float2 *v) {
    const float2 a = v[0], b = v[1];
    v[0] = pk_add(a, b);
    v[1] = pk_sub(a, b);
}
template <int S> DEV void pk_bfly4(float2 *v) {
    const float2 a = pk_add(v[0], v[2]), b = pk_sub(v[0], v[2]);
    const float2 c = pk_add(v[1], v[3]), d = pk_rot<S>(pk_sub(v[1], v[3]));
    v[0] = pk_add(a, c);
    v[1] = pk_add(b, d);
    v[2] = pk_sub(a, c);
    v[3] = pk_sub(b, d);
}
template <int S> DEV void pk_bfly8(float2 *v) {
    const float h = 0.70710678118654752440f;
    float2 e[4] = {v[0], v[2], v[4], v[6]};
    float2 o[4] = {v[1], v[3], v[5], v[7]};
    pk_bfly4<S>(e);
    pk_bfly4<S>(o);
    /* W8^k o[k], W8 = exp(S 2 pi i / 8): (1 + S i)/sqrt2, S i, (-1 + S i)/sqrt2 */
    const float2 t1 = pk_scale(pk_add(o[1], pk_rot<S>(o[1])), h);
    const float2 t2 = pk_rot<S>(o[2]);
    const float2 t3 = pk_scale(pk_sub(pk_rot<S>(o[3]), o[3]), h);
    v[0] = pk_add(e[0], o[0]); v[4] = pk_sub(e[0], o[0]);
    v[1] = pk_add(e[1], t1);   v[5] = pk_sub(e[1], t1);
    v[2] = pk_add(e[2], t2);   v[6] = pk_sub(e[2], t2);
    v[3] = pk_add(e[3], t3);   v[7] = pk_sub(e[3], t3);
}
template <int RAD, int S> DEV void pk_dft(float2 *v) {
    if constexpr (RAD == 8) pk_bfly8<S>(v);
    else if constexpr (RAD == 4) pk_bfly4<S>(v);
    else pk_bfly2<S>(v);
}

/* ------------------------------------------------------------------ stage plan of a length-N line
 * stages: radix 8 at NS = 1, 8, 64 while they fit, then one radix 2/4/8 stage that completes N. */
template <int N> struct Pow2Plan {
    static_assert(N >= 16 && N <= 2048 && (N & (N - 1)) == 0, "power of two 16..2048");
    static constexpr int STEP = N / 8;
    static constexpr int NSTAGE = N <= 64 ? 2 : (N <= 512 ? 3 : 4);
    static constexpr int ns(int s) { return s == 0 ? 1 : (s == 1 ? 8 : (s == 2 ? 64 : 512)); }
    static constexpr int rad(int s) { return (s == NSTAGE - 1) ? N / ns(s) : 8; }
    /* float4 twiddle rows: stage s (s >= 1) holds ns(s) phases x (rad - 1) multipliers */
    static constexpr int tw_base(int s) { return s <= 1 ? 0 : (s == 2 ? 56 : 504); }
    static constexpr int TW_TOTAL = tw_base(NSTAGE - 1) + ns(NSTAGE - 1) * (rad(NSTAGE - 1) - 1);
};

/* fill the shared twiddle table from the global exp(-2 pi i k / (tws N)) table (tws = 1 or 2) */
template <int N, int SIGN> DEV void fill_twiddles(float4 *twS, const float2 *__restrict__ tw, int tws, int nthreads) {
    using P = Pow2Plan<N>;
#pragma unroll
    for (int s = 1; s < P::NSTAGE; s++) {
        const int NS = P::ns(s), RAD = P::rad(s);
        for (int e = threadIdx.x; e < NS * (RAD - 1); e += nthreads) {
            const int k = e / (RAD - 1), q = 1 + e - k * (RAD - 1);
            float2 w = ldg(&tw[q * k * (N / (NS * RAD)) * tws]);
            if (SIGN > 0) w.y = -w.y;
            twS[P::tw_base(s) + e] = make_float4(w.x, w.y, -w.y, w.x);
        }
    }
}

/* one Stockham stage on the registers of a thread.  twp = row of the thread's phase for
   butterfly 0 (phases of butterfly b are b * STEP rows further: only in the last stage, where
   t + b STEP < NS, see Pow2Plan) */
template <int N, int RAD, int NS, int SIGN> DEV void stage_regs(float2 (&v)[8], const float4 *twp) {
    constexpr int NB = 8 / RAD;
    constexpr int STEP = N / 8;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        float2 u[RAD];
#pragma unroll
        for (int q = 0; q < RAD; q++) u[q] = v[b + NB * q];
        if constexpr (NS > 1) {
#pragma unroll
            for (int q = 1; q < RAD; q++) u[q] = pk_cmul4(u[q], twp[b * STEP * (RAD - 1) + (q - 1)]);
        }
        pk_dft<RAD, SIGN>(u);
#pragma unroll
        for (int q = 0; q < RAD; q++) v[b + NB * q] = u[q];
    }
}

/* Register-resident twiddles for kernels whose threads keep the same butterfly phases for every
   line they transform (the z kernels: one line per LT lanes, so the shared-memory rows would be
   read with 32 distinct addresses per warp, 4 wavefronts per LDS.128, for every line).  A radix-8
   stage keeps w, w^2, w^4 and derives the other four powers with one complex multiply each (two
   roundings at most); the final radix-4 / radix-2 stage keeps (w, w^2) resp. w per butterfly. */
template <int N> struct RegTwiddles {
    float2 w[Pow2Plan<N>::NSTAGE - 1][4];
};
template <int N> DEV void load_reg_twiddles(RegTwiddles<N> &r, int t, const float4 *twS) {
    using P = Pow2Plan<N>;
    constexpr int STEP = N / 8;
#pragma unroll
    for (int s = 1; s < P::NSTAGE; s++) {
        const int NS = P::ns(s), RAD = P::rad(s);
        const float4 *row = twS + P::tw_base(s) + (t & (NS - 1)) * (RAD - 1);
        if (RAD == 8) {
            r.w[s - 1][0] = make_float2(row[0].x, row[0].y);
            r.w[s - 1][1] = make_float2(row[1].x, row[1].y);
            r.w[s - 1][2] = make_float2(row[3].x, row[3].y);
            r.w[s - 1][3] = make_float2(0.f, 0.f);
        } else if (RAD == 4) {
#pragma unroll
            for (int b = 0; b < 2; b++) {
                r.w[s - 1][2 * b] = make_float2(row[b * STEP * 3].x, row[b * STEP * 3].y);
                r.w[s - 1][2 * b + 1] = make_float2(row[b * STEP * 3 + 1].x, row[b * STEP * 3 + 1].y);
            }
        } else {
#pragma unroll
            for (int b = 0; b < 4; b++) r.w[s - 1][b] = make_float2(row[b * STEP].x, row[b * STEP].y);
        }
    }
}
template <int N, int RAD, int SIGN> DEV void stage_regs_rt(float2 (&v)[8], const float2 (&w)[4]) {
    constexpr int NB = 8 / RAD;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        float2 u[RAD];
#pragma unroll
        for (int q = 0; q < RAD; q++) u[q] = v[b + NB * q];
        if constexpr (RAD == 8) {
            const float2 w1 = w[0], w2 = w[1], w4 = w[2];
            const float2 w3 = pk_cmul(w1, w2);
            u[1] = pk_cmul(u[1], w1);
            u[2] = pk_cmul(u[2], w2);
            u[3] = pk_cmul(u[3], w3);
            u[4] = pk_cmul(u[4], w4);
            u[5] = pk_cmul(u[5], pk_cmul(w4, w1));
            u[6] = pk_cmul(u[6], pk_cmul(w4, w2));
            u[7] = pk_cmul(u[7], pk_cmul(w4, w3));
        } else if constexpr (RAD == 4) {
            const float2 w1 = w[2 * b], w2 = w[2 * b + 1];
            u[1] = pk_cmul(u[1], w1);
            u[2] = pk_cmul(u[2], w2);
            u[3] = pk_cmul(u[3], pk_cmul(w1, w2));
        } else {
            u[1] = pk_cmul(u[1], w[b]);
        }
        pk_dft<RAD, SIGN>(u);
#pragma unroll
        for (int q = 0; q < RAD; q++) v[b + NB * q] = u[q];
    }
}

/* where a line's point p lives in shared memory; offq<D>(p0, q) addresses point p0 + q D with the
   q-dependent part a compile-time constant wherever the layout allows */
template <int TL> struct TilePolicy { /* TL lines interleaved: [point][line] */
    float2 *A, *B; /* the two exchange buffers */
    int c;
    DEV int off(int p) const { return p * TL + c; }
    template <int D> DEV int offq(int p0, int q) const { return off(p0) + q * D * TL; }
    DEV void sync() const { __syncthreads(); }
};
template <int LT> struct LinePolicy { /* one line per LT threads, padded by one slot every 8 points */
    float2 *A, *B;
    int bar_id;
    DEV int off(int p) const { return p + (p >> 3); }
    template <int D> DEV int offq(int p0, int q) const {
        if constexpr (D % 8 == 0) return off(p0) + q * (D + D / 8);
        else return off(p0 + q * D);
    }
    DEV void sync() const {
        if (LT <= 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(LT) : "memory");
    }
};

/* transpose between stages: outputs of stage (RAD, NS) -> inputs t + m STEP of the next stage */
template <int N, int RAD, int NS, class Pol, bool WAIT_ASYNC = false>
DEV void exchange(float2 (&v)[8], int t, float2 *buf, const Pol &pol) {
    constexpr int NB = 8 / RAD;
    constexpr int STEP = N / 8;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int j = t + b * STEP;
        const int k = j & (NS - 1);
        const int j0 = (j - k) * RAD + k;
#pragma unroll
        for (int q = 0; q < RAD; q++) buf[pol.template offq<NS>(j0, q)] = v[b + NB * q];
    }
    if (WAIT_ASYNC) cp_async_wait_all(); /* this thread's staged copies land before the barrier publishes them */
    pol.sync();
#pragma unroll
    for (int m = 0; m < 8; m++) v[m] = buf[pol.template offq<STEP>(t, m)];
}

/* full length-N transform of the line whose points t + m N/8 live in v.  Exchanges alternate
   between pol.A and pol.B; a caller that loops must keep one barrier between two transforms
   only if N has a single exchange (N <= 64), see the kernels. */
template <int N, int SIGN, class Pol, bool WAIT_ASYNC = false>
DEV void fft_line_regs(float2 (&v)[8], int t, const Pol &pol, const float4 *twS) {
    using P = Pow2Plan<N>;
    stage_regs<N, 8, 1, SIGN>(v, nullptr);
    exchange<N, 8, 1, Pol, WAIT_ASYNC && P::NSTAGE == 2>(v, t, pol.A, pol);
    if constexpr (P::NSTAGE == 2) {
        stage_regs<N, P::rad(1), 8, SIGN>(v, twS + (t & 7) * (P::rad(1) - 1));
    } else {
        stage_regs<N, 8, 8, SIGN>(v, twS + (t & 7) * 7);
        exchange<N, 8, 8, Pol, WAIT_ASYNC && P::NSTAGE == 3>(v, t, pol.B, pol);
        if constexpr (P::NSTAGE == 3) {
            stage_regs<N, P::rad(2), 64, SIGN>(v, twS + P::tw_base(2) + (t & 63) * (P::rad(2) - 1));
        } else {
            stage_regs<N, 8, 64, SIGN>(v, twS + P::tw_base(2) + (t & 63) * 7);
            /* A again: every thread read it before the barrier of the second exchange */
            exchange<N, 8, 64, Pol, WAIT_ASYNC>(v, t, pol.A, pol);
            stage_regs<N, P::rad(3), 512, SIGN>(v, twS + P::tw_base(3) + (t & 511) * (P::rad(3) - 1));
        }
    }
}
/* same transform with the twiddles of RegTwiddles (z kernels) */
template <int N, int SIGN, class Pol> DEV void fft_line_regs_rt(float2 (&v)[8], int t, const Pol &pol, const RegTwiddles<N> &rt) {
    using P = Pow2Plan<N>;
    stage_regs<N, 8, 1, SIGN>(v, nullptr);
    exchange<N, 8, 1, Pol>(v, t, pol.A, pol);
    if constexpr (P::NSTAGE == 2) {
        stage_regs_rt<N, P::rad(1), SIGN>(v, rt.w[0]);
    } else {
        stage_regs_rt<N, 8, SIGN>(v, rt.w[0]);
        exchange<N, 8, 8, Pol>(v, t, pol.B, pol);
        if constexpr (P::NSTAGE == 3) {
            stage_regs_rt<N, P::rad(2), SIGN>(v, rt.w[1]);
        } else {
            stage_regs_rt<N, 8, SIGN>(v, rt.w[1]);
            exchange<N, 8, 64, Pol>(v, t, pol.A, pol);
            stage_regs_rt<N, P::rad(3), SIGN>(v, rt.w[2]);
        }
    }
}
/* number of exchanges that touch buffer A / need a trailing barrier before the next transform */
template <int N> struct Pow2Sync {
    /* after fft_line_regs returns, the last exchange buffer may still be read by slower threads of
       the line; the next transform's first write goes to A.  With >= 2 exchanges the barrier of
       the second exchange already orders "all reads of A done" before anyone returns -- unless
       the last exchange itself used A (4 stages). */
    static constexpr bool NEED_TAIL_SYNC = (Pow2Plan<N>::NSTAGE != 3);
};

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
 * SM.  MODE: 0 plain, 1 window from the |n|^2 table, 2 generic k-space multiplier (apply_kmul). */
enum { PM_PLAIN = 0, PM_WTAB = 1, PM_GENERIC = 2 };
/* cp.async the window rows of one tile of the x pass into shared memory: row |nx| holds the TL
   columns of the tile, fetched as 16-byte chunks of 4 columns (a chunk never straddles a (y)
   row because the pitch is a multiple of 8).  Source: KMul::wtab3. */
template <int N, int TL> DEV void stage_window(float *Wbuf, const StridedArgs &a, int tile, int tiles_per_group, int ntiles) {
    constexpr int NA = N / 2 + 1, CH = TL / 4;
    if (tile < ntiles) {
        const int col0 = (tile % tiles_per_group) * TL;
        for (int e = threadIdx.x; e < NA * CH; e += Pow2Cfg<N>::THREADS) {
            const int ax = e / CH, j = e - ax * CH;
            const int col = col0 + 4 * j;
            if (col < a.ncols) {
                const int iyl = col / a.pitch, iz = col - iyl * a.pitch;
                const int iy = iyl + a.y_off;
                const int ay = (iy > a.ny / 2) ? a.ny - iy : iy;
                cp_async_16(Wbuf + ax * TL + 4 * j, a.wtab3 + ((long long)ax * a.w3_xstride + (long long)ay * a.pitch + iz));
            }
        }
    }
    cp_async_commit();
}
template <int N, int SIGN, int MODE, bool SCATTER>
__global__ void __launch_bounds__(Pow2Cfg<N>::THREADS, Pow2Cfg<N>::MIN_CTAS)
fft_strided_pow2_kernel(const float2 *__restrict__ src, float2 *__restrict__ dst, StridedArgs a, int tiles_per_group,
                        int ntiles) {
    using P = Pow2Plan<N>;
    constexpr int TL = Pow2Cfg<N>::TL;
    constexpr int STEP = N / 8;
    DYN_SMEM(float2, S);
    float4 *twS = reinterpret_cast<float4 *>(S + 2 * N * TL);
    constexpr int WROWS = N / 2 + 1;
    float *Wst = reinterpret_cast<float *>(twS + P::TW_TOTAL); /* [2][WROWS][TL] staged window rows (PM_WTAB) */
    fill_twiddles<N, SIGN>(twS, a.tw, 1, Pow2Cfg<N>::THREADS);
    const int c = threadIdx.x & (TL - 1), t = threadIdx.x / TL;
    const TilePolicy<TL> pol{S, S + N * TL, c};
    const long long rs = (long long)STEP * a.line_stride; /* elements between a thread's points */
    const long long toff = (long long)t * a.line_stride;
    float2 v[8], vn[8];
    int tile = blockIdx.x;
    /* prologue: first tile */
    {
        const int g = tile / tiles_per_group, col = (tile - g * tiles_per_group) * TL + c;
        const float2 *p = src + ((long long)g * a.group_stride + col + toff);
#pragma unroll
        for (int m = 0; m < 8; m++) {
            vn[m] = make_float2(0.f, 0.f);
            if (tile < ntiles && col < a.ncols) vn[m] = p[m * rs];
        }
    }
    if constexpr (MODE == PM_WTAB) {
        stage_window<N, TL>(Wst, a, tile, tiles_per_group, ntiles);
        cp_async_wait_all();
    }
    __syncthreads(); /* twiddle table (and the first tile's window rows) visible */
    for (int it = 0; tile < ntiles; tile += gridDim.x, it++) {
        const int g = tile / tiles_per_group, col = (tile - g * tiles_per_group) * TL + c;
        const bool live = col < a.ncols;
        const long long base = (long long)g * a.group_stride + col + toff;
#pragma unroll
        for (int m = 0; m < 8; m++) v[m] = vn[m];
        /* prefetch the next tile of this CTA */
        {
            const int nt = tile + gridDim.x;
            const int gn = nt / tiles_per_group, coln = (nt - gn * tiles_per_group) * TL + c;
            const float2 *p = src + ((long long)gn * a.group_stride + coln + toff);
            const bool liven = nt < ntiles && coln < a.ncols;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                vn[m] = make_float2(0.f, 0.f);
                if (liven) vn[m] = p[m * rs];
            }
        }
        if constexpr (MODE == PM_WTAB) {
            /* stage the window rows of the NEXT tile (landed before this tile's last barrier), then
               multiply by this tile's rows: |nx| = t + m STEP (m < 4) or N - (t + m STEP) */
            stage_window<N, TL>(Wst + ((it + 1) & 1) * WROWS * TL, a, tile + gridDim.x, tiles_per_group, ntiles);
            if (live) {
                const float *w = Wst + (it & 1) * WROWS * TL + c;
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    const int ax = m < 4 ? t + m * STEP : N - (t + m * STEP);
                    v[m] = pk_scale(v[m], w[ax * TL]);
                }
            }
        } else if constexpr (MODE == PM_GENERIC) {
            if (live) {
#pragma unroll
                for (int m = 0; m < 8; m++) v[m] = apply_kmul(v[m], t + m * STEP, col, a);
            }
        }
        fft_line_regs<N, SIGN, TilePolicy<TL>, MODE == PM_WTAB>(v, t, pol, twS);
        if (live) {
            if constexpr (SCATTER) {
                /* the transpose of the slab FFT: point t + m STEP of the line belongs to rank
                   (t + m STEP) / sc_nl; the TL lanes of a tile row write one contiguous segment
                   of that rank's receive buffer (local HBM or a peer's over NVLink) */
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    float2 *q = scatter_dst(a, t + m * STEP, g, col);
                    *q = (a.scale != 1.f) ? pk_scale(v[m], a.scale) : v[m];
                }
            } else {
                float2 *q = dst + base;
                if (a.scale != 1.f) {
#pragma unroll
                    for (int m = 0; m < 8; m++) q[m * rs] = pk_scale(v[m], a.scale);
                } else {
#pragma unroll
                    for (int m = 0; m < 8; m++) q[m * rs] = v[m];
                }
            }
        }
        if (Pow2Sync<N>::NEED_TAIL_SYNC) __syncthreads();
    }
}

template <int N, int SIGN, int MODE, bool SCATTER>
static void launch_strided_pow2_inst(const float2 *src, float2 *dst, const StridedArgs &a, int ngroups) {
    using P = Pow2Plan<N>;
    constexpr int TL = Pow2Cfg<N>::TL;
    const size_t smem = (size_t)2 * N * TL * sizeof(float2) + (size_t)P::TW_TOTAL * sizeof(float4) +
                        (MODE == PM_WTAB ? (size_t)2 * (N / 2 + 1) * TL * sizeof(float) : 0);
    const int tiles_per_group = (a.ncols + TL - 1) / TL;
    const long long ntiles = (long long)tiles_per_group * ngroups;
    auto k = &fft_strided_pow2_kernel<N, SIGN, MODE, SCATTER>;
    /* persistent grid: exactly the CTAs that are resident at once */
    allow_smem(k, smem);
    static int per_sm = 0;
    if (per_sm == 0) {
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, Pow2Cfg<N>::THREADS, smem));
        if (per_sm < 1) per_sm = 1;
    }
    long long grid = (long long)dev_num_sms() * per_sm;
    if (grid > ntiles) grid = ntiles;
    B200_LAUNCH_T(SCATTER ? "fft_strided_pow2_scatter_kernel" : "fft_strided_pow2_kernel", k, dim3((unsigned)grid),
                  Pow2Cfg<N>::THREADS, smem, src, dst, a, tiles_per_group, (int)ntiles);
}
template <int N, int MODE> static void launch_strided_pow2_mode(const float2 *src, float2 *dst, const StridedArgs &a, int ngroups) {
    if (a.sc_on) {
        /* scatter stores exist for the passes the slab transforms use: the forward y pass (plain)
           and the inverse x pass (any multiplier) */
        if (a.sign < 0) {
            if constexpr (MODE == PM_PLAIN) launch_strided_pow2_inst<N, -1, PM_PLAIN, true>(src, dst, a, ngroups);
            else b200_throw(B200_ValueError, "slab FFT: no forward scatter pass with a k-space multiplier");
        } else {
            launch_strided_pow2_inst<N, 1, MODE, true>(src, dst, a, ngroups);
        }
        return;
    }
    if (a.sign < 0) launch_strided_pow2_inst<N, -1, MODE, false>(src, dst, a, ngroups);
    else launch_strided_pow2_inst<N, 1, MODE, false>(src, dst, a, ngroups);
}
template <int N> static void launch_strided_pow2(const float2 *src, float2 *dst, const StridedArgs &a, int ngroups) {
    const bool has_kmul = a.kmul != KMUL_NONE || a.op != KOP_NONE;
    if (!has_kmul) launch_strided_pow2_mode<N, PM_PLAIN>(src, dst, a, ngroups);
    else if (a.wtab3 && a.op == KOP_NONE) launch_strided_pow2_mode<N, PM_WTAB>(src, dst, a, ngroups);
    else launch_strided_pow2_mode<N, PM_GENERIC>(src, dst, a, ngroups);
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
   inter-stage transposes use private padded strips of shared memory, and lines synchronise with
   __syncwarp / a per-line named barrier -- no CTA-wide barrier in the loop. */
template <int NH> struct ZLineCfg {
    static constexpr int LT = NH / 8;                       /* threads per line */
    static constexpr int THREADS = LT > 256 ? LT : 256;
    static constexpr int LINES = THREADS / LT;              /* lines per CTA */
    static constexpr int STRIP = NH + NH / 8;               /* padded points per line and buffer */
    static constexpr size_t SMEM = (size_t)2 * LINES * STRIP * sizeof(float2) + (size_t)Pow2Plan<NH>::TW_TOTAL * sizeof(float4) +
                                   (size_t)NH * sizeof(float2);
};
template <int NH>
__global__ void __launch_bounds__(ZLineCfg<NH>::THREADS)
fft_c2r_z_pow2_kernel(const float2 *__restrict__ src, float *__restrict__ dst, ZArgs a) {
    using Cfg = ZLineCfg<NH>;
    constexpr int LT = Cfg::LT, STEP = NH / 8;
    DYN_SMEM(float2, S);
    __shared__ float red_min[32], red_max[32];
    float4 *twS = reinterpret_cast<float4 *>(S + 2 * Cfg::LINES * Cfg::STRIP);
    float2 *mrg = reinterpret_cast<float2 *>(twS + Pow2Plan<NH>::TW_TOTAL); /* e^{+2 pi i n / NR}, n < NH */
    fill_twiddles<NH, 1>(twS, a.tw, 2, Cfg::THREADS);
    for (int n = threadIdx.x; n < NH; n += Cfg::THREADS) {
        float2 w = ldg(&a.tw[n]);
        mrg[n] = make_float2(w.x, -w.y);
    }
    __syncthreads();
    const int line = threadIdx.x / LT, t = threadIdx.x - line * LT;
    const LinePolicy<LT> pol{S + (2 * line) * Cfg::STRIP, S + (2 * line + 1) * Cfg::STRIP, 1 + line};
    RegTwiddles<NH> rtw;
    load_reg_twiddles<NH>(rtw, t, twS);
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
            const float2 B = make_float2(xb[m].x, -xb[m].y);
            const float2 sum = pk_add(xa[m], B), dif = pk_sub(xa[m], B);
            const float2 p = pk_cmul(dif, mrg[t + m * STEP]);
            v[m] = pk_add(sum, pk_rot<1>(p));
        }
        fft_line_regs_rt<NH, 1>(v, t, pol, rtw);
        /* z[n] = (x[2n], x[2n+1]) */
#pragma unroll
        for (int m = 0; m < 8; m++) {
            float2 x = pk_scale(v[m], a.scale);
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
        if (Pow2Sync<NH>::NEED_TAIL_SYNC) pol.sync();
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

/* NR = 2 NH reals per row -> NH + 1 complex values.  Same line-per-LT-threads layout: the real row
   is read as NH (even, odd) pairs, coalesced, straight into the FFT registers; the Hermitian
   split needs Z[k] and Z[NH-k] of the same line, exchanged through the line's strip. */
template <int NH>
__global__ void __launch_bounds__(ZLineCfg<NH>::THREADS)
fft_r2c_z_pow2_kernel(const float *__restrict__ src, float2 *__restrict__ dst, ZArgs a) {
    using Cfg = ZLineCfg<NH>;
    constexpr int LT = Cfg::LT, STEP = NH / 8;
    DYN_SMEM(float2, S);
    float4 *twS = reinterpret_cast<float4 *>(S + 2 * Cfg::LINES * Cfg::STRIP);
    float2 *mrg = reinterpret_cast<float2 *>(twS + Pow2Plan<NH>::TW_TOTAL); /* e^{-2 pi i k / NR}, k < NH */
    fill_twiddles<NH, -1>(twS, a.tw, 2, Cfg::THREADS);
    for (int n = threadIdx.x; n < NH; n += Cfg::THREADS) mrg[n] = ldg(&a.tw[n]);
    __syncthreads();
    const int line = threadIdx.x / LT, t = threadIdx.x - line * LT;
    const LinePolicy<LT> pol{S + (2 * line) * Cfg::STRIP, S + (2 * line + 1) * Cfg::STRIP, 1 + line};
    RegTwiddles<NH> rtw;
    load_reg_twiddles<NH>(rtw, t, twS);
    const float2 *src2 = reinterpret_cast<const float2 *>(src);
    const long long row_stride2 = a.real_row_stride / 2;
    const bool pre = a.premul != 1.f || a.clip;
    for (long long row0 = (long long)blockIdx.x * Cfg::LINES; row0 < a.nrows;
         row0 += (long long)gridDim.x * Cfg::LINES) {
        const long long row = row0 + line;
        const bool live = row < a.nrows;
        float2 v[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            float2 x = live ? src2[row * row_stride2 + t + m * STEP] : make_float2(0.f, 0.f);
            if (pre) {
                /* prepare_box_for_filtering (IonisationBox.c:343-346): double product, clamp */
                double c0 = (double)x.x * (double)a.premul, c1 = (double)x.y * (double)a.premul;
                if (a.clip) {
                    c0 = fmax(fmin(c0, (double)a.clip_hi), (double)a.clip_lo);
                    c1 = fmax(fmin(c1, (double)a.clip_hi), (double)a.clip_lo);
                }
                x = make_float2((float)c0, (float)c1);
            }
            v[m] = x;
        }
        fft_line_regs_rt<NH, -1>(v, t, pol, rtw);
        /* Z[k] at k = t + m STEP; partner Z[(NH - k) mod NH] through the strip that the last
           exchange did NOT use (no hazard with slower threads still reading the other one) */
        float2 *strip = (Pow2Plan<NH>::NSTAGE == 3) ? pol.A : pol.B;
#pragma unroll
        for (int m = 0; m < 8; m++) strip[pol.off(t + m * STEP)] = v[m];
        pol.sync();
        /* X[k] = (Z[k] + conj Z[NH-k]) / 2 - (i/2) e^{-2 pi i k / NR} (Z[k] - conj Z[NH-k]) */
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int k = t + m * STEP;
            const float2 zc = strip[pol.off((NH - k) & (NH - 1))];
            const float2 Zc = make_float2(zc.x, -zc.y);
            const float2 s = pk_add(v[m], Zc), d = pk_sub(v[m], Zc);
            const float2 p = pk_cmul(d, mrg[k]);
            float2 X = pk_scale(pk_add(s, pk_rot<-1>(p)), 0.5f * a.scale);
            if (live) dst[row * a.pitch + k] = X;
            if (k == 0 && live) dst[row * a.pitch + NH] = make_float2((v[m].x - v[m].y) * a.scale, 0.f);
        }
        pol.sync();
    }
}

template <int NH> static int z_grid(int per_sm, int nrows) {
    using Cfg = ZLineCfg<NH>;
    const long long groups = ((long long)nrows + Cfg::LINES - 1) / Cfg::LINES;
    long long grid = (long long)dev_num_sms() * per_sm;
    if (grid > groups) grid = groups;
    return (int)grid;
}
template <int NH> static void launch_c2r_z_pow2(const float2 *src, float *dst, const ZArgs &a) {
    using Cfg = ZLineCfg<NH>;
    auto k = &fft_c2r_z_pow2_kernel<NH>;
    allow_smem(k, Cfg::SMEM);
    static int per_sm = 0;
    if (per_sm == 0) {
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, Cfg::THREADS, Cfg::SMEM));
        if (per_sm < 1) per_sm = 1;
    }
    B200_LAUNCH_T("fft_c2r_z_pow2_kernel", k, dim3((unsigned)z_grid<NH>(per_sm, a.nrows)), Cfg::THREADS, Cfg::SMEM, src, dst, a);
}
template <int NH> static void launch_r2c_z_pow2(const float *src, float2 *dst, const ZArgs &a) {
    using Cfg = ZLineCfg<NH>;
    auto k = &fft_r2c_z_pow2_kernel<NH>;
    allow_smem(k, Cfg::SMEM);
    static int per_sm = 0;
    if (per_sm == 0) {
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, Cfg::THREADS, Cfg::SMEM));
        if (per_sm < 1) per_sm = 1;
    }
    B200_LAUNCH_T("fft_r2c_z_pow2_kernel", k, dim3((unsigned)z_grid<NH>(per_sm, a.nrows)), Cfg::THREADS, Cfg::SMEM, src, dst, a);
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
