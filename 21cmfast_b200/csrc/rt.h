/*
 * rt.h -- thin runtime layer shared by every translation unit of lib21cmfast_b200.so.
 *
 * Product build (nvcc, sm_100a): CUDA runtime, one library-owned stream, launch counting,
 * error -> status-code mapping.
 *
 * B200_EMU build (g++ only, used by tests/emu to check kernel *index logic* on a machine
 * without a GPU): the same kernel sources are compiled with the CUDA execution model mapped to
 * one logical thread per block (every kernel here is written as block-stride loops separated by
 * __syncthreads(), so this is exact) and blocks spread over OpenMP threads.  The emulation
 * library is test infrastructure: it is built into tests/_emu/, never into the package, and
 * Backend() never loads it.
 */
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>

#include "../../include/py21cmfast_b200.h"

/* status codes: exceptions.h:12-21 of the reference, plus 10 for CUDA failures */
enum {
    B200_OK = 0, B200_IOError = 1, B200_GSLError = 2, B200_ValueError = 3,
    B200_PhotonConsError = 4, B200_TableGenerationError = 5, B200_TableEvaluationError = 6,
    B200_InfinityorNaNError = 7, B200_MassDepZetaError = 8, B200_MemoryAllocError = 9,
    B200_CUDAError = 10
};

struct B200Error {
    int code;
    char msg[256];
};
[[noreturn]] void b200_throw(int code, const char *fmt, ...);

struct CallStats {
    long long launches, h2d, d2h;
    double ms;
};
extern CallStats g_stats;

#ifndef B200_EMU
/* ------------------------------------------------------------------ real CUDA */
#include <cuda_runtime.h>

extern cudaStream_t g_stream;
void rt_init();

#define CUDA_CHECK(expr)                                                                     \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            b200_throw(_e == cudaErrorMemoryAllocation ? B200_MemoryAllocError               \
                                                       : B200_CUDAError,                     \
                       "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)

/* optional per-kernel timing with CUDA events on the launching stream (b200_profile_enable) */
extern int g_profile;
int prof_begin(const char *name);
void prof_end(int slot);

#define B200_LAUNCH(kernel, grid, block, smem, ...)                  \
    do {                                                             \
        int _ps = g_profile ? prof_begin(#kernel) : -1;              \
        kernel<<<(grid), (block), (smem), g_stream>>>(__VA_ARGS__);  \
        if (_ps >= 0) prof_end(_ps);                                 \
        g_stats.launches++;                                          \
        CUDA_CHECK(cudaGetLastError());                              \
    } while (0)

/* same, for a kernel given as a function pointer (template instantiations) with an explicit label */
#define B200_LAUNCH_T(label, kptr, grid, block, smem, ...)           \
    do {                                                             \
        int _ps = g_profile ? prof_begin(label) : -1;                \
        kptr<<<(grid), (block), (smem), g_stream>>>(__VA_ARGS__);    \
        if (_ps >= 0) prof_end(_ps);                                 \
        g_stats.launches++;                                          \
        CUDA_CHECK(cudaGetLastError());                              \
    } while (0)

#define DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char _dyn_smem[]; \
    type *name = reinterpret_cast<type *>(_dyn_smem)

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
template <typename T> DEV T ldg(const T *p) { return __ldg(p); }
DEV void atomic_add_u64(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }
DEV unsigned int atomic_fetch_add_u32(unsigned int *p, unsigned int v) { return atomicAdd(p, v); }
DEV void atomic_or_u32(unsigned int *p, unsigned int v) { atomicOr(p, v); }
DEV void atomic_add_f64(double *p, double v) { atomicAdd(p, v); }
DEV void atomic_min_i32(int *p, int v) { atomicMin(p, v); }
DEV void atomic_max_i32(int *p, int v) { atomicMax(p, v); }
DEV int float_as_int_bits(float f) { return __float_as_int(f); }
DEV float int_bits_as_float(int i) { return __int_as_float(i); }
/* asynchronous 16-byte global -> shared copies (LDGSTS): in-flight loads that hold no registers */
DEV void cp_async_16(void *smem_dst, const void *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int N> DEV void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
/* two single-precision lanes per instruction (Blackwell FFMA2 / FADD2 / FMUL2) */
DEV float2 f2_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
DEV float2 f2_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
DEV float2 f2_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
DEV int float_to_int_floor(float x) { return __float2int_rd(x); }
#else /* host-only translation units (host_physics.cpp, host_numerics.cpp) built by g++ */
#define HD inline
#define DEV inline
#endif

#else
/* ------------------------------------------------------------------ host emulation */
#include <omp.h>
struct uint3e { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct uint2 { unsigned int x, y; };
struct uchar4 { unsigned char x, y, z, w; };
static inline uint2 make_uint2(unsigned int a, unsigned int b) { uint2 r = {a, b}; return r; }
static inline float2 make_float2(float a, float b) { float2 r = {a, b}; return r; }
static inline float4 make_float4(float a, float b, float c, float d) { float4 r = {a, b, c, d}; return r; }
extern thread_local uint3e blockIdx;
extern thread_local uint3e gridDim;
extern thread_local unsigned char *b200_emu_smem;
static const uint3e threadIdx = {0, 0, 0};
static const uint3e blockDim = {1, 1, 1};
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __syncthreads() ((void)0)
using std::isfinite;
#define HD inline
#define DEV inline
typedef int cudaStream_t;
static inline void rt_init() {}
template <typename T> inline T ldg(const T *p) { return *p; }
inline void atomic_add_u64(unsigned long long *p, unsigned long long v) {
#pragma omp atomic
    *p += v;
}
inline void atomic_add_f64(double *p, double v) {
#pragma omp atomic
    *p += v;
}
inline unsigned int atomic_fetch_add_u32(unsigned int *p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline void atomic_or_u32(unsigned int *p, unsigned int v) { __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline void atomic_min_i32(int *p, int v) {
#pragma omp critical(b200_minmax)
    { if (v < *p) *p = v; }
}
inline void atomic_max_i32(int *p, int v) {
#pragma omp critical(b200_minmax)
    { if (v > *p) *p = v; }
}
inline int float_as_int_bits(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float int_bits_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float2 f2_fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
inline float2 f2_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float2 f2_mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline int float_to_int_floor(float x) { return (int)floorf(x); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline void sincosf_emu(float x, float *s, float *c) { *s = sinf(x); *c = cosf(x); }
static inline void sincospi(double x, double *s, double *c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }

#define DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(b200_emu_smem)
#define B200_LAUNCH_T(label, kptr, grid, block, smem, ...) B200_LAUNCH(kptr, grid, block, smem, __VA_ARGS__)

#define B200_LAUNCH(kernel, grid, block, smem, ...)                                         \
    do {                                                                                    \
        dim3 _g = dim3(grid);                                                                    \
        long long _nb = (long long)_g.x * _g.y * _g.z;                                      \
        g_stats.launches++;                                                                 \
        _Pragma("omp parallel")                                                             \
        {                                                                                   \
            unsigned char *_sm = (unsigned char *)malloc((size_t)(smem) + 64);              \
            b200_emu_smem = _sm;                                                            \
            gridDim.x = _g.x; gridDim.y = _g.y; gridDim.z = _g.z;                           \
            _Pragma("omp for schedule(dynamic, 1)")                                         \
            for (long long _b = 0; _b < _nb; _b++) {                                        \
                blockIdx.x = (unsigned)(_b % _g.x);                                         \
                blockIdx.y = (unsigned)((_b / _g.x) % _g.y);                                \
                blockIdx.z = (unsigned)(_b / ((long long)_g.x * _g.y));                     \
                kernel(__VA_ARGS__);                                                        \
            }                                                                               \
            free(_sm);                                                                      \
        }                                                                                   \
    } while (0)
#endif

/* ------------------------------------------------------------------ memory helpers */
void *dev_alloc(size_t bytes);
void dev_free(void *p);
void dev_zero(void *p, size_t bytes);
void h2d(void *dst, const void *src, size_t bytes);
void d2h(void *dst, const void *src, size_t bytes);
void d2d(void *dst, const void *src, size_t bytes);
void dev_sync();
int dev_num_sms();
/* pinned host staging + asynchronous copies + events, for pipelines that must not stall the stream */
void *host_pinned_alloc(size_t bytes);
void host_pinned_free(void *p);
void h2d_async(void *dst, const void *pinned_src, size_t bytes);
void d2h_async(void *pinned_dst, const void *src, size_t bytes);
/* second stream for host<->device copies that overlap kernels on the main stream; `slot` indexes a
   small pool of events (copy_event_record on the copy stream, main_wait_copy_event makes the
   main stream wait; copy_wait_main makes the copy stream wait for everything enqueued so far
   on the main stream).  In the emulation build the copies are synchronous and the rest no-ops. */
void h2d_copy_stream(void *dst, const void *src, size_t bytes);
void d2h_copy_stream(void *dst, const void *src, size_t bytes);
void copy_event_record(int slot);
void main_wait_copy_event(int slot);
void copy_wait_main();
void copy_stream_sync();
/* auxiliary compute stream: small independent kernels (the per-radius window tables) run beside the
   main stream's big passes.  rt_use_aux(true) redirects B200_LAUNCH to the auxiliary stream until
   rt_use_aux(false); rt_event_record(slot) records on the CURRENT stream, rt_stream_wait(slot) makes
   the CURRENT stream wait for that record.  No-ops in the emulation build (everything is in order). */
void rt_use_aux(bool on);
void rt_event_record(int slot);
void rt_stream_wait(int slot);
void *dev_event_create();
void dev_event_record(void *ev);
void dev_event_wait_host(void *ev);
void dev_event_destroy(void *ev);

/* order-preserving float <-> int key (so that integer atomicMin/Max order floats) */
HD int float_order_key(int bits) { return bits >= 0 ? bits : bits ^ 0x7fffffff; }
inline float float_from_order_key(int key) {
    int bits = key >= 0 ? key : key ^ 0x7fffffff;
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* RAII device buffer */
template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        if (count) p = (T *)dev_alloc(count * sizeof(T));
        n = count;
    }
    void ensure(size_t count) { if (count > n) alloc(count); }
    void release() { if (p) dev_free(p); p = nullptr; n = 0; }
    operator T *() const { return p; }
};

/* Opt-in residency of OUTPUT boxes between the reference's own entry points: with
   b200_residency(1) a call that produces a box the next call of the chain reads (perturbed density -> ionized
   box -> brightness temperature) leaves its device copy behind under the HOST pointer the caller received it
   in, and the consumer uses that copy instead of uploading the host array again.  The caller promises not to
   modify or free those host arrays while residency is on (b200_residency(0) drops every copy); the last
   RESIDENT_SLOTS boxes are kept.  Off by default: the reference reads the caller's arrays on every call. */
#define RESIDENT_SLOTS 6
bool resident_enabled();
void resident_put(const void *host, float *dev_owned, size_t n); /* takes ownership of a dev_alloc'ed buffer */
const float *resident_get(const void *host, size_t n);            /* device mirror of the host array, or null */
void resident_clear();

/* device timer (CUDA events on g_stream; wall clock in emulation) */
struct DevTimer {
    void *a = nullptr, *b = nullptr;
    double t0 = 0;
    void start();
    double stop_ms();
};
