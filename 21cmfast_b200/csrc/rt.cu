/* rt.cu -- stream, error plumbing, pooled device allocations, copy accounting. */
#include "rt.h"

#include <cstdarg>
#include <map>
#include <vector>

CallStats g_stats = {0, 0, 0, 0.0};

void b200_throw(int code, const char *fmt, ...) {
    B200Error e;
    e.code = code;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(e.msg, sizeof(e.msg), fmt, ap);
    va_end(ap);
    if (getenv("B200_VERBOSE")) fprintf(stderr, "[21cmfast_b200] error %d: %s\n", code, e.msg);
    throw e;
}

/* Freed buffers are kept and handed back for the next request of the same size: a Compute* call
   allocates the same few multi-GB work boxes every time, and cudaMalloc/cudaFree of those costs
   milliseconds and synchronises the device. */
static std::multimap<size_t, void *> g_free_pool;
static std::map<void *, size_t> g_live;

#ifndef B200_EMU
cudaStream_t g_stream = nullptr;
static cudaStream_t g_copy_stream = nullptr;
static cudaStream_t g_main_stream = nullptr, g_aux_stream = nullptr;
static cudaEvent_t g_rt_events[128];
static bool g_rt_events_ready = false;
static cudaEvent_t g_copy_events[64];
static cudaEvent_t g_main_event = nullptr;
static int g_device = -1;
static int g_sms = 0;

void rt_init() {
    if (g_stream) return;
    if (g_device < 0) {
        int dev = 0;
        const char *lr = getenv("B200_DEVICE");
        if (lr) dev = atoi(lr);
        g_device = dev;
    }
    CUDA_CHECK(cudaSetDevice(g_device));
    CUDA_CHECK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, g_device));
    g_sms = prop.multiProcessorCount;
}
int dev_num_sms() { rt_init(); return g_sms; }

extern "C" int b200_set_device(int device) {
    try {
        if (g_stream && device != g_device) {
            b200_release_device_cache();
            cudaStreamDestroy(g_stream);
            g_stream = nullptr;
            if (g_rt_events_ready) {
                cudaStreamDestroy(g_aux_stream);
                g_aux_stream = nullptr;
                for (int i = 0; i < 128; i++) cudaEventDestroy(g_rt_events[i]);
                g_rt_events_ready = false;
            }
            if (g_copy_stream) {
                cudaStreamDestroy(g_copy_stream);
                g_copy_stream = nullptr;
                for (int i = 0; i < 64; i++) cudaEventDestroy(g_copy_events[i]);
                cudaEventDestroy(g_main_event);
            }
        }
        g_device = device;
        rt_init();
    } catch (B200Error &e) { return e.code; }
    return 0;
}

void *dev_alloc(size_t bytes) {
    rt_init();
    auto it = g_free_pool.find(bytes);
    void *p = nullptr;
    if (it != g_free_pool.end()) {
        p = it->second;
        g_free_pool.erase(it);
    } else {
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) { /* retry once after dropping the pool */
            cudaGetLastError();
            for (auto &kv : g_free_pool) cudaFree(kv.second);
            g_free_pool.clear();
            CUDA_CHECK(cudaMalloc(&p, bytes));
        }
    }
    g_live[p] = bytes;
    return p;
}
void dev_free(void *p) {
    auto it = g_live.find(p);
    if (it == g_live.end()) return;
    g_free_pool.insert({it->second, p});
    g_live.erase(it);
}
static void pool_drop() {
    if (g_stream) cudaStreamSynchronize(g_stream);
    for (auto &kv : g_free_pool) cudaFree(kv.second);
    g_free_pool.clear();
}
void dev_zero(void *p, size_t bytes) { CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, g_stream)); }
void h2d(void *dst, const void *src, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream));
    g_stats.h2d += (long long)bytes;
}
void d2h(void *dst, const void *src, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
    CUDA_CHECK(cudaStreamSynchronize(g_stream));
    g_stats.d2h += (long long)bytes;
}
void d2d(void *dst, const void *src, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_stream));
}
void dev_sync() { CUDA_CHECK(cudaStreamSynchronize(g_stream)); }
void *host_pinned_alloc(size_t bytes) {
    void *p = nullptr;
    CUDA_CHECK(cudaMallocHost(&p, bytes));
    return p;
}
void host_pinned_free(void *p) { if (p) cudaFreeHost(p); }
void h2d_async(void *dst, const void *src, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream));
}
void d2h_async(void *dst, const void *src, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
}
static void copy_stream_init() {
    rt_init();
    if (g_copy_stream) return;
    CUDA_CHECK(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 64; i++) CUDA_CHECK(cudaEventCreateWithFlags(&g_copy_events[i], cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&g_main_event, cudaEventDisableTiming));
}
void h2d_copy_stream(void *dst, const void *src, size_t bytes) {
    copy_stream_init();
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_copy_stream));
    g_stats.h2d += (long long)bytes;
}
void d2h_copy_stream(void *dst, const void *src, size_t bytes) {
    copy_stream_init();
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_copy_stream));
    g_stats.d2h += (long long)bytes;
}
void copy_event_record(int slot) {
    copy_stream_init();
    CUDA_CHECK(cudaEventRecord(g_copy_events[slot & 63], g_copy_stream));
}
void main_wait_copy_event(int slot) { CUDA_CHECK(cudaStreamWaitEvent(g_stream, g_copy_events[slot & 63], 0)); }
void copy_wait_main() {
    copy_stream_init();
    CUDA_CHECK(cudaEventRecord(g_main_event, g_stream));
    CUDA_CHECK(cudaStreamWaitEvent(g_copy_stream, g_main_event, 0));
}
void copy_stream_sync() { if (g_copy_stream) CUDA_CHECK(cudaStreamSynchronize(g_copy_stream)); }
static void aux_init() {
    rt_init();
    if (g_rt_events_ready) return;
    CUDA_CHECK(cudaStreamCreateWithFlags(&g_aux_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 128; i++) CUDA_CHECK(cudaEventCreateWithFlags(&g_rt_events[i], cudaEventDisableTiming));
    g_rt_events_ready = true;
}
void rt_use_aux(bool on) {
    aux_init();
    if (on) {
        if (g_stream != g_aux_stream) { g_main_stream = g_stream; g_stream = g_aux_stream; }
    } else if (g_stream == g_aux_stream) {
        g_stream = g_main_stream;
    }
}
void rt_event_record(int slot) {
    aux_init();
    CUDA_CHECK(cudaEventRecord(g_rt_events[slot & 127], g_stream));
}
void rt_stream_wait(int slot) {
    aux_init();
    CUDA_CHECK(cudaStreamWaitEvent(g_stream, g_rt_events[slot & 127], 0));
}
void *dev_event_create() {
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return e;
}
void dev_event_record(void *ev) { CUDA_CHECK(cudaEventRecord((cudaEvent_t)ev, g_stream)); }
void dev_event_wait_host(void *ev) { CUDA_CHECK(cudaEventSynchronize((cudaEvent_t)ev)); }
void dev_event_destroy(void *ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }

void DevTimer::start() {
    rt_init();
    if (!a) {
        cudaEvent_t ea, eb;
        CUDA_CHECK(cudaEventCreate(&ea));
        CUDA_CHECK(cudaEventCreate(&eb));
        a = ea; b = eb;
    }
    CUDA_CHECK(cudaEventRecord((cudaEvent_t)a, g_stream));
}
double DevTimer::stop_ms() {
    CUDA_CHECK(cudaEventRecord((cudaEvent_t)b, g_stream));
    CUDA_CHECK(cudaEventSynchronize((cudaEvent_t)b));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b));
    return ms;
}

/* ---- per-kernel event timing ------------------------------------------------------------ */
#include <string>
int g_profile = 0;
struct ProfPair { std::string name; cudaEvent_t a, b; };
static std::vector<ProfPair> g_prof_pairs;
static std::map<std::string, std::pair<long long, double>> g_prof_acc;
int prof_begin(const char *name) {
    ProfPair p;
    p.name = name;
    if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return -1;
    cudaEventRecord(p.a, g_stream);
    g_prof_pairs.push_back(p);
    return (int)g_prof_pairs.size() - 1;
}
void prof_end(int slot) { cudaEventRecord(g_prof_pairs[slot].b, g_stream); }
static void prof_collect() {
    if (g_stream) cudaStreamSynchronize(g_stream);
    for (auto &p : g_prof_pairs) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            auto &acc = g_prof_acc[p.name];
            acc.first++;
            acc.second += ms;
        }
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    g_prof_pairs.clear();
}
extern "C" void b200_profile_enable(int on) {
    prof_collect();
    if (on) g_prof_acc.clear();
    g_profile = on;
}
/* writes "name count total_ms\n" lines; returns the number of bytes needed */
extern "C" int b200_profile_report(char *buf, int buflen) {
    prof_collect();
    std::string out;
    char line[256];
    for (auto &kv : g_prof_acc) {
        snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (buf && buflen > 0) {
        strncpy(buf, out.c_str(), buflen - 1);
        buf[buflen - 1] = 0;
    }
    return (int)out.size() + 1;
}
#else
/* ---------------------------------------------------------------- emulation */
extern "C" void b200_profile_enable(int) {}
extern "C" int b200_profile_report(char *buf, int buflen) { if (buf && buflen > 0) buf[0] = 0; return 1; }
thread_local uint3e blockIdx = {0, 0, 0};
thread_local uint3e gridDim = {1, 1, 1};
thread_local unsigned char *b200_emu_smem = nullptr;
int dev_num_sms() { return 4; }
extern "C" int b200_set_device(int) { return 0; }
void *dev_alloc(size_t bytes) {
    void *p = nullptr;
    if (posix_memalign(&p, 256, bytes ? bytes : 256)) b200_throw(B200_MemoryAllocError, "malloc");
    return p;
}
void dev_free(void *p) { free(p); }
static void pool_drop() {}
void dev_zero(void *p, size_t bytes) { memset(p, 0, bytes); }
void h2d(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); g_stats.h2d += bytes; }
void d2h(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); g_stats.d2h += bytes; }
void d2d(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
void dev_sync() {}
void *host_pinned_alloc(size_t bytes) { return malloc(bytes); }
void host_pinned_free(void *p) { free(p); }
void h2d_async(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
void d2h_async(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
void h2d_copy_stream(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); g_stats.h2d += bytes; }
void d2h_copy_stream(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); g_stats.d2h += bytes; }
void copy_event_record(int) {}
void main_wait_copy_event(int) {}
void copy_wait_main() {}
void copy_stream_sync() {}
void rt_use_aux(bool) {}
void rt_event_record(int) {}
void rt_stream_wait(int) {}
void *dev_event_create() { return nullptr; }
void dev_event_record(void *) {}
void dev_event_wait_host(void *) {}
void dev_event_destroy(void *) {}
void DevTimer::start() { t0 = omp_get_wtime(); }
double DevTimer::stop_ms() { return (omp_get_wtime() - t0) * 1e3; }
#endif

void ics_cache_drop();  /* perturb.cu */
void fft_plans_drop();  /* fft.cu */
void ps_device_tables_drop(); /* host_physics.cpp: device copy of the CLASS spline nodes */

/* ------------------------------------------------------------------ output residency (rt.h) */
struct ResidentEntry { const void *host; float *dev; size_t n; unsigned long long stamp; };
static ResidentEntry g_resident[RESIDENT_SLOTS];
static unsigned long long g_resident_clock = 0;
static bool g_resident_on = false;
bool resident_enabled() { return g_resident_on; }
void resident_clear() {
    for (auto &e : g_resident) {
        if (e.dev) dev_free(e.dev);
        e = ResidentEntry{nullptr, nullptr, 0, 0};
    }
}
void resident_put(const void *host, float *dev_owned, size_t n) {
    if (!g_resident_on || !host) { dev_free(dev_owned); return; }
    ResidentEntry *slot = nullptr;
    for (auto &e : g_resident)
        if (e.host == host) slot = &e; /* the same host array produced again: replace its copy */
    if (!slot) {
        slot = &g_resident[0];
        for (auto &e : g_resident)
            if (e.stamp < slot->stamp) slot = &e; /* empty slots have stamp 0: used first, then the oldest */
    }
    if (slot->dev) dev_free(slot->dev);
    *slot = ResidentEntry{host, dev_owned, n, ++g_resident_clock};
}
const float *resident_get(const void *host, size_t n) {
    if (!g_resident_on || !host) return nullptr;
    for (auto &e : g_resident)
        if (e.host == host && e.n == n && e.dev) { e.stamp = ++g_resident_clock; return e.dev; }
    return nullptr;
}
extern "C" void b200_residency(int enable) {
    g_resident_on = enable != 0;
    if (!g_resident_on) resident_clear();
}

extern "C" void b200_release_device_cache(void) {
    resident_clear();
    ps_device_tables_drop();
    ics_cache_drop();
    fft_plans_drop();
    pool_drop();
}

extern "C" void b200_last_call_stats(long long *launches, long long *h2d_b, long long *d2h_b,
                                     double *ms) {
    if (launches) *launches = g_stats.launches;
    if (h2d_b) *h2d_b = g_stats.h2d;
    if (d2h_b) *d2h_b = g_stats.d2h;
    if (ms) *ms = g_stats.ms;
}
