/* dist.cu -- symmetric peer memory, stream-ordered barrier and small gathers (see dist.h). */
#include "dist.h"

#ifdef B200_EMU
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#endif

DistCtx g_dist;

#define DIST_TIMEOUT_NS 30000000000ULL /* a barrier that waits longer than 30 s gives up and flags an error */

struct BarrierArgs {
    DistControl *ctl[DIST_MAX_RANKS];
    int rank, world;
    unsigned long long target;
    int mode; /* 0 barrier only, 1 gather 64-bit words, 2 combine {min key, max key}, 3 interleaved gather of 32-bit words */
    const unsigned long long *src[DIST_MAX_RANKS];
    unsigned long long *dst;
    int words;
};

#ifndef B200_EMU
DEV void sys_signal(unsigned long long *p) {
    __threadfence_system(); /* everything this GPU stored before (earlier kernels included) is ordered before the signal */
    atomicAdd_system(p, 1ULL);
}
DEV unsigned long long sys_load_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
DEV unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
DEV void spin_pause() { __nanosleep(64); }
DEV unsigned long long peer_load(const unsigned long long *p) { return __ldcv(p); }
DEV unsigned int peer_load32(const unsigned int *p) { return __ldcv(p); }
#else
inline void sys_signal(unsigned long long *p) { __atomic_fetch_add(p, 1ULL, __ATOMIC_SEQ_CST); }
inline unsigned long long sys_load_acquire(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline unsigned long long now_ns() { return (unsigned long long)(omp_get_wtime() * 1e9); }
inline void spin_pause() { sched_yield(); }
inline unsigned long long peer_load(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline unsigned int peer_load32(const unsigned int *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
#endif

/* one CTA: signal every rank (self included), wait until all `world` signals of this epoch have
   arrived here, then optionally fetch a few words from every peer */
__global__ void dist_barrier_kernel(BarrierArgs a) {
    for (int r = threadIdx.x; r < a.world; r += blockDim.x) sys_signal(&a.ctl[r]->arrive);
    if (threadIdx.x == 0) {
        const unsigned long long *mine = &a.ctl[a.rank]->arrive;
        const unsigned long long t0 = now_ns();
        while (sys_load_acquire(mine) < a.target) {
            spin_pause();
            if (now_ns() - t0 > DIST_TIMEOUT_NS) { a.ctl[a.rank]->error = 1; break; }
        }
    }
    __syncthreads();
    if (a.mode == 1) {
        for (int i = threadIdx.x; i < a.world * a.words; i += blockDim.x) {
            const int r = i / a.words, j = i - r * a.words;
            a.dst[i] = peer_load(a.src[r] + j);
        }
    } else if (a.mode == 3) {
        unsigned int *dst = reinterpret_cast<unsigned int *>(a.dst);
        for (int i = threadIdx.x; i < a.words; i += blockDim.x) {
            const int r = i % a.world, j = i / a.world;
            dst[i] = peer_load32(reinterpret_cast<const unsigned int *>(a.src[r]) + j);
        }
    } else if (a.mode == 2) {
        if (threadIdx.x == 0) {
            int lo = 2147483647, hi = -2147483647 - 1;
            for (int r = 0; r < a.world; r++) {
                const unsigned long long w = peer_load(a.src[r]);
                const int k0 = (int)(unsigned int)(w & 0xffffffffULL), k1 = (int)(unsigned int)(w >> 32);
                lo = k0 < lo ? k0 : lo;
                hi = k1 > hi ? k1 : hi;
            }
            int *out = reinterpret_cast<int *>(a.dst);
            out[0] = lo;
            out[1] = hi;
        }
    }
}

void dist_require() {
    if (!g_dist.ready) b200_throw(B200_ValueError, "the slab-decomposed entry points need b200_dist_init / b200_dist_connect first");
}

void *dist_alloc(size_t bytes) {
    dist_require();
    const size_t off = (g_dist.bump + 255) & ~(size_t)255;
    if (off + bytes > g_dist.heap_bytes)
        b200_throw(B200_MemoryAllocError, "symmetric heap exhausted: need %zu more bytes (heap %zu); pass a larger heap to b200_dist_init",
                   off + bytes - g_dist.heap_bytes, g_dist.heap_bytes);
    g_dist.bump = off + bytes;
    return g_dist.heap[g_dist.rank] + off;
}
void dist_reset() { g_dist.bump = DIST_CONTROL_BYTES; }
size_t dist_mark() { return g_dist.bump; }
void dist_release(size_t mark) { g_dist.bump = mark; }

static void launch_barrier(int mode, const unsigned long long *src_sym, unsigned long long *dst, int words) {
    dist_require();
    BarrierArgs a;
    memset(&a, 0, sizeof(a));
    a.rank = g_dist.rank; a.world = g_dist.world;
    g_dist.epoch++;
    a.target = g_dist.epoch * (unsigned long long)g_dist.world;
    a.mode = mode; a.dst = dst; a.words = words;
    for (int r = 0; r < g_dist.world; r++) {
        a.ctl[r] = reinterpret_cast<DistControl *>(g_dist.heap[r]);
        a.src[r] = src_sym ? dist_peer(const_cast<unsigned long long *>(src_sym), r) : nullptr;
    }
    B200_LAUNCH(dist_barrier_kernel, 1, 128, 0, a);
}
void dist_barrier() { launch_barrier(0, nullptr, nullptr, 0); }
void dist_barrier_gather(const unsigned long long *src_sym, unsigned long long *dst, int words) {
    launch_barrier(1, src_sym, dst, words);
}
void dist_barrier_gather_interleaved32(const unsigned int *src_sym, unsigned int *dst, int total_words) {
    launch_barrier(3, reinterpret_cast<const unsigned long long *>(src_sym), reinterpret_cast<unsigned long long *>(dst), total_words);
}
void dist_barrier_minmax(const int *keys_sym, int *out) {
    launch_barrier(2, reinterpret_cast<const unsigned long long *>(keys_sym), reinterpret_cast<unsigned long long *>(out), 1);
}

void dist_check() {
    if (!g_dist.ready) return;
    dev_sync();
    int err = 0;
    DistControl *ctl = reinterpret_cast<DistControl *>(g_dist.heap[g_dist.rank]);
#ifndef B200_EMU
    CUDA_CHECK(cudaMemcpy(&err, &ctl->error, sizeof(int), cudaMemcpyDeviceToHost));
#else
    err = ctl->error;
#endif
    if (err) b200_throw(B200_CUDAError, "a cross-GPU barrier timed out (a peer rank failed or left the call sequence)");
}

/* ------------------------------------------------------------------ set-up / tear-down (C ABI) */
#ifndef B200_EMU
extern "C" int b200_dist_shutdown(void) {
    if (g_dist.heap[g_dist.rank] == nullptr) return 0;
    if (g_stream) cudaStreamSynchronize(g_stream);
    for (int r = 0; r < g_dist.world; r++) {
        if (!g_dist.heap[r]) continue;
        if (r == g_dist.rank) cudaFree(g_dist.heap[r]);
        else cudaIpcCloseMemHandle(g_dist.heap[r]);
        g_dist.heap[r] = nullptr;
    }
    g_dist = DistCtx();
    return 0;
}
extern "C" int b200_dist_init(int rank, int world, unsigned long long heap_bytes, void *handle_out) {
    try {
        rt_init();
        static_assert(sizeof(cudaIpcMemHandle_t) == DIST_HANDLE_BYTES, "handle size");
        if (world < 1 || world > DIST_MAX_RANKS || rank < 0 || rank >= world || heap_bytes < 2 * DIST_CONTROL_BYTES || !handle_out)
            b200_throw(B200_ValueError, "b200_dist_init: bad arguments (1 <= world <= %d)", DIST_MAX_RANKS);
        b200_dist_shutdown();
        void *p = nullptr;
        CUDA_CHECK(cudaMalloc(&p, heap_bytes));
        CUDA_CHECK(cudaMemset(p, 0, DIST_CONTROL_BYTES));
        CUDA_CHECK(cudaDeviceSynchronize());
        cudaIpcMemHandle_t h;
        CUDA_CHECK(cudaIpcGetMemHandle(&h, p));
        memcpy(handle_out, &h, DIST_HANDLE_BYTES);
        g_dist.rank = rank; g_dist.world = world; g_dist.heap_bytes = heap_bytes;
        g_dist.heap[rank] = (unsigned char *)p;
        g_dist.bump = DIST_CONTROL_BYTES;
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] b200_dist_init: %s\n", e.msg);
        return e.code;
    }
    return 0;
}
extern "C" int b200_dist_connect(const void *handles) {
    try {
        if (!g_dist.heap[g_dist.rank] || !handles) b200_throw(B200_ValueError, "b200_dist_connect before b200_dist_init");
        for (int r = 0; r < g_dist.world; r++) {
            if (r == g_dist.rank) continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, (const unsigned char *)handles + (size_t)r * DIST_HANDLE_BYTES, DIST_HANDLE_BYTES);
            void *p = nullptr;
            CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            g_dist.heap[r] = (unsigned char *)p;
        }
        g_dist.ready = true;
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] b200_dist_connect: %s\n", e.msg);
        return e.code;
    }
    return 0;
}
#else
static char g_shm_name[DIST_MAX_RANKS][DIST_HANDLE_BYTES];
extern "C" int b200_dist_shutdown(void) {
    for (int r = 0; r < g_dist.world; r++) {
        if (!g_dist.heap[r]) continue;
        munmap(g_dist.heap[r], g_dist.heap_bytes);
        if (r == g_dist.rank) shm_unlink(g_shm_name[r]);
        g_dist.heap[r] = nullptr;
    }
    g_dist = DistCtx();
    return 0;
}
extern "C" int b200_dist_init(int rank, int world, unsigned long long heap_bytes, void *handle_out) {
    try {
        if (world < 1 || world > DIST_MAX_RANKS || rank < 0 || rank >= world || heap_bytes < 2 * DIST_CONTROL_BYTES || !handle_out)
            b200_throw(B200_ValueError, "b200_dist_init: bad arguments");
        b200_dist_shutdown();
        static int counter = 0;
        snprintf(g_shm_name[rank], DIST_HANDLE_BYTES, "/b200emu_%d_%d_%d", (int)getpid(), rank, counter++);
        const int fd = shm_open(g_shm_name[rank], O_CREAT | O_RDWR | O_EXCL, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)heap_bytes) != 0) b200_throw(B200_MemoryAllocError, "shm_open / ftruncate failed");
        void *p = mmap(nullptr, heap_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (p == MAP_FAILED) b200_throw(B200_MemoryAllocError, "mmap of the symmetric heap failed");
        memset(p, 0, DIST_CONTROL_BYTES);
        memset(handle_out, 0, DIST_HANDLE_BYTES);
        memcpy(handle_out, g_shm_name[rank], strlen(g_shm_name[rank]));
        g_dist.rank = rank; g_dist.world = world; g_dist.heap_bytes = heap_bytes;
        g_dist.heap[rank] = (unsigned char *)p;
        g_dist.bump = DIST_CONTROL_BYTES;
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] b200_dist_init: %s\n", e.msg);
        return e.code;
    }
    return 0;
}
extern "C" int b200_dist_connect(const void *handles) {
    try {
        if (!g_dist.heap[g_dist.rank] || !handles) b200_throw(B200_ValueError, "b200_dist_connect before b200_dist_init");
        for (int r = 0; r < g_dist.world; r++) {
            if (r == g_dist.rank) continue;
            char name[DIST_HANDLE_BYTES + 1];
            memcpy(name, (const unsigned char *)handles + (size_t)r * DIST_HANDLE_BYTES, DIST_HANDLE_BYTES);
            name[DIST_HANDLE_BYTES] = 0;
            const int fd = shm_open(name, O_RDWR, 0600);
            if (fd < 0) b200_throw(B200_MemoryAllocError, "cannot open the peer heap %s", name);
            void *p = mmap(nullptr, g_dist.heap_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            close(fd);
            if (p == MAP_FAILED) b200_throw(B200_MemoryAllocError, "mmap of a peer heap failed");
            g_dist.heap[r] = (unsigned char *)p;
        }
        g_dist.ready = true;
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] b200_dist_connect: %s\n", e.msg);
        return e.code;
    }
    return 0;
}
#endif

extern "C" int b200_dist_rank(void) { return g_dist.ready ? g_dist.rank : -1; }
extern "C" int b200_dist_world(void) { return g_dist.ready ? g_dist.world : 0; }
/* stand-alone barrier for callers that interleave their own work with the library's (tests) */
extern "C" int b200_dist_barrier(void) {
    try {
        dist_barrier();
        dist_check();
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] b200_dist_barrier: %s\n", e.msg);
        return e.code;
    }
    return 0;
}
