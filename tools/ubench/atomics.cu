// micro-benchmark: throughput of the atomic flavours the CIC deposit could use (B200)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
constexpr int TILE = 4096; // 3 x 16 KB of u32 limbs = 48 KB static
template <int MODE> __global__ void __launch_bounds__(256) k(unsigned long long *g, size_t gmask, int iters) {
    __shared__ unsigned s32[TILE * 3];
    unsigned long long *s64 = reinterpret_cast<unsigned long long *>(s32); // TILE u64 cells fit in the first 2 limb planes
    float *sf = reinterpret_cast<float *>(s32);
    for (int i = threadIdx.x; i < TILE * 3; i += 256) s32[i] = 0;
    __syncthreads();
    unsigned base = hash(blockIdx.x * 977 + 13);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int it = 0; it < iters; it++) {
        // neighbouring lanes hit neighbouring cells (like a z-run of groups), random row per warp-iter
        unsigned r = hash(base + it * 8 + warp);
        unsigned cell = (r % (TILE - 64)) + lane;
        unsigned long long v = (unsigned long long)(r | 1) << 13;
        if (MODE == 0) atomicAdd(&g[((size_t)r * 32 + lane) & gmask], v);                 // global u64 red, spread
        if (MODE == 1) { atomicAdd(&s32[cell], (unsigned)(v & 0xfffff)); atomicAdd(&s32[TILE + cell], (unsigned)((v >> 20) & 0xfffff)); atomicAdd(&s32[2 * TILE + cell], (unsigned)(v >> 40)); } // 3 limbs
        if (MODE == 2) atomicAdd(&s64[cell], v);                                           // smem u64 (CAS loop)
        if (MODE == 3) atomicAdd(&sf[cell], 1.0f);                                         // smem f32
        if (MODE == 4) atomicAdd(&s32[cell], (unsigned)v);                                 // smem u32 single
        if (MODE == 5) { unsigned old = atomicAdd(&s32[cell], (unsigned)v); if (old + (unsigned)v < old) atomicAdd(&s32[TILE + cell], 1u + (unsigned)(v >> 32)); else atomicAdd(&s32[TILE + cell], (unsigned)(v >> 32)); } // lo/hi carry
    }
    __syncthreads();
    if (MODE != 0) { unsigned long long a = 0; for (int i = threadIdx.x; i < TILE * 3; i += 256) a += s32[i]; if (a == 0x123456789ULL) g[0] = a; }
}
template <int MODE> int run(const char *name, unsigned long long *g, size_t gmask) {
    int iters = 2000, grid = 148 * 4;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<grid, 256>>>(g, gmask, 100);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    k<MODE><<<grid, 256>>>(g, gmask, iters);
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    double n = (double)grid * 256 * iters;
    printf("%-28s %8.3f ms  %8.1f G contributions/s\n", name, ms, n / ms * 1e-6);
    return 0;
}
int main() {
    size_t n = (size_t)1 << 27; // 1 GiB of u64: 512^3 grid
    unsigned long long *g; CK(cudaMalloc(&g, n * 8)); CK(cudaMemset(g, 0, n * 8));
    run<0>("global red.u64 spread", g, n - 1);
    run<1>("smem 3x red.u32 (20-bit limbs)", g, n - 1);
    run<2>("smem atomicAdd u64 (CAS)", g, n - 1);
    run<3>("smem atomicAdd f32", g, n - 1);
    run<4>("smem red.u32 single", g, n - 1);
    run<5>("smem lo/hi with carry", g, n - 1);
    return 0;
}
