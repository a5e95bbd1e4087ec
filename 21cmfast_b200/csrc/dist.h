/*
 * dist.h -- one coeval box across the GPUs of a node: symmetric peer memory, barriers and small
 * gathers issued from the library's own stream (slab-decomposed FFT with the
 * transpose over NVLink, slab-local particle deposit with a halo exchange).
 *
 * One process per GPU.  Every rank owns a "symmetric heap" (one cudaMalloc, exported with a CUDA IPC
 * handle and mapped by every peer), carved by a bump allocator that all ranks drive with the same
 * sequence of requests, so an object lives at the same offset on every rank and
 * dist_peer(ptr, r) is plain pointer arithmetic.  Kernels store straight into a peer's heap over
 * NVLink (the all-to-all transpose of the slab FFT is the store stage of a 1-D pass, see
 * fft.cu::fft_c2r_slab) and a one-CTA barrier kernel on the same stream orders those stores
 * against the peer's next read: system-scope release increments of every rank's arrival counter,
 * acquire spin on the own one.  Nothing here calls NCCL; the Python side only uses
 * torch.distributed to exchange the 64-byte handles once.
 *
 * B200_EMU build (CPU test tier): the heap is a POSIX shared-memory segment, the handle its name,
 * the barrier the same counters with __atomic builtins -- two gloo ranks exercise the real index
 * logic and the real concurrency on a machine without a GPU.
 */
#pragma once
#include "rt.h"

#define DIST_MAX_RANKS 8
#define DIST_HANDLE_BYTES 64

struct DistCtx {
    bool ready = false;
    int rank = 0, world = 1;
    size_t heap_bytes = 0;
    unsigned char *heap[DIST_MAX_RANKS] = {nullptr}; /* heap[r]: rank r's heap as mapped in this process */
    size_t bump = 0;                                 /* next free offset (after the control block) */
    unsigned long long epoch = 0;                    /* barriers issued so far (same on every rank) */
    unsigned long long transforms = 0;               /* slab transforms issued so far: selects the receive buffer */
};
extern DistCtx g_dist;

/* control block at the start of every heap */
struct DistControl {
    unsigned long long arrive;   /* arrival counter of the barrier (monotonic) */
    unsigned long long pad0[15];
    int error;                   /* set by a barrier that timed out */
    int pad1[31];
};
#define DIST_CONTROL_BYTES 4096

void dist_require();
/* symmetric allocation (256-byte aligned); valid until the next dist_reset() */
void *dist_alloc(size_t bytes);
void dist_reset();
size_t dist_mark();               /* current bump offset: dist_release(mark) frees everything allocated after it */
void dist_release(size_t mark);
template <typename T> static inline T *dist_peer(T *local, int r) {
    return reinterpret_cast<T *>(g_dist.heap[r] + (reinterpret_cast<unsigned char *>(local) - g_dist.heap[g_dist.rank]));
}
/* stream-ordered barrier over all ranks: every store (local or to a peer) issued on the library's
   stream before it is visible to every rank's work issued after it */
void dist_barrier();
/* barrier + gather of `words` 64-bit words per rank: dst[r * words + j] = (rank r's src)[j].
   src must live in the symmetric heap, dst is local memory. */
void dist_barrier_gather(const unsigned long long *src_sym, unsigned long long *dst, int words);
/* barrier + interleaved gather of 32-bit words: dst[i] = (rank i % world's src)[i / world] for i < total_words --
   a table whose entries were computed round-robin over the ranks (the per-radius f_coll tables of the ladder) */
void dist_barrier_gather_interleaved32(const unsigned int *src_sym, unsigned int *dst, int total_words);
/* barrier + combine of {min key, max key} pairs (float_order_key integers): keys_sym (symmetric,
   int[2]) holds this rank's pair; out[0] = min over ranks, out[1] = max over ranks (local memory) */
void dist_barrier_minmax(const int *keys_sym, int *out);
/* throws if a barrier of this call timed out (checked after the stream has been drained) */
void dist_check();
