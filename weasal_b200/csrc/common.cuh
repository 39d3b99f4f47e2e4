// Shared helpers for the weasal_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include "../../include/weasal_b200.h"
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <string>

namespace kp {

// ---- error plumbing: every C-ABI entry returns a status; kp_last_error() holds the message -----------------------
extern thread_local std::string g_last_error;

// status codes: enum kp_status of the public header (KP_OK, KP_ERR_*)

inline int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define KP_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return ::kp::fail(KP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// every kernel launch is followed by this: it counts the launch (kp_launch_count) and surfaces launch errors
extern std::atomic<long long> g_launch_count;
#define KP_CHECK_LAUNCH()                     \
    do {                                      \
        ::kp::g_launch_count.fetch_add(1);    \
        KP_CUDA(cudaGetLastError());          \
    } while (0)

// ---- optional per-kernel timing (bench.py's roofline leg): CUDA events recorded on the launching stream ---------------
extern bool g_profile_on;
void profile_push(const char* tag, cudaEvent_t a, cudaEvent_t b);
struct ProfileScope {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s;
    const char* tag;
    ProfileScope(const char* t, cudaStream_t st) : s(st), tag(t) {
        if (g_profile_on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s); }
    }
    ~ProfileScope() {
        if (a) { cudaEventRecord(b, s); profile_push(tag, a, b); }
    }
};

// ---- scratch memory ----------------------------------------------------------------------------------------------------
// Every entry point carves its temporaries out of a per-(thread, stream) arena that persists across calls: calls on
// one stream execute in order, so the next call may reuse the bytes of the previous one without any allocator
// traffic (a training step makes ~400 temporaries). The arena grows by adding geometrically larger blocks (first one
// 128 MB) and never returns memory, so steady state is a bump pointer over one or two blocks.
struct ArenaBlock { char* base; size_t cap; };
struct Arena {
    ArenaBlock blocks[16];
    int nblocks = 0;
    int cur = 0;
    size_t off = 0;
    int hold = 0;  // > 0: nested entry points append to the arena instead of restarting it
};
Arena* arena_for(cudaStream_t stream);
int arena_begin(Arena* a, cudaStream_t stream);
void* arena_alloc(Arena* a, size_t bytes);

struct Scratch {
    cudaStream_t stream;
    Arena* arena;
    int status = KP_OK;
    explicit Scratch(cudaStream_t s) : stream(s) {
        arena = arena_for(s);
        if (arena->hold == 0) status = arena_begin(arena, s);  // a held arena keeps its contents (see ArenaHold)
    }
    template <typename T>
    T* alloc(size_t count) {
        if (status != KP_OK) return nullptr;
        void* p = arena_alloc(arena, (count ? count : 1) * sizeof(T));
        if (!p) {
            status = fail(KP_ERR_CUDA, "scratch arena: cudaMalloc failed");
            return nullptr;
        }
        return reinterpret_cast<T*>(p);
    }
};

// While alive, entry points called from inside another entry point keep what the outer one allocated.
struct ArenaHold {
    Arena* a;
    explicit ArenaHold(Scratch& s) : a(s.arena) { a->hold++; }
    ~ArenaHold() { a->hold--; }
};

// small host arrays (batch offsets, grid rotations) travel to the device as a kernel argument: no pageable
// host->device copy, which costs ~25 us of host time each
constexpr int SMALL_WORDS = 252;
struct SmallBlob {
    int n;
    int w[SMALL_WORDS];
};
// (the same launch can zero up to a few ints elsewhere: result slots, error flags)
int upload_small(const void* host, size_t bytes, void* d_dst, cudaStream_t stream, int* zero_dst = nullptr, int zero_n = 0);
inline int upload_offsets(const int* host_offsets, int n_plus_1, int* d_dst, cudaStream_t stream, int* zero_dst = nullptr,
                          int zero_n = 0) {
    return upload_small(host_offsets, (size_t)n_plus_1 * sizeof(int), d_dst, stream, zero_dst, zero_n);
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline int num_bits(unsigned long long v) {  // bits needed to represent values in [0, v]
    int b = 0;
    while (v) { b++; v >>= 1; }
    return b ? b : 1;
}

// ---- device helpers ------------------------------------------------------------------------------------------------
// The reference's C++ was built without FMA contraction: every expression whose rounding
// matters for bit-exact indices is written with explicit round-to-nearest intrinsics, never a*b+c.
__device__ __forceinline__ float sq_dist_ref(float ax, float ay, float az, float bx, float by, float bz) {
    // cloud.h:71-74 / nanoflann.hpp:432-440: ((dx*dx + dy*dy) + dz*dz), d = a - b
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// order-preserving float <-> uint encoding for atomicMin / atomicMax on floats
__device__ __forceinline__ unsigned int f2ord(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {  // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

// batch element of a stacked index: offsets[0..nb] ascending, offsets[nb] = total
__device__ __forceinline__ int batch_of(const int* __restrict__ offsets, int nb, int i) {
    int lo = 0, hi = nb - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (offsets[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// exclusive scan of one int per thread across the CTA (blockDim.x a multiple of 32, <= 1024);
// smem_warp needs 33 ints; *total receives the CTA-wide sum.
__device__ __forceinline__ int block_exclusive_scan(int v, int* smem_warp, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) smem_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < (blockDim.x >> 5)) ? smem_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        smem_warp[lane] = wi - w;  // exclusive warp offsets
        if (lane == 31) smem_warp[32] = wi;
    }
    __syncthreads();
    int res = incl - v + smem_warp[warp];
    if (total) *total = smem_warp[32];
    __syncthreads();
    return res;
}

// ---- device-wide utilities (scan.cu / radix_sort.cu) -----------------------------------------------------------------
// Exclusive prefix sum of n int32 values; writes the grand total to *d_total when non-null. `tmp` must hold
// scan_tmp_ints(n) ints.
size_t scan_tmp_ints(int n);
int exclusive_scan(const int* d_in, int* d_out, int n, int* d_total, int* tmp, cudaStream_t stream);

// Stable LSD radix sort of (key u32, value u32) pairs on key bits [0, nbits). Result ends in keys_out / vals_out
// regardless of the pass count. tmp must hold sort_tmp_ints(n) ints.
size_t sort_tmp_ints(int n);
int stable_sort_pairs(const unsigned* keys_in, const unsigned* vals_in, unsigned* keys_out, unsigned* vals_out,
                      unsigned* keys_alt, unsigned* vals_alt, int n, int nbits, int* tmp, cudaStream_t stream);

}  // namespace kp
