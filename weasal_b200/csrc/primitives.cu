// Device-wide building blocks used by the precompute kernels: exclusive scan and a stable LSD radix sort.
// Hand-written (no CUB / Thrust): the sizes on this path (10^3 .. 10^7 keys, 8..24 significant bits) are
// launch- and HBM-bound, so the kernels are kept few and fat: one histogram, one single-CTA scan and one
// ranked scatter per 8-bit digit.
#include "common.cuh"

#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

namespace kp {

thread_local std::string g_last_error;
std::atomic<long long> g_launch_count{0};
bool g_profile_on = false;
namespace {
struct ProfRec { std::string tag; cudaEvent_t a, b; };
std::vector<ProfRec> g_prof;
std::mutex g_prof_mu;
}  // namespace
void profile_push(const char* tag, cudaEvent_t a, cudaEvent_t b) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back({tag, a, b});
}
// "tag count total_ms" lines; drains the record list
int profile_read(char* buf, int buflen) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<long long, double>> acc;
    for (auto& r : g_prof) {
        float ms = 0.f;
        cudaEventSynchronize(r.b);
        cudaEventElapsedTime(&ms, r.a, r.b);
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
        auto& e = acc[r.tag];
        e.first++;
        e.second += ms;
    }
    g_prof.clear();
    std::string out;
    for (auto& kv : acc) {
        char line[160];
        snprintf(line, sizeof line, "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if ((int)out.size() + 1 > buflen) return -1;
    memcpy(buf, out.c_str(), out.size() + 1);
    return (int)out.size();
}

namespace {
struct ArenaKey { cudaStream_t stream; int device; Arena arena; };
thread_local std::vector<ArenaKey*> g_arenas;
}  // namespace

Arena* arena_for(cudaStream_t stream) {
    int dev = 0;
    cudaGetDevice(&dev);
    for (auto* k : g_arenas)
        if (k->stream == stream && k->device == dev) return &k->arena;
    auto* k = new ArenaKey{stream, dev, Arena()};
    g_arenas.push_back(k);
    return &k->arena;
}

int arena_begin(Arena* a, cudaStream_t stream) {
    (void)stream;
    // blocks are never merged or freed: releasing device memory synchronises the whole device (measured: a merge
    // cost a 120 ms step); blocks grow geometrically, so there are only a handful
    a->cur = 0;
    a->off = 0;
    return KP_OK;
}

void* arena_alloc(Arena* a, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    while (a->cur < a->nblocks) {
        ArenaBlock& b = a->blocks[a->cur];
        if (a->off + bytes <= b.cap) {
            void* p = b.base + a->off;
            a->off += bytes;
            return p;
        }
        a->cur++;
        a->off = 0;
    }
    if (a->nblocks >= 16) return nullptr;
    size_t prev = a->nblocks ? a->blocks[a->nblocks - 1].cap : 0;
    size_t cap = bytes > 2 * prev ? bytes : 2 * prev;
    if (cap < ((size_t)128 << 20)) cap = (size_t)128 << 20;
    void* p = nullptr;
    if (cudaMalloc(&p, cap) != cudaSuccess) return nullptr;
    if (getenv("WEASAL_DEBUG")) fprintf(stderr, "[weasal_b200] arena grows: +%zu MB (block %d)\n", cap >> 20, a->nblocks);
    a->blocks[a->nblocks] = {(char*)p, cap};
    a->cur = a->nblocks++;
    a->off = bytes;
    return p;
}

__global__ void write_small_kernel(SmallBlob b, int* dst, int* zero_dst, int zero_n) {
    for (int i = threadIdx.x; i < b.n; i += blockDim.x) dst[i] = b.w[i];
    for (int i = threadIdx.x; i < zero_n; i += blockDim.x) zero_dst[i] = 0;
}

int upload_small(const void* host, size_t bytes, void* d_dst, cudaStream_t stream, int* zero_dst, int zero_n) {
    if (bytes == 0 && zero_n == 0) return KP_OK;
    if ((bytes & 3) == 0 && bytes <= SMALL_WORDS * sizeof(int)) {
        SmallBlob b;
        b.n = (int)(bytes / 4);
        memcpy(b.w, host, bytes);
        write_small_kernel<<<1, 128, 0, stream>>>(b, (int*)d_dst, zero_dst, zero_n);
        KP_CHECK_LAUNCH();
    } else {
        KP_CUDA(cudaMemcpyAsync(d_dst, host, bytes, cudaMemcpyHostToDevice, stream));
        if (zero_n) KP_CUDA(cudaMemsetAsync(zero_dst, 0, (size_t)zero_n * sizeof(int), stream));
    }
    return KP_OK;
}

// ------------------------------------------------------------------------------------------------------------ scan
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(const int* __restrict__ in, int* __restrict__ out,
                                                                  int n, int* __restrict__ block_sums,
                                                                  int* __restrict__ total_out) {
    __shared__ int s_warp[33];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        sum += v[i];
    }
    int total;
    int excl = block_exclusive_scan(sum, s_warp, &total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = excl;
        excl += v[i];
    }
    if (threadIdx.x == 0) {
        if (block_sums) block_sums[blockIdx.x] = total;
        if (gridDim.x == 1 && total_out) *total_out = total;
    }
}

// single CTA, out of place: the whole scan in one launch for arrays up to 64k entries (launch-bound regime)
__global__ void __launch_bounds__(SCAN_THREADS) scan_small_kernel(const int* __restrict__ in, int* __restrict__ out, int m,
                                                                 int* __restrict__ total_out) {
    __shared__ int s_warp[33];
    const int chunk = (m + SCAN_THREADS - 1) / SCAN_THREADS;
    const int lo = min((int)threadIdx.x * chunk, m), hi = min(lo + chunk, m);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += in[i];
    int total;
    int excl = block_exclusive_scan(sum, s_warp, &total);
    for (int i = lo; i < hi; i++) {
        const int v = in[i];
        out[i] = excl;
        excl += v;
    }
    if (threadIdx.x == 0 && total_out) *total_out = total;
}

// single CTA: exclusive scan of m values in place (m <= 1024 * chunk)
__global__ void __launch_bounds__(SCAN_THREADS) scan_single_kernel(int* __restrict__ data, int m,
                                                                  int* __restrict__ total_out) {
    __shared__ int s_warp[33];
    const int chunk = (m + SCAN_THREADS - 1) / SCAN_THREADS;
    const int lo = threadIdx.x * chunk, hi = min(lo + chunk, m);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += data[i];
    int total;
    int excl = block_exclusive_scan(sum, s_warp, &total);
    for (int i = lo; i < hi; i++) {
        int v = data[i];
        data[i] = excl;
        excl += v;
    }
    if (threadIdx.x == 0 && total_out) *total_out = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(int* __restrict__ out, int n,
                                                               const int* __restrict__ block_sums) {
    const int off = block_sums[blockIdx.x];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) out[base + i] += off;
}

size_t scan_tmp_ints(int n) { return (size_t)ceil_div(n > 0 ? n : 1, SCAN_TILE) + 1; }

int exclusive_scan(const int* d_in, int* d_out, int n, int* d_total, int* tmp, cudaStream_t stream) {
    if (n <= 0) {
        if (d_total) KP_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int), stream));
        return KP_OK;
    }
    // one CTA, one launch up to 8k entries; beyond that its per-thread sequential walk (n/1024 dependent, uncoalesced
    // loads: 25-34 us at 40k entries) loses to the three coalesced launches below (~10 us)
    if (n <= 8192 && d_in != d_out) {
        scan_small_kernel<<<1, SCAN_THREADS, 0, stream>>>(d_in, d_out, n, d_total);
        KP_CHECK_LAUNCH();
        return KP_OK;
    }
    const int nblk = ceil_div(n, SCAN_TILE);
    scan_tiles_kernel<<<nblk, SCAN_THREADS, 0, stream>>>(d_in, d_out, n, nblk > 1 ? tmp : nullptr, d_total);
    KP_CHECK_LAUNCH();
    if (nblk > 1) {
        scan_single_kernel<<<1, SCAN_THREADS, 0, stream>>>(tmp, nblk, d_total);
        KP_CHECK_LAUNCH();
        scan_add_kernel<<<nblk, SCAN_THREADS, 0, stream>>>(d_out, n, tmp);
        KP_CHECK_LAUNCH();
    }
    return KP_OK;
}

// ------------------------------------------------------------------------------------------------------ radix sort
constexpr int RS_THREADS = 512;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;  // 4096 keys
constexpr int RS_MAX_BLOCKS = 296;               // 2 CTAs per SM on 148 SMs

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const unsigned* __restrict__ keys, int n, int shift,
                                                            int tiles_per_block, int* __restrict__ hist) {
    __shared__ int s_hist[256];
    for (int i = threadIdx.x; i < 256; i += RS_THREADS) s_hist[i] = 0;
    __syncthreads();
    const long long lo = (long long)blockIdx.x * tiles_per_block * RS_TILE;
    const long long hi = min((long long)n, lo + (long long)tiles_per_block * RS_TILE);
    for (long long i = lo + threadIdx.x; i < hi; i += RS_THREADS) atomicAdd(&s_hist[(keys[i] >> shift) & 255u], 1);
    __syncthreads();
    for (int d = threadIdx.x; d < 256; d += RS_THREADS) hist[d * gridDim.x + blockIdx.x] = s_hist[d];
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const unsigned* __restrict__ keys_in,
                                                               const unsigned* __restrict__ vals_in,
                                                               unsigned* __restrict__ keys_out,
                                                               unsigned* __restrict__ vals_out, int n, int shift,
                                                               int tiles_per_block, const int* __restrict__ offs) {
    __shared__ int s_run[256];
    __shared__ int s_wcnt[RS_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int d = threadIdx.x; d < 256; d += RS_THREADS) s_run[d] = offs[d * gridDim.x + blockIdx.x];
    for (int t = 0; t < tiles_per_block; t++) {
        const long long tile_base = ((long long)blockIdx.x * tiles_per_block + t) * RS_TILE;
        if (tile_base >= n) break;
        for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&s_wcnt[0][0])[i] = 0;
        __syncthreads();
        unsigned k[RS_ROUNDS], v[RS_ROUNDS];
        const long long seg = tile_base + warp * (32 * RS_ROUNDS);
#pragma unroll
        for (int r = 0; r < RS_ROUNDS; r++) {
            const long long i = seg + r * 32 + lane;
            const bool ok = i < n;
            k[r] = ok ? keys_in[i] : 0xffffffffu;
            v[r] = ok ? vals_in[i] : 0u;
            const unsigned d = ok ? ((k[r] >> shift) & 255u) : 256u;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            if (ok && (peers & lt_mask) == 0) s_wcnt[warp][d] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            int o = s_run[threadIdx.x];
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) {
                int c = s_wcnt[w][threadIdx.x];
                s_wcnt[w][threadIdx.x] = o;
                o += c;
            }
            s_run[threadIdx.x] = o;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RS_ROUNDS; r++) {
            const long long i = seg + r * 32 + lane;
            const bool ok = i < n;
            const unsigned d = ok ? ((k[r] >> shift) & 255u) : 256u;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            int pos = 0;
            if (ok) pos = s_wcnt[warp][d] + __popc(peers & lt_mask);
            __syncwarp();
            if (ok && (peers & lt_mask) == 0) s_wcnt[warp][d] += __popc(peers);
            __syncwarp();
            if (ok) {
                keys_out[pos] = k[r];
                vals_out[pos] = v[r];
            }
        }
        __syncthreads();
    }
}

size_t sort_tmp_ints(int n) {
    (void)n;
    return (size_t)256 * RS_MAX_BLOCKS + 8;
}

int stable_sort_pairs(const unsigned* keys_in, const unsigned* vals_in, unsigned* keys_out, unsigned* vals_out,
                      unsigned* keys_alt, unsigned* vals_alt, int n, int nbits, int* tmp, cudaStream_t stream) {
    if (n <= 0) return KP_OK;
    const int passes = (nbits + 7) / 8;
    const int ntiles = ceil_div(n, RS_TILE);
    const int grid = ntiles < RS_MAX_BLOCKS ? ntiles : RS_MAX_BLOCKS;
    const int tpb = ceil_div(ntiles, grid);
    const int grid2 = ceil_div(ntiles, tpb);
    if (passes == 0) {
        KP_CUDA(cudaMemcpyAsync(keys_out, keys_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream));
        KP_CUDA(cudaMemcpyAsync(vals_out, vals_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream));
        return KP_OK;
    }
    // ping-pong so that the last pass lands in (keys_out, vals_out)
    const unsigned* src_k = keys_in;
    const unsigned* src_v = vals_in;
    for (int p = 0; p < passes; p++) {
        const bool to_out = ((passes - 1 - p) % 2) == 0;
        unsigned* dst_k = to_out ? keys_out : keys_alt;
        unsigned* dst_v = to_out ? vals_out : vals_alt;
        rs_hist_kernel<<<grid2, RS_THREADS, 0, stream>>>(src_k, n, p * 8, tpb, tmp);
        KP_CHECK_LAUNCH();
        scan_single_kernel<<<1, SCAN_THREADS, 0, stream>>>(tmp, 256 * grid2, nullptr);
        KP_CHECK_LAUNCH();
        rs_scatter_kernel<<<grid2, RS_THREADS, 0, stream>>>(src_k, src_v, dst_k, dst_v, n, p * 8, tpb, tmp);
        KP_CHECK_LAUNCH();
        src_k = dst_k;
        src_v = dst_v;
    }
    return KP_OK;
}

}  // namespace kp
