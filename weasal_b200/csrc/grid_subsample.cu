// Grid subsampling on the GPU — replaces cpp_wrappers/cpp_subsampling (grid_subsampling.cpp:5-106, :109-211).
//
// Pipeline (all on one stream, one host sync at the end to learn the voxel count):
//   bbox          per-batch-element min/max of the (optionally rotated) points
//   key + insert  voxel key per point in the reference's f32 arithmetic; open-addressing hash insert -> voxel slot;
//                 atomicMin gives each voxel its first input index
//   first-flags   exclusive scan of "is first point of its voxel" = the voxel's rank in first-occurrence order
//   order         (reference order only) one CTA per batch element replays libstdc++'s unordered_map
//                 insert / rehash rounds in parallel to get the iteration order the reference emits voxels in
//   sort          stable radix sort of points by their voxel's output position (few bits: log2 N)
//   reduce        one thread per voxel: in-input-order f32 sums (no FMA), barycentre = sum * (float)(1.0/count),
//                 features = sum / (float)count, labels = first maximum of the histogram in unordered_map<int,int>
//                 iteration order
// Voxel membership, output order, barycentres, features and labels are bit-identical to the reference built with
// the same libstdc++ (the rehash schedule is taken from std::__detail::_Prime_rehash_policy at run time).
#include "common.cuh"

#include <cooperative_groups.h>
#include <cstdlib>

#include <climits>
#include <unordered_map>
#include <vector>

namespace kp {

constexpr unsigned long long HT_EMPTY = ~0ULL;
constexpr int MAX_SCHED = 48;
constexpr int MAX_LABELS = 59;  // distinct labels per voxel handled exactly (two rehashes of the histogram map)

struct Sched {
    long long elt[MAX_SCHED];
    long long bkt[MAX_SCHED];
    int n;
};

// rehash schedule of std::unordered_map with libstdc++'s prime policy: while inserting, when size == elt[i]
// the table is rebuilt with bkt[i] buckets (bits/hashtable_policy.h _Prime_rehash_policy::_M_need_rehash).
static Sched make_schedule(long long max_elts) {
    Sched s;
    s.n = 0;
    std::__detail::_Prime_rehash_policy pol;
    size_t bkt = 1;
    long long e = 0;
    while (e <= max_elts && s.n < MAX_SCHED) {
        auto r = pol._M_need_rehash(bkt, (size_t)e, 1);
        if (r.first) {
            bkt = r.second;
            s.elt[s.n] = e;
            s.bkt[s.n] = (long long)bkt;
            s.n++;
        }
        long long nxt = (long long)pol._M_next_resize;
        e = nxt > e ? nxt : e + 1;
    }
    return s;
}

struct GsParams {
    const float* pts;
    int n, nb;
    const int* offsets;  // [nb+1] device
    const float* rot;    // [nb*9] device or null
    float dl, inv_dl;
};

// datasets/common.py:118: out[j] = (p0*R[0][j] + p1*R[1][j]) + p2*R[2][j], f32, no FMA
__device__ __forceinline__ void rotate_fwd(const float* R, float& x, float& y, float& z) {
    float a = __fadd_rn(__fadd_rn(__fmul_rn(x, R[0]), __fmul_rn(y, R[3])), __fmul_rn(z, R[6]));
    float b = __fadd_rn(__fadd_rn(__fmul_rn(x, R[1]), __fmul_rn(y, R[4])), __fmul_rn(z, R[7]));
    float c = __fadd_rn(__fadd_rn(__fmul_rn(x, R[2]), __fmul_rn(y, R[5])), __fmul_rn(z, R[8]));
    x = a; y = b; z = c;
}
// datasets/common.py:134: multiply by R.T
__device__ __forceinline__ void rotate_bwd(const float* R, float& x, float& y, float& z) {
    float a = __fadd_rn(__fadd_rn(__fmul_rn(x, R[0]), __fmul_rn(y, R[1])), __fmul_rn(z, R[2]));
    float b = __fadd_rn(__fadd_rn(__fmul_rn(x, R[3]), __fmul_rn(y, R[4])), __fmul_rn(z, R[5]));
    float c = __fadd_rn(__fadd_rn(__fmul_rn(x, R[6]), __fmul_rn(y, R[7])), __fmul_rn(z, R[8]));
    x = a; y = b; z = c;
}

__device__ __forceinline__ void load_point(const GsParams& P, int i, int b, float& x, float& y, float& z) {
    x = P.pts[3 * (size_t)i]; y = P.pts[3 * (size_t)i + 1]; z = P.pts[3 * (size_t)i + 2];
    if (P.rot) rotate_fwd(P.rot + 9 * b, x, y, z);
}

__global__ void __launch_bounds__(256) gs_bbox_kernel(GsParams P, unsigned* __restrict__ bbox) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = i < P.n;
    int b = -1;
    unsigned v[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    if (ok) {
        b = batch_of(P.offsets, P.nb, i);
        float x, y, z;
        load_point(P, i, b, x, y, z);
        v[0] = v[3] = f2ord(x); v[1] = v[4] = f2ord(y); v[2] = v[5] = f2ord(z);
    }
    // one atomic per warp when the whole warp sits in one batch element (the common case)
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    const bool uniform = __all_sync(0xffffffffu, b == b0 || !ok) && b0 >= 0;
    if (uniform) {
#pragma unroll
        for (int a = 0; a < 6; a++) {
            unsigned r = (a < 3) ? __reduce_min_sync(0xffffffffu, v[a]) : __reduce_max_sync(0xffffffffu, v[a]);
            if ((threadIdx.x & 31) == 0) {
                if (a < 3) atomicMin(&bbox[b0 * 6 + a], r); else atomicMax(&bbox[b0 * 6 + a], r);
            }
        }
    } else if (ok) {
#pragma unroll
        for (int a = 0; a < 6; a++) {
            if (a < 3) atomicMin(&bbox[b * 6 + a], v[a]); else atomicMax(&bbox[b * 6 + a], v[a]);
        }
    }
}

// grid_subsampling.cpp:27-31, 53-56
__device__ __forceinline__ unsigned long long voxel_key(const GsParams& P, const unsigned* __restrict__ bbox, int b,
                                                        float x, float y, float z) {
    const float ox = __fmul_rn(floorf(__fmul_rn(ord2f(bbox[b * 6 + 0]), P.inv_dl)), P.dl);
    const float oy = __fmul_rn(floorf(__fmul_rn(ord2f(bbox[b * 6 + 1]), P.inv_dl)), P.dl);
    const float oz = __fmul_rn(floorf(__fmul_rn(ord2f(bbox[b * 6 + 2]), P.inv_dl)), P.dl);
    const unsigned long long NX = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(ord2f(bbox[b * 6 + 3]), ox), P.dl)) + 1ULL;
    const unsigned long long NY = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(ord2f(bbox[b * 6 + 4]), oy), P.dl)) + 1ULL;
    const unsigned long long ix = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(x, ox), P.dl));
    const unsigned long long iy = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(y, oy), P.dl));
    const unsigned long long iz = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(z, oz), P.dl));
    return ix + NX * iy + NX * NY * iz;
}

// one launch resets the hash table, the bounding boxes and the error flag
__global__ void gs_init_kernel(unsigned long long* tkeys, int* tfirst, int tsize, unsigned* bbox, int nb, int* err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < tsize) { tkeys[i] = HT_EMPTY; tfirst[i] = INT_MAX; }
    if (i < nb * 6) bbox[i] = ((i % 6) < 3) ? 0xffffffffu : 0u;
    if (i == 0) *err = 0;
}

__global__ void __launch_bounds__(256) gs_insert_kernel(GsParams P, const unsigned* __restrict__ bbox,
                                                       unsigned long long* __restrict__ tkeys,
                                                       int* __restrict__ tfirst, int tmask,
                                                       unsigned long long* __restrict__ pkey, int* __restrict__ pslot,
                                                       int* __restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const int b = batch_of(P.offsets, P.nb, i);
    float x, y, z;
    load_point(P, i, b, x, y, z);
    const unsigned long long key = voxel_key(P, bbox, b, x, y, z);
    if (key >= (1ULL << 52)) atomicOr(err, 1);  // composite key below would overflow
    const unsigned long long ck = key * (unsigned long long)P.nb + (unsigned long long)b;
    pkey[i] = key;
    unsigned slot = (unsigned)mix64(ck) & (unsigned)tmask;
    while (true) {
        unsigned long long cur = tkeys[slot];
        if (cur == HT_EMPTY) cur = atomicCAS(&tkeys[slot], HT_EMPTY, ck);
        if (cur == HT_EMPTY || cur == ck) break;
        slot = (slot + 1) & (unsigned)tmask;
    }
    pslot[i] = (int)slot;
    atomicMin(&tfirst[slot], i);
}

__global__ void __launch_bounds__(256) gs_flag_kernel(int n, const int* __restrict__ pslot,
                                                     const int* __restrict__ tfirst, int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (tfirst[pslot[i]] == i) ? 1 : 0;
}

// sequence of distinct voxels in first-occurrence order: seq_key[rank], seq_slot[rank]; seq_start[b] per batch element
__global__ void __launch_bounds__(256) gs_sequence_kernel(int n, int nb, const int* __restrict__ offsets,
                                                         const int* __restrict__ flags, const int* __restrict__ rank,
                                                         const int* __restrict__ total,
                                                         const unsigned long long* __restrict__ pkey,
                                                         const int* __restrict__ pslot,
                                                         unsigned long long* __restrict__ seq_key,
                                                         int* __restrict__ seq_slot, int* __restrict__ seq_start) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) { seq_key[rank[i]] = pkey[i]; seq_slot[rank[i]] = pslot[i]; }
    if (i <= nb) seq_start[i] = (i < nb && offsets[i] < n) ? rank[offsets[i]] : *total;
}

// ---- libstdc++ unordered_map iteration order, replayed in parallel --------------------------------------------------
// Within one table generation (bucket count P) a node whose bucket is empty goes to the list head, otherwise to the
// front of its bucket's run; a rehash re-inserts the list, in list order, by the same rule. So for the insertion
// sequence S of a generation: position(t) = (number of nodes in buckets first touched after bucket(t)) + (number of
// nodes of bucket(t) inserted after t). One CTA per batch element walks the generations.
constexpr int ORD_THREADS = 1024;

__device__ __forceinline__ int ord_elem(const int* cur, int len0, int e0, int t) { return t < len0 ? cur[t] : e0 + (t - len0); }

__global__ void __launch_bounds__(ORD_THREADS) gs_order_kernel(const unsigned long long* __restrict__ seq_key,
                                                              const int* __restrict__ seq_start, Sched sched,
                                                              int* __restrict__ scratch,
                                                              const long long* __restrict__ scratch_off,
                                                              const int* __restrict__ in_offsets,
                                                              int* __restrict__ pos_out) {
    __shared__ int s_warp[33];
    const int b = blockIdx.x;
    const int s0 = seq_start[b];
    const int m = seq_start[b + 1] - s0;
    if (m <= 0) return;
    const unsigned long long* keys = seq_key + s0;
    const int cap = in_offsets[b + 1] - in_offsets[b];
    int* base_ptr = scratch + scratch_off[b];
    int* cur = base_ptr;
    int* nxt = cur + cap;
    int* mem = nxt + cap;
    int* wsum = mem + cap;
    int* bkt = wsum + cap;   // bucket of every element of the current generation (one modulo per element)
    int* bfirst = bkt + cap;
    // bucket arrays sized for the last generation this element can reach
    long long pcap = sched.bkt[0];
    for (int i = 0; i < sched.n; i++) if (sched.elt[i] < cap || i == 0) pcap = sched.bkt[i];
    int* bcnt = bfirst + pcap;
    int* bbase = bcnt + pcap;
    const int tid = threadIdx.x;
    // keys below 2^32 (any realistic voxel grid) take the 32-bit modulo, several times cheaper than the 64-bit one
    __shared__ int s_wide;
    if (tid == 0) s_wide = 0;
    __syncthreads();
    for (int t = tid; t < m; t += ORD_THREADS) if (keys[t] >> 32) s_wide = 1;
    __syncthreads();
    const bool wide = s_wide != 0;

    for (int g = 0; g < sched.n; g++) {
        const int e0 = (int)sched.elt[g];
        if (e0 >= m && g > 0) break;
        const unsigned long long Pn = (unsigned long long)sched.bkt[g];
        const int end = (g + 1 < sched.n && sched.elt[g + 1] < m) ? (int)sched.elt[g + 1] : m;
        const int len0 = e0;            // list length at the rehash
        const int ns = len0 + (end - e0);
        for (int j = tid; j < (int)Pn; j += ORD_THREADS) { bfirst[j] = INT_MAX; bcnt[j] = 0; }
        __syncthreads();
        for (int t = tid; t < ns; t += ORD_THREADS) {
            const unsigned long long key = keys[ord_elem(cur, len0, e0, t)];
            const int bk = wide ? (int)(key % Pn) : (int)((unsigned)key % (unsigned)Pn);
            bkt[t] = bk;
            atomicMin(&bfirst[bk], t);
            atomicAdd(&bcnt[bk], 1);
        }
        __syncthreads();
        // suffix sums of first-toucher weights (chunked: each thread owns a contiguous range)
        const int chunk = (ns + ORD_THREADS - 1) / ORD_THREADS;
        const int lo = min(tid * chunk, ns), hi = min(lo + chunk, ns);
        int local = 0;
        for (int t = lo; t < hi; t++) {
            const int bk = bkt[t];
            const int w = (bfirst[bk] == t) ? bcnt[bk] : 0;
            wsum[t] = w;
            local += w;
        }
        int total;
        int excl = block_exclusive_scan(local, s_warp, &total);
        for (int t = lo; t < hi; t++) {
            const int w = wsum[t];
            excl += w;
            if (w) bbase[bkt[t]] = total - excl;  // nodes in buckets first touched after this one
        }
        __syncthreads();
        for (int j = tid; j < (int)Pn; j += ORD_THREADS) bfirst[j] = 0;  // reuse as fill cursor
        __syncthreads();
        for (int t = tid; t < ns; t += ORD_THREADS) {
            const int bk = bkt[t];
            mem[bbase[bk] + atomicAdd(&bfirst[bk], 1)] = t;
        }
        __syncthreads();
        for (int t = tid; t < ns; t += ORD_THREADS) {
            const int bk = bkt[t];
            const int r0 = bbase[bk], c = bcnt[bk];
            int later = 0;
            for (int u = 0; u < c; u++) later += (mem[r0 + u] > t) ? 1 : 0;
            nxt[r0 + later] = ord_elem(cur, len0, e0, t);
        }
        __syncthreads();
        int* sw = cur; cur = nxt; nxt = sw;
        if (end >= m) break;
    }
    for (int r = tid; r < m; r += ORD_THREADS) pos_out[s0 + cur[r]] = r;
}

// The same replay spread over the whole device for large clouds (the dataset-preparation call on a whole tile,
// datasets/Vaihingen3D_PseudoLabel.py:802-805: 1e5 .. 1e7 points in ONE batch element): a cooperative launch, every phase
// of a generation a grid-stride loop, grid-wide barriers where the single-CTA kernel has __syncthreads. The work of a
// generation is embarrassingly parallel apart from one prefix sum; generations remain sequential (a rehash re-inserts the
// list in list order) but their sizes form a geometric series, so the last two carry most of the work. Batch elements
// are processed one after the other.
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(ORD_THREADS) gs_order_coop_kernel(const unsigned long long* __restrict__ seq_key,
                                                                   const int* __restrict__ seq_start, int nb, Sched sched,
                                                                   int* __restrict__ scratch,
                                                                   const long long* __restrict__ scratch_off,
                                                                   const int* __restrict__ in_offsets,
                                                                   int* __restrict__ block_sums,
                                                                   int* __restrict__ pos_out) {
    cg::grid_group grid = cg::this_grid();
    __shared__ int s_warp[33];
    __shared__ int s_base;
    const int tid = threadIdx.x;
    const long long gtid = (long long)blockIdx.x * ORD_THREADS + tid;
    const long long gsize = (long long)gridDim.x * ORD_THREADS;
    for (int b = 0; b < nb; b++) {
        const int s0 = seq_start[b];
        const int m = seq_start[b + 1] - s0;
        if (m <= 0) continue;
        const unsigned long long* keys = seq_key + s0;
        const int cap = in_offsets[b + 1] - in_offsets[b];
        int* cur = scratch + scratch_off[b];
        int* nxt = cur + cap;
        int* mem = nxt + cap;
        int* wsum = mem + cap;
        int* bkt = wsum + cap;
        int* bfirst = bkt + cap;
        long long pcap = sched.bkt[0];
        for (int i = 0; i < sched.n; i++) if (sched.elt[i] < cap || i == 0) pcap = sched.bkt[i];
        int* bcnt = bfirst + pcap;
        int* bbase = bcnt + pcap;
        for (int g = 0; g < sched.n; g++) {
            const int e0 = (int)sched.elt[g];
            if (e0 >= m && g > 0) break;
            const unsigned long long Pn = (unsigned long long)sched.bkt[g];
            const int end = (g + 1 < sched.n && sched.elt[g + 1] < m) ? (int)sched.elt[g + 1] : m;
            const int len0 = e0;
            const int ns = len0 + (end - e0);
            for (long long j = gtid; j < (long long)Pn; j += gsize) { bfirst[j] = INT_MAX; bcnt[j] = 0; }
            grid.sync();
            for (long long t = gtid; t < ns; t += gsize) {
                const unsigned long long key = keys[ord_elem(cur, len0, e0, (int)t)];
                const int bk = (key >> 32) ? (int)(key % Pn) : (int)((unsigned)key % (unsigned)Pn);
                bkt[t] = bk;
                atomicMin(&bfirst[bk], (int)t);
                atomicAdd(&bcnt[bk], 1);
            }
            grid.sync();
            // suffix sums of the first-toucher weights: every thread owns a contiguous range, blocks meet through block_sums
            const long long chunk = (ns + gsize - 1) / gsize;
            const long long lo = min(gtid * chunk, (long long)ns), hi = min(lo + chunk, (long long)ns);
            int local = 0;
            for (long long t = lo; t < hi; t++) {
                const int bk = bkt[t];
                const int w = (bfirst[bk] == (int)t) ? bcnt[bk] : 0;
                wsum[t] = w;
                local += w;
            }
            int btotal;
            int excl = block_exclusive_scan(local, s_warp, &btotal);
            if (tid == 0) block_sums[blockIdx.x] = btotal;
            grid.sync();
            if (tid == 0) {
                int before = 0;
                for (int j = 0; j < (int)blockIdx.x; j++) before += block_sums[j];
                s_base = before;
            }
            __syncthreads();
            excl += s_base;   // weights before this thread's range; total = ns (every node is counted once)
            for (long long t = lo; t < hi; t++) {
                const int w = wsum[t];
                excl += w;
                if (w) bbase[bkt[t]] = ns - excl;  // nodes in buckets first touched after this one
            }
            for (long long j = gtid; j < (long long)Pn; j += gsize) bfirst[j] = 0;  // reuse as fill cursor (not read above)
            grid.sync();
            for (long long t = gtid; t < ns; t += gsize) {
                const int bk = bkt[t];
                mem[bbase[bk] + atomicAdd(&bfirst[bk], 1)] = (int)t;
            }
            grid.sync();
            for (long long t = gtid; t < ns; t += gsize) {
                const int bk = bkt[t];
                const int r0 = bbase[bk], c = bcnt[bk];
                int later = 0;
                for (int u = 0; u < c; u++) later += (mem[r0 + u] > (int)t) ? 1 : 0;
                nxt[r0 + later] = ord_elem(cur, len0, e0, (int)t);
            }
            grid.sync();
            int* sw = cur; cur = nxt; nxt = sw;
            if (end >= m) break;
        }
        for (long long r = gtid; r < m; r += gsize) pos_out[s0 + cur[r]] = (int)r;
        grid.sync();
    }
}

// per voxel (sequence index) -> global output position, honouring max_p truncation (grid_subsampling.cpp:181-204)
__global__ void __launch_bounds__(256) gs_outpos_kernel(int nb, const int* __restrict__ seq_start,
                                                       const int* __restrict__ seq_slot,
                                                       const int* __restrict__ pos_local, int max_p,
                                                       int* __restrict__ slot_outpos, int* __restrict__ out_lens,
                                                       int invalid) {
    const int total = seq_start[nb];
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < total) {
        const int b = batch_of(seq_start, nb, g);
        int outbase = 0;
        for (int j = 0; j < b; j++) outbase += min(seq_start[j + 1] - seq_start[j], max_p);
        const int pl = pos_local ? pos_local[g] : g - seq_start[b];
        slot_outpos[seq_slot[g]] = (pl < max_p) ? outbase + pl : invalid;
    }
    if (g < nb) out_lens[g] = min(seq_start[g + 1] - seq_start[g], max_p);
    if (g == 0) {
        int tot = 0;
        for (int j = 0; j < nb; j++) tot += min(seq_start[j + 1] - seq_start[j], max_p);
        out_lens[nb] = tot;
    }
}

__global__ void __launch_bounds__(256) gs_sortkey_kernel(int n, const int* __restrict__ pslot,
                                                        const int* __restrict__ slot_outpos,
                                                        unsigned* __restrict__ skey, unsigned* __restrict__ sval) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { skey[i] = (unsigned)slot_outpos[pslot[i]]; sval[i] = (unsigned)i; }
}

__global__ void __launch_bounds__(256) gs_segment_kernel(int n, const unsigned* __restrict__ skey, int invalid,
                                                        int* __restrict__ seg_start) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = skey[i];
    if (k != (unsigned)invalid && (i == 0 || skey[i - 1] != k)) seg_start[k] = i;
}

// iteration order of unordered_map<int,int> holding `nl` distinct labels inserted in first-occurrence order
// (literal list simulation, grid_subsampling.cpp:99-101); returns the label of the first maximum.
__device__ int label_argmax(const int* lab, const int* cnt, int nl, const Sched& sched, int* overflow) {
    if (nl > MAX_LABELS) { atomicOr(overflow, 2); nl = MAX_LABELS; }
    int nextn[MAX_LABELS];
    int bkt[MAX_LABELS];  // node before the bucket's first node; -1 = list head sentinel; -2 = empty
    int head = -1, nbk = 1, sp = 0;
    bkt[0] = -2;
    for (int i = 0; i < nl; i++) {
        if (sp < sched.n && i == (int)sched.elt[sp]) {  // rehash: re-insert the list in list order
            nbk = (int)sched.bkt[sp]; sp++;
            for (int j = 0; j < nbk; j++) bkt[j] = -2;
            int p = head; head = -1; int bbegin = 0;
            while (p != -1) {
                const int nx = nextn[p];
                const int bk = (int)((unsigned long long)(long long)lab[p] % (unsigned long long)nbk);
                if (bkt[bk] == -2) {
                    nextn[p] = head; head = p; bkt[bk] = -1;
                    if (nextn[p] != -1) bkt[bbegin] = p;
                    bbegin = bk;
                } else {
                    const int bf = bkt[bk];
                    if (bf == -1) { nextn[p] = head; head = p; } else { nextn[p] = nextn[bf]; nextn[bf] = p; }
                }
                p = nx;
            }
        }
        const int bk = (int)((unsigned long long)(long long)lab[i] % (unsigned long long)nbk);
        if (bkt[bk] != -2) {
            const int bf = bkt[bk];
            if (bf == -1) { nextn[i] = head; head = i; } else { nextn[i] = nextn[bf]; nextn[bf] = i; }
        } else {
            nextn[i] = head; head = i;
            if (nextn[i] != -1) bkt[(int)((unsigned long long)(long long)lab[nextn[i]] % (unsigned long long)nbk)] = i;
            bkt[bk] = -1;
        }
    }
    int best = head;
    for (int p = (head >= 0 ? nextn[head] : -1); p != -1; p = nextn[p]) if (cnt[best] < cnt[p]) best = p;
    return lab[best];
}

__global__ void __launch_bounds__(128) gs_reduce_kernel(GsParams P, int m_total, const unsigned* __restrict__ skey,
                                                       const unsigned* __restrict__ sval,
                                                       const int* __restrict__ seg_start,
                                                       const float* __restrict__ feats, int fdim,
                                                       const int* __restrict__ classes, int ldim, Sched lsched,
                                                       float* __restrict__ out_pts, float* __restrict__ out_feats,
                                                       int* __restrict__ out_classes, int* __restrict__ err) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m_total) return;
    const int s0 = seg_start[r];
    int s1 = s0;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    int b = 0;
    while (s1 < P.n && skey[s1] == (unsigned)r) {
        const int i = (int)sval[s1];
        if (s1 == s0) b = batch_of(P.offsets, P.nb, i);
        float x, y, z;
        load_point(P, i, b, x, y, z);
        sx = __fadd_rn(sx, x); sy = __fadd_rn(sy, y); sz = __fadd_rn(sz, z);  // SampledData::update_*, in input order
        s1++;
    }
    const int cnt = s1 - s0;
    const float w = (float)(1.0 / (double)cnt);  // grid_subsampling.cpp:87: double reciprocal narrowed to float
    float bx = __fmul_rn(sx, w), by = __fmul_rn(sy, w), bz = __fmul_rn(sz, w);
    if (P.rot) rotate_bwd(P.rot + 9 * b, bx, by, bz);
    out_pts[3 * (size_t)r] = bx; out_pts[3 * (size_t)r + 1] = by; out_pts[3 * (size_t)r + 2] = bz;
    if (feats) {
        const float fc = (float)cnt;
        for (int d = 0; d < fdim; d++) {
            float acc = 0.f;
            for (int t = s0; t < s1; t++) acc = __fadd_rn(acc, feats[(size_t)sval[t] * fdim + d]);
            out_feats[(size_t)r * fdim + d] = __fdiv_rn(acc, fc);  // :90-94
        }
    }
    if (classes) {
        for (int l = 0; l < ldim; l++) {
            int lab[MAX_LABELS], lc[MAX_LABELS], nl = 0;
            bool over = false;
            for (int t = s0; t < s1; t++) {
                const int c = classes[(size_t)sval[t] * ldim + l];
                int f = -1;
                for (int u = 0; u < nl; u++) if (lab[u] == c) { f = u; break; }
                if (f >= 0) lc[f]++;
                else if (nl < MAX_LABELS) { lab[nl] = c; lc[nl] = 1; nl++; }
                else over = true;
            }
            if (over) atomicOr(err, 2);
            out_classes[(size_t)r * ldim + l] = label_argmax(lab, lc, nl, lsched, err);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------- host side
// All pointers are device pointers except lens_host / rot_host / out_lens_host / m_host. Output buffers must hold n
// rows (the voxel count is only known after the run); *m_host receives the voxel count.
int grid_subsample_device(const float* pts, int n, const int* lens_host, int nb, const float* feats, int fdim,
                          const int* classes, int ldim, float dl, int max_p, int order_mode, const float* rot_host,
                          float* out_pts, int* out_lens_host, float* out_feats, int* out_classes, int* m_host,
                          cudaStream_t stream) {
    if (n < 0 || nb <= 0 || !(dl > 0.f)) return fail(KP_ERR_ARG, "grid_subsample: bad n / nb / sampleDl");
    if (classes && ldim > 1 && nb > 1)
        return fail(KP_ERR_UNSUPPORTED, "grid_subsample: batched label columns > 1 (the reference's slice is wrong there)");
    std::vector<int> offs(nb + 1, 0);
    for (int b = 0; b < nb; b++) {
        if (lens_host[b] < 0) return fail(KP_ERR_ARG, "grid_subsample: negative batch length");
        offs[b + 1] = offs[b] + lens_host[b];
    }
    if (offs[nb] != n) return fail(KP_ERR_ARG, "grid_subsample: batch lengths do not sum to N");
    for (int b = 0; b < nb; b++) out_lens_host[b] = 0;
    *m_host = 0;
    if (n == 0) return KP_OK;
    if (max_p < 1) max_p = n;

    Scratch S(stream);
    int* d_offs = S.alloc<int>(nb + 1);
    float* d_rot = rot_host ? S.alloc<float>((size_t)nb * 9) : nullptr;
    unsigned* d_bbox = S.alloc<unsigned>((size_t)nb * 6);
    int tsize = 1024;
    while (tsize < 2 * n) tsize <<= 1;
    unsigned long long* d_tkeys = S.alloc<unsigned long long>(tsize);
    int* d_tfirst = S.alloc<int>(tsize);
    int* d_slot_outpos = S.alloc<int>(tsize);
    unsigned long long* d_pkey = S.alloc<unsigned long long>(n);
    int* d_pslot = S.alloc<int>(n);
    int* d_flags = S.alloc<int>(n);
    int* d_rank = S.alloc<int>(n);
    int* d_total = S.alloc<int>(1);
    int* d_scan_tmp = S.alloc<int>(scan_tmp_ints(n));
    unsigned long long* d_seq_key = S.alloc<unsigned long long>(n);
    int* d_seq_slot = S.alloc<int>(n);
    int* d_seq_start = S.alloc<int>(nb + 1);
    int* d_out_lens = S.alloc<int>(nb + 1);
    int* d_err = S.alloc<int>(1);
    unsigned* d_skey = S.alloc<unsigned>(n);
    unsigned* d_sval = S.alloc<unsigned>(n);
    unsigned* d_skey2 = S.alloc<unsigned>(n);
    unsigned* d_sval2 = S.alloc<unsigned>(n);
    unsigned* d_skey3 = S.alloc<unsigned>(n);
    unsigned* d_sval3 = S.alloc<unsigned>(n);
    int* d_sort_tmp = S.alloc<int>(sort_tmp_ints(n));
    int* d_seg_start = S.alloc<int>(n + 1);
    if (S.status != KP_OK) return S.status;

    { int rc0 = upload_offsets(offs.data(), nb + 1, d_offs, stream); if (rc0 != KP_OK) return rc0; }
    if (rot_host) { int rc0 = upload_small(rot_host, (size_t)nb * 9 * sizeof(float), d_rot, stream); if (rc0 != KP_OK) return rc0; }

    GsParams P;
    P.pts = pts; P.n = n; P.nb = nb; P.offsets = d_offs; P.rot = d_rot; P.dl = dl;
    P.inv_dl = 1.0f / dl;  // grid_subsampling.cpp:27: `1/sampleDl` is an f32 division
    const int nblk = ceil_div(n, 256);

    ProfileScope* ps = new ProfileScope("gs_hash", stream);
    gs_init_kernel<<<ceil_div(tsize > nb * 6 ? tsize : nb * 6, 256), 256, 0, stream>>>(d_tkeys, d_tfirst, tsize, d_bbox, nb,
                                                                                    d_err);
    KP_CHECK_LAUNCH();
    gs_bbox_kernel<<<nblk, 256, 0, stream>>>(P, d_bbox);
    KP_CHECK_LAUNCH();
    gs_insert_kernel<<<nblk, 256, 0, stream>>>(P, d_bbox, d_tkeys, d_tfirst, tsize - 1, d_pkey, d_pslot, d_err);
    KP_CHECK_LAUNCH();
    gs_flag_kernel<<<nblk, 256, 0, stream>>>(n, d_pslot, d_tfirst, d_flags);
    KP_CHECK_LAUNCH();
    int rc = exclusive_scan(d_flags, d_rank, n, d_total, d_scan_tmp, stream);
    if (rc != KP_OK) return rc;
    gs_sequence_kernel<<<ceil_div(n > nb ? n : nb + 1, 256), 256, 0, stream>>>(n, nb, d_offs, d_flags, d_rank, d_total,
                                                                             d_pkey, d_pslot, d_seq_key, d_seq_slot,
                                                                             d_seq_start);
    KP_CHECK_LAUNCH();

    delete ps;
    int* d_pos_local = nullptr;
    if (order_mode == 1) {
        Sched sched = make_schedule(n);
        std::vector<long long> soff(nb);
        long long tot = 0;
        for (int b = 0; b < nb; b++) {
            soff[b] = tot;
            long long pcap = sched.bkt[0];
            for (int i = 0; i < sched.n; i++) if (sched.elt[i] < lens_host[b] || i == 0) pcap = sched.bkt[i];
            tot += 5LL * lens_host[b] + 3LL * pcap + 8;
        }
        if (tot > 0x7fffffffLL * 2) return fail(KP_ERR_UNSUPPORTED, "grid_subsample: cloud too large for reference-order scratch");
        int* d_scratch = S.alloc<int>((size_t)tot);
        long long* d_soff = S.alloc<long long>(nb);
        d_pos_local = S.alloc<int>(n);
        if (S.status != KP_OK) return S.status;
        { int rc0 = upload_small(soff.data(), nb * sizeof(long long), d_soff, stream); if (rc0 != KP_OK) return rc0; }
        ProfileScope pso("gs_order", stream);
        int max_len = 0;
        for (int b = 0; b < nb; b++) max_len = lens_host[b] > max_len ? lens_host[b] : max_len;
        static int coop_ctas = -1;   // co-resident CTAs of the cooperative kernel (0: cooperative launch not available)
        if (coop_ctas < 0) {
            int dev = 0, sms = 0, per_sm = 0, can = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gs_order_coop_kernel, ORD_THREADS, 0);
            coop_ctas = can ? sms * per_sm : 0;
        }
        static const int coop_min = getenv("WEASAL_GS_COOP_MIN") ? atoi(getenv("WEASAL_GS_COOP_MIN")) : 60000;
        if (coop_ctas > 1 && max_len >= coop_min) {
            // large elements: the replay on every SM (grid-wide barriers cost a few microseconds each, ~100 of them)
            int* d_bsums = S.alloc<int>(coop_ctas);
            if (S.status != KP_OK) return S.status;
            int ctas = ceil_div(max_len, ORD_THREADS);
            if (ctas > coop_ctas) ctas = coop_ctas;
            const unsigned long long* a0 = d_seq_key;
            const int* a1 = d_seq_start;
            int a2 = nb;
            int* a4 = d_scratch;
            const long long* a5 = d_soff;
            const int* a6 = d_offs;
            int* a8 = d_pos_local;
            void* args[] = {&a0, &a1, &a2, &sched, &a4, &a5, &a6, &d_bsums, &a8};
            KP_CUDA(cudaLaunchCooperativeKernel((const void*)gs_order_coop_kernel, dim3(ctas), dim3(ORD_THREADS), args, 0, stream));
            g_launch_count.fetch_add(1);
        } else {
            gs_order_kernel<<<nb, ORD_THREADS, 0, stream>>>(d_seq_key, d_seq_start, sched, d_scratch, d_soff, d_offs,
                                                           d_pos_local);
            KP_CHECK_LAUNCH();
        }
    }
    ProfileScope* ps3 = new ProfileScope("gs_sort", stream);
    const int invalid = n;  // sort key of points whose voxel was truncated by max_p
    gs_outpos_kernel<<<ceil_div(n > nb ? n : nb + 1, 256), 256, 0, stream>>>(nb, d_seq_start, d_seq_slot, d_pos_local,
                                                                           max_p, d_slot_outpos, d_out_lens, invalid);
    KP_CHECK_LAUNCH();
    gs_sortkey_kernel<<<nblk, 256, 0, stream>>>(n, d_pslot, d_slot_outpos, d_skey, d_sval);
    KP_CHECK_LAUNCH();
    rc = stable_sort_pairs(d_skey, d_sval, d_skey2, d_sval2, d_skey3, d_sval3, n, num_bits((unsigned long long)n),
                           d_sort_tmp, stream);
    if (rc != KP_OK) return rc;
    gs_segment_kernel<<<nblk, 256, 0, stream>>>(n, d_skey2, invalid, d_seg_start);
    delete ps3;
    KP_CHECK_LAUNCH();

    // the voxel count decides the reduce grid: the one host sync of this call
    std::vector<int> h_lens(nb + 1);
    int h_err = 0;
    KP_CUDA(cudaMemcpyAsync(h_lens.data(), d_out_lens, (nb + 1) * sizeof(int), cudaMemcpyDeviceToHost, stream));
    KP_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, stream));
    KP_CUDA(cudaStreamSynchronize(stream));
    if (h_err & 1) return fail(KP_ERR_UNSUPPORTED, "grid_subsample: voxel grid too large (key overflow)");
    const int m_total = h_lens[nb];
    Sched lsched = make_schedule(MAX_LABELS + 1);
    if (m_total > 0) {
        ProfileScope psr("gs_reduce", stream);
        gs_reduce_kernel<<<ceil_div(m_total, 128), 128, 0, stream>>>(P, m_total, d_skey2, d_sval2, d_seg_start, feats,
                                                                    fdim, classes, ldim, lsched, out_pts, out_feats,
                                                                    out_classes, d_err);
        KP_CHECK_LAUNCH();
    }
    if (classes) {
        KP_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, stream));
        KP_CUDA(cudaStreamSynchronize(stream));
        if (h_err & 2) return fail(KP_ERR_UNSUPPORTED, "grid_subsample: more than 59 distinct labels in one voxel");
    }
    for (int b = 0; b < nb; b++) out_lens_host[b] = h_lens[b];
    *m_host = m_total;
    return KP_OK;
}

}  // namespace kp
