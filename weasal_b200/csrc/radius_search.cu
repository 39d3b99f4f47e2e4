// Batch radius search on the GPU — replaces cpp_wrappers/cpp_neighbors (neighbors.cpp:211-332 as wired in at
// wrapper.cpp:197-198; ordering contract of neighbors.cpp:125-208).
//
// Semantics reproduced bit-exactly: neighbours of query i = supports j of the SAME batch element with
// d2 = ((qx-sx)^2 + (qy-sy)^2) + (qz-sz)^2 < r2 = radius*radius, all in unfused f32 (nanoflann.hpp:432-440,
// neighbors.cpp:226), strict '<' (nanoflann.hpp:249-251); rows sorted by d2 ascending; global support indices;
// rows padded with Ns (neighbors.cpp:322-324).
// Stated tie-break: (d2 ascending, support index ascending) — that of batch_ordered_neighbors (upper_bound
// insertion, neighbors.cpp:176-181). The wired-in nanoflann path sorts by d2 only (std::sort, nanoflann.hpp:208-214),
// so it can differ from this order only inside groups of exactly equal d2.
//
// Design: uniform hash grid over the supports (cell edge slightly above the radius so that every neighbour lies in
// the 27 cells around the query's cell), supports scattered into cell-contiguous order as float4 (x,y,z,index),
// one warp per query: 27 lanes probe the 27 cells, the warp sweeps each non-empty cell 32 candidates at a time,
// ballot-compacts the hits into shared memory, rank-sorts them by (d2, index) and writes the row.
// The row buffer is [Nq, cap]; rows keep their `cap` closest neighbours (what big_neighborhood_filter,
// datasets/common.py:336-346, does afterwards) and the true maximum count is returned for the caller to slice.
#include "common.cuh"

#include <cstdlib>
#include <vector>

namespace kp {

constexpr unsigned long long CELL_EMPTY = ~0ULL;
// hits staged per query in shared memory: the fast variant keeps 256 (2 KB per warp, 16 warps per CTA, four CTAs per
// SM); a query with more neighbours than that raises RS_ERR_RETRY and the call is repeated with the 1024-hit variant
constexpr int RS_HITS_SMALL = 256, RS_WARPS_SMALL = 16;
constexpr int RS_HITS_BIG = 1024, RS_WARPS_BIG = 4;
constexpr int RS_ERR_GRID = 1, RS_ERR_DENSE = 2, RS_ERR_RETRY = 4;

struct SearchParams {
    const float* q; int nq;
    const float* s; int ns;
    const int* q_off; const int* s_off; int nb;
    float r2;
    int shadow;    // padding value of the rows (the reference pads with ns, neighbors.cpp:324)
    int nq_rows;   // rows of `out` (>= nq); rows [nq, nq_rows) are filled with `shadow` (static-shape batches)
};

// one launch resets everything a search needs: hash table, bounding boxes, result slots, the cell-storage cursor
__global__ void __launch_bounds__(256) rs_init_kernel(unsigned long long* tkeys, int* tcount, int tsize, unsigned* bbox,
                                                     int nb, int* result2, int* cursor) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < tsize) { tkeys[i] = CELL_EMPTY; tcount[i] = 0; }
    if (i < nb * 6) bbox[i] = ((i % 6) < 3) ? 0xffffffffu : 0u;
    if (i < 2) result2[i] = 0;
    if (i == 0) *cursor = 0;
}

__global__ void __launch_bounds__(256) rs_bbox_kernel(const float* __restrict__ s, int ns,
                                                     const int* __restrict__ s_off, int nb,
                                                     unsigned* __restrict__ bbox) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = i < ns;
    int b = -1;
    unsigned v[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    if (ok) {
        b = batch_of(s_off, nb, i);
        v[0] = v[3] = f2ord(s[3 * (size_t)i]);
        v[1] = v[4] = f2ord(s[3 * (size_t)i + 1]);
        v[2] = v[5] = f2ord(s[3 * (size_t)i + 2]);
    }
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

// Cell edge = radius * (1 + slack). The slack covers the f32 rounding of (p - origin) * inv_cell, which grows with
// the cell coordinate: |error| <= ~4 * Vmax * 2^-24 cells, so slack = 2^-10 + Vmax * 2^-20 keeps every true
// neighbour within +-1 cell on each axis. Every CTA derives the same value from the bounding boxes (thread 0, then
// broadcast through shared memory); returns false when the grid would not fit 18 bits per axis.
__device__ __forceinline__ float block_inv_cell(const unsigned* __restrict__ bbox, int nb, float radius, float* s_slot,
                                                bool* fits) {
    if (threadIdx.x == 0) {
        float ext = 0.f;
        for (int b = 0; b < nb; b++)
            for (int a = 0; a < 3; a++) {
                const unsigned lo = bbox[b * 6 + a], hi = bbox[b * 6 + 3 + a];
                if (lo == 0xffffffffu) continue;  // empty batch element
                ext = fmaxf(ext, ord2f(hi) - ord2f(lo));
            }
        const float vmax = ext / radius + 2.f;
        const float slack = 0.0009765625f + vmax * 9.5367431640625e-07f;
        s_slot[0] = 1.f / (radius * (1.f + slack));
        s_slot[1] = (vmax < 260000.f) ? 1.f : 0.f;
    }
    __syncthreads();
    if (fits) *fits = s_slot[1] != 0.f;
    return s_slot[0];
}

__device__ __forceinline__ int cell_coord(float p, float origin, float inv_cell) {
    float v = floorf((p - origin) * inv_cell);
    v = fminf(fmaxf(v, -4.f), 262150.f);
    return (int)v;
}

__device__ __forceinline__ unsigned long long cell_key(int b, int cx, int cy, int cz) {
    return ((unsigned long long)b << 54) | ((unsigned long long)cz << 36) | ((unsigned long long)cy << 18) |
           (unsigned long long)cx;
}

__global__ void __launch_bounds__(256) rs_insert_kernel(const float* __restrict__ s, int ns,
                                                       const int* __restrict__ s_off, int nb,
                                                       const unsigned* __restrict__ bbox, float radius,
                                                       unsigned long long* __restrict__ tkeys,
                                                       int* __restrict__ tcount, int tmask, int* __restrict__ sslot,
                                                       int* __restrict__ srank) {
    __shared__ float s_plan[2];
    const float inv = block_inv_cell(bbox, nb, radius, s_plan, nullptr);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    const int b = batch_of(s_off, nb, i);
    const int cx = cell_coord(s[3 * (size_t)i], ord2f(bbox[b * 6 + 0]), inv);
    const int cy = cell_coord(s[3 * (size_t)i + 1], ord2f(bbox[b * 6 + 1]), inv);
    const int cz = cell_coord(s[3 * (size_t)i + 2], ord2f(bbox[b * 6 + 2]), inv);
    const unsigned long long key = cell_key(b, cx, cy, cz);
    unsigned slot = (unsigned)mix64(key) & (unsigned)tmask;
    while (true) {
        unsigned long long cur = tkeys[slot];
        if (cur == CELL_EMPTY) cur = atomicCAS(&tkeys[slot], CELL_EMPTY, key);
        if (cur == CELL_EMPTY || cur == key) break;
        slot = (slot + 1) & (unsigned)tmask;
    }
    sslot[i] = (int)slot;
    srank[i] = atomicAdd(&tcount[slot], 1);
}

// every occupied cell claims a contiguous range of the cell-sorted support array (the order of the ranges does not
// matter: rows are sorted by (d2, index) at the end)
__global__ void __launch_bounds__(256) rs_assign_kernel(const int* __restrict__ tcount, int* __restrict__ tstart,
                                                       int tsize, int* __restrict__ cursor) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (i < tsize) ? tcount[i] : 0;
    // one atomic per warp instead of one per occupied cell (tens of thousands of atomics on a single address serialise)
    const int lane = threadIdx.x & 31;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(cursor, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (c > 0) tstart[i] = base + incl - c;
}

__global__ void __launch_bounds__(256) rs_fill_kernel(const float* __restrict__ s, int ns,
                                                     const int* __restrict__ sslot, const int* __restrict__ srank,
                                                     const int* __restrict__ tstart, float4* __restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    float4 v;
    v.x = s[3 * (size_t)i]; v.y = s[3 * (size_t)i + 1]; v.z = s[3 * (size_t)i + 2];
    v.w = __int_as_float(i);
    sorted[tstart[sslot[i]] + srank[i]] = v;
}

template <typename OutT, int HITS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) rs_search_kernel(
    SearchParams P, const unsigned* __restrict__ bbox, float radius,
    const unsigned long long* __restrict__ tkeys, const int* __restrict__ tcount, const int* __restrict__ tstart,
    int tmask, const float4* __restrict__ sorted, OutT* __restrict__ out, int cap, int* __restrict__ hmax,
    int* __restrict__ err) {
    __shared__ float s_d2[WARPS][HITS];
    __shared__ int s_idx[WARPS][HITS];
    __shared__ float s_plan[2];
    bool fits;
    const float inv = block_inv_cell(bbox, P.nb, radius, s_plan, &fits);
    if (!fits && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(err, RS_ERR_GRID);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int qi = blockIdx.x * WARPS + warp;
    if (qi >= P.nq) {
        if (qi < P.nq_rows)
            for (int h = lane; h < cap; h += 32) out[(size_t)qi * cap + h] = (OutT)P.shadow;
        return;
    }
    float* hd = s_d2[warp];
    int* hi = s_idx[warp];

    const float qx = P.q[3 * (size_t)qi], qy = P.q[3 * (size_t)qi + 1], qz = P.q[3 * (size_t)qi + 2];
    const int b = batch_of(P.q_off, P.nb, qi);
    int count = 0;
    if (P.s_off[b + 1] > P.s_off[b]) {
        const int cx = cell_coord(qx, ord2f(bbox[b * 6 + 0]), inv);
        const int cy = cell_coord(qy, ord2f(bbox[b * 6 + 1]), inv);
        const int cz = cell_coord(qz, ord2f(bbox[b * 6 + 2]), inv);
        // lanes 0..26 look up the 27 surrounding cells
        int c_start = 0, c_cnt = 0;
        if (lane < 27) {
            const int nx = cx + (lane % 3) - 1, ny = cy + ((lane / 3) % 3) - 1, nz = cz + (lane / 9) - 1;
            if (nx >= 0 && ny >= 0 && nz >= 0 && nx < 262144 && ny < 262144 && nz < 262144) {
                const unsigned long long key = cell_key(b, nx, ny, nz);
                unsigned slot = (unsigned)mix64(key) & (unsigned)tmask;
                while (true) {
                    const unsigned long long cur = tkeys[slot];
                    if (cur == key) { c_start = tstart[slot]; c_cnt = tcount[slot]; break; }
                    if (cur == CELL_EMPTY) break;
                    slot = (slot + 1) & (unsigned)tmask;
                }
            }
        }
        // the candidates of all cells form one flat list (prefix sums over the lanes); the warp sweeps it 32 at a
        // time, so the loads of different cells are in flight together
        int incl = c_cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int excl = incl - c_cnt;
        for (int base = 0; base < total; base += 32) {
            const int c = base + lane;
            int l = 0;  // first lane whose inclusive prefix exceeds c
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const int t = __shfl_sync(0xffffffffu, incl, l + step - 1);
                if (t <= c) l += step;
            }
            l = min(l, 31);
            const int st = __shfl_sync(0xffffffffu, c_start, l);
            const int ex = __shfl_sync(0xffffffffu, excl, l);
            bool hit = false;
            float d2 = 0.f;
            int sj = 0;
            if (c < total) {
                const float4 sp = sorted[st + (c - ex)];
                d2 = sq_dist_ref(qx, qy, qz, sp.x, sp.y, sp.z);
                sj = __float_as_int(sp.w);
                hit = d2 < P.r2;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const int pos = count + __popc(m & lt_mask);
                if (pos < HITS) { hd[pos] = d2; hi[pos] = sj; }
            }
            count += __popc(m);
        }
    }
    if (count > HITS) {
        if (lane == 0) atomicOr(err, HITS == RS_HITS_BIG ? RS_ERR_DENSE : RS_ERR_RETRY);
        count = HITS;
    }
    if (lane == 0 && count > *(volatile int*)hmax) atomicMax(hmax, count);
    __syncwarp();
    // rank sort by (d2, index): stable, O(n^2 / 32) per query, n is a few dozen
    OutT* row = out + (size_t)qi * cap;
    for (int i = lane; i < count; i += 32) {
        const float di = hd[i];
        const int ii = hi[i];
        int rank = 0;
        for (int j = 0; j < count; j++) {
            const float dj = hd[j];
            rank += (dj < di || (dj == di && hi[j] < ii)) ? 1 : 0;
        }
        if (rank < cap) row[rank] = (OutT)ii;
    }
    for (int h = count + lane; h < cap; h += 32) row[h] = (OutT)P.shadow;
}

// ------------------------------------------------------------------------------------------ the fast path: two kernels
// An experiment, kept for the record and selectable with WEASAL_RS_FAST=1 (results at its call site): the warp-per-query
// kernel above spends ~600 warp instructions per query on shuffles, ballots and a rank sort (ncu: issue-bound at 57-72 %
// issue slots with DRAM at 1 %). This path splits the work by what each part is good at:
//   rs_collect   ONE THREAD per query walks the 27 cells and their candidates with plain loads and compares (~2000 thread
//                instructions = ~60 warp instructions per query) and appends the hits, as packed 64-bit keys
//                (d2 bits << 32 | support index: d2 >= 0, so the unsigned order of the keys IS the (d2, index) order of the
//                stated tie-break), to the query's row of a scratch matrix. Queries are taken in the cell-sorted order
//                of a grid built over them (`qorder`: for a conv search the support grid itself), so the 32 lanes of a warp
//                sit in the same or adjacent cells, probe the same hash slots and read the same candidates: their loads
//                coalesce into broadcasts.
//   rs_rows      one warp per query sorts its keys (n <= 32: one key per lane, rank by 32 shuffles; else through shared
//                memory) and writes the row: the `cap` smallest keys' indices, padded with the shadow value.
// A query with more than RS_FAST_HITS neighbours raises RS_ERR_RETRY: the caller repeats with the 1024-hit kernel.
constexpr int RS_FAST_HITS = 128;

__global__ void __launch_bounds__(128) rs_collect_kernel(SearchParams P, const unsigned* __restrict__ bbox, float radius,
                                                        const unsigned long long* __restrict__ tkeys,
                                                        const int* __restrict__ tcount, const int* __restrict__ tstart,
                                                        int tmask, const float4* __restrict__ sorted,
                                                        const float4* __restrict__ qorder,
                                                        unsigned long long* __restrict__ keys, int* __restrict__ cnt,
                                                        int* __restrict__ err) {
    __shared__ float s_plan[2];
    bool fits;
    const float inv = block_inv_cell(bbox, P.nb, radius, s_plan, &fits);
    if (!fits && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(err, RS_ERR_GRID);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.nq) return;
    const int qi = qorder ? __float_as_int(qorder[t].w) : t;
    const float qx = P.q[3 * (size_t)qi], qy = P.q[3 * (size_t)qi + 1], qz = P.q[3 * (size_t)qi + 2];
    const int b = batch_of(P.q_off, P.nb, qi);
    unsigned long long* row = keys + (size_t)qi * RS_FAST_HITS;
    int count = 0;
    if (P.s_off[b + 1] > P.s_off[b]) {
        const int cx = cell_coord(qx, ord2f(bbox[b * 6 + 0]), inv);
        const int cy = cell_coord(qy, ord2f(bbox[b * 6 + 1]), inv);
        const int cz = cell_coord(qz, ord2f(bbox[b * 6 + 2]), inv);
        for (int c = 0; c < 27; c++) {
            const int nx = cx + (c % 3) - 1, ny = cy + ((c / 3) % 3) - 1, nz = cz + (c / 9) - 1;
            if (nx < 0 || ny < 0 || nz < 0 || nx >= 262144 || ny >= 262144 || nz >= 262144) continue;
            const unsigned long long key = cell_key(b, nx, ny, nz);
            unsigned slot = (unsigned)mix64(key) & (unsigned)tmask;
            int c_start = 0, c_cnt = 0;
            while (true) {
                const unsigned long long cur = tkeys[slot];
                if (cur == key) { c_start = tstart[slot]; c_cnt = tcount[slot]; break; }
                if (cur == CELL_EMPTY) break;
                slot = (slot + 1) & (unsigned)tmask;
            }
            for (int e0 = 0; e0 < c_cnt; e0 += 4) {   // four candidates in flight
                float4 sp[4];
#pragma unroll
                for (int u = 0; u < 4; u++) sp[u] = sorted[c_start + min(e0 + u, c_cnt - 1)];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const float d2 = sq_dist_ref(qx, qy, qz, sp[u].x, sp[u].y, sp[u].z);
                    if (e0 + u < c_cnt && d2 < P.r2) {
                        if (count < RS_FAST_HITS)
                            row[count] = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(sp[u].w);
                        count++;
                    }
                }
            }
        }
    }
    if (count > RS_FAST_HITS) {
        atomicOr(err, RS_ERR_RETRY);
        count = RS_FAST_HITS;
    }
    cnt[qi] = count;
}

constexpr int RS_ROWS_WARPS = 8;
template <typename OutT>
__global__ void __launch_bounds__(RS_ROWS_WARPS * 32) rs_rows_kernel(int nq, int nq_rows, int shadow,
                                                                    const unsigned long long* __restrict__ keys,
                                                                    const int* __restrict__ cnt, OutT* __restrict__ out,
                                                                    int cap, int* __restrict__ hmax) {
    __shared__ unsigned long long s_keys[RS_ROWS_WARPS][RS_FAST_HITS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = blockIdx.x * RS_ROWS_WARPS + warp;
    if (qi >= nq_rows) return;
    OutT* row = out + (size_t)qi * cap;
    const int n = qi < nq ? cnt[qi] : 0;
    if (n > 0) {
        const unsigned long long* src = keys + (size_t)qi * RS_FAST_HITS;
        if (lane == 0 && n > *(volatile int*)hmax) atomicMax(hmax, n);
        if (n <= 32) {
            const unsigned long long mine = lane < n ? src[lane] : ~0ULL;
            int rank = 0;
            for (int j = 0; j < n; j++) rank += (__shfl_sync(0xffffffffu, mine, j) < mine) ? 1 : 0;   // keys are distinct
            if (lane < n && rank < cap) row[rank] = (OutT)(unsigned)(mine & 0xffffffffULL);
        } else {
            unsigned long long* sk = s_keys[warp];
            for (int i = lane; i < n; i += 32) sk[i] = src[i];
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                const unsigned long long mine = sk[i];
                int rank = 0;
                for (int j = 0; j < n; j++) rank += (sk[j] < mine) ? 1 : 0;
                if (rank < cap) row[rank] = (OutT)(unsigned)(mine & 0xffffffffULL);
            }
        }
    }
    for (int h = n + lane; h < cap; h += 32) row[h] = (OutT)shadow;
}

// ---------------------------------------------------------------------------------------------------------- host side
// A search grid lives in ONE caller-owned device buffer (grid_bytes) so that it can outlive the call that built it: in
// the pyramid the grid over layer l+1 at radius 2r serves the upsample search of layer l and the conv and pool
// searches of layer l+1 (datasets/common.py:505, 531, 534), i.e. 5 builds per batch instead of 13.
struct GridView {
    int ns, nb, tsize;
    float radius;
    unsigned long long* tkeys;
    int *tcount, *tstart, *s_off, *cursor;
    unsigned* bbox;
    float4* sorted;
};

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

static int grid_tsize(int ns) {
    int tsize = 1024;
    while (tsize < 2 * ns) tsize <<= 1;
    return tsize;
}

size_t grid_bytes(int ns, int nb) {
    const size_t t = (size_t)grid_tsize(ns);
    return align256(t * 8) + 2 * align256(t * 4) + align256((size_t)(nb + 1) * 4) + align256(4) + align256((size_t)nb * 24) +
           align256((size_t)(ns > 0 ? ns : 1) * 16);
}

static GridView grid_view(void* buf, int ns, int nb, float radius) {
    GridView g;
    g.ns = ns; g.nb = nb; g.radius = radius; g.tsize = grid_tsize(ns);
    char* p = (char*)buf;
    g.tkeys = (unsigned long long*)p; p += align256((size_t)g.tsize * 8);
    g.tcount = (int*)p; p += align256((size_t)g.tsize * 4);
    g.tstart = (int*)p; p += align256((size_t)g.tsize * 4);
    g.s_off = (int*)p; p += align256((size_t)(nb + 1) * 4);
    g.cursor = (int*)p; p += align256(4);
    g.bbox = (unsigned*)p; p += align256((size_t)nb * 24);
    g.sorted = (float4*)p;
    return g;
}

static int check_batches(const int* b_host, int nb, int n, std::vector<int>& off) {
    off.assign(nb + 1, 0);
    for (int b = 0; b < nb; b++) {
        if (b_host[b] < 0) return fail(KP_ERR_ARG, "batch_query: negative batch length");
        off[b + 1] = off[b] + b_host[b];
    }
    if (off[nb] != n) return fail(KP_ERR_ARG, "batch_query: batch lengths do not sum to N");
    return KP_OK;
}

// supports -> grid (5 launches): init, bounding boxes, hash insert + per-cell counts, cell ranges, cell-sorted copy
int grid_build_device(const float* s, int ns, const int* sb_host, int nb, float radius, void* grid_buf,
                      cudaStream_t stream) {
    if (ns < 0 || nb <= 0 || !(radius > 0.f)) return fail(KP_ERR_ARG, "batch_query: bad sizes / radius");
    if (nb > 1023) return fail(KP_ERR_UNSUPPORTED, "batch_query: more than 1023 batch elements");
    std::vector<int> soff;
    int rc = check_batches(sb_host, nb, ns, soff);
    if (rc != KP_OK) return rc;
    GridView g = grid_view(grid_buf, ns, nb, radius);
    Scratch S(stream);
    int* d_sslot = S.alloc<int>(ns);
    int* d_srank = S.alloc<int>(ns);
    int* d_dummy = S.alloc<int>(2);
    if (S.status != KP_OK) return S.status;
    rc = upload_offsets(soff.data(), nb + 1, g.s_off, stream);
    if (rc != KP_OK) return rc;
    ProfileScope ps("rs_build", stream);
    rs_init_kernel<<<ceil_div(g.tsize > nb * 6 ? g.tsize : nb * 6, 256), 256, 0, stream>>>(g.tkeys, g.tcount, g.tsize, g.bbox,
                                                                                        nb, d_dummy, g.cursor);
    KP_CHECK_LAUNCH();
    if (ns > 0) {
        rs_bbox_kernel<<<ceil_div(ns, 256), 256, 0, stream>>>(s, ns, g.s_off, nb, g.bbox);
        KP_CHECK_LAUNCH();
        rs_insert_kernel<<<ceil_div(ns, 256), 256, 0, stream>>>(s, ns, g.s_off, nb, g.bbox, radius, g.tkeys, g.tcount,
                                                              g.tsize - 1, d_sslot, d_srank);
        KP_CHECK_LAUNCH();
        rs_assign_kernel<<<ceil_div(g.tsize, 256), 256, 0, stream>>>(g.tcount, g.tstart, g.tsize, g.cursor);
        KP_CHECK_LAUNCH();
        rs_fill_kernel<<<ceil_div(ns, 256), 256, 0, stream>>>(s, ns, d_sslot, d_srank, g.tstart, g.sorted);
        KP_CHECK_LAUNCH();
    }
    return KP_OK;
}

// queries against a built grid. out is [nq, cap] (int32 or int64); rows keep their cap closest neighbours.
// d_result != null: device int[2] receives {true max count, error bits}, no host sync. Otherwise the call synchronises,
// escalates to the 1024-hit kernel if needed and stores the true max count in *hmax_host.
int grid_query_device_ex(const void* grid_buf, int ns, int nb, float radius, const float* q, int nq, const int* qb_host,
                         void* out, int out_is_i64, int cap, int* hmax_host, int* d_result, int shadow, int nq_rows,
                         cudaStream_t stream, const void* qorder_grid = nullptr);
int grid_query_device(const void* grid_buf, int ns, int nb, float radius, const float* q, int nq, const int* qb_host,
                      void* out, int out_is_i64, int cap, int* hmax_host, int* d_result, cudaStream_t stream) {
    return grid_query_device_ex(grid_buf, ns, nb, radius, q, nq, qb_host, out, out_is_i64, cap, hmax_host, d_result, ns, nq,
                                stream);
}

// shadow: the value rows are padded with; nq_rows >= nq: rows of `out` (the extra ones are filled with `shadow`).
// qorder_grid: a grid built over the QUERY points themselves (any radius; the support grid itself when queries are the
// supports): its cell-sorted point order is the order queries are processed in (coherent warps). Null: index order.
int grid_query_device_ex(const void* grid_buf, int ns, int nb, float radius, const float* q, int nq, const int* qb_host,
                         void* out, int out_is_i64, int cap, int* hmax_host, int* d_result, int shadow, int nq_rows,
                         cudaStream_t stream, const void* qorder_grid) {
    if (nq < 0 || cap < 0 || nq_rows < nq) return fail(KP_ERR_ARG, "batch_query: bad sizes");
    std::vector<int> qoff;
    int rc = check_batches(qb_host, nb, nq, qoff);
    if (rc != KP_OK) return rc;
    if (hmax_host) *hmax_host = 0;
    if (nq_rows == 0) {
        if (d_result) KP_CUDA(cudaMemsetAsync(d_result, 0, 2 * sizeof(int), stream));
        return KP_OK;
    }
    GridView g = grid_view(const_cast<void*>(grid_buf), ns, nb, radius);
    Scratch S(stream);
    int* d_qoff = S.alloc<int>(nb + 1 + 2);
    int* d_hmax = d_result ? d_result : S.alloc<int>(2);
    if (S.status != KP_OK) return S.status;
    int* d_err = d_hmax + 1;
    {   // offsets and the zeroed result slots travel in one kernel-argument upload when the slots are ours
        rc = upload_offsets(qoff.data(), nb + 1, d_qoff, stream, d_hmax, 2);
        if (rc != KP_OK) return rc;
    }
    SearchParams P;
    P.q = q; P.nq = nq; P.s = nullptr; P.ns = ns; P.q_off = d_qoff; P.s_off = g.s_off; P.nb = nb;
    P.r2 = radius * radius;  // neighbors.cpp:226, f32
    P.shadow = shadow; P.nq_rows = nq_rows;
    // Measured on B200 (profiles/r2_radius_search_ab.txt): the thread-per-query path is NOT faster than the warp-per-query
    // kernel (0.49 vs 0.50 ms at 420 k queries, 0.17 vs 0.14 ms at 42 k, 0.89 vs 0.39 ms over the 13 searches of a training
    // batch, where 38 k threads cannot fill the device), so it stays an opt-in experiment (WEASAL_RS_FAST=1).
    static const bool fast_off = !(getenv("WEASAL_RS_FAST") && atoi(getenv("WEASAL_RS_FAST")) == 1);
    unsigned long long* d_keys = nullptr;
    int* d_cnt = nullptr;
    if (!fast_off && nq > 0) {
        d_keys = S.alloc<unsigned long long>((size_t)nq * RS_FAST_HITS);
        d_cnt = S.alloc<int>(nq);
        if (S.status != KP_OK) return S.status;
    }
    const float4* qorder = qorder_grid ? grid_view(const_cast<void*>(qorder_grid), nq, nb, 1.f).sorted : nullptr;
    auto launch = [&](bool big) {
        ProfileScope ps2("rs_search", stream);
        if (!big && d_keys) {
            rs_collect_kernel<<<ceil_div(nq, 128), 128, 0, stream>>>(P, g.bbox, radius, g.tkeys, g.tcount, g.tstart, g.tsize - 1,
                                                                    g.sorted, qorder, d_keys, d_cnt, d_err);
            const int grid = ceil_div(nq_rows, RS_ROWS_WARPS);
            if (out_is_i64)
                rs_rows_kernel<long long><<<grid, RS_ROWS_WARPS * 32, 0, stream>>>(nq, nq_rows, shadow, d_keys, d_cnt, (long long*)out, cap, d_hmax);
            else
                rs_rows_kernel<int><<<grid, RS_ROWS_WARPS * 32, 0, stream>>>(nq, nq_rows, shadow, d_keys, d_cnt, (int*)out, cap, d_hmax);
        } else if (!big) {
            const int grid = ceil_div(nq_rows, RS_WARPS_SMALL);
            if (out_is_i64)
                rs_search_kernel<long long, RS_HITS_SMALL, RS_WARPS_SMALL><<<grid, RS_WARPS_SMALL * 32, 0, stream>>>(
                    P, g.bbox, radius, g.tkeys, g.tcount, g.tstart, g.tsize - 1, g.sorted, (long long*)out, cap, d_hmax, d_err);
            else
                rs_search_kernel<int, RS_HITS_SMALL, RS_WARPS_SMALL><<<grid, RS_WARPS_SMALL * 32, 0, stream>>>(
                    P, g.bbox, radius, g.tkeys, g.tcount, g.tstart, g.tsize - 1, g.sorted, (int*)out, cap, d_hmax, d_err);
        } else {
            const int grid = ceil_div(nq_rows, RS_WARPS_BIG);
            if (out_is_i64)
                rs_search_kernel<long long, RS_HITS_BIG, RS_WARPS_BIG><<<grid, RS_WARPS_BIG * 32, 0, stream>>>(
                    P, g.bbox, radius, g.tkeys, g.tcount, g.tstart, g.tsize - 1, g.sorted, (long long*)out, cap, d_hmax, d_err);
            else
                rs_search_kernel<int, RS_HITS_BIG, RS_WARPS_BIG><<<grid, RS_WARPS_BIG * 32, 0, stream>>>(
                    P, g.bbox, radius, g.tkeys, g.tcount, g.tstart, g.tsize - 1, g.sorted, (int*)out, cap, d_hmax, d_err);
        }
    };
    launch(false);
    KP_CHECK_LAUNCH();
    if (d_result) return KP_OK;
    int h[2] = {0, 0};
    KP_CUDA(cudaMemcpyAsync(h, d_hmax, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    KP_CUDA(cudaStreamSynchronize(stream));
    if (h[1] & RS_ERR_RETRY) {  // a query outgrew the 256-hit staging: repeat with the 1024-hit variant
        KP_CUDA(cudaMemsetAsync(d_hmax, 0, 2 * sizeof(int), stream));
        launch(true);
        KP_CHECK_LAUNCH();
        KP_CUDA(cudaMemcpyAsync(h, d_hmax, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
        KP_CUDA(cudaStreamSynchronize(stream));
    }
    if (h[1] & RS_ERR_GRID) return fail(KP_ERR_UNSUPPORTED, "batch_query: cloud extent / radius exceeds 2^18 cells per axis");
    if (h[1] & RS_ERR_DENSE) return fail(KP_ERR_TOO_DENSE, "batch_query: more than 1024 neighbours for one query");
    *hmax_host = h[0];
    return KP_OK;
}

// one-shot search: the grid lives in this call's scratch
int batch_query_device(const float* q, int nq, const float* s, int ns, const int* qb_host, const int* sb_host, int nb,
                       float radius, void* out, int out_is_i64, int cap, int* hmax_host, int* d_result,
                       cudaStream_t stream) {
    if (nq < 0 || ns < 0 || nb <= 0 || cap < 0 || !(radius > 0.f)) return fail(KP_ERR_ARG, "batch_query: bad sizes / radius");
    void* buf = nullptr;
    {
        // the grid must survive the Scratch objects of the two calls below, which share this thread's arena: take it
        // from the arena first and let the nested calls allocate after it
        Scratch S(stream);
        buf = S.alloc<char>(grid_bytes(ns, nb));
        if (S.status != KP_OK) return S.status;
        ArenaHold hold(S);
        int rc = grid_build_device(s, ns, sb_host, nb, radius, buf, stream);
        if (rc != KP_OK) return rc;
        return grid_query_device_ex(buf, ns, nb, radius, q, nq, qb_host, out, out_is_i64, cap, hmax_host, d_result, ns, nq,
                                    stream, (q == s && nq == ns) ? buf : nullptr);
    }
}

}  // namespace kp
