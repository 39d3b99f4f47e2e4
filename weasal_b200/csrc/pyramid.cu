// Device pyramid builder — the per-layer walk of PointCloudDataset.segmentation_inputs (datasets/common.py:461-577)
// as ONE native call: conv search, grid subsampling (with the random grid orientation of common.py:89-135), pool
// search and upsample search per layer, every index matrix cropped to its neighbourhood limit (common.py:336-346).
//
// Why native: at ~40k points per batch the pyramid is ~150 small launches and the step is bound by host launch work.
// Issued from one C call that holds no Python lock, it can run on a side stream from a prefetch thread (the
// counterpart of the reference's DataLoader workers) while the training thread launches the network.
//
// Outputs are carved out of one caller-owned device slab in the order they become known (a layer's point count is
// only known after its subsampling), so a batch costs the caller one allocation; offsets into the slab are returned.
// Search grids are shared: the grid over layer l+1 at radius 2r serves the upsample search of layer l and the conv and
// pool searches of layer l+1.
#include "common.cuh"

#include <vector>

namespace kp {

size_t grid_bytes(int ns, int nb);
int grid_build_device(const float* s, int ns, const int* sb_host, int nb, float radius, void* grid_buf, cudaStream_t stream);
int grid_query_device_ex(const void* grid_buf, int ns, int nb, float radius, const float* q, int nq, const int* qb_host,
                         void* out, int out_is_i64, int cap, int* hmax_host, int* d_result, int shadow, int nq_rows,
                         cudaStream_t stream, const void* qorder_grid = nullptr);
int grid_subsample_device(const float* pts, int n, const int* lens_host, int nb, const float* feats, int fdim,
                          const int* classes, int ldim, float dl, int max_p, int order_mode, const float* rot_host,
                          float* out_pts, int* out_lens_host, float* out_feats, int* out_classes, int* m_host,
                          cudaStream_t stream);

// dst[0, n_dst) = src[0, n_src) followed by `pad` (static-shape batches: fixed row counts, padded tails)
template <typename T>
__global__ void __launch_bounds__(256) pad_copy_kernel(const T* __restrict__ src, long long n_src, T* __restrict__ dst,
                                                      long long n_dst, T pad) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_dst; i += (long long)gridDim.x * blockDim.x)
        dst[i] = i < n_src ? src[i] : pad;
}
template <typename T>
static int pad_copy(const T* src, long long n_src, T* dst, long long n_dst, T pad, cudaStream_t stream) {
    if (n_dst <= 0) return KP_OK;
    const int grid = ceil_div(n_dst, 256) < 1184 ? ceil_div(n_dst, 256) : 1184;
    pad_copy_kernel<T><<<grid, 256, 0, stream>>>(src, n_src, dst, n_dst, pad);
    KP_CHECK_LAUNCH();
    return KP_OK;
}

namespace {
struct Slab {
    char* base;
    long long cap, off = 0;
    bool over = false;
    long long take(long long bytes) {  // returns the offset of a 256-byte aligned range (or -1 once the slab is full)
        const long long o = (off + 255) & ~255LL;
        if (o + bytes > cap) { over = true; off = o + bytes; return -1; }
        off = o + bytes;
        return o;
    }
};

struct PendingSearch {
    int layer, kind;       // kind: 0 conv, 1 pool, 2 upsample
    const void* grid;
    int ns, nq, shadow, rows;
    float radius;
    const float* q;
    std::vector<int> qlens;
    void* out;
    int width;             // row stride of the matrix = columns kept
    bool limited;
};
}  // namespace

// layout of the integer outputs (all HOST arrays):
//   n_out[L]              points per layer
//   lens_out[L*nb]        batch lengths per layer
//   offs[5*L+3]           slab byte offsets: [0,L) points (layer 0: -1, the caller's own tensor), [L,2L) conv matrices,
//                         [2L,3L) pool matrices, [3L,4L) upsample matrices, [4L,5L) lengths (int32 [nb]); -1 = absent;
//                         [5L] features, [5L+1] labels, [5L+2] true widths int32 [3L] (static mode only)
//   widths[3*L]           true maximum neighbour count of the conv / pool / upsample search of each layer
//   strides[3*L]          row stride (columns stored) of those matrices
// Static mode (n_cap != null): every layer has a FIXED row count n_cap[l] (its points, its matrices and the layer-0
// features / labels are padded: points with 1e6, index rows with the shadow value, which is the support layer's
// n_cap, features with 0, labels with label_pad), so the slab layout depends on the caps alone and a consumer can bind
// it once (a captured CUDA graph). A layer that outgrows its cap returns KP_ERR_CAPACITY with *need_cap = -(l + 1).
// Returns KP_ERR_CAPACITY when the slab is too small (*need_bytes = required size) or when an unlimited search found
// rows wider than `cap` (*need_cap = required cap): the caller grows and repeats.
int pyramid_build_device(const float* pts0, int n0, const int* lens0, int nb, int L, const float* conv_r,
                         const float* pool_r, const float* up_r, const float* dl, const float* rot, const int* limits,
                         int order, int idx_is_i64, int cap, void* slab, long long slab_bytes, long long* offs, int* n_out,
                         int* lens_out, int* widths, int* strides, long long* need_bytes, int* need_cap,
                         const int* n_cap, const float* feats, int fdim, const long long* labels, long long label_pad,
                         cudaStream_t stream) {
    if (n0 <= 0 || nb <= 0 || L <= 0 || L > 16 || cap <= 0) return fail(KP_ERR_ARG, "pyramid: bad sizes");
    const int isz = idx_is_i64 ? 8 : 4;
    for (int i = 0; i < 5 * L + 3; i++) offs[i] = -1;
    if (n_cap && n0 > n_cap[0]) {
        *need_cap = -1;
        return fail(KP_ERR_CAPACITY, "pyramid: layer 0 has more points than its static capacity");
    }
    for (int i = 0; i < 3 * L; i++) { widths[i] = 0; strides[i] = 0; }
    *need_bytes = 0;
    *need_cap = cap;

    Scratch S(stream);
    int* d_results = S.alloc<int>((size_t)6 * L);
    if (S.status != KP_OK) return S.status;
    ArenaHold hold(S);  // nested entry points append to this call's scratch: the grids must outlive them
    KP_CUDA(cudaMemsetAsync(d_results, 0, (size_t)6 * L * sizeof(int), stream));

    Slab sl;
    sl.base = (char*)slab;
    sl.cap = slab_bytes;
    std::vector<PendingSearch> pend;

    const float* cur = pts0;
    int cur_n = n0;
    std::vector<int> cur_lens(lens0, lens0 + nb);
    const void* cur_grid = nullptr;
    float cur_grid_r = 0.f;

    auto lim = [&](int layer) { return (limits && layer < L && limits[layer] > 0) ? limits[layer] : 0; };
    auto get_grid = [&](const float* p, int n, const std::vector<int>& lens, float r, const void** g) -> int {
        void* buf = S.alloc<char>(grid_bytes(n, nb));
        if (S.status != KP_OK) return S.status;
        int rc = grid_build_device(p, n, lens.data(), nb, r, buf, stream);
        *g = buf;
        return rc;
    };
    // (rows, shadow): the matrix has `rows` rows (>= nq) and pads with `shadow`
    // qorder: a grid over the query points (their cell-sorted order makes the search's warps coherent)
    auto search = [&](int layer, int kind, const void* grid, int ns, float r, const float* q, int nq,
                      const std::vector<int>& qlens, int limit, int rows, int shadow, const void* qorder) -> int {
        const int width = limit > 0 ? limit : cap;
        const long long o = sl.take((long long)rows * width * isz);
        offs[(1 + kind) * L + layer] = o;
        strides[kind * L + layer] = width;
        if (o < 0) return KP_OK;  // slab exhausted: keep walking to learn the full size
        PendingSearch ps;
        ps.layer = layer; ps.kind = kind; ps.grid = grid; ps.ns = ns; ps.nq = nq; ps.radius = r; ps.q = q;
        ps.qlens = qlens; ps.out = sl.base + o; ps.width = width; ps.limited = limit > 0;
        ps.shadow = shadow; ps.rows = rows;
        pend.push_back(ps);
        return grid_query_device_ex(grid, ns, nb, r, q, nq, qlens.data(), ps.out, idx_is_i64, width, nullptr,
                                    d_results + 2 * (kind * L + layer), shadow, rows, stream, qorder);
    };

    int rc;
    if (n_cap) {  // layer 0 lives in the slab too, padded to its capacity, next to the padded features / labels
        const long long po = sl.take((long long)n_cap[0] * 12);
        offs[0] = po;
        if (po >= 0 && (rc = pad_copy<float>(pts0, (long long)n0 * 3, (float*)(sl.base + po), (long long)n_cap[0] * 3, 1e6f, stream)) != KP_OK) return rc;
        if (feats && fdim > 0) {
            const long long fo = sl.take((long long)n_cap[0] * fdim * 4);
            offs[5 * L] = fo;
            if (fo >= 0 && (rc = pad_copy<float>(feats, (long long)n0 * fdim, (float*)(sl.base + fo), (long long)n_cap[0] * fdim, 0.f, stream)) != KP_OK) return rc;
        }
        if (labels) {
            const long long lo = sl.take((long long)n_cap[0] * 8);
            offs[5 * L + 1] = lo;
            if (lo >= 0 && (rc = pad_copy<long long>(labels, n0, (long long*)(sl.base + lo), n_cap[0], label_pad, stream)) != KP_OK) return rc;
        }
        offs[5 * L + 2] = sl.take((long long)3 * L * 4);  // the true widths, uploaded once they are known
        if (po >= 0) cur = (const float*)(sl.base + po);
    }
    for (int l = 0; l < L; l++) {
        n_out[l] = cur_n;
        for (int b = 0; b < nb; b++) lens_out[l * nb + b] = cur_lens[b];
        {
            const long long o = sl.take((long long)nb * 4);
            offs[4 * L + l] = o;
            if (o >= 0 && (rc = upload_small(cur_lens.data(), (size_t)nb * 4, sl.base + o, stream)) != KP_OK) return rc;
        }
        if (sl.over) break;
        if (conv_r[l] > 0.f) {
            if (!cur_grid || cur_grid_r != conv_r[l]) {
                if ((rc = get_grid(cur, cur_n, cur_lens, conv_r[l], &cur_grid)) != KP_OK) return rc;
                cur_grid_r = conv_r[l];
            }
            if ((rc = search(l, 0, cur_grid, cur_n, conv_r[l], cur, cur_n, cur_lens, lim(l), n_cap ? n_cap[l] : cur_n,
                             n_cap ? n_cap[l] : cur_n, cur_grid)) != KP_OK) return rc;
        }
        if (l + 1 >= L || !(dl[l] > 0.f)) break;
        // next layer's points: room for cur_n rows now, trimmed to the voxel count once it is known (static mode: the
        // subsampling writes to scratch and the result is copied into its fixed, padded range)
        const long long po = sl.take((long long)(n_cap ? n_cap[l + 1] : cur_n) * 12);
        if (po < 0) {  // cannot continue without the next layer: report a generous size
            sl.off += (long long)cur_n * (3LL * cap * isz + 12) * 2;
            break;
        }
        float* next = (float*)(sl.base + po);
        float* sub_out = next;
        if (n_cap) {
            sub_out = S.alloc<float>((size_t)cur_n * 3);
            if (S.status != KP_OK) return S.status;
        }
        std::vector<int> next_lens(nb, 0);
        int m = 0;
        rc = grid_subsample_device(cur, cur_n, cur_lens.data(), nb, nullptr, 0, nullptr, 0, dl[l], 0, order,
                                   rot ? rot + (size_t)l * nb * 9 : nullptr, sub_out, next_lens.data(), nullptr, nullptr,
                                   &m, stream);
        if (rc != KP_OK) return rc;
        if (m <= 0) return fail(KP_ERR_EMPTY, "pyramid: a layer came out empty");
        if (n_cap) {
            if (m > n_cap[l + 1]) {
                KP_CUDA(cudaStreamSynchronize(stream));
                *need_cap = -(l + 2);
                return fail(KP_ERR_CAPACITY, "pyramid: a layer has more points than its static capacity");
            }
            if ((rc = pad_copy<float>(sub_out, (long long)m * 3, next, (long long)n_cap[l + 1] * 3, 1e6f, stream)) != KP_OK) return rc;
        } else {
            sl.off = po + (long long)m * 12;
        }
        offs[l + 1] = po;
        // pool: queries = next layer, supports = this layer
        const void* pool_grid = cur_grid;
        if (!pool_grid || cur_grid_r != pool_r[l]) {
            if ((rc = get_grid(cur, cur_n, cur_lens, pool_r[l], &pool_grid)) != KP_OK) return rc;
        }
        // the grid over the next layer (supports of the upsample search; the next layer's conv grid when radii agree) is
        // built first: it also gives the pool search its query order
        const void* up_grid = nullptr;
        if ((rc = get_grid(next, m, next_lens, up_r[l], &up_grid)) != KP_OK) return rc;
        if ((rc = search(l, 1, pool_grid, cur_n, pool_r[l], next, m, next_lens, lim(l), n_cap ? n_cap[l + 1] : m,
                         n_cap ? n_cap[l] : cur_n, up_grid)) != KP_OK) return rc;
        // upsample: queries = this layer, supports = next layer
        if ((rc = search(l, 2, up_grid, m, up_r[l], cur, cur_n, cur_lens, lim(l + 1), n_cap ? n_cap[l] : cur_n,
                         n_cap ? n_cap[l + 1] : m, pool_grid)) != KP_OK) return rc;
        cur = next; cur_n = m; cur_lens = next_lens;
        cur_grid = up_grid; cur_grid_r = up_r[l];
    }
    if (sl.over) {
        KP_CUDA(cudaStreamSynchronize(stream));
        *need_bytes = sl.off + (sl.off >> 2);
        return fail(KP_ERR_CAPACITY, "pyramid: output slab too small");
    }

    // one read-back for all searches
    std::vector<int> res((size_t)6 * L, 0);
    KP_CUDA(cudaMemcpyAsync(res.data(), d_results, (size_t)6 * L * sizeof(int), cudaMemcpyDeviceToHost, stream));
    KP_CUDA(cudaStreamSynchronize(stream));
    int want_cap = cap;
    for (auto& ps : pend) {
        int* r2 = &res[2 * (ps.kind * L + ps.layer)];
        if (r2[1] & 1) return fail(KP_ERR_UNSUPPORTED, "batch_query: cloud extent / radius exceeds 2^18 cells per axis");
        if (r2[1] & 2) return fail(KP_ERR_TOO_DENSE, "batch_query: more than 1024 neighbours for one query");
        if (r2[1] & 4) {  // a row outgrew the 256-hit staging of the fast kernel: redo this search synchronously
            int h = 0;
            rc = grid_query_device_ex(ps.grid, ps.ns, nb, ps.radius, ps.q, ps.nq, ps.qlens.data(), ps.out, idx_is_i64,
                                      ps.width, &h, nullptr, ps.shadow, ps.rows, stream);
            if (rc != KP_OK) return rc;
            r2[0] = h;
        }
        widths[ps.kind * L + ps.layer] = r2[0];
        if (!ps.limited && r2[0] > ps.width && r2[0] > want_cap) want_cap = r2[0];
    }
    if (want_cap > cap) {
        *need_cap = want_cap;
        return fail(KP_ERR_CAPACITY, "pyramid: neighbour rows wider than cap");
    }
    if (n_cap && offs[5 * L + 2] >= 0) {  // fixed-width consumers (max_pool) need the true widths on the device
        if ((rc = upload_small(widths, (size_t)3 * L * 4, sl.base + offs[5 * L + 2], stream)) != KP_OK) return rc;
        KP_CUDA(cudaStreamSynchronize(stream));
    }
    *need_bytes = sl.off;
    return KP_OK;
}

}  // namespace kp
