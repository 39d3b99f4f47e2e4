/* weasal_b200 — C ABI of the B200 (sm_100a) KPConv hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / numpy types. Each entry point names the
 * reference interface it replaces (paths relative to the WeaSAL repository root). INTEGRATION.md shows the
 * reference-side binding (the Python modules cpp_wrappers.cpp_neighbors.radius_neighbors,
 * cpp_wrappers.cpp_subsampling.grid_subsampling and the class models.blocks.KPConv) on top of these calls.
 *
 * Conventions
 *   - every function returns 0 on success or a negative kp_status; kp_last_error() returns the message of the
 *     calling thread's last failure;
 *   - "_host" entry points take HOST buffers and do their own host<->device copies (this is what the reference's
 *     numpy-facing extension modules bind); "_dev" entry points take DEVICE pointers plus a cudaStream_t passed as
 *     void* (what the torch-facing KPConv module and the device pyramid builder bind);
 *   - batch-length arrays (q_batches, s_batches, batches) and rotation matrices are always HOST pointers;
 *   - there is no CPU fallback: without a CUDA device every compute entry fails with KP_ERR_CUDA.
 */
#ifndef WEASAL_B200_H
#define WEASAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum kp_status {
    KP_OK = 0,
    KP_ERR_CUDA = -1,        /* CUDA runtime failure */
    KP_ERR_ARG = -2,         /* malformed argument */
    KP_ERR_CAPACITY = -3,    /* caller buffer too small */
    KP_ERR_TOO_DENSE = -4,   /* > 1024 neighbours for one query */
    KP_ERR_EMPTY = -5,       /* empty result; the reference raises RuntimeError("Error") here */
    KP_ERR_UNSUPPORTED = -6
};

const char* kp_last_error(void);
int kp_version(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
long long kp_launch_count(void);
void kp_free_host(void* p);
/* Per-kernel timing for bench.py's roofline leg: while enabled, the library brackets its main kernels with CUDA
 * events on the launching stream. kp_profile_read synchronises those events, writes "tag launches total_ms" lines
 * into buf (returns the length, -1 if buf is too small) and clears the records. */
void kp_profile_enable(int on);
int kp_profile_read(char* buf, int buflen);
/* Host-side planning only (no device work; exported so that the CPU tests can check it): over how many CTAs the KPConv /
 * unary kernels split the reduction axis of one 128-point tile, given the tiles, the 128-column chunks of that axis and
 * the CTAs the device holds at once (296 for the 8-warp kernels, 148 for the 16-warp ones). */
int kp_plan_ksplit(int n_tiles, int n_chunks, int slots);
/* Streams confined to a partition of the device's SMs (CUDA green context, driver API): `n_streams` non-blocking streams
 * of the given priority whose kernels run on a group of at least `min_sms` SMs (rounded up to the architecture's
 * granularity, 8 on sm_100) and on no others. The training step's prefetch stage (the reference's DataLoader workers,
 * datasets/Vaihingen3D_PseudoLabel.py:243-252) runs on such streams, so that its kernels do not take SM slots from the
 * training step on the rest of the device. streams_out receives cudaStream_t handles (they live as long as the process),
 * *sms_granted the partition's size. KP_ERR_UNSUPPORTED when the driver lacks the API. */
int kp_sm_partition_streams(int min_sms, int n_streams, int priority, void** streams_out, int* sms_granted);

/* ------------------------------------------------------------------------------------------------------------------
 * Batch radius search.
 * Replaces: cpp_wrappers/cpp_neighbors/wrapper.cpp:58-238 `batch_query(queries, supports, q_batches, s_batches,
 *           radius)` -> neighbors.cpp:211-332 batch_nanoflann_neighbors; called from datasets/common.py:185-196.
 * Result rows: same-batch supports with d2 < radius^2 (f32, unfused), sorted by (d2, support index) ascending,
 * global indices, padded with Ns.
 *
 * kp_batch_query_host: *out is malloc'd int32 [nq, *hmax] (free with kp_free_host), hmax = max neighbour count.
 *   nq == 0 or hmax == 0 returns KP_ERR_EMPTY like the reference's RuntimeError("Error") (wrapper.cpp:201-205).
 * kp_batch_query_dev: out is a device buffer [nq, cap] of int32 (out_is_i64 = 0) or int64 (= 1). Rows keep their
 *   `cap` closest neighbours (the crop datasets/common.py:336-346 applies afterwards); *hmax receives the true
 *   maximum count so the caller can slice [:, :min(hmax, cap)]. Synchronises the stream once.
 */
int kp_batch_query_host(const float* queries, int nq, const float* supports, int ns, const int* q_batches,
                        const int* s_batches, int nb, float radius, int** out, int* hmax);
int kp_batch_query_dev(const float* queries, int nq, const float* supports, int ns, const int* q_batches,
                       const int* s_batches, int nb, float radius, void* out, int out_is_i64, int cap, int* hmax,
                       void* stream);

/* Same search without the host synchronisation: d_result is a DEVICE int[2] that receives {true maximum neighbour
 * count, error bits (1 = extent/radius above 2^18 cells per axis, 2 = more than 1024 neighbours for one query,
 * 4 = a query had more than 256 neighbours: repeat the search with kp_batch_query_dev, which escalates itself)}.
 * The device pyramid builder issues all searches of a batch this way and reads the results back with one copy. */
int kp_batch_query_dev_async(const float* queries, int nq, const float* supports, int ns, const int* q_batches,
                             const int* s_batches, int nb, float radius, void* out, int out_is_i64, int cap,
                             int* d_result, void* stream);

/* Split form of the search, for callers that query the same supports at the same radius more than once (in the
 * pyramid of datasets/common.py:505-534 the grid over layer l+1 at radius 2r serves the upsample search of layer l
 * and the conv and pool searches of layer l+1). `grid` is a caller-owned DEVICE buffer of kp_search_grid_bytes(ns, nb)
 * bytes; build once, query any number of times with the same (ns, nb, radius). Query semantics are those of
 * kp_batch_query_dev (hmax != NULL: synchronising) / kp_batch_query_dev_async (d_result != NULL: not synchronising). */
long long kp_search_grid_bytes(int ns, int nb);
int kp_search_grid_build_dev(const float* supports, int ns, const int* s_batches, int nb, float radius, void* grid,
                             void* stream);
int kp_search_grid_query_dev(const void* grid, int ns, int nb, float radius, const float* queries, int nq,
                             const int* q_batches, void* out, int out_is_i64, int cap, int* hmax, int* d_result,
                             void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Grid subsampling.
 * Replaces: cpp_wrappers/cpp_subsampling/wrapper.cpp:62-333 `subsample_batch(points, batches, features, classes,
 *           sampleDl, method, max_p, verbose)` -> grid_subsampling.cpp:109-211, and wrapper.cpp:338-566
 *           `subsample(points, features, classes, sampleDl, method, verbose)` -> grid_subsampling.cpp:5-106
 *           (= nb 1, max_p 0); called from datasets/common.py:44-182.
 * order: 1 = the reference's output order (iteration order of libstdc++'s unordered_map, reproduced exactly for the
 *        libstdc++ this library is built against); 0 = first-occurrence order of the voxels (cheaper).
 * rot:   optional HOST [nb,3,3] f32 — the per-element random grid orientation datasets/common.py:89-135 applies
 *        around the call: points are rotated by R before voxelisation and barycentres by R^T afterwards, in the
 *        same f32 order numpy uses. NULL = no rotation.
 * features [n,fdim] / classes [n,ldim] may be NULL.
 *
 * _host: outputs are malloc'd ([m,3], [m,fdim], [m,ldim]); out_batches is a caller array [nb]; KP_ERR_EMPTY when
 *        no voxel is produced (wrapper.cpp:266-270).
 * _dev:  output buffers are caller-allocated device arrays with n rows; *m receives the voxel count; out_batches
 *        is a HOST array [nb]. Synchronises the stream once.
 */
int kp_grid_subsample_host(const float* points, int n, const int* batches, int nb, const float* features, int fdim,
                           const int* classes, int ldim, float sampleDl, int max_p, int order, const float* rot,
                           float** out_points, int* out_batches, float** out_features, int** out_classes, int* m);
int kp_grid_subsample_dev(const float* points, int n, const int* batches, int nb, const float* features, int fdim,
                          const int* classes, int ldim, float sampleDl, int max_p, int order, const float* rot,
                          float* out_points, int* out_batches, float* out_features, int* out_classes, int* m,
                          void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Pyramid builder: the whole per-layer walk of one batch in one call.
 * Replaces: datasets/common.py:461-577 `PointCloudDataset.segmentation_inputs` (the loop over config.architecture):
 *           per layer l the conv search (common.py:505), `batch_grid_subsampling` to layer l+1 incl. its random grid
 *           orientation (common.py:521, 89-135), the pool search (:531), the upsample search at twice the radius (:534)
 *           and the `big_neighborhood_filter` column crop (:336-346, 544-547).
 * points0 [n0,3] is a DEVICE pointer; every other input array is a HOST array:
 *   lengths0 [nb]; conv_radius / pool_radius / up_radius / sample_dl [n_layers] (conv_radius[l] == 0: no conv search at
 *   layer l; pool / upsample / sample_dl entries of the last layer are ignored); rot [n_layers-1][nb][3][3] f32 grid
 *   orientations or NULL; limits [n_layers] neighbourhood limits or NULL (0 = unlimited).
 * Outputs are carved from the caller's DEVICE `slab` (slab_bytes); offsets [5*n_layers + 3] receives byte offsets into it:
 *   [0,L) points of layer l ([n,3] f32; layer 0 = -1, the caller's points0), [L,2L) conv matrices [n_l, stride],
 *   [2L,3L) pool matrices [n_{l+1}, stride], [3L,4L) upsample matrices [n_l, stride], [4L,5L) batch lengths (int32
 *   [nb]); -1 = absent. n_out [L], lengths_out [L*nb], widths [3*L] (true maximum neighbour counts, conv / pool /
 *   upsample blocks of L) and strides [3*L] (row strides = limit, or cap when unlimited) are HOST arrays. Index
 *   matrices are int64 (idx_is_i64) or int32, rows sorted by (d2, index) and padded with the support count.
 * Returns KP_ERR_CAPACITY with *need_bytes (slab too small) or *need_cap (an unlimited search found rows wider than
 * cap) set to what a repeat call needs. Synchronises the stream; issues no Python-visible work, so it can run from a
 * prefetch thread on a side stream while another thread launches the network.
 */
int kp_pyramid_build_dev(const float* points0, int n0, const int* lengths0, int nb, int n_layers,
                         const float* conv_radius, const float* pool_radius, const float* up_radius,
                         const float* sample_dl, const float* rot, const int* limits, int order, int idx_is_i64, int cap,
                         void* slab, long long slab_bytes, long long* offsets, int* n_out, int* lengths_out, int* widths,
                         int* strides, long long* need_bytes, int* need_cap, void* stream);

/* Static-shape variant for consumers that bind their inputs once (a captured CUDA graph of the training step): layer l
 * always has n_cap[l] rows (HOST array [n_layers]). Points are padded with 1e6, index rows with the shadow value, which
 * is the SUPPORT layer's n_cap (so that "index == row count of the support tensor" still marks a shadow, the convention
 * of models/blocks.py:278/357), the optional layer-0 features [n0,fdim] f32 with 0 and labels [n0] int64 with
 * label_pad (DEVICE pointers or NULL; they travel in the slab so the consumer copies one buffer). Layer 0's points
 * live in the slab too (offsets[0] >= 0); offsets has 5*n_layers + 3 entries ([5L] features, [5L+1] labels, [5L+2] the 3*n_layers true widths as int32,
 * conv / pool / upsample blocks of L, for kp_max_pool_forward_width_dev). The slab
 * layout depends on (n_cap, limits / cap, nb, fdim) alone. Padded query rows have no neighbours and nothing refers to
 * a padded support row, so every operator of the path computes the same values on the real rows. A layer that outgrows
 * its capacity returns KP_ERR_CAPACITY with *need_cap = -(layer + 1). */
int kp_pyramid_build_static_dev(const float* points0, int n0, const int* lengths0, int nb, int n_layers,
                                const float* conv_radius, const float* pool_radius, const float* up_radius,
                                const float* sample_dl, const float* rot, const int* limits, int order, int idx_is_i64,
                                int cap, const int* n_cap, const float* features, int fdim, const long long* labels,
                                long long label_pad, void* slab, long long slab_bytes, long long* offsets, int* n_out,
                                int* lengths_out, int* widths, int* strides, long long* need_bytes, int* need_cap,
                                void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * KPConv (rigid, 'linear' influence, 'sum' aggregation), forward and backward.
 * Replaces: models/blocks.py:238-374 `KPConv.forward(q_pts, s_pts, neighb_inds, x)` and the autograd backward of
 *           that expression (gradients w.r.t. x and weights only; blocks.py:235-236).
 *   q_pts [nq,3] f32, s_pts [ns,3] f32, neighb_inds [nq,H] int32/int64 with row stride idx_stride (elements),
 *   shadow index == ns, x [ns,cin] f32, weights [K,cin,cout] f32, kernel_points [K,3] f32, out [nq,cout] f32.
 * All pointers are device pointers. The second contraction runs on tcgen05 tensor cores with TF32 operands
 * (rounded to nearest) and fp32 accumulation; the kernel-point-weighted gather is exact fp32.
 */
int kp_kpconv_forward_dev(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                          int idx_is_i64, int H, int idx_stride, const float* x, int cin, const float* weights,
                          int cout, const float* kernel_points, int K, float KP_extent, float* out, void* stream);
/* d_x [ns,cin] and d_weights [K,cin,cout] are overwritten (not accumulated). */
int kp_kpconv_backward_dev(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                           int idx_is_i64, int H, int idx_stride, const float* x, int cin, const float* weights,
                           int cout, const float* kernel_points, int K, float KP_extent, const float* d_out,
                           float* d_x, float* d_weights, void* stream);

/* Training variant: the forward pass leaves its influence lists (which depend only on the geometry and the kernel
 * points) in two caller-owned device buffers of the sizes kp_kpconv_lists_bytes reports (`lists_hdr`: 4 control ints +
 * one header of 272 ints per tile of 128 centres; `lists_entries`: (neighbour | row << 25, weight) records, worst case
 * 15 per table cell), and the backward pass of the same call reuses them instead of rebuilding them. Backward also
 * accepts the transposed neighbour table (kp_transpose_table_dev: CSR over the supports, rowptr int32 [ns+2] = ns+1 row
 * pointers followed by the longest row's length, col int32 [nq*H]), which depends on the index matrix only and can
 * therefore be shared by every KPConv that uses that matrix; NULL = build it internally.
 * Results are identical to the plain pair above. */
int kp_transpose_table_dev(const void* neighb_inds, int idx_is_i64, int nq, int H, int idx_stride, int ns,
                           int* rowptr, int* col, void* stream);
void kp_kpconv_lists_bytes(int nq, int H, long long* hdr_bytes, long long* entries_bytes);
int kp_kpconv_forward_keep_dev(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                               int idx_is_i64, int H, int idx_stride, const float* x, int cin, const float* weights,
                               int cout, const float* kernel_points, int K, float KP_extent, float* out,
                               void* lists_hdr, void* lists_entries, void* stream);
int kp_kpconv_backward_kept_dev(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                                int idx_is_i64, int H, int idx_stride, const float* x, int cin, const float* weights,
                                int cout, const float* kernel_points, int K, float KP_extent, const float* d_out,
                                float* d_x, float* d_weights, const void* lists_hdr, const void* lists_entries,
                                const int* t_rowptr, const int* t_col, void* stream);

/* Split form, for callers that know the geometry before the features (a training loop builds the lists of batch t+1 in
 * its prefetch stage, on a side stream, while batch t trains; KPConv's forward / backward of models/blocks.py:238-374
 * then starts at the tensor-core kernels):
 *   kp_kpconv_lists_build_dev  lists of one pass. Forward and dW: centres = q_pts, others = s_pts, the index matrix as
 *       given, kp_sign = +1. dX: centres = s_pts, others = q_pts, the TRANSPOSED table (t_rowptr / t_col from
 *       kp_transpose_table_dev with n_pairs = nq*H, or the matrix itself when it is symmetric), kp_sign = -1.
 *       entries_cap = capacity of lists_entries in records (the buffer itself must hold entries_cap + 2: the kernels fetch
 *       16-byte aligned windows); when a calibrated capacity is exceeded the overflow flag
 *       lists_hdr[1] is set (the affected tiles are left empty) and the caller must rebuild with a larger buffer.
 *   kp_kpconv_apply_lists_dev  out[nc, cout] = (gather of x[n_x_rows, cin] through the lists) x weights. Forward:
 *       transpose_w = 0, weights [K, cin, cout]. dX: x = d_out, cin = the conv's out_channels, cout = its in_channels,
 *       transpose_w = 1 and the conv's own weights [K, cout, cin]. weights_packed != 0: `weights` holds the operand
 *       images made by kp_pack_weights_dev (kind 0 / 1). out_slope != 1: LeakyReLU fused into the epilogue.
 *   kp_kpconv_dw_lists_dev     d_weights [K, cin, cout] = gather^T x d_out (overwritten), lists of the forward pass. */
int kp_kpconv_lists_build_dev(const float* centres, int nc, const float* others, int no, const void* neighb_inds,
                              int idx_is_i64, int H, int idx_stride, const int* t_rowptr, const int* t_col,
                              long long n_pairs, const float* kernel_points, int K, float kp_sign, float KP_extent,
                              void* lists_hdr, void* lists_entries, long long entries_cap, void* stream);
int kp_kpconv_apply_lists_dev(int nc, const float* x, int n_x_rows, int cin, const float* weights, int weights_packed,
                              int transpose_w, int cout, int K, const void* lists_hdr, const void* lists_entries,
                              float* out, float out_slope, void* stream);
int kp_kpconv_dw_lists_dev(int nq, const float* x, int ns, int cin, const float* d_out, int cout, int K,
                           const void* lists_hdr, const void* lists_entries, float* d_weights, void* stream);

/* All of a batch's geometry-only KPConv work in ONE call, for a prefetch thread: the jobs run in order on `stream`.
 *   kind 0  lists over a padded index matrix (centres, others, neighb_inds / idx_is_i64 / H / idx_stride, kernel points,
 *           kp_sign, KP_extent) -> (hdr, entries, entries_cap)        [kp_kpconv_lists_build_dev]
 *   kind 1  transposed table of a padded index matrix with nc rows over `no` supports -> (rowptr [no+2], col [nc*H])
 *                                                                    [kp_transpose_table_dev]
 *   kind 2  lists over a CSR table (rowptr, col of an earlier kind-1 job, n_pairs = its nc*H) -> (hdr, entries, cap)
 * overflow_flag (optional DEVICE int): set to 1 when any list outgrew its entries_cap. */
typedef struct kp_list_job {
    int kind;
    const float* centres; int nc;
    const float* others; int no;
    const void* neighb_inds; int idx_is_i64, H, idx_stride;
    int* rowptr; int* col; long long n_pairs;
    const float* kernel_points; int K; float kp_sign, KP_extent;
    void* hdr; void* entries; long long entries_cap;
} kp_list_job;
int kp_kpconv_prepare_dev(const kp_list_job* jobs, int n_jobs, int* overflow_flag, void* stream);

/* Operand images of the tensor-core contractions (TF32-rounded weights in the UMMA shared-memory layout, one image per
 * 64 reduction columns). One launch packs any number of them: a training step calls it once per step for every KPConv
 * and unary block instead of once per operator call. kinds: 0 KPConv forward (weights [K,cin,cout]), 1 KPConv dX (same
 * weights, read transposed), 2 linear forward (weight [cout,cin]), 3 linear dX (same weight). kinds / Ks / cins / couts /
 * weights / images are HOST arrays of n_jobs entries (device pointers inside); kp_pack_image_floats = image size. */
long long kp_pack_image_floats(int kind, int K, int cin, int cout);
int kp_pack_weights_dev(int n_jobs, const int* kinds, const float* const* weights, const int* Ks, const int* cins,
                        const int* couts, float* const* images, void* stream);

/* Backward for a SYMMETRIC neighbour table: queries == supports (pts [n,3]) and no row lost a neighbour to a crop, as
 * for the conv matrices `neighbors[l]` of datasets/common.py:505 when no neighbourhood limit bites. Then j is in row i
 * exactly when i is in row j (the f32 squared distance is exactly symmetric), the table is its own transpose and the
 * dX pass needs no transposed copy. The caller asserts the symmetry; results equal kp_kpconv_backward_dev's.
 * lists_hdr / lists_entries: the forward pass's lists (kp_kpconv_forward_keep_dev) or NULL. */
int kp_kpconv_backward_sym_dev(const float* pts, int n, const void* neighb_inds, int idx_is_i64, int H, int idx_stride,
                               const float* x, int cin, const float* weights, int cout, const float* kernel_points,
                               int K, float KP_extent, const float* d_out, float* d_x, float* d_weights,
                               const void* lists_hdr, const void* lists_entries, void* stream);

/* fp32 CUDA-core pieces (bring-up / cross-check of the tensor-core path; not the product path):
 *   wf [nq, K*cin] = kernel-point-weighted neighbour features; dx [ns,cin] += adjoint scatter of dwf [nq,K*cin]. */
int kp_kpconv_wf_dev(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds, int idx_is_i64,
                     int H, int idx_stride, const float* x, int cin, const float* kernel_points, int K,
                     float KP_extent, float* wf, void* stream);
int kp_kpconv_dx_atomic_dev(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                            int idx_is_i64, int H, int idx_stride, const float* dwf, int cin,
                            const float* kernel_points, int K, float KP_extent, float* dx, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Unary blocks next to KPConv: Linear (+ bias) + LeakyReLU as one tcgen05 kernel each way.
 * Replaces: models/blocks.py:467-507 `UnaryBlock.forward` = `nn.Linear(in, out, bias=False)` -> `BatchNormBlock`
 *           (blocks.py:430-465: the identity on 2-D features when use_bn, `x + bias` otherwise) -> `LeakyReLU(0.1)`,
 *           and its autograd backward.
 *   x [n,cin] f32, weight [cout,cin] f32 (nn.Linear layout), bias [cout] or NULL, y [n,cout];
 *   y = leaky_relu(x weight^T + bias, negative_slope); negative_slope = 1 means no activation.
 *   backward: g = d_y * (y > 0 ? 1 : negative_slope) with y the forward OUTPUT (NULL when there is no activation);
 *   d_x [n,cin] = g weight (may be NULL), d_weight [cout,cin] = g^T x (overwritten). The bias gradient (column sums
 *   of g) is left to the caller. TF32 operands rounded to nearest, fp32 accumulation, like the KPConv contraction. */
int kp_linear_forward_dev(const float* x, int n, int cin, const float* weight, const float* bias, int cout,
                          float negative_slope, float* y, void* stream);
int kp_linear_backward_dev(const float* x, int n, int cin, const float* weight, int cout, const float* y,
                           float negative_slope, const float* d_y, float* d_x, float* d_weight, void* stream);
/* The same with ready-made operand images (kp_pack_weights_dev kinds 2 / 3) and the two halves of the backward pass as
 * separate calls, so that a caller can run them on different streams. */
int kp_linear_forward_packed_dev(const float* x, int n, int cin, const float* images, const float* bias, int cout,
                                 float negative_slope, float* y, void* stream);
int kp_linear_dx_packed_dev(int n, int cin, const float* images_t, int cout, const float* y, float negative_slope,
                            const float* d_y, float* d_x, void* stream);
int kp_linear_dw_dev(const float* x, int n, int cin, int cout, const float* y, float negative_slope, const float* d_y,
                     float* d_weight, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Pooling gathers next to KPConv (the callers' side of the path).
 * Replaces: models/blocks.py:93-112 `max_pool(x, inds)` (shadow row = zeros, so a shadow entry contributes 0 to the
 *           max) and models/blocks.py:77-90 `closest_pool(x, inds)` (first column only), plus their adjoints.
 *   x [ns,channels] f32, inds [nq,H] int32/int64 (row stride idx_stride), out [nq,channels];
 *   argmax [nq,channels] int32 receives the winning support row (-1 = the shadow zero) for the backward pass.
 *   kp_closest_pool_dev: backward == 0: dst[nq,ch] = src[inds[:,0]];  backward != 0: dst[ns,ch] = scatter-add of src[nq,ch].
 */
int kp_max_pool_forward_dev(const float* x, int ns, int channels, const void* inds, int idx_is_i64, int nq, int H,
                            int idx_stride, float* out, int* argmax, void* stream);
/* Same with the matrix' TRUE width as a device scalar (d_width, int32): columns at or beyond *d_width are ignored. For
 * fixed-width (static-shape) matrices: the reference's matrix is only as wide as the batch's widest row
 * (cpp_neighbors/wrapper.cpp:211), so a full row has no shadow entry and no zero candidate in the maximum. */
int kp_max_pool_forward_width_dev(const float* x, int ns, int channels, const void* inds, int idx_is_i64, int nq, int H,
                                  int idx_stride, const int* d_width, float* out, int* argmax, void* stream);
int kp_max_pool_backward_dev(const float* d_out, const int* argmax, int nq, int channels, float* d_x, int ns,
                             void* stream);
int kp_closest_pool_dev(const float* src, int ns, int channels, const void* inds, int idx_is_i64, int nq,
                        int idx_stride, float* dst, int backward, void* stream);
/* Same with a row stride (in elements, >= channels) on src: a gradient that is a column slice of a wider matrix — the
 * backward of the decoder's torch.cat (models/architectures.py:339-340) — is read in place, not copied first. */
int kp_closest_pool_strided_dev(const float* src, int src_row_stride, int ns, int channels, const void* inds,
                                int idx_is_i64, int nq, int idx_stride, float* dst, int backward, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Either side of the network: input spheres and test-time votes.
 *
 * kp_extract_spheres_dev
 *   Replaces: datasets/Vaihingen3D_PseudoLabel.py:345-365 (`input_trees[c].query_radius(center_point, r=in_radius)`, then
 *             `(points[input_inds] - center_point).astype(np.float32)`), same code in the three sibling datasets.
 *   cloud [n,3] f32 DEVICE; centres [nb,3] f64 HOST (nb <= 16); membership like sklearn's KDTree: float64 distance <= radius.
 *   Outputs (DEVICE, capacity `cap` rows): centred points f32 [*,3] and cloud indices int64 [*], spheres stacked in centre
 *   order, rows of a sphere in ascending cloud index (sklearn returns tree order; only the row order differs).
 *   lengths [nb] HOST. KP_ERR_CAPACITY when the spheres hold more than `cap` points. Synchronises the stream.
 * kp_augment_spheres_dev
 *   Replaces: datasets/common.py:252-334 `augmentation_transform` (points only) and the feature assembly of
 *             Vaihingen3D_PseudoLabel.py:383, 423-430.
 *   out = (points . R_b) * scale_b + noise in float32, products and sums in numpy's order (no FMA); R [nb,3,3], scale [nb,3],
 *   lengths [nb], centre_z [nb], color_keep [nb] are HOST arrays drawn by the caller in the reference's RNG order; noise is a
 *   DEVICE [n,3] array or NULL. out_features [n,fdim] (or NULL) = [1, colors[inds] * color_keep, z_aug + centre_z, z_aug][:fdim]
 *   with colors a DEVICE [N_cloud, ncol] array and inds the indices kp_extract_spheres_dev returned.
 * kp_vote_update_dev
 *   Replaces: utils/tester_PseudoLabel.py:176-195. mode 0: test_probs[inds] = smooth * test_probs[inds] + (1 - smooth) * probs
 *   for the points within radius_limit of their sphere centre (radius_limit <= 0: all), spheres applied in order; mode 1:
 *   the order-independent accumulation used when spheres are sharded over GPUs (test_probs += probs, weight += 1).
 * kp_vote_reproject_dev
 *   Replaces: tester_PseudoLabel.py:270-283 (`test_probs[test_proj]`, argmax) and utils/metrics.py:35-118 `fast_confusion`
 *   for labels 0..C-1. weight = NULL for EMA votes; proj = NULL: identity; any of out_probs [m,C], out_pred [m] int32,
 *   (truth [m] int32, confusion [C,C] int64, rows = truth) may be NULL. */
int kp_extract_spheres_dev(const float* cloud, long long n, const double* centres, int nb, double radius,
                           float* out_points, long long* out_inds, long long cap, int* lengths, void* stream);
int kp_augment_spheres_dev(const float* points, const int* lengths, int nb, const float* R, const float* scale,
                           const float* noise, float* out_points, const float* colors, int ncol, const long long* inds,
                           const float* centre_z, const float* color_keep, float* out_features, int fdim, void* stream);
int kp_vote_update_dev(const float* probs, const float* points, const long long* inds, const int* lengths, int nb,
                       int n_classes, float radius_limit, float smooth, int mode, float* test_probs, float* weight,
                       void* stream);
int kp_vote_reproject_dev(const float* test_probs, const float* weight, const long long* proj, long long m, int n_classes,
                          float* out_probs, int* out_pred, const int* truth, long long* confusion, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WEASAL_B200_H */
