// extern "C" surface of libweasal_b200.so — see include/weasal_b200.h for the contract of every entry point.
#include "../../include/weasal_b200.h"
#include "common.cuh"

#include <cstdlib>
#include <vector>

namespace kp {
int batch_query_device(const float* q, int nq, const float* s, int ns, const int* qb_host, const int* sb_host, int nb,
                       float radius, void* out, int out_is_i64, int cap, int* hmax_host, int* d_result,
                       cudaStream_t stream);
size_t grid_bytes(int ns, int nb);
int grid_build_device(const float* s, int ns, const int* sb_host, int nb, float radius, void* grid_buf, cudaStream_t stream);
int grid_query_device(const void* grid_buf, int ns, int nb, float radius, const float* q, int nq, const int* qb_host,
                      void* out, int out_is_i64, int cap, int* hmax_host, int* d_result, cudaStream_t stream);
int grid_subsample_device(const float* pts, int n, const int* lens_host, int nb, const float* feats, int fdim,
                          const int* classes, int ldim, float dl, int max_p, int order_mode, const float* rot_host,
                          float* out_pts, int* out_lens_host, float* out_feats, int* out_classes, int* m_host,
                          cudaStream_t stream);
int kpconv_wf_device(const float* q, int nq, const float* s, int ns, const void* idx, int idx_is_i64, int H,
                     int idx_stride, const float* x, int cin, const float* kp, int K, float extent, float* wf,
                     cudaStream_t stream);
int kpconv_dx_atomic_device(const float* q, int nq, const float* s, int ns, const void* idx, int idx_is_i64, int H,
                            int idx_stride, const float* dwf, int cin, const float* kp, int K, float extent, float* dx,
                            cudaStream_t stream);
void kpconv_lists_bytes(int nc, long long n_pairs, long long* hdr_bytes, long long* entries_bytes);
int kpconv_forward_device(const float* q, int nq, const float* s, int ns, const void* idx, int idx_is_i64, int H,
                          int idx_stride, const float* x, int cin, const float* w, int cout, const float* kp, int K,
                          float extent, float* out, void* lists_hdr, void* lists_entries, cudaStream_t stream);
int kpconv_backward_device(const float* q, int nq, const float* s, int ns, const void* idx, int idx_is_i64, int H,
                           int idx_stride, const float* x, int cin, const float* w, int cout, const float* kp, int K,
                           float extent, const float* dout, float* dx, float* dw, const void* lists_hdr,
                           const void* lists_entries, const int* t_rowptr, const int* t_col, int table_symmetric,
                           cudaStream_t stream);
int kpconv_lists_build_device(const float* centres, int nc, const float* others, int no, const void* idx, int idx_is_i64,
                              int H, int idx_stride, const int* rowptr, const int* col, long long n_pairs,
                              const float* kp, int K, float kp_sign, float extent, void* hdr, void* entries,
                              long long entries_cap, cudaStream_t stream);
int kpconv_apply_lists_device(int nc, const float* x, int n_x_rows, int cin, const float* w, int w_packed, int transpose_w,
                              int cout, int K, const void* hdr, const void* entries, float* out, float slope_out,
                              cudaStream_t stream);
int kpconv_dw_lists_device(int nq, const float* x, int ns, int cin, const float* dout, int cout, int K, const void* hdr,
                           const void* entries, float* dw, cudaStream_t stream);
long long pack_image_floats(int kind, int K, int cin, int cout);
int kpconv_prepare_device(const kp_list_job* jobs, int n_jobs, int* overflow_flag, cudaStream_t stream);
int pack_weights_device(int n_jobs, const int* kinds, const float* const* weights, const int* Ks, const int* cins,
                        const int* couts, float* const* images, cudaStream_t stream);
int transpose_table_entry(const void* idx, int idx_is_i64, int nq, int H, int idx_stride, int ns, int* rowptr,
                          int* col_sorted, cudaStream_t stream);
int profile_read(char* buf, int buflen);
int plan_ksplit(int n_tiles, int n_chunks, int slots);
int sm_partition_streams(int min_sms, int n_streams, int priority, void** streams_out, int* sms_granted);
int linear_forward_device(const float* x, int n, int cin, const float* w, int w_packed, const float* bias, int cout,
                          float slope, float* y, cudaStream_t stream);
int linear_backward_device(const float* x, int n, int cin, const float* w, int w_packed, int cout, const float* y,
                           float slope, const float* dy, float* dx, float* dw, cudaStream_t stream);
int pyramid_build_device(const float* pts0, int n0, const int* lens0, int nb, int L, const float* conv_r,
                         const float* pool_r, const float* up_r, const float* dl, const float* rot, const int* limits,
                         int order, int idx_is_i64, int cap, void* slab, long long slab_bytes, long long* offs, int* n_out,
                         int* lens_out, int* widths, int* strides, long long* need_bytes, int* need_cap,
                         const int* n_cap, const float* feats, int fdim, const long long* labels, long long label_pad,
                         cudaStream_t stream);
int extract_spheres_device(const float* cloud, long long n, const double* centres_host, int nb, double radius,
                           float* out_pts, long long* out_inds, long long cap, int* lengths_host, cudaStream_t stream);
int augment_device(const float* pts, const int* lengths_host, int nb, const float* R_host, const float* scale_host,
                   const float* noise, float* out, const float* colors, int ncol, const long long* inds,
                   const float* centre_z_host, const float* keep_host, float* feats, int fdim, cudaStream_t stream);
int vote_device(const float* probs, const float* pts, const long long* inds, const int* lengths_host, int nb, int C,
                float radius_limit, float smooth, int mode, float* acc, float* weight, cudaStream_t stream);
int reproject_device(const float* acc, const float* weight, const long long* proj, long long m, int C, float* out_probs,
                     int* out_pred, const int* truth, long long* conf, cudaStream_t stream);
int max_pool_fwd_device(const float* x, int ns, int C, const void* idx, int is_i64, int nq, int H, int stride,
                        float* out, int* arg, const int* d_width, cudaStream_t stream);
int max_pool_bwd_device(const float* dout, const int* arg, int nq, int C, float* dx, int ns, cudaStream_t stream);
int closest_pool_device(const float* src, int ns, int C, const void* idx, int is_i64, int nq, int stride, float* dst,
                        int backward, int src_ld, cudaStream_t stream);
}  // namespace kp

using namespace kp;

namespace {
struct DevBuf {  // synchronous device buffer for the host-facing entry points
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        KP_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
        return KP_OK;
    }
};
}  // namespace

extern "C" {

const char* kp_last_error(void) { return g_last_error.c_str(); }
int kp_version(void) { return 100; }
long long kp_launch_count(void) { return g_launch_count.load(); }
void kp_free_host(void* p) { free(p); }
void kp_profile_enable(int on) { g_profile_on = on != 0; }
int kp_profile_read(char* buf, int buflen) { return profile_read(buf, buflen); }
int kp_sm_partition_streams(int min_sms, int n_streams, int priority, void** streams_out, int* sms_granted) {
    return kp::sm_partition_streams(min_sms, n_streams, priority, streams_out, sms_granted);
}

int kp_plan_ksplit(int n_tiles, int n_chunks, int slots) {
    if (n_tiles <= 0 || n_chunks <= 0 || slots <= 0) return fail(KP_ERR_ARG, "plan_ksplit: bad sizes");
    return plan_ksplit(n_tiles, n_chunks, slots);
}

int kp_batch_query_dev(const float* queries, int nq, const float* supports, int ns, const int* q_batches,
                       const int* s_batches, int nb, float radius, void* out, int out_is_i64, int cap, int* hmax,
                       void* stream) {
    return batch_query_device(queries, nq, supports, ns, q_batches, s_batches, nb, radius, out, out_is_i64, cap, hmax,
                              nullptr, (cudaStream_t)stream);
}

int kp_batch_query_dev_async(const float* queries, int nq, const float* supports, int ns, const int* q_batches,
                             const int* s_batches, int nb, float radius, void* out, int out_is_i64, int cap,
                             int* d_result, void* stream) {
    if (!d_result) return fail(KP_ERR_ARG, "batch_query_async: d_result is required");
    return batch_query_device(queries, nq, supports, ns, q_batches, s_batches, nb, radius, out, out_is_i64, cap, nullptr,
                              d_result, (cudaStream_t)stream);
}

int kp_batch_query_host(const float* queries, int nq, const float* supports, int ns, const int* q_batches,
                        const int* s_batches, int nb, float radius, int** out, int* hmax) {
    *out = nullptr;
    *hmax = 0;
    if (nq <= 0) return fail(KP_ERR_EMPTY, "Error");
    DevBuf dq, ds, dout;
    int rc;
    if ((rc = dq.alloc((size_t)nq * 12)) || (rc = ds.alloc((size_t)ns * 12))) return rc;
    KP_CUDA(cudaMemcpy(dq.p, queries, (size_t)nq * 12, cudaMemcpyHostToDevice));
    if (ns > 0) KP_CUDA(cudaMemcpy(ds.p, supports, (size_t)ns * 12, cudaMemcpyHostToDevice));
    int cap = 64;
    while (true) {
        if ((rc = dout.alloc((size_t)nq * cap * sizeof(int)))) return rc;
        rc = batch_query_device((const float*)dq.p, nq, (const float*)ds.p, ns, q_batches, s_batches, nb, radius, dout.p,
                                0, cap, hmax, nullptr, 0);
        if (rc != KP_OK) return rc;
        if (*hmax <= cap) break;
        cap = *hmax;  // rare: a row was wider than the first guess, redo with the exact width
        cudaFree(dout.p);
        dout.p = nullptr;
    }
    if (*hmax == 0) return fail(KP_ERR_EMPTY, "Error");
    *out = (int*)malloc((size_t)nq * (*hmax) * sizeof(int));
    KP_CUDA(cudaMemcpy2D(*out, (size_t)(*hmax) * sizeof(int), dout.p, (size_t)cap * sizeof(int),
                         (size_t)(*hmax) * sizeof(int), nq, cudaMemcpyDeviceToHost));
    return KP_OK;
}

long long kp_search_grid_bytes(int ns, int nb) { return (long long)grid_bytes(ns, nb); }

int kp_search_grid_build_dev(const float* supports, int ns, const int* s_batches, int nb, float radius, void* grid,
                             void* stream) {
    return grid_build_device(supports, ns, s_batches, nb, radius, grid, (cudaStream_t)stream);
}

int kp_search_grid_query_dev(const void* grid, int ns, int nb, float radius, const float* queries, int nq,
                             const int* q_batches, void* out, int out_is_i64, int cap, int* hmax, int* d_result,
                             void* stream) {
    if (!hmax && !d_result) return fail(KP_ERR_ARG, "search_grid_query: hmax or d_result is required");
    return grid_query_device(grid, ns, nb, radius, queries, nq, q_batches, out, out_is_i64, cap, hmax, d_result,
                             (cudaStream_t)stream);
}

int kp_grid_subsample_dev(const float* points, int n, const int* batches, int nb, const float* features, int fdim,
                          const int* classes, int ldim, float sampleDl, int max_p, int order, const float* rot,
                          float* out_points, int* out_batches, float* out_features, int* out_classes, int* m,
                          void* stream) {
    return grid_subsample_device(points, n, batches, nb, features, fdim, classes, ldim, sampleDl, max_p, order, rot,
                                 out_points, out_batches, out_features, out_classes, m, (cudaStream_t)stream);
}

int kp_grid_subsample_host(const float* points, int n, const int* batches, int nb, const float* features, int fdim,
                           const int* classes, int ldim, float sampleDl, int max_p, int order, const float* rot,
                           float** out_points, int* out_batches, float** out_features, int** out_classes, int* m) {
    *out_points = nullptr;
    if (out_features) *out_features = nullptr;
    if (out_classes) *out_classes = nullptr;
    *m = 0;
    if (n <= 0) return fail(KP_ERR_EMPTY, "Error");
    DevBuf dp, df, dc, op, of, oc;
    int rc;
    if ((rc = dp.alloc((size_t)n * 12)) || (rc = op.alloc((size_t)n * 12))) return rc;
    KP_CUDA(cudaMemcpy(dp.p, points, (size_t)n * 12, cudaMemcpyHostToDevice));
    if (features) {
        if ((rc = df.alloc((size_t)n * fdim * 4)) || (rc = of.alloc((size_t)n * fdim * 4))) return rc;
        KP_CUDA(cudaMemcpy(df.p, features, (size_t)n * fdim * 4, cudaMemcpyHostToDevice));
    }
    if (classes) {
        if ((rc = dc.alloc((size_t)n * ldim * 4)) || (rc = oc.alloc((size_t)n * ldim * 4))) return rc;
        KP_CUDA(cudaMemcpy(dc.p, classes, (size_t)n * ldim * 4, cudaMemcpyHostToDevice));
    }
    rc = grid_subsample_device((const float*)dp.p, n, batches, nb, (const float*)df.p, fdim, (const int*)dc.p, ldim,
                               sampleDl, max_p, order, rot, (float*)op.p, out_batches, (float*)of.p, (int*)oc.p, m, 0);
    if (rc != KP_OK) return rc;
    if (*m <= 0) return fail(KP_ERR_EMPTY, "Error");
    *out_points = (float*)malloc((size_t)(*m) * 12);
    KP_CUDA(cudaMemcpy(*out_points, op.p, (size_t)(*m) * 12, cudaMemcpyDeviceToHost));
    if (features) {
        *out_features = (float*)malloc((size_t)(*m) * fdim * 4);
        KP_CUDA(cudaMemcpy(*out_features, of.p, (size_t)(*m) * fdim * 4, cudaMemcpyDeviceToHost));
    }
    if (classes) {
        *out_classes = (int*)malloc((size_t)(*m) * ldim * 4);
        KP_CUDA(cudaMemcpy(*out_classes, oc.p, (size_t)(*m) * ldim * 4, cudaMemcpyDeviceToHost));
    }
    return KP_OK;
}

int kp_kpconv_wf_dev(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds, int idx_is_i64,
                     int H, int idx_stride, const float* x, int cin, const float* kernel_points, int K,
                     float KP_extent, float* wf, void* stream) {
    return kpconv_wf_device(q_pts, nq, s_pts, ns, neighb_inds, idx_is_i64, H, idx_stride, x, cin, kernel_points, K,
                            KP_extent, wf, (cudaStream_t)stream);
}

int kp_kpconv_dx_atomic_dev(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                            int idx_is_i64, int H, int idx_stride, const float* dwf, int cin,
                            const float* kernel_points, int K, float KP_extent, float* dx, void* stream) {
    return kpconv_dx_atomic_device(q_pts, nq, s_pts, ns, neighb_inds, idx_is_i64, H, idx_stride, dwf, cin,
                                   kernel_points, K, KP_extent, dx, (cudaStream_t)stream);
}

int kp_kpconv_forward_dev(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                          int idx_is_i64, int H, int idx_stride, const float* x, int cin, const float* weights,
                          int cout, const float* kernel_points, int K, float KP_extent, float* out, void* stream) {
    return kpconv_forward_device(q_pts, nq, s_pts, ns, neighb_inds, idx_is_i64, H, idx_stride, x, cin, weights, cout,
                                 kernel_points, K, KP_extent, out, nullptr, nullptr, (cudaStream_t)stream);
}

void kp_kpconv_lists_bytes(int nq, int H, long long* hdr_bytes, long long* entries_bytes) {
    kpconv_lists_bytes(nq, (long long)nq * H, hdr_bytes, entries_bytes);
}

int kp_kpconv_lists_build_dev(const float* centres, int nc, const float* others, int no, const void* neighb_inds,
                              int idx_is_i64, int H, int idx_stride, const int* t_rowptr, const int* t_col,
                              long long n_pairs, const float* kernel_points, int K, float kp_sign, float KP_extent,
                              void* lists_hdr, void* lists_entries, long long entries_cap, void* stream) {
    return kpconv_lists_build_device(centres, nc, others, no, neighb_inds, idx_is_i64, H, idx_stride, t_rowptr, t_col,
                                     n_pairs, kernel_points, K, kp_sign, KP_extent, lists_hdr, lists_entries, entries_cap,
                                     (cudaStream_t)stream);
}

int kp_kpconv_apply_lists_dev(int nc, const float* x, int n_x_rows, int cin, const float* weights, int weights_packed,
                              int transpose_w, int cout, int K, const void* lists_hdr, const void* lists_entries,
                              float* out, float out_slope, void* stream) {
    return kpconv_apply_lists_device(nc, x, n_x_rows, cin, weights, weights_packed, transpose_w, cout, K, lists_hdr,
                                     lists_entries, out, out_slope, (cudaStream_t)stream);
}

int kp_kpconv_dw_lists_dev(int nq, const float* x, int ns, int cin, const float* d_out, int cout, int K,
                           const void* lists_hdr, const void* lists_entries, float* d_weights, void* stream) {
    return kpconv_dw_lists_device(nq, x, ns, cin, d_out, cout, K, lists_hdr, lists_entries, d_weights, (cudaStream_t)stream);
}

long long kp_pack_image_floats(int kind, int K, int cin, int cout) { return pack_image_floats(kind, K, cin, cout); }

int kp_kpconv_prepare_dev(const kp_list_job* jobs, int n_jobs, int* overflow_flag, void* stream) {
    return kpconv_prepare_device(jobs, n_jobs, overflow_flag, (cudaStream_t)stream);
}

int kp_pack_weights_dev(int n_jobs, const int* kinds, const float* const* weights, const int* Ks, const int* cins,
                        const int* couts, float* const* images, void* stream) {
    if (n_jobs < 0 || (n_jobs > 0 && (!kinds || !weights || !Ks || !cins || !couts || !images)))
        return fail(KP_ERR_ARG, "pack_weights: bad arguments");
    return pack_weights_device(n_jobs, kinds, weights, Ks, cins, couts, images, (cudaStream_t)stream);
}

int kp_kpconv_forward_keep_dev(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                               int idx_is_i64, int H, int idx_stride, const float* x, int cin, const float* weights,
                               int cout, const float* kernel_points, int K, float KP_extent, float* out,
                               void* lists_hdr, void* lists_entries, void* stream) {
    return kpconv_forward_device(q_pts, nq, s_pts, ns, neighb_inds, idx_is_i64, H, idx_stride, x, cin, weights, cout,
                                 kernel_points, K, KP_extent, out, lists_hdr, lists_entries, (cudaStream_t)stream);
}

int kp_kpconv_backward_kept_dev(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                                int idx_is_i64, int H, int idx_stride, const float* x, int cin, const float* weights,
                                int cout, const float* kernel_points, int K, float KP_extent, const float* d_out,
                                float* d_x, float* d_weights, const void* lists_hdr, const void* lists_entries,
                                const int* t_rowptr, const int* t_col, void* stream) {
    return kpconv_backward_device(q_pts, nq, s_pts, ns, neighb_inds, idx_is_i64, H, idx_stride, x, cin, weights, cout,
                                  kernel_points, K, KP_extent, d_out, d_x, d_weights, lists_hdr, lists_entries,
                                  t_rowptr, t_col, 0, (cudaStream_t)stream);
}

int kp_kpconv_backward_sym_dev(const float* pts, int n, const void* neighb_inds, int idx_is_i64, int H, int idx_stride,
                               const float* x, int cin, const float* weights, int cout, const float* kernel_points,
                               int K, float KP_extent, const float* d_out, float* d_x, float* d_weights,
                               const void* lists_hdr, const void* lists_entries, void* stream) {
    return kpconv_backward_device(pts, n, pts, n, neighb_inds, idx_is_i64, H, idx_stride, x, cin, weights, cout,
                                  kernel_points, K, KP_extent, d_out, d_x, d_weights, lists_hdr, lists_entries, nullptr,
                                  nullptr, 1, (cudaStream_t)stream);
}

int kp_transpose_table_dev(const void* neighb_inds, int idx_is_i64, int nq, int H, int idx_stride, int ns,
                           int* rowptr, int* col, void* stream) {
    return transpose_table_entry(neighb_inds, idx_is_i64, nq, H, idx_stride, ns, rowptr, col, (cudaStream_t)stream);
}

int kp_kpconv_backward_dev(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                           int idx_is_i64, int H, int idx_stride, const float* x, int cin, const float* weights,
                           int cout, const float* kernel_points, int K, float KP_extent, const float* d_out,
                           float* d_x, float* d_weights, void* stream) {
    return kpconv_backward_device(q_pts, nq, s_pts, ns, neighb_inds, idx_is_i64, H, idx_stride, x, cin, weights, cout,
                                  kernel_points, K, KP_extent, d_out, d_x, d_weights, nullptr, nullptr, nullptr, nullptr,
                                  0, (cudaStream_t)stream);
}

int kp_pyramid_build_dev(const float* points0, int n0, const int* lengths0, int nb, int n_layers,
                         const float* conv_radius, const float* pool_radius, const float* up_radius,
                         const float* sample_dl, const float* rot, const int* limits, int order, int idx_is_i64, int cap,
                         void* slab, long long slab_bytes, long long* offsets, int* n_out, int* lengths_out, int* widths,
                         int* strides, long long* need_bytes, int* need_cap, void* stream) {
    return pyramid_build_device(points0, n0, lengths0, nb, n_layers, conv_radius, pool_radius, up_radius, sample_dl, rot,
                                limits, order, idx_is_i64, cap, slab, slab_bytes, offsets, n_out, lengths_out, widths,
                                strides, need_bytes, need_cap, nullptr, nullptr, 0, nullptr, 0, (cudaStream_t)stream);
}

int kp_pyramid_build_static_dev(const float* points0, int n0, const int* lengths0, int nb, int n_layers,
                                const float* conv_radius, const float* pool_radius, const float* up_radius,
                                const float* sample_dl, const float* rot, const int* limits, int order, int idx_is_i64,
                                int cap, const int* n_cap, const float* features, int fdim, const long long* labels,
                                long long label_pad, void* slab, long long slab_bytes, long long* offsets, int* n_out,
                                int* lengths_out, int* widths, int* strides, long long* need_bytes, int* need_cap,
                                void* stream) {
    if (!n_cap) return fail(KP_ERR_ARG, "pyramid_build_static: n_cap is required");
    return pyramid_build_device(points0, n0, lengths0, nb, n_layers, conv_radius, pool_radius, up_radius, sample_dl, rot,
                                limits, order, idx_is_i64, cap, slab, slab_bytes, offsets, n_out, lengths_out, widths,
                                strides, need_bytes, need_cap, n_cap, features, fdim, labels, label_pad,
                                (cudaStream_t)stream);
}

int kp_linear_forward_dev(const float* x, int n, int cin, const float* weight, const float* bias, int cout,
                          float negative_slope, float* y, void* stream) {
    return linear_forward_device(x, n, cin, weight, 0, bias, cout, negative_slope, y, (cudaStream_t)stream);
}
int kp_linear_forward_packed_dev(const float* x, int n, int cin, const float* images, const float* bias, int cout,
                                 float negative_slope, float* y, void* stream) {
    return linear_forward_device(x, n, cin, images, 1, bias, cout, negative_slope, y, (cudaStream_t)stream);
}
int kp_linear_backward_dev(const float* x, int n, int cin, const float* weight, int cout, const float* y,
                           float negative_slope, const float* d_y, float* d_x, float* d_weight, void* stream) {
    return linear_backward_device(x, n, cin, weight, 0, cout, y, negative_slope, d_y, d_x, d_weight, (cudaStream_t)stream);
}
int kp_linear_dx_packed_dev(int n, int cin, const float* images_t, int cout, const float* y, float negative_slope,
                            const float* d_y, float* d_x, void* stream) {
    return linear_backward_device(nullptr, n, cin, images_t, 1, cout, y, negative_slope, d_y, d_x, nullptr, (cudaStream_t)stream);
}
int kp_linear_dw_dev(const float* x, int n, int cin, int cout, const float* y, float negative_slope, const float* d_y,
                     float* d_weight, void* stream) {
    return linear_backward_device(x, n, cin, nullptr, 0, cout, y, negative_slope, d_y, nullptr, d_weight, (cudaStream_t)stream);
}

int kp_max_pool_forward_dev(const float* x, int ns, int channels, const void* inds, int idx_is_i64, int nq, int H,
                            int idx_stride, float* out, int* argmax, void* stream) {
    return max_pool_fwd_device(x, ns, channels, inds, idx_is_i64, nq, H, idx_stride, out, argmax, nullptr,
                               (cudaStream_t)stream);
}
int kp_max_pool_forward_width_dev(const float* x, int ns, int channels, const void* inds, int idx_is_i64, int nq, int H,
                                  int idx_stride, const int* d_width, float* out, int* argmax, void* stream) {
    return max_pool_fwd_device(x, ns, channels, inds, idx_is_i64, nq, H, idx_stride, out, argmax, d_width,
                               (cudaStream_t)stream);
}
int kp_max_pool_backward_dev(const float* d_out, const int* argmax, int nq, int channels, float* d_x, int ns,
                             void* stream) {
    return max_pool_bwd_device(d_out, argmax, nq, channels, d_x, ns, (cudaStream_t)stream);
}
int kp_closest_pool_dev(const float* src, int ns, int channels, const void* inds, int idx_is_i64, int nq,
                        int idx_stride, float* dst, int backward, void* stream) {
    return closest_pool_device(src, ns, channels, inds, idx_is_i64, nq, idx_stride, dst, backward, channels,
                               (cudaStream_t)stream);
}
int kp_closest_pool_strided_dev(const float* src, int src_row_stride, int ns, int channels, const void* inds,
                                int idx_is_i64, int nq, int idx_stride, float* dst, int backward, void* stream) {
    return closest_pool_device(src, ns, channels, inds, idx_is_i64, nq, idx_stride, dst, backward, src_row_stride,
                               (cudaStream_t)stream);
}

int kp_extract_spheres_dev(const float* cloud, long long n, const double* centres, int nb, double radius,
                           float* out_points, long long* out_inds, long long cap, int* lengths, void* stream) {
    return extract_spheres_device(cloud, n, centres, nb, radius, out_points, out_inds, cap, lengths, (cudaStream_t)stream);
}
int kp_augment_spheres_dev(const float* points, const int* lengths, int nb, const float* R, const float* scale,
                           const float* noise, float* out_points, const float* colors, int ncol, const long long* inds,
                           const float* centre_z, const float* color_keep, float* out_features, int fdim, void* stream) {
    return augment_device(points, lengths, nb, R, scale, noise, out_points, colors, ncol, inds, centre_z, color_keep,
                          out_features, fdim, (cudaStream_t)stream);
}
int kp_vote_update_dev(const float* probs, const float* points, const long long* inds, const int* lengths, int nb,
                       int n_classes, float radius_limit, float smooth, int mode, float* test_probs, float* weight,
                       void* stream) {
    return vote_device(probs, points, inds, lengths, nb, n_classes, radius_limit, smooth, mode, test_probs, weight,
                       (cudaStream_t)stream);
}
int kp_vote_reproject_dev(const float* test_probs, const float* weight, const long long* proj, long long m, int n_classes,
                          float* out_probs, int* out_pred, const int* truth, long long* confusion, void* stream) {
    return reproject_device(test_probs, weight, proj, m, n_classes, out_probs, out_pred, truth, confusion, (cudaStream_t)stream);
}

}  // extern "C"
