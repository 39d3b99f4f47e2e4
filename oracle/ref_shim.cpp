// TEST INFRASTRUCTURE — not product code.
//
// extern "C" shim over the UNMODIFIED reference C++ cores, compiled from the sources where they
// lie under /root/reference (see oracle/Makefile; output goes to oracle/_ref/, git-ignored).
// It wraps raw pointers into vector<PointXYZ> exactly the way the reference's CPython wrappers do
// (cpp_wrappers/cpp_neighbors/wrapper.cpp:184-198, cpp_wrappers/cpp_subsampling/wrapper.cpp:238-263)
// and calls:
//   batch_nanoflann_neighbors  (neighbors.cpp:211-332)  -> ref_batch_neighbors   (the wired-in path)
//   batch_ordered_neighbors    (neighbors.cpp:125-208)  -> ref_batch_ordered     (stable tie-break arbiter)
//   grid_subsampling           (grid_subsampling.cpp:5-106)    -> ref_subsample
//   batch_grid_subsampling     (grid_subsampling.cpp:109-211)  -> ref_subsample_batch
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// the resulting library.
#include "cpp_wrappers/cpp_neighbors/neighbors/neighbors.h"
#include "cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.h"
#include <cstdlib>
#include <cstring>

static vector<PointXYZ> wrap_points(const float* p, int n) {
    return vector<PointXYZ>((PointXYZ*)p, (PointXYZ*)p + n);
}

extern "C" {

void ref_free(void* p) { free(p); }

// Returns 0 on success; *out is malloc'd [nq, *hmax] int32 (caller frees with ref_free).
// An empty result returns 1, mirroring the wrapper's RuntimeError("Error") (wrapper.cpp:201-205).
int ref_batch_neighbors(const float* q, int nq, const float* s, int ns, const int* qb, const int* sb, int nb,
                        float radius, int** out, int* hmax) {
    vector<PointXYZ> queries = wrap_points(q, nq), supports = wrap_points(s, ns);
    vector<int> q_batches(qb, qb + nb), s_batches(sb, sb + nb), res;
    batch_nanoflann_neighbors(queries, supports, q_batches, s_batches, res, radius);
    if (res.size() < 1) return 1;
    *hmax = (int)(res.size() / (size_t)nq);
    *out = (int*)malloc(res.size() * sizeof(int));
    memcpy(*out, res.data(), res.size() * sizeof(int));
    return 0;
}

int ref_batch_ordered(const float* q, int nq, const float* s, int ns, const int* qb, const int* sb, int nb,
                      float radius, int** out, int* hmax) {
    vector<PointXYZ> queries = wrap_points(q, nq), supports = wrap_points(s, ns);
    vector<int> q_batches(qb, qb + nb), s_batches(sb, sb + nb), res;
    batch_ordered_neighbors(queries, supports, q_batches, s_batches, res, radius);
    if (res.size() < 1) return 1;
    *hmax = (int)(res.size() / (size_t)nq);
    *out = (int*)malloc(res.size() * sizeof(int));
    memcpy(*out, res.data(), res.size() * sizeof(int));
    return 0;
}

// feats / classes may be NULL (fdim / ldim then ignored). Outputs malloc'd; *nout = number of voxels.
int ref_subsample(const float* p, int n, const float* feats, int fdim, const int* classes, int ldim, float dl,
                  float** out_p, float** out_f, int** out_c, int* nout) {
    vector<PointXYZ> pts = wrap_points(p, n), sp;
    vector<float> f, sf;
    vector<int> c, sc;
    if (feats) f.assign(feats, feats + (size_t)n * fdim);
    if (classes) c.assign(classes, classes + (size_t)n * ldim);
    grid_subsampling(pts, sp, f, sf, c, sc, dl, 0);
    if (sp.size() < 1) return 1;
    *nout = (int)sp.size();
    *out_p = (float*)malloc(sp.size() * 3 * sizeof(float));
    memcpy(*out_p, sp.data(), sp.size() * 3 * sizeof(float));
    if (feats) { *out_f = (float*)malloc(sf.size() * sizeof(float)); memcpy(*out_f, sf.data(), sf.size() * sizeof(float)); }
    if (classes) { *out_c = (int*)malloc(sc.size() * sizeof(int)); memcpy(*out_c, sc.data(), sc.size() * sizeof(int)); }
    return 0;
}

int ref_subsample_batch(const float* p, int n, const int* batches, int nb, const float* feats, int fdim,
                        const int* classes, int ldim, float dl, int max_p,
                        float** out_p, int* out_b, float** out_f, int** out_c, int* nout) {
    vector<PointXYZ> pts = wrap_points(p, n), sp;
    vector<float> f, sf;
    vector<int> c, sc, b(batches, batches + nb), sb;
    if (feats) f.assign(feats, feats + (size_t)n * fdim);
    if (classes) c.assign(classes, classes + (size_t)n * ldim);
    batch_grid_subsampling(pts, sp, f, sf, c, sc, b, sb, dl, max_p);
    if (sp.size() < 1) return 1;
    *nout = (int)sp.size();
    *out_p = (float*)malloc(sp.size() * 3 * sizeof(float));
    memcpy(*out_p, sp.data(), sp.size() * 3 * sizeof(float));
    memcpy(out_b, sb.data(), nb * sizeof(int));
    if (feats) { *out_f = (float*)malloc(sf.size() * sizeof(float)); memcpy(*out_f, sf.data(), sf.size() * sizeof(float)); }
    if (classes) { *out_c = (int*)malloc(sc.size() * sizeof(int)); memcpy(*out_c, sc.data(), sc.size() * sizeof(int)); }
    return 0;
}

}  // extern "C"
