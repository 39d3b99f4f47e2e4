// The two stages either side of the network that the reference runs on the CPU per batch:
//
//   input side   potential_item of datasets/Vaihingen3D_PseudoLabel.py:310-395 — sklearn KDTree.query_radius around a sphere
//                centre (:347), re-centring (:365), augmentation_transform (datasets/common.py:252-334), feature assembly
//                (:383, :423-430);
//   output side  cloud_segmentation_test of utils/tester_PseudoLabel.py:176-195 — the per-sphere update of the voted class
//                probabilities, :270-283 the reprojection onto the evaluation points, and utils/metrics.py:35-118
//                fast_confusion.
//
// Sizes: a cloud is 0.4-12 M points, a batch 4 spheres; everything here is one or two passes over index / coordinate
// arrays, bound by HBM bandwidth, so the kernels are plain grid-stride loops with coalesced accesses.
#include "common.cuh"

#include <vector>

namespace kp {

// ------------------------------------------------------------------------------------------------ sphere extraction
// Membership as sklearn's KDTree.query_radius decides it: double-precision distance, point kept when dist <= r. Output
// order = ascending cloud index (sklearn returns tree order; the network is permutation-equivariant in the points).
constexpr int SPH_THREADS = 256;
constexpr int SPH_MAX_B = 16;   // spheres per call

struct SphCentres {
    int nb;
    double c[SPH_MAX_B][3];
    double r2;
};

__device__ __forceinline__ bool in_sphere(const float* __restrict__ cloud, long long i, const SphCentres& S, int b) {
    const double dx = (double)cloud[3 * i] - S.c[b][0], dy = (double)cloud[3 * i + 1] - S.c[b][1],
                 dz = (double)cloud[3 * i + 2] - S.c[b][2];
    return dx * dx + dy * dy + dz * dz <= S.r2;
}

// counts[b * n_blocks + block] = points of this block's chunk inside sphere b
__global__ void __launch_bounds__(SPH_THREADS) sph_count_kernel(const float* __restrict__ cloud, long long n,
                                                               const __grid_constant__ SphCentres S, int n_blocks,
                                                               int* __restrict__ counts) {
    __shared__ int s_cnt[SPH_MAX_B];
    if (threadIdx.x < SPH_MAX_B) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const long long chunk = (n + n_blocks - 1) / n_blocks;
    const long long lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
    for (long long i0 = lo; i0 < hi; i0 += SPH_THREADS) {
        const long long i = i0 + threadIdx.x;
        for (int b = 0; b < S.nb; b++) {
            const bool in = i < hi && in_sphere(cloud, i, S, b);
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt[b], __popc(m));
        }
    }
    __syncthreads();
    if (threadIdx.x < S.nb) counts[threadIdx.x * n_blocks + blockIdx.x] = s_cnt[threadIdx.x];
}

// offsets = exclusive scan of counts in (sphere, block) order = the stacked output position of every block's first hit
__global__ void __launch_bounds__(SPH_THREADS) sph_fill_kernel(const float* __restrict__ cloud, long long n,
                                                              const __grid_constant__ SphCentres S, int n_blocks,
                                                              const int* __restrict__ offsets, long long cap,
                                                              float* __restrict__ out_pts, long long* __restrict__ out_inds) {
    __shared__ int s_warp[SPH_THREADS / 32];
    const long long chunk = (n + n_blocks - 1) / n_blocks;
    const long long lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = 0; b < S.nb; b++) {
        long long run = offsets[b * n_blocks + blockIdx.x];
        for (long long i0 = lo; i0 < hi; i0 += SPH_THREADS) {
            const long long i = i0 + threadIdx.x;
            const bool in = i < hi && in_sphere(cloud, i, S, b);
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (lane == 0) s_warp[warp] = __popc(m);
            __syncthreads();
            int before = 0, total = 0;
            for (int w = 0; w < SPH_THREADS / 32; w++) {
                if (w < warp) before += s_warp[w];
                total += s_warp[w];
            }
            if (in) {
                const long long o = run + before + __popc(m & ((1u << lane) - 1u));
                if (o < cap) {
                    // (points[input_inds] - center_point).astype(np.float32): the subtraction runs in float64
                    out_pts[3 * o] = (float)((double)cloud[3 * i] - S.c[b][0]);
                    out_pts[3 * o + 1] = (float)((double)cloud[3 * i + 1] - S.c[b][1]);
                    out_pts[3 * o + 2] = (float)((double)cloud[3 * i + 2] - S.c[b][2]);
                    out_inds[o] = i;
                }
            }
            run += total;
            __syncthreads();
        }
    }
}

int extract_spheres_device(const float* cloud, long long n, const double* centres_host, int nb, double radius,
                           float* out_pts, long long* out_inds, long long cap, int* lengths_host, cudaStream_t stream) {
    if (n <= 0 || nb <= 0 || nb > SPH_MAX_B || !(radius > 0.0) || cap <= 0 || !cloud || !centres_host || !out_pts || !out_inds || !lengths_host)
        return fail(KP_ERR_ARG, "extract_spheres: bad arguments");
    Scratch S(stream);
    const int n_blocks = (int)std::min<long long>((n + 4 * SPH_THREADS - 1) / (4 * SPH_THREADS), 148 * 8);
    int* counts = S.alloc<int>((size_t)nb * n_blocks + 1);
    int* offsets = S.alloc<int>((size_t)nb * n_blocks + 1);
    int* total = S.alloc<int>(1);
    int* tmp = S.alloc<int>(scan_tmp_ints(nb * n_blocks));
    if (S.status != KP_OK) return S.status;
    SphCentres C;
    C.nb = nb;
    C.r2 = radius * radius;
    for (int b = 0; b < nb; b++)
        for (int d = 0; d < 3; d++) C.c[b][d] = centres_host[3 * b + d];
    ProfileScope ps("sph_extract", stream);
    sph_count_kernel<<<n_blocks, SPH_THREADS, 0, stream>>>(cloud, n, C, n_blocks, counts);
    KP_CHECK_LAUNCH();
    int rc = exclusive_scan(counts, offsets, nb * n_blocks, total, tmp, stream);
    if (rc != KP_OK) return rc;
    sph_fill_kernel<<<n_blocks, SPH_THREADS, 0, stream>>>(cloud, n, C, n_blocks, offsets, cap, out_pts, out_inds);
    KP_CHECK_LAUNCH();
    // sphere lengths = differences of the offsets at the sphere boundaries
    std::vector<int> starts(nb + 1, 0);
    for (int b = 0; b < nb; b++)
        KP_CUDA(cudaMemcpyAsync(&starts[b], offsets + (size_t)b * n_blocks, sizeof(int), cudaMemcpyDeviceToHost, stream));
    KP_CUDA(cudaMemcpyAsync(&starts[nb], total, sizeof(int), cudaMemcpyDeviceToHost, stream));
    KP_CUDA(cudaStreamSynchronize(stream));
    for (int b = 0; b < nb; b++) lengths_host[b] = starts[b + 1] - starts[b];
    if ((long long)starts[nb] > cap) return fail(KP_ERR_CAPACITY, "extract_spheres: output capacity too small");
    return KP_OK;
}

// -------------------------------------------------------------------------------------- augmentation + feature assembly
// augmented = np.sum(np.expand_dims(points, 2) * R, axis=1) * scale + noise        (datasets/common.py:318), float32:
// products rounded individually, summed left to right, no fused multiply-add. Per-sphere R [3,3] and scale [3] travel as a
// kernel argument. Features (Vaihingen3D_PseudoLabel.py:383, 423-430): [1, colours * keep, z_aug + centre_z, z_aug][:fdim].
struct AugArgs {
    int nb;
    int start[SPH_MAX_B + 1];
    float R[SPH_MAX_B][9], scale[SPH_MAX_B][3], centre_z[SPH_MAX_B], keep[SPH_MAX_B];
};

__global__ void __launch_bounds__(256) sph_augment_kernel(const float* __restrict__ pts, const float* __restrict__ noise,
                                                         const __grid_constant__ AugArgs A, float* __restrict__ out,
                                                         const float* __restrict__ colors, int ncol,
                                                         const long long* __restrict__ inds, float* __restrict__ feats,
                                                         int fdim) {
    const int n = A.start[A.nb];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int b = 0;
        while (b + 1 < A.nb && i >= A.start[b + 1]) b++;
        const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        float o[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const float s = __fadd_rn(__fadd_rn(__fmul_rn(x, A.R[b][j]), __fmul_rn(y, A.R[b][3 + j])), __fmul_rn(z, A.R[b][6 + j]));
            o[j] = __fadd_rn(__fmul_rn(s, A.scale[b][j]), noise ? noise[3 * i + j] : 0.f);
            out[3 * i + j] = o[j];
        }
        if (feats) {
            int c = 0;
            feats[(size_t)i * fdim + c++] = 1.f;
            for (int k = 0; k < ncol && c < fdim; k++)
                feats[(size_t)i * fdim + c++] = colors[(size_t)(inds ? inds[i] : i) * ncol + k] * A.keep[b];
            if (c < fdim) feats[(size_t)i * fdim + c++] = __fadd_rn(o[2], A.centre_z[b]);
            if (c < fdim) feats[(size_t)i * fdim + c++] = o[2];
            while (c < fdim) feats[(size_t)i * fdim + c++] = 0.f;
        }
    }
}

int augment_device(const float* pts, const int* lengths_host, int nb, const float* R_host, const float* scale_host,
                   const float* noise, float* out, const float* colors, int ncol, const long long* inds,
                   const float* centre_z_host, const float* keep_host, float* feats, int fdim, cudaStream_t stream) {
    if (nb <= 0 || nb > SPH_MAX_B || !pts || !lengths_host || !R_host || !scale_host || !out || (feats && fdim <= 0))
        return fail(KP_ERR_ARG, "augment: bad arguments");
    AugArgs A;
    A.nb = nb;
    A.start[0] = 0;
    for (int b = 0; b < nb; b++) {
        A.start[b + 1] = A.start[b] + lengths_host[b];
        for (int k = 0; k < 9; k++) A.R[b][k] = R_host[9 * b + k];
        for (int k = 0; k < 3; k++) A.scale[b][k] = scale_host[3 * b + k];
        A.centre_z[b] = centre_z_host ? centre_z_host[b] : 0.f;
        A.keep[b] = keep_host ? keep_host[b] : 1.f;
    }
    const int n = A.start[nb];
    if (n == 0) return KP_OK;
    ProfileScope ps("sph_augment", stream);
    sph_augment_kernel<<<ceil_div(n, 256) < 1184 ? ceil_div(n, 256) : 1184, 256, 0, stream>>>(pts, noise, A, out, colors, ncol, inds, feats, fdim);
    KP_CHECK_LAUNCH();
    return KP_OK;
}

// ----------------------------------------------------------------------------------------------------------- voting
// utils/tester_PseudoLabel.py:176-195, one sphere: points within test_radius_ratio * in_radius of the sphere centre update
//     test_probs[inds] = smooth * test_probs[inds] + (1 - smooth) * probs
// (a sphere never names a cloud point twice, so the update is race free; the spheres of a batch are applied one launch
// after the other, like the reference's loop, because overlapping spheres update the same rows in order).
// mode 1 = the order-independent form used when spheres are sharded over ranks: sum[inds] += probs, weight[inds] += 1.
__global__ void __launch_bounds__(256) vote_kernel(const float* __restrict__ probs, const float* __restrict__ pts,
                                                  const long long* __restrict__ inds, int n, int C, float r2, float smooth,
                                                  int mode, float* __restrict__ acc, float* __restrict__ weight) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)n * C; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / C), c = (int)(t - (long long)i * C);
        if (r2 > 0.f) {
            // np.sum(points ** 2, axis=1) < (ratio * in_radius) ** 2, float32, left to right
            const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
            if (!(d2 < r2)) continue;
        }
        const long long j = inds[i];
        float* a = acc + j * C + c;
        if (mode == 0) *a = __fadd_rn(__fmul_rn(smooth, *a), __fmul_rn(1.f - smooth, probs[t]));
        else {
            atomicAdd(a, probs[t]);
            if (c == 0) atomicAdd(weight + j, 1.f);
        }
    }
}

int vote_device(const float* probs, const float* pts, const long long* inds, const int* lengths_host, int nb, int C,
                float radius_limit, float smooth, int mode, float* acc, float* weight, cudaStream_t stream) {
    if (nb <= 0 || C <= 0 || !probs || !inds || !lengths_host || !acc || (mode == 1 && !weight) || (radius_limit > 0.f && !pts))
        return fail(KP_ERR_ARG, "vote: bad arguments");
    ProfileScope ps("vote", stream);
    long long i0 = 0;
    for (int b = 0; b < nb; b++) {
        const int n = lengths_host[b];
        if (n > 0) {
            const long long tot = (long long)n * C;
            vote_kernel<<<ceil_div(tot, 256) < 2368 ? ceil_div(tot, 256) : 2368, 256, 0, stream>>>(
                probs + i0 * C, pts ? pts + 3 * i0 : nullptr, inds + i0, n, C, radius_limit > 0.f ? radius_limit * radius_limit : 0.f,
                smooth, mode, acc, weight);
            KP_CHECK_LAUNCH();
        }
        i0 += n;
    }
    return KP_OK;
}

// tester_PseudoLabel.py:270-283: probs = test_probs[test_proj]; predictions = label_values[argmax(probs, axis=1)];
// metrics.py:35-118: confusion = bincount(true * C + pred) for labels 0..C-1
__global__ void __launch_bounds__(256) reproject_kernel(const float* __restrict__ acc, const float* __restrict__ weight,
                                                       const long long* __restrict__ proj, long long m, int C,
                                                       float* __restrict__ out_probs, int* __restrict__ out_pred,
                                                       const int* __restrict__ truth, unsigned long long* __restrict__ conf) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        const long long j = proj ? proj[i] : i;
        const float w = weight ? fmaxf(weight[j], 1e-12f) : 1.f;
        int best = 0;
        float bv = -1e30f;
        for (int c = 0; c < C; c++) {
            const float v = acc[j * C + c] / w;
            if (out_probs) out_probs[i * C + c] = v;
            if (v > bv) { bv = v; best = c; }   // first maximum, like np.argmax
        }
        if (out_pred) out_pred[i] = best;
        if (conf && truth) {
            const int t = truth[i];
            if (t >= 0 && t < C) atomicAdd(conf + (size_t)t * C + best, 1ull);
        }
    }
}

int reproject_device(const float* acc, const float* weight, const long long* proj, long long m, int C, float* out_probs,
                     int* out_pred, const int* truth, long long* conf, cudaStream_t stream) {
    if (m < 0 || C <= 0 || !acc) return fail(KP_ERR_ARG, "reproject: bad arguments");
    if (conf) KP_CUDA(cudaMemsetAsync(conf, 0, (size_t)C * C * sizeof(long long), stream));
    if (m == 0) return KP_OK;
    ProfileScope ps("vote_reproject", stream);
    reproject_kernel<<<ceil_div(m, 256) < 2368 ? ceil_div(m, 256) : 2368, 256, 0, stream>>>(
        acc, weight, proj, m, C, out_probs, out_pred, truth, (unsigned long long*)conf);
    KP_CHECK_LAUNCH();
    return KP_OK;
}

}  // namespace kp
