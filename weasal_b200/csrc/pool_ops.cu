// Gather-style pooling operators that sit next to KPConv on every strided / decoder block:
//   max_pool      models/blocks.py:93-112   out[i,:] = max_h xpad[idx[i,h],:]   (xpad = x with one zero row: shadow -> 0)
//   closest_pool  models/blocks.py:77-90    out[i,:] = xpad[idx[i,0],:]         (nearest upsampling)
// and their adjoints. The reference builds these from cat + expand + Tensor.gather + max (materialising [N,H,C]);
// here each is one pass: a warp owns a query row, lanes stride the channels, the neighbour index is warp-uniform.
#include "common.cuh"

namespace kp {

// A warp owns a query row. The row's indices are read 32 at a time (one per lane) and the real ones compacted with a
// ballot (about half of a distance-sorted, shadow-padded row is padding); the warp then walks the real neighbours four at
// a time, each lane keeping the running maxima of its (up to CPL) channels, so that 4 x CPL feature loads are in flight
// per lane instead of one dependent index load + one feature load per neighbour.
// Semantics of blocks.py:93-112: out = max over the row of xpad[idx], xpad = x with one zero row, so a shadow entry
// contributes the value 0 (argmax -1, no gradient); among equal values the earliest real neighbour wins, and a shadow's
// 0 replaces the running maximum only when it is strictly greater (as if the padding came last, which is where the
// search puts it); a row without real neighbours gives 0.
template <typename IdxT, int CPL>
__global__ void __launch_bounds__(256) max_pool_fwd_kernel(const float* __restrict__ x, int ns, int C,
                                                          const IdxT* __restrict__ idx, int nq, int H, int stride,
                                                          float* __restrict__ out, int* __restrict__ arg,
                                                          const int* __restrict__ d_width) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= nq) return;
    // d_width (optional, device scalar): the matrix' true width; columns beyond it do not exist for this operator (a
    // fixed-width matrix must not hand a full row the extra zero candidate of a shadow column, see engine.py)
    if (d_width) H = min(H, max(*d_width, 0));
    const IdxT* row = idx + (size_t)i * stride;
    for (int cb = 0; cb < C; cb += 32 * CPL) {
        float best[CPL];
        int bj[CPL];
#pragma unroll
        for (int u = 0; u < CPL; u++) { best[u] = 0.f; bj[u] = -1; }
        bool first = true, shadow_seen = false;
        for (int hb = 0; hb < H; hb += 32) {
            const long long jl = (hb + lane < H) ? (long long)row[hb + lane] : -1;
            const bool real_l = jl >= 0 && jl < ns;
            unsigned m = __ballot_sync(0xffffffffu, real_l);
            shadow_seen = shadow_seen || (__ballot_sync(0xffffffffu, !real_l && hb + lane < H) != 0u);
            while (m) {
                int jv[4];
                float v[4][CPL];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int src = m ? (__ffs(m) - 1) : 0;
                    jv[q] = m ? (int)__shfl_sync(0xffffffffu, jl, src) : -1;
                    m &= m - 1u;
#pragma unroll
                    for (int u = 0; u < CPL; u++) {
                        const int c = cb + u * 32 + lane;
                        v[q][u] = (jv[q] >= 0 && c < C) ? __ldg(x + (size_t)jv[q] * C + c) : 0.f;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (jv[q] >= 0) {
#pragma unroll
                        for (int u = 0; u < CPL; u++)
                            if (first || v[q][u] > best[u]) { best[u] = v[q][u]; bj[u] = jv[q]; }
                        first = false;
                    }
                }
            }
        }
        if (shadow_seen) {
#pragma unroll
            for (int u = 0; u < CPL; u++)
                if (first || 0.f > best[u]) { best[u] = 0.f; bj[u] = -1; }
        }
#pragma unroll
        for (int u = 0; u < CPL; u++) {
            const int c = cb + u * 32 + lane;
            if (c < C) {
                out[(size_t)i * C + c] = best[u];
                arg[(size_t)i * C + c] = bj[u];
            }
        }
    }
}

__global__ void __launch_bounds__(256) max_pool_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ arg,
                                                          long long total, int C, float* __restrict__ dx) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int j = arg[t];
    if (j >= 0) atomicAdd(&dx[(size_t)j * C + (t % C)], dout[t]);
}

template <typename IdxT>
__global__ void __launch_bounds__(256) closest_pool_kernel(const float* __restrict__ src, int ns, int C,
                                                          const IdxT* __restrict__ idx, int nq, int stride,
                                                          float* __restrict__ dst, int backward, int src_ld) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= nq) return;
    const long long j = (long long)idx[(size_t)i * stride];
    const bool real = j >= 0 && j < ns;
    if (!backward) {
        for (int c = lane; c < C; c += 32) dst[(size_t)i * C + c] = real ? src[(size_t)j * src_ld + c] : 0.f;
    } else if (real) {
        for (int c = lane; c < C; c += 32) atomicAdd(&dst[(size_t)j * C + c], src[(size_t)i * src_ld + c]);
    }
}

int max_pool_fwd_device(const float* x, int ns, int C, const void* idx, int is_i64, int nq, int H, int stride,
                        float* out, int* arg, const int* d_width, cudaStream_t stream) {
    if (nq == 0 || C == 0) return KP_OK;
    if (C <= 64) {
        if (is_i64) max_pool_fwd_kernel<long long, 2><<<ceil_div(nq, 8), 256, 0, stream>>>(x, ns, C, (const long long*)idx, nq, H, stride, out, arg, d_width);
        else max_pool_fwd_kernel<int, 2><<<ceil_div(nq, 8), 256, 0, stream>>>(x, ns, C, (const int*)idx, nq, H, stride, out, arg, d_width);
    } else {
        if (is_i64) max_pool_fwd_kernel<long long, 8><<<ceil_div(nq, 8), 256, 0, stream>>>(x, ns, C, (const long long*)idx, nq, H, stride, out, arg, d_width);
        else max_pool_fwd_kernel<int, 8><<<ceil_div(nq, 8), 256, 0, stream>>>(x, ns, C, (const int*)idx, nq, H, stride, out, arg, d_width);
    }
    KP_CHECK_LAUNCH();
    return KP_OK;
}

int max_pool_bwd_device(const float* dout, const int* arg, int nq, int C, float* dx, int ns, cudaStream_t stream) {
    KP_CUDA(cudaMemsetAsync(dx, 0, (size_t)ns * C * sizeof(float), stream));
    const long long total = (long long)nq * C;
    if (total == 0) return KP_OK;
    max_pool_bwd_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(dout, arg, total, C, dx);
    KP_CHECK_LAUNCH();
    return KP_OK;
}

// src_ld: row stride of src in elements (>= C): a gradient that is a column slice of a wider matrix (the backward of the
// decoder's torch.cat) is read in place instead of being copied first
int closest_pool_device(const float* src, int ns, int C, const void* idx, int is_i64, int nq, int stride, float* dst,
                        int backward, int src_ld, cudaStream_t stream) {
    if (src_ld < C) return fail(KP_ERR_ARG, "closest_pool: source row stride smaller than the channel count");
    if (backward) KP_CUDA(cudaMemsetAsync(dst, 0, (size_t)ns * C * sizeof(float), stream));
    if (nq == 0 || C == 0) return KP_OK;
    if (is_i64) closest_pool_kernel<long long><<<ceil_div(nq, 8), 256, 0, stream>>>(src, ns, C, (const long long*)idx, nq, stride, dst, backward, src_ld);
    else closest_pool_kernel<int><<<ceil_div(nq, 8), 256, 0, stream>>>(src, ns, C, (const int*)idx, nq, stride, dst, backward, src_ld);
    KP_CHECK_LAUNCH();
    return KP_OK;
}

}  // namespace kp
