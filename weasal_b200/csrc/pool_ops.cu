// Gather-style pooling operators that sit next to KPConv on every strided / decoder block:
//   max_pool      models/blocks.py:93-112   out[i,:] = max_h xpad[idx[i,h],:]   (xpad = x with one zero row: shadow -> 0)
//   closest_pool  models/blocks.py:77-90    out[i,:] = xpad[idx[i,0],:]         (nearest upsampling)
// and their adjoints. The reference builds these from cat + expand + Tensor.gather + max (materialising [N,H,C]);
// here each is one pass: a warp owns a query row, lanes stride the channels, the neighbour index is warp-uniform.
#include "common.cuh"

namespace kp {

// lanes stride the channels; the neighbour loop is outermost so each index is read once per row, and a lane keeps the
// running maxima of its (up to CPL) channels in registers. C > 32*CPL falls back to repeating the sweep per block of
// 32*CPL channels.
template <typename IdxT, int CPL>
__global__ void __launch_bounds__(256) max_pool_fwd_kernel(const float* __restrict__ x, int ns, int C,
                                                          const IdxT* __restrict__ idx, int nq, int H, int stride,
                                                          float* __restrict__ out, int* __restrict__ arg) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= nq) return;
    const IdxT* row = idx + (size_t)i * stride;
    for (int cb = 0; cb < C; cb += 32 * CPL) {
        float best[CPL];
        int bj[CPL];
#pragma unroll
        for (int u = 0; u < CPL; u++) { best[u] = 0.f; bj[u] = -1; }
        for (int h = 0; h < H; h++) {
            const long long j = (long long)row[h];
            const bool real = j >= 0 && j < ns;
#pragma unroll
            for (int u = 0; u < CPL; u++) {
                const int c = cb + u * 32 + lane;
                const float v = (real && c < C) ? __ldg(x + (size_t)j * C + c) : 0.f;
                if (h == 0 || v > best[u]) { best[u] = v; bj[u] = real ? (int)j : -1; }
            }
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
                                                          float* __restrict__ dst, int backward) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= nq) return;
    const long long j = (long long)idx[(size_t)i * stride];
    const bool real = j >= 0 && j < ns;
    if (!backward) {
        for (int c = lane; c < C; c += 32) dst[(size_t)i * C + c] = real ? src[(size_t)j * C + c] : 0.f;
    } else if (real) {
        for (int c = lane; c < C; c += 32) atomicAdd(&dst[(size_t)j * C + c], src[(size_t)i * C + c]);
    }
}

int max_pool_fwd_device(const float* x, int ns, int C, const void* idx, int is_i64, int nq, int H, int stride,
                        float* out, int* arg, cudaStream_t stream) {
    if (nq == 0 || C == 0) return KP_OK;
    if (C <= 64) {
        if (is_i64) max_pool_fwd_kernel<long long, 2><<<ceil_div(nq, 8), 256, 0, stream>>>(x, ns, C, (const long long*)idx, nq, H, stride, out, arg);
        else max_pool_fwd_kernel<int, 2><<<ceil_div(nq, 8), 256, 0, stream>>>(x, ns, C, (const int*)idx, nq, H, stride, out, arg);
    } else {
        if (is_i64) max_pool_fwd_kernel<long long, 8><<<ceil_div(nq, 8), 256, 0, stream>>>(x, ns, C, (const long long*)idx, nq, H, stride, out, arg);
        else max_pool_fwd_kernel<int, 8><<<ceil_div(nq, 8), 256, 0, stream>>>(x, ns, C, (const int*)idx, nq, H, stride, out, arg);
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

int closest_pool_device(const float* src, int ns, int C, const void* idx, int is_i64, int nq, int stride, float* dst,
                        int backward, cudaStream_t stream) {
    if (backward) KP_CUDA(cudaMemsetAsync(dst, 0, (size_t)ns * C * sizeof(float), stream));
    if (nq == 0 || C == 0) return KP_OK;
    if (is_i64) closest_pool_kernel<long long><<<ceil_div(nq, 8), 256, 0, stream>>>(src, ns, C, (const long long*)idx, nq, stride, dst, backward);
    else closest_pool_kernel<int><<<ceil_div(nq, 8), 256, 0, stream>>>(src, ns, C, (const int*)idx, nq, stride, dst, backward);
    KP_CHECK_LAUNCH();
    return KP_OK;
}

}  // namespace kp
