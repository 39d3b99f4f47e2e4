// KPConv building blocks on CUDA cores (fp32): the kernel-point-weighted neighbour gather ("WF"), its
// adjoint scatter, and the transposed neighbour table used by the atomics-free dX path.
// Math follows models/blocks.py:277-374 (rigid, 'linear' influence, 'sum' aggregation):
//   w[i,k,h]  = max(0, 1 - ||(s[idx[i,h]] - q[i]) - kp[k]|| / KP_extent)          (:281-298, :337)
//   WF[i,k,:] = sum_h w[i,k,h] * x[idx[i,h],:]                                    (:357-363)
// A shadow neighbour (idx == Ns) is the point (1e6,1e6,1e6) with zero features (:278, :357): it contributes
// exactly zero and is skipped.
// These kernels are the fp32 cross-check / bring-up path; the product forward/backward is kpconv_tc.cu.
#include "common.cuh"

namespace kp {

__device__ __forceinline__ float influence(float rx, float ry, float rz, float kx, float ky, float kz, float inv_ext) {
    const float dx = rx - kx, dy = ry - ky, dz = rz - kz;
    const float d2 = dx * dx + dy * dy + dz * dz;
    return fmaxf(0.f, 1.f - sqrtf(d2) * inv_ext);
}

template <typename IdxT>
__global__ void __launch_bounds__(128) kpconv_wf_kernel(const float* __restrict__ q, int nq,
                                                       const float* __restrict__ s, int ns,
                                                       const IdxT* __restrict__ idx, int H, int idx_stride,
                                                       const float* __restrict__ x, int cin,
                                                       const float* __restrict__ kp, int K, float inv_ext,
                                                       float* __restrict__ wf) {
    extern __shared__ float s_acc[];  // [warps][K*cin]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float* acc = s_acc + (size_t)warp * K * cin;
    float kx = 0.f, ky = 0.f, kz = 0.f;
    if (lane < K) { kx = kp[3 * lane]; ky = kp[3 * lane + 1]; kz = kp[3 * lane + 2]; }
    for (int i = blockIdx.x * nwarps + warp; i < nq; i += gridDim.x * nwarps) {
        for (int t = lane; t < K * cin; t += 32) acc[t] = 0.f;
        __syncwarp();
        const float qx = q[3 * (size_t)i], qy = q[3 * (size_t)i + 1], qz = q[3 * (size_t)i + 2];
        for (int h = 0; h < H; h++) {
            const long long j = (long long)idx[(size_t)i * idx_stride + h];
            if (j < 0 || j >= ns) continue;
            const float rx = s[3 * j] - qx, ry = s[3 * j + 1] - qy, rz = s[3 * j + 2] - qz;
            const float w = (lane < K) ? influence(rx, ry, rz, kx, ky, kz, inv_ext) : 0.f;
            unsigned m = __ballot_sync(0xffffffffu, w > 0.f);
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                const float wk = __shfl_sync(0xffffffffu, w, k);
                for (int c = lane; c < cin; c += 32) acc[k * cin + c] += wk * x[(size_t)j * cin + c];
            }
        }
        __syncwarp();
        for (int t = lane; t < K * cin; t += 32) wf[(size_t)i * K * cin + t] = acc[t];
        __syncwarp();
    }
}

// dX[j,:] += sum_k w[i,k,h] * dWF[i,k,:]   (adjoint of the gather; atomics)
template <typename IdxT>
__global__ void __launch_bounds__(128) kpconv_dx_atomic_kernel(const float* __restrict__ q, int nq,
                                                              const float* __restrict__ s, int ns,
                                                              const IdxT* __restrict__ idx, int H, int idx_stride,
                                                              const float* __restrict__ dwf, int cin,
                                                              const float* __restrict__ kp, int K, float inv_ext,
                                                              float* __restrict__ dx) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float kx = 0.f, ky = 0.f, kz = 0.f;
    if (lane < K) { kx = kp[3 * lane]; ky = kp[3 * lane + 1]; kz = kp[3 * lane + 2]; }
    for (int i = blockIdx.x * nwarps + warp; i < nq; i += gridDim.x * nwarps) {
        const float qx = q[3 * (size_t)i], qy = q[3 * (size_t)i + 1], qz = q[3 * (size_t)i + 2];
        for (int h = 0; h < H; h++) {
            const long long j = (long long)idx[(size_t)i * idx_stride + h];
            if (j < 0 || j >= ns) continue;
            const float rx = s[3 * j] - qx, ry = s[3 * j + 1] - qy, rz = s[3 * j + 2] - qz;
            const float w = (lane < K) ? influence(rx, ry, rz, kx, ky, kz, inv_ext) : 0.f;
            const unsigned m0 = __ballot_sync(0xffffffffu, w > 0.f);
            if (!m0) continue;
            for (int c0 = 0; c0 < cin; c0 += 32) {  // warp-uniform trip count: the shuffles below need all lanes
                const int c = c0 + lane;
                float v = 0.f;
                unsigned m = m0;
                while (m) {
                    const int k = __ffs(m) - 1;
                    m &= m - 1;
                    const float wk = __shfl_sync(0xffffffffu, w, k);
                    if (c < cin) v += wk * dwf[((size_t)i * K + k) * cin + c];
                }
                if (c < cin) atomicAdd(&dx[(size_t)j * cin + c], v);
            }
        }
    }
}

int kpconv_wf_device(const float* q, int nq, const float* s, int ns, const void* idx, int idx_is_i64, int H,
                     int idx_stride, const float* x, int cin, const float* kp, int K, float extent, float* wf,
                     cudaStream_t stream) {
    if (nq == 0) return KP_OK;
    if (K > 32) return fail(KP_ERR_UNSUPPORTED, "kpconv: more than 32 kernel points");
    int warps = 4;
    size_t smem = (size_t)warps * K * cin * sizeof(float);
    while (smem > 96 * 1024 && warps > 1) { warps >>= 1; smem = (size_t)warps * K * cin * sizeof(float); }
    if (smem > 200 * 1024) return fail(KP_ERR_UNSUPPORTED, "kpconv_wf: in_channels too large");
    const int grid = ceil_div(nq, warps);
    const float inv_ext = 1.f / extent;
    if (idx_is_i64) {
        KP_CUDA(cudaFuncSetAttribute(kpconv_wf_kernel<long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kpconv_wf_kernel<long long><<<grid, warps * 32, smem, stream>>>(q, nq, s, ns, (const long long*)idx, H, idx_stride,
                                                                    x, cin, kp, K, inv_ext, wf);
    } else {
        KP_CUDA(cudaFuncSetAttribute(kpconv_wf_kernel<int>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kpconv_wf_kernel<int><<<grid, warps * 32, smem, stream>>>(q, nq, s, ns, (const int*)idx, H, idx_stride, x, cin,
                                                              kp, K, inv_ext, wf);
    }
    KP_CHECK_LAUNCH();
    return KP_OK;
}

// dx must be zero-initialised by the caller.
int kpconv_dx_atomic_device(const float* q, int nq, const float* s, int ns, const void* idx, int idx_is_i64, int H,
                            int idx_stride, const float* dwf, int cin, const float* kp, int K, float extent, float* dx,
                            cudaStream_t stream) {
    if (nq == 0) return KP_OK;
    if (K > 32) return fail(KP_ERR_UNSUPPORTED, "kpconv: more than 32 kernel points");
    const int grid = ceil_div(nq, 4);
    const float inv_ext = 1.f / extent;
    if (idx_is_i64)
        kpconv_dx_atomic_kernel<long long><<<grid, 128, 0, stream>>>(q, nq, s, ns, (const long long*)idx, H, idx_stride, dwf,
                                                                  cin, kp, K, inv_ext, dx);
    else
        kpconv_dx_atomic_kernel<int><<<grid, 128, 0, stream>>>(q, nq, s, ns, (const int*)idx, H, idx_stride, dwf, cin, kp,
                                                            K, inv_ext, dx);
    KP_CHECK_LAUNCH();
    return KP_OK;
}

}  // namespace kp
