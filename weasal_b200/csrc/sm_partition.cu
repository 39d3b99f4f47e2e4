// Streams confined to a group of SMs (CUDA green contexts, driver API fetched through cudaGetDriverEntryPoint: no
// libcuda link). The prefetch stage of a training step (pyramid + influence lists of the NEXT batches: the counterpart
// of the reference's DataLoader workers, datasets/Vaihingen3D_PseudoLabel.py:243-252) runs on such streams: its
// kernels then never hold SM slots outside their partition, and the training step's chain of small dependent kernels
// finds free SMs whenever a node becomes ready.
#include "common.cuh"

#include <cuda.h>
#include <mutex>

namespace kp {

namespace {
template <typename Fn>
Fn driver_fn(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<Fn>(p);
}
}  // namespace

int sm_partition_streams(int min_sms, int n_streams, int priority, void** streams_out, int* sms_granted) {
    if (min_sms <= 0 || n_streams <= 0 || !streams_out) return fail(KP_ERR_ARG, "sm_partition_streams: bad arguments");
    typedef CUresult (*DeviceGetFn)(CUdevice*, int);
    typedef CUresult (*GetDevResourceFn)(CUdevice, CUdevResource*, CUdevResourceType);
    typedef CUresult (*SplitFn)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
    typedef CUresult (*GenDescFn)(CUdevResourceDesc*, CUdevResource*, unsigned int);
    typedef CUresult (*GreenCreateFn)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
    typedef CUresult (*GreenStreamFn)(CUstream*, CUgreenCtx, unsigned int, int);
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    KP_CUDA(cudaFree(nullptr));  // the primary context must be current
    int ord = 0;
    KP_CUDA(cudaGetDevice(&ord));
    DeviceGetFn device_get = driver_fn<DeviceGetFn>("cuDeviceGet");
    GetDevResourceFn get_res = driver_fn<GetDevResourceFn>("cuDeviceGetDevResource");
    SplitFn split = driver_fn<SplitFn>("cuDevSmResourceSplitByCount");
    GenDescFn gen_desc = driver_fn<GenDescFn>("cuDevResourceGenerateDesc");
    GreenCreateFn green_create = driver_fn<GreenCreateFn>("cuGreenCtxCreate");
    GreenStreamFn green_stream = driver_fn<GreenStreamFn>("cuGreenCtxStreamCreate");
    if (!device_get || !get_res || !split || !gen_desc || !green_create || !green_stream)
        return fail(KP_ERR_UNSUPPORTED, "sm_partition_streams: the driver has no green-context API");
    CUdevice dev;
    CUdevResource all, group, rest;
    unsigned int n_groups = 1;
    CUdevResourceDesc desc;
    CUgreenCtx green;
    if (device_get(&dev, ord) != CUDA_SUCCESS || get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS)
        return fail(KP_ERR_CUDA, "sm_partition_streams: cuDeviceGetDevResource failed");
    if ((unsigned)min_sms >= all.sm.smCount) return fail(KP_ERR_ARG, "sm_partition_streams: partition as large as the device");
    if (split(&group, &n_groups, &all, &rest, 0, (unsigned)min_sms) != CUDA_SUCCESS || n_groups < 1)
        return fail(KP_ERR_CUDA, "sm_partition_streams: cuDevSmResourceSplitByCount failed");
    if (gen_desc(&desc, &group, 1) != CUDA_SUCCESS || green_create(&green, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS)
        return fail(KP_ERR_CUDA, "sm_partition_streams: cuGreenCtxCreate failed");
    for (int i = 0; i < n_streams; i++) {
        CUstream s;
        if (green_stream(&s, green, CU_STREAM_NON_BLOCKING, priority) != CUDA_SUCCESS)
            return fail(KP_ERR_CUDA, "sm_partition_streams: cuGreenCtxStreamCreate failed");
        streams_out[i] = (void*)s;
    }
    if (sms_granted) *sms_granted = (int)group.sm.smCount;
    return KP_OK;  // (the green context lives as long as the process: its streams are handed out)
}

}  // namespace kp
