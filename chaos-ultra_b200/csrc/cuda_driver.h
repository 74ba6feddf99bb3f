/*
 * cuda_driver.h -- lazily bound CUDA driver API.
 *
 * libchaos_ultra.so must load on a machine without a driver (the build host, CI) and fail
 * loudly only when a provider is created, so libcuda is dlopen'ed at that point instead of
 * being a link-time dependency.  Role in the reference: JCuda's native loader
 * (cudarenderer/CudaHelpers.java:27-45 and the UnsatisfiedLinkError path in
 * CudaFractalRendererProvider.java:80-88).
 */
#ifndef CHAOS_CUDA_DRIVER_H
#define CHAOS_CUDA_DRIVER_H

#include <cuda.h>

#define CHAOS_CU_FUNCS(X)                       \
    X(cuInit)                                   \
    X(cuDriverGetVersion)                       \
    X(cuDeviceGet)                              \
    X(cuDeviceGetCount)                         \
    X(cuDeviceGetAttribute)                     \
    X(cuDeviceGetName)                          \
    X(cuDevicePrimaryCtxRetain)                 \
    X(cuDevicePrimaryCtxRelease)                \
    X(cuCtxPushCurrent)                         \
    X(cuCtxPopCurrent)                          \
    X(cuCtxSynchronize)                         \
    X(cuModuleLoadData)                         \
    X(cuModuleUnload)                           \
    X(cuModuleGetFunction)                      \
    X(cuModuleGetGlobal)                        \
    X(cuMemAlloc)                               \
    X(cuMemAllocPitch)                          \
    X(cuMemFree)                                \
    X(cuMemHostAlloc)                           \
    X(cuMemFreeHost)                            \
    X(cuMemHostGetDevicePointer)                \
    X(cuMemHostRegister)                        \
    X(cuMemHostUnregister)                      \
    X(cuIpcGetMemHandle)                        \
    X(cuIpcOpenMemHandle)                       \
    X(cuIpcCloseMemHandle)                      \
    X(cuMemcpyHtoD)                             \
    X(cuMemcpyDtoH)                             \
    X(cuMemcpyHtoDAsync)                        \
    X(cuMemcpyDtoHAsync)                        \
    X(cuMemcpy2D)                               \
    X(cuMemsetD8Async)                          \
    X(cuMemsetD32Async)                         \
    X(cuStreamCreate)                           \
    X(cuStreamDestroy)                          \
    X(cuStreamSynchronize)                      \
    X(cuStreamWaitEvent)                        \
    X(cuEventCreate)                            \
    X(cuEventRecord)                            \
    X(cuEventSynchronize)                       \
    X(cuEventElapsedTime)                       \
    X(cuEventDestroy)                           \
    X(cuLaunchKernel)                           \
    X(cuFuncSetAttribute)                       \
    X(cuFuncGetAttribute)                       \
    X(cuOccupancyMaxActiveBlocksPerMultiprocessor) \
    X(cuGetErrorString)                         \
    X(cuGetErrorName)

struct chaos_cuda_driver {
#define X(name) decltype(&name) p_##name;
    CHAOS_CU_FUNCS(X)
#undef X
    void *handle;
};

/* returns NULL and fills err (if non-NULL) when libcuda cannot be loaded */
const chaos_cuda_driver *chaos_cuda_driver_get(const char **err);

#endif
