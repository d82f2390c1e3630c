// Stand-in for <tensorflow/core/util/cuda_launch_config.h> (see ../framework/tensor.h): GetCuda2DLaunchConfig.
// The reference's kernels are grid-stride loops (CUDA_AXIS_KERNEL_LOOP, csrc/tf_cuda_utils.h:10-12), so the launch
// geometry chosen here cannot change a result; it follows TensorFlow's shape (256-thread blocks, x first).
#ifndef DIRT_REF_SHIM_LAUNCH_CONFIG_H
#define DIRT_REF_SHIM_LAUNCH_CONFIG_H

#include "tensorflow/core/framework/tensor.h"

namespace tensorflow {

struct Cuda2DLaunchConfig {
    dim3 virtual_thread_count = dim3(0, 0, 0);
    dim3 thread_per_block = dim3(0, 0, 0);
    dim3 block_count = dim3(0, 0, 0);
};

inline Cuda2DLaunchConfig GetCuda2DLaunchConfig(int xdim, int ydim, const Eigen::GpuDevice&)
{
    Cuda2DLaunchConfig config;
    if (xdim <= 0 || ydim <= 0) return config;
    const int threads = 256;
    const int block_cols = std::min(xdim, threads);
    const int block_rows = std::max(threads / block_cols, 1);
    int device = 0, sms = 1, threads_per_sm = 2048;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&threads_per_sm, cudaDevAttrMaxThreadsPerMultiProcessor, device);
    const int max_blocks = std::max(sms * threads_per_sm / threads, 1);
    const int grid_x = std::min((xdim + block_cols - 1) / block_cols, max_blocks);
    const int grid_y = std::min(std::max(max_blocks / grid_x, 1), std::max(ydim / block_rows, 1));
    config.virtual_thread_count = dim3(xdim, ydim, 1);
    config.thread_per_block = dim3(block_cols, block_rows, 1);
    config.block_count = dim3(grid_x, grid_y, 1);
    return config;
}

}  // namespace tensorflow

#endif
