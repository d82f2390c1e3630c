// Stand-in for <tensorflow/core/framework/tensor.h>: just enough of TensorFlow / Eigen for the reference's
// csrc/rasterise_grad_egl.cu to compile UNMODIFIED with nvcc (no TensorFlow in this image).
//
// TEST INFRASTRUCTURE (oracle/_ref): it exists so that the reference's own gradient kernel `assemble_grads`
// (csrc/rasterise_grad_egl.cu:93-236) can be run on a B200 and pin the oracle and the CUDA path.  Nothing under
// dirt_b200/ uses it.  What is provided, and the TensorFlow facility each piece stands in for:
//   tensorflow::TTypes<T,N>::{Tensor,ConstTensor}  Eigen::TensorMap<Eigen::Tensor<T,N,RowMajor>>: operator()(i...),
//                                                  dimension(i), data(); unchecked row-major indexing, exactly what
//                                                  Eigen does under -DNDEBUG (csrc/CMakeLists.txt:41)
//   tensorflow::Tensor                             dim_size(i), NumElements(), tensor<T,N>() over a device pointer
//   Eigen::GpuDevice                               stream()
//   LOG(FATAL) << ...                              prints and aborts
#ifndef DIRT_REF_SHIM_TENSOR_H
#define DIRT_REF_SHIM_TENSOR_H

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <sstream>

namespace Eigen {
struct GpuDevice {
    cudaStream_t stream_ = nullptr;
    cudaStream_t stream() const { return stream_; }
};
}  // namespace Eigen

namespace tensorflow {

template <typename T, int N>
struct TensorMapShim {
    T* data_;
    long dims_[N];

    __host__ __device__ long dimension(int i) const { return dims_[i]; }
    __host__ __device__ T* data() const { return data_; }
    __host__ __device__ T& operator()(long i0, long i1) const
    {
        static_assert(N == 2, "rank mismatch");
        return data_[i0 * dims_[1] + i1];
    }
    __host__ __device__ T& operator()(long i0, long i1, long i2) const
    {
        static_assert(N == 3, "rank mismatch");
        return data_[(i0 * dims_[1] + i1) * dims_[2] + i2];
    }
    __host__ __device__ T& operator()(long i0, long i1, long i2, long i3) const
    {
        static_assert(N == 4, "rank mismatch");
        return data_[((i0 * dims_[1] + i1) * dims_[2] + i2) * dims_[3] + i3];
    }
};

template <typename T, int N>
struct TTypes {
    typedef TensorMapShim<T, N> Tensor;
    typedef TensorMapShim<const T, N> ConstTensor;
};

class Tensor {
public:
    Tensor() : data_(nullptr), rank_(0) { for (long& d : dims_) d = 0; }
    Tensor(void* device_data, std::initializer_list<long> dims) : data_(device_data), rank_((int)dims.size())
    {
        int i = 0;
        for (long d : dims) dims_[i++] = d;
        for (; i < 8; ++i) dims_[i] = 1;
    }
    long dim_size(int i) const { return dims_[i]; }
    long NumElements() const
    {
        long n = 1;
        for (int i = 0; i < rank_; ++i) n *= dims_[i];
        return n;
    }
    template <typename T, int N>
    typename TTypes<T, N>::Tensor tensor()
    {
        typename TTypes<T, N>::Tensor m;
        m.data_ = static_cast<T*>(data_);
        for (int i = 0; i < N; ++i) m.dims_[i] = dims_[i];
        return m;
    }
    template <typename T, int N>
    typename TTypes<T, N>::ConstTensor tensor() const
    {
        typename TTypes<T, N>::ConstTensor m;
        m.data_ = static_cast<const T*>(data_);
        for (int i = 0; i < N; ++i) m.dims_[i] = dims_[i];
        return m;
    }

private:
    void* data_;
    int rank_;
    long dims_[8];
};

namespace shim {
struct FatalStream {
    std::ostringstream s;
    template <typename U> FatalStream& operator<<(const U& v) { s << v; return *this; }
    [[noreturn]] ~FatalStream()
    {
        std::fprintf(stderr, "FATAL (reference code): %s\n", s.str().c_str());
        std::abort();
    }
};
}  // namespace shim

}  // namespace tensorflow

#define FATAL 3
#define LOG(severity) ::tensorflow::shim::FatalStream()

#endif
