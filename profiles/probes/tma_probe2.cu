// tma_probe2.cu -- variants of one tensor-map box load, ONE per process (an illegal instruction kills the context):
//   ./tma_probe2 <rank 2|3> <x> <box_w> <type 0=int32 1=float32> <mode 0=own asm shared::cluster, 1=own asm shared::cta, 2=libcu++ wrapper> <l2promo 0|1>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
namespace cde = cuda::device::experimental;
using barrier_t = cuda::barrier<cuda::thread_scope_block>;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK, int MODE>
__global__ void k_load(const __grid_constant__ CUtensorMap map, int x, int y, int z, int box_words, int* out)
{
    __shared__ alignas(128) int dst[2048];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier_t bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier_t::arrival_token token;
    if (threadIdx.x == 0) {
        if (MODE == 2) {
            if (RANK == 2) cde::cp_async_bulk_tensor_2d_global_to_shared(dst, &map, x, y, bar);
            else cde::cp_async_bulk_tensor_3d_global_to_shared(dst, &map, x, y, z, bar);
        } else {
            const uint32_t mb = smem_u32(cuda::device::barrier_native_handle(bar));
            if (RANK == 2) {
                if (MODE == 0)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(x), "r"(y), "r"(mb) : "memory");
                else
                    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(x), "r"(y), "r"(mb) : "memory");
            } else {
                if (MODE == 0)
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(x), "r"(y), "r"(z), "r"(mb) : "memory");
                else
                    asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(x), "r"(y), "r"(z), "r"(mb) : "memory");
            }
        }
        token = cuda::device::barrier_arrive_tx(bar, 1, box_words * 4);
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < box_words; i += blockDim.x) out[i] = dst[i];
}

int main(int argc, char** argv)
{
    if (argc < 7) { printf("usage\n"); return 2; }
    const int rank = atoi(argv[1]), x = atoi(argv[2]), box_w = atoi(argv[3]), type = atoi(argv[4]), mode = atoi(argv[5]), promo = atoi(argv[6]);
    const int W = 64, H = 48, B = 3, y = 5, z = 2;
    std::vector<int> img((size_t)W * H * B);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (int)i;
    int *d_img, *d_out; cudaMalloc(&d_img, img.size() * 4); cudaMalloc(&d_out, 1 << 16);
    cudaMemcpy(d_img, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
    PFN_cuTensorMapEncodeTiled enc = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q) != cudaSuccess || !enc) { printf("no encoder\n"); return 1; }
    alignas(64) CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, 10u, 1u};
    const cuuint32_t es[3] = {1u, 1u, 1u};
    CUresult r = enc(&map, type ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_INT32, rank, rank == 2 ? (void*)(d_img + (size_t)z * W * H) : (void*)d_img,
                     dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("rank=%d x=%d box=%d type=%d mode=%d promo=%d: encode failed %d\n", rank, x, box_w, type, mode, promo, (int)r); return 1; }
    const int words = box_w * 10;
    if (rank == 2) { if (mode == 0) k_load<2, 0><<<1, 32>>>(map, x, y, z, words, d_out); else if (mode == 1) k_load<2, 1><<<1, 32>>>(map, x, y, z, words, d_out); else k_load<2, 2><<<1, 32>>>(map, x, y, z, words, d_out); }
    else { if (mode == 0) k_load<3, 0><<<1, 32>>>(map, x, y, z, words, d_out); else if (mode == 1) k_load<3, 1><<<1, 32>>>(map, x, y, z, words, d_out); else k_load<3, 2><<<1, 32>>>(map, x, y, z, words, d_out); }
    cudaError_t e = cudaDeviceSynchronize();
    int bad = -1;
    if (e == cudaSuccess) {
        std::vector<int> h(words);
        cudaMemcpy(h.data(), d_out, words * 4, cudaMemcpyDeviceToHost);
        bad = 0;
        for (int r2 = 0; r2 < 10; ++r2)
            for (int c = 0; c < box_w; ++c)
                if (h[r2 * box_w + c] != (z * H + y + r2) * W + x + c) ++bad;
    }
    printf("rank=%d x=%d box=%d type=%d mode=%d promo=%d: %s, mismatches %d\n", rank, x, box_w, type, mode, promo, e == cudaSuccess ? "ok" : cudaGetErrorString(e), bad);
    return e == cudaSuccess ? 0 : 1;
}
