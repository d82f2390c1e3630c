// tma_probe.cu -- stand-alone checks of the asynchronous-copy building blocks backward.cu uses, one kernel per stage so
// that a failing stage names itself:  nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// stage 1: mbarrier init / arrive / wait, no copies
__global__ void k_mbar(int* out, int with_fence)
{
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    if (with_fence) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (threadIdx.x == 0) mbar_arrive(bar);
    mbar_wait(bar, 0);
    if (threadIdx.x == 0) out[0] = 1 + ((int)(smem_u32(sm) & 1023));   // also reports the alignment of the dynamic smem base
}

// stage 2: 1-D bulk copy
__global__ void k_bulk1d(const int* src, int* out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    int* dst = reinterpret_cast<int*>(sm);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 1024);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, 256);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(src), "r"(256), "r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(bar, 0);
    out[threadIdx.x] = dst[threadIdx.x] + dst[threadIdx.x + 32];
}

// stage 3: 3-D tensor-map box, as backward.cu issues it
__global__ void k_tma3d(const __grid_constant__ CUtensorMap map, int x, int y, int z, int box_words, int* out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    int* dst = reinterpret_cast<int*>(sm);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 4096);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, box_words * 4);
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < box_words; i += 32) out[i] = dst[i];
}

#define CK(what) do { cudaError_t e = cudaDeviceSynchronize(); if (e == cudaSuccess) e = cudaGetLastError(); \
    printf("%-40s %s\n", what, e == cudaSuccess ? "ok" : cudaGetErrorString(e)); if (e != cudaSuccess) return 1; } while (0)

int main()
{
    int* d_out; cudaMalloc(&d_out, 1 << 16);
    std::vector<int> h(1 << 14);
    k_mbar<<<1, 32, 8192>>>(d_out, 0); CK("mbarrier (no init fence)");
    k_mbar<<<1, 32, 8192>>>(d_out, 1); CK("mbarrier + fence.mbarrier_init");
    cudaMemcpy(h.data(), d_out, 4, cudaMemcpyDeviceToHost);
    printf("dynamic smem base & 1023 = %d\n", h[0] - 1);

    const int W = 64, H = 48, B = 3;
    std::vector<int> img((size_t)W * H * B);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (int)i;
    int* d_img; cudaMalloc(&d_img, img.size() * 4);
    cudaMemcpy(d_img, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
    k_bulk1d<<<1, 32, 8192>>>(d_img, d_out); CK("cp.async.bulk 1-D");

    PFN_cuTensorMapEncodeTiled enc = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q) != cudaSuccess || !enc) { printf("no encoder\n"); return 1; }
    for (int elems = 1; elems <= 4; elems += (elems == 1 ? 2 : 1)) {   // 1, 3, 4 "channels": W/elems pixels per row
        CUtensorMap map;
        const int Wp = W / 4 * 4;   // row of Wp words viewed as Wp/elems pixels (only the box arithmetic matters here)
        const cuuint64_t dims[3] = {(cuuint64_t)Wp, (cuuint64_t)H, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
        const cuuint32_t box[3] = {(cuuint32_t)(12 * elems), 10u, 1u};
        const cuuint32_t es[3] = {1u, 1u, 1u};
        CUresult r = enc(&map, elems == 1 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_img, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode elems=%d -> %d\n", elems, (int)r);
        if (r != CUDA_SUCCESS) continue;
        const int x = 7, y = 5, z = 2, words = 12 * elems * 10;
        k_tma3d<<<1, 32, 8192>>>(map, x, y, z, words, d_out);
        char name[64]; snprintf(name, sizeof name, "TMA 3-D box %dx10 words", 12 * elems); CK(name);
        cudaMemcpy(h.data(), d_out, words * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < 10; ++r2)
            for (int c = 0; c < 12 * elems; ++c)
                if (h[r2 * 12 * elems + c] != (z * H + y + r2) * W + x + c) ++bad;
        printf("  contents: %d mismatches\n", bad);
    }
    return 0;
}
