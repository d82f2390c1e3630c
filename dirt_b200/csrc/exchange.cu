// Sum of the batch-shared vertex gradient over the GPUs of one node, over peer memory (NVLink / NVSwitch): the one
// exchange of the multi-GPU path (the reference has none -- tests/multi_gpu_test.py only checks that two devices
// do not crash).  One kernel per rank and step does both halves:
//
//   push    CTA p copies this rank's buffer (count floats, 82 KB for the 5k-triangle mesh) into slot [parity][rank] of
//           peer p's exchange area with plain 16-byte stores to the peer mapping, fences, and releases flag [parity][rank]
//           of peer p with this step's sequence number;
//   reduce  the same CTA then waits until all `world` flags of its OWN area carry the sequence number and sums slice p of
//           the `world` slots in rank order into `out` -- every rank adds the same numbers in the same order, so the result
//           is bit-identical on all ranks (an all-reduce by ring or tree gives no such guarantee across algorithms).
//
// `world` CTAs in all: the kernel sits next to the rasteriser's own kernels without taking more than a few SM slots.
// Slot reuse needs no back-pressure: slots alternate with the step parity, and rank s can only push step k+2 after its
// kernel of step k+1 finished, which waited for every peer's push of k+1, which that peer issued after its own kernel
// of step k (with the reads of the slot) had completed -- the calls are stream-ordered on every rank.
#include "common.cuh"
#include "../../include/dirt_b200.h"

namespace dirt {

constexpr int EXCHANGE_THREADS = 512;
constexpr int EXCHANGE_MAX_WORLD = 16;

struct PeerTable {
    float* slots[EXCHANGE_MAX_WORLD];          // peer p's exchange area: [2][world][count_padded] floats
    unsigned int* flags[EXCHANGE_MAX_WORLD];   // peer p's flags: [2][world]
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(EXCHANGE_THREADS) peer_exchange_kernel(const float* __restrict__ local, float* __restrict__ out,
                                                                           PeerTable peers, int world, int rank,
                                                                           long long count4, unsigned int sequence)
{
    const int p = blockIdx.x;
    const int parity = (int)(sequence & 1u);
    const long long slot4 = count4;   // float4s per slot
    // push: the whole local buffer into slot [parity][rank] of peer p
    {
        const float4* src = reinterpret_cast<const float4*>(local);
        float4* dst = reinterpret_cast<float4*>(peers.slots[p]) + ((long long)parity * world + rank) * slot4;
        for (long long i = threadIdx.x; i < count4; i += EXCHANGE_THREADS) dst[i] = __ldg(src + i);
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) st_release_sys(peers.flags[p] + parity * world + rank, sequence);
    }
    // reduce: slice p of the sum over this rank's own slots, once every peer's push of this step has landed
    if (threadIdx.x < world) {
        const unsigned int* flag = peers.flags[rank] + parity * world + threadIdx.x;
        while ((int)(ld_acquire_sys(flag) - sequence) < 0) __nanosleep(64);
    }
    __syncthreads();
    const long long per = (count4 + world - 1) / world;
    const long long begin = (long long)p * per, end = min(begin + per, count4);
    const float4* mine = reinterpret_cast<const float4*>(peers.slots[rank]) + (long long)parity * world * slot4;
    for (long long i = begin + threadIdx.x; i < end; i += EXCHANGE_THREADS) {
        float4 acc = __ldcg(mine + i);   // written by a peer over NVLink: read past L1
        for (int q = 1; q < world; ++q) {
            const float4 v = __ldcg(mine + (long long)q * slot4 + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(out)[i] = acc;
    }
}

}  // namespace dirt

extern "C" {

size_t dirt_peer_exchange_bytes(int world, long long count)
{
    if (world < 1 || world > dirt::EXCHANGE_MAX_WORLD || count < 0) return 0;
    const long long count4 = (count + 3) / 4;
    return (size_t)2 * (size_t)world * (size_t)count4 * 16;
}

int dirt_peer_exchange(const float* local, float* out, void* const* peer_slots, void* const* peer_flags,
                       int world, int rank, long long count, unsigned int sequence, void* cuda_stream)
{
    if (local == nullptr || out == nullptr || peer_slots == nullptr || peer_flags == nullptr) return DIRT_ERR_NULL_POINTER;
    if (world < 1 || world > dirt::EXCHANGE_MAX_WORLD || rank < 0 || rank >= world || count <= 0 || (count & 3) != 0 ||
        sequence == 0 || local == out)
        return DIRT_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(local) & 15) != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0) return DIRT_ERR_MISALIGNED;
    dirt::PeerTable table;
    for (int p = 0; p < world; ++p) {
        if (peer_slots[p] == nullptr || peer_flags[p] == nullptr) return DIRT_ERR_NULL_POINTER;
        if ((reinterpret_cast<uintptr_t>(peer_slots[p]) & 15) != 0 || (reinterpret_cast<uintptr_t>(peer_flags[p]) & 3) != 0) return DIRT_ERR_MISALIGNED;
        table.slots[p] = static_cast<float*>(peer_slots[p]);
        table.flags[p] = static_cast<unsigned int*>(peer_flags[p]);
    }
    for (int p = world; p < dirt::EXCHANGE_MAX_WORLD; ++p) { table.slots[p] = nullptr; table.flags[p] = nullptr; }
    dirt::peer_exchange_kernel<<<world, dirt::EXCHANGE_THREADS, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        local, out, table, world, rank, count / 4, sequence);
    return cudaGetLastError() == cudaSuccess ? DIRT_OK : DIRT_ERR_CUDA;
}

}  // extern "C"
