// raster.cu -- forward rasteriser: one warp per 8x8 screen tile, z-buffer in registers.
//
// Replaces the GL draw loop + upload_background/download_pixels of the reference
// (csrc/rasterise_egl.cpp:349-396, csrc/rasterise_egl.cu:10-38,65-91): the background is read
// and the output written directly in the [B,H,W,C] tensors, top row first.
//
// Per tile: the tile's binned face list and the image's large-face list are consumed in chunks
// of 32 (one face per lane: load the 64-B coverage record, move the edge functions to the tile
// origin in int64, reject faces that cannot touch the tile), survivors are parked in shared
// memory and broadcast one at a time to all lanes, each lane testing its two pixels (S5, S7).
// The winner is min (depth key, face index) -- order independent, so list order does not matter.
#include "common.cuh"

namespace dirt {

constexpr int WARPS_PER_BLOCK = 4;

struct __align__(16) Slot {
    int32_t A0, B0, A1, B1;
    int32_t A2, B2, Q0, Q1;
    int32_t Q2;
    float zA, zB, zC;
    int32_t face, kind, pad0, pad1;
};
static_assert(sizeof(Slot) == 64, "Slot must be 64 bytes");

struct PixelPair {
    uint32_t key0, key1;
    int32_t face0, face1;
};

__device__ __forceinline__ int32_t clamp_q(int64_t q)
{
    const int64_t lim = (int64_t)1 << 30;
    return (int32_t)(q > lim ? lim : (q < -lim ? -lim : q));
}

// hard faces (H1-H3): homogeneous double-precision evaluation at this lane's two pixels
__device__ __noinline__ void hard_face_pixels(const float* __restrict__ verts, const TriInterp* __restrict__ itp_b,
                                              int face, int H, int W, int col, int row0, bool& in0, bool& in1,
                                              uint32_t& key0, uint32_t& key1)
{
    const TriInterp t = load_interp(itp_b + face);
    float p[3][4];
    const int vid[3] = {t.v0, t.v1, t.v2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(verts) + vid[k]);
        p[k][0] = v.x; p[k][1] = v.y; p[k][2] = v.z; p[k][3] = v.w;
    }
    double gq[3][3], gs[3], gz[3];
    in0 = in1 = false;
    key0 = key1 = KEY_EMPTY;
    if (!exact::planes_double(p, H, W, gq, gs, gz)) return;
#pragma unroll
    for (int pix = 0; pix < 2; ++pix) {
        const int row = row0 + pix;
        bool in = true;
        double qv[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double v = exact::plane_double(gq[k], col, row);
            qv[k] = v;
            const bool own = gq[k][0] > 0.0 || (gq[k][0] == 0.0 && gq[k][1] > 0.0);
            if (!(v > 0.0 || (v == 0.0 && own))) in = false;
        }
        const double sum = __dadd_rn(__dadd_rn(qv[0], qv[1]), qv[2]);
        in = in && (sum > 0.0);
        const float z = (float)exact::plane_double(gz, col, row);
        const uint32_t key = exact::depth_key(z);
        if (pix == 0) { in0 = in; key0 = key; } else { in1 = in; key1 = key; }
    }
}

// Consume one face list for this warp's tile.
__device__ __forceinline__ void consume_list(const int* __restrict__ list, int count, const TriCov* __restrict__ cov_b,
                                             const TriInterp* __restrict__ itp_b, const float* __restrict__ verts,
                                             Slot* slots, int lane, int tcol0, int trow0, int H, int W,
                                             PixelPair& best)
{
    const int dx = lane & 7, dy = (lane >> 3) * 2;
    const int col = tcol0 + dx, row0 = trow0 + dy;
    const float colf = (float)col, row0f = (float)row0, row1f = (float)(row0 + 1);
    for (int base = 0; base < count; base += 32) {
        const int i = base + lane;
        const int f = (i < count) ? __ldg(&list[i]) : -1;
        bool alive = false;
        if (f >= 0) {
            const TriCov c = load_cov(cov_b + f);
            Slot s;
            s.face = f; s.kind = (int32_t)c.kind; s.pad0 = s.pad1 = 0;
            s.A0 = c.A0; s.B0 = c.B0; s.A1 = c.A1; s.B1 = c.B1; s.A2 = c.A2; s.B2 = c.B2;
            s.zA = c.zA; s.zB = c.zB; s.zC = c.zC;
            if (c.kind == 1u) {
                const int64_t Q0 = c.q0 + (int64_t)c.A0 * tcol0 + (int64_t)c.B0 * trow0;
                const int64_t Q1 = c.q1 + (int64_t)c.A1 * tcol0 + (int64_t)c.B1 * trow0;
                const int64_t Q2 = c.q2 + (int64_t)c.A2 * tcol0 + (int64_t)c.B2 * trow0;
                // the largest value each edge function takes on the tile's 8x8 pixel centres
                const int64_t m0 = Q0 + max(0, 7 * c.A0) + max(0, 7 * c.B0);
                const int64_t m1 = Q1 + max(0, 7 * c.A1) + max(0, 7 * c.B1);
                const int64_t m2 = Q2 + max(0, 7 * c.A2) + max(0, 7 * c.B2);
                alive = (m0 >= 0) && (m1 >= 0) && (m2 >= 0);
                s.Q0 = clamp_q(Q0); s.Q1 = clamp_q(Q1); s.Q2 = clamp_q(Q2);
            } else if (c.kind == 2u) {
                alive = true;
                s.Q0 = s.Q1 = s.Q2 = 0;
            }
            if (alive) {
                uint4* dst = reinterpret_cast<uint4*>(&slots[lane]);
                const uint4* src = reinterpret_cast<const uint4*>(&s);
                dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
            }
        }
        unsigned mask = __ballot_sync(0xffffffffu, alive);
        __syncwarp();
        while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const Slot s = slots[j];
            bool in0, in1;
            uint32_t key0, key1;
            if (s.kind == 1) {
                const int32_t n0 = s.Q0 + s.A0 * dx + s.B0 * dy;
                const int32_t n1 = s.Q1 + s.A1 * dx + s.B1 * dy;
                const int32_t n2 = s.Q2 + s.A2 * dx + s.B2 * dy;
                in0 = (n0 | n1 | n2) >= 0;
                in1 = ((n0 + s.B0) | (n1 + s.B1) | (n2 + s.B2)) >= 0;
                key0 = exact::depth_key(exact::depth_normal(s.zA, s.zB, s.zC, colf, row0f));
                key1 = exact::depth_key(exact::depth_normal(s.zA, s.zB, s.zC, colf, row1f));
            } else {
                hard_face_pixels(verts, itp_b, s.face, H, W, col, row0, in0, in1, key0, key1);
            }
            if (in0 && (key0 < best.key0 || (key0 == best.key0 && s.face < best.face0))) {
                best.key0 = key0; best.face0 = s.face;
            }
            if (in1 && (key1 < best.key1 || (key1 == best.key1 && s.face < best.face1))) {
                best.key1 = key1; best.face1 = s.face;
            }
        }
        __syncwarp();
    }
}

// MODE 0: colour forward (pixels [+ face ids]); MODE 1: visibility (face ids and/or G-buffer)
template <int MODE, int CT>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) raster_kernel(
    const float* __restrict__ vertices, const float* __restrict__ background,
    const float* __restrict__ vertex_colors, float* __restrict__ pixels, int32_t* __restrict__ face_ids_out,
    float* __restrict__ gbuffer_out, Workspace ws, Dims d)
{
    __shared__ Slot slots_all[WARPS_PER_BLOCK][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tile_global = (long long)blockIdx.x * WARPS_PER_BLOCK + warp;
    if (tile_global >= (long long)d.B * d.tiles) return;
    const int b = (int)(tile_global / d.tiles);
    const int t = (int)(tile_global - (long long)b * d.tiles);
    const int ty = t / d.tiles_x, tx = t - ty * d.tiles_x;
    const int tcol0 = tx * TILE, trow0 = ty * TILE;

    const TriCov* cov_b = ws.cov + (size_t)b * d.F;
    const TriInterp* itp_b = ws.itp + (size_t)b * d.F;
    const float* verts = vertices + (size_t)b * d.V * 4;

    PixelPair best;
    best.key0 = best.key1 = KEY_EMPTY;
    best.face0 = best.face1 = -1;

    const int2 range = ws.tile_range[tile_global];
    consume_list(ws.refs + range.x, range.y, cov_b, itp_b, verts, slots_all[warp], lane, tcol0, trow0, d.H, d.W, best);
    const int nlarge = ws.large_count[b];
    if (nlarge > 0)
        consume_list(ws.large_list + (size_t)b * d.F, nlarge, cov_b, itp_b, verts, slots_all[warp], lane, tcol0, trow0,
                     d.H, d.W, best);

    const int col = tcol0 + (lane & 7), row0 = trow0 + (lane >> 3) * 2;
    if (col >= d.W) return;
#pragma unroll
    for (int pix = 0; pix < 2; ++pix) {
        const int row = row0 + pix;
        if (row >= d.H) break;
        const int face = pix ? best.face1 : best.face0;
        const size_t p = ((size_t)b * d.H + row) * d.W + col;
        if (MODE == 1) {
            if (face_ids_out) face_ids_out[p] = face;
            if (gbuffer_out) {
                float4 g = make_float4(-1.f, -1.f, -1.f, __int_as_float(0x7f800000));
                if (face >= 0) g = exact::gbuffer_at(load_interp(itp_b + face), col, row);
                reinterpret_cast<float4*>(gbuffer_out)[p] = g;
            }
        } else {
            if (face_ids_out) face_ids_out[p] = face;
            const int C = (CT > 0) ? CT : d.C;
            if (face < 0) {
                if (CT == 4) {
                    reinterpret_cast<float4*>(pixels)[p] = __ldg(reinterpret_cast<const float4*>(background) + p);
                } else {
                    for (int ch = 0; ch < C; ++ch) pixels[p * C + ch] = __ldg(&background[p * C + ch]);
                }
            } else {
                const TriInterp ti = load_interp(itp_b + face);
                const float4 g = exact::gbuffer_at(ti, col, row);
                const float* cols = vertex_colors + (size_t)b * d.V * C;
                if (CT == 4) {
                    const float4 c0 = __ldg(reinterpret_cast<const float4*>(cols) + ti.v0);
                    const float4 c1 = __ldg(reinterpret_cast<const float4*>(cols) + ti.v1);
                    const float4 c2 = __ldg(reinterpret_cast<const float4*>(cols) + ti.v2);
                    // c2 + b0*(c0-c2) + b1*(c1-c2): exact for equal vertex colours (tests/square_test.py)
                    float4 o;
                    o.x = fmaf(g.x, c0.x - c2.x, fmaf(g.y, c1.x - c2.x, c2.x));
                    o.y = fmaf(g.x, c0.y - c2.y, fmaf(g.y, c1.y - c2.y, c2.y));
                    o.z = fmaf(g.x, c0.z - c2.z, fmaf(g.y, c1.z - c2.z, c2.z));
                    o.w = fmaf(g.x, c0.w - c2.w, fmaf(g.y, c1.w - c2.w, c2.w));
                    reinterpret_cast<float4*>(pixels)[p] = o;
                } else {
                    const float* c0 = cols + (size_t)ti.v0 * C;
                    const float* c1 = cols + (size_t)ti.v1 * C;
                    const float* c2 = cols + (size_t)ti.v2 * C;
                    for (int ch = 0; ch < C; ++ch) {
                        const float a2 = __ldg(&c2[ch]);
                        pixels[p * C + ch] = fmaf(g.x, __ldg(&c0[ch]) - a2, fmaf(g.y, __ldg(&c1[ch]) - a2, a2));
                    }
                }
            }
        }
    }
}

cudaError_t launch_raster_forward(const float* vertices, const float* background, const float* vertex_colors, float* pixels,
                                  int32_t* face_ids_out, const Workspace& ws, const Dims& d, cudaStream_t stream,
                                  int* launches)
{
    const long long total_tiles = (long long)d.B * d.tiles;
    if (total_tiles == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((total_tiles + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    const float* v = vertices;
    ScopedKernelTimer timer(1, stream);
    const bool vec4 = d.C == 4 && ((uintptr_t)background % 16 == 0) && ((uintptr_t)pixels % 16 == 0) &&
                      ((uintptr_t)vertex_colors % 16 == 0);
    if (vec4)
        raster_kernel<0, 4><<<grid, WARPS_PER_BLOCK * 32, 0, stream>>>(v, background, vertex_colors, pixels, face_ids_out,
                                                                       nullptr, ws, d);
    else
        raster_kernel<0, 0><<<grid, WARPS_PER_BLOCK * 32, 0, stream>>>(v, background, vertex_colors, pixels, face_ids_out,
                                                                       nullptr, ws, d);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_raster_visibility(const float* vertices, int32_t* face_ids, float* gbuffer, const Workspace& ws, const Dims& d,
                                     cudaStream_t stream, int* launches)
{
    const long long total_tiles = (long long)d.B * d.tiles;
    if (total_tiles == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((total_tiles + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    raster_kernel<1, 0><<<grid, WARPS_PER_BLOCK * 32, 0, stream>>>(vertices, nullptr, nullptr, nullptr,
                                                                   face_ids, gbuffer, ws, d);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace dirt
