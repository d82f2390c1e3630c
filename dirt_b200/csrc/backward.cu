// backward.cu -- the RasteriseGrad pass: restates assemble_grads
// (csrc/rasterise_grad_egl.cu:93-236) on top of the face-id visibility buffer.
//
// Per pixel: Scharr filter of `pixels` per channel group (frame-edge clamp), colour-gradient
// splat with the undilated barycentrics, background gradient, occluder-edge dilation from the
// +-1 neighbour along the dominant-gradient axis (dithered by (x+y)%2), position-gradient splat.
// Decision quantities (Scharr sums, their L1 norms, clip_w) follow the fixed fp32 operation order
// of DESIGN.md so that every discrete choice matches the oracle; accumulated values are ordinary fp32.
#include "common.cuh"

namespace dirt {

#ifndef DIRT_BWD_WARPS
#define DIRT_BWD_WARPS 4
#endif
constexpr int BWD_WARPS_PER_BLOCK = DIRT_BWD_WARPS;
#ifndef DIRT_ABLATE
#define DIRT_ABLATE 0   // timing experiments only: 1 = no per-face reduction, 2 = no Scharr/dilation, 3 = both
#endif
#ifndef DIRT_BWD_SMALL_FACE
#define DIRT_BWD_SMALL_FACE 12  // faces owning at most this many records in a tile are added directly (0: always reduce); profiles/r01_sweep_small_face.txt
#endif
#ifndef DIRT_BWD_SMALL_LANES
#define DIRT_BWD_SMALL_LANES 0  // > 0: the criterion is the number of lanes holding a record of the face instead
#endif
// 8x8 tiles (neighbours in x, = one 16x8 tile of the coverage flags) per warp.  Measured (profiles/r01_sweep_bwd_tiles.txt):
// pairs win for C = 3 (cfg5 backward 1.36 -> 1.26 ms: a background-only pair is copied with both tiles' loads in flight)
// and lose for C = 4, whose larger per-pixel state spills more when the tile body sits in a loop.
#ifndef DIRT_BWD_TILES_C4
#define DIRT_BWD_TILES_C4 1
#endif
#ifndef DIRT_BWD_TILES_C3
#define DIRT_BWD_TILES_C3 2
#endif
// warps per CTA (32 warps per SM resident either way: 32 / warps CTAs of <= 64 registers per thread)
#ifndef DIRT_BWD_WARPS_C4
#define DIRT_BWD_WARPS_C4 DIRT_BWD_WARPS
#endif
#ifndef DIRT_BWD_WARPS_C3
#define DIRT_BWD_WARPS_C3 DIRT_BWD_WARPS
#endif
template <int C> struct BwdTiles {
    static constexpr int value = (C == 4) ? DIRT_BWD_TILES_C4 : DIRT_BWD_TILES_C3;
    static constexpr int warps = (C == 4) ? DIRT_BWD_WARPS_C4 : DIRT_BWD_WARPS_C3;
};
#ifndef DIRT_BWD_PREFETCH_GP
#define DIRT_BWD_PREFETCH_GP 1   // measured: 0.429 -> 0.419 ms at cfg3 (profiles/r01_sweep_prefetch2.txt)
#endif
#ifndef DIRT_BWD_MIN_BLOCKS
#define DIRT_BWD_MIN_BLOCKS 8   // <= 64 registers: measured best (profiles/r01_sweep_bounds.txt)
#endif
constexpr int TILE = 8;   // backward tile edge: one warp per 8x8 tile, two pixels per lane

struct V3 { float x, y, z; };

// at(): nearest frame pixel for out-of-range taps; three components of the channel group starting at
// c0 (width n).  For n == 1 the reference reads "channels" 1 and 2 of a contiguous [B,H,W,1] tensor,
// i.e. the next two pixels in flat order (0 past the end of the tensor).
__device__ __forceinline__ V3 group_at(const float* __restrict__ pixels, int b, int r, int c, const Dims& d, int c0, int n)
{
    r = max(0, min(d.H - 1, r));
    c = max(0, min(d.W - 1, c));
    const size_t lin = ((size_t)b * d.H + r) * d.W + c;
    V3 v;
    if (n == 3) {
        const float* p = pixels + lin * d.C + c0;
        v.x = __ldg(p); v.y = __ldg(p + 1); v.z = __ldg(p + 2);
    } else {
        const size_t total = (size_t)d.B * d.H * d.W;
        v.x = __ldg(pixels + lin * d.C + c0);
        v.y = (lin + 1 < total) ? __ldg(pixels + (lin + 1) * d.C + c0) : 0.f;
        v.z = (lin + 2 < total) ? __ldg(pixels + (lin + 2) * d.C + c0) : 0.f;
    }
    return v;
}

// (a + b - c - d) * 3/32 + (e - f) * 10/32 in the order of operations of the reference's compiled kernel
// (csrc/rasterise_grad_egl.cu:126-127 as nvcc contracts it: FMUL (e-f)*10/32, then FFMA (a+b-c-d)*3/32 + that;
// profiles/r02_ref_assemble_grads_scharr_sass.txt)
__device__ __forceinline__ float scharr_comp(float a, float b, float c, float dd, float e, float f)
{
    const float X = __fsub_rn(__fsub_rn(__fadd_rn(a, b), c), dd);
    const float Y = __fsub_rn(e, f);
    return __fmaf_rn(X, 0.09375f, __fmul_rn(Y, 0.3125f));
}

__device__ __forceinline__ float l1(const float s[3])
{
    return __fadd_rn(__fadd_rn(fabsf(s[0]), fabsf(s[1])), fabsf(s[2]));
}

// Generic fallback: any channel count / any grouping, one atomic per term (slow, reference-shaped).
__global__ void __launch_bounds__(BWD_WARPS_PER_BLOCK * 32) backward_generic_kernel(
    const float* __restrict__ vertices, const float* __restrict__ pixels, const float* __restrict__ grad_pixels,
    const int32_t* __restrict__ face_ids, float* __restrict__ grad_background, float* __restrict__ grad_vertices,
    float* __restrict__ grad_vertex_colors, Workspace ws, Dims d, GroupSpec groups)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tile_global = (long long)blockIdx.x * BWD_WARPS_PER_BLOCK + warp;
    if (tile_global >= (long long)d.B * d.btiles) return;
    const int b = (int)(tile_global / d.btiles);
    const int t = (int)(tile_global - (long long)b * d.btiles);
    const int ty = t / d.btiles_x, tx = t - ty * d.btiles_x;
    const int col = tx * TILE + (lane & 7), row0 = ty * TILE + (lane >> 3) * 2;
    if (col >= d.W) return;

    const TriInterp* itp_b = ws.itp + (size_t)b * d.F;
    const float* verts = vertices + (size_t)b * d.V * 4;
    const int32_t* ids = face_ids + (size_t)b * d.H * d.W;
    float* gverts = grad_vertices + (size_t)b * d.V * 4;
    float* gcols = grad_vertex_colors + (size_t)b * d.V * d.C;
    const int C = d.C;
    const float inf = __int_as_float(0x7f800000);

    for (int pix = 0; pix < 2; ++pix) {
        const int row = row0 + pix;
        if (row >= d.H) break;
        const size_t p = ((size_t)b * d.H + row) * d.W + col;
        const int f_own = ids[row * d.W + col];
        TriInterp t_own;
        float4 g_own = make_float4(-1.f, -1.f, -1.f, inf);
        if (f_own >= 0) {
            t_own = load_interp(itp_b + f_own);
            g_own = exact::gbuffer_at(t_own, col, row);
            const int vid[3] = {t_own.v0, t_own.v1, t_own.v2};
            const float bary[3] = {g_own.x, g_own.y, g_own.z};
            for (int ch = 0; ch < C; ++ch) {
                const float gp = __ldg(&grad_pixels[p * C + ch]);
#pragma unroll
                for (int k = 0; k < 3; ++k) atomicAdd(&gcols[(size_t)vid[k] * C + ch], gp * bary[k]);
                grad_background[p * C + ch] = 0.f;
            }
        } else {
            for (int ch = 0; ch < C; ++ch) grad_background[p * C + ch] = __ldg(&grad_pixels[p * C + ch]);
        }

        const bool interior = col > 0 && row > 0 && col < d.W - 1 && row < d.H - 1;
        int c0 = 0;
        for (int gi = 0; gi < groups.n; ++gi) {
            const int n = groups.width[gi];
            // at(ox,oy) is image (row - oy, col + ox)
            const V3 a_mm = group_at(pixels, b, row + 1, col - 1, d, c0, n), a_mp = group_at(pixels, b, row - 1, col - 1, d, c0, n);
            const V3 a_pm = group_at(pixels, b, row + 1, col + 1, d, c0, n), a_pp = group_at(pixels, b, row - 1, col + 1, d, c0, n);
            const V3 a_m0 = group_at(pixels, b, row, col - 1, d, c0, n), a_p0 = group_at(pixels, b, row, col + 1, d, c0, n);
            const V3 a_0m = group_at(pixels, b, row + 1, col, d, c0, n), a_0p = group_at(pixels, b, row - 1, col, d, c0, n);
            float sx[3], sy[3];
            sx[0] = scharr_comp(a_mm.x, a_mp.x, a_pm.x, a_pp.x, a_m0.x, a_p0.x);
            sx[1] = scharr_comp(a_mm.y, a_mp.y, a_pm.y, a_pp.y, a_m0.y, a_p0.y);
            sx[2] = scharr_comp(a_mm.z, a_mp.z, a_pm.z, a_pp.z, a_m0.z, a_p0.z);
            sy[0] = scharr_comp(a_mm.x, a_pm.x, a_mp.x, a_pp.x, a_0m.x, a_0p.x);
            sy[1] = scharr_comp(a_mm.y, a_pm.y, a_mp.y, a_pp.y, a_0m.y, a_0p.y);
            sy[2] = scharr_comp(a_mm.z, a_pm.z, a_mp.z, a_pp.z, a_0m.z, a_0p.z);

            int f = f_own;
            float4 g = g_own;
            TriInterp tf = t_own;
            if (interior) {
                int dx = (l1(sx) > l1(sy)) ? 1 : 0, dy = 1 - dx;  // buffer (y-up) orientation
                if ((col + row) & 1) { dx = -dx; dy = -dy; }
                for (int attempt = 0; attempt < 2; ++attempt) {
                    const int nc = col + dx, nr = row - dy;
                    const int fn = ids[nr * d.W + nc];
                    if (fn >= 0) {
                        const TriInterp tn = load_interp(itp_b + fn);
                        const bool differs = (f_own < 0) || tn.v0 != t_own.v0 || tn.v1 != t_own.v1 || tn.v2 != t_own.v2;
                        const float4 gn = exact::gbuffer_at(tn, nc, nr);
                        if (differs && g_own.w > gn.w) {
                            g = gn; f = fn; tf = tn;
                            break;
                        }
                    }
                    dx = -dx; dy = -dy;
                }
            }
            if (f >= 0) {
                float dLdx = 0.f, dLdy = 0.f;
                for (int ch = 0; ch < n; ++ch) {
                    const float gp = __ldg(&grad_pixels[p * C + c0 + ch]);
                    dLdx += gp * sx[ch];
                    dLdy += gp * sy[ch];
                }
                const int vid[3] = {tf.v0, tf.v1, tf.v2};
                const float bary[3] = {g.x, g.y, g.z};
                float clip_x = 0.f, clip_y = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float2 xy = __ldg(reinterpret_cast<const float2*>(verts + (size_t)vid[k] * 4));
                    clip_x += bary[k] * xy.x;
                    clip_y += bary[k] * xy.y;
                }
                const float inv_w = 1.f / g.w;
                const float dxv_dxc = 0.5f * (float)d.W * inv_w, dyv_dyc = 0.5f * (float)d.H * inv_w;
                const float dxv_dwc = -dxv_dxc * clip_x * inv_w, dyv_dwc = -dyv_dyc * clip_y * inv_w;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float ax = dLdx * bary[k], ay = dLdy * bary[k];
                    atomicAdd(&gverts[(size_t)vid[k] * 4 + 0], ax * dxv_dxc);
                    atomicAdd(&gverts[(size_t)vid[k] * 4 + 1], ay * dyv_dyc);
                    atomicAdd(&gverts[(size_t)vid[k] * 4 + 3], ax * dxv_dwc + ay * dyv_dwc);
                }
            }
            c0 += n;
        }
    }
}


// =================================================================================================
// The tile kernel: one launch handles one channel group of width 3 or 1 on its slice [c0, c0+C) of the cs channels, or
// -- C = 4 -- the fused pair {3,1} of a 4-channel tensor.  Any channel count / grouping is a sequence of such launches
// (what the reference does at the Python level, dirt/rasterise_ops.py:86-108, without slicing or copying).
//
// One warp per 8x8 tile (or per pair of tiles, see BwdTiles), two vertically adjacent pixels per lane.
//  (0) 16x8 coverage flags written by the forward pass short-cut tiles that no face can reach:
//      grad_background = grad_pixels, nothing else.
//  (1) The tile of `pixels` plus its halo (10 rows x 12 columns: one pixel around for the Scharr taps and
//      two more to the right for the flat-order reads of 1-wide groups) is staged into shared memory with
//      cp.async, rows/columns clamped to the frame exactly as at() clamps its taps; all taps are then
//      LDS at constant offsets.  Tiles whose halo crosses the right frame edge read the taps from global
//      memory instead (there the flat-order reads wrap into the next row).
//  (2) Every per-pixel term of assemble_grads is a product  scalar * barycentric  destined for vertex k
//      of one face: C colour scalars (grad_pixels) keyed by the pixel's own face, three position scalars
//      (a, b, c = dL/d clip x, y, w of the fragment) keyed by the possibly dilated face.
//  (3) The warp loops over the distinct faces present in the tile (REDUX.MIN over the keys) and reduces the
//      3*(C+3) sums of each face over its 32 lanes with a transposed butterfly: at every shuffle step a lane
//      keeps half of the sums and hands the other half to its partner, so 21 sums need 11+6+3+2+1 = 23
//      shuffles instead of 21*5, and every lane ends up owning one finished sum: ONE warp-wide RED per
//      (face, tile) instead of the reference's atomicAdd per pixel per term
//      (csrc/rasterise_grad_egl.cu:139,227-229).  The first step works on operands (the barycentric
//      weights are stored per half-warp), the destination of a lane's sum comes from a constant-memory
//      table, and faces with only a few records in the tile skip the butterfly (vector REDs).
// =================================================================================================

constexpr int HALO_ROWS = TILE + 2;   // 10
constexpr int HALO_COLS = TILE + 4;   // 12: col-1 .. col+10

// 16-byte vector reduction to global memory (sm_90+): four fp32 adds in one RED
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct PixelTerms {      // everything one pixel contributes (fast path)
    int key_col;         // own face (-1: uncovered)
    int key_pos;         // face receiving the position gradient (-1: none)
    // barycentrics (c: undilated, for the colour terms; p: of the fragment receiving the position gradient), arranged for
    // the first butterfly step: A = the vertex whose sums this lane keeps (vertex 0 on lanes 0-15, vertex 1 on lanes
    // 16-31), B = the one it hands to its partner, 2 = vertex 2
    float cA, cB, c2;
    float pA, pB, p2;
};

struct Fragment {   // a face seen at a pixel: G-buffer entry + vertex ids
    int face;
    int v0, v1, v2;
    float4 g;       // bary0, bary1, bary2, clip_w
};

__device__ __forceinline__ Fragment fragment_at(const TriInterp* __restrict__ itp_b, int face, int col, int row)
{
    Fragment fr;
    fr.face = face;
    if (face >= 0) {
        const TriInterp t = load_interp(itp_b + face);
        fr.v0 = t.v0; fr.v1 = t.v1; fr.v2 = t.v2;
        fr.g = exact::gbuffer_at(t, col, row);
    } else {
        fr.v0 = fr.v1 = fr.v2 = -1;
        fr.g = make_float4(-1.f, -1.f, -1.f, __int_as_float(0x7f800000));
    }
    return fr;
}

// dilation (csrc/rasterise_grad_egl.cu:155-194).  The preferred neighbour offset (buffer, y-up orientation) depends on
// the group's Scharr sums; given the offset, the outcome depends on the visibility buffer only.
// Offsets are encoded as code = dx + 3*dy with (dx,dy) in {(1,0),(-1,0),(0,1),(0,-1)}: 1, -1, 3, -3.
__device__ __forceinline__ int dilation_code(const float (&sx)[3], const float (&sy)[3], int col, int row)
{
    int code = (l1(sx) > l1(sy)) ? 1 : 3;
    if ((col + row) & 1) code = -code;
    return code;
}

struct Neighbours { int left, right, up, down; };   // face ids at (col-1,row), (col+1,row), (col,row-1), (col,row+1)

__device__ __forceinline__ Fragment dilate(const Fragment& own, int code, const Neighbours nb,
                                           const TriInterp* __restrict__ itp_b, int col, int row, int& src)
{
    src = 0;
    // code = +-1: the horizontal pair (right, left); +-3: the vertical pair (up, down).  The first attempt looks along
    // the sign of `code`, the second one the other way.
    const bool horiz = (unsigned)(code + 1) <= 2u;
    const int f_plus = horiz ? nb.right : nb.up, f_minus = horiz ? nb.left : nb.down;
#pragma unroll
    for (int attempt = 0; attempt < 2; ++attempt) {
        const bool minus = (code < 0) != (attempt == 1);
        const int fn = minus ? f_minus : f_plus;
        if (fn >= 0 && fn != own.face) {
            // buffer offset (dx,dy) is image (col + dx, row - dy)
            const int step = minus ? -1 : 1;
            const Fragment n = fragment_at(itp_b, fn, horiz ? col + step : col, horiz ? row : row - step);
            const bool differs = (own.face < 0) || n.v0 != own.v0 || n.v1 != own.v1 || n.v2 != own.v2;
            if (differs && own.g.w > n.g.w) {
                src = attempt ? -code : code;   // distinguishes the four neighbours, never 0
                return n;
            }
        }
    }
    return own;
}

__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Scharr sums of the first group (width N0) from the staged tile.  lr/lc: pixel position inside the halo tile.
template <int C, int N0>
__device__ __forceinline__ void scharr_smem(const float* __restrict__ tile, int lr, int lc, float (&sx)[3], float (&sy)[3])
{
    // comp k of the tap at (lr+dr, lc+dc): channel k (N0 == 3) or channel 0 of the pixel k places to the right (N0 == 1)
    auto T = [&](int dr, int dc, int k) -> float {
        return (N0 == 3) ? tile[((lr + dr) * HALO_COLS + lc + dc) * C + k] : tile[((lr + dr) * HALO_COLS + lc + dc + k) * C];
    };
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        // at(ox,oy) is image (row - oy, col + ox)
        const float a_mm = T(+1, -1, k), a_mp = T(-1, -1, k), a_pm = T(+1, +1, k), a_pp = T(-1, +1, k);
        const float a_m0 = T(0, -1, k), a_p0 = T(0, +1, k), a_0m = T(+1, 0, k), a_0p = T(-1, 0, k);
        sx[k] = scharr_comp(a_mm, a_mp, a_pm, a_pp, a_m0, a_p0);
        sy[k] = scharr_comp(a_mm, a_pm, a_mp, a_pp, a_0m, a_0p);
    }
}

// C == 4, groups {3,1}: both groups from 15 vector taps (rows lr-1..lr+1, columns lc-1..lc+3).
// Group {0,1,2} uses .xyz of columns lc-1..lc+1; group {3} uses .w, comp k being the pixel k places to the right.
__device__ __forceinline__ void scharr_smem_c4(const float* __restrict__ tile, int lr, int lc, float (&sx)[3], float (&sy)[3],
                                               float (&sx1)[3], float (&sy1)[3])
{
    const float4* t4 = reinterpret_cast<const float4*>(tile) + (lr * HALO_COLS + lc);
    float4 up[5], mid[5], dn[5];   // image rows lr-1 (at oy=+1), lr, lr+1 (at oy=-1); columns lc-1 .. lc+3
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        up[j] = t4[-HALO_COLS + j - 1];
        mid[j] = t4[j - 1];
        dn[j] = t4[HALO_COLS + j - 1];
    }
    // at(ox,oy) is image (row - oy, col + ox):  a_mm = dn[0], a_mp = up[0], a_pm = dn[2], a_pp = up[2], a_m0 = mid[0],
    // a_p0 = mid[2], a_0m = dn[1], a_0p = up[1]
    sx[0] = scharr_comp(dn[0].x, up[0].x, dn[2].x, up[2].x, mid[0].x, mid[2].x);
    sx[1] = scharr_comp(dn[0].y, up[0].y, dn[2].y, up[2].y, mid[0].y, mid[2].y);
    sx[2] = scharr_comp(dn[0].z, up[0].z, dn[2].z, up[2].z, mid[0].z, mid[2].z);
    sy[0] = scharr_comp(dn[0].x, dn[2].x, up[0].x, up[2].x, dn[1].x, up[1].x);
    sy[1] = scharr_comp(dn[0].y, dn[2].y, up[0].y, up[2].y, dn[1].y, up[1].y);
    sy[2] = scharr_comp(dn[0].z, dn[2].z, up[0].z, up[2].z, dn[1].z, up[1].z);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        sx1[k] = scharr_comp(dn[k].w, up[k].w, dn[k + 2].w, up[k + 2].w, mid[k].w, mid[k + 2].w);
        sy1[k] = scharr_comp(dn[k].w, dn[k + 2].w, up[k].w, up[k + 2].w, dn[k + 1].w, up[k + 1].w);
    }
}

// global-memory taps (tiles at the right frame edge): three components of group [c0, c0+N0) at the clamped pixel
template <int C, int N0>
__device__ __forceinline__ void scharr_global(const float* __restrict__ pixels, int b, int row, int col, const Dims& d,
                                              int c0, float (&sx)[3], float (&sy)[3])
{
    const V3 a_mm = group_at(pixels, b, row + 1, col - 1, d, c0, N0), a_mp = group_at(pixels, b, row - 1, col - 1, d, c0, N0);
    const V3 a_pm = group_at(pixels, b, row + 1, col + 1, d, c0, N0), a_pp = group_at(pixels, b, row - 1, col + 1, d, c0, N0);
    const V3 a_m0 = group_at(pixels, b, row, col - 1, d, c0, N0), a_p0 = group_at(pixels, b, row, col + 1, d, c0, N0);
    const V3 a_0m = group_at(pixels, b, row + 1, col, d, c0, N0), a_0p = group_at(pixels, b, row - 1, col, d, c0, N0);
    sx[0] = scharr_comp(a_mm.x, a_mp.x, a_pm.x, a_pp.x, a_m0.x, a_p0.x);
    sx[1] = scharr_comp(a_mm.y, a_mp.y, a_pm.y, a_pp.y, a_m0.y, a_p0.y);
    sx[2] = scharr_comp(a_mm.z, a_mp.z, a_pm.z, a_pp.z, a_m0.z, a_p0.z);
    sy[0] = scharr_comp(a_mm.x, a_pm.x, a_mp.x, a_pp.x, a_0m.x, a_0p.x);
    sy[1] = scharr_comp(a_mm.y, a_pm.y, a_mp.y, a_pp.y, a_0m.y, a_0p.y);
    sy[2] = scharr_comp(a_mm.z, a_pm.z, a_mp.z, a_pp.z, a_0m.z, a_0p.z);
}

// Transposed butterfly over the lanes that differ in bits `bit`, bit/2, ..., 1: at every step a lane keeps half of its
// sums and hands the other half to its partner.  After STEPS steps each lane is left with OUT = ceil(N / 2^STEPS) sums.
template <int N, int STEPS>
struct TransposedReduce {
    static constexpr int H = (N + 1) / 2;
    static constexpr int OUT = TransposedReduce<H, STEPS - 1>::OUT;
    static __device__ __forceinline__ void run(float (&v)[N], int lane, int bit, float (&out)[OUT])
    {
        float w[H];
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const float hi = (H + i < N) ? v[H + i] : 0.f;
            const float send = up ? v[i] : hi;
            const float keep = up ? hi : v[i];
            w[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
        TransposedReduce<H, STEPS - 1>::run(w, lane, bit >> 1, out);
    }
    // global index of out[0] on this lane: out[i] is sum base+i (entries beyond the valid range are padding)
    static __host__ __device__ constexpr int base(int lane, int bit)
    {
        return ((lane & bit) ? H : 0) + TransposedReduce<H, STEPS - 1>::base(lane, bit >> 1);
    }
    // is out[i] a real sum?  (checks the padding introduced at every level)
    static __host__ __device__ constexpr bool valid(int lane, int bit, int i)
    {
        // p: position inside this level's kept half
        return TransposedReduce<H, STEPS - 1>::valid(lane, bit >> 1, i) &&
               (((lane & bit) ? H : 0) + TransposedReduce<H, STEPS - 1>::base(lane, bit >> 1) + i < N);
    }
};
template <int N>
struct TransposedReduce<N, 0> {
    static constexpr int OUT = N;
    static __device__ __forceinline__ void run(float (&v)[N], int, int, float (&out)[N])
    {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = v[i];
    }
    static __host__ __device__ constexpr int base(int, int) { return 0; }
    static __host__ __device__ constexpr bool valid(int, int, int i) { return i < N; }
};

// Which finished sum a lane owns after the butterfly, as a destination: bits 0-1 = vertex k of the face, bits 2-3 =
// component inside the vertex's row, bit 4 = row of grad_vertices (else grad_vertex_colors); -1 = none.  A table in
// constant memory because the compiler, short of registers, otherwise recomputes the index arithmetic (~40 instructions)
// for every face of every tile.
template <int C>
struct OwnerTable {
    int meta[32];
};
template <int C>
constexpr OwnerTable<C> make_owner_table()
{
    // layout after the first (operand-level) butterfly step, see the per-face reduction: NS sums of vertex 0 (lanes
    // 0-15) or vertex 1 (lanes 16-31), then the lower / upper half of the NS sums of vertex 2
    constexpr int NS = C + 3, H2 = (NS + 1) / 2;
    using Red = TransposedReduce<NS + H2, 4>;
    OwnerTable<C> t{};
    for (int lane = 0; lane < 32; ++lane) {
        int m = -1;
        if (Red::valid(lane, 8, 0)) {
            const int q = Red::base(lane, 8);
            const bool up = (lane & 16) != 0;
            const int k = q < NS ? (up ? 1 : 0) : 2;
            const int j = q < NS ? q : (q - NS) + (up ? H2 : 0);
            if (j < NS) {
                const bool pos = j >= C;
                const int comp = pos ? (j - C == 2 ? 3 : j - C) : j;   // a, b, c go to x, y, w of the vertex row
                m = k | (comp << 2) | (pos ? 16 : 0);
            }
        }
        t.meta[lane] = m;
    }
    return t;
}
__constant__ OwnerTable<1> c_owner1 = make_owner_table<1>();
__constant__ OwnerTable<3> c_owner3 = make_owner_table<3>();
__constant__ OwnerTable<4> c_owner4 = make_owner_table<4>();
template <int C>
__device__ __forceinline__ int owner_meta(int lane)
{
    return C == 1 ? c_owner1.meta[lane] : C == 3 ? c_owner3.meta[lane] : c_owner4.meta[lane];
}

template <int C, int BWD_TILES, int NW>
__global__ void __launch_bounds__(NW * 32, DIRT_BWD_MIN_BLOCKS * DIRT_BWD_WARPS / NW) backward_tile_kernel(
    const float* __restrict__ vertices, const float* __restrict__ pixels, const float* __restrict__ grad_pixels,
    const int32_t* __restrict__ face_ids, float* __restrict__ grad_background, float* __restrict__ grad_vertices,
    float* __restrict__ grad_vertex_colors, Workspace ws, Dims d, const unsigned char* __restrict__ tile_flags,
    int cs, int c0)   // cs: channels per pixel in the tensors, c0: first channel of the group this launch handles (width C)
{
    constexpr int NS = C + 3;                  // scalars per pixel: C colour + (a,b,c)
    constexpr int N0 = (C == 1) ? 1 : 3;       // width of the first group
    constexpr bool TWO_GROUPS = (C == 4);      // {3,1}
    constexpr int REACH = (C == 3) ? 1 : 3;    // columns to the right of the pixel that its taps read

    __shared__ __align__(16) float tile_all[NW][HALO_ROWS * HALO_COLS * C];

    // grid: x = groups of BWD_WARPS_PER_BLOCK * BWD_TILES tiles along a tile row, y = tile row, z = image
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int txb = (blockIdx.x * NW + warp) * BWD_TILES, ty = blockIdx.y;
    if (txb >= d.btiles_x) return;
    const int trow0 = ty * TILE;
    const int lcol = lane & 7, lrow0 = (lane >> 3) * 2;
    const int row0 = trow0 + lrow0;
    const int H = d.H, W = d.W;
    float* tile = tile_all[warp];

    for (int b = blockIdx.z; b < d.B; b += gridDim.z) {
    const TriInterp* itp_b = ws.itp + (size_t)b * d.F;
    const float* verts = vertices + (size_t)b * d.V * 4;
    const int32_t* ids = face_ids + (size_t)b * H * W;
    float* gverts = grad_vertices + (size_t)b * d.V * 4;
    float* gcols = grad_vertex_colors + (size_t)b * d.V * cs + c0;
    const size_t img = (size_t)b * H * W;
    const float halfW = 0.5f * (float)W, halfH = 0.5f * (float)H;

    if (BWD_TILES == 2) {
    // The two 8x8 tiles of this warp are exactly one 16x8 tile of the forward pass's coverage flags.
#if DIRT_BWD_PREFETCH_GP
    // start the DRAM read of the pair's grad_pixels while the flag is still on its way
    if (C == 4 && lane < 16 && trow0 + (lane >> 1) < H && (txb + (lane & 1)) * TILE < W)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(grad_pixels + (img + (size_t)(trow0 + (lane >> 1)) * W + (txb + (lane & 1)) * TILE) * 4));
#endif
    // ---- background-only pair: nothing can reach an unflagged tile, grad_background = grad_pixels and we are done.
    // Both tiles' loads are issued before the first store: such a pair is pure latency (flag -> load -> store).
    if (tile_flags != nullptr && tile_flags[(size_t)b * d.tiles + ty * d.tiles_x + (txb >> 1)] == 0) {
        if (C == 4) {
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = row0 + (i & 1), col = (txb + (i >> 1)) * TILE + lcol;
                if (col < W && row < H) v[i] = __ldg(reinterpret_cast<const float4*>(grad_pixels) + img + (size_t)row * W + col);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = row0 + (i & 1), col = (txb + (i >> 1)) * TILE + lcol;
                if (col < W && row < H) reinterpret_cast<float4*>(grad_background)[img + (size_t)row * W + col] = v[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = row0 + (i & 1), col = (txb + (i >> 1)) * TILE + lcol;
                if (col >= W || row >= H) continue;
                const size_t p = img + (size_t)row * W + col;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) grad_background[p * cs + c0 + ch] = __ldg(grad_pixels + p * cs + c0 + ch);
            }
        }
        continue;
    }
    }
    for (int sub = 0; sub < BWD_TILES; ++sub) {
    const int tx = txb + sub;
    if (tx >= d.btiles_x) break;
    const int tcol0 = tx * TILE;
    const int col = tcol0 + lcol;

    if (BWD_TILES == 1) {
#if DIRT_BWD_PREFETCH_GP
    // start the DRAM read of this tile's grad_pixels while the tile flag is still on its way
    if (C == 4 && lane < 8 && trow0 + lane < H && tcol0 < W)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(grad_pixels + (img + (size_t)(trow0 + lane) * W + tcol0) * 4));
#endif
    // ---- background-only tiles: the forward pass flagged every 16x8 tile that shows a face or touches one that does.
    // Nothing can reach an unflagged tile: grad_background = grad_pixels and we are done.
    if (tile_flags != nullptr && tile_flags[(size_t)b * d.tiles + ty * d.tiles_x + (tx >> 1)] == 0) {
#pragma unroll
        for (int pix = 0; pix < 2; ++pix) {
            const int row = row0 + pix;
            if (col >= W || row >= H) continue;
            const size_t p = img + (size_t)row * W + col;
            if (C == 4) reinterpret_cast<float4*>(grad_background)[p] = __ldg(reinterpret_cast<const float4*>(grad_pixels) + p);
            else {
#pragma unroll
                for (int ch = 0; ch < C; ++ch) grad_background[p * cs + c0 + ch] = __ldg(grad_pixels + p * cs + c0 + ch);
            }
        }
        continue;
    }
    }

    // ---- this lane's two pixels and their six outer neighbours in the visibility buffer; grad_pixels ------------
    // rows row0-1 .. row0+2 at column col, and columns col-1 / col+1 at rows row0, row0+1 (-1 outside the frame)
    int id_up, id_0, id_1, id_dn, id_l0, id_r0, id_l1, id_r1;
    if (tcol0 > 0 && trow0 > 0 && tcol0 + TILE < W && trow0 + TILE < H) {   // warp-uniform: no bounds checks needed
        const int32_t* p0 = ids + row0 * W + col;
        id_up = __ldg(p0 - W); id_0 = __ldg(p0); id_1 = __ldg(p0 + W); id_dn = __ldg(p0 + 2 * W);
        id_l0 = __ldg(p0 - 1); id_r0 = __ldg(p0 + 1); id_l1 = __ldg(p0 + W - 1); id_r1 = __ldg(p0 + W + 1);
    } else {
        const bool col_in = col < W;
        auto id_at = [&](int r, int c) -> int { return (r >= 0 && r < H && c >= 0 && c < W) ? __ldg(&ids[r * W + c]) : -1; };
        id_up = col_in ? id_at(row0 - 1, col) : -1; id_0 = col_in ? id_at(row0, col) : -1;
        id_1 = col_in ? id_at(row0 + 1, col) : -1; id_dn = col_in ? id_at(row0 + 2, col) : -1;
        id_l0 = id_at(row0, col - 1); id_r0 = id_at(row0, col + 1);
        id_l1 = id_at(row0 + 1, col - 1); id_r1 = id_at(row0 + 1, col + 1);
    }
    const Neighbours nb0 = {id_l0, id_r0, id_up, id_1};
    const Neighbours nb1 = {id_l1, id_r1, id_0, id_dn};
    float gp[2][C];
#pragma unroll
    for (int pix = 0; pix < 2; ++pix) {
        const int row = row0 + pix;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) gp[pix][ch] = 0.f;
        if (col >= W || row >= H) continue;
        const size_t p = img + (size_t)row * W + col;
        if (C == 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(grad_pixels) + p);
            gp[pix][0] = v.x; gp[pix][1 % C] = v.y; gp[pix][2 % C] = v.z; gp[pix][3 % C] = v.w;
        } else {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) gp[pix][ch] = __ldg(grad_pixels + p * cs + c0 + ch);
        }
    }

    // ---- grad_background; does anything reach this tile? -------------------------------------------------------
    int f_own[2];
    bool near = false;
#pragma unroll
    for (int pix = 0; pix < 2; ++pix) {
        const int row = row0 + pix;
        f_own[pix] = -2;
        if (col >= W || row >= H) continue;
        const size_t p = img + (size_t)row * W + col;
        const int f = pix ? id_1 : id_0;
        // grad_background: grad_pixels where uncovered, 0 elsewhere (:143-148, memset :247)
        if (C == 4) {
            reinterpret_cast<float4*>(grad_background)[p] =
                f < 0 ? make_float4(gp[pix][0], gp[pix][1 % C], gp[pix][2 % C], gp[pix][3 % C]) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) grad_background[p * cs + c0 + ch] = f < 0 ? gp[pix][ch] : 0.f;
        }
        // own coverage, or (interior pixels only) a covered 4-neighbour that could dilate into this pixel
        bool n = f >= 0;
        if (!n && col > 0 && row > 0 && col < W - 1 && row < H - 1)
            n = pix ? ((nb1.left & nb1.right & nb1.up & nb1.down) >= 0) : ((nb0.left & nb0.right & nb0.up & nb0.down) >= 0);   // any of the four non-negative
        if (n) f_own[pix] = f;   // -2: nothing can reach this pixel
        near = near || n;
    }
    if (!__any_sync(0xffffffffu, near)) continue;

    // ---- stage the tile of `pixels` (+halo) ----------------------------------------------------------
    const bool staged = (tcol0 + TILE - 1 + REACH) <= W - 1;   // warp-uniform
    if (staged) {
        constexpr int ELEMS = HALO_ROWS * HALO_COLS;
        for (int e = lane; e < ELEMS; e += 32) {
            const int hr = e / HALO_COLS, hc = e - hr * HALO_COLS;
            const int r = max(0, min(H - 1, trow0 - 1 + hr));
            const int c = max(0, min(W - 1, tcol0 - 1 + hc));
            const float* src = pixels + (img + (size_t)r * W + c) * (C == 4 ? 4 : cs) + (C == 4 ? 0 : c0);
            if (C == 4) cp_async_16(tile + e * 4, src);
            else {
#pragma unroll
                for (int ch = 0; ch < C; ++ch) cp_async_4(tile + e * C + ch, src + ch);
            }
        }
    }

    // ---- own fragments (overlaps the staging copies) ---------------------------------------------------
    Fragment own[2];
    own[0] = fragment_at(itp_b, max(f_own[0], -1), col, row0);
    if (f_own[1] == f_own[0] && f_own[0] >= 0) {
        // same face one row down: only the G-buffer entry changes
        own[1] = own[0];
        own[1].g = exact::gbuffer_at(load_interp(itp_b + f_own[1]), col, row0 + 1);
    } else {
        own[1] = fragment_at(itp_b, max(f_own[1], -1), col, row0 + 1);
    }
    if (staged) cp_async_wait_all();
    __syncwarp();

    // ---- per-pixel terms -------------------------------------------------------------------------------------
    const bool upper = (lane & 16) != 0;   // which half of the warp: decides the arrangement of the weights in PixelTerms
    PixelTerms term[2];
    float sc[2][NS];    // scalars: [0,C) grad_pixels, C..C+2 = a,b,c
#pragma unroll
    for (int pix = 0; pix < 2; ++pix) {
        const int row = row0 + pix;
        PixelTerms& T = term[pix];
        T.key_col = T.key_pos = -1;
        T.cA = T.cB = T.c2 = T.pA = T.pB = T.p2 = 0.f;
#pragma unroll
        for (int i = 0; i < NS; ++i) sc[pix][i] = (i < C) ? gp[pix][i % C] : 0.f;
        if (f_own[pix] == -2) continue;
        const Fragment& me = own[pix];
        const bool interior = col > 0 && row > 0 && col < W - 1 && row < H - 1;
        T.key_col = me.face;
        if (me.face >= 0) { T.cA = upper ? me.g.y : me.g.x; T.cB = upper ? me.g.x : me.g.y; T.c2 = me.g.z; }
#if DIRT_ABLATE >= 2
        T.key_pos = me.face; T.pA = T.cA; T.pB = T.cB; T.p2 = T.c2;
        sc[pix][C] = gp[pix][0]; sc[pix][C + 1] = gp[pix][0]; sc[pix][C + 2] = gp[pix][0];
        continue;
#endif

        float sx[3], sy[3], sx1[3], sy1[3];
        if (staged) {
            if (C == 4) scharr_smem_c4(tile, lrow0 + pix + 1, lcol + 1, sx, sy, sx1, sy1);
            else scharr_smem<C, N0>(tile, lrow0 + pix + 1, lcol + 1, sx, sy);
        } else {
            scharr_global<C, N0>(pixels, b, row, col, d, c0, sx, sy);
            if (TWO_GROUPS) scharr_global<C, 1>(pixels, b, row, col, d, c0 + 3, sx1, sy1);
        }
        // C = 4: reduce the Scharr sums to what the rest needs (two gradient scalars and a dilation code per group) BEFORE
        // the dilation's loads -- twelve sums fewer live across them (0.4225 -> 0.4077 ms at cfg3).  The single-group
        // kernels are faster with the sums consumed after the dilation (cfg5: 1.258 vs 1.308 ms), so they keep that order.
        float dLdx = 0.f, dLdy = 0.f;
        float gx1 = 0.f, gy1 = 0.f;
        int code0 = 0, code1 = 0, src0 = 0;
        Fragment pos0 = me;
        if (TWO_GROUPS) {
#pragma unroll
            for (int ch = 0; ch < N0; ++ch) { dLdx += gp[pix][ch] * sx[ch]; dLdy += gp[pix][ch] * sy[ch]; }
            code0 = interior ? dilation_code(sx, sy, col, row) : 0;
            gx1 = gp[pix][3 % C] * sx1[0]; gy1 = gp[pix][3 % C] * sy1[0];
            code1 = interior ? dilation_code(sx1, sy1, col, row) : 0;
            if (interior) pos0 = dilate(me, code0, pix ? nb1 : nb0, itp_b, col, row, src0);
        } else {
            if (interior) {
                code0 = dilation_code(sx, sy, col, row);
                pos0 = dilate(me, code0, pix ? nb1 : nb0, itp_b, col, row, src0);
            }
#pragma unroll
            for (int ch = 0; ch < N0; ++ch) { dLdx += gp[pix][ch] * sx[ch]; dLdy += gp[pix][ch] * sy[ch]; }
        }

        auto position_terms = [&](const Fragment& fr, float gx, float gy, float& a, float& bb, float& cc) {
            // a = dL/dx_clip, b = dL/dy_clip, c = dL/dw_clip of the fragment (:196-232)
            const float2 p0 = __ldg(reinterpret_cast<const float2*>(verts + (size_t)fr.v0 * 4));
            const float2 p1 = __ldg(reinterpret_cast<const float2*>(verts + (size_t)fr.v1 * 4));
            const float2 p2 = __ldg(reinterpret_cast<const float2*>(verts + (size_t)fr.v2 * 4));
            const float clip_x = fr.g.x * p0.x + fr.g.y * p1.x + fr.g.z * p2.x;
            const float clip_y = fr.g.x * p0.y + fr.g.y * p1.y + fr.g.z * p2.y;
            const float inv_w = __fdividef(1.0f, fr.g.w);
            a = gx * halfW * inv_w;
            bb = gy * halfH * inv_w;
            cc = -(a * clip_x + bb * clip_y) * inv_w;
        };

        if (TWO_GROUPS) {
            // the second group dilates to the same fragment whenever it prefers the same neighbour (the usual case):
            // the outcome of a dilation depends on the offset and on the visibility buffer only
            if (code1 == code0) {
                dLdx += gx1; dLdy += gy1;
            } else {
                int src1;
                const Fragment pos1 = dilate(me, code1, pix ? nb1 : nb0, itp_b, col, row, src1);
                if (pos1.face >= 0) {
                    if (pos1.face == pos0.face && src1 == src0) {
                        dLdx += gx1; dLdy += gy1;
                    } else {
                        // the two groups dilated differently (rare): this group's terms go out one by one
                        float a1, b1, c1;
                        position_terms(pos1, gx1, gy1, a1, b1, c1);
                        const int vid[3] = {pos1.v0, pos1.v1, pos1.v2};
                        const float bary[3] = {pos1.g.x, pos1.g.y, pos1.g.z};
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            atomicAdd(&gverts[(size_t)vid[k] * 4 + 0], a1 * bary[k]);
                            atomicAdd(&gverts[(size_t)vid[k] * 4 + 1], b1 * bary[k]);
                            atomicAdd(&gverts[(size_t)vid[k] * 4 + 3], c1 * bary[k]);
                        }
                    }
                }
            }
        }
        if (pos0.face >= 0) {
            T.key_pos = pos0.face;
            T.pA = upper ? pos0.g.y : pos0.g.x; T.pB = upper ? pos0.g.x : pos0.g.y; T.p2 = pos0.g.z;
            position_terms(pos0, dLdx, dLdy, sc[pix][C], sc[pix][C + 1], sc[pix][C + 2]);
        }
    }

    // ---- per-face reduction -------------------------------------------------------------------------------------
    // One iteration per distinct face of the tile (REDUX.MIN over the keys): the 3*(C+3) sums of the face are reduced
    // over the 32 lanes with a transposed butterfly (each lane ends up owning one finished sum) and leave the SM as ONE
    // warp-wide RED.  Faces that own only a few records in this tile skip the butterfly: their records are added
    // directly, all such faces of the tile together, in one pass of vector REDs at the end.
#if DIRT_ABLATE != 1
    {
        const int owner = owner_meta<C>(lane);
        // destination of this lane's finished sum: component (owner >> 2) & 3 of row `vid` of grad_vertices / grad_vertex_colors
        float* const owner_row = ((owner & 16) ? gverts : gcols) + ((owner >> 2) & 3);
        const int owner_stride = (owner & 16) ? 4 : (C == 4 ? 4 : cs);
        const int kc0 = term[0].key_col, kc1 = term[1].key_col, kp0 = term[0].key_pos, kp1 = term[1].key_pos;
        unsigned direct = 0;   // bit 0/1: colour record of pixel 0/1, bit 2/3: position record of pixel 0/1
        int last = -1;
        while (true) {
            // next distinct face key greater than `last`
            unsigned cand = 0x7fffffffu;
            if (kc0 > last) cand = min(cand, (unsigned)kc0);
            if (kc1 > last) cand = min(cand, (unsigned)kc1);
            if (kp0 > last) cand = min(cand, (unsigned)kp0);
            if (kp1 > last) cand = min(cand, (unsigned)kp1);
            const unsigned fmin = __reduce_min_sync(0xffffffffu, cand);
            if (fmin == 0x7fffffffu) break;
            const int f = (int)fmin;
            last = f;
            const bool mc0 = kc0 == f, mc1 = kc1 == f, mp0 = kp0 == f, mp1 = kp1 == f;
#if DIRT_BWD_SMALL_FACE > 0
#if DIRT_BWD_SMALL_LANES > 0
            // lanes holding a record of this face (each holds up to four): one vote instead of four
            const int records = __popc(__ballot_sync(0xffffffffu, mc0 | mc1 | mp0 | mp1));
            if (records <= DIRT_BWD_SMALL_LANES) {
#else
            const int records = __popc(__ballot_sync(0xffffffffu, mc0)) + __popc(__ballot_sync(0xffffffffu, mc1)) +
                                __popc(__ballot_sync(0xffffffffu, mp0)) + __popc(__ballot_sync(0xffffffffu, mp1));
            if (records <= DIRT_BWD_SMALL_FACE) {
#endif
                direct |= (mc0 ? 1u : 0u) | (mc1 ? 2u : 0u) | (mp0 ? 4u : 0u) | (mp1 ? 8u : 0u);
                continue;
            }
#endif
            // First butterfly step at operand level: this lane keeps the NS sums of vertex A and hands those of vertex B
            // to lane^16 (the weights were arranged per half-warp when they were stored), and the halves swap one half
            // each of the NS sums of vertex 2.  The remaining steps are the generic transposed butterfly.
            constexpr int H2 = (NS + 1) / 2;
            float keep[NS], send[NS], third[NS];
#pragma unroll
            for (int pix = 0; pix < 2; ++pix) {
                const PixelTerms& T = term[pix];
                const bool mc = pix ? mc1 : mc0, mp = pix ? mp1 : mp0;
                const float wcA = mc ? T.cA : 0.f, wcB = mc ? T.cB : 0.f, wc2 = mc ? T.c2 : 0.f;
                const float wpA = mp ? T.pA : 0.f, wpB = mp ? T.pB : 0.f, wp2 = mp ? T.p2 : 0.f;
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const float a = (j < C) ? wcA : wpA, bb = (j < C) ? wcB : wpB, cc = (j < C) ? wc2 : wp2;
                    keep[j] = (pix == 0) ? a * sc[pix][j] : fmaf(a, sc[pix][j], keep[j]);
                    send[j] = (pix == 0) ? bb * sc[pix][j] : fmaf(bb, sc[pix][j], send[j]);
                    third[j] = (pix == 0) ? cc * sc[pix][j] : fmaf(cc, sc[pix][j], third[j]);
                }
            }
            float v[NS + H2];
#pragma unroll
            for (int j = 0; j < NS; ++j) v[j] = keep[j] + __shfl_xor_sync(0xffffffffu, send[j], 16);
#pragma unroll
            for (int i = 0; i < H2; ++i) {
                const float hi = (H2 + i < NS) ? third[H2 + i] : 0.f;
                const float snd = upper ? third[i] : hi, kp = upper ? hi : third[i];
                v[NS + i] = kp + __shfl_xor_sync(0xffffffffu, snd, 16);
            }
            float total[1];
            TransposedReduce<NS + H2, 4>::run(v, lane, 8, total);
            if (owner >= 0) {
                const int4 q = __ldg(reinterpret_cast<const int4*>(itp_b + f) + 2);   // {sC, v0, v1, v2}
                const int vid = (owner & 1) ? q.z : ((owner & 2) ? q.w : q.y);
                atomicAdd(owner_row + (size_t)vid * owner_stride, total[0]);
            }
        }
#if DIRT_BWD_SMALL_FACE > 0
        if (__any_sync(0xffffffffu, direct != 0u)) {
#pragma unroll
            for (int rec = 0; rec < 4; ++rec) {
                if (!(direct & (1u << rec))) continue;
                const int pix = rec & 1;
                const PixelTerms& T = term[pix];
                const bool colour = rec < 2;
                const int f = colour ? T.key_col : T.key_pos;
                const int4 q = __ldg(reinterpret_cast<const int4*>(itp_b + f) + 2);
                const int vid[3] = {q.y, q.z, q.w};
                const float wA = colour ? T.cA : T.pA, wB = colour ? T.cB : T.pB;
                const float w[3] = {upper ? wB : wA, upper ? wA : wB, colour ? T.c2 : T.p2};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (colour) {
                        if (C == 4) red_add_v4(gcols + (size_t)vid[k] * 4, w[k] * sc[pix][0], w[k] * sc[pix][1 % NS], w[k] * sc[pix][2 % NS], w[k] * sc[pix][3 % NS]);
                        else {
#pragma unroll
                            for (int j = 0; j < C; ++j) atomicAdd(gcols + (size_t)vid[k] * cs + j, w[k] * sc[pix][j]);
                        }
                    } else {
                        red_add_v4(gverts + (size_t)vid[k] * 4, w[k] * sc[pix][C], w[k] * sc[pix][C + 1], 0.f, w[k] * sc[pix][C + 2]);
                    }
                }
            }
        }
#endif
    }
#endif
    }   // sub
    }   // b
}

cudaError_t launch_backward(const float* vertices, const float* pixels, const float* grad_pixels,
                            const int32_t* face_ids, float* grad_background, float* grad_vertices,
                            float* grad_vertex_colors, const Workspace& ws, const Dims& d, const GroupSpec& groups,
                            bool tile_flags_valid, cudaStream_t stream, int* launches)
{
    cudaError_t e;
    if ((e = cudaMemsetAsync(grad_vertices, 0, sizeof(float) * (size_t)d.B * d.V * 4, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(grad_vertex_colors, 0, sizeof(float) * (size_t)d.B * d.V * d.C, stream)) != cudaSuccess) return e;
    const long long total_tiles = (long long)d.B * d.btiles;
    if (total_tiles == 0) return cudaSuccess;
    ScopedKernelTimer timer(2, stream);
    // C == 4 with the default grouping {3,1} and 16-byte aligned tensors: one fused launch.  Everything else: one launch
    // per channel group (width 3 or 1) on its slice of the channels -- what the reference does at the Python level
    // (dirt/rasterise_ops.py:86-108), except that nothing is sliced or copied and grad_vertices accumulates in place.
    const bool fused4 = d.C == 4 && groups.n == 2 && groups.width[0] == 3 && groups.width[1] == 1 &&
                        (((uintptr_t)pixels | (uintptr_t)grad_pixels | (uintptr_t)grad_background | (uintptr_t)grad_vertex_colors) % 16 == 0);
    const unsigned char* flags = tile_flags_valid ? ws.tile_flags : nullptr;
    auto grid_for = [&](int tiles_per_warp, int warps) {
        return dim3((unsigned)((d.btiles_x + warps * tiles_per_warp - 1) / (warps * tiles_per_warp)), (unsigned)d.btiles_y,
                    (unsigned)min(d.B, 65535));
    };
#ifdef DIRT_FORCE_GENERIC_BACKWARD   // the reference-shaped kernel (one atomic per term), kept for debugging
    {
        const unsigned grid = (unsigned)((total_tiles + BWD_WARPS_PER_BLOCK - 1) / BWD_WARPS_PER_BLOCK);
        backward_generic_kernel<<<grid, BWD_WARPS_PER_BLOCK * 32, 0, stream>>>(vertices, pixels, grad_pixels, face_ids, grad_background,
                                                                              grad_vertices, grad_vertex_colors, ws, d, groups);
        ++*launches;
        return cudaGetLastError();
    }
#endif
    if (fused4) {
        backward_tile_kernel<4, BwdTiles<4>::value, BwdTiles<4>::warps>
            <<<grid_for(BwdTiles<4>::value, BwdTiles<4>::warps), BwdTiles<4>::warps * 32, 0, stream>>>(
                vertices, pixels, grad_pixels, face_ids, grad_background, grad_vertices, grad_vertex_colors, ws, d, flags, 4, 0);
        ++*launches;
        return cudaGetLastError();
    }
    int c0 = 0;
    for (int g = 0; g < groups.n; ++g) {
        if (groups.width[g] == 3)
            backward_tile_kernel<3, BwdTiles<3>::value, BwdTiles<3>::warps>
                <<<grid_for(BwdTiles<3>::value, BwdTiles<3>::warps), BwdTiles<3>::warps * 32, 0, stream>>>(
                    vertices, pixels, grad_pixels, face_ids, grad_background, grad_vertices, grad_vertex_colors, ws, d, flags, d.C, c0);
        else
            backward_tile_kernel<1, BwdTiles<1>::value, BwdTiles<1>::warps>
                <<<grid_for(BwdTiles<1>::value, BwdTiles<1>::warps), BwdTiles<1>::warps * 32, 0, stream>>>(
                    vertices, pixels, grad_pixels, face_ids, grad_background, grad_vertices, grad_vertex_colors, ws, d, flags, d.C, c0);
        c0 += groups.width[g];
        ++*launches;
        const cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) return le;
    }
    return cudaSuccess;
}

}  // namespace dirt
