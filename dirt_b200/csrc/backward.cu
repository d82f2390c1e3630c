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

constexpr int BWD_WARPS_PER_BLOCK = 4;

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

// (a + b - c - d) * 3/32 + (e - f) * 10/32 in the specified order
__device__ __forceinline__ float scharr_comp(float a, float b, float c, float dd, float e, float f)
{
    const float X = __fsub_rn(__fsub_rn(__fadd_rn(a, b), c), dd);
    const float Y = __fsub_rn(e, f);
    return __fmaf_rn(Y, 0.3125f, __fmul_rn(X, 0.09375f));
}

__device__ __forceinline__ float l1(const float s[3])
{
    return __fadd_rn(__fadd_rn(fabsf(s[0]), fabsf(s[1])), fabsf(s[2]));
}

__global__ void __launch_bounds__(BWD_WARPS_PER_BLOCK * 32) backward_kernel(
    const float* __restrict__ vertices, const float* __restrict__ pixels, const float* __restrict__ grad_pixels,
    const int32_t* __restrict__ face_ids, float* __restrict__ grad_background, float* __restrict__ grad_vertices,
    float* __restrict__ grad_vertex_colors, Workspace ws, Dims d, GroupSpec groups)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tile_global = (long long)blockIdx.x * BWD_WARPS_PER_BLOCK + warp;
    if (tile_global >= (long long)d.B * d.tiles) return;
    const int b = (int)(tile_global / d.tiles);
    const int t = (int)(tile_global - (long long)b * d.tiles);
    const int ty = t / d.tiles_x, tx = t - ty * d.tiles_x;
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

cudaError_t launch_backward(const float* vertices, const float* pixels, const float* grad_pixels,
                            const int32_t* face_ids, float* grad_background, float* grad_vertices,
                            float* grad_vertex_colors, const Workspace& ws, const Dims& d, const GroupSpec& groups,
                            cudaStream_t stream, int* launches)
{
    cudaError_t e;
    if ((e = cudaMemsetAsync(grad_vertices, 0, sizeof(float) * (size_t)d.B * d.V * 4, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(grad_vertex_colors, 0, sizeof(float) * (size_t)d.B * d.V * d.C, stream)) != cudaSuccess) return e;
    const long long total_tiles = (long long)d.B * d.tiles;
    if (total_tiles == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((total_tiles + BWD_WARPS_PER_BLOCK - 1) / BWD_WARPS_PER_BLOCK);
    ScopedKernelTimer timer(2, stream);
    backward_kernel<<<grid, BWD_WARPS_PER_BLOCK * 32, 0, stream>>>(vertices, pixels, grad_pixels, face_ids, grad_background,
                                                                   grad_vertices, grad_vertex_colors, ws, d, groups);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace dirt
