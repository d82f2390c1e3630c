// common.cuh -- record layouts, workspace carving and the exactly-specified arithmetic shared by
// the sm_100a kernels of libdirt_b200.so.
//
// The arithmetic in `exact::` implements the visibility specification S1-S7 / H1-H3 / G written
// out in DESIGN.md (and restated independently in oracle/dirt_oracle.c): every operation that can
// change which face a pixel shows is an explicitly rounded intrinsic (__fmul_rn, __dadd_rn, ...)
// so that nvcc's fused-multiply-add contraction cannot alter it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dirt {

constexpr int TILE_W = 16;                 // forward raster / binning tile: one warp per 16x8 tile, a 2x2 quad per lane
constexpr int TILE_H = 8;
constexpr int TILE_W_SHIFT = 4;
constexpr int TILE_H_SHIFT = 3;
constexpr uint32_t KEY_EMPTY = 0x00800000u; // depth key of the cleared depth buffer (1.0)
constexpr int SMALL_TILE_LIMIT = 16;       // faces whose bbox spans <= this many tiles are binned per tile;
                                           // larger ones go to the per-image "large" list
constexpr int BIN_CAP = 128;               // face references a tile's bin holds; further ones go to the overflow list of the tile's row
constexpr int OVF_ROW_CAP = 1024;          // capacity of the overflow list of one row of tiles of one image
constexpr int MAX_GROUPS = 128;            // channel groups (each 1 or 3 wide)

constexpr uint32_t KIND_CULLED = 0, KIND_SMALL = 1, KIND_HARD = 2, KIND_LARGE = 3;

// ---- per-face records written by the setup kernel (64 B each, 16-B vector loadable) ------------
// Coverage + depth: everything the z-buffer loop needs.  Edge k: n_k(col,row) = A_k*col + B_k*row + q_k >= 0
// inside (S5); depth plane (zA,zB,zC) in absolute (col,row) (S6/S7).  Two layouts share the 64 bytes:
//   KIND_SMALL (binned per tile): q relative to the bbox corner (cref,rref) so that everything the raster loop
//              does fits int32;
//   KIND_LARGE / KIND_HARD: absolute int64 q (hard faces only use the depth plane and the vertex ids).
struct __align__(16) TriCov {
    int32_t A0, B0, A1, B1;
    int32_t A2, B2;
    union {
        struct { int32_t q0r, q1r, q2r; float zA, zB, zC; int32_t cref, rref, pad; uint32_t kind; } s;
        struct { float zA, zB; int64_t q0, q1, q2; float zC; uint32_t kind; } l;   // `kind` sits at byte 60 in both
    };
};
#define DIRT_COV_KIND(c) ((c).s.kind)
static_assert(sizeof(TriCov) == 64, "TriCov must be 64 bytes");

struct __align__(16) TriInterp { // interpolation planes relative to (cref,rref) + vertex ids (G)
    float q0A, q0B, q0C, q1A;
    float q1B, q1C, sA, sB;
    float sC;
    int32_t v0, v1, v2;
    int32_t cref, rref;
    int32_t pad0, pad1;
};
static_assert(sizeof(TriInterp) == 64, "TriInterp must be 64 bytes");

// Forward shading of a face whose tensors have C <= 4 channels: value_c(p) = N_c(p) / S(p), where N_c is the plane of
// sum_k (beta_k / w_k) * colour_kc and S the plane of 1 / clip_w, both relative to the face's reference pixel.  One
// 64-byte gather per (pixel, face) instead of the interpolation record plus three vertex-colour rows.
struct __align__(16) TriShade {
    float sA, sB, sC;
    uint32_t ref;        // cref | rref << 16 (frames up to 65536 x 65536)
    float n[4][3];       // (A, B, C) of N_c, c = 0..3
};
static_assert(sizeof(TriShade) == 64, "TriShade must be 64 bytes");
constexpr int SHADE_MAX_CHANNELS = 4;
constexpr int SHADE_MAX_EXTENT = 65536;

struct __align__(16) TriXY {     // clip-space x,y of the face's three vertices (the backward pass's clip_x / clip_y,
    float x0, y0, x1, y1;        // csrc/rasterise_grad_egl.cu:210-215): stored next to the planes so that a tile's face
    float x2, y2, pad0, pad1;    // table is filled in one hop
};
static_assert(sizeof(TriXY) == 32, "TriXY must be 32 bytes");

// ---- workspace ---------------------------------------------------------------------------------
struct Header {
    unsigned long long tag;   // workspace_tag() of the call that filled the setup records (0: none)
    int error;                // set by a backward call that was promised setup records which are not there
    int pad;
};

__host__ __device__ inline unsigned long long workspace_tag(const void* vertices, const void* faces, int B, int H, int W, int V, int F)
{
    // FNV-1a over the identity of the geometry tensors and the sizes; never 0
    unsigned long long h = 1469598103934665603ull;
    const unsigned long long words[7] = {(unsigned long long)(uintptr_t)vertices, (unsigned long long)(uintptr_t)faces,
                                         (unsigned long long)B, (unsigned long long)H, (unsigned long long)W,
                                         (unsigned long long)V, (unsigned long long)F};
    for (int i = 0; i < 7; ++i)
        for (int k = 0; k < 8; ++k) { h ^= (words[i] >> (8 * k)) & 0xffull; h *= 1099511628211ull; }
    return h ? h : 1ull;
}

struct Workspace {
    TriCov* cov;          // [B*F]
    TriInterp* itp;       // [B*F]
    TriXY* xy;            // [B*F]
    TriShade* shade;      // [B*F]  (written by forward calls with C <= SHADE_MAX_CHANNELS)
    // One-pass binning: the setup kernel appends a face to the bin of every tile its bounding box touches
    // (position = atomicAdd on the tile's count).  A bin holds BIN_CAP references at a fixed place, so the raster
    // kernel fetches a tile's count and its references in ONE hop; what does not fit goes to the overflow list of the
    // tile's ROW of tiles ((tile, face) pairs, read only by the tiles of that row whose bins were full), and what does
    // not fit there either to the image's large list (scanned by every tile).
    int* tile_count;      // [B*T]  references binned to the tile (may exceed BIN_CAP: the excess is in the overflow list)
    unsigned char* tile_flags;  // [B*T]  1 if the 16x8 tile or one of its 8 neighbours shows a face (written by the raster kernel)
    unsigned char* face_in_large;  // [B*F]  the face already sits in the large list (last-resort path only)
    int* large_count;     // [B]
    int* ovf_count;       // [B*tiles_y]
    int* large_list;      // [B*F]   faces spanning > SMALL_TILE_LIMIT tiles, hard faces
    int* bins;            // [B*T*BIN_CAP]
    int2* ovf;            // [B*tiles_y*OVF_ROW_CAP]  (tile, face)
    struct Header* header; // identity of the (vertices, faces, sizes) the setup records belong to; error flag
    int32_t* face_ids;    // [B*H*W] (last block; used only by a backward call without face ids)
    float* gc_pad;        // [B*V*4] (C == 3 only) grad_vertex_colors accumulated in 16-byte rows: one vector RED per vertex
    size_t zero_bytes;    // bytes from tile_count that the forward pass zeroes (counts, flags, list counts, header)
    size_t bytes_without_face_ids;
    size_t bytes;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline Workspace carve_workspace(void* base, int B, int H, int W, int C, int V, int F)
{
    Workspace ws;
    const size_t tiles = (size_t)((W + TILE_W - 1) / TILE_W) * ((H + TILE_H - 1) / TILE_H);
    const size_t BF = (size_t)B * F, BT = (size_t)B * tiles;
    size_t off = 0;
    char* p = (char*)base;
    auto take = [&](size_t bytes) { char* r = p + off; off = align_up(off + bytes, 256); return r; };
    ws.cov = (TriCov*)take(BF * sizeof(TriCov));
    ws.itp = (TriInterp*)take(BF * sizeof(TriInterp));
    ws.xy = (TriXY*)take(BF * sizeof(TriXY));
    ws.shade = (TriShade*)take(BF * sizeof(TriShade));
    // the blocks the forward pass must zero are adjacent: one memset covers [tile_count, zero_end)
    ws.tile_count = (int*)take(BT * sizeof(int));
    ws.tile_flags = (unsigned char*)take(BT);
    ws.face_in_large = (unsigned char*)take(BF);
    ws.large_count = (int*)take((size_t)B * sizeof(int));
    const size_t rows = (size_t)B * ((H + TILE_H - 1) / TILE_H);
    ws.ovf_count = (int*)take(rows * sizeof(int));
    ws.header = (Header*)take(256);
    ws.zero_bytes = (size_t)((p + off) - (char*)ws.tile_count);
    ws.large_list = (int*)take(BF * sizeof(int));
    ws.bins = (int*)take(BT * BIN_CAP * sizeof(int));
    ws.ovf = (int2*)take(rows * OVF_ROW_CAP * sizeof(int2));
    ws.gc_pad = (float*)take(C == 3 ? (size_t)B * V * 4 * sizeof(float) : 0);
    ws.bytes_without_face_ids = off;   // all a call needs whose caller holds the face ids (dirt_workspace_bytes_min)
    ws.face_ids = (int32_t*)take((size_t)B * H * W * sizeof(int32_t));
    ws.bytes = off;
    return ws;
}

// S6 constants of the NDC -> pixel-index map, divided once on the host (IEEE double, same values as on the device)
struct PixelScale {
    double two_over_W, two_over_H, inv_W, inv_H;
};

// ---- exactly specified arithmetic --------------------------------------------------------------
namespace exact {

__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }

// NDC plane (a,b,c) -> pixel-index plane g = (gA,gB,gC)
__device__ __forceinline__ void ndc_to_pixel_plane(double a, double b, double c, double two_over_W,
                                                   double two_over_H, double inv_W, double inv_H, double g[3])
{
    g[0] = __dmul_rn(a, two_over_W);
    g[1] = -__dmul_rn(b, two_over_H);
    double t0 = __dmul_rn(a, dsub(inv_W, 1.0));
    double t1 = __dmul_rn(b, dsub(1.0, inv_H));
    g[2] = __dadd_rn(__dadd_rn(t0, t1), c);
}

// S6: planes of q_k = beta_k/w_k (k=0..2), S = 1/clip_w and window depth, in double, absolute pixel
// indices.  p[k] = (x,y,z,w) of vertex k.  Returns false when the face is degenerate.
__device__ inline bool planes_double(const float p[3][4], const PixelScale& ps, double gq[3][3], double gs[3], double gz[3])
{
    const double x0 = p[0][0], y0 = p[0][1], w0 = p[0][3];
    const double x1 = p[1][0], y1 = p[1][1], w1 = p[1][3];
    const double x2 = p[2][0], y2 = p[2][1], w2 = p[2][3];
    const double c00 = dsub(__dmul_rn(y1, w2), __dmul_rn(y2, w1));
    const double c01 = dsub(__dmul_rn(y2, w0), __dmul_rn(y0, w2));
    const double c02 = dsub(__dmul_rn(y0, w1), __dmul_rn(y1, w0));
    const double c10 = dsub(__dmul_rn(w1, x2), __dmul_rn(w2, x1));
    const double c11 = dsub(__dmul_rn(w2, x0), __dmul_rn(w0, x2));
    const double c12 = dsub(__dmul_rn(w0, x1), __dmul_rn(w1, x0));
    const double c20 = dsub(__dmul_rn(x1, y2), __dmul_rn(x2, y1));
    const double c21 = dsub(__dmul_rn(x2, y0), __dmul_rn(x0, y2));
    const double c22 = dsub(__dmul_rn(x0, y1), __dmul_rn(x1, y0));
    const double det = __dadd_rn(__dadd_rn(__dmul_rn(x0, c00), __dmul_rn(y0, c10)), __dmul_rn(w0, c20));
    if (!(det != 0.0) || !isfinite(det)) return false;
    const double rdet = __ddiv_rn(1.0, det);   // one division, then products (as the specification says)
    double inv[3][3];
    inv[0][0] = __dmul_rn(c00, rdet); inv[0][1] = __dmul_rn(c01, rdet); inv[0][2] = __dmul_rn(c02, rdet);
    inv[1][0] = __dmul_rn(c10, rdet); inv[1][1] = __dmul_rn(c11, rdet); inv[1][2] = __dmul_rn(c12, rdet);
    inv[2][0] = __dmul_rn(c20, rdet); inv[2][1] = __dmul_rn(c21, rdet); inv[2][2] = __dmul_rn(c22, rdet);
    const double two_over_W = ps.two_over_W, two_over_H = ps.two_over_H, inv_W = ps.inv_W, inv_H = ps.inv_H;
#pragma unroll
    for (int k = 0; k < 3; ++k)
        ndc_to_pixel_plane(inv[0][k], inv[1][k], inv[2][k], two_over_W, two_over_H, inv_W, inv_H, gq[k]);
    ndc_to_pixel_plane(__dadd_rn(__dadd_rn(inv[0][0], inv[0][1]), inv[0][2]),
                       __dadd_rn(__dadd_rn(inv[1][0], inv[1][1]), inv[1][2]),
                       __dadd_rn(__dadd_rn(inv[2][0], inv[2][1]), inv[2][2]),
                       two_over_W, two_over_H, inv_W, inv_H, gs);
    const double z0 = p[0][2], z1 = p[1][2], z2 = p[2][2];
    const double a = __dadd_rn(__dadd_rn(__dmul_rn(inv[0][0], z0), __dmul_rn(inv[0][1], z1)), __dmul_rn(inv[0][2], z2));
    const double b = __dadd_rn(__dadd_rn(__dmul_rn(inv[1][0], z0), __dmul_rn(inv[1][1], z1)), __dmul_rn(inv[1][2], z2));
    const double c = __dadd_rn(__dadd_rn(__dmul_rn(inv[2][0], z0), __dmul_rn(inv[2][1], z1)), __dmul_rn(inv[2][2], z2));
    double g[3];
    ndc_to_pixel_plane(a, b, c, two_over_W, two_over_H, inv_W, inv_H, g);
    gz[0] = __dmul_rn(0.5, g[0]);
    gz[1] = __dmul_rn(0.5, g[1]);
    gz[2] = __dadd_rn(__dmul_rn(0.5, g[2]), 0.5);
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if (!isfinite(gz[j]) || !isfinite(gs[j]) || !isfinite(gq[0][j]) || !isfinite(gq[1][j]) || !isfinite(gq[2][j]))
            return false;
    return true;
}

// S7: depth key from a window depth z; a fragment exists iff key < KEY_EMPTY
__device__ __forceinline__ uint32_t depth_key(float z)
{
    return __float_as_uint(__fmaf_rn(z, 8388608.0f, 8388608.0f)) - 0x4B000000u;
}

// S7: depth of a normal face at absolute pixel (col,row)
__device__ __forceinline__ float depth_normal(float zA, float zB, float zC, float col, float row)
{
    return __fmaf_rn(zA, col, __fmaf_rn(zB, row, zC));
}

// H1: a double plane at absolute (col,row): gA*col + (gB*row + gC)
__device__ __forceinline__ double plane_double(const double g[3], int col, int row)
{
    return __dadd_rn(__dmul_rn(g[0], (double)col), __dadd_rn(__dmul_rn(g[1], (double)row), g[2]));
}

// Correctly rounded 1/x without the subroutine call __frcp_rn compiles to: MUFU.RCP (<= 1 ulp) followed by one
// fused Newton step is correctly rounded for every x whose reciprocal is a normal number; the rare remaining
// magnitudes take the library path.  (tests/test_gpu_parity.py::test_visibility_gbuffer_matches_oracle compares the
// result bit for bit with the oracle's IEEE 1.0f / x.)
__device__ __forceinline__ float rcp_rn(float x)
{
    const float ax = fabsf(x);
    if (ax > 1.0e-30f && ax < 1.0e30f) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        const float e = __fmaf_rn(-x, r, 1.0f);
        return __fmaf_rn(r, e, r);
    }
    return __frcp_rn(x);
}

// G: barycentrics and clip-w of a face at pixel (col,row)
__device__ __forceinline__ float4 gbuffer_at(const TriInterp& t, int col, int row)
{
    const float dc = (float)(col - t.cref), dr = (float)(row - t.rref);
    const float S = __fmaf_rn(t.sA, dc, __fmaf_rn(t.sB, dr, t.sC));
    const float cw = rcp_rn(S);   // correctly rounded reciprocal == the oracle's IEEE 1.0f / S
    const float q0 = __fmaf_rn(t.q0A, dc, __fmaf_rn(t.q0B, dr, t.q0C));
    const float q1 = __fmaf_rn(t.q1A, dc, __fmaf_rn(t.q1B, dr, t.q1C));
    const float b0 = __fmul_rn(q0, cw), b1 = __fmul_rn(q1, cw);
    return make_float4(b0, b1, __fsub_rn(__fsub_rn(1.0f, b0), b1), cw);
}

}  // namespace exact

// 64-byte record loads through the read-only path
__device__ __forceinline__ TriInterp load_interp(const TriInterp* p)
{
    union { TriInterp t; uint4 u[4]; } r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    r.u[0] = __ldg(s); r.u[1] = __ldg(s + 1); r.u[2] = __ldg(s + 2); r.u[3] = __ldg(s + 3);
    return r.t;
}
__device__ __forceinline__ TriCov load_cov(const TriCov* p)
{
    union { TriCov t; uint4 u[4]; } r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    r.u[0] = __ldg(s); r.u[1] = __ldg(s + 1); r.u[2] = __ldg(s + 2); r.u[3] = __ldg(s + 3);
    return r.t;
}

// ---- launch parameter blocks -------------------------------------------------------------------
struct Dims {
    int B, H, W, C, V, F;
    PixelScale ps;
    int tiles_x, tiles_y, tiles;     // per image, forward/binning tiles (TILE_W x TILE_H)
    int btiles_x, btiles_y, btiles;  // per image, backward tiles (8 x 8)
};

inline Dims make_dims(int B, int H, int W, int C, int V, int F)
{
    Dims d;
    d.B = B; d.H = H; d.W = W; d.C = C; d.V = V; d.F = F;
    d.tiles_x = (W + TILE_W - 1) / TILE_W;
    d.tiles_y = (H + TILE_H - 1) / TILE_H;
    d.tiles = d.tiles_x * d.tiles_y;
    d.btiles_x = (W + 7) / 8;
    d.btiles_y = (H + 7) / 8;
    d.btiles = d.btiles_x * d.btiles_y;
    d.ps.two_over_W = 2.0 / (double)W; d.ps.two_over_H = 2.0 / (double)H;
    d.ps.inv_W = 1.0 / (double)W; d.ps.inv_H = 1.0 / (double)H;
    return d;
}

struct GroupSpec {
    int n;
    unsigned char width[MAX_GROUPS];
};

// ---- optional per-kernel timing (dirt_kernel_timer_enable) -------------------------------------
struct KernelTimer {
    int which = 0;  // 0 off, 1 forward raster kernel, 2 backward kernel
    cudaEvent_t start = nullptr, stop = nullptr;
    bool recorded = false;
};
KernelTimer& kernel_timer();  // thread-local, defined in api.cu

struct ScopedKernelTimer {
    KernelTimer& t;
    cudaStream_t stream;
    bool on;
    ScopedKernelTimer(int which, cudaStream_t s) : t(kernel_timer()), stream(s), on(t.which == which && t.start)
    {
        if (on) cudaEventRecord(t.start, stream);
    }
    ~ScopedKernelTimer()
    {
        if (on) { cudaEventRecord(t.stop, stream); t.recorded = true; }
    }
};

// ---- host-side launchers (one per .cu) ----------------------------------------------------------
// All return cudaError_t of the launch and add the number of kernels they launched to *launches.
// Both setup launchers leave workspace_tag(vertices, faces, sizes) in the workspace header.
// vertex_colors != nullptr (and shade_records_ok(d)): the shading records are written as well.
inline bool shade_records_ok(const Dims& d) { return d.C <= SHADE_MAX_CHANNELS && d.W <= SHADE_MAX_EXTENT && d.H <= SHADE_MAX_EXTENT; }
cudaError_t launch_setup_and_bin(const float* vertices, const int32_t* faces, const float* vertex_colors, const Workspace& ws,
                                 const Dims& d, cudaStream_t stream, int* launches);
cudaError_t launch_setup_only(const float* vertices, const int32_t* faces, const Workspace& ws, const Dims& d,
                              cudaStream_t stream, int* launches);
cudaError_t launch_raster_forward(const float* vertices, const float* background, const float* vertex_colors, float* pixels,
                                  int32_t* face_ids_out, const Workspace& ws, const Dims& d, cudaStream_t stream,
                                  int* launches);
cudaError_t launch_raster_visibility(const float* vertices, int32_t* face_ids, float* gbuffer, const Workspace& ws, const Dims& d,
                                     cudaStream_t stream, int* launches);
cudaError_t launch_backward(const float* vertices, const float* pixels, const float* grad_pixels,
                            const int32_t* face_ids, float* grad_background, float* grad_vertices,
                            float* grad_vertex_colors, const Workspace& ws, const Dims& d, const GroupSpec& groups,
                            bool tile_flags_valid, int flags, unsigned long long expect_tag, cudaStream_t stream,
                            int* launches);   // flags: DIRT_BWD_* of include/dirt_b200.h; expect_tag != 0: the records are
                                              // promised to carry this tag (checked on the device)
constexpr int BWD_SHARED_GEOMETRY = 1, BWD_SKIP_POSITION = 2, BWD_SKIP_COLOUR = 4;   // == DIRT_BWD_* (static_assert in api.cu)

}  // namespace dirt
