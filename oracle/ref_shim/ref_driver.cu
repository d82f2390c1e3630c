// ref_driver.cu -- runs the reference's OWN gradient kernel on a caller-supplied G-buffer.
//
// TEST INFRASTRUCTURE (oracle/_ref).  The Makefile compiles /root/reference/csrc/rasterise_grad_egl.cu UNMODIFIED
// (against the stand-in headers in this directory) together with this file into oracle/_ref/libdirt_ref_grad.so.
// This file only does what RasteriseGradOpGpu::Compute does around launch_grad_assembly
// (csrc/rasterise_grad_egl.cpp:381-472): allocate the four outputs, lay the per-pixel (barycentrics, clip_w) and
// vertex-index images out as the tiled, y-up RGBA32F framebuffer textures GL would have rendered
// (frame placement :405-411,436-437; clear values :442-445; indices stored as floats, csrc/shaders.cpp:71), wrap
// them in CUDA arrays and call the reference's launch_grad_assembly.  All arithmetic that produces a gradient is
// the reference's.
//
// The G-buffer is an INPUT: in the reference it comes from the OpenGL driver, here from whichever
// visibility implementation is being pinned (the CPU oracle's, in tests/ and tests/golden/make_ref_golden.py).
#include "tensorflow/core/framework/tensor.h"
#include "tensorflow/core/util/cuda_launch_config.h"

#include "rasterise_grad_common.h"   // the reference's header (csrc/), found through -I /root/reference/csrc

#include <cstdint>
#include <vector>

namespace {

struct DeviceBuffer {
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 4); }
    ~DeviceBuffer() { if (p) cudaFree(p); }
};

struct DeviceArray {
    cudaArray_t a = nullptr;
    ~DeviceArray() { if (a) cudaFreeArray(a); }
};

}  // namespace

#define TRY(expr)                                        \
    do {                                                 \
        cudaError_t e__ = (expr);                        \
        if (e__ != cudaSuccess) {                        \
            std::fprintf(stderr, "ref_driver: %s failed: %s\n", #expr, cudaGetErrorString(e__)); \
            return -(int)e__;                            \
        }                                                \
    } while (0)

// All pointers are HOST pointers.  Layouts are the op's (row-major, channels last, row 0 = top of the image):
//   gbuffer [B,H,W,4] = (bary0, bary1, bary2, clip_w), (-1,-1,-1,+inf) where uncovered
//   vertex_ids [B,H,W,3] = the visible face's three vertex indices, -1 where uncovered
//   pixels, grad_pixels, grad_background [B,H,W,C] with C = 1 or 3 (one native-op call, csrc/hwc.h:27)
//   vertices, grad_vertices [B,V,4]; grad_vertex_colors [B,V,C]; debug_thingy [B,H,W,3] (may be NULL)
// Returns 0 on success, a negated cudaError_t otherwise.
extern "C" int dirt_ref_assemble_grads(const float* gbuffer, const int32_t* vertex_ids, const float* pixels,
                                       const float* grad_pixels, const float* vertices, float* grad_vertices,
                                       float* grad_vertex_colors, float* grad_background, float* debug_thingy, int B,
                                       int H, int W, int C, int V)
{
    if (B <= 0 || H <= 0 || W <= 0 || (C != 1 && C != 3) || V <= 0) return -1000;
    // framebuffer tiling as in csrc/rasterise_grad_egl.cpp:405-411
    const int horizontal_count = static_cast<int>(std::sqrt(static_cast<float>(B)) + .1f);
    const int vertical_count = B / horizontal_count + (B % horizontal_count == 0 ? 0 : 1);
    const int buffer_width = W * horizontal_count, buffer_height = H * vertical_count;
    const int frames_per_row = buffer_width / W;

    const float inf = std::numeric_limits<float>::infinity();
    std::vector<float4> bary_depth((size_t)buffer_width * buffer_height, make_float4(-1.f, -1.f, -1.f, inf));
    std::vector<float4> indices((size_t)buffer_width * buffer_height, make_float4(-1.f, -1.f, -1.f, -1.f));
    for (int b = 0; b < B; ++b) {
        const int frame_x = (b % frames_per_row) * W, frame_y = (b / frames_per_row) * H;
        for (int r = 0; r < H; ++r)
            for (int c = 0; c < W; ++c) {
                const size_t src = ((size_t)b * H + r) * W + c;
                const size_t dst = (size_t)(frame_y + (H - 1 - r)) * buffer_width + frame_x + c;   // GL rows run bottom-up
                bary_depth[dst] = make_float4(gbuffer[src * 4], gbuffer[src * 4 + 1], gbuffer[src * 4 + 2], gbuffer[src * 4 + 3]);
                // the fragment shader writes a vec3 into an RGBA32F attachment; alpha is never read by assemble_grads
                indices[dst] = make_float4((float)vertex_ids[src * 3], (float)vertex_ids[src * 3 + 1], (float)vertex_ids[src * 3 + 2], 1.f);
            }
    }

    const cudaChannelFormatDesc desc = cudaCreateChannelDesc<float4>();
    DeviceArray bary_array, index_array;
    TRY(cudaMallocArray(&bary_array.a, &desc, buffer_width, buffer_height, cudaArraySurfaceLoadStore));
    TRY(cudaMallocArray(&index_array.a, &desc, buffer_width, buffer_height, cudaArraySurfaceLoadStore));
    const size_t pitch = (size_t)buffer_width * sizeof(float4);
    TRY(cudaMemcpy2DToArray(bary_array.a, 0, 0, bary_depth.data(), pitch, pitch, buffer_height, cudaMemcpyHostToDevice));
    TRY(cudaMemcpy2DToArray(index_array.a, 0, 0, indices.data(), pitch, pitch, buffer_height, cudaMemcpyHostToDevice));

    const size_t n_pix = (size_t)B * H * W;
    DeviceBuffer d_pixels, d_grad_pixels, d_vertices, d_gv, d_gc, d_gb, d_dbg;
    // `pixels` gets 8 zero floats of tail padding: for C = 1 the reference's at() reads "channels" 1 and 2 of the
    // last pixels of the last image past the end of the tensor (csrc/rasterise_grad_egl.cu:119-123); those reads are
    // defined as 0 here, as in the oracle
    TRY(d_pixels.alloc((n_pix * C + 8) * sizeof(float)));
    TRY(cudaMemset(d_pixels.p, 0, (n_pix * C + 8) * sizeof(float)));
    TRY(cudaMemcpy(d_pixels.p, pixels, n_pix * C * sizeof(float), cudaMemcpyHostToDevice));
    TRY(d_grad_pixels.alloc(n_pix * C * sizeof(float)));
    TRY(cudaMemcpy(d_grad_pixels.p, grad_pixels, n_pix * C * sizeof(float), cudaMemcpyHostToDevice));
    TRY(d_vertices.alloc((size_t)B * V * 4 * sizeof(float)));
    TRY(cudaMemcpy(d_vertices.p, vertices, (size_t)B * V * 4 * sizeof(float), cudaMemcpyHostToDevice));
    TRY(d_gv.alloc((size_t)B * V * 4 * sizeof(float)));
    TRY(d_gc.alloc((size_t)B * V * C * sizeof(float)));
    TRY(d_gb.alloc(n_pix * C * sizeof(float)));
    TRY(d_dbg.alloc(n_pix * 3 * sizeof(float)));

    tensorflow::Tensor t_gv(d_gv.p, {B, V, 4}), t_gc(d_gc.p, {B, V, C}), t_gb(d_gb.p, {B, H, W, C}), t_dbg(d_dbg.p, {B, H, W, 3});
    const tensorflow::Tensor t_pixels(d_pixels.p, {B, H, W, C}), t_grad_pixels(d_grad_pixels.p, {B, H, W, C}), t_vertices(d_vertices.p, {B, V, 4});
    Eigen::GpuDevice device;
    const cudaArray_t bary_handle = bary_array.a, index_handle = index_array.a;
    // the reference's launcher: four memsets, two surface objects, assemble_grads (csrc/rasterise_grad_egl.cu:238-278)
    launch_grad_assembly(t_gv, t_gc, t_gb, t_dbg, bary_handle, index_handle, t_pixels, t_grad_pixels, t_vertices, buffer_width,
                         buffer_height, device);
    TRY(cudaGetLastError());
    TRY(cudaDeviceSynchronize());

    TRY(cudaMemcpy(grad_vertices, d_gv.p, (size_t)B * V * 4 * sizeof(float), cudaMemcpyDeviceToHost));
    TRY(cudaMemcpy(grad_vertex_colors, d_gc.p, (size_t)B * V * C * sizeof(float), cudaMemcpyDeviceToHost));
    TRY(cudaMemcpy(grad_background, d_gb.p, n_pix * C * sizeof(float), cudaMemcpyDeviceToHost));
    if (debug_thingy) TRY(cudaMemcpy(debug_thingy, d_dbg.p, n_pix * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

// The reference's vertex expansion kernel upload_vertices (csrc/rasterise_grad_egl.cu:12-34) on host arrays:
// expanded [B, 3F] Vertex records (36 bytes each, csrc/rasterise_grad_common.h:5-11).
extern "C" int dirt_ref_upload_vertices(const float* vertices, const int32_t* faces, void* expanded, int B, int V, int F)
{
    if (B <= 0 || V <= 0 || F <= 0) return -1000;
    DeviceBuffer d_vertices, d_faces, d_out;
    TRY(d_vertices.alloc((size_t)B * V * 4 * sizeof(float)));
    TRY(cudaMemcpy(d_vertices.p, vertices, (size_t)B * V * 4 * sizeof(float), cudaMemcpyHostToDevice));
    TRY(d_faces.alloc((size_t)B * F * 3 * sizeof(int32_t)));
    TRY(cudaMemcpy(d_faces.p, faces, (size_t)B * F * 3 * sizeof(int32_t), cudaMemcpyHostToDevice));
    TRY(d_out.alloc((size_t)B * F * 3 * sizeof(Vertex)));
    tensorflow::TTypes<Vertex, 2>::Tensor buffer;
    buffer.data_ = static_cast<Vertex*>(d_out.p);
    buffer.dims_[0] = B; buffer.dims_[1] = (long)F * 3;
    const tensorflow::Tensor t_vertices(d_vertices.p, {B, V, 4}), t_faces(d_faces.p, {B, F, 3});
    Eigen::GpuDevice device;
    launch_vertex_upload(buffer, t_vertices, t_faces, device);
    TRY(cudaGetLastError());
    TRY(cudaDeviceSynchronize());
    TRY(cudaMemcpy(expanded, d_out.p, (size_t)B * F * 3 * sizeof(Vertex), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int dirt_ref_sizeof_vertex(void) { return (int)sizeof(Vertex); }
