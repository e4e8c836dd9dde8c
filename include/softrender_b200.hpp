// softrender_b200.hpp -- header-only C++17 mirror of the reference's typed builder API over the C ABI.
//
// The reference is Rust; no Rust toolchain exists in this image, so the host side above the C ABI is
// written in C++ with the same verbs, argument meaning, ownership (stage transitions consume the stage,
// one draw in flight per pipeline) and error behaviour (reference panics become softrender::Error).
//
//   Pipeline::from_framebuffer(fb, uniforms)             src/pipeline/mod.rs:120
//     .render_mesh(Triangle{}, mesh, stencil)            src/pipeline/mod.rs:146            -> VertexShader
//        .run(vs)                                        src/pipeline/stages/vertex.rs:87   -> GeometryShader
//           .run(gs) / .clip_primitives()                src/pipeline/stages/geometry.rs:132,261
//           .finish(viewport)                            src/pipeline/stages/geometry.rs:60 -> FragmentShader
//        .run_to_fragment(viewport, vs)                  src/pipeline/stages/vertex.rs:123  -> FragmentShader
//              .with_blend / cull_faces / tile_size ..   src/pipeline/stages/fragment.rs:82-160
//              .run(fs)                                  src/pipeline/stages/fragment.rs:168
#pragma once

#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "softrender_b200.h"

namespace softrender {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string &m) : std::runtime_error(m), status(s) {}
};
inline void check(int status) {
    if (status != SR_OK) throw Error(status, sr_last_error());
}

// zero-sized primitive markers (src/primitive.rs:76,102,133)
struct Point { static constexpr uint32_t id = SR_POINT; };
struct Line { static constexpr uint32_t id = SR_LINE; };
struct Triangle { static constexpr uint32_t id = SR_TRIANGLE; };

struct Dimensions { uint32_t width, height; };  // src/geometry/dimension.rs:4

inline sr_viewport Viewport(Dimensions d, float near_, float far_, uint32_t x = 0, uint32_t y = 0) {
    return sr_viewport{(float)x, (float)y, (float)d.width, (float)d.height, near_, far_};  // Viewport::new (clipvertex.rs:50-59)
}

class Context {
public:
    explicit Context(int device = 0) { check(sr_context_create(device, &h_)); }
    ~Context() { sr_context_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    sr_context *handle() const { return h_; }
    void synchronize() { check(sr_context_synchronize(h_)); }
private:
    sr_context *h_ = nullptr;
};

// one pixel of RenderBuffer<ColorDepthAttachments<RGBAf32Color, f32>> (20 bytes)
struct PixelCD { float r, g, b, a, depth; };

class RenderBuffer {
public:
    static RenderBuffer with_dimensions(Context &c, Dimensions d, bool stencil = false) {
        RenderBuffer fb;
        check(sr_framebuffer_create(c.handle(), d.width, d.height, stencil ? SR_FB_RGBAF32_DF32_S8 : SR_FB_RGBAF32_DF32, &fb.h_));
        fb.dim_ = d;
        return fb;
    }
    RenderBuffer(RenderBuffer &&o) noexcept : h_(o.h_), dim_(o.dim_) { o.h_ = nullptr; }
    ~RenderBuffer() { if (h_) sr_framebuffer_destroy(h_); }
    Dimensions dimensions() const { return dim_; }
    void clear(const float (&color)[4]) { check(sr_framebuffer_clear(h_, color)); }
    /* one colour of Framebuffer::clear's tuple for a texture buffer with several colour planes (texturebuffer.rs:181-197) */
    void clear_attachment(uint32_t index, const float (&color)[4]) { check(sr_framebuffer_clear_attachment(h_, index, color)); }
    void download_attachment(uint32_t index, float *color) { check(sr_framebuffer_download_attachment(h_, index, color)); }
    std::vector<PixelCD> pixels() {
        std::vector<PixelCD> out((size_t)dim_.width * dim_.height);
        check(sr_framebuffer_download(h_, out.data(), out.size() * sizeof(PixelCD)));
        return out;
    }
    // presentation read-back (realtime_example/src/main.rs:100-116): `(c * 255.0) as u8` per channel, converted on the device;
    // abgr = the byte order the example writes into SDL's RGBA8888 streaming texture
    std::vector<uint8_t> rgba8(bool abgr = false) {
        std::vector<uint8_t> out((size_t)dim_.width * dim_.height * 4);
        check(sr_framebuffer_download_rgba8(h_, out.data(), out.size(), abgr ? 1u : 0u));
        return out;
    }
    PixelCD pixel(uint32_t x, uint32_t y) {  // checked accessor: throws Error{SR_ERR_INVALID_PIXEL_COORDINATE}
        PixelCD p;
        check(sr_framebuffer_get_pixel(h_, x, y, &p.r, &p.depth, nullptr));
        return p;
    }
    sr_framebuffer *handle() const { return h_; }
private:
    RenderBuffer() = default;
    sr_framebuffer *h_ = nullptr;
    Dimensions dim_{0, 0};
};

class Mesh {  // Arc<Mesh<V>>
public:
    Mesh(Context &c, const float *vertices, uint64_t nverts, uint32_t vin_floats, const uint32_t *indices, uint64_t nindices) {
        check(sr_mesh_create(c.handle(), vertices, nverts, vin_floats, indices, nindices, 4, &h_));
    }
    ~Mesh() { sr_mesh_destroy(h_); }
    Mesh(const Mesh &) = delete;
    sr_mesh *handle() const { return h_; }
private:
    sr_mesh *h_ = nullptr;
};

class Stage {
protected:
    explicit Stage(sr_draw *h) : h_(h) {}
    Stage(Stage &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    ~Stage() { if (h_) sr_draw_destroy(h_); }
    sr_draw *take() { sr_draw *h = h_; h_ = nullptr; return h; }
    sr_draw *h_;
};

class FragmentShader : public Stage {
public:
    explicit FragmentShader(sr_draw *h) : Stage(h) {}
    FragmentShader(FragmentShader &&) = default;
    FragmentShader duplicate() { sr_draw *d; check(sr_draw_duplicate(h_, &d)); return FragmentShader(d); }
    FragmentShader &&cull_faces(std::optional<sr_winding> w) && { check(sr_fragment_set_cull_faces(h_, w ? *w : SR_CULL_NONE)); return std::move(*this); }
    FragmentShader &&antialiased_lines(bool e) && { check(sr_fragment_set_antialiased_lines(h_, e)); return std::move(*this); }
    FragmentShader &&tile_size(Dimensions d) && { check(sr_fragment_set_tile_size(h_, d.width, d.height)); return std::move(*this); }
    FragmentShader &&with_blend(sr_blend b) && { check(sr_fragment_set_blend(h_, b)); return std::move(*this); }
    void run(sr_fragment_shader fs) && { check(sr_fragment_run(h_, fs)); }
};

class GeometryShader : public Stage {
public:
    explicit GeometryShader(sr_draw *h) : Stage(h) {}
    GeometryShader(GeometryShader &&) = default;
    GeometryShader duplicate() { sr_draw *d; check(sr_draw_duplicate(h_, &d)); return GeometryShader(d); }
    GeometryShader run(sr_geometry_shader gs) && { check(sr_geometry_run(h_, gs)); return GeometryShader(take()); }
    GeometryShader clip_primitives() && { check(sr_geometry_clip_primitives(h_)); return GeometryShader(take()); }
    FragmentShader finish(const sr_viewport &vp) && { check(sr_geometry_finish(h_, &vp)); return FragmentShader(take()); }
};

class VertexShader : public Stage {
public:
    explicit VertexShader(sr_draw *h) : Stage(h) {}
    VertexShader(VertexShader &&) = default;
    GeometryShader run(sr_vertex_shader vs) && { check(sr_vertex_run(h_, vs)); return GeometryShader(take()); }
    FragmentShader run_to_fragment(const sr_viewport &vp, sr_vertex_shader vs) && {
        check(sr_vertex_run_to_fragment(h_, &vp, vs));
        return FragmentShader(take());
    }
};

class Pipeline {
public:
    static Pipeline from_framebuffer(Context &c, RenderBuffer &fb, const sr_uniforms &u) {
        Pipeline p;
        check(sr_pipeline_create(c.handle(), fb.handle(), &u, &p.h_));
        p.fb_ = &fb;
        return p;
    }
    Pipeline(Pipeline &&o) noexcept : h_(o.h_), fb_(o.fb_) { o.h_ = nullptr; }
    ~Pipeline() { if (h_) sr_pipeline_destroy(h_); }
    RenderBuffer &framebuffer() { return *fb_; }
    void set_uniforms(const sr_uniforms &u) { check(sr_pipeline_set_uniforms(h_, &u)); }  // *uniforms_mut() = u
    void set_stencil_config(sr_stencil_test t, sr_stencil_op o) { check(sr_pipeline_set_stencil_config(h_, t, o)); }
    /* render-to-texture: the colour of `src` sampled in place (TextureBufferRef, src/framebuffer/texturebuffer.rs:12-58) */
    void bind_framebuffer_texture(RenderBuffer *src) { check(sr_pipeline_bind_framebuffer_texture(h_, src ? src->handle() : nullptr)); }
    /* colour plane `index` of a texture buffer declared with several (the named accessors of declare_texture_buffer!, :110-117) */
    void bind_framebuffer_attachment(RenderBuffer &src, uint32_t index) { check(sr_pipeline_bind_framebuffer_attachment(h_, src.handle(), index)); }
    /* Filter / Edge of texture(t, coord, filter, edge), src/texture.rs:14-45 */
    void set_sampler(sr_texture_filter f, sr_texture_edge e, const float *border_rgba = nullptr) { check(sr_pipeline_set_sampler(h_, f, e, border_rgba)); }
    template <class T>
    VertexShader render_mesh(T, const Mesh &mesh, std::optional<uint32_t> stencil = std::nullopt) {
        sr_draw *d;
        check(sr_render_mesh(h_, mesh.handle(), T::id, stencil ? 1 : 0, stencil.value_or(0), &d));
        return VertexShader(d);
    }
private:
    Pipeline() = default;
    sr_pipeline *h_ = nullptr;
    RenderBuffer *fb_ = nullptr;
};

}  // namespace softrender
