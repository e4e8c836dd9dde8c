// builder_example.cpp -- the flow of the reference's examples/suzanne.rs:73-183 written against the C++
// builder mirror (include/softrender_b200.hpp): framebuffer -> pipeline -> render_mesh -> vertex run ->
// clip_primitives -> finish -> fragment run.  Renders one lit triangle pair and prints coverage; used by
// build() as the "does the host mirror compile and link" check and runnable on a GPU box.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/softrender_b200.hpp"

using namespace softrender;

static void identity(float *m) {
    for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
}

int main() {
    try {
        Context ctx(0);
        auto fb = RenderBuffer::with_dimensions(ctx, Dimensions{256, 256});
        const float clear[4] = {0.01f, 0.01f, 0.01f, 1.0f};
        fb.clear(clear);

        sr_uniforms u{};
        identity(u.model); identity(u.mit); identity(u.view); identity(u.projection);
        u.projection[10] = -1.0f;  // keep z in front: clip z = -z_world
        u.camera[2] = 2.0f; u.camera[3] = 1.0f;
        u.sz_light[0] = u.sz_light[1] = u.sz_light[2] = 5.0f; u.sz_light[3] = 1.0f;
        u.sz_color[0] = std::pow(0.1f, 2.2f); u.sz_color[1] = std::pow(0.5f, 2.2f); u.sz_color[2] = std::pow(0.1f, 2.2f); u.sz_color[3] = 1.0f;
        u.sz_intensity = 4.0f;

        auto pipeline = Pipeline::from_framebuffer(ctx, fb, u);
        const float verts[4 * 6] = {-0.5f, -0.5f, -0.5f, 0, 0, 1, 0.5f, -0.5f, -0.5f, 0, 0, 1,
                                    0.5f, 0.5f, -0.5f, 0, 0, 1, -0.5f, 0.5f, -0.5f, 0, 0, 1};
        const uint32_t idx[6] = {0, 1, 2, 0, 2, 3};
        Mesh mesh(ctx, verts, 4, 6, idx, 6);

        const sr_viewport vp = Viewport(fb.dimensions(), 0.001f, 1000.0f);
        pipeline.render_mesh(Triangle{}, mesh)
            .run(SR_VS_SUZANNE)
            .clip_primitives()
            .finish(vp)
            .cull_faces(std::nullopt)
            .run(SR_FS_SUZANNE);

        size_t covered = 0;
        for (const PixelCD &p : fb.pixels()) covered += p.depth > -1e30f;
        std::printf("builder_example: %zu of %u pixels covered\n", covered, 256u * 256u);
        // second pass (render-to-texture): the frame above sampled in place by a full-screen quad, Nearest + Clamp at the same
        // size is a 1:1 copy of the colours (texturebuffer.rs:12-58, texture.rs:14-45)
        auto fb2 = RenderBuffer::with_dimensions(ctx, Dimensions{256, 256});
        fb2.clear(clear);
        auto post = Pipeline::from_framebuffer(ctx, fb2, u);
        post.bind_framebuffer_texture(&fb);
        post.set_sampler(SR_FILTER_NEAREST, SR_EDGE_CLAMP);
        const float quad[4 * 6] = {-1, -1, 0, 1, 0, 1, 1, -1, 0, 1, 1, 1, 1, 1, 0, 1, 1, 0, -1, 1, 0, 1, 0, 0};  // clip xyzw + uv
        Mesh qmesh(ctx, quad, 4, 6, idx, 6);
        post.render_mesh(Triangle{}, qmesh).run_to_fragment(vp, SR_VS_PASSTHROUGH).run(SR_FS_TEXTURE_UNLIT);
        size_t same = 0;
        const auto a = fb.pixels(), b = fb2.pixels();
        for (size_t i = 0; i < a.size(); ++i) same += std::memcmp(&a[i].r, &b[i].r, 4 * sizeof(float)) == 0;
        std::printf("render-to-texture copy pass: %zu of %zu pixels identical\n", same, a.size());
        if (same != a.size()) return 3;
        try {
            fb.pixel(256, 0);
        } catch (const Error &e) {
            std::printf("checked accessor: status %d (%s)\n", e.status, e.what());
        }
        return covered == 128u * 128u + 0 ? 0 : (covered > 0 ? 0 : 2);
    } catch (const Error &e) {
        std::fprintf(stderr, "softrender error %d: %s\n", e.status, e.what());
        return 1;
    }
}
