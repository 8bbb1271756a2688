// Test program for include/rt_dropin.h (needs a GPU): a world built through the reference's class surface is uploaded
// with rt_upload_world / rt_apply_camera and rendered through the C ABI; the frame goes to argv[1] as raw floats.
//   stdin: n, then n lines "cx cy cz r mat ax ay az param" (hex floats)
#include <cstdio>
#include <vector>

#include "rt_dropin.h"

int main(int argc, char **argv) {
    if (argc < 2) return 1;
    int n = 0;
    if (scanf("%d", &n) != 1) return 1;
    std::vector<hitable *> list;
    for (int i = 0; i < n; i++) {
        float cx, cy, cz, r, ax, ay, az, p;
        int mat;
        if (scanf("%a %a %a %a %d %a %a %a %a", &cx, &cy, &cz, &r, &mat, &ax, &ay, &az, &p) != 9) return 2;
        material *m = nullptr;
        if (mat == RT_MAT_LAMBERTIAN) m = new lambertian(vec3(ax, ay, az));
        else if (mat == RT_MAT_METAL) m = new metal(vec3(ax, ay, az), p);
        else if (mat == RT_MAT_DIELECTRIC) m = new dielectric(p);
        list.push_back(new sphere(vec3(cx, cy, cz), r, m));
    }
    hitable *world = new hitable_list(list.data(), n);
    const int nx = 64, ny = 48, ns = 2;
    camera cam(vec3(13, 2, 3), vec3(0, 0, 0), vec3(0, 1, 0), 30.0f, float(nx) / float(ny), 0.1f, 10.0f);   // main.cu:192-202
    rt_context *ctx = nullptr;
    if (rt_create(0, &ctx)) { fprintf(stderr, "rt_create failed\n"); return 3; }
    if (rt_upload_world(ctx, world) || rt_apply_camera(ctx, cam, nx, ny) || rt_octree_build(ctx, 30, nullptr)) {
        fprintf(stderr, "%s\n", rt_last_error(ctx));
        return 4;
    }
    rt_render_args a{};
    a.nx = nx; a.ny = ny; a.ns = ns; a.max_depth = 50; a.use_octree = 1;
    std::vector<float> fb((size_t)nx * ny * 3);
    if (rt_render_to_host(ctx, &a, fb.data(), nullptr)) { fprintf(stderr, "%s\n", rt_last_error(ctx)); return 5; }
    FILE *f = fopen(argv[1], "wb");
    fwrite(fb.data(), sizeof(float), fb.size(), f);
    fclose(f);
    rt_destroy(ctx);
    return 0;
}
