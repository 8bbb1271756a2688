// raytracing_main.cpp — the reference's CLI, `RayTracing [mode]` (main.cu:347-477), over the C ABI.
//
//   mode 0 = P3 PPM to stdout (default), 1 = no output, 3 = output.ppm; 2 (OpenGL preview) is accepted and ignored.
// Same banner on stderr, same "took X seconds." line, same exit(99) on a CUDA error.  The reference's knobs keep
// their names; they are #ifndef-guarded so -D works, and can be overridden at run time without touching the
// positional argument: RT_NUM_SPHERES, RT_SPHERES_PER_LEAF, RT_USE_OCTREE, RT_USE_FP16, RT_NX, RT_NY, RT_NS
// (and RT_SEED_MODE=1 for the upstream per-pixel seeding curand_init(1984, pixel_index, 0), main.cu:90).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "rt_abi.h"

#ifndef NUM_SPHERES
#define NUM_SPHERES 8000          // main.cu:22
#endif
#ifndef SPHERE_RADIUS
#define SPHERE_RADIUS 0.1f        // main.cu:23
#endif
#ifndef SPHERES_PER_LEAF
#define SPHERES_PER_LEAF 30       // acceleration_structure.h:15
#endif
#ifndef RT_NO_OCTREE
#define USE_OCTREE                // main.cu:24
#endif

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// check_cuda (main.cu:29-37): message, then exit(99)
static void check(rt_context *ctx, int rc, const char *what, const char *file, int line) {
    if (rc) {
        std::cerr << "CUDA error = " << static_cast<unsigned int>(rc) << " at " << file << ":" << line << " '" << what << "' \n";
        if (ctx) std::cerr << rt_last_error(ctx) << "\n";
        exit(99);
    }
}
#define CHECK(call) check(ctx, (call), #call, __FILE__, __LINE__)

int main(int argc, char **argv) {
    const int nx = env_int("RT_NX", 1200), ny = env_int("RT_NY", 800), ns = env_int("RT_NS", 10);   // main.cu:348-350
    const int tx = 8, ty = 8;
    const int n = env_int("RT_NUM_SPHERES", NUM_SPHERES);
    const int spl = env_int("RT_SPHERES_PER_LEAF", SPHERES_PER_LEAF);
#ifdef USE_OCTREE
    const int use_octree = env_int("RT_USE_OCTREE", 1);
#else
    const int use_octree = env_int("RT_USE_OCTREE", 0);
#endif
    std::cerr << "Rendering a " << nx << "x" << ny << " image with " << ns << " samples per pixel ";
    std::cerr << "in " << tx << "x" << ty << " blocks.\n";
    std::cerr << "Number of spheres: " << n << "\n";
    std::cerr << "Sphere radius: " << SPHERE_RADIUS << "\n";
    std::cerr << (use_octree ? "Use octree: ON\n" : "Use octree: OFF\n");
#ifdef USE_FP16                   // precision_types.h:8
    const int precision = env_int("RT_USE_FP16", 1) ? RT_PREC_FP16 : RT_PREC_FP32;
#else
    const int precision = env_int("RT_USE_FP16", 0) ? RT_PREC_FP16 : RT_PREC_FP32;
#endif
    int output_mode = 0;
    if (argc > 1) output_mode = std::stoi(argv[1]);      // throws on garbage, as the reference does
    std::cerr << "Output mode: " << output_mode << "\n";

    rt_context *ctx = nullptr;
    CHECK(rt_create(0, &ctx));
    CHECK(rt_scene_generate_ex(ctx, n, SPHERE_RADIUS, precision));
    if (use_octree) CHECK(rt_octree_build_ex(ctx, spl, precision, nullptr));
    CHECK(rt_camera_set(ctx, nullptr, nx, ny));

    rt_render_args a{};
    a.nx = nx; a.ny = ny; a.ns = ns; a.max_depth = 50; a.use_octree = use_octree;
    a.precision = precision;
    a.seed_mode = env_int("RT_SEED_MODE", RT_SEED_HEAD);
    rt_render_stats st{};
    const auto t0 = std::chrono::steady_clock::now();
    size_t need = 0;
    const bool want_text = output_mode == 0 || output_mode == 3;
    std::vector<float> fb(want_text ? 0 : (size_t)nx * ny * 3);
    if (want_text) CHECK(rt_render_to_ppm(ctx, &a, &st, &need));          // the frame is quantised and formatted on the GPU
    else CHECK(rt_render_to_host(ctx, &a, fb.data(), &st));
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cerr << "took " << secs << " seconds.\n";
    if (getenv("RT_VERBOSE"))
        std::cerr << "rays " << st.rays << ", kernel " << st.kernel_ms << " ms, " << (st.rays / (st.kernel_ms * 1e3)) << " Mrays/s\n";

    if (output_mode == 0 || output_mode == 3) {
        std::string txt(need, '\0');
        CHECK(rt_ppm_read(ctx, &txt[0], need));
        if (output_mode == 0) {
            std::cout.write(txt.data(), (std::streamsize)need);
        } else {
            std::ofstream out("output.ppm");
            out.write(txt.data(), (std::streamsize)need);
        }
    }
    rt_destroy(ctx);
    return 0;
}
