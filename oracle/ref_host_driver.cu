/*
 * ref_host_driver.cu — CPU ORACLE support (test infrastructure, NOT product code).
 *
 * Harness that host-compiles the REFERENCE'S OWN sources (taken where they lie under /root/reference by
 * oracle/Makefile, never committed here) and exposes them through a small C API, so that
 *   (a) oracle/rt_oracle.c (the restatement) can be validated against the real code, and
 *   (b) bench.py has a CPU baseline of kind "reference" (OpenMP over image rows).
 *
 * What runs is the reference's code: sphere::hit, hitable_list::hit, buildOctree/insert, hitTree/traverseTree,
 * lambertian/metal/dielectric::scatter, camera, and color() from main.cu.  What this file restates (because the
 * originals are __global__ kernels that cannot run on the host) is only:
 *   - create_world's body (main.cu:146-204) with the device's left-to-right argument evaluation,
 *   - render's per-pixel sample loop (main.cu:96-117).
 * The Makefile applies four documented one-line patches to a throw-away copy of the sources:
 *   main.cu: USE_OCTREE made switchable (-DRTO_USE_OCTREE), acceleration_structure.h: SPHERES_PER_LEAF taken
 *   from -DRTO_SPL, material.h:33 and camera.h:15: the RNG draws sequenced left to right as the device does.
 *
 * Shim: cuRAND's QUALIFIERS made __host__ __device__, and __device__ functions made host-callable.
 */
#define QUALIFIERS static __forceinline__ __host__ __device__
#include <curand_kernel.h>
#include <cuda_fp16.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <fstream>
#include <iostream>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#undef __device__
#define __device__ __host__ __location__(device)

/* instrumentation counters (thread-local), bumped by hooks the Makefile patches into the throw-away copy
 * (sphere.h:18, acceleration_structure.h:227, main.cu:48); compiled out of the (unused) device pass */
static thread_local uint64_t g_sphere_tests, g_aabb_tests, g_rays;
#ifdef __CUDA_ARCH__
#define RTO_COUNT_SPHERE
#define RTO_COUNT_AABB
#define RTO_COUNT_RAY
#else
#define RTO_COUNT_SPHERE (g_sphere_tests++);
#define RTO_COUNT_AABB (g_aabb_tests++);
#define RTO_COUNT_RAY (g_rays++);
#endif

#define main reference_main_unused
#include "main.cu" /* the patched throw-away copy; pulls in every reference header */
#undef main

extern "C" {

struct refh_sphere {          /* same layout as rto_sphere (oracle/rt_oracle.h) */
    float cx, cy, cz, radius;
    int32_t mat;
    float ax, ay, az, param;
};
struct refh_counters { uint64_t rays, sphere_tests, aabb_tests, paths; uint32_t max_depth; };
struct refh_params {          /* same layout as rto_render_params */
    int nx, ny, ns, use_octree, spl, arith, seed_mode, max_depth;
    int i0, i1, istep, j0, j1, jstep, threads;
};

struct refh_world {
    int n;
    sphere *list;             /* the reference's AoS sphere array */
    hitable **ptrs;
    hitable *world;           /* hitable_list */
    hitable **d_world;
    camera *cam;
    Octree *octree;
    std::vector<material *> mats;
    std::unordered_map<const material *, int> mat_index;
};

#define RND (curand_uniform(&local_rand_state))

/* main.cu:146-204 with explicit left-to-right draws */
refh_world *refh_create_world(int n, float sphere_radius, int nx, int ny) {
    refh_world *w = new refh_world();
    w->n = n;
    w->list = static_cast<sphere *>(calloc((size_t)n, sizeof(sphere)));   /* zero-filled */
    w->mats.assign((size_t)n, nullptr);
    curandState local_rand_state;
    curand_init(1984, 0, 0, &local_rand_state);                            /* rand_init, main.cu:80 */
    auto put = [&](int i, vec3 c, real_t r, material *m) {
        new (&w->list[i]) sphere(c, r, m);
        w->mats[(size_t)i] = m;
        w->mat_index[m] = i;
    };
    put(0, vec3(0, -1000.0, -1), 1000, new lambertian(vec3(0.5, 0.5, 0.5)));
    int i = 1;
    put(i++, vec3(0, 1, 0), 1.0, new dielectric(1.5));
    put(i++, vec3(-4, 1, 0), 1.0, new lambertian(vec3(0.4, 0.2, 0.1)));
    put(i++, vec3(4, 1, 0), 1.0, new metal(vec3(0.7, 0.6, 0.5), 0.0));
    const int spheres_per_dim = sqrtf((float)n - 4);
    const double spacing = 20. / spheres_per_dim;
    for (double a = -10; a < 10; a += spacing) {
        for (double b = -10; b < 10 && i < n; b += spacing) {
            const real_t choose_mat = RND;
            const float r1 = RND;
            const float r2 = RND;
            const vec3 center(a + r1, sphere_radius, b + r2);
            if (choose_mat < real_t(0.8f)) {
                const float q0 = RND, q1 = RND, q2 = RND, q3 = RND, q4 = RND, q5 = RND;
                put(i++, center, sphere_radius, new lambertian(vec3(q0 * q1, q2 * q3, q4 * q5)));
            } else if (choose_mat < real_t(0.95f)) {
                const float q0 = RND, q1 = RND, q2 = RND, q3 = RND;
                put(i++, center, sphere_radius,
                    new metal(vec3(0.5f * (1.0f + q0), 0.5f * (1.0f + q1), 0.5f * (1.0f + q2)), 0.5f * q3));
            } else {
                put(i++, center, sphere_radius, new dielectric(1.5));
            }
        }
    }
    w->ptrs = new hitable *[n];
    for (int k = 0; k < n; k++) w->ptrs[k] = &w->list[k];
    w->world = new hitable_list(w->ptrs, n);
    w->d_world = new hitable *[1];
    w->d_world[0] = w->world;
    const vec3 lookfrom(13, 2, 3);
    const vec3 lookat(0, 0, 0);
    const real_t dist_to_focus = 10.0;
    const real_t aperture = 0.1;
    w->cam = new camera(lookfrom, lookat, vec3(0, 1, 0), 30.0, real_t(nx) / real_t(ny), aperture, dist_to_focus);
    w->octree = nullptr;
    return w;
}

int refh_initialised(const refh_world *w) {
    int c = 0;
    for (int k = 0; k < w->n; k++) c += w->mats[(size_t)k] != nullptr;
    return c;
}

/* flatten the reference objects for comparison with the restated scene */
void refh_export_spheres(const refh_world *w, refh_sphere *out) {
    for (int k = 0; k < w->n; k++) {
        const sphere &s = w->list[k];
        refh_sphere o;
        memset(&o, 0, sizeof o);
        o.mat = -1;
        material *m = w->mats[(size_t)k];
        if (m) {
            o.cx = float(s.center.x()); o.cy = float(s.center.y()); o.cz = float(s.center.z());
            o.radius = float(s.radius);
            if (auto *l = dynamic_cast<lambertian *>(m)) {
                o.mat = 0; o.ax = float(l->albedo.x()); o.ay = float(l->albedo.y()); o.az = float(l->albedo.z());
            } else if (auto *mt = dynamic_cast<metal *>(m)) {
                o.mat = 1; o.ax = float(mt->albedo.x()); o.ay = float(mt->albedo.y()); o.az = float(mt->albedo.z());
                o.param = float(mt->fuzz);
            } else if (auto *d = dynamic_cast<dielectric *>(m)) {
                o.mat = 2; o.param = float(d->ref_idx);
            }
        }
        out[k] = o;
    }
}

void refh_export_camera(const refh_world *w, float *out22) {
    const camera &c = *w->cam;
    const vec3 *v[7] = {&c.origin, &c.lower_left_corner, &c.horizontal, &c.vertical, &c.u, &c.v, &c.w};
    for (int k = 0; k < 7; k++)
        for (int e = 0; e < 3; e++) out22[3 * k + e] = float((*v[k])[e]);
    out22[21] = float(c.lens_radius);
}

size_t refh_octree_sizeof(void) { return sizeof(Octree); }
int refh_spl(void) { return SPHERES_PER_LEAF; }
int refh_use_octree(void) {
#ifdef USE_OCTREE
    return 1;
#else
    return 0;
#endif
}
int refh_fp16(void) {
#ifdef USE_FP16
    return 1;
#else
    return 0;
#endif
}

/* acceleration_structure.h:195 — the reference's own serial build; blob receives sizeof(Octree) bytes */
void refh_build_octree(refh_world *w, void *blob) {
    if (w->octree) delete w->octree;
    w->octree = buildOctree(w->list, w->n);
    if (blob) memcpy(blob, w->octree, sizeof(Octree));
}

/* main.cu:96-117 pixel loop around the reference's own color() */
int refh_render(refh_world *w, const refh_params *p, float *fb_gamma, float *fb_linear, refh_counters *out) {
#ifdef USE_OCTREE
    if (!w->octree) refh_build_octree(w, nullptr);
#endif
    refh_counters total;
    memset(&total, 0, sizeof total);
    const int nrows = (p->j1 - p->j0 + p->jstep - 1) / p->jstep;
#ifdef _OPENMP
    if (p->threads > 0) omp_set_num_threads(p->threads);
#endif
#pragma omp parallel
    {
        uint64_t paths = 0;
        g_sphere_tests = 0;
        g_aabb_tests = 0;
        g_rays = 0;
#pragma omp for schedule(dynamic, 1)
        for (int r = 0; r < nrows; r++) {
            const int j = p->j0 + r * p->jstep;
            for (int i = p->i0; i < p->i1; i += p->istep) {
                const int pixel_index = j * p->nx + i;
                curandState local_rand_state;
                if (p->seed_mode == 1) curand_init(1984, pixel_index, 0, &local_rand_state);   /* main.cu:90, the upstream form */
                else curand_init(1984 + pixel_index, 0, 0, &local_rand_state);                 /* main.cu:93, HEAD */
                vec3 col(0, 0, 0);
                for (int s = 0; s < p->ns; s++) {
                    real_t u = real_t(i + curand_uniform(&local_rand_state)) / real_t(p->nx);
                    real_t v = real_t(j + curand_uniform(&local_rand_state)) / real_t(p->ny);
                    ray ry = w->cam->get_ray(u, v, &local_rand_state);
                    col += color(ry, w->d_world, &local_rand_state, w->octree, nullptr);
                    paths++;
                }
                if (fb_linear)
                    for (int c = 0; c < 3; c++) fb_linear[3 * (size_t)pixel_index + c] = float(col[c]);
                col /= real_t(p->ns);
                col[0] = sqrt(col[0]);
                col[1] = sqrt(col[1]);
                col[2] = sqrt(col[2]);
                if (fb_gamma)
                    for (int c = 0; c < 3; c++) fb_gamma[3 * (size_t)pixel_index + c] = float(col[c]);
            }
        }
#pragma omp critical
        {
            total.paths += paths;
            total.rays += g_rays;
            total.sphere_tests += g_sphere_tests;
            total.aabb_tests += g_aabb_tests;
        }
    }
    if (out) *out = total;
    return 0;
}

/* one ray through the reference's own closest-hit code; returns sphere index or -1 */
/* cuRAND itself (the toolkit header, host-compiled): state words {d, v0..v4} after curand_init(seed, subsequence, 0) */
void refh_curand_state(unsigned long long seed, unsigned long long subsequence, unsigned int out6[6]) {
    curandState s;
    curand_init(seed, subsequence, 0, &s);
    out6[0] = s.d;
    for (int k = 0; k < 5; k++) out6[1 + k] = s.v[k];
}

int refh_closest_hit(refh_world *w, const float *o, const float *d, float *t_out) {
    ray r(vec3(o[0], o[1], o[2]), vec3(d[0], d[1], d[2]));
    hit_record rec;
    bool hit;
#ifdef USE_OCTREE
    if (!w->octree) refh_build_octree(w, nullptr);
    hit = hitTree(w->octree, r, rec, w->d_world);
#else
    hit = w->world->hit(r, 0.001f, FLT_MAX, rec);
#endif
    if (!hit) return -1;
    *t_out = float(rec.t);
    auto it = w->mat_index.find(rec.mat_ptr);
    return it == w->mat_index.end() ? -2 : it->second;
}

void refh_destroy(refh_world *w) {
    if (!w) return;
    for (material *m : w->mats) {   /* material has no virtual destructor: delete through the concrete type */
        if (auto *l = dynamic_cast<lambertian *>(m)) delete l;
        else if (auto *mt = dynamic_cast<metal *>(m)) delete mt;
        else if (auto *d = dynamic_cast<dielectric *>(m)) delete d;
    }
    delete[] w->ptrs;
    delete (hitable_list *)w->world;
    delete[] w->d_world;
    delete w->cam;
    if (w->octree) delete w->octree;
    free(w->list);
    delete w;
}

} /* extern "C" */
