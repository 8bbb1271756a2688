/*
 * rt_dropin.h — the reference's world-building class surface, host side, over the C ABI of rt_abi.h.
 *
 * In the reference these classes live on the device: create_world (main.cu:146-204) news up `sphere`, `lambertian`,
 * `metal`, `dielectric`, a `hitable_list` and a `camera` inside a one-thread kernel, and render chases their vtables.
 * Code written against that surface keeps compiling against this header — same class names, constructor signatures
 * and accessors (vec3.h, ray.h, hitable.h:11-21, sphere.h:7-15, hitable_list.h:6-14, material.h:47-116,
 * camera.h:20-58) — but the objects are plain host descriptions: `rt_upload_world` flattens a `hitable` tree into the
 * SoA scene of librt_b200.so (rt_scene_upload) and `rt_apply_camera` hands the constructor arguments to
 * rt_camera_set.  Nothing here runs on the GPU; the rendering arithmetic is the library's.
 *
 * `hitable::hit` is provided for host-side picking / debugging with the reference's semantics (closest hit, strict
 * '<'); it is NOT the render path.  `material::scatter` has no host equivalent (it draws from a device cuRAND state);
 * materials only describe themselves.
 */
#ifndef RT_DROPIN_H
#define RT_DROPIN_H

#include <math.h>

#include <vector>

#include "rt_abi.h"

typedef float real_t;                                   /* precision_types.h:179; USE_FP16 is a run-time switch of the library */

class vec3 {
public:
    vec3() { e[0] = e[1] = e[2] = 0; }
    vec3(real_t e0, real_t e1, real_t e2) { e[0] = e0; e[1] = e1; e[2] = e2; }
    real_t x() const { return e[0]; }
    real_t y() const { return e[1]; }
    real_t z() const { return e[2]; }
    real_t r() const { return e[0]; }
    real_t g() const { return e[1]; }
    real_t b() const { return e[2]; }
    const vec3 &operator+() const { return *this; }
    vec3 operator-() const { return vec3(-e[0], -e[1], -e[2]); }
    real_t operator[](int i) const { return e[i]; }
    real_t &operator[](int i) { return e[i]; }
    vec3 &operator+=(const vec3 &v) { for (int i = 0; i < 3; i++) e[i] += v.e[i]; return *this; }
    vec3 &operator-=(const vec3 &v) { for (int i = 0; i < 3; i++) e[i] -= v.e[i]; return *this; }
    vec3 &operator*=(const vec3 &v) { for (int i = 0; i < 3; i++) e[i] *= v.e[i]; return *this; }
    vec3 &operator/=(const vec3 &v) { for (int i = 0; i < 3; i++) e[i] /= v.e[i]; return *this; }
    vec3 &operator*=(real_t t) { for (int i = 0; i < 3; i++) e[i] *= t; return *this; }
    vec3 &operator/=(real_t t) { const real_t k = (real_t)(1.0 / t); for (int i = 0; i < 3; i++) e[i] *= k; return *this; }
    real_t squared_length() const { return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; }
    real_t length() const { return sqrtf(squared_length()); }
    void make_unit_vector() { *this /= length(); }
    real_t e[3];
};
inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.e[0] + b.e[0], a.e[1] + b.e[1], a.e[2] + b.e[2]); }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.e[0] - b.e[0], a.e[1] - b.e[1], a.e[2] - b.e[2]); }
inline vec3 operator*(const vec3 &a, const vec3 &b) { return vec3(a.e[0] * b.e[0], a.e[1] * b.e[1], a.e[2] * b.e[2]); }
inline vec3 operator/(const vec3 &a, const vec3 &b) { return vec3(a.e[0] / b.e[0], a.e[1] / b.e[1], a.e[2] / b.e[2]); }
inline vec3 operator*(real_t t, const vec3 &v) { return vec3(t * v.e[0], t * v.e[1], t * v.e[2]); }
inline vec3 operator*(const vec3 &v, real_t t) { return t * v; }
inline vec3 operator/(const vec3 &v, real_t t) { return vec3(v.e[0] / t, v.e[1] / t, v.e[2] / t); }
inline real_t dot(const vec3 &a, const vec3 &b) { return a.e[0] * b.e[0] + a.e[1] * b.e[1] + a.e[2] * b.e[2]; }
inline vec3 cross(const vec3 &a, const vec3 &b) {
    return vec3(a.e[1] * b.e[2] - a.e[2] * b.e[1], -(a.e[0] * b.e[2] - a.e[2] * b.e[0]), a.e[0] * b.e[1] - a.e[1] * b.e[0]);
}
inline vec3 unit_vector(const vec3 &v) { return v / v.length(); }

class ray {
public:
    ray() {}
    ray(const vec3 &a, const vec3 &b) : A(a), B(b) {}
    vec3 origin() const { return A; }
    vec3 direction() const { return B; }
    vec3 point_at_parameter(real_t t) const { return A + t * B; }
    vec3 A, B;
};

/* ---- materials (material.h:47-116): descriptions only ------------------------------------------------------------------ */
class material {
public:
    virtual ~material() {}
    virtual void describe(rt_sphere_desc &d) const = 0;      /* fills mat, albedo, param */
};
class lambertian : public material {
public:
    explicit lambertian(const vec3 &a) : albedo(a) {}
    void describe(rt_sphere_desc &d) const override { d.mat = RT_MAT_LAMBERTIAN; d.ax = albedo.x(); d.ay = albedo.y(); d.az = albedo.z(); d.param = 0; }
    vec3 albedo;
};
class metal : public material {
public:
    metal(const vec3 &a, real_t f) : albedo(a), fuzz(f < 1.0f ? f : 1.0f) {}       /* the clamp of material.h:66 */
    void describe(rt_sphere_desc &d) const override { d.mat = RT_MAT_METAL; d.ax = albedo.x(); d.ay = albedo.y(); d.az = albedo.z(); d.param = fuzz; }
    vec3 albedo;
    real_t fuzz;
};
class dielectric : public material {
public:
    explicit dielectric(real_t ri) : ref_idx(ri) {}
    void describe(rt_sphere_desc &d) const override { d.mat = RT_MAT_DIELECTRIC; d.ax = d.ay = d.az = 0; d.param = ref_idx; }
    real_t ref_idx;
};

/* ---- geometry (hitable.h, sphere.h, hitable_list.h) --------------------------------------------------------------------- */
struct hit_record {
    real_t t;
    vec3 p;
    vec3 normal;
    material *mat_ptr;
};
class hitable {
public:
    virtual ~hitable() {}
    virtual bool hit(const ray &r, real_t t_min, real_t t_max, hit_record &rec) const = 0;
    virtual void flatten(std::vector<rt_sphere_desc> &out) const = 0;     /* appends this object's spheres, in list order */
};
class sphere : public hitable {
public:
    sphere() : radius(0), mat_ptr(nullptr) {}
    sphere(vec3 cen, real_t r, material *m) : center(cen), radius(r), mat_ptr(m) {}
    bool hit(const ray &r, real_t t_min, real_t t_max, hit_record &rec) const override {
        const vec3 oc = r.origin() - center;
        const real_t a = dot(r.direction(), r.direction()), b = dot(oc, r.direction()), c = dot(oc, oc) - radius * radius;
        const real_t disc = b * b - a * c;
        if (!(disc > 0)) return false;
        const real_t sq = sqrtf(disc);
        for (int k = 0; k < 2; k++) {
            const real_t t = (k == 0 ? (-b - sq) : (-b + sq)) / a;
            if (t < t_max && t > t_min) {
                rec.t = t;
                rec.p = r.point_at_parameter(t);
                rec.normal = (rec.p - center) / radius;
                rec.mat_ptr = mat_ptr;
                return true;
            }
        }
        return false;
    }
    void flatten(std::vector<rt_sphere_desc> &out) const override {
        rt_sphere_desc d;
        d.cx = center.x(); d.cy = center.y(); d.cz = center.z(); d.radius = radius;
        d.mat = RT_MAT_NONE; d.ax = d.ay = d.az = d.param = 0;
        if (mat_ptr) mat_ptr->describe(d);
        out.push_back(d);
    }
    vec3 center;
    real_t radius;
    material *mat_ptr;
};
class hitable_list : public hitable {
public:
    hitable_list() : list(nullptr), list_size(0) {}
    hitable_list(hitable **l, int n) : list(l), list_size(n) {}
    bool hit(const ray &r, real_t t_min, real_t t_max, hit_record &rec) const override {
        hit_record tmp;
        bool any = false;
        real_t closest = t_max;
        for (int i = 0; i < list_size; i++)
            if (list[i]->hit(r, t_min, closest, tmp)) { any = true; closest = tmp.t; rec = tmp; }
        return any;
    }
    void flatten(std::vector<rt_sphere_desc> &out) const override { for (int i = 0; i < list_size; i++) list[i]->flatten(out); }
    hitable **list;
    int list_size;
};

/* ---- camera (camera.h:20-58): keeps the constructor's arguments ------------------------------------------------------------ */
class camera {
public:
    camera(vec3 lookfrom, vec3 lookat, vec3 vup, real_t vfov, real_t aspect, real_t aperture, real_t focus_dist) {
        for (int k = 0; k < 3; k++) { desc.lookfrom[k] = lookfrom[k]; desc.lookat[k] = lookat[k]; desc.vup[k] = vup[k]; }
        desc.vfov = vfov; desc.aspect = aspect; desc.aperture = aperture; desc.focus_dist = focus_dist;
    }
    rt_camera_desc desc;
};

/* the world a `hitable*` describes -> the library's scene; index 0 of the flattened list is the sphere hitTree tests
 * unconditionally (the "ground" of acceleration_structure.h:322), as in the reference's d_list */
inline int rt_upload_world(rt_context *ctx, const hitable *world) {
    std::vector<rt_sphere_desc> flat;
    world->flatten(flat);
    return rt_scene_upload(ctx, flat.data(), (int)flat.size());
}
inline int rt_apply_camera(rt_context *ctx, const camera &cam, int nx, int ny) { return rt_camera_set(ctx, &cam.desc, nx, ny); }

#endif
