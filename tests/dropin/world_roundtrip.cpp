// Test program for include/rt_dropin.h (no GPU needed): rebuilds a scene through the reference's class surface —
// sphere(center, radius, new <material>(...)), hitable_list(list, n), camera(...) — from a flat description read on
// stdin, flattens it back and prints it; also answers a few closest-hit queries with hitable_list::hit.
//   stdin : n, then n lines "cx cy cz r mat ax ay az param", then q, then q lines "ox oy oz dx dy dz"
//   stdout: n lines of the flattened descriptors (hex floats), then q lines "idx t"
#include <cstdio>
#include <vector>

#include "rt_dropin.h"

int main() {
    int n = 0;
    if (scanf("%d", &n) != 1) return 1;
    std::vector<hitable *> list;
    std::vector<material *> mats;
    for (int i = 0; i < n; i++) {
        float cx, cy, cz, r, ax, ay, az, p;
        int mat;
        if (scanf("%a %a %a %a %d %a %a %a %a", &cx, &cy, &cz, &r, &mat, &ax, &ay, &az, &p) != 9) return 2;
        material *m = nullptr;
        if (mat == RT_MAT_LAMBERTIAN) m = new lambertian(vec3(ax, ay, az));
        else if (mat == RT_MAT_METAL) m = new metal(vec3(ax, ay, az), p);
        else if (mat == RT_MAT_DIELECTRIC) m = new dielectric(p);
        mats.push_back(m);
        list.push_back(new sphere(vec3(cx, cy, cz), r, m));
    }
    hitable *world = new hitable_list(list.data(), n);
    std::vector<rt_sphere_desc> flat;
    world->flatten(flat);
    for (const rt_sphere_desc &d : flat) printf("%a %a %a %a %d %a %a %a %a\n", d.cx, d.cy, d.cz, d.radius, d.mat, d.ax, d.ay, d.az, d.param);
    int q = 0;
    if (scanf("%d", &q) != 1) return 3;
    for (int k = 0; k < q; k++) {
        float o[3], d[3];
        if (scanf("%a %a %a %a %a %a", &o[0], &o[1], &o[2], &d[0], &d[1], &d[2]) != 6) return 4;
        hit_record rec;
        int idx = -1;
        const bool hit = world->hit(ray(vec3(o[0], o[1], o[2]), vec3(d[0], d[1], d[2])), 0.001f, 3.402823466e+38f, rec);
        if (hit)
            for (int i = 0; i < n; i++)
                if (mats[(size_t)i] && mats[(size_t)i] == rec.mat_ptr) { idx = i; break; }     // one material object per sphere
        printf("%d %a\n", idx, hit ? rec.t : 0.f);
    }
    camera cam(vec3(13, 2, 3), vec3(0, 0, 0), vec3(0, 1, 0), 30.0f, 1.5f, 0.1f, 10.0f);
    printf("camera %a %a %a %a\n", cam.desc.vfov, cam.desc.aspect, cam.desc.aperture, cam.desc.focus_dist);
    return 0;
}
