/*
 * ref_cuda_driver.cu — ORACLE support (test infrastructure, NOT product code).
 *
 * Runs the REFERENCE'S OWN CUDA kernels (rand_init, create_world, render_init, render, free_world and the host
 * buildOctree — all from a throw-away patched copy of /root/reference/main.cu made by oracle/Makefile, never
 * committed) on the GPU, recompiled for sm_100, with the launch sequence of the reference's main()
 * (main.cu:347-477), and dumps what the reference itself never writes out: the float framebuffer, the sphere
 * list, the camera and the Octree blob, plus cudaEvent timings of every phase.  It is the golden-vector
 * generator for tests/golden/ref_cuda/ and the "reference CUDA build on the same B200" that bench.py reports
 * next to the product's number.
 *
 * Patches applied to the copy (constants are not -D overridable in the reference):
 *   main.cu:22 NUM_SPHERES <- -DRTO_N, main.cu:24 USE_OCTREE <- -DRTO_USE_OCTREE,
 *   acceleration_structure.h:15 SPHERES_PER_LEAF <- -DRTO_SPL.  Image size / spp are locals of main() and become
 *   command-line arguments here.  No kernel or device function is modified.
 *
 * usage: ref_cuda_<variant> nx ny ns[,ns2,...] [--fb f.bin] [--spheres s.bin] [--camera c.bin] [--octree o.bin] [--reps R]
 * A comma-separated ns list renders every sample count in ONE process (create_world alone takes ~2 minutes at 100 k
 * spheres: one thread device-news every material), one JSON line per count; --fb dumps the frame of the LAST count.
 */
#define main reference_main_unused
#include "main.cu"
#undef main

#include <cstdio>
#include <cstdlib>
#include <cstring>

struct flat_sphere { float cx, cy, cz, radius; int mat; float ax, ay, az, param; };

/* classify the device-heap material objects without touching reference code: vtables differ per class */
__global__ void rto_flatten(sphere (*d_list)[NUM_SPHERES], flat_sphere *out, const void *vt_l, const void *vt_m,
                            const void *vt_d) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NUM_SPHERES) return;
    const sphere &s = (*d_list)[i];
    flat_sphere o = {0, 0, 0, 0, -1, 0, 0, 0, 0};
    if (s.mat_ptr) {
        o.cx = float(s.center.x()); o.cy = float(s.center.y()); o.cz = float(s.center.z()); o.radius = float(s.radius);
        const void *vt = *(const void *const *)s.mat_ptr;
        if (vt == vt_l) { const lambertian *m = (const lambertian *)s.mat_ptr; o.mat = 0; o.ax = float(m->albedo.x()); o.ay = float(m->albedo.y()); o.az = float(m->albedo.z()); }
        else if (vt == vt_m) { const metal *m = (const metal *)s.mat_ptr; o.mat = 1; o.ax = float(m->albedo.x()); o.ay = float(m->albedo.y()); o.az = float(m->albedo.z()); o.param = float(m->fuzz); }
        else if (vt == vt_d) { const dielectric *m = (const dielectric *)s.mat_ptr; o.mat = 2; o.param = float(m->ref_idx); }
        else o.mat = -2;
    }
    out[i] = o;
}
__global__ void rto_vtables(const void **out) {
    lambertian *l = new lambertian(vec3(0, 0, 0));
    metal *m = new metal(vec3(0, 0, 0), 0);
    dielectric *d = new dielectric(1.5);
    out[0] = *(const void **)l; out[1] = *(const void **)m; out[2] = *(const void **)d;
    delete l; delete m; delete d;
}
__global__ void rto_camera_dump(camera **cam, float *out22) {
    const camera &c = **cam;
    const vec3 *v[7] = {&c.origin, &c.lower_left_corner, &c.horizontal, &c.vertical, &c.u, &c.v, &c.w};
    for (int k = 0; k < 7; k++)
        for (int e = 0; e < 3; e++) out22[3 * k + e] = float((*v[k])[e]);
    out22[21] = float(c.lens_radius);
}

/* The body of the reference's render kernel (main.cu:96-117) for a LIST of pixels instead of the whole frame: same device
 * functions (camera::get_ray, color), same per-pixel stream (the state render_init left in rand_state[pixel_index]), same
 * accumulation; the running linear sum `col` is written out after every sample, so a frame of the full-size configuration can
 * be compared pixel by pixel and sample by sample without tracing all of it.  Test infrastructure, like the rest of this file. */
__global__ void rto_render_pixels(const int *pix_ij, int npix, int max_x, int max_y, int ns, camera **cam, hitable **world,
                                  curandState *rand_state, Octree *d_octree, sphere (*d_list)[NUM_SPHERES], float *prefix) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npix) return;
    int i = pix_ij[2 * k], j = pix_ij[2 * k + 1];
    if (i < 0 || j < 0 || i >= max_x || j >= max_y) return;
    int pixel_index = j * max_x + i;
    curandState local_rand_state = rand_state[pixel_index];
    vec3 col(0, 0, 0);
    for (int s = 0; s < ns; s++) {
        real_t u = real_t(i + curand_uniform(&local_rand_state)) / real_t(max_x);
        real_t v = real_t(j + curand_uniform(&local_rand_state)) / real_t(max_y);
        ray r = (*cam)->get_ray(u, v, &local_rand_state);
        col += color(r, world, &local_rand_state, d_octree, d_list);
        for (int c = 0; c < 3; c++) prefix[((size_t)k * ns + s) * 3 + c] = float(col[c]);
    }
}

static void dump(const char *path, const void *p, size_t bytes) {
    FILE *f = fopen(path, "wb");
    if (!f || fwrite(p, 1, bytes, f) != bytes) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
    fclose(f);
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s nx ny ns [--fb f] [--spheres f] [--camera f] [--octree f] [--reps R]\n", argv[0]); return 1; }
    const int nx = atoi(argv[1]), ny = atoi(argv[2]);
    int ns_list[16], ns_count = 0;
    for (const char *q = argv[3]; *q && ns_count < 16;) {
        ns_list[ns_count++] = atoi(q);
        while (*q && *q != ',') q++;
        if (*q == ',') q++;
    }
    if (ns_count < 1 || ns_list[0] < 1) { fprintf(stderr, "bad ns list\n"); return 1; }
    const char *fb_path = 0, *sph_path = 0, *cam_path = 0, *oct_path = 0, *pix_path = 0, *pix_out = 0;
    int reps = 1, no_free = 0;
    for (int a = 4; a < argc; a++)
        if (!strcmp(argv[a], "--no-free")) no_free = 1;
    for (int a = 4; a + 1 < argc; a += 2) {
        if (!strcmp(argv[a], "--fb")) fb_path = argv[a + 1];
        else if (!strcmp(argv[a], "--spheres")) sph_path = argv[a + 1];
        else if (!strcmp(argv[a], "--camera")) cam_path = argv[a + 1];
        else if (!strcmp(argv[a], "--octree")) oct_path = argv[a + 1];
        else if (!strcmp(argv[a], "--reps")) reps = atoi(argv[a + 1]);
        else if (!strcmp(argv[a], "--pixels")) pix_path = argv[a + 1];       /* text file: "i j" per line */
        else if (!strcmp(argv[a], "--pixels-out")) pix_out = argv[a + 1];    /* npix * ns * 3 floats: running sums per sample */
    }
    const int tx = 8, ty = 8;                                     /* main.cu:351-352 */
    const size_t num_pixels = (size_t)nx * ny;
    /* create_world device-news one material per sphere plus an N-pointer array */
    size_t heap = (size_t)NUM_SPHERES * 96 + (64u << 20);
    checkCudaErrors(cudaDeviceSetLimit(cudaLimitMallocHeapSize, heap));

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_world = 0, ms_init = 0, ms_render = 0;

    vec3 *fb;
    checkCudaErrors(cudaMallocManaged(reinterpret_cast<void **>(&fb), num_pixels * sizeof(vec3)));
    curandState *d_rand_state, *d_rand_state2;
    checkCudaErrors(cudaMalloc(reinterpret_cast<void **>(&d_rand_state), num_pixels * sizeof(curandState)));
    checkCudaErrors(cudaMalloc(reinterpret_cast<void **>(&d_rand_state2), sizeof(curandState)));
    rand_init<<<1, 1>>>(d_rand_state2);
    checkCudaErrors(cudaDeviceSynchronize());

    sphere(*d_list)[NUM_SPHERES];
    checkCudaErrors(cudaMalloc(reinterpret_cast<void **>(&d_list), NUM_SPHERES * sizeof(sphere)));
    checkCudaErrors(cudaMemset(d_list, 0, NUM_SPHERES * sizeof(sphere)));   /* define the slots create_world skips (D3) */
    hitable **d_world;
    checkCudaErrors(cudaMalloc(reinterpret_cast<void **>(&d_world), sizeof(hitable *)));
    camera **d_camera;
    checkCudaErrors(cudaMalloc(reinterpret_cast<void **>(&d_camera), sizeof(camera *)));
    cudaEventRecord(e0);
    create_world<<<1, 1>>>(d_list, d_world, d_camera, nx, ny, d_rand_state2);
    cudaEventRecord(e1);
    checkCudaErrors(cudaGetLastError());
    checkCudaErrors(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms_world, e0, e1);

    sphere *cpu_spheres = static_cast<sphere *>(malloc(sizeof(sphere) * NUM_SPHERES));
    checkCudaErrors(cudaMemcpy(cpu_spheres, d_list, NUM_SPHERES * sizeof(sphere), cudaMemcpyDeviceToHost));
    clock_t c0 = clock();
    Octree *octree = buildOctree(cpu_spheres, NUM_SPHERES);
    double ms_build = 1e3 * double(clock() - c0) / CLOCKS_PER_SEC;
    Octree *d_octree;
    checkCudaErrors(cudaMalloc(reinterpret_cast<void **>(&d_octree), sizeof(Octree)));
    checkCudaErrors(cudaMemcpy(d_octree, octree, sizeof(Octree), cudaMemcpyHostToDevice));
    checkCudaErrors(cudaDeviceSynchronize());

    dim3 blocks((nx + tx - 1) / tx, (ny + ty - 1) / ty);
    dim3 threads(tx, ty);
    int use_octree = 0, fp16 = 0;
#ifdef USE_OCTREE
    use_octree = 1;
#endif
#ifdef USE_FP16
    fp16 = 1;
#endif
    if (pix_path && pix_out) {      /* pixel-list mode: no full frame is traced */
        FILE *pf = fopen(pix_path, "r");
        if (!pf) { fprintf(stderr, "cannot read %s\n", pix_path); return 2; }
        int cap = 1024, npix = 0, *h_ij = (int *)malloc(sizeof(int) * 2 * cap);
        while (fscanf(pf, "%d %d", &h_ij[2 * npix], &h_ij[2 * npix + 1]) == 2) {
            if (++npix == cap) { cap *= 2; h_ij = (int *)realloc(h_ij, sizeof(int) * 2 * cap); }
        }
        fclose(pf);
        const int ns = ns_list[0];
        int *d_ij;
        float *d_prefix;
        checkCudaErrors(cudaMalloc(&d_ij, sizeof(int) * 2 * (npix + 1)));
        checkCudaErrors(cudaMalloc(&d_prefix, sizeof(float) * 3 * (size_t)ns * (npix + 1)));
        checkCudaErrors(cudaMemset(d_prefix, 0, sizeof(float) * 3 * (size_t)ns * (npix + 1)));
        checkCudaErrors(cudaMemcpy(d_ij, h_ij, sizeof(int) * 2 * npix, cudaMemcpyHostToDevice));
        render_init<<<blocks, threads>>>(nx, ny, d_rand_state);
        checkCudaErrors(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        rto_render_pixels<<<(npix + 31) / 32, 32>>>(d_ij, npix, nx, ny, ns, d_camera, d_world, d_rand_state, d_octree, d_list, d_prefix);
        cudaEventRecord(e1);
        checkCudaErrors(cudaGetLastError());
        checkCudaErrors(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms_render, e0, e1);
        float *h_prefix = (float *)malloc(sizeof(float) * 3 * (size_t)ns * (npix + 1));
        checkCudaErrors(cudaMemcpy(h_prefix, d_prefix, sizeof(float) * 3 * (size_t)ns * npix, cudaMemcpyDeviceToHost));
        dump(pix_out, h_prefix, sizeof(float) * 3 * (size_t)ns * npix);
        printf("{\"impl\": \"ref_cuda\", \"mode\": \"pixels\", \"n\": %d, \"spl\": %d, \"use_octree\": %d, \"fp16\": %d, \"nx\": %d, \"ny\": %d, "
               "\"ns\": %d, \"pixels\": %d, \"render_ms\": %.3f, \"create_world_ms\": %.3f}\n",
               NUM_SPHERES, SPHERES_PER_LEAF, use_octree, fp16, nx, ny, ns, npix, ms_render, ms_world);
        ns_count = 0;
        free(h_prefix); free(h_ij); cudaFree(d_ij); cudaFree(d_prefix);
    }
    for (int k = 0; k < ns_count; k++) {
    const int ns = ns_list[k];
    float best_render = 1e30f, best_init = 1e30f, sum_render = 0;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        render_init<<<blocks, threads>>>(nx, ny, d_rand_state);
        cudaEventRecord(e1);
        checkCudaErrors(cudaGetLastError());
        checkCudaErrors(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms_init, e0, e1);
        cudaEventRecord(e0);
        render<<<blocks, threads>>>(fb, nx, ny, ns, d_camera, d_world, d_rand_state, d_octree, d_list);
        cudaEventRecord(e1);
        checkCudaErrors(cudaGetLastError());
        checkCudaErrors(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms_render, e0, e1);
        if (ms_render < best_render) best_render = ms_render;
        if (ms_init < best_init) best_init = ms_init;
        sum_render += ms_render;
    }
    printf("{\"impl\": \"ref_cuda\", \"n\": %d, \"spl\": %d, \"use_octree\": %d, \"fp16\": %d, \"nx\": %d, \"ny\": %d, \"ns\": %d, "
           "\"reps\": %d, \"create_world_ms\": %.3f, \"octree_build_host_ms\": %.3f, \"octree_bytes\": %zu, "
           "\"render_init_ms\": %.4f, \"render_ms\": %.4f, \"render_ms_mean\": %.4f, \"node_count\": %d, \"leaf_count\": %d}\n",
           NUM_SPHERES, SPHERES_PER_LEAF, use_octree, fp16, nx, ny, ns, reps, ms_world, ms_build, sizeof(Octree),
           best_init, best_render, sum_render / reps, octree->nodeCount, octree->leafCount);
    fflush(stdout);
    }

    if (fb_path) dump(fb_path, fb, num_pixels * sizeof(vec3));
    if (oct_path) dump(oct_path, octree, sizeof(Octree));
    if (cam_path) {
        float *d_c, h_c[22];
        cudaMalloc(&d_c, sizeof h_c);
        rto_camera_dump<<<1, 1>>>(d_camera, d_c);
        checkCudaErrors(cudaMemcpy(h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost));
        dump(cam_path, h_c, sizeof h_c);
        cudaFree(d_c);
    }
    if (sph_path) {
        const void **d_vt, *h_vt[3];
        cudaMalloc(&d_vt, sizeof h_vt);
        rto_vtables<<<1, 1>>>(d_vt);
        checkCudaErrors(cudaMemcpy(h_vt, d_vt, sizeof h_vt, cudaMemcpyDeviceToHost));
        flat_sphere *d_f, *h_f = (flat_sphere *)malloc(sizeof(flat_sphere) * NUM_SPHERES);
        cudaMalloc(&d_f, sizeof(flat_sphere) * NUM_SPHERES);
        rto_flatten<<<(NUM_SPHERES + 255) / 256, 256>>>(d_list, d_f, h_vt[0], h_vt[1], h_vt[2]);
        checkCudaErrors(cudaMemcpy(h_f, d_f, sizeof(flat_sphere) * NUM_SPHERES, cudaMemcpyDeviceToHost));
        dump(sph_path, h_f, sizeof(flat_sphere) * NUM_SPHERES);
        free(h_f); cudaFree(d_f); cudaFree(d_vt);
    }

    if (!no_free) {     /* one device thread deletes every material: ~90 s at 100 k spheres; process exit releases the heap as well */
        free_world<<<1, 1>>>(d_list, d_world, d_camera);
        checkCudaErrors(cudaDeviceSynchronize());
    }
    cudaFree(d_camera); cudaFree(d_world); cudaFree(d_list); cudaFree(d_rand_state); cudaFree(d_rand_state2);
    cudaFree(fb); cudaFree(d_octree);
    delete octree;
    free(cpu_spheres);
    return 0;
}
