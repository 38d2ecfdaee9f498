/*
 * oracle/raycast.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of what the reference's render_scene() (example_renderer.py:153-157:
 * mi.load_file + mi.render) computes for the scene generate_xml_content() emits
 * (example_renderer.py:113-128): which sphere (or the ground rectangle, or nothing)
 * each pixel-centre camera ray sees first, and a shaded sRGB8 image of it.
 *
 * PARITY UNPINNED for the image: the arithmetic of render_scene lives in Mitsuba 3
 * (requirements.txt:3, no version pinned), which is not in /root/reference and cannot be
 * installed here; the reference ships no tests, golden images or sample data.  What IS
 * pinned: the scene handed to this file (centres, radius, camera, floor, light) is the
 * one parsed back out of the reference's own generate_xml_content() output
 * (oracle/scene_from_xml.py, tests/golden/).
 *
 * Conventions restated from Mitsuba 3's perspective sensor / look_at:
 *   dir = normalize(target-origin); left = normalize(up x dir); newup = dir x left
 *   sample_x = 0.5 - 0.5*cot(fov/2)*x_c/z_c ; sample_y = 0.5 - 0.5*(W/H)*cot(fov/2)*y_c/z_c
 *   pixel (i,j) centre = ((i+0.5)/W, (j+0.5)/H), row 0 on top; fov is horizontal.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this.  Build: make -C oracle  (gcc -O2 -ffp-contract=off -fopenmp).
 *
 * The f32 visibility arithmetic below is written out operation by operation because
 * the CUDA path must reproduce it bit for bit (DESIGN.md §3 "VA-1"); every product,
 * sum and fused multiply-add is one IEEE-754 binary32 operation, round-to-nearest-even.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_ID_FLOOR 0xFFFFFFFEu
#define ORC_ID_MISS 0xFFFFFFFFu
#define ORC_KEY_MISS 0x7F800000FFFFFFFFull

typedef struct orc_frame {
    float L[3], U[3], D[3], O[3];
    float T, Th, TW;
    float near_clip, far_clip;
    int32_t W, H;
} orc_frame;

/* TAIL constants (example_renderer.py:55-72) + BALL_SEGMENT material (:41-53). */
typedef struct orc_scene {
    int32_t has_floor;
    float floor_z, floor_min[2], floor_max[2];
    float floor_albedo;
    float light_z, light_half, radiance;
    float bounce;
} orc_scene;

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* ---- camera frame: double precision on the host, rounded once to f32 ------------- */
void orc_camera_frame(const float origin[3], const float target[3], const float up[3],
                      float fov_x_deg, float near_clip, float far_clip, int W, int H,
                      orc_frame* f)
{
    double o[3], d[3], u[3], l[3], nu[3];
    for (int k = 0; k < 3; ++k) { o[k] = origin[k]; d[k] = (double)target[k] - (double)origin[k]; u[k] = up[k]; }
    double len = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int k = 0; k < 3; ++k) d[k] = d[k] / len;
    l[0] = u[1] * d[2] - u[2] * d[1];
    l[1] = u[2] * d[0] - u[0] * d[2];
    l[2] = u[0] * d[1] - u[1] * d[0];
    len = sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
    for (int k = 0; k < 3; ++k) l[k] = l[k] / len;
    nu[0] = d[1] * l[2] - d[2] * l[1];
    nu[1] = d[2] * l[0] - d[0] * l[2];
    nu[2] = d[0] * l[1] - d[1] * l[0];
    double T = tan((double)fov_x_deg * 3.14159265358979323846 / 360.0);
    for (int k = 0; k < 3; ++k) {
        f->L[k] = (float)l[k]; f->U[k] = (float)nu[k]; f->D[k] = (float)d[k]; f->O[k] = (float)o[k];
    }
    f->T = (float)T;
    f->Th = (float)(T * (double)H / (double)W);
    f->TW = (float)(T / (double)W);
    f->near_clip = near_clip; f->far_clip = far_clip; f->W = W; f->H = H;
}

/* ---- per-pixel ray constants ------------------------------------------------------- */
static inline float pix_u(const orc_frame* f, int i) { return fmaf(-(float)(2 * i + 1), f->TW, f->T); }
static inline float pix_w(const orc_frame* f, int j) { return fmaf(-(float)(2 * j + 1), f->TW, f->Th); }

/* ---- sphere centre in camera space -------------------------------------------------- */
static inline void to_camera(const orc_frame* f, const float* p, float* c)
{
    float dx = p[0] - f->O[0], dy = p[1] - f->O[1], dz = p[2] - f->O[2];
    c[0] = fmaf(dz, f->L[2], fmaf(dy, f->L[1], dx * f->L[0]));
    c[1] = fmaf(dz, f->U[2], fmaf(dy, f->U[1], dx * f->U[0]));
    c[2] = fmaf(dz, f->D[2], fmaf(dy, f->D[1], dx * f->D[0]));
}

/* ---- the ray-sphere test.  Ray = s*(u,w,1), s = camera-space depth.  Stable form:
 * |c x v|^2 <= r^2 |v|^2  (perpendicular distance), depth = (v.c - sqrt(disc)) / |v|^2. */
static inline int sphere_depth(float cx, float cy, float cz, float r2, float u, float w,
                               float vv, float inv_vv, float near_clip, float far_clip,
                               float* depth)
{
    float a = fmaf(-cz, w, cy);
    float b = fmaf(cz, u, -cx);
    float e = fmaf(cx, w, -(cy * u));
    float m = fmaf(e, e, fmaf(b, b, a * a));
    float disc = fmaf(r2, vv, -m);
    if (!(disc >= 0.0f)) return 0;
    float vc = fmaf(cy, w, fmaf(cx, u, cz));
    float t = (vc - sqrtf(disc)) * inv_vv;
    if (!(t >= near_clip && t <= far_clip)) return 0;
    *depth = t;
    return 1;
}

/* ---- ground rectangle (TAIL, example_renderer.py:56-62) ----------------------------- */
static inline uint64_t floor_key(const orc_frame* f, const orc_scene* s, float u, float w)
{
    if (!s->has_floor) return ORC_KEY_MISS;
    float dwx = fmaf(w, f->U[0], fmaf(u, f->L[0], f->D[0]));
    float dwy = fmaf(w, f->U[1], fmaf(u, f->L[1], f->D[1]));
    float dwz = fmaf(w, f->U[2], fmaf(u, f->L[2], f->D[2]));
    float t = (s->floor_z - f->O[2]) / dwz;
    if (!(t >= f->near_clip && t <= f->far_clip)) return ORC_KEY_MISS;
    float hx = fmaf(t, dwx, f->O[0]);
    float hy = fmaf(t, dwy, f->O[1]);
    if (!(hx >= s->floor_min[0] && hx <= s->floor_max[0] && hy >= s->floor_min[1] && hy <= s->floor_max[1]))
        return ORC_KEY_MISS;
    return ((uint64_t)f2u(t) << 32) | ORC_ID_FLOOR;
}

/* ---- conservative pixel bounding box of a sphere (only ever used to SKIP work; the
 * brute-force mode below never calls it, and tests check both modes agree). ----------- */
static int sphere_bbox(const orc_frame* f, const float* c, float r, int* i0, int* i1, int* j0, int* j1)
{
    const int W = f->W, H = f->H;
    /* a non-finite centre or radius can never pass the ray test (disc is NaN or -inf, or the
     * depth is -inf); r enters the test only as r*r */
    if (!(isfinite(c[0]) && isfinite(c[1]) && isfinite(c[2]) && isfinite(r))) return 0;
    r = fabsf(r);
    if (c[2] + r < f->near_clip) return 0;
    if (c[2] - r <= 1e-6f) { *i0 = 0; *i1 = W - 1; *j0 = 0; *j1 = H - 1; return 1; }
    double cz = c[2], rr = (double)r * 1.0001 + 1e-7;
    double den = cz * cz - rr * rr;
    double sx = rr * sqrt(c[0] * (double)c[0] + den), sy = rr * sqrt(c[1] * (double)c[1] + den);
    double umin = (c[0] * cz - sx) / den, umax = (c[0] * cz + sx) / den;
    double wmin = (c[1] * cz - sy) / den, wmax = (c[1] * cz + sy) / den;
    /* u_i = T - (2i+1)TW  =>  i = (T-u)/(2TW) - 0.5 */
    double inv = 1.0 / (2.0 * (double)f->TW);
    double fi0 = ((double)f->T - umax) * inv - 0.5, fi1 = ((double)f->T - umin) * inv - 0.5;
    double fj0 = ((double)f->Th - wmax) * inv - 0.5, fj1 = ((double)f->Th - wmin) * inv - 0.5;
    double a0 = ceil(fi0 - 0.01), a1 = floor(fi1 + 0.01), b0 = ceil(fj0 - 0.01), b1 = floor(fj1 + 0.01);
    if (a0 < 0) a0 = 0;
    if (b0 < 0) b0 = 0;
    if (a1 > W - 1) a1 = W - 1;
    if (b1 > H - 1) b1 = H - 1;
    if (a0 > a1 || b0 > b1) return 0;
    *i0 = (int)a0; *i1 = (int)a1; *j0 = (int)b0; *j1 = (int)b1;
    return 1;
}

/* ---- velocity trails: a trail is the straight `linearcurve` the reference emits
 * from  position - v_hat * L  to  position  with radius 0.0007 (traj_ball_renderer.py:98-188),
 * modelled as a capsule = cylinder body + the two end spheres.  "VA-2": the test below is a
 * fixed binary32 operation sequence shared with the CUDA kernel (capsule_depth in pcr_kernels.cuh):
 * body first; a ray that enters the infinite cylinder beyond an end can only hit that end's sphere.  It is the cancellation-free form of the ray-cylinder quadratic: with
 * P = v x d, T = d . (A x v) (a scalar triple product built from the small moment components),
 * the discriminant is dd * (r^2 |P|^2 - T^2). -------------------------------------------------- */
static inline int capsule_depth(const float* A, const float* B, float r2, float u, float w, float vv, float inv_vv,
                                float near_clip, float far_clip, float* depth)
{
    float dx = B[0] - A[0], dy = B[1] - A[1], dz = B[2] - A[2];
    float dd = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    float ma = fmaf(-A[2], w, A[1]);
    float mb = fmaf(A[2], u, -A[0]);
    float me = fmaf(A[0], w, -(A[1] * u));
    float T = fmaf(dz, me, fmaf(dy, mb, dx * ma));
    float px = fmaf(w, dz, -dy);
    float py = fmaf(-u, dz, dx);
    float pz = fmaf(u, dy, -(w * dx));
    float PP = fmaf(pz, pz, fmaf(py, py, px * px));
    float disc = fmaf(r2, PP, -(T * T));
    if (!(disc >= 0.0f)) return 0;                  /* the ray misses the infinite cylinder, hence the capsule */
    if (PP > 0.0f) {
        float va = fmaf(A[1], w, fmaf(A[0], u, A[2]));
        float vd = fmaf(dy, w, fmaf(dx, u, dz));
        float da = fmaf(dz, A[2], fmaf(dy, A[1], dx * A[0]));
        float PQ = fmaf(va, dd, -(vd * da));
        float s = (PQ - sqrtf(dd * disc)) / PP;
        float y = fmaf(s, vd, -da);
        if (y >= 0.0f && y <= dd) {                 /* enters through the body: that is the nearest hit */
            if (!(s >= near_clip && s <= far_clip)) return 0;
            *depth = s;
            return 1;
        }
        /* enters the cylinder beyond one end: only that end's sphere can be hit first */
        const float* C = (y < 0.0f) ? A : B;
        return sphere_depth(C[0], C[1], C[2], r2, u, w, vv, inv_vv, near_clip, far_clip, depth);
    }
    /* ray parallel to the axis: nearest of the two end spheres */
    float ta, tb;
    int ha = sphere_depth(A[0], A[1], A[2], r2, u, w, vv, inv_vv, near_clip, far_clip, &ta);
    int hb = sphere_depth(B[0], B[1], B[2], r2, u, w, vv, inv_vv, near_clip, far_clip, &tb);
    if (ha && (!hb || ta <= tb)) { *depth = ta; return 1; }
    if (hb) { *depth = tb; return 1; }
    return 0;
}

/* conservative pixel bbox of a capsule: hull of the two end spheres' (unclamped) boxes */
static int capsule_bbox(const orc_frame* f, const float* ca, const float* cb, float r, int* i0, int* i1, int* j0, int* j1)
{
    const int W = f->W, H = f->H;
    if (!(isfinite(ca[0]) && isfinite(ca[1]) && isfinite(ca[2]) && isfinite(cb[0]) && isfinite(cb[1]) && isfinite(cb[2]) && isfinite(r))) return 0;
    r = fabsf(r);
    if (ca[2] + r < f->near_clip && cb[2] + r < f->near_clip) return 0;
    if (ca[2] - r <= 1e-3f || cb[2] - r <= 1e-3f) { *i0 = 0; *i1 = W - 1; *j0 = 0; *j1 = H - 1; return 1; }
    double lo_i = 1e30, hi_i = -1e30, lo_j = 1e30, hi_j = -1e30;
    for (int e = 0; e < 2; ++e) {
        const float* c = e ? cb : ca;
        double cz = c[2], rr = (double)r * 1.0001 + 1e-7;
        double den = cz * cz - rr * rr;
        double sx = rr * sqrt(c[0] * (double)c[0] + den), sy = rr * sqrt(c[1] * (double)c[1] + den);
        double umin = (c[0] * cz - sx) / den, umax = (c[0] * cz + sx) / den;
        double wmin = (c[1] * cz - sy) / den, wmax = (c[1] * cz + sy) / den;
        double inv = 1.0 / (2.0 * (double)f->TW);
        double fi0 = ((double)f->T - umax) * inv - 0.5, fi1 = ((double)f->T - umin) * inv - 0.5;
        double fj0 = ((double)f->Th - wmax) * inv - 0.5, fj1 = ((double)f->Th - wmin) * inv - 0.5;
        if (fi0 < lo_i) lo_i = fi0;
        if (fi1 > hi_i) hi_i = fi1;
        if (fj0 < lo_j) lo_j = fj0;
        if (fj1 > hi_j) hi_j = fj1;
    }
    double a0 = ceil(lo_i - 0.01), a1 = floor(hi_i + 0.01), b0 = ceil(lo_j - 0.01), b1 = floor(hi_j + 0.01);
    if (a0 < 0) a0 = 0;
    if (b0 < 0) b0 = 0;
    if (a1 > W - 1) a1 = W - 1;
    if (b1 > H - 1) b1 = H - 1;
    if (!(a0 <= a1 && b0 <= b1)) return 0;
    *i0 = (int)a0; *i1 = (int)a1; *j0 = (int)b0; *j1 = (int)b1;
    return 1;
}

/*
 * Visibility buffer.  pos4 = n x (x,y,z,r) world-space spheres (what BALL_SEGMENT emits);
 * vis = H*W keys, row 0 on top.  mode 0: brute force, every pixel against every sphere
 * (the definition).  mode 1: each sphere against the pixels of its conservative bbox,
 * threads own row bands (min is order independent, so the result is the same).
 */
void orc_visibility(const float* pos4, int64_t n, uint32_t id_base, const orc_frame* f,
                    const orc_scene* s, uint64_t* vis, int mode)
{
    const int W = f->W, H = f->H;
    float* cam = (float*)malloc((size_t)(n > 0 ? n : 1) * 5 * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) {
        float c[3];
        to_camera(f, pos4 + 4 * k, c);
        float r = pos4[4 * k + 3];
        cam[5 * k + 0] = c[0]; cam[5 * k + 1] = c[1]; cam[5 * k + 2] = c[2];
        cam[5 * k + 3] = r * r; cam[5 * k + 4] = r;
    }
    if (mode == 0) {
#pragma omp parallel for schedule(dynamic, 4)
        for (int j = 0; j < H; ++j) {
            float w = pix_w(f, j);
            for (int i = 0; i < W; ++i) {
                float u = pix_u(f, i);
                float vv = fmaf(u, u, fmaf(w, w, 1.0f));
                float inv_vv = 1.0f / vv;
                uint64_t best = floor_key(f, s, u, w);
                for (int64_t k = 0; k < n; ++k) {
                    float t;
                    const float* c = cam + 5 * k;
                    if (sphere_depth(c[0], c[1], c[2], c[3], u, w, vv, inv_vv, f->near_clip, f->far_clip, &t)) {
                        uint64_t key = ((uint64_t)f2u(t) << 32) | (uint64_t)(id_base + (uint32_t)k);
                        if (key < best) best = key;
                    }
                }
                vis[(size_t)j * W + i] = best;
            }
        }
    } else {
        int* box = (int*)malloc((size_t)(n > 0 ? n : 1) * 4 * sizeof(int));
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < n; ++k) {
            int* b = box + 4 * k;
            if (!sphere_bbox(f, cam + 5 * k, cam[5 * k + 4], b, b + 1, b + 2, b + 3)) { b[0] = 1; b[1] = 0; b[2] = 1; b[3] = 0; }
        }
        const int band = 8;
        const int nb = (H + band - 1) / band;
#pragma omp parallel for schedule(dynamic, 1)
        for (int bi = 0; bi < nb; ++bi) {
            int r0 = bi * band, r1 = r0 + band - 1;
            if (r1 > H - 1) r1 = H - 1;
            for (int j = r0; j <= r1; ++j) {
                float w = pix_w(f, j);
                for (int i = 0; i < W; ++i) vis[(size_t)j * W + i] = floor_key(f, s, pix_u(f, i), w);
            }
            for (int64_t k = 0; k < n; ++k) {
                const int* b = box + 4 * k;
                int j0 = b[2] > r0 ? b[2] : r0, j1 = b[3] < r1 ? b[3] : r1;
                if (j0 > j1 || b[0] > b[1]) continue;
                const float* c = cam + 5 * k;
                for (int j = j0; j <= j1; ++j) {
                    float w = pix_w(f, j);
                    for (int i = b[0]; i <= b[1]; ++i) {
                        float u = pix_u(f, i);
                        float vv = fmaf(u, u, fmaf(w, w, 1.0f));
                        float a = fmaf(-c[2], w, c[1]);
                        float bb = fmaf(c[2], u, -c[0]);
                        float e = fmaf(c[0], w, -(c[1] * u));
                        float m = fmaf(e, e, fmaf(bb, bb, a * a));
                        float disc = fmaf(c[3], vv, -m);
                        if (!(disc >= 0.0f)) continue;
                        float inv_vv = 1.0f / vv;
                        float vc = fmaf(c[1], w, fmaf(c[0], u, c[2]));
                        float t = (vc - sqrtf(disc)) * inv_vv;
                        if (!(t >= f->near_clip && t <= f->far_clip)) continue;
                        uint64_t key = ((uint64_t)f2u(t) << 32) | (uint64_t)(id_base + (uint32_t)k);
                        uint64_t* dst = vis + (size_t)j * W + i;
                        if (key < *dst) *dst = key;
                    }
                }
            }
        }
        free(box);
    }
    free(cam);
}

/*
 * Merge m capsules (trails) into an existing visibility buffer: cap_a4 = m x (ax,ay,az,r),
 * cap_b4 = m x (bx,by,bz,valid) in WORLD space; capsule j gets id cap_id_base + j.  mode 0: every
 * pixel against every capsule; mode 1: bbox-accelerated, threads own row bands.
 */
void orc_visibility_caps(const float* cap_a4, const float* cap_b4, int64_t m, uint32_t cap_id_base, const orc_frame* f,
                         uint64_t* vis, int mode)
{
    const int W = f->W, H = f->H;
    float* cam = (float*)malloc((size_t)(m > 0 ? m : 1) * 8 * sizeof(float));
    int* box = (int*)malloc((size_t)(m > 0 ? m : 1) * 4 * sizeof(int));
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < m; ++k) {
        float* c = cam + 8 * k;
        int* b = box + 4 * k;
        to_camera(f, cap_a4 + 4 * k, c);
        to_camera(f, cap_b4 + 4 * k, c + 4);
        float r = cap_a4[4 * k + 3];
        c[3] = r * r; c[7] = r;
        int ok = cap_b4[4 * k + 3] != 0.0f;
        if (ok && mode == 1) ok = capsule_bbox(f, c, c + 4, r, b, b + 1, b + 2, b + 3);
        else if (ok) { b[0] = 0; b[1] = W - 1; b[2] = 0; b[3] = H - 1; }
        if (!ok) { b[0] = 1; b[1] = 0; b[2] = 1; b[3] = 0; }
    }
    const int band = 8;
    const int nb = (H + band - 1) / band;
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < nb; ++bi) {
        int r0 = bi * band, r1 = r0 + band - 1;
        if (r1 > H - 1) r1 = H - 1;
        for (int64_t k = 0; k < m; ++k) {
            const int* b = box + 4 * k;
            int j0 = b[2] > r0 ? b[2] : r0, j1 = b[3] < r1 ? b[3] : r1;
            if (j0 > j1 || b[0] > b[1]) continue;
            const float* c = cam + 8 * k;
            for (int j = j0; j <= j1; ++j) {
                float w = pix_w(f, j);
                for (int i = b[0]; i <= b[1]; ++i) {
                    float u = pix_u(f, i);
                    float vv = fmaf(u, u, fmaf(w, w, 1.0f));
                    float inv_vv = 1.0f / vv;
                    float t;
                    if (!capsule_depth(c, c + 4, c[3], u, w, vv, inv_vv, f->near_clip, f->far_clip, &t)) continue;
                    uint64_t key = ((uint64_t)f2u(t) << 32) | (uint64_t)(cap_id_base + (uint32_t)k);
                    uint64_t* dst = vis + (size_t)j * W + i;
                    if (key < *dst) *dst = key;
                }
            }
        }
    }
    free(box);
    free(cam);
}

/* ---- shading: the stated look model, evaluated in double -------------- */

/* Form factor (cosine-weighted solid angle / pi) of the square emitter |x|,|y|<=a at
 * z = lz (TAIL light, example_renderer.py:64-72) seen from p with unit normal nrm:
 * Lambert's polygon formula after clipping the polygon to the horizon of nrm. */
static double rect_form_factor(const double p[3], const double nrm[3], double a, double lz)
{
    double v[8][3], q[8][3];
    const double cx[4] = { -a, a, a, -a }, cy[4] = { -a, -a, a, a };
    int nv = 4;
    for (int k = 0; k < 4; ++k) { v[k][0] = cx[k] - p[0]; v[k][1] = cy[k] - p[1]; v[k][2] = lz - p[2]; }
    /* clip against nrm . x >= 0 */
    int nq = 0;
    for (int k = 0; k < nv; ++k) {
        const double* A = v[k];
        const double* B = v[(k + 1) % nv];
        double da = A[0] * nrm[0] + A[1] * nrm[1] + A[2] * nrm[2];
        double db = B[0] * nrm[0] + B[1] * nrm[1] + B[2] * nrm[2];
        if (da >= 0.0) { q[nq][0] = A[0]; q[nq][1] = A[1]; q[nq][2] = A[2]; ++nq; }
        if ((da >= 0.0) != (db >= 0.0)) {
            double t = da / (da - db);
            q[nq][0] = A[0] + t * (B[0] - A[0]); q[nq][1] = A[1] + t * (B[1] - A[1]); q[nq][2] = A[2] + t * (B[2] - A[2]);
            ++nq;
        }
    }
    if (nq < 3) return 0.0;
    for (int k = 0; k < nq; ++k) {
        double l = sqrt(q[k][0] * q[k][0] + q[k][1] * q[k][1] + q[k][2] * q[k][2]);
        if (l < 1e-30) return 0.0;
        q[k][0] /= l; q[k][1] /= l; q[k][2] /= l;
    }
    double sum = 0.0;
    for (int k = 0; k < nq; ++k) {
        const double* A = q[k];
        const double* B = q[(k + 1) % nq];
        double c[3] = { A[1] * B[2] - A[2] * B[1], A[2] * B[0] - A[0] * B[2], A[0] * B[1] - A[1] * B[0] };
        double cl = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        if (cl < 1e-12) continue;
        double d = A[0] * B[0] + A[1] * B[1] + A[2] * B[2];
        if (d > 1.0) d = 1.0;
        if (d < -1.0) d = -1.0;
        sum += acos(d) * (c[0] * nrm[0] + c[1] * nrm[1] + c[2] * nrm[2]) / cl;
    }
    return fabs(sum) / (2.0 * 3.14159265358979323846);
}

static inline uint8_t srgb8(double c)
{
    double s = c <= 0.0031308 ? 12.92 * c : 1.055 * pow(c, 1.0 / 2.4) - 0.055;
    if (!(s > 0.0)) s = 0.0;
    if (s > 1.0) s = 1.0;
    return (uint8_t)(int)(s * 255.0 + 0.5);
}

/*
 * rgba[H][W][4] from a visibility buffer.  owner_only mirrors pcr_shade: non-local sphere
 * ids give 0,0,0,0; floor/miss pixels are written only when id_base == 0.
 */
void orc_shade(const uint64_t* vis, const float* pos4, const float* attr4, int64_t n,
               uint32_t id_base, int owner_only, const orc_frame* f, const orc_scene* s,
               uint8_t* rgba)
{
    const int W = f->W, H = f->H;
    const double up[3] = { 0.0, 0.0, 1.0 };
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = 0; j < H; ++j) {
        float w = pix_w(f, j);
        for (int i = 0; i < W; ++i) {
            uint8_t* px = rgba + ((size_t)j * W + i) * 4;
            uint64_t key = vis[(size_t)j * W + i];
            uint32_t id = (uint32_t)(key & 0xFFFFFFFFu);
            float t = u2f((uint32_t)(key >> 32));
            double rgb[3] = { 0.0, 0.0, 0.0 };
            int write = 1;
            if (id == ORC_ID_MISS) {
                if (owner_only && id_base != 0) write = 0;
            } else {
                float u = pix_u(f, i);
                float dwx = fmaf(w, f->U[0], fmaf(u, f->L[0], f->D[0]));
                float dwy = fmaf(w, f->U[1], fmaf(u, f->L[1], f->D[1]));
                float dwz = fmaf(w, f->U[2], fmaf(u, f->L[2], f->D[2]));
                double P[3] = { (double)fmaf(t, dwx, f->O[0]), (double)fmaf(t, dwy, f->O[1]), (double)fmaf(t, dwz, f->O[2]) };
                if (id == ORC_ID_FLOOR) {
                    if (owner_only && id_base != 0) write = 0;
                    else if (f->O[2] > s->floor_z) {
                        double L = (double)s->floor_albedo * (double)s->radiance * rect_form_factor(P, up, s->light_half, s->light_z);
                        rgb[0] = rgb[1] = rgb[2] = L;
                    }
                } else {
                    int64_t k = (int64_t)id - (int64_t)id_base;
                    if (k < 0 || k >= n) write = owner_only ? 0 : 1;
                    else {
                        const float* c = pos4 + 4 * k;
                        double nr[3] = { P[0] - (double)c[0], P[1] - (double)c[1], P[2] - (double)c[2] };
                        double l = sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
                        if (l > 0.0) { nr[0] /= l; nr[1] /= l; nr[2] /= l; } else { nr[2] = 1.0; }
                        double Ld = (double)s->radiance * rect_form_factor(P, nr, s->light_half, s->light_z);
                        double Li = 0.0;
                        if (s->has_floor) {
                            double Pf[3] = { P[0], P[1], (double)s->floor_z };
                            double B = (double)s->floor_albedo * (double)s->radiance * rect_form_factor(Pf, up, s->light_half, s->light_z);
                            Li = (double)s->bounce * B * 0.5 * (1.0 - nr[2]);
                        }
                        for (int ch = 0; ch < 3; ++ch) rgb[ch] = (double)attr4[4 * k + ch] * (Ld + Li);
                    }
                }
            }
            if (write) { px[0] = srgb8(rgb[0]); px[1] = srgb8(rgb[1]); px[2] = srgb8(rgb[2]); px[3] = 255; }
            else { px[0] = px[1] = px[2] = px[3] = 0; }
        }
    }
}

/*
 * Shade the pixels won by capsules (ids cap_id_base .. cap_id_base+m-1) on top of an image produced by
 * orc_shade: diffuse trail colour, normal = from the nearest point of the capsule axis to the hit.
 */
void orc_shade_caps(const uint64_t* vis, const float* cap_a4, const float* cap_b4, int64_t m, uint32_t cap_id_base,
                    const float trail_rgb[3], const orc_frame* f, const orc_scene* s, uint8_t* rgba)
{
    const int W = f->W, H = f->H;
    const double up[3] = { 0.0, 0.0, 1.0 };
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = 0; j < H; ++j) {
        float w = pix_w(f, j);
        for (int i = 0; i < W; ++i) {
            uint64_t key = vis[(size_t)j * W + i];
            uint32_t id = (uint32_t)(key & 0xFFFFFFFFu);
            if (id < cap_id_base || (int64_t)(id - cap_id_base) >= m || id >= ORC_ID_FLOOR) continue;
            int64_t k = id - cap_id_base;
            float t = u2f((uint32_t)(key >> 32));
            float u = pix_u(f, i);
            float dwx = fmaf(w, f->U[0], fmaf(u, f->L[0], f->D[0]));
            float dwy = fmaf(w, f->U[1], fmaf(u, f->L[1], f->D[1]));
            float dwz = fmaf(w, f->U[2], fmaf(u, f->L[2], f->D[2]));
            double P[3] = { (double)fmaf(t, dwx, f->O[0]), (double)fmaf(t, dwy, f->O[1]), (double)fmaf(t, dwz, f->O[2]) };
            const float* A = cap_a4 + 4 * k;
            const float* B = cap_b4 + 4 * k;
            double d[3] = { (double)B[0] - A[0], (double)B[1] - A[1], (double)B[2] - A[2] };
            double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            double h = dd > 0.0 ? ((P[0] - A[0]) * d[0] + (P[1] - A[1]) * d[1] + (P[2] - A[2]) * d[2]) / dd : 0.0;
            if (h < 0.0) h = 0.0;
            if (h > 1.0) h = 1.0;
            double nr[3] = { P[0] - (A[0] + h * d[0]), P[1] - (A[1] + h * d[1]), P[2] - (A[2] + h * d[2]) };
            double l = sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
            if (l > 0.0) { nr[0] /= l; nr[1] /= l; nr[2] /= l; } else { nr[2] = 1.0; }
            double Ld = (double)s->radiance * rect_form_factor(P, nr, s->light_half, s->light_z);
            double Li = 0.0;
            if (s->has_floor) {
                double Pf[3] = { P[0], P[1], (double)s->floor_z };
                double Bn = (double)s->floor_albedo * (double)s->radiance * rect_form_factor(Pf, up, s->light_half, s->light_z);
                Li = (double)s->bounce * Bn * 0.5 * (1.0 - nr[2]);
            }
            uint8_t* px = rgba + ((size_t)j * W + i) * 4;
            for (int ch = 0; ch < 3; ++ch) px[ch] = srgb8((double)trail_rgb[ch] * (Ld + Li));
            px[3] = 255;
        }
    }
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm of bench.py asks for all cores */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void)
{
    int n = 1;
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    n = omp_get_max_threads();
#endif
    return n;
}
