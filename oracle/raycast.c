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

/* =========================================================================================
 * Droplet scene of traj_renderer.py / traj_vel_renderer.py: every point is an
 * instance of the droplet OBJ mesh (_create_droplet_mesh, traj_renderer.py:102-153) placed by the
 * 4x4 matrix of DROPLET_SEGMENT (:45-55), plus one `linearcurve` polyline per point
 * (_add_trail_lines :204-396 or _add_velocity_trail, traj_vel_renderer.py:194-288).
 *
 * Arithmetic contract "VA-3" (ray-triangle), PARITY UNPINNED against Mitsuba like VA-1:
 *   world vertex   X_k = fma(M[k][0], vx, fma(M[k][1], vy, fma(M[k][2], vz, M[k][3])))   (M, v binary32)
 *   camera vertex  to_camera(X)                                   (the sphere-centre transform)
 *   ray (u,w,1):   e1 = v1-v0, e2 = v2-v0, p = d x e2, det = e1.p, s = -v0, q = s x e1
 *                  bu = (s.p)/det, bv = (d.q)/det, depth = (e2.q)/det, with 1/det one division;
 *                  hit iff det != 0, bu >= 0, bv >= 0, bu + bv <= 1, near <= depth <= far.
 * Every line below is one IEEE binary32 operation; pcr_kernels.cuh:triangle_depth repeats them.
 * ========================================================================================= */
static inline int triangle_depth(const float* v0, const float* v1, const float* v2, float u, float w,
                                 float near_clip, float far_clip, float* depth)
{
    const float e1x = v1[0] - v0[0], e1y = v1[1] - v0[1], e1z = v1[2] - v0[2];
    const float e2x = v2[0] - v0[0], e2y = v2[1] - v0[1], e2z = v2[2] - v0[2];
    const float px = fmaf(w, e2z, -e2y);
    const float py = fmaf(-u, e2z, e2x);
    const float pz = fmaf(u, e2y, -(w * e2x));
    const float det = fmaf(e1z, pz, fmaf(e1y, py, e1x * px));
    if (!(det != 0.0f)) return 0;
    const float inv = 1.0f / det;
    const float sx = -v0[0], sy = -v0[1], sz = -v0[2];
    const float bu = fmaf(sz, pz, fmaf(sy, py, sx * px)) * inv;
    if (!(bu >= 0.0f && bu <= 1.0f)) return 0;
    const float qx = fmaf(sy, e1z, -(sz * e1y));
    const float qy = fmaf(sz, e1x, -(sx * e1z));
    const float qz = fmaf(sx, e1y, -(sy * e1x));
    const float bv = fmaf(w, qy, fmaf(u, qx, qz)) * inv;
    if (!(bv >= 0.0f && bu + bv <= 1.0f)) return 0;
    const float t = fmaf(e2z, qz, fmaf(e2y, qy, e2x * qx)) * inv;
    if (!(t >= near_clip && t <= far_clip)) return 0;
    *depth = t;
    return 1;
}

/*
 * Merge n mesh instances into an existing visibility buffer.  verts = nv x 3 (object space, as a
 * loader reads the OBJ text), faces = nf x 3 zero-based, xf = n x 12 rows [R | t] (binary32).
 * Instance k gets id id_base + k.  mode 0: every pixel against every triangle of every instance;
 * mode 1: pixels inside the (padded) screen bbox of the instance's projected vertices only.
 */
void orc_visibility_mesh(const float* verts, int32_t nv, const int32_t* faces, int32_t nf, const float* xf, int64_t n,
                         uint32_t id_base, const orc_frame* f, uint64_t* vis, int mode)
{
    const int W = f->W, H = f->H;
    float* cam = (float*)malloc((size_t)(n > 0 ? n : 1) * (size_t)nv * 3 * sizeof(float));
    int* box = (int*)malloc((size_t)(n > 0 ? n : 1) * 4 * sizeof(int));
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) {
        const float* M = xf + 12 * k;
        float* cv = cam + (size_t)k * nv * 3;
        int* b = box + 4 * k;
        float lo_i = 1e30f, hi_i = -1e30f, lo_j = 1e30f, hi_j = -1e30f, zmin = 1e30f, zmax = -1e30f;
        int finite = 1;
        for (int v = 0; v < nv; ++v) {
            const float* o = verts + 3 * v;
            float X[3];
            for (int r = 0; r < 3; ++r) X[r] = fmaf(M[4 * r + 0], o[0], fmaf(M[4 * r + 1], o[1], fmaf(M[4 * r + 2], o[2], M[4 * r + 3])));
            to_camera(f, X, cv + 3 * v);
            const float* c = cv + 3 * v;
            if (!(isfinite(c[0]) && isfinite(c[1]) && isfinite(c[2]))) finite = 0;
            if (c[2] < zmin) zmin = c[2];
            if (c[2] > zmax) zmax = c[2];
            if (c[2] > 1e-3f) {
                float fi = (f->T - c[0] / c[2]) / (2.0f * f->TW) - 0.5f, fj = (f->Th - c[1] / c[2]) / (2.0f * f->TW) - 0.5f;
                if (fi < lo_i) lo_i = fi;
                if (fi > hi_i) hi_i = fi;
                if (fj < lo_j) lo_j = fj;
                if (fj > hi_j) hi_j = fj;
            }
        }
        if (!finite || zmax < f->near_clip) { b[0] = 1; b[1] = 0; b[2] = 1; b[3] = 0; }
        else if (mode == 0 || !(zmin > 1e-3f)) { b[0] = 0; b[1] = W - 1; b[2] = 0; b[3] = H - 1; }
        else {
            float a0 = ceilf(lo_i - 1.0f), a1 = floorf(hi_i + 1.0f), b0 = ceilf(lo_j - 1.0f), b1 = floorf(hi_j + 1.0f);
            if (a0 < 0.0f) a0 = 0.0f;
            if (b0 < 0.0f) b0 = 0.0f;
            if (a1 > (float)(W - 1)) a1 = (float)(W - 1);
            if (b1 > (float)(H - 1)) b1 = (float)(H - 1);
            if (!(a0 <= a1 && b0 <= b1)) { b[0] = 1; b[1] = 0; b[2] = 1; b[3] = 0; }
            else { b[0] = (int)a0; b[1] = (int)a1; b[2] = (int)b0; b[3] = (int)b1; }
        }
    }
    const int band = 8;
    const int nb = (H + band - 1) / band;
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < nb; ++bi) {
        int r0 = bi * band, r1 = r0 + band - 1;
        if (r1 > H - 1) r1 = H - 1;
        for (int64_t k = 0; k < n; ++k) {
            const int* b = box + 4 * k;
            int j0 = b[2] > r0 ? b[2] : r0, j1 = b[3] < r1 ? b[3] : r1;
            if (j0 > j1 || b[0] > b[1]) continue;
            const float* cv = cam + (size_t)k * nv * 3;
            for (int j = j0; j <= j1; ++j) {
                float w = pix_w(f, j);
                for (int i = b[0]; i <= b[1]; ++i) {
                    float u = pix_u(f, i);
                    uint64_t* dst = vis + (size_t)j * W + i;
                    for (int t = 0; t < nf; ++t) {
                        const int32_t* fc = faces + 3 * t;
                        float d;
                        if (!triangle_depth(cv + 3 * fc[0], cv + 3 * fc[1], cv + 3 * fc[2], u, w, f->near_clip, f->far_clip, &d)) continue;
                        uint64_t key = ((uint64_t)f2u(d) << 32) | (uint64_t)(id_base + (uint32_t)k);
                        if (key < *dst) *dst = key;
                    }
                }
            }
        }
    }
    free(box);
    free(cam);
}

/*
 * Merge n polylines (curve files) into an existing visibility buffer: ctrl = n x max_ctrl x 3 control
 * points (world space), count[k] of them valid (0 or >= 2); every pair of consecutive control points
 * is one capsule of the given radius (VA-2); all segments of polyline k share id cap_id_base + k.
 */
void orc_visibility_polylines(const float* ctrl, const int32_t* count, int64_t n, int32_t max_ctrl, float radius,
                              uint32_t cap_id_base, const orc_frame* f, uint64_t* vis, int mode)
{
    const int W = f->W, H = f->H;
    const int ms = max_ctrl - 1;
    float* cam = (float*)malloc((size_t)(n > 0 ? n : 1) * max_ctrl * 3 * sizeof(float));
    int* box = (int*)malloc((size_t)(n > 0 ? n : 1) * ms * 4 * sizeof(int));
    const float r2 = radius * radius;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) {
        float* c = cam + (size_t)k * max_ctrl * 3;
        for (int p = 0; p < count[k]; ++p) to_camera(f, ctrl + ((size_t)k * max_ctrl + p) * 3, c + 3 * p);
        for (int sg = 0; sg < ms; ++sg) {
            int* b = box + ((size_t)k * ms + sg) * 4;
            int ok = sg + 1 < count[k];
            if (ok && mode == 1) ok = capsule_bbox(f, c + 3 * sg, c + 3 * sg + 3, radius, b, b + 1, b + 2, b + 3);
            else if (ok) { b[0] = 0; b[1] = W - 1; b[2] = 0; b[3] = H - 1; }
            if (!ok) { b[0] = 1; b[1] = 0; b[2] = 1; b[3] = 0; }
        }
    }
    const int band = 8;
    const int nb = (H + band - 1) / band;
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < nb; ++bi) {
        int r0 = bi * band, r1 = r0 + band - 1;
        if (r1 > H - 1) r1 = H - 1;
        for (int64_t k = 0; k < n; ++k) {
            for (int sg = 0; sg + 1 < count[k]; ++sg) {
                const int* b = box + ((size_t)k * ms + sg) * 4;
                int j0 = b[2] > r0 ? b[2] : r0, j1 = b[3] < r1 ? b[3] : r1;
                if (j0 > j1 || b[0] > b[1]) continue;
                const float* A = cam + ((size_t)k * max_ctrl + sg) * 3;
                for (int j = j0; j <= j1; ++j) {
                    float w = pix_w(f, j);
                    for (int i = b[0]; i <= b[1]; ++i) {
                        float u = pix_u(f, i);
                        float vv = fmaf(u, u, fmaf(w, w, 1.0f));
                        float inv_vv = 1.0f / vv;
                        float t;
                        if (!capsule_depth(A, A + 3, r2, u, w, vv, inv_vv, f->near_clip, f->far_clip, &t)) continue;
                        uint64_t key = ((uint64_t)f2u(t) << 32) | (uint64_t)(cap_id_base + (uint32_t)k);
                        uint64_t* dst = vis + (size_t)j * W + i;
                        if (key < *dst) *dst = key;
                    }
                }
            }
        }
    }
    free(box);
    free(cam);
}

/* Lit radiance of a diffuse surface point with unit normal nr (the stated look model). */
static double lit(const double P[3], const double nr[3], const orc_scene* s)
{
    const double up[3] = { 0.0, 0.0, 1.0 };
    double Ld = (double)s->radiance * rect_form_factor(P, nr, s->light_half, s->light_z);
    double Li = 0.0;
    if (s->has_floor) {
        double Pf[3] = { P[0], P[1], (double)s->floor_z };
        double Bn = (double)s->floor_albedo * (double)s->radiance * rect_form_factor(Pf, up, s->light_half, s->light_z);
        Li = (double)s->bounce * Bn * 0.5 * (1.0 - nr[2]);
    }
    return Ld + Li;
}

static inline void hit_point(const orc_frame* f, int i, int j, float t, double P[3])
{
    float u = pix_u(f, i), w = pix_w(f, j);
    float dwx = fmaf(w, f->U[0], fmaf(u, f->L[0], f->D[0]));
    float dwy = fmaf(w, f->U[1], fmaf(u, f->L[1], f->D[1]));
    float dwz = fmaf(w, f->U[2], fmaf(u, f->L[2], f->D[2]));
    P[0] = (double)fmaf(t, dwx, f->O[0]); P[1] = (double)fmaf(t, dwy, f->O[1]); P[2] = (double)fmaf(t, dwz, f->O[2]);
}

/* Overlay the pixels won by polylines: normal = from the nearest point of the nearest segment's axis. */
void orc_shade_polylines(const uint64_t* vis, const float* ctrl, const int32_t* count, int64_t n, int32_t max_ctrl,
                         uint32_t cap_id_base, const float trail_rgb[3], const orc_frame* f, const orc_scene* s, uint8_t* rgba)
{
    const int W = f->W, H = f->H;
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = 0; j < H; ++j) {
        for (int i = 0; i < W; ++i) {
            uint64_t key = vis[(size_t)j * W + i];
            uint32_t id = (uint32_t)(key & 0xFFFFFFFFu);
            if (id < cap_id_base || (int64_t)(id - cap_id_base) >= n || id >= ORC_ID_FLOOR) continue;
            int64_t k = id - cap_id_base;
            double P[3];
            hit_point(f, i, j, u2f((uint32_t)(key >> 32)), P);
            double best = 1e300, nr[3] = { 0.0, 0.0, 1.0 };
            for (int sg = 0; sg + 1 < count[k]; ++sg) {
                const float* A = ctrl + ((size_t)k * max_ctrl + sg) * 3;
                const float* B = A + 3;
                double d[3] = { (double)B[0] - A[0], (double)B[1] - A[1], (double)B[2] - A[2] };
                double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                double h = dd > 0.0 ? ((P[0] - A[0]) * d[0] + (P[1] - A[1]) * d[1] + (P[2] - A[2]) * d[2]) / dd : 0.0;
                if (h < 0.0) h = 0.0;
                if (h > 1.0) h = 1.0;
                double q[3] = { P[0] - (A[0] + h * d[0]), P[1] - (A[1] + h * d[1]), P[2] - (A[2] + h * d[2]) };
                double l2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
                if (l2 < best) { best = l2; nr[0] = q[0]; nr[1] = q[1]; nr[2] = q[2]; }
            }
            double l = sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
            if (l > 0.0) { nr[0] /= l; nr[1] /= l; nr[2] /= l; } else { nr[0] = 0.0; nr[1] = 0.0; nr[2] = 1.0; }
            double Lo = lit(P, nr, s);
            uint8_t* px = rgba + ((size_t)j * W + i) * 4;
            for (int ch = 0; ch < 3; ++ch) px[ch] = srgb8((double)trail_rgb[ch] * Lo);
            px[3] = 255;
        }
    }
}

/*
 * Overlay the pixels won by droplet instances (ids id_base .. id_base+n-1).  Shading normal = the smooth
 * normal of the surface of revolution the mesh samples: prof = (n_rings+1) x 4 rows (ring radius, ring z,
 * ring normal radial component, ring normal z component) in object space, ring z strictly decreasing; the
 * hit point is taken to object space, the ring normals either side of its z are interpolated linearly and
 * swung around the axis to the hit's azimuth.
 */
void orc_shade_droplets(const uint64_t* vis, const float* xf, int64_t n, uint32_t id_base, const float* prof, int32_t n_rings,
                        const float rgb[3], const orc_frame* f, const orc_scene* s, uint8_t* rgba)
{
    const int W = f->W, H = f->H;
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = 0; j < H; ++j) {
        for (int i = 0; i < W; ++i) {
            uint64_t key = vis[(size_t)j * W + i];
            uint32_t id = (uint32_t)(key & 0xFFFFFFFFu);
            if (id < id_base || (int64_t)(id - id_base) >= n || id >= ORC_ID_FLOOR) continue;
            const float* M = xf + 12 * (int64_t)(id - id_base);
            double P[3];
            hit_point(f, i, j, u2f((uint32_t)(key >> 32)), P);
            double d[3] = { P[0] - (double)M[3], P[1] - (double)M[7], P[2] - (double)M[11] };
            double q[3];
            for (int c = 0; c < 3; ++c) q[c] = (double)M[c] * d[0] + (double)M[4 + c] * d[1] + (double)M[8 + c] * d[2];   /* R^T d */
            int b = 0;
            while (b + 1 < n_rings && q[2] < (double)prof[4 * (b + 1) + 1]) ++b;
            double z0 = prof[4 * b + 1], z1 = prof[4 * (b + 1) + 1];
            double fr = z0 > z1 ? (z0 - q[2]) / (z0 - z1) : 0.0;
            if (fr < 0.0) fr = 0.0;
            if (fr > 1.0) fr = 1.0;
            double nrr = (double)prof[4 * b + 2] + fr * ((double)prof[4 * (b + 1) + 2] - (double)prof[4 * b + 2]);
            double nzz = (double)prof[4 * b + 3] + fr * ((double)prof[4 * (b + 1) + 3] - (double)prof[4 * b + 3]);
            double l = sqrt(nrr * nrr + nzz * nzz);
            if (l > 0.0) { nrr /= l; nzz /= l; } else { nrr = 1.0; nzz = 0.0; }
            double rho = sqrt(q[0] * q[0] + q[1] * q[1]);
            double cx = rho > 1e-12 ? q[0] / rho : 1.0, cy = rho > 1e-12 ? q[1] / rho : 0.0;
            double no[3] = { nrr * cx, nrr * cy, nzz }, nr[3];
            for (int c = 0; c < 3; ++c) nr[c] = (double)M[4 * c] * no[0] + (double)M[4 * c + 1] * no[1] + (double)M[4 * c + 2] * no[2];
            l = sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
            if (l > 0.0) { nr[0] /= l; nr[1] /= l; nr[2] /= l; } else { nr[0] = 0.0; nr[1] = 0.0; nr[2] = 1.0; }
            double Lo = lit(P, nr, s);
            uint8_t* px = rgba + ((size_t)j * W + i) * 4;
            for (int ch = 0; ch < 3; ++ch) px[ch] = srgb8((double)rgb[ch] * Lo);
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
