// pcr_droplets.cuh — droplet scene of traj_renderer.py / traj_vel_renderer.py:
// every point is an instance of the droplet OBJ mesh (_create_droplet_mesh, traj_renderer.py:102-153)
// placed by the Rodrigues matrix of generate_rotation_matrix_from_velocity (:159-202), plus one
// `linearcurve` polyline per point: the Catmull-Rom history trail of _add_trail_lines (:204-396) or the
// straight velocity trail of traj_vel_renderer.py:194-288.
//
// What the reference does per point in python (600 us/point with trails, one temp file per trail) is one
// thread here; the mesh is never instanced in memory — a CTA per droplet transforms its 340 vertices into
// shared memory and tests the pixels of its screen box against them (arithmetic contract "VA-3", same
// operation sequence as oracle/raycast.c:triangle_depth).
#pragma once
#include "pcr_kernels.cuh"

namespace pcr {

constexpr int MAX_CTRL = 21;            // 20 spline samples + the current position (traj_renderer.py:272,331)
constexpr int HISTORY_FRAMES = 20;      // trail_length_frames (traj_renderer.py:218)
constexpr int TRAIL_SAMPLES = 20;
constexpr int DROP_MAX_RINGS = 64, DROP_MAX_SEGS = 64, DROP_MAX_VERTS = 2048;

// one kept spline sample of the reference's plan for a history of h frames: segment index and the
// float32 images of t, t^2, t^3 (python floats cast when they meet the float32 arrays)
struct TrailSample { int seg; float t1, t2, t3; };
struct TrailPlan { TrailSample s[HISTORY_FRAMES + 1][TRAIL_SAMPLES]; };    // indexed by h (3..20); h == 2: t1 = t, t2 = 1 - t

struct DropletMeshDev {
    const float* verts;      // [nv][3] object space, as a loader reads the OBJ text
    const float4* prof;      // [n_rings+1] ring radius, ring z, smooth normal (radial, z)
    float4 bound;            // bounding sphere of the whole mesh (0, 0, zc, R)
    int n_rings, n_segs, nv;
};

// generate_rotation_matrix_from_velocity (traj_renderer.py:159-202) in float64 with the constant default
// direction (0,0,-1) folded in (dot = -t_z, axis = (t_y, -t_x, 0)), same operations as
// oracle/droplet_oracle.py:rotation_from_velocity; the result is what a loader reads from the `{}`-formatted
// matrix: the float64 values rounded to float32.
__device__ __forceinline__ void rotation_from_velocity(float fx, float fy, float fz, float* R9)
{
    double R[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};
    const double vx = (double)fx, vy = (double)fy, vz = (double)fz;
    const double vn = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));
    bool rotate = vn >= 1e-6;
    double ax[3] = {0.0, 0.0, 0.0}, ang = 0.0;
    if (rotate) {
        const double tx = __ddiv_rn(vx, vn), ty = __ddiv_rn(vy, vn), tz = __ddiv_rn(vz, vn);
        const double d = fmin(fmax(-tz, -1.0), 1.0);
        ax[0] = ty; ax[1] = -tx; ax[2] = 0.0;
        double an = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(ax[0], ax[0]), __dmul_rn(ax[1], ax[1])), 0.0));
        if (an < 1e-8) {
            if (d > 0.999) rotate = false;
            else {
                const bool use_x = fabs(tx) < 0.9;
                const double m0 = use_x ? 1.0 : 0.0, m1 = use_x ? 0.0 : 1.0, m2 = 0.0;
                ax[0] = __dsub_rn(__dmul_rn(ty, m2), __dmul_rn(tz, m1));
                ax[1] = __dsub_rn(__dmul_rn(tz, m0), __dmul_rn(tx, m2));
                ax[2] = __dsub_rn(__dmul_rn(tx, m1), __dmul_rn(ty, m0));
                an = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(ax[0], ax[0]), __dmul_rn(ax[1], ax[1])), __dmul_rn(ax[2], ax[2])));
                if (an > 1e-8) { ax[0] = __ddiv_rn(ax[0], an); ax[1] = __ddiv_rn(ax[1], an); ax[2] = __ddiv_rn(ax[2], an); }
                else { ax[0] = 0.0; ax[1] = 1.0; ax[2] = 0.0; }
                ang = 3.141592653589793;
            }
        } else {
            ax[0] = __ddiv_rn(ax[0], an); ax[1] = __ddiv_rn(ax[1], an); ax[2] = __ddiv_rn(ax[2], an);
            ang = acos(d);
        }
    }
    if (rotate) {
        const double c = cos(ang), s = sin(ang), omc = __dsub_rn(1.0, c);
        const double K[9] = {0.0, -ax[2], ax[1], ax[2], 0.0, -ax[0], -ax[1], ax[0], 0.0};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double kk = __dadd_rn(__dadd_rn(__dmul_rn(K[3 * i], K[j]), __dmul_rn(K[3 * i + 1], K[3 + j])), __dmul_rn(K[3 * i + 2], K[6 + j]));
                R[3 * i + j] = __dadd_rn(__dadd_rn(i == j ? 1.0 : 0.0, __dmul_rn(s, K[3 * i + j])), __dmul_rn(omc, kk));
            }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) R9[k] = (float)R[k];
}

__device__ __forceinline__ float round6f(double x) { return (float)__ddiv_rn(rint(__dmul_rn(x, 1e6)), 1e6); }

// _add_trail_lines (traj_renderer.py:204-396) for one point: pa[0..h) = its history positions (oldest first,
// already the last <= 20), pos = its current position.  Writes the control points of the curve file (after
// the 6-decimal text round trip, float32) and returns how many (0 = no trail).  Same operations as
// oracle/droplet_oracle.py:history_trails.
__device__ __forceinline__ int history_trail(const float (*pa)[3], int h, const float* pos, const TrailPlan& plan, float (*out)[3])
{
    if (h < 2) return 0;
    double kept[3] = {0.0, 0.0, 0.0}, first[3] = {0.0, 0.0, 0.0};
    int cnt = 0;
    for (int k = 0; k <= TRAIL_SAMPLES; ++k) {
        float q[3];
        if (k == TRAIL_SAMPLES) { q[0] = pos[0]; q[1] = pos[1]; q[2] = pos[2]; }
        else {
            const TrailSample sm = plan.s[h][k];
            if (h == 2) {
#pragma unroll
                for (int c = 0; c < 3; ++c) q[c] = __fadd_rn(__fmul_rn(sm.t2, pa[0][c]), __fmul_rn(sm.t1, pa[1][c]));
            } else {
                const int s = sm.seg, last = h - 2;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float p0, p1 = pa[s][c], p2 = pa[s + 1][c], p3;
                    if (s == 0) { p0 = __fsub_rn(pa[0][c], __fsub_rn(pa[1][c], pa[0][c])); p3 = pa[min(2, h - 1)][c]; }
                    else if (s == last) { p0 = pa[s - 1][c]; p3 = __fadd_rn(pa[s + 1][c], __fsub_rn(pa[s + 1][c], pa[s][c])); }
                    else { p0 = pa[s - 1][c]; p3 = pa[min(s + 2, h - 1)][c]; }
                    const float a = __fmul_rn(__fadd_rn(-p0, p2), sm.t1);
                    const float b = __fmul_rn(__fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(2.0f, p0), __fmul_rn(5.0f, p1)), __fmul_rn(4.0f, p2)), p3), sm.t2);
                    const float d = __fmul_rn(__fadd_rn(__fsub_rn(__fadd_rn(-p0, __fmul_rn(3.0f, p1)), __fmul_rn(3.0f, p2)), p3), sm.t3);
                    q[c] = __fmul_rn(0.5f, __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(2.0f, p1), a), b), d));
                }
            }
        }
        if (!(isfinite(q[0]) && isfinite(q[1]) && isfinite(q[2]))) continue;
        const double x = (double)q[0], y = (double)q[1], z = (double)q[2];
        if (cnt > 0) {
            const double dx = __dsub_rn(x, kept[0]), dy = __dsub_rn(y, kept[1]), dz = __dsub_rn(z, kept[2]);
            const double dist = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            if (!(dist > 1e-5)) continue;
        } else { first[0] = x; first[1] = y; first[2] = z; }
        kept[0] = x; kept[1] = y; kept[2] = z;
        out[cnt][0] = round6f(x); out[cnt][1] = round6f(y); out[cnt][2] = round6f(z);
        ++cnt;
    }
    if (cnt >= 2) {            // an (almost) closed curve loses its last point (traj_renderer.py:368-373)
        const double dx = __dsub_rn(first[0], kept[0]), dy = __dsub_rn(first[1], kept[1]), dz = __dsub_rn(first[2], kept[2]);
        const double dist = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        if (dist < 1e-5) --cnt;
    }
    return cnt >= 2 ? cnt : 0;
}

// Per point of every frame of the batch: the droplet's to-world matrix [R | position] and the control points
// of its trail.  RAW: positions / velocities come from the caller's raw trajectory (K1 evaluated here, history
// frames included); otherwise from already transformed float32 arrays (the geometry-only entries).
//   trails 1: straight velocity trail (2 control points)   2: Catmull-Rom history trail (<= 21)
template <typename T>
__global__ void __launch_bounds__(128)
k_droplet_prepare(RawFrames<T> raw, long long n, int g0, StyleDev st, const FrameDev* __restrict__ frames, const float* __restrict__ rot,
                  const TrailPlan* __restrict__ plan, float* __restrict__ xf, float* __restrict__ ctrl, int* __restrict__ count)
{
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = g0 + b;                                            // frame index inside the caller's buffer
    const T* q = raw.in + (size_t)g * raw.frame_stride + i * raw.cols;
    const float4 p = k1_position<T>(__ldg(q), __ldg(q + 1), __ldg(q + 2), raw.stats + (size_t)g * 10, st, 0.0f);
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (raw.cols == 6) { v = k1_velocity<T>(q, st); rotation_from_velocity(v.x, v.y, v.z, R); }
    else if (rot) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = __ldg(rot + i * 9 + k);
    }
    float* M = xf + ((size_t)b * n + i) * 12;
    const float pos[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int r = 0; r < 3; ++r) { M[4 * r] = R[3 * r]; M[4 * r + 1] = R[3 * r + 1]; M[4 * r + 2] = R[3 * r + 2]; M[4 * r + 3] = pos[r]; }
    float (*out)[3] = reinterpret_cast<float (*)[3]>(ctrl + ((size_t)b * n + i) * MAX_CTRL * 3);
    int cnt = 0;
    if (raw.cols == 6 && st.trails == 1) {
        float tail[3], head[3];
        if (trail_ends(p, v, st, frames[b].trail_scale, tail, head)) {
            cnt = 2;
#pragma unroll
            for (int c = 0; c < 3; ++c) { out[0][c] = tail[c]; out[1][c] = head[c]; }
        }
    } else if (raw.cols == 6 && st.trails == 2) {
        const int h = min(HISTORY_FRAMES, g);
        float pa[HISTORY_FRAMES][3];
        for (int k = 0; k < h; ++k) {
            const int gh = g - h + k;
            const T* qh = raw.in + (size_t)gh * raw.frame_stride + i * raw.cols;
            const float4 ph = k1_position<T>(__ldg(qh), __ldg(qh + 1), __ldg(qh + 2), raw.stats + (size_t)gh * 10, st, 0.0f);
            pa[k][0] = ph.x; pa[k][1] = ph.y; pa[k][2] = ph.z;
        }
        cnt = history_trail(pa, h, pos, *plan, out);
    }
    count[(size_t)b * n + i] = cnt;
}

// geometry-only entries (pcr_droplet_transforms / pcr_history_trails) on transformed float32 arrays
__global__ void __launch_bounds__(128)
k_droplet_transforms(const float* __restrict__ pcl, long long n, int cols, const float* __restrict__ rot, float* __restrict__ xf)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* q = pcl + i * cols;
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    if (cols == 6) rotation_from_velocity(__ldg(q + 3), __ldg(q + 4), __ldg(q + 5), R);
    else if (rot) { for (int k = 0; k < 9; ++k) R[k] = __ldg(rot + i * 9 + k); }
    float* M = xf + i * 12;
    for (int r = 0; r < 3; ++r) { M[4 * r] = R[3 * r]; M[4 * r + 1] = R[3 * r + 1]; M[4 * r + 2] = R[3 * r + 2]; M[4 * r + 3] = __ldg(q + r); }
}

__global__ void __launch_bounds__(128)
k_history_trails(const float* __restrict__ hist, int h_all, const float* __restrict__ posn, long long n, const TrailPlan* __restrict__ plan,
                 float* __restrict__ ctrl, int* __restrict__ count)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int h = min(HISTORY_FRAMES, h_all);
    float pa[HISTORY_FRAMES][3];
    for (int k = 0; k < h; ++k)
        for (int c = 0; c < 3; ++c) pa[k][c] = __ldg(hist + ((size_t)(h_all - h + k) * n + i) * 3 + c);
    const float pos[3] = {__ldg(posn + i * 3), __ldg(posn + i * 3 + 1), __ldg(posn + i * 3 + 2)};
    count[i] = history_trail(pa, h, pos, *plan, reinterpret_cast<float (*)[3]>(ctrl + i * MAX_CTRL * 3));
}

// stats of frames that are already standardised (pcr_style.xform == 2): centre 0, scale 1 — k1_position then
// returns the coordinates unchanged ((x - 0) / 1 is exact)
__global__ void k_identity_stats(double* __restrict__ stats, int n_frames)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_frames) return;
    double* S = stats + (size_t)g * 10;
    S[0] = S[1] = S[2] = 0.0;
    S[3] = S[4] = S[5] = 0.0;
    S[6] = S[7] = S[8] = 1.0;
    S[9] = 1.0;
}

// floor / miss keys for every pixel (the droplet path has no tile lists to fill from)
__global__ void __launch_bounds__(256)
k_fill_floor(const FrameDev* __restrict__ frames, StyleDev st, unsigned long long* __restrict__ vis, long long vis_stride)
{
    const int b = blockIdx.z;
    const FrameDev& f = frames[b];
    const int px = blockIdx.x * 64 + (threadIdx.x & 63), py = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (px >= f.W || py >= f.H) return;
    vis[(size_t)b * vis_stride + (size_t)py * f.W + px] = floor_key(f, st, pix_u(f, px), pix_w(f, py));
}

__device__ __forceinline__ void to_camera(const FrameDev& f, const float* X, float* c)
{
    const float dx = __fsub_rn(X[0], f.O[0]), dy = __fsub_rn(X[1], f.O[1]), dz = __fsub_rn(X[2], f.O[2]);
    c[0] = fmaf(dz, f.L[2], fmaf(dy, f.L[1], __fmul_rn(dx, f.L[0])));
    c[1] = fmaf(dz, f.U[2], fmaf(dy, f.U[1], __fmul_rn(dx, f.U[0])));
    c[2] = fmaf(dz, f.D[2], fmaf(dy, f.D[1], __fmul_rn(dx, f.D[0])));
}

// Polylines: one warp per (frame, point) walks the segments of its curve; per segment the lanes stride over
// the pixels of its conservative box, skip those far from the projected axis, run VA-2 on the rest and merge
// with atomicMin.  Un-binned on purpose: a trail covers a few dozen pixels.
__global__ void __launch_bounds__(256)
k_raster_polylines(const FrameDev* __restrict__ frames, long long n, float radius, uint32_t cap_id_base, const float* __restrict__ ctrl,
                   const int* __restrict__ count, unsigned long long* __restrict__ vis, long long vis_stride,
                   const unsigned int* __restrict__ hz, int hz_stride)
{
    const int b = blockIdx.y;
    const FrameDev& f = frames[b];
    const int lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int cnt = __ldg(count + (size_t)b * n + i);
    if (cnt < 2) return;
    const float* cp = ctrl + ((size_t)b * n + i) * MAX_CTRL * 3;
    unsigned long long* out = vis + (size_t)b * vis_stride;
    const unsigned long long id = (unsigned long long)(cap_id_base + (uint32_t)i);
    const float r2 = __fmul_rn(radius, radius);
    float A[3], B[3];
    { const float X[3] = {__ldg(cp), __ldg(cp + 1), __ldg(cp + 2)}; to_camera(f, X, B); }
    for (int s = 0; s + 1 < cnt; ++s) {
        A[0] = B[0]; A[1] = B[1]; A[2] = B[2];
        { const float X[3] = {__ldg(cp + 3 * s + 3), __ldg(cp + 3 * s + 4), __ldg(cp + 3 * s + 5)}; to_camera(f, X, B); }
        int x0, x1, y0, y1;
        if (!capsule_bbox(f, A, B, radius, x0, x1, y0, y1)) continue;
        // occlusion cull (see k_raster_droplets): a segment entirely behind the farthest pre-pass winner of every 8x4 pixel
        // block its box touches cannot win a pixel
        if (hz && nearest_depth_bits(fminf(A[2], B[2]), radius) > hiz_far_bits_warp(hz + (size_t)b * hz_stride, (f.W + HZ_W - 1) / HZ_W, x0, x1, y0, y1, lane))
            continue;
        const CapsuleScreen cs = capsule_screen(f, A, B, radius);
        const float pad = cs.pad - 11.4f + 1.0f;                       // pixel half diagonal instead of the tile's
        const float ex = cs.bi - cs.ai, ey = cs.bj - cs.aj, ee = ex * ex + ey * ey;
        const int bw = x1 - x0 + 1;
        const long long npx = (long long)bw * (y1 - y0 + 1);
        for (long long k = lane; k < npx; k += 32) {
            const int px = x0 + (int)(k % bw), py = y0 + (int)(k / bw);
            if (!cs.all) {
                float hh = ee > 0.0f ? __fdividef(((float)px - cs.ai) * ex + ((float)py - cs.aj) * ey, ee) : 0.0f;
                hh = fminf(fmaxf(hh, 0.0f), 1.0f);
                const float qx = (float)px - (cs.ai + hh * ex), qy = (float)py - (cs.aj + hh * ey);
                if (qx * qx + qy * qy > pad * pad) continue;
            }
            const float u = pix_u(f, px), w = pix_w(f, py);
            const float vv = fmaf(u, u, fmaf(w, w, 1.0f));
            const float inv_vv = __fdiv_rn(1.0f, vv);
            float t;
            if (capsule_depth(A[0], A[1], A[2], B[0], B[1], B[2], r2, u, w, vv, inv_vv, f.near_clip, f.far_clip, t))
                atomicMin(out + (size_t)py * f.W + px, ((unsigned long long)__float_as_uint(t) << 32) | id);
        }
    }
}

// VA-3 — ray-triangle test for the ray s*(u,w,1) against camera-space vertices.  Same operation sequence as
// oracle/raycast.c:triangle_depth.
__device__ __forceinline__ bool triangle_depth(const float* v0, const float* v1, const float* v2, float u, float w,
                                               float near_clip, float far_clip, float& depth)
{
    const float e1x = __fsub_rn(v1[0], v0[0]), e1y = __fsub_rn(v1[1], v0[1]), e1z = __fsub_rn(v1[2], v0[2]);
    const float e2x = __fsub_rn(v2[0], v0[0]), e2y = __fsub_rn(v2[1], v0[1]), e2z = __fsub_rn(v2[2], v0[2]);
    const float px = fmaf(w, e2z, -e2y);
    const float py = fmaf(-u, e2z, e2x);
    const float pz = fmaf(u, e2y, -__fmul_rn(w, e2x));
    const float det = fmaf(e1z, pz, fmaf(e1y, py, __fmul_rn(e1x, px)));
    if (!(det != 0.0f)) return false;
    const float inv = __fdiv_rn(1.0f, det);
    const float sx = -v0[0], sy = -v0[1], sz = -v0[2];
    const float bu = __fmul_rn(fmaf(sz, pz, fmaf(sy, py, __fmul_rn(sx, px))), inv);
    if (!(bu >= 0.0f && bu <= 1.0f)) return false;
    const float qx = fmaf(sy, e1z, -__fmul_rn(sz, e1y));
    const float qy = fmaf(sz, e1x, -__fmul_rn(sx, e1z));
    const float qz = fmaf(sx, e1y, -__fmul_rn(sy, e1x));
    const float bv = __fmul_rn(fmaf(w, qy, fmaf(u, qx, qz)), inv);
    if (!(bv >= 0.0f && __fadd_rn(bu, bv) <= 1.0f)) return false;
    const float t = __fmul_rn(fmaf(e2z, qz, fmaf(e2y, qy, __fmul_rn(e2x, qx))), inv);
    if (!(t >= near_clip && t <= far_clip)) return false;
    depth = t;
    return true;
}

// Droplets: one WARP per (frame, point), triangle-order raster.  The warp transforms the mesh's vertices to
// camera space in its slice of shared memory (world vertex = M v, then the sphere-centre camera transform —
// VA-3); then each lane takes triangles, projects the three vertices to get the triangle's (padded) pixel box —
// one to a few pixels for a 2.5 mm triangle — and runs the VA-3 test on those pixel centres only, merging hits
// with atomicMin.  Same (pixel, triangle) arithmetic as testing every pixel against every triangle; the min is
// order independent.  A triangle with a vertex at or behind the eye plane takes the instance's whole screen box.
//
// Occlusion cull (dense scenes: 100 k droplets of ~300 pixels each bury one another a hundred deep — 6 % of C4D's
// droplets show a pixel): the frame's droplets are sorted into DROP_SLABS depth slabs (k_droplet_depth_bins /
// k_droplet_slab_edges / k_droplet_slab_order) and rasterised nearest slab first; after every slab k_hiz_from_vis turns the keys so far into the
// farthest depth per 8x4 pixel block, and the next slab drops every instance whose bounding sphere lies entirely behind
// the farthest winner of every block its screen box touches: it cannot win a pixel, so the keys are the same with the
// cull on or off.  (A subsample as occluders, like the sphere path's pre-pass, culled 28 % here: a droplet's box spans
// thirty blocks and a sparse sample leaves holes in them; the slabs in front of a droplet hold everything that hides it.)
// order == NULL: one pass over everything, no cull.
constexpr int DROP_SLABS = 8, DROP_BINS = 256;
// cumulative share of a frame's droplets in front of slab s (sixty-fourths): thin slabs in front, where the visible
// surface is; the deep half of the cloud goes in two passes that the Hi-Z empties almost completely
__constant__ int c_slab_share[DROP_SLABS + 1] = {0, 1, 2, 4, 8, 16, 32, 48, 64};

__global__ void __launch_bounds__(256)
k_raster_droplets(const FrameDev* __restrict__ frames, long long n, DropletMeshDev mesh, uint32_t id_base, const float* __restrict__ xf,
                  unsigned long long* __restrict__ vis, long long vis_stride, const int* __restrict__ order, const int* __restrict__ starts,
                  int slab, const unsigned int* __restrict__ hz, int hz_stride, unsigned long long* __restrict__ dbg)
{
    // persistent: the grid's warps stride over the slab's droplet list (order == NULL: droplets 0..n-1)
    extern __shared__ __align__(16) float s_cam[];   // [warps][nv][3] camera-space vertices
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const FrameDev& f = frames[b];
    const int first = order ? __ldg(starts + b * (DROP_SLABS + 1) + slab) : 0;
    const long long count = order ? (long long)(__ldg(starts + b * (DROP_SLABS + 1) + slab + 1) - first) : n;
    const int* list = order ? order + (size_t)b * n + first : nullptr;
    float* cam = s_cam + (size_t)warp * mesh.nv * 3;
    unsigned long long* out = vis + (size_t)b * vis_stride;
    const int ns = mesh.n_segs, ntri = 2 * mesh.n_rings * ns;
    for (long long k = (long long)blockIdx.x * warps + warp; k < count; k += (long long)gridDim.x * warps) {
        const long long i = list ? (long long)__ldg(list + k) : k;
        const float* M = xf + ((size_t)b * n + i) * 12;
        float m[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) m[q] = __ldg(M + q);
        // instance bounding sphere -> screen box (also the early out for droplets off screen / behind the eye)
        int bx0 = 0, bx1 = -1, by0 = 0, by1 = -1;
        {
            const float4 o = mesh.bound;
            const float X[3] = {m[2] * o.z + m[3], m[6] * o.z + m[7], m[10] * o.z + m[11]};
            float c[3];
            to_camera(f, X, c);
            const float R = o.w * 1.01f + 1e-5f + 4e-6f * (fabsf(c[0]) + fabsf(c[1]) + fabsf(c[2]));
            if (!(isfinite(m[0] + m[1] + m[2] + m[4] + m[5] + m[6] + m[8] + m[9] + m[10]) && sphere_bbox(f, c[0], c[1], c[2], R, bx0, bx1, by0, by1)))
                continue;                                // warp-uniform
#ifdef PCR_RASTER_STATS
            if (lane == 0 && hz) atomicAdd(dbg + 11, 1ull);
#endif
            if (hz && nearest_depth_bits(c[2], R) > hiz_far_bits_warp(hz + (size_t)b * hz_stride, (f.W + HZ_W - 1) / HZ_W, bx0, bx1, by0, by1, lane)) {
#ifdef PCR_RASTER_STATS
                if (lane == 0) atomicAdd(dbg + 12, 1ull);
#endif
                continue;                                // buried (warp-uniform)
            }
        }
        __syncwarp();                                    // the previous droplet's triangles are done with the vertices
        for (int v = lane; v < mesh.nv; v += 32) {
            const float ox = __ldg(mesh.verts + 3 * v), oy = __ldg(mesh.verts + 3 * v + 1), oz = __ldg(mesh.verts + 3 * v + 2);
            float X[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) X[r] = fmaf(m[4 * r], ox, fmaf(m[4 * r + 1], oy, fmaf(m[4 * r + 2], oz, m[4 * r + 3])));
            to_camera(f, X, cam + 3 * v);
        }
        __syncwarp();
        const unsigned long long id = (unsigned long long)(id_base + (uint32_t)i);
        for (int t = lane; t < ntri; t += 32) {
            const int quad = t >> 1, r = quad / ns, j = quad - r * ns, jn = j + 1 == ns ? 0 : j + 1;
            const float* ring0 = cam + 3 * r * ns;
            const float* ring1 = ring0 + 3 * ns;
            // the reference's faces per quad: (v0, v2, v1) and (v1, v2, v3)
            const float* v0 = (t & 1) ? ring0 + 3 * jn : ring0 + 3 * j;
            const float* v1 = ring1 + 3 * j;
            const float* v2 = (t & 1) ? ring1 + 3 * jn : ring0 + 3 * jn;
            int x0 = bx0, x1 = bx1, y0 = by0, y1 = by1;
            if (v0[2] > 1e-3f && v1[2] > 1e-3f && v2[2] > 1e-3f) {
                float i0, j0, i1, j1, i2, j2;
                pixel_of(f, v0[0], v0[1], v0[2], i0, j0);
                pixel_of(f, v1[0], v1[1], v1[2], i1, j1);
                pixel_of(f, v2[0], v2[1], v2[2], i2, j2);
                // A pixel centre can pass VA-3 only inside the projected triangle, up to the float error of the
                // barycentrics — which, in pixels, grows as the triangle turns edge-on (error ~1e-4 of an edge length
                // divided by the cosine of the tilt).  Half a pixel of padding covers tilts down to cos ~ 2e-4.
                constexpr float PAD = 0.5f;
                x0 = max(x0, (int)ceilf(fminf(i0, fminf(i1, i2)) - PAD)); x1 = min(x1, (int)floorf(fmaxf(i0, fmaxf(i1, i2)) + PAD));
                y0 = max(y0, (int)ceilf(fminf(j0, fminf(j1, j2)) - PAD)); y1 = min(y1, (int)floorf(fmaxf(j0, fmaxf(j1, j2)) + PAD));
            }
            for (int py = y0; py <= y1; ++py) {
                const float w = pix_w(f, py);
                for (int px = x0; px <= x1; ++px) {
                    float d;
                    if (triangle_depth(v0, v1, v2, pix_u(f, px), w, f.near_clip, f.far_clip, d))
                        atomicMin(out + (size_t)py * f.W + px, ((unsigned long long)__float_as_uint(d) << 32) | id);
                }
            }
        }
    }
}

// Depth bin of every droplet of a frame (camera depth of its bounding sphere's centre over [depth of the world origin -+
// 1.8]: a standardised cloud lies within sqrt(3) of it) and the frame's histogram of them; then the slab edges (bins
// [edges[s], edges[s+1]) = slab s, nearest first) and the droplet indices in slab order.
__global__ void __launch_bounds__(256)
k_droplet_depth_bins(const FrameDev* __restrict__ frames, long long n, float4 bound, const float* __restrict__ xf,
                     unsigned char* __restrict__ dbin, unsigned int* __restrict__ hist)
{
    __shared__ unsigned int s_h[DROP_BINS];
    const int b = blockIdx.y;
    const FrameDev& f = frames[b];
    s_h[threadIdx.x] = 0u;
    __syncthreads();
    const float z0 = -(f.D[0] * f.O[0] + f.D[1] * f.O[1] + f.D[2] * f.O[2]) - 1.8f;
    for (long long i = (long long)blockIdx.x * 1024 + threadIdx.x; i < min(n, (long long)(blockIdx.x + 1) * 1024); i += 256) {
        const float* M = xf + ((size_t)b * n + i) * 12;
        const float X[3] = {__ldg(M + 2) * bound.z + __ldg(M + 3), __ldg(M + 6) * bound.z + __ldg(M + 7), __ldg(M + 10) * bound.z + __ldg(M + 11)};
        const float cz = (X[0] - f.O[0]) * f.D[0] + (X[1] - f.O[1]) * f.D[1] + (X[2] - f.O[2]) * f.D[2];
        const float t = (cz - z0) * ((float)DROP_BINS / 3.6f);
        const int bin = t >= 0.0f ? min((int)t, DROP_BINS - 1) : 0;           // NaN -> 0
        dbin[(size_t)b * n + i] = (unsigned char)bin;
        atomicAdd(&s_h[bin], 1u);
    }
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(hist + b * DROP_BINS + threadIdx.x, s_h[threadIdx.x]);
}

__global__ void k_droplet_slab_edges(unsigned int* __restrict__ hist, long long n, int* __restrict__ edges, int* __restrict__ starts, int* __restrict__ cursor)
{
    // edges[s] = first depth bin of slab s, starts[s] = droplets in front of it; a slab ends at the first bin boundary at
    // or beyond its share c_slab_share[s + 1] / 64 of the frame
    const int b = blockIdx.x;
    if (threadIdx.x != 0) return;
    unsigned int* h = hist + b * DROP_BINS;
    int* e = edges + b * (DROP_SLABS + 1);
    int* st = starts + b * (DROP_SLABS + 1);
    long long acc = 0;
    int s = 1;
    e[0] = 0; st[0] = 0;
    for (int k = 0; k < DROP_BINS; ++k) {
        acc += h[k];
        h[k] = 0u;                                   // ready for the next batch
        while (s < DROP_SLABS && acc >= (n * c_slab_share[s] + 63) / 64) { e[s] = k + 1; st[s] = (int)acc; ++s; }
    }
    while (s <= DROP_SLABS) { e[s] = DROP_BINS; st[s] = (int)acc; ++s; }
    for (int k = 0; k < DROP_SLABS; ++k) cursor[b * DROP_SLABS + k] = st[k];
}

// order[b][starts[s] ...) = the droplets of slab s (any order): per block, shared-memory counts per slab, one range
// reservation per (block, slab), ranks inside the block
__global__ void __launch_bounds__(256)
k_droplet_slab_order(long long n, const unsigned char* __restrict__ dbin, const int* __restrict__ edges, int* __restrict__ cursor,
                     int* __restrict__ order)
{
    __shared__ int s_cnt[DROP_SLABS], s_base[DROP_SLABS], s_edge[DROP_SLABS + 1];
    const int b = blockIdx.y;
    if (threadIdx.x < DROP_SLABS) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x <= DROP_SLABS) s_edge[threadIdx.x] = edges[b * (DROP_SLABS + 1) + threadIdx.x];
    __syncthreads();
    int slab[4], rank[4];
    const long long i0 = (long long)blockIdx.x * 1024 + threadIdx.x;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const long long i = i0 + q * 256;
        slab[q] = -1;
        if (i < n) {
            const int bin = (int)dbin[(size_t)b * n + i];
            int s = 0;
            while (s + 1 < DROP_SLABS && bin >= s_edge[s + 1]) ++s;
            slab[q] = s;
            rank[q] = atomicAdd(&s_cnt[s], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < DROP_SLABS && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(cursor + b * DROP_SLABS + threadIdx.x, s_cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (slab[q] >= 0) order[(size_t)b * n + s_base[slab[q]] + rank[q]] = (int)(i0 + q * 256);
}

// Level 1 of the Hi-Z straight from a visibility buffer: farthest depth (float bits) per 8x4 pixel block.  One warp per
// strip of 32 x 4 pixels (four blocks): a lane reads its column's four keys, then a max over each group of 8 lanes.
__global__ void __launch_bounds__(256)
k_hiz_from_vis(const FrameDev* __restrict__ frames, const unsigned long long* __restrict__ vis, long long vis_stride,
               unsigned int* __restrict__ hz, int hz_stride)
{
    const int b = blockIdx.z;
    const FrameDev& f = frames[b];
    const int W = f.W, H = f.H, hzw = (W + HZ_W - 1) / HZ_W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = (blockIdx.x * 8 + warp) * 32 + lane, by = blockIdx.y;          // strip column, block row
    if ((px & ~31) >= W || by * HZ_H >= H) return;
    const unsigned long long* v = vis + (size_t)b * vis_stride;
    unsigned int far_bits = 0u;
    if (px < W)
#pragma unroll
        for (int r = 0; r < HZ_H; ++r) {
            const int py = by * HZ_H + r;
            if (py < H) far_bits = max(far_bits, (unsigned int)(v[(size_t)py * W + px] >> 32));
        }
    far_bits = max(far_bits, __shfl_xor_sync(0xffffffffu, far_bits, 4));
    far_bits = max(far_bits, __shfl_xor_sync(0xffffffffu, far_bits, 2));
    far_bits = max(far_bits, __shfl_xor_sync(0xffffffffu, far_bits, 1));
    static_assert(HZ_W == 8 && HZ_H == 4, "strip layout");
    if ((lane & 7) == 0 && px < W) hz[(size_t)b * hz_stride + by * hzw + px / HZ_W] = far_bits;
}

// radiance leaving a diffuse surface point P with unit normal (nx,ny,nz): emitter + ground bounce
__device__ __forceinline__ float lit_radiance(const StyleDev& st, const FloorLut& lut, float Px, float Py, float Pz, float nx, float ny, float nz)
{
    const float Fd = st.light_z > Pz ? rect_form_factor_clipped(Px, Py, Pz, nx, ny, nz, st.light_half, st.light_z)
                                     : rect_form_factor(Px, Py, Pz, nx, ny, nz, st.light_half, st.light_z);
    float Li = 0.0f;
    if (st.has_floor) Li = st.bounce * st.floor_albedo * st.radiance * floor_form_factor(lut, st, Px, Py) * 0.5f * (1.0f - nz);
    return st.radiance * Fd + Li;
}

// K4 of the droplet scene: ids [0,n) droplets (smooth surface-of-revolution normal from the ring profile),
// [n,2n) trails (normal from the nearest segment's axis), floor / miss as in k_shade.
template <typename T>
__global__ void __launch_bounds__(256)
k_shade_droplets(const FrameDev* __restrict__ frames, StyleDev st, FloorLut lut, const uint64_t* __restrict__ vis, long long vis_stride,
                 RawFrames<T> raw, int g0, long long n, DropletMeshDev mesh, const float* __restrict__ xf, const float* __restrict__ ctrl,
                 const int* __restrict__ count, uint32_t* __restrict__ rgba, long long rgba_stride)
{
    const int b = blockIdx.z;
    const FrameDev& f = frames[b];
    const int px = blockIdx.x * 64 + (threadIdx.x & 63), py = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (px >= f.W || py >= f.H) return;
    const size_t p = (size_t)py * f.W + px;
    const uint64_t key = __ldg(vis + (size_t)b * vis_stride + p);
    const uint32_t id = (uint32_t)key;
    uint32_t* dst = rgba + (size_t)b * rgba_stride + p;
    if (id >= ID_FLOOR || (long long)id >= 2 * n) {
        *dst = shade_pixel<float, false>(f, st, lut, key, px, py, nullptr, nullptr, RawFrames<float>{}, b, 0, 0u, 0);
        return;
    }
    const float t = __uint_as_float((uint32_t)(key >> 32));
    const float u = pix_u(f, px), w = pix_w(f, py);
    const float dwx = fmaf(w, f.U[0], fmaf(u, f.L[0], f.D[0]));
    const float dwy = fmaf(w, f.U[1], fmaf(u, f.L[1], f.D[1]));
    const float dwz = fmaf(w, f.U[2], fmaf(u, f.L[2], f.D[2]));
    const float Px = fmaf(t, dwx, f.O[0]), Py = fmaf(t, dwy, f.O[1]), Pz = fmaf(t, dwz, f.O[2]);
    float nx = 0.f, ny = 0.f, nz = 1.f, rgb[3];
    if ((long long)id < n) {
        const float* M = xf + ((size_t)b * n + id) * 12;
        float m[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) m[k] = __ldg(M + k);
        const float dx = Px - m[3], dy = Py - m[7], dz = Pz - m[11];
        const float qx = m[0] * dx + m[4] * dy + m[8] * dz, qy = m[1] * dx + m[5] * dy + m[9] * dz, qz = m[2] * dx + m[6] * dy + m[10] * dz;
        int r = 0;
        while (r + 1 < mesh.n_rings && qz < __ldg(&mesh.prof[r + 1].y)) ++r;
        const float4 a = __ldg(mesh.prof + r), c = __ldg(mesh.prof + r + 1);
        float fr = a.y > c.y ? (a.y - qz) / (a.y - c.y) : 0.0f;
        fr = fminf(fmaxf(fr, 0.0f), 1.0f);
        float nr = a.z + fr * (c.z - a.z), nzz = a.w + fr * (c.w - a.w);
        float l = sqrtf(nr * nr + nzz * nzz);
        if (l > 0.0f) { nr /= l; nzz /= l; } else { nr = 1.0f; nzz = 0.0f; }
        const float rho = sqrtf(qx * qx + qy * qy);
        const float cx = rho > 1e-12f ? qx / rho : 1.0f, cy = rho > 1e-12f ? qy / rho : 0.0f;
        const float ox = nr * cx, oy = nr * cy;
        nx = m[0] * ox + m[1] * oy + m[2] * nzz; ny = m[4] * ox + m[5] * oy + m[6] * nzz; nz = m[8] * ox + m[9] * oy + m[10] * nzz;
        l = sqrtf(nx * nx + ny * ny + nz * nz);
        if (l > 0.0f) { nx /= l; ny /= l; nz /= l; } else { nx = 0.f; ny = 0.f; nz = 1.f; }
        const int g = g0 + b;
        float speed = 0.0f;
        if (raw.cols == 6) speed = k1_velocity<T>(raw.in + (size_t)g * raw.frame_stride + (long long)id * raw.cols, st).w;
        k1_colour<T>(make_float4(m[3], m[7], m[11], 0.f), speed, raw.stats + (size_t)g * 10, st, raw.user_rgb, (long long)id, rgb);
    } else {
        const long long k = (long long)id - n;
        const int cnt = __ldg(count + (size_t)b * n + k);
        const float* cp = ctrl + ((size_t)b * n + k) * MAX_CTRL * 3;
        float best = INFINITY;
        for (int s = 0; s + 1 < cnt; ++s) {
            const float ax = __ldg(cp + 3 * s), ay = __ldg(cp + 3 * s + 1), az = __ldg(cp + 3 * s + 2);
            const float dx = __ldg(cp + 3 * s + 3) - ax, dy = __ldg(cp + 3 * s + 4) - ay, dz = __ldg(cp + 3 * s + 5) - az;
            const float dd = dx * dx + dy * dy + dz * dz;
            float h = dd > 0.0f ? ((Px - ax) * dx + (Py - ay) * dy + (Pz - az) * dz) / dd : 0.0f;
            h = fminf(fmaxf(h, 0.0f), 1.0f);
            const float qx = Px - (ax + h * dx), qy = Py - (ay + h * dy), qz = Pz - (az + h * dz);
            const float l2 = qx * qx + qy * qy + qz * qz;
            if (l2 < best) { best = l2; nx = qx; ny = qy; nz = qz; }
        }
        const float l = sqrtf(nx * nx + ny * ny + nz * nz);
        if (l > 0.0f) { nx /= l; ny /= l; nz /= l; } else { nx = 0.f; ny = 0.f; nz = 1.f; }
        rgb[0] = st.trail_rgb[0]; rgb[1] = st.trail_rgb[1]; rgb[2] = st.trail_rgb[2];
    }
    const float Lo = lit_radiance(st, lut, Px, Py, Pz, nx, ny, nz);
    *dst = srgb8(rgb[0] * Lo) | (srgb8(rgb[1] * Lo) << 8) | (srgb8(rgb[2] * Lo) << 16) | 0xFF000000u;
}

}  // namespace pcr
