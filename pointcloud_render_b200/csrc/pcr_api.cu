// pcr_api.cu — host side of the C ABI declared in include/pcr.h.
// Replaces the reference's generate_xml_content -> save_xml -> mi.load_file -> mi.render ->
// write_bitmap(sRGB8) chain (example_renderer.py:113-161) with stream-ordered kernel launches.
#include "pcr.h"
#include "pcr_kernels.cuh"
#include "pcr_droplets.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

using namespace pcr;

namespace {

constexpr int RING_SLOTS = 16;         // pinned camera-frame ring: the host may run this many uploads ahead of the device
constexpr int MAX_STAT_BLOCKS = 1184;   // 148 SMs x 8
// K0 products (standardisation statistics + the compact pre-pass sample) of up to PREP_SLOTS batches can exist ahead of
// the batch being rendered; region PREP_SLOTS of the stats scratch belongs to the entries that run K0 in line on the
// caller's stream (pcr_standardize, pcr_stats_partial, pcr_render_transformed, the droplet path).
constexpr int PREP_SLOTS = 6;
constexpr int HOST_STAGES = 6;          // staging slots of the host-buffer entry (chunks in flight: H2D | K0 + serial mean | kernels | D2H)
constexpr int HOST_TICKETS = 16;
constexpr int INLINE_REGION = PREP_SLOTS;

enum KernelId { KID_STATS = 0, KID_TRANSFORM, KID_PROJECT, KID_SCAN, KID_SCATTER, KID_RASTER, KID_HIZ, KID_SHADE,
                KID_ZMIN, KID_AXIS, KID_FILL, KID_MEAN, KID_LUT, KID_DROP_PREP, KID_FILL_FLOOR, KID_RASTER_POLY, KID_RASTER_DROP,
                KID_SHADE_DROP, KID_PEER_INIT, KID_SHADE_PEER, KID_SHADE_FLOOR, KID_TRAILS, KID_COUNT };
const char* const kKernelNames[KID_COUNT] = {"k_stats", "k_transform", "k_project_count", "k_scan_tiles", "k_scatter",
                                             "k_raster_tiles", "k_hiz", "k_shade", "k_zmin", "k_axis_transform",
                                             "k_fill_tiles", "k_mean_sequential", "k_build_floor_lut", "k_droplet_prepare",
                                             "k_fill_floor", "k_raster_polylines", "k_raster_droplets", "k_shade_droplets", "k_peer_init_rows",
                                             "k_shade_peer", "k_shade_floor_tiles", "k_raster_trails"};
constexpr size_t PROF_MAX_RECORDS = 1 << 16;

struct ProfRec { int kid; cudaEvent_t a, b; };

}  // namespace

struct pcr_ctx {
    int device = 0;
    long long max_points = 0;
    int max_w = 0, max_h = 0, max_batch = 1;
    long long pair_cap = 0;
    int tiles_cap = 0;
    int num_sms = 148;
    std::string err;
    long long launches = 0;

    // scratch, all [max_batch][...]
    float4* sph = nullptr;            // per survivor: camera-space sphere (cx, cy, cz, r)
    uint4* rect = nullptr;            // per survivor: pixel bbox + sphere index (K2a -> K2b, K3)
    unsigned int* surv_count = nullptr;
    int gx_cap = 0;
    double *partials = nullptr, *stats = nullptr;
    unsigned int* done = nullptr;
    unsigned int *counts = nullptr, *offsets = nullptr, *cursor = nullptr, *overflow = nullptr;
    float4* p_sph = nullptr;          // per (tile, primitive) pair, in tile order (K2b -> K3): centre + r^2,
    uint2* p_ci = nullptr;            // cull word + key id,
    unsigned int* tile_state = nullptr;   // [max_batch][tiles_cap] lazy floor fill (BinDev::tile_state)
    unsigned long long* scan_part = nullptr;    // k_scan_tiles: per-stripe totals / ready flags (BinDev)
    unsigned int* scan_ready = nullptr;
    unsigned int *fill_list = nullptr, *fill_count = nullptr;    // lazy floor fill: tiles k_fill_tiles has to visit (BinDev)
    int scan_stripes = 1;
    unsigned int scan_epoch = 0;
    int lazy_fill = 1;                // PCR_LAZY_FILL=0 disables (diagnostics)
    void* sample = nullptr;           // [2][max_batch][ceil(n/step)][3] every step-th point of a frame, written by K0 for the pre-pass
    size_t sample_bytes = 0;
    int sample_prepass = 1;           // PCR_SAMPLE_PREPASS=0 disables (diagnostics)
    int stats_ahead = 1;              // K0 of the next batches on side streams while a batch renders (PCR_STATS_AHEAD=0: in line)
    int raster_ctas_per_sm = 4;       // k_raster_tiles: resident CTAs per SM (occupancy query)
    unsigned long long* stat_pairs = nullptr;
    unsigned int *item_count = nullptr, *item_next = nullptr;
    uint4* items = nullptr;
    int item_cap = 0;
    int smem_optin = 48 * 1024;       // max dynamic shared memory per block (opt-in), capped at 200 KB
    int smem_max = 48 * 1024;         // ... uncapped (the serial mean's exclusive blocks)
    float* lut = nullptr;             // floor form-factor table (LUT_N^2), rebuilt when the scene constants change
    float lut_key[7] = {0, 0, 0, 0, 0, 0, 0};
    bool lut_valid = false;
    unsigned int* hz = nullptr;       // [max_batch][hz_cap] farthest pre-pass depth per 8x4 pixel block
    int hz_cap = 0;
    int occlusion = -1;               // -1 auto (n >= occlusion_min_points), 0 off, 1 always
    int scatter_threads = BIN_THREADS; // K2b threads per chunk (PCR_SCATTER_THREADS: diagnostics)
    int two_phase = 1;                // K2a's coarse-then-fine Hi-Z cull (PCR_TWO_PHASE=0 disables: diagnostics)
    // Occluder pre-pass: the sampled spheres deeper than the cloud's centre plane + prepass_zcut (standardised units: the cloud's
    // largest extent is 1) lose to nearer ones almost everywhere; only every prepass_back-th of them takes part.  Measured on H
    // (profiles/r02xy_prepass_sweep.md): step 16 / no cut 21.8 k frames/s, step 8 / cut -0.1 / back 8 24.1 k, no back part at all 24.5 k.
    float prepass_zcut = -0.1f;       // PCR_PREPASS_ZCUT (1e38 = off)
    int prepass_back = 8;             // ... but every prepass_back-th (a power of two) of those still takes part (PCR_PREPASS_BACK)
    int mean_fused = 1;               // the serial mean's kernel also produces min / max / scale / sample: no k_stats (PCR_MEAN_FUSED=0: diagnostics)
    int scatter_merge = 0;            // K2b blocks take several K2a chunks (PCR_SCATTER_MERGE=1; measured slower on H: 156 vs 124 us per launch)
    int mean_exclusive = 1;           // the serial mean's blocks keep their SMs to themselves (PCR_MEAN_EXCLUSIVE=0: diagnostics)
    int cull4 = 1;                    // k_project_cull4 where it applies (PCR_CULL4=0 disables: diagnostics)
    int occlusion_step = 0;           // the pre-pass rasterises every step-th point (of those in front of prepass_zcut, see below);
                                      // 0 = by the size of the cloud (prepass_step): 8 up to 1.5 M points, 64 from 8 M
    long long occlusion_min_points = 1 << 17;
    // Two nested pre-passes for large clouds (n >= occlusion_min_points2): every (step2 * ratio)-th point first, then every
    // step2-th point culled by the first one's Hi-Z, then all points culled by the second one's — the occluders
    // themselves are mostly buried, and the coarser level removes them before they cost a list entry.
    int occlusion_levels = 1;         // 1: the single pre-pass (default); 2 (PCR_OCCLUSION_LEVELS=2): the two nested ones — measured on H
                                      // they halve the main pass's pairs but cost as much as they save (20.3 k vs 21.2 k frames/s)
    int occlusion_step2 = 8, occlusion_ratio = 8;
    long long occlusion_min_points2 = 1 << 19;
    unsigned int* hz_b = nullptr;     // second Hi-Z buffer (the finer pre-pass builds its own while it is culled by the coarser one's)
    uint64_t* vis = nullptr;          // lazily allocated when the caller passes d_vis == NULL
    FrameDev* d_frames = nullptr;
    FrameDev* h_frames = nullptr;     // pinned ring: RING_SLOTS x max_batch
    cudaEvent_t ring_ev[RING_SLOTS] = {};
    bool ring_used[RING_SLOTS] = {};
    int ring_next = 0;

    // host-buffer pipeline (pcr_render_frames_host), lazily created
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[HOST_STAGES] = {}, ev_comp[HOST_STAGES] = {}, ev_d2h[HOST_STAGES] = {};
    void* stage_in[HOST_STAGES] = {};
    uint8_t* stage_rgba[HOST_STAGES] = {};
    uint64_t* stage_vis[HOST_STAGES] = {};      // allocated when a caller first asks for the keys
    size_t stage_in_bytes = 0;
    unsigned long long host_chunks = 0;         // chunks submitted so far, over all calls (chunk c uses slot c % HOST_STAGES)
    // PCR_HOST_TRACE=1 (diagnostics): timed events at the stages of every chunk of the host pipeline, printed by pcr_host_wait(-1)
    int host_trace = 0;
    struct TraceRec { unsigned long long chunk; int nb; cudaEvent_t h2d0, h2d1, ready, comp0, comp1, d2h1; };
    std::vector<TraceRec> trace;
    cudaEvent_t ticket_ev[HOST_TICKETS] = {};
    long long tickets = 0;                      // calls submitted so far
    float *stage_radius = nullptr, *stage_rgb = nullptr;

    long long last_overflow_frames = 0;

    // fused z-merge over peer memory (pcr_peer_*): this rank's merged z-buffer / image (cudaMalloc, so that
    // cudaIpcGetMemHandle maps them at offset 0) and the pointer table of every rank's
    void* peer_merged = nullptr;
    void* peer_image = nullptr;
    int peer_w = 0, peer_h = 0;
    PeerDev peer = {};
    bool peer_tiles_valid = false;    // tile_state holds the tiles the last pcr_render_shard_peer drew in (k_shade_peer)

    // droplet scene (pcr_render_droplet_frames): mesh tables, spline plan, per-frame stats of the whole
    // buffer, per-point matrices / control points of one batch — all lazily allocated
    float* mesh_verts = nullptr;
    float4* mesh_prof = nullptr;
    float4 mesh_bound = {0.f, 0.f, 0.f, 0.f};
    int mesh_rings = 0, mesh_segs = 0;
    TrailPlan* plan = nullptr;
    double* dstats = nullptr;
    size_t dstats_frames = 0;
    float *dxf = nullptr, *dctrl = nullptr;
    int* dcount = nullptr;
    unsigned char* dbin = nullptr;    // depth bin of every droplet (occlusion cull by depth slabs)
    unsigned int* dhist = nullptr;    // [max_batch][DROP_BINS], zero between batches
    int* dedges = nullptr;            // [max_batch][DROP_SLABS + 1] first depth bin of every slab
    int* dstarts = nullptr;           // [max_batch][DROP_SLABS + 1] droplets in front of every slab
    int* dcursor = nullptr;           // [max_batch][DROP_SLABS]
    int* dorder = nullptr;            // [batch * n] droplet indices in slab order
    size_t dprep_points = 0;          // capacity of dxf / dctrl / dcount / dbin in points (batch * n)

    // The scratch is shared by every entry point, so work issued on DIFFERENT streams must not
    // overlap: each entry waits for the previous entry's last event when the stream changed.
    cudaEvent_t ev_last = nullptr;
    cudaStream_t last_stream = nullptr;
    bool has_last = false;

    // stats-ahead pipeline: K0 — and the serial reference-exact mean, 2 ms per million points of pure latency — of up
    // to PREP_SLOTS batches runs on side streams (one per slot, so the serial chains of different batches overlap)
    // while earlier batches render.  A slot is keyed by the batch's device pointer and shape; pcr_render_frames looks
    // its batches up before computing anything, so a caller (or pcr_render_frames_host, or pcr_prefetch_frames) can
    // have the statistics of frames it will render later computed now.
    struct PrepSlot {
        const void* in = nullptr; int is_f64 = 0; long long n = 0; int cols = 0, nb = 0, mean_mode = 0, sampled = 0, sstep = 0;
        bool valid = false;           // holds products nobody has consumed yet
        bool used = false;            // ev_free has been recorded at least once
        unsigned long long stamp = 0; // age (eviction order of unconsumed hints)
        cudaStream_t stream = nullptr;
        cudaEvent_t ev_fork = nullptr, ev_ready = nullptr, ev_free = nullptr;
    } prep[PREP_SLOTS];
    unsigned long long prep_stamp = 0;

    // per-kernel CUDA-event timing (pcr_profile / pcr_profile_read)
    bool profiling = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> prof_pool;
};

namespace {

int fail(pcr_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess)
{
    if (c) {
        c->err = what;
        if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
    }
    return code;
}

#define CK(call)                                                            \
    do {                                                                    \
        cudaError_t e_ = (call);                                            \
        if (e_ != cudaSuccess) return fail(ctx, PCR_ERR_CUDA, #call, e_);   \
    } while (0)

cudaEvent_t prof_event(pcr_ctx* c)
{
    cudaEvent_t e = nullptr;
    if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
    if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return e;
}

inline void prof_begin(pcr_ctx* c, int kid, cudaStream_t s)
{
    if (!c->profiling || c->prof.size() >= PROF_MAX_RECORDS) return;
    ProfRec r{kid, prof_event(c), prof_event(c)};
    if (!r.a || !r.b) return;
    cudaEventRecord(r.a, s);
    c->prof.push_back(r);
}

inline void prof_end(pcr_ctx* c, int kid, cudaStream_t s)
{
    if (!c->profiling || c->prof.empty() || c->prof.back().kid != kid) return;
    cudaEventRecord(c->prof.back().b, s);
}

// LAUNCH(kernel id, stream, kernel<<<...>>>(...)) — counts the launch and, when profiling is on,
// brackets it with two events on the launching stream.
#define LAUNCH(kid, stream, ...)                                                        \
    do {                                                                                \
        prof_begin(ctx, kid, stream);                                                   \
        __VA_ARGS__;                                                                    \
        cudaError_t e_ = cudaGetLastError();                                            \
        prof_end(ctx, kid, stream);                                                     \
        if (e_ != cudaSuccess) return fail(ctx, PCR_ERR_CUDA, kKernelNames[kid], e_);   \
        ctx->launches++;                                                                \
    } while (0)

StyleDev to_style_dev(const pcr_style* s)
{
    StyleDev d;
    d.color_mode = s->color_mode;
    for (int k = 0; k < 3; ++k) d.const_rgb[k] = s->const_rgb[k];
    d.radius = s->radius; d.flip_x = s->flip_x; d.z_lift = s->z_lift; d.vel_norm = s->vel_norm;
    d.has_floor = s->has_floor; d.floor_z = s->floor_z;
    for (int k = 0; k < 2; ++k) { d.floor_min[k] = s->floor_min[k]; d.floor_max[k] = s->floor_max[k]; }
    d.floor_albedo = s->floor_albedo; d.light_z = s->light_z; d.light_half = s->light_half;
    d.radiance = s->radiance; d.bounce = s->bounce; d.xform = s->xform;
    d.trails = s->trails; d.trail_radius = s->trail_radius;
    for (int k = 0; k < 3; ++k) d.trail_rgb[k] = s->trail_rgb[k];
    d.trail_len_min = s->trail_len_min; d.trail_len_max = s->trail_len_max;
    return d;
}

// Camera frame in double on the host, rounded once to f32.  Mitsuba 3 look_at /
// perspective conventions for the <sensor> block of XMLTemplates.HEAD (example_renderer.py:16-31).
int camera_frame_host(const pcr_camera* cam, pcr_frame* f)
{
    if (!cam || !f || cam->width <= 0 || cam->height <= 0) return PCR_ERR_INVALID;
    double o[3], d[3], u[3], l[3], nu[3];
    for (int k = 0; k < 3; ++k) { o[k] = cam->origin[k]; d[k] = (double)cam->target[k] - (double)cam->origin[k]; u[k] = cam->up[k]; }
    double len = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (!(len > 0.0)) return PCR_ERR_INVALID;
    for (int k = 0; k < 3; ++k) d[k] = d[k] / len;
    l[0] = u[1] * d[2] - u[2] * d[1];
    l[1] = u[2] * d[0] - u[0] * d[2];
    l[2] = u[0] * d[1] - u[1] * d[0];
    len = sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
    if (!(len > 0.0)) return PCR_ERR_INVALID;
    for (int k = 0; k < 3; ++k) l[k] = l[k] / len;
    nu[0] = d[1] * l[2] - d[2] * l[1];
    nu[1] = d[2] * l[0] - d[0] * l[2];
    nu[2] = d[0] * l[1] - d[1] * l[0];
    double T = tan((double)cam->fov_x_deg * 3.14159265358979323846 / 360.0);
    for (int k = 0; k < 3; ++k) { f->L[k] = (float)l[k]; f->U[k] = (float)nu[k]; f->D[k] = (float)d[k]; f->O[k] = (float)o[k]; }
    f->T = (float)T;
    f->Th = (float)(T * (double)cam->height / (double)cam->width);
    f->TW = (float)(T / (double)cam->width);
    f->near_clip = cam->near_clip; f->far_clip = cam->far_clip;
    f->W = cam->width; f->H = cam->height;
    return PCR_OK;
}

void to_frame_dev(const pcr_frame& f, FrameDev* d)
{
    for (int k = 0; k < 3; ++k) { d->L[k] = f.L[k]; d->U[k] = f.U[k]; d->D[k] = f.D[k]; d->O[k] = f.O[k]; }
    d->T = f.T; d->Th = f.Th; d->TW = f.TW; d->inv2TW = 1.0f / (2.0f * f.TW);
    d->near_clip = f.near_clip; d->far_clip = f.far_clip;
    d->W = f.W; d->H = f.H;
    d->tiles_x = (f.W + TILE - 1) / TILE; d->tiles_y = (f.H + TILE - 1) / TILE;
    d->trail_scale = 0.0;
}

BinDev bin_of(pcr_ctx* c)
{
    BinDev b;
    b.counts = c->counts; b.offsets = c->offsets; b.cursor = c->cursor;
    b.p_sph = c->p_sph; b.p_ci = c->p_ci; b.tile_state = nullptr;
    b.scan_part = c->scan_part; b.scan_ready = c->scan_ready; b.scan_stripes = c->scan_stripes;
    b.fill_list = c->fill_list; b.fill_count = c->fill_count;
    b.overflow = c->overflow; b.stat_pairs = c->stat_pairs; b.tiles_cap = c->tiles_cap; b.pair_cap = c->pair_cap;
    b.item_count = c->item_count; b.item_next = c->item_next; b.items = c->items; b.item_cap = c->item_cap;
    b.surv_count = c->surv_count; b.gx_cap = c->gx_cap;
    return b;
}

// Floor form-factor table for this style (built once per distinct set of scene constants).
int floor_lut(pcr_ctx* ctx, const StyleDev& st, cudaStream_t stream, FloorLut* out)
{
    out->data = nullptr; out->x0 = out->y0 = out->inv_cx = out->inv_cy = 0.0f;
    const float w = st.floor_max[0] - st.floor_min[0], h = st.floor_max[1] - st.floor_min[1];
    if (!st.has_floor || !(st.light_z > st.floor_z) || !(w > 0.0f) || !(h > 0.0f)) return PCR_OK;     // evaluated directly
    const float key[7] = {st.floor_z, st.floor_min[0], st.floor_min[1], st.floor_max[0], st.floor_max[1], st.light_z, st.light_half};
    const float cx = w / (float)(LUT_N - 1), cy = h / (float)(LUT_N - 1);
    if (!ctx->lut) CK(cudaMalloc((void**)&ctx->lut, sizeof(float) * LUT_N * LUT_N));
    if (!ctx->lut_valid || memcmp(key, ctx->lut_key, sizeof(key)) != 0) {
        dim3 grid((LUT_N + 255) / 256, LUT_N);
        LAUNCH(KID_LUT, stream, k_build_floor_lut<<<grid, 256, 0, stream>>>(ctx->lut, st.floor_min[0], st.floor_min[1], cx, cy,
                                                                            st.light_z - st.floor_z, st.light_half));
        memcpy(ctx->lut_key, key, sizeof(key));
        ctx->lut_valid = true;
    }
    out->data = ctx->lut; out->x0 = st.floor_min[0]; out->y0 = st.floor_min[1]; out->inv_cx = 1.0f / cx; out->inv_cy = 1.0f / cy;
    return PCR_OK;
}

// Upload `nb` cameras into d_frames through the pinned ring.
int upload_frames(pcr_ctx* ctx, const pcr_camera* cams, int nb, cudaStream_t stream)
{
    int slot = ctx->ring_next;
    ctx->ring_next = (slot + 1) % RING_SLOTS;
    if (ctx->ring_used[slot]) CK(cudaEventSynchronize(ctx->ring_ev[slot]));
    FrameDev* h = ctx->h_frames + (size_t)slot * ctx->max_batch;
    for (int b = 0; b < nb; ++b) {
        pcr_frame f;
        if (camera_frame_host(cams + b, &f) != PCR_OK) return fail(ctx, PCR_ERR_INVALID, "degenerate camera");
        if (f.W > ctx->max_w || f.H > ctx->max_h) return fail(ctx, PCR_ERR_CAPACITY, "frame larger than the context");
        to_frame_dev(f, h + b);
        h[b].trail_scale = cams[b].trail_scale;
    }
    CK(cudaMemcpyAsync(ctx->d_frames, h, sizeof(FrameDev) * nb, cudaMemcpyHostToDevice, stream));
    CK(cudaEventRecord(ctx->ring_ev[slot], stream));
    ctx->ring_used[slot] = true;
    return PCR_OK;
}

// region: which copy of the partials / done scratch this launch uses (a prepared-batch slot, or INLINE_REGION)
int launch_stats(pcr_ctx* ctx, const void* d_in, int in_is_f64, long long n, int cols, long long frame_stride,
                 int nb, int region, double* stats, int finalize, cudaStream_t stream, int mean_mode = PCR_MEAN_F64,
                 void* sample = nullptr, long long sample_stride = 0, int sample_step = 1)
{
    double* partials = ctx->partials + (size_t)region * ctx->max_batch * MAX_STAT_BLOCKS * 9;
    unsigned int* done = ctx->done + (size_t)region * ctx->max_batch;
    // 32 points per thread: the per-block reduction (f64 shuffles, last-block fold) must not outweigh the streaming
    int blocks = (int)std::min<long long>((n + 256 * 32 - 1) / (256 * 32), MAX_STAT_BLOCKS);
    if ((long long)blocks * nb < 2 * ctx->num_sms)        // few frames: keep every SM busy with smaller chunks
        blocks = (int)std::min<long long>((n + 2047) / 2048, (2 * ctx->num_sms + nb - 1) / nb);
    blocks = std::min(std::max(blocks, 1), MAX_STAT_BLOCKS);
    dim3 grid(blocks, nb);
    // the reference's own (sequential, input-dtype) mean replaces the f64 one unless the caller asked for PCR_MEAN_F64; its
    // kernel then produces the rest of K0 as well (min / max / scale / pre-pass sample: helper warps on the chains' SMs) and
    // k_stats is not launched at all
    const bool sequential = finalize == 1 && mean_mode != PCR_MEAN_F64;
    // (float frames on 16-byte boundaries only: the helpers must keep out of the chains' way, which takes 16-byte loads)
    const bool aligned16 = !in_is_f64 && ((uintptr_t)d_in % 16 == 0) && (nb == 1 || (frame_stride * 4) % 16 == 0);
    const bool fused_k0 = sequential && ctx->mean_fused && aligned16;
    // float4 streaming needs 3 columns and every frame base on a 16-byte boundary
    const int vec = aligned16 && cols == 3;
    if (fused_k0) {
    } else if (in_is_f64)
        LAUNCH(KID_STATS, stream, k_stats<double><<<grid, 256, 0, stream>>>((const double*)d_in, n, cols, frame_stride, partials, MAX_STAT_BLOCKS, stats, done, finalize, 0,
                                                                        (double*)sample, sample_stride, (unsigned int)sample_step));
    else
        LAUNCH(KID_STATS, stream, k_stats<float><<<grid, 256, 0, stream>>>((const float*)d_in, n, cols, frame_stride, partials, MAX_STAT_BLOCKS, stats, done, finalize, vec,
                                                                       (float*)sample, sample_stride, (unsigned int)sample_step));
    if (sequential) {
        // MEAN_LANES frames per block: three chain warps (one per axis, lane = frame) + one producer warp (+ the helper warps)
        const int mean_blocks = (nb + MEAN_LANES - 1) / MEAN_LANES;
        const size_t mean_smem = ctx->mean_exclusive ? std::max<size_t>(MEAN_SMEM_BYTES, (size_t)ctx->smem_max) : MEAN_SMEM_BYTES;   // see MEAN_SMEM_BYTES
        const int mean_threads = fused_k0 ? 128 + 32 * MEAN_HELPERS : 128;
#define PCR_MEAN(T, C) LAUNCH(KID_MEAN, stream, (k_mean_sequential<T, C><<<mean_blocks, mean_threads, mean_smem, stream>>>((const T*)d_in, n, frame_stride, stats, nb, \
                                                 fused_k0 ? 1 : 0, fused_k0 ? (T*)sample : nullptr, sample_stride, (unsigned int)std::max(sample_step, 1))))
        if (in_is_f64) { if (cols == 3) PCR_MEAN(double, 3); else PCR_MEAN(double, 6); }
        else { if (cols == 3) PCR_MEAN(float, 3); else PCR_MEAN(float, 6); }
#undef PCR_MEAN
    }
    return PCR_OK;
}

// Every how-many-th point the occluder pre-pass takes.  What it needs is a roughly constant NUMBER of near spheres — enough to
// cover the cloud's silhouette a few times over, and a sphere's share of that silhouette does not depend on the film size —
// so the step grows with the cloud: 8 at the 1 M-point headline (sweep: profiles/r02xy_prepass_sweep.md), 64 for 50 M
// points (C5 on one GPU, gpurun_out/r03s_*: step 8 543 frames/s, 16 696, 32 790, 64 795).
int prepass_step(const pcr_ctx* ctx, long long n)
{
    if (ctx->occlusion_step > 0) return ctx->occlusion_step;
    return (int)std::min<long long>(64, std::max<long long>(8, 8 * ((n + 500000) / 1000000)));
}

double* inline_stats(pcr_ctx* ctx) { return ctx->stats + (size_t)INLINE_REGION * ctx->max_batch * 10; }
double* slot_stats(pcr_ctx* ctx, int slot) { return ctx->stats + (size_t)slot * ctx->max_batch * 10; }

// two nested occluder pre-passes (pcr_ctx::occlusion_levels) for this cloud size?
bool two_prepasses(const pcr_ctx* ctx, long long n)
{
    return ctx->occlusion_levels >= 2 && n >= ctx->occlusion_min_points2 && n > (long long)ctx->occlusion_step2 * ctx->occlusion_ratio * 64;
}

// ---- prepared batches (see pcr_ctx::PrepSlot) ----------------------------------------------------------------------
// What the occluder pre-pass needs from K0 for frames of n points: every sstep-th point, compact.
struct SamplePlan { bool sampled; int sstep; long long stride; };          // stride in elements per frame
SamplePlan sample_plan(const pcr_ctx* ctx, long long n)
{
    const bool occl = ctx->occlusion > 0 || (ctx->occlusion < 0 && n >= ctx->occlusion_min_points);
    SamplePlan sp;
    sp.sstep = two_prepasses(ctx, n) ? ctx->occlusion_step2 : prepass_step(ctx, n);
    sp.sampled = ctx->sample_prepass && occl && n > sp.sstep;
    sp.stride = sp.sampled ? ((n + sp.sstep - 1) / sp.sstep) * 3 : 0;
    return sp;
}

int ensure_sample(pcr_ctx* ctx, const SamplePlan& sp, size_t elem)
{
    if (!sp.sampled) return PCR_OK;
    const size_t need = (size_t)PREP_SLOTS * ((ctx->max_batch * (size_t)sp.stride * elem + 255) & ~(size_t)255);
    if (ctx->sample_bytes >= need) return PCR_OK;
    CK(cudaDeviceSynchronize());
    for (pcr_ctx::PrepSlot& p : ctx->prep) p.valid = false;          // their samples are gone
    if (ctx->sample) CK(cudaFree(ctx->sample));
    ctx->sample = nullptr; ctx->sample_bytes = 0;
    CK(cudaMalloc(&ctx->sample, need));
    ctx->sample_bytes = need;
    return PCR_OK;
}

// (a slot's region starts at a multiple of the allocation's quarter, whatever n is: a hint for frames of one size is
// never overwritten by the K0 of frames of another size that landed in a different slot)
char* slot_sample(pcr_ctx* ctx, int slot, const SamplePlan& sp, size_t)
{
    return sp.sampled ? (char*)ctx->sample + (size_t)slot * (ctx->sample_bytes / PREP_SLOTS) : nullptr;
}

int find_prepared(pcr_ctx* ctx, const void* in, int is_f64, long long n, int cols, int nb, int mean_mode, const SamplePlan& sp)
{
    for (int k = 0; k < PREP_SLOTS; ++k) {
        const pcr_ctx::PrepSlot& p = ctx->prep[k];
        if (p.valid && p.in == in && p.is_f64 == is_f64 && p.n == n && p.cols == cols && p.nb == nb && p.mean_mode == mean_mode &&
            p.sampled == (int)sp.sampled && p.sstep == sp.sstep)
            return k;
    }
    return -1;
}

// K0 (+ the serial mean) of one batch into a free slot, ordered after everything `after` holds at this point (the
// frames may still be in flight there).  on_side: run on the slot's own high-priority stream (so it overlaps whatever
// `after` does next); otherwise on `after` itself.  Returns the slot through *slot_out.
int prepare_batch(pcr_ctx* ctx, const void* in, int is_f64, long long n, int cols, int nb, int mean_mode, const SamplePlan& sp,
                  cudaStream_t after, bool on_side, int* slot_out)
{
    int slot = -1;
    for (int k = 0; k < PREP_SLOTS && slot < 0; ++k) if (!ctx->prep[k].valid) slot = k;
    if (slot >= 0) {                                   // among the free ones, the one released longest ago
        for (int k = 0; k < PREP_SLOTS; ++k)
            if (!ctx->prep[k].valid && ctx->prep[k].stamp < ctx->prep[slot].stamp) slot = k;
    } else {                                           // every slot holds an unconsumed hint: drop the oldest
        slot = 0;
        for (int k = 1; k < PREP_SLOTS; ++k) if (ctx->prep[k].stamp < ctx->prep[slot].stamp) slot = k;
    }
    pcr_ctx::PrepSlot& p = ctx->prep[slot];
    if (!p.ev_ready) {
        CK(cudaEventCreateWithFlags(&p.ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p.ev_ready, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p.ev_free, cudaEventDisableTiming));
    }
    cudaStream_t q = after;
    if (on_side) {
        if (!p.stream) {
            int lo = 0, hi = 0;
            CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CK(cudaStreamCreateWithPriority(&p.stream, cudaStreamNonBlocking, hi));
        }
        q = p.stream;
        CK(cudaEventRecord(p.ev_fork, after));
        CK(cudaStreamWaitEvent(q, p.ev_fork, 0));
    }
    if (p.used) CK(cudaStreamWaitEvent(q, p.ev_free, 0));           // the batch that last used this slot has been rendered
    const size_t elem = is_f64 ? 8 : 4;
    int rc = launch_stats(ctx, in, is_f64, n, cols, n * cols, nb, slot, slot_stats(ctx, slot), 1, q, mean_mode,
                          slot_sample(ctx, slot, sp, elem), sp.stride, sp.sstep);
    if (rc) return rc;
    CK(cudaEventRecord(p.ev_ready, q));
    p.in = in; p.is_f64 = is_f64; p.n = n; p.cols = cols; p.nb = nb; p.mean_mode = mean_mode; p.sampled = sp.sampled; p.sstep = sp.sstep;
    p.valid = true; p.stamp = ++ctx->prep_stamp;
    *slot_out = slot;
    return PCR_OK;
}

int launch_transform(pcr_ctx* ctx, const void* d_in, int in_is_f64, long long n, int cols, long long frame_stride,
                     int nb, const float* d_radius, const float* d_rgb, const double* stats, const StyleDev& st,
                     float4* pos, float4* attr, float4* vel, long long out_stride, cudaStream_t stream)
{
    dim3 grid((unsigned)((n + 255) / 256), nb);
    if (in_is_f64)
        LAUNCH(KID_TRANSFORM, stream, k_transform<double><<<grid, 256, 0, stream>>>((const double*)d_in, n, cols, frame_stride, d_radius, d_rgb, stats, st, pos, attr, vel, out_stride));
    else
        LAUNCH(KID_TRANSFORM, stream, k_transform<float><<<grid, 256, 0, stream>>>((const float*)d_in, n, cols, frame_stride, d_radius, d_rgb, stats, st, pos, attr, vel, out_stride));
    return PCR_OK;
}

// K2a -> scan -> K2b -> K3 (-> fallback) (-> K4) for nb frames whose cameras are already in d_frames.
// Type-erased raw source of the fused path (pcr_render_frames): K1 is evaluated inside K2a / K4.
struct RawSrc {
    const void* in; int is_f64; long long frame_stride; int cols; const double* stats; const float* radius; const float* rgb;
    const void* sample = nullptr; long long sample_stride = 0;    // every occlusion_step-th point (x, y, z), compact: K0 -> pre-pass
};

template <typename T>
RawFrames<T> raw_frames(const RawSrc* r)
{
    RawFrames<T> f;
    f.in = r ? (const T*)r->in : nullptr; f.frame_stride = r ? r->frame_stride : 0; f.cols = r ? r->cols : 3;
    f.stats = r ? r->stats : nullptr; f.radius = r ? r->radius : nullptr; f.user_rgb = r ? r->rgb : nullptr;
    return f;
}

int launch_shade(pcr_ctx* ctx, const StyleDev& st, const uint64_t* vis_in, long long vis_stride, const float4* pos, const float4* attr,
                 long long in_stride, const RawSrc* raw, long long n, int nb, uint32_t id_base, int owner_only, int W, int H,
                 uint8_t* rgba, long long rgba_stride, cudaStream_t stream, const unsigned int* tile_state = nullptr)
{
    uint64_t* vis = const_cast<uint64_t*>(vis_in);      // written only where tile_state says the keys do not exist yet
    const int tiles_cap = ctx->tiles_cap;
    FloorLut lut;
    int rc = floor_lut(ctx, st, stream, &lut);
    if (rc) return rc;
    dim3 grid((unsigned)((W + 63) / 64), (unsigned)((H + 4 * SHADE_ROWS - 1) / (4 * SHADE_ROWS)), nb);
    if (tile_state) {
        // the tiles nothing was drawn in: keys + ground shading in one go; k_shade takes the others from a list
        const int tiles = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
        LAUNCH(KID_SHADE_FLOOR, stream, k_active_tiles<<<nb, 1024, 0, stream>>>(ctx->d_frames, ctx->tile_state, tiles_cap));
        grid = dim3((unsigned)std::max(1, std::min((tiles + 3) / 4, 2 * ctx->num_sms * 4 / nb + 1)), 1, nb);
        dim3 fgrid((unsigned)std::max(1, std::min((tiles + 7) / 8, ctx->num_sms * 32 / nb)), nb);
        LAUNCH(KID_SHADE_FLOOR, stream, k_shade_floor_tiles<<<fgrid, 256, 0, stream>>>(ctx->d_frames, st, lut, tile_state, tiles_cap,
                                                                                     (unsigned long long*)vis, vis_stride, (uint32_t*)rgba, rgba_stride,
                                                                                     owner_only && id_base != 0 ? 1 : 0));
    }
    if (!raw)
        LAUNCH(KID_SHADE, stream, k_shade<float, false><<<grid, 256, 0, stream>>>(ctx->d_frames, st, lut, vis, vis_stride, pos, attr, in_stride,
                                                                                 raw_frames<float>(nullptr), n, id_base, owner_only, (uint32_t*)rgba, rgba_stride, tile_state, tiles_cap));
    else if (raw->is_f64)
        LAUNCH(KID_SHADE, stream, k_shade<double, true><<<grid, 256, 0, stream>>>(ctx->d_frames, st, lut, vis, vis_stride, nullptr, nullptr, 0,
                                                                                 raw_frames<double>(raw), n, id_base, owner_only, (uint32_t*)rgba, rgba_stride, tile_state, tiles_cap));
    else
        LAUNCH(KID_SHADE, stream, k_shade<float, true><<<grid, 256, 0, stream>>>(ctx->d_frames, st, lut, vis, vis_stride, nullptr, nullptr, 0,
                                                                                raw_frames<float>(raw), n, id_base, owner_only, (uint32_t*)rgba, rgba_stride, tile_state, tiles_cap));
    return PCR_OK;
}

int launch_render(pcr_ctx* ctx, const float4* pos, const float4* attr, long long in_stride, const RawSrc* raw, long long n, int nb,
                  uint32_t id_base, const StyleDev& st, int W, int H, uint64_t* vis, long long vis_stride,
                  uint8_t* rgba, long long rgba_stride, int owner_only, cudaStream_t stream, const PeerDev* push = nullptr)
{
    BinDev bin = bin_of(ctx);
    PeerDev no_peer = {};                              // world == 0: nothing is pushed
    const PeerDev peer_final = push ? *push : no_peer;
    const int tiles = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    // K2 blocks own contiguous point chunks and histogram their pairs in shared memory when the
    // tile count allows it (count: 4 B/tile, scatter: 8 B/tile)
    const int use_smem = (size_t)tiles * 8 <= (size_t)ctx->smem_optin ? 1 : 0;
    const int resident = ctx->num_sms * (2048 / BIN_THREADS);
    unsigned long long* v = (unsigned long long*)vis;
    // velocity trails (a line raster into the finished keys, k_raster_trails) only exist in the fused whole-path entry
    const int trails = raw && st.trails == 1 && raw->cols == 6 ? 1 : 0;
    if (trails && 2 * (unsigned long long)n > 0xFFFFFFF0ull) return fail(ctx, PCR_ERR_INVALID, "too many points for trail ids (n + i)");
    const uint32_t cap_id_base = trails ? (uint32_t)n : 0u;
    // lazy floor fill: tiles in which nothing is drawn get their keys from K4 — only when K4 follows in this very call
    // (trails are merged into the finished keys of every tile with atomicMin: the tiles nothing was binned in need their
    // floor keys first, so frames with trails take the eager fill)
    // The fused peer merge is lazy too: a rank's local keys are only ever read in the tiles its own passes drew in (the
    // seeded raster, and k_shade_peer through the same tile states), the merged buffer has its own floor keys.
    const bool lazy = ctx->lazy_fill && (rgba != nullptr || push != nullptr) && !trails;
    ctx->peer_tiles_valid = lazy && push != nullptr;
    if (lazy) bin.tile_state = ctx->tile_state;
    const long long slots = ctx->max_points;           // survivor slots per frame

    // one binning + raster pass over `np` spheres (sphere i = point i*step)
    auto pass = [&](long long np, int step, const unsigned int* hz, int seeded, const PeerDev& peer, unsigned int* hz_out, int sample_step) -> int {
        unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>((np + 2047) / 2048, std::max(1, 2 * resident / nb)));
        // (films too large for the shared-memory histograms use per-pair global atomics; the same large chunks serve them
        // best — 2048-point chunks made 24 k tiny blocks of a 50 M-point cloud: K2a / K2b 435 / 454 -> 379 / 390 us on C5)
        gx = std::min<unsigned>(gx, (unsigned)ctx->gx_cap);
        if (np > 0) {
            dim3 grid(gx, nb);
            // two-phase cull (see k_project_count): main pass over a Hi-Z, tile histogram + coarse Hi-Z level + the
            // warps' rings must fit in shared memory
            const int hz_w1 = (W + HZ_W - 1) / HZ_W, hz_h1 = (H + HZ_H - 1) / HZ_H;
            const size_t sm2 = (size_t)tiles * 4 + (size_t)((hz_w1 + 3) / 4) * ((hz_h1 + 3) / 4) * 4 + (size_t)(BIN_THREADS / 32) * 5 * RING_CAP * 4;
            const int two_phase = (ctx->two_phase && hz && use_smem && sm2 <= (size_t)ctx->smem_optin / 2) ? 1 : 0;
            const size_t sm = two_phase ? sm2 : (use_smem ? (size_t)tiles * 4 : 0);
            // the pre-pass reads the compact sample K0 wrote (point i of the pass = sample row i) when there is one
            RawSrc rsrc = raw ? *raw : RawSrc{};
            int fetch_step = step;
            // (the sample holds every sample_step-th point; a coarser pass strides over it)
            if (raw && step > 1 && raw->sample && step % sample_step == 0) { rsrc.in = raw->sample; rsrc.frame_stride = raw->sample_stride; rsrc.cols = 3; fetch_step = step / sample_step; }
            const RawSrc* rawp = raw ? &rsrc : nullptr;
            // the headline case — float (n,3) frames on 16-byte boundaries, n a multiple of 4, one radius, every point, a Hi-Z to
            // cull against — takes the four-points-per-thread kernel with the coarse-cell table (k_project_cull4)
            const size_t sm4 = (size_t)((tiles + 3) & ~3) * 4 + (size_t)((hz_w1 + 3) / 4 + 2) * ((hz_h1 + 3) / 4 + 2) * 16 + (size_t)(BIN_THREADS / 32) * 4 * RING4_CAP * 4;
            const bool cull4 = ctx->cull4 && raw && !raw->is_f64 && raw->cols == 3 && step == 1 && !raw->radius && hz && use_smem && ctx->two_phase &&
                               sm4 <= (size_t)ctx->smem_optin / 2 && (np & 3) == 0 && np < (1ll << 31) && ((uintptr_t)raw->in & 15) == 0 &&
                               ((raw->frame_stride * 4) & 15) == 0;
            if (cull4) {
                LAUNCH(KID_PROJECT, stream, (k_project_cull4<<<grid, BIN_THREADS, sm4, stream>>>((const float*)raw->in, np, raw->frame_stride, raw->stats, st, ctx->d_frames,
                                                                                            ctx->sph, ctx->rect, slots, bin, hz, ctx->hz_cap)));
            } else
#define PCR_PROJECT(T, RAWB, posarg, strarg, rawarg)                                                                             \
    LAUNCH(KID_PROJECT, stream, (k_project_count<T, RAWB><<<grid, BIN_THREADS, sm, stream>>>(                                    \
        posarg, np, strarg, rawarg, st, fetch_step, ctx->d_frames, ctx->sph, ctx->rect, slots, bin, use_smem, hz, ctx->hz_cap, two_phase, step, (hz_out && !hz && step > 1) ? ctx->prepass_zcut : 3.0e38f, (unsigned int)ctx->prepass_back - 1u)))
            { if (!raw) PCR_PROJECT(float, false, pos, in_stride, raw_frames<float>(nullptr));
            else if (raw->is_f64) PCR_PROJECT(double, true, nullptr, 0, raw_frames<double>(rawp));
            else PCR_PROJECT(float, true, nullptr, 0, raw_frames<float>(rawp)); }
#undef PCR_PROJECT
        }
        if (++ctx->scan_epoch == 0u) ctx->scan_epoch = 1u;            // 0 = the flags' initial value
        LAUNCH(KID_SCAN, stream, k_scan_tiles<<<dim3((unsigned)((tiles + 4095) / 4096), nb), 1024, 0, stream>>>(ctx->d_frames, bin, np, lazy ? (seeded ? 2 : 1) : 0, ctx->scan_epoch,
                                                                                                                    lazy ? hz_out : nullptr, ctx->hz_cap));
        if (lazy && seeded) {
            // tiles only the main pass touches need the floor keys the raster starts from
            dim3 grid((unsigned)std::max(1, std::min(std::min((tiles + 7) / 8, ctx->num_sms * 32 / nb), 16)), nb);   // a short list per frame
            LAUNCH(KID_FILL, stream, k_fill_tiles<<<grid, 256, 0, stream>>>(ctx->d_frames, st, bin, v, vis_stride, nullptr, ctx->hz_cap, 2));
        }
        if (np > 0) {
            // a K2b block takes several K2a chunks, so that the whole launch is one wave of resident blocks
            const int per_frame = std::max(1, PCR_SCATTER_BLOCKS * ctx->num_sms / nb);
            const int merge = ctx->scatter_merge ? std::min<int>(SCATTER_MERGE_MAX, std::max<int>(1, ((int)gx + per_frame - 1) / per_frame)) : 1;
            dim3 grid((gx + merge - 1) / merge, nb);
            LAUNCH(KID_SCATTER, stream, k_scatter<<<grid, ctx->scatter_threads, use_smem ? tiles * 8 : 0, stream>>>(
                np, ctx->d_frames, ctx->sph, ctx->rect, slots, bin, use_smem, id_base, (uint32_t)step, (int)gx, merge));
        }
        if (!seeded) {
            // floor keys of empty tiles, all-ones preset of split tiles
            dim3 grid((unsigned)std::max(1, std::min(std::min((tiles + 7) / 8, ctx->num_sms * 32 / nb), lazy ? 16 : 1 << 30)), nb);
            LAUNCH(KID_FILL, stream, k_fill_tiles<<<grid, 256, 0, stream>>>(ctx->d_frames, st, bin, v, vis_stride, hz_out, ctx->hz_cap, lazy ? 1 : 0));
        }
        if (np > 0) {
            // persistent raster: CTAs pull (tile, <= ITEM_SPHERES spheres) items from per-frame queues
            const int raster_ctas = ctx->num_sms * ctx->raster_ctas_per_sm;
            dim3 grid((unsigned)std::max(1, std::min(raster_ctas, tiles * nb)));
            LAUNCH(KID_RASTER, stream, k_raster_tiles<<<grid, RASTER_CTA_THREADS, sizeof(RasterStage) * RASTER_STAGES, stream>>>(
                ctx->d_frames, st, ctx->sph, ctx->rect, slots, bin, id_base, (uint32_t)step, v, vis_stride, nb, np, seeded, (int)gx, peer, hz_out, ctx->hz_cap));
        }
        return PCR_OK;
    };

    const bool occl = ctx->occlusion > 0 || (ctx->occlusion < 0 && n >= ctx->occlusion_min_points);
    const unsigned int* trail_hz = nullptr;
    const int hzn2p = (((W + HZ_W - 1) / HZ_W + 3) / 4 + 2) * (((H + HZ_H - 1) / HZ_H + 3) / 4 + 2);
    // Hi-Z of a finished pre-pass: level-1 entries were written by k_fill_tiles (empty tiles), the scan (untouched tiles, lazy
    // fill) and the raster (single-item tiles); the tiles split into several items are re-read, then level 2 is built
    auto finish_hiz = [&](unsigned int* hzbuf, int listed) -> int {
        dim3 grid((unsigned)(listed ? std::min((tiles + 7) / 8, 8) : (tiles + 7) / 8), nb);
        LAUNCH(KID_HIZ, stream, k_hiz_split<<<grid, 256, 0, stream>>>(ctx->d_frames, bin, v, vis_stride, hzbuf, ctx->hz_cap, listed));
        dim3 grid2((unsigned)((hzn2p + 255) / 256), nb);                // one thread per cell of the padded level-2 grid
        LAUNCH(KID_HIZ, stream, k_hiz2<<<grid2, 256, 0, stream>>>(ctx->d_frames, hzbuf, ctx->hz_cap));
        return PCR_OK;
    };
    if (occl && two_prepasses(ctx, n)) {
        // every (step2 * ratio)-th point -> Hi-Z A; every step2-th point, culled by A and seeded with the first pass's keys
        // -> Hi-Z B (which starts as a copy of A: tiles the second pass does not touch keep their entries); then all points
        const int s1 = ctx->occlusion_step2, s0 = s1 * ctx->occlusion_ratio;
        if (!ctx->hz_b) CK(cudaMalloc((void**)&ctx->hz_b, sizeof(unsigned int) * (size_t)ctx->max_batch * (size_t)ctx->hz_cap));
        int rc = pass((n + s0 - 1) / s0, s0, nullptr, 0, no_peer, ctx->hz, s1);
        if (rc) return rc;
        if ((rc = finish_hiz(ctx->hz, lazy ? 1 : 0))) return rc;
        CK(cudaMemcpyAsync(ctx->hz_b, ctx->hz, sizeof(unsigned int) * (size_t)nb * (size_t)ctx->hz_cap, cudaMemcpyDeviceToDevice, stream));
        rc = pass((n + s1 - 1) / s1, s1, ctx->hz, 1, no_peer, ctx->hz_b, s1);
        if (rc) return rc;
        if ((rc = finish_hiz(ctx->hz_b, 0))) return rc;             // (the seeded scan's list holds other tiles: look at every tile)
        rc = pass(n, 1, ctx->hz_b, 1, peer_final, nullptr, s1);
        trail_hz = ctx->hz_b;
        if (rc) return rc;
    } else if (occl && n > prepass_step(ctx, n)) {
        // occluder pre-pass (every step-th point, true ids) -> Hi-Z -> main pass seeded with its keys
        const int step = prepass_step(ctx, n);
        int rc = pass((n + step - 1) / step, step, nullptr, 0, no_peer, ctx->hz, step);       // only the final pass pushes to the peers
        if (rc) return rc;
        if ((rc = finish_hiz(ctx->hz, lazy ? 1 : 0))) return rc;
        rc = pass(n, 1, ctx->hz, 1, peer_final, nullptr, step);
        trail_hz = ctx->hz;
        if (rc) return rc;
    } else {
        int rc = pass(n, 1, nullptr, 0, peer_final, nullptr, 1);
        if (rc) return rc;
    }
    if (trails) {
        // the velocity trail of every point, as a line raster straight into the finished keys (k_raster_trails); culled by
        // the occluder pre-pass's Hi-Z where there was one
        dim3 grid((unsigned)((n + 7) / 8), nb);
        if (raw->is_f64)
            LAUNCH(KID_TRAILS, stream, k_raster_trails<double><<<grid, 256, 0, stream>>>(raw_frames<double>(raw), n, st, ctx->d_frames, cap_id_base, v, vis_stride, trail_hz, ctx->hz_cap));
        else
            LAUNCH(KID_TRAILS, stream, k_raster_trails<float><<<grid, 256, 0, stream>>>(raw_frames<float>(raw), n, st, ctx->d_frames, cap_id_base, v, vis_stride, trail_hz, ctx->hz_cap));
    }
    if (rgba) return launch_shade(ctx, st, vis, vis_stride, pos, attr, in_stride, raw, n, nb, id_base, owner_only, W, H, rgba, rgba_stride, stream,
                                  lazy ? ctx->tile_state : nullptr);
    return PCR_OK;
}

// Serialise entries that use the context's scratch across streams (see pcr_ctx::ev_last).
int enter(pcr_ctx* ctx, cudaStream_t s)
{
    if (ctx->has_last && ctx->last_stream != s) CK(cudaStreamWaitEvent(s, ctx->ev_last, 0));
    return PCR_OK;
}

int leave(pcr_ctx* ctx, cudaStream_t s)
{
    if (!ctx->ev_last) CK(cudaEventCreateWithFlags(&ctx->ev_last, cudaEventDisableTiming));
    CK(cudaEventRecord(ctx->ev_last, s));
    ctx->last_stream = s;
    ctx->has_last = true;
    return PCR_OK;
}

int check_common(pcr_ctx* ctx, long long n, int cols, const pcr_style* style)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (!style) return fail(ctx, PCR_ERR_INVALID, "style is NULL");
    if (n < 0 || (cols != 3 && cols != 6)) return fail(ctx, PCR_ERR_INVALID, "n < 0 or cols not 3|6");
    if (style->mean_mode < 0 || style->mean_mode > 2) return fail(ctx, PCR_ERR_INVALID, "mean_mode not in {0,1,2}");
    if (n > ctx->max_points) return fail(ctx, PCR_ERR_CAPACITY, "n exceeds the context's max_points");
    return PCR_OK;
}

int ensure_vis(pcr_ctx* ctx)
{
    if (!ctx->vis) CK(cudaMalloc(&ctx->vis, sizeof(uint64_t) * (size_t)ctx->max_batch * ctx->max_w * ctx->max_h));
    return PCR_OK;
}

}  // namespace

extern "C" {

int pcr_abi_version(void) { return PCR_ABI_VERSION; }

const char* pcr_last_error(const pcr_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int pcr_camera_frame(const pcr_camera* cam, pcr_frame* out) { return camera_frame_host(cam, out); }

int pcr_create(pcr_ctx** out, int device, int64_t max_points, int max_w, int max_h, int max_batch, int64_t pair_capacity)
{
    if (!out || max_points < 1 || max_points > 0xFFFFFFF0ll || max_w < 1 || max_h < 1 || max_w > 65535 || max_h > 65535 || max_batch < 1 || max_batch > 64)
        return PCR_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return PCR_ERR_CUDA;
    pcr_ctx* ctx = new (std::nothrow) pcr_ctx();
    if (!ctx) return PCR_ERR_NOMEM;
    ctx->device = device; ctx->max_points = max_points; ctx->max_w = max_w; ctx->max_h = max_h; ctx->max_batch = max_batch;
    // a pair costs 24 bytes: the default leaves room for 24 tile entries per point up to 262144 points (large
    // projected spheres on small clouds), 12 above (+ padding of every tile's range to a multiple of 4); a frame that
    // needs more takes the unbinned raster (overflow path)
    ctx->pair_cap = pair_capacity > 0 ? pair_capacity : (max_points <= 262144 ? 24 : 12) * max_points + 65536;
    if (ctx->pair_cap > 0xFFFFFFF0ll) ctx->pair_cap = 0xFFFFFFF0ll;
    ctx->pair_cap = (ctx->pair_cap + 3) & ~3ll;
    ctx->tiles_cap = (((max_w + TILE - 1) / TILE) * ((max_h + TILE - 1) / TILE) + 3) & ~3;   // multiple of 4: k_scan_tiles uses uint4
    ctx->item_cap = ctx->tiles_cap + (int)(ctx->pair_cap / ITEM_SPHERES) + 1;
    {   // level 1 (8x4 pixel blocks) + level 2 (4x4 groups of them), see hiz_far_bits
        const int w1 = (max_w + HZ_W - 1) / HZ_W, h1 = (max_h + HZ_H - 1) / HZ_H;
        // level 1, level 2, and the coarse-cell table of k_project_cull4 (k_hiz2); a multiple of 4 words: every frame's table is 16-byte aligned
        ctx->hz_cap = hiz_table_offset(w1, h1) + 4 * ((w1 + 3) / 4 + 2) * ((h1 + 3) / 4 + 2);
    }
    if (const char* e = getenv("PCR_OCCLUSION")) ctx->occlusion = atoi(e);
    if (const char* e = getenv("PCR_TWO_PHASE")) ctx->two_phase = atoi(e);
    if (const char* e = getenv("PCR_CULL4")) ctx->cull4 = atoi(e);
    if (const char* e = getenv("PCR_MEAN_EXCLUSIVE")) ctx->mean_exclusive = atoi(e);
    if (const char* e = getenv("PCR_SCATTER_MERGE")) ctx->scatter_merge = atoi(e);
    if (const char* e = getenv("PCR_MEAN_FUSED")) ctx->mean_fused = atoi(e);
    if (const char* e = getenv("PCR_PREPASS_ZCUT")) ctx->prepass_zcut = (float)atof(e);
    if (const char* e = getenv("PCR_PREPASS_BACK")) { int v = std::max(1, atoi(e)); while (v & (v - 1)) v &= v - 1; ctx->prepass_back = v; }
    if (const char* e = getenv("PCR_LAZY_FILL")) ctx->lazy_fill = atoi(e);
    if (const char* e = getenv("PCR_SAMPLE_PREPASS")) ctx->sample_prepass = atoi(e);
    if (const char* e = getenv("PCR_STATS_AHEAD")) ctx->stats_ahead = atoi(e);
    if (const char* e = getenv("PCR_SCATTER_THREADS")) ctx->scatter_threads = std::min(BIN_THREADS, std::max(32, atoi(e) & ~31));
    if (const char* e = getenv("PCR_OCCLUSION_STEP")) ctx->occlusion_step = std::max(2, atoi(e));
    if (const char* e = getenv("PCR_OCCLUSION_LEVELS")) ctx->occlusion_levels = atoi(e);
    if (const char* e = getenv("PCR_HOST_TRACE")) ctx->host_trace = atoi(e);
    if (const char* e = getenv("PCR_OCCLUSION_STEP2")) ctx->occlusion_step2 = std::max(2, atoi(e));
    if (const char* e = getenv("PCR_OCCLUSION_RATIO")) ctx->occlusion_ratio = std::max(2, atoi(e));
    const size_t B = (size_t)max_batch, N = (size_t)max_points, Tn = (size_t)ctx->tiles_cap;
    cudaError_t e = cudaSetDevice(device);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) {
        ctx->num_sms = prop.multiProcessorCount;
        ctx->smem_optin = (int)std::min<size_t>(prop.sharedMemPerBlockOptin, 200 * 1024);
        ctx->smem_max = (int)prop.sharedMemPerBlockOptin;
        e = cudaFuncSetAttribute(k_project_count<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_project_count<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_project_count<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin);
        {
            const int mean_smem = (int)std::max<size_t>(MEAN_SMEM_BYTES, (size_t)ctx->smem_max);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mean_sequential<float, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mean_smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mean_sequential<float, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, mean_smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mean_sequential<double, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mean_smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mean_sequential<double, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, mean_smem);
        }
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_project_cull4, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_raster_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(RasterStage) * RASTER_STAGES));
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->raster_ctas_per_sm, k_raster_tiles, RASTER_CTA_THREADS, sizeof(RasterStage) * RASTER_STAGES);
        if (e == cudaSuccess && ctx->raster_ctas_per_sm < 1) e = cudaErrorLaunchOutOfResources;
    }
#define ALLOC(ptr, bytes) if (e == cudaSuccess) e = cudaMalloc((void**)&(ptr), (bytes))
    ALLOC(ctx->sph, sizeof(float4) * B * N);
    ALLOC(ctx->rect, sizeof(uint4) * B * N);
    ctx->gx_cap = (int)std::max<long long>(2 * ctx->num_sms * (2048 / BIN_THREADS), (max_points + BIN_THREADS * 4 - 1) / (BIN_THREADS * 4)) + 1;
    ALLOC(ctx->surv_count, sizeof(unsigned int) * B * (size_t)ctx->gx_cap);
    ALLOC(ctx->partials, sizeof(double) * B * MAX_STAT_BLOCKS * 9 * (PREP_SLOTS + 1));   // one region per prepared-batch slot + the in-line one
    ALLOC(ctx->stats, sizeof(double) * B * 10 * (PREP_SLOTS + 1));
    ALLOC(ctx->done, sizeof(unsigned int) * B * (PREP_SLOTS + 1));
    ALLOC(ctx->counts, sizeof(unsigned int) * B * Tn);
    ALLOC(ctx->offsets, sizeof(unsigned int) * B * (Tn + 4));   // per-frame stride tiles_cap + 4 keeps uint4 alignment
    ALLOC(ctx->cursor, sizeof(unsigned int) * B * Tn);
    ctx->scan_stripes = (int)((Tn + 4095) / 4096);
    ALLOC(ctx->scan_part, sizeof(unsigned long long) * B * (size_t)ctx->scan_stripes * 3);
    ALLOC(ctx->fill_list, sizeof(unsigned int) * B * Tn);
    ALLOC(ctx->fill_count, sizeof(unsigned int) * B);
    ALLOC(ctx->scan_ready, sizeof(unsigned int) * B * (size_t)ctx->scan_stripes);
    ALLOC(ctx->tile_state, sizeof(unsigned int) * (2 * B * Tn + B));     // [state | active list | active count]
    ALLOC(ctx->p_sph, sizeof(float4) * B * (size_t)ctx->pair_cap);
    ALLOC(ctx->p_ci, sizeof(uint2) * B * (size_t)ctx->pair_cap);
    ALLOC(ctx->overflow, sizeof(unsigned int) * B);
    ALLOC(ctx->stat_pairs, sizeof(unsigned long long) * (B + 16));
    ALLOC(ctx->item_count, sizeof(unsigned int) * B);
    ALLOC(ctx->item_next, sizeof(unsigned int) * B);
    ALLOC(ctx->items, sizeof(uint4) * B * (size_t)ctx->item_cap);
    ALLOC(ctx->hz, sizeof(unsigned int) * B * (size_t)ctx->hz_cap);
    ALLOC(ctx->d_frames, sizeof(FrameDev) * B);
#undef ALLOC
    if (e == cudaSuccess) e = cudaMallocHost((void**)&ctx->h_frames, sizeof(FrameDev) * B * RING_SLOTS);
    if (e == cudaSuccess) e = cudaMemset(ctx->counts, 0, sizeof(unsigned int) * B * Tn);
    if (e == cudaSuccess) e = cudaMemset(ctx->done, 0, sizeof(unsigned int) * B * (PREP_SLOTS + 1));
    if (e == cudaSuccess) e = cudaMemset(ctx->scan_ready, 0, sizeof(unsigned int) * B * (size_t)ctx->scan_stripes);
    if (e == cudaSuccess) e = cudaMemset(ctx->overflow, 0, sizeof(unsigned int) * B);
    if (e == cudaSuccess) e = cudaMemset(ctx->stat_pairs, 0, sizeof(unsigned long long) * (B + 16));
    for (int k = 0; k < RING_SLOTS && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&ctx->ring_ev[k], cudaEventDisableTiming);
    {
        // the prepared-batch slots' streams and events, all of them now: created on first use, a slot that had only served
        // in-line work so far cost its first hint a stream creation — 6 ms of idle GPU in the middle of a pipeline (measured)
        int lo = 0, hi = 0;
        if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
        for (pcr_ctx::PrepSlot& p : ctx->prep) {
            if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&p.stream, cudaStreamNonBlocking, hi);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p.ev_fork, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p.ev_ready, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p.ev_free, cudaEventDisableTiming);
        }
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        int code = e == cudaErrorMemoryAllocation ? PCR_ERR_NOMEM : PCR_ERR_CUDA;
        pcr_destroy(ctx);
        cudaGetLastError();
        return code;
    }
    *out = ctx;
    return PCR_OK;
}

void pcr_destroy(pcr_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    void* frees[] = {ctx->surv_count, ctx->sph, ctx->rect, ctx->partials, ctx->stats, ctx->done, ctx->counts, ctx->offsets,
                     ctx->cursor, ctx->p_sph, ctx->p_ci, ctx->tile_state, ctx->sample, ctx->scan_part, ctx->scan_ready, ctx->fill_list, ctx->fill_count, ctx->overflow, ctx->stat_pairs, ctx->item_count, ctx->item_next, ctx->items, ctx->hz, ctx->hz_b, ctx->lut, ctx->vis, ctx->d_frames, ctx->stage_radius, ctx->stage_rgb,
                     ctx->peer_merged, ctx->peer_image, ctx->mesh_verts, ctx->mesh_prof, ctx->plan, ctx->dstats, ctx->dxf, ctx->dctrl, ctx->dcount, ctx->dbin, ctx->dhist, ctx->dedges, ctx->dstarts, ctx->dcursor, ctx->dorder};
    for (void* p : frees) if (p) cudaFree(p);
    if (ctx->h_frames) cudaFreeHost(ctx->h_frames);
    for (int k = 0; k < RING_SLOTS; ++k) if (ctx->ring_ev[k]) cudaEventDestroy(ctx->ring_ev[k]);
    for (int k = 0; k < HOST_STAGES; ++k) {
        if (ctx->ev_h2d[k]) cudaEventDestroy(ctx->ev_h2d[k]);
        if (ctx->ev_comp[k]) cudaEventDestroy(ctx->ev_comp[k]);
        if (ctx->ev_d2h[k]) cudaEventDestroy(ctx->ev_d2h[k]);
        for (void* q : {(void*)ctx->stage_in[k], (void*)ctx->stage_rgba[k], (void*)ctx->stage_vis[k]}) if (q) cudaFree(q);
    }
    for (cudaEvent_t ev : ctx->ticket_ev) if (ev) cudaEventDestroy(ev);
    for (ProfRec& r : ctx->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
    for (pcr_ctx::PrepSlot& p : ctx->prep) {
        if (p.stream) cudaStreamDestroy(p.stream);
        for (cudaEvent_t ev : {p.ev_fork, p.ev_ready, p.ev_free}) if (ev) cudaEventDestroy(ev);
    }
    if (ctx->ev_last) cudaEventDestroy(ctx->ev_last);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_comp) cudaStreamDestroy(ctx->s_comp);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    delete ctx;
}

int pcr_stats_partial(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols, double* d_partial9, void* stream)
{
    pcr_style dummy; memset(&dummy, 0, sizeof(dummy));
    int rc = check_common(ctx, n, cols, &dummy);
    if (rc) return rc;
    if (!d_in || !d_partial9 || n < 1) return fail(ctx, PCR_ERR_INVALID, "pcr_stats_partial: NULL buffer or n < 1");
    CK(cudaSetDevice(ctx->device));
    if ((rc = enter(ctx, (cudaStream_t)stream))) return rc;
    int rc2 = launch_stats(ctx, d_in, in_is_f64, n, cols, 0, 1, INLINE_REGION, d_partial9, 2, (cudaStream_t)stream);
    if (rc2) return rc2;
    return leave(ctx, (cudaStream_t)stream);
}

int pcr_finalize_stats(pcr_ctx* ctx, const double* d_partials, int n_shards, int64_t n_total, int in_is_f64, double* d_stats10,
                       void* stream)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (!d_partials || !d_stats10 || n_shards < 1 || n_total < 1) return fail(ctx, PCR_ERR_INVALID, "pcr_finalize_stats: NULL buffer or empty cloud");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (in_is_f64)
        LAUNCH(KID_STATS, s, k_finalize_partials<double><<<1, 32, 0, s>>>(d_partials, n_shards, n_total, d_stats10));
    else
        LAUNCH(KID_STATS, s, k_finalize_partials<float><<<1, 32, 0, s>>>(d_partials, n_shards, n_total, d_stats10));
    return PCR_OK;
}

int pcr_standardize_with_stats(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols, const float* d_radius,
                               const float* d_rgb, const pcr_style* style, const double* d_stats10, float* d_pos_out,
                               float* d_attr_out, float* d_vel_out, void* stream)
{
    int rc = check_common(ctx, n, cols, style);
    if (rc) return rc;
    if (n == 0) return PCR_OK;
    if (!d_in || !d_stats10 || !d_pos_out || !d_attr_out) return fail(ctx, PCR_ERR_INVALID, "pcr_standardize: NULL buffer");
    CK(cudaSetDevice(ctx->device));
    return launch_transform(ctx, d_in, in_is_f64, n, cols, 0, 1, d_radius, d_rgb, d_stats10, to_style_dev(style),
                            (float4*)d_pos_out, (float4*)d_attr_out, (float4*)d_vel_out, 0, (cudaStream_t)stream);
}

int pcr_standardize(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols, const float* d_radius, const float* d_rgb,
                    const pcr_style* style, float* d_pos_out, float* d_attr_out, float* d_vel_out, double* d_stats, void* stream)
{
    int rc = check_common(ctx, n, cols, style);
    if (rc) return rc;
    if (n == 0) return PCR_OK;
    if (!d_in || !d_pos_out || !d_attr_out) return fail(ctx, PCR_ERR_INVALID, "pcr_standardize: NULL buffer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = enter(ctx, s))) return rc;
    double* stats = inline_stats(ctx);
    rc = launch_stats(ctx, d_in, in_is_f64, n, cols, 0, 1, INLINE_REGION, stats, 1, s, style->mean_mode);
    if (rc) return rc;
    rc = launch_transform(ctx, d_in, in_is_f64, n, cols, 0, 1, d_radius, d_rgb, stats, to_style_dev(style),
                          (float4*)d_pos_out, (float4*)d_attr_out, (float4*)d_vel_out, 0, s);
    if (rc) return rc;
    if (d_stats) CK(cudaMemcpyAsync(d_stats, stats, sizeof(double) * 10, cudaMemcpyDeviceToDevice, s));
    return leave(ctx, s);
}

int pcr_transform_coordinates(pcr_ctx* ctx, const float* d_in, int64_t n, int cols, int flip_x, float z_lift, float* d_out,
                              void* stream)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (n < 0 || (cols != 3 && cols != 6)) return fail(ctx, PCR_ERR_INVALID, "n < 0 or cols not 3|6");
    if (n == 0) return PCR_OK;
    if (!d_in || !d_out || d_in == d_out) return fail(ctx, PCR_ERR_INVALID, "pcr_transform_coordinates: NULL or aliased buffer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH(KID_AXIS, s, k_axis_transform<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_in, n, cols, flip_x, z_lift, d_out));
    return PCR_OK;
}

int pcr_velocity_trails(pcr_ctx* ctx, const float* d_pcl6, int64_t n, const pcr_style* style, double trail_scale, float* d_tail,
                        float* d_head, uint8_t* d_valid, void* stream)
{
    int rc = check_common(ctx, n, 6, style);
    if (rc) return rc;
    if (n == 0) return PCR_OK;
    if (!d_pcl6 || !d_tail || !d_head || !d_valid) return fail(ctx, PCR_ERR_INVALID, "pcr_velocity_trails: NULL buffer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH(KID_AXIS, s, k_trail_ends<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_pcl6, n, to_style_dev(style), trail_scale, d_tail, d_head, d_valid));
    return PCR_OK;
}

int pcr_render(pcr_ctx* ctx, const float* d_pos, const float* d_attr, int64_t n, uint32_t id_base, const pcr_camera* cam,
               const pcr_style* style, uint64_t* d_vis, uint8_t* d_rgba, void* stream)
{
    int rc = check_common(ctx, n, 3, style);
    if (rc) return rc;
    if (!cam || !d_vis || (n > 0 && !d_pos) || (d_rgba && n > 0 && !d_attr)) return fail(ctx, PCR_ERR_INVALID, "pcr_render: NULL buffer");
    if ((unsigned long long)id_base + (unsigned long long)n > 0xFFFFFFFEull) return fail(ctx, PCR_ERR_INVALID, "point id overflow");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = enter(ctx, s))) return rc;
    rc = upload_frames(ctx, cam, 1, s);
    if (rc) return rc;
    const long long px = (long long)cam->width * cam->height;
    rc = launch_render(ctx, (const float4*)d_pos, (const float4*)d_attr, 0, nullptr, n, 1, id_base, to_style_dev(style), cam->width,
                       cam->height, d_vis, px, d_rgba, px, 0, s);
    if (rc) return rc;
    return leave(ctx, s);
}

int pcr_render_transformed(pcr_ctx* ctx, const float* d_pcl, int64_t n, int cols, const float* d_radius, const float* d_rgb,
                           const pcr_camera* cam, const pcr_style* style, uint64_t* d_vis, uint8_t* d_rgba, void* stream)
{
    int rc = check_common(ctx, n, cols, style);
    if (rc) return rc;
    if (!cam || !d_rgba || (n > 0 && !d_pcl)) return fail(ctx, PCR_ERR_INVALID, "pcr_render_transformed: NULL buffer");
    if (style->trails != 0 && style->trails != 1) return fail(ctx, PCR_ERR_INVALID, "pcr_render_transformed: trails must be 0 or 1");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (!d_vis) { rc = ensure_vis(ctx); if (rc) return rc; }
    if ((rc = enter(ctx, s))) return rc;
    if ((rc = upload_frames(ctx, cam, 1, s))) return rc;
    // the frame is used as it is: centre 0 and scale 1 make K1's (x - c) / s exact, xform = 1 skips the axis transform;
    // the cloud's min / max (position colormap) are still taken from the data
    double* stats = inline_stats(ctx);
    if (n > 0) {
        rc = launch_stats(ctx, d_pcl, 0, n, cols, 0, 1, INLINE_REGION, stats, 1, s, PCR_MEAN_F64);
        if (rc) return rc;
    }
    LAUNCH(KID_STATS, s, k_stats_identity<<<1, 32, 0, s>>>(stats, n > 0 ? 1 : 0));
    StyleDev st = to_style_dev(style);
    st.xform = 1;
    const RawSrc raw = {d_pcl, 0, n * cols, cols, stats, d_radius, d_rgb};
    const long long px = (long long)cam->width * cam->height;
    rc = launch_render(ctx, nullptr, nullptr, 0, &raw, n, 1, 0u, st, cam->width, cam->height, d_vis ? d_vis : ctx->vis, px, d_rgba, px, 0, s);
    if (rc) return rc;
    return leave(ctx, s);
}

int pcr_shade(pcr_ctx* ctx, const uint64_t* d_vis, const float* d_pos, const float* d_attr, int64_t n, uint32_t id_base,
              int owner_only, const pcr_camera* cam, const pcr_style* style, uint8_t* d_rgba, void* stream)
{
    int rc = check_common(ctx, n, 3, style);
    if (rc) return rc;
    if (!cam || !d_vis || !d_rgba || (n > 0 && (!d_pos || !d_attr))) return fail(ctx, PCR_ERR_INVALID, "pcr_shade: NULL buffer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = enter(ctx, s))) return rc;
    rc = upload_frames(ctx, cam, 1, s);
    if (rc) return rc;
    const long long px = (long long)cam->width * cam->height;
    rc = launch_shade(ctx, to_style_dev(style), d_vis, px, (const float4*)d_pos, (const float4*)d_attr, 0, nullptr, n, 1, id_base, owner_only,
                      cam->width, cam->height, d_rgba, px, s);
    if (rc) return rc;
    return leave(ctx, s);
}

int pcr_render_shard(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols, const float* d_radius, const float* d_rgb,
                     const double* d_stats10, uint32_t id_base, const pcr_camera* cam, const pcr_style* style, uint64_t* d_vis,
                     void* stream)
{
    int rc = check_common(ctx, n, cols, style);
    if (rc) return rc;
    if (!cam || !d_vis || !d_stats10 || (n > 0 && !d_in)) return fail(ctx, PCR_ERR_INVALID, "pcr_render_shard: NULL buffer");
    if ((unsigned long long)id_base + (unsigned long long)n > 0xFFFFFFFEull) return fail(ctx, PCR_ERR_INVALID, "point id overflow");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = enter(ctx, s))) return rc;
    rc = upload_frames(ctx, cam, 1, s);
    if (rc) return rc;
    StyleDev st = to_style_dev(style);
    st.trails = 0;                                       // trail ids (n + i) are not defined across shards
    const RawSrc raw = {d_in, in_is_f64, n * cols, cols, d_stats10, d_radius, d_rgb};
    const long long px = (long long)cam->width * cam->height;
    rc = launch_render(ctx, nullptr, nullptr, 0, &raw, n, 1, id_base, st, cam->width, cam->height, d_vis, px, nullptr, px, 0, s);
    if (rc) return rc;
    return leave(ctx, s);
}

int pcr_shade_shard(pcr_ctx* ctx, const uint64_t* d_vis, const void* d_in, int in_is_f64, int64_t n, int cols, const float* d_radius,
                    const float* d_rgb, const double* d_stats10, uint32_t id_base, int owner_only, const pcr_camera* cam,
                    const pcr_style* style, uint8_t* d_rgba, void* stream)
{
    int rc = check_common(ctx, n, cols, style);
    if (rc) return rc;
    if (!cam || !d_vis || !d_rgba || !d_stats10 || (n > 0 && !d_in)) return fail(ctx, PCR_ERR_INVALID, "pcr_shade_shard: NULL buffer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = enter(ctx, s))) return rc;
    rc = upload_frames(ctx, cam, 1, s);
    if (rc) return rc;
    StyleDev st = to_style_dev(style);
    st.trails = 0;
    const RawSrc raw = {d_in, in_is_f64, n * cols, cols, d_stats10, d_radius, d_rgb};
    const long long px = (long long)cam->width * cam->height;
    rc = launch_shade(ctx, st, d_vis, px, nullptr, nullptr, 0, &raw, n, 1, id_base, owner_only, cam->width, cam->height, d_rgba, px, s);
    if (rc) return rc;
    return leave(ctx, s);
}

int pcr_render_frames(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols, int n_frames, const float* d_radius,
                      const float* d_rgb, const pcr_camera* cams, const pcr_style* style, uint64_t* d_vis, uint8_t* d_rgba,
                      void* stream)
{
    int rc = check_common(ctx, n, cols, style);
    if (rc) return rc;
    if (n_frames < 0 || !cams || !d_rgba || (n > 0 && !d_in)) return fail(ctx, PCR_ERR_INVALID, "pcr_render_frames: NULL buffer");
    if (n < 1) return fail(ctx, PCR_ERR_INVALID, "pcr_render_frames: empty frames cannot be standardised");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    const StyleDev st = to_style_dev(style);
    const int W = cams[0].width, H = cams[0].height;
    for (int f = 0; f < n_frames; ++f)
        if (cams[f].width != W || cams[f].height != H) return fail(ctx, PCR_ERR_INVALID, "all frames of a call must share W x H");
    const long long px = (long long)W * H;
    const size_t elem = in_is_f64 ? 8 : 4;
    const long long frame_stride = n * cols;
    if (!d_vis) { rc = ensure_vis(ctx); if (rc) return rc; }
    if ((rc = enter(ctx, s))) return rc;
    const int B = ctx->max_batch;
    const int nbatches = (n_frames + B - 1) / B;
    // K0 — and the serial reference-exact mean, 2 ms of pure latency per million points — of the next batches runs on
    // side streams while a batch renders: every batch is looked up among the prepared slots first (an earlier
    // pcr_prefetch_frames, pcr_render_frames_host, or this loop's own look-ahead), and only computed in line when nobody did.
    const bool ahead = ctx->stats_ahead != 0;
    const SamplePlan sp = sample_plan(ctx, n);
    if ((rc = ensure_sample(ctx, sp, elem))) return rc;
    auto batch_in = [&](int k) { return (const char*)d_in + (size_t)k * B * frame_stride * elem; };
    auto batch_nb = [&](int k) { return std::min(B, n_frames - k * B); };
    auto look_ahead = [&](int from) -> int {             // keep up to PREP_SLOTS - 1 later batches in preparation
        for (int j = from; ahead && j < std::min(nbatches, from + PREP_SLOTS - 1); ++j) {
            if (find_prepared(ctx, batch_in(j), in_is_f64, n, cols, batch_nb(j), style->mean_mode, sp) >= 0) continue;
            int free_slots = 0;
            for (const pcr_ctx::PrepSlot& p : ctx->prep) free_slots += !p.valid;
            if (!free_slots) break;                      // never evict what this very call still needs
            int slot;
            int rc2 = prepare_batch(ctx, batch_in(j), in_is_f64, n, cols, batch_nb(j), style->mean_mode, sp, s, true, &slot);
            if (rc2) return rc2;
        }
        return PCR_OK;
    };
    for (int k = 0; k < nbatches; ++k) {
        const int f0 = k * B;
        const int nb = batch_nb(k);
        rc = upload_frames(ctx, cams + f0, nb, s);
        if (rc) return rc;
        int slot = find_prepared(ctx, batch_in(k), in_is_f64, n, cols, nb, style->mean_mode, sp);
        if (slot < 0) {
            if (getenv("PCR_DEBUG_PREP")) fprintf(stderr, "[pcr] render_frames: batch %p not prepared, K0 in line\n", (const void*)batch_in(k));
            // nobody prepared this batch: on a side stream when more batches follow (their K0 then overlaps this one's
            // render), in line otherwise
            rc = prepare_batch(ctx, batch_in(k), in_is_f64, n, cols, nb, style->mean_mode, sp, s, ahead && nbatches > 1, &slot);
            if (rc) return rc;
        }
        if ((rc = look_ahead(k + 1))) return rc;
        pcr_ctx::PrepSlot& ps = ctx->prep[slot];
        CK(cudaStreamWaitEvent(s, ps.ev_ready, 0));
        // no K1 launch: K2a and K4 evaluate standardise/transform/colour from the raw frames on the fly
        RawSrc raw = {batch_in(k), in_is_f64, frame_stride, cols, slot_stats(ctx, slot), d_radius, d_rgb};
        raw.sample = slot_sample(ctx, slot, sp, elem); raw.sample_stride = sp.stride;
        uint64_t* vis = d_vis ? d_vis + (size_t)f0 * px : ctx->vis;
        const long long vis_stride = d_vis ? px : (long long)ctx->max_w * ctx->max_h;
        rc = launch_render(ctx, nullptr, nullptr, 0, &raw, n, nb, 0u, st, W, H, vis, vis_stride,
                           d_rgba + (size_t)f0 * px * 4, px, 0, s);
        if (rc) return rc;
        CK(cudaEventRecord(ps.ev_free, s));
        ps.valid = false; ps.used = true; ps.stamp = ++ctx->prep_stamp;
    }
    return leave(ctx, s);
}

int pcr_prefetch_frames(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols, int n_frames, const pcr_style* style, void* stream)
{
    int rc = check_common(ctx, n, cols, style);
    if (rc) return rc;
    if (n_frames < 0 || (n_frames > 0 && (!d_in || n < 1))) return fail(ctx, PCR_ERR_INVALID, "pcr_prefetch_frames: NULL buffer or empty frames");
    if (n_frames == 0) {                                  // drop every hint
        for (pcr_ctx::PrepSlot& p : ctx->prep) p.valid = false;
        return PCR_OK;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t elem = in_is_f64 ? 8 : 4;
    const SamplePlan sp = sample_plan(ctx, n);
    if ((rc = ensure_sample(ctx, sp, elem))) return rc;
    const int B = ctx->max_batch;
    const int nbatches = std::min((n_frames + B - 1) / B, PREP_SLOTS);       // a hint: what does not fit is computed later
    for (int k = 0; k < nbatches; ++k) {
        const char* in = (const char*)d_in + (size_t)k * B * (size_t)(n * cols) * elem;
        const int nb = std::min(B, n_frames - k * B);
        if (find_prepared(ctx, in, in_is_f64, n, cols, nb, style->mean_mode, sp) >= 0) continue;
        int free_slots = 0;
        for (const pcr_ctx::PrepSlot& p : ctx->prep) free_slots += !p.valid;
        if (free_slots <= 1) {                                 // a hint never displaces an earlier one (those frames come first), and one slot stays free for un-hinted work
            if (getenv("PCR_DEBUG_PREP")) fprintf(stderr, "[pcr] prefetch_frames: hint %p dropped (no free slot)\n", (const void*)in);
            break;
        }
        if (getenv("PCR_DEBUG_PREP")) fprintf(stderr, "[pcr] prefetch_frames: hint %p prepared\n", (const void*)in);
        int slot;
        if ((rc = prepare_batch(ctx, in, in_is_f64, n, cols, nb, style->mean_mode, sp, s, ctx->stats_ahead != 0, &slot))) return rc;
    }
    return PCR_OK;
}

int pcr_render_frames_host_submit(pcr_ctx* ctx, const void* h_in, int in_is_f64, int64_t n, int cols, int n_frames, const float* h_radius,
                                  const float* h_rgb, const pcr_camera* cams, const pcr_style* style, uint64_t* h_vis, uint8_t* h_rgba,
                                  int64_t* ticket)
{
    int rc = check_common(ctx, n, cols, style);
    if (rc) return rc;
    if (n_frames < 0 || !cams || !h_rgba || !h_in || n < 1) return fail(ctx, PCR_ERR_INVALID, "pcr_render_frames_host: NULL buffer or n < 1");
    CK(cudaSetDevice(ctx->device));
    const int W = n_frames ? cams[0].width : 1, H = n_frames ? cams[0].height : 1;
    if (W > ctx->max_w || H > ctx->max_h) return fail(ctx, PCR_ERR_CAPACITY, "frame larger than the context");
    const size_t px = (size_t)W * H, elem = in_is_f64 ? 8 : 4;
    const size_t frame_bytes = (size_t)n * cols * elem;
    const int B = ctx->max_batch;
    if (!ctx->s_h2d) {
        CK(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_comp, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
        for (int k = 0; k < HOST_STAGES; ++k) {
            CK(cudaEventCreateWithFlags(&ctx->ev_h2d[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_comp[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_d2h[k], cudaEventDisableTiming));
            CK(cudaMalloc((void**)&ctx->stage_rgba[k], (size_t)B * ctx->max_w * ctx->max_h * 4));
        }
        for (int k = 0; k < HOST_TICKETS; ++k) CK(cudaEventCreateWithFlags(&ctx->ticket_ev[k], cudaEventDisableTiming));
        CK(cudaMalloc((void**)&ctx->stage_radius, sizeof(float) * ctx->max_points));
        CK(cudaMalloc((void**)&ctx->stage_rgb, sizeof(float) * 3 * ctx->max_points));
    }
    if (h_vis && !ctx->stage_vis[0]) {
        CK(cudaDeviceSynchronize());
        for (int k = 0; k < HOST_STAGES; ++k) CK(cudaMalloc((void**)&ctx->stage_vis[k], (size_t)B * ctx->max_w * ctx->max_h * 8));
    }
    if (ctx->stage_in_bytes < (size_t)B * frame_bytes) {
        CK(cudaDeviceSynchronize());
        for (pcr_ctx::PrepSlot& p : ctx->prep) p.valid = false;          // hints on the old staging buffers are void
        for (int k = 0; k < HOST_STAGES; ++k) {
            if (ctx->stage_in[k]) CK(cudaFree(ctx->stage_in[k]));
            ctx->stage_in[k] = nullptr;
            CK(cudaMalloc(&ctx->stage_in[k], (size_t)B * frame_bytes));
        }
        ctx->stage_in_bytes = (size_t)B * frame_bytes;
    }
    // Work that OTHER entry points left on a caller's stream may still be using the scratch: the kernels wait for it.
    // (Earlier host-buffer calls are ordered by the staging slots' own events — the copies of this call overlap the
    // kernels of the previous one.)
    if ((rc = enter(ctx, ctx->s_comp))) return rc;
    const float *d_radius = nullptr, *d_rgb = nullptr;
    if (h_radius) { CK(cudaMemcpyAsync(ctx->stage_radius, h_radius, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->s_comp)); d_radius = ctx->stage_radius; }
    if (h_rgb) { CK(cudaMemcpyAsync(ctx->stage_rgb, h_rgb, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, ctx->s_comp)); d_rgb = ctx->stage_rgb; }
    // Chunks of at most B frames, but at least ~4 chunks per call so that the H2D copy of chunk k+1, the K0 (+ serial
    // mean) of chunk k, the kernels of chunk k-1 and the D2H copy of chunk k-2 overlap even for short calls.
    // With the parallel float64 mean a chunk's work is proportional to its size, so the tail of the call is cut into ever
    // smaller chunks (C, ..., C, C/2, C/4, ..., 1): what is left to do after the last byte has arrived shrinks with it.
    // With the reference's serial mean every chunk carries the same ~3-4 ms of latency however small it is (measured:
    // PCR_HOST_TRACE=1) and small chunks only use up staging slots: equal chunks, a quarter of the call each.
    const bool serial_mean = style->mean_mode != PCR_MEAN_F64;
    int C = std::max(1, std::min(B, (n_frames + 3) / 4));
    if (serial_mean) C = std::min(C, 8);            // (PREP_SLOTS chunks of 8 frames in flight outrun the H2D copy by far)
    for (int f0 = 0, nb = 0; f0 < n_frames; f0 += nb) {
        const int left = n_frames - f0;
        nb = serial_mean ? std::min(C, left) : (left > C ? C : std::max(1, left / 2));
        const unsigned long long chunk = ctx->host_chunks++;
        const int k = (int)(chunk % HOST_STAGES);
        const bool reused = chunk >= (unsigned long long)HOST_STAGES;
        if (reused) CK(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_comp[k], 0));       // input slot free again
        pcr_ctx::TraceRec tr = {chunk, nb, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        if (ctx->host_trace) {
            for (cudaEvent_t* ev : {&tr.h2d0, &tr.h2d1, &tr.ready, &tr.comp0, &tr.comp1, &tr.d2h1}) CK(cudaEventCreate(ev));
            CK(cudaEventRecord(tr.h2d0, ctx->s_h2d));
        }
        CK(cudaMemcpyAsync(ctx->stage_in[k], (const char*)h_in + (size_t)f0 * frame_bytes, (size_t)nb * frame_bytes,
                           cudaMemcpyHostToDevice, ctx->s_h2d));
        CK(cudaEventRecord(ctx->ev_h2d[k], ctx->s_h2d));
        if (ctx->host_trace) CK(cudaEventRecord(tr.h2d1, ctx->s_h2d));
        // K0 (+ the serial reference-exact mean) of this chunk starts the moment its bytes have landed, on a side
        // stream: it overlaps the previous chunks' kernels instead of delaying this chunk's
        if ((rc = pcr_prefetch_frames(ctx, ctx->stage_in[k], in_is_f64, n, cols, nb, style, ctx->s_h2d))) return rc;
        CK(cudaStreamWaitEvent(ctx->s_comp, ctx->ev_h2d[k], 0));
        if (reused) CK(cudaStreamWaitEvent(ctx->s_comp, ctx->ev_d2h[k], 0));      // output slot drained
        if (ctx->host_trace) {
            // (the prepared slot's ready event, seen from its own stream, and the start of this chunk's kernels)
            for (pcr_ctx::PrepSlot& p : ctx->prep) if (p.valid && p.in == ctx->stage_in[k]) CK(cudaEventRecord(tr.ready, p.stream ? p.stream : ctx->s_comp));
            CK(cudaEventRecord(tr.comp0, ctx->s_comp));
        }
        rc = pcr_render_frames(ctx, ctx->stage_in[k], in_is_f64, n, cols, nb, d_radius, d_rgb, cams + f0, style,
                               h_vis ? ctx->stage_vis[k] : nullptr, ctx->stage_rgba[k], ctx->s_comp);
        if (rc) return rc;
        CK(cudaEventRecord(ctx->ev_comp[k], ctx->s_comp));
        if (ctx->host_trace) CK(cudaEventRecord(tr.comp1, ctx->s_comp));
        CK(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_comp[k], 0));
        CK(cudaMemcpyAsync(h_rgba + (size_t)f0 * px * 4, ctx->stage_rgba[k], (size_t)nb * px * 4, cudaMemcpyDeviceToHost, ctx->s_d2h));
        if (h_vis) CK(cudaMemcpyAsync(h_vis + (size_t)f0 * px, ctx->stage_vis[k], (size_t)nb * px * 8, cudaMemcpyDeviceToHost, ctx->s_d2h));
        CK(cudaEventRecord(ctx->ev_d2h[k], ctx->s_d2h));
        if (ctx->host_trace) { CK(cudaEventRecord(tr.d2h1, ctx->s_d2h)); ctx->trace.push_back(tr); }
    }
    const long long t = ctx->tickets++;
    CK(cudaEventRecord(ctx->ticket_ev[t % HOST_TICKETS], ctx->s_d2h));
    if (ticket) *ticket = t;
    return PCR_OK;
}

int pcr_host_wait(pcr_ctx* ctx, int64_t ticket)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (!ctx->s_d2h || ctx->tickets == 0) return PCR_OK;
    CK(cudaSetDevice(ctx->device));
    if (ticket < 0 || ticket >= ctx->tickets) {              // everything submitted so far
        CK(cudaStreamSynchronize(ctx->s_d2h));
        CK(cudaStreamSynchronize(ctx->s_comp));
        if (ctx->host_trace && !ctx->trace.empty()) {
            cudaDeviceSynchronize();
            const cudaEvent_t t0 = ctx->trace.front().h2d0;
            fprintf(stderr, "[pcr host trace] chunk frames | h2d start end | stats+mean ready | kernels start end | d2h end   (ms)\n");
            for (pcr_ctx::TraceRec& r : ctx->trace) {
                float a = 0, b = 0, c = -1, d = 0, e = 0, f = 0;
                cudaEventElapsedTime(&a, t0, r.h2d0); cudaEventElapsedTime(&b, t0, r.h2d1);
                if (cudaEventQuery(r.ready) == cudaSuccess) cudaEventElapsedTime(&c, t0, r.ready);
                cudaGetLastError();
                cudaEventElapsedTime(&d, t0, r.comp0); cudaEventElapsedTime(&e, t0, r.comp1); cudaEventElapsedTime(&f, t0, r.d2h1);
                fprintf(stderr, "[pcr host trace] %5llu %3d | %8.3f %8.3f | %8.3f | %8.3f %8.3f | %8.3f\n", r.chunk, r.nb, a, b, c, d, e, f);
            }
            for (pcr_ctx::TraceRec& r : ctx->trace)
                for (cudaEvent_t ev : {r.h2d0, r.h2d1, r.ready, r.comp0, r.comp1, r.d2h1}) if (ev && ev != t0) cudaEventDestroy(ev);
            cudaEventDestroy(t0);
            ctx->trace.clear();
        }
        return PCR_OK;
    }
    // a ticket older than the ring has been overwritten by a later call's event: waiting for that is conservative
    CK(cudaEventSynchronize(ctx->ticket_ev[ticket % HOST_TICKETS]));
    return PCR_OK;
}

int pcr_render_frames_host(pcr_ctx* ctx, const void* h_in, int in_is_f64, int64_t n, int cols, int n_frames, const float* h_radius,
                           const float* h_rgb, const pcr_camera* cams, const pcr_style* style, uint64_t* h_vis, uint8_t* h_rgba)
{
    int64_t ticket = -1;
    int rc = pcr_render_frames_host_submit(ctx, h_in, in_is_f64, n, cols, n_frames, h_radius, h_rgb, cams, style, h_vis, h_rgba, &ticket);
    if (rc) return rc;
    return pcr_host_wait(ctx, -1);
}

int pcr_zmin(pcr_ctx* ctx, uint64_t* d_dst, const uint64_t* d_src, int64_t n_px, void* stream)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (!d_dst || !d_src || n_px < 0) return fail(ctx, PCR_ERR_INVALID, "pcr_zmin: NULL buffer");
    if (n_px == 0) return PCR_OK;
    CK(cudaSetDevice(ctx->device));
    int blocks = (int)std::min<long long>((n_px + 255) / 256, (long long)ctx->num_sms * 16);
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH(KID_ZMIN, s, k_zmin<<<blocks, 256, 0, s>>>((unsigned long long*)d_dst, (const unsigned long long*)d_src, n_px));
    return PCR_OK;
}

int pcr_zmerge_nccl(pcr_ctx* ctx, uint64_t* d_vis, int64_t n_px, void* comm, void* stream)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (!d_vis || !comm || n_px < 0) return fail(ctx, PCR_ERR_INVALID, "pcr_zmerge_nccl: NULL buffer or communicator");
    // ncclResult_t ncclAllReduce(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)
    typedef int (*allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
    static allreduce_fn fn = nullptr;
    if (!fn) {
        fn = (allreduce_fn)dlsym(RTLD_DEFAULT, "ncclAllReduce");
        if (!fn) {
            void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (h) fn = (allreduce_fn)dlsym(h, "ncclAllReduce");
        }
        if (!fn) return fail(ctx, PCR_ERR_NCCL, "ncclAllReduce not found in this process");
    }
    const int kNcclUint64 = 5, kNcclMin = 3;   // nccl.h: ncclUint64 = 5, ncclMin = 3
    int r = fn(d_vis, d_vis, (size_t)n_px, kNcclUint64, kNcclMin, comm, (cudaStream_t)stream);
    if (r != 0) return fail(ctx, PCR_ERR_NCCL, "ncclAllReduce(uint64, min) failed");
    return PCR_OK;
}

int pcr_counters(pcr_ctx* ctx, int64_t out[4], void* stream)
{
    if (!ctx || !out) return PCR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    unsigned long long pairs = 0;
    std::vector<unsigned int> ov(ctx->max_batch);
    CK(cudaMemcpy(&pairs, ctx->stat_pairs, sizeof(pairs), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ov.data(), ctx->overflow, sizeof(unsigned int) * ctx->max_batch, cudaMemcpyDeviceToHost));
#ifdef PCR_RASTER_STATS
    {
        unsigned long long dbg[16];
        CK(cudaMemcpy(dbg, ctx->stat_pairs, sizeof(dbg), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[pcr stats] warp-candidates %llu  cull-iterations %llu  box-pass %llu  droplets tested %llu culled %llu\n", dbg[8], dbg[9], dbg[10], dbg[11], dbg[12]);
        CK(cudaMemset(ctx->stat_pairs + 8, 0, 8 * sizeof(unsigned long long)));
    }
#endif
    long long nov = 0;
    for (unsigned int v : ov) nov += v != 0;
    out[0] = ctx->launches; out[1] = (int64_t)pairs; out[2] = nov; out[3] = 0;
    return PCR_OK;
}

int pcr_profile(pcr_ctx* ctx, int enable)
{
    if (!ctx) return PCR_ERR_INVALID;
    ctx->profiling = enable != 0;
    return PCR_OK;
}

int pcr_profile_read(pcr_ctx* ctx, double* ms_out, int64_t* count_out, int capacity)
{
    if (!ctx || !ms_out || !count_out || capacity < KID_COUNT) return ctx ? fail(ctx, PCR_ERR_INVALID, "pcr_profile_read: buffers too small") : PCR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    for (int k = 0; k < capacity; ++k) { ms_out[k] = 0.0; count_out[k] = 0; }
    for (ProfRec& r : ctx->prof) {
        float ms = 0.0f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            ms_out[r.kid] += (double)ms;
            count_out[r.kid] += 1;
        } else {
            cudaGetLastError();
        }
        ctx->prof_pool.push_back(r.a);
        ctx->prof_pool.push_back(r.b);
    }
    ctx->prof.clear();
    return KID_COUNT;
}

int pcr_set_occlusion(pcr_ctx* ctx, int mode, int step, int64_t min_points)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (mode < -1 || mode > 1 || (step != 0 && step != -1 && step < 2)) return fail(ctx, PCR_ERR_INVALID, "pcr_set_occlusion: mode in {-1,0,1}, step >= 2 (0 keep, -1 automatic)");
    ctx->occlusion = mode;
    if (step == -1) ctx->occlusion_step = 0;                        // by the size of the cloud (prepass_step)
    else if (step) { ctx->occlusion_step = step; ctx->occlusion_step2 = std::max(2, step / 2); }
    if (min_points > 0) ctx->occlusion_min_points = min_points;
    return PCR_OK;
}


// ------------------------------------------------------------------------------------------
// Droplet scene (traj_renderer.py / traj_vel_renderer.py)
// ------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

// The reference's sampling plan for a history of h frames (traj_renderer.py:272-316): samples_per_segment =
// max(2, 20 // (h-1)) samples on every segment, thinned with linspace(0, total-1, 20).astype(int) when that
// gives more than 20, padded with the last one when fewer.  t, t*t, (t*t)*t are python floats there and meet
// float32 arrays, so they are rounded to float32 once.
void build_trail_plan(TrailPlan* plan)
{
    memset(plan, 0, sizeof(*plan));
    for (int k = 0; k < TRAIL_SAMPLES; ++k) {
        const double t = (double)k / (double)(TRAIL_SAMPLES - 1);
        plan->s[2][k] = TrailSample{0, (float)t, (float)(1.0 - t), 0.0f};
    }
    for (int h = 3; h <= HISTORY_FRAMES; ++h) {
        const int n_seg = h - 1, sps = std::max(2, TRAIL_SAMPLES / n_seg), total = n_seg * sps;
        auto sample = [&](int idx) {
            const int seg = idx / sps, i = idx % sps;
            const double t = (double)i / (double)(sps - 1);
            return TrailSample{seg, (float)t, (float)(t * t), (float)((t * t) * t)};
        };
        for (int k = 0; k < TRAIL_SAMPLES; ++k) {
            int idx;
            if (total > TRAIL_SAMPLES) {
                const double step = (double)(total - 1) / (double)(TRAIL_SAMPLES - 1);
                idx = k == TRAIL_SAMPLES - 1 ? total - 1 : (int)((double)k * step);
            } else {
                idx = std::min(k, total - 1);
            }
            plan->s[h][k] = sample(idx);
        }
    }
}

int ensure_plan(pcr_ctx* ctx)
{
    if (ctx->plan) return PCR_OK;
    TrailPlan* host = new (std::nothrow) TrailPlan;
    if (!host) return fail(ctx, PCR_ERR_NOMEM, "trail plan");
    build_trail_plan(host);
    cudaError_t e = cudaMalloc((void**)&ctx->plan, sizeof(TrailPlan));
    if (e == cudaSuccess) e = cudaMemcpy(ctx->plan, host, sizeof(TrailPlan), cudaMemcpyHostToDevice);
    delete host;
    if (e != cudaSuccess) return fail(ctx, PCR_ERR_CUDA, "trail plan upload", e);
    return PCR_OK;
}

DropletMeshDev mesh_dev(const pcr_ctx* ctx)
{
    DropletMeshDev m;
    m.verts = ctx->mesh_verts; m.prof = ctx->mesh_prof; m.bound = ctx->mesh_bound;
    m.n_rings = ctx->mesh_rings; m.n_segs = ctx->mesh_segs; m.nv = (ctx->mesh_rings + 1) * ctx->mesh_segs;
    return m;
}

}  // namespace

extern "C" {

int pcr_set_droplet_mesh(pcr_ctx* ctx, const float* h_verts, int n_rings, int n_segments)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (!h_verts || n_rings < 1 || n_rings > DROP_MAX_RINGS || n_segments < 3 || n_segments > DROP_MAX_SEGS ||
        (n_rings + 1) * n_segments > DROP_MAX_VERTS)
        return fail(ctx, PCR_ERR_INVALID, "pcr_set_droplet_mesh: bad ring / segment count");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());                       // a previous mesh may still be in use
    const int nv = (n_rings + 1) * n_segments;
    // ring profile (vertex 0 of every ring lies in the xz half-plane) and its smooth normals: the normalised
    // sum of the two adjacent band normals, the poles along the axis (same as oracle/droplet_oracle.py:ring_profile)
    std::vector<double> pr(n_rings + 1), pz(n_rings + 1), bnr(n_rings), bnz(n_rings);
    for (int i = 0; i <= n_rings; ++i) { pr[i] = h_verts[(size_t)i * n_segments * 3]; pz[i] = h_verts[(size_t)i * n_segments * 3 + 2]; }
    for (int i = 0; i < n_rings; ++i) {
        if (!(pz[i + 1] < pz[i])) return fail(ctx, PCR_ERR_INVALID, "pcr_set_droplet_mesh: ring z must decrease strictly");
        const double dr = pr[i + 1] - pr[i], dz = pz[i + 1] - pz[i], l = sqrt(dz * dz + dr * dr);
        bnr[i] = l > 0.0 ? -dz / l : 0.0; bnz[i] = l > 0.0 ? dr / l : 0.0;
    }
    std::vector<float4> prof(n_rings + 1);
    for (int i = 0; i <= n_rings; ++i) {
        double nr = 0.0, nz = i == 0 ? 1.0 : -1.0;
        if (i > 0 && i < n_rings) { nr = bnr[i - 1] + bnr[i]; nz = bnz[i - 1] + bnz[i]; }
        const double l = sqrt(nr * nr + nz * nz);
        if (l > 0.0) { nr /= l; nz /= l; }
        prof[i] = make_float4((float)pr[i], (float)pz[i], (float)nr, (float)nz);
    }
    double zlo = 1e300, zhi = -1e300, r2 = 0.0;
    for (int v = 0; v < nv; ++v) { zlo = std::min(zlo, (double)h_verts[3 * v + 2]); zhi = std::max(zhi, (double)h_verts[3 * v + 2]); }
    const double zc = 0.5 * (zlo + zhi);
    for (int v = 0; v < nv; ++v) {
        const double x = h_verts[3 * v], y = h_verts[3 * v + 1], z = h_verts[3 * v + 2] - zc;
        r2 = std::max(r2, x * x + y * y + z * z);
    }
    ctx->mesh_bound = make_float4(0.f, 0.f, (float)zc, (float)sqrt(r2));
    for (void* p : {(void*)ctx->mesh_verts, (void*)ctx->mesh_prof}) if (p) cudaFree(p);
    ctx->mesh_verts = nullptr; ctx->mesh_prof = nullptr; ctx->mesh_rings = 0;
    CK(cudaMalloc((void**)&ctx->mesh_verts, sizeof(float) * 3 * nv));
    CK(cudaMalloc((void**)&ctx->mesh_prof, sizeof(float4) * (n_rings + 1)));
    CK(cudaMemcpy(ctx->mesh_verts, h_verts, sizeof(float) * 3 * nv, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->mesh_prof, prof.data(), sizeof(float4) * (n_rings + 1), cudaMemcpyHostToDevice));
    ctx->mesh_rings = n_rings; ctx->mesh_segs = n_segments;
    CK(cudaFuncSetAttribute(k_raster_droplets, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin));
    return PCR_OK;
}

int pcr_droplet_transforms(pcr_ctx* ctx, const float* d_pcl, int64_t n, int cols, const float* d_rot, float* d_xf, void* stream)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (n < 0 || (cols != 3 && cols != 6)) return fail(ctx, PCR_ERR_INVALID, "n < 0 or cols not 3|6");
    if (n == 0) return PCR_OK;
    if (!d_pcl || !d_xf) return fail(ctx, PCR_ERR_INVALID, "pcr_droplet_transforms: NULL buffer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH(KID_DROP_PREP, s, k_droplet_transforms<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_pcl, n, cols, d_rot, d_xf));
    return PCR_OK;
}

int pcr_history_trails(pcr_ctx* ctx, const float* d_hist, int n_history, const float* d_pos, int64_t n, float* d_ctrl, int32_t* d_count,
                       void* stream)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (n < 0 || n_history < 0) return fail(ctx, PCR_ERR_INVALID, "pcr_history_trails: negative size");
    if (n == 0) return PCR_OK;
    if ((n_history > 0 && !d_hist) || !d_pos || !d_ctrl || !d_count) return fail(ctx, PCR_ERR_INVALID, "pcr_history_trails: NULL buffer");
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_plan(ctx);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH(KID_DROP_PREP, s, k_history_trails<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_hist, n_history, d_pos, n, ctx->plan, d_ctrl, d_count));
    return PCR_OK;
}

int pcr_render_droplet_frames(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols, int n_frames, int n_history,
                              const float* d_rot, const pcr_camera* cams, const pcr_style* style, uint64_t* d_vis, uint8_t* d_rgba,
                              void* stream)
{
    int rc = check_common(ctx, n, cols, style);
    if (rc) return rc;
    if (n_frames < 0 || n_history < 0 || !cams || !d_rgba || !d_in) return fail(ctx, PCR_ERR_INVALID, "pcr_render_droplet_frames: NULL buffer");
    if (n < 1) return fail(ctx, PCR_ERR_INVALID, "pcr_render_droplet_frames: empty frames cannot be standardised");
    if (!ctx->mesh_rings) return fail(ctx, PCR_ERR_INVALID, "pcr_render_droplet_frames: call pcr_set_droplet_mesh first");
    if (style->trails < 0 || style->trails > 2) return fail(ctx, PCR_ERR_INVALID, "trails not in {0,1,2}");
    if (2 * (unsigned long long)n > 0xFFFFFFF0ull) return fail(ctx, PCR_ERR_INVALID, "too many points for trail ids (n + i)");
    if (n_frames == 0) return PCR_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    StyleDev st = to_style_dev(style);
    const bool prestandardised = style->xform == 2;       // frames as process() hands them to generate_xml_content
    if (prestandardised) st.xform = 1;
    const int W = cams[0].width, H = cams[0].height;
    for (int f = 0; f < n_frames; ++f)
        if (cams[f].width != W || cams[f].height != H) return fail(ctx, PCR_ERR_INVALID, "all frames of a call must share W x H");
    const long long px = (long long)W * H;
    const size_t elem = in_is_f64 ? 8 : 4;
    const long long frame_stride = n * cols;
    const int B = ctx->max_batch;
    const int total = n_history + n_frames;
    if (!d_vis) { rc = ensure_vis(ctx); if (rc) return rc; }
    if ((rc = ensure_plan(ctx))) return rc;
    if ((rc = enter(ctx, s))) return rc;
    // scratch that earlier stream-ordered calls may still be reading is only ever grown after a sync
    const size_t need_pts = (size_t)std::min(B, n_frames) * (size_t)n;
    if (ctx->dstats_frames < (size_t)total || ctx->dprep_points < need_pts) {
        CK(cudaDeviceSynchronize());
        if (ctx->dstats_frames < (size_t)total) {
            if (ctx->dstats) CK(cudaFree(ctx->dstats));
            ctx->dstats = nullptr; ctx->dstats_frames = 0;
            CK(cudaMalloc((void**)&ctx->dstats, sizeof(double) * 10 * (size_t)total));
            ctx->dstats_frames = (size_t)total;
        }
        if (ctx->dprep_points < need_pts) {
            for (void* p : {(void*)ctx->dxf, (void*)ctx->dctrl, (void*)ctx->dcount, (void*)ctx->dbin, (void*)ctx->dorder}) if (p) CK(cudaFree(p));
            ctx->dxf = nullptr; ctx->dctrl = nullptr; ctx->dcount = nullptr; ctx->dbin = nullptr; ctx->dorder = nullptr; ctx->dprep_points = 0;
            CK(cudaMalloc((void**)&ctx->dbin, need_pts));
            CK(cudaMalloc((void**)&ctx->dorder, sizeof(int) * need_pts));
            if (!ctx->dhist) {
                CK(cudaMalloc((void**)&ctx->dhist, sizeof(unsigned int) * (size_t)ctx->max_batch * DROP_BINS));
                CK(cudaMemset(ctx->dhist, 0, sizeof(unsigned int) * (size_t)ctx->max_batch * DROP_BINS));
                CK(cudaMalloc((void**)&ctx->dedges, sizeof(int) * (size_t)ctx->max_batch * (DROP_SLABS + 1)));
                CK(cudaMalloc((void**)&ctx->dstarts, sizeof(int) * (size_t)ctx->max_batch * (DROP_SLABS + 1)));
                CK(cudaMalloc((void**)&ctx->dcursor, sizeof(int) * (size_t)ctx->max_batch * DROP_SLABS));
            }
            CK(cudaMalloc((void**)&ctx->dxf, sizeof(float) * 12 * need_pts));
            CK(cudaMalloc((void**)&ctx->dctrl, sizeof(float) * 3 * MAX_CTRL * need_pts));
            CK(cudaMalloc((void**)&ctx->dcount, sizeof(int) * need_pts));
            ctx->dprep_points = need_pts;
        }
    }
    // standardisation statistics of every frame in the buffer (the history frames are standardised one by
    // one, like the reference's all_frame_data, traj_renderer.py:728-741); only the frames actually used
    const int first_used = st.trails == 2 && cols == 6 ? std::max(0, n_history - HISTORY_FRAMES) : n_history;
    if (prestandardised)
        LAUNCH(KID_STATS, s, k_identity_stats<<<(total + 127) / 128, 128, 0, s>>>(ctx->dstats, total));
    for (int g = first_used; g < total && !prestandardised; g += B) {
        const int nb = std::min(B, total - g);
        rc = launch_stats(ctx, (const char*)d_in + (size_t)g * frame_stride * elem, in_is_f64, n, cols, frame_stride, nb, INLINE_REGION,
                          ctx->dstats + (size_t)g * 10, 1, s, style->mean_mode);
        if (rc) return rc;
    }
    const DropletMeshDev mesh = mesh_dev(ctx);
    // one warp per droplet; as many warps per CTA as the camera-space vertices of their instances fit in shared memory
    const int drop_warps = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)ctx->smem_optin / ((size_t)mesh.nv * 12)));
    const size_t smem = (size_t)drop_warps * mesh.nv * 12;
    const RawSrc raw = {d_in, in_is_f64, frame_stride, cols, ctx->dstats, nullptr, nullptr};
    FloorLut lut;
    if ((rc = floor_lut(ctx, st, s, &lut))) return rc;
    for (int f0 = 0; f0 < n_frames; f0 += B) {
        const int nb = std::min(B, n_frames - f0);
        const int g0 = n_history + f0;
        rc = upload_frames(ctx, cams + f0, nb, s);
        if (rc) return rc;
        unsigned long long* vis = (unsigned long long*)(d_vis ? d_vis + (size_t)f0 * px : ctx->vis);
        const long long vis_stride = d_vis ? px : (long long)ctx->max_w * ctx->max_h;
        {
            dim3 grid((unsigned)((n + 127) / 128), nb);
            if (in_is_f64)
                LAUNCH(KID_DROP_PREP, s, k_droplet_prepare<double><<<grid, 128, 0, s>>>(raw_frames<double>(&raw), n, g0, st, ctx->d_frames, d_rot,
                                                                                      ctx->plan, ctx->dxf, ctx->dctrl, ctx->dcount));
            else
                LAUNCH(KID_DROP_PREP, s, k_droplet_prepare<float><<<grid, 128, 0, s>>>(raw_frames<float>(&raw), n, g0, st, ctx->d_frames, d_rot,
                                                                                     ctx->plan, ctx->dxf, ctx->dctrl, ctx->dcount));
        }
        dim3 pgrid((unsigned)((W + 63) / 64), (unsigned)((H + 3) / 4), nb);
        LAUNCH(KID_FILL_FLOOR, s, k_fill_floor<<<pgrid, 256, 0, s>>>(ctx->d_frames, st, vis, vis_stride));
        // dense scenes: depth slabs nearest first, the Hi-Z rebuilt after each (see k_raster_droplets); the trails last,
        // culled by everything in front of them
        const bool cull = ctx->occlusion > 0 || (ctx->occlusion < 0 && n >= 16384);
        const unsigned int* hzp = nullptr;
        // persistent grids: the warps stride over the (slab's) droplet list
        const int drop_blocks_sm = std::max(1, std::min(2048 / (32 * drop_warps), (int)((size_t)ctx->smem_optin / std::max<size_t>(smem, 1))));
        const long long per_pass = cull ? (n + 1) / 2 : n;                                   // the largest slab holds a quarter; leave slack
        const dim3 dgrid((unsigned)std::max<long long>(1, std::min<long long>((per_pass + drop_warps - 1) / drop_warps, (long long)ctx->num_sms * drop_blocks_sm * 4 / nb + 1)), nb);
        if (cull) {
            LAUNCH(KID_DROP_PREP, s, k_droplet_depth_bins<<<dim3((unsigned)((n + 1023) / 1024), nb), 256, 0, s>>>(ctx->d_frames, n, mesh.bound, ctx->dxf, ctx->dbin, ctx->dhist));
            LAUNCH(KID_DROP_PREP, s, k_droplet_slab_edges<<<nb, 32, 0, s>>>(ctx->dhist, n, ctx->dedges, ctx->dstarts, ctx->dcursor));
            LAUNCH(KID_DROP_PREP, s, k_droplet_slab_order<<<dim3((unsigned)((n + 1023) / 1024), nb), 256, 0, s>>>(n, ctx->dbin, ctx->dedges, ctx->dcursor, ctx->dorder));
            const dim3 hgrid((unsigned)((W + 255) / 256), (unsigned)((H + HZ_H - 1) / HZ_H), nb);
            for (int slab = 0; slab < DROP_SLABS; ++slab) {
                LAUNCH(KID_RASTER_DROP, s, k_raster_droplets<<<dgrid, 32 * drop_warps, smem, s>>>(ctx->d_frames, n, mesh, 0u, ctx->dxf, vis, vis_stride, ctx->dorder,
                                                                                                  ctx->dstarts, slab, hzp, ctx->hz_cap, ctx->stat_pairs));
                if (slab + 1 < DROP_SLABS) {
                    LAUNCH(KID_HIZ, s, k_hiz_from_vis<<<hgrid, 256, 0, s>>>(ctx->d_frames, vis, vis_stride, ctx->hz, ctx->hz_cap));
                    hzp = ctx->hz;
                }
            }
        } else {
            LAUNCH(KID_RASTER_DROP, s, k_raster_droplets<<<dgrid, 32 * drop_warps, smem, s>>>(ctx->d_frames, n, mesh, 0u, ctx->dxf, vis, vis_stride, nullptr, nullptr, 0,
                                                                                              nullptr, ctx->hz_cap, ctx->stat_pairs));
        }
        if (st.trails && cols == 6) {
            dim3 grid((unsigned)((n + 7) / 8), nb);
            LAUNCH(KID_RASTER_POLY, s, k_raster_polylines<<<grid, 256, 0, s>>>(ctx->d_frames, n, st.trail_radius, (uint32_t)n, ctx->dctrl,
                                                                              ctx->dcount, vis, vis_stride, hzp, ctx->hz_cap));
        }
        uint32_t* rgba = (uint32_t*)(d_rgba + (size_t)f0 * px * 4);
        if (in_is_f64)
            LAUNCH(KID_SHADE_DROP, s, k_shade_droplets<double><<<pgrid, 256, 0, s>>>(ctx->d_frames, st, lut, (const uint64_t*)vis, vis_stride,
                                                                                   raw_frames<double>(&raw), g0, n, mesh, ctx->dxf, ctx->dctrl,
                                                                                   ctx->dcount, rgba, px));
        else
            LAUNCH(KID_SHADE_DROP, s, k_shade_droplets<float><<<pgrid, 256, 0, s>>>(ctx->d_frames, st, lut, (const uint64_t*)vis, vis_stride,
                                                                                  raw_frames<float>(&raw), g0, n, mesh, ctx->dxf, ctx->dctrl,
                                                                                  ctx->dcount, rgba, px));
    }
    return leave(ctx, s);
}

// ------------------------------------------------------------------------------------------
// Fused z-merge over peer memory (point-sharded clouds)
// ------------------------------------------------------------------------------------------
int pcr_peer_alloc(pcr_ctx* ctx, int width, int height, void** d_merged, void** d_image)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (width < 1 || height < 1 || width > ctx->max_w || height > ctx->max_h) return fail(ctx, PCR_ERR_CAPACITY, "pcr_peer_alloc: frame larger than the context");
    CK(cudaSetDevice(ctx->device));
    if (ctx->peer_w != width || ctx->peer_h != height) {
        CK(cudaDeviceSynchronize());
        if (ctx->peer_merged) CK(cudaFree(ctx->peer_merged));
        if (ctx->peer_image) CK(cudaFree(ctx->peer_image));
        ctx->peer_merged = ctx->peer_image = nullptr; ctx->peer_w = ctx->peer_h = 0; ctx->peer.world = 0;
        CK(cudaMalloc(&ctx->peer_merged, sizeof(uint64_t) * (size_t)width * height));
        CK(cudaMalloc(&ctx->peer_image, sizeof(uint32_t) * (size_t)width * height));
        ctx->peer_w = width; ctx->peer_h = height;
    }
    if (d_merged) *d_merged = ctx->peer_merged;
    if (d_image) *d_image = ctx->peer_image;
    return PCR_OK;
}

int pcr_ipc_export(pcr_ctx* ctx, const void* d_ptr, uint8_t handle[64])
{
    if (!ctx) return PCR_ERR_INVALID;
    if (!d_ptr || !handle) return fail(ctx, PCR_ERR_INVALID, "pcr_ipc_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    memcpy(handle, &h, 64);
    return PCR_OK;
}

int pcr_ipc_open(pcr_ctx* ctx, const uint8_t handle[64], void** d_ptr)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (!d_ptr || !handle) return fail(ctx, PCR_ERR_INVALID, "pcr_ipc_open: NULL argument");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PCR_OK;
}

int pcr_ipc_close(pcr_ctx* ctx, void* d_ptr)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (!d_ptr) return PCR_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaIpcCloseMemHandle(d_ptr));
    return PCR_OK;
}

int pcr_peer_set(pcr_ctx* ctx, int rank, int world, int dst_rank, void* const* merged_ptrs, void* const* image_ptrs)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (world == 0) { ctx->peer.world = 0; return PCR_OK; }
    if (world < 1 || world > MAX_PEERS || rank < 0 || rank >= world || dst_rank < 0 || dst_rank >= world || !merged_ptrs || !image_ptrs)
        return fail(ctx, PCR_ERR_INVALID, "pcr_peer_set: bad rank / world (at most 8 ranks)");
    if (!ctx->peer_merged) return fail(ctx, PCR_ERR_INVALID, "pcr_peer_set: call pcr_peer_alloc first");
    PeerDev p = {};
    for (int r = 0; r < world; ++r) {
        if (!merged_ptrs[r] || !image_ptrs[r]) return fail(ctx, PCR_ERR_INVALID, "pcr_peer_set: NULL peer pointer");
        p.merged[r] = (unsigned long long*)merged_ptrs[r];
        p.image[r] = (uint32_t*)image_ptrs[r];
    }
    if (p.merged[rank] != ctx->peer_merged || p.image[rank] != ctx->peer_image)
        return fail(ctx, PCR_ERR_INVALID, "pcr_peer_set: entry [rank] must be this context's own buffers");
    p.world = world; p.rank = rank; p.dst = dst_rank;
    p.base = ctx->peer_h / world; p.rem = ctx->peer_h % world;
    ctx->peer = p;
    return PCR_OK;
}

int pcr_peer_begin_frame(pcr_ctx* ctx, const pcr_camera* cam, const pcr_style* style, void* stream)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (!cam || !style) return fail(ctx, PCR_ERR_INVALID, "pcr_peer_begin_frame: NULL argument");
    if (ctx->peer.world < 1) return fail(ctx, PCR_ERR_INVALID, "pcr_peer_begin_frame: call pcr_peer_set first");
    if (cam->width != ctx->peer_w || cam->height != ctx->peer_h) return fail(ctx, PCR_ERR_INVALID, "camera does not match pcr_peer_alloc");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc = enter(ctx, s);
    if (rc) return rc;
    if ((rc = upload_frames(ctx, cam, 1, s))) return rc;
    const PeerDev& p = ctx->peer;
    const int y0 = p.rank * p.base + std::min(p.rank, p.rem), y1 = y0 + p.base + (p.rank < p.rem ? 1 : 0);
    if (y1 > y0) {
        dim3 grid((unsigned)((cam->width + 63) / 64), (unsigned)((y1 - y0 + 3) / 4));
        LAUNCH(KID_PEER_INIT, s, k_peer_init_rows<<<grid, 256, 0, s>>>(ctx->d_frames, to_style_dev(style), p, y0, y1));
    }
    return leave(ctx, s);
}

int pcr_render_shard_peer(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols, const float* d_radius, const float* d_rgb,
                          const double* d_stats10, uint32_t id_base, const pcr_camera* cam, const pcr_style* style, uint64_t* d_vis,
                          void* stream)
{
    int rc = check_common(ctx, n, cols, style);
    if (rc) return rc;
    if (!cam || !d_vis || !d_stats10 || (n > 0 && !d_in)) return fail(ctx, PCR_ERR_INVALID, "pcr_render_shard_peer: NULL buffer");
    if (ctx->peer.world < 1) return fail(ctx, PCR_ERR_INVALID, "pcr_render_shard_peer: call pcr_peer_set first");
    if (cam->width != ctx->peer_w || cam->height != ctx->peer_h) return fail(ctx, PCR_ERR_INVALID, "camera does not match pcr_peer_alloc");
    if ((unsigned long long)id_base + (unsigned long long)n > 0xFFFFFFFEull) return fail(ctx, PCR_ERR_INVALID, "point id overflow");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = enter(ctx, s))) return rc;
    if ((rc = upload_frames(ctx, cam, 1, s))) return rc;
    StyleDev st = to_style_dev(style);
    st.trails = 0;
    const RawSrc raw = {d_in, in_is_f64, n * cols, cols, d_stats10, d_radius, d_rgb};
    const long long px = (long long)cam->width * cam->height;
    rc = launch_render(ctx, nullptr, nullptr, 0, &raw, n, 1, id_base, st, cam->width, cam->height, d_vis, px, nullptr, px, 0, s, &ctx->peer);
    if (rc) return rc;
    return leave(ctx, s);
}

int pcr_shade_shard_peer(pcr_ctx* ctx, const uint64_t* d_vis, const void* d_in, int in_is_f64, int64_t n, int cols, const float* d_radius,
                         const float* d_rgb, const double* d_stats10, uint32_t id_base, const pcr_camera* cam, const pcr_style* style,
                         void* stream)
{
    int rc = check_common(ctx, n, cols, style);
    if (rc) return rc;
    if (!cam || !d_vis || !d_stats10 || (n > 0 && !d_in)) return fail(ctx, PCR_ERR_INVALID, "pcr_shade_shard_peer: NULL buffer");
    if (ctx->peer.world < 1) return fail(ctx, PCR_ERR_INVALID, "pcr_shade_shard_peer: call pcr_peer_set first");
    if (cam->width != ctx->peer_w || cam->height != ctx->peer_h) return fail(ctx, PCR_ERR_INVALID, "camera does not match pcr_peer_alloc");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = enter(ctx, s))) return rc;
    if ((rc = upload_frames(ctx, cam, 1, s))) return rc;
    StyleDev st = to_style_dev(style);
    st.trails = 0;
    FloorLut lut;
    if ((rc = floor_lut(ctx, st, s, &lut))) return rc;
    const RawSrc raw = {d_in, in_is_f64, n * cols, cols, d_stats10, d_radius, d_rgb};
    // one block per tile this rank has something to do in (drawn in, or inside its row band); the rest are skipped
    const int tiles = ((cam->width + TILE - 1) / TILE) * ((cam->height + TILE - 1) / TILE);
    const unsigned grid = (unsigned)std::max(1, std::min(tiles, ctx->num_sms * 16));
    // (with the lazy floor fill the local keys only exist in the tiles the render call drew in: its tile states must still be there)
    if (ctx->lazy_fill && !ctx->peer_tiles_valid)
        return fail(ctx, PCR_ERR_INVALID, "pcr_shade_shard_peer must follow the pcr_render_shard_peer of the same frame on this context (no other render in between)");
    const unsigned int* tstate = ctx->peer_tiles_valid ? ctx->tile_state : nullptr;
    if (in_is_f64)
        LAUNCH(KID_SHADE_PEER, s, k_shade_peer<double><<<grid, 256, 0, s>>>(ctx->d_frames, st, lut, d_vis, raw_frames<double>(&raw), n, id_base, ctx->peer, tstate));
    else
        LAUNCH(KID_SHADE_PEER, s, k_shade_peer<float><<<grid, 256, 0, s>>>(ctx->d_frames, st, lut, d_vis, raw_frames<float>(&raw), n, id_base, ctx->peer, tstate));
    return leave(ctx, s);
}

int pcr_selftest_scale_div(pcr_ctx* ctx, const float* h_divisors, int n, uint64_t* mismatches)
{
    if (!ctx) return PCR_ERR_INVALID;
    if (!h_divisors || !mismatches || n < 1 || n > 65535) return fail(ctx, PCR_ERR_INVALID, "divisors / mismatches NULL or n out of range");
    CK(cudaSetDevice(ctx->device));
    float* d_div = nullptr;
    unsigned long long* d_bad = nullptr;
    CK(cudaMalloc((void**)&d_div, sizeof(float) * n));
    cudaError_t e = cudaMalloc((void**)&d_bad, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemcpy(d_div, h_divisors, sizeof(float) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d_bad, 0, sizeof(unsigned long long));
    if (e == cudaSuccess) {
        k_selftest_scale_div<<<dim3((unsigned)(ctx->num_sms * 8), (unsigned)n), 256>>>(d_div, n, d_bad);
        e = cudaGetLastError();
    }
    unsigned long long bad = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
    cudaFree(d_div);
    if (d_bad) cudaFree(d_bad);
    if (e != cudaSuccess) return fail(ctx, PCR_ERR_CUDA, "pcr_selftest_scale_div", e);
    *mismatches = bad;
    return PCR_OK;
}

const char* pcr_kernel_name(int kernel_id) { return kernel_id >= 0 && kernel_id < KID_COUNT ? kKernelNames[kernel_id] : ""; }

}  // extern "C"
