/*
 * pcr.h — C ABI of the B200-native point-cloud sphere renderer ("pcr").
 *
 * This is the drop-in boundary for the one hot path of EvaShenLu/PointCloud_Render:
 * everything the reference does between "a frame's (N,3|6) point array" and
 * "an image", i.e. its L2 scene emission + L1 Mitsuba render:
 *
 *   standardize_point_cloud     example_renderer.py:94-98, traj_ball_renderer.py:190-202
 *   axis transform              example_renderer.py:171-173, traj_ball_renderer.py:204-221,
 *                               traj_b0.py:62-82 (no x flip)
 *   compute_color hook          example_renderer.py:89-92,115-124
 *   generate_xml_content        example_renderer.py:113-128   (eliminated: data stays in HBM)
 *   render_scene                example_renderer.py:153-157   (mi.load_file + mi.render)
 *   save_scene (sRGB8 part)     example_renderer.py:159-161   (linear -> sRGB -> u8; PNG stays on host)
 *
 * The reference has no FFI of its own; its seam is the render_scene/save_scene
 * pair.  INTEGRATION.md shows the ctypes stub a maintainer adds.
 *
 * Conventions
 *   - plain C types only; every d_* pointer is CALLER-OWNED DEVICE memory, every
 *     h_* pointer is host memory; the context owns only scratch.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *     All work is stream-ordered and asynchronous; no entry point synchronises
 *     the device unless its comment says so.
 *   - return 0 on success, negative pcr_status otherwise; the message is kept
 *     per context (pcr_last_error). No C++ exception crosses this boundary.
 *   - one context per (GPU, host thread); a context is not thread-safe.
 *   - there is NO CPU fallback: without a CUDA device pcr_create fails.
 */
#ifndef PCR_H_
#define PCR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCR_ABI_VERSION 4

/* Point ids stored in the low 32 bits of a visibility key. */
#define PCR_ID_FLOOR 0xFFFFFFFEu
#define PCR_ID_MISS  0xFFFFFFFFu
/* key = (float_as_uint(depth) << 32) | id ; depth = camera-space z of the hit (> 0).
 * A pixel that nothing covers holds depth = +inf, id = PCR_ID_MISS.              */
#define PCR_KEY_MISS 0x7F800000FFFFFFFFull

typedef enum pcr_status {
    PCR_OK = 0,
    PCR_ERR_INVALID = -1,   /* bad argument                                   */
    PCR_ERR_CUDA = -2,      /* a CUDA runtime call failed                     */
    PCR_ERR_CAPACITY = -3,  /* n / W / H / batch larger than the context      */
    PCR_ERR_NCCL = -4,      /* NCCL missing or a collective failed            */
    PCR_ERR_NOMEM = -5
} pcr_status;

/* Colour hook modes (compute_color, example_renderer.py:89-92).  Mode 0 is the
 * reference's behaviour; 1-3 are the extensions BASELINE.json's north_star names. */
enum {
    PCR_COLOR_CONST = 0,     /* constant const_rgb (reference: 0.3 grey)                  */
    PCR_COLOR_POSITION = 1,  /* colormap of (p-min)/(range+1e-8), example_renderer.py:121 */
    PCR_COLOR_VELOCITY = 2,  /* ramp on min(|v|/vel_norm,1), traj_ball_renderer.py:134    */
    PCR_COLOR_USER = 3       /* caller-supplied per-point RGB                             */
};

/* <sensor type="perspective"> block of XMLTemplates.HEAD (example_renderer.py:16-31). */
typedef struct pcr_camera {
    float origin[3];
    float target[3];
    float up[3];
    float fov_x_deg;   /* horizontal field of view, degrees */
    float near_clip;
    float far_clip;
    int32_t width;
    int32_t height;
    double trail_scale; /* length_scale of _add_velocity_trail for THIS frame (traj_ball_renderer.py:119-124,
                           traj_vel_renderer.py:215-224; 1.0 in traj_original/b0/b1); only read when
                           pcr_style.trails is set; <= 0 draws no trails */
} pcr_camera;

/* Scene constants of XMLTemplates.BALL_SEGMENT / TAIL (example_renderer.py:41-72)
 * plus the axis-transform flavour (traj_ball_renderer.py:204-221 vs traj_b0.py:62-82). */
typedef struct pcr_style {
    int32_t color_mode;
    float const_rgb[3];
    float radius;        /* sphere radius when no per-point radius is given (0.01) */
    int32_t flip_x;      /* 1: pos' = (-z, x, y+z_lift); 0: (z, x, y+z_lift)       */
    float z_lift;        /* 0.0125                                                */
    float vel_norm;      /* 10.0                                                  */
    int32_t has_floor;
    float floor_z;
    float floor_min[2];  /* world x,y extent of the ground rectangle              */
    float floor_max[2];
    float floor_albedo;  /* diffuse reflectance of the ground (1.0)               */
    float light_z;       /* area emitter: square |x|,|y| <= light_half at z       */
    float light_half;
    float radiance;
    float bounce;        /* weight of the ground-bounce term on spheres (1.0)     */
    int32_t xform;       /* 0: the reference's axis transform (permute, flip, lift);
                            1: none (standardise only) — lets the facade expose
                            standardize_point_cloud / transform_coordinates separately;
                            2 (pcr_render_droplet_frames only): the frames are ALREADY standardised
                            and transformed — used as they are, no statistics taken (the position
                            colormap has no range then) */
    int32_t mean_mode;   /* how the centre of standardize_point_cloud is summed (PCR_MEAN_*)      */
    /* velocity trails (_add_velocity_trail, traj_ball_renderer.py:98-188): pcr_render_frames[_host] with
     * 6-column frames draws, per point, the straight trail  position - v_hat*L -> position,
     * L = (trail_len_min + (trail_len_max - trail_len_min) * min(|v|/vel_norm, 1)) * camera.trail_scale,
     * as a capsule of radius trail_radius; its id in the visibility buffer is n + point index.   */
    int32_t trails;          /* 0: none (default), 1: the straight velocity trail, 2 (pcr_render_droplet_frames
                                only): the Catmull-Rom history trail of traj_renderer.py:204-396           */
    float trail_radius;      /* 0.0007  (traj_ball_renderer.py:160)                               */
    float trail_rgb[3];      /* 0.2, 1.0, 0.4  (:179)                                             */
    double trail_len_min;    /* 0.07 (:132) — doubles: the reference computes the length in f64   */
    double trail_len_max;    /* 0.3  (:133)                                                       */
} pcr_style;

/* The reference's np.mean(axis=0) is a SEQUENTIAL sum in the input dtype (example_renderer.py:96),
 * whose float32 rounding error grows with N.  PCR_MEAN_SEQUENTIAL reproduces it bit for bit (one
 * thread per axis: ~2 ms of latency per million points, which the whole-path entries hide by computing it for
 * later batches on side streams while earlier ones render — see pcr_prefetch_frames); PCR_MEAN_F64 sums in float64
 * in parallel and rounds once (order independent, more accurate, not bit-identical to numpy for float32 input).
 * PCR_MEAN_AUTO (the default) = the reference's arithmetic wherever one GPU sees the whole frame, i.e. SEQUENTIAL
 * in every entry that takes whole frames, at any size; only the point-sharded entries (pcr_stats_partial ->
 * pcr_finalize_stats), where a float fold cannot be split across shards, use the float64 mean. */
enum { PCR_MEAN_AUTO = 0, PCR_MEAN_SEQUENTIAL = 1, PCR_MEAN_F64 = 2 };

/* f32 camera frame derived on the HOST in double precision from a pcr_camera.
 * Kernels and the parity oracle consume exactly these numbers. */
typedef struct pcr_frame {
    float L[3];          /* camera x axis ("left")  = normalize(up x dir)   */
    float U[3];          /* camera y axis ("newup") = dir x left            */
    float D[3];          /* viewing direction       = normalize(target-origin) */
    float O[3];          /* eye                                              */
    float T;             /* tan(fov_x/2)                                     */
    float Th;            /* T * H / W                                        */
    float TW;            /* T / W                                            */
    float near_clip;
    float far_clip;
    int32_t W;
    int32_t H;
} pcr_frame;

typedef struct pcr_ctx pcr_ctx;

int pcr_abi_version(void);

/* Allocates scratch for up to max_points points per frame, max_w x max_h pixels and
 * max_batch frames in flight per launch.  pair_capacity = max (tile, sphere) pairs per
 * frame (0 = default 24*max_points + 65536 up to 262144 points, 12*max_points + 65536 above;
 * 24 bytes of scratch each); frames that exceed it take the slower un-binned raster, results are
 * identical.  Synchronous. */
int pcr_create(pcr_ctx** out, int device, int64_t max_points, int max_w, int max_h,
               int max_batch, int64_t pair_capacity);
void pcr_destroy(pcr_ctx* ctx);
const char* pcr_last_error(const pcr_ctx* ctx);

/* Host helper: the derived camera frame (no GPU work). */
int pcr_camera_frame(const pcr_camera* cam, pcr_frame* out);

/* K0+K1 — standardize_point_cloud + axis transform + colour hook for ONE frame.
 *   d_in       (n, cols) row-major, float32 or float64 (in_is_f64), cols = 3 or 6
 *   d_radius   optional per-point radius [n] (NULL -> style->radius)
 *   d_rgb      optional per-point RGB [n][3] for PCR_COLOR_USER
 *   d_pos_out  [n] float4 = (x', y', z', radius)      transformed positions
 *   d_attr_out [n] float4 = (r, g, b, |v|)            colour hook output
 *   d_vel_out  optional [n] float4 = (vx', vy', vz', 0)  transformed velocities (cols==6)
 *   d_stats    optional device double[10]: mean xyz, min xyz, max xyz, scale (input axes)
 */
int pcr_standardize(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols,
                    const float* d_radius, const float* d_rgb, const pcr_style* style,
                    float* d_pos_out, float* d_attr_out, float* d_vel_out, double* d_stats,
                    void* stream);

/* transform_coordinates alone (traj_ball_renderer.py:204-221; no-flip: traj_b0.py:62-82,
 * traj_original.py:40-60) on an already standardised (n, cols) float32 array:
 * pos' = (-+z, x, y + z_lift), vel' = (-+vz, vx, vy).  d_out must not alias d_in. */
int pcr_transform_coordinates(pcr_ctx* ctx, const float* d_in, int64_t n, int cols, int flip_x,
                              float z_lift, float* d_out, void* stream);

/* _add_velocity_trail's geometry alone (traj_ball_renderer.py:98-176): for an already transformed
 * (n, 6) float32 array, the first and last control point of the curve file the reference writes
 * for every point — tail = position - v_hat*L and head = position, both after the file's 6-decimal
 * text round trip, as float32 — and whether the reference draws a trail at all (|v| >= 1e-6 and
 * trail_scale > 0).  d_tail, d_head: [n][3] float32; d_valid: [n] uint8. */
int pcr_velocity_trails(pcr_ctx* ctx, const float* d_pcl6, int64_t n, const pcr_style* style,
                        double trail_scale, float* d_tail, float* d_head, uint8_t* d_valid, void* stream);

/* K2+K3(+K4) — render already-transformed spheres (what generate_xml_content would emit).
 *   d_pos    [n] float4 (x,y,z,r) world space ; d_attr [n] float4 (r,g,b,_)
 *   id_base  added to the local index to form the stored point id (point sharding)
 *   d_vis    [H][W] uint64 keys, written in full (required)
 *   d_rgba   [H][W][4] uint8 sRGB image, or NULL to skip shading
 */
int pcr_render(pcr_ctx* ctx, const float* d_pos, const float* d_attr, int64_t n,
               uint32_t id_base, const pcr_camera* cam, const pcr_style* style,
               uint64_t* d_vis, uint8_t* d_rgba, void* stream);

/* render_scene for a frame that is ALREADY standardised and transformed — the (n, 3|6) float32 array process() hands to
 * generate_xml_content (traj_ball_renderer.py:309-333, example_renderer.py:113-128): one sphere per row and, for 6
 * columns with style->trails == 1, the velocity trail _add_velocity_trail draws for every point (id n + index,
 * length scale cam->trail_scale).  Positions / velocities are used bit for bit as given (style->xform and mean_mode are
 * ignored).  d_vis may be NULL.  This is what the facade's process() calls, so a reference-named entry point draws what
 * the reference draws. */
int pcr_render_transformed(pcr_ctx* ctx, const float* d_pcl, int64_t n, int cols, const float* d_radius, const float* d_rgb,
                           const pcr_camera* cam, const pcr_style* style, uint64_t* d_vis, uint8_t* d_rgba, void* stream);

/* K4 alone — shade a (possibly merged) visibility buffer.  Points whose id is outside
 * [id_base, id_base+n) are shaded only when owner_only == 0 is impossible for them, so:
 *   owner_only = 0: every sphere id must be local (single GPU);
 *   owner_only = 1: pixels won by a non-local sphere are written as 0,0,0,0 and floor /
 *                   miss pixels are written only when id_base == 0 (rank 0), so that a
 *                   byte-wise MAX all-reduce of the images assembles the frame. */
int pcr_shade(pcr_ctx* ctx, const uint64_t* d_vis, const float* d_pos, const float* d_attr,
              int64_t n, uint32_t id_base, int owner_only, const pcr_camera* cam,
              const pcr_style* style, uint8_t* d_rgba, void* stream);

/* Whole hot path for a trajectory: n_frames frames of (n, cols) points, resident on the
 * device, -> visibility keys and sRGB8 images.  Frames are batched max_batch per launch.
 *   d_in   [n_frames][n][cols] ; cams: HOST array of n_frames cameras
 *   d_vis  [n_frames][H][W] or NULL ; d_rgba [n_frames][H][W][4] (required)
 */
int pcr_render_frames(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols,
                      int n_frames, const float* d_radius, const float* d_rgb,
                      const pcr_camera* cams, const pcr_style* style,
                      uint64_t* d_vis, uint8_t* d_rgba, void* stream);

/* Hint: the n_frames frames at d_in (as they are once everything already submitted to `stream` has run) will be handed
 * to pcr_render_frames later, unchanged, with the same n / cols / dtype / style->mean_mode and at the same addresses.
 * Their K0 products — the standardisation statistics, incl. the serial reference-exact mean (~2 ms of latency per
 * million points), and the occluder pre-pass sample — are computed NOW on side streams, one per batch, so that they
 * overlap whatever renders in the meantime (the reference does the same on the host: traj_renderer.py:718-743
 * standardises every frame before it renders the first).  pcr_render_frames looks its batches up (by address and
 * shape, max_batch frames at a time) and computes only what nobody prepared; results never depend on hints.  At
 * most 6 batches are kept (further frames are simply not prefetched; an unconsumed hint is dropped when its slot is
 * needed).  n_frames == 0 drops every hint.  The caller must not modify hinted frames before rendering them. */
int pcr_prefetch_frames(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols, int n_frames,
                        const pcr_style* style, void* stream);

/* Same with HOST buffers: h_in is copied host->device and h_rgba (and h_vis if not NULL)
 * device->host in chunks on two internal streams so copies overlap the kernels.
 * Pinned host memory makes the copies asynchronous.  Synchronous: returns when the
 * images are in h_rgba. */
int pcr_render_frames_host(pcr_ctx* ctx, const void* h_in, int in_is_f64, int64_t n, int cols,
                           int n_frames, const float* h_radius, const float* h_rgb,
                           const pcr_camera* cams, const pcr_style* style,
                           uint64_t* h_vis, uint8_t* h_rgba);

/* The same, asynchronously: submit enqueues the call's copies and kernels and returns a ticket; pcr_host_wait blocks
 * until that call's images (and keys) are in the caller's buffers (ticket < 0: every call submitted so far).  Two calls
 * in flight keep the H2D stream busy across calls: the kernels, the serial mean and the D2H copy of the end of call k
 * overlap the input copy of call k+1 (Mitsuba's write_bitmap is asynchronous in the same way, example_renderer.py:161).
 * h_in / h_rgba / h_vis / cams' frames must stay valid and untouched until the ticket has been waited for. */
int pcr_render_frames_host_submit(pcr_ctx* ctx, const void* h_in, int in_is_f64, int64_t n, int cols,
                                  int n_frames, const float* h_radius, const float* h_rgb,
                                  const pcr_camera* cams, const pcr_style* style,
                                  uint64_t* h_vis, uint8_t* h_rgba, int64_t* ticket);
int pcr_host_wait(pcr_ctx* ctx, int64_t ticket);

/* d_dst[i] = min(d_dst[i], d_src[i]) on uint64 keys — the local half of a z-buffer merge. */
int pcr_zmin(pcr_ctx* ctx, uint64_t* d_dst, const uint64_t* d_src, int64_t n_px, void* stream);

/* C1 — in-place ncclAllReduce(d_vis, n_px, ncclUint64, ncclMin) over NVLink.
 * `comm` is an ncclComm_t; NCCL is resolved from the running process (dlsym), so the
 * library has no link-time NCCL dependency. */
int pcr_zmerge_nccl(pcr_ctx* ctx, uint64_t* d_vis, int64_t n_px, void* comm, void* stream);

/* Point-sharded standardisation, step 1 and 2 (C0 lives between them):
 *   pcr_stats_partial  writes double[9] = sum xyz, min xyz, max xyz of the local shard;
 *   the caller all-reduces (sum / min / max) and passes the global double[10]
 *   (mean xyz, min xyz, max xyz, scale) to pcr_standardize_with_stats. */
int pcr_stats_partial(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols,
                      double* d_partial9, void* stream);
/* Device half of C0: d_partials = [n_shards][9] doubles (every rank's pcr_stats_partial output,
 * all-gathered in rank order), n_total = points of the whole cloud -> d_stats10, with the same
 * roundings as a single-GPU frame.  No host synchronisation. */
int pcr_finalize_stats(pcr_ctx* ctx, const double* d_partials, int n_shards, int64_t n_total,
                       int in_is_f64, double* d_stats10, void* stream);
int pcr_standardize_with_stats(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols,
                               const float* d_radius, const float* d_rgb, const pcr_style* style,
                               const double* d_stats10, float* d_pos_out, float* d_attr_out,
                               float* d_vel_out, void* stream);

/* Point-sharded whole path without materialising the transformed arrays (K1 fused into K2a / K4,
 * like pcr_render_frames) for ONE shard of a cloud whose global stats are already known:
 *   pcr_stats_partial -> all-gather -> pcr_finalize_stats -> pcr_render_shard -> z-merge ->
 *   pcr_shade_shard(owner_only = 1) -> byte-MAX assembly.
 * d_in: this shard's raw (n, cols) points; ids stored = id_base + local index.  No trails. */
int pcr_render_shard(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols,
                     const float* d_radius, const float* d_rgb, const double* d_stats10, uint32_t id_base,
                     const pcr_camera* cam, const pcr_style* style, uint64_t* d_vis, void* stream);
int pcr_shade_shard(pcr_ctx* ctx, const uint64_t* d_vis, const void* d_in, int in_is_f64, int64_t n, int cols,
                    const float* d_radius, const float* d_rgb, const double* d_stats10, uint32_t id_base,
                    int owner_only, const pcr_camera* cam, const pcr_style* style, uint8_t* d_rgba, void* stream);

/* ---- droplet scene: traj_renderer.py / traj_vel_renderer.py --------------------------
 * Those two scripts draw every point as an instance of a droplet mesh (_create_droplet_mesh,
 * traj_renderer.py:102-153: a surface of revolution of (n_rings+1) x n_segments vertices written to an OBJ
 * file) placed by the 4x4 matrix of generate_rotation_matrix_from_velocity (:159-202), and one `linearcurve`
 * polyline per point: the Catmull-Rom history trail of _add_trail_lines (:204-396) or the straight velocity
 * trail of traj_vel_renderer.py:194-288.  Ids in the visibility buffer: droplet of point i = i, its trail = n + i. */
#define PCR_HISTORY_FRAMES 20   /* trail_length_frames, traj_renderer.py:218 */
#define PCR_MAX_CTRL 21         /* 20 spline samples (:272) + the current position (:331) */

/* The mesh: h_verts = HOST [(n_rings+1) * n_segments][3] float32, ring-major, exactly as a loader reads the
 * OBJ text (6 decimals); faces are the reference's (v0,v2,v1),(v1,v2,v3) per quad.  Ring z must decrease
 * strictly from ring 0 to ring n_rings.  Synchronous; must be called before pcr_render_droplet_frames. */
int pcr_set_droplet_mesh(pcr_ctx* ctx, const float* h_verts, int n_rings, int n_segments);

/* generate_rotation_matrix_from_velocity alone: d_pcl = transformed (n, cols) float32; cols == 6: rotation from
 * the velocity columns; cols == 3: d_rot[n][9] (generate_random_rotation_matrix, :398-418 — the numpy legacy
 * generator stays on the host) or identity when NULL.  d_xf[n][12] = rows of [R | position], float32 — what a
 * loader reads from the matrix DROPLET_SEGMENT prints. */
int pcr_droplet_transforms(pcr_ctx* ctx, const float* d_pcl, int64_t n, int cols, const float* d_rot, float* d_xf, void* stream);

/* _add_trail_lines alone: d_hist = [n_history][n][3] transformed float32 positions of the previous frames
 * (oldest first; the last 20 are used), d_pos = [n][3] current positions.  d_ctrl[n][PCR_MAX_CTRL][3] = control
 * points of the curve file the reference writes (after its 6-decimal text round trip), d_count[n] = how many
 * are valid (0 = the reference draws no trail for that point). */
int pcr_history_trails(pcr_ctx* ctx, const float* d_hist, int n_history, const float* d_pos, int64_t n, float* d_ctrl,
                       int32_t* d_count, void* stream);

/* Whole path.  d_in = [n_history + n_frames][n][cols] raw frames on the device: the n_frames frames to render,
 * preceded by the n_history frames before them (a frame-sharded caller passes a 20-frame halo).  Every frame
 * is standardised on its own, like the reference's all_frame_data (traj_renderer.py:728-741).  Frame j uses
 * cams[j]; with style->trails == 2 its history is the up to 20 buffer frames before it; == 1 draws the
 * velocity trail (cams[j].trail_scale); cols == 3 draws no trails and takes d_rot (or identity). */
int pcr_render_droplet_frames(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols, int n_frames, int n_history,
                              const float* d_rot, const pcr_camera* cams, const pcr_style* style, uint64_t* d_vis,
                              uint8_t* d_rgba, void* stream);

/* ---- fused z-merge over peer memory (point-sharded clouds; no reference counterpart) ----------
 * Alternative to pcr_render_shard -> ncclAllReduce(min) -> pcr_shade_shard -> byte-MAX all-reduce: every rank owns a
 * band of image rows of the merged z-buffer; the raster pushes its winners straight into the owner's rows with
 * 64-bit atomicMin over NVLink WHILE it rasterises the remaining tiles, and the shade kernel reads the merged key
 * back from the owner and stores each pixel it won directly into rank dst_rank's image.  The only collectives
 * left are three tiny ones used as barriers (the 72-byte stats all-gather and two 1-element all-reduces).
 *
 * Per frame, on every rank, stream-ordered:
 *   pcr_peer_begin_frame                      floor keys into the rows this rank owns
 *   pcr_stats_partial -> all-gather -> pcr_finalize_stats      (the all-gather also orders every rank's
 *                                                               begin_frame before anybody's pushes)
 *   pcr_render_shard_peer                     K2/K3 into the local d_vis + pushes into the owners' rows
 *   barrier  (any collective)                 all pushes have landed
 *   pcr_shade_shard_peer                      K4: pixels this rank won / owns -> image of dst_rank (must follow this frame's
 *                                             pcr_render_shard_peer on the same context with no other render in between: d_vis
 *                                             holds valid keys only in the tiles that call drew in, and their list lives in the context)
 *   barrier                                   the image on dst_rank is complete; merged rows may be reused
 * Set-up, once: pcr_peer_alloc on every rank, exchange the buffers' IPC handles (pcr_ipc_export / pcr_ipc_open;
 * ranks inside ONE process pass raw pointers), pcr_peer_set.  At most 8 ranks. */
int pcr_peer_alloc(pcr_ctx* ctx, int width, int height, void** d_merged, void** d_image);
int pcr_ipc_export(pcr_ctx* ctx, const void* d_ptr, uint8_t handle[64]);      /* cudaIpcGetMemHandle */
int pcr_ipc_open(pcr_ctx* ctx, const uint8_t handle[64], void** d_ptr);       /* cudaIpcOpenMemHandle */
int pcr_ipc_close(pcr_ctx* ctx, void* d_ptr);
/* merged_ptrs / image_ptrs: HOST arrays of `world` device pointers in rank order; entry [rank] must be this
 * context's own buffers.  world == 0 detaches. */
int pcr_peer_set(pcr_ctx* ctx, int rank, int world, int dst_rank, void* const* merged_ptrs, void* const* image_ptrs);
int pcr_peer_begin_frame(pcr_ctx* ctx, const pcr_camera* cam, const pcr_style* style, void* stream);
int pcr_render_shard_peer(pcr_ctx* ctx, const void* d_in, int in_is_f64, int64_t n, int cols, const float* d_radius,
                          const float* d_rgb, const double* d_stats10, uint32_t id_base, const pcr_camera* cam,
                          const pcr_style* style, uint64_t* d_vis, void* stream);
int pcr_shade_shard_peer(pcr_ctx* ctx, const uint64_t* d_vis, const void* d_in, int in_is_f64, int64_t n, int cols,
                         const float* d_radius, const float* d_rgb, const double* d_stats10, uint32_t id_base,
                         const pcr_camera* cam, const pcr_style* style, void* stream);

/* Counters of the last pcr_render / pcr_render_frames call (synchronises `stream`):
 * out[0] = kernels launched, out[1] = (tile,sphere) pairs of the last frame,
 * out[2] = frames that overflowed pair_capacity, out[3] = spheres culled (last frame). */
int pcr_counters(pcr_ctx* ctx, int64_t out[4], void* stream);

/* Occlusion pre-pass of pcr_render / pcr_render_frames.  Dense clouds bury most spheres (depth
 * complexity in the hundreds at 1 M points): the pre-pass rasterises every step-th point (of the
 * points behind the cloud's centre plane only every 8th of those: they lose to nearer ones almost
 * everywhere), builds a per-8x4-pixel farthest-depth map, and the main pass drops every sphere that
 * lies entirely behind it before it is binned.  Purely a work-skipping device: the keys are identical.
 *   mode -1: automatic (on when n >= min_points; default min_points 131072), 0: off, 1: always
 *   step  0: keep ; -1: automatic (the default: every 8th point up to 1.5 M points per frame, growing to every 64th
 *         from 8 M — the pre-pass needs a number of near spheres, not a share of the cloud) ; min_points 0: keep */
int pcr_set_occlusion(pcr_ctx* ctx, int mode, int step, int64_t min_points);

/* Per-kernel timing.  While enabled, every kernel launch is bracketed by two CUDA events on the
 * launching stream.  pcr_profile_read waits for the recorded events, writes the summed
 * milliseconds and launch counts per kernel id (index k <-> pcr_kernel_name(k)), clears the
 * records and returns the number of kernel ids (<= capacity), or a negative status. */
int pcr_profile(pcr_ctx* ctx, int enable);
int pcr_profile_read(pcr_ctx* ctx, double* ms_out, int64_t* count_out, int capacity);
const char* pcr_kernel_name(int kernel_id);

/* Diagnostics.  The standardisation (p - centre) / scale divides every coordinate of a frame by one scale; the
 * kernels hoist the reciprocal refinement of the hardware's IEEE division sequence out of the per-point loop
 * (scale_div in pcr_kernels.cuh).  This entry checks that shortcut against __fdiv_rn for EVERY binary32 dividend and
 * each of the n given divisors (host array) and writes the number of differing quotient bit patterns (expected 0).
 * Synchronous. */
int pcr_selftest_scale_div(pcr_ctx* ctx, const float* h_divisors, int n, uint64_t* mismatches);

#ifdef __cplusplus
}
#endif
#endif /* PCR_H_ */
