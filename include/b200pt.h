/* libb200pt -- C ABI of the B200-native (sm_100a) path-tracing hot path.
 *
 * This is the drop-in boundary for the Taichi hot path of fakhirsh/path-tracer-python.  The seam is
 * the compiled-array boundary between the reference's host compile step (L2) and its device
 * state + kernels (L1/L0), i.e. what `TaichiRenderer.__init__` hands to `_upload_to_gpu`
 * (reference: src/render_server/taichi_renderer/renderer.py:84-98).  Every entry point names the
 * reference interface it replaces (paths relative to src/render_server/taichi_renderer/).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.  Pointers documented "device" must be
 *     CUDA device pointers on the context's GPU (the Python host owns them as torch tensors);
 *     pointers documented "host" are ordinary host memory.  The library never frees either.
 *   - every function returns PT_OK (0) or a negative PT_ERR_* code; pt_last_error() gives the
 *     message of the calling thread's last failure.  (The reference prints and continues,
 *     scene_compiler.py:353; kernels never report.)
 *   - all launches go to the caller's `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream); no hidden synchronisation except where stated.
 *   - one context per GPU / host thread (thread-compatible, not thread-safe).
 *   - arithmetic: PT-F32 contract (strict IEEE binary32, contract-defined elementary functions
 *     and per-path random stream) -- see DESIGN.md section 3 and oracle/oracle_math.h.
 */
#ifndef B200PT_H
#define B200PT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PT_OK 0
#define PT_ERR_INVALID_ARG (-1)
#define PT_ERR_CUDA (-2)
#define PT_ERR_NO_SCENE (-3)
#define PT_ERR_WORKSPACE (-4)
#define PT_ERR_CAPACITY (-5)

#define PT_MAX_IMAGES 16 /* fields.py:117 MAX_IMAGE_TEXTURES */

/* Primitive / material / texture codes: scene_compiler.py:9-25 */
#define PT_PRIM_SPHERE 0
#define PT_PRIM_TRIANGLE 1
#define PT_PRIM_QUAD 2

typedef struct PtContext PtContext;

/* Per-primitive-type material/texture/medium arrays exactly as compile_materials returns them
 * (scene_compiler.py:254-439; fields.py:70-138).  All host pointers, n entries each. */
typedef struct PtHostMaterials {
    const int32_t* material_type;   /* 0 Lambert 1 Metal 2 Dielectric 3 Emissive 4 Isotropic */
    const float* albedo;            /* [n][3] */
    const float* fuzz;
    const float* ir;
    const float* emit;              /* [n][3] */
    const int32_t* texture_type;    /* 0 solid 1 checker 2 image 3 noise */
    const float* texture_scale;
    const float* color1;            /* [n][3] */
    const float* color2;            /* [n][3] */
    const int32_t* image_idx;
    const int32_t* is_medium;
    const float* medium_density;
    const float* medium_albedo;     /* [n][3] */
} PtHostMaterials;

/* The reference-layout scene: output of compile_scene (scene_compiler.py:931-965) and compile_bvh
 * (bvh_compiler.py:132-168) plus the Perlin tables (fields.py:145-153) and image textures
 * (renderer.py:202-214; 8-bit RGB, texel = u8/255.0f as util/rtw_image.py:66). Host pointers. */
typedef struct PtHostScene {
    int32_t num_spheres, num_quads, num_triangles, num_nodes, num_images;
    const float* sphere_data;                                   /* [n][4] cx cy cz r */
    const float *quad_Q, *quad_u, *quad_v, *quad_normal, *quad_D, *quad_w;
    const float *tri_v0, *tri_edge1, *tri_edge2, *tri_normal;
    PtHostMaterials mat[3];                                     /* by prim type: sphere, triangle, quad */
    const float *bvh_bbox_min, *bvh_bbox_max;                   /* [nodes][3] */
    const int32_t *bvh_left, *bvh_right, *bvh_prim_type, *bvh_prim_idx;
    const float* perlin_randvec;                                /* [256][3] */
    const int32_t *perm_x, *perm_y, *perm_z;                    /* [256] */
    const uint8_t* images[PT_MAX_IMAGES];                       /* [h][w][3] */
    int32_t image_w[PT_MAX_IMAGES], image_h[PT_MAX_IMAGES];
    /* "full" feature mode -- what the Taichi path dropped but src/core renders: */
    const float* sphere_motion;                                 /* [n][3] centre(1)-centre(0) (core/sphere.py:22-36) or NULL */
    int32_t full_features;                                      /* bit field; 0 = taichi_compat.  Bit 0: moving spheres + material
                                                                 * types 5/6 (subsurface_simple / _volumetric, core/material.py:
                                                                 * 144-276).  Bit 1: next-event estimation at Lambertian hits
                                                                 * (the technique docs/optimizations.md:53 claims and
                                                                 * docs/__optimizations.md:7-10 lists as missing) */
} PtHostScene;

/* Layout of the packed device blob.  Offsets are bytes from blob start,
 * every section 256-byte aligned. */
typedef struct PtPackedInfo {
    uint64_t total_bytes;
    uint64_t off_nodes;      /* num_inner * node_bytes : the children's boxes + child refs of every inner node */
    uint64_t off_spheres;    /* num_spheres * 16 B */
    uint64_t off_quads;      /* num_quads * 64 B */
    uint64_t off_tris;       /* num_triangles * 48 B */
    uint64_t off_prim_mat;   /* (num_spheres+num_triangles+num_quads) * 4 B material ids */
    uint64_t off_materials;  /* num_materials * 48 B de-duplicated records */
    uint64_t off_perlin;     /* 256*16 B vectors + 768 B permutations */
    uint64_t off_images;     /* RGBA8 texels, image after image */
    uint64_t image_off[PT_MAX_IMAGES]; /* from off_images */
    int32_t image_w[PT_MAX_IMAGES], image_h[PT_MAX_IMAGES];
    int32_t num_inner, num_spheres, num_quads, num_triangles, num_materials, num_images;
    int32_t max_depth_tree;  /* depth of the tree (stack bound check) */
    int32_t has_media;       /* any is_medium flag set */
    uint64_t off_sphere_motion; /* num_spheres * 16 B, present iff full_features */
    int32_t full_features;
    int32_t node_bytes;      /* bytes per packed node: 64 (two children per node) or 128 (the 4-wide build, -DPT_WIDE=1) */
    uint64_t off_lights;     /* num_lights * 32 B: (leaf ref, area, emit.rg)(emit.b, -, -, -) per emissive primitive */
    int32_t num_lights;
    float nee_r2;            /* squared switch-over distance of the next-event estimator: 0.25 * mean emitter area */
} PtPackedInfo;

/* Camera + render settings: what _upload_camera writes (renderer.py:230-247; fields.py:159-172) */
typedef struct PtCamera {
    float center[3], pixel00[3], delta_u[3], delta_v[3], defocus_disk_u[3], defocus_disk_v[3];
    float defocus_angle;
    float background[3];
    int32_t max_depth;
    int32_t jitter;          /* sub-pixel offset of get_ray (kernels.py:183-184): 0 = two draws of the path's random stream, as the
                              * reference; 1 = the sample's point of a per-pixel XOR-scrambled Sobol' (0,2)-sequence (low-discrepancy
                              * sampling, docs/__optimizations.md:7-10): same expected image, better-stratified pixel footprints */
} PtCamera;

typedef struct PtCounters {
    uint64_t rays;        /* BVH traversals (primary + bounce + medium-exit searches) */
    uint64_t box_tests;   /* child-box slab tests (32 B of node data each) */
    uint64_t sphere_tests, triangle_tests, quad_tests;
    uint64_t paths;
    uint64_t bounces;     /* path segments */
} PtCounters;

const char* pt_last_error(void);
const char* pt_version(void);

/* ---- host side of the seam: replaces _upload_to_gpu's per-element field writes
 *      (renderer.py:102-228) with one packed blob that the caller copies to the device. */
int pt_pack_scene_info(const PtHostScene* scene, PtPackedInfo* info);
int pt_pack_scene(const PtHostScene* scene, const PtPackedInfo* info, void* host_blob);

/* ---- native binned-SAH builder: replaces SAHBVHBuilder.build + flatten
 *      (sah_bvh_builder.py:167-418) and produces bit-identical arrays.  Inputs are per-primitive
 *      float32 boxes/centroids as add_sphere/add_quad/add_triangle compute them (:104-165).
 *      Outputs must hold 2*n-1 nodes.  Returns the node count (>= 0) or an error code. */
int pt_bvh_build_sah(int32_t n_prims, const float* prim_bbox_min, const float* prim_bbox_max,
                     const float* prim_centroid, const int32_t* prim_type, const int32_t* prim_idx,
                     float* out_bbox_min, float* out_bbox_max, int32_t* out_left, int32_t* out_right,
                     int32_t* out_parent, int32_t* out_prim_type, int32_t* out_prim_idx);

/* ---- context */
int pt_create(int device_ordinal, PtContext** out_ctx);   /* replaces ti.init (renderer.py:16) */
int pt_destroy(PtContext* ctx);
/* Scheduling knobs -- none of them changes a result bit (tests/test_gpu_renderer.py):
 *   "block_threads" (megakernel CTA size), "smem_nodes" (inner nodes staged in shared memory; only acts in builds with -DPT_SNODES=1: the default build fetches every node through L1, which measured faster),
 *   "burst_num" / "wf_burst_num" (a warp leaves the traversal loop when fewer than n/32 of the lanes
 *   that entered it still traverse), "vote_num" / "wf_vote_num" (the warp visits a node when at least
 *   n lanes want to, else it tests primitives; 0 = while-while order), "sync_samples",
 *   "wf_blocks_per_sm", "wf_extend_minb" (4/5/6: register cap of k_wf_extend), "wf_waves" (full waves
 *   before the tail kernel), "wf_fuse_primary" (leading waves run inside k_wf_primary), "wf_tile"
 *   (8x4 pixel tiles in the primary wave), "wavefront_sort" (material binning, 0/1),
 *   "wf_overlap" / "wf_pipelines" (1..4 batches of a call in flight on context-owned streams; default 2),
 *   "wf_primary_threads", "wf_grid_waves" (grid size of the statically split kernels in waves of resident CTAs),
 *   "wf_tail_vote" (the tail as a persistent voting kernel), "wf_rel_primary" / "wf_oct_nodes" (origin-relative boxes /
 *   direction-octant variants of the node array for the primary wave, built per call in the workspace),
 *   "wf_shade_split" (one shade kernel per class run), "wf_shade_stage" (k_wf_shade collects its survivors per warp in
 *   shared memory and appends two rounds with one atomic on the queue counter; default 1), "wf_extend_bulk" (ray staging of k_wf_extend with
 *   cp.async.bulk + mbarrier instead of cp.async), "wf_extend_carveout" (its preferred shared-memory carveout in
 *   percent, -1 = driver's choice) */
int pt_set_option(PtContext* ctx, const char* name, int64_t value);

/* `device_blob` (device, info->total_bytes) stays owned by the caller and must outlive its use. */
int pt_set_scene(PtContext* ctx, const void* device_blob, const PtPackedInfo* info);
int pt_set_camera(PtContext* ctx, const PtCamera* cam);  /* replaces _upload_camera, renderer.py:230 */

/* Row-interleaved windows for the multi-GPU tile split (no reference counterpart; SURVEY.md 8e "use
 * interleaved row blocks"): after pt_set_row_interleave(ctx, B, S) the window [y0, y1) of the render
 * calls means y1 - y0 LOCAL rows (a multiple of B), local row ly being image row
 * y0 + (ly / B) * S + ly % B -- rank r of N passes y0 = r * B, S = N * B and renders blocks r, r + N, ...
 * in ONE call.  Rows beyond H are skipped.  B: power of two >= 4; block_rows <= 0 switches it off. */
int pt_set_row_interleave(PtContext* ctx, int block_rows, int stride_rows);

/* ---- megakernel: replaces `render_sample` (kernels.py:1177-1187) launched once per sample by
 *      render() (renderer.py:405-409).  Adds samples [sample_begin, sample_begin+sample_count) of
 *      rows [y0,y1) to accum (device, H*W*3 f32, row-major like fields.accum_buffer). */
int pt_render_megakernel(PtContext* ctx, float* accum, int W, int H, int y0, int y1, int sample_begin,
                         int sample_count, uint32_t seed, void* stream);

/* ---- wavefront: replaces generate_camera_rays / intersect_rays / shade_miss_rays /
 *      reset_next_ray_count / shade_and_scatter / swap_ray_buffers (kernels.py:1219-1418) and
 *      their host loop (renderer.py:305-334).  workspace: device, >= pt_wavefront_workspace_bytes. */
size_t pt_wavefront_workspace_bytes(int W, int rows);   /* minimum: one sample of the window in flight */
/* workspace for `samples_in_flight` samples of every pixel in flight at once (bigger waves: fewer,
 * fuller launches).  pt_render_wavefront uses as many as the workspace it is given can hold. */
size_t pt_wavefront_workspace_bytes_for(int W, int rows, int samples_in_flight);
/* Give the call 64 B per inner BVH node on top of that and the primary wave of a defocus-free camera
 * runs against origin-relative node boxes (a per-call scratch copy; without the room the plain test
 * is used).  With more than one batch per call the tail of each batch overlaps the next batch's
 * primary wave on a second, context-owned stream that is joined before the call returns to `stream`. */
int pt_render_wavefront(PtContext* ctx, float* accum, int W, int H, int y0, int y1, int sample_begin,
                        int sample_count, uint32_t seed, void* workspace, size_t workspace_bytes, void* stream);

/* ---- stage-level entry points: ONE batch -- `samples` samples of every pixel of the window,
 *      sample indices sample_base .. -- driven one wavefront kernel at a time, as the reference's host loop drives
 *      its kernels (renderer.py:305-334): generate, then per bounce extend + shade until no ray is left, then resolve.
 *      The batch lives in the caller's workspace (>= pt_wavefront_workspace_bytes_for(W, rows, samples)); pt_wf_begin
 *      binds it to the context.  Queue q = bounce & 1 is the one wave `bounce` reads.
 *        pt_wf_generate  = generate_camera_rays (kernels.py:1219-1239): camera rays of the batch -> queue 0
 *        pt_wf_extend    = intersect_rays (:1242-1263) for the rays of queue bounce & 1, plus the exit search of
 *                          apply_constant_medium (:417-419) for hits on a medium boundary -> hit records
 *        pt_wf_shade     = shade_miss_rays + reset_next_ray_count + shade_and_scatter (:1266-1399): material binning,
 *                          radiance contributions, survivors appended (compacted) to the other queue; the reference's
 *                          swap_ray_buffers copy (:1402-1418) is the role swap of the two queues
 *        pt_wf_compact   = the host's read of active_ray_count (renderer.py:315): live rays in the queue that wave
 *                          `bounce` reads (compaction itself is part of pt_wf_shade's append); synchronises `stream`
 *        pt_wf_read_rays / pt_wf_read_hits = the queue / the hit records as plain device arrays ([n][3] floats, pixel
 *                          index = fields.ray_pixel_index, prim = (type << 28) | index or -1); any pointer may be NULL
 *        pt_wf_resolve   = adds the batch's radiance to accum (H*W*3 f32), sample after sample (kernels.py:1280/:1373) */
int pt_wf_begin(PtContext* ctx, int W, int H, int y0, int y1, int sample_base, int samples, uint32_t seed, void* workspace,
                size_t workspace_bytes);
int pt_wf_generate(PtContext* ctx, void* stream);
int pt_wf_extend(PtContext* ctx, int bounce, void* stream);
int pt_wf_shade(PtContext* ctx, int bounce, void* stream);
int pt_wf_compact(PtContext* ctx, int bounce, int* live_rays, void* stream);
int pt_wf_read_rays(PtContext* ctx, int bounce, int n, float* o, float* d, float* throughput, int32_t* pixel, int32_t* depth, void* stream);
int pt_wf_read_hits(PtContext* ctx, int n, float* t, int32_t* prim, float* t_exit, void* stream);
int pt_wf_resolve(PtContext* ctx, float* accum, void* stream);

/* ---- stage-level probes used by the parity tests */
/* get_ray for one sample of every pixel (kernels.py:176-201, :1219-1239); o,d: device [H*W][3] */
int pt_generate_camera_rays(PtContext* ctx, int W, int H, uint32_t seed, int sample, float* o, float* d, void* stream);
/* traverse_bvh over a ray list (intersect_rays, kernels.py:1242-1263): out_t = closest t (t_max on a
 * miss); out_prim = (prim_type << 28) | prim_idx, or -1 on a miss; out_p/out_n may be NULL. */
int pt_first_hit(PtContext* ctx, const float* o, const float* d, int n, float t_min, float t_max, float* out_t,
                 int32_t* out_prim, float* out_p, float* out_n, void* stream);
/* the contract's elementary functions: fn 0 sin 1 cos 2 log 3 acos 4 atan2(a,b) 5 pow(a,5);
 * fn 6: a/b unused, out[i] = i-th draw of the random stream with state seed (uint32 bits of a[0]);
 * fn 7 exp; fn 8 the slab test's reciprocal direction (|x| > 1e-8 ? 1/x : 1e8) of a[i] (even i) or
 * b[i] (odd i) */
int pt_math_eval(PtContext* ctx, int fn, int n, const float* a, const float* b, float* out, void* stream);

/* ---- roofline probe (no reference counterpart: SURVEY.md section 8d asks for the measured L2 /
 *      L1 roof of the node-fetch stream).  Issues the traversal kernels' node-fetch sequence -- four
 *      16-byte read-only loads of one 64-byte record -- at full rate, every lane active, from `buffer`
 *      (device, `bytes` bytes, finite floats): pattern 0 = every lane its own pseudo-random record,
 *      1 = one record per warp.  4 * iters records per lane per launch, `repeats` timed launches after
 *      one warm-up; returns the best gathered GB/s.  Synchronises `stream`. */
int pt_probe_gather(PtContext* ctx, const void* buffer, size_t bytes, int pattern, int iters, int repeats,
                    double* out_gbytes_per_s, void* stream);

/* ---- output: replaces LivePreview.buffer_to_image (preview.py:117-132) on the device */
int pt_tonemap(PtContext* ctx, const float* accum, int n_floats, int sample_count, uint8_t* out_rgb, void* stream);

/* ---- statistics (synchronises the context's device).  Counting is off unless enabled; the
 *      counting kernels are separate instantiations so the timed path carries no atomics. */
int pt_enable_counters(PtContext* ctx, int enable);
int pt_get_counters(PtContext* ctx, PtCounters* out, int reset);

/* ---- debug-bounds build (libb200pt_dbg.so, -DPT_DEBUG_BOUNDS=1): the kernels check every node / primitive /
 *      material / texel / queue / bin / slab / stack index before use and record the first violation.  Returns it
 *      (check code << 24 | source line of pt_device.cuh / b200pt.cu; 0 = none) and clears it; always 0 in the
 *      product builds, whose kernels carry no checks.  Synchronises the device.  (No reference counterpart:
 *      Taichi's debug=True bounds checking, renderer.py:16, is the closest.) */
int pt_get_debug_error(PtContext* ctx, uint32_t* out_code);

/* ---- per-stage device times of the LAST pt_render_wavefront call (what renderer.py:305-334 prints
 *      as its per-bounce timing): when enabled, every kernel launch of the call is bracketed by CUDA
 *      events on the launching stream.  pt_get_stage_times synchronises the device, writes up to
 *      max_entries records in launch order and returns the total number in *n_entries. */
#define PT_STAGE_RESET 0
#define PT_STAGE_GENERATE 1
#define PT_STAGE_PRIMARY 2
#define PT_STAGE_EXTEND 3
#define PT_STAGE_BIN 4
#define PT_STAGE_SHADE 5
#define PT_STAGE_TAIL 6
#define PT_STAGE_RESOLVE 7
typedef struct PtStageTime {
    int32_t kernel; /* PT_STAGE_* */
    int32_t wave;   /* bounce index of the launch (0 for per-batch stages) */
    float ms;
} PtStageTime;
int pt_enable_stage_timing(PtContext* ctx, int enable);
int pt_get_stage_times(PtContext* ctx, PtStageTime* out, int max_entries, int* n_entries);
/* with counting AND stage timing enabled: the cumulative counters as they stood after each launch of
 * the last call (same order as pt_get_stage_times); differences give per-launch ray / test counts. */
int pt_get_stage_counters(PtContext* ctx, PtCounters* out, int max_entries, int* n_entries);

#ifdef __cplusplus
}
#endif
#endif /* B200PT_H */
