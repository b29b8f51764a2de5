// libb200pt: CUDA kernels (sm_100a) and the C ABI declared in include/b200pt.h.
//
// Megakernel  : one persistent lane per pixel runs ALL samples of the launch as a state machine
//               (generate -> traverse -> shade / medium-exit search -> regenerate), accumulating in
//               registers in sample order; replaces kernels.py:1177-1187 + trace_ray :1024-1170.
// Wavefront   : generate / extend / shade over SoA float4 ray queues in HBM, compaction and
//               material binning with warp-aggregated atomics, no host read-back inside a sample;
//               replaces kernels.py:1219-1418 + the host loop renderer.py:305-334.
// Compile with: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false (PT-F32 contract).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "pt_device.cuh"

using namespace pt;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
extern "C" std::string* pt_error_slot(void);   // pt_host.cpp: one thread-local slot for the whole library
static int set_error(int code, const std::string& msg) {
    *pt_error_slot() = msg;
    return code;
}
#define PT_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return set_error(PT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

#ifndef PT_MEGA_THREADS
#define PT_MEGA_THREADS 256
#define PT_MEGA_MINB 3
#endif
// launch bound of k_wf_primary: 128-thread CTAs x 7 per SM cap it at 72 registers = 28 warps per SM; the
// kernel waits mostly on fixed-latency dependencies, which more warps hide (256 x 3 = 80 registers,
// 24 warps: 5.08 ms per batch; 128 x 7: 4.94 ms with 28 B of spills; 256 x 4 = 64 registers: 212 B of spills)
#ifndef PT_PRIMARY_MINB
#define PT_PRIMARY_MINB 7
#endif
#ifndef PT_PRIMARY_THREADS
#define PT_PRIMARY_THREADS 128
#endif

constexpr int kMaxPipes = 4;
struct WfParams;
struct PtContext {
    int device = 0;
    bool has_scene = false;
    bool has_camera = false;
    DScene scene{};
    PtPackedInfo info{};
    PtCamera cam{};
    int block_threads = PT_MEGA_THREADS;
    int smem_nodes = -1;      // -1: auto
    int wavefront_sort = 1;
    int burst_num = 12;      // megakernel: leave a burst when < 12/32 of the entering lanes still traverse
    int sync_samples = 0;
    int vote_num = 12;       // megakernel: descend when >= 12 lanes want to (tools/sweep.py, config 2)
    int wf_vote_num = 16;    // the same threshold for k_wf_extend (tools/sweep_wf.py)
    int wf_waves = 9;        // full waves per batch before the tail kernel takes over (config 2, final build: 8 / 9 / 10 waves = 3209 / 3225 / 3221 Mpaths/s)
    int wf_primary_threads = PT_PRIMARY_THREADS;   // CTA size of k_wf_primary (<= its launch bound)
    int wf_overlap = 1;      // two pipelines: consecutive batches on two streams with their own queue sets (pt_render_wavefront)
    int wf_pipelines = 2;    // batches in flight at once when wf_overlap is on (each with its own stream and queue set)
    cudaStream_t aux_stream[kMaxPipes] = {};          // [0] unused: pipeline 0 runs on the caller's stream
    cudaEvent_t ev_resolve[kMaxPipes] = {}, ev_start = nullptr;
    int wf_tail_vote = 1;    // the tail as a persistent voting state machine (k_wf_tail_vote) instead of one lane per ray
    int wf_rel_primary = 1;  // camera rays without defocus: slab test against origin-relative node boxes
    int wf_oct_nodes = 1;    // direction-octant variants of the node array (pt_device.cuh trav_root), built per call in the workspace
    int wf_grid_waves = 16;  // static-partition kernels launch this many full waves of resident CTAs
    int wf_shade_stage = 1;  // k_wf_shade<FULL, kShadeStage>: survivors staged per warp in shared memory, one queue-counter atomic per ~4 rounds
    int dbg_shade_smem = 0;  // experiment knob: dynamic shared memory reserved per k_wf_shade CTA (limits residency)
    int wf_fuse_primary = 1; // wave 0 (camera rays) as one fused kernel
    int wf_burst_num = 20;
    int wf_blocks_per_sm = 7;
    int wf_extend_minb = PT_FAST_SLAB ? 6 : 5;  // launch bound of k_wf_extend: the one that compiles to <= 56 registers without spills, so
                                           // that six 192-thread CTAs (36 warps) are resident per SM (5 -> 56 in both numerics; with the FMA
                                           // slab test, PT_FAST_SLAB: 6 -> 55, 5 -> 61)
    int wf_extend_carveout = -1;   // preferred shared-memory carveout of k_wf_extend in percent of the SM's 228 KB (-1: the driver's choice)
    int wf_extend_bulk = 0;  // ray staging of k_wf_extend with cp.async.bulk + mbarrier (k_wf_extend<..., BULK>; measured -1 %)
    int wf_tile = 1;
    int wf_shade_split = 0;  // 1: k_wf_shade_cls<KIND> per class run instead of one k_wf_shade for all classes (measured slower:
                             // 5.99 vs 5.12 ms per config-2 batch -- three dependent launches, each with its own ramp and tail,
                             // where one kernel lets warps of different classes hide each other's memory latency)
    int row_shift = -1, row_stride = 0;   // pt_set_row_interleave
    int num_sms = 148;
    int max_smem_optin = 0;
    WfParams* stage_P = nullptr; // pt_wf_begin: the batch bound for the stage-level entry points
    bool count = false;
    DCounters* d_counters = nullptr;
    // stage timing (pt_enable_stage_timing): events around every launch of the last render call
    bool stage_timing = false;
    std::vector<cudaEvent_t> ev_pool;
    struct StageMark { int kernel, wave, ev_begin, ev_end; };
    std::vector<StageMark> marks;
    int ev_used = 0;
    DCounters* d_stage_snap = nullptr;      // counters after every timed launch (counting + timing both on)
};
static const int kMaxStageSnaps = 4096;

// ------------------------------------------------------------------------------------------------
// shared helpers
// ------------------------------------------------------------------------------------------------
// (PT_SNODES=1 builds only: copies the top n_snodes nodes into the CTA's shared memory; the default build
// fetches every node through L1 and this is a no-op without a barrier.)
__device__ __forceinline__ void stage_nodes(float4* s_node_mem, const DScene& S, int n_snodes) {
#if PT_SNODES
    const int n4 = n_snodes * 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) s_node_mem[i] = __ldg(S.nodes + i);
    __syncthreads();
#else
    (void)s_node_mem; (void)S; (void)n_snodes;
#endif
}

__device__ __forceinline__ void flush_counters(DCounters* dc, const LocalCounters& lc) {
    // one atomic per warp per counter (warp-aggregated)
    unsigned v[7] = {lc.rays, lc.box_tests, lc.sphere_tests, lc.triangle_tests, lc.quad_tests, lc.paths, lc.bounces};
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(dc);
#pragma unroll
    for (int k = 0; k < 7; k++) {
        unsigned long long s = v[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(dst + k, s);
    }
}

// The constant-medium decision once the exit search is done (kernels.py:421-448). Returns true if
// the ray scatters inside the medium (t_scatter valid).
__device__ __forceinline__ bool medium_sample(bool hit_exit, float t_exit_hit, float t_entry, f3 d, float density,
                                              ptm::Rng& rng, float& t_exit, float& t_scatter) {
    t_exit = 0.0f;
    t_scatter = 0.0f;
    if (!hit_exit) return false;
    t_exit = t_exit_hit;
    float t1 = max2(t_entry, 0.001f);
    float t2 = min2(t_exit, 1e10f);
    if (t1 < t2) {
        if (t1 < 0.0f) t1 = 0.0f;
        float ray_length = ptm::fsqrt(dot(d, d));
        float distance_inside = (t2 - t1) * ray_length;
        float hit_distance = ptm::fdiv(-ptm::log_c(max2(ptm::rng_next(rng), 1e-10f)), density);
        if (hit_distance < distance_inside) {
            t_scatter = t1 + ptm::fdiv(hit_distance, ray_length);
            return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// megakernel
// ------------------------------------------------------------------------------------------------
// Row-interleaved windows (the multi-GPU tile split, distributed.py): the window's local row ly is the
// image row  y0 + (ly >> row_shift) * row_stride + (ly & (block - 1)),  block = 1 << row_shift rows, so
// that rank r of N renders blocks r, r + N, r + 2N, ... (row_stride = N * block) with ONE call; rows
// that fall beyond the image are skipped.  row_shift < 0: the plain contiguous window [y0, y1).
__device__ __forceinline__ int win_row(int y0, int row_shift, int row_stride, int ly) {
    return row_shift < 0 ? y0 + ly : y0 + (ly >> row_shift) * row_stride + (ly & ((1 << row_shift) - 1));
}

struct MegaParams {
    DScene S;
    PtCamera cam;
    float* accum;
    int W, H, y0, y1;
    int row_shift, row_stride;   // row interleave (win_row); row_shift < 0 = contiguous window
    int sample_begin, sample_count;
    uint32_t seed;
    int n_snodes;
    int tiles_x;
    int burst_num;       // leave a traversal burst when < burst_num/32 of the live lanes still traverse
    int sync_samples;    // > 0: lanes of a warp re-synchronise every sync_samples samples
    int vote_num;        // > 0: warp-voting traversal; descend when >= vote_num/32 of the traversing lanes want to
    DCounters* counters;
};

// Lane states of the persistent megakernel
enum : int { LANE_NEED_PATH = 0, LANE_HAVE_RESULT = 1, LANE_TRAVERSING = 2, LANE_DONE = 3, LANE_WAIT = 4, LANE_BEGIN = 5 };

template <bool COUNT, bool FULL>
__global__ void __launch_bounds__(PT_MEGA_THREADS, PT_MEGA_MINB) k_megakernel(const __grid_constant__ MegaParams P) {
    extern __shared__ float4 s_node_mem[];
    const DScene& S = P.S;
    stage_nodes(s_node_mem, S, P.n_snodes);
    const SNodeBase s_nodes = snode_base(s_node_mem);

    // one warp = one 8x4 pixel tile (coherent primary rays)
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int tx = gwarp % P.tiles_x, ty = gwarp / P.tiles_x;
    const int px = tx * 8 + (lane & 7);
    const int ly = ty * 4 + (lane >> 3);
    const int py = win_row(P.y0, P.row_shift, P.row_stride, ly);
    const bool valid = (px < P.W) && (ly < P.y1 - P.y0) && (py < P.H);
    const int pixel = py * P.W + px;
    float* ap = P.accum + 3 * (size_t)(valid ? pixel : 0);
    LocalCounters lc;
    f3 acc = mk3(0.0f, 0.0f, 0.0f);
    if (valid) acc = mk3(ap[0], ap[1], ap[2]);
    const int max_depth = P.cam.max_depth;
    const f3 bg = mk3(P.cam.background[0], P.cam.background[1], P.cam.background[2]);
    int s = P.sample_begin;
    const int s_end = P.sample_begin + P.sample_count;

    // per-path state
    f3 thr = mk3(1.0f, 1.0f, 1.0f), col = mk3(0.0f, 0.0f, 0.0f);
    ptm::Rng rng{0u, 1u};
    float time = 0.0f;       // shutter time of the current path (full feature mode)
    int depth = 0;
    int prev_nee = 0;        // features "nee": 1 = the last scatter was a Lambertian vertex with next-event estimation,
                             // 2 = ... and the ray has passed through a medium since (nee_direct, pt_device.cuh)
    int mode = 0;            // 0: closest-hit search, 1: medium-exit search (kernels.py:417)
    float t_entry = 0.0f;
    int ref_entry = kRefNone;
    // traversal state (persists across bursts)
    TravState T;
    T.cur = kRefNone;
    StackEntry stack_mem[kStackSlots];
    const LocalStack stack{stack_mem};
    stack_init(stack);
    int st = (valid && s < s_end && max_depth > 0) ? LANE_NEED_PATH : LANE_DONE;

    for (;;) {
        // ---- (a) lanes whose traversal finished: medium logic / shading / Russian roulette
        if (st == LANE_HAVE_RESULT) {
            if (mode == 0) trav_finish<FULL>(T, S);         // equal-t hits: the primitive the reference would have kept
            const f3 o = T.o, d = T.d;
            const float t = T.best_t;
            const int ref = T.best_ref;
            bool end_path = false, do_shade = false, scattered = false, passthrough = false, restart = false;
            int sref = ref;
            float stt = t;
            f3 hp = o, sdir = d, atten = mk3(1.0f, 1.0f, 1.0f);
            Mat m;
            if (mode == 0) {
                if (ref == kRefNone) {
                    col = col + thr * bg;                   // kernels.py:1166
                    end_path = true;
                } else {
                    m = load_mat(S, mat_id_of(S, ref));
                    if (mat_is_medium(m.header)) {
                        mode = 1;                           // the exit search is simply the next traversal
                        t_entry = t;
                        ref_entry = ref;
                        trav_restart<COUNT>(T, S, t + 0.0001f, 1e10f, lc);
                        restart = true;
                    } else {
                        do_shade = true;
                    }
                }
            } else {
                mode = 0;
                const int mid = mat_id_of(S, ref_entry);
                const float4 m2 = __ldg(S.materials + 3 * mid + 2);   // (density, medium albedo)
                float t_exit, t_scatter;
                if (medium_sample(ref != kRefNone, t, t_entry, d, m2.x, rng, t_exit, t_scatter)) {
                    hp = o + t_scatter * d;                 // kernels.py:446
                    sdir = random_unit_vector(rng);         // kernels.py:1088
                    atten = mk3(m2.y, m2.z, m2.w);
                    scattered = true;
                } else if (t_exit > 0.0f) {
                    passthrough = true;                     // kernels.py:1100-1110
                    float eps_t = ptm::fdiv(0.001f, ptm::fsqrt(dot(d, d)));
                    hp = o + d * (t_exit + eps_t);
                } else {
                    m = load_mat(S, mid);                   // fallback: shade the boundary, :1111-1119
                    sref = ref_entry;
                    stt = t_entry;
                    do_shade = true;
                }
            }
            bool did_nee = false;
            if (do_shade) {
                f3 n, emit;
                finalize_hit<FULL>(S, sref, o, d, stt, hp, n, time);
                if (FULL && nee_on(S) && mat_type(m.header) == 0) {       // direct light at a Lambertian vertex, before its scatter draws
                    col = col + thr * nee_direct<FULL>(S, hp, n, eval_texture(S, m, sref, hp), rng, time);
                    did_nee = true;
                }
                scattered = shade_surface<FULL>(S, m, sref, d, hp, n, rng, emit, sdir, atten);
                // kernels.py:1116/1124; after a vertex with next-event estimation only the close pairs are taken from here
                if (!(FULL && prev_nee != 0) || (prev_nee == 1 && (stt * stt) * dot(d, d) < S.nee_r2)) col = col + thr * emit;
            }
            if (FULL && !restart) prev_nee = passthrough ? (prev_nee != 0 ? 2 : 0) : (did_nee ? 1 : 0);
            if (!restart) {
                if (!end_path) {
                    if (scattered) {
                        thr = thr * atten;
                        if (depth + 1 >= max_depth) {
                            end_path = true;
                        } else if (depth + 1 >= 5) {        // Russian roulette, kernels.py:1145-1157
                            float survival = min2(max2(max2(thr.x, thr.y), thr.z), 0.95f);
                            if (ptm::rng_next(rng) > survival) end_path = true;
                            else thr = vdiv(thr, survival);
                        }
                        depth++;
                    } else if (passthrough) {
                        depth++;                            // the for-loop index still advances
                        if (depth >= max_depth) end_path = true;
                    } else {
                        end_path = true;                    // absorbed
                    }
                }
                if (end_path) {
                    acc = acc + col;                        // kernels.py:1187
                    s++;
                    // optional warp re-synchronisation every sync_samples samples: the lanes of a
                    // tile start their next primary rays together (coherent again)
                    st = (P.sync_samples > 0 && s < s_end && ((s - P.sample_begin) % P.sync_samples) == 0) ? LANE_WAIT
                                                                                                           : LANE_NEED_PATH;
                } else {
                    if (COUNT) lc.bounces++;
                    T.o = hp;
                    T.d = sdir;
                    st = LANE_BEGIN;
                }
            } else {
                st = LANE_TRAVERSING;
            }
        }
        // ---- (b) path regeneration: the lane moves on to its next sample
        if (st == LANE_NEED_PATH) {
            if (s < s_end) {
                rng = ptm::rng_seed(P.seed, (uint32_t)pixel, (uint32_t)s);
                time = FULL ? ptm::path_time(rng) : 0.0f;
                f3 o, d;
                get_ray(P.cam, px, py, rng, o, d, P.seed, (uint32_t)pixel, (uint32_t)s);
                d = normalized(d);                          // kernels.py:1042
                thr = mk3(1.0f, 1.0f, 1.0f);
                col = mk3(0.0f, 0.0f, 0.0f);
                depth = 0;
                prev_nee = 0;
                mode = 0;
                if (COUNT) { lc.paths++; lc.bounces++; }
                T.o = o;
                T.d = d;
                st = LANE_BEGIN;
            } else {
                st = LANE_DONE;
            }
        }
        // bounce rays and fresh camera rays start their traversal together (one copy of the set-up
        // code, executed once per iteration by both groups of lanes)
        if (st == LANE_BEGIN) {
            trav_begin<COUNT>(T, S, T.o, T.d, 0.001f, 1e10f, lc, time);
            st = LANE_TRAVERSING;
        }
        // ---- (c) traversal burst: leave when fewer than ~burst_frac of the live lanes still traverse
        const unsigned live = __ballot_sync(0xffffffffu, st != LANE_DONE);
        if (live == 0u) break;
        if (P.sync_samples > 0) {
            const unsigned busy = __ballot_sync(0xffffffffu, st == LANE_TRAVERSING || st == LANE_HAVE_RESULT || st == LANE_NEED_PATH);
            if (busy == 0u && st == LANE_WAIT) st = LANE_NEED_PATH;
        }
        // leave the burst when fewer than burst_num/32 of the lanes that entered it still traverse
        const int min_active = (__popc(__ballot_sync(0xffffffffu, st == LANE_TRAVERSING)) * P.burst_num + 31) >> 5;
        // vote_num <= 0 ("no voting") is the vote with threshold 1: descend whenever any lane wants to,
        // i.e. while-while order -- one copy of the traversal code in the kernel instead of two
        bool fin;
        trav_run_vote<COUNT, FULL>(T, st == LANE_TRAVERSING, fin, S, s_nodes, P.n_snodes, stack,
                             min_active > 0 ? min_active : 1, P.vote_num > 0 ? P.vote_num : 1, lc);
        if (fin) st = LANE_HAVE_RESULT;
    }
    PT_BOUND(!valid || (pixel >= 0 && pixel < P.W * P.H), PT_DBG_PIXEL);
    if (valid) { ap[0] = acc.x; ap[1] = acc.y; ap[2] = acc.z; }
    if (COUNT) flush_counters(P.counters, lc);
}

// ------------------------------------------------------------------------------------------------
// wavefront
// ------------------------------------------------------------------------------------------------
// A *batch* of B samples of every pixel of the window is in flight at once (B*N paths, N = W*rows;
// B is whatever fits the caller's workspace).  Per wave: k_wf_extend (closest hit, dynamic ray
// fetch) then k_wf_shade (material-binned, compaction).  A path's radiance contribution is written
// to slab[slot][pixel] and k_wf_resolve adds the slabs to the accumulation buffer in sample order,
// so the sums are the ones the reference's one-sample-at-a-time loop produces.
//
// Ray record (48 B): a 32-byte pair (o.xyz d.x)(d.yz thr.xy) -- what k_wf_extend reads, and one whole
// sector for the gather of k_wf_shade -- plus a float4 plane (thr.z path_id rng depth)
//   path_id = slot * N + local pixel.        Hit record (16 B): (t, ref, t_exit, exit_hit)
struct WfQueues {
    float4 *ra[2];           // [capacity][2]: (o.xyz d.x)(d.yz thr.xy), one 32-byte sector per ray
    float4 *r2[2];           // [capacity]: (thr.z path_id rng depth)
    float4* hit;
    uint32_t* bins;          // kNumBins regions of `capacity` ray indices (material binning)
    int* bin_count;          // [kNumBins]
    float* slab;             // [B][N][3] radiance contribution of each path of the batch
    int* qcount;             // [2] ray counts of the two queues
    int* cursor;             // dynamic-fetch cursor of k_wf_extend
    int* tie_list;           // [capacity] paths / rays whose closest hit is shared by several primitives (resolve_tie, pt_device.cuh)
    int* tie_count;          // [2]: [0] the fused primary wave's list, [1] this wave's k_wf_extend list
    int capacity;            // B * N
};
constexpr int kNodeBytes = 16 * kNodeQuads;   // one packed inner node: 64 B (two children) or 128 B (PT_WIDE: four)
constexpr int kNumBins = 8;  // one bin per shading class (pt_device.cuh PT_CLS_*)

struct WfParams {
    DScene S;
    PtCamera cam;
    WfQueues Q;
    float* accum;
    int W, H, y0, y1;
    int row_shift, row_stride;   // row interleave (win_row); row_shift < 0 = contiguous window
    float inv_W;             // 1.0f / W (win_pixel)
    int N;                   // W * (y1 - y0)
    float inv_N;             // 1.0f / N (wf_path_key)
    int sample_base;         // first sample index of this batch
    int batch;               // samples in this batch (<= B)
    uint32_t seed;
    int n_snodes;
    int bounce;              // parity selects queue
    int sort;
    int burst_num;
    int vote_num;
    int fused_waves;         // waves k_wf_primary runs in registers before handing paths to the queues
    int tile;                // k_wf_primary walks the window in 8x4-pixel tiles (one warp = one tile)
    DCounters* counters;
};

// Image pixel index of the window's local pixel (row-major within the window).
__device__ __forceinline__ int win_pixel(const WfParams& P, int local) {
    if (P.row_shift < 0) return P.y0 * P.W + local;
    int ly = (int)((float)local * P.inv_W);          // local / W by one float multiply + exact fix-up
    int px = local - ly * P.W;
    if (px < 0) { ly--; px += P.W; }
    else if (px >= P.W) { ly++; px -= P.W; }
    return win_row(P.y0, P.row_shift, P.row_stride, ly) * P.W + px;
}

// The key of the path that owns a queue entry -- (start, stride) of its random stream -- recomputed from
// its id: the ray record carries only the stream position, the stride is a pure function of
// (seed, pixel, sample).  path_id = slot * N + local pixel with slot <= 256, so the quotient comes
// from one float multiply and an exact fix-up instead of an integer division.
__device__ __forceinline__ ptm::Rng wf_path_key(const WfParams& P, int path_id) {
    int slot = (int)((float)path_id * P.inv_N);
    int local = path_id - slot * P.N;
    if (local < 0) { slot--; local += P.N; }
    else if (local >= P.N) { slot++; local -= P.N; }
    return ptm::rng_seed(P.seed, (uint32_t)win_pixel(P, local), (uint32_t)(P.sample_base + slot));
}
// full feature mode: the shutter time of that path
template <bool FULL>
__device__ __forceinline__ float wf_path_time(const WfParams& P, int path_id) {
    if (!FULL) return 0.0f;
    return ptm::path_time(wf_path_key(P, path_id));
}

__global__ void __launch_bounds__(256) k_wf_generate(const __grid_constant__ WfParams P) {
    const int total = P.N * P.batch;
    const int stride = gridDim.x * blockDim.x;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < total; id += stride) {
        const int slot = id / P.N;
        const int pixel = win_pixel(P, id - slot * P.N);
        const int py = pixel / P.W, px = pixel - py * P.W;
        ptm::Rng rng = ptm::rng_seed(P.seed, (uint32_t)pixel, (uint32_t)(P.sample_base + slot));
        f3 o, d;
        get_ray(P.cam, px, py, rng, o, d, P.seed, (uint32_t)pixel, (uint32_t)(P.sample_base + slot));     // NOT normalised, kernels.py:1229-1233
        P.Q.ra[0][2 * (size_t)id] = make_float4(o.x, o.y, o.z, d.x);
        P.Q.ra[0][2 * (size_t)id + 1] = make_float4(d.y, d.z, 1.0f, 1.0f);
        P.Q.r2[0][id] = make_float4(1.0f, __int_as_float(id), __uint_as_float(rng.s), __int_as_float(0));
        P.Q.slab[3 * (size_t)id] = 0.0f;
        P.Q.slab[3 * (size_t)id + 1] = 0.0f;
        P.Q.slab[3 * (size_t)id + 2] = 0.0f;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.Q.qcount[0] = total;
        P.Q.qcount[1] = 0;
        P.Q.cursor[0] = 0;
        P.Q.cursor[1] = 0;
        P.Q.tie_count[0] = 0;
        P.Q.tie_count[1] = 0;
    }
}

// warp-aggregated append: lanes with the same key share one atomicAdd
__device__ __forceinline__ int warp_agg_append(int* counter_base, int key, bool active) {
    const unsigned lane = threadIdx.x & 31;
    unsigned amask = __ballot_sync(0xffffffffu, active);
    int pos = 0;
    if (active) {
        unsigned peers = __match_any_sync(amask, key);
        int leader = __ffs(peers) - 1;
        int rank = __popc(peers & ((1u << lane) - 1));
        int base = 0;
        if ((int)lane == leader) base = atomicAdd(counter_base + key, __popc(peers));
        base = __shfl_sync(peers, base, leader);
        pos = base + rank;
    }
    return pos;
}

// ---- cp.async helpers (LDGSTS: global -> shared without passing through registers)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ---- bulk async copy (TMA, UBLKCP): ONE instruction from one lane moves a contiguous block global -> shared and reports
// its bytes to an mbarrier; the waiting side polls the barrier's phase parity.  k_wf_extend's ray staging as a variant
// (BULK): the 32 rays a warp stages next are one contiguous kilobyte of the queue, so the 64 LDGSTS of a
// staging step (two per lane) become one UBLKCP issued by lane 0 plus an expect-tx arrive, and the switch to the other
// buffer polls that buffer's mbarrier instead of cp.async.wait_group.  Bit-identical; measured on config 2 (tools/ab_head.sh,
// same box, three rounds): k_wf_extend 12.49 vs 12.25 ms per 32-sample batch, frame 3181 vs 3214 Mpaths/s (-1.0 %) -- the
// kernel is not short of the issue slots the LDGSTS take, the bulk copy's completion latency is the longer one, and the
// barrier words plus the phase bits cost a register under the 56-register bound.  Not the default: option "wf_extend_bulk"
// selects the k_wf_extend<..., BULK = true> instantiations (bit-identical in the options test).
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\t"
                 "bra WAIT_%=;\n\t"
                 "DONE_%=:\n\t}"
                 ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

constexpr int kWfReserve = 128;   // ray indices a warp reserves per atomic on the global cursor
constexpr int kWfWarps = 8;       // warps per CTA of the wavefront kernels (256 threads)

// intersect_rays (kernels.py:1242-1263) + the exit search of apply_constant_medium (:417-419).
// Persistent warps with dynamic ray fetch.  Each warp reserves kWfReserve consecutive ray indices
// with one atomic, stages them 32 at a time into shared memory with cp.async (double-buffered, so
// the HBM latency of the queue read overlaps traversal), and hands staged rays to lanes as they
// finish; the other lanes keep traversing (resumable warp-voting traversal).  Results go to the
// hit-record array in ray order: no atomics, no binning here.
#ifndef PT_EXT_WARPS
#define PT_EXT_WARPS 6       // warps per CTA of k_wf_extend: 192 threads x 6 resident CTAs = 36 warps per SM at 56
                             // registers without spills (256 x 4 = 32 warps; 256 x 5 = 40 warps at 48 registers with spills:
                             // 2812 / 2846 / 2881 Mpaths/s on config 2)
#endif
// Traversal-stack slots per lane that k_wf_extend keeps in shared memory (slot 0 = sentinel; HybridStack,
// pt_device.cuh).  Measured on config 2, extend per 32-sample frame: 1 slot 13.81 ms, 4: 13.12, 8: 13.26, 12: 13.60,
// 16: 14.04 (more slots take L1 capacity away from the node fetches); the all-local stack of round 1: 13.25.
#ifndef PT_EXT_SMEM_SLOTS
#define PT_EXT_SMEM_SLOTS 4
#endif
constexpr int kExtSmemSlots = PT_EXT_SMEM_SLOTS;
template <bool COUNT, bool FULL, int MINB, bool BULK = false>
__global__ void __launch_bounds__(PT_EXT_WARPS * 32, MINB) k_wf_extend(const __grid_constant__ WfParams P) {
    extern __shared__ float4 s_dyn[];
    __shared__ __align__(128) float4 s_ray[PT_EXT_WARPS][2][32][2];      // (o.xyz d.x)(d.yz ..) of staged rays, 2 KB per warp
    __shared__ __align__(8) unsigned long long s_bar[BULK ? PT_EXT_WARPS : 1][2];   // BULK: one mbarrier per staging buffer
    __shared__ StackEntry s_stack[kExtSmemSlots][PT_EXT_WARPS * 32];
    float4* s_node_mem = s_dyn;
    const DScene& S = P.S;
    const int cur = P.bounce & 1;
    const int count = P.Q.qcount[cur];
    // reset_next_ray_count (:1283) -- also when this wave is empty, otherwise the stale count of
    // the wave before last would be picked up again two launches later
    if (blockIdx.x == 0 && threadIdx.x == 0) P.Q.qcount[cur ^ 1] = 0;
    if (blockIdx.x == 0 && threadIdx.x < kNumBins) P.Q.bin_count[threadIdx.x] = 0;   // for this wave's k_wf_bin
    // small waves: only as many CTAs as there are 256-ray slices take part (the rest would just
    // hammer the cursor with atomics and stage nodes for nothing)
    if ((long long)blockIdx.x * (PT_EXT_WARPS * 32) >= count) return;
    stage_nodes(s_node_mem, S, P.n_snodes);
    const SNodeBase s_nodes = snode_base(s_node_mem);
    LocalCounters lc;
    const unsigned lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const float4* __restrict__ qa = P.Q.ra[cur];

    // warp-uniform staging state
    int res_next = 0, res_end = 0;      // reserved, not yet staged ray indices
    bool exhausted = false;             // the global cursor is past the end of the queue
    int stg_base[2] = {0, 0}, stg_n[2] = {0, 0};
    int buf = 0, stg_pos = 0;
    // small waves reserve 32 rays at a time so that the work spreads over all warps
    const int reserve = (long long)count >= (long long)gridDim.x * PT_EXT_WARPS * 2 * kWfReserve ? kWfReserve : 32;
    auto stage = [&](int b) {           // start the async copy of up to 32 more rays into buffer b
        if (res_next >= res_end && !exhausted) {
            int base = 0;
            if (lane == 0) base = atomicAdd(P.Q.cursor, reserve);
            base = __shfl_sync(0xffffffffu, base, 0);
            res_next = base;
            res_end = base + reserve < count ? base + reserve : count;
            if (base >= count) { exhausted = true; res_next = res_end = 0; }
        }
        int n = res_end - res_next;
        n = n > 32 ? 32 : (n < 0 ? 0 : n);
        if (BULK) {
            __syncwarp();                   // every lane has taken its ray out of buffer b
            if (lane == 0 && n > 0) {
                PT_BOUND(res_next >= 0 && res_next + n <= count, PT_DBG_QUEUE);
                mbar_expect_tx(&s_bar[BULK ? warp : 0][b], (unsigned)n * 32u);
                bulk_g2s(&s_ray[warp][b][0][0], qa + 2 * (size_t)res_next, (unsigned)n * 32u, &s_bar[BULK ? warp : 0][b]);
            }
        } else {
            if ((int)lane < n) {
                PT_BOUND(res_next >= 0 && res_next + (int)lane < count, PT_DBG_QUEUE);
                cp_async16(&s_ray[warp][b][lane][0], qa + 2 * (size_t)(res_next + lane));
                cp_async16(&s_ray[warp][b][lane][1], qa + 2 * (size_t)(res_next + lane) + 1);
            }
            cp_async_commit();
        }
        stg_base[b] = res_next;
        stg_n[b] = n;
        res_next += n;
    };
    unsigned phase = 0;                 // BULK, bit b: the parity buffer b's next completion will have
    if (BULK) {
        if (lane == 0) { mbar_init(&s_bar[BULK ? warp : 0][0], 1); mbar_init(&s_bar[BULK ? warp : 0][1], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        __syncwarp();
    }
    auto wait_staged = [&](int b) {     // the copy into buffer b has landed (BULK: only called when one was started, stg_n[b] > 0)
        if (BULK) {
            mbar_wait(&s_bar[BULK ? warp : 0][b], (phase >> b) & 1u);
            phase ^= 1u << b;
        } else {
            cp_async_wait_all();
            __syncwarp();
        }
    };
    stage(0);
    if (!BULK || stg_n[0] > 0) wait_staged(0);
    stage(1);

    TravState T;
    T.cur = kRefNone;
    // stack: the first kExtSmemSlots slots of every lane in shared memory, the rest in local memory
    StackEntry stack_mem[kStackSlots - kExtSmemSlots > 0 ? kStackSlots - kExtSmemSlots : 1];
    const HybridStack<kExtSmemSlots> stack{&s_stack[0][threadIdx.x], PT_EXT_WARPS * 32, stack_mem};
    stack_init(stack);
    int st = LANE_NEED_PATH;     // NEED_PATH: wants a ray; TRAVERSING; HAVE_RESULT; DONE: queue exhausted
    int ray = 0, mode = 0, ref_entry = kRefNone;
    float t_entry = 0.0f;

    for (;;) {
        // ---- finished lanes: second traversal for medium boundaries, else write the hit record
        if (st == LANE_HAVE_RESULT) {
            if (mode == 0) {
                // equal-t hits (pt_device.cuh resolve_tie): listed for k_wf_fix_ties, which rewrites the record where the
                // reference would have kept another primitive; the counting instantiation resolves in place
                if (COUNT) {
                    trav_finish<FULL>(T, S);
                } else if (trav_take_tie(T)) {
                    const int k = atomicAdd(P.Q.tie_count + 1, 1);
                    PT_BOUND(k >= 0 && k < P.Q.capacity, PT_DBG_QUEUE);
                    P.Q.tie_list[k] = ray;
                }
                const int ref = T.best_ref;
                // shading class of the hit = its material bin (0 = miss)
                const int cls = ref != kRefNone ? ref_class(ref) : 0;
                if (cls == PT_CLS_MEDIUM) {
                    mode = 1;
                    t_entry = T.best_t;
                    ref_entry = ref;
                    trav_restart<COUNT>(T, S, t_entry + 0.0001f, 1e10f, lc);
                    st = LANE_TRAVERSING;
                } else {
                    PT_BOUND(ray >= 0 && ray < count && count <= P.Q.capacity, PT_DBG_QUEUE);
                    P.Q.hit[ray] = make_float4(T.best_t, __int_as_float(ref), 0.0f, __int_as_float(cls << 8));
                    st = LANE_NEED_PATH;
                }
            } else {
                mode = 0;
                PT_BOUND(ray >= 0 && ray < count && count <= P.Q.capacity, PT_DBG_QUEUE);
                P.Q.hit[ray] = make_float4(t_entry, __int_as_float(ref_entry), T.best_t,
                                           __int_as_float((T.best_ref != kRefNone ? 1 : 0) | (PT_CLS_MEDIUM << 8)));
                st = LANE_NEED_PATH;
            }
        }
        // ---- hand staged rays to idle lanes
        unsigned want = __ballot_sync(0xffffffffu, st == LANE_NEED_PATH);
        while (want) {
            const int avail = stg_n[buf] - stg_pos;
            if (avail == 0) {
                if (stg_n[buf ^ 1] == 0) {               // nothing staged, nothing left to stage
                    if (st == LANE_NEED_PATH) st = LANE_DONE;
                    break;
                }
                wait_staged(buf ^ 1);
                buf ^= 1;
                stg_pos = 0;
                stage(buf ^ 1);
                continue;
            }
            const int n_want = __popc(want);
            const int take = avail < n_want ? avail : n_want;
            const int rank = __popc(want & ((1u << lane) - 1));
            if (st == LANE_NEED_PATH && rank < take) {
                const int slot = stg_pos + rank;
                const float4 a = s_ray[warp][buf][slot][0], b = s_ray[warp][buf][slot][1];
                ray = stg_base[buf] + slot;
                if (COUNT) lc.bounces++;
                const float time = FULL ? wf_path_time<FULL>(P, __float_as_int(P.Q.r2[cur][ray].y)) : 0.0f;
                trav_begin<COUNT>(T, S, mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), 0.001f, 1e10f, lc, time);
                mode = 0;
                st = LANE_TRAVERSING;
            }
            stg_pos += take;
            want = __ballot_sync(0xffffffffu, st == LANE_NEED_PATH);
        }
        const unsigned live = __ballot_sync(0xffffffffu, st != LANE_DONE);
        if (live == 0u) break;
        const int min_active = (__popc(live) * P.burst_num + 31) >> 5;
        // vote_num <= 0 ("no voting") is the vote with threshold 1: descend whenever any lane wants to,
        // i.e. while-while order -- one copy of the traversal code in the kernel instead of two
        bool fin;
        trav_run_vote<COUNT, FULL>(T, st == LANE_TRAVERSING, fin, S, s_nodes, P.n_snodes, stack,
                             min_active > 0 ? min_active : 1, P.vote_num > 0 ? P.vote_num : 1, lc);
        if (fin) st = LANE_HAVE_RESULT;
    }
    cp_async_wait_all();
    if (COUNT) flush_counters(P.counters, lc);
}

// Equal-t hits of a wave, resolved with every lane busy (pt_device.cuh resolve_tie).  For each ray k_wf_extend listed:
// the primitive the reference's visiting order keeps; when that is not the one in the hit record the record is rewritten
// -- with the exit search if the hit turns INTO a medium boundary, without it if it stops being one.
template <bool FULL>
__global__ void __launch_bounds__(128) k_wf_fix_ties(const __grid_constant__ WfParams P) {
    const DScene& S = P.S;
    const int cur = P.bounce & 1;
    const int n = P.Q.tie_count[1];
    LocalCounters lc;
    const int stride = gridDim.x * blockDim.x;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const int ray = P.Q.tie_list[k];
        PT_BOUND(ray >= 0 && ray < P.Q.qcount[cur] && n <= P.Q.capacity, PT_DBG_QUEUE);
        const float4 a = P.Q.ra[cur][2 * (size_t)ray], b = P.Q.ra[cur][2 * (size_t)ray + 1];
        const float4 h = P.Q.hit[ray];
        const f3 o = mk3(a.x, a.y, a.z), d = mk3(a.w, b.x, b.y);
        const float time = FULL ? wf_path_time<FULL>(P, __float_as_int(P.Q.r2[cur][ray].y)) : 0.0f;
        f3 inv;
        ptm::inv_dir3(d.x, d.y, d.z, inv.x, inv.y, inv.z);
        const int old_ref = __float_as_int(h.y);
        const int ref = resolve_tie<FULL>(S, o, d, inv, 0.001f, h.x, time);
        if (ref == kRefNone || ref == old_ref) continue;
        const int cls = ref_class(ref);
        if (cls != PT_CLS_MEDIUM) {
            P.Q.hit[ray] = make_float4(h.x, __int_as_float(ref), 0.0f, __int_as_float(cls << 8));
        } else if (ref_class(old_ref) == PT_CLS_MEDIUM) {
            P.Q.hit[ray] = make_float4(h.x, __int_as_float(ref), h.z, h.w);       // same entry point: same exit search
        } else {
            float t_exit;
            int eref;
            traverse<false, FULL>(S, (SNodeBase)0, 0, o, d, h.x + 0.0001f, 1e10f, t_exit, eref, lc, time);
            P.Q.hit[ray] = make_float4(h.x, __int_as_float(ref), t_exit, __int_as_float((eref != kRefNone ? 1 : 0) | (PT_CLS_MEDIUM << 8)));
        }
    }
}

// Closest hit of a ray and, when that hit is a constant-medium boundary, the exit search behind it
// (kernels.py:417-419) -- as ONE traversal loop run up to twice, so that the kernels that inline it
// (k_wf_primary, k_wf_tail) carry one copy of the traversal code, not two (instruction cache).
// DEFER: an exact-t tie is not resolved here -- the function returns true at once and the caller hands the path to a
// launch that does (k_wf_primary<..., TIES>).
template <bool COUNT, bool FULL, bool REL = false, bool DEFER = false, bool OCT = false>
__device__ __forceinline__ bool closest_hit_and_exit(const DScene& S, SNodeBase s_nodes, int n_snodes, f3 o, f3 d, float time,
                                                     float& t, int& ref, float& t_exit, int& eref, LocalCounters& lc) {
    StackEntry stack_mem[kStackSlots];
    const LocalStack stack{stack_mem};
    stack_init(stack);
    TravState T;
    float t_min = 0.001f;
    t_exit = 0.0f;
    eref = kRefNone;
    trav_begin<COUNT>(T, S, o, d, t_min, 1e10f, lc, time);
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        if (pass) trav_restart<COUNT>(T, S, t_min, 1e10f, lc);
        if (OCT) T.cur = trav_root(S, T.inv);       // S.nodes holds the direction-octant variants
        trav_run<COUNT, FULL, REL, LocalStack, OCT>(T, S, s_nodes, n_snodes, stack, 0, lc);
        if (pass == 0) {
            if (DEFER) { if (trav_take_tie(T)) return true; }
            else trav_finish<FULL>(T, S);               // equal-t hits: the primitive the reference would have kept
            t = T.best_t;
            ref = T.best_ref;
            if (ref == kRefNone || !ref_is_medium(ref)) break;
            t_min = t + 0.0001f;
        } else {
            t_exit = T.best_t;
            eref = T.best_ref;
        }
    }
    return false;
}

// One ray of shade_miss_rays + shade_and_scatter (kernels.py:1266-1399): adds the ray's radiance
// contribution (miss -> background, hit -> emission) to `sp` and returns true with the next ray
// (origin, direction, throughput, depth) if the path goes on.
// A path adds radiance exactly once, at the event that ends it (a miss, or an emissive hit --
// emitters do not scatter, kernels.py:904-906), and its slab entry was zeroed when the path was
// generated; so "slab += add" is written as "slab = 0 + add": same bits, but a plain store instead
// of a scattered, latency-bound read-modify-write.
template <bool FULL>
__device__ __forceinline__ bool wf_shade_one(const DScene& S, f3 bg, int max_depth, float* sp, f3 o, f3 d, f3 thr,
                                             int depth, ptm::Rng& rng, float t, int ref, float t_exit_hit, bool exit_hit,
                                             f3& no, f3& nd, f3& nthr, int& ndepth, float time = 0.0f) {
    // features "nee" (FULL instantiations only): bits 29-30 of the depth word carry the path's next-event-estimation state
    // (nee_direct, pt_device.cuh) -- an emitter hit now may already have been counted at the previous vertex.  A path
    // then contributes at several vertices, so its slab entry is accumulated (+=) instead of written once.
    const bool nee = FULL && nee_on(S);
    const int prev_nee = nee ? ((depth >> 29) & 3) : 0;
    if (nee) depth &= 0x1fffffff;
    if (ref == kRefNone) {                           // shade_miss_rays, kernels.py:1266-1280
        f3 add = thr * bg;
        const f3 base = nee ? mk3(sp[0], sp[1], sp[2]) : mk3(0.0f, 0.0f, 0.0f);
        sp[0] = base.x + add.x; sp[1] = base.y + add.y; sp[2] = base.z + add.z;     // == slab + add, see above
        return false;
    }
    const int mid = mat_id_of(S, ref);               // shade_and_scatter, kernels.py:1289-1399
    Mat m = load_mat(S, mid);
    bool scattered = false, do_shade = true, emit_next = false, did_nee = false;
    f3 hp, sdir, atten, emit = mk3(0.0f, 0.0f, 0.0f);
    if (mat_is_medium(m.header)) {
        const float4 m2 = __ldg(S.materials + 3 * mid + 2);
        float t_exit, t_scatter;
        if (medium_sample(exit_hit, t_exit_hit, t, d, m2.x, rng, t_exit, t_scatter)) {
            hp = o + t_scatter * d;
            sdir = random_unit_vector(rng);
            atten = mk3(m2.y, m2.z, m2.w);
            scattered = true;
            do_shade = false;
        } else if (t_exit > 0.0f) {                  // passthrough: re-enqueue, same depth (:1342-1350)
            float eps_t = ptm::fdiv(0.001f, ptm::fsqrt(dot(d, d)));
            no = o + d * (t_exit + eps_t);
            nd = d;
            nthr = thr;
            ndepth = depth | (prev_nee != 0 ? (2 << 29) : 0);           // passed through a medium since the NEE vertex
            emit_next = true;
            do_shade = false;
        }
    }
    if (do_shade) {
        f3 n;
        finalize_hit<FULL>(S, ref, o, d, t, hp, n, time);
        if (nee && mat_type(m.header) == 0) {        // direct light at a Lambertian vertex, before its scatter draws
            const f3 add = thr * nee_direct<FULL>(S, hp, n, eval_texture(S, m, ref, hp), rng, time);
            sp[0] = sp[0] + add.x; sp[1] = sp[1] + add.y; sp[2] = sp[2] + add.z;
            did_nee = true;
        }
        scattered = shade_surface<FULL>(S, m, ref, d, hp, n, rng, emit, sdir, atten);
    }
    // kernels.py:1368-1375; after a vertex with next-event estimation only the close pairs are taken from here
    if ((emit.x > 0.0f || emit.y > 0.0f || emit.z > 0.0f) && (prev_nee == 0 || (prev_nee == 1 && (t * t) * dot(d, d) < S.nee_r2))) {
        f3 add = thr * emit;
        if (scattered || nee) {                      // no scattering emitter today: keep the general form
            sp[0] = sp[0] + add.x; sp[1] = sp[1] + add.y; sp[2] = sp[2] + add.z;
        } else {
            sp[0] = 0.0f + add.x; sp[1] = 0.0f + add.y; sp[2] = 0.0f + add.z;
        }
    }
    if (scattered) {
        f3 new_thr = thr * atten;
        int new_depth = depth + 1;
        if (new_depth < max_depth) {
            bool cont = true;
            if (new_depth >= 5) {
                float survival = min2(max2(max2(new_thr.x, new_thr.y), new_thr.z), 0.95f);
                if (ptm::rng_next(rng) > survival) cont = false;
                else new_thr = vdiv(new_thr, survival);
            }
            if (cont) {
                no = hp; nd = sdir; nthr = new_thr; ndepth = new_depth | (did_nee ? (1 << 29) : 0);
                emit_next = true;
            }
        }
    }
    return emit_next;
}

// Material binning: a streaming pass over the hit records that appends every ray index to the bin of
// its class (written by k_wf_extend into the record's flag word).  Each CTA handles a contiguous
// slice of 2048 rays: classes are counted per warp with ballots, reduced in shared memory, and the
// CTA reserves its space in every bin with ONE atomic per class.  k_wf_shade then walks the bins in
// class order, so its warps shade one material at a time.
constexpr int kBinRounds = 8;
__global__ void __launch_bounds__(256) k_wf_bin(const __grid_constant__ WfParams P) {
    __shared__ int s_cnt[kNumBins][kWfWarps];
    __shared__ int s_base[kNumBins];
    const int cur = P.bounce & 1;
    const int count = P.Q.qcount[cur];
    const unsigned lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int slice = 256 * kBinRounds;
    for (long long base = (long long)blockIdx.x * slice; base < count; base += (long long)gridDim.x * slice) {
        int cls[kBinRounds], rank[kBinRounds];
        int mine[kNumBins];
#pragma unroll
        for (int k = 0; k < kNumBins; k++) mine[k] = 0;
#pragma unroll
        for (int r = 0; r < kBinRounds; r++) {
            const long long i = base + r * 256 + threadIdx.x;
            cls[r] = i < count ? ((__float_as_int(P.Q.hit[i].w) >> 8) & (kNumBins - 1)) : -1;
            rank[r] = 0;
#pragma unroll
            for (int k = 0; k < kNumBins; k++) {
                const unsigned m = __ballot_sync(0xffffffffu, cls[r] == k);
                if (cls[r] == k) rank[r] = mine[k] + __popc(m & ((1u << lane) - 1));
                mine[k] += __popc(m);                  // warp-uniform running count of class k
            }
        }
        if (lane < kNumBins) {
            int v = 0;
#pragma unroll
            for (int k = 0; k < kNumBins; k++) v = lane == k ? mine[k] : v;
            s_cnt[lane][warp] = v;
        }
        __syncthreads();
        if (threadIdx.x < kNumBins) {
            int tot = 0;
            for (int w = 0; w < kWfWarps; w++) { int v = s_cnt[threadIdx.x][w]; s_cnt[threadIdx.x][w] = tot; tot += v; }
            s_base[threadIdx.x] = tot ? atomicAdd(P.Q.bin_count + threadIdx.x, tot) : 0;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kBinRounds; r++) {
            if (cls[r] >= 0) {
                const int pos = s_base[cls[r]] + s_cnt[cls[r]][warp] + rank[r];
                PT_BOUND(pos >= 0 && pos < P.Q.capacity, PT_DBG_BIN);
                P.Q.bins[(size_t)cls[r] * P.Q.capacity + pos] = (uint32_t)(base + r * 256 + threadIdx.x);
            }
        }
        __syncthreads();
    }
}

// shade_miss_rays + shade_and_scatter (kernels.py:1266-1399) over the material bins; survivors are
// appended to the other queue with one atomic per warp.
#ifndef PT_SHADE_THREADS
#define PT_SHADE_THREADS 256
#define PT_SHADE_MINB 3
#endif
// STAGE > 0: survivors are collected per warp in shared memory (room for STAGE rounds of 32) and appended in one go --
// one atomic on the queue counter per ~STAGE rounds instead of one per round, and the queue is written in whole
// 128-byte lines.  Why it matters: the append counter is ONE address, the L2 serialises atomics on it (measured: an
// extra returning atomic per warp and round on a neighbouring word DOUBLES the kernel's time, wave after wave --
// ~1.15 ns per atomic), and with one atomic per 32 rays that ceiling, not HBM, is what k_wf_shade ran into.
#ifndef PT_SHADE_STAGE
#define PT_SHADE_STAGE 2     // 2 / 3 / 4 rounds: k_wf_shade 3.80 / 3.78 / 3.80 ms per config-2 batch (4.98 unstaged) -- with half the atomics the
                             // counter is no longer the limit; 24 KB of shared memory per CTA instead of 48
#endif
// (The same staging for the survivors of the fused primary wave: k_wf_primary 4.07 vs 3.84 ms -- that kernel is issue-bound,
// two million atomics in 3.8 ms are below the counter's ceiling, and the staging code costs it issue slots.)
constexpr int kShadeStage = PT_SHADE_STAGE;
template <bool FULL, int STAGE = 0>
__global__ void __launch_bounds__(PT_SHADE_THREADS, PT_SHADE_MINB) k_wf_shade(const __grid_constant__ WfParams P) {
    extern __shared__ float4 s_stage[];       // STAGE: [warps][STAGE * 32][2] ray pairs, then [warps][STAGE * 32] third planes
    const DScene& S = P.S;
    const int cur = P.bounce & 1, nxt = cur ^ 1;
    const int count = P.Q.qcount[cur];
    if (blockIdx.x == 0 && threadIdx.x == 0) { P.Q.cursor[0] = 0; P.Q.tie_count[1] = 0; }       // for the next k_wf_extend
    if ((long long)blockIdx.x * blockDim.x >= count) return;
    // bin boundaries in shared memory: nine values that would otherwise sit in registers through the
    // whole shading code of this register-bound kernel
    __shared__ int prefix[kNumBins + 1];
    if (P.sort) {
        if (threadIdx.x == 0) {
            int acc = 0;
            prefix[0] = 0;
            for (int b = 0; b < kNumBins; b++) { acc += P.Q.bin_count[b]; prefix[b + 1] = acc; }
        }
        __syncthreads();
    }
    const int max_depth = P.cam.max_depth;
    const f3 bg = mk3(P.cam.background[0], P.cam.background[1], P.cam.background[2]);
    const int stride = gridDim.x * blockDim.x;
    const int rounds = (count + stride - 1) / stride;
    constexpr int kSlots = STAGE > 0 ? STAGE * 32 : 1;
    constexpr int kWarpsPerCta = PT_SHADE_THREADS / 32;
    const unsigned lane = threadIdx.x & 31;
    float4* const s_ab = s_stage + (size_t)(threadIdx.x >> 5) * kSlots * 2;
    float4* const s_c = s_stage + (size_t)kWarpsPerCta * kSlots * 2 + (size_t)(threadIdx.x >> 5) * kSlots;
    int fill = 0;                       // STAGE: rays in this warp's buffer (warp-uniform)
    auto flush = [&]() {                // STAGE: append the buffer to the next queue -- one atomic, coalesced line writes
        __syncwarp();
        int base = 0;
        if (lane == 0) base = atomicAdd(P.Q.qcount + nxt, fill);
        base = __shfl_sync(0xffffffffu, base, 0);
        PT_BOUND(base >= 0 && base + fill <= P.Q.capacity, PT_DBG_QUEUE);
        float4* ga = P.Q.ra[nxt] + 2 * (size_t)base;
        float4* gc = P.Q.r2[nxt] + (size_t)base;
        for (int i = lane; i < 2 * fill; i += 32) ga[i] = s_ab[i];
        for (int i = lane; i < fill; i += 32) gc[i] = s_c[i];
        __syncwarp();
        fill = 0;
    };
    for (int r = 0; r < rounds; r++) {
        const int j = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool active = j < count;
        float4 a, bq, c, h;
        if (active) {
            int i = j;
            if (P.sort) {
                int b = 0;
#pragma unroll
                for (int k = 1; k < kNumBins; k++) b += (j >= prefix[k]) ? 1 : 0;
                i = (int)P.Q.bins[(size_t)b * P.Q.capacity + (j - prefix[b])];
                PT_BOUND(j - prefix[b] >= 0 && j - prefix[b] < P.Q.capacity && i >= 0 && i < count, PT_DBG_BIN);
            }
            ld256(P.Q.ra[cur] + 2 * (size_t)i, a, bq);          // the 32-byte ray pair = one sector, one 256-bit load
            c = P.Q.r2[cur][i]; h = P.Q.hit[i];
        }
        bool emit_next = false;
        f3 no, nd, nthr;
        int path_id = 0, ndepth = 0;
        ptm::Rng rng{0u, 1u};
        if (active) {
            path_id = __float_as_int(c.y);
            PT_BOUND(path_id >= 0 && path_id < P.N * P.batch, PT_DBG_SLAB);
            const ptm::Rng key = wf_path_key(P, path_id);
            rng.s = __float_as_uint(c.z);
            rng.inc = key.inc;
            emit_next = wf_shade_one<FULL>(S, bg, max_depth, P.Q.slab + 3 * (size_t)path_id, mk3(a.x, a.y, a.z),
                                     mk3(a.w, bq.x, bq.y), mk3(bq.z, bq.w, c.x), __float_as_int(c.w), rng, h.x,
                                     __float_as_int(h.y), h.z, (__float_as_int(h.w) & 1) != 0, no, nd, nthr, ndepth,
                                     FULL ? ptm::path_time(key) : 0.0f);
        }
        // compaction: one atomic per warp (kernels.py:1345/:1394 use one per ray).  Aggregating
        // further, one atomic per CTA through shared memory and two barriers, measured slower
        // (shade 5.11 vs 4.88 ms per batch): the barriers couple the warps of a CTA, and this kernel
        // lives on its warps hiding each other's memory latency.
        if (STAGE > 0) {
            const unsigned m = __ballot_sync(0xffffffffu, emit_next);
            if (emit_next) {
                const int slot = fill + __popc(m & ((1u << lane) - 1));
                s_ab[2 * slot] = make_float4(no.x, no.y, no.z, nd.x);
                s_ab[2 * slot + 1] = make_float4(nd.y, nd.z, nthr.x, nthr.y);
                s_c[slot] = make_float4(nthr.z, __int_as_float(path_id), __uint_as_float(rng.s), __int_as_float(ndepth));
            }
            fill += __popc(m);
            if (fill > kSlots - 32) flush();
        } else {
            const int pos = warp_agg_append(P.Q.qcount + nxt, 0, emit_next);
            if (emit_next) {
                PT_BOUND(pos >= 0 && pos < P.Q.capacity, PT_DBG_QUEUE);
                st256(P.Q.ra[nxt] + 2 * (size_t)pos, make_float4(no.x, no.y, no.z, nd.x), make_float4(nd.y, nd.z, nthr.x, nthr.y));
                P.Q.r2[nxt][pos] = make_float4(nthr.z, __int_as_float(path_id), __uint_as_float(rng.s), __int_as_float(ndepth));
            }
        }
    }
    if (STAGE > 0 && fill > 0) flush();
}

// The same work split by shading class: the bins are class-sorted, so each of three kernels owns a
// contiguous run of the index space and is compiled for what that run needs --
//   KIND 0  miss + emissive: the path ends here; one multiply and one 12-byte store per ray, no next ray,
//           no random numbers (reads 32 of the ray's 48 bytes and, for emitters, the hit's primitive);
//   KIND 1  Lambert with a solid colour (the bulk of every scene): hit point, normal, cosine-weighted
//           direction, Russian roulette;
//   KIND 2  everything else through the general code (textures with Perlin noise / image lookups, metal,
//           dielectric, medium boundaries, full-mode subsurface).
// One kernel for all classes runs the trivial bins at the register budget of the worst one (80 registers,
// 3 CTAs per SM); split, the terminal kernel is a pure streaming pass at full occupancy.  Per-ray arithmetic
// is wf_shade_one's, operation for operation (bit-identical; options test "wf_shade_split").
template <bool FULL, int KIND>
__global__ void __launch_bounds__(256, KIND == 0 ? 8 : (KIND == 1 ? 4 : 3)) k_wf_shade_cls(const __grid_constant__ WfParams P) {
    constexpr int c_lo = KIND == 0 ? PT_CLS_MISS : (KIND == 1 ? PT_CLS_LAMBERT_SOLID : PT_CLS_LAMBERT_TEX);
    constexpr int c_hi = KIND == 0 ? PT_CLS_LAMBERT_SOLID : (KIND == 1 ? PT_CLS_LAMBERT_TEX : kNumBins);
    const DScene& S = P.S;
    const int cur = P.bounce & 1, nxt = cur ^ 1;
    if (KIND == 0 && blockIdx.x == 0 && threadIdx.x == 0) { P.Q.cursor[0] = 0; P.Q.tie_count[1] = 0; }       // for the next k_wf_extend
    __shared__ int prefix[c_hi - c_lo + 1];       // starts of this kernel's classes within its own run
    if (threadIdx.x == 0) {
        int acc = 0;
        prefix[0] = 0;
        for (int b = c_lo; b < c_hi; b++) { acc += P.Q.bin_count[b]; prefix[b - c_lo + 1] = acc; }
    }
    __syncthreads();
    const int count = prefix[c_hi - c_lo];
    if ((long long)blockIdx.x * blockDim.x >= count) return;
    const int max_depth = P.cam.max_depth;
    const f3 bg = mk3(P.cam.background[0], P.cam.background[1], P.cam.background[2]);
    const int stride = gridDim.x * blockDim.x;
    const int rounds = (count + stride - 1) / stride;
    for (int r = 0; r < rounds; r++) {
        const int j = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool active = j < count;
        bool emit_next = false;
        f3 no, nd, nthr;
        int path_id = 0, ndepth = 0;
        ptm::Rng rng{0u, 1u};
        if (active) {
            int b = 0;
#pragma unroll
            for (int k = 1; k < c_hi - c_lo; k++) b += (j >= prefix[k]) ? 1 : 0;
            const int i = (int)P.Q.bins[(size_t)(c_lo + b) * P.Q.capacity + (j - prefix[b])];
            PT_BOUND(i >= 0 && i < P.Q.qcount[cur] && P.Q.qcount[cur] <= P.Q.capacity, PT_DBG_BIN);
            const float4 bq = P.Q.ra[cur][2 * (size_t)i + 1], c = P.Q.r2[cur][i];
            const f3 thr = mk3(bq.z, bq.w, c.x);
            path_id = __float_as_int(c.y);
            float* sp = P.Q.slab + 3 * (size_t)path_id;
            if (KIND == 0) {
                f3 add;
                bool store = true;
                if (b == 0) {                                    // shade_miss_rays, kernels.py:1266-1280
                    add = thr * bg;
                } else {                                         // emitter: emitted() on both faces, no scatter (:904-906)
                    const int ref = __float_as_int(P.Q.hit[i].y);
                    const Mat m = load_mat(S, mat_id_of(S, ref));
                    const f3 emit = m.c1;
                    store = emit.x > 0.0f || emit.y > 0.0f || emit.z > 0.0f;      // kernels.py:1368
                    add = thr * emit;
                }
                if (store) { sp[0] = 0.0f + add.x; sp[1] = 0.0f + add.y; sp[2] = 0.0f + add.z; }   // == slab + add (wf_shade_one)
            } else {
                const float4 a = P.Q.ra[cur][2 * (size_t)i], h = P.Q.hit[i];
                const f3 o = mk3(a.x, a.y, a.z), d = mk3(a.w, bq.x, bq.y);
                const int depth = __float_as_int(c.w);
                const ptm::Rng key = wf_path_key(P, path_id);
                rng.s = __float_as_uint(c.z);
                rng.inc = key.inc;
                const float time = FULL ? ptm::path_time(key) : 0.0f;
                if (KIND == 1) {
                    const int ref = __float_as_int(h.y);
                    const Mat m = load_mat(S, mat_id_of(S, ref));
                    f3 hp, n;
                    finalize_hit<FULL>(S, ref, o, d, h.x, hp, n, time);
                    const f3 sdir = random_cosine_direction(n, rng);       // scatter(), Lambertian: kernels.py:833-838
                    f3 new_thr = thr * m.c1;                                // eval_texture, solid colour (:931)
                    const int new_depth = depth + 1;
                    if (new_depth < max_depth) {
                        bool cont = true;
                        if (new_depth >= 5) {                               // Russian roulette, :1379-1391
                            const float survival = min2(max2(max2(new_thr.x, new_thr.y), new_thr.z), 0.95f);
                            if (ptm::rng_next(rng) > survival) cont = false;
                            else new_thr = vdiv(new_thr, survival);
                        }
                        if (cont) { no = hp; nd = sdir; nthr = new_thr; ndepth = new_depth; emit_next = true; }
                    }
                } else {
                    emit_next = wf_shade_one<FULL>(S, bg, max_depth, sp, o, d, thr, depth, rng, h.x, __float_as_int(h.y), h.z,
                                                   (__float_as_int(h.w) & 1) != 0, no, nd, nthr, ndepth, time);
                }
            }
        }
        if (KIND != 0) {
            const int pos = warp_agg_append(P.Q.qcount + nxt, 0, emit_next);       // one atomic per warp
            if (emit_next) {
                PT_BOUND(pos >= 0 && pos < P.Q.capacity, PT_DBG_QUEUE);
                st256(P.Q.ra[nxt] + 2 * (size_t)pos, make_float4(no.x, no.y, no.z, nd.x), make_float4(nd.y, nd.z, nthr.x, nthr.y));
                P.Q.r2[nxt][pos] = make_float4(nthr.z, __int_as_float(path_id), __uint_as_float(rng.s), __int_as_float(ndepth));
            }
        }
    }
}

// Lane -> pixel map of k_wf_primary: the window is walked in 8x4-pixel tiles so that the 32 primary
// rays of a warp span a compact patch (and so do the bounce rays they append, which the next wave
// fetches 32 at a time); columns / rows that do not fill a tile follow in scanline order.  Any
// bijection is allowed here: a path is keyed by its (pixel, sample), never by the lane that runs it.
__device__ __forceinline__ void tile_local_pixel(int local, int W, int rows, int& px, int& py) {
    const int tw = W >> 3, th = rows >> 2;
    const int tiled = tw * th * 32;
    if (local < tiled) {
        const int t = local >> 5, in = local & 31;
        const int ty = t / tw, tx = t - ty * tw;
        px = tx * 8 + (in & 7);
        py = ty * 4 + (in >> 3);
    } else {
        int r = local - tiled;
        const int rw = W - tw * 8, right = rw * th * 4;
        if (r < right) {
            py = r / rw;
            px = tw * 8 + (r - py * rw);
        } else {
            r -= right;
            py = th * 4 + r / W;
            px = r - (r / W) * W;
        }
    }
}

// Wave 0 fused: generate_camera_rays + intersect_rays + shade_miss_rays + shade_and_scatter for the
// primary rays (kernels.py:1219-1399) in one kernel.  Camera rays are coherent, so nothing is gained
// by sending them through the queues: the ray is generated in registers, traced, shaded, and only
// the surviving bounce ray is appended (one atomic per warp) to queue 1, where wave 1 picks it up.
// (The OCT instantiations hold fewer values across the node visit and run 8 CTAs per SM: 62 registers, +0.4 % on config 2.)
// ONE: exactly one fused wave (the default schedule) -- throughput (1,1,1) and depth 0 are then compile-time
// constants, the wave loop disappears and with it the Russian-roulette code (depth + 1 < 5) and several
// registers that were live across the traversal.
// Equal-t hits (pt_device.cuh resolve_tie): the ONE instantiation does not resolve them -- it appends the path id to
// WfQueues.tie_list and goes on; a second, small launch of the TIES instantiation runs exactly those paths from their
// camera ray again, resolving in place.  A path is keyed by (pixel, sample), not by the launch or lane that runs it, and
// the order of the queue it lands in is free, so the frame is the same bits.  (The counting instantiation resolves in
// place: every traversal is counted once.)
// OCT: the launch traverses the direction-octant variants of the node array (pt_device.cuh trav_root; P.S.nodes and
// P.S.oct_stride are set for this launch only).
template <bool COUNT, bool FULL, bool REL, bool ONE, bool TIES = false, bool OCT = false>
__global__ void __launch_bounds__(PT_PRIMARY_THREADS, PT_PRIMARY_MINB + (OCT && PT_PRIMARY_MINB == 7 ? 1 : 0)) k_wf_primary(const __grid_constant__ WfParams P) {
    extern __shared__ float4 s_node_mem[];
    const DScene& S = P.S;
    stage_nodes(s_node_mem, S, P.n_snodes);
    const SNodeBase s_nodes = snode_base(s_node_mem);
    LocalCounters lc;
    const int max_depth = P.cam.max_depth;
    const f3 bg = mk3(P.cam.background[0], P.cam.background[1], P.cam.background[2]);
    constexpr bool DEFER = ONE && !COUNT && !TIES && PT_TIE;
    const int total = TIES ? P.Q.tie_count[0] : P.N * P.batch;
    const int stride = gridDim.x * blockDim.x;
    const int rounds = (total + stride - 1) / stride;
    const int qd = ONE ? 1 : (P.fused_waves & 1);   // the queue wave `fused_waves` reads from
    for (int r = 0; r < rounds; r++) {
        const int j = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        int id = j;
        bool emit_next = false;
        f3 no, nd, nthr;
        int ndepth = 0;
        ptm::Rng rng{0u, 1u};
        if (j < total) {
            const int jj = TIES ? P.Q.tie_list[j] : j;          // TIES: the listed path ids are (slot, row, column) already
            const int slot = jj / P.N;
            int px, ly;                                         // column, row within the window
            if (TIES) {
                const int l = jj - slot * P.N;
                ly = l / P.W;
                px = l - ly * P.W;
            } else if (P.tile) {
                tile_local_pixel(j - slot * P.N, P.W, P.y1 - P.y0, px, ly);
            } else {
                const int l = j - slot * P.N;
                ly = l / P.W;
                px = l - ly * P.W;
            }
            const int local = ly * P.W + px;
            id = slot * P.N + local;
            const int py = win_row(P.y0, P.row_shift, P.row_stride, ly);
            const int pixel = py * P.W + px;
            rng = ptm::rng_seed(P.seed, (uint32_t)pixel, (uint32_t)(P.sample_base + slot));
            const float time = FULL ? ptm::path_time(rng) : 0.0f;
            f3 o, d;
            get_ray(P.cam, px, py, rng, o, d, P.seed, (uint32_t)pixel, (uint32_t)(P.sample_base + slot));  // NOT normalised, kernels.py:1229-1233
            float* sp = P.Q.slab + 3 * (size_t)id;
            sp[0] = 0.0f; sp[1] = 0.0f; sp[2] = 0.0f;
            if (COUNT) lc.paths++;
            f3 thr = mk3(1.0f, 1.0f, 1.0f);
            int depth = 0;
            emit_next = max_depth > 0 && py < P.H;          // rows of an interleaved window beyond the image: no path
            const int fused_waves = ONE ? 1 : P.fused_waves;
#pragma unroll 1
            for (int w = 0; w < fused_waves && emit_next; w++) {
                float t, t_exit = 0.0f;
                int ref, eref = kRefNone;
                if (COUNT) lc.bounces++;
                if (closest_hit_and_exit<COUNT, FULL, REL, DEFER, OCT>(S, s_nodes, P.n_snodes, o, d, time, t, ref, t_exit, eref, lc)) {
                    const int k = atomicAdd(P.Q.tie_count, 1);      // DEFER only: this path runs again in the TIES launch
                    PT_BOUND(k >= 0 && k < P.Q.capacity, PT_DBG_QUEUE);
                    P.Q.tie_list[k] = id;
                    emit_next = false;
                    break;
                }
                emit_next = wf_shade_one<FULL>(S, bg, max_depth, sp, o, d, thr, depth, rng, t, ref, t_exit, eref != kRefNone,
                                         no, nd, nthr, ndepth, time);
                o = no; d = nd; thr = nthr; depth = ndepth;
            }
        }
        int pos = warp_agg_append(P.Q.qcount + qd, 0, emit_next);
        if (emit_next) {
            PT_BOUND(pos >= 0 && pos < P.Q.capacity && id >= 0 && id < P.N * P.batch, PT_DBG_QUEUE);
            const int id2 = id;
            st256(P.Q.ra[qd] + 2 * (size_t)pos, make_float4(no.x, no.y, no.z, nd.x), make_float4(nd.y, nd.z, nthr.x, nthr.y));
            P.Q.r2[qd][pos] = make_float4(nthr.z, __int_as_float(id2), __uint_as_float(rng.s), __int_as_float(ndepth));
        }
    }
    if (COUNT) flush_counters(P.counters, lc);
}

// Node boxes relative to one ray origin: (b - o) rounded once, exactly as the slab test would round
// it; the child references ride along unchanged.
__global__ void __launch_bounds__(256) k_make_rel_nodes(const float4* __restrict__ nodes, float4* __restrict__ rel, int num_inner,
                                                        float ox, float oy, float oz) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kNodeQuads * num_inner; i += stride) {
        float4 v = nodes[i];
#if PT_WIDE
        const int row = i & 7, axis = row >> 1;          // rows 0..5: (min, max) of x, y, z; rows 6, 7: references
        const float o = axis == 0 ? ox : (axis == 1 ? oy : oz);
        if (row < 6) v = make_float4(v.x - o, v.y - o, v.z - o, v.w - o);
#else
        const int row = i & 3;
        const float o = row == 0 ? ox : (row == 1 ? oy : oz);
        if (row < 3) v = make_float4(v.x - o, v.y - o, v.z - o, v.w - o);
#endif
        rel[i] = v;
    }
}

#if !PT_WIDE
// The eight direction-octant variants of the node array (trav_root): variant k swaps the (min, max) slots of both child
// boxes on the axes whose bit is set in k and points its inner references into itself (+ k * num_inner); with `rel`
// the origin is subtracted as in k_make_rel_nodes.  One thread per (variant, node).
__global__ void __launch_bounds__(256) k_make_oct_nodes(const float4* __restrict__ nodes, float4* __restrict__ out, int num_inner,
                                                        int rel, float ox, float oy, float oz) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 8 * num_inner; i += stride) {
        const int k = i / num_inner, j = i - k * num_inner;
        float4 v[3];
        const float o[3] = {ox, oy, oz};
        for (int a = 0; a < 3; a++) {
            v[a] = nodes[4 * j + a];
            if (rel) v[a] = make_float4(v[a].x - o[a], v[a].y - o[a], v[a].z - o[a], v[a].w - o[a]);
            if ((k >> a) & 1) v[a] = make_float4(v[a].y, v[a].x, v[a].w, v[a].z);
        }
        float4 r = nodes[4 * j + 3];
        int lref = __float_as_int(r.x), rref = __float_as_int(r.y);
        if (lref >= 0) lref += k * num_inner;
        if (rref >= 0) rref += k * num_inner;
        r.x = __int_as_float(lref);
        r.y = __int_as_float(rref);
        float4* q = out + 4 * (size_t)i;
        q[0] = v[0]; q[1] = v[1]; q[2] = v[2]; q[3] = r;
    }
}
#endif

// zeroes the per-batch counters before k_wf_primary
__global__ void k_wf_reset(const __grid_constant__ WfParams P) {
    if (threadIdx.x < kNumBins) P.Q.bin_count[threadIdx.x] = 0;
    if (threadIdx.x == 0) { P.Q.qcount[0] = 0; P.Q.qcount[1] = 0; P.Q.cursor[0] = 0; P.Q.cursor[1] = 0; P.Q.tie_count[0] = 0; P.Q.tie_count[1] = 0; }   // cursor[1]: k_wf_tail_vote's
}

// The tail of a batch: after `P.bounce` full waves only a small fraction of the paths is still
// alive, and running them as ever smaller waves (three launches each, latency-bound) costs more
// than the work itself.  This kernel finishes them megakernel-style: one lane per remaining ray,
// looping wave by wave (traverse, exit search for medium boundaries, wf_shade_one) until the path
// ends or the wave budget max_depth (renderer.py:313) is used up -- the same per-ray operations in
// the same order as the wave kernels, so results are identical.
template <bool COUNT, bool FULL>
__global__ void __launch_bounds__(256, 3) k_wf_tail(const __grid_constant__ WfParams P) {
    extern __shared__ float4 s_node_mem[];
    const DScene& S = P.S;
    const int cur = P.bounce & 1;
    const int count = P.Q.qcount[cur];
    if ((long long)blockIdx.x * blockDim.x >= count) return;
    stage_nodes(s_node_mem, S, P.n_snodes);
    const SNodeBase s_nodes = snode_base(s_node_mem);
    LocalCounters lc;
    const int max_depth = P.cam.max_depth;
    const f3 bg = mk3(P.cam.background[0], P.cam.background[1], P.cam.background[2]);
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const float4 a = P.Q.ra[cur][2 * (size_t)i], bq = P.Q.ra[cur][2 * (size_t)i + 1], c = P.Q.r2[cur][i];
        f3 o = mk3(a.x, a.y, a.z), d = mk3(a.w, bq.x, bq.y), thr = mk3(bq.z, bq.w, c.x);
        const int path_id = __float_as_int(c.y);
        const ptm::Rng key = wf_path_key(P, path_id);
        ptm::Rng rng{__float_as_uint(c.z), key.inc};
        int depth = __float_as_int(c.w);
        float* sp = P.Q.slab + 3 * (size_t)path_id;
        const float time = FULL ? ptm::path_time(key) : 0.0f;
        for (int wave = P.bounce; wave < max_depth; wave++) {
            float t, t_exit = 0.0f;
            int ref, eref = kRefNone;
            if (COUNT) lc.bounces++;
            closest_hit_and_exit<COUNT, FULL>(S, s_nodes, P.n_snodes, o, d, time, t, ref, t_exit, eref, lc);
            f3 no, nd, nthr;
            int ndepth;
            if (!wf_shade_one<FULL>(S, bg, max_depth, sp, o, d, thr, depth, rng, t, ref, t_exit, eref != kRefNone, no, nd, nthr, ndepth, time))
                break;
            o = no; d = nd; thr = nthr; depth = ndepth;
        }
    }
    if (COUNT) { __syncwarp(); flush_counters(P.counters, lc); }
}

// The same tail as a persistent state machine (the megakernel's scheme applied to the queue): warps
// pull the remaining rays from a cursor, lanes run traverse -> (exit search) -> wf_shade_one ->
// next bounce independently, and the traversal is the resumable warp-voting one, so a lane whose
// path has ended picks up the next ray while the others keep traversing.  Per-ray operations and
// their order are those of k_wf_tail: identical results.
template <bool COUNT, bool FULL>
__global__ void __launch_bounds__(256, 3) k_wf_tail_vote(const __grid_constant__ WfParams P) {
    const DScene& S = P.S;
    const int cur = P.bounce & 1;
    const int count = P.Q.qcount[cur];
    if ((long long)blockIdx.x * blockDim.x >= count) return;
    const SNodeBase s_nodes = 0;
    LocalCounters lc;
    const int max_depth = P.cam.max_depth;
    const f3 bg = mk3(P.cam.background[0], P.cam.background[1], P.cam.background[2]);
    const unsigned lane = threadIdx.x & 31;
    TravState T;
    T.cur = kRefNone;
    StackEntry stack_mem[kStackSlots];
    const LocalStack stack{stack_mem};
    stack_init(stack);
    f3 thr = mk3(1.0f, 1.0f, 1.0f);
    ptm::Rng rng{0u, 1u};
    int path_id = 0, depth = 0, wave = 0, mode = 0, ref_entry = kRefNone;
    float t_entry = 0.0f, time = 0.0f;
    int st = LANE_NEED_PATH;
    bool exhausted = false;                 // warp-uniform: the cursor ran past the end of the queue
    for (;;) {
        if (st == LANE_HAVE_RESULT) {
            bool shade = true;
            float t = T.best_t, t_exit = 0.0f;
            int ref = T.best_ref;
            bool exit_hit = false;
            if (mode == 0) {
                trav_finish<FULL>(T, S);                // equal-t hits: the primitive the reference would have kept
                ref = T.best_ref;
                if (ref != kRefNone && ref_is_medium(ref)) {          // exit search behind a medium boundary
                    mode = 1;
                    t_entry = t;
                    ref_entry = ref;
                    trav_restart<COUNT>(T, S, t_entry + 0.0001f, 1e10f, lc);
                    st = LANE_TRAVERSING;
                    shade = false;
                }
            } else {
                mode = 0;
                t_exit = T.best_t;
                exit_hit = T.best_ref != kRefNone;
                t = t_entry;
                ref = ref_entry;
            }
            if (shade) {
                f3 no, nd, nthr;
                int ndepth;
                const bool go = wf_shade_one<FULL>(S, bg, max_depth, P.Q.slab + 3 * (size_t)path_id, T.o, T.d, thr, depth, rng, t, ref,
                                                   t_exit, exit_hit, no, nd, nthr, ndepth, time);
                wave++;
                if (go && wave < max_depth) {
                    thr = nthr;
                    depth = ndepth;
                    if (COUNT) lc.bounces++;
                    trav_begin<COUNT>(T, S, no, nd, 0.001f, 1e10f, lc, time);
                    st = LANE_TRAVERSING;
                } else {
                    st = LANE_NEED_PATH;
                }
            }
        }
        // ---- idle lanes take the next rays of the queue (one atomic per warp and round)
        const unsigned want = __ballot_sync(0xffffffffu, st == LANE_NEED_PATH);
        if (want && !exhausted) {
            int base = 0;
            if (lane == 0) base = atomicAdd(P.Q.cursor + 1, __popc(want));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= count) exhausted = true;
            if (st == LANE_NEED_PATH) {
                const int i = base + __popc(want & ((1u << lane) - 1));
                if (i < count) {
                    PT_BOUND(i >= 0 && count <= P.Q.capacity, PT_DBG_QUEUE);
                    const float4 a = P.Q.ra[cur][2 * (size_t)i], bq = P.Q.ra[cur][2 * (size_t)i + 1], c = P.Q.r2[cur][i];
                    thr = mk3(bq.z, bq.w, c.x);
                    path_id = __float_as_int(c.y);
                    const ptm::Rng key = wf_path_key(P, path_id);
                    rng.s = __float_as_uint(c.z);
                    rng.inc = key.inc;
                    depth = __float_as_int(c.w);
                    wave = P.bounce;
                    time = FULL ? ptm::path_time(key) : 0.0f;
                    mode = 0;
                    if (wave < max_depth) {
                        if (COUNT) lc.bounces++;
                        trav_begin<COUNT>(T, S, mk3(a.x, a.y, a.z), mk3(a.w, bq.x, bq.y), 0.001f, 1e10f, lc, time);
                        st = LANE_TRAVERSING;
                    }
                } else {
                    st = LANE_DONE;
                }
            }
        } else if (want && st == LANE_NEED_PATH) {
            st = LANE_DONE;
        }
        const unsigned live = __ballot_sync(0xffffffffu, st != LANE_DONE);
        if (live == 0u) break;
        const int min_active = (__popc(__ballot_sync(0xffffffffu, st == LANE_TRAVERSING)) * P.burst_num + 31) >> 5;
        bool fin;
        trav_run_vote<COUNT, FULL>(T, st == LANE_TRAVERSING, fin, S, s_nodes, 0, stack, min_active > 0 ? min_active : 1,
                             P.vote_num > 0 ? P.vote_num : 1, lc);
        if (fin) st = LANE_HAVE_RESULT;
    }
    if (COUNT) flush_counters(P.counters, lc);
}

// stage-level probes: the queue a wave reads / the hit records a wave wrote, unpacked into plain arrays
__global__ void __launch_bounds__(256) k_wf_unpack_rays(const __grid_constant__ WfParams P, int cur, int n, float* o, float* d, float* thr,
                                                        int32_t* pixel, int32_t* depth) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = P.Q.ra[cur][2 * (size_t)i], b = P.Q.ra[cur][2 * (size_t)i + 1], c = P.Q.r2[cur][i];
    if (o) { o[3 * i] = a.x; o[3 * i + 1] = a.y; o[3 * i + 2] = a.z; }
    if (d) { d[3 * i] = a.w; d[3 * i + 1] = b.x; d[3 * i + 2] = b.y; }
    if (thr) { thr[3 * i] = b.z; thr[3 * i + 1] = b.w; thr[3 * i + 2] = c.x; }
    if (pixel) {
        const int path_id = __float_as_int(c.y);
        int slot = (int)((float)path_id * P.inv_N);
        int local = path_id - slot * P.N;
        if (local < 0) local += P.N; else if (local >= P.N) local -= P.N;
        pixel[i] = win_pixel(P, local);                 // ray_pixel_index (fields.py:205)
    }
    if (depth) depth[i] = __float_as_int(c.w) & 0x1fffffff;          // (bits 29-30: the next-event-estimation state of features "nee")
}
__global__ void __launch_bounds__(256) k_wf_unpack_hits(const __grid_constant__ WfParams P, int n, float* t, int32_t* prim, float* t_exit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 h = P.Q.hit[i];
    const int ref = __float_as_int(h.y);
    if (t) t[i] = h.x;
    if (prim) prim[i] = ref == kRefNone ? -1 : ((ref_type(ref) << 28) | ref_index(ref));
    if (t_exit) t_exit[i] = h.z;
}

// accum[pixel] += slab[0][pixel]; += slab[1][pixel]; ... in sample order (kernels.py:1280/:1373)
__global__ void __launch_bounds__(256) k_wf_resolve(const __grid_constant__ WfParams P) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.N; i += stride) {
        const int pixel = win_pixel(P, i);
        if (pixel >= P.W * P.H) continue;               // interleaved window: row beyond the image
        PT_BOUND(pixel >= 0, PT_DBG_PIXEL);
        float* ap = P.accum + 3 * (size_t)pixel;
        f3 acc = mk3(ap[0], ap[1], ap[2]);
        for (int s = 0; s < P.batch; s++) {
            const float* sp = P.Q.slab + 3 * ((size_t)s * P.N + i);
            acc = acc + mk3(sp[0], sp[1], sp[2]);
        }
        ap[0] = acc.x; ap[1] = acc.y; ap[2] = acc.z;
    }
}

// ------------------------------------------------------------------------------------------------
// probes, tonemap
// ------------------------------------------------------------------------------------------------
__global__ void k_generate_rays(PtCamera cam, int W, int H, uint32_t seed, int sample, float* o_out, float* d_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    int py = i / W, px = i - py * W;
    ptm::Rng rng = ptm::rng_seed(seed, (uint32_t)i, (uint32_t)sample);
    f3 o, d;
    get_ray(cam, px, py, rng, o, d, seed, (uint32_t)i, (uint32_t)sample);
    o_out[3 * i] = o.x; o_out[3 * i + 1] = o.y; o_out[3 * i + 2] = o.z;
    d_out[3 * i] = d.x; d_out[3 * i + 1] = d.y; d_out[3 * i + 2] = d.z;
}

struct FirstHitParams {
    DScene S;
    const float* o;
    const float* d;
    int n;
    float t_min, t_max;
    float* out_t;
    int32_t* out_prim;
    float* out_p;
    float* out_n;
    int n_snodes;
    DCounters* counters;
};

template <bool COUNT>
__global__ void __launch_bounds__(256) k_first_hit(const __grid_constant__ FirstHitParams P) {
    extern __shared__ float4 s_node_mem[];
    stage_nodes(s_node_mem, P.S, P.n_snodes);
    const SNodeBase s_nodes = snode_base(s_node_mem);
    LocalCounters lc;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        f3 o = mk3(P.o[3 * i], P.o[3 * i + 1], P.o[3 * i + 2]);
        f3 d = mk3(P.d[3 * i], P.d[3 * i + 1], P.d[3 * i + 2]);
        float t;
        int ref;
        traverse<COUNT>(P.S, s_nodes, P.n_snodes, o, d, P.t_min, P.t_max, t, ref, lc);
        P.out_t[i] = t;
        P.out_prim[i] = ref == kRefNone ? -1 : ((ref_type(ref) << 28) | ref_index(ref));
        if (P.out_p || P.out_n) {
            f3 p = mk3(0.0f, 0.0f, 0.0f), n = mk3(0.0f, 0.0f, 0.0f);
            if (ref != kRefNone) finalize_hit(P.S, ref, o, d, t, p, n);
            if (P.out_p) { P.out_p[3 * i] = p.x; P.out_p[3 * i + 1] = p.y; P.out_p[3 * i + 2] = p.z; }
            if (P.out_n) { P.out_n[3 * i] = n.x; P.out_n[3 * i + 1] = n.y; P.out_n[3 * i + 2] = n.z; }
        }
    }
    if (COUNT) {
        // lanes leave the loop at different times; re-converge before the shuffles
        __syncwarp();
        flush_counters(P.counters, lc);
    }
}

__global__ void k_math_eval(int fn, int n, const float* a, const float* b, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (fn == 6) {
        if (i == 0) {
            ptm::Rng s{__float_as_uint(a[0]), __float_as_uint(a[1])};      // (position, stride) of the stream
            for (int k = 0; k < n; k++) out[k] = ptm::rng_next(s);
        }
        return;
    }
    if (i >= n) return;
    float x = a[i], r = 0.0f;
    switch (fn) {
        case 0: r = ptm::sin_c(x); break;
        case 1: r = ptm::cos_c(x); break;
        case 2: r = ptm::log_c(x); break;
        case 3: r = ptm::acos_c(x); break;
        case 4: r = ptm::atan2_c(x, b[i]); break;
        case 5: r = ptm::pow5_c(x); break;
        case 7: r = ptm::exp_c(x); break;
        case 8: { float iy, iz; ptm::inv_dir3(x, b[i], 1.0f, r, iy, iz); r = (i & 1) ? iy : r; break; }   // traversal reciprocal
    }
    out[i] = r;
}

// Roofline probe: the node-fetch instruction sequence of the traversal kernels (four 16-byte read-only
// loads of one 64-byte record, load_node) issued at full rate with every lane active -- no stack, no
// arithmetic beyond a checksum, four independent records in flight per lane.  pattern 0: every lane
// its own pseudo-random record (what incoherent bounce rays do at the leaves of the tree); pattern 1:
// one record per warp (what coherent rays do at the top of the tree).  The measured GB/s over a
// working set that fits L1, L2 or neither are the memory roofs the traversal kernels sit under
// ("report against L2 bandwidth ... microbenchmark it on the box").
__global__ void __launch_bounds__(PT_EXT_WARPS * 32, 6) k_probe_gather(const float4* __restrict__ buf, unsigned n_records, int iters,
                                                                      int pattern, float* __restrict__ sink) {
    const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned who = pattern == 1 ? (gtid >> 5) : gtid;
    unsigned h = ptm::mix32(who * 0x9e3779b9u + 0x85ebca6bu);
    float acc = 0.0f;
    for (int i = 0; i < iters; i++) {
        unsigned idx[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            h = h * 0x915f77f5u + 0x6a09e667u;
            idx[k] = (unsigned)(((unsigned long long)(h ^ (h >> 15)) * n_records) >> 32);
        }
        float4 v[4][4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(buf) + (unsigned long long)idx[k] * 64ull);
            v[k][0] = __ldg(p); v[k][1] = __ldg(p + 1); v[k][2] = __ldg(p + 2); v[k][3] = __ldg(p + 3);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) acc += (v[k][0].x + v[k][1].y) + (v[k][2].z + v[k][3].w);
    }
    if (acc == 123.456f) sink[0] = acc;          // keeps the loads alive; never true for the zero/finite test buffer
}

// preview.py:117-132: u8 = clip(sqrt(max(0, accum * (1/spp))) * 255.999, 0, 255)
__global__ void k_tonemap(const float* accum, int n, float scale, uint8_t* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = accum[i] * scale;
    v = v > 0.0f ? v : 0.0f;
    float g = ptm::fsqrt(v) * 255.999f;
    g = g < 0.0f ? 0.0f : (g > 255.0f ? 255.0f : g);
    out[i] = (uint8_t)g;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static int pick_snodes(const PtContext* ctx, int blocks_per_sm_hint) {
#if !PT_SNODES
    (void)ctx; (void)blocks_per_sm_hint;
    return 0;                                   // no shared-memory node staging in this build (pt_device.cuh load_node)
#endif
    int n = ctx->smem_nodes;
    const int cap_bytes = ctx->max_smem_optin / (blocks_per_sm_hint > 0 ? blocks_per_sm_hint : 1) - 1024;
    int cap = cap_bytes > 0 ? cap_bytes / 64 : 0;
    if (n < 0) n = 256;                        // auto: the top 16 KB of the tree (more would eat L1)
    if (n > cap) n = cap;
    if (n > ctx->scene.num_inner) n = ctx->scene.num_inner;
    return n < 0 ? 0 : n;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_error(PT_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    return PT_OK;
}

static void wf_stage_free(PtContext* ctx);

extern "C" {

const char* pt_last_error(void) { return pt_error_slot()->c_str(); }
const char* pt_version(void) {
#if PT_FAST
    return PT_FAST_SLAB ? "b200pt 0.2 (sm_100a, PT-F32 fast: FMA slab test, FMA + approximate reciprocal)"
                        : "b200pt 0.2 (sm_100a, PT-F32 fast: FMA + approximate reciprocal around the exactly rounded intersection tests)";
#elif PT_DEBUG_BOUNDS
    return "b200pt 0.2 (sm_100a, PT-F32 strict, debug-bounds build)";
#elif PT_WIDE
    return "b200pt 0.2 (sm_100a, PT-F32 strict, 4-wide BVH)";
#else
    return "b200pt 0.2 (sm_100a, PT-F32 strict)";
#endif
}

int pt_create(int device_ordinal, PtContext** out_ctx) {
    if (!out_ctx) return set_error(PT_ERR_INVALID_ARG, "pt_create: out_ctx is NULL");
    int n = 0;
    PT_CUDA(cudaGetDeviceCount(&n));
    if (device_ordinal < 0 || device_ordinal >= n) return set_error(PT_ERR_INVALID_ARG, "pt_create: bad device ordinal");
    PT_CUDA(cudaSetDevice(device_ordinal));
    PtContext* c = new PtContext();
    c->device = device_ordinal;
    cudaDeviceProp prop;
    PT_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
    c->num_sms = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    PT_CUDA(cudaMalloc(&c->d_counters, sizeof(DCounters)));
    PT_CUDA(cudaMemset(c->d_counters, 0, sizeof(DCounters)));
    *out_ctx = c;
    return PT_OK;
}

int pt_destroy(PtContext* ctx) {
    if (!ctx) return PT_OK;
    cudaSetDevice(ctx->device);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->d_stage_snap) cudaFree(ctx->d_stage_snap);
    for (int i = 0; i < kMaxPipes; i++) {
        if (ctx->ev_resolve[i]) cudaEventDestroy(ctx->ev_resolve[i]);
        if (ctx->aux_stream[i]) cudaStreamDestroy(ctx->aux_stream[i]);
    }
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    wf_stage_free(ctx);
    delete ctx;
    return PT_OK;
}

int pt_set_option(PtContext* ctx, const char* name, int64_t value) {
    if (!ctx || !name) return set_error(PT_ERR_INVALID_ARG, "pt_set_option: NULL argument");
    std::string k(name);
    if (k == "block_threads") {
        if ((value != 64 && value != 128 && value != 256) || value > PT_MEGA_THREADS)
            return set_error(PT_ERR_INVALID_ARG, "block_threads must be 64, 128 or 256 and within the kernel's launch bound");
        ctx->block_threads = (int)value;
    } else if (k == "smem_nodes") {
        ctx->smem_nodes = (int)value;
    } else if (k == "burst_num") {
        if (value < 0 || value > 32) return set_error(PT_ERR_INVALID_ARG, "burst_num must be in [0, 32]");
        ctx->burst_num = (int)value;
    } else if (k == "vote_num") {
        if (value < 0 || value > 32) return set_error(PT_ERR_INVALID_ARG, "vote_num must be in [0, 32]");
        ctx->vote_num = (int)value;
    } else if (k == "wf_primary_threads") {
        if ((value != 64 && value != 128 && value != 256) || value > PT_PRIMARY_THREADS)
            return set_error(PT_ERR_INVALID_ARG, "wf_primary_threads must be 64, 128 or 256 and within the kernel's launch bound");
        ctx->wf_primary_threads = (int)value;
    } else if (k == "wf_overlap") {
        ctx->wf_overlap = value ? 1 : 0;
    } else if (k == "wf_pipelines") {
        if (value < 1 || value > kMaxPipes) return set_error(PT_ERR_INVALID_ARG, "wf_pipelines must be in [1, 4]");
        ctx->wf_pipelines = (int)value;
    } else if (k == "wf_tail_vote") {
        ctx->wf_tail_vote = value ? 1 : 0;
    } else if (k == "wf_rel_primary") {
        ctx->wf_rel_primary = value ? 1 : 0;
    } else if (k == "wf_extend_carveout") {
        if (value < -1 || value > 100) return set_error(PT_ERR_INVALID_ARG, "wf_extend_carveout must be -1 or a percentage");
        ctx->wf_extend_carveout = (int)value;
    } else if (k == "wf_extend_bulk") {
        ctx->wf_extend_bulk = value ? 1 : 0;
    } else if (k == "wf_oct_nodes") {
        ctx->wf_oct_nodes = value ? 1 : 0;
    } else if (k == "wf_grid_waves") {
        if (value < 1 || value > 4096) return set_error(PT_ERR_INVALID_ARG, "wf_grid_waves must be in [1, 4096]");
        ctx->wf_grid_waves = (int)value;
    } else if (k == "wf_shade_stage") {
        ctx->wf_shade_stage = value ? 1 : 0;
    } else if (k == "dbg_shade_smem") {
        if (value < 0 || value > 200 * 1024) return set_error(PT_ERR_INVALID_ARG, "dbg_shade_smem must be in [0, 200 KiB]");
        ctx->dbg_shade_smem = (int)value;
    } else if (k == "wf_vote_num") {
        if (value < 0 || value > 32) return set_error(PT_ERR_INVALID_ARG, "wf_vote_num must be in [0, 32]");
        ctx->wf_vote_num = (int)value;
    } else if (k == "wf_fuse_primary") {
        ctx->wf_fuse_primary = value < 0 ? 0 : (int)value;   // number of leading waves run fused (0 = off)
    } else if (k == "wf_waves") {
        if (value < 1) return set_error(PT_ERR_INVALID_ARG, "wf_waves must be >= 1");
        ctx->wf_waves = (int)value;
    } else if (k == "wf_burst_num") {
        if (value < 0 || value > 32) return set_error(PT_ERR_INVALID_ARG, "wf_burst_num must be in [0, 32]");
        ctx->wf_burst_num = (int)value;
    } else if (k == "wf_shade_split") {
        ctx->wf_shade_split = value ? 1 : 0;
    } else if (k == "wf_tile") {
        ctx->wf_tile = value ? 1 : 0;
    } else if (k == "wf_extend_minb") {
        if (value < 4 || value > 6) return set_error(PT_ERR_INVALID_ARG, "wf_extend_minb must be 4, 5 or 6");
        ctx->wf_extend_minb = (int)value;
    } else if (k == "wf_blocks_per_sm") {
        if (value < 1 || value > 8) return set_error(PT_ERR_INVALID_ARG, "wf_blocks_per_sm must be in [1, 8]");
        ctx->wf_blocks_per_sm = (int)value;
    } else if (k == "sync_samples") {
        ctx->sync_samples = value < 0 ? 0 : (int)value;
    } else if (k == "wavefront_sort") {
        ctx->wavefront_sort = value ? 1 : 0;
    } else {
        return set_error(PT_ERR_INVALID_ARG, "pt_set_option: unknown option " + k);
    }
    return PT_OK;
}

int pt_set_scene(PtContext* ctx, const void* device_blob, const PtPackedInfo* info) {
    if (!ctx || !info) return set_error(PT_ERR_INVALID_ARG, "pt_set_scene: NULL argument");
    if (!device_blob && info->total_bytes) return set_error(PT_ERR_INVALID_ARG, "pt_set_scene: NULL blob");
    // Validate everything before touching the context: a rejected scene leaves the previous one
    // (and has_scene) exactly as they were.
    if (info->max_depth_tree > kStackSize)
        return set_error(PT_ERR_CAPACITY, "pt_set_scene: BVH deeper than the 64-entry traversal stack (kernels.py:649)");
    if (info->node_bytes != kNodeBytes)
        return set_error(PT_ERR_INVALID_ARG, "pt_set_scene: the blob was packed for another node layout (node_bytes)");
    if (info->num_inner < 0 || info->num_spheres < 0 || info->num_quads < 0 || info->num_triangles < 0 ||
        info->num_materials < 0 || info->num_images < 0 || info->num_images > PT_MAX_IMAGES)
        return set_error(PT_ERR_INVALID_ARG, "pt_set_scene: bad counts in PtPackedInfo");
    const uint64_t total = info->total_bytes;
    auto inside = [&](uint64_t off, uint64_t bytes) { return off <= total && bytes <= total - off; };
    const uint64_t n_prims = (uint64_t)info->num_spheres + (uint64_t)info->num_triangles + (uint64_t)info->num_quads;
    bool ok = inside(info->off_nodes, (uint64_t)info->num_inner * kNodeBytes) && inside(info->off_spheres, (uint64_t)info->num_spheres * 16) &&
              inside(info->off_quads, (uint64_t)info->num_quads * 64) && inside(info->off_tris, (uint64_t)info->num_triangles * 48) &&
              inside(info->off_prim_mat, n_prims * 4) && inside(info->off_materials, (uint64_t)info->num_materials * 48) &&
              inside(info->off_perlin, 256 * 16 + 768) &&
              (!info->full_features || inside(info->off_sphere_motion, (uint64_t)info->num_spheres * 16)) &&
              info->num_lights >= 0 && inside(info->off_lights, (uint64_t)info->num_lights * 32);
    for (int i = 0; i < info->num_images && ok; i++)
        ok = info->image_w[i] > 0 && info->image_h[i] > 0 && inside(info->off_images, info->image_off[i]) &&
             inside(info->off_images + info->image_off[i], (uint64_t)info->image_w[i] * info->image_h[i] * 4);
    if (!ok) return set_error(PT_ERR_INVALID_ARG, "pt_set_scene: a section of PtPackedInfo lies outside the blob");
    const char* base = static_cast<const char*>(device_blob);
    DScene S;
    memset(&S, 0, sizeof S);
    S.nodes = reinterpret_cast<const float4*>(base + info->off_nodes);
    S.spheres = reinterpret_cast<const float4*>(base + info->off_spheres);
    S.quads = reinterpret_cast<const float4*>(base + info->off_quads);
    S.tris = reinterpret_cast<const float4*>(base + info->off_tris);
    S.prim_mat = reinterpret_cast<const uint32_t*>(base + info->off_prim_mat);
    S.nodes_plain = S.nodes;
    S.materials = reinterpret_cast<const float4*>(base + info->off_materials);
    S.perlin_vec = reinterpret_cast<const float4*>(base + info->off_perlin);
    S.perlin_perm = reinterpret_cast<const uint8_t*>(base + info->off_perlin + 256 * 16);
    for (int i = 0; i < info->num_images; i++) {
        S.image[i] = reinterpret_cast<const uchar4*>(base + info->off_images + info->image_off[i]);
        S.image_w[i] = info->image_w[i];
        S.image_h[i] = info->image_h[i];
    }
    S.num_inner = info->num_inner;
    S.n_prims[PT_PRIM_SPHERE] = info->num_spheres;
    S.n_prims[PT_PRIM_TRIANGLE] = info->num_triangles;
    S.n_prims[PT_PRIM_QUAD] = info->num_quads;
    S.num_materials = info->num_materials;
    S.mat_base[PT_PRIM_SPHERE] = 0;
    S.mat_base[PT_PRIM_TRIANGLE] = info->num_spheres;
    S.mat_base[PT_PRIM_QUAD] = info->num_spheres + info->num_triangles;
    S.has_media = info->has_media;
    S.full_features = info->full_features;
    S.sphere_motion = reinterpret_cast<const float4*>(base + info->off_sphere_motion);
    S.lights = reinterpret_cast<const float4*>(base + info->off_lights);
    S.n_lights = info->num_lights;
    S.nee_r2 = info->nee_r2;
    ctx->scene = S;
    ctx->info = *info;
    ctx->has_scene = true;
    return PT_OK;
}

int pt_set_camera(PtContext* ctx, const PtCamera* cam) {
    if (!ctx || !cam) return set_error(PT_ERR_INVALID_ARG, "pt_set_camera: NULL argument");
    if (cam->jitter != 0 && cam->jitter != 1) return set_error(PT_ERR_INVALID_ARG, "pt_set_camera: jitter must be 0 (random) or 1 (Sobol' (0,2))");
    ctx->cam = *cam;
    ctx->has_camera = true;
    return PT_OK;
}

static int check_render_args(PtContext* ctx, const float* accum, int W, int H, int y0, int y1, int sample_count) {
    if (!ctx || !accum) return set_error(PT_ERR_INVALID_ARG, "render: NULL argument");
    if (!ctx->has_scene || !ctx->has_camera) return set_error(PT_ERR_NO_SCENE, "render: scene or camera not set");
    if (W <= 0 || H <= 0 || y0 < 0 || y0 > y1 || sample_count < 0 || (long long)W * H > 0x7fffffffll)
        return set_error(PT_ERR_INVALID_ARG, "render: bad image window / sample range");
    if (ctx->row_shift < 0) {
        if (y1 > H) return set_error(PT_ERR_INVALID_ARG, "render: window rows beyond the image");
    } else {
        // interleaved window: y1 - y0 local rows, mapped by win_row; the last block may stick out of the image
        // (those rows are skipped), but not by a whole block
        const int rows = y1 - y0, blk = 1 << ctx->row_shift;
        if (rows > 0 && (rows % blk != 0 || y0 + (long long)(rows / blk - 1) * ctx->row_stride >= H))
            return set_error(PT_ERR_INVALID_ARG, "render: interleaved window needs whole blocks that start inside the image");
    }
    return PT_OK;
}

int pt_set_row_interleave(PtContext* ctx, int block_rows, int stride_rows) {
    if (!ctx) return set_error(PT_ERR_INVALID_ARG, "pt_set_row_interleave: NULL context");
    if (block_rows <= 0) { ctx->row_shift = -1; ctx->row_stride = 0; return PT_OK; }
    int sh = 0;
    while ((1 << sh) < block_rows) sh++;
    if ((1 << sh) != block_rows || block_rows < 4 || stride_rows < block_rows)
        return set_error(PT_ERR_INVALID_ARG, "pt_set_row_interleave: block_rows must be a power of two >= 4 and stride_rows >= block_rows");
    ctx->row_shift = sh;
    ctx->row_stride = stride_rows;
    return PT_OK;
}

int pt_render_megakernel(PtContext* ctx, float* accum, int W, int H, int y0, int y1, int sample_begin, int sample_count,
                         uint32_t seed, void* stream) {
    int rc = check_render_args(ctx, accum, W, H, y0, y1, sample_count);
    if (rc) return rc;
    if (y0 == y1 || sample_count == 0) return PT_OK;
    PT_CUDA(cudaSetDevice(ctx->device));
    MegaParams P;
    P.S = ctx->scene;
    P.cam = ctx->cam;
    P.accum = accum;
    P.W = W; P.H = H; P.y0 = y0; P.y1 = y1;
    P.row_shift = ctx->row_shift; P.row_stride = ctx->row_stride;
    P.sample_begin = sample_begin; P.sample_count = sample_count;
    P.seed = seed;
    P.tiles_x = (W + 7) / 8;
    P.burst_num = ctx->burst_num;
    P.sync_samples = ctx->sync_samples;
    P.vote_num = ctx->vote_num;
    P.counters = ctx->d_counters;
    const int tiles_y = (y1 - y0 + 3) / 4;
    const long long warps = (long long)P.tiles_x * tiles_y;
    const int bt = ctx->block_threads;
    const int wpb = bt / 32;
    const int blocks = (int)((warps + wpb - 1) / wpb);
    P.n_snodes = pick_snodes(ctx, 2);
    const size_t smem = (size_t)P.n_snodes * 64;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool full = ctx->scene.full_features != 0;
    auto go = [&](auto kernel) -> int {
        int r = set_smem(kernel, smem);
        if (r) return r;
        kernel<<<blocks, bt, smem, st>>>(P);
        return PT_OK;
    };
    if (ctx->count) rc = full ? go(k_megakernel<true, true>) : go(k_megakernel<true, false>);
    else rc = full ? go(k_megakernel<false, true>) : go(k_megakernel<false, false>);
    if (rc) return rc;
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Stage timing: brackets a launch with two events on the launching stream (only when enabled; the
// timed product path records nothing).
static int stage_event(PtContext* ctx, cudaStream_t st) {
    if (ctx->ev_used == (int)ctx->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return -1;
        ctx->ev_pool.push_back(e);
    }
    const int id = ctx->ev_used++;
    cudaEventRecord(ctx->ev_pool[id], st);
    return id;
}
struct StageScope {
    PtContext* ctx; cudaStream_t st; int kernel, wave, b;
    StageScope(PtContext* c, cudaStream_t s, int k, int w) : ctx(c), st(s), kernel(k), wave(w), b(-1) {
        if (ctx->stage_timing) b = stage_event(ctx, st);
    }
    ~StageScope() {
        if (ctx->stage_timing && b >= 0) {
            const int e = stage_event(ctx, st);
            if (e < 0) return;
            // marks and counter snapshots stay in step: once the snapshot buffer is full nothing more is recorded
            if ((int)ctx->marks.size() >= kMaxStageSnaps) return;
            if (ctx->count && ctx->d_stage_snap)
                cudaMemcpyAsync(ctx->d_stage_snap + ctx->marks.size(), ctx->d_counters, sizeof(DCounters),
                                cudaMemcpyDeviceToDevice, st);
            ctx->marks.push_back({kernel, wave, b, e});
        }
    }
};

// bytes of workspace one in-flight sample of an N-pixel window needs
static size_t wf_bytes_per_sample(size_t n) {
    return 6 * n * 16      // two queues x three float4 planes
           + n * 16        // hit records
           + (size_t)kNumBins * n * 4   // material bins
           + n * 12        // radiance slab
           + n * 4         // tie list (k_wf_fix_ties, k_wf_primary<..., TIES>)
           + n * 8;        // spare (round 1 kept a second slab here; the size formula is part of the ABI)
}
static const size_t kMaxPathsInFlight = 0x78000000;
static const size_t kWfFixedBytes = 4096 + 32 * 256;   // counter sets + per-section alignment slack

size_t pt_wavefront_workspace_bytes(int W, int rows) { return pt_wavefront_workspace_bytes_for(W, rows, 1); }

size_t pt_wavefront_workspace_bytes_for(int W, int rows, int samples_in_flight) {
    if (W <= 0 || rows <= 0 || samples_in_flight <= 0) return 0;
    return wf_bytes_per_sample((size_t)W * rows) * (size_t)samples_in_flight + kWfFixedBytes;
}

int pt_render_wavefront(PtContext* ctx, float* accum, int W, int H, int y0, int y1, int sample_begin, int sample_count,
                        uint32_t seed, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_render_args(ctx, accum, W, H, y0, y1, sample_count);
    if (rc) return rc;
    if (y0 == y1 || sample_count == 0) return PT_OK;
    const int rows = y1 - y0;
    const size_t n = (size_t)W * rows;
    if (!workspace || workspace_bytes < pt_wavefront_workspace_bytes(W, rows))
        return set_error(PT_ERR_WORKSPACE, "pt_render_wavefront: workspace too small");
    // The origin-relative node copy of the primary wave (64 B per inner node) comes out of the caller's
    // workspace too, when that leaves room for at least one sample in flight; it is rebuilt by every
    // call (microseconds), so the context owns no device memory besides its counters.
    size_t rel_bytes = align_up((size_t)ctx->scene.num_inner * kNodeBytes, 256);
    // With room for eight copies (and up to 1 GiB of them), the primary wave traverses the direction-octant variants of
    // that array (pt_device.cuh trav_root).
    bool oct = false;
#if !PT_WIDE       // (fast numerics: fma(b, inv, noi) is monotone in b like (b - o) * inv, so the variants hold there too)
    if (ctx->wf_oct_nodes && ctx->scene.num_inner > 0 && 8 * rel_bytes <= ((size_t)1 << 30) &&
        workspace_bytes >= kWfFixedBytes + 8 * rel_bytes + wf_bytes_per_sample(n)) {
        oct = true;
        rel_bytes *= 8;
    }
#endif
    if (workspace_bytes < kWfFixedBytes + rel_bytes + wf_bytes_per_sample(n)) rel_bytes = 0;
    // samples in flight: as many as the workspace holds (and the 31-bit path id allows)
    const size_t S_cap = (workspace_bytes - kWfFixedBytes - rel_bytes) / wf_bytes_per_sample(n);
    // Two pipelines: the workspace is split into two complete sets of queues and consecutive batches run
    // on two streams (the caller's and a context-owned one).  Every stage of a batch depends only on the
    // stage before it, so the two batches are independent until their resolves, which are chained in
    // batch order.  What this buys: the small launches of a batch -- late waves that cannot fill 148 SMs,
    // the tail's latency chain, bin / resolve passes -- run in the shadow of the other batch's big ones,
    // and the block scheduler back-fills every kernel's ramp-down with the other stream's CTAs.
    int n_pipe = (ctx->wf_overlap && !ctx->stage_timing) ? ctx->wf_pipelines : 1;
    if ((size_t)n_pipe > S_cap) n_pipe = (int)S_cap;
    if (n_pipe > sample_count) n_pipe = sample_count;
    size_t B = S_cap / (size_t)n_pipe;
    if (B > ((size_t)sample_count + n_pipe - 1) / n_pipe) B = ((size_t)sample_count + n_pipe - 1) / n_pipe;
    if (B > 256) B = 256;
    // paths in flight are indexed with 32-bit ints; the grid-stride loops overshoot `total` by up to one
    // grid (< 2^26 threads) before their bounds test, so the cap leaves that much headroom
    while (B > 1 && B * n > kMaxPathsInFlight) B--;
    if (B == 0 || B * n > kMaxPathsInFlight) return set_error(PT_ERR_CAPACITY, "pt_render_wavefront: window too large");
    // equal batches: 64 samples with room for 63 in flight run as 32 + 32, not 63 + 1
    {
        const size_t nb = ((size_t)sample_count + B - 1) / B;
        B = ((size_t)sample_count + nb - 1) / nb;
    }
    const int n_batches = (sample_count + (int)B - 1) / (int)B;
    if (n_batches < n_pipe) n_pipe = n_batches;
    PT_CUDA(cudaSetDevice(ctx->device));
    const size_t cap = B * n;
    char* p = static_cast<char*>(workspace);
    float4* rel_nodes = rel_bytes ? reinterpret_cast<float4*>(p) : nullptr;
    p += rel_bytes;
    WfParams P;
    P.S = ctx->scene;
    P.cam = ctx->cam;
    auto take = [&](size_t bytes) { char* q = p; p += align_up(bytes, 256); return q; };
    WfQueues sets[kMaxPipes];
    for (int k = 0; k < n_pipe; k++) {
        WfQueues& Q = sets[k];
        for (int q = 0; q < 2; q++) {
            Q.ra[q] = reinterpret_cast<float4*>(take(cap * 32));
            Q.r2[q] = reinterpret_cast<float4*>(take(cap * 16));
        }
        Q.hit = reinterpret_cast<float4*>(take(cap * 16));
        Q.bins = reinterpret_cast<uint32_t*>(take((size_t)kNumBins * cap * 4));
        Q.slab = reinterpret_cast<float*>(take(cap * 12));
        Q.tie_list = reinterpret_cast<int*>(take(cap * 4));
        int* counters = reinterpret_cast<int*>(take(256));       // 64 ints
        Q.qcount = counters;
        Q.bin_count = counters + 8;
        Q.cursor = counters + 32;
        Q.tie_count = counters + 40;
        Q.capacity = (int)cap;
    }
    P.Q = sets[0];
    P.accum = accum;
    P.W = W; P.H = H; P.y0 = y0; P.y1 = y1;
    P.row_shift = ctx->row_shift; P.row_stride = ctx->row_stride;
    P.inv_W = 1.0f / (float)W;
    P.N = (int)n;
    P.inv_N = 1.0f / (float)n;
    P.seed = seed;
    P.sort = ctx->wavefront_sort;
    P.burst_num = ctx->wf_burst_num;
    P.vote_num = ctx->wf_vote_num;
    P.tile = ctx->wf_tile;
    P.counters = ctx->d_counters;
    P.n_snodes = pick_snodes(ctx, 4);
    const size_t smem = (size_t)P.n_snodes * 64;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int bt = 256;
    const int grid_persist = ctx->num_sms * ctx->wf_blocks_per_sm;
    const bool full = ctx->scene.full_features != 0;
    // kernel instantiations for this context: counting on/off x feature mode
    void (*k_primary)(const WfParams) = nullptr;
    void (*k_extend)(const WfParams) = nullptr;
    void (*k_tail)(const WfParams) = nullptr;
    // exact-t ties: the launches that resolve what k_wf_extend / the fused primary wave only listed (not in counting mode)
    void (*k_fix_ties)(const WfParams) = full ? k_wf_fix_ties<true> : k_wf_fix_ties<false>;
    void (*k_primary_ties)(const WfParams) = nullptr;
    const int grid_fix = ctx->num_sms * 2;
    void (*k_shade)(const WfParams) = full ? k_wf_shade<true> : k_wf_shade<false>;
    size_t shade_smem = (size_t)ctx->dbg_shade_smem;
    if (ctx->wf_shade_stage) {
        k_shade = full ? k_wf_shade<true, kShadeStage> : k_wf_shade<false, kShadeStage>;
        const size_t need = (size_t)(PT_SHADE_THREADS / 32) * kShadeStage * 32 * 48;
        if (shade_smem < need) shade_smem = need;
    }
    const bool split = ctx->wf_shade_split && ctx->wavefront_sort && !ctx->dbg_shade_smem && !(ctx->scene.full_features & 2);
    void (*k_shade_cls[3])(const WfParams) = {full ? k_wf_shade_cls<true, 0> : k_wf_shade_cls<false, 0>,
                                              full ? k_wf_shade_cls<true, 1> : k_wf_shade_cls<false, 1>,
                                              full ? k_wf_shade_cls<true, 2> : k_wf_shade_cls<false, 2>};
    int occ_cls[3] = {1, 1, 1};
    // camera rays without defocus all start at the camera centre: the fused primary wave then tests
    // origin-relative node boxes (one subtraction less per box coordinate).  Only when that kernel
    // runs wave 0 alone -- later waves start elsewhere.
    // (not with next-event estimation: its shadow rays start at the hit points and need the plain node boxes)
    bool rel = !PT_FAST_SLAB && !(ctx->scene.full_features & 2) && ctx->wf_rel_primary && ctx->wf_fuse_primary == 1 && !(ctx->cam.defocus_angle > 0.0f) &&
               ctx->scene.num_inner > 0 && ctx->cam.max_depth > 0 && !ctx->count && rel_nodes != nullptr;
    // ... and its direction-octant variants (8 copies, with the origin subtracted when `rel`), same conditions minus the
    // shared origin: camera rays are coherent in direction with or without defocus
    oct = oct && !(ctx->scene.full_features & 2) && ctx->wf_fuse_primary == 1 && ctx->scene.num_inner > 0 && ctx->cam.max_depth > 0 &&
          !ctx->count && rel_nodes != nullptr;
#if !PT_WIDE
    if (oct) {
        int g = (int)(((size_t)ctx->scene.num_inner * 8 + 255) / 256);
        if (g > ctx->num_sms * 8) g = ctx->num_sms * 8;
        k_make_oct_nodes<<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(ctx->scene.nodes, rel_nodes, ctx->scene.num_inner, rel ? 1 : 0,
                                                                          ctx->cam.center[0], ctx->cam.center[1], ctx->cam.center[2]);
        PT_CUDA(cudaGetLastError());
    } else
#endif
    if (rel) {
        int g = (int)(((size_t)ctx->scene.num_inner * kNodeQuads + 255) / 256);
        if (g > ctx->num_sms * 8) g = ctx->num_sms * 8;
        k_make_rel_nodes<<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(ctx->scene.nodes, rel_nodes, ctx->scene.num_inner,
                                                                          ctx->cam.center[0], ctx->cam.center[1], ctx->cam.center[2]);
        PT_CUDA(cudaGetLastError());
    }
    if (ctx->count) {
        k_primary = full ? k_wf_primary<true, true, false, false> : k_wf_primary<true, false, false, false>;
        k_extend = full ? k_wf_extend<true, true, 4> : k_wf_extend<true, false, 4>;
        if (ctx->wf_tail_vote) k_tail = full ? k_wf_tail_vote<true, true> : k_wf_tail_vote<true, false>;
        else k_tail = full ? k_wf_tail<true, true> : k_wf_tail<true, false>;
    } else {
        const bool one = ctx->wf_fuse_primary == 1 && ctx->cam.max_depth > 0;   // then fused == 1 below
        if (one && PT_TIE) k_primary_ties = full ? k_wf_primary<false, true, false, true, true> : k_wf_primary<false, false, false, true, true>;
        if (oct && rel) k_primary = full ? k_wf_primary<false, true, true, true, false, true> : k_wf_primary<false, false, true, true, false, true>;
        else if (oct) k_primary = full ? k_wf_primary<false, true, false, true, false, true> : k_wf_primary<false, false, false, true, false, true>;
        else if (rel) k_primary = full ? k_wf_primary<false, true, true, true> : k_wf_primary<false, false, true, true>;
        else if (one) k_primary = full ? k_wf_primary<false, true, false, true> : k_wf_primary<false, false, false, true>;
        else k_primary = full ? k_wf_primary<false, true, false, false> : k_wf_primary<false, false, false, false>;
        // MINB = resident CTAs per SM the register allocation is capped for (4: 64 regs, 5: 48, 6: 40)
        if (ctx->wf_extend_bulk) k_extend = full ? k_wf_extend<false, true, 6, true> : k_wf_extend<false, false, 6, true>;
        else if (ctx->wf_extend_minb >= 6) k_extend = full ? k_wf_extend<false, true, 6> : k_wf_extend<false, false, 6>;
        else if (ctx->wf_extend_minb == 5) k_extend = full ? k_wf_extend<false, true, 5> : k_wf_extend<false, false, 5>;
        else k_extend = full ? k_wf_extend<false, true, 4> : k_wf_extend<false, false, 4>;
        if (ctx->wf_tail_vote) k_tail = full ? k_wf_tail_vote<false, true> : k_wf_tail_vote<false, false>;
        else k_tail = full ? k_wf_tail<false, true> : k_wf_tail<false, false>;
    }
    if ((rc = set_smem(k_primary, smem)) || (rc = set_smem(k_extend, smem)) || (rc = set_smem(k_tail, smem))) return rc;
    PT_CUDA(cudaFuncSetAttribute(k_extend, cudaFuncAttributePreferredSharedMemoryCarveout, ctx->wf_extend_carveout));
    if (k_primary_ties && (rc = set_smem(k_primary_ties, smem))) return rc;
    if (shade_smem && (rc = set_smem(k_shade, shade_smem))) return rc;
    int occ_primary = 1, occ_shade = 1, occ_tail = 1;
    const int bt_primary = ctx->wf_primary_threads;
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_primary, k_primary, bt_primary, smem));
    const int bt_shade = PT_SHADE_THREADS;
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_shade, k_shade, bt_shade, shade_smem));
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_tail, k_tail, bt, smem));
    for (int k = 0; k < 3; k++) {
        PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cls[k], k_shade_cls[k], 256, 0));
        if (occ_cls[k] < 1) occ_cls[k] = 1;
        if (ctx->wf_grid_waves > 1) occ_cls[k] *= ctx->wf_grid_waves;
    }
    if (occ_primary < 1) occ_primary = 1;
    if (occ_shade < 1) occ_shade = 1;
    if (occ_tail < 1) occ_tail = 1;
    if (ctx->wf_grid_waves > 1) { occ_primary *= ctx->wf_grid_waves; occ_shade *= ctx->wf_grid_waves; occ_tail *= ctx->wf_grid_waves; }
    const int max_depth = ctx->cam.max_depth;
    ctx->marks.clear();
    ctx->ev_used = 0;
    const int waves = max_depth < ctx->wf_waves ? max_depth : ctx->wf_waves;
    const int fused = (ctx->wf_fuse_primary && max_depth > 0) ? (ctx->wf_fuse_primary < waves ? ctx->wf_fuse_primary : waves) : 0;
    const int tail_bounce = waves > fused ? waves : fused;
    cudaStream_t streams[kMaxPipes] = {st, st, st, st};
    if (n_pipe > 1) {
        bool ok = ctx->ev_start || cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < n_pipe && ok; i++) {
            if (i > 0 && !ctx->aux_stream[i]) ok = cudaStreamCreateWithFlags(&ctx->aux_stream[i], cudaStreamNonBlocking) == cudaSuccess;
            if (ok && !ctx->ev_resolve[i]) ok = cudaEventCreateWithFlags(&ctx->ev_resolve[i], cudaEventDisableTiming) == cudaSuccess;
        }
        if (!ok) {
            (void)cudaGetLastError();
            return set_error(PT_ERR_CUDA, "pt_render_wavefront: cannot create the pipeline streams / events");
        }
        // the other streams start after everything the caller queued before this call (and the node copy)
        PT_CUDA(cudaEventRecord(ctx->ev_start, st));
        for (int i = 1; i < n_pipe; i++) {
            streams[i] = ctx->aux_stream[i];
            PT_CUDA(cudaStreamWaitEvent(streams[i], ctx->ev_start, 0));
        }
    }
    int batch_index = 0;
    for (int s = sample_begin; s < sample_begin + sample_count; s += (int)B, batch_index++) {
        const int pipe = batch_index % n_pipe;
        cudaStream_t bs = streams[pipe];
        P.Q = sets[pipe];
        P.sample_base = s;
        P.batch = (int)B < sample_begin + sample_count - s ? (int)B : sample_begin + sample_count - s;
        const size_t total = n * (size_t)P.batch;
        int grid_full = (int)((total + bt - 1) / bt);
        const int grid_full_extend = (int)((total + PT_EXT_WARPS * 32 - 1) / (PT_EXT_WARPS * 32));
        const int grid_extend = grid_persist < grid_full_extend ? grid_persist : grid_full_extend;
        int grid_wide = ctx->num_sms * 8 < grid_full ? ctx->num_sms * 8 : grid_full;
        // Kernels that split their work statically over the grid (primary, shade, tail) launch a
        // whole number of waves of RESIDENT CTAs (occupancy x SMs).  Measured on config 2
        // (tools/stage_times.py): 1 wave 6.07 / 6.44 ms (primary / shade per batch: per-CTA cost
        // varies with the image region, the slowest CTA sets the time), 8/3 waves -- the former
        // "8 CTAs per SM" with 3 resident -- 5.67 / 4.9, 16 waves 5.43 / 4.12, 1024 waves 5.91 / 6.02
        // (per-CTA set-up dominates).
        const int grid_full_primary = (int)((total + bt_primary - 1) / bt_primary);
        const int grid_primary = ctx->num_sms * occ_primary < grid_full_primary ? ctx->num_sms * occ_primary : grid_full_primary;
        const int grid_full_shade = (int)((total + bt_shade - 1) / bt_shade);
        const int grid_shade = ctx->num_sms * occ_shade < grid_full_shade ? ctx->num_sms * occ_shade : grid_full_shade;
        const int grid_tail = ctx->num_sms * occ_tail < grid_full ? ctx->num_sms * occ_tail : grid_full;
        int first_wave = 0;
        P.bounce = 0;
        if (fused > 0) {
            // the first wave(s) in one kernel; survivors land in the queue the next wave reads
            P.fused_waves = fused;
            { StageScope t(ctx, bs, PT_STAGE_RESET, 0); k_wf_reset<<<1, 32, 0, bs>>>(P); }
            {
                StageScope t(ctx, bs, PT_STAGE_PRIMARY, 0);
                WfParams Pp = P;
                if (rel || oct) Pp.S.nodes = rel_nodes;                    // this launch only
                if (oct) Pp.S.oct_stride = ctx->scene.num_inner;
                k_primary<<<grid_primary, bt_primary, smem, bs>>>(Pp);
                if (k_primary_ties) k_primary_ties<<<ctx->num_sms, bt_primary, smem, bs>>>(P);      // plain node boxes
            }
            first_wave = P.fused_waves;
        } else {
            { StageScope t(ctx, bs, PT_STAGE_GENERATE, 0); k_wf_generate<<<grid_wide, bt, 0, bs>>>(P); }
        }
        for (int b = first_wave; b < waves; b++) {        // renderer.py:313; no host read-back of the ray count
            P.bounce = b;
            {
                StageScope t(ctx, bs, PT_STAGE_EXTEND, b);
                k_extend<<<grid_extend, PT_EXT_WARPS * 32, smem, bs>>>(P);
                if (!ctx->count && PT_TIE) k_fix_ties<<<grid_fix, 128, 0, bs>>>(P);
            }
            if (P.sort) { StageScope t(ctx, bs, PT_STAGE_BIN, b); k_wf_bin<<<grid_wide, bt, 0, bs>>>(P); }
            if (split) {
                for (int k = 0; k < 3; k++) {
                    const int g = ctx->num_sms * occ_cls[k] < grid_full ? ctx->num_sms * occ_cls[k] : grid_full;
                    StageScope t(ctx, bs, PT_STAGE_SHADE, b);
                    k_shade_cls[k]<<<g, 256, 0, bs>>>(P);
                }
            } else {
                StageScope t(ctx, bs, PT_STAGE_SHADE, b);
                k_shade<<<grid_shade, bt_shade, shade_smem, bs>>>(P);
            }
        }
        if (waves < max_depth) {                          // whatever is still alive: finish in one launch
            P.bounce = tail_bounce;
            { StageScope t(ctx, bs, PT_STAGE_TAIL, P.bounce); k_tail<<<grid_tail, bt, smem, bs>>>(P); }
        }
        // resolves add to the accumulation buffer in batch order, whichever stream their batch ran on: the
        // sums are those of the serial schedule (bit-identical, options test wf_overlap=0/1)
        if (n_pipe > 1 && batch_index >= 1) PT_CUDA(cudaStreamWaitEvent(bs, ctx->ev_resolve[(batch_index - 1) % n_pipe], 0));
        int grid_res = (int)((n + bt - 1) / bt);
        if (grid_res > ctx->num_sms * 8) grid_res = ctx->num_sms * 8;
        { StageScope t(ctx, bs, PT_STAGE_RESOLVE, 0); k_wf_resolve<<<grid_res, bt, 0, bs>>>(P); }
        if (n_pipe > 1) PT_CUDA(cudaEventRecord(ctx->ev_resolve[pipe], bs));
    }
    // join: the caller's stream continues after the last resolve (which follows every other resolve, and each
    // stream's last kernel is a resolve)
    if (n_pipe > 1 && (batch_index - 1) % n_pipe != 0) PT_CUDA(cudaStreamWaitEvent(st, ctx->ev_resolve[(batch_index - 1) % n_pipe], 0));
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

// ---- stage-level entry points: one batch, one of the reference's wavefront kernels per call ----------------
int pt_wf_begin(PtContext* ctx, int W, int H, int y0, int y1, int sample_base, int samples, uint32_t seed, void* workspace,
                size_t workspace_bytes) {
    float dummy;
    int rc = check_render_args(ctx, &dummy, W, H, y0, y1, samples);
    if (rc) return rc;
    const int rows = y1 - y0;
    const size_t n = (size_t)W * rows;
    if (samples <= 0 || rows <= 0) return set_error(PT_ERR_INVALID_ARG, "pt_wf_begin: empty batch");
    if (!workspace || workspace_bytes < pt_wavefront_workspace_bytes_for(W, rows, samples))
        return set_error(PT_ERR_WORKSPACE, "pt_wf_begin: workspace too small for this many samples in flight");
    if ((size_t)samples * n > kMaxPathsInFlight || samples > 256) return set_error(PT_ERR_CAPACITY, "pt_wf_begin: batch too large");
    if (!ctx->stage_P) ctx->stage_P = new WfParams();
    WfParams& P = *ctx->stage_P;
    const size_t cap = (size_t)samples * n;
    char* p = static_cast<char*>(workspace);
    auto take = [&](size_t bytes) { char* q = p; p += align_up(bytes, 256); return q; };
    P.S = ctx->scene;
    P.cam = ctx->cam;
    for (int q = 0; q < 2; q++) {
        P.Q.ra[q] = reinterpret_cast<float4*>(take(cap * 32));
        P.Q.r2[q] = reinterpret_cast<float4*>(take(cap * 16));
    }
    P.Q.hit = reinterpret_cast<float4*>(take(cap * 16));
    P.Q.bins = reinterpret_cast<uint32_t*>(take((size_t)kNumBins * cap * 4));
    P.Q.slab = reinterpret_cast<float*>(take(cap * 12));
    P.Q.tie_list = reinterpret_cast<int*>(take(cap * 4));
    int* counters = reinterpret_cast<int*>(take(256));
    P.Q.qcount = counters;
    P.Q.bin_count = counters + 8;
    P.Q.cursor = counters + 32;
    P.Q.tie_count = counters + 40;
    P.Q.capacity = (int)cap;
    P.accum = nullptr;
    P.W = W; P.H = H; P.y0 = y0; P.y1 = y1;
    P.row_shift = ctx->row_shift; P.row_stride = ctx->row_stride;
    P.inv_W = 1.0f / (float)W;
    P.N = (int)n;
    P.inv_N = 1.0f / (float)n;
    P.sample_base = sample_base;
    P.batch = samples;
    P.seed = seed;
    P.n_snodes = 0;
    P.bounce = 0;
    P.sort = 1;
    P.burst_num = ctx->wf_burst_num;
    P.vote_num = ctx->wf_vote_num;
    P.fused_waves = 0;
    P.tile = ctx->wf_tile;
    P.counters = ctx->d_counters;
    return PT_OK;
}

static int wf_stage_params(PtContext* ctx, WfParams** P, const char* who) {
    if (!ctx) return set_error(PT_ERR_INVALID_ARG, std::string(who) + ": NULL context");
    if (!ctx->stage_P) return set_error(PT_ERR_NO_SCENE, std::string(who) + ": no batch bound (pt_wf_begin)");
    if (!ctx->has_scene || !ctx->has_camera) return set_error(PT_ERR_NO_SCENE, std::string(who) + ": scene or camera not set");
    PT_CUDA(cudaSetDevice(ctx->device));
    ctx->stage_P->S = ctx->scene;         // the scene / camera as they are now
    ctx->stage_P->cam = ctx->cam;
    *P = ctx->stage_P;
    return PT_OK;
}

int pt_wf_generate(PtContext* ctx, void* stream) {
    WfParams* P;
    int rc = wf_stage_params(ctx, &P, "pt_wf_generate");
    if (rc) return rc;
    const long long total = (long long)P->N * P->batch;
    int grid = (int)((total + 255) / 256);
    if (grid > ctx->num_sms * 8) grid = ctx->num_sms * 8;
    P->bounce = 0;
    k_wf_generate<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*P);
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_wf_extend(PtContext* ctx, int bounce, void* stream) {
    WfParams* P;
    int rc = wf_stage_params(ctx, &P, "pt_wf_extend");
    if (rc) return rc;
    if (bounce < 0) return set_error(PT_ERR_INVALID_ARG, "pt_wf_extend: negative bounce");
    P->bounce = bounce;
    const bool full = ctx->scene.full_features != 0;
    void (*k)(const WfParams) = ctx->count ? (full ? k_wf_extend<true, true, 4> : k_wf_extend<true, false, 4>)
                                           : (full ? k_wf_extend<false, true, 5> : k_wf_extend<false, false, 5>);
    const long long total = (long long)P->N * P->batch;
    const int grid_full = (int)((total + PT_EXT_WARPS * 32 - 1) / (PT_EXT_WARPS * 32));
    const int grid_persist = ctx->num_sms * ctx->wf_blocks_per_sm;
    k<<<grid_persist < grid_full ? grid_persist : grid_full, PT_EXT_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(*P);
    if (!ctx->count && PT_TIE) {
        if (full) k_wf_fix_ties<true><<<ctx->num_sms * 2, 128, 0, static_cast<cudaStream_t>(stream)>>>(*P);
        else k_wf_fix_ties<false><<<ctx->num_sms * 2, 128, 0, static_cast<cudaStream_t>(stream)>>>(*P);
    }
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_wf_shade(PtContext* ctx, int bounce, void* stream) {
    WfParams* P;
    int rc = wf_stage_params(ctx, &P, "pt_wf_shade");
    if (rc) return rc;
    if (bounce < 0) return set_error(PT_ERR_INVALID_ARG, "pt_wf_shade: negative bounce");
    P->bounce = bounce;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = (long long)P->N * P->batch;
    int grid = (int)((total + 255) / 256);
    const int grid_wide = ctx->num_sms * 8 < grid ? ctx->num_sms * 8 : grid;
    k_wf_bin<<<grid_wide, 256, 0, st>>>(*P);
    if (ctx->scene.full_features) k_wf_shade<true><<<grid_wide, PT_SHADE_THREADS, 0, st>>>(*P);
    else k_wf_shade<false><<<grid_wide, PT_SHADE_THREADS, 0, st>>>(*P);
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_wf_compact(PtContext* ctx, int bounce, int* live_rays, void* stream) {
    WfParams* P;
    int rc = wf_stage_params(ctx, &P, "pt_wf_compact");
    if (rc) return rc;
    if (!live_rays || bounce < 0) return set_error(PT_ERR_INVALID_ARG, "pt_wf_compact: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int v = 0;
    PT_CUDA(cudaMemcpyAsync(&v, P->Q.qcount + (bounce & 1), sizeof v, cudaMemcpyDeviceToHost, st));
    PT_CUDA(cudaStreamSynchronize(st));
    *live_rays = v;
    return PT_OK;
}

int pt_wf_read_rays(PtContext* ctx, int bounce, int n, float* o, float* d, float* throughput, int32_t* pixel, int32_t* depth, void* stream) {
    WfParams* P;
    int rc = wf_stage_params(ctx, &P, "pt_wf_read_rays");
    if (rc) return rc;
    if (n < 0 || n > P->Q.capacity || bounce < 0) return set_error(PT_ERR_INVALID_ARG, "pt_wf_read_rays: bad count / bounce");
    if (n == 0) return PT_OK;
    k_wf_unpack_rays<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(*P, bounce & 1, n, o, d, throughput, pixel, depth);
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_wf_read_hits(PtContext* ctx, int n, float* t, int32_t* prim, float* t_exit, void* stream) {
    WfParams* P;
    int rc = wf_stage_params(ctx, &P, "pt_wf_read_hits");
    if (rc) return rc;
    if (n < 0 || n > P->Q.capacity) return set_error(PT_ERR_INVALID_ARG, "pt_wf_read_hits: bad count");
    if (n == 0) return PT_OK;
    k_wf_unpack_hits<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(*P, n, t, prim, t_exit);
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_wf_resolve(PtContext* ctx, float* accum, void* stream) {
    WfParams* P;
    int rc = wf_stage_params(ctx, &P, "pt_wf_resolve");
    if (rc) return rc;
    if (!accum) return set_error(PT_ERR_INVALID_ARG, "pt_wf_resolve: NULL accumulation buffer");
    WfParams R = *P;
    R.accum = accum;
    int grid = (R.N + 255) / 256;
    if (grid > ctx->num_sms * 8) grid = ctx->num_sms * 8;
    k_wf_resolve<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(R);
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_generate_camera_rays(PtContext* ctx, int W, int H, uint32_t seed, int sample, float* o, float* d, void* stream) {
    if (!ctx || !o || !d) return set_error(PT_ERR_INVALID_ARG, "pt_generate_camera_rays: NULL argument");
    if (!ctx->has_camera) return set_error(PT_ERR_NO_SCENE, "pt_generate_camera_rays: camera not set");
    PT_CUDA(cudaSetDevice(ctx->device));
    const int n = W * H;
    k_generate_rays<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(ctx->cam, W, H, seed, sample, o, d);
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_first_hit(PtContext* ctx, const float* o, const float* d, int n, float t_min, float t_max, float* out_t,
                 int32_t* out_prim, float* out_p, float* out_n, void* stream) {
    if (!ctx || !o || !d || !out_t || !out_prim) return set_error(PT_ERR_INVALID_ARG, "pt_first_hit: NULL argument");
    if (!ctx->has_scene) return set_error(PT_ERR_NO_SCENE, "pt_first_hit: scene not set");
    if (n <= 0) return PT_OK;
    PT_CUDA(cudaSetDevice(ctx->device));
    FirstHitParams P;
    P.S = ctx->scene;
    P.o = o; P.d = d; P.n = n; P.t_min = t_min; P.t_max = t_max;
    P.out_t = out_t; P.out_prim = out_prim; P.out_p = out_p; P.out_n = out_n;
    P.n_snodes = pick_snodes(ctx, 4);
    P.counters = ctx->d_counters;
    const size_t smem = (size_t)P.n_snodes * 64;
    int grid = (n + 255) / 256;
    const int cap = ctx->num_sms * 8;
    if (grid > cap) grid = cap;
    int rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (ctx->count) {
        if ((rc = set_smem(k_first_hit<true>, smem))) return rc;
        k_first_hit<true><<<grid, 256, smem, st>>>(P);
    } else {
        if ((rc = set_smem(k_first_hit<false>, smem))) return rc;
        k_first_hit<false><<<grid, 256, smem, st>>>(P);
    }
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_math_eval(PtContext* ctx, int fn, int n, const float* a, const float* b, float* out, void* stream) {
    if (!ctx || !a || !out) return set_error(PT_ERR_INVALID_ARG, "pt_math_eval: NULL argument");
    if (n <= 0) return PT_OK;
    PT_CUDA(cudaSetDevice(ctx->device));
    k_math_eval<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(fn, n, a, b ? b : a, out);
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_tonemap(PtContext* ctx, const float* accum, int n_floats, int sample_count, uint8_t* out_rgb, void* stream) {
    if (!ctx || !accum || !out_rgb) return set_error(PT_ERR_INVALID_ARG, "pt_tonemap: NULL argument");
    if (n_floats <= 0) return PT_OK;
    PT_CUDA(cudaSetDevice(ctx->device));
    const float scale = 1.0f / (float)(sample_count > 1 ? sample_count : 1);
    k_tonemap<<<(n_floats + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(accum, n_floats, scale, out_rgb);
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_probe_gather(PtContext* ctx, const void* buffer, size_t bytes, int pattern, int iters, int repeats,
                    double* out_gbytes_per_s, void* stream) {
    if (!ctx || !buffer || !out_gbytes_per_s) return set_error(PT_ERR_INVALID_ARG, "pt_probe_gather: NULL argument");
    if (bytes < 64 || bytes / 64 > 0xffffffffull || iters <= 0 || repeats <= 0 || pattern < 0 || pattern > 1)
        return set_error(PT_ERR_INVALID_ARG, "pt_probe_gather: bad size / pattern / iteration count");
    PT_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int threads = PT_EXT_WARPS * 32, grid = ctx->num_sms * 6;           // the launch shape of k_wf_extend
    cudaEvent_t e0, e1;
    PT_CUDA(cudaEventCreate(&e0));
    PT_CUDA(cudaEventCreate(&e1));
    float best_ms = 0.0f;
    for (int r = 0; r < repeats + 1; r++) {                                    // first pass warms the caches
        cudaEventRecord(e0, st);
        k_probe_gather<<<grid, threads, 0, st>>>(static_cast<const float4*>(buffer), (unsigned)(bytes / 64), iters, pattern,
                                                 reinterpret_cast<float*>(ctx->d_counters));
        cudaEventRecord(e1, st);
        cudaError_t e = cudaEventSynchronize(e1);
        float ms = 0.0f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return set_error(PT_ERR_CUDA, cudaGetErrorString(e)); }
        if (r > 0 && (best_ms == 0.0f || ms < best_ms)) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    PT_CUDA(cudaGetLastError());
    const double gathered = (double)grid * threads * (double)iters * 4.0 * 64.0;
    *out_gbytes_per_s = gathered / (best_ms * 1e-3) / 1e9;
    return PT_OK;
}

int pt_get_debug_error(PtContext* ctx, uint32_t* out_code) {
    if (!ctx || !out_code) return set_error(PT_ERR_INVALID_ARG, "pt_get_debug_error: NULL argument");
    *out_code = 0;
#if PT_DEBUG_BOUNDS
    PT_CUDA(cudaSetDevice(ctx->device));
    PT_CUDA(cudaDeviceSynchronize());
    unsigned int v = 0;
    PT_CUDA(cudaMemcpyFromSymbol(&v, g_pt_debug_error, sizeof v));
    *out_code = v;
    if (v) {
        const unsigned int zero = 0;
        PT_CUDA(cudaMemcpyToSymbol(g_pt_debug_error, &zero, sizeof zero));
    }
#endif
    return PT_OK;
}

int pt_enable_counters(PtContext* ctx, int enable) {
    if (!ctx) return set_error(PT_ERR_INVALID_ARG, "pt_enable_counters: NULL context");
    ctx->count = enable != 0;
    return PT_OK;
}

int pt_get_counters(PtContext* ctx, PtCounters* out, int reset) {
    if (!ctx || !out) return set_error(PT_ERR_INVALID_ARG, "pt_get_counters: NULL argument");
    PT_CUDA(cudaSetDevice(ctx->device));
    PT_CUDA(cudaDeviceSynchronize());
    DCounters h;
    PT_CUDA(cudaMemcpy(&h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    out->rays = h.rays; out->box_tests = h.box_tests; out->sphere_tests = h.sphere_tests;
    out->triangle_tests = h.triangle_tests; out->quad_tests = h.quad_tests; out->paths = h.paths; out->bounces = h.bounces;
    if (reset) PT_CUDA(cudaMemset(ctx->d_counters, 0, sizeof(DCounters)));
    return PT_OK;
}

int pt_enable_stage_timing(PtContext* ctx, int enable) {
    if (!ctx) return set_error(PT_ERR_INVALID_ARG, "pt_enable_stage_timing: NULL context");
    ctx->stage_timing = enable != 0;
    if (enable && !ctx->d_stage_snap) {
        PT_CUDA(cudaSetDevice(ctx->device));
        PT_CUDA(cudaMalloc(&ctx->d_stage_snap, sizeof(DCounters) * kMaxStageSnaps));
        PT_CUDA(cudaMemset(ctx->d_stage_snap, 0, sizeof(DCounters) * kMaxStageSnaps));
    }
    return PT_OK;
}

int pt_get_stage_counters(PtContext* ctx, PtCounters* out, int max_entries, int* n_entries) {
    if (!ctx || !n_entries || (max_entries > 0 && !out)) return set_error(PT_ERR_INVALID_ARG, "pt_get_stage_counters: NULL argument");
    PT_CUDA(cudaSetDevice(ctx->device));
    PT_CUDA(cudaDeviceSynchronize());
    int n = (ctx->count && ctx->d_stage_snap) ? (int)ctx->marks.size() : 0;
    if (n > kMaxStageSnaps) n = kMaxStageSnaps;
    *n_entries = n;
    const int m = n < max_entries ? n : max_entries;
    if (m > 0) {
        std::vector<DCounters> h(m);
        PT_CUDA(cudaMemcpy(h.data(), ctx->d_stage_snap, sizeof(DCounters) * m, cudaMemcpyDeviceToHost));
        for (int i = 0; i < m; i++) {
            out[i].rays = h[i].rays; out[i].box_tests = h[i].box_tests; out[i].sphere_tests = h[i].sphere_tests;
            out[i].triangle_tests = h[i].triangle_tests; out[i].quad_tests = h[i].quad_tests;
            out[i].paths = h[i].paths; out[i].bounces = h[i].bounces;
        }
    }
    return PT_OK;
}

int pt_get_stage_times(PtContext* ctx, PtStageTime* out, int max_entries, int* n_entries) {
    if (!ctx || !n_entries || (max_entries > 0 && !out)) return set_error(PT_ERR_INVALID_ARG, "pt_get_stage_times: NULL argument");
    PT_CUDA(cudaSetDevice(ctx->device));
    PT_CUDA(cudaDeviceSynchronize());
    *n_entries = (int)ctx->marks.size();
    for (int i = 0; i < (int)ctx->marks.size() && i < max_entries; i++) {
        const PtContext::StageMark& m = ctx->marks[i];
        float ms = 0.0f;
        PT_CUDA(cudaEventElapsedTime(&ms, ctx->ev_pool[m.ev_begin], ctx->ev_pool[m.ev_end]));
        out[i].kernel = m.kernel;
        out[i].wave = m.wave;
        out[i].ms = ms;
    }
    return PT_OK;
}

}  // extern "C"

static void wf_stage_free(PtContext* ctx) {
    delete ctx->stage_P;
    ctx->stage_P = nullptr;
}
