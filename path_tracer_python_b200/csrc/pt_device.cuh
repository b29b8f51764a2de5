// Device-side building blocks of the B200 path tracer: vector type, packed-scene view, two-child
// BVH traversal, primitive tests, material evaluation.  Semantics follow the reference's Taichi
// functions (src/render_server/taichi_renderer/kernels.py, cited per function); the structure does
// not: nodes hold both children's boxes (one 64-byte fetch per visit instead of three 44-byte
// fetches, kernels.py:664,707-708), leaves are folded into child references (which also carry the
// shading class), every node is fetched through L1 (the whole tree is L2-resident), hit point / normal
// are reconstructed after the loop instead of being carried through it, and materials are de-duplicated
// records behind a per-primitive id.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pt_math.cuh"
#include "../../include/b200pt.h"

// ---- debug-bounds build (-DPT_DEBUG_BOUNDS=1 -> libb200pt_dbg.so): every index the kernels form -- node, primitive,
// material, image texel, queue slot, bin slot, slab slot, stack slot -- is checked before use.  The first violation is
// recorded in a device word (code << 24 | source line; pt_get_debug_error) and the kernel carries on, so that the
// word can still be read afterwards.  compute-sanitizer is not available on this GPU pool; tests/test_gpu_debug.py
// runs the golden suite through this build instead.  The product builds compile the checks away.
#ifndef PT_DEBUG_BOUNDS
#define PT_DEBUG_BOUNDS 0
#endif
#if PT_DEBUG_BOUNDS
__device__ unsigned int g_pt_debug_error = 0u;
__device__ __noinline__ void pt_debug_fail(unsigned code, unsigned line) { atomicCAS(&g_pt_debug_error, 0u, (code << 24) | (line & 0xffffffu)); }
#define PT_BOUND(cond, code) do { if (!(cond)) pt_debug_fail((code), __LINE__); } while (0)
#else
#define PT_BOUND(cond, code) ((void)0)
#endif
enum : unsigned { PT_DBG_NODE = 1, PT_DBG_PRIM = 2, PT_DBG_MATERIAL = 3, PT_DBG_STACK = 4, PT_DBG_QUEUE = 5, PT_DBG_BIN = 6,
                  PT_DBG_SLAB = 7, PT_DBG_TEXEL = 8, PT_DBG_PIXEL = 9 };

namespace pt {

struct f3 {
    float x, y, z;
};
__device__ __forceinline__ f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return f3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return f3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator-(f3 a) { return f3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return f3{a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return f3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return f3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ f3 vdiv(f3 a, float s) { return f3{ptm::fdiv(a.x, s), ptm::fdiv(a.y, s), ptm::fdiv(a.z, s)}; }
// Taichi Vector.dot: products summed left to right
__device__ __forceinline__ float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
    return f3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Taichi Vector.normalized(): (1 / norm) * v
__device__ __forceinline__ f3 normalized(f3 a) {
    float inv = ptm::frcp(ptm::fsqrt(dot(a, a)));
    return inv * a;
}
__device__ __forceinline__ float min2(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float max2(float a, float b) { return a > b ? a : b; }

// ---- packed scene as the kernels see it (pointers into the caller's device blob)
#ifndef PT_WIDE
#define PT_WIDE 0
#endif
#ifndef PT_TIE
#define PT_TIE (!PT_WIDE)   // exact-t ties resolved the reference's way (resolve_tie); the 4-wide variant build keeps its visiting order
#endif
constexpr int kRefNone = (int)0x80000000;   // "no child" / "no hit"
constexpr int kRefTie = 0x40000000;         // set in every leaf reference (primitive index < 2^25).  TravState.best_ref with the
                                            // bit CLEARED: another primitive was hit at exactly best_t (trav_take_tie)
constexpr int kStackSize = 64;              // kernels.py:649

struct DScene {
    const float4* nodes;      // 4 x float4 per inner node: (lmin.x lmax.x rmin.x rmax.x) (..y) (..z) (left,right,-,-)
    const float4* spheres;    // (cx cy cz r)
    const float4* quads;      // (n.xyz D) (Q.xyz w.x) (u.xyz w.y) (v.xyz w.z)
    const float4* tris;       // (v0.xyz n.x) (e1.xyz n.y) (e2.xyz n.z)
    const uint32_t* prim_mat; // material id per primitive: [spheres | triangles | quads]
    const float4* nodes_plain; // = nodes, except in the launch that traverses the origin-relative copy (resolve_tie)
    const float4* materials;  // 3 x float4 per unique material
    const float4* perlin_vec; // 256 gradient vectors
    const uint8_t* perlin_perm; // perm_x | perm_y | perm_z (256 each)
    const uchar4* image[PT_MAX_IMAGES];
    int image_w[PT_MAX_IMAGES], image_h[PT_MAX_IMAGES];
    int num_inner;
    int oct_stride;           // 0: `nodes` is the plain array.  N = num_inner: `nodes` holds 8 direction-octant variants of
                              // it (variant k at node index k * N, its inner references offset likewise): see trav_root
    int n_prims[3];           // spheres, triangles, quads (by PT_PRIM_*)
    int num_materials;
    int mat_base[3];          // offset of each prim type inside prim_mat
    int has_media;
    // "full" feature mode (moving spheres, subsurface materials); 0 = the reference's Taichi path
    const float4* sphere_motion;   // centre(1) - centre(0) per sphere
    int full_features;             // bit 0: motion blur + subsurface; bit 1: next-event estimation
    // next-event estimation (features "nee"): one record per emissive primitive, (leaf ref, area, emit.rg)(emit.b, -, -, -)
    const float4* lights;
    int n_lights;
    float nee_r2;                  // squared switch-over distance between the two sampling strategies (nee_direct)
};

struct DCounters {
    unsigned long long rays, box_tests, sphere_tests, triangle_tests, quad_tests, paths, bounces;
};

struct LocalCounters {
    unsigned int rays = 0, box_tests = 0, sphere_tests = 0, triangle_tests = 0, quad_tests = 0, paths = 0, bounces = 0;
};

__device__ __forceinline__ int ref_type(int ref) { return (~ref) & 3; }
__device__ __forceinline__ int ref_index(int ref) { return (~ref) >> 5; }
// Shading class packed into the reference by the host packer (pt_host.cpp shading_class); 0 is what the
// wavefront kernels use for a miss.  The order is the order of the material bins, chosen so that each
// specialised shade kernel owns a contiguous run: [miss, emissive] end the path with one store,
// [Lambert with a solid colour] is the bulk of every scene, the rest needs the general code.
enum : int { PT_CLS_MISS = 0, PT_CLS_EMISSIVE = 1, PT_CLS_LAMBERT_SOLID = 2, PT_CLS_LAMBERT_TEX = 3, PT_CLS_METAL = 4,
             PT_CLS_DIELECTRIC = 5, PT_CLS_MEDIUM = 6, PT_CLS_OTHER = 7 };
__device__ __forceinline__ int ref_class(int ref) { return ((~ref) >> 2) & 7; }
__device__ __forceinline__ bool ref_is_medium(int ref) { return ref_class(ref) == PT_CLS_MEDIUM; }

// ---- primitive tests: arithmetic identical to hit_sphere / hit_triangle / hit_quad; they return
//      only t (the caller keeps t and the primitive reference).
//      They are written with the explicitly rounded intrinsics (__fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn /
//      __fsqrt_rn: never contracted, never approximated), so BOTH builds of the library evaluate them
//      bit-identically: the reference's formulas are ill-conditioned where it matters (|oc|^2 - r^2 for a
//      radius-1000 ground sphere, D - n.o for a far plane), and any other rounding of them moves t by 1e-5 ..
//      1e-4 relative -- more than north_star's tolerance.  The fast build therefore keeps the strict
//      primitive tests and spends its freedom on the slab test and the shading arithmetic.
namespace sx {
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ f3 sub(f3 a, f3 b) { return f3{sub(a.x, b.x), sub(a.y, b.y), sub(a.z, b.z)}; }
__device__ __forceinline__ float dot(f3 a, f3 b) { return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z)); }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
    return f3{sub(mul(a.y, b.z), mul(a.z, b.y)), sub(mul(a.z, b.x), mul(a.x, b.z)), sub(mul(a.x, b.y), mul(a.y, b.x))};
}
}  // namespace sx
// kernels.py:208-248
// centre(time) = centre(0) + time * motion (core/sphere.py:36); time is only used in full mode
template <bool FULL>
__device__ __forceinline__ f3 sphere_center(const DScene& S, int idx, float4 s, float time) {
    f3 c = mk3(s.x, s.y, s.z);
    if (FULL) {
        float4 m = __ldg(S.sphere_motion + idx);
        c = mk3(sx::add(c.x, sx::mul(time, m.x)), sx::add(c.y, sx::mul(time, m.y)), sx::add(c.z, sx::mul(time, m.z)));
    }
    return c;
}
template <bool FULL>
__device__ __forceinline__ bool hit_sphere_t(const DScene& S, int idx, f3 o, f3 d, float t_min, float t_max, float& t,
                                             float time) {
    PT_BOUND(idx >= 0 && idx < S.n_prims[PT_PRIM_SPHERE], PT_DBG_PRIM);
    float4 s = __ldg(S.spheres + idx);
    f3 oc = sx::sub(sphere_center<FULL>(S, idx, s, time), o);
    float a = sx::dot(d, d);
    float h = sx::dot(d, oc);
    float c = sx::sub(sx::dot(oc, oc), sx::mul(s.w, s.w));
    float disc = sx::sub(sx::mul(h, h), sx::mul(a, c));
    if (disc >= 0.0f) {
        float sq = __fsqrt_rn(disc);
        float root = sx::div(sx::sub(h, sq), a);
        if (root < t_min || root > t_max) root = sx::div(sx::add(h, sq), a);
        if (root >= t_min && root <= t_max) { t = root; return true; }
    }
    return false;
}
// kernels.py:251-307
__device__ __forceinline__ bool hit_triangle_t(const DScene& S, int idx, f3 o, f3 d, float t_min, float t_max, float& t) {
    PT_BOUND(idx >= 0 && idx < S.n_prims[PT_PRIM_TRIANGLE], PT_DBG_PRIM);
    const float4* p = S.tris + 3 * idx;
    float4 a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2);
    f3 v0 = mk3(a0.x, a0.y, a0.z), e1 = mk3(a1.x, a1.y, a1.z), e2 = mk3(a2.x, a2.y, a2.z);
    f3 h = sx::cross(d, e2);
    float det = sx::dot(e1, h);
    if (fabsf(det) >= 1e-8f) {
        float inv_det = __frcp_rn(det);
        f3 s = sx::sub(o, v0);
        float u = sx::mul(inv_det, sx::dot(s, h));
        if (u >= 0.0f && u <= 1.0f) {
            f3 q = sx::cross(s, e1);
            float v = sx::mul(inv_det, sx::dot(d, q));
            if (v >= 0.0f && sx::add(u, v) <= 1.0f) {
                float tt = sx::mul(inv_det, sx::dot(e2, q));
                if (tt >= t_min && tt <= t_max) { t = tt; return true; }
            }
        }
    }
    return false;
}
// kernels.py:310-362
__device__ __forceinline__ bool hit_quad_t(const DScene& S, int idx, f3 o, f3 d, float t_min, float t_max, float& t) {
    PT_BOUND(idx >= 0 && idx < S.n_prims[PT_PRIM_QUAD], PT_DBG_PRIM);
    const float4* p = S.quads + 4 * idx;
    float4 q0 = __ldg(p);
    f3 n = mk3(q0.x, q0.y, q0.z);
    float denom = sx::dot(n, d);
    if (fabsf(denom) >= 1e-8f) {
        float tt = sx::div(sx::sub(q0.w, sx::dot(n, o)), denom);
        if (tt >= t_min && tt <= t_max) {
            float4 q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3);
            f3 isect = mk3(sx::add(o.x, sx::mul(tt, d.x)), sx::add(o.y, sx::mul(tt, d.y)), sx::add(o.z, sx::mul(tt, d.z)));
            f3 ph = sx::sub(isect, mk3(q1.x, q1.y, q1.z));
            f3 u = mk3(q2.x, q2.y, q2.z), v = mk3(q3.x, q3.y, q3.z), w = mk3(q1.w, q2.w, q3.w);
            float alpha = sx::dot(w, sx::cross(ph, v));
            float beta = sx::dot(w, sx::cross(u, ph));
            if (alpha >= 0.0f && alpha <= 1.0f && beta >= 0.0f && beta <= 1.0f) { t = tt; return true; }
        }
    }
    return false;
}

// hit point and shading normal of the closest hit, exactly as hit_* would have returned them
template <bool FULL = false>
__device__ __forceinline__ void finalize_hit(const DScene& S, int ref, f3 o, f3 d, float t, f3& p, f3& n,
                                             float time = 0.0f) {
    int type = ref_type(ref), idx = ref_index(ref);
    PT_BOUND(ref < 0 && ref != kRefNone && type <= 2 && idx < S.n_prims[type], PT_DBG_PRIM);
    p = o + t * d;
    if (type == PT_PRIM_SPHERE) {
        float4 s = __ldg(S.spheres + idx);
        n = vdiv(p - sphere_center<FULL>(S, idx, s, time), s.w);     // kernels.py:246 (never flipped)
    } else if (type == PT_PRIM_QUAD) {
        float4 q0 = __ldg(S.quads + 4 * idx);
        f3 nn = mk3(q0.x, q0.y, q0.z);
        n = dot(nn, d) < 0.0f ? nn : -nn;                      // kernels.py:355-360
    } else {
        const float4* tp = S.tris + 3 * idx;
        f3 nn = mk3(__ldg(tp).w, __ldg(tp + 1).w, __ldg(tp + 2).w);
        n = dot(d, nn) > 0.0f ? -nn : nn;                      // kernels.py:300-305
    }
}

// ---- BVH traversal: closest hit over the same primitive set the reference's traverse_bvh_legacy
//      (kernels.py:624-742) reaches; the slab test is hit_aabb_optimized (kernels.py:600-621) with
//      the reference's clamped reciprocal direction (kernels.py:642-646).  Children are visited
//      nearest-entry first; a deferred child is skipped at pop time when its entry distance exceeds
//      the current closest hit -- the same predicate the reference re-evaluates on pop (:667).
//
//      The traversal is *resumable*: all of its state lives in TravState (+ the per-lane stack), so a
//      warp can leave the loop when too few of its lanes are still traversing, let the finished
//      lanes shade / fetch new work, and come back (persistent-lane scheduling).

// min/max of the slab test map to single FMNMX / FMNMX3 instructions (sm_100a).  Float min/max is
// exact and order-independent for non-NaN inputs; the results only feed comparisons, so the sign of
// a zero cannot matter either.
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// PT_FAST_SLAB=1: the FMA form of the slab test (one FMA per plane against -(o * inv), conservative box test) -- what
// the fast-numerics build used until it was measured against the strict test in the same build (config 2, same box):
// wavefront 3105 vs 3287 Mpaths/s, megakernel 2345 vs 2468.  Packed, the strict test is 6 FADD2 + 6 FMUL2 per node
// visit against 6 FFMA2 + 2 FFMA, and it needs four registers less per ray (no -(o * inv), no slack) in kernels
// that sit on their register cap; with the strict boxes the primary wave also gets the origin-relative node copy.
// So the fast build (PT_FAST=1) now keeps the strict box test too -- its first hits ARE the strict build's -- and is
// "fast" in everything else: FMA contraction, approximate reciprocal / division / square root.  -DPT_FAST_SLAB=1
// rebuilds the FMA variant.
#ifndef PT_FAST_SLAB
#define PT_FAST_SLAB 0
#endif
// One plane of the slab test, hit_aabb_optimized (kernels.py:609-614): (b - o) * inv, two roundings
// (strict) -- or one FFMA against the precomputed -(o * inv) in the fast-numerics build.
#if PT_FAST_SLAB
#define PT_SLAB(b, o_, inv_, noi_) __fmaf_rn((b), (inv_), (noi_))
// fl(b * inv - fl(o * inv)) is off by up to eps * (|o * inv| + |t|) from the exact plane distance, which for a
// ray that starts far from a thin box (|o * inv| >> thickness * inv) is enough to make the box vanish.  The
// fast build therefore tests  tmax * (1 + 2^-20) + slack >= tmin  with slack = 2^-21 * max |o * inv| per
// ray: never culls a box the strict test keeps (one FFMA per box; the result is a superset traversal).
#define PT_BOX_HIT(tmax, tmin, slack) (__fmaf_rn((tmax), 1.00000095f, (slack)) >= (tmin))
#else
#define PT_SLAB(b, o_, inv_, noi_) (((b) - (o_)) * (inv_))
#define PT_BOX_HIT(tmax, tmin, slack) ((tmax) >= (tmin))
#endif
// Blackwell's packed fp32 arithmetic (sub / mul .f32x2 -> FADD2 / FMUL2, sm_100+): two IEEE round-to-nearest
// operations per issue slot, the scalar operand broadcast by the instruction itself (Rn.F32).  The two planes
// of one axis of one child box arrive as an aligned register pair out of the 256-bit node load, so the strict
// slab test's 12 FADD + 12 FMUL per node visit become 6 FADD2 + 6 FMUL2 -- same roundings, same bits
// (no .ftz: subnormals are kept, as in the scalar code under -ftz=false).
// The fast-numerics build packs its one-FFMA plane test the same way: fma.rn.f32x2 (FFMA2) against the broadcast
// reciprocal direction and -(o * inv).
#ifndef PT_F32X2
#define PT_F32X2 1
#endif
__device__ __forceinline__ void slab2(float b0, float b1, float o, float inv, float noi, float& r0, float& r1) {
#if PT_F32X2 && PT_FAST_SLAB
    asm("{\n\t.reg .b64 v, ii, nn;\n\t"
        "mov.b64 v, {%2, %3};\n\tmov.b64 ii, {%4, %4};\n\tmov.b64 nn, {%5, %5};\n\t"
        "fma.rn.f32x2 v, v, ii, nn;\n\t"
        "mov.b64 {%0, %1}, v;\n\t}"
        : "=f"(r0), "=f"(r1) : "f"(b0), "f"(b1), "f"(inv), "f"(noi));
#elif PT_F32X2
    asm("{\n\t.reg .b64 v, oo, ii;\n\t"
        "mov.b64 v, {%2, %3};\n\tmov.b64 oo, {%4, %4};\n\tmov.b64 ii, {%5, %5};\n\t"
        "sub.rn.f32x2 v, v, oo;\n\tmul.rn.f32x2 v, v, ii;\n\t"
        "mov.b64 {%0, %1}, v;\n\t}"
        : "=f"(r0), "=f"(r1) : "f"(b0), "f"(b1), "f"(o), "f"(inv));
#else
    r0 = PT_SLAB(b0, o, inv, noi);
    r1 = PT_SLAB(b1, o, inv, noi);
#endif
}
// ... and the origin-relative form (boxes with the shared ray origin already subtracted): one FMUL2
__device__ __forceinline__ void slab2_rel(float b0, float b1, float inv, float& r0, float& r1) {
#if PT_F32X2
    asm("{\n\t.reg .b64 v, ii;\n\t"
        "mov.b64 v, {%2, %3};\n\tmov.b64 ii, {%4, %4};\n\t"
        "mul.rn.f32x2 v, v, ii;\n\t"
        "mov.b64 {%0, %1}, v;\n\t}"
        : "=f"(r0), "=f"(r1) : "f"(b0), "f"(b1), "f"(inv));
#else
    r0 = b0 * inv;
    r1 = b1 * inv;
#endif
}

typedef unsigned long long StackEntry;
#ifndef PT_STACK_TOP
#define PT_STACK_TOP 1      // see trav_push / trav_pop
#endif

struct TravState {
    f3 o, d, inv;
#if PT_FAST_SLAB
    f3 noi;              // -(o * inv): the slab test is fma(b, inv, noi) per plane (fast numerics only)
    float slack;         // 2^-21 * max |o * inv| (PT_BOX_HIT)
#endif
    float time;          // shutter time of the path (full feature mode; 0 otherwise)
    float t_min, best_t;
    int best_ref;
    int cur;     // >= 0 inner node, < 0 leaf reference, kRefNone = traversal finished
    int sp;      // entries in the local-memory part of the stack
#if PT_STACK_TOP
    unsigned long long top;   // newest stack entry, held in registers (low word kRefNone = stack empty)
#endif
};

// ---- direction-octant node variants.  Variant k of the node array (bit 0 / 1 / 2 of k = the ray runs towards -x / -y / -z)
// stores the two planes of an axis in the order the ray meets them: (max, min) instead of (min, max) on the axes the ray
// descends.  The slab distances (b - o) * inv are the same numbers whichever slot they come from, and
// min(a0, a1) is a0 by construction (fp subtraction and multiplication are monotone), so for a ray that traverses ITS
// variant the six pair-wise FMNMX of a child box disappear: 8 min/max per node visit instead of 20 -- on the ALU pipe,
// the most loaded one in the traversal kernels.  A variant is a complete tree (its inner references point into itself),
// so picking it costs nothing but the root index.  Built per render call in the caller's workspace (k_make_oct_nodes).
// Measured on config 2: the fused primary wave (coherent camera rays: a warp stays in one or two variants) -19 %;
// k_wf_extend (incoherent bounce rays: eight times the node footprint in L1) +5.5 %, and a run-time switch between
// the two forms of the test costs that kernel another 5 % -- so only trav_run<..., OCT> (the primary wave) has it.
__device__ __forceinline__ int trav_root(const DScene& S, f3 inv) {
    if (S.num_inner <= 0) return kRefNone;
    const int k = (inv.x < 0.0f ? 1 : 0) | (inv.y < 0.0f ? 2 : 0) | (inv.z < 0.0f ? 4 : 0);
    return k * S.oct_stride;
}

template <bool COUNT>
__device__ __forceinline__ void trav_begin(TravState& T, const DScene& S, f3 o, f3 d, float t_min, float t_max,
                                           LocalCounters& lc, float time = 0.0f) {
    T.o = o;
    T.d = d;
    T.time = time;
    ptm::inv_dir3(d.x, d.y, d.z, T.inv.x, T.inv.y, T.inv.z);
#if PT_FAST_SLAB
    T.noi = mk3(-(o.x * T.inv.x), -(o.y * T.inv.y), -(o.z * T.inv.z));
    T.slack = 4.76837158e-7f * fmaxf(fmaxf(fabsf(T.noi.x), fabsf(T.noi.y)), fabsf(T.noi.z));
#endif
    T.t_min = t_min;
    T.best_t = t_max;
    T.best_ref = kRefNone;
    T.cur = S.num_inner > 0 ? 0 : kRefNone;
    T.sp = 0;
#if PT_STACK_TOP
    T.top = (unsigned long long)(unsigned)kRefNone;
#endif
    if (COUNT) lc.rays++;
}

// Second search along the SAME ray from a later t_min (the exit search behind a constant-medium
// boundary, kernels.py:417-419): origin, direction and reciprocal direction stay as they are.
template <bool COUNT>
__device__ __forceinline__ void trav_restart(TravState& T, const DScene& S, float t_min, float t_max, LocalCounters& lc) {
    T.t_min = t_min;
    T.best_t = t_max;
    T.best_ref = kRefNone;
    T.cur = S.num_inner > 0 ? 0 : kRefNone;
    T.sp = 0;
#if PT_STACK_TOP
    T.top = (unsigned long long)(unsigned)kRefNone;
#endif
    if (COUNT) lc.rays++;
}

// A stack entry packs the deferred child's reference and its slab entry distance into one 8-byte
// word, so that a push / pop is a single local-memory instruction (STL.64 / LDL.64).
__device__ __forceinline__ StackEntry stack_pack(int ref, float t) {
    return ((StackEntry)__float_as_uint(t) << 32) | (StackEntry)(unsigned)ref;
}
// PT_STACK_TOP=1 (default) keeps the newest entry in registers (T.top) and local memory holds the
// rest, so that a pop hands out T.top and only *starts* the load of its successor: the local-memory
// latency leaves the critical path of the pop.  It only pays if that load lands in T.top's own
// registers -- with the natural code order the compiler loads into a temporary and copies, which
// waits for the load on the spot (measured +-0 %); with the order below: wavefront +0.8 %,
// megakernel +0.7 % on config 2 (tools/ab.py, same box).  PT_STACK_TOP=0 builds the plain stack.
// Either way, entries whose slab entry distance is already beyond the closest hit are skipped on
// pop (the box test would reject them anyway).
// (PT_STACK_TOP=1) memory slots 1..T.sp hold the older entries and slot 0 a permanent sentinel
// (stack_init), so that the refill of T.top after a pop is one unconditional load.
constexpr int kStackSlots = kStackSize + PT_STACK_TOP;
// Where the stack slots live.  LocalStack: the per-lane array in local memory (every kernel but k_wf_extend).
// HybridStack<S>: slots 0..S-1 in shared memory, one column per thread (entry i of thread t at s[i * stride + t]:
// the bank depends on t only, so lanes at different depths never conflict), deeper slots in local memory.
// A push or pop in shared memory is two conflict-free wavefronts whatever the lanes' depths; the same access
// to the lane-interleaved local array touches one 256-byte line per distinct depth in the warp and competes
// with the node fetches for L1's tag stage (ncu round 1: 262 M local-memory sectors against 397 M global-load
// sectors per wave-1 launch of k_wf_extend).
// the entry distance of a stack entry: all of it (binary build) or without the two child-slot bits (wide build)
#if PT_WIDE
constexpr unsigned kDistMask = ~3u;
#else
constexpr unsigned kDistMask = ~0u;
#endif
struct LocalStack {
    StackEntry* p;
    __device__ __forceinline__ StackEntry load(int i) const { return p[i]; }
    __device__ __forceinline__ void store(int i, StackEntry e) const { p[i] = e; }
};
template <int S>
struct HybridStack {
    StackEntry* s;       // this thread's column of the CTA's shared-memory slots
    int stride;          // threads per CTA
    StackEntry* l;       // local-memory overflow, slots S..
    __device__ __forceinline__ StackEntry load(int i) const { return i < S ? s[i * stride] : l[i - S]; }
    __device__ __forceinline__ void store(int i, StackEntry e) const {
        if (i < S) s[i * stride] = e;
        else l[i - S] = e;
    }
};
template <class STK>
__device__ __forceinline__ void stack_init(const STK& stack) {
#if PT_STACK_TOP
    stack.store(0, (StackEntry)(unsigned)kRefNone);
#endif
}
template <class STK>
__device__ __forceinline__ void trav_push(TravState& T, const STK& stack, StackEntry e) {
#if PT_STACK_TOP
    if ((int)(unsigned)T.top != kRefNone) {      // cannot overflow: pt_set_scene rejects trees deeper than kStackSize
        PT_BOUND(T.sp + 1 < kStackSlots, PT_DBG_STACK);
        T.sp++;
        stack.store(T.sp, T.top);
    }
    T.top = e;
#else
    if (T.sp < kStackSize) {
        stack.store(T.sp, e);
        T.sp++;
    }
#endif
}
template <class STK>
__device__ __forceinline__ int trav_pop(TravState& T, const STK& stack) {
    int next = kRefNone;
#if PT_STACK_TOP
    while ((int)(unsigned)T.top != kRefNone) {
        const bool take = __uint_as_float((unsigned)(T.top >> 32) & kDistMask) <= T.best_t;
        const int cand = (int)(unsigned)T.top;
        // the old top is consumed above; the barrier keeps the compiler from hoisting the load over
        // those uses, so it can land in T.top's own registers and nothing waits for it here
        asm volatile("" ::: "memory");
        T.top = stack.load(T.sp);            // slot 0 = sentinel: no branch
        T.sp = T.sp > 0 ? T.sp - 1 : 0;
        if (take) { next = cand; break; }
    }
#else
    while (T.sp > 0) {
        --T.sp;
        const StackEntry e = stack.load(T.sp);
        if (__uint_as_float((unsigned)(e >> 32) & kDistMask) <= T.best_t) { next = (int)(unsigned)e; break; }
    }
#endif
    return next;
}

// Fetch of one 64-byte node: four read-only loads through L1 (hit rate 70-75 % on config 2; the whole
// tree is L2-resident).
// History, all measured on config 2 with tools/ab.py: staging the top 256 nodes (breadth-first
// order) in shared memory first paid (+1 % over none), and was refined twice -- shared-window base
// formed once per kernel instead of per visit, then ONE generic load sequence through a selected
// address instead of a predicated LDS/LDG pair (+1.4 % wavefront, +4 % megakernel).  With the rest
// of the node visit trimmed, the address select alone (ISETP, 2 x SEL, shift) costs more than the
// staged levels save: no staging at all is +3.0 % / +3.2 %, and returns 16 KB per CTA to L1.
// PT_SNODES=1 rebuilds the staged variant (options "smem_nodes" then takes effect again).
#ifndef PT_SNODES
#define PT_SNODES 0
#endif
#ifndef PT_LDG256
#define PT_LDG256 1
#endif
// two consecutive float4 (32-byte aligned) through the read-only path in one 256-bit load
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
// ... and the plain (coherent) 256-bit load / store for queue records that other kernels write
__device__ __forceinline__ void ld256(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p) : "memory");
}
__device__ __forceinline__ void st256(float4* p, float4 a, float4 b) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
typedef unsigned long long SNodeBase;
__device__ __forceinline__ SNodeBase snode_base(const float4* s_node_mem) {
#if PT_SNODES
    unsigned long long a = (unsigned long long)s_node_mem, r;
    asm volatile("mov.u64 %0, %1;" : "=l"(r) : "l"(a));      // opaque: the generic address stays in a register
    return r;
#else
    return 0ull;
#endif
}
__device__ __forceinline__ void load_node(const float4* __restrict__ nodes, SNodeBase s_nodes, int n_snodes, int cur,
                                          float4& nx, float4& ny, float4& nz, float4& nc, int n_nodes_dbg = 0) {
#if PT_SNODES
    const unsigned long long base = cur < n_snodes ? s_nodes : (unsigned long long)nodes;
    const float4* p = reinterpret_cast<const float4*>(base) + 4 * cur;
    nx = p[0]; ny = p[1]; nz = p[2];
    const float2 c2 = *reinterpret_cast<const float2*>(p + 3);
#else
    PT_BOUND(cur >= 0 && cur < n_nodes_dbg, PT_DBG_NODE);
    // one IMAD.WIDE: the 64-byte stride applied in 64-bit arithmetic (4 * cur as an int costs a shift first)
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(nodes) + (unsigned long long)(unsigned)cur * 64ull);
#if PT_LDG256
    // Blackwell's 256-bit global load (LDG.E.256, sm_100+): the 64-byte node in TWO load instructions instead of
    // four -- half the LSU issue slots and half the L1 tag look-ups of a divergent fetch (every lane its own node)
    float2 c2;
    float pad0, pad1;
    ldg256(p, nx, ny);
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(nz.x), "=f"(nz.y), "=f"(nz.z), "=f"(nz.w), "=f"(c2.x), "=f"(c2.y), "=f"(pad0), "=f"(pad1) : "l"(p + 2));
#else
    nx = __ldg(p); ny = __ldg(p + 1); nz = __ldg(p + 2);
    const float2 c2 = __ldg(reinterpret_cast<const float2*>(p + 3));
#endif
#endif
    nc = make_float4(c2.x, c2.y, 0.0f, 0.0f);
}

#if PT_WIDE
// ---- the 4-wide node visit (PT_WIDE=1 build): one 128-byte node = four 256-bit loads, four slab tests, the hit
// children sorted by entry distance with a 5-comparator network on 32-bit keys (entry distance with the child slot in
// its two lowest mantissa bits; distances are positive, so their bit patterns order like the numbers), the nearest
// one descended into and the others pushed farthest first.  Returns the next node / leaf reference (or pops).
constexpr int kNodeQuads = 8;        // float4 per node
__device__ __forceinline__ unsigned umin2(unsigned a, unsigned b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned umax2(unsigned a, unsigned b) { return a > b ? a : b; }
__device__ __forceinline__ int sel4(float4 r, unsigned slot) {
    const float lo = (slot & 1u) ? r.y : r.x, hi = (slot & 1u) ? r.w : r.z;
    return __float_as_int((slot & 2u) ? hi : lo);
}
template <bool REL, class STK>
__device__ __forceinline__ int visit_wide(TravState& T, const float4* __restrict__ nodes, int cur, f3 o, f3 inv, f3 noi, float slack,
                                          float t_min, const STK& stack, int n_nodes_dbg) {
    PT_BOUND(cur >= 0 && cur < n_nodes_dbg, PT_DBG_NODE);
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(nodes) + (unsigned long long)(unsigned)cur * 128ull);
    float4 mnx, mxx, mny, mxy, mnz, mxz, refs, pad;
    ldg256(p, mnx, mxx);
    ldg256(p + 2, mny, mxy);
    ldg256(p + 4, mnz, mxz);
    ldg256(p + 6, refs, pad);
    const float best_t = T.best_t;
    unsigned key[4];
#define PT_WIDE_CHILD(K, C)                                                                                          \
    {                                                                                                                \
        float a0, a1, b0, b1, c0, c1;                                                                                \
        if (REL) {                                                                                                   \
            a0 = mnx.C * inv.x; a1 = mxx.C * inv.x; b0 = mny.C * inv.y; b1 = mxy.C * inv.y; c0 = mnz.C * inv.z; c1 = mxz.C * inv.z; \
        } else {                                                                                                     \
            a0 = PT_SLAB(mnx.C, o.x, inv.x, noi.x); a1 = PT_SLAB(mxx.C, o.x, inv.x, noi.x);                          \
            b0 = PT_SLAB(mny.C, o.y, inv.y, noi.y); b1 = PT_SLAB(mxy.C, o.y, inv.y, noi.y);                          \
            c0 = PT_SLAB(mnz.C, o.z, inv.z, noi.z); c1 = PT_SLAB(mxz.C, o.z, inv.z, noi.z);                          \
        }                                                                                                            \
        const float tn = fmax3(fminf(a0, a1), fminf(b0, b1), fmaxf(fminf(c0, c1), t_min));                           \
        const float tf = fmin3(fmaxf(a0, a1), fmaxf(b0, b1), fminf(fmaxf(c0, c1), best_t));                          \
        key[K] = PT_BOX_HIT(tf, tn, slack) ? ((__float_as_uint(tn) & ~3u) | (unsigned)(K)) : 0xffffffffu;            \
    }
    PT_WIDE_CHILD(0, x)
    PT_WIDE_CHILD(1, y)
    PT_WIDE_CHILD(2, z)
    PT_WIDE_CHILD(3, w)
#undef PT_WIDE_CHILD
    // sorting network (0,1)(2,3)(0,2)(1,3)(1,2): ascending, misses (0xffffffff) last
    unsigned a = umin2(key[0], key[1]), b = umax2(key[0], key[1]), c = umin2(key[2], key[3]), d = umax2(key[2], key[3]);
    const unsigned k0 = umin2(a, c), m1 = umax2(a, c), m2 = umin2(b, d), k3 = umax2(b, d);
    const unsigned k1 = umin2(m1, m2), k2 = umax2(m1, m2);
    if (k0 == 0xffffffffu) return trav_pop(T, stack);
    // farthest first, so that the nearest deferred child is popped first; the key (distance, low bits = slot) rides in the entry
    if (k3 != 0xffffffffu) trav_push(T, stack, ((StackEntry)k3 << 32) | (StackEntry)(unsigned)sel4(refs, k3 & 3u));
    if (k2 != 0xffffffffu) trav_push(T, stack, ((StackEntry)k2 << 32) | (StackEntry)(unsigned)sel4(refs, k2 & 3u));
    if (k1 != 0xffffffffu) trav_push(T, stack, ((StackEntry)k1 << 32) | (StackEntry)(unsigned)sel4(refs, k1 & 3u));
    return sel4(refs, k0 & 3u);
}
#else
constexpr int kNodeQuads = 4;
#endif

// Runs the traversal until it finishes (returns true) or, when min_active > 0, until fewer than
// min_active lanes of the warp are still inside the loop (returns false; call again to resume).
// REL: S.nodes holds the node boxes with the ray origin already subtracted (every ray of the launch
// shares that origin: camera rays without defocus; b200pt.cu k_make_rel_nodes).  fl(b - o) is the
// same single rounding whoever performs it, so the slab test drops its 12 subtractions per node.
template <bool COUNT, bool FULL = false, bool REL = false, class STK = LocalStack, bool OCT = false>
__device__ __forceinline__ bool trav_run(TravState& T, const DScene& S, SNodeBase s_nodes, int n_snodes,
                                         const STK& stack, int min_active, LocalCounters& lc) {
    const f3 o = REL ? mk3(0.0f, 0.0f, 0.0f) : T.o;
    const f3 inv = T.inv;
#if PT_FAST_SLAB
    const f3 noi = T.noi;
    const float slack = T.slack;
#else
    const f3 noi = mk3(0.0f, 0.0f, 0.0f);
    const float slack = 0.0f;
#endif
    const float t_min = T.t_min;
    const float4* const nodes = S.nodes;
    while (T.cur != kRefNone) {
        while (T.cur >= 0) {
#if PT_WIDE
            if (COUNT) lc.box_tests += 4;
            T.cur = visit_wide<REL>(T, nodes, T.cur, o, inv, noi, slack, t_min, stack, S.num_inner);
            continue;
#endif
            float4 nx, ny, nz, nc;
            load_node(nodes, s_nodes, n_snodes, T.cur, nx, ny, nz, nc, S.num_inner + 7 * S.oct_stride);
            if (COUNT) lc.box_tests += 2;
            const float best_t = T.best_t;
            float l0x, l1x, l0y, l1y, l0z, l1z, r0x, r1x, r0y, r1y, r0z, r1z;
#if PT_F32X2
            if (REL) {
                slab2_rel(nx.x, nx.y, inv.x, l0x, l1x); slab2_rel(ny.x, ny.y, inv.y, l0y, l1y); slab2_rel(nz.x, nz.y, inv.z, l0z, l1z);
                slab2_rel(nx.z, nx.w, inv.x, r0x, r1x); slab2_rel(ny.z, ny.w, inv.y, r0y, r1y); slab2_rel(nz.z, nz.w, inv.z, r0z, r1z);
            } else {
                slab2(nx.x, nx.y, o.x, inv.x, noi.x, l0x, l1x); slab2(ny.x, ny.y, o.y, inv.y, noi.y, l0y, l1y); slab2(nz.x, nz.y, o.z, inv.z, noi.z, l0z, l1z);
                slab2(nx.z, nx.w, o.x, inv.x, noi.x, r0x, r1x); slab2(ny.z, ny.w, o.y, inv.y, noi.y, r0y, r1y); slab2(nz.z, nz.w, o.z, inv.z, noi.z, r0z, r1z);
            }
#else
            if (REL) {
                l0x = nx.x * inv.x; l1x = nx.y * inv.x; l0y = ny.x * inv.y; l1y = ny.y * inv.y; l0z = nz.x * inv.z; l1z = nz.y * inv.z;
                r0x = nx.z * inv.x; r1x = nx.w * inv.x; r0y = ny.z * inv.y; r1y = ny.w * inv.y; r0z = nz.z * inv.z; r1z = nz.w * inv.z;
            } else {
                l0x = PT_SLAB(nx.x, o.x, inv.x, noi.x); l1x = PT_SLAB(nx.y, o.x, inv.x, noi.x);
                l0y = PT_SLAB(ny.x, o.y, inv.y, noi.y); l1y = PT_SLAB(ny.y, o.y, inv.y, noi.y);
                l0z = PT_SLAB(nz.x, o.z, inv.z, noi.z); l1z = PT_SLAB(nz.y, o.z, inv.z, noi.z);
                r0x = PT_SLAB(nx.z, o.x, inv.x, noi.x); r1x = PT_SLAB(nx.w, o.x, inv.x, noi.x);
                r0y = PT_SLAB(ny.z, o.y, inv.y, noi.y); r1y = PT_SLAB(ny.w, o.y, inv.y, noi.y);
                r0z = PT_SLAB(nz.z, o.z, inv.z, noi.z); r1z = PT_SLAB(nz.w, o.z, inv.z, noi.z);
            }
#endif
            float ltmin, ltmax, rtmin, rtmax;
            if (OCT) {      // slot 0 = the plane the ray meets first (trav_root)
                ltmin = fmax3(l0x, l0y, fmaxf(l0z, t_min)); ltmax = fmin3(l1x, l1y, fminf(l1z, best_t));
                rtmin = fmax3(r0x, r0y, fmaxf(r0z, t_min)); rtmax = fmin3(r1x, r1y, fminf(r1z, best_t));
            } else {
                ltmin = fmax3(fminf(l0x, l1x), fminf(l0y, l1y), fmaxf(fminf(l0z, l1z), t_min));
                ltmax = fmin3(fmaxf(l0x, l1x), fmaxf(l0y, l1y), fminf(fmaxf(l0z, l1z), best_t));
                rtmin = fmax3(fminf(r0x, r1x), fminf(r0y, r1y), fmaxf(fminf(r0z, r1z), t_min));
                rtmax = fmin3(fmaxf(r0x, r1x), fmaxf(r0y, r1y), fminf(fmaxf(r0z, r1z), best_t));
            }
            const int lref = __float_as_int(nc.x), rref = __float_as_int(nc.y);
            const bool hl = PT_BOX_HIT(ltmax, ltmin, slack);        // both children always exist (host packer)
            const bool hr = PT_BOX_HIT(rtmax, rtmin, slack);
            if (hl && hr) {
                const bool left_first = ltmin <= rtmin;
                trav_push(T, stack, stack_pack(left_first ? rref : lref, left_first ? rtmin : ltmin));
                T.cur = left_first ? lref : rref;
            } else if (hl) {
                T.cur = lref;
            } else if (hr) {
                T.cur = rref;
            } else {
                T.cur = trav_pop(T, stack);
            }
        }
        if (T.cur != kRefNone) {   // a leaf reference
            const int type = ref_type(T.cur), idx = ref_index(T.cur);
            float t;
            bool hit;
            if (type == PT_PRIM_SPHERE) { if (COUNT) lc.sphere_tests++; hit = hit_sphere_t<FULL>(S, idx, T.o, T.d, t_min, T.best_t, t, T.time); }
            else if (type == PT_PRIM_QUAD) { if (COUNT) lc.quad_tests++; hit = hit_quad_t(S, idx, T.o, T.d, t_min, T.best_t, t); }
            else { if (COUNT) lc.triangle_tests++; hit = hit_triangle_t(S, idx, T.o, T.d, t_min, T.best_t, t); }
            // the primitive tests accept t <= best_t; exactly equal: trav_take_tie (kRefNone stays kRefNone).  Two separate
            // predicated statements: as if / else the compiler routes best_t and best_ref through copies on every path
            if (PT_TIE && hit && t == T.best_t) T.best_ref &= ~kRefTie;
            if (hit && t < T.best_t) { T.best_t = t; T.best_ref = T.cur; }   // kernels.py:691 (strict)
            T.cur = trav_pop(T, stack);
        }
        if (min_active > 0 && __popc(__activemask()) < min_active) break;
    }
    return T.cur == kRefNone;
}

// Warp-voting variant of trav_run for persistent kernels.  ALL 32 lanes of the warp call it together
// (`active` = this lane is traversing).  Each iteration the warp executes exactly one of the two
// code blocks -- the node visit or the primitive test -- chosen by majority among the lanes that
// still traverse, instead of while-while's "descend until every lane holds a leaf", where one slow
// lane keeps 31 waiting (measured: 10.7 of 32 lanes active in the node-visit code; a trace-driven
// model of this policy predicts ~25).  Returns when fewer than min_active lanes still traverse.
// `done` is set for a lane when its traversal has finished.  Visiting order per lane is unchanged,
// so results are bit-identical to trav_run.
template <bool COUNT, bool FULL = false, class STK = LocalStack>
__device__ __forceinline__ void trav_run_vote(TravState& T, bool active, bool& done, const DScene& S,
                                              SNodeBase s_nodes, int n_snodes, const STK& stack,
                                              int min_active, int vote_num, LocalCounters& lc) {
    const f3 o = T.o, inv = T.inv;
#if PT_FAST_SLAB
    const f3 noi = T.noi;
    const float slack = T.slack;
#else
    const f3 noi = mk3(0.0f, 0.0f, 0.0f);
    const float slack = 0.0f;
#endif
    const float t_min = T.t_min;
    const float4* const nodes = S.nodes;
    // lanes that do not traverse look finished: one register carries "active and where"
    int cur = active ? T.cur : kRefNone;
    bool progressed = false;     // at least one block is executed per call (no livelock whatever min_active is)
    for (;;) {
        // common case first: enough lanes want to descend -> one ballot, no further bookkeeping
        const bool want_node = cur >= 0;
        const unsigned m_node = __ballot_sync(0xffffffffu, want_node);
        bool do_node = __popc(m_node) >= vote_num && m_node != 0u;
        if (!do_node) {
            const bool want_leaf = (unsigned)cur > (unsigned)kRefNone;      // negative, not kRefNone
            const unsigned m_leaf = __ballot_sync(0xffffffffu, want_leaf);
            const int n_all = __popc(m_node | m_leaf);
            if (n_all == 0 || (progressed && n_all < min_active)) break;
            do_node = (m_leaf == 0u);
        }
        progressed = true;
        if (do_node) {
#if PT_WIDE
            if (want_node) {
                if (COUNT) lc.box_tests += 4;
                cur = visit_wide<false>(T, nodes, cur, o, inv, noi, slack, t_min, stack, S.num_inner);
            }
#else
            if (want_node) {
                float4 nx, ny, nz, nc;
                load_node(nodes, s_nodes, n_snodes, cur, nx, ny, nz, nc, S.num_inner + 7 * S.oct_stride);
                if (COUNT) lc.box_tests += 2;
                const float best_t = T.best_t;
#if PT_F32X2
                float l0x, l1x, l0y, l1y, l0z, l1z, r0x, r1x, r0y, r1y, r0z, r1z;
                slab2(nx.x, nx.y, o.x, inv.x, noi.x, l0x, l1x); slab2(ny.x, ny.y, o.y, inv.y, noi.y, l0y, l1y); slab2(nz.x, nz.y, o.z, inv.z, noi.z, l0z, l1z);
                slab2(nx.z, nx.w, o.x, inv.x, noi.x, r0x, r1x); slab2(ny.z, ny.w, o.y, inv.y, noi.y, r0y, r1y); slab2(nz.z, nz.w, o.z, inv.z, noi.z, r0z, r1z);
#else
                float l0x = PT_SLAB(nx.x, o.x, inv.x, noi.x), l1x = PT_SLAB(nx.y, o.x, inv.x, noi.x);
                float l0y = PT_SLAB(ny.x, o.y, inv.y, noi.y), l1y = PT_SLAB(ny.y, o.y, inv.y, noi.y);
                float l0z = PT_SLAB(nz.x, o.z, inv.z, noi.z), l1z = PT_SLAB(nz.y, o.z, inv.z, noi.z);
                float r0x = PT_SLAB(nx.z, o.x, inv.x, noi.x), r1x = PT_SLAB(nx.w, o.x, inv.x, noi.x);
                float r0y = PT_SLAB(ny.z, o.y, inv.y, noi.y), r1y = PT_SLAB(ny.w, o.y, inv.y, noi.y);
                float r0z = PT_SLAB(nz.z, o.z, inv.z, noi.z), r1z = PT_SLAB(nz.w, o.z, inv.z, noi.z);
#endif
                const float ltmin = fmax3(fminf(l0x, l1x), fminf(l0y, l1y), fmaxf(fminf(l0z, l1z), t_min));
                const float ltmax = fmin3(fmaxf(l0x, l1x), fmaxf(l0y, l1y), fminf(fmaxf(l0z, l1z), best_t));
                const float rtmin = fmax3(fminf(r0x, r1x), fminf(r0y, r1y), fmaxf(fminf(r0z, r1z), t_min));
                const float rtmax = fmin3(fmaxf(r0x, r1x), fmaxf(r0y, r1y), fminf(fmaxf(r0z, r1z), best_t));
                const int lref = __float_as_int(nc.x), rref = __float_as_int(nc.y);
                const bool hl = PT_BOX_HIT(ltmax, ltmin, slack);        // both children always exist (host packer)
                const bool hr = PT_BOX_HIT(rtmax, rtmin, slack);
                if (hl && hr) {
                    const bool left_first = ltmin <= rtmin;
                    trav_push(T, stack, stack_pack(left_first ? rref : lref, left_first ? rtmin : ltmin));
                    cur = left_first ? lref : rref;
                } else if (hl) {
                    cur = lref;
                } else if (hr) {
                    cur = rref;
                } else {
                    cur = trav_pop(T, stack);
                }
            }
#endif
        } else {
            if ((unsigned)cur > (unsigned)kRefNone) {
                const int type = ref_type(cur), idx = ref_index(cur);
                float t;
                bool hit;
                if (type == PT_PRIM_SPHERE) { if (COUNT) lc.sphere_tests++; hit = hit_sphere_t<FULL>(S, idx, T.o, T.d, t_min, T.best_t, t, T.time); }
                else if (type == PT_PRIM_QUAD) { if (COUNT) lc.quad_tests++; hit = hit_quad_t(S, idx, T.o, T.d, t_min, T.best_t, t); }
                else { if (COUNT) lc.triangle_tests++; hit = hit_triangle_t(S, idx, T.o, T.d, t_min, T.best_t, t); }
                // the primitive tests accept t <= best_t; exactly equal: trav_take_tie (kRefNone stays kRefNone)
                if (PT_TIE && hit && t == T.best_t) T.best_ref &= ~kRefTie;
                if (hit && t < T.best_t) { T.best_t = t; T.best_ref = cur; }   // kernels.py:691 (strict)
                cur = trav_pop(T, stack);
            }
        }
    }
    if (active) T.cur = cur;
    done = active && cur == kRefNone;
}

// ---- equal-t hits.  The reference keeps the FIRST of several primitives hit at exactly the same t (strict
// `t < closest_t`, kernels.py:691), and it meets them in the order of ITS traversal: at every inner node the child whose
// box centre projects nearer along the ray first (kernels.py:703-732).  The traversals above meet them in another order
// (nearest slab entry first, warp votes), so they only note that it happened -- bit kRefTie of best_ref cleared; a closer
// hit sets a fresh reference and with it the bit again, which matters: 0.8 % of config 2's rays meet two coincident box
// walls BEHIND the surface they finally stop at, 0.02 % end on a shared hit.  For those, resolve_tie walks the tree the
// reference's way and returns the first primitive it meets that is hit at exactly t.  Pruning nodes never changes the
// relative order in which a depth-first walk meets the leaves that remain, so the walk only follows boxes the ray is
// inside of at the hit itself: slab interval meeting [t (1 - 2^-7), t], a margin five orders of magnitude above the
// rounding of a slab distance -- a root-to-leaf path or two.
// Where it runs: measured on config 2, ANY resolve code inside k_wf_extend -- a call or an inlined loop, executed for
// one ray in 5000 -- slows that kernel by 10 % (register allocation / scheduling of the voting loop around it), while
// noting costs under 1 %.  So the two kernels that carry the frame defer: k_wf_extend lists the ray (WfQueues.tie_list)
// for k_wf_fix_ties, the fused primary wave lists the path for a second, small launch of its resolving instantiation.
// Whole cost on config 2: -1.5 % (note 0.8, the extra launches 0.5, list appends 0.2); megakernel -2.7 %.
// The kernels that shade in place and are not on the critical path of the frame (k_megakernel, the tails, probes, the
// counting instantiations) call trav_finish.  Coincident faces of neighbouring boxes and a medium standing on a floor
// are everyday modelling, and when the tied primitives have different materials the choice shows in the image.
template <bool FULL>
__device__ __noinline__ int resolve_tie(const DScene& S, f3 o, f3 d, f3 inv, float t_min, float t, float time) {
#if PT_WIDE
    return kRefNone;
#else
    int stack[kStackSize];
    int sp = 0;
    int cur = S.num_inner > 0 ? 0 : kRefNone;
    const float t_lo = fmaxf(t_min, t * 0.9921875f);
    while (cur != kRefNone) {
        if (cur >= 0) {
            PT_BOUND(cur < S.num_inner, PT_DBG_NODE);
            const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(S.nodes_plain) + (unsigned long long)(unsigned)cur * 64ull);
            const float4 nx = __ldg(p), ny = __ldg(p + 1), nz = __ldg(p + 2);
            const float2 c2 = __ldg(reinterpret_cast<const float2*>(p + 3));
            const int lref = __float_as_int(c2.x), rref = __float_as_int(c2.y);
            const float l0x = (nx.x - o.x) * inv.x, l1x = (nx.y - o.x) * inv.x, l0y = (ny.x - o.y) * inv.y, l1y = (ny.y - o.y) * inv.y;
            const float l0z = (nz.x - o.z) * inv.z, l1z = (nz.y - o.z) * inv.z;
            const float r0x = (nx.z - o.x) * inv.x, r1x = (nx.w - o.x) * inv.x, r0y = (ny.z - o.y) * inv.y, r1y = (ny.w - o.y) * inv.y;
            const float r0z = (nz.z - o.z) * inv.z, r1z = (nz.w - o.z) * inv.z;
            const bool hl = fmin3(fmaxf(l0x, l1x), fmaxf(l0y, l1y), fminf(fmaxf(l0z, l1z), t)) >=
                            fmax3(fminf(l0x, l1x), fminf(l0y, l1y), fmaxf(fminf(l0z, l1z), t_lo));
            const bool hr = fmin3(fmaxf(r0x, r1x), fmaxf(r0y, r1y), fminf(fmaxf(r0z, r1z), t)) >=
                            fmax3(fminf(r0x, r1x), fminf(r0y, r1y), fmaxf(fminf(r0z, r1z), t_lo));
            const f3 lc = mk3((nx.x + nx.y) * 0.5f, (ny.x + ny.y) * 0.5f, (nz.x + nz.y) * 0.5f);
            const f3 rc = mk3((nx.z + nx.w) * 0.5f, (ny.z + ny.w) * 0.5f, (nz.z + nz.w) * 0.5f);
            const bool left_first = dot(lc - o, d) < dot(rc - o, d);        // kernels.py:716-721
            const int first = left_first ? lref : rref, second = left_first ? rref : lref;
            const bool hf = left_first ? hl : hr, hs = left_first ? hr : hl;
            if (hf && hs) { if (sp < kStackSize) stack[sp++] = second; cur = first; }
            else if (hf) cur = first;
            else if (hs) cur = second;
            else cur = sp > 0 ? stack[--sp] : kRefNone;
        } else {
            const int type = ref_type(cur), idx = ref_index(cur);
            float tt = 0.0f;
            bool hit;
            if (type == PT_PRIM_SPHERE) hit = hit_sphere_t<FULL>(S, idx, o, d, t_min, t, tt, time);
            else if (type == PT_PRIM_QUAD) hit = hit_quad_t(S, idx, o, d, t_min, t, tt);
            else hit = hit_triangle_t(S, idx, o, d, t_min, t, tt);
            if (hit && tt == t) return cur;
            cur = sp > 0 ? stack[--sp] : kRefNone;
        }
    }
    return kRefNone;
#endif
}
// the tie flag of a finished closest-hit traversal; leaves best_ref a plain reference again
__device__ __forceinline__ bool trav_take_tie(TravState& T) {
    const bool tie = PT_TIE && !(T.best_ref & kRefTie) && T.best_ref != kRefNone;
    if (tie) T.best_ref |= kRefTie;
    return tie;
}
// call when a closest-hit traversal has finished and the primitive id matters (resolves in place)
template <bool FULL>
__device__ __forceinline__ void trav_finish(TravState& T, const DScene& S) {
    if (trav_take_tie(T)) {
        const int r = resolve_tie<FULL>(S, T.o, T.d, T.inv, T.t_min, T.best_t, T.time);
        if (r != kRefNone) T.best_ref = r;
    }
}

// One-shot traversal (probe / wavefront kernels).
template <bool COUNT, bool FULL = false>
__device__ __forceinline__ void traverse(const DScene& S, SNodeBase s_nodes, int n_snodes, f3 o, f3 d,
                                         float t_min, float t_max, float& best_t, int& best_ref, LocalCounters& lc,
                                         float time = 0.0f) {
    StackEntry stack_mem[kStackSlots];
    const LocalStack stack{stack_mem};
    stack_init(stack);
    TravState T;
    trav_begin<COUNT>(T, S, o, d, t_min, t_max, lc, time);
    trav_run<COUNT, FULL>(T, S, s_nodes, n_snodes, stack, 0, lc);
    trav_finish<FULL>(T, S);
    best_t = T.best_t;
    best_ref = T.best_ref;
}

// ---- samplers (kernels.py:16-71)
__device__ __forceinline__ f3 random_in_unit_disk(ptm::Rng& rng) {
    for (;;) {
        float px = ptm::rng_next(rng) * 2.0f - 1.0f;
        float py = ptm::rng_next(rng) * 2.0f - 1.0f;
        f3 p = mk3(px, py, 0.0f);
        if (dot(p, p) < 1.0f) return p;
    }
}
__device__ __forceinline__ f3 random_unit_vector(ptm::Rng& rng) {
    for (;;) {
        float px = ptm::rng_next(rng) * 2.0f - 1.0f;
        float py = ptm::rng_next(rng) * 2.0f - 1.0f;
        float pz = ptm::rng_next(rng) * 2.0f - 1.0f;
        f3 p = mk3(px, py, pz);
        float l = dot(p, p);
        if (l < 1.0f && l > 1e-20f) return normalized(p);
    }
}
__device__ __forceinline__ f3 random_cosine_direction(f3 normal, ptm::Rng& rng) {
    float r1 = ptm::rng_next(rng);
    float r2 = ptm::rng_next(rng);
    float z = ptm::fsqrt(1.0f - r2);
    float phi = 6.283185307179586f * r1;
    float st = ptm::fsqrt(r2);
    float sn, cs;
    ptm::sincos_c(phi, &sn, &cs);
    float x = cs * st, y = sn * st;
    f3 w = normalized(normal);
    f3 a;
    if (fabsf(w.x) < 0.9f) a = mk3(1.0f, 0.0f, 0.0f);
    else if (fabsf(w.y) < 0.9f) a = mk3(0.0f, 1.0f, 0.0f);
    else a = mk3(0.0f, 0.0f, 1.0f);
    f3 v = normalized(cross(a, w));
    f3 u = cross(v, w);
    return normalized((x * u + y * v) + z * w);
}

// ---- camera (kernels.py:176-201)
__device__ __forceinline__ void get_ray(const PtCamera& cam, int px, int py, ptm::Rng& rng, f3& o, f3& d,
                                        uint32_t seed, uint32_t pixel, uint32_t sample) {
    float ox, oy;
    if (cam.jitter == 1) {                      // low-discrepancy sub-pixel offsets; the random stream is left for the bounces
        ptm::ld_jitter(seed, pixel, sample, ox, oy);
        ox = ox - 0.5f;
        oy = oy - 0.5f;
    } else {
        ox = ptm::rng_next(rng) - 0.5f;
        oy = ptm::rng_next(rng) - 0.5f;
    }
    f3 p00 = mk3(cam.pixel00[0], cam.pixel00[1], cam.pixel00[2]);
    f3 du = mk3(cam.delta_u[0], cam.delta_u[1], cam.delta_u[2]);
    f3 dv = mk3(cam.delta_v[0], cam.delta_v[1], cam.delta_v[2]);
    f3 c = mk3(cam.center[0], cam.center[1], cam.center[2]);
    f3 ps = (p00 + ((float)px + ox) * du) + ((float)py + oy) * dv;
    o = c;
    if (cam.defocus_angle > 0.0f) {
        f3 p = random_in_unit_disk(rng);
        f3 ddu = mk3(cam.defocus_disk_u[0], cam.defocus_disk_u[1], cam.defocus_disk_u[2]);
        f3 ddv = mk3(cam.defocus_disk_v[0], cam.defocus_disk_v[1], cam.defocus_disk_v[2]);
        o = (c + p.x * ddu) + p.y * ddv;
    }
    d = ps - o;
}

// ---- textures (kernels.py:78-169, :924-1017)
__device__ __forceinline__ float perlin_noise(const DScene& S, f3 p) {
    float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    float u = p.x - fx, v = p.y - fy, w = p.z - fz;
    int i = (int)fx, j = (int)fy, k = (int)fz;
    float uu = u * u * (3.0f - 2.0f * u);
    float vv = v * v * (3.0f - 2.0f * v);
    float ww = w * w * (3.0f - 2.0f * w);
    float accum = 0.0f;
#pragma unroll
    for (int di = 0; di < 2; di++)
#pragma unroll
        for (int dj = 0; dj < 2; dj++)
#pragma unroll
            for (int dk = 0; dk < 2; dk++) {
                int idx = S.perlin_perm[(i + di) & 255] ^ S.perlin_perm[256 + ((j + dj) & 255)] ^
                          S.perlin_perm[512 + ((k + dk) & 255)];
                float4 g = __ldg(S.perlin_vec + idx);
                float fi = (float)di, fj = (float)dj, fk = (float)dk;
                f3 wt = mk3(u - fi, v - fj, w - fk);
                accum = accum + (fi * uu + (1.0f - fi) * (1.0f - uu)) * (fj * vv + (1.0f - fj) * (1.0f - vv)) *
                                    (fk * ww + (1.0f - fk) * (1.0f - ww)) * dot(mk3(g.x, g.y, g.z), wt);
            }
    return accum;
}
__device__ __forceinline__ float perlin_turb3(const DScene& S, f3 p) {
    float accum = 0.0f, weight = 1.0f;
    f3 tp = p;
    for (int i = 0; i < 3; i++) {            // depth = 3, kernels.py:1013
        accum = accum + weight * perlin_noise(S, tp);
        weight = weight * 0.5f;
        tp = tp * 2.0f;
    }
    return fabsf(accum);
}

// material record accessors
struct Mat {
    int header;     // mat_type | tex_type<<4 | is_medium<<8 | (image_idx+1)<<16
    float scalar;   // Lambert/isotropic: texture scale; metal: fuzz; dielectric: ir
    f3 c1, c2;      // Lambert: texture colours; metal: albedo; emissive: emit colour
};
__device__ __forceinline__ int mat_id_of(const DScene& S, int ref) {
    PT_BOUND(ref < 0 && ref != kRefNone && ref_type(ref) <= 2 && ref_index(ref) < S.n_prims[ref_type(ref)], PT_DBG_PRIM);
    return (int)__ldg(S.prim_mat + S.mat_base[ref_type(ref)] + ref_index(ref));
}
__device__ __forceinline__ Mat load_mat(const DScene& S, int mid) {
    PT_BOUND(mid >= 0 && mid < S.num_materials, PT_DBG_MATERIAL);
    float4 m0 = __ldg(S.materials + 3 * mid), m1 = __ldg(S.materials + 3 * mid + 1);
    Mat m;
    m.header = __float_as_int(m0.x);
    m.scalar = m0.y;
    m.c1 = mk3(m0.z, m0.w, m1.x);
    m.c2 = mk3(m1.y, m1.z, m1.w);
    return m;
}
__device__ __forceinline__ int mat_type(int header) { return header & 15; }
__device__ __forceinline__ int tex_type(int header) { return (header >> 4) & 15; }
__device__ __forceinline__ bool mat_is_medium(int header) { return (header >> 8) & 1; }
__device__ __forceinline__ int mat_image(int header) { return (header >> 16) - 1; }

// kernels.py:924-1017
__device__ __forceinline__ f3 eval_texture(const DScene& S, const Mat& m, int ref, f3 hp) {
    int tt = tex_type(m.header);
    f3 result = mk3(1.0f, 1.0f, 1.0f);
    if (tt == 0) {
        result = m.c1;
    } else if (tt == 1) {
        float inv_scale = ptm::frcp(m.scalar);
        int s = (int)floorf(inv_scale * hp.x) + (int)floorf(inv_scale * hp.y) + (int)floorf(inv_scale * hp.z);
        result = (s & 1) == 0 ? m.c1 : m.c2;     // Python-style (s % 2 == 0) for negative s too
    } else if (tt == 2) {
        if (ref_type(ref) == PT_PRIM_SPHERE) {
            float4 sp = __ldg(S.spheres + ref_index(ref));
            f3 on = normalized(hp - mk3(sp.x, sp.y, sp.z));                 // get_sphere_uv, :78-102
            float phi = ptm::acos_c(-on.y);
            float theta = ptm::atan2_c(-on.z, on.x) + 3.141592653589793f;
            float u = ptm::fdiv(theta, 6.283185307179586f);
            float v = ptm::fdiv(phi, 3.141592653589793f);
            int img = mat_image(m.header);
            if (img >= 0) {
                int iw = S.image_w[img], ih = S.image_h[img];
                u = max2(0.0f, min2(1.0f, u));
                v = 1.0f - max2(0.0f, min2(1.0f, v));
                int i = (int)(u * (float)iw), j = (int)(v * (float)ih);
                i = max(0, min(iw - 1, i));
                j = max(0, min(ih - 1, j));
                PT_BOUND(img < PT_MAX_IMAGES && S.image[img] != nullptr && i >= 0 && i < iw && j >= 0 && j < ih, PT_DBG_TEXEL);
                uchar4 t = __ldg(S.image[img] + (size_t)j * iw + i);
                result = mk3(ptm::fdiv((float)t.x, 255.0f), ptm::fdiv((float)t.y, 255.0f), ptm::fdiv((float)t.z, 255.0f));
            }
        } else {
            result = mk3(1.0f, 0.0f, 1.0f);
        }
    } else if (tt == 3) {
        float nv = ptm::sin_c(m.scalar * hp.z + 10.0f * perlin_turb3(S, hp));
        result = (m.c1 * 0.5f) * (1.0f + nv);
    }
    return result;
}

__device__ __forceinline__ f3 reflect(f3 v, f3 n) { return v - (2.0f * dot(v, n)) * n; }
__device__ __forceinline__ f3 refract(f3 uv, f3 n, float eta) {
    float ct = min2(-dot(uv, n), 1.0f);
    f3 perp = eta * (uv + ct * n);
    f3 par = (-ptm::fsqrt(fabsf(1.0f - dot(perp, perp)))) * n;
    return perp + par;
}
__device__ __forceinline__ float reflectance(float cosine, float ref_idx) {
    float r0 = ptm::fdiv(1.0f - ref_idx, 1.0f + ref_idx);
    r0 = r0 * r0;
    return r0 + (1.0f - r0) * ptm::pow5_c(1.0f - cosine);
}

// core/material.py:258-276 _sample_henyey_greenstein (full feature mode)
__device__ __noinline__ f3 sample_henyey_greenstein(f3 incident, float g, ptm::Rng& rng) {
    if (fabsf(g) < 1e-3f) return random_unit_vector(rng);
    float sqr_term = ptm::fdiv(1.0f - g * g, 1.0f - g + 2.0f * g * ptm::rng_next(rng));
    float cos_theta = ptm::fdiv(1.0f + g * g - sqr_term * sqr_term, 2.0f * g);
    float sin_theta = ptm::fsqrt(max2(0.0f, 1.0f - cos_theta * cos_theta));
    float phi = 6.2831853f * ptm::rng_next(rng);
    f3 w = normalized(incident);
    f3 a = fabsf(w.x) > 0.9f ? mk3(1.0f, 0.0f, 0.0f) : mk3(0.0f, 1.0f, 0.0f);
    f3 v = normalized(cross(w, a));
    f3 u = cross(w, v);
    float sn, cs;
    ptm::sincos_c(phi, &sn, &cs);
    return normalized(((u * cs) * sin_theta + (v * sn) * sin_theta) + w * cos_theta);
}

// subsurface_simple / subsurface_volumetric (core/material.py:144-256), full feature mode only.
// `hp` is replaced by the (possibly displaced) origin of the scattered ray.
__device__ __noinline__ bool shade_subsurface(const Mat& m, f3 ray_dir, f3& hp, f3 normal, ptm::Rng& rng,
                                              f3& scatter_dir, f3& attenuation) {
    const f3 nf = dot(ray_dir, normal) < 0.0f ? normal : -normal;     // face-forwarded, core/hittable.py:17-22
    const f3 albedo = m.c1;
    if (mat_type(m.header) == 5) {
        const float scatter_distance = m.scalar;
        f3 dir;
        if (ptm::rng_next(rng) < 0.5f) {
            dir = nf + random_unit_vector(rng);
        } else {
            f3 ruv = random_unit_vector(rng);
            f3 displacement = (ruv * scatter_distance) * ptm::rng_next(rng);
            hp = hp + displacement;
            dir = nf + random_unit_vector(rng);
        }
        if (fabsf(dir.x) < 1e-8f && fabsf(dir.y) < 1e-8f && fabsf(dir.z) < 1e-8f) dir = nf;
        scatter_dir = dir;
        attenuation = albedo;
        return true;
    }
    const float sigma_s = m.scalar, sigma_a = m.c2.x, g = m.c2.y;
    const float sigma_t = sigma_s + sigma_a;
    const f3 entry = hp;
    f3 pos = hp - nf * 0.001f;
    f3 dirn = normalized(ray_dir);
    f3 thr = mk3(1.0f, 1.0f, 1.0f);
    for (int it = 0; it < 64; it++) {
        float t = ptm::fdiv(-ptm::log_c(max2(ptm::rng_next(rng), 1e-10f)), sigma_t);
        pos = pos + dirn * t;
        f3 dv = pos - entry;
        float dist = ptm::fsqrt(dot(dv, dv));
        float exit_probability = 1.0f - ptm::exp_c(-dist * 0.5f);
        if (ptm::rng_next(rng) < exit_probability) {
            f3 dir = nf + random_unit_vector(rng);
            if (fabsf(dir.x) < 1e-8f && fabsf(dir.y) < 1e-8f && fabsf(dir.z) < 1e-8f) dir = nf;
            hp = pos;
            scatter_dir = dir;
            attenuation = thr * albedo;
            return true;
        }
        if (ptm::rng_next(rng) < ptm::fdiv(sigma_a, sigma_t)) { attenuation = mk3(0.0f, 0.0f, 0.0f); return false; }
        dirn = sample_henyey_greenstein(dirn, g, rng);
        thr = thr * albedo;
    }
    return false;
}

// emitted + scatter for a surface hit (kernels.py:789-917). Returns scattered.  `hp` is in/out: the
// scattered ray starts at the hit point, except for the full-mode subsurface materials.
template <bool FULL = false>
__device__ __forceinline__ bool shade_surface(const DScene& S, const Mat& m, int ref, f3 ray_dir, f3& hp, f3 normal,
                                              ptm::Rng& rng, f3& emit, f3& scatter_dir, f3& attenuation) {
    int mt = mat_type(m.header);
    emit = mk3(0.0f, 0.0f, 0.0f);
    scatter_dir = mk3(0.0f, 0.0f, 0.0f);
    attenuation = mk3(1.0f, 1.0f, 1.0f);
    if (mt >= 5) return FULL ? shade_subsurface(m, ray_dir, hp, normal, rng, scatter_dir, attenuation) : false;
    if (mt == 0) {
        attenuation = eval_texture(S, m, ref, hp);
        scatter_dir = random_cosine_direction(normal, rng);
        return true;
    } else if (mt == 1) {
        f3 reflected = reflect(normalized(ray_dir), normal);
        scatter_dir = reflected + m.scalar * random_unit_vector(rng);
        if (dot(scatter_dir, normal) > 0.0f) { attenuation = m.c1; return true; }
        return false;
    } else if (mt == 2) {
        float ir = m.scalar;
        bool front = dot(ray_dir, normal) < 0.0f;
        f3 nf = front ? normal : -normal;
        float ratio = front ? ptm::frcp(ir) : ir;
        f3 ud = normalized(ray_dir);
        float ct = min2(-dot(ud, nf), 1.0f);
        float st = ptm::fsqrt(1.0f - ct * ct);
        bool cannot = ratio * st > 1.0f;
        bool refl = cannot;
        if (!refl) refl = reflectance(ct, ratio) > ptm::rng_next(rng);   // `or` short-circuits, :894
        scatter_dir = refl ? reflect(ud, nf) : refract(ud, nf, ratio);
        return true;
    } else if (mt == 3) {
        emit = m.c1;
        return false;
    } else if (mt == 4) {
        scatter_dir = random_unit_vector(rng);
        attenuation = eval_texture(S, m, ref, hp);
        return true;
    }
    return false;
}

// ---- next-event estimation (features="nee"; only in the FULL kernel instantiations).
// At a Lambertian hit one emitter is chosen uniformly, a point on it is sampled uniformly by area, and its radiance is
// added if the shadow ray reaches it:  L = albedo/pi * Le * cos_s * cos_l / dist^2 * (area * n_lights) * T.
// cos_s is taken against the SAME normal the cosine-weighted scatter uses (spheres: outward, never flipped,
// kernels.py:246), emitters emit from both faces (kernels.py:904-906).  The shadow ray follows the reference's
// constant-medium semantics so that both estimators have the same expectation: a medium boundary is not an occluder
// but a transmittance exp(-density * length inside), "inside" ends at the next hit of ANY primitive
// (kernels.py:417-419), and the ray then continues beyond that hit (kernels.py:1100-1110); an exit search that finds
// nothing makes the boundary a surface (kernels.py:1111-1119), i.e. opaque here.
// Area sampling is a poor strategy for an emitter that is close (a ceiling point 1 unit above the Cornell light sees
// area * n / (pi * dist^2) ~ 4000): the two strategies that can produce a (vertex, emitter point) pair -- this one and
// the cosine-sampled next ray hitting the emitter -- are therefore combined with a partition of unity on their distance:
// pairs farther apart than sqrt(nee_r2) (= half the side of the mean emitter) are taken from here, closer ones from the
// scattered ray (the caller keeps a per-path flag and counts a directly hit emitter after such a vertex only when it is
// closer than that).  Pairs whose connecting segment crosses a medium interval always belong to this strategy (after a
// pass-through the scattered ray restarts from a new origin and no longer knows its distance to the vertex).  Every
// pair is counted exactly once either way, so the estimator stays unbiased.  Path state (2 bits): 0 = the last
// scatter had no next-event estimation, 1 = it had, 2 = it had and the ray has since passed through a medium.
template <bool FULL>
__device__ f3 nee_direct(const DScene& S, f3 hp, f3 n_s, f3 albedo, ptm::Rng& rng, float time) {
    const f3 zero = mk3(0.0f, 0.0f, 0.0f);
    int li = (int)(ptm::rng_next(rng) * (float)S.n_lights);
    li = li < S.n_lights ? li : S.n_lights - 1;
    const float4 L0 = __ldg(S.lights + 2 * li), L1 = __ldg(S.lights + 2 * li + 1);
    const int lref = __float_as_int(L0.x);
    const int type = ref_type(lref), idx = ref_index(lref);
    f3 y, nl;
    if (type == PT_PRIM_SPHERE) {
        const float4 sp = __ldg(S.spheres + idx);
        const f3 dirn = random_unit_vector(rng);
        y = sphere_center<FULL>(S, idx, sp, time) + sp.w * dirn;
        nl = dirn;
    } else if (type == PT_PRIM_QUAD) {
        const float4* p = S.quads + 4 * idx;
        const float4 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3);
        const float a = ptm::rng_next(rng), b = ptm::rng_next(rng);
        y = (mk3(q1.x, q1.y, q1.z) + a * mk3(q2.x, q2.y, q2.z)) + b * mk3(q3.x, q3.y, q3.z);
        nl = mk3(q0.x, q0.y, q0.z);
    } else {
        const float4* p = S.tris + 3 * idx;
        const float4 a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2);
        float r1 = ptm::rng_next(rng), r2 = ptm::rng_next(rng);
        if (r1 + r2 > 1.0f) { r1 = 1.0f - r1; r2 = 1.0f - r2; }
        y = (mk3(a0.x, a0.y, a0.z) + r1 * mk3(a1.x, a1.y, a1.z)) + r2 * mk3(a2.x, a2.y, a2.z);
        nl = mk3(a0.w, a1.w, a2.w);
    }
    const f3 dv = y - hp;
    const float dist2 = dot(dv, dv);
    if (!(dist2 > 1e-12f)) return zero;
    const bool close = !(dist2 >= S.nee_r2);                  // close pairs belong to the scattered ray, unless a medium lies between
    bool crossed = false;
    const float dist = ptm::fsqrt(dist2);
    const f3 wi = vdiv(dv, dist);
    const float cos_s = dot(n_s, wi), cos_l = fabsf(dot(nl, wi));
    if (!(cos_s > 0.0f) || !(cos_l > 0.0f)) return zero;
    const float t_end = dist * 0.9999f;                       // stop short of the emitter itself
    const float len = ptm::fsqrt(dot(wi, wi));
    float trans = 1.0f, t_cur = 0.001f;
    LocalCounters lc;
    for (int seg = 0; seg < 16; seg++) {
        if (!(t_cur < t_end)) break;
        float t;
        int ref;
        traverse<false, FULL>(S, 0ull, 0, hp, wi, t_cur, t_end, t, ref, lc, time);
        if (ref == kRefNone) break;                           // nothing in the way
        if (!ref_is_medium(ref)) return zero;                 // occluded
        float t_exit;
        int eref;
        traverse<false, FULL>(S, 0ull, 0, hp, wi, t + 0.0001f, 1e10f, t_exit, eref, lc, time);
        if (eref == kRefNone) return zero;                    // boundary acts as a surface (fallback branch)
        crossed = true;
        const float4 m2 = __ldg(S.materials + 3 * mat_id_of(S, ref) + 2);     // (density, medium albedo)
        const float t1 = max2(t, 0.001f), t2 = min2(t_exit, t_end);
        if (t1 < t2) trans = trans * ptm::exp_c(-(m2.x * ((t2 - t1) * len)));
        // the "inside" interval reaches the emitter itself: a ray that passes such an interval continues BEYOND its closing
        // surface (kernels.py:1108), i.e. it never sees this point of the emitter (a sphere is then seen from inside, at
        // its far side -- the pair this loop reaches when the sampled point lies there)
        if (!(t_exit < t_end)) return zero;
        t_cur = t_exit + ptm::fdiv(0.001f, len);
        if (seg == 15) return zero;                           // pathological nesting: give up (unbiasedness is not at stake in practice)
    }
    if (close && !crossed) return zero;
    const float geom = ptm::fdiv(((cos_s * cos_l) * L0.y) * (float)S.n_lights, 3.14159265358979f * dist2);
    return (albedo * mk3(L0.z, L0.w, L1.x)) * (geom * trans);
}
__device__ __forceinline__ bool nee_on(const DScene& S) { return (S.full_features & 2) != 0 && S.n_lights > 0; }

}  // namespace pt
