/* TEST INFRASTRUCTURE ONLY -- the CPU parity oracle.  Not part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * A restatement, in plain scalar C++, of the reference's Taichi hot path
 *   /root/reference/src/render_server/taichi_renderer/kernels.py   (all @ti.func / @ti.kernel)
 * over the reference's own device layout
 *   /root/reference/src/render_server/taichi_renderer/fields.py:25-222
 * i.e. the arrays `compile_scene` (scene_compiler.py:931) and `compile_bvh` (bvh_compiler.py:132)
 * produce.  Control flow, constants, traversal order and evaluation order follow kernels.py
 * statement by statement; every function cites the lines it restates.  Arithmetic follows the
 * PT-F32 contract (oracle_math.h): strict IEEE binary32, no contraction (build with
 * -ffp-contract=off), contract-defined elementary functions and random stream.
 *
 * Full-feature mode (OScene.full_features): moving spheres (core/sphere.py:22-36) and the two
 * subsurface materials (core/material.py:144-276), which exist only in the reference's pure-Python
 * tracer; semantics restated here in the same f32 contract and pinned *statistically* against that
 * tracer (tests/test_full_features.py) -- the Taichi path has nothing to compare bits with.
 *
 * Pinning: tests/test_oracle_golden.py checks this file bit-for-bit against fixtures produced by
 * executing the reference's unmodified renderer.py/kernels.py/fields.py under oracle/taichi_shim
 * (a pure-Python stand-in for the absent `taichi` module); see oracle/gen_golden.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle_math.h"

extern "C" {

typedef struct {
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
} OMaterials;

typedef struct {
    int32_t num_spheres, num_quads, num_triangles, num_nodes, num_images;
    const float* sphere_data;       /* [n][4] cx cy cz r                 fields.py:25 */
    const float *quad_Q, *quad_u, *quad_v, *quad_normal, *quad_D, *quad_w;   /* fields.py:38-43 */
    const float *tri_v0, *tri_edge1, *tri_edge2, *tri_normal;                /* fields.py:29-34 */
    OMaterials mat[3];              /* indexed by prim_type: 0 sphere, 1 triangle, 2 quad */
    const float *bvh_bbox_min, *bvh_bbox_max;                                /* fields.py:52-63 */
    const int32_t *bvh_left, *bvh_right, *bvh_prim_type, *bvh_prim_idx;
    const float* perlin_randvec;    /* [256][3] */
    const int32_t *perm_x, *perm_y, *perm_z;
    const uint8_t* images[16];      /* [h][w][3] u8; texel value = u8 / 255.0f (rtw_image.py:66) */
    int32_t image_w[16], image_h[16];
    /* ---- full-feature mode (features the Taichi path dropped but the reference's pure-Python
     * tracer has; SURVEY.md "three facts" #2): moving spheres and subsurface materials */
    const float* sphere_motion;     /* [n][3] centre(1) - centre(0), or NULL (core/sphere.py:22-35) */
    int32_t full_features;
} OScene;

typedef struct {
    float center[3], pixel00[3], delta_u[3], delta_v[3], disk_u[3], disk_v[3];
    float defocus_angle;
    float bg[3];
    int32_t max_depth;
    int32_t jitter;      /* 0: random sub-pixel offsets (kernels.py:183-184); 1: scrambled Sobol' (0,2) points */
} OCamera;

typedef struct {
    uint64_t rays;          /* traverse_bvh invocations */
    uint64_t node_tests;    /* hit_aabb_optimized evaluations */
    uint64_t sphere_tests, triangle_tests, quad_tests;
    uint64_t paths;
    uint64_t bounces;       /* trace_ray loop iterations / wavefront ray-waves */
} OCounters;

} /* extern "C" */

namespace {

struct V3 {
    float x, y, z;
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 ld3(const float* p) { return V3{p[0], p[1], p[2]}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
/* Taichi Vector.dot = sum of elementwise products, accumulated left to right */
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
/* Taichi Vector.normalized(): invlen = 1 / (norm() + eps), eps = 0; return invlen * v */
inline V3 normalized(V3 a) {
    float invlen = 1.0f / (sqrtf(dot(a, a)) + 0.0f);
    return invlen * a;
}
inline float fmin2(float a, float b) { return a < b ? a : b; }   /* ti.min / ti.max on f32 */
inline float fmax2(float a, float b) { return a > b ? a : b; }

struct Rng {
    ORng s;          /* (position, stride) of the path's stream, oracle_math.h */
    float next() { return o_rng_next(&s); }
};

struct Hit {
    bool hit;
    float t;
    V3 p, n;
    int prim_type, prim_idx;
};

/* next-event estimation (full_features bit 1): one record per emissive primitive that is not a medium boundary */
struct OLight {
    int prim_type, prim_idx;
    float area;
    V3 emit;
};

struct Ctx {
    const OScene* sc;
    const OCamera* cam;
    OCounters cnt;
    float time = 0.0f;   /* shutter time of the path being traced (full-feature mode) */
    const OLight* lights = nullptr;
    int n_lights = 0;
    float nee_r2 = 0.0f;   /* squared switch-over distance between light sampling and the scattered ray (nee_direct) */
};

/* centre of sphere idx at the current path's time: Ray.at(time) = origin + time * direction
 * (core/sphere.py:36, util/ray.py:38) */
inline V3 sphere_center(const Ctx& c, int idx) {
    const float* s = c.sc->sphere_data + 4 * idx;
    V3 center = v3(s[0], s[1], s[2]);
    if (c.sc->full_features && c.sc->sphere_motion) center = center + c.time * ld3(c.sc->sphere_motion + 3 * idx);
    return center;
}

/* kernels.py:16-25 */
V3 random_in_unit_disk(Rng& rng) {
    for (;;) {
        float px = rng.next() * 2.0f - 1.0f;
        float py = rng.next() * 2.0f - 1.0f;
        V3 p = v3(px, py, 0.0f);
        if (dot(p, p) < 1.0f) return p;
    }
}

/* kernels.py:28-38 */
V3 random_unit_vector(Rng& rng) {
    for (;;) {
        float px = rng.next() * 2.0f - 1.0f;
        float py = rng.next() * 2.0f - 1.0f;
        float pz = rng.next() * 2.0f - 1.0f;
        V3 p = v3(px, py, pz);
        float lensq = dot(p, p);
        if (lensq < 1.0f && lensq > 1e-20f) return normalized(p);
    }
}

/* kernels.py:41-71 */
V3 random_cosine_direction(V3 normal, Rng& rng) {
    float r1 = rng.next();
    float r2 = rng.next();
    float z = sqrtf(1.0f - r2);
    float phi = 6.283185307179586f * r1; /* 2.0 * math.pi folded in Python, narrowed to f32 */
    float sin_theta = sqrtf(r2);
    float x = o_cosf(phi) * sin_theta;
    float y = o_sinf(phi) * sin_theta;
    V3 w = normalized(normal);
    V3 a;
    if (fabsf(w.x) < 0.9f) a = v3(1.0f, 0.0f, 0.0f);
    else if (fabsf(w.y) < 0.9f) a = v3(0.0f, 1.0f, 0.0f);
    else a = v3(0.0f, 0.0f, 1.0f);
    V3 v = normalized(cross(a, w));
    V3 u = cross(v, w);
    return normalized((x * u + y * v) + z * w);
}

/* kernels.py:78-102 */
void get_sphere_uv(V3 p, V3 center, float* u, float* v) {
    V3 on = normalized(p - center);
    float phi = o_acosf(-on.y);
    float theta = o_atan2f(-on.z, on.x) + 3.141592653589793f;
    *u = theta / 6.283185307179586f;
    *v = phi / 3.141592653589793f;
}

/* kernels.py:109-151 */
float perlin_noise(const OScene* sc, V3 p) {
    float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    float u = p.x - fx, v = p.y - fy, w = p.z - fz;
    int i = (int)fx, j = (int)fy, k = (int)fz;
    float uu = u * u * (3.0f - 2.0f * u);
    float vv = v * v * (3.0f - 2.0f * v);
    float ww = w * w * (3.0f - 2.0f * w);
    float accum = 0.0f;
    for (int di = 0; di < 2; di++)
        for (int dj = 0; dj < 2; dj++)
            for (int dk = 0; dk < 2; dk++) {
                int idx = sc->perm_x[(i + di) & 255] ^ sc->perm_y[(j + dj) & 255] ^ sc->perm_z[(k + dk) & 255];
                V3 g = ld3(sc->perlin_randvec + 3 * idx);
                float fi = (float)di, fj = (float)dj, fk = (float)dk;
                V3 weight = v3(u - fi, v - fj, w - fk);
                accum = accum + (fi * uu + (1.0f - fi) * (1.0f - uu)) *
                                    (fj * vv + (1.0f - fj) * (1.0f - vv)) *
                                    (fk * ww + (1.0f - fk) * (1.0f - ww)) * dot(g, weight);
            }
    return accum;
}

/* kernels.py:154-169 */
float perlin_turb(const OScene* sc, V3 p, int depth) {
    float accum = 0.0f;
    V3 temp_p = p;
    float weight = 1.0f;
    for (int i = 0; i < depth; i++) {
        accum = accum + weight * perlin_noise(sc, temp_p);
        weight = weight * 0.5f;
        temp_p = temp_p * 2.0f;
    }
    return fabsf(accum);
}

/* kernels.py:176-201 */
void get_ray(const OCamera* cam, int px, int py, Rng& rng, V3* o, V3* d, uint32_t seed, uint32_t pixel, uint32_t sample) {
    float offset_x, offset_y;
    if (cam->jitter == 1) {
        o_ld_jitter(seed, pixel, sample, &offset_x, &offset_y);
        offset_x = offset_x - 0.5f;
        offset_y = offset_y - 0.5f;
    } else {
        offset_x = rng.next() - 0.5f;
        offset_y = rng.next() - 0.5f;
    }
    V3 pixel_sample = (ld3(cam->pixel00) + ((float)px + offset_x) * ld3(cam->delta_u)) +
                      ((float)py + offset_y) * ld3(cam->delta_v);
    V3 origin = ld3(cam->center);
    if (cam->defocus_angle > 0.0f) {
        V3 p = random_in_unit_disk(rng);
        origin = (ld3(cam->center) + p.x * ld3(cam->disk_u)) + p.y * ld3(cam->disk_v);
    }
    *o = origin;
    *d = pixel_sample - origin;
}

/* kernels.py:208-248 */
bool hit_sphere(Ctx& c, int idx, V3 o, V3 d, float t_min, float t_max, float* t, V3* p, V3* n) {
    c.cnt.sphere_tests++;
    const float* s = c.sc->sphere_data + 4 * idx;
    V3 center = sphere_center(c, idx);
    float radius = s[3];
    V3 oc = center - o;
    float a = dot(d, d);
    float h = dot(d, oc);
    float cc = dot(oc, oc) - radius * radius;
    float disc = h * h - a * cc;
    if (disc >= 0.0f) {
        float sqrtd = sqrtf(disc);
        float root = (h - sqrtd) / a;
        if (root < t_min || root > t_max) root = (h + sqrtd) / a;
        if (root >= t_min && root <= t_max) {
            *t = root;
            *p = o + root * d;
            *n = (*p - center) / radius;
            return true;
        }
    }
    return false;
}

/* kernels.py:251-307 */
bool hit_triangle(Ctx& c, int idx, V3 o, V3 d, float t_min, float t_max, float* t, V3* p, V3* n) {
    c.cnt.triangle_tests++;
    const OScene* sc = c.sc;
    V3 v0 = ld3(sc->tri_v0 + 3 * idx), e1 = ld3(sc->tri_edge1 + 3 * idx), e2 = ld3(sc->tri_edge2 + 3 * idx);
    V3 normal = ld3(sc->tri_normal + 3 * idx);
    V3 h = cross(d, e2);
    float det = dot(e1, h);
    if (fabsf(det) >= 1e-8f) {
        float inv_det = 1.0f / det;
        V3 s = o - v0;
        float u = inv_det * dot(s, h);
        if (u >= 0.0f && u <= 1.0f) {
            V3 q = cross(s, e1);
            float v = inv_det * dot(d, q);
            if (v >= 0.0f && u + v <= 1.0f) {
                float tt = inv_det * dot(e2, q);
                if (tt >= t_min && tt <= t_max) {
                    *t = tt;
                    *p = o + tt * d;
                    *n = dot(d, normal) > 0.0f ? -normal : normal;
                    return true;
                }
            }
        }
    }
    return false;
}

/* kernels.py:310-362 */
bool hit_quad(Ctx& c, int idx, V3 o, V3 d, float t_min, float t_max, float* t, V3* p, V3* n) {
    c.cnt.quad_tests++;
    const OScene* sc = c.sc;
    V3 Q = ld3(sc->quad_Q + 3 * idx), u = ld3(sc->quad_u + 3 * idx), v = ld3(sc->quad_v + 3 * idx);
    V3 normal = ld3(sc->quad_normal + 3 * idx), w = ld3(sc->quad_w + 3 * idx);
    float D = sc->quad_D[idx];
    float denom = dot(normal, d);
    if (fabsf(denom) >= 1e-8f) {
        float tt = (D - dot(normal, o)) / denom;
        if (tt >= t_min && tt <= t_max) {
            V3 isect = o + tt * d;
            V3 ph = isect - Q;
            float alpha = dot(w, cross(ph, v));
            float beta = dot(w, cross(u, ph));
            if (alpha >= 0.0f && alpha <= 1.0f && beta >= 0.0f && beta <= 1.0f) {
                *t = tt;
                *p = isect;
                *n = denom < 0.0f ? normal : -normal;
                return true;
            }
        }
    }
    return false;
}

/* kernels.py:600-621 */
bool hit_aabb_optimized(V3 bmin, V3 bmax, V3 o, V3 inv, float t_min, float t_max) {
    V3 t0 = (bmin - o) * inv;
    V3 t1 = (bmax - o) * inv;
    V3 lo = v3(fmin2(t0.x, t1.x), fmin2(t0.y, t1.y), fmin2(t0.z, t1.z));
    V3 hi = v3(fmax2(t0.x, t1.x), fmax2(t0.y, t1.y), fmax2(t0.z, t1.z));
    float tmin = fmax2(fmax2(lo.x, lo.y), fmax2(lo.z, t_min));
    float tmax = fmin2(fmin2(hi.x, hi.y), fmin2(hi.z, t_max));
    return tmax >= tmin;
}

/* kernels.py:624-742 (traverse_bvh_legacy; the active branch of traverse_bvh, :749-759) */
Hit traverse_bvh(Ctx& c, V3 o, V3 d, float t_min, float t_max) {
    const OScene* sc = c.sc;
    c.cnt.rays++;
    Hit best;
    best.hit = false;
    best.t = t_max; /* closest_t */
    best.p = v3(0, 0, 0);
    best.n = v3(0, 0, 0);
    best.prim_type = 0;
    best.prim_idx = 0;
    V3 inv = v3(fabsf(d.x) > 1e-8f ? 1.0f / d.x : 1e8f, fabsf(d.y) > 1e-8f ? 1.0f / d.y : 1e8f,
                fabsf(d.z) > 1e-8f ? 1.0f / d.z : 1e8f);
    int stack[64];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        int ni = stack[--sp];
        if (ni < 0 || ni >= sc->num_nodes) continue;
        V3 bmin = ld3(sc->bvh_bbox_min + 3 * ni), bmax = ld3(sc->bvh_bbox_max + 3 * ni);
        c.cnt.node_tests++;
        if (!hit_aabb_optimized(bmin, bmax, o, inv, t_min, best.t)) continue;
        int prim_idx = sc->bvh_prim_idx[ni];
        if (prim_idx >= 0) {
            int prim_type = sc->bvh_prim_type[ni];
            bool hit = false;
            float t = 0.0f;
            V3 p = v3(0, 0, 0), n = v3(0, 0, 0);
            if (prim_type == 0) hit = hit_sphere(c, prim_idx, o, d, t_min, best.t, &t, &p, &n);
            else if (prim_type == 1) hit = hit_triangle(c, prim_idx, o, d, t_min, best.t, &t, &p, &n);
            else if (prim_type == 2) hit = hit_quad(c, prim_idx, o, d, t_min, best.t, &t, &p, &n);
            if (hit && t < best.t) {
                best.hit = true;
                best.t = t;
                best.p = p;
                best.n = n;
                best.prim_type = prim_type;
                best.prim_idx = prim_idx;
            }
        } else {
            int l = sc->bvh_left[ni], r = sc->bvh_right[ni];
            if (l >= 0 && r >= 0) {
                V3 lc = (ld3(sc->bvh_bbox_min + 3 * l) + ld3(sc->bvh_bbox_max + 3 * l)) * 0.5f;
                V3 rc = (ld3(sc->bvh_bbox_min + 3 * r) + ld3(sc->bvh_bbox_max + 3 * r)) * 0.5f;
                float ld = dot(lc - o, d);
                float rd = dot(rc - o, d);
                if (ld < rd) {
                    if (sp < 64) stack[sp++] = r;
                    if (sp < 64) stack[sp++] = l;
                } else {
                    if (sp < 64) stack[sp++] = l;
                    if (sp < 64) stack[sp++] = r;
                }
            } else {
                if (r >= 0 && sp < 64) stack[sp++] = r;
                if (l >= 0 && sp < 64) stack[sp++] = l;
            }
        }
    }
    return best;
}

/* kernels.py:766-786 */
V3 reflect(V3 v, V3 n) { return v - (2.0f * dot(v, n)) * n; }
V3 refract(V3 uv, V3 n, float eta) {
    float cos_theta = fmin2(-dot(uv, n), 1.0f);
    V3 perp = eta * (uv + cos_theta * n);
    V3 par = (-sqrtf(fabsf(1.0f - dot(perp, perp)))) * n;
    return perp + par;
}
float reflectance(float cosine, float ref_idx) {
    float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
    r0 = r0 * r0;
    return r0 + (1.0f - r0) * o_pow5f(1.0f - cosine);
}

/* kernels.py:789-814 */
V3 emitted(const OScene* sc, int prim_type, int prim_idx) {
    const OMaterials& m = sc->mat[prim_type];
    if (m.material_type[prim_idx] == 3) return ld3(m.emit + 3 * prim_idx);
    return v3(0, 0, 0);
}

/* kernels.py:924-1017 */
V3 eval_texture(const OScene* sc, int prim_type, int prim_idx, V3 hp) {
    const OMaterials& m = sc->mat[prim_type];
    int tex_type = m.texture_type[prim_idx];
    V3 color1 = ld3(m.color1 + 3 * prim_idx), color2 = ld3(m.color2 + 3 * prim_idx);
    float scale = m.texture_scale[prim_idx];
    int img_idx = m.image_idx[prim_idx];
    V3 result = v3(1, 1, 1);
    if (tex_type == 0) {
        result = color1;
    } else if (tex_type == 1) {
        float inv_scale = 1.0f / scale;
        int xi = (int)floorf(inv_scale * hp.x), yi = (int)floorf(inv_scale * hp.y), zi = (int)floorf(inv_scale * hp.z);
        int s = xi + yi + zi;
        bool is_even = (((s % 2) + 2) % 2) == 0; /* Python-style modulo */
        result = is_even ? color1 : color2;
    } else if (tex_type == 2) {
        if (prim_type == 0) {
            const float* sd = sc->sphere_data + 4 * prim_idx;
            float u, v;
            get_sphere_uv(hp, v3(sd[0], sd[1], sd[2]), &u, &v);
            if (img_idx >= 0 && img_idx < sc->num_images) {
                int iw = sc->image_w[img_idx], ih = sc->image_h[img_idx];
                u = fmax2(0.0f, fmin2(1.0f, u));
                v = 1.0f - fmax2(0.0f, fmin2(1.0f, v));
                int i = (int)(u * (float)iw);
                int j = (int)(v * (float)ih);
                i = i > iw - 1 ? iw - 1 : i; i = i < 0 ? 0 : i;   /* max(0, min(w-1, i)) */
                j = j > ih - 1 ? ih - 1 : j; j = j < 0 ? 0 : j;
                const uint8_t* px = sc->images[img_idx] + 3 * ((size_t)j * iw + i);
                result = v3((float)px[0] / 255.0f, (float)px[1] / 255.0f, (float)px[2] / 255.0f);
            }
        } else {
            result = v3(1.0f, 0.0f, 1.0f);
        }
    } else if (tex_type == 3) {
        float noise_val = o_sinf(scale * hp.z + 10.0f * perlin_turb(sc, hp, 3));
        result = (color1 * 0.5f) * (1.0f + noise_val);
    }
    return result;
}

/* kernels.py:817-917 */
/* core/material.py:258-276 _sample_henyey_greenstein */
V3 sample_henyey_greenstein(V3 incident, float g, Rng& rng) {
    if (fabsf(g) < 1e-3f) return random_unit_vector(rng);
    float sqr_term = (1.0f - g * g) / (1.0f - g + 2.0f * g * rng.next());
    float cos_theta = (1.0f + g * g - sqr_term * sqr_term) / (2.0f * g);
    float sin_theta = sqrtf(fmax2(0.0f, 1.0f - cos_theta * cos_theta));
    float phi = 6.2831853f * rng.next();                       /* 2 * 3.14159265 */
    V3 w = normalized(incident);
    V3 a = fabsf(w.x) > 0.9f ? v3(1, 0, 0) : v3(0, 1, 0);
    V3 v = normalized(cross(w, a));
    V3 u = cross(w, v);
    return normalized(((u * o_cosf(phi)) * sin_theta + (v * o_sinf(phi)) * sin_theta) + w * cos_theta);
}

bool near_zero(V3 v) { return fabsf(v.x) < 1e-8f && fabsf(v.y) < 1e-8f && fabsf(v.z) < 1e-8f; }

bool scatter(const OScene* sc, V3 ray_dir, V3 hp, V3 normal, int prim_type, int prim_idx, Rng& rng,
             V3* scatter_dir, V3* attenuation, V3* scatter_origin) {
    const OMaterials& m = sc->mat[prim_type];
    int mat_type = m.material_type[prim_idx];
    *scatter_dir = v3(0, 0, 0);
    *attenuation = v3(1, 1, 1);
    *scatter_origin = hp;
    bool scattered = false;
    if (sc->full_features && (mat_type == 5 || mat_type == 6)) {
        /* the pure-Python tracer's materials see the face-forwarded normal (core/hittable.py:17-22) */
        V3 nf = dot(ray_dir, normal) < 0.0f ? normal : -normal;
        V3 albedo = ld3(m.albedo + 3 * prim_idx);
        if (mat_type == 5) {                                   /* subsurface_simple, core/material.py:144-173 */
            float scatter_distance = m.fuzz[prim_idx];
            V3 dir;
            if (rng.next() < 0.5f) {
                dir = nf + random_unit_vector(rng);
            } else {
                V3 ruv = random_unit_vector(rng);
                V3 displacement = (ruv * scatter_distance) * rng.next();
                *scatter_origin = hp + displacement;
                dir = nf + random_unit_vector(rng);
            }
            if (near_zero(dir)) dir = nf;
            *scatter_dir = dir;
            *attenuation = albedo;
            return true;
        }
        /* subsurface_volumetric, core/material.py:175-256 */
        float sigma_s = m.fuzz[prim_idx], sigma_a = m.ir[prim_idx], g = m.texture_scale[prim_idx];
        float sigma_t = sigma_s + sigma_a;
        V3 pos = hp - nf * 0.001f;
        V3 dirn = normalized(ray_dir);
        V3 thr = v3(1, 1, 1);
        for (int it = 0; it < 64; it++) {
            float t = -o_logf(fmax2(rng.next(), 1e-10f)) / sigma_t;
            pos = pos + dirn * t;
            V3 dv = pos - hp;
            float dist = sqrtf(dot(dv, dv));
            float exit_probability = 1.0f - o_expf(-dist * 0.5f);
            if (rng.next() < exit_probability) {
                V3 dir = nf + random_unit_vector(rng);
                if (near_zero(dir)) dir = nf;
                *scatter_origin = pos;
                *scatter_dir = dir;
                *attenuation = thr * albedo;
                return true;
            }
            if (rng.next() < sigma_a / sigma_t) { *attenuation = v3(0, 0, 0); return false; }
            dirn = sample_henyey_greenstein(dirn, g, rng);
            thr = thr * albedo;
        }
        return false;
    }
    if (mat_type == 0) {
        V3 albedo = eval_texture(sc, prim_type, prim_idx, hp);
        *scatter_dir = random_cosine_direction(normal, rng);
        *attenuation = albedo;
        scattered = true;
    } else if (mat_type == 1) {
        V3 albedo = ld3(m.albedo + 3 * prim_idx);
        float fuzz = m.fuzz[prim_idx];
        V3 reflected = reflect(normalized(ray_dir), normal);
        *scatter_dir = reflected + fuzz * random_unit_vector(rng);
        if (dot(*scatter_dir, normal) > 0.0f) {
            *attenuation = albedo;
            scattered = true;
        }
    } else if (mat_type == 2) {
        float ir = m.ir[prim_idx];
        bool front_face = dot(ray_dir, normal) < 0.0f;
        V3 nf = front_face ? normal : -normal;
        float ratio = front_face ? (1.0f / ir) : ir;
        V3 unit_dir = normalized(ray_dir);
        float cos_theta = fmin2(-dot(unit_dir, nf), 1.0f);
        float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
        bool cannot_refract = ratio * sin_theta > 1.0f;
        /* `a or b` short-circuits: the random number is drawn only when refraction is possible */
        if (cannot_refract || reflectance(cos_theta, ratio) > rng.next()) *scatter_dir = reflect(unit_dir, nf);
        else *scatter_dir = refract(unit_dir, nf, ratio);
        *attenuation = v3(1, 1, 1);
        scattered = true;
    } else if (mat_type == 3) {
        scattered = false;
    } else if (mat_type == 4) {
        *scatter_dir = random_unit_vector(rng);
        *attenuation = eval_texture(sc, prim_type, prim_idx, hp);
        scattered = true;
    }
    return scattered;
}

struct MediumResult {
    bool is_medium_hit;
    float t_scatter;
    V3 scatter_point;
    float t_exit;
};

/* kernels.py:365-450 */
MediumResult apply_constant_medium(Ctx& c, int prim_type, int prim_idx, V3 o, V3 d, float t_min, float t_max,
                                   float t_entry, Rng& rng) {
    MediumResult r;
    r.is_medium_hit = false;
    r.t_scatter = 0.0f;
    r.scatter_point = v3(0, 0, 0);
    r.t_exit = 0.0f;
    const OMaterials& m = c.sc->mat[prim_type];
    int is_medium = m.is_medium[prim_idx];
    float density = m.medium_density[prim_idx];
    if (is_medium > 0) {
        Hit ex = traverse_bvh(c, o, d, t_entry + 0.0001f, 1e10f);
        if (ex.hit) {
            r.t_exit = ex.t;
            float t1 = fmax2(t_entry, t_min);
            float t2 = fmin2(r.t_exit, t_max);
            if (t1 < t2) {
                if (t1 < 0.0f) t1 = 0.0f;
                float ray_length = sqrtf(dot(d, d));
                float distance_inside = (t2 - t1) * ray_length;
                float hit_distance = -o_logf(fmax2(rng.next(), 1e-10f)) / density;
                if (hit_distance < distance_inside) {
                    r.t_scatter = t1 + hit_distance / ray_length;
                    r.scatter_point = o + r.t_scatter * d;
                    r.is_medium_hit = true;
                }
            }
        }
    }
    return r;
}

/* ---- next-event estimation (features "nee"; no counterpart in the reference's code -- its docs claim it,
 * docs/optimizations.md:53, and list it as missing, docs/__optimizations.md:7-10).  This is the CPU twin of
 * pt_device.cuh nee_direct(), operation for operation: at a Lambertian hit one emitter is chosen uniformly, a point
 * on it is sampled uniformly by area and its radiance added if the shadow ray reaches it,
 *     L = albedo/pi * Le * cos_s * cos_l / dist^2 * (area * n_lights) * T,
 * where the shadow ray follows the reference's constant-medium semantics (kernels.py:365-450, :1100-1119): a medium
 * boundary is a transmittance exp(-density * length inside), "inside" ends at the next hit of any primitive and the
 * ray continues beyond it; an exit search that finds nothing makes the boundary an (opaque) surface. */
void build_lights(const OScene* sc, std::vector<OLight>& out) {
    const int counts[3] = {sc->num_spheres, sc->num_triangles, sc->num_quads};
    for (int pt = 0; pt < 3; pt++) {
        const OMaterials& m = sc->mat[pt];
        for (int i = 0; i < counts[pt]; i++) {
            if (m.material_type[i] != 3 || m.is_medium[i] > 0) continue;
            const float* e = m.emit + 3 * i;
            if (!(e[0] > 0.0f || e[1] > 0.0f || e[2] > 0.0f)) continue;
            double area = 0.0;
            if (pt == 0) {
                const double r = sc->sphere_data[4 * i + 3];
                area = 4.0 * 3.14159265358979323846 * r * r;
            } else {
                const float* a = pt == 2 ? sc->quad_u + 3 * i : sc->tri_edge1 + 3 * i;
                const float* b = pt == 2 ? sc->quad_v + 3 * i : sc->tri_edge2 + 3 * i;
                const double cx = (double)a[1] * b[2] - (double)a[2] * b[1], cy = (double)a[2] * b[0] - (double)a[0] * b[2],
                             cz = (double)a[0] * b[1] - (double)a[1] * b[0];
                area = sqrt(cx * cx + cy * cy + cz * cz) * (pt == 2 ? 1.0 : 0.5);
            }
            if (!(area > 0.0)) continue;
            out.push_back(OLight{pt, i, (float)area, v3(e[0], e[1], e[2])});
        }
    }
}
inline bool nee_on(const Ctx& c) { return (c.sc->full_features & 2) != 0 && c.n_lights > 0; }
/* 0.25 * mean emitter area: (vertex, emitter point) pairs farther apart than its square root are estimated by light
 * sampling, closer ones by the cosine-sampled ray hitting the emitter -- a partition of unity, so each pair is counted
 * exactly once (area sampling has unbounded variance for an emitter that is close) */
float nee_switch_r2(const std::vector<OLight>& lights) {
    double sum = 0.0;
    for (const OLight& l : lights) sum += (double)l.area;
    return lights.empty() ? 0.0f : (float)(0.25 * (sum / (double)lights.size()));
}

V3 nee_direct(Ctx& c, V3 hp, V3 n_s, V3 albedo, Rng& rng) {
    const OScene* sc = c.sc;
    const V3 zero = v3(0, 0, 0);
    int li = (int)(rng.next() * (float)c.n_lights);
    li = li < c.n_lights ? li : c.n_lights - 1;
    const OLight& L = c.lights[li];
    V3 y, nl;
    if (L.prim_type == 0) {
        V3 dirn = random_unit_vector(rng);
        y = sphere_center(c, L.prim_idx) + sc->sphere_data[4 * L.prim_idx + 3] * dirn;
        nl = dirn;
    } else if (L.prim_type == 2) {
        float a = rng.next(), b = rng.next();
        y = (ld3(sc->quad_Q + 3 * L.prim_idx) + a * ld3(sc->quad_u + 3 * L.prim_idx)) + b * ld3(sc->quad_v + 3 * L.prim_idx);
        nl = ld3(sc->quad_normal + 3 * L.prim_idx);
    } else {
        float r1 = rng.next(), r2 = rng.next();
        if (r1 + r2 > 1.0f) { r1 = 1.0f - r1; r2 = 1.0f - r2; }
        y = (ld3(sc->tri_v0 + 3 * L.prim_idx) + r1 * ld3(sc->tri_edge1 + 3 * L.prim_idx)) + r2 * ld3(sc->tri_edge2 + 3 * L.prim_idx);
        nl = ld3(sc->tri_normal + 3 * L.prim_idx);
    }
    V3 dv = y - hp;
    float dist2 = dot(dv, dv);
    if (!(dist2 > 1e-12f)) return zero;
    const bool close = !(dist2 >= c.nee_r2);   /* close pairs belong to the scattered ray, unless a medium lies between */
    bool crossed = false;
    float dist = sqrtf(dist2);
    V3 wi = dv / dist;
    float cos_s = dot(n_s, wi), cos_l = fabsf(dot(nl, wi));
    if (!(cos_s > 0.0f) || !(cos_l > 0.0f)) return zero;
    const float t_end = dist * 0.9999f;
    const float len = sqrtf(dot(wi, wi));
    float trans = 1.0f, t_cur = 0.001f;
    OCounters saved = c.cnt;                                   /* shadow rays are not counted (nor are they on the device) */
    for (int seg = 0; seg < 16; seg++) {
        if (!(t_cur < t_end)) break;
        Hit h = traverse_bvh(c, hp, wi, t_cur, t_end);
        if (!h.hit) break;
        if (!(sc->mat[h.prim_type].is_medium[h.prim_idx] > 0)) { c.cnt = saved; return zero; }
        Hit e = traverse_bvh(c, hp, wi, h.t + 0.0001f, 1e10f);
        if (!e.hit) { c.cnt = saved; return zero; }
        crossed = true;
        float density = sc->mat[h.prim_type].medium_density[h.prim_idx];
        float t1 = fmax2(h.t, 0.001f), t2 = fmin2(e.t, t_end);
        if (t1 < t2) trans = trans * o_expf(-(density * ((t2 - t1) * len)));
        /* the "inside" interval reaches the emitter itself: a ray that passes such an interval continues BEYOND its closing
         * surface (kernels.py:1108), i.e. it never sees this point of the emitter (a sphere is then seen from inside, at
         * its far side -- the pair this loop reaches when the sampled point lies there) */
        if (!(e.t < t_end)) { c.cnt = saved; return zero; }
        t_cur = e.t + 0.001f / len;
        if (seg == 15) { c.cnt = saved; return zero; }
    }
    c.cnt = saved;
    if (close && !crossed) return zero;
    float geom = (((cos_s * cos_l) * L.area) * (float)c.n_lights) / (3.14159265358979f * dist2);
    return (albedo * L.emit) * (geom * trans);
}

/* kernels.py:1024-1170 */
V3 trace_ray(Ctx& c, V3 ray_origin, V3 ray_dir, Rng& rng) {
    const OScene* sc = c.sc;
    const int max_depth = c.cam->max_depth;
    V3 color = v3(0, 0, 0);
    V3 throughput = v3(1, 1, 1);
    V3 cur_o = ray_origin;
    V3 cur_d = normalized(ray_dir);
    const bool nee = nee_on(c);
    int prev_nee = 0;   /* 1: the last scatter was a Lambertian vertex with next-event estimation; 2: ... and the ray has
                         * passed through a medium since (pairs across a medium always belong to the light sampler) */
    for (int depth = 0; depth < max_depth; depth++) {
        c.cnt.bounces++;
        Hit h = traverse_bvh(c, cur_o, cur_d, 0.001f, 1e10f);
        if (h.hit) {
            int is_cm = sc->mat[h.prim_type].is_medium[h.prim_idx];
            bool scattered = false, did_nee = false, surface = false;
            V3 scatter_dir = v3(0, 0, 0), attenuation = v3(1, 1, 1);
            bool volume_passthrough = false;
            V3 hit_point = h.p, normal = h.n;
            if (is_cm > 0) {
                MediumResult mr = apply_constant_medium(c, h.prim_type, h.prim_idx, cur_o, cur_d, 0.001f, 1e10f, h.t, rng);
                if (mr.is_medium_hit) {
                    hit_point = mr.scatter_point;
                    normal = v3(1, 0, 0);
                    scatter_dir = random_unit_vector(rng);
                    attenuation = ld3(sc->mat[h.prim_type].medium_albedo + 3 * h.prim_idx);
                    scattered = true;
                } else if (mr.t_exit > 0.0f) {
                    volume_passthrough = true;
                    float ray_length = sqrtf(dot(cur_d, cur_d));
                    float epsilon_t = 0.001f / ray_length;
                    cur_o = cur_o + cur_d * (mr.t_exit + epsilon_t);
                } else {
                    surface = true;
                }
            } else {
                surface = true;
            }
            if (surface) {
                if (nee && sc->mat[h.prim_type].material_type[h.prim_idx] == 0) {
                    color = color + throughput * nee_direct(c, hit_point, normal, eval_texture(sc, h.prim_type, h.prim_idx, hit_point), rng);
                    did_nee = true;
                }
                V3 emit = emitted(sc, h.prim_type, h.prim_idx);
                if (prev_nee == 0 || (prev_nee == 1 && (h.t * h.t) * dot(cur_d, cur_d) < c.nee_r2)) color = color + throughput * emit;
                scattered = scatter(sc, cur_d, hit_point, normal, h.prim_type, h.prim_idx, rng, &scatter_dir, &attenuation, &hit_point);
            }
            if (nee) prev_nee = volume_passthrough ? (prev_nee != 0 ? 2 : 0) : (did_nee ? 1 : 0);
            if (scattered) {
                cur_o = hit_point;     /* = the hit point, or the displaced exit point of a subsurface material */
                cur_d = scatter_dir;
                throughput = throughput * attenuation;
                if (depth + 1 >= max_depth) break;
                if (depth + 1 >= 5) {
                    float survival = fmin2(fmax2(fmax2(throughput.x, throughput.y), throughput.z), 0.95f);
                    if (rng.next() > survival) break;
                    throughput = throughput / survival;
                }
            } else if (!volume_passthrough) {
                break;
            }
        } else {
            color = color + throughput * ld3(c.cam->bg);
            break;
        }
    }
    return color;
}

void add_counters(OCounters* dst, const OCounters& s) {
    dst->rays += s.rays; dst->node_tests += s.node_tests; dst->sphere_tests += s.sphere_tests;
    dst->triangle_tests += s.triangle_tests; dst->quad_tests += s.quad_tests; dst->paths += s.paths;
    dst->bounces += s.bounces;
}

} /* namespace */

extern "C" {

/* Elementary functions, vectorised, for tests/test_gpu_math.py and oracle/taichi_shim */
void o_math_eval(int fn, int n, const float* a, const float* b, float* out) {
    for (int i = 0; i < n; i++) {
        switch (fn) {
            case 0: out[i] = o_sinf(a[i]); break;
            case 1: out[i] = o_cosf(a[i]); break;
            case 2: out[i] = o_logf(a[i]); break;
            case 3: out[i] = o_acosf(a[i]); break;
            case 4: out[i] = o_atan2f(a[i], b[i]); break;
            case 5: out[i] = o_pow5f(a[i]); break;
            case 7: out[i] = o_expf(a[i]); break;
            default: out[i] = 0.0f;
        }
    }
}

/* a stream's state crosses the C ABI as one 64-bit word: position in the low half, stride in the high half */
uint64_t o_rng_seed_state(uint32_t seed, uint32_t pixel, uint32_t sample) {
    ORng r = o_rng_seed(seed, pixel, sample);
    return ((uint64_t)r.inc << 32) | r.s;
}
/* low-discrepancy pixel jitter: the first n sample points of one pixel, (x, y) pairs in [0, 1) */
void o_ld_jitter_points(uint32_t seed, uint32_t pixel, int n, float* out) {
    for (int i = 0; i < n; i++) o_ld_jitter(seed, pixel, (uint32_t)i, out + 2 * i, out + 2 * i + 1);
}
/* batch form for the stream-uniqueness tests: states of n (pixel, sample) keys */
void o_rng_seed_states(uint32_t seed, int64_t n, const uint32_t* pixels, const uint32_t* samples, uint64_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) out[i] = o_rng_seed_state(seed, pixels[i], samples[i]);
}
/* first draw of each of n streams (u32 before the float conversion), same purpose */
void o_rng_first_draws(int64_t n, const uint64_t* states, int k, uint32_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        uint32_t s = (uint32_t)states[i], inc = (uint32_t)(states[i] >> 32);
        for (int j = 0; j < k; j++) { s += inc; out[i * k + j] = o_mix32(s); }
    }
}
void o_rng_draw(uint64_t* state, int n, float* out) {
    ORng r = {(uint32_t)*state, (uint32_t)(*state >> 32)};
    for (int i = 0; i < n; i++) out[i] = o_rng_next(&r);
    *state = ((uint64_t)r.inc << 32) | r.s;
}

/* get_ray for sample `sample` of every pixel (kernels.py:176-201; generate_camera_rays :1219-1239).
 * Also returns the per-path RNG state after ray generation. */
void o_generate_camera_rays(const OCamera* cam, int W, int H, uint32_t seed, int sample, float* out_o, float* out_d,
                            uint32_t* out_rng) {
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            int idx = py * W + px;
            Rng rng{o_rng_seed(seed, (uint32_t)idx, (uint32_t)sample)};
            V3 o, d;
            get_ray(cam, px, py, rng, &o, &d, seed, (uint32_t)idx, (uint32_t)sample);
            out_o[3 * idx] = o.x; out_o[3 * idx + 1] = o.y; out_o[3 * idx + 2] = o.z;
            out_d[3 * idx] = d.x; out_d[3 * idx + 1] = d.y; out_d[3 * idx + 2] = d.z;
            if (out_rng) out_rng[idx] = rng.s.s;
        }
}

/* traverse_bvh over an arbitrary ray list (intersect_rays, kernels.py:1242-1263) */
void o_first_hit(const OScene* sc, int n, const float* o, const float* d, float t_min, float t_max, int32_t* out_valid,
                 float* out_t, int32_t* out_type, int32_t* out_idx, float* out_p, float* out_n, OCounters* counters,
                 int nthreads) {
    OCounters total;
    memset(&total, 0, sizeof total);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        Ctx c;
        c.sc = sc;
        c.cam = nullptr;
        memset(&c.cnt, 0, sizeof c.cnt);
#pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < n; i++) {
            Hit h = traverse_bvh(c, ld3(o + 3 * i), ld3(d + 3 * i), t_min, t_max);
            out_valid[i] = h.hit ? 1 : 0;
            out_t[i] = h.t; /* traverse_bvh returns closest_t (kernels.py:742): t_max on a miss */
            out_type[i] = h.prim_type;
            out_idx[i] = h.prim_idx;
            if (out_p) { out_p[3 * i] = h.p.x; out_p[3 * i + 1] = h.p.y; out_p[3 * i + 2] = h.p.z; }
            if (out_n) { out_n[3 * i] = h.n.x; out_n[3 * i + 1] = h.n.y; out_n[3 * i + 2] = h.n.z; }
        }
#pragma omp critical
        add_counters(&total, c.cnt);
    }
    if (counters) add_counters(counters, total);
}

/* render_sample (kernels.py:1177-1187) for samples [sample_begin, sample_begin+sample_count) and
 * rows [y0, y1): accum[py,px] += trace_ray(get_ray(px,py)), one sample after the other. */
void o_render_megakernel(const OScene* sc, const OCamera* cam, int W, int H, int y0, int y1, int sample_begin,
                         int sample_count, uint32_t seed, float* accum, OCounters* counters, int nthreads) {
    OCounters total;
    memset(&total, 0, sizeof total);
    std::vector<OLight> lights;
    if (sc->full_features & 2) build_lights(sc, lights);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        Ctx c;
        c.sc = sc;
        c.cam = cam;
        c.lights = lights.data();
        c.n_lights = (int)lights.size();
        c.nee_r2 = nee_switch_r2(lights);
        memset(&c.cnt, 0, sizeof c.cnt);
#pragma omp for schedule(dynamic, 1)
        for (int py = y0; py < y1; py++)
            for (int px = 0; px < W; px++) {
                int idx = py * W + px;
                V3 acc = ld3(accum + 3 * (size_t)idx);
                for (int s = sample_begin; s < sample_begin + sample_count; s++) {
                    Rng rng{o_rng_seed(seed, (uint32_t)idx, (uint32_t)s)};
                    c.time = sc->full_features ? o_path_time(rng.s) : 0.0f;
                    V3 o, d;
                    get_ray(cam, px, py, rng, &o, &d, seed, (uint32_t)idx, (uint32_t)s);
                    V3 col = trace_ray(c, o, d, rng);
                    acc = acc + col;
                    c.cnt.paths++;
                }
                accum[3 * (size_t)idx] = acc.x; accum[3 * (size_t)idx + 1] = acc.y; accum[3 * (size_t)idx + 2] = acc.z;
            }
#pragma omp critical
        add_counters(&total, c.cnt);
    }
    if (counters) add_counters(counters, total);
}

/* The wavefront pipeline (kernels.py:1219-1418 driven by renderer.py:305-334), one sample at a
 * time: generate_camera_rays; then per bounce intersect_rays, shade_miss_rays,
 * reset_next_ray_count, shade_and_scatter, swap_ray_buffers, for at most max_depth waves.
 * A path never forks, so each pixel owns at most one ray per wave and the queue order chosen by
 * the atomic append (kernels.py:1345, :1394) cannot change any pixel's result; the restatement
 * therefore processes paths one after the other, wave by wave, with the per-path random stream
 * carried in the ray record. */
void o_render_wavefront(const OScene* sc, const OCamera* cam, int W, int H, int y0, int y1, int sample_begin,
                        int sample_count, uint32_t seed, float* accum, OCounters* counters, int nthreads) {
    OCounters total;
    memset(&total, 0, sizeof total);
    const int max_depth = cam->max_depth;
    std::vector<OLight> lights;
    if (sc->full_features & 2) build_lights(sc, lights);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        Ctx c;
        c.sc = sc;
        c.cam = cam;
        c.lights = lights.data();
        c.n_lights = (int)lights.size();
        c.nee_r2 = nee_switch_r2(lights);
        memset(&c.cnt, 0, sizeof c.cnt);
#pragma omp for schedule(dynamic, 1)
        for (int py = y0; py < y1; py++)
            for (int px = 0; px < W; px++) {
                int idx = py * W + px;
                float* A = accum + 3 * (size_t)idx;
                for (int s = sample_begin; s < sample_begin + sample_count; s++) {
                    c.cnt.paths++;
                    Rng rng{o_rng_seed(seed, (uint32_t)idx, (uint32_t)s)};
                    c.time = sc->full_features ? o_path_time(rng.s) : 0.0f;
                    /* generate_camera_rays: direction NOT normalised (kernels.py:1229-1233) */
                    V3 ro, rd;
                    get_ray(cam, px, py, rng, &ro, &rd, seed, (uint32_t)idx, (uint32_t)s);
                    V3 thr = v3(1, 1, 1);
                    int depth = 0;
                    bool active = true;
                    const bool nee = nee_on(c);
                    int prev_nee = 0;
                    /* the path's radiance is summed on its own and added to the pixel when the path ends (the device
                     * keeps it in a per-path slab that k_wf_resolve adds in sample order): with one contribution per
                     * path -- the reference's estimator -- that is the same sum as adding each event straight away */
                    V3 pc = v3(0, 0, 0);
                    for (int bounce = 0; bounce < max_depth && active; bounce++) { /* renderer.py:313 */
                        c.cnt.bounces++;
                        /* intersect_rays */
                        Hit h = traverse_bvh(c, ro, rd, 0.001f, 1e10f);
                        if (!h.hit) { /* shade_miss_rays, kernels.py:1266-1280 */
                            V3 add = thr * ld3(cam->bg);
                            pc = pc + add;
                            active = false;
                            break;
                        }
                        /* shade_and_scatter, kernels.py:1289-1399 */
                        V3 hit_point = h.p, normal = h.n;
                        int is_cm = sc->mat[h.prim_type].is_medium[h.prim_idx];
                        bool scattered = false, passthrough = false, surface = false, did_nee = false;
                        V3 scatter_dir = v3(0, 0, 0), attenuation = v3(1, 1, 1), emit = v3(0, 0, 0);
                        V3 next_o = ro;
                        if (is_cm > 0) {
                            MediumResult mr = apply_constant_medium(c, h.prim_type, h.prim_idx, ro, rd, 0.001f, 1e10f, h.t, rng);
                            if (mr.is_medium_hit) {
                                hit_point = mr.scatter_point;
                                scatter_dir = random_unit_vector(rng);
                                attenuation = ld3(sc->mat[h.prim_type].medium_albedo + 3 * h.prim_idx);
                                scattered = true;
                            } else if (mr.t_exit > 0.0f) {
                                float epsilon_t = 0.001f / sqrtf(dot(rd, rd));
                                next_o = ro + rd * (mr.t_exit + epsilon_t);
                                passthrough = true; /* re-enqueued with unchanged depth/throughput */
                            } else {
                                surface = true;
                            }
                        } else {
                            surface = true;
                        }
                        if (surface) {
                            if (nee && sc->mat[h.prim_type].material_type[h.prim_idx] == 0) {
                                V3 add = thr * nee_direct(c, hit_point, normal, eval_texture(sc, h.prim_type, h.prim_idx, hit_point), rng);
                                pc = pc + add;
                                did_nee = true;
                            }
                            emit = emitted(sc, h.prim_type, h.prim_idx);
                            scattered = scatter(sc, rd, hit_point, normal, h.prim_type, h.prim_idx, rng, &scatter_dir, &attenuation, &hit_point);
                        }
                        if ((emit.x > 0.0f || emit.y > 0.0f || emit.z > 0.0f) && (prev_nee == 0 || (prev_nee == 1 && (h.t * h.t) * dot(rd, rd) < c.nee_r2))) {
                            V3 add = thr * emit;
                            pc = pc + add;
                        }
                        if (nee) prev_nee = passthrough ? (prev_nee != 0 ? 2 : 0) : (did_nee ? 1 : 0);
                        active = false;
                        if (passthrough) {
                            ro = next_o;
                            active = true;
                        }
                        if (scattered) {
                            V3 new_thr = thr * attenuation;
                            int new_depth = depth + 1;
                            if (new_depth < max_depth) {
                                bool cont = true;
                                if (new_depth >= 5) {
                                    float survival = fmin2(fmax2(fmax2(new_thr.x, new_thr.y), new_thr.z), 0.95f);
                                    if (rng.next() > survival) cont = false;
                                    else new_thr = new_thr / survival;
                                }
                                if (cont) {
                                    ro = hit_point;
                                    rd = scatter_dir;
                                    thr = new_thr;
                                    depth = new_depth;
                                    active = true;
                                }
                            }
                        }
                    }
                    A[0] = A[0] + pc.x; A[1] = A[1] + pc.y; A[2] = A[2] + pc.z;
                }
            }
#pragma omp critical
        add_counters(&total, c.cnt);
    }
    if (counters) add_counters(counters, total);
}

/* preview.py:117-132 buffer_to_image */
void o_buffer_to_image(const float* accum, int n_floats, int sample_count, uint8_t* out) {
    float scale = 1.0f / (float)(sample_count > 1 ? sample_count : 1);
    for (int i = 0; i < n_floats; i++) {
        float v = accum[i] * scale;
        v = v > 0.0f ? v : 0.0f;
        float g = sqrtf(v) * 255.999f;
        g = g < 0.0f ? 0.0f : (g > 255.0f ? 255.0f : g);
        out[i] = (uint8_t)g;
    }
}

int o_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

} /* extern "C" */
