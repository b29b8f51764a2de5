/* TEST INFRASTRUCTURE ONLY (CPU oracle) -- never linked into or called by the product path.
 *
 * PT-F32 numeric contract, part 1: the elementary functions.
 *
 * The reference executes its kernels through Taichi, an un-pinned third-party JIT
 * (`import taichi as ti`, /root/reference/src/render_server/taichi_renderer/kernels.py:8) whose
 * `ti.sin/cos/log/acos/atan2/pow` lower to backend-specific intrinsics; none of the reference's
 * tests pins a value at that boundary.  To make "same inputs -> same bits"
 * a testable property, the contract fixes each of them to a published single-precision algorithm
 * (Cephes `sinf/cosf/logf/asinf/atanf`, S. Moshier, netlib cephes/single) evaluated in strict IEEE
 * binary32: one rounding per operation, no FMA contraction, left-to-right evaluation.
 * `sqrt` and `/` are the correctly rounded IEEE operations.
 *
 * The CUDA implementation (path_tracer_python_b200/csrc/pt_math.cuh) is written independently
 * against this same specification and is compared bit-for-bit in tests/test_gpu_math.py.
 */
#ifndef PT_ORACLE_MATH_H
#define PT_ORACLE_MATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint32_t o_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float o_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* Cody-Waite split of pi/4 used by both sin and cos (exact products for |j| < 8192). */
#define O_FOPI 1.27323954473516f
#define O_DP1 0.78515625f
#define O_DP2 2.4187564849853515625e-4f
#define O_DP3 3.77489497744594108e-8f

static inline float o_sin_poly(float x, float z) {
    float y = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * x;
    return y + x;
}
static inline float o_cos_poly(float z) {
    float y = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z;
    y = y - 0.5f * z;
    return y + 1.0f;
}

/* ti.sin (kernels.py:55, :1013) */
static inline float o_sinf(float xx) {
    int sign = 1;
    float x = xx;
    if (x < 0.0f) { sign = -1; x = -x; }
    if (!(x <= 16777215.0f)) return 0.0f; /* total loss of precision (and NaN) */
    int j = (int)(O_FOPI * x);
    float y = (float)j;
    if (j & 1) { j += 1; y = y + 1.0f; }
    j &= 7;
    if (j > 3) { sign = -sign; j -= 4; }
    x = ((x - y * O_DP1) - y * O_DP2) - y * O_DP3;
    float z = x * x;
    float r = (j == 1 || j == 2) ? o_cos_poly(z) : o_sin_poly(x, z);
    return sign < 0 ? -r : r;
}

/* ti.cos (kernels.py:54) */
static inline float o_cosf(float xx) {
    int sign = 1;
    float x = xx;
    if (x < 0.0f) x = -x;
    if (!(x <= 16777215.0f)) return 0.0f;
    int j = (int)(O_FOPI * x);
    float y = (float)j;
    if (j & 1) { j += 1; y = y + 1.0f; }
    j &= 7;
    if (j > 3) { j -= 4; sign = -sign; }
    if (j > 1) sign = -sign;
    x = ((x - y * O_DP1) - y * O_DP2) - y * O_DP3;
    float z = x * x;
    float r = (j == 1 || j == 2) ? o_sin_poly(x, z) : o_cos_poly(z);
    return sign < 0 ? -r : r;
}

/* ti.log (kernels.py:441); argument is max(rand, 1e-10) in (0, 1]; x <= 0 -> -inf like libm */
static inline float o_logf(float xx) {
    if (!(xx > 0.0f)) return -INFINITY;
    uint32_t u = o_f2u(xx);
    int e;
    if ((u >> 23) == 0) { /* subnormal: renormalise */
        float s = xx * 8388608.0f;
        u = o_f2u(s);
        e = (int)(u >> 23) - 126 - 23;
    } else {
        e = (int)(u >> 23) - 126;
    }
    float x = o_u2f((u & 0x007fffffu) | 0x3f000000u); /* mantissa in [0.5, 1) */
    if (x < 0.707106781186547524f) { e -= 1; x = x + x - 1.0f; }
    else { x = x - 1.0f; }
    float z = x * x;
    float y = ((((((((7.0376836292e-2f * x - 1.1514610310e-1f) * x + 1.1676998740e-1f) * x
                     - 1.2420140846e-1f) * x + 1.4249322787e-1f) * x - 1.6668057665e-1f) * x
                  + 2.0000714765e-1f) * x - 2.4999993993e-1f) * x + 3.3333331174e-1f) * x * z;
    float fe = (float)e;
    if (e) y = y + -2.12194440e-4f * fe;
    y = y + -0.5f * z;
    z = x + y;
    if (e) z = z + 0.693359375f * fe;
    return z;
}

static inline float o_asinf_core(float a) { /* a in [0,1] */
    float x, z;
    int flag;
    if (a < 1.0e-4f) return a;
    if (a > 0.5f) { z = 0.5f * (1.0f - a); x = sqrtf(z); flag = 1; }
    else { x = a; z = x * x; flag = 0; }
    z = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z
         + 1.6666752422e-1f) * z * x + x;
    if (flag) { z = z + z; z = 1.5707963267948966192f - z; }
    return z;
}
static inline float o_asinf(float x) {
    float a = x < 0.0f ? -x : x;
    float r = o_asinf_core(a);
    return x < 0.0f ? -r : r;
}

/* ti.acos (kernels.py:96); the contract clamps the argument to [-1,1] (normalised-vector
 * components can exceed 1 by an ulp; libm would give NaN there). */
static inline float o_acosf(float x) {
    if (x > 1.0f) x = 1.0f;
    if (x < -1.0f) x = -1.0f;
    if (x > 0.5f) return 2.0f * o_asinf(sqrtf(0.5f * (1.0f - x)));
    if (x < -0.5f) return 3.14159265358979323846f - 2.0f * o_asinf(sqrtf(0.5f * (1.0f + x)));
    return 1.5707963267948966192f - o_asinf(x);
}

static inline float o_atanf(float xx) {
    int neg = xx < 0.0f;
    float x = neg ? -xx : xx;
    float y;
    if (x > 2.414213562373095f) { y = 1.5707963267948966192f; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = 0.7853981633974483096f; x = (x - 1.0f) / (x + 1.0f); }
    else { y = 0.0f; }
    float z = x * x;
    y = y + ((((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z
              - 3.33329491539e-1f) * z * x + x);
    return neg ? -y : y;
}

/* ti.atan2 (kernels.py:97) */
static inline float o_atan2f(float y, float x) {
    int code = 0;
    if (x < 0.0f) code = 2;
    if (y < 0.0f) code |= 1;
    if (x == 0.0f) {
        if (code & 1) return -1.5707963267948966192f;
        if (y == 0.0f) return 0.0f;
        return 1.5707963267948966192f;
    }
    if (y == 0.0f) return (code & 2) ? 3.14159265358979323846f : 0.0f;
    float w = 0.0f;
    if (code == 2) w = 3.14159265358979323846f;
    else if (code == 3) w = -3.14159265358979323846f;
    return w + o_atanf(y / x);
}

/* exp (full-feature mode only: subsurface_volumetric's exit probability, core/material.py:213).
 * Cephes expf: range reduction by ln2 split in two constants, degree-5 polynomial, ldexp. */
static inline float o_ldexpf(float z, int n) {
    if (n < -126) { z = z * o_u2f(0x00800000u); n += 126; if (n < -126) n = -126; }   /* 2^-126 first */
    if (n > 127) { z = z * o_u2f(0x7f000000u); n -= 127; if (n > 127) n = 127; }
    return z * o_u2f((uint32_t)(n + 127) << 23);
}
static inline float o_expf(float xx) {
    float x = xx;
    if (x > 88.72283905206835f) return INFINITY;
    if (x < -103.278929903431851103f) return 0.0f;
    float z = floorf(1.44269504088896341f * x + 0.5f);
    x = x - z * 0.693359375f;
    x = x - z * -2.12194440e-4f;
    int n = (int)z;
    z = x * x;
    z = (((((1.9875691500e-4f * x + 1.3981999507e-3f) * x + 8.3334519073e-3f) * x + 4.1665795894e-2f) * x
          + 1.6666665459e-1f) * x + 5.0000001201e-1f) * z + x + 1.0f;
    return o_ldexpf(z, n);
}

/* ti.pow(x, 5.0) (kernels.py:786): the only call site has the literal exponent 5 */
static inline float o_pow5f(float x) {
    float x2 = x * x;
    return x2 * x2 * x;
}

/* PT-F32 contract, part 2: the per-path random stream that stands in for `ti.random()`
 * (kernels.py:21, :33, :47, :184, :441, :894, :1150, :1388; Taichi's own generator is an
 * un-seeded, backend-specific per-thread xorshift and is not reproducible).
 * A stream = a Weyl sequence on 32 bits with its own odd increment: s += inc; output = 2-round
 * multiply-xorshift finaliser of s; f32 = top 24 bits * 2^-24.  (start, increment) come from
 * (seed, pixel, sample) through three Feistel steps, each a bijection on 32 x 31 bits: for one seed
 * no two (pixel, sample) pairs share a stream (63 bits of stream identity). */
static inline uint32_t o_mix32(uint32_t z) {
    z ^= z >> 16; z *= 0x21f0aaadu;
    z ^= z >> 15; z *= 0x735a2d97u;
    z ^= z >> 15;
    return z;
}
typedef struct ORng { uint32_t s, inc; } ORng;
static inline ORng o_rng_seed(uint32_t seed, uint32_t pixel_index, uint32_t sample_index) {
    ORng r;
    uint32_t a1 = o_mix32(pixel_index ^ o_mix32(seed ^ 0xa511e9b3u));
    uint32_t b1 = (sample_index + (o_mix32(a1 ^ 0x3c6ef372u) >> 1)) & 0x7fffffffu;
    r.s = a1 + o_mix32(b1 ^ 0xdaa66d2bu);
    r.inc = 2u * b1 + 1u;
    return r;
}
/* full-feature mode: a path's shutter time (camera.py:136 `ray_time = random()`) is a pure function
 * of the path key, so that both kernel architectures can recompute it without carrying it */
static inline float o_path_time(ORng key) {
    return (float)(o_mix32(key.s ^ 0x7157a3d1u ^ (key.inc * 0x9e3779b9u)) >> 8) * 5.9604644775390625e-8f;
}
/* low-discrepancy pixel jitter (OCamera.jitter = 1): point `sample` of the Sobol' (0,2)-sequence, XOR-scrambled per pixel */
static inline uint32_t o_bitrev32(uint32_t v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
}
static inline uint32_t o_sobol2(uint32_t i) {
    uint32_t r = 0u, v;
    for (v = 0x80000000u; i; i >>= 1, v ^= v >> 1)
        if (i & 1u) r ^= v;
    return r;
}
static inline void o_ld_jitter(uint32_t seed, uint32_t pixel, uint32_t sample, float* jx, float* jy) {
    uint32_t sx = o_mix32(pixel ^ o_mix32(seed ^ 0x68bc21ebu)), sy = o_mix32(sx ^ 0x02e5be93u);
    *jx = (float)((o_bitrev32(sample) ^ sx) >> 8) * 5.9604644775390625e-8f;
    *jy = (float)((o_sobol2(sample) ^ sy) >> 8) * 5.9604644775390625e-8f;
}
static inline float o_rng_next(ORng* r) {
    r->s += r->inc;
    return (float)(o_mix32(r->s) >> 8) * 5.9604644775390625e-8f; /* 2^-24, exact */
}
#endif
