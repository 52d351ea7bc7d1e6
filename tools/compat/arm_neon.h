/* x86 stand-in for <arm_neon.h>.
 *
 * TEST INFRASTRUCTURE ONLY (see the checker README).  The reference's comp.h and
 * comp_prelu.h include <arm_neon.h> unconditionally (cpp_impl/comp.h:6,
 * cpp_impl/comp_prelu.h:6) although none of the NEON kernels is registered at
 * HEAD (cpp_impl/main.cpp:160-172).  This stub only has to let those never
 * executed templates parse under g++ on x86_64; it is put on the include path
 * by oracle/Makefile (and by the side-by-side test driver, host/Makefile REF=...) when the reference is compiled where it lies.
 */
#pragma once
#include <stdint.h>

typedef float float32x4_t __attribute__((vector_size(16)));
typedef uint32_t uint32x4_t __attribute__((vector_size(16)));

static inline float32x4_t vdupq_n_f32(float v) { return (float32x4_t){v, v, v, v}; }
static inline float32x4_t vaddq_f32(float32x4_t a, float32x4_t b) { return a + b; }
static inline float32x4_t vsubq_f32(float32x4_t a, float32x4_t b) { return a - b; }
static inline float32x4_t vmulq_f32(float32x4_t a, float32x4_t b) { return a * b; }
static inline float vaddvq_f32(float32x4_t a) { return (a[0] + a[1]) + (a[2] + a[3]); }
static inline float32x4_t vld1q_f32(const float *p) { float32x4_t r; __builtin_memcpy(&r, p, 16); return r; }
static inline void vst1q_f32(float *p, float32x4_t v) { __builtin_memcpy(p, &v, 16); }
static inline uint32x4_t vcgtq_f32(float32x4_t a, float32x4_t b) { return (uint32x4_t)(a > b); }
static inline float32x4_t vbslq_f32(uint32x4_t m, float32x4_t a, float32x4_t b)
{
    return (float32x4_t)((m & (uint32x4_t)a) | (~m & (uint32x4_t)b));
}
