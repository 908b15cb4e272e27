/* oracle/blocks.h — TEST INFRASTRUCTURE (CPU oracle), never linked into the product.
 *
 * Wire format of the ggml quantised blocks on the MUL_MAT hot path, restated from the layouts
 * the reference declares in ggml/src/ggml-common.h:
 *   q4_0 :170-175   q8_0 :219-224   q4_K :295-306   q5_K :312-324   q6_K :330-336   q8_K :339-344
 * All multi-byte scalars are little-endian; `half` is IEEE binary16 stored as uint16_t.
 */
#ifndef ORACLE_BLOCKS_H
#define ORACLE_BLOCKS_H
#include <stdint.h>
#include <stddef.h>

#define OR_QK   32   /* elements per legacy block (q4_0/q8_0) */
#define OR_QKK 256   /* elements per K-quant super-block     */

#pragma pack(push, 1)
typedef struct { uint16_t d; uint8_t qs[16]; }                                   or_q4_0;  /* 18 B  */
typedef struct { uint16_t d; int8_t  qs[32]; }                                   or_q8_0;  /* 34 B  */
typedef struct { uint16_t d, dmin; uint8_t sc[12]; uint8_t qs[128]; }            or_q4_K;  /* 144 B */
typedef struct { uint16_t d, dmin; uint8_t sc[12]; uint8_t qh[32]; uint8_t qs[128]; } or_q5_K; /* 176 B */
typedef struct { uint8_t ql[128]; uint8_t qh[64]; int8_t sc[16]; uint16_t d; }   or_q6_K;  /* 210 B */
typedef struct { float d; int8_t qs[256]; int16_t bsums[16]; }                   or_q8_K;  /* 292 B */
#pragma pack(pop)

_Static_assert(sizeof(or_q4_0) == 18,  "q4_0");
_Static_assert(sizeof(or_q8_0) == 34,  "q8_0");
_Static_assert(sizeof(or_q4_K) == 144, "q4_K");
_Static_assert(sizeof(or_q5_K) == 176, "q5_K");
_Static_assert(sizeof(or_q6_K) == 210, "q6_K");
_Static_assert(sizeof(or_q8_K) == 292, "q8_K");

/* ggml_type ids (ggml/include/ggml.h:379-421) for the types this path handles */
enum { OR_F32 = 0, OR_F16 = 1, OR_Q4_0 = 2, OR_Q8_0 = 8, OR_Q4_K = 12, OR_Q5_K = 13, OR_Q6_K = 14, OR_Q8_K = 15,
       OR_BF16 = 30 };

static inline float or_h2f(uint16_t h) { _Float16 v; __builtin_memcpy(&v, &h, 2); return (float) v; }
static inline uint16_t or_f2h(float f) { _Float16 v = (_Float16) f; uint16_t h; __builtin_memcpy(&h, &v, 2); return h; }

#endif
