/* oracle/gsl_shim/gsl/gsl_shim_all.h
 *
 * TEST INFRASTRUCTURE ONLY -- part of the CPU oracle, never linked into the product.
 *
 * Header-only stand-in for the subset of the GNU Scientific Library that the
 * reference's hot path uses (GSL itself is not installed in this image; the
 * reference requires gsl >= 1.10 and pins no version, btk/configure.in:119-128).
 * Written from the published GSL API semantics; nothing here comes from GSL's
 * sources.  Call sites that define the needed surface:
 *   btk/modulated/modulated.cc:439,603      gsl_fft_complex_radix2_{backward,forward}
 *   btk/beamformer/beamformer.cc:1181,2430  gsl_blas_zdotc / gsl_blas_zgemv
 *   btk/beamformer/beamformer.cc:2533       gsl_sf_sinc
 * The FFT is a plain double-precision radix-2 (bit reversal, then decimation in time)
 * whose twiddles advance by the standard trigonometric recurrence
 *   w <- w - (2 sin^2(theta/2)) w + j sin(theta) w        (Numerical Recipes 5.5),
 * one sin() pair per stage and no cos/sin inside the butterfly loops, which is the cost
 * profile of a real radix-2 library routine (results differ from GSL's in the last ulp
 * at most; the parity budget is 1e-4).
 */
#ifndef BTKB200_GSL_SHIM_ALL_H
#define BTKB200_GSL_SHIM_ALL_H

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

#define GSL_SUCCESS 0
#define GSL_FAILURE (-1)
#define GSL_EDOM 1
#define GSL_EINVAL 4
#define GSL_DBL_EPSILON 2.2204460492503131e-16
#define GSL_POSINF (INFINITY)
#define GSL_NEGINF (-INFINITY)
#define GSL_NAN (NAN)
#define GSL_MAX(a, b) ((a) > (b) ? (a) : (b))
#define GSL_MIN(a, b) ((a) < (b) ? (a) : (b))

/* ---------------------------------------------------------------- complex */
typedef struct { double dat[2]; } gsl_complex;
typedef struct { float dat[2]; } gsl_complex_float;
#define GSL_REAL(z) ((z).dat[0])
#define GSL_IMAG(z) ((z).dat[1])
#define GSL_SET_COMPLEX(zp, x, y) do { (zp)->dat[0] = (x); (zp)->dat[1] = (y); } while (0)
#define GSL_SET_REAL(zp, x) do { (zp)->dat[0] = (x); } while (0)
#define GSL_SET_IMAG(zp, y) do { (zp)->dat[1] = (y); } while (0)
#define GSL_COMPLEX_ONE (gsl_complex_rect(1.0, 0.0))
#define GSL_COMPLEX_ZERO (gsl_complex_rect(0.0, 0.0))

static inline gsl_complex gsl_complex_rect(double x, double y) { gsl_complex z; z.dat[0] = x; z.dat[1] = y; return z; }
static inline gsl_complex gsl_complex_polar(double r, double th) { return gsl_complex_rect(r * cos(th), r * sin(th)); }
static inline double gsl_complex_abs2(gsl_complex z) { return z.dat[0] * z.dat[0] + z.dat[1] * z.dat[1]; }
static inline double gsl_complex_abs(gsl_complex z) { return hypot(z.dat[0], z.dat[1]); }
static inline double gsl_complex_arg(gsl_complex z) { return (z.dat[0] == 0.0 && z.dat[1] == 0.0) ? 0.0 : atan2(z.dat[1], z.dat[0]); }
static inline double gsl_complex_logabs(gsl_complex z) { return log(gsl_complex_abs(z)); }
static inline gsl_complex gsl_complex_add(gsl_complex a, gsl_complex b) { return gsl_complex_rect(a.dat[0] + b.dat[0], a.dat[1] + b.dat[1]); }
static inline gsl_complex gsl_complex_sub(gsl_complex a, gsl_complex b) { return gsl_complex_rect(a.dat[0] - b.dat[0], a.dat[1] - b.dat[1]); }
static inline gsl_complex gsl_complex_mul(gsl_complex a, gsl_complex b) {
  return gsl_complex_rect(a.dat[0] * b.dat[0] - a.dat[1] * b.dat[1], a.dat[0] * b.dat[1] + a.dat[1] * b.dat[0]);
}
static inline gsl_complex gsl_complex_div(gsl_complex a, gsl_complex b) {
  double s = 1.0 / gsl_complex_abs(b);
  double sbr = s * b.dat[0], sbi = s * b.dat[1];
  return gsl_complex_rect((a.dat[0] * sbr + a.dat[1] * sbi) * s, (a.dat[1] * sbr - a.dat[0] * sbi) * s);
}
static inline gsl_complex gsl_complex_add_real(gsl_complex a, double x) { return gsl_complex_rect(a.dat[0] + x, a.dat[1]); }
static inline gsl_complex gsl_complex_sub_real(gsl_complex a, double x) { return gsl_complex_rect(a.dat[0] - x, a.dat[1]); }
static inline gsl_complex gsl_complex_mul_real(gsl_complex a, double x) { return gsl_complex_rect(a.dat[0] * x, a.dat[1] * x); }
static inline gsl_complex gsl_complex_div_real(gsl_complex a, double x) { return gsl_complex_rect(a.dat[0] / x, a.dat[1] / x); }
static inline gsl_complex gsl_complex_add_imag(gsl_complex a, double y) { return gsl_complex_rect(a.dat[0], a.dat[1] + y); }
static inline gsl_complex gsl_complex_mul_imag(gsl_complex a, double y) { return gsl_complex_rect(-y * a.dat[1], y * a.dat[0]); }
static inline gsl_complex gsl_complex_conjugate(gsl_complex a) { return gsl_complex_rect(a.dat[0], -a.dat[1]); }
static inline gsl_complex gsl_complex_negative(gsl_complex a) { return gsl_complex_rect(-a.dat[0], -a.dat[1]); }
static inline gsl_complex gsl_complex_inverse(gsl_complex a) {
  double s = 1.0 / gsl_complex_abs(a);
  return gsl_complex_rect((a.dat[0] * s) * s, -(a.dat[1] * s) * s);
}
static inline gsl_complex gsl_complex_exp(gsl_complex a) { double r = exp(a.dat[0]); return gsl_complex_rect(r * cos(a.dat[1]), r * sin(a.dat[1])); }
static inline gsl_complex gsl_complex_log(gsl_complex a) { return gsl_complex_rect(gsl_complex_logabs(a), gsl_complex_arg(a)); }
static inline gsl_complex gsl_complex_sqrt(gsl_complex a) {
  if (a.dat[0] == 0.0 && a.dat[1] == 0.0) return gsl_complex_rect(0, 0);
  double x = fabs(a.dat[0]), y = fabs(a.dat[1]), w;
  if (x >= y) { double t = y / x; w = sqrt(x) * sqrt(0.5 * (1.0 + sqrt(1.0 + t * t))); }
  else { double t = x / y; w = sqrt(y) * sqrt(0.5 * (t + sqrt(1.0 + t * t))); }
  if (a.dat[0] >= 0.0) return gsl_complex_rect(w, a.dat[1] / (2.0 * w));
  double vi = (a.dat[1] >= 0) ? w : -w;
  return gsl_complex_rect(a.dat[1] / (2.0 * vi), vi);
}
static inline gsl_complex gsl_complex_sqrt_real(double x) { return x >= 0 ? gsl_complex_rect(sqrt(x), 0) : gsl_complex_rect(0, sqrt(-x)); }
static inline gsl_complex gsl_complex_pow_real(gsl_complex a, double b) {
  if (a.dat[0] == 0 && a.dat[1] == 0) return gsl_complex_rect(b == 0 ? 1.0 : 0.0, 0.0);
  double rho = exp(gsl_complex_logabs(a) * b), beta = gsl_complex_arg(a) * b;
  return gsl_complex_rect(rho * cos(beta), rho * sin(beta));
}

/* ---------------------------------------------------------------- blocks / vectors / matrices */
#define BTKSHIM_DECL_CONTAINER(SUF, T, MULT)                                                      \
  typedef struct { size_t size; T* data; } gsl_block##SUF;                                        \
  typedef struct { size_t size; size_t stride; T* data; gsl_block##SUF* block; int owner; } gsl_vector##SUF; \
  typedef struct { size_t size1; size_t size2; size_t tda; T* data; gsl_block##SUF* block; int owner; } gsl_matrix##SUF; \
  static inline gsl_vector##SUF* gsl_vector##SUF##_alloc(size_t n) {                              \
    gsl_vector##SUF* v = (gsl_vector##SUF*)malloc(sizeof(gsl_vector##SUF));                       \
    gsl_block##SUF* b = (gsl_block##SUF*)malloc(sizeof(gsl_block##SUF));                          \
    b->size = n; b->data = (T*)malloc(sizeof(T) * (MULT) * (n ? n : 1));                          \
    v->size = n; v->stride = 1; v->data = b->data; v->block = b; v->owner = 1; return v; }        \
  static inline void gsl_vector##SUF##_set_zero(gsl_vector##SUF* v) {                             \
    for (size_t i = 0; i < v->size; i++) for (int k = 0; k < (MULT); k++) v->data[(MULT) * i * v->stride + k] = 0; } \
  static inline gsl_vector##SUF* gsl_vector##SUF##_calloc(size_t n) {                             \
    gsl_vector##SUF* v = gsl_vector##SUF##_alloc(n); gsl_vector##SUF##_set_zero(v); return v; }   \
  static inline void gsl_vector##SUF##_free(gsl_vector##SUF* v) {                                 \
    if (!v) return; if (v->owner && v->block) { free(v->block->data); free(v->block); } free(v); } \
  static inline int gsl_vector##SUF##_memcpy(gsl_vector##SUF* d, const gsl_vector##SUF* s) {      \
    if (d->size != s->size) { fprintf(stderr, "gsl shim: vector memcpy size mismatch\n"); abort(); } \
    for (size_t i = 0; i < s->size; i++) for (int k = 0; k < (MULT); k++)                         \
      d->data[(MULT) * i * d->stride + k] = s->data[(MULT) * i * s->stride + k];                  \
    return GSL_SUCCESS; }                                                                         \
  static inline gsl_matrix##SUF* gsl_matrix##SUF##_alloc(size_t n1, size_t n2) {                  \
    gsl_matrix##SUF* m = (gsl_matrix##SUF*)malloc(sizeof(gsl_matrix##SUF));                       \
    gsl_block##SUF* b = (gsl_block##SUF*)malloc(sizeof(gsl_block##SUF));                          \
    b->size = n1 * n2; b->data = (T*)malloc(sizeof(T) * (MULT) * (n1 * n2 ? n1 * n2 : 1));        \
    m->size1 = n1; m->size2 = n2; m->tda = n2; m->data = b->data; m->block = b; m->owner = 1; return m; } \
  static inline void gsl_matrix##SUF##_set_zero(gsl_matrix##SUF* m) {                             \
    for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++)                   \
      for (int k = 0; k < (MULT); k++) m->data[(MULT) * (i * m->tda + j) + k] = 0; }              \
  static inline gsl_matrix##SUF* gsl_matrix##SUF##_calloc(size_t n1, size_t n2) {                 \
    gsl_matrix##SUF* m = gsl_matrix##SUF##_alloc(n1, n2); gsl_matrix##SUF##_set_zero(m); return m; } \
  static inline void gsl_matrix##SUF##_free(gsl_matrix##SUF* m) {                                 \
    if (!m) return; if (m->owner && m->block) { free(m->block->data); free(m->block); } free(m); } \
  static inline int gsl_matrix##SUF##_memcpy(gsl_matrix##SUF* d, const gsl_matrix##SUF* s) {      \
    if (d->size1 != s->size1 || d->size2 != s->size2) { fprintf(stderr, "gsl shim: matrix memcpy size mismatch\n"); abort(); } \
    for (size_t i = 0; i < s->size1; i++) for (size_t j = 0; j < s->size2; j++)                   \
      for (int k = 0; k < (MULT); k++) d->data[(MULT) * (i * d->tda + j) + k] = s->data[(MULT) * (i * s->tda + j) + k]; \
    return GSL_SUCCESS; }

#define BTKSHIM_DECL_REAL_ACCESS(SUF, T)                                                          \
  static inline T gsl_vector##SUF##_get(const gsl_vector##SUF* v, size_t i) { return v->data[i * v->stride]; } \
  static inline void gsl_vector##SUF##_set(gsl_vector##SUF* v, size_t i, T x) { v->data[i * v->stride] = x; } \
  static inline T* gsl_vector##SUF##_ptr(gsl_vector##SUF* v, size_t i) { return v->data + i * v->stride; } \
  static inline void gsl_vector##SUF##_set_all(gsl_vector##SUF* v, T x) { for (size_t i = 0; i < v->size; i++) v->data[i * v->stride] = x; } \
  static inline int gsl_vector##SUF##_scale(gsl_vector##SUF* v, double x) { for (size_t i = 0; i < v->size; i++) v->data[i * v->stride] = (T)(v->data[i * v->stride] * x); return GSL_SUCCESS; } \
  static inline int gsl_vector##SUF##_add(gsl_vector##SUF* a, const gsl_vector##SUF* b) { for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] += b->data[i * b->stride]; return GSL_SUCCESS; } \
  static inline int gsl_vector##SUF##_sub(gsl_vector##SUF* a, const gsl_vector##SUF* b) { for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] -= b->data[i * b->stride]; return GSL_SUCCESS; } \
  static inline int gsl_vector##SUF##_add_constant(gsl_vector##SUF* a, double x) { for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] = (T)(a->data[i * a->stride] + x); return GSL_SUCCESS; } \
  static inline T gsl_vector##SUF##_max(const gsl_vector##SUF* v) { T m = v->data[0]; for (size_t i = 1; i < v->size; i++) if (v->data[i * v->stride] > m) m = v->data[i * v->stride]; return m; } \
  static inline T gsl_vector##SUF##_min(const gsl_vector##SUF* v) { T m = v->data[0]; for (size_t i = 1; i < v->size; i++) if (v->data[i * v->stride] < m) m = v->data[i * v->stride]; return m; } \
  static inline int gsl_vector##SUF##_fwrite(FILE* fp, const gsl_vector##SUF* v) { for (size_t i = 0; i < v->size; i++) if (fwrite(v->data + i * v->stride, sizeof(T), 1, fp) != 1) return GSL_FAILURE; return GSL_SUCCESS; } \
  static inline int gsl_vector##SUF##_fread(FILE* fp, gsl_vector##SUF* v) { for (size_t i = 0; i < v->size; i++) if (fread(v->data + i * v->stride, sizeof(T), 1, fp) != 1) return GSL_FAILURE; return GSL_SUCCESS; } \
  static inline T gsl_matrix##SUF##_get(const gsl_matrix##SUF* m, size_t i, size_t j) { return m->data[i * m->tda + j]; } \
  static inline void gsl_matrix##SUF##_set(gsl_matrix##SUF* m, size_t i, size_t j, T x) { m->data[i * m->tda + j] = x; } \
  static inline void gsl_matrix##SUF##_set_all(gsl_matrix##SUF* m, T x) { for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] = x; } \
  static inline void gsl_matrix##SUF##_set_identity(gsl_matrix##SUF* m) { for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] = (T)(i == j ? 1 : 0); } \
  static inline int gsl_matrix##SUF##_scale(gsl_matrix##SUF* m, double x) { for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] = (T)(m->data[i * m->tda + j] * x); return GSL_SUCCESS; } \
  static inline int gsl_matrix##SUF##_add(gsl_matrix##SUF* a, const gsl_matrix##SUF* b) { for (size_t i = 0; i < a->size1; i++) for (size_t j = 0; j < a->size2; j++) a->data[i * a->tda + j] += b->data[i * b->tda + j]; return GSL_SUCCESS; } \
  static inline int gsl_matrix##SUF##_sub(gsl_matrix##SUF* a, const gsl_matrix##SUF* b) { for (size_t i = 0; i < a->size1; i++) for (size_t j = 0; j < a->size2; j++) a->data[i * a->tda + j] -= b->data[i * b->tda + j]; return GSL_SUCCESS; } \
  static inline int gsl_matrix##SUF##_get_row(gsl_vector##SUF* v, const gsl_matrix##SUF* m, size_t i) { for (size_t j = 0; j < m->size2; j++) v->data[j * v->stride] = m->data[i * m->tda + j]; return GSL_SUCCESS; } \
  static inline int gsl_matrix##SUF##_get_col(gsl_vector##SUF* v, const gsl_matrix##SUF* m, size_t j) { for (size_t i = 0; i < m->size1; i++) v->data[i * v->stride] = m->data[i * m->tda + j]; return GSL_SUCCESS; } \
  static inline int gsl_matrix##SUF##_set_row(gsl_matrix##SUF* m, size_t i, const gsl_vector##SUF* v) { for (size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] = v->data[j * v->stride]; return GSL_SUCCESS; } \
  static inline int gsl_matrix##SUF##_set_col(gsl_matrix##SUF* m, size_t j, const gsl_vector##SUF* v) { for (size_t i = 0; i < m->size1; i++) m->data[i * m->tda + j] = v->data[i * v->stride]; return GSL_SUCCESS; } \
  static inline int gsl_matrix##SUF##_transpose_memcpy(gsl_matrix##SUF* d, const gsl_matrix##SUF* s) { for (size_t i = 0; i < s->size1; i++) for (size_t j = 0; j < s->size2; j++) d->data[j * d->tda + i] = s->data[i * s->tda + j]; return GSL_SUCCESS; }

BTKSHIM_DECL_CONTAINER(, double, 1)
BTKSHIM_DECL_REAL_ACCESS(, double)
BTKSHIM_DECL_CONTAINER(_float, float, 1)
BTKSHIM_DECL_REAL_ACCESS(_float, float)
BTKSHIM_DECL_CONTAINER(_short, short, 1)
BTKSHIM_DECL_REAL_ACCESS(_short, short)
BTKSHIM_DECL_CONTAINER(_char, char, 1)
BTKSHIM_DECL_REAL_ACCESS(_char, char)
BTKSHIM_DECL_CONTAINER(_int, int, 1)
BTKSHIM_DECL_REAL_ACCESS(_int, int)
BTKSHIM_DECL_CONTAINER(_uchar, unsigned char, 1)
BTKSHIM_DECL_REAL_ACCESS(_uchar, unsigned char)
BTKSHIM_DECL_CONTAINER(_complex, double, 2)
BTKSHIM_DECL_CONTAINER(_complex_float, float, 2)

/* complex double access */
static inline gsl_complex gsl_vector_complex_get(const gsl_vector_complex* v, size_t i) {
  return gsl_complex_rect(v->data[2 * i * v->stride], v->data[2 * i * v->stride + 1]);
}
static inline void gsl_vector_complex_set(gsl_vector_complex* v, size_t i, gsl_complex z) {
  v->data[2 * i * v->stride] = z.dat[0]; v->data[2 * i * v->stride + 1] = z.dat[1];
}
static inline gsl_complex* gsl_vector_complex_ptr(gsl_vector_complex* v, size_t i) { return (gsl_complex*)(v->data + 2 * i * v->stride); }
static inline void gsl_vector_complex_set_all(gsl_vector_complex* v, gsl_complex z) { for (size_t i = 0; i < v->size; i++) gsl_vector_complex_set(v, i, z); }
static inline int gsl_vector_complex_add(gsl_vector_complex* a, const gsl_vector_complex* b) {
  for (size_t i = 0; i < a->size; i++) gsl_vector_complex_set(a, i, gsl_complex_add(gsl_vector_complex_get(a, i), gsl_vector_complex_get(b, i)));
  return GSL_SUCCESS;
}
static inline int gsl_vector_complex_sub(gsl_vector_complex* a, const gsl_vector_complex* b) {
  for (size_t i = 0; i < a->size; i++) gsl_vector_complex_set(a, i, gsl_complex_sub(gsl_vector_complex_get(a, i), gsl_vector_complex_get(b, i)));
  return GSL_SUCCESS;
}
static inline int gsl_vector_complex_scale(gsl_vector_complex* a, gsl_complex x) {
  for (size_t i = 0; i < a->size; i++) gsl_vector_complex_set(a, i, gsl_complex_mul(gsl_vector_complex_get(a, i), x));
  return GSL_SUCCESS;
}
static inline gsl_complex gsl_matrix_complex_get(const gsl_matrix_complex* m, size_t i, size_t j) {
  return gsl_complex_rect(m->data[2 * (i * m->tda + j)], m->data[2 * (i * m->tda + j) + 1]);
}
static inline void gsl_matrix_complex_set(gsl_matrix_complex* m, size_t i, size_t j, gsl_complex z) {
  m->data[2 * (i * m->tda + j)] = z.dat[0]; m->data[2 * (i * m->tda + j) + 1] = z.dat[1];
}
static inline void gsl_matrix_complex_set_all(gsl_matrix_complex* m, gsl_complex z) {
  for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) gsl_matrix_complex_set(m, i, j, z);
}
static inline void gsl_matrix_complex_set_identity(gsl_matrix_complex* m) {
  for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) gsl_matrix_complex_set(m, i, j, gsl_complex_rect(i == j ? 1.0 : 0.0, 0.0));
}
static inline int gsl_matrix_complex_scale(gsl_matrix_complex* m, gsl_complex x) {
  for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) gsl_matrix_complex_set(m, i, j, gsl_complex_mul(gsl_matrix_complex_get(m, i, j), x));
  return GSL_SUCCESS;
}
static inline int gsl_matrix_complex_add(gsl_matrix_complex* a, const gsl_matrix_complex* b) {
  for (size_t i = 0; i < a->size1; i++) for (size_t j = 0; j < a->size2; j++) gsl_matrix_complex_set(a, i, j, gsl_complex_add(gsl_matrix_complex_get(a, i, j), gsl_matrix_complex_get(b, i, j)));
  return GSL_SUCCESS;
}
static inline int gsl_matrix_complex_sub(gsl_matrix_complex* a, const gsl_matrix_complex* b) {
  for (size_t i = 0; i < a->size1; i++) for (size_t j = 0; j < a->size2; j++) gsl_matrix_complex_set(a, i, j, gsl_complex_sub(gsl_matrix_complex_get(a, i, j), gsl_matrix_complex_get(b, i, j)));
  return GSL_SUCCESS;
}
static inline int gsl_matrix_complex_get_row(gsl_vector_complex* v, const gsl_matrix_complex* m, size_t i) {
  for (size_t j = 0; j < m->size2; j++) gsl_vector_complex_set(v, j, gsl_matrix_complex_get(m, i, j)); return GSL_SUCCESS;
}
static inline int gsl_matrix_complex_get_col(gsl_vector_complex* v, const gsl_matrix_complex* m, size_t j) {
  for (size_t i = 0; i < m->size1; i++) gsl_vector_complex_set(v, i, gsl_matrix_complex_get(m, i, j)); return GSL_SUCCESS;
}
static inline int gsl_matrix_complex_set_row(gsl_matrix_complex* m, size_t i, const gsl_vector_complex* v) {
  for (size_t j = 0; j < m->size2; j++) gsl_matrix_complex_set(m, i, j, gsl_vector_complex_get(v, j)); return GSL_SUCCESS;
}
static inline int gsl_matrix_complex_set_col(gsl_matrix_complex* m, size_t j, const gsl_vector_complex* v) {
  for (size_t i = 0; i < m->size1; i++) gsl_matrix_complex_set(m, i, j, gsl_vector_complex_get(v, i)); return GSL_SUCCESS;
}
static inline int gsl_matrix_complex_transpose_memcpy(gsl_matrix_complex* d, const gsl_matrix_complex* s) {
  for (size_t i = 0; i < s->size1; i++) for (size_t j = 0; j < s->size2; j++) gsl_matrix_complex_set(d, j, i, gsl_matrix_complex_get(s, i, j)); return GSL_SUCCESS;
}

/* views (only the trivially needed ones) */
typedef struct { gsl_vector vector; } gsl_vector_view;
typedef struct { gsl_vector vector; } gsl_vector_const_view;
typedef struct { gsl_vector_complex vector; } gsl_vector_complex_view;
typedef struct { gsl_matrix matrix; } gsl_matrix_view;
typedef struct { gsl_matrix_complex matrix; } gsl_matrix_complex_view;
static inline gsl_vector_view gsl_vector_view_array(double* base, size_t n) {
  gsl_vector_view v; v.vector.size = n; v.vector.stride = 1; v.vector.data = base; v.vector.block = 0; v.vector.owner = 0; return v;
}
static inline gsl_vector_complex_view gsl_matrix_complex_row(gsl_matrix_complex* m, size_t i) {
  gsl_vector_complex_view v; v.vector.size = m->size2; v.vector.stride = 1; v.vector.data = m->data + 2 * i * m->tda; v.vector.block = 0; v.vector.owner = 0; return v;
}
static inline gsl_vector_complex_view gsl_matrix_complex_column(gsl_matrix_complex* m, size_t j) {
  gsl_vector_complex_view v; v.vector.size = m->size1; v.vector.stride = m->tda; v.vector.data = m->data + 2 * j; v.vector.block = 0; v.vector.owner = 0; return v;
}

/* ---------------------------------------------------------------- CBLAS enums + the BLAS calls used */
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
enum CBLAS_UPLO { CblasUpper = 121, CblasLower = 122 };
enum CBLAS_DIAG { CblasNonUnit = 131, CblasUnit = 132 };
enum CBLAS_SIDE { CblasLeft = 141, CblasRight = 142 };
typedef enum CBLAS_TRANSPOSE CBLAS_TRANSPOSE_t;
typedef enum CBLAS_UPLO CBLAS_UPLO_t;
typedef enum CBLAS_DIAG CBLAS_DIAG_t;
typedef enum CBLAS_SIDE CBLAS_SIDE_t;

/* dotc = conj(x)^T y ; dotu = x^T y */
static inline int gsl_blas_zdotc(const gsl_vector_complex* x, const gsl_vector_complex* y, gsl_complex* dotc) {
  double re = 0, im = 0;
  for (size_t i = 0; i < x->size; i++) {
    gsl_complex a = gsl_vector_complex_get(x, i), b = gsl_vector_complex_get(y, i);
    re += a.dat[0] * b.dat[0] + a.dat[1] * b.dat[1];
    im += a.dat[0] * b.dat[1] - a.dat[1] * b.dat[0];
  }
  dotc->dat[0] = re; dotc->dat[1] = im; return GSL_SUCCESS;
}
static inline int gsl_blas_zdotu(const gsl_vector_complex* x, const gsl_vector_complex* y, gsl_complex* dotu) {
  double re = 0, im = 0;
  for (size_t i = 0; i < x->size; i++) {
    gsl_complex a = gsl_vector_complex_get(x, i), b = gsl_vector_complex_get(y, i);
    re += a.dat[0] * b.dat[0] - a.dat[1] * b.dat[1];
    im += a.dat[0] * b.dat[1] + a.dat[1] * b.dat[0];
  }
  dotu->dat[0] = re; dotu->dat[1] = im; return GSL_SUCCESS;
}
static inline double gsl_blas_dznrm2(const gsl_vector_complex* x) {
  double s = 0; for (size_t i = 0; i < x->size; i++) s += gsl_complex_abs2(gsl_vector_complex_get(x, i)); return sqrt(s);
}
static inline double gsl_blas_dnrm2(const gsl_vector* x) { double s = 0; for (size_t i = 0; i < x->size; i++) s += x->data[i * x->stride] * x->data[i * x->stride]; return sqrt(s); }
static inline int gsl_blas_ddot(const gsl_vector* x, const gsl_vector* y, double* r) { double s = 0; for (size_t i = 0; i < x->size; i++) s += x->data[i * x->stride] * y->data[i * y->stride]; *r = s; return GSL_SUCCESS; }
static inline void gsl_blas_zscal(const gsl_complex a, gsl_vector_complex* x) { gsl_vector_complex_scale(x, a); }
static inline void gsl_blas_zdscal(double a, gsl_vector_complex* x) { for (size_t i = 0; i < x->size; i++) gsl_vector_complex_set(x, i, gsl_complex_mul_real(gsl_vector_complex_get(x, i), a)); }
static inline void gsl_blas_dscal(double a, gsl_vector* x) { for (size_t i = 0; i < x->size; i++) x->data[i * x->stride] *= a; }
static inline int gsl_blas_zaxpy(const gsl_complex a, const gsl_vector_complex* x, gsl_vector_complex* y) {
  for (size_t i = 0; i < x->size; i++) gsl_vector_complex_set(y, i, gsl_complex_add(gsl_vector_complex_get(y, i), gsl_complex_mul(a, gsl_vector_complex_get(x, i))));
  return GSL_SUCCESS;
}
static inline int gsl_blas_daxpy(double a, const gsl_vector* x, gsl_vector* y) { for (size_t i = 0; i < x->size; i++) y->data[i * y->stride] += a * x->data[i * x->stride]; return GSL_SUCCESS; }

/* ---------------------------------------------------------------- real dense algebra used by modulated/prototypeDesign.cc
 * (views :283-323, dgemv/dgemm :346-548, SVD :276-300).  Textbook operations; the SVD is a one-sided Jacobi (Hestenes)
 * iteration in double with the singular values sorted in decreasing order like gsl_linalg_SV_decomp delivers them. */
#ifndef HUGE
#define HUGE HUGE_VAL
#endif
static inline gsl_matrix_view gsl_matrix_submatrix(gsl_matrix* m, size_t k1, size_t k2, size_t n1, size_t n2) {
  gsl_matrix_view v; v.matrix.size1 = n1; v.matrix.size2 = n2; v.matrix.tda = m->tda; v.matrix.data = m->data + k1 * m->tda + k2;
  v.matrix.block = m->block; v.matrix.owner = 0; return v; }
static inline gsl_vector_view gsl_vector_subvector(gsl_vector* x, size_t k, size_t n) {
  gsl_vector_view v; v.vector.size = n; v.vector.stride = x->stride; v.vector.data = x->data + k * x->stride;
  v.vector.block = x->block; v.vector.owner = 0; return v; }
static inline int gsl_blas_dgemv(CBLAS_TRANSPOSE_t t, double alpha, const gsl_matrix* A, const gsl_vector* x, double beta, gsl_vector* y) {
  const size_t rows = (t == CblasNoTrans) ? A->size1 : A->size2, cols = (t == CblasNoTrans) ? A->size2 : A->size1;
  double* tmp = (double*)malloc(sizeof(double) * (rows ? rows : 1));
  for (size_t i = 0; i < rows; i++) {
    double acc = 0;
    for (size_t j = 0; j < cols; j++) acc += ((t == CblasNoTrans) ? A->data[i * A->tda + j] : A->data[j * A->tda + i]) * x->data[j * x->stride];
    tmp[i] = alpha * acc + (beta != 0.0 ? beta * y->data[i * y->stride] : 0.0);
  }
  for (size_t i = 0; i < rows; i++) y->data[i * y->stride] = tmp[i];
  free(tmp); return GSL_SUCCESS; }
static inline int gsl_blas_dgemm(CBLAS_TRANSPOSE_t ta, CBLAS_TRANSPOSE_t tb, double alpha, const gsl_matrix* A, const gsl_matrix* B, double beta, gsl_matrix* C) {
  const size_t n1 = C->size1, n2 = C->size2, kk = (ta == CblasNoTrans) ? A->size2 : A->size1;
  double* tmp = (double*)malloc(sizeof(double) * (n1 * n2 ? n1 * n2 : 1));
  for (size_t i = 0; i < n1; i++) for (size_t j = 0; j < n2; j++) {
    double acc = 0;
    for (size_t k = 0; k < kk; k++)
      acc += ((ta == CblasNoTrans) ? A->data[i * A->tda + k] : A->data[k * A->tda + i]) * ((tb == CblasNoTrans) ? B->data[k * B->tda + j] : B->data[j * B->tda + k]);
    tmp[i * n2 + j] = alpha * acc + (beta != 0.0 ? beta * C->data[i * C->tda + j] : 0.0);
  }
  for (size_t i = 0; i < n1; i++) for (size_t j = 0; j < n2; j++) C->data[i * C->tda + j] = tmp[i * n2 + j];
  free(tmp); return GSL_SUCCESS; }
/* A (M x N, M >= N) = U S V^T : A is overwritten by U, V is N x N, S holds the singular values in decreasing order. */
static inline int gsl_linalg_SV_decomp(gsl_matrix* A, gsl_matrix* V, gsl_vector* S, gsl_vector* work) {
  (void)work;
  const size_t M = A->size1, N = A->size2;
  for (size_t i = 0; i < N; i++) for (size_t j = 0; j < N; j++) V->data[i * V->tda + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (size_t p = 0; p + 1 < N; p++) for (size_t q = p + 1; q < N; q++) {
      double a = 0, b = 0, g = 0;
      for (size_t i = 0; i < M; i++) { const double x = A->data[i * A->tda + p], y = A->data[i * A->tda + q]; a += x * x; b += y * y; g += x * y; }
      if (g == 0.0 || a == 0.0 || b == 0.0) continue;
      const double rel = fabs(g) / sqrt(a * b);
      if (rel > off) off = rel;
      if (rel < 1e-15) continue;
      const double zeta = (b - a) / (2.0 * g), tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      const double c = 1.0 / sqrt(1.0 + tt * tt), sn = c * tt;
      for (size_t i = 0; i < M; i++) { const double x = A->data[i * A->tda + p], y = A->data[i * A->tda + q]; A->data[i * A->tda + p] = c * x - sn * y; A->data[i * A->tda + q] = sn * x + c * y; }
      for (size_t i = 0; i < N; i++) { const double x = V->data[i * V->tda + p], y = V->data[i * V->tda + q]; V->data[i * V->tda + p] = c * x - sn * y; V->data[i * V->tda + q] = sn * x + c * y; }
    }
    if (off < 1e-15) break;
  }
  double svmax = 0;
  for (size_t j = 0; j < N; j++) {
    double n2 = 0; for (size_t i = 0; i < M; i++) n2 += A->data[i * A->tda + j] * A->data[i * A->tda + j];
    const double sv = sqrt(n2); S->data[j * S->stride] = sv;
    if (sv > svmax) svmax = sv;
  }
  /* U.  GSL delivers a U with orthonormal columns whatever the rank (its U is a product of orthogonal transformations;
   * prototypeDesign.cc:280-287 relies on it: for a wide matrix it decomposes the transpose and takes this U as "V", whose
   * trailing columns are the null space).  Hestenes' u_j = g_j / s_j is rounding noise for a vanishing s_j, so those
   * columns are rebuilt: unit vectors, orthogonalised twice against every other column. */
  std::vector<size_t> nullcols;
  for (size_t j = 0; j < N; j++) {
    const double sv = S->data[j * S->stride];
    if (sv > 1e-13 * svmax && sv > 0) for (size_t i = 0; i < M; i++) A->data[i * A->tda + j] /= sv;
    else { nullcols.push_back(j); for (size_t i = 0; i < M; i++) A->data[i * A->tda + j] = 0.0; }
  }
  if (!nullcols.empty() && M >= N) {
    size_t cand = 0;
    std::vector<double> v(M);
    for (size_t q = 0; q < nullcols.size(); q++) {
      const size_t j = nullcols[q];
      for (; cand < M; cand++) {
        for (size_t i = 0; i < M; i++) v[i] = (i == cand) ? 1.0 : 0.0;
        for (int pass = 0; pass < 2; pass++)
          for (size_t k = 0; k < N; k++) {
            if (k == j) continue;
            double d = 0; for (size_t i = 0; i < M; i++) d += v[i] * A->data[i * A->tda + k];
            if (d != 0.0) for (size_t i = 0; i < M; i++) v[i] -= d * A->data[i * A->tda + k];
          }
        double n2 = 0; for (size_t i = 0; i < M; i++) n2 += v[i] * v[i];
        if (n2 > 0.25) { const double nn = sqrt(n2); for (size_t i = 0; i < M; i++) A->data[i * A->tda + j] = v[i] / nn; cand++; break; }
      }
    }
  }
  for (size_t j = 0; j + 1 < N; j++) {          /* selection sort, decreasing */
    size_t best = j;
    for (size_t k = j + 1; k < N; k++) if (S->data[k * S->stride] > S->data[best * S->stride]) best = k;
    if (best != j) {
      const double t = S->data[j * S->stride]; S->data[j * S->stride] = S->data[best * S->stride]; S->data[best * S->stride] = t;
      for (size_t i = 0; i < M; i++) { const double x = A->data[i * A->tda + j]; A->data[i * A->tda + j] = A->data[i * A->tda + best]; A->data[i * A->tda + best] = x; }
      for (size_t i = 0; i < N; i++) { const double x = V->data[i * V->tda + j]; V->data[i * V->tda + j] = V->data[i * V->tda + best]; V->data[i * V->tda + best] = x; }
    }
  }
  return GSL_SUCCESS; }
static inline gsl_complex btkshim_op(const gsl_matrix_complex* A, CBLAS_TRANSPOSE_t t, size_t i, size_t j) {
  if (t == CblasNoTrans) return gsl_matrix_complex_get(A, i, j);
  gsl_complex z = gsl_matrix_complex_get(A, j, i);
  return t == CblasConjTrans ? gsl_complex_conjugate(z) : z;
}
/* y = alpha op(A) x + beta y */
static inline int gsl_blas_zgemv(CBLAS_TRANSPOSE_t t, const gsl_complex alpha, const gsl_matrix_complex* A,
                                 const gsl_vector_complex* x, const gsl_complex beta, gsl_vector_complex* y) {
  size_t rows = (t == CblasNoTrans) ? A->size1 : A->size2, cols = (t == CblasNoTrans) ? A->size2 : A->size1;
  std::vector<gsl_complex> out(rows);
  for (size_t i = 0; i < rows; i++) {
    gsl_complex s = gsl_complex_rect(0, 0);
    for (size_t j = 0; j < cols; j++) s = gsl_complex_add(s, gsl_complex_mul(btkshim_op(A, t, i, j), gsl_vector_complex_get(x, j)));
    out[i] = gsl_complex_add(gsl_complex_mul(alpha, s), gsl_complex_mul(beta, gsl_vector_complex_get(y, i)));
  }
  for (size_t i = 0; i < rows; i++) gsl_vector_complex_set(y, i, out[i]);
  return GSL_SUCCESS;
}
/* C = alpha op(A) op(B) + beta C */
static inline int gsl_blas_zgemm(CBLAS_TRANSPOSE_t ta, CBLAS_TRANSPOSE_t tb, const gsl_complex alpha, const gsl_matrix_complex* A,
                                 const gsl_matrix_complex* B, const gsl_complex beta, gsl_matrix_complex* C) {
  size_t M = C->size1, N = C->size2, K = (ta == CblasNoTrans) ? A->size2 : A->size1;
  std::vector<gsl_complex> out(M * N);
  for (size_t i = 0; i < M; i++) for (size_t j = 0; j < N; j++) {
    gsl_complex s = gsl_complex_rect(0, 0);
    for (size_t k = 0; k < K; k++) s = gsl_complex_add(s, gsl_complex_mul(btkshim_op(A, ta, i, k), btkshim_op(B, tb, k, j)));
    out[i * N + j] = gsl_complex_add(gsl_complex_mul(alpha, s), gsl_complex_mul(beta, gsl_matrix_complex_get(C, i, j)));
  }
  for (size_t i = 0; i < M; i++) for (size_t j = 0; j < N; j++) gsl_matrix_complex_set(C, i, j, out[i * N + j]);
  return GSL_SUCCESS;
}
/* A += alpha x y^T */
static inline int gsl_blas_zgeru(const gsl_complex alpha, const gsl_vector_complex* x, const gsl_vector_complex* y, gsl_matrix_complex* A) {
  for (size_t i = 0; i < A->size1; i++) for (size_t j = 0; j < A->size2; j++)
    gsl_matrix_complex_set(A, i, j, gsl_complex_add(gsl_matrix_complex_get(A, i, j),
        gsl_complex_mul(alpha, gsl_complex_mul(gsl_vector_complex_get(x, i), gsl_vector_complex_get(y, j)))));
  return GSL_SUCCESS;
}
/* A += alpha x y^H */
static inline int gsl_blas_zgerc(const gsl_complex alpha, const gsl_vector_complex* x, const gsl_vector_complex* y, gsl_matrix_complex* A) {
  for (size_t i = 0; i < A->size1; i++) for (size_t j = 0; j < A->size2; j++)
    gsl_matrix_complex_set(A, i, j, gsl_complex_add(gsl_matrix_complex_get(A, i, j),
        gsl_complex_mul(alpha, gsl_complex_mul(gsl_vector_complex_get(x, i), gsl_complex_conjugate(gsl_vector_complex_get(y, j))))));
  return GSL_SUCCESS;
}

/* ---------------------------------------------------------------- special functions */
/* normalised sinc: sin(pi x)/(pi x), sinc(0)=1 (gsl_sf_sinc's definition) */
static inline double gsl_sf_sinc(double x) {
  double ax = fabs(x);
  if (ax < 1e-8) return 1.0 - (M_PI * M_PI * x * x) / 6.0;
  return sin(M_PI * x) / (M_PI * x);
}

/* ---------------------------------------------------------------- FFT (packed complex double, radix-2) */
typedef enum { gsl_fft_forward = -1, gsl_fft_backward = +1 } gsl_fft_direction;
typedef double* gsl_complex_packed_array;
static inline int btkshim_fft_radix2(double* data, size_t stride, size_t n, int sign) {
  if (n == 0 || (n & (n - 1))) { fprintf(stderr, "gsl shim: radix-2 FFT length %zu is not a power of 2\n", n); return GSL_EINVAL; }
  if (n == 1) return GSL_SUCCESS;
  /* bit reversal */
  for (size_t i = 0, j = 0; i < n - 1; i++) {
    if (i < j) {
      double tr = data[2 * stride * i], ti = data[2 * stride * i + 1];
      data[2 * stride * i] = data[2 * stride * j]; data[2 * stride * i + 1] = data[2 * stride * j + 1];
      data[2 * stride * j] = tr; data[2 * stride * j + 1] = ti;
    }
    size_t k = n >> 1;
    while (k <= j) { j -= k; k >>= 1; }
    j += k;
  }
  /* Decimation in time; per stage ONE sin() pair and the recurrence w <- w + (-s2 w + s (j w)) inside the stage, i.e. no
     cos/sin in the butterfly loops.  An earlier
     version of this stand-in called cos/sin per (stage, a), which made the reference's CPU chain look 1.5-2 x slower
     than it is with the real library. */
  for (size_t dual = 1; dual < n; dual <<= 1) {
    double wr = 1.0, wi = 0.0;
    const double theta = 2.0 * (double)sign * M_PI / (2.0 * (double)dual);
    const double sn = sin(theta), t = sin(theta / 2.0), s2 = 2.0 * t * t;
    for (size_t a = 0; a < dual; a++) {
      for (size_t b = a; b < n; b += 2 * dual) {
        size_t p = b, q = b + dual;
        double xr = data[2 * stride * q], xi = data[2 * stride * q + 1];
        double tr = wr * xr - wi * xi, ti = wr * xi + wi * xr;
        data[2 * stride * q] = data[2 * stride * p] - tr; data[2 * stride * q + 1] = data[2 * stride * p + 1] - ti;
        data[2 * stride * p] += tr; data[2 * stride * p + 1] += ti;
      }
      { const double nr = wr - sn * wi - s2 * wr, ni = wi + sn * wr - s2 * wi; wr = nr; wi = ni; }
    }
  }
  return GSL_SUCCESS;
}
static inline int gsl_fft_complex_radix2_forward(gsl_complex_packed_array d, size_t stride, size_t n) { return btkshim_fft_radix2(d, stride, n, -1); }
static inline int gsl_fft_complex_radix2_backward(gsl_complex_packed_array d, size_t stride, size_t n) { return btkshim_fft_radix2(d, stride, n, +1); }
static inline int gsl_fft_complex_radix2_inverse(gsl_complex_packed_array d, size_t stride, size_t n) {
  int r = btkshim_fft_radix2(d, stride, n, +1);
  for (size_t i = 0; i < n; i++) { d[2 * stride * i] /= (double)n; d[2 * stride * i + 1] /= (double)n; }
  return r;
}
static inline int gsl_fft_complex_radix2_transform(gsl_complex_packed_array d, size_t stride, size_t n, gsl_fft_direction s) { return btkshim_fft_radix2(d, stride, n, (int)s); }

#endif /* BTKB200_GSL_SHIM_ALL_H */
