/* check_const_div.c -- verifies the division-by-constant shortcut of track_hessian.cu (div169, div_h):
 *   float  x / 169.f   exhaustively over all 2^32 bit patterns,
 *   double x / 0.02 and x / 0.01 over 6.4e9 random inputs of the forms the tracker produces.
 * build: gcc -O2 -fopenmp -ffp-contract=off -o check_const_div tools/check_const_div.c -lm   (about a minute)
 * expected output: one float mismatch (x = -0, result +0 instead of -0; a sum of non-negative terms is never -0)
 * and zero double mismatches. */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static long check_float(void) {
  const float d = 169.f, y = 1.0f / 169.f;
  long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
  for (long long i = 0; i < (1LL << 32); ++i) {
    uint32_t u = (uint32_t)i;
    float x;
    memcpy(&x, &u, 4);
    if (isnan(x) || isinf(x)) continue;
    float ref = x / d, q = x * y, r = fmaf(-q, d, x), q2 = fmaf(r, y, q);
    if (memcmp(&ref, &q2, 4) != 0) {
      printf("float mismatch: x=%a ref=%a got=%a\n", x, ref, q2);
      bad++;
    }
  }
  return bad;
}

static inline uint64_t rng(uint64_t* s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }

static long check_double(void) {
  const double hs[2] = {0.02, 0.01};
  long bad = 0;
#pragma omp parallel reduction(+ : bad)
  {
    uint64_t s = 0x9E3779B97F4A7C15ULL ^ (uint64_t)(1 + omp_get_thread_num()) * 0xD1B54A32D192ED03ULL;
    for (long it = 0; it < 400000000L; ++it)
      for (int k = 0; k < 2; ++k) {
        const double h = hs[k], y = 1.0 / h;
        double x;
        int mode = it % 4;
        uint64_t r1 = rng(&s), r2 = rng(&s);
        if (mode == 0) { /* arbitrary doubles away from overflow / underflow */
          memcpy(&x, &r1, 8);
          if (isnan(x) || isinf(x) || fabs(x) > 1e300 || (fabs(x) < 1e-300 && x != 0)) continue;
        } else {
          uint32_t a = (uint32_t)r1, b = (uint32_t)r2;
          float fa, fb;
          memcpy(&fa, &a, 4);
          memcpy(&fb, &b, 4);
          if (isnan(fa) || isinf(fa) || isnan(fb) || isinf(fb)) continue;
          if (mode == 1) { /* two close scores */
            uint32_t b2 = a + (uint32_t)(r2 % 4096) - 2048;
            memcpy(&fb, &b2, 4);
            if (isnan(fb) || isinf(fb)) continue;
          }
          x = (double)fa - (double)fb;                 /* hessian.h:163-169 numerators */
          if (mode == 2) x *= 0.5;
          if (mode == 3) {                             /* second stage: difference of two quotients */
            x = ((double)fa) / h - ((double)fb) / h;
            if (isinf(x) || isnan(x)) continue;
          }
        }
        double ref = x / h, q = x * y, r = fma(-q, h, x), q2 = fma(r, y, q);
        if (memcmp(&ref, &q2, 8) != 0 && x != 0) bad++;
      }
  }
  return bad;
}

int main(void) {
  printf("float  x/169: %ld mismatches\n", check_float());
  printf("double x/h  : %ld mismatches\n", check_double());
  return 0;
}
