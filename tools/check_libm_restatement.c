// Checks the restatement of glibc sinf/cosf/expf used by the cpu-exact mode (csrc/common.cuh: glibc_sincosf, glibc_expf) against the libm of this image:
//   gcc -O2 -mfma -ffp-contract=off -o /tmp/chk tools/check_libm_restatement.c -lm && /tmp/chk     (expected: 0 mismatches)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
typedef struct { double sign[4]; double hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3; } sincos_t;
static const sincos_t T2[2] = {
  {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
  {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};
static const uint32_t inv_pio4[24] = {0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529, 0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1,
  0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, 0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};
static inline uint32_t asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline uint32_t abstop12(float x) { return (asuint(x) >> 20) & 0x7ff; }
static inline float sinf_poly(double x, double x2, const sincos_t *p, int n) {
  if ((n & 1) == 0) { double x3 = x * x2, s1 = fma(x2, p->s3, p->s2), x7 = x3 * x2, s = fma(x3, p->s1, x); return (float)fma(x7, s1, s); }
  double x4 = x2 * x2, c2 = fma(x2, p->c4, p->c3), c1 = fma(x2, p->c1, p->c0), x6 = x4 * x2, c = fma(x4, p->c2, c1); return (float)fma(x6, c2, c);
}
static inline double reduce_fast(double x, const sincos_t *p, int *np) { double r = x * p->hpi_inv; int n = ((int32_t)r + 0x800000) >> 24; *np = n; return fma(-(double)n, p->hpi, x); }
static inline double reduce_large(uint32_t xi, int *np) {
  const uint32_t *arr = &inv_pio4[(xi >> 26) & 15]; int shift = (xi >> 23) & 7; uint64_t n, res0, res1, res2;
  xi = (xi & 0xffffff) | 0x800000; xi <<= shift; res0 = xi * arr[0]; res1 = (uint64_t)xi * arr[4]; res2 = (uint64_t)xi * arr[8];
  res0 = (res2 >> 32) | (res0 << 32); res0 += res1; n = (res0 + (1ULL << 61)) >> 62; res0 -= n << 62; double x = (int64_t)res0; *np = (int)n; return x * 0x1.921FB54442D18p-62;
}
float my_sinf(float y) {
  double x = y, s; int n; const sincos_t *p = &T2[0];
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) { s = x * x; if (abstop12(y) < abstop12(0x1p-12f)) return y; return sinf_poly(x, s, p, 0); }
  else if (abstop12(y) < abstop12(120.0f)) { x = reduce_fast(x, p, &n); s = p->sign[n & 3]; if (n & 2) p = &T2[1]; return sinf_poly(x * s, x * x, p, n); }
  else { uint32_t xi = asuint(y); int sign = xi >> 31; x = reduce_large(xi, &n); s = p->sign[(n + sign) & 3]; if ((n + sign) & 2) p = &T2[1]; return sinf_poly(x * s, x * x, p, n); }
}
float my_cosf(float y) {
  double x = y, s; int n; const sincos_t *p = &T2[0];
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) { double x2 = x * x; if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f; return sinf_poly(x, x2, p, 1); }
  else if (abstop12(y) < abstop12(120.0f)) { x = reduce_fast(x, p, &n); s = p->sign[n & 3]; if (n & 2) p = &T2[1]; return sinf_poly(x * s, x * x, p, n ^ 1); }
  else { uint32_t xi = asuint(y); int sign = xi >> 31; x = reduce_large(xi, &n); s = p->sign[(n + sign) & 3]; if ((n + sign) & 2) p = &T2[1]; return sinf_poly(x * s, x * x, p, n ^ 1); }
}
static uint64_t TT[32];
float my_expf(float x) {
  double xd = x; uint32_t abstop = (asuint(x) >> 20) & 0x7ff;
  if (abstop >= ((asuint(88.0f) >> 20))) return expf(x);   // special range: not needed by softmax (args <= 0) beyond underflow
  const double IL = 0x1.71547652b82fep+0 * 32; double kd = fma(IL, xd, 0x1.8p+52); uint64_t ki; memcpy(&ki, &kd, 8); kd -= 0x1.8p+52; double r = fma(IL, xd, -kd); double z;
  uint64_t t = TT[ki % 32]; t += ki << (52 - 5); double s; memcpy(&s, &t, 8);
  const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32, C2 = 0x1.62e42ff0c52d6p-1 / 32;
  z = fma(C0, r, C1); double r2 = r * r; double y = fma(C2, r, 1.0); y = fma(z, r2, y); y = y * s; return (float)y;
}
int main() {
  for (int i = 0; i < 32; i++) { double v = exp2((double)i / 32); uint64_t u; memcpy(&u, &v, 8); TT[i] = u - ((uint64_t)i << 47); }
  uint64_t st = 88172645463325252ull; long bad_s = 0, bad_c = 0, bad_e = 0, n = 0;
  for (long it = 0; it < 60000000; it++) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    float scale = (it % 4 == 0) ? 1.0f : (it % 4 == 1) ? 100.0f : (it % 4 == 2) ? 5000.0f : 140000.0f;
    float x = (float)((double)(st >> 11) / 9007199254740992.0 * 2 - 1) * scale;
    if (my_sinf(x) != sinf(x)) { if (bad_s < 3) printf("sin %a: %a vs %a\n", x, my_sinf(x), sinf(x)); bad_s++; }
    if (my_cosf(x) != cosf(x)) { if (bad_c < 3) printf("cos %a: %a vs %a\n", x, my_cosf(x), cosf(x)); bad_c++; }
    float xe = -fabsf(x) * (80.0f / scale);
    if (my_expf(xe) != expf(xe)) { if (bad_e < 3) printf("exp %a: %a vs %a\n", xe, my_expf(xe), expf(xe)); bad_e++; }
    n++;
  }
  printf("n=%ld mismatches sin %ld cos %ld exp %ld\n", n, bad_s, bad_c, bad_e);
  long bad_rn = 0; st = 1234567;
  for (long it = 0; it < 20000000; it++) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; float x = (float)((double)(st >> 11) / 9007199254740992.0) * 5000.0f;
    if ((float)sin((double)x) != sinf(x)) bad_rn++; }
  printf("(float)sin(double) vs sinf mismatches: %ld of 20M\n", bad_rn);
  return 0;
}
