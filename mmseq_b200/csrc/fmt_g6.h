/* fmt_g6.h — "%g" at precision 6 (what operator<<(double) writes at the stream defaults: the reference's number format,
 * SURVEY.md appendix B) fast enough for the 4e8 values of the trace files.  tests/test_cli_args.py checks it against
 * printf's characters. */
#ifndef MMQ_FMT_G6_H
#define MMQ_FMT_G6_H
#include <stdint.h>
#include <string.h>

#include <charconv>

/* "%g" (precision 6) of a double, the characters printf / operator<< produce, without the general machinery: six significant
 * digits from one multiplication by an exact power of ten; anything unusual (non-finite, zero, negative, outside 1e-17..1e27,
 * or within 1e-7 of a rounding tie) goes to to_chars, which is specified to give printf's characters. */
static inline char* fmt_g6(char* buf, double v) {
  static const double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  if (!(v >= 1e-17 && v < 1e27)) return std::to_chars(buf, buf + 40, v, std::chars_format::general, 6).ptr;
  uint64_t bits;
  memcpy(&bits, &v, 8);
  const int e2 = (int)(bits >> 52) - 1023;
  int X = (e2 * 1233) >> 12; /* floor(e2 log10 2), within one of floor(log10 v) */
  if (e2 < 0) X = -(((-e2) * 1233 + 4095) >> 12);
  double w = X <= 5 ? v * P10[5 - X] : v / P10[X - 5];
  if (w < 1e5) { --X; w = X <= 5 ? v * P10[5 - X] : v / P10[X - 5]; }
  else if (w >= 1e6) { ++X; w = X <= 5 ? v * P10[5 - X] : v / P10[X - 5]; }
  if (!(w >= 1e5 && w < 1e6)) return std::to_chars(buf, buf + 40, v, std::chars_format::general, 6).ptr;
  uint32_t D = (uint32_t)w;
  const double fr = w - (double)D;
  if (fr > 0.5 - 1e-7) {
    if (fr < 0.5 + 1e-7) return std::to_chars(buf, buf + 40, v, std::chars_format::general, 6).ptr;
    if (++D == 1000000u) { D = 100000u; ++X; }
  }
  int nd = 6;
  while (nd > 1 && D % 10u == 0u) { D /= 10u; --nd; }
  char dg[8];
  for (int i = nd - 1; i >= 0; --i) { dg[i] = (char)('0' + D % 10u); D /= 10u; }
  char* p = buf;
  if (X < -4 || X >= 6) {
    *p++ = dg[0];
    if (nd > 1) { *p++ = '.'; for (int i = 1; i < nd; ++i) *p++ = dg[i]; }
    *p++ = 'e';
    int ax = X;
    if (X < 0) { *p++ = '-'; ax = -X; } else *p++ = '+';
    if (ax >= 100) { *p++ = (char)('0' + ax / 100); ax %= 100; }
    *p++ = (char)('0' + ax / 10);
    *p++ = (char)('0' + ax % 10);
  } else if (X >= 0) {
    const int ip = X + 1; /* digits in front of the point */
    for (int i = 0; i < ip; ++i) *p++ = i < nd ? dg[i] : '0';
    if (nd > ip) { *p++ = '.'; for (int i = ip; i < nd; ++i) *p++ = dg[i]; }
  } else {
    *p++ = '0'; *p++ = '.';
    for (int i = 0; i < -X - 1; ++i) *p++ = '0';
    for (int i = 0; i < nd; ++i) *p++ = dg[i];
  }
  return p;
}

#endif
