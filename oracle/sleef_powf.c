// TEST INFRASTRUCTURE (oracle/): scalar C restatement of Sleef's powf (u10, FMA build) — the function torch's CPU kernel
// runs for x.pow(p - 1) in the HQQ solver's shrink operator (/root/reference/amq/kernel/hqq/hqq/core/optimize.py:96-108
// through ATen's Vectorized<float>::pow -> Sleef_powf{8,16}_u10).  Sleef is a third-party dependency of torch that is not
// in /root/reference (torch 2.11 bundles Sleef 3.6); the published algorithm (sleefsimdsp.c: xpowf = expkf(logkf(|x|) * y)
// in double-float arithmetic) is restated here operation by operation and pinned bit for bit against torch.pow itself
// (tests/test_oracle_golden.py::test_sleef_powf_restatement_matches_torch, 4 M inputs).  The CUDA solver
// (amq_b200/csrc/hqq_quant.cu sleef_powf_u10) carries the same sequence with __fmaf_rn / __fmul_rn / __fadd_rn / __fdiv_rn.
// Build: gcc -O2 -mfma -ffp-contract=off -shared -fPIC -o oracle/_ref/libsleefpow.so oracle/sleef_powf.c -lm
#include <math.h>
#include <stdint.h>
#include <string.h>
typedef struct { float x, y; } f2;
static inline float fmapn(float a, float b, float c) { return fmaf(a, b, -c); }   // a*b - c
static inline float fmanp(float a, float b, float c) { return fmaf(-a, b, c); }   // -a*b + c
static inline f2 dfadd2_f_f(float x, float y) { f2 r; r.x = x + y; float v = r.x - x; r.y = (x - (r.x - v)) + (y - v); return r; }
static inline f2 dfadd2_f2_f(f2 x, float y) { f2 r; r.x = x.x + y; float v = r.x - x.x; r.y = (x.x - (r.x - v)) + (y - v); r.y = r.y + x.y; return r; }
static inline f2 dfadd_f2_f2(f2 x, f2 y) { f2 r; r.x = x.x + y.x; r.y = x.x - r.x + y.x + x.y + y.y; return r; }
static inline f2 dfadd2_f2_f2(f2 x, f2 y) { f2 r; r.x = x.x + y.x; float v = r.x - x.x; r.y = (x.x - (r.x - v)) + (y.x - v); r.y = r.y + (x.y + y.y); return r; }
static inline f2 dfadd_f_f2(float x, f2 y) { f2 r; r.x = x + y.x; r.y = x - r.x + y.x + y.y; return r; }
static inline f2 dfmul_f2_f(f2 x, float y) { f2 r; r.x = x.x * y; r.y = fmaf(x.y, y, fmapn(x.x, y, r.x)); return r; }
static inline f2 dfmul_f2_f2(f2 x, f2 y) { f2 r; r.x = x.x * y.x; r.y = fmaf(x.x, y.y, fmaf(x.y, y.x, fmapn(x.x, y.x, r.x))); return r; }
static inline f2 dfsqu(f2 x) { f2 r; r.x = x.x * x.x; r.y = fmaf(x.x + x.x, x.y, fmapn(x.x, x.x, r.x)); return r; }
static inline f2 dfdiv(f2 n, f2 d) {
  float t = 1.0f / d.x; float s = n.x * t; float u = fmapn(t, n.x, s);
  float v = fmanp(d.y, t, fmanp(d.x, t, 1.0f));
  f2 r; r.x = s; r.y = fmaf(s, v, fmaf(n.y, t, u)); return r;
}
static inline f2 dfscale(f2 x, float s) { f2 r = {x.x * s, x.y * s}; return r; }
static inline f2 dfnorm(f2 t) { f2 s; s.x = t.x + t.y; s.y = t.x - s.x + t.y; return s; }
static f2 logkf(float d) {
  // getexp(d * (1/0.75)), getmant in [0.75, 1.5)
  float dd = d * (1.0f / 0.75f);
  int e; frexpf(dd, &e); e -= 1;            // floor(log2(dd))
  float m = ldexpf(d, -e);
  f2 x = dfdiv(dfadd2_f_f(-1.0f, m), dfadd2_f_f(1.0f, m));
  f2 x2 = dfsqu(x);
  float t = 0.240320354700088500976562f;
  t = fmaf(t, x2.x, 0.285112679004669189453125f);
  t = fmaf(t, x2.x, 0.400007992982864379882812f);
  f2 c = {0.66666662693023681640625f, 3.69183861259614332084311e-09f};
  f2 ln2 = {0.69314718246459960938f, -1.904654323148236017e-09f};
  f2 s = dfmul_f2_f(ln2, (float)e);
  s = dfadd_f2_f2(s, dfscale(x, 2.0f));
  s = dfadd_f2_f2(s, dfmul_f2_f2(dfmul_f2_f2(x2, x), dfadd2_f2_f2(dfmul_f2_f(x2, t), c)));
  return s;
}
static float expkf(f2 d) {
  float u = (d.x + d.y) * 1.442695040888963407359924681001892137426645954152985934135449406931f;
  int q = (int)rintf(u);
  f2 s = dfadd2_f2_f(d, (float)q * -0.693145751953125f);
  s = dfadd2_f2_f(s, (float)q * -1.428606765330187045e-06f);
  s = dfnorm(s);
  u = 0.00136324646882712841033936f;
  u = fmaf(u, s.x, 0.00836596917361021041870117f);
  u = fmaf(u, s.x, 0.0416710823774337768554688f);
  u = fmaf(u, s.x, 0.166665524244308471679688f);
  u = fmaf(u, s.x, 0.499999850988388061523438f);
  f2 t = dfadd_f2_f2(s, dfmul_f2_f(dfsqu(s), u));
  t = dfadd_f_f2(1.0f, t);
  u = t.x + t.y;
  u = ldexpf(u, q);
  if (d.x < -104.0f) u = 0.0f;
  return u;
}
void sleef_powf(const float* x, float y, float* out, long n) {
  for (long i = 0; i < n; ++i) {
    f2 l = logkf(fabsf(x[i]));
    f2 d = dfmul_f2_f(l, y);
    out[i] = expkf(d);
  }
}
