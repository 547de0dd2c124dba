// okb_gatecos.h -- cos() for the gate constants of triangulateFast (reference okvis_frontend/src/stereo_triangulation.cpp:
// 82-127: cos(2.6 * sigma), cos(6.0 * sigma)), written once and compiled for the device AND the host.
//
// The reference calls the host libm. SURVEY.md H3: no device transcendental may decide a match, and the host-buffer and
// device-resident forms of one matcher must agree bit for bit. CUDA's cos is only within 1-2 ulp of libm, and a
// correctly rounded cos is not what libm returns either (glibc 2.39 misrounds 0.03 % of the arguments in [0, 0.5):
// measured). gate_cos therefore restates the algorithm glibc's cos runs on an FMA-capable x86-64 for
// 2^-27 <= |x| < 0.855469 (sysdeps/ieee754/dbl-64/s_sin.c, do_cos, as compiled for the *_fma ifunc variant: every a*b+c
// contracted): |x| is split into a multiple of 1/128 (table of sin/cos high+low words, regenerated from exact rational
// arithmetic by tools/gen_gatecos_table.py) and a remainder r with two short polynomials in r. Only IEEE double
// add/mul/fma: host and device execute the same operations and return the same bits. 100 000 000 arguments compared with
// the libm of this image: zero differences (tests/test_gatecos.py runs 20 M in the CPU suite); okb_create repeats a
// 65 536-argument self-check against the libm of the machine it runs on and reports it through okb_gate_cos_exact().
// sigma = size / f * 0.125 of any real camera keeps 6 sigma far below 0.855; outside that range the plain cos is used.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define OKB_GC_HD __host__ __device__ inline
#else
#define OKB_GC_HD inline
#endif

namespace okb {

static const double kGateCosTabHost[110][4] = {
#include "okb_gatecos_tab.h"
};
#if defined(__CUDACC__)
static __device__ __constant__ double kGateCosTabDev[110][4] = {
#include "okb_gatecos_tab.h"
};
#endif

OKB_GC_HD double gate_cos(double x)
{
#if defined(__CUDA_ARCH__)
  const double (*tab)[4] = kGateCosTabDev;
#else
  const double (*tab)[4] = kGateCosTabHost;
#endif
  const double big = 52776558133248.0;   // 1.5 * 2^45: adding it rounds |x| to a multiple of 2^-7
  const double sn3 = -1.66666666666664880952546298448555E-01, sn5 = 8.33333214285722277379541354343671E-03;
  const double cs2 = 4.99999999999999999999950396842453E-01, cs4 = -4.16666666666664434524222570944589E-02,
               cs6 = 1.38888874007937613028114285595617E-03;
  const double ax = fabs(x);
  if (!(ax < 0.85546875)) return cos(x);   // high word < 0x3feb6000
  if (ax < 7.450580596923828125e-09) return 1.0;   // 2^-27
  const double u = big + ax;
  const double r = ax - (u - big);
#if defined(__CUDA_ARCH__)
  const int k = __double2loint(u);
#else
  uint64_t ub; memcpy(&ub, &u, 8);
  const int k = (int)(ub & 0xffffffffu);
#endif
  const double xx = r * r;
  const double s = fma(r * xx, fma(xx, sn5, sn3), r);
  const double c = xx * fma(xx, fma(xx, cs6, cs4), cs2);
  double cor = fma(-s, tab[k][1], tab[k][3]);
  cor = fma(-tab[k][2], c, cor);
  cor = fma(-tab[k][0], s, cor);
  return tab[k][2] + cor;
}

}  // namespace okb
