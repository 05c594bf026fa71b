// Host harness for zigzag_b200/csrc/zz_fastmath.cuh (the functions are __host__ __device__): prints the worst error of the
// table-driven exp / log against long-double libm over a seeded sample.  Built and run by tests/test_fastmath.py.
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <random>
struct double2 { double x, y; };
#include "zz_fastmath.cuh"
using namespace zz;
static double ulp_err(double got, long double ref) {
  double r = (double)ref; int e; frexp(r, &e);
  long double ulp = ldexpl(1.0L, e - 53);
  return (double)fabsl(((long double)got - ref) / ulp);
}
int main(int argc, char **argv) {
  long n = argc > 1 ? atol(argv[1]) : 2000000;
  static MathTabs T;
  for (int j = 0; j < 256; ++j) memcpy(&T.exp_t[j], &kExpTab[j], 8);
  for (int j = 0; j < 512; ++j) { memcpy(&T.log_t[j].x, &kLogTab[2 * j], 8); memcpy(&T.log_t[j].y, &kLogTab[2 * j + 1], 8); }
  std::mt19937_64 g(1); std::uniform_real_distribution<double> U(0, 1);
  double me = 0, ml = 0, mla = 0, me2 = 0, mp2 = 0, mrc = 0, msq = 0, menb = 0;
  for (long it = 0; it < n; ++it) {
    double x = (it & 1) ? (U(g) - 0.5) * 20 : (U(g) - 0.5) * 1416;
    double ue = ulp_err(fexp(x, T.exp_t), expl((long double)x)); if (ue > me) me = ue;
    {  // branch-free helpers
      double xr = ldexp(1 + U(g), (int)(U(g) * 600) - 300);
      double urc = ulp_err(frcp_nb(xr), 1.0L / (long double)xr); if (urc > mrc) mrc = urc;
      double xs = (it & 1) ? U(g) * 44.4 + 4.6e-10 : U(g) * 1e-6 + 4.6e-10;
      double usq = ulp_err(fsqrt_nb(xs), sqrtl((long double)xs)); if (usq > msq) msq = usq;
      double xe = (U(g) - 0.5) * 1400; double uenb = ulp_err(fexp_nb(xe, T.exp_t), expl((long double)xe)); if (uenb > menb) menb = uenb;
    }
    {  // base-2 detection exponential: 1 - 2^-b against 1 - exp(-a) with b = a * log2(e) (rounded)
      double a = (it % 4 == 0) ? U(g) * 60 : U(g) * U(g) * 4;
      double bb = a * ZZ_KLOG2E;
      double e2 = fexp2_neg_clamped(bb, T.exp_t);
      long double ref = exp2l(-(long double)bb);
      double u2 = ulp_err(e2, ref); if (u2 > me2) me2 = u2;
      double dp = fabs((1 - e2) - (double)(1 - expl(-(long double)a))); if (dp > mp2) mp2 = dp;
    }
    double y = ldexp(U(g) + 1e-9, (int)(U(g) * 120) - 60);
    if (it % 3 == 0) y = 1 - U(g) * 1e-3 * U(g);
    if (it % 3 == 1) y = 1 + U(g) * 1e-2;
    long double rl = logl((long double)y); double l = flog(y, T.log_t);
    if (fabsl(rl) > 0.01L) { double ul = ulp_err(l, rl); if (ul > ml) ml = ul; } else { double ae = (double)fabsl((long double)l - rl); if (ae > mla) mla = ae; }
  }
  printf("exp_max_ulp %.4f\nlog_max_ulp %.4f\nlog_max_abs_near_one %.4e\nexp2_max_ulp %.4f\npx_max_abs_err %.4e\n", me, ml, mla, me2, mp2);
  printf("rcp_max_ulp %.4f\nsqrt_max_ulp %.4f\nexp_nb_max_ulp %.4f\nexp_nb_m750 %.17g\nexp_nb_m720 %.17g\nexp_nb_800 %.17g\n", mrc, msq, menb, fexp_nb(-750, T.exp_t), fexp_nb(-720, T.exp_t), fexp_nb(800, T.exp_t));
  printf("log1 %.17g\nlog0 %.17g\nlogneg %.17g\nloginf %.17g\nlogsub %.17g %.17g\n", flog(1.0, T.log_t), flog(0.0, T.log_t), flog(-1.0, T.log_t), flog(INFINITY, T.log_t), flog(1e-310, T.log_t), log(1e-310));
  printf("exp_m800 %.17g\nexp_800 %.17g\nexp_m720 %.17g %.17g\nexp_nan %.17g\n", fexp(-800, T.exp_t), fexp(800, T.exp_t), fexp(-720, T.exp_t), exp(-720.), fexp(NAN, T.exp_t));
  const double a[] = {0, 1e-300, 1e-17, 0.5, 36.7, 36.8, 37.5, 40, 699, 700, 701, 1e5, INFINITY};
  int bad = 0;
  for (double v : a) bad += fabs((1 - fexp_neg_clamped(v, T.exp_t)) - (1 - exp(-v))) > 3e-16;
  for (double v : a) bad += fabs((1 - fexp2_neg_clamped(v * ZZ_KLOG2E, T.exp_t)) - (1 - exp(-v))) > 3e-16;
  printf("px_mismatch %d\n", bad);
  {  // det_q / det_log_product (straight-line tile) against the per-library path it replaces: sum_r log(q_r), q = undetected ? 1 - p : p
    double worst = 0, worst_q = 0; int zero_bad = 0;
    for (long it = 0; it < n / 16; ++it) {
      const int R = 2 + (int)(U(g) * 17);                  // 2..18 libraries
      double t = (it % 5 == 0) ? U(g) * 1e-6 : (it % 5 == 1) ? U(g) * 40 : U(g) * 3;
      long double want = 0; double P = 1; bool zero = false;
      for (int r = 0; r < R; ++r) {
        const double alpha = 0.5 + U(g) * 3, a2 = alpha * ZZ_KLOG2E;
        const unsigned nd = (U(g) < 0.2);
        const double tt = t * a2 > 1010 ? 1010 / a2 : t;
        const double e = fexp2_neg_clamped(tt * a2, T.exp_t), p = 1 - e, qref = nd ? 1 - p : p;
        const double q = det_q(tt, -256.0 * a2, nd, T);
        // the two exponent roundings differ by <= 1 ulp of e, i.e. <= 2^-53 absolute on q
        const double dq = fabs(q - qref) / 1.2e-16; if (dq > worst_q) worst_q = dq;
        if (q == 0) zero = true; else want += logl((long double)q);
        P *= q;
      }
      const double got = det_log_product(P, T.log_t);
      if (zero) { zero_bad += !(got == -INFINITY); continue; }
      zero_bad += !(got > -INFINITY);
      // R roundings of the product (2^-53 relative each) + the logarithm's own 2 ulp
      const double tol = (R + 1) * 1.2e-16 + 3e-16 * fabsl(want);
      const double err = (double)fabsl((long double)got - want) / tol; if (err > worst) worst = err;
    }
    printf("det_product_worst %.4f\ndet_q_worst %.4f\ndet_product_zero_bad %d\n", worst, worst_q, zero_bad);
  }
  return 0;
}
