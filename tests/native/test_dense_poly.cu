// Host check of the dense Horner evaluator of csrc/common.cuh (what the quadrature loops of the element kernels use)
// against the generic monomial evaluator, for every instantiation the launchers pick.  Built and run by
// tests/test_host.py::test_dense_polynomial_evaluator (nvcc, host code only -- no GPU needed).
#include "common.cuh"
#include <cmath>
#include <cstdlib>
using namespace cb;
namespace cb { void set_error(const char*, ...) {} std::atomic<long long> g_launch_count{0}; thread_local long long t_launch_count = 0; }

template <int DEG, bool ZDEP>
static double run(int trials) {
  double worst = 0;
  for (int trial = 0; trial < trials; ++trial) {
    cb_poly p;
    memset(&p, 0, sizeof(p));
    p.nterms = 1 + rand() % 24;
    for (int t = 0; t < p.nterms; ++t) {
      const int i = rand() % (DEG + 1), j = rand() % (DEG + 1 - i), k = ZDEP ? rand() % (DEG + 1 - i - j) : 0;
      p.ex[t] = i; p.ey[t] = j; p.ez[t] = k;
      p.coef[t] = (rand() / (double)RAND_MAX - 0.5) * 4;
    }
    DensePoly d;
    if (!dense_from_poly(p, d)) return 1.0;
    bool zdep = false;
    const int deg = dense_degree(d, &zdep);
    if (deg > DEG || (zdep && !ZDEP)) return 2.0;
    const double x = rand() / (double)RAND_MAX, y = rand() / (double)RAND_MAX, z = rand() / (double)RAND_MAX;
    const double a = poly_eval(p, x, y, z);
    worst = fmax(worst, fabs(a - dense_eval<3, DEG, ZDEP>(d, x, y, z)));
    if (!ZDEP) worst = fmax(worst, fabs(a - dense_eval<2, DEG, false>(d, x, y, 0.0)));
    if (DEG < DP_DEG) worst = fmax(worst, fabs(a - dense_eval<3, DP_DEG, true>(d, x, y, z)));
  }
  return worst;
}

int main() {
  srand(1);
  double worst = 0;
  worst = fmax(worst, run<4, true>(2000));
  worst = fmax(worst, run<4, false>(2000));
  worst = fmax(worst, run<2, true>(2000));
  worst = fmax(worst, run<2, false>(2000));
  worst = fmax(worst, run<3, true>(500));
  worst = fmax(worst, run<1, false>(500));
  worst = fmax(worst, run<0, false>(10));
  cb_poly p;
  memset(&p, 0, sizeof(p));
  p.nterms = 1; p.ex[0] = 5; p.coef[0] = 1;
  DensePoly d;
  const bool rejected = !dense_from_poly(p, d);
  printf("worst %.3e, degree-5 rejected %d\n", worst, (int)rejected);
  return (worst < 1e-13 && rejected) ? 0 : 1;
}
