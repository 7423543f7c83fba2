// Host build of oak_b200/csrc/tridiag_math.cuh for CPU checks (tools/proto_tridiag.py --host):
//   g++ -O2 -shared -fPIC -ffp-contract=off tools/tridiag_host.cpp -o /tmp/libtridiag_host.so
#include "../oak_b200/csrc/tridiag_math.cuh"
extern "C" int host_tql(int n, double *d, double *e, int s, double tn) { return tql_eigenvalues(n, d, e, s, tn); }
extern "C" double host_twisted(int n, const double *d, const double *e, int sd, double lam, double pivmin,
                               double *w, int sw, double *gam) {
  return twisted_vector(n, d, e, sd, lam, pivmin, w, sw, gam);
}
extern "C" int host_pwk(int n, double *d, double *e, int s, double tn) { return pwk_eigenvalues(n, d, e, s, tn); }
extern "C" double host_twisted2(int n, const double *d, const double *e, int sd, double lam, double pivmin,
                                double *w, int sw, double *gam) {
  return twisted_vector2(n, d, e, sd, lam, pivmin, w, sw, gam);
}
extern "C" double host_twisted3(int n, const double *d, const double *e, int sd, double lam, double scale,
                                double *w, int sw, double *gam) {
  // the per-matrix arrays k_tvec keeps in shared memory
  double ds[256], e2[256], en[256];
  for (int i = 0; i < n; i++) {
    ds[i] = scale * d[i * sd];
    const double es = i < n - 1 ? scale * e[i * sd] : 0.;
    e2[i] = es * es; en[i] = -es;
  }
  return twisted_vector3(n, ds, e2, en, scale * lam, 1. / scale, w, sw, gam);
}
extern "C" int host_pwk_pf(int n, double *d, double *e, int s, double tn) { return pwk_eigenvalues_t<6>(n, d, e, s, tn); }
