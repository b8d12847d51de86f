/* C99 harness for include/ace_b200.h: proves that the header is plain C (no C++ constructs, no torch types) and that
 * the library's entry points bind through real C linkage.  Only host-side entry points are called, so it runs on a
 * machine without a GPU; on a GPU box it additionally runs one tiny kernel build + inverse through the C ABI. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ace_b200.h"

int main(void) {
  const char* v = ace_version();
  if (!v || strncmp(v, "ace_b200", 8) != 0) {
    printf("FAIL version\n");
    return 1;
  }
  int blocks[2] = {-1, -1}, width = -1;
  if (ace_shard_plan(16384, 8, 3, blocks, &width) != 0 || blocks[0] != 3 || blocks[1] != 12 || width != 1024) {
    printf("FAIL shard_plan %d %d %d\n", blocks[0], blocks[1], width);
    return 1;
  }
  struct ace_fit_config cfg;
  ace_fit_default_config(&cfg);
  if (cfg.kernel != ACE_KERNEL_SE || cfg.optimizer != ACE_OPT_NADAM || cfg.learning_rate != 0.01 || cfg.norm_clip != 1) {
    printf("FAIL default_config\n");
    return 1;
  }
  /* host-side optimiser step through the ABI (src/optimizer_cpp.cpp:23-42) */
  double m[3] = {0, 0, 0}, vv[3] = {0, 0, 0}, g[3] = {0.5, -0.25, 0.0}, par[3] = {1, 2, 3};
  if (ace_Nadam_cpp(1.0, 0.01, 0.9, 0.999, 1e-8, m, vv, g, par, 3) != 1 || !(par[0] > 1.0) || !(par[1] < 2.0) ||
      par[2] != 3.0) {
    printf("FAIL Nadam %g %g %g\n", par[0], par[1], par[2]);
    return 1;
  }
  const int ndev = ace_device_count();
  printf("OK %s devices=%d\n", v, ndev);
  if (ndev > 0) {
    /* K = kernmat_SE_symmetric(X, Z, theta); inv = invkernel(K, sigma); check inv * (K + e^sigma I) = I on a probe */
    enum { N = 40, P = 2, BZ = 1, B = 2, NPAR = 2 + B + B * P };
    double X[N * P], Z[N * BZ], th[NPAR], K[N * N], cube[N * N * B], eig[N], inv[N * N];
    int i, j;
    for (i = 0; i < N * P; ++i) X[i] = sin(0.37 * i);
    for (i = 0; i < N; ++i) Z[i] = cos(0.11 * i);
    th[0] = log(0.3);
    th[1] = 0.0;
    for (i = 2; i < 2 + B; ++i) th[i] = 0.0;
    for (i = 2 + B; i < NPAR; ++i) th[i] = log(2.0);
    if (ace_kernmat_SE_symmetric_cpp(X, Z, N, P, BZ, th, K, cube) != 0) {
      printf("FAIL kernmat: %s\n", ace_last_error());
      return 1;
    }
    if (ace_invkernel_cpp(K, N, th[0], eig, inv) != 0) {
      printf("FAIL invkernel: %s\n", ace_last_error());
      return 1;
    }
    double worst = 0.0;
    for (i = 0; i < N; ++i)
      for (j = 0; j < N; ++j) {
        double s = 0.0;
        int k;
        for (k = 0; k < N; ++k) s += inv[i + N * k] * (K[k + N * j] + (k == j ? exp(th[0]) : 0.0));
        s -= (i == j) ? 1.0 : 0.0;
        if (fabs(s) > worst) worst = fabs(s);
      }
    printf("GPU residual %.3e\n", worst);
    if (!(worst < 1e-9)) return 1;
  }
  return 0;
}
