/* host_air3d.c -- a host WITHOUT Python driving the hot path through the C-ABI (include/hjb200.h).
 *
 * The air3D game of the reference (ValueFuncs/hji_solver.py:509-599 over DubinsVehicleRel, dubins_relative.py:44-111):
 * grid + initial cylinder built here, then per CFL step  dt = min(factorCFL * stepBound, tEnd - t)  (ode_cfl_3.py:142)
 * and one hj_step (three fused TVD-RK3 stage kernels + the minVOverTime epilogue); the field stays resident in HBM.
 *
 *   gcc -std=c99 -O2 -ffp-contract=off -Iinclude examples/host_air3d.c -o host_air3d levelsetpy_b200/_hjb200.so -lm
 *   ./host_air3d [N [steps]]        prints one line:  N steps t sum min max
 *
 * Exit code 3 with the library's own message when there is no CUDA device (there is no CPU fallback).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "hjb200.h"

#define CK(call)                                                                \
  do {                                                                          \
    int rc_ = (call);                                                           \
    if (rc_ != HJ_OK) {                                                         \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, hj_last_error());           \
      return 3;                                                                 \
    }                                                                           \
  } while (0)

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 101;
  const int steps = argc > 2 ? atoi(argv[2]) : 5;
  const double pi = 3.14159265358979323846;
  const double lo[3] = {-6.0, -10.0, 0.0};
  const double hi[3] = {20.0, 10.0, 2.0 * pi * (1.0 - 1.0 / n)};     /* periodic heading: the last node is 2 pi - dx */
  const int64_t N[3] = {n, n, n};
  const int bc[3] = {HJ_BC_EXTRAPOLATE, HJ_BC_EXTRAPOLATE, HJ_BC_PERIODIC};   /* grid.bdry */
  const int tz[3] = {0, 0, 0};                                                 /* grid.bdryData[d].towardZero */
  double dx[3];
  double* vs[3];
  for (int d = 0; d < 3; ++d) {
    dx[d] = (hi[d] - lo[d]) / (n - 1);                              /* process_grid.py:185 */
    vs[d] = (double*)malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i) vs[d][i] = i * dx[d] + lo[d];       /* np.linspace(min, max, N), process_grid.py:204 */
    vs[d][n - 1] = hi[d];
  }

  hj_ctx* ctx = NULL;
  CK(hj_create(&ctx, 0, 3, N, dx, bc, tz, HJ_WENO_AS_SHIPPED));
  for (int d = 0; d < 3; ++d) CK(hj_set_axis(ctx, d, vs[d], n));
  double* cs = (double*)malloc(sizeof(double) * n);
  double* sn = (double*)malloc(sizeof(double) * n);
  for (int i = 0; i < n; ++i) { cs[i] = cos(vs[2][i]); sn[i] = sin(vs[2][i]); }
  CK(hj_set_table(ctx, 0, cs, n));                                  /* cos(grid.vs[2]) */
  CK(hj_set_table(ctx, 1, sn, n));                                  /* sin(grid.vs[2]) */
  const double u_bound = 5.0, w_bound = 1.0;
  const double params[5] = {u_bound, u_bound, w_bound, w_bound, w_bound};   /* v_e v_p w(1) w_e w_p, dubins_relative.py:44-61 */
  CK(hj_set_system(ctx, HJ_SYS_DUBINS_REL, params, 5));

  /* shapeCylinder(grid, 2, 0, 5): sqrt(x0^2 + x1^2) - 5, constant along the heading (InitialConditions/cylinder.py) */
  const int64_t nodes = hj_num_nodes(ctx);
  double* y = (double*)malloc(sizeof(double) * (size_t)nodes);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      const double v = sqrt(vs[0][i] * vs[0][i] + vs[1][j] * vs[1][j]) - 5.0;
      for (int k = 0; k < n; ++k) y[((int64_t)i * n + j) * n + k] = v;
    }
  CK(hj_upload(ctx, NULL, HJ_FIELD_STATE, y, 1));

  const double factor_cfl = 0.8, t_end = 1.0;
  double t = 0.0, alpha_max[3], step_bound = 0.0;
  CK(hj_alpha_max(ctx, NULL, t, alpha_max, &step_bound));           /* state-only alpha: one bound for the whole run */
  for (int s = 0; s < steps && t_end - t >= 1e-4; ++s) {
    double dt = factor_cfl * step_bound;
    if (t_end - t < dt) dt = t_end - t;                             /* ode_cfl_3.py:142-143 */
    CK(hj_step(ctx, NULL, t, dt, NULL, HJ_COMP_MIN_OVER_TIME, 0, 0));
    const double t1 = t + dt, t2 = t1 + dt;                         /* ode_cfl_3.py:145,178,187,220,236: the time bookkeeping */
    const double t_half = 0.25 * (3.0 * t + t2);
    const double t_three_half = t_half + dt;
    t = (1.0 / 3.0) * (t + 2.0 * t_three_half);
  }
  CK(hj_download(ctx, NULL, HJ_FIELD_STATE, y, 1));                 /* synchronises */

  double sum = 0.0, mn = y[0], mx = y[0];
  for (int64_t i = 0; i < nodes; ++i) {
    sum += y[i];
    if (y[i] < mn) mn = y[i];
    if (y[i] > mx) mx = y[i];
  }
  printf("%d %d %.17g %.17g %.17g %.17g %lld\n", n, steps, t, sum, mn, mx, (long long)hj_launch_count());
  CK(hj_destroy(ctx));
  for (int d = 0; d < 3; ++d) free(vs[d]);
  free(cs); free(sn); free(y);
  return 0;
}
