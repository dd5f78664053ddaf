// hj_common.cuh -- shared device code: kernel parameter blocks, on-the-fly ghost cells, the fifth-order
// upwind schemes, RK stage algebra and block reductions.  sm_100a only.
//
// Reference behaviour restated here (paths relative to the LevelSetPy checkout):
//   ghost cells        BoundaryCondition/add_ghost_extrapolate.py:88-110, add_ghost_periodic.py:78-87
//   candidates/weights SpatialDerivative/ENO3aHelper.py:76-189, upwind_first_weno5a.py:107-196
//   LF term            ExplicitIntegration/Term/term_lax_friedrich.py:106-128
//   GLF dissipation    ExplicitIntegration/Dissipation/artificial_diss_glf.py:80-109
//   RK3 stage algebra  ExplicitIntegration/Integration/ode_cfl_3.py:151-241
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hjb200.h"

#define HJ_DEV __device__ __forceinline__

struct KGrid {
  int D;
  int N[HJ_MAX_DIM];            // nodes per dim (this slab, without halo planes)
  long long stride[HJ_MAX_DIM]; // element stride per dim of the field being read
  double dx[HJ_MAX_DIM];
  double dxinv[HJ_MAX_DIM];     // 1/dx as the host computed it (ENO3aHelper.py:57)
  int bc[HJ_MAX_DIM];           // hj_bc
  double slope_mult[HJ_MAX_DIM];// +1 / -1 (towardZero), add_ghost_extrapolate.py:61-64
  const double* vs[HJ_MAX_DIM]; // grid.vs[d] on the device
  // host-precomputed coefficients of the fixed-weight scheme (constant-bank operands in the TMA kernel):
  //   derivC = ca1 (v4-v2) + ca2 (v5-v1) + ca3 (v6-v0);  0.5 (R-L) = cb (v0+v6 - 6 (v1+v5) + 15 (v2+v4) - 20 v3)
  double ca1[HJ_MAX_DIM], ca2[HJ_MAX_DIM], ca3[HJ_MAX_DIM], cb[HJ_MAX_DIM];
};

struct KSys {
  double p[HJ_MAX_PARAMS];
  const double* tab[HJ_MAX_TABLES];
};

struct KStage {
  int stage;        // 0: ydot only (termLaxFriedrichs); 1,2,3: fused RK3 stages
  int comp;         // hj_comp (stage 3 only)
  int use_obs;      // stage 3 only
  int want_reduce;
  int restrict_sign;  // termRestrictUpdate (term_restrict_update.py:91-94): +1 ydot = max(ydot,0), -1 min(ydot,0), 0 off
  double fin_a, fin_b;  // final-stage combination y = fin_a*(y0 + fin_b*yLast): RK3 (1/3, 2) ode_cfl_3.py:241, RK2 (1/2, 1) ode_cfl_2.py
  double dt;
  const double* dt_arr; // batch contexts: one dt per batch element (dim 0), else nullptr
  const double* in;   // field the stencil reads
  const double* y0;   // y at the start of the step (stages 2,3), pointwise
  const double* tmp;  // dimension-split path, pass 2: in + dt * F_B(in) written by pass 1 (pointwise)
  const double* aux;  // target / V0 (stage 3, comp 3/4)
  const double* obs;  // obstacle (stage 3)
  double* out;
  long long out_stride[HJ_MAX_DIM];  // stride of out/y0/aux/obs (pitched or dense)
  // slab jobs, pass 2 of the dimension-split path with fused halo pushes (hj_halo_set_fused): the nodes of my three
  // lowest / highest dim-0 planes are ALSO stored into the lower / upper neighbour's stored halo planes of the output
  // buffer (peer memory over NVLink): push_lo / push_hi + the node's own offset is the peer address, or nullptr
  double* push_lo;
  double* push_hi;
  unsigned long long* red;           // HJ_REDUCE_LEN(D) ordered-uint64 slots, or nullptr
  const unsigned long long* epsmax;  // D ordered-uint64 raw max(D1^2) (intended WENO), or nullptr
};

// ---------------------------------------------------------------- ordered encoding for fp64 atomics
HJ_DEV unsigned long long enc_ordered(double x) {
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
HJ_DEV double dec_ordered(unsigned long long e) {
  unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
  return __longlong_as_double((long long)b);
}

HJ_DEV double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
HJ_DEV double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- ghost cells
// addGhostExtrapolate: slope = m*|edge-next|*sign(edge); ghost `dist` cells outside = edge + dist*slope.
// Un-fused multiplies/adds so the ghost values are bit-identical to numpy's.
HJ_DEV double ghost_extrapolate(double edge, double next, int dist, double m) {
  double sgn = (edge > 0.0) ? 1.0 : ((edge < 0.0) ? -1.0 : 0.0);
  double slope = __dmul_rn(__dmul_rn(m, fabs(__dsub_rn(edge, next))), sgn);
  return __dadd_rn(edge, __dmul_rn((double)dist, slope));
}

// v[k] = phi[i + k - 3], k = 0..6, along one dim, ghosts resolved on the fly.
// `p` points at node i; `s` is the element stride of the dim.
HJ_DEV void load_stencil(const double* __restrict__ p, int i, int n, long long s, int bc, double m, double v[7]) {
  if ((i >= HJ_GHOST && i < n - HJ_GHOST) || bc == HJ_BC_HALO) {
#pragma unroll
    for (int k = 0; k < 7; ++k) v[k] = __ldg(p + (long long)(k - 3) * s);
  } else if (bc == HJ_BC_PERIODIC) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      int j = i + k - 3;
      j = (j < 0) ? j + n : ((j >= n) ? j - n : j);
      v[k] = __ldg(p + (long long)(j - i) * s);
    }
  } else {
    double e0 = 0, e1 = 0, f0 = 0, f1 = 0;
    if (i < HJ_GHOST) {
      e0 = __ldg(p + (long long)(0 - i) * s);
      e1 = __ldg(p + (long long)(1 - i) * s);
    }
    if (i >= n - HJ_GHOST) {
      f0 = __ldg(p + (long long)(n - 1 - i) * s);
      f1 = __ldg(p + (long long)(n - 2 - i) * s);
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      int j = i + k - 3;
      if (j < 0) v[k] = ghost_extrapolate(e0, e1, -j, m);
      else if (j >= n) v[k] = ghost_extrapolate(f0, f1, j - (n - 1), m);
      else v[k] = __ldg(p + (long long)(k - 3) * s);
    }
  }
}

// ---------------------------------------------------------------- fifth-order upwind derivative pair
// v[0..6] = phi[i-3..i+3].  Returns derivL (phi^-) and derivR (phi^+).
//
// AS_SHIPPED: upwind_first_weno5a.py as shipped has all three smoothness indicators == 0 (aliasing at :97),
// so weightWENO (:189-194) reduces to the fixed weights (.1,.6,.3)/(.3,.6,.1) applied to the ENO3 candidates of
// ENO3aHelper.py:116-189.  Written on the six first differences; agrees with the reference to a few ulp.
HJ_DEV void upwind5_linear(const double v[7], double dxinv, double& L, double& R) {
  const double d0 = v[1] - v[0], d1 = v[2] - v[1], d2 = v[3] - v[2];
  const double d3 = v[4] - v[3], d4 = v[5] - v[4], d5 = v[6] - v[5];
  const double c2 = 2.0 / 60.0, c13 = 13.0 / 60.0, c47 = 47.0 / 60.0, c27 = 27.0 / 60.0, c3 = 3.0 / 60.0;
  L = dxinv * (c2 * d0 - c13 * d1 + c47 * d2 + c27 * d3 - c3 * d4);
  R = dxinv * (-c3 * d1 + c27 * d2 + c47 * d3 - c13 * d4 + c2 * d5);
}

// INTENDED: Osher & Fedkiw (3.32)-(3.41) / Mitchell's upwindFirstWENO5a, i.e. what upwind_first_weno5a.py:107-172
// computes without the aliasing bug: v_k = D1 (unstripped), smoothness S1..S3, alpha_k = w_k/(S_k+eps)^2 with
// eps = 1e-6*max(D1^2)+1e-99.  inv_eps = 1/eps.  The weights are evaluated as w_k * prod_{j!=k} q_j with
// q_j = (S_j/eps + 1)^2, which is alpha_k * (eps^2 * q0 q1 q2): same ratio, one division, and no under/overflow
// when eps == 1e-99 (a field that is flat along this dim).
HJ_DEV double weno_combine(double c0, double c1, double c2, double s0, double s1, double s2, double w0, double w1,
                           double w2, double inv_eps) {
  double q0 = fma(s0, inv_eps, 1.0), q1 = fma(s1, inv_eps, 1.0), q2 = fma(s2, inv_eps, 1.0);
  q0 *= q0; q1 *= q1; q2 *= q2;
  const double a0 = w0 * (q1 * q2), a1 = w1 * (q0 * q2), a2 = w2 * (q0 * q1);
  return (a0 * c0 + a1 * c1 + a2 * c2) / (a0 + a1 + a2);
}

HJ_DEV void upwind5_weno(const double v[7], double dxinv, double inv_eps, double& L, double& R) {
  const double v1 = dxinv * (v[1] - v[0]), v2 = dxinv * (v[2] - v[1]), v3 = dxinv * (v[3] - v[2]);
  const double v4 = dxinv * (v[4] - v[3]), v5 = dxinv * (v[5] - v[4]), v6 = dxinv * (v[6] - v[5]);
  const double k6 = 1.0 / 6.0;
  // third-order candidates on {i-3..i}, {i-2..i+1}, {i-1..i+2}, {i..i+3}  (dL[1]==dR[0], dL[2]==dR[1])
  const double c0 = k6 * (2.0 * v1 - 7.0 * v2 + 11.0 * v3);
  const double c1 = k6 * (-v2 + 5.0 * v3 + 2.0 * v4);
  const double c2 = k6 * (2.0 * v3 + 5.0 * v4 - v5);
  const double c3 = k6 * (11.0 * v4 - 7.0 * v5 + 2.0 * v6);
  const double k13 = 13.0 / 12.0;
  const double t123 = v1 - 2.0 * v2 + v3, t234 = v2 - 2.0 * v3 + v4, t345 = v3 - 2.0 * v4 + v5,
               t456 = v4 - 2.0 * v5 + v6;
  double u;
  // left: S1(v1,v2,v3), S2(v2,v3,v4), S3(v3,v4,v5)
  u = v1 - 4.0 * v2 + 3.0 * v3; const double sl0 = k13 * t123 * t123 + 0.25 * u * u;
  u = v2 - v4;                  const double sl1 = k13 * t234 * t234 + 0.25 * u * u;
  u = 3.0 * v3 - 4.0 * v4 + v5; const double sl2 = k13 * t345 * t345 + 0.25 * u * u;
  // right: the same three forms one position up: S1(v2,v3,v4), S2(v3,v4,v5), S3(v4,v5,v6)
  u = v2 - 4.0 * v3 + 3.0 * v4; const double sr0 = k13 * t234 * t234 + 0.25 * u * u;
  u = v3 - v5;                  const double sr1 = k13 * t345 * t345 + 0.25 * u * u;
  u = 3.0 * v4 - 4.0 * v5 + v6; const double sr2 = k13 * t456 * t456 + 0.25 * u * u;
  L = weno_combine(c0, c1, c2, sl0, sl1, sl2, 0.1, 0.6, 0.3, inv_eps);
  R = weno_combine(c1, c2, c3, sr0, sr1, sr2, 0.3, 0.6, 0.1, inv_eps);
}

// ENO divided differences around node i from v[0..6] = phi[i-3..i+3], with the reference's operation order
// (ENO3aHelper.py:76-88: scalar products first) and un-fused arithmetic, so that the tables -- hence the
// minimum-modulus choices, which are discontinuous in them -- are bit-identical to numpy's:
//   a[k] = dxInv (v[k+1]-v[k])                 a[2] = D1[i], a[3] = D1[i+1]
//   b[k] = (0.5 dxInv)(a[k+1]-a[k])            b[1], b[2], b[3] = D2[i], D2[i+1], D2[i+2]  (centred at i-1, i, i+1)
//   c[k] = ((1/3) dxInv)(b[k+1]-b[k])          c[0..3] = D3[i..i+3]                         (centred at i-3/2 ..)
struct EnoTables { double a[6], b[5], c[4]; };
HJ_DEV void eno_tables(const double v[7], double dxinv, EnoTables& T, bool third) {
  const double h2 = __dmul_rn(0.5, dxinv), h3 = __dmul_rn(1.0 / 3.0, dxinv);
#pragma unroll
  for (int k = 0; k < 6; ++k) T.a[k] = __dmul_rn(dxinv, __dsub_rn(v[k + 1], v[k]));
#pragma unroll
  for (int k = 0; k < 5; ++k) T.b[k] = __dmul_rn(h2, __dsub_rn(T.a[k + 1], T.a[k]));
  if (third) {
#pragma unroll
    for (int k = 0; k < 4; ++k) T.c[k] = __dmul_rn(h3, __dsub_rn(T.b[k + 1], T.b[k]));
  }
}

// upwindFirstENO3a (upwind_first_eno3a.py:87-142): candidates of ENO3aHelper.py:116-189, the one built on the
// minimum-modulus D2 and then D3 neighbours is taken (:104-140).
// the three left and three right third-order candidates of upwindFirstENO3aHelper (ENO3aHelper.py:116-189) -- what
// upwindFirstWENO5a / upwindFirstENO3a return for generateAll = True (upwind_first_weno5a.py:73-75)
HJ_DEV void eno3a_candidates(const EnoTables& T, double dx, double dL[3], double dR[3]) {
  const double dx2 = __dmul_rn(dx, dx), cLL = __dmul_rn(2.0, dx2), cLR = -dx2;
  const double l01 = __dadd_rn(T.a[2], __dmul_rn(dx, T.b[1])), l2 = __dadd_rn(T.a[2], __dmul_rn(dx, T.b[2]));
  const double r01 = __dadd_rn(T.a[3], __dmul_rn(-dx, T.b[2])), r2 = __dadd_rn(T.a[3], __dmul_rn(-dx, T.b[3]));
  dL[0] = __dadd_rn(l01, __dmul_rn(cLL, T.c[0]));
  dL[1] = __dadd_rn(l01, __dmul_rn(cLL, T.c[1]));
  dL[2] = __dadd_rn(l2, __dmul_rn(cLR, T.c[2]));
  dR[0] = __dadd_rn(r01, __dmul_rn(cLR, T.c[1]));
  dR[1] = __dadd_rn(r01, __dmul_rn(cLR, T.c[2]));
  dR[2] = __dadd_rn(r2, __dmul_rn(cLL, T.c[3]));
}

HJ_DEV void upwind_eno3a(const double v[7], double dx, double dxinv, double& L, double& R) {
  EnoTables T;
  eno_tables(v, dxinv, T, true);
  double cL[3], cR[3];
  eno3a_candidates(T, dx, cL, cR);
  const double dL0 = cL[0], dL1 = cL[1], dL2 = cL[2], dR0 = cR[0], dR1 = cR[1], dR2 = cR[2];
  const bool t0 = fabs(T.c[0]) < fabs(T.c[1]), t1 = fabs(T.c[1]) < fabs(T.c[2]), t2 = fabs(T.c[2]) < fabs(T.c[3]);
  {  // left: index i of the masks
    const bool sL = fabs(T.b[1]) < fabs(T.b[2]);
    const bool LL = t0 && sL, RR = !t1 && !sL;
    L = LL ? dL0 : (RR ? dL2 : dL1);
  }
  {  // right: index i+1 of the masks
    const bool sL = fabs(T.b[2]) < fabs(T.b[3]);
    const bool LL = t1 && sL, RR = !t2 && !sL;
    R = LL ? dR0 : (RR ? dR2 : dR1);
  }
}

// upwindFirstENO2 (upwind_first_eno2.py:50-150): two ghost cells (v[1..5]), minimum-modulus second-order term.
HJ_DEV void upwind_eno2(const double v[7], double dx, double dxinv, double& L, double& R) {
  EnoTables T;
  eno_tables(v, dxinv, T, false);
  const double dL0 = __dadd_rn(T.a[2], __dmul_rn(dx, T.b[1])), dL1 = __dadd_rn(T.a[2], __dmul_rn(dx, T.b[2]));
  const double dR0 = __dsub_rn(T.a[3], __dmul_rn(dx, T.b[2])), dR1 = __dsub_rn(T.a[3], __dmul_rn(dx, T.b[3]));
  L = fabs(T.b[1]) < fabs(T.b[2]) ? dL0 : dL1;
  R = fabs(T.b[2]) < fabs(T.b[3]) ? dR0 : dR1;
}

template <int WENO>
HJ_DEV void upwind5(const double v[7], double dxinv, double inv_eps, double& L, double& R, double dx = 0.0) {
  if (WENO == HJ_WENO_AS_SHIPPED) upwind5_linear(v, dxinv, L, R);
  else if (WENO == HJ_WENO_INTENDED) upwind5_weno(v, dxinv, inv_eps, L, R);
  else if (WENO == HJ_SCHEME_ENO3A) upwind_eno3a(v, dx, dxinv, L, R);
  else upwind_eno2(v, dx, dxinv, L, R);
}

// max over the D1 entries this node is responsible for (unstripped table: node pairs (-3,-2)..(N+1,N+2)):
// pair (i,i+1) always; node 0 adds the three pairs below it, node N-1 the two pairs above (i+1,i+2),(i+2,i+3).
HJ_DEV double d1sq_local(const double v[7], double dxinv, int i, int n) {
  double d = dxinv * (v[4] - v[3]);
  double m = d * d;
  if (i == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { d = dxinv * (v[k + 1] - v[k]); m = fmax(m, d * d); }
  }
  if (i == n - 1) {
#pragma unroll
    for (int k = 4; k < 6; ++k) { d = dxinv * (v[k + 1] - v[k]); m = fmax(m, d * d); }
  }
  return m;
}

HJ_DEV double inv_eps_from_max(unsigned long long enc) {
  const double mx = dec_ordered(enc);
  return 1.0 / (1e-6 * mx + 1e-99);   // upwind_first_weno5a.py:156
}

// ---------------------------------------------------------------- RK3 stage algebra + driver epilogue
// ode_cfl_3.py:151 (y1), :184,:193 (y2, yHalf), :226,:241 (yThreeHalf, y); hji_solver.py:571-599, :641-644.
// Driver epilogue and termRestrictUpdate min / max: hji_solver.py:571-599 and term_restrict_update.py:92-94 use
// np.minimum / np.maximum, which PROPAGATE a NaN; fmin / fmax would return the other operand and silently repair a
// blown-up node, so the NaN check of hji_solver.py:544 could never fire.  The second operands (y0 / target /
// obstacle / 0) are finite inputs: propagate the NaN of the integrated value.
HJ_DEV double nan_min(double y, double ref) { return (y == y) ? fmin(y, ref) : y; }
HJ_DEV double nan_max(double y, double ref) { return (y == y) ? fmax(y, ref) : y; }
HJ_DEV double restrict_update(double ydot, int sign) {
  return sign > 0 ? nan_max(ydot, 0.0) : (sign < 0 ? nan_min(ydot, 0.0) : ydot);
}

HJ_DEV double comp_epilogue(double y, int comp, double y0, double aux) {
  switch (comp) {
    case HJ_COMP_MIN_OVER_TIME: return nan_min(y, y0);
    case HJ_COMP_MAX_OVER_TIME: return nan_max(y, y0);
    case HJ_COMP_MIN_WITH_AUX: return nan_min(y, aux);
    case HJ_COMP_MAX_WITH_AUX: return nan_max(y, aux);
    default: return y;
  }
}

HJ_DEV double stage_update(const KStage& st, double yin, double ydot, long long oidx) {
  ydot = restrict_update(ydot, st.restrict_sign);
  if (st.stage == 0) return ydot;
  if (st.stage == 1) return yin + st.dt * ydot;
  const double y0 = st.y0[oidx];
  if (st.stage == 2) {
    const double y2 = yin + st.dt * ydot;
    return 0.25 * (3.0 * y0 + y2);
  }
  const double y32 = yin + st.dt * ydot;
  double y = st.fin_a * (y0 + st.fin_b * y32);
  if (st.comp == HJ_COMP_MIN_WITH_AUX || st.comp == HJ_COMP_MAX_WITH_AUX) y = comp_epilogue(y, st.comp, y0, st.aux[oidx]);
  else y = comp_epilogue(y, st.comp, y0, 0.0);
  if (st.use_obs) y = nan_max(y, -st.obs[oidx]);
  return y;
}

// ---------------------------------------------------------------- per-thread reduction accumulator
template <int D>
struct RedAcc {
  double amax[D], dmin[D], dmax[D];
  int nan;
  HJ_DEV void init() {
#pragma unroll
    for (int d = 0; d < D; ++d) { amax[d] = -INFINITY; dmin[d] = INFINITY; dmax[d] = -INFINITY; }
    nan = 0;
  }
  // block-wide: warp shuffles, then at most one atomic per warp per slot -- and only when the warp's value would
  // change the record (a plain load of the current record filters almost all of them: half a million CTAs hammering
  // 3 D addresses serialise in L2 otherwise; a stale read only costs a redundant atomic)
  HJ_DEV void flush(unsigned long long* red) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double a = warp_max(amax[d]), lo = warp_min(dmin[d]), hi = warp_max(dmax[d]);
      if ((threadIdx.x + threadIdx.y * blockDim.x) % 32 == 0) {
        const volatile unsigned long long* cur = red;
        const unsigned long long ea = enc_ordered(a), el = enc_ordered(lo), eh = enc_ordered(hi);
        if (ea > cur[d]) atomicMax(red + d, ea);
        if (el < cur[D + d]) atomicMin(red + D + d, el);
        if (eh > cur[2 * D + d]) atomicMax(red + 2 * D + d, eh);
      }
    }
    const int any = __any_sync(0xffffffffu, nan);
    if (any && (threadIdx.x + threadIdx.y * blockDim.x) % 32 == 0) atomicOr(red + 3 * D, 1ull);
  }
};
