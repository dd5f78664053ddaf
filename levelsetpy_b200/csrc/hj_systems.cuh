// hj_systems.cuh -- compiled device functors for schemeData.hamFunc / schemeData.partialFunc.
//
// Each functor restates one DynamicalSystems class of the reference:
//   DubinsRelF   dubins_relative.py:63-111      H = p1(v_e - v_p cos x3) - p2 v_p sin x3 - w|p1 x2 - p2 x1 - p3| + w|p3|
//   DoubleIntF   double_integrator.py:49-89     H = -(p1 x2 - |p2| u)
//   FlockF       flock.py:190-258, bird.py:235-372   H = min_j {H_abs(j), H_attacked}; scalar coefficients per bird
//   PairF<A,B>   product construction of SURVEY.md 8(d): H = H_A + H_B on disjoint dim blocks
//
// alpha_d (the partialFunc of artificial_diss_glf.py:98) is state-only for every one of them, so the GLF
// stepBound does not depend on the field.  Products/sums that feed max_x alpha_d use un-fused (_rn) arithmetic and
// host-evaluated trig tables so that the device maximum -- hence dt -- is bit-identical to numpy's.
#pragma once
#include "hj_common.cuh"

// Parameter block layouts (doubles) ------------------------------------------------------------------------
//  DubinsRel : [0]=v_e [1]=v_p [2]=w(1) [3]=w_e [4]=w_p ; tables TB+0 = cos(vs[2]), TB+1 = sin(vs[2])
//  DoubleInt : [0]=u_bound
//  Flock     : [0]=K (# un-attacked birds) [1]=has_attacked [2]=W [3]=a1 [4]=a2 [5]=x_att [6]=y_att
//              [7..9]=alpha_0..2 (host scalars)  [10+3j..] = (c1,c2,c3)_j : H_abs_j = p1 c1 + p2 c2 + p3 c3
#define HJ_DUBINS_NP 5
#define HJ_DINT_NP 1
#define HJ_FLOCK_HDR 10

template <int BASE, int PB, int TB>
struct DubinsRelF {
  static constexpr int ND = 3, BASE_DIM = BASE, NSCRATCH = 0;
  // x1, x2 and the two x3-only coefficients of dubins_relative.py:81-82, evaluated un-fused like numpy does
  // awx1/awx2 = |w x1|, |w x2| (the state-only parts of alpha_1 / alpha_0) are kept per node so that the marched
  // plane body adds them instead of re-multiplying
  struct Pt { double x1, x2, p1c, p2c, awx1, awx2; };
  HJ_DEV static void set3(Pt& q, int i3, const KSys& k) {
    q.p1c = __dsub_rn(k.p[PB + 0], __dmul_rn(k.p[PB + 1], __ldg(k.tab[TB + 0] + i3)));   // v_e - v_p cos x3
    q.p2c = __dmul_rn(k.p[PB + 1], __ldg(k.tab[TB + 1] + i3));                          // v_p sin x3
  }
  HJ_DEV static Pt load(const int* idx, const KGrid& g, const KSys& k, const double* = nullptr) {
    Pt q;
    q.x1 = __ldg(g.vs[BASE + 0] + idx[BASE + 0]);
    q.x2 = __ldg(g.vs[BASE + 1] + idx[BASE + 1]);
    q.awx1 = fabs(__dmul_rn(k.p[PB + 2], q.x1));
    q.awx2 = fabs(__dmul_rn(k.p[PB + 2], q.x2));
    set3(q, idx[BASE + 2], k);
    return q;
  }
  // refresh only what depends on global dim GD (the marching dim of the plane-ring kernel), in two phases so the
  // table reads for plane z+1 can be in flight while plane z is computed
  template <int GD>
  HJ_DEV static double2 fetch(int i, const KGrid& g, const KSys& k) {
    if (GD == BASE + 0) return make_double2(__ldg(g.vs[BASE + 0] + i), 0.0);
    if (GD == BASE + 1) return make_double2(__ldg(g.vs[BASE + 1] + i), 0.0);
    if (GD == BASE + 2) return make_double2(__ldg(k.tab[TB + 0] + i), __ldg(k.tab[TB + 1] + i));
    return make_double2(0.0, 0.0);
  }
  template <int GD>
  HJ_DEV static void apply(Pt& q, const double2 r, const KSys& k) {
    if (GD == BASE + 0) { q.x1 = r.x; q.awx1 = fabs(__dmul_rn(k.p[PB + 2], r.x)); }
    if (GD == BASE + 1) { q.x2 = r.x; q.awx2 = fabs(__dmul_rn(k.p[PB + 2], r.x)); }
    if (GD == BASE + 2) {
      q.p1c = __dsub_rn(k.p[PB + 0], __dmul_rn(k.p[PB + 1], r.x));
      q.p2c = __dmul_rn(k.p[PB + 1], r.y);
    }
  }
  HJ_DEV static double ham(const Pt& q, const double* p, const KSys& k) {
    const double w = k.p[PB + 2];
    const double p1 = p[BASE + 0], p2 = p[BASE + 1], p3 = p[BASE + 2];
    return p1 * q.p1c - p2 * q.p2c - w * fabs(p1 * q.x2 - p2 * q.x1 - p3) + w * fabs(p3);
  }
  HJ_DEV static double alpha(int dl, const Pt& q, const KSys& k) {
    if (dl == 0) return __dadd_rn(fabs(q.p1c), q.awx2);
    if (dl == 1) return __dadd_rn(fabs(q.p2c), q.awx1);
    return __dadd_rn(k.p[PB + 3], k.p[PB + 4]);
  }
  // global dims alpha_dl depends on (bit d = dim d): max_x alpha_dl only needs that sub-grid (hj_alpha_max)
  static constexpr unsigned alpha_dims(int dl) {
    return dl == 0 ? (1u << (BASE + 1)) | (1u << (BASE + 2)) : (dl == 1 ? (1u << (BASE + 0)) | (1u << (BASE + 2)) : 0u);
  }
};

template <int BASE, int PB>
struct DoubleIntF {
  static constexpr int ND = 2, BASE_DIM = BASE, NSCRATCH = 0;
  struct Pt { double x2; };
  HJ_DEV static Pt load(const int* idx, const KGrid& g, const KSys& k, const double* = nullptr) {
    Pt q;
    q.x2 = __ldg(g.vs[BASE + 1] + idx[BASE + 1]);
    return q;
  }
  template <int GD>
  HJ_DEV static double2 fetch(int i, const KGrid& g, const KSys& k) {
    if (GD == BASE + 1) return make_double2(__ldg(g.vs[BASE + 1] + i), 0.0);
    return make_double2(0.0, 0.0);
  }
  template <int GD>
  HJ_DEV static void apply(Pt& q, const double2 r, const KSys& k) {
    if (GD == BASE + 1) q.x2 = r.x;
  }
  HJ_DEV static double ham(const Pt& q, const double* p, const KSys& k) {
    return -(p[BASE + 0] * q.x2 - fabs(p[BASE + 1]) * k.p[PB + 0]);
  }
  HJ_DEV static double alpha(int dl, const Pt& q, const KSys& k) {
    return dl == 0 ? fabs(q.x2) : fabs(k.p[PB + 0]);
  }
  static constexpr unsigned alpha_dims(int dl) { return dl == 0 ? (1u << (BASE + 1)) : 0u; }
};

struct FlockF {
  static constexpr int ND = 3, BASE_DIM = 0, NSCRATCH = 0;
  struct Pt { int dummy; };
  HJ_DEV static Pt load(const int*, const KGrid&, const KSys&, const double* = nullptr) { return Pt{0}; }
  template <int GD>
  HJ_DEV static double2 fetch(int, const KGrid&, const KSys&) { return make_double2(0.0, 0.0); }
  template <int GD>
  HJ_DEV static void apply(Pt&, const double2, const KSys&) {}
  HJ_DEV static double ham(const Pt&, const double* p, const KSys& k) {
    const int K = (int)k.p[0];
    const double p1 = p[0], p2 = p[1], p3 = p[2];
    double h = INFINITY;
    for (int j = 0; j < K; ++j) {
      const double* c = k.p + HJ_FLOCK_HDR + 3 * j;
      h = fmin(h, p1 * c[0] + p2 * c[1] + p3 * c[2]);                      // bird.py:266-273
    }
    if (k.p[1] != 0.0) {
      const double W = k.p[2];
      const double ha = (p1 * k.p[3] - p2 * k.p[4]) + W * fabs(p2 * k.p[5] - p1 * k.p[6] + p3) + W * fabs(p3);
      h = fmin(h, ha);                                                     // bird.py:305-316, flock.py:232-233
    }
    return h;
  }
  HJ_DEV static double alpha(int dl, const Pt&, const KSys& k) { return k.p[7 + dl]; }
  static constexpr unsigned alpha_dims(int) { return 0u; }
};

// Batch of independent Flock grids: the field is [nbatch, N0, N1, N2], dim 0 is the batch
// index and carries no stencil; every batch element has its own parameter block (layout as FlockF) in the device
// table k.tab[HJ_BATCH_TABLE] (k.p[0] = doubles per block), copied once per CTA into shared-memory scratch.
#define HJ_BATCH_TABLE (HJ_MAX_TABLES - 1)
struct FlockBatchF {
  static constexpr int ND = 3, BASE_DIM = 1, NSCRATCH = HJ_MAX_PARAMS;
  struct Pt { const double* P; };
  HJ_DEV static void fill_scratch(double* scratch, long long batch, const KSys& k, int tid, int nthreads) {
    const int n = (int)k.p[0];
    const double* src = k.tab[HJ_BATCH_TABLE] + batch * n;
    for (int i = tid; i < n; i += nthreads) scratch[i] = __ldg(src + i);
  }
  HJ_DEV static Pt load(const int*, const KGrid&, const KSys&, const double* scratch) { return Pt{scratch}; }
  template <int GD>
  HJ_DEV static double2 fetch(int, const KGrid&, const KSys&) { return make_double2(0.0, 0.0); }
  template <int GD>
  HJ_DEV static void apply(Pt&, const double2, const KSys&) {}
  HJ_DEV static double ham(const Pt& q, const double* p, const KSys&) {
    const double* P = q.P;
    const int K = (int)P[0];
    const double p1 = p[1], p2 = p[2], p3 = p[3];
    double h = INFINITY;
    for (int j = 0; j < K; ++j) {
      const double* c = P + HJ_FLOCK_HDR + 3 * j;
      h = fmin(h, p1 * c[0] + p2 * c[1] + p3 * c[2]);                      // bird.py:266-273
    }
    if (P[1] != 0.0) {
      const double W = P[2];
      const double ha = (p1 * P[3] - p2 * P[4]) + W * fabs(p2 * P[5] - p1 * P[6] + p3) + W * fabs(p3);
      h = fmin(h, ha);                                                     // bird.py:305-316, flock.py:232-233
    }
    return h;
  }
  HJ_DEV static double alpha(int dl, const Pt& q, const KSys&) { return q.P[7 + dl]; }
  static constexpr unsigned alpha_dims(int) { return 0u; }
};

// genericHam / genericPartial (Hamiltonians/generic_ham.py:5-57, generic_partial.py:6-58) with a DEVICE dynSys: the
// reference evaluates them through three Python methods of schemeData.dynSys -- get_opt_u(t, deriv, uMode, x),
// get_opt_v(t, deriv, dMode, x), dynamics(t, x, u, d) -- which cannot run inside a fused kernel.  A device dynSys is a
// struct `Dyn` with
//   NX, NU, NDST, NPAR                  state / control / disturbance dimensions, its own scalar parameters
//   Pt, load, fetch<GD>, apply<GD>      the state-dependent drift terms of a node (as the other functors)
//   opt_u(par, p, sgn, u)               get_opt_u for the costate p: sgn = +1 for uMode 'max', -1 for 'min'
//   opt_d(par, p, sgn, d)               get_opt_v likewise
//   f(i, q, par, u, d)                  dynamics(t, x, u, d)[i], un-fused arithmetic in the reference's order
//   alpha_dims(i)                       the dims |f_i| depends on through x
// and GenericF<Dyn> turns it into the functor interface of the stage kernels:
//   ham   = sum_i p_i f_i(x, u*(p), d*(p)), negated for tMode 'backward'                  (generic_ham.py:26-50)
//   alpha = max(|f_i(uU,dU)|, |f_i(uU,dL)|, |f_i(uL,dL)|, |f_i(uL,dU)|)                    (generic_partial.py:44-56)
// where uU / uL / dU / dL are the dynSys's optimal inputs at derivMax / derivMin -- GRID-WIDE scalars that the host obtains
// by calling the dynSys's own get_opt_u / get_opt_v on the reduced derivative range (hj_deriv_range) exactly as
// generic_partial.py:28-40 does, and hands over in the parameter block.  alpha therefore changes with every RHS evaluation.
// Block layout (doubles): [0] uSign [1] dSign [2] hamSign (-1: tMode 'backward') | uU[NU] uL[NU] dU[NDST] dL[NDST] | Dyn's NPAR.
// Registered dynSys: DubinsCarDyn (the Dubins car of the helperOC toolbox this API was written for):
//   dx0 = speed cos x2 + d0,  dx1 = speed sin x2 + d1,  dx2 = u + d2,  |u| <= wMax, |d_i| <= dMax_i;  par = speed wMax dMax[3]
template <int TB>
struct DubinsCarDyn {
  static constexpr int NX = 3, NU = 1, NDST = 3, NPAR = 5;
  struct Pt { double vc, vs; };                               // speed cos x2, speed sin x2 (products rounded like numpy's)
  HJ_DEV static Pt load(const int* idx, const KGrid&, const KSys& k, const double* par) {
    Pt q;
    q.vc = __dmul_rn(par[0], __ldg(k.tab[TB + 0] + idx[2]));
    q.vs = __dmul_rn(par[0], __ldg(k.tab[TB + 1] + idx[2]));
    return q;
  }
  template <int GD>
  HJ_DEV static double2 fetch(int i, const KGrid&, const KSys& k) {
    if (GD == 2) return make_double2(__ldg(k.tab[TB + 0] + i), __ldg(k.tab[TB + 1] + i));
    return make_double2(0.0, 0.0);
  }
  template <int GD>
  HJ_DEV static void apply(Pt& q, const double2 r, const double* par) {
    if (GD == 2) { q.vc = __dmul_rn(par[0], r.x); q.vs = __dmul_rn(par[0], r.y); }
  }
  // (deriv >= 0) * wMax + (deriv < 0) * (-wMax) for 'max', the opposite for 'min'
  HJ_DEV static void opt_u(const double* par, const double* p, double sgn, double* u) {
    u[0] = (p[2] >= 0.0) ? sgn * par[1] : -(sgn * par[1]);
  }
  HJ_DEV static void opt_d(const double* par, const double* p, double sgn, double* d) {
#pragma unroll
    for (int i = 0; i < 3; ++i) d[i] = (p[i] >= 0.0) ? sgn * par[2 + i] : -(sgn * par[2 + i]);
  }
  HJ_DEV static double f(int i, const Pt& q, const double*, const double* u, const double* d) {
    if (i == 0) return __dadd_rn(q.vc, d[0]);
    if (i == 1) return __dadd_rn(q.vs, d[1]);
    return __dadd_rn(u[0], d[2]);
  }
  static constexpr unsigned alpha_dims(int i) { return i < 2 ? (1u << 2) : 0u; }
};

template <class Dyn>
struct GenericF {
  static constexpr int ND = Dyn::NX, BASE_DIM = 0, NSCRATCH = 0;
  static constexpr int O_UU = 3, O_UL = O_UU + Dyn::NU, O_DU = O_UL + Dyn::NU, O_DL = O_DU + Dyn::NDST,
                       O_PAR = O_DL + Dyn::NDST, NP = O_PAR + Dyn::NPAR;
  using Pt = typename Dyn::Pt;
  HJ_DEV static Pt load(const int* idx, const KGrid& g, const KSys& k, const double* = nullptr) {
    return Dyn::load(idx, g, k, k.p + O_PAR);
  }
  template <int GD>
  HJ_DEV static double2 fetch(int i, const KGrid& g, const KSys& k) { return Dyn::template fetch<GD>(i, g, k); }
  template <int GD>
  HJ_DEV static void apply(Pt& q, const double2 r, const KSys& k) { Dyn::template apply<GD>(q, r, k.p + O_PAR); }
  HJ_DEV static double ham(const Pt& q, const double* p, const KSys& k) {
    double u[Dyn::NU], d[Dyn::NDST];
    Dyn::opt_u(k.p + O_PAR, p, k.p[0], u);                     // generic_ham.py:27
    Dyn::opt_d(k.p + O_PAR, p, k.p[1], d);                     // :32
    double h = 0.0;
#pragma unroll
    for (int i = 0; i < Dyn::NX; ++i) h += p[i] * Dyn::f(i, q, k.p + O_PAR, u, d);   // :45-47
    return k.p[2] < 0.0 ? -h : h;                              // :54-55
  }
  HJ_DEV static double alpha(int dl, const Pt& q, const KSys& k) {
    const double* par = k.p + O_PAR;
    const double *uU = k.p + O_UU, *uL = k.p + O_UL, *dU = k.p + O_DU, *dL = k.p + O_DL;
    double a = fmax(fabs(Dyn::f(dl, q, par, uU, dU)), fabs(Dyn::f(dl, q, par, uU, dL)));     // generic_partial.py:52
    a = fmax(a, fabs(Dyn::f(dl, q, par, uL, dL)));                                           // :53
    return fmax(a, fabs(Dyn::f(dl, q, par, uL, dU)));                                        // :54
  }
  static constexpr unsigned alpha_dims(int dl) { return Dyn::alpha_dims(dl); }
};

template <class A, class B>
struct PairF {
  static constexpr int ND = A::ND + B::ND, BASE_DIM = 0, NSCRATCH = 0;
  using First = A;     // dim block [0, A::ND)
  using Second = B;    // dim block [A::ND, ND): the trailing (contiguous) dims
  struct Pt { typename A::Pt a; typename B::Pt b; };
  HJ_DEV static Pt load(const int* idx, const KGrid& g, const KSys& k, const double* = nullptr) {
    Pt q;
    q.a = A::load(idx, g, k);
    q.b = B::load(idx, g, k);
    return q;
  }
  // a global dim belongs to exactly one of the two blocks, so one double2 carries either block's fetch
  template <int GD>
  HJ_DEV static double2 fetch(int i, const KGrid& g, const KSys& k) {
    if (GD < A::ND) return A::template fetch<GD>(i, g, k);
    return B::template fetch<GD>(i, g, k);
  }
  template <int GD>
  HJ_DEV static void apply(Pt& q, const double2 r, const KSys& k) {
    if (GD < A::ND) A::template apply<GD>(q.a, r, k);
    else B::template apply<GD>(q.b, r, k);
  }
  HJ_DEV static double ham(const Pt& q, const double* p, const KSys& k) {
    return A::ham(q.a, p, k) + B::ham(q.b, p, k);
  }
  HJ_DEV static double alpha(int dl, const Pt& q, const KSys& k) {
    return dl < A::ND ? A::alpha(dl, q.a, k) : B::alpha(dl - A::ND, q.b, k);
  }
  static constexpr unsigned alpha_dims(int dl) { return dl < A::ND ? A::alpha_dims(dl) : B::alpha_dims(dl - A::ND); }
};

// dimension-split trait: a product system is advanced as two passes, one per dim block (hj_vec_kernel.cuh)
template <class S>
struct SysSplit { static constexpr bool value = false; };
template <class A, class B>
struct SysSplit<PairF<A, B>> { static constexpr bool value = true; };

using SysDubinsRel = DubinsRelF<0, 0, 0>;
using SysDoubleInt = DoubleIntF<0, 0>;
using SysFlock = FlockF;
using SysDubinsRelPair = PairF<DubinsRelF<0, 0, 0>, DubinsRelF<3, HJ_DUBINS_NP, 2>>;
using SysDoubleIntPair = PairF<DoubleIntF<0, 0>, DoubleIntF<2, HJ_DINT_NP>>;
using SysFlockBatch = FlockBatchF;
using SysGenericDubinsCar = GenericF<DubinsCarDyn<0>>;
#define HJ_SYS_FLOCK_BATCH 100   // internal id: HJ_SYS_FLOCK registered on a batch context

// host-side dispatch helper: calls f.template operator()<Sys>() for the functor registered under `id`
template <class F>
inline bool hj_dispatch_system(int id, F&& f) {
  switch (id) {
    case HJ_SYS_DUBINS_REL: f.template operator()<SysDubinsRel>(); return true;
    case HJ_SYS_DOUBLE_INT: f.template operator()<SysDoubleInt>(); return true;
    case HJ_SYS_FLOCK: f.template operator()<SysFlock>(); return true;
    case HJ_SYS_DUBINS_REL_PAIR: f.template operator()<SysDubinsRelPair>(); return true;
    case HJ_SYS_DOUBLE_INT_PAIR: f.template operator()<SysDoubleIntPair>(); return true;
    case HJ_SYS_FLOCK_BATCH: f.template operator()<SysFlockBatch>(); return true;
    case HJ_SYS_GENERIC_DUBINS_CAR: f.template operator()<SysGenericDubinsCar>(); return true;
    default: return false;
  }
}
inline int hj_system_ndim(int id) {
  switch (id) {
    case HJ_SYS_DUBINS_REL: return 3;
    case HJ_SYS_DOUBLE_INT: return 2;
    case HJ_SYS_FLOCK: return 3;
    case HJ_SYS_DUBINS_REL_PAIR: return 6;
    case HJ_SYS_DOUBLE_INT_PAIR: return 4;
    case HJ_SYS_FLOCK_BATCH: return 4;   // grid dims incl. the batch dim
    case HJ_SYS_GENERIC_DUBINS_CAR: return 3;
    default: return -1;
  }
}
