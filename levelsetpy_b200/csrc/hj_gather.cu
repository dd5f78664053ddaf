// hj_gather.cu -- "gather" backend: one thread per grid node, neighbours fetched through L1/L2 with ghost
// cells generated on the fly.  Works for every shape/BC/system; it is the correctness baseline on the device,
// the backend of the per-operator entry points (hj_deriv, hj_add_ghost, hj_rhs on dense user arrays) and the
// fallback for shapes the TMA ring kernel does not take.  It is still a single fused kernel per RK stage:
// no padded copies, no derivative arrays, no separate reductions.
#include "hj_internal.h"
#include "hj_systems.cuh"

namespace {

constexpr int BX = 32, BY = 8;

template <int D>
HJ_DEV void decompose_outer(long long o, const KGrid& g, int* idx) {
#pragma unroll
  for (int d = D - 3; d >= 0; --d) {
    idx[d] = (int)(o % g.N[d]);
    o /= g.N[d];
  }
}

// One fused RHS / RK-stage kernel: for every node, for every dim: 7-point stencil with on-the-fly ghosts ->
// (derivL, derivR) -> derivC, R-L; then H(x, derivC), GLF dissipation, stage algebra, and (optionally) the
// derivative min/max, alpha max and NaN reductions.
template <class Sys, int WENO>
__global__ void __launch_bounds__(BX* BY, 2) k_stage_gather(const KGrid g, const KSys ks, const KStage st,
                                                          const long long nouter) {
  constexpr int D = Sys::ND;
  const int NX = g.N[D - 1], NY = g.N[D - 2];
  const int xt = (NX + BX - 1) / BX;
  const int ix = (blockIdx.x % xt) * BX + threadIdx.x;
  const int iy = (blockIdx.x / xt) * BY + threadIdx.y;
  const bool active = ix < NX && iy < NY;

  double inv_eps[D];
#pragma unroll
  for (int d = 0; d < D; ++d) inv_eps[d] = (WENO == HJ_WENO_INTENDED) ? inv_eps_from_max(st.epsmax[d]) : 0.0;

  RedAcc<D> acc;
  acc.init();

  for (long long o = blockIdx.y; o < nouter; o += gridDim.y) {
    if (!active) continue;
    int idx[D];
    decompose_outer<D>(o, g, idx);
    idx[D - 2] = iy;
    idx[D - 1] = ix;
    long long off = 0, ooff = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      off += (long long)idx[d] * g.stride[d];
      ooff += (long long)idx[d] * st.out_stride[d];
    }
    double pc[D], dd[D];
    double yin = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      double v[7], L, R;
      load_stencil(st.in + off, idx[d], g.N[d], g.stride[d], g.bc[d], g.slope_mult[d], v);
      upwind5<WENO>(v, g.dxinv[d], inv_eps[d], L, R, g.dx[d]);
      pc[d] = 0.5 * (L + R);           // term_lax_friedrich.py:108
      dd[d] = R - L;                   // artificial_diss_glf.py:90
      if (d == 0) yin = v[3];
      if (st.want_reduce) {
        acc.dmin[d] = fmin(acc.dmin[d], fmin(L, R));
        acc.dmax[d] = fmax(acc.dmax[d], fmax(L, R));
      }
    }
    const typename Sys::Pt pt = Sys::load(idx, g, ks);
    const double ham = Sys::ham(pt, pc, ks);
    double diss = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double a = Sys::alpha(d, pt, ks);
      diss += (0.5 * dd[d]) * a;       // artificial_diss_glf.py:100
      if (st.want_reduce) acc.amax[d] = fmax(acc.amax[d], a);
    }
    const double ydot = -(ham - diss); // term_lax_friedrich.py:124-128
    const double y = stage_update(st, yin, ydot, ooff);
    st.out[ooff] = y;
    if (st.want_reduce && y != y) acc.nan = 1;
  }
  if (st.want_reduce) acc.flush(st.red);
}

// upwindFirstWENO5a(grid, data, dim) -> derivL, derivR (dense in, dense out)
template <int WENO>
__global__ void __launch_bounds__(256) k_deriv(const KGrid g, const double* __restrict__ in, const int dim,
                                               double* __restrict__ dl, double* __restrict__ dr,
                                               const unsigned long long* epsmax, const long long n) {
  const double inv_eps = (WENO == HJ_WENO_INTENDED) ? inv_eps_from_max(epsmax[dim]) : 0.0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)((e / g.stride[dim]) % g.N[dim]);
    double v[7], L, R;
    load_stencil(in + e, i, g.N[dim], g.stride[dim], g.bc[dim], g.slope_mult[dim], v);
    upwind5<WENO>(v, g.dxinv[dim], inv_eps, L, R, g.dx[dim]);
    dl[e] = L;
    dr[e] = R;
  }
}

// upwindFirstENO3aHelper (ENO3aHelper.py:11): the six third-order candidates, out = [dL0, dL1, dL2, dR0, dR1, dR2][n]
__global__ void __launch_bounds__(256) k_deriv_all(const KGrid g, const double* __restrict__ in, const int dim,
                                                   double* __restrict__ out, const long long n) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)((e / g.stride[dim]) % g.N[dim]);
    double v[7], dL[3], dR[3];
    load_stencil(in + e, i, g.N[dim], g.stride[dim], g.bc[dim], g.slope_mult[dim], v);
    EnoTables T;
    eno_tables(v, g.dxinv[dim], T, true);
    eno3a_candidates(T, g.dx[dim], dL, dR);
#pragma unroll
    for (int k = 0; k < 3; ++k) { out[(long long)k * n + e] = dL[k]; out[(long long)(3 + k) * n + e] = dR[k]; }
  }
}

// addGhostExtrapolate / addGhostPeriodic: dense (.., N_dim, ..) -> dense (.., N_dim + 2*width, ..)
__global__ void __launch_bounds__(256) k_add_ghost(const KGrid g, const double* __restrict__ in, const int dim,
                                                   const int width, double* __restrict__ out, const long long nout) {
  const long long inner = g.stride[dim];           // dense: product of N[dim+1..]
  const int n = g.N[dim], no = n + 2 * width;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nout; e += (long long)gridDim.x * blockDim.x) {
    const long long in_ = e % inner;
    const long long t = e / inner;
    const int jo = (int)(t % no);
    const long long outer = t / no;
    const double* base = in + outer * (long long)n * inner + in_;
    const int j = jo - width;
    double val;
    if (j >= 0 && j < n) val = base[(long long)j * inner];
    else if (g.bc[dim] == HJ_BC_PERIODIC) val = base[(long long)((j < 0) ? j + n : j - n) * inner];
    else if (j < 0) val = ghost_extrapolate(base[0], base[inner], -j, g.slope_mult[dim]);
    else val = ghost_extrapolate(base[(long long)(n - 1) * inner], base[(long long)(n - 2) * inner], j - (n - 1),
                                 g.slope_mult[dim]);
    out[e] = val;
  }
}

// max_x alpha_d without a field (state-only partialFunc): artificial_diss_glf.py:104 evaluated once.  alpha_dl only
// depends on the dims Sys::alpha_dims(dl) names (e.g. relative Dubins: alpha_0 = |v_e - v_p cos x3| + |w x2|), so the
// maximum is taken over that sub-grid with every other index held at 0 -- the same set of alpha values, hence the
// same maximum bit for bit, from N^2 evaluations instead of N^D (41^6: 0.8 s -> microseconds).
template <class Sys>
__global__ void __launch_bounds__(256) k_alpha_max(const KGrid g, const KSys ks, unsigned long long* red, const int dl,
                                                   const unsigned mask, const long long count) {
  constexpr int D = Sys::BASE_DIM + Sys::ND;
  double amax = -INFINITY;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < count; e += (long long)gridDim.x * blockDim.x) {
    int idx[D];
    long long r = e;
#pragma unroll
    for (int d = D - 1; d >= 0; --d) {
      if (mask & (1u << d)) { idx[d] = (int)(r % g.N[d]); r /= g.N[d]; }
      else idx[d] = 0;
    }
    const typename Sys::Pt pt = Sys::load(idx, g, ks);
    amax = fmax(amax, Sys::alpha(dl, pt, ks));
  }
  amax = warp_max(amax);
  if (threadIdx.x % 32 == 0 && enc_ordered(amax) > *(const volatile unsigned long long*)(red + dl))
    atomicMax(red + dl, enc_ordered(amax));
}

// intended WENO: eps_d = 1e-6 * max(D1_d^2) + 1e-99 over the unstripped D1 table (upwind_first_weno5a.py:154-156).
// This prepass produces the raw maxima; the stage kernel turns them into 1/eps.
template <int D>
__global__ void __launch_bounds__(BX* BY) k_maxd1sq(const KGrid g, const double* __restrict__ in,
                                                     unsigned long long* epsmax, const long long nouter,
                                                     const int only_dim) {
  const int NX = g.N[D - 1], NY = g.N[D - 2];
  const int xt = (NX + BX - 1) / BX;
  const int ix = (blockIdx.x % xt) * BX + threadIdx.x;
  const int iy = (blockIdx.x / xt) * BY + threadIdx.y;
  const bool active = ix < NX && iy < NY;
  double mx[D];
#pragma unroll
  for (int d = 0; d < D; ++d) mx[d] = 0.0;
  // a CTA walks a CONTIGUOUS run of outer indices (the +1 neighbour along the slowest dims is then the next
  // iteration's node: an L1 / L2 hit), and the grid is a few CTAs per SM, not one CTA per 256 nodes
  const long long per = (nouter + gridDim.y - 1) / gridDim.y;
  const long long o_end = min(nouter, (long long)(blockIdx.y + 1) * per);
  for (long long o = (long long)blockIdx.y * per; o < o_end; ++o) {
    if (!active) continue;
    int idx[D];
    decompose_outer<D>(o, g, idx);
    idx[D - 2] = iy;
    idx[D - 1] = ix;
    long long off = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) off += (long long)idx[d] * g.stride[d];
    const double c = __ldg(in + off);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if (only_dim >= 0 && d != only_dim) continue;
      const int i = idx[d], n = g.N[d];
      if ((i > 0 && i < n - 1) || (g.bc[d] == HJ_BC_HALO && i < n - 1)) {
        const double dl = g.dxinv[d] * (__ldg(in + off + g.stride[d]) - c);
        mx[d] = fmax(mx[d], dl * dl);
        if (g.bc[d] == HJ_BC_HALO && i == 0) {   // pairs reaching into the lower halo planes belong to this rank's table
          double v[7];
          load_stencil(in + off, i, n, g.stride[d], g.bc[d], g.slope_mult[d], v);
          mx[d] = fmax(mx[d], d1sq_local(v, g.dxinv[d], 0, n + 1));
        }
      } else {
        double v[7];
        load_stencil(in + off, i, n, g.stride[d], g.bc[d], g.slope_mult[d], v);
        mx[d] = fmax(mx[d], d1sq_local(v, g.dxinv[d], i, n));
      }
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double m = warp_max(mx[d]);
    // only a value that raises the record is worth an atomic (see RedAcc::flush)
    if ((threadIdx.x + threadIdx.y * blockDim.x) % 32 == 0 && m > 0.0 &&
        enc_ordered(m) > *(const volatile unsigned long long*)(epsmax + d))
      atomicMax(epsmax + d, enc_ordered(m));
  }
}

// derivMin / derivMax of artificial_diss_glf.py:82-88 on their own: min and max over the grid of (derivL, derivR) of every
// dim, without a system -- what genericPartial (generic_partial.py:28-40) needs BEFORE the dissipation of the same RHS can
// be formed.  Reads the field once per dim through L1 / L2; the record is the one the stage kernels write (RedAcc).
template <int D, int WENO>
__global__ void __launch_bounds__(BX* BY) k_deriv_range(const KGrid g, const double* __restrict__ in,
                                                         const unsigned long long* epsmax, unsigned long long* red,
                                                         const long long nouter) {
  const int NX = g.N[D - 1], NY = g.N[D - 2];
  const int xt = (NX + BX - 1) / BX;
  const int ix = (blockIdx.x % xt) * BX + threadIdx.x;
  const int iy = (blockIdx.x / xt) * BY + threadIdx.y;
  const bool active = ix < NX && iy < NY;
  double inv_eps[D];
#pragma unroll
  for (int d = 0; d < D; ++d) inv_eps[d] = (WENO == HJ_WENO_INTENDED) ? inv_eps_from_max(epsmax[d]) : 0.0;
  RedAcc<D> acc;
  acc.init();
  const long long per = (nouter + gridDim.y - 1) / gridDim.y;
  const long long o_end = min(nouter, (long long)(blockIdx.y + 1) * per);
  for (long long o = (long long)blockIdx.y * per; o < o_end; ++o) {
    if (!active) continue;
    int idx[D];
    decompose_outer<D>(o, g, idx);
    idx[D - 2] = iy;
    idx[D - 1] = ix;
    long long off = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) off += (long long)idx[d] * g.stride[d];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      double v[7], L, R;
      load_stencil(in + off, idx[d], g.N[d], g.stride[d], g.bc[d], g.slope_mult[d], v);
      upwind5<WENO>(v, g.dxinv[d], inv_eps[d], L, R, g.dx[d]);
      acc.dmin[d] = fmin(acc.dmin[d], fmin(L, R));
      acc.dmax[d] = fmax(acc.dmax[d], fmax(L, R));
    }
  }
  acc.flush(red);
}

__global__ void k_init_reduce(unsigned long long* red, int D) {
  const int i = threadIdx.x;
  if (i < D) red[i] = enc_ordered(-INFINITY);
  else if (i < 2 * D) red[i] = enc_ordered(INFINITY);
  else if (i < 3 * D) red[i] = enc_ordered(-INFINITY);
  else if (i == 3 * D) red[i] = 0ull;
}
__global__ void k_init_eps(unsigned long long* eps, int D) {
  if ((int)threadIdx.x < D) eps[threadIdx.x] = enc_ordered(0.0);
}

// dense <-> pitched (rows of the innermost dim)
__global__ void __launch_bounds__(256) k_repitch(const double* __restrict__ src, double* __restrict__ dst, const int nx,
                                                 const long long rows, const long long src_pitch,
                                                 const long long dst_pitch) {
  const long long total = rows * nx;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / nx;
    const int x = (int)(e - r * nx);
    dst[r * dst_pitch + x] = src[r * src_pitch + x];
  }
}

// slab edge ranks: the 3 stored halo planes outside a non-periodic global boundary of dim 0 are the
// addGhostExtrapolate ghosts of this rank's own first / last two planes (add_ghost_extrapolate.py:88-110)
__global__ void __launch_bounds__(256) k_edge_halo(double* __restrict__ buf, const long long plane, const int n0,
                                                   const int side, const double m) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < plane; e += (long long)gridDim.x * blockDim.x) {
    if (side == 0) {
      const double edge = buf[(long long)HJ_GHOST * plane + e], next = buf[(long long)(HJ_GHOST + 1) * plane + e];
#pragma unroll
      for (int k = 0; k < HJ_GHOST; ++k) buf[(long long)k * plane + e] = ghost_extrapolate(edge, next, HJ_GHOST - k, m);
    } else {
      const double edge = buf[(long long)(HJ_GHOST + n0 - 1) * plane + e], next = buf[(long long)(HJ_GHOST + n0 - 2) * plane + e];
#pragma unroll
      for (int k = 0; k < HJ_GHOST; ++k) buf[(long long)(HJ_GHOST + n0 + k) * plane + e] = ghost_extrapolate(edge, next, k + 1, m);
    }
  }
}

template <int D>
long long outer_count(const KGrid& g) {
  long long n = 1;
  for (int d = 0; d <= D - 3; ++d) n *= g.N[d];
  return n;
}

inline dim3 tile_grid(const KGrid& g, int D, long long nouter) {
  const int xt = (g.N[D - 1] + BX - 1) / BX, yt = (g.N[D - 2] + BY - 1) / BY;
  return dim3((unsigned)(xt * yt), (unsigned)(nouter < 65535 ? nouter : 65535), 1);
}

// (X, Y) tiles x as many runs of the outer index as give ~16 CTAs per SM
inline dim3 walk_grid(const KGrid& g, int D, long long nouter) {
  const int xt = (g.N[D - 1] + BX - 1) / BX, yt = (g.N[D - 2] + BY - 1) / BY;
  long long gy = (148LL * 16 + (long long)xt * yt - 1) / ((long long)xt * yt);
  gy = gy < 1 ? 1 : (gy > nouter ? nouter : gy);
  return dim3((unsigned)(xt * yt), (unsigned)(gy < 65535 ? gy : 65535), 1);
}

inline int flat_blocks(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = 148LL * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// The reference's per-callable operators on dense arrays, for hosts that keep termLaxFriedrichs and swap single hooks
// (hamFunc, partialFunc, dissFunc):
//   OP 0  ham   = hamFunc(t, data, derivC, schemeData)                     (dubins_relative.py:63-88 etc.)
//   OP 1  alpha = partialFunc(t, data, derivMin, derivMax, schemeData, dl) as a dense array (state-only alphas)
//   OP 2  diss  = sum_d 0.5 (R_d - L_d) alpha_d, summed d = 0..D-1 from 0 like artificial_diss_glf.py:100, plus the
//         derivative min / max and max alpha reductions that give stepBound (:82-88, :104-109)
struct KPtrs {
  const double* a[HJ_MAX_DIM];
  const double* b[HJ_MAX_DIM];
};
template <class Sys, int OP>
__global__ void __launch_bounds__(256) k_sys_op(const KGrid g, const KSys ks, const KPtrs in, double* __restrict__ out,
                                                const int dl, unsigned long long* red, const long long n) {
  constexpr int D = Sys::ND;
  RedAcc<D> acc;
  acc.init();
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int idx[D];
    long long r = e;
#pragma unroll
    for (int d = D - 1; d >= 0; --d) { idx[d] = (int)(r % g.N[d]); r /= g.N[d]; }
    const typename Sys::Pt pt = Sys::load(idx, g, ks);
    if (OP == 0) {
      double pc[D];
#pragma unroll
      for (int d = 0; d < D; ++d) pc[d] = __ldg(in.a[d] + e);
      out[e] = Sys::ham(pt, pc, ks);
    } else if (OP == 1) {
      out[e] = Sys::alpha(dl, pt, ks);
    } else {
      double diss = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double L = __ldg(in.a[d] + e), Rr = __ldg(in.b[d] + e), al = Sys::alpha(d, pt, ks);
        diss = __dadd_rn(diss, __dmul_rn(__dmul_rn(0.5, __dsub_rn(Rr, L)), al));
        acc.dmin[d] = fmin(acc.dmin[d], fmin(L, Rr));
        acc.dmax[d] = fmax(acc.dmax[d], fmax(L, Rr));
        acc.amax[d] = fmax(acc.amax[d], al);
      }
      out[e] = diss;
    }
  }
  if (OP == 2) acc.flush(red);
}

struct SysOpLauncher {
  int op;
  const KGrid& g;
  const KSys& ks;
  const KPtrs& in;
  double* out;
  int dl;
  unsigned long long* red;
  cudaStream_t s;
  bool ok = true;
  template <class Sys>
  void operator()() {
    if constexpr (Sys::BASE_DIM == 0 && Sys::NSCRATCH == 0) {
      long long n = 1;
      for (int d = 0; d < Sys::ND; ++d) n *= g.N[d];
      switch (op) {
        case 0: k_sys_op<Sys, 0><<<flat_blocks(n), 256, 0, s>>>(g, ks, in, out, dl, red, n); break;
        case 1: k_sys_op<Sys, 1><<<flat_blocks(n), 256, 0, s>>>(g, ks, in, out, dl, red, n); break;
        case 2: k_sys_op<Sys, 2><<<flat_blocks(n), 256, 0, s>>>(g, ks, in, out, dl, red, n); break;
        default: ok = false;
      }
    } else {
      ok = false;
    }
  }
};

struct StageLauncher {
  int weno;
  const KGrid& g;
  const KSys& ks;
  const KStage& st;
  cudaStream_t s;
  bool ok = true;
  template <class Sys>
  void operator()() {
    if constexpr (Sys::BASE_DIM == 0) {
      constexpr int D = Sys::ND;
      const long long no = outer_count<D>(g);
      const dim3 grid = tile_grid(g, D, no), block(BX, BY);
      switch (weno) {
        case HJ_WENO_AS_SHIPPED: k_stage_gather<Sys, HJ_WENO_AS_SHIPPED><<<grid, block, 0, s>>>(g, ks, st, no); break;
        case HJ_WENO_INTENDED: k_stage_gather<Sys, HJ_WENO_INTENDED><<<grid, block, 0, s>>>(g, ks, st, no); break;
        case HJ_SCHEME_ENO3A: k_stage_gather<Sys, HJ_SCHEME_ENO3A><<<grid, block, 0, s>>>(g, ks, st, no); break;
        case HJ_SCHEME_ENO2: k_stage_gather<Sys, HJ_SCHEME_ENO2><<<grid, block, 0, s>>>(g, ks, st, no); break;
        default: ok = false;
      }
    } else {
      ok = false;    // batch functors only exist for the plane-ring kernel
    }
  }
};

struct AlphaLauncher {
  const KGrid& g;
  const KSys& ks;
  unsigned long long* red;
  cudaStream_t s;
  bool ok = true;
  int launches = 0;
  template <class Sys>
  void operator()() {
    if constexpr (Sys::BASE_DIM == 0) {
      for (int dl = 0; dl < Sys::ND; ++dl) {
        const unsigned mask = Sys::alpha_dims(dl);
        long long count = 1;
        for (int d = 0; d < Sys::ND; ++d)
          if (mask & (1u << d)) count *= g.N[d];
        k_alpha_max<Sys><<<flat_blocks(count), 256, 0, s>>>(g, ks, red, dl, mask, count);
        ++launches;
      }
    } else {
      ok = false;
    }
  }
};

}  // namespace

cudaError_t hj_launch_stage_gather(int system_id, int weno, const KGrid& g, const KSys& ks, const KStage& st,
                                   cudaStream_t s) {
  StageLauncher l{weno, g, ks, st, s};
  if (!hj_dispatch_system(system_id, l)) return cudaErrorInvalidValue;
  if (!l.ok) return cudaErrorNotSupported;
  hj_count_launch(1);
  return cudaGetLastError();
}

cudaError_t hj_launch_sys_op(int system_id, int op, const KGrid& g, const KSys& ks, const double* const* a,
                             const double* const* b, double* out, int dl, unsigned long long* red, cudaStream_t s) {
  KPtrs in{};
  for (int d = 0; d < g.D; ++d) {
    in.a[d] = a ? a[d] : nullptr;
    in.b[d] = b ? b[d] : nullptr;
  }
  SysOpLauncher l{op, g, ks, in, out, dl, red, s};
  if (!hj_dispatch_system(system_id, l)) return cudaErrorInvalidValue;
  if (!l.ok) return cudaErrorNotSupported;
  hj_count_launch(1);
  return cudaGetLastError();
}

cudaError_t hj_launch_deriv(int weno, const KGrid& g, const double* in, int dim, double* dl, double* dr,
                            const unsigned long long* epsmax, cudaStream_t s) {
  long long n = 1;
  for (int d = 0; d < g.D; ++d) n *= g.N[d];
  switch (weno) {
    case HJ_WENO_AS_SHIPPED: k_deriv<HJ_WENO_AS_SHIPPED><<<flat_blocks(n), 256, 0, s>>>(g, in, dim, dl, dr, epsmax, n); break;
    case HJ_WENO_INTENDED: k_deriv<HJ_WENO_INTENDED><<<flat_blocks(n), 256, 0, s>>>(g, in, dim, dl, dr, epsmax, n); break;
    case HJ_SCHEME_ENO3A: k_deriv<HJ_SCHEME_ENO3A><<<flat_blocks(n), 256, 0, s>>>(g, in, dim, dl, dr, epsmax, n); break;
    case HJ_SCHEME_ENO2: k_deriv<HJ_SCHEME_ENO2><<<flat_blocks(n), 256, 0, s>>>(g, in, dim, dl, dr, epsmax, n); break;
    default: return cudaErrorInvalidValue;
  }
  hj_count_launch(1);
  return cudaGetLastError();
}

cudaError_t hj_launch_deriv_all(const KGrid& g, const double* in, int dim, double* out6, cudaStream_t s) {
  long long n = 1;
  for (int d = 0; d < g.D; ++d) n *= g.N[d];
  k_deriv_all<<<flat_blocks(n), 256, 0, s>>>(g, in, dim, out6, n);
  hj_count_launch(1);
  return cudaGetLastError();
}

cudaError_t hj_launch_add_ghost(const KGrid& g, const double* in, int dim, int width, double* out, cudaStream_t s) {
  long long n = 1;
  for (int d = 0; d < g.D; ++d) n *= (d == dim) ? (g.N[d] + 2 * width) : g.N[d];
  k_add_ghost<<<flat_blocks(n), 256, 0, s>>>(g, in, dim, width, out, n);
  hj_count_launch(1);
  return cudaGetLastError();
}

cudaError_t hj_launch_alpha_max(int system_id, const KGrid& g, const KSys& ks, unsigned long long* red,
                                cudaStream_t s) {
  AlphaLauncher l{g, ks, red, s};
  if (!hj_dispatch_system(system_id, l)) return cudaErrorInvalidValue;
  if (!l.ok) return cudaErrorNotSupported;
  hj_count_launch(l.launches);
  return cudaGetLastError();
}

cudaError_t hj_launch_maxd1sq(const KGrid& g, const double* in, unsigned long long* epsmax, int only_dim,
                              cudaStream_t s) {
  const dim3 block(BX, BY);
  switch (g.D) {
    case 2: { long long no = outer_count<2>(g); k_maxd1sq<2><<<walk_grid(g, 2, no), block, 0, s>>>(g, in, epsmax, no, only_dim); break; }
    case 3: { long long no = outer_count<3>(g); k_maxd1sq<3><<<walk_grid(g, 3, no), block, 0, s>>>(g, in, epsmax, no, only_dim); break; }
    case 4: { long long no = outer_count<4>(g); k_maxd1sq<4><<<walk_grid(g, 4, no), block, 0, s>>>(g, in, epsmax, no, only_dim); break; }
    case 5: { long long no = outer_count<5>(g); k_maxd1sq<5><<<walk_grid(g, 5, no), block, 0, s>>>(g, in, epsmax, no, only_dim); break; }
    case 6: { long long no = outer_count<6>(g); k_maxd1sq<6><<<walk_grid(g, 6, no), block, 0, s>>>(g, in, epsmax, no, only_dim); break; }
    default: return cudaErrorInvalidValue;
  }
  hj_count_launch(1);
  return cudaGetLastError();
}

template <int D>
static cudaError_t launch_deriv_range(int weno, const KGrid& g, const double* in, const unsigned long long* epsmax,
                                      unsigned long long* red, cudaStream_t s) {
  const long long no = outer_count<D>(g);
  const dim3 grid = walk_grid(g, D, no), block(BX, BY);
  switch (weno) {
    case HJ_WENO_AS_SHIPPED: k_deriv_range<D, HJ_WENO_AS_SHIPPED><<<grid, block, 0, s>>>(g, in, epsmax, red, no); break;
    case HJ_WENO_INTENDED: k_deriv_range<D, HJ_WENO_INTENDED><<<grid, block, 0, s>>>(g, in, epsmax, red, no); break;
    case HJ_SCHEME_ENO3A: k_deriv_range<D, HJ_SCHEME_ENO3A><<<grid, block, 0, s>>>(g, in, epsmax, red, no); break;
    case HJ_SCHEME_ENO2: k_deriv_range<D, HJ_SCHEME_ENO2><<<grid, block, 0, s>>>(g, in, epsmax, red, no); break;
    default: return cudaErrorNotSupported;
  }
  return cudaGetLastError();
}

cudaError_t hj_launch_deriv_range(int weno, const KGrid& g, const double* in, const unsigned long long* epsmax,
                                  unsigned long long* red, cudaStream_t s) {
  cudaError_t e;
  switch (g.D) {
    case 2: e = launch_deriv_range<2>(weno, g, in, epsmax, red, s); break;
    case 3: e = launch_deriv_range<3>(weno, g, in, epsmax, red, s); break;
    case 4: e = launch_deriv_range<4>(weno, g, in, epsmax, red, s); break;
    case 5: e = launch_deriv_range<5>(weno, g, in, epsmax, red, s); break;
    case 6: e = launch_deriv_range<6>(weno, g, in, epsmax, red, s); break;
    default: return cudaErrorInvalidValue;
  }
  hj_count_launch(1);
  return e;
}

cudaError_t hj_launch_init_reduce(unsigned long long* red, int D, cudaStream_t s) {
  k_init_reduce<<<1, 32, 0, s>>>(red, D);
  hj_count_launch(1);
  return cudaGetLastError();
}
cudaError_t hj_launch_init_eps(unsigned long long* eps, int D, cudaStream_t s) {
  k_init_eps<<<1, 32, 0, s>>>(eps, D);
  hj_count_launch(1);
  return cudaGetLastError();
}

cudaError_t hj_launch_edge_halo(double* buf, long long plane, int n0, int side, double m, cudaStream_t s) {
  k_edge_halo<<<flat_blocks(plane), 256, 0, s>>>(buf, plane, n0, side, m);
  hj_count_launch(1);
  return cudaGetLastError();
}

// driver epilogue helpers (hji_solver.py:603-672), pointwise over the pitched interior of a field
// max |a - b| and a NaN flag of a (the change since the last frame: stopConverge, :661-672; the NaN check of :544)
__global__ void __launch_bounds__(256) k_change(const double* __restrict__ a, const double* __restrict__ b, long long n,
                                                unsigned long long* red) {
  double m = 0.0;
  int nan = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double x = a[i];
    if (x != x) nan = 1;
    else m = fmax(m, fabs(x - b[i]));
  }
  m = warp_max(m);
  nan = __any_sync(0xffffffffu, nan);
  if (threadIdx.x % 32 == 0) {
    atomicMax(red, enc_ordered(m));
    if (nan) atomicOr(red + 1, 1ull);
  }
}
// discounting: mode 0 (:603-611) y = gamma y + (1 - gamma) ref;  mode 1 ("Kene", :615-637) shift below zero by
// max |l|, discount, min / max with the shifted target, shift back
__global__ void __launch_bounds__(256) k_discount(double* __restrict__ y, const double* __restrict__ ref, long long n,
                                                  double gamma, int mode, int take_max, double max_val) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = y[i];
    if (mode == 2) {                 // the obstacle mask on its own (:641-644): y = max(y, -obstacle)
      v = nan_max(v, -ref[i]);
    } else if (mode == 0) {
      v = __dmul_rn(v, gamma);
      v = __dadd_rn(v, __dmul_rn(1.0 - gamma, ref[i]));
    } else {
      double yt = __dmul_rn(__dsub_rn(v, max_val), gamma);
      const double tt = __dsub_rn(ref[i], max_val);
      yt = take_max ? nan_max(yt, tt) : nan_min(yt, tt);
      v = __dadd_rn(yt, max_val);
    }
    y[i] = v;
  }
}

cudaError_t hj_launch_change(const double* a, const double* b, long long n, unsigned long long* red, cudaStream_t s) {
  k_change<<<flat_blocks(n), 256, 0, s>>>(a, b, n, red);
  hj_count_launch(1);
  return cudaGetLastError();
}
cudaError_t hj_launch_discount(double* y, const double* ref, long long n, double gamma, int mode, int take_max,
                               double max_val, cudaStream_t s) {
  k_discount<<<flat_blocks(n), 256, 0, s>>>(y, ref, n, gamma, mode, take_max, max_val);
  hj_count_launch(1);
  return cudaGetLastError();
}

static long long rows_of(const KGrid& g) {
  long long r = 1;
  for (int d = 0; d < g.D - 1; ++d) r *= g.N[d];
  return r;
}

cudaError_t hj_launch_pack(const double* dense, double* pitched, const KGrid& gd, const KGrid& gp, cudaStream_t s) {
  const long long rows = rows_of(gd);
  const int nx = gd.N[gd.D - 1];
  k_repitch<<<flat_blocks(rows * nx), 256, 0, s>>>(dense, pitched, nx, rows, gd.stride[gd.D - 2], gp.stride[gp.D - 2]);
  hj_count_launch(1);
  return cudaGetLastError();
}
cudaError_t hj_launch_unpack(const double* pitched, double* dense, const KGrid& gd, const KGrid& gp, cudaStream_t s) {
  const long long rows = rows_of(gd);
  const int nx = gd.N[gd.D - 1];
  k_repitch<<<flat_blocks(rows * nx), 256, 0, s>>>(pitched, dense, nx, rows, gp.stride[gp.D - 2], gd.stride[gd.D - 2]);
  hj_count_launch(1);
  return cudaGetLastError();
}
