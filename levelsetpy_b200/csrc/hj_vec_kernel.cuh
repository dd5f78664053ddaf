// hj_vec_kernel.cuh -- second pass of the dimension-split path for product systems (device side).
//
// A product system (the 4-D double-integrator pair, the 6-D relative-Dubins pair) has
//     H(x, p) = H_A(x_A, p_A) + H_B(x_B, p_B),   alpha_d = alpha_d(x_block(d)),
// so one RHS is  F(y) = F_A(y) + F_B(y)  where F_A only differentiates along the leading dim block A and F_B along
// the trailing (contiguous) block B.  A star stencil of radius 3 in 6 dims has no tile that fits shared memory or L2
// with useful reuse in all dims, but each block alone is a 2-D/3-D stencil.  The stage is therefore two kernels:
//     pass 1 (k_stage_tma on block B):  tmp = F_B(in)                                             8 B r + 8 B w
//     pass 2 (this kernel, block A):    out = RK_s(y0, in + dt * (tmp + F_A(in)))  [+ epilogue]   16(24) B r + 8 B w
// i.e. 40 / 48 / 48 B per node for the three stages instead of the 16 / 24 / 24 B of a (hypothetical) fully fused
// 6-D tile, but every byte is streamed once, coalesced, at HBM speed.
//
// Layout of pass 2: the trailing dims are flattened into one contiguous "vector" axis of V = stride[NS-1] elements
// that carries no stencil.  A CTA owns VB consecutive vector elements (VP thread pairs) times a TA x TB tile of the
// first block dims and marches along the last block dim.  Each plane of the tile with its 3-cell halo in the tiled
// dims -- a box {VB, 1, TB+6, TA+6} of a 4-D tensor map [V, N2, N1, N0] (3-D for a 2-dim block) -- arrives by one
// TMA load into an R-slot ring; the marching dim uses a 3-deep register queue + the ring slots of planes z+1..z+3,
// exactly like the plane-ring kernel.  Both nodes of a thread's pair share their block-A state (x_A does not vary along the vector
// axis).  Ghost cells are made in registers from the TMA zero fill, as in hj_tma_kernel.cuh.
#pragma once
#include "hj_tma_kernel.cuh"

struct VecGeom {
  int nvc;          // chunks along the vector axis
  int vc0;          // first chunk this launch advances (pass 2 in column pieces: hj_stage_pass_cols)
  int nta, ntb;     // tiles along the two tiled block dims (ntb = 1 for a 2-dim block)
  int nzc, cz;      // chunks / planes per chunk along the marching block dim
  int zcoord0;      // TMA coordinate shift of dim 0 (stored halo planes of a slab context)
  int pitch, NX;    // innermost padded / true extent: pad columns are excluded from stores and reductions
};

// The marching dim MD is the LAST dim of the block: for the relative-Dubins block that is the periodic heading, whose
// wrap-around costs nothing when it is the marched dim (the ring simply loads plane z +- N), and on a slab context
// dim 0 (thin, with stored halo planes) is then a tiled dim that one tile covers.
// GW_: ghost warp, as in TmaCfg -- one extra warp writes the ghost rows of the tiled dims into the landed slot.
template <int NS_, int R_, int MINB_, int VP_, int TA_, int TB_, int GW_ = 0, int OPT_ = 0>
struct VecCfg {
  static constexpr int OPT = OPT_;        // 256: tuning harness only -- the ghost warp forwards the barrier, no fill
  static constexpr int NGW = GW_;
  static constexpr bool GW = GW_ > 0;
  static constexpr int NS = NS_;          // dims of the leading block (2 or 3)
  static constexpr int MD = NS_ - 1;      // marching block dim
  static constexpr int DA = 0, DB = NS_ == 3 ? 1 : -1;             // tiled block dims (DA slower in memory)
  static constexpr int R = R_, MINB = MINB_;
  static constexpr int VP = VP_, VB = 2 * VP_;                    // thread pairs / doubles along the vector axis
  static constexpr int TA = TA_, TB = NS_ == 3 ? TB_ : 1;          // tile of the tiled dims
  static constexpr int HA = TA + 6, HB = NS_ == 3 ? TB + 6 : 1;    // haloed tile
  static constexpr int SB = VB, SA = HB * VB;                      // slot strides (doubles) of dims DB, DA
  static constexpr int NACTIVE = VP * TA * TB;
  static constexpr int NCONS = (NACTIVE + 31) / 32 * 32;
  static constexpr int NTHREADS = NCONS + 32 * GW_;
  static constexpr int BOX = HA * HB * VB;
  static constexpr int SLOT = (BOX + 15) / 16 * 16;
  static constexpr size_t smem_bytes() { return (size_t)R * SLOT * 8 + 3 * R * 8; }
};

namespace hjtma {

HJ_DEV void tma_load_4d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"((uint64_t)tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <class Blk, int GD, int WENO, bool RED, int STAGE, class Cfg>
__global__ void __launch_bounds__(Cfg::NTHREADS, Cfg::MINB)
k_stage_vec(const __grid_constant__ CUtensorMap tmap, const KGrid g, const KSys ks, const KStage st, const VecGeom geo) {
  constexpr int NS = Cfg::NS, MD = Cfg::MD, DA = Cfg::DA, DB = Cfg::DB >= 0 ? Cfg::DB : 0;
  static_assert(Blk::BASE_DIM == 0 && Blk::ND == NS && NS < GD, "pass 2 takes the leading dim block");
  static_assert(STAGE >= 1 && STAGE <= 3, "RK stages only");
  constexpr int R = Cfg::R, SLOT = Cfg::SLOT, VB = Cfg::VB, VP = Cfg::VP, TA = Cfg::TA, TB = Cfg::TB;
  constexpr int SA = Cfg::SA, SB = Cfg::SB;
  constexpr int NWARPS = Cfg::NCONS / 32;
  static_assert((SLOT * 8) % 128 == 0, "slot must keep 128-byte alignment");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);
  const uint32_t ring_s = smem_u32(smem_raw);
  const uint32_t full_s = ring_s + R * SLOT * 8;
  const uint32_t empty_s = full_s + R * 8;
  const uint32_t ready_s = empty_s + R * 8;

  const int tid = threadIdx.x;
  long long b = blockIdx.x;
  const int tb = (int)(b % geo.ntb); b /= geo.ntb;
  const int ta = (int)(b % geo.nta); b /= geo.nta;
  const int zc = (int)(b % geo.nzc); b /= geo.nzc;
  const int v0 = ((int)b + geo.vc0) * VB;
  const int NM = g.N[MD], NA = g.N[DA], NB = NS == 3 ? g.N[DB] : 1;
  const long long V = g.stride[NS - 1];
  const int ia0 = ta * TA, ib0 = tb * TB, z0 = zc * geo.cz;
  const int z1 = min(z0 + geo.cz, NM);
  const int bcm = g.bc[MD], bca = g.bc[DA], bcb = NS == 3 ? g.bc[DB] : HJ_BC_EXTRAPOLATE;
  const unsigned klast = (unsigned)((z1 - 1 + 3) - (z0 - 3));
  // stored halo planes of dim 0 (slab context) shift its TMA coordinate; dim 0 is the tiled dim DA
  const int ca0 = ia0 - 3 + geo.zcoord0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < R; ++s) {
      mbar_init(full_s + 8 * s, 1);
      mbar_init(empty_s + 8 * s, NWARPS);
      mbar_init(ready_s + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto tma_plane = [&](unsigned s, int zsrc) {
    const uint32_t fb = full_s + 8 * s;
    mbar_expect_tx(fb, Cfg::BOX * 8);
    // tensor dims, fastest first: [V, N2, N1, N0] resp. [V, N1, N0]; the marching dim is the block's last dim
    if constexpr (NS == 3) tma_load_4d(ring_s + s * (SLOT * 8), &tmap, fb, v0, zsrc, ib0 - 3, ca0);
    else tma_load_3d(ring_s + s * (SLOT * 8), &tmap, fb, v0, zsrc, ca0);
  };
  // plane with ring position k -> slot s: TMA load, or a bare arrival for a computed ghost plane
  auto issue = [&](unsigned k, unsigned s) {
    const int zp = z0 - 3 + (int)k;
    int zsrc = zp;
    bool load = true;
    if (zp < 0 || zp >= NM) {
      if (bcm == HJ_BC_PERIODIC) zsrc = zp < 0 ? zp + NM : zp - NM;
      else load = false;                                        // extrapolated ghost plane (the marched dim is never dim 0)
    }
    if (load) tma_plane(s, zsrc);
    else mbar_arrive(full_s + 8 * s);
  };
  if (tid == 0) {
    for (unsigned k = 0; k < (unsigned)R && k <= klast; ++k) issue(k, k);
  }

  const int lane = tid & 31;
  const bool need_patch_a = bca != HJ_BC_HALO && (ia0 - 3 < 0 || ia0 + TA + 2 >= NA);
  const bool need_patch_b = NS == 3 && (ib0 - 3 < 0 || ib0 + TB + 2 >= NB);

  // ================================================================== ghost warp (see k_stage_tma)
  // Ghost rows of the tiled dims: for dim A rows a = -3..-1 / NA..NA+2 over the tile's B rows, for dim B likewise over
  // the tile's A rows; every row is VB doubles, moved as double2.
  const bool gw_on = Cfg::GW && (need_patch_a || need_patch_b);
  if constexpr (Cfg::GW) {
    if (tid >= Cfg::NCONS) {
      if (!gw_on) return;
      const double ma = g.slope_mult[DA], mb = g.slope_mult[DB];
      const bool whole_a = NA <= TA, whole_b = NB <= TB;
      const int na = min(TA, NA - ia0), nb = NS == 3 ? min(TB, NB - ib0) : 1;      // tile rows inside the grid
      for (unsigned k = (unsigned)(tid - Cfg::NCONS) / 32; k <= klast; k += Cfg::NGW) {
        const unsigned s = k % R;
        mbar_wait(full_s + 8 * s, (k / R) & 1);
        const int zp = z0 - 3 + (int)k;
        int zsrc = zp;
        bool loaded = true;
        if (zp < 0 || zp >= NM) {
          if (bcm == HJ_BC_PERIODIC) zsrc = zp < 0 ? zp + NM : zp - NM;
          else loaded = false;
        }
        if ((Cfg::OPT & 256) == 0 && loaded && k >= 3 && k + 3 <= klast) {
          double* sl = ring + (size_t)s * SLOT;
          const double* gplane = st.in + (long long)zsrc * g.stride[MD] + v0;
          auto ghost2 = [](double2 e, double2 n, int dist, double m) {
            return make_double2(ghost_extrapolate(e.x, n.x, dist, m), ghost_extrapolate(e.y, n.y, dist, m));
          };
          // one lane per (row of the other tiled dim, vector pair): the two edge values give the three ghost rows
          if (need_patch_a) {
            for (int it = lane; it < nb * VP; it += 32) {
              const int bb = it / VP, vp2 = it % VP;
              double* colp = sl + (NS == 3 ? bb + 3 : 0) * SB + 2 * vp2 + (3 - ia0) * SA;   // colp[a * SA] = row a of dim A
              const double* gcol = gplane + (NS == 3 ? (long long)(ib0 + bb) * g.stride[DB] : 0) + 2 * vp2;
              if (ia0 == 0) {
                if (bca == HJ_BC_PERIODIC) {
#pragma unroll
                  for (int a = -3; a < 0; ++a)
                    *reinterpret_cast<double2*>(colp + a * SA) = whole_a ? lds2(colp + (a + NA) * SA) : ldg2(gcol + (long long)(a + NA) * g.stride[DA]);
                } else {
                  const double2 e0 = lds2(colp), e1 = lds2(colp + SA);
#pragma unroll
                  for (int a = -3; a < 0; ++a) *reinterpret_cast<double2*>(colp + a * SA) = ghost2(e0, e1, -a, ma);
                }
              }
              if (ia0 + TA + 2 >= NA) {
                const int amax = min(NA + 2, ia0 + TA + 2);
                if (bca == HJ_BC_PERIODIC) {
                  for (int a = NA; a <= amax; ++a)
                    *reinterpret_cast<double2*>(colp + a * SA) = whole_a ? lds2(colp + (a - NA) * SA) : ldg2(gcol + (long long)(a - NA) * g.stride[DA]);
                } else {
                  const double2 f0 = lds2(colp + (NA - 1) * SA), f1 = lds2(colp + (NA - 2) * SA);
                  for (int a = NA; a <= amax; ++a) *reinterpret_cast<double2*>(colp + a * SA) = ghost2(f0, f1, a - (NA - 1), ma);
                }
              }
            }
          }
          if (NS == 3 && need_patch_b) {
            for (int it = lane; it < na * VP; it += 32) {
              const int aa2 = it / VP, vp2 = it % VP;
              double* colp = sl + (aa2 + 3) * SA + 2 * vp2 + (3 - ib0) * SB;                 // colp[b * SB] = row b of dim B
              const double* gcol = gplane + (long long)(ia0 + aa2) * g.stride[DA] + 2 * vp2;
              if (ib0 == 0) {
                if (bcb == HJ_BC_PERIODIC) {
#pragma unroll
                  for (int q2 = -3; q2 < 0; ++q2)
                    *reinterpret_cast<double2*>(colp + q2 * SB) = whole_b ? lds2(colp + (q2 + NB) * SB) : ldg2(gcol + (long long)(q2 + NB) * g.stride[DB]);
                } else {
                  const double2 e0 = lds2(colp), e1 = lds2(colp + SB);
#pragma unroll
                  for (int q2 = -3; q2 < 0; ++q2) *reinterpret_cast<double2*>(colp + q2 * SB) = ghost2(e0, e1, -q2, mb);
                }
              }
              if (ib0 + TB + 2 >= NB) {
                const int bmax = min(NB + 2, ib0 + TB + 2);
                if (bcb == HJ_BC_PERIODIC) {
                  for (int q2 = NB; q2 <= bmax; ++q2)
                    *reinterpret_cast<double2*>(colp + q2 * SB) = whole_b ? lds2(colp + (q2 - NB) * SB) : ldg2(gcol + (long long)(q2 - NB) * g.stride[DB]);
                } else {
                  const double2 f0 = lds2(colp + (NB - 1) * SB), f1 = lds2(colp + (NB - 2) * SB);
                  for (int q2 = NB; q2 <= bmax; ++q2) *reinterpret_cast<double2*>(colp + q2 * SB) = ghost2(f0, f1, q2 - (NB - 1), mb);
                }
              }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(ready_s + 8 * s);
      }
      return;
    }
  }
  const uint32_t land_s = gw_on ? ready_s : full_s;

  // ================================================================== consumers
  const bool live = tid < Cfg::NACTIVE;
  const int vp = tid % VP;
  const int pos = live ? tid / VP : 0;
  const int ab = pos % TB, aa = pos / TB;
  const long long iv = (long long)v0 + 2 * vp;
  const int ia = ia0 + aa, ib = ib0 + ab;
  const bool inb = live && iv < V && ia < NA && ib < NB;
  const int xcol = (int)(iv % geo.pitch);                    // innermost index of node A: pad columns are not nodes
  const bool ok0 = inb && xcol < geo.NX, ok1 = inb && xcol + 1 < geo.NX;
  long long off = (long long)z0 * g.stride[MD] + (long long)ia * g.stride[DA] + iv;
  if (NS == 3) off += (long long)ib * g.stride[DB];
  const long long zstride = g.stride[MD];

  double inv_eps[NS];
#pragma unroll
  for (int d = 0; d < NS; ++d) inv_eps[d] = (WENO == HJ_WENO_INTENDED) ? inv_eps_from_max(st.epsmax[d]) : 0.0;
  int idx[GD];
#pragma unroll
  for (int d = 0; d < GD; ++d) idx[d] = 0;
  idx[MD] = z0;
  idx[DA] = min(ia, NA - 1);
  if (NS == 3) idx[DB] = min(ib, NB - 1);
  typename Blk::Pt pt = Blk::load(idx, g, ks);               // x_A is shared by the two nodes of my pair

  const int myoff = ((aa + 3) * Cfg::HB + (NS == 3 ? ab + 3 : 0)) * VB + 2 * vp;
  RedAcc<GD> acc;
  acc.init();
  // does this thread's dim-0 plane belong to a neighbour's halo?  (dim 0 is the tiled dim DA; loop-invariant)
  const bool push_lo = st.push_lo != nullptr && ia < HJ_GHOST;
  const bool push_hi = st.push_hi != nullptr && ia >= NA - HJ_GHOST;

  // ---- prologue (see k_stage_tma)
  double2 q[3];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    mbar_wait(land_s + 8 * k, 0);
    if (k < 3) q[k] = lds2(ring + (size_t)k * SLOT + myoff);
  }
  if (bcm == HJ_BC_EXTRAPOLATE && z0 == 0) {
    const double2 e0 = lds2(ring + (size_t)3 * SLOT + myoff), e1 = lds2(ring + (size_t)4 * SLOT + myoff);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      q[k].x = ghost_extrapolate(e0.x, e1.x, 3 - k, g.slope_mult[MD]);
      q[k].y = ghost_extrapolate(e0.y, e1.y, 3 - k, g.slope_mult[MD]);
    }
  }
  __syncwarp();
  if (lane == 0) { mbar_arrive(empty_s + 0); mbar_arrive(empty_s + 8); mbar_arrive(empty_s + 16); }
  if (tid == 0) {
    for (unsigned k = R; k < (unsigned)R + 3 && k <= klast; ++k) {
      mbar_wait(empty_s + 8 * (k - R), 0);
      issue(k, k - R);
    }
  }

  // ---- march
  unsigned kc = 3;
  unsigned s_prev = 2 % R, s_cur = 3 % R, s_p1 = 4 % R, s_p2 = 5 % R, s_new = 6 % R;
  unsigned p_prev = 0, p_cur = 0, p_new = (6 / R) & 1;
  int z = z0;
  double2 raw_next = Blk::template fetch<MD>(z0, g, ks);
  const double2 zero2 = make_double2(0.0, 0.0);
  double2 tmp_next = inb ? ldg2(st.tmp + off) : zero2;       // pass-1 result F_B(in): fetched one plane ahead
  double2 y0_next = (STAGE >= 2 && inb) ? ldg2(st.y0 + off) : zero2;   // so is y0 (stage 3 overwrites it in place, but
                                                             // only at this thread's own node of the CURRENT plane)

  // SIMPLE (fast march only): the common epilogue -- no termRestrictUpdate; stage 3 = minVOverTime without obstacle --
  // compiled in, so that the steady-state loop carries no epilogue dispatch (as in k_stage_tma)
  auto plane = [&]<bool FAST, bool SIMPLE = false>() {
    if constexpr (FAST) {
      if (tid == 0) {
        mbar_wait(empty_s + 8 * s_prev, p_prev);
        tma_plane(s_prev, z + R - 1);
      }
    } else {
      if (tid == 0 && kc >= 4 && kc - 1 + R <= klast) {
        mbar_wait(empty_s + 8 * s_prev, p_prev);
        issue(kc - 1 + R, s_prev);
      }
    }
    Blk::template apply<MD>(pt, raw_next, ks);
    raw_next = Blk::template fetch<MD>(min(z + 1, NM - 1), g, ks);
    // pointwise streams: issued first, consumed last
    const double2 tmpv = tmp_next;
    const double2 y0v = y0_next;
    if (inb && z + 1 < z1) {
      tmp_next = ldg2(st.tmp + off + zstride);
      if (STAGE >= 2) y0_next = ldg2(st.y0 + off + zstride);
    }
    double2 auxv = zero2, obsv = zero2;
    if (STAGE == 3 && !SIMPLE && inb) {
      if (st.comp == HJ_COMP_MIN_WITH_AUX || st.comp == HJ_COMP_MAX_WITH_AUX) auxv = ldg2(st.aux + off);
      if (st.use_obs) obsv = ldg2(st.obs + off);
    }

    double pcA[NS], hdA[NS], pcB[NS], hdB[NS];
    double L, Rr;
    constexpr bool red = RED;
#define HJ_RED(d, ok)                                              \
  if (red && (ok)) {                                               \
    acc.dmin[d] = fmin(acc.dmin[d], fmin(L, Rr));                  \
    acc.dmax[d] = fmax(acc.dmax[d], fmax(L, Rr));                  \
  }
    // One dim at a time (loads next to their use, compiler barriers in between): three haloed dims at once would
    // need 72 registers of neighbour data.
    const double* cur = ring + (size_t)s_cur * SLOT + myoff;
    const double2 ctr = lds2(cur);
    {  // tiled dim DA
      double2 am3 = lds2(cur - 3 * SA), am2 = lds2(cur - 2 * SA), am1 = lds2(cur - 1 * SA);
      double2 ap1 = lds2(cur + 1 * SA), ap2 = lds2(cur + 2 * SA), ap3 = lds2(cur + 3 * SA);
      if constexpr (!FAST) {
        if (!gw_on && need_patch_a && inb)
          patch_y<SA>(am3, am2, am1, ap1, ap2, ap3, ia, ia0, NA, bca, g.slope_mult[DA], cur - (aa + 3) * SA,
                      st.in + off - (long long)ia * g.stride[DA], g.stride[DA], NA <= TA);
      }
      pc_hd<WENO>(am3.x, am2.x, am1.x, ctr.x, ap1.x, ap2.x, ap3.x, g, DA, inv_eps[DA], pcA[DA], hdA[DA], L, Rr, red);
      HJ_RED(DA, ok0)
      pc_hd<WENO>(am3.y, am2.y, am1.y, ctr.y, ap1.y, ap2.y, ap3.y, g, DA, inv_eps[DA], pcB[DA], hdB[DA], L, Rr, red);
      HJ_RED(DA, ok1)
    }
    asm volatile("" ::: "memory");
    if constexpr (NS == 3) {  // tiled dim DB
      double2 bm3 = lds2(cur - 3 * SB), bm2 = lds2(cur - 2 * SB), bm1 = lds2(cur - 1 * SB);
      double2 bp1 = lds2(cur + 1 * SB), bp2 = lds2(cur + 2 * SB), bp3 = lds2(cur + 3 * SB);
      if constexpr (!FAST) {
        if (!gw_on && need_patch_b && inb)
          patch_y<SB>(bm3, bm2, bm1, bp1, bp2, bp3, ib, ib0, NB, bcb, g.slope_mult[DB], cur - (ab + 3) * SB,
                      st.in + off - (long long)ib * g.stride[DB], g.stride[DB], NB <= TB);
      }
      pc_hd<WENO>(bm3.x, bm2.x, bm1.x, ctr.x, bp1.x, bp2.x, bp3.x, g, DB, inv_eps[DB], pcA[DB], hdA[DB], L, Rr, red);
      HJ_RED(DB, ok0)
      pc_hd<WENO>(bm3.y, bm2.y, bm1.y, ctr.y, bp1.y, bp2.y, bp3.y, g, DB, inv_eps[DB], pcB[DB], hdB[DB], L, Rr, red);
      HJ_RED(DB, ok1)
      asm volatile("" ::: "memory");
    }
    {  // marching dim: planes z+1, z+2 landed earlier; plane z+3 is the newest one of the ring
      double2 zp1 = lds2(ring + (size_t)s_p1 * SLOT + myoff), zp2 = lds2(ring + (size_t)s_p2 * SLOT + myoff), zp3;
      mbar_wait(land_s + 8 * s_new, p_new);
      zp3 = lds2(ring + (size_t)s_new * SLOT + myoff);
      if constexpr (!FAST) {
        if (bcm == HJ_BC_EXTRAPOLATE && z + 3 >= NM) {
          const int ke = NM - 1 - z;
          const double2 ed = ke == 0 ? ctr : (ke == 1 ? zp1 : zp2);
          const double2 nx = ke == 0 ? q[2] : (ke == 1 ? ctr : zp1);
          const double m = g.slope_mult[MD];
          if (ke < 1) zp1 = make_double2(ghost_extrapolate(ed.x, nx.x, 1 - ke, m), ghost_extrapolate(ed.y, nx.y, 1 - ke, m));
          if (ke < 2) zp2 = make_double2(ghost_extrapolate(ed.x, nx.x, 2 - ke, m), ghost_extrapolate(ed.y, nx.y, 2 - ke, m));
          zp3 = make_double2(ghost_extrapolate(ed.x, nx.x, 3 - ke, m), ghost_extrapolate(ed.y, nx.y, 3 - ke, m));
        }
      }
      // this warp is done with the current plane's slot
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_s + 8 * s_cur);
      pc_hd<WENO>(q[0].x, q[1].x, q[2].x, ctr.x, zp1.x, zp2.x, zp3.x, g, MD, inv_eps[MD], pcA[MD], hdA[MD], L, Rr, red);
      HJ_RED(MD, ok0)
      pc_hd<WENO>(q[0].y, q[1].y, q[2].y, ctr.y, zp1.y, zp2.y, zp3.y, g, MD, inv_eps[MD], pcB[MD], hdB[MD], L, Rr, red);
      HJ_RED(MD, ok1)
    }
#undef HJ_RED

    double ydA = -Blk::ham(pt, pcA, ks), ydB = -Blk::ham(pt, pcB, ks);
#pragma unroll
    for (int d = 0; d < NS; ++d) {
      const double a = Blk::alpha(d, pt, ks);
      ydA = fma(hdA[d], a, ydA);
      ydB = fma(hdB[d], a, ydB);
      if (red && ok0) acc.amax[d] = fmax(acc.amax[d], a);
    }

    // tmp holds F_B(in) from pass 1: total ydot, termRestrictUpdate, then the RK algebra + driver epilogue
    ydA = tmpv.x + ydA;
    ydB = tmpv.y + ydB;
    if constexpr (!SIMPLE) {
      ydA = restrict_update(ydA, st.restrict_sign);
      ydB = restrict_update(ydB, st.restrict_sign);
    }
    const double vA = ctr.x + st.dt * ydA, vB = ctr.y + st.dt * ydB;
    double oA, oB;
    if (STAGE == 1) { oA = vA; oB = vB; }
    else if (STAGE == 2) { oA = 0.25 * (3.0 * y0v.x + vA); oB = 0.25 * (3.0 * y0v.y + vB); }
    else {
      oA = st.fin_a * (y0v.x + st.fin_b * vA);
      oB = st.fin_a * (y0v.y + st.fin_b * vB);
      if constexpr (SIMPLE) {
        oA = nan_min(oA, y0v.x); oB = nan_min(oB, y0v.y);
      } else {
        oA = comp_epilogue(oA, st.comp, y0v.x, auxv.x);
        oB = comp_epilogue(oB, st.comp, y0v.y, auxv.y);
        if (st.use_obs) { oA = nan_max(oA, -obsv.x); oB = nan_max(oB, -obsv.y); }
      }
    }
    if (ok1) *reinterpret_cast<double2*>(st.out + off) = make_double2(oA, oB);
    else if (ok0) st.out[off] = oA;
    // fused halo push: the same values go straight into the neighbours' halo planes over NVLink (posted stores: the
    // transfer for the next stage overlaps this kernel instead of following it)
    if (push_lo) {
      if (ok1) *reinterpret_cast<double2*>(st.push_lo + off) = make_double2(oA, oB);
      else if (ok0) st.push_lo[off] = oA;
    }
    if (push_hi) {
      if (ok1) *reinterpret_cast<double2*>(st.push_hi + off) = make_double2(oA, oB);
      else if (ok0) st.push_hi[off] = oA;
    }
    if (red && ((ok0 && oA != oA) || (ok1 && oB != oB))) acc.nan = 1;

    q[0] = q[1]; q[1] = q[2]; q[2] = ctr;
    ++z; ++kc; off += zstride;
    s_prev = s_cur; p_prev = p_cur;
    s_cur = s_p1; if (s_cur == 0) p_cur ^= 1;
    s_p1 = s_p2; s_p2 = s_new;
    if (++s_new == (unsigned)R) { s_new = 0; p_new ^= 1; }
  };

  const int zf_end = ((need_patch_a || need_patch_b) && !gw_on) ? z0 : z1 - R + 1;
  plane.template operator()<false>();
  const bool simple = st.restrict_sign == 0 && (STAGE != 3 || (st.comp == HJ_COMP_MIN_OVER_TIME && !st.use_obs));
  if (simple) {
    while (z < zf_end) plane.template operator()<true, true>();
  } else {
    while (z < zf_end) plane.template operator()<true>();
  }
  while (z < z1) plane.template operator()<false>();
  if (RED) acc.flush(st.red);
}

}  // namespace hjtma
