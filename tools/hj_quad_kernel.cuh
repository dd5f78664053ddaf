// hj_quad_kernel.cuh -- the plane-ring stage kernel with a 2 x 2 node block per thread (whole 3-D systems).
// DEVELOPER HARNESS ONLY (tools/tune_tma.cu): measured, bit-identical to the production kernel, and SLOWER -- see the
// results at the end of this header comment.  It is not compiled into the library.
//
// Same machinery as k_stage_tma (hj_tma_kernel.cuh): a CTA owns a TY x TX tile of (Y, X), marches a chunk of Z planes,
// every haloed plane box arrives by one TMA load into a ring of shared-memory slots (full[] / empty[] mbarriers, lane 0
// of warp 0 is the producer), planes z-3..z-1 of the Z stencil live in a register queue, ghost cells are made in
// registers.  What differs is the work per thread: a thread owns two X-adjacent nodes in each of two Y-adjacent rows.
// Per plane it reads two 10-column X windows (10 LDS.128), the six rows above / below its row pair (6: the two rows
// are each other's Y neighbours) and planes z+1..z+3 of both rows (6): 22 LDS.128 for 4 nodes = 11 doubles per node
// against 14 in the 1 x 2 kernel, and the ~100 non-FP64 instructions of a plane body (ring bookkeeping, barriers,
// addresses, the producer) are paid once per 4 nodes instead of once per 2.  The 1 x 2 kernel is bound by the
// shared-memory pipe and the issue slots next to the FP64 pipe; this one is made to lean on the FP64
// pipe alone.  It needs ~168 registers per thread: 192 threads (a 32 x 24 tile) x 2 CTAs per SM.
//
// Results are bit-identical to k_stage_tma: every node sees the same operands in the same order.
//
// Measured (B200, air3D 512^3, as_shipped, ms per launch for stages 1 / 2 / 3; profiles/r02_tune_quad.txt):
//   production 1 x 2 kernel, 32 x 16 tile, 2 x 256 threads, 120-128 registers          0.77 / 0.83 / 0.93   (2.53 per step)
//   2 x 2, 32 x 24 tile, 2 x 192 threads, 168 registers (stage 3 spills in the loop)   0.86 / 1.01 / 1.50   (64-plane chunks)
//   2 x 2, 32 x 32 tile, 1 x 256 threads, 218-237 registers, X windows prefetched      0.86 / 0.92 / 1.17
//   2 x 2, 32 x 16 tile, 3 x 128 threads, 168 registers                                0.85 / 1.00 / 1.49
// 21 % fewer shared-memory reads and 20 % fewer issued instructions per node (386 against 2 x 239 SASS instructions per
// plane body) do not pay for going from 16 to 12 (or 8) resident warps: the plane body is a chain of dependent
// LDS -> DADD/DFMA groups, and what hides their latencies is the number of warps, not the work per warp.
#pragma once
#include "../levelsetpy_b200/csrc/hj_tma_kernel.cuh"

// OPT bit 1: X windows of plane z+1 are loaded before the Z arithmetic of plane z (software pipeline across planes; costs
//            40 registers across the loop edge)
template <int R_, int MINB_, int TY_ = 24, int TXP_ = 16, int OPT_ = 0>
struct QuadCfg {
  static constexpr int OPT = OPT_;
  static constexpr int R = R_, MINB = MINB_, TXP = TXP_, TX = 2 * TXP_, TY = TY_;
  static_assert(TY_ % 2 == 0, "row pairs");
  static constexpr int NTHREADS = TXP * TY / 2;
  static_assert(NTHREADS % 32 == 0, "whole warps");
  static constexpr int BW = TX + 8, BH = TY + 6;
  static constexpr int BOX = BW * BH, YBOX = TX * TY;
  static constexpr int SLOT = (BOX + 15) / 16 * 16, YSLOT_FULL = (YBOX + 15) / 16 * 16;
  template <int STAGE>
  static constexpr size_t smem_bytes() {
    return (size_t)R * (SLOT + (STAGE >= 2 ? YSLOT_FULL : 0)) * 8 + 2 * R * 8;
  }
};

namespace hjtma {

template <class Sys, int WENO, bool RED, int STAGE, class Cfg>
__global__ void __launch_bounds__(Cfg::NTHREADS, Cfg::MINB)
k_stage_quad(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_y0, const KGrid g,
             const KSys ks, const KStage st, const TmaGeom geo) {
  static_assert(Sys::BASE_DIM == 0 && Sys::ND == 3 && Sys::NSCRATCH == 0, "whole 3-D systems");
  static_assert(STAGE >= 1 && STAGE <= 3, "RK stages");
  constexpr int D = 3, DX = 2, DY = 1, DZ = 0;
  constexpr int TX = Cfg::TX, TY = Cfg::TY, BW = Cfg::BW, PAIRS = Cfg::TXP, R = Cfg::R;
  constexpr int NWARPS = Cfg::NTHREADS / 32;
  constexpr int SLOT = Cfg::SLOT, YSLOT_FULL = Cfg::YSLOT_FULL;
  static_assert((SLOT * 8) % 128 == 0 && (YSLOT_FULL * 8) % 128 == 0, "slots must keep 128-byte alignment");
  static_assert(R >= 6 && R <= 16, "ring depth");
  constexpr int YSLOT = (STAGE >= 2) ? YSLOT_FULL : 0;
  constexpr int YBOX = (STAGE >= 2) ? Cfg::YBOX : 0;
  constexpr bool PIPE = (Cfg::OPT & 1) != 0;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);
  double* yring = ring + (size_t)R * SLOT;
  const uint32_t ring_s = smem_u32(smem_raw);
  const uint32_t yring_s = ring_s + R * SLOT * 8;
  const uint32_t full_s = ring_s + R * (SLOT + YSLOT) * 8;
  const uint32_t empty_s = full_s + R * 8;

  const int tid = threadIdx.x;
  long long b = blockIdx.x;
  const int xt = (int)(b % geo.nxt); b /= geo.nxt;
  const int yt = (int)(b % geo.nyt); b /= geo.nyt;
  const int zc = (int)(b % geo.nzc);
  const int NX = g.N[DX], NY = g.N[DY], NZ = g.N[DZ];
  const int x0 = xt * TX, y0 = yt * TY, z0 = geo.zbeg + zc * geo.cz;
  const int z1 = min(z0 + geo.cz, geo.zend);
  const int bcx = g.bc[DX], bcy = g.bc[DY], bcz = g.bc[DZ];
  const unsigned klast = (unsigned)((z1 - 1 + 3) - (z0 - 3));   // ring position of the last plane this chunk needs
  const int zcoord_base = (int)geo.zcoord0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < R; ++s) {
      mbar_init(full_s + 8 * s, 1);
      mbar_init(empty_s + 8 * s, NWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const double dt = st.dt;
  __syncthreads();

  // plane with ring position k -> slot s: TMA load, or a bare arrival for a computed ghost plane
  auto issue = [&](unsigned k, unsigned s) {
    const int zp = z0 - 3 + (int)k;
    int zsrc = zp;
    bool load = true;
    if (zp < 0 || zp >= NZ) {
      if (bcz == HJ_BC_PERIODIC) zsrc = zp < 0 ? zp + NZ : zp - NZ;
      else if (bcz == HJ_BC_EXTRAPOLATE) load = false;          // ghost plane: computed from the register queue
    }
    // the y0 tile rides along for planes that will be "current" (ring positions 3 .. klast-3)
    const bool ytile = STAGE >= 2 && k >= 3 && k + 3 <= klast;
    const uint32_t fb = full_s + 8 * s;
    if (load) {
      mbar_expect_tx(fb, (Cfg::BOX + (ytile ? YBOX : 0)) * 8);
      tma_load_3d(ring_s + s * (SLOT * 8), &tmap, fb, x0 - 4, y0 - 3, zcoord_base + zsrc);
      if (ytile) tma_load_3d(yring_s + s * (YSLOT * 8), &tmap_y0, fb, x0, y0, zcoord_base + zp);
    } else {
      mbar_arrive(fb);
    }
  };
  if (tid == 0) {
    for (unsigned k = 0; k < (unsigned)R && k <= klast; ++k) issue(k, k);
  }
  // does any node of this tile have an X / Y stencil that leaves the grid?  (CTA-uniform)
  const bool need_patch_x = x0 - 3 < 0 || x0 + TX + 2 >= NX;
  const bool need_patch_y = y0 - 3 < 0 || y0 + TY + 2 >= NY;
  const int lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);       // warp index, known warp-uniform to the compiler

  // my 2 x 2 block: columns ix, ix+1 of rows iy, iy+1
  const int tp = tid % PAIRS, tr = tid / PAIRS;
  const int ix = x0 + 2 * tp, iy = y0 + 2 * tr;
  bool ok[2][2];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 2; ++c) ok[r][c] = ix + c < NX && iy + r < NY;
  const long long ystride = g.stride[DY], zstride = g.stride[DZ];
  long long off = (long long)iy * ystride + ix + (long long)z0 * zstride;   // stride[DX] == 1

  // system state of my four nodes: everything that does not depend on the marching dim is loaded once
  typename Sys::Pt pt[2][2];
  {
    int idx[D];
    idx[DZ] = z0;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        idx[DY] = min(iy + r, NY - 1);                         // clamp: masked nodes must not read past the axis tables
        idx[DX] = min(ix + c, NX - 1);
        pt[r][c] = Sys::load(idx, g, ks, nullptr);
      }
  }
  double inv_eps[D];
#pragma unroll
  for (int d = 0; d < D; ++d) inv_eps[d] = (WENO == HJ_WENO_INTENDED) ? inv_eps_from_max(st.epsmax[d]) : 0.0;

  const int myoff = (2 * tr + 3) * BW + 4 + 2 * tp;          // my upper pair inside a slot (doubles); the lower one: + BW
  const bool full_tile = x0 + TX <= NX && y0 + TY <= NY;

  RedAcc<D> acc;
  acc.init();

  // ---- prologue: planes z0-3 .. z0-1 (ring positions 0..2) go into the Z register queues and their slots are handed
  // back; planes z0 .. z0+2 must have landed before the march starts (the march only waits for plane z+3)
  double2 q[2][3];                                           // planes z-3, z-2, z-1 of my two pairs
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    mbar_wait(full_s + 8 * k, 0);
    if (k < 3) {
      q[0][k] = lds2(ring + (size_t)k * SLOT + myoff);
      q[1][k] = lds2(ring + (size_t)k * SLOT + myoff + BW);
    }
  }
  if (bcz == HJ_BC_EXTRAPOLATE && z0 < 3) {
    // planes below the grid (ring positions k < 3 - z0; a plane range may start at z0 = 1 or 2) from planes 0, 1
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const double2 e0 = lds2(ring + (size_t)(3 - z0) * SLOT + myoff + r * BW);
      const double2 e1 = lds2(ring + (size_t)(4 - z0) * SLOT + myoff + r * BW);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (k < 3 - z0) {
          q[r][k].x = ghost_extrapolate(e0.x, e1.x, 3 - z0 - k, g.slope_mult[DZ]);
          q[r][k].y = ghost_extrapolate(e0.y, e1.y, 3 - z0 - k, g.slope_mult[DZ]);
        }
      }
    }
  }
  __syncwarp();
  if (lane == 0) { mbar_arrive(empty_s + 0); mbar_arrive(empty_s + 8); mbar_arrive(empty_s + 16); }
  if (tid == 0) {
    for (unsigned k = R; k < (unsigned)R + 3 && k <= klast; ++k) {     // the slots of z0-3..z0-1 get planes R..R+2
      mbar_wait(empty_s + 8 * (k - R), 0);
      issue(k, k - R);
    }
  }

  // ---- march
  unsigned kc = 3;                                           // ring position of the current plane
  unsigned s_prev = 2 % R, s_cur = 3 % R, s_p1 = 4 % R, s_p2 = 5 % R, s_new = 6 % R;
  unsigned p_prev = 0, p_cur = 0, p_new = (6 / R) & 1;
  int z = z0;
  double2 raw_next = Sys::template fetch<DZ>(z0, g, ks);
  double2 xw[2][5];                                          // PIPE: X windows of the next plane

  auto load_xw = [&](unsigned slot) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const double2* rowp = reinterpret_cast<const double2*>(ring + (size_t)slot * SLOT + myoff + r * BW);
#pragma unroll
      for (int j = 0; j < 5; ++j) xw[r][j] = rowp[j - 2];
    }
  };

  // one plane.  FAST: interior plane of an interior tile -- the plane to prefetch (z+R-1) exists and will be "current" in
  // this chunk, plane z+3 exists, no stencil leaves the grid in X/Y: no ghost code in the loop body.
  auto plane = [&]<bool FAST, bool SIMPLE = false>() {
    if constexpr (FAST) {
      if (warp_u == 0) {                                     // uniform branch: warps 1.. skip the divergent section
        if (lane == 0) {                                     // producer duty: recycle the slot of plane z-1
          mbar_wait(empty_s + 8 * s_prev, p_prev);
          const uint32_t fb = full_s + 8 * s_prev;
          mbar_expect_tx(fb, (Cfg::BOX + YBOX) * 8);
          tma_load_3d(ring_s + s_prev * (SLOT * 8), &tmap, fb, x0 - 4, y0 - 3, zcoord_base + z + R - 1);
          if (STAGE >= 2) tma_load_3d(yring_s + s_prev * (YSLOT * 8), &tmap_y0, fb, x0, y0, zcoord_base + z + R - 1);
        }
      }
    } else {
      if (tid == 0 && kc >= 4 && kc - 1 + R <= klast) {
        mbar_wait(empty_s + 8 * s_prev, p_prev);
        issue(kc - 1 + R, s_prev);
      }
    }
    // plane z+3 (the newest the stencil needs) has normally landed long ago: probe now, consume the answer later
    const uint32_t landed = mbar_test(full_s + 8 * s_new, p_new);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 2; ++c) Sys::template apply<DZ>(pt[r][c], raw_next, ks);   // the marching dim is shared
    raw_next = Sys::template fetch<DZ>(min(z + 1, NZ - 1), g, ks);

    // early global loads: aux / obstacle pairs
    double2 auxv[2], obsv[2], y0v[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) auxv[r] = obsv[r] = y0v[r] = make_double2(0.0, 0.0);
    if (STAGE == 3 && !SIMPLE) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (ok[r][0]) {
          if (st.comp == HJ_COMP_MIN_WITH_AUX || st.comp == HJ_COMP_MAX_WITH_AUX) auxv[r] = ldg2(st.aux + off + r * ystride);
          if (st.use_obs) obsv[r] = ldg2(st.obs + off + r * ystride);
        }
      }
    }

    double pc[2][2][D], hd[2][2][D];                         // [row][col][dim]
    double L, Rr;
    constexpr bool red = RED;
#define HJ_QRED(d, okk)                                            \
  if (red && (okk)) {                                              \
    acc.dmin[d] = fmin(acc.dmin[d], fmin(L, Rr));                  \
    acc.dmax[d] = fmax(acc.dmax[d], fmax(L, Rr));                  \
  }
#define HJ_QBARRIER asm volatile("" ::: "memory");
    const double* cur = ring + (size_t)s_cur * SLOT;
    // ---- X windows: columns ix-4 .. ix+5 of my two rows (w[r][2] = my pair of row r)
    double2 w[2][5];
    if constexpr (PIPE && FAST) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int j = 0; j < 5; ++j) w[r][j] = xw[r][j];      // loaded during the previous plane
    } else {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const double2* rowp = reinterpret_cast<const double2*>(cur + myoff + r * BW);
#pragma unroll
        for (int j = 0; j < 5; ++j) w[r][j] = rowp[j - 2];
      }
    }
    if constexpr (!FAST) {
      if (need_patch_x) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
          if (ok[r][0])
            patch_x(w[r][0], w[r][1], w[r][2], w[r][3], w[r][4], ix, x0, NX, bcx, g.slope_mult[DX],
                    cur + (2 * tr + r + 3) * BW, st.in + off + r * ystride - ix, NX <= TX);
      }
    }
    const double2 ctr[2] = {w[0][2], w[1][2]};
    // ---- Y rows iy-3, iy-2, iy-1 and iy+2, iy+3, iy+4 at my column pair (rows iy, iy+1 are the centres)
    double2 yr[6];
    yr[0] = lds2(cur + myoff - 3 * BW); yr[1] = lds2(cur + myoff - 2 * BW); yr[2] = lds2(cur + myoff - 1 * BW);
    yr[3] = lds2(cur + myoff + 2 * BW); yr[4] = lds2(cur + myoff + 3 * BW); yr[5] = lds2(cur + myoff + 4 * BW);
    HJ_QBARRIER
    // ---- X arithmetic: node (r, 0) uses columns ix-3..ix+3, node (r, 1) is shifted by one
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      pc_hd<WENO>(w[r][0].y, w[r][1].x, w[r][1].y, w[r][2].x, w[r][2].y, w[r][3].x, w[r][3].y, g, DX, inv_eps[DX],
                  pc[r][0][DX], hd[r][0][DX], L, Rr, red);
      HJ_QRED(DX, ok[r][0])
      pc_hd<WENO>(w[r][1].x, w[r][1].y, w[r][2].x, w[r][2].y, w[r][3].x, w[r][3].y, w[r][4].x, g, DX, inv_eps[DX],
                  pc[r][1][DX], hd[r][1][DX], L, Rr, red);
      HJ_QRED(DX, ok[r][1])
    }
    // ---- Z neighbours above: planes z+1, z+2 landed earlier; plane z+3 is the newest one of the ring
    double2 zp1[2], zp2[2], zp3[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (STAGE >= 2) y0v[r] = lds2(yring + (size_t)s_cur * YSLOT + (2 * tr + r) * TX + 2 * tp);
      zp1[r] = lds2(ring + (size_t)s_p1 * SLOT + myoff + r * BW);
      zp2[r] = lds2(ring + (size_t)s_p2 * SLOT + myoff + r * BW);
    }
    if (!landed) mbar_wait(full_s + 8 * s_new, p_new);
#pragma unroll
    for (int r = 0; r < 2; ++r) zp3[r] = lds2(ring + (size_t)s_new * SLOT + myoff + r * BW);
    HJ_QBARRIER
    // ---- Y arithmetic: row 0 sees (iy-3, iy-2, iy-1 | iy | iy+1, iy+2, iy+3), row 1 the same shifted by one row
    {
      double2 a[2][6];
      a[0][0] = yr[0]; a[0][1] = yr[1]; a[0][2] = yr[2]; a[0][3] = ctr[1]; a[0][4] = yr[3]; a[0][5] = yr[4];
      a[1][0] = yr[1]; a[1][1] = yr[2]; a[1][2] = ctr[0]; a[1][3] = yr[3]; a[1][4] = yr[4]; a[1][5] = yr[5];
      if constexpr (!FAST) {
        if (need_patch_y) {
#pragma unroll
          for (int r = 0; r < 2; ++r)
            if (ok[r][0])
              patch_y<BW>(a[r][0], a[r][1], a[r][2], a[r][3], a[r][4], a[r][5], iy + r, y0, NY, bcy, g.slope_mult[DY],
                          cur + 4 + 2 * tp, st.in + off - (long long)iy * ystride, ystride, NY <= TY);
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        pc_hd<WENO>(a[r][0].x, a[r][1].x, a[r][2].x, ctr[r].x, a[r][3].x, a[r][4].x, a[r][5].x, g, DY, inv_eps[DY],
                    pc[r][0][DY], hd[r][0][DY], L, Rr, red);
        HJ_QRED(DY, ok[r][0])
        pc_hd<WENO>(a[r][0].y, a[r][1].y, a[r][2].y, ctr[r].y, a[r][3].y, a[r][4].y, a[r][5].y, g, DY, inv_eps[DY],
                    pc[r][1][DY], hd[r][1][DY], L, Rr, red);
        HJ_QRED(DY, ok[r][1])
      }
    }
    if constexpr (!FAST) {
      if (bcz == HJ_BC_EXTRAPOLATE && z + 3 >= NZ) {         // ghost planes above the grid: edge plane NZ-1 = z+ke
        const int ke = NZ - 1 - z;                           // 0..2
        const double m = g.slope_mult[DZ];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const double2 ed = ke == 0 ? ctr[r] : (ke == 1 ? zp1[r] : zp2[r]);
          const double2 nx = ke == 0 ? q[r][2] : (ke == 1 ? ctr[r] : zp1[r]);
          if (ke < 1) zp1[r] = make_double2(ghost_extrapolate(ed.x, nx.x, 1 - ke, m), ghost_extrapolate(ed.y, nx.y, 1 - ke, m));
          if (ke < 2) zp2[r] = make_double2(ghost_extrapolate(ed.x, nx.x, 2 - ke, m), ghost_extrapolate(ed.y, nx.y, 2 - ke, m));
          zp3[r] = make_double2(ghost_extrapolate(ed.x, nx.x, 3 - ke, m), ghost_extrapolate(ed.y, nx.y, 3 - ke, m));
        }
      }
    }
    // this warp is done with the current plane's slot
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_s + 8 * s_cur);
    if constexpr (PIPE && FAST) {                            // X windows of plane z+1 (resident since two planes ago)
      load_xw(s_p1);
      HJ_QBARRIER
    }
    // ---- Z arithmetic
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      pc_hd<WENO>(q[r][0].x, q[r][1].x, q[r][2].x, ctr[r].x, zp1[r].x, zp2[r].x, zp3[r].x, g, DZ, inv_eps[DZ],
                  pc[r][0][DZ], hd[r][0][DZ], L, Rr, red);
      HJ_QRED(DZ, ok[r][0])
      pc_hd<WENO>(q[r][0].y, q[r][1].y, q[r][2].y, ctr[r].y, zp1[r].y, zp2[r].y, zp3[r].y, g, DZ, inv_eps[DZ],
                  pc[r][1][DZ], hd[r][1][DZ], L, Rr, red);
      HJ_QRED(DZ, ok[r][1])
    }
#undef HJ_QRED
#undef HJ_QBARRIER

    // ---- Hamiltonian + GLF dissipation, RK stage algebra, driver epilogue: node by node, as k_stage_tma does
    double o[2][2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        double yd = -Sys::ham(pt[r][c], pc[r][c], ks);       // ydot = -(ham - diss)
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const double al = Sys::alpha(d, pt[r][c], ks);
          yd = fma(hd[r][c][d], al, yd);                     // artificial_diss_glf.py:100
          if (red && ok[r][c]) acc.amax[d] = fmax(acc.amax[d], al);
        }
        if constexpr (!SIMPLE) yd = restrict_update(yd, st.restrict_sign);
        const double cv = c == 0 ? ctr[r].x : ctr[r].y;
        const double y0c = c == 0 ? y0v[r].x : y0v[r].y;
        double ov;
        if (STAGE == 1) ov = cv + dt * yd;
        else if (STAGE == 2) ov = 0.25 * (3.0 * y0c + (cv + dt * yd));
        else {
          ov = st.fin_a * (y0c + st.fin_b * (cv + dt * yd));
          if constexpr (SIMPLE) {
            ov = nan_min(ov, y0c);
          } else {
            const bool with_aux = st.comp == HJ_COMP_MIN_WITH_AUX || st.comp == HJ_COMP_MAX_WITH_AUX;
            ov = comp_epilogue(ov, st.comp, y0c, with_aux ? (c == 0 ? auxv[r].x : auxv[r].y) : 0.0);
            if (st.use_obs) ov = nan_max(ov, -(c == 0 ? obsv[r].x : obsv[r].y));
          }
        }
        o[r][c] = ov;
        if (red && ok[r][c] && ov != ov) acc.nan = 1;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      double* op = st.out + off + r * ystride;
      if (FAST && full_tile) *reinterpret_cast<double2*>(op) = make_double2(o[r][0], o[r][1]);
      else if (ok[r][1]) *reinterpret_cast<double2*>(op) = make_double2(o[r][0], o[r][1]);
      else if (ok[r][0]) op[0] = o[r][0];
    }

#pragma unroll
    for (int r = 0; r < 2; ++r) { q[r][0] = q[r][1]; q[r][1] = q[r][2]; q[r][2] = ctr[r]; }
    ++z; ++kc; off += zstride;
    s_prev = s_cur; p_prev = p_cur;
    s_cur = s_p1; if (s_cur == 0) p_cur ^= 1;
    s_p1 = s_p2; s_p2 = s_new;
    if (++s_new == (unsigned)R) { s_new = 0; p_new ^= 1; }
  };

  // head (general body) -> fast segment z in [z0+1, z1-R] (the prefetched plane z+R-1 <= z1-1 <= NZ-1) -> tail
  const int zf_end = (need_patch_x || need_patch_y) ? z0 : z1 - R + 1;
  plane.template operator()<false>();
  if (PIPE && z < zf_end) load_xw(s_cur);                    // pipeline prologue: X windows of the first fast plane
  // SIMPLE: the fast march specialised for the common epilogue (no termRestrictUpdate; stage 3 = minVOverTime without
  // obstacle), so that the steady-state loop carries no epilogue dispatch
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
