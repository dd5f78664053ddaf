// hj_tma_kernel.cuh -- the fused RHS + RK-stage kernel as a streamed plane ring (TMA backend), device side.
//
// Layout of the work (grid dims named X = D-1 (contiguous), Y = D-2, Z = D-3, "slow" = 0..D-4):
//   * a CTA owns a TY x TX tile of (Y, X) and marches a chunk of CZ planes along Z;
//   * each Z-plane of the tile, with its 3-cell X/Y halo (box (TY+6) x (TX+8) doubles; 4 columns on each side in X
//     keep every row and every thread's 2-node pair 16-byte aligned), is brought into a ring of R shared-memory
//     slots by ONE cp.async.bulk.tensor (TMA) per plane, signalled on an mbarrier; out-of-range box parts are
//     zero-filled by the TMA unit and the threads that own a boundary node replace them IN REGISTERS with the
//     extrapolated / periodic ghost cells (add_ghost_extrapolate.py:88-110, add_ghost_periodic.py:78-87) -- no
//     padded copy of the field ever exists, the ring is never written by a thread, and no CTA-wide barrier runs
//     in the steady state;
//   * every thread owns two X-adjacent nodes: planes z-3..z-1 of the Z stencil live in a 3-deep register queue,
//     planes z+1..z+3 are read from their ring slots (they are resident anyway), the X and Y stencils are read
//     from the current plane's slot, all with 16-byte LDS; slow-dim neighbours (D >= 4) come straight from L2/HBM
//     with 16-byte read-only loads;
//   * the march is split into a head, a FAST segment (interior planes of interior tiles: no ghost code in the loop
//     body) and a tail, so that the steady-state loop is compact;
//   * the Hamiltonian, GLF dissipation, RK stage algebra and the driver epilogue are applied in registers and
//     the result leaves with one 16-byte store per thread.  DRAM traffic per node: 8 B read + 8 B write
//     (+8 B for y0 in stages 2/3) = the algorithmic 16/24/24 B.
//
// Reference behaviour restated: see hj_common.cuh header.
#pragma once
#include <cuda.h>

#include "hj_common.cuh"
#include "hj_systems.cuh"

struct TmaGeom {
  int nxt, nyt, nzc, cz;      // tiles in X, Y; Z chunks; planes per chunk
  long long nslow;            // product of slow dims
  long long zcoord0;          // TMA dim-2 coordinate of (slow = 0, z = 0)  (halo planes on dim 0 shift it)
  int NZ;
  int zbeg, zend;             // planes [zbeg, zend) of Z are advanced by this launch (whole dim: 0, NZ)
};

// tuning knobs of the plane-ring kernel
// OPT_ bits (all on in production; the tuning harness switches them off one at a time):
//   1  probe the barrier of plane z+3 (mbarrier.test_wait) at the top of the plane body, spin only if it failed
//   2  producer duty behind a warp-uniform branch (only warp 0 sees the divergent lane-0 section)
//   4  fast march: unconditional 16-byte store (interior tiles have no masked node)
//   8  fast march software-pipelined across dims (see PIPE in the plane body; +2 % on stage 3)
// 128  fast march specialised for the common epilogue (SIMPLE)
//  16/32/64  tuning harness only: no arithmetic / every load hits L2 / no store (bound-finding experiments)
//
// GW_ (ghost warp): one extra warp per CTA writes the extrapolated / periodic ghost cells of every landed plane INTO its
// ring slot (the columns / rows the TMA unit zero-filled), so that tiles touching the domain boundary march through the
// same ghost-free FAST body as interior tiles.  Made for planes that are one or two tiles wide (41 x 41, 101 x 101: the
// trailing block of the 6-D pair, the Flock batch), where EVERY tile touches the boundary.  Consumers then wait on a
// second barrier per slot (ready[s], armed by the ghost warp) instead of the TMA barrier.
//
// XPAD_: extra columns loaded at the right end of every slot row.  A row of the slot is TX + 8 + XPAD doubles; thread
// pair t of tile row r reads the 16-byte bank group (t + r (4 + XPAD / 2) + 2) mod 8, so a quarter-warp that straddles
// two tile rows (TXP not a multiple of 8: the 21-pair rows of 41-wide planes) hits 2-way bank conflicts unless
// XPAD = 8 makes the row-to-row shift a whole 128 bytes.
template <int R_, int MINB_, int UNROLL_, int TY_ = 16, int TXP_ = 16, bool SEQ_ = false, int OPT_ = 143, int GW_ = 0,
          int XPAD_ = 0>
struct TmaCfg {
  static constexpr int OPT = OPT_;
  static constexpr int NGW = GW_;          // ghost warps (0 = none; 2: they take alternate planes)
  static constexpr bool GW = GW_ > 0;
  static constexpr bool SEQ = SEQ_;       // evaluate the stencil one dim at a time (smaller live set)
  static constexpr int R = R_;            // ring slots (planes z+1..z+3 are needed, the rest is prefetch distance)
  static constexpr int MINB = MINB_;      // resident CTAs per SM the register allocation is sized for
  static constexpr int UNROLL = UNROLL_;  // planes per trip of the fast march loop
  static constexpr int TXP = TXP_;        // node pairs per tile row: one node pair per thread
  static constexpr int TX = 2 * TXP_, TY = TY_;   // tile (X, Y)
  static constexpr int NACTIVE = TXP * TY;
  static constexpr int NCONS = (NACTIVE + 31) / 32 * 32;           // consumer threads: whole warps; the surplus idles
  static constexpr int NTHREADS = NCONS + 32 * GW_;                // + the ghost warp(s)
  static constexpr int BW = TX + 8 + XPAD_, BH = TY + 6;           // haloed plane box (doubles)
  static constexpr int BOX = BW * BH, YBOX = TX * TY;              // doubles the TMA unit writes per plane
  static constexpr int SLOT = (BOX + 15) / 16 * 16;                // slot strides keep 128-byte alignment
  static constexpr int YSLOT_FULL = (YBOX + 15) / 16 * 16;
  template <int STAGE>
  static constexpr size_t smem_bytes() {
    return (size_t)R * (SLOT + (STAGE >= 2 ? YSLOT_FULL : 0)) * 8 + 3 * R * 8;
  }
};

namespace hjtma {

// ------------------------------------------------------------------------------------------ PTX helpers
HJ_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
HJ_DEV void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
HJ_DEV void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
HJ_DEV void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
HJ_DEV void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
HJ_DEV uint32_t mbar_test(uint32_t bar, uint32_t parity) {      // non-blocking probe
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
HJ_DEV void tma_load_3d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"((uint64_t)tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
HJ_DEV double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
HJ_DEV double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// ------------------------------------------------------------------------------------------ stencil math on pairs
// derivC and 0.5*(derivR - derivL) of the as-shipped (fixed-weight) scheme straight from the 7 nodes; the
// coefficients are host-precomputed (KGrid::ca1..cb) and used as constant-bank operands:
//   derivC    = ca1 (v4-v2) + ca2 (v5-v1) + ca3 (v6-v0)                       (= 0.5*(L+R), 6 flops)
//   0.5*(R-L) = cb (v0+v6 - 6 (v1+v5) + 15 (v2+v4) - 20 v3)                   (sixth difference, 7 flops)
// i.e. 13 fp64 instructions per node per dim instead of ~60 for the divided-difference tables + weightWENO.
template <int WENO, int DBG = 0>
HJ_DEV void pc_hd(const double v0, const double v1, const double v2, const double v3, const double v4, const double v5,
                  const double v6, const KGrid& g, const int d, double inv_eps, double& pc, double& hd, double& L,
                  double& Rr, const bool need_lr) {
  if (DBG) {                      // tuning harness only: keep the loads alive, drop the arithmetic
    pc = v0 + v6; hd = v3 + v1; L = Rr = 0.0;
    asm volatile("" :: "d"(v2), "d"(v4), "d"(v5));
  } else if (WENO == HJ_WENO_AS_SHIPPED) {
    pc = g.ca1[d] * (v4 - v2) + g.ca2[d] * (v5 - v1) + g.ca3[d] * (v6 - v0);
    double t = fma(-6.0, v1 + v5, v0 + v6);
    t = fma(15.0, v2 + v4, t);
    t = fma(-20.0, v3, t);
    hd = g.cb[d] * t;
    if (need_lr) { L = pc - hd; Rr = pc + hd; }
  } else {
    // true WENO5, or the ENO3a / ENO2 candidates picked by minimum modulus (upwind_first_eno3a.py:104-140,
    // upwind_first_eno2.py:137-148) -- the same device functions the gather backend runs, so the divided-difference
    // tables and the choices made from them are bit-identical on both backends
    const double v[7] = {v0, v1, v2, v3, v4, v5, v6};
    upwind5<WENO>(v, g.dxinv[d], inv_eps, L, Rr, g.dx[d]);
    pc = 0.5 * (L + Rr);
    hd = 0.5 * (Rr - L);
  }
}

// one slow-dim neighbour pair (k = -3..3, k != 0) with on-the-fly boundary handling; CTA-uniform branches
HJ_DEV double2 slow_neighbor(const double* p, int i, int k, int n, long long s, int bc, double m) {
  const int j = i + k;
  if ((j >= 0 && j < n) || bc == HJ_BC_HALO) return ldg2(p + (long long)k * s);
  if (bc == HJ_BC_PERIODIC) return ldg2(p + (long long)((j < 0 ? j + n : j - n) - i) * s);
  const int e = j < 0 ? 0 : n - 1, nx = j < 0 ? 1 : n - 2, dist = j < 0 ? -j : j - (n - 1);
  const double2 a = ldg2(p + (long long)(e - i) * s), b = ldg2(p + (long long)(nx - i) * s);
  return make_double2(ghost_extrapolate(a.x, b.x, dist, m), ghost_extrapolate(a.y, b.y, dist, m));
}

// ------------------------------------------------------------------------------------------ ghost cells in registers
// X window of a thread: columns ix-4 .. ix+5 in (w0.x w0.y w1.x w1.y w2.x w2.y w3.x w3.y w4.x w4.y); the stencils of
// its two nodes use columns ix-3 .. ix+4.  Columns outside [0, NX) were zero-filled by the TMA unit (or belong to a
// neighbouring tile of a periodic dim) and are replaced here.  `srow` = my row inside the current slot (column
// x0-4), `grow` = column 0 of my row in global memory.  Only called by threads with a node inside the grid.
HJ_DEV void patch_x(double2& w0, double2& w1, double2& w2, double2& w3, double2& w4, const int ix, const int x0,
                    const int NX, const int bc, const double m, const double* srow, const double* grow,
                    const bool whole) {
  // `whole`: the tile spans the whole row (x0 == 0, NX <= TX), so periodic images are in the slot too
  if (ix < 3) {                                               // columns ix-3 .. ix-1 may be < 0
    double e0 = 0.0, e1 = 0.0;
    if (bc != HJ_BC_PERIODIC) { e0 = srow[0 - x0 + 4]; e1 = srow[1 - x0 + 4]; }
    auto gl = [&](int c) {
      if (bc == HJ_BC_PERIODIC) return whole ? srow[c + NX + 4] : __ldg(grow + c + NX);
      return ghost_extrapolate(e0, e1, -c, m);
    };
    if (ix - 3 < 0) w0.y = gl(ix - 3);
    if (ix - 2 < 0) w1.x = gl(ix - 2);
    if (ix - 1 < 0) w1.y = gl(ix - 1);
  }
  if (ix + 4 >= NX) {                                         // columns ix+1 .. ix+4 may be >= NX
    double f0 = 0.0, f1 = 0.0;
    if (bc != HJ_BC_PERIODIC) { f0 = srow[NX - 1 - x0 + 4]; f1 = srow[NX - 2 - x0 + 4]; }
    auto gr = [&](int c) {
      if (bc == HJ_BC_PERIODIC) return whole ? srow[c - NX + 4] : __ldg(grow + c - NX);
      return ghost_extrapolate(f0, f1, c - (NX - 1), m);
    };
    if (ix + 1 >= NX) w2.y = gr(ix + 1);
    if (ix + 2 >= NX) w3.x = gr(ix + 2);
    if (ix + 3 >= NX) w3.y = gr(ix + 3);
    if (ix + 4 >= NX) w4.x = gr(ix + 4);
  }
}

// Y neighbours of a thread's pair: rows iy-3 .. iy+3.  `scol` = row y0-3 of the current slot at my column pair,
// `gcol` = row 0 at my column pair in global memory, `ys` = global row stride.
template <int BW>
HJ_DEV void patch_y(double2& ym3, double2& ym2, double2& ym1, double2& yp1, double2& yp2, double2& yp3, const int iy,
                    const int y0, const int NY, const int bc, const double m, const double* scol, const double* gcol,
                    const long long ys, const bool whole) {
  // `whole`: the tile spans the whole dim (y0 == 0, NY <= tile), so periodic images are in the slot too
  if (iy < 3) {
    double2 e0 = make_double2(0.0, 0.0), e1 = e0;
    if (bc != HJ_BC_PERIODIC) { e0 = lds2(scol + (0 - y0 + 3) * BW); e1 = lds2(scol + (1 - y0 + 3) * BW); }
    auto gt = [&](int r) {
      if (bc == HJ_BC_PERIODIC) return whole ? lds2(scol + (r + NY + 3) * BW) : ldg2(gcol + (long long)(r + NY) * ys);
      return make_double2(ghost_extrapolate(e0.x, e1.x, -r, m), ghost_extrapolate(e0.y, e1.y, -r, m));
    };
    if (iy - 3 < 0) ym3 = gt(iy - 3);
    if (iy - 2 < 0) ym2 = gt(iy - 2);
    if (iy - 1 < 0) ym1 = gt(iy - 1);
  }
  if (iy + 3 >= NY) {
    double2 f0 = make_double2(0.0, 0.0), f1 = f0;
    if (bc != HJ_BC_PERIODIC) { f0 = lds2(scol + (NY - 1 - y0 + 3) * BW); f1 = lds2(scol + (NY - 2 - y0 + 3) * BW); }
    auto gb = [&](int r) {
      if (bc == HJ_BC_PERIODIC) return whole ? lds2(scol + (r - NY + 3) * BW) : ldg2(gcol + (long long)(r - NY) * ys);
      const int dist = r - (NY - 1);
      return make_double2(ghost_extrapolate(f0.x, f1.x, dist, m), ghost_extrapolate(f0.y, f1.y, dist, m));
    };
    if (iy + 1 >= NY) yp1 = gb(iy + 1);
    if (iy + 2 >= NY) yp2 = gb(iy + 2);
    if (iy + 3 >= NY) yp3 = gb(iy + 3);
  }
}

// ------------------------------------------------------------------------------------------ the kernel
// Producer/consumer ring without a CTA-wide barrier:
//   full[s]  : armed by the producer thread (expect_tx), completed by the TMA unit when plane box s has landed
//   empty[s] : one arrival per warp when that warp no longer needs the plane in slot s
// All 8 warps are consumers (one node pair per thread); lane 0 of warp 0 doubles as the producer: at the top of
// step z it waits until every warp has released plane z-1 (normally already true: the ring runs R-4 planes
// ahead of need) and re-arms that slot with plane z+R-1.
//
// `Sys` is the functor of the dim block [BASE_DIM, BASE_DIM + ND) of a GD-dimensional grid and the kernel
// differentiates along those dims only.  A whole system has BASE = 0, ND = GD.  The trailing block of a product
// system (BASE + ND == GD, dimension-split path, see hj_vec_kernel.cuh) makes this kernel the first of two passes:
// out = F_B(in) with STAGE = 0.
template <class Sys, int GD, int WENO, bool RED, int STAGE, class Cfg>
__global__ void __launch_bounds__(Cfg::NTHREADS, Cfg::MINB)
k_stage_tma(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_y0, const KGrid g,
            const KSys ks, const KStage st, const TmaGeom geo) {
  constexpr int D = GD;
  constexpr int B0 = Sys::BASE_DIM;                              // first differentiated dim
  static_assert(B0 + Sys::ND == GD, "the plane-ring kernel takes the trailing dim block");
  static_assert(D >= 3, "the plane-ring kernel needs a Z dim");
  constexpr int DX = D - 1, DY = D - 2, DZ = D - 3, NSLOW = D - 3;
  static_assert(DY >= B0, "X and Y must belong to the block");
  constexpr bool ZIN = DZ >= B0;                             // is the marching dim differentiated?
  constexpr int TX = Cfg::TX, BW = Cfg::BW, PAIRS = Cfg::TXP;
  constexpr int R = Cfg::R;
  constexpr int TY = Cfg::TY, NTHREADS = Cfg::NTHREADS, NCONS_WARPS = Cfg::NCONS / 32;
  constexpr int SLOT = Cfg::SLOT, YSLOT_FULL = Cfg::YSLOT_FULL;
  static_assert((SLOT * 8) % 128 == 0 && (YSLOT_FULL * 8) % 128 == 0, "slots must keep 128-byte alignment");
  static_assert(R >= 6 && R <= 16, "ring depth");
  // stages 2/3 also stream the un-haloed y0 tile (TY x TX) of each plane through the ring, on the same barrier
  constexpr int YSLOT = (STAGE >= 2) ? YSLOT_FULL : 0;
  constexpr int YBOX = (STAGE >= 2) ? Cfg::YBOX : 0;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);
  double* yring = ring + (size_t)R * SLOT;
  const uint32_t ring_s = smem_u32(smem_raw);
  const uint32_t yring_s = ring_s + R * SLOT * 8;
  const uint32_t full_s = ring_s + R * (SLOT + YSLOT) * 8;
  const uint32_t empty_s = full_s + R * 8;
  const uint32_t ready_s = empty_s + R * 8;                      // ghost-warp configurations only

  const int tid = threadIdx.x;
  long long b = blockIdx.x;
  const int xt = (int)(b % geo.nxt); b /= geo.nxt;
  const int yt = (int)(b % geo.nyt); b /= geo.nyt;
  const int zc = (int)(b % geo.nzc); b /= geo.nzc;
  const long long slow_flat = b;
  const int NX = g.N[DX], NY = g.N[DY], NZ = g.N[DZ];
  const int x0 = xt * TX, y0 = yt * TY, z0 = geo.zbeg + zc * geo.cz;
  const int z1 = min(z0 + geo.cz, geo.zend);
  const int bcx = g.bc[DX], bcy = g.bc[DY], bcz = g.bc[DZ];
  const unsigned klast = (unsigned)((z1 - 1 + 3) - (z0 - 3));   // ring position of the last plane this chunk needs
  const int zcoord_base = (int)(geo.zcoord0 + slow_flat * NZ);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < R; ++s) {
      mbar_init(full_s + 8 * s, 1);
      mbar_init(empty_s + 8 * s, NCONS_WARPS);
      mbar_init(ready_s + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // batch functors keep the parameter block of this CTA's batch element (= the slow index) in shared memory
  double* scratch = reinterpret_cast<double*>(smem_raw + Cfg::template smem_bytes<STAGE>());
  if constexpr (Sys::NSCRATCH > 0) {
    static_assert(NSLOW == 1, "batch functors: the only slow dim is the batch index");
    Sys::fill_scratch(scratch, slow_flat, ks, tid, NTHREADS);
  }
  const double dt = st.dt_arr ? __ldg(st.dt_arr + slow_flat) : st.dt;
  __syncthreads();
  // batch contexts: an element whose dt is 0 has reached its t_end -- the per-grid loop of the reference would not step
  // it any more (ode_cfl_3.py:125), so its field stays bit for bit as it is (CTA-uniform exit; buffer 0 is untouched
  // because all three stages skip)
  if (st.dt_arr && dt == 0.0) return;

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

  // ================================================================== ghost warp
  // Plane by plane, behind the TMA unit: write the ghost cells of the landed plane into its slot, then release it to the
  // consumers (ready[s]).  Only planes that will be "current" need them (the Z stencil reads interior nodes only).
  // The slot was written through the async proxy and will be again: fence.proxy.async orders this warp's generic-proxy
  // stores before the consumers' release of the slot lets the producer re-arm it.
  const bool gw_on = Cfg::GW && (need_patch_x || need_patch_y);
  if constexpr (Cfg::GW) {
    if (tid >= Cfg::NCONS) {
      if (!gw_on) return;
      long long base = 0;
      {
        long long r = slow_flat;
#pragma unroll
        for (int d = NSLOW - 1; d >= 0; --d) { base += (long long)(r % g.N[d]) * g.stride[d]; r /= g.N[d]; }
      }
      const long long ystride = g.stride[DY];
      const double mx = g.slope_mult[DX], my = g.slope_mult[DY];
      const bool whole_x = NX <= TX, whole_y = NY <= TY;
      const int nrow = min(TY, NY - y0);                       // tile rows inside the grid
      for (unsigned k = (unsigned)(tid - Cfg::NCONS) / 32; k <= klast; k += Cfg::NGW) {
        const unsigned s = k % R;
        mbar_wait(full_s + 8 * s, (k / R) & 1);
        const int zp = z0 - 3 + (int)k;
        int zsrc = zp;
        bool loaded = true;
        if (zp < 0 || zp >= NZ) {
          if (bcz == HJ_BC_PERIODIC) zsrc = zp < 0 ? zp + NZ : zp - NZ;
          else if (bcz == HJ_BC_EXTRAPOLATE) loaded = false;
        }
        if ((Cfg::OPT & 256) == 0 && loaded && k >= 3 && k + 3 <= klast) {     // 256: tuning harness only, no fill
          double* sl = ring + (size_t)s * SLOT;
          const double* gplane = st.in + base + (long long)zsrc * g.stride[DZ];
          // One lane per row (X sides) / per column pair (Y sides): the two edge values are read once and give all the
          // ghost cells of that side (the slope of add_ghost_extrapolate.py:88-100 is common to them).
          if (need_patch_x) {
            for (int row = lane; row < nrow; row += 32) {
              double* rowp = sl + (row + 3) * BW + 4 - x0;          // rowp[c] = column c of the grid
              const double* grow = gplane + (long long)(y0 + row) * ystride;
              if (x0 == 0) {                                        // columns -3..-1
                if (bcx == HJ_BC_PERIODIC) {
#pragma unroll
                  for (int c = -3; c < 0; ++c) rowp[c] = whole_x ? rowp[c + NX] : __ldg(grow + c + NX);
                } else {
                  const double e0 = rowp[0], e1 = rowp[1];
#pragma unroll
                  for (int c = -3; c < 0; ++c) rowp[c] = ghost_extrapolate(e0, e1, -c, mx);
                }
              }
              if (x0 + TX + 2 >= NX) {                              // columns NX..NX+3 (as far as the box reaches)
                const int cmax = min(NX + 3, x0 + TX + 3);
                if (bcx == HJ_BC_PERIODIC) {
                  for (int c = NX; c <= cmax; ++c) rowp[c] = whole_x ? rowp[c - NX] : __ldg(grow + c - NX);
                } else {
                  const double f0 = rowp[NX - 1], f1 = rowp[NX - 2];
                  for (int c = NX; c <= cmax; ++c) rowp[c] = ghost_extrapolate(f0, f1, c - (NX - 1), mx);
                }
              }
            }
          }
          if (need_patch_y) {
            for (int cp = lane; cp < TX / 2; cp += 32) {
              if (x0 + 2 * cp >= NX) continue;
              double* colp = sl + 4 + 2 * cp + (3 - y0) * BW;       // colp[r * BW] = row r of the grid, my column pair
              const double* gcol = gplane + x0 + 2 * cp;
              if (y0 == 0) {                                        // rows -3..-1
                if (bcy == HJ_BC_PERIODIC) {
#pragma unroll
                  for (int r = -3; r < 0; ++r)
                    *reinterpret_cast<double2*>(colp + r * BW) = whole_y ? lds2(colp + (r + NY) * BW) : ldg2(gcol + (long long)(r + NY) * ystride);
                } else {
                  const double2 e0 = lds2(colp), e1 = lds2(colp + BW);
#pragma unroll
                  for (int r = -3; r < 0; ++r)
                    *reinterpret_cast<double2*>(colp + r * BW) =
                        make_double2(ghost_extrapolate(e0.x, e1.x, -r, my), ghost_extrapolate(e0.y, e1.y, -r, my));
                }
              }
              if (y0 + TY + 2 >= NY) {                              // rows NY..NY+2 (as far as the box reaches)
                const int rmax = min(NY + 2, y0 + TY + 2);
                if (bcy == HJ_BC_PERIODIC) {
                  for (int r = NY; r <= rmax; ++r)
                    *reinterpret_cast<double2*>(colp + r * BW) = whole_y ? lds2(colp + (r - NY) * BW) : ldg2(gcol + (long long)(r - NY) * ystride);
                } else {
                  const double2 f0 = lds2(colp + (NY - 1) * BW), f1 = lds2(colp + (NY - 2) * BW);
                  for (int r = NY; r <= rmax; ++r)
                    *reinterpret_cast<double2*>(colp + r * BW) =
                        make_double2(ghost_extrapolate(f0.x, f1.x, r - (NY - 1), my), ghost_extrapolate(f0.y, f1.y, r - (NY - 1), my));
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
  // consumers see a plane when it has landed (and, with a working ghost warp, when its ghost cells are in place)
  const uint32_t land_s = gw_on ? ready_s : full_s;

  // ================================================================== consumers
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);       // warp index, known warp-uniform to the compiler
  const bool live = tid < Cfg::NACTIVE;                     // surplus threads of the last warp only keep the barriers
  const int tp = tid % PAIRS, ty = live ? tid / PAIRS : TY - 1;
  int idx[D];
  {
    long long r = slow_flat;
#pragma unroll
    for (int d = NSLOW - 1; d >= 0; --d) { idx[d] = (int)(r % g.N[d]); r /= g.N[d]; }
  }
  const int ix = x0 + 2 * tp, iy = y0 + ty;
  const bool ok0 = live && ix < NX && iy < NY, ok1 = live && ix + 1 < NX && iy < NY;
  long long off = (long long)iy * g.stride[DY] + ix + (long long)z0 * g.stride[DZ];   // stride[DX] == 1
#pragma unroll
  for (int d = 0; d < NSLOW; ++d) off += (long long)idx[d] * g.stride[d];
  const long long zstride = g.stride[DZ];

  double inv_eps[D];
#pragma unroll
  for (int d = 0; d < D; ++d) inv_eps[d] = (WENO == HJ_WENO_INTENDED) ? inv_eps_from_max(st.epsmax[d]) : 0.0;
  // system state of my two nodes: everything that does not depend on the marching dim is loaded once
  idx[DZ] = z0;
  idx[DY] = min(iy, NY - 1);                                 // clamp: masked threads must not read past the axis tables
  idx[DX] = min(ix, NX - 1);
  typename Sys::Pt ptA = Sys::load(idx, g, ks, scratch);
  idx[DX] = min(ix + 1, NX - 1);
  typename Sys::Pt ptB = Sys::load(idx, g, ks, scratch);

  const int myoff = (ty + 3) * BW + 4 + 2 * tp;              // my pair inside a slot (doubles); 16-byte aligned
  // a tile with nodes outside the grid keeps its stores masked in the fast march too
  const bool full_tile = x0 + TX <= NX && y0 + TY <= NY;

  RedAcc<D> acc;
  acc.init();

  // ---- prologue: planes z0-3 .. z0-1 (ring positions 0..2) go into the Z register queue and their slots are handed
  // back; planes z0 .. z0+2 must have landed before the march starts (the march only waits for plane z+3)
  double2 q[3];                                              // planes z-3, z-2, z-1 of my pair
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    mbar_wait(land_s + 8 * k, 0);
    if (k < 3) q[k] = lds2(ring + (size_t)k * SLOT + myoff);
  }
  if (bcz == HJ_BC_EXTRAPOLATE && z0 < 3) {
    // planes below the grid (ring positions k < 3 - z0; a plane range may start at z0 = 1 or 2) from planes 0, 1
    const double2 e0 = lds2(ring + (size_t)(3 - z0) * SLOT + myoff), e1 = lds2(ring + (size_t)(4 - z0) * SLOT + myoff);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k < 3 - z0) {
        q[k].x = ghost_extrapolate(e0.x, e1.x, 3 - z0 - k, g.slope_mult[DZ]);
        q[k].y = ghost_extrapolate(e0.y, e1.y, 3 - z0 - k, g.slope_mult[DZ]);
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
  // slot of ring position kc + j (j = -1..3) and the phase parity of positions kc-1 / kc+3, maintained incrementally
  unsigned s_prev = 2 % R, s_cur = 3 % R, s_p1 = 4 % R, s_p2 = 5 % R, s_new = 6 % R;
  unsigned p_prev = 0, p_cur = 0, p_new = (6 / R) & 1;
  int z = z0;
  double2 raw_next = Sys::template fetch<DZ>(z0, g, ks);
  double2 xw[5];                                             // pipelined fast march: X window of the next plane

  // one plane.  FAST: interior plane of an interior tile -- the plane to prefetch (z+R-1) exists and will be
  // "current" in this chunk, plane z+3 exists, no stencil leaves the grid in X/Y: no ghost code in the loop body.
  auto plane = [&]<bool FAST, bool SIMPLE = false>() {
    if constexpr (FAST) {
      auto produce = [&]() {                                 // producer duty: recycle the slot of plane z-1
        mbar_wait(empty_s + 8 * s_prev, p_prev);
        const uint32_t fb = full_s + 8 * s_prev;
        mbar_expect_tx(fb, (Cfg::BOX + YBOX) * 8);
        if constexpr ((Cfg::OPT & 32) != 0) {               // tuning harness only: every load hits L2
          tma_load_3d(ring_s + s_prev * (SLOT * 8), &tmap, fb, 28, 13, (z & 7) + 8);
          if (STAGE >= 2) tma_load_3d(yring_s + s_prev * (YSLOT * 8), &tmap_y0, fb, 32, 16, (z & 7) + 8);
          return;
        }
        tma_load_3d(ring_s + s_prev * (SLOT * 8), &tmap, fb, x0 - 4, y0 - 3, zcoord_base + z + R - 1);
        if (STAGE >= 2) tma_load_3d(yring_s + s_prev * (YSLOT * 8), &tmap_y0, fb, x0, y0, zcoord_base + z + R - 1);
      };
      if constexpr ((Cfg::OPT & 2) != 0) {
        if (warp_u == 0) {                                   // uniform branch: warps 1.. skip the divergent section
          if (lane == 0) produce();
        }
      } else {
        if (tid == 0) produce();
      }
    } else {
      if (tid == 0 && kc >= 4 && kc - 1 + R <= klast) {
        mbar_wait(empty_s + 8 * s_prev, p_prev);
        issue(kc - 1 + R, s_prev);
      }
    }
    // plane z+3 (the newest the stencil needs) has normally landed long ago: probe now, consume the answer later
    uint32_t landed = 0;
    if constexpr ((Cfg::OPT & 1) != 0) landed = mbar_test(land_s + 8 * s_new, p_new);
    Sys::template apply<DZ>(ptA, raw_next, ks);              // the marching dim is shared by my two nodes
    Sys::template apply<DZ>(ptB, raw_next, ks);
    raw_next = Sys::template fetch<DZ>(min(z + 1, NZ - 1), g, ks);

    // early global loads: aux / obstacle pairs, slow-dim neighbours
    double2 y0v = make_double2(0.0, 0.0), auxv = y0v, obsv = y0v;
    if (STAGE == 3 && !SIMPLE && ok0) {
      if (st.comp == HJ_COMP_MIN_WITH_AUX || st.comp == HJ_COMP_MAX_WITH_AUX) auxv = ldg2(st.aux + off);
      if (st.use_obs) obsv = ldg2(st.obs + off);
    }
    // slow dims with few neighbours in flight: the pair loads of ONE slow dim are issued early (they overlap the
    // shared-memory phase); the remaining slow dims are loaded one dim at a time (6 x 16 B in flight per thread)
    // so that a 6-D system does not need 72 registers of neighbour data
    double2 sn0[6];
    if (NSLOW > B0 && ok0) {
#pragma unroll
      for (int k = 0; k < 6; ++k)
        sn0[k] = slow_neighbor(st.in + off, idx[NSLOW - 1], k < 3 ? k - 3 : k - 2, g.N[NSLOW - 1], g.stride[NSLOW - 1],
                               g.bc[NSLOW - 1], g.slope_mult[NSLOW - 1]);
    }

    double pcA[D], hdA[D], pcB[D], hdB[D];
    double L, Rr;
    constexpr bool red = RED;
#define HJ_RED(d, ok)                                              \
  if (red && (ok)) {                                               \
    acc.dmin[d] = fmin(acc.dmin[d], fmin(L, Rr));                  \
    acc.dmax[d] = fmax(acc.dmax[d], fmax(L, Rr));                  \
  }
    // Cfg::SEQ: one dim at a time (loads next to their use, compiler barriers in between) -- trades shared-memory
    // latency hiding inside a warp for a smaller live set, i.e. more resident warps
#define HJ_SEQ_BARRIER if constexpr (Cfg::SEQ) asm volatile("" ::: "memory");
    // PIPE (fast march only): the shared-memory loads of one dim are issued before the arithmetic of the previous
    // dim -- Y loads | X math | Z loads | Y math | X loads of plane z+1 | Z math + Hamiltonian + stage algebra -- so
    // that the shared-memory pipe and the FP64 pipe overlap inside every warp instead of alternating CTA-wide
    constexpr bool PIPE = FAST && (Cfg::OPT & 8) != 0;
#define HJ_PIPE_BARRIER if constexpr (PIPE) asm volatile("" ::: "memory");
    const double* cur = ring + (size_t)s_cur * SLOT;
    // X window: columns ix-4 .. ix+5 of my row (w2 = my pair)
    double2 w0, w1, w2, w3, w4;
    if constexpr (PIPE) {
      w0 = xw[0]; w1 = xw[1]; w2 = xw[2]; w3 = xw[3]; w4 = xw[4];   // loaded during the previous plane
    } else {
      const double2* rowp = reinterpret_cast<const double2*>(cur + myoff);
      w0 = rowp[-2]; w1 = rowp[-1]; w2 = rowp[0]; w3 = rowp[1]; w4 = rowp[2];
    }
    if constexpr (!FAST) {
      // ghost cells of the current plane in X (tiles touching the domain boundary only), in registers
      if (!gw_on && need_patch_x && ok0)
        patch_x(w0, w1, w2, w3, w4, ix, x0, NX, bcx, g.slope_mult[DX], cur + (ty + 3) * BW, st.in + off - ix, NX <= TX);
    }
    const double2 ctr = w2;
    double2 ym3, ym2, ym1, yp1, yp2, yp3;
    auto load_y = [&]() {
      ym3 = lds2(cur + myoff - 3 * BW); ym2 = lds2(cur + myoff - 2 * BW); ym1 = lds2(cur + myoff - 1 * BW);
      yp1 = lds2(cur + myoff + 1 * BW); yp2 = lds2(cur + myoff + 2 * BW); yp3 = lds2(cur + myoff + 3 * BW);
    };
    if constexpr (PIPE) { load_y(); HJ_PIPE_BARRIER }
    // X: node A uses columns ix-3..ix+3 = (w0.y, w1.x, w1.y, w2.x, w2.y, w3.x, w3.y); node B is shifted by one
    pc_hd<WENO, (Cfg::OPT & 16)>(w0.y, w1.x, w1.y, w2.x, w2.y, w3.x, w3.y, g, DX, inv_eps[DX], pcA[DX], hdA[DX], L, Rr, red);
    HJ_RED(DX, ok0)
    pc_hd<WENO, (Cfg::OPT & 16)>(w1.x, w1.y, w2.x, w2.y, w3.x, w3.y, w4.x, g, DX, inv_eps[DX], pcB[DX], hdB[DX], L, Rr, red);
    HJ_RED(DX, ok1)
    HJ_SEQ_BARRIER
    // Z neighbours above: planes z+1, z+2 landed earlier; plane z+3 is the newest one of the ring
    double2 zp1 = make_double2(0.0, 0.0), zp2 = zp1, zp3 = zp1;
    auto load_z = [&]() {
      if (STAGE >= 2) y0v = lds2(yring + (size_t)s_cur * YSLOT + ty * TX + 2 * tp);
      if (ZIN) { zp1 = lds2(ring + (size_t)s_p1 * SLOT + myoff); zp2 = lds2(ring + (size_t)s_p2 * SLOT + myoff); }
      if (!landed) mbar_wait(land_s + 8 * s_new, p_new);
      if (ZIN) zp3 = lds2(ring + (size_t)s_new * SLOT + myoff);
    };
    if constexpr (PIPE) { load_z(); HJ_PIPE_BARRIER }
    {  // Y neighbours of the pair
      if constexpr (!PIPE) load_y();
      if constexpr (!FAST) {
        if (!gw_on && need_patch_y && ok0)
          patch_y<BW>(ym3, ym2, ym1, yp1, yp2, yp3, iy, y0, NY, bcy, g.slope_mult[DY], cur + 4 + 2 * tp,
                      st.in + off - (long long)iy * g.stride[DY], g.stride[DY], NY <= TY);
      }
      pc_hd<WENO, (Cfg::OPT & 16)>(ym3.x, ym2.x, ym1.x, ctr.x, yp1.x, yp2.x, yp3.x, g, DY, inv_eps[DY], pcA[DY], hdA[DY], L, Rr, red);
      HJ_RED(DY, ok0)
      pc_hd<WENO, (Cfg::OPT & 16)>(ym3.y, ym2.y, ym1.y, ctr.y, yp1.y, yp2.y, yp3.y, g, DY, inv_eps[DY], pcB[DY], hdB[DY], L, Rr, red);
      HJ_RED(DY, ok1)
    }
    HJ_SEQ_BARRIER
    {
      if constexpr (!PIPE) load_z();
      if constexpr (!FAST) {
        if (ZIN && bcz == HJ_BC_EXTRAPOLATE && z + 3 >= NZ) {  // ghost planes above the grid: edge plane NZ-1 = z+ke
          const int ke = NZ - 1 - z;                           // 0..2
          const double2 ed = ke == 0 ? ctr : (ke == 1 ? zp1 : zp2);
          const double2 nx = ke == 0 ? q[2] : (ke == 1 ? ctr : zp1);
          const double m = g.slope_mult[DZ];
          if (ke < 1) zp1 = make_double2(ghost_extrapolate(ed.x, nx.x, 1 - ke, m), ghost_extrapolate(ed.y, nx.y, 1 - ke, m));
          if (ke < 2) zp2 = make_double2(ghost_extrapolate(ed.x, nx.x, 2 - ke, m), ghost_extrapolate(ed.y, nx.y, 2 - ke, m));
          zp3 = make_double2(ghost_extrapolate(ed.x, nx.x, 3 - ke, m), ghost_extrapolate(ed.y, nx.y, 3 - ke, m));
        }
      }
      // this warp is done with the current plane's slot
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_s + 8 * s_cur);
      if constexpr (PIPE) {                                    // X window of plane z+1 (resident since two planes ago)
        const double2* rowp = reinterpret_cast<const double2*>(ring + (size_t)s_p1 * SLOT + myoff);
        xw[0] = rowp[-2]; xw[1] = rowp[-1]; xw[2] = rowp[0]; xw[3] = rowp[1]; xw[4] = rowp[2];
        HJ_PIPE_BARRIER
      }
      if constexpr (ZIN) {
        pc_hd<WENO, (Cfg::OPT & 16)>(q[0].x, q[1].x, q[2].x, ctr.x, zp1.x, zp2.x, zp3.x, g, DZ, inv_eps[DZ], pcA[DZ], hdA[DZ], L, Rr, red);
        HJ_RED(DZ, ok0)
        pc_hd<WENO, (Cfg::OPT & 16)>(q[0].y, q[1].y, q[2].y, ctr.y, zp1.y, zp2.y, zp3.y, g, DZ, inv_eps[DZ], pcB[DZ], hdB[DZ], L, Rr, red);
        HJ_RED(DZ, ok1)
      }
    }
#undef HJ_PIPE_BARRIER
    HJ_SEQ_BARRIER
    // slow dims (of the block)
#pragma unroll
    for (int d = NSLOW - 1; d >= B0; --d) {
      double2 sn[6];
      if (d == NSLOW - 1) {
#pragma unroll
        for (int k = 0; k < 6; ++k) sn[k] = sn0[k];
      } else if (ok0) {
#pragma unroll
        for (int k = 0; k < 6; ++k)
          sn[k] = slow_neighbor(st.in + off, idx[d], k < 3 ? k - 3 : k - 2, g.N[d], g.stride[d], g.bc[d], g.slope_mult[d]);
      }
      pc_hd<WENO, (Cfg::OPT & 16)>(sn[0].x, sn[1].x, sn[2].x, ctr.x, sn[3].x, sn[4].x, sn[5].x, g, d, inv_eps[d], pcA[d], hdA[d], L, Rr, red);
      HJ_RED(d, ok0)
      pc_hd<WENO, (Cfg::OPT & 16)>(sn[0].y, sn[1].y, sn[2].y, ctr.y, sn[3].y, sn[4].y, sn[5].y, g, d, inv_eps[d], pcB[d], hdB[d], L, Rr, red);
      HJ_RED(d, ok1)
    }
#undef HJ_RED
#undef HJ_SEQ_BARRIER

    // Hamiltonian + GLF dissipation (artificial_diss_glf.py:100: diss += 0.5*(R-L)*alpha)
    double hamA, hamB;
    if constexpr ((Cfg::OPT & 16) != 0) { hamA = pcA[DX] + pcA[DY]; hamB = pcB[DX] + pcB[DY]; }
    else { hamA = Sys::ham(ptA, pcA, ks); hamB = Sys::ham(ptB, pcB, ks); }
    double ydA = -hamA, ydB = -hamB;                         // ydot = -(ham - diss)
#pragma unroll
    for (int d = B0; d < D; ++d) {
      const double aA = Sys::alpha(d - B0, ptA, ks), aB = Sys::alpha(d - B0, ptB, ks);
      ydA = fma(hdA[d], aA, ydA);
      ydB = fma(hdB[d], aB, ydB);
      if (red) {
        if (ok0) acc.amax[d] = fmax(acc.amax[d], aA);
        if (ok1) acc.amax[d] = fmax(acc.amax[d], aB);
      }
    }

    if constexpr (!SIMPLE) {
      ydA = restrict_update(ydA, st.restrict_sign);
      ydB = restrict_update(ydB, st.restrict_sign);
    }
    // RK stage algebra + driver epilogue (see stage_update in hj_common.cuh), on the pair
    double oA, oB;
    if (STAGE == 0) { oA = ydA; oB = ydB; }
    else if (STAGE == 1) { oA = ctr.x + dt * ydA; oB = ctr.y + dt * ydB; }
    else if (STAGE == 2) {
      oA = 0.25 * (3.0 * y0v.x + (ctr.x + dt * ydA));
      oB = 0.25 * (3.0 * y0v.y + (ctr.y + dt * ydB));
    } else {
      oA = st.fin_a * (y0v.x + st.fin_b * (ctr.x + dt * ydA));
      oB = st.fin_a * (y0v.y + st.fin_b * (ctr.y + dt * ydB));
      if constexpr (SIMPLE) {
        oA = nan_min(oA, y0v.x); oB = nan_min(oB, y0v.y);
      } else {
        const bool with_aux = st.comp == HJ_COMP_MIN_WITH_AUX || st.comp == HJ_COMP_MAX_WITH_AUX;
        oA = comp_epilogue(oA, st.comp, y0v.x, with_aux ? auxv.x : 0.0);
        oB = comp_epilogue(oB, st.comp, y0v.y, with_aux ? auxv.y : 0.0);
      }
      if (!SIMPLE && st.use_obs) { oA = nan_max(oA, -obsv.x); oB = nan_max(oB, -obsv.y); }
    }
    if (FAST && (Cfg::OPT & 64)) { if (oA == 1.2345e300) st.out[off] = oB; }      // tuning harness only: no store
    else if (FAST && (Cfg::OPT & 4) && Cfg::NACTIVE == Cfg::NCONS && (!Cfg::GW || full_tile)) *reinterpret_cast<double2*>(st.out + off) = make_double2(oA, oB);
    else if (ok1) *reinterpret_cast<double2*>(st.out + off) = make_double2(oA, oB);
    else if (ok0) st.out[off] = oA;
    if (red && ((ok0 && oA != oA) || (ok1 && oB != oB))) acc.nan = 1;

    q[0] = q[1]; q[1] = q[2]; q[2] = ctr;
    ++z; ++kc; off += zstride;
    s_prev = s_cur; p_prev = p_cur;
    s_cur = s_p1; if (s_cur == 0) p_cur ^= 1;
    s_p1 = s_p2; s_p2 = s_new;
    if (++s_new == (unsigned)R) { s_new = 0; p_new ^= 1; }
  };

  // head (general body) -> fast segment z in [z0+1, z1-R] (the prefetched plane z+R-1 <= z1-1 <= NZ-1) -> tail
  const int zf_end = ((need_patch_x || need_patch_y) && !gw_on) ? z0 : z1 - R + 1;
  plane.template operator()<false>();
  if ((Cfg::OPT & 8) != 0 && z < zf_end) {                   // pipeline prologue: X window of the first fast plane
    const double2* rowp = reinterpret_cast<const double2*>(ring + (size_t)s_cur * SLOT + myoff);
    xw[0] = rowp[-2]; xw[1] = rowp[-1]; xw[2] = rowp[0]; xw[3] = rowp[1]; xw[4] = rowp[2];
  }
  // SIMPLE: the fast march specialised for the common epilogue (no termRestrictUpdate; stage 3 = minVOverTime
  // without obstacle), so that the steady-state loop carries no epilogue dispatch
  const bool simple = (Cfg::OPT & 128) != 0 && st.restrict_sign == 0 &&
                      (STAGE != 3 || (st.comp == HJ_COMP_MIN_OVER_TIME && !st.use_obs));
  if (simple) {
    while (z < zf_end) plane.template operator()<true, true>();
  } else {
    while (z < zf_end) {
#pragma unroll
      for (int u = 0; u < Cfg::UNROLL; ++u) {
        plane.template operator()<true>();
        if (z >= zf_end) break;
      }
    }
  }
  while (z < z1) plane.template operator()<false>();
  if (RED) acc.flush(st.red);
}

}  // namespace hjtma
