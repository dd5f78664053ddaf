// hj_tma.cu -- TMA backend: the fused RHS + RK-stage kernel as a streamed plane ring.
//
// Layout of the work (grid dims named X = D-1 (contiguous), Y = D-2, Z = D-3, "slow" = 0..D-4):
//   * a CTA owns a TY x TX tile of (Y, X) and marches a chunk of CZ planes along Z;
//   * each Z-plane of the tile, with its 3-cell X/Y halo (box (TY+6) x (TX+8) doubles; 4 columns on each side in X
//     keep every row and every thread's 2-node pair 16-byte aligned), is brought into a ring of R shared-memory
//     slots by ONE cp.async.bulk.tensor (TMA) per plane, signalled on an mbarrier; out-of-range box parts are
//     zero-filled by the TMA unit and then overwritten with extrapolated / periodic ghost cells by the threads
//     (boundary CTAs only) -- no padded copy of the field ever exists;
//   * every thread owns two X-adjacent nodes: the Z stencil lives in a 7-deep register queue (one 16-byte
//     shared-memory read per plane), the X and Y stencils are read from the current plane's slot with 16-byte
//     LDS; slow-dim neighbours (D >= 4) come straight from L2/HBM with 16-byte read-only loads;
//   * R = 8: planes z+1..z+3 are needed, z+4..z+8 are prefetch distance, so HBM latency is hidden by the ring
//     rather than by occupancy;
//   * the Hamiltonian, GLF dissipation, RK stage algebra and the driver epilogue are applied in registers and
//     the result leaves with one 16-byte store per thread.  DRAM traffic per node: 8 B read + 8 B write
//     (+8 B for y0 in stages 2/3) = the algorithmic 16/24/24 B.
//
// Reference behaviour restated: see hj_common.cuh header.
#include <cuda.h>

#include <cstdio>
#include <cstring>

#include "hj_internal.h"
#include "hj_systems.cuh"

namespace {

constexpr int R = 8;          // ring slots

struct TmaGeom {
  int nxt, nyt, nzc, cz;      // tiles in X, Y; Z chunks; planes per chunk
  long long nslow;            // product of slow dims
  long long zcoord0;          // TMA dim-2 coordinate of (slow = 0, z = 0)  (halo planes on dim 0 shift it)
  int NZ;
};

// ------------------------------------------------------------------------------------------ PTX helpers
HJ_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
HJ_DEV void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
HJ_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
HJ_DEV void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
HJ_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
HJ_DEV void tma_load_3d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
HJ_DEV double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

// ------------------------------------------------------------------------------------------ stencil math on pairs
// derivC and 0.5*(derivR - derivL) of the as-shipped (fixed-weight) scheme straight from the 7 nodes; the
// coefficients are host-precomputed (KGrid::ca1..cb) and used as constant-bank operands:
//   derivC    = ca1 (v4-v2) + ca2 (v5-v1) + ca3 (v6-v0)                       (= 0.5*(L+R), 6 flops)
//   0.5*(R-L) = cb (v0+v6 - 6 (v1+v5) + 15 (v2+v4) - 20 v3)                   (sixth difference, 7 flops)
// i.e. 13 fp64 instructions per node per dim instead of ~60 for the divided-difference tables + weightWENO.
template <int WENO>
HJ_DEV void pc_hd(const double v0, const double v1, const double v2, const double v3, const double v4, const double v5,
                  const double v6, const KGrid& g, const int d, double inv_eps, double& pc, double& hd, double& L,
                  double& Rr, const bool need_lr) {
  if (WENO == HJ_WENO_AS_SHIPPED) {
    pc = g.ca1[d] * (v4 - v2) + g.ca2[d] * (v5 - v1) + g.ca3[d] * (v6 - v0);
    double t = fma(-6.0, v1 + v5, v0 + v6);
    t = fma(15.0, v2 + v4, t);
    t = fma(-20.0, v3, t);
    hd = g.cb[d] * t;
    if (need_lr) { L = pc - hd; Rr = pc + hd; }
  } else {
    const double v[7] = {v0, v1, v2, v3, v4, v5, v6};
    upwind5_weno(v, g.dxinv[d], inv_eps, L, Rr);
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

// ------------------------------------------------------------------------------------------ the kernel
// Producer/consumer ring without a CTA-wide barrier in the steady state:
//   full[s]  : armed by the producer thread (expect_tx), completed by the TMA unit when plane box s has landed
//   empty[s] : one arrival per warp when that warp no longer needs the plane in slot s
// All 8 warps are consumers (one node pair per thread); lane 0 of warp 0 doubles as the producer: at the top of
// step z it waits until every warp has released plane z-1 (normally already true: the ring runs R-4 = 4 planes
// ahead of need) and re-arms that slot with plane z+R-1.  Consumer warps therefore never wait for each other in
// interior tiles; only tiles that touch the domain boundary pay a 256-thread named barrier per plane for the
// ghost-cell patch.
constexpr int NCONS_WARPS = 8;
constexpr int NCONS = NCONS_WARPS * 32;          // 256 threads
constexpr int NTHREADS_WS = NCONS;

template <class Sys, int WENO, int TX, int TY, bool RED, int STAGE>
__global__ void __launch_bounds__(NTHREADS_WS, 2)
k_stage_tma(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_y0, const KGrid g,
            const KSys ks, const KStage st, const TmaGeom geo) {
  constexpr int D = Sys::ND;
  static_assert(D >= 3, "the plane-ring kernel needs a Z dim");
  constexpr int DX = D - 1, DY = D - 2, DZ = D - 3, NSLOW = D - 3;
  constexpr int PAIRS = TX / 2;
  static_assert(PAIRS * TY == NCONS, "tile must give every consumer thread one node pair");
  constexpr int BW = TX + 8, BH = TY + 6, SLOT = BW * BH;        // doubles
  static_assert((SLOT * 8) % 128 == 0, "slot must keep 128-byte alignment");
  static_assert((R & (R - 1)) == 0, "R must be a power of two");
  // stages 2/3 also stream the un-haloed y0 tile (TY x TX) of each plane through the ring, on the same barrier
  constexpr int YSLOT = (STAGE >= 2) ? TX * TY : 0;
  static_assert((YSLOT * 8) % 128 == 0, "y0 slot must keep 128-byte alignment");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);
  double* yring = ring + (size_t)R * SLOT;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)R * (SLOT + YSLOT) * 8);
  uint64_t* empty = full + R;

  const int tid = threadIdx.x;
  long long b = blockIdx.x;
  const int xt = (int)(b % geo.nxt); b /= geo.nxt;
  const int yt = (int)(b % geo.nyt); b /= geo.nyt;
  const int zc = (int)(b % geo.nzc); b /= geo.nzc;
  const long long slow_flat = b;
  const int NX = g.N[DX], NY = g.N[DY], NZ = g.N[DZ];
  const int x0 = xt * TX, y0 = yt * TY, z0 = zc * geo.cz;
  const int z1 = min(z0 + geo.cz, NZ);
  const int bcx = g.bc[DX], bcy = g.bc[DY], bcz = g.bc[DZ];
  const unsigned klast = (unsigned)((z1 - 1 + 3) - (z0 - 3));   // ring position of the last plane this chunk needs
  const int zcoord_base = (int)(geo.zcoord0 + slow_flat * NZ);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < R; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCONS_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // plane with ring position k -> slot k & (R-1): TMA load, or a bare arrival for a computed ghost plane
  auto issue = [&](unsigned k) {
    const unsigned s = k & (R - 1);
    const int zp = z0 - 3 + (int)k;
    int zsrc = zp;
    bool load = true;
    if (zp < 0 || zp >= NZ) {
      if (bcz == HJ_BC_PERIODIC) zsrc = zp < 0 ? zp + NZ : zp - NZ;
      else if (bcz == HJ_BC_EXTRAPOLATE) load = false;          // ghost plane: computed from the register queue
    }
    // the y0 tile rides along for planes that will be "current" (ring positions 3 .. klast-3)
    const bool ytile = STAGE >= 2 && k >= 3 && k + 3 <= klast;
    if (load) {
      mbar_expect_tx(&full[s], (SLOT + (ytile ? YSLOT : 0)) * 8);
      tma_load_3d(ring + (size_t)s * SLOT, &tmap, &full[s], x0 - 4, y0 - 3, zcoord_base + zsrc);
      if (ytile) tma_load_3d(yring + (size_t)s * YSLOT, &tmap_y0, &full[s], x0, y0, zcoord_base + zp);
    } else {
      mbar_arrive(&full[s]);
    }
  };
  if (tid == 0) {
    for (unsigned k = 0; k < R && k <= klast; ++k) issue(k);
  }

  // ================================================================== consumer warps
  const int lane = tid & 31;
  const int tp = tid % PAIRS, ty = tid / PAIRS;
  int idx[D];
  {
    long long r = slow_flat;
#pragma unroll
    for (int d = NSLOW - 1; d >= 0; --d) { idx[d] = (int)(r % g.N[d]); r /= g.N[d]; }
  }
  const int ix = x0 + 2 * tp, iy = y0 + ty;
  const bool ok0 = ix < NX && iy < NY, ok1 = ix + 1 < NX && iy < NY;
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
  typename Sys::Pt ptA = Sys::load(idx, g, ks);
  idx[DX] = min(ix + 1, NX - 1);
  typename Sys::Pt ptB = Sys::load(idx, g, ks);

  const int myoff = (ty + 3) * BW + 4 + 2 * tp;              // my pair inside a slot (doubles); 16-byte aligned
  const bool need_patch_x = (bcx != HJ_BC_HALO) && (x0 - 3 < 0 || x0 + TX + 2 >= NX);
  const bool need_patch_y = (bcy != HJ_BC_HALO) && (y0 - 3 < 0 || y0 + TY + 2 >= NY);

  RedAcc<D> acc;
  acc.init();

  // ---- prologue: fill the Z register queue with planes z0-3 .. z0+2, then hand slots 0..2 back
  double2 q[7];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    mbar_wait(&full[k], 0);
    q[k] = *reinterpret_cast<const double2*>(ring + (size_t)k * SLOT + myoff);
  }
  if (bcz == HJ_BC_EXTRAPOLATE && z0 == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {                            // planes -3,-2,-1 from planes 0,1
      q[k].x = ghost_extrapolate(q[3].x, q[4].x, 3 - k, g.slope_mult[DZ]);
      q[k].y = ghost_extrapolate(q[3].y, q[4].y, 3 - k, g.slope_mult[DZ]);
    }
  }
  __syncwarp();
  if (lane == 0) { mbar_arrive(&empty[0]); mbar_arrive(&empty[1]); mbar_arrive(&empty[2]); }
  if (tid == 0) {
    for (unsigned k = R; k < R + 3 && k <= klast; ++k) {     // planes z0+5..z0+7 into the slots of z0-3..z0-1
      mbar_wait(&empty[k & (R - 1)], 0);
      issue(k);
    }
  }

  // ---- march
  unsigned kc = 3;                                           // ring position of the current plane
  double2 raw_next = Sys::template fetch<DZ>(z0, g, ks);
  for (int z = z0; z < z1; ++z, ++kc, off += zstride) {
    if (tid == 0 && kc >= 4 && kc - 1 + R <= klast) {       // producer duty: recycle the slot of plane z-1
      const unsigned kp = kc - 1;
      mbar_wait(&empty[kp & (R - 1)], (kp / R) & 1);
      issue(kp + R);
    }
    Sys::template apply<DZ>(ptA, raw_next, ks);              // the marching dim is shared by my two nodes
    Sys::template apply<DZ>(ptB, raw_next, ks);
    raw_next = Sys::template fetch<DZ>(min(z + 1, NZ - 1), g, ks);

    // early global loads: aux / obstacle pairs, slow-dim neighbours
    double2 y0v = make_double2(0.0, 0.0), auxv = y0v, obsv = y0v;
    if (STAGE == 3 && ok0) {
      if (st.comp == HJ_COMP_MIN_WITH_AUX || st.comp == HJ_COMP_MAX_WITH_AUX) auxv = ldg2(st.aux + off);
      if (st.use_obs) obsv = ldg2(st.obs + off);
    }
    double2 sn[NSLOW > 0 ? NSLOW : 1][6];
    if (ok0) {
#pragma unroll
      for (int d = 0; d < NSLOW; ++d) {
#pragma unroll
        for (int k = 0; k < 6; ++k)
          sn[d][k] = slow_neighbor(st.in + off, idx[d], k < 3 ? k - 3 : k - 2, g.N[d], g.stride[d], g.bc[d], g.slope_mult[d]);
      }
    }

    // newest plane (z+3): into the queue, or a computed ghost plane
    {
      const unsigned k = kc + 3, s = k & (R - 1);
      mbar_wait(&full[s], (k / R) & 1);
      if (bcz == HJ_BC_EXTRAPOLATE && z + 3 >= NZ) {
        const int dist = z + 3 - (NZ - 1);                   // 1..3 ; edge plane NZ-1 sits at queue index 6-dist
        const double2 ed = dist == 1 ? q[5] : (dist == 2 ? q[4] : q[3]);
        const double2 nx = dist == 1 ? q[4] : (dist == 2 ? q[3] : q[2]);
        q[6].x = ghost_extrapolate(ed.x, nx.x, dist, g.slope_mult[DZ]);
        q[6].y = ghost_extrapolate(ed.y, nx.y, dist, g.slope_mult[DZ]);
      } else {
        q[6] = *reinterpret_cast<const double2*>(ring + (size_t)s * SLOT + myoff);
      }
    }
    const unsigned scur = kc & (R - 1);
    double* cur = ring + (size_t)scur * SLOT;
    if (STAGE >= 2) y0v = *reinterpret_cast<const double2*>(yring + (size_t)scur * YSLOT + ty * TX + 2 * tp);

    // ghost cells of the current plane in X / Y (tiles touching the domain boundary only)
    if (need_patch_x || need_patch_y) {
      if (need_patch_x) {
        for (int e = tid; e < 6 * TY; e += NCONS) {
          const int r = e / 6 + 3, j = e % 6;
          const int x = j < 3 ? j - 3 : NX + (j - 3);        // ghost node index
          const int c = x - x0 + 4;
          if (c < 1 || c >= BW - 1 || y0 + r - 3 >= NY) continue;
          double val;
          if (bcx == HJ_BC_PERIODIC) {
            const int xs = x < 0 ? x + NX : x - NX;
            val = __ldg(st.in + off - ix - (long long)ty * g.stride[DY] + (long long)(r - 3) * g.stride[DY] + xs);
          } else {
            const int e0 = x < 0 ? 0 : NX - 1, e1 = x < 0 ? 1 : NX - 2, dist = x < 0 ? -x : x - (NX - 1);
            val = ghost_extrapolate(cur[r * BW + e0 - x0 + 4], cur[r * BW + e1 - x0 + 4], dist, g.slope_mult[DX]);
          }
          cur[r * BW + c] = val;
        }
      }
      if (need_patch_y) {
        for (int e = tid; e < 6 * TX; e += NCONS) {
          const int cc = e % TX + 4, j = e / TX;
          const int y = j < 3 ? j - 3 : NY + (j - 3);
          const int r = y - y0 + 3;
          if (r < 0 || r >= BH || x0 + cc - 4 >= NX) continue;
          double val;
          if (bcy == HJ_BC_PERIODIC) {
            const int ys = y < 0 ? y + NY : y - NY;
            val = __ldg(st.in + off - ix - (long long)iy * g.stride[DY] + (long long)ys * g.stride[DY] + (x0 + cc - 4));
          } else {
            const int e0 = y < 0 ? 0 : NY - 1, e1 = y < 0 ? 1 : NY - 2, dist = y < 0 ? -y : y - (NY - 1);
            val = ghost_extrapolate(cur[(e0 - y0 + 3) * BW + cc], cur[(e1 - y0 + 3) * BW + cc], dist, g.slope_mult[DY]);
          }
          cur[r * BW + cc] = val;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // patched cells will later be overwritten by TMA
      asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory");
    }

    // X window: columns c-4 .. c+5 of my row (c = my pair's first column)
    const double2* rowp = reinterpret_cast<const double2*>(cur + myoff);
    const double2 w0 = rowp[-2], w1 = rowp[-1], w2 = rowp[0], w3 = rowp[1], w4 = rowp[2];
    // Y neighbours of the pair
    const double2 ym3 = *reinterpret_cast<const double2*>(cur + myoff - 3 * BW);
    const double2 ym2 = *reinterpret_cast<const double2*>(cur + myoff - 2 * BW);
    const double2 ym1 = *reinterpret_cast<const double2*>(cur + myoff - 1 * BW);
    const double2 yp1 = *reinterpret_cast<const double2*>(cur + myoff + 1 * BW);
    const double2 yp2 = *reinterpret_cast<const double2*>(cur + myoff + 2 * BW);
    const double2 yp3 = *reinterpret_cast<const double2*>(cur + myoff + 3 * BW);
    const double2 ctr = q[3];
    // this warp is done with the current plane's slot
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[scur]);

    double pcA[D], hdA[D], pcB[D], hdB[D];
    double L, Rr;
    constexpr bool red = RED;
#define HJ_RED(d, ok)                                              \
  if (red && (ok)) {                                               \
    acc.dmin[d] = fmin(acc.dmin[d], fmin(L, Rr));                  \
    acc.dmax[d] = fmax(acc.dmax[d], fmax(L, Rr));                  \
  }
    // X: node A uses columns c-3..c+3 = (w0.y, w1.x, w1.y, w2.x, w2.y, w3.x, w3.y); node B is shifted by one
    pc_hd<WENO>(w0.y, w1.x, w1.y, w2.x, w2.y, w3.x, w3.y, g, DX, inv_eps[DX], pcA[DX], hdA[DX], L, Rr, red);
    HJ_RED(DX, ok0)
    pc_hd<WENO>(w1.x, w1.y, w2.x, w2.y, w3.x, w3.y, w4.x, g, DX, inv_eps[DX], pcB[DX], hdB[DX], L, Rr, red);
    HJ_RED(DX, ok1)
    // Y
    pc_hd<WENO>(ym3.x, ym2.x, ym1.x, ctr.x, yp1.x, yp2.x, yp3.x, g, DY, inv_eps[DY], pcA[DY], hdA[DY], L, Rr, red);
    HJ_RED(DY, ok0)
    pc_hd<WENO>(ym3.y, ym2.y, ym1.y, ctr.y, yp1.y, yp2.y, yp3.y, g, DY, inv_eps[DY], pcB[DY], hdB[DY], L, Rr, red);
    HJ_RED(DY, ok1)
    // Z (register queue)
    pc_hd<WENO>(q[0].x, q[1].x, q[2].x, q[3].x, q[4].x, q[5].x, q[6].x, g, DZ, inv_eps[DZ], pcA[DZ], hdA[DZ], L, Rr, red);
    HJ_RED(DZ, ok0)
    pc_hd<WENO>(q[0].y, q[1].y, q[2].y, q[3].y, q[4].y, q[5].y, q[6].y, g, DZ, inv_eps[DZ], pcB[DZ], hdB[DZ], L, Rr, red);
    HJ_RED(DZ, ok1)
    // slow dims
#pragma unroll
    for (int d = 0; d < NSLOW; ++d) {
      pc_hd<WENO>(sn[d][0].x, sn[d][1].x, sn[d][2].x, ctr.x, sn[d][3].x, sn[d][4].x, sn[d][5].x, g, d, inv_eps[d], pcA[d], hdA[d], L, Rr, red);
      HJ_RED(d, ok0)
      pc_hd<WENO>(sn[d][0].y, sn[d][1].y, sn[d][2].y, ctr.y, sn[d][3].y, sn[d][4].y, sn[d][5].y, g, d, inv_eps[d], pcB[d], hdB[d], L, Rr, red);
      HJ_RED(d, ok1)
    }
#undef HJ_RED

    // Hamiltonian + GLF dissipation (artificial_diss_glf.py:100: diss += 0.5*(R-L)*alpha)
    const double hamA = Sys::ham(ptA, pcA, ks), hamB = Sys::ham(ptB, pcB, ks);
    double ydA = -hamA, ydB = -hamB;                         // ydot = -(ham - diss)
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double aA = Sys::alpha(d, ptA, ks), aB = Sys::alpha(d, ptB, ks);
      ydA = fma(hdA[d], aA, ydA);
      ydB = fma(hdB[d], aB, ydB);
      if (red) {
        if (ok0) acc.amax[d] = fmax(acc.amax[d], aA);
        if (ok1) acc.amax[d] = fmax(acc.amax[d], aB);
      }
    }

    // RK stage algebra + driver epilogue (see stage_update in hj_common.cuh), on the pair
    double oA, oB;
    if (STAGE == 0) { oA = ydA; oB = ydB; }
    else if (STAGE == 1) { oA = ctr.x + st.dt * ydA; oB = ctr.y + st.dt * ydB; }
    else if (STAGE == 2) {
      oA = 0.25 * (3.0 * y0v.x + (ctr.x + st.dt * ydA));
      oB = 0.25 * (3.0 * y0v.y + (ctr.y + st.dt * ydB));
    } else {
      oA = (1.0 / 3.0) * (y0v.x + 2.0 * (ctr.x + st.dt * ydA));
      oB = (1.0 / 3.0) * (y0v.y + 2.0 * (ctr.y + st.dt * ydB));
      switch (st.comp) {
        case HJ_COMP_MIN_OVER_TIME: oA = fmin(oA, y0v.x); oB = fmin(oB, y0v.y); break;
        case HJ_COMP_MAX_OVER_TIME: oA = fmax(oA, y0v.x); oB = fmax(oB, y0v.y); break;
        case HJ_COMP_MIN_WITH_AUX: oA = fmin(oA, auxv.x); oB = fmin(oB, auxv.y); break;
        case HJ_COMP_MAX_WITH_AUX: oA = fmax(oA, auxv.x); oB = fmax(oB, auxv.y); break;
        default: break;
      }
      if (st.use_obs) { oA = fmax(oA, -obsv.x); oB = fmax(oB, -obsv.y); }
    }
    if (ok1) *reinterpret_cast<double2*>(st.out + off) = make_double2(oA, oB);
    else if (ok0) st.out[off] = oA;
    if (red && ((ok0 && oA != oA) || (ok1 && oB != oB))) acc.nan = 1;

#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] = q[k + 1];
  }
  if (RED) acc.flush(st.red);
}

}  // namespace

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct HjTmaPlan {
  CUtensorMap tmap[3];
  CUtensorMap tmap_y0;     // un-haloed TY x TX box on buffer 0 (y at the start of the step)
  TmaGeom geo;
  int tx, ty;
  size_t smem;
  long long nblocks;
};

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

template <class Sys, int WENO, int TX, int TY, bool RED, int STAGE>
static cudaError_t launch_one(const HjTmaPlan* p, const CUtensorMap& tm, const KGrid& g, const KSys& ks,
                              const KStage& st, cudaStream_t s) {
  auto kern = k_stage_tma<Sys, WENO, TX, TY, RED, STAGE>;
  const size_t smem = (size_t)R * ((TX + 8) * (TY + 6) + (STAGE >= 2 ? TX * TY : 0)) * 8 + 2 * R * 8;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  kern<<<(unsigned)p->nblocks, NTHREADS_WS, smem, s>>>(tm, p->tmap_y0, g, ks, st, p->geo);
  return cudaGetLastError();
}

struct TmaLauncher {
  const HjTmaPlan* p;
  const CUtensorMap& tm;
  int weno;
  const KGrid& g;
  const KSys& ks;
  const KStage& st;
  cudaStream_t s;
  cudaError_t err = cudaSuccess;
  template <class Sys, int WENO>
  cudaError_t by_stage(bool red) {
    switch (st.stage) {
      case 1: return red ? launch_one<Sys, WENO, 32, 16, true, 1>(p, tm, g, ks, st, s) : launch_one<Sys, WENO, 32, 16, false, 1>(p, tm, g, ks, st, s);
      case 2: return red ? launch_one<Sys, WENO, 32, 16, true, 2>(p, tm, g, ks, st, s) : launch_one<Sys, WENO, 32, 16, false, 2>(p, tm, g, ks, st, s);
      case 3: return red ? launch_one<Sys, WENO, 32, 16, true, 3>(p, tm, g, ks, st, s) : launch_one<Sys, WENO, 32, 16, false, 3>(p, tm, g, ks, st, s);
      default: return cudaErrorNotSupported;
    }
  }
  template <class Sys>
  void operator()() {
    if constexpr (Sys::ND >= 3) {
      const bool red = st.want_reduce != 0;
      if (weno == HJ_WENO_AS_SHIPPED) err = by_stage<Sys, HJ_WENO_AS_SHIPPED>(red);
      else err = by_stage<Sys, HJ_WENO_INTENDED>(red);
    } else {
      err = cudaErrorNotSupported;
    }
  }
};

HjTmaPlan* hj_tma_plan_create(const KGrid& g, int system_id, int weno, double* const bufs[3], int halo0, char* err,
                              int errlen) {
  (void)weno;
  const int D = g.D;
  if (D < 3) { snprintf(err, errlen, "2-D grids use the gather backend"); return nullptr; }
  if (hj_system_ndim(system_id) != D) { snprintf(err, errlen, "system/grid dim mismatch"); return nullptr; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) { snprintf(err, errlen, "cuTensorMapEncodeTiled not available from the driver"); return nullptr; }
  const int TX = 32, TY = 16;
  const int NX = g.N[D - 1], NY = g.N[D - 2], NZ = g.N[D - 3];
  if (NZ < 4) { snprintf(err, errlen, "Z extent too small"); return nullptr; }
  const long long pitch = g.stride[D - 2];
  if (pitch % 2) { snprintf(err, errlen, "row pitch must be even"); return nullptr; }
  long long nslow = 1;
  for (int d = 0; d < D - 3; ++d) nslow *= g.N[d];
  // number of (Y,X) planes in a buffer, counting halo planes of dim 0
  long long planes = nslow * NZ;
  long long zcoord0 = 0;
  if (halo0) {
    long long per0 = planes / g.N[0];          // (Y,X) planes per dim-0 plane
    planes += 2LL * HJ_GHOST * per0;
    zcoord0 = (long long)HJ_GHOST * per0;
  }
  if (planes > 0x7fffffffLL) { snprintf(err, errlen, "too many planes for a 32-bit TMA coordinate"); return nullptr; }
  HjTmaPlan* p = new HjTmaPlan();
  p->tx = TX; p->ty = TY;
  p->geo.nxt = (NX + TX - 1) / TX;
  p->geo.nyt = (NY + TY - 1) / TY;
  // Z chunk: long enough to amortise the 6-plane lead-in, short enough to give >= ~4 waves of CTAs
  int cz = NZ;
  const long long tiles = (long long)p->geo.nxt * p->geo.nyt * nslow;
  while (cz > 32 && tiles * ((NZ + cz - 1) / cz) < 148LL * 2 * 4) cz = (cz + 1) / 2;
  if (cz > 128) cz = 128;
  p->geo.cz = cz;
  p->geo.nzc = (NZ + cz - 1) / cz;
  p->geo.nslow = nslow;
  p->geo.zcoord0 = zcoord0;
  p->geo.NZ = NZ;
  p->nblocks = tiles * p->geo.nzc;
  p->smem = (size_t)R * (TX + 8) * (TY + 6) * 8 + 2 * R * 8;
  if (p->nblocks > 0x7fffffffLL) { delete p; snprintf(err, errlen, "grid too large"); return nullptr; }
  for (int bidx = 0; bidx < 4; ++bidx) {                    // 0..2: haloed boxes on the RK buffers; 3: y0 tile on buffer 0
    cuuint64_t dims[3] = {(cuuint64_t)NX, (cuuint64_t)NY, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 8, (cuuint64_t)pitch * NY * 8};
    cuuint32_t box[3] = {(cuuint32_t)(bidx < 3 ? TX + 8 : TX), (cuuint32_t)(bidx < 3 ? TY + 6 : TY), 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(bidx < 3 ? &p->tmap[bidx] : &p->tmap_y0, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
                     (void*)bufs[bidx < 3 ? bidx : 0], dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      snprintf(err, errlen, "cuTensorMapEncodeTiled failed (%d)", (int)r);
      delete p;
      return nullptr;
    }
  }
  return p;
}

void hj_tma_plan_destroy(HjTmaPlan* p) { delete p; }

cudaError_t hj_launch_stage_tma(HjTmaPlan* plan, int system_id, int weno, const KGrid& g, const KSys& ks,
                                const KStage& st, int in_buf, cudaStream_t s) {
  TmaLauncher l{plan, plan->tmap[in_buf], weno, g, ks, st, s};
  if (!hj_dispatch_system(system_id, l)) return cudaErrorInvalidValue;
  if (l.err == cudaSuccess) hj_count_launch(1);
  return l.err;
}
