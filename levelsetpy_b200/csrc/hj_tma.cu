// hj_tma.cu -- TMA backend, host side: tensor maps, tile geometry and the per-system launchers of the plane-ring
// kernel (device side: hj_tma_kernel.cuh).
#include <cuda.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "hj_internal.h"
#include "hj_tma_kernel.cuh"
#include "hj_tma_plan.h"
#include "hj_vec_kernel.cuh"

using namespace hjtma;

// production configurations (see tools/tune_tma.cu for the measured alternatives).
//   whole 3-D systems: 32 x 16 tile, 8-slot ring, 2 CTAs/SM.
//   product systems (dimension-split, hj_vec_kernel.cuh): pass 1 takes a tile shaped for the trailing block's
//   plane (42 x 12 for the 41 x 41 planes of the 6-D pair -- measured faster than the bank-conflict-free 48 x 7
//   at 6 warps per CTA; 54 x 9 for the 161 x 161 planes of the 4-D pair), pass 2 a {vector pairs, TA, TB} tile of
//   the leading block.  Pass 2 of the 6-D pair is bound by the TMA unit's row rate: 128-byte rows (8 pairs) with a
//   4 x 7 tile and a 6-slot ring (2 CTAs/SM) measured best: 53.6 / 65.5 / 65.3 ms per launch at 41^6 against
//   73.7 / 90.5 / 90.0 ms for 64-byte rows with a 7 x 7 tile.
using ProdCfg = TmaCfg<8, 2, 1>;
// whole systems whose planes are a few tiles wide run a ghost-warp configuration (hj_tma_kernel.cuh): the Flock batch
// (101 x 101 planes, rows of 101 = one 102-wide tile, 13 tiles of 8 rows; 13 consumer warps + the ghost warp, 1 CTA/SM)
template <class Sys> struct WholeCfg { using type = ProdCfg; };
#ifndef HJ_FB_TY
#define HJ_FB_TY 8
#define HJ_FB_TXP 51
#define HJ_FB_MINB 1
#define HJ_FB_GW 1
#define HJ_FB_XPAD 0
#define HJ_FB_R 8
#endif
template <> struct WholeCfg<SysFlockBatch> { using type = TmaCfg<HJ_FB_R, HJ_FB_MINB, 1, HJ_FB_TY, HJ_FB_TXP, false, 143, HJ_FB_GW, HJ_FB_XPAD>; };
// 'intended' (true) WENO5 of a whole system keeps the production tile: the body spills 90-400 bytes per thread at the
// 128 registers of two 256-thread CTAs per SM, but one 384-thread CTA at 168 registers (no spills) measured SLOWER
// (256^3: 0.63 / 0.59 / 1.81 ms per stage against 0.58 / 0.51 / 0.55 ms) -- the FP64 pipe is the bound (70 % busy) and
// 16 resident warps feed it better than 12.  The ENO functors (divided-difference tables of both sides live together)
// take the roomy configuration.
#ifndef HJ_INT_TY
#define HJ_INT_TY 24
#define HJ_INT_MINB 1
#endif
template <class Sys, int WENO> struct WholeCfgW { using type = typename WholeCfg<Sys>::type; };
// upwindFirstENO3a / upwindFirstENO2 as CoStateCalc on the plane-ring backend: whole 3-D systems, the
// same roomy configuration (the divided-difference tables of both sides are live together)
template <> struct WholeCfgW<SysDubinsRel, HJ_SCHEME_ENO3A> { using type = TmaCfg<8, HJ_INT_MINB, 1, HJ_INT_TY, 16>; };
template <> struct WholeCfgW<SysDubinsRel, HJ_SCHEME_ENO2> { using type = TmaCfg<8, HJ_INT_MINB, 1, HJ_INT_TY, 16>; };
template <> struct WholeCfgW<SysFlock, HJ_SCHEME_ENO3A> { using type = TmaCfg<8, HJ_INT_MINB, 1, HJ_INT_TY, 16>; };
template <> struct WholeCfgW<SysFlock, HJ_SCHEME_ENO2> { using type = TmaCfg<8, HJ_INT_MINB, 1, HJ_INT_TY, 16>; };
template <class Sys> struct SplitCfg;
// tile shapes of the 6-D pair; the -D overrides are a developer hook (HJ_EXTRA_NVCC_FLAGS in build.py).
//   pass 1: 41 x 41 planes as two 42 x 21 tiles (14 consumer warps + two ghost warps on alternate planes, 1 CTA/SM,
//           94 % of the lanes on nodes, slot rows padded to a conflict-free 58 doubles; measured on 8 x 41^5: 5.8 ms
//           against 9.9 ms for the 42 x 12 tile without ghost warps, which ran every plane in the general body, 7.9 ms
//           with one ghost warp -- the fill of one plane is a latency chain of about a plane time);
//   pass 2: 8 vector pairs x (4 x 7) tile, 7 consumer warps + the ghost warp, 2 CTAs/SM.
#ifndef HJ_P2_R
#define HJ_P2_R 6
#define HJ_P2_MINB 2
#define HJ_P2_VP 8
#define HJ_P2_TA 4
#define HJ_P2_TB 7
#define HJ_P2_GW 1
#define HJ_P2_OPT 0
#endif
#ifndef HJ_P1_TY
#define HJ_P1_MINB 1
#define HJ_P1_TY 21
#define HJ_P1_TXP 21
#define HJ_P1_GW 2
#define HJ_P1_XPAD 8
#define HJ_P1_R 8
#define HJ_P1_OPT 143
#endif
#define HJ_P2_6D VecCfg<3, HJ_P2_R, HJ_P2_MINB, HJ_P2_VP, HJ_P2_TA, HJ_P2_TB, HJ_P2_GW, HJ_P2_OPT>
#define HJ_P1_6D TmaCfg<HJ_P1_R, HJ_P1_MINB, 1, HJ_P1_TY, HJ_P1_TXP, false, HJ_P1_OPT, HJ_P1_GW, HJ_P1_XPAD>
// P2T: pass-2 tile for THIN dim-0 extents (slabs of a multi-GPU job: 41 planes over 8 ranks are 5..6 planes each, of
// which a 4-row tile wastes a third); chosen per context by pick_thin() below.  6 x 4 with a ghost warp (224 threads)
// measured 6.9 / 7.5 / 7.7 ms per stage on a 6 x 41^5 slab against 8.3 / 9.0 / 10.7 ms for 6 x 5 without one (which
// would need 288 threads and spill at 112 registers with one)
#ifndef HJ_P2T_TB
#define HJ_P2T_TB 4
#define HJ_P2T_GW 1
#endif
#define HJ_P2T_6D VecCfg<3, HJ_P2_R, HJ_P2_MINB, HJ_P2_VP, 6, HJ_P2T_TB, HJ_P2T_GW>
template <> struct SplitCfg<SysDubinsRelPair> { using P1 = HJ_P1_6D; using P2 = HJ_P2_6D; using P2T = HJ_P2T_6D; };
#ifndef HJ_P1_4D_GW
#define HJ_P1_4D_GW 1
#define HJ_P2_4D_GW 1
#endif
// 4-D pair: both kernels with a ghost warp (288 threads, 2 CTAs/SM; measured at 161^4: pass 1 3.07 ms against 3.80,
// pass 2 4.4 / 5.1 / 5.2 ms against 4.9 / 5.5 / 5.8, 23.9 ms per step against 27.5).  Pass 2 tiles dim 0 in 16 rows; slabs of 20-21 planes (161 over 8 ranks) take the 11-row tile (2 tiles = 22
// rows instead of 32)
template <> struct SplitCfg<SysDoubleIntPair> {
  using P1 = TmaCfg<8, 2, 1, 9, 27, false, 143, HJ_P1_4D_GW>;
  using P2 = VecCfg<2, 8, 2, 16, 16, 1, HJ_P2_4D_GW>;
  using P2T = VecCfg<2, 8, 2, 16, 11, 1, HJ_P2_4D_GW>;
};
// rows of dim 0 a tiling of height `ta` processes per useful row
static double tile_waste(int n0, int ta) { return (double)((n0 + ta - 1) / ta * ta) / n0; }
template <class P2, class P2T>
static bool pick_thin(int n0) {
  if (P2::TA == P2T::TA) return false;
  return tile_waste(n0, P2T::TA) < 0.9 * tile_waste(n0, P2::TA);
}

// ------------------------------------------------------------------------------------------ host side
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute of a kernel: remember per device
// whether this instantiation has it yet (a process may hold contexts on several devices; bit d = device d done)
template <class K>
static cudaError_t ensure_smem_attr(K kern, size_t smem, std::atomic<unsigned long long>& done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
  return e;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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

template <class Sys, int GD, int WENO, bool RED, int STAGE, class Cfg>
static cudaError_t launch_one(const HjTmaPlan* p, const CUtensorMap& tm, const KGrid& g, const KSys& ks,
                              const KStage& st, cudaStream_t s) {
  auto kern = k_stage_tma<Sys, GD, WENO, RED, STAGE, Cfg>;
  constexpr size_t smem = Cfg::template smem_bytes<STAGE>() + Sys::NSCRATCH * 8;
  static std::atomic<unsigned long long> attr_done{0};
  if (cudaError_t e = ensure_smem_attr(kern, smem, attr_done); e != cudaSuccess) return e;
  kern<<<(unsigned)p->nblocks, Cfg::NTHREADS, smem, s>>>(tm, p->tmap_y0, g, ks, st, p->geo);
  return cudaGetLastError();
}

template <class Blk, int GD, int WENO, bool RED, int STAGE, class Cfg>
static cudaError_t launch_vec(const HjTmaPlan* p, const CUtensorMap& tm, const KGrid& g, const KSys& ks,
                              const KStage& st, cudaStream_t s) {
  auto kern = k_stage_vec<Blk, GD, WENO, RED, STAGE, Cfg>;
  constexpr size_t smem = Cfg::smem_bytes();
  static std::atomic<unsigned long long> attr_done{0};
  if (cudaError_t e = ensure_smem_attr(kern, smem, attr_done); e != cudaSuccess) return e;
  kern<<<(unsigned)p->vblocks, Cfg::NTHREADS, smem, s>>>(tm, g, ks, st, p->vgeo);
  return cudaGetLastError();
}

struct TmaLauncher {
  const HjTmaPlan* p;
  int in_buf;
  int weno;
  const KGrid& g;
  const KSys& ks;
  const KStage& st;
  cudaStream_t s;
  int which_pass = 0;            // product systems: 0 = both passes, 1 = pass 1 only, 2 = pass 2 only
  cudaError_t err = cudaSuccess;
  int launches = 0;
  // whole system: one fused kernel per stage
  template <class Sys, int WENO, bool RED>
  cudaError_t by_stage() {
    const CUtensorMap& tm = p->tmap[in_buf];
    launches = 1;
    switch (st.stage) {
      case 1: return launch_one<Sys, Sys::BASE_DIM + Sys::ND, WENO, RED, 1, typename WholeCfgW<Sys, WENO>::type>(p, tm, g, ks, st, s);
      case 2: return launch_one<Sys, Sys::BASE_DIM + Sys::ND, WENO, RED, 2, typename WholeCfgW<Sys, WENO>::type>(p, tm, g, ks, st, s);
      case 3: return launch_one<Sys, Sys::BASE_DIM + Sys::ND, WENO, RED, 3, typename WholeCfgW<Sys, WENO>::type>(p, tm, g, ks, st, s);
      default: return cudaErrorNotSupported;
    }
  }
  // product system: pass 1 (trailing block, tmp = F_B(in)) then pass 2 (leading block + stage algebra)
  template <class Sys, int WENO, bool RED>
  cudaError_t split_by_stage() {
    using P1 = typename SplitCfg<Sys>::P1;
    using P2 = typename SplitCfg<Sys>::P2;
    KStage s1 = st;
    s1.stage = 0;                 // ydot of the trailing block only
    s1.comp = HJ_COMP_NONE;
    s1.use_obs = 0;
    s1.restrict_sign = 0;         // termRestrictUpdate acts on the total ydot, in pass 2
    s1.out = const_cast<double*>(st.tmp);
    if (which_pass != 2) {
      cudaError_t e = launch_one<typename Sys::Second, Sys::ND, WENO, RED, 0, P1>(p, p->tmap[in_buf], g, ks, s1, s);
      if (e != cudaSuccess) return e;
      ++launches;
      if (which_pass == 1) return cudaSuccess;
    }
    ++launches;
    const CUtensorMap& vm = p->vmap[in_buf];
    using P2T = typename SplitCfg<Sys>::P2T;
    if constexpr (!std::is_same<P2, P2T>::value) {
      if (p->thin) switch (st.stage) {
        case 1: return launch_vec<typename Sys::First, Sys::ND, WENO, RED, 1, P2T>(p, vm, g, ks, st, s);
        case 2: return launch_vec<typename Sys::First, Sys::ND, WENO, RED, 2, P2T>(p, vm, g, ks, st, s);
        case 3: return launch_vec<typename Sys::First, Sys::ND, WENO, RED, 3, P2T>(p, vm, g, ks, st, s);
        default: return cudaErrorNotSupported;
      }
    }
    switch (st.stage) {
      case 1: return launch_vec<typename Sys::First, Sys::ND, WENO, RED, 1, P2>(p, vm, g, ks, st, s);
      case 2: return launch_vec<typename Sys::First, Sys::ND, WENO, RED, 2, P2>(p, vm, g, ks, st, s);
      case 3: return launch_vec<typename Sys::First, Sys::ND, WENO, RED, 3, P2>(p, vm, g, ks, st, s);
      default: return cudaErrorNotSupported;
    }
  }
  template <class Sys>
  void operator()() {
    const bool red = st.want_reduce != 0;
    if constexpr (SysSplit<Sys>::value) {
      if (weno == HJ_WENO_AS_SHIPPED) err = red ? split_by_stage<Sys, HJ_WENO_AS_SHIPPED, true>() : split_by_stage<Sys, HJ_WENO_AS_SHIPPED, false>();
      else err = red ? split_by_stage<Sys, HJ_WENO_INTENDED, true>() : split_by_stage<Sys, HJ_WENO_INTENDED, false>();
    } else if constexpr (Sys::ND >= 3) {
      if (weno == HJ_WENO_AS_SHIPPED) err = red ? by_stage<Sys, HJ_WENO_AS_SHIPPED, true>() : by_stage<Sys, HJ_WENO_AS_SHIPPED, false>();
      else if (weno == HJ_WENO_INTENDED) err = red ? by_stage<Sys, HJ_WENO_INTENDED, true>() : by_stage<Sys, HJ_WENO_INTENDED, false>();
      else if constexpr (Sys::NSCRATCH == 0) {           // ENO functors: whole 3-D systems (not the batch functor)
        if (weno == HJ_SCHEME_ENO3A) err = red ? by_stage<Sys, HJ_SCHEME_ENO3A, true>() : by_stage<Sys, HJ_SCHEME_ENO3A, false>();
        else if (weno == HJ_SCHEME_ENO2) err = red ? by_stage<Sys, HJ_SCHEME_ENO2, true>() : by_stage<Sys, HJ_SCHEME_ENO2, false>();
        else err = cudaErrorNotSupported;
      } else err = cudaErrorNotSupported;
    } else {
      err = cudaErrorNotSupported;
    }
  }
};

// tile shapes of a system's kernels, for the plan
struct PlanShape {
  int txp = ProdCfg::TXP, ty = ProdCfg::TY, bw = ProdCfg::BW;
  bool split = false, thin = false;
  int n0 = 0, weno = HJ_WENO_AS_SHIPPED;
  int ns = 0, vb = 0, ta = 0, tb = 0;
  template <class Sys>
  void operator()() {
    if constexpr (SysSplit<Sys>::value) {
      using P1 = typename SplitCfg<Sys>::P1;
      using P2 = typename SplitCfg<Sys>::P2;
      using P2T = typename SplitCfg<Sys>::P2T;
      txp = P1::TXP; ty = P1::TY; bw = P1::BW;
      split = true;
      thin = pick_thin<P2, P2T>(n0);
      ns = P2::NS; vb = P2::VB; ta = thin ? P2T::TA : P2::TA; tb = thin ? P2T::TB : P2::TB;
    } else if (weno == HJ_SCHEME_ENO3A || weno == HJ_SCHEME_ENO2) {
      using W = typename WholeCfgW<Sys, HJ_SCHEME_ENO3A>::type;
      txp = W::TXP; ty = W::TY; bw = W::BW;
    } else {
      using W = typename WholeCfg<Sys>::type;
      txp = W::TXP; ty = W::TY; bw = W::BW;
    }
  }
};

HjTmaPlan* hj_tma_plan_create(const KGrid& g, int system_id, int weno, double* const bufs[3], int halo0, char* err,
                              int errlen, int tile_y) {
  const int D = g.D;
  PlanShape shape;
  shape.n0 = g.N[0];
  shape.weno = weno;
  if (!hj_dispatch_system(system_id, shape)) { snprintf(err, errlen, "unknown system"); return nullptr; }
  const int TY = tile_y > 0 ? tile_y : shape.ty;
  const int TX = 2 * shape.txp;
  const int BWD = shape.bw;                            // slot row = TX + 8 (+ XPAD) doubles
  if (D < 3) { snprintf(err, errlen, "2-D grids use the gather backend"); return nullptr; }
  if (hj_system_ndim(system_id) != D) { snprintf(err, errlen, "system/grid dim mismatch"); return nullptr; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) { snprintf(err, errlen, "cuTensorMapEncodeTiled not available from the driver"); return nullptr; }
  const int NX = g.N[D - 1], NY = g.N[D - 2], NZ = g.N[D - 3];
  // the tiled / marched dims of the plane-ring kernel (outer dims -- a batch index, the leading dims of a >= 4-D grid --
  // carry no such limit)
  if (NX < 4 || NY < 4 || NZ < 4) { snprintf(err, errlen, "X/Y/Z extent too small for the plane-ring backend"); return nullptr; }
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
  p->geo.zbeg = 0;
  p->geo.zend = NZ;
  p->tiles = tiles;
  p->nblocks = tiles * p->geo.nzc;
  if (p->nblocks > 0x7fffffffLL) { delete p; snprintf(err, errlen, "grid too large"); return nullptr; }
  for (int bidx = 0; bidx < 4; ++bidx) {                    // 0..2: haloed boxes on the RK buffers; 3: y0 tile on buffer 0
    cuuint64_t dims[3] = {(cuuint64_t)NX, (cuuint64_t)NY, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 8, (cuuint64_t)pitch * NY * 8};
    cuuint32_t box[3] = {(cuuint32_t)(bidx < 3 ? BWD : TX), (cuuint32_t)(bidx < 3 ? TY + 6 : TY), 1};
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
  p->split = shape.split;
  p->thin = shape.thin;
  if (shape.split) {
    // pass 2: [V, (N2,) N1, N0 (+ halo planes)] with V = the flattened trailing dims; the block's last dim is marched
    // (box extent 1), the others are tiled with a 3-cell halo
    const int NS = shape.ns, MD = NS - 1;
    const long long V = g.stride[NS - 1];
    const long long planes0 = g.N[0] + (halo0 ? 2 * HJ_GHOST : 0);
    VecGeom& vg = p->vgeo;
    vg.nvc = (int)((V + shape.vb - 1) / shape.vb);
    vg.vc0 = 0;
    p->vb = shape.vb;
    p->vlen = V;
    vg.nta = (g.N[0] + shape.ta - 1) / shape.ta;
    vg.ntb = NS == 3 ? (g.N[1] + shape.tb - 1) / shape.tb : 1;
    vg.cz = g.N[MD];
    vg.nzc = 1;
    vg.zcoord0 = halo0 ? HJ_GHOST : 0;
    vg.pitch = (int)pitch;
    vg.NX = NX;
    p->vblocks = (long long)vg.nvc * vg.nta * vg.ntb * vg.nzc;
    if (g.N[MD] < 4 || V > 0x7fffffffLL || p->vblocks > 0x7fffffffLL) {
      delete p;
      snprintf(err, errlen, "grid shape not supported by the dimension-split path");
      return nullptr;
    }
    for (int bidx = 0; bidx < 3; ++bidx) {
      cuuint64_t dims[4];
      cuuint64_t strides[3];
      cuuint32_t box[4];
      cuuint32_t es[4] = {1, 1, 1, 1};
      int rank;
      if (NS == 3) {
        rank = 4;
        dims[0] = (cuuint64_t)V; dims[1] = (cuuint64_t)g.N[2]; dims[2] = (cuuint64_t)g.N[1]; dims[3] = (cuuint64_t)planes0;
        strides[0] = (cuuint64_t)g.stride[2] * 8; strides[1] = (cuuint64_t)g.stride[1] * 8; strides[2] = (cuuint64_t)g.stride[0] * 8;
        box[0] = shape.vb; box[1] = 1; box[2] = shape.tb + 6; box[3] = shape.ta + 6;
      } else {
        rank = 3;
        dims[0] = (cuuint64_t)V; dims[1] = (cuuint64_t)g.N[1]; dims[2] = (cuuint64_t)planes0;
        strides[0] = (cuuint64_t)g.stride[1] * 8; strides[1] = (cuuint64_t)g.stride[0] * 8;
        box[0] = shape.vb; box[1] = 1; box[2] = shape.ta + 6;
      }
      CUresult r = enc(&p->vmap[bidx], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, (void*)bufs[bidx], dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        snprintf(err, errlen, "cuTensorMapEncodeTiled (pass 2) failed (%d)", (int)r);
        delete p;
        return nullptr;
      }
    }
  }
  return p;
}

void hj_tma_plan_destroy(HjTmaPlan* p) { delete p; }
bool hj_tma_plan_is_split(const HjTmaPlan* p) { return p && p->split; }

bool hj_tma_plan_cols(const HjTmaPlan* p, long long* vlen, int* quantum) {
  if (!p || !p->split) return false;
  if (vlen) *vlen = p->vlen;
  if (quantum) *quantum = p->vb;
  return true;
}

cudaError_t hj_launch_stage_tma(HjTmaPlan* plan, int system_id, int weno, const KGrid& g, const KSys& ks,
                                const KStage& st, int in_buf, cudaStream_t s, int zbeg, int zend, int which_pass,
                                long long col_begin, long long col_end) {
  HjTmaPlan sub;
  if (col_end > col_begin) {       // pass 2 on columns [col_begin, col_end) of the vector axis only
    if (!plan->split || which_pass != 2 || col_begin % plan->vb) return cudaErrorInvalidValue;
    sub = *plan;
    sub.vgeo.vc0 = (int)(col_begin / plan->vb);
    const long long c1 = (col_end + plan->vb - 1) / plan->vb;
    sub.vblocks = plan->vblocks / plan->vgeo.nvc * (c1 - sub.vgeo.vc0);
    plan = &sub;
  }
  if (zend > zbeg) {               // advance only planes [zbeg, zend) of Z (pipelined host <-> device stepping)
    if (plan->split) return cudaErrorNotSupported;
    sub = *plan;
    sub.geo.zbeg = zbeg;
    sub.geo.zend = zend;
    sub.geo.nzc = (zend - zbeg + sub.geo.cz - 1) / sub.geo.cz;
    sub.nblocks = sub.tiles * sub.geo.nzc;
    plan = &sub;
  }
  TmaLauncher l{plan, in_buf, weno, g, ks, st, s};
  if (which_pass) {
    if (!plan->split) return cudaErrorNotSupported;
    l.which_pass = which_pass;
  }
  if (!hj_dispatch_system(system_id, l)) return cudaErrorInvalidValue;
  hj_count_launch(l.launches);
  return l.err;
}
