// hj_tma.cu -- TMA backend, host side: tensor maps, tile geometry and the per-system launchers of the plane-ring
// kernel (device side: hj_tma_kernel.cuh).
#include <cuda.h>

#include <cstdio>
#include <cstring>

#include "hj_internal.h"
#include "hj_tma_kernel.cuh"
#include "hj_tma_plan.h"

using namespace hjtma;

// production configuration of the ring kernel (see tools/tune_tma.cu for the measured alternatives)
using ProdCfg = TmaCfg<8, 2, 1>;

// ------------------------------------------------------------------------------------------ host side
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

template <class Sys, int WENO, bool RED, int STAGE, class Cfg>
static cudaError_t launch_one(const HjTmaPlan* p, const CUtensorMap& tm, const KGrid& g, const KSys& ks,
                              const KStage& st, cudaStream_t s) {
  auto kern = k_stage_tma<Sys, WENO, RED, STAGE, Cfg>;
  constexpr size_t smem = Cfg::template smem_bytes<STAGE>();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  kern<<<(unsigned)p->nblocks, Cfg::NTHREADS, smem, s>>>(tm, p->tmap_y0, g, ks, st, p->geo);
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
  template <class Sys, int WENO, bool RED>
  cudaError_t by_stage() {
    switch (st.stage) {
      case 1: return launch_one<Sys, WENO, RED, 1, ProdCfg>(p, tm, g, ks, st, s);
      case 2: return launch_one<Sys, WENO, RED, 2, ProdCfg>(p, tm, g, ks, st, s);
      case 3: return launch_one<Sys, WENO, RED, 3, ProdCfg>(p, tm, g, ks, st, s);
      default: return cudaErrorNotSupported;
    }
  }
  template <class Sys>
  void operator()() {
    if constexpr (Sys::ND >= 3) {
      const bool red = st.want_reduce != 0;
      if (weno == HJ_WENO_AS_SHIPPED) err = red ? by_stage<Sys, HJ_WENO_AS_SHIPPED, true>() : by_stage<Sys, HJ_WENO_AS_SHIPPED, false>();
      else err = red ? by_stage<Sys, HJ_WENO_INTENDED, true>() : by_stage<Sys, HJ_WENO_INTENDED, false>();
    } else {
      err = cudaErrorNotSupported;
    }
  }
};

HjTmaPlan* hj_tma_plan_create(const KGrid& g, int system_id, int weno, double* const bufs[3], int halo0, char* err,
                              int errlen, int tile_y) {
  (void)weno;
  const int D = g.D;
  const int TY = tile_y > 0 ? tile_y : ProdCfg::TY;
  if (D < 3) { snprintf(err, errlen, "2-D grids use the gather backend"); return nullptr; }
  if (hj_system_ndim(system_id) != D) { snprintf(err, errlen, "system/grid dim mismatch"); return nullptr; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) { snprintf(err, errlen, "cuTensorMapEncodeTiled not available from the driver"); return nullptr; }
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
