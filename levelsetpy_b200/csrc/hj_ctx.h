// hj_ctx.h -- the context behind the opaque hj_ctx* of include/hjb200.h, shared by the C-ABI translation units
// (hj_api.cu: fields, operators, stepping; hj_halo.cu: slab halos over peer memory).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "hj_internal.h"

struct HjHalo;   // peer-memory halo transport of a slab context (hj_halo.cu)

struct hj_ctx {
  int device = 0, D = 0, weno = 0, backend = HJ_BACKEND_AUTO, system_id = HJ_SYS_NONE, nparams = 0;
  int halo0 = 0;                 // dim 0 carries stored halo planes (slab decomposition)
  long long pitch = 0;           // padded innermost extent
  long long plane = 0;           // pitched elements of one dim-0 plane
  long long elems = 0;           // pitched elements of a whole field incl. halo planes
  long long origin = 0;          // element offset of the first interior node
  long long nodes = 0;           // prod N
  KGrid gp{}, gd{};              // pitched (resident fields) and dense (user arrays) views
  KSys ks{};
  double* vs_dev[HJ_MAX_DIM] = {};
  double* tab_dev[HJ_MAX_TABLES] = {};
  unsigned axes_set = 0;
  double* buf[3] = {};           // y, y1, yHalf (base pointers incl. halo planes)
  double* aux = nullptr;
  double* obs = nullptr;
  double* snap = nullptr;        // the state at the last hj_snapshot (driver: the frame at tau[i-1])
  double* staging = nullptr;     // dense staging for host <-> pitched conversion
  unsigned long long* red = nullptr;  // 4 reduction records (3 stages + scratch) + eps record
  unsigned long long* eps = nullptr;
  double* pinned = nullptr;      // host scratch
  bool have_state = false, alpha_valid = false;
  double alpha_cache[HJ_MAX_DIM] = {};
  double step_bound_cache = 0.0;
  int restrict_sign = 0;         // termRestrictUpdate: 0 off, +1 / -1
  int nbatch = 0;                // > 0: batch context (dim 0 of the internal grid is the batch index)
  double* batch_dt = nullptr;    // [nbatch] per-element dt of the current step
  double* batch_params = nullptr;// [3][nbatch][nparams] per-stage parameter blocks
  // pipelined host <-> device stepping (hj_ode_cfl3_single with host buffers)
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  std::vector<cudaEvent_t> ev_up, ev_done;
  cudaEvent_t ev_start = nullptr;
  int pipe_planes = 0;           // dim-0 planes per chunk of the pipelined step (0: the default policy)
  HjTmaPlan* plan = nullptr;
  bool plan_tried = false;
  std::string plan_err;
  HjHalo* halo = nullptr;         // peer-memory halo transport (hj_halo_attach)
};

void hj_halo_fused_targets(hj_ctx* c, int out_buf, double** lo, double** hi);   // hj_halo.cu
void hj_halo_destroy(hj_ctx* c);   // hj_halo.cu: unmap neighbours, free the flag array (called by hj_destroy)

// sets the thread's hj_last_error() text and returns `code`
int hj_fail(int code, const char* fmt, ...);
#define HJ_CK(call)                                                                                       \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) return hj_fail(HJ_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));   \
  } while (0)
