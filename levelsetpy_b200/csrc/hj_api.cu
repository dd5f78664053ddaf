// hj_api.cu -- the C-ABI (include/hjb200.h): context, resident fields, operator entry points.
#include <atomic>
#include <cstdarg>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "hj_ctx.h"
#include "hj_internal.h"
#include "hj_systems.cuh"

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};
void hj_count_launch(int n) { g_launches += n; }

int hj_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define fail hj_fail
#define CK HJ_CK

static const int RED_STRIDE = 32;  // slots per reduction record (>= HJ_REDUCE_LEN(6) = 19)

extern "C" {

static int ensure_buffers(hj_ctx* c);

const char* hj_version(void) { return "levelsetpy_b200 0.1 (sm_100a)"; }
const char* hj_last_error(void) { return g_err.c_str(); }
int64_t hj_launch_count(void) { return (int64_t)g_launches.load(); }

int hj_create(hj_ctx** out, int device, int ndim, const int64_t* N, const double* dx, const int* bc_kind,
              const int* bc_toward_zero, int weno_mode) {
  if (!out || !N || !dx || !bc_kind) return fail(HJ_ERR_INVALID, "hj_create: null argument");
  if (ndim < 2 || ndim > HJ_MAX_DIM) return fail(HJ_ERR_UNSUPPORTED, "hj_create: grid.dim must be 2..%d, got %d", HJ_MAX_DIM, ndim);
  if (weno_mode < HJ_WENO_AS_SHIPPED || weno_mode > HJ_SCHEME_ENO2) return fail(HJ_ERR_INVALID, "hj_create: bad weno_mode");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(HJ_ERR_CUDA, "hj_create: no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(HJ_ERR_INVALID, "hj_create: device %d out of range", device);
  CK(cudaSetDevice(device));
  hj_ctx* c = new hj_ctx();
  c->device = device;
  c->D = ndim;
  c->weno = weno_mode;
  for (int d = 0; d < ndim; ++d) {
    // smallest extents the reference's ghost cells accept: addGhostExtrapolate reads the edge node and its neighbour
    // (add_ghost_extrapolate.py:88-100), addGhostPeriodic copies `width` = 3 nodes (add_ghost_periodic.py:78-87)
    const int64_t nmin = bc_kind[d] == HJ_BC_EXTRAPOLATE ? 2 : HJ_GHOST;
    if (N[d] < nmin || N[d] > 0x7fffffff) { delete c; return fail(HJ_ERR_INVALID, "hj_create: N[%d]=%lld unsupported (need >= %lld)", d, (long long)N[d], (long long)nmin); }
    if (!(dx[d] > 0.0)) { delete c; return fail(HJ_ERR_INVALID, "hj_create: grid cell size dx must be strictly positive"); }
    if (bc_kind[d] != HJ_BC_EXTRAPOLATE && bc_kind[d] != HJ_BC_PERIODIC && !(bc_kind[d] == HJ_BC_HALO && d == 0)) {
      delete c;
      return fail(HJ_ERR_UNSUPPORTED, "hj_create: boundary kind %d on dim %d unsupported", bc_kind[d], d);
    }
  }
  c->halo0 = bc_kind[0] == HJ_BC_HALO;
  const int D = ndim;
  c->pitch = (N[D - 1] + 1) & ~1LL;
  KGrid& gp = c->gp;
  KGrid& gd = c->gd;
  gp.D = gd.D = D;
  long long sp = 1, sd = 1;
  c->nodes = 1;
  for (int d = D - 1; d >= 0; --d) {
    gp.N[d] = gd.N[d] = (int)N[d];
    gp.dx[d] = gd.dx[d] = dx[d];
    gp.dxinv[d] = gd.dxinv[d] = 1 / dx[d];
    gp.bc[d] = gd.bc[d] = bc_kind[d];
    gp.slope_mult[d] = gd.slope_mult[d] = (bc_toward_zero && bc_toward_zero[d]) ? -1.0 : 1.0;
    gp.ca1[d] = gd.ca1[d] = gp.dxinv[d] * (45.0 / 60.0);
    gp.ca2[d] = gd.ca2[d] = gp.dxinv[d] * (-9.0 / 60.0);
    gp.ca3[d] = gd.ca3[d] = gp.dxinv[d] * (1.0 / 60.0);
    gp.cb[d] = gd.cb[d] = gp.dxinv[d] * (1.0 / 60.0);
    gp.stride[d] = sp;
    gd.stride[d] = sd;
    sp *= (d == D - 1) ? c->pitch : N[d];
    sd *= N[d];
    c->nodes *= N[d];
  }
  c->plane = gp.stride[0];
  c->origin = c->halo0 ? HJ_GHOST * c->plane : 0;
  c->elems = c->plane * (N[0] + (c->halo0 ? 2 * HJ_GHOST : 0));
  cudaError_t e = cudaMalloc(&c->red, 5 * RED_STRIDE * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(c->red, 0, 5 * RED_STRIDE * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMallocHost(&c->pinned, 5 * RED_STRIDE * sizeof(double));
  if (e != cudaSuccess) {
    hj_destroy(c);
    return fail(HJ_ERR_CUDA, "hj_create: allocation failed: %s", cudaGetErrorString(e));
  }
  c->eps = c->red + 4 * RED_STRIDE;
  *out = c;
  return HJ_OK;
}

int hj_destroy(hj_ctx* c) {
  if (!c) return HJ_OK;
  cudaSetDevice(c->device);
  hj_halo_destroy(c);
  if (c->plan) hj_tma_plan_destroy(c->plan);
  for (cudaEvent_t e : c->ev_up) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_done) cudaEventDestroy(e);
  if (c->ev_start) cudaEventDestroy(c->ev_start);
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  for (int b = 0; b < 3; ++b) cudaFree(c->buf[b]);
  for (int d = 0; d < HJ_MAX_DIM; ++d) cudaFree(c->vs_dev[d]);
  for (int t = 0; t < HJ_MAX_TABLES; ++t) cudaFree(c->tab_dev[t]);
  cudaFree(c->batch_dt);
  cudaFree(c->batch_params);
  cudaFree(c->aux);
  cudaFree(c->obs);
  cudaFree(c->snap);
  cudaFree(c->staging);
  cudaFree(c->red);
  cudaFreeHost(c->pinned);
  delete c;
  return HJ_OK;
}

int hj_set_backend(hj_ctx* c, int backend) {
  if (!c) return fail(HJ_ERR_INVALID, "null ctx");
  if (backend < HJ_BACKEND_AUTO || backend > HJ_BACKEND_TMA) return fail(HJ_ERR_INVALID, "hj_set_backend: bad backend %d", backend);
  c->backend = backend;
  return HJ_OK;
}

static int upload_vec(double** dst, const double* host, int64_t n) {
  if (*dst) { cudaFree(*dst); *dst = nullptr; }
  CK(cudaMalloc(dst, n * sizeof(double)));
  CK(cudaMemcpy(*dst, host, n * sizeof(double), cudaMemcpyHostToDevice));
  return HJ_OK;
}

int hj_set_axis(hj_ctx* c, int dim, const double* vs_host, int64_t n) {
  if (!c || !vs_host) return fail(HJ_ERR_INVALID, "hj_set_axis: null argument");
  if (c->nbatch) dim += 1;       // user dims of a batch context sit behind the batch dim
  if (dim < (c->nbatch ? 1 : 0) || dim >= c->D) return fail(HJ_ERR_INVALID, "hj_set_axis: Illegal dim parameter");
  if (n != c->gp.N[dim]) return fail(HJ_ERR_INVALID, "hj_set_axis: vs[%d] has %lld entries, grid.N is %d", dim, (long long)n, c->gp.N[dim]);
  CK(cudaSetDevice(c->device));
  int r = upload_vec(&c->vs_dev[dim], vs_host, n);
  if (r) return r;
  c->gp.vs[dim] = c->gd.vs[dim] = c->vs_dev[dim];
  c->axes_set |= 1u << dim;
  c->alpha_valid = false;
  return HJ_OK;
}

int hj_set_table(hj_ctx* c, int slot, const double* tab_host, int64_t n) {
  if (!c || !tab_host) return fail(HJ_ERR_INVALID, "hj_set_table: null argument");
  if (slot < 0 || slot >= HJ_MAX_TABLES || n <= 0) return fail(HJ_ERR_INVALID, "hj_set_table: bad slot/size");
  CK(cudaSetDevice(c->device));
  int r = upload_vec(&c->tab_dev[slot], tab_host, n);
  if (r) return r;
  c->ks.tab[slot] = c->tab_dev[slot];
  c->alpha_valid = false;
  return HJ_OK;
}

int hj_set_system(hj_ctx* c, int system_id, const double* params, int nparams) {
  if (!c) return fail(HJ_ERR_INVALID, "null ctx");
  if (c->nbatch) {
    if (system_id != HJ_SYS_FLOCK)
      return fail(HJ_ERR_UNSUPPORTED, "hj_set_system: only HJ_SYS_FLOCK has a batch functor (got system id %d)", system_id);
    system_id = HJ_SYS_FLOCK_BATCH;
  }
  const int nd = hj_system_ndim(system_id);
  if (nd < 0) return fail(HJ_ERR_UNSUPPORTED, "hj_set_system: system id %d has no registered device functor", system_id);
  if (nd != c->D) return fail(HJ_ERR_INVALID, "hj_set_system: system is %d-D but the grid is %d-D", nd - (c->nbatch ? 1 : 0), c->D - (c->nbatch ? 1 : 0));
  if (nparams < 0 || nparams > HJ_MAX_PARAMS || (nparams && !params && !c->nbatch)) return fail(HJ_ERR_INVALID, "hj_set_system: bad parameter block");
  if (system_id == HJ_SYS_FLOCK_BATCH) {
    if (nparams < HJ_FLOCK_HDR || nparams > HJ_MAX_PARAMS) return fail(HJ_ERR_INVALID, "hj_set_system: batch flock block length must be %d..%d doubles", HJ_FLOCK_HDR, HJ_MAX_PARAMS);
    params = nullptr;            // per-element blocks arrive with hj_step_batch
  } else if (system_id == HJ_SYS_FLOCK) {
    if (nparams < HJ_FLOCK_HDR) return fail(HJ_ERR_INVALID, "hj_set_system: flock block needs >= %d doubles", HJ_FLOCK_HDR);
    const int K = (int)params[0];
    if (K < 0 || HJ_FLOCK_HDR + 3 * K > nparams) return fail(HJ_ERR_INVALID, "hj_set_system: flock block too short for %d birds", K);
  }
  c->system_id = system_id;
  c->nparams = nparams;
  std::memset(c->ks.p, 0, sizeof c->ks.p);
  if (nparams && params) std::memcpy(c->ks.p, params, nparams * sizeof(double));
  c->alpha_valid = false;
  return HJ_OK;
}

int64_t hj_num_nodes(const hj_ctx* c) { return c ? c->nodes : 0; }
int64_t hj_field_elems(const hj_ctx* c) { return c ? c->elems : 0; }
int64_t hj_plane_elems(const hj_ctx* c) { return c ? c->plane : 0; }

int hj_state_ptr(hj_ctx* c, int which, double** p) {
  if (!c || !p || which < 0 || which > 2) return fail(HJ_ERR_INVALID, "hj_state_ptr: bad argument");
  CK(cudaSetDevice(c->device));
  int r = ensure_buffers(c);
  if (r) return r;
  *p = c->buf[which];
  if (which == 0) c->have_state = true;   // the caller now owns the contents of the resident state
  return HJ_OK;
}

// the three RK buffers are allocated on first use so that operator-only contexts (hj_deriv, hj_add_ghost,
// hj_rhs on dense arrays) stay light
static int ensure_buffers(hj_ctx* c) {
  for (int b = 0; b < 3; ++b) {
    if (c->buf[b]) continue;
    cudaError_t e = cudaMalloc(&c->buf[b], c->elems * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(c->buf[b], 0, c->elems * sizeof(double));
    if (e != cudaSuccess)
      return fail(HJ_ERR_CUDA, "allocation of RK buffer %d (%.2f GB) failed: %s", b, c->elems * 8e-9, cudaGetErrorString(e));
  }
  return HJ_OK;
}

static double** field_slot(hj_ctx* c, int field) {
  switch (field) {
    case HJ_FIELD_STATE: return &c->buf[0];
    case HJ_FIELD_AUX: return &c->aux;
    case HJ_FIELD_OBSTACLE: return &c->obs;
    default: return nullptr;
  }
}

static int need_staging(hj_ctx* c) {
  if (!c->staging) CK(cudaMalloc(&c->staging, c->nodes * sizeof(double)));
  return HJ_OK;
}

int hj_upload(hj_ctx* c, void* stream, int field, const double* dense, int is_host) {
  if (!c || !dense) return fail(HJ_ERR_INVALID, "hj_upload: null argument");
  double** slot = field_slot(c, field);
  if (!slot) return fail(HJ_ERR_INVALID, "hj_upload: bad field %d", field);
  CK(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (field == HJ_FIELD_STATE) {
    int r = ensure_buffers(c);
    if (r) return r;
  }
  if (!*slot) {
    CK(cudaMalloc(slot, c->elems * sizeof(double)));
    CK(cudaMemsetAsync(*slot, 0, c->elems * sizeof(double), s));
  }
  double* dst = *slot + c->origin;
  const bool same_layout = c->pitch == c->gp.N[c->D - 1];
  if (same_layout) {
    CK(cudaMemcpyAsync(dst, dense, c->nodes * sizeof(double), is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
  } else {
    const double* src = dense;
    if (is_host) {
      int r = need_staging(c);
      if (r) return r;
      CK(cudaMemcpyAsync(c->staging, dense, c->nodes * sizeof(double), cudaMemcpyHostToDevice, s));
      src = c->staging;
    }
    CK(hj_launch_pack(src, dst, c->gd, c->gp, s));
  }
  if (field == HJ_FIELD_STATE) c->have_state = true;
  return HJ_OK;
}

int hj_download(hj_ctx* c, void* stream, int field, double* dense, int is_host) {
  if (!c || !dense) return fail(HJ_ERR_INVALID, "hj_download: null argument");
  double** slot = field_slot(c, field);
  if (!slot || !*slot) return fail(HJ_ERR_STATE, "hj_download: field %d not resident", field);
  CK(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  const double* src = *slot + c->origin;
  const bool same_layout = c->pitch == c->gp.N[c->D - 1];
  if (same_layout) {
    CK(cudaMemcpyAsync(dense, src, c->nodes * sizeof(double), is_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s));
  } else if (is_host) {
    int r = need_staging(c);
    if (r) return r;
    CK(hj_launch_unpack(src, c->staging, c->gd, c->gp, s));
    CK(cudaMemcpyAsync(dense, c->staging, c->nodes * sizeof(double), cudaMemcpyDeviceToHost, s));
  } else {
    CK(hj_launch_unpack(src, dense, c->gd, c->gp, s));
  }
  if (is_host) CK(cudaStreamSynchronize(s));
  return HJ_OK;
}

static int check_ready(hj_ctx* c, bool need_system) {
  if (!c) return fail(HJ_ERR_INVALID, "null ctx");
  if (need_system) {
    if (c->system_id == HJ_SYS_NONE) return fail(HJ_ERR_STATE, "no system registered: call hj_set_system first");
    if (c->axes_set != (1u << c->D) - 1) return fail(HJ_ERR_STATE, "grid.vs not set for every dim: call hj_set_axis");
  }
  return HJ_OK;
}

int hj_deriv(hj_ctx* c, void* stream, const double* data_dev, int dim, double* dl, double* dr) {
  if (!c || !data_dev || !dl || !dr) return fail(HJ_ERR_INVALID, "hj_deriv: null argument");
  if (dim < 0 || dim >= c->D) return fail(HJ_ERR_INVALID, "Illegal dim parameter");   // upwind_first_weno5a.py:60
  if (c->halo0) return fail(HJ_ERR_UNSUPPORTED, "hj_deriv: not available on a slab (halo) context");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (c->weno == HJ_WENO_INTENDED) {
    CK(hj_launch_init_eps(c->eps, c->D, s));
    CK(hj_launch_maxd1sq(c->gd, data_dev, c->eps, dim, s));
  }
  CK(hj_launch_deriv(c->weno, c->gd, data_dev, dim, dl, dr, c->eps, s));
  return HJ_OK;
}

int hj_deriv_candidates(hj_ctx* c, void* stream, const double* data_dev, int dim, double* out6_dev) {
  if (!c || !data_dev || !out6_dev) return fail(HJ_ERR_INVALID, "hj_deriv_candidates: null argument");
  if (dim < 0 || dim >= c->D) return fail(HJ_ERR_INVALID, "Illegal dim parameter");   // ENO3aHelper.py:54-55
  if (c->halo0) return fail(HJ_ERR_UNSUPPORTED, "hj_deriv_candidates: not available on a slab (halo) context");
  CK(cudaSetDevice(c->device));
  CK(hj_launch_deriv_all(c->gd, data_dev, dim, out6_dev, (cudaStream_t)stream));
  return HJ_OK;
}

int hj_add_ghost(hj_ctx* c, void* stream, const double* data_dev, int dim, int width, double* out_dev) {
  if (!c || !data_dev || !out_dev) return fail(HJ_ERR_INVALID, "hj_add_ghost: null argument");
  if (dim < 0 || dim >= c->D) return fail(HJ_ERR_INVALID, "Illegal dim parameter");
  if (width < 0 || width > c->gd.N[dim]) return fail(HJ_ERR_INVALID, "Illegal width parameter");  // add_ghost_extrapolate.py:58
  if (c->gd.bc[dim] == HJ_BC_HALO) return fail(HJ_ERR_UNSUPPORTED, "hj_add_ghost: halo dim");
  CK(cudaSetDevice(c->device));
  CK(hj_launch_add_ghost(c->gd, data_dev, dim, width, out_dev, (cudaStream_t)stream));
  return HJ_OK;
}

static void decode_record(const unsigned long long* enc, int D, double* out) {
  for (int i = 0; i < 3 * D; ++i) {
    unsigned long long e = enc[i];
    unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
    std::memcpy(&out[i], &b, 8);
  }
  out[3 * D] = enc[3 * D] ? 1.0 : 0.0;
}

static double step_bound_from_alpha(const hj_ctx* c, const double* amax) {
  double inv = 0;
  for (int d = 0; d < c->D; ++d) inv = inv + (amax[d] / c->gp.dx[d]);   // artificial_diss_glf.py:107, dims in order
  return 1 / inv;                                                       // :109
}

int hj_rhs(hj_ctx* c, void* stream, double t, const double* y_dev, double* ydot_dev, double* step_bound,
           double* reduce_host) {
  (void)t;
  int r = check_ready(c, true);
  if (r) return r;
  if (!y_dev || !ydot_dev) return fail(HJ_ERR_INVALID, "hj_rhs: null argument");
  if (c->halo0 || c->nbatch) return fail(HJ_ERR_UNSUPPORTED, "hj_rhs: dense-array entry point is not available on a slab / batch context");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* red = c->red + 3 * RED_STRIDE;
  CK(hj_launch_init_reduce(red, c->D, s));
  if (c->weno == HJ_WENO_INTENDED) {
    CK(hj_launch_init_eps(c->eps, c->D, s));
    CK(hj_launch_maxd1sq(c->gd, y_dev, c->eps, -1, s));
  }
  KStage st{};
  st.stage = 0;
  st.want_reduce = 1;
  st.restrict_sign = c->restrict_sign;
  st.in = y_dev;
  st.out = ydot_dev;
  for (int d = 0; d < c->D; ++d) st.out_stride[d] = c->gd.stride[d];
  st.red = red;
  st.epsmax = c->eps;
  CK(hj_launch_stage_gather(c->system_id, c->weno, c->gd, c->ks, st, s));
  CK(cudaMemcpyAsync(c->pinned, red, RED_STRIDE * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  double rec[RED_STRIDE];
  decode_record((const unsigned long long*)c->pinned, c->D, rec);
  if (reduce_host) std::memcpy(reduce_host, rec, HJ_REDUCE_LEN(c->D) * sizeof(double));
  if (step_bound) *step_bound = step_bound_from_alpha(c, rec);
  return HJ_OK;
}

static bool use_tma(hj_ctx* c);

int hj_deriv_range(hj_ctx* c, void* stream, const double* y_dev, int stage, double* deriv_min, double* deriv_max) {
  if (!c || !deriv_min || !deriv_max) return fail(HJ_ERR_INVALID, "hj_deriv_range: null argument");
  if (c->nbatch) return fail(HJ_ERR_UNSUPPORTED, "hj_deriv_range: not available on a batch context");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  const double* in;
  const KGrid* g;
  if (y_dev) {
    if (c->halo0) return fail(HJ_ERR_UNSUPPORTED, "hj_deriv_range: dense arrays are not available on a slab context");
    in = y_dev;
    g = &c->gd;
  } else {
    if (stage < 1 || stage > 4) return fail(HJ_ERR_INVALID, "hj_deriv_range: stage must be 1..3 (odeCFL3) or 4 (final stage of odeCFL2)");
    if (!c->have_state) return fail(HJ_ERR_STATE, "hj_deriv_range: no resident state (hj_upload first)");
    static const int in_[5] = {0, 0, 1, 2, 1};
    in = c->buf[in_[stage]] + c->origin;
    g = &c->gp;
  }
  unsigned long long* red = c->red + 3 * RED_STRIDE;
  CK(hj_launch_init_reduce(red, c->D, s));
  if (c->weno == HJ_WENO_INTENDED) {
    CK(hj_launch_init_eps(c->eps, c->D, s));
    CK(hj_launch_maxd1sq(*g, in, c->eps, -1, s));
  }
  if (!y_dev && !c->halo0 && c->system_id != HJ_SYS_NONE && use_tma(c) && !hj_tma_plan_is_split(c->plan)) {
    // resident state of a whole system on the plane-ring backend: the reduce-only pass is the stage-1 kernel itself with
    // dt = 0 and its reductions on, its output parked in the buffer that is dead at this point of the step (stage 1 / 3:
    // buffer 1, stage 2 and the RK2 final stage: buffer 2 -- the real stage launch that follows overwrites it or no
    // longer reads it).  The
    // derivative range does not depend on the system's parameter block, so a stale block is harmless.  One field read
    // through the TMA ring (1.5 ms at 512^3) instead of D cached gathers per node (2.1 ms).
    static const int in_idx[5] = {0, 0, 1, 2, 1}, dead[5] = {0, 1, 2, 1, 2};
    KStage st{};
    st.stage = 1;
    st.comp = HJ_COMP_NONE;
    st.want_reduce = 1;
    st.fin_a = 1.0 / 3.0;
    st.fin_b = 2.0;
    st.dt = 0.0;
    st.in = in;
    st.y0 = c->buf[0] + c->origin;
    st.tmp = c->buf[dead[stage]] + c->origin;
    st.out = c->buf[dead[stage]] + c->origin;
    for (int d = 0; d < c->D; ++d) st.out_stride[d] = c->gp.stride[d];
    st.red = red;
    st.epsmax = c->eps;
    CK(hj_launch_stage_tma(c->plan, c->system_id, c->weno, c->gp, c->ks, st, in_idx[stage], s, 0, 0, 0, 0, 0));
  } else {
    CK(hj_launch_deriv_range(c->weno, *g, in, c->eps, red, s));
  }
  CK(cudaMemcpyAsync(c->pinned, red, RED_STRIDE * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  double rec[RED_STRIDE];
  decode_record((const unsigned long long*)c->pinned, c->D, rec);
  for (int d = 0; d < c->D; ++d) {
    deriv_min[d] = rec[c->D + d];
    deriv_max[d] = rec[2 * c->D + d];
  }
  return HJ_OK;
}

static int sys_op_ready(hj_ctx* c, const char* who) {
  int r = check_ready(c, true);
  if (r) return r;
  if (c->halo0 || c->nbatch) return fail(HJ_ERR_UNSUPPORTED, "%s: dense-array entry point is not available on a slab / batch context", who);
  return HJ_OK;
}

int hj_ham(hj_ctx* c, void* stream, double t, const double* const* deriv_c_dev, double* ham_dev) {
  (void)t;
  int r = sys_op_ready(c, "hj_ham");
  if (r) return r;
  if (!deriv_c_dev || !ham_dev) return fail(HJ_ERR_INVALID, "hj_ham: null argument");
  for (int d = 0; d < c->D; ++d)
    if (!deriv_c_dev[d]) return fail(HJ_ERR_INVALID, "hj_ham: deriv[%d] is null", d);
  CK(cudaSetDevice(c->device));
  CK(hj_launch_sys_op(c->system_id, 0, c->gd, c->ks, deriv_c_dev, nullptr, ham_dev, 0, nullptr, (cudaStream_t)stream));
  return HJ_OK;
}

int hj_alpha(hj_ctx* c, void* stream, double t, int dim, double* alpha_dev) {
  (void)t;
  int r = sys_op_ready(c, "hj_alpha");
  if (r) return r;
  if (!alpha_dev) return fail(HJ_ERR_INVALID, "hj_alpha: null argument");
  if (dim < 0 || dim >= c->D) return fail(HJ_ERR_INVALID, "Illegal dim parameter");
  CK(cudaSetDevice(c->device));
  CK(hj_launch_sys_op(c->system_id, 1, c->gd, c->ks, nullptr, nullptr, alpha_dev, dim, nullptr, (cudaStream_t)stream));
  return HJ_OK;
}

int hj_diss_glf(hj_ctx* c, void* stream, double t, const double* const* deriv_l_dev, const double* const* deriv_r_dev,
                double* diss_dev, double* step_bound, double* reduce_host) {
  (void)t;
  int r = sys_op_ready(c, "hj_diss_glf");
  if (r) return r;
  if (!deriv_l_dev || !deriv_r_dev || !diss_dev) return fail(HJ_ERR_INVALID, "hj_diss_glf: null argument");
  for (int d = 0; d < c->D; ++d)
    if (!deriv_l_dev[d] || !deriv_r_dev[d]) return fail(HJ_ERR_INVALID, "hj_diss_glf: derivL/derivR[%d] is null", d);
  CK(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* red = c->red + 3 * RED_STRIDE;
  CK(hj_launch_init_reduce(red, c->D, s));
  CK(hj_launch_sys_op(c->system_id, 2, c->gd, c->ks, deriv_l_dev, deriv_r_dev, diss_dev, 0, red, s));
  CK(cudaMemcpyAsync(c->pinned, red, RED_STRIDE * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  double rec[RED_STRIDE];
  decode_record((const unsigned long long*)c->pinned, c->D, rec);
  if (reduce_host) std::memcpy(reduce_host, rec, HJ_REDUCE_LEN(c->D) * sizeof(double));
  if (step_bound) *step_bound = step_bound_from_alpha(c, rec);
  return HJ_OK;
}

int hj_alpha_max(hj_ctx* c, void* stream, double t, double* alpha_max_host, double* step_bound) {
  (void)t;
  int r = check_ready(c, true);
  if (r) return r;
  if (c->nbatch) return fail(HJ_ERR_UNSUPPORTED, "hj_alpha_max: batch flock alphas are host scalars of the parameter blocks");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (!c->alpha_valid) {
    unsigned long long* red = c->red + 3 * RED_STRIDE;
    CK(hj_launch_init_reduce(red, c->D, s));
    CK(hj_launch_alpha_max(c->system_id, c->gp, c->ks, red, s));
    CK(cudaMemcpyAsync(c->pinned, red, RED_STRIDE * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    double rec[RED_STRIDE];
    decode_record((const unsigned long long*)c->pinned, c->D, rec);
    for (int d = 0; d < c->D; ++d) c->alpha_cache[d] = rec[d];
    c->step_bound_cache = step_bound_from_alpha(c, rec);
    c->alpha_valid = true;
  }
  if (alpha_max_host) std::memcpy(alpha_max_host, c->alpha_cache, c->D * sizeof(double));
  if (step_bound) *step_bound = c->step_bound_cache;
  return HJ_OK;
}

int hj_stage_io(const hj_ctx* c, int stage, int* in_buffer, int* out_buffer) {
  if (!c || stage < 1 || stage > 3) return fail(HJ_ERR_INVALID, "hj_stage_io: stage must be 1..3");
  static const int in_[4] = {0, 0, 1, 2}, out_[4] = {0, 1, 2, 0};
  if (in_buffer) *in_buffer = in_[stage];
  if (out_buffer) *out_buffer = out_[stage];
  return HJ_OK;
}

int hj_eps_prepass(hj_ctx* c, void* stream, int buf, uint64_t** eps_dev) {
  if (!c || buf < 0 || buf > 2) return fail(HJ_ERR_INVALID, "hj_eps_prepass: bad argument");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  CK(hj_launch_init_eps(c->eps, c->D, s));
  CK(hj_launch_maxd1sq(c->gp, c->buf[buf] + c->origin, c->eps, -1, s));
  if (eps_dev) *eps_dev = (uint64_t*)c->eps;
  return HJ_OK;
}

int hj_fill_edge_halo(hj_ctx* c, void* stream, int buf, int side) {
  if (!c || buf < 0 || buf > 2 || side < 0 || side > 1) return fail(HJ_ERR_INVALID, "hj_fill_edge_halo: bad argument");
  if (!c->halo0) return fail(HJ_ERR_STATE, "hj_fill_edge_halo: context has no stored halo planes (dim 0 is not HJ_BC_HALO)");
  CK(cudaSetDevice(c->device));
  int r = ensure_buffers(c);
  if (r) return r;
  CK(hj_launch_edge_halo(c->buf[buf], c->plane, c->gp.N[0], side, c->gp.slope_mult[0], (cudaStream_t)stream));
  return HJ_OK;
}

int hj_dev_alloc(int device, int64_t bytes, void** out) {
  if (!out || bytes < 0) return fail(HJ_ERR_INVALID, "hj_dev_alloc: bad argument");
  CK(cudaSetDevice(device));
  CK(cudaMalloc(out, (size_t)(bytes > 0 ? bytes : 8)));
  return HJ_OK;
}
int hj_dev_free(void* p) {
  if (p) CK(cudaFree(p));
  return HJ_OK;
}
int hj_memcpy(void* dst, const void* src, int64_t bytes, int kind, void* stream, int sync) {
  // kind: 1 = host->device, 2 = device->host, 3 = device->device
  cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : (kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
  CK(cudaMemcpyAsync(dst, src, (size_t)bytes, k, (cudaStream_t)stream));
  if (sync) CK(cudaStreamSynchronize((cudaStream_t)stream));
  return HJ_OK;
}
int hj_stream_sync(void* stream) {
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  return HJ_OK;
}
int hj_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

static bool use_tma(hj_ctx* c) {
  if (c->backend == HJ_BACKEND_GATHER) return false;
  if ((c->weno == HJ_SCHEME_ENO3A || c->weno == HJ_SCHEME_ENO2) &&
      !(c->system_id == HJ_SYS_DUBINS_REL || c->system_id == HJ_SYS_FLOCK)) {
    // the ENO functors are compiled into the plane-ring kernel for whole 3-D systems; product systems, 2-D grids and
    // batches take the gather backend
    c->plan_err = "upwindFirstENO2 / upwindFirstENO3a: plane-ring kernels exist for whole 3-D systems only";
    return false;
  }
  if (!c->plan_tried) {
    c->plan_tried = true;
    char err[256] = {0};
    c->plan = hj_tma_plan_create(c->gp, c->system_id, c->weno, c->buf, c->halo0, err, sizeof err);
    c->plan_err = err;
  }
  return c->plan != nullptr;
}

// stage 1..3: the TVD-RK3 stages (ode_cfl_3.py:151,184-193,226-241); stage 4: the final stage of the RK2 scheme
// (ode_cfl_2.py: y = 0.5 (y + (y1 + dt f(y1)))), which is the stage-3 kernel reading buffer 1
static int stage_impl(hj_ctx* c, cudaStream_t s, int stage, double dt, const double* params, int comp, int use_obs,
                      int want_reduce, bool run_prepass, bool batch = false, int zbeg = 0, int zend = 0,
                      int which_pass = 0, long long col_begin = 0, long long col_end = 0) {
  static const int in_[5] = {0, 0, 1, 2, 1}, out_[5] = {0, 1, 2, 0, 0};
  const bool final_stage = stage >= 3;
  const int slot = stage == 4 ? 1 : stage - 1;   // reduction record / batch parameter set of this stage
  KSys ks = c->ks;
  if (params) std::memcpy(ks.p, params, c->nparams * sizeof(double));
  KStage st{};
  if (batch) {                    // per-element parameter blocks of this stage + per-element dt
    ks.p[0] = (double)c->nparams;
    ks.tab[HJ_BATCH_TABLE] = c->batch_params + (size_t)slot * c->nbatch * c->nparams;
    st.dt_arr = c->batch_dt;
  }
  st.stage = final_stage ? 3 : stage;
  st.comp = final_stage ? comp : HJ_COMP_NONE;
  st.use_obs = final_stage ? use_obs : 0;
  st.want_reduce = want_reduce;
  st.restrict_sign = c->restrict_sign;
  st.fin_a = stage == 4 ? 0.5 : 1.0 / 3.0;
  st.fin_b = stage == 4 ? 1.0 : 2.0;
  st.dt = dt;
  st.in = c->buf[in_[stage]] + c->origin;
  st.y0 = c->buf[0] + c->origin;
  // dimension-split path: pass 1 parks F_B(in) in the stage's output buffer (stage 3: buffer 1, which is free by
  // then -- buffer 0 still holds y0; RK2 final stage: buffer 2)
  st.tmp = (stage == 3 ? c->buf[1] : (stage == 4 ? c->buf[2] : c->buf[out_[stage]])) + c->origin;
  st.aux = c->aux ? c->aux + c->origin : nullptr;
  st.obs = c->obs ? c->obs + c->origin : nullptr;
  st.out = c->buf[out_[stage]] + c->origin;
  for (int d = 0; d < c->D; ++d) st.out_stride[d] = c->gp.stride[d];
  st.red = c->red + slot * RED_STRIDE;
  st.epsmax = c->eps;
  if (c->halo && which_pass == 2) hj_halo_fused_targets(c, out_[stage], &st.push_lo, &st.push_hi);
  if (st.comp == HJ_COMP_MIN_WITH_AUX || st.comp == HJ_COMP_MAX_WITH_AUX)
    if (!st.aux) return fail(HJ_ERR_STATE, "Need to define target function l(x)!");   // hji_solver.py:584
  if (st.use_obs && !st.obs) return fail(HJ_ERR_STATE, "obstacle field not uploaded");
  if (which_pass && !(use_tma(c) && hj_tma_plan_is_split(c->plan)))
    return fail(HJ_ERR_UNSUPPORTED, "hj_stage_pass: this context does not advance a product system on the dimension-split path");
  if (want_reduce == 1 && which_pass != 2) CK(hj_launch_init_reduce(st.red, c->D, s));   // 2: accumulate only
  if (c->weno == HJ_WENO_INTENDED && run_prepass && which_pass != 2) {
    CK(hj_launch_init_eps(c->eps, c->D, s));
    CK(hj_launch_maxd1sq(c->gp, st.in, c->eps, -1, s));
  }
  if (use_tma(c)) {
    CK(hj_launch_stage_tma(c->plan, c->system_id, c->weno, c->gp, ks, st, in_[stage], s, zbeg, zend, which_pass, col_begin,
                           col_end));
  } else if (c->nbatch) {
    return fail(HJ_ERR_UNSUPPORTED, "batch contexts run on the TMA backend only: %s", c->plan_err.c_str());
  } else {
    if (c->backend == HJ_BACKEND_TMA)
      return fail(HJ_ERR_UNSUPPORTED, "TMA backend unavailable for this grid: %s", c->plan_err.c_str());
    CK(hj_launch_stage_gather(c->system_id, c->weno, c->gp, ks, st, s));
  }
  return HJ_OK;
}

int hj_stage(hj_ctx* c, void* stream, int stage, double t, double dt, const double* params, int comp, int use_obstacle,
             int want_reduce) {
  (void)t;
  int r = check_ready(c, true);
  if (r) return r;
  if (c->nbatch) return fail(HJ_ERR_UNSUPPORTED, "hj_stage: use hj_step_batch on a batch context");
  if (!c->have_state) return fail(HJ_ERR_STATE, "hj_stage: no resident state (hj_upload first)");
  if (stage < 1 || stage > 4) return fail(HJ_ERR_INVALID, "hj_stage: stage must be 1..3 (odeCFL3) or 4 (final stage of odeCFL2)");
  if (stage == 4 && c->halo0) return fail(HJ_ERR_UNSUPPORTED, "hj_stage: slab contexts drive the RK3 stages");
  CK(cudaSetDevice(c->device));
  // on a slab the caller runs hj_eps_prepass + allreduce itself before each stage
  return stage_impl(c, (cudaStream_t)stream, stage, dt, params, comp, use_obstacle, want_reduce, !c->halo0);
}

int hj_stage_range(hj_ctx* c, void* stream, int stage, int64_t z_begin, int64_t z_end, double t, double dt,
                   const double* params, int comp, int use_obstacle, int want_reduce) {
  (void)t;
  int r = check_ready(c, true);
  if (r) return r;
  if (c->nbatch) return fail(HJ_ERR_UNSUPPORTED, "hj_stage_range: use hj_step_batch on a batch context");
  if (!c->have_state) return fail(HJ_ERR_STATE, "hj_stage_range: no resident state (hj_upload first)");
  if (stage < 1 || stage > 3) return fail(HJ_ERR_INVALID, "hj_stage_range: stage must be 1..3");
  if (c->D < 3) return fail(HJ_ERR_UNSUPPORTED, "hj_stage_range: needs a grid of >= 3 dims");
  if (c->weno == HJ_WENO_INTENDED) return fail(HJ_ERR_UNSUPPORTED, "hj_stage_range: the intended-WENO eps pre-pass needs the whole field");
  CK(cudaSetDevice(c->device));
  if (!use_tma(c) || hj_tma_plan_is_split(c->plan))
    return fail(HJ_ERR_UNSUPPORTED, "hj_stage_range: whole-system plane-ring (TMA) contexts only");
  const int64_t NZ = c->gp.N[c->D - 3];
  if (z_begin < 0 || z_end > NZ || z_begin >= z_end) return fail(HJ_ERR_INVALID, "hj_stage_range: need 0 <= z_begin < z_end <= N[D-3]");
  return stage_impl(c, (cudaStream_t)stream, stage, dt, params, comp, use_obstacle, want_reduce, false, false, (int)z_begin,
                    (int)z_end);
}

int hj_is_split(const hj_ctx* c) {
  if (!c) return 0;
  hj_ctx* m = const_cast<hj_ctx*>(c);
  if (m->system_id == HJ_SYS_NONE || !m->buf[0]) return 0;
  return use_tma(m) && hj_tma_plan_is_split(m->plan) ? 1 : 0;
}

int hj_stage_pass(hj_ctx* c, void* stream, int stage, int which_pass, double t, double dt, const double* params,
                  int comp, int use_obstacle, int want_reduce) {
  (void)t;
  int r = check_ready(c, true);
  if (r) return r;
  if (c->nbatch) return fail(HJ_ERR_UNSUPPORTED, "hj_stage_pass: use hj_step_batch on a batch context");
  if (!c->have_state) return fail(HJ_ERR_STATE, "hj_stage_pass: no resident state (hj_upload first)");
  if (stage < 1 || stage > 3) return fail(HJ_ERR_INVALID, "hj_stage_pass: stage must be 1..3");
  if (which_pass < 1 || which_pass > 2) return fail(HJ_ERR_INVALID, "hj_stage_pass: pass must be 1 or 2");
  CK(cudaSetDevice(c->device));
  return stage_impl(c, (cudaStream_t)stream, stage, dt, params, comp, use_obstacle, want_reduce, !c->halo0, false, 0, 0,
                    which_pass);
}

int hj_split_cols(hj_ctx* c, int64_t* vlen, int* quantum) {
  if (!c) return fail(HJ_ERR_INVALID, "null ctx");
  long long v = 0;
  int q = 0;
  if (c->system_id == HJ_SYS_NONE || !c->buf[0] || !use_tma(c) || !hj_tma_plan_cols(c->plan, &v, &q))
    return fail(HJ_ERR_UNSUPPORTED, "hj_split_cols: this context does not advance a product system on the dimension-split path");
  if (vlen) *vlen = v;
  if (quantum) *quantum = q;
  return HJ_OK;
}

int hj_stage_pass_cols(hj_ctx* c, void* stream, int stage, int64_t col_begin, int64_t col_end, double t, double dt,
                       const double* params, int comp, int use_obstacle, int want_reduce) {
  (void)t;
  int r = check_ready(c, true);
  if (r) return r;
  if (c->nbatch) return fail(HJ_ERR_UNSUPPORTED, "hj_stage_pass_cols: not available on a batch context");
  if (!c->have_state) return fail(HJ_ERR_STATE, "hj_stage_pass_cols: no resident state (hj_upload first)");
  if (stage < 1 || stage > 3) return fail(HJ_ERR_INVALID, "hj_stage_pass_cols: stage must be 1..3");
  CK(cudaSetDevice(c->device));
  int64_t v = 0;
  int q = 0;
  if ((r = hj_split_cols(c, &v, &q))) return r;
  if (col_begin < 0 || col_begin >= col_end || col_end > v || col_begin % q || (col_end % q && col_end != v))
    return fail(HJ_ERR_INVALID, "hj_stage_pass_cols: need 0 <= col_begin < col_end <= %lld, both multiples of %d (col_end may be the axis length)", (long long)v, q);
  return stage_impl(c, (cudaStream_t)stream, stage, dt, params, comp, use_obstacle, want_reduce, false, false, 0, 0, 2,
                    col_begin, col_end);
}

int hj_step(hj_ctx* c, void* stream, double t, double dt, const double* stage_params, int comp, int use_obstacle,
            int want_reduce) {
  (void)t;
  int r = check_ready(c, true);
  if (r) return r;
  if (c->nbatch) return fail(HJ_ERR_UNSUPPORTED, "hj_step: use hj_step_batch on a batch context");
  if (!c->have_state) return fail(HJ_ERR_STATE, "hj_step: no resident state (hj_upload first)");
  if (c->halo0) return fail(HJ_ERR_UNSUPPORTED, "hj_step: on a slab context drive hj_stage and exchange halos between stages");
  CK(cudaSetDevice(c->device));
  for (int stage = 1; stage <= 3; ++stage) {
    const double* p = stage_params ? stage_params + (stage - 1) * c->nparams : nullptr;
    r = stage_impl(c, (cudaStream_t)stream, stage, dt, p, comp, use_obstacle, want_reduce, true);
    if (r) return r;
  }
  return HJ_OK;
}

int hj_snapshot(hj_ctx* c, void* stream) {
  if (!c) return fail(HJ_ERR_INVALID, "null ctx");
  if (!c->have_state) return fail(HJ_ERR_STATE, "hj_snapshot: no resident state (hj_upload first)");
  CK(cudaSetDevice(c->device));
  if (!c->snap) CK(cudaMalloc(&c->snap, c->elems * sizeof(double)));
  CK(cudaMemcpyAsync(c->snap, c->buf[0], c->elems * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return HJ_OK;
}

int hj_change(hj_ctx* c, void* stream, double* max_abs_change, int* has_nan) {
  if (!c) return fail(HJ_ERR_INVALID, "null ctx");
  if (!c->have_state) return fail(HJ_ERR_STATE, "hj_change: no resident state (hj_upload first)");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* red = c->red + 3 * RED_STRIDE;
  CK(cudaMemsetAsync(red, 0, 2 * sizeof(unsigned long long), s));      // enc_ordered(x) > 0 for every x >= 0
  const double* cur = c->buf[0] + c->origin;
  // without a snapshot the field is compared with itself: change 0, NaN flag only
  const double* ref = c->snap ? c->snap + c->origin : cur;
  CK(hj_launch_change(cur, ref, c->plane * c->gp.N[0], red, s));
  CK(cudaMemcpyAsync(c->pinned, red, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const unsigned long long* h = (const unsigned long long*)c->pinned;
  double m = 0.0;
  if (h[0]) {
    unsigned long long b = (h[0] >> 63) ? (h[0] & 0x7fffffffffffffffull) : ~h[0];
    std::memcpy(&m, &b, 8);
  }
  if (max_abs_change) *max_abs_change = m;
  if (has_nan) *has_nan = h[1] ? 1 : 0;
  return HJ_OK;
}

int hj_discount(hj_ctx* c, void* stream, double gamma, int mode, int take_max, double max_val) {
  if (!c) return fail(HJ_ERR_INVALID, "null ctx");
  if (!c->have_state) return fail(HJ_ERR_STATE, "hj_discount: no resident state (hj_upload first)");
  if (mode < 0 || mode > 2) return fail(HJ_ERR_INVALID, "check your discountFactor and discountMode");     // :638
  const double* ref = mode == 2 ? c->obs : c->aux;
  if (!ref) return fail(HJ_ERR_STATE, mode == 2 ? "obstacle field not uploaded" : "Need to define target function l(x)!");   // :617
  CK(cudaSetDevice(c->device));
  CK(hj_launch_discount(c->buf[0] + c->origin, ref + c->origin, c->plane * c->gp.N[0], gamma, mode, take_max, max_val,
                        (cudaStream_t)stream));
  return HJ_OK;
}

int hj_set_pipeline_planes(hj_ctx* c, int planes) {
  if (!c) return fail(HJ_ERR_INVALID, "null ctx");
  if (planes < 0) return fail(HJ_ERR_INVALID, "hj_set_pipeline_planes: planes must be >= 0 (0 = default)");
  c->pipe_planes = planes;
  return HJ_OK;
}

int hj_set_restrict(hj_ctx* c, int sign) {
  if (!c) return fail(HJ_ERR_INVALID, "null ctx");
  c->restrict_sign = sign > 0 ? 1 : (sign < 0 ? -1 : 0);
  return HJ_OK;
}

int hj_step_rk2(hj_ctx* c, void* stream, double t, double dt, const double* stage_params, int comp, int use_obstacle,
                int want_reduce) {
  (void)t;
  int r = check_ready(c, true);
  if (r) return r;
  if (c->nbatch) return fail(HJ_ERR_UNSUPPORTED, "hj_step_rk2: not available on a batch context");
  if (!c->have_state) return fail(HJ_ERR_STATE, "hj_step_rk2: no resident state (hj_upload first)");
  if (c->halo0) return fail(HJ_ERR_UNSUPPORTED, "hj_step_rk2: slab contexts drive hj_stage (RK3)");
  CK(cudaSetDevice(c->device));
  for (int k = 0; k < 2; ++k) {
    const double* p = stage_params ? stage_params + k * c->nparams : nullptr;
    r = stage_impl(c, (cudaStream_t)stream, k == 0 ? 1 : 4, dt, p, comp, use_obstacle, want_reduce, true);
    if (r) return r;
  }
  return HJ_OK;
}

int hj_step_reductions(hj_ctx* c, void* stream, double* reduce_host) {
  if (!c || !reduce_host) return fail(HJ_ERR_INVALID, "hj_step_reductions: null argument");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(c->pinned, c->red, 3 * RED_STRIDE * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const int L = HJ_REDUCE_LEN(c->D);
  for (int k = 0; k < 3; ++k) decode_record((const unsigned long long*)c->pinned + k * RED_STRIDE, c->D, reduce_host + k * L);
  return HJ_OK;
}

// Host-buffer stepping as a software pipeline (3-D grids whose dim 0 is the marched Z dim): dim 0 is uploaded in chunks;
// as soon as chunk w is up, stage 1 advances every plane whose +3-plane stencil is now resident (up to hi(w) - 3), stage 2
// the planes 3 further down (up to hi(w) - 6), stage 3 those up to hi(w) - 9, and these finished planes go down on a third
// stream at once: the download runs one chunk + 9 planes behind the upload.  Stream order on the compute stream satisfies
// every +-3-plane dependency; stage 3 writes buffer 0 in place only below hi(w) - 9, which no later launch reads and no
// later upload touches.  With pinned host memory the two PCIe directions overlap each other and the compute.
static bool can_pipeline(hj_ctx* c, int is_host) {
  if (!is_host || c->D != 3 || c->halo0 || c->nbatch || c->weno != HJ_WENO_AS_SHIPPED) return false;
  if (c->pitch != c->gp.N[2] || c->gp.bc[0] == HJ_BC_PERIODIC || c->gp.N[0] < 64) return false;
  if (ensure_buffers(c) != HJ_OK) return false;
  return use_tma(c) && !hj_tma_plan_is_split(c->plan);
}

// default chunk of the pipelined step.  Measured at 512^3 (tools/e2e_chunk_sweep.py, profiles/r02_e2e_chunk_sweep.jsonl):
// 16- and 32-plane chunks (32 / 64 MB) 24.0 ms per step, 8 planes 25.5, 4 planes 28.1 (a chunk's three launches re-load a
// 6-plane lead-in); the bare full-duplex transfer of the field is 21.6 ms, the same chunked copies without kernels 22.4.
static constexpr long long PIPE_CHUNK_BYTES = 32ll << 20;

static int pipelined_step(hj_ctx* c, cudaStream_t s, double dt, const double* y_host, double* y_out, int comp, int use_obs) {
  const int N0 = c->gp.N[0];
  int P = c->pipe_planes;
  if (P <= 0) P = (int)((PIPE_CHUNK_BYTES + c->plane * 8 - 1) / (c->plane * 8));
  if (P > (N0 + 7) / 8) P = (N0 + 7) / 8;   // at least 8 chunks
  if (P < 1) P = 1;
  const int C = (N0 + P - 1) / P;
  if (!c->s_h2d) {
    CK(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
  }
  while ((int)c->ev_up.size() < C) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_up.push_back(e);
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_done.push_back(e);
  }
  // copies start after whatever the caller queued on its stream before this call
  CK(cudaEventRecord(c->ev_start, s));
  CK(cudaStreamWaitEvent(c->s_h2d, c->ev_start, 0));
  CK(cudaStreamWaitEvent(c->s_d2h, c->ev_start, 0));
  const size_t plane = (size_t)c->plane;
  auto lo = [&](int k) { return k * P; };
  auto hi = [&](int k) { return (k + 1) * P < N0 ? (k + 1) * P : N0; };
  for (int k = 0; k < C; ++k) {
    CK(cudaMemcpyAsync(c->buf[0] + lo(k) * plane, y_host + lo(k) * plane, (hi(k) - lo(k)) * plane * sizeof(double),
                       cudaMemcpyHostToDevice, c->s_h2d));
    CK(cudaEventRecord(c->ev_up[k], c->s_h2d));
  }
  c->have_state = true;
  int done[4] = {0, 0, 0, 0};                    // planes [0, done[st]) of stage st are launched
  for (int w = 0; w < C; ++w) {
    CK(cudaStreamWaitEvent(s, c->ev_up[w], 0));
    const int d2h_from = done[3];
    for (int stg = 1; stg <= 3; ++stg) {
      int upto = w == C - 1 ? N0 : hi(w) - 3 * stg;
      if (upto <= done[stg]) continue;
      int r;
      if ((r = stage_impl(c, s, stg, dt, nullptr, comp, use_obs, 0, false, false, done[stg], upto))) return r;
      done[stg] = upto;
    }
    if (done[3] > d2h_from) {
      CK(cudaEventRecord(c->ev_done[w], s));
      CK(cudaStreamWaitEvent(c->s_d2h, c->ev_done[w], 0));
      CK(cudaMemcpyAsync(y_out + d2h_from * plane, c->buf[0] + d2h_from * plane,
                         (size_t)(done[3] - d2h_from) * plane * sizeof(double), cudaMemcpyDeviceToHost, c->s_d2h));
    }
  }
  CK(cudaStreamSynchronize(c->s_d2h));
  CK(cudaStreamSynchronize(s));
  return HJ_OK;
}

int hj_ode_cfl3_step(hj_ctx* c, void* stream, double t, double t_end, double factor_cfl, double max_step,
                     const double* y_in, double* y_out, int is_host, int comp, int use_obstacle, double* t_new,
                     double* dt_out) {
  int r = check_ready(c, true);
  if (r) return r;
  if (!y_in || !y_out) return fail(HJ_ERR_INVALID, "hj_ode_cfl3_step: null y");
  if (c->halo0 || c->nbatch) return fail(HJ_ERR_UNSUPPORTED, "hj_ode_cfl3_step: not available on a slab / batch context");
  // a Flock re-derives its parameter block on each of the three RHS evaluations (flock.py:213) and its alphas are host
  // scalars of those blocks: this entry point takes neither, so it must not step one with a frozen block
  if (c->system_id == HJ_SYS_FLOCK)
    return fail(HJ_ERR_UNSUPPORTED, "hj_ode_cfl3_step: HJ_SYS_FLOCK needs per-stage parameter blocks; use hj_upload + hj_step(stage_params) + hj_download");
  if (factor_cfl < 0.0) return fail(HJ_ERR_INVALID, "FactorCFL must be a positive scalar double value");   // ode_cfl_set.py:104
  if (max_step < 0.0) return fail(HJ_ERR_INVALID, "MaxStep must be a positive scalar double value");       // ode_cfl_set.py:106
  double sb = 0;
  r = hj_alpha_max(c, stream, t, nullptr, &sb);
  if (r) return r;
  // ode_cfl_3.py:142-143
  double dt = factor_cfl * sb;
  if (t_end - t < dt) dt = t_end - t;
  if (max_step < dt) dt = max_step;
  if (can_pipeline(c, is_host)) {
    r = pipelined_step(c, (cudaStream_t)stream, dt, y_in, y_out, comp, use_obstacle);
    if (r) return r;
  } else {
    r = hj_upload(c, stream, HJ_FIELD_STATE, y_in, is_host);
    if (r) return r;
    r = hj_step(c, stream, t, dt, nullptr, comp, use_obstacle, 0);
    if (r) return r;
    r = hj_download(c, stream, HJ_FIELD_STATE, y_out, is_host);
    if (r) return r;
    if (!is_host) CK(cudaStreamSynchronize((cudaStream_t)stream));
  }
  // ode_cfl_3.py:145,178,187,220,236 -- time arithmetic kept verbatim
  const double t1 = t + dt;
  const double t2 = t1 + dt;
  const double t_half = 0.25 * (3 * t + t2);
  const double t_three_half = t_half + dt;
  if (t_new) *t_new = (1.0 / 3.0) * (t + 2 * t_three_half);
  if (dt_out) *dt_out = dt;
  return HJ_OK;
}

int hj_ode_cfl3_single(hj_ctx* c, void* stream, double t, double t_end, double factor_cfl, double max_step,
                       double* y_inout, int is_host, int comp, int use_obstacle, double* t_new, double* dt_out) {
  if (!y_inout) return fail(HJ_ERR_INVALID, "hj_ode_cfl3_single: null y");
  return hj_ode_cfl3_step(c, stream, t, t_end, factor_cfl, max_step, y_inout, y_inout, is_host, comp, use_obstacle, t_new,
                          dt_out);
}

/* pinned host memory for the host-buffer entry points (cudaHostAlloc: both PCIe directions run at full rate and the
 * chunked copies of the pipelined step overlap its kernels) */
int hj_host_alloc(int64_t bytes, void** out) {
  if (!out || bytes < 0) return fail(HJ_ERR_INVALID, "hj_host_alloc: bad argument");
  CK(cudaHostAlloc(out, (size_t)(bytes > 0 ? bytes : 8), cudaHostAllocDefault));
  return HJ_OK;
}
int hj_host_free(void* p) {
  if (p) CK(cudaFreeHost(p));
  return HJ_OK;
}

int hj_create_batch(hj_ctx** out, int device, int nbatch, int ndim, const int64_t* N, const double* dx,
                    const int* bc_kind, const int* bc_toward_zero, int weno_mode) {
  if (!out || !N || !dx || !bc_kind) return fail(HJ_ERR_INVALID, "hj_create_batch: null argument");
  if (nbatch < 1) return fail(HJ_ERR_INVALID, "hj_create_batch: nbatch must be >= 1");
  if (ndim != 3) return fail(HJ_ERR_UNSUPPORTED, "hj_create_batch: batches of 3-D grids only (got %d-D)", ndim);
  if (weno_mode != HJ_WENO_AS_SHIPPED)
    return fail(HJ_ERR_UNSUPPORTED, "hj_create_batch: the 'maxOverGrid' epsilon of the intended WENO is per grid; batch contexts run as_shipped only");
  int64_t Nb[HJ_MAX_DIM];
  double dxb[HJ_MAX_DIM];
  int bcb[HJ_MAX_DIM], tzb[HJ_MAX_DIM];
  Nb[0] = nbatch; dxb[0] = 1.0; bcb[0] = HJ_BC_EXTRAPOLATE; tzb[0] = 0;
  for (int d = 0; d < ndim; ++d) {
    if (bc_kind[d] == HJ_BC_HALO) return fail(HJ_ERR_UNSUPPORTED, "hj_create_batch: halo dims are for slab contexts");
    Nb[d + 1] = N[d]; dxb[d + 1] = dx[d]; bcb[d + 1] = bc_kind[d]; tzb[d + 1] = bc_toward_zero ? bc_toward_zero[d] : 0;
  }
  hj_ctx* c = nullptr;
  // hj_create checks every dim against the smallest extent its ghost cells need (2 for an extrapolated dim); the
  // batch dim is never differentiated, so a batch of 1 grid is legal: create with a padded extent and shrink it
  const int64_t nb_create = nbatch < 2 ? 2 : nbatch;
  Nb[0] = nb_create;
  int r = hj_create(&c, device, ndim + 1, Nb, dxb, bcb, tzb, weno_mode);
  if (r) return r;
  if (nb_create != nbatch) {         // shrink the batch extent: strides of dims >= 1 do not depend on N[0]
    c->gp.N[0] = c->gd.N[0] = nbatch;
    c->nodes = c->nodes / nb_create * nbatch;
    c->elems = c->plane * nbatch;
  }
  c->nbatch = nbatch;
  c->axes_set |= 1u;                 // the batch dim has no axis
  cudaError_t e = cudaMalloc(&c->batch_dt, nbatch * sizeof(double));
  if (e != cudaSuccess) { hj_destroy(c); return fail(HJ_ERR_CUDA, "hj_create_batch: allocation failed: %s", cudaGetErrorString(e)); }
  *out = c;
  return HJ_OK;
}

int hj_step_batch(hj_ctx* c, void* stream, const double* dt_host, const double* stage_params_host, int comp,
                  int use_obstacle) {
  int r = check_ready(c, true);
  if (r) return r;
  if (!c->nbatch) return fail(HJ_ERR_STATE, "hj_step_batch: not a batch context (hj_create_batch)");
  if (!dt_host || !stage_params_host) return fail(HJ_ERR_INVALID, "hj_step_batch: null argument");
  if (!c->have_state) return fail(HJ_ERR_STATE, "hj_step_batch: no resident state (hj_upload first)");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t np = (size_t)3 * c->nbatch * c->nparams;
  if (!c->batch_params) CK(cudaMalloc(&c->batch_params, (size_t)3 * c->nbatch * HJ_MAX_PARAMS * sizeof(double)));
  CK(cudaMemcpyAsync(c->batch_dt, dt_host, c->nbatch * sizeof(double), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(c->batch_params, stage_params_host, np * sizeof(double), cudaMemcpyHostToDevice, s));
  for (int stage = 1; stage <= 3; ++stage) {
    r = stage_impl(c, s, stage, 0.0, nullptr, comp, use_obstacle, 0, false, true);
    if (r) return r;
  }
  return HJ_OK;
}

int hj_batch_size(const hj_ctx* c) { return c ? c->nbatch : 0; }

}  // extern "C"
