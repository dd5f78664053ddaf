// hj_internal.h -- host-side declarations shared by the C-ABI (hj_api.cu) and the kernel launchers.
#pragma once
#include <cuda_runtime.h>

#include "hj_common.cuh"

struct HjLaunchInfo {
  long long nouter;   // product of N[0..D-3]
};

// gather backend (hj_gather.cu)
cudaError_t hj_launch_stage_gather(int system_id, int weno, const KGrid& g, const KSys& ks, const KStage& st,
                                   cudaStream_t s);
cudaError_t hj_launch_deriv(int weno, const KGrid& g, const double* in, int dim, double* dl, double* dr,
                            const unsigned long long* epsmax, cudaStream_t s);
cudaError_t hj_launch_sys_op(int system_id, int op, const KGrid& g, const KSys& ks, const double* const* a,
                             const double* const* b, double* out, int dl, unsigned long long* red, cudaStream_t s);
cudaError_t hj_launch_deriv_all(const KGrid& g, const double* in, int dim, double* out6, cudaStream_t s);
cudaError_t hj_launch_add_ghost(const KGrid& g, const double* in, int dim, int width, double* out, cudaStream_t s);
cudaError_t hj_launch_alpha_max(int system_id, const KGrid& g, const KSys& ks, unsigned long long* red,
                                cudaStream_t s);
cudaError_t hj_launch_maxd1sq(const KGrid& g, const double* in, unsigned long long* epsmax, int only_dim,
                              cudaStream_t s);
cudaError_t hj_launch_deriv_range(int weno, const KGrid& g, const double* in, const unsigned long long* epsmax,
                                  unsigned long long* red, cudaStream_t s);
cudaError_t hj_launch_init_reduce(unsigned long long* red, int D, cudaStream_t s);
cudaError_t hj_launch_init_eps(unsigned long long* eps, int D, cudaStream_t s);
cudaError_t hj_launch_edge_halo(double* buf, long long plane, int n0, int side, double m, cudaStream_t s);
cudaError_t hj_launch_change(const double* a, const double* b, long long n, unsigned long long* red, cudaStream_t s);
cudaError_t hj_launch_discount(double* y, const double* ref, long long n, double gamma, int mode, int take_max,
                               double max_val, cudaStream_t s);
cudaError_t hj_launch_pack(const double* dense, double* pitched, const KGrid& g_dense, const KGrid& g_pitched,
                           cudaStream_t s);
cudaError_t hj_launch_unpack(const double* pitched, double* dense, const KGrid& g_dense, const KGrid& g_pitched,
                             cudaStream_t s);

// TMA backend (hj_tma.cu)
struct HjTmaPlan;   // tensor maps + tile geometry for one context (hj_tma_plan.h)
HjTmaPlan* hj_tma_plan_create(const KGrid& g_pitched, int system_id, int weno, double* const bufs[3], int halo0,
                              char* err, int errlen, int tile_y = 0);
void hj_tma_plan_destroy(HjTmaPlan* p);
cudaError_t hj_launch_stage_tma(HjTmaPlan* plan, int system_id, int weno, const KGrid& g, const KSys& ks,
                                const KStage& st, int in_buf, cudaStream_t s, int zbeg = 0, int zend = 0,
                                int which_pass = 0, long long col_begin = 0, long long col_end = 0);
bool hj_tma_plan_cols(const HjTmaPlan* p, long long* vlen, int* quantum);
bool hj_tma_plan_is_split(const HjTmaPlan* p);

void hj_count_launch(int n);
