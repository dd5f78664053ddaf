// tune_tma.cu -- developer harness (not part of the product): times variants of the plane-ring stage kernel
// (hj_tma_kernel.cuh) on an air3D-shaped problem and checks that every variant writes the same bits.
//
//   build : python tools/build_tune.py            (links against the objects of levelsetpy_b200/build)
//   run   : tools/tune_tma [N=512] [reps=10]
//
// Prints one line per variant: ms per launch for stages 1/2/3, their sum (= ms per TVD-RK3 step) and the fraction
// of the HBM roofline (64 B per point-step at the peak given as argv[3], default 6550 GB/s).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../levelsetpy_b200/csrc/hj_internal.h"
#include "../levelsetpy_b200/csrc/hj_tma_kernel.cuh"
#include "hj_quad_kernel.cuh"
#include "../levelsetpy_b200/csrc/hj_tma_plan.h"

using namespace hjtma;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } \
  } while (0)

__global__ void k_fill(double* p, long long n, int N, double a) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % N), y = (int)((i / N) % N), z = (int)(i / ((long long)N * N));
    const double fx = -6.0 + 26.0 * z / (N - 1), fy = -10.0 + 20.0 * y / (N - 1), th = 6.283185307179586 * x / N;
    p[i] = sqrt(fx * fx + fy * fy) - 5.0 + a * sin(th + 0.3 * fx) * cos(0.2 * fy);
  }
}
__global__ void k_xor(const unsigned long long* p, long long n, unsigned long long* out) {
  unsigned long long acc = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc ^= p[i] * (unsigned long long)(2 * (i % 1000003) + 1);
  for (int o = 16; o > 0; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicXor(out, acc);
}

// max |a - b|, number of differing words and the first differing index (debugging a new variant against the reference)
__global__ void k_diff(const double* a, const double* b, long long n, unsigned long long* out) {
  double m = 0.0;
  unsigned long long cnt = 0, first = ~0ull;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double x = a[i], y = b[i];
    if (__double_as_longlong(x) != __double_as_longlong(y)) {
      ++cnt;
      if ((unsigned long long)i < first) first = (unsigned long long)i;
      const double d = fabs(x - y);
      m = (d == d) ? fmax(m, d) : INFINITY;
    }
  }
  atomicAdd(out, cnt);
  atomicMin(out + 1, first);
  atomicMax(out + 2, (unsigned long long)__double_as_longlong(m));   // non-negative doubles order like integers
}

struct Problem {
  int N;
  KGrid g;
  KSys ks;
  double* buf[3];
  HjTmaPlan* plans[64];      // by tile height
  const char* only;          // substring filter on the variant name
  unsigned long long* xor_dev;
  double peak;
  double* ref_out[3];        // stage outputs of the first variant (device copies), for k_diff
  int cz_override;           // > 0: planes per Z chunk for the quad variants
};

static unsigned long long checksum(Problem& P, const double* p) {
  CK(cudaMemset(P.xor_dev, 0, 8));
  k_xor<<<1184, 256>>>((const unsigned long long*)p, (long long)P.N * P.N * P.N, P.xor_dev);
  unsigned long long h;
  CK(cudaMemcpy(&h, P.xor_dev, 8, cudaMemcpyDeviceToHost));
  return h;
}

template <class Cfg, int STAGE, bool QUAD>
static auto pick_kernel() {
  if constexpr (QUAD) return k_stage_quad<SysDubinsRel, HJ_WENO_AS_SHIPPED, false, STAGE, Cfg>;
  else return k_stage_tma<SysDubinsRel, 3, HJ_WENO_AS_SHIPPED, false, STAGE, Cfg>;
}

template <class Cfg, int STAGE, bool QUAD = false>
static float run_stage(Problem& P, int reps, unsigned long long* sum) {
  auto kern = pick_kernel<Cfg, STAGE, QUAD>();
  constexpr size_t smem = Cfg::template smem_bytes<STAGE>();
  constexpr int NTHREADS = Cfg::NTHREADS;
  if (!P.plans[Cfg::TY]) {
    char err[256] = {0};
    P.plans[Cfg::TY] = hj_tma_plan_create(P.g, HJ_SYS_DUBINS_REL, HJ_WENO_AS_SHIPPED, P.buf, 0, err, sizeof err, Cfg::TY);
    if (!P.plans[Cfg::TY]) { printf("plan: %s\n", err); exit(1); }
  }
  HjTmaPlan planv = *P.plans[Cfg::TY];
  HjTmaPlan* plan = &planv;
  if (QUAD && P.cz_override > 0) {
    plan->geo.cz = P.cz_override;
    plan->geo.nzc = (P.N + plan->geo.cz - 1) / plan->geo.cz;
    plan->nblocks = plan->tiles * plan->geo.nzc;
  }
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHREADS, smem));
  KStage st{};
  st.stage = STAGE;
  st.comp = STAGE == 3 ? HJ_COMP_MIN_OVER_TIME : HJ_COMP_NONE;
  st.dt = 1e-3;
  st.fin_a = 1.0 / 3.0;   // TVD-RK3 final combination (ode_cfl_3.py:241), as hj_api.cu sets it
  st.fin_b = 2.0;
  const int in_buf = STAGE - 1;
  st.in = P.buf[in_buf];
  st.y0 = P.buf[0];
  st.out = STAGE == 2 ? P.buf[2] : P.buf[1];   // stage 3 writes buf1 here so that every launch sees the same inputs
  for (int d = 0; d < 3; ++d) st.out_stride[d] = P.g.stride[d];
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) kern<<<(unsigned)plan->nblocks, NTHREADS, smem>>>(plan->tmap[in_buf], plan->tmap_y0, P.g, P.ks, st, plan->geo);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) kern<<<(unsigned)plan->nblocks, NTHREADS, smem>>>(plan->tmap[in_buf], plan->tmap_y0, P.g, P.ks, st, plan->geo);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  *sum = checksum(P, st.out);
  const long long n = (long long)P.N * P.N * P.N;
  if (!P.ref_out[STAGE - 1]) {
    CK(cudaMalloc(&P.ref_out[STAGE - 1], n * 8));
    CK(cudaMemcpy(P.ref_out[STAGE - 1], st.out, n * 8, cudaMemcpyDeviceToDevice));
  } else if (QUAD) {
    unsigned long long h[3] = {0, ~0ull, 0}, *d;
    CK(cudaMalloc(&d, 24));
    CK(cudaMemcpy(d, h, 24, cudaMemcpyHostToDevice));
    k_diff<<<1184, 256>>>(st.out, P.ref_out[STAGE - 1], n, d);
    CK(cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost));
    CK(cudaFree(d));
    if (h[0]) {
      double m;
      memcpy(&m, &h[2], 8);
      const long long i = (long long)h[1];
      printf("\n   stage %d: %llu words differ, first at (z %lld, y %lld, x %lld), max |diff| %.3e\n   ", STAGE, h[0],
             i / ((long long)P.N * P.N), (i / P.N) % P.N, i % P.N, m);
    }
  }
  if (STAGE == 1) printf("[occ %d regs ", occ);
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kern));
  printf("%d%s", fa.numRegs, STAGE == 3 ? "] " : "/");
  return ms / reps;
}

template <class Cfg, class Cfg23 = Cfg, bool QUAD = false>
static void run_variant(Problem& P, const char* name, int reps, unsigned long long ref[3]) {
  if (P.only && ref[0] && !strstr(name, P.only)) return;     // the first variant always runs: it is the reference
  printf("%-22s ", name);
  unsigned long long s[3];
  // stage order matters: stage 2 reads buf1 (stage-1 output), stage 3 reads buf2 and overwrites buf1
  const float t1 = run_stage<Cfg, 1, QUAD>(P, reps, &s[0]);
  const float t2 = run_stage<Cfg23, 2, QUAD>(P, reps, &s[1]);
  const float t3 = run_stage<Cfg23, 3, QUAD>(P, reps, &s[2]);
  bool same = true;
  for (int i = 0; i < 3; ++i) {
    if (!ref[i]) ref[i] = s[i];
    same = same && ref[i] == s[i];
  }
  const double pts = (double)P.N * P.N * P.N, tot = t1 + t2 + t3;
  printf("s1 %.3f  s2 %.3f  s3 %.3f  step %.3f ms  %.1f Gpt-steps/s  %.1f%% of %.0f GB/s  %s\n", t1, t2, t3, tot,
         pts / tot * 1e-6, 100.0 * pts * 64.0 / (tot * 1e-3) / (P.peak * 1e9), P.peak, same ? "bits==" : "BITS DIFFER");
  fflush(stdout);
}

int main(int argc, char** argv) {
  Problem P{};
  P.N = argc > 1 ? atoi(argv[1]) : 512;
  const int reps = argc > 2 ? atoi(argv[2]) : 10;
  P.peak = argc > 3 ? atof(argv[3]) : 6550.0;
  const int N = P.N;
  if (N % 2) { printf("N must be even\n"); return 1; }
  const long long n = (long long)N * N * N;
  for (int b = 0; b < 3; ++b) CK(cudaMalloc(&P.buf[b], n * 8));
  CK(cudaMalloc(&P.xor_dev, 8));
  k_fill<<<1184, 256>>>(P.buf[0], n, N, 0.1);
  CK(cudaMemset(P.buf[1], 0, n * 8));
  CK(cudaMemset(P.buf[2], 0, n * 8));
  // air3D box: [-6,20] x [-10,10] x [0, 2 pi (1 - 1/N)], dim 2 periodic
  const double lo[3] = {-6, -10, 0}, hi[3] = {20, 10, 6.283185307179586 * (1.0 - 1.0 / N)};
  KGrid& g = P.g;
  g.D = 3;
  long long s = 1;
  std::vector<double> vs(N), cs(N), sn(N);
  for (int d = 2; d >= 0; --d) {
    g.N[d] = N;
    g.dx[d] = (hi[d] - lo[d]) / (N - 1);
    g.dxinv[d] = 1 / g.dx[d];
    g.bc[d] = d == 2 ? HJ_BC_PERIODIC : HJ_BC_EXTRAPOLATE;
    g.slope_mult[d] = 1.0;
    g.ca1[d] = g.dxinv[d] * (45.0 / 60.0); g.ca2[d] = g.dxinv[d] * (-9.0 / 60.0);
    g.ca3[d] = g.dxinv[d] * (1.0 / 60.0);  g.cb[d] = g.dxinv[d] * (1.0 / 60.0);
    g.stride[d] = s;
    s *= N;
    for (int i = 0; i < N; ++i) vs[i] = lo[d] + i * g.dx[d];
    double* dv;
    CK(cudaMalloc(&dv, N * 8));
    CK(cudaMemcpy(dv, vs.data(), N * 8, cudaMemcpyHostToDevice));
    g.vs[d] = dv;
    if (d == 2) {
      for (int i = 0; i < N; ++i) { cs[i] = cos(vs[i]); sn[i] = sin(vs[i]); }
      double *dc, *dsn;
      CK(cudaMalloc(&dc, N * 8)); CK(cudaMalloc(&dsn, N * 8));
      CK(cudaMemcpy(dc, cs.data(), N * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dsn, sn.data(), N * 8, cudaMemcpyHostToDevice));
      P.ks.tab[0] = dc; P.ks.tab[1] = dsn;
    }
  }
  P.ks.p[0] = 5; P.ks.p[1] = 5; P.ks.p[2] = 1; P.ks.p[3] = 1; P.ks.p[4] = 1;
  P.only = argc > 4 && strcmp(argv[4], "-") ? argv[4] : nullptr;
  P.cz_override = argc > 5 ? atoi(argv[5]) : 0;
  printf("N=%d reps=%d\n", N, reps);
  unsigned long long ref[3] = {0, 0, 0};
#define V(R, MINB, U, TY, SEQ, OPT) \
  run_variant<TmaCfg<R, MINB, U, TY, 16, SEQ, OPT>>(P, "R" #R "_b" #MINB "_u" #U "_ty" #TY "_seq" #SEQ "_opt" #OPT, reps, ref)
  V(8, 2, 1, 16, false, 143);   // production
  // 2 x 2 nodes per thread (hj_quad_kernel.cuh): 32 x 24 tile, 192 threads, 2 CTAs/SM; ring 8 for stage 1, 7 for stages
  // 2/3 (the y0 tiles share the shared memory); Q1: X windows of plane z+1 prefetched
#define Q(R1, R23, MINB, TY, OPT) \
  run_variant<QuadCfg<R1, MINB, TY, 16, OPT>, QuadCfg<R23, MINB, TY, 16, OPT>, true>(P, "quad_R" #R1 "_" #R23 "_b" #MINB "_ty" #TY "_opt" #OPT, reps, ref)
  Q(8, 7, 2, 24, 0);
  Q(8, 7, 2, 24, 1);
  Q(8, 8, 1, 24, 0);            // 1 CTA/SM: no register cap
  Q(8, 8, 1, 32, 1);            // 256 threads, 1 CTA/SM
  Q(6, 6, 3, 16, 0);            // 32 x 16 tile, 128 threads, 3 CTAs/SM
  V(8, 3, 1, 12, false, 143);   // 3 CTAs of 192 threads (18 warps/SM, <= 112 registers)
  V(6, 3, 1, 12, false, 143);
  V(8, 2, 1, 16, true, 143);    // one dim at a time
  V(8, 2, 1, 16, false, 135);   // without the dim-pipelined load order
  V(8, 2, 1, 16, false, 7);     // without the SIMPLE fast loop
  V(8, 2, 1, 16, false, 0);     // round-start plane body
  // bound-finding experiments: results differ by construction
  V(8, 2, 1, 16, false, 23);    // no arithmetic
  V(8, 2, 1, 16, false, 39);    // every load hits L2
  V(8, 2, 1, 16, false, 103);   // every load hits L2, no store
  V(8, 2, 1, 16, false, 119);   // nothing but the ring: no arithmetic, no DRAM traffic
  return 0;
}
