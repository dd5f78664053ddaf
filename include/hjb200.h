/* hjb200.h -- C-ABI of the B200-native explicit Hamilton-Jacobi time-stepping hot path.
 *
 * Drop-in boundary for robotsorcerer/LevelSetPy's WENO5 + global Lax-Friedrichs + TVD-RK3 path.
 * The reference has no FFI layer (it is Python callables stored in Bundles); each
 * entry point below names the reference callable whose work it performs.  Paths are relative to the
 * reference checkout.
 *
 * Conventions
 *   - plain C: opaque context, raw pointers, sizes; no C++/torch types; no exceptions cross.
 *   - every function returns 0 on success, a negative hj_status on failure; hj_last_error() gives text.
 *   - all fields are fp64, C-order ("dense": last dim contiguous, exactly y.reshape(grid.shape)).
 *     Inside the context fields live in a pitched layout (innermost row padded to an even count so
 *     rows are 16-byte aligned for TMA / vector access); upload/download convert.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All device work is
 *     enqueued on it.  Functions documented "synchronises" wait for that stream before returning.
 *   - one context per (process, device, slab).  Not thread-safe per context.
 *   - the caller owns every buffer it passes; the context owns its state, scratch and reductions.
 *   - there is NO CPU fallback: without a CUDA device hj_create fails.
 */
#ifndef HJB200_H
#define HJB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HJ_MAX_DIM 6
#define HJ_MAX_PARAMS 96
#define HJ_MAX_TABLES 8
#define HJ_GHOST 3 /* stencil half-width of upwindFirstWENO5a (ENO3aHelper.py:61) */

typedef struct hj_ctx hj_ctx;

typedef enum hj_status {
  HJ_OK = 0,
  HJ_ERR_INVALID = -1,      /* bad argument (the reference raises ValueError / assert)      */
  HJ_ERR_CUDA = -2,         /* CUDA runtime / driver error                                  */
  HJ_ERR_UNSUPPORTED = -3,  /* valid in the reference but outside this library's hot path   */
  HJ_ERR_STATE = -4,        /* call order (e.g. step before system/state is set)            */
  HJ_ERR_NAN = -5           /* NaN met in the integrated field (hji_solver.py:544)          */
} hj_status;

/* grid.bdry[d]: BoundaryCondition/add_ghost_extrapolate.py:16, add_ghost_periodic.py:12.
 * HJ_BC_HALO: the 3 ghost planes are stored in the field itself (slab decomposition, dim 0 only). */
typedef enum hj_bc { HJ_BC_EXTRAPOLATE = 0, HJ_BC_PERIODIC = 1, HJ_BC_HALO = 2 } hj_bc;

/* SpatialDerivative/upwind_first_weno5a.py:13-196 semantics. */
typedef enum hj_weno {
  HJ_WENO_AS_SHIPPED = 0, /* bug-compatible: aliasing at :97 => fixed-weight 5th-order upwind      */
  HJ_WENO_INTENDED = 1,   /* true WENO5 smoothness indicators + weights (ENO3bHelper.py:136-160)   */
  /* the other CoStateCalc functors of the reference, gather backend:                */
  HJ_SCHEME_ENO3A = 2,    /* upwindFirstENO3a / upwindFirstENO3  SpatialDerivative/upwind_first_eno3a.py:87-142 */
  HJ_SCHEME_ENO2 = 3      /* upwindFirstENO2                     SpatialDerivative/upwind_first_eno2.py:50-150  */
} hj_weno;

/* Compiled device functors for schemeData.hamFunc / schemeData.partialFunc.  Parameter blocks: see
 * INTEGRATION.md ("system parameter blocks").
 *   DUBINS_REL      DynamicalSystems/dubins_relative.py:63-111   3-D
 *   DOUBLE_INT      DynamicalSystems/double_integrator.py:49-89  2-D
 *   FLOCK           DynamicalSystems/flock.py:190-258 + bird.py:235-372 (also a lone Bird)  3-D
 *   DUBINS_REL_PAIR product of two DUBINS_REL on dims 0-2 / 3-5  6-D  
 *   DOUBLE_INT_PAIR product of two DOUBLE_INT on dims 0-1 / 2-3  4-D  
 *   GENERIC_DUBINS_CAR  genericHam / genericPartial (Hamiltonians/generic_ham.py:5, generic_partial.py:6) over a
 *                   device dynSys (csrc/hj_systems.cuh: GenericF<Dyn>), here the 3-D Dubins car
 *                   dx = (v cos x2 + d0, v sin x2 + d1, u + d2).  Its alpha depends on the derivative range of
 *                   the field (hj_deriv_range): the parameter block is refreshed before every RK stage.     */
typedef enum hj_system {
  HJ_SYS_NONE = 0,
  HJ_SYS_DUBINS_REL = 1,
  HJ_SYS_DOUBLE_INT = 2,
  HJ_SYS_FLOCK = 3,
  HJ_SYS_DUBINS_REL_PAIR = 4,
  HJ_SYS_DOUBLE_INT_PAIR = 5,
  HJ_SYS_GENERIC_DUBINS_CAR = 6
} hj_system;

/* Driver epilogue fused into the last RK stage (ValueFuncs/hji_solver.py:566-599). */
typedef enum hj_comp {
  HJ_COMP_NONE = 0,          /* 'set' / 'none'                                   */
  HJ_COMP_MIN_OVER_TIME = 1, /* 'minVOverTime': y = min(y, yLast)   :571-573     */
  HJ_COMP_MAX_OVER_TIME = 2, /* 'maxVOverTime'                                   */
  HJ_COMP_MIN_WITH_AUX = 3,  /* 'minVWithTarget' / 'minVWithV0': y = min(y, aux) */
  HJ_COMP_MAX_WITH_AUX = 4   /* 'maxVWithTarget' / 'maxVWithV0'                  */
} hj_comp;

typedef enum hj_field { HJ_FIELD_STATE = 0, HJ_FIELD_AUX = 1, HJ_FIELD_OBSTACLE = 2 } hj_field;

/* Kernel family used by hj_step / hj_rhs. */
typedef enum hj_backend {
  HJ_BACKEND_AUTO = 0,
  HJ_BACKEND_GATHER = 1, /* one thread per node, neighbours through L1/L2 (any shape)             */
  HJ_BACKEND_TMA = 2     /* TMA-staged shared-memory plane ring, streamed along dim D-3            */
} hj_backend;

/* Layout of the reduction record written by hj_rhs / hj_step (doubles):
 *   [0 .. D)      alphaMax_d  = max_x alpha_d        (artificial_diss_glf.py:104)
 *   [D .. 2D)     derivMin_d  = min(min L_d, min R_d) (:82-84)
 *   [2D .. 3D)    derivMax_d                          (:86-88)
 *   [3D]          nan flag (1.0 if any output was NaN, hji_solver.py:544)                         */
#define HJ_REDUCE_LEN(D) (3 * (D) + 1)

const char* hj_version(void);
const char* hj_last_error(void);

/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t hj_launch_count(void);

/* Grid + scheme.  Replaces the reads of grid.{dim,N,dx,bdry,bdryData} done by
 * upwindFirstENO3aHelper (SpatialDerivative/ENO3aHelper.py:57-64) and termLaxFriedrichs
 * (ExplicitIntegration/Term/term_lax_friedrich.py:91-97).
 *   N[d]              nodes along dim d of THIS context's slab (excluding halo planes)
 *   dx[d]             grid.dx[d]
 *   bc_kind[d]        hj_bc; HJ_BC_HALO only for d == 0
 *   bc_toward_zero[d] grid.bdryData[d].towardZero (extrapolate only)                              */
int hj_create(hj_ctx** out, int device, int ndim, const int64_t* N, const double* dx, const int* bc_kind,
              const int* bc_toward_zero, int weno_mode);
int hj_destroy(hj_ctx* ctx);
int hj_set_backend(hj_ctx* ctx, int backend);

/* grid.vs[d] (Grids/process_grid.py:204): the n == N[d] node coordinates of dim d, host pointer. */
int hj_set_axis(hj_ctx* ctx, int dim, const double* vs_host, int64_t n);

/* Host-evaluated 1-D table attached to dim `dim` (e.g. numpy cos/sin of grid.vs[2], so device values
 * are bit-identical to the reference's cp.cos(grid.xs[2]), dubins_relative.py:81-82).              */
int hj_set_table(hj_ctx* ctx, int slot, const double* tab_host, int64_t n);

/* schemeData.hamFunc / partialFunc -> device functor + scalar parameter block. */
int hj_set_system(hj_ctx* ctx, int system_id, const double* params, int nparams);

/* Resident fields.  `dense` is host memory (is_host != 0; pageable or pinned) or device memory. */
int hj_upload(hj_ctx* ctx, void* stream, int field, const double* dense, int is_host);
int hj_download(hj_ctx* ctx, void* stream, int field, double* dense, int is_host); /* synchronises if is_host */
int64_t hj_num_nodes(const hj_ctx* ctx);          /* prod N[d]                      */
int64_t hj_field_elems(const hj_ctx* ctx);        /* pitched elements incl. halos   */
/* Device pointer of RK buffer 0..2 (pitched layout, halo planes included): slab halo exchange, and on-device
 * initialisation of grids too large to stage on the host.  Taking buffer 0 marks the state as resident: the
 * caller owns its contents from then on. */
int hj_state_ptr(hj_ctx* ctx, int which_buffer, double** dev_ptr);
int64_t hj_plane_elems(const hj_ctx* ctx);        /* pitched elements of one dim-0 plane */

/* derivMin / derivMax of artificialDissipationGLF (artificial_diss_glf.py:82-88) on their own: per dim, the minimum and
 * the maximum over the grid of the upwind pair (derivL, derivR) of the context's CoStateCalc -- what genericPartial
 * (Hamiltonians/generic_partial.py:28-40) feeds the dynSys's get_opt_u / get_opt_v BEFORE the dissipation of the same RHS
 * exists.  y_dev: a dense device array, or NULL for the resident buffer RK stage `stage` (1..3; 4 = final stage of
 * odeCFL2) reads.  deriv_min /
 * deriv_max: D host doubles each.  One reduce-only kernel; synchronises. */
int hj_deriv_range(hj_ctx* ctx, void* stream, const double* y_dev, int stage, double* deriv_min, double* deriv_max);

/* upwindFirstWENO5a(grid, data, dim) -> (derivL, derivR)   SpatialDerivative/upwind_first_weno5a.py:13.
 * data/derivL/derivR: dense device arrays of hj_num_nodes doubles (HJ_BC_HALO contexts: not supported). */
int hj_deriv(hj_ctx* ctx, void* stream, const double* data_dev, int dim, double* derivL_dev, double* derivR_dev);

/* dL, dR, _ = upwindFirstENO3aHelper(grid, data, dim) -- SpatialDerivative/ENO3aHelper.py:11-191: the three left and
 * three right third-order ENO candidates (what upwindFirstWENO5a / upwindFirstENO3a return for generateAll = True,
 * upwind_first_weno5a.py:73-75).  out6_dev: 6 dense arrays of n doubles, [dL0, dL1, dL2, dR0, dR1, dR2].              */
int hj_deriv_candidates(hj_ctx* ctx, void* stream, const double* data_dev, int dim, double* out6_dev);

/* addGhostExtrapolate / addGhostPeriodic (dataIn, dim, width) -> dataOut, dense device arrays.
 * BoundaryCondition/add_ghost_extrapolate.py:16, add_ghost_periodic.py:12.                          */
int hj_add_ghost(hj_ctx* ctx, void* stream, const double* data_dev, int dim, int width, double* out_dev);

/* termLaxFriedrichs(t, y, schemeData) -> (ydot, stepBound)  ExplicitIntegration/Term/term_lax_friedrich.py:8
 * with dissFunc = artificialDissipationGLF (ExplicitIntegration/Dissipation/artificial_diss_glf.py:7).
 * y/ydot dense device arrays.  reduce_host[HJ_REDUCE_LEN(D)] and *step_bound are filled; synchronises. */
int hj_rhs(hj_ctx* ctx, void* stream, double t, const double* y_dev, double* ydot_dev, double* step_bound,
           double* reduce_host);

/* The single hooks of the reference's operator API on dense device arrays, for a host that keeps the
 * reference's termLaxFriedrichs and swaps one callable at a time; the fused stage kernels do not use them.
 *   hj_ham       ham = hamFunc(t, data, derivC, schemeData): deriv_c_dev[d] = derivC of dim d (n doubles each)
 *                (DynamicalSystems/dubins_relative.py:63-88, double_integrator.py:49-74, bird.py:266-316, flock.py:190-233)
 *   hj_alpha     alpha = partialFunc(t, data, derivMin, derivMax, schemeData, dim) as a dense array of n doubles (every
 *                registered system's alpha is state-only; dubins_relative.py:90-111, double_integrator.py:76-89)
 *   hj_diss_glf  diss, stepBound = artificialDissipationGLF(t, data, derivL, derivR, schemeData)
 *                (ExplicitIntegration/Dissipation/artificial_diss_glf.py:7-111): diss = sum_d 0.5 (R_d - L_d) alpha_d in the
 *                reference's order; reduce_host (optional) receives HJ_REDUCE_LEN(D) doubles like hj_rhs; synchronises. */
int hj_ham(hj_ctx* ctx, void* stream, double t, const double* const* deriv_c_dev, double* ham_dev);
int hj_alpha(hj_ctx* ctx, void* stream, double t, int dim, double* alpha_dev);
int hj_diss_glf(hj_ctx* ctx, void* stream, double t, const double* const* deriv_l_dev, const double* const* deriv_r_dev,
                double* diss_dev, double* step_bound, double* reduce_host);

/* max_x alpha_d for d = 0..D-1 without touching a field (state-only partialFunc); synchronises.
 * stepBound = 1 / sum_d alpha_max[d] / dx[d]  (artificial_diss_glf.py:104-109).                       */
int hj_alpha_max(hj_ctx* ctx, void* stream, double t, double* alpha_max_host, double* step_bound);

/* One TVD-RK3 step of the resident state with a caller-chosen dt: the loop body of odeCFL3
 * (ExplicitIntegration/Integration/ode_cfl_3.py:125-251) as three fused stage kernels, plus the driver
 * epilogue of hji_solver.py:566-644 fused into stage 3.
 *   stage_params  NULL, or 3*nparams doubles: the system parameter block for each of the three RHS
 *                 evaluations (Flock mutates its headings on every hamFunc call, flock.py:213).
 *   comp          hj_comp;  use_obstacle != 0: y = max(y, -obstacle) afterwards (:641-644, pointwise).
 *   want_reduce   != 0: also produce the per-stage reduction records (device side, no sync).
 * Asynchronous: nothing is read back.                                                                 */
int hj_step(hj_ctx* ctx, void* stream, double t, double dt, const double* stage_params, int comp, int use_obstacle,
            int want_reduce);
/* The three reduction records of the last hj_step(want_reduce=1): 3*HJ_REDUCE_LEN(D) doubles; synchronises. */
int hj_step_reductions(hj_ctx* ctx, void* stream, double* reduce_host);

/* Multi-GPU slab support: run stage `stage` (1..3) only, so the caller can exchange halos between stages.
 * hj_step == hj_stage(1); hj_stage(2); hj_stage(3) on a single device.
 * stage 4 (not on slab contexts): the final stage of odeCFL2 (ode_cfl_2.py: y = 0.5 (y + (y1 + dt f(y1)))), so that
 * hj_step_rk2 == hj_stage(1); hj_stage(4) -- for systems whose parameter block has to be refreshed between the two
 * RHS evaluations from the field itself (genericPartial, with hj_deriv_range(stage) before each).          */
int hj_stage(hj_ctx* ctx, void* stream, int stage, double t, double dt, const double* params, int comp,
             int use_obstacle, int want_reduce);
/* Product systems on the dimension-split path: a stage is two kernels.  Pass 1 (trailing dim block)
 * reads no dim-0 halo plane, so a slab job posts its halo exchange, runs pass 1 under it, and runs pass 2 (leading
 * dim block + stage algebra) once the halos have landed.  hj_stage == hj_stage_pass(1); hj_stage_pass(2).
 * hj_is_split: 1 if this context (system and state set) advances its system that way, else 0.                 */
int hj_stage_pass(hj_ctx* ctx, void* stream, int stage, int which_pass, double t, double dt, const double* params,
                  int comp, int use_obstacle, int want_reduce);
int hj_is_split(const hj_ctx* ctx);
/* Pass 2 in pieces: the trailing dims of a product system are one flattened "vector" axis without stencil (length
 * *vlen doubles, the pitched stride of the last leading dim); hj_stage_pass_cols runs pass 2 of `stage` on columns
 * [col_begin, col_end) of it only (multiples of *quantum; col_end may be *vlen).  Disjoint pieces covering the axis
 * equal hj_stage_pass(2) bit for bit.  A slab job pushes each finished piece of its edge planes to the neighbours
 * (hj_halo_push with the same columns) while the next piece is computed.  want_reduce as for hj_stage_range.      */
int hj_split_cols(hj_ctx* ctx, int64_t* vlen, int* quantum);
int hj_stage_pass_cols(hj_ctx* ctx, void* stream, int stage, int64_t col_begin, int64_t col_end, double t, double dt,
                       const double* params, int comp, int use_obstacle, int want_reduce);
/* Whole (non-product) systems on the plane-ring backend: run stage `stage` on planes [z_begin, z_end) of the marched
 * dim D-3 only (dim 0 of a 3-D grid, i.e. the slab dim).  A slab job posts its halo exchange, advances the planes
 * whose stencil stays inside the slab ([3, N0-3)) under it, and advances the two 3-plane edge ranges once the halos
 * have landed.  The union of disjoint ranges covering [0, N[D-3]) equals hj_stage bit for bit.  as_shipped WENO only.
 * want_reduce: 0 none, 1 reset the stage's reduction record then accumulate, 2 accumulate into it.             */
int hj_stage_range(hj_ctx* ctx, void* stream, int stage, int64_t z_begin, int64_t z_end, double t, double dt,
                   const double* params, int comp, int use_obstacle, int want_reduce);
/* Which internal buffer (0..2) stage `stage` reads with its stencil (needs valid halos) / writes. */
int hj_stage_io(const hj_ctx* ctx, int stage, int* in_buffer, int* out_buffer);
/* intended-WENO only: per-dim max(D1^2) prepass of buffer `buf` into the context's eps record
 * (upwind_first_weno5a.py:154-156).  eps_dev returns the device address of the D raw maxima
 * (ordered-uint64 encoded) so a slab job can max-allreduce them before the stage.                       */
int hj_eps_prepass(hj_ctx* ctx, void* stream, int buf, uint64_t** eps_dev);

/* Slab edge ranks: fill the 3 stored halo planes of buffer `buf` on `side` (0 = below plane 0, 1 = above the last
 * plane) with addGhostExtrapolate ghosts of the slab's own edge planes (add_ghost_extrapolate.py:88-110), for a
 * global dim-0 boundary that is not periodic.  Interior slab faces get their halos from the neighbour instead. */
int hj_fill_edge_halo(hj_ctx* ctx, void* stream, int buf, int side);

/* Slab halos over NVLink peer memory.  The reference has no multi-GPU path: the planes a
 * rank receives are the ghost planes addGhostPeriodic / the interior of a larger array would have supplied
 * (BoundaryCondition/add_ghost_periodic.py:78-87, SpatialDerivative/ENO3aHelper.py:64).  A slab context (dim 0 =
 * HJ_BC_HALO) PUSHES its 3 edge planes into its neighbours' stored halo planes with the copy engines -- no NCCL, no
 * SM copy kernel -- so any host that can move HJ_HALO_DESC_BYTES bytes between ranks can run slabs:
 *   hj_halo_export   fills `desc` (CUDA IPC handles of the three RK buffers and of the arrival counters; POD bytes).
 *   hj_halo_attach   maps the lower / upper neighbour's descriptor (NULL = no neighbour on that side; contexts of one
 *                    process are attached by pointer).  For a two-rank periodic ring pass the same descriptor twice.
 *   hj_halo_push     ordered behind `stream`: copy my edge planes of RK buffer `buf` into the halos of the neighbours in
 *                    `sides` (bit 0 lower, bit 1 upper; 3 = both) on dedicated copy streams, then bump their arrival
 *                    counters.  col_begin/col_end/row_len select
 *                    columns of every row_len-element row of the planes (pushing a buffer in pieces while later
 *                    pieces are still being computed); 0, 0, 0 = whole planes.
 *   hj_halo_wait     makes `stream` wait until `npush` further pushes of `buf` from each neighbour have landed and
 *                    my own outbound copies have left.  Every rank must push and await each buffer equally often.
 *   hj_halo_attached bit 0 / bit 1: a lower / upper neighbour is attached.
 *   hj_halo_set_fused  product systems on the dimension-split path: pass 2 (hj_stage_pass(2) / hj_stage_pass_cols) stores
 *                    the nodes of its three lowest / highest dim-0 planes ALSO into the halo planes of the lower / upper
 *                    neighbour (`sides` bit 0 / bit 1; 0 = off) of the buffer it writes -- one kernel computes the stage
 *                    and moves its halo over NVLink (peer stores), the transfer overlaps the kernel tile by tile.  The
 *                    host then calls hj_halo_signal instead of hj_halo_push for those sides.  With one side fused and
 *                    the other pushed by the copy engines the two transports share the link.
 *   hj_halo_signal   ordered behind `stream`: bump the arrival counters of `buf` at the neighbours in `sides` (after a
 *                    fused pass 2).                                                                                */
#define HJ_HALO_DESC_BYTES 512
int hj_halo_export(hj_ctx* ctx, void* desc);
int hj_halo_attach(hj_ctx* ctx, const void* lower_desc, const void* upper_desc);
int hj_halo_detach(hj_ctx* ctx);
int hj_halo_attached(const hj_ctx* ctx);
int hj_halo_set_fused(hj_ctx* ctx, int sides);
int hj_halo_signal(hj_ctx* ctx, void* stream, int buf, int sides);
int hj_halo_push(hj_ctx* ctx, void* stream, int buf, int sides, int64_t col_begin, int64_t col_end, int64_t row_len);
int hj_halo_wait(hj_ctx* ctx, void* stream, int buf, int npush);

/* odeCFL3(schemeFunc, [t, t_end], y, options{factorCFL,maxStep,singleStep='on'}, schemeData)
 * ExplicitIntegration/Integration/ode_cfl_3.py:11 for one CFL-limited step on a dense array that may live
 * on the host (is_host) -- upload, dt = min(factorCFL*stepBound, t_end-t, maxStep) (:142-143), step,
 * download.  Synchronises.  This is the "host buffers in, host buffers out" call the e2e number times.
 * Systems whose parameter block changes between the three RHS evaluations (HJ_SYS_FLOCK) are refused
 * (HJ_ERR_UNSUPPORTED): step them with hj_upload + hj_step(stage_params) + hj_download.                    */
int hj_ode_cfl3_single(hj_ctx* ctx, void* stream, double t, double t_end, double factor_cfl, double max_step,
                       double* y_inout, int is_host, int comp, int use_obstacle, double* t_new, double* dt_out);
/* The same call out of place, as odeCFL3 returns a fresh y (ode_cfl_3.py:11: `t, y, schemeData = odeCFL3(...)`):
 * reads y_in, writes y_out (may alias).  With pinned host buffers (hj_host_alloc) on a 3-D grid the step is a software
 * pipeline -- chunked H2D, a wavefront of stage launches on plane ranges, chunked D2H -- bit-identical to the
 * resident step.  This is what levelsetpy_b200.odeCFL3(..., singleStep='on') calls for a host array.               */
int hj_ode_cfl3_step(hj_ctx* ctx, void* stream, double t, double t_end, double factor_cfl, double max_step,
                     const double* y_in, double* y_out, int is_host, int comp, int use_obstacle, double* t_new,
                     double* dt_out);
/* Chunk height (dim-0 planes) of that software pipeline; 0 = the library's default.  A tuning knob: results do not depend
 * on it (every chunking is bit-identical to the resident step). */
int hj_set_pipeline_planes(hj_ctx* ctx, int planes);
/* Pinned (page-locked) host memory for those buffers. */
int hj_host_alloc(int64_t bytes, void** out);
int hj_host_free(void* p);

/* Driver epilogues of HJIPDE_solve beyond the fused min / max (ValueFuncs/hji_solver.py), on the resident state:
 *   hj_snapshot   keeps a device copy of the state (the frame at tau[i-1]; :509-533).
 *   hj_change     max |state - snapshot| (stopConverge's "change", :661-672) and a NaN flag of the state (the check of
 *                 :544) from one device reduction; synchronises; no field leaves the device.
 *   hj_discount   discounting after a step: mode 0 (:603-611) y = gamma y + (1 - gamma) l with l = the HJ_FIELD_AUX field
 *                 (target, or data0); mode 1 ("Kene", :615-637) y = gamma (y - m) min|max (l - m) + m with m = max_val =
 *                 max |l| (take_max selects max: maxVWithL); mode 2: the obstacle mask on its own, y = max(y, -obstacle)
 *                 (:641-644), for when discounting sits between the compMethod epilogue and the mask.                */
int hj_snapshot(hj_ctx* ctx, void* stream);
int hj_change(hj_ctx* ctx, void* stream, double* max_abs_change, int* has_nan);
int hj_discount(hj_ctx* ctx, void* stream, double gamma, int mode, int take_max, double max_val);

/* termRestrictUpdate (ExplicitIntegration/Term/term_restrict_update.py:56-96): restrict the sign of the update,
 * ydot = max(ydot, 0) (sign > 0, schemeData.positive true) or min(ydot, 0) (sign < 0), fused into every stage kernel
 * and into hj_rhs; sign = 0 switches it off. */
int hj_set_restrict(hj_ctx* ctx, int sign);

/* odeCFL2 (ExplicitIntegration/Integration/ode_cfl_2.py): one step of the second-order TVD Runge-Kutta scheme on the
 * resident state, y1 = y + dt f(y); y = 0.5 (y + (y1 + dt f(y1))), as two fused stage kernels; arguments as hj_step
 * (stage_params: 2 parameter blocks or NULL). */
int hj_step_rk2(hj_ctx* ctx, void* stream, double t, double dt, const double* stage_params, int comp, int use_obstacle,
                int want_reduce);

/* Batch contexts (many small independent grids, e.g. one 101^3 grid per Flock).
 * nbatch grids of identical shape / dx / boundary kinds share one context: fields are [nbatch, N0, N1, N2]
 * (hj_upload / hj_download move all of them at once), every grid has its own system parameter block and its own
 * dt, and one launch per RK stage advances the whole batch.  Replaces the reference's Python loop over grids, each
 * iteration of which is a full odeCFL3 call (ode_cfl_3.py:11) on an L2-sized field.  3-D grids, HJ_SYS_FLOCK,
 * as_shipped WENO, TMA backend.  hj_set_system(ctx, HJ_SYS_FLOCK, NULL, nparams) registers the block length;
 * hj_step_batch takes dt[nbatch] and the per-stage blocks [3][nbatch][nparams] (flock.py:213 re-derives the
 * headings on each of the three RHS evaluations) from host memory. */
int hj_create_batch(hj_ctx** ctx, int device, int nbatch, int ndim, const int64_t* N, const double* dx,
                    const int* bc_kind, const int* bc_toward_zero, int weno_mode);
int hj_step_batch(hj_ctx* ctx, void* stream, const double* dt_host, const double* stage_params_host, int comp,
                  int use_obstacle);
int hj_batch_size(const hj_ctx* ctx);


/* Plain device-memory helpers so a host language without its own CUDA binding can drive the dense-array entry
 * points (the Python shim uses them when it is handed numpy arrays).  kind: 1 = H2D, 2 = D2H, 3 = D2D.       */
int hj_device_count(void);
int hj_dev_alloc(int device, int64_t bytes, void** out);
int hj_dev_free(void* p);
int hj_memcpy(void* dst, const void* src, int64_t bytes, int kind, void* stream, int sync);
int hj_stream_sync(void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HJB200_H */
