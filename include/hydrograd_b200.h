/* hydrograd_b200.h -- C ABI of the B200-native 2-D shallow-water RHS / VJP path.
 *
 * This is the drop-in boundary for ONE hot path of psu-efd/Hydrograd.jl: the function
 *
 *     swe_2d_rhs(dQdt, Q, params_vector, t, p_extra)
 *         src/fvm/discretization/semi_discretize_swe_2D.jl:18-277
 *
 * that src/applications/solve_swe_2D.jl:281-307 wraps into the ODEFunction handed to
 * OrdinaryDiffEq / SciMLSensitivity, its reverse-mode derivative (what Zygote derives from it,
 * contract at src/utilities/debug_AD.jl:60,75), and the hand-rolled explicit Euler stepper
 * src/ode_solvers/custom_ODE_solvers.jl:5-95.
 *
 * The reference has no FFI of its own (it is 100 % Julia); the entry points below are what a Julia
 * `@ccall` shim binds (hydrograd.jl_b200/julia/HydrogradB200.jl).  Conventions:
 *   - plain pointers and sizes only; every function returns 0 on success, non-zero on error;
 *     the message is available from hg_last_error(ctx) (or hg_last_error(NULL) for hg_create).
 *   - the caller owns every buffer it passes; the library copies what it needs during the call and
 *     never keeps a host pointer.  Host-pointer calls are synchronous.
 *   - one caller per ctx at a time (the reference is single threaded); different ctxs are independent.
 *   - fp64 everywhere; ids are int64 as in Julia; `index_base` says whether ids start at 1 (Julia) or 0.
 *   - state vector layout Q = [xi(1:N); q_x(1:N); q_y(1:N)]  (semi_discretize_swe_2D.jl:93-95).
 *   - there is NO CPU fallback: every compute entry point needs a CUDA device (sm_100a build).
 */
#ifndef HYDROGRAD_B200_H
#define HYDROGRAD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HG_ABI_VERSION 3

#if defined(__GNUC__)
#define HG_API __attribute__((visibility("default")))
#else
#define HG_API
#endif

typedef struct hg_ctx hg_ctx;

/* active_param_name of SWE2D_Extra_Parameters (src/applications/application_commons.jl:9):
 * which model parameter `params_vector` carries (semi_discretize_swe_2D.jl:114,153,190). */
enum {
  HG_PARAM_NONE = 0,     /* forward simulation: params_vector unused                          */
  HG_PARAM_ZB = 1,       /* "zb":       params = zb_cells[N]        -> update_bed_data        */
  HG_PARAM_MANNING = 2,  /* "ManningN": params = n per zone[n_mat]  -> n_cells = p[matID+1]   */
  HG_PARAM_Q = 3,        /* "Q":        params = inletQ_TotalQ[n_inletq]                      */
  HG_PARAM_UDE = 4       /* bPerform_UDE: params = the network parameters [hg_ude_desc.n_params] of the
                            model set by hg_set_ude_model (semi_discretize_swe_2D.jl:165-178)  */
};

/* error codes */
enum {
  HG_OK = 0,
  HG_ERR_ARG = 1,         /* bad argument / wrong array length / inconsistent mesh             */
  HG_ERR_CUDA = 2,        /* CUDA runtime failure or no sm_100 device                          */
  HG_ERR_CONVEYANCE = 3,  /* inlet-q total conveyance <= 1e-10 (bc_2D.jl:678-680 assert)       */
  HG_ERR_SOLVER = 4,      /* unknown Riemann solver (semi_discretize_swe_2D.jl:356-361)        */
  HG_ERR_STATE = 5,       /* call sequence error (e.g. device state never set)                 */
  HG_ERR_COMM = 6         /* library-owned halo exchange: a neighbour's push did not arrive in time */
};

/* ---- mesh_2D (src/meshes/mesh_2D.jl:2-71): the fields swe_2d_rhs reads, flattened. -------------
 * N x ld tables are column-major like Julia's `cellFacesList` (entry (i,j) at i + N*j, 0-based i,j).
 * cell_neighbors / cell_normals are `cellNeighbors_Dict` / `cell_normals` written into the same
 * N x ld shape by the shim.  cell j-th face of cell i:  face id  cell_faces[i + N*j],
 * neighbour cell id -- or GHOST id when face_is_boundary[face] -- cell_neighbors[i + N*j],
 * outward unit normal (cell_normals[i + N*(j + ld*0)], cell_normals[i + N*(j + ld*1)]).          */
typedef struct {
  int64_t n_cells;                 /* numOfCells                                                 */
  int64_t n_faces;                 /* numOfFaces                                                 */
  int64_t n_ghost;                 /* numOfAllBounaryFaces (= number of ghost cells)             */
  int64_t ld;                      /* second dim of the tables (8 = gMax_Nodes_per_Element)      */
  int32_t index_base;              /* 1: ids as in Julia, 0: C ids                               */
  const int64_t* cell_nfaces;      /* [N]       cellNodesCount                                   */
  const int64_t* cell_faces;       /* [N*ld]    cellFacesList (sign ignored)                     */
  const int64_t* cell_neighbors;   /* [N*ld]    cellNeighbors_Dict                               */
  const double* cell_normals;      /* [N*ld*2]  cell_normals                                     */
  const uint8_t* face_is_boundary; /* [F]       bFace_is_boundary                                */
  const double* face_lengths;      /* [F]       face_lengths                                     */
  const double* cell_areas;        /* [N]       cell_areas                                       */
  const double* cell_centroids;    /* [N*2] column-major cell_centroids, or NULL (only used to
                                      order cells for locality / to partition across GPUs)      */
} hg_mesh_desc;

/* ---- BoundaryConditions2D (src/fvm/boundary_conditions/bc_2D.jl:3-47), static part, flattened.
 * Boundaries are listed in the reference's processing order: all inlet-q, then exit-h, wall, symm
 * (bc_2D.jl:279-295); entry arrays are the per-boundary vectors concatenated in that order.       */
typedef struct {
  int64_t n_inletq, n_exith, n_wall, n_symm; /* nInletQ_BCs, nExitH_BCs, nWall_BCs, nSymm_BCs    */
  const int64_t* bc_ptr;           /* [n_bc+1] 0-based offsets of each boundary's entries        */
  const int64_t* ghost_ids;        /* [B] *_ghostCellIDs            (all_boundary_ghost_ids)     */
  const int64_t* internal_cells;   /* [B] *_internalCellIDs                                      */
  const double* outward_normals;   /* [B*2] column-major *_faceOutwardNormals / *_outwardNormals */
  const double* face_lengths;      /* [B] inletQ_Length (read for inlet-q entries only)          */
  /* ---- multi-GPU extension (no counterpart in the reference, which is a single serial process): after the
   * symm boundaries come n_halo "halo boundaries", one per neighbouring rank.  Their entries are the faces
   * this rank shares with that neighbour, in an order both ranks agree on (sorted by the global ids of the
   * two cells); the ghost cell of such an entry mirrors the neighbour's cell and is refreshed through the
   * halo buffers (hg_halo_*) before each RHS.  NULL / 0 on a single GPU.                             */
  int64_t n_halo;
  const uint8_t* halo_flip;        /* [B] 1 where the ghost (remote) cell has the smaller GLOBAL id: the face is
                                      then evaluated with the remote cell as L so that every rank computes the
                                      same bits as a single-GPU run                                     */
  const double* halo_area;         /* [B] area of the remote cell (VJP needs lambda/area of both sides)   */
} hg_bc_desc;

/* ---- SWE2D_Extra_Parameters (src/applications/application_commons.jl:7-44) + swe_2D_consts
 * (src/constants/swe_2D_constants.jl:4-17): the frozen fields the RHS reads.                     */
typedef struct {
  double g, k_n, h_small;          /* swe_2D_constants.g, .k_n, .h_small                         */
  const char* riemann_solver;      /* swe_2D_constants.RiemannSolver; only "Roe" exists          */
  const double* hstill;            /* [N]                                                        */
  const double* hstill_ghost;      /* [B]   hstill_ghostCells                                    */
  const double* zb_cells;          /* [N]                                                        */
  const double* zb_ghost;          /* [B]   zb_ghostCells                                        */
  const double* S0_cells;          /* [2N]  N x 2 column-major                                   */
  const double* ManningN_cells;    /* [N]                                                        */
  const int64_t* matID_cells;      /* [N]   srh_all_Dict["matID_cells"], 0-based zone ids; may be
                                            NULL when HG_PARAM_MANNING is never used             */
  int64_t n_mat;                   /* number of Manning zones (length of the ManningN params)    */
  const double* inletQ_TotalQ;     /* [n_inletq]                                                 */
  const double* exitH_WSE;         /* [n_exith]                                                  */
} hg_fields_desc;

typedef struct {
  int32_t device;                  /* CUDA device ordinal                                        */
  int32_t tile_cells;              /* cells per CTA tile of the fused kernel; 0 = default        */
  int32_t reorder;                 /* 1 = renumber cells for locality (needs cell_centroids)     */
  int32_t strict;                  /* 1 = reference evaluation order, no FMA contraction         */
  int32_t path;                    /* 0 = fused tile kernel, 1 = plain 3-kernel path (ghost/face/cell) */
  int32_t reserved[11];            /* reserved[0]: threads per CTA of the fused kernel (0 = default), tuning only;
                                      reserved[1]: chunks of the host-buffer pipeline of hg_rhs / hg_rhs_vjp (0 = default), tuning only;
                                      reserved[2]: launch-shape variant of the VJP kernel (0 = default), tuning only;
                                      reserved[3]: L2 prefetch distance in tiles (0 = one residency ahead, -1 = off);
                                      reserved[4]: 1 = do not regroup a tile's faces into bank-conflict-free blocks of 16;
                                      reserved[5]: 1 = run the UDE network through the generic (run-time shape) kernels */
} hg_options;

/* Manning's n as a function of the state, the reference's forward-simulation option ManningN_option = "variable"
 * (semi_discretize_swe_2D.jl:140-149; closures of parameters/process_ManningN_2D.jl:119-213). */
enum hg_manning_function {
  HG_MANNING_CONSTANT = 0,   /* ManningN_cells as bound (default)                                                  */
  HG_MANNING_POWER_LAW = 1,  /* n = n_lower + (n_upper - n_lower) (h + eps)^(-k)                                    */
  HG_MANNING_SIGMOID = 2,    /* n = n_lower + (n_upper - n_lower) / (1 + exp(k (h - h_mid)))                        */
  HG_MANNING_INVERSE = 3,    /* n = n_lower + (n_upper - n_lower) / (1 + k h)                                       */
  HG_MANNING_H_UMAG_KS = 4   /* Cheng (2008) friction factor f(Re, h/ks) -> n = sqrt(f/8) h^(1/6) / sqrt(9.81)      */
};

/* fills *opt with the defaults */
HG_API void hg_default_options(hg_options* opt);

HG_API int hg_abi_version(void);

/* Build a context (device copies of the mesh in the internal face/tile layout).  Replaces the
 * per-call unpacking of p_extra at semi_discretize_swe_2D.jl:26-67.                              */
HG_API int hg_create(hg_ctx** out, const hg_mesh_desc* mesh, const hg_bc_desc* bc,
              const hg_fields_desc* fields, const hg_options* opt);
HG_API void hg_destroy(hg_ctx* ctx);
HG_API const char* hg_last_error(const hg_ctx* ctx);

HG_API int64_t hg_n_cells(const hg_ctx* ctx);

/* Replace frozen fields after creation (setup_* in solve_swe_2D.jl:148-221); NULL = keep.        */
HG_API int hg_set_fields(hg_ctx* ctx, const double* ManningN_cells, const double* zb_cells,
                  const double* zb_ghost, const double* S0_cells, const double* inletQ_TotalQ,
                  const double* exitH_WSE);

/* Selects the Manning's n closure evaluated inside every RHS from the clamped (h, q) of each cell; params = {n_lower,
 * n_upper, k, h_mid}; ks_cells[N] (reference order) is the roughness height per cell, needed by HG_MANNING_H_UMAG_KS
 * (ks per material zone gathered through matID, process_ManningN_2D.jl:56-60).  Forward simulation only, like the
 * reference: the VJP / adjoint / ensemble entry points and active = HG_PARAM_MANNING return HG_ERR_ARG while a closure is set. */
HG_API int hg_set_manning_function(hg_ctx* ctx, int32_t type, const double* params, const double* ks_cells);

/* ---- UDE: Manning's n of every cell from a small neural network of the cell's own state, the reference's
 * settings.bPerform_UDE with UDE_choice "ManningN_h" / "ManningN_h_Umag_ks" (semi_discretize_swe_2D.jl:165-178 ->
 * update_ManningN_UDE, parameters/process_ManningN_2D.jl:216-272).  The network is the one create_NN_model builds
 * (UDE/process_UDE.jl:2-65): per hidden layer Dense(in, width, activation) followed by LayerNorm(width), then
 * Dense(in, 1) and the bounded output  n = lo + (hi - lo) sigmoid(z).  Inputs are (h, |U|, ks) of the clamped state,
 * each mapped to [-1, 1] by 2 (x - lo) / (hi - lo) - 1 (process_ManningN_2D.jl:233-240).  Lux (third party, Lux = "1.2.3",
 * Project.toml:88) is absent from this image, so its layers are restated from their published definitions:
 *   Dense      y = activation(W x + b), W [out x in] column-major;
 *   LayerNorm  y = (x - mean) / sqrt(var + epsilon) * scale + bias, var uncorrected, epsilon = 1f-5, scale / bias [width].
 *              Which entries `mean` and `var` run over is Lux's `dims`: the reference passes none, and Lux's default
 *              `dims = Colon()` is documented as "over the whole input array" -- every hidden unit of EVERY cell of the
 *              batch (HG_LN_WHOLE_ARRAY; n of one cell then depends on all cells through two scalars per layer, which
 *              the kernels obtain with deterministic reductions).  HG_LN_PER_CELL normalises over the hidden units of
 *              each cell (Lux `dims = 1`, the textbook layer norm).  The shim picks the mode from the layer's `dims`.
 * With active_param = HG_PARAM_UDE, params_vector is the flat parameter vector theta (a ComponentArray of the Lux
 * parameters); the off_* fields say where each array starts in it, so any flattening order works.  The VJP then returns
 * pbar = d(lambda . rhs)/d theta and Qbar includes the path through n(Q).  (UDE_choice "FlowResistance",
 * semi_discretize_swe_2D.jl:517-532, is not built.) */
#define HG_UDE_MAX_HIDDEN 3
#define HG_UDE_MAX_WIDTH 8
enum { HG_UDE_MANNING_H = 1, HG_UDE_MANNING_H_UMAG_KS = 2 };
enum { HG_ACT_IDENTITY = 0, HG_ACT_RELU = 1, HG_ACT_LEAKYRELU = 2, HG_ACT_SIGMOID = 3, HG_ACT_TANH = 4, HG_ACT_SOFTPLUS = 5 };
enum { HG_LN_NONE = 0, HG_LN_PER_CELL = 1, HG_LN_WHOLE_ARRAY = 2 };
typedef struct {
  int32_t choice;                            /* HG_UDE_MANNING_H (input h) or HG_UDE_MANNING_H_UMAG_KS (h, |U|, ks) */
  int32_t n_hidden;                          /* hidden_layers: 1 .. HG_UDE_MAX_HIDDEN                               */
  int32_t width[HG_UDE_MAX_HIDDEN];          /* units per hidden layer, 1 .. HG_UDE_MAX_WIDTH                       */
  int32_t activation[HG_UDE_MAX_HIDDEN];     /* HG_ACT_* of each hidden Dense (get_activation, process_UDE.jl:91-105) */
  int32_t layernorm;                         /* HG_LN_*                                                             */
  double ln_epsilon;                         /* Lux default 1f-5 = 9.999999747378752e-06                            */
  double h_bounds[2], umag_bounds[2], ks_bounds[2], output_bounds[2];   /* UDE_NN_config                           */
  int64_t n_params;                          /* length of theta                                                     */
  int64_t off_weight[HG_UDE_MAX_HIDDEN + 1]; /* 0-based offsets into theta; entry n_hidden = the output Dense       */
  int64_t off_bias[HG_UDE_MAX_HIDDEN + 1];
  int64_t off_ln_scale[HG_UDE_MAX_HIDDEN];   /* ignored with HG_LN_NONE                                             */
  int64_t off_ln_bias[HG_UDE_MAX_HIDDEN];
} hg_ude_desc;

/* Sets (desc != NULL) or clears (desc == NULL) the UDE model.  ks_cells[N] (reference order) is needed by
 * HG_UDE_MANNING_H_UMAG_KS (process_ManningN_2D.jl:48-60).  While a model is set, HG_PARAM_UDE is the only active
 * parameter accepted besides NONE (which then evaluates the RHS with the bound ManningN_cells, no network).  Not
 * available on the strict path, for ensembles, or -- HG_LN_WHOLE_ARRAY only -- on multi-rank contexts. */
HG_API int hg_set_ude_model(hg_ctx* ctx, const hg_ude_desc* desc, const double* ks_cells);

/* dQdt = swe_2d_rhs(Q, params, t)   -- host buffers, reference cell order.
 * semi_discretize_swe_2D.jl:18-277.  `t` is accepted and unused exactly like the reference.      */
HG_API int hg_rhs(hg_ctx* ctx, const double* Q, const double* params, int64_t n_params,
           int32_t active_param, double t, double* dQdt);

/* Vector-Jacobian product of the same call: Qbar = (d rhs/dQ)^T lambda  [3N],
 * pbar = (d rhs/dparams)^T lambda  [n_params] (NULL allowed when active_param = NONE),
 * ncell_bar (optional, may be NULL) = d(lambda . rhs)/d ManningN_cells  [N]  (UDE hook).
 * Replaces Zygote.pullback on swe_2d_rhs (debug_AD.jl:60,75; swe_2D_inversion.jl:339).
 * Q = NULL: differentiate at the state that is resident on the device -- the one the last hg_rhs / hg_set_state call put
 * there.  That is the shape of a Zygote pullback: `y, back = Zygote.pullback(swe_2d_rhs, Q, p)` evaluates the primal
 * once (hg_rhs uploads Q) and `back(lambda)` needs only the cotangent, so the pullback ships 24 instead of 48 bytes per
 * cell over PCIe.  hg_state_generation guards the reuse: read it after the forward call, pass Q = NULL only while it is
 * unchanged (HG_ERR_STATE when nothing is resident).                                               */
HG_API int hg_rhs_vjp(hg_ctx* ctx, const double* Q, const double* params, int64_t n_params,
               int32_t active_param, double t, const double* lambda, double* Qbar, double* pbar,
               double* ncell_bar);

/* ---- device-resident state (no host round trip per stage) ---------------------------------- */
/* Forward mode: dQdt_dot = J_Q v + J_p pdot (and dQdt when not NULL) -- one partial of a ForwardDiff.Dual pass through
 * swe_2d_rhs (swe_2D_sensitivity.jl:80 wraps the whole solve in ForwardDiff.jacobian; ForwardDiffSensitivity /
 * ForwardSensitivity of solve_swe_2D.jl:230-235).  v[3N]; pdot[n_params] or NULL (zero).  Fused contexts run the forward-mode
 * tile kernel (hg_fjvp.cu: each face once on the staged tile, K directions per launch); contexts created with strict = 1 run
 * on the plain tables in the reference's evaluation order.  No state-dependent Manning closure, no UDE network, one GPU.   */
HG_API int hg_rhs_jvp(hg_ctx* ctx, const double* Q, const double* params, int64_t n_params, int32_t active_param, double t,
               const double* v, const double* pdot, double* dQdt, double* dQdt_dot);
/* K directions at once (one ForwardDiff chunk): V[K][3N], Pdot[K][n_params] or NULL, JV[K][3N]; dQdt may be NULL */
HG_API int hg_rhs_jvp_multi(hg_ctx* ctx, const double* Q, const double* params, int64_t n_params, int32_t active_param, double t,
                     int64_t K, const double* V, const double* Pdot, double* dQdt, double* JV);
/* Forward sensitivity solve = the reference's sensitivity driver (swe_2D_sensitivity.jl:34-80: ForwardDiff.jacobian around
 * solve(prob, Tsit5(), adaptive, dt; abstol, reltol)) on the device: values and one partial per entry of the active
 * parameter (zb, ManningN or Q) advance together, the error estimate includes the partials like DiffEqBase's norm of Dual
 * numbers, PI controller with the powers of hg_set_controller_pow.  t_save[n_save] / Q_save[n_save][3N] (n_save may be 0):
 * the VALUES at the driver's save times by Tsit5's dense output, i.e. the columns of forward_simulation_results.json
 * (swe_2D_sensitivity.jl:60-72).  Q_T[3N] (may be NULL); S[n_params][3N], row k =
 * d Q(t1) / d p_k (the transpose of the reference's 3N x n_params Jacobian = its column-major JSON layout); stats =
 * {accepted, rejected, augmented RHS evaluations}; hg_last_steps afterwards gives the accepted steps.  Fused and strict
 * contexts (the K partials of a stage go through one launch of the respective forward-mode kernel).                         */
HG_API int hg_solve_tsit5_sens(hg_ctx* ctx, const double* Q0, const double* params, int64_t n_params, int32_t active_param,
                        double t0, double t1, double dt, int32_t adaptive, double abstol, double reltol, const double* t_save,
                        int64_t n_save, double* Q_save, double* Q_T, double* S, int64_t* stats);
HG_API int hg_set_state(hg_ctx* ctx, const double* Q);          /* host [3N] -> device                  */
HG_API int hg_get_state(hg_ctx* ctx, double* Q);                /* device -> host [3N]                  */
HG_API int hg_set_params(hg_ctx* ctx, const double* params, int64_t n_params, int32_t active_param);
/* dQdt of the resident state into a resident buffer; hg_get_rhs copies it out.  Asynchronous on
 * the ctx stream; hg_sync waits.                                                                 */
HG_API int hg_rhs_resident(hg_ctx* ctx);
HG_API int hg_get_rhs(hg_ctx* ctx, double* dQdt);
HG_API int hg_sync(hg_ctx* ctx);

/* nsteps of custom_ODE_update_cells (custom_ODE_solvers.jl:5-33) on the resident state:
 * Q+ = Q + dt*rhs(Q); where xi+ < h_small: xi+ = h_small, q+ = 0 (the reference's mask is on xi).
 * Fused into the RHS kernel (one launch per step, captured in a CUDA graph).                      */
HG_API int hg_step_euler(hg_ctx* ctx, double dt, int64_t nsteps);

/* nsteps of the classical fixed-step RK4 on the resident state (what `solve(prob, RK4(), adaptive=false, dt=dt)` does
 * in swe_2D_forward_simulation.jl:44), four fused RHS launches + three axpy kernels per step, no host round trip.   */
HG_API int hg_step_rk4(hg_ctx* ctx, double dt, int64_t nsteps);

/* The other explicit fixed-step solvers the control files accept (forward_simulation_ode_solver / inversion ode_solver:
 * swe_2D_forward_simulation.jl:40-47, swe_2D_inversion.jl:272-275), on the resident state:
 *   hg_step_ode_euler  OrdinaryDiffEq's `Euler()`: u+ = u + dt f(u) -- no dry mask, unlike hg_step_euler;
 *   hg_step_ab3        OrdinaryDiffEq's `AB3()`: u+ = u + dt/12 (23 f_n - 16 f_{n-1} + 5 f_{n-2}), started with two steps of
 *                      Ralston's method u+ = u + dt/4 (k1 + 3 f(u + 2/3 dt k1)).  The two older slopes stay on the device
 *                      between calls; restart != 0, hg_set_state or any other solver call starts the sequence again.
 * (`Rosenbrock23()` / `TRBDF2()` are implicit and not built.) */
HG_API int hg_step_ode_euler(hg_ctx* ctx, double dt, int64_t nsteps);
HG_API int hg_step_ab3(hg_ctx* ctx, double dt, int64_t nsteps, int32_t restart);

/* Tsit5 (adaptive with OrdinaryDiffEq's PI controller, or fixed-step when adaptive = 0) on the resident state from t0 to
 * t1 -- `solve(prob, Tsit5(), adaptive=..., dt=dt, saveat=t_save; abstol=1e-6, reltol=1e-3)` of
 * swe_2D_forward_simulation.jl:38-41 / swe_2D_sensitivity.jl:38-43 without a host round trip per stage.  dt is the initial
 * (adaptive) or the fixed step.  t_save[n_save] in [t0, t1] are stops at which the state is copied to Q_save[n_save][3N]
 * (reference order); OrdinaryDiffEq interpolates its saves instead (hg_solve_tsit5_dense), both agree within the
 * integration tolerance.
 * stats[3] (may be NULL) = accepted steps, rejected steps, RHS evaluations.
 * Multi-rank contexts (after hg_comm_connect / hg_comm_init_shm): collective, every stage exchanges halos; an adaptive
 * solve additionally needs hg_comm_set_allreduce (the error norm is summed over ranks), HG_ERR_ARG without it. */
HG_API int hg_solve_tsit5(hg_ctx* ctx, double t0, double t1, double dt, int32_t adaptive, double abstol, double reltol,
                          const double* t_save, int64_t n_save, double* Q_save, int64_t* stats);

/* The powers EEst^beta1 and qold^beta2 of the PI controller: mode 0 (default) the correctly rounded pow; mode 1
 * DiffEqBase's `fastpow` (Float32 rational-approximation log2 + exp2, ~1e-5 relative error), which OrdinaryDiffEq used in the
 * generation the reference ran -- with it an adaptive solve follows OrdinaryDiffEq's own step sequence (the reference's
 * committed Savannah forward runs are reproduced to 1e-9; with mode 0 to 1e-7).  hg_fastpow exposes the function. */
HG_API int hg_set_controller_pow(hg_ctx* ctx, int32_t mode);
HG_API double hg_fastpow(double x, double y);

/* The same solve with OrdinaryDiffEq's saveat semantics (savevalues!, Tsit5's dense output `Tsit5Interp`): the steps do
 * not stop at the save times (only t1 is a stop); after every accepted step [t, t + h] the fourth-order interpolant
 * u + h sum_i b_i(theta) k_i, theta = (t_save - t) / h, is evaluated on the device for every save time the step has
 * passed, and a save time equal to the step end copies the new state.  This reproduces the reference's step sequence
 * (the step sizes do not depend on t_save); arguments as for hg_solve_tsit5. */
HG_API int hg_solve_tsit5_dense(hg_ctx* ctx, double t0, double t1, double dt, int32_t adaptive, double abstol, double reltol,
                                const double* t_save, int64_t n_save, double* Q_save, int64_t* stats);

/* Discrete adjoint of nsteps of hg_step_euler (what SciMLSensitivity + Zygote produce for the "customized" Euler
 * solver inside compute_loss_inversion, swe_2D_inversion.jl:339): given lambda_T = d loss / d Q(T) it returns
 * Q0bar = d loss / d Q0 [3N] and pbar = d loss / d params [n_params] (NULL when active_param = NONE).  The forward
 * sweep keeps sqrt(nsteps)-spaced checkpoints on the device, the reverse sweep recomputes each segment and calls the
 * VJP kernel once per step; the dry mask of custom_ODE_update_cells is a constant selector, like every other clamp.
 * Multi-rank contexts with a connected transport (also hg_rk_adjoint / hg_rk_adjoint_steps): collective; Q0bar covers the
 * owned cells and pbar is this rank's partial sum -- add the ranks' pbar in a fixed order. */
HG_API int hg_euler_adjoint(hg_ctx* ctx, const double* Q0, const double* params, int64_t n_params, int32_t active_param,
                     double dt, int64_t nsteps, const double* lambda_T, double* Q_T, double* Q0bar, double* pbar);

/* Discrete adjoint of nsteps of a fixed-step explicit Runge-Kutta method -- method 0: classical RK4 (hg_step_rk4),
 * 1: Tsit5 (hg_solve_tsit5 with adaptive = 0) -- i.e. the exact derivative of those steppers ("discretise, then
 * differentiate"), the counterpart of hg_euler_adjoint for the SciML solvers of swe_2D_inversion.jl:339.  Same arguments
 * and outputs as hg_euler_adjoint; stage states are recomputed per step, ~sqrt(nsteps) checkpoints are kept. */
HG_API int hg_rk_adjoint(hg_ctx* ctx, int32_t method, const double* Q0, const double* params, int64_t np, int32_t active,
                         double dt, int64_t nsteps, const double* lambda_T, double* Q_T, double* Q0bar, double* pbar);

/* The same discrete adjoint over a GIVEN sequence of step sizes h[nsteps] -- in particular the accepted steps of an adaptive
 * solve, which hg_last_steps returns after hg_solve_tsit5(_dense): the gradient of the reference's default integrator
 * (`solve(prob, Tsit5(), adaptive=true, ...; sensealg=...)`, swe_2D_inversion.jl:339) with the step sizes held constant, as
 * every discretise-then-differentiate adjoint does.  hg_last_steps: n receives the number of accepted steps, h (may be NULL)
 * the first min(capacity, n) of them. */
HG_API int hg_rk_adjoint_steps(hg_ctx* ctx, int32_t method, const double* Q0, const double* params, int64_t np, int32_t active,
                               const double* h, int64_t nsteps, const double* lambda_T, double* Q_T, double* Q0bar, double* pbar);
HG_API int hg_last_steps(const hg_ctx* ctx, double* h, int64_t capacity, int64_t* n);

/* custom_ODE_solve (custom_ODE_solvers.jl:36-95): steps over t_start:dt:t_end, saving every
 * state; sol is [3N x n_saves] column-major, n_saves_capacity columns available; *n_saves out.   */
HG_API int hg_custom_ode_solve(hg_ctx* ctx, const double* Q0, const double* params, int64_t n_params,
                        int32_t active_param, double t_start, double t_end, double dt, double* sol,
                        int64_t n_saves_capacity, int64_t* n_saves);

/* ---- multi-GPU halo exchange (one context per rank; the transport -- NCCL send/recv -- is the caller's).
 * Buffers are device memory owned by the context: for neighbour k (k-th halo boundary, n_k entries) a block of
 * 6*n_k doubles [xi | q_x | q_y | lambda_0 | lambda_1 | lambda_2] at offset 6*(entries before k).  A RHS-only
 * exchange moves the first 3*n_k doubles of each block.                                              */
HG_API int hg_halo_info(const hg_ctx* ctx, int64_t* n_neighbors, int64_t* n_entries);
HG_API int hg_halo_counts(const hg_ctx* ctx, int64_t* counts /* [n_neighbors] */);
HG_API int hg_halo_buffers(hg_ctx* ctx, double** d_send, double** d_recv, int64_t* n_doubles);
HG_API int hg_halo_pack(hg_ctx* ctx, int32_t with_lambda);   /* resident state (and lambda) -> send buffer      */
/* run every kernel of this context on the caller's CUDA stream (e.g. torch's current stream) so that packing,
 * the NCCL transfers and the RHS are ordered without host synchronisation; NULL restores the own stream   */
HG_API int hg_set_stream(hg_ctx* ctx, void* cuda_stream);
/* resident cotangent for hg_vjp_resident / hg_time_vjp */
HG_API int hg_set_lambda(hg_ctx* ctx, const double* lambda);
HG_API int hg_vjp_resident(hg_ctx* ctx);

/* ---- library-owned halo exchange over NVLink peer memory (hg_comm.cu): replaces the caller-side NCCL send/recv.
 * Each rank exports a handle of its receive block, collects its neighbours' handles by whatever messaging the host has
 * (MPI, Distributed.jl, torch.distributed: plumbing only) and connects; or lets hg_comm_init_shm do the rendezvous through
 * POSIX shared memory (one process per GPU on one box, no messaging layer needed).  Ranks living in ONE process (one host
 * process driving several GPUs) connect the same way -- peer access instead of CUDA IPC.  Once connected, every resident RHS /
 * VJP / stepper call pushes the cut cells' states (and cotangents) straight into the neighbours' buffers and the tile kernel
 * itself waits for the halo only in the tiles that touch it: calls are COLLECTIVE (same sequence on every rank).
 *   peer_handles        [n_neighbors * HG_COMM_HANDLE_BYTES], neighbour k = k-th halo boundary of this context
 *   peer_entry_offset   [n_neighbors] halo entries that precede THIS rank's block in neighbour k's own boundary list
 *   peer_flag_index     [n_neighbors] index of THIS rank in neighbour k's neighbour list                                  */
#define HG_COMM_HANDLE_BYTES 128
HG_API int hg_comm_export(hg_ctx* ctx, void* handle /* [HG_COMM_HANDLE_BYTES] */);
HG_API int hg_comm_connect(hg_ctx* ctx, int64_t n_neighbors, const void* peer_handles, const int64_t* peer_entry_offset,
                           const int64_t* peer_flag_index);
HG_API int hg_comm_init_shm(hg_ctx* ctx, const char* job_name, int32_t rank, int32_t world, const int32_t* neighbor_ranks /* [n_neighbors] */);
/* auto (default 1): every resident RHS / VJP evaluation starts with an exchange of the state it evaluates.  auto = 0: the
 * caller issues hg_comm_exchange itself (several contexts of ONE process on ONE device must push all before any consumes). */
HG_API int hg_comm_set_auto(hg_ctx* ctx, int32_t on);
HG_API int hg_comm_exchange(hg_ctx* ctx, int32_t with_lambda);
HG_API int hg_comm_disconnect(hg_ctx* ctx);   /* call after a barrier: peers must not push any more */
/* Host-side sum over ranks for the few scalars that are global: the error norm of an adaptive hg_solve_tsit5 on a
 * multi-rank context (two doubles per step).  fn adds inout[0..n) over all ranks in place and returns 0 (MPI_Allreduce,
 * torch.distributed.all_reduce, Distributed.jl ...).  Everything else -- fixed-step solvers, the time adjoints -- needs no
 * global sum: Q0bar covers the owned cells, pbar is the rank's partial sum.                                              */
typedef int (*hg_allreduce_fn)(double* inout, int64_t n, void* user);
HG_API int hg_comm_set_allreduce(hg_ctx* ctx, hg_allreduce_fn fn, void* user);

/* Multi-GPU overlap of the halo exchange with the tiles that need no remote cell.  phase 1: everything that does not
 * touch a halo face (launch it right after hg_halo_pack, while the exchange is in flight on another stream); phase 2: the
 * band of tiles with halo faces (launch it once the received block is complete); phase 0 = hg_rhs_resident /
 * hg_vjp_resident.  Phases 1 + 2 together give bitwise the result of phase 0. */
HG_API int hg_rhs_resident_phase(hg_ctx* ctx, int32_t phase);
HG_API int hg_vjp_resident_phase(hg_ctx* ctx, int32_t phase);
HG_API int hg_get_vjp(hg_ctx* ctx, double* Qbar, double* pbar, double* ncell_bar);

/* ---- parameter ensembles (sensitivity / uncertainty runs: BASELINE config "1024 parameter sets"): M independent
 * members on ONE mesh advance in ONE launch per step (member index fastest in the CTA order, so the mesh tables of
 * a tile are fetched from HBM once and shared through L2).  Members differ in state, and optionally in the Manning
 * zone values (active = HG_PARAM_MANNING) or the inlet discharges (HG_PARAM_Q).  No communication: shard members
 * across GPUs by creating one context per GPU.                                                      */
HG_API int hg_ensemble_alloc(hg_ctx* ctx, int64_t n_members, int32_t per_member_manning);
HG_API int hg_ensemble_set_member(hg_ctx* ctx, int64_t member, const double* Q, const double* params, int64_t n_params,
                           int32_t active_param);
HG_API int hg_ensemble_step_euler(hg_ctx* ctx, double dt, int64_t nsteps);
HG_API int hg_ensemble_rhs(hg_ctx* ctx);                                /* dQdt of every member (resident)       */
HG_API int hg_ensemble_get_member(hg_ctx* ctx, int64_t member, int32_t what /* 0 state, 1 dQdt */, double* out);
HG_API int hg_time_ensemble(hg_ctx* ctx, int32_t n_steps, double dt, float* ms_total);

/* ---- SRH-2D case reader + mesh / boundary / bed builder, host only, linear time (replaces, for callers without the
 * Julia package, utilities/SRH_2D/* + meshes/mesh_2D.jl:75-652 + bc_2D.jl:50-570 + process_bed_2D.jl:9-66; the
 * reference builder is O(N*B) and cannot feed million-cell meshes).  The case owns its arrays; fetch them by name:
 * cell_nfaces cell_faces cell_neighbors cell_nodes (int64), cell_normals face_lengths cell_areas cell_centroids
 * bc_normals bc_lengths zb_cells zb_ghost S0_cells ManningN_cells ManningN_zone inletQ_TotalQ exitH_WSE node_coords
 * (float64), bc_ptr bc_ghost_ids bc_internal_cells matID_cells (int64), face_is_boundary (uint8).  Ids are 1-based
 * (index_base = 1), N x 8 tables column-major.  dims = {N, F, B, ld, index_base, n_inletq, n_exith, n_wall, n_symm,
 * n_mat, n_nodes, ...}.                                                                              */
typedef struct hg_case hg_case;
HG_API int hg_case_load_srh2d(hg_case** out, const char* srhhydro_path, char* err, int64_t errlen);
HG_API void hg_case_free(hg_case* c);
HG_API int hg_case_dims(const hg_case* c, int64_t* dims /* [16] */);
HG_API int hg_case_array(const hg_case* c, const char* name, const void** ptr, int64_t* count, int32_t* dtype /* 0 f64, 1 i64, 2 u8 */);

/* ---- domain decomposition for multi-GPU runs, host only (csrc/hg_partition.cpp; no counterpart in the reference, which is a
 * single serial process).  hg_partition_rcb: recursive coordinate bisection of the centroids into P parts; cell groups
 * (group_ptr [n_groups+1] into group_cells, 0-based ids; e.g. the cells of each inlet-q boundary) are moved as a whole to the
 * rank owning most of them.  hg_partition_extract: the rank-local mesh of `rank` as an hg_case (free with hg_case_free) whose
 * arrays carry the names of the descriptor fields they feed -- cell_nfaces cell_faces cell_neighbors bc_ptr bc_ghost_ids
 * bc_internal_cells matID_cells (int64), cell_normals face_lengths cell_areas cell_centroids bc_normals bc_lengths halo_area
 * hstill hstill_ghost zb_cells zb_ghost S0_cells ManningN_cells inletQ_TotalQ exitH_WSE (float64), face_is_boundary halo_flip
 * (uint8), 0-based ids -- plus own (global ids of the owned cells), neighbors (ranks, = halo boundaries in order), counts
 * (entries per neighbour), halo_cells (local cell of every halo entry), halo_remote (global id of its remote cell).
 * dims = {N, F, B, ld, index_base, n_inletq, n_exith, n_wall, n_symm, n_mat, n_halo, n_halo_entries}.  `gid` (optional) are the
 * global ids of the input cells when the input is itself a piece of a larger mesh.                                     */
HG_API int hg_partition_rcb(int64_t n_cells, const double* cx, const double* cy, int32_t n_parts, int64_t n_groups,
                            const int64_t* group_ptr, const int64_t* group_cells, int32_t* part /* [n_cells] out */);
HG_API int hg_partition_extract(hg_case** out, const hg_mesh_desc* mesh, const hg_bc_desc* bc, const hg_fields_desc* fields,
                                const int32_t* part, int32_t rank, const int64_t* gid, char* err, int64_t errlen);

/* ---- results writers and the derived fields of the forward driver, host only (csrc/hg_results.cpp): the output side
 * of the path.  Replaces, for callers without the Julia package, postprocess_forward_simulation_results_swe_2D
 * (applications/forward_simulation/process_forward_simulation_results_2D.jl:4-86), swe_2D_save_results_SciML and
 * export_to_vtk_2D (utilities/swe_2D_tools.jl:10-98, 145-214), and the JSON3.pretty calls of the sensitivity driver
 * (applications/sensitivity/swe_2D_sensitivity.jl:70,90).  Files are byte-compatible with the reference's: numbers are
 * formatted the way Julia prints a Float64 (shortest round-trip digits, Base.Ryu.writeshortest's layout), and in JSON the
 * way JSON3.pretty leaves them (whole values as integers).                                                        */
enum { HG_FMT_JULIA = 0, HG_FMT_JSON3 = 1 };
/* formats x into out (cap >= 32), NUL-terminated; returns the length, -1 on a bad argument */
HG_API int hg_format_f64(double x, int32_t style, char* out, int64_t cap);

/* streaming JSON3.pretty writer of one top-level object: key, then a number / string / (nested) array of numbers */
typedef struct hg_json hg_json;
/* style HG_FMT_JSON3: numbers as JSON3.pretty (JSON3 1.14) leaves them -- whole values as integers, except in arrays
 * whose first element is not whole; HG_FMT_JULIA: every number as Julia prints a Float64 */
HG_API int hg_json_open(hg_json** out, const char* path, int32_t style, char* err, int64_t errlen);
HG_API const char* hg_json_error(const hg_json* w);
HG_API int hg_json_key(hg_json* w, const char* key);
HG_API int hg_json_begin_array(hg_json* w);
HG_API int hg_json_end_array(hg_json* w);
HG_API int hg_json_numbers(hg_json* w, const double* x, int64_t n);   /* n elements of the open array; NaN / Inf refused like JSON3 */
HG_API int hg_json_number(hg_json* w, double x);
HG_API int hg_json_string(hg_json* w, const char* value);
HG_API int hg_json_close(hg_json* w, int32_t trailing_newline);       /* always frees w                                */

typedef struct {
  const char* name;
  const double* data;   /* scalars: [n_cells]; vectors: n_cells x 2 column-major (hcat(x, y))                          */
} hg_named_array;
/* export_to_vtk_2D: node_xyz [n_nodes][3] (x y z per node), cell_nodes N x ld column-major with ids from index_base,
 * FIELD block only when field_name is not empty.                                                                  */
HG_API int hg_write_vtk_2d(const char* path, int64_t n_nodes, const double* node_xyz, int64_t n_cells, int64_t ld, int32_t index_base,
                    const int64_t* cell_nodes, const int64_t* cell_nnodes, const char* field_name, const char* field_type,
                    double field_value, const hg_named_array* scalars, int64_t n_scalars, const hg_named_array* vectors,
                    int64_t n_vectors, char* err, int64_t errlen);
/* xi, wse, h, u, v, friction_x/y of a saved state Q[3N] (any output may be NULL) */
HG_API int hg_forward_truth_fields(int64_t N, const double* Q, const double* hstill, const double* wstill, const double* ManningN_cells,
                            double g, double k_n, double h_small, double* xi, double* wse, double* h, double* u, double* v,
                            double* friction_x, double* friction_y);
/* update_ManningN_forward_simulation on the host: n and the closures' diagnostics h/ks, f, Re (NULL to skip) */
HG_API int hg_manning_function_cells(int32_t type, const double* params, int64_t N, const double* h, const double* umag, const double* ks,
                              double* n, double* h_ks, double* f, double* Re);
/* process_dry_wet_flags (fvm/discretization/process_dry_wet.jl:2-35), read by the VTK file only */
HG_API int hg_dry_wet_flags(int64_t N, int64_t ld, int32_t index_base, const int64_t* cell_nfaces, const int64_t* cell_faces,
                     const int64_t* cell_neighbors, const uint8_t* face_is_boundary, int64_t n_faces, const double* h,
                     const double* zb_cells, double h_small, uint8_t* b_dry_wet, uint8_t* adjacent_to_dry_land,
                     uint8_t* adjacent_to_high_dry_land);
/* swe_2D_calc_total_water_volume: pairwise sum(h .* cell_areas) */
HG_API double hg_total_water_volume(int64_t N, const double* h, const double* cell_areas);

/* ---- timing hooks used by bench.py (device time of the last N launches, CUDA events on the
 * ctx stream) and introspection for the roofline arithmetic.                                     */
HG_API int hg_time_rhs(hg_ctx* ctx, int32_t n_launches, int32_t fused_euler, double dt, float* ms_total);
HG_API int hg_time_vjp(hg_ctx* ctx, int32_t n_launches, float* ms_total);
HG_API int hg_time_jvp(hg_ctx* ctx, int32_t n_directions, int32_t n_launches, float* ms_total);   /* fused forward-mode kernel */
HG_API int64_t hg_kernel_launches(const hg_ctx* ctx);            /* kernels launched so far             */
/* A counter that changes whenever the resident state changes (uploads by hg_rhs / hg_rhs_vjp / hg_set_state / the forward-
 * mode calls, every stepper, solve and time adjoint).  Equal values before and after = the state of the earlier call is
 * still on the device, so hg_rhs_vjp may be called with Q = NULL.  -1 for a NULL context.                             */
HG_API int64_t hg_state_generation(const hg_ctx* ctx);
HG_API int hg_mesh_stats(const hg_ctx* ctx, int64_t* n_cells, int64_t* n_faces, int64_t* sum_cell_faces,
                  int64_t* n_tiles, int64_t* device_bytes);
/* Host-only: run the hg_create preprocessing (validation, renumbering, tiling) WITHOUT touching a GPU
 * and report the plan: stats[0..7] = n_tiles, max local cells, max tile faces, shared-memory bytes per
 * CTA, total halo cells, total tile faces, interior tile faces, sum of cell faces.  perm_out (optional,
 * [N]) receives the internal->reference cell permutation.                                          */
HG_API int hg_plan_stats(const hg_mesh_desc* mesh, const hg_bc_desc* bc, const hg_fields_desc* fields,
                  const hg_options* opt, int64_t* stats, int64_t* perm_out);
/* Host-only: the stage tables of the host-buffer pipeline (hg_rhs / hg_rhs_vjp with host pointers on >= 1M cells) for this mesh.
 * header[5] = n_chunks, nominal rows per chunk, n_tiles, tile_cells, alignment margin in rows; tile_stage[n_tiles] (may be NULL)
 * = the chunk after which a tile's cells and halo have all landed; chunk_done[n_chunks] (may be NULL) = the stage after which
 * result chunk c may leave.  opt->reserved[1] > 0 overrides the chunk count (tuning).                                    */
HG_API int hg_plan_pipeline(const hg_mesh_desc* mesh, const hg_bc_desc* bc, const hg_fields_desc* fields, const hg_options* opt,
                     int64_t* header, int32_t* tile_stage, int32_t* chunk_done);
/* Host-only: rows [r0, r1) that chunk c (of K, nominal size rows_per_chunk) of a [N] component moves when the component starts at
 * host address host_addr: boundaries are shifted so that every copy starts at a multiple of 256 bytes (test hook). */
HG_API int hg_debug_chunk_rows(uint64_t host_addr, int32_t c, int32_t K, int64_t rows_per_chunk, int64_t N, int64_t* r0, int64_t* r1);

/* Host-only: the tile tables hg_create would upload for this mesh, by name (test / inspection hook; no GPU touched).
 * hg_plan_array: "dims" (i64: N, B, n_tiles, tile_cells, slots per cell, component stride, ints per tile descriptor, n_chunks,
 * tiles without halo faces, first band position of comm_order); "tile_order" "band_order" "comm_order" (i32: the launch orders of
 * the host-buffer pipeline, of the two-phase overlap and of the library-owned transport; empty on a mesh without halo faces);
 * "bcell_ref" "bcell_ptr" "bcell_ent" (i32: distinct boundary-adjacent cells -> their boundary entries) "cf_rev" (i32: per cell-face
 * of the reference-order CSR, where the same face sits in the neighbour's list; -1 on boundary faces);
 * "perm" "iperm" "tile_desc" "halo" "bface_e" "bc_type" "bc_group" "bc_ghost" "bc_cell_ref" "inlet_ptr" (i32); "face_lr" (u32);
 * "cf_idx" (u16); "face_nx" "face_ny" "face_len" "bc_nx" "bc_ny" "bc_l53" "bc_l23" "bc_hstill" "bc_zb" (f64).  The pointers
 * stay valid until hg_plan_close.  The layout is described in DESIGN.md section 3.                                      */
typedef struct hg_plan hg_plan;
HG_API int hg_plan_open(const hg_mesh_desc* mesh, const hg_bc_desc* bc, const hg_fields_desc* fields, const hg_options* opt, hg_plan** out);
HG_API void hg_plan_close(hg_plan* p);
HG_API int hg_plan_array(const hg_plan* p, const char* name, const void** ptr, int64_t* count,
                  int32_t* dtype /* 0 f64, 1 i64, 2 u8, 3 i32, 4 u32, 5 u16 */);
/* Accuracy probe of the kernels' branch-free fp64 helpers (hg_device.cuh): out[i] = f(x[i]) evaluated on the device,
 * kind 0 = 1/x, 1 = 1/sqrt(x), 2 = sqrt(x), 3 = sqrt(x^2 + eps) (smooth abs), 4 = x^(-7/3); x > 0, host pointers.          */
HG_API int hg_debug_math(hg_ctx* ctx, int32_t kind, int64_t n, const double* x, double* out);
/* write 256 MiB of device scratch (L2 flush between timed iterations)                             */
HG_API int hg_flush_l2(hg_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* HYDROGRAD_B200_H */
