/*
 * txasm.h -- C ABI of the B200-native finite-element assembly path.
 *
 * Drop-in boundary for the ONE hot path of hillyuan/Tianxin (Panzer fork): the fp64 residual +
 * Jacobian volume assembly behind panzer::AssemblyEngine<EvalT>::evaluate, i.e.
 *
 *   GatherSolution_Tpetra -> DOF/DOFGradient -> Integrator_* -> ScatterResidual_Tpetra
 *   bracketed by TpetraLinearObjFactory::globalToGhostContainer / ghostToGlobalContainer.
 *
 * A reference-side maintainer binds these symbols from the (unchanged) host C++:
 * disc-fe's AssemblyEngine forwards its stages here; dof-mgr's GlobalIndexer hands over the
 * LID table; lof's TpetraLinearObjContainer hands over the raw Tpetra local views.  The stub
 * is shown in INTEGRATION.md.  Every entry point cites the reference interface it replaces
 * (paths relative to the reference tree).
 *
 * Conventions
 *  - plain C, plain pointers and sizes; ordinals as in core/cmake/PanzerCore_config.hpp.in:19-32:
 *    LocalOrdinal = int, GlobalOrdinal = long long; CSR row offsets are 64-bit.
 *  - memory spaces: every array argument may be a HOST or a DEVICE pointer; the library asks
 *    the CUDA runtime (cudaPointerGetAttributes).  Device (or managed) arrays are used in
 *    place (zero copy, e.g. Kokkos device views); host arrays are staged through
 *    library-owned device buffers (setup arrays once, solution/result arrays on every call).
 *  - ownership: the caller owns every array it passes; arrays passed to setup calls by DEVICE
 *    pointer must outlive the handle.  The library owns its handle and its device scratch.
 *  - errors: every function returns 0 (TXASM_OK) or a negative TXASM_E* code and never
 *    throws; txasm_last_error() returns the message.  CUDA / NCCL errors are sticky.
 *  - threading: a handle is bound to one device and one stream; calls on one handle must be
 *    serialised by the caller.  All work is stream-ordered; calls return after enqueue
 *    unless host arrays have to be filled (then they synchronise) or txasm_sync is called.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *    TXASM_ECUDA.
 */
#ifndef TXASM_H
#define TXASM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TXASM_VERSION_MAJOR 0
#define TXASM_VERSION_MINOR 2

typedef struct txasm_handle_s *txasm_handle;

enum {
  TXASM_OK = 0,
  TXASM_EINVAL = -1,       /* bad argument */
  TXASM_ECUDA = -2,        /* CUDA runtime error (incl. "no device") */
  TXASM_ENOMEM = -3,
  TXASM_ESTATE = -4,       /* call order violated (e.g. evaluate before setup) */
  TXASM_ENCCL = -5,
  TXASM_EUNSUPPORTED = -6  /* element / basis / term not implemented */
};

/* cell topologies and bases named as Panzer_IntrepidBasisFactory.hpp:138-235 selects them */
enum { TXASM_TOPO_HEX8 = 8, TXASM_TOPO_HEX27 = 27, TXASM_TOPO_TET4 = 4, TXASM_TOPO_TET10 = 10 };
enum { TXASM_BASIS_HGRAD_C1 = 1,   /* Basis_HGRAD_HEX_C1 / _TET_C1   (Panzer_IntrepidBasisFactory.hpp:153-184) */
       TXASM_BASIS_HGRAD_C2 = 2,   /* Basis_HGRAD_HEX_C2 / _TET_C2 */
       TXASM_BASIS_HCURL_I1 = 3 }; /* Basis_HCURL_HEX_I1           (:157-158,169) */

/* evaluation types: panzer::Traits::Residual / ::Jacobian (disc-fe/src/Panzer_Traits.hpp) */
enum { TXASM_RESIDUAL = 0, TXASM_JACOBIAN = 1 };

/* AssemblyEngine<EvalT>::EvaluationFlags (disc-fe/src/Panzer_AssemblyEngine.hpp:72-86): same bit values */
enum {
  TXASM_FLAG_INITIALIZE = 1,      /* globalToGhost of x, dxdt (halo import) */
  TXASM_FLAG_VOLUMETRIC_FILL = 2, /* evaluateVolume */
  TXASM_FLAG_BOUNDARY_FILL = 4,   /* evaluateDirichletCondition */
  TXASM_FLAG_SCATTER = 8,         /* ghostToGlobal of f, A (halo export, ADD) */
  TXASM_FLAG_ALL = 15
};

/* How element contributions reach f and A.
 *  ROWTILE: owner-computes.  Rows are grouped into spatial tiles; one CTA computes every
 *           element touching its rows and writes each row of A (and f) exactly once with
 *           plain coalesced stores: no atomics, no zero fill, bitwise reproducible.
 *  ATOMIC:  element-parallel, per entry binary search of the CSR row + red.global.add.f64 --
 *           the literal restatement of ScatterResidual_Tpetra's
 *           jac.sumIntoValues(lid, lids, N, vals, true, true)
 *           (disc-fe/src/evaluators/Panzer_ScatterResidual_Tpetra_impl.hpp:390-412).
 *  AUTO:    ROWTILE where the mesh qualifies, the general row-gather otherwise. */
enum { TXASM_SCATTER_AUTO = 0, TXASM_SCATTER_ROWTILE = 1, TXASM_SCATTER_ATOMIC = 2, TXASM_SCATTER_ROWGATHER = 3,
       TXASM_SCATTER_GENERIC = 4 /* reported by txasm_info for handles built from txasm_gblock_add blocks */ };

typedef struct {
  int    device;        /* CUDA device ordinal */
  void  *stream;        /* cudaStream_t to enqueue on; NULL = the library creates a (blocking) stream, which is
                           ordered against work on the legacy default stream.  Arrays produced on OTHER
                           streams must be complete before they are passed in */
  int    scatter_mode;  /* TXASM_SCATTER_* */
  double affine_tol;    /* an element is treated as a parallelepiped (constant Jacobian, exact
                           integration) when its vertices deviate from one by less than
                           affine_tol * (longest edge); <0 = never; default 1e-13 when 0 */
  int    reserved[8];
} txasm_config;

/* Integrand terms.  They mirror the integrator evaluators an equation set registers
 * (adapters-stk/example/PoissonExample/Example_PoissonEquationSet_impl.hpp:150-195):
 *   GRADGRAD : Integrator_GradBasisDotVector   r_b += sum_q wgrad_b(q) . M grad(u)(q)
 *              (disc-fe/src/evaluators/Panzer_Integrator_GradBasisDotVector_impl.hpp:220-294)
 *   MASS     : Integrator_BasisTimesScalar     r_b += sum_q wphi_b(q) M u(q)
 *              (disc-fe/src/evaluators/Panzer_Integrator_BasisTimesScalar_impl.hpp:210-270)
 *   SOURCE   : Integrator_BasisTimesScalar on a closure-model field s(x_q)  (M = -1 in the example)
 * `vec` selects which solution vector the integrand is gathered from, which also selects the
 * forward-mode seed exactly as GatherSolution_Tpetra<Jacobian> does
 * (disc-fe/src/evaluators/Panzer_GatherSolution_Tpetra_impl.hpp:554-572): X -> beta,
 * XDOT -> alpha.  XDOTDOT -> gamma is an extension: the reference plumbs d2xdt2 through its
 * containers but has no gather for it. */
enum { TXASM_TERM_GRADGRAD = 1, TXASM_TERM_MASS = 2, TXASM_TERM_SOURCE = 3,
       TXASM_TERM_TRANSIENT_MASS = 4  /* Integrator_TransientBasisTimesScalar: MASS that contributes only when
                                         txasm_inargs.evaluate_transient_terms is set (workset.evaluate_transient_terms,
                                         disc-fe/src/evaluators/Panzer_Integrator_TransientBasisTimesScalar_impl.hpp) */ };
enum { TXASM_VEC_X = 0, TXASM_VEC_XDOT = 1, TXASM_VEC_XDOTDOT = 2 };
/* built-in closure models for SOURCE */
enum {
  TXASM_SOURCE_CONSTANT = 2,    /* s = 1 */
  TXASM_SOURCE_SIN3 = 1,        /* s = 12 pi^2 sin(2 pi x) sin(2 pi y) sin(2 pi z): 3-D analogue of
                                   Example_SimpleSource_impl.hpp:88-96 */
  TXASM_SOURCE_IP_ARRAY = 100   /* s given at the integration points: double[n_cells][n_qp] */
};
typedef struct {
  int    kind;        /* TXASM_TERM_* */
  int    vec;         /* TXASM_VEC_*  (GRADGRAD, MASS) */
  double multiplier;  /* "Multiplier" of the integrator */
  int    source_id;   /* TXASM_SOURCE_* (SOURCE) */
  const double *ip_values; /* TXASM_SOURCE_IP_ARRAY */
  int    gather_seed_index1; /* 0: seed by `vec` (beta / alpha / gamma).  k > 0: the gather of this term's DOF has
                                "Gather Seed Index" k-1 and is seeded with txasm_inargs.gather_seeds[k-1]
                                (Panzer_GatherSolution_Tpetra_impl.hpp:554-572) */
  int    reserved;
  const double *field_multiplier_ip; /* GRADGRAD / MASS: product of the integrator's "Field Multipliers" at the integration
                                points, double[n_cells][n_qp] (Panzer_Integrator_GradBasisDotVector_impl.hpp:257-294:
                                r_b += sum_q wgrad_b(q) . M grad u(q) prod_k fm_k(q)), or NULL.  All GRADGRAD terms of a
                                block must name the same array (likewise all MASS terms); cells of such a block take the
                                general 2x2x2 path (the coefficient is not constant over a cell).  Set before txasm_setup. */
} txasm_term;

/* panzer::AssemblyEngineInArgs scalars (disc-fe/src/Panzer_AssemblyEngine_InArgs.hpp:92-107)
 * copied into every workset by evaluateVolume (Panzer_AssemblyEngine_impl.hpp:163-169). */
typedef struct {
  double alpha, beta, gamma;
  double time, step_size, stage_number;
  int    evaluate_transient_terms;
  int    zero_outputs;   /* 1: f and A are zeroed first (what initializeGhostedContainer /
                            setAllToScalar(0) do around the reference's evaluate); the ROWTILE
                            and ROWGATHER modes overwrite every entry and ignore this */
  int    n_gather_seeds; /* AssemblyEngineInArgs::gather_seeds (Panzer_AssemblyEngine_InArgs.hpp:104) */
  const double *gather_seeds;   /* host array */
} txasm_inargs;

/* the stage timers of AssemblyEngine::evaluate (Panzer_AssemblyEngine_impl.hpp:74,89,102,107,112,118), in milliseconds
 * of device time for the last txasm_evaluate.  evaluate_interfacebcs is always 0: interface conditions are not
 * implemented (the Poisson / elasticity / Maxwell paths have none).  A Dirichlet stage fused into the fill reads ~0. */
typedef struct {
  double evaluate_gather, evaluate_volume, evaluate_neumannbcs, evaluate_interfacebcs,
         evaluate_dirichletbcs, evaluate_scatter;
} txasm_timers;

typedef struct {
  int64_t n_cells, n_rows, nnz;
  int64_t n_affine_cells;      /* cells on the constant-Jacobian path */
  int64_t n_regular_rows;      /* rows on the register-accumulated 27-point path */
  int     scatter_mode;        /* mode actually selected */
  int     n_tiles, tile_rows_max, tile_cells_max;
  int     smem_bytes, threads_per_cta, ctas_per_sm;
  int     kernel_launches_last_evaluate;
  int     n_sm;
  /* appended in 0.2 (the struct only grows at the end) */
  int     n_uniform_tiles;     /* tiles [0, n_uniform_tiles): congruent axis-aligned cells, every row interior with canonical
                                  column order -- eligible for the lean uniform-tile kernel */
  int     n_brick_tiles;       /* ... of which the node set is a full tensor brick (k_fill_brick) */
  int     uniform_kernel_used; /* last evaluate: 0 general row-tile kernel only, 1 k_fill_uniform, 2 k_fill_brick */
  int     dirichlet_fused;     /* last evaluate: Dirichlet rows written by the fill kernel itself (no separate launch) */
  int     export_overlapped;   /* last evaluate: halo export ran under the uniform-tile kernel */
  int     n_edge_tiles;        /* lattice tiles with rows on their faces (mesh / rank boundary): k_fill_edge */
  int     reserved_i[2];
  double  setup_ms;            /* wall time of the last txasm_setup */
} txasm_info;

/* ------------------------------------------------------------------------------------------ */
/* life cycle                                                                                   */
int txasm_version(int *major, int *minor);
int txasm_create(const txasm_config *cfg, txasm_handle *out);
int txasm_destroy(txasm_handle h);
const char *txasm_last_error(txasm_handle h);   /* h may be NULL: last creation error */

/* ------------------------------------------------------------------------------------------ */
/* setup (once; the analogue of FieldManagerBuilder::setupVolumeFieldManagers + LOF construction) */

/* Element block = one Phalanx volume field manager and its worksets.
 * lids: panzer::GlobalIndexer::getLIDs() -- int[n_cells][dofs_per_cell], LayoutRight
 *       (dof-mgr/src/Panzer_GlobalIndexer.hpp:254-326,569-601).
 * cell_coords: the worksets' cell_vertex_coordinates double[n_cells][8][3] concatenated in cell
 *       order (disc-fe/src/Panzer_Workset_Builder_impl.hpp:153-187).  May be NULL when
 *       node_coords (double[n_rows][3], indexed by LID) is given instead. */
int txasm_block_add(txasm_handle h, int topology, int basis, int cubature_degree,
                    int64_t n_cells, int dofs_per_cell, const int *lids,
                    const double *cell_coords, const double *node_coords, int64_t n_rows);

/* General element blocks (BASELINE.json configs 3-5): any number of blocks per handle -- the reference loops over them at
 * Panzer_AssemblyEngine_impl.hpp:152 -- each with its own topology, basis and field layout:
 *   HEX8 / HGRAD_C1, HEX27 / HGRAD_C2, TET4 / HGRAD_C1, TET10 / HGRAD_C2  one scalar field, or three interleaved fields on HEX8
 *   HEX8 / HCURL_I1                                                        one edge-element field, orientation signs required
 * cell_vertex_coords: the worksets' cell_vertex_coordinates [n_cells][8 or 4][3] (geometry is always vertex based, as in
 *   panzer::Workset).  lids: GlobalIndexer::getLIDs() rows of this block [n_cells][dofs_per_cell].  field_offsets:
 *   getGIDFieldOffsets(block, field) for every field, [n_fields][basis cardinality]; NULL = FieldAggPattern's interleaving
 *   (dof-mgr/src/Panzer_FieldAggPattern.cpp:201-276): position = basis * n_fields + field.  orientation_signs: +-1 per
 *   (cell, edge), +1 when the edge's first Shards vertex has the smaller global vertex id (Panzer_IntrepidOrientation.cpp:96-99).
 * n_rows: local DOFs of the handle (all blocks share one LID space and one graph, set with txasm_graph_set).
 * Such a handle cannot also hold a txasm_block_add block; it assembles with searched atomic adds (sumIntoValues semantics:
 * a column absent from the row is skipped), the positions found once at txasm_setup. */
typedef struct {
  int topology, basis, cubature_degree;
  int64_t n_cells;
  const double *cell_vertex_coords;
  int n_fields, dofs_per_cell;
  const int *lids;
  const int *field_offsets;
  const signed char *orientation_signs;
} txasm_block_desc;
int txasm_gblock_add(txasm_handle h, const txasm_block_desc *desc, int64_t n_rows, int *block_id);

/* The block's integrands (one operator per block; multipliers as the equation sets pass them to the integrators):
 *   TXASM_OP_DIFFUSION   Integrator_GradBasisDotVector + Integrator_BasisTimesScalar (Example_PoissonEquationSet_impl.hpp:150-195)
 *                        params = {kappa, react (on x), mass_dot (on xdot), mass_dotdot (on xdotdot), constant source}
 *   TXASM_OP_ELASTICITY  three HGRAD fields u_i: rows int grad(phi) . sigma_i(u), sigma = lambda tr(eps) I + 2 mu eps, plus
 *                        rho int phi d2u_i/dt2 and c int phi du_i/dt -- the reference has no elasticity equation set
 *                        (SURVEY.md appendix B): operator defined here, parity unpinned.
 *                        params = {lambda, mu, rho, c, body force x, y, z}
 *   TXASM_OP_CURLCURL    Integrator_CurlBasisDotVector + Integrator_BasisTimesVector
 *                        (adapters-stk/example/CurlLaplacianExample/Example_CurlLaplacianEquationSet_impl.hpp:163-212)
 *                        params = {curl-curl multiplier, mass multiplier (on x), mass multiplier (on xdot), -, source x, y, z}
 * Jacobian seeds: beta for x, alpha for xdot, gamma for xdotdot. */
enum { TXASM_OP_DIFFUSION = 1, TXASM_OP_ELASTICITY = 2, TXASM_OP_CURLCURL = 3 };
int txasm_gblock_terms_set(txasm_handle h, int block_id, int op, const double *params, int n_params);

/* The ghosted local matrix exactly as Tpetra::CrsMatrix::getLocalMatrixDevice() exposes it
 * (graph.row_map, graph.entries): rows sorted by local column index
 * (disc-fe/src/lof/Panzer_TpetraLinearObjFactory_impl.hpp:558-650). */
int txasm_graph_set(txasm_handle h, int64_t n_rows, const int64_t *rowptr, const int *colind);

/* Alternative: build that graph on the device from the LID table (the work buildGhostedGraph does
 * with insertGlobalIndices + fillComplete).  Call with colind == NULL to obtain nnz. */
int txasm_graph_build(txasm_handle h, int64_t *nnz_out);
int txasm_graph_get(txasm_handle h, int64_t *rowptr, int *colind);
/* Rows [first_row, first_row + n_rows) of the graph to host arrays: rowptr[n_rows + 1] rebased to 0, colind (may be NULL).
 * The ghost rows are all the Import/Export negotiation needs on the host (a message of surface size). */
int txasm_graph_get_rows(txasm_handle h, int64_t first_row, int64_t n_rows, int64_t *rowptr, int *colind);
/* The fill graph on the device: insert the columns cols[i] into the rows rows[i] of the device-resident graph (pairs that
 * exist already or repeat are fine) -- what TpetraLinearObjFactory::buildGraph's Export(INSERT) does to the owned rows
 * (lof/Panzer_TpetraLinearObjFactory_impl.hpp:534-556).  pos[i] (host or device, may be NULL) = index of (rows[i], cols[i])
 * in the new A_values: the static plan of the matrix ADD of ghostToGlobalContainer (:151-205) for txasm_halo_set_matrix.
 * The handle must be set up again afterwards. */
int txasm_graph_merge_columns(txasm_handle h, int64_t n, const int *rows, const int *cols, int64_t *pos, int64_t *nnz_out);

int txasm_terms_set(txasm_handle h, const txasm_term *terms, int n_terms);

/* TianXin::DirichletEvalautor's nodeset DOFs (m_local_dofs, m_values):
 * disc-fe/src/evaluators/TianXin_Dirichlet_impl.hpp:59-81 */
int txasm_dirichlet_set(txasm_handle h, int n, const int *local_dofs, const double *values);

/* TianXin::CLoadEvalautor's concentrated loads (disc-fe/src/evaluators/TianXin_CLoad_impl.hpp:56-78 ->
 * TpetraLinearObjContainer::applyConcentratedLoad, lof/Panzer_TpetraLinearObjContainer.hpp:334-357):
 * Residual evaluation: f[local_dofs[i]] += values[i]; the Jacobian evaluator does nothing, as in the reference.
 * Applied in the BOUNDARY_FILL stage before the Dirichlet rows. */
int txasm_cload_set(txasm_handle h, int n, const int *local_dofs, const double *values);

/* TianXin::Flux (disc-fe/src/evaluators/TianXin_Neumann_impl.hpp:143-160) on side worksets: for every listed side
 * f[lid(cell,b)] += value * int_side phi_b dGamma  (2x2 Gauss on the face, Shards Hexahedron<8> side ordinals 0..5;
 * cells are indices into the block's cell list).  The value does not depend on x, so Residual and Jacobian type
 * evaluations add the same numbers to f and nothing to A.  Applied first in the BOUNDARY_FILL stage (the reference
 * order: Neumann, then Dirichlet -- Panzer_AssemblyEngine_impl.hpp:100-115). */
int txasm_neumann_set(txasm_handle h, int n_sides, const int *cells, const int *local_sides, const double *values);

/* Functional response: value = sum over this handle's cells of
 *   sum_qp scalar(cell,qp) * weighted_measure(cell,qp)            (panzer::Integrator_Scalar +
 *   ResponseScatterEvaluator_Functional, disc-fe/src/evaluators/Panzer_Integrator_Scalar_impl.hpp:117-142,
 *   disc-fe/src/responses/Panzer_ResponseScatterEvaluator_Functional_impl.hpp:137-142)
 * followed by the global sum over the communicator of txasm_comm_init (Response_Functional::scatterResponse), if any.
 * Integrands: TXASM_RESP_INTEGRAL the field itself; TXASM_RESP_L2_ERROR (A-B)^2 and TXASM_RESP_H1_ERROR
 * (A-B)^2 + |grad A - grad B|^2 -- the example's "L2 ERROR_CALC" / "H1 ERROR_CALC" closure models
 * (adapters-stk/example/PoissonExample/Example_ClosureModel_Factory_impl.hpp:147-275) -- with A the field of x at the
 * integration points and B the exact solution `solution_id` (TXASM_SOURCE_SIN3: sin2pix sin2piy sin2piz; 3: the
 * example's sin2pix sin2piy).  Tensor Gauss cubature of `cubature_degree` (the example uses 10).  x: ghosted
 * vector by LID, host or device.  The value (not its square root) is written to *value on the host; synchronous. */
enum { TXASM_RESP_INTEGRAL = 1, TXASM_RESP_L2_ERROR = 2, TXASM_RESP_H1_ERROR = 3,
       TXASM_RESP_IP_ARRAY = 4 /* internal: integrand given at the integration points (txasm_response_integral) */ };
int txasm_response_functional(txasm_handle h, int kind, int solution_id, int cubature_degree, const double *x, double *value);

/* TianXin::Response_Integral<Residual> (disc-fe/src/responses/TianXin_Response_Integral_impl.hpp:106-133): the integrand is
 * any field at the integration points, cell_ip_values[n_cells][n_qp] (n_qp = (cubature_degree/2+1)^3, x fastest);
 *   value = sum over cells and points of cellvalue(cell, qp) * weighted_measure(cell, qp), reduced over the communicator;
 * then response_vector[0] += value, as the reference's tVector_->sumIntoLocalValue(0, glbValue) does (response_vector is a
 * host array and must be given -- the reference throws "reponse vector not defined" otherwise).  *value (optional)
 * receives the global value.  The reference's Jacobian specialisation is unfinished (it adds 100.0 per LID, :224-230) and
 * is not reproduced. */
int txasm_response_integral(txasm_handle h, int cubature_degree, const double *cell_ip_values, double *response_vector, double *value);

/* Finalise: classify cells, build row tiles / adjacency / slot tables, size shared memory. */
int txasm_setup(txasm_handle h);

/* Run-time switches (tests cross-check the kernel variants against each other bit for bit; tuning).  Each has an
 * environment variable of the same meaning that sets the default when the handle is created.
 *   "uniform_kernel"  (TXASM_NO_UNIFORM_KERNEL=1 -> 0)  1: uniform tiles go to the lean kernels, 0: every tile takes k_fill_rowtile
 *   "brick_kernel"    (TXASM_NO_BRICK_KERNEL=1 -> 0)    1: brick tiles go to k_fill_brick, 0: to k_fill_uniform
 *   "export_overlap"  (TXASM_EXPORT_OVERLAP=0/1)        1: halo export on a side stream under the uniform-tile kernel
 *   "fuse_dirichlet"  (TXASM_NO_FUSE_DIRICHLET=1 -> 0)  1: evaluate(All) writes Dirichlet rows from the fill kernel
 *   "concurrent_fill" (TXASM_NO_CONCURRENT_FILL=1 -> 0) 1: boundary-tile kernels on a side stream beside the uniform-tile kernel
 *   "edge_kernel"     (TXASM_NO_EDGE_KERNEL=1 -> 0)     1: lattice tiles with rows on their faces (mesh / rank boundary) go to k_fill_edge
 *                                                        (entry-code streams), 0: to k_fill_rowtile
 *   "stage_timers"    (default 0)                        1: CUDA events between the stages of every evaluate (txasm_timers_get,
 *                                                        txasm_last_fill_ms); ~2 % of a 1 ms step, so off unless asked for
 *   "fill_event_ring" (default 0)                        N > 0: keep the fill spans of the last N evaluates (txasm_fill_ms_history)
 *   "block_atomic"    (default 1)                        general blocks: 1 = atomic adds at planned positions like ScatterResidual_Tpetra,
 *                                                        0 = element rows to scratch, then an owner-computes gather (no atomics, bitwise
 *                                                        reproducible; 1.4-1.9x slower as measured on B200)
 *   "dmma"            (default 1)                        1: Q2-hexahedron blocks form their element matrix on the FP64 tensor cores
 *   "halo_p2p"        (default 1)                        1: halo over peer memory once connected, 0: NCCL send/recv
 *   "grid_cap"        (default 0 = none)                 > 0: persistent fill kernels launch at most this many CTAs
 * Unknown names return TXASM_EINVAL.  Cheap; may be called between evaluates. */
int txasm_option_set(txasm_handle h, const char *name, int value);
int txasm_option_get(txasm_handle h, const char *name, int *value);
int txasm_info_get(txasm_handle h, txasm_info *info);

/* ------------------------------------------------------------------------------------------ */
/* the hot path                                                                                 */

/* AssemblyEngine<EvalT>::evaluate(in, flags)  (disc-fe/src/Panzer_AssemblyEngine_impl.hpp:65-129).
 * x, xdot, xdotdot: ghosted vectors (length n_rows = owned ++ ghosted); the owned prefix is the
 *   global container's vector, the ghost tail is filled by the INITIALIZE stage when a halo is set.
 * f: ghosted residual double[n_rows]; A_values: values of the ghosted CSR matrix double[nnz]
 *   (NULL for TXASM_RESIDUAL).
 * Dirichlet (BOUNDARY_FILL): Jacobian -> rows := identity and f = x - value
 *   (lof/Panzer_TpetraLinearObjContainer.hpp:228-237,306-317); Residual -> f = x - value. */
int txasm_evaluate(txasm_handle h, int eval_type, int flags, const txasm_inargs *in,
                   const double *x, const double *xdot, const double *xdotdot,
                   double *f, double *A_values);

/* Introspection of the row-tile tables (tests / tuning): rows[tile_rows_max] (row id or -1),
 * cells[tile_cells_max] (cell id or -1, in staging order), adjl[tile_rows_max*8] (staging position of the
 * cell that has the row as local vertex a, 0xFFFF if none).  *n_cells_out = cells used by the tile. */
int txasm_tile_get(txasm_handle h, int tile, int *rows, int *cells, unsigned short *adjl, int *n_cells_out);

/* Diagnostic (environment TXASM_TIMELINE=1): start / end of k_fill_brick, k_fill_edge, k_fill_rowtile inside the last
 * evaluate, microseconds from the earliest start (%globaltimer stamps by one thread per CTA); -1: the kernel did not run. */
int txasm_debug_timeline(txasm_handle h, double out[6]);

int txasm_sync(txasm_handle h);
/* Stage timers of the last evaluate.  Needs txasm_option_set(h, "stage_timers", 1) BEFORE that evaluate: the CUDA events
 * between the stages are off by default (nine event records cost ~2 % of a 1 ms step); TXASM_ESTATE otherwise. */
int txasm_timers_get(txasm_handle h, txasm_timers *t);
/* device time (ms) of the dominant fill kernel in the last evaluate, measured with CUDA events
 * on the handle's stream */
int txasm_last_fill_ms(txasm_handle h, double *ms);
/* With txasm_option_set(h, "fill_event_ring", R): the fill time (as txasm_last_fill_ms) of each of the last
 * min(R, cap, evaluates since) evaluates, oldest first, without a synchronisation between them -- the average launch
 * duration of the fill over a timed region of back-to-back evaluates.  Synchronises the stream. */
int txasm_fill_ms_history(txasm_handle h, double *ms, int cap, int *n);

/* Diagnostic: DFMA throughput of the device in TFLOP/s (dependent-free chains on every SM, CUDA events on the
 * handle's stream) -- the second ceiling of the general-hexahedron fill. */
int txasm_measure_fp64_peak(txasm_handle h, double *tflops);

/* ------------------------------------------------------------------------------------------ */
/* multi-GPU: replaces Tpetra Import/Export (lof/Panzer_TpetraLinearObjFactory_impl.hpp:124-219) */

/* NCCL bootstrap.  Rank 0 calls txasm_comm_unique_id and distributes the 128 bytes. */
int txasm_comm_unique_id(void *id128);
int txasm_comm_init(txasm_handle h, int nranks, int rank, const void *id128);

/* Neighbour lists.  For neighbour k (rank nbr_rank[k]):
 *   import: I receive x for my ghost LIDs recv_lids[recv_off[k]..recv_off[k+1]) and the neighbour
 *           sends me its owned LIDs listed in ITS send list (same order, by GID);
 *   export: the reverse direction carries f (ADD at the owner) and ghost-row Jacobian values.
 * send_lids are owned LIDs of mine that neighbour k ghosts, in the order k expects them. */
int txasm_halo_set(txasm_handle h, int64_t n_owned, int n_nbr, const int *nbr_rank,
                   const int64_t *send_off, const int *send_lids,
                   const int64_t *recv_off, const int *recv_lids);
/* For the matrix export: for every ghost row I send (recv_lids order) the values of the whole
 * row; the owner adds entry j of that row into A_values[row_map[...]].  recv side map:
 *   mat_send_rowlen is implied by the graph; mat_recv_pos[k-range] gives, for every value the
 *   owner receives from neighbour k, the destination index into A_values or -1 (column absent). */
int txasm_halo_set_matrix(txasm_handle h, const int64_t *mat_recv_off, const int64_t *mat_recv_pos);

/* Peer-memory exchange (NVLink / NVSwitch, one process per GPU): after txasm_halo_set + txasm_halo_set_matrix every rank
 * calls txasm_halo_p2p_export (allocates its receive slab, fills a blob of txasm_halo_p2p_blob_size() bytes holding the
 * cudaIpc handle and the segment of each neighbour), the host all-gathers the blobs (rank-major) and every rank calls
 * txasm_halo_p2p_connect.  From then on the INITIALIZE / SCATTER stages push straight into the neighbours' slabs with
 * plain stores and flag the epoch; NCCL is only used by txasm_response_functional's all-reduce.  Option "halo_p2p" = 0
 * switches back to grouped ncclSend/ncclRecv.  txasm_halo_p2p_status: *timed_out = k+1 when neighbour k never delivered. */
int txasm_halo_p2p_blob_size(void);
int txasm_halo_p2p_export(txasm_handle h, void *blob);
int txasm_halo_p2p_connect(txasm_handle h, int nranks, const void *blobs);
int txasm_halo_p2p_status(txasm_handle h, int *timed_out);

#ifdef __cplusplus
}
#endif
#endif /* TXASM_H */
