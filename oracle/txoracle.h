/*
 * txoracle.h -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A deliberately naive, unfused, host-only restatement of the finite-element
 * assembly hot path of hillyuan/Tianxin (Panzer fork), written from the
 * reference's algorithm description (file:line cited at every function in
 * txoracle.c).  It is NOT part of the product: only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it,
 * and only as the checker / CPU baseline.  Nothing under tianxin_b200/ links,
 * imports or executes anything in oracle/.
 *
 * Parity status: the reference itself cannot be compiled here (needs Trilinos:
 * Kokkos, Tpetra, Phalanx, Sacado, Intrepid2, STK + MPI; none present).  The
 * oracle is pinned against every golden vector the reference's own tests hold
 * for this path (tests/test_oracle_golden.py lists them one by one):
 *   - GIDs / connectivity: tCubeHexMeshDOFManager.cpp:147-198,
 *     tSquareQuadMeshDOFManager.cpp:145-200,328-376,412-446
 *   - basis / geometry identities: basis_values2.cpp:264-286,
 *     integration_values2.cpp:106-115
 *   - fake-indexer owned/ghosted sizes: UnitTest_GlobalIndexer.cpp:122-270
 * Element-matrix VALUES have no stored golden in the reference ("pinned by
 * identities, not by vectors").
 */
#ifndef TXORACLE_H
#define TXORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int nx, ny, nz;      /* "X/Y/Z Elements" (one block) */
  int px, py, pz;      /* "X/Y/Z Procs" */
  double x0, xf, y0, yf, z0, zf;
} orc_mesh_params;

/* ---- mesh: Panzer_STK_CubeHexMeshFactory.cpp ---- */
int     orc_default_proc_grid(int nranks, int *px, int *py, int *pz);
int64_t orc_mesh_num_elems(const orc_mesh_params *p, int rank);
int     orc_mesh_build(const orc_mesh_params *p, int rank,
                       int64_t *elem_ids,    /* [ne]        stk element ids (1-based), ascending */
                       int64_t *elem_nodes,  /* [ne][8]     stk node ids (1-based), Shards Hex8 order */
                       double  *cell_coords  /* [ne][8][3]  cell_vertex_coordinates */);

/* ---- DOF numbering: Panzer_DOFManager.cpp (GUN), all ranks simulated in one process ---- */
typedef struct orc_dofs orc_dofs;
orc_dofs *orc_dofs_create(int nranks, int ids_per_elem, int nfields);
void      orc_dofs_destroy(orc_dofs *d);
/* conn = STKConnManager connectivity ids (Panzer_STKConnManager.cpp:201-226), [ne][ids_per_elem] */
int       orc_dofs_set_conn(orc_dofs *d, int rank, int64_t ne, const int64_t *conn);
int       orc_dofs_build(orc_dofs *d);
int64_t   orc_dofs_num_elems(const orc_dofs *d, int rank);
int64_t   orc_dofs_num_owned(const orc_dofs *d, int rank);
int64_t   orc_dofs_num_ghosted(const orc_dofs *d, int rank);
int       orc_dofs_gids_per_elem(const orc_dofs *d);
int       orc_dofs_get_elem_gids(const orc_dofs *d, int rank, int64_t *out /*[ne][nfields*ids_per_elem]*/);
int       orc_dofs_get_elem_lids(const orc_dofs *d, int rank, int *out);
int       orc_dofs_get_owned(const orc_dofs *d, int rank, int64_t *out);
int       orc_dofs_get_ghosted(const orc_dofs *d, int rank, int64_t *out);
/* getGIDFieldOffsets(block, field): Panzer_FieldAggPattern.cpp:201-276 */
int       orc_dofs_field_offsets(const orc_dofs *d, int field, int *out /*[ids_per_elem]*/);

/* ---- ghosted graph: Panzer_TpetraLinearObjFactory_impl.hpp:558-650 ---- */
/* two-call protocol: first with colind==NULL to get rowptr (and nnz = rowptr[n_rows]) */
int orc_ghosted_graph(int64_t ne, int n_per_elem, const int *lids, int n_rows,
                      int64_t *rowptr /*[n_rows+1]*/, int *colind /* or NULL */);

/* ---- geometry + basis tables for Q1 hex, 2x2x2 Gauss: IntegrationValues2 / BasisValues2 ---- */
typedef struct {
  int64_t ne;
  double *jac;          /* [ne][8][3][3]  jac(c,q,d,e) = dx_d/dxi_e */
  double *jac_inv;      /* [ne][8][3][3] */
  double *jac_det;      /* [ne][8] */
  double *wm;           /* [ne][8]        weighted_measure */
  double *ip;           /* [ne][8][3]     ip_coordinates */
  double *basis;        /* [ne][8][8]     basis_scalar(c,b,q) */
  double *wbasis;       /* [ne][8][8]     weighted_basis_scalar */
  double *gbasis;       /* [ne][8][8][3]  grad_basis(c,b,q,d) */
  double *wgbasis;      /* [ne][8][8][3]  weighted_grad_basis */
} orc_tables;
void orc_ref_cubature(double *pts /*[8][3]*/, double *wts /*[8]*/);
void orc_ref_basis(const double *pt /*[3]*/, double *val /*[8]*/, double *grad /*[8][3]*/);
int  orc_tables_build(int64_t ne, const double *cell_coords, orc_tables *t);  /* caller allocates */

/* ---- the assembly pipeline ---- */
typedef struct {
  int    eval_type;        /* 0 = Residual, 1 = Jacobian */
  int    workset_size;     /* 20 in every reference driver */
  double alpha, beta;      /* seeds: alpha for dxdt gathers, beta for x gathers */
  double kappa;            /* multiplier of  int grad(v).grad(T)        (thermal_conductivity) */
  double mass_dot;         /* multiplier of  int v Tdot   (0 = no transient term) */
  double react;            /* multiplier of  int v T      (0 = none; used by the fe_assembly identity) */
  double source_mult;      /* multiplier of  int v s      (-1 in the Poisson example; 0 = none) */
  int    source_id;        /* 0 none, 1: 12 pi^2 sin2pix sin2piy sin2piz, 2: constant 1,
                              3: 8 pi^2 sin2pix sin2piy (2-D example source) */
  int    nthreads;         /* OpenMP threads over worksets (1 = serial, deterministic) */
  double gamma;            /* seed of the D2XDT2 gather (extension, parity unpinned) */
  double mass_dotdot;      /* multiplier of int phi d2T/dt2 */
  const double *fm_grad;   /* product of the "Field Multipliers" of Integrator_GradBasisDotVector at the IPs [ne][8], or NULL
                              (disc-fe/src/evaluators/Panzer_Integrator_GradBasisDotVector_impl.hpp:257-294) */
  const double *fm_mass;   /* ... of the Integrator_BasisTimesScalar terms on the solution fields, or NULL (:268-298) */
} orc_terms;

int orc_evaluate_volume(const orc_terms *terms, int64_t ne, const int *lids /*[ne][8]*/,
                        const orc_tables *t,
                        const double *x /*[n_local]*/, const double *xdot /* or NULL */,
                        int n_rows, const int64_t *rowptr, const int *colind,
                        double *f /*[n_local] accumulated into*/, double *A /*[nnz] accumulated into, NULL for residual*/);

int orc_evaluate_volume2(const orc_terms *terms, int64_t ne, const int *lids, const orc_tables *t,
                         const double *x, const double *xdot, const double *xdotdot /* or NULL */,
                         int n_rows, const int64_t *rowptr, const int *colind, double *f, double *A);

/* TianXin_Dirichlet_impl.hpp:59-81 + TpetraLinearObjContainer.hpp:228-237,306-317 */
int orc_dirichlet(int eval_type, int n, const int *local_dofs, const double *values,
                  const double *x, double *f, const int64_t *rowptr, const int *colind, double *A);

/* TianXin_CLoad_impl.hpp:56-78 */
int orc_dirichlet_rows_and_columns(int n, const int *local_dofs, int64_t n_rows, const int64_t *rowptr, const int *colind, double *A);
int orc_cload(int eval_type, int n, const int *local_dofs, const double *values, double *f);

/* TianXin_Neumann_impl.hpp:143-160 (Flux) on side worksets */
int orc_neumann_flux(int64_t n_sides, const int *cells, const int *sides, const double *values, const int *lids,
                     const double *cell_coords, double *f);

/* Functional responses: Panzer_Integrator_Scalar_impl.hpp:117-142 + ResponseScatterEvaluator_Functional (rank-local sum) */
int orc_gauss_legendre(int n, double *x, double *w);
int orc_response_functional(int kind /*1 integral of the field, 2 L2 error^2, 3 H1 error^2*/, int solution_id /*1: 3-D sine product, 3: the example's 2-D one*/,
                            int cub_degree, int64_t ne, const int *lids, const double *cell_coords, const double *x, double *value);

/* in-process Tpetra Import/Export restatement: TpetraLinearObjFactory_impl.hpp:124-219 */
int orc_response_integral(int64_t ne, int nq, const double *cellvalue, const double *wm, double *response_vector, double *value);
int orc_global_to_ghost(const orc_dofs *d, const double *const *x_owned /*[nranks]*/, int rank, double *x_ghosted);
int orc_ghost_to_global_vec(const orc_dofs *d, const double *const *f_ghosted /*[nranks]*/, int rank, double *f_owned);

/* ======================================================================== */
/* Element blocks beyond the scalar Q1 hexahedron (txblocks.c): BASELINE.json configs 3-5               */
/* ======================================================================== */
/* elem: 1 HEX8 / HGRAD C1, 2 HEX27 / HGRAD C2, 3 TET4 / HGRAD C1, 4 TET10 / HGRAD C2, 5 HEX8 / HCURL I1
 * op:   1 scalar diffusion   p = {kappa, react, mass_dot, mass_dotdot, constant source}
 *       2 linear elastodynamics (3 interleaved fields)  p = {lambda, mu, rho (on d2u/dt2), damping (on du/dt), body force[3]}
 *       3 curl-curl + mass (HCURL)  p = {curl multiplier, mass multiplier, mass_dot multiplier, -, source[3]}          */
typedef struct {
  int elem, op, cub_degree, eval_type;
  double alpha, beta, gamma;
  double p[8];
} orb_spec;
int  orb_num_basis(int elem);
int  orb_num_vertices(int elem);
void orb_ref_basis(int elem, const double *pt, double *val, double *der);
int  orb_cubature(int elem, int deg, double *pts, double *wts);
int  orb_evaluate(const orb_spec *sp, int64_t ne, const double *cell_coords /*[ne][nv][3]*/, int ndof, const int *lids /*[ne][ndof]*/,
                  const int *field_offsets /*[nfields][nb] or NULL = interleaved*/, const signed char *signs /*HCURL [ne][12] or NULL*/,
                  const double *x, const double *xdot, const double *xdotdot,
                  const int64_t *rowptr, const int *colind, double *f, double *A);
int  orb_q2_hex_lids(int nx, int ny, int nz, int *lids);
int  orb_hcurl_hex_lids(int nx, int ny, int nz, int *lids, signed char *signs);

#ifdef __cplusplus
}
#endif
#endif
