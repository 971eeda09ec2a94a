/*
 * txhost.h -- C interface of the host-side mirror (libtxhost.so, C++17, no CUDA).
 *
 * The reference's host code (adapters-stk mesh factory + STKConnManager, dof-mgr DOFManager,
 * disc-fe TpetraLinearObjFactory) stays as it is in a real deployment and calls txasm.h.  It cannot
 * be built here (needs Trilinos + MPI), so tests and the bench drive the C ABI through this
 * stand-in, which mirrors the reference classes one to one:
 *
 *   txhost_mesh     panzer_stk::CubeHexMeshFactory + STK_Interface + STKConnManager
 *                   (adapters-stk/src/stk_interface/Panzer_STK_CubeHexMeshFactory.cpp,
 *                    adapters-stk/src/Panzer_STKConnManager.cpp)
 *   txhost_dofmgr   panzer::DOFManager / GlobalIndexer (dof-mgr/src/Panzer_DOFManager.cpp)
 *   txhost_lof      panzer::TpetraLinearObjFactory (disc-fe/src/lof/Panzer_TpetraLinearObjFactory_impl.hpp):
 *                   ghosted graph, Import/Export plans
 *
 * Distributed algorithms are written as state machines around ONE primitive, an all-to-all of
 * int64 records (what Tpetra's Directory / Import / Export do over MPI): call *_step with what was
 * received, get what to send next, until *done.  The caller moves the data (torch.distributed in
 * this repo, MPI in the reference); with one rank the exchange is a local copy.
 */
#ifndef TXHOST_H
#define TXHOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct txhost_mesh_s *txhost_mesh;
typedef struct txhost_dofmgr_s *txhost_dofmgr;
typedef struct txhost_lof_s *txhost_lof;

const char *txhost_last_error(void);
/* OpenMP threads of the host mirror (launchers export OMP_NUM_THREADS=1 per rank); returns the count now in use */
int     txhost_set_num_threads(int n);

/* ---- CubeHexMeshFactory: "X/Y/Z Elements", "X/Y/Z Procs" (-1,-1,-1 = factory default grid,
 *      px=-1 only = x slabs), "X0".."Zf"; one element block eblock-0_0_0 */
txhost_mesh txhost_cube_hex_mesh(int nx, int ny, int nz, int px, int py, int pz,
                                 double x0, double xf, double y0, double yf, double z0, double zf,
                                 int rank, int nranks);
/* The rank's brick of elements and the processor grid without building the mesh arrays: out[9] = {xs, xn, ys, yn, zs, zn,
 * px, py, pz} (Panzer_STK_CubeHexMeshFactory.cpp:89-136, 463-535, 918-927).  tianxin_b200/device_setup.py forms ids,
 * connectivity and coordinates from these on the GPU. */
int     txhost_cube_hex_brick(int nx, int ny, int nz, int px, int py, int pz, int rank, int nranks, int64_t *out);
void    txhost_mesh_destroy(txhost_mesh m);
int64_t txhost_mesh_num_elems(txhost_mesh m);
int     txhost_mesh_proc_grid(txhost_mesh m, int *px, int *py, int *pz);
int     txhost_mesh_get(txhost_mesh m, int64_t *elem_ids, int64_t *elem_nodes /*[ne][8]*/, double *cell_coords /*[ne][8][3]*/);
/* STKConnManager::buildConnectivity for a nodal pattern: connectivity id = stk node id - 1 */
int     txhost_mesh_connectivity(txhost_mesh m, int64_t *conn /*[ne][8]*/);
/* synthetic perturbation of interior nodes: amp*h*(u-1/2), splitmix64 */
int     txhost_mesh_perturb(txhost_mesh m, double amp);
/* side sets left/right (x), front/back (y), bottom/top (z) as in CubeHexMeshFactory.cpp:352-357:
 * node ids (stk, 1-based) of my elements lying on that side; returns count (out may be NULL) */
int64_t txhost_mesh_sideset_nodes(txhost_mesh m, const char *name, int64_t *out);

/* ---- DOFManager (GUN numbering).  All fields nodal CG on the same ids. */
txhost_dofmgr txhost_dofmgr_create(int rank, int nranks, int ids_per_elem, int nfields);
/* A finished DOF manager from results computed elsewhere (device_setup.py): owned / ghosted GIDs in LID order, ghost owners,
 * first GID of this rank.  Enough for TpetraLinearObjFactory in compact mode (txhost_lof_set_ghost_rows). */
txhost_dofmgr txhost_dofmgr_from_arrays(int rank, int nranks, int ids_per_elem, int nfields, int64_t n_owned, const int64_t *owned,
                                        int64_t n_ghosted, const int64_t *ghosted, const int *ghosted_owner, int64_t my_offset);
void    txhost_dofmgr_destroy(txhost_dofmgr d);
int     txhost_dofmgr_set_connectivity(txhost_dofmgr d, int64_t ne, const int64_t *conn);
/* buildGlobalUnknowns as a state machine.  recv_counts[P], recv = concatenated records received in
 * the previous exchange (ignored on the first call).  On return *send_counts / *send point to
 * library-owned buffers valid until the next call; *done = 1 when finished (nothing to send). */
int     txhost_dofmgr_step(txhost_dofmgr d, const int64_t *recv_counts, const int64_t *recv,
                           const int64_t **send_counts, const int64_t **send, int *done);
int64_t txhost_dofmgr_num_owned(txhost_dofmgr d);
int64_t txhost_dofmgr_num_ghosted(txhost_dofmgr d);
int     txhost_dofmgr_get_owned(txhost_dofmgr d, int64_t *out);          /* getOwnedIndices */
int     txhost_dofmgr_get_ghosted(txhost_dofmgr d, int64_t *out);        /* getGhostedIndices */
int     txhost_dofmgr_get_ghosted_owner(txhost_dofmgr d, int *out);      /* owning rank of each ghosted index */
int     txhost_dofmgr_get_elem_gids(txhost_dofmgr d, int64_t *out);      /* getElementGIDs, [ne][gpe] */
int     txhost_dofmgr_get_elem_lids(txhost_dofmgr d, int *out);          /* getLIDs(), [ne][gpe] LayoutRight */
int     txhost_dofmgr_field_offsets(txhost_dofmgr d, int field, int *out); /* getGIDFieldOffsets */

/* ---- TpetraLinearObjFactory: ghosted graph + Import/Export plans */
txhost_lof txhost_lof_create(txhost_dofmgr d);
void    txhost_lof_destroy(txhost_lof l);
/* buildGhostedGraph: rows/cols = owned++ghosted, rows sorted by local column */
int     txhost_lof_ghosted_graph(txhost_lof l, int64_t *nnz);
int     txhost_lof_get_ghosted_graph(txhost_lof l, int64_t *rowptr, int *colind);
/* adopt a ghosted graph built elsewhere (e.g. on the device by txasm_graph_build) */
int     txhost_lof_set_ghosted_graph(txhost_lof l, const int64_t *rowptr, const int *colind);
/* Compact mode: only the ghost rows [n_owned, n_local) of the ghosted graph (rowptr[n_ghost + 1] rebased to 0) -- all the
 * negotiation needs.  The plan then has no fill graph / matrix positions; it carries every received (owned row, local
 * column) pair in plan order (txhost_lof_get_pairs) for txasm_graph_merge_columns, which holds the graph on the device. */
int     txhost_lof_set_ghost_rows(txhost_lof l, const int64_t *rowptr, const int *colind);
int64_t txhost_lof_num_pairs(txhost_lof l);
int     txhost_lof_get_pairs(txhost_lof l, int *rows, int *cols);
/* Plan construction (one exchange).  Afterwards the "fill graph" is available: the ghosted graph
 * whose OWNED rows also carry the columns other ranks contribute (the global matrix's columns,
 * buildGraph's Export INSERT), remote-only columns numbered n_local, n_local+1, ... */
int     txhost_lof_step(txhost_lof l, const int64_t *recv_counts, const int64_t *recv,
                        const int64_t **send_counts, const int64_t **send, int *done);
int     txhost_lof_num_neighbors(txhost_lof l);
int     txhost_lof_get_halo(txhost_lof l, int *nbr_rank, int64_t *send_off, int *send_lids,
                            int64_t *recv_off, int *recv_lids);     /* pass NULLs to query sizes via *_sizes */
int     txhost_lof_halo_sizes(txhost_lof l, int64_t *n_send, int64_t *n_recv, int64_t *n_mat_recv,
                              int64_t *fill_nnz, int64_t *n_cols);
int     txhost_lof_get_fill_graph(txhost_lof l, int64_t *rowptr, int *colind, int64_t *col_gids /*[n_cols]*/);
int     txhost_lof_get_matrix_plan(txhost_lof l, int64_t *mat_recv_off /*[n_nbr+1]*/, int64_t *mat_recv_pos);

#ifdef __cplusplus
}
#endif
#endif
