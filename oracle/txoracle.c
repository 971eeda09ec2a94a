/*
 * txoracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY (see txoracle.h).
 *
 * Restates, function by function, the reference algorithm for the assembly hot
 * path.  All paths below are relative to the reference tree.  The code is meant
 * to be read next to the reference, not to be fast: one loop per reference
 * loop, one temporary per reference field, forward-mode derivative arrays of
 * length N = DOFs per element exactly as Sacado::Fad::DFad<double> would carry.
 *
 * Third-party arithmetic that is NOT in the reference tree (Trilinos, version
 * unpinned -- README.md:13): Intrepid2 Basis_HGRAD_HEX_C1_FEM, tensor Gauss
 * cubature, CellTools::setJacobian/Inv/Det, FunctionSpaceTools::
 * HGRADtransformGRAD / multiplyMeasure / computeCellMeasure, Sacado DFad,
 * KokkosSparse sumIntoValues, Tpetra createOneToOne / CrsGraph::fillComplete.
 * Their published algorithms are restated here and anchored on the reference's
 * call sites; the conventions are pinned by the reference's own unit tests
 * (tests/test_oracle_golden.py).
 */
#include "txoracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ======================================================================== */
/* A. Inline cube mesh: adapters-stk/src/stk_interface/Panzer_STK_CubeHexMeshFactory.cpp */
/* ======================================================================== */

/* :89-133 -- default processor grid when "X/Y/Z Procs" are all -1 ("copied from galeri") */
int orc_default_proc_grid(int nranks, int *px, int *py, int *pz)
{
  int x, y, z;
  x = y = z = (int)pow((double)nranks, 0.333334);
  if (x * y * z != nranks) {
    enum { maxFactor = 50 };
    int factors[maxFactor];
    int ProcTemp = nranks;
    x = y = z = 1;
    for (int jj = 0; jj < maxFactor; jj++) factors[jj] = 0;
    for (int jj = 2; jj < maxFactor; jj++) {
      int flag = 1;
      while (flag) {
        int temp = ProcTemp / jj;
        if (temp * jj == ProcTemp) { factors[jj]++; ProcTemp = temp; }
        else flag = 0;
      }
    }
    x = ProcTemp;
    for (int jj = maxFactor - 1; jj > 0; jj--) {
      while (factors[jj] != 0) {
        if ((x <= y) && (x <= z)) x = x * jj;
        else if ((y <= x) && (y <= z)) y = y * jj;
        else z = z * jj;
        factors[jj]--;
      }
    }
  }
  *px = x; *py = y; *pz = z;
  return 0;
}

/* :918-927 procRankToProcTuple */
static void rank_to_tuple(const orc_mesh_params *p, int rank, int *i, int *j, int *k)
{
  *k = rank / (p->px * p->py); rank = rank % (p->px * p->py);
  *j = rank / p->px;           rank = rank % p->px;
  *i = rank;
}

/* :463-535 determine{X,Y,Z}ElemSizeAndStart -- first "extra" procs get one more layer */
static void size_and_start(int nelems, int size, int loc, int64_t *start, int64_t *nume)
{
  int64_t minElements = nelems / size;
  int64_t extra = nelems - minElements * size;
  if (loc < extra) { *nume = minElements + 1; *start = loc * (minElements + 1); }
  else { *nume = minElements; *start = extra * (minElements + 1) + (loc - extra) * minElements; }
}

/* Panzer_STK_MeshFactory.hpp:161-168 getMeshCoord */
static double mesh_coord(int64_t nx, double deltaX, double x0)
{
  double x = (double)nx * deltaX;
  double modX = fabs(x), modX0 = fabs(x0);
  double val = x + x0;
  if ((x0 * x < 0.0) && (fabs(modX - modX0) < DBL_EPSILON * modX0)) val = 0.0;
  return val;
}

int64_t orc_mesh_num_elems(const orc_mesh_params *p, int rank)
{
  int i, j, k; int64_t s, nx, ny, nz;
  rank_to_tuple(p, rank, &i, &j, &k);
  size_and_start(p->nx, p->px, i, &s, &nx);
  size_and_start(p->ny, p->py, j, &s, &ny);
  size_and_start(p->nz, p->pz, k, &s, &nz);
  return nx * ny * nz;
}

/* :401-461 buildBlock.  Node id = nz(NY+1)(NX+1)+ny(NX+1)+nx+1 (:433); element id =
 * NX*NY*nz+NX*ny+nx+1 (:446); Hex8 node order (:447-455).  Local element order =
 * ascending element id (STK_Interface::buildLocalElementIDs over get_selected_entities,
 * Panzer_STK_Interface.cpp:2342-2365 -- assumption A1 of SURVEY.md appendix A). */
int orc_mesh_build(const orc_mesh_params *p, int rank, int64_t *elem_ids, int64_t *elem_nodes, double *cell_coords)
{
  int pi, pj, pk; int64_t xs, ys, zs, xn, yn, zn;
  rank_to_tuple(p, rank, &pi, &pj, &pk);
  size_and_start(p->nx, p->px, pi, &xs, &xn);
  size_and_start(p->ny, p->py, pj, &ys, &yn);
  size_and_start(p->nz, p->pz, pk, &zs, &zn);
  const int64_t NX = p->nx, NY = p->ny;
  const double dX = (p->xf - p->x0) / (double)p->nx;
  const double dY = (p->yf - p->y0) / (double)p->ny;
  const double dZ = (p->zf - p->z0) / (double)p->nz;
  int64_t e = 0;
  for (int64_t nz = zs; nz < zs + zn; ++nz)
    for (int64_t ny = ys; ny < ys + yn; ++ny)
      for (int64_t nx = xs; nx < xs + xn; ++nx, ++e) {
        int64_t n[8];
        n[0] = nx + 1 + ny * (NX + 1) + nz * (NY + 1) * (NX + 1);
        n[1] = n[0] + 1;
        n[2] = n[1] + (NX + 1);
        n[3] = n[2] - 1;
        n[4] = n[0] + (NY + 1) * (NX + 1);
        n[5] = n[1] + (NY + 1) * (NX + 1);
        n[6] = n[2] + (NY + 1) * (NX + 1);
        n[7] = n[3] + (NY + 1) * (NX + 1);
        if (elem_ids) elem_ids[e] = NX * NY * nz + NX * ny + nx + 1;
        for (int a = 0; a < 8; ++a) {
          if (elem_nodes) elem_nodes[e * 8 + a] = n[a];
          if (cell_coords) {
            int64_t id0 = n[a] - 1;
            int64_t ix = id0 % (NX + 1), iy = (id0 / (NX + 1)) % (NY + 1), iz = id0 / ((NX + 1) * (NY + 1));
            cell_coords[(e * 8 + a) * 3 + 0] = mesh_coord(ix, dX, p->x0);
            cell_coords[(e * 8 + a) * 3 + 1] = mesh_coord(iy, dY, p->y0);
            cell_coords[(e * 8 + a) * 3 + 2] = mesh_coord(iz, dZ, p->z0);
          }
        }
      }
  return 0;
}

/* ======================================================================== */
/* B/C. DOF numbering: dof-mgr/src/Panzer_DOFManager.cpp                     */
/* ======================================================================== */

struct orc_dofs {
  int nranks, ipe, nfields;
  int64_t *ne;          /* [nranks] */
  int64_t **conn;       /* [nranks] -> [ne][ipe] */
  /* results */
  int64_t **ov;         /* sorted unique overlap ids per rank (std::set order, :1261-1290) */
  int64_t *n_ov;
  int64_t **ov_gid0;    /* first GID (field 0) of each overlap id */
  int64_t **egids;      /* [ne][ipe*nfields] */
  int **elids;
  int64_t **owned; int64_t *n_owned;
  int64_t **ghosted; int64_t *n_ghosted;
  int built;
};

orc_dofs *orc_dofs_create(int nranks, int ids_per_elem, int nfields)
{
  orc_dofs *d = (orc_dofs *)calloc(1, sizeof(orc_dofs));
  d->nranks = nranks; d->ipe = ids_per_elem; d->nfields = nfields;
  d->ne = (int64_t *)calloc(nranks, sizeof(int64_t));
  d->conn = (int64_t **)calloc(nranks, sizeof(int64_t *));
  d->ov = (int64_t **)calloc(nranks, sizeof(int64_t *));
  d->n_ov = (int64_t *)calloc(nranks, sizeof(int64_t));
  d->ov_gid0 = (int64_t **)calloc(nranks, sizeof(int64_t *));
  d->egids = (int64_t **)calloc(nranks, sizeof(int64_t *));
  d->elids = (int **)calloc(nranks, sizeof(int *));
  d->owned = (int64_t **)calloc(nranks, sizeof(int64_t *));
  d->n_owned = (int64_t *)calloc(nranks, sizeof(int64_t));
  d->ghosted = (int64_t **)calloc(nranks, sizeof(int64_t *));
  d->n_ghosted = (int64_t *)calloc(nranks, sizeof(int64_t));
  return d;
}

void orc_dofs_destroy(orc_dofs *d)
{
  if (!d) return;
  for (int r = 0; r < d->nranks; ++r) {
    free(d->conn[r]); free(d->ov[r]); free(d->ov_gid0[r]); free(d->egids[r]);
    free(d->elids[r]); free(d->owned[r]); free(d->ghosted[r]);
  }
  free(d->ne); free(d->conn); free(d->ov); free(d->n_ov); free(d->ov_gid0); free(d->egids);
  free(d->elids); free(d->owned); free(d->n_owned); free(d->ghosted); free(d->n_ghosted);
  free(d);
}

int orc_dofs_set_conn(orc_dofs *d, int rank, int64_t ne, const int64_t *conn)
{
  if (rank < 0 || rank >= d->nranks) return -1;
  free(d->conn[rank]);
  d->ne[rank] = ne;
  d->conn[rank] = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ne * d->ipe + 1));
  memcpy(d->conn[rank], conn, sizeof(int64_t) * (size_t)(ne * d->ipe));
  return 0;
}

static int cmp_i64(const void *a, const void *b)
{
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return (x < y) ? -1 : (x > y);
}
typedef struct { int64_t id; int rank; } id_rank;
static int cmp_id_rank(const void *a, const void *b)
{
  const id_rank *x = (const id_rank *)a, *y = (const id_rank *)b;
  if (x->id != y->id) return (x->id < y->id) ? -1 : 1;
  return (x->rank < y->rank) ? -1 : (x->rank > y->rank);
}
static int64_t find_i64(const int64_t *v, int64_t n, int64_t key)
{
  int64_t lo = 0, hi = n - 1;
  while (lo <= hi) {
    int64_t mid = (lo + hi) / 2;
    if (v[mid] == key) return mid;
    if (v[mid] < key) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

/* DOFManager::buildGlobalUnknowns (:474-714) for nodal CG fields that all live on
 * the same ids (nfields DOFs on every connectivity id):
 *  - overlap map = std::set of my elements' ids, ascending (:1261-1290)
 *  - createOneToOne with GreedyTieBreak: owner = smallest rank holding the id
 *    (:108-133, :743-748); the one-to-one map keeps the overlap map's order
 *    restricted to the ids I own (Tpetra behaviour, assumption A2)
 *  - local count -> exclusive scan -> my offset (:794-813)
 *  - GID assignment: rows of the non-overlap MV in order, fields in order,
 *    which_id += ndof (:823-843)
 *  - import back, fill elementGIDs_: per connectivity id, per field (:1293-1349)
 *  - owned_: first touch over my elements in order, owned ids only (:580-636)
 *  - ghosted_: first touch of everything not in owned_ (:650-695)
 *  - LIDs = position in owned_ ++ ghosted_ (Panzer_GlobalIndexer.hpp:604-638)
 */
int orc_dofs_build(orc_dofs *d)
{
  const int P = d->nranks, ipe = d->ipe, nf = d->nfields;
  int64_t total = 0;
  for (int r = 0; r < P; ++r) {
    int64_t n = d->ne[r] * ipe;
    int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    memcpy(tmp, d->conn[r], sizeof(int64_t) * (size_t)n);
    qsort(tmp, (size_t)n, sizeof(int64_t), cmp_i64);
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) if (i == 0 || tmp[i] != tmp[i - 1]) tmp[m++] = tmp[i];
    d->ov[r] = tmp; d->n_ov[r] = m; total += m;
  }
  /* owner of each id = min rank that has it */
  id_rank *all = (id_rank *)malloc(sizeof(id_rank) * (size_t)(total + 1));
  int64_t t = 0;
  for (int r = 0; r < P; ++r)
    for (int64_t i = 0; i < d->n_ov[r]; ++i) { all[t].id = d->ov[r][i]; all[t].rank = r; ++t; }
  qsort(all, (size_t)total, sizeof(id_rank), cmp_id_rank);
  int64_t nuniq = 0;
  int64_t *uid = (int64_t *)malloc(sizeof(int64_t) * (size_t)(total + 1));
  int *uowner = (int *)malloc(sizeof(int) * (size_t)(total + 1));
  for (int64_t i = 0; i < total; ++i)
    if (i == 0 || all[i].id != all[i - 1].id) { uid[nuniq] = all[i].id; uowner[nuniq] = all[i].rank; ++nuniq; }
  free(all);
  /* owned ids per rank in ascending id order; offsets by exclusive scan */
  int64_t *cnt = (int64_t *)calloc(P + 1, sizeof(int64_t));
  for (int64_t i = 0; i < nuniq; ++i) cnt[uowner[i]]++;
  int64_t *offset = (int64_t *)calloc(P + 1, sizeof(int64_t));
  for (int r = 1; r < P; ++r) offset[r] = offset[r - 1] + cnt[r - 1] * nf;
  /* GID of (id, field 0): offset[owner] + nf * position among owner's ids (ascending) */
  int64_t *ugid0 = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nuniq + 1));
  int64_t *pos = (int64_t *)calloc(P + 1, sizeof(int64_t));
  for (int64_t i = 0; i < nuniq; ++i) { int o = uowner[i]; ugid0[i] = offset[o] + nf * pos[o]; pos[o]++; }

  for (int r = 0; r < P; ++r) {
    const int64_t ne = d->ne[r], m = d->n_ov[r];
    d->ov_gid0[r] = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m + 1));
    char *is_owned = (char *)malloc((size_t)(m + 1));
    for (int64_t i = 0; i < m; ++i) {
      int64_t u = find_i64(uid, nuniq, d->ov[r][i]);
      d->ov_gid0[r][i] = ugid0[u];
      is_owned[i] = (uowner[u] == r);
    }
    const int gpe = ipe * nf;
    d->egids[r] = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ne * gpe + 1));
    d->elids[r] = (int *)malloc(sizeof(int) * (size_t)(ne * gpe + 1));
    int64_t *eov = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ne * ipe + 1));
    for (int64_t e = 0; e < ne; ++e)
      for (int c = 0; c < ipe; ++c) {
        int64_t i = find_i64(d->ov[r], m, d->conn[r][e * ipe + c]);
        eov[e * ipe + c] = i;
        for (int f = 0; f < nf; ++f) d->egids[r][e * gpe + c * nf + f] = d->ov_gid0[r][i] + f;
      }
    /* first-touch owned_, then first-touch ghosted_; touched[] is per (overlap id) since all
       nf fields of an id are consecutive in every element's GID list */
    int64_t *lid0 = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m + 1));
    for (int64_t i = 0; i < m; ++i) lid0[i] = -1;
    int64_t n_owned_ids = 0, n_ghost_ids = 0;
    for (int64_t i = 0; i < m; ++i) { if (is_owned[i]) n_owned_ids++; else n_ghost_ids++; }
    d->owned[r] = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_owned_ids * nf + 1));
    d->ghosted[r] = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_ghost_ids * nf + 1));
    int64_t no = 0, ng = 0;
    for (int64_t e = 0; e < ne; ++e)
      for (int c = 0; c < ipe; ++c) {
        int64_t i = eov[e * ipe + c];
        if (is_owned[i] && lid0[i] < 0) {
          lid0[i] = no;
          for (int f = 0; f < nf; ++f) d->owned[r][no++] = d->ov_gid0[r][i] + f;
        }
      }
    /* :635-636 leftovers (owned but not touched by my own elements) would be appended in
       unordered_set order -- implementation defined; cannot happen for ids taken from my own
       elements (every overlap id is touched by one of my elements).  Assert (assumption A3). */
    if (no != n_owned_ids * nf) return -2;
    for (int64_t e = 0; e < ne; ++e)
      for (int c = 0; c < ipe; ++c) {
        int64_t i = eov[e * ipe + c];
        if (!is_owned[i] && lid0[i] < 0) {
          lid0[i] = no + ng;
          for (int f = 0; f < nf; ++f) d->ghosted[r][ng++] = d->ov_gid0[r][i] + f;
        }
      }
    d->n_owned[r] = no; d->n_ghosted[r] = ng;
    for (int64_t e = 0; e < ne; ++e)
      for (int c = 0; c < ipe; ++c)
        for (int f = 0; f < nf; ++f)
          d->elids[r][e * gpe + c * nf + f] = (int)(lid0[eov[e * ipe + c]] + f);
    free(lid0); free(eov); free(is_owned);
  }
  free(uid); free(uowner); free(cnt); free(offset); free(ugid0); free(pos);
  d->built = 1;
  return 0;
}

int64_t orc_dofs_num_elems(const orc_dofs *d, int rank) { return d->ne[rank]; }
int64_t orc_dofs_num_owned(const orc_dofs *d, int rank) { return d->n_owned[rank]; }
int64_t orc_dofs_num_ghosted(const orc_dofs *d, int rank) { return d->n_ghosted[rank]; }
int orc_dofs_gids_per_elem(const orc_dofs *d) { return d->ipe * d->nfields; }
int orc_dofs_get_elem_gids(const orc_dofs *d, int rank, int64_t *out)
{ memcpy(out, d->egids[rank], sizeof(int64_t) * (size_t)(d->ne[rank] * d->ipe * d->nfields)); return 0; }
int orc_dofs_get_elem_lids(const orc_dofs *d, int rank, int *out)
{ memcpy(out, d->elids[rank], sizeof(int) * (size_t)(d->ne[rank] * d->ipe * d->nfields)); return 0; }
int orc_dofs_get_owned(const orc_dofs *d, int rank, int64_t *out)
{ memcpy(out, d->owned[rank], sizeof(int64_t) * (size_t)d->n_owned[rank]); return 0; }
int orc_dofs_get_ghosted(const orc_dofs *d, int rank, int64_t *out)
{ memcpy(out, d->ghosted[rank], sizeof(int64_t) * (size_t)d->n_ghosted[rank]); return 0; }

/* FieldAggPattern::buildFieldPatternData (Panzer_FieldAggPattern.cpp:201-276): subcell by
 * subcell, on each subcell the fields in field order -> offset of (field, basis b) for nodal
 * fields = b*nfields + field. */
int orc_dofs_field_offsets(const orc_dofs *d, int field, int *out)
{
  for (int b = 0; b < d->ipe; ++b) out[b] = b * d->nfields + field;
  return 0;
}

/* ======================================================================== */
/* D. Ghosted graph: disc-fe/src/lof/Panzer_TpetraLinearObjFactory_impl.hpp:558-650 */
/* ======================================================================== */
static int cmp_int(const void *a, const void *b)
{
  int x = *(const int *)a, y = *(const int *)b;
  return (x < y) ? -1 : (x > y);
}

/* Row map = column map = owned_++ghosted_ (:517-532), so local column index == LID.  Every
 * element inserts all of its GIDs into the row of each of its GIDs (:615-647); fillComplete
 * sorts each row by local column index and merges duplicates (Tpetra, assumption A4). */
int orc_ghosted_graph(int64_t ne, int npe, const int *lids, int n_rows, int64_t *rowptr, int *colind)
{
  int64_t *cnt = (int64_t *)calloc((size_t)n_rows + 1, sizeof(int64_t));
  for (int64_t e = 0; e < ne; ++e)
    for (int j = 0; j < npe; ++j) cnt[lids[e * npe + j]] += npe;       /* nEntriesPerRow (:598-601) */
  int64_t *start = (int64_t *)malloc(sizeof(int64_t) * ((size_t)n_rows + 1));
  start[0] = 0;
  for (int i = 0; i < n_rows; ++i) start[i + 1] = start[i] + cnt[i];
  int *raw = (int *)malloc(sizeof(int) * (size_t)(start[n_rows] + 1));
  memset(cnt, 0, sizeof(int64_t) * ((size_t)n_rows + 1));
  for (int64_t e = 0; e < ne; ++e)
    for (int j = 0; j < npe; ++j) {
      int row = lids[e * npe + j];
      for (int k = 0; k < npe; ++k) raw[start[row] + cnt[row]++] = lids[e * npe + k];   /* insertGlobalIndices (:646) */
    }
  rowptr[0] = 0;
  for (int i = 0; i < n_rows; ++i) {
    int *row = raw + start[i];
    int64_t n = cnt[i];
    qsort(row, (size_t)n, sizeof(int), cmp_int);
    int64_t m = 0;
    for (int64_t k = 0; k < n; ++k) if (k == 0 || row[k] != row[k - 1]) row[m++] = row[k];
    cnt[i] = m;
    rowptr[i + 1] = rowptr[i] + m;
  }
  if (colind)
    for (int i = 0; i < n_rows; ++i) memcpy(colind + rowptr[i], raw + start[i], sizeof(int) * (size_t)cnt[i]);
  free(cnt); free(start); free(raw);
  return 0;
}

/* ======================================================================== */
/* E. Geometry and basis tables                                              */
/* ======================================================================== */

/* Shards Hexahedron<8> reference vertices on [-1,1]^3; matches the node order the mesh factory
 * writes (Panzer_STK_CubeHexMeshFactory.cpp:447-455). */
static const double HEX_S[8][3] = {
  {-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};

/* Intrepid2 DefaultCubatureFactory::create(hex, degree 2|3) selected by
 * Panzer_IntegrationRule.cpp:153-166: tensor product of 2-point Gauss-Legendre rules, points
 * +-1/sqrt(3), weights 1, first direction fastest.  (The order of the points inside the rule is
 * third-party; it only changes the order of the sum over q.) */
void orc_ref_cubature(double *pts, double *wts)
{
  const double g = 1.0 / sqrt(3.0);
  const double gp[2] = {-g, g};
  for (int k = 0; k < 2; ++k)
    for (int j = 0; j < 2; ++j)
      for (int i = 0; i < 2; ++i) {
        int q = i + 2 * (j + 2 * k);
        pts[q * 3 + 0] = gp[i]; pts[q * 3 + 1] = gp[j]; pts[q * 3 + 2] = gp[k];
        wts[q] = 1.0;
      }
}

/* Intrepid2::Basis_HGRAD_HEX_C1_FEM::getValues(OPERATOR_VALUE / OPERATOR_GRAD), chosen by
 * Panzer_IntrepidBasisFactory.hpp:153-156; phi_i = (1 +- x)(1 +- y)(1 +- z)/8.  2-D analogue pinned
 * by disc-fe/test/core_tests/basis_values2.cpp:268-270. */
void orc_ref_basis(const double *pt, double *val, double *grad)
{
  const double x = pt[0], y = pt[1], z = pt[2];
  for (int i = 0; i < 8; ++i) {
    const double sx = HEX_S[i][0], sy = HEX_S[i][1], sz = HEX_S[i][2];
    if (val) val[i] = (1.0 + sx * x) * (1.0 + sy * y) * (1.0 + sz * z) / 8.0;
    if (grad) {
      grad[i * 3 + 0] = sx * (1.0 + sy * y) * (1.0 + sz * z) / 8.0;
      grad[i * 3 + 1] = (1.0 + sx * x) * sy * (1.0 + sz * z) / 8.0;
      grad[i * 3 + 2] = (1.0 + sx * x) * (1.0 + sy * y) * sz / 8.0;
    }
  }
}

/* IntegrationValues2<double>::evaluateValues (disc-fe/src/Panzer_IntegrationValues2.cpp):
 *   getJacobian :946-983 (CellTools::setJacobian: J(c,q,d,e) = sum_n X(c,n,d) dphi_n/dxi_e(q))
 *   getJacobianDeterminant :1021-1054, getJacobianInverse :985-1019 (3x3 cofactors)
 *   getWeightedMeasure :1056-1221 (computeCellMeasure: detJ * w_q)
 *   getCubaturePoints :1559-1625 (mapToPhysicalFrame: sum_n X(c,n,d) phi_n(q))
 * BasisValues2<double>::evaluateValues (disc-fe/src/Panzer_BasisValues2_impl.hpp):
 *   getBasisValues :1036-1190 (HGRADtransformVALUE = copy of reference values; multiplyMeasure)
 *   getGradBasisValues :1376-1521 (HGRADtransformGRAD: sum_e Jinv(c,q,e,d) dphi_b/dxi_e;
 *   multiplyMeasure: weighted_measure * grad_basis)
 * Identities pinned by disc-fe/test/core_tests/basis_values2.cpp:264-286. */
int orc_tables_build(int64_t ne, const double *X, orc_tables *t)
{
  double pts[24], wts[8], rv[8][8], rg[8][8][3];
  orc_ref_cubature(pts, wts);
  for (int q = 0; q < 8; ++q) {
    double v[8], g[24];
    orc_ref_basis(pts + 3 * q, v, g);
    for (int b = 0; b < 8; ++b) { rv[b][q] = v[b]; for (int d = 0; d < 3; ++d) rg[b][q][d] = g[b * 3 + d]; }
  }
  t->ne = ne;
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < ne; ++c) {
    const double *Xc = X + c * 24;
    for (int q = 0; q < 8; ++q) {
      double J[3][3], Ji[3][3];
      for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e) {
          double s = 0.0;
          for (int n = 0; n < 8; ++n) s += Xc[n * 3 + d] * rg[n][q][e];
          J[d][e] = s;
        }
      const double c0 = J[1][1] * J[2][2] - J[2][1] * J[1][2];
      const double c1 = -J[1][0] * J[2][2] + J[2][0] * J[1][2];
      const double c2 = J[1][0] * J[2][1] - J[2][0] * J[1][1];
      const double det = J[0][0] * c0 + J[0][1] * c1 + J[0][2] * c2;
      Ji[0][0] = c0 / det; Ji[1][0] = c1 / det; Ji[2][0] = c2 / det;
      Ji[0][1] = (-J[0][1] * J[2][2] + J[0][2] * J[2][1]) / det;
      Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
      Ji[2][1] = (-J[0][0] * J[2][1] + J[0][1] * J[2][0]) / det;
      Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
      Ji[1][2] = (-J[0][0] * J[1][2] + J[0][2] * J[1][0]) / det;
      Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
      const double wm = det * wts[q];
      for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e) {
          t->jac[((c * 8 + q) * 3 + d) * 3 + e] = J[d][e];
          t->jac_inv[((c * 8 + q) * 3 + d) * 3 + e] = Ji[d][e];
        }
      t->jac_det[c * 8 + q] = det;
      t->wm[c * 8 + q] = wm;
      for (int d = 0; d < 3; ++d) {
        double s = 0.0;
        for (int n = 0; n < 8; ++n) s += Xc[n * 3 + d] * rv[n][q];
        t->ip[(c * 8 + q) * 3 + d] = s;
      }
      for (int b = 0; b < 8; ++b) {
        t->basis[(c * 8 + b) * 8 + q] = rv[b][q];
        t->wbasis[(c * 8 + b) * 8 + q] = wm * rv[b][q];
        for (int d = 0; d < 3; ++d) {
          double s = 0.0;
          for (int e = 0; e < 3; ++e) s += Ji[e][d] * rg[b][q][e];
          t->gbasis[((c * 8 + b) * 8 + q) * 3 + d] = s;
          t->wgbasis[((c * 8 + b) * 8 + q) * 3 + d] = wm * s;
        }
      }
    }
  }
  return 0;
}

/* ======================================================================== */
/* F. The workset pipeline                                                   */
/* ======================================================================== */

#define NB 8   /* basis functions  */
#define NQ 8   /* quadrature points */
#define ND 3
#define NFAD 8 /* derivative length = DOFs per element (Panzer_FieldManagerBuilder.cpp:871-874) */

typedef struct { double val; double dx[NFAD]; } fad;   /* Sacado::Fad::DFad<double> stand-in */

static double source_value(int id, const double *ip)
{
  switch (id) {
    case 1: return 12.0 * M_PI * M_PI * sin(2.0 * M_PI * ip[0]) * sin(2.0 * M_PI * ip[1]) * sin(2.0 * M_PI * ip[2]);
    case 2: return 1.0;
    case 3: /* adapters-stk/example/PoissonExample/Example_SimpleSource_impl.hpp:88-96 */
      return 8.0 * M_PI * M_PI * sin(2.0 * M_PI * ip[0]) * sin(2.0 * M_PI * ip[1]);
    default: return 0.0;
  }
}

/* KokkosSparse::CrsMatrix::sumIntoValues(row, cols, n, vals, is_sorted=true, force_atomic=true)
 * as called at disc-fe/src/evaluators/Panzer_ScatterResidual_Tpetra_impl.hpp:410: for every
 * column, binary-search the sorted row; present -> atomic add; absent -> skipped. */
static void sum_into_values(const int64_t *rowptr, const int *colind, double *A, int row,
                            const int *cols, int n, const double *vals, int atomic)
{
  const int64_t b = rowptr[row];
  const int len = (int)(rowptr[row + 1] - b);
  for (int k = 0; k < n; ++k) {
    int lo = 0, hi = len - 1, at = -1;
    while (lo <= hi) {
      int mid = (lo + hi) / 2;
      int c = colind[b + mid];
      if (c == cols[k]) { at = mid; break; }
      if (c < cols[k]) lo = mid + 1; else hi = mid - 1;
    }
    if (at < 0) continue;
    if (atomic) {
#pragma omp atomic
      A[b + at] += vals[k];
    } else A[b + at] += vals[k];
  }
}

/* One workset of nc cells: the Phalanx DAG of the Poisson equation set
 * (adapters-stk/example/PoissonExample/Example_PoissonEquationSet_impl.hpp:150-195) in
 * topological order. */
static void evaluate_workset(const orc_terms *tm, int nc, int64_t c0, const int *lids, const orc_tables *t,
                             const double *x, const double *xdot, const double *xdotdot, const int64_t *rowptr, const int *colind,
                             double *f, double *A, int atomic)
{
  const int jac = tm->eval_type == 1;
  const int transient = (tm->mass_dot != 0.0) && xdot;
  /* second-order-in-time mass term: EXTENSION.  The reference registers D2XDT2_<dof>
     (disc-fe/src/Panzer_EquationSet_DefaultImpl_impl.hpp:962-981) and carries d2xdt2 in its containers
     (lof/Panzer_TpetraLinearObjContainer.hpp:116-117) but has no gather for it; here it is gathered from xdotdot
     with seed gamma, by analogy with the DXDT gather (seed alpha).  Parity unpinned. */
  const int second = (tm->mass_dotdot != 0.0) && xdotdot;
  fad *Tdd = second ? (fad *)calloc((size_t)nc * NB, sizeof(fad)) : NULL;        /* D2XDT2_TEMPERATURE <Cell,BASIS> */
  fad *Tddip = second ? (fad *)calloc((size_t)nc * NQ, sizeof(fad)) : NULL;
  /* per-workset MDFields */
  fad *T = (fad *)calloc((size_t)nc * NB, sizeof(fad));            /* TEMPERATURE        <Cell,BASIS> */
  fad *Tdot = (fad *)calloc((size_t)nc * NB, sizeof(fad));         /* DXDT_TEMPERATURE   <Cell,BASIS> */
  fad *gradT = (fad *)calloc((size_t)nc * NQ * ND, sizeof(fad));   /* GRAD_TEMPERATURE   <Cell,IP,Dim> */
  fad *Tip = (fad *)calloc((size_t)nc * NQ, sizeof(fad));          /* TEMPERATURE at IP  <Cell,IP> */
  fad *Tdip = (fad *)calloc((size_t)nc * NQ, sizeof(fad));         /* DXDT_TEMPERATURE at IP */
  double *src = (double *)calloc((size_t)nc * NQ, sizeof(double)); /* SOURCE_TEMPERATURE <Cell,IP> */
  fad *R = (fad *)calloc((size_t)nc * NB, sizeof(fad));            /* RESIDUAL_TEMPERATURE <Cell,BASIS> */

  /* K1: GlobalIndexer::getElementLIDs scratch copy (Panzer_GlobalIndexer.hpp:278-326) */
  int *slids = (int *)malloc(sizeof(int) * (size_t)nc * NB);
  for (int c = 0; c < nc; ++c) for (int i = 0; i < NB; ++i) slids[c * NB + i] = lids[(c0 + c) * NB + i];

  /* K2/K3: GatherSolution_Tpetra (Panzer_GatherSolution_Tpetra_impl.hpp:201-210 Residual,
     :554-572 seed choice, :615-633 Jacobian functor).  offsets(basis)=basis for one nodal field. */
  {
    double seed = tm->beta;                          /* gatherSeedIndex_<0 -> workset.beta */
    for (int c = 0; c < nc; ++c)
      for (int b = 0; b < NB; ++b) {
        const int offset = b, lid = slids[c * NB + offset];
        T[c * NB + b].val = x[lid];
        if (jac && seed != 0.0) T[c * NB + b].dx[offset] = seed;
      }
    if (transient) {
      seed = tm->alpha;                              /* useTimeDerivativeSolutionVector_ -> workset.alpha */
      for (int c = 0; c < nc; ++c)
        for (int b = 0; b < NB; ++b) {
          const int offset = b, lid = slids[c * NB + offset];
          Tdot[c * NB + b].val = xdot[lid];
          if (jac && seed != 0.0) Tdot[c * NB + b].dx[offset] = seed;
        }
    }
    if (second) {
      seed = tm->gamma;
      for (int c = 0; c < nc; ++c)
        for (int b = 0; b < NB; ++b) {
          const int offset = b, lid = slids[c * NB + offset];
          Tdd[c * NB + b].val = xdotdot[lid];
          if (jac && seed != 0.0) Tdd[c * NB + b].dx[offset] = seed;
        }
    }
  }
  if (second)                                     /* DOF evaluator on D2XDT2 */
    for (int c = 0; c < nc; ++c)
      for (int q = 0; q < NQ; ++q) {
        const double *bs = t->basis + (c0 + c) * NB * NQ;
        fad *o = &Tddip[c * NQ + q];
        o->val = Tdd[c * NB].val * bs[0 * NQ + q];
        for (int k = 0; k < NFAD; ++k) o->dx[k] = Tdd[c * NB].dx[k] * bs[0 * NQ + q];
        for (int bf = 1; bf < NB; ++bf) {
          o->val += Tdd[c * NB + bf].val * bs[bf * NQ + q];
          for (int k = 0; k < NFAD; ++k) o->dx[k] += Tdd[c * NB + bf].dx[k] * bs[bf * NQ + q];
        }
      }

  /* K5: DOFGradient::evaluateFields (Panzer_DOFGradient_impl.hpp:89-97): initialise with the
     b=0 product, then accumulate b=1..7 */
  for (int c = 0; c < nc; ++c)
    for (int q = 0; q < NQ; ++q)
      for (int d = 0; d < ND; ++d) {
        const double *gb = t->gbasis + ((c0 + c) * NB * NQ) * ND;
        fad *g = &gradT[(c * NQ + q) * ND + d];
        const double g0 = gb[(0 * NQ + q) * ND + d];
        g->val = T[c * NB].val * g0;
        for (int k = 0; k < NFAD; ++k) g->dx[k] = T[c * NB].dx[k] * g0;
        for (int bf = 1; bf < NB; ++bf) {
          const double gk = gb[(bf * NQ + q) * ND + d];
          g->val += T[c * NB + bf].val * gk;
          for (int k = 0; k < NFAD; ++k) g->dx[k] += T[c * NB + bf].dx[k] * gk;
        }
      }

  /* K4: DOF (EvaluateDOFWithSens_Scalar, Panzer_DOF_Functors.hpp:176-186) for fields used at IPs */
  if (tm->react != 0.0 || transient)
    for (int c = 0; c < nc; ++c)
      for (int q = 0; q < NQ; ++q) {
        const double *bs = t->basis + (c0 + c) * NB * NQ;
        for (int which = 0; which < 2; ++which) {
          if (which == 0 && tm->react == 0.0) continue;
          if (which == 1 && !transient) continue;
          const fad *u = which ? Tdot : T;
          fad *o = which ? &Tdip[c * NQ + q] : &Tip[c * NQ + q];
          o->val = u[c * NB].val * bs[0 * NQ + q];
          for (int k = 0; k < NFAD; ++k) o->dx[k] = u[c * NB].dx[k] * bs[0 * NQ + q];
          for (int bf = 1; bf < NB; ++bf) {
            o->val += u[c * NB + bf].val * bs[bf * NQ + q];
            for (int k = 0; k < NFAD; ++k) o->dx[k] += u[c * NB + bf].dx[k] * bs[bf * NQ + q];
          }
        }
      }

  /* K9: closure model, e.g. Example_SimpleSource_impl.hpp:88-96: source at ip_coordinates */
  if (tm->source_id && tm->source_mult != 0.0)
    for (int c = 0; c < nc; ++c)
      for (int q = 0; q < NQ; ++q) src[c * NQ + q] = source_value(tm->source_id, t->ip + ((c0 + c) * NQ + q) * 3);

  /* K6: Integrator_GradBasisDotVector, EVALUATES style, no field multipliers
     (Panzer_Integrator_GradBasisDotVector_impl.hpp:237-255): zero, then q outer, dim, basis */
  for (int c = 0; c < nc; ++c) {
    const double *wgb = t->wgbasis + ((c0 + c) * NB * NQ) * ND;
    for (int b = 0; b < NB; ++b) memset(&R[c * NB + b], 0, sizeof(fad));
    for (int q = 0; q < NQ; ++q)
      for (int d = 0; d < ND; ++d)
        for (int b = 0; b < NB; ++b) {
          double cf = wgb[(b * NQ + q) * ND + d] * tm->kappa;           /* basis_*multiplier_ ... */
          if (tm->fm_grad) cf *= tm->fm_grad[(c0 + c) * NQ + q];        /* ... * kokkosFieldMults_(fm)(cell, qp) (:276-281) */
          const fad *v = &gradT[(c * NQ + q) * ND + d];                /* ... *vector_ */
          R[c * NB + b].val += cf * v->val;
          for (int k = 0; k < NFAD; ++k) R[c * NB + b].dx[k] += cf * v->dx[k];
        }
  }

  /* K7: Integrator_BasisTimesScalar, CONTRIBUTES (Panzer_Integrator_BasisTimesScalar_impl.hpp:
     234-239): tmp = multiplier*scalar(c,q); field(c,b) += basis(c,b,q)*tmp.
     Registration order in the equation set: transient (:158-166), then source (:186-193);
     `react` is the analogous mass term on TEMPERATURE (dof-mgr/test/fe_assembly identity). */
  for (int c = 0; c < nc; ++c) {
    const double *wb = t->wbasis + (c0 + c) * NB * NQ;
    if (second)
      for (int q = 0; q < NQ; ++q) {
        const double fmm = tm->fm_mass ? tm->fm_mass[(c0 + c) * NQ + q] : 1.0;
        fad tmp; tmp.val = tm->mass_dotdot * fmm * Tddip[c * NQ + q].val;
        for (int k = 0; k < NFAD; ++k) tmp.dx[k] = tm->mass_dotdot * fmm * Tddip[c * NQ + q].dx[k];
        for (int b = 0; b < NB; ++b) {
          R[c * NB + b].val += wb[b * NQ + q] * tmp.val;
          for (int k = 0; k < NFAD; ++k) R[c * NB + b].dx[k] += wb[b * NQ + q] * tmp.dx[k];
        }
      }
    if (transient)
      for (int q = 0; q < NQ; ++q) {
        const double fmm = tm->fm_mass ? tm->fm_mass[(c0 + c) * NQ + q] : 1.0;
        fad tmp; tmp.val = tm->mass_dot * fmm * Tdip[c * NQ + q].val;
        for (int k = 0; k < NFAD; ++k) tmp.dx[k] = tm->mass_dot * fmm * Tdip[c * NQ + q].dx[k];
        for (int b = 0; b < NB; ++b) {
          R[c * NB + b].val += wb[b * NQ + q] * tmp.val;
          for (int k = 0; k < NFAD; ++k) R[c * NB + b].dx[k] += wb[b * NQ + q] * tmp.dx[k];
        }
      }
    if (tm->react != 0.0)
      for (int q = 0; q < NQ; ++q) {
        const double fmm = tm->fm_mass ? tm->fm_mass[(c0 + c) * NQ + q] : 1.0;
        fad tmp; tmp.val = tm->react * fmm * Tip[c * NQ + q].val;
        for (int k = 0; k < NFAD; ++k) tmp.dx[k] = tm->react * fmm * Tip[c * NQ + q].dx[k];
        for (int b = 0; b < NB; ++b) {
          R[c * NB + b].val += wb[b * NQ + q] * tmp.val;
          for (int k = 0; k < NFAD; ++k) R[c * NB + b].dx[k] += wb[b * NQ + q] * tmp.dx[k];
        }
      }
    if (tm->source_id && tm->source_mult != 0.0)
      for (int q = 0; q < NQ; ++q) {
        const double tmp = tm->source_mult * src[c * NQ + q];
        for (int b = 0; b < NB; ++b) R[c * NB + b].val += wb[b * NQ + q] * tmp;
      }
  }

  /* K11/K12: ScatterResidual_Tpetra (Panzer_ScatterResidual_Tpetra_impl.hpp:374-413 Jacobian
     functor, :415-440 Residual functor) */
  for (int c = 0; c < nc; ++c)
    for (int b = 0; b < NB; ++b) {
      const int offset = b, lid = slids[c * NB + offset];
      if (f) {
        if (atomic) {
#pragma omp atomic
          f[lid] += R[c * NB + b].val;
        } else f[lid] += R[c * NB + b].val;
      }
      if (jac && A) {
        double vals[NFAD];
        for (int s = 0; s < NFAD; ++s) vals[s] = R[c * NB + b].dx[s];   /* scratch_vals_ */
        sum_into_values(rowptr, colind, A, lid, &slids[c * NB], NFAD, vals, atomic);
      }
    }

  free(T); free(Tdot); free(gradT); free(Tip); free(Tdip); free(src); free(R); free(slids); free(Tdd); free(Tddip);
}

/* AssemblyEngine<EvalT>::evaluateVolume (disc-fe/src/Panzer_AssemblyEngine_impl.hpp:134-182):
 * sequential loop over worksets of <= workset_size cells in element order
 * (Panzer_Workset_Builder_impl.hpp:128-151).  nthreads>1 spreads worksets over OpenMP threads
 * for the CPU baseline (scatter then uses atomics, as Kokkos::atomic_add does). */
int orc_evaluate_volume(const orc_terms *tm, int64_t ne, const int *lids, const orc_tables *t,
                        const double *x, const double *xdot, int n_rows, const int64_t *rowptr, const int *colind,
                        double *f, double *A)
{
  return orc_evaluate_volume2(tm, ne, lids, t, x, xdot, NULL, n_rows, rowptr, colind, f, A);
}

int orc_evaluate_volume2(const orc_terms *tm, int64_t ne, const int *lids, const orc_tables *t,
                         const double *x, const double *xdot, const double *xdotdot, int n_rows, const int64_t *rowptr,
                         const int *colind, double *f, double *A)
{
  (void)n_rows;
  const int W = tm->workset_size > 0 ? tm->workset_size : 20;
  const int64_t nws = (ne + W - 1) / W;
  int nt = tm->nthreads > 0 ? tm->nthreads : 1;
#ifndef _OPENMP
  nt = 1;
#endif
  if (nt == 1) {
    for (int64_t w = 0; w < nws; ++w) {
      int64_t c0 = w * W;
      int nc = (int)((ne - c0) < W ? (ne - c0) : W);
      evaluate_workset(tm, nc, c0, lids, t, x, xdot, xdotdot, rowptr, colind, f, A, 0);
    }
  } else {
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
    for (int64_t w = 0; w < nws; ++w) {
      int64_t c0 = w * W;
      int nc = (int)((ne - c0) < W ? (ne - c0) : W);
      evaluate_workset(tm, nc, c0, lids, t, x, xdot, xdotdot, rowptr, colind, f, A, 1);
    }
  }
  return 0;
}

/* ======================================================================== */
/* G. Dirichlet                                                              */
/* ======================================================================== */
/* TianXin::DirichletEvalautor::evaluateFields (disc-fe/src/evaluators/TianXin_Dirichlet_impl.hpp:
 * 59-81): Residual -> evalDirichletResidual: f[l] = x[l] - value
 * (lof/Panzer_TpetraLinearObjContainer.hpp:306-317); Jacobian ->
 * Tpetra::applyDirichletBoundaryConditionToLocalMatrixRows (row l: diagonal 1, other stored
 * entries 0; columns untouched) followed by the same residual (:228-237). */
int orc_dirichlet(int eval_type, int n, const int *local_dofs, const double *values,
                  const double *x, double *f, const int64_t *rowptr, const int *colind, double *A)
{
  for (int i = 0; i < n; ++i) {
    const int l = local_dofs[i];
    if (eval_type == 1 && A)
      for (int64_t k = rowptr[l]; k < rowptr[l + 1]; ++k) A[k] = (colind[k] == l) ? 1.0 : 0.0;
    if (f) f[l] = x[l] - values[i];
  }
  return 0;
}

/* Jacobian evaluation with f == null (the eigenvalue path): TpetraLinearObjContainer::applyDirichletBoundaryCondition
 * calls Tpetra::applyDirichletBoundaryConditionToLocalMatrixRowsAndColumns (lof/Panzer_TpetraLinearObjContainer.hpp:
 * 223-226): rows := identity AND every stored entry in the columns of the listed DOFs := 0.  Naive scan of all rows. */
int orc_dirichlet_rows_and_columns(int n, const int *local_dofs, int64_t n_rows, const int64_t *rowptr, const int *colind, double *A)
{
  char *is_dir = (char *)calloc((size_t)n_rows, 1);
  if (!is_dir) return -1;
  for (int i = 0; i < n; ++i) is_dir[local_dofs[i]] = 1;
  for (int64_t r = 0; r < n_rows; ++r)
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      const int c = colind[k];
      if (is_dir[r]) A[k] = (c == r) ? 1.0 : 0.0;
      else if (is_dir[c]) A[k] = 0.0;
    }
  free(is_dir);
  return 0;
}

/* TianXin::CLoadEvalautor<Residual>::evaluateFields (disc-fe/src/evaluators/TianXin_CLoad_impl.hpp:56-61) ->
 * TpetraLinearObjContainer::applyConcentratedLoad (lof/Panzer_TpetraLinearObjContainer.hpp:343-350):
 * fview(local_dofs(i)) += values(i).  The Jacobian evaluator does nothing (:73-78). */
int orc_cload(int eval_type, int n, const int *local_dofs, const double *values, double *f)
{
  if (eval_type != 0 || !f) return 0;
  for (int i = 0; i < n; ++i) f[local_dofs[i]] += values[i];
  return 0;
}

/* TianXin::Flux<EvalT>::evaluateFields (disc-fe/src/evaluators/TianXin_Neumann_impl.hpp:143-160) on a SIDE workset:
 *   residual(cell,b) = val * sum_qp weighted_basis_scalar(cell,b,qp),  scattered (added) into f by ScatterResidual.
 * Side worksets integrate on the face: cubature of the side topology (Quadrilateral<4>, degree 2 -> 2x2 Gauss, weights
 * 1) mapped into the cell by CellTools::mapToReferenceSubcell; weighted measure by FunctionSpaceTools::computeFaceMeasure
 * = w_q * || J t1 x J t2 ||  with the reference face tangents t1, t2 (IntegrationValues2::getWeightedMeasure, side
 * branch); weighted_basis_scalar = N_b(xi_q) * weighted_measure.  Shards Hexahedron<8> side -> vertex map below.
 * The value is constant in x, so the Jacobian-type evaluation adds the same numbers to f and nothing to A. */
static const int HEX_SIDE_NODES[6][4] = {{0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {0, 4, 7, 3}, {0, 3, 2, 1}, {4, 5, 6, 7}};
int orc_neumann_flux(int64_t n_sides, const int *cells, const int *sides, const double *values, const int *lids,
                     const double *cell_coords, double *f)
{
  const double g = 0.57735026918962576451;
  const double qs[4][2] = {{-g, -g}, {g, -g}, {g, g}, {-g, g}};
  for (int64_t i = 0; i < n_sides; ++i) {
    const int c = cells[i], sd = sides[i];
    if (sd < 0 || sd > 5) return -1;
    const double *Xc = cell_coords + (int64_t)c * 24;
    const int *sn = HEX_SIDE_NODES[sd];
    double V[4][3], t1[3], t2[3], r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 4; ++k) for (int d = 0; d < 3; ++d) V[k][d] = HEX_S[sn[k]][d];
    for (int d = 0; d < 3; ++d) { t1[d] = 0.5 * (V[1][d] - V[0][d]); t2[d] = 0.5 * (V[3][d] - V[0][d]); }
    for (int q = 0; q < 4; ++q) {
      const double s = qs[q][0], t = qs[q][1];
      double pt[3], val[8], grad[24], J[3][3], T1[3], T2[3];
      for (int d = 0; d < 3; ++d)
        pt[d] = 0.25 * ((1 - s) * (1 - t) * V[0][d] + (1 + s) * (1 - t) * V[1][d] + (1 + s) * (1 + t) * V[2][d] + (1 - s) * (1 + t) * V[3][d]);
      orc_ref_basis(pt, val, grad);
      for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e) {
          double a = 0.0;
          for (int n = 0; n < 8; ++n) a += Xc[n * 3 + d] * grad[n * 3 + e];
          J[d][e] = a;
        }
      for (int d = 0; d < 3; ++d) {
        T1[d] = J[d][0] * t1[0] + J[d][1] * t1[1] + J[d][2] * t1[2];
        T2[d] = J[d][0] * t2[0] + J[d][1] * t2[1] + J[d][2] * t2[2];
      }
      const double nx = T1[1] * T2[2] - T1[2] * T2[1], ny = T1[2] * T2[0] - T1[0] * T2[2], nz = T1[0] * T2[1] - T1[1] * T2[0];
      const double wm = sqrt(nx * nx + ny * ny + nz * nz);
      for (int b = 0; b < 8; ++b) r[b] += values[i] * (val[b] * wm);
    }
    for (int b = 0; b < 8; ++b) f[lids[(int64_t)c * 8 + b]] += r[b];
  }
  return 0;
}

/* ======================================================================== */
/* G2. Functional responses                           */
/* ======================================================================== */

/* Gauss-Legendre rule with n points on [-1,1] (Intrepid2 CubatureDirectLineGauss tabulates these; Newton on P_n
 * from the Chebyshev guess reproduces them to the last bits) */
int orc_gauss_legendre(int n, double *x, double *w)
{
  if (n < 1 || n > 16) return -1;
  for (int i = 0; i < n; ++i) {
    double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 1; j <= n; ++j) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j; }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      const double dz = p1 / pp;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    x[n - 1 - i] = z;
    w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
  return 0;
}

/* Response_Functional: value = sum over cells of Integrator_Scalar's cell integral
 *   integral(cell) = sum_qp multiplier * scalar(cell,qp) * weighted_measure(cell,qp)
 *   (disc-fe/src/evaluators/Panzer_Integrator_Scalar_impl.hpp:117-142;
 *    disc-fe/src/responses/Panzer_ResponseScatterEvaluator_Functional_impl.hpp:137-142; global sum by
 *    Response_Functional::scatterResponse, Panzer_Response_Functional_impl.hpp:130-145 -- done by the caller here).
 * The integrands are the example's closure models (adapters-stk/example/PoissonExample/
 * Example_ClosureModel_Factory_impl.hpp:147-275): "L2 ERROR_CALC" = (A - B)^2, "H1 ERROR_CALC" = (A - B)^2 +
 * |grad A - grad B|^2 with A = the DOF field at the integration points (DOF / DOFGradient evaluators) and B the
 * exact solution (Example_SimpleSolution_impl.hpp:85-105: sin 2pi x sin 2pi y; solution_id 1 is the 3-D product
 * that matches source 1).  kind 1: integrand A.  Cubature: tensor Gauss of the given degree (points = degree/2+1),
 * the example asks for degree 10. */
int orc_response_functional(int kind, int solution_id, int cub_degree, int64_t ne, const int *lids,
                            const double *cell_coords, const double *x, double *value)
{
  const int np = cub_degree / 2 + 1;
  double gx[16], gw[16];
  if (orc_gauss_legendre(np, gx, gw)) return -1;
  double total = 0.0;
  for (int64_t c = 0; c < ne; ++c) {
    const double *Xc = cell_coords + c * 24;
    double u[8];
    for (int a = 0; a < 8; ++a) u[a] = x[lids[c * 8 + a]];
    double integral = 0.0;
    for (int k = 0; k < np; ++k)
      for (int j = 0; j < np; ++j)
        for (int i = 0; i < np; ++i) {
          const double pt[3] = {gx[i], gx[j], gx[k]};
          double val[8], grad[24], J[3][3], P[3] = {0, 0, 0}, A = 0.0, gA_ref[3] = {0, 0, 0};
          orc_ref_basis(pt, val, grad);
          for (int d = 0; d < 3; ++d)
            for (int e = 0; e < 3; ++e) {
              double a = 0.0;
              for (int n = 0; n < 8; ++n) a += Xc[n * 3 + d] * grad[n * 3 + e];
              J[d][e] = a;
            }
          for (int n = 0; n < 8; ++n) {
            A += val[n] * u[n];
            for (int d = 0; d < 3; ++d) { P[d] += val[n] * Xc[n * 3 + d]; gA_ref[d] += grad[n * 3 + d] * u[n]; }
          }
          const double c0 = J[1][1] * J[2][2] - J[2][1] * J[1][2];
          const double c1 = -J[1][0] * J[2][2] + J[2][0] * J[1][2];
          const double c2 = J[1][0] * J[2][1] - J[2][0] * J[1][1];
          const double det = J[0][0] * c0 + J[0][1] * c1 + J[0][2] * c2;
          const double wm = det * gw[i] * gw[j] * gw[k];
          double sc;
          if (kind == 1) sc = A;
          else {
            const double sx = sin(2 * M_PI * P[0]), sy = sin(2 * M_PI * P[1]), cx = cos(2 * M_PI * P[0]), cy = cos(2 * M_PI * P[1]);
            const double sz = (solution_id == 1) ? sin(2 * M_PI * P[2]) : 1.0, cz = (solution_id == 1) ? cos(2 * M_PI * P[2]) : 0.0;
            const double B = sx * sy * sz;
            sc = (A - B) * (A - B);
            if (kind == 3) {
              double Ji[3][3];
              Ji[0][0] = c0 / det; Ji[1][0] = c1 / det; Ji[2][0] = c2 / det;
              Ji[0][1] = (-J[0][1] * J[2][2] + J[0][2] * J[2][1]) / det;
              Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
              Ji[2][1] = (-J[0][0] * J[2][1] + J[0][1] * J[2][0]) / det;
              Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
              Ji[1][2] = (-J[0][0] * J[1][2] + J[0][2] * J[1][0]) / det;
              Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
              const double gB[3] = {2 * M_PI * cx * sy * sz, 2 * M_PI * sx * cy * sz, 2 * M_PI * sx * sy * cz};
              for (int d = 0; d < 3; ++d) {
                const double gA = Ji[0][d] * gA_ref[0] + Ji[1][d] * gA_ref[1] + Ji[2][d] * gA_ref[2];
                sc += (gA - gB[d]) * (gA - gB[d]);
              }
            }
          }
          integral += sc * wm;
        }
    total += integral;
  }
  *value = total;
  return 0;
}

/* TianXin::Response_Integral<Residual>::evaluateFields (disc-fe/src/responses/TianXin_Response_Integral_impl.hpp:106-133):
 *   result = sum_cell sum_qp cellvalue_(cell, qp) * wm(cell, qp)   (Kokkos::parallel_reduce per workset)
 *   reduceAll(REDUCE_SUM) -> glbValue (the caller sums the ranks here); value_ = glbValue;
 *   tVector_->sumIntoLocalValue(0, glbValue).  wm = weighted_measure of the workset's integration rule. */
int orc_response_integral(int64_t ne, int nq, const double *cellvalue, const double *wm, double *response_vector, double *value)
{
  double result = 0.0;
  for (int64_t c = 0; c < ne; ++c) {
    double cell_integral = 0.0;
    for (int q = 0; q < nq; ++q) cell_integral += cellvalue[c * nq + q] * wm[c * nq + q];
    result += cell_integral;
  }
  if (!response_vector) return -1;       /* "reponse vector not defined" */
  response_vector[0] += result;
  if (value) *value = result;
  return 0;
}

/* ======================================================================== */
/* H. Import / Export between owned and ghosted vectors                      */
/* ======================================================================== */
/* TpetraLinearObjFactory::globalToGhostTpetraVector (lof/..._impl.hpp:207-219):
 * out.putScalar(0); out.doImport(in, importer, INSERT)  -- ghosted map = owned_++ghosted_ */
int orc_global_to_ghost(const orc_dofs *d, const double *const *x_owned, int rank, double *xg)
{
  const int64_t no = d->n_owned[rank], ng = d->n_ghosted[rank];
  for (int64_t i = 0; i < no; ++i) xg[i] = x_owned[rank][i];
  for (int64_t i = 0; i < ng; ++i) {
    const int64_t gid = d->ghosted[rank][i];
    int found = 0;
    for (int r = 0; r < d->nranks && !found; ++r)
      for (int64_t k = 0; k < d->n_owned[r]; ++k)
        if (d->owned[r][k] == gid) { xg[no + i] = x_owned[r][k]; found = 1; break; }
    if (!found) return -1;
  }
  return 0;
}

/* ghostToGlobalTpetraVector (:170-180): out.putScalar(0); out.doExport(in, exporter, ADD) */
int orc_ghost_to_global_vec(const orc_dofs *d, const double *const *fg, int rank, double *fo)
{
  const int64_t no = d->n_owned[rank];
  for (int64_t i = 0; i < no; ++i) fo[i] = 0.0;
  for (int64_t i = 0; i < no; ++i) {
    const int64_t gid = d->owned[rank][i];
    for (int r = 0; r < d->nranks; ++r) {          /* contributions in rank order */
      const int64_t nl = d->n_owned[r] + d->n_ghosted[r];
      for (int64_t k = 0; k < nl; ++k) {
        const int64_t g = (k < d->n_owned[r]) ? d->owned[r][k] : d->ghosted[r][k - d->n_owned[r]];
        if (g == gid) { fo[i] += fg[r][k]; break; }
      }
    }
  }
  return 0;
}
