// txhost.cpp -- host-side mirror of the reference classes that FEED the assembly hot path
// (see include/txhost.h).  C++17, host only, OpenMP for the large sorts.  Independent of oracle/.
#include "../../../include/txhost.h"
#include <algorithm>
#include <omp.h>
#include <climits>
#include <cstdint>
#include <cmath>
#include <cfloat>
#include <cstring>
#include <string>
#include <vector>
#include <unordered_map>
#include <parallel/algorithm>

namespace {
thread_local std::string g_err;
int fail(const std::string &m) { g_err = m; return -1; }
}  // namespace

extern "C" const char *txhost_last_error(void) { return g_err.c_str(); }
// launchers such as torchrun export OMP_NUM_THREADS=1 to every rank; the caller knows how many cores the rank really has
extern "C" int txhost_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }

// =====================================================================================
// CubeHexMeshFactory  (adapters-stk/src/stk_interface/Panzer_STK_CubeHexMeshFactory.cpp)
// =====================================================================================
struct txhost_mesh_s {
  int nx, ny, nz, px, py, pz, rank, nranks;
  double x0, xf, y0, yf, z0, zf;
  int64_t xs, xn, ys, yn, zs, zn;      // my brick of elements: start, count per axis
  std::vector<int64_t> elem_ids, elem_nodes;
  std::vector<double> cell_coords;
};

namespace {

// :89-133 default processor grid
void default_grid(int size, int &px, int &py, int &pz)
{
  px = py = pz = (int)std::pow((double)size, 0.333334);
  if (px * py * pz != size) {
    px = py = pz = 1;
    const int maxFactor = 50;
    int ProcTemp = size;
    int factors[maxFactor];
    for (int jj = 0; jj < maxFactor; jj++) factors[jj] = 0;
    for (int jj = 2; jj < maxFactor; jj++) {
      bool flag = true;
      while (flag) {
        int temp = ProcTemp / jj;
        if (temp * jj == ProcTemp) { factors[jj]++; ProcTemp = temp; }
        else flag = false;
      }
    }
    px = ProcTemp;
    for (int jj = maxFactor - 1; jj > 0; jj--)
      while (factors[jj] != 0) {
        if ((px <= py) && (px <= pz)) px = px * jj;
        else if ((py <= px) && (py <= pz)) py = py * jj;
        else pz = pz * jj;
        factors[jj]--;
      }
  }
}

// :463-535 determine*ElemSizeAndStart
void size_and_start(int n, int size, int loc, int64_t &start, int64_t &nume)
{
  const int64_t minE = n / size, extra = n - minE * size;
  if (loc < extra) { nume = minE + 1; start = loc * (minE + 1); }
  else { nume = minE; start = extra * (minE + 1) + (loc - extra) * minE; }
}

// Panzer_STK_MeshFactory.hpp:161-168
double mesh_coord(int64_t nx, double delta, double x0)
{
  const double x = (double)nx * delta;
  double val = x + x0;
  if ((x0 * x < 0.0) && (std::fabs(std::fabs(x) - std::fabs(x0)) < DBL_EPSILON * std::fabs(x0))) val = 0.0;
  return val;
}

uint64_t splitmix64(uint64_t z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace

extern "C" {

txhost_mesh txhost_cube_hex_mesh(int nx, int ny, int nz, int px, int py, int pz, double x0, double xf, double y0,
                                 double yf, double z0, double zf, int rank, int nranks)
{
  if (nx < 1 || ny < 1 || nz < 1 || nranks < 1 || rank < 0 || rank >= nranks) { fail("cube_hex_mesh: bad arguments"); return nullptr; }
  if (px == -1 && py == -1 && pz == -1) default_grid(nranks, px, py, pz);   // :89-127
  else if (px == -1) { px = nranks; py = 1; pz = 1; }                        // :128-133
  if (px * py * pz != nranks) { fail("the product of X/Y/Z Procs must equal the number of processors"); return nullptr; }  // :134-136
  if (nx / px < 1 || ny / py < 1 || nz / pz < 1) { fail("fewer elements than processors along an axis"); return nullptr; }
  auto *m = new txhost_mesh_s();
  m->nx = nx; m->ny = ny; m->nz = nz; m->px = px; m->py = py; m->pz = pz; m->rank = rank; m->nranks = nranks;
  m->x0 = x0; m->xf = xf; m->y0 = y0; m->yf = yf; m->z0 = z0; m->zf = zf;
  // :918-927 procRankToProcTuple
  int r = rank;
  const int k = r / (px * py); r = r % (px * py);
  const int j = r / px; r = r % px;
  const int i = r;
  size_and_start(nx, px, i, m->xs, m->xn);
  size_and_start(ny, py, j, m->ys, m->yn);
  size_and_start(nz, pz, k, m->zs, m->zn);
  const int64_t ne = m->xn * m->yn * m->zn, NX = nx, NY = ny;
  m->elem_ids.resize(ne); m->elem_nodes.resize(ne * 8); m->cell_coords.resize(ne * 24);
  const double dX = (xf - x0) / (double)nx, dY = (yf - y0) / (double)ny, dZ = (zf - z0) / (double)nz;
  // :401-461 buildBlock; local element order = ascending element id (x fastest)
#pragma omp parallel for schedule(static)
  for (int64_t lz = 0; lz < m->zn; ++lz)
    for (int64_t ly = 0; ly < m->yn; ++ly)
      for (int64_t lx = 0; lx < m->xn; ++lx) {
        const int64_t e = lx + m->xn * (ly + m->yn * lz);
        const int64_t ex = m->xs + lx, ey = m->ys + ly, ez = m->zs + lz;
        int64_t n[8];
        n[0] = ex + 1 + ey * (NX + 1) + ez * (NY + 1) * (NX + 1);
        n[1] = n[0] + 1;
        n[2] = n[1] + (NX + 1);
        n[3] = n[2] - 1;
        for (int a = 0; a < 4; ++a) n[4 + a] = n[a] + (NY + 1) * (NX + 1);
        m->elem_ids[e] = NX * NY * ez + NX * ey + ex + 1;
        for (int a = 0; a < 8; ++a) {
          m->elem_nodes[e * 8 + a] = n[a];
          const int64_t id0 = n[a] - 1;
          const int64_t ix = id0 % (NX + 1), iy = (id0 / (NX + 1)) % (NY + 1), iz = id0 / ((NX + 1) * (NY + 1));
          m->cell_coords[(e * 8 + a) * 3 + 0] = mesh_coord(ix, dX, x0);
          m->cell_coords[(e * 8 + a) * 3 + 1] = mesh_coord(iy, dY, y0);
          m->cell_coords[(e * 8 + a) * 3 + 2] = mesh_coord(iz, dZ, z0);
        }
      }
  return m;
}

// The rank's brick of elements and the processor grid, without building any array (the device mesh builder forms ids,
// connectivity and coordinates from these on the GPU): out = {xs, xn, ys, yn, zs, zn, px, py, pz}
int txhost_cube_hex_brick(int nx, int ny, int nz, int px, int py, int pz, int rank, int nranks, int64_t *out)
{
  if (nx < 1 || ny < 1 || nz < 1 || nranks < 1 || rank < 0 || rank >= nranks || !out) return fail("cube_hex_brick: bad arguments");
  if (px == -1 && py == -1 && pz == -1) default_grid(nranks, px, py, pz);
  else if (px == -1) { px = nranks; py = 1; pz = 1; }
  if (px * py * pz != nranks) return fail("the product of X/Y/Z Procs must equal the number of processors");
  if (nx / px < 1 || ny / py < 1 || nz / pz < 1) return fail("fewer elements than processors along an axis");
  int r = rank;
  const int k = r / (px * py); r = r % (px * py);
  const int j = r / px; r = r % px;
  const int i = r;
  size_and_start(nx, px, i, out[0], out[1]);
  size_and_start(ny, py, j, out[2], out[3]);
  size_and_start(nz, pz, k, out[4], out[5]);
  out[6] = px; out[7] = py; out[8] = pz;
  return 0;
}

void txhost_mesh_destroy(txhost_mesh m) { delete m; }
int64_t txhost_mesh_num_elems(txhost_mesh m) { return m ? (int64_t)m->elem_ids.size() : -1; }
int txhost_mesh_proc_grid(txhost_mesh m, int *px, int *py, int *pz) { *px = m->px; *py = m->py; *pz = m->pz; return 0; }

int txhost_mesh_get(txhost_mesh m, int64_t *elem_ids, int64_t *elem_nodes, double *cell_coords)
{
  if (elem_ids) memcpy(elem_ids, m->elem_ids.data(), m->elem_ids.size() * sizeof(int64_t));
  if (elem_nodes) memcpy(elem_nodes, m->elem_nodes.data(), m->elem_nodes.size() * sizeof(int64_t));
  if (cell_coords) memcpy(cell_coords, m->cell_coords.data(), m->cell_coords.size() * sizeof(double));
  return 0;
}

// STKConnManager::addSubcellConnectivities (adapters-stk/src/Panzer_STKConnManager.cpp:201-226):
// offset + idCnt*(stk_id-1) + i with nodeOffset = 0, idCnt = 1
int txhost_mesh_connectivity(txhost_mesh m, int64_t *conn)
{
  const int64_t n = (int64_t)m->elem_nodes.size();
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) conn[i] = m->elem_nodes[i] - 1;
  return 0;
}

int txhost_mesh_perturb(txhost_mesh m, double amp)
{
  const int64_t NX = m->nx, NY = m->ny, NZ = m->nz;
  const double h[3] = {(m->xf - m->x0) / NX, (m->yf - m->y0) / NY, (m->zf - m->z0) / NZ};
  const int64_t n = (int64_t)m->elem_nodes.size();
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const int64_t id0 = m->elem_nodes[i] - 1;
    const int64_t ix = id0 % (NX + 1), iy = (id0 / (NX + 1)) % (NY + 1), iz = id0 / ((NX + 1) * (NY + 1));
    if (ix == 0 || ix == NX || iy == 0 || iy == NY || iz == 0 || iz == NZ) continue;
    for (int c = 0; c < 3; ++c) {
      const double u = (double)splitmix64(0x5EEDull + (uint64_t)id0 * 3ull + (uint64_t)c) / 18446744073709551616.0;
      m->cell_coords[i * 3 + c] += amp * h[c] * (u - 0.5);
    }
  }
  return 0;
}

// CubeHexMeshFactory::addSideSets (:604-750): nx==0 left, nx==max right, ny==0 bottom, ny==max top,
// nz==0 back, nz==max front.  Returns the distinct nodes of my elements on that side, ascending.
int64_t txhost_mesh_sideset_nodes(txhost_mesh m, const char *name, int64_t *out)
{
  const std::string s(name);
  int axis, hi;
  if (s == "left") { axis = 0; hi = 0; } else if (s == "right") { axis = 0; hi = 1; }
  else if (s == "bottom") { axis = 1; hi = 0; } else if (s == "top") { axis = 1; hi = 1; }
  else if (s == "back") { axis = 2; hi = 0; } else if (s == "front") { axis = 2; hi = 1; }
  else { fail("unknown sideset " + s); return -1; }
  const int64_t NX = m->nx, NY = m->ny, NZ = m->nz;
  const int64_t N[3] = {NX, NY, NZ}, s0[3] = {m->xs, m->ys, m->zs}, cn[3] = {m->xn, m->yn, m->zn};
  const int64_t plane = hi ? N[axis] : 0;                       // node index of the side
  if (!(s0[axis] <= plane && plane <= s0[axis] + cn[axis])) return 0;
  if (hi ? (s0[axis] + cn[axis] != N[axis]) : (s0[axis] != 0)) return 0;
  int64_t lo[3] = {m->xs, m->ys, m->zs}, up[3] = {m->xs + m->xn, m->ys + m->yn, m->zs + m->zn};
  lo[axis] = up[axis] = plane;
  int64_t cnt = 0;
  for (int64_t z = lo[2]; z <= up[2]; ++z)
    for (int64_t y = lo[1]; y <= up[1]; ++y)
      for (int64_t x = lo[0]; x <= up[0]; ++x) {
        if (out) out[cnt] = z * (NY + 1) * (NX + 1) + y * (NX + 1) + x + 1;
        ++cnt;
      }
  return cnt;
}

}  // extern "C"

// =====================================================================================
// DOFManager::buildGlobalUnknowns  (dof-mgr/src/Panzer_DOFManager.cpp:474-714, GUN :719-865)
// =====================================================================================
struct txhost_dofmgr_s {
  int rank, P, ipe, nf;
  int64_t ne = 0;
  std::vector<int64_t> conn;
  int state = 0;
  // overlap map (std::set order = ascending ids, :1261-1290)
  std::vector<int64_t> ov;            // sorted unique ids of my elements
  std::vector<int> ov_owner;          // owning rank
  std::vector<int64_t> ov_gid0;       // first GID of the id (field 0)
  std::vector<int64_t> eov;           // [ne*ipe] index into ov
  // directory side (ids hashed to me by id % P)
  std::vector<int64_t> dir_ids; std::vector<int> dir_from;   // received requests in arrival order
  // results
  std::vector<int64_t> owned, ghosted, egids;
  std::vector<int> ghosted_owner, elids;
  int64_t my_offset = 0;
  // exchange buffers
  std::vector<int64_t> scounts, sbuf;
  std::vector<std::vector<int64_t>> req_ids;    // ids I asked each rank about (to pair replies)
};

namespace {
using DM = txhost_dofmgr_s;

void pack(DM *d, const std::vector<std::vector<int64_t>> &per)
{
  d->scounts.assign(d->P, 0);
  d->sbuf.clear();
  for (int r = 0; r < d->P; ++r) { d->scounts[r] = (int64_t)per[r].size(); d->sbuf.insert(d->sbuf.end(), per[r].begin(), per[r].end()); }
}
}  // namespace

extern "C" {

txhost_dofmgr txhost_dofmgr_create(int rank, int nranks, int ids_per_elem, int nfields)
{
  if (nranks < 1 || rank < 0 || rank >= nranks || ids_per_elem < 1 || nfields < 1) { fail("dofmgr_create: bad arguments"); return nullptr; }
  auto *d = new DM();
  d->rank = rank; d->P = nranks; d->ipe = ids_per_elem; d->nf = nfields;
  return d;
}
// A finished DOF manager from its results (built elsewhere, e.g. on the device): what TpetraLinearObjFactory needs for the
// plan negotiation in compact mode -- owned / ghosted GIDs in LID order, the owner of every ghost, this rank's first GID.
txhost_dofmgr txhost_dofmgr_from_arrays(int rank, int nranks, int ids_per_elem, int nfields, int64_t n_owned, const int64_t *owned,
                                        int64_t n_ghosted, const int64_t *ghosted, const int *ghosted_owner, int64_t my_offset)
{
  txhost_dofmgr d = txhost_dofmgr_create(rank, nranks, ids_per_elem, nfields);
  if (!d) return nullptr;
  if (n_owned < 0 || n_ghosted < 0 || (n_owned && !owned) || (n_ghosted && (!ghosted || !ghosted_owner))) { delete d; fail("dofmgr_from_arrays: bad arguments"); return nullptr; }
  d->owned.assign(owned, owned + n_owned);
  d->ghosted.assign(ghosted, ghosted + n_ghosted);
  d->ghosted_owner.assign(ghosted_owner, ghosted_owner + n_ghosted);
  d->my_offset = my_offset;
  d->state = 6;
  return d;
}
void txhost_dofmgr_destroy(txhost_dofmgr d) { delete d; }

int txhost_dofmgr_set_connectivity(txhost_dofmgr d, int64_t ne, const int64_t *conn)
{
  d->ne = ne; d->conn.assign(conn, conn + ne * d->ipe); d->state = 0;
  return 0;
}

int txhost_dofmgr_step(txhost_dofmgr d, const int64_t *rc, const int64_t *rb, const int64_t **sc, const int64_t **sb, int *done)
{
  const int P = d->P;
  *done = 0;
  std::vector<std::vector<int64_t>> per(P);
  switch (d->state) {
    case 0: {  // overlap map; ask the directory (rank id % P) who owns each id -- Tpetra::createOneToOne
      const int64_t n = (int64_t)d->conn.size();
      d->eov.resize(n);
      int64_t lo = INT64_MAX, hi = INT64_MIN;
#pragma omp parallel for schedule(static) reduction(min : lo) reduction(max : hi)
      for (int64_t i = 0; i < n; ++i) { lo = std::min(lo, d->conn[i]); hi = std::max(hi, d->conn[i]); }
      if (n > 0 && (hi - lo) / 16 <= n) {
        // ids of a mesh are dense in [lo, hi]: a bitmap + prefix popcounts give the sorted unique list and the rank of
        // every id in O(n) (same result as sort + unique + binary search)
        const int64_t nw = (hi - lo) / 64 + 1;
        std::vector<uint64_t> bits((size_t)nw, 0);
        for (int64_t i = 0; i < n; ++i) { const int64_t k = d->conn[i] - lo; bits[(size_t)(k >> 6)] |= 1ull << (k & 63); }
        std::vector<int64_t> pre((size_t)nw + 1, 0);
        for (int64_t w = 0; w < nw; ++w) pre[(size_t)w + 1] = pre[(size_t)w] + __builtin_popcountll(bits[(size_t)w]);
        d->ov.resize((size_t)pre[(size_t)nw]);
#pragma omp parallel for schedule(static)
        for (int64_t w = 0; w < nw; ++w) {
          uint64_t b = bits[(size_t)w];
          int64_t o = pre[(size_t)w];
          while (b) { const int t = __builtin_ctzll(b); d->ov[(size_t)o++] = lo + w * 64 + t; b &= b - 1; }
        }
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
          const int64_t k = d->conn[i] - lo;
          d->eov[i] = pre[(size_t)(k >> 6)] + __builtin_popcountll(bits[(size_t)(k >> 6)] & ((1ull << (k & 63)) - 1));
        }
      } else {
        d->ov = d->conn;
        __gnu_parallel::sort(d->ov.begin(), d->ov.end());
        d->ov.erase(std::unique(d->ov.begin(), d->ov.end()), d->ov.end());
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i)
          d->eov[i] = std::lower_bound(d->ov.begin(), d->ov.end(), d->conn[i]) - d->ov.begin();
      }
      d->req_ids.assign(P, {});
      for (int r = 0; r < P; ++r) d->req_ids[r].reserve(d->ov.size() / P + 16);
      for (int64_t id : d->ov) d->req_ids[(int)(id % P)].push_back(id);
      pack(d, d->req_ids);
      break;
    }
    case 1: {  // directory: GreedyTieBreak -> smallest rank holding the id (:108-133); reply in request order
      int64_t tot = 0;
      for (int r = 0; r < P; ++r) tot += rc[r];
      int64_t lo = INT64_MAX, hi = INT64_MIN;
      for (int64_t k = 0; k < tot; ++k) { lo = std::min(lo, rb[k]); hi = std::max(hi, rb[k]); }
      if (tot > 0 && (hi - lo) / P / 16 <= tot) {
        // the ids hashed to me are lo, lo + P, ...: one table slot per id, requests visited by ascending rank
        std::vector<int> own((size_t)((hi - lo) / P + 1), -1);
        int64_t off = 0;
        for (int r = 0; r < P; ++r) {
          for (int64_t k = 0; k < rc[r]; ++k) { int &o = own[(size_t)((rb[off + k] - lo) / P)]; if (o < 0) o = r; }
          off += rc[r];
        }
        off = 0;
        for (int r = 0; r < P; ++r) {
          per[r].resize((size_t)rc[r]);
          for (int64_t k = 0; k < rc[r]; ++k) per[r][(size_t)k] = own[(size_t)((rb[off + k] - lo) / P)];
          off += rc[r];
        }
      } else {
        std::vector<std::pair<int64_t, int>> all;
        int64_t off = 0;
        for (int r = 0; r < P; ++r) { for (int64_t k = 0; k < rc[r]; ++k) all.emplace_back(rb[off + k], r); off += rc[r]; }
        std::vector<std::pair<int64_t, int>> srt = all;
        __gnu_parallel::sort(srt.begin(), srt.end());
        // first entry of each id in (id, rank) order is the owner
        std::vector<int64_t> uid; std::vector<int> uown;
        for (size_t i = 0; i < srt.size(); ++i)
          if (i == 0 || srt[i].first != srt[i - 1].first) { uid.push_back(srt[i].first); uown.push_back(srt[i].second); }
        for (auto &q : all) {
          const size_t u = std::lower_bound(uid.begin(), uid.end(), q.first) - uid.begin();
          per[q.second].push_back(uown[u]);
        }
      }
      pack(d, per);
      break;
    }
    case 2: {  // owners known; count my owned ids and all-gather the counts (Teuchos::scan, :794-813)
      d->ov_owner.assign(d->ov.size(), -1);
      std::vector<int64_t> cursor(P, 0), start(P + 1, 0);
      for (int r = 0; r < P; ++r) start[r + 1] = start[r] + rc[r];
      for (size_t i = 0; i < d->ov.size(); ++i) {
        const int dr = (int)(d->ov[i] % P);
        d->ov_owner[i] = (int)rb[start[dr] + cursor[dr]++];
      }
      int64_t cnt = 0;
      for (int o : d->ov_owner) cnt += (o == d->rank);
      for (int r = 0; r < P; ++r) per[r].push_back(cnt * d->nf);
      pack(d, per);
      break;
    }
    case 3: {  // my offset = exclusive scan; number my owned ids in overlap-map order, fields inner (:823-843);
               // ask owners for the GIDs of the ids I do not own (Import REPLACE, :858)
      int64_t off = 0;
      for (int r = 0; r < d->rank; ++r) off += rb[r];   // rc[r]==1 each
      d->my_offset = off;
      d->ov_gid0.assign(d->ov.size(), -1);
      int64_t which = 0;
      d->req_ids.assign(P, {});
      for (size_t i = 0; i < d->ov.size(); ++i) {
        if (d->ov_owner[i] == d->rank) { d->ov_gid0[i] = off + which; which += d->nf; }
        else d->req_ids[d->ov_owner[i]].push_back(d->ov[i]);
      }
      pack(d, d->req_ids);
      break;
    }
    case 4: {  // owner: answer GID requests
      int64_t off = 0;
      for (int r = 0; r < P; ++r) {
        for (int64_t k = 0; k < rc[r]; ++k) {
          const int64_t id = rb[off + k];
          const size_t i = std::lower_bound(d->ov.begin(), d->ov.end(), id) - d->ov.begin();
          if (i >= d->ov.size() || d->ov[i] != id || d->ov_owner[i] != d->rank) return fail("GID request for an id this rank does not own");
          per[r].push_back(d->ov_gid0[i]);
        }
        off += rc[r];
      }
      pack(d, per);
      break;
    }
    case 5: {  // finalize: elementGIDs_ (:1293-1349), owned_ (:580-636), ghosted_ (:650-695), LIDs
      std::vector<int64_t> start(P + 1, 0), cursor(P, 0);
      for (int r = 0; r < P; ++r) start[r + 1] = start[r] + rc[r];
      for (size_t i = 0; i < d->ov.size(); ++i) {
        const int o = d->ov_owner[i];
        if (o != d->rank) d->ov_gid0[i] = rb[start[o] + cursor[o]++];
      }
      const int gpe = d->ipe * d->nf;
      const int64_t ne = d->ne;
      d->egids.resize(ne * gpe); d->elids.resize(ne * gpe);
#pragma omp parallel for schedule(static)
      for (int64_t e = 0; e < ne; ++e)
        for (int c = 0; c < d->ipe; ++c)
          for (int f = 0; f < d->nf; ++f) d->egids[e * gpe + c * d->nf + f] = d->ov_gid0[d->eov[e * d->ipe + c]] + f;
      std::vector<int64_t> lid0(d->ov.size(), -1);
      d->owned.clear(); d->ghosted.clear(); d->ghosted_owner.clear();
      // first touch in element order: owned ids first (:580-636), then everything else (:650-695).  One pass records the
      // order in which the ids are met; owned and ghosted ids keep that order inside their class.
      std::vector<int64_t> order;
      order.reserve(d->ov.size());
      for (int64_t i = 0; i < ne * d->ipe; ++i) {
        const int64_t o = d->eov[i];
        if (lid0[o] < 0) { lid0[o] = 0; order.push_back(o); }
      }
      // :635-636 "remaining owned" cannot occur: every overlap id comes from one of my elements
      for (int64_t o : order)
        if (d->ov_owner[o] == d->rank) {
          lid0[o] = (int64_t)d->owned.size();
          for (int f = 0; f < d->nf; ++f) d->owned.push_back(d->ov_gid0[o] + f);
        }
      const int64_t no = (int64_t)d->owned.size();
      for (int64_t o : order)
        if (d->ov_owner[o] != d->rank) {
          lid0[o] = no + (int64_t)d->ghosted.size();
          for (int f = 0; f < d->nf; ++f) { d->ghosted.push_back(d->ov_gid0[o] + f); d->ghosted_owner.push_back(d->ov_owner[o]); }
        }
#pragma omp parallel for schedule(static)
      for (int64_t e = 0; e < ne; ++e)
        for (int c = 0; c < d->ipe; ++c)
          for (int f = 0; f < d->nf; ++f) d->elids[e * gpe + c * d->nf + f] = (int)(lid0[d->eov[e * d->ipe + c]] + f);
      d->scounts.assign(P, 0); d->sbuf.clear();
      *done = 1;
      break;
    }
    default: return fail("dofmgr_step called after completion");
  }
  d->state++;
  *sc = d->scounts.data(); *sb = d->sbuf.data();
  return 0;
}

int64_t txhost_dofmgr_num_owned(txhost_dofmgr d) { return (int64_t)d->owned.size(); }
int64_t txhost_dofmgr_num_ghosted(txhost_dofmgr d) { return (int64_t)d->ghosted.size(); }
int txhost_dofmgr_get_owned(txhost_dofmgr d, int64_t *o) { memcpy(o, d->owned.data(), d->owned.size() * 8); return 0; }
int txhost_dofmgr_get_ghosted(txhost_dofmgr d, int64_t *o) { memcpy(o, d->ghosted.data(), d->ghosted.size() * 8); return 0; }
int txhost_dofmgr_get_ghosted_owner(txhost_dofmgr d, int *o) { memcpy(o, d->ghosted_owner.data(), d->ghosted_owner.size() * 4); return 0; }
int txhost_dofmgr_get_elem_gids(txhost_dofmgr d, int64_t *o) { memcpy(o, d->egids.data(), d->egids.size() * 8); return 0; }
int txhost_dofmgr_get_elem_lids(txhost_dofmgr d, int *o) { memcpy(o, d->elids.data(), d->elids.size() * 4); return 0; }
// FieldAggPattern (dof-mgr/src/Panzer_FieldAggPattern.cpp:201-276): per subcell, fields in order
int txhost_dofmgr_field_offsets(txhost_dofmgr d, int field, int *o) { for (int b = 0; b < d->ipe; ++b) o[b] = b * d->nf + field; return 0; }

}  // extern "C"

// =====================================================================================
// TpetraLinearObjFactory  (disc-fe/src/lof/Panzer_TpetraLinearObjFactory_impl.hpp)
// =====================================================================================
struct txhost_lof_s {
  txhost_dofmgr d;
  int state = 0;
  int64_t n_owned = 0, n_local = 0;
  std::vector<int64_t> rowptr; std::vector<int> colind;          // ghosted graph
  std::vector<int64_t> gid_of_lid;
  // halo
  std::vector<int> nbr;                                          // neighbour ranks, ascending
  std::vector<int64_t> send_off, recv_off; std::vector<int> send_lids, recv_lids;
  // fill graph + matrix plan
  std::vector<int64_t> frowptr; std::vector<int> fcolind; std::vector<int64_t> col_gids;
  std::vector<int64_t> mrecv_off, mrecv_pos;
  std::vector<int> pair_rows, pair_cols;                         // every received (owned row, local column), in plan order
  bool compact = false;                                          // only the ghost rows of the graph are here (the rest stays on the device)
  std::vector<int64_t> scounts, sbuf;
};

namespace {
// buildGhostedGraph (:558-650): every element inserts all its GIDs into the row of each of its GIDs;
// fillComplete sorts by local column index and merges duplicates.
void ghosted_graph(txhost_lof_s *l)
{
  auto *d = l->d;
  const int gpe = d->ipe * d->nf;
  const int64_t nl = l->n_local, ne = d->ne;
  std::vector<int64_t> cnt(nl + 1, 0);
  for (int64_t i = 0; i < ne * gpe; ++i) cnt[d->elids[i]] += gpe;
  std::vector<int64_t> start(nl + 1, 0);
  for (int64_t i = 0; i < nl; ++i) start[i + 1] = start[i] + cnt[i];
  std::vector<int> raw((size_t)start[nl]);
  std::fill(cnt.begin(), cnt.end(), 0);
  for (int64_t e = 0; e < ne; ++e)
    for (int j = 0; j < gpe; ++j) {
      const int row = d->elids[e * gpe + j];
      int *dst = raw.data() + start[row] + cnt[row];
      for (int k = 0; k < gpe; ++k) dst[k] = d->elids[e * gpe + k];
      cnt[row] += gpe;
    }
  l->rowptr.assign(nl + 1, 0);
#pragma omp parallel for schedule(dynamic, 4096)
  for (int64_t i = 0; i < nl; ++i) {
    int *b = raw.data() + start[i], *e = b + cnt[i];
    std::sort(b, e);
    cnt[i] = std::unique(b, e) - b;
  }
  for (int64_t i = 0; i < nl; ++i) l->rowptr[i + 1] = l->rowptr[i] + cnt[i];
  l->colind.resize((size_t)l->rowptr[nl]);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nl; ++i) std::copy(raw.data() + start[i], raw.data() + start[i] + cnt[i], l->colind.data() + l->rowptr[i]);
}
}  // namespace

extern "C" {

txhost_lof txhost_lof_create(txhost_dofmgr d)
{
  if (!d || d->state < 6) { fail("lof_create: DOF manager not built"); return nullptr; }
  auto *l = new txhost_lof_s();
  l->d = d;
  l->n_owned = (int64_t)d->owned.size();
  l->n_local = l->n_owned + (int64_t)d->ghosted.size();
  l->gid_of_lid = d->owned;
  l->gid_of_lid.insert(l->gid_of_lid.end(), d->ghosted.begin(), d->ghosted.end());
  return l;
}
void txhost_lof_destroy(txhost_lof l) { delete l; }

int txhost_lof_ghosted_graph(txhost_lof l, int64_t *nnz)
{
  if (l->rowptr.empty()) ghosted_graph(l);
  if (nnz) *nnz = l->rowptr.back();
  return 0;
}
// adopt a ghosted graph built elsewhere (e.g. on the device by txasm_graph_build)
int txhost_lof_set_ghosted_graph(txhost_lof l, const int64_t *rowptr, const int *colind)
{
  l->rowptr.assign(rowptr, rowptr + l->n_local + 1);
  l->colind.assign(colind, colind + l->rowptr.back());
  return 0;
}
// Compact mode: only the ghost rows [n_owned, n_local) of the ghosted graph (rowptr rebased to 0).  The plan then
// carries the received (row, column) pairs instead of the fill graph; whoever holds the graph (the device,
// txasm_graph_merge_columns) inserts them and returns the matrix positions.
int txhost_lof_set_ghost_rows(txhost_lof l, const int64_t *rowptr, const int *colind)
{
  const int64_t ng = l->n_local - l->n_owned;
  l->rowptr.assign((size_t)l->n_local + 1, 0);
  for (int64_t i = 0; i <= ng; ++i) l->rowptr[(size_t)(l->n_owned + i)] = rowptr[i];
  l->colind.assign(colind, colind + rowptr[ng]);
  l->compact = true;
  return 0;
}
int64_t txhost_lof_num_pairs(txhost_lof l) { return l->state == 2 ? (int64_t)l->pair_rows.size() : -1; }
int txhost_lof_get_pairs(txhost_lof l, int *rows, int *cols)
{
  if (l->state != 2) return fail("lof plan not built");
  if (rows) memcpy(rows, l->pair_rows.data(), l->pair_rows.size() * 4);
  if (cols) memcpy(cols, l->pair_cols.data(), l->pair_cols.size() * 4);
  return 0;
}
int txhost_lof_get_ghosted_graph(txhost_lof l, int64_t *rowptr, int *colind)
{
  if (l->rowptr.empty()) ghosted_graph(l);
  if (rowptr) memcpy(rowptr, l->rowptr.data(), l->rowptr.size() * 8);
  if (colind) memcpy(colind, l->colind.data(), l->colind.size() * 4);
  return 0;
}

// One exchange builds the Import/Export plans (what Tpetra::Import/Export constructors negotiate):
//  send to the owner of each ghost row, in my ghosted_ order:  [gid, ncols, (col_gid, col_owner) x ncols]
//  the owner derives (a) which of its owned LIDs I ghost, in my order  -> its send_lids / my recv_lids
//                    (b) the fill graph: owned rows gain the remote columns (buildGraph's Export INSERT, :534-556)
//                    (c) for every value it will receive, the index into its A_values
int txhost_lof_step(txhost_lof l, const int64_t *rc, const int64_t *rb, const int64_t **sc, const int64_t **sb, int *done)
{
  auto *d = l->d;
  const int P = d->P;
  *done = 0;
  if (l->rowptr.empty()) ghosted_graph(l);
  if (l->state == 0) {
    std::vector<std::vector<int64_t>> per(P);
    // ghost rows grouped by owner (ascending rank), inside a group in ghosted_ order
    std::vector<std::vector<int>> rows(P);
    for (size_t g = 0; g < d->ghosted.size(); ++g) rows[d->ghosted_owner[g]].push_back((int)(l->n_owned + g));
    // owner of a local column: me for owned LIDs, ghosted_owner otherwise
    auto owner_of = [&](int lid) { return lid < l->n_owned ? d->rank : d->ghosted_owner[lid - l->n_owned]; };
    l->recv_off.assign(1, 0); l->recv_lids.clear();
    std::vector<int> nb;
    for (int r = 0; r < P; ++r) {
      if (rows[r].empty()) continue;
      nb.push_back(r);
      for (int lid : rows[r]) {
        l->recv_lids.push_back(lid);
        per[r].push_back(l->gid_of_lid[lid]);
        per[r].push_back(l->rowptr[lid + 1] - l->rowptr[lid]);
        for (int64_t k = l->rowptr[lid]; k < l->rowptr[lid + 1]; ++k) {
          per[r].push_back(l->gid_of_lid[l->colind[k]]);
          per[r].push_back(owner_of(l->colind[k]));
        }
      }
    }
    l->nbr = nb;   // provisional: ranks I ghost from; ranks ghosting from me are added in state 1
    l->scounts.assign(P, 0); l->sbuf.clear();
    for (int r = 0; r < P; ++r) { l->scounts[r] = (int64_t)per[r].size(); l->sbuf.insert(l->sbuf.end(), per[r].begin(), per[r].end()); }
    l->state = 1;
    *sc = l->scounts.data(); *sb = l->sbuf.data();
    return 0;
  }
  if (l->state == 1) {
    // GID -> local LID: the owned GIDs are the contiguous range [my_offset, my_offset + n_owned) (buildGlobalUnknowns state 3)
    // in first-touch LID order -> one direct table; the ghost GIDs (surface size) -> a sorted list
    const int64_t g0 = d->my_offset, no = l->n_owned;
    std::vector<int> lid_owned((size_t)no);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < no; ++i) lid_owned[(size_t)(l->gid_of_lid[i] - g0)] = (int)i;
    std::vector<std::pair<int64_t, int>> ghost_lid;
    ghost_lid.reserve((size_t)(l->n_local - no));
    for (int64_t i = no; i < l->n_local; ++i) ghost_lid.emplace_back(l->gid_of_lid[i], (int)i);
    std::sort(ghost_lid.begin(), ghost_lid.end());
    auto find_lid = [&](int64_t g) -> int {
      if (g >= g0 && g < g0 + no) return lid_owned[(size_t)(g - g0)];
      auto it = std::lower_bound(ghost_lid.begin(), ghost_lid.end(), std::make_pair(g, -1));
      return (it != ghost_lid.end() && it->first == g) ? it->second : -1;
    };
    // pass 1: the received rows, neighbour by neighbour, row by row, column by column
    std::vector<std::vector<int>> rec_rows(P);          // owned LID of every received row
    std::vector<int64_t> pair_gid; std::vector<int> pair_row, pair_from;   // every received (row, column GID)
    std::vector<std::pair<int, int64_t>> newcols;       // (owner, gid) of columns unknown to my ghosted map
    int64_t off = 0;
    for (int r = 0; r < P; ++r) {
      const int64_t end = off + rc[r];
      while (off < end) {
        const int64_t gid = rb[off++], nc = rb[off++];
        if (gid < g0 || gid >= g0 + no) return fail("received a ghost row this rank does not own");
        const int row = lid_owned[(size_t)(gid - g0)];
        rec_rows[r].push_back(row);
        for (int64_t k = 0; k < nc; ++k) {
          const int64_t cg = rb[off++]; const int co = (int)rb[off++];
          pair_row.push_back(row); pair_gid.push_back(cg); pair_from.push_back(r);
          if (find_lid(cg) < 0) newcols.emplace_back(co, cg);
        }
      }
    }
    // neighbours = ranks I receive x from (recv side) U ranks that sent me rows (send side)
    std::vector<char> isn(P, 0);
    for (int r : l->nbr) isn[r] = 1;
    for (int r = 0; r < P; ++r) if (!rec_rows[r].empty()) isn[r] = 1;
    // rebuild recv_off per final neighbour list (recv_lids were pushed grouped by owner ascending)
    std::vector<std::vector<int>> myrecv(P);
    {
      size_t p = 0;
      std::vector<int64_t> cntr(P, 0);
      for (size_t g = 0; g < d->ghosted.size(); ++g) cntr[d->ghosted_owner[g]]++;
      for (int r = 0; r < P; ++r) { myrecv[r].assign(l->recv_lids.begin() + p, l->recv_lids.begin() + p + cntr[r]); p += cntr[r]; }
    }
    l->nbr.clear();
    for (int r = 0; r < P; ++r) if (isn[r]) l->nbr.push_back(r);
    const int nn = (int)l->nbr.size();
    l->recv_off.assign(nn + 1, 0); l->send_off.assign(nn + 1, 0);
    l->recv_lids.clear(); l->send_lids.clear();
    for (int k = 0; k < nn; ++k) {
      const int r = l->nbr[k];
      l->recv_lids.insert(l->recv_lids.end(), myrecv[r].begin(), myrecv[r].end());
      l->recv_off[k + 1] = (int64_t)l->recv_lids.size();
      l->send_lids.insert(l->send_lids.end(), rec_rows[r].begin(), rec_rows[r].end());
      l->send_off[k + 1] = (int64_t)l->send_lids.size();
    }
    // remote-only columns: appended after my n_local columns, grouped by owning rank then GID
    // (Tpetra's makeColMap order for remote GIDs -- assumption A5 of SURVEY.md appendix A)
    std::sort(newcols.begin(), newcols.end());
    newcols.erase(std::unique(newcols.begin(), newcols.end()), newcols.end());
    l->col_gids = l->gid_of_lid;
    std::vector<std::pair<int64_t, int>> new_lid;       // (gid, column index) of the remote-only columns
    for (auto &nc : newcols) { new_lid.emplace_back(nc.second, (int)l->col_gids.size()); l->col_gids.push_back(nc.second); }
    std::sort(new_lid.begin(), new_lid.end());
    // local column of every received pair; the pairs are already in plan order (rank ascending = neighbour order)
    const int64_t np = (int64_t)pair_row.size();
    l->pair_rows = pair_row;
    l->pair_cols.resize((size_t)np);
    for (int64_t i = 0; i < np; ++i) {
      int lc = find_lid(pair_gid[(size_t)i]);
      if (lc < 0) {
        auto it = std::lower_bound(new_lid.begin(), new_lid.end(), std::make_pair(pair_gid[(size_t)i], -1));
        lc = it->second;
      }
      l->pair_cols[(size_t)i] = lc;
    }
    l->mrecv_off.assign(nn + 1, 0);
    {
      std::vector<int64_t> per_rank(P, 0);
      for (int64_t i = 0; i < np; ++i) per_rank[pair_from[(size_t)i]]++;
      for (int k = 0; k < nn; ++k) l->mrecv_off[k + 1] = l->mrecv_off[k] + per_rank[l->nbr[k]];
    }
    l->mrecv_pos.clear();
    l->frowptr.clear(); l->fcolind.clear();
    if (!l->compact) {
      // fill graph: ghost rows unchanged, owned rows = local columns U received columns, sorted by local column.
      // Only the rows that received something are merged; the others are copied.
      std::vector<std::pair<int, int>> ex((size_t)np);
      for (int64_t i = 0; i < np; ++i) ex[(size_t)i] = {l->pair_rows[(size_t)i], l->pair_cols[(size_t)i]};
      std::sort(ex.begin(), ex.end());
      ex.erase(std::unique(ex.begin(), ex.end()), ex.end());
      l->frowptr.assign(l->n_local + 1, 0);
      {
        size_t c = 0;
        for (int64_t i = 0; i < l->n_local; ++i) {
          const int *b = l->colind.data() + l->rowptr[i], *e = l->colind.data() + l->rowptr[i + 1];
          int64_t len = e - b;
          for (; c < ex.size() && ex[c].first == i; ++c) if (!std::binary_search(b, e, ex[c].second)) ++len;
          l->frowptr[i + 1] = l->frowptr[i] + len;
        }
      }
      l->fcolind.resize((size_t)l->frowptr[l->n_local]);
      {
        size_t c = 0;
        for (int64_t i = 0; i < l->n_local; ++i) {
          const int *b = l->colind.data() + l->rowptr[i], *e = l->colind.data() + l->rowptr[i + 1];
          int *o = l->fcolind.data() + l->frowptr[i];
          if (c >= ex.size() || ex[c].first != i) { std::copy(b, e, o); continue; }
          size_t c1 = c;
          while (c1 < ex.size() && ex[c1].first == i) ++c1;
          std::vector<int> add;
          for (size_t q = c; q < c1; ++q) add.push_back(ex[q].second);
          int *oe = std::set_union(b, e, add.begin(), add.end(), o);
          (void)oe;
          c = c1;
        }
      }
      // positions of the values I will receive, in plan order
      l->mrecv_pos.resize((size_t)np);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < np; ++i) {
        const int row = l->pair_rows[(size_t)i], lc = l->pair_cols[(size_t)i];
        const int *b = l->fcolind.data() + l->frowptr[row], *e = l->fcolind.data() + l->frowptr[row + 1];
        const int *p = std::lower_bound(b, e, lc);
        l->mrecv_pos[(size_t)i] = (p != e && *p == lc) ? (int64_t)(l->frowptr[row] + (p - b)) : -1;
      }
    }
    l->scounts.assign(P, 0); l->sbuf.clear();
    l->state = 2;
    *done = 1;
    *sc = l->scounts.data(); *sb = l->sbuf.data();
    return 0;
  }
  return fail("lof_step called after completion");
}

int txhost_lof_num_neighbors(txhost_lof l) { return (int)l->nbr.size(); }

int txhost_lof_halo_sizes(txhost_lof l, int64_t *ns, int64_t *nr, int64_t *nmr, int64_t *fnnz, int64_t *ncols)
{
  if (l->state != 2) return fail("lof plan not built");
  if (ns) *ns = (int64_t)l->send_lids.size();
  if (nr) *nr = (int64_t)l->recv_lids.size();
  if (nmr) *nmr = (int64_t)l->mrecv_pos.size();
  if (fnnz) *fnnz = l->frowptr.empty() ? 0 : l->frowptr.back();
  if (ncols) *ncols = (int64_t)l->col_gids.size();
  return 0;
}

int txhost_lof_get_halo(txhost_lof l, int *nbr_rank, int64_t *send_off, int *send_lids, int64_t *recv_off, int *recv_lids)
{
  if (l->state != 2) return fail("lof plan not built");
  if (nbr_rank) memcpy(nbr_rank, l->nbr.data(), l->nbr.size() * 4);
  if (send_off) memcpy(send_off, l->send_off.data(), l->send_off.size() * 8);
  if (send_lids) memcpy(send_lids, l->send_lids.data(), l->send_lids.size() * 4);
  if (recv_off) memcpy(recv_off, l->recv_off.data(), l->recv_off.size() * 8);
  if (recv_lids) memcpy(recv_lids, l->recv_lids.data(), l->recv_lids.size() * 4);
  return 0;
}

int txhost_lof_get_fill_graph(txhost_lof l, int64_t *rowptr, int *colind, int64_t *col_gids)
{
  if (l->state != 2) return fail("lof plan not built");
  if (rowptr && !l->frowptr.empty()) memcpy(rowptr, l->frowptr.data(), l->frowptr.size() * 8);
  if (colind && !l->fcolind.empty()) memcpy(colind, l->fcolind.data(), l->fcolind.size() * 4);
  if (col_gids) memcpy(col_gids, l->col_gids.data(), l->col_gids.size() * 8);
  return 0;
}

int txhost_lof_get_matrix_plan(txhost_lof l, int64_t *off, int64_t *pos)
{
  if (l->state != 2) return fail("lof plan not built");
  if (off) memcpy(off, l->mrecv_off.data(), l->mrecv_off.size() * 8);
  if (pos) memcpy(pos, l->mrecv_pos.data(), l->mrecv_pos.size() * 8);
  return 0;
}

}  // extern "C"
