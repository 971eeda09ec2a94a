// fill_brick.cu -- k_fill_brick: the uniform tiles whose cells form a full tensor brick (every tile of an inline
// CubeHexMeshFactory mesh away from the domain boundary).
//
// Same contract as k_fill_rowtile / k_fill_uniform (fill_rowtile.cu): owner-computes, every value of A and f is
// written once, no atomics, bitwise reproducible.  What changes is where the per-tile work goes.  For a brick tile
// the nodes the tile touches are a (cx+1) x (cy+1) x (cz+1) lattice (<= 10 x 10 x 6 for a 256-row tile), so
//
//   * the tile is described by ONE record (BrickRec, 2.9 KB, one TMA bulk load, prefetched two tiles ahead): the
//     LID of every lattice node and the lattice position of every row -- instead of the 405 x 8 tile-ordered LID copy,
//     the per-row cell table and the row list (18.5 KB) the cell-based kernels read;
//   * phase 1 gathers the solution ONCE PER NODE (600 loads) instead of once per cell vertex (3240), into a lattice
//     in shared memory;
//   * the separable source load vector  int phi_row s  = det * X(i) Y(j) Z(k)  needs the 1-D factors of the lattice
//     lines only: 4 sine evaluations for each of the 8 + 8 + 4 interior lattice coordinates per tile instead of 6 per
//     cell (2430 per tile);
//   * phase 2 (thread per row) is the 27-point stencil  f = sum_j Kf[j] u[node + off_j] + source  on that lattice;
//   * A is the constant row image cK*Kf (+ cM*Mf) streamed out by TMA bulk stores exactly as in k_fill_uniform.
//
// Reference semantics reproduced: GatherSolution -> DOFGradient/DOF -> Integrator_GradBasisDotVector /
// Integrator_BasisTimesScalar -> ScatterResidual (disc-fe/src/evaluators/Panzer_*_impl.hpp, see fill_rowtile.cu) for
// cells with a constant diagonal Jacobian, where the sum over the 8 cells around a node collapses to the stencil
// above (Kf, Mf are sums of the exact element integrals of elem_q1hex.cuh).  Values agree with the cell-by-cell
// kernels to rounding (different summation order), never bit for bit; tests compare both against the oracle.
#include "tiles.hpp"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <vector>

namespace txasm {

constexpr int BRICK_NODE_CAP = 600;
constexpr int BRICK_ROWS = 256;
constexpr int BRICK_DIM_CAP = 16;       // lattice nodes per axis (positions travel in 4 bits)
constexpr int BRICK_CELL_CAP = 1024;    // cells per tile the classification kernel can hold
constexpr int BRICK_RUNS_INLINE = 32;   // an 8 x 8 x 4-node tile of a lexicographically numbered mesh has 32 x-line runs
constexpr int SHAPE_STRIDE = 32;        // Kf[27] | Jxx Jyy Jzz det | pad -- a row of Tiles::d_tile_kf

struct BrickRec {
  int nxs, nys, nzs;                    // lattice nodes per axis
  int n_rows;
  int shape;                            // row of the shape table
  int n_nodes;
  int n_runs;                           // runs of rows contiguous in A (TMA bulk stores)
  int runs_inline;                      // 1: they are in runs[] below, 0: in Tiles::d_runs at run_ptr[tile]
  RowRun runs[BRICK_RUNS_INLINE];
  unsigned short rowpos[BRICK_ROWS];    // per tile row (slot order = ascending LID): i | j << 4 | k << 8, 0xFFFF: no row
  int nodes[BRICK_NODE_CAP];            // LID of lattice node (i, j, k) at i + nxs * (j + nys * k)
};
static_assert(sizeof(BrickRec) % 16 == 0, "BrickRec is moved by 16-byte cp.async");

// Lattice tiles with rows on their faces (k_fill_edge): thin slabs on a face of the mesh reach 18 x 18 x 3 nodes, so the
// lattice is larger and positions travel in 5 bits per axis.
constexpr int EDGE_CODE_CAP = BRICK_ROWS * 27 + 8;       // entry codes of one tile in shared memory (+ the 16-byte staging granule)
constexpr int EDGE_NODE_CAP = 1024;
constexpr int EDGE_DIM_CAP = 32;
struct EdgeRec {
  int nxs, nys, nzs;
  int n_rows;
  int shape;
  int n_nodes;
  int n_runs;                           // runs of the tile (Tiles::d_runs at run_beg)
  int code_cnt;                         // entry codes of its non-uniform runs: Tiles::d_edge_code[code_beg, code_beg + code_cnt)
  long long run_beg;
  long long code_beg;
  unsigned short rowpos[BRICK_ROWS];    // i | j << 5 | k << 10, 0xFFFF: no row
  int nodes[EDGE_NODE_CAP];
};
static_assert(sizeof(EdgeRec) % 16 == 0 && offsetof(EdgeRec, nodes) % 16 == 0, "EdgeRec is read by 16-byte loads");

__device__ __forceinline__ unsigned long long brick_dbl_key(double v)
{
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double brick_key_dbl(unsigned long long k)
{
  unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// One CTA per tile.  WRITE = false: flag[t] = 1 when the tile is a brick tile.  WRITE = true: fill rec[t] (the tile
// is known to qualify).  A tile qualifies when it is uniform (class 7: congruent axis-aligned cells, every row
// interior with canonical column order), its cells tile a cx x cy x cz box exactly once each with positive axes,
// the lattice fits the record, and cells sharing a lattice node agree on its LID.
// EDGE = true: the tile need not be uniform -- any tile of congruent axis-aligned cells whose cells fill a box; its rows may
// sit anywhere on the lattice (faces, edges, corners of the mesh or of the rank's brick): k_fill_edge takes those.
template <bool WRITE, bool EDGE>
__global__ void __launch_bounds__(BRICK_ROWS) k_brick_scan(int t0, int n_tiles, const int *__restrict__ tile_rows,
                                                           const int64_t *__restrict__ cell_ptr, const int *__restrict__ cells,
                                                           const unsigned short *__restrict__ adjl, const int *__restrict__ lids,
                                                           const double *__restrict__ xyz, const double *__restrict__ tile_kf,
                                                           const unsigned char *__restrict__ tile_cong, double tol,
                                                           unsigned char *__restrict__ flag, void *__restrict__ rec_)
{
  constexpr int NODE_CAP = EDGE ? EDGE_NODE_CAP : BRICK_NODE_CAP, DIM_CAP = EDGE ? EDGE_DIM_CAP : BRICK_DIM_CAP;
  constexpr int PB = EDGE ? 5 : 4, PM = (1 << PB) - 1;           // bits per lattice coordinate
  __shared__ int nodes[EDGE_NODE_CAP];
  __shared__ int occ[BRICK_CELL_CAP];
  __shared__ unsigned short cellpos[BRICK_CELL_CAP];
  __shared__ unsigned long long xmin[3];
  __shared__ int dims[3], bad, nrows;
  const int t = t0 + blockIdx.x, tid = threadIdx.x;
  if (t >= n_tiles) return;
  const int64_t cb = cell_ptr[t];
  const int nc = (int)(cell_ptr[t + 1] - cb);
  const double h[3] = {tile_kf[(int64_t)t * KF_STRIDE + 27], tile_kf[(int64_t)t * KF_STRIDE + 28], tile_kf[(int64_t)t * KF_STRIDE + 29]};
  if (tid == 0) {
    bad = (EDGE ? ((tile_cong[t] & 3) != 3 || (tile_cong[t] & 8) != 0) : ((tile_cong[t] & 7) != 7)) || nc <= 0 || nc > BRICK_CELL_CAP ||
          !(h[0] > 0.0) || !(h[1] > 0.0) || !(h[2] > 0.0);
    xmin[0] = xmin[1] = xmin[2] = ~0ull;
    dims[0] = dims[1] = dims[2] = 0;
    nrows = 0;
  }
  __syncthreads();
  if (bad) { if (!WRITE && tid == 0) flag[t] = 0; return; }
  for (int j = tid; j < nc; j += BRICK_ROWS) {
    const int64_t l0 = lids[(int64_t)cells[cb + j] * 8];
    for (int d = 0; d < 3; ++d) atomicMin(&xmin[d], brick_dbl_key(xyz[l0 * 3 + d]));
  }
  __syncthreads();
  const double hmax = fmax(h[0], fmax(h[1], h[2]));
  for (int j = tid; j < nc; j += BRICK_ROWS) {
    const int64_t l0 = lids[(int64_t)cells[cb + j] * 8];
    int p[3];
    bool ok = true;
    for (int d = 0; d < 3; ++d) {
      const double x0 = brick_key_dbl(xmin[d]), x = xyz[l0 * 3 + d];
      const double q = rint((x - x0) / (2.0 * h[d]));
      ok = ok && q >= 0.0 && q < (double)(DIM_CAP - 1) && fabs(x - (x0 + 2.0 * h[d] * q)) <= 4.0 * tol * hmax * (q + 1.0);
      p[d] = ok ? (int)q : 0;
      if (ok) atomicMax(&dims[d], p[d] + 1);
    }
    if (!ok) bad = 1;
    cellpos[j] = (unsigned short)(p[0] | (p[1] << PB) | (p[2] << (2 * PB)));
  }
  __syncthreads();
  const int cx = dims[0], cy = dims[1], cz = dims[2];
  const int nxs = cx + 1, nys = cy + 1, nzs = cz + 1, nn = nxs * nys * nzs;
  if (bad || cx * cy * cz != nc || nn > NODE_CAP) { if (!WRITE && tid == 0) flag[t] = 0; return; }
  for (int i = tid; i < nn; i += BRICK_ROWS) nodes[i] = -1;
  for (int i = tid; i < nc; i += BRICK_ROWS) occ[i] = 0;
  __syncthreads();
  for (int j = tid; j < nc; j += BRICK_ROWS) {
    const int cp = cellpos[j], pi = cp & PM, pj = (cp >> PB) & PM, pk = cp >> (2 * PB);
    if (atomicExch(&occ[pi + cx * (pj + cy * pk)], 1)) bad = 1;            // two cells at one lattice position
    const int *l = lids + (int64_t)cells[cb + j] * 8;
    for (int a = 0; a < 8; ++a) {
      const int idx = (pi + (hex_sx(a) > 0)) + nxs * ((pj + (hex_sy(a) > 0)) + nys * (pk + (hex_sz(a) > 0)));
      const int old = atomicCAS(&nodes[idx], -1, l[a]);
      if (old != -1 && old != l[a]) bad = 1;                                // cells disagree on the node
    }
  }
  __syncthreads();
  // rows: vertex a of one of the cells around them (brick tiles: vertex 6 (+,+,+) of the cell at (i-1, j-1, k-1))
  const int row = tile_rows[(int64_t)t * BRICK_ROWS + tid];
  unsigned short rp = 0xFFFF;
  if (row >= 0 && !bad) {
    int a = 6, el = adjl[((int64_t)t * BRICK_ROWS + tid) * 8 + 6];
    if (EDGE)                                 // a row on the boundary: any cell that has it as a vertex
      for (int aa = 0; aa < 8 && el == 0xFFFF; ++aa) { a = aa; el = adjl[((int64_t)t * BRICK_ROWS + tid) * 8 + aa]; }
    if (el == 0xFFFF || el >= nc) bad = 1;
    else {
      const int cp = cellpos[el];
      const int i = (cp & PM) + (hex_sx(a) > 0), j = ((cp >> PB) & PM) + (hex_sy(a) > 0), k = (cp >> (2 * PB)) + (hex_sz(a) > 0);
      const bool inside = i >= 1 && i <= nxs - 2 && j >= 1 && j <= nys - 2 && k >= 1 && k <= nzs - 2;
      if ((!EDGE && !inside) || i > PM || j > PM || k > PM || nodes[i + nxs * (j + nys * k)] != row) bad = 1;
      rp = (unsigned short)(i | (j << PB) | (k << (2 * PB)));
      atomicAdd(&nrows, 1);
    }
  }
  __syncthreads();
  if (!WRITE) { if (tid == 0) flag[t] = bad ? 0 : 1; return; }
  if (EDGE) {
    EdgeRec *r = reinterpret_cast<EdgeRec *>(rec_) + (t - t0);
    if (tid == 0) { r->nxs = nxs; r->nys = nys; r->nzs = nzs; r->n_rows = nrows; r->shape = bad ? -1 : 0; r->n_nodes = nn; }
    r->rowpos[tid] = rp;
    for (int i = tid; i < EDGE_NODE_CAP; i += BRICK_ROWS) r->nodes[i] = (i < nn) ? nodes[i] : 0;
  } else {
    BrickRec *r = reinterpret_cast<BrickRec *>(rec_) + (t - t0);
    if (tid == 0) { r->nxs = nxs; r->nys = nys; r->nzs = nzs; r->n_rows = nrows; r->shape = bad ? -1 : 0; r->n_nodes = nn; }
    r->rowpos[tid] = rp;
    for (int i = tid; i < BRICK_NODE_CAP; i += BRICK_ROWS) r->nodes[i] = (i < nn) ? nodes[i] : 0;
  }
}

__global__ void k_brick_mark(int n, const unsigned char *__restrict__ flag, unsigned char *__restrict__ tile_cong, int bit)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) tile_cong[t] = (unsigned char)((tile_cong[t] & (bit == 8 ? 7 : 15)) | (flag[t] ? bit : 0));
}
// rows of the brick tiles whose runs travel in the record: row -> shape id (else -1)
__global__ void k_brick_row_shape(int n, const int *__restrict__ shape, const int64_t *__restrict__ run_ptr,
                                  const int *__restrict__ tile_rows, int *__restrict__ row_shape)
{
  const int t = blockIdx.x, tid = threadIdx.x;
  if (t >= n || run_ptr[t + 1] - run_ptr[t] > BRICK_RUNS_INLINE) return;
  const int row = tile_rows[(int64_t)t * BRICK_ROWS + tid];
  if (row >= 0) row_shape[row] = shape[t];
}

// Shape id and run table of every record.  Runs are the maximal sequences of tile rows contiguous in A (as k_tile_runs
// finds them).  Where a run borders a row of ANOTHER brick tile of the same shape -- whose values are the same periodic
// row image -- the border is moved up to the next multiple of BRICK_ALIGN doubles: the tile on the left writes the
// few leading entries of its neighbour's row, the tile on the right starts at the aligned address.  Both tiles then
// store whole 128-byte lines; partially written sectors (a read-modify-write in DRAM) only remain at the edge of the
// uniform region.  (tools/micro/wbench.cu: 4.5 -> 5.2 TB/s for this store pattern on B200.)
constexpr int BRICK_ALIGN = 16;
__global__ void k_brick_set_shape(int n, int64_t n_rows, const int *__restrict__ shape, const int *__restrict__ tile_rows,
                                  const int64_t *__restrict__ rowptr, const int *__restrict__ row_shape,
                                  const int64_t *__restrict__ run_ptr, BrickRec *__restrict__ rec)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  BrickRec &R = rec[t];
  R.shape = shape[t];
  const int nr = (int)(run_ptr[t + 1] - run_ptr[t]);
  R.n_runs = nr;
  R.runs_inline = nr <= BRICK_RUNS_INLINE ? 1 : 0;
  for (int i = 0; i < BRICK_RUNS_INLINE; ++i) R.runs[i] = RowRun{0, 0, 0};
  if (!R.runs_inline) return;
  int i = -1, r0 = -1, r1 = -1;
  long long run_beg = 0, prev_end = -1;
  auto close = [&]() {
    if (i < 0) return;
    long long b = run_beg, e = prev_end;
    const int sh = shape[t];
    if (r0 > 0 && row_shape[r0 - 1] == sh && rowptr[r0 - 1] + 27 == rowptr[r0]) b = (b + BRICK_ALIGN - 1) / BRICK_ALIGN * BRICK_ALIGN;
    if (r1 + 1 < n_rows && row_shape[r1 + 1] == sh && rowptr[r1 + 2] - rowptr[r1 + 1] == 27) e = (e + BRICK_ALIGN - 1) / BRICK_ALIGN * BRICK_ALIGN;
    R.runs[i] = RowRun{b, (int)(e - b) | RUN_UNIFORM, (int)(b - run_beg)};     // soff: phase of the row image at b
  };
  for (int sl = 0; sl < BRICK_ROWS; ++sl) {
    const int row = tile_rows[(int64_t)t * BRICK_ROWS + sl];
    if (row < 0) continue;
    const long long beg = rowptr[row];
    if (prev_end != beg) { close(); ++i; run_beg = beg; r0 = row; }
    prev_end = rowptr[row + 1];
    r1 = row;
  }
  close();
}

// Entry codes of the lattice tiles with rows on their faces.  A CSR entry of such a tile is one of 27 x 27 values: the
// canonical neighbour c of a row in boundary state st (which of its 8 cells exist).  The codes st * 27 + c are laid out
// in A order, run by run, so k_fill_edge streams them: one coalesced 2-byte load, one table look-up, one coalesced store
// per entry, whatever the column order.  Runs of uniform rows need none (their values repeat with period 27).
constexpr unsigned EDGE_CODE_ZERO = 729u;      // an entry no local cell writes (value 0)
constexpr unsigned EDGE_CODE_DIR = 0x8000u;    // the row is a TianXin Dirichlet row ...
constexpr unsigned EDGE_CODE_DIAG = 0x4000u;   // ... and this is its diagonal
__global__ void k_edge_run_count(int64_t r0, int64_t n_runs, const RowRun *__restrict__ runs, int64_t *__restrict__ cnt)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r < n_runs) cnt[r] = (runs[r0 + r].n & RUN_UNIFORM) ? 0 : (int64_t)(runs[r0 + r].n & ~RUN_UNIFORM);
}
__global__ void __launch_bounds__(BRICK_ROWS) k_edge_codes(int t0, EdgeRec *__restrict__ rec, const int *__restrict__ tile_rows,
                                                           const int64_t *__restrict__ rowptr, const unsigned char *__restrict__ tile_perm,
                                                           const int64_t *__restrict__ run_ptr, const RowRun *__restrict__ runs,
                                                           const int64_t *__restrict__ eoff, const int *__restrict__ row_dir,
                                                           unsigned short *__restrict__ code)
{
  const int t = t0 + blockIdx.x, tid = threadIdx.x;
  EdgeRec &R = rec[blockIdx.x];
  if (tid == 0) {
    const int64_t g0 = run_ptr[t] - run_ptr[t0], g1 = run_ptr[t + 1] - run_ptr[t0];
    R.n_runs = (int)(g1 - g0);
    R.run_beg = run_ptr[t];
    R.code_beg = eoff[g0];
    int64_t end = eoff[g0];
    for (int64_t g = g0; g < g1; ++g)
      if (!(runs[run_ptr[t0] + g].n & RUN_UNIFORM)) end = eoff[g] + runs[run_ptr[t0] + g].n;
    R.code_cnt = (int)(end - eoff[g0]);
  }
  const unsigned rp = R.rowpos[tid];
  if (rp == 0xFFFFu) return;
  const int64_t slot = (int64_t)t * BRICK_ROWS + tid;
  const int row = tile_rows[slot];
  if (row < 0) return;
  const int i = rp & 31, j = (rp >> 5) & 31, k = (rp >> 10) & 31;
  const int sx = (i == 0) ? 1 : (i == R.nxs - 1 ? 2 : 0), sy = (j == 0) ? 1 : (j == R.nys - 1 ? 2 : 0), sz = (k == 0) ? 1 : (k == R.nzs - 1 ? 2 : 0);
  const int st = sx + 3 * sy + 9 * sz;
  const int64_t base = rowptr[row];
  const int64_t rb = run_ptr[t];
  int lo = 0, hi = (int)(run_ptr[t + 1] - rb) - 1;          // the run that holds this row (runs ascend in A)
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (runs[rb + mid].beg <= base) lo = mid; else hi = mid - 1;
  }
  const RowRun rr = runs[rb + lo];
  if (rr.n & RUN_UNIFORM) return;
  unsigned short *o = code + eoff[rb + lo - run_ptr[t0]] + (base - rr.beg);
  const unsigned char *pm = tile_perm + slot * PERM_STRIDE;
  const unsigned dirf = (row_dir && row_dir[row] >= 0) ? EDGE_CODE_DIR : 0u;      // TianXin Dirichlet row (used when the BC is fused)
  for (int c = 0; c < 27; ++c)
    if (pm[c] != 0xFFu) o[pm[c]] = (unsigned short)((unsigned)(st * 27 + c) | dirf | ((dirf && c == 13) ? EDGE_CODE_DIAG : 0u));
}
__global__ void k_fill_u16(int64_t n, unsigned short v, unsigned short *__restrict__ out)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = v;
}
// (Re)writes the entry codes; called at setup and whenever the set of Dirichlet rows changes.
int edge_codes_refresh(txasm_handle h)
{
  Tiles *T = h->tiles;
  if (!T || !T->d_edge_code || T->n_edge <= T->n_uni) return TXASM_OK;
  k_fill_u16<<<1184, 256, 0, h->stream>>>(T->n_edge_code, (unsigned short)EDGE_CODE_ZERO, T->d_edge_code);
  k_edge_codes<<<T->n_edge - T->n_uni, BRICK_ROWS, 0, h->stream>>>(T->n_uni, (EdgeRec *)T->d_edge_rec, T->d_tile_rows, h->d_rowptr, T->d_tile_perm,
                                                                   T->d_run_ptr, T->d_runs, T->d_edge_eoff, h->d_row_dir, T->d_edge_code);
  TX_CUDA(h, cudaGetLastError());
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  return TXASM_OK;
}

static double brick_tol(txasm_handle h) { return h->cfg.affine_tol == 0.0 ? 1e-13 : h->cfg.affine_tol; }

void brick_free(txasm_handle h)
{
  Tiles *T = h->tiles;
  if (!T) return;
  if (T->d_brick_rec) { dev_free(h, T->d_brick_rec); T->d_brick_rec = nullptr; }
  if (T->d_brick_flag) { dev_free(h, T->d_brick_flag); T->d_brick_flag = nullptr; }
  if (T->d_shapes) { dev_free(h, T->d_shapes); T->d_shapes = nullptr; }
  if (T->d_edge_rec) { dev_free(h, T->d_edge_rec); T->d_edge_rec = nullptr; }
  if (T->d_edge_code) { dev_free(h, T->d_edge_code); T->d_edge_code = nullptr; }
  if (T->d_edge_eoff) { dev_free(h, T->d_edge_eoff); T->d_edge_eoff = nullptr; }
  T->n_brick = 0; T->n_shapes = 0; T->n_edge = 0;
}

// bit 3 of tile_cong: the tile is a brick tile.  Deterministic in the tile's own tables, so it can be re-run after the
// tiles have been renumbered.
int brick_classify(txasm_handle h)
{
  Tiles *T = h->tiles;
  if (!T || T->n_tiles == 0) return TXASM_OK;
  if (T->TR != BRICK_ROWS || !T->all_affine || brick_tol(h) < 0.0) return TXASM_OK;
  if (T->d_brick_flag) { dev_free(h, T->d_brick_flag); T->d_brick_flag = nullptr; }
  int rc = dev_alloc(h, &T->d_brick_flag, (size_t)T->n_tiles);
  if (rc) return rc;
  k_brick_scan<false, false><<<T->n_tiles, BRICK_ROWS, 0, h->stream>>>(0, T->n_tiles, T->d_tile_rows, T->d_tile_cell_ptr, T->d_tile_cells,
                                                                      T->d_adjl, h->d_lids, h->d_xyz, T->d_tile_kf, T->d_tile_cong,
                                                                      brick_tol(h), T->d_brick_flag, nullptr);
  k_brick_mark<<<(T->n_tiles + 255) / 256, 256, 0, h->stream>>>(T->n_tiles, T->d_brick_flag, T->d_tile_cong, 8);
  // bit 4: a lattice tile with rows on its faces (k_fill_edge)
  k_brick_scan<false, true><<<T->n_tiles, BRICK_ROWS, 0, h->stream>>>(0, T->n_tiles, T->d_tile_rows, T->d_tile_cell_ptr, T->d_tile_cells,
                                                                     T->d_adjl, h->d_lids, h->d_xyz, T->d_tile_kf, T->d_tile_cong,
                                                                     brick_tol(h), T->d_brick_flag, nullptr);
  k_brick_mark<<<(T->n_tiles + 255) / 256, 256, 0, h->stream>>>(T->n_tiles, T->d_brick_flag, T->d_tile_cong, 16);
  TX_CUDA(h, cudaGetLastError());
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  return TXASM_OK;
}

// Records of the tiles [0, n_brick) and the table of their distinct cell shapes.  Tiles whose stiffness rows agree
// within the affine tolerance share a shape (the cells of an inline mesh differ by the rounding of i*h + x0), so a
// CTA rebuilds its row image only where the mesh really changes.
int brick_build(txasm_handle h)
{
  Tiles *T = h->tiles;
  if (!T) return TXASM_OK;
  int rc;
  if (T->n_edge > T->n_uni) {               // lattice tiles with rows on their faces: records only
    const int ne = T->n_edge - T->n_uni;
    if ((rc = dev_alloc(h, &T->d_edge_rec, (size_t)ne * sizeof(EdgeRec)))) return rc;
    TX_CUDA(h, cudaMemsetAsync(T->d_edge_rec, 0, (size_t)ne * sizeof(EdgeRec), h->stream));
    k_brick_scan<true, true><<<ne, BRICK_ROWS, 0, h->stream>>>(T->n_uni, T->n_edge, T->d_tile_rows, T->d_tile_cell_ptr, T->d_tile_cells,
                                                              T->d_adjl, h->d_lids, h->d_xyz, T->d_tile_kf, T->d_tile_cong, brick_tol(h),
                                                              nullptr, T->d_edge_rec);
    TX_CUDA(h, cudaGetLastError());
    TX_CUDA(h, cudaStreamSynchronize(h->stream));
    // entry codes of their non-uniform runs
    int64_t rp2[2];
    TX_CUDA(h, copy_to_device_sync(h, &rp2[0], T->d_run_ptr + T->n_uni, sizeof(int64_t)));
    TX_CUDA(h, copy_to_device_sync(h, &rp2[1], T->d_run_ptr + T->n_edge, sizeof(int64_t)));
    const int64_t nre = rp2[1] - rp2[0];
    T->edge_run0 = rp2[0];
    if ((rc = dev_alloc(h, &T->d_edge_eoff, (size_t)(nre + 1)))) return rc;
    k_edge_run_count<<<(unsigned)((nre + 255) / 256), 256, 0, h->stream>>>(rp2[0], nre, T->d_runs, T->d_edge_eoff);
    std::vector<int64_t> eo((size_t)nre + 1);
    TX_CUDA(h, copy_to_device_sync(h, eo.data(), T->d_edge_eoff, sizeof(int64_t) * (size_t)nre));
    std::vector<int64_t> rp((size_t)ne + 1);
    TX_CUDA(h, copy_to_device_sync(h, rp.data(), T->d_run_ptr + T->n_uni, sizeof(int64_t) * (size_t)(ne + 1)));
    int64_t acc = 0;
    T->edge_ok = true;
    for (int t = 0; t < ne; ++t) {            // a tile's codes start on a 16-byte boundary (they are staged by 16-byte loads)
      acc = (acc + 7) / 8 * 8;
      const int64_t a0 = acc;
      for (int64_t r = rp[(size_t)t] - rp2[0]; r < rp[(size_t)t + 1] - rp2[0]; ++r) { const int64_t c = eo[(size_t)r]; eo[(size_t)r] = acc; acc += c; }
      // what k_fill_edge stages per tile must fit its shared memory (rows longer than the 27-point stencil -- columns
      // inserted with txasm_graph_merge_columns -- can exceed it): such a handle keeps k_fill_rowtile on these tiles
      if (acc - a0 > EDGE_CODE_CAP - 8 || rp[(size_t)t + 1] - rp[(size_t)t] > BRICK_ROWS) T->edge_ok = false;
    }
    if (acc / 8 >= (int64_t)0xFFFFFFFFll || nre >= (int64_t)0xFFFFFFFFll) T->edge_ok = false;    // (32-bit offsets in the prefetch registers)
    eo[(size_t)nre] = acc;
    acc += 8;                                 // (the staging loop reads whole 16-byte words)
    TX_CUDA(h, copy_to_device_sync(h, T->d_edge_eoff, eo.data(), sizeof(int64_t) * (size_t)(nre + 1)));
    if ((rc = dev_alloc(h, &T->d_edge_code, (size_t)std::max<int64_t>(acc, 1)))) return rc;
    T->n_edge_code = std::max<int64_t>(acc, 1);
    if ((rc = edge_codes_refresh(h))) return rc;
  }
  if (T->n_brick == 0) return TXASM_OK;
  const int nb = T->n_brick;
  if ((rc = dev_alloc(h, &T->d_brick_rec, (size_t)nb * sizeof(BrickRec)))) return rc;
  k_brick_scan<true, false><<<nb, BRICK_ROWS, 0, h->stream>>>(0, nb, T->d_tile_rows, T->d_tile_cell_ptr, T->d_tile_cells, T->d_adjl,
                                                             h->d_lids, h->d_xyz, T->d_tile_kf, T->d_tile_cong, brick_tol(h), nullptr,
                                                             T->d_brick_rec);
  TX_CUDA(h, cudaGetLastError());
  std::vector<double> kf((size_t)nb * KF_STRIDE);
  TX_CUDA(h, copy_to_device_sync(h, kf.data(), T->d_tile_kf, sizeof(double) * kf.size()));
  const double tol = brick_tol(h);
  std::vector<int> shape(nb), reps;                 // reps: tile whose Kf row represents the shape
  auto same = [&](int a, int b) {
    const double *p = &kf[(size_t)a * KF_STRIDE], *q = &kf[(size_t)b * KF_STRIDE];
    double m = 0.0;
    for (int i = 0; i < 31; ++i) m = std::max(m, std::fabs(q[i]));
    for (int i = 0; i < 27; ++i) if (std::fabs(p[i] - q[i]) > tol * m) return false;
    for (int i = 27; i < 31; ++i) if (std::fabs(p[i] - q[i]) > tol * std::fabs(q[i])) return false;
    return true;
  };
  int last = -1;
  for (int t = 0; t < nb; ++t) {
    int s = -1;
    if (last >= 0 && same(t, reps[last])) s = last;
    for (int k = (int)reps.size() - 1, tried = 0; s < 0 && k >= 0 && tried < 256; --k, ++tried)
      if (same(t, reps[k])) s = k;
    if (s < 0) { reps.push_back(t); s = (int)reps.size() - 1; }
    shape[t] = s; last = s;
  }
  T->n_shapes = (int)reps.size();
  std::vector<double> sh((size_t)T->n_shapes * SHAPE_STRIDE);
  for (int s = 0; s < T->n_shapes; ++s)
    std::copy(&kf[(size_t)reps[s] * KF_STRIDE], &kf[(size_t)reps[s] * KF_STRIDE] + KF_STRIDE, &sh[(size_t)s * SHAPE_STRIDE]);
  rc = dev_alloc(h, &T->d_shapes, sh.size());
  if (rc) return rc;
  TX_CUDA(h, copy_to_device_sync(h, T->d_shapes, sh.data(), sizeof(double) * sh.size()));
  int *d_shape = nullptr;
  TX_CUDA(h, cudaMalloc(&d_shape, sizeof(int) * nb));
  cudaError_t e = copy_to_device_sync(h, d_shape, shape.data(), sizeof(int) * nb);
  if (e == cudaSuccess) {
    int *d_row_shape = nullptr;
    e = cudaMalloc(&d_row_shape, sizeof(int) * (size_t)h->n_rows);
    if (e == cudaSuccess) {
      cudaMemsetAsync(d_row_shape, 0xFF, sizeof(int) * (size_t)h->n_rows, h->stream);
      k_brick_row_shape<<<nb, BRICK_ROWS, 0, h->stream>>>(nb, d_shape, T->d_run_ptr, T->d_tile_rows, d_row_shape);
      k_brick_set_shape<<<(nb + 127) / 128, 128, 0, h->stream>>>(nb, h->n_rows, d_shape, T->d_tile_rows, h->d_rowptr, d_row_shape,
                                                                T->d_run_ptr, (BrickRec *)T->d_brick_rec);
      e = cudaStreamSynchronize(h->stream);
      cudaFree(d_row_shape);
    }
  }
  cudaFree(d_shape);
  TX_CUDA(h, e);
  return TXASM_OK;
}

// ============================================================================ the kernel
struct BrickArgs {
  const BrickRec *rec;
  int n_tiles;
  const double *shapes;
  const int64_t *run_ptr;
  const RowRun *runs;
  int ablate;                           // profiling only (TXASM_BRICK_ABLATE): 1 no gathers, 2 no phase 2, 4 no stores
};

constexpr int BRICK_U = 608;            // lattice buffer (doubles), >= BRICK_NODE_CAP, 16-byte multiple
constexpr int BRICK_S1D = 48;           // 1-D source factors of the lattice lines: X[16] Y[16] Z[16]
constexpr int BRICK_NBUF = 4;           // record buffers (3 in use at prefetch depth 1, 4 at depth 2)
constexpr int BRICK_GPT = (BRICK_NODE_CAP + BRICK_ROWS - 1) / BRICK_ROWS;   // gathers per thread (3)

// node mass stencil of a box, per unit det: m(dx) m(dy) m(dz) with m(0) = 4/3, m(+-1) = 1/3 -- the sum of
// aff_mass(a, b) = prod_d (1 + p_d/3)/2 over the cells around the node (two cells share the node along an axis)
__device__ __forceinline__ double brick_mass(int c)
{
  const int dx = c % 3, dy = (c / 3) % 3, dz = c / 9;
  return ((dx == 1) ? 4.0 / 3.0 : 1.0 / 3.0) * ((dy == 1) ? 4.0 / 3.0 : 1.0 / 3.0) * ((dz == 1) ? 4.0 / 3.0 : 1.0 / 3.0);
}

template <bool MASS>
__host__ __device__ constexpr int brick_smem()
{
  return BRICK_NBUF * (int)sizeof(BrickRec) + 2 * BRICK_U * 8 * (MASS ? 2 : 1) + 2 * BRICK_S1D * 8 + (IMG_DOUBLES + 28) * 8 +
         3 * 32 * 8 + 16;
}

// The gathers of tile t + G are issued while tile t is in its stencil / store phases and land in registers; they are
// written to the lattice buffer at the top of the next iteration.  So the only global-memory latency a tile waits
// for is the one it could not hide behind the previous tile.
template <bool MASS>
struct BrickPrefetch {
  double g[BRICK_GPT], m[MASS ? BRICK_GPT : 1];
  double xl, xm;                        // lattice-line coordinates of this thread's 1-D source factor
};

template <bool MASS>
__device__ __forceinline__ void brick_gather(const FillArgs &A, const BrickRec *rec, int tid, bool has_src, int ablate,
                                             BrickPrefetch<MASS> &pf)
{
  const int nn = rec->n_nodes;
#pragma unroll
  for (int r = 0; r < BRICK_GPT; ++r) {
    const int n = tid + r * BRICK_ROWS;
    double g = 0.0, m = 0.0;
    if (n < nn && !(ablate & 1)) {
      const int lid = rec->nodes[n];
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        if (A.c.kg[v] != 0.0 || (MASS && A.c.km[v] != 0.0)) {
          const double xv = __ldg(A.x[v] + lid);
          g = fma(A.c.kg[v], xv, g);
          if (MASS) m = fma(A.c.km[v], xv, m);
        }
      }
    }
    pf.g[r] = g;
    if (MASS) pf.m[r] = m;
  }
  pf.xl = pf.xm = 0.0;
  if (has_src) {
    const int nxs = rec->nxs, nys = rec->nys, nx2 = nxs - 2, ny2 = nys - 2, nz2 = rec->nzs - 2;
    const int q = tid - 128;
    if (q >= 0 && q < nx2 + ny2 + nz2) {
      const int d = (q < nx2) ? 0 : (q < nx2 + ny2 ? 1 : 2);
      const int i = 1 + ((d == 0) ? q : (d == 1 ? q - nx2 : q - nx2 - ny2));
      const int stride = (d == 0) ? 1 : (d == 1 ? nxs : nxs * nys);
      pf.xl = __ldg(A.xyz + (int64_t)rec->nodes[(i - 1) * stride] * 3 + d);
      pf.xm = __ldg(A.xyz + (int64_t)rec->nodes[i * stride] * 3 + d);
    }
  }
}

// DEPTH: how many tiles ahead the gathers run (1: 64 registers, 4 CTAs/SM; 2: one more register set, 3 CTAs/SM)
template <bool MASS, int DEPTH>
__global__ void __launch_bounds__(BRICK_ROWS, (MASS || DEPTH > 1) ? 3 : 4) k_fill_brick(FillArgs A, BrickArgs B)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BrickRec *recb = reinterpret_cast<BrickRec *>(smem_raw);
  double *ub = reinterpret_cast<double *>(smem_raw + BRICK_NBUF * sizeof(BrickRec));      // [2][BRICK_U] (x2 with MASS)
  double *s1d = ub + 2 * BRICK_U * (MASS ? 2 : 1);                                         // [2][BRICK_S1D]
  double *img = s1d + 2 * BRICK_S1D;                                                       // row image, see k_tile_runs
  double *kf = img + IMG_DOUBLES + 28;                                                     // Kf[27] .. geo at 27..30
  double *mf = kf + 32;                                                                    // Mf[27]
  double *cst = mf + 32;                                                                   // cS, cC
  const unsigned rec_s = (unsigned)__cvta_generic_to_shared(recb);
  const unsigned img_s = (unsigned)__cvta_generic_to_shared(img);
  const int tid = threadIdx.x, G = gridDim.x;
  const bool has_src = A.c.n_src > 0;
  const bool jac = A.jacobian && A.A;

  // Tile records travel by cp.async (LDGSTS, 16 bytes per thread): the TMA queue of the SM is kept for the row
  // stores, behind which a bulk load of the record would wait.
  constexpr int REC_CHUNKS = (int)sizeof(BrickRec) / 16;
  auto rec_fetch = [&](int tile, int buf) {
    if (tid < REC_CHUNKS && tile < B.n_tiles)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(rec_s + (unsigned)(buf * sizeof(BrickRec)) + 16u * tid),
                   "l"(reinterpret_cast<const unsigned char *>(B.rec + tile) + 16 * tid) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int t = blockIdx.x;
  tx_stamp(A.dbg, 0, false);
  for (int d = 0; d <= DEPTH; ++d) rec_fetch(t + d * G, d);
  asm volatile("cp.async.wait_group 1;" ::: "memory");     // all but the last record requested
  __syncthreads();
  int cur_shape = -1;
  double hx = 0.0, hy = 0.0, hz = 0.0;
  BrickPrefetch<MASS> pf, pf2;
  brick_gather<MASS>(A, recb, tid, has_src, B.ablate, pf);
  if (DEPTH > 1 && t + G < B.n_tiles) brick_gather<MASS>(A, recb + 1, tid, has_src, B.ablate, pf2);

  for (int it = 0; t < B.n_tiles; t += G, ++it) {
    const int rb_i = it % BRICK_NBUF, ub_i = it & 1;
    const BrickRec *rec = recb + rb_i;
    double *u = ub + ub_i * BRICK_U, *um = ub + (2 + ub_i) * BRICK_U, *s1 = s1d + ub_i * BRICK_S1D;
    const int nxs = rec->nxs, nys = rec->nys, nn = rec->n_nodes, shape = rec->shape;
    const int nrun = jac ? rec->n_runs : 0;
    const bool runs_in = rec->runs_inline != 0;
    int64_t rb = 0;
    if (jac && !runs_in) rb = B.run_ptr[t];

    if (shape != cur_shape) {            // first tile of the CTA, or the cell shape changes: row image, stencils, box
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // stores still reading the old image
      __syncthreads();
      const double *sp = B.shapes + (int64_t)shape * SHAPE_STRIDE;
      const double det = __ldg(sp + 30);
      if (tid < 27) { kf[tid] = __ldg(sp + tid); mf[tid] = det * brick_mass(tid); }
      if (tid >= 27 && tid < 31) kf[tid] = __ldg(sp + tid);
      if (tid == 32) {
        double cs = 0.0, cc = 0.0;
        for (int s = 0; s < A.c.n_src; ++s) {
          if (A.c.src_id[s] == TXASM_SOURCE_SIN3) cs += A.c.src_mult[s] * 118.43525281307230 * det;
          else cc += A.c.src_mult[s] * det * 8.0;
        }
        cst[0] = cs; cst[1] = cc;
      }
      if (jac)
        for (int i = tid; i < IMG_DOUBLES; i += BRICK_ROWS) {
          const int c = i % 27;
          double v = A.c.cK * __ldg(sp + c);
          if (MASS) v = fma(A.c.cM, det * brick_mass(c), v);
          img[i] = v;
        }
      hx = __ldg(sp + 27); hy = __ldg(sp + 28); hz = __ldg(sp + 29);
      cur_shape = shape;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      // (the phase-1 barrier below publishes all of it)
    }

    // ---------------- phase 1: the prefetched solution values onto the lattice; 1-D source factors of the lattice lines
#pragma unroll
    for (int r = 0; r < BRICK_GPT; ++r) {
      const int n = tid + r * BRICK_ROWS;
      if (n < nn) { u[n] = pf.g[r]; if (MASS) um[n] = pf.m[r]; }
    }
    if (has_src) {
      // interior lattice coordinate i of axis d: the node is the + vertex of the cell on its left and the - vertex of
      // the cell on its right; their Gauss points are (left node + h) -+ h/sqrt3 and (node + h) -+ h/sqrt3
      const int nx2 = nxs - 2, ny2 = nys - 2, nz2 = rec->nzs - 2;
      const int q = tid - 128;
      if (q >= 0 && q < nx2 + ny2 + nz2) {
        const int d = (q < nx2) ? 0 : (q < nx2 + ny2 ? 1 : 2);
        const int i = 1 + ((d == 0) ? q : (d == 1 ? q - nx2 : q - nx2 - ny2));
        const double hd = (d == 0) ? hx : (d == 1 ? hy : hz);
        constexpr double wl = 0.5 * (1.0 - TX_INV_SQRT3), wh = 0.5 * (1.0 + TX_INV_SQRT3);
        const double dq = hd * TX_INV_SQRT3;
        const double f0 = sin2pi_fast(pf.xl + hd - dq), f1 = sin2pi_fast(pf.xl + hd + dq);
        const double g0 = sin2pi_fast(pf.xm + hd - dq), g1 = sin2pi_fast(pf.xm + hd + dq);
        s1[d * 16 + i] = (wl * f0 + wh * f1) + (wh * g0 + wl * g1);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");     // my piece of the next tile's record (requested a tile ago)
    __syncthreads();                     // lattice and next record complete; every thread is past tile t - G
    rec_fetch(t + (DEPTH + 1) * G, (it + DEPTH + 1) % BRICK_NBUF);
    const int my_run = (tid & 31) * (BRICK_ROWS / 32) + (tid >> 5);
    RowRun rr0{0, 0, 0};
    if (jac && my_run < nrun) rr0 = runs_in ? rec->runs[my_run] : B.runs[rb + my_run];
    // the gathers of the next tile: its record arrived a tile ago; the loads fly under this tile's phases 2 and 3
    if (DEPTH > 1) {
      pf = pf2;
      if (t + 2 * G < B.n_tiles) brick_gather<MASS>(A, recb + (it + 2) % BRICK_NBUF, tid, has_src, B.ablate, pf2);
    } else if (t + G < B.n_tiles) brick_gather<MASS>(A, recb + (it + 1) % BRICK_NBUF, tid, has_src, B.ablate, pf);

    // ---------------- phase 2: f = sum_j Kf[j] u[node + off_j] (+ Mf[j] um[...]) + source
    const unsigned rp = rec->rowpos[tid];
    if (rp != 0xFFFFu && A.f && !(B.ablate & 2)) {
      const int i = rp & 15, j = (rp >> 4) & 15, k = rp >> 8;
      const int c = i + nxs * (j + nys * k), sy = nxs, sz = nxs * nys;
      double fr = 0.0;
#pragma unroll
      for (int dz = 0; dz < 3; ++dz)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int o = c + (dx - 1) + (dy - 1) * sy + (dz - 1) * sz;
            fr = fma(kf[dx + 3 * dy + 9 * dz], u[o], fr);
            if (MASS) fr = fma(mf[dx + 3 * dy + 9 * dz], um[o], fr);
          }
      if (has_src) fr += fma(cst[0] * s1[i], s1[16 + j] * s1[32 + k], cst[1]);
      A.f[rec->nodes[c]] = fr;
    }

    // ---------------- A: every run straight from the constant image.  Element k of a run is img[(phase + k) % 27];
    //                  the image holds that pattern for 27 + 26 + 216 elements, so a copy of up to 8 rows starts at
    //                  img + phase (phase even) or img + phase + 27 (phase odd): both 16-byte aligned.
    if (jac && !(B.ablate & 4)) {
      for (int r = my_run; r < nrun; r += BRICK_ROWS) {
        RowRun rr = (r == my_run) ? rr0 : B.runs[rb + r];
        rr.n &= ~RUN_UNIFORM;
        int ph = runs_in ? rr.soff : 0;
        double *g = A.A + rr.beg;
        const int head = (int)(rr.beg & 1);
        const int mid = (rr.n - head) & ~1;
        if (head) g[0] = img[ph];
        if (rr.n - head - mid) g[rr.n - 1] = img[(ph + rr.n - 1) % 27];
        ph = (ph + head) % 27;
        const unsigned src = img_s + 8u * (unsigned)(ph + ((ph & 1) ? 27 : 0));
        for (int o = 0; o < mid; o += 216) {
          const int m = (mid - o < 216) ? mid - o : 216;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       ::"l"(g + head + o), "r"(src), "r"((unsigned)m * 8u) : "memory");
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // stores may still be reading the image
  tx_stamp(A.dbg, 0, true);
}

// ============================================================================ lattice tiles with rows on their faces
// The tiles k_fill_brick cannot take because some rows sit on a face of the lattice -- the boundary of the mesh or of the
// rank's brick, where cells are missing -- but whose cells are still congruent, axis-aligned and fill a box (every
// boundary tile of an inline mesh).  For a box grid the element integrals factorise per axis, so the row of node p is
//     A[p, p + o] = kx(ox) my(oy) mz(oz) + mx(ox) ky(oy) mz(oz) + mx(ox) my(oy) kz(oz),     M[p, p + o] = mx my mz
// with the 1-D stiffness / mass factors of a node that has a cell on its left (sL) and / or right (sR), cell length L:
//     k(0) = (sL + sR) / L,  k(-1) = -sL / L,  k(+1) = -sR / L;     m(0) = (sL + sR) L / 3,  m(-1) = sL L / 6,  m(+1) = sR L / 6
// -- the sums over the cells around p of the exact integrals of elem_q1hex.cuh (the same numbers k_fill_rowtile accumulates
// cell by cell, in another order).  One CTA per tile, one thread per row: gather once per lattice node, 27 entries
// computed and stored through the row's canonical -> CSR-slot permutation (any column order; slots no local cell writes
// are zeroed; TianXin Dirichlet rows become identity rows).  Small on purpose (64 registers, 12 KB shared memory, CTAs
// that come and go): it runs on a side stream BESIDE the persistent k_fill_brick instead of after it.
struct EdgeArgs {
  const EdgeRec *rec;
  int t0, n_tiles;                      // tiles [t0, n_tiles)
  const double *tile_kf;                // [tile][KF_STRIDE]: .. | Jxx Jyy Jzz det at 27..30
  const int *tile_rows;                 // [tile * 256 + slot] row id or -1
  const RowRun *runs;
  const unsigned short *code;           // entry codes of the non-uniform runs (k_edge_codes)
  const int64_t *eoff;                  // [run - run0] first code of the run
  int64_t run0;                         // run_ptr[t0]
  int tma_store;                        // A is 16-byte aligned: runs of uniform rows leave from the constant image
  const int *row_dir;                   // fused Dirichlet (or NULL)
  const double *dir_vals;
  double tol;                           // cell sizes that agree within tol (relative) share the value tables
  int ablate;                           // profiling only (TXASM_EDGE_ABLATE): 1 no gathers, 2 no code stores, 4 no TMA stores, 8 no f
};

constexpr int EDGE_U = EDGE_NODE_CAP;
// 1-D assembled stiffness (k) and mass (m) factor of a lattice node towards its neighbour at offset o - 1 along an axis of
// cell length L; st: 0 cells on both sides, 1 only on the right (low face), 2 only on the left (high face).
__device__ __forceinline__ void edge_fac(double L, int st, int o, double &k, double &m)
{
  const double sL = (st == 1) ? 0.0 : 1.0, sR = (st == 2) ? 0.0 : 1.0;
  k = (o == 1) ? (sL + sR) / L : -((o == 0) ? sL : sR) / L;
  m = (o == 1) ? (sL + sR) * L * (1.0 / 3.0) : ((o == 0) ? sL : sR) * L * (1.0 / 6.0);
}

template <bool MASS>
__host__ __device__ constexpr int edge_smem()
{
  return 8 * ((IMG_DOUBLES + 28) + 27 * 27 * (MASS ? 2 : 1) + (27 * 27 + 1) + EDGE_U * (MASS ? 2 : 1) + 3 * EDGE_DIM_CAP + BRICK_ROWS) +
         4 * 2 * BRICK_ROWS + 2 * EDGE_CODE_CAP + 16;
}

// What a CTA fetches one tile ahead (registers): the header of the record, the lattice nodes this thread gathers, its row.
struct EdgePre {
  unsigned dimsrun;                     // nxs | nys << 6 | nzs << 12 | n_runs << 18
  unsigned nncode;                      // n_nodes | code_cnt << 11
  unsigned rb, cb8;                     // first run (relative to run0), first code / 8
  int node[EDGE_NODE_CAP / BRICK_ROWS];
  int row;
  unsigned rp;
};
__device__ __forceinline__ void edge_prefetch(const EdgeArgs &E, int t, int tid, EdgePre &P)
{
  if (t >= E.n_tiles) { P.dimsrun = 0; P.nncode = 0; P.rb = P.cb8 = 0; P.row = -1; P.rp = 0xFFFFu; return; }
  const EdgeRec *rec = E.rec + (t - E.t0);
  const int4 h0 = __ldg(reinterpret_cast<const int4 *>(rec)), h1 = __ldg(reinterpret_cast<const int4 *>(rec) + 1);
  const longlong2 h2 = __ldg(reinterpret_cast<const longlong2 *>(rec) + 2);
  P.dimsrun = (unsigned)h0.x | ((unsigned)h0.y << 6) | ((unsigned)h0.z << 12) | ((unsigned)h1.z << 18);
  P.nncode = (unsigned)h1.y | ((unsigned)h1.w << 11);
  P.rb = (unsigned)(h2.x - E.run0); P.cb8 = (unsigned)(h2.y >> 3);
#pragma unroll
  for (int q = 0; q < EDGE_NODE_CAP / BRICK_ROWS; ++q) P.node[q] = __ldg(rec->nodes + tid + q * BRICK_ROWS);   // (padded with 0 beyond n_nodes)
  P.rp = __ldg(rec->rowpos + tid);
  P.row = __ldg(E.tile_rows + (int64_t)t * BRICK_ROWS + tid);
}

template <bool MASS>
__global__ void __launch_bounds__(BRICK_ROWS, 4) k_fill_edge(FillArgs A, EdgeArgs E)
{
  extern __shared__ __align__(16) unsigned char edge_smem_raw[];
  double *img = reinterpret_cast<double *>(edge_smem_raw);      // [IMG_DOUBLES + 28] row image of the INTERIOR rows (see k_fill_uniform)
  double *sK = img + IMG_DOUBLES + 28;                          // [state 0..26 = sx + 3 sy + 9 sz][canonical neighbour]: stiffness
  double *sM = sK + 27 * 27;                                    // ... mass (MASS only)
  double *sV = sM + (MASS ? 27 * 27 : 0);                       // cK * sK + cM * sM: the Jacobian entry of a code; [729] = 0
  double *u = sV + 27 * 27 + 1;                                 // [EDGE_U] lattice: gathered solution (stiffness weights)
  double *um = u + EDGE_U;                                      // ... mass weights (MASS only)
  double *s1 = um + (MASS ? EDGE_U : 0);                        // [3][EDGE_DIM_CAP] 1-D source factors
  long long *sRunBeg = reinterpret_cast<long long *>(s1 + 3 * EDGE_DIM_CAP);   // per run: first A index
  int *sRunN = reinterpret_cast<int *>(sRunBeg + BRICK_ROWS);   // per run: length | RUN_UNIFORM
  int *sRunOff = sRunN + BRICK_ROWS;                            // per run: first code in sCode
  unsigned short *sCode = reinterpret_cast<unsigned short *>((reinterpret_cast<uintptr_t>(sRunOff + BRICK_ROWS) + 15) & ~(uintptr_t)15);   // [EDGE_CODE_CAP]
  const int tid = threadIdx.x;
  tx_stamp(A.dbg, 1, false);
  const bool has_src = A.c.n_src > 0;
  const bool jac = A.jacobian && A.A;
  const bool use_img = jac && E.tma_store;     // runs of uniform (interior, canonical order) rows: from the image, by TMA
  const bool fused = E.row_dir != nullptr;     // the Dirichlet rows flagged in the codes become identity rows
  double th0 = -1.0, th1 = -1.0, th2 = -1.0;   // cell size the tables in shared memory were made for
  EdgePre P;
  edge_prefetch(E, E.t0 + (int)blockIdx.x, tid, P);
  for (int t = E.t0 + (int)blockIdx.x; t < E.n_tiles; t += (int)gridDim.x) {     // persistent CTAs
    const EdgeRec *rec = E.rec + (t - E.t0);
    const int nxs = P.dimsrun & 63, nys = (P.dimsrun >> 6) & 63, nzs = (P.dimsrun >> 12) & 63, nrun = (int)(P.dimsrun >> 18);
    const int nn = (int)(P.nncode & 2047u), ncode = (int)(P.nncode >> 11);
    const int64_t rb = E.run0 + P.rb, cb = (int64_t)P.cb8 << 3;
    const unsigned rp = P.rp;
    const int row = P.row;
    // ---- phase 1: everything the tile reads from global memory, issued together (their addresses came one tile ahead):
    //      the solution once per lattice node, the runs and their entry codes, the Dirichlet index of the thread's row
    double g[EDGE_NODE_CAP / BRICK_ROWS], gm[MASS ? EDGE_NODE_CAP / BRICK_ROWS : 1];
#pragma unroll
    for (int q = 0; q < EDGE_NODE_CAP / BRICK_ROWS; ++q) {
      g[q] = 0.0;
      if (MASS) gm[q] = 0.0;
      if (tid + q * BRICK_ROWS < nn && !(E.ablate & 1)) {
#pragma unroll
        for (int v = 0; v < 3; ++v)
          if (A.c.kg[v] != 0.0 || (MASS && A.c.km[v] != 0.0)) {
            const double xv = __ldg(A.x[v] + P.node[q]);
            g[q] = fma(A.c.kg[v], xv, g[q]);
            if (MASS) gm[q] = fma(A.c.km[v], xv, gm[q]);
          }
      }
    }
    const int dir_i = (fused && row >= 0) ? __ldg(E.row_dir + row) : -1;
    if (jac) {
      for (int i = tid; i * 8 < ncode; i += BRICK_ROWS)
        reinterpret_cast<uint4 *>(sCode)[i] = __ldg(reinterpret_cast<const uint4 *>(E.code + cb) + i);
      for (int r = tid; r < nrun; r += BRICK_ROWS) {
        const RowRun rr = E.runs[rb + r];
        sRunBeg[r] = rr.beg; sRunN[r] = rr.n;
        sRunOff[r] = (rr.n & RUN_UNIFORM) ? 0 : (int)(__ldg(E.eoff + (rb - E.run0) + r) - cb);
      }
    }
    const double *geo = E.tile_kf + (int64_t)t * KF_STRIDE + 27;
    const double hx = __ldg(geo), hy = __ldg(geo + 1), hz = __ldg(geo + 2), det = __ldg(geo + 3);
    // (the cells of an inline mesh differ by the rounding of i * h + x0: same tables within the affine tolerance)
    const bool retab = !(fabs(hx - th0) <= 4.0 * E.tol * hx && fabs(hy - th1) <= 4.0 * E.tol * hy && fabs(hz - th2) <= 4.0 * E.tol * hz);
    if (retab) { th0 = hx; th1 = hy; th2 = hz; }
    if (retab) {                               // the stencil of every boundary state
      for (int e = tid; e < 27 * 27; e += BRICK_ROWS) {
        const int st = e / 27, c = e - st * 27;
        double kx, mx, ky, my, kz, mz;
        edge_fac(2.0 * hx, st % 3, c % 3, kx, mx);
        edge_fac(2.0 * hy, (st / 3) % 3, (c / 3) % 3, ky, my);
        edge_fac(2.0 * hz, st / 9, c / 9, kz, mz);
        const double kk = kx * my * mz + mx * ky * mz + mx * my * kz, mm = mx * my * mz;
        sK[e] = kk;
        if (MASS) sM[e] = mm;
        sV[e] = MASS ? fma(A.c.cK, kk, A.c.cM * mm) : A.c.cK * kk;
      }
      if (tid == 0) sV[27 * 27] = 0.0;
      if (use_img)                            // (state 0 = a node with all 8 cells; same formula as sV above)
        for (int i = tid; i < IMG_DOUBLES; i += BRICK_ROWS) {
          const int c = i % 27;
          double kx, mx, ky, my, kz, mz;
          edge_fac(2.0 * hx, 0, c % 3, kx, mx);
          edge_fac(2.0 * hy, 0, (c / 3) % 3, ky, my);
          edge_fac(2.0 * hz, 0, c / 9, kz, mz);
          const double kk = kx * my * mz + mx * ky * mz + mx * my * kz, mm = mx * my * mz;
          img[i] = MASS ? fma(A.c.cK, kk, A.c.cM * mm) : A.c.cK * kk;
        }
    }
    if (has_src && tid < nxs + nys + nzs) {    // 1-D source factors (one-sided at the faces)
      const int d = (tid < nxs) ? 0 : (tid < nxs + nys ? 1 : 2);
      const int i = (d == 0) ? tid : (d == 1 ? tid - nxs : tid - nxs - nys);
      const int nd = (d == 0) ? nxs : (d == 1 ? nys : nzs), stride = (d == 0) ? 1 : (d == 1 ? nxs : nxs * nys);
      const double hd = (d == 0) ? hx : (d == 1 ? hy : hz);
      constexpr double wl = 0.5 * (1.0 - TX_INV_SQRT3), wh = 0.5 * (1.0 + TX_INV_SQRT3);
      const double dq = hd * TX_INV_SQRT3;
      const double xm = __ldg(A.xyz + (int64_t)rec->nodes[i * stride] * 3 + d);
      double v = 0.0;
      if (i > 0) {                           // + vertex of the cell on the left
        const double xl = __ldg(A.xyz + (int64_t)rec->nodes[(i - 1) * stride] * 3 + d);
        v += wl * sin2pi_fast(xl + hd - dq) + wh * sin2pi_fast(xl + hd + dq);
      }
      if (i < nd - 1) v += wh * sin2pi_fast(xm + hd - dq) + wl * sin2pi_fast(xm + hd + dq);   // - vertex of the cell on the right
      s1[d * EDGE_DIM_CAP + i] = v;
    }
#pragma unroll
    for (int q = 0; q < EDGE_NODE_CAP / BRICK_ROWS; ++q)
      if (tid + q * BRICK_ROWS < nn) {
        u[tid + q * BRICK_ROWS] = g[q];
        if (MASS) um[tid + q * BRICK_ROWS] = gm[q];
      }
    edge_prefetch(E, t + (int)gridDim.x, tid, P);          // the next tile's addresses (in flight during the stores below)
    if (retab && use_img) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- the Jacobian, run by run (a run = rows of the tile that are contiguous in A)
    if (jac) {
      if (use_img) {                          // runs of uniform rows: TMA bulk stores from the image (as in k_fill_uniform)
        const unsigned img_s = (unsigned)__cvta_generic_to_shared(img);
        for (int r = tid; r < nrun; r += BRICK_ROWS) {
          int n = sRunN[r];
          if (!(n & RUN_UNIFORM) || (E.ablate & 4)) continue;
          n &= ~RUN_UNIFORM;
          const long long beg = sRunBeg[r];
          double *gA = A.A + beg;
          const int head = (int)(beg & 1);
          const int mid = (n - head) & ~1;
          if (head) gA[0] = img[0];
          if (n - head - mid) gA[n - 1] = img[(n - 1) % 27];
          const unsigned src = img_s + (head ? 28u * 8u : 0u);
          for (int o = 0; o < mid; o += 216) {
            const int m = (mid - o < 216) ? mid - o : 216;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gA + head + o), "r"(src), "r"((unsigned)m * 8u) : "memory");
          }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      // the other runs: a warp per run streams the run's entry codes (shared memory) -- one look-up, one coalesced store each
      const int lane = tid & 31;
      for (int r = tid >> 5; r < nrun; r += BRICK_ROWS / 32) {
        const int nf = sRunN[r];
        double *gA = A.A + sRunBeg[r];
        if (nf & RUN_UNIFORM) {
          if (!use_img) {
            const int n = nf & ~RUN_UNIFORM;
            for (int j = lane; j < n; j += 32) gA[j] = sV[j % 27];
          }
          continue;
        }
        const unsigned short *cd = sCode + sRunOff[r];
        if (E.ablate & 2) continue;
#pragma unroll 4
        for (int j = lane; j < nf; j += 32) {
          const unsigned c = cd[j];
          double v = sV[c & 1023u];
          if (fused && (c & EDGE_CODE_DIR)) v = (c & EDGE_CODE_DIAG) ? 1.0 : 0.0;
          gA[j] = v;
        }
      }
    }
    // ---- one thread per row: f (27-point stencil of the row's state)
    if (A.f && rp != 0xFFFFu && !(E.ablate & 8)) {
      const int ri = rp & 31, rj = (rp >> 5) & 31, rk = (rp >> 10) & 31;
      const int c0 = ri + nxs * (rj + nys * rk);
      // state per axis: 0 cells on both sides, 1 only on the right (low face), 2 only on the left (high face)
      const int sx = (ri == 0) ? 1 : (ri == nxs - 1 ? 2 : 0), sy = (rj == 0) ? 1 : (rj == nys - 1 ? 2 : 0), sz = (rk == 0) ? 1 : (rk == nzs - 1 ? 2 : 0);
      const int st = sx + 3 * sy + 9 * sz;
      double fr = 0.0;
      if (dir_i >= 0) fr = __ldg(A.x[0] + row) - __ldg(E.dir_vals + dir_i);     // TianXin Dirichlet: f = x - value
      else {
        const double *kr = sK + st * 27, *mr = sM + (MASS ? st * 27 : 0);
#pragma unroll
        for (int c = 0; c < 27; ++c) {
          const int dx = c % 3, dy = (c / 3) % 3, dz = c / 9;
          const bool ok = !(dx == 0 && sx == 1) && !(dx == 2 && sx == 2) && !(dy == 0 && sy == 1) && !(dy == 2 && sy == 2) &&
                          !(dz == 0 && sz == 1) && !(dz == 2 && sz == 2);
          if (ok) {
            const int o = c0 + (dx - 1) + (dy - 1) * nxs + (dz - 1) * nxs * nys;
            fr = fma(kr[c], u[o], fr);
            if (MASS) fr = fma(mr[c], um[o], fr);
          }
        }
        if (has_src) {
          double cs = 0.0, cc = 0.0;
          for (int s = 0; s < A.c.n_src; ++s) {
            if (A.c.src_id[s] == TXASM_SOURCE_SIN3) cs += A.c.src_mult[s] * 118.43525281307230 * det;
            else cc += A.c.src_mult[s] * det;
          }
          const double ncell = ((ri > 0) + (ri < nxs - 1)) * ((rj > 0) + (rj < nys - 1)) * ((rk > 0) + (rk < nzs - 1));
          fr += fma(cs * s1[ri], s1[EDGE_DIM_CAP + rj] * s1[2 * EDGE_DIM_CAP + rk], cc * ncell);
        }
      }
      A.f[row] = fr;
    }
    if (use_img) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the image must outlive the reads of its stores
    __syncthreads();                              // shared memory is rewritten by the next tile
  }
  tx_stamp(A.dbg, 1, true);
}

bool fill_edge_eligible(txasm_handle h, const FillArgs &a)
{
  const Tiles *T = h->tiles;
  if (!h->opt_uniform || !h->opt_edge || !T || T->n_edge <= T->n_uni || !T->d_edge_rec || !T->d_edge_code || !T->edge_ok) return false;
  for (int i = 0; i < a.c.n_src; ++i)
    if (a.c.src_id[i] != TXASM_SOURCE_SIN3 && a.c.src_id[i] != TXASM_SOURCE_CONSTANT) return false;
  return true;
}

int launch_fill_edge(txasm_handle h, const FillArgs &a, cudaStream_t stream, const int *row_dir, const double *dir_vals)
{
  Tiles *T = h->tiles;
  const int tma_ok = (a.A && (((uintptr_t)a.A) & 15) == 0) ? 1 : 0;
  static const int edge_ablate = [] { const char *e = getenv("TXASM_EDGE_ABLATE"); return e ? atoi(e) : 0; }();
  EdgeArgs e{(const EdgeRec *)T->d_edge_rec, T->n_uni, T->n_edge, T->d_tile_kf, T->d_tile_rows, T->d_runs,
             T->d_edge_code, T->d_edge_eoff, T->edge_run0, tma_ok, row_dir, dir_vals, brick_tol(h), edge_ablate};
  if (!T->edge_attr_set) {
    TX_CUDA(h, cudaFuncSetAttribute(k_fill_edge<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, edge_smem<true>()));
    TX_CUDA(h, cudaFuncSetAttribute(k_fill_edge<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, edge_smem<false>()));
    T->edge_attr_set = true;
  }
  auto k = a.c.has_mass ? k_fill_edge<true> : k_fill_edge<false>;
  const int smem = a.c.has_mass ? edge_smem<true>() : edge_smem<false>();
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, BRICK_ROWS, smem);
  int grid = std::min(T->n_edge - T->n_uni, std::max(1, occ) * h->n_sm);
  if (h->opt_grid_cap > 0) grid = std::min(grid, h->opt_grid_cap);
  k<<<grid, BRICK_ROWS, smem, stream>>>(a, e);
  TX_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return TXASM_OK;
}

bool fill_brick_eligible(txasm_handle h, const FillArgs &a)
{
  const Tiles *T = h->tiles;
  if (!h->opt_uniform || !h->opt_brick || !T || T->n_brick == 0 || !T->d_brick_rec) return false;
  if (a.jacobian && (!a.A || (((uintptr_t)a.A) & 15) != 0)) return false;
  for (int i = 0; i < a.c.n_src; ++i)
    if (a.c.src_id[i] != TXASM_SOURCE_SIN3 && a.c.src_id[i] != TXASM_SOURCE_CONSTANT) return false;
  return true;
}

int launch_fill_brick(txasm_handle h, const FillArgs &a, cudaStream_t stream)
{
  Tiles *T = h->tiles;
  const bool mass = a.c.has_mass != 0;
  static const int depth = [] { const char *e = getenv("TXASM_BRICK_DEPTH"); return e ? atoi(e) : 1; }();
  auto k = mass ? k_fill_brick<true, 1> : (depth > 1 ? k_fill_brick<false, 2> : k_fill_brick<false, 1>);
  const int smem = mass ? brick_smem<true>() : brick_smem<false>();
  if (!T->brick_attr_set) {
    TX_CUDA(h, cudaFuncSetAttribute(k_fill_brick<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, brick_smem<true>()));
    TX_CUDA(h, cudaFuncSetAttribute(k_fill_brick<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, brick_smem<false>()));
    TX_CUDA(h, cudaFuncSetAttribute(k_fill_brick<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, brick_smem<false>()));
    tx_set_carveout(k_fill_brick<true, 1>); tx_set_carveout(k_fill_brick<false, 1>); tx_set_carveout(k_fill_brick<false, 2>);
    tx_set_carveout(k_fill_edge<true>); tx_set_carveout(k_fill_edge<false>);
    T->brick_attr_set = true;
  }
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, BRICK_ROWS, smem);
  if (h->opt_brick_ctas > 0) occ = std::min(occ, h->opt_brick_ctas);
  if (h->brick_ctas_limit > 0) occ = std::min(occ, h->brick_ctas_limit);
  int grid = std::min(T->n_brick, std::max(1, occ) * h->n_sm);
  if (h->opt_grid_cap > 0) grid = std::min(grid, h->opt_grid_cap);
  T->ctas_per_sm = occ;
  static const int ablate = [] { const char *e = getenv("TXASM_BRICK_ABLATE"); return e ? atoi(e) : 0; }();
  BrickArgs ba{(const BrickRec *)T->d_brick_rec, T->n_brick, T->d_shapes, T->d_run_ptr, T->d_runs, ablate};
  k<<<grid, BRICK_ROWS, smem, stream>>>(a, ba);
  TX_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return TXASM_OK;
}

}  // namespace txasm
