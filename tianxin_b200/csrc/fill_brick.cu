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
static_assert(sizeof(BrickRec) % 16 == 0, "BrickRec is moved by a TMA bulk copy");

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
template <bool WRITE>
__global__ void __launch_bounds__(BRICK_ROWS) k_brick_scan(int n_tiles, const int *__restrict__ tile_rows,
                                                           const int64_t *__restrict__ cell_ptr, const int *__restrict__ cells,
                                                           const unsigned short *__restrict__ adjl, const int *__restrict__ lids,
                                                           const double *__restrict__ xyz, const double *__restrict__ tile_kf,
                                                           const unsigned char *__restrict__ tile_cong, double tol,
                                                           unsigned char *__restrict__ flag, BrickRec *__restrict__ rec)
{
  __shared__ int nodes[BRICK_NODE_CAP];
  __shared__ int occ[BRICK_CELL_CAP];
  __shared__ unsigned short cellpos[BRICK_CELL_CAP];
  __shared__ unsigned long long xmin[3];
  __shared__ int dims[3], bad, nrows;
  const int t = blockIdx.x, tid = threadIdx.x;
  if (t >= n_tiles) return;
  const int64_t cb = cell_ptr[t];
  const int nc = (int)(cell_ptr[t + 1] - cb);
  const double h[3] = {tile_kf[(int64_t)t * KF_STRIDE + 27], tile_kf[(int64_t)t * KF_STRIDE + 28], tile_kf[(int64_t)t * KF_STRIDE + 29]};
  if (tid == 0) {
    bad = ((tile_cong[t] & 7) != 7) || nc <= 0 || nc > BRICK_CELL_CAP || !(h[0] > 0.0) || !(h[1] > 0.0) || !(h[2] > 0.0);
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
      ok = ok && q >= 0.0 && q < (double)(BRICK_DIM_CAP - 1) && fabs(x - (x0 + 2.0 * h[d] * q)) <= 4.0 * tol * hmax * (q + 1.0);
      p[d] = ok ? (int)q : 0;
      if (ok) atomicMax(&dims[d], p[d] + 1);
    }
    if (!ok) bad = 1;
    cellpos[j] = (unsigned short)(p[0] | (p[1] << 4) | (p[2] << 8));
  }
  __syncthreads();
  const int cx = dims[0], cy = dims[1], cz = dims[2];
  const int nxs = cx + 1, nys = cy + 1, nzs = cz + 1, nn = nxs * nys * nzs;
  if (bad || cx * cy * cz != nc || nn > BRICK_NODE_CAP) { if (!WRITE && tid == 0) flag[t] = 0; return; }
  for (int i = tid; i < nn; i += BRICK_ROWS) nodes[i] = -1;
  for (int i = tid; i < nc; i += BRICK_ROWS) occ[i] = 0;
  __syncthreads();
  for (int j = tid; j < nc; j += BRICK_ROWS) {
    const int cp = cellpos[j], pi = cp & 15, pj = (cp >> 4) & 15, pk = cp >> 8;
    if (atomicExch(&occ[pi + cx * (pj + cy * pk)], 1)) bad = 1;            // two cells at one lattice position
    const int *l = lids + (int64_t)cells[cb + j] * 8;
    for (int a = 0; a < 8; ++a) {
      const int idx = (pi + (hex_sx(a) > 0)) + nxs * ((pj + (hex_sy(a) > 0)) + nys * (pk + (hex_sz(a) > 0)));
      const int old = atomicCAS(&nodes[idx], -1, l[a]);
      if (old != -1 && old != l[a]) bad = 1;                                // cells disagree on the node
    }
  }
  __syncthreads();
  // rows: the row is vertex 6 (+,+,+) of the cell at (i-1, j-1, k-1)
  const int row = tile_rows[(int64_t)t * BRICK_ROWS + tid];
  unsigned short rp = 0xFFFF;
  if (row >= 0 && !bad) {
    const int el = adjl[((int64_t)t * BRICK_ROWS + tid) * 8 + 6];
    if (el == 0xFFFF || el >= nc) bad = 1;
    else {
      const int cp = cellpos[el], i = (cp & 15) + 1, j = ((cp >> 4) & 15) + 1, k = (cp >> 8) + 1;
      if (i < 1 || i > nxs - 2 || j < 1 || j > nys - 2 || k < 1 || k > nzs - 2 || nodes[i + nxs * (j + nys * k)] != row) bad = 1;
      rp = (unsigned short)(i | (j << 4) | (k << 8));
      atomicAdd(&nrows, 1);
    }
  }
  __syncthreads();
  if (!WRITE) { if (tid == 0) flag[t] = bad ? 0 : 1; return; }
  BrickRec *r = rec + t;
  if (tid == 0) {
    r->nxs = nxs; r->nys = nys; r->nzs = nzs; r->n_rows = nrows; r->shape = bad ? -1 : 0; r->n_nodes = nn;
  }
  r->rowpos[tid] = rp;
  for (int i = tid; i < BRICK_NODE_CAP; i += BRICK_ROWS) r->nodes[i] = (i < nn) ? nodes[i] : 0;
}

__global__ void k_brick_mark(int n, const unsigned char *__restrict__ flag, unsigned char *__restrict__ tile_cong)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) tile_cong[t] = (unsigned char)((tile_cong[t] & 7) | (flag[t] ? 8 : 0));
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

static double brick_tol(txasm_handle h) { return h->cfg.affine_tol == 0.0 ? 1e-13 : h->cfg.affine_tol; }

void brick_free(txasm_handle h)
{
  Tiles *T = h->tiles;
  if (!T) return;
  if (T->d_brick_rec) { dev_free(h, T->d_brick_rec); T->d_brick_rec = nullptr; }
  if (T->d_brick_flag) { dev_free(h, T->d_brick_flag); T->d_brick_flag = nullptr; }
  if (T->d_shapes) { dev_free(h, T->d_shapes); T->d_shapes = nullptr; }
  T->n_brick = 0; T->n_shapes = 0;
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
  k_brick_scan<false><<<T->n_tiles, BRICK_ROWS, 0, h->stream>>>(T->n_tiles, T->d_tile_rows, T->d_tile_cell_ptr, T->d_tile_cells,
                                                               T->d_adjl, h->d_lids, h->d_xyz, T->d_tile_kf, T->d_tile_cong,
                                                               brick_tol(h), T->d_brick_flag, nullptr);
  k_brick_mark<<<(T->n_tiles + 255) / 256, 256, 0, h->stream>>>(T->n_tiles, T->d_brick_flag, T->d_tile_cong);
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
  if (!T || T->n_brick == 0) return TXASM_OK;
  const int nb = T->n_brick;
  int rc = dev_alloc(h, &T->d_brick_rec, (size_t)nb * sizeof(BrickRec));
  if (rc) return rc;
  k_brick_scan<true><<<nb, BRICK_ROWS, 0, h->stream>>>(nb, T->d_tile_rows, T->d_tile_cell_ptr, T->d_tile_cells, T->d_adjl,
                                                      h->d_lids, h->d_xyz, T->d_tile_kf, T->d_tile_cong, brick_tol(h), nullptr,
                                                      (BrickRec *)T->d_brick_rec);
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
constexpr int BRICK_NBUF = 3;           // record buffers: the record of tile t + 2G is in flight while tile t computes
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

template <bool MASS>
__global__ void __launch_bounds__(BRICK_ROWS, MASS ? 3 : 4) k_fill_brick(FillArgs A, BrickArgs B)
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
  rec_fetch(t, 0);
  rec_fetch(t + G, 1);
  asm volatile("cp.async.wait_group 1;" ::: "memory");     // record of the first tile
  __syncthreads();
  int cur_shape = -1;
  double hx = 0.0, hy = 0.0, hz = 0.0;
  BrickPrefetch<MASS> pf;
  brick_gather<MASS>(A, recb, tid, has_src, B.ablate, pf);

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
    rec_fetch(t + 2 * G, (it + 2) % BRICK_NBUF);
    const int my_run = (tid & 31) * (BRICK_ROWS / 32) + (tid >> 5);
    RowRun rr0{0, 0, 0};
    if (jac && my_run < nrun) rr0 = runs_in ? rec->runs[my_run] : B.runs[rb + my_run];
    // the gathers of the next tile: its record arrived a tile ago; the loads fly under this tile's phases 2 and 3
    if (t + G < B.n_tiles) brick_gather<MASS>(A, recb + (it + 1) % BRICK_NBUF, tid, has_src, B.ablate, pf);

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
  auto k = mass ? k_fill_brick<true> : k_fill_brick<false>;
  const int smem = mass ? brick_smem<true>() : brick_smem<false>();
  if (!T->brick_attr_set) {
    TX_CUDA(h, cudaFuncSetAttribute(k_fill_brick<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, brick_smem<true>()));
    TX_CUDA(h, cudaFuncSetAttribute(k_fill_brick<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, brick_smem<false>()));
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
