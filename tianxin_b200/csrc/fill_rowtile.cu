// fill_rowtile.cu -- TXASM_SCATTER_ROWTILE: the B200 fast path.  Owner-computes row tiles.
//
// Replaces, for one element block, the whole workset loop of AssemblyEngine::evaluateVolume
// (disc-fe/src/Panzer_AssemblyEngine_impl.hpp:152-181) and the ~9 Kokkos dispatches per 20-cell
// workset behind it with ONE kernel launch:
//
//   setup (once):  rows (local DOFs) are ordered along a Morton curve of their node coordinates; tiles are the
//     leaves of the Morton octree with <= TR rows; each tile gets the list of cells touching its rows (own + halo
//     cells, in cell-id order), a per-row table "which tile cell has me as local vertex a", per row the
//     permutation from the canonical 27-point neighbour index to the CSR slot of that column, the runs of rows
//     contiguous in A, and -- when all its cells are translates of one parallelepiped ("congruent") -- the
//     stiffness row of an interior node.
//   evaluate:      persistent CTAs walk the tiles; the next tile's LID block arrives by a TMA bulk load.
//     phase 1 (thread per tile cell): gather LIDs, coordinates, solution; geometry; stage per-cell
//       data in shared memory (constant-Jacobian cells: 14 metric terms + gathered u + source load;
//       general cells: the 36+8 element matrix/vector from the full 2x2x2 rule).
//     phase 2 (thread per row): walk the <=8 cells around the node with the local vertex index as a
//       compile-time constant, accumulate the 27 row entries in REGISTERS (canonical neighbour
//       index is compile time), residual alongside; interior rows of congruent tiles take the tile's
//       precomputed stiffness row and gather their 27 neighbour values instead.
//     phase 3: permute into CSR slot order through shared memory; every run of rows leaves by one TMA bulk store.
//   => no atomics, no zero-fill pass, no colind reads, each A value and f value written exactly
//      once, bitwise reproducible.  Halo cells are recomputed by neighbouring tiles (~1.6x phase 1).
//
// Rows whose neighbourhood is not a regular 27-point patch (irregular valence, repeated local
// index, > 64 entries) are left to the general row-gather kernel (fill_rowgather.cu).
#include "tiles.hpp"
#include <cub/cub.cuh>
#include <algorithm>

namespace txasm {

int launch_fill_rowgather_list(txasm_handle h, const FillArgs &a, const int *row_list, int64_t n);

// ============================================================================ setup kernels
__global__ void k_row_regular(int64_t n_rows, const int64_t *__restrict__ adj_ptr, const int *__restrict__ adj,
                              const int *__restrict__ lids, const int64_t *__restrict__ rowptr,
                              const int *__restrict__ colind, unsigned char *__restrict__ regular,
                              unsigned char *__restrict__ perm, int *__restrict__ adjcell, int *__restrict__ maxlen)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  bool ok = true;
  int ac[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) ac[a] = -1;
  for (int64_t k = adj_ptr[r]; k < adj_ptr[r + 1]; ++k) {
    const int a = adj[k] & 7, cell = adj[k] >> 3;
    if (ac[a] != -1) ok = false;
    ac[a] = cell;
  }
  int colc[27];
  for (int c = 0; c < 27; ++c) colc[c] = -1;
  for (int a = 0; a < 8 && ok; ++a) {
    if (ac[a] < 0) continue;
    for (int b = 0; b < 8; ++b) {
      const int c = canon(a, b);
      const int col = lids[(int64_t)ac[a] * 8 + b];
      if (colc[c] == -1) colc[c] = col;
      else if (colc[c] != col) ok = false;
    }
  }
  for (int c1 = 0; c1 < 27 && ok; ++c1)
    for (int c2 = c1 + 1; c2 < 27; ++c2)
      if (colc[c1] >= 0 && colc[c1] == colc[c2]) ok = false;
  if (ok && colc[13] != (int)r) ok = false;
  const int64_t b0 = rowptr[r];
  const int len = (int)(rowptr[r + 1] - b0);
  if (len > LROW_CAP) ok = false;
  int nvalid = 0;
  for (int c = 0; c < PERM_STRIDE; ++c) {
    unsigned char p = 0xFF;
    if (c == 27) p = (unsigned char)((nvalid != len) ? 1 : 0);      // row has slots no local cell writes
    if (ok && c < 27 && colc[c] >= 0) {
      int lo = 0, hi = len - 1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int v = colind[b0 + mid];
        if (v == colc[c]) { p = (unsigned char)mid; ++nvalid; break; }
        if (v < colc[c]) lo = mid + 1; else hi = mid - 1;
      }
    }
    perm[r * PERM_STRIDE + c] = p;
  }
  regular[r] = ok ? 1 : 0;
#pragma unroll
  for (int a = 0; a < 8; ++a) adjcell[r * 8 + a] = ok ? ac[a] : -1;
  if (ok) atomicMax(maxlen, len);
}

__device__ __forceinline__ unsigned long long dbl_key(double v)
{
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_dbl(unsigned long long k)
{
  unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}
__global__ void k_bbox(int64_t n, const double *__restrict__ xyz, unsigned long long *__restrict__ mn, unsigned long long *__restrict__ mx)
{
  // grid-stride, per-thread min/max, warp shuffle reduction, one atomic per warp and axis
  unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0, 0, 0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    for (int d = 0; d < 3; ++d) {
      const unsigned long long k = dbl_key(xyz[i * 3 + d]);
      lo[d] = k < lo[d] ? k : lo[d];
      hi[d] = k > hi[d] ? k : hi[d];
    }
  for (int d = 0; d < 3; ++d) {
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo[d], o), b = __shfl_xor_sync(0xffffffffu, hi[d], o);
      lo[d] = a < lo[d] ? a : lo[d];
      hi[d] = b > hi[d] ? b : hi[d];
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mn[d], lo[d]); atomicMax(&mx[d], hi[d]); }
  }
}
__device__ __forceinline__ unsigned long long spread3(unsigned long long v)
{  // 21 bits -> every third bit
  v &= 0x1fffffull;
  v = (v | (v << 32)) & 0x1f00000000ffffull;
  v = (v | (v << 16)) & 0x1f0000ff0000ffull;
  v = (v | (v << 8)) & 0x100f00f00f00f00full;
  v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}
// smallest bounding-box extent of a cell along each axis (the mesh spacing of a uniform grid)
__global__ void k_cell_hmin(int64_t n_cells, const int *__restrict__ lids, const double *__restrict__ xyz, unsigned long long *__restrict__ hmin)
{
  double h[3] = {1e300, 1e300, 1e300};
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_cells; e += (int64_t)gridDim.x * blockDim.x)
    for (int d = 0; d < 3; ++d) {
      double lo = 1e300, hi = -1e300;
      for (int a = 0; a < 8; ++a) { const double v = xyz[(int64_t)lids[e * 8 + a] * 3 + d]; lo = fmin(lo, v); hi = fmax(hi, v); }
      if (hi > lo) h[d] = fmin(h[d], hi - lo);
    }
  for (int d = 0; d < 3; ++d) {
    for (int o = 16; o > 0; o >>= 1) h[d] = fmin(h[d], __shfl_xor_sync(0xffffffffu, h[d], o));
    if ((threadIdx.x & 31) == 0) atomicMin(&hmin[d], dbl_key(h[d]));
  }
}
__global__ void k_morton(int64_t n, const double *__restrict__ xyz, const unsigned long long *__restrict__ mn,
                         const unsigned long long *__restrict__ mx, const unsigned long long *__restrict__ hmin,
                         const unsigned char *__restrict__ regular, unsigned long long *__restrict__ keys, int *__restrict__ vals)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  vals[i] = (int)i;
  if (!regular[i]) { keys[i] = ~0ull; return; }
  // Quantum per axis = the smallest cell extent / 2^k: on a uniform grid the nodes then sit on multiples of 2^k
  // quanta and the octree boxes hold exactly 8 x 8 x 4 of them, whatever the number of nodes per axis.  k is the
  // largest power that keeps 21 bits per axis; a mesh too anisotropic for that gets one isotropic quantum.
  double ext[3], h[3], cells_max = 0.0, ext_max = 0.0;
  for (int d = 0; d < 3; ++d) {
    ext[d] = key_dbl(mx[d]) - key_dbl(mn[d]);
    h[d] = key_dbl(hmin[d]);
    ext_max = fmax(ext_max, ext[d]);
    if (h[d] > 0.0 && h[d] < 1e299) cells_max = fmax(cells_max, ext[d] / h[d]); else cells_max = 1e300;
  }
  double scale[3];
  if (cells_max < 1048576.0) {
    double f = 1.0;
    while (cells_max * f * 2.0 < 2097151.0 && f < 4096.0) f *= 2.0;
    for (int d = 0; d < 3; ++d) scale[d] = f / h[d];
  } else {
    for (int d = 0; d < 3; ++d) scale[d] = (ext_max > 0.0) ? 1048576.0 / ext_max : 0.0;
  }
  unsigned long long q[3];
  for (int d = 0; d < 3; ++d) {
    const double t = (xyz[i * 3 + d] - key_dbl(mn[d])) * scale[d];
    q[d] = (unsigned long long)fmin(fmax(t + 0.5, 0.0), 2097151.0);
  }
  keys[i] = spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2);
}

__global__ void k_tile_rows(int64_t n_slots, int64_t n_regular, const int *__restrict__ sorted, int *__restrict__ tile_rows)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_slots) tile_rows[i] = (i < n_regular) ? sorted[i] : -1;
}
// Tiles are the leaves of the Morton octree with at most TR rows: the coarsest aligned key block (key >> L equal)
// that holds <= TR of the sorted keys.  Unlike "every TR consecutive keys", a leaf never straddles two boxes, so
// tiles stay compact (cells per tile <= 405 for any uniform grid, not only 2^k+1 nodes per axis) at the price of
// some tiles with fewer than TR rows.  flag[i] = 1 when sorted position i starts a leaf.
// max_level caps the box: with the per-axis quantum of k_morton a leaf spans at most 16 nodes per axis, so the lines of
// nodes along an edge of the mesh (1 x 1 x N: one sparse high-level box) become lattice tiles of 16 rows instead of one
// tile with a 2 x 2 x (N + 1) lattice.
__global__ void k_leaf_starts(int64_t n, const unsigned long long *__restrict__ keys, int TR, int max_level, int *__restrict__ flag)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = keys[i];
  // members of the level-L block around i, counted inside the window [i-TR, i+TR]; > TR when it reaches the window edge
  auto bounds = [&](int L, int64_t &lb, int64_t &ub) {
    const unsigned long long lo_key = (L >= 64) ? 0ull : ((key >> L) << L);
    const unsigned long long hi_key = (L >= 64) ? ~0ull : (lo_key | ((L == 0) ? 0ull : ((~0ull) >> (64 - L))));
    int64_t a = (i - TR - 1 > 0) ? i - TR - 1 : 0, b = i;           // first j in [a, i] with keys[j] >= lo_key
    while (a < b) { const int64_t m = (a + b) >> 1; if (keys[m] >= lo_key) b = m; else a = m + 1; }
    lb = a;
    a = i; b = (i + TR + 2 < n) ? i + TR + 2 : n;                   // first j in (i, b] with keys[j] > hi_key
    while (a < b) { const int64_t m = (a + b) >> 1; if (keys[m] > hi_key) b = m; else a = m + 1; }
    ub = a;
  };
  int lo = 0, hi = max_level;            // largest L <= max_level with count <= TR (count is monotone in L)
  int64_t lb, ub;
  bounds(0, lb, ub);
  if (ub - lb > TR) {                    // more than TR coincident keys: cut by count
    flag[i] = ((i - lb) % TR == 0) ? 1 : 0;
    return;
  }
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    bounds(mid, lb, ub);
    if (ub - lb <= TR) lo = mid; else hi = mid - 1;
  }
  bounds(lo, lb, ub);
  flag[i] = (lb == i) ? 1 : 0;
}
__global__ void k_tile_starts(int64_t n, const int *__restrict__ flag, const int *__restrict__ tile_of, int *__restrict__ start)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && flag[i]) start[tile_of[i] - 1] = (int)i;
}
// key = (tile index, row id): sorting it orders the rows INSIDE each tile by row id, so that consecutive
// threads own consecutive rows (contiguous A / f stores) and touch consecutive tile cells (no bank conflicts)
__global__ void k_tile_keys(int64_t n, const int *__restrict__ tile_of, const int *__restrict__ sorted, unsigned long long *__restrict__ keys)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((unsigned long long)(tile_of[i] - 1) << 32) | (unsigned int)sorted[i];
}
__global__ void k_tile_rows_from_keys(int64_t n, int TR, const unsigned long long *__restrict__ keys, const int *__restrict__ start,
                                      int *__restrict__ tile_rows)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = (int)(keys[i] >> 32);
  tile_rows[(int64_t)t * TR + (i - start[t])] = (int)(keys[i] & 0xffffffffull);
}
// tile-ordered row tables: CSR begin/length and the perm bytes of every tile row, contiguous per tile
__global__ void k_tile_rowtables(int64_t n_slots, const int *__restrict__ tile_rows, const int64_t *__restrict__ rowptr,
                                 const unsigned char *__restrict__ perm, unsigned char *__restrict__ tperm)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  const int row = tile_rows[i];
  uint4 a = make_uint4(~0u, ~0u, ~0u, ~0u), b = a;
  if (row >= 0) {
    const uint4 *pp = reinterpret_cast<const uint4 *>(perm + (int64_t)row * 32);
    a = pp[0]; b = pp[1];
  }
  uint4 *o = reinterpret_cast<uint4 *>(tperm + i * 32);
  o[0] = a; o[1] = b;
}
// Runs: maximal sequences of tile rows (slot order) whose CSR rows are contiguous in A (and, in a congruent tile, all
// uniform or all not).  Each run is written to
// global memory by ONE TMA bulk store from the shared-memory out buffer, in which the run sits at an offset with
// the same 16-byte phase as its global address.  Pass FILL=false counts runs and the out-buffer size per tile.
// A row of a congruent tile is UNIFORM when all 8 cells around it exist and its 27 columns sit in canonical order
// (perm = identity; true for interior nodes under lexicographic and under first-touch numbering): its values are
// cK * Kf[0..26] in storage order, the same 216 bytes for every such row.  A run of uniform rows (RUN_UNIFORM in
// RowRun::n) is stored straight from a constant shared-memory image of that pattern, and a tile whose rows are all
// uniform (tile_cong = 2) needs neither the placement phase nor its barriers nor the drain wait.
template <bool FILL>
__global__ void k_tile_runs(int n_tiles, int TR, const int *__restrict__ tile_rows, const int64_t *__restrict__ rowptr,
                            const unsigned char *__restrict__ tperm, const unsigned short *__restrict__ adjl,
                            unsigned char *__restrict__ tile_cong, int *__restrict__ nruns, int *__restrict__ outsize,
                            const int64_t *__restrict__ run_ptr, RowRun *__restrict__ runs, unsigned *__restrict__ rowinfo)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  int nr = 0, cursor = 0, soff = 0;
  long long run_beg = 0, prev_end = -1;
  const int64_t rb = FILL ? run_ptr[t] : 0;
  const bool cong = tile_cong[t] != 0;
  bool run_uni = true, all_uni = cong, prev_uni = false;
  for (int sl = 0; sl < TR; ++sl) {
    const int64_t slot = (int64_t)t * TR + sl;
    const int row = tile_rows[slot];
    if (row < 0) { if (FILL) rowinfo[slot] = 0xFFFFu; continue; }
    const long long beg = rowptr[row];
    const int len = (int)(rowptr[row + 1] - beg);
    bool uni = cong && len == 27;
    for (int c = 0; c < 27 && uni; ++c) uni = (tperm[slot * 32 + c] == c);
    for (int a = 0; a < 8 && uni; ++a) uni = (adjl[slot * 8 + a] != 0xFFFF);
    // start a run: the row does not follow its predecessor in A, or it changes between uniform and not (a run of a
    // congruent tile is either stored whole from the constant row image or not at all: on the faces of a mesh an
    // x-line = one boundary row + interior rows, and only the boundary row needs values of its own)
    if (prev_end != beg || (cong && uni != prev_uni)) {
      if (FILL && nr > 0) runs[rb + nr - 1].n = (int)(prev_end - run_beg) | (run_uni ? RUN_UNIFORM : 0);
      soff = ((cursor + 1) & ~1) + (int)(beg & 1);            // same parity (16-byte phase) as the global index
      run_beg = beg;
      run_uni = true;
      if (FILL) { runs[rb + nr].beg = beg; runs[rb + nr].soff = soff; }
      ++nr;
    }
    run_uni = run_uni && uni;
    all_uni = all_uni && uni;
    prev_uni = uni;
    const int off = soff + (int)(beg - run_beg);
    if (FILL) rowinfo[slot] = (unsigned)off | ((unsigned)len << 16) | ((unsigned)(tperm[slot * 32 + 27] & 1) << 24);
    cursor = off + len;
    prev_end = beg + len;
  }
  if (FILL && nr > 0) runs[rb + nr - 1].n = (int)(prev_end - run_beg) | (run_uni ? RUN_UNIFORM : 0);
  if (FILL) {
    // a row is only treated as uniform when its whole run is (the run then never touches the out buffer)
    for (int i = 0, sl = 0; sl < TR; ++sl) {
      const int64_t slot = (int64_t)t * TR + sl;
      const int row = tile_rows[slot];
      if (row < 0) continue;
      while (i + 1 < nr && rowptr[row] >= runs[rb + i + 1].beg) ++i;
      if (runs[rb + i].n & RUN_UNIFORM) rowinfo[slot] |= ROW_UNIFORM;
    }
    if (all_uni && nr > 0) tile_cong[t] |= 4;
  }
  if (!FILL) { nruns[t] = nr; outsize[t] = cursor; }
}
// per-tile copy of the LID table in tile-cell order: phase 1 reads it fully coalesced, one dependent load less
__global__ void k_tile_lids(int64_t n, const int *__restrict__ cells, const int *__restrict__ lids, int *__restrict__ out)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // (tile cell, vertex)
  if (i < n * 8) { const int c = cells[i >> 3]; out[i] = (c >= 0) ? lids[(int64_t)c * 8 + (i & 7)] : -1; }
}

// One CTA per tile: sorted unique list of the cells around the tile's rows, and the tile-local index
// of each (row, a) cell.  CAP = TR*8 candidates are bitonic-sorted in shared memory.
template <int TR>
__global__ void __launch_bounds__(TR) k_tile_cells(const int *__restrict__ tile_rows, const int *__restrict__ adjcell,
                                                   int *__restrict__ ncells, int *__restrict__ cells_tmp,
                                                   unsigned short *__restrict__ adjl)
{
  constexpr int CAP = TR * 8;
  __shared__ int s[CAP];
  __shared__ int u[CAP];
  __shared__ int s_count;
  const int t = blockIdx.x, tid = threadIdx.x;
  const int row = tile_rows[(int64_t)t * TR + tid];
  int mine[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    mine[a] = (row >= 0) ? adjcell[(int64_t)row * 8 + a] : -1;
    s[tid * 8 + a] = mine[a] >= 0 ? mine[a] : 0x7fffffff;
  }
  __syncthreads();
  for (int k = 2; k <= CAP; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < CAP; i += TR) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const int a = s[i], b = s[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { s[i] = b; s[ixj] = a; }
        }
      }
      __syncthreads();
    }
  if (tid == 0) s_count = 0;
  __syncthreads();
  // unique: each thread owns 8 consecutive sorted entries
  int flags = 0, cnt = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = tid * 8 + k;
    const bool first = (s[i] != 0x7fffffff) && (i == 0 || s[i] != s[i - 1]);
    if (first) { flags |= 1 << k; ++cnt; }
  }
  typedef cub::BlockScan<int, TR> Scan;
  __shared__ typename Scan::TempStorage tmp;
  int off, total;
  Scan(tmp).ExclusiveSum(cnt, off, total);
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (flags & (1 << k)) u[off++] = s[tid * 8 + k];
  __syncthreads();
  // Staging position of a cell = its rank in cell-id order.  With x-fastest cell numbering the cells met by
  // consecutive rows of an x-line -- including the line's x-1 halo cell -- are consecutive for every local
  // vertex a, so the 16-byte phase-2 loads of a quarter-warp fall into distinct banks.  (Anchoring cells at the
  // slot of their vertex-0 row left that halo cell at an arbitrary position: 118M -> 52M conflict wavefronts
  // and 2.62 -> 2.42 ms at 256^3 when it was replaced by this order.)
  auto find = [&](int cell) {
    int lo = 0, hi = total - 1;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1;
      if (u[mid] == cell) return mid;
      if (u[mid] < cell) lo = mid + 1; else hi = mid - 1;
    }
    return 0xFFFF;
  };
  if (tid == 0) ncells[t] = total;
  for (int i = tid; i < total; i += TR) cells_tmp[(int64_t)t * CAP + i] = u[i];
#pragma unroll
  for (int a = 0; a < 8; ++a)
    adjl[((int64_t)t * TR + tid) * 8 + a] = (mine[a] >= 0) ? (unsigned short)find(mine[a]) : (unsigned short)0xFFFF;
}

__global__ void k_compact_cells(int n_tiles, int cap, const int *__restrict__ ncells, const int64_t *__restrict__ ptr,
                                const int *__restrict__ tmp, int *__restrict__ out)
{
  const int t = blockIdx.x;
  if (t >= n_tiles) return;
  const int n = ncells[t];
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[ptr[t] + i] = tmp[(int64_t)t * cap + i];
}

// A tile is CONGRUENT when all its cells are translates of its first cell: the Jacobian columns (half the edge
// vectors from vertex 0 to vertices 1, 3, 4 -- what phase 1 computes) agree within tol * (largest entry), tol being
// the same affine_tol (default 1e-13) that already decides whether a cell counts as a parallelepiped.  (On an
// inline mesh the cells differ only by the rounding of i*h + x0, about eps * n relative.)  Such a tile stages the
// metric split D/O once (cell 0) and the row threads read it as a shared-memory broadcast.
__global__ void k_tile_congruent(int n_tiles, const int64_t *__restrict__ cell_ptr, const int *__restrict__ cells,
                                 const int *__restrict__ lids, const double *__restrict__ xyz,
                                 const unsigned char *__restrict__ aff, double tol, unsigned char *__restrict__ flag)
{
  const int t = blockIdx.x;
  if (t >= n_tiles) return;
  const int64_t cb = cell_ptr[t];
  const int nc = (int)(cell_ptr[t + 1] - cb);
  __shared__ double J0[9];
  __shared__ double jmax;
  __shared__ int bad, skew;
  auto jac = [&](int cell, double (&J)[9]) {
    const int *l = lids + (int64_t)cell * 8;
    for (int d = 0; d < 3; ++d) {
      const double x0 = xyz[(int64_t)l[0] * 3 + d];
      J[d * 3 + 0] = 0.5 * (xyz[(int64_t)l[1] * 3 + d] - x0);
      J[d * 3 + 1] = 0.5 * (xyz[(int64_t)l[3] * 3 + d] - x0);
      J[d * 3 + 2] = 0.5 * (xyz[(int64_t)l[4] * 3 + d] - x0);
    }
  };
  if (threadIdx.x == 0) {
    bad = (nc == 0); skew = 0;
    if (nc) {
      double J[9]; jac(cells[cb], J);
      double m = 0.0;
      for (int k = 0; k < 9; ++k) { J0[k] = J[k]; m = fmax(m, fabs(J[k])); }
      jmax = m;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nc; i += blockDim.x) {
    const int c = cells[cb + i];
    if (!aff[c]) { bad = 1; continue; }
    double J[9]; jac(c, J);
    for (int k = 0; k < 9; ++k) {
      if (fabs(J[k] - J0[k]) > tol * jmax) bad = 1;
      if ((k % 4) != 0 && J[k] != 0.0) skew = 1;          // off-diagonal entry: the cell is not axis-aligned
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) flag[t] = bad ? 0 : (skew ? 1 : 3);   // bit 0 congruent, bit 1 axis-aligned (bit 2: k_tile_runs)
}

__global__ void k_tile_affine(int64_t n, const int *__restrict__ cells, const unsigned char *__restrict__ aff, int *__restrict__ n_non)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && cells[i] >= 0 && !aff[cells[i]]) atomicAdd(n_non, 1);
}

__global__ void k_mark_rows(int64_t n_slots, const int *__restrict__ tile_rows, unsigned char *__restrict__ mark)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_slots && tile_rows[i] >= 0) mark[tile_rows[i]] = 1;
}
__global__ void k_count_marked(int64_t n, const int *__restrict__ rows, const unsigned char *__restrict__ mark, int *__restrict__ count)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && rows[i] >= 0 && mark[rows[i]]) atomicAdd(count, 1);
}

__global__ void k_permute_tiles(int64_t n_slots, int TR, const int *__restrict__ src_tile, const int *__restrict__ in, int *__restrict__ out)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_slots) out[i] = in[(int64_t)src_tile[i / TR] * TR + i % TR];
}

__global__ void k_list_irregular(int64_t n_rows, const unsigned char *__restrict__ regular, int *__restrict__ list, int *__restrict__ count)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_rows && !regular[i]) list[atomicAdd(count, 1)] = (int)i;
}

// ============================================================================ the fill kernel
// Staging per tile cell, k-major with compile-time stride TEP (so shared-memory offsets are immediates):
//   affine : D[8] | O1[3] O2[3] | ug[8] | (Mm[4] um[8]) | src[8]
//   general: K sym[36] | r[8]
// D/O are the constant-Jacobian stiffness split by sign pattern: for vertices a,b of a cell let
// p_d = s^d_a s^d_b (+-1).  Then  int grad phi_a . grad phi_b = D[p] + sum_{pairs d<e, p_d=p_e} sgn O{1|2}_de
// with D[p] = sum_d G_dd p_d c_e c_f / 8, c = 1 + p/3, O1 = G_de/3 (c_f = 4/3), O2 = G_de/6 (c_f = 2/3),
// sgn = p_d s^d_a s^e_a: an entry costs 1-3 additions, no constants.  Mm[k] = det c_x c_y c_z / 8 for k minus signs.
__host__ __device__ constexpr int stage_doubles(bool affine, bool mass, bool src)
{
  return affine ? (14 + 8 + (mass ? 12 : 0) + (src ? 8 : 0)) : (36 + 8);
}
__host__ __device__ constexpr int pidx(int a, int b)
{
  return (hex_sx(a) * hex_sx(b) < 0 ? 1 : 0) | (hex_sy(a) * hex_sy(b) < 0 ? 2 : 0) | (hex_sz(a) * hex_sz(b) < 0 ? 4 : 0);
}
__host__ __device__ constexpr int pminus(int p) { return (p & 1) + ((p >> 1) & 1) + ((p >> 2) & 1); }
// coefficient of G_dd in D[p]
__host__ __device__ constexpr double dcoef(int p, int d)
{
  const int pd = ((p >> d) & 1) ? -1 : 1;
  const int e = (d + 1) % 3, f = (d + 2) % 3;
  const double ce = ((p >> e) & 1) ? 2.0 / 3.0 : 4.0 / 3.0, cf = ((p >> f) & 1) ? 2.0 / 3.0 : 4.0 / 3.0;
  return 0.125 * pd * ce * cf;
}

// Shared-memory staging is stored as PAIRS: logical entry k of cell j lives in double2 slot (k/2)*TEP + j,
// component k%2, so phase 1 writes and phase 2 reads 16 bytes per instruction (half the LDS/STS count).
template <int TEP> __device__ __forceinline__ double2 ld2(const double *sm, int q, int j)
{ return reinterpret_cast<const double2 *>(sm)[q * TEP + j]; }
template <int TEP> __device__ __forceinline__ void st2(double *sm, int q, int j, double x, double y)
{ reinterpret_cast<double2 *>(sm)[q * TEP + j] = make_double2(x, y); }
template <int TEP> __device__ __forceinline__ double ld1(const double *sm, int k, int j)
{ return sm[(((k) >> 1) * TEP + j) * 2 + ((k) & 1)]; }
template <int TEP> __device__ __forceinline__ void st1(double *sm, int k, int j, double v)
{ sm[(((k) >> 1) * TEP + j) * 2 + ((k) & 1)] = v; }

// logical staging layout (affine): 0-7 D[p] | 8+2*pr+{0:O1,1:O2}, pr = xy,yz,zx | 14-21 ug | (22-25 Mm, 26-33 um) | src[8]
template <int A, int B, bool JAC>
__device__ __forceinline__ void kab_affine(const double2 (&dq)[4], const double2 (&oq)[3], double cK, double ub, double &acc, double &fr)
{
  constexpr int p = pidx(A, B);
  double t = (p & 1) ? dq[p >> 1].y : dq[p >> 1].x;
  // pairs (x,y) f=z, (y,z) f=x, (z,x) f=y
#define TX_OFF(PR, D_, E_, F_)                                                                       \
  if (((p >> D_) & 1) == ((p >> E_) & 1)) {                                                          \
    constexpr int sgn = (((p >> D_) & 1) ? -1 : 1) * hex_s(A, D_) * hex_s(A, E_);                    \
    const double o = ((p >> F_) & 1) ? oq[PR].y : oq[PR].x;                                          \
    t = (sgn > 0) ? t + o : t - o;                                                                   \
  }
  TX_OFF(0, 0, 1, 2) TX_OFF(1, 1, 2, 0) TX_OFF(2, 2, 0, 1)
#undef TX_OFF
  if (JAC) acc = fma(cK, t, acc);
  fr = fma(t, ub, fr);
}

template <int TEP, int A, bool JAC>
__device__ __forceinline__ void row_accum_affine(const double *__restrict__ sm, int el, int elm, const FillCoef &c,
                                                 bool has_mass, bool has_src, double (&acc)[27], double &fr)
{
  double2 dq[4], oq[3], uq[4];          // metric split from cell elm (= el, or the tile's first cell: a broadcast load)
#pragma unroll
  for (int q = 0; q < 4; ++q) dq[q] = ld2<TEP>(sm, q, elm);
#pragma unroll
  for (int q = 0; q < 3; ++q) oq[q] = ld2<TEP>(sm, 4 + q, elm);
#pragma unroll
  for (int q = 0; q < 4; ++q) uq[q] = ld2<TEP>(sm, 7 + q, el);
#define TX_KAB(B) kab_affine<A, B, JAC>(dq, oq, c.cK, ((B) & 1) ? uq[(B) >> 1].y : uq[(B) >> 1].x, acc[canon(A, B)], fr);
  TX_KAB(0) TX_KAB(1) TX_KAB(2) TX_KAB(3) TX_KAB(4) TX_KAB(5) TX_KAB(6) TX_KAB(7)
#undef TX_KAB
  if (has_mass) {
    const double2 m01 = ld2<TEP>(sm, 11, el), m23 = ld2<TEP>(sm, 12, el);
    const double mm[4] = {m01.x, m01.y, m23.x, m23.y};
    double2 vq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) vq[q] = ld2<TEP>(sm, 13 + q, el);
#define TX_MAB(B)                                                               \
    { const double m = mm[pminus(pidx(A, B))];                                  \
      if (JAC) acc[canon(A, B)] = fma(c.cM, m, acc[canon(A, B)]);               \
      fr = fma(m, ((B) & 1) ? vq[(B) >> 1].y : vq[(B) >> 1].x, fr); }
    TX_MAB(0) TX_MAB(1) TX_MAB(2) TX_MAB(3) TX_MAB(4) TX_MAB(5) TX_MAB(6) TX_MAB(7)
#undef TX_MAB
  }
  if (has_src) fr += ld1<TEP>(sm, (has_mass ? 34 : 22) + A, el);
}

// General hexahedra: row A of the element matrix of `cell` straight from the records k_elem_general wrote
// (fill_general.cu: K upper triangle [36] | r[8], 352 bytes per cell).  Every matrix row is read by exactly one thread of the
// whole grid -- the owner of that DOF -- so there is no halo traffic and nothing to stage.
template <int A, bool JAC>
__device__ __forceinline__ void row_accum_general(const double *__restrict__ elem, int64_t cell, double (&acc)[27], double &fr)
{
  const double *rec = elem + cell * ELEM_REC;
  if (JAC) {
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[canon(A, b)] += __ldg(rec + sym_idx(A, b));      // (compile-time offsets; K[A][A..7] is contiguous)
  }
  fr += __ldg(rec + 36 + A);
}

// phase 1, constant-Jacobian cell: geometry from vertices 0,1,3,4 (a parallelepiped is fixed by them),
// metric split D/O, gathered solution, source load vector
template <int TEP>
__device__ __forceinline__ void stage_affine(double *__restrict__ sm, int j, int64_t e, const double (&X0)[3], const double (&X1)[3],
                                             const double (&X3)[3], const double (&X4)[3], const double (&ug)[8],
                                             const FillCoef &c, bool has_mass, bool has_src, bool metric)
{
  double J[3][3], xc[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    J[d][0] = 0.5 * (X1[d] - X0[d]); J[d][1] = 0.5 * (X3[d] - X0[d]); J[d][2] = 0.5 * (X4[d] - X0[d]);
    xc[d] = X0[d] + (J[d][0] + J[d][1] + J[d][2]);
  }
  const double c0 = J[1][1] * J[2][2] - J[2][1] * J[1][2];
  const double c1 = J[2][0] * J[1][2] - J[1][0] * J[2][2];
  const double c2 = J[1][0] * J[2][1] - J[2][0] * J[1][1];
  const double det = J[0][0] * c0 + J[0][1] * c1 + J[0][2] * c2;
  if (metric) {                           // (congruent tile: only its first cell stages the metric split)
  const double idet = 1.0 / det;
  double Ji[3][3];  // Ji[e][d] = d xi_e / d x_d
  Ji[0][0] = c0 * idet; Ji[1][0] = c1 * idet; Ji[2][0] = c2 * idet;
  Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * idet;
  Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * idet;
  Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * idet;
  Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * idet;
  Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * idet;
  Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * idet;
  double G[6];   // det * Jinv Jinv^T : xx yy zz xy yz zx
  G[0] = det * (Ji[0][0] * Ji[0][0] + Ji[0][1] * Ji[0][1] + Ji[0][2] * Ji[0][2]);
  G[1] = det * (Ji[1][0] * Ji[1][0] + Ji[1][1] * Ji[1][1] + Ji[1][2] * Ji[1][2]);
  G[2] = det * (Ji[2][0] * Ji[2][0] + Ji[2][1] * Ji[2][1] + Ji[2][2] * Ji[2][2]);
  G[3] = det * (Ji[0][0] * Ji[1][0] + Ji[0][1] * Ji[1][1] + Ji[0][2] * Ji[1][2]);
  G[4] = det * (Ji[1][0] * Ji[2][0] + Ji[1][1] * Ji[2][1] + Ji[1][2] * Ji[2][2]);
  G[5] = det * (Ji[2][0] * Ji[0][0] + Ji[2][1] * Ji[0][1] + Ji[2][2] * Ji[0][2]);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    st2<TEP>(sm, q, j, G[0] * dcoef(2 * q, 0) + G[1] * dcoef(2 * q, 1) + G[2] * dcoef(2 * q, 2),
             G[0] * dcoef(2 * q + 1, 0) + G[1] * dcoef(2 * q + 1, 1) + G[2] * dcoef(2 * q + 1, 2));
#pragma unroll
  for (int k = 0; k < 3; ++k) st2<TEP>(sm, 4 + k, j, G[3 + k] * (1.0 / 3.0), G[3 + k] * (1.0 / 6.0));
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) st2<TEP>(sm, 7 + q, j, ug[2 * q], ug[2 * q + 1]);
  int base = 22;
  if (has_mass) {
    st2<TEP>(sm, 11, j, det * (8.0 / 27.0), det * (4.0 / 27.0));
    st2<TEP>(sm, 12, j, det * (2.0 / 27.0), det * (1.0 / 27.0));
    base = 34;                            // um[8] at 26..33 is staged by the caller's mass pass
  }
  if (has_src) {
    // source load vector  det * sum_q N_a(xi_q) sum_s mult_s s_s(x_q),  x_q = xc + J xi_q.
    // Axis-aligned cell (diagonal J): x_q, y_q, z_q take two values each, so a separable closure model
    // needs 6 instead of 24 function evaluations -- the same 8 point values.
    const bool diag = (J[0][1] == 0.0) & (J[0][2] == 0.0) & (J[1][0] == 0.0) & (J[1][2] == 0.0) &
                      (J[2][0] == 0.0) & (J[2][1] == 0.0);
    bool slow = false;
    for (int s = 0; s < c.n_src; ++s) slow |= (c.src_id[s] == TXASM_SOURCE_SIN3) && !diag;
    if (slow) {
      // general position: evaluate the closure models at the 8 points one by one
      double bl[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) bl[a] = 0.0;
#pragma unroll 1
      for (int q = 0; q < 8; ++q) {
        const double xi = (q & 1) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
        const double et = (q & 2) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
        const double ze = (q & 4) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
        const double xq = xc[0] + J[0][0] * xi + J[0][1] * et + J[0][2] * ze;
        const double yq = xc[1] + J[1][0] * xi + J[1][1] * et + J[1][2] * ze;
        const double zq = xc[2] + J[2][0] * xi + J[2][1] * et + J[2][2] * ze;
        double sq = 0.0;
        for (int s = 0; s < c.n_src; ++s) {
          const double v = (c.src_id[s] == TXASM_SOURCE_IP_ARRAY) ? c.src_ip[s][e * 8 + q] : source_eval(c.src_id[s], xq, yq, zq);
          sq = fma(c.src_mult[s], v, sq);
        }
        sq *= 0.125 * det;
#pragma unroll
        for (int a = 0; a < 8; ++a)
          bl[a] = fma((1.0 + hex_sx(a) * xi) * (1.0 + hex_sy(a) * et), (1.0 + hex_sz(a) * ze) * sq, bl[a]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) st2<TEP>(sm, (base >> 1) + q, j, bl[2 * q], bl[2 * q + 1]);
      return;
    }
    double sq8[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) sq8[q] = 0.0;
    for (int s = 0; s < c.n_src; ++s) {
      const int id = c.src_id[s];
      const double mult = c.src_mult[s];
      if (id == TXASM_SOURCE_IP_ARRAY) {
#pragma unroll
        for (int q = 0; q < 8; ++q) sq8[q] = fma(mult, c.src_ip[s][e * 8 + q], sq8[q]);
      } else if (id == TXASM_SOURCE_CONSTANT) {
#pragma unroll
        for (int q = 0; q < 8; ++q) sq8[q] += mult;
      } else if (id == TXASM_SOURCE_SIN3) {     // diag
        const double dx = J[0][0] * TX_INV_SQRT3, dy = J[1][1] * TX_INV_SQRT3, dz = J[2][2] * TX_INV_SQRT3;
        const double fx0 = sin2pi_fast(xc[0] - dx), fx1 = sin2pi_fast(xc[0] + dx);
        const double fy0 = sin2pi_fast(xc[1] - dy), fy1 = sin2pi_fast(xc[1] + dy);
        const double fz0 = (mult * 118.43525281307230) * sin2pi_fast(xc[2] - dz);
        const double fz1 = (mult * 118.43525281307230) * sin2pi_fast(xc[2] + dz);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          sq8[q] = fma(((q & 1) ? fx1 : fx0) * ((q & 2) ? fy1 : fy0), (q & 4) ? fz1 : fz0, sq8[q]);
      }
    }
    // N_a(xi_q) is a product of (1 +- 1/sqrt3)/2 factors: sum over q by sum factorisation (x, then y, then z)
    constexpr double wl = 0.5 * (1.0 - TX_INV_SQRT3), wh = 0.5 * (1.0 + TX_INV_SQRT3);
    double tx[2][4];   // [sx][qy,qz]
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      tx[0][r] = wh * sq8[2 * r] + wl * sq8[2 * r + 1];     // vertex at xi=-1: weight (1-xi_q)/2
      tx[1][r] = wl * sq8[2 * r] + wh * sq8[2 * r + 1];
    }
    double ty[2][2][2];  // [sx][sy][qz]
#pragma unroll
    for (int ix = 0; ix < 2; ++ix)
#pragma unroll
      for (int qz = 0; qz < 2; ++qz) {
        ty[ix][0][qz] = wh * tx[ix][2 * qz] + wl * tx[ix][2 * qz + 1];
        ty[ix][1][qz] = wl * tx[ix][2 * qz] + wh * tx[ix][2 * qz + 1];
      }
    double bl[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int ix = hex_sx(a) > 0, iy = hex_sy(a) > 0, iz = hex_sz(a) > 0;
      const double v = iz ? (wl * ty[ix][iy][0] + wh * ty[ix][iy][1]) : (wh * ty[ix][iy][0] + wl * ty[ix][iy][1]);
      bl[a] = det * v;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) st2<TEP>(sm, (base >> 1) + q, j, bl[2 * q], bl[2 * q + 1]);
  }
}

// Stiffness row of a node whose 8 cells are all translates of one parallelepiped: Kf[c] = sum over the cells (the
// node being vertex A of cell A) and their vertices B with canon(A,B) = c of  int grad phi_A . grad phi_B.  It only
// depends on the cell shape, so a congruent tile gets it once at setup (as the reference precomputes its
// IntegrationValues2 / BasisValues2 tables) and interior rows read it instead of visiting 8 cells.
__global__ void k_tile_kf(int n_tiles, const int64_t *__restrict__ cell_ptr, const int *__restrict__ cells,
                          const int *__restrict__ lids, const double *__restrict__ xyz,
                          const unsigned char *__restrict__ cong, double *__restrict__ kf)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  double *out = kf + (int64_t)t * KF_STRIDE;
  for (int c = 0; c < KF_STRIDE; ++c) out[c] = 0.0;
  if (!cong[t]) return;
  const int *l = lids + (int64_t)cells[cell_ptr[t]] * 8;
  double J[3][3];
  for (int d = 0; d < 3; ++d) {
    const double x0 = xyz[(int64_t)l[0] * 3 + d];
    J[d][0] = 0.5 * (xyz[(int64_t)l[1] * 3 + d] - x0);
    J[d][1] = 0.5 * (xyz[(int64_t)l[3] * 3 + d] - x0);
    J[d][2] = 0.5 * (xyz[(int64_t)l[4] * 3 + d] - x0);
  }
  const double c0 = J[1][1] * J[2][2] - J[2][1] * J[1][2];
  const double c1 = J[2][0] * J[1][2] - J[1][0] * J[2][2];
  const double c2 = J[1][0] * J[2][1] - J[2][0] * J[1][1];
  const double det = J[0][0] * c0 + J[0][1] * c1 + J[0][2] * c2;
  out[27] = J[0][0]; out[28] = J[1][1]; out[29] = J[2][2]; out[30] = det;
  const double idet = 1.0 / det;
  double Ji[3][3];
  Ji[0][0] = c0 * idet; Ji[1][0] = c1 * idet; Ji[2][0] = c2 * idet;
  Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * idet;
  Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * idet;
  Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * idet;
  Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * idet;
  Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * idet;
  Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * idet;
  double G[6];
  G[0] = det * (Ji[0][0] * Ji[0][0] + Ji[0][1] * Ji[0][1] + Ji[0][2] * Ji[0][2]);
  G[1] = det * (Ji[1][0] * Ji[1][0] + Ji[1][1] * Ji[1][1] + Ji[1][2] * Ji[1][2]);
  G[2] = det * (Ji[2][0] * Ji[2][0] + Ji[2][1] * Ji[2][1] + Ji[2][2] * Ji[2][2]);
  G[3] = det * (Ji[0][0] * Ji[1][0] + Ji[0][1] * Ji[1][1] + Ji[0][2] * Ji[1][2]);
  G[4] = det * (Ji[1][0] * Ji[2][0] + Ji[1][1] * Ji[2][1] + Ji[1][2] * Ji[2][2]);
  G[5] = det * (Ji[2][0] * Ji[0][0] + Ji[2][1] * Ji[0][1] + Ji[2][2] * Ji[0][2]);
  double D[8], O1[3], O2[3];
  for (int p = 0; p < 8; ++p) D[p] = G[0] * dcoef(p, 0) + G[1] * dcoef(p, 1) + G[2] * dcoef(p, 2);
  for (int k = 0; k < 3; ++k) { O1[k] = G[3 + k] * (1.0 / 3.0); O2[k] = G[3 + k] * (1.0 / 6.0); }
  for (int a = 0; a < 8; ++a)
    for (int b = 0; b < 8; ++b) {
      const int p = pidx(a, b);
      double tt = D[p];
      const int de[3][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}};     // pairs (x,y) f=z, (y,z) f=x, (z,x) f=y
      for (int pr = 0; pr < 3; ++pr) {
        const int d = de[pr][0], e = de[pr][1], f = de[pr][2];
        if (((p >> d) & 1) == ((p >> e) & 1)) {
          const int sgn = (((p >> d) & 1) ? -1 : 1) * hex_s(a, d) * hex_s(a, e);
          const double o = ((p >> f) & 1) ? O2[pr] : O1[pr];
          tt = (sgn > 0) ? tt + o : tt - o;
        }
      }
      out[canon(a, b)] += tt;
    }
}

// Persistent kernel: each CTA walks tiles blockIdx.x, +gridDim.x, ...  The LID block of the NEXT tile is pulled
// into shared memory by one TMA bulk copy while the current tile computes, and the node data the next tile will
// gather is prefetched into L2, so phase 1 starts from shared memory and hits in cache.
#ifdef TX_MINB_OVERRIDE
#define TX_MINB(TR, AFFINE) TX_MINB_OVERRIDE
#else
#define TX_MINB(TR, AFFINE) ((AFFINE) ? 512 / (TR) : 4)
#endif
template <int TR, int TEP, bool AFFINE, bool JAC>
__global__ void __launch_bounds__(TR, TX_MINB(TR, AFFINE)) k_fill_rowtile(FillArgs A, TileArgs T)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  int *lidbuf = reinterpret_cast<int *>(smem_raw + T.stage_bytes);
  const unsigned lidbuf_s = (unsigned)__cvta_generic_to_shared(lidbuf);
  const unsigned mbar = lidbuf_s + TEP * 32;
  // constant image of the uniform row pattern, cK*Kf repeated with period 27 (see k_tile_runs), and the values it holds
  double *img = reinterpret_cast<double *>(smem_raw + T.stage_bytes + TEP * 32 + 16);
  double *kfc = img + IMG_DOUBLES;
  const unsigned img_s = mbar + 16;
  const int tid = threadIdx.x, G = gridDim.x;
  const bool has_mass = A.c.has_mass != 0, has_src = A.c.n_src > 0;
  bool need_cell = !AFFINE;            // the global cell id is only needed to index per-cell IP arrays
  for (int s = 0; s < A.c.n_src; ++s) need_cell |= (A.c.src_id[s] == TXASM_SOURCE_IP_ARRAY);
  // the common case: exactly one solution vector feeds the GRADGRAD integrands
  int gv = -1, ngv = 0;
#pragma unroll
  for (int v = 0; v < 3; ++v) if (A.c.kg[v] != 0.0) { gv = v; ++ngv; }
  const bool one_g = (ngv == 1);
  const double *__restrict__ xg = one_g ? A.x[gv] : nullptr;
  const double kgv = one_g ? A.c.kg[gv] : 0.0;

  int t = T.t_begin + (int)blockIdx.x;
  tx_stamp(A.dbg, 2, false);
  int64_t cb = T.tile_cell_ptr[t];
  int ncell = (int)(T.tile_cell_ptr[t + 1] - cb);
  if (tid == 0) {
    mbar_init(mbar, 1);
    if (AFFINE) bulk_load(lidbuf_s, T.tile_lids + cb * 8, (unsigned)ncell * 32u, mbar);
  }
  if (AFFINE && JAC && tid < 28) kfc[tid] = __longlong_as_double(0x7ff8000000000000LL);   // no image yet
  __syncthreads();
  unsigned parity = 0;
  // general hexahedra: the global ids of the 8 cells around this thread's row, fetched one tile ahead (two dependent
  // loads off the critical path: the row pass then only waits for the element records)
  int gcell[8];
  if (!AFFINE) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(T.adjl + ((int64_t)t * TR + tid) * 8));
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int el = (int)((w[a >> 1] >> (16 * (a & 1))) & 0xFFFFu);
      gcell[a] = (el != 0xFFFF) ? __ldg(T.tile_cells + cb + el) : -1;
    }
  }

  for (; t < T.n_tiles; t += G) {
    const int tn = t + G;
    int64_t cbn = 0;
    int ncelln = 0;
    if (tn < T.n_tiles) { cbn = T.tile_cell_ptr[tn]; ncelln = (int)(T.tile_cell_ptr[tn + 1] - cbn); }
    uint4 alv_n = make_uint4(~0u, ~0u, ~0u, ~0u);
    if (!AFFINE && tn < T.n_tiles) alv_n = __ldg(reinterpret_cast<const uint4 *>(T.adjl + ((int64_t)tn * TR + tid) * 8));
    const int64_t slot = (int64_t)t * TR + tid;
    const int tcls = AFFINE ? (int)T.tile_cong[t] : 0;          // bit 0 congruent, bit 1 axis-aligned, bit 2 all rows uniform
    const bool cong = (tcls & 1) != 0;
    int64_t rb = 0;
    int nrun = 0;
    if (JAC) { rb = T.run_ptr[t]; nrun = (int)(T.run_ptr[t + 1] - rb); }   // used after phase 3; in flight meanwhile

    if (AFFINE) {
      mbar_wait(mbar, parity);           // LIDs of tile t are in shared memory
      parity ^= 1u;
    }

    // ---------------- phase 1: one thread per tile cell
    for (int j = tid; AFFINE && j < ncell; j += TR) {
      int lid[8];
      {
        const int4 *p = reinterpret_cast<const int4 *>(lidbuf + j * 8);
        const int4 v0 = p[0], v1 = p[1];
        lid[0] = v0.x; lid[1] = v0.y; lid[2] = v0.z; lid[3] = v0.w;
        lid[4] = v1.x; lid[5] = v1.y; lid[6] = v1.z; lid[7] = v1.w;
      }
      const int64_t e = need_cell ? T.tile_cells[cb + j] : 0;
      if (AFFINE) {
        double X[4][3], ug[8];
        constexpr int vn[4] = {0, 1, 3, 4};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int64_t l = lid[vn[k]];
          X[k][0] = __ldg(A.xyz + l * 3); X[k][1] = __ldg(A.xyz + l * 3 + 1); X[k][2] = __ldg(A.xyz + l * 3 + 2);
        }
        if (one_g) {
#pragma unroll
          for (int n = 0; n < 8; ++n) ug[n] = kgv * __ldg(xg + lid[n]);
        } else {
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const int64_t l = lid[n];
            double g = 0.0;
#pragma unroll
            for (int v = 0; v < 3; ++v)
              if (A.c.kg[v] != 0.0) g = fma(A.c.kg[v], __ldg(A.x[v] + l), g);
            ug[n] = g;
          }
        }
        stage_affine<TEP>(sm, j, e, X[0], X[1], X[2], X[3], ug, A.c, has_mass, has_src, !cong || j == 0);
        if (has_mass) {                   // mass pass: the combined solution the MASS integrands see
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const int64_t l = lid[n];
            double m = 0.0;
#pragma unroll
            for (int v = 0; v < 3; ++v)
              if (A.c.km[v] != 0.0) m = fma(A.c.km[v], __ldg(A.x[v] + l), m);
            st1<TEP>(sm, 26 + n, j, m);
          }
        }
      }
      // (general hexahedra: nothing to stage -- phase 2 reads the element records of k_elem_general directly)
    }
    // this row's cell table (one 16-byte load) and id: in flight across the barrier
    const uint4 alv = AFFINE ? __ldg(reinterpret_cast<const uint4 *>(T.adjl + slot * 8)) : make_uint4(0u, 0u, 0u, 0u);
    const int row = T.tile_rows[slot];
    // congruent tile: the interior stiffness row (same address for every thread), also in flight across the barrier
    const bool use_kf = AFFINE && cong && !has_mass;
    double acc[27];
#pragma unroll
    for (int c = 0; c < 27; ++c) acc[c] = use_kf ? __ldg(T.tile_kf + (int64_t)t * KF_STRIDE + c) : 0.0;
    const bool img_ok = AFFINE && JAC && use_kf && T.tma_store != 0;   // uniform runs leave from the constant image
    const bool uni = img_ok && (tcls & 4);                             // ... and the tile has nothing else
    double kfv = 0.0;
    bool stale = false;
    if (img_ok && tid < 27) {
      kfv = A.c.cK * __ldg(T.tile_kf + (int64_t)t * KF_STRIDE + tid);
      stale = !(kfv == kfc[tid]);
    }
    const int rebuild = (AFFINE && JAC) ? __syncthreads_or(stale) : (__syncthreads(), 0);   // staging complete; lidbuf free
    if (AFFINE && tid == 0 && tn < T.n_tiles) bulk_load(lidbuf_s, T.tile_lids + cbn * 8, (unsigned)ncelln * 32u, mbar);
    if (rebuild) {                       // (first tile of the CTA, or the cell shape changed)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // stores still reading the old image
      __syncthreads();
      for (int i = tid; i < IMG_DOUBLES; i += TR) img[i] = A.c.cK * __ldg(T.tile_kf + (int64_t)t * KF_STRIDE + i % 27);
      if (tid < 27) kfc[tid] = kfv;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }
    // this thread's run (runs are dealt round-robin over the warps, see the store loop): in flight during phase 2
    const int my_run = (tid & 31) * (TR / 32) + (tid >> 5);
    RowRun rr0{0, 0, 0};
    if (JAC && my_run < nrun) rr0 = T.runs[rb + my_run];
    const unsigned alw[4] = {alv.x, alv.y, alv.z, alv.w};
    unsigned rinfo = 0xFFFFu;
    uint4 p0 = make_uint4(~0u, ~0u, ~0u, ~0u), p1 = p0;
    int dir_i = -1;                      // fused Dirichlet row: index of its value (consumed after phase 2)
    if (T.row_dir && row >= 0) dir_i = __ldg(T.row_dir + row);
    if (JAC) {                           // consumed in phase 3; in flight during phase 2
      rinfo = T.tile_rowinfo[slot];
      const uint4 *pp = reinterpret_cast<const uint4 *>(T.tile_perm + slot * PERM_STRIDE);
      p0 = __ldg(pp); p1 = __ldg(pp + 1);
    }

    int gnext[8];
    if (!AFFINE) {                       // next tile's cell ids (its table row arrived during the barrier above)
      const unsigned w[4] = {alv_n.x, alv_n.y, alv_n.z, alv_n.w};
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int el = (int)((w[a >> 1] >> (16 * (a & 1))) & 0xFFFFu);
        gnext[a] = (el != 0xFFFF) ? __ldg(T.tile_cells + cbn + el) : -1;
      }
    }

    // ---------------- phase 2: one thread per row, 27 entries in registers
    double fr = 0.0;
    const bool full = ((alw[0] & 0xFFFFu) != 0xFFFFu) & ((alw[0] >> 16) != 0xFFFFu) & ((alw[1] & 0xFFFFu) != 0xFFFFu) &
                      ((alw[1] >> 16) != 0xFFFFu) & ((alw[2] & 0xFFFFu) != 0xFFFFu) & ((alw[2] >> 16) != 0xFFFFu) &
                      ((alw[3] & 0xFFFFu) != 0xFFFFu) & ((alw[3] >> 16) != 0xFFFFu);
#define TX_EL(AA) ((int)((alw[(AA) >> 1] >> (16 * ((AA) & 1))) & 0xFFFFu))
    if (use_kf && full) {
      // interior row of a congruent tile: f = sum_j Kf[j] ug_j + sources; A row = cK Kf.  Neighbour j is vertex
      // nb_vert(j) of the cell in which this row is vertex nb_cell(j).
#define TX_NB(J) fr = fma(acc[J], ld1<TEP>(sm, 14 + nb_vert(J), TX_EL(nb_cell(J))), fr);
      TX_NB(0) TX_NB(1) TX_NB(2) TX_NB(3) TX_NB(4) TX_NB(5) TX_NB(6) TX_NB(7) TX_NB(8) TX_NB(9) TX_NB(10) TX_NB(11) TX_NB(12)
      TX_NB(13) TX_NB(14) TX_NB(15) TX_NB(16) TX_NB(17) TX_NB(18) TX_NB(19) TX_NB(20) TX_NB(21) TX_NB(22) TX_NB(23) TX_NB(24)
      TX_NB(25) TX_NB(26)
#undef TX_NB
      if (has_src) {
#define TX_SRC(AA) fr += ld1<TEP>(sm, 22 + (AA), TX_EL(AA));
        TX_SRC(0) TX_SRC(1) TX_SRC(2) TX_SRC(3) TX_SRC(4) TX_SRC(5) TX_SRC(6) TX_SRC(7)
#undef TX_SRC
      }
      if (JAC && !(img_ok && (rinfo & ROW_UNIFORM))) {
#pragma unroll
        for (int c = 0; c < 27; ++c) acc[c] *= A.c.cK;
      }
    } else {
    if (use_kf) {
#pragma unroll
      for (int c = 0; c < 27; ++c) acc[c] = 0.0;
    }
#define TX_ROW(AA)                                                                                   \
    { if (AFFINE) {                                                                                  \
        const int el = (int)((alw[(AA) >> 1] >> (16 * ((AA) & 1))) & 0xFFFFu);                       \
        if (el != 0xFFFF) row_accum_affine<TEP, AA, JAC>(sm, el, cong ? 0 : el, A.c, has_mass, has_src, acc, fr);\
      } else if (gcell[AA] >= 0) row_accum_general<AA, JAC>(A.elem, (int64_t)gcell[AA], acc, fr);    \
    }
    TX_ROW(0) TX_ROW(1) TX_ROW(2) TX_ROW(3) TX_ROW(4) TX_ROW(5) TX_ROW(6) TX_ROW(7)
    }
#undef TX_ROW
#undef TX_EL
    if (dir_i >= 0) {
      // TianXin Dirichlet row, fused (TianXin_Dirichlet_impl.hpp:59-81 -> applyDirichletBoundaryConditionToLocalMatrixRows +
      // evalDirichletResidual, lof/Panzer_TpetraLinearObjContainer.hpp:228-237,306-317): row := identity, f = x - value
      fr = __ldg(A.x[0] + row) - __ldg(T.dir_vals + dir_i);
      if (JAC) {
#pragma unroll
        for (int c = 0; c < 27; ++c) acc[c] = (c == 13) ? 1.0 : 0.0;
      }
    }
    if (row >= 0 && A.f) A.f[row] = fr;

    if (JAC) {
      // ---------------- phase 3: permute to CSR slot order in shared memory, then TMA bulk stores of row runs
      double *out = sm;
      if (!uni) {
      __syncthreads();                   // staging is dead; reuse it as the out buffer
      const int my_len = (int)((rinfo >> 16) & 0xFFu);
      if (row >= 0 && !(img_ok && (rinfo & ROW_UNIFORM))) {
        double *o = out + (rinfo & 0xFFFFu);
        if ((rinfo >> 24) & 1u)          // the row has slots no local cell writes (zero them)
          for (int s = 0; s < my_len; ++s) o[s] = 0.0;
        const unsigned w[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
        for (int c = 0; c < 27; ++c) {
          const unsigned p = (w[c >> 2] >> (8 * (c & 3))) & 0xFFu;
          if (p != 0xFFu) o[p] = acc[c];
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my writes -> visible to the TMA engine
      __syncthreads();
      }
      {
        const unsigned out_s = (unsigned)__cvta_generic_to_shared(out);
        // a bulk copy takes warp-uniform operands, so a warp issues its lanes' copies one after the other: deal the
        // runs round-robin over the warps (run i -> warp i % 8) instead of giving the first 32 to warp 0
        for (int i = my_run; i < nrun; i += TR) {
          RowRun rr = (i == my_run) ? rr0 : T.runs[rb + i];
          double *g = A.A + rr.beg;
          const double *so = out + rr.soff;
          const bool from_img = img_ok && (rr.n & RUN_UNIFORM);
          rr.n &= ~RUN_UNIFORM;
          if (from_img) {
            // the run is the period-27 pattern: element k of the run is img[k % 27].  The 16-byte aligned middle
            // starts at pattern phase `head`; img + 28 has phase 1, img + 0 phase 0; 216 = 8 rows keeps the phase.
            const int head = (int)(rr.beg & 1);
            const int mid = (rr.n - head) & ~1;
            if (head) g[0] = img[0];
            if (rr.n - head - mid) g[rr.n - 1] = img[(rr.n - 1) % 27];
            const unsigned src = img_s + (head ? 28u * 8u : 0u);
            for (int o = 0; o < mid; o += 216) {
              const int m = (mid - o < 216) ? mid - o : 216;
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                           ::"l"(g + head + o), "r"(src), "r"((unsigned)m * 8u) : "memory");
            }
          } else if (T.tma_store) {
            // 16-byte aligned middle by one bulk store; at most one leading and one trailing element by hand
            const int head = (int)(rr.beg & 1);
            const int mid = (rr.n - head) & ~1;
            if (head) g[0] = so[0];
            if (rr.n - head - mid) g[rr.n - 1] = so[rr.n - 1];
            if (mid)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                           ::"l"(g + head), "r"(out_s + (unsigned)(rr.soff + head) * 8u), "r"((unsigned)mid * 8u) : "memory");
          } else {
            for (int k = 0; k < rr.n; ++k) g[k] = so[k];               // A_values not 16-byte aligned: plain copy
          }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }

    // ---------------- warm L2 with the row tables of the next tile
    if (tn < T.n_tiles) {
      {                                  // its row tables: 16 + 8 + 32 bytes per row, contiguous per tile
        const int64_t sn = (int64_t)tn * TR;
        if (tid < TR / 8) prefetch_l2(T.adjl + (sn + tid * 8) * 8);
        if (tid < TR / 32) prefetch_l2(T.tile_rowinfo + sn + tid * 32);
        if (tid < TR / 4) prefetch_l2(T.tile_perm + (sn + tid * 4) * PERM_STRIDE);
        if (tid < TR / 32) prefetch_l2(T.tile_rows + sn + tid * 32);
      }
    }
    if (JAC && !uni) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();                     // staging (and the out buffer) dead before the next tile stages into it
    cb = cbn; ncell = ncelln;
    if (!AFFINE) {
#pragma unroll
      for (int a = 0; a < 8; ++a) gcell[a] = gnext[a];
    }
  }
  if (JAC) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // image-sourced stores may still be reading
  tx_stamp(A.dbg, 2, true);
}

// Source load vector of one closure model for a cell in general position (or a per-IP array): the 8 point values,
// then sum factorisation of  b_a = det * sum_q N_a(xi_q) s_q.  Out of line: sinpi() would crowd the fast path.
// (scalars by value: a reference to the kernel's FillCoef would force a per-thread copy of it into local memory)
__device__ __noinline__ void source_load_general(int id, double mult, const double *__restrict__ ip, int64_t e,
                                                 double xc0, double xc1, double xc2,
                                                 double j00, double j01, double j02, double j10, double j11, double j12,
                                                 double j20, double j21, double j22, double det, double *__restrict__ bl)
{
  constexpr double wl = 0.5 * (1.0 - TX_INV_SQRT3), wh = 0.5 * (1.0 + TX_INV_SQRT3);
  double sq8[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (id == TXASM_SOURCE_IP_ARRAY) sq8[q] = mult * ip[e * 8 + q];
    else {
      const double xi = (q & 1) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
      const double et = (q & 2) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
      const double ze = (q & 4) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
      sq8[q] = mult * source_eval(id, xc0 + j00 * xi + j01 * et + j02 * ze, xc1 + j10 * xi + j11 * et + j12 * ze,
                                  xc2 + j20 * xi + j21 * et + j22 * ze);
    }
  }
  double tx[2][4];   // [sx][qy,qz]
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    tx[0][k] = wh * sq8[2 * k] + wl * sq8[2 * k + 1];
    tx[1][k] = wl * sq8[2 * k] + wh * sq8[2 * k + 1];
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int ix = hex_sx(a) > 0, iy = hex_sy(a) > 0, iz = hex_sz(a) > 0;
    const double y0 = iy ? (wl * tx[ix][0] + wh * tx[ix][1]) : (wh * tx[ix][0] + wl * tx[ix][1]);
    const double y1 = iy ? (wl * tx[ix][2] + wh * tx[ix][3]) : (wh * tx[ix][2] + wl * tx[ix][3]);
    bl[a] = det * (iz ? (wl * y0 + wh * y1) : (wh * y0 + wl * y1));
  }
}

// ============================================================================ uniform tiles
// Tiles [0, n_uni): congruent cells, every row interior with canonical column order (class 2, see k_tile_runs).  All
// of A for such a tile is the constant row image, so the kernel has no metric staging, no 27 accumulators, no
// placement and no row tables beyond the cell table: 16 doubles of staging per cell (gathered u, source load) and
// <= 80 registers -> 3 CTAs of 256 threads per SM.  Jacobian-type evaluations without mass terms only; anything
// else goes through k_fill_rowtile for all tiles.
template <int TEP>
__global__ void __launch_bounds__(256, 3) k_fill_uniform(FillArgs A, TileArgs T)
{
  constexpr int TR = 256;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);                     // 8 double2 slots per cell, stride TEP
  int *lidbuf = reinterpret_cast<int *>(smem_raw + TEP * 128);
  const unsigned lidbuf_s = (unsigned)__cvta_generic_to_shared(lidbuf);
  const unsigned mbar = lidbuf_s + TEP * 32;
  double *img = reinterpret_cast<double *>(smem_raw + TEP * 128 + TEP * 32 + 16);
  double *kfc = img + IMG_DOUBLES;       // cK*Kf the image holds
  double *kfu = kfc + 28;                // Kf itself (residual)
  double *geo = kfu + 28;                // Jxx Jyy Jzz det of the cells (axis-aligned translates of one box)
  const unsigned img_s = mbar + 16;
  const int tid = threadIdx.x, G = gridDim.x;
  const bool has_src = A.c.n_src > 0;
  bool need_cell = false;
  for (int s = 0; s < A.c.n_src; ++s) need_cell |= (A.c.src_id[s] == TXASM_SOURCE_IP_ARRAY);

  int t = T.t_begin + (int)blockIdx.x;
  int64_t cb = T.tile_cell_ptr[t];
  int ncell = (int)(T.tile_cell_ptr[t + 1] - cb);
  if (tid == 0) {
    mbar_init(mbar, 1);
    bulk_load(lidbuf_s, T.tile_lids + cb * 8, (unsigned)ncell * 32u, mbar);
  }
  if (tid < 28) kfc[tid] = __longlong_as_double(0x7ff8000000000000LL);   // no image yet
  __syncthreads();
  unsigned parity = 0;
  // Barrier that also makes the per-shape constants valid for `tile`: row image cK*Kf, Kf, box geometry.  They are
  // rebuilt when cK*Kf of that tile differs from what the image holds (first tile of the CTA, change of cell shape;
  // Kf determines the box).  Called before a tile's phase 1: in the prologue and as each tile's closing barrier.
  auto ensure = [&](int tile) {
    double kfv = 0.0, kf0 = 0.0;
    bool stale = false;
    if (tile < T.n_tiles && tid < 27) {
      kf0 = __ldg(T.tile_kf + (int64_t)tile * KF_STRIDE + tid);
      kfv = A.c.cK * kf0;
      stale = !(kfv == kfc[tid]);
    }
    if (__syncthreads_or(stale)) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // stores still reading the old image
      __syncthreads();
      for (int i = tid; i < IMG_DOUBLES; i += TR) img[i] = A.c.cK * __ldg(T.tile_kf + (int64_t)tile * KF_STRIDE + i % 27);
      if (tid < 27) { kfc[tid] = kfv; kfu[tid] = kf0; }
      if (tid >= 32 && tid < 36) geo[tid - 32] = __ldg(T.tile_kf + (int64_t)tile * KF_STRIDE + 27 + (tid - 32));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }
  };
  ensure(t);

  for (; t < T.n_tiles; t += G) {        // (n_tiles = n_uni for this launch)
    const int tn = t + G;
    int64_t cbn = 0;
    int ncelln = 0;
    if (tn < T.n_tiles) { cbn = T.tile_cell_ptr[tn]; ncelln = (int)(T.tile_cell_ptr[tn + 1] - cbn); }
    const int64_t slot = (int64_t)t * TR + tid;
    const int64_t rb = T.run_ptr[t];
    const int nrun = (int)(T.run_ptr[t + 1] - rb);

    mbar_wait(mbar, parity);
    parity ^= 1u;

    // ---------------- phase 1: gathered solution and source load vector of every tile cell
    for (int j = tid; j < ncell; j += TR) {
      int lid[8];
      {
        const int4 *p = reinterpret_cast<const int4 *>(lidbuf + j * 8);
        const int4 v0 = p[0], v1 = p[1];
        lid[0] = v0.x; lid[1] = v0.y; lid[2] = v0.z; lid[3] = v0.w;
        lid[4] = v1.x; lid[5] = v1.y; lid[6] = v1.z; lid[7] = v1.w;
      }
      {
        double ug[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          double g = 0.0;
#pragma unroll
          for (int v = 0; v < 3; ++v)
            if (A.c.kg[v] != 0.0) g = fma(A.c.kg[v], __ldg(A.x[v] + lid[n]), g);
          ug[n] = g;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) st2<TEP>(sm, q, j, ug[2 * q], ug[2 * q + 1]);
      }
      if (has_src) {
        // cell = cell 0 of the tile translated: one vertex is gathered, the box comes from the setup table
        const double hx = geo[0], hy = geo[1], hz = geo[2], det = geo[3];
        double xc[3];
        {
          const double *p0 = A.xyz + (int64_t)lid[0] * 3;
          xc[0] = __ldg(p0) + hx; xc[1] = __ldg(p0 + 1) + hy; xc[2] = __ldg(p0 + 2) + hz;
        }
        constexpr double wl = 0.5 * (1.0 - TX_INV_SQRT3), wh = 0.5 * (1.0 + TX_INV_SQRT3);
        double bl[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) bl[a] = 0.0;
#pragma unroll 1
        for (int s = 0; s < A.c.n_src; ++s) {
          if (A.c.src_id[s] == TXASM_SOURCE_SIN3) {
            // separable model on an axis-aligned cell (all cells of these tiles are): b_a = det X_a Y_a Z_a
            const double dx = hx * TX_INV_SQRT3, dy = hy * TX_INV_SQRT3, dz = hz * TX_INV_SQRT3;
            double f0 = sin2pi_fast(xc[0] - dx), f1 = sin2pi_fast(xc[0] + dx);
            const double Xm = wh * f0 + wl * f1, Xp = wl * f0 + wh * f1;
            f0 = sin2pi_fast(xc[1] - dy); f1 = sin2pi_fast(xc[1] + dy);
            const double Ym = wh * f0 + wl * f1, Yp = wl * f0 + wh * f1;
            const double cz = A.c.src_mult[s] * 118.43525281307230 * det;
            f0 = cz * sin2pi_fast(xc[2] - dz); f1 = cz * sin2pi_fast(xc[2] + dz);
            const double Zm = wh * f0 + wl * f1, Zp = wl * f0 + wh * f1;
            const double xy[2][2] = {{Xm * Ym, Xm * Yp}, {Xp * Ym, Xp * Yp}};
#pragma unroll
            for (int a = 0; a < 8; ++a) bl[a] = fma(xy[hex_sx(a) > 0][hex_sy(a) > 0], hex_sz(a) > 0 ? Zp : Zm, bl[a]);
          } else {                         // TXASM_SOURCE_CONSTANT (the launch admits nothing else): b_a = mult * det
            const double v = A.c.src_mult[s] * det;
#pragma unroll
            for (int a = 0; a < 8; ++a) bl[a] += v;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) st2<TEP>(sm, 4 + q, j, bl[2 * q], bl[2 * q + 1]);
      }
    }
    const uint4 alv = __ldg(reinterpret_cast<const uint4 *>(T.adjl + slot * 8));
    const int row = T.tile_rows[slot];
    __syncthreads();                     // staging complete; lidbuf free
    if (tid == 0 && tn < T.n_tiles) bulk_load(lidbuf_s, T.tile_lids + cbn * 8, (unsigned)ncelln * 32u, mbar);
    const int my_run = (tid & 31) * (TR / 32) + (tid >> 5);
    RowRun rr0{0, 0, 0};
    if (my_run < nrun) rr0 = T.runs[rb + my_run];
    const unsigned alw[4] = {alv.x, alv.y, alv.z, alv.w};

    // ---------------- phase 2: f = sum_j Kf[j] u_j + sources (every row of the tile has its 8 cells)
    if (row >= 0 && A.f) {
      double fr = 0.0;
#define TX_EL(AA) ((int)((alw[(AA) >> 1] >> (16 * ((AA) & 1))) & 0xFFFFu))
#define TX_NB(J) fr = fma(kfu[J], ld1<TEP>(sm, nb_vert(J), TX_EL(nb_cell(J))), fr);
      TX_NB(0) TX_NB(1) TX_NB(2) TX_NB(3) TX_NB(4) TX_NB(5) TX_NB(6) TX_NB(7) TX_NB(8) TX_NB(9) TX_NB(10) TX_NB(11) TX_NB(12)
      TX_NB(13) TX_NB(14) TX_NB(15) TX_NB(16) TX_NB(17) TX_NB(18) TX_NB(19) TX_NB(20) TX_NB(21) TX_NB(22) TX_NB(23) TX_NB(24)
      TX_NB(25) TX_NB(26)
#undef TX_NB
      if (has_src) {
#define TX_SRC(AA) fr += ld1<TEP>(sm, 8 + (AA), TX_EL(AA));
        TX_SRC(0) TX_SRC(1) TX_SRC(2) TX_SRC(3) TX_SRC(4) TX_SRC(5) TX_SRC(6) TX_SRC(7)
#undef TX_SRC
      }
#undef TX_EL
      A.f[row] = fr;
    }

    // ---------------- A: every run straight from the constant image
    for (int i = my_run; i < nrun; i += TR) {
      RowRun rr = (i == my_run) ? rr0 : T.runs[rb + i];
      rr.n &= ~RUN_UNIFORM;
      double *g = A.A + rr.beg;
      const int head = (int)(rr.beg & 1);
      const int mid = (rr.n - head) & ~1;
      if (head) g[0] = img[0];
      if (rr.n - head - mid) g[rr.n - 1] = img[(rr.n - 1) % 27];
      const unsigned src = img_s + (head ? 28u * 8u : 0u);
      for (int o = 0; o < mid; o += 216) {
        const int m = (mid - o < 216) ? mid - o : 216;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(g + head + o), "r"(src), "r"((unsigned)m * 8u) : "memory");
      }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (tn < T.n_tiles) {                // warm L2 with the cell table and row ids of the next tile
      const int64_t sn = (int64_t)tn * TR;
      if (tid < TR / 8) prefetch_l2(T.adjl + (sn + tid * 8) * 8);
      if (tid < TR / 32) prefetch_l2(T.tile_rows + sn + tid * 32);
    }
    ensure(tn);                          // staging dead before the next tile writes it; constants valid for it
    cb = cbn; ncell = ncelln;
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // stores may still be reading the image
}

// ============================================================================ host side
template <class T>
static void free_dev(txasm_handle h, T *&p) { if (p) { dev_free(h, (void *)p); p = nullptr; } }

void tiles_free(txasm_handle h)
{
  if (!h->tiles) return;
  Tiles *T = h->tiles;
  brick_free(h);
  free_dev(h, T->d_tile_rows); free_dev(h, T->d_tile_cell_ptr); free_dev(h, T->d_tile_cells); free_dev(h, T->d_tile_lids);
  free_dev(h, T->d_adjl); free_dev(h, T->d_perm); free_dev(h, T->d_irregular);
  free_dev(h, T->d_tile_perm); free_dev(h, T->d_tile_cong); free_dev(h, T->d_tile_kf); free_dev(h, T->d_row_uniform);
  free_dev(h, T->d_tile_rowinfo); free_dev(h, T->d_run_ptr); free_dev(h, T->d_runs);
  delete T;
  h->tiles = nullptr;
}

template <int TR>
static int build_cells(txasm_handle h, Tiles *T, const int *adjcell)
{
  const int cap = TR * 8;
  int *ncells = nullptr, *tmp = nullptr;
  int64_t *ncells64 = nullptr;
  TX_CUDA(h, cudaMalloc(&ncells, sizeof(int) * (T->n_tiles + 1)));
  TX_CUDA(h, cudaMalloc(&ncells64, sizeof(int64_t) * (T->n_tiles + 1)));
  TX_CUDA(h, cudaMalloc(&tmp, sizeof(int) * (size_t)T->n_tiles * cap));
  TX_CUDA(h, cudaMemsetAsync(ncells, 0, sizeof(int) * (T->n_tiles + 1), h->stream));
  int rc = dev_alloc(h, &T->d_adjl, (size_t)T->n_tiles * 8 * TR);
  if (rc) return rc;
  k_tile_cells<TR><<<T->n_tiles, TR, 0, h->stream>>>(T->d_tile_rows, adjcell, ncells, tmp, T->d_adjl);
  TX_CUDA(h, cudaGetLastError());
  // max and prefix sum of the per-tile counts
  {
    size_t tb = 0; int *dmax = nullptr;
    TX_CUDA(h, cudaMalloc(&dmax, sizeof(int)));
    cub::DeviceReduce::Max(nullptr, tb, ncells, dmax, T->n_tiles, h->stream);
    void *t2 = nullptr; TX_CUDA(h, cudaMalloc(&t2, tb ? tb : 1));
    cub::DeviceReduce::Max(t2, tb, ncells, dmax, T->n_tiles, h->stream);
    TX_CUDA(h, cudaMemcpyAsync(&T->te_max, dmax, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    TX_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(t2); cudaFree(dmax);
  }
  {
    std::vector<int> hc(T->n_tiles + 1);
    TX_CUDA(h, copy_to_device_sync(h, hc.data(), ncells, sizeof(int) * (T->n_tiles + 1)));
    std::vector<int64_t> hp(T->n_tiles + 1);
    int64_t s = 0;
    for (int i = 0; i < T->n_tiles; ++i) { hp[i] = s; s += hc[i]; }
    hp[T->n_tiles] = s;
    rc = dev_alloc(h, &T->d_tile_cell_ptr, (size_t)T->n_tiles + 1);
    if (rc) return rc;
    TX_CUDA(h, copy_to_device_sync(h, T->d_tile_cell_ptr, hp.data(), sizeof(int64_t) * (T->n_tiles + 1)));
    rc = dev_alloc(h, &T->d_tile_cells, (size_t)s);
    if (rc) return rc;
    k_compact_cells<<<T->n_tiles, 128, 0, h->stream>>>(T->n_tiles, cap, ncells, T->d_tile_cell_ptr, tmp, T->d_tile_cells);
    TX_CUDA(h, cudaGetLastError());
    rc = dev_alloc(h, &T->d_tile_lids, (size_t)s * 8);
    if (rc) return rc;
    if (s) k_tile_lids<<<(unsigned)((s * 8 + 255) / 256), 256, 0, h->stream>>>(s, T->d_tile_cells, h->d_lids, T->d_tile_lids);
    TX_CUDA(h, cudaGetLastError());
    free_dev(h, T->d_tile_cong);
    rc = dev_alloc(h, &T->d_tile_cong, (size_t)T->n_tiles);
    if (rc) return rc;
    {
      const char *e = getenv("TXASM_NO_CONGRUENT");
      if (e && e[0] == '1') { TX_CUDA(h, cudaMemsetAsync(T->d_tile_cong, 0, (size_t)T->n_tiles, h->stream)); }
      else k_tile_congruent<<<T->n_tiles, 128, 0, h->stream>>>(T->n_tiles, T->d_tile_cell_ptr, T->d_tile_cells, h->d_lids, h->d_xyz,
                                                             h->d_cell_affine, h->cfg.affine_tol == 0.0 ? 1e-13 : h->cfg.affine_tol,
                                                             T->d_tile_cong);
      TX_CUDA(h, cudaGetLastError());
    }
    free_dev(h, T->d_tile_kf);
    rc = dev_alloc(h, &T->d_tile_kf, (size_t)T->n_tiles * KF_STRIDE);
    if (rc) return rc;
    k_tile_kf<<<(T->n_tiles + 127) / 128, 128, 0, h->stream>>>(T->n_tiles, T->d_tile_cell_ptr, T->d_tile_cells, h->d_lids, h->d_xyz,
                                                              T->d_tile_cong, T->d_tile_kf);
    TX_CUDA(h, cudaGetLastError());
    // are all tile cells affine?
    int *d_non = nullptr, non = 0;
    TX_CUDA(h, cudaMalloc(&d_non, sizeof(int)));
    TX_CUDA(h, cudaMemsetAsync(d_non, 0, sizeof(int), h->stream));
    if (s) k_tile_affine<<<(unsigned)((s + 255) / 256), 256, 0, h->stream>>>(s, T->d_tile_cells, h->d_cell_affine, d_non);
    TX_CUDA(h, cudaMemcpyAsync(&non, d_non, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    TX_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(d_non);
    T->all_affine = (non == 0);
  }
  cudaFree(ncells); cudaFree(ncells64); cudaFree(tmp);
  return TXASM_OK;
}

// kernel instantiations: (TR, TEP) pairs per variant; the smallest TEP >= te_max is used
typedef void (*TileKernel)(FillArgs, TileArgs);
struct KernelChoice { int TR, TEP; bool affine; TileKernel jac, res; };
#define TX_KC(TRv, TEPv, AFF) {TRv, TEPv, AFF, k_fill_rowtile<TRv, TEPv, AFF, true>, k_fill_rowtile<TRv, TEPv, AFF, false>}
static const KernelChoice g_kernels[] = {
  TX_KC(256, 416, true), TX_KC(256, 448, true), TX_KC(256, 512, true), TX_KC(256, 640, true), TX_KC(128, 256, true), TX_KC(128, 384, true),
  TX_KC(128, 256, false), TX_KC(128, 288, false), TX_KC(128, 320, false), TX_KC(128, 384, false), TX_KC(128, 512, false),
};
#undef TX_KC

static const KernelChoice *pick_kernel(int TR, bool affine, int te_max)
{
  const KernelChoice *best = nullptr;
  for (const KernelChoice &k : g_kernels)
    if (k.TR == TR && k.affine == affine && k.TEP >= te_max && (!best || k.TEP < best->TEP)) best = &k;
  return best;
}

static int smem_need(const Tiles *T, bool affine, int TR, bool mass = true, bool src = true)
{
  // setup sizes for the worst case (mass + source on); a launch asks for what its term list needs
  const int per_cell = stage_doubles(affine, mass, src);
  const int stage = affine ? per_cell * T->tep * 8 : 0;   // general hexahedra: nothing is staged (records are read directly)
  const int out = T->out_doubles * 8 + 16;
  return (std::max(stage, out) + 15) & ~15;          // the staging / out region
}
// + LID buffer + mbarrier + (affine) constant row image
static int smem_total(const Tiles *T, int stage_bytes) { return stage_bytes + T->tep * 32 + 16 + (T->all_affine ? IMG_BYTES : 0); }

int tiles_build(txasm_handle h)
{
  tiles_free(h);
  const int64_t nr = h->n_rows;
  Tiles *T = new Tiles();
  h->tiles = T;
  int rc;
  // 1. regular rows, perm table, per-row cell table
  unsigned char *regular = nullptr;
  int *adjcell = nullptr, *d_maxlen = nullptr;
  TX_CUDA(h, cudaMalloc(&regular, (size_t)nr));
  TX_CUDA(h, cudaMalloc(&adjcell, sizeof(int) * (size_t)nr * 8));
  TX_CUDA(h, cudaMalloc(&d_maxlen, sizeof(int)));
  TX_CUDA(h, cudaMemsetAsync(d_maxlen, 0, sizeof(int), h->stream));
  if ((rc = dev_alloc(h, &T->d_perm, (size_t)nr * PERM_STRIDE))) return rc;
  k_row_regular<<<(unsigned)((nr + 127) / 128), 128, 0, h->stream>>>(nr, h->d_adj_ptr, h->d_adj, h->d_lids, h->d_rowptr,
                                                                    h->d_colind, regular, T->d_perm, adjcell, d_maxlen);
  TX_CUDA(h, cudaGetLastError());
  TX_CUDA(h, cudaMemcpyAsync(&T->lrow, d_maxlen, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  // irregular rows
  int *d_cnt = nullptr;
  TX_CUDA(h, cudaMalloc(&d_cnt, sizeof(int)));
  TX_CUDA(h, cudaMemsetAsync(d_cnt, 0, sizeof(int), h->stream));
  int *irr_tmp = nullptr;
  TX_CUDA(h, cudaMalloc(&irr_tmp, sizeof(int) * (size_t)nr));
  k_list_irregular<<<(unsigned)((nr + 255) / 256), 256, 0, h->stream>>>(nr, regular, irr_tmp, d_cnt);
  int n_irr = 0;
  TX_CUDA(h, cudaMemcpyAsync(&n_irr, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_cnt); cudaFree(d_maxlen);
  T->n_irregular = n_irr; T->n_regular = nr - n_irr;
  if (T->lrow < 27) T->lrow = 27;
  if ((T->lrow & 1) == 0) T->lrow += 1;            // odd stride: conflict-free row-strided shared stores
  if (n_irr) {
    if ((rc = dev_alloc(h, &T->d_irregular, (size_t)n_irr))) return rc;
    // deterministic order
    std::vector<int> hi(n_irr);
    TX_CUDA(h, copy_to_device_sync(h, hi.data(), irr_tmp, sizeof(int) * n_irr));
    std::sort(hi.begin(), hi.end());
    TX_CUDA(h, copy_to_device_sync(h, T->d_irregular, hi.data(), sizeof(int) * n_irr));
  }
  cudaFree(irr_tmp);
  if (T->n_regular == 0) { cudaFree(regular); cudaFree(adjcell); tiles_free(h); return set_err(h, TXASM_EUNSUPPORTED, "no regular rows"); }

  // 2. Morton order of the regular rows
  unsigned long long *bb = nullptr, *keys = nullptr, *keys2 = nullptr;
  int *vals = nullptr, *vals2 = nullptr;
  TX_CUDA(h, cudaMalloc(&bb, sizeof(unsigned long long) * 9));
  {
    unsigned long long init[9] = {~0ull, ~0ull, ~0ull, 0, 0, 0, ~0ull, ~0ull, ~0ull};
    TX_CUDA(h, copy_to_device_sync(h, bb, init, sizeof(init)));
  }
  TX_CUDA(h, cudaMalloc(&keys, sizeof(unsigned long long) * (size_t)nr));
  TX_CUDA(h, cudaMalloc(&keys2, sizeof(unsigned long long) * (size_t)nr));
  TX_CUDA(h, cudaMalloc(&vals, sizeof(int) * (size_t)nr));
  TX_CUDA(h, cudaMalloc(&vals2, sizeof(int) * (size_t)nr));
  k_bbox<<<h->n_sm * 8, 256, 0, h->stream>>>(nr, h->d_xyz, bb, bb + 3);
  k_cell_hmin<<<h->n_sm * 8, 256, 0, h->stream>>>(h->n_cells, h->d_lids, h->d_xyz, bb + 6);
  k_morton<<<(unsigned)((nr + 255) / 256), 256, 0, h->stream>>>(nr, h->d_xyz, bb, bb + 3, bb + 6, regular, keys, vals);
  {
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, vals, vals2, (int)nr, 0, 64, h->stream);
    void *tmp = nullptr; TX_CUDA(h, cudaMalloc(&tmp, tb ? tb : 1));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, vals, vals2, (int)nr, 0, 64, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    TX_CUDA(h, e);
  }
  int max_level = 63;
  {                                                   // the quantum of k_morton, on the host: 2^k quanta per cell -> boxes of <= 16 nodes per axis
    unsigned long long hb[9];
    TX_CUDA(h, copy_to_device_sync(h, hb, bb, sizeof(hb)));
    auto kd = [](unsigned long long k) {
      const unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
      double d; memcpy(&d, &u, 8); return d;
    };
    double cells_max = 0.0;
    for (int d = 0; d < 3; ++d) {
      const double ext = kd(hb[3 + d]) - kd(hb[d]), hd = kd(hb[6 + d]);
      if (hd > 0.0 && hd < 1e299) cells_max = std::max(cells_max, ext / hd); else cells_max = 1e300;
    }
    if (cells_max < 1048576.0) {
      int k = 0;
      double f = 1.0;
      while (cells_max * f * 2.0 < 2097151.0 && f < 4096.0) { f *= 2.0; ++k; }
      max_level = std::min(63, 3 * (k + 4));
    }
  }
  cudaFree(bb); cudaFree(keys); cudaFree(vals);       // keys2: sorted Morton keys, vals2: rows in that order

  // 3. tiles: TR rows each; try the large tile first, shrink if shared memory does not fit
  const int try_tr[2] = {256, 128};
  bool done = false;
  for (int attempt = 0; attempt < 2 && !done; ++attempt) {
    const int TR = try_tr[attempt];
    free_dev(h, T->d_tile_rows); free_dev(h, T->d_tile_cell_ptr); free_dev(h, T->d_tile_cells); free_dev(h, T->d_tile_lids); free_dev(h, T->d_adjl);
    T->TR = TR;
    {
      const int64_t nreg = T->n_regular;
      int *flag = nullptr, *tile_of = nullptr, *start = nullptr;
      TX_CUDA(h, cudaMalloc(&flag, sizeof(int) * (size_t)nreg));
      TX_CUDA(h, cudaMalloc(&tile_of, sizeof(int) * (size_t)nreg));
      k_leaf_starts<<<(unsigned)((nreg + 255) / 256), 256, 0, h->stream>>>(nreg, keys2, TR, max_level, flag);
      TX_CUDA(h, cudaGetLastError());
      {
        size_t tb = 0;
        cub::DeviceScan::InclusiveSum(nullptr, tb, flag, tile_of, (int)nreg, h->stream);
        void *tmp = nullptr; TX_CUDA(h, cudaMalloc(&tmp, tb ? tb : 1));
        cudaError_t e = cub::DeviceScan::InclusiveSum(tmp, tb, flag, tile_of, (int)nreg, h->stream);
        TX_CUDA(h, cudaMemcpyAsync(&T->n_tiles, tile_of + nreg - 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        cudaStreamSynchronize(h->stream);
        cudaFree(tmp);
        TX_CUDA(h, e);
      }
      const int64_t slots = (int64_t)T->n_tiles * TR;
      if (slots > 0x7fffffffLL) {
        cudaFree(flag); cudaFree(tile_of); cudaFree(vals2); cudaFree(keys2); cudaFree(regular); cudaFree(adjcell);
        tiles_free(h);
        return set_err(h, TXASM_EUNSUPPORTED, "too many tile slots");
      }
      TX_CUDA(h, cudaMalloc(&start, sizeof(int) * (size_t)T->n_tiles));
      k_tile_starts<<<(unsigned)((nreg + 255) / 256), 256, 0, h->stream>>>(nreg, flag, tile_of, start);
      if ((rc = dev_alloc(h, &T->d_tile_rows, (size_t)slots))) return rc;
      TX_CUDA(h, cudaMemsetAsync(T->d_tile_rows, 0xFF, sizeof(int) * (size_t)slots, h->stream));
      unsigned long long *k1 = nullptr, *k2 = nullptr;
      TX_CUDA(h, cudaMalloc(&k1, sizeof(unsigned long long) * (size_t)nreg));
      TX_CUDA(h, cudaMalloc(&k2, sizeof(unsigned long long) * (size_t)nreg));
      k_tile_keys<<<(unsigned)((nreg + 255) / 256), 256, 0, h->stream>>>(nreg, tile_of, vals2, k1);
      size_t tb = 0;
      cub::DeviceRadixSort::SortKeys(nullptr, tb, k1, k2, (int)nreg, 0, 64, h->stream);
      void *tmp = nullptr; TX_CUDA(h, cudaMalloc(&tmp, tb ? tb : 1));
      cudaError_t e = cub::DeviceRadixSort::SortKeys(tmp, tb, k1, k2, (int)nreg, 0, 64, h->stream);
      k_tile_rows_from_keys<<<(unsigned)((nreg + 255) / 256), 256, 0, h->stream>>>(nreg, TR, k2, start, T->d_tile_rows);
      cudaStreamSynchronize(h->stream);
      cudaFree(tmp); cudaFree(k1); cudaFree(k2); cudaFree(flag); cudaFree(tile_of); cudaFree(start);
      TX_CUDA(h, e);
      TX_CUDA(h, cudaGetLastError());
    }
    rc = (TR == 256) ? build_cells<256>(h, T, adjcell) : build_cells<128>(h, T, adjcell);
    if (rc) return rc;
    const KernelChoice *kc = pick_kernel(TR, T->all_affine, T->te_max);
    if (!kc) continue;                                // general cells only have the 128-row instantiations
    T->tep = kc->TEP;
    T->smem_bytes = smem_total(T, smem_need(T, T->all_affine, TR));
    if (T->smem_bytes <= h->smem_optin) done = true;
  }
  cudaFree(vals2); cudaFree(keys2); cudaFree(regular);
  if (!done) {
    const int need = T->smem_bytes;
    cudaFree(adjcell); tiles_free(h);
    return set_err(h, TXASM_EUNSUPPORTED, "row tiles need %d bytes of shared memory", need);
  }

  // 4. opt in to the shared memory size
  const KernelChoice *kc = pick_kernel(T->TR, T->all_affine, T->te_max);
  TX_CUDA(h, cudaFuncSetAttribute(kc->jac, cudaFuncAttributeMaxDynamicSharedMemorySize, T->smem_bytes));
  TX_CUDA(h, cudaFuncSetAttribute(kc->res, cudaFuncAttributeMaxDynamicSharedMemorySize, T->smem_bytes));
  tx_set_carveout(kc->jac); tx_set_carveout(kc->res);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kc->jac, T->TR, smem_total(T, smem_need(T, T->all_affine, T->TR, false, true)));
  T->ctas_per_sm = occ;
  // tile-ordered row tables (CSR begin/length, perm), runs and the out-buffer layout
  auto build_row_tables = [&]() -> int {
    const int64_t slots = (int64_t)T->n_tiles * T->TR;
    if ((rc = dev_alloc(h, &T->d_tile_perm, (size_t)slots * PERM_STRIDE))) return rc;
    k_tile_rowtables<<<(unsigned)((slots + 255) / 256), 256, 0, h->stream>>>(slots, T->d_tile_rows, h->d_rowptr, T->d_perm,
                                                                           T->d_tile_perm);
    TX_CUDA(h, cudaGetLastError());
    TX_CUDA(h, cudaStreamSynchronize(h->stream));
    // runs of rows contiguous in A (TMA bulk stores) and the out-buffer layout
    int *d_nr = nullptr, *d_os = nullptr;
    TX_CUDA(h, cudaMalloc(&d_nr, sizeof(int) * T->n_tiles));
    TX_CUDA(h, cudaMalloc(&d_os, sizeof(int) * T->n_tiles));
    k_tile_runs<false><<<(T->n_tiles + 127) / 128, 128, 0, h->stream>>>(T->n_tiles, T->TR, T->d_tile_rows, h->d_rowptr, T->d_tile_perm,
                                                                        T->d_adjl, T->d_tile_cong,
                                                                        d_nr, d_os, nullptr, nullptr, nullptr);
    std::vector<int> hn(T->n_tiles), ho(T->n_tiles);
    TX_CUDA(h, cudaStreamSynchronize(h->stream));     // h->stream is non-blocking: cudaMemcpy does not wait for it
    TX_CUDA(h, copy_to_device_sync(h, hn.data(), d_nr, sizeof(int) * T->n_tiles));
    TX_CUDA(h, copy_to_device_sync(h, ho.data(), d_os, sizeof(int) * T->n_tiles));
    cudaFree(d_nr); cudaFree(d_os);
    std::vector<int64_t> rp(T->n_tiles + 1, 0);
    int omax = 0;
    for (int i = 0; i < T->n_tiles; ++i) { rp[i + 1] = rp[i] + hn[i]; omax = std::max(omax, ho[i]); }
    T->out_doubles = omax;
    if ((rc = dev_alloc(h, &T->d_run_ptr, (size_t)T->n_tiles + 1))) return rc;
    if ((rc = dev_alloc(h, &T->d_runs, (size_t)rp[T->n_tiles]))) return rc;
    if ((rc = dev_alloc(h, &T->d_tile_rowinfo, (size_t)slots))) return rc;
    TX_CUDA(h, copy_to_device_sync(h, T->d_run_ptr, rp.data(), sizeof(int64_t) * (T->n_tiles + 1)));
    k_tile_runs<true><<<(T->n_tiles + 127) / 128, 128, 0, h->stream>>>(T->n_tiles, T->TR, T->d_tile_rows, h->d_rowptr, T->d_tile_perm,
                                                                       T->d_adjl, T->d_tile_cong,
                                                                       nullptr, nullptr, T->d_run_ptr, T->d_runs, T->d_tile_rowinfo);
    TX_CUDA(h, cudaGetLastError());
    TX_CUDA(h, cudaStreamSynchronize(h->stream));
    // shared memory may have grown with the run padding: re-check
    T->smem_bytes = smem_total(T, smem_need(T, T->all_affine, T->TR));
    if (T->smem_bytes > h->smem_optin) {
      const int need = T->smem_bytes;
      tiles_free(h);
      return set_err(h, TXASM_EUNSUPPORTED, "row tiles need %d bytes of shared memory", need);
    }
    TX_CUDA(h, cudaFuncSetAttribute(kc->jac, cudaFuncAttributeMaxDynamicSharedMemorySize, T->smem_bytes));
    TX_CUDA(h, cudaFuncSetAttribute(kc->res, cudaFuncAttributeMaxDynamicSharedMemorySize, T->smem_bytes));
    return TXASM_OK;
  };
  if ((rc = build_row_tables())) return rc;
  if ((rc = brick_classify(h))) return rc;            // bit 3: the uniform tile is a brick tile (fill_brick.cu)

  // 5. Brick tiles first, then the other tiles whose rows are all uniform: the store-only tiles form the contiguous
  //    ranges [0, n_brick) and [n_brick, n_uni) (one kernel per range, no tile list to chase).  The order inside each
  //    class stays the Morton order.
  {
    std::vector<unsigned char> cls(T->n_tiles);
    TX_CUDA(h, copy_to_device_sync(h, cls.data(), T->d_tile_cong, (size_t)T->n_tiles));
    std::vector<int> src;                              // new tile -> old tile
    for (int i = 0; i < T->n_tiles; ++i) if ((cls[i] & 15) == 15) src.push_back(i);
    T->n_brick = (int)src.size();
    for (int i = 0; i < T->n_tiles; ++i) if ((cls[i] & 15) == 7) src.push_back(i);
    T->n_uni = (int)src.size();
    for (int i = 0; i < T->n_tiles; ++i) if ((cls[i] & 7) != 7 && (cls[i] & 16)) src.push_back(i);      // lattice tiles with face rows
    T->n_edge = (int)src.size();
    for (int i = 0; i < T->n_tiles; ++i) if ((cls[i] & 7) != 7 && !(cls[i] & 16)) src.push_back(i);
    bool moved = false;
    for (int i = 0; i < T->n_tiles; ++i) moved |= (src[i] != i);
    if (moved) {
      int *d_src = nullptr, *rows2 = nullptr;
      const int64_t slots = (int64_t)T->n_tiles * T->TR;
      TX_CUDA(h, cudaMalloc(&d_src, sizeof(int) * T->n_tiles));
      TX_CUDA(h, cudaMalloc(&rows2, sizeof(int) * (size_t)slots));
      TX_CUDA(h, copy_to_device_sync(h, d_src, src.data(), sizeof(int) * T->n_tiles));
      k_permute_tiles<<<(unsigned)((slots + 255) / 256), 256, 0, h->stream>>>(slots, T->TR, d_src, T->d_tile_rows, rows2);
      TX_CUDA(h, cudaMemcpyAsync(T->d_tile_rows, rows2, sizeof(int) * (size_t)slots, cudaMemcpyDeviceToDevice, h->stream));
      TX_CUDA(h, cudaStreamSynchronize(h->stream));
      cudaFree(d_src); cudaFree(rows2);
      free_dev(h, T->d_tile_cell_ptr); free_dev(h, T->d_tile_cells); free_dev(h, T->d_tile_lids); free_dev(h, T->d_adjl);
      free_dev(h, T->d_tile_cong); free_dev(h, T->d_tile_kf); free_dev(h, T->d_tile_perm);
      free_dev(h, T->d_run_ptr); free_dev(h, T->d_runs); free_dev(h, T->d_tile_rowinfo);
      rc = (T->TR == 256) ? build_cells<256>(h, T, adjcell) : build_cells<128>(h, T, adjcell);
      if (rc) return rc;
      if ((rc = build_row_tables())) return rc;
      if ((rc = brick_classify(h))) return rc;
    }
    if (T->n_brick || T->n_edge > T->n_uni) {          // the classification is a function of the tile alone: same tiles, new numbers
      TX_CUDA(h, copy_to_device_sync(h, cls.data(), T->d_tile_cong, (size_t)T->n_tiles));
      for (int i = 0; i < T->n_tiles; ++i)
        if (((cls[i] & 15) == 15) != (i < T->n_brick) || ((cls[i] & 7) == 7) != (i < T->n_uni) ||
            ((cls[i] & 7) != 7 && (cls[i] & 16) != 0) != (i >= T->n_uni && i < T->n_edge)) {
          cudaFree(adjcell); tiles_free(h);
          return set_err(h, TXASM_ESTATE, "tile classes changed under renumbering (tile %d: class %d)", i, (int)cls[i]);
        }
      if ((rc = brick_build(h))) return rc;
    }
  }
  cudaFree(adjcell);
  free_dev(h, T->d_perm);
  // rows of the uniform range (the export may run under k_fill_uniform when it touches none of them)
  if ((rc = dev_alloc(h, &T->d_row_uniform, (size_t)nr))) return rc;
  TX_CUDA(h, cudaMemsetAsync(T->d_row_uniform, 0, (size_t)nr, h->stream));
  if (T->n_uni) k_mark_rows<<<(unsigned)(((int64_t)T->n_uni * T->TR + 255) / 256), 256, 0, h->stream>>>((int64_t)T->n_uni * T->TR, T->d_tile_rows, T->d_row_uniform);
  TX_CUDA(h, cudaGetLastError());
  return TXASM_OK;
}

int tiles_get(txasm_handle h, int tile, int *rows, int *cells, unsigned short *adjl, int *n_cells_out)
{
  const Tiles *T = h->tiles;
  if (!T || tile < 0 || tile >= T->n_tiles) return set_err(h, TXASM_EINVAL, "tile_get: no such tile");
  int64_t ptr[2];
  TX_CUDA(h, copy_to_device_sync(h, ptr, T->d_tile_cell_ptr + tile, sizeof(ptr)));
  if (rows) TX_CUDA(h, copy_to_device_sync(h, rows, T->d_tile_rows + (int64_t)tile * T->TR, sizeof(int) * T->TR));
  if (cells) TX_CUDA(h, copy_to_device_sync(h, cells, T->d_tile_cells + ptr[0], sizeof(int) * (ptr[1] - ptr[0])));
  if (adjl) TX_CUDA(h, copy_to_device_sync(h, adjl, T->d_adjl + (int64_t)tile * T->TR * 8, sizeof(unsigned short) * T->TR * 8));
  if (n_cells_out) *n_cells_out = (int)(ptr[1] - ptr[0]);
  return TXASM_OK;
}

int tiles_count(txasm_handle h) { return h->tiles ? h->tiles->n_tiles : 0; }

int tiles_info(txasm_handle h, txasm_info *info)
{
  const Tiles *T = h->tiles;
  info->n_regular_rows = T->n_regular;
  info->n_tiles = T->n_tiles; info->tile_rows_max = T->TR; info->tile_cells_max = T->te_max;
  info->smem_bytes = T->smem_bytes; info->threads_per_cta = T->TR; info->ctas_per_sm = T->ctas_per_sm;
  info->n_uniform_tiles = T->n_uni; info->n_brick_tiles = T->n_brick; info->n_edge_tiles = T->n_edge - T->n_uni;
  return TXASM_OK;
}

typedef void (*UniKernel)(FillArgs, TileArgs);
struct UniChoice { int TEP; UniKernel k; };
static const UniChoice g_uni_kernels[] = {{416, k_fill_uniform<416>}};
static int uni_smem(int tep) { return tep * 128 + tep * 32 + 16 + (IMG_DOUBLES + 56 + 4) * 8; }

// the uniform range [0, n_uni) goes to the lean kernel: Jacobian type, no mass terms, closed-form sources, aligned A
bool fill_uniform_eligible(txasm_handle h, const FillArgs &a)
{
  const Tiles *T = h->tiles;
  if (!h->opt_uniform || !T || T->n_uni == 0 || !a.jacobian || !a.A || (((uintptr_t)a.A) & 15) != 0 || a.c.has_mass) return false;
  if (!T->all_affine || T->TR != 256 || T->tep != g_uni_kernels[0].TEP) return false;
  for (int i = 0; i < a.c.n_src; ++i)
    if (a.c.src_id[i] != TXASM_SOURCE_SIN3 && a.c.src_id[i] != TXASM_SOURCE_CONSTANT) return false;
  return true;
}

int rows_touch_uniform_tiles(txasm_handle h, const int *d_rows, int64_t n, bool *touch)
{
  *touch = false;
  const Tiles *T = h->tiles;
  if (!T || !T->d_row_uniform || n == 0 || T->n_uni == 0) return TXASM_OK;
  int *d_cnt = nullptr, cnt = 0;
  TX_CUDA(h, cudaMalloc(&d_cnt, sizeof(int)));
  TX_CUDA(h, cudaMemsetAsync(d_cnt, 0, sizeof(int), h->stream));
  k_count_marked<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, d_rows, T->d_row_uniform, d_cnt);
  TX_CUDA(h, cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_cnt);
  *touch = cnt != 0;
  return TXASM_OK;
}

// Which kernel takes which tiles in this evaluate: [0, e_brick) k_fill_brick, [e_brick, e_uni) k_fill_uniform,
// [e_uni, n_tiles) k_fill_rowtile (+ the irregular rows by k_fill_rowgather).
// ... and, when the uniform range is covered, [n_uni, e_edge) by k_fill_edge.
void fill_ranges(txasm_handle h, const FillArgs &a, int *e_brick, int *e_uni, int *e_edge)
{
  const Tiles *T = h->tiles;
  *e_brick = fill_brick_eligible(h, a) ? T->n_brick : 0;
  *e_uni = fill_uniform_eligible(h, a) ? T->n_uni : *e_brick;
  if (e_edge) *e_edge = (*e_uni == T->n_uni && fill_edge_eligible(h, a)) ? T->n_edge : *e_uni;
}

// rows of the tiles [t0, n_tiles) that are not stored from the constant image, and the irregular rows, may carry a
// fused Dirichlet condition; count the Dirichlet rows among them
__global__ void k_count_fusable(int64_t s0, int64_t s1, const int *__restrict__ tile_rows, const unsigned *__restrict__ rowinfo,
                                const int *__restrict__ row_dir, int *__restrict__ count)
{
  const int64_t i = s0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < s1) {
    const int r = tile_rows[i];
    if (r >= 0 && row_dir[r] >= 0 && !(rowinfo[i] & ROW_UNIFORM)) atomicAdd(count, 1);
  }
}
__global__ void k_row_dir(int n, const int *__restrict__ dofs, int *__restrict__ row_dir)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicMax(&row_dir[dofs[i]], i);      // a DOF listed twice: the later entry wins
}
__global__ void k_count_nonneg(int64_t n, const int *__restrict__ v, int *__restrict__ count)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && v[i] >= 0) atomicAdd(count, 1);
}

// Can the Dirichlet rows be written by k_fill_rowtile itself?  Yes when every one of them is a row of a tile outside
// the uniform range and is not stored from the constant row image (rows on the boundary of an inline mesh always
// qualify: they have fewer than 27 entries).  Builds the row -> Dirichlet index table.
int dirichlet_fuse_prepare(txasm_handle h)
{
  Tiles *T = h->tiles;
  h->dir_fusable = 0;
  if (!T || h->n_dir == 0 || h->mode != TXASM_SCATTER_ROWTILE) return TXASM_OK;
  if (h->d_row_dir) { dev_free(h, h->d_row_dir); h->d_row_dir = nullptr; }
  int rc = dev_alloc(h, &h->d_row_dir, (size_t)h->n_rows);
  if (rc) return rc;
  TX_CUDA(h, cudaMemsetAsync(h->d_row_dir, 0xFF, sizeof(int) * (size_t)h->n_rows, h->stream));
  k_row_dir<<<(h->n_dir + 255) / 256, 256, 0, h->stream>>>(h->n_dir, h->d_dir_dofs, h->d_row_dir);
  int *d_cnt = nullptr, cnt[2] = {0, 0};
  TX_CUDA(h, cudaMalloc(&d_cnt, 2 * sizeof(int)));
  TX_CUDA(h, cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int), h->stream));
  k_count_nonneg<<<(unsigned)((h->n_rows + 255) / 256), 256, 0, h->stream>>>(h->n_rows, h->d_row_dir, d_cnt);
  const int64_t s0 = (int64_t)T->n_uni * T->TR, s1 = (int64_t)T->n_tiles * T->TR;
  if (s1 > s0) k_count_fusable<<<(unsigned)((s1 - s0 + 255) / 256), 256, 0, h->stream>>>(s0, s1, T->d_tile_rows, T->d_tile_rowinfo, h->d_row_dir, d_cnt + 1);
  TX_CUDA(h, cudaMemcpyAsync(cnt, d_cnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_cnt);
  h->dir_fusable = (cnt[0] > 0 && cnt[0] == cnt[1]) ? 1 : 0;
  return edge_codes_refresh(h);            // (the entry codes of k_fill_edge flag the Dirichlet rows)
}

int launch_fill_rowtile(txasm_handle h, const FillArgs &a_in, int part, cudaStream_t st, bool fuse_dir)
{
  FillArgs a = a_in;
  Tiles *T = h->tiles;
  const int stage = smem_need(T, T->all_affine, T->TR, a.c.has_mass != 0, a.c.n_src > 0);
  const int smem = smem_total(T, stage);
  const KernelChoice *kc = pick_kernel(T->TR, T->all_affine, T->te_max);
  TileKernel k = a.jacobian ? kc->jac : kc->res;
  const int tma_ok = (a.A && (((uintptr_t)a.A) & 15) == 0) ? 1 : 0;
  int e_brick = 0, e_uni = 0, e_edge = 0;
  fill_ranges(h, a, &e_brick, &e_uni, &e_edge);
  if (part != FILL_REST) {
    h->uniform_used = e_brick > 0 ? 2 : (e_uni > 0 ? 1 : 0);
    if (e_brick > 0) {
      int rc = launch_fill_brick(h, a, st);
      if (rc) return rc;
    }
    if (e_uni > e_brick) {
      const UniChoice &u = g_uni_kernels[0];
      const int us = uni_smem(u.TEP);
      if (!T->uni_attr_set) { TX_CUDA(h, cudaFuncSetAttribute(u.k, cudaFuncAttributeMaxDynamicSharedMemorySize, us)); T->uni_attr_set = true; }
      int occ = 1;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, u.k, 256, us);
      int grid = std::min(e_uni - e_brick, std::max(1, occ) * h->n_sm);
      if (h->opt_grid_cap > 0) grid = std::min(grid, h->opt_grid_cap);
      if (e_brick == 0) T->ctas_per_sm = occ;      // (reported by txasm_info_get: the kernel that covers most tiles)
      TileArgs ta{T->d_tile_rows, T->d_tile_cell_ptr, T->d_tile_cells, T->d_tile_lids, T->d_adjl,
                  T->d_tile_rowinfo, T->d_run_ptr, T->d_runs, T->d_tile_perm, T->lrow, e_uni, stage, tma_ok, e_brick,
                  T->d_tile_cong, T->d_tile_kf, nullptr, nullptr};
      u.k<<<grid, 256, us, st>>>(a, ta);
      TX_CUDA(h, cudaGetLastError());
      h->launches += 1;
    }
  }
  if (part == FILL_UNIFORM) return TXASM_OK;
  if (e_edge > e_uni) {                  // lattice tiles with rows on their faces (Dirichlet rows fused there as well)
    int rc = launch_fill_edge(h, a, st, fuse_dir ? h->d_row_dir : nullptr, fuse_dir ? h->d_dir_vals : nullptr);
    if (rc) return rc;
  }
  const int t_begin = e_edge;
  if (!T->all_affine) {                  // general hexahedra: every cell's element matrix once, then the tiles gather
    int rc = launch_elem_general(h, a, st);
    if (rc) return rc;
  }
  if (t_begin < T->n_tiles) {
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, T->TR, smem);
    if (h->opt_rest_ctas > 0 && t_begin > 0) occ = std::min(occ, h->opt_rest_ctas);
    int grid = std::min(T->n_tiles - t_begin, std::max(1, occ) * h->n_sm);      // persistent CTAs
    if (h->opt_grid_cap > 0) grid = std::min(grid, h->opt_grid_cap);
    if (e_uni == 0) T->ctas_per_sm = occ;
    TileArgs ta{T->d_tile_rows, T->d_tile_cell_ptr, T->d_tile_cells, T->d_tile_lids, T->d_adjl,
                T->d_tile_rowinfo, T->d_run_ptr, T->d_runs, T->d_tile_perm, T->lrow, T->n_tiles, stage, tma_ok, t_begin,
                T->d_tile_cong, T->d_tile_kf, fuse_dir ? h->d_row_dir : nullptr, fuse_dir ? h->d_dir_vals : nullptr};
    k<<<grid, T->TR, smem, st>>>(a, ta);
    TX_CUDA(h, cudaGetLastError());
    h->launches += 1;
  }
  if (T->n_irregular) {
    cudaStream_t keep = h->stream;
    h->stream = st;
    int rc = launch_fill_rowgather_list(h, a, T->d_irregular, T->n_irregular);
    h->stream = keep;
    if (rc) return rc;
  }
  return TXASM_OK;
}

}  // namespace txasm
