// fill_rowtile.cu -- TXASM_SCATTER_ROWTILE: the B200 fast path.  Owner-computes row tiles.
//
// Replaces, for one element block, the whole workset loop of AssemblyEngine::evaluateVolume
// (disc-fe/src/Panzer_AssemblyEngine_impl.hpp:152-181) and the ~9 Kokkos dispatches per 20-cell
// workset behind it with ONE kernel launch:
//
//   setup (once):  rows (local DOFs) are ordered along a Morton curve of their node coordinates and
//     cut into tiles of TR rows; each tile gets the list of cells touching its rows (own + halo
//     cells), a per-row table "which tile cell has me as local vertex a" and, per row, the
//     permutation from the canonical 27-point neighbour index to the CSR slot of that column.
//   evaluate:      one CTA per tile.
//     phase 1 (thread per tile cell): gather LIDs, coordinates, solution; geometry; stage per-cell
//       data in shared memory (constant-Jacobian cells: 6 metric terms + gathered u + source load;
//       general cells: the 36+8 element matrix/vector from the full 2x2x2 rule).
//     phase 2 (thread per row): walk the <=8 cells around the node with the local vertex index as a
//       compile-time constant, accumulate the 27 row entries in REGISTERS (canonical neighbour
//       index is compile time), residual alongside.
//     phase 3: permute into CSR slot order through shared memory and write every row of A once,
//       with plain coalesced stores.
//   => no atomics, no zero-fill pass, no colind reads, each A value and f value written exactly
//      once, bitwise reproducible.  Halo cells are recomputed by neighbouring tiles (~1.5x phase 1).
//
// Rows whose neighbourhood is not a regular 27-point patch (irregular valence, repeated local
// index, > 64 entries) are left to the general row-gather kernel (fill_rowgather.cu).
#include "txasm_internal.hpp"
#include "elem_q1hex.cuh"
#include <cub/cub.cuh>
#include <algorithm>

namespace txasm {

int launch_fill_rowgather_list(txasm_handle h, const FillArgs &a, const int *row_list, int64_t n);

constexpr int PERM_STRIDE = 32;      // bytes per row in the perm table (27 used)
constexpr int LROW_CAP = 64;         // longest row the tile path takes

struct Tiles {
  int TR = 0;                        // rows per tile (= threads per CTA)
  int n_tiles = 0;
  int64_t n_regular = 0, n_irregular = 0;
  int te_max = 0;                    // max cells per tile
  int tep = 0;                       // padded cell stride in shared memory
  int lrow = 27;                     // longest regular row
  bool all_affine = false;
  int *d_tile_rows = nullptr;        // [n_tiles*TR] row ids (Morton order), -1 padding
  int64_t *d_tile_cell_ptr = nullptr;// [n_tiles+1]
  int *d_tile_cells = nullptr;       // cell ids per tile, ascending
  unsigned short *d_adjl = nullptr;  // [n_tiles][8][TR] tile-local cell index of the cell having row r as vertex a
  unsigned char *d_perm = nullptr;   // [n_rows][32] canonical neighbour -> CSR slot (0xFF absent)
  int *d_irregular = nullptr;        // list of irregular rows
  int smem_bytes = 0;
  int ctas_per_sm = 0;
};

// canonical 27-point neighbour index of vertex b seen from vertex a of the same cell
__host__ __device__ constexpr int canon(int a, int b)
{
  return ((hex_sx(b) - hex_sx(a)) / 2 + 1) + 3 * ((hex_sy(b) - hex_sy(a)) / 2 + 1) + 9 * ((hex_sz(b) - hex_sz(a)) / 2 + 1);
}

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
  for (int c = 0; c < PERM_STRIDE; ++c) {
    unsigned char p = 0xFF;
    if (ok && c < 27 && colc[c] >= 0) {
      int lo = 0, hi = len - 1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int v = colind[b0 + mid];
        if (v == colc[c]) { p = (unsigned char)mid; break; }
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
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int d = 0; d < 3; ++d) {
    const unsigned long long k = dbl_key(xyz[i * 3 + d]);
    atomicMin(&mn[d], k);
    atomicMax(&mx[d], k);
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
__global__ void k_morton(int64_t n, const double *__restrict__ xyz, const unsigned long long *__restrict__ mn,
                         const unsigned long long *__restrict__ mx, const unsigned char *__restrict__ regular,
                         unsigned long long *__restrict__ keys, int *__restrict__ vals)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  vals[i] = (int)i;
  if (!regular[i]) { keys[i] = ~0ull; return; }
  // one common quantum for the three axes keeps the curve cells cubic
  double ext = 0.0;
  for (int d = 0; d < 3; ++d) ext = fmax(ext, key_dbl(mx[d]) - key_dbl(mn[d]));
  const double scale = (ext > 0.0) ? 1048576.0 / ext : 0.0;   // 2^20 quanta over the longest axis
  unsigned long long q[3];
  for (int d = 0; d < 3; ++d) {
    const double t = (xyz[i * 3 + d] - key_dbl(mn[d])) * scale;
    q[d] = (unsigned long long)fmin(fmax(t + 0.5, 0.0), 2097151.0);
  }
  keys[i] = spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2);
}

__global__ void k_tile_rows(int64_t n_slots, int64_t n_regular, const int *__restrict__ sorted, int *__restrict__ tile_rows)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_slots) tile_rows[i] = (i < n_regular) ? sorted[i] : -1;
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
  if (tid == 0) ncells[t] = total;
  for (int i = tid; i < total; i += TR) cells_tmp[(int64_t)t * CAP + i] = u[i];
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    unsigned short v = 0xFFFF;
    if (mine[a] >= 0) {
      int lo = 0, hi = total - 1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (u[mid] == mine[a]) { v = (unsigned short)mid; break; }
        if (u[mid] < mine[a]) lo = mid + 1; else hi = mid - 1;
      }
    }
    adjl[((int64_t)t * 8 + a) * TR + tid] = v;
  }
}

__global__ void k_compact_cells(int n_tiles, int cap, const int *__restrict__ ncells, const int64_t *__restrict__ ptr,
                                const int *__restrict__ tmp, int *__restrict__ out)
{
  const int t = blockIdx.x;
  if (t >= n_tiles) return;
  const int n = ncells[t];
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[ptr[t] + i] = tmp[(int64_t)t * cap + i];
}

__global__ void k_tile_affine(int64_t n, const int *__restrict__ cells, const unsigned char *__restrict__ aff, int *__restrict__ n_non)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && !aff[cells[i]]) atomicAdd(n_non, 1);
}

__global__ void k_list_irregular(int64_t n_rows, const unsigned char *__restrict__ regular, int *__restrict__ list, int *__restrict__ count)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_rows && !regular[i]) list[atomicAdd(count, 1)] = (int)i;
}

// ============================================================================ the fill kernel
struct TileArgs {
  const int *tile_rows;
  const int64_t *tile_cell_ptr;
  const int *tile_cells;
  const unsigned short *adjl;
  const unsigned char *perm;
  int tep;     // cell stride in shared memory
  int lrow;    // out-buffer row stride
};

// per-cell staging size in doubles
__host__ __device__ constexpr int stage_doubles(bool affine, bool mass, bool src)
{
  return affine ? (6 + 8 + (mass ? 9 : 0) + (src ? 8 : 0)) : (36 + 8);
}

template <int A, bool JAC>
__device__ __forceinline__ void row_accum_affine(const double *__restrict__ sm, int tep, int el, const FillCoef &c,
                                                 bool has_mass, bool has_src, double (&acc)[27], double &fr)
{
  // staging layout (k-major, stride tep): G[6] | ug[8] | (det, um[8]) | src[8]
  double G[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) G[k] = sm[k * tep + el];
  const double *su = sm + 6 * tep;
  {
    double t;
#define TX_KAB(B)                                                               \
    t = aff_kab<A, B>(G);                                                       \
    if (JAC) acc[canon(A, B)] = fma(c.cK, t, acc[canon(A, B)]);                 \
    fr = fma(t, su[(B) * tep + el], fr);
    TX_KAB(0) TX_KAB(1) TX_KAB(2) TX_KAB(3) TX_KAB(4) TX_KAB(5) TX_KAB(6) TX_KAB(7)
#undef TX_KAB
  }
  int base = 14;
  if (has_mass) {
    const double det = sm[base * tep + el];
    const double *sv = sm + (base + 1) * tep;
#define TX_MAB(B)                                                               \
    { const double m = det * aff_mass(A, B);                                    \
      if (JAC) acc[canon(A, B)] = fma(c.cM, m, acc[canon(A, B)]);               \
      fr = fma(m, sv[(B) * tep + el], fr); }
    TX_MAB(0) TX_MAB(1) TX_MAB(2) TX_MAB(3) TX_MAB(4) TX_MAB(5) TX_MAB(6) TX_MAB(7)
#undef TX_MAB
    base += 9;
  }
  if (has_src) fr += sm[(base + A) * tep + el];
}

template <int A, bool JAC>
__device__ __forceinline__ void row_accum_general(const double *__restrict__ sm, int tep, int el, double (&acc)[27], double &fr)
{
  // staging layout: K sym[36] | r[8]
  if (JAC) {
#define TX_GAB(B) acc[canon(A, B)] += sm[sym_idx(A, B) * tep + el];
    TX_GAB(0) TX_GAB(1) TX_GAB(2) TX_GAB(3) TX_GAB(4) TX_GAB(5) TX_GAB(6) TX_GAB(7)
#undef TX_GAB
  }
  fr += sm[(36 + A) * tep + el];
}

template <int TR, bool AFFINE, bool JAC>
__global__ void __launch_bounds__(TR, AFFINE ? 512 / TR : 1) k_fill_rowtile(FillArgs A, TileArgs T)
{
  extern __shared__ double sm[];
  const int t = blockIdx.x, tid = threadIdx.x;
  const int tep = T.tep;
  const int64_t cb = T.tile_cell_ptr[t];
  const int ncell = (int)(T.tile_cell_ptr[t + 1] - cb);
  const bool has_mass = A.c.has_mass != 0, has_src = A.c.n_src > 0;

  // ---------------- phase 1: one thread per tile cell
  for (int j = tid; j < ncell; j += TR) {
    const int64_t e = T.tile_cells[cb + j];
    int lid[8];
    {
      const int4 *p = reinterpret_cast<const int4 *>(A.lids + e * 8);
      const int4 v0 = __ldg(p), v1 = __ldg(p + 1);
      lid[0] = v0.x; lid[1] = v0.y; lid[2] = v0.z; lid[3] = v0.w;
      lid[4] = v1.x; lid[5] = v1.y; lid[6] = v1.z; lid[7] = v1.w;
    }
    double X[8][3], ug[8], um[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int64_t l = lid[n];
      X[n][0] = __ldg(A.xyz + l * 3); X[n][1] = __ldg(A.xyz + l * 3 + 1); X[n][2] = __ldg(A.xyz + l * 3 + 2);
      double g = 0.0, m = 0.0;
#pragma unroll
      for (int v = 0; v < 3; ++v)
        if (A.c.has_vec[v]) {
          const double xv = __ldg(A.x[v] + l);
          g = fma(A.c.kg[v], xv, g);
          m = fma(A.c.km[v], xv, m);
        }
      ug[n] = g; um[n] = m;
    }
    if (AFFINE) {
      double J[3][3];
      AffineGeom g;
      affine_geom(X, J, g);
#pragma unroll
      for (int k = 0; k < 6; ++k) sm[k * tep + j] = g.G[k];
#pragma unroll
      for (int b = 0; b < 8; ++b) sm[(6 + b) * tep + j] = ug[b];
      int base = 14;
      if (has_mass) {
        sm[base * tep + j] = g.det;
#pragma unroll
        for (int b = 0; b < 8; ++b) sm[(base + 1 + b) * tep + j] = um[b];
        base += 9;
      }
      if (has_src) {
        // source load vector: det * sum_q N_a(xi_q) sum_s mult_s s_s(x_q),  x_q = Xc + J xi_q
        double xc[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          double s = 0.0;
#pragma unroll
          for (int n = 0; n < 8; ++n) s += X[n][d];
          xc[d] = 0.125 * s;
        }
        double bl[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) bl[a] = 0.0;
        // Axis-aligned cell (diagonal Jacobian): x_q, y_q, z_q take two values each, so a separable
        // closure model needs 6 instead of 24 function evaluations.  Exactly the same 8 point values.
        const bool diag = (J[0][1] == 0.0) & (J[0][2] == 0.0) & (J[1][0] == 0.0) & (J[1][2] == 0.0) &
                          (J[2][0] == 0.0) & (J[2][1] == 0.0);
        double sq8[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) sq8[q] = 0.0;
        for (int s = 0; s < A.c.n_src; ++s) {
          const int id = A.c.src_id[s];
          const double mult = A.c.src_mult[s];
          if (id == TXASM_SOURCE_IP_ARRAY) {
#pragma unroll
            for (int q = 0; q < 8; ++q) sq8[q] = fma(mult, A.c.src_ip[s][e * 8 + q], sq8[q]);
          } else if (id == TXASM_SOURCE_CONSTANT) {
#pragma unroll
            for (int q = 0; q < 8; ++q) sq8[q] += mult;
          } else if (diag && id == TXASM_SOURCE_SIN3) {
            double fx[2], fy[2], fz[2];
            fx[0] = sinpi(2.0 * (xc[0] - J[0][0] * TX_INV_SQRT3)); fx[1] = sinpi(2.0 * (xc[0] + J[0][0] * TX_INV_SQRT3));
            fy[0] = sinpi(2.0 * (xc[1] - J[1][1] * TX_INV_SQRT3)); fy[1] = sinpi(2.0 * (xc[1] + J[1][1] * TX_INV_SQRT3));
            fz[0] = sinpi(2.0 * (xc[2] - J[2][2] * TX_INV_SQRT3)); fz[1] = sinpi(2.0 * (xc[2] + J[2][2] * TX_INV_SQRT3));
            const double cm = mult * 118.43525281307230;
#pragma unroll
            for (int q = 0; q < 8; ++q) sq8[q] = fma(cm * fx[q & 1], fy[(q >> 1) & 1] * fz[(q >> 2) & 1], sq8[q]);
          } else {
#pragma unroll 1
            for (int q = 0; q < 8; ++q) {
              const double xi = (q & 1) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
              const double et = (q & 2) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
              const double ze = (q & 4) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
              const double xq = xc[0] + J[0][0] * xi + J[0][1] * et + J[0][2] * ze;
              const double yq = xc[1] + J[1][0] * xi + J[1][1] * et + J[1][2] * ze;
              const double zq = xc[2] + J[2][0] * xi + J[2][1] * et + J[2][2] * ze;
              sq8[q] = fma(mult, source_eval(id, xq, yq, zq), sq8[q]);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const double sq = sq8[q] * g.det;
#pragma unroll
          for (int a = 0; a < 8; ++a) {
            // N_a(xi_q) = (1 +- 1/sqrt3)^k (1 -+ 1/sqrt3)^(3-k) / 8: compile-time per (a,q)
            const double na = 0.125 * (1.0 + hex_sx(a) * ((q & 1) ? TX_INV_SQRT3 : -TX_INV_SQRT3)) *
                              (1.0 + hex_sy(a) * ((q & 2) ? TX_INV_SQRT3 : -TX_INV_SQRT3)) *
                              (1.0 + hex_sz(a) * ((q & 4) ? TX_INV_SQRT3 : -TX_INV_SQRT3));
            bl[a] = fma(na, sq, bl[a]);
          }
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) sm[(base + a) * tep + j] = bl[a];
      }
    } else {
      double K[36], r[8];
      elem_general<JAC>(X, ug, um, A.c, e, K, r);
      if (JAC) {
#pragma unroll
        for (int k = 0; k < 36; ++k) sm[k * tep + j] = K[k];
      }
#pragma unroll
      for (int a = 0; a < 8; ++a) sm[(36 + a) * tep + j] = r[a];
    }
  }
  __syncthreads();

  // ---------------- phase 2: one thread per row, 27 entries in registers
  const int row = T.tile_rows[(int64_t)t * TR + tid];
  double acc[27];
#pragma unroll
  for (int c = 0; c < 27; ++c) acc[c] = 0.0;
  double fr = 0.0;
  if (row >= 0) {
    const unsigned short *al = T.adjl + (int64_t)t * 8 * TR + tid;
#define TX_ROW(AA)                                                                                   \
    { const int el = al[(AA) * TR];                                                                  \
      if (el != 0xFFFF) {                                                                            \
        if (AFFINE) row_accum_affine<AA, JAC>(sm, tep, el, A.c, has_mass, has_src, acc, fr);         \
        else row_accum_general<AA, JAC>(sm, tep, el, acc, fr);                                       \
      } }
    TX_ROW(0) TX_ROW(1) TX_ROW(2) TX_ROW(3) TX_ROW(4) TX_ROW(5) TX_ROW(6) TX_ROW(7)
#undef TX_ROW
    if (A.f) A.f[row] = fr;
  }
  if (!JAC) return;

  // ---------------- phase 3: permute to CSR slot order in shared memory, coalesced row stores
  __syncthreads();                       // staging is dead; reuse it as out[TR][lrow]
  const int lrow = T.lrow;
  double *out = sm;
  int64_t my_beg = 0;
  int my_len = 0;
  if (row >= 0) {
    my_beg = A.rowptr[row];
    my_len = (int)(A.rowptr[row + 1] - my_beg);
    double *o = out + tid * lrow;
    for (int s = 0; s < lrow; ++s) o[s] = 0.0;
    const uint4 *pp = reinterpret_cast<const uint4 *>(T.perm + (int64_t)row * PERM_STRIDE);
    const uint4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
    const unsigned w[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
    for (int c = 0; c < 27; ++c) {
      const unsigned p = (w[c >> 2] >> (8 * (c & 3))) & 0xFFu;
      if (p != 0xFFu) o[p] = acc[c];
    }
  }
  __syncwarp();
  // each warp stores the 32 rows its own lanes just staged: one coalesced store per row
  const int lane = tid & 31;
  const double *wout = out + (tid - lane) * lrow;
#pragma unroll 8
  for (int i = 0; i < 32; ++i) {
    const int64_t b = __shfl_sync(0xffffffffu, my_beg, i);
    const int len = __shfl_sync(0xffffffffu, my_len, i);
    double *dst = A.A + b;
    if (lane < len) dst[lane] = wout[i * lrow + lane];
    for (int s = lane + 32; s < len; s += 32) dst[s] = wout[i * lrow + s];
  }
}

// ============================================================================ host side
template <class T>
static void free_dev(txasm_handle h, T *&p) { if (p) { dev_free(h, (void *)p); p = nullptr; } }

void tiles_free(txasm_handle h)
{
  if (!h->tiles) return;
  Tiles *T = h->tiles;
  free_dev(h, T->d_tile_rows); free_dev(h, T->d_tile_cell_ptr); free_dev(h, T->d_tile_cells);
  free_dev(h, T->d_adjl); free_dev(h, T->d_perm); free_dev(h, T->d_irregular);
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
    TX_CUDA(h, cudaMemcpy(hc.data(), ncells, sizeof(int) * (T->n_tiles + 1), cudaMemcpyDeviceToHost));
    std::vector<int64_t> hp(T->n_tiles + 1);
    int64_t s = 0;
    for (int i = 0; i < T->n_tiles; ++i) { hp[i] = s; s += hc[i]; }
    hp[T->n_tiles] = s;
    rc = dev_alloc(h, &T->d_tile_cell_ptr, (size_t)T->n_tiles + 1);
    if (rc) return rc;
    TX_CUDA(h, cudaMemcpy(T->d_tile_cell_ptr, hp.data(), sizeof(int64_t) * (T->n_tiles + 1), cudaMemcpyHostToDevice));
    rc = dev_alloc(h, &T->d_tile_cells, (size_t)s);
    if (rc) return rc;
    k_compact_cells<<<T->n_tiles, 128, 0, h->stream>>>(T->n_tiles, cap, ncells, T->d_tile_cell_ptr, tmp, T->d_tile_cells);
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

static int smem_need(const Tiles *T, bool affine, int TR, bool mass = true, bool src = true)
{
  // setup sizes for the worst case (mass + source on); a launch asks for what its term list needs
  const int per_cell = stage_doubles(affine, mass, src);
  const int stage = per_cell * T->tep * 8;
  const int out = TR * T->lrow * 8;
  return std::max(stage, out);
}

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
    TX_CUDA(h, cudaMemcpy(hi.data(), irr_tmp, sizeof(int) * n_irr, cudaMemcpyDeviceToHost));
    std::sort(hi.begin(), hi.end());
    TX_CUDA(h, cudaMemcpy(T->d_irregular, hi.data(), sizeof(int) * n_irr, cudaMemcpyHostToDevice));
  }
  cudaFree(irr_tmp);
  if (T->n_regular == 0) { cudaFree(regular); cudaFree(adjcell); tiles_free(h); return set_err(h, TXASM_EUNSUPPORTED, "no regular rows"); }

  // 2. Morton order of the regular rows
  unsigned long long *bb = nullptr, *keys = nullptr, *keys2 = nullptr;
  int *vals = nullptr, *vals2 = nullptr;
  TX_CUDA(h, cudaMalloc(&bb, sizeof(unsigned long long) * 6));
  {
    unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0, 0, 0};
    TX_CUDA(h, cudaMemcpy(bb, init, sizeof(init), cudaMemcpyHostToDevice));
  }
  TX_CUDA(h, cudaMalloc(&keys, sizeof(unsigned long long) * (size_t)nr));
  TX_CUDA(h, cudaMalloc(&keys2, sizeof(unsigned long long) * (size_t)nr));
  TX_CUDA(h, cudaMalloc(&vals, sizeof(int) * (size_t)nr));
  TX_CUDA(h, cudaMalloc(&vals2, sizeof(int) * (size_t)nr));
  k_bbox<<<(unsigned)((nr + 255) / 256), 256, 0, h->stream>>>(nr, h->d_xyz, bb, bb + 3);
  k_morton<<<(unsigned)((nr + 255) / 256), 256, 0, h->stream>>>(nr, h->d_xyz, bb, bb + 3, regular, keys, vals);
  {
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, vals, vals2, (int)nr, 0, 64, h->stream);
    void *tmp = nullptr; TX_CUDA(h, cudaMalloc(&tmp, tb ? tb : 1));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, vals, vals2, (int)nr, 0, 64, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    TX_CUDA(h, e);
  }
  cudaFree(bb); cudaFree(keys); cudaFree(keys2); cudaFree(vals);

  // 3. tiles: TR rows each; try the large tile first, shrink if shared memory does not fit
  const int try_tr[2] = {256, 128};
  bool done = false;
  for (int attempt = 0; attempt < 2 && !done; ++attempt) {
    const int TR = try_tr[attempt];
    free_dev(h, T->d_tile_rows); free_dev(h, T->d_tile_cell_ptr); free_dev(h, T->d_tile_cells); free_dev(h, T->d_adjl);
    T->TR = TR;
    T->n_tiles = (int)((T->n_regular + TR - 1) / TR);
    const int64_t slots = (int64_t)T->n_tiles * TR;
    if ((rc = dev_alloc(h, &T->d_tile_rows, (size_t)slots))) return rc;
    k_tile_rows<<<(unsigned)((slots + 255) / 256), 256, 0, h->stream>>>(slots, T->n_regular, vals2, T->d_tile_rows);
    TX_CUDA(h, cudaGetLastError());
    rc = (TR == 256) ? build_cells<256>(h, T, adjcell) : build_cells<128>(h, T, adjcell);
    if (rc) return rc;
    T->tep = T->te_max | 1;                           // odd stride
    T->smem_bytes = smem_need(T, T->all_affine, TR);
    // general cells need the big staging: prefer the small tile for them
    if (!T->all_affine && TR == 256) continue;
    if (T->smem_bytes <= h->smem_optin) done = true;
  }
  cudaFree(vals2); cudaFree(regular); cudaFree(adjcell);
  if (!done) { tiles_free(h); return set_err(h, TXASM_EUNSUPPORTED, "row tiles need %d bytes of shared memory", T->smem_bytes); }

  // 4. opt in to the shared memory size for every instantiation
#define TX_ATTR(K)                                                                                                  \
  TX_CUDA(h, cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, T->smem_bytes))
  if (T->TR == 256) {
    TX_ATTR((k_fill_rowtile<256, true, true>)); TX_ATTR((k_fill_rowtile<256, true, false>));
    TX_ATTR((k_fill_rowtile<256, false, true>)); TX_ATTR((k_fill_rowtile<256, false, false>));
  } else {
    TX_ATTR((k_fill_rowtile<128, true, true>)); TX_ATTR((k_fill_rowtile<128, true, false>));
    TX_ATTR((k_fill_rowtile<128, false, true>)); TX_ATTR((k_fill_rowtile<128, false, false>));
  }
#undef TX_ATTR
  int occ = 0;
  if (T->TR == 256) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fill_rowtile<256, true, true>, 256, T->smem_bytes);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fill_rowtile<128, false, true>, 128, T->smem_bytes);
  T->ctas_per_sm = occ;
  return TXASM_OK;
}

int tiles_info(txasm_handle h, txasm_info *info)
{
  const Tiles *T = h->tiles;
  info->n_regular_rows = T->n_regular;
  info->n_tiles = T->n_tiles; info->tile_rows_max = T->TR; info->tile_cells_max = T->te_max;
  info->smem_bytes = T->smem_bytes; info->threads_per_cta = T->TR; info->ctas_per_sm = T->ctas_per_sm;
  return TXASM_OK;
}

int launch_fill_rowtile(txasm_handle h, const FillArgs &a)
{
  Tiles *T = h->tiles;
  TileArgs ta{T->d_tile_rows, T->d_tile_cell_ptr, T->d_tile_cells, T->d_adjl, T->d_perm, T->tep, T->lrow};
  const int smem = smem_need(T, T->all_affine, T->TR, a.c.has_mass != 0, a.c.n_src > 0);
#define TX_LAUNCH(TRv, AFF, JACv) k_fill_rowtile<TRv, AFF, JACv><<<T->n_tiles, TRv, smem, h->stream>>>(a, ta)
  if (T->TR == 256) {
    if (T->all_affine) { if (a.jacobian) TX_LAUNCH(256, true, true); else TX_LAUNCH(256, true, false); }
    else { if (a.jacobian) TX_LAUNCH(256, false, true); else TX_LAUNCH(256, false, false); }
  } else {
    if (T->all_affine) { if (a.jacobian) TX_LAUNCH(128, true, true); else TX_LAUNCH(128, true, false); }
    else { if (a.jacobian) TX_LAUNCH(128, false, true); else TX_LAUNCH(128, false, false); }
  }
#undef TX_LAUNCH
  TX_CUDA(h, cudaGetLastError());
  h->launches += 1;
  if (T->n_irregular) {
    int rc = launch_fill_rowgather_list(h, a, T->d_irregular, T->n_irregular);
    if (rc) return rc;
  }
  return TXASM_OK;
}

}  // namespace txasm
