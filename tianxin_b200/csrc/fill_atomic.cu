// fill_atomic.cu -- TXASM_SCATTER_ATOMIC: element-parallel fill with searched atomic scatter.
//
// One thread per cell runs the whole Phalanx DAG of the block in registers (gather -> geometry ->
// basis -> gradient -> integrate), then scatters exactly like
// ScatterResidual_Tpetra<Jacobian>::evaluateFields
// (disc-fe/src/evaluators/Panzer_ScatterResidual_Tpetra_impl.hpp:374-413): atomic add of the residual,
// and per row KokkosSparse sumIntoValues(lid, lids, N, vals, is_sorted=true, force_atomic=true), i.e. a
// binary search of the sorted CSR row per column and red.global.add.f64; absent columns are skipped.
// Needs f and A zeroed first.  Kept as the literal restatement of the reference scatter and as the
// cross-check of the atomics-free paths; the row-tile path is the fast one.
#include "txasm_internal.hpp"
#include "elem_q1hex.cuh"

namespace txasm {

template <bool JAC>
__global__ void __launch_bounds__(128) k_fill_atomic(FillArgs A)
{
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= A.n_cells) return;
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
  double K[36], r[8];
  elem_general<JAC>(X, ug, um, A.c, e, K, r);
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int row = lid[a];
    if (A.f) atomicAdd(A.f + row, r[a]);
    if (JAC) {
      const int64_t b0 = A.rowptr[row];
      const int len = (int)(A.rowptr[row + 1] - b0);
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int col = lid[b];
        int lo = 0, hi = len - 1, at = -1;
        while (lo <= hi) {
          const int mid = (lo + hi) >> 1;
          const int c = __ldg(A.colind + b0 + mid);
          if (c == col) { at = mid; break; }
          if (c < col) lo = mid + 1; else hi = mid - 1;
        }
        if (at >= 0) atomicAdd(A.A + b0 + at, K[sym_idx(a, b)]);
      }
    }
  }
}

int launch_fill_atomic(txasm_handle h, const FillArgs &a)
{
  const unsigned grid = (unsigned)((a.n_cells + 127) / 128);
  if (a.jacobian) k_fill_atomic<true><<<grid, 128, 0, h->stream>>>(a);
  else k_fill_atomic<false><<<grid, 128, 0, h->stream>>>(a);
  TX_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return TXASM_OK;
}

}  // namespace txasm
