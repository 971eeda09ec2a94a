// setup_kernels.cu -- one-time device-side setup: node coordinates by LID, node->element
// adjacency, ghosted CSR graph, affine classification.
//
// These replace host work the reference does at setup time and caches:
//   * workset coordinate copies        disc-fe/src/Panzer_Workset_Builder_impl.hpp:153-187
//   * TpetraLinearObjFactory::buildGhostedGraph (insertGlobalIndices + fillComplete)
//                                      disc-fe/src/lof/Panzer_TpetraLinearObjFactory_impl.hpp:558-650
#include "txasm_internal.hpp"
#include "elem_q1hex.cuh"
#include <cub/cub.cuh>

namespace txasm {

// ------------------------------------------------------------------ node coordinates by LID
__global__ void k_scatter_coords(int64_t n_cells, const int *__restrict__ lids, const double *__restrict__ cc,
                                 double *__restrict__ xyz)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // (cell, vertex)
  if (i >= n_cells * 8) return;
  const int lid = lids[i];
  // every cell holding the vertex writes the same three values: benign
  xyz[(int64_t)lid * 3 + 0] = cc[i * 3 + 0];
  xyz[(int64_t)lid * 3 + 1] = cc[i * 3 + 1];
  xyz[(int64_t)lid * 3 + 2] = cc[i * 3 + 2];
}

int build_node_coords(txasm_handle h, const double *d_cc)
{
  const int64_t n = h->n_cells * 8;
  k_scatter_coords<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->n_cells, h->d_lids, d_cc, h->d_xyz);
  TX_CUDA(h, cudaGetLastError());
  return TXASM_OK;
}

// ------------------------------------------------------------------ adjacency (transpose of the LID table)
__global__ void k_adj_count(int64_t n, const int *__restrict__ lids, int *__restrict__ cnt)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&cnt[lids[i]], 1);
}
__global__ void k_adj_fill(int64_t n, const int *__restrict__ lids, const int64_t *__restrict__ ptr,
                           int *__restrict__ cursor, int *__restrict__ adj)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int row = lids[i];
  const int p = atomicAdd(&cursor[row], 1);
  adj[ptr[row] + p] = (int)i;   // packed cell*8 + local vertex
}
__global__ void k_adj_sort(int64_t n_rows, const int64_t *__restrict__ ptr, int *__restrict__ adj)
{
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int64_t b = ptr[r];
  const int n = (int)(ptr[r + 1] - b);
  for (int i = 1; i < n; ++i) {   // insertion sort, n ~ 8
    const int v = adj[b + i];
    int j = i - 1;
    while (j >= 0 && adj[b + j] > v) { adj[b + j + 1] = adj[b + j]; --j; }
    adj[b + j + 1] = v;
  }
}
__global__ void k_widen(int64_t n, const int *__restrict__ in, int64_t *__restrict__ out)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

static int exclusive_scan_i64(txasm_handle h, const int64_t *in, int64_t *out, int64_t n)
{
  size_t tmp_bytes = 0;
  TX_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, n, h->stream));
  void *tmp = nullptr;
  TX_CUDA(h, cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
  cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, n, h->stream);
  cudaStreamSynchronize(h->stream);
  cudaFree(tmp);
  TX_CUDA(h, e);
  return TXASM_OK;
}

int build_adjacency(txasm_handle h)
{
  if (h->d_adj_ptr) return TXASM_OK;
  const int64_t n = h->n_cells * 8, nr = h->n_rows;
  if (h->n_cells >= (int64_t(1) << 28)) return set_err(h, TXASM_EUNSUPPORTED, "more than 2^28 cells per rank");
  int *cnt = nullptr;
  int64_t *cnt64 = nullptr;
  TX_CUDA(h, cudaMalloc(&cnt, sizeof(int) * (nr + 1)));
  TX_CUDA(h, cudaMalloc(&cnt64, sizeof(int64_t) * (nr + 1)));
  TX_CUDA(h, cudaMemsetAsync(cnt, 0, sizeof(int) * (nr + 1), h->stream));
  k_adj_count<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, h->d_lids, cnt);
  k_widen<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, h->stream>>>(nr + 1, cnt, cnt64);
  int rc = dev_alloc(h, &h->d_adj_ptr, nr + 1);
  if (rc) return rc;
  rc = exclusive_scan_i64(h, cnt64, h->d_adj_ptr, nr + 1);
  if (rc) return rc;
  // max adjacency
  {
    size_t tb = 0; int *dmax = nullptr;
    TX_CUDA(h, cudaMalloc(&dmax, sizeof(int)));
    cub::DeviceReduce::Max(nullptr, tb, cnt, dmax, (int)nr, h->stream);
    void *tmp = nullptr; TX_CUDA(h, cudaMalloc(&tmp, tb ? tb : 1));
    cub::DeviceReduce::Max(tmp, tb, cnt, dmax, (int)nr, h->stream);
    TX_CUDA(h, cudaMemcpyAsync(&h->max_adj, dmax, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    TX_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(tmp); cudaFree(dmax);
  }
  rc = dev_alloc(h, &h->d_adj, (size_t)n);
  if (rc) return rc;
  TX_CUDA(h, cudaMemsetAsync(cnt, 0, sizeof(int) * (nr + 1), h->stream));
  k_adj_fill<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, h->d_lids, h->d_adj_ptr, cnt, h->d_adj);
  k_adj_sort<<<(unsigned)((nr + 255) / 256), 256, 0, h->stream>>>(nr, h->d_adj_ptr, h->d_adj);
  TX_CUDA(h, cudaGetLastError());
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(cnt); cudaFree(cnt64);
  if (h->max_adj > MAX_ADJ) return set_err(h, TXASM_EUNSUPPORTED, "a node has %d adjacent cells (max %d)", h->max_adj, MAX_ADJ);
  return TXASM_OK;
}

// ------------------------------------------------------------------ ghosted graph on the device
// Row r gets every LID of every cell containing r; sorted by local column index, duplicates merged
// (what CrsGraph::fillComplete leaves).  Pass 1 counts, pass 2 writes.
template <bool WRITE>
__global__ void k_graph_rows(int64_t n_rows, const int64_t *__restrict__ adj_ptr, const int *__restrict__ adj,
                             const int *__restrict__ lids, int64_t *__restrict__ cnt_or_ptr, int *__restrict__ colind)
{
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int cols[MAX_ADJ * 8];
  int n = 0;
  for (int64_t k = adj_ptr[r]; k < adj_ptr[r + 1]; ++k) {
    const int64_t cell = adj[k] >> 3;
    for (int b = 0; b < 8; ++b) {
      const int c = lids[cell * 8 + b];
      // sorted insert with dedup
      int j = n - 1;
      bool dup = false;
      while (j >= 0 && cols[j] >= c) { if (cols[j] == c) { dup = true; break; } --j; }
      if (dup) continue;
      for (int m = n - 1; m > j; --m) cols[m + 1] = cols[m];
      cols[j + 1] = c;
      ++n;
    }
  }
  if (!WRITE) cnt_or_ptr[r] = n;
  else {
    const int64_t b = cnt_or_ptr[r];
    for (int i = 0; i < n; ++i) colind[b + i] = cols[i];
  }
}

int build_graph_device(txasm_handle h, int64_t *nnz_out)
{
  int rc = build_adjacency(h);
  if (rc) return rc;
  const int64_t nr = h->n_rows;
  int64_t *cnt = nullptr, *rowptr = nullptr;
  TX_CUDA(h, cudaMalloc(&cnt, sizeof(int64_t) * (nr + 1)));
  TX_CUDA(h, cudaMemsetAsync(cnt, 0, sizeof(int64_t) * (nr + 1), h->stream));
  k_graph_rows<false><<<(unsigned)((nr + 127) / 128), 128, 0, h->stream>>>(nr, h->d_adj_ptr, h->d_adj, h->d_lids, cnt, nullptr);
  rc = dev_alloc(h, &rowptr, nr + 1);
  if (rc) return rc;
  rc = exclusive_scan_i64(h, cnt, rowptr, nr + 1);
  cudaFree(cnt);
  if (rc) return rc;
  int64_t nnz = 0;
  TX_CUDA(h, copy_to_device_sync(h, &nnz, rowptr + nr, sizeof(int64_t)));
  int *colind = nullptr;
  rc = dev_alloc(h, &colind, (size_t)nnz);
  if (rc) return rc;
  k_graph_rows<true><<<(unsigned)((nr + 127) / 128), 128, 0, h->stream>>>(nr, h->d_adj_ptr, h->d_adj, h->d_lids, rowptr, colind);
  TX_CUDA(h, cudaGetLastError());
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  h->d_rowptr = rowptr; h->d_colind = colind; h->nnz = nnz; h->have_graph = true;
  if (nnz_out) *nnz_out = nnz;
  return TXASM_OK;
}

// ------------------------------------------------------------------ affine classification
__global__ void k_classify(int64_t n_cells, const int *__restrict__ lids, const double *__restrict__ xyz,
                           double tol, unsigned char *__restrict__ flag, unsigned long long *__restrict__ n_aff)
{
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int is = 0;
  if (e < n_cells) {
    double X[8][3];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int64_t l = lids[e * 8 + n];
      X[n][0] = xyz[l * 3]; X[n][1] = xyz[l * 3 + 1]; X[n][2] = xyz[l * 3 + 2];
    }
    is = (tol >= 0.0) && (hex_nonaffinity(X) <= tol);
    flag[e] = (unsigned char)is;
  }
  const unsigned m = __ballot_sync(0xffffffffu, is);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_aff, (unsigned long long)__popc(m));
}

int classify_cells(txasm_handle h)
{
  if (h->d_cell_affine) return TXASM_OK;
  int rc = dev_alloc(h, &h->d_cell_affine, (size_t)h->n_cells);
  if (rc) return rc;
  unsigned long long *d_n = nullptr;
  TX_CUDA(h, cudaMalloc(&d_n, sizeof(unsigned long long)));
  TX_CUDA(h, cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), h->stream));
  double tol = h->cfg.affine_tol == 0.0 ? 1e-13 : h->cfg.affine_tol;
  if (h->force_general) tol = -1.0;      // field multipliers: the coefficient varies inside a cell, no exact-integration shortcut
  k_classify<<<(unsigned)((h->n_cells + 127) / 128), 128, 0, h->stream>>>(h->n_cells, h->d_lids, h->d_xyz, tol, h->d_cell_affine, d_n);
  unsigned long long n = 0;
  TX_CUDA(h, cudaMemcpyAsync(&n, d_n, sizeof(n), cudaMemcpyDeviceToHost, h->stream));
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_n);
  h->n_affine = (int64_t)n;
  return TXASM_OK;
}

}  // namespace txasm
