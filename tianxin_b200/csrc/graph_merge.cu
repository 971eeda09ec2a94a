// graph_merge.cu -- the fill graph on the device.
//
// TpetraLinearObjFactory::buildGraph (lof/Panzer_TpetraLinearObjFactory_impl.hpp:534-556) exports the ghosted graph with
// INSERT: an owned row gains the columns that other ranks' cells contribute to it.  The host side negotiates WHICH
// (row, column) pairs arrive (a message of surface size, txhost.cpp); this file inserts them into the device-resident
// CSR graph -- sort + unique of the pairs, per-row counts, scan, one merge pass -- and returns, for every pair, the index
// of its entry in the new A_values (the static plan ghostToGlobalContainer's ADD of the matrix uses).  The 2 GB graph
// never visits the host.
#include "txasm_internal.hpp"
#include <cub/cub.cuh>
#include <algorithm>

namespace txasm {

__global__ void k_pair_keys(int64_t n, const int *__restrict__ rows, const int *__restrict__ cols, unsigned long long *__restrict__ keys)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((unsigned long long)(unsigned)rows[i] << 32) | (unsigned)cols[i];
}
__device__ __forceinline__ bool row_has(const int *__restrict__ b, int len, int c)
{
  int lo = 0, hi = len;
  while (lo < hi) { const int m = (lo + hi) >> 1; if (b[m] < c) lo = m + 1; else hi = m; }
  return lo < len && b[lo] == c;
}
// add[row] += 1 for every unique pair whose column the row does not have yet; such keys are flagged
__global__ void k_pair_count(int64_t m, const unsigned long long *__restrict__ keys, const int64_t *__restrict__ rowptr,
                             const int *__restrict__ colind, int64_t n_rows, int *__restrict__ add, unsigned char *__restrict__ is_new,
                             int *__restrict__ bad)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int64_t row = (int64_t)(keys[i] >> 32);
  const int col = (int)(keys[i] & 0xffffffffull);
  if (row >= n_rows || col < 0) { *bad = 1; is_new[i] = 0; return; }
  const int64_t b = rowptr[row];
  const bool has = row_has(colind + b, (int)(rowptr[row + 1] - b), col);
  is_new[i] = has ? 0 : 1;
  if (!has) atomicAdd(&add[row], 1);
}
__global__ void k_new_len(int64_t n_rows, const int64_t *__restrict__ rowptr, const int *__restrict__ add, int64_t *__restrict__ len)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r < n_rows) len[r] = rowptr[r + 1] - rowptr[r] + add[r];
  if (r == n_rows) len[r] = 0;
}
// one warp per row: rows without insertions are copied, the others merged with their (sorted) new columns
__global__ void k_merge_rows(int64_t n_rows, const int64_t *__restrict__ rowptr, const int *__restrict__ colind,
                             const int *__restrict__ add, const int64_t *__restrict__ new_rowptr, int64_t m,
                             const unsigned long long *__restrict__ keys, const unsigned char *__restrict__ is_new,
                             int *__restrict__ new_colind)
{
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const int64_t b = rowptr[r], nb = new_rowptr[r];
  const int len = (int)(rowptr[r + 1] - b);
  if (add[r] == 0) {
    for (int k = lane; k < len; k += 32) new_colind[nb + k] = colind[b + k];
    return;
  }
  if (lane) return;
  int64_t lo = 0, hi = m;                                   // first key of this row
  const unsigned long long k0 = (unsigned long long)r << 32;
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (keys[mid] < k0) lo = mid + 1; else hi = mid; }
  int i = 0;
  int64_t j = lo, o = nb;
  auto next_new = [&]() { while (j < m && (int64_t)(keys[j] >> 32) == r && !is_new[j]) ++j; return j < m && (int64_t)(keys[j] >> 32) == r; };
  bool have = next_new();
  while (i < len || have) {
    const int a = (i < len) ? colind[b + i] : 0x7fffffff;
    const int c = have ? (int)(keys[j] & 0xffffffffull) : 0x7fffffff;
    if (a <= c) { new_colind[o++] = a; ++i; }
    else { new_colind[o++] = c; ++j; have = next_new(); }
  }
}
__global__ void k_pair_pos(int64_t n, const int *__restrict__ rows, const int *__restrict__ cols, const int64_t *__restrict__ rowptr,
                           const int *__restrict__ colind, int64_t *__restrict__ pos)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t b = rowptr[rows[i]];
  const int len = (int)(rowptr[rows[i] + 1] - b);
  int lo = 0, hi = len;
  const int c = cols[i];
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (colind[b + mid] < c) lo = mid + 1; else hi = mid; }
  pos[i] = (lo < len && colind[b + lo] == c) ? b + lo : -1;
}

}  // namespace txasm

using namespace txasm;

#define TX_CHECK_H(h)                                                    \
  do {                                                                   \
    if (!(h)) return TXASM_EINVAL;                                       \
    if ((h)->sticky) return TXASM_ECUDA;                                 \
    cudaError_t e__ = cudaSetDevice((h)->device);                        \
    if (e__ != cudaSuccess) return cuda_fail(h, e__, "cudaSetDevice", __FILE__, __LINE__); \
  } while (0)

// frees p only when the handle allocated it (graph arrays handed over by the caller stay the caller's)
static void free_if_owned(txasm_handle h, const void *p)
{
  if (p && std::find(h->owned.begin(), h->owned.end(), (void *)p) != h->owned.end()) dev_free(h, (void *)p);
}

extern "C" {

int txasm_graph_get_rows(txasm_handle h, int64_t first_row, int64_t n_rows, int64_t *rowptr, int *colind)
{
  TX_CHECK_H(h);
  if (!h->have_graph) return set_err(h, TXASM_ESTATE, "no graph");
  if (first_row < 0 || n_rows < 0 || first_row + n_rows > h->n_rows || !rowptr) return set_err(h, TXASM_EINVAL, "graph_get_rows: bad arguments");
  TX_CUDA(h, copy_to_device_sync(h, rowptr, h->d_rowptr + first_row, sizeof(int64_t) * (size_t)(n_rows + 1)));
  const int64_t b = rowptr[0], e = rowptr[n_rows];
  for (int64_t i = 0; i <= n_rows; ++i) rowptr[i] -= b;
  if (colind && e > b) TX_CUDA(h, copy_to_device_sync(h, colind, h->d_colind + b, sizeof(int) * (size_t)(e - b)));
  return TXASM_OK;
}

int txasm_graph_merge_columns(txasm_handle h, int64_t n, const int *rows, const int *cols, int64_t *pos, int64_t *nnz_out)
{
  TX_CHECK_H(h);
  if (!h->have_graph) return set_err(h, TXASM_ESTATE, "graph_merge_columns before graph_build / graph_set");
  if (n < 0 || (n && (!rows || !cols))) return set_err(h, TXASM_EINVAL, "graph_merge_columns: bad arguments");
  if (h->d_dir_plan) { dev_free(h, h->d_dir_plan); h->d_dir_plan = nullptr; }
  h->is_setup = false;
  if (n == 0) { if (nnz_out) *nnz_out = h->nnz; return TXASM_OK; }
  const int64_t nr = h->n_rows;
  const int *d_rows = nullptr, *d_cols = nullptr;
  int rc;
  if ((rc = to_device(h, rows, (size_t)n, &d_rows))) return rc;
  if ((rc = to_device(h, cols, (size_t)n, &d_cols))) return rc;
  unsigned long long *keys = nullptr, *keys2 = nullptr, *uniq = nullptr;
  int64_t *d_m = nullptr, *len = nullptr, *new_rowptr = nullptr;
  int *add = nullptr, *d_bad = nullptr, *new_colind = nullptr;
  unsigned char *is_new = nullptr;
  void *tmp = nullptr;
  auto cleanup = [&]() {
    cudaFree(keys); cudaFree(keys2); cudaFree(uniq); cudaFree(d_m); cudaFree(len); cudaFree(add); cudaFree(d_bad); cudaFree(is_new); cudaFree(tmp);
    if (d_rows != rows) free_if_owned(h, d_rows);
    if (d_cols != cols) free_if_owned(h, d_cols);
  };
#define TXM(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return cuda_fail(h, e__, #call, __FILE__, __LINE__); } } while (0)
  TXM(cudaMalloc(&keys, sizeof(unsigned long long) * (size_t)n));
  TXM(cudaMalloc(&keys2, sizeof(unsigned long long) * (size_t)n));
  TXM(cudaMalloc(&uniq, sizeof(unsigned long long) * (size_t)n));
  TXM(cudaMalloc(&d_m, sizeof(int64_t)));
  TXM(cudaMalloc(&add, sizeof(int) * (size_t)(nr + 1)));
  TXM(cudaMalloc(&d_bad, sizeof(int)));
  TXM(cudaMalloc(&is_new, (size_t)n));
  TXM(cudaMalloc(&len, sizeof(int64_t) * (size_t)(nr + 1)));
  TXM(cudaMemsetAsync(add, 0, sizeof(int) * (size_t)(nr + 1), h->stream));
  TXM(cudaMemsetAsync(d_bad, 0, sizeof(int), h->stream));
  k_pair_keys<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, d_rows, d_cols, keys);
  size_t tb = 0, tb2 = 0, tb3 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tb, keys, keys2, (int)n, 0, 64, h->stream);
  cub::DeviceSelect::Unique(nullptr, tb2, keys2, uniq, d_m, (int)n, h->stream);
  cub::DeviceScan::ExclusiveSum(nullptr, tb3, len, len, (int)(nr + 1), h->stream);
  tb = std::max(tb, std::max(tb2, tb3));
  TXM(cudaMalloc(&tmp, tb ? tb : 1));
  TXM(cub::DeviceRadixSort::SortKeys(tmp, tb, keys, keys2, (int)n, 0, 64, h->stream));
  TXM(cub::DeviceSelect::Unique(tmp, tb, keys2, uniq, d_m, (int)n, h->stream));
  int64_t m = 0;
  TXM(copy_to_device_sync(h, &m, d_m, sizeof(int64_t)));
  k_pair_count<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(m, uniq, h->d_rowptr, h->d_colind, nr, add, is_new, d_bad);
  int bad = 0;
  TXM(copy_to_device_sync(h, &bad, d_bad, sizeof(int)));
  if (bad) { cleanup(); return set_err(h, TXASM_EINVAL, "graph_merge_columns: a row index is out of range"); }
  k_new_len<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, h->stream>>>(nr, h->d_rowptr, add, len);
  if ((rc = dev_alloc(h, &new_rowptr, (size_t)(nr + 1)))) { cleanup(); return rc; }
  TXM(cub::DeviceScan::ExclusiveSum(tmp, tb, len, new_rowptr, (int)(nr + 1), h->stream));
  int64_t nnz = 0;
  TXM(copy_to_device_sync(h, &nnz, new_rowptr + nr, sizeof(int64_t)));
  if ((rc = dev_alloc(h, &new_colind, (size_t)nnz))) { cleanup(); return rc; }
  k_merge_rows<<<(unsigned)((nr * 32 + 255) / 256), 256, 0, h->stream>>>(nr, h->d_rowptr, h->d_colind, add, new_rowptr, m, uniq, is_new, new_colind);
  TXM(cudaGetLastError());
  if (pos) {
    int64_t *d_pos = nullptr;
    const bool pos_dev = is_device_ptr(pos);
    if (pos_dev) d_pos = pos; else TXM(cudaMalloc(&d_pos, sizeof(int64_t) * (size_t)n));
    k_pair_pos<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, d_rows, d_cols, new_rowptr, new_colind, d_pos);
    if (!pos_dev) {
      cudaError_t e = copy_to_device_sync(h, pos, d_pos, sizeof(int64_t) * (size_t)n);
      cudaFree(d_pos);
      TXM(e);
    }
  }
  TXM(cudaStreamSynchronize(h->stream));
#undef TXM
  free_if_owned(h, h->d_rowptr);         // (arrays the caller handed over with txasm_graph_set stay the caller's)
  free_if_owned(h, h->d_colind);
  h->d_rowptr = new_rowptr; h->d_colind = new_colind; h->nnz = nnz;
  cleanup();
  if (nnz_out) *nnz_out = nnz;
  return TXASM_OK;
}

}  // extern "C"
