// txasm_capi.cu -- the extern "C" boundary declared in include/txasm.h.
#include "txasm_internal.hpp"
#include <cstdarg>
#include <cstring>
#include <algorithm>
#include <chrono>

namespace txasm {

static thread_local std::string g_create_err;

int set_err(txasm_handle h, int code, const char *fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_err = buf;
  return code;
}

int cuda_fail(txasm_handle h, cudaError_t e, const char *what, const char *file, int line)
{
  if (h) h->sticky = true;
  return set_err(h, TXASM_ECUDA, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
}

bool is_device_ptr(const void *p)
{
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

void dev_free(txasm_handle h, void *p)
{
  if (!p) return;
  auto it = std::find(h->owned.begin(), h->owned.end(), p);
  if (it != h->owned.end()) h->owned.erase(it);
  cudaFree(p);
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

// stage-timer events (txasm_timers_get, txasm_last_fill_ms): skipped when the option stage_timers is off
#define TX_TIME_EV(i) do { if (h->opt_timers) cudaEventRecord(h->ev[i], h->stream); } while (0)

static inline void fill_ring_record(txasm_handle h, int k)
{
  if (h->fill_ring.empty()) return;
  cudaEventRecord(h->fill_ring[(size_t)(h->fill_ring_count % h->fill_ring_n) * 4 + k], h->stream);
}

extern "C" {

int txasm_version(int *major, int *minor)
{
  if (major) *major = TXASM_VERSION_MAJOR;
  if (minor) *minor = TXASM_VERSION_MINOR;
  return TXASM_OK;
}

const char *txasm_last_error(txasm_handle h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int txasm_create(const txasm_config *cfg, txasm_handle *out)
{
  if (!out) return TXASM_EINVAL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_err(nullptr, TXASM_ECUDA, "no CUDA device available (%s); txasm has no CPU fallback",
                   e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  txasm_handle h = new (std::nothrow) txasm_handle_s();
  if (!h) return TXASM_ENOMEM;
  if (cfg) h->cfg = *cfg;
  h->device = h->cfg.device;
  if (h->device < 0 || h->device >= ndev) { delete h; return set_err(nullptr, TXASM_EINVAL, "device %d out of range", h->device); }
  e = cudaSetDevice(h->device);
  if (e != cudaSuccess) { int rc = cuda_fail(nullptr, e, "cudaSetDevice", __FILE__, __LINE__); delete h; return rc; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, h->device);
  if (prop.major < 10) {
    delete h;
    return set_err(nullptr, TXASM_EUNSUPPORTED, "device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
  }
  h->n_sm = prop.multiProcessorCount;
  h->smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (h->cfg.stream) { h->stream = (cudaStream_t)h->cfg.stream; h->own_stream = false; }
  else {
    // a BLOCKING stream: it orders itself against the legacy default stream, on which most callers (Kokkos,
    // torch) produce the arrays they hand over
    e = cudaStreamCreate(&h->stream);
    if (e != cudaSuccess) { int rc = cuda_fail(nullptr, e, "cudaStreamCreate", __FILE__, __LINE__); delete h; return rc; }
    h->own_stream = true;
  }
  for (auto &ev : h->ev) cudaEventCreate(&ev);
  {
    auto env = [](const char *n) { const char *e = getenv(n); return e && e[0] == '1'; };
    h->opt_uniform = env("TXASM_NO_UNIFORM_KERNEL") ? 0 : 1;
    h->opt_brick = env("TXASM_NO_BRICK_KERNEL") ? 0 : 1;
    h->opt_fuse_dir = env("TXASM_NO_FUSE_DIRICHLET") ? 0 : 1;
    h->opt_concurrent = env("TXASM_NO_CONCURRENT_FILL") ? 0 : 1;
    h->opt_edge = env("TXASM_NO_EDGE_KERNEL") ? 0 : 1;
    const char *ov = getenv("TXASM_EXPORT_OVERLAP");
    h->opt_overlap = ov ? (ov[0] == '1') : 1;
  }
  *out = h;
  return TXASM_OK;
}

int txasm_destroy(txasm_handle h)
{
  if (!h) return TXASM_EINVAL;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  tiles_free(h);
  halo_free(h);
  gblocks_free(h);
  for (void *p : h->owned) cudaFree(p);
  for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
  for (auto &e : h->fill_ring) if (e) cudaEventDestroy(e);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return TXASM_OK;
}

int txasm_block_add(txasm_handle h, int topology, int basis, int cubature_degree, int64_t n_cells, int dofs_per_cell,
                    const int *lids, const double *cell_coords, const double *node_coords, int64_t n_rows)
{
  TX_CHECK_H(h);
  if (h->have_block) return set_err(h, TXASM_EUNSUPPORTED, "one element block per handle in this version");
  if (topology != TXASM_TOPO_HEX8 || basis != TXASM_BASIS_HGRAD_C1 || dofs_per_cell != 8)
    return set_err(h, TXASM_EUNSUPPORTED, "only HEX8 / HGRAD C1 / one scalar field (8 DOFs per cell) is implemented");
  if (cubature_degree != 2 && cubature_degree != 3)
    return set_err(h, TXASM_EUNSUPPORTED, "cubature degree %d: only the 2x2x2 Gauss rule (degree 2 or 3) is implemented", cubature_degree);
  if (n_cells <= 0 || n_rows <= 0 || !lids || (!cell_coords && !node_coords)) return set_err(h, TXASM_EINVAL, "block_add: bad arguments");
  if (n_rows > 0x7fffffffLL) return set_err(h, TXASM_EINVAL, "n_rows exceeds LocalOrdinal range");
  h->n_cells = n_cells; h->n_rows = n_rows;
  int rc = to_device(h, lids, (size_t)n_cells * 8, &h->d_lids);
  if (rc) return rc;
  rc = dev_alloc(h, &h->d_xyz, (size_t)n_rows * 3);
  if (rc) return rc;
  if (node_coords) {
    TX_CUDA(h, cudaMemcpyAsync(h->d_xyz, node_coords, sizeof(double) * n_rows * 3, cudaMemcpyDefault, h->stream));
    TX_CUDA(h, cudaStreamSynchronize(h->stream));
  } else {
    const double *d_cc = nullptr;
    const size_t before = h->owned.size();
    rc = to_device(h, cell_coords, (size_t)n_cells * 24, &d_cc);
    if (rc) return rc;
    rc = build_node_coords(h, d_cc);
    if (rc) return rc;
    TX_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->owned.size() > before) dev_free(h, (void *)d_cc);   // staging copy of the per-cell coordinates
  }
  h->have_block = true;
  h->is_setup = false;
  return TXASM_OK;
}

int txasm_graph_set(txasm_handle h, int64_t n_rows, const int64_t *rowptr, const int *colind)
{
  if (h && h->d_dir_plan) { dev_free(h, h->d_dir_plan); h->d_dir_plan = nullptr; }   // Dirichlet row plan follows the graph
  TX_CHECK_H(h);
  if (!h->have_block && !gblocks_count(h)) return set_err(h, TXASM_ESTATE, "graph_set before block_add");
  if (n_rows != h->n_rows || !rowptr || !colind) return set_err(h, TXASM_EINVAL, "graph_set: bad arguments");
  int rc = to_device(h, rowptr, (size_t)n_rows + 1, &h->d_rowptr);
  if (rc) return rc;
  TX_CUDA(h, copy_to_device_sync(h, &h->nnz, h->d_rowptr + n_rows, sizeof(int64_t)));
  rc = to_device(h, colind, (size_t)h->nnz, &h->d_colind);
  if (rc) return rc;
  h->have_graph = true;
  h->is_setup = false;
  return TXASM_OK;
}

int txasm_graph_build(txasm_handle h, int64_t *nnz_out)
{
  if (h && h->d_dir_plan) { dev_free(h, h->d_dir_plan); h->d_dir_plan = nullptr; }   // Dirichlet row plan follows the graph
  TX_CHECK_H(h);
  if (!h->have_block && !gblocks_count(h)) return set_err(h, TXASM_ESTATE, "graph_build before block_add");
  if (h->have_graph) { if (nnz_out) *nnz_out = h->nnz; return TXASM_OK; }
  h->is_setup = false;
  if (gblocks_count(h)) return gblocks_graph_build(h, nnz_out);
  return build_graph_device(h, nnz_out);
}

int txasm_graph_get(txasm_handle h, int64_t *rowptr, int *colind)
{
  TX_CHECK_H(h);
  if (!h->have_graph) return set_err(h, TXASM_ESTATE, "no graph");
  if (rowptr) TX_CUDA(h, copy_to_device_sync(h, rowptr, h->d_rowptr, sizeof(int64_t) * (h->n_rows + 1)));
  if (colind) TX_CUDA(h, copy_to_device_sync(h, colind, h->d_colind, sizeof(int) * h->nnz));
  return TXASM_OK;
}

int txasm_terms_set(txasm_handle h, const txasm_term *terms, int n)
{
  TX_CHECK_H(h);
  if (n < 0 || (n && !terms)) return set_err(h, TXASM_EINVAL, "terms_set: bad arguments");
  std::vector<txasm_term> t(terms, terms + n);
  std::vector<const double *> ip(n, nullptr), fm(n, nullptr);
  bool any_fm = false;
  int nsrc = 0;
  for (int i = 0; i < n; ++i) {
    switch (t[i].kind) {
      case TXASM_TERM_GRADGRAD: case TXASM_TERM_MASS: case TXASM_TERM_TRANSIENT_MASS:
        if (t[i].gather_seed_index1 < 0) return set_err(h, TXASM_EINVAL, "term %d: bad gather seed index", i);
        if (t[i].vec < 0 || t[i].vec > 2) return set_err(h, TXASM_EINVAL, "term %d: bad vec", i);
        if (t[i].field_multiplier_ip) {
          if (!h->have_block) return set_err(h, TXASM_EINVAL, "term %d: field multipliers need a block", i);
          for (int j = 0; j < i && !fm[i]; ++j) if (t[j].field_multiplier_ip == t[i].field_multiplier_ip) fm[i] = fm[j];   // one copy per array
          if (!fm[i]) { int rc = to_device(h, t[i].field_multiplier_ip, (size_t)h->n_cells * NQ, &fm[i]); if (rc) return rc; }
          any_fm = true;
        }
        break;
      case TXASM_TERM_SOURCE:
        if (++nsrc > MAX_SRC) return set_err(h, TXASM_EUNSUPPORTED, "more than %d source terms", MAX_SRC);
        if (t[i].source_id == TXASM_SOURCE_IP_ARRAY) {
          if (!t[i].ip_values || !h->have_block) return set_err(h, TXASM_EINVAL, "term %d: ip_values needs a block and a pointer", i);
          int rc = to_device(h, t[i].ip_values, (size_t)h->n_cells * NQ, &ip[i]);
          if (rc) return rc;
        } else if (t[i].source_id != TXASM_SOURCE_SIN3 && t[i].source_id != TXASM_SOURCE_CONSTANT)
          return set_err(h, TXASM_EUNSUPPORTED, "term %d: unknown source id %d", i, t[i].source_id);
        break;
      default: return set_err(h, TXASM_EUNSUPPORTED, "term %d: kind %d not implemented", i, t[i].kind);
    }
  }
  h->terms = t;
  h->d_src_ip = ip;
  h->d_field_mult = fm;
  if (any_fm != h->force_general) {      // the affine classification depends on it
    h->force_general = any_fm;
    if (h->d_cell_affine) { dev_free(h, h->d_cell_affine); h->d_cell_affine = nullptr; }
    h->is_setup = false;
  }
  return TXASM_OK;
}

int txasm_dirichlet_set(txasm_handle h, int n, const int *local_dofs, const double *values)
{
  TX_CHECK_H(h);
  if (n < 0 || (n && (!local_dofs || !values))) return set_err(h, TXASM_EINVAL, "dirichlet_set: bad arguments");
  if (h->d_dir_dofs) { dev_free(h, h->d_dir_dofs); h->d_dir_dofs = nullptr; }
  if (h->d_dir_vals) { dev_free(h, h->d_dir_vals); h->d_dir_vals = nullptr; }
  if (h->d_dir_plan) { dev_free(h, h->d_dir_plan); h->d_dir_plan = nullptr; }
  h->overlap_state = 0;
  h->dir_fusable = -1;
  h->n_dir = n;
  if (n == 0) return TXASM_OK;
  int rc = dev_alloc(h, &h->d_dir_dofs, (size_t)n);
  if (rc) return rc;
  rc = dev_alloc(h, &h->d_dir_vals, (size_t)n);
  if (rc) return rc;
  TX_CUDA(h, cudaMemcpyAsync(h->d_dir_dofs, local_dofs, sizeof(int) * n, cudaMemcpyDefault, h->stream));
  TX_CUDA(h, cudaMemcpyAsync(h->d_dir_vals, values, sizeof(double) * n, cudaMemcpyDefault, h->stream));
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  return TXASM_OK;
}

int txasm_neumann_set(txasm_handle h, int n, const int *cells, const int *local_sides, const double *values)
{
  TX_CHECK_H(h);
  if (n < 0 || (n && (!cells || !local_sides || !values))) return set_err(h, TXASM_EINVAL, "neumann_set: bad arguments");
  if (h->d_neu_cells) { dev_free(h, h->d_neu_cells); h->d_neu_cells = nullptr; }
  if (h->d_neu_sides) { dev_free(h, h->d_neu_sides); h->d_neu_sides = nullptr; }
  if (h->d_neu_vals) { dev_free(h, h->d_neu_vals); h->d_neu_vals = nullptr; }
  h->n_neu = n;
  if (n == 0) return TXASM_OK;
  int rc;
  if ((rc = dev_alloc(h, &h->d_neu_cells, (size_t)n))) return rc;
  if ((rc = dev_alloc(h, &h->d_neu_sides, (size_t)n))) return rc;
  if ((rc = dev_alloc(h, &h->d_neu_vals, (size_t)n))) return rc;
  {
    std::vector<int> hs((size_t)n);
    TX_CUDA(h, copy_to_device_sync(h, hs.data(), local_sides, sizeof(int) * n));
    for (int s : hs) if (s < 0 || s > 5) return set_err(h, TXASM_EINVAL, "neumann_set: side ordinal %d is not a hexahedron side", s);
    TX_CUDA(h, copy_to_device_sync(h, h->d_neu_sides, hs.data(), sizeof(int) * n));
  }
  TX_CUDA(h, cudaMemcpyAsync(h->d_neu_cells, cells, sizeof(int) * n, cudaMemcpyDefault, h->stream));
  TX_CUDA(h, cudaMemcpyAsync(h->d_neu_vals, values, sizeof(double) * n, cudaMemcpyDefault, h->stream));
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  return TXASM_OK;
}

int txasm_cload_set(txasm_handle h, int n, const int *local_dofs, const double *values)
{
  TX_CHECK_H(h);
  if (n < 0 || (n && (!local_dofs || !values))) return set_err(h, TXASM_EINVAL, "cload_set: bad arguments");
  if (h->d_cload_dofs) { dev_free(h, h->d_cload_dofs); h->d_cload_dofs = nullptr; }
  if (h->d_cload_vals) { dev_free(h, h->d_cload_vals); h->d_cload_vals = nullptr; }
  h->n_cload = n;
  if (n == 0) return TXASM_OK;
  int rc = dev_alloc(h, &h->d_cload_dofs, (size_t)n);
  if (rc) return rc;
  rc = dev_alloc(h, &h->d_cload_vals, (size_t)n);
  if (rc) return rc;
  TX_CUDA(h, cudaMemcpyAsync(h->d_cload_dofs, local_dofs, sizeof(int) * n, cudaMemcpyDefault, h->stream));
  TX_CUDA(h, cudaMemcpyAsync(h->d_cload_vals, values, sizeof(double) * n, cudaMemcpyDefault, h->stream));
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  return TXASM_OK;
}

int txasm_setup(txasm_handle h)
{
  if (h) { h->overlap_state = 0; h->dir_fusable = -1; }
  TX_CHECK_H(h);
  const auto t0 = std::chrono::steady_clock::now();
  if (gblocks_count(h)) {                  // general element blocks: scatter plan only
    if (!h->have_graph) return set_err(h, TXASM_ESTATE, "setup needs a graph (txasm_graph_set)");
    int rc = gblocks_setup(h);
    if (rc) return rc;
    h->mode = TXASM_SCATTER_GENERIC;
    TX_CUDA(h, cudaStreamSynchronize(h->stream));
    h->setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    h->is_setup = true;
    return TXASM_OK;
  }
  if (!h->have_block || !h->have_graph) return set_err(h, TXASM_ESTATE, "setup needs a block and a graph");
  int rc = build_adjacency(h);
  if (rc) return rc;
  rc = classify_cells(h);
  if (rc) return rc;
  int want = h->cfg.scatter_mode;
  if (want == TXASM_SCATTER_AUTO || want == TXASM_SCATTER_ROWTILE) {
    rc = tiles_build(h);
    if (rc == TXASM_OK) h->mode = TXASM_SCATTER_ROWTILE;
    else if (want == TXASM_SCATTER_ROWTILE || rc != TXASM_EUNSUPPORTED) return rc;
    else h->mode = TXASM_SCATTER_ROWGATHER;
  } else if (want == TXASM_SCATTER_ATOMIC || want == TXASM_SCATTER_ROWGATHER) h->mode = want;
  else return set_err(h, TXASM_EINVAL, "unknown scatter mode %d", want);
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  h->setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  h->is_setup = true;
  return TXASM_OK;
}

int txasm_info_get(txasm_handle h, txasm_info *info)
{
  if (!h || !info) return TXASM_EINVAL;
  memset(info, 0, sizeof(*info));
  info->n_cells = h->n_cells; info->n_rows = h->n_rows; info->nnz = h->nnz;
  info->n_affine_cells = h->n_affine;
  info->scatter_mode = h->mode;
  info->kernel_launches_last_evaluate = h->launches;
  info->n_sm = h->n_sm;
  info->uniform_kernel_used = h->uniform_used; info->dirichlet_fused = h->dir_fused ? 1 : 0;
  info->export_overlapped = h->overlap_used ? 1 : 0; info->setup_ms = h->setup_ms;
  if (h->tiles) tiles_info(h, info);
  return TXASM_OK;
}

// resolve a caller vector: device pointer -> itself; host pointer -> staging copy (H2D when `in`)
static int stage_in(txasm_handle h, const double *p, size_t n, double **stage, const double **out)
{
  if (!p) { *out = nullptr; return TXASM_OK; }
  if (is_device_ptr(p)) { *out = p; return TXASM_OK; }
  if (!*stage) { int rc = dev_alloc(h, stage, n); if (rc) return rc; }
  TX_CUDA(h, cudaMemcpyAsync(*stage, p, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  *out = *stage;
  return TXASM_OK;
}

static int ensure_side_stream(txasm_handle h)
{
  if (h->side_stream) return TXASM_OK;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  TX_CUDA(h, cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, hi));
  return TXASM_OK;
}

int txasm_evaluate(txasm_handle h, int eval_type, int flags, const txasm_inargs *in,
                   const double *x, const double *xdot, const double *xdotdot, double *f, double *A_values)
{
  TX_CHECK_H(h);
  if (!h->is_setup) return set_err(h, TXASM_ESTATE, "evaluate before setup");
  if (!in || flags <= 0 || flags > TXASM_FLAG_ALL) return set_err(h, TXASM_EINVAL, "evaluate: bad inargs / flags");
  if (eval_type != TXASM_RESIDUAL && eval_type != TXASM_JACOBIAN) return set_err(h, TXASM_EINVAL, "evaluate: bad eval_type");
  const int jac = (eval_type == TXASM_JACOBIAN);
  if (jac && !A_values) return set_err(h, TXASM_EINVAL, "Jacobian evaluation needs A_values");
  h->launches = 0;
  h->uniform_used = 0; h->dir_fused = false; h->overlap_used = false; h->neu_recorded = false;

  // consolidate the term list into coefficients
  FillCoef c;
  memset(&c, 0, sizeof(c));
  double seed[3] = {in->beta, in->alpha, in->gamma};
  for (size_t i = 0; i < h->terms.size(); ++i) {
    const txasm_term &t = h->terms[i];
    // Integrator_TransientBasisTimesScalar skips its contribution unless workset.evaluate_transient_terms
    // (disc-fe/src/evaluators/Panzer_Integrator_TransientBasisTimesScalar_impl.hpp: evaluateFields)
    if (t.kind == TXASM_TERM_TRANSIENT_MASS && !in->evaluate_transient_terms) continue;
    // GatherSolution_Tpetra<Jacobian> seed choice (Panzer_GatherSolution_Tpetra_impl.hpp:554-572): gather_seeds[i] when
    // the gather has a "Gather Seed Index" >= 0, else alpha for the time-derivative vector, beta otherwise
    double sd = 0.0;
    if (t.kind == TXASM_TERM_GRADGRAD || t.kind == TXASM_TERM_MASS || t.kind == TXASM_TERM_TRANSIENT_MASS) {
      sd = seed[t.vec];
      if (t.gather_seed_index1 > 0) {
        if (t.gather_seed_index1 > in->n_gather_seeds || !in->gather_seeds)
          return set_err(h, TXASM_EINVAL, "term %d wants gather_seeds[%d] but inargs carries %d seeds", (int)i, t.gather_seed_index1 - 1, in->n_gather_seeds);
        sd = in->gather_seeds[t.gather_seed_index1 - 1];
      }
    }
    if (t.kind == TXASM_TERM_GRADGRAD) {
      if (c.kg[0] == 0.0 && c.kg[1] == 0.0 && c.kg[2] == 0.0) c.fmK = h->d_field_mult[i];
      else if (c.fmK != h->d_field_mult[i]) return set_err(h, TXASM_EUNSUPPORTED, "the GRADGRAD terms must share one field-multiplier array");
      c.kg[t.vec] += t.multiplier; c.cK += t.multiplier * sd;
    } else if (t.kind == TXASM_TERM_MASS || t.kind == TXASM_TERM_TRANSIENT_MASS) {
      if (c.km[0] == 0.0 && c.km[1] == 0.0 && c.km[2] == 0.0) c.fmM = h->d_field_mult[i];
      else if (c.fmM != h->d_field_mult[i]) return set_err(h, TXASM_EUNSUPPORTED, "the MASS terms must share one field-multiplier array");
      c.km[t.vec] += t.multiplier; c.cM += t.multiplier * sd;
    }
    else if (t.kind == TXASM_TERM_SOURCE) {
      c.src_id[c.n_src] = t.source_id; c.src_mult[c.n_src] = t.multiplier; c.src_ip[c.n_src] = h->d_src_ip[i]; c.n_src++;
    }
  }
  for (int v = 0; v < 3; ++v) {
    c.has_vec[v] = (c.kg[v] != 0.0 || c.km[v] != 0.0);
    if (c.km[v] != 0.0) c.has_mass = 1;
  }
  if (jac && c.cM != 0.0) c.has_mass = 1;
  const double *xin[3] = {x, xdot, xdotdot};
  const bool generic = (h->mode == TXASM_SCATTER_GENERIC);
  if (generic) for (int v = 0; v < 3; ++v) c.has_vec[v] = xin[v] != nullptr;      // the blocks' operators say what they read
  for (int v = 0; v < 3; ++v)
    if (!generic && c.has_vec[v] && !xin[v]) return set_err(h, TXASM_EINVAL, "a term reads solution vector %d but it is NULL", v);
  if (generic && h->n_neu > 0) return set_err(h, TXASM_EUNSUPPORTED, "Neumann side sets are implemented for the Q1 hexahedron block only");

  FillArgs a;
  memset(&a, 0, sizeof(a));
  a.n_cells = h->n_cells; a.n_rows = h->n_rows; a.lids = h->d_lids; a.xyz = h->d_xyz;
  a.rowptr = h->d_rowptr; a.colind = h->d_colind; a.jacobian = jac; a.c = c;
  int rc;
  {
    static const bool timeline = [] { const char *e = getenv("TXASM_TIMELINE"); return e && e[0] == '1'; }();
    if (timeline) {
      if (!h->d_dbg) { rc = dev_alloc(h, &h->d_dbg, 8); if (rc) return rc; }
      const unsigned long long init[8] = {~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0};
      TX_CUDA(h, cudaMemcpyAsync(h->d_dbg, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
      a.dbg = h->d_dbg;
    }
  }
  bool x_host[3] = {false, false, false};
  for (int v = 0; v < 3; ++v) {
    const double *src = (c.has_vec[v] || v == 0) ? xin[v] : nullptr;
    x_host[v] = src && !is_device_ptr(src);
    rc = stage_in(h, src, (size_t)h->n_rows, &h->st_x[v], &a.x[v]);
    if (rc) return rc;
  }
  const bool f_host = f && !is_device_ptr(f), A_host = jac && !is_device_ptr(A_values);
  if (f_host && !h->st_f) { rc = dev_alloc(h, &h->st_f, (size_t)h->n_rows); if (rc) return rc; }
  if (A_host && !h->st_A) { rc = dev_alloc(h, &h->st_A, (size_t)h->nnz); if (rc) return rc; }
  a.f = f ? (f_host ? h->st_f : f) : nullptr;
  a.A = jac ? (A_host ? h->st_A : A_values) : nullptr;
  // Host outputs are staged.  Unless this call overwrites every entry (a volume fill by an owner-computes mode, or
  // by the atomic mode with zero_outputs), the staging buffer must start from the caller's contents: split-stage
  // calls (BoundaryFill or Scatter alone) and accumulating atomic fills read-modify-write f and A.
  {
    const bool owner = (h->mode == TXASM_SCATTER_ROWTILE || h->mode == TXASM_SCATTER_ROWGATHER || (generic && !h->opt_block_atomic));
    const bool overwrites = (flags & TXASM_FLAG_VOLUMETRIC_FILL) && (owner || in->zero_outputs);
    if (!overwrites) {
      if (f_host) TX_CUDA(h, cudaMemcpyAsync(h->st_f, f, sizeof(double) * h->n_rows, cudaMemcpyHostToDevice, h->stream));
      if (A_host) TX_CUDA(h, cudaMemcpyAsync(h->st_A, A_values, sizeof(double) * h->nnz, cudaMemcpyHostToDevice, h->stream));
    }
  }

  const bool rowtile = (h->mode == TXASM_SCATTER_ROWTILE);
  const bool vol = (flags & TXASM_FLAG_VOLUMETRIC_FILL) != 0, bnd = (flags & TXASM_FLAG_BOUNDARY_FILL) != 0;
  int e_brick = 0, e_uni = 0, e_edge = 0;
  if (rowtile) fill_ranges(h, a, &e_brick, &e_uni, &e_edge);
  const bool have_uni = e_uni > 0;
  // Dirichlet rows written by the fill kernel itself: both stages requested, nothing else in the boundary stage that
  // would have to run between them (Neumann and concentrated loads precede Dirichlet in the reference's order)
  bool fuse_dir = false;
  if (h->opt_fuse_dir && rowtile && vol && bnd && h->n_dir > 0 && h->n_neu == 0 && (h->n_cload == 0 || jac) && a.x[0] && a.f) {
    if (h->dir_fusable < 0) { rc = dirichlet_fuse_prepare(h); if (rc) return rc; }
    fuse_dir = h->dir_fusable == 1;
  }
  h->dir_fused = fuse_dir;

  // Overlapped schedule (all four stages, neighbours present, uniform tile range in use): the export only touches
  // ghost rows and the owned rows on rank interfaces, none of which lies in a uniform tile, and the same holds for
  // the Dirichlet rows.  So: import | tiles outside the uniform range | Dirichlet | export on a side stream, under
  // the uniform-tile kernels on the main stream.  Same numbers as the sequential order; the export's latency disappears.
  bool overlap = false;
  if (h->opt_overlap && flags == TXASM_FLAG_ALL && rowtile && halo_n_neighbours(h) > 0 && h->n_neu == 0 && have_uni) {
    if (h->overlap_state == 0) {
      bool t1 = false, t2 = false;
      rc = halo_rows_touch_uniform_tiles(h, &t1);
      if (rc) return rc;
      if (h->n_dir) { rc = rows_touch_uniform_tiles(h, h->d_dir_dofs, h->n_dir, &t2); if (rc) return rc; }
      h->overlap_state = (t1 || t2) ? 2 : 1;
    }
    overlap = h->overlap_state == 1;
  }
  // Without a halo to hide: the tiles on the boundary on a side stream beside the uniform-tile kernels (disjoint rows)
  const bool concurrent = !overlap && h->opt_concurrent && rowtile && vol && have_uni && h->tiles && e_uni < tiles_count(h);
  if (overlap || concurrent) { rc = ensure_side_stream(h); if (rc) return rc; }
  h->overlap_used = overlap;

  TX_TIME_EV(0);
  if (flags & TXASM_FLAG_INITIALIZE) {
    double *xs[3] = {(double *)a.x[0], (double *)a.x[1], (double *)a.x[2]};
    rc = halo_import(h, xs);
    if (rc) return rc;
    // the ghost tail of a host x is part of the caller's ghosted container (globalToGhostContainer writes it)
    const int64_t no = halo_n_owned(h);
    if (halo_n_neighbours(h) > 0 && no < h->n_rows)
      for (int v = 0; v < 3; ++v)
        if (x_host[v])
          TX_CUDA(h, cudaMemcpyAsync((double *)xin[v] + no, a.x[v] + no, sizeof(double) * (h->n_rows - no), cudaMemcpyDeviceToHost, h->stream));
  }
  TX_TIME_EV(1);
  if (overlap) {
    TX_TIME_EV(5);
    fill_ring_record(h, 0);
    rc = launch_fill_rowtile(h, a, FILL_REST, h->stream, fuse_dir);
    if (rc) return rc;
    TX_TIME_EV(6);
    fill_ring_record(h, 1);
    TX_TIME_EV(2);
    if (jac == 0 && h->n_cload > 0) { rc = launch_cload(h, a.f); if (rc) return rc; }
    if (h->n_dir > 0 && !fuse_dir) { rc = launch_dirichlet(h, jac, a.x[0], a.f, a.A); if (rc) return rc; }
    TX_TIME_EV(3);
    cudaEventRecord(h->ev[8], h->stream);                       // fork
    TX_CUDA(h, cudaStreamWaitEvent(h->side_stream, h->ev[8], 0));
    {
      cudaStream_t main_stream = h->stream;
      h->stream = h->side_stream;
      rc = halo_export(h, a.f, a.A, jac);
      h->stream = main_stream;
      if (rc) return rc;
    }
    cudaEventRecord(h->ev[9], h->side_stream);
    TX_TIME_EV(10);
    fill_ring_record(h, 2);
    h->brick_ctas_limit = 3;             // leave room on every SM for the (small) exchange kernels
    rc = launch_fill_rowtile(h, a, FILL_UNIFORM, h->stream, false);
    h->brick_ctas_limit = 0;
    if (rc) return rc;
    TX_TIME_EV(11);
    fill_ring_record(h, 3);
    TX_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev[9], 0));   // join
    TX_TIME_EV(4);
  } else {
    if (vol) {
      const bool overwrite = (h->mode == TXASM_SCATTER_ROWTILE || h->mode == TXASM_SCATTER_ROWGATHER || (generic && !h->opt_block_atomic));
      if (!overwrite && in->zero_outputs) {
        if (a.f) TX_CUDA(h, cudaMemsetAsync(a.f, 0, sizeof(double) * h->n_rows, h->stream));
        if (a.A) TX_CUDA(h, cudaMemsetAsync(a.A, 0, sizeof(double) * h->nnz, h->stream));
        h->launches += (a.f ? 1 : 0) + (a.A ? 1 : 0);
      }
      TX_TIME_EV(5);
      fill_ring_record(h, 0);
      if (concurrent) {
        cudaEventRecord(h->ev[8], h->stream);                     // fork: boundary tiles on the side stream
        TX_CUDA(h, cudaStreamWaitEvent(h->side_stream, h->ev[8], 0));
        rc = launch_fill_rowtile(h, a, FILL_REST, h->side_stream, fuse_dir);
        if (rc) return rc;
        cudaEventRecord(h->ev[9], h->side_stream);
        static const int side_limit = [] { const char *e = getenv("TXASM_BRICK_SIDE_LIMIT"); return e ? atoi(e) : 0; }();
        if (e_edge > e_uni && e_brick > 0) h->brick_ctas_limit = side_limit;   // (tuning: CTAs per SM left to k_fill_brick beside k_fill_edge)
        rc = launch_fill_rowtile(h, a, FILL_UNIFORM, h->stream, false);
        h->brick_ctas_limit = 0;
        if (rc) return rc;
        TX_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev[9], 0));  // join
      } else if (rowtile) rc = launch_fill_rowtile(h, a, FILL_ALL, h->stream, fuse_dir);
      else if (generic) rc = launch_gblocks(h, jac, in, a.x, a.f, a.A);
      else if (h->mode == TXASM_SCATTER_ROWGATHER) rc = launch_fill_rowgather(h, a);
      else rc = launch_fill_atomic(h, a);
      if (rc) return rc;
      TX_TIME_EV(6);
      fill_ring_record(h, 1);
    }
    TX_TIME_EV(2);
    h->neu_recorded = false;
    if (bnd && h->n_neu > 0) {
      TX_TIME_EV(12);
      rc = launch_neumann(h, a.f);
      if (rc) return rc;
      TX_TIME_EV(13);
      h->neu_recorded = true;
    }
    if (bnd && h->n_cload > 0 && !jac) {   // CLoadEvalautor<Jacobian> is a no-op in the reference
      rc = launch_cload(h, a.f);
      if (rc) return rc;
    }
    if (bnd && h->n_dir > 0 && !fuse_dir) {
      rc = launch_dirichlet(h, jac, a.x[0], a.f, a.A);
      if (rc) return rc;
    }
    TX_TIME_EV(3);
    if (flags & TXASM_FLAG_SCATTER) {
      rc = halo_export(h, a.f, a.A, jac);
      if (rc) return rc;
    }
    TX_TIME_EV(4);
  }
  h->vol_recorded = vol && h->opt_timers;
  h->timers_recorded = h->opt_timers != 0;
  if (vol && !h->fill_ring.empty()) {
    h->fill_ring_segs[h->fill_ring_count % h->fill_ring_n] = overlap ? 2 : 1;
    h->fill_ring_count += 1;
  }
  if (f_host) TX_CUDA(h, cudaMemcpyAsync(f, h->st_f, sizeof(double) * h->n_rows, cudaMemcpyDeviceToHost, h->stream));
  if (A_host) TX_CUDA(h, cudaMemcpyAsync(A_values, h->st_A, sizeof(double) * h->nnz, cudaMemcpyDeviceToHost, h->stream));
  if (f_host || A_host || x_host[0] || x_host[1] || x_host[2]) TX_CUDA(h, cudaStreamSynchronize(h->stream));
  return TXASM_OK;
}

static const struct { const char *name; int txasm_handle_s::*field; } g_options[] = {
  {"uniform_kernel", &txasm_handle_s::opt_uniform}, {"brick_kernel", &txasm_handle_s::opt_brick},
  {"export_overlap", &txasm_handle_s::opt_overlap}, {"fuse_dirichlet", &txasm_handle_s::opt_fuse_dir},
  {"stage_timers", &txasm_handle_s::opt_timers}, {"concurrent_fill", &txasm_handle_s::opt_concurrent}, {"grid_cap", &txasm_handle_s::opt_grid_cap},
  {"brick_ctas_per_sm", &txasm_handle_s::opt_brick_ctas}, {"halo_p2p", &txasm_handle_s::opt_p2p},
  {"rest_ctas_per_sm", &txasm_handle_s::opt_rest_ctas}, {"edge_kernel", &txasm_handle_s::opt_edge}, {"dmma", &txasm_handle_s::opt_dmma}, {"block_atomic", &txasm_handle_s::opt_block_atomic},
};

int txasm_option_set(txasm_handle h, const char *name, int value)
{
  if (!h || !name) return TXASM_EINVAL;
  if (!strcmp(name, "fill_event_ring")) {       // keep the fill spans of the last `value` evaluates (txasm_fill_ms_history)
    for (auto &e : h->fill_ring) cudaEventDestroy(e);
    h->fill_ring.clear(); h->fill_ring_segs.clear();
    h->fill_ring_n = value > 0 ? value : 0;
    h->fill_ring_count = 0;
    h->fill_ring.resize((size_t)h->fill_ring_n * 4);
    h->fill_ring_segs.assign((size_t)h->fill_ring_n, 0);
    for (auto &e : h->fill_ring) TX_CUDA(h, cudaEventCreate(&e));
    return TXASM_OK;
  }
  for (const auto &o : g_options)
    if (!strcmp(name, o.name)) { 
      const bool counted = (o.field == &txasm_handle_s::opt_grid_cap || o.field == &txasm_handle_s::opt_brick_ctas ||
                            o.field == &txasm_handle_s::opt_rest_ctas);
      h->*(o.field) = counted ? (value > 0 ? value : 0) : (value ? 1 : 0);
      return TXASM_OK;
    }
  return set_err(h, TXASM_EINVAL, "unknown option \"%s\"", name);
}

int txasm_option_get(txasm_handle h, const char *name, int *value)
{
  if (!h || !name || !value) return TXASM_EINVAL;
  if (!strcmp(name, "fill_event_ring")) { *value = h->fill_ring_n; return TXASM_OK; }
  for (const auto &o : g_options)
    if (!strcmp(name, o.name)) { *value = h->*(o.field); return TXASM_OK; }
  return set_err(h, TXASM_EINVAL, "unknown option \"%s\"", name);
}

int txasm_response_functional(txasm_handle h, int kind, int solution_id, int cubature_degree, const double *x, double *value)
{
  TX_CHECK_H(h);
  if (!x || !value) return set_err(h, TXASM_EINVAL, "response_functional: x and value are required");
  if (kind < TXASM_RESP_INTEGRAL || kind > TXASM_RESP_H1_ERROR) return set_err(h, TXASM_EINVAL, "response_functional: kind %d", kind);
  if (kind != TXASM_RESP_INTEGRAL && solution_id != TXASM_SOURCE_SIN3 && solution_id != 3)
    return set_err(h, TXASM_EINVAL, "response_functional: exact solution %d", solution_id);
  if (!h->d_lids || !h->d_xyz) return set_err(h, TXASM_ESTATE, "response_functional needs a block");
  const double *xd = nullptr;
  int rc = stage_in(h, x, (size_t)h->n_rows, &h->st_x[0], &xd);
  if (rc) return rc;
  return response_functional(h, kind, solution_id, cubature_degree, xd, value);
}

// TianXin::Response_Integral<Residual>::evaluateFields (disc-fe/src/responses/TianXin_Response_Integral_impl.hpp:106-133)
int txasm_response_integral(txasm_handle h, int cubature_degree, const double *cell_ip_values, double *response_vector, double *value)
{
  TX_CHECK_H(h);
  if (!cell_ip_values) return set_err(h, TXASM_EINVAL, "response_integral: cell_ip_values is required");
  if (!response_vector) return set_err(h, TXASM_ESTATE, "TianXin::Response_Integral: reponse vector not defined. Please call setVector() before calling this method");
  if (!h->d_lids || !h->d_xyz) return set_err(h, TXASM_ESTATE, "response_integral needs a block");
  const int np = cubature_degree / 2 + 1;
  if (cubature_degree < 0 || np > 16) return set_err(h, TXASM_EINVAL, "response_integral: cubature degree %d", cubature_degree);
  const double *d_ip = nullptr;
  double *tmp = nullptr;
  const size_t n = (size_t)h->n_cells * np * np * np;
  if (is_device_ptr(cell_ip_values)) d_ip = cell_ip_values;
  else {
    TX_CUDA(h, cudaMalloc(&tmp, sizeof(double) * n));
    cudaError_t e = copy_to_device_sync(h, tmp, cell_ip_values, sizeof(double) * n);
    if (e != cudaSuccess) { cudaFree(tmp); return cuda_fail(h, e, "copy", __FILE__, __LINE__); }
    d_ip = tmp;
  }
  double glb = 0.0;
  int rc = response_functional(h, TXASM_RESP_IP_ARRAY, 0, cubature_degree, d_ip, &glb);
  if (tmp) cudaFree(tmp);
  if (rc) return rc;
  response_vector[0] += glb;             // tVector_->sumIntoLocalValue(0, glbValue)
  if (value) *value = glb;               // value_.deep_copy(glbValue)
  return TXASM_OK;
}

// profiling (TXASM_TIMELINE=1): start / end of k_fill_brick, k_fill_edge, k_fill_rowtile in the last evaluate, in microseconds
// relative to the earliest start; -1 where a kernel did not run
int txasm_debug_timeline(txasm_handle h, double out[6])
{
  TX_CHECK_H(h);
  if (!out) return TXASM_EINVAL;
  for (int i = 0; i < 6; ++i) out[i] = -1.0;
  if (!h->d_dbg) return TXASM_OK;
  unsigned long long v[8];
  TX_CUDA(h, copy_to_device_sync(h, v, h->d_dbg, sizeof(v)));
  unsigned long long t0 = ~0ull;
  for (int k = 0; k < 3; ++k) if (v[2 * k] != ~0ull && v[2 * k] < t0) t0 = v[2 * k];
  for (int k = 0; k < 3; ++k) if (v[2 * k] != ~0ull) { out[2 * k] = (double)(v[2 * k] - t0) * 1e-3; out[2 * k + 1] = (double)(v[2 * k + 1] - t0) * 1e-3; }
  return TXASM_OK;
}

int txasm_tile_get(txasm_handle h, int tile, int *rows, int *cells, unsigned short *adjl, int *n_cells_out)
{
  TX_CHECK_H(h);
  if (!h->is_setup || !h->tiles) return set_err(h, TXASM_ESTATE, "tile_get: no row tiles (setup with ROWTILE first)");
  return tiles_get(h, tile, rows, cells, adjl, n_cells_out);
}

int txasm_sync(txasm_handle h)
{
  TX_CHECK_H(h);
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  return TXASM_OK;
}

int txasm_timers_get(txasm_handle h, txasm_timers *t)
{
  TX_CHECK_H(h);
  if (!t) return TXASM_EINVAL;
  if (!h->timers_recorded) return set_err(h, TXASM_ESTATE, "no stage timers: the last evaluate ran with option stage_timers = 0 (or there was none)");
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  txasm_timers o{};
  if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) o.evaluate_gather = ms;
  if (cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]) == cudaSuccess) o.evaluate_volume = ms;
  float uni_ms = 0.f;            // overlapped schedule: the uniform tiles are filled after the boundary stage, under the export
  if (h->overlap_used && cudaEventElapsedTime(&uni_ms, h->ev[10], h->ev[11]) == cudaSuccess) o.evaluate_volume += uni_ms;
  float neu_ms = 0.f;
  if (h->neu_recorded && !h->overlap_used && cudaEventElapsedTime(&neu_ms, h->ev[12], h->ev[13]) == cudaSuccess) o.evaluate_neumannbcs = neu_ms;
  if (cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]) == cudaSuccess) o.evaluate_dirichletbcs = (ms > neu_ms) ? ms - neu_ms : 0.0;
  if (cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]) == cudaSuccess) o.evaluate_scatter = (ms > uni_ms) ? ms - uni_ms : 0.0;
  cudaGetLastError();
  *t = o;
  return TXASM_OK;
}

int txasm_last_fill_ms(txasm_handle h, double *out)
{
  TX_CHECK_H(h);
  if (!out) return TXASM_EINVAL;
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  if (!h->vol_recorded) return set_err(h, TXASM_ESTATE, "no fill recorded");
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev[5], h->ev[6]) != cudaSuccess) { cudaGetLastError(); return set_err(h, TXASM_ESTATE, "no fill recorded"); }
  if (h->overlap_used) {
    float ms2 = 0.f;
    if (cudaEventElapsedTime(&ms2, h->ev[10], h->ev[11]) == cudaSuccess) ms += ms2;
  }
  *out = ms;
  return TXASM_OK;
}

int txasm_fill_ms_history(txasm_handle h, double *ms, int cap, int *n)
{
  TX_CHECK_H(h);
  if (!ms || !n || cap < 0) return TXASM_EINVAL;
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  const long have = std::min<long>(h->fill_ring_count, h->fill_ring_n);
  const long take = std::min<long>(have, cap);
  for (long i = 0; i < take; ++i) {
    const long k = (h->fill_ring_count - take + i) % h->fill_ring_n;
    float a = 0.f, b = 0.f;
    if (cudaEventElapsedTime(&a, h->fill_ring[k * 4], h->fill_ring[k * 4 + 1]) != cudaSuccess) { cudaGetLastError(); return set_err(h, TXASM_ESTATE, "fill ring"); }
    if (h->fill_ring_segs[k] == 2 && cudaEventElapsedTime(&b, h->fill_ring[k * 4 + 2], h->fill_ring[k * 4 + 3]) != cudaSuccess) { cudaGetLastError(); b = 0.f; }
    ms[i] = (double)a + (double)b;
  }
  *n = (int)take;
  return TXASM_OK;
}

}  // extern "C"
