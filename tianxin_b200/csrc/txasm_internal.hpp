// txasm_internal.hpp -- handle layout and helpers shared by the .cu files of libtxasm.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <cstdio>
#include "../../include/txasm.h"

namespace txasm {

constexpr int NB = 8;        // Q1 hex: basis functions = vertices
constexpr int NQ = 8;        // 2x2x2 Gauss
constexpr int MAX_SRC = 4;   // source terms per block
constexpr int MAX_ADJ = 32;
constexpr int ELEM_REC = 44;  // doubles per element record of the general-hexahedron path: K upper triangle [36] (sym_idx) | r[8]  // elements around a node handled by the row paths

// Consolidated integrand coefficients for one evaluate (see DESIGN.md "terms"):
//   residual  r = K (sum_v kg[v] u_v) + M (sum_v km[v] u_v) + sum_s src_mult[s] int(phi s_s)
//   jacobian  J = cK K + cM M,  cK = sum_v kg[v] seed[v],  cM = sum_v km[v] seed[v]
// which is what the Fad chain gather(seed) -> DOFGradient/DOF -> Integrator_* -> scatter yields
// for these (linear) integrands.
struct FillCoef {
  double kg[3];
  double km[3];
  double cK, cM;
  int    n_src;
  int    src_id[MAX_SRC];
  double src_mult[MAX_SRC];
  const double *src_ip[MAX_SRC];
  int    has_mass;   // any km != 0
  int    has_vec[3]; // vector v is read at all
  // Integrator field multipliers (the "Field Multipliers" of Integrator_GradBasisDotVector / _BasisTimesScalar,
  // Panzer_Integrator_GradBasisDotVector_impl.hpp:257-294): their product at the integration points, [n_cells][8];
  // NULL = 1.  One array for the GRADGRAD terms, one for the MASS terms.  Cells then take the general 2x2x2 path.
  const double *fmK, *fmM;
};

struct FillArgs {
  // mesh / dofs
  int64_t n_cells;
  int64_t n_rows;
  const int *lids;          // [n_cells][8]
  const double *xyz;        // [n_rows][3] node coordinates by LID
  // graph
  const int64_t *rowptr;
  const int *colind;
  // state
  const double *x[3];       // x, xdot, xdotdot (ghosted, by LID); may be null when unused
  double *f;
  double *A;
  int jacobian;             // fill A
  FillCoef c;
  const double *elem;       // general hexahedra: element records (ELEM_REC doubles) computed by k_elem_general
  unsigned long long *dbg;  // profiling (TXASM_TIMELINE=1): [2k] = earliest start, [2k+1] = latest end (globaltimer ns) of fill kernel k
};

struct Tiles;
struct Halo;
struct GBlocks;

}  // namespace txasm

struct txasm_handle_s {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  bool sticky = false;
  txasm_config cfg{};
  int n_sm = 0;
  int smem_optin = 0;

  // block (one block in this version)
  bool have_block = false;
  int64_t n_cells = 0, n_rows = 0, nnz = 0;
  const int *d_lids = nullptr;
  double *d_xyz = nullptr;              // owned: node coordinates by LID
  // graph
  bool have_graph = false;
  const int64_t *d_rowptr = nullptr;
  const int *d_colind = nullptr;
  // row -> (element, local node) adjacency, CSR, entries packed e*8+a, sorted
  int64_t *d_adj_ptr = nullptr;
  int *d_adj = nullptr;
  int max_adj = 0;
  // classification
  unsigned char *d_cell_affine = nullptr;
  int64_t n_affine = 0;
  // terms
  std::vector<txasm_term> terms;
  std::vector<const double *> d_src_ip;  // device copies of ip arrays
  std::vector<const double *> d_field_mult;   // device copies of the terms' field multipliers
  bool force_general = false;            // a term carries field multipliers: no cell is treated as affine
  // dirichlet
  int n_dir = 0;
  int *d_dir_dofs = nullptr;
  double *d_dir_vals = nullptr;
  void *d_dir_plan = nullptr;        // per Dirichlet row: CSR begin / length / diagonal position (bc_halo.cu)
  int n_neu = 0;
  int *d_neu_cells = nullptr, *d_neu_sides = nullptr;
  double *d_neu_vals = nullptr;
  int n_cload = 0;
  int *d_cload_dofs = nullptr;
  double *d_cload_vals = nullptr;
  unsigned long long *d_dbg = nullptr;  // kernel timeline slots (TXASM_TIMELINE=1)
  double *d_elem = nullptr;             // [n_cells][ELEM_REC] element records of the general-hexahedron path (lazily allocated)
  // row-tile path (filled by setup)
  txasm::Tiles *tiles = nullptr;
  int mode = 0;                         // scatter mode selected at setup
  bool is_setup = false;
  // host staging for evaluate with host arrays
  double *st_x[3] = {nullptr, nullptr, nullptr};
  double *st_f = nullptr, *st_A = nullptr;
  // timing
  cudaEvent_t ev[14] = {};          // 0-4 stage boundaries, 5/6 fill (first part), 8/9 fork/join of the export, 10/11 fill (uniform part), 12/13 Neumann
  cudaStream_t side_stream = nullptr; // the export runs here under the uniform-tile kernel (see txasm_evaluate)
  int overlap_state = 0;              // 0 unknown, 1 export may overlap the uniform tiles, 2 it may not
  bool overlap_used = false;          // last evaluate used the overlapped schedule
  int launches = 0;
  // run-time switches (txasm_option_set; defaults from the environment at creation)
  int opt_uniform = 1, opt_brick = 1, opt_overlap = 0, opt_fuse_dir = 1, opt_concurrent = 1;
  int opt_block_atomic = 1;           // general blocks: 1 = planned atomic adds (ScatterResidual semantics; the faster of the two as measured), 0 = owner-computes gather (no atomics, reproducible)
  int opt_dmma = 1;                   // Q2 hexahedra: element matrix on the FP64 tensor cores (k_gblock_q2_dmma)
  int opt_p2p = 1;                    // 1: the halo goes over peer memory once txasm_halo_p2p_connect has run, 0: NCCL send/recv
  std::vector<cudaEvent_t> fill_ring;   // option fill_event_ring: 4 events per evaluate (begin/end of up to two fill segments)
  std::vector<int> fill_ring_segs;      // segments recorded in the slot
  int fill_ring_n = 0;
  long fill_ring_count = 0;
  int brick_ctas_limit = 0;           // set per evaluate: CTAs per SM left to k_fill_brick when the export runs beside it
  bool timers_recorded = false;       // the last evaluate recorded its stage events
  int opt_timers = 0;                 // 1: CUDA events around the stages of every evaluate (txasm_timers_get / txasm_last_fill_ms); costs ~2 % of a 256^3 step
  int opt_edge = 1;                   // 1: lattice tiles with rows on their faces go to k_fill_edge (0: to k_fill_rowtile)
  int opt_rest_ctas = 0;              // > 0: CTAs per SM of the boundary-tile kernel (tuning: co-residency with k_fill_brick)
  int opt_brick_ctas = 0;             // > 0: CTAs per SM of k_fill_brick (tuning)
  int opt_grid_cap = 0;               // > 0: persistent kernels launch at most this many CTAs (tests: many tiles per CTA on small meshes)
  int uniform_used = 0;               // last evaluate: 0 none, 1 k_fill_uniform, 2 k_fill_brick
  bool dir_fused = false;             // last evaluate: Dirichlet rows written by the fill kernel
  int dir_fusable = -1;               // -1 unknown, 0 no (a Dirichlet row lies outside the general tiles), 1 yes
  int *d_row_dir = nullptr;           // [n_rows] index into the Dirichlet arrays or -1 (fused Dirichlet)
  double setup_ms = 0.0;
  bool neu_recorded = false;          // events 12/13 of the last evaluate bracket the Neumann side sets
  bool vol_recorded = false;          // events 5/6 of the last evaluate bracket a fill
  // owned allocations
  std::vector<void *> owned;
  // halo / nccl
  txasm::Halo *halo = nullptr;
  txasm::GBlocks *gblocks = nullptr;    // general element blocks (gblock.cu); exclusive with the Q1 fast-path block
};

typedef struct ncclComm *ncclComm_t;

namespace txasm {

// CSR by destination of an additive unpack: destination u gets the sum of buf[src[ptr[u] .. ptr[u+1])] in that order
// (neighbour order), so the Export ADD is one launch and bitwise reproducible
struct UnpackPlan {
  int64_t n_dst = 0;
  int64_t *d_dst = nullptr;      // destination index (into f or A)
  int64_t *d_ptr = nullptr;      // [n_dst + 1]
  int64_t *d_src = nullptr;      // positions in the receive buffer
};
struct P2P;

struct Halo {
  int nranks = 1, rank = 0;
  ncclComm_t comm = nullptr;
  int64_t n_owned = 0;
  int n_nbr = 0;
  std::vector<int> nbr;
  std::vector<int64_t> send_off, recv_off;         // vector halo
  int *d_send_lids = nullptr, *d_recv_lids = nullptr;
  double *d_sbuf = nullptr, *d_rbuf = nullptr;     // max(send,recv) sized, reused for x / f
  // matrix export
  bool have_mat = false;
  std::vector<int64_t> msend_off, mrecv_off;
  int64_t *d_msend_src = nullptr;                  // index into A of every value I send
  int64_t *d_mrecv_pos = nullptr;                  // destination index into A (or -1) of every value I receive
  double *d_msbuf = nullptr, *d_mrbuf = nullptr;
  UnpackPlan up_f, up_A;                           // fused, ordered unpack of the export
  P2P *p2p = nullptr;                              // peer-memory exchange (halo_p2p.cu); NULL: NCCL send/recv
};

int set_err(txasm_handle h, int code, const char *fmt, ...);
int cuda_fail(txasm_handle h, cudaError_t e, const char *what, const char *file, int line);

#define TX_CUDA(h, call)                                                          \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) return txasm::cuda_fail(h, e__, #call, __FILE__, __LINE__); \
  } while (0)

// true if p is a device-accessible (device or managed) pointer
bool is_device_ptr(const void *p);

template <class T>
int dev_alloc(txasm_handle h, T **out, size_t n)
{
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, (n ? n : 1) * sizeof(T));
  if (e != cudaSuccess) return cuda_fail(h, e, "cudaMalloc", __FILE__, __LINE__);
  h->owned.push_back(p);
  *out = (T *)p;
  return TXASM_OK;
}
void dev_free(txasm_handle h, void *p);

// device view of a caller array: borrowed if device pointer, else owned copy
template <class T>
int to_device(txasm_handle h, const T *src, size_t n, const T **out)
{
  if (!src) { *out = nullptr; return TXASM_OK; }
  if (is_device_ptr(src)) { *out = src; return TXASM_OK; }
  T *d = nullptr;
  int rc = dev_alloc(h, &d, n);
  if (rc) return rc;
  cudaError_t e = cudaMemcpyAsync(d, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream);
  if (e != cudaSuccess) return cuda_fail(h, e, "cudaMemcpyAsync(H2D)", __FILE__, __LINE__);
  e = cudaStreamSynchronize(h->stream);   // src may be pageable / freed by the caller
  if (e != cudaSuccess) return cuda_fail(h, e, "cudaStreamSynchronize", __FILE__, __LINE__);
  *out = d;
  return TXASM_OK;
}

// ---- setup kernels (setup_kernels.cu)
int build_node_coords(txasm_handle h, const double *d_cell_coords);
int build_adjacency(txasm_handle h);
int build_graph_device(txasm_handle h, int64_t *nnz_out);
int classify_cells(txasm_handle h);

// ---- fill paths
int launch_fill_atomic(txasm_handle h, const FillArgs &a);        // fill_atomic.cu
int launch_fill_rowgather(txasm_handle h, const FillArgs &a);     // fill_rowgather.cu
int launch_elem_general(txasm_handle h, FillArgs &a, cudaStream_t st);   // fill_general.cu: fills a.elem
int tiles_build(txasm_handle h);                                  // fill_rowtile.cu
void tiles_free(txasm_handle h);
enum { FILL_ALL = 0, FILL_REST = 1, FILL_UNIFORM = 2 };   // all tiles | everything but the uniform range | the uniform range
int launch_fill_rowtile(txasm_handle h, const FillArgs &a, int part, cudaStream_t st, bool fuse_dir);
bool fill_uniform_eligible(txasm_handle h, const FillArgs &a);
void fill_ranges(txasm_handle h, const FillArgs &a, int *e_brick, int *e_uni, int *e_edge = nullptr);   // tiles [0,e_brick) brick, [e_brick,e_uni) uniform kernel
int dirichlet_fuse_prepare(txasm_handle h);      // builds d_row_dir, sets dir_fusable
int rows_touch_uniform_tiles(txasm_handle h, const int *d_rows, int64_t n, bool *touch);
int halo_rows_touch_uniform_tiles(txasm_handle h, bool *touch);
int halo_n_neighbours(txasm_handle h);
int64_t halo_n_owned(txasm_handle h);
int tiles_count(txasm_handle h);
int tiles_info(txasm_handle h, txasm_info *info);
int tiles_get(txasm_handle h, int tile, int *rows, int *cells, unsigned short *adjl, int *n_cells_out);

// ---- boundary / halo (bc_halo.cu)
// Copy in either direction, ordered on the handle's stream, complete on return.  Plain cudaMemcpy from pageable
// memory may return before the data has landed and only orders with BLOCKING streams; the handle's stream can be a
// caller's non-blocking one (torch), whose kernels would then read the destination too early.
inline cudaError_t copy_to_device_sync(txasm_handle h, void *dst, const void *src, size_t bytes)
{
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, h->stream);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(h->stream);
}
int launch_dirichlet(txasm_handle h, int jacobian, const double *x, double *f, double *A);
int launch_cload(txasm_handle h, double *f);
int launch_neumann(txasm_handle h, double *f);
int response_functional(txasm_handle h, int kind, int solution_id, int cub_degree, const double *x_dev, double *value_host);
int halo_allreduce_sum(txasm_handle h, double *d_value);   // no-op without a communicator
void halo_free(txasm_handle h);
int halo_import(txasm_handle h, double *const x[3]);
int halo_export(txasm_handle h, double *f, double *A, int jacobian);
int unpack_plan_build(txasm_handle h, UnpackPlan &P, const std::vector<int64_t> &dst /* -1: skip */);
int launch_unpack_add(txasm_handle h, const UnpackPlan &P, const double *buf, double *v);
// gblock.cu
void gblocks_free(txasm_handle h);
int gblocks_count(txasm_handle h);
int gblocks_setup(txasm_handle h);
int gblocks_graph_build(txasm_handle h, int64_t *nnz_out);
int launch_gblocks(txasm_handle h, int jacobian, const txasm_inargs *in, const double *const x[3], double *f, double *A);
// halo_p2p.cu
void p2p_free(txasm_handle h);
bool p2p_active(txasm_handle h);
int p2p_import(txasm_handle h, double *const x[3]);
int p2p_export(txasm_handle h, double *f, double *A, int jac);

}  // namespace txasm
