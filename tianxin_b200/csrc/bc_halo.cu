// bc_halo.cu -- Dirichlet rows and the owned<->ghosted halo exchange.
//
// Dirichlet: TianXin::DirichletEvalautor (disc-fe/src/evaluators/TianXin_Dirichlet_impl.hpp:59-81) ->
//   TpetraLinearObjContainer::applyDirichletBoundaryCondition / evalDirichletResidual
//   (disc-fe/src/lof/Panzer_TpetraLinearObjContainer.hpp:228-237, 306-317).
// Halo: replaces Tpetra Import(INSERT) / Export(ADD) of
//   TpetraLinearObjFactory::globalToGhostContainer / ghostToGlobalContainer
//   (disc-fe/src/lof/Panzer_TpetraLinearObjFactory_impl.hpp:124-219) with grouped ncclSend/ncclRecv
//   between neighbour ranks on the handle's stream.  NCCL is bound with dlopen so that the library
//   loads on machines without NCCL and shares the copy a host framework already loaded.
#include "txasm_internal.hpp"
#include "elem_q1hex.cuh"
#include <dlfcn.h>
#include <cstring>
#include <algorithm>

// minimal NCCL surface (ABI-stable since 2.x)
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64_ = 8 };

namespace txasm {

struct Nccl {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static Nccl *nccl_get(std::string *why)
{
  static Nccl n;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // share the already loaded copy
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    n.lib = lib;
    if (lib) {
      n.GetUniqueId = (decltype(n.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
      n.CommInitRank = (decltype(n.CommInitRank))dlsym(lib, "ncclCommInitRank");
      n.CommDestroy = (decltype(n.CommDestroy))dlsym(lib, "ncclCommDestroy");
      n.Send = (decltype(n.Send))dlsym(lib, "ncclSend");
      n.Recv = (decltype(n.Recv))dlsym(lib, "ncclRecv");
      n.AllReduce = (decltype(n.AllReduce))dlsym(lib, "ncclAllReduce");
      n.GroupStart = (decltype(n.GroupStart))dlsym(lib, "ncclGroupStart");
      n.GroupEnd = (decltype(n.GroupEnd))dlsym(lib, "ncclGroupEnd");
      n.GetErrorString = (decltype(n.GetErrorString))dlsym(lib, "ncclGetErrorString");
    }
  }
  if (!n.lib || !n.GetUniqueId || !n.CommInitRank || !n.Send || !n.Recv || !n.GroupStart || !n.GroupEnd) {
    if (why) *why = "libnccl.so.2 not found or incomplete";
    return nullptr;
  }
  return &n;
}

#define TX_NCCL(h, n, call)                                                                     \
  do {                                                                                          \
    ncclResult_t r__ = (call);                                                                  \
    if (r__ != 0) { h->sticky = true; return set_err(h, TXASM_ENCCL, "NCCL error %d (%s): %s", r__, \
                    (n)->GetErrorString ? (n)->GetErrorString(r__) : "?", #call); }             \
  } while (0)

void halo_free(txasm_handle h)
{
  if (!h->halo) return;
  Nccl *n = nccl_get(nullptr);
  p2p_free(h);
  if (h->halo->comm && n && n->CommDestroy) n->CommDestroy(h->halo->comm);
  delete h->halo;
  h->halo = nullptr;
}

// ------------------------------------------------------------------ kernels
// per Dirichlet row: CSR begin, length and the position of the diagonal (built once per dirichlet_set / graph, so the
// apply kernel has one independent 16-byte load per row instead of the chain dofs -> rowptr -> colind)
struct DirRow { long long beg; int len; int diag; };
__global__ void k_dirichlet_plan(int n, const int *__restrict__ dofs, const int64_t *__restrict__ rowptr,
                                 const int *__restrict__ colind, DirRow *__restrict__ plan)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = dofs[i];
  const int64_t b = rowptr[l], e = rowptr[l + 1];
  int diag = -1;
  for (int64_t k = b; k < e; ++k) if (colind[k] == l) diag = (int)(k - b);
  plan[i] = DirRow{(long long)b, (int)(e - b), diag};
}
__global__ void k_dirichlet(int n, const int *__restrict__ dofs, const double *__restrict__ vals, int jac,
                            const double *__restrict__ x, double *__restrict__ f,
                            const DirRow *__restrict__ plan, double *__restrict__ A)
{
  // one warp per Dirichlet row
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  if (jac && A) {
    const DirRow r = plan[w];
    for (int k = lane; k < r.len; k += 32) A[r.beg + k] = (k == r.diag) ? 1.0 : 0.0;
  }
  if (lane == 0 && f) { const int l = dofs[w]; f[l] = x[l] - vals[w]; }
}

// Jacobian evaluation without a residual (f == NULL, the eigenvalue path): the reference calls
// Tpetra::applyDirichletBoundaryConditionToLocalMatrixRowsAndColumns (lof/Panzer_TpetraLinearObjContainer.hpp:223-226),
// which also zeroes the COLUMNS of the Dirichlet DOFs.  One warp per Dirichlet row l: the graph is structurally
// symmetric (every cell couples all its DOFs), so the rows holding column l are the columns of row l.
__global__ void k_dirichlet_cols(int n, const int *__restrict__ dofs, const int64_t *__restrict__ rowptr,
                                 const int *__restrict__ colind, double *__restrict__ A)
{
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const int l = dofs[w];
  const int64_t b = rowptr[l], e = rowptr[l + 1];
  for (int64_t k = b + lane; k < e; k += 32) {
    const int r = colind[k];
    if (r == l) continue;
    int64_t lo = rowptr[r], hi = rowptr[r + 1] - 1;
    while (lo <= hi) {
      const int64_t mid = (lo + hi) >> 1;
      const int v = colind[mid];
      if (v == l) { A[mid] = 0.0; break; }
      if (v < l) lo = mid + 1; else hi = mid - 1;
    }
  }
}

// Neumann flux on element sides: one thread per side, 2x2 Gauss on the face, atomic adds into f (a node belongs to
// up to four listed sides; the reference scatters with atomic adds as well)
__global__ void k_neumann(int n, const int *__restrict__ cells, const int *__restrict__ sides, const double *__restrict__ vals,
                          const int *__restrict__ lids, const double *__restrict__ xyz, double *__restrict__ f)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int side_nodes[6][4] = {{0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {0, 4, 7, 3}, {0, 3, 2, 1}, {4, 5, 6, 7}};
  const int sd = sides[i];
  const int *l = lids + (int64_t)cells[i] * 8;
  double X[8][3];
  for (int a = 0; a < 8; ++a)
    for (int d = 0; d < 3; ++d) X[a][d] = xyz[(int64_t)l[a] * 3 + d];
  double V[4][3], t1[3], t2[3];
  for (int k = 0; k < 4; ++k) {
    const int v = side_nodes[sd][k];
    V[k][0] = hex_sx(v); V[k][1] = hex_sy(v); V[k][2] = hex_sz(v);
  }
  for (int d = 0; d < 3; ++d) { t1[d] = 0.5 * (V[1][d] - V[0][d]); t2[d] = 0.5 * (V[3][d] - V[0][d]); }
  double r[4] = {0.0, 0.0, 0.0, 0.0};
  for (int q = 0; q < 4; ++q) {
    const double s = (q == 1 || q == 2) ? TX_INV_SQRT3 : -TX_INV_SQRT3, t = (q >= 2) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
    double pt[3], J[3][3];
    for (int d = 0; d < 3; ++d)
      pt[d] = 0.25 * ((1 - s) * (1 - t) * V[0][d] + (1 + s) * (1 - t) * V[1][d] + (1 + s) * (1 + t) * V[2][d] + (1 - s) * (1 + t) * V[3][d]);
    for (int d = 0; d < 3; ++d)
      for (int e = 0; e < 3; ++e) J[d][e] = 0.0;
    for (int a = 0; a < 8; ++a) {
      const double fx = 1.0 + hex_sx(a) * pt[0], fy = 1.0 + hex_sy(a) * pt[1], fz = 1.0 + hex_sz(a) * pt[2];
      const double g[3] = {0.125 * hex_sx(a) * fy * fz, 0.125 * fx * hex_sy(a) * fz, 0.125 * fx * fy * hex_sz(a)};
      for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e) J[d][e] += X[a][d] * g[e];
    }
    double T1[3], T2[3];
    for (int d = 0; d < 3; ++d) {
      T1[d] = J[d][0] * t1[0] + J[d][1] * t1[1] + J[d][2] * t1[2];
      T2[d] = J[d][0] * t2[0] + J[d][1] * t2[1] + J[d][2] * t2[2];
    }
    const double nx = T1[1] * T2[2] - T1[2] * T2[1], ny = T1[2] * T2[0] - T1[0] * T2[2], nz = T1[0] * T2[1] - T1[1] * T2[0];
    const double wm = sqrt(nx * nx + ny * ny + nz * nz);
    for (int k = 0; k < 4; ++k) {          // the four vertices of the face; the others vanish on it
      const int v = side_nodes[sd][k];
      const double N = 0.125 * (1.0 + hex_sx(v) * pt[0]) * (1.0 + hex_sy(v) * pt[1]) * (1.0 + hex_sz(v) * pt[2]);
      r[k] += vals[i] * (N * wm);
    }
  }
  for (int k = 0; k < 4; ++k) atomicAdd(&f[l[side_nodes[sd][k]]], r[k]);
}

__global__ void k_cload(int n, const int *__restrict__ dofs, const double *__restrict__ vals, double *__restrict__ f)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&f[dofs[i]], vals[i]);      // a DOF may be listed twice
}

__global__ void k_pack(int64_t n, const int *__restrict__ idx, const double *__restrict__ v, double *__restrict__ buf)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) buf[i] = v[idx[i]];
}
__global__ void k_unpack_insert(int64_t n, const int *__restrict__ idx, const double *__restrict__ buf, double *__restrict__ v)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) v[idx[i]] = buf[i];
}
__global__ void k_pack64(int64_t n, const int64_t *__restrict__ idx, const double *__restrict__ v, double *__restrict__ buf)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) buf[i] = v[idx[i]];
}
__global__ void k_unpack_add_plan(int64_t n, const int64_t *__restrict__ dst, const int64_t *__restrict__ ptr,
                                  const int64_t *__restrict__ src, const double *__restrict__ buf, double *__restrict__ v)
{
  const int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (u >= n) return;
  const int64_t d = dst[u];
  double a = v[d];
  for (int64_t k = ptr[u]; k < ptr[u + 1]; ++k) a += buf[src[k]];     // neighbour order: reproducible
  v[d] = a;
}

static inline unsigned nblk(int64_t n) { return (unsigned)((n + 255) / 256); }

int unpack_plan_build(txasm_handle h, UnpackPlan &P, const std::vector<int64_t> &dst)
{
  std::vector<int64_t> order;
  order.reserve(dst.size());
  for (int64_t i = 0; i < (int64_t)dst.size(); ++i) if (dst[i] >= 0) order.push_back(i);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return dst[a] < dst[b]; });
  std::vector<int64_t> ud, ptr;
  for (int64_t k = 0; k < (int64_t)order.size(); ++k) {
    if (k == 0 || dst[order[k]] != dst[order[k - 1]]) { ud.push_back(dst[order[k]]); ptr.push_back(k); }
  }
  ptr.push_back((int64_t)order.size());
  P.n_dst = (int64_t)ud.size();
  int rc;
  if ((rc = dev_alloc(h, &P.d_dst, ud.size()))) return rc;
  if ((rc = dev_alloc(h, &P.d_ptr, ptr.size()))) return rc;
  if ((rc = dev_alloc(h, &P.d_src, order.size()))) return rc;
  if (!ud.empty()) TX_CUDA(h, copy_to_device_sync(h, P.d_dst, ud.data(), sizeof(int64_t) * ud.size()));
  TX_CUDA(h, copy_to_device_sync(h, P.d_ptr, ptr.data(), sizeof(int64_t) * ptr.size()));
  if (!order.empty()) TX_CUDA(h, copy_to_device_sync(h, P.d_src, order.data(), sizeof(int64_t) * order.size()));
  return TXASM_OK;
}

int launch_unpack_add(txasm_handle h, const UnpackPlan &P, const double *buf, double *v)
{
  if (P.n_dst == 0) return TXASM_OK;
  k_unpack_add_plan<<<nblk(P.n_dst), 256, 0, h->stream>>>(P.n_dst, P.d_dst, P.d_ptr, P.d_src, buf, v);
  h->launches += 1;
  return TXASM_OK;
}

int launch_dirichlet(txasm_handle h, int jac, const double *x, double *f, double *A)
{
  if (h->n_dir == 0) return TXASM_OK;
  if (f && !x) return set_err(h, TXASM_EINVAL, "Dirichlet residual needs x");
  if (jac && A && !h->d_dir_plan) {     // first Jacobian apply after dirichlet_set / a new graph
    int rc = dev_alloc(h, (DirRow **)&h->d_dir_plan, (size_t)h->n_dir);
    if (rc) return rc;
    k_dirichlet_plan<<<(h->n_dir + 127) / 128, 128, 0, h->stream>>>(h->n_dir, h->d_dir_dofs, h->d_rowptr, h->d_colind,
                                                                     (DirRow *)h->d_dir_plan);
  }
  const int threads = 128, warps_per_block = threads / 32;
  k_dirichlet<<<(h->n_dir + warps_per_block - 1) / warps_per_block, threads, 0, h->stream>>>(
      h->n_dir, h->d_dir_dofs, h->d_dir_vals, jac, x, f, (const DirRow *)h->d_dir_plan, A);
  h->launches += 1;
  if (jac && A && !f) {
    k_dirichlet_cols<<<(h->n_dir + warps_per_block - 1) / warps_per_block, threads, 0, h->stream>>>(h->n_dir, h->d_dir_dofs, h->d_rowptr,
                                                                                                      h->d_colind, A);
    h->launches += 1;
  }
  TX_CUDA(h, cudaGetLastError());
  return TXASM_OK;
}

int launch_neumann(txasm_handle h, double *f)
{
  if (h->n_neu == 0 || !f) return TXASM_OK;
  k_neumann<<<(h->n_neu + 127) / 128, 128, 0, h->stream>>>(h->n_neu, h->d_neu_cells, h->d_neu_sides, h->d_neu_vals, h->d_lids,
                                                          h->d_xyz, f);
  TX_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return TXASM_OK;
}

int launch_cload(txasm_handle h, double *f)
{
  if (h->n_cload == 0 || !f) return TXASM_OK;
  k_cload<<<(h->n_cload + 255) / 256, 256, 0, h->stream>>>(h->n_cload, h->d_cload_dofs, h->d_cload_vals, f);
  TX_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return TXASM_OK;
}

// ---------------------------------------------------------------- functional responses
struct GaussRule { int np; double x[16], w[16]; };
static int gauss_legendre(int n, GaussRule &g)
{
  if (n < 1 || n > 16) return -1;
  g.np = n;
  for (int i = 0; i < n; ++i) {
    double z = cos(3.14159265358979323846 * (i + 0.75) / (n + 0.5)), pp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 1; j <= n; ++j) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j; }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      const double dz = p1 / pp;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    g.x[n - 1 - i] = z;
    g.w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
  return 0;
}

// one thread per cell, tensor Gauss points in a loop; block tree reduction -> one partial per block
__global__ void __launch_bounds__(128) k_response(int64_t n_cells, const int *__restrict__ lids, const double *__restrict__ xyz,
                                                  const double *__restrict__ x, int kind, int solution_id, GaussRule g,
                                                  double *__restrict__ partial)
{
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double integral = 0.0;
  if (c < n_cells) {
    double X[8][3], u[8];
    for (int a = 0; a < 8; ++a) {
      const int64_t l = lids[c * 8 + a];
      u[a] = (kind == TXASM_RESP_IP_ARRAY) ? 0.0 : x[l];
      for (int d = 0; d < 3; ++d) X[a][d] = xyz[l * 3 + d];
    }
    const double twopi = 6.28318530717958647692;
    for (int k = 0; k < g.np; ++k)
      for (int j = 0; j < g.np; ++j)
        for (int i = 0; i < g.np; ++i) {
          const double pt[3] = {g.x[i], g.x[j], g.x[k]};
          double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, P[3] = {0, 0, 0}, A = 0.0, gr[3] = {0, 0, 0};
#pragma unroll
          for (int a = 0; a < 8; ++a) {
            const double fx = 1.0 + hex_sx(a) * pt[0], fy = 1.0 + hex_sy(a) * pt[1], fz = 1.0 + hex_sz(a) * pt[2];
            const double N = 0.125 * fx * fy * fz;
            const double dN[3] = {0.125 * hex_sx(a) * fy * fz, 0.125 * fx * hex_sy(a) * fz, 0.125 * fx * fy * hex_sz(a)};
            A += N * u[a];
            for (int d = 0; d < 3; ++d) {
              P[d] += N * X[a][d];
              gr[d] += dN[d] * u[a];
              for (int e = 0; e < 3; ++e) J[d][e] += X[a][d] * dN[e];
            }
          }
          const double c0 = J[1][1] * J[2][2] - J[2][1] * J[1][2];
          const double c1 = -J[1][0] * J[2][2] + J[2][0] * J[1][2];
          const double c2 = J[1][0] * J[2][1] - J[2][0] * J[1][1];
          const double det = J[0][0] * c0 + J[0][1] * c1 + J[0][2] * c2;
          const double wm = det * g.w[i] * g.w[j] * g.w[k];
          double sc;
          if (kind == TXASM_RESP_INTEGRAL) sc = A;
          else if (kind == TXASM_RESP_IP_ARRAY) sc = x[c * (int64_t)(g.np * g.np * g.np) + (i + g.np * (j + g.np * k))];
          else {
            double sx, cx, sy, cy, sz = 1.0, cz = 0.0;
            sincos(twopi * P[0], &sx, &cx); sincos(twopi * P[1], &sy, &cy);
            if (solution_id == TXASM_SOURCE_SIN3) sincos(twopi * P[2], &sz, &cz);
            const double B = sx * sy * sz;
            sc = (A - B) * (A - B);
            if (kind == TXASM_RESP_H1_ERROR) {
              const double id = 1.0 / det;
              double Ji[3][3];
              Ji[0][0] = c0 * id; Ji[1][0] = c1 * id; Ji[2][0] = c2 * id;
              Ji[0][1] = (-J[0][1] * J[2][2] + J[0][2] * J[2][1]) * id;
              Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
              Ji[2][1] = (-J[0][0] * J[2][1] + J[0][1] * J[2][0]) * id;
              Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
              Ji[1][2] = (-J[0][0] * J[1][2] + J[0][2] * J[1][0]) * id;
              Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
              const double gB[3] = {twopi * cx * sy * sz, twopi * sx * cy * sz, twopi * sx * sy * cz};
              for (int d = 0; d < 3; ++d) {
                const double gA = Ji[0][d] * gr[0] + Ji[1][d] * gr[1] + Ji[2][d] * gr[2];
                sc += (gA - gB[d]) * (gA - gB[d]);
              }
            }
          }
          integral += sc * wm;
        }
  }
  __shared__ double red[128];
  red[threadIdx.x] = integral;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
// fixed-order sum of the per-block partials (one block): bitwise reproducible
__global__ void __launch_bounds__(1024) k_sum_partials(int64_t n, const double *__restrict__ partial, double *__restrict__ out)
{
  __shared__ double red[1024];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

int response_functional(txasm_handle h, int kind, int solution_id, int cub_degree, const double *x_dev, double *value_host)
{
  GaussRule g;
  if (cub_degree < 0 || gauss_legendre(cub_degree / 2 + 1, g)) return set_err(h, TXASM_EINVAL, "response: cubature degree %d", cub_degree);
  const int64_t nb = (h->n_cells + 127) / 128;
  double *d_part = nullptr;
  TX_CUDA(h, cudaMalloc(&d_part, sizeof(double) * (size_t)(nb + 1)));
  if (nb) k_response<<<(unsigned)nb, 128, 0, h->stream>>>(h->n_cells, h->d_lids, h->d_xyz, x_dev, kind, solution_id, g, d_part);
  k_sum_partials<<<1, 1024, 0, h->stream>>>(nb, d_part, d_part + nb);
  TX_CUDA(h, cudaGetLastError());
  h->launches += 2;
  int rc = halo_allreduce_sum(h, d_part + nb);
  if (rc) { cudaFree(d_part); return rc; }
  TX_CUDA(h, cudaMemcpyAsync(value_host, d_part + nb, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_part);
  return TXASM_OK;
}

// does the export read or write a row of the uniform tile range?  (it never should: ghost rows are not interior
// rows and the owned rows it adds into carry remote columns -- checked once, the overlapped schedule depends on it)
int halo_rows_touch_uniform_tiles(txasm_handle h, bool *touch)
{
  *touch = false;
  Halo *H = h->halo;
  if (!H || H->n_nbr == 0) return TXASM_OK;
  bool t1 = false, t2 = false;
  int rc = rows_touch_uniform_tiles(h, H->d_send_lids, H->send_off[H->n_nbr], &t1);
  if (rc) return rc;
  rc = rows_touch_uniform_tiles(h, H->d_recv_lids, H->recv_off[H->n_nbr], &t2);
  if (rc) return rc;
  *touch = t1 || t2;
  return TXASM_OK;
}

int halo_n_neighbours(txasm_handle h) { return h->halo ? h->halo->n_nbr : 0; }
int64_t halo_n_owned(txasm_handle h) { return h->halo ? h->halo->n_owned : h->n_rows; }

int halo_allreduce_sum(txasm_handle h, double *d_value)
{
  Halo *H = h->halo;
  if (!H || !H->comm || H->nranks <= 1) return TXASM_OK;
  Nccl *n = nccl_get(nullptr);
  if (!n->AllReduce) return set_err(h, TXASM_ENCCL, "ncclAllReduce not found");
  TX_NCCL(h, n, n->AllReduce(d_value, d_value, 1, ncclFloat64_, 0 /*ncclSum*/, H->comm, h->stream));
  return TXASM_OK;
}

// x ghosts := owner values.  One grouped send/recv per vector.
int halo_import(txasm_handle h, double *const x[3])
{
  Halo *H = h->halo;
  if (!H || H->n_nbr == 0) return TXASM_OK;
  if (p2p_active(h)) return p2p_import(h, x);
  Nccl *n = nccl_get(nullptr);
  const int64_t ns = H->send_off[H->n_nbr], nr = H->recv_off[H->n_nbr];
  for (int v = 0; v < 3; ++v) {
    if (!x[v]) continue;
    if (ns) k_pack<<<nblk(ns), 256, 0, h->stream>>>(ns, H->d_send_lids, x[v], H->d_sbuf);
    TX_NCCL(h, n, n->GroupStart());
    for (int k = 0; k < H->n_nbr; ++k) {
      const int64_t cs = H->send_off[k + 1] - H->send_off[k], cr = H->recv_off[k + 1] - H->recv_off[k];
      if (cs) TX_NCCL(h, n, n->Send(H->d_sbuf + H->send_off[k], (size_t)cs, ncclFloat64_, H->nbr[k], H->comm, h->stream));
      if (cr) TX_NCCL(h, n, n->Recv(H->d_rbuf + H->recv_off[k], (size_t)cr, ncclFloat64_, H->nbr[k], H->comm, h->stream));
    }
    TX_NCCL(h, n, n->GroupEnd());
    if (nr) k_unpack_insert<<<nblk(nr), 256, 0, h->stream>>>(nr, H->d_recv_lids, H->d_rbuf, x[v]);
    h->launches += 2;
  }
  TX_CUDA(h, cudaGetLastError());
  return TXASM_OK;
}

// f: ghost entries travel to their owner and are added (Export ADD); A: whole ghost rows.  Both exchanges share
// one grouped send/recv (one NCCL launch); the adds run neighbour by neighbour so the result is deterministic.
int halo_export(txasm_handle h, double *f, double *A, int jac)
{
  Halo *H = h->halo;
  if (!H || H->n_nbr == 0) return TXASM_OK;
  Nccl *n = nccl_get(nullptr);
  const bool do_f = f != nullptr, do_A = jac && A && H->have_mat;
  if (!do_f && !do_A) return TXASM_OK;
  if (p2p_active(h)) return p2p_export(h, f, A, jac);
  const int64_t ns = H->recv_off[H->n_nbr];   // I send my ghost entries ...
  const int64_t ms = do_A ? H->msend_off[H->n_nbr] : 0;
  if (do_f && ns) k_pack<<<nblk(ns), 256, 0, h->stream>>>(ns, H->d_recv_lids, f, H->d_rbuf);
  if (do_A && ms) k_pack64<<<nblk(ms), 256, 0, h->stream>>>(ms, H->d_msend_src, A, H->d_msbuf);
  TX_NCCL(h, n, n->GroupStart());
  for (int k = 0; k < H->n_nbr; ++k) {
    if (do_f) {
      const int64_t cs = H->recv_off[k + 1] - H->recv_off[k], cr = H->send_off[k + 1] - H->send_off[k];
      if (cs) TX_NCCL(h, n, n->Send(H->d_rbuf + H->recv_off[k], (size_t)cs, ncclFloat64_, H->nbr[k], H->comm, h->stream));
      if (cr) TX_NCCL(h, n, n->Recv(H->d_sbuf + H->send_off[k], (size_t)cr, ncclFloat64_, H->nbr[k], H->comm, h->stream));
    }
    if (do_A) {
      const int64_t cs = H->msend_off[k + 1] - H->msend_off[k], cr = H->mrecv_off[k + 1] - H->mrecv_off[k];
      if (cs) TX_NCCL(h, n, n->Send(H->d_msbuf + H->msend_off[k], (size_t)cs, ncclFloat64_, H->nbr[k], H->comm, h->stream));
      if (cr) TX_NCCL(h, n, n->Recv(H->d_mrbuf + H->mrecv_off[k], (size_t)cr, ncclFloat64_, H->nbr[k], H->comm, h->stream));
    }
  }
  TX_NCCL(h, n, n->GroupEnd());
  // ... and add what the neighbours sent: one launch per container, contributions in neighbour order
  if (do_f) { int rc = launch_unpack_add(h, H->up_f, H->d_sbuf, f); if (rc) return rc; }
  if (do_A) { int rc = launch_unpack_add(h, H->up_A, H->d_mrbuf, A); if (rc) return rc; }
  h->launches += (do_f ? 1 : 0) + (do_A ? 1 : 0);     // the pack kernels
  TX_CUDA(h, cudaGetLastError());
  return TXASM_OK;
}

}  // namespace txasm

using namespace txasm;

extern "C" {

int txasm_comm_unique_id(void *id128)
{
  if (!id128) return TXASM_EINVAL;
  std::string why;
  Nccl *n = nccl_get(&why);
  if (!n) return set_err(nullptr, TXASM_ENCCL, "%s", why.c_str());
  ncclUniqueId id;
  if (n->GetUniqueId(&id) != 0) return set_err(nullptr, TXASM_ENCCL, "ncclGetUniqueId failed");
  memcpy(id128, &id, sizeof(id));
  return TXASM_OK;
}

int txasm_comm_init(txasm_handle h, int nranks, int rank, const void *id128)
{
  if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return TXASM_EINVAL;
  cudaSetDevice(h->device);
  std::string why;
  Nccl *n = nccl_get(&why);
  if (!n) return set_err(h, TXASM_ENCCL, "%s", why.c_str());
  if (!h->halo) h->halo = new Halo();
  h->halo->nranks = nranks; h->halo->rank = rank;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  TX_NCCL(h, n, n->CommInitRank(&h->halo->comm, nranks, id, rank));
  return TXASM_OK;
}

int txasm_halo_set(txasm_handle h, int64_t n_owned, int n_nbr, const int *nbr_rank, const int64_t *send_off,
                   const int *send_lids, const int64_t *recv_off, const int *recv_lids)
{
  if (!h) return TXASM_EINVAL;
  cudaSetDevice(h->device);
  if (!h->halo) h->halo = new Halo();
  Halo *H = h->halo;
  if (n_nbr > 0 && !H->comm) return set_err(h, TXASM_ESTATE, "halo_set with neighbours needs txasm_comm_init first");
  if (n_nbr < 0 || (n_nbr && (!nbr_rank || !send_off || !recv_off))) return set_err(h, TXASM_EINVAL, "halo_set: bad arguments");
  H->n_owned = n_owned; H->n_nbr = n_nbr;
  h->overlap_state = 0;
  H->nbr.assign(nbr_rank, nbr_rank + n_nbr);
  H->send_off.assign(send_off, send_off + n_nbr + 1);
  H->recv_off.assign(recv_off, recv_off + n_nbr + 1);
  const int64_t ns = n_nbr ? send_off[n_nbr] : 0, nr = n_nbr ? recv_off[n_nbr] : 0;
  int rc;
  if ((rc = dev_alloc(h, &H->d_send_lids, (size_t)ns))) return rc;
  if ((rc = dev_alloc(h, &H->d_recv_lids, (size_t)nr))) return rc;
  if ((rc = dev_alloc(h, &H->d_sbuf, (size_t)ns))) return rc;
  if ((rc = dev_alloc(h, &H->d_rbuf, (size_t)nr))) return rc;
  if (ns) TX_CUDA(h, copy_to_device_sync(h, H->d_send_lids, send_lids, sizeof(int) * ns));
  if (nr) TX_CUDA(h, copy_to_device_sync(h, H->d_recv_lids, recv_lids, sizeof(int) * nr));
  {
    std::vector<int> sl((size_t)ns);
    if (ns) TX_CUDA(h, copy_to_device_sync(h, sl.data(), H->d_send_lids, sizeof(int) * ns));   // (send_lids may be a device array)
    std::vector<int64_t> dst(sl.begin(), sl.end());
    if ((rc = unpack_plan_build(h, H->up_f, dst))) return rc;
  }
  return TXASM_OK;
}

int txasm_halo_set_matrix(txasm_handle h, const int64_t *mat_recv_off, const int64_t *mat_recv_pos)
{
  if (!h || !h->halo) return TXASM_EINVAL;
  cudaSetDevice(h->device);
  Halo *H = h->halo;
  if (!h->have_graph) return set_err(h, TXASM_ESTATE, "halo_set_matrix needs the graph");
  if (H->n_nbr == 0) { H->have_mat = true; H->msend_off.assign(1, 0); H->mrecv_off.assign(1, 0); return TXASM_OK; }
  // send side: every ghost row (recv_lids order), whole row
  const int64_t nr = H->recv_off[H->n_nbr];
  std::vector<int> rl((size_t)nr);
  std::vector<int64_t> rp((size_t)h->n_rows + 1);
  TX_CUDA(h, copy_to_device_sync(h, rl.data(), H->d_recv_lids, sizeof(int) * nr));
  TX_CUDA(h, copy_to_device_sync(h, rp.data(), h->d_rowptr, sizeof(int64_t) * (h->n_rows + 1)));
  H->msend_off.assign(H->n_nbr + 1, 0);
  std::vector<int64_t> src;
  for (int k = 0; k < H->n_nbr; ++k) {
    for (int64_t i = H->recv_off[k]; i < H->recv_off[k + 1]; ++i)
      for (int64_t p = rp[rl[i]]; p < rp[rl[i] + 1]; ++p) src.push_back(p);
    H->msend_off[k + 1] = (int64_t)src.size();
  }
  H->mrecv_off.assign(mat_recv_off, mat_recv_off + H->n_nbr + 1);
  const int64_t ms = (int64_t)src.size(), mr = H->mrecv_off[H->n_nbr];
  int rc;
  if ((rc = dev_alloc(h, &H->d_msend_src, (size_t)ms))) return rc;
  if ((rc = dev_alloc(h, &H->d_mrecv_pos, (size_t)mr))) return rc;
  if ((rc = dev_alloc(h, &H->d_msbuf, (size_t)ms))) return rc;
  if ((rc = dev_alloc(h, &H->d_mrbuf, (size_t)mr))) return rc;
  if (ms) TX_CUDA(h, copy_to_device_sync(h, H->d_msend_src, src.data(), sizeof(int64_t) * ms));
  if (mr) TX_CUDA(h, copy_to_device_sync(h, H->d_mrecv_pos, mat_recv_pos, sizeof(int64_t) * mr));
  {
    std::vector<int64_t> dst((size_t)mr);
    if (mr) TX_CUDA(h, copy_to_device_sync(h, dst.data(), H->d_mrecv_pos, sizeof(int64_t) * mr));
    if ((rc = unpack_plan_build(h, H->up_A, dst))) return rc;
  }
  H->have_mat = true;
  return TXASM_OK;
}

}  // extern "C"
