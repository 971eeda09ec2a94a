// halo_p2p.cu -- the owned <-> ghosted exchange over NVLink peer memory, without NCCL in the loop.
//
// Replaces Tpetra Import(INSERT) / Export(ADD) of TpetraLinearObjFactory::globalToGhostContainer /
// ghostToGlobalContainer (disc-fe/src/lof/Panzer_TpetraLinearObjFactory_impl.hpp:124-219) like bc_halo.cu does, but as
// the device-side half of the assembly step instead of a library collective:
//
//   * every rank owns one receive slab (cudaMalloc) and publishes it with cudaIpcGetMemHandle; neighbours map it with
//     cudaIpcOpenMemHandle (one process per GPU, NVSwitch: every peer at full NVLink bandwidth);
//   * PUSH: one kernel gathers what the neighbours need (x of my owned DOFs / f and the whole A rows of my ghost DOFs)
//     and stores it straight into their slabs with plain st.global over NVLink; the last CTA to finish fences
//     (__threadfence_system) and raises one flag per neighbour (an epoch counter, so nothing is ever reset);
//   * WAIT: one warp polls my flags with ld.acquire.sys until every neighbour has delivered this epoch (bounded: a
//     neighbour that never arrives sets a sticky error instead of hanging the GPU);
//   * UNPACK: x ghosts are inserted; f / A contributions are added by destination in neighbour order (UnpackPlan:
//     one launch, bitwise reproducible).
//
// All kernels are small (256 threads, < 64 registers, no shared memory), so they fit beside the persistent fill
// kernels: txasm_evaluate runs the export on a side stream under k_fill_brick.  Buffer reuse needs no double
// buffering: a rank can only reach push(e+1) after its wait(e+1 import) resp. wait(e export) saw every neighbour's
// flag, which the neighbour raises after its own unpack of the previous epoch in stream order.
#include "txasm_internal.hpp"
#include <cstring>
#include <vector>

namespace txasm {

constexpr int P2P_MAX_NBR = 32;
struct P2PPeerDev {                 // where neighbour k receives from me (pointers into ITS slab, mapped here)
  double *x, *f, *A;
  unsigned long long *flag_import, *flag_export;
};
struct P2PSeg { int64_t send_off[P2P_MAX_NBR + 1], recv_off[P2P_MAX_NBR + 1], msend_off[P2P_MAX_NBR + 1]; };

struct P2PBlobEntry { int nbr; int flag_slot; int64_t x_off, f_off, A_off; };     // byte offsets into the owner's slab
struct P2PBlob {
  cudaIpcMemHandle_t handle;
  int rank, n_nbr;
  int64_t flags_off;                // byte offset of flags[2][P2P_MAX_NBR]
  P2PBlobEntry e[P2P_MAX_NBR];
};

struct P2P {
  bool connected = false;
  unsigned char *slab = nullptr;
  size_t slab_bytes = 0;
  double *recv_x = nullptr, *recv_f = nullptr, *recv_A = nullptr;
  unsigned long long *flags = nullptr;        // [2][P2P_MAX_NBR]: import, export
  std::vector<void *> opened;
  P2PPeerDev *d_peers = nullptr;
  P2PSeg seg{};
  P2PSeg *d_seg = nullptr;
  unsigned long long epoch_import = 0, epoch_export = 0;
  unsigned int *d_done = nullptr;             // CTA completion counters of the push kernels
  int *d_timeout = nullptr;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// the last CTA of a push raises the flags
__device__ __forceinline__ void p2p_signal(unsigned int *done, unsigned int n_ctas, int n_nbr, const P2PPeerDev *peers, int which,
                                           unsigned long long epoch)
{
  __threadfence_system();                     // my remote stores before my arrival
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(done, 1u) == n_ctas - 1);
  __syncthreads();
  if (last) {
    __threadfence_system();
    if ((int)threadIdx.x < n_nbr) st_release_sys(which ? peers[threadIdx.x].flag_export : peers[threadIdx.x].flag_import, epoch);
    if (threadIdx.x == 0) *done = 0;          // ready for the next push (stream order)
  }
}

// import: x of my owned DOFs on neighbour k's send list -> its recv_x segment
__global__ void __launch_bounds__(256) k_p2p_push_x(int n_nbr, const P2PSeg *__restrict__ seg, const P2PPeerDev *__restrict__ peers,
                                                    const int *__restrict__ send_lids, const double *__restrict__ x0,
                                                    const double *__restrict__ x1, const double *__restrict__ x2, int64_t stride,
                                                    unsigned int *done, unsigned long long epoch)
{
  const int64_t ns = seg->send_off[n_nbr];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ns; i += (int64_t)gridDim.x * blockDim.x) {
    int k = 0;
    while (i >= seg->send_off[k + 1]) ++k;
    const int64_t o = i - seg->send_off[k];
    const int l = send_lids[i];
    double *dst = peers[k].x + o;
    const int64_t cnt = seg->send_off[k + 1] - seg->send_off[k];
    if (x0) dst[0] = x0[l];
    if (x1) dst[cnt] = x1[l];                 // the vectors of one neighbour lie one after the other
    if (x2) dst[2 * cnt] = x2[l];
  }
  (void)stride;
  p2p_signal(done, gridDim.x, n_nbr, peers, 0, epoch);
}

// export: f of my ghost DOFs and the values of my ghost rows -> the owners' recv_f / recv_A segments
__global__ void __launch_bounds__(256) k_p2p_push_export(int n_nbr, const P2PSeg *__restrict__ seg, const P2PPeerDev *__restrict__ peers,
                                                         const int *__restrict__ recv_lids, const int64_t *__restrict__ msend_src,
                                                         const double *__restrict__ f, const double *__restrict__ A,
                                                         unsigned int *done, unsigned long long epoch)
{
  const int64_t nf = f ? seg->recv_off[n_nbr] : 0, nA = A ? seg->msend_off[n_nbr] : 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nf + nA; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < nf) {
      int k = 0;
      while (i >= seg->recv_off[k + 1]) ++k;
      peers[k].f[i - seg->recv_off[k]] = f[recv_lids[i]];
    } else {
      const int64_t j = i - nf;
      int k = 0;
      while (j >= seg->msend_off[k + 1]) ++k;
      peers[k].A[j - seg->msend_off[k]] = A[msend_src[j]];
    }
  }
  p2p_signal(done, gridDim.x, n_nbr, peers, 1, epoch);
}

__global__ void k_p2p_wait(int n_nbr, const unsigned long long *flags, unsigned long long epoch, int *timeout)
{
  const int k = threadIdx.x;
  if (k >= n_nbr) return;
  const long long t0 = clock64();
  while (ld_acquire_sys(flags + k) < epoch) {
    if (clock64() - t0 > 20000000000LL) { atomicExch(timeout, 1 + k); break; }     // ~10 s at 2 GHz
    __nanosleep(200);
  }
}

__global__ void k_p2p_unpack_x(int n_nbr, const P2PSeg *__restrict__ seg, const int *__restrict__ recv_lids,
                               const double *__restrict__ buf, double *__restrict__ x0, double *__restrict__ x1, double *__restrict__ x2)
{
  const int64_t nr = seg->recv_off[n_nbr];
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nr) return;
  int k = 0;
  while (i >= seg->recv_off[k + 1]) ++k;
  const int64_t cnt = seg->recv_off[k + 1] - seg->recv_off[k], o = i - seg->recv_off[k];
  const double *src = buf + 3 * seg->recv_off[k] + o;          // three vectors per neighbour segment
  const int l = recv_lids[i];
  if (x0) x0[l] = src[0];
  if (x1) x1[l] = src[cnt];
  if (x2) x2[l] = src[2 * cnt];
}

void p2p_free(txasm_handle h)
{
  if (!h->halo || !h->halo->p2p) return;
  P2P *P = h->halo->p2p;
  cudaStreamSynchronize(h->stream);
  for (void *p : P->opened) cudaIpcCloseMemHandle(p);
  if (P->slab) cudaFree(P->slab);
  if (P->d_peers) cudaFree(P->d_peers);
  if (P->d_seg) cudaFree(P->d_seg);
  if (P->d_done) cudaFree(P->d_done);
  if (P->d_timeout) cudaFree(P->d_timeout);
  delete P;
  h->halo->p2p = nullptr;
}

bool p2p_active(txasm_handle h) { return h->halo && h->halo->p2p && h->halo->p2p->connected && h->opt_p2p; }

static int p2p_check_timeout(txasm_handle h) { (void)h; return TXASM_OK; }

int p2p_import(txasm_handle h, double *const x[3])
{
  Halo *H = h->halo;
  P2P *P = H->p2p;
  const int64_t ns = H->send_off[H->n_nbr], nr = H->recv_off[H->n_nbr];
  ++P->epoch_import;
  const int grid = (int)std::min<int64_t>(std::max<int64_t>((ns + 255) / 256, 1), 4 * h->n_sm);
  k_p2p_push_x<<<grid, 256, 0, h->stream>>>(H->n_nbr, P->d_seg, P->d_peers, H->d_send_lids, x[0], x[1], x[2], 0, P->d_done, P->epoch_import);
  k_p2p_wait<<<1, 32, 0, h->stream>>>(H->n_nbr, P->flags, P->epoch_import, P->d_timeout);
  if (nr) k_p2p_unpack_x<<<(unsigned)((nr + 255) / 256), 256, 0, h->stream>>>(H->n_nbr, P->d_seg, H->d_recv_lids, P->recv_x, x[0], x[1], x[2]);
  TX_CUDA(h, cudaGetLastError());
  h->launches += 3;
  return p2p_check_timeout(h);
}

int p2p_export(txasm_handle h, double *f, double *A, int jac)
{
  Halo *H = h->halo;
  P2P *P = H->p2p;
  const bool do_A = jac && A && H->have_mat;
  const int64_t n = (f ? H->recv_off[H->n_nbr] : 0) + (do_A ? H->msend_off[H->n_nbr] : 0);
  ++P->epoch_export;
  const int grid = (int)std::min<int64_t>(std::max<int64_t>((n + 1023) / 1024, 1), 2 * h->n_sm);
  k_p2p_push_export<<<grid, 256, 0, h->stream>>>(H->n_nbr, P->d_seg, P->d_peers, H->d_recv_lids, H->d_msend_src, f, do_A ? A : nullptr,
                                                P->d_done + 1, P->epoch_export);
  k_p2p_wait<<<1, 32, 0, h->stream>>>(H->n_nbr, P->flags + P2P_MAX_NBR, P->epoch_export, P->d_timeout);
  TX_CUDA(h, cudaGetLastError());
  h->launches += 2;
  int rc;
  if (f && (rc = launch_unpack_add(h, H->up_f, P->recv_f, f))) return rc;
  if (do_A && (rc = launch_unpack_add(h, H->up_A, P->recv_A, A))) return rc;
  return TXASM_OK;
}

}  // namespace txasm

using namespace txasm;

extern "C" {

int txasm_halo_p2p_blob_size(void) { return (int)sizeof(P2PBlob); }

// Allocate my receive slab and describe it: `blob` (txasm_halo_p2p_blob_size() bytes) is what the neighbours need.
int txasm_halo_p2p_export(txasm_handle h, void *blob)
{
  if (!h || !blob || !h->halo) return TXASM_EINVAL;
  TX_CUDA(h, cudaSetDevice(h->device));
  Halo *H = h->halo;
  if (H->n_nbr > P2P_MAX_NBR) return set_err(h, TXASM_EUNSUPPORTED, "%d neighbours (max %d)", H->n_nbr, P2P_MAX_NBR);
  if (!H->have_mat) return set_err(h, TXASM_ESTATE, "halo_p2p_export needs halo_set and halo_set_matrix first");
  p2p_free(h);
  P2P *P = new P2P();
  H->p2p = P;
  const int64_t nr = H->n_nbr ? H->recv_off[H->n_nbr] : 0, ns = H->n_nbr ? H->send_off[H->n_nbr] : 0, mr = H->n_nbr ? H->mrecv_off[H->n_nbr] : 0;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t x_bytes = up(3 * nr * 8), f_bytes = up(ns * 8), A_bytes = up(mr * 8), fl_bytes = up(2 * P2P_MAX_NBR * 8);
  P->slab_bytes = x_bytes + f_bytes + A_bytes + fl_bytes;
  TX_CUDA(h, cudaMalloc((void **)&P->slab, P->slab_bytes));
  TX_CUDA(h, cudaMemsetAsync(P->slab, 0, P->slab_bytes, h->stream));
  P->recv_x = (double *)P->slab;
  P->recv_f = (double *)(P->slab + x_bytes);
  P->recv_A = (double *)(P->slab + x_bytes + f_bytes);
  P->flags = (unsigned long long *)(P->slab + x_bytes + f_bytes + A_bytes);
  TX_CUDA(h, cudaMalloc((void **)&P->d_done, 2 * sizeof(unsigned int)));
  TX_CUDA(h, cudaMalloc((void **)&P->d_timeout, sizeof(int)));
  TX_CUDA(h, cudaMemsetAsync(P->d_done, 0, 2 * sizeof(unsigned int), h->stream));
  TX_CUDA(h, cudaMemsetAsync(P->d_timeout, 0, sizeof(int), h->stream));
  for (int k = 0; k <= H->n_nbr; ++k) { P->seg.send_off[k] = H->send_off[k]; P->seg.recv_off[k] = H->recv_off[k]; P->seg.msend_off[k] = H->msend_off[k]; }
  TX_CUDA(h, cudaMalloc((void **)&P->d_seg, sizeof(P2PSeg)));
  TX_CUDA(h, cudaMemcpyAsync(P->d_seg, &P->seg, sizeof(P2PSeg), cudaMemcpyHostToDevice, h->stream));
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  P2PBlob b;
  memset(&b, 0, sizeof(b));
  TX_CUDA(h, cudaIpcGetMemHandle(&b.handle, P->slab));
  b.rank = H->rank; b.n_nbr = H->n_nbr;
  b.flags_off = (int64_t)(x_bytes + f_bytes + A_bytes);
  for (int k = 0; k < H->n_nbr; ++k) {
    b.e[k].nbr = H->nbr[k]; b.e[k].flag_slot = k;
    b.e[k].x_off = 3 * H->recv_off[k] * 8;                       // neighbour k delivers my ghosts it owns: three vectors
    b.e[k].f_off = (int64_t)x_bytes + H->send_off[k] * 8;        // ... and the f of my owned DOFs it ghosts
    b.e[k].A_off = (int64_t)(x_bytes + f_bytes) + H->mrecv_off[k] * 8;
  }
  memcpy(blob, &b, sizeof(b));
  return TXASM_OK;
}

// blobs: the blobs of ALL ranks, rank-major (txasm_halo_p2p_blob_size() bytes each).  Maps the neighbours' slabs.
int txasm_halo_p2p_connect(txasm_handle h, int nranks, const void *blobs)
{
  if (!h || !blobs || !h->halo || !h->halo->p2p) return TXASM_EINVAL;
  TX_CUDA(h, cudaSetDevice(h->device));
  Halo *H = h->halo;
  P2P *P = H->p2p;
  const P2PBlob *B = (const P2PBlob *)blobs;
  std::vector<P2PPeerDev> peers(H->n_nbr);
  for (int k = 0; k < H->n_nbr; ++k) {
    const int r = H->nbr[k];
    if (r < 0 || r >= nranks || B[r].rank != r) return set_err(h, TXASM_EINVAL, "p2p_connect: no blob for rank %d", r);
    const P2PBlobEntry *me = nullptr;
    for (int j = 0; j < B[r].n_nbr; ++j) if (B[r].e[j].nbr == H->rank) me = &B[r].e[j];
    if (!me) return set_err(h, TXASM_EINVAL, "p2p_connect: rank %d does not list rank %d as a neighbour", r, H->rank);
    void *base = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&base, B[r].handle, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(h, TXASM_ECUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e)); }
    P->opened.push_back(base);
    unsigned char *b = (unsigned char *)base;
    peers[k].x = (double *)(b + me->x_off);
    peers[k].f = (double *)(b + me->f_off);
    peers[k].A = (double *)(b + me->A_off);
    unsigned long long *fl = (unsigned long long *)(b + B[r].flags_off);
    peers[k].flag_import = fl + me->flag_slot;
    peers[k].flag_export = fl + P2P_MAX_NBR + me->flag_slot;
  }
  TX_CUDA(h, cudaMalloc((void **)&P->d_peers, sizeof(P2PPeerDev) * std::max(1, H->n_nbr)));
  if (H->n_nbr) TX_CUDA(h, copy_to_device_sync(h, P->d_peers, peers.data(), sizeof(P2PPeerDev) * H->n_nbr));
  P->connected = true;
  h->overlap_state = 0;
  return TXASM_OK;
}

// 0: no neighbour timed out; k+1: neighbour k never delivered (the wait kernel gave up)
int txasm_halo_p2p_status(txasm_handle h, int *timed_out)
{
  if (!h || !timed_out) return TXASM_EINVAL;
  *timed_out = 0;
  if (!h->halo || !h->halo->p2p) return TXASM_OK;
  TX_CUDA(h, cudaSetDevice(h->device));
  TX_CUDA(h, copy_to_device_sync(h, timed_out, h->halo->p2p->d_timeout, sizeof(int)));
  return TXASM_OK;
}

}  // extern "C"
