// tiles.hpp -- row-tile tables and the TMA / mbarrier helpers shared by fill_rowtile.cu and fill_brick.cu
#pragma once
#include "txasm_internal.hpp"
#include "elem_q1hex.cuh"

namespace txasm {

constexpr int PERM_STRIDE = 32;
constexpr int KF_STRIDE = 32;         // per congruent tile: interior stiffness row [27] | Jxx Jyy Jzz det of its cells | pad
constexpr int IMG_DOUBLES = 272;      // 28 + 8 rows of 27, rounded to 16 bytes
constexpr int IMG_BYTES = (IMG_DOUBLES + 28) * 8;      // bytes per row in the perm table (27 used)
constexpr int LROW_CAP = 63;         // longest row the tile path takes (length travels in 6 bits)

constexpr int RUN_UNIFORM = 1 << 30;      // RowRun::n flag: the run consists of uniform rows (stored from the constant image)
constexpr unsigned ROW_UNIFORM = 1u << 25;  // rowinfo flag of such a row
struct RowRun { long long beg; int n; int soff; };     // first A index, length in doubles, out-buffer offset in doubles
struct Tiles {
  int TR = 0;                        // rows per tile (= threads per CTA)
  int n_tiles = 0;
  int64_t n_regular = 0, n_irregular = 0;
  int te_max = 0;                    // max cells per tile
  int tep = 0;                       // compile-time cell stride of the chosen kernel instantiation
  int lrow = 27;                     // longest regular row
  bool all_affine = false;
  int *d_tile_rows = nullptr;        // [n_tiles*TR] row ids (Morton order), -1 padding
  int64_t *d_tile_cell_ptr = nullptr;// [n_tiles+1]
  int *d_tile_cells = nullptr;       // cell ids per tile, ascending
  int *d_tile_lids = nullptr;        // [sum ncells][8] LIDs in tile-cell order
  unsigned short *d_adjl = nullptr;  // [n_tiles][TR][8] tile-local cell index of the cell having row r as vertex a
  unsigned char *d_perm = nullptr;   // [n_rows][32] canonical neighbour -> CSR slot (0xFF absent); freed after setup
  unsigned *d_tile_rowinfo = nullptr;  // [n_tiles*TR] out-buffer offset | len<<16 | zero-fill<<24 (0xFFFF: no row)
  int64_t *d_run_ptr = nullptr;        // [n_tiles+1]
  RowRun *d_runs = nullptr;
  int out_doubles = 0;                 // out-buffer size (doubles)
  unsigned char *d_tile_perm = nullptr;
  unsigned char *d_tile_cong = nullptr;   // [n_tiles] 1: all cells of the tile are translates of its first cell; 2: and all rows uniform
  unsigned char *d_row_uniform = nullptr;   // [n_rows] 1: the row belongs to a tile of the uniform range
  int n_uni = 0;                          // tiles [0, n_uni): congruent, axis-aligned, all rows uniform (class 7)
  int n_brick = 0;                        // tiles [0, n_brick) of those: the cells form a full tensor brick (class 15)
  unsigned char *d_brick_rec = nullptr;   // [n_brick] records of BRICK_REC_BYTES (fill_brick.cu)
  unsigned char *d_brick_flag = nullptr;  // [n_tiles] scratch of brick_classify
  int n_edge = 0;                         // tiles [n_uni, n_edge): lattice tiles with rows on their faces (k_fill_edge)
  unsigned char *d_edge_rec = nullptr;    // [n_edge - n_uni] records
  unsigned short *d_edge_code = nullptr;  // per CSR entry of their non-uniform runs, in A order: state * 27 + canonical neighbour (0xFFFF: zero)
  bool edge_attr_set = false;
  bool edge_ok = false;                   // every edge tile fits k_fill_edge's shared memory (checked when the codes are laid out)
  int64_t n_edge_code = 0;
  int64_t edge_run0 = 0;                  // run_ptr[n_uni]
  int64_t *d_edge_eoff = nullptr;         // [runs of those tiles + 1] first code of the run (index: run - run_ptr[n_uni])
  double *d_shapes = nullptr;             // [n_shapes][SHAPE_STRIDE] distinct cell shapes of the brick tiles
  int n_shapes = 0;
  bool brick_attr_set = false;
  bool uni_attr_set = false;              // shared-memory opt-in of k_fill_uniform done for this handle's device
  double *d_tile_kf = nullptr;            // [n_tiles][27] stiffness row of an interior node of a congruent tile
  int grid = 0;
  int *d_irregular = nullptr;        // list of irregular rows
  int smem_bytes = 0;
  int ctas_per_sm = 0;
};

// canonical 27-point neighbour index of vertex b seen from vertex a of the same cell
__host__ __device__ constexpr int canon(int a, int b)
{
  return ((hex_sx(b) - hex_sx(a)) / 2 + 1) + 3 * ((hex_sy(b) - hex_sy(a)) / 2 + 1) + 9 * ((hex_sz(b) - hex_sz(a)) / 2 + 1);
}

struct TileArgs {
  const int *tile_rows;                 // [n_tiles*TR] row id or -1
  const int64_t *tile_cell_ptr;
  const int *tile_cells;
  const int *tile_lids;                 // [sum ncells][8]
  const unsigned short *adjl;           // [n_tiles][TR][8]
  const unsigned *tile_rowinfo;         // [n_tiles*TR] out offset | len<<16 | zero<<24
  const int64_t *run_ptr;               // [n_tiles+1]
  const RowRun *runs;
  const unsigned char *tile_perm;       // [n_tiles*TR][32]
  int lrow;                             // out-buffer row stride
  int n_tiles;
  int stage_bytes;                      // offset of the LID buffer in dynamic shared memory
  int tma_store;                        // A_values is 16-byte aligned: row runs leave by TMA bulk stores
  int t_begin;                          // this launch covers tiles [t_begin, n_tiles)
  const unsigned char *tile_cong;       // [n_tiles] congruent-tile flags
  const double *tile_kf;                // [n_tiles][27] interior stiffness row of congruent tiles
  const int *row_dir;                   // fused Dirichlet: [n_rows] index into dir_vals or -1; NULL = not fused
  const double *dir_vals;
};

// interior row: canonical neighbour j = (dx,dy,dz)+1 is vertex nb_vert(j) of the cell in which the row is vertex nb_cell(j)
__host__ __device__ constexpr int hex_vertex(int bx, int by, int bz) { return 4 * bz + 2 * by + (bx ^ by); }
__host__ __device__ constexpr int nb_cell(int j)
{
  return hex_vertex((j % 3 - 1) < 0 ? 1 : 0, ((j / 3) % 3 - 1) < 0 ? 1 : 0, (j / 9 - 1) < 0 ? 1 : 0);
}
__host__ __device__ constexpr int nb_vert(int j)
{
  return hex_vertex((j % 3 - 1) > 0 ? 1 : 0, ((j / 3) % 3 - 1) > 0 ? 1 : 0, (j / 9 - 1) > 0 ? 1 : 0);
}
// ---- mbarrier / TMA bulk copy helpers (sm_90+ PTX; SASS: SYNCS / UBLKCP)
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(mbar), "r"(parity) : "memory");
}
// one thread: announce `bytes` and start the bulk copy global -> shared; completion flips the mbarrier phase
__device__ __forceinline__ void bulk_load(unsigned dst, const void *src, unsigned bytes, unsigned mbar)
{
  if (bytes == 0) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory"); return; }
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
// kernel timeline (TXASM_TIMELINE=1): one thread per CTA stamps the earliest start / latest end of kernel `k`
__device__ __forceinline__ unsigned long long tx_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void tx_stamp(unsigned long long *dbg, int k, bool end)
{
  if (!dbg || threadIdx.x != 0) return;
  if (end) atomicMax(dbg + 2 * k + 1, tx_globaltimer()); else atomicMin(dbg + 2 * k, tx_globaltimer());
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }


// fill_brick.cu
int brick_classify(txasm_handle h);                       // sets bit 3 of tile_cong for brick tiles (before the reorder)
int brick_build(txasm_handle h);                          // records + shape table for tiles [0, n_brick) (after it)
void brick_free(txasm_handle h);
bool fill_brick_eligible(txasm_handle h, const FillArgs &a);
int launch_fill_brick(txasm_handle h, const FillArgs &a, cudaStream_t stream);
bool fill_edge_eligible(txasm_handle h, const FillArgs &a);
int launch_fill_edge(txasm_handle h, const FillArgs &a, cudaStream_t stream, const int *row_dir, const double *dir_vals);

// Shared-memory carve-out of the fill kernels (TXASM_CARVEOUT, percent of the SM's L1/shared array; default: the driver's
// choice).  Tuning aid: two kernels share an SM only under one carve-out.  Measured: no effect on the lattice kernels (they do
// not co-reside anyway), and 100 % costs the general-hexahedron row pass its L1 (9.8 -> 12.2 ms).
inline int tx_carveout()
{
  static const int v = [] { const char *e = getenv("TXASM_CARVEOUT"); return e ? atoi(e) : -1; }();
  return v;
}
int edge_codes_refresh(txasm_handle h);   // fill_brick.cu: entry codes of the edge tiles (again after the Dirichlet rows changed)
template <typename K> inline void tx_set_carveout(K k)
{
  if (tx_carveout() >= 0) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, tx_carveout());
}

}  // namespace txasm
