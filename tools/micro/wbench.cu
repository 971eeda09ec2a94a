// wbench.cu -- write-bandwidth microbenchmarks behind the design of the uniform-tile stores.
// How fast can 3.6 GB of fp64 leave the SMs as (a) memset, (b) plain vector stores, (c) TMA bulk stores from a constant
// shared-memory image, in linear order and in the order / chunking the row tiles produce (x-lines of 8 rows = 1728 B,
// 8 x 8 x 4-node tiles in Morton order on a 257^3 lattice)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o wbench wbench.cu && ./wbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_plain(double2 *p, size_t n2) {
  const double2 v = make_double2(1.0, 2.0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// chunk list: each entry = start (in doubles, even) ; every chunk has `len` doubles (even).  CTA b takes chunks b, b+G, ...
// grouped by `per` consecutive chunks per "tile" (a CTA issues the `per` chunks of a tile from 32/.. lanes like the kernel).
__global__ void __launch_bounds__(256) k_bulk(double *A, const long long *starts, long long n_chunks, int len, int per, int strided) {
  extern __shared__ __align__(16) double img[];
  for (int i = threadIdx.x; i < len; i += blockDim.x) img[i] = (double)(i % 27);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const unsigned img_s = (unsigned)__cvta_generic_to_shared(img);
  const long long n_tiles = (n_chunks + per - 1) / per;
  const long long per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  for (long long it = 0; it < per_cta; ++it) {
    const long long t = strided ? (blockIdx.x + it * gridDim.x) : (blockIdx.x * per_cta + it);
    if (t >= n_tiles) break;
    const int r = (threadIdx.x & 31) * 8 + (threadIdx.x >> 5);
    for (int c = r; c < per; c += 256) {
      const long long ci = t * per + c;
      if (ci < n_chunks)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(A + starts[ci]), "r"(img_s), "r"((unsigned)len * 8u) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if ((it & 7) == 7) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__global__ void __launch_bounds__(256) k_bulk_var(double *A, const long long *starts, const int *lens, long long n_chunks, int per) {
  extern __shared__ __align__(16) double img[];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) img[i] = (double)(i % 27);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const unsigned img_s = (unsigned)__cvta_generic_to_shared(img);
  const long long n_tiles = (n_chunks + per - 1) / per;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int r = (threadIdx.x & 31) * 8 + (threadIdx.x >> 5);
    if (r < per && t * per + r < n_chunks) {
      const long long ci = t * per + r;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(A + starts[ci]), "r"(img_s), "r"((unsigned)lens[ci] * 8u) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

static unsigned long long spread3(unsigned long long v) {
  v &= 0x1fffffull; v = (v | (v << 32)) & 0x1f00000000ffffull; v = (v | (v << 16)) & 0x1f0000ff0000ffull;
  v = (v | (v << 8)) & 0x100f00f00f00f00full; v = (v | (v << 4)) & 0x10c30c30c30c30c3ull; v = (v | (v << 2)) & 0x1249249249249249ull; return v;
}

int main() {
  const int N = 256;                          // interior lattice 256^3 rows (pretend), 27 doubles per row
  const size_t n_rows = (size_t)N * N * N, n = n_rows * 27;
  double *A; CK(cudaMalloc(&A, n * 8 + 4096));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char *name, float ms) { printf("%-64s %7.3f ms  %7.1f GB/s\n", name, ms, n * 8.0 / ms / 1e6); fflush(stdout); };
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); CK(cudaMemsetAsync(A, 0, n * 8)); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
  }
  report("cudaMemset", ms);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_plain<<<148 * 8, 256>>>((double2 *)A, n / 2); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
  }
  report("plain st.v2.f64, grid-stride", ms);

  struct Cfg { const char *name; int bx, by, bz; int morton; int strided; int ctas; };
  const Cfg cfgs[] = {
    {"bulk: 8x8x4 tiles, Morton order, stride-G CTAs, 4/SM (as k_fill_brick)", 8, 8, 4, 1, 1, 4},
    {"bulk: 8x8x4 tiles, Morton order, stride-G CTAs, 3/SM", 8, 8, 4, 1, 1, 3},
    {"bulk: 8x8x4 tiles, Morton order, stride-G CTAs, 2/SM", 8, 8, 4, 1, 1, 2},
    {"bulk: 8x8x4 tiles, Morton order, stride-G CTAs, 1/SM", 8, 8, 4, 1, 1, 1},
    {"bulk: 8x8x4 tiles, Morton order, contiguous tile ranges per CTA, 4/SM", 8, 8, 4, 1, 0, 4},
    {"bulk: 8x8x4 tiles, lexicographic tile order, stride-G, 4/SM", 8, 8, 4, 0, 1, 4},
    {"bulk: 16x4x4 tiles, lexicographic, stride-G, 4/SM", 16, 4, 4, 0, 1, 4},
    {"bulk: 32x4x2 tiles, lexicographic, stride-G, 4/SM", 32, 4, 2, 0, 1, 4},
    {"bulk: 64x2x2 tiles, lexicographic, stride-G, 4/SM", 64, 2, 2, 0, 1, 4},
    {"bulk: 256x1x1 tiles (whole x-lines), lexicographic, stride-G, 4/SM", 256, 1, 1, 0, 1, 4},
    {"bulk: 256x1x1 tiles, lexicographic, stride-G, 2/SM", 256, 1, 1, 0, 1, 2},
  };
  for (const Cfg &c : cfgs) {
    // chunks: x-lines of bx rows; 216-double (8 rows) copies like the kernel => chunk = min(bx, 8) rows... keep whole x-line of the tile, split in 8-row copies
    const int tx = N / c.bx, ty = N / c.by, tz = N / c.bz;
    std::vector<std::pair<unsigned long long, int>> order;   // (key, tile id)
    for (int k = 0; k < tz; ++k) for (int j = 0; j < ty; ++j) for (int i = 0; i < tx; ++i) {
      unsigned long long key = c.morton ? (spread3(i * c.bx) | (spread3(j * c.by) << 1) | (spread3(k * c.bz) << 2)) : ((unsigned long long)((k * ty + j)) * tx + i);
      order.push_back({key, (k * ty + j) * tx + i});
    }
    std::sort(order.begin(), order.end());
    const int len_rows = 8;                                 // rows per bulk copy (216 doubles, 1728 B)
    const int copies_per_line = c.bx / len_rows, per = c.by * c.bz * copies_per_line;
    std::vector<long long> starts; starts.reserve(n_rows / len_rows);
    for (auto &o : order) {
      const int id = o.second, i = id % tx, j = (id / tx) % ty, k = id / (tx * ty);
      for (int kk = 0; kk < c.bz; ++kk) for (int jj = 0; jj < c.by; ++jj) for (int cc = 0; cc < copies_per_line; ++cc) {
        const long long row = ((long long)(k * c.bz + kk) * N + (j * c.by + jj)) * N + i * c.bx + cc * len_rows;
        starts.push_back(row * 27);
      }
    }
    long long *d_s; CK(cudaMalloc(&d_s, starts.size() * 8)); CK(cudaMemcpy(d_s, starts.data(), starts.size() * 8, cudaMemcpyHostToDevice));
    // starts must be even (16-byte aligned): row*27 even iff row even; rows are multiples of 8 here
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_bulk<<<148 * c.ctas, 256, 216 * 8>>>(A, d_s, (long long)starts.size(), 216, per, c.strided);
      cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
    }
    report(c.name, ms);
    cudaFree(d_s);
  }
  // the real mesh: 257 nodes per line, runs start at arbitrary rows => chunk starts are not sector (32 B) aligned.
  // (a) as is: every chunk [216 r, 216 r + 1728) bytes, r = 2 + 8 m (16-byte aligned, half of them split a sector)
  // (b) boundaries between x-adjacent chunks moved up to the next multiple of 128 bytes (the neighbour's image is the same)
  for (int variant = 0; variant < 3; ++variant) {
    const int M = 257, per_line = 31;
    std::vector<std::pair<unsigned long long, long long>> tiles;    // Morton key of the 8x8x4 tile, first row
    std::vector<long long> starts; std::vector<int> lens;
    for (int k0 = 1; k0 + 4 <= M - 1; k0 += 4) for (int j0 = 1; j0 + 8 <= M - 1; j0 += 8) for (int m = 0; m < per_line; ++m)
      tiles.push_back({spread3(m) | (spread3(j0 / 8) << 1) | (spread3(k0 / 4) << 2), ((long long)k0 * M + j0) * M + 2 + 8 * m});
    std::sort(tiles.begin(), tiles.end());
    for (auto &t : tiles)
      for (int kk = 0; kk < 4; ++kk) for (int jj = 0; jj < 8; ++jj) {
        const long long row = t.second + ((long long)kk * M + jj) * M;
        long long b = row * 27, e = b + 216;
        const int m = (int)((row % M - 2) / 8);
        if (variant >= 1) {
          const long long al = variant == 1 ? 16 : 4;                  // 128-byte or 32-byte boundaries
          if (m > 0) b = (b + al - 1) / al * al;
          if (m < per_line - 1) e = (e + al - 1) / al * al;
        }
        b &= ~1LL; e = (e + 1) & ~1LL;                                 // 16-byte alignment of the bulk copy (the kernel stores odd ends by hand)
        starts.push_back(b); lens.push_back((int)(e - b));
      }
    long long *d_s; int *d_l;
    CK(cudaMalloc(&d_s, starts.size() * 8)); CK(cudaMemcpy(d_s, starts.data(), starts.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_l, lens.size() * 4)); CK(cudaMemcpy(d_l, lens.data(), lens.size() * 4, cudaMemcpyHostToDevice));
    double bytes = 0; for (int l : lens) bytes += 8.0 * l;
    for (int ctas : {2, 4}) {
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_bulk_var<<<148 * ctas, 256, 256 * 8>>>(A, d_s, d_l, (long long)starts.size(), 32);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("257-lattice 8x8x4 Morton tiles, %s, %d CTA/SM: %7.3f ms  %7.1f GB/s (%.2f GB)\n",
             variant == 0 ? "true row boundaries (unaligned)" : (variant == 1 ? "boundaries moved to 128 B" : "boundaries moved to 32 B"), ctas, ms, bytes / ms / 1e6, bytes / 1e9);
    }
    cudaFree(d_s); cudaFree(d_l);
  }
  // longer bulk copies on whole x-lines: 32 rows (6912 B) and 256 rows per copy from a longer image
  for (int rows_per_copy : {16, 32, 64}) {
    std::vector<long long> starts;
    for (long long row = 0; row < (long long)n_rows; row += rows_per_copy) starts.push_back(row * 27);
    long long *d_s; CK(cudaMalloc(&d_s, starts.size() * 8)); CK(cudaMemcpy(d_s, starts.data(), starts.size() * 8, cudaMemcpyHostToDevice));
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 27 * 8);
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_bulk<<<148 * 4, 256, rows_per_copy * 27 * 8>>>(A, d_s, (long long)starts.size(), rows_per_copy * 27, 256 / rows_per_copy * 4, 1);
      cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
    }
    char nm[128]; snprintf(nm, sizeof nm, "bulk: linear, %d rows (%d B) per copy, 4/SM", rows_per_copy, rows_per_copy * 216);
    report(nm, ms);
    cudaFree(d_s);
  }
  return 0;
}
