// Octree partition of a point cloud into blocks on the GPU: a stable counting sort of the points by the Morton key of their
// block, output-compatible with the reference's src/utils/octree_coding.py:68-113 (7.6 s of Python on
// longdress, octree_coding.py:66) -- same blocks in the same (Morton) order, points of a block in their input order, local
// coordinates.  Besides the float64 rows the reference returns, the kernels emit what the block loops consume directly:
// int16 (block, z, y, x) rows, already on the device.
//
//   1. keys      per point: block id = floor(p / block_size) per axis -> (z,y,x)-interleaved Morton key; presence table
//   2. rank      exclusive scan of the presence table: key -> block index in Morton order (one CTA; <= 2^18 keys)
//   3. tile hist per tile of 1024 points: number of its points in every block it touches (global atomics, integers)
//   4. col scan  per block: running sum over the tiles -> first output row of (tile, block); block offsets
//   5. scatter   per tile: a point's rank among the earlier points of its tile with the same block (stable), rows written
//                with the block origin subtracted
// All arithmetic on indices is integer; the result does not depend on the schedule.
#include "common.cuh"

namespace pccgeo {

constexpr int OT_TILE = 1024;

__device__ __forceinline__ uint32_t morton3(uint32_t x, uint32_t y, uint32_t z, int level) {
  uint32_t key = 0;
  for (int b = level - 1; b >= 0; --b) key = (key << 3) | (((z >> b) & 1u) << 2) | (((y >> b) & 1u) << 1) | ((x >> b) & 1u);
  return key;   // octree_coding.py:93-97: most significant bit first, z then y then x
}

__global__ void octree_keys_kernel(const double* __restrict__ rows, long long n, int cols, double block_size, int level,
                                   uint32_t* __restrict__ key, int* __restrict__ present, int* __restrict__ err) {
  const uint32_t side = 1u << level;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double* r = rows + i * cols;
    const double fx = floor(r[0] / block_size), fy = floor(r[1] / block_size), fz = floor(r[2] / block_size);
    if (!(fx >= 0 && fy >= 0 && fz >= 0 && fx < side && fy < side && fz < side)) { *err = 1; key[i] = 0; continue; }
    const uint32_t k = morton3((uint32_t)fx, (uint32_t)fy, (uint32_t)fz, level);
    key[i] = k;
    present[k] = 1;
  }
}

// rank_of_key[k] = number of present keys below k; *n_blocks = number of present keys.  One CTA of 1024 threads.
__global__ void octree_rank_kernel(const int* __restrict__ present, int n_keys, int* __restrict__ rank_of_key, int* __restrict__ n_blocks) {
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_keys; base += blockDim.x) {
    const int k = base + threadIdx.x;
    const int v = k < n_keys ? present[k] : 0;
    int s = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if ((threadIdx.x & 31) >= o) s += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = wsum[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += t;
      }
      wsum[threadIdx.x] = w;
    }
    __syncthreads();
    const int before = carry + (threadIdx.x >= 32 ? wsum[(threadIdx.x >> 5) - 1] : 0) + s - v;
    if (k < n_keys) rank_of_key[k] = before;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_blocks = carry;
}

__global__ void octree_tile_hist_kernel(const uint32_t* __restrict__ key, const int* __restrict__ rank_of_key, long long n, int nb,
                                        int* __restrict__ tile_count) {
  const long long i = (long long)blockIdx.x * OT_TILE + threadIdx.x;
  if (i < n) atomicAdd(&tile_count[(long long)blockIdx.x * nb + rank_of_key[key[i]]], 1);
}

// per block b (thread): tile_count[t][b] -> exclusive running sum over tiles (in place); totals[b]
__global__ void octree_col_scan_kernel(int* __restrict__ tile_count, int n_tiles, int nb, long long* __restrict__ totals) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  int run = 0;
  for (int t = 0; t < n_tiles; ++t) {
    const int c = tile_count[(long long)t * nb + b];
    tile_count[(long long)t * nb + b] = run;
    run += c;
  }
  totals[b] = run;
}

// offsets[b] = sum of totals below b (nb + 1 entries) -- one CTA, sequential over chunks (nb <= 2^18)
__global__ void octree_offsets_kernel(const long long* __restrict__ totals, int nb, long long* __restrict__ offsets) {
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int b = 0; b < nb; ++b) { offsets[b] = run; run += totals[b]; }
    offsets[nb] = run;
  }
}

__global__ void __launch_bounds__(OT_TILE) octree_scatter_kernel(const double* __restrict__ rows, const uint32_t* __restrict__ key,
                                                                 const int* __restrict__ rank_of_key, const int* __restrict__ tile_base,
                                                                 const long long* __restrict__ offsets, long long n, int cols, int nb,
                                                                 double block_size, int level, double* __restrict__ out_rows,
                                                                 int16_t* __restrict__ out_coords) {
  __shared__ int blk[OT_TILE];
  const long long i = (long long)blockIdx.x * OT_TILE + threadIdx.x;
  const uint32_t k = i < n ? key[i] : 0xffffffffu;
  const int b = i < n ? rank_of_key[k] : -1;
  blk[threadIdx.x] = b;
  __syncthreads();
  if (i >= n) return;
  int local = 0;   // earlier points of this tile in the same block: keeps the input order inside a block
  for (int j = 0; j < (int)threadIdx.x; ++j) local += blk[j] == b;
  const long long dst = offsets[b] + tile_base[(long long)blockIdx.x * nb + b] + local;
  // block origin from the key (de-interleave)
  uint32_t bx = 0, by = 0, bz = 0;
  for (int l = 0; l < level; ++l) {
    bx |= ((k >> (3 * l)) & 1u) << l;
    by |= ((k >> (3 * l + 1)) & 1u) << l;
    bz |= ((k >> (3 * l + 2)) & 1u) << l;
  }
  const double* r = rows + i * cols;
  const double lx = r[0] - bx * block_size, ly = r[1] - by * block_size, lz = r[2] - bz * block_size;
  if (out_rows) {
    double* o = out_rows + dst * cols;
    o[0] = lx; o[1] = ly; o[2] = lz;
    for (int c = 3; c < cols; ++c) o[c] = r[c];
  }
  if (out_coords) {   // (block, i0, i1, i2) like pccgeo_blocks_to_coords_host: the block loops' densify input
    short4 q;
    q.x = (short)(b & 0x7fff); q.y = (short)lx; q.z = (short)ly; q.w = (short)lz;
    reinterpret_cast<short4*>(out_coords)[dst] = q;
  }
}

}  // namespace pccgeo

using namespace pccgeo;

static inline size_t ot_keys_bytes(long long n) { return (size_t)((n * 4 + 255) / 256 * 256); }

// workspace of stage 1: keys (n uint32, padded), presence and rank tables (2 x 8^level int32), {n_blocks, err} (64 B)
extern "C" size_t pccgeo_octree_ws_bytes(long long n, int level) {
  return ot_keys_bytes(n) + (size_t)(1LL << (3 * level)) * 8 + 64;
}
// workspace of stage 2: per-block totals (int64) + the (tiles x blocks) running-count table (int32)
extern "C" size_t pccgeo_octree_ws2_bytes(long long n, int n_blocks) {
  const long long n_tiles = (n + OT_TILE - 1) / OT_TILE;
  return (size_t)n_blocks * 8 + (size_t)n_tiles * n_blocks * 4 + 256;
}

// Stage 1 (async): keys, presence, ranks.  The caller reads {n_blocks, err} (two int32 at ws + pccgeo_octree_ws_bytes - 64) after
// a stream sync, then calls stage 2.  rows: device float64 (n, cols >= 3); 1 <= level <= 6.  Replaces octree_coding.py:85-101.
extern "C" int pccgeo_octree_partition_keys(const double* rows, long long n, int cols, double block_size, int level, void* ws, void* stream) {
  PCCGEO_REQUIRE(rows && ws && n > 0 && cols >= 3 && level >= 1 && level <= 6 && block_size >= 1, "octree_partition_keys: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_keys = 1 << (3 * level);
  uint8_t* w = (uint8_t*)ws;
  uint32_t* key = (uint32_t*)w;
  int* present = (int*)(w + ot_keys_bytes(n));
  int* rank_of_key = present + n_keys;
  int* meta = rank_of_key + n_keys;   // [0] n_blocks, [1] err
  PCCGEO_CUDA(cudaMemsetAsync(present, 0, sizeof(int) * n_keys, st));
  PCCGEO_CUDA(cudaMemsetAsync(meta, 0, 64, st));
  long long g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  octree_keys_kernel<<<(int)g, 256, 0, st>>>(rows, n, cols, block_size, level, key, present, meta + 1);
  int rc = check_launch("octree_keys_kernel");
  if (rc) return rc;
  octree_rank_kernel<<<1, 1024, 0, st>>>(present, n_keys, rank_of_key, meta);
  return check_launch("octree_rank_kernel");
}

// Stage 2 (async): grouped rows (float64, like the reference's blocks: local coordinates, input order inside a block) and / or
// int16 (block, i0, i1, i2) rows for pccgeo_densify_from, + offsets (n_blocks + 1, device int64).  Replaces the per-point loop
// of octree_coding.py:103-111.
extern "C" int pccgeo_octree_partition_scatter(const double* rows, long long n, int cols, double block_size, int level, int n_blocks,
                                               const void* ws, void* ws2, double* out_rows, int16_t* out_coords, long long* offsets,
                                               void* stream) {
  PCCGEO_REQUIRE(rows && ws && ws2 && offsets && n > 0 && n_blocks > 0 && (out_rows || out_coords), "octree_partition_scatter: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_keys = 1 << (3 * level);
  const long long n_tiles = (n + OT_TILE - 1) / OT_TILE;
  const uint8_t* w = (const uint8_t*)ws;
  const uint32_t* key = (const uint32_t*)w;
  const int* rank_of_key = (const int*)(w + ot_keys_bytes(n)) + n_keys;
  long long* totals = (long long*)ws2;
  int* tile_count = (int*)(totals + n_blocks);
  PCCGEO_CUDA(cudaMemsetAsync(tile_count, 0, sizeof(int) * n_tiles * n_blocks, st));
  octree_tile_hist_kernel<<<(int)n_tiles, OT_TILE, 0, st>>>(key, rank_of_key, n, n_blocks, tile_count);
  int rc = check_launch("octree_tile_hist_kernel");
  if (rc) return rc;
  octree_col_scan_kernel<<<(n_blocks + 127) / 128, 128, 0, st>>>(tile_count, (int)n_tiles, n_blocks, totals);
  rc = check_launch("octree_col_scan_kernel");
  if (rc) return rc;
  octree_offsets_kernel<<<1, 32, 0, st>>>(totals, n_blocks, offsets);
  rc = check_launch("octree_offsets_kernel");
  if (rc) return rc;
  octree_scatter_kernel<<<(int)n_tiles, OT_TILE, 0, st>>>(rows, key, rank_of_key, tile_count, offsets, n, cols, n_blocks, block_size, level,
                                                          out_rows, out_coords);
  return check_launch("octree_scatter_kernel");
}
