// Batched GICP alignment on sm_100a.
//
// Replaces RegistrationGICP::RegisterPointClouds (reference src/RegistrationGICP.cc:5-20) and the
// small_gicp code behind it (Thirdparty/small_gicp/include/small_gicp: util/downsampling.hpp,
// ann/kdtree.hpp, util/normal_estimation.hpp, factors/gicp_factor.hpp,
// registration/reduction_omp.hpp, registration/optimizer.hpp, util/lie.hpp) for a batch of
// independent cloud pairs.  All arithmetic is fp64 as in the reference.
//
//   k_group_insert / k_group_rank / k_group_fill   deterministic "group by 63-bit key" built on an
//                      open-addressing hash table: voxels (leaf 0.02 m) for the downsampling, grid
//                      cells (0.05 m) for the neighbour search.  Groups are numbered by first
//                      occurrence and member lists are in input order, so every sum below has a
//                      fixed order (results are run-to-run reproducible).
//   k_voxel_mean       per-voxel mean in input order          (voxelgrid_sampling :23-78)
//   k_knn_cov          exact 10-NN by expanding grid shells (KdTree::knn_search semantics: exact
//                      k nearest) + covariance regularisation (normal_estimation.hpp:66-92)
//   k_nn_corr          exact 1-NN within max_correspondence_distance (seeded with the previous correspondence)
//   k_linearize        GICPFactor::linearize (gicp_factor.hpp:34-73), per-warp partial sums of H, b, e
//   k_lm_begin / k_error / k_lm_decide              LevenbergMarquardtOptimizer (optimizer.hpp:83-148)
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.cuh"

#include <type_traits>
namespace gfs {

static const unsigned long long KEY_EMPTY = 0xffffffffffffffffull;
static const int RED_N = 29;  // 21 (H upper) + 6 (b) + 1 (e) + 1 (inlier count)

struct GicpDev {
  int nmax, hsize;  // points capacity per cloud, hash slots per cloud (power of two)
  int nblk;         // partial-sum blocks per pair
  double voxel, cell, max_dist, rot_eps, trans_eps;
  int k, max_iter;
  // which clouds a launch works on: cloud(y) = cbase + y * cstep.  align: all 2 * pairs clouds (0, 1); track: only the
  // new cloud of every sequence (slot, 2).  swap = 0: cloud 2p is the target of pair p and 2p + 1 its source; 1: the
  // other way round (track mode alternates, so last call's source is this call's target without being rebuilt).
  int cbase, cstep, swap;
  int inputAll;     // 1: every cloud's raw points come from the `src` array (track mode); 0: even clouds from `tgt`
  int errPpt;       // source points per thread of k_error (GFS_GICP_ERR_PPT, even, 2..32)
  int linPpt;       // source points per thread of k_linearize (GFS_GICP_LIN_PPT, 1..32)
  int cellOrder;    // 1: neighbour-search kernels take their queries in grid-cell order (rec[]), not in point order
  int octSorted;    // 1: k_cell_sort ran (an octant kernel is selected): k_cell_pack takes the cell's members from slotOf[]
  int dense;        // 1: the k-NN grid is the dense sorted grid (dS / cellBox = its region), 0: the hash grid (tab)
  int dcap, daxis;  // dense grid: cells per cloud it can hold, per-axis limit applied when a cloud's box would not fit
  int dstride;      // dcap + 4: entries between two clouds' arrays (a multiple of 4: every array starts 16-byte aligned)
  unsigned* dS;               // [clouds][dstride] dense grid: first record of every cell of the region, x fastest (an exclusive
                              // prefix sum of the cell populations: a run of cells xa..xb of one row is records S[xa] .. S[xb + 1])
  unsigned long long* dSum;   // [clouds][3] sum of the points' cell coordinates (centres the region of an oversized cloud)
  // per cloud (2 * pairs clouds; cloud 2p = target of pair p, 2p+1 = source)
  unsigned long long* keys;  // [clouds][hsize]
  int* minIdx;               // [clouds][hsize]
  int* count;                // [clouds][hsize]
  int* start;                // [clouds][hsize]
  int* cursor;               // [clouds][hsize]
  int* rank;                 // [clouds][hsize]
  int* slotOf;               // [clouds][nmax]
  int* members;              // [clouds][nmax]
  int* nIn;                  // [clouds] input point count
  int* nDown;                // [clouds] downsampled point count
  int* nCells;               // [clouds] occupied grid cells
  int* nFall;                // [clouds] queries the cell-centric 10-NN kernel left to the per-query kernel
  int* cellBox;              // [clouds][6] min/max cell coordinate of the grid
  double* pts;               // [clouds][nmax][4]
  uint4* tab;                // [clouds][hsize] kNN grid slot {key lo, key hi, start, count}: one 16-byte probe
  double* rec;               // [clouds][nmax][4] downsampled points in cell order {x, y, z, index bits}
  float4* recf;              // [clouds][nmax] the same records as float32 RELATIVE TO THEIR CELL's origin {x, y, z, index bits}
  int* nbr;                  // [clouds][nmax][10] the 10 nearest points of every point (itself first), in (distance, index) order
  double* nbrR2;             // [clouds][nmax] squared distance to the 10th of them (< 0: fewer than 10 found -- no certificate)
  int* knnList;              // [clouds][nmax][16] candidates k_knn_cov_warp hands to k_knn_select_cov
  unsigned char* knnCnt;     // [clouds][nmax] their number (0: the query went to the hand-over list instead)
  uint2* oct;                // [clouds][hsize] per occupied cell: its records are sorted by octant (bit 0 / 1 / 2 = upper half in
                             // x / y / z); eight 8-bit counts.  0xffffffff, 0xffffffff: more than 255 records, not sorted
  float nnBoundA;            // error bound of a float32 cell-local squared distance v: nnBoundA * sqrt(v) + 5e-7 * v + 1e-14
  double* cov;               // [clouds][nmax][6]
  // per pair
  int* corr;                 // [pairs][nmax] target index of source point (-1: none)
  double* maha;              // [pairs][nmax][6]
  double* partial;           // [pairs][nblk][warps per block][RED_N]
  double* partialE;          // [pairs][nblk]
  double* state;             // [pairs][LM_STATE]
  int* istate;               // [pairs][LM_ISTATE]
  int* counters;             // [4]: needTrial, active
  int* activeList;           // [2][listStride] pairs with work left, written by the decisions of a round the host checks (list
                             // listWrite), read by the rounds after it (list listRead): the optimiser kernels then launch only as many
                             // CTA rows as there are unfinished pairs.  Without it the 6 % of pairs that run all 20 iterations kept
                             // every kernel at full grid size: 16 of 21 rounds spent their time retiring empty CTAs.
  int listRead, listWrite, listStride;   // -1: none
  int* tickets;              // [pairs][2] blocks of the pair that finished k_linearize / k_error (the last one does the pair's LM step)
};

// per-pair LM state layout (doubles)
enum { S_T = 0, S_NEWT = 12, S_H = 24, S_B = 45, S_E = 51, S_LAMBDA = 52, S_DELTA = 53, LM_STATE = 60 };
// per-pair LM state layout (ints)
enum { I_ACTIVE = 0, I_NEED = 1, I_CONV = 2, I_ITER = 3, I_TRIAL = 4, I_INL = 5, I_INNER = 6, I_SUCCESS = 7, I_OUTER = 8, LM_ISTATE = 12 };
// Every pair walks its own LM state machine (optimizer.hpp:97-141) through the launches the host enqueues in "rounds" of
// [search, linearize, begin, error, decide]: a pair BETWEEN outer iterations (I_ACTIVE && !I_NEED) takes part in the first three
// kernels of a round, a pair IN an iteration (I_NEED: a lambda trial is waiting for its error) in the last two -- a rejected
// trial simply uses the next round's error / decide slots.  I_OUTER is the pair's own outer-iteration index.  The host no
// longer steers each trial: it only reads the number of unfinished pairs every few rounds to know when to stop.

__device__ __forceinline__ int cloud_of(const GicpDev& D, int y) { return D.cbase + y * D.cstep; }
// the pair a CTA row of an optimiser kernel works on: row y itself until the host has seen the first list of unfinished pairs
__device__ __forceinline__ int pair_of(const GicpDev& D, int y) { return D.listRead < 0 ? y : D.activeList[D.listRead * D.listStride + y]; }
__device__ __forceinline__ int tgt_cloud(const GicpDev& D, int p) { return 2 * p + D.swap; }
__device__ __forceinline__ int src_cloud(const GicpDev& D, int p) { return 2 * p + 1 - D.swap; }
__device__ __forceinline__ const float* raw_points(const GicpDev& D, int c, const float* tgt, const float* src, int stride) {
  return ((D.inputAll || (c & 1)) ? src : tgt) + (size_t)(c >> 1) * stride * 4;
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
__device__ __forceinline__ int fast_floor_d(double v) {  // util/fast_floor.hpp:12-15
  const int n = (int)v;
  return n - (v < (double)n);
}
// 63-bit key of a point for cell size 1/inv (downsampling.hpp:41-55); KEY_EMPTY if out of range
__device__ __forceinline__ unsigned long long point_key(double x, double y, double z, double inv) {
  const long long off = 1 << 20, mask = (1 << 21) - 1;
  const long long cx = (long long)fast_floor_d(x * inv) + off, cy = (long long)fast_floor_d(y * inv) + off,
                  cz = (long long)fast_floor_d(z * inv) + off;
  if (cx < 0 || cy < 0 || cz < 0 || cx > mask || cy > mask || cz > mask) return KEY_EMPTY;
  return (unsigned long long)cx | ((unsigned long long)cy << 21) | ((unsigned long long)cz << 42);
}

__global__ void k_group_clear(GicpDev D, int clouds) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)clouds * D.hsize;
  if (i < n) {
    const size_t j = (size_t)cloud_of(D, (int)(i / D.hsize)) * D.hsize + (size_t)(i % D.hsize);
    D.keys[j] = KEY_EMPTY;
    D.minIdx[j] = 0x7fffffff;
    D.count[j] = 0;
    D.cursor[j] = 0;
  }
  if (i < (long long)clouds * 6) D.cellBox[cloud_of(D, (int)(i / 6)) * 6 + (i % 6)] = ((i % 6) < 3) ? 0x7fffffff : -0x7fffffff;
}

// mode 0: keys of the raw float4 input points, voxel leaf; mode 1: keys of the downsampled points, grid cell
__global__ void __launch_bounds__(256) k_group_insert(GicpDev D, int mode, const float* __restrict__ tgt,
                                                      const float* __restrict__ src, int stride) {
  const int c = cloud_of(D, blockIdx.y), i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = mode == 0 ? D.nIn[c] : D.nDown[c];
  if (i >= n) return;
  double x, y, z, inv;
  if (mode == 0) {
    const float* p = raw_points(D, c, tgt, src, stride) + (size_t)i * 4;
    const float4 v = *reinterpret_cast<const float4*>(p);
    x = (double)v.x; y = (double)v.y; z = (double)v.z;
    inv = 1.0 / D.voxel;
  } else {
    const double* p = D.pts + ((size_t)c * D.nmax + i) * 4;
    x = p[0]; y = p[1]; z = p[2];
    inv = 1.0 / D.cell;
  }
  const unsigned long long key = point_key(x, y, z, inv);
  int slot = -1;
  if (key != KEY_EMPTY) {
    const int hm = D.hsize - 1;
    unsigned long long* keys = D.keys + (size_t)c * D.hsize;
    int h = (int)(mix64(key) & hm);
    while (true) {
      const unsigned long long prev = atomicCAS(&keys[h], KEY_EMPTY, key);
      if (prev == KEY_EMPTY || prev == key) break;
      h = (h + 1) & hm;
    }
    slot = h;
    atomicMin(&D.minIdx[(size_t)c * D.hsize + h], i);
    atomicAdd(&D.count[(size_t)c * D.hsize + h], 1);
    if (mode == 1) {
      int* box = D.cellBox + c * 6;
      const int cx = (int)(key & 0x1fffff), cy = (int)((key >> 21) & 0x1fffff), cz = (int)((key >> 42) & 0x1fffff);
      atomicMin(&box[0], cx); atomicMin(&box[1], cy); atomicMin(&box[2], cz);
      atomicMax(&box[3], cx); atomicMax(&box[4], cy); atomicMax(&box[5], cz);
    }
  }
  D.slotOf[(size_t)c * D.nmax + i] = slot;
}

// One CTA per cloud: number the groups by first occurrence and lay their member lists out
// contiguously (start = exclusive prefix of the group sizes in that order).
__global__ void __launch_bounds__(1024) k_group_rank(GicpDev D, int mode, int* __restrict__ nGroups) {
  __shared__ int s_a[32], s_b[32];
  __shared__ int s_ca, s_cb;
  const int c = cloud_of(D, blockIdx.x), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = mode == 0 ? D.nIn[c] : D.nDown[c];
  const int* slotOf = D.slotOf + (size_t)c * D.nmax;
  const size_t hb = (size_t)c * D.hsize;
  if (tid == 0) { s_ca = 0; s_cb = 0; }
  __syncthreads();
  for (int base = 0; base < n; base += 4096) {   // four consecutive points per thread and trip
    int lead[4], cnt[4], slot[4];
    int a = 0, b = 0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = base + 4 * tid + u;
      lead[u] = 0; cnt[u] = 0; slot[u] = -1;
      if (i < n) {
        slot[u] = slotOf[i];
        if (slot[u] >= 0 && D.minIdx[hb + slot[u]] == i) { lead[u] = 1; cnt[u] = D.count[hb + slot[u]]; }
      }
      a += lead[u]; b += cnt[u];
    }
    const int ownA = a, ownB = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int ta = __shfl_up_sync(0xffffffffu, a, o), tb = __shfl_up_sync(0xffffffffu, b, o);
      if (lane >= o) { a += ta; b += tb; }
    }
    if (lane == 31) { s_a[warp] = a; s_b[warp] = b; }
    __syncthreads();
    if (warp == 0) {
      int wa = s_a[lane], wb = s_b[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, wa, o), tb = __shfl_up_sync(0xffffffffu, wb, o);
        if (lane >= o) { wa += ta; wb += tb; }
      }
      s_a[lane] = wa; s_b[lane] = wb;
    }
    __syncthreads();
    int pa = s_ca + (warp ? s_a[warp - 1] : 0) + a - ownA;
    int pb = s_cb + (warp ? s_b[warp - 1] : 0) + b - ownB;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (lead[u]) { D.rank[hb + slot[u]] = pa; D.start[hb + slot[u]] = pb; }
      pa += lead[u]; pb += cnt[u];
    }
    __syncthreads();
    if (tid == 0) { s_ca += s_a[31]; s_cb += s_b[31]; }
    __syncthreads();
  }
  if (tid == 0) nGroups[c] = s_ca;
}

// Append every point to its group's member list.  The order inside a group is whatever the atomics give:
// the voxel mean sorts its (short) list back into input order, the k-NN grid does not depend on it.
__global__ void __launch_bounds__(256) k_group_fill(GicpDev D, int mode) {
  const int c = cloud_of(D, blockIdx.y), i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = mode == 0 ? D.nIn[c] : D.nDown[c];
  if (i >= n) return;
  const int slot = D.slotOf[(size_t)c * D.nmax + i];
  if (slot < 0) return;
  const size_t hb = (size_t)c * D.hsize;
  const int pos = atomicAdd(&D.cursor[hb + slot], 1);
  D.members[(size_t)c * D.nmax + D.start[hb + slot] + pos] = i;
}

// ascending in-place sort of one member list by its owning thread: insertion sort for the usual handful
// of points per voxel, heap sort beyond that
__device__ void sort_members(int* m, int n) {
  if (n <= 24) {
    for (int i = 1; i < n; i++) {
      const int v = m[i];
      int j = i - 1;
      while (j >= 0 && m[j] > v) { m[j + 1] = m[j]; j--; }
      m[j + 1] = v;
    }
    return;
  }
  auto sift = [&](int root, int end) {
    const int v = m[root];
    while (true) {
      int ch = 2 * root + 1;
      if (ch >= end) break;
      if (ch + 1 < end && m[ch + 1] > m[ch]) ch++;
      if (m[ch] <= v) break;
      m[root] = m[ch];
      root = ch;
    }
    m[root] = v;
  };
  for (int i = n / 2 - 1; i >= 0; i--) sift(i, n);
  for (int e = n - 1; e > 0; e--) {
    const int t = m[0]; m[0] = m[e]; m[e] = t;
    sift(0, e);
  }
}

// voxelgrid_sampling: mean of each voxel's points, summed in input order (downsampling.hpp:60-75)
__global__ void __launch_bounds__(256) k_voxel_mean(GicpDev D, const float* __restrict__ tgt, const float* __restrict__ src,
                                                    int stride) {
  const int c = cloud_of(D, blockIdx.y), i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.nIn[c]) return;
  const int slot = D.slotOf[(size_t)c * D.nmax + i];
  if (slot < 0) return;
  const size_t hb = (size_t)c * D.hsize;
  if (D.minIdx[hb + slot] != i) return;
  const float* base = raw_points(D, c, tgt, src, stride);
  int* mem = D.members + (size_t)c * D.nmax + D.start[hb + slot];
  const int cnt = D.count[hb + slot];
  sort_members(mem, cnt);
  double sx = 0, sy = 0, sz = 0, w = 0;
  for (int j = 0; j < cnt; j++) {
    const float4 v = *reinterpret_cast<const float4*>(base + (size_t)mem[j] * 4);
    sx += (double)v.x; sy += (double)v.y; sz += (double)v.z; w += 1.0;
  }
  double* o = D.pts + ((size_t)c * D.nmax + D.rank[hb + slot]) * 4;
  o[0] = sx / w; o[1] = sy / w; o[2] = sz / w; o[3] = 1.0;
}

// ---- octant order inside a grid cell.  One thread per hash slot: the cell's member list is counting-sorted by the octant of
// the cell each point falls in (decided in fp64 from the cell's origin, the expression every search recomputes), written to
// slotOf[] (free between the grouping and the searches) where k_cell_pack picks it up, and the eight counts are packed into
// oct[].  A search that knows how far it has to look visits only the octants its ball reaches (k_nn_corr2), and the 10-NN
// kernel skips whole octants that cannot hold a neighbour (k_knn_cov_warp): a 0.1 m cell holds ~25 points, an octant ~3-6.
__global__ void __launch_bounds__(256) k_cell_sort(GicpDev D, int clouds) {
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= (long long)clouds * D.hsize) return;
  const int c = cloud_of(D, (int)(i0 / D.hsize));
  const size_t i = (size_t)c * D.hsize + (size_t)(i0 % D.hsize);
  const unsigned long long key = D.keys[i];
  if (key == KEY_EMPTY) { D.oct[i] = make_uint2(0u, 0u); return; }
  const int cnt = D.count[i], st = D.start[i];
  const int* mem = D.members + (size_t)c * D.nmax + st;
  int* out = D.slotOf + (size_t)c * D.nmax + st;
  if (cnt > 255) {
    for (int j = 0; j < cnt; j++) out[j] = mem[j];
    D.oct[i] = make_uint2(0xffffffffu, 0xffffffffu);
    return;
  }
  const int off = 1 << 20;
  const double ox = (double)((int)(key & 0x1fffff) - off) * D.cell, oy = (double)((int)((key >> 21) & 0x1fffff) - off) * D.cell,
               oz = (double)((int)((key >> 42) & 0x1fffff) - off) * D.cell;
  const double half = 0.5 * D.cell;
  const double* pts = D.pts + (size_t)c * D.nmax * 4;
  auto octant = [&](int pi) -> int {
    const double* p = pts + (size_t)pi * 4;
    return ((p[0] - ox >= half) ? 1 : 0) | ((p[1] - oy >= half) ? 2 : 0) | ((p[2] - oz >= half) ? 4 : 0);
  };
  unsigned long long n8 = 0;  // eight byte counters
  for (int j = 0; j < cnt; j++) n8 += 1ull << (8 * octant(mem[j]));
  unsigned long long cur = n8 * 0x0101010101010100ull;  // byte k = sum of the bytes below k (no carries: total <= 255)
  for (int j = 0; j < cnt; j++) {
    const int pi = mem[j], o = octant(pi);
    out[(int)((cur >> (8 * o)) & 0xffull)] = pi;
    cur += 1ull << (8 * o);
  }
  D.oct[i] = make_uint2((unsigned)n8, (unsigned)(n8 >> 32));
}

// ---- k-NN grid: packed slot table + points stored in cell order
__global__ void __launch_bounds__(256) k_cell_pack(GicpDev D, int clouds) {
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 < (long long)clouds * D.hsize) {
    const size_t cb = (size_t)cloud_of(D, (int)(i0 / D.hsize)) * D.hsize, i = cb + (size_t)(i0 % D.hsize);
    const unsigned long long k = D.keys[i];
    D.tab[i] = make_uint4((unsigned)k, (unsigned)(k >> 32), (unsigned)D.start[i], (unsigned)D.count[i]);
    // occupied cells in first-occurrence order -> slot (the fill cursors are dead by now: reuse their storage)
    if (k != KEY_EMPTY) D.cursor[cb + D.rank[i]] = (int)(i0 % D.hsize);
  }
  if (i0 < clouds) D.nFall[cloud_of(D, (int)i0)] = 0;
  if (i0 < (long long)clouds * D.nmax) {
    const int c = cloud_of(D, (int)(i0 / D.nmax)), m = (int)(i0 % D.nmax);
    const size_t i = (size_t)c * D.nmax + m;
    if (m < D.nDown[c]) {
      const int pi = D.octSorted ? D.slotOf[i] : D.members[i];   // octant order (k_cell_sort) only when an octant kernel needs it
      const double2* p = reinterpret_cast<const double2*>(D.pts + ((size_t)c * D.nmax + pi) * 4);
      double2* o = reinterpret_cast<double2*>(D.rec + (size_t)i * 4);
      const double2 xy = p[0];
      o[0] = xy;
      o[1] = make_double2(p[1].x, __longlong_as_double((long long)pi));
      // float32 copy relative to the origin of the record's own cell (|.| < cell: 6e-9 m of rounding at cell = 0.1 m);
      // the same floor as point_key, so the origin is the one a search computes from the cell it visits
      const double inv = 1.0 / D.cell;
      const double ox = (double)fast_floor_d(xy.x * inv) * D.cell, oy = (double)fast_floor_d(xy.y * inv) * D.cell,
                   oz = (double)fast_floor_d(p[1].x * inv) * D.cell;
      D.recf[i] = make_float4((float)(xy.x - ox), (float)(xy.y - oy), (float)(p[1].x - oz), __int_as_float(pi));
    }
  }
}

struct Grid {
  const uint4* tab;
  const double2* rec;
  const float4* recf;
  const uint2* oct;
  const double* pts;
  int hm;
};
__device__ __forceinline__ Grid make_grid(const GicpDev& D, int c) {
  Grid g;
  g.tab = D.tab + (size_t)c * D.hsize;
  g.rec = reinterpret_cast<const double2*>(D.rec + (size_t)c * D.nmax * 4);
  g.recf = D.recf + (size_t)c * D.nmax;
  g.oct = D.oct + (size_t)c * D.hsize;
  g.pts = D.pts + (size_t)c * D.nmax * 4;
  g.hm = D.hsize - 1;
  return g;
}
// occupied cell -> (first record, count); false for an empty cell
__device__ __forceinline__ bool grid_find(const Grid& g, int cx, int cy, int cz, int& start, int& count) {
  if ((unsigned)cx > 0x1fffffu || (unsigned)cy > 0x1fffffu || (unsigned)cz > 0x1fffffu) return false;
  const unsigned long long key = (unsigned long long)cx | ((unsigned long long)cy << 21) | ((unsigned long long)cz << 42);
  int h = (int)(mix64(key) & g.hm);
  while (true) {
    const uint4 e = __ldg(&g.tab[h]);
    const unsigned long long k = (unsigned long long)e.x | ((unsigned long long)e.y << 32);
    if (k == key) { start = (int)e.z; count = (int)e.w; return true; }
    if (k == KEY_EMPTY) return false;
    h = (h + 1) & g.hm;
  }
}
// the same, also returning the hash slot (index into oct[])
__device__ __forceinline__ bool grid_find_slot(const Grid& g, int cx, int cy, int cz, int& start, int& count, int& slot) {
  if ((unsigned)cx > 0x1fffffu || (unsigned)cy > 0x1fffffu || (unsigned)cz > 0x1fffffu) return false;
  const unsigned long long key = (unsigned long long)cx | ((unsigned long long)cy << 21) | ((unsigned long long)cz << 42);
  int h = (int)(mix64(key) & g.hm);
  while (true) {
    const uint4 e = __ldg(&g.tab[h]);
    const unsigned long long k = (unsigned long long)e.x | ((unsigned long long)e.y << 32);
    if (k == key) { start = (int)e.z; count = (int)e.w; slot = h; return true; }
    if (k == KEY_EMPTY) return false;
    h = (h + 1) & g.hm;
  }
}
// record -> squared distance to the query and the point index
template <class G>
__device__ __forceinline__ double rec_sqdist(const G& g, int r, double qx, double qy, double qz, int& index) {
  const double2 a = __ldg(&g.rec[2 * (size_t)r]), b = __ldg(&g.rec[2 * (size_t)r + 1]);
  index = (int)__double_as_longlong(b.y);
  // Eigen Vector4d::squaredNorm with SSE2 packets: (dx^2 + dz^2) + dy^2
  const double dx = a.x - qx, dy = a.y - qy, dz = b.x - qz;
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dz, dz)), __dmul_rn(dy, dy));
}
__device__ __forceinline__ double sqdist3(const double* p, double qx, double qy, double qz) {
  // Eigen Vector4d::squaredNorm with SSE2 packets: (dx^2 + dz^2) + dy^2
  const double dx = p[0] - qx, dy = p[1] - qy, dz = p[2] - qz;
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dz, dz)), __dmul_rn(dy, dy));
}

// k nearest under the total order (distance, index): independent of the visiting order
template <int K>
struct KnnAcc {
  double d[K];
  int id[K];
  int found;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < K; i++) { d[i] = DBL_MAX; id[i] = 0x7fffffff; }
    found = 0;
  }
  __device__ __forceinline__ void push(int index, double dist) {
    if (dist > d[K - 1] || (dist == d[K - 1] && index >= id[K - 1])) return;
    double cd = dist;
    int ci = index;
#pragma unroll
    for (int i = 0; i < K; i++) {
      const bool before = cd < d[i] || (cd == d[i] && ci < id[i]);
      if (before) {
        const double td = d[i]; const int ti = id[i];
        d[i] = cd; id[i] = ci;
        cd = td; ci = ti;
      }
    }
    if (found < K) found++;
  }
};

template <int K, class G>
__device__ __forceinline__ void scan_cell(const G& g, int s, int n, double qx, double qy, double qz, KnnAcc<K>& acc) {
  for (int j = 0; j < n; j++) {
    int pi;
    const double d = rec_sqdist(g, s + j, qx, qy, qz, pi);
    acc.push(pi, d);
  }
}

// ---- shell traversal with cell pruning
// Query geometry: home cell (cx,cy,cz) and the offsets (fx,fy,fz) of the query inside it.
struct ShellQuery {
  int cx, cy, cz;
  double fx, fy, fz, cell, margin;
};
__device__ __forceinline__ ShellQuery make_shell_query(double cell, double qx, double qy, double qz) {
  ShellQuery s;
  const double inv = 1.0 / cell;
  const int off = 1 << 20;
  s.cx = fast_floor_d(qx * inv) + off; s.cy = fast_floor_d(qy * inv) + off; s.cz = fast_floor_d(qz * inv) + off;
  s.fx = qx - (double)(s.cx - off) * cell; s.fy = qy - (double)(s.cy - off) * cell; s.fz = qz - (double)(s.cz - off) * cell;
  s.cell = cell;
  // distance from the query to the nearest face of its own cell
  const double m = fmin(fmin(s.fx, cell - s.fx), fmin(fmin(s.fy, cell - s.fy), fmin(s.fz, cell - s.fz)));
  s.margin = fmax(m, 0.0) * 0.999999;  // guard the bound against rounding in fx..fz
  return s;
}
// squared lower bound of the distance from the query to any point of a cell `d` cells away along one axis
// (shrunk a little so that rounding in the cell assignment can never prune a cell that matters)
__device__ __forceinline__ double axis_gap2(int d, double f, double cell) {
  if (d == 0) return 0.0;
  const double g = d > 0 ? (double)d * cell - f : f - (double)(d + 1) * cell;
  const double gg = fmax(g * 0.999999 - 1e-12, 0.0);
  return gg * gg;
}
// Visit the occupied cells of shell r (Chebyshev ring) whose lower bound does not exceed limit().
// cls > 0 (only meaningful for r == 1): visit only the cells with exactly `cls` non-zero offsets -- 1 = the six face
// neighbours, 2 = the twelve edge neighbours, 3 = the eight corners -- so the caller can tighten limit() in between.
template <class Limit, class Visit>
__device__ __forceinline__ void visit_shell(const Grid& g, const int* box, const ShellQuery& q, int r, Limit limit, Visit visit, int cls = 0) {
  const int x0 = max(q.cx - r, box[0]), x1 = min(q.cx + r, box[3]);
  const int y0 = max(q.cy - r, box[1]), y1 = min(q.cy + r, box[4]);
  const int z0 = max(q.cz - r, box[2]), z1 = min(q.cz + r, box[5]);
  for (int z = z0; z <= z1; z++) {
    const double gz2 = axis_gap2(z - q.cz, q.fz, q.cell);
    if (gz2 > limit()) continue;
    for (int y = y0; y <= y1; y++) {
      const double gyz2 = gz2 + axis_gap2(y - q.cy, q.fy, q.cell);
      if (gyz2 > limit()) continue;
      const bool face = (z == q.cz - r) || (z == q.cz + r) || (y == q.cy - r) || (y == q.cy + r);
      const int nzy = (z != q.cz) + (y != q.cy);
      if (face) {
        for (int x = x0; x <= x1; x++) {
          if (cls && nzy + (x != q.cx) != cls) continue;
          if (gyz2 + axis_gap2(x - q.cx, q.fx, q.cell) > limit()) continue;
          int cs, cn;
          if (grid_find(g, x, y, z, cs, cn)) visit(cs, cn);
        }
      } else {
        if (cls && nzy + 1 != cls) continue;
        if (q.cx - r >= x0 && !(gyz2 + axis_gap2(-r, q.fx, q.cell) > limit())) {
          int cs, cn;
          if (grid_find(g, q.cx - r, y, z, cs, cn)) visit(cs, cn);
        }
        if (r > 0 && q.cx + r <= x1 && !(gyz2 + axis_gap2(r, q.fx, q.cell) > limit())) {
          int cs, cn;
          if (grid_find(g, q.cx + r, y, z, cs, cn)) visit(cs, cn);
        }
      }
    }
  }
}
__device__ __forceinline__ bool box_covered(const int* box, const ShellQuery& q, int r) {
  return q.cx - r <= box[0] && q.cx + r >= box[3] && q.cy - r <= box[1] && q.cy + r >= box[4] && q.cz - r <= box[2] &&
         q.cz + r >= box[5];
}

// ---- dense sorted grid (GFS_GICP_GRID=1, the default for the default kernels).
// The hash grid costs a query 27 dependent probes (mix64, a 16-byte slot from a 1 MB table, key compare) for the 27 cells around
// it, about 60 % of them empty, and a cell's records lie wherever the cell was first seen.  Depth-camera clouds fit a box of a
// few hundred thousand 0.11 m cells, so the grid can simply be an array over that box: S[cell] = index of the cell's first
// record, cells numbered x-fastest and the records stored in that order (S is the exclusive prefix sum of the cell
// populations, k_dense_*).  A lookup is one 4-byte load, and -- the point of sorting -- the cells xa..xb of one row are ONE
// contiguous run of records S[xa] .. S[xb + 1]: the 27-cell neighbourhood is 9 runs, found with 18 loads from 9 sectors.
// A cloud whose box exceeds the array (outliers far away) keeps a region of at most `daxis` cells per axis around its mean cell;
// points AND queries outside are clamped into the region's border cells.  That keeps every bound valid: a border cell then
// also holds points beyond its outer face, which are farther from any query inside the region than the face the bounds are
// computed from, and a clamped query sits in the border cell with an offset outside [0, cell) -- its margin is 0 and the gaps to
// the inner cells are its true distances to their faces.  (Any cell assignment gives the same exact result; only speed differs.)
struct DGrid {
  const unsigned* S;
  const double2* rec;
  const double* pts;
  int ox, oy, oz, nx, ny, nz;
};
__device__ __forceinline__ DGrid make_dgrid(const GicpDev& D, int c) {
  DGrid g;
  g.S = D.dS + (size_t)c * (size_t)D.dstride;
  g.rec = reinterpret_cast<const double2*>(D.rec + (size_t)c * D.nmax * 4);
  g.pts = D.pts + (size_t)c * D.nmax * 4;
  const int* box = D.cellBox + c * 6;
  g.ox = box[0]; g.oy = box[1]; g.oz = box[2];
  g.nx = box[3] - box[0] + 1; g.ny = box[4] - box[1] + 1; g.nz = box[5] - box[2] + 1;
  return g;
}
// cell coordinate of a point along one axis, clamped into [lo, hi] (saturating: any finite or infinite input is fine)
__device__ __forceinline__ int dense_coord(double v, double inv, int lo, int hi) {
  const double t = fmin(fmax(v * inv, -2.0e6), 2.0e6);
  const int c = fast_floor_d(t) + (1 << 20);
  return min(max(c, lo), hi);
}
__device__ __forceinline__ ShellQuery make_shell_query(const DGrid& g, double cell, double qx, double qy, double qz) {
  ShellQuery s;
  const double inv = 1.0 / cell;
  const int off = 1 << 20;
  s.cx = dense_coord(qx, inv, g.ox, g.ox + g.nx - 1); s.cy = dense_coord(qy, inv, g.oy, g.oy + g.ny - 1);
  s.cz = dense_coord(qz, inv, g.oz, g.oz + g.nz - 1);
  s.fx = qx - (double)(s.cx - off) * cell; s.fy = qy - (double)(s.cy - off) * cell; s.fz = qz - (double)(s.cz - off) * cell;
  s.cell = cell;
  const double m = fmin(fmin(s.fx, cell - s.fx), fmin(fmin(s.fy, cell - s.fy), fmin(s.fz, cell - s.fz)));
  s.margin = fmax(m, 0.0) * 0.999999;
  return s;
}
__device__ __forceinline__ ShellQuery make_shell_query(const Grid&, double cell, double qx, double qy, double qz) {
  return make_shell_query(cell, qx, qy, qz);
}
// records of the cells xa..xb (inside the region) of row (y, z)
__device__ __forceinline__ void dense_run(const DGrid& g, int y, int z, int xa, int xb, int& start, int& count) {
  const unsigned* row = g.S + ((size_t)(z - g.oz) * g.ny + (size_t)(y - g.oy)) * g.nx;
  const unsigned a = __ldg(row + (xa - g.ox)), b = __ldg(row + (xb - g.ox + 1));
  start = (int)a; count = (int)(b - a);
}
// visit_shell for the dense grid: visit(first record, count) once per ROW of the shell.  rows > 0 (only meaningful for r == 1):
// only the rows with exactly rows - 1 non-zero offsets in (y, z) -- 1 = the query's own row (its two end cells), 2 = the four
// rows next to it, 3 = the four diagonal rows -- so that the caller can tighten limit() in between.
template <class Limit, class Visit>
__device__ __forceinline__ void visit_shell(const DGrid& g, const int*, const ShellQuery& q, int r, Limit limit, Visit visit, int rows = 0) {
  const int x0 = max(q.cx - r, g.ox), x1 = min(q.cx + r, g.ox + g.nx - 1);
  const int y0 = max(q.cy - r, g.oy), y1 = min(q.cy + r, g.oy + g.ny - 1);
  const int z0 = max(q.cz - r, g.oz), z1 = min(q.cz + r, g.oz + g.nz - 1);
  for (int z = z0; z <= z1; z++) {
    const double gz2 = axis_gap2(z - q.cz, q.fz, q.cell);
    if (gz2 > limit()) continue;
    for (int y = y0; y <= y1; y++) {
      if (rows && (z != q.cz) + (y != q.cy) != rows - 1) continue;
      const double gyz2 = gz2 + axis_gap2(y - q.cy, q.fy, q.cell);
      if (gyz2 > limit()) continue;
      const bool face = (z == q.cz - r) || (z == q.cz + r) || (y == q.cy - r) || (y == q.cy + r);
      int cs, cn;
      if (face) {
        // the gap grows with |x - cx|: the cells that survive pruning are one run around cx
        int xa = x0, xb = x1;
        while (xa < q.cx && gyz2 + axis_gap2(xa - q.cx, q.fx, q.cell) > limit()) xa++;
        while (xb > q.cx && gyz2 + axis_gap2(xb - q.cx, q.fx, q.cell) > limit()) xb--;
        dense_run(g, y, z, xa, xb, cs, cn);
        if (cn) visit(cs, cn);
      } else {
        if (q.cx - r >= x0 && !(gyz2 + axis_gap2(-r, q.fx, q.cell) > limit())) {
          dense_run(g, y, z, q.cx - r, q.cx - r, cs, cn);
          if (cn) visit(cs, cn);
        }
        if (r > 0 && q.cx + r <= x1 && !(gyz2 + axis_gap2(r, q.fx, q.cell) > limit())) {
          dense_run(g, y, z, q.cx + r, q.cx + r, cs, cn);
          if (cn) visit(cs, cn);
        }
      }
    }
  }
}
__device__ __forceinline__ bool box_covered(const DGrid& g, const int*, const ShellQuery& q, int r) {
  return q.cx - r <= g.ox && q.cx + r >= g.ox + g.nx - 1 && q.cy - r <= g.oy && q.cy + r >= g.oy + g.ny - 1 && q.cz - r <= g.oz &&
         q.cz + r >= g.oz + g.nz - 1;
}
__device__ __forceinline__ bool box_covered(const Grid&, const int* box, const ShellQuery& q, int r) { return box_covered(box, q, r); }

// ---- dense sorted grid: construction (replaces the hash grouping of the downsampled points + k_cell_pack)
__global__ void k_dense_init(GicpDev D, int clouds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= clouds) return;
  const int c = cloud_of(D, i);
  for (int k = 0; k < 3; k++) { D.cellBox[c * 6 + k] = 0x7fffffff; D.cellBox[c * 6 + 3 + k] = -0x7fffffff; D.dSum[c * 3 + k] = 0ull; }
  D.nFall[c] = 0;
}
// bounding box (in cells) and coordinate sums of every cloud
__global__ void __launch_bounds__(256) k_dense_box(GicpDev D) {
  const int c = cloud_of(D, blockIdx.y);
  const int n = D.nDown[c];
  if ((int)(blockIdx.x * blockDim.x * 4) >= n) return;   // 4 points per thread, 256 apart
  int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-0x7fffffff, -0x7fffffff, -0x7fffffff};
  unsigned long long sum[3] = {0ull, 0ull, 0ull};
  const double inv = 1.0 / D.cell;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int i = (blockIdx.x * 4 + j) * blockDim.x + threadIdx.x;
    if (i >= n) break;
    const double* p = D.pts + ((size_t)c * D.nmax + i) * 4;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int v = dense_coord(p[k], inv, 0, (1 << 21) - 1);
      lo[k] = min(lo[k], v); hi[k] = max(hi[k], v); sum[k] += (unsigned long long)v;
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = min(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = max(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
      sum[k] += __shfl_xor_sync(0xffffffffu, sum[k], o);
    }
  }
  // warps -> CTA in shared memory, then nine atomics per CTA (per warp they queued up on the same nine addresses of a cloud:
  // 240 us for 64 clouds)
  __shared__ int s_lo[8][3], s_hi[8][3];
  __shared__ unsigned long long s_sum[8][3];
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) { s_lo[warp][k] = lo[k]; s_hi[warp][k] = hi[k]; s_sum[warp][k] = sum[k]; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int k = threadIdx.x;
    int l = s_lo[0][k], h = s_hi[0][k];
    unsigned long long t = s_sum[0][k];
    for (int w = 1; w < 8; w++) { l = min(l, s_lo[w][k]); h = max(h, s_hi[w][k]); t += s_sum[w][k]; }
    atomicMin(&D.cellBox[c * 6 + k], l);
    atomicMax(&D.cellBox[c * 6 + 3 + k], h);
    atomicAdd(&D.dSum[c * 3 + k], t);
  }
}
// one CTA per cloud: fix the region (the box, or `daxis` cells per oversized axis around the mean cell) and clear its counters
__global__ void __launch_bounds__(1024) k_dense_region(GicpDev D) {
  __shared__ long long s_vol;
  const int c = cloud_of(D, blockIdx.x);
  if (threadIdx.x == 0) {
    int* box = D.cellBox + c * 6;
    const int n = D.nDown[c];
    int lo[3], hi[3];
    for (int k = 0; k < 3; k++) { lo[k] = n > 0 ? box[k] : 0; hi[k] = n > 0 ? box[3 + k] : 0; }
    long long vol = 1;
    for (int k = 0; k < 3; k++) vol *= (long long)(hi[k] - lo[k] + 1);
    if (vol > (long long)D.dcap) {
      for (int k = 0; k < 3; k++) {
        if (hi[k] - lo[k] + 1 <= D.daxis) continue;
        const int mean = (int)(D.dSum[c * 3 + k] / (unsigned long long)n);
        const int l = min(max(mean - D.daxis / 2, lo[k]), hi[k] - D.daxis + 1);
        lo[k] = l; hi[k] = l + D.daxis - 1;
      }
      vol = 1;
      for (int k = 0; k < 3; k++) vol *= (long long)(hi[k] - lo[k] + 1);
    }
    for (int k = 0; k < 3; k++) { box[k] = lo[k]; box[3 + k] = hi[k]; }
    s_vol = vol;
  }
  __syncthreads();
  unsigned* S = D.dS + (size_t)c * (size_t)D.dstride;
  const long long vol = s_vol;
  for (long long j = threadIdx.x; j <= vol; j += blockDim.x) S[j] = 0u;
}
// population of every cell; a point remembers its cell (slotOf) and its arrival number inside it (members)
__global__ void __launch_bounds__(256) k_dense_count(GicpDev D) {
  const int c = cloud_of(D, blockIdx.y), i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.nDown[c]) return;
  const int* box = D.cellBox + c * 6;
  const int ox = box[0], oy = box[1], oz = box[2], nx = box[3] - box[0] + 1, ny = box[4] - box[1] + 1;
  const double* p = D.pts + ((size_t)c * D.nmax + i) * 4;
  const double inv = 1.0 / D.cell;
  const int cx = dense_coord(p[0], inv, ox, box[3]), cy = dense_coord(p[1], inv, oy, box[4]), cz = dense_coord(p[2], inv, oz, box[5]);
  const int idx = ((cz - oz) * ny + (cy - oy)) * nx + (cx - ox);
  unsigned* S = D.dS + (size_t)c * (size_t)D.dstride;
  D.slotOf[(size_t)c * D.nmax + i] = idx;
  D.members[(size_t)c * D.nmax + i] = (int)atomicAdd(&S[idx], 1u);
}
// one CTA per cloud: exclusive prefix sum of the populations, in place, 4096 cells per trip (uint4 per thread); S[vol] = n
__global__ void __launch_bounds__(1024) k_dense_scan(GicpDev D) {
  __shared__ unsigned s_w[32];
  __shared__ unsigned s_carry;
  const int c = cloud_of(D, blockIdx.x);
  const int* box = D.cellBox + c * 6;
  const long long vol = (long long)(box[3] - box[0] + 1) * (box[4] - box[1] + 1) * (box[5] - box[2] + 1);
  unsigned* S = D.dS + (size_t)c * (size_t)D.dstride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0u;
  __syncthreads();
  for (long long base = 0; base < vol; base += 4096) {
    const long long j = base + 4 * (long long)tid;
    unsigned v[4];
    if (j + 3 < vol) {
      const uint4 t = *reinterpret_cast<const uint4*>(S + j);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int u = 0; u < 4; u++) v[u] = (j + u < vol) ? S[j + u] : 0u;
    }
    const unsigned own = v[0] + v[1] + v[2] + v[3];
    unsigned incl = own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned w = s_w[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      s_w[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    unsigned run = s_carry + (warp ? s_w[warp - 1] : 0u) + incl - own;
    if (j + 3 < vol) {
      *reinterpret_cast<uint4*>(S + j) = make_uint4(run, run + v[0], run + v[0] + v[1], run + v[0] + v[1] + v[2]);
    } else {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (j + u < vol) S[j + u] = run;
        run += v[u];
      }
    }
    __syncthreads();
    if (tid == 0) s_carry += s_w[31];
    __syncthreads();
  }
  if (tid == 0) { S[vol] = s_carry; D.nCells[c] = 0; }
}
// the points again, in cell order {x, y, z, index bits}
__global__ void __launch_bounds__(256) k_dense_fill(GicpDev D) {
  const int c = cloud_of(D, blockIdx.y), i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.nDown[c]) return;
  const unsigned* S = D.dS + (size_t)c * (size_t)D.dstride;
  const size_t pos = (size_t)c * D.nmax + S[D.slotOf[(size_t)c * D.nmax + i]] + (unsigned)D.members[(size_t)c * D.nmax + i];
  const double2* p = reinterpret_cast<const double2*>(D.pts + ((size_t)c * D.nmax + i) * 4);
  double2* o = reinterpret_cast<double2*>(D.rec + pos * 4);
  o[0] = p[0];
  o[1] = make_double2(p[1].x, __longlong_as_double((long long)i));
}

// Exact k-NN: shells of grid cells around the query until the k-th distance is provably final,
// brute force over the cloud beyond MAX_SHELL.  If max_r2 >= 0 the search may stop as soon as no
// unvisited point can be closer than sqrt(max_r2) (bounded 1-NN for correspondences).
static const int MAX_SHELL = 8;
template <int K, class G>
__device__ bool grid_knn(const G& g, const int* box, int nPts, double cell, double qx, double qy, double qz,
                         double max_r2, KnnAcc<K>& acc, bool seeded = false) {
  if (!seeded) acc.init();
  const ShellQuery q = make_shell_query(g, cell, qx, qy, qz);
  const double cap = max_r2 >= 0.0 ? max_r2 : DBL_MAX;
  for (int r = 0; r <= MAX_SHELL; r++) {
    visit_shell(g, box, q, r,
                [&]() { return acc.found == K ? fmin(acc.d[K - 1], cap) : cap; },
                [&](int cs, int cn) { scan_cell<K>(g, cs, cn, qx, qy, qz, acc); });
    const double bound = (double)r * cell + q.margin;  // every unvisited point is farther than this
    const double b2 = bound * bound;
    if (acc.found == K && acc.d[K - 1] <= b2) return true;
    if (max_r2 >= 0.0 && b2 >= max_r2) return true;
    if (box_covered(g, box, q, r)) return true;
  }
  return false;  // far / sparse query: the caller finishes it by brute force over the cloud
}

// Exact k-NN of ONE query by the whole warp: every lane scans a 1/32 slice of the cloud (coalesced), the 32
// sorted lists are merged by K rounds of warp-wide (distance, index) minima.  The result is returned in
// `acc` of lane `dst`.  Used for the few sparse / far queries the shell search cannot finish: done by a single
// lane it is 42k dependent iterations, the longest tail of the kernel.
template <int K>
__device__ void warp_brute_knn(const double* pts, int nPts, double qx, double qy, double qz, int dst, KnnAcc<K>& acc) {
  const int lane = threadIdx.x & 31;
  KnnAcc<K> loc;
  loc.init();
  for (int i = lane; i < nPts; i += 32) loc.push(i, sqdist3(pts + (size_t)i * 4, qx, qy, qz));
  int found = 0;
#pragma unroll
  for (int r = 0; r < K; r++) {
    double bd = loc.d[0];
    int bi = loc.id[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    if (loc.id[0] == bi && bi != 0x7fffffff) {  // this lane held the minimum: pop it
#pragma unroll
      for (int k = 0; k + 1 < K; k++) { loc.d[k] = loc.d[k + 1]; loc.id[k] = loc.id[k + 1]; }
      loc.d[K - 1] = DBL_MAX; loc.id[K - 1] = 0x7fffffff;
    }
    if (bi != 0x7fffffff) found = r + 1;
    if (lane == dst) { acc.d[r] = bd; acc.id[r] = bi; }
  }
  if (lane == dst) acc.found = found;
}


// Exact 10-NN in two passes, shaped for SIMT.  Per shell the occupied cells that survive pruning are
// first collected into a per-thread list in shared memory; the candidates of those cells are then
// walked by ONE flat loop (a lane never waits in a per-cell inner loop of a neighbour lane).
// Pass A keeps only the K smallest distances, rounded UP to float, in a branch-free min/max chain and
// yields a radius rho >= the true k-th distance; pass B walks the listed cells again and appends the few
// candidates with d <= rho to a short per-thread list; the exact (distance, index) selection then runs
// once over that list.  Falls back to grid_knn when a list overflows (ties, dense cells, clouds with
// fewer than K points) or the shells run out.
static const int KNN_K = 10;
static const int KNN_LIST = 16;    // candidates kept for the exact selection
static const int KNN_CELLS = 48;   // occupied cells remembered per query, packed (start << 8 | count)
static const int KNN_THREADS = 128;
static const size_t KNN_SMEM = (size_t)KNN_THREADS * (KNN_LIST * 12 + KNN_CELLS * 4);
static const int KS_LIST = 12, KS_CELLS = 28;   // k_knn_search: 256 bytes of lists per query, 32 KB per CTA (up to 7 CTAs per SM)
static const size_t KS_SMEM = (size_t)KNN_THREADS * (KS_LIST * 12 + KS_CELLS * 4);
static const int KD_LIST = 12, KD_CELLS = 16;   // dense grid: a list entry is a ROW run (shells 0-1: at most 11), 208 bytes per query
static const size_t KD_SMEM = (size_t)KNN_THREADS * (KD_LIST * 12 + KD_CELLS * 4);

// One flat loop over the records of cells[from..to), four records per trip: their eight 16-byte loads are issued
// before the first one is used (the loop used to wait out a full L1 / L2 latency per record: 37 % of the kernel's
// stall samples).  body(squared distance, point index) is called in record order.
template <class G, class Body>
__device__ __forceinline__ void walk_cells(const G& g, double qx, double qy, double qz, const unsigned* cells, int from, int to, Body body) {
  int ci = from - 1, r = 0, end = 0;
  auto next = [&]() -> int {
    while (r >= end) {
      if (++ci >= to) return -1;
      const unsigned c = cells[ci * KNN_THREADS];
      r = (int)(c >> 8); end = r + (int)(c & 0xffu);
    }
    return r++;
  };
  for (;;) {
    int idx[4];
    idx[0] = next();
    if (idx[0] < 0) break;
#pragma unroll
    for (int u = 1; u < 4; u++) idx[u] = next();
    double2 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const size_t rr = 2 * (size_t)(idx[u] >= 0 ? idx[u] : idx[0]);
      a[u] = __ldg(&g.rec[rr]); b[u] = __ldg(&g.rec[rr + 1]);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (idx[u] < 0) break;
      // Eigen Vector4d::squaredNorm with SSE2 packets: (dx^2 + dz^2) + dy^2
      const double dx = a[u].x - qx, dy = a[u].y - qy, dz = b[u].x - qz;
      body(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dz, dz)), __dmul_rn(dy, dy)), (int)__double_as_longlong(b[u].y));
    }
  }
}

template <int LIST = KNN_LIST, int CELLS = KNN_CELLS, class G = Grid>
__device__ __forceinline__ bool knn10_two_pass(const G& g, const int* box, int nPts, double cell, double qx, double qy, double qz,
                                               double* s_d, int* s_id, unsigned* s_cells, KnnAcc<KNN_K>& acc) {
  const ShellQuery q = make_shell_query(g, cell, qx, qy, qz);
  float top[KNN_K];
#pragma unroll
  for (int i = 0; i < KNN_K; i++) top[i] = FLT_MAX;
  int seen = 0, nc = 0;
  bool ok = false, overflow = false;
  for (int r = 0; r <= MAX_SHELL && !overflow; r++) {
    // shell 1 goes in three rounds -- face, edge, corner neighbours -- so that the k-th distance found so far prunes the
    // farther cells (with ~8 points per cell the home cell alone rarely holds k: nothing would be pruned otherwise)
    const int rounds = r == 1 ? 3 : 1;
    for (int cls = 1; cls <= rounds; cls++) {
      const int from = nc;
      const double lim = (double)top[KNN_K - 1];
      visit_shell(g, box, q, r, [&]() { return lim; },
                  [&](int cs, int cn) {
                    if (nc < CELLS && cn < 256 && cs < (1 << 24)) s_cells[nc++ * KNN_THREADS] = ((unsigned)cs << 8) | (unsigned)cn;
                    else overflow = true;
                  }, r == 1 ? cls : 0);
      walk_cells(g, qx, qy, qz, s_cells, from, nc, [&](double dd, int) {
        float v = __double2float_ru(dd);
#pragma unroll
        for (int i = 0; i < KNN_K; i++) {
          const float lo = fminf(top[i], v);
          v = fmaxf(top[i], v);
          top[i] = lo;
        }
        seen++;
      });
    }
    const double bound = (double)r * cell + q.margin;
    if ((seen >= KNN_K && (double)top[KNN_K - 1] <= bound * bound) || box_covered(g, box, q, r)) { ok = true; break; }
  }
  ok = ok && !overflow && seen >= KNN_K && top[KNN_K - 1] < FLT_MAX;
  int cnt = 0;
  if (ok) {
    const double rho = (double)top[KNN_K - 1];
    walk_cells(g, qx, qy, qz, s_cells, 0, nc, [&](double dd, int pi) {
      if (dd <= rho) {
        if (cnt < LIST) { s_d[cnt * KNN_THREADS] = dd; s_id[cnt * KNN_THREADS] = pi; }
        cnt++;
      }
    });
    ok = cnt <= LIST;
  }
  if (!ok) return grid_knn<KNN_K, G>(g, box, nPts, cell, qx, qy, qz, -1.0, acc);
  acc.init();
  for (int j = 0; j < cnt; j++) acc.push(s_id[j * KNN_THREADS], s_d[j * KNN_THREADS]);
  return true;
}

// ---- Eigen SelfAdjointEigenSolver<Matrix3d>::computeDirect restated (see oracle/gicp_oracle.cpp)
__device__ void sym3_roots(const double m[3][3], double roots[3]) {
  const double s_inv3 = 1.0 / 3.0, s_sqrt3 = sqrt(3.0);
  const double c0 = m[0][0] * m[1][1] * m[2][2] + 2.0 * m[1][0] * m[2][0] * m[2][1] - m[0][0] * m[2][1] * m[2][1] -
                    m[1][1] * m[2][0] * m[2][0] - m[2][2] * m[1][0] * m[1][0];
  const double c1 = m[0][0] * m[1][1] - m[1][0] * m[1][0] + m[0][0] * m[2][2] - m[2][0] * m[2][0] + m[1][1] * m[2][2] -
                    m[2][1] * m[2][1];
  const double c2 = m[0][0] + m[1][1] + m[2][2];
  const double c2_over_3 = c2 * s_inv3;
  double a_over_3 = (c2 * c2_over_3 - c1) * s_inv3;
  a_over_3 = fmax(a_over_3, 0.0);
  const double half_b = 0.5 * (c0 + c2_over_3 * (2.0 * c2_over_3 * c2_over_3 - c1));
  double q = a_over_3 * a_over_3 * a_over_3 - half_b * half_b;
  q = fmax(q, 0.0);
  const double rho = sqrt(a_over_3);
  const double theta = atan2(sqrt(q), half_b) * s_inv3;
  double st, ct;
  sincos(theta, &st, &ct);
  roots[0] = c2_over_3 - rho * (ct + s_sqrt3 * st);
  roots[1] = c2_over_3 - rho * (ct - s_sqrt3 * st);
  roots[2] = c2_over_3 + 2.0 * rho * ct;
}
__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ void extract_kernel(const double m[3][3], double res[3], double rep[3]) {
  int i0 = 0;
  double best = fabs(m[0][0]);
  if (fabs(m[1][1]) > best) { best = fabs(m[1][1]); i0 = 1; }
  if (fabs(m[2][2]) > best) { best = fabs(m[2][2]); i0 = 2; }
  const int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
  double colA[3], colB[3];
  for (int r = 0; r < 3; r++) { rep[r] = m[r][i0]; colA[r] = m[r][i1]; colB[r] = m[r][i2]; }
  double c0[3], c1[3];
  cross3(rep, colA, c0);
  cross3(rep, colB, c1);
  const double n0 = c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2];
  const double n1 = c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2];
  if (n0 > n1) { const double s = sqrt(n0); for (int r = 0; r < 3; r++) res[r] = c0[r] / s; }
  else { const double s = sqrt(n1); for (int r = 0; r < 3; r++) res[r] = c1[r] / s; }
}
__device__ void eig3_direct(const double A[3][3], double V[3][3]) {
  const double eps = DBL_EPSILON;
  double m[3][3], evals[3];
  const double shift = (A[0][0] + A[1][1] + A[2][2]) / 3.0;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) m[r][c] = (r >= c) ? A[r][c] : A[c][r];
  for (int i = 0; i < 3; i++) m[i][i] -= shift;
  double scale = 0;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) scale = fmax(scale, fabs(m[r][c]));
  if (scale > 0)
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) m[r][c] /= scale;
  sym3_roots(m, evals);
  double v[3][3];
  if ((evals[2] - evals[0]) <= eps) {
    for (int k = 0; k < 3; k++)
      for (int r = 0; r < 3; r++) v[k][r] = (k == r) ? 1.0 : 0.0;
  } else {
    double tmp[3][3];
    double d0 = evals[2] - evals[1], d1 = evals[1] - evals[0];
    int k = 0, l = 2;
    if (d0 > d1) { k = 2; l = 0; d0 = d1; }
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) tmp[r][c] = m[r][c];
    for (int i = 0; i < 3; i++) tmp[i][i] -= evals[k];
    extract_kernel(tmp, v[k], v[l]);
    if (d0 <= 2 * eps * d1) {
      const double dot = v[k][0] * v[l][0] + v[k][1] * v[l][1] + v[k][2] * v[l][2];
      for (int r = 0; r < 3; r++) v[l][r] -= dot * v[l][r];
      const double nn = sqrt(v[l][0] * v[l][0] + v[l][1] * v[l][1] + v[l][2] * v[l][2]);
      for (int r = 0; r < 3; r++) v[l][r] /= nn;
    } else {
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) tmp[r][c] = m[r][c];
      for (int i = 0; i < 3; i++) tmp[i][i] -= evals[l];
      double dummy[3];
      extract_kernel(tmp, v[l], dummy);
    }
    double c[3];
    cross3(v[2], v[0], c);
    const double nn = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    for (int r = 0; r < 3; r++) v[1][r] = c[r] / nn;
  }
  for (int k = 0; k < 3; k++)
    for (int r = 0; r < 3; r++) V[r][k] = v[k][r];
}

// CovarianceSetter (normal_estimation.hpp:66-92): regularised covariance of the k neighbours, in (distance, index) order
__device__ void cov_from_knn(const Grid& g, const KnnAcc<KNN_K>& acc, double* out) {
  const int nf = acc.found;
  if (nf < 5) {
    out[0] = 1; out[1] = 0; out[2] = 0; out[3] = 1; out[4] = 0; out[5] = 1;
    return;
  }
  double s[3] = {0, 0, 0}, cr[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int j = 0; j < nf; j++) {
    const double* p = g.pts + (size_t)acc.id[j] * 4;
    const double v[3] = {p[0], p[1], p[2]};
    for (int a = 0; a < 3; a++) {
      s[a] += v[a];
      for (int b = 0; b < 3; b++) cr[a][b] += v[a] * v[b];
    }
  }
  double C[3][3], V[3][3];
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) C[a][b] = (cr[a][b] - (s[a] / nf) * s[b]) / nf;
  eig3_direct(C, V);
  const double val[3] = {1e-3, 1.0, 1.0};
  double R[3][3];
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) {
      double t = 0;
      for (int k = 0; k < 3; k++) t += (V[a][k] * val[k]) * V[b][k];
      R[a][b] = t;
    }
  out[0] = R[0][0]; out[1] = R[0][1]; out[2] = R[0][2]; out[3] = R[1][1]; out[4] = R[1][2]; out[5] = R[2][2];
}

// the neighbour list itself is kept too: k_nn_corr3 searches a query's new correspondence among the neighbours of its old one
__device__ __forceinline__ void store_knn(const GicpDev& D, int c, int i, const KnnAcc<KNN_K>& acc) {
  int* o = D.nbr + ((size_t)c * D.nmax + i) * KNN_K;
#pragma unroll
  for (int k = 0; k < KNN_K; k++) o[k] = acc.id[k];
  D.nbrR2[(size_t)c * D.nmax + i] = acc.found == KNN_K ? acc.d[KNN_K - 1] : -1.0;
}

// estimate_local_features<CovarianceSetter>, one thread per query (shell search around the query).
// use_list = 0: every point of the cloud; use_list = 1: only the queries k_knn_cov_cells handed over
// (slotOf[] holds their point indices, nFall[] their number).
__global__ void __launch_bounds__(KNN_THREADS, 4) k_knn_cov(GicpDev D, int use_list) {
  extern __shared__ __align__(16) unsigned char s_knn[];
  double* s_d = reinterpret_cast<double*>(s_knn);
  unsigned* s_cells = reinterpret_cast<unsigned*>(s_d + KNN_LIST * KNN_THREADS);
  int* s_id = reinterpret_cast<int*>(s_cells + KNN_CELLS * KNN_THREADS);
  const int c = cloud_of(D, blockIdx.y);
  const int nPts = D.nDown[c];
  const int n = use_list ? D.nFall[c] : nPts;
  const int* list = D.slotOf + (size_t)c * D.nmax;
  const Grid g = make_grid(D, c);
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const int t = base + threadIdx.x;
    const bool active = t < n;
    int i = active ? (use_list ? list[t] : t) : 0;
    double qx, qy, qz;
    if (!use_list && D.cellOrder) {
      // queries in grid-cell order: the lanes of a warp share their home cell (a 0.1 m cell holds about a warp's worth of
      // points), so they probe the same hash slots and walk the same records -- the loads of a trip hit the same lines
      // and the per-lane loops have (nearly) the same trip counts.  Every downsampled point has a record: its cell key
      // is valid because its voxel key was.
      const double2 a = __ldg(&g.rec[2 * (size_t)i]), b = __ldg(&g.rec[2 * (size_t)i + 1]);
      qx = a.x; qy = a.y; qz = b.x;
      i = (int)__double_as_longlong(b.y);
    } else {
      const double* q = g.pts + (size_t)i * 4;
      qx = q[0]; qy = q[1]; qz = q[2];
    }
    KnnAcc<KNN_K> acc;
    bool finished = true;
    if (active)
      finished = knn10_two_pass(g, D.cellBox + c * 6, nPts, D.cell, qx, qy, qz, s_d + threadIdx.x, s_id + threadIdx.x,
                                s_cells + threadIdx.x, acc);
    // queries the shell search could not finish: the warp brute-forces them together, one at a time
    unsigned need = __ballot_sync(0xffffffffu, active && !finished);
    while (need) {
      const int src = __ffs(need) - 1;
      need &= need - 1;
      warp_brute_knn<KNN_K>(g.pts, nPts, __shfl_sync(0xffffffffu, qx, src), __shfl_sync(0xffffffffu, qy, src),
                            __shfl_sync(0xffffffffu, qz, src), src, acc);
    }
    if (active) {
      cov_from_knn(g, acc, D.cov + ((size_t)c * D.nmax + i) * 6);
      store_knn(D, c, i, acc);
    }
  }
}

// ---- the same search, split from the covariance (GFS_GICP_KNN=2).  k_knn_cov holds 128 registers and 48 KB of lists per
// CTA, so four CTAs (16 warps) share an SM and the search -- a chain of dependent loads -- waits with 37 % of the issue slots
// used.  The search alone, with lists sized to what a query really needs (<= 27 cells of shells 0 and 1, 10 candidates + ties;
// anything longer takes the shell-walk fallback as before), fits MINB CTAs; the covariance (the register-heavy 3x3
// eigen-decomposition) runs afterwards from the stored neighbour lists, one thread per point.  Bit-identical results.
template <int MINB, int LIST, int CELLS, bool DENSE>
__global__ void __launch_bounds__(KNN_THREADS, MINB) k_knn_search(GicpDev D) {
  extern __shared__ __align__(16) unsigned char s_knn[];
  double* s_d = reinterpret_cast<double*>(s_knn);
  unsigned* s_cells = reinterpret_cast<unsigned*>(s_d + LIST * KNN_THREADS);
  int* s_id = reinterpret_cast<int*>(s_cells + CELLS * KNN_THREADS);
  const int c = cloud_of(D, blockIdx.y);
  const int nPts = D.nDown[c];
  const auto g = [&] { if constexpr (DENSE) return make_dgrid(D, c); else return make_grid(D, c); }();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((int)(blockIdx.x * blockDim.x) >= nPts) return;
  const bool active = t < nPts;
  // queries in grid-cell order (see k_knn_cov)
  const int r = active ? t : 0;
  const double2 a = __ldg(&g.rec[2 * (size_t)r]), b = __ldg(&g.rec[2 * (size_t)r + 1]);
  double qx = a.x, qy = a.y, qz = b.x;
  const int i = (int)__double_as_longlong(b.y);
  KnnAcc<KNN_K> acc;
  bool finished = true;
  if (active)
    finished = knn10_two_pass<LIST, CELLS>(g, D.cellBox + c * 6, nPts, D.cell, qx, qy, qz, s_d + threadIdx.x, s_id + threadIdx.x,
                                           s_cells + threadIdx.x, acc);
  unsigned need = __ballot_sync(0xffffffffu, active && !finished);
  while (need) {
    const int src = __ffs(need) - 1;
    need &= need - 1;
    warp_brute_knn<KNN_K>(g.pts, nPts, __shfl_sync(0xffffffffu, qx, src), __shfl_sync(0xffffffffu, qy, src),
                          __shfl_sync(0xffffffffu, qz, src), src, acc);
  }
  if (active) store_knn(D, c, i, acc);   // missing neighbours (clouds with fewer than K points) keep the id 0x7fffffff
}
__global__ void __launch_bounds__(128) k_cov_nbr(GicpDev D) {
  const int c = cloud_of(D, blockIdx.y);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.nDown[c]) return;
  const Grid g = make_grid(D, c);
  const int2* o = reinterpret_cast<const int2*>(D.nbr + ((size_t)c * D.nmax + i) * KNN_K);
  KnnAcc<KNN_K> acc;
  acc.found = 0;
#pragma unroll
  for (int k = 0; k < KNN_K / 2; k++) {
    const int2 v = o[k];
    acc.id[2 * k] = v.x; acc.id[2 * k + 1] = v.y;
    acc.found += (v.x != 0x7fffffff) + (v.y != 0x7fffffff);
  }
  cov_from_knn(g, acc, D.cov + ((size_t)c * D.nmax + i) * 6);
}

// ---- cell-centric exact 10-NN + covariance.
// Eight lanes own one occupied grid cell.  They look its 26 neighbour cells up together (one hash probe
// chain per lane and round instead of ~27 dependent chains per query), copy the records of the 3x3x3
// block into shared memory once, and every lane then scans that list for one query of the cell: pass A
// keeps the K smallest distances rounded UP to float in a branch-free min/max chain (radius rho >= the
// true k-th distance), pass B remembers the positions of the candidates with d <= rho (one byte each, 16
// packed in four registers), the exact (distance, index) selection runs once over those.  The block is
// complete for a query when its k-th distance does not exceed cell + (distance to the nearest face of its
// own cell); the other queries (sparse surroundings, overfull blocks, ties) go to k_knn_cov(use_list = 1).
static const int KC_THREADS = 128;
static const int KC_LANES = 8;
static const int KC_CAP = 192;                        // candidates of one 3x3x3 block (positions fit a byte)
static const int KC_STRIDE = KC_CAP * 28 + 8;         // per-group bytes; = 8 mod 128: groups start on different banks
static const size_t KC_SMEM = (size_t)(KC_THREADS / KC_LANES) * KC_STRIDE;

__global__ void __launch_bounds__(KC_THREADS) k_knn_cov_cells(GicpDev D) {
  extern __shared__ __align__(16) unsigned char s_kc[];
  const int c = cloud_of(D, blockIdx.y);
  const int nCells = D.nCells[c];
  const int grp = threadIdx.x / KC_LANES, gl = threadIdx.x % KC_LANES;
  double* X = reinterpret_cast<double*>(s_kc + (size_t)grp * KC_STRIDE);
  double* Y = X + KC_CAP;
  double* Z = Y + KC_CAP;
  int* I = reinterpret_cast<int*>(Z + KC_CAP);
  const Grid g = make_grid(D, c);
  const int* cellSlot = D.cursor + (size_t)c * D.hsize;
  int* fall = D.slotOf + (size_t)c * D.nmax;
  for (int cell0 = blockIdx.x * (KC_THREADS / KC_LANES); cell0 < nCells; cell0 += gridDim.x * (KC_THREADS / KC_LANES)) {
    const int ci = cell0 + grp;
    const bool have = ci < nCells;
    int qs = 0, qn = 0;
    int st[4] = {0, 0, 0, 0}, cn[4] = {0, 0, 0, 0};
    if (have) {
      const uint4 e = __ldg(&g.tab[cellSlot[ci]]);
      const unsigned long long key = (unsigned long long)e.x | ((unsigned long long)e.y << 32);
      const int cx = (int)(key & 0x1fffff), cy = (int)((key >> 21) & 0x1fffff), cz = (int)((key >> 42) & 0x1fffff);
      qs = (int)e.z; qn = (int)e.w;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int nb = gl + KC_LANES * k;
        if (nb == 13) { st[k] = qs; cn[k] = qn; }
        else if (nb < 27) {
          int s0, n0;
          if (grid_find(g, cx + nb % 3 - 1, cy + (nb / 3) % 3 - 1, cz + nb / 9 - 1, s0, n0)) { st[k] = s0; cn[k] = n0; }
        }
      }
    }
    // offsets of this lane's cells in the group's candidate list (exclusive scan over the 8 lanes, 4 rounds)
    int off[4], total = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      int inc = cn[k];
#pragma unroll
      for (int o = 1; o < KC_LANES; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o, KC_LANES);
        if (gl >= o) inc += t;
      }
      off[k] = total + inc - cn[k];
      total += __shfl_sync(0xffffffffu, inc, KC_LANES - 1, KC_LANES);
    }
    const bool fits = total <= KC_CAP;
    __syncwarp();  // the previous cell's scans are done before its list is overwritten
    if (have && fits) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        for (int j = 0; j < cn[k]; j += 2) {
          const size_t r0 = 2 * (size_t)(st[k] + j);
          const bool two = j + 1 < cn[k];
          const double2 a0 = __ldg(&g.rec[r0]), b0 = __ldg(&g.rec[r0 + 1]);
          double2 a1 = a0, b1 = b0;
          if (two) { a1 = __ldg(&g.rec[r0 + 2]); b1 = __ldg(&g.rec[r0 + 3]); }
          const int o = off[k] + j;
          X[o] = a0.x; Y[o] = a0.y; Z[o] = b0.x; I[o] = (int)__double_as_longlong(b0.y);
          if (two) { X[o + 1] = a1.x; Y[o + 1] = a1.y; Z[o + 1] = b1.x; I[o + 1] = (int)__double_as_longlong(b1.y); }
        }
      }
    }
    __syncwarp();
    const int home = __shfl_sync(0xffffffffu, off[1], 13 - KC_LANES, KC_LANES);  // nb 13 = lane 5, round 1
    for (int q0 = 0; q0 < qn; q0 += KC_LANES) {
      const int qi = q0 + gl;
      if (qi >= qn) continue;
      int pidx;
      double qx, qy, qz;
      if (fits) { qx = X[home + qi]; qy = Y[home + qi]; qz = Z[home + qi]; pidx = I[home + qi]; }
      else {
        const double2 a = __ldg(&g.rec[2 * (size_t)(qs + qi)]), b = __ldg(&g.rec[2 * (size_t)(qs + qi) + 1]);
        qx = a.x; qy = a.y; qz = b.x; pidx = (int)__double_as_longlong(b.y);
      }
      bool ok = fits && total >= KNN_K;
      unsigned w0 = 0, w1 = 0, w2 = 0, w3 = 0;
      int cnt = 0;
      if (ok) {
        float top[KNN_K];
#pragma unroll
        for (int i = 0; i < KNN_K; i++) top[i] = FLT_MAX;
#pragma unroll 2
        for (int j = 0; j < total; j++) {
          const double dx = X[j] - qx, dy = Y[j] - qy, dz = Z[j] - qz;
          float v = __double2float_ru(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dz, dz)), __dmul_rn(dy, dy)));
#pragma unroll
          for (int i = 0; i < KNN_K; i++) {
            const float lo = fminf(top[i], v);
            v = fmaxf(top[i], v);
            top[i] = lo;
          }
        }
        const ShellQuery sq = make_shell_query(D.cell, qx, qy, qz);
        const double bound = D.cell + sq.margin;
        const double rho = (double)top[KNN_K - 1];
        ok = rho <= bound * bound;
        if (ok) {
          for (int j = 0; j < total; j++) {
            const double dx = X[j] - qx, dy = Y[j] - qy, dz = Z[j] - qz;
            const double dd = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dz, dz)), __dmul_rn(dy, dy));
            if (dd <= rho) {
              w3 = __funnelshift_l(w2, w3, 8); w2 = __funnelshift_l(w1, w2, 8); w1 = __funnelshift_l(w0, w1, 8);
              w0 = (w0 << 8) | (unsigned)j;
              cnt++;
            }
          }
          ok = cnt <= 16;
        }
      }
      if (!ok) {
        fall[atomicAdd(&D.nFall[c], 1)] = pidx;
        continue;
      }
      KnnAcc<KNN_K> acc;
      acc.init();
      for (int t = 0; t < cnt; t++) {
        const int j = (int)(w0 & 0xffu);
        w0 = __funnelshift_r(w0, w1, 8); w1 = __funnelshift_r(w1, w2, 8); w2 = __funnelshift_r(w2, w3, 8); w3 >>= 8;
        const double dx = X[j] - qx, dy = Y[j] - qy, dz = Z[j] - qz;
        acc.push(I[j], __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dz, dz)), __dmul_rn(dy, dy)));
      }
      cov_from_knn(g, acc, D.cov + ((size_t)c * D.nmax + pidx) * 6);
      store_knn(D, c, pidx, acc);
    }
  }
}

// ---- 10-NN + covariance, third generation: ONE WARP PER GRID CELL (the default; GFS_GICP_KNN=0 selects k_knn_cov alone).
// The lanes of a warp take the points of one cell as their queries (a 0.1 m cell holds about a warp's worth), so the
// candidate set -- the 3x3x3 block of cells around it -- is the same for all of them:
//  * 27 lanes probe the 27 cells at once (one hash chain each instead of 27 dependent chains per query);
//  * the block's records are staged ONCE per cell in shared memory as float32 relative to the home cell's origin (coalesced
//    16-byte loads of recf[]), octant by octant (k_cell_sort), with a table of the occupied octants ("blocks", ~4 points);
//  * every lane then runs the branch-free top-10 min / max chain over the staged candidates out of shared memory (broadcast
//    reads, all lanes active, no global load in the loop).  A block is skipped for the whole warp when no lane can still use
//    it -- its box is farther from every query than that query's current 10th distance -- which leaves roughly the blocks
//    inside (cell + 2 x 10-NN radius)^3, about half of the 3x3x3 block;
//  * pass B collects the <= 16 candidates within the float32 radius + error bound, the exact (distance, index) selection and
//    the covariance run in fp64 on those, exactly as in k_knn_cov.
// Error bound of a float32 squared distance v between two staged points (coordinates below 2 cells, each rounded twice):
//     E(v) = 3e-6 cell sqrt(v) + 1e-6 v + 1e-12      (twice the worst case 2 sqrt(3) d (4.4e-8 m at cell = 0.1) + 2.4e-7 d^2)
// A query is finished here iff the block provably holds its 10 nearest points (10th distance + E <= (cell + distance to the
// nearest face of its own cell)^2) and pass B's list did not overflow; the others -- sparse surroundings, overfull blocks,
// ties -- are appended to the cloud's hand-over list and done by k_knn_cov(use_list = 1).  Results are bit-identical.
static const int KW_WARPS = 4;
static const int KW_CAP = 384;     // staged candidates per cell block
static const int KW_BLOCKS = 216;  // 27 cells x 8 octants
static const int KW_LIST = 16;
struct KwSmem {
  float4 cand[KW_CAP];
  unsigned blk[KW_BLOCKS];         // start (10 bits) | count (8) | half-cell index x, y, z (3 bits each, 0..5)
  int list[KW_LIST * 32];
};
static const size_t KW_SMEM = sizeof(KwSmem) * KW_WARPS;
// visiting order of the 27 cells: home, 6 face, 12 edge, 8 corner neighbours (index = (dx+1) + 3 (dy+1) + 9 (dz+1))
__device__ const unsigned char KW_ORDER[27] = {13, 12, 14, 10, 16, 4, 22, 9, 11, 15, 17, 3, 5, 21, 23, 1, 7, 19, 25, 0, 2, 6, 8, 18, 20, 24, 26};

__device__ __forceinline__ float kw_err(float v, float ea) { return ea * sqrtf(v) + 1e-6f * v + 1e-12f; }
// squared distance from a query at (lx, ly, lz) (home-cell coordinates) to the half-cell box of a block, shrunk by 1 um
__device__ __forceinline__ float kw_gap2(unsigned e, float lx, float ly, float lz, float half) {
  const float bx = (float)((int)((e >> 18) & 7u) - 2) * half, by = (float)((int)((e >> 21) & 7u) - 2) * half,
              bz = (float)((int)((e >> 24) & 7u) - 2) * half;
  const float gx = fmaxf(fmaxf(bx - lx, lx - (bx + half)) - 1e-6f, 0.f), gy = fmaxf(fmaxf(by - ly, ly - (by + half)) - 1e-6f, 0.f),
              gz = fmaxf(fmaxf(bz - lz, lz - (bz + half)) - 1e-6f, 0.f);
  return gx * gx + gy * gy + gz * gz;
}

__global__ void __launch_bounds__(KW_WARPS * 32) k_knn_cov_warp(GicpDev D) {
  extern __shared__ __align__(16) unsigned char s_kw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  KwSmem& S = reinterpret_cast<KwSmem*>(s_kw)[warp];
  const int c = cloud_of(D, blockIdx.y);
  const int nCells = D.nCells[c];
  const Grid g = make_grid(D, c);
  const int* cellSlot = D.cursor + (size_t)c * D.hsize;
  int* fall = D.slotOf + (size_t)c * D.nmax;
  const float cellf = (float)D.cell, half = 0.5f * cellf, ea = (float)(3e-6 * D.cell);
  const unsigned full = 0xffffffffu;
  for (int ci = blockIdx.x * KW_WARPS + warp; ci < nCells; ci += gridDim.x * KW_WARPS) {
    const int hslot = cellSlot[ci];
    const uint4 he = __ldg(&g.tab[hslot]);
    const unsigned long long hkey = (unsigned long long)he.x | ((unsigned long long)he.y << 32);
    const int hx = (int)(hkey & 0x1fffff), hy = (int)((hkey >> 21) & 0x1fffff), hz = (int)((hkey >> 42) & 0x1fffff);
    const int qs = (int)he.z, qn = (int)he.w;
    // ---- the 27 cells: lane k probes cell KW_ORDER[k]
    int st = 0, cn = 0, dx = 0, dy = 0, dz = 0;
    unsigned long long n8 = 0;
    bool unsorted = false;
    if (lane < 27) {
      const int nb = KW_ORDER[lane];
      dx = nb % 3 - 1; dy = (nb / 3) % 3 - 1; dz = nb / 9 - 1;
      int slot = hslot;
      bool have = true;
      if (lane == 0) { st = qs; cn = qn; }
      else have = grid_find_slot(g, hx + dx, hy + dy, hz + dz, st, cn, slot);
      if (have) {
        const uint2 oc = __ldg(&g.oct[slot]);
        unsorted = oc.x == 0xffffffffu && oc.y == 0xffffffffu;
        n8 = (unsigned long long)oc.x | ((unsigned long long)oc.y << 32);
      } else { st = 0; cn = 0; }
    }
    int inc = cn;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(full, inc, o);
      if (lane >= o) inc += t;
    }
    const int base = inc - cn, total = __shfl_sync(full, inc, 31);
    const bool bad = __any_sync(full, unsorted) || total > KW_CAP || total < KNN_K;
    __syncwarp();  // the previous cell's queries are done with S
    if (bad) {     // hand every query of this cell over
      for (int qi = lane; qi < qn; qi += 32) {
        const int pidx = __float_as_int(__ldg(&g.recf[qs + qi]).w);
        fall[atomicAdd(&D.nFall[c], 1)] = pidx;
        D.knnCnt[(size_t)c * D.nmax + pidx] = 0;
      }
      continue;
    }
    // ---- stage the candidates cell by cell and list the occupied octants
    int nblk = 0;
    for (int k = 0; k < 27; k++) {
      const int kcn = __shfl_sync(full, cn, k);
      if (kcn == 0) continue;
      const int kst = __shfl_sync(full, st, k), kbase = __shfl_sync(full, base, k);
      const int kdx = __shfl_sync(full, dx, k), kdy = __shfl_sync(full, dy, k), kdz = __shfl_sync(full, dz, k);
      const unsigned long long kn8 = __shfl_sync(full, n8, k);
      const float ox = (float)kdx * cellf, oy = (float)kdy * cellf, oz = (float)kdz * cellf;
      for (int j = lane; j < kcn; j += 32) {
        float4 r = __ldg(&g.recf[kst + j]);
        r.x += ox; r.y += oy; r.z += oz;
        S.cand[kbase + j] = r;
      }
      const int ocnt = lane < 8 ? (int)((kn8 >> (8 * lane)) & 0xffull) : 0;
      int pre = ocnt;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const int t = __shfl_up_sync(full, pre, o);
        if (lane >= o) pre += t;
      }
      const unsigned m = __ballot_sync(full, ocnt > 0);
      if (ocnt > 0) {
        const int pos = nblk + __popc(m & ((1u << lane) - 1u));
        const unsigned hxh = (unsigned)((kdx + 1) * 2 + (lane & 1)), hyh = (unsigned)((kdy + 1) * 2 + ((lane >> 1) & 1)),
                       hzh = (unsigned)((kdz + 1) * 2 + ((lane >> 2) & 1));
        S.blk[pos] = (unsigned)(kbase + pre - ocnt) | ((unsigned)ocnt << 10) | (hxh << 18) | (hyh << 21) | (hzh << 24);
      }
      nblk += __popc(m);
    }
    __syncwarp();
    // ---- queries: the home cell's records (staged first, at offset 0), 32 per pass
    for (int q0 = 0; q0 < qn; q0 += 32) {
      const int qi = q0 + lane;
      const bool act = qi < qn;
      const float4 me = S.cand[act ? qi : 0];
      const float lx = me.x, ly = me.y, lz = me.z;
      const int pidx = __float_as_int(me.w);
      float top[KNN_K];
#pragma unroll
      for (int i = 0; i < KNN_K; i++) top[i] = FLT_MAX;
      // pass A: the 10 smallest float32 distances
      for (int b = 0; b < nblk; b++) {
        const unsigned e = S.blk[b];
        if (!__any_sync(full, act && kw_gap2(e, lx, ly, lz, half) <= top[KNN_K - 1])) continue;
        const int s0 = (int)(e & 0x3ffu), n0 = (int)((e >> 10) & 0xffu);
        for (int j = 0; j < n0; j++) {
          const float4 cd = S.cand[s0 + j];
          const float ddx = cd.x - lx, ddy = cd.y - ly, ddz = cd.z - lz;
          float v = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
#pragma unroll
          for (int i = 0; i < KNN_K; i++) {
            const float lo = fminf(top[i], v);
            v = fmaxf(top[i], v);
            top[i] = lo;
          }
        }
      }
      // is the block complete for this query?  (exact coordinates, fp64)
      const double* qp = g.pts + (size_t)(act ? pidx : 0) * 4;
      const double qx = qp[0], qy = qp[1], qz = qp[2];
      const float rho = top[KNN_K - 1];
      bool ok = act && rho < FLT_MAX;
      if (ok) {
        const ShellQuery sq = make_shell_query(D.cell, qx, qy, qz);
        const double bound = D.cell + sq.margin;
        ok = (double)rho + (double)kw_err(rho, ea) <= bound * bound;
      }
      // pass B: everything within the radius + twice the error bound goes to the exact selection
      const float T = ok ? (rho + 2.f * kw_err(rho, ea)) * 1.000001f : -1.f;
      int cnt = 0;
      for (int b = 0; b < nblk; b++) {
        const unsigned e = S.blk[b];
        if (!__any_sync(full, ok && kw_gap2(e, lx, ly, lz, half) <= T)) continue;
        const int s0 = (int)(e & 0x3ffu), n0 = (int)((e >> 10) & 0xffu);
        for (int j = 0; j < n0; j++) {
          const float4 cd = S.cand[s0 + j];
          const float ddx = cd.x - lx, ddy = cd.y - ly, ddz = cd.z - lz;
          const float v = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
          if (v <= T) {
            if (cnt < KW_LIST) S.list[cnt * 32 + lane] = __float_as_int(cd.w);
            cnt++;
          }
        }
      }
      ok = ok && cnt <= KW_LIST && cnt >= KNN_K;
      if (act && !ok) fall[atomicAdd(&D.nFall[c], 1)] = pidx;
      if (act) {
        // hand the short list to k_knn_select_cov (the fp64 selection and the eigen decomposition want ~120 registers;
        // keeping them out of this kernel lets four times as many warps hide the shared-memory latency of the scans)
        D.knnCnt[(size_t)c * D.nmax + pidx] = ok ? (unsigned char)cnt : (unsigned char)0;
        if (ok) {
          int* lst = D.knnList + ((size_t)c * D.nmax + pidx) * KW_LIST;
          for (int t = 0; t < cnt; t++) lst[t] = S.list[t * 32 + lane];
        }
      }
    }
  }
}

// exact (distance, index) selection over the <= 16 listed candidates of every query k_knn_cov_warp finished + covariance
__global__ void __launch_bounds__(128) k_knn_select_cov(GicpDev D) {
  const int c = cloud_of(D, blockIdx.y);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.nDown[c]) return;
  const int cnt = D.knnCnt[(size_t)c * D.nmax + i];
  if (cnt == 0) return;  // handed over to k_knn_cov(use_list = 1)
  const Grid g = make_grid(D, c);
  const double* q = g.pts + (size_t)i * 4;
  const double qx = q[0], qy = q[1], qz = q[2];
  const int4* lst = reinterpret_cast<const int4*>(D.knnList + ((size_t)c * D.nmax + i) * KW_LIST);
  KnnAcc<KNN_K> acc;
  acc.init();
  for (int t4 = 0; t4 < (cnt + 3) / 4; t4++) {
    const int4 v = __ldg(&lst[t4]);
    const int id[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (4 * t4 + u < cnt) acc.push(id[u], sqdist3(g.pts + (size_t)id[u] * 4, qx, qy, qz));
  }
  cov_from_knn(g, acc, D.cov + ((size_t)c * D.nmax + i) * 6);
  store_knn(D, c, i, acc);
}

// ---- block reduction of NV doubles per thread into out[NV] (fixed order: lanes, then warps)
// Warp sums of NV (<= 32) per-lane values by recursive halving: in the step with offset o a lane keeps one half of its
// remaining values and hands the other half to lane ^ o, so the whole vector costs 16 + 8 + 4 + 2 + 1 = 31 shuffles of a
// double instead of 5 per value (145 for the 29 sums of k_linearize, which made the kernel shuffle-bound: one 32-lane
// shuffle per clock and SM).  Lane l returns the total of v[l] (l < NV); the summation tree is fixed: (l, l^16), ^8, ... ^1.
template <int NV>
__device__ __forceinline__ double warp_reduce_vec(const double (&in)[NV]) {
  static_assert(NV <= 32, "one value per lane");
  const int lane = threadIdx.x & 31;
  double v[32];
#pragma unroll
  for (int k = 0; k < 32; k++) v[k] = k < NV ? in[k] : 0.0;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int k = 0; k < o; k++) {
      const double send = upper ? v[k] : v[k + o];
      const double keep = upper ? v[k + o] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}
template <int NV, int THREADS>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* out) {
  __shared__ double s_red[THREADS / 32][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if constexpr (NV < 7) {  // 5 shuffles per value beat the 31 of the halving scheme
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double x = v[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) s_red[warp][k] = x;
    }
  } else {
    const double x = warp_reduce_vec<NV>(v);
    if (lane < NV) s_red[warp][lane] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) t += s_red[w][threadIdx.x];
    out[threadIdx.x] = t;
  }
}

__device__ __forceinline__ void xform(const double* T, const double* p, double o[3]) {
  // T: rows of [R | t], 12 doubles
  for (int r = 0; r < 3; r++) o[r] = ((T[4 * r] * p[0] + T[4 * r + 1] * p[1]) + T[4 * r + 2] * p[2]) + T[4 * r + 3];
}

static const int LIN_THREADS = 128;

// 1-NN of every transformed source point in the target cloud within max_correspondence_distance
// (KdTree::nearest_neighbor_search + DistanceRejector, gicp_factor.hpp:40-48).  A kernel of its own: the
// search is a chain of dependent loads (hash probe -> cell records) and wants many resident warps, the
// linearisation is register-heavy arithmetic.  Candidates of a cell are fetched four at a time so four
// record loads are in flight per lane.
struct Nn1 {
  double d;
  int id;
  __device__ __forceinline__ void push(int index, double dist) {
    if (dist < d || (dist == d && index < id)) { d = dist; id = index; }
  }
};
template <class G>
__device__ __forceinline__ void scan_cell_nn1(const G& g, int s, int n, double qx, double qy, double qz, Nn1& nn) {
  for (int j = 0; j < n; j += 4) {
    double2 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const size_t r = 2 * (size_t)(s + min(j + u, n - 1));
      a[u] = __ldg(&g.rec[r]); b[u] = __ldg(&g.rec[r + 1]);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const double dx = a[u].x - qx, dy = a[u].y - qy, dz = b[u].x - qz;
      nn.push((int)__double_as_longlong(b[u].y), __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dz, dz)), __dmul_rn(dy, dy)));
    }
  }
}
static const int NN_THREADS = 128;
__global__ void __launch_bounds__(NN_THREADS, 6) k_nn_corr(GicpDev D) {
  const int p = pair_of(D, blockIdx.y);
  if (!D.istate[p * LM_ISTATE + I_ACTIVE] || D.istate[p * LM_ISTATE + I_NEED]) return;
  const int iter = D.istate[p * LM_ISTATE + I_OUTER];
  const int ct = tgt_cloud(D, p), cs = src_cloud(D, p);
  const int i = blockIdx.x * NN_THREADS + threadIdx.x;
  if (i >= D.nDown[cs]) return;
  const double* T = D.state + (size_t)p * LM_STATE + S_T;
  const double* ps = D.pts + ((size_t)cs * D.nmax + i) * 4;
  double q[3];
  xform(T, ps, q);
  const Grid g = make_grid(D, ct);
  const double max_d2 = D.max_dist * D.max_dist;
  const double cap = max_d2 * 1.0000001;
  Nn1 nn;
  nn.d = DBL_MAX; nn.id = 0x7fffffff;
  // the previous iteration's correspondence is a real target point: starting from it only tightens
  // the pruning radius, the result is still the exact nearest neighbour
  const int prev = iter > 0 ? D.corr[(size_t)p * D.nmax + i] : -1;
  if (prev >= 0) nn.push(prev, sqdist3(g.pts + (size_t)prev * 4, q[0], q[1], q[2]));
  const ShellQuery sq = make_shell_query(D.cell, q[0], q[1], q[2]);
  const int* box = D.cellBox + ct * 6;
  bool done = false;
  for (int r = 0; r <= MAX_SHELL && !done; r++) {
    visit_shell(g, box, sq, r, [&]() { return fmin(nn.d, cap); },
                [&](int cs_, int cn_) { scan_cell_nn1(g, cs_, cn_, q[0], q[1], q[2], nn); });
    const double bound = (double)r * D.cell + sq.margin;  // every unvisited point is farther than this
    const double b2 = bound * bound;
    done = nn.d <= b2 || b2 >= cap || box_covered(box, sq, r);
  }
  if (!done) {
    // unreachable while max_dist <= (MAX_SHELL + 1) cells; kept so the search stays exact for any setting
    nn.d = DBL_MAX; nn.id = 0x7fffffff;
    for (int t = 0; t < D.nDown[ct]; t++) nn.push(t, sqdist3(g.pts + (size_t)t * 4, q[0], q[1], q[2]));
  }
  D.corr[(size_t)p * D.nmax + i] = (nn.id != 0x7fffffff && !(nn.d > max_d2)) ? nn.id : -1;  // DistanceRejector: sq_dist > max_dist_sq
}

// ---- correspondence search, second generation (the default; GFS_GICP_NN=0 selects k_nn_corr above).
// Same result as k_nn_corr -- the exact nearest target point under (distance, index), or none within
// max_correspondence_distance -- reached with less work per query:
//  * queries are taken in the SOURCE cloud's grid-cell order (rec[]): a rigid transform keeps neighbours together,
//    so the lanes of a warp land in the same one to eight target cells and their probes / record loads hit the same
//    lines (in point order a warp's queries were strung out along an image row, ~0.6 m, six cells apart);
//  * the search region is the ball around the query whose radius is the best distance so far (the previous
//    iteration's correspondence seeds it: ~1 cm once the clouds are roughly aligned).  The home cell is scanned
//    first, then only the cells whose index range the ball reaches -- typically one to three more -- instead of
//    testing all 26 neighbours of shell 1 one by one;
//  * F32 = true: the scan runs on the float32 cell-local records (one 16-byte load and six float operations per
//    record instead of two loads and eight double operations) and keeps best + runner-up.  With the guaranteed
//    one-sided error bound nn_bound() the cells it visits are a superset of the exact search's; if the runner-up is
//    provably farther than the best, the best IS the exact nearest neighbour and its exact fp64 distance feeds the
//    rejection test; otherwise (about 1 query in 40 000) the lane repeats the search in fp64.
// Exactness of the region: a target point closer than `lim` lies in a cell whose lower bound axis_gap2 (shrunk by
// 1e-6 relative + 1e-12 against rounding in the cell assignment) does not exceed lim, and whose index is inside
// [floor((f - R) / cell), floor((f + R) / cell)] per axis for R = sqrt(lim) inflated the same way.
//
// Error bound of the float32 distance (scripts/gicp_prefilter_study.py measures it): a record's coordinates are below
// `cell`, a visited cell's origin is within max_dist + cell of the query per axis, each is rounded once (relative
// 2^-24), and d(d^2) = 2 sum |d_c| e_c plus the roundings of the float products and sums (< 5 * 2^-24 d^2):
//     bound(v) = A * sqrt(v) + 5e-7 * v + 1e-14,   A = 1.2 * 2 sqrt(3) 2^-24 (2 cell + max_dist)   (20 % of slack)
__device__ __forceinline__ float nn_bound(float v, float A) { return A * sqrtf(v) + 5e-7f * v + 1e-14f; }

struct Nn1f {
  float v1, v2;
  int id1;
};
__device__ __forceinline__ void scan_cell_nn1f(const Grid& g, int s, int n, float qx, float qy, float qz, Nn1f& nn) {
  for (int j = 0; j < n; j += 4) {
    float4 f[4];
#pragma unroll
    for (int u = 0; u < 4; u++) f[u] = __ldg(&g.recf[s + min(j + u, n - 1)]);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (j + u >= n) break;
      const int idx = __float_as_int(f[u].w);
      if (idx == nn.id1) continue;  // the seed (or the current best) met again: its fp32 twin must not become the runner-up
      const float dx = f[u].x - qx, dy = f[u].y - qy, dz = f[u].z - qz;
      const float v = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      if (v < nn.v1) { nn.v2 = nn.v1; nn.v1 = v; nn.id1 = idx; }
      else if (v < nn.v2) nn.v2 = v;
    }
  }
}

// cells the ball of squared radius limit() around the query reaches, home cell first; limit() may shrink while visiting
template <bool FAST, class Limit, class Visit>
__device__ __forceinline__ void visit_ball(const Grid& g, const int* box, const ShellQuery& q, Limit limit, Visit visit) {
  {
    int cs, cn, slot;
    if (grid_find_slot(g, q.cx, q.cy, q.cz, cs, cn, slot)) visit(cs, cn, q.cx, q.cy, q.cz, slot);
  }
  const double lim0 = limit();
  if (lim0 <= q.margin * q.margin) return;  // the ball stays inside the home cell
  const double R = sqrt(lim0) * 1.000001 + 1e-12, inv = 1.0 / q.cell;
  if (FAST && 2.0 * R < q.cell) {
    // The usual case once the clouds are roughly aligned (R ~ 1 cm, cell = 10 cm): the ball is narrower than a cell, so along
    // each axis it reaches at most ONE neighbour.  The seven candidate cells are the non-empty subsets of those axis steps:
    // a fixed seven-trip loop with three precomputed gaps instead of a triple loop over index ranges (the lanes of a warp
    // stay together; the triple loop ran at 12 active lanes).
    const int sx = (q.fx - R < 0.0) ? -1 : ((q.fx + R >= q.cell) ? 1 : 0), sy = (q.fy - R < 0.0) ? -1 : ((q.fy + R >= q.cell) ? 1 : 0),
              sz = (q.fz - R < 0.0) ? -1 : ((q.fz + R >= q.cell) ? 1 : 0);
    const double gx2 = axis_gap2(sx, q.fx, q.cell), gy2 = axis_gap2(sy, q.fy, q.cell), gz2 = axis_gap2(sz, q.fz, q.cell);
#pragma unroll
    for (int m = 1; m < 8; m++) {
      const bool ux = m & 1, uy = m & 2, uz = m & 4;
      if ((ux && !sx) || (uy && !sy) || (uz && !sz)) continue;
      const double g2 = (ux ? gx2 : 0.0) + (uy ? gy2 : 0.0) + (uz ? gz2 : 0.0);
      if (g2 > limit()) continue;
      const int x = q.cx + (ux ? sx : 0), y = q.cy + (uy ? sy : 0), z = q.cz + (uz ? sz : 0);
      if (x < box[0] || x > box[3] || y < box[1] || y > box[4] || z < box[2] || z > box[5]) continue;
      int cs, cn, slot;
      if (grid_find_slot(g, x, y, z, cs, cn, slot)) visit(cs, cn, x, y, z, slot);
    }
    return;
  }
  const int x0 = max(q.cx + fast_floor_d((q.fx - R) * inv), box[0]), x1 = min(q.cx + fast_floor_d((q.fx + R) * inv), box[3]);
  const int y0 = max(q.cy + fast_floor_d((q.fy - R) * inv), box[1]), y1 = min(q.cy + fast_floor_d((q.fy + R) * inv), box[4]);
  const int z0 = max(q.cz + fast_floor_d((q.fz - R) * inv), box[2]), z1 = min(q.cz + fast_floor_d((q.fz + R) * inv), box[5]);
  for (int z = z0; z <= z1; z++) {
    const double gz2 = axis_gap2(z - q.cz, q.fz, q.cell);
    if (gz2 > limit()) continue;
    for (int y = y0; y <= y1; y++) {
      const double gyz2 = gz2 + axis_gap2(y - q.cy, q.fy, q.cell);
      if (gyz2 > limit()) continue;
      for (int x = x0; x <= x1; x++) {
        if (x == q.cx && y == q.cy && z == q.cz) continue;
        if (gyz2 + axis_gap2(x - q.cx, q.fx, q.cell) > limit()) continue;
        int cs, cn, slot;
        if (grid_find_slot(g, x, y, z, cs, cn, slot)) visit(cs, cn, x, y, z, slot);
      }
    }
  }
}

// visit_ball for the dense grid: visit(first record, count) once per row of the ball's bounding block
template <bool FAST, class Limit, class Visit>
__device__ __forceinline__ void visit_ball(const DGrid& g, const ShellQuery& q, Limit limit, Visit visit) {
  int cs, cn;
  dense_run(g, q.cy, q.cz, q.cx, q.cx, cs, cn);
  if (cn) visit(cs, cn);
  const double lim0 = limit();
  if (lim0 <= q.margin * q.margin) return;  // the ball stays inside the home cell
  const double R = sqrt(lim0) * 1.000001 + 1e-12, inv = 1.0 / q.cell;
  const int xlo = g.ox, xhi = g.ox + g.nx - 1, ylo = g.oy, yhi = g.oy + g.ny - 1, zlo = g.oz, zhi = g.oz + g.nz - 1;
  if (FAST && 2.0 * R < q.cell) {
    // the ball reaches at most one neighbour per axis: four rows (own, y-step, z-step, both), in each the home column and / or
    // the x-step column -- adjacent cells, one run
    const int sx = (q.fx - R < 0.0) ? -1 : ((q.fx + R >= q.cell) ? 1 : 0), sy = (q.fy - R < 0.0) ? -1 : ((q.fy + R >= q.cell) ? 1 : 0),
              sz = (q.fz - R < 0.0) ? -1 : ((q.fz + R >= q.cell) ? 1 : 0);
    const double gx2 = axis_gap2(sx, q.fx, q.cell), gy2 = axis_gap2(sy, q.fy, q.cell), gz2 = axis_gap2(sz, q.fz, q.cell);
    const bool xin = sx != 0 && q.cx + sx >= xlo && q.cx + sx <= xhi;
#pragma unroll
    for (int m = 0; m < 4; m++) {
      const bool uy = m & 1, uz = m & 2;
      if ((uy && !sy) || (uz && !sz)) continue;
      const double gyz2 = (uy ? gy2 : 0.0) + (uz ? gz2 : 0.0);
      if (m && gyz2 > limit()) continue;
      const int y = q.cy + (uy ? sy : 0), z = q.cz + (uz ? sz : 0);
      if (y < ylo || y > yhi || z < zlo || z > zhi) continue;
      const bool side = xin && !(gyz2 + gx2 > limit());
      if (!m && !side) continue;
      const int xa = m ? (side && sx < 0 ? q.cx - 1 : q.cx) : q.cx + sx, xb = m ? (side && sx > 0 ? q.cx + 1 : q.cx) : q.cx + sx;
      dense_run(g, y, z, xa, xb, cs, cn);
      if (cn) visit(cs, cn);
    }
    return;
  }
  // the ball's bounding block in cells, BOTH ends clamped into the region: a ball that lies beyond the region (a clamped query)
  // maps to the border cells, which hold everything beyond
  auto span = [&](int c, double f, int lo_, int hi_, int& a, int& b) {
    const double lim = 3.0e6;
    a = min(max(c + fast_floor_d(fmin(fmax((f - R) * inv, -lim), lim)), lo_), hi_);
    b = min(max(c + fast_floor_d(fmin(fmax((f + R) * inv, -lim), lim)), lo_), hi_);
  };
  int x0, x1, y0, y1, z0, z1;
  span(q.cx, q.fx, xlo, xhi, x0, x1);
  span(q.cy, q.fy, ylo, yhi, y0, y1);
  span(q.cz, q.fz, zlo, zhi, z0, z1);
  for (int z = z0; z <= z1; z++) {
    const double gz2 = axis_gap2(z - q.cz, q.fz, q.cell);
    if (gz2 > limit()) continue;
    for (int y = y0; y <= y1; y++) {
      const double gyz2 = gz2 + axis_gap2(y - q.cy, q.fy, q.cell);
      if (gyz2 > limit()) continue;
      int xa = x0, xb = x1;
      while (xa < q.cx && gyz2 + axis_gap2(xa - q.cx, q.fx, q.cell) > limit()) xa++;
      while (xb > q.cx && gyz2 + axis_gap2(xb - q.cx, q.fx, q.cell) > limit()) xb--;
      if (y == q.cy && z == q.cz) {  // the home cell was visited first: the runs on either side of it
        if (xa < q.cx) { dense_run(g, y, z, xa, min(q.cx - 1, xb), cs, cn); if (cn) visit(cs, cn); }
        if (xb > q.cx) { dense_run(g, y, z, max(q.cx + 1, xa), xb, cs, cn); if (cn) visit(cs, cn); }
      } else {
        dense_run(g, y, z, xa, xb, cs, cn);
        if (cn) visit(cs, cn);
      }
    }
  }
}

// The octants of cell (x, y, z) a ball of squared radius lim around the query can reach, as an 8-bit mask (bit o = octant o of
// k_cell_sort).  Per axis the lower half [0, cell/2) and the upper half [cell/2, cell) of the cell are tested against
// [l - R, l + R] (l = query coordinate relative to the cell's origin, R inflated by 1e-6 relative + 1e-9: the octant of a
// point is decided in fp64 from the same origin, so 1e-9 m covers the rounding of that decision many times over).
__device__ __forceinline__ unsigned octant_mask(const double q[3], int x, int y, int z, double cell, double lim) {
  const int off = 1 << 20;
  const double R = sqrt(lim) * 1.000001 + 1e-9, half = 0.5 * cell;
  const double l[3] = {q[0] - (double)(x - off) * cell, q[1] - (double)(y - off) * cell, q[2] - (double)(z - off) * cell};
  const unsigned lo[3] = {0x55u, 0x33u, 0x0fu}, hi[3] = {0xaau, 0xccu, 0xf0u};  // octants in the lower / upper half along x, y, z
  unsigned m = 0xffu;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    unsigned ma = 0;
    if (l[a] - R < half && l[a] + R >= 0.0) ma |= lo[a];
    if (l[a] + R >= half && l[a] - R < cell) ma |= hi[a];
    m &= ma;
  }
  return m;
}
// scan(first record, count) for every octant of the cell whose bit is set in `mask`
template <class Scan>
__device__ __forceinline__ void for_octants(const Grid& g, int slot, int cs, int cn, unsigned mask, Scan scan) {
  const uint2 oc = __ldg(&g.oct[slot]);
  if (oc.x == 0xffffffffu && oc.y == 0xffffffffu) { scan(cs, cn); return; }  // an unsorted (overfull) cell: all of it
  unsigned long long n8 = (unsigned long long)oc.x | ((unsigned long long)oc.y << 32);
  int r = cs;
#pragma unroll
  for (int o = 0; o < 8; o++) {
    const int c = (int)(n8 & 0xffull);
    n8 >>= 8;
    if (c && ((mask >> o) & 1u)) scan(r, c);
    r += c;
  }
}

template <bool F32, bool OCT, bool FAST>
__global__ void __launch_bounds__(NN_THREADS, 6) k_nn_corr2(GicpDev D) {
  const int p = pair_of(D, blockIdx.y);
  if (!D.istate[p * LM_ISTATE + I_ACTIVE] || D.istate[p * LM_ISTATE + I_NEED]) return;
  const int iter = D.istate[p * LM_ISTATE + I_OUTER];
  const int ct = tgt_cloud(D, p), cs = src_cloud(D, p);
  const int t = blockIdx.x * NN_THREADS + threadIdx.x;
  if (t >= D.nDown[cs]) return;
  const double* T = D.state + (size_t)p * LM_STATE + S_T;
  int i = t;
  double ps[3];
  if (D.cellOrder) {
    const double2* rs = reinterpret_cast<const double2*>(D.rec + ((size_t)cs * D.nmax + t) * 4);
    const double2 a = __ldg(&rs[0]), b = __ldg(&rs[1]);
    ps[0] = a.x; ps[1] = a.y; ps[2] = b.x;
    i = (int)__double_as_longlong(b.y);
  } else {
    const double* pp = D.pts + ((size_t)cs * D.nmax + t) * 4;
    ps[0] = pp[0]; ps[1] = pp[1]; ps[2] = pp[2];
  }
  double q[3];
  xform(T, ps, q);
  const Grid g = make_grid(D, ct);
  const double max_d2 = D.max_dist * D.max_dist;
  const double cap = max_d2 * 1.0000001;
  const int prev = iter > 0 ? D.corr[(size_t)p * D.nmax + i] : -1;
  const ShellQuery sq = make_shell_query(D.cell, q[0], q[1], q[2]);
  const int* box = D.cellBox + ct * 6;
  int id = 0x7fffffff;
  double d = DBL_MAX;
  bool exact = !F32;
  if (F32) {
    const double capf = cap * 1.000001 + 1e-12;  // pruning cap of the float32 pass: never tighter than the exact one
    Nn1f nf;
    nf.v1 = FLT_MAX; nf.v2 = FLT_MAX; nf.id1 = 0x7fffffff;
    if (prev >= 0) {
      // seed: an upper bound of the exact distance to the previous correspondence
      nf.v1 = __double2float_ru(sqdist3(g.pts + (size_t)prev * 4, q[0], q[1], q[2]));
      nf.id1 = prev;
    }
    auto ub = [&]() -> double { return nf.v1 < FLT_MAX ? (double)nf.v1 + (double)nn_bound(nf.v1, D.nnBoundA) : DBL_MAX; };
    const int off = 1 << 20;
    visit_ball<FAST>(g, box, sq, [&]() { return fmin(ub(), capf); },
               [&](int cs_, int cn_, int x, int y, int z, int slot) {
                 const float qx = (float)(q[0] - (double)(x - off) * D.cell), qy = (float)(q[1] - (double)(y - off) * D.cell),
                             qz = (float)(q[2] - (double)(z - off) * D.cell);
                 if (OCT)
                   for_octants(g, slot, cs_, cn_, octant_mask(q, x, y, z, D.cell, fmin(ub(), capf)),
                               [&](int r0, int rn) { scan_cell_nn1f(g, r0, rn, qx, qy, qz, nf); });
                 else
                   scan_cell_nn1f(g, cs_, cn_, qx, qy, qz, nf);
               });
    // the best is the exact nearest neighbour iff no other candidate can be as close: the runner-up's lower bound is
    // above the best's upper bound (a seed that was never beaten carries an exact upper bound already)
    const bool clear_winner = nf.id1 != 0x7fffffff &&
                              (nf.v2 == FLT_MAX || (double)nf.v2 - (double)nn_bound(nf.v2, D.nnBoundA) > (double)nf.v1 + (double)nn_bound(nf.v1, D.nnBoundA));
    if (clear_winner) {
      id = nf.id1;
      d = sqdist3(g.pts + (size_t)id * 4, q[0], q[1], q[2]);
    } else if (nf.id1 != 0x7fffffff) {
      exact = true;  // ambiguous: redo this query in fp64
    }
  }
  if (exact) {
    Nn1 nn;
    nn.d = DBL_MAX; nn.id = 0x7fffffff;
    // the previous iteration's correspondence is a real target point: starting from it only tightens the radius
    if (prev >= 0) nn.push(prev, sqdist3(g.pts + (size_t)prev * 4, q[0], q[1], q[2]));
    visit_ball<FAST>(g, box, sq, [&]() { return fmin(nn.d, cap); },
               [&](int cs_, int cn_, int x, int y, int z, int slot) {
                 if (OCT)
                   for_octants(g, slot, cs_, cn_, octant_mask(q, x, y, z, D.cell, fmin(nn.d, cap)),
                               [&](int r0, int rn) { scan_cell_nn1(g, r0, rn, q[0], q[1], q[2], nn); });
                 else
                   scan_cell_nn1(g, cs_, cn_, q[0], q[1], q[2], nn);
               });
    id = nn.id; d = nn.d;
  }
  D.corr[(size_t)p * D.nmax + i] = (id != 0x7fffffff && !(d > max_d2)) ? id : -1;  // DistanceRejector: sq_dist > max_dist_sq
}

// ---- correspondence search, third generation (GFS_GICP_NN=7, the default): look among the NEIGHBOURS of the old correspondence first.
// k_knn_cov leaves, for every target point t, its 10 nearest points N(t) (t itself first) and r(t) = the distance to the 10th.  A
// point outside N(t) is at least r(t) away from t, hence at least r(t) - |q - t| away from a query q.  So for a query whose previous
// correspondence was t: scan the ten points of N(t); if the best of them is closer than r(t) - |q - t|, it IS the exact nearest
// target point (ties can only occur inside N(t), where the (distance, index) order decides as everywhere else).  That is eleven
// independent 32-byte gathers and no hash probe, no cell walk, no data-dependent loop: the lanes of a warp stay together (the ball
// walk runs at 13 active lanes).  Queries without a certificate -- first round, no previous correspondence, large residuals, a big
// step of the optimiser -- are COMPACTED inside the block and searched with the ball walk by the first lanes of the block, so
// the slow path is executed by full warps and only as many of them as needed.  The result is the exact search's, bit for bit.
template <class G>
__device__ __forceinline__ void nn1_ball_search(const GicpDev& D, const G& g, const int* box, const double q[3], int prev, double cap,
                                                int& id, double& d) {
  const ShellQuery sq = make_shell_query(g, D.cell, q[0], q[1], q[2]);
  Nn1 nn;
  nn.d = DBL_MAX; nn.id = 0x7fffffff;
  if (prev >= 0) nn.push(prev, sqdist3(g.pts + (size_t)prev * 4, q[0], q[1], q[2]));
  if constexpr (std::is_same<G, DGrid>::value)
    visit_ball<true>(g, sq, [&]() { return fmin(nn.d, cap); }, [&](int cs_, int cn_) { scan_cell_nn1(g, cs_, cn_, q[0], q[1], q[2], nn); });
  else
    visit_ball<true>(g, box, sq, [&]() { return fmin(nn.d, cap); },
                     [&](int cs_, int cn_, int, int, int, int) { scan_cell_nn1(g, cs_, cn_, q[0], q[1], q[2], nn); });
  id = nn.id; d = nn.d;
}

template <int MINB, bool DENSE>
__global__ void __launch_bounds__(NN_THREADS, MINB) k_nn_corr3(GicpDev D) {
  __shared__ double s_q[NN_THREADS][3];
  __shared__ int s_i[NN_THREADS], s_prev[NN_THREADS], s_list[NN_THREADS];
  __shared__ int s_wcnt[NN_THREADS / 32], s_total;
  const int p = pair_of(D, blockIdx.y);
  if (!D.istate[p * LM_ISTATE + I_ACTIVE] || D.istate[p * LM_ISTATE + I_NEED]) return;
  const int iter = D.istate[p * LM_ISTATE + I_OUTER];
  const int ct = tgt_cloud(D, p), cs = src_cloud(D, p);
  const int ns = D.nDown[cs];
  if ((int)(blockIdx.x * NN_THREADS) >= ns) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = blockIdx.x * NN_THREADS + tid;
  const bool valid = t < ns;
  const double* T = D.state + (size_t)p * LM_STATE + S_T;
  const auto g = [&] { if constexpr (DENSE) return make_dgrid(D, ct); else return make_grid(D, ct); }();
  const double max_d2 = D.max_dist * D.max_dist;
  const double cap = max_d2 * 1.0000001;
  int i = 0, prev = -1;
  double q[3] = {0, 0, 0};
  bool need = false;
  if (valid) {
    double ps[3];
    i = t;
    if (D.cellOrder) {
      const double2* rs = reinterpret_cast<const double2*>(D.rec + ((size_t)cs * D.nmax + t) * 4);
      const double2 a = __ldg(&rs[0]), b = __ldg(&rs[1]);
      ps[0] = a.x; ps[1] = a.y; ps[2] = b.x;
      i = (int)__double_as_longlong(b.y);
    } else {
      const double* pp = D.pts + ((size_t)cs * D.nmax + t) * 4;
      ps[0] = pp[0]; ps[1] = pp[1]; ps[2] = pp[2];
    }
    xform(T, ps, q);
    prev = iter > 0 ? D.corr[(size_t)p * D.nmax + i] : -1;
    need = true;
    if (prev >= 0) {
      const double r2 = __ldg(&D.nbrR2[(size_t)ct * D.nmax + prev]);
      if (r2 >= 0.0) {
        const int2* nb = reinterpret_cast<const int2*>(D.nbr + ((size_t)ct * D.nmax + prev) * KNN_K);
        int ids[KNN_K];
#pragma unroll
        for (int k = 0; k < KNN_K / 2; k++) { const int2 v = __ldg(&nb[k]); ids[2 * k] = v.x; ids[2 * k + 1] = v.y; }
        Nn1 nn;
        nn.d = DBL_MAX; nn.id = 0x7fffffff;
        double dprev2 = 0.0;
#pragma unroll
        for (int k = 0; k < KNN_K; k++) {
          const double dd = sqdist3(g.pts + (size_t)ids[k] * 4, q[0], q[1], q[2]);
          if (ids[k] == prev) dprev2 = dd;
          nn.push(ids[k], dd);
        }
        // prev heads its own list unless other points coincide with it exactly; should it be missing from the list altogether
        // (ten coincident points: r = 0, no certificate anyway) its distance is measured directly
        if (ids[0] != prev && dprev2 == 0.0) dprev2 = sqdist3(g.pts + (size_t)prev * 4, q[0], q[1], q[2]);
        const double bound = sqrt(r2) * (1.0 - 1e-9) - 1e-12 - sqrt(dprev2) * (1.0 + 1e-9);
        if (bound > 0.0 && sqrt(nn.d) * (1.0 + 1e-9) < bound) {
          D.corr[(size_t)p * D.nmax + i] = !(nn.d > max_d2) ? nn.id : -1;  // DistanceRejector: sq_dist > max_dist_sq
          need = false;
        }
      }
    }
  }
  // ---- compact the queries that still need the full search: ballot + per-warp offsets
  const unsigned m = __ballot_sync(0xffffffffu, need);
  if (lane == 0) s_wcnt[warp] = __popc(m);
  if (need) { s_q[tid][0] = q[0]; s_q[tid][1] = q[1]; s_q[tid][2] = q[2]; s_i[tid] = i; s_prev[tid] = prev; }
  __syncthreads();
  if (tid == 0) {
    int a = 0;
    for (int w = 0; w < NN_THREADS / 32; w++) { const int c = s_wcnt[w]; s_wcnt[w] = a; a += c; }
    s_total = a;
  }
  __syncthreads();
  if (need) s_list[s_wcnt[warp] + __popc(m & ((1u << lane) - 1u))] = tid;
  __syncthreads();
  const int total = s_total;
  if (tid >= total) return;
  const int src = s_list[tid];
  const double qq[3] = {s_q[src][0], s_q[src][1], s_q[src][2]};
  int id;
  double d;
  nn1_ball_search(D, g, D.cellBox + ct * 6, qq, s_prev[src], cap, id, d);
  D.corr[(size_t)p * D.nmax + s_i[src]] = (id != 0x7fffffff && !(d > max_d2)) ? id : -1;
}

__device__ void lm_begin_block(const GicpDev& D, int p);   // defined below (need lm_trial)
__device__ void lm_decide_thread(const GicpDev& D, int p);
// true in every thread of exactly one block of pair p: the one that finished last among the `nblocks` participating ones
// (classic threadfence reduction: results written before the fence are visible to the block that draws the last ticket)
__device__ __forceinline__ bool last_block_of_pair(const GicpDev& D, int p, int which, int nblocks) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(&D.tickets[2 * p + which], 1);
    s_last = (t == nblocks - 1);
    if (s_last) D.tickets[2 * p + which] = 0;  // ready for the next round
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// GICPFactor::linearize for every source point of every active pair + partial sums; the pair's last block then sums the
// partials in their fixed order and starts the first lambda trial (what k_lm_begin did in a launch of its own)
template <int MINB>
__global__ void __launch_bounds__(LIN_THREADS, MINB) k_linearize(GicpDev D) {
  const int p = pair_of(D, blockIdx.y);
  if (!D.istate[p * LM_ISTATE + I_ACTIVE] || D.istate[p * LM_ISTATE + I_NEED]) return;
  const int ct = tgt_cloud(D, p), cs = src_cloud(D, p);
  const int ns = D.nDown[cs];
  const int PPT = D.linPpt;
  const int nbp = max((ns + LIN_THREADS * PPT - 1) / (LIN_THREADS * PPT), 1);  // participating blocks (block 0 always: an empty cloud still steps)
  if ((int)blockIdx.x >= nbp) return;
  double acc[RED_N];
#pragma unroll
  for (int k = 0; k < RED_N; k++) acc[k] = 0.0;
  const double* T = D.state + (size_t)p * LM_STATE + S_T;
  // PPT points per thread, LIN_THREADS apart (coalesced), summed in that order before the warp reduction
  // the correspondence of the NEXT point is fetched one trip ahead: a trip's gathers (target point, target covariance) then start
  // at once instead of after the index load they depend on
  const int* corr = D.corr + (size_t)p * D.nmax;
  const int i0 = blockIdx.x * PPT * LIN_THREADS + threadIdx.x;
  int tgtNext = i0 < ns ? corr[i0] : -1;
#pragma unroll 1
  for (int j = 0; j < PPT; j++) {
    const int i = i0 + j * LIN_THREADS;
    if (i >= ns) break;
    const int tgt = tgtNext;
    tgtNext = (j + 1 < PPT && i + LIN_THREADS < ns) ? corr[i + LIN_THREADS] : -1;
    const double* ps = D.pts + ((size_t)cs * D.nmax + i) * 4;
    double q[3];
    xform(T, ps, q);
    const Grid g = make_grid(D, ct);
    if (tgt >= 0) {
      const double* cS = D.cov + ((size_t)cs * D.nmax + i) * 6;
      const double* cT = D.cov + ((size_t)ct * D.nmax + tgt) * 6;
      const double Cs[3][3] = {{cS[0], cS[1], cS[2]}, {cS[1], cS[3], cS[4]}, {cS[2], cS[4], cS[5]}};
      const double Ct[3][3] = {{cT[0], cT[1], cT[2]}, {cT[1], cT[3], cT[4]}, {cT[2], cT[4], cT[5]}};
      double R[3][3];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) R[r][c] = T[4 * r + c];
      double RC[3][3], A[3][3];
      for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) RC[r][k] = R[r][0] * Cs[0][k] + R[r][1] * Cs[1][k] + R[r][2] * Cs[2][k];
      for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) A[r][k] = Ct[r][k] + (RC[r][0] * R[k][0] + RC[r][1] * R[k][1] + RC[r][2] * R[k][2]);
      // 3x3 inverse by cofactors
      const double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1], c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2],
                   c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
      const double id = 1.0 / (A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02);
      double M[3][3];
      M[0][0] = c00 * id; M[1][0] = c01 * id; M[2][0] = c02 * id;
      M[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id;
      M[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
      M[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id;
      M[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
      M[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
      M[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
      double* mo = D.maha + ((size_t)p * D.nmax + i) * 9;
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) mo[3 * r + c] = M[r][c];
      const double* pt = g.pts + (size_t)tgt * 4;
      const double res[3] = {pt[0] - q[0], pt[1] - q[1], pt[2] - q[2]};
      const double S[3][3] = {{0, -ps[2], ps[1]}, {ps[2], 0, -ps[0]}, {-ps[1], ps[0], 0}};
      double J[3][6];
      for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) {
          J[r][k] = R[r][0] * S[0][k] + R[r][1] * S[1][k] + R[r][2] * S[2][k];
          J[r][3 + k] = -R[r][k];
        }
      double MJ[3][6], Mr[3];
      for (int r = 0; r < 3; r++) {
        for (int k = 0; k < 6; k++) MJ[r][k] = M[r][0] * J[0][k] + M[r][1] * J[1][k] + M[r][2] * J[2][k];
        Mr[r] = M[r][0] * res[0] + M[r][1] * res[1] + M[r][2] * res[2];
      }
      int n = 0;
#pragma unroll
      for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = a; b < 6; b++) acc[n++] += J[0][a] * MJ[0][b] + J[1][a] * MJ[1][b] + J[2][a] * MJ[2][b];
#pragma unroll
      for (int a = 0; a < 6; a++) acc[21 + a] += J[0][a] * Mr[0] + J[1][a] * Mr[1] + J[2][a] * Mr[2];
      acc[27] += 0.5 * (res[0] * Mr[0] + res[1] * Mr[1] + res[2] * Mr[2]);
      acc[28] += 1.0;
    }
  }
  // per-warp partial sums (no block barrier: a warp whose lanes all found their neighbour quickly retires
  // without waiting for a slow sibling); k_lm_begin adds them in a fixed order
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* out = D.partial + (((size_t)p * D.nblk + blockIdx.x) * (LIN_THREADS / 32) + warp) * RED_N;
    const double x = warp_reduce_vec<RED_N>(acc);
    if (lane < RED_N) out[lane] = x;
  }
  if (last_block_of_pair(D, p, 0, nbp)) lm_begin_block(D, p);
}

// GICPFactor::error with the stored correspondences / Mahalanobis matrices; the pair's last block then decides the trial
// (what k_lm_decide did in a launch of its own)
__global__ void __launch_bounds__(LIN_THREADS) k_error(GicpDev D) {
  const int p = pair_of(D, blockIdx.y);
  if (!D.istate[p * LM_ISTATE + I_NEED]) return;
  const int ct = tgt_cloud(D, p), cs = src_cloud(D, p);
  const int ns = D.nDown[cs];
  const int per = LIN_THREADS * D.errPpt;
  const int nbp = max((ns + per - 1) / per, 1);
  if ((int)blockIdx.x >= nbp) return;
  // errPpt points per thread, two per trip: both correspondences, then both gathers are in flight before the first is used
  // (one point per thread left the kernel waiting out a dependent corr -> target-point chain at 14 % of the issue slots, and a
  // block of 128 points paid a fence, a ticket and a partial sum for 15 flops per thread)
  const double* T = D.state + (size_t)p * LM_STATE + S_NEWT;
  const double* sp = D.pts + (size_t)cs * D.nmax * 4;
  const double* tp = D.pts + (size_t)ct * D.nmax * 4;
  const int* corr = D.corr + (size_t)p * D.nmax;
  const double* maha = D.maha + (size_t)p * D.nmax * 9;
  double acc[1] = {0.0};
#pragma unroll 1
  for (int j = 0; j < D.errPpt; j += 2) {
    int i[2], tgt[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      i[u] = (blockIdx.x * D.errPpt + j + u) * LIN_THREADS + threadIdx.x;
      tgt[u] = (j + u < D.errPpt && i[u] < ns) ? corr[i[u]] : -1;
    }
    double ps[2][3], pt[2][3], M[2][9];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      if (tgt[u] < 0) continue;
#pragma unroll
      for (int k = 0; k < 3; k++) { ps[u][k] = sp[(size_t)i[u] * 4 + k]; pt[u][k] = tp[(size_t)tgt[u] * 4 + k]; }
#pragma unroll
      for (int k = 0; k < 9; k++) M[u][k] = maha[(size_t)i[u] * 9 + k];
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      if (tgt[u] < 0) continue;
      double q[3];
      xform(T, ps[u], q);
      const double res[3] = {pt[u][0] - q[0], pt[u][1] - q[1], pt[u][2] - q[2]};
      double Mr[3];
      for (int r = 0; r < 3; r++) Mr[r] = M[u][3 * r] * res[0] + M[u][3 * r + 1] * res[1] + M[u][3 * r + 2] * res[2];
      acc[0] += 0.5 * (res[0] * Mr[0] + res[1] * Mr[1] + res[2] * Mr[2]);
    }
  }
  block_reduce_store<1, LIN_THREADS>(acc, D.partialE + (size_t)p * D.nblk + blockIdx.x);
  if (last_block_of_pair(D, p, 1, nbp) && threadIdx.x == 0) lm_decide_thread(D, p);
}

// ---- 6x6 LDL^T solve, se3_exp, compose (one thread)
__device__ void ldlt_solve6(const double Hin[6][6], const double rhs[6], double x[6]) {
  double L[6][6], Dg[6];
  for (int a = 0; a < 6; a++)
    for (int b = 0; b < 6; b++) L[a][b] = 0;
  for (int j = 0; j < 6; j++) {
    double d = Hin[j][j];
    for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k] * Dg[k];
    Dg[j] = d;
    L[j][j] = 1;
    for (int i = j + 1; i < 6; i++) {
      double v = Hin[i][j];
      for (int k = 0; k < j; k++) v -= L[i][k] * L[j][k] * Dg[k];
      L[i][j] = v / d;
    }
  }
  double y[6];
  for (int i = 0; i < 6; i++) {
    double v = rhs[i];
    for (int k = 0; k < i; k++) v -= L[i][k] * y[k];
    y[i] = v;
  }
  for (int i = 0; i < 6; i++) y[i] /= Dg[i];
  for (int i = 5; i >= 0; i--) {
    double v = y[i];
    for (int k = i + 1; k < 6; k++) v -= L[k][i] * x[k];
    x[i] = v;
  }
}
__device__ void se3_exp_dev(const double a[6], double T[12]) {  // util/lie.hpp:52-96
  const double w[3] = {a[0], a[1], a[2]};
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double imag, real;
  if (theta_sq < 1e-10) {
    const double tq = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * tq;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * tq;
  } else {
    const double theta = sqrt(theta_sq), half = 0.5 * theta;
    imag = sin(half) / theta;
    real = cos(half);
  }
  const double qw = real, qx = imag * w[0], qy = imag * w[1], qz = imag * w[2];
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy,
               tyz = tz * qy, tzz = tz * qz;
  double R[3][3];
  R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz; R[0][2] = txz + twy;
  R[1][0] = txy + twz; R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy; R[2][1] = tyz + twx; R[2][2] = 1 - (txx + tyy);
  const double theta = sqrt(theta_sq);
  const double tr[3] = {a[3], a[4], a[5]};
  double t[3];
  if (theta < 1e-10) {
    for (int r = 0; r < 3; r++) t[r] = R[r][0] * tr[0] + R[r][1] * tr[1] + R[r][2] * tr[2];
  } else {
    const double O[3][3] = {{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}};
    const double k1 = (1.0 - cos(theta)) / theta_sq, k2 = (theta - sin(theta)) / (theta_sq * theta);
    for (int r = 0; r < 3; r++) {
      double Vr[3];
      for (int c = 0; c < 3; c++) {
        double oo = 0;
        for (int k = 0; k < 3; k++) oo += O[r][k] * O[k][c];
        Vr[c] = (r == c ? 1.0 : 0.0) + k1 * O[r][c] + k2 * oo;
      }
      t[r] = Vr[0] * tr[0] + Vr[1] * tr[1] + Vr[2] * tr[2];
    }
  }
  for (int r = 0; r < 3; r++) { T[4 * r] = R[r][0]; T[4 * r + 1] = R[r][1]; T[4 * r + 2] = R[r][2]; T[4 * r + 3] = t[r]; }
}
__device__ void compose12(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) C[4 * r + c] = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c];
    C[4 * r + 3] = A[4 * r] * B[3] + A[4 * r + 1] * B[7] + A[4 * r + 2] * B[11] + A[4 * r + 3];
  }
}
// delta = (H + lambda I)^-1 (-b); newT = T * se3_exp(delta)       (optimizer.hpp:107-112)
__device__ void lm_trial(double* st) {
  double H[6][6], nb[6], delta[6];
  int n = 0;
  for (int a = 0; a < 6; a++)
    for (int b = a; b < 6; b++) { H[a][b] = H[b][a] = st[S_H + n]; n++; }
  for (int a = 0; a < 6; a++) { H[a][a] += st[S_LAMBDA]; nb[a] = -st[S_B + a]; }
  ldlt_solve6(H, nb, delta);
  double dT[12];
  se3_exp_dev(delta, dT);
  compose12(st + S_T, dT, st + S_NEWT);
  for (int a = 0; a < 6; a++) st[S_DELTA + a] = delta[a];
}

// all LIN_THREADS threads of one block: sum the pair's per-warp partials in block order, start the first trial (:97-112)
__device__ __noinline__ void lm_begin_block(const GicpDev& D, int p) {
  __shared__ double s_tot[RED_N];
  const int tid = threadIdx.x;
  int* is = D.istate + p * LM_ISTATE;
  const int iter = is[I_OUTER];
  double* st = D.state + (size_t)p * LM_STATE;
  // fixed-order sum of the per-warp partials: thread t takes partials t, t+128, ...; then lanes, then warps
  const int per = LIN_THREADS * D.linPpt;
  const int nw = max((D.nDown[src_cloud(D, p)] + per - 1) / per, 1) * (LIN_THREADS / 32);
  double acc[RED_N];
#pragma unroll
  for (int k = 0; k < RED_N; k++) acc[k] = 0.0;
  const double* part = D.partial + (size_t)p * D.nblk * (LIN_THREADS / 32) * RED_N;
  for (int b = tid; b < nw; b += LIN_THREADS) {
#pragma unroll
    for (int k = 0; k < RED_N; k++) acc[k] += __ldcg(&part[(size_t)b * RED_N + k]);  // written by other blocks: read through L2
  }
  block_reduce_store<RED_N, LIN_THREADS>(acc, s_tot);
  __syncthreads();
  if (tid < RED_N) {
    const double s = s_tot[tid];
    if (tid < 21) st[S_H + tid] = s;
    else if (tid < 27) st[S_B + tid - 21] = s;
    else if (tid == 27) st[S_E] = s;
    else if (tid == 28) is[I_INL] = (int)(s + 0.5);
  }
  __syncthreads();
  if (tid == 0) {
    is[I_ITER] = iter;
    is[I_TRIAL] = 0;
    is[I_SUCCESS] = 0;
    lm_trial(st);
    __threadfence();
    is[I_NEED] = 1;
  }
}

// one thread: new_e, accept / reject, next trial or end of the outer iteration (:114-143)
__device__ __noinline__ void lm_decide_thread(const GicpDev& D, int p) {
  int* is = D.istate + p * LM_ISTATE;
  const int iter = is[I_OUTER];
  double* st = D.state + (size_t)p * LM_STATE;
  const int nb = max((D.nDown[src_cloud(D, p)] + LIN_THREADS * D.errPpt - 1) / (LIN_THREADS * D.errPpt), 1);
  // fixed-order sum: lane-strided partial sums are NOT order-preserving, so one thread sums serially
  double new_e = 0;
  for (int b = 0; b < nb; b++) new_e += __ldcg(&D.partialE[(size_t)p * D.nblk + b]);
  is[I_INNER]++;
  if (new_e <= st[S_E]) {
    const double* d = st + S_DELTA;
    const double dr = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const double dt = sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
    is[I_CONV] = (dr <= D.rot_eps && dt <= D.trans_eps) ? 1 : 0;
    for (int k = 0; k < 12; k++) st[S_T + k] = st[S_NEWT + k];
    st[S_LAMBDA] /= 10.0;
    is[I_SUCCESS] = 1;
    is[I_NEED] = 0;
  } else {
    st[S_LAMBDA] *= 10.0;
    is[I_TRIAL]++;
    if (is[I_TRIAL] >= 10) is[I_NEED] = 0;
    else lm_trial(st);  // the next round's error slot evaluates it
  }
  if (!is[I_NEED]) {  // the outer iteration is over for this pair
    if (!is[I_SUCCESS] || is[I_CONV] || iter + 1 >= D.max_iter) is[I_ACTIVE] = 0;
    else is[I_OUTER] = iter + 1;
  }
  if (is[I_ACTIVE] && D.listWrite >= 0)   // pairs with work left (the host reads the number in the rounds it checks)
    D.activeList[D.listWrite * D.listStride + atomicAdd(&D.counters[1], 1)] = p;
}

__global__ void k_lm_init(GicpDev D, int pairs, const double* __restrict__ T0) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pairs) return;
  double* st = D.state + (size_t)p * LM_STATE;
  for (int k = 0; k < LM_STATE; k++) st[k] = 0;
  for (int k = 0; k < 12; k++) st[S_T + k] = T0[(size_t)p * 16 + k];
  st[S_LAMBDA] = 1e-3;
  int* is = D.istate + p * LM_ISTATE;
  for (int k = 0; k < LM_ISTATE; k++) is[k] = 0;
  is[I_ACTIVE] = D.max_iter > 0 ? 1 : 0;
  D.tickets[2 * p] = 0; D.tickets[2 * p + 1] = 0;
}

__global__ void k_gicp_result(GicpDev D, int pairs, GfsGicpResult* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pairs) return;
  const double* st = D.state + (size_t)p * LM_STATE;
  const int* is = D.istate + p * LM_ISTATE;
  GfsGicpResult r;
  for (int k = 0; k < 12; k++) r.T[k] = st[S_T + k];
  r.T[12] = r.T[13] = r.T[14] = 0; r.T[15] = 1;
  int n = 0;
  for (int a = 0; a < 6; a++)
    for (int b = a; b < 6; b++) { r.H[6 * a + b] = r.H[6 * b + a] = st[S_H + n]; n++; }
  for (int a = 0; a < 6; a++) r.b[a] = st[S_B + a];
  r.error = st[S_E];
  r.iterations = is[I_ITER];
  r.num_inliers = is[I_INL];
  r.converged = is[I_CONV];
  r.n_target = D.nDown[tgt_cloud(D, p)];
  r.n_source = D.nDown[src_cloud(D, p)];
  r.inner_evals = is[I_INNER];
  out[p] = r;
}

// per-cloud input counts, clamped to [0, stride]: a count beyond the batch stride would make the grouping kernels read
// the next pair's cloud.  nt == nullptr (track mode): only the new cloud of every sequence, cloud 2p + cbase.
__global__ void k_set_counts(GicpDev D, int pairs, const int* __restrict__ nt, const int* __restrict__ ns, int stride) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pairs) return;
  const int lim = min(stride, D.nmax);
  if (nt) {
    D.nIn[2 * p] = min(max(nt[p], 0), lim);
    D.nIn[2 * p + 1] = min(max(ns[p], 0), lim);
  } else {
    D.nIn[2 * p + D.cbase] = min(max(ns[p], 0), lim);
  }
}

}  // namespace gfs

using namespace gfs;

struct GfsGicp {
  GicpDev dev;
  int maxPairs = 0;
  DevBuf b_keys, b_minIdx, b_count, b_start, b_cursor, b_rank, b_slotOf, b_members, b_nIn, b_nDown, b_nCells, b_nFall, b_box, b_pts, b_cov, b_tab, b_rec, b_recf, b_oct, b_knnList, b_knnCnt, b_nbr, b_nbrR2,
      b_corr, b_maha, b_partial, b_partialE, b_state, b_istate, b_counters, b_tickets, b_dS, b_dSum, b_active;
  DevBuf b_tgt, b_src, b_n, b_T0, b_res;
  PinnedBuf h_counters;
  int launches = 0;
  bool cellKnn = false;  // GFS_GICP_KNN_CELLS=1: cell-centric 10-NN kernel first (same results)
  int nnMode = 8;        // GFS_GICP_NN: 0 = k_nn_corr (shell walk), 1 = k_nn_corr2 (ball walk, fp64), 2 = + float32 prefilter, 3 = ball walk over
                         // octants (fp64), 4 = octants + float32 prefilter, 5 = ball walk with the seven-cell fast path (fp64), 6 = 5 + float32
                         // prefilter, 7..10 = k_nn_corr3: neighbours of the old correspondence first, certified by the 10-NN radius, ball walk
                         // for the compacted rest, compiled for 6 / 8 (default) / 7 / 5 CTAs per SM.  0..6 use the hash grid.
  int knnMode = 3;       // GFS_GICP_KNN: 2..5 = k_knn_search + k_cov_nbr (search split from the covariance) compiled for 6 / 8 (default; 7 on
                         // the hash grid) / 5 / 4 CTAs per SM, 0 = k_knn_cov (thread per query, search + covariance in one kernel),
                         // 1 = k_knn_cov_warp (warp per cell, octant skipping) + k_knn_cov for what it hands over.  0 and 1 use the hash grid.
  int linMinb = 3;       // GFS_GICP_LIN_MINB: CTAs per SM k_linearize is compiled for (3 = no register cap, 4 = 128 registers, 5 = 96)
  int trackCalls = 0;    // gfs_gicp_track_*: calls since the last reset (the new cloud goes to slot trackCalls & 1)
  int trackSeqs = 0;
  // optional per-stage CUDA-event timing of one call (gfs_gicp_set_profiling): an event after every stage, read back at
  // the end of the call; stage = GFS_GICP_STAGE_* index
  bool profiling = false;
  std::vector<cudaEvent_t> evPool;
  std::vector<int> evStage;
  size_t evUsed = 0;
  float stageMs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int stageLaunches[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

// stage indices of gfs_gicp_get_profile: grouping + voxel means + cell pack, 10-NN covariances, correspondence search,
// linearisation, LM bookkeeping (partial-sum reduction, error evaluations, decisions incl. the host's control reads)
enum { ST_GROUP = 0, ST_KNN = 1, ST_NN = 2, ST_LIN = 3, ST_LM = 4 };

static void prof_begin(GfsGicp* h, cudaStream_t st) {
  if (!h->profiling) return;
  h->evUsed = 0;
  h->evStage.clear();
  for (int k = 0; k < 8; k++) { h->stageMs[k] = 0; h->stageLaunches[k] = 0; }
  if (h->evPool.empty()) { h->evPool.resize(1); cudaEventCreate(&h->evPool[0]); }
  cudaEventRecord(h->evPool[0], st);
  h->evUsed = 1;
}
// everything enqueued since the previous mark belongs to `stage`
static void prof_mark(GfsGicp* h, cudaStream_t st, int stage, int launches = 1) {
  if (!h->profiling) return;
  if (h->evUsed == h->evPool.size()) { h->evPool.push_back(nullptr); cudaEventCreate(&h->evPool.back()); }
  cudaEventRecord(h->evPool[h->evUsed++], st);
  h->evStage.push_back(stage);
  h->stageLaunches[stage] += launches;
}
static void prof_end(GfsGicp* h, cudaStream_t st) {
  if (!h->profiling || h->evUsed < 2) return;
  cudaStreamSynchronize(st);
  for (size_t i = 1; i < h->evUsed; i++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->evPool[i - 1], h->evPool[i]) == cudaSuccess) h->stageMs[h->evStage[i - 1]] += ms;
  }
}

extern "C" {

void gfs_gicp_default_setting(GfsGicpSetting* s) {
  if (!s) return;
  s->downsampling_resolution = 0.02;      // RegistrationGICP.cc:11
  s->max_correspondence_distance = 0.1;   // RegistrationGICP.cc:12-14
  s->rotation_eps = 0.1 * M_PI / 180.0;   // registration_helper.hpp:44
  s->translation_eps = 1e-3;              // :45
  s->num_neighbors = 10;                  // registration_helper.cpp:59-60
  s->max_iterations = 20;                 // registration_helper.hpp:47
  s->num_threads = 4;                     // RegistrationGICP.cc:10 (ignored on the GPU)
}

int gfs_gicp_create(const GfsGicpSetting* setting, int max_points, int max_pairs, GfsGicp** out) {
  GFS_REQUIRE(out, GFS_ERR_INVALID, "out is null");
  *out = nullptr;
  GFS_REQUIRE(max_points > 0 && max_pairs > 0, GFS_ERR_INVALID, "bad capacity");
  GfsGicpSetting s;
  gfs_gicp_default_setting(&s);
  if (setting) s = *setting;
  GFS_REQUIRE(s.downsampling_resolution > 0 && s.max_correspondence_distance > 0 && s.max_iterations >= 0, GFS_ERR_INVALID,
              "bad setting");
  GFS_REQUIRE(s.num_neighbors == 10, GFS_ERR_INVALID, "num_neighbors must be 10 (the value align() hard-codes)");
  int rc = gfs_device_check();
  if (rc) return rc;
  GFS_CUDA(cudaFuncSetAttribute(k_knn_cov, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KNN_SMEM));
  GFS_CUDA(cudaFuncSetAttribute(k_knn_search<6, KS_LIST, KS_CELLS, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  GFS_CUDA(cudaFuncSetAttribute(k_knn_search<7, KS_LIST, KS_CELLS, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  GFS_CUDA(cudaFuncSetAttribute(k_knn_search<6, KD_LIST, KD_CELLS, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  GFS_CUDA(cudaFuncSetAttribute(k_knn_search<8, KD_LIST, KD_CELLS, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  GFS_CUDA(cudaFuncSetAttribute(k_knn_cov_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KC_SMEM));
  GFS_CUDA(cudaFuncSetAttribute(k_knn_cov_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KW_SMEM));
  GfsGicp* h = new GfsGicp();
  GicpDev& D = h->dev;
  memset(&D, 0, sizeof(D));
  D.nmax = max_points;
  int hs = 64;
  while (hs < 2 * max_points) hs <<= 1;
  D.hsize = hs;
  D.nblk = div_up(max_points, LIN_THREADS);
  D.voxel = s.downsampling_resolution;
  // grid cell = the correspondence radius (the bounded 1-NN search then ends after shells 0-1), never finer than
  // 4 voxels (a 3x3x3 block must hold the 10 nearest neighbours of almost every point).  Measured on the
  // configs[2] clouds: 0.05 m -> 1816 pairs/s, 0.065 -> 2447, 0.08 -> 2445, 0.1 -> 2505 (profiles/r01_summary.md)
  // ... and 10 % above that: with cell == radius the first, unseeded search (R = radius) reaches two cells down an axis whenever the
  // query sits right at a face; 0.11 m measured 4401 pairs/s against 4214 at 0.10 and 3954 at 0.125 (profiles/r02_summary.md)
  D.cell = 1.1 * std::max(s.max_correspondence_distance, 4.0 * s.downsampling_resolution);
  if (const char* ce = getenv("GFS_GICP_CELL")) {  // tuning knob (metres); any value gives the same results
    const double v = atof(ce);
    if (v > 0) D.cell = v;
  }
  D.max_dist = s.max_correspondence_distance;
  D.nnBoundA = (float)(1.2 * 2.0 * 1.7320508 * 5.9604645e-8 * (2.0 * D.cell + s.max_correspondence_distance));
  D.cbase = 0; D.cstep = 1; D.swap = 0; D.inputAll = 0;
  D.listRead = -1; D.listWrite = -1; D.listStride = max_pairs;
  D.cellOrder = 1;
  D.linPpt = 8;
  D.errPpt = 8;
  if (const char* e = getenv("GFS_GICP_ERR_PPT")) { const int v = atoi(e) & ~1; D.errPpt = v < 2 ? 2 : (v > 32 ? 32 : v); }
  if (const char* e = getenv("GFS_GICP_LIN_PPT")) { const int v = atoi(e); D.linPpt = v < 1 ? 1 : (v > 32 ? 32 : v); }
  if (const char* e = getenv("GFS_GICP_ORDER")) D.cellOrder = atoi(e) != 0;  // 0: queries in point order (first generation)
  if (const char* e = getenv("GFS_GICP_LIN_MINB")) h->linMinb = atoi(e);
  if (const char* e = getenv("GFS_GICP_NN")) h->nnMode = atoi(e);
  if (const char* e = getenv("GFS_GICP_KNN")) h->knnMode = atoi(e);
  D.octSorted = (h->knnMode == 1 || h->nnMode == 3 || h->nnMode == 4) ? 1 : 0;
  h->cellKnn = getenv("GFS_GICP_KNN_CELLS") != nullptr;
  // the dense sorted grid serves the default kernels (k_knn_search, k_nn_corr3); the earlier generations keep the hash grid
  D.dense = (D.cellOrder && h->knnMode >= 2 && h->nnMode >= 7 && !h->cellKnn) ? 1 : 0;
  if (const char* e = getenv("GFS_GICP_GRID")) D.dense = D.dense && atoi(e) != 0;
  D.dcap = 1 << 20;   // 101^3 cells: an 11 m cube at the default 0.11 m cell
  if (const char* e = getenv("GFS_GICP_DENSE_CAP")) { const int v = atoi(e); if (v >= 64) D.dcap = v; }
  D.daxis = (int)floor(cbrt((double)D.dcap) + 1e-9);
  D.dstride = ((D.dcap + 1 + 3) / 4) * 4;
  D.rot_eps = s.rotation_eps;
  D.trans_eps = s.translation_eps;
  D.k = s.num_neighbors;
  D.max_iter = s.max_iterations;
  h->maxPairs = max_pairs;
  const size_t C = 2 * (size_t)max_pairs, P = max_pairs, N = max_points, H = hs;
#define RES(buf, bytes, field, type)            \
  if ((rc = h->buf.reserve(bytes))) {           \
    delete h;                                   \
    return rc;                                  \
  }                                             \
  D.field = (type)h->buf.p;
  RES(b_keys, C * H * 8, keys, unsigned long long*)
  RES(b_minIdx, C * H * 4, minIdx, int*)
  RES(b_count, C * H * 4, count, int*)
  RES(b_start, C * H * 4, start, int*)
  RES(b_cursor, C * H * 4, cursor, int*)
  RES(b_rank, C * H * 4, rank, int*)
  RES(b_slotOf, C * N * 4, slotOf, int*)
  RES(b_members, C * N * 4, members, int*)
  RES(b_nIn, C * 4, nIn, int*)
  RES(b_nDown, C * 4, nDown, int*)
  RES(b_nCells, C * 4, nCells, int*)
  RES(b_nFall, C * 4, nFall, int*)
  RES(b_box, C * 6 * 4, cellBox, int*)
  RES(b_pts, C * N * 32, pts, double*)
  RES(b_cov, C * N * 48, cov, double*)
  RES(b_tab, C * H * 16, tab, uint4*)
  RES(b_rec, C * N * 32, rec, double*)
  RES(b_recf, C * N * 16, recf, float4*)
  RES(b_oct, C * H * 8, oct, uint2*)
  RES(b_nbr, C * N * 40, nbr, int*)
  RES(b_nbrR2, C * N * 8, nbrR2, double*)
  RES(b_knnList, C * N * 64, knnList, int*)
  RES(b_knnCnt, C * N, knnCnt, unsigned char*)
  RES(b_corr, P * N * 4, corr, int*)
  RES(b_maha, P * N * 72, maha, double*)
  RES(b_partial, P * D.nblk * (LIN_THREADS / 32) * RED_N * 8, partial, double*)
  RES(b_partialE, P * D.nblk * 8, partialE, double*)
  RES(b_state, P * LM_STATE * 8, state, double*)
  RES(b_istate, P * LM_ISTATE * 4, istate, int*)
  RES(b_counters, 16, counters, int*)
  RES(b_tickets, P * 8, tickets, int*)
  RES(b_active, 2 * P * 4, activeList, int*)
  RES(b_dS, (D.dense ? C * (size_t)D.dstride : 4) * 4, dS, unsigned*)
  RES(b_dSum, C * 3 * 8, dSum, unsigned long long*)
#undef RES
  if ((rc = h->h_counters.reserve(16))) { delete h; return rc; }
  *out = h;
  return GFS_OK;
}

int gfs_gicp_destroy(GfsGicp* h) {
  if (!h) return GFS_OK;
  DevBuf* d[] = {&h->b_keys, &h->b_minIdx, &h->b_count, &h->b_start, &h->b_cursor, &h->b_rank, &h->b_slotOf, &h->b_members,
                 &h->b_nIn, &h->b_nDown, &h->b_nCells, &h->b_nFall, &h->b_box, &h->b_pts, &h->b_cov, &h->b_tab, &h->b_rec, &h->b_recf, &h->b_oct, &h->b_knnList, &h->b_knnCnt, &h->b_nbr, &h->b_nbrR2, &h->b_corr, &h->b_maha, &h->b_partial,
                 &h->b_partialE, &h->b_state, &h->b_istate, &h->b_counters, &h->b_tickets, &h->b_dS, &h->b_dSum, &h->b_active, &h->b_tgt, &h->b_src, &h->b_n, &h->b_T0, &h->b_res};
  for (DevBuf* b : d) b->release();
  h->h_counters.release();
  for (cudaEvent_t e : h->evPool) cudaEventDestroy(e);
  delete h;
  return GFS_OK;
}

int gfs_gicp_last_launches(const GfsGicp* h) { return h ? h->launches : GFS_ERR_INVALID; }

int gfs_gicp_set_profiling(GfsGicp* h, int on) {
  GFS_REQUIRE(h, GFS_ERR_INVALID, "null handle");
  h->profiling = on != 0;
  return GFS_OK;
}
int gfs_gicp_get_profile(const GfsGicp* h, float* ms8, int* launches8) {
  GFS_REQUIRE(h && ms8, GFS_ERR_INVALID, "null argument");
  for (int k = 0; k < 8; k++) { ms8[k] = h->stageMs[k]; if (launches8) launches8[k] = h->stageLaunches[k]; }
  return GFS_OK;
}

static int group_build(GfsGicp* h, const GicpDev& D, cudaStream_t st, int mode, int clouds, const float* tgt, const float* src,
                       int stride, int* nGroupsOut) {
  const long long slots = (long long)clouds * D.hsize;
  k_group_clear<<<(unsigned)((slots + 255) / 256), 256, 0, st>>>(D, clouds);
  k_group_insert<<<dim3(div_up(D.nmax, 256), clouds), 256, 0, st>>>(D, mode, tgt, src, stride);
  k_group_rank<<<clouds, 1024, 0, st>>>(D, mode, nGroupsOut);
  k_group_fill<<<dim3(div_up(D.nmax, 256), clouds), 256, 0, st>>>(D, mode);
  h->launches += 4;
  return GFS_OK;
}

// Voxel downsampling, k-NN grid and covariances of `clouds` clouds (those D.cbase / D.cstep select).
static int preprocess_clouds(GfsGicp* h, const GicpDev& D, cudaStream_t st, int clouds, const float* d_target, const float* d_source,
                             int stride) {
  prof_begin(h, st);
  group_build(h, D, st, 0, clouds, d_target, d_source, stride, D.nDown);
  k_voxel_mean<<<dim3(div_up(D.nmax, 256), clouds), 256, 0, st>>>(D, d_target, d_source, stride);
  if (D.dense) {
    const dim3 gp(div_up(D.nmax, 256), clouds);
    k_dense_init<<<div_up(clouds, 128), 128, 0, st>>>(D, clouds);
    k_dense_box<<<dim3(div_up(D.nmax, 1024), clouds), 256, 0, st>>>(D);
    k_dense_region<<<clouds, 1024, 0, st>>>(D);
    k_dense_count<<<gp, 256, 0, st>>>(D);
    k_dense_scan<<<clouds, 1024, 0, st>>>(D);
    k_dense_fill<<<gp, 256, 0, st>>>(D);
    h->launches += 6 - 5;
  } else {
    // grid over the downsampled points (reuses the hash-table storage); group count is not needed
    group_build(h, D, st, 1, clouds, nullptr, nullptr, 0, D.nCells);
    if (D.octSorted) {
      const long long slots = (long long)clouds * D.hsize;
      k_cell_sort<<<(unsigned)((slots + 255) / 256), 256, 0, st>>>(D, clouds);
      h->launches += 1;
    }
    const long long work = (long long)clouds * std::max(D.hsize, D.nmax);
    k_cell_pack<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(D, clouds);
  }
  prof_mark(h, st, ST_GROUP, 11);
  if (h->knnMode == 1 && !h->cellKnn) {
    // warp per cell; what it cannot finish (sparse surroundings, overfull blocks, ties) goes to the per-query kernel
    k_knn_cov_warp<<<dim3(2 * 148, clouds), KW_WARPS * 32, KW_SMEM, st>>>(D);
    k_knn_select_cov<<<dim3(div_up(D.nmax, 128), clouds), 128, 0, st>>>(D);
    k_knn_cov<<<dim3(8, clouds), KNN_THREADS, KNN_SMEM, st>>>(D, 1);
    h->launches += 2;
  } else if (h->knnMode >= 2 && !h->cellKnn && D.cellOrder) {
    const dim3 gk(div_up(D.nmax, KNN_THREADS), clouds);
    if (D.dense) {
      if (h->knnMode == 3) k_knn_search<8, KD_LIST, KD_CELLS, true><<<gk, KNN_THREADS, KD_SMEM, st>>>(D);
      else if (h->knnMode == 4) k_knn_search<5, KD_LIST, KD_CELLS, true><<<gk, KNN_THREADS, KD_SMEM, st>>>(D);
      else k_knn_search<6, KD_LIST, KD_CELLS, true><<<gk, KNN_THREADS, KD_SMEM, st>>>(D);
    } else if (h->knnMode == 2) k_knn_search<6, KS_LIST, KS_CELLS, false><<<gk, KNN_THREADS, KS_SMEM, st>>>(D);
    else if (h->knnMode == 3) k_knn_search<7, KS_LIST, KS_CELLS, false><<<gk, KNN_THREADS, KS_SMEM, st>>>(D);
    else if (h->knnMode == 4) k_knn_search<5, KS_LIST, KS_CELLS, false><<<gk, KNN_THREADS, KS_SMEM, st>>>(D);
    else k_knn_search<4, KS_LIST, KS_CELLS, false><<<gk, KNN_THREADS, KS_SMEM, st>>>(D);
    k_cov_nbr<<<dim3(div_up(D.nmax, 128), clouds), 128, 0, st>>>(D);
    h->launches += 1;
  } else if (!h->cellKnn) {
    k_knn_cov<<<dim3(div_up(D.nmax, KNN_THREADS), clouds), KNN_THREADS, KNN_SMEM, st>>>(D, 0);
  } else {
    // ~3 resident CTAs per SM in flight over the whole batch; every CTA strides over its cloud's cells
    const int gx = std::min(std::max(div_up(148 * 6, clouds), 24), div_up(D.nmax, KC_THREADS / KC_LANES));
    k_knn_cov_cells<<<dim3(gx, clouds), KC_THREADS, KC_SMEM, st>>>(D);
    k_knn_cov<<<dim3(8, clouds), KNN_THREADS, KNN_SMEM, st>>>(D, 1);
    h->launches += 1;
  }
  prof_mark(h, st, ST_KNN, h->knnMode == 1 && !h->cellKnn ? 3 : (h->cellKnn ? 2 : 1));
  h->launches += 3;
  GFS_CUDA(cudaGetLastError());
  return GFS_OK;
}

// LevenbergMarquardtOptimizer::optimize for every pair (registration/optimizer.hpp:83-148)
static int optimize_pairs(GfsGicp* h, const GicpDev& D, cudaStream_t st, int pairs, const double* d_T0, GfsGicpResult* d_out) {
  k_lm_init<<<div_up(pairs, 128), 128, 0, st>>>(D, pairs, d_T0);
  h->launches += 1;
  int* hc = (int*)h->h_counters.p;
  // Rounds of [search, linearize, begin, error, decide]; every pair advances through its own state machine (see I_OUTER).  A pair
  // needs one round per lambda trial, i.e. at least one per outer iteration; the host only looks every few rounds (after the
  // 3rd, then every 2nd: most pairs converge within 3-8 iterations) whether any pair is left.  GFS_GICP_CHECK_EVERY=1 gives the
  // round-1 behaviour of one host synchronisation per round.
  static const int checkEvery = [] { const char* e = getenv("GFS_GICP_CHECK_EVERY"); const int v = e ? atoi(e) : 2; return v > 0 ? v : 2; }();
  const int maxRounds = D.max_iter * 10;  // every outer iteration may take up to 10 trials (optimizer.hpp:107)
  int nextCheck = checkEvery == 1 ? 0 : 2;
  GicpDev L = D;            // per-launch copy: which list of unfinished pairs the kernels read / write
  L.listRead = -1; L.listWrite = -1; L.listStride = h->maxPairs;
  int rows = pairs;         // CTA rows = unfinished pairs as of the last check
  for (int round = 0; round < maxRounds; round++) {
    const dim3 gn(div_up(D.nmax, NN_THREADS), rows);
    if (h->nnMode == 0) k_nn_corr<<<gn, NN_THREADS, 0, st>>>(L);
    else if (h->nnMode == 1) k_nn_corr2<false, false, false><<<gn, NN_THREADS, 0, st>>>(L);
    else if (h->nnMode == 2) k_nn_corr2<true, false, false><<<gn, NN_THREADS, 0, st>>>(L);
    else if (h->nnMode == 3) k_nn_corr2<false, true, false><<<gn, NN_THREADS, 0, st>>>(L);
    else if (h->nnMode == 4) k_nn_corr2<true, true, false><<<gn, NN_THREADS, 0, st>>>(L);
    else if (h->nnMode == 5) k_nn_corr2<false, false, true><<<gn, NN_THREADS, 0, st>>>(L);
    else if (h->nnMode == 6) k_nn_corr2<true, false, true><<<gn, NN_THREADS, 0, st>>>(L);
    else if (D.dense) {
      if (h->nnMode == 8) k_nn_corr3<8, true><<<gn, NN_THREADS, 0, st>>>(L);
      else if (h->nnMode == 9) k_nn_corr3<7, true><<<gn, NN_THREADS, 0, st>>>(L);
      else k_nn_corr3<6, true><<<gn, NN_THREADS, 0, st>>>(L);
    }
    else if (h->nnMode == 8) k_nn_corr3<8, false><<<gn, NN_THREADS, 0, st>>>(L);
    else if (h->nnMode == 9) k_nn_corr3<7, false><<<gn, NN_THREADS, 0, st>>>(L);
    else if (h->nnMode == 10) k_nn_corr3<5, false><<<gn, NN_THREADS, 0, st>>>(L);
    else k_nn_corr3<6, false><<<gn, NN_THREADS, 0, st>>>(L);
    prof_mark(h, st, ST_NN);
    const dim3 gl(div_up(D.nblk, D.linPpt), rows);   // + the pair's LM begin in its last block
    if (h->linMinb == 4) k_linearize<4><<<gl, LIN_THREADS, 0, st>>>(L);
    else if (h->linMinb == 5) k_linearize<5><<<gl, LIN_THREADS, 0, st>>>(L);
    else k_linearize<3><<<gl, LIN_THREADS, 0, st>>>(L);
    prof_mark(h, st, ST_LIN);
    const bool check = round >= nextCheck;
    if (check) GFS_CUDA(cudaMemsetAsync(D.counters, 0, 8, st));
    L.listWrite = check ? (L.listRead < 0 ? 0 : 1 - L.listRead) : -1;
    k_error<<<dim3(div_up(D.nblk, D.errPpt), rows), LIN_THREADS, 0, st>>>(L);       // + the pair's LM decision in its last block
    h->launches += 3;
    if (check) {
      GFS_CUDA(cudaMemcpyAsync(hc, D.counters, 8, cudaMemcpyDeviceToHost, st));
      GFS_CUDA(gfs::stream_wait(st));
      prof_mark(h, st, ST_LM, 1);
      if (hc[1] == 0) break;  // every pair converged / failed / hit max_iterations
      rows = hc[1];
      L.listRead = L.listWrite;
      nextCheck = round + checkEvery;
    } else {
      prof_mark(h, st, ST_LM, 1);
    }
  }
  k_gicp_result<<<div_up(pairs, 128), 128, 0, st>>>(D, pairs, d_out);
  h->launches += 1;
  prof_mark(h, st, ST_LM);
  prof_end(h, st);
  GFS_CUDA(cudaGetLastError());
  return GFS_OK;
}

int gfs_gicp_align_batch_device(GfsGicp* h, void* stream, const float* d_target, const int* d_nt, const float* d_source,
                                const int* d_ns, int pairs, int stride, const double* d_T0, GfsGicpResult* d_out) {
  GFS_REQUIRE(h, GFS_ERR_INVALID, "null handle");
  GFS_REQUIRE(d_target && d_nt && d_source && d_ns && d_T0 && d_out, GFS_ERR_INVALID, "null pointer");
  GFS_REQUIRE(pairs > 0 && pairs <= h->maxPairs, GFS_ERR_CAPACITY, "pairs exceeds the handle's max_pairs");
  GFS_REQUIRE(stride > 0 && stride <= h->dev.nmax, GFS_ERR_CAPACITY, "stride exceeds the handle's max_points");
  cudaStream_t st = (cudaStream_t)stream;
  GicpDev D = h->dev;
  D.cbase = 0; D.cstep = 1; D.swap = 0; D.inputAll = 0;
  h->trackCalls = 0;  // the clouds of a tracked sequence are overwritten
  h->launches = 0;
  k_set_counts<<<div_up(pairs, 128), 128, 0, st>>>(D, pairs, d_nt, d_ns, stride);
  h->launches += 1;
  int rc = preprocess_clouds(h, D, st, 2 * pairs, d_target, d_source, stride);
  if (rc) return rc;
  return optimize_pairs(h, D, st, pairs, d_T0, d_out);
}

// ---- tracking mode: one new cloud per sequence and call (Tracking::PredictStateICP, reference src/Tracking.cc:3364-3413,
// registers the current frame's cloud -- the source -- against the last frame's -- the target).  small_gicp::align
// preprocesses both clouds on every call (registration_helper.cpp:23-35, 56-68), so a frame's cloud is downsampled,
// indexed and given covariances twice: once as the source, once more one frame later as the target.  Preprocessing is a
// function of the cloud alone, so here each cloud is preprocessed ONCE and stays in HBM for the next call; the results
// are those of gfs_gicp_align_batch on the pairs (cloud[k-1], cloud[k]) bit for bit (tests/test_gpu_gicp.py).
int gfs_gicp_track_reset(GfsGicp* h) {
  GFS_REQUIRE(h, GFS_ERR_INVALID, "null handle");
  h->trackCalls = 0;
  h->trackSeqs = 0;
  return GFS_OK;
}

int gfs_gicp_track_calls(const GfsGicp* h) { return h ? h->trackCalls : GFS_ERR_INVALID; }

int gfs_gicp_track_batch_device(GfsGicp* h, void* stream, const float* d_cloud, const int* d_n, int seqs, int stride,
                                const double* d_T0, GfsGicpResult* d_out) {
  GFS_REQUIRE(h, GFS_ERR_INVALID, "null handle");
  GFS_REQUIRE(d_cloud && d_n, GFS_ERR_INVALID, "null pointer");
  GFS_REQUIRE(seqs > 0 && seqs <= h->maxPairs, GFS_ERR_CAPACITY, "seqs exceeds the handle's max_pairs");
  GFS_REQUIRE(stride > 0 && stride <= h->dev.nmax, GFS_ERR_CAPACITY, "stride exceeds the handle's max_points");
  GFS_REQUIRE(h->trackCalls == 0 || seqs == h->trackSeqs, GFS_ERR_INVALID, "the number of sequences changed: call gfs_gicp_track_reset first");
  GFS_REQUIRE(h->trackCalls == 0 || (d_T0 && d_out), GFS_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  GicpDev D = h->dev;
  const int slot = h->trackCalls & 1;
  D.cbase = slot; D.cstep = 2; D.swap = 1 - slot; D.inputAll = 1;
  h->launches = 0;
  k_set_counts<<<div_up(seqs, 128), 128, 0, st>>>(D, seqs, nullptr, d_n, stride);
  h->launches += 1;
  int rc = preprocess_clouds(h, D, st, seqs, nullptr, d_cloud, stride);
  if (rc) return rc;
  const bool first = h->trackCalls == 0;
  h->trackCalls++;
  h->trackSeqs = seqs;
  if (first) { prof_end(h, st); return GFS_OK; }  // nothing to register the first cloud against
  return optimize_pairs(h, D, st, seqs, d_T0, d_out);
}

int gfs_gicp_track_batch(GfsGicp* h, void* stream, const float* cloud, const int* n, int seqs, int stride, const double* T0,
                         GfsGicpResult* out) {
  GFS_REQUIRE(h, GFS_ERR_INVALID, "null handle");
  GFS_REQUIRE(cloud && n, GFS_ERR_INVALID, "null pointer");
  GFS_REQUIRE(seqs > 0 && seqs <= h->maxPairs, GFS_ERR_CAPACITY, "seqs exceeds the handle's max_pairs");
  GFS_REQUIRE(stride > 0 && stride <= h->dev.nmax, GFS_ERR_CAPACITY, "stride exceeds the handle's max_points");
  const bool first = h->trackCalls == 0;
  GFS_REQUIRE(first || (T0 && out), GFS_ERR_INVALID, "null pointer");
  for (int i = 0; i < seqs; i++) GFS_REQUIRE(n[i] >= 0 && n[i] <= stride, GFS_ERR_INVALID, "point count outside [0, stride]");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t P = seqs, cb = P * stride * 16;
  int rc;
  if ((rc = h->b_src.reserve(cb))) return rc;
  if ((rc = h->b_n.reserve(P * 8))) return rc;
  if ((rc = h->b_T0.reserve(P * 128))) return rc;
  if ((rc = h->b_res.reserve(P * sizeof(GfsGicpResult)))) return rc;
  GFS_CUDA(cudaMemcpyAsync(h->b_src.p, cloud, cb, cudaMemcpyHostToDevice, st));
  GFS_CUDA(cudaMemcpyAsync(h->b_n.p, n, P * 4, cudaMemcpyHostToDevice, st));
  if (!first) GFS_CUDA(cudaMemcpyAsync(h->b_T0.p, T0, P * 128, cudaMemcpyHostToDevice, st));
  rc = gfs_gicp_track_batch_device(h, stream, (const float*)h->b_src.p, (const int*)h->b_n.p, seqs, stride,
                                   (const double*)h->b_T0.p, (GfsGicpResult*)h->b_res.p);
  if (rc) return rc;
  if (!first) GFS_CUDA(cudaMemcpyAsync(out, h->b_res.p, P * sizeof(GfsGicpResult), cudaMemcpyDeviceToHost, st));
  GFS_CUDA(gfs::stream_wait(st));
  return GFS_OK;
}

int gfs_gicp_align_batch(GfsGicp* h, void* stream, const float* target, const int* nt, const float* source, const int* ns,
                         int pairs, int stride, const double* T0, GfsGicpResult* out) {
  GFS_REQUIRE(h, GFS_ERR_INVALID, "null handle");
  GFS_REQUIRE(target && nt && source && ns && T0 && out, GFS_ERR_INVALID, "null pointer");
  GFS_REQUIRE(pairs > 0 && pairs <= h->maxPairs, GFS_ERR_CAPACITY, "pairs exceeds the handle's max_pairs");
  GFS_REQUIRE(stride > 0 && stride <= h->dev.nmax, GFS_ERR_CAPACITY, "stride exceeds the handle's max_points");
  for (int i = 0; i < pairs; i++)
    GFS_REQUIRE(nt[i] >= 0 && nt[i] <= stride && ns[i] >= 0 && ns[i] <= stride, GFS_ERR_INVALID, "point count outside [0, stride]");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t P = pairs, cb = P * stride * 16;
  int rc;
  if ((rc = h->b_tgt.reserve(cb))) return rc;
  if ((rc = h->b_src.reserve(cb))) return rc;
  if ((rc = h->b_n.reserve(P * 8))) return rc;
  if ((rc = h->b_T0.reserve(P * 128))) return rc;
  if ((rc = h->b_res.reserve(P * sizeof(GfsGicpResult)))) return rc;
  GFS_CUDA(cudaMemcpyAsync(h->b_tgt.p, target, cb, cudaMemcpyHostToDevice, st));
  GFS_CUDA(cudaMemcpyAsync(h->b_src.p, source, cb, cudaMemcpyHostToDevice, st));
  GFS_CUDA(cudaMemcpyAsync(h->b_n.p, nt, P * 4, cudaMemcpyHostToDevice, st));
  GFS_CUDA(cudaMemcpyAsync((int*)h->b_n.p + P, ns, P * 4, cudaMemcpyHostToDevice, st));
  GFS_CUDA(cudaMemcpyAsync(h->b_T0.p, T0, P * 128, cudaMemcpyHostToDevice, st));
  rc = gfs_gicp_align_batch_device(h, stream, (const float*)h->b_tgt.p, (const int*)h->b_n.p, (const float*)h->b_src.p,
                                   (const int*)h->b_n.p + P, pairs, stride, (const double*)h->b_T0.p,
                                   (GfsGicpResult*)h->b_res.p);
  if (rc) return rc;
  GFS_CUDA(cudaMemcpyAsync(out, h->b_res.p, P * sizeof(GfsGicpResult), cudaMemcpyDeviceToHost, st));
  GFS_CUDA(gfs::stream_wait(st));
  return GFS_OK;
}

int gfs_gicp_align(GfsGicp* h, void* stream, const float* target, int nt, const float* source, int ns, const double* T0,
                   GfsGicpResult* out) {
  GFS_REQUIRE(h && out, GFS_ERR_INVALID, "null handle/output");
  GFS_REQUIRE(nt >= 0 && ns >= 0 && nt <= h->dev.nmax && ns <= h->dev.nmax, GFS_ERR_CAPACITY, "cloud larger than max_points");
  GFS_REQUIRE((target || nt == 0) && (source || ns == 0) && T0, GFS_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int stride = std::max(std::max(nt, ns), 1);
  int rc;
  if ((rc = h->b_tgt.reserve((size_t)stride * 16))) return rc;
  if ((rc = h->b_src.reserve((size_t)stride * 16))) return rc;
  if ((rc = h->b_n.reserve(8))) return rc;
  if ((rc = h->b_T0.reserve(128))) return rc;
  if ((rc = h->b_res.reserve(sizeof(GfsGicpResult)))) return rc;
  if (nt) GFS_CUDA(cudaMemcpyAsync(h->b_tgt.p, target, (size_t)nt * 16, cudaMemcpyHostToDevice, st));
  if (ns) GFS_CUDA(cudaMemcpyAsync(h->b_src.p, source, (size_t)ns * 16, cudaMemcpyHostToDevice, st));
  const int n2[2] = {nt, ns};
  GFS_CUDA(cudaMemcpyAsync(h->b_n.p, n2, 8, cudaMemcpyHostToDevice, st));
  GFS_CUDA(cudaMemcpyAsync(h->b_T0.p, T0, 128, cudaMemcpyHostToDevice, st));
  rc = gfs_gicp_align_batch_device(h, stream, (const float*)h->b_tgt.p, (const int*)h->b_n.p, (const float*)h->b_src.p,
                                   (const int*)h->b_n.p + 1, 1, stride, (const double*)h->b_T0.p, (GfsGicpResult*)h->b_res.p);
  if (rc) return rc;
  GFS_CUDA(cudaMemcpyAsync(out, h->b_res.p, sizeof(GfsGicpResult), cudaMemcpyDeviceToHost, st));
  GFS_CUDA(gfs::stream_wait(st));
  return GFS_OK;
}

int gfs_gicp_get_knn_stats(GfsGicp* h, void* stream, int cloud, int* n_cells, int* n_per_query) {
  GFS_REQUIRE(h && n_cells && n_per_query && cloud >= 0 && cloud < 2 * h->maxPairs, GFS_ERR_INVALID, "bad handle/cloud");
  GFS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  GFS_CUDA(cudaMemcpy(n_cells, h->dev.nCells + cloud, 4, cudaMemcpyDeviceToHost));
  GFS_CUDA(cudaMemcpy(n_per_query, h->dev.nFall + cloud, 4, cudaMemcpyDeviceToHost));
  return GFS_OK;
}

// Parity hook: downsampled points (xyz) and covariances (xx xy xz yy yz zz) of one cloud of the last
// batch (cloud = 2*pair for the target, 2*pair+1 for the source), copied to the host.
int gfs_gicp_get_cloud(GfsGicp* h, void* stream, int cloud, double* out_xyz, double* out_cov6, int cap, int* n) {
  GFS_REQUIRE(h && n && cloud >= 0 && cloud < 2 * h->maxPairs, GFS_ERR_INVALID, "bad handle/cloud");
  cudaStream_t st = (cudaStream_t)stream;
  GFS_CUDA(gfs::stream_wait(st));
  int m = 0;
  GFS_CUDA(cudaMemcpy(&m, h->dev.nDown + cloud, 4, cudaMemcpyDeviceToHost));
  *n = m;
  const int k = std::min(m, cap);
  if (k > 0 && out_xyz) {
    std::vector<double> tmp((size_t)k * 4);
    GFS_CUDA(cudaMemcpy(tmp.data(), h->dev.pts + (size_t)cloud * h->dev.nmax * 4, tmp.size() * 8, cudaMemcpyDeviceToHost));
    for (int i = 0; i < k; i++) { out_xyz[3 * i] = tmp[4 * i]; out_xyz[3 * i + 1] = tmp[4 * i + 1]; out_xyz[3 * i + 2] = tmp[4 * i + 2]; }
  }
  if (k > 0 && out_cov6)
    GFS_CUDA(cudaMemcpy(out_cov6, h->dev.cov + (size_t)cloud * h->dev.nmax * 6, (size_t)k * 48, cudaMemcpyDeviceToHost));
  return GFS_OK;
}
}
