// Descriptor matching on sm_100a: brute-force 256-bit Hamming argmin and the GMS grid filter.
//
// Replaces, for whole batches of frame pairs,
//   * cv::BFMatcher(NORM_HAMMING).match(d1, d2)          reference call sites src/ORBmatcher.cc:755-756, 805-806, 888-889
//   * ORBmatcher::DescriptorDistance                      src/ORBmatcher.cc:2536-2550
//   * gms_matcher::GetInlierMask(mask, false, false)      Thirdparty/GMS/include/gms_matcher.h:236-246 -> run(1) :385-419
// Integer work throughout: results are bit-exact against oracle/match_oracle.cpp.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace gfs {

// ------------------------------------------------------------------------------------------------
// k_bf_hamming: CTA = 8 warps; the pair's train descriptors are staged once per CTA in shared
// memory, word-major (T[w][row]) so a warp's 32 lanes read 32 consecutive words; each warp keeps
// 4 query descriptors in registers and sweeps the train rows lane-strided.  Ties resolve to the
// lowest train index (each lane scans its rows in increasing order with '<', the warp reduction
// orders by (distance, index)).
// ------------------------------------------------------------------------------------------------
static const int BF_WARPS = 8;
static const int BF_QPW = 4;                         // queries per warp per sweep
static const int BF_QPB = BF_WARPS * BF_QPW * 2;     // queries per CTA (two sweeps)

__global__ void __launch_bounds__(BF_WARPS * 32) k_bf_hamming(const uint8_t* __restrict__ dq, const int* __restrict__ nq,
                                                              const uint8_t* __restrict__ dt, const int* __restrict__ nt,
                                                              int stride, int* __restrict__ out_idx,
                                                              int* __restrict__ out_dist) {
  extern __shared__ uint32_t T[];  // [8][tp]
  const int pair = blockIdx.y;
  const int nQ = nq[pair], nT = nt[pair];
  const int q0 = blockIdx.x * BF_QPB;
  if (q0 >= nQ) return;
  const int tp = (nT + 31) & ~31;
  const uint32_t* gt = (const uint32_t*)(dt + (size_t)pair * stride * 32);
  for (int i = threadIdx.x; i < tp * 8; i += blockDim.x) {
    const int row = i >> 3, w = i & 7;
    T[w * tp + row] = (row < nT) ? __ldg(gt + i) : 0u;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t* gq = (const uint32_t*)(dq + (size_t)pair * stride * 32);
  for (int sweep = 0; sweep < 2; sweep++) {
    const int qb = q0 + (sweep * BF_WARPS + warp) * BF_QPW;
    if (qb >= nQ) continue;
    uint32_t qw[BF_QPW][8];
#pragma unroll
    for (int j = 0; j < BF_QPW; j++) {
      const int q = min(qb + j, nQ - 1);
#pragma unroll
      for (int w = 0; w < 8; w++) qw[j][w] = __ldg(gq + (size_t)q * 8 + w);
    }
    int best[BF_QPW], bidx[BF_QPW];
#pragma unroll
    for (int j = 0; j < BF_QPW; j++) { best[j] = 0x7fffffff; bidx[j] = -1; }
    for (int t = lane; t < nT; t += 32) {
      uint32_t tw[8];
#pragma unroll
      for (int w = 0; w < 8; w++) tw[w] = T[w * tp + t];
#pragma unroll
      for (int j = 0; j < BF_QPW; j++) {
        int d = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) d += __popc(qw[j][w] ^ tw[w]);
        if (d < best[j]) { best[j] = d; bidx[j] = t; }
      }
    }
#pragma unroll
    for (int j = 0; j < BF_QPW; j++) {
      // (distance, index) packed so that min() prefers the smaller distance, then the lower index
      unsigned long long v = ((unsigned long long)(unsigned)best[j] << 32) | (unsigned)(bidx[j] < 0 ? 0x7fffffff : bidx[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
      const int q = qb + j;
      if (lane == 0 && q < nQ) {
        const bool none = nT == 0;
        out_idx[(size_t)pair * stride + q] = none ? -1 : (int)(v & 0xffffffffu);
        out_dist[(size_t)pair * stride + q] = none ? -1 : (int)(v >> 32);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_bf_hamming_mma: the same argmin, with the 256-bit XOR-popcount done by the tensor-core binary MMA
// (mma.sync m16n8k256 .b1 xor.popc: one warp instruction = 16 x 8 Hamming distances).  Experimental:
// on sm_100a ptxas lowers the instruction onto IMMA.16832.U8 sequences and the kernel is slower than the
// POPC one (2.73 vs 1.90 ms per 1023 pairs); kept, bit-exact, behind GFS_BF_MMA=1 for comparison.
// CTA = 8 warps x 16 queries; train descriptors in shared memory, row-major.
// ------------------------------------------------------------------------------------------------
static const int BFM_WARPS = 8;
static const int BFM_QPB = BFM_WARPS * 16;

__global__ void __launch_bounds__(BFM_WARPS * 32) k_bf_hamming_mma(const uint8_t* __restrict__ dq, const int* __restrict__ nq,
                                                                   const uint8_t* __restrict__ dt, const int* __restrict__ nt,
                                                                   int stride, int* __restrict__ out_idx,
                                                                   int* __restrict__ out_dist) {
  extern __shared__ uint32_t T[];  // [tp][8]
  const int pair = blockIdx.y;
  const int nQ = nq[pair], nT = nt[pair];
  const int q0 = blockIdx.x * BFM_QPB;
  if (q0 >= nQ) return;
  const int tp = (nT + 7) & ~7;
  const uint32_t* gt = (const uint32_t*)(dt + (size_t)pair * stride * 32);
  for (int i = threadIdx.x; i < tp * 8; i += blockDim.x) T[i] = ((i >> 3) < nT) ? __ldg(gt + i) : 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int qb = q0 + warp * 16;
  if (qb >= nQ) return;
  const uint32_t* gq = (const uint32_t*)(dq + (size_t)pair * stride * 32);
  const int qa = min(qb + g, nQ - 1), qc = min(qb + g + 8, nQ - 1);
  const uint32_t a0 = __ldg(gq + (size_t)qa * 8 + t), a1 = __ldg(gq + (size_t)qc * 8 + t);
  const uint32_t a2 = __ldg(gq + (size_t)qa * 8 + t + 4), a3 = __ldg(gq + (size_t)qc * 8 + t + 4);
  int bestA = 0x7fffffff, idxA = 0x7fffffff, bestB = 0x7fffffff, idxB = 0x7fffffff;
  for (int c = 0; c < tp; c += 8) {
    const uint32_t b0 = T[(c + g) * 8 + t], b1 = T[(c + g) * 8 + t + 4];
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    asm volatile(
        "mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.xor.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    const int j0 = c + 2 * t, j1 = j0 + 1;  // train rows of this thread's two columns (ascending)
    if (j0 < nT) {
      if (c0 < bestA) { bestA = c0; idxA = j0; }
      if (c2 < bestB) { bestB = c2; idxB = j0; }
    }
    if (j1 < nT) {
      if (c1 < bestA) { bestA = c1; idxA = j1; }
      if (c3 < bestB) { bestB = c3; idxB = j1; }
    }
  }
  unsigned long long va = ((unsigned long long)(unsigned)bestA << 32) | (unsigned)idxA;
  unsigned long long vb = ((unsigned long long)(unsigned)bestB << 32) | (unsigned)idxB;
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    va = min(va, __shfl_xor_sync(0xffffffffu, va, o));
    vb = min(vb, __shfl_xor_sync(0xffffffffu, vb, o));
  }
  if (t == 0) {
    const bool none = nT == 0;
    if (qb + g < nQ) {
      out_idx[(size_t)pair * stride + qb + g] = none ? -1 : (int)(va & 0xffffffffu);
      out_dist[(size_t)pair * stride + qb + g] = none ? -1 : (int)(va >> 32);
    }
    if (qb + g + 8 < nQ) {
      out_idx[(size_t)pair * stride + qb + g + 8] = none ? -1 : (int)(vb & 0xffffffffu);
      out_dist[(size_t)pair * stride + qb + g + 8] = none ? -1 : (int)(vb >> 32);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_gms: one CTA per frame pair.  The reference's dense 400x400 motion-statistics table (640 KB,
// cleared four times per call) is replaced by a shared-memory hash table of the <= nm occupied (left cell,
// right cell) keys; every table read becomes a lookup, which is exact.
// ------------------------------------------------------------------------------------------------
static const int GMS_G = 20, GMS_NG = 400, GMS_THREADS = 256;

__device__ __forceinline__ int gms_nb9(int idx, int k) {
  const int x = idx % GMS_G + (k % 3 - 1), y = idx / GMS_G + (k / 3 - 1);
  if (x < 0 || x >= GMS_G || y < 0 || y >= GMS_G) return -1;
  return x + y * GMS_G;
}
// Occupied (left cell, right cell) keys live in an open-addressing hash table in shared memory
// (key -> number of matches): building it is one atomicCAS + two atomicAdd per match, every read of the
// reference's motion-statistics table is a one- or two-probe lookup.  (First version: bitonic sort of the
// keys + binary searches -- 55 barrier-separated stages per grid type, 45 % of the kernel's instructions.)
static const uint32_t GMS_EMPTY = 0xffffffffu;
__device__ __forceinline__ unsigned gms_hash(uint32_t key, int tbits) { return (key * 2654435761u) >> (32 - tbits); }
__device__ __forceinline__ int gms_count(const uint32_t* hkey, const uint32_t* hcnt, int tbits, uint32_t key) {
  const unsigned tmask = (1u << tbits) - 1u;
  unsigned h = gms_hash(key, tbits);
  while (true) {
    const uint32_t k = hkey[h];
    if (k == key) return (int)hcnt[h];
    if (k == GMS_EMPTY) return 0;
    h = (h + 1) & tmask;
  }
}

__global__ void __launch_bounds__(GMS_THREADS) k_gms(const GfsKeyPoint* __restrict__ kp1, const int* __restrict__ n1,
                                                     const GfsKeyPoint* __restrict__ kp2, const int* __restrict__ n2,
                                                     const int* __restrict__ mq, const int* __restrict__ mt,
                                                     const int* __restrict__ nmArr, int stride, int mstride, int tbits,
                                                     int w1, int h1, int w2, int h2, uint8_t* __restrict__ out_inlier,
                                                     int* __restrict__ out_count) {
  extern __shared__ uint32_t gsm[];
  const int tsize = 1 << tbits;
  uint32_t* hkey = gsm;                          // [tsize] l * 400 + r, GMS_EMPTY when free
  uint32_t* hcnt = hkey + tsize;                 // [tsize] matches with that key
  short* ml = (short*)(hcnt + tsize);            // [mstride] left cell of match i (this grid type)
  short* mr = ml + mstride;                      // [mstride] right cell (grid type 1)
  int* nLeft = (int*)(ml + 2 * mstride);        // [400] (2*mstride shorts keep 4-byte alignment)
  int* cellPair = nLeft + GMS_NG;                // [400]
  __shared__ int s_cnt;
  const int pair = blockIdx.x, tid = threadIdx.x;
  const int nm = nmArr ? nmArr[pair] : n1[pair];
  const int N1 = n1[pair], N2 = n2[pair];
  const GfsKeyPoint* K1 = kp1 + (size_t)pair * stride;
  const GfsKeyPoint* K2 = kp2 + (size_t)pair * stride;
  const int* MQ = mq ? mq + (size_t)pair * mstride : nullptr;
  const int* MT = mt + (size_t)pair * mstride;
  uint8_t* mask = out_inlier + (size_t)pair * mstride;
  for (int i = tid; i < nm; i += GMS_THREADS) mask[i] = 0;
  if (tid == 0) s_cnt = 0;
  const unsigned tmask = (unsigned)tsize - 1u;

  for (int type = 1; type <= 4; type++) {
    __syncthreads();
    for (int i = tid; i < tsize; i += GMS_THREADS) { hkey[i] = GMS_EMPTY; hcnt[i] = 0u; }
    for (int c = tid; c < GMS_NG; c += GMS_THREADS) { nLeft[c] = 0; cellPair[c] = 0; }
    __syncthreads();
    for (int i = tid; i < nm; i += GMS_THREADS) {
      const int q = MQ ? MQ[i] : i, t = MT[i];
      int l = -1, r = -1;
      if (q >= 0 && q < N1 && t >= 0 && t < N2) {
        // NormalizePoints :104-115 and GetGridIndexLeft/Right :125-160
        const float lx = __fmul_rn(__fdiv_rn(K1[q].x, (float)w1), (float)GMS_G);
        const float ly = __fmul_rn(__fdiv_rn(K1[q].y, (float)h1), (float)GMS_G);
        const int x = (type == 2 || type == 4) ? (int)floor((double)lx + 0.5) : (int)floorf(lx);
        const int y = (type == 3 || type == 4) ? (int)floor((double)ly + 0.5) : (int)floorf(ly);
        l = (x >= GMS_G || y >= GMS_G) ? -1 : x + y * GMS_G;
        if (type == 1) {
          const float rx = __fmul_rn(__fdiv_rn(K2[t].x, (float)w2), (float)GMS_G);
          const float ry = __fmul_rn(__fdiv_rn(K2[t].y, (float)h2), (float)GMS_G);
          r = (int)floorf(rx) + (int)floorf(ry) * GMS_G;
          // keep -1 / -2 exact (the reference compares raw indices against mCellPairs, :406)
          r = (r < -2) ? -3 : min(r, 32000);
          mr[i] = (short)r;
        } else {
          r = mr[i];
        }
      } else if (type == 1) {
        mr[i] = -1;
      }
      ml[i] = (short)max(-1, min(l, 32000));
      if (l >= 0 && r >= 0 && l < GMS_NG && r < GMS_NG) {
        const uint32_t key = (uint32_t)(l * GMS_NG + r);
        unsigned h = gms_hash(key, tbits);
        while (true) {  // the table is at most 2/3 full: a free or matching slot always turns up
          const uint32_t prev = atomicCAS(&hkey[h], GMS_EMPTY, key);
          if (prev == GMS_EMPTY || prev == key) break;
          h = (h + 1) & tmask;
        }
        atomicAdd(&hcnt[h], 1u);
        atomicAdd(&nLeft[l], 1);
      }
    }
    __syncthreads();
    // per left cell: first-maximum right cell (VerifyCellPairs :338-356) = max of (count, -right cell)
    for (int sidx = tid; sidx < tsize; sidx += GMS_THREADS) {
      const uint32_t k = hkey[sidx];
      if (k != GMS_EMPTY) {
        const int l = (int)(k / GMS_NG), r = (int)(k - (uint32_t)l * GMS_NG);
        atomicMax((unsigned*)&cellPair[l], (hcnt[sidx] << 16) | (uint32_t)(0xFFFF - r));
      }
    }
    __syncthreads();
    // neighbourhood support test (:358-381); decisions are staged so every cell reads the unmodified tables
    int dec[(GMS_NG + GMS_THREADS - 1) / GMS_THREADS];
    {
      int u = 0;
      for (int c = tid; c < GMS_NG; c += GMS_THREADS, u++) {
        const unsigned packed = (unsigned)cellPair[c];
        const int rt = packed ? 0xFFFF - (int)(packed & 0xFFFFu) : -1;  // -1 when the row is empty
        dec[u] = rt;
        if (rt < 0) continue;
        int score = 0, numpair = 0;
        double thresh = 0;
        for (int k = 0; k < 9; k++) {
          const int ll = gms_nb9(c, k), rr = gms_nb9(rt, k);
          if (ll == -1 || rr == -1) continue;
          score += gms_count(hkey, hcnt, tbits, (uint32_t)(ll * GMS_NG + rr));
          thresh += nLeft[ll];
          numpair++;
        }
        thresh = 6.0 * sqrt(thresh / numpair);
        if (score < thresh) dec[u] = -2;
      }
    }
    __syncthreads();
    {
      int u = 0;
      for (int c = tid; c < GMS_NG; c += GMS_THREADS, u++) cellPair[c] = dec[u];
    }
    __syncthreads();
    for (int i = tid; i < nm; i += GMS_THREADS) {
      const int l = ml[i];
      if (l >= 0 && l < GMS_NG && cellPair[l] == (int)mr[i]) mask[i] = 1;
    }
  }
  __syncthreads();
  int c = 0;
  for (int i = tid; i < nm; i += GMS_THREADS) c += mask[i];
  c = __reduce_add_sync(0xffffffffu, c);
  if ((tid & 31) == 0 && c) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (tid == 0) out_count[pair] = s_cnt;
}

static int next_pow2(int n) {
  int p = 32;
  while (p < n) p <<= 1;
  return p;
}

static int launch_gms(cudaStream_t st, const GfsKeyPoint* kp1, const int* n1, const GfsKeyPoint* kp2, const int* n2,
                      const int* mq, const int* mt, const int* nm, int pairs, int stride, int mstride, int w1, int h1,
                      int w2, int h2, uint8_t* inl, int* cnt) {
  // hash table: a power of two >= 1.5 x the matches of a pair (load factor <= 2/3)
  const int tsize = next_pow2(mstride + mstride / 2 + 1);
  int tbits = 0;
  while ((1 << tbits) < tsize) tbits++;
  const size_t smem = (size_t)tsize * 8 + (size_t)mstride * 4 + 2 * GMS_NG * 4;
  GFS_REQUIRE(smem <= 200 * 1024 && mstride < 65536, GFS_ERR_CAPACITY, "too many matches per pair for the GMS kernel (max ~16k)");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    GFS_CUDA(cudaFuncSetAttribute(k_gms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  k_gms<<<pairs, GMS_THREADS, smem, st>>>(kp1, n1, kp2, n2, mq, mt, nm, stride, mstride, tbits, w1, h1, w2, h2, inl, cnt);
  GFS_CUDA(cudaGetLastError());
  return GFS_OK;
}

// match_umma.cu
int launch_bf_hamming_umma(cudaStream_t st, const uint8_t* d_dq, const int* d_nq, const uint8_t* d_dt, const int* d_nt,
                           int pairs, int stride, int* d_out_idx, int* d_out_dist);
}  // namespace gfs

using namespace gfs;

extern "C" {

int gfs_match_bf_hamming_batch_device(void* stream, const uint8_t* d_dq, const int* d_nq, const uint8_t* d_dt,
                                      const int* d_nt, int pairs, int stride, int* d_out_idx, int* d_out_dist) {
  GFS_REQUIRE(d_dq && d_nq && d_dt && d_nt && d_out_idx && d_out_dist, GFS_ERR_INVALID, "null pointer");
  GFS_REQUIRE(pairs > 0 && stride > 0, GFS_ERR_INVALID, "bad pairs/stride");
  // Default: tensor-core kernel (tcgen05.mma.kind::i8 on +-1 expanded descriptors, match_umma.cu).
  // GFS_BF_POPC=1 / GFS_BF_MMA=1 select the two CUDA-core / legacy-MMA kernels below (kept for comparison;
  // all three are bit-identical).
  static int variant = -1;
  if (variant < 0) variant = getenv("GFS_BF_MMA") ? 2 : (getenv("GFS_BF_POPC") ? 1 : 0);
  if (variant == 0 && stride <= 65536)
    return launch_bf_hamming_umma((cudaStream_t)stream, d_dq, d_nq, d_dt, d_nt, pairs, stride, d_out_idx, d_out_dist);
  const size_t smem = (size_t)((stride + 31) & ~31) * 32;
  GFS_REQUIRE(smem <= 200 * 1024, GFS_ERR_CAPACITY, "train set too large for shared memory (max 6400 rows)");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    GFS_CUDA(cudaFuncSetAttribute(k_bf_hamming, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  // GFS_BF_MMA=1 selects the binary-MMA kernel.  Measured on B200 (1023 pairs x 1000 x 1000): POPC kernel
  // 1.90 ms, b1-MMA kernel 2.73 ms -- sm_100a has no native b1 MMA, ptxas emulates it with IMMA.16832.U8
  // sequences, so the POPC-pipe kernel stays the default.
  if (variant == 2) {
    static size_t configured2 = 0;
    if (smem > 48 * 1024 && smem > configured2) {
      GFS_CUDA(cudaFuncSetAttribute(k_bf_hamming_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured2 = smem;
    }
    k_bf_hamming_mma<<<dim3(div_up(stride, BFM_QPB), pairs), BFM_WARPS * 32, smem, (cudaStream_t)stream>>>(
        d_dq, d_nq, d_dt, d_nt, stride, d_out_idx, d_out_dist);
  } else {
    k_bf_hamming<<<dim3(div_up(stride, BF_QPB), pairs), BF_WARPS * 32, smem, (cudaStream_t)stream>>>(
        d_dq, d_nq, d_dt, d_nt, stride, d_out_idx, d_out_dist);
  }
  GFS_CUDA(cudaGetLastError());
  return GFS_OK;
}

int gfs_match_bf_hamming(void* stream, const uint8_t* dq, int nq, const uint8_t* dt, int nt, int* out_idx,
                         int* out_dist) {
  GFS_REQUIRE(nq >= 0 && nt >= 0, GFS_ERR_INVALID, "negative size");
  GFS_REQUIRE((dq || nq == 0) && (dt || nt == 0), GFS_ERR_INVALID, "null descriptors");
  if (nq == 0) return GFS_OK;
  GFS_REQUIRE(out_idx && out_dist, GFS_ERR_INVALID, "null output");
  int rc = gfs_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int stride = std::max(std::max(nq, nt), 1);
  DevBuf b;
  const size_t dsz = (size_t)stride * 32;
  if ((rc = b.reserve(2 * dsz + 2 * sizeof(int) * (size_t)stride + 16))) return rc;
  uint8_t* d_q = (uint8_t*)b.p;
  uint8_t* d_t = d_q + dsz;
  int* d_idx = (int*)(d_t + dsz);
  int* d_dist = d_idx + stride;
  int* d_n = d_dist + stride;
  const int hn[2] = {nq, nt};
  auto fail = [&](int code) { b.release(); return code; };
  if (cudaMemcpyAsync(d_q, dq, (size_t)nq * 32, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      (nt && cudaMemcpyAsync(d_t, dt, (size_t)nt * 32, cudaMemcpyHostToDevice, st) != cudaSuccess) ||
      cudaMemcpyAsync(d_n, hn, sizeof(hn), cudaMemcpyHostToDevice, st) != cudaSuccess) {
    set_error("H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(GFS_ERR_CUDA);
  }
  rc = gfs_match_bf_hamming_batch_device(stream, d_q, d_n, d_t, d_n + 1, 1, stride, d_idx, d_dist);
  if (rc) return fail(rc);
  if (cudaMemcpyAsync(out_idx, d_idx, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaMemcpyAsync(out_dist, d_dist, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) {
    set_error("D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(GFS_ERR_CUDA);
  }
  b.release();
  return GFS_OK;
}

int gfs_gms_filter_batch_device(void* stream, const GfsKeyPoint* d_kp1, const int* d_n1, const GfsKeyPoint* d_kp2,
                                const int* d_n2, const int* d_train_idx, int pairs, int stride, int w1, int h1,
                                int w2, int h2, uint8_t* d_out_inlier, int* d_out_count) {
  GFS_REQUIRE(d_kp1 && d_n1 && d_kp2 && d_n2 && d_train_idx && d_out_inlier && d_out_count, GFS_ERR_INVALID, "null pointer");
  GFS_REQUIRE(pairs > 0 && stride > 0 && w1 > 0 && h1 > 0 && w2 > 0 && h2 > 0, GFS_ERR_INVALID, "bad sizes");
  return launch_gms((cudaStream_t)stream, d_kp1, d_n1, d_kp2, d_n2, nullptr, d_train_idx, nullptr, pairs, stride, stride,
                    w1, h1, w2, h2, d_out_inlier, d_out_count);
}

int gfs_gms_filter(void* stream, const GfsKeyPoint* kp1, int n1, int w1, int h1, const GfsKeyPoint* kp2, int n2,
                   int w2, int h2, const int* matches_qt, int nm, uint8_t* out_inlier, int* out_count) {
  GFS_REQUIRE(n1 >= 0 && n2 >= 0 && nm >= 0 && w1 > 0 && h1 > 0 && w2 > 0 && h2 > 0, GFS_ERR_INVALID, "bad sizes");
  GFS_REQUIRE(out_count, GFS_ERR_INVALID, "null output");
  *out_count = 0;
  if (nm == 0) return GFS_OK;
  GFS_REQUIRE(kp1 && kp2 && matches_qt && out_inlier, GFS_ERR_INVALID, "null pointer");
  int rc = gfs_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int stride = std::max(std::max(n1, n2), 1);
  std::vector<int> mq(nm), mt(nm);
  for (int i = 0; i < nm; i++) { mq[i] = matches_qt[2 * i]; mt[i] = matches_qt[2 * i + 1]; }
  DevBuf b;
  const size_t ksz = (size_t)stride * sizeof(GfsKeyPoint);
  if ((rc = b.reserve(2 * ksz + (size_t)nm * 8 + align_up((size_t)nm, 16) + 64))) return rc;
  GfsKeyPoint* d_k1 = (GfsKeyPoint*)b.p;
  GfsKeyPoint* d_k2 = d_k1 + stride;
  int* d_mq = (int*)(d_k2 + stride);
  int* d_mt = d_mq + nm;
  int* d_n = d_mt + nm;  // n1, n2, nm, count
  uint8_t* d_mask = (uint8_t*)(d_n + 4);
  const int hn[4] = {n1, n2, nm, 0};
  bool ok = true;
  if (n1) ok &= cudaMemcpyAsync(d_k1, kp1, (size_t)n1 * sizeof(GfsKeyPoint), cudaMemcpyHostToDevice, st) == cudaSuccess;
  if (n2) ok &= cudaMemcpyAsync(d_k2, kp2, (size_t)n2 * sizeof(GfsKeyPoint), cudaMemcpyHostToDevice, st) == cudaSuccess;
  ok &= cudaMemcpyAsync(d_mq, mq.data(), (size_t)nm * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
  ok &= cudaMemcpyAsync(d_mt, mt.data(), (size_t)nm * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
  ok &= cudaMemcpyAsync(d_n, hn, sizeof(hn), cudaMemcpyHostToDevice, st) == cudaSuccess;
  if (!ok) {
    set_error("H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    b.release();
    return GFS_ERR_CUDA;
  }
  rc = launch_gms(st, d_k1, d_n, d_k2, d_n + 1, d_mq, d_mt, d_n + 2, 1, stride, nm, w1, h1, w2, h2, d_mask, d_n + 3);
  if (rc) { b.release(); return rc; }
  ok = cudaMemcpyAsync(out_inlier, d_mask, nm, cudaMemcpyDeviceToHost, st) == cudaSuccess;
  ok &= cudaMemcpyAsync(out_count, d_n + 3, sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess;
  ok &= cudaStreamSynchronize(st) == cudaSuccess;
  b.release();
  if (!ok) {
    set_error("D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return GFS_ERR_CUDA;
  }
  return GFS_OK;
}
}
