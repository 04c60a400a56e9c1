// Brute-force Hamming matcher on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces cv::BFMatcher(NORM_HAMMING)::match as called from ORBmatcher::SearchWithGMS
// (reference src/ORBmatcher.cc:752-755; DescriptorDistance :2536 is the per-pair popcount).
//
// A 256-bit descriptor is expanded to 256 int8 values v = 1 - 2*bit.  For two descriptors
//     dot(a, b) = 256 - 2 * hamming(a, b)
// exactly, in integers, so argmin hamming = argmax dot and the distance is recovered without error.
// The 1000 x 1000 x 256 products per frame pair run as tcgen05.mma.kind::i8 (M = 128 queries,
// N = 128 train rows, K = 32 per instruction, s32 accumulators in TMEM).
//
// CTA = one tile of 128 queries against every train tile of the pair:
//   warps 0-3  epilogue: tcgen05.ld the accumulators (one query row per thread), running
//              max of dot * 65536 + (65535 - n): larger dot wins, ties go to the lowest train index
//              (BFMatcher keeps the first minimum)
//   warps 4-7  producers: expand packed descriptors into the canonical K-major, no-swizzle UMMA
//              shared-memory layout (8 x 16 B core matrices)
//   warp  8    allocates TMEM, one lane issues the MMAs
// Two train-tile stages in shared memory and two accumulator stages in TMEM, so expansion, MMA
// and epilogue of consecutive tiles overlap; 96 KB + 256 TMEM columns per CTA -> 2 CTAs per SM.
#include <cstdint>

#include "common.cuh"

namespace gfs {

static const int UM_M = 128, UM_N = 128;       // MMA tile: queries x train rows
static const int UM_KB = 256;                  // bytes of K per row after expansion (256 int8)
static const int UM_TILE_BYTES = UM_N * UM_KB; // 32 KB per operand tile
static const int UM_THREADS = 9 * 32;
static const int UM_SBO = 16 * 128;            // byte stride between 8-row groups (16 core matrices of 128 B)
static const int UM_LBO = 128;                 // byte stride between the two core matrices of one K = 32 step
static const size_t UM_SMEM = 3 * (size_t)UM_TILE_BYTES + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(UM_LBO >> 4) << 16) | ((uint64_t)(UM_SBO >> 4) << 32) |
         (1ull << 46);
}
// instruction descriptor: D = s32, A = B = s8, both K-major, N = 128, M = 128
static const uint32_t UM_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(UM_N >> 3) << 17) | ((uint32_t)(UM_M >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(UM_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// 4 descriptor bits -> 4 int8 (1 - 2*bit), bit k in byte k
__device__ __forceinline__ uint32_t expand_nibble(uint32_t nib) {
  const uint32_t t = (nib * 0x00204081u) & 0x01010101u;
  return t * 0xFEu + 0x01010101u;
}
// Expansion of `rows` packed descriptors (32 B each, row r at g + r*32; rows >= limit re-read row
// limit-1) into one operand tile, in two halves so that the global loads of the next tile can be in
// flight while the current one is written.  A warp step covers 8 rows x 8 K-chunks of 16 B: every
// quarter-warp writes one contiguous 128-B core matrix (conflict-free 128-bit stores).  The bit ->
// int8 expansion is arithmetic (a 256-entry shared-memory table was tried: its bank conflicts made the
// shared-memory pipe the bottleneck).
template <int NW>  // warps sharing the tile; 32 / NW steps per warp
struct TileWords {
  uint32_t w[32 / NW];
  __device__ __forceinline__ void load(const uint8_t* __restrict__ g, int row0, int limit, int worker) {
    const int lane = worker & 31, wp = worker >> 5;
    const int r8 = lane & 7, cp = lane >> 3;
#pragma unroll
    for (int i = 0; i < 32 / NW; i++) {
      const int step = wp + i * NW, grp = step >> 1, half = step & 1;
      const int row = min(row0 + grp * 8 + r8, limit - 1);
      w[i] = __ldg(reinterpret_cast<const uint32_t*>(g + (size_t)row * 32) + half * 4 + cp);
    }
  }
  __device__ __forceinline__ void store(uint8_t* tile, int worker) const {
    const int lane = worker & 31, wp = worker >> 5;
    const int r8 = lane & 7, cp = lane >> 3;
#pragma unroll
    for (int i = 0; i < 32 / NW; i++) {
      const int step = wp + i * NW, grp = step >> 1, half = step & 1;
      const int pair = half * 4 + cp;  // K-chunks 2*pair, 2*pair+1 <-> descriptor bytes 4*pair .. 4*pair+3
      uint8_t* dst = tile + grp * UM_SBO + (2 * pair) * 128 + r8 * 16;
      const uint32_t bits = w[i];
      *reinterpret_cast<uint4*>(dst) = make_uint4(expand_nibble(bits & 15u), expand_nibble((bits >> 4) & 15u),
                                                  expand_nibble((bits >> 8) & 15u), expand_nibble((bits >> 12) & 15u));
      *reinterpret_cast<uint4*>(dst + 128) = make_uint4(expand_nibble((bits >> 16) & 15u), expand_nibble((bits >> 20) & 15u),
                                                        expand_nibble((bits >> 24) & 15u), expand_nibble(bits >> 28));
    }
  }
};

__global__ void __launch_bounds__(UM_THREADS) k_bf_hamming_umma(const uint8_t* __restrict__ dq, const int* __restrict__ nq,
                                                                const uint8_t* __restrict__ dt, const int* __restrict__ nt,
                                                                int stride, int* __restrict__ out_idx,
                                                                int* __restrict__ out_dist) {
  extern __shared__ __align__(1024) uint8_t um_smem[];
  __shared__ __align__(8) unsigned long long s_bar[8];  // full[2], empty[2], accFull[2], accEmpty[2]
  __shared__ uint32_t s_tmem;
  const int pair = blockIdx.y;
  const int nQ = nq[pair], nT = nt[pair];
  const int q0 = blockIdx.x * UM_M;
  if (q0 >= nQ) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint8_t* gq = dq + (size_t)pair * stride * 32;
  const uint8_t* gt = dt + (size_t)pair * stride * 32;
  if (nT <= 0) {
    for (int q = q0 + tid; q < min(q0 + UM_M, nQ); q += UM_THREADS) {
      out_idx[(size_t)pair * stride + q] = -1;
      out_dist[(size_t)pair * stride + q] = -1;
    }
    return;
  }
  uint8_t* base = (uint8_t*)(((uintptr_t)um_smem + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;
  uint8_t* sB = base + UM_TILE_BYTES;
  const uint32_t bar0 = smem_u32(&s_bar[0]);
  auto BAR = [&](int kind, int s) { return bar0 + 8u * (uint32_t)(kind * 2 + s); };  // 0 full, 1 empty, 2 accFull, 3 accEmpty
  const int nTiles = (nT + UM_N - 1) / UM_N;

  if (tid == 0) {
    for (int s = 0; s < 2; s++) {
      mbar_init(BAR(0, s), 128);  // producers
      mbar_init(BAR(1, s), 1);    // tcgen05.commit
      mbar_init(BAR(2, s), 1);    // tcgen05.commit
      mbar_init(BAR(3, s), 128);  // epilogue threads
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // query tile: expanded once by warps 0-7
  TileWords<8> aw;
  if (warp < 8) {
    aw.load(gq, q0, stride, tid);
    aw.store(sA, tid);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  if (warp >= 4 && warp < 8) {
    // ---- producers
    // packed words of the next two tiles stay in flight while the current one is expanded
    TileWords<4> t0, t1, t2;
    t0.load(gt, 0, stride, tid - 128);
    if (nTiles > 1) t1.load(gt, UM_N, stride, tid - 128);
    for (int j = 0; j < nTiles; j++) {
      const int s = j & 1;
      if (j + 2 < nTiles) t2.load(gt, (j + 2) * UM_N, stride, tid - 128);
      mbar_wait(BAR(1, s), ((j >> 1) & 1) ^ 1);
      t0.store(sB + s * UM_TILE_BYTES, tid - 128);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(BAR(0, s));
      t0 = t1;
      t1 = t2;
    }
  } else if (warp == 8) {
    // ---- MMA issuer
    if (lane == 0) {
      const uint64_t descA = umma_smem_desc(smem_u32(sA));
      for (int j = 0; j < nTiles; j++) {
        const int s = j & 1;
        mbar_wait(BAR(0, s), (j >> 1) & 1);
        mbar_wait(BAR(3, s), ((j >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t descB = umma_smem_desc(smem_u32(sB + s * UM_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < UM_KB / 32; k++) {
          // one K = 32 step = two 128-B core matrices further along K: +256 B = +16 in the address field
          umma_i8(tmem + (uint32_t)(s * UM_N), descA + (uint64_t)(k * 16), descB + (uint64_t)(k * 16), k > 0 ? 1u : 0u);
        }
        umma_commit(BAR(1, s));  // train stage may be refilled
        umma_commit(BAR(2, s));  // accumulators are complete
      }
    }
    __syncwarp();
  } else {
    // ---- epilogue: thread = query row (TMEM lane 32*warp + lane)
    int best = -0x7fffffff;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int j = 0; j < nTiles; j++) {
      const int s = j & 1;
      mbar_wait(BAR(2, s), (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int nBase = j * UM_N;
      const bool full = nBase + UM_N <= nT;
#pragma unroll 1
      for (int c = 0; c < UM_N / 32; c++) {
        uint32_t v[32];
        tmem_ld32(lane_addr + (uint32_t)(s * UM_N + c * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int n0 = nBase + c * 32;
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; i++) best = max(best, (int)v[i] * 65536 + (65535 - (n0 + i)));
        } else {
#pragma unroll
          for (int i = 0; i < 32; i++)
            if (n0 + i < nT) best = max(best, (int)v[i] * 65536 + (65535 - (n0 + i)));
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(BAR(3, s));
    }
    const int q = q0 + warp * 32 + lane;
    if (q < nQ) {
      const int dot = best >> 16, n = 65535 - (best & 0xffff);
      out_idx[(size_t)pair * stride + q] = n;
      out_dist[(size_t)pair * stride + q] = (256 - dot) >> 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

int launch_bf_hamming_umma(cudaStream_t st, const uint8_t* d_dq, const int* d_nq, const uint8_t* d_dt, const int* d_nt,
                           int pairs, int stride, int* d_out_idx, int* d_out_dist) {
  // the dynamic shared memory limit is a per-device attribute: one flag per device ordinal, not per process
  static bool configured[64] = {};
  int devId = 0;
  GFS_CUDA(cudaGetDevice(&devId));
  if (devId < 0 || devId >= 64 || !configured[devId]) {
    GFS_CUDA(cudaFuncSetAttribute(k_bf_hamming_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UM_SMEM));
    if (devId >= 0 && devId < 64) configured[devId] = true;
  }
  k_bf_hamming_umma<<<dim3(div_up(stride, UM_M), pairs), UM_THREADS, UM_SMEM, st>>>(d_dq, d_nq, d_dt, d_nt, stride, d_out_idx,
                                                                                   d_out_dist);
  GFS_CUDA(cudaGetLastError());
  return GFS_OK;
}

}  // namespace gfs
