// ORB extractor for batches of frames on sm_100a.
//
// Replaces ORB_SLAM3::ORBextractor::operator() (reference src/ORBextractor.cc:1145-1225) and the
// OpenCV primitives underneath it.  One launch sequence processes a whole batch of frames:
//
//   k_pyr_level   x (nlevels-1)  INTER_AREA pyramid (ComputePyramid :1227-1251, cv::resize)
//   k_fast_cells2                per 35-px cell FAST-9/16 + 3x3 NMS + ini->min threshold retry
//                                (k_fast_cells: first generation, GFS_FAST_V1=1)
//                                (ComputeKeyPointsOctTree cell loop :794-851, cv::FAST)
//   k_octree                     quadtree keypoint distribution (DistributeOctTree :567-768), one
//                                warp per (frame, level), exact std::list / std::sort semantics
//   k_blur7                      7x7 sigma=2 fixed-point Gaussian per level (:1188-1189)
//   k_orient_desc                IC_Angle (:71-95) + rBRIEF (:99-160) + output packing (:1196-1218),
//                                one warp per keypoint
//   k_pack_lapping               only when a lapping area is given (:1208-1217)
//
// All arithmetic follows the oracle (oracle/orb_oracle.cpp): integer stages exactly, float stages
// with explicit _rn intrinsics so nvcc cannot contract them into FMAs.
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>   // CUtensorMap and its enums only: cuTensorMapEncodeTiled is fetched through the runtime (no libcuda link)

#include "common.cuh"
#include "gcc_sort.h"
#include "../../include/gfs_orb_pattern.h"

namespace gfs {

static const int MAX_LEVELS = 12;
static const int EDGE_THRESHOLD = 19;
static const int HALF_PATCH = 15;
static const int PATCH_SIZE = 31;
static const int MIN_BORDER = EDGE_THRESHOLD - 3;  // 16

struct LevelDev {
  int w, h, pitch;
  long long off;  // byte offset of this level inside a frame's pyramid / blur block
  int nCols, nRows, wCell, hCell, cellBase;
  int maxBorderX, maxBorderY;
  int nFeat, nIni;
  float hX;
  int keyOff;          // offset (in keys) inside a frame's key buffers
  int selOff, selCap;  // slot range inside a frame's selected-keypoint array
  int tabX, tabY;      // first entry of this level's area tables
  float scale;
  int patchSize;
};

struct OrbDev {
  int nlevels, iniTh, minTh;
  int cellCap, totalCells, keysPerFrame, selPerFrame, kpStride, nodeCap;
  int roiPitch, roiRows, scPitch, offSc, offCorner, offCand, offKept;  // shared-memory geometry of k_fast_cells
  long long pyrStride;
  float p1, p3, p5, p7, factorPI;  // fastAtan2 polynomial (degrees) and deg->rad
  int umax[16];
  LevelDev lv[MAX_LEVELS];
};

struct AreaEntry {
  int s0, n;
  float a[4];
};


// ------------------------------------------------------------------------------------------------
// k_pyr_level: one thread per destination pixel.  OpenCV's area resize accumulates, per source
// row, buf = S0*a0 (+ S1*a1 (+ S2*a2)) in float and then sum = b0*buf0 (+ b1*buf1 ...) -- the same
// order is kept here, every product and sum rounded separately.
// ------------------------------------------------------------------------------------------------
static const int PYR_TW = 64, PYR_TH = 48;
__global__ void __launch_bounds__(256) k_pyr_level(OrbDev P, int l, const uint8_t* __restrict__ src_base,
                                                   long long src_stride, int src_pitch, uint8_t* __restrict__ pyr,
                                                   const AreaEntry* __restrict__ tabs) {
  // CTA = 64 x 32 destination tile.  Horizontal pass: one float per (source row, destination column)
  // into shared memory (a source row feeds up to two destination rows); vertical pass from there.
  extern __shared__ float s_buf[];  // [source rows of the tile][PYR_TW]
  const LevelDev& L = P.lv[l];
  const int tid = threadIdx.x, frame = blockIdx.z;
  const int dx0 = blockIdx.x * PYR_TW, dy0 = blockIdx.y * PYR_TH;
  const int th = min(PYR_TH, L.h - dy0);
  const int c = tid & (PYR_TW - 1), dx = dx0 + c;
  const uint8_t* src = src_base + (long long)frame * src_stride;
  const AreaEntry eyF = tabs[L.tabY + dy0], eyL = tabs[L.tabY + dy0 + th - 1];
  const int r0 = eyF.s0, nr = eyL.s0 + eyL.n - r0;
  if (dx < L.w) {
    const AreaEntry ex = tabs[L.tabX + dx];
    for (int r = tid >> 6; r < nr; r += 256 / PYR_TW) {
      const uint8_t* S = src + (long long)(r0 + r) * src_pitch + ex.s0;
      float buf = __fmul_rn((float)__ldg(S), ex.a[0]);
      for (int k = 1; k < ex.n; k++) buf = __fadd_rn(buf, __fmul_rn((float)__ldg(S + k), ex.a[k]));
      s_buf[r * PYR_TW + c] = buf;
    }
  }
  __syncthreads();
  if (dx >= L.w) return;
  for (int ry = tid >> 6; ry < th; ry += 256 / PYR_TW) {
    const AreaEntry ey = tabs[L.tabY + dy0 + ry];
    const float* bcol = s_buf + (ey.s0 - r0) * PYR_TW + c;
    float sum = __fmul_rn(ey.a[0], bcol[0]);
    for (int r = 1; r < ey.n; r++) sum = __fadd_rn(sum, __fmul_rn(ey.a[r], bcol[r * PYR_TW]));
    int v = __float2int_rn(sum);
    v = min(255, max(0, v));
    pyr[(long long)frame * P.pyrStride + L.off + (long long)(dy0 + ry) * L.pitch + dx] = (uint8_t)v;
  }
}

// ------------------------------------------------------------------------------------------------
// k_pyr_level3: the same arithmetic for scale factors below 2 (every destination pixel covers 2 or 3
// source pixels; the tables are padded to exactly 3 taps with zero weights, which adds an exact 0).
// CTA = 64 x 32 destination tile.  Source window -> shared memory as floats (word loads, each source
// pixel converted once) -> horizontal pass into shared memory -> vertical pass, 4 pixels per thread,
// packed 32-bit stores.
// ------------------------------------------------------------------------------------------------
struct Area3 {
  int s0;
  float a0, a1, a2;
};
__device__ __forceinline__ unsigned sat_u8_rn(float v) {  // saturate_cast<uchar>(cvRound(v))
  unsigned r;
  asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__global__ void __launch_bounds__(256) k_pyr_level3(OrbDev P, int l, const uint8_t* __restrict__ src_base,
                                                    long long src_stride, int src_pitch, int word_ok,
                                                    uint8_t* __restrict__ pyr, const Area3* __restrict__ tabs, int srcRows,
                                                    int srcPitchF) {
  extern __shared__ float s_buf[];
  float* s_src = s_buf;                        // [srcRows][srcPitchF]
  float* s_h = s_buf + srcRows * srcPitchF;    // [srcRows][PYR_TW]
  const LevelDev& L = P.lv[l];
  const int tid = threadIdx.x, frame = blockIdx.z;
  const int dx0 = blockIdx.x * PYR_TW, dy0 = blockIdx.y * PYR_TH;
  const int tw = min(PYR_TW, L.w - dx0), th = min(PYR_TH, L.h - dy0);
  const Area3* tx = tabs + L.tabX + dx0;
  const Area3* ty = tabs + L.tabY + dy0;
  const int xs = tx[0].s0 & ~3;
  const int words = ((tx[tw - 1].s0 + 2 - xs) >> 2) + 1;
  const int r0 = ty[0].s0, nr = ty[th - 1].s0 + 3 - r0;
  const uint8_t* src = src_base + (long long)frame * src_stride + (long long)r0 * src_pitch + xs;
  {
    const int wd = tid & 31;
    if (wd < words) {
      const bool whole = word_ok && (xs + 4 * wd + 4 <= src_pitch);
      const uint8_t* S0 = src + 4 * wd;
      float* D0 = s_src + 4 * wd;
      if (whole) {
        // four rows per trip: all four loads are issued before the first conversion waits on one
        for (int r = tid >> 5; r < nr; r += 32) {
          unsigned v[4];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int rr = min(r + 8 * k, nr - 1);  // clamped rows reload the last row (stored again below, same value)
            v[k] = __ldg(reinterpret_cast<const unsigned*>(S0 + (long long)rr * src_pitch));
          }
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int rr = min(r + 8 * k, nr - 1);
            *reinterpret_cast<float4*>(D0 + rr * srcPitchF) =
                make_float4((float)(v[k] & 0xffu), (float)((v[k] >> 8) & 0xffu), (float)((v[k] >> 16) & 0xffu), (float)(v[k] >> 24));
          }
        }
      } else {
        const int lim = src_pitch - 1 - (xs + 4 * wd);  // last readable byte of the row
        for (int r = tid >> 5; r < nr; r += 8) {
          const uint8_t* S = S0 + (long long)r * src_pitch;
          *reinterpret_cast<float4*>(D0 + r * srcPitchF) =
              make_float4((float)__ldg(S + min(0, lim)), (float)__ldg(S + min(1, lim)), (float)__ldg(S + min(2, lim)),
                          (float)__ldg(S + min(3, lim)));
        }
      }
    }
  }
  __syncthreads();
  {
    const int c = tid & (PYR_TW - 1);
    if (c < tw) {
      const Area3 e = tx[c];
      const float* p = s_src + (e.s0 - xs);
#pragma unroll 2
      for (int r = tid >> 6; r < nr; r += 256 / PYR_TW) {
        const float* q = p + r * srcPitchF;
        float buf = __fmul_rn(q[0], e.a0);
        buf = __fadd_rn(buf, __fmul_rn(q[1], e.a1));
        buf = __fadd_rn(buf, __fmul_rn(q[2], e.a2));
        s_h[r * PYR_TW + c] = buf;
      }
    }
  }
  __syncthreads();
  {
    const int cg = (tid & 15) * 4;
    if (cg < tw) {
      uint8_t* out = pyr + (long long)frame * P.pyrStride + L.off + dx0 + cg;
      for (int ry = tid >> 4; ry < th; ry += 16) {
        const Area3 e = ty[ry];
        const float* q = s_h + (e.s0 - r0) * PYR_TW + cg;
        const float4 h0 = *reinterpret_cast<const float4*>(q), h1 = *reinterpret_cast<const float4*>(q + PYR_TW),
                     h2 = *reinterpret_cast<const float4*>(q + 2 * PYR_TW);
        float sx = __fmul_rn(e.a0, h0.x), sy = __fmul_rn(e.a0, h0.y), sz = __fmul_rn(e.a0, h0.z), sw = __fmul_rn(e.a0, h0.w);
        sx = __fadd_rn(sx, __fmul_rn(e.a1, h1.x)); sy = __fadd_rn(sy, __fmul_rn(e.a1, h1.y));
        sz = __fadd_rn(sz, __fmul_rn(e.a1, h1.z)); sw = __fadd_rn(sw, __fmul_rn(e.a1, h1.w));
        sx = __fadd_rn(sx, __fmul_rn(e.a2, h2.x)); sy = __fadd_rn(sy, __fmul_rn(e.a2, h2.y));
        sz = __fadd_rn(sz, __fmul_rn(e.a2, h2.z)); sw = __fadd_rn(sw, __fmul_rn(e.a2, h2.w));
        const unsigned v = sat_u8_rn(sx) | (sat_u8_rn(sy) << 8) | (sat_u8_rn(sz) << 16) | (sat_u8_rn(sw) << 24);
        *reinterpret_cast<unsigned*>(out + (long long)(dy0 + ry) * L.pitch) = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_fast_cells: one CTA per (cell, frame).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool has_arc9(unsigned m16) {
  unsigned m = m16 | (m16 << 16);
  unsigned r = m & (m >> 1);
  r &= r >> 2;
  r &= r >> 4;
  r &= m >> 8;
  return (r & 0xFFFFu) != 0;
}

// best = max over the 16 arcs of 9 contiguous ring pixels of max(min d, min -d).
// NOTE (toolchain): written with min-only networks over d and -d.  The equivalent
// max(best, max(min d, -(max d))) form is MISCOMPILED by ptxas 12.9 for sm_100a at -O1 and above
// (it returned max(d) for rings such as {33,32,30,29,31,27,30,20,-1,-1,0,-2,-2,7,36,37}; correct
// at -Xptxas -O0) -- found by the oracle parity test, see DESIGN.md "ptxas min/max miscompile".
__device__ __forceinline__ int fast_best16(const int (&d)[16]) {
  int nd[16];
#pragma unroll
  for (int k = 0; k < 16; k++) nd[k] = -d[k];
  int a2[16], b2[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    a2[k] = min(d[k], d[(k + 1) & 15]);
    b2[k] = min(nd[k], nd[(k + 1) & 15]);
  }
  int a4[16], b4[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    a4[k] = min(a2[k], a2[(k + 2) & 15]);
    b4[k] = min(b2[k], b2[(k + 2) & 15]);
  }
  int bestD = -1000, bestB = -1000;
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const int a = min(min(a4[k], a4[(k + 4) & 15]), d[(k + 8) & 15]);
    const int b = min(min(b4[k], b4[(k + 4) & 15]), nd[(k + 8) & 15]);
    bestD = max(bestD, a);
    bestB = max(bestB, b);
  }
  return max(bestD, bestB);
}

static const int FAST_THREADS = 128;

// Shared memory: ROI pixels (word-aligned rows), the score map with a zero border, the list of
// detected corners and the list of NMS survivors.  Per-pixel work is the antipodal-pair reject
// (4 loads); only the survivors pay for the 16-pixel ring test, only corners for score + NMS.
__global__ void __launch_bounds__(FAST_THREADS) k_fast_cells(OrbDev P, const uint8_t* __restrict__ img0,
                                                              long long img_stride, int pitch0,
                                                              const uint8_t* __restrict__ pyr,
                                                              uint32_t* __restrict__ cellKeys,
                                                              int* __restrict__ cellCount) {
  extern __shared__ __align__(16) uint8_t sm[];
  __shared__ int s_nCand, s_nCorner, s_nKept;
  const int tid = threadIdx.x;
  const int cell = blockIdx.x, frame = blockIdx.y;
  int l = 0;
  while (l + 1 < P.nlevels && cell >= P.lv[l + 1].cellBase) l++;
  const LevelDev& L = P.lv[l];
  const int c = cell - L.cellBase;
  const int ci = c / L.nCols, cj = c - ci * L.nCols;
  const int iniY = MIN_BORDER + ci * L.hCell, iniX = MIN_BORDER + cj * L.wCell;
  int maxY = iniY + L.hCell + 6, maxX = iniX + L.wCell + 6;
  const bool skip = (iniY >= L.maxBorderY - 3) || (iniX >= L.maxBorderX - 6);
  maxY = min(maxY, L.maxBorderY);
  maxX = min(maxX, L.maxBorderX);
  const int rw = maxX - iniX, rh = maxY - iniY;
  int* outCount = cellCount + (long long)frame * P.totalCells + cell;
  if (skip || rw < 7 || rh < 7) {
    if (tid == 0) *outCount = 0;
    return;
  }
  const uint8_t* src;
  int pitch;
  if (l == 0) {
    src = img0 + (long long)frame * img_stride;
    pitch = pitch0;
  } else {
    src = pyr + (long long)frame * P.pyrStride + L.off;
    pitch = L.pitch;
  }
  const int RP = P.roiPitch, SP = P.scPitch;
  uint8_t* sc = sm + P.offSc;                       // [(dh+2)][SP], zero border
  uint16_t* corners = (uint16_t*)(sm + P.offCorner);
  uint16_t* cand = (uint16_t*)(sm + P.offCand);
  uint32_t* keptList = (uint32_t*)(sm + P.offKept);
  const int dw = rw - 6, dh = rh - 6;
  const int shift = iniX & 3;
  const uint8_t* roi = sm + shift;                  // pixel (y, x) of the ROI at roi[y * RP + x]
  if (((pitch & 3) == 0) && ((((size_t)src) & 3) == 0)) {
    const int nW = (shift + rw + 3) >> 2;
    const uint32_t* s32 = (const uint32_t*)(src + (long long)iniY * pitch + (iniX - shift));
    const int p32 = pitch >> 2, rp32 = RP >> 2;
    uint32_t* d32 = (uint32_t*)sm;
    if (nW <= 16) {
      // 16 lanes per row, 8 rows per step: no index division
      const int x = tid & 15;
      if (x < nW)
        for (int y = tid >> 4; y < rh; y += FAST_THREADS / 16) d32[y * rp32 + x] = __ldg(s32 + (long long)y * p32 + x);
    } else {
      for (int idx = tid; idx < nW * rh; idx += FAST_THREADS) {
        const int y = idx / nW, x = idx - y * nW;
        d32[y * rp32 + x] = __ldg(s32 + (long long)y * p32 + x);
      }
    }
  } else {
    for (int idx = tid; idx < rw * rh; idx += FAST_THREADS) {
      const int y = idx / rw, x = idx - y * rw;
      sm[shift + y * RP + x] = __ldg(src + (long long)(iniY + y) * pitch + iniX + x);
    }
  }
  {
    uint32_t* z = (uint32_t*)sc;
    for (int idx = tid; idx < ((dh + 2) * SP) >> 2; idx += FAST_THREADS) z[idx] = 0u;
  }
  if (tid == 0) { s_nCorner = 0; s_nCand = 0; s_nKept = 0; }
  __syncthreads();

  const int npx = dw * dh;
  const int thMin = P.minTh, thIni = P.iniTh;
  const int lane = tid & 31;
  int nKept = 0;
  // cv::FAST(cell, iniThFAST) first; only a cell without any surviving corner is redone at minThFAST
  // (ORBextractor.cc:809-826).  At threshold th the score map only holds corners at th, which is all
  // the 3x3 NMS of that FAST call ever sees.
  for (int pass = 0; pass < 2; pass++) {
    const int th = pass == 0 ? thIni : thMin;
    // phase 1 (every pixel, all lanes busy): antipodal-pair reject.  Every 9-arc of the 16-ring holds
    // one pixel of each antipodal pair, so both (0,8) and (4,12) must have a bright (dark) member.
    {
      int x = tid % dw, y = tid / dw;
      const int sx = FAST_THREADS % dw, sy = FAST_THREADS / dw;
      const int rounds = (npx + FAST_THREADS - 1) / FAST_THREADS;
      for (int r = 0; r < rounds; r++) {
        const int idx = r * FAST_THREADS + tid;
        bool ok = false;
        if (idx < npx) {
          const uint8_t* p = roi + (y + 3) * RP + (x + 3);
          const int cval = p[0];
          const int hi = cval + th, lo = cval - th;
          const int v0 = p[3 * RP], v8 = p[-3 * RP], v4 = p[3], v12 = p[-3];
          const bool br = ((v0 > hi) | (v8 > hi)) & ((v4 > hi) | (v12 > hi));
          const bool dk = ((v0 < lo) | (v8 < lo)) & ((v4 < lo) | (v12 < lo));
          ok = br | dk;
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m) {
          int basePos = 0;
          if (lane == 0) basePos = atomicAdd(&s_nCand, __popc(m));
          basePos = __shfl_sync(0xffffffffu, basePos, 0);
          if (ok) cand[basePos + __popc(m & ((1u << lane) - 1u))] = (uint16_t)idx;
        }
        x += sx; y += sy;
        if (x >= dw) { x -= dw; y++; }
      }
    }
    __syncthreads();
    // phase 2 (survivors only, densely packed): full ring test, score, corner list
    {
      const int nCand = s_nCand;
      for (int i = tid; i < nCand; i += FAST_THREADS) {
        const int idx = cand[i];
        const int y = idx / dw, x = idx - y * dw;
        const uint8_t* p = roi + (y + 3) * RP + (x + 3);
        const int cval = p[0];
        const int hi = cval + th, lo = cval - th;
        int v[16];
        v[0] = p[3 * RP];       v[1] = p[3 * RP + 1];   v[2] = p[2 * RP + 2];   v[3] = p[RP + 3];
        v[4] = p[3];            v[5] = p[-RP + 3];      v[6] = p[-2 * RP + 2];  v[7] = p[-3 * RP + 1];
        v[8] = p[-3 * RP];      v[9] = p[-3 * RP - 1];  v[10] = p[-2 * RP - 2]; v[11] = p[-RP - 3];
        v[12] = p[-3];          v[13] = p[RP - 3];      v[14] = p[2 * RP - 2];  v[15] = p[3 * RP - 1];
        unsigned bright = 0, dark = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
          bright |= (unsigned)(v[k] > hi) << k;
          dark |= (unsigned)(v[k] < lo) << k;
        }
        if (has_arc9(bright) || has_arc9(dark)) {
          int d[16];
#pragma unroll
          for (int k = 0; k < 16; k++) d[k] = cval - v[k];
          const int best = fast_best16(d);
          sc[(y + 1) * SP + (x + 1)] = (uint8_t)(best - 1);  // best > th >= 0, best <= 255
          corners[atomicAdd(&s_nCorner, 1)] = (uint16_t)idx;
        }
      }
    }
    __syncthreads();
    // 3x3 NMS with strict '>' over the corners of this pass; pixels outside the cell's tested region
    // (and non-corners) score 0
    {
      const int nCorner = s_nCorner;
      for (int i = tid; i < nCorner; i += FAST_THREADS) {
        const int idx = corners[i];
        const int y = idx / dw, x = idx - y * dw;
        const uint8_t* q = sc + (y + 1) * SP + (x + 1);
        const int sv = q[0];
        if (sv == 0) continue;
        int m = max(max(q[-SP - 1], q[-SP]), max(q[-SP + 1], q[-1]));
        m = max(m, max(max(q[1], q[SP - 1]), max(q[SP], q[SP + 1])));
        if (sv > m) keptList[atomicAdd(&s_nKept, 1)] = (uint32_t)idx | ((uint32_t)sv << 16);
      }
    }
    __syncthreads();
    nKept = s_nKept;
    if (nKept > 0 || thIni == thMin) break;
    // nothing at iniTh: clear and redo the cell at minTh
    __syncthreads();
    {
      uint32_t* z = (uint32_t*)sc;
      for (int idx = tid; idx < ((dh + 2) * SP) >> 2; idx += FAST_THREADS) z[idx] = 0u;
    }
    if (tid == 0) { s_nCorner = 0; s_nCand = 0; s_nKept = 0; }
    __syncthreads();
  }
  // output in cv::FAST order (row-major): rank = number of survivors with a smaller pixel index
  uint32_t* out = cellKeys + ((long long)frame * P.totalCells + cell) * P.cellCap;
  for (int i = tid; i < nKept; i += FAST_THREADS) {
    const uint32_t e = keptList[i];
    const int idx = (int)(e & 0xffffu);
    int rank = 0;
    for (int j = 0; j < nKept; j++) rank += (int)(keptList[j] & 0xffffu) < idx;
    const int y = idx / dw, x = idx - y * dw;
    const unsigned kx = (unsigned)(x + 3 + cj * L.wCell), ky = (unsigned)(y + 3 + ci * L.hCell);
    out[rank] = kx | (ky << 12) | ((e >> 16) << 24);
  }
  if (tid == 0) *outCount = nKept;
}

// ------------------------------------------------------------------------------------------------
// k_fast_cells2: same contract as k_fast_cells (one CTA per (cell, frame), same outputs), fewer
// instructions per pixel.
//   phase 1: one thread per two aligned 4-pixel words of the tested region.  The words and their ring
//            neighbours (3 rows up / down, 3 columns left / right) are split into even / odd bytes
//            as 16-bit lanes, so `v > c + th` and `v < c - th` are one 32-bit add each for two
//            pixels (bit 15 of `c + th + 0x8000 - v` / `v + 0x8000 - c + th`; no lane ever
//            borrows).  Survivors of the antipodal-pair reject are appended as (y << 8 | x) keys.
//   phase 2: per survivor the 16 ring differences are packed as (d + 256, -d + 256) in the two
//            halves of a register (one IMAD: v * 0xFFFF + const) and the nine-arc min / max network
//            runs on both polarities at once with VIMNMX(3).S16x2.  corner <=> best > th, which is
//            the arc test itself (min over an arc of +-d > th).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned fast_best16_packed(const unsigned (&q)[16]) {
  unsigned a2[16], a4[16];
#pragma unroll
  for (int k = 0; k < 16; k++) a2[k] = __vmins2(q[k], q[(k + 1) & 15]);
#pragma unroll
  for (int k = 0; k < 16; k++) a4[k] = __vmins2(a2[k], a2[(k + 2) & 15]);
  unsigned best = 0u;
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    const unsigned e0 = __vimin3_s16x2(a4[k], a4[(k + 4) & 15], q[(k + 8) & 15]);
    const unsigned e1 = __vimin3_s16x2(a4[k + 1], a4[(k + 5) & 15], q[(k + 9) & 15]);
    best = __vimax3_s16x2(best, e0, e1);
  }
  return best;
}

// Tensor maps of the pyramid levels for k_fast_cells2: level l of ALL frames of the batch as one rank-3 tensor (x, y, frame), box
// = one cell's ROI (roiPitch x roiRows x 1).  A CTA fetches its ROI with ONE cp.async.bulk.tensor.3d (UTMALDG) at the ROI's own
// coordinates rounded down to a 16-byte column (the tensor path traps on a box whose innermost start is not a multiple of 16
// bytes -- scripts/probe/tmap_probe3.cu; the box is wide enough for the shift, as with the per-row copies).  mask bit l is
// clear when level l cannot have a map (base / pitch / frame stride not multiples of 16): the CTA then stages as before.
struct alignas(64) FastTmaps {
  CUtensorMap m[MAX_LEVELS];
  unsigned mask;
};

// One precomputed record per cell (all levels of a frame): level and ROI of ComputeKeyPointsOctTree's
// cell loop (:794-808), so the CTA does not search the level table or divide.
struct FastCell {
  int l, iniX, iniY, rwrh;  // rw | rh << 16; 0: nothing to test in this cell
};

__device__ __forceinline__ void fast_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(FAST_THREADS) k_fast_cells2(OrbDev P, const __grid_constant__ FastTmaps tm,
                                                               const FastCell* __restrict__ cellTab,
                                                               const uint8_t* __restrict__ img0,
                                                               long long img_stride, int pitch0,
                                                               const uint8_t* __restrict__ pyr,
                                                               uint32_t* __restrict__ cellKeys,
                                                               int* __restrict__ cellCount) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  // the tensor-map load wants a 128-byte aligned destination: the dynamic window starts wherever the static variables end
  uint8_t* sm = sm_raw + ((128u - ((uint32_t)__cvta_generic_to_shared(sm_raw) & 127u)) & 127u);
  __shared__ int s_nCand, s_nCorner, s_nKept;
  __shared__ unsigned s_vm[FAST_THREADS];
  __shared__ __align__(8) unsigned long long s_bar;
  const int tid = threadIdx.x;
  const int cell = blockIdx.x, frame = blockIdx.y;
  const int4 ct = __ldg(reinterpret_cast<const int4*>(cellTab) + cell);
  const int l = ct.x, iniX = ct.y, iniY = ct.z;
  const int rw = ct.w & 0xffff, rh = ct.w >> 16;
  int* outCount = cellCount + (long long)frame * P.totalCells + cell;
  if (ct.w == 0) {
    if (tid == 0) *outCount = 0;
    return;
  }
  const uint8_t* src;
  int pitch;
  if (l == 0) {
    src = img0 + (long long)frame * img_stride;
    pitch = pitch0;
  } else {
    src = pyr + (long long)frame * P.pyrStride + P.lv[l].off;
    pitch = P.lv[l].pitch;
  }
  const int RP = P.roiPitch, SP = P.scPitch;
  uint8_t* sc = sm + P.offSc;                       // [(dh+2)][SP], zero border
  uint16_t* corners = (uint16_t*)(sm + P.offCorner);
  uint16_t* cand = (uint16_t*)(sm + P.offCand);
  uint32_t* keptList = (uint32_t*)(sm + P.offKept);
  const int dw = rw - 6, dh = rh - 6;
  const int rp32 = RP >> 2;
  // The ROI is staged so that its 16-byte (4-byte) groups are the source's: pixel (y, x) at
  // sm[y * RP + shift + x].  Rows whose 16-byte groups are addressable go through the bulk-copy
  // engine (one cp.async.bulk per row, completion on an mbarrier) while the CTA clears the score
  // map; otherwise word / byte loads.
  const bool tmap = (tm.mask >> l) & 1u;
  const bool bulk = !tmap && ((pitch & 15) == 0) && ((((size_t)src) & 15) == 0);
  int shift;
  if (tmap) {
    shift = iniX & 15;   // the box must start on a 16-byte boundary of the innermost dimension (measured: anything else traps)
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(RP * P.roiRows)) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
              (uint32_t)__cvta_generic_to_shared(sm)),
          "l"(&tm.m[l]), "r"(iniX - shift), "r"(iniY), "r"(frame), "r"(bar)
          : "memory");
    }
    __syncthreads();   // the barrier's initialisation is visible to the threads that will wait on it
  } else if (bulk) {
    shift = iniX & 15;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    const uint32_t rowBytes = (uint32_t)((shift + rw + 15) & ~15);  // <= RP, stays inside the source row (pitch % 16 == 0)
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(rowBytes * (uint32_t)rh) : "memory");
    }
    __syncthreads();
    const uint8_t* g0 = src + (long long)iniY * pitch + (iniX - shift);
    const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(sm);
    for (int y = tid; y < rh; y += FAST_THREADS) {
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d0 + (uint32_t)(y * RP)),
          "l"(g0 + (long long)y * pitch), "r"(rowBytes), "r"(bar)
          : "memory");
    }
  } else if ((pitch & 3) == 0) {
    shift = (int)(((size_t)src + (size_t)iniX) & 3);
    const int nW = (shift + rw + 3) >> 2;
    const uint32_t* s32 = (const uint32_t*)(src + (long long)iniY * pitch + (iniX - shift));
    const int p32 = pitch >> 2;
    uint32_t* d32 = (uint32_t*)sm;
    for (int idx = tid; idx < nW * rh; idx += FAST_THREADS) {
      const int y = idx / nW, x = idx - y * nW;
      d32[y * rp32 + x] = __ldg(s32 + (long long)y * p32 + x);
    }
  } else {
    shift = (int)(((size_t)src + (size_t)iniX) & 3);
    for (int idx = tid; idx < rw * rh; idx += FAST_THREADS) {
      const int y = idx / rw, x = idx - y * rw;
      sm[shift + y * RP + x] = __ldg(src + (long long)(iniY + y) * pitch + iniX + x);
    }
  }
  const uint8_t* roi = sm + shift;
  {
    uint4* z = (uint4*)sc;  // offSc and the region's size are multiples of 16
    for (int idx = tid; idx < ((dh + 2) * SP + 15) >> 4; idx += FAST_THREADS) z[idx] = make_uint4(0u, 0u, 0u, 0u);
  }
  const int wLo = (shift + 3) >> 2, wHi = (shift + rw - 4) >> 2, nWt = wHi - wLo + 1;
  // i / nPair == (i * inv) >> 16 for i < 4096 (nPair <= 32: the float quotient cannot round across an integer)
  const int nPair = (nWt + 1) >> 1;  // a thread of phase 1 takes two adjacent words
  const unsigned inv = (unsigned)(65536.0f / (float)nPair) + 1u;
  if (tid <= nWt) {  // entry nWt (all zero) pads an odd row
    // which bytes of word wLo + tid are tested pixels (column in [0, dw)), as bits 14 / 15 / 30 / 31
    const int xb = 4 * (wLo + tid) - shift - 3;
    unsigned vm = 0u;
    if (xb >= 0 && xb < dw) vm |= 0x00004000u;
    if (xb + 1 >= 0 && xb + 1 < dw) vm |= 0x00008000u;
    if (xb + 2 >= 0 && xb + 2 < dw) vm |= 0x40000000u;
    if (xb + 3 >= 0 && xb + 3 < dw) vm |= 0x80000000u;
    s_vm[tid] = vm;
  }
  if (tid == 0) { s_nCorner = 0; s_nCand = 0; s_nKept = 0; }
  if (bulk || tmap) fast_mbar_wait((uint32_t)__cvta_generic_to_shared(&s_bar), 0u);
  __syncthreads();

  const int thMin = P.minTh, thIni = P.iniTh;
  const int items = nPair * dh;  // (row, word pair) grid of the words that hold tested pixels (x in [3, rw - 3))
  const uint32_t* sm32 = (const uint32_t*)sm;
  int nKept = 0;
  for (int pass = 0; pass < 2; pass++) {
    const int th = pass == 0 ? thIni : thMin;
    // phase 1: antipodal-pair reject, eight pixels (two adjacent words) per thread.  Every 9-arc of the
    // 16-ring holds one pixel of each antipodal pair, so both (0,8) and (4,12) must have a bright (dark)
    // member.  E(w) / O(w) = even / odd bytes of a word as 16-bit lanes; the neighbours 3 columns to the
    // left / right of the even pixels are odd bytes and vice versa, half of them straddle two words.
    {
      const unsigned K = (unsigned)(th + 0x8000) * 0x00010001u;
      const uint32_t* r0 = sm32 + 3 * rp32 + wLo;
      const int key0 = 4 * wLo - shift - 3;
      auto reject = [K](unsigned c, unsigned u, unsigned d, unsigned lf, unsigned rt) -> unsigned {
        const unsigned HB = c + K, LB = K - c;
        const unsigned notBright = ((HB - u) & (HB - d)) | ((HB - lf) & (HB - rt));
        const unsigned notDark = ((LB + u) & (LB + d)) | ((LB + lf) & (LB + rt));
        return ~(notBright & notDark) & 0x80008000u;
      };
      for (int it = tid; it < items; it += FAST_THREADS) {
        const int y = (int)(((unsigned)it * inv) >> 16), w = 2 * (it - y * nPair);
        const uint32_t* r = r0 + y * rp32 + w;
        const unsigned key = (unsigned)((y << 8) + 4 * w + key0);  // (row << 8 | column) of byte 0
        const unsigned Wm = r[-1], C0 = r[0], C1 = r[1], Wp = r[2];
        const unsigned U0 = r[-3 * rp32], U1 = r[-3 * rp32 + 1], D0 = r[3 * rp32], D1 = r[3 * rp32 + 1];
        const unsigned eM = __byte_perm(Wm, 0u, 0x4240), oM = __byte_perm(Wm, 0u, 0x4341);
        const unsigned e0 = __byte_perm(C0, 0u, 0x4240), o0 = __byte_perm(C0, 0u, 0x4341);
        const unsigned e1 = __byte_perm(C1, 0u, 0x4240), o1 = __byte_perm(C1, 0u, 0x4341);
        const unsigned eP = __byte_perm(Wp, 0u, 0x4240), oP = __byte_perm(Wp, 0u, 0x4341);
        const unsigned t0e = reject(e0, __byte_perm(U0, 0u, 0x4240), __byte_perm(D0, 0u, 0x4240), oM, __byte_perm(o0, o1, 0x5432));
        const unsigned t0o = reject(o0, __byte_perm(U0, 0u, 0x4341), __byte_perm(D0, 0u, 0x4341), __byte_perm(eM, e0, 0x5432), e1);
        const unsigned t1e = reject(e1, __byte_perm(U1, 0u, 0x4240), __byte_perm(D1, 0u, 0x4240), o0, __byte_perm(o1, oP, 0x5432));
        const unsigned t1o = reject(o1, __byte_perm(U1, 0u, 0x4341), __byte_perm(D1, 0u, 0x4341), __byte_perm(e0, e1, 0x5432), eP);
        // bits 14 / 15 / 30 / 31 = pixels 0 / 1 / 2 / 3 of a word
        const unsigned ta = ((t0e >> 1) | t0o) & s_vm[w], tb = ((t1e >> 1) | t1o) & s_vm[w + 1];
        if (ta | tb) {
          int pos = atomicAdd(&s_nCand, __popc(ta) + __popc(tb));
          if (ta & 0x00004000u) cand[pos++] = (uint16_t)key;
          if (ta & 0x00008000u) cand[pos++] = (uint16_t)(key + 1u);
          if (ta & 0x40000000u) cand[pos++] = (uint16_t)(key + 2u);
          if (ta & 0x80000000u) cand[pos++] = (uint16_t)(key + 3u);
          if (tb & 0x00004000u) cand[pos++] = (uint16_t)(key + 4u);
          if (tb & 0x00008000u) cand[pos++] = (uint16_t)(key + 5u);
          if (tb & 0x40000000u) cand[pos++] = (uint16_t)(key + 6u);
          if (tb & 0x80000000u) cand[pos] = (uint16_t)(key + 7u);
        }
      }
    }
    __syncthreads();
    // phase 2 (survivors only, densely packed): score over the nine-arcs, corner list
    {
      const int nCand = s_nCand;
      for (int i = tid; i < nCand; i += FAST_THREADS) {
        const unsigned key = cand[i];
        const int y = (int)(key >> 8), x = (int)(key & 255u);
        const uint8_t* p = roi + (y + 3) * RP + (x + 3);
        const unsigned cval = p[0];
        const unsigned CC = (cval + 256u) + ((256u - cval) << 16);
        unsigned q[16];
        q[0] = p[3 * RP];       q[1] = p[3 * RP + 1];   q[2] = p[2 * RP + 2];   q[3] = p[RP + 3];
        q[4] = p[3];            q[5] = p[-RP + 3];      q[6] = p[-2 * RP + 2];  q[7] = p[-3 * RP + 1];
        q[8] = p[-3 * RP];      q[9] = p[-3 * RP - 1];  q[10] = p[-2 * RP - 2]; q[11] = p[-RP - 3];
        q[12] = p[-3];          q[13] = p[RP - 3];      q[14] = p[2 * RP - 2];  q[15] = p[3 * RP - 1];
#pragma unroll
        for (int k = 0; k < 16; k++) q[k] = q[k] * 0xFFFFu + CC;  // (c - v + 256) | (v - c + 256) << 16
        const unsigned b2 = fast_best16_packed(q);
        const int best = (int)max(b2 & 0xFFFFu, b2 >> 16) - 256;
        if (best > th) {
          sc[(y + 1) * SP + (x + 1)] = (uint8_t)(best - 1);  // th >= 0, best <= 255
          corners[atomicAdd(&s_nCorner, 1)] = (uint16_t)key;
        }
      }
    }
    __syncthreads();
    // 3x3 NMS with strict '>' over the corners of this pass; pixels outside the cell's tested region
    // (and non-corners) score 0
    {
      const int nCorner = s_nCorner;
      for (int i = tid; i < nCorner; i += FAST_THREADS) {
        const unsigned key = corners[i];
        const int y = (int)(key >> 8), x = (int)(key & 255u);
        const uint8_t* q = sc + (y + 1) * SP + (x + 1);
        const int sv = q[0];
        if (sv == 0) continue;
        int m = max(max(q[-SP - 1], q[-SP]), max(q[-SP + 1], q[-1]));
        m = max(m, max(max(q[1], q[SP - 1]), max(q[SP], q[SP + 1])));
        if (sv > m) keptList[atomicAdd(&s_nKept, 1)] = key | ((uint32_t)sv << 16);
      }
    }
    __syncthreads();
    nKept = s_nKept;
    if (nKept > 0 || thIni == thMin) break;
    // nothing at iniTh: clear and redo the cell at minTh
    __syncthreads();
    {
      uint32_t* z = (uint32_t*)sc;
      for (int idx = tid; idx < ((dh + 2) * SP) >> 2; idx += FAST_THREADS) z[idx] = 0u;
    }
    if (tid == 0) { s_nCorner = 0; s_nCand = 0; s_nKept = 0; }
    __syncthreads();
  }
  // output in cv::FAST order (row-major): rank = number of survivors with a smaller key
  uint32_t* out = cellKeys + ((long long)frame * P.totalCells + cell) * P.cellCap;
  for (int i = tid; i < nKept; i += FAST_THREADS) {
    const uint32_t e = keptList[i];
    const unsigned key = e & 0xffffu;
    int rank = 0;
    for (int j = 0; j < nKept; j++) rank += (keptList[j] & 0xffffu) < key;
    const unsigned kx = (key & 255u) + (unsigned)(3 + iniX - MIN_BORDER), ky = (key >> 8) + (unsigned)(3 + iniY - MIN_BORDER);
    out[rank] = kx | (ky << 12) | ((e >> 16) << 24);
  }
  if (tid == 0) *outCount = nKept;
}

// ------------------------------------------------------------------------------------------------
// k_octree: DistributeOctTree with one warp per (frame, level).  Lane 0 owns the linked list and
// the sort; all lanes cooperate on the stable 4-way key partition (DivideNode) and on the
// per-node best-response pick.  Keys live in two ping-pong global buffers; a node's keys are a
// contiguous range, children reuse the parent's range in the other buffer.
// ------------------------------------------------------------------------------------------------
static const int NIL = -1;
struct OctNode {
  short ulx, uly, brx, bry;
  int begin, count;
  short prev, next;
  short buf, noMore;
};
struct OctState {
  int head, tail, size, freeHead, nvsz, nprev, nToExpand, heapsorted;
};
struct NodeLess {
  const OctNode* nodes;
  GFS_HD bool operator()(int af, int as, int bf, int bs) const {
    if (af < bf) return true;
    if (af > bf) return false;
    return nodes[as].ulx < nodes[bs].ulx;
  }
};

static const int OCT_WARPS = 4;

__device__ __forceinline__ int oct_alloc(OctState* S, OctNode* nodes) {
  const int i = S->freeHead;
  S->freeHead = nodes[i].next;
  return i;
}
__device__ __forceinline__ void oct_free(OctState* S, OctNode* nodes, int i) {
  nodes[i].next = (short)S->freeHead;
  S->freeHead = i;
}
__device__ __forceinline__ void oct_push_front(OctState* S, OctNode* nodes, int i) {
  nodes[i].prev = NIL;
  nodes[i].next = (short)S->head;
  if (S->head != NIL) nodes[S->head].prev = (short)i;
  else S->tail = i;
  S->head = i;
  S->size++;
}
__device__ __forceinline__ void oct_push_back(OctState* S, OctNode* nodes, int i) {
  nodes[i].next = NIL;
  nodes[i].prev = (short)S->tail;
  if (S->tail != NIL) nodes[S->tail].next = (short)i;
  else S->head = i;
  S->tail = i;
  S->size++;
}
__device__ __forceinline__ void oct_erase(OctState* S, OctNode* nodes, int i) {
  const int p = nodes[i].prev, n = nodes[i].next;
  if (p != NIL) nodes[p].next = (short)n;
  else S->head = n;
  if (n != NIL) nodes[n].prev = (short)p;
  else S->tail = p;
  S->size--;
}

// ExtractorNode::DivideNode (:502-550) + the push_front / erase bookkeeping around it (:635-669).
__device__ void oct_divide(OctState* S, OctNode* nodes, int* vszF, int* vszS, uint32_t* keys0, uint32_t* keys1,
                           int cur, bool countExpand, int lane) {
  const OctNode nd = nodes[cur];
  const int halfX = (nd.brx - nd.ulx + 1) >> 1;  // ceil(float(w)/2), w >= 0
  const int halfY = (nd.bry - nd.uly + 1) >> 1;
  const int midX = nd.ulx + halfX, midY = nd.uly + halfY;
  const uint32_t* src = (nd.buf ? keys1 : keys0) + nd.begin;
  uint32_t* dst = (nd.buf ? keys0 : keys1) + nd.begin;
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  const unsigned lt = (1u << lane) - 1u;
  if (nd.count <= 128) {
    // the node's keys fit four registers per lane: one trip to memory instead of a count and a scatter pass
    uint32_t kk[4];
    int qq[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = lane + 32 * u;
      qq[u] = -1;
      kk[u] = 0;
      if (i < nd.count) kk[u] = src[i];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (lane + 32 * u < nd.count) {
        const int x = kk[u] & 0xFFF, y = (kk[u] >> 12) & 0xFFF;
        qq[u] = (x < midX) ? ((y < midY) ? 0 : 2) : ((y < midY) ? 1 : 3);
      }
      c0 += (qq[u] == 0); c1 += (qq[u] == 1); c2 += (qq[u] == 2); c3 += (qq[u] == 3);
    }
    c0 = __reduce_add_sync(0xffffffffu, c0);
    c1 = __reduce_add_sync(0xffffffffu, c1);
    c2 = __reduce_add_sync(0xffffffffu, c2);
    c3 = __reduce_add_sync(0xffffffffu, c3);
    int b0 = 0, b1 = c0, b2 = c0 + c1, b3 = c0 + c1 + c2;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (32 * u < nd.count) {  // warp-uniform
        const int q = qq[u];
        const unsigned m0 = __ballot_sync(0xffffffffu, q == 0);
        const unsigned m1 = __ballot_sync(0xffffffffu, q == 1);
        const unsigned m2 = __ballot_sync(0xffffffffu, q == 2);
        const unsigned m3 = __ballot_sync(0xffffffffu, q == 3);
        if (q == 0) dst[b0 + __popc(m0 & lt)] = kk[u];
        else if (q == 1) dst[b1 + __popc(m1 & lt)] = kk[u];
        else if (q == 2) dst[b2 + __popc(m2 & lt)] = kk[u];
        else if (q == 3) dst[b3 + __popc(m3 & lt)] = kk[u];
        b0 += __popc(m0); b1 += __popc(m1); b2 += __popc(m2); b3 += __popc(m3);
      }
    }
  } else {
    for (int i = lane; i < nd.count; i += 32) {
      const uint32_t k = src[i];
      const int x = k & 0xFFF, y = (k >> 12) & 0xFFF;
      const int q = (x < midX) ? ((y < midY) ? 0 : 2) : ((y < midY) ? 1 : 3);
      c0 += (q == 0); c1 += (q == 1); c2 += (q == 2); c3 += (q == 3);
    }
    c0 = __reduce_add_sync(0xffffffffu, c0);
    c1 = __reduce_add_sync(0xffffffffu, c1);
    c2 = __reduce_add_sync(0xffffffffu, c2);
    c3 = __reduce_add_sync(0xffffffffu, c3);
    int b0 = 0, b1 = c0, b2 = c0 + c1, b3 = c0 + c1 + c2;
    for (int s = 0; s < nd.count; s += 32) {
      const int i = s + lane;
      const bool valid = i < nd.count;
      uint32_t k = 0;
      int q = -1;
      if (valid) {
        k = src[i];
        const int x = k & 0xFFF, y = (k >> 12) & 0xFFF;
        q = (x < midX) ? ((y < midY) ? 0 : 2) : ((y < midY) ? 1 : 3);
      }
      const unsigned m0 = __ballot_sync(0xffffffffu, q == 0);
      const unsigned m1 = __ballot_sync(0xffffffffu, q == 1);
      const unsigned m2 = __ballot_sync(0xffffffffu, q == 2);
      const unsigned m3 = __ballot_sync(0xffffffffu, q == 3);
      if (q == 0) dst[b0 + __popc(m0 & lt)] = k;
      else if (q == 1) dst[b1 + __popc(m1 & lt)] = k;
      else if (q == 2) dst[b2 + __popc(m2 & lt)] = k;
      else if (q == 3) dst[b3 + __popc(m3 & lt)] = k;
      b0 += __popc(m0); b1 += __popc(m1); b2 += __popc(m2); b3 += __popc(m3);
    }
  }
  __syncwarp();
  if (lane == 0) {
    const int cnt[4] = {c0, c1, c2, c3};
    const int beg[4] = {nd.begin, nd.begin + c0, nd.begin + c0 + c1, nd.begin + c0 + c1 + c2};
    const short ulx[4] = {nd.ulx, (short)midX, nd.ulx, (short)midX};
    const short uly[4] = {nd.uly, nd.uly, (short)midY, (short)midY};
    const short brx[4] = {(short)midX, nd.brx, (short)midX, nd.brx};
    const short bry[4] = {(short)midY, (short)midY, nd.bry, nd.bry};
    for (int q = 0; q < 4; q++) {
      if (cnt[q] > 0) {
        const int i = oct_alloc(S, nodes);
        nodes[i].ulx = ulx[q]; nodes[i].uly = uly[q]; nodes[i].brx = brx[q]; nodes[i].bry = bry[q];
        nodes[i].begin = beg[q]; nodes[i].count = cnt[q];
        nodes[i].buf = (short)(nd.buf ^ 1);
        nodes[i].noMore = (short)(cnt[q] == 1);
        oct_push_front(S, nodes, i);
        if (cnt[q] > 1) {
          if (countExpand) S->nToExpand++;
          vszF[S->nvsz] = cnt[q];
          vszS[S->nvsz] = i;
          S->nvsz++;
        }
      }
    }
    oct_erase(S, nodes, cur);
    oct_free(S, nodes, cur);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(OCT_WARPS * 32) k_octree(OrbDev P, int batch, const uint32_t* __restrict__ cellKeys,
                                                           const int* __restrict__ cellCount, uint32_t* keysA,
                                                           uint32_t* keysB, uint32_t* __restrict__ sel,
                                                           int* __restrict__ selCount, int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t osm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int inst = blockIdx.x * OCT_WARPS + warp;
  if (inst >= batch * P.nlevels) return;
  // heavy (fine) levels first so the tail of the grid is made of cheap instances
  const int l = inst / batch, frame = inst - l * batch;
  const LevelDev& L = P.lv[l];
  const int nodeCap = P.nodeCap;
  const size_t perWarp = sizeof(OctState) + (size_t)nodeCap * (sizeof(OctNode) + 4 * sizeof(int));
  uint8_t* base = osm + warp * ((perWarp + 15) / 16 * 16);
  OctState* S = (OctState*)base;
  OctNode* nodes = (OctNode*)(base + sizeof(OctState));
  int* vszF = (int*)(nodes + nodeCap);
  int* vszS = vszF + nodeCap;
  int* prvF = vszS + nodeCap;
  int* prvS = prvF + nodeCap;

  uint32_t* keys0 = keysA + (long long)frame * P.keysPerFrame + L.keyOff;
  uint32_t* keys1 = keysB + (long long)frame * P.keysPerFrame + L.keyOff;
  const int N = L.nFeat;

  // ---- gather the level's candidates in reference order (cell row-major, in-cell row-major)
  const int nCells = L.nCols * L.nRows;
  const int* cc = cellCount + (long long)frame * P.totalCells + L.cellBase;
  const uint32_t* ck = cellKeys + ((long long)frame * P.totalCells + L.cellBase) * P.cellCap;
  int ncand = 0;
  for (int cb = 0; cb < nCells; cb += 32) {
    const int c = cb + lane;
    const int n = (c < nCells) ? cc[c] : 0;
    int inc = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    // every lane copies its own cell: 32 independent copy chains instead of one cell at a time
    const int excl = ncand + inc - n;
    const uint32_t* s = ck + (long long)c * P.cellCap;
#pragma unroll 4
    for (int k = 0; k < n; k++) keys1[excl + k] = s[k];
    ncand += __shfl_sync(0xffffffffu, inc, 31);
  }
  __syncwarp();

  // ---- initial nodes (:573-612)
  if (lane == 0) {
    S->head = NIL; S->tail = NIL; S->size = 0; S->nvsz = 0; S->nprev = 0; S->nToExpand = 0; S->heapsorted = 0;
    for (int i = 0; i < nodeCap; i++) nodes[i].next = (short)((i + 1 < nodeCap) ? i + 1 : NIL);
    S->freeHead = 0;
  }
  __syncwarp();
  {
    int off = 0;
    for (int i = 0; i < L.nIni; i++) {
      // stable filter keys1 -> keys0 of the keys whose bucket (int)(x / hX) is i
      int w = off;
      for (int s = 0; s < ncand; s += 32) {
        const int k = s + lane;
        bool take = false;
        uint32_t key = 0;
        if (k < ncand) {
          key = keys1[k];
          take = (L.nIni == 1) || ((int)__fdiv_rn((float)(key & 0xFFF), L.hX) == i);
        }
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (take) keys0[w + __popc(m & ((1u << lane) - 1u))] = key;
        w += __popc(m);
      }
      const int cnt = w - off;
      if (lane == 0 && cnt > 0) {  // empty initial nodes are erased (:605-606)
        const int n = oct_alloc(S, nodes);
        nodes[n].ulx = (short)(int)__fmul_rn(L.hX, (float)i);
        nodes[n].brx = (short)(int)__fmul_rn(L.hX, (float)(i + 1));
        nodes[n].uly = 0;
        nodes[n].bry = (short)(L.maxBorderY - MIN_BORDER);
        nodes[n].begin = off; nodes[n].count = cnt; nodes[n].buf = 0;
        nodes[n].noMore = (short)(cnt == 1);
        oct_push_back(S, nodes, n);
      }
      off = w;
    }
  }
  __syncwarp();

  // ---- subdivision (:614-747)
  bool finish = false;
  while (!finish) {
    int prevSize = S->size;
    int cur = S->head;
    __syncwarp();
    if (lane == 0) { S->nvsz = 0; S->nToExpand = 0; }
    __syncwarp();
    while (cur != NIL) {
      const int nxt = nodes[cur].next;
      const bool nm = nodes[cur].noMore != 0;
      __syncwarp();
      if (!nm) oct_divide(S, nodes, vszF, vszS, keys0, keys1, cur, true, lane);
      cur = nxt;
    }
    const int size = S->size;
    if (size >= N || size == prevSize) {
      finish = true;
    } else if (size + S->nToExpand * 3 > N) {
      while (!finish) {
        prevSize = S->size;
        __syncwarp();
        if (lane == 0) {
          const int n = S->nvsz;
          for (int i = 0; i < n; i++) { prvF[i] = vszF[i]; prvS[i] = vszS[i]; }
          S->nprev = n;
          S->nvsz = 0;
          GccSort<NodeLess> srt{{prvF, prvS}, NodeLess{nodes}};
          srt.sort(n, &S->heapsorted);
        }
        __syncwarp();
        const int np = S->nprev;
        for (int j = np - 1; j >= 0; j--) {
          oct_divide(S, nodes, vszF, vszS, keys0, keys1, prvS[j], false, lane);
          if (S->size >= N) break;
        }
        if (S->size >= N || S->size == prevSize) finish = true;
      }
    }
  }

  // ---- best response per node, in list order (:749-767); first maximum wins.  Lane 0 flattens the list,
  // then every lane takes whole nodes (a node holds a handful of keys by now); only nodes too large for
  // one lane are reduced by the whole warp.
  uint32_t* out = sel + (long long)frame * P.selPerFrame + L.selOff;
  __syncwarp();
  if (lane == 0) {
    int n = 0;
    for (int cur = S->head; cur != NIL; cur = nodes[cur].next) vszS[n++] = cur;
    S->nvsz = n;
  }
  __syncwarp();
  const int nout = S->nvsz;
  for (int jb = 0; jb < nout; jb += 32) {
    const int j = jb + lane;
    bool big = false;
    if (j < nout) {
      const OctNode nd = nodes[vszS[j]];
      if (nd.count <= 48) {
        const uint32_t* src = (nd.buf ? keys1 : keys0) + nd.begin;
        uint32_t bk = src[0];
#pragma unroll 4
        for (int i = 1; i < nd.count; i++) {
          const uint32_t k = src[i];
          if ((k >> 24) > (bk >> 24)) bk = k;
        }
        if (j < L.selCap) out[j] = bk;
      } else {
        big = true;
      }
    }
    unsigned bm = __ballot_sync(0xffffffffu, big);
    while (bm) {
      const int jj = jb + __ffs(bm) - 1;
      bm &= bm - 1;
      const OctNode nd = nodes[vszS[jj]];
      const uint32_t* src = (nd.buf ? keys1 : keys0) + nd.begin;
      // order by (score desc, index asc): pack score in the high bits, inverted index below
      unsigned long long bestv = 0ull;
      for (int i = lane; i < nd.count; i += 32) {
        const uint32_t k = src[i];
        const unsigned long long v = ((unsigned long long)(k >> 24) << 32) | (unsigned)(0x7fffffff - i);
        bestv = max(bestv, v);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) bestv = max(bestv, __shfl_xor_sync(0xffffffffu, bestv, o));
      const int bi = 0x7fffffff - (int)(bestv & 0xffffffffu);
      if (lane == 0 && jj < L.selCap) out[jj] = src[bi];
    }
  }
  if (lane == 0) {
    selCount[frame * P.nlevels + l] = min(nout, L.selCap);
    if (nout > L.selCap) atomicOr(status, 1);
    if (S->heapsorted) atomicOr(status, 2);
  }
}

// ------------------------------------------------------------------------------------------------
// k_blur7: cv::GaussianBlur 7x7 sigma 2 on u8 = fixed-point separable [18,34,48,56,48,34,18]/256.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
  return p;
}

static const int BL_TW = 64, BL_TH = 32;
__global__ void __launch_bounds__(256) k_blur7(OrbDev P, int l, const uint8_t* __restrict__ img0, long long img_stride,
                                               int pitch0, const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur) {
  __shared__ uint8_t s_in[BL_TH + 6][BL_TW + 8];
  __shared__ uint16_t s_h[BL_TH + 6][BL_TW];
  const LevelDev& L = P.lv[l];
  const int frame = blockIdx.z;
  const uint8_t* src;
  int pitch;
  if (l == 0) { src = img0 + (long long)frame * img_stride; pitch = pitch0; }
  else { src = pyr + (long long)frame * P.pyrStride + L.off; pitch = L.pitch; }
  const int x0 = blockIdx.x * BL_TW, y0 = blockIdx.y * BL_TH;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < (BL_TH + 6) * (BL_TW + 6); idx += 256) {
    const int yy = idx / (BL_TW + 6), xx = idx - yy * (BL_TW + 6);
    const int gy = reflect101(y0 + yy - 3, L.h), gx = reflect101(x0 + xx - 3, L.w);
    s_in[yy][xx] = __ldg(src + (long long)gy * pitch + gx);
  }
  __syncthreads();
  for (int idx = tid; idx < (BL_TH + 6) * BL_TW; idx += 256) {
    const int yy = idx / BL_TW, xx = idx - yy * BL_TW;
    const uint8_t* p = &s_in[yy][xx];
    const int acc = 18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3];
    s_h[yy][xx] = (uint16_t)acc;
  }
  __syncthreads();
  uint8_t* dst = blur + (long long)frame * P.pyrStride + L.off;
  for (int idx = tid; idx < BL_TH * BL_TW; idx += 256) {
    const int yy = idx / BL_TW, xx = idx - yy * BL_TW;
    const int gx = x0 + xx, gy = y0 + yy;
    if (gx < L.w && gy < L.h) {
      const unsigned acc = 18u * (s_h[yy][xx] + s_h[yy + 6][xx]) + 34u * (s_h[yy + 1][xx] + s_h[yy + 5][xx]) +
                           48u * (s_h[yy + 2][xx] + s_h[yy + 4][xx]) + 56u * s_h[yy + 3][xx];
      dst[(long long)gy * L.pitch + gx] = (uint8_t)((acc + 32768u) >> 16);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_orient_desc: one warp per selected keypoint slot.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_atan2_deg(const OrbDev& P, float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  float a;
  if (ax >= ay) {
    const float c = __fdiv_rn(ay, __fadd_rn(ax, (float)DBL_EPSILON));
    const float c2 = __fmul_rn(c, c);
    float t = __fadd_rn(__fmul_rn(P.p7, c2), P.p5);
    t = __fadd_rn(__fmul_rn(t, c2), P.p3);
    t = __fadd_rn(__fmul_rn(t, c2), P.p1);
    a = __fmul_rn(t, c);
  } else {
    const float c = __fdiv_rn(ax, __fadd_rn(ay, (float)DBL_EPSILON));
    const float c2 = __fmul_rn(c, c);
    float t = __fadd_rn(__fmul_rn(P.p7, c2), P.p5);
    t = __fadd_rn(__fmul_rn(t, c2), P.p3);
    t = __fadd_rn(__fmul_rn(t, c2), P.p1);
    a = __fsub_rn(90.f, __fmul_rn(t, c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

static const int OD_WARPS = 8;
static const int PATCH_R = 21;                 // 18 (largest rounded rotated pattern radius) + 3 (blur taps)
static const int PATCH_W = 2 * PATCH_R + 1;    // 43
static const int RAW_PITCH = 52;               // 13 words per row: odd, so column walks hit distinct banks
static const int HB_W = 37, HB_PITCH = 40;     // horizontally blurred rows, radius 18, u16 (10 quads per row)

// The warp stages the keypoint's 43x43 neighbourhood of the level image in shared memory (one
// coalesced sweep), takes the IC_Angle moments from it, runs the horizontal pass of the 7x7 sigma-2
// fixed-point Gaussian (ORBextractor.cc:1188-1189) over it and evaluates the vertical pass only at
// the 512 rotated sample positions -- the blurred level image is never materialised in HBM.
__global__ void __launch_bounds__(OD_WARPS * 32) k_orient_desc(OrbDev P, const uint8_t* __restrict__ img0,
                                                               long long img_stride, int pitch0,
                                                               const uint8_t* __restrict__ pyr,
                                                               const uint4* __restrict__ pattern,
                                                               const uint32_t* __restrict__ sel,
                                                               const int* __restrict__ selCount,
                                                               GfsKeyPoint* __restrict__ out_kp,
                                                               uint8_t* __restrict__ out_desc, int* __restrict__ out_n,
                                                               int* __restrict__ out_mono) {
  __shared__ __align__(16) uint8_t s_raw[OD_WARPS][PATCH_W * RAW_PITCH];
  __shared__ __align__(16) uint16_t s_hb[OD_WARPS][PATCH_W * HB_PITCH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = blockIdx.x * OD_WARPS + warp;
  const int frame = blockIdx.y;
  if (slot >= P.selPerFrame) return;
  int l = 0;
  while (l + 1 < P.nlevels && slot >= P.lv[l + 1].selOff) l++;
  const LevelDev& L = P.lv[l];
  const int i = slot - L.selOff;
  const int* sc = selCount + frame * P.nlevels;
  int base = 0, total = 0;
  for (int k = 0; k < P.nlevels; k++) {
    const int c = sc[k];
    if (k < l) base += c;
    total += c;
  }
  if (slot == 0 && lane == 0) { out_n[frame] = total; out_mono[frame] = total; }
  if (i >= sc[l]) return;
  const uint32_t key = sel[(long long)frame * P.selPerFrame + slot];
  const int cx = (int)(key & 0xFFF) + MIN_BORDER, cy = (int)((key >> 12) & 0xFFF) + MIN_BORDER;
  const int score = (int)(key >> 24);
  const uint8_t* src;
  int pitch;
  if (l == 0) { src = img0 + (long long)frame * img_stride; pitch = pitch0; }
  else { src = pyr + (long long)frame * P.pyrStride + L.off; pitch = L.pitch; }

  // ---- stage the raw patch: rows cy-21..cy+21, columns cx-21..cx+21 (REFLECT_101 at the image border)
  uint8_t* raw = s_raw[warp];
  const int x0 = cx - PATCH_R, y0 = cy - PATCH_R;
  int shift = 0;
  const bool interior = x0 >= 0 && y0 >= 0 && cx + PATCH_R < L.w && cy + PATCH_R < L.h && ((pitch & 3) == 0) &&
                        ((((size_t)src) & 3) == 0);
  bool fast = false;
  if (interior) {
    shift = x0 & 3;
    const int nW = (shift + PATCH_W + 3) >> 2;  // <= 12 words
    if (x0 - shift + 4 * nW <= pitch) {
      fast = true;
      const uint32_t* s32 = (const uint32_t*)(src + (long long)y0 * pitch + (x0 - shift));
      const int p32 = pitch >> 2;
      uint32_t* d32 = (uint32_t*)raw;
      // 16 lanes per row (nW <= 12 of them active), two rows per step: no index division
      const int c = lane & 15;
      if (c < nW) {
        for (int r = lane >> 4; r < PATCH_W; r += 2) d32[r * (RAW_PITCH / 4) + c] = __ldg(s32 + (long long)r * p32 + c);
      }
    }
  }
  if (!fast) {
    // border keypoints: the lane's (up to) two reflected columns once, then row by row
    shift = 0;
    const int gx0 = reflect101(x0 + lane, L.w), gx1 = reflect101(x0 + min(lane + 32, PATCH_W - 1), L.w);
    for (int r = 0; r < PATCH_W; r++) {
      const uint8_t* row = src + (long long)reflect101(y0 + r, L.h) * pitch;
      raw[r * RAW_PITCH + lane] = __ldg(row + gx0);
      if (lane < PATCH_W - 32) raw[r * RAW_PITCH + lane + 32] = __ldg(row + gx1);
    }
  }
  __syncwarp();
  const uint8_t* rp = raw + shift;  // patch pixel (r, c) at rp[r * RAW_PITCH + c]

  // ---- IC_Angle (:71-95): integer moments over the radius-15 disc, one column per lane
  int m10 = 0, m01 = 0;
  if (lane < 2 * HALF_PATCH + 1) {
    const int u = lane - HALF_PATCH;
    const int d = P.umax[u < 0 ? -u : u];  // the disc is symmetric: column u spans |v| <= umax[|u|]
    const uint8_t* col = rp + PATCH_R * RAW_PITCH + PATCH_R + u;
    int cs = 0;
    for (int v = -d; v <= d; v++) {
      const int val = col[v * RAW_PITCH];
      m01 += v * val;
      cs += val;
    }
    m10 = u * cs;
  }
  m10 = __reduce_add_sync(0xffffffffu, m10);
  m01 = __reduce_add_sync(0xffffffffu, m01);
  const float angle = fast_atan2_deg(P, (float)m01, (float)m10);

  // ---- horizontal blur pass: hb[r][c] = sum_k K[k] * raw[r][c + k], c in [0, 40) (37 used).  The aligned
  // words of the raw row are realigned to the patch origin (funnel shifts), every output is two 4-tap
  // byte dot products (dp4a) with the kernel (18 34 48 56 | 48 34 18 0).
  uint16_t* hb = s_hb[warp];
  {
    const unsigned sh8 = 8u * (unsigned)shift;
    const uint32_t* raw32 = (const uint32_t*)raw;
    const uint32_t kA = 18u | (34u << 8) | (48u << 16) | (56u << 24), kB = 48u | (34u << 8) | (18u << 16);
    // an item is eight adjacent outputs (five aligned words in, one 16-byte store out)
    for (int item = lane; item < PATCH_W * (HB_PITCH / 8); item += 32) {
      const int r = item / (HB_PITCH / 8), q8 = item - r * (HB_PITCH / 8);
      const uint32_t* w = raw32 + r * (RAW_PITCH / 4) + 2 * q8;
      const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
      const uint32_t s0 = __funnelshift_r(w0, w1, sh8), s1 = __funnelshift_r(w1, w2, sh8), s2 = __funnelshift_r(w2, w3, sh8),
                     s3 = __funnelshift_r(w3, w4, sh8);
      uint32_t a[12];
      a[0] = s0; a[1] = __byte_perm(s0, s1, 0x4321); a[2] = __byte_perm(s0, s1, 0x5432); a[3] = __byte_perm(s0, s1, 0x6543);
      a[4] = s1; a[5] = __byte_perm(s1, s2, 0x4321); a[6] = __byte_perm(s1, s2, 0x5432); a[7] = __byte_perm(s1, s2, 0x6543);
      a[8] = s2; a[9] = __byte_perm(s2, s3, 0x4321); a[10] = __byte_perm(s2, s3, 0x5432); a[11] = __byte_perm(s2, s3, 0x6543);
      unsigned o[8];
#pragma unroll
      for (int k = 0; k < 8; k++) o[k] = __dp4a(a[k + 4], kB, __dp4a(a[k], kA, 0u));
      *reinterpret_cast<uint4*>(hb + r * HB_PITCH + 8 * q8) =
          make_uint4(o[0] | (o[1] << 16), o[2] | (o[3] << 16), o[4] | (o[5] << 16), o[6] | (o[7] << 16));
    }
  }
  __syncwarp();

  // ---- rBRIEF (:99-160): lane b computes descriptor byte b (tests 8b .. 8b+7); the vertical blur
  // pass is evaluated at the sample positions only: ((sum_k K[k] * hb[y+k][x]) + 2^15) >> 16
  const float ar = __fmul_rn(angle, P.factorPI);
  double sd, cd;
  sincos((double)ar, &sd, &cd);
  const float a = (float)cd, b = (float)sd;
  // the lane's 16 pattern points (32 signed bytes) live in 8 registers; a per-lane indexed
  // __constant__ read would serialise in the address-divergence unit (ncu: adu pipe 88 %)
  const uint4 pw0 = __ldg(pattern + 2 * lane), pw1 = __ldg(pattern + 2 * lane + 1);
  const unsigned pw[8] = {pw0.x, pw0.y, pw0.z, pw0.w, pw1.x, pw1.y, pw1.z, pw1.w};
  int val = 0;
#pragma unroll
  for (int t = 0; t < 8; t++) {
    int s[2];
#pragma unroll
    for (int p = 0; p < 2; p++) {
      // point index 2t+p -> bytes 4t+2p (x) and 4t+2p+1 (y) of the lane's record
      const unsigned w = pw[t];
      const float px = (float)((int)(w << (24 - 16 * p)) >> 24), py = (float)((int)(w << (16 - 16 * p)) >> 24);
      const float r1 = __fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a));
      const float r2 = __fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b));
      const int ry = (int)roundf(r1), rx = (int)roundf(r2);
      const uint16_t* h = hb + (ry + 18) * HB_PITCH + (rx + 18);  // rows ry+18 .. ry+24 = blur taps -3 .. +3
      const unsigned acc = 18u * (h[0] + h[6 * HB_PITCH]) + 34u * (h[HB_PITCH] + h[5 * HB_PITCH]) +
                           48u * (h[2 * HB_PITCH] + h[4 * HB_PITCH]) + 56u * h[3 * HB_PITCH];
      s[p] = (int)((acc + 32768u) >> 16);
    }
    val |= (s[0] < s[1]) << t;
  }
  const long long o = (long long)frame * P.kpStride + base + i;
  out_desc[o * 32 + lane] = (uint8_t)val;
  if (lane == 0) {
    GfsKeyPoint kp;
    kp.x = (l == 0) ? (float)cx : __fmul_rn((float)cx, L.scale);
    kp.y = (l == 0) ? (float)cy : __fmul_rn((float)cy, L.scale);
    kp.size = (float)L.patchSize;
    kp.angle = angle;
    kp.response = (float)score;
    kp.octave = l;
    out_kp[o] = kp;
  }
}

// ------------------------------------------------------------------------------------------------
// k_pack_lapping: mono keypoints from the front, lapping-area keypoints from the back (:1208-1217).
// One CTA per frame; in-place via a copy in global scratch.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_pack_lapping(int kpStride, int lap0, int lap1, const GfsKeyPoint* __restrict__ tkp,
                                                       const uint8_t* __restrict__ tdesc, GfsKeyPoint* __restrict__ out_kp,
                                                       uint8_t* __restrict__ out_desc, const int* __restrict__ out_n,
                                                       int* __restrict__ out_mono) {
  __shared__ int s_w[32];
  __shared__ int s_carry;
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = out_n[frame];
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int b = 0; b < n; b += 1024) {
    const int i = b + tid;
    GfsKeyPoint kp;
    int lapf = 0;
    if (i < n) {
      kp = tkp[(long long)frame * kpStride + i];
      lapf = (kp.x >= (float)lap0 && kp.x <= (float)lap1);
    }
    const unsigned m = __ballot_sync(0xffffffffu, lapf);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    int before = s_carry;
    for (int w2 = 0; w2 < warp; w2++) before += s_w[w2];
    before += __popc(m & ((1u << lane) - 1u));
    if (i < n) {
      const int dstI = lapf ? (n - 1 - before) : (i - before);
      out_kp[(long long)frame * kpStride + dstI] = kp;
      const uint4* s4 = (const uint4*)(tdesc + ((long long)frame * kpStride + i) * 32);
      uint4* d4 = (uint4*)(out_desc + ((long long)frame * kpStride + dstI) * 32);
      d4[0] = s4[0];
      d4[1] = s4[1];
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w2 = 0; w2 < 32; w2++) t += s_w[w2];
      s_carry += t;
    }
    __syncthreads();
  }
  if (tid == 0) out_mono[frame] = n - s_carry;
}

}  // namespace gfs

// ================================================================================================
// Host side
// ================================================================================================
using namespace gfs;

struct GfsOrb {
  int nfeatures, nlevels, iniTh, minTh;
  double scaleFactor;
  float sf[MAX_LEVELS], isf[MAX_LEVELS];
  int nPerLevel[MAX_LEVELS];
  int maxW, maxH, maxBatch;
  int geomW = 0, geomH = 0;
  OrbDev dev;
  DevBuf d_pyr, d_blur, d_cellKeys, d_cellCount, d_keysA, d_keysB, d_sel, d_selCount, d_tabs, d_status;
  DevBuf d_in, d_okp, d_odesc, d_on, d_omono, d_tkp, d_tdesc, d_pattern;
  PinnedBuf h_in, h_okp, h_odesc, h_on;
  size_t fastSmem = 0, octSmem = 0, pyrSmem = 0, pyr3Smem = 0;
  bool fastV1 = getenv("GFS_FAST_V1") != nullptr;  // first-generation FAST kernel (A/B runs); same results
  int pyr3Rows = 0, pyr3PitchF = 0;  // k_pyr_level3 shared-memory geometry; pyr3Rows == 0: generic kernel
  unsigned lastTmapMask = 0;          // levels whose FAST ROIs the last call staged through a tensor map
  DevBuf d_tabs3, d_cellTab;
  // optional per-stage CUDA-event timing (bench roofline): pyramid, fast, octree, blur, orient/desc, pack
  bool profiling = false;
  cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // last batch (debug hooks)
  cudaStream_t auxStream = nullptr;  // second stream of gfs_orb_extract_batch_device
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  const uint8_t* lastImgs = nullptr;
  long long lastStride = 0;
  int lastPitch = 0, lastBatch = 0;
};

static inline int cv_round_f(double v) { return (int)std::nearbyint(v); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tmap_encoder() {
  static EncodeTiledFn fn = [] {
    if (const char* e = getenv("GFS_ORB_TMAP")) if (atoi(e) == 0) return (EncodeTiledFn) nullptr;   // A/B runs: per-row bulk copies
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return (EncodeTiledFn) nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}
// level images of `frames` frames as a rank-3 u8 tensor (x, y, frame) with a roiPitch x roiRows x 1 box; false: no map possible
static bool encode_level_map(CUtensorMap* m, const uint8_t* base, int w, int hgt, int frames, long long pitch, long long frameStride,
                             int boxW, int boxH) {
  EncodeTiledFn enc = tmap_encoder();
  long long fs = frameStride;
  if (frames == 1 && (fs < pitch * hgt || (fs & 15))) fs = (pitch * hgt + 15) / 16 * 16;   // a single frame: any valid stride will do
  if (!enc || (((uintptr_t)base | (uintptr_t)pitch | (uintptr_t)fs) & 15) != 0 || fs < pitch * hgt || pitch < w || boxW > 256 ||
      boxH > 256 || (boxW & 15))
    return false;
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)hgt, (cuuint64_t)frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fs};
  const cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1u};
  const cuuint32_t es[3] = {1u, 1u, 1u};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static void area_tab(int ssize, int dsize, std::vector<AreaEntry>& out) {
  // cv::resize INTER_AREA table (oracle/orb_oracle.cpp area_tab)
  const double scale = 1.0 / ((double)dsize / (double)ssize);
  for (int dx = 0; dx < dsize; dx++) {
    AreaEntry e{};
    e.s0 = -1;
    const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    const double cw = std::min(scale, ssize - fsx1);
    int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
    sx2 = std::min(sx2, ssize - 1);
    sx1 = std::min(sx1, sx2);
    auto push = [&](int s, float a) {
      if (e.n == 0) e.s0 = s;
      if (e.n < 4) e.a[e.n] = a;
      e.n++;
    };
    if (sx1 - fsx1 > 1e-3) push(sx1 - 1, (float)((sx1 - fsx1) / cw));
    for (int sx = sx1; sx < sx2; sx++) push(sx, (float)(1.0 / cw));
    if (fsx2 - sx2 > 1e-3) push(sx2, (float)(std::min(std::min(fsx2 - sx2, 1.), cw) / cw));
    out.push_back(e);
  }
}

static int orb_set_geometry(GfsOrb* h, int w, int ht) {
  if (h->geomW == w && h->geomH == ht) return GFS_OK;
  OrbDev D = h->dev;  // built on a copy; committed only when everything succeeded
  D.nlevels = h->nlevels;
  D.iniTh = h->iniTh;
  D.minTh = h->minTh;
  std::vector<AreaEntry> tabs;
  long long off = 0;
  int cellBase = 0, keyOff = 0, selOff = 0, cellCap = 1, nodeCap = 8, maxWC = 0, maxHC = 0;
  for (int l = 0; l < h->nlevels; l++) {
    LevelDev& L = D.lv[l];
    L.w = cv_round_f((float)w * h->isf[l]);
    L.h = cv_round_f((float)ht * h->isf[l]);
    L.pitch = (int)align_up(L.w, 16);
    L.off = off;
    off += (long long)L.pitch * L.h;
    off = (long long)align_up((size_t)off, 256);
    L.maxBorderX = L.w - EDGE_THRESHOLD + 3;
    L.maxBorderY = L.h - EDGE_THRESHOLD + 3;
    const float width = (float)(L.maxBorderX - MIN_BORDER), height = (float)(L.maxBorderY - MIN_BORDER);
    L.nCols = (int)(width / 35.f);
    L.nRows = (int)(height / 35.f);
    if (L.nCols < 1 || L.nRows < 1) {
      set_error("image %dx%d too small: pyramid level %d (%dx%d) has no 35-px FAST cell", w, ht, l, L.w, L.h);
      return GFS_ERR_INVALID;
    }
    L.wCell = (int)std::ceil(width / L.nCols);
    L.hCell = (int)std::ceil(height / L.nRows);
    if (L.w > 4095 || L.h > 4095) {
      set_error("level %d size %dx%d exceeds the 12-bit key range", l, L.w, L.h);
      return GFS_ERR_INVALID;
    }
    maxWC = std::max(maxWC, L.wCell);
    maxHC = std::max(maxHC, L.hCell);
    L.cellBase = cellBase;
    cellBase += L.nCols * L.nRows;
    cellCap = std::max(cellCap, ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2));
    L.nFeat = h->nPerLevel[l];
    L.nIni = (int)std::round((float)(L.maxBorderX - MIN_BORDER) / (L.maxBorderY - MIN_BORDER));
    if (L.nIni == 0) L.nIni = 1;
    L.hX = (float)(L.maxBorderX - MIN_BORDER) / L.nIni;
    L.selOff = selOff;
    L.selCap = std::max(L.nFeat + 3, 4 * L.nIni);
    selOff += L.selCap;
    nodeCap = std::max(nodeCap, L.selCap + 8);
    L.scale = h->sf[l];
    L.patchSize = (int)(PATCH_SIZE * h->sf[l]);
    if (l > 0) {
      const LevelDev& Pv = D.lv[l - 1];
      L.tabX = (int)tabs.size();
      area_tab(Pv.w, L.w, tabs);
      L.tabY = (int)tabs.size();
      area_tab(Pv.h, L.h, tabs);
    } else {
      L.tabX = L.tabY = 0;
    }
  }
  int maxSrcRows = 1;
  for (int l = 1; l < h->nlevels; l++) {
    const LevelDev& L = D.lv[l];
    for (int dy0 = 0; dy0 < L.h; dy0 += PYR_TH) {
      const AreaEntry& f = tabs[L.tabY + dy0];
      const AreaEntry& e = tabs[L.tabY + std::min(dy0 + PYR_TH, L.h) - 1];
      maxSrcRows = std::max(maxSrcRows, e.s0 + e.n - f.s0);
    }
  }
  h->pyrSmem = (size_t)maxSrcRows * PYR_TW * sizeof(float);
  // 3-tap tables (same indexing as `tabs`) when no entry needs more than 3 source pixels
  std::vector<Area3> tabs3;
  bool three = !tabs.empty();
  for (const AreaEntry& e : tabs) three = three && e.n >= 1 && e.n <= 3;
  int pyr3Rows = 0, pyr3PitchF = 0;
  if (three) {
    tabs3.resize(tabs.size());
    auto pad = [&](int first, int count, int ssize) {
      for (int i = first; i < first + count && three; i++) {
        const AreaEntry& e = tabs[i];
        float a[3] = {0.f, 0.f, 0.f};
        int s0 = e.s0;
        if (ssize < 3) { three = false; break; }
        const int shift = std::max(0, s0 + 3 - ssize);  // zero taps go in front when the window would leave the image
        s0 -= shift;
        for (int k = 0; k < e.n; k++) a[k + shift] = e.a[k];
        tabs3[i] = Area3{s0, a[0], a[1], a[2]};
      }
    };
    for (int l = 1; l < h->nlevels; l++) {
      pad(D.lv[l].tabX, D.lv[l].w, D.lv[l - 1].w);
      pad(D.lv[l].tabY, D.lv[l].h, D.lv[l - 1].h);
    }
    for (int l = 1; l < h->nlevels && three; l++) {
      const LevelDev& L = D.lv[l];
      for (int dy0 = 0; dy0 < L.h; dy0 += PYR_TH)
        pyr3Rows = std::max(pyr3Rows, tabs3[L.tabY + std::min(dy0 + PYR_TH, L.h) - 1].s0 + 3 - tabs3[L.tabY + dy0].s0);
      for (int dx0 = 0; dx0 < L.w; dx0 += PYR_TW) {
        const int xs = tabs3[L.tabX + dx0].s0 & ~3;
        const int words = ((tabs3[L.tabX + std::min(dx0 + PYR_TW, L.w) - 1].s0 + 2 - xs) >> 2) + 1;
        if (words > 32) three = false;
        pyr3PitchF = std::max(pyr3PitchF, 4 * words);
      }
    }
  }
  if (!three) pyr3Rows = pyr3PitchF = 0;
  const size_t pyr3Smem = (size_t)pyr3Rows * (pyr3PitchF + PYR_TW) * sizeof(float);
  for (const AreaEntry& e : tabs)
    if (e.n > 4 || e.n < 1) {
      set_error("scale factor outside the supported (1, 3) range");
      return GFS_ERR_INVALID;
    }
  for (int l = 0; l < h->nlevels; l++) {
    LevelDev& L = D.lv[l];
    L.keyOff = keyOff;
    keyOff += L.nCols * L.nRows * cellCap;
  }
  D.pyrStride = off;
  D.cellCap = cellCap;
  D.totalCells = cellBase;
  D.keysPerFrame = keyOff;
  D.selPerFrame = selOff;
  D.nodeCap = nodeCap;
  // ROI row pitch: room for the 16-byte alignment shift of the first column and the rounded-up row
  // of the bulk copy; a pitch of 4 (mod 8) words keeps eight consecutive rows on distinct banks
  D.roiPitch = (int)align_up(maxWC + 6 + 15, 16);
  if (((D.roiPitch >> 2) & 7) != 4) D.roiPitch += 16;
  if (((D.roiPitch >> 2) & 7) != 4) D.roiPitch += 16;
  D.roiRows = maxHC + 6;
  D.scPitch = (int)align_up(maxWC + 2, 4);
  D.offSc = (int)align_up((size_t)D.roiRows * D.roiPitch + 4, 16);
  D.offCorner = D.offSc + (int)align_up((size_t)(maxHC + 2) * D.scPitch, 16);
  D.offCand = D.offCorner + (int)align_up((size_t)2 * maxWC * maxHC, 16);
  D.offKept = D.offCand + (int)align_up((size_t)2 * maxWC * maxHC, 16);
  if (selOff > D.kpStride) {
    set_error("internal: selected-slot count %d exceeds keypoint stride %d", selOff, D.kpStride);
    return GFS_ERR_INVALID;
  }
  if (nodeCap > 32000) {
    set_error("nfeatures too large for the quadtree node pool");
    return GFS_ERR_CAPACITY;
  }
  const float scale = (float)(180.0 / M_PI);
  D.p1 = 0.9997878412794807f * scale;
  D.p3 = -0.3258083974640975f * scale;
  D.p5 = 0.1555786518463281f * scale;
  D.p7 = -0.04432655554792128f * scale;
  D.factorPI = (float)(M_PI / 180.f);
  {  // umax (ctor :464-478)
    int v, v0, vmax = (int)std::floor(HALF_PATCH * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(HALF_PATCH * std::sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH * HALF_PATCH;
    for (v = 0; v <= vmax; ++v) D.umax[v] = cv_round_f(std::sqrt(hp2 - v * v));
    for (v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
      while (D.umax[v0] == D.umax[v0 + 1]) ++v0;
      D.umax[v] = v0;
      ++v0;
    }
  }
  const size_t B = (size_t)h->maxBatch;
  int rc;
  if ((rc = h->d_pyr.reserve(B * D.pyrStride))) return rc;
  if ((rc = h->d_cellKeys.reserve(B * D.totalCells * (size_t)D.cellCap * 4))) return rc;
  if ((rc = h->d_cellCount.reserve(B * D.totalCells * 4))) return rc;
  if ((rc = h->d_keysA.reserve(B * (size_t)D.keysPerFrame * 4))) return rc;
  if ((rc = h->d_keysB.reserve(B * (size_t)D.keysPerFrame * 4))) return rc;
  if ((rc = h->d_sel.reserve(B * (size_t)D.selPerFrame * 4))) return rc;
  if ((rc = h->d_selCount.reserve(B * MAX_LEVELS * 4))) return rc;
  if ((rc = h->d_status.reserve(16))) return rc;
  {
    std::vector<FastCell> cells((size_t)D.totalCells);
    for (int l = 0; l < h->nlevels; l++) {
      const LevelDev& L = D.lv[l];
      for (int ci = 0; ci < L.nRows; ci++)
        for (int cj = 0; cj < L.nCols; cj++) {
          FastCell& c = cells[(size_t)L.cellBase + (size_t)ci * L.nCols + cj];
          c.l = l;
          c.iniY = MIN_BORDER + ci * L.hCell;
          c.iniX = MIN_BORDER + cj * L.wCell;
          const int maxY = std::min(c.iniY + L.hCell + 6, L.maxBorderY), maxX = std::min(c.iniX + L.wCell + 6, L.maxBorderX);
          const bool skip = (c.iniY >= L.maxBorderY - 3) || (c.iniX >= L.maxBorderX - 6);
          const int rw = maxX - c.iniX, rh = maxY - c.iniY;
          c.rwrh = (skip || rw < 7 || rh < 7) ? 0 : (rw | (rh << 16));
        }
    }
    if ((rc = h->d_cellTab.reserve(cells.size() * sizeof(FastCell)))) return rc;
    GFS_CUDA(cudaMemcpy(h->d_cellTab.p, cells.data(), cells.size() * sizeof(FastCell), cudaMemcpyHostToDevice));
  }
  if ((rc = h->d_tabs.reserve(std::max<size_t>(tabs.size(), 1) * sizeof(AreaEntry)))) return rc;
  if (!tabs.empty()) GFS_CUDA(cudaMemcpy(h->d_tabs.p, tabs.data(), tabs.size() * sizeof(AreaEntry), cudaMemcpyHostToDevice));
  if (pyr3Rows) {
    if ((rc = h->d_tabs3.reserve(tabs3.size() * sizeof(Area3)))) return rc;
    GFS_CUDA(cudaMemcpy(h->d_tabs3.p, tabs3.data(), tabs3.size() * sizeof(Area3), cudaMemcpyHostToDevice));
  }
  GFS_CUDA(cudaMemset(h->d_status.p, 0, 16));
  if ((rc = h->d_pattern.reserve(1024))) return rc;
  GFS_CUDA(cudaMemcpy(h->d_pattern.p, GFS_ORB_PATTERN, 1024, cudaMemcpyHostToDevice));
  h->fastSmem = (size_t)D.offKept + (size_t)4 * cellCap + 16 + 128;   // + the slack k_fast_cells2 uses to align its window to 128 bytes
  const size_t perWarp = align_up(sizeof(OctState) + (size_t)nodeCap * (sizeof(OctNode) + 4 * sizeof(int)), 16);
  h->octSmem = perWarp * OCT_WARPS;
  if (h->octSmem > 200 * 1024) {
    set_error("nfeatures too large: quadtree needs %zu B shared memory", h->octSmem);
    return GFS_ERR_CAPACITY;
  }
  // the attribute is per function (shared by every handle): only ever raise it
  static size_t octMax = 48 * 1024, fastMax = 48 * 1024, pyrMax = 48 * 1024, pyr3Max = 48 * 1024;
  if (pyr3Smem > pyr3Max) {
    GFS_CUDA(cudaFuncSetAttribute(k_pyr_level3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pyr3Smem));
    pyr3Max = pyr3Smem;
  }
  if (h->pyrSmem > pyrMax) {
    GFS_CUDA(cudaFuncSetAttribute(k_pyr_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->pyrSmem));
    pyrMax = h->pyrSmem;
  }
  if (h->octSmem > octMax) {
    GFS_CUDA(cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->octSmem));
    octMax = h->octSmem;
  }
  if (h->fastSmem > fastMax) {
    GFS_CUDA(cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fastSmem));
    GFS_CUDA(cudaFuncSetAttribute(k_fast_cells2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fastSmem));
    fastMax = h->fastSmem;
  }
  h->pyr3Rows = pyr3Rows;
  h->pyr3PitchF = pyr3PitchF;
  h->pyr3Smem = pyr3Smem;
  h->dev = D;
  h->geomW = w;
  h->geomH = ht;
  return GFS_OK;
}

extern "C" {

int gfs_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, int max_w,
                   int max_h, int max_batch, GfsOrb** out) {
  GFS_REQUIRE(out, GFS_ERR_INVALID, "out is null");
  *out = nullptr;
  GFS_REQUIRE(nfeatures > 0 && nlevels >= 1 && nlevels <= MAX_LEVELS, GFS_ERR_INVALID, "bad nfeatures/nlevels");
  GFS_REQUIRE(scale_factor > 1.0f && scale_factor < 3.0f, GFS_ERR_INVALID, "scale_factor must be in (1,3)");
  GFS_REQUIRE(ini_th_fast >= 0 && ini_th_fast <= 255 && min_th_fast >= 0 && min_th_fast <= 255, GFS_ERR_INVALID,
              "FAST thresholds must be in [0,255]");
  GFS_REQUIRE(max_w > 0 && max_h > 0 && max_batch > 0, GFS_ERR_INVALID, "bad workspace size");
  int rc = gfs_device_check();
  if (rc) return rc;
  GfsOrb* h = new GfsOrb();
  h->nfeatures = nfeatures;
  h->nlevels = nlevels;
  h->iniTh = ini_th_fast;
  h->minTh = min_th_fast;
  h->scaleFactor = scale_factor;  // float -> double member, as the reference (ORBextractor.h:103)
  h->maxW = max_w;
  h->maxH = max_h;
  h->maxBatch = max_batch;
  // scale tables and per-level budgets, ctor :428-458
  h->sf[0] = 1.0f;
  for (int i = 1; i < nlevels; i++) h->sf[i] = (float)(h->sf[i - 1] * h->scaleFactor);
  for (int i = 0; i < nlevels; i++) h->isf[i] = 1.0f / h->sf[i];
  const float factor = (float)(1.0f / h->scaleFactor);
  float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
  int sum = 0;
  for (int l = 0; l < nlevels - 1; l++) {
    h->nPerLevel[l] = cv_round_f(nDesired);
    sum += h->nPerLevel[l];
    nDesired *= factor;
  }
  h->nPerLevel[nlevels - 1] = std::max(nfeatures - sum, 0);
  int total = 0;
  for (int l = 0; l < nlevels; l++) total += std::max(h->nPerLevel[l] + 3, 16);
  memset(&h->dev, 0, sizeof(h->dev));
  h->dev.kpStride = (int)align_up((size_t)total, 32);
  *out = h;
  return GFS_OK;
}

int gfs_orb_destroy(GfsOrb* h) {
  if (!h) return GFS_OK;
  DevBuf* d[] = {&h->d_pyr, &h->d_blur, &h->d_cellKeys, &h->d_cellCount, &h->d_keysA, &h->d_keysB, &h->d_sel,
                 &h->d_selCount, &h->d_tabs, &h->d_tabs3, &h->d_cellTab, &h->d_status, &h->d_in, &h->d_okp, &h->d_odesc, &h->d_on, &h->d_omono,
                 &h->d_tkp, &h->d_tdesc, &h->d_pattern};
  for (DevBuf* b : d) b->release();
  h->h_in.release(); h->h_okp.release(); h->h_odesc.release(); h->h_on.release();
  for (int i = 0; i < 7; i++)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->evFork) cudaEventDestroy(h->evFork);
  if (h->evJoin) cudaEventDestroy(h->evJoin);
  if (h->auxStream) cudaStreamDestroy(h->auxStream);
  delete h;
  return GFS_OK;
}

int gfs_orb_max_keypoints(const GfsOrb* h) { return h ? h->dev.kpStride : GFS_ERR_INVALID; }

int gfs_orb_tables(const GfsOrb* h, float* scale_factors, int* features_per_level) {
  GFS_REQUIRE(h, GFS_ERR_INVALID, "null handle");
  for (int i = 0; i < h->nlevels; i++) {
    if (scale_factors) scale_factors[i] = h->sf[i];
    if (features_per_level) features_per_level[i] = h->nPerLevel[i];
  }
  return GFS_OK;
}

int gfs_orb_level_size(const GfsOrb* h, int w, int h_img, int level, int* lw, int* lh) {
  GFS_REQUIRE(h && level >= 0 && level < h->nlevels, GFS_ERR_INVALID, "bad handle/level");
  if (lw) *lw = cv_round_f((float)w * h->isf[level]);
  if (lh) *lh = cv_round_f((float)h_img * h->isf[level]);
  return GFS_OK;
}

int gfs_orb_set_profiling(GfsOrb* h, int enable) {
  GFS_REQUIRE(h, GFS_ERR_INVALID, "null handle");
  if (enable && !h->ev[0])
    for (int i = 0; i < 7; i++) GFS_CUDA(cudaEventCreate(&h->ev[i]));
  h->profiling = enable != 0;
  return GFS_OK;
}

int gfs_orb_get_profile(GfsOrb* h, float* ms6) {
  GFS_REQUIRE(h && ms6 && h->profiling && h->ev[0], GFS_ERR_INVALID, "profiling not enabled");
  GFS_CUDA(cudaEventSynchronize(h->ev[6]));
  for (int i = 0; i < 6; i++) GFS_CUDA(cudaEventElapsedTime(&ms6[i], h->ev[i], h->ev[i + 1]));
  return GFS_OK;
}

int gfs_orb_launches_per_call(const GfsOrb* h, int lap0, int lap1) {
  if (!h) return GFS_ERR_INVALID;
  // pyramid (nlevels-1) + fast + octree + orient/desc with the fused blur (+ pack)
  return (h->nlevels - 1) + 1 + 1 + 1 + ((lap0 != 0 || lap1 != 0) ? 1 : 0);
}

// Frames of the chunk use scratch slots slot0 .. slot0+batch-1 of the handle, so chunks with disjoint
// slot ranges may run concurrently on different streams.
static int orb_run_chunk(GfsOrb* h, cudaStream_t st, int slot0, const uint8_t* d_imgs, int batch, int w, int ht, int pitch,
                         size_t img_stride, int lap0, int lap1, GfsKeyPoint* d_kp, uint8_t* d_desc, int* d_n,
                         int* d_mono) {
  const OrbDev& D = h->dev;
  uint8_t* pyr = (uint8_t*)h->d_pyr.p + (size_t)slot0 * D.pyrStride;
  uint32_t* cellKeys = (uint32_t*)h->d_cellKeys.p + (size_t)slot0 * D.totalCells * D.cellCap;
  int* cellCount = (int*)h->d_cellCount.p + (size_t)slot0 * D.totalCells;
  uint32_t* keysA = (uint32_t*)h->d_keysA.p + (size_t)slot0 * D.keysPerFrame;
  uint32_t* keysB = (uint32_t*)h->d_keysB.p + (size_t)slot0 * D.keysPerFrame;
  uint32_t* selKeys = (uint32_t*)h->d_sel.p + (size_t)slot0 * D.selPerFrame;
  int* selCount = (int*)h->d_selCount.p + (size_t)slot0 * D.nlevels;
  auto mark = [&](int i) { if (h->profiling) cudaEventRecord(h->ev[i], st); };
  mark(0);
  for (int l = 1; l < h->nlevels; l++) {
    const LevelDev& L = D.lv[l];
    dim3 blk(256), grd(div_up(L.w, PYR_TW), div_up(L.h, PYR_TH), batch);
    const uint8_t* src = (l == 1) ? d_imgs : pyr + D.lv[l - 1].off;
    const long long ss = (l == 1) ? (long long)img_stride : D.pyrStride;
    const int sp = (l == 1) ? pitch : D.lv[l - 1].pitch;
    if (h->pyr3Rows) {
      const int word_ok = (((uintptr_t)src | (uintptr_t)ss | (uintptr_t)sp) & 3) == 0;
      k_pyr_level3<<<grd, blk, h->pyr3Smem, st>>>(D, l, src, ss, sp, word_ok, pyr, (const Area3*)h->d_tabs3.p, h->pyr3Rows,
                                                  h->pyr3PitchF);
    } else {
      k_pyr_level<<<grd, blk, h->pyrSmem, st>>>(D, l, src, ss, sp, pyr, (const AreaEntry*)h->d_tabs.p);
    }
  }
  mark(1);
  if (h->fastV1)
    k_fast_cells<<<dim3(D.totalCells, batch), FAST_THREADS, h->fastSmem, st>>>(
        D, d_imgs, (long long)img_stride, pitch, pyr, cellKeys, cellCount);
  else {
    FastTmaps tm;
    memset(&tm, 0, sizeof(tm));
    for (int l = 0; l < h->nlevels; l++) {
      const LevelDev& L = D.lv[l];
      const bool ok = l == 0 ? encode_level_map(&tm.m[0], d_imgs, L.w, L.h, batch, pitch, (long long)img_stride, D.roiPitch, D.roiRows)
                             : encode_level_map(&tm.m[l], pyr + L.off, L.w, L.h, batch, L.pitch, D.pyrStride, D.roiPitch, D.roiRows);
      if (ok) tm.mask |= 1u << l;
    }
    h->lastTmapMask = tm.mask;
    k_fast_cells2<<<dim3(D.totalCells, batch), FAST_THREADS, h->fastSmem, st>>>(
        D, tm, (const FastCell*)h->d_cellTab.p, d_imgs, (long long)img_stride, pitch, pyr, cellKeys, cellCount);
  }
  mark(2);
  k_octree<<<div_up(batch * h->nlevels, OCT_WARPS), OCT_WARPS * 32, h->octSmem, st>>>(
      D, batch, cellKeys, cellCount, keysA, keysB, selKeys, selCount, (int*)h->d_status.p);
  mark(3);
  mark(4);  // (the 7x7 blur is fused into k_orient_desc; k_blur7 only serves gfs_orb_get_level)
  const bool lapping = (lap0 != 0 || lap1 != 0);
  GfsKeyPoint* kpDst = d_kp;
  uint8_t* descDst = d_desc;
  if (lapping) {
    int rc;
    if ((rc = h->d_tkp.reserve((size_t)h->maxBatch * D.kpStride * sizeof(GfsKeyPoint)))) return rc;
    if ((rc = h->d_tdesc.reserve((size_t)h->maxBatch * D.kpStride * 32))) return rc;
    kpDst = (GfsKeyPoint*)h->d_tkp.p + (size_t)slot0 * D.kpStride;
    descDst = (uint8_t*)h->d_tdesc.p + (size_t)slot0 * D.kpStride * 32;
  }
  k_orient_desc<<<dim3(div_up(D.selPerFrame, OD_WARPS), batch), OD_WARPS * 32, 0, st>>>(
      D, d_imgs, (long long)img_stride, pitch, pyr, (const uint4*)h->d_pattern.p, selKeys, selCount, kpDst, descDst, d_n,
      d_mono);
  mark(5);
  if (lapping)
    k_pack_lapping<<<batch, 1024, 0, st>>>(D.kpStride, lap0, lap1, kpDst, descDst, d_kp, d_desc, d_n, d_mono);
  mark(6);
  GFS_CUDA(cudaGetLastError());
  if (slot0 == 0) {
    h->lastImgs = d_imgs;
    h->lastStride = (long long)img_stride;
    h->lastPitch = pitch;
  }
  h->lastBatch = slot0 + batch;
  return GFS_OK;
}

// internal entry used by the front-end pipeline (frontend.cu): one chunk on the caller's stream, scratch
// slots slot0 .. slot0+nb-1
extern "C++" {
namespace gfs {
int orb_extract_slots(GfsOrb* h, cudaStream_t st, int slot0, const uint8_t* d_imgs, int nb, int w, int h_img, int pitch,
                      size_t img_stride, GfsKeyPoint* d_kp, uint8_t* d_desc, int* d_n, int* d_mono) {
  GFS_REQUIRE(h && slot0 >= 0 && nb > 0 && slot0 + nb <= h->maxBatch, GFS_ERR_CAPACITY, "slot range exceeds max_batch");
  GFS_REQUIRE(w <= h->maxW && h_img <= h->maxH, GFS_ERR_CAPACITY, "image larger than the handle's max_w/max_h");
  int rc = orb_set_geometry(h, w, h_img);
  if (rc) return rc;
  return orb_run_chunk(h, st, slot0, d_imgs, nb, w, h_img, pitch, img_stride, 0, 0, d_kp, d_desc, d_n, d_mono);
}
}  // namespace gfs
}  // extern "C++"

int gfs_orb_extract_batch_device(GfsOrb* h, void* stream, const uint8_t* d_imgs, int batch, int w, int h_img,
                                 int pitch, size_t img_stride, int lap0, int lap1, GfsKeyPoint* d_out_kp,
                                 uint8_t* d_out_desc, int* d_out_n, int* d_out_mono) {
  GFS_REQUIRE(h, GFS_ERR_INVALID, "null handle");
  GFS_REQUIRE(d_imgs && w > 0 && h_img > 0, GFS_ERR_EMPTY, "empty image");
  GFS_REQUIRE(batch > 0 && pitch >= w, GFS_ERR_INVALID, "bad batch/pitch");
  GFS_REQUIRE(d_out_kp && d_out_desc && d_out_n && d_out_mono, GFS_ERR_INVALID, "null output");
  GFS_REQUIRE(w <= h->maxW && h_img <= h->maxH, GFS_ERR_CAPACITY, "image larger than the handle's max_w/max_h");
  int rc = orb_set_geometry(h, w, h_img);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int ks = h->dev.kpStride;
  if (!h->profiling && batch >= 128 && batch <= h->maxBatch) {
    // Two halves on two streams: the latency-bound quadtree kernel of one half (one warp per frame and
    // level) overlaps the throughput-bound kernels of the other.
    if (!h->auxStream) {
      GFS_CUDA(cudaStreamCreateWithFlags(&h->auxStream, cudaStreamNonBlocking));
      GFS_CUDA(cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming));
      GFS_CUDA(cudaEventCreateWithFlags(&h->evJoin, cudaEventDisableTiming));
    }
    const int n0 = batch / 2, n1 = batch - n0;
    GFS_CUDA(cudaEventRecord(h->evFork, st));
    GFS_CUDA(cudaStreamWaitEvent(h->auxStream, h->evFork, 0));
    rc = orb_run_chunk(h, st, 0, d_imgs, n0, w, h_img, pitch, img_stride, lap0, lap1, d_out_kp, d_out_desc, d_out_n,
                       d_out_mono);
    if (rc) return rc;
    rc = orb_run_chunk(h, h->auxStream, n0, d_imgs + (size_t)n0 * img_stride, n1, w, h_img, pitch, img_stride, lap0, lap1,
                       d_out_kp + (size_t)n0 * ks, d_out_desc + (size_t)n0 * ks * 32, d_out_n + n0, d_out_mono + n0);
    if (rc) return rc;
    GFS_CUDA(cudaEventRecord(h->evJoin, h->auxStream));
    GFS_CUDA(cudaStreamWaitEvent(st, h->evJoin, 0));
    h->lastImgs = d_imgs;
    return GFS_OK;
  }
  for (int b0 = 0; b0 < batch; b0 += h->maxBatch) {
    const int nb = std::min(h->maxBatch, batch - b0);
    rc = orb_run_chunk(h, st, 0, d_imgs + (size_t)b0 * img_stride, nb, w, h_img, pitch, img_stride, lap0, lap1,
                       d_out_kp + (size_t)b0 * ks, d_out_desc + (size_t)b0 * ks * 32, d_out_n + b0, d_out_mono + b0);
    if (rc) return rc;
  }
  return GFS_OK;
}

int gfs_orb_extract_batch(GfsOrb* h, void* stream, const uint8_t* imgs, int batch, int w, int h_img, int pitch,
                          size_t img_stride, int lap0, int lap1, GfsKeyPoint* out_kp, uint8_t* out_desc, int* out_n,
                          int* out_mono) {
  GFS_REQUIRE(h, GFS_ERR_INVALID, "null handle");
  GFS_REQUIRE(imgs && w > 0 && h_img > 0, GFS_ERR_EMPTY, "empty image");
  GFS_REQUIRE(batch > 0 && pitch >= w, GFS_ERR_INVALID, "bad batch/pitch");
  GFS_REQUIRE(out_kp && out_desc && out_n && out_mono, GFS_ERR_INVALID, "null output");
  cudaStream_t st = (cudaStream_t)stream;
  const int ks = h->dev.kpStride;
  const int cb = std::min(batch, h->maxBatch);
  const size_t dpitch = align_up((size_t)w, 16), dstride = dpitch * h_img;
  int rc;
  if ((rc = h->d_in.reserve(cb * dstride))) return rc;
  if ((rc = h->d_okp.reserve((size_t)cb * ks * sizeof(GfsKeyPoint)))) return rc;
  if ((rc = h->d_odesc.reserve((size_t)cb * ks * 32))) return rc;
  if ((rc = h->d_on.reserve((size_t)cb * 2 * sizeof(int)))) return rc;
  const bool pin_in = is_pinned_host(imgs);
  const bool pin_out = is_pinned_host(out_kp) && is_pinned_host(out_desc) && is_pinned_host(out_n) && is_pinned_host(out_mono);
  if (!pin_in && (rc = h->h_in.reserve(cb * (size_t)w * h_img))) return rc;
  if (!pin_out) {
    if ((rc = h->h_okp.reserve((size_t)cb * ks * sizeof(GfsKeyPoint)))) return rc;
    if ((rc = h->h_odesc.reserve((size_t)cb * ks * 32))) return rc;
    if ((rc = h->h_on.reserve((size_t)cb * 2 * sizeof(int)))) return rc;
  }
  for (int b0 = 0; b0 < batch; b0 += cb) {
    const int nb = std::min(cb, batch - b0);
    const uint8_t* src = imgs + (size_t)b0 * img_stride;
    if (pin_in) {
      if (img_stride == (size_t)pitch * h_img) {
        GFS_CUDA(cudaMemcpy2DAsync(h->d_in.p, dpitch, src, pitch, w, (size_t)h_img * nb, cudaMemcpyHostToDevice, st));
      } else {
        for (int i = 0; i < nb; i++)
          GFS_CUDA(cudaMemcpy2DAsync((uint8_t*)h->d_in.p + i * dstride, dpitch, src + i * img_stride, pitch, w, h_img,
                                     cudaMemcpyHostToDevice, st));
      }
    } else {
      // a previous chunk's H2D from the staging buffer must have completed before it is reused
      GFS_CUDA(cudaStreamSynchronize(st));
      uint8_t* stg = (uint8_t*)h->h_in.p;
      for (int i = 0; i < nb; i++)
        for (int y = 0; y < h_img; y++) memcpy(stg + ((size_t)i * h_img + y) * w, src + i * img_stride + (size_t)y * pitch, w);
      GFS_CUDA(cudaMemcpy2DAsync(h->d_in.p, dpitch, stg, w, w, (size_t)h_img * nb, cudaMemcpyHostToDevice, st));
    }
    int* d_n = (int*)h->d_on.p;
    int* d_mono = d_n + cb;
    rc = gfs_orb_extract_batch_device(h, stream, (const uint8_t*)h->d_in.p, nb, w, h_img, (int)dpitch, dstride, lap0,
                                      lap1, (GfsKeyPoint*)h->d_okp.p, (uint8_t*)h->d_odesc.p, d_n, d_mono);
    if (rc) return rc;
    GfsKeyPoint* okp = pin_out ? out_kp + (size_t)b0 * ks : (GfsKeyPoint*)h->h_okp.p;
    uint8_t* odesc = pin_out ? out_desc + (size_t)b0 * ks * 32 : (uint8_t*)h->h_odesc.p;
    int* on = pin_out ? out_n + b0 : (int*)h->h_on.p;
    int* omono = pin_out ? out_mono + b0 : (int*)h->h_on.p + cb;
    GFS_CUDA(cudaMemcpyAsync(okp, h->d_okp.p, (size_t)nb * ks * sizeof(GfsKeyPoint), cudaMemcpyDeviceToHost, st));
    GFS_CUDA(cudaMemcpyAsync(odesc, h->d_odesc.p, (size_t)nb * ks * 32, cudaMemcpyDeviceToHost, st));
    GFS_CUDA(cudaMemcpyAsync(on, d_n, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost, st));
    GFS_CUDA(cudaMemcpyAsync(omono, d_mono, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost, st));
    GFS_CUDA(cudaStreamSynchronize(st));
    if (!pin_out) {
      memcpy(out_kp + (size_t)b0 * ks, okp, (size_t)nb * ks * sizeof(GfsKeyPoint));
      memcpy(out_desc + (size_t)b0 * ks * 32, odesc, (size_t)nb * ks * 32);
      memcpy(out_n + b0, on, (size_t)nb * sizeof(int));
      memcpy(out_mono + b0, omono, (size_t)nb * sizeof(int));
    }
  }
  int status = 0;
  GFS_CUDA(cudaMemcpy(&status, h->d_status.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (status & 1) {
    set_error("quadtree produced more keypoints than the per-level slot capacity");
    return GFS_ERR_CAPACITY;
  }
  return GFS_OK;
}

int gfs_orb_extract(GfsOrb* h, void* stream, const uint8_t* img, int w, int h_img, int pitch, int lap0, int lap1,
                    GfsKeyPoint* out_kp, uint8_t* out_desc, int* out_n, int* out_mono) {
  if (!img || w <= 0 || h_img <= 0) {
    if (out_n) *out_n = 0;
    if (out_mono) *out_mono = -1;
    set_error("empty image");
    return GFS_ERR_EMPTY;
  }
  return gfs_orb_extract_batch(h, stream, img, 1, w, h_img, pitch, (size_t)pitch * h_img, lap0, lap1, out_kp, out_desc,
                               out_n, out_mono);
}

int gfs_orb_get_level(GfsOrb* h, void* stream, int frame, int level, int blurred, uint8_t* out) {
  GFS_REQUIRE(h && out && h->geomW > 0, GFS_ERR_INVALID, "no batch processed yet");
  GFS_REQUIRE(frame >= 0 && frame < h->lastBatch && level >= 0 && level < h->nlevels, GFS_ERR_INVALID, "bad frame/level");
  cudaStream_t st = (cudaStream_t)stream;
  const LevelDev& L = h->dev.lv[level];
  const uint8_t* src;
  size_t sp;
  if (blurred) {
    // the blurred working copy (ORBextractor.cc:1188-1189) is not kept by the extraction path any more:
    // materialise this level on demand with the stand-alone blur kernel
    int rc = h->d_blur.reserve((size_t)h->maxBatch * h->dev.pyrStride);
    if (rc) return rc;
    k_blur7<<<dim3(div_up(L.w, BL_TW), div_up(L.h, BL_TH), h->lastBatch), 256, 0, st>>>(
        h->dev, level, h->lastImgs, h->lastStride, h->lastPitch, (const uint8_t*)h->d_pyr.p, (uint8_t*)h->d_blur.p);
    GFS_CUDA(cudaGetLastError());
    src = (const uint8_t*)h->d_blur.p + (size_t)frame * h->dev.pyrStride + L.off;
    sp = L.pitch;
  } else if (level == 0) {
    src = h->lastImgs + (size_t)frame * h->lastStride;
    sp = h->lastPitch;
  } else {
    src = (const uint8_t*)h->d_pyr.p + (size_t)frame * h->dev.pyrStride + L.off;
    sp = L.pitch;
  }
  GFS_CUDA(cudaMemcpy2DAsync(out, L.w, src, sp, L.w, L.h, cudaMemcpyDeviceToHost, st));
  GFS_CUDA(cudaStreamSynchronize(st));
  return GFS_OK;
}

int gfs_orb_get_candidates(GfsOrb* h, void* stream, int frame, int level, float* out_xyr, int cap, int* n) {
  GFS_REQUIRE(h && n && h->geomW > 0, GFS_ERR_INVALID, "no batch processed yet");
  GFS_REQUIRE(frame >= 0 && frame < h->lastBatch && level >= 0 && level < h->nlevels, GFS_ERR_INVALID, "bad frame/level");
  cudaStream_t st = (cudaStream_t)stream;
  const OrbDev& D = h->dev;
  const LevelDev& L = D.lv[level];
  const int nc = L.nCols * L.nRows;
  std::vector<int> cnt(nc);
  std::vector<uint32_t> keys((size_t)nc * D.cellCap);
  GFS_CUDA(cudaStreamSynchronize(st));
  GFS_CUDA(cudaMemcpy(cnt.data(), (int*)h->d_cellCount.p + (size_t)frame * D.totalCells + L.cellBase, nc * sizeof(int),
                      cudaMemcpyDeviceToHost));
  GFS_CUDA(cudaMemcpy(keys.data(), (uint32_t*)h->d_cellKeys.p + ((size_t)frame * D.totalCells + L.cellBase) * D.cellCap,
                      keys.size() * 4, cudaMemcpyDeviceToHost));
  int k = 0;
  for (int c = 0; c < nc; c++)
    for (int i = 0; i < cnt[c]; i++, k++) {
      if (k < cap && out_xyr) {
        const uint32_t key = keys[(size_t)c * D.cellCap + i];
        out_xyr[3 * k] = (float)(key & 0xFFF);
        out_xyr[3 * k + 1] = (float)((key >> 12) & 0xFFF);
        out_xyr[3 * k + 2] = (float)(key >> 24);
      }
    }
  *n = k;
  return GFS_OK;
}

// Host-only debug hook: the libstdc++-introsort restatement used by k_octree, exposed so the CPU
// test-suite can compare it against std::sort (tests/test_host_logic.py).  key: (first, ulx).
int gfs_debug_gcc_sort(int* first, int* second, const int* ulx_by_second, int n, int* heapsorted) {
  struct L {
    const int* u;
    GFS_HD bool operator()(int af, int as, int bf, int bs) const {
      if (af < bf) return true;
      if (af > bf) return false;
      return u[as] < u[bs];
    }
  };
  int hs = 0;
  GccSort<L> s{{first, second}, L{ulx_by_second}};
  s.sort(n, &hs);
  if (heapsorted) *heapsorted = hs;
  return GFS_OK;
}
}
