// Optical-flow front end on sm_100a, batched over frames / frame pairs.
//
// Replaces
//   cv::buildOpticalFlowPyramid(image, mImGray, winSize, 3)                 reference src/Frame.cc:370-373
//   ORBmatcher::fbKltTracking / Tracking::fbKltTracking                     src/ORBmatcher.cc:2186-2293, src/Tracking.cc:3262-3360
//     (two cv::calcOpticalFlowPyrLK calls with OPTFLOW_USE_INITIAL_FLOW | OPTFLOW_LK_GET_MIN_EIGENVALS, 30 iterations,
//      eps 0.01: forward over nbpyrlvl levels, status / min-eigenvalue / inBorder filter, backward at level 0,
//      forward-backward distance check)
// OpenCV's arithmetic restated (modules/imgproc/src/pyramids.cpp pyrDown 8U; modules/video/src/lkpyramid.cpp
// calcScharrDeriv, LKTrackerInvoker): everything up to the 2x2 solve is integer / fixed point and is reproduced
// exactly; the sums of integer products OpenCV accumulates in float SIMD lanes are accumulated exactly in int64
// (order-free, so the result does not depend on the lane layout) and narrowed once.
//
//   k_clahe_lut/apply  cv::CLAHE (Frame.cc:366-368): per-tile histogram + clip + LUT, then the bilinear LUT blend
//   k_klt_pyr_down   CTA per 32x8 destination tile: the 68x20 source window staged in shared memory, separable
//                    [1 4 6 4 1], (sum + 128) >> 8
//   k_klt_scharr     CTA per 128x32 tile: int16 (dI/dx, dI/dy), reflect-101 inside the image, 16-byte stores
//   k_klt_track      one warp per point, all pyramid levels and both passes in ONE launch: the 35x35 template patch
//                    (int16 intensities with 5 fractional bits + int16 derivatives, interpolated in place from the
//                    staged 36x36 derivative window) lives in shared memory, the 36x36 window of the other image is
//                    staged per iteration (aligned 32-bit words inside the image, reflect-101 bytes at the border),
//                    the five sums are warp-reduced int64
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>   // CUtensorMap and its enums only (cuTensorMapEncodeTiled is fetched through the runtime)
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace gfs {
namespace klt {

static const int MAX_LEVELS = 8;
static const int TRACK_WARPS = 4;

struct PyrGeom {
  int levels;                 // number of levels above level 0
  int w[MAX_LEVELS + 1], h[MAX_LEVELS + 1];
  int off[MAX_LEVELS + 1];    // pixel offset of the level inside the packed image / derivative arrays
  int npix;                   // total pixels over all levels
  int tilesX[MAX_LEVELS + 1], tileBase[MAX_LEVELS + 2];  // 128x32 tiles of k_klt_scharr: per-level grid width and first block
  size_t imgBytes;            // npix rounded up to 16: the derivatives start there
  size_t derEnd;              // imgBytes + 4 * npix rounded up to 16: the padded level copies start there
  // every level once more with a reflect-101 border of KLT_PAD_X / KLT_PAD_Y pixels (what cv::buildOpticalFlowPyramid's
  // copyMakeBorder produces), row pitch a multiple of 16 bytes: a tracking window that leaves the image is then an ordinary
  // box of the padded copy, and EVERY image window can come through the tensor-map path
  int padW[MAX_LEVELS + 1], padH[MAX_LEVELS + 1];
  size_t padOff[MAX_LEVELS + 1];   // byte offset inside the frame
  size_t frameBytes;
};
static const int KLT_PAD_X = 48, KLT_PAD_Y = 40;   // >= the largest window the padded path serves (win + 1 <= 36 ... 40)

static PyrGeom make_geom(int w, int h, int levels) {
  PyrGeom g;
  memset(&g, 0, sizeof(g));
  g.levels = levels;
  int off = 0;
  for (int l = 0; l <= levels; l++) {
    g.w[l] = w; g.h[l] = h; g.off[l] = off;
    off += w * h;
    w = (w + 1) / 2; h = (h + 1) / 2;
  }
  g.npix = off;
  int tb = 0;
  for (int l = 0; l <= levels; l++) {
    g.tilesX[l] = (g.w[l] + 127) / 128;
    g.tileBase[l] = tb;
    tb += g.tilesX[l] * ((g.h[l] + 31) / 32);
  }
  g.tileBase[levels + 1] = tb;
  g.imgBytes = align_up((size_t)off, 16);
  g.derEnd = align_up(g.imgBytes + (size_t)4 * off, 16);
  size_t po = g.derEnd;
  for (int l = 0; l <= levels; l++) {
    g.padW[l] = (int)align_up((size_t)g.w[l] + 2 * KLT_PAD_X, 16);
    g.padH[l] = g.h[l] + 2 * KLT_PAD_Y;
    g.padOff[l] = po;
    po += (size_t)g.padW[l] * g.padH[l];
  }
  g.frameBytes = align_up(po, 16);
  return g;
}

// Highest level cv::buildOpticalFlowPyramid(img, pyr, Size(win, win), levels) actually builds: it stops as soon as the
// next level would be <= the window in either dimension (its return value), and calcOpticalFlowPyrLK clamps maxLevel to
// what the pyramids hold.  VGA with the 35-px window keeps all 3 levels; QVGA keeps 2.
static int effective_levels(int w, int h, int levels, int win) {
  int l = 0;
  while (l < levels) {
    const int nw = (w + 1) / 2, nh = (h + 1) / 2;
    if (nw <= win || nh <= win) break;
    w = nw; h = nh; l++;
  }
  return l;
}

__device__ __forceinline__ int reflect101(int p, int len) {  // cv::borderInterpolate(BORDER_REFLECT_101)
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
  return p;
}

// ---- cv::CLAHE::apply, CV_8UC1 (modules/imgproc/src/clahe.cpp), as Frame::Frame applies it (reference
// src/Frame.cc:366-368: createCLAHE(3.0, Size(8, 8)), in place).
struct ClaheGeom { int w, h, tilesX, tilesY, tw, th, clipLimit; float lutScale; };
// CTA per (tile, frame): histogram in shared memory, clip + redistribute exactly as CLAHE_CalcLut_Body, prefix sum -> LUT
__global__ void __launch_bounds__(256) k_clahe_lut(const uint8_t* __restrict__ src, int pitch, size_t imgStride, ClaheGeom G,
                                                   uint8_t* __restrict__ lut) {
  __shared__ int s_hist[256];
  __shared__ int s_scan[256];
  __shared__ int s_clipped;
  const int tid = threadIdx.x, k = blockIdx.x, ty = k / G.tilesX, tx = k % G.tilesX;
  const uint8_t* img = src + (size_t)blockIdx.y * imgStride;
  s_hist[tid] = 0;
  if (tid == 0) s_clipped = 0;
  __syncthreads();
  const bool interior = (tx + 1) * G.tw <= G.w && (ty + 1) * G.th <= G.h;
  for (int i = tid; i < G.tw * G.th; i += 256) {
    const int y = i / G.tw, x = i - y * G.tw;
    const int X = tx * G.tw + x, Y = ty * G.th + y;
    const uint8_t v = interior ? img[(size_t)Y * pitch + X] : img[(size_t)reflect101(Y, G.h) * pitch + reflect101(X, G.w)];
    atomicAdd(&s_hist[v], 1);
  }
  __syncthreads();
  int hv = s_hist[tid];
  if (G.clipLimit > 0) {
    if (hv > G.clipLimit) { atomicAdd(&s_clipped, hv - G.clipLimit); hv = G.clipLimit; }
    __syncthreads();
    const int clipped = s_clipped;
    const int redistBatch = clipped / 256;
    const int residual = clipped - redistBatch * 256;
    hv += redistBatch;
    if (residual != 0) {
      const int step = max(256 / residual, 1);
      // for (i = 0; i < 256 && residual > 0; i += step, residual--) hist[i]++
      if (tid % step == 0 && tid / step < residual) hv++;
    }
  }
  // inclusive prefix sum over the 256 bins (Hillis-Steele in shared memory: integers, order-free)
  s_scan[tid] = hv;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    const int t = tid >= o ? s_scan[tid - o] : 0;
    __syncthreads();
    s_scan[tid] += t;
    __syncthreads();
  }
  const int v = __float2int_rn((float)s_scan[tid] * G.lutScale);  // saturate_cast<uchar>(sum * lutScale)
  lut[((size_t)blockIdx.y * G.tilesX * G.tilesY + k) * 256 + tid] = (uint8_t)min(max(v, 0), 255);
}
// thread per pixel: bilinear blend of the four surrounding tile LUTs (CLAHE_Interpolation_Body)
__global__ void __launch_bounds__(256) k_clahe_apply(const uint8_t* __restrict__ src, int pitch, size_t imgStride, ClaheGeom G,
                                                     const uint8_t* __restrict__ lut, uint8_t* __restrict__ dst, int dstPitch,
                                                     size_t dstStride) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= G.w || y >= G.h) return;
  const float inv_tw = 1.0f / G.tw, inv_th = 1.0f / G.th;
  const float tyf = y * inv_th - 0.5f, txf = x * inv_tw - 0.5f;
  int ty1 = (int)floorf(tyf), tx1 = (int)floorf(txf);
  int ty2 = ty1 + 1, tx2 = tx1 + 1;
  const float ya = tyf - ty1, ya1 = 1.0f - ya, xa = txf - tx1, xa1 = 1.0f - xa;
  ty1 = max(ty1, 0); ty2 = min(ty2, G.tilesY - 1); tx1 = max(tx1, 0); tx2 = min(tx2, G.tilesX - 1);
  const int v = src[(size_t)blockIdx.z * imgStride + (size_t)y * pitch + x];
  const uint8_t* L = lut + (size_t)blockIdx.z * G.tilesX * G.tilesY * 256;
  const float l11 = L[(ty1 * G.tilesX + tx1) * 256 + v], l12 = L[(ty1 * G.tilesX + tx2) * 256 + v];
  const float l21 = L[(ty2 * G.tilesX + tx1) * 256 + v], l22 = L[(ty2 * G.tilesX + tx2) * 256 + v];
  const float res = (l11 * xa1 + l12 * xa) * ya1 + (l21 * xa1 + l22 * xa) * ya;
  dst[(size_t)blockIdx.z * dstStride + (size_t)y * dstPitch + x] = (uint8_t)min(max(__float2int_rn(res), 0), 255);
}

// ---- level 0 of the pyramid = the image itself (buildOpticalFlowPyramid copies it into its padded block): one launch
// for the whole batch, 16 bytes per thread when pitch, width and the pointers allow it
__global__ void __launch_bounds__(256) k_klt_copy_level0(const uint8_t* __restrict__ imgs, int pitch, size_t imgStride, int w, int h,
                                                         uint8_t* __restrict__ pyr, size_t frameBytes, int vec) {
  const uint8_t* src = imgs + (size_t)blockIdx.y * imgStride;
  uint8_t* dst = pyr + (size_t)blockIdx.y * frameBytes;
  if (vec) {
    const int wq = w >> 4, n = wq * h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const int y = i / wq, x = i - y * wq;
      reinterpret_cast<uint4*>(dst + (size_t)y * w)[x] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)y * pitch) + x);
    }
  } else {
    const int n = w * h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const int y = i / w, x = i - y * w;
      dst[i] = src[(size_t)y * pitch + x];
    }
  }
}

// ---- cv::pyrDown, 8U: dst (dw x dh) from src (sw x sh), one frame per blockIdx.z
static const int PD_TX = 32, PD_TY = 8;
__global__ void __launch_bounds__(PD_TX * PD_TY) k_klt_pyr_down(uint8_t* __restrict__ pyr, size_t frameBytes, int srcOff, int sw, int sh,
                                                                int dstOff, int dw, int dh) {
  __shared__ uint8_t s_src[2 * PD_TY + 3][2 * PD_TX + 4];
  __shared__ int s_row[2 * PD_TY + 3][PD_TX];
  const uint8_t* src = pyr + (size_t)blockIdx.z * frameBytes + srcOff;
  uint8_t* dst = pyr + (size_t)blockIdx.z * frameBytes + dstOff;
  const int x0 = blockIdx.x * PD_TX, y0 = blockIdx.y * PD_TY;
  const int tid = threadIdx.y * PD_TX + threadIdx.x;
  for (int i = tid; i < (2 * PD_TY + 3) * (2 * PD_TX + 4); i += PD_TX * PD_TY) {
    const int r = i / (2 * PD_TX + 4), c = i % (2 * PD_TX + 4);
    s_src[r][c] = src[(size_t)reflect101(2 * y0 - 2 + r, sh) * sw + reflect101(2 * x0 - 2 + c, sw)];
  }
  __syncthreads();
  for (int i = tid; i < (2 * PD_TY + 3) * PD_TX; i += PD_TX * PD_TY) {
    const int r = i / PD_TX, x = i % PD_TX;
    const uint8_t* s = &s_src[r][2 * x];
    s_row[r][x] = s[0] + s[4] + 4 * (s[1] + s[3]) + 6 * s[2];
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x < dw && y < dh) {
    const int r = 2 * threadIdx.y, c = threadIdx.x;
    const int v = s_row[r][c] + s_row[r + 4][c] + 4 * (s_row[r + 1][c] + s_row[r + 3][c]) + 6 * s_row[r + 2][c];
    dst[(size_t)y * dw + x] = (uint8_t)((v + 128) >> 8);
  }
}

// ---- calcScharrDeriv for every level of every frame.  CTA per 128x32 tile of one level: the 130x34 source window goes to
// shared memory (aligned 32-bit loads inside the image, reflect-101 bytes on the rim), every thread produces four
// pixels and writes them as one 16-byte store when the level's geometry allows it.  HBM-bound: 1 byte read + 4 bytes
// written per pixel.
static const int SC_TX = 128, SC_TY = 32;
// the padded copies of all levels: 16 bytes of a row per thread
__global__ void __launch_bounds__(256) k_klt_pad_level(uint8_t* __restrict__ pyr, PyrGeom G) {
  const int l = blockIdx.z;   // one launch for all levels: the grid is sized for level 0, the CTAs beyond a level's size retire
  const int w = G.w[l], h = G.h[l], pw = G.padW[l], ph = G.padH[l];
  const int groups = pw / 16;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups * ph) return;
  const int y = i / groups, gx = (i - y * groups) * 16;
  const uint8_t* src = pyr + (size_t)blockIdx.y * G.frameBytes + G.off[l] + (size_t)reflect101(y - KLT_PAD_Y, h) * w;
  uint8_t* dst = pyr + (size_t)blockIdx.y * G.frameBytes + G.padOff[l] + (size_t)y * pw + gx;
  unsigned v[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    unsigned word = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) word |= (unsigned)src[reflect101(gx + 4 * k + b - KLT_PAD_X, w)] << (8 * b);
    v[k] = word;
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
}

__global__ void __launch_bounds__(256) k_klt_scharr(uint8_t* __restrict__ pyr, PyrGeom G) {
  __shared__ __align__(16) uint8_t s_t[SC_TY + 2][SC_TX + 8];  // column c of the tile at [.][c + 4]: words stay aligned
  int b = blockIdx.x, l = 0;
  while (l < G.levels && b >= G.tileBase[l + 1]) l++;
  b -= G.tileBase[l];
  const int w = G.w[l], h = G.h[l];
  const int x0 = (b % G.tilesX[l]) * SC_TX, y0 = (b / G.tilesX[l]) * SC_TY;
  const uint8_t* img = pyr + (size_t)blockIdx.y * G.frameBytes + G.off[l];
  short2* der = reinterpret_cast<short2*>(pyr + (size_t)blockIdx.y * G.frameBytes + G.imgBytes) + G.off[l];
  const int tid = threadIdx.x;
  const bool wordRows = (w % 4 == 0) && (G.off[l] % 4 == 0) && x0 + SC_TX <= w;
  if (wordRows) {
    for (int i = tid; i < (SC_TY + 2) * (SC_TX / 4); i += 256) {
      const int r = i / (SC_TX / 4), k = i - r * (SC_TX / 4);
      const int Y = reflect101(y0 - 1 + r, h);
      reinterpret_cast<unsigned*>(&s_t[r][4])[k] = __ldg(reinterpret_cast<const unsigned*>(img + (size_t)Y * w + x0) + k);
    }
    if (tid < 2 * (SC_TY + 2)) {  // the two rim columns (68 threads)
      const int r = tid >> 1, right = tid & 1;
      const int Y = reflect101(y0 - 1 + r, h);
      s_t[r][right ? 4 + SC_TX : 3] = img[(size_t)Y * w + reflect101(right ? x0 + SC_TX : x0 - 1, w)];
    }
  } else {
    for (int i = tid; i < (SC_TY + 2) * (SC_TX + 2); i += 256) {
      const int r = i / (SC_TX + 2), c = i - r * (SC_TX + 2);
      s_t[r][3 + c] = img[(size_t)reflect101(y0 - 1 + r, h) * w + reflect101(x0 - 1 + c, w)];
    }
  }
  __syncthreads();
  const int cx = (tid & 31) * 4, x = x0 + cx;
  if (x >= w) return;
  const bool vecStore = wordRows && x + 4 <= w;
  for (int r = tid >> 5; r < SC_TY; r += 8) {
    const int y = y0 + r;
    if (y >= h) break;
    // t0 = (s(y-1) + s(y+1)) * 3 + s(y) * 10 ; t1 = s(y+1) - s(y-1) for the six columns x-1 .. x+4
    int t0[6], t1[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const int a = s_t[r][3 + cx + k], m = s_t[r + 1][3 + cx + k], c = s_t[r + 2][3 + cx + k];
      t0[k] = (a + c) * 3 + m * 10;
      t1[k] = c - a;
    }
    short2 o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) o[k] = make_short2((short)(t0[k + 2] - t0[k]), (short)((t1[k + 2] + t1[k]) * 3 + t1[k + 1] * 10));
    short2* dst = der + (size_t)y * w + x;
    if (vecStore) {
      auto pk = [](short2 v) { return (unsigned)(unsigned short)v.x | ((unsigned)(unsigned short)v.y << 16); };
      *reinterpret_cast<uint4*>(dst) = make_uint4(pk(o[0]), pk(o[1]), pk(o[2]), pk(o[3]));
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (x + k < w) dst[k] = o[k];
    }
  }
}

// ---- LKTrackerInvoker for one point by one warp
// Tensor maps of the pyramid levels 0..KLT_MAP_LEVELS-1 of both pyramids of a call, images (u8) and derivatives (short2 as u32), each a
// rank-3 tensor (x, y, frame).  A warp fetches a window with ONE cp.async.bulk.tensor.3d instead of a 36-trip row loop of dependent
// load -> store pairs (the row loops were 30 % of k_klt_track's samples).  The box starts on a 16-byte column (the hardware traps
// otherwise, scripts/probe/tmap_probe3.cu), so windows are staged with a column offset: derivative windows 40 x 36 entries (offset
// 0..3), image windows 64 x 36 bytes (offset 0..15).  Out-of-bounds entries read as zero -- exactly calcScharrDeriv's
// BORDER_CONSTANT for the derivative window; image windows that leave the image (reflect-101) keep the byte loop.
static const int KLT_MAP_LEVELS = 5;
struct alignas(64) KltMaps {
  CUtensorMap img[2][KLT_MAP_LEVELS], der[2][KLT_MAP_LEVELS];   // [0] = prevPyr, [1] = curPyr
  unsigned imgMask[2], derMask[2];
};
struct LevelView { const uint8_t* img; const short2* der; int w, h; const CUtensorMap *imgMap, *derMap; int frame; };
__device__ __forceinline__ LevelView level_view(const uint8_t* frame, const PyrGeom& G, int l, const KltMaps& M, int which, int f) {
  LevelView v;
  v.img = frame + G.off[l];
  v.der = reinterpret_cast<const short2*>(frame + G.imgBytes) + G.off[l];
  v.w = G.w[l]; v.h = G.h[l];
  v.imgMap = (l < KLT_MAP_LEVELS && ((M.imgMask[which] >> l) & 1u)) ? &M.img[which][l] : nullptr;
  v.derMap = (l < KLT_MAP_LEVELS && ((M.derMask[which] >> l) & 1u)) ? &M.der[which][l] : nullptr;
  v.frame = f;
  return v;
}
// per-warp transaction barrier: lane 0 announces the bytes and issues the loads, every lane waits for the phase
struct WarpTma { uint32_t bar; uint32_t parity; };
__device__ __forceinline__ void tma_expect(const WarpTma& t, uint32_t bytes) {
  // the windows were read / written through the generic proxy until now: order those accesses before the async-proxy writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(t.bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load3(const WarpTma& t, const CUtensorMap* map, void* dst, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(dst)),
               "l"(map), "r"(x), "r"(y), "r"(z), "r"(t.bar)
               : "memory");
}
__device__ __forceinline__ void tma_wait(WarpTma& t) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(t.bar), "r"(t.parity) : "memory");
  t.parity ^= 1u;
}
__device__ __forceinline__ int descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Stage the (win+1) x (win+1) window of `img` whose top-left pixel is (x0, y0) into shared memory, row pitch `pitch`
// bytes.  A window inside the image is copied as aligned 32-bit words (the row keeps its misalignment: pixel (x, y)
// sits at s_win[y * pitch + ((a0 + y * w) & 3) + x], a0 = the returned low address bits of the first pixel); a
// window that leaves the image is copied byte by byte with the reflect-101 rule and a0 = NO_SHIFT.
static const unsigned NO_SHIFT = 0xffffffffu;
__device__ __host__ __forceinline__ int win_pitch(int win) { return (win + 1 + 15 + 15) / 16 * 16; }   // image window: room for a 0..15 column offset, 16-byte rows
__device__ __host__ __forceinline__ int der_pitch(int win) { return (win + 1 + 3 + 3) / 4 * 4; }        // derivative window: 0..3 entries of offset, 16-byte rows
// where pixel (x, y) of a staged window sits: s_win[y * pitch + x + (uniform ? off : (off + y * w) & 3)]
struct WinRef { unsigned off; bool uniform; };
// the image maps describe the PADDED level copies (reflect-101 border): any window within the border is a plain box of them
__device__ __forceinline__ bool win_tma_ok(const LevelView& L, int x0, int y0, int win) {
  return L.imgMap != nullptr && x0 >= -KLT_PAD_X && y0 >= -KLT_PAD_Y && x0 + win + 1 <= L.w + KLT_PAD_X && y0 + win + 1 <= L.h + KLT_PAD_Y;
}
// the generic-proxy staging: aligned 32-bit words inside the image (every row keeps the source's word alignment), reflect-101 bytes
// at the border
__device__ __forceinline__ WinRef stage_window(const LevelView& L, int x0, int y0, int win, uint8_t* s_win) {
  const int lane = threadIdx.x & 31, ww = win + 1, pitch = win_pitch(win);
  const bool inside = x0 >= 0 && y0 >= 0 && x0 + ww <= L.w && y0 + ww <= L.h;
  WinRef ref;
  ref.off = 0; ref.uniform = true;
  if (inside) {
    const uint8_t* p0 = L.img + (size_t)y0 * L.w + x0;
    ref.off = (unsigned)(reinterpret_cast<size_t>(p0) & 3);
    ref.uniform = (L.w & 3) == 0;
    const int wprS = pitch / 4, nW = (ww + 3 + 3) >> 2;   // words per shared-memory row / words copied per row (covers any alignment)
    int r = 0, k = lane;
    while (k >= nW) { k -= nW; r++; }
    for (int i = lane; i < ww * nW; i += 32) {
      const uint8_t* row = p0 + (size_t)r * L.w;
      const unsigned* src = reinterpret_cast<const unsigned*>(row - (reinterpret_cast<size_t>(row) & 3));
      reinterpret_cast<unsigned*>(s_win)[r * wprS + k] = __ldg(src + k);
      k += 32;
      while (k >= nW) { k -= nW; r++; }
    }
  } else {
    // rows outer, lanes over the columns: the reflected column of a lane is the same for every row
    const int c0 = lane < ww ? reflect101(x0 + lane, L.w) : 0, c1 = lane + 32 < ww ? reflect101(x0 + lane + 32, L.w) : 0;
    for (int y = 0; y < ww; y++) {
      const uint8_t* row = L.img + (size_t)reflect101(y0 + y, L.h) * L.w;
      if (lane < ww) s_win[y * pitch + lane] = row[c0];
      if (lane + 32 < ww) s_win[y * pitch + lane + 32] = row[c1];
    }
  }
  __syncwarp();
  return ref;
}
__device__ __forceinline__ const uint8_t* win_px(const uint8_t* s_win, int pitch, WinRef ref, int w, int x, int y) {
  return s_win + y * pitch + x + (ref.uniform ? (int)ref.off : (int)((ref.off + (unsigned)(y * w)) & 3u));
}
// One pyramid level.  (nx, ny) in/out as OpenCV's nextPts[ptidx]; status / err updated as the invoker does.
__device__ void lk_level(const LevelView& I, const LevelView& J, int level, int maxLevel, int win, int maxCount, float eps2,
                         bool useInitial, float px, float py, float& nx, float& ny, int& status, float& err, short* s_I, short2* s_dI,
                         uint8_t* s_win, WarpTma& T) {
  const int lane = threadIdx.x & 31;
  const float halfWin = (win - 1) * 0.5f;
  const float sc = (float)(1. / (1 << level));
  float prevx = px * sc, prevy = py * sc;
  float nextx, nexty;
  if (level == maxLevel) {
    if (useInitial) { nextx = nx * sc; nexty = ny * sc; }
    else { nextx = prevx; nexty = prevy; }
  } else {
    nextx = nx * 2.f; nexty = ny * 2.f;
  }
  nx = nextx; ny = nexty;
  prevx -= halfWin; prevy -= halfWin;
  const int ipx = (int)floorf(prevx), ipy = (int)floorf(prevy);
  if (ipx < -win || ipx >= I.w || ipy < -win || ipy >= I.h) {
    if (level == 0) { status = 0; err = 0; }
    return;
  }
  float a = prevx - ipx, b = prevy - ipy;
  const int W_BITS = 14;
  const float FLT_SCALE = 1.f / (1 << 20);
  int iw00 = __float2int_rn((1.f - a) * (1.f - b) * (1 << W_BITS));
  int iw01 = __float2int_rn(a * (1.f - b) * (1 << W_BITS));
  int iw10 = __float2int_rn((1.f - a) * b * (1 << W_BITS));
  int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
  __syncwarp();
  const int pitch = win_pitch(win), DP = der_pitch(win), ww = win + 1;
  // template side: the image window and the (win+1)^2 derivative window (zero outside the image: derivBorder = BORDER_CONSTANT).
  // Through the tensor maps both are in flight together (one transaction phase); otherwise row loops.
  const bool tI = win_tma_ok(I, ipx, ipy, win), tD = I.derMap != nullptr;
  if ((tI || tD) && lane == 0) {
    tma_expect(T, (uint32_t)((tI ? pitch * ww : 0) + (tD ? DP * ww * 4 : 0)));
    if (tI) tma_load3(T, I.imgMap, s_win, (ipx + KLT_PAD_X) & ~15, ipy + KLT_PAD_Y, I.frame);
    if (tD) tma_load3(T, I.derMap, s_dI, ipx & ~3, ipy, I.frame);
  }
  WinRef aI;
  aI.off = (unsigned)((ipx + KLT_PAD_X) & 15); aI.uniform = true;
  if (!tI) aI = stage_window(I, ipx, ipy, win, s_win);
  const int dOff = tD ? (ipx & 3) : 0;   // column of the window's first entry inside its staged row
  long long sA11 = 0, sA12 = 0, sA22 = 0;
  {
    // the derivative window is interpolated IN PLACE: output (x, y) is written to entry (x, y) of the row and needs the inputs at
    // (dOff + x .. + 1, y .. y + 1), which no later output of the raster order reads once this batch of 32 has loaded them
    const int dw = win + 1;
    if (!tD) {
      // columns 0..31: one row per trip, every lane its own column (the column test is loop-invariant)
      const int X = ipx + lane;
      const bool colIn = lane < dw && X >= 0 && X < I.w;
      for (int y = 0; y < dw; y++) {
        const int Y = ipy + y;
        short2 v = make_short2(0, 0);
        if (colIn && Y >= 0 && Y < I.h) v = I.der[(size_t)Y * I.w + X];
        if (lane < dw) s_dI[y * DP + lane] = v;
      }
      // the columns beyond 32 (four of them for the 35 x 35 window): flattened over the lanes instead of a second
      // trip per row with four lanes at work
      const int rem = dw - 32;
      for (int i = lane; i < rem * dw; i += 32) {
        const int y = rem == 4 ? (i >> 2) : i / rem, x = 32 + i - y * rem;
        const int X2 = ipx + x, Y = ipy + y;
        s_dI[y * DP + x] = (Y >= 0 && Y < I.h && X2 >= 0 && X2 < I.w) ? I.der[(size_t)Y * I.w + X2] : make_short2(0, 0);
      }
    }
    if (tI || tD) tma_wait(T);
    __syncwarp();
    int x = lane, y = 0;
    while (x >= win) { x -= win; y++; }
    const int n = win * win;
    for (int i0 = 0; i0 < n; i0 += 32) {  // uniform trip count: every lane takes part in the barriers
      const int i = i0 + lane;
      const bool valid = i < n;
      int ival = 0, ixval = 0, iyval = 0;
      if (valid) {
        const short2* dr = s_dI + y * DP + dOff + x;
        const short2 d00 = dr[0], d01 = dr[1], d10 = dr[DP], d11 = dr[DP + 1];
        const uint8_t* s0 = win_px(s_win, pitch, aI, I.w, x, y);
        const uint8_t* s1 = win_px(s_win, pitch, aI, I.w, x, y + 1);
        ival = descale(s0[0] * iw00 + s0[1] * iw01 + s1[0] * iw10 + s1[1] * iw11, W_BITS - 5);
        ixval = descale(d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11, W_BITS);
        iyval = descale(d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11, W_BITS);
        sA11 += (long long)ixval * ixval; sA12 += (long long)ixval * iyval; sA22 += (long long)iyval * iyval;
      }
      __syncwarp();
      if (valid) {
        s_I[i] = (short)ival;
        s_dI[y * DP + x] = make_short2((short)ixval, (short)iyval);
      }
      __syncwarp();
      x += 32;
      while (x >= win) { x -= win; y++; }
    }
  }
  sA11 = warp_sum_ll(sA11); sA12 = warp_sum_ll(sA12); sA22 = warp_sum_ll(sA22);
  const float A11 = (float)sA11 * FLT_SCALE, A12 = (float)sA12 * FLT_SCALE, A22 = (float)sA22 * FLT_SCALE;
  float D = A11 * A22 - A12 * A12;
  const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
  err = minEig;  // OPTFLOW_LK_GET_MIN_EIGENVALS
  if (minEig < 1e-4f || D < FLT_EPSILON) {
    if (level == 0) status = 0;
    return;
  }
  D = 1.f / D;
  nextx -= halfWin; nexty -= halfWin;
  float pdx = 0, pdy = 0;
  int stagedX = INT_MIN, stagedY = INT_MIN;   // origin of the J window s_win holds (it held the I window until here)
  WinRef aJ;
  aJ.off = 0; aJ.uniform = true;
  for (int j = 0; j < maxCount; j++) {
    const int inx = (int)floorf(nextx), iny = (int)floorf(nexty);
    if (inx < -win || inx >= J.w || iny < -win || iny >= J.h) {
      if (level == 0) status = 0;
      break;
    }
    a = nextx - inx; b = nexty - iny;
    iw00 = __float2int_rn((1.f - a) * (1.f - b) * (1 << W_BITS));
    iw01 = __float2int_rn(a * (1.f - b) * (1 << W_BITS));
    iw10 = __float2int_rn((1.f - a) * b * (1 << W_BITS));
    iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
    __syncwarp();
    // the window is staged again only when its integer origin moved: once the iteration is down to sub-pixel steps
    // (most of its trips) the bytes in shared memory are already the ones it needs
    if (inx != stagedX || iny != stagedY) {
      if (win_tma_ok(J, inx, iny, win)) {
        if (lane == 0) {
          tma_expect(T, (uint32_t)(pitch * ww));
          tma_load3(T, J.imgMap, s_win, (inx + KLT_PAD_X) & ~15, iny + KLT_PAD_Y, J.frame);
        }
        aJ.off = (unsigned)((inx + KLT_PAD_X) & 15); aJ.uniform = true;
        tma_wait(T);
      } else {
        aJ = stage_window(J, inx, iny, win, s_win);
      }
      stagedX = inx; stagedY = iny;
    }
    long long sb1 = 0, sb2 = 0;
    if (((win * win + 31) >> 5) <= 64 && aJ.uniform) {
      // Common case (level widths that are multiples of 4, or a reflected window): every staged row has the same
      // misalignment, so a pixel's address is one multiply-add; and a lane's <= 64 products (|diff| <= 8160 =
      // 255 * 32, |d| <= 4080 = 16 * 255: each < 2^25) sum exactly in 32 bits.
      // Every lane takes ONE contiguous run of the window in raster order (39 pixels of the 35 x 35 window) instead of every
      // 32nd pixel: the left column of a pixel's 2 x 2 neighbourhood is the right column of the pixel before it, so a step
      // loads two new bytes instead of four and advances its pointers by one (the sums are exact integers: any order gives
      // the same bits).
      const uint8_t* wbase = s_win + (int)aJ.off;
      int a1 = 0, a2 = 0;
      const int n = win * win, per = (n + 31) >> 5;
      int i = lane * per;
      const int iend = min(i + per, n);
      if (i < iend) {
        int y = i / win, x = i - y * win;
        const uint8_t* s0 = wbase + y * pitch + x;
        const short2* dp = s_dI + y * DP + x;
        int p00 = s0[0], p10 = s0[pitch];
        for (; i < iend; i++) {
          const int p01 = s0[1], p11 = s0[pitch + 1];
          const int diff = descale(p00 * iw00 + p01 * iw01 + p10 * iw10 + p11 * iw11, W_BITS - 5) - s_I[i];
          const short2 d = *dp;
          a1 += diff * d.x;
          a2 += diff * d.y;
          if (++x == win) {   // next row: skip the window's extra column, reload the left column
            x = 0;
            s0 += pitch - win + 1; dp += DP - win + 1;
            p00 = s0[0]; p10 = s0[pitch];
          } else {
            s0++; dp++;
            p00 = p01; p10 = p11;
          }
        }
      }
      sb1 = a1; sb2 = a2;
    } else {
      int x = lane, y = 0;
      while (x >= win) { x -= win; y++; }
      for (int i = lane; i < win * win; i += 32) {
        const uint8_t* s0 = win_px(s_win, pitch, aJ, J.w, x, y);
        const uint8_t* s1 = win_px(s_win, pitch, aJ, J.w, x, y + 1);
        const int diff = descale(s0[0] * iw00 + s0[1] * iw01 + s1[0] * iw10 + s1[1] * iw11, W_BITS - 5) - s_I[i];
        const short2 d = s_dI[y * DP + x];
        sb1 += (long long)(diff * d.x);  // |diff| < 2^14, |d| < 2^13: the product fits an int
        sb2 += (long long)(diff * d.y);
        x += 32;
        while (x >= win) { x -= win; y++; }
      }
    }
    sb1 = warp_sum_ll(sb1); sb2 = warp_sum_ll(sb2);
    const float b1 = (float)sb1 * FLT_SCALE, b2 = (float)sb2 * FLT_SCALE;
    const float dx = (A12 * b2 - A22 * b1) * D, dy = (A12 * b1 - A11 * b2) * D;
    nextx += dx; nexty += dy;
    nx = nextx + halfWin; ny = nexty + halfWin;
    if ((double)dx * dx + (double)dy * dy <= (double)eps2) break;
    if (j > 0 && (double)fabsf(dx + pdx) < 0.01 && (double)fabsf(dy + pdy) < 0.01) {
      nx -= dx * 0.5f; ny -= dy * 0.5f;
      break;
    }
    pdx = dx; pdy = dy;
  }
}

struct TrackArgs {
  const uint8_t *prevPyr, *curPyr;   // [batch][frameBytes]
  const float* kps;                  // [batch][stride][2]
  float* next;                       // [batch][stride][2] in/out
  const int* n;                      // [batch]
  uint8_t* status;                   // [batch][stride]
  float* err;                        // [batch][stride] (may be null)
  int stride, win, maxLevel, maxCount, useInitial, fb;
  float eps2, ferr, fbDist;
};

// mode fb = 0: cv::calcOpticalFlowPyrLK(prev, cur) with OPTFLOW_LK_GET_MIN_EIGENVALS; fb = 1: ORBmatcher::fbKltTracking
__device__ __host__ __forceinline__ size_t up128(size_t v) { return (v + 127) / 128 * 128; }
__global__ void __launch_bounds__(TRACK_WARPS * 32) k_klt_track(TrackArgs A, PyrGeom G, const __grid_constant__ KltMaps M) {
  extern __shared__ __align__(16) unsigned char s_raw0[];
  __shared__ __align__(8) unsigned long long s_bar[TRACK_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.y, p = blockIdx.x * TRACK_WARPS + warp;
  if (p >= A.n[f]) return;
  const int win = A.win, ww = win + 1;
  // per warp: derivative window / template derivatives, image window, template intensities -- each on a 128-byte boundary
  // (the tensor-map loads want their destination aligned)
  unsigned char* s_raw = s_raw0 + ((128u - ((uint32_t)__cvta_generic_to_shared(s_raw0) & 127u)) & 127u);
  const size_t szD = up128((size_t)der_pitch(win) * ww * 4), szW = up128((size_t)win_pitch(win) * ww), szI = up128((size_t)win * win * 2);
  unsigned char* base = s_raw + (szD + szW + szI) * warp;
  short2* s_dI = reinterpret_cast<short2*>(base);
  uint8_t* s_win = base + szD;
  short* s_I = reinterpret_cast<short*>(base + szD + szW);
  WarpTma T;
  T.bar = (uint32_t)__cvta_generic_to_shared(&s_bar[warp]);
  T.parity = 0u;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(T.bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint8_t* P = A.prevPyr + (size_t)f * G.frameBytes;
  const uint8_t* C = A.curPyr + (size_t)f * G.frameBytes;
  const size_t o = (size_t)f * A.stride + p;
  const float kx = A.kps[2 * o], ky = A.kps[2 * o + 1];
  float nx = A.next[2 * o], ny = A.next[2 * o + 1];
  int status = 1;
  float err = 0;
  const int maxLevel = min(A.maxLevel, G.levels);
  for (int l = maxLevel; l >= 0; l--)
    lk_level(level_view(P, G, l, M, 0, f), level_view(C, G, l, M, 1, f), l, maxLevel, win, A.maxCount, A.eps2, A.useInitial != 0, kx, ky, nx,
             ny, status, err, s_I, s_dI, s_win, T);
  if (!A.fb) {
    if (lane == 0) { A.next[2 * o] = nx; A.next[2 * o + 1] = ny; A.status[o] = (uint8_t)status; if (A.err) A.err[o] = err; }
    return;
  }
  // fbKltTracking: forward filter, backward pass at level 0 from the tracked position with the keypoint as initial flow
  bool ok = status != 0 && !(err > A.ferr) && 1.f <= nx && nx < (float)G.w[0] - 1.f && 1.f <= ny && ny < (float)G.h[0] - 1.f;
  if (ok) {
    float bx = kx, by = ky, err2 = 0;
    int st2 = 1;
    lk_level(level_view(C, G, 0, M, 1, f), level_view(P, G, 0, M, 0, f), 0, 0, win, A.maxCount, A.eps2, true, nx, ny, bx, by, st2, err2, s_I, s_dI,
             s_win, T);
    const double ddx = (double)kx - (double)bx, ddy = (double)ky - (double)by;
    ok = st2 != 0 && !(sqrt(ddx * ddx + ddy * ddy) > (double)A.fbDist);
  }
  if (lane == 0) { A.next[2 * o] = nx; A.next[2 * o + 1] = ny; A.status[o] = ok ? 1 : 0; if (A.err) A.err[o] = err; }
}

}  // namespace klt
}  // namespace gfs

using namespace gfs;
using namespace gfs::klt;

struct GfsKlt {
  int maxW = 0, maxH = 0, levels = 0, maxPoints = 0, maxBatch = 0;
  DevBuf d_img, d_pyrA, d_pyrB, d_kps, d_next, d_n, d_status, d_err;
  int launches = 0;
};

extern "C" {

int gfs_klt_create(int max_w, int max_h, int levels, int max_points, int max_batch, GfsKlt** out) {
  GFS_REQUIRE(out, GFS_ERR_INVALID, "out is null");
  *out = nullptr;
  GFS_REQUIRE(max_w > 0 && max_h > 0 && levels >= 0 && levels <= MAX_LEVELS && max_points > 0 && max_batch > 0, GFS_ERR_INVALID,
              "bad capacity");
  int rc = gfs_device_check();
  if (rc) return rc;
  GfsKlt* h = new GfsKlt();
  h->maxW = max_w; h->maxH = max_h; h->levels = levels; h->maxPoints = max_points; h->maxBatch = max_batch;
  *out = h;
  return GFS_OK;
}

int gfs_klt_destroy(GfsKlt* h) {
  if (!h) return GFS_OK;
  DevBuf* d[] = {&h->d_img, &h->d_pyrA, &h->d_pyrB, &h->d_kps, &h->d_next, &h->d_n, &h->d_status, &h->d_err};
  for (DevBuf* b : d) b->release();
  delete h;
  return GFS_OK;
}

int gfs_klt_last_launches(const GfsKlt* h) { return h ? h->launches : GFS_ERR_INVALID; }

size_t gfs_klt_pyramid_bytes(const GfsKlt* h, int w, int h_img) {
  if (!h || w <= 0 || h_img <= 0) return 0;
  return make_geom(w, h_img, h->levels).frameBytes;
}
int gfs_klt_pyramid_layout(const GfsKlt* h, int w, int h_img, int* level_w, int* level_h, int* level_off, size_t* deriv_offset) {
  GFS_REQUIRE(h && level_w && level_h && level_off && deriv_offset, GFS_ERR_INVALID, "null argument");
  const PyrGeom g = make_geom(w, h_img, h->levels);
  for (int l = 0; l <= g.levels; l++) { level_w[l] = g.w[l]; level_h[l] = g.h[l]; level_off[l] = g.off[l]; }
  *deriv_offset = g.imgBytes;
  return GFS_OK;
}

int gfs_klt_build_pyramid_batch_device(GfsKlt* h, void* stream, const uint8_t* d_imgs, int batch, int w, int h_img, int pitch,
                                       size_t img_stride, uint8_t* d_pyr) {
  GFS_REQUIRE(h && d_imgs && d_pyr, GFS_ERR_INVALID, "null argument");
  GFS_REQUIRE(batch > 0 && batch <= h->maxBatch, GFS_ERR_CAPACITY, "batch exceeds the handle's max_batch");
  GFS_REQUIRE(w > 0 && h_img > 0 && w <= h->maxW && h_img <= h->maxH && pitch >= w, GFS_ERR_CAPACITY, "image larger than the handle's size");
  cudaStream_t st = (cudaStream_t)stream;
  const PyrGeom g = make_geom(w, h_img, h->levels);
  // level 0 = the image itself (buildOpticalFlowPyramid copies it into the padded pyramid)
  {
    const bool vec = (w % 16 == 0) && (pitch % 16 == 0) && (img_stride % 16 == 0) && (g.frameBytes % 16 == 0) &&
                     (reinterpret_cast<size_t>(d_imgs) % 16 == 0) && (reinterpret_cast<size_t>(d_pyr) % 16 == 0);
    const int work = vec ? (w / 16) * h_img : w * h_img;
    k_klt_copy_level0<<<dim3(std::min(div_up(work, 256), 148 * 4), batch), 256, 0, st>>>(d_imgs, pitch, img_stride, w, h_img, d_pyr,
                                                                                     g.frameBytes, vec ? 1 : 0);
  }
  h->launches = 1;
  for (int l = 0; l < g.levels; l++) {
    k_klt_pyr_down<<<dim3(div_up(g.w[l + 1], PD_TX), div_up(g.h[l + 1], PD_TY), batch), dim3(PD_TX, PD_TY), 0, st>>>(
        d_pyr, g.frameBytes, g.off[l], g.w[l], g.h[l], g.off[l + 1], g.w[l + 1], g.h[l + 1]);
    h->launches++;
  }
  k_klt_scharr<<<dim3(g.tileBase[g.levels + 1], batch), 256, 0, st>>>(d_pyr, g);
  h->launches++;
  k_klt_pad_level<<<dim3(div_up(g.padW[0] / 16 * g.padH[0], 256), batch, g.levels + 1), 256, 0, st>>>(d_pyr, g);
  h->launches++;
  GFS_CUDA(cudaGetLastError());
  return GFS_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tmap_encoder() {
  static EncodeTiledFn fn = [] {
    if (const char* e = getenv("GFS_KLT_TMAP")) if (atoi(e) == 0) return (EncodeTiledFn) nullptr;   // A/B runs: row loops only
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return (EncodeTiledFn) nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}
// one pyramid level of `frames` frames as a rank-3 tensor (x, y, frame) of `esize`-byte entries; false: no map possible
static bool encode_level_map(CUtensorMap* m, const void* base, int esize, int w, int hgt, int frames, size_t frameStride, int boxW, int boxH) {
  EncodeTiledFn enc = tmap_encoder();
  const size_t pitch = (size_t)w * esize;
  if (!enc || (((uintptr_t)base | pitch | frameStride) & 15) != 0 || frameStride < pitch * hgt || boxW > 256 || boxH > 256 ||
      ((boxW * esize) & 15))
    return false;
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)hgt, (cuuint64_t)frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frameStride};
  const cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1u};
  const cuuint32_t es[3] = {1u, 1u, 1u};
  return enc(m, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int launch_track(GfsKlt* h, cudaStream_t st, const TrackArgs& A, const PyrGeom& g, int batch, int maxPts) {
  const int win = A.win, ww = win + 1;
  const size_t perWarp = up128((size_t)der_pitch(win) * ww * 4) + up128((size_t)win_pitch(win) * ww) + up128((size_t)win * win * 2);
  const size_t smem = perWarp * TRACK_WARPS + 128;   // + the slack the kernel uses to align its window to 128 bytes
  if (smem > 48 * 1024) GFS_CUDA(cudaFuncSetAttribute(k_klt_track, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KltMaps M;
  memset(&M, 0, sizeof(M));
  for (int which = 0; which < 2; which++) {
    const uint8_t* pyr = which ? A.curPyr : A.prevPyr;
    for (int l = 0; l <= g.levels && l < KLT_MAP_LEVELS; l++) {
      if (ww <= std::min(KLT_PAD_X, KLT_PAD_Y) &&
          encode_level_map(&M.img[which][l], pyr + g.padOff[l], 1, g.padW[l], g.padH[l], batch, g.frameBytes, win_pitch(win), ww))
        M.imgMask[which] |= 1u << l;
      if (encode_level_map(&M.der[which][l], pyr + g.imgBytes + (size_t)4 * g.off[l], 4, g.w[l], g.h[l], batch, g.frameBytes, der_pitch(win), ww))
        M.derMask[which] |= 1u << l;
    }
  }
  k_klt_track<<<dim3(div_up(maxPts, TRACK_WARPS), batch), TRACK_WARPS * 32, smem, st>>>(A, g, M);
  h->launches++;
  GFS_CUDA(cudaGetLastError());
  return GFS_OK;
}

int gfs_klt_calc_batch_device(GfsKlt* h, void* stream, const uint8_t* d_prev_pyr, const uint8_t* d_cur_pyr, int batch, int w, int h_img,
                              const float* d_pts, float* d_next, const int* d_n, int stride, int win, int max_level, int max_count,
                              float eps, int use_initial_flow, uint8_t* d_status, float* d_err) {
  GFS_REQUIRE(h && d_prev_pyr && d_cur_pyr && d_pts && d_next && d_n && d_status, GFS_ERR_INVALID, "null argument");
  GFS_REQUIRE(batch > 0 && batch <= h->maxBatch && stride > 0 && stride <= h->maxPoints, GFS_ERR_CAPACITY, "batch / stride exceed the handle");
  GFS_REQUIRE(win >= 3 && win <= 63 && max_level >= 0, GFS_ERR_INVALID, "bad window / level");
  const PyrGeom g = make_geom(w, h_img, h->levels);
  TrackArgs A;
  memset(&A, 0, sizeof(A));
  A.prevPyr = d_prev_pyr; A.curPyr = d_cur_pyr; A.kps = d_pts; A.next = d_next; A.n = d_n; A.status = d_status; A.err = d_err;
  A.stride = stride; A.win = win; A.maxLevel = std::min(max_level, effective_levels(w, h_img, h->levels, win)); A.maxCount = std::min(std::max(max_count, 0), 100);
  const float e = std::min(std::max(eps, 0.f), 10.f);
  A.eps2 = e * e; A.useInitial = use_initial_flow; A.fb = 0;
  h->launches = 0;
  return launch_track(h, (cudaStream_t)stream, A, g, batch, stride);
}

int gfs_klt_fb_track_batch_device(GfsKlt* h, void* stream, const uint8_t* d_prev_pyr, const uint8_t* d_cur_pyr, int batch, int w, int h_img,
                                  const float* d_kps, float* d_priors, const int* d_n, int stride, int win, int nbpyrlvl, float ferr,
                                  float fmax_fbklt_dist, uint8_t* d_status) {
  GFS_REQUIRE(h && d_prev_pyr && d_cur_pyr && d_kps && d_priors && d_n && d_status, GFS_ERR_INVALID, "null argument");
  GFS_REQUIRE(batch > 0 && batch <= h->maxBatch && stride > 0 && stride <= h->maxPoints, GFS_ERR_CAPACITY, "batch / stride exceed the handle");
  GFS_REQUIRE(win >= 3 && win <= 63 && nbpyrlvl >= 0, GFS_ERR_INVALID, "bad window / level");
  const PyrGeom g = make_geom(w, h_img, h->levels);
  TrackArgs A;
  memset(&A, 0, sizeof(A));
  A.prevPyr = d_prev_pyr; A.curPyr = d_cur_pyr; A.kps = d_kps; A.next = d_priors; A.n = d_n; A.status = d_status; A.err = nullptr;
  A.stride = stride; A.win = win; A.maxLevel = std::min(nbpyrlvl, effective_levels(w, h_img, h->levels, win)); A.maxCount = 30; A.eps2 = 0.01f * 0.01f; A.useInitial = 1; A.fb = 1;
  A.ferr = ferr; A.fbDist = fmax_fbklt_dist;
  h->launches = 0;
  return launch_track(h, (cudaStream_t)stream, A, g, batch, stride);
}

// Host-pointer convenience: one frame pair, images in, tracks out (both pyramids are built on the way).
int gfs_klt_fb_track(GfsKlt* h, void* stream, const uint8_t* prev_img, const uint8_t* cur_img, int w, int h_img, int pitch, const float* kps,
                     float* priors, int n, int win, int nbpyrlvl, float ferr, float fmax_fbklt_dist, uint8_t* status) {
  GFS_REQUIRE(h && prev_img && cur_img, GFS_ERR_INVALID, "null image");
  GFS_REQUIRE(n >= 0 && n <= h->maxPoints, GFS_ERR_CAPACITY, "n exceeds the handle's max_points");
  GFS_REQUIRE(w > 0 && h_img > 0 && w <= h->maxW && h_img <= h->maxH && pitch >= w, GFS_ERR_CAPACITY, "image larger than the handle's size");
  if (n == 0) return GFS_OK;  // fbKltTracking returns before touching its outputs (ORBmatcher.cc:2197-2199)
  GFS_REQUIRE(kps && priors && status, GFS_ERR_INVALID, "null point arrays");
  cudaStream_t st = (cudaStream_t)stream;
  const PyrGeom g = make_geom(w, h_img, h->levels);
  int rc;
  if ((rc = h->d_img.reserve((size_t)2 * pitch * h_img)) || (rc = h->d_pyrA.reserve(g.frameBytes)) || (rc = h->d_pyrB.reserve(g.frameBytes)) ||
      (rc = h->d_kps.reserve((size_t)n * 8)) || (rc = h->d_next.reserve((size_t)n * 8)) || (rc = h->d_n.reserve(4)) ||
      (rc = h->d_status.reserve((size_t)n)))
    return rc;
  uint8_t* dimg = (uint8_t*)h->d_img.p;
  GFS_CUDA(cudaMemcpyAsync(dimg, prev_img, (size_t)pitch * h_img, cudaMemcpyHostToDevice, st));
  GFS_CUDA(cudaMemcpyAsync(dimg + (size_t)pitch * h_img, cur_img, (size_t)pitch * h_img, cudaMemcpyHostToDevice, st));
  GFS_CUDA(cudaMemcpyAsync(h->d_kps.p, kps, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  GFS_CUDA(cudaMemcpyAsync(h->d_next.p, priors, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  GFS_CUDA(cudaMemcpyAsync(h->d_n.p, &n, 4, cudaMemcpyHostToDevice, st));
  if ((rc = gfs_klt_build_pyramid_batch_device(h, stream, dimg, 1, w, h_img, pitch, 0, (uint8_t*)h->d_pyrA.p))) return rc;
  int l1 = h->launches;
  if ((rc = gfs_klt_build_pyramid_batch_device(h, stream, dimg + (size_t)pitch * h_img, 1, w, h_img, pitch, 0, (uint8_t*)h->d_pyrB.p))) return rc;
  l1 += h->launches;
  rc = gfs_klt_fb_track_batch_device(h, stream, (const uint8_t*)h->d_pyrA.p, (const uint8_t*)h->d_pyrB.p, 1, w, h_img, (const float*)h->d_kps.p,
                                     (float*)h->d_next.p, (const int*)h->d_n.p, n, win, nbpyrlvl, ferr, fmax_fbklt_dist, (uint8_t*)h->d_status.p);
  if (rc) return rc;
  h->launches += l1;
  GFS_CUDA(cudaMemcpyAsync(priors, h->d_next.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  GFS_CUDA(cudaMemcpyAsync(status, h->d_status.p, (size_t)n, cudaMemcpyDeviceToHost, st));
  GFS_CUDA(gfs::stream_wait(st));
  return GFS_OK;
}

// cv::createCLAHE(clip_limit, Size(tiles_x, tiles_y))->apply for a batch of 8-bit gray frames (device pointers);
// d_dst may alias d_src (Frame::Frame applies it in place).
int gfs_clahe_apply_batch_device(void* stream, const uint8_t* d_src, int batch, int w, int h_img, int pitch, size_t img_stride,
                                 double clip_limit, int tiles_x, int tiles_y, uint8_t* d_dst, int dst_pitch, size_t dst_stride) {
  GFS_REQUIRE(d_src && d_dst, GFS_ERR_INVALID, "null pointer");
  GFS_REQUIRE(batch > 0 && w > 0 && h_img > 0 && pitch >= w && dst_pitch >= w && tiles_x > 0 && tiles_y > 0, GFS_ERR_INVALID, "bad geometry");
  int rc = gfs_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  ClaheGeom G;
  G.w = w; G.h = h_img; G.tilesX = tiles_x; G.tilesY = tiles_y;
  int ew = w, eh = h_img;  // copyMakeBorder(BORDER_REFLECT_101) on the right / bottom when the size is not divisible
  if (!(w % tiles_x == 0 && h_img % tiles_y == 0)) { ew = w + (tiles_x - (w % tiles_x)); eh = h_img + (tiles_y - (h_img % tiles_y)); }
  G.tw = ew / tiles_x; G.th = eh / tiles_y;
  const int total = G.tw * G.th;
  G.lutScale = static_cast<float>(255) / total;
  G.clipLimit = 0;
  if (clip_limit > 0.0) G.clipLimit = std::max(static_cast<int>(clip_limit * total / 256), 1);
  uint8_t* lut = nullptr;
  GFS_CUDA(cudaMallocAsync((void**)&lut, (size_t)batch * tiles_x * tiles_y * 256, st));
  k_clahe_lut<<<dim3(tiles_x * tiles_y, batch), 256, 0, st>>>(d_src, pitch, img_stride, G, lut);
  k_clahe_apply<<<dim3(div_up(w, 32), div_up(h_img, 8), batch), 256, 0, st>>>(d_src, pitch, img_stride, G, lut, d_dst, dst_pitch, dst_stride);
  const cudaError_t e = cudaGetLastError();
  GFS_CUDA(cudaFreeAsync(lut, st));
  GFS_CUDA(e);
  return GFS_OK;
}
// one frame, HOST pointers
int gfs_clahe_apply(void* stream, const uint8_t* src, int w, int h_img, int pitch, double clip_limit, int tiles_x, int tiles_y, uint8_t* dst) {
  GFS_REQUIRE(src && dst && w > 0 && h_img > 0 && pitch >= w, GFS_ERR_INVALID, "bad argument");
  int rc = gfs_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* d = nullptr;
  GFS_CUDA(cudaMallocAsync((void**)&d, (size_t)pitch * h_img, st));
  GFS_CUDA(cudaMemcpyAsync(d, src, (size_t)pitch * h_img, cudaMemcpyHostToDevice, st));
  rc = gfs_clahe_apply_batch_device(stream, d, 1, w, h_img, pitch, 0, clip_limit, tiles_x, tiles_y, d, pitch, 0);
  if (rc == GFS_OK) {
    const cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)w, d, (size_t)pitch, (size_t)w, (size_t)h_img, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) { gfs::set_error("cudaMemcpy2DAsync -> %s", cudaGetErrorString(e)); rc = GFS_ERR_CUDA; }
  }
  cudaFreeAsync(d, st);
  if (rc) return rc;
  GFS_CUDA(gfs::stream_wait(st));
  return GFS_OK;
}

}  // extern "C"
