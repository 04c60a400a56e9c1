// Projection-window descriptor search and depth back-projection on sm_100a.
//
//   k_search_by_projection   ORBmatcher::SearchByProjection, both overloads (reference
//                            src/ORBmatcher.cc:43-207 local-map search = mode 1, :1853-2063 frame-to-
//                            frame search = mode 0) over Frame::GetFeaturesInArea (src/Frame.cc:1007-1071)
//                            and the 64x48 Frame grid (AssignFeaturesToGrid / PosInGrid, :734-761, :1073-1084).
//   k_depth_to_cloud         Frame::ConvertDepthToPointCloud (src/Frame.cc:590-623).
//
// The reference loop is order dependent: a keypoint already taken by an earlier map point with
// Observations() > 0 is skipped (:87-88, :1924-1925).  The kernel reproduces first-come-first-served
// exactly by iterating to the fixed point of "query i ignores keypoints claimed by a blocking query
// j < i": query 0 never depends on anyone, query 1 only on query 0, ... so the fixed point is the
// sequential result; it is reached in 2-3 rounds because few map points compete for a keypoint.
#include <climits>
#include <vector>

#include "common.cuh"

namespace gfs {

static const int PG_COLS = 64, PG_ROWS = 48, PG_CELLS = PG_COLS * PG_ROWS;
static const int P_THREADS = 256;
static const int P_TH_HIGH = 100, P_HISTO = 30;

__device__ __forceinline__ int desc_dist(const uint32_t (&a)[8], const uint32_t* __restrict__ b) {
  int d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d += __popc(a[i] ^ __ldg(b + i));
  return d;
}

__global__ void __launch_bounds__(P_THREADS) k_search_by_projection(
    int mode, float nnratio, int checkOri, const GfsProjQuery* __restrict__ queries, const int* __restrict__ nqArr, int qstride,
    const GfsKeyPoint* __restrict__ kps, const float* __restrict__ uRight, const uint8_t* __restrict__ desc,
    const uint8_t* __restrict__ occupied, const int* __restrict__ nArr, int kstride, float minX, float minY, float invW,
    float invH, int* __restrict__ claim /*[pairs][qstride] scratch*/, int* __restrict__ out_assign, int* __restrict__ out_n) {
  extern __shared__ int psm[];
  int* cellStart = psm;                       // [PG_CELLS + 1]
  int* prevMin = cellStart + PG_CELLS + 1;    // [kstride]
  int* newMin = prevMin + kstride;            // [kstride]
  unsigned short* items = (unsigned short*)(newMin + kstride);  // [kstride]
  __shared__ int s_hist[P_HISTO], s_ind[3], s_count, s_carry;
  const int f = blockIdx.x, tid = threadIdx.x;
  const int nq = nqArr[f], n = nArr[f];
  const GfsProjQuery* Q = queries + (size_t)f * qstride;
  const GfsKeyPoint* K = kps + (size_t)f * kstride;
  const float* UR = uRight + (size_t)f * kstride;
  const uint8_t* DS = desc + (size_t)f * kstride * 32;
  const uint8_t* OC = occupied ? occupied + (size_t)f * kstride : nullptr;
  int* CL = claim + (size_t)f * qstride;
  int* AS = out_assign + (size_t)f * kstride;

  // ---- Frame grid: counting sort of the keypoints by cell, lists ascending in keypoint index
  for (int i = tid; i <= PG_CELLS; i += P_THREADS) cellStart[i] = 0;
  if (tid == 0) { s_count = 0; s_carry = 0; }
  __syncthreads();
  auto cell_of = [&](int i) -> int {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(K[i].x, minX), invW)), py = (int)roundf(__fmul_rn(__fsub_rn(K[i].y, minY), invH));
    if (px < 0 || px >= PG_COLS || py < 0 || py >= PG_ROWS) return -1;
    return px * PG_ROWS + py;
  };
  for (int i = tid; i < n; i += P_THREADS) {
    const int c = cell_of(i);
    if (c >= 0) atomicAdd(&cellStart[c + 1], 1);
  }
  __syncthreads();
  // inclusive scan of cellStart[1..PG_CELLS] by warp 0 in chunks of 32
  if (tid < 32) {
    int carry = 0;
    for (int base = 1; base <= PG_CELLS; base += 32) {
      const int i = base + tid;
      const int v = (i <= PG_CELLS) ? cellStart[i] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (tid >= o) inc += t;
      }
      if (i <= PG_CELLS) cellStart[i] = carry + inc;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();
  // fill: position = start + rank among the cell's keypoints with a smaller index (keeps index order)
  for (int i = tid; i < n; i += P_THREADS) prevMin[i] = cell_of(i);  // prevMin doubles as the cell-id scratch
  __syncthreads();
  for (int i = tid; i < n; i += P_THREADS) {
    const int c = prevMin[i];
    if (c < 0) continue;
    int r = 0;
    for (int j = 0; j < i; j++) r += (prevMin[j] == c);
    items[cellStart[c] + r] = (unsigned short)i;
  }
  __syncthreads();
  for (int i = tid; i < n; i += P_THREADS) { prevMin[i] = INT_MAX; newMin[i] = INT_MAX; }
  __syncthreads();

  // ---- fixed-point rounds
  for (int round = 0; round <= nq + 1; round++) {
    for (int i = tid; i < nq; i += P_THREADS) {
      const GfsProjQuery q = Q[i];
      int cl = -1;
      if (q.radius >= 0.f) {
        const float x = q.u, y = q.v, r = q.radius;
        const int c0x = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, minX), r), invW)));
        const int c1x = min(PG_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, minX), r), invW)));
        const int c0y = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, minY), r), invH)));
        const int c1y = min(PG_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, minY), r), invH)));
        if (c0x < PG_COLS && c1x >= 0 && c0y < PG_ROWS && c1y >= 0) {
          uint32_t qd[8];
#pragma unroll
          for (int w = 0; w < 8; w++) qd[w] = ((const uint32_t*)q.desc)[w];
          const bool check = (q.min_level > 0) || (q.max_level >= 0);
          int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
          for (int ix = c0x; ix <= c1x; ix++)
            for (int iy = c0y; iy <= c1y; iy++) {
              const int c = ix * PG_ROWS + iy;
              for (int p = cellStart[c]; p < cellStart[c + 1]; p++) {
                const int idx = items[p];
                const GfsKeyPoint kp = K[idx];
                if (check) {
                  if (kp.octave < q.min_level) continue;
                  if (q.max_level >= 0 && kp.octave > q.max_level) continue;
                }
                const float dx = __fsub_rn(kp.x, x), dy = __fsub_rn(kp.y, y);
                if (!(fabsf(dx) < r && fabsf(dy) < r)) continue;
                if ((OC && OC[idx]) || prevMin[idx] < i) continue;  // taken by an earlier blocking map point
                const float ur = UR[idx];
                if (ur > 0) {
                  const float er = fabsf(__fsub_rn(q.ur, ur));
                  if (er > r) continue;
                }
                const int dist = desc_dist(qd, (const uint32_t*)(DS + (size_t)idx * 32));
                if (dist < bestDist) {
                  bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kp.octave; bestIdx = idx;
                } else if (mode == 1 && dist < bestDist2) {
                  bestLevel2 = kp.octave; bestDist2 = dist;
                }
              }
            }
          if (bestDist <= P_TH_HIGH) {
            if (mode == 1) {
              const float lim = __fmul_rn(nnratio, (float)bestDist2);
              const bool reject = (bestLevel == bestLevel2) && ((float)bestDist > lim);
              if (!reject && (bestLevel != bestLevel2 || (float)bestDist <= lim)) cl = bestIdx;
            } else {
              cl = bestIdx;
            }
          }
        }
      }
      CL[i] = cl;
      if (cl >= 0 && q.blocks) atomicMin(&newMin[cl], i);
    }
    __syncthreads();
    int changed = 0;
    for (int i = tid; i < n; i += P_THREADS) {
      if (newMin[i] != prevMin[i]) changed = 1;
      prevMin[i] = newMin[i];
      newMin[i] = INT_MAX;
    }
    changed = __syncthreads_or(changed);
    if (!changed) break;
  }

  // ---- mvpMapPoints[idx] = last claimer; rotation-histogram consistency for mode 0 (:2031-2050)
  int* lastClaim = newMin;  // reuse
  for (int i = tid; i < n; i += P_THREADS) lastClaim[i] = -1;
  for (int i = tid; i < P_HISTO; i += P_THREADS) s_hist[i] = 0;
  __syncthreads();
  int mine = 0;
  const float factor = 1.0f / P_HISTO;
  for (int i = tid; i < nq; i += P_THREADS) {
    const int cl = CL[i];
    if (cl < 0) continue;
    mine++;
    atomicMax(&lastClaim[cl], i);
    if (mode == 0 && checkOri) {
      float rot = __fsub_rn(Q[i].angle, K[cl].angle);
      if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
      int bin = (int)roundf(__fmul_rn(rot, factor));
      if (bin == P_HISTO) bin = 0;
      atomicAdd(&s_hist[bin], 1);
    }
  }
  atomicAdd(&s_count, mine);
  __syncthreads();
  if (mode == 0 && checkOri) {
    if (tid == 0) {  // ComputeThreeMaxima (:2500-2532)
      int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
      for (int i = 0; i < P_HISTO; i++) {
        const int s = s_hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
      }
      if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
      s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
    }
    __syncthreads();
    int removed = 0;
    for (int i = tid; i < nq; i += P_THREADS) {
      const int cl = CL[i];
      if (cl < 0) continue;
      float rot = __fsub_rn(Q[i].angle, K[cl].angle);
      if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
      int bin = (int)roundf(__fmul_rn(rot, factor));
      if (bin == P_HISTO) bin = 0;
      if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) {
        prevMin[cl] = -2;  // mark "nulled" (prevMin is free now); any claimer in a discarded bin clears the keypoint
        removed++;
      }
    }
    atomicSub(&s_count, removed);
    __syncthreads();
    for (int i = tid; i < n; i += P_THREADS)
      if (prevMin[i] == -2) lastClaim[i] = -1;
    __syncthreads();
  }
  for (int i = tid; i < n; i += P_THREADS) AS[i] = lastClaim[i];
  if (tid == 0) out_n[f] = s_count;
}

// Frame::ConvertDepthToPointCloud: one CTA per frame, ordered compaction of the strided samples
__global__ void __launch_bounds__(1024) k_depth_to_cloud(const float* __restrict__ depth, int w, int h, int pitch_f,
                                                         long long frame_stride_f, int stride, float fx, float fy, float cx,
                                                         float cy, float4* __restrict__ out, int cap, int* __restrict__ out_n) {
  __shared__ int s_w[32];
  __shared__ int s_base;
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* D = depth + (long long)f * frame_stride_f;
  float4* O = out + (size_t)f * cap;
  const int nu = (w + stride - 1) / stride, nv = (h + stride - 1) / stride, total = nu * nv;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int b0 = 0; b0 < total; b0 += 1024) {
    const int i = b0 + tid;
    float d = 0.f;
    int u = 0, v = 0;
    bool ok = false;
    if (i < total) {
      v = (i / nu) * stride; u = (i - (i / nu) * nu) * stride;
      d = __ldg(D + (long long)v * pitch_f + u);
      ok = d > 0.0f && d < 10.0f;
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    int pos = s_base;
    for (int w2 = 0; w2 < warp; w2++) pos += s_w[w2];
    pos += __popc(m & ((1u << lane) - 1u));
    if (ok && pos < cap) {
      const float x = __fdiv_rn(__fmul_rn(__fsub_rn((float)u, cx), d), fx);
      const float y = __fdiv_rn(__fmul_rn(__fsub_rn((float)v, cy), d), fy);
      O[pos] = make_float4(x, y, d, 1.0f);
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w2 = 0; w2 < 32; w2++) t += s_w[w2];
      s_base += t;
    }
    __syncthreads();
  }
  if (tid == 0) out_n[f] = s_base;
}

}  // namespace gfs

using namespace gfs;

extern "C" {

int gfs_search_by_projection_batch_device(void* stream, int mode, float nnratio, int check_orientation,
                                          const GfsProjQuery* d_queries, const int* d_nq, int qstride,
                                          const GfsKeyPoint* d_kps_un, const float* d_u_right, const uint8_t* d_desc,
                                          const uint8_t* d_occupied, const int* d_n, int kstride, int frames, float min_x,
                                          float min_y, float inv_w, float inv_h, int* d_scratch_claim, int* d_out_assign,
                                          int* d_out_nmatches) {
  GFS_REQUIRE(d_queries && d_nq && d_kps_un && d_u_right && d_desc && d_n && d_scratch_claim && d_out_assign && d_out_nmatches,
              GFS_ERR_INVALID, "null pointer");
  GFS_REQUIRE(mode == 0 || mode == 1, GFS_ERR_INVALID, "mode must be 0 (frame-to-frame) or 1 (local map)");
  GFS_REQUIRE(frames > 0 && qstride > 0 && kstride > 0 && kstride <= 65535, GFS_ERR_INVALID, "bad sizes");
  const size_t smem = (size_t)(PG_CELLS + 1 + 2 * kstride) * 4 + (size_t)kstride * 2 + 16;
  GFS_REQUIRE(smem <= 200 * 1024, GFS_ERR_CAPACITY, "too many keypoints per frame");
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    GFS_CUDA(cudaFuncSetAttribute(k_search_by_projection, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  k_search_by_projection<<<frames, P_THREADS, smem, (cudaStream_t)stream>>>(
      mode, nnratio, check_orientation, d_queries, d_nq, qstride, d_kps_un, d_u_right, d_desc, d_occupied, d_n, kstride, min_x,
      min_y, inv_w, inv_h, d_scratch_claim, d_out_assign, d_out_nmatches);
  GFS_CUDA(cudaGetLastError());
  return GFS_OK;
}

int gfs_search_by_projection(void* stream, int mode, float nnratio, int check_orientation, const GfsProjQuery* queries, int nq,
                             const GfsKeyPoint* kps_un, const float* u_right, const uint8_t* desc, const uint8_t* occupied,
                             int n, float min_x, float min_y, float inv_w, float inv_h, int* out_assign, int* out_nmatches) {
  GFS_REQUIRE(nq >= 0 && n >= 0 && out_nmatches, GFS_ERR_INVALID, "bad sizes");
  *out_nmatches = 0;
  if (n == 0) return GFS_OK;
  GFS_REQUIRE(out_assign && kps_un && u_right && desc && (queries || nq == 0), GFS_ERR_INVALID, "null pointer");
  if (nq == 0) {
    for (int i = 0; i < n; i++) out_assign[i] = -1;
    return GFS_OK;
  }
  int rc = gfs_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  DevBuf b;
  const size_t oq = 0, ok = align_up(oq + (size_t)nq * sizeof(GfsProjQuery), 16), ou = align_up(ok + (size_t)n * sizeof(GfsKeyPoint), 16),
               od = align_up(ou + (size_t)n * 4, 16), oo = align_up(od + (size_t)n * 32, 16), oc = align_up(oo + (size_t)n, 16),
               oa = align_up(oc + (size_t)nq * 4, 16), on = align_up(oa + (size_t)n * 4, 16), tot = on + 64;
  if ((rc = b.reserve(tot))) return rc;
  uint8_t* p = (uint8_t*)b.p;
  const int hn[3] = {nq, n, 0};
  bool good = true;
  good &= cudaMemcpyAsync(p + oq, queries, (size_t)nq * sizeof(GfsProjQuery), cudaMemcpyHostToDevice, st) == cudaSuccess;
  good &= cudaMemcpyAsync(p + ok, kps_un, (size_t)n * sizeof(GfsKeyPoint), cudaMemcpyHostToDevice, st) == cudaSuccess;
  good &= cudaMemcpyAsync(p + ou, u_right, (size_t)n * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
  good &= cudaMemcpyAsync(p + od, desc, (size_t)n * 32, cudaMemcpyHostToDevice, st) == cudaSuccess;
  if (occupied) good &= cudaMemcpyAsync(p + oo, occupied, (size_t)n, cudaMemcpyHostToDevice, st) == cudaSuccess;
  good &= cudaMemcpyAsync(p + on, hn, sizeof(hn), cudaMemcpyHostToDevice, st) == cudaSuccess;
  if (!good) { set_error("H2D copy failed: %s", cudaGetErrorString(cudaGetLastError())); b.release(); return GFS_ERR_CUDA; }
  int* dn = (int*)(p + on);
  rc = gfs_search_by_projection_batch_device(stream, mode, nnratio, check_orientation, (const GfsProjQuery*)(p + oq), dn, nq,
                                             (const GfsKeyPoint*)(p + ok), (const float*)(p + ou), p + od,
                                             occupied ? p + oo : nullptr, dn + 1, n, 1, min_x, min_y, inv_w, inv_h,
                                             (int*)(p + oc), (int*)(p + oa), dn + 2);
  if (rc) { b.release(); return rc; }
  good = cudaMemcpyAsync(out_assign, p + oa, (size_t)n * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
  good &= cudaMemcpyAsync(out_nmatches, dn + 2, 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
  good &= cudaStreamSynchronize(st) == cudaSuccess;
  b.release();
  if (!good) { set_error("D2H copy failed: %s", cudaGetErrorString(cudaGetLastError())); return GFS_ERR_CUDA; }
  return GFS_OK;
}

int gfs_depth_to_cloud_batch_device(void* stream, const float* d_depth, int frames, int w, int h, int pitch_floats,
                                    size_t frame_stride_floats, int stride, float fx, float fy, float cx, float cy,
                                    float* d_out_xyz1, int cap, int* d_out_n) {
  GFS_REQUIRE(d_depth && d_out_xyz1 && d_out_n, GFS_ERR_INVALID, "null pointer");
  GFS_REQUIRE(frames > 0 && w > 0 && h > 0 && stride > 0 && cap > 0 && pitch_floats >= w, GFS_ERR_INVALID, "bad sizes");
  k_depth_to_cloud<<<frames, 1024, 0, (cudaStream_t)stream>>>(d_depth, w, h, pitch_floats, (long long)frame_stride_floats,
                                                              stride, fx, fy, cx, cy, (float4*)d_out_xyz1, cap, d_out_n);
  GFS_CUDA(cudaGetLastError());
  return GFS_OK;
}

int gfs_depth_to_cloud(void* stream, const float* depth, int w, int h, int stride, float fx, float fy, float cx, float cy,
                       float* out_xyz1, int cap, int* out_n) {
  GFS_REQUIRE(depth && out_xyz1 && out_n && w > 0 && h > 0 && stride > 0 && cap > 0, GFS_ERR_INVALID, "bad arguments");
  int rc = gfs_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  DevBuf b;
  const size_t od = 0, oo = align_up((size_t)w * h * 4, 16), on = oo + (size_t)cap * 16;
  if ((rc = b.reserve(on + 16))) return rc;
  uint8_t* p = (uint8_t*)b.p;
  if (cudaMemcpyAsync(p + od, depth, (size_t)w * h * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) {
    set_error("H2D copy failed"); b.release(); return GFS_ERR_CUDA;
  }
  rc = gfs_depth_to_cloud_batch_device(stream, (const float*)(p + od), 1, w, h, w, (size_t)w * h, stride, fx, fy, cx, cy,
                                       (float*)(p + oo), cap, (int*)(p + on));
  if (rc) { b.release(); return rc; }
  int n = 0;
  bool good = cudaMemcpyAsync(&n, p + on, 4, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
  if (good && n > 0) good = cudaMemcpy(out_xyz1, p + oo, (size_t)std::min(n, cap) * 16, cudaMemcpyDeviceToHost) == cudaSuccess;
  b.release();
  if (!good) { set_error("D2H copy failed"); return GFS_ERR_CUDA; }
  *out_n = n;
  return GFS_OK;
}
}
