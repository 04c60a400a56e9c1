// Batched tracking front-end: ORB extraction of `batch` frames followed by BF-Hamming + GMS
// between consecutive frames (frame i -> frame i+1).  This is the per-frame sequence
// Frame::ExtractORB (reference src/Frame.cc:768-777 -> ORBextractor::operator()) then
// ORBmatcher::SearchWithGMS (src/ORBmatcher.cc:744-778) of System::TrackRGBD, applied to a batch
// of independent frames -- BASELINE.json configs[1].
#include <algorithm>
#include <vector>

#include <cstdlib>

#include "common.cuh"

namespace gfs {
// orb.cu: one chunk of frames on scratch slots slot0 .. slot0+nb-1 of the extractor
int orb_extract_slots(GfsOrb* h, cudaStream_t st, int slot0, const uint8_t* d_imgs, int nb, int w, int h_img, int pitch,
                      size_t img_stride, GfsKeyPoint* d_kp, uint8_t* d_desc, int* d_n, int* d_mono);
}
using namespace gfs;

struct GfsFrontend {
  GfsOrb* orb = nullptr;
  int maxBatch = 0, stride = 0;
  DevBuf d_in, d_kp, d_desc, d_n, d_idx, d_dist, d_inl, d_cnt;
  PinnedBuf h_in;
  // host-buffer path: H2D of chunk k+1 and D2H of chunk k-1 overlap the kernels of chunk k
  cudaStream_t copyStream = nullptr, outStream = nullptr, auxStream = nullptr;
  cudaStream_t aux[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // compute streams 1..7 (aux[0] == auxStream)
  cudaEvent_t evStart = nullptr;
  std::vector<cudaEvent_t> evIn, evDone, evExt;
  int chunk = 128;
  bool profiling = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // start, after orb, after bf, after gms
};

extern "C" {

int gfs_frontend_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, int max_w,
                        int max_h, int max_batch, GfsFrontend** out) {
  GFS_REQUIRE(out, GFS_ERR_INVALID, "out is null");
  *out = nullptr;
  GfsOrb* orb = nullptr;
  int rc = gfs_orb_create(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast, max_w, max_h, max_batch, &orb);
  if (rc) return rc;
  GfsFrontend* f = new GfsFrontend();
  f->orb = orb;
  f->maxBatch = max_batch;
  f->stride = gfs_orb_max_keypoints(orb);
  *out = f;
  return GFS_OK;
}

int gfs_frontend_destroy(GfsFrontend* f) {
  if (!f) return GFS_OK;
  gfs_orb_destroy(f->orb);
  DevBuf* d[] = {&f->d_in, &f->d_kp, &f->d_desc, &f->d_n, &f->d_idx, &f->d_dist, &f->d_inl, &f->d_cnt};
  for (DevBuf* b : d) b->release();
  f->h_in.release();
  for (cudaEvent_t e : f->ev)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : f->evIn) cudaEventDestroy(e);
  for (cudaEvent_t e : f->evDone) cudaEventDestroy(e);
  for (cudaEvent_t e : f->evExt) cudaEventDestroy(e);
  if (f->copyStream) cudaStreamDestroy(f->copyStream);
  if (f->outStream) cudaStreamDestroy(f->outStream);
  for (cudaStream_t a : f->aux)
    if (a) cudaStreamDestroy(a);
  if (f->evStart) cudaEventDestroy(f->evStart);
  delete f;
  return GFS_OK;
}

int gfs_frontend_max_keypoints(const GfsFrontend* f) { return f ? f->stride : GFS_ERR_INVALID; }
GfsOrb* gfs_frontend_orb(GfsFrontend* f) { return f ? f->orb : nullptr; }

int gfs_frontend_set_profiling(GfsFrontend* f, int enable) {
  GFS_REQUIRE(f, GFS_ERR_INVALID, "null handle");
  if (enable && !f->ev[0])
    for (int i = 0; i < 4; i++) GFS_CUDA(cudaEventCreate(&f->ev[i]));
  f->profiling = enable != 0;
  return gfs_orb_set_profiling(f->orb, enable);
}

// ms8 = {pyramid, fast_cells, octree, blur, orient_desc, pack_lapping, bf_hamming, gms}
int gfs_frontend_get_profile(GfsFrontend* f, float* ms8) {
  GFS_REQUIRE(f && ms8 && f->profiling, GFS_ERR_INVALID, "profiling not enabled");
  int rc = gfs_orb_get_profile(f->orb, ms8);
  if (rc) return rc;
  GFS_CUDA(cudaEventSynchronize(f->ev[3]));
  GFS_CUDA(cudaEventElapsedTime(&ms8[6], f->ev[1], f->ev[2]));
  GFS_CUDA(cudaEventElapsedTime(&ms8[7], f->ev[2], f->ev[3]));
  return GFS_OK;
}

int gfs_frontend_launches_per_call(const GfsFrontend* f, int batch) {
  if (!f) return GFS_ERR_INVALID;
  // gfs_frontend_run_device: the extractor splits batches of >= 128 frames into two halves on two streams
  // (gfs_orb_extract_batch_device), each half launching the whole kernel sequence; then matcher + GMS
  const int halves = (!f->profiling && batch >= 128 && batch <= f->maxBatch) ? 2 : 1;
  return halves * gfs_orb_launches_per_call(f->orb, 0, 0) + (batch > 1 ? 2 : 0);
}

int gfs_frontend_run_device(GfsFrontend* f, void* stream, const uint8_t* d_imgs, int batch, int w, int h_img,
                            int pitch, size_t img_stride, GfsKeyPoint* d_kp, uint8_t* d_desc, int* d_n, int* d_mono,
                            int* d_train_idx, int* d_dist, uint8_t* d_inlier, int* d_inlier_count) {
  GFS_REQUIRE(f, GFS_ERR_INVALID, "null handle");
  GFS_REQUIRE(batch > 0 && batch <= f->maxBatch, GFS_ERR_CAPACITY, "batch exceeds the handle's max_batch");
  cudaStream_t st = (cudaStream_t)stream;
  if (f->profiling) cudaEventRecord(f->ev[0], st);
  int rc = gfs_orb_extract_batch_device(f->orb, stream, d_imgs, batch, w, h_img, pitch, img_stride, 0, 0, d_kp, d_desc,
                                        d_n, d_mono);
  if (rc) return rc;
  if (f->profiling) cudaEventRecord(f->ev[1], st);
  if (batch > 1) {
    const int s = f->stride;
    // pair p: query = frame p, train = frame p + 1
    rc = gfs_match_bf_hamming_batch_device(stream, d_desc, d_n, d_desc + (size_t)s * 32, d_n + 1, batch - 1, s,
                                           d_train_idx, d_dist);
    if (rc) return rc;
    if (f->profiling) cudaEventRecord(f->ev[2], st);
    rc = gfs_gms_filter_batch_device(stream, d_kp, d_n, d_kp + s, d_n + 1, d_train_idx, batch - 1, s, w, h_img, w,
                                     h_img, d_inlier, d_inlier_count);
    if (rc) return rc;
  } else if (f->profiling) {
    cudaEventRecord(f->ev[2], st);
  }
  if (f->profiling) cudaEventRecord(f->ev[3], st);
  return GFS_OK;
}

// Host-pointer variant: H2D of the images, the device pipeline, D2H of every result.
// Outputs: out_kp [batch][stride], out_desc [batch][stride][32], out_n/out_mono [batch],
// out_train_idx/out_dist [batch-1][stride], out_inlier [batch-1][stride], out_inlier_count [batch-1].
int gfs_frontend_run(GfsFrontend* f, void* stream, const uint8_t* imgs, int batch, int w, int h_img, int pitch,
                     size_t img_stride, GfsKeyPoint* out_kp, uint8_t* out_desc, int* out_n, int* out_mono,
                     int* out_train_idx, int* out_dist, uint8_t* out_inlier, int* out_inlier_count) {
  GFS_REQUIRE(f, GFS_ERR_INVALID, "null handle");
  GFS_REQUIRE(imgs && w > 0 && h_img > 0, GFS_ERR_EMPTY, "empty image");
  GFS_REQUIRE(batch > 0 && batch <= f->maxBatch, GFS_ERR_CAPACITY, "batch exceeds the handle's max_batch");
  GFS_REQUIRE(out_kp && out_desc && out_n && out_mono, GFS_ERR_INVALID, "null output");
  GFS_REQUIRE(batch == 1 || (out_train_idx && out_dist && out_inlier && out_inlier_count), GFS_ERR_INVALID, "null match output");
  cudaStream_t st = (cudaStream_t)stream;
  const int s = f->stride;
  const size_t B = (size_t)batch, dpitch = align_up((size_t)w, 16), dstride = dpitch * h_img;
  int rc;
  if ((rc = f->d_in.reserve(B * dstride))) return rc;
  if ((rc = f->d_kp.reserve(B * s * sizeof(GfsKeyPoint)))) return rc;
  if ((rc = f->d_desc.reserve(B * s * 32))) return rc;
  if ((rc = f->d_n.reserve(B * 2 * sizeof(int)))) return rc;
  if ((rc = f->d_idx.reserve(B * s * sizeof(int)))) return rc;
  if ((rc = f->d_dist.reserve(B * s * sizeof(int)))) return rc;
  if ((rc = f->d_inl.reserve(B * s))) return rc;
  if ((rc = f->d_cnt.reserve(B * sizeof(int)))) return rc;
  int* d_n = (int*)f->d_n.p;
  int* d_mono = d_n + batch;
  GfsKeyPoint* d_kp = (GfsKeyPoint*)f->d_kp.p;
  uint8_t* d_desc = (uint8_t*)f->d_desc.p;
  uint8_t* d_in = (uint8_t*)f->d_in.p;
  const bool pinned = is_pinned_host(imgs) && is_pinned_host(out_kp) && is_pinned_host(out_desc) &&
                      (batch == 1 || (is_pinned_host(out_train_idx) && is_pinned_host(out_dist) && is_pinned_host(out_inlier)));
  bool matchedInChunks = false;
  // H2D is faster than the kernels, so after the first chunk the copy stream stays ahead: only the first
  // copy is exposed.  Chunks of batch/12 keep it short while the kernels still see >= 64 frames.
  // tuning knobs, read once per process
  static const int envChunks = [] { const char* e = getenv("GFS_FRONTEND_CHUNKS"); return e ? std::max(1, atoi(e)) : 12; }();
  static const int envFirst = [] { const char* e = getenv("GFS_FRONTEND_FIRST"); return e ? std::max(1, atoi(e)) : 0; }();
  static const int envStreams = [] { const char* e = getenv("GFS_FRONTEND_STREAMS"); return e ? std::min(8, std::max(1, atoi(e))) : 8; }();
  f->chunk = std::max(64, div_up(batch, envChunks));
  const int firstChunk = envFirst ? std::min(f->chunk, envFirst) : std::max(32, f->chunk / 2);
  if (pinned && batch > f->chunk) {
    // ---- pipelined: chunked H2D on a copy stream, kernels on the caller's stream, D2H on a third
    if (!f->copyStream) {
      GFS_CUDA(cudaStreamCreateWithFlags(&f->copyStream, cudaStreamNonBlocking));
      GFS_CUDA(cudaStreamCreateWithFlags(&f->outStream, cudaStreamNonBlocking));
      for (int i = 0; i < 7; i++) GFS_CUDA(cudaStreamCreateWithFlags(&f->aux[i], cudaStreamNonBlocking));
      f->auxStream = f->aux[0];
      GFS_CUDA(cudaEventCreateWithFlags(&f->evStart, cudaEventDisableTiming));
    }
    const int nChunks = 1 + div_up(batch - firstChunk, f->chunk);
    while ((int)f->evIn.size() < nChunks) {
      cudaEvent_t a, b2, c2;
      GFS_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
      GFS_CUDA(cudaEventCreateWithFlags(&b2, cudaEventDisableTiming));
      GFS_CUDA(cudaEventCreateWithFlags(&c2, cudaEventDisableTiming));
      f->evIn.push_back(a);
      f->evDone.push_back(b2);
      f->evExt.push_back(c2);
    }
    // the side streams must not run ahead of work already queued on the caller's stream
    GFS_CUDA(cudaEventRecord(f->evStart, st));
    GFS_CUDA(cudaStreamWaitEvent(f->copyStream, f->evStart, 0));
    GFS_CUDA(cudaStreamWaitEvent(f->outStream, f->evStart, 0));
    for (int i = 0; i < 7; i++) GFS_CUDA(cudaStreamWaitEvent(f->aux[i], f->evStart, 0));
    // Chunks rotate over nStreams compute streams: the latency-bound quadtree kernel of a chunk (~0.4 ms whatever
    // the chunk size) needs the throughput kernels of several other chunks to hide behind.  Measured (B200, 1024
    // VGA frames, frames/s end to end; profiles/r01_summary.md): 8 chunks on 2 / 3 / 4 streams 100.0k / 104.9k /
    // 108.8k; with the final kernels 4 streams x 8 chunks 123.5k, 6 x 10 125.3k, 8 x 12 127.0k (default).
    const int nStreams = envStreams;
    cudaStream_t streams[8] = {st, f->aux[0], f->aux[1], f->aux[2], f->aux[3], f->aux[4], f->aux[5], f->aux[6]};
    // A chunk's frame
    // pairs (and the pair that straddles the previous chunk) are matched and copied out right behind
    // its extraction, so the tail after the last chunk is one chunk's matcher, not the batch's.
    for (int c = 0; c < nChunks; c++) {
      const size_t b0 = c == 0 ? 0 : (size_t)firstChunk + (size_t)(c - 1) * f->chunk;
      const size_t nb = c == 0 ? (size_t)firstChunk : std::min<size_t>(f->chunk, B - b0);
      cudaStream_t cs = streams[c % nStreams];
      if (img_stride == (size_t)pitch * h_img) {
        GFS_CUDA(cudaMemcpy2DAsync(d_in + b0 * dstride, dpitch, imgs + b0 * img_stride, pitch, w, (size_t)h_img * nb,
                                   cudaMemcpyHostToDevice, f->copyStream));
      } else {
        for (size_t i = 0; i < nb; i++)
          GFS_CUDA(cudaMemcpy2DAsync(d_in + (b0 + i) * dstride, dpitch, imgs + (b0 + i) * img_stride, pitch, w, h_img,
                                     cudaMemcpyHostToDevice, f->copyStream));
      }
      GFS_CUDA(cudaEventRecord(f->evIn[c], f->copyStream));
      GFS_CUDA(cudaStreamWaitEvent(cs, f->evIn[c], 0));
      rc = orb_extract_slots(f->orb, cs, (int)b0, d_in + b0 * dstride, (int)nb, w, h_img, (int)dpitch, dstride, d_kp + b0 * s,
                             d_desc + b0 * s * 32, d_n + b0, d_mono + b0);
      if (rc) return rc;
      GFS_CUDA(cudaEventRecord(f->evExt[c], cs));
      // the pairs both of whose frames are extracted now: [p0, p1); frame b0 - 1 belongs to the previous chunk
      const size_t p0 = c == 0 ? 0 : b0 - 1, p1 = b0 + nb - 1;
      if (p1 > p0) {
        if (c > 0) GFS_CUDA(cudaStreamWaitEvent(cs, f->evExt[c - 1], 0));
        rc = gfs_match_bf_hamming_batch_device(cs, d_desc + p0 * s * 32, d_n + p0, d_desc + (p0 + 1) * s * 32, d_n + p0 + 1,
                                               (int)(p1 - p0), s, (int*)f->d_idx.p + p0 * s, (int*)f->d_dist.p + p0 * s);
        if (rc) return rc;
        rc = gfs_gms_filter_batch_device(cs, d_kp + p0 * s, d_n + p0, d_kp + (p0 + 1) * s, d_n + p0 + 1,
                                         (int*)f->d_idx.p + p0 * s, (int)(p1 - p0), s, w, h_img, w, h_img,
                                         (uint8_t*)f->d_inl.p + p0 * s, (int*)f->d_cnt.p + p0);
        if (rc) return rc;
      }
      GFS_CUDA(cudaEventRecord(f->evDone[c], cs));
      GFS_CUDA(cudaStreamWaitEvent(f->outStream, f->evDone[c], 0));
      GFS_CUDA(cudaMemcpyAsync(out_kp + b0 * s, d_kp + b0 * s, nb * s * sizeof(GfsKeyPoint), cudaMemcpyDeviceToHost, f->outStream));
      GFS_CUDA(cudaMemcpyAsync(out_desc + b0 * s * 32, d_desc + b0 * s * 32, nb * s * 32, cudaMemcpyDeviceToHost, f->outStream));
      if (p1 > p0) {
        const size_t np = p1 - p0;
        GFS_CUDA(cudaMemcpyAsync(out_train_idx + p0 * s, (int*)f->d_idx.p + p0 * s, np * s * sizeof(int), cudaMemcpyDeviceToHost, f->outStream));
        GFS_CUDA(cudaMemcpyAsync(out_dist + p0 * s, (int*)f->d_dist.p + p0 * s, np * s * sizeof(int), cudaMemcpyDeviceToHost, f->outStream));
        GFS_CUDA(cudaMemcpyAsync(out_inlier + p0 * s, (uint8_t*)f->d_inl.p + p0 * s, np * s, cudaMemcpyDeviceToHost, f->outStream));
      }
    }
    for (int c = 0; c < nChunks; c++)
      if (c % nStreams) GFS_CUDA(cudaStreamWaitEvent(st, f->evDone[c], 0));
    if (f->profiling) for (int i = 0; i < 4; i++) cudaEventRecord(f->ev[i], st);  // stages interleave: no per-stage split here
    matchedInChunks = true;
  } else {
    const uint8_t* src = imgs;
    size_t spitch = pitch;
    if (!is_pinned_host(imgs)) {  // pageable input: stage once through pinned memory
      if ((rc = f->h_in.reserve(B * (size_t)w * h_img))) return rc;
      uint8_t* stg = (uint8_t*)f->h_in.p;
      for (size_t i = 0; i < B; i++)
        for (int y = 0; y < h_img; y++) memcpy(stg + (i * h_img + y) * w, imgs + i * img_stride + (size_t)y * pitch, w);
      src = stg;
      spitch = w;
      img_stride = (size_t)w * h_img;
    }
    if (img_stride == spitch * h_img) {
      GFS_CUDA(cudaMemcpy2DAsync(d_in, dpitch, src, spitch, w, (size_t)h_img * B, cudaMemcpyHostToDevice, st));
    } else {
      for (size_t i = 0; i < B; i++)
        GFS_CUDA(cudaMemcpy2DAsync(d_in + i * dstride, dpitch, src + i * img_stride, spitch, w, h_img, cudaMemcpyHostToDevice, st));
    }
    rc = gfs_frontend_run_device(f, stream, d_in, batch, w, h_img, (int)dpitch, dstride, d_kp, d_desc, d_n, d_mono,
                                 (int*)f->d_idx.p, (int*)f->d_dist.p, (uint8_t*)f->d_inl.p, (int*)f->d_cnt.p);
    if (rc) return rc;
    GFS_CUDA(cudaMemcpyAsync(out_kp, d_kp, B * s * sizeof(GfsKeyPoint), cudaMemcpyDeviceToHost, st));
    GFS_CUDA(cudaMemcpyAsync(out_desc, d_desc, B * s * 32, cudaMemcpyDeviceToHost, st));
  }
  GFS_CUDA(cudaMemcpyAsync(out_n, d_n, B * sizeof(int), cudaMemcpyDeviceToHost, st));
  GFS_CUDA(cudaMemcpyAsync(out_mono, d_mono, B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (batch > 1) {
    if (!matchedInChunks) {
      GFS_CUDA(cudaMemcpyAsync(out_train_idx, f->d_idx.p, (B - 1) * s * sizeof(int), cudaMemcpyDeviceToHost, st));
      GFS_CUDA(cudaMemcpyAsync(out_dist, f->d_dist.p, (B - 1) * s * sizeof(int), cudaMemcpyDeviceToHost, st));
      GFS_CUDA(cudaMemcpyAsync(out_inlier, f->d_inl.p, (B - 1) * s, cudaMemcpyDeviceToHost, st));
    }
    GFS_CUDA(cudaMemcpyAsync(out_inlier_count, f->d_cnt.p, (B - 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  if (f->outStream) GFS_CUDA(cudaStreamSynchronize(f->outStream));
  GFS_CUDA(cudaStreamSynchronize(st));
  return GFS_OK;
}
}
