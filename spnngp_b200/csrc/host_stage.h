// host_stage.h -- pageable host arrays -> device through a small ring of pinned buffers, filled by a few
// host threads while earlier chunks are in flight (nngp_set_data's upload path).
//
// A plain cudaMemcpy from pageable memory stages through one driver thread (~6-8 GB/s here); the working
// set of a fit (coords, neighbour ids, X, y: ~100 MB at n = 10^6, m = 15) then costs more wall time than
// twenty sampler iterations.  Here T threads each own two pinned chunks and a stream: thread t takes chunks
// t, t+T, ... of the job list, converts/copies one into pinned memory (fusing the (x, y) interleave and the
// neighbour-id range check into that pass) and queues its DMA, so host copies and PCIe transfers overlap.
#pragma once
#include <cuda_runtime.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

namespace nngp {

struct StageJob {
  enum Kind { RAW, XY_INTERLEAVE, INT_RANGE };
  Kind kind;
  void *dst;        // device (run) / host, pageable (run_d2h)
  const void *src;  // host, pageable (run) / device (run_d2h)
  size_t bytes;     // bytes written to dst
  long long n;      // XY_INTERLEAVE: number of sites (src = x[n] then y[n]); INT_RANGE: valid ids are [0, n)
};

// true when `p` is page-locked host memory the device can DMA from / to directly (cudaHostAlloc, cudaHostRegister,
// nngp_host_alloc): such arrays skip the pinned ring altogether
inline bool host_is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

class HostStager {
 public:
  static constexpr size_t kChunk = (size_t)2 << 20;
  static constexpr int kMaxThreads = 8;

  ~HostStager() {
    for (int t = 0; t < nthreads_; t++) {
      for (int b = 0; b < 2; b++) {
        if (ev_[t][b]) cudaEventDestroy(ev_[t][b]);
        if (buf_[t][b]) cudaFreeHost(buf_[t][b]);
      }
      if (st_[t]) cudaStreamDestroy(st_[t]);
    }
  }

  cudaError_t init(int device) {
    device_ = device;
    unsigned hc = std::thread::hardware_concurrency();
    nthreads_ = hc >= 16 ? kMaxThreads : (hc >= 8 ? 4 : (hc >= 4 ? 2 : 1));  // multi-GPU boxes run one stager per rank
    for (int t = 0; t < nthreads_; t++) {
      cudaError_t e = cudaStreamCreateWithFlags(&st_[t], cudaStreamNonBlocking);
      if (e != cudaSuccess) return e;
      for (int b = 0; b < 2; b++) {
        if ((e = cudaHostAlloc(&buf_[t][b], kChunk, cudaHostAllocDefault)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&ev_[t][b], cudaEventDisableTiming)) != cudaSuccess) return e;
      }
    }
    return cudaSuccess;
  }

  // Runs all jobs; returns when every byte is on the device.  *bad_id is set if an INT_RANGE job saw an id
  // outside [0, n) (the upload still completes).
  // *min_id / *max_id (optional): the smallest / largest id any INT_RANGE job saw (INT_MAX / -1 if none).
  cudaError_t run(const std::vector<StageJob> &jobs, bool *bad_id, int *min_id = nullptr, int *max_id = nullptr) {
    struct Chunk { int job; size_t off, len; };
    std::vector<Chunk> chunks;
    bool direct = false;
    for (size_t j = 0; j < jobs.size(); j++) {
      if (jobs[j].kind == StageJob::RAW && jobs[j].bytes >= kChunk && host_is_pinned(jobs[j].src)) {
        // the caller's array is page-locked: one DMA straight from it (PCIe rate instead of the host's memcpy rate)
        cudaError_t e = cudaMemcpyAsync(jobs[j].dst, jobs[j].src, jobs[j].bytes, cudaMemcpyHostToDevice, st_[0]);
        if (e != cudaSuccess) return e;
        direct = true;
        continue;
      }
      for (size_t off = 0; off < jobs[j].bytes; off += kChunk)
        chunks.push_back({(int)j, off, jobs[j].bytes - off < kChunk ? jobs[j].bytes - off : kChunk});
    }
    if (chunks.empty()) {
      if (bad_id) *bad_id = false;
      if (min_id) *min_id = 2147483647;
      if (max_id) *max_id = -1;
      return direct ? cudaStreamSynchronize(st_[0]) : cudaSuccess;
    }
    std::atomic<int> err{(int)cudaSuccess};
    std::atomic<bool> bad{false};
    std::atomic<int> lowest{2147483647}, highest{-1};
    auto worker = [&](int t) {
      cudaSetDevice(device_);
      int b = 0;
      for (size_t ci = (size_t)t; ci < chunks.size(); ci += (size_t)nthreads_, b ^= 1) {
        const Chunk &ch = chunks[ci];
        const StageJob &J = jobs[ch.job];
        cudaError_t e = cudaEventSynchronize(ev_[t][b]);  // this buffer's previous DMA has finished
        if (e != cudaSuccess) { err = (int)e; return; }
        char *pin = (char *)buf_[t][b];
        if (J.kind == StageJob::RAW) {
          memcpy(pin, (const char *)J.src + ch.off, ch.len);
        } else if (J.kind == StageJob::XY_INTERLEAVE) {
          const double *x = (const double *)J.src, *y = x + J.n;
          const size_t i0 = ch.off / 16, cnt = ch.len / 16;
          double *o = (double *)pin;
          for (size_t i = 0; i < cnt; i++) {
            o[2 * i] = x[i0 + i];
            o[2 * i + 1] = y[i0 + i];
          }
        } else {
          const int *s = (const int *)((const char *)J.src + ch.off);
          int *o = (int *)pin;
          const size_t cnt = ch.len / 4;
          unsigned worst = 0;  // as unsigned, a negative id is huge
          int low = 2147483647;
          for (size_t i = 0; i < cnt; i++) {
            const int v = s[i];
            o[i] = v;
            worst = (unsigned)v > worst ? (unsigned)v : worst;
            low = v < low ? v : low;
          }
          if ((long long)worst >= J.n) bad = true;
          int cur = lowest.load();
          while (low < cur && !lowest.compare_exchange_weak(cur, low)) {
          }
          const int high = (worst > 2147483647u) ? 2147483647 : (int)worst;
          cur = highest.load();
          while (high > cur && !highest.compare_exchange_weak(cur, high)) {
          }
        }
        e = cudaMemcpyAsync((char *)J.dst + ch.off, pin, ch.len, cudaMemcpyHostToDevice, st_[t]);
        if (e == cudaSuccess) e = cudaEventRecord(ev_[t][b], st_[t]);
        if (e != cudaSuccess) { err = (int)e; return; }
      }
      cudaError_t e = cudaStreamSynchronize(st_[t]);
      if (e != cudaSuccess) err = (int)e;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads_; t++) th.emplace_back(worker, t);
    worker(0);
    for (auto &x : th) x.join();
    if (bad_id) *bad_id = bad.load();
    if (min_id) *min_id = lowest.load();
    if (max_id) *max_id = highest.load();
    return (cudaError_t)err.load();
  }

  // The other direction: device -> pageable host arrays (RAW jobs; dst = host, src = device).  Each thread drains
  // its chunks through its two pinned buffers: the DMA of chunk k+1 runs while chunk k is copied out (first touch of
  // a fresh result array page-faults, which is why the threads matter more here than on the way in).
  cudaError_t run_d2h(const std::vector<StageJob> &jobs) {
    struct Chunk { int job; size_t off, len; };
    std::vector<Chunk> chunks;
    // page-locked destination (checked once: the jobs of a call write into one result array): straight DMAs
    const bool direct = !jobs.empty() && host_is_pinned(jobs.front().dst) && host_is_pinned(jobs.back().dst);
    if (direct) {
      for (size_t j = 0; j < jobs.size(); j++) {
        cudaError_t e = cudaMemcpyAsync(jobs[j].dst, jobs[j].src, jobs[j].bytes, cudaMemcpyDeviceToHost, st_[0]);
        if (e != cudaSuccess) return e;
      }
      return cudaStreamSynchronize(st_[0]);
    }
    for (size_t j = 0; j < jobs.size(); j++)
      for (size_t off = 0; off < jobs[j].bytes; off += kChunk)
        chunks.push_back({(int)j, off, jobs[j].bytes - off < kChunk ? jobs[j].bytes - off : kChunk});
    std::atomic<int> err{(int)cudaSuccess};
    auto worker = [&](int t) {
      cudaSetDevice(device_);
      std::vector<size_t> mine;
      for (size_t ci = (size_t)t; ci < chunks.size(); ci += (size_t)nthreads_) mine.push_back(ci);
      auto issue = [&](size_t k) {
        const Chunk &ch = chunks[mine[k]];
        cudaError_t e = cudaMemcpyAsync(buf_[t][k & 1], (const char *)jobs[ch.job].src + ch.off, ch.len,
                                        cudaMemcpyDeviceToHost, st_[t]);
        if (e == cudaSuccess) e = cudaEventRecord(ev_[t][k & 1], st_[t]);
        if (e != cudaSuccess) err = (int)e;
      };
      if (!mine.empty()) issue(0);
      for (size_t k = 0; k < mine.size(); k++) {
        if (k + 1 < mine.size()) issue(k + 1);
        cudaError_t e = cudaEventSynchronize(ev_[t][k & 1]);
        if (e != cudaSuccess) { err = (int)e; return; }
        const Chunk &ch = chunks[mine[k]];
        memcpy((char *)jobs[ch.job].dst + ch.off, buf_[t][k & 1], ch.len);
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads_; t++) th.emplace_back(worker, t);
    worker(0);
    for (auto &x : th) x.join();
    return (cudaError_t)err.load();
  }

 private:
  int device_ = 0, nthreads_ = 1;
  cudaStream_t st_[kMaxThreads] = {};
  void *buf_[kMaxThreads][2] = {};
  cudaEvent_t ev_[kMaxThreads][2] = {};
};

}  // namespace nngp
