// Host side of the pseudo-labelling loop (workflows/pseudo_label_generator.py:189-211 seen from the CPU): what the
// reference does per batch with `data['images'].cuda()` (:190), per image with cv2.imwrite (:43-46) and per batch with
// numpy bookkeeping (:82-105) becomes three native pieces, so that the interpreter issues a handful of foreign calls per
// WINDOW of batches and never waits for the GPU:
//
//   stager   a ring of device slots fed by cudaMemcpyAsync on a copy stream; two events per slot order the copy
//            against the consumer stream in both directions (slot free -> copy, copy done -> consumer);
//   emit     ONE call queues phase C, the mean-prob EMA, the PNG encoder and every device-to-host copy of a window
//            (file blob, offset table, per-image class counts, per-group confidence sums and thresholds);
//   writer   a persistent pool: a dispatcher thread sleeps on the window's event (cudaEventBlockingSync: no spinning
//            core), fetches what the predicted blob copy missed, and POSIX writer threads put the files on disk.
//            The Python side only waits on a ticket when it is about to reuse the window's pinned buffers.
#include <fcntl.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

using namespace hiast;

// ------------------------------------------------------------------------------------------ stager
namespace {

struct Stager {
  int device = 0;
  std::vector<cudaEvent_t> ready;        // per slot: its copy has landed (recorded on the copy stream)
  std::vector<cudaEvent_t> freed;        // per RELEASE (round robin): the consumer stream has consumed the released slots
  std::vector<long long> slot_gen;       // per slot: generation of the release that freed it (0 = never used)
  long long gen = 0;                     // releases so far
  long long waited_gen = 0;              // newest generation the copy stream already waits for
};

}  // namespace

extern "C" int hiast_stager_create(int n_slots, void** handle_out) {
  if (n_slots < 1 || n_slots > (1 << 20) || !handle_out) return HIAST_ERR_INVALID_ARG;
  std::unique_ptr<Stager> s(new Stager);
  HIAST_CUDA_TRY(cudaGetDevice(&s->device));
  s->ready.resize(n_slots);
  s->freed.resize(n_slots + 1);          // a slot's generation is at most n_slots releases old when it is pushed again
  s->slot_gen.assign(n_slots, 0);
  for (auto& e : s->ready) HIAST_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : s->freed) HIAST_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *handle_out = s.release();
  return HIAST_OK;
}

extern "C" int hiast_stager_destroy(void* handle) {
  Stager* s = static_cast<Stager*>(handle);
  if (!s) return HIAST_OK;
  for (cudaEvent_t e : s->ready) cudaEventDestroy(e);
  for (cudaEvent_t e : s->freed) cudaEventDestroy(e);
  delete s;
  return HIAST_OK;
}

extern "C" int hiast_stager_push(void* handle, int slot, void* dst_device, const void* src_host, size_t nbytes,
                                 void* copy_stream, void* consumer_stream) {
  Stager* s = static_cast<Stager*>(handle);
  if (!s || slot < 0 || slot >= static_cast<int>(s->ready.size()) || (nbytes && (!dst_device || !src_host)))
    return HIAST_ERR_INVALID_ARG;
  cudaStream_t cs = as_stream(copy_stream), ms = as_stream(consumer_stream);
  // Releases happen in order on ONE consumer stream, so waiting for generation g covers every older one: the copy stream
  // waits once per release it has not seen yet -- not once per copy.  (A cross-stream wait in front of every 5 MB copy
  // kept the DMA engine from running copies back to back: 175 us instead of 95 us per copy with four ranks on the node.)
  const long long g = s->slot_gen[slot];
  if (g > s->waited_gen) {
    HIAST_CUDA_TRY(cudaStreamWaitEvent(cs, s->freed[g % static_cast<long long>(s->freed.size())], 0));
    s->waited_gen = g;
  }
  if (nbytes) HIAST_CUDA_TRY(cudaMemcpyAsync(dst_device, src_host, nbytes, cudaMemcpyHostToDevice, cs));
  HIAST_CUDA_TRY(cudaEventRecord(s->ready[slot], cs));
  HIAST_CUDA_TRY(cudaStreamWaitEvent(ms, s->ready[slot], 0));
  return HIAST_OK;
}

extern "C" int hiast_stager_release(void* handle, int first_slot, int n_slots, void* consumer_stream) {
  Stager* s = static_cast<Stager*>(handle);
  if (!s || first_slot < 0 || n_slots < 0 || first_slot + n_slots > static_cast<int>(s->ready.size()))
    return HIAST_ERR_INVALID_ARG;
  if (n_slots == 0) return HIAST_OK;
  // one event for the whole run of slots: they were all consumed by work already queued on the consumer stream
  const long long g = ++s->gen;
  HIAST_CUDA_TRY(cudaEventRecord(s->freed[g % static_cast<long long>(s->freed.size())], as_stream(consumer_stream)));
  for (int i = 0; i < n_slots; ++i) s->slot_gen[first_slot + i] = g;
  return HIAST_OK;
}

// ------------------------------------------------------------------------------------------ emit
extern "C" int hiast_ias_emit_window(const HiastWindowEmit* a, void* stream, void* copy_stream) {
  if (!a || !a->conf || !a->label || !a->thr_groups || !a->plbl || !a->counts || !a->confsum) return HIAST_ERR_INVALID_ARG;
  if (a->n_images < 0 || a->H < 1 || a->W < 1 || a->C < 1 || a->C > HIAST_MAX_CLASSES || a->group_size < 1)
    return HIAST_ERR_INVALID_ARG;
  if (a->n_images == 0) return HIAST_OK;
  cudaStream_t st = as_stream(stream);
  const int n = a->n_images, C = a->C;
  const int g = (n + a->group_size - 1) / a->group_size;
  const int64_t HW = static_cast<int64_t>(a->H) * a->W;
  HIAST_CUDA_TRY(cudaMemsetAsync(a->counts, 0, sizeof(int64_t) * n * C, st));
  HIAST_CUDA_TRY(cudaMemsetAsync(a->confsum, 0, sizeof(uint64_t) * g * C, st));
  HIAST_TRY(hiast_ias_select(a->conf, a->label, a->thr_groups, n, HW, C, a->group_size, a->plbl, a->counts, a->confsum, stream));
  if (a->mean_state)
    HIAST_TRY(hiast_ias_meanprob_scan(a->confsum, a->counts, n, a->group_size, g, C, a->cp_gamma, a->mean_state, stream));
  if (a->blob_dev) {
    if (!a->offsets_dev || !a->png_ws || !a->offsets_host || !a->blob_host) return HIAST_ERR_INVALID_ARG;
    HIAST_TRY(hiast_png_encode(a->plbl, n, a->H, a->W, a->blob_dev, a->blob_capacity, a->offsets_dev, a->png_ws,
                               a->png_ws_bytes, stream));
  }
  // the device-to-host copies leave the compute stream: behind one event they run on copy_stream, beside the next
  // window's kernels (a window's label maps are 134 MB = 2.4 ms of PCIe at full resolution)
  cudaStream_t cs = st;
  if (copy_stream && as_stream(copy_stream) != st) {
    static thread_local cudaEvent_t bridge[64] = {nullptr};
    int dev = 0;
    HIAST_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return HIAST_ERR_UNSUPPORTED;
    if (!bridge[dev]) HIAST_CUDA_TRY(cudaEventCreateWithFlags(&bridge[dev], cudaEventDisableTiming));
    cs = as_stream(copy_stream);
    HIAST_CUDA_TRY(cudaEventRecord(bridge[dev], st));
    HIAST_CUDA_TRY(cudaStreamWaitEvent(cs, bridge[dev], 0));
  }
  if (a->blob_dev) {
    HIAST_CUDA_TRY(cudaMemcpyAsync(a->offsets_host, a->offsets_dev, sizeof(int64_t) * (n + 1), cudaMemcpyDeviceToHost, cs));
    const size_t copy = a->blob_copy_bytes < a->blob_capacity ? a->blob_copy_bytes : a->blob_capacity;
    if (copy) HIAST_CUDA_TRY(cudaMemcpyAsync(a->blob_host, a->blob_dev, copy, cudaMemcpyDeviceToHost, cs));
  } else if (a->plbl_host) {
    HIAST_CUDA_TRY(cudaMemcpyAsync(a->plbl_host, a->plbl, static_cast<size_t>(n) * HW, cudaMemcpyDeviceToHost, cs));
  }
  if (a->counts_host)
    HIAST_CUDA_TRY(cudaMemcpyAsync(a->counts_host, a->counts, sizeof(int64_t) * n * C, cudaMemcpyDeviceToHost, cs));
  if (a->confsum_host)
    HIAST_CUDA_TRY(cudaMemcpyAsync(a->confsum_host, a->confsum, sizeof(uint64_t) * g * C, cudaMemcpyDeviceToHost, cs));
  if (a->thr_groups_host)
    HIAST_CUDA_TRY(cudaMemcpyAsync(a->thr_groups_host, a->thr_groups, sizeof(double) * g * C, cudaMemcpyDeviceToHost, cs));
  return HIAST_OK;
}

// ------------------------------------------------------------------------------------------ writer
namespace {

struct WriteJob {
  int64_t ticket = 0;
  int device = 0;
  cudaEvent_t event = nullptr;
  std::vector<std::string> paths;
  const uint8_t* blob_host = nullptr;
  size_t blob_host_capacity = 0;
  const int64_t* offsets_host = nullptr;
  size_t bytes_copied = 0;
  const uint8_t* blob_dev = nullptr;
  std::atomic<int> next{0}, left{0}, first_errno{0};
  int status = HIAST_OK;
};

struct Writer {
  std::mutex mu;
  std::condition_variable cv_jobs, cv_files, cv_done;
  std::deque<std::shared_ptr<WriteJob>> pending;       // submitted, event not yet seen
  std::deque<std::shared_ptr<WriteJob>> writing;       // files being written
  std::vector<std::pair<int64_t, std::pair<int, int>>> finished;   // ticket -> (status, errno), recent
  int64_t next_ticket = 1, done_upto = 0;
  bool stop = false;
  std::thread dispatcher;
  std::vector<std::thread> workers;
  cudaStream_t topup_stream = nullptr;
  int topup_device = -1;

  void write_one(WriteJob& j, int i) {
    const int fd = ::open(j.paths[i].c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    int err = 0;
    if (fd < 0) {
      err = errno;
    } else {
      const uint8_t* p = j.blob_host + j.offsets_host[i];
      int64_t left = j.offsets_host[i + 1] - j.offsets_host[i];
      while (left > 0) {
        const ssize_t w = ::write(fd, p, static_cast<size_t>(left));
        if (w < 0) {
          if (errno == EINTR) continue;
          err = errno;
          break;
        }
        p += w;
        left -= w;
      }
      if (::close(fd) != 0 && err == 0) err = errno;
    }
    if (err != 0) {
      int expected = 0;
      j.first_errno.compare_exchange_strong(expected, err);
    }
  }

  void finish(const std::shared_ptr<WriteJob>& j) {     // mu held
    int status = j->status;
    if (status == HIAST_OK && j->first_errno.load() != 0) status = HIAST_ERR_IO;
    finished.emplace_back(j->ticket, std::make_pair(status, j->first_errno.load()));
    if (finished.size() > 64) finished.erase(finished.begin());
    for (auto it = writing.begin(); it != writing.end(); ++it)
      if (it->get() == j.get()) {
        writing.erase(it);
        break;
      }
    // tickets complete in order of submission as far as callers are concerned
    int64_t upto = next_ticket - 1;
    for (auto& p : pending) upto = std::min(upto, p->ticket - 1);
    for (auto& w : writing) upto = std::min(upto, w->ticket - 1);
    done_upto = upto;
    if (j->event) cudaEventDestroy(j->event);
    cv_done.notify_all();
  }

  void dispatch_loop() {
    for (;;) {
      std::shared_ptr<WriteJob> j;
      {
        std::unique_lock<std::mutex> lock(mu);
        cv_jobs.wait(lock, [&] { return stop || !pending.empty(); });
        if (pending.empty()) return;                     // stop and drained
        j = pending.front();
      }
      cudaSetDevice(j->device);
      cudaError_t e = cudaEventSynchronize(j->event);    // blocking-sync event: this thread sleeps
      const int n = static_cast<int>(j->paths.size());
      if (e != cudaSuccess) {
        j->status = HIAST_ERR_CUDA;
      } else {
        const int64_t total = j->offsets_host[n];
        if (total < 0 || static_cast<size_t>(total) > j->blob_host_capacity) {
          j->status = HIAST_ERR_WORKSPACE;
        } else if (static_cast<size_t>(total) > j->bytes_copied) {
          // the files were larger than the predicted copy: fetch the rest (the device blob is intact until the slot's
          // next window, which the caller only starts after waiting on this ticket)
          if (!topup_stream || topup_device != j->device) {
            if (cudaStreamCreateWithFlags(&topup_stream, cudaStreamNonBlocking) != cudaSuccess) topup_stream = nullptr;
            topup_device = j->device;
          }
          e = cudaMemcpyAsync(const_cast<uint8_t*>(j->blob_host) + j->bytes_copied, j->blob_dev + j->bytes_copied,
                              static_cast<size_t>(total) - j->bytes_copied, cudaMemcpyDeviceToHost, topup_stream);
          if (e == cudaSuccess) e = cudaStreamSynchronize(topup_stream);
          if (e != cudaSuccess) j->status = HIAST_ERR_CUDA;
        }
      }
      std::unique_lock<std::mutex> lock(mu);
      pending.pop_front();
      if (j->status != HIAST_OK || n == 0) {
        writing.push_back(j);
        finish(j);
        continue;
      }
      j->left.store(n);
      writing.push_back(j);
      cv_files.notify_all();
    }
  }

  void worker_loop() {
    for (;;) {
      std::shared_ptr<WriteJob> j;
      int i = -1;
      {
        std::unique_lock<std::mutex> lock(mu);
        for (;;) {
          for (auto& w : writing) {
            const int n = static_cast<int>(w->paths.size());
            if (w->status == HIAST_OK && w->next.load() < n) {
              const int k = w->next.fetch_add(1);
              if (k < n) {
                j = w;
                i = k;
                break;
              }
            }
          }
          if (j || stop) break;
          cv_files.wait(lock);
        }
        if (!j) return;
      }
      write_one(*j, i);
      if (j->left.fetch_sub(1) == 1) {
        std::unique_lock<std::mutex> lock(mu);
        finish(j);
      }
    }
  }
};

}  // namespace

extern "C" int hiast_writer_create(int n_threads, void** handle_out) {
  if (n_threads < 1 || n_threads > 256 || !handle_out) return HIAST_ERR_INVALID_ARG;
  Writer* w = new Writer;
  w->dispatcher = std::thread([w] { w->dispatch_loop(); });
  for (int i = 0; i < n_threads; ++i) w->workers.emplace_back([w] { w->worker_loop(); });
  *handle_out = w;
  return HIAST_OK;
}

extern "C" int64_t hiast_writer_submit(void* handle, const char* const* paths_host, int n_files, const uint8_t* blob_host,
                                       size_t blob_host_capacity, const int64_t* offsets_host, size_t bytes_copied,
                                       const uint8_t* blob_dev, void* stream) {
  Writer* w = static_cast<Writer*>(handle);
  if (!w || n_files < 0 || (n_files && (!paths_host || !blob_host || !offsets_host))) return HIAST_ERR_INVALID_ARG;
  auto j = std::make_shared<WriteJob>();
  HIAST_CUDA_TRY(cudaGetDevice(&j->device));
  HIAST_CUDA_TRY(cudaEventCreateWithFlags(&j->event, cudaEventDisableTiming | cudaEventBlockingSync));
  cudaError_t e = cudaEventRecord(j->event, as_stream(stream));
  if (e != cudaSuccess) {
    cudaEventDestroy(j->event);
    return cuda_fail(e);
  }
  j->paths.reserve(n_files);
  for (int i = 0; i < n_files; ++i) j->paths.emplace_back(paths_host[i]);
  j->blob_host = blob_host;
  j->blob_host_capacity = blob_host_capacity;
  j->offsets_host = offsets_host;
  j->bytes_copied = bytes_copied;
  j->blob_dev = blob_dev;
  std::unique_lock<std::mutex> lock(w->mu);
  j->ticket = w->next_ticket++;
  w->pending.push_back(j);
  w->cv_jobs.notify_one();
  return j->ticket;
}

extern "C" int hiast_writer_wait(void* handle, int64_t ticket, int* errno_out) {
  Writer* w = static_cast<Writer*>(handle);
  if (!w || ticket < 0) return HIAST_ERR_INVALID_ARG;
  std::unique_lock<std::mutex> lock(w->mu);
  if (ticket >= w->next_ticket) return HIAST_ERR_INVALID_ARG;
  w->cv_done.wait(lock, [&] { return w->done_upto >= ticket; });
  int status = HIAST_OK, err = 0;
  for (auto& f : w->finished)
    if (f.first <= ticket && f.second.first != HIAST_OK && status == HIAST_OK) {
      status = f.second.first;
      err = f.second.second;
    }
  if (errno_out) *errno_out = err;
  return status;
}

extern "C" int hiast_writer_destroy(void* handle) {
  Writer* w = static_cast<Writer*>(handle);
  if (!w) return HIAST_OK;
  {
    std::unique_lock<std::mutex> lock(w->mu);
    w->cv_done.wait(lock, [&] { return w->pending.empty() && w->writing.empty(); });
    w->stop = true;
    w->cv_jobs.notify_all();
    w->cv_files.notify_all();
  }
  w->dispatcher.join();
  for (auto& t : w->workers) t.join();
  if (w->topup_stream) cudaStreamDestroy(w->topup_stream);
  delete w;
  return HIAST_OK;
}
