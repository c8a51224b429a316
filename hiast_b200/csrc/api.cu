// Library-level entry points, device queries and host-side test hooks.
#include <fcntl.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"
#include "scan_math.h"

namespace hiast {

thread_local int g_last_cuda_error = 0;

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int ensure_dyn_smem_impl(const void* kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return HIAST_OK;           // the default limit needs no opt-in
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> granted;
  int dev = 0;
  HIAST_CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = granted[std::make_pair(dev, kernel)];
  if (have >= bytes) return HIAST_OK;
  HIAST_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  have = bytes;
  return HIAST_OK;
}

}  // namespace hiast

extern "C" int hiast_dev_variants(void) {
#ifdef HIAST_DEV_VARIANTS
  return 1;
#else
  return 0;
#endif
}

extern "C" int hiast_version(void) { return 1000 * 0 + 2; }

extern "C" const char* hiast_status_string(int status) {
  switch (status) {
    case HIAST_OK: return "ok";
    case HIAST_ERR_INVALID_ARG: return "invalid argument";
    case HIAST_ERR_UNSUPPORTED: return "unsupported configuration";
    case HIAST_ERR_CUDA: return "CUDA error (see hiast_last_cuda_error)";
    case HIAST_ERR_WORKSPACE: return "workspace too small";
    case HIAST_ERR_IO: return "file I/O error (errno in errno_out)";
    default: return "unknown status";
  }
}

extern "C" int hiast_last_cuda_error(void) { return hiast::g_last_cuda_error; }

extern "C" int hiast_device_sm_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return 0;
  }
  return hiast::sm_count();
}

extern "C" double hiast_testhook_powi(double x, int n) { return hiast::powi_dd(x, n); }

extern "C" double hiast_testhook_threshold_step(const uint32_t* prefix_row_host, int key_lo, double thr, double alpha,
                                                double beta, double gamma, float* temp_out, int* error_out) {
  const int nb = HIAST_KEY_ONE - key_lo + 1;
  float temp = 0.f;
  int err = 0;
  const double r = hiast::ias_threshold_step(prefix_row_host, nb, key_lo, thr, alpha, beta, gamma, &temp, &err);
  if (temp_out) *temp_out = temp;
  if (error_out) *error_out = err;
  return r;
}

// ---- host-side file writer of the device PNG path -------------------------------------------------------------------
// pseudo_label_generator.py:43-46 ends in a file per image.  With the files encoded on the device the remaining host work is
// open / write / close; done from Python threads it is throttled by the interpreter lock (every call boundary re-acquires it
// while the main thread is busy launching the next window).  One foreign call writes all files of a window with n_threads
// plain POSIX writers; ctypes releases the lock for the duration.
extern "C" int hiast_write_files(const char* const* paths, const uint8_t* blob, const int64_t* offsets, int n_files,
                                 int n_threads, int* errno_out) {
  if (!paths || !blob || !offsets || n_files < 0) return HIAST_ERR_INVALID_ARG;
  if (n_files == 0) return HIAST_OK;
  n_threads = std::max(1, std::min(n_threads, n_files));
  std::atomic<int> next(0), first_errno(0);
  auto worker = [&]() {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n_files) return;
      const int fd = ::open(paths[i], O_WRONLY | O_CREAT | O_TRUNC, 0644);
      int err = 0;
      if (fd < 0) {
        err = errno;
      } else {
        const uint8_t* p = blob + offsets[i];
        int64_t left = offsets[i + 1] - offsets[i];
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
        first_errno.compare_exchange_strong(expected, err);
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker);
  worker();
  for (auto& th : pool) th.join();
  if (errno_out) *errno_out = first_errno.load();
  return first_errno.load() == 0 ? HIAST_OK : HIAST_ERR_IO;
}
