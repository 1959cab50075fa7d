// Error reporting, version and device checks of the C ABI.
#include "common.h"

#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

namespace avcer {

static thread_local char g_err[1024] = "";

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

int& sm_limit_ref() {
  static thread_local int limit = 0;
  return limit;
}

}  // namespace avcer

extern "C" int avcer_set_sm_limit(int n_sms) {
  if (n_sms < 0 || (n_sms & 1)) return avcer::set_error("set_sm_limit: %d must be 0 (all) or a positive even SM count", n_sms);
  avcer::sm_limit_ref() = n_sms;
  return 0;
}

extern "C" const char* avcer_last_error(void) { return avcer::g_err; }

extern "C" int avcer_version(void) { return 100; }

extern "C" const char* avcer_storage_type(void) { return AVCER_STORAGE_NAME; }

extern "C" int avcer_device_check(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return avcer::set_error("no CUDA device visible (%s); avcer_b200 has no CPU fallback",
                            e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  int dev = 0, major = 0;
  AVCER_CUDA(cudaGetDevice(&dev));
  AVCER_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10)
    return avcer::set_error("device compute capability %d.x is not sm_100a (B200); kernels are sm_100a only", major);
  return 0;
}

extern "C" int avcer_num_sms(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}

// ------------------------------------------------------------------ host-side file staging (no device work)
// Reads n files into one caller-owned host buffer (normally pinned), file i at dst + offsets[i] (offsets are written here:
// every file starts on a 16-byte boundary), sizes[i] = its length.  The reference reads its face crops one cv2.imread at a
// time (get_prob_video.py:95); from Python, open() + read() costs ~20 us per 30 KB file of interpreter and syscall overhead
// -- more than the GPU needs to decode it.  `threads` workers pull file indices from a shared counter: pass 1 stats the
// files (sizes -> offsets), pass 2 reads them.  Returns 0, or 1 with avcer_last_error() naming the file that failed or
// the capacity that was needed (*needed is set in both cases).
extern "C" int avcer_read_files(const char* const* paths, int n, uint8_t* dst, int64_t capacity, int64_t* offsets, int64_t* sizes,
                                int64_t* needed, int threads) {
  if (n < 0 || paths == nullptr || offsets == nullptr || sizes == nullptr || needed == nullptr)
    return avcer::set_error("read_files: bad arguments");
  if (threads < 1) threads = 1;
  if (threads > 64) threads = 64;
  std::atomic<int> next{0}, bad{-1};
  auto run = [&](auto&& body) {
    next = 0;
    std::vector<std::thread> pool;
    for (int t = 1; t < threads && t < n; ++t) pool.emplace_back(body);
    body();
    for (auto& th : pool) th.join();
  };
  run([&] {
    for (int i; (i = next.fetch_add(1)) < n;) {
      struct stat st;
      if (stat(paths[i], &st) != 0) { bad = i; sizes[i] = 0; } else sizes[i] = (int64_t)st.st_size;
    }
  });
  if (bad >= 0) return avcer::set_error("read_files: cannot stat %s", paths[bad.load()]);
  int64_t off = 0;
  for (int i = 0; i < n; ++i) {
    offsets[i] = off;
    off += (sizes[i] + 15) / 16 * 16;
  }
  *needed = off;
  if (off > capacity || dst == nullptr) return avcer::set_error("read_files: %lld bytes needed, buffer holds %lld", (long long)off, (long long)capacity);
  run([&] {
    for (int i; (i = next.fetch_add(1)) < n;) {
      const int fd = open(paths[i], O_RDONLY);
      if (fd < 0) { bad = i; continue; }
      int64_t got = 0;
      while (got < sizes[i]) {
        const ssize_t r = read(fd, dst + offsets[i] + got, (size_t)(sizes[i] - got));
        if (r <= 0) break;
        got += r;
      }
      close(fd);
      if (got != sizes[i]) bad = i;
    }
  });
  if (bad >= 0) return avcer::set_error("read_files: cannot read %s", paths[bad.load()]);
  return 0;
}
