// Host-side batched reader of basecaller output files (no GPU work here): the command line's loader threads spent
// their time holding Python's interpreter lock (open / read / header parsing per 100 KB file), so at a few thousand
// pairs per second the files, not the kernels, set the pace.  These two entry points do the per-file work on native
// threads; the logarithm of the probabilities stays numpy's (decode.py:45; numpy's float32 log is not a fixed function
// across builds, so only the same ufunc on the same host reproduces the reference bit for bit).
//
// replaces: the np.load calls of decode.load_logits (decode.py:41-51) and bonito's branch of model_from_trace
//           (decode.py:76-80) for plain .npy tables.
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/poreover_b200.h"

namespace {

struct NpyInfo {
  int64_t rows = 0, cols = 0;
  int64_t data_off = 0;
  int flag = POB_NPY_OTHER;
};

// Header of a version 1-3 .npy file holding a little-endian float32 C-order array of one or two dimensions.
bool parse_header(const char* buf, size_t len, NpyInfo& info) {
  if (len < 12 || memcmp(buf, "\x93NUMPY", 6) != 0) return false;
  const int major = (unsigned char)buf[6];
  size_t hlen, start;
  if (major == 1) { hlen = (unsigned char)buf[8] | ((size_t)(unsigned char)buf[9] << 8); start = 10; }
  else if (major == 2 || major == 3) {
    hlen = (unsigned char)buf[8] | ((size_t)(unsigned char)buf[9] << 8) | ((size_t)(unsigned char)buf[10] << 16) |
           ((size_t)(unsigned char)buf[11] << 24);
    start = 12;
  } else return false;
  if (start + hlen > len) return false;
  const std::string h(buf + start, hlen);
  if (h.find("'descr': '<f4'") == std::string::npos) return false;
  if (h.find("'fortran_order': False") == std::string::npos) return false;
  const size_t sp = h.find("'shape': (");
  if (sp == std::string::npos) return false;
  const size_t se = h.find(')', sp);
  if (se == std::string::npos) return false;
  int64_t dims[4];
  int nd = 0;
  const char* p = h.c_str() + sp + 10;
  const char* end = h.c_str() + se;
  while (p < end && nd < 4) {
    while (p < end && (*p == ' ' || *p == ',')) ++p;
    if (p >= end) break;
    char* q;
    const long long v = strtoll(p, &q, 10);
    if (q == p) return false;
    dims[nd++] = v;
    p = q;
  }
  if (nd != 2) return false;  // 3-D logits (PoreOverNet windows) and vectors take the reference loader path
  info.rows = dims[0]; info.cols = dims[1];
  info.data_off = (int64_t)(start + hlen);
  return true;
}

template <typename F>
void parallel_for(int n, int threads, F&& body) {
  if (threads < 1) threads = 1;
  if (threads > n) threads = n > 0 ? n : 1;
  std::atomic<int> next(0);
  auto loop = [&] {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n) return;
      body(i);
    }
  };
  if (threads == 1) { loop(); return; }
  std::vector<std::thread> ts;
  for (int t = 1; t < threads; ++t) ts.emplace_back(loop);
  loop();
  for (auto& t : ts) t.join();
}

bool read_all(int fd, char* dst, size_t bytes, off_t off) {
  size_t done = 0;
  while (done < bytes) {
    const ssize_t r = pread(fd, dst + done, bytes - done, off + (off_t)done);
    if (r < 0) { if (errno == EINTR) continue; return false; }
    if (r == 0) return false;
    done += (size_t)r;
  }
  return true;
}

}  // namespace

extern "C" {

int pob_npy_probe(const char* const* paths, int n, int threads, int64_t* rows, int64_t* cols, int64_t* data_off,
                  int32_t* flags, float* first_row_sum) {
  if (n < 0 || (n > 0 && (!paths || !rows || !cols || !data_off || !flags))) return POB_EINVAL;
  parallel_for(n, threads, [&](int i) {
    rows[i] = cols[i] = data_off[i] = 0;
    flags[i] = POB_NPY_UNREADABLE;
    if (first_row_sum) first_row_sum[i] = 0.f;
    const int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
    if (fd < 0) return;
    char buf[1024];
    const ssize_t got = pread(fd, buf, sizeof(buf), 0);
    NpyInfo info;
    if (got >= 12 && parse_header(buf, (size_t)got, info)) {
      struct stat st;
      if (fstat(fd, &st) == 0 && st.st_size >= info.data_off + info.rows * info.cols * 4) {
        rows[i] = info.rows; cols[i] = info.cols; data_off[i] = info.data_off;
        flags[i] = POB_NPY_F32_2D;
        if (first_row_sum && info.rows > 0 && info.cols > 0 && info.cols <= 64) {
          float row[64];
          if (read_all(fd, (char*)row, (size_t)info.cols * 4, (off_t)info.data_off)) {
            // numpy's float32 sum of fewer than 8 elements is the plain left-to-right loop (pairwise_sum's base case)
            float s = row[0];
            for (int k = 1; k < info.cols; ++k) s += row[k];
            first_row_sum[i] = s;
          } else flags[i] = POB_NPY_UNREADABLE;
        }
      }
    } else if (got >= 0) {
      flags[i] = POB_NPY_OTHER;
    }
    close(fd);
  });
  return POB_OK;
}

int pob_npy_read(const char* const* paths, int n, int threads, const int64_t* rows, int64_t cols, const int64_t* data_off,
                 const int64_t* row_off, float* dst, int32_t* ok) {
  if (n < 0 || (n > 0 && (!paths || !rows || !data_off || !row_off || !dst))) return POB_EINVAL;
  parallel_for(n, threads, [&](int i) {
    if (ok) ok[i] = 0;
    const int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
    if (fd < 0) return;
    float* out = dst + (size_t)row_off[i] * (size_t)cols;
    const size_t bytes = (size_t)rows[i] * (size_t)cols * 4;
    const bool good = read_all(fd, (char*)out, bytes, (off_t)data_off[i]);
    close(fd);
    // alignment rows between this read and the next
    const size_t pad_rows = (size_t)(row_off[i + 1] - row_off[i]) - (size_t)rows[i];
    if (pad_rows) memset(out + (size_t)rows[i] * cols, 0, pad_rows * (size_t)cols * 4);
    if (ok) ok[i] = good ? 1 : 0;
  });
  return POB_OK;
}

}  // extern "C"
