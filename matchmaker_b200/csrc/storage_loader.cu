// Block loader of the dense-retrieval storage layout: file segments -> one contiguous device buffer.
//
// The reference keeps the encoded collection in numpy memmaps `token_reps_<n>.npy`
// (matchmaker/dense_retrieval.py:201-265, re-opened at :291-302) and hands them to the indexer as host arrays
// (:328), i.e. every row crosses  disk -> page cache -> (numpy copy) -> pageable cudaMemcpy.  Here the rows a rank
// owns are read with pread() straight into two pinned staging buffers and leave for the GPU with cudaMemcpyAsync:
// the read of segment piece i+1 overlaps the PCIe transfer of piece i, no Python objects and no pageable bounce
// buffer are involved, and only the byte ranges the rank needs are touched.  Host code only (no kernel).
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstring>
#include <mutex>

#include "host_util.cuh"

namespace mmb {

namespace {

struct Staging {
  void* buf[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  size_t bytes = 0;
};

std::mutex g_stage_mu;
Staging g_stage;

int ensure_staging(size_t bytes) {
  if (g_stage.bytes >= bytes) return MMB200_OK;
  for (int i = 0; i < 2; ++i) {
    if (g_stage.buf[i]) MMB_CHECK_CUDA(cudaFreeHost(g_stage.buf[i]));
    g_stage.buf[i] = nullptr;
  }
  g_stage.bytes = 0;
  for (int i = 0; i < 2; ++i) {
    MMB_CHECK_CUDA(cudaHostAlloc(&g_stage.buf[i], bytes, cudaHostAllocDefault));
    if (!g_stage.done[i]) MMB_CHECK_CUDA(cudaEventCreateWithFlags(&g_stage.done[i], cudaEventDisableTiming));
  }
  g_stage.bytes = bytes;
  return MMB200_OK;
}

}  // namespace

}  // namespace mmb

extern "C" int mmb200_storage_load(const char* const* paths, const int64_t* file_offsets, const int64_t* nbytes,
                                   int32_t n_segments, void* dst_device, int64_t staging_bytes, void* stream_) {
  using namespace mmb;
  MMB_REQUIRE(n_segments >= 0, "negative segment count");
  if (n_segments == 0) return MMB200_OK;
  MMB_REQUIRE(paths && file_offsets && nbytes && dst_device, "null pointer");
  if (staging_bytes <= 0) staging_bytes = 32ll << 20;
  staging_bytes = std::max<int64_t>(staging_bytes, 4096);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  std::lock_guard<std::mutex> lock(g_stage_mu);
  if (int rc = ensure_staging((size_t)staging_bytes)) return rc;
  uint8_t* dst = static_cast<uint8_t*>(dst_device);
  int slot = 0;
  bool used[2] = {false, false};
  for (int s = 0; s < n_segments; ++s) {
    MMB_REQUIRE(paths[s] && file_offsets[s] >= 0 && nbytes[s] >= 0, "bad segment");
    if (nbytes[s] == 0) continue;
    const int fd = open(paths[s], O_RDONLY);
    if (fd < 0) {
      set_error(std::string("mmb200_storage_load: cannot open ") + paths[s] + ": " + strerror(errno));
      return MMB200_ERR_INVALID;
    }
    int64_t done = 0;
    while (done < nbytes[s]) {
      const int64_t want = std::min<int64_t>(staging_bytes, nbytes[s] - done);
      if (used[slot]) {  // the copy that last read this staging buffer must have finished
        cudaError_t e = cudaEventSynchronize(g_stage.done[slot]);
        if (e != cudaSuccess) {
          close(fd);
          set_error(std::string("cudaEventSynchronize failed: ") + cudaGetErrorString(e));
          return MMB200_ERR_CUDA;
        }
      }
      int64_t got = 0;
      while (got < want) {
        const ssize_t r = pread(fd, static_cast<uint8_t*>(g_stage.buf[slot]) + got, (size_t)(want - got),
                                (off_t)(file_offsets[s] + done + got));
        if (r <= 0) {
          close(fd);
          set_error(std::string("mmb200_storage_load: short read from ") + paths[s] +
                    (r < 0 ? std::string(": ") + strerror(errno) : std::string(" (file shorter than the segment)")));
          return MMB200_ERR_INVALID;
        }
        got += r;
      }
      cudaError_t e = cudaMemcpyAsync(dst, g_stage.buf[slot], (size_t)want, cudaMemcpyHostToDevice, stream);
      if (e == cudaSuccess) e = cudaEventRecord(g_stage.done[slot], stream);
      if (e != cudaSuccess) {
        close(fd);
        set_error(std::string("cudaMemcpyAsync failed: ") + cudaGetErrorString(e));
        return MMB200_ERR_CUDA;
      }
      used[slot] = true;
      slot ^= 1;
      dst += want;
      done += want;
    }
    close(fd);
  }
  // the staging buffers are reused by the next call: wait until the last copies have left them
  for (int i = 0; i < 2; ++i)
    if (used[i]) MMB_CHECK_CUDA(cudaEventSynchronize(g_stage.done[i]));
  return MMB200_OK;
}
