// Library plumbing: error reporting, launch counter, device-wide scan.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "scan.cuh"

namespace fd {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int exclusive_scan_i32(const int32_t* d_in, int32_t* d_out, int64_t n, int32_t* d_total, void* d_tmp,
                       cudaStream_t stream) {
  return scan_impl(LoadI32{d_in}, d_out, n, d_total, d_tmp, stream);
}
int exclusive_scan_popc(const uint32_t* d_in, int32_t* d_out, int64_t n, int32_t* d_total,
                        void* d_tmp, cudaStream_t stream) {
  return scan_impl(LoadPopc{d_in}, d_out, n, d_total, d_tmp, stream);
}

__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

}  // namespace fd

extern "C" {

int fd_version(void) { return FD_ABI_VERSION; }
const char* fd_last_error(void) { return fd::g_err; }
int64_t fd_launch_count(void) { return fd::g_launches.load(); }

size_t fd_scan_tmp_bytes(int64_t n) {
  if (n < 0) n = 0;
  return (size_t)(fd::ceil_div(n, fd::kScanTile) + 1) * sizeof(int32_t);
}

int fd_fill_i32(int32_t* d_ptr, int64_t n, int32_t value, void* stream) {
  FD_REQUIRE(d_ptr != nullptr || n == 0, "fd_fill_i32: null pointer");
  if (n <= 0) return 0;
  int grid = fd::persistent_grid(fd::ceil_div(n, 256), 8);
  fd::fill_i32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_ptr, n, value);
  FD_LAUNCHED();
  return 0;
}

}  // extern "C"
