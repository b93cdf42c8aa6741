#include "common.cuh"

#include <stdarg.h>
#include <string.h>

namespace pccgeo {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

__global__ void finish_sum_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
  __shared__ double sm[32];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
  v = block_sum(v, sm);
  if (threadIdx.x == 0) *out = v;
}

}  // namespace pccgeo

extern "C" {
const char* pccgeo_last_error(void) { return pccgeo::g_err; }
int pccgeo_version(void) { return 100; }
long long pccgeo_launch_count(void) { return pccgeo::g_launches.load(); }
size_t pccgeo_reduce_ws_doubles(void) { return pccgeo::kReduceBlocks; }
}
