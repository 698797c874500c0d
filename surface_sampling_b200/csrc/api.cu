// Library-level entry points of libvssr_b200.so (see include/vssr_b200.h).
#include "common.cuh"

long long g_vssr_launches = 0;

extern "C" int vssr_version(void) { return 100; }

extern "C" int64_t vssr_launch_count(void) { return (int64_t)g_vssr_launches; }

extern "C" int vssr_device_cc(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return -1;
  return p.major * 10 + p.minor;
}
