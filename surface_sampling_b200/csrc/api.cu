// Library-level entry points of libvssr_b200.so (see include/vssr_b200.h).
#include "common.cuh"

long long g_vssr_launches = 0;
long long g_vssr_graph_launches = 0;

extern "C" int vssr_version(void) { return 100; }

extern "C" int64_t vssr_launch_count(void) { return (int64_t)g_vssr_launches; }
extern "C" int64_t vssr_graph_launch_count(void) { return (int64_t)g_vssr_graph_launches; }

extern "C" int vssr_device_cc(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return -1;
  return p.major * 10 + p.minor;
}

// ---- optional per-kernel-class profile: cudaEvent pairs on the launching stream ----
namespace {
constexpr int kMaxPairs = 1 << 16;
bool g_prof_on = false;
int g_prof_n = 0;
cudaEvent_t* g_ev0 = nullptr;
cudaEvent_t* g_ev1 = nullptr;
int* g_cls = nullptr;
int g_created = 0;
}  // namespace

void vssr_prof_begin(int cls, cudaStream_t st) {
  if (!g_prof_on || g_prof_n >= kMaxPairs) return;
  if (g_prof_n >= g_created) {
    cudaEventCreate(&g_ev0[g_prof_n]);
    cudaEventCreate(&g_ev1[g_prof_n]);
    g_created = g_prof_n + 1;
  }
  g_cls[g_prof_n] = cls;
  cudaEventRecord(g_ev0[g_prof_n], st);
}
bool vssr_prof_active() { return g_prof_on; }
void vssr_prof_end(int cls, cudaStream_t st) {
  (void)cls;
  if (!g_prof_on || g_prof_n >= kMaxPairs) return;
  cudaEventRecord(g_ev1[g_prof_n], st);
  ++g_prof_n;
}

extern "C" int vssr_profile_enable(int on) {
  if (on && !g_ev0) {
    g_ev0 = new cudaEvent_t[kMaxPairs];
    g_ev1 = new cudaEvent_t[kMaxPairs];
    g_cls = new int[kMaxPairs];
  }
  g_prof_on = on != 0;
  g_prof_n = 0;
  return VSSR_OK;
}

// Synchronises the device, then sums elapsed ms and launch counts per kernel class.
extern "C" int vssr_profile_collect(double* ms /*[n_class]*/, int64_t* launches /*[n_class]*/, int n_class) {
  if (!ms || !launches || n_class < VSSR_K_NCLASS) return VSSR_ERR_ARG;
  VSSR_CUDA(cudaDeviceSynchronize());
  for (int k = 0; k < n_class; ++k) { ms[k] = 0.0; launches[k] = 0; }
  for (int i = 0; i < g_prof_n; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_ev0[i], g_ev1[i]) == cudaSuccess) {
      ms[g_cls[i]] += (double)t;
      launches[g_cls[i]] += 1;
    }
  }
  g_prof_n = 0;
  return VSSR_OK;
}

extern "C" int vssr_kernel_class_count(void) { return VSSR_K_NCLASS; }
