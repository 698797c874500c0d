// Ensemble statistics, batched FIRE and the fused PaiNN relaxation driver.
// Replaces: EnsembleNFF mean/std + unit/offset handling (mcmc/calculators/calculators.py:468-489
// via nff EnsembleNFF.calculate), ase.optimize.FIRE.step + Dynamics.irun + FixAtoms as driven by
// optimize_slab (mcmc/dynamics.py:120-168), get_system_val (mcmc/uncertainty/prediction.py:181-223).
// CPU restatement: oracle/relax.py, oracle/painn.py::EnsembleOracle.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr double KCAL_PER_EV = 23.06052;  // nff.utils.constants.EV_TO_KCAL_MOL

__global__ void ensemble_energy_kernel(const double* __restrict__ energy, const double* __restrict__ offset_ev,
                                       int M, int B, double* __restrict__ e_mean, double* __restrict__ e_std) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double off = offset_ev ? offset_ev[b] : 0.0;
  double s = 0.0;
  for (int m = 0; m < M; ++m) s += energy[(long long)m * B + b] * (1.0 / KCAL_PER_EV) + off;
  const double mean = s / M;
  double v = 0.0;
  for (int m = 0; m < M; ++m) {
    const double d = (energy[(long long)m * B + b] * (1.0 / KCAL_PER_EV) + off) - mean;
    v += d * d;
  }
  e_mean[b] = mean;
  e_std[b] = sqrt(v / M);
}

// fp32 like numpy on the reference's float32 arrays: g_m*(1/23.06052), sequential sum, / M
__global__ void ensemble_force_kernel(const float* __restrict__ grad, int M, long long n3, float* __restrict__ f_mean,
                                      float* __restrict__ f_std) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n3) return;
  const float c = (float)(1.0 / KCAL_PER_EV);
  float s = 0.f;
  for (int m = 0; m < M; ++m) s = __fadd_rn(s, __fmul_rn(grad[(long long)m * n3 + k], c));
  const float mean = s / (float)M;
  f_mean[k] = -mean;
  if (f_std) {
    float v = 0.f;
    for (int m = 0; m < M; ++m) {
      const float d = __fsub_rn(__fmul_rn(grad[(long long)m * n3 + k], c), mean);
      v = __fadd_rn(v, __fmul_rn(d, d));
    }
    f_std[k] = sqrtf(v / (float)M);
  }
}

__global__ void atom_norm_kernel(const float* __restrict__ vec, int n, float* __restrict__ out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const float x = vec[3 * a], y = vec[3 * a + 1], z = vec[3 * a + 2];
  out[a] = sqrtf(x * x + y * y + z * z);
}

// one warp per structure: sum, max, min, mean, mean_squared, rms
__global__ void __launch_bounds__(128) system_reduce_kernel(const float* __restrict__ x, const int32_t* __restrict__ atom_ptr,
                                                            int B, float* __restrict__ out) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  const int a0 = atom_ptr[b], a1 = atom_ptr[b + 1];
  float s = 0.f, s2 = 0.f, mx = -INFINITY, mn = INFINITY;
  for (int a = a0 + lane; a < a1; a += 32) {
    const float v = x[a];
    s += v; s2 += v * v; mx = fmaxf(mx, v); mn = fminf(mn, v);
  }
  s = warp_sum(s); s2 = warp_sum(s2); mx = warp_max(mx); mn = warp_min(mn);
  if (lane == 0) {
    const float n = (float)(a1 - a0);
    float* o = out + 6 * b;
    o[0] = s; o[1] = mx; o[2] = mn; o[3] = s / n; o[4] = s2 / n; o[5] = sqrtf(s2 / n);
  }
}

__global__ void fire_init_kernel(double* __restrict__ state, double* __restrict__ vel, int B, long long n3) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n3) vel[k] = 0.0;
  if (k < B) {
    double* s = state + 8 * k;
    s[0] = 0.1; s[1] = 0.1; s[2] = 0.0; s[3] = 0.0; s[4] = 0.0; s[5] = 0.0; s[6] = 0.0; s[7] = 0.0;
  }
}

// fixed-order block reduction (128 threads) of up to 3 doubles
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double (*red)[3]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  __syncthreads();
  if (lane == 0) { red[wid][0] = a; red[wid][1] = b; red[wid][2] = c; }
  __syncthreads();
  a = (red[0][0] + red[1][0]) + (red[2][0] + red[3][0]);
  b = (red[0][1] + red[1][1]) + (red[2][1] + red[3][1]);
  c = (red[0][2] + red[1][2]) + (red[2][2] + red[3][2]);
}
__device__ __forceinline__ void block_max2(double& a, double& b, double (*red)[3]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  a = warp_max(a); b = warp_max(b);
  __syncthreads();
  if (lane == 0) { red[wid][0] = a; red[wid][1] = b; }
  __syncthreads();
  a = fmax(fmax(red[0][0], red[1][0]), fmax(red[2][0], red[3][0]));
  b = fmax(fmax(red[0][1], red[1][1]), fmax(red[2][1], red[3][1]));
}

// One CTA per structure.  ASE FIRE (no masses, whole-structure norms) in fp64.
__global__ void __launch_bounds__(128) fire_step_kernel(double* __restrict__ pos, float* __restrict__ pos32,
                                                        double* __restrict__ vel, const float* __restrict__ forces,
                                                        const uint8_t* __restrict__ fixed,
                                                        const int32_t* __restrict__ atom_ptr, double* __restrict__ state,
                                                        int max_steps, double fmax_tol) {
  __shared__ double red[4][3];
  const int b = blockIdx.x;
  double* s = state + 8 * b;
  if (s[4] != 0.0) return;  // already converged: ASE stopped calling the calculator
  const int a0 = atom_ptr[b], a1 = atom_ptr[b + 1];
  const int tid = threadIdx.x;
  // masked force statistics
  double fm2 = 0.0, fabsmax = 0.0;
  for (int a = a0 + tid; a < a1; a += blockDim.x) {
    const double fx = forces[3 * a], fy = forces[3 * a + 1], fz = forces[3 * a + 2];
    fabsmax = fmax(fabsmax, fmax(fabs(fx), fmax(fabs(fy), fabs(fz))));
    if (!fixed[a]) fm2 = fmax(fm2, fx * fx + fy * fy + fz * fz);
  }
  block_max2(fm2, fabsmax, red);
  const int nsteps = (int)s[3];
  const bool conv = fm2 < fmax_tol * fmax_tol;
  __syncthreads();
  if (tid == 0) {
    s[6] = sqrt(fm2);
    s[7] = fabsmax;
    if (conv) s[4] = 1.0;
  }
  if (conv || nsteps >= max_steps) return;

  double dt = s[0], alpha = s[1];
  int npos = (int)s[2];
  const bool has_v = s[5] != 0.0;
  if (has_v) {
    double vf = 0.0, ff = 0.0, vv = 0.0;
    for (int a = a0 + tid; a < a1; a += blockDim.x) {
      if (fixed[a]) continue;
      for (int c = 0; c < 3; ++c) {
        const double f = forces[3 * a + c], v = vel[3 * a + c];
        vf += f * v; ff += f * f; vv += v * v;
      }
    }
    block_sum3(vf, ff, vv, red);
    if (vf > 0.0) {
      const double scale = alpha / sqrt(ff) * sqrt(vv);
      for (int a = a0 + tid; a < a1; a += blockDim.x) {
        for (int c = 0; c < 3; ++c) {
          const double f = fixed[a] ? 0.0 : (double)forces[3 * a + c];
          vel[3 * a + c] = (1.0 - alpha) * vel[3 * a + c] + f * scale;
        }
      }
      if (npos > 5) {
        dt = fmin(dt * 1.1, 1.0);
        alpha *= 0.99;
      }
      npos += 1;
    } else {
      for (int a = a0 + tid; a < a1; a += blockDim.x)
        for (int c = 0; c < 3; ++c) vel[3 * a + c] = 0.0;
      alpha = 0.1;
      dt *= 0.5;
      npos = 0;
    }
  }
  // v += dt f ; dr = dt v
  double n2 = 0.0, z0 = 0.0, z1 = 0.0;
  for (int a = a0 + tid; a < a1; a += blockDim.x) {
    for (int c = 0; c < 3; ++c) {
      const double f = fixed[a] ? 0.0 : (double)forces[3 * a + c];
      const double v = vel[3 * a + c] + dt * f;
      vel[3 * a + c] = v;
      const double dr = dt * v;
      n2 += dr * dr;
    }
  }
  block_sum3(n2, z0, z1, red);
  const double normdr = sqrt(n2);
  const double sc = normdr > 0.2 ? 0.2 / normdr : 1.0;
  for (int a = a0 + tid; a < a1; a += blockDim.x) {
    if (fixed[a]) continue;
    for (int c = 0; c < 3; ++c) {
      const double dr = dt * vel[3 * a + c];
      const double x = pos[3 * a + c] + (normdr > 0.2 ? 0.2 * dr / normdr : dr);
      pos[3 * a + c] = x;
      pos32[3 * a + c] = (float)x;
    }
  }
  (void)sc;
  if (tid == 0) {
    s[0] = dt; s[1] = alpha; s[2] = (double)npos; s[3] = (double)(nsteps + 1); s[5] = 1.0;
  }
}

__global__ void to_float_kernel(const double* __restrict__ src, long long n, float* __restrict__ dst) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[k] = (float)src[k];
}

// out[b] = [E (clamped), E_std, raw E, max|F|, nsteps, converged, oob, n_evals]
// max|F| is taken over ALL atoms of the last evaluation's raw forces (calc.results["forces"], mcmc/dynamics.py:155)
__global__ void relax_finalize_kernel(const double* __restrict__ e_mean, const double* __restrict__ e_std,
                                      const double* __restrict__ state, const float* __restrict__ forces,
                                      const int32_t* __restrict__ atom_ptr, int B, double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* s = state + 8 * b;
  const double e = e_mean[b];
  double fmx = 0.0;
  for (int k = 3 * atom_ptr[b]; k < 3 * atom_ptr[b + 1]; ++k) fmx = fmax(fmx, fabs((double)forces[k]));
  const bool oob = fabs(e) > 1000.0 || fmx > 1000.0;  // mcmc/dynamics.py:159
  double* o = out + 8 * b;
  o[0] = oob ? 1000.0 : e;
  o[1] = e_std ? e_std[b] : 0.0;
  o[2] = e;
  o[3] = fmx;
  o[4] = s[3];
  o[5] = s[4];
  o[6] = oob ? 1.0 : 0.0;
  o[7] = s[3] + 1.0;
}

struct RelaxWs {
  float* pos32; double* vel; double* state; int32_t* deg; int32_t* rowptr; int32_t* col; int8_t* shift;
  double* energy; float* grad; double* e_mean; double* e_std; void* painn; size_t painn_bytes; size_t bytes;
};

RelaxWs carve_relax(void* base, int M, int A, int Bmax, long long e_cap) {
  RelaxWs w;
  size_t off = 0;
  auto take = [&](size_t nbytes) -> void* {
    void* p = base ? reinterpret_cast<char*>(base) + off : nullptr;
    off += ((nbytes + 255) / 256) * 256;
    return p;
  };
  w.pos32 = (float*)take((size_t)A * 3 * 4);
  w.vel = (double*)take((size_t)A * 3 * 8);
  w.state = (double*)take((size_t)Bmax * 8 * 8);
  w.deg = (int32_t*)take((size_t)A * 4);
  w.rowptr = (int32_t*)take((size_t)(A + 1) * 4);
  w.col = (int32_t*)take((size_t)e_cap * 4);
  w.shift = (int8_t*)take((size_t)e_cap * 4);
  w.energy = (double*)take((size_t)M * Bmax * 8);
  w.grad = (float*)take((size_t)M * A * 3 * 4);
  w.e_mean = (double*)take((size_t)Bmax * 8);
  w.e_std = (double*)take((size_t)Bmax * 8);
  w.painn_bytes = vssr_painn_workspace_bytes(M, A, e_cap);
  w.painn = take(w.painn_bytes);
  w.bytes = off;
  return w;
}

}  // namespace

extern "C" int vssr_ensemble_stats(const double* energy, const float* grad, const double* offset_ev,
                                   const int32_t* atom_ptr, int32_t n_models, int32_t n_struct, int32_t n_atoms,
                                   double* e_mean, double* e_std, float* f_mean, float* f_std, void* stream) {
  (void)atom_ptr;
  if (!energy || !grad || !e_mean || !e_std || !f_mean) return VSSR_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  VSSR_PROF(VSSR_K_ENSEMBLE, st, ensemble_energy_kernel<<<ceil_div(n_struct, 128), 128, 0, st>>>(energy, offset_ev, n_models, n_struct, e_mean, e_std));
  const long long n3 = 3LL * n_atoms;
  VSSR_PROF(VSSR_K_ENSEMBLE, st, ensemble_force_kernel<<<ceil_div(n3, 256), 256, 0, st>>>(grad, n_models, n3, f_mean, f_std));
  return VSSR_OK;
}

extern "C" int vssr_system_reduce(const float* per_atom, const int32_t* atom_ptr, int32_t n_struct, float* out,
                                  void* stream) {
  if (!per_atom || !atom_ptr || !out) return VSSR_ERR_ARG;
  VSSR_PROF(VSSR_K_ENSEMBLE, (cudaStream_t)stream, system_reduce_kernel<<<ceil_div(n_struct, 4), 128, 0, (cudaStream_t)stream>>>(per_atom, atom_ptr, n_struct, out));
  return VSSR_OK;
}

extern "C" int vssr_atom_norm(const float* vec, int32_t n_atoms, float* out, void* stream) {
  if (!vec || !out) return VSSR_ERR_ARG;
  VSSR_PROF(VSSR_K_ENSEMBLE, (cudaStream_t)stream, atom_norm_kernel<<<ceil_div(n_atoms, 256), 256, 0, (cudaStream_t)stream>>>(vec, n_atoms, out));
  return VSSR_OK;
}

extern "C" int vssr_fire_init(double* fire_state, double* vel, int32_t n_struct, int32_t n_atoms, void* stream) {
  if (!fire_state || !vel) return VSSR_ERR_ARG;
  const long long n3 = 3LL * n_atoms;
  const long long n = n3 > n_struct ? n3 : n_struct;
  VSSR_PROF(VSSR_K_FIRE, (cudaStream_t)stream, fire_init_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(fire_state, vel, n_struct, n3));
  return VSSR_OK;
}

extern "C" int vssr_fire_step(double* pos, float* pos32, double* vel, const float* forces, const uint8_t* fixed,
                              const int32_t* atom_ptr, int32_t n_struct, double* fire_state, int32_t max_steps,
                              double fmax, void* stream) {
  if (!pos || !pos32 || !vel || !forces || !fixed || !atom_ptr || !fire_state) return VSSR_ERR_ARG;
  VSSR_PROF(VSSR_K_FIRE, (cudaStream_t)stream, fire_step_kernel<<<n_struct, 128, 0, (cudaStream_t)stream>>>(pos, pos32, vel, forces, fixed, atom_ptr, fire_state,
                                                               max_steps, fmax));
  return VSSR_OK;
}

extern "C" size_t vssr_painn_relax_workspace_bytes(int32_t n_models, int32_t n_atoms, int64_t e_cap) {
  // the structure count is bounded by the atom count
  return carve_relax(nullptr, n_models, n_atoms, n_atoms, e_cap).bytes;
}

extern "C" int vssr_painn_relax_edge_stats(const void* relax_workspace, int32_t n_models, int32_t n_atoms, int64_t e_cap,
                                           const int32_t* atom_ptr, int32_t n_struct, const void* filter_cache,
                                           int32_t fc_n0, int64_t fc_e_cap0, int64_t* out5, void* stream) {
  if (!relax_workspace) return VSSR_ERR_ARG;
  const RelaxWs w = carve_relax(const_cast<void*>(relax_workspace), n_models, n_atoms, n_atoms, e_cap);
  return vssr_painn_edge_stats(w.painn, n_models, n_atoms, e_cap, atom_ptr, n_struct, w.rowptr, filter_cache, fc_n0,
                               fc_e_cap0, out5, stream);
}

namespace {
// Executable graphs of relaxations still in flight: destroyed once the event recorded after their last launch has
// completed (polled at the next call; the oldest is waited for when 32 are pending).
struct GraphPool {
  struct Item { cudaGraphExec_t exec; cudaEvent_t done; };
  cudaStream_t capture_stream = nullptr;
  Item items[32];
  int n = 0;
  int reap(bool all) {
    int kept = 0;
    for (int k = 0; k < n; ++k) {
      cudaError_t q = all ? cudaEventSynchronize(items[k].done) : cudaEventQuery(items[k].done);
      if (q == cudaSuccess) {
        cudaGraphExecDestroy(items[k].exec);
        cudaEventDestroy(items[k].done);
      } else if (q == cudaErrorNotReady) {
        items[kept++] = items[k];
      } else {
        return (int)q;
      }
    }
    n = kept;
    return VSSR_OK;
  }
  int retire(cudaGraphExec_t exec, cudaStream_t st) {
    if (n == 32) {
      const int r = reap(true);
      if (r) return r;
    }
    Item it{exec, nullptr};
    VSSR_CUDA(cudaEventCreateWithFlags(&it.done, cudaEventDisableTiming));
    VSSR_CUDA(cudaEventRecord(it.done, st));
    items[n++] = it;
    return VSSR_OK;
  }
};
GraphPool& graph_pool() {       // per thread AND per device: the capture stream lives on one device
  static thread_local GraphPool pools[64];
  int dev = 0;
  cudaGetDevice(&dev);
  return pools[dev & 63];
}
}  // namespace

extern "C" int vssr_painn_relax(const float* weights, int32_t n_models, double* pos, const int32_t* z,
                                const uint8_t* fixed, const int32_t* atom_ptr, const float* cell, const uint8_t* pbc,
                                const double* offset_ev, int32_t n_struct, int32_t n_atoms,
                                int32_t max_atoms_per_struct, float cutoff, float skin,
                                int32_t relax_steps, double fmax, int64_t e_cap, const void* filter_cache,
                                int32_t fc_n0, int64_t fc_e_cap0, int32_t fc_flags, void* workspace,
                                size_t workspace_bytes,
                                double* out, float* forces, float* forces_std,
                                int32_t* status, void* stream) {
  if (!weights || !pos || !z || !fixed || !atom_ptr || !cell || !pbc || !workspace || !out || !forces || !status)
    return VSSR_ERR_ARG;
  if (n_struct <= 0 || n_atoms <= 0 || n_struct > n_atoms || relax_steps < 0) return VSSR_ERR_ARG;
  RelaxWs w = carve_relax(workspace, n_models, n_atoms, n_atoms, e_cap);
  if (w.bytes > workspace_bytes) return VSSR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  const long long n3 = 3LL * n_atoms;
  VSSR_PROF(VSSR_K_FIRE, st, to_float_kernel<<<ceil_div(n3, 256), 256, 0, st>>>(pos, n3, w.pos32));
  if ((rc = vssr_nbr_build(w.pos32, atom_ptr, cell, pbc, n_struct, n_atoms, cutoff + skin, w.deg, w.rowptr, w.col,
                           w.shift, e_cap, status, stream)))
    return rc;
  if ((rc = vssr_fire_init(w.state, w.vel, n_struct, n_atoms, stream))) return rc;
  // one optimiser iteration: ensemble evaluation, statistics, FIRE step.  Iterations 0 .. relax_steps-1 are the SAME
  // launches with the same arguments (the step counter lives in w.state); the last evaluation is the one whose raw
  // forces on ALL atoms feed the out-of-bounds test (mcmc/dynamics.py:155-159): it always computes the full gradient
  auto iteration = [&](bool last, void* s) -> int {
    const int32_t flags_it = last ? (fc_flags & ~VSSR_FC_CONSTRAINED_GRAD) : fc_flags;
    int r;
    if ((r = vssr_painn_energy_grad(weights, n_models, w.pos32, z, atom_ptr, cell, n_struct, n_atoms,
                                    max_atoms_per_struct, w.rowptr, w.col, w.shift, e_cap, cutoff, filter_cache, fc_n0,
                                    fc_e_cap0, flags_it, w.painn, w.painn_bytes, w.energy, w.grad, nullptr, s)))
      return r;
    if ((r = vssr_ensemble_stats(w.energy, w.grad, offset_ev, atom_ptr, n_models, n_struct, n_atoms, w.e_mean,
                                 w.e_std, forces, last ? forces_std : nullptr, s)))
      return r;
    return vssr_fire_step(pos, w.pos32, w.vel, forces, fixed, atom_ptr, n_struct, w.state, relax_steps, fmax, s);
  };
  // Iterations 1 .. relax_steps-1 are replayed from a CUDA graph captured once per call (~75 kernel nodes): the host
  // issues relax_steps graph launches instead of ~1500 kernel launches and never fills the launch queue.  Iteration 0
  // runs directly (it configures function attributes and tensor maps -- nothing of that happens inside the capture).
  // Off while the per-class event profile is recording, and with VSSR_NO_GRAPH=1 (bench documentation switch).
  static const bool no_graph = getenv("VSSR_NO_GRAPH") && atoi(getenv("VSSR_NO_GRAPH")) > 0;
  if (relax_steps < 3 || no_graph || vssr_prof_active()) {
    for (int it = 0; it <= relax_steps; ++it)
      if ((rc = iteration(it == relax_steps, stream))) return rc;
  } else {
    if ((rc = iteration(false, stream))) return rc;
    GraphPool& gp = graph_pool();
    if ((rc = gp.reap(false))) return rc;
    if (!gp.capture_stream) VSSR_CUDA(cudaStreamCreateWithFlags(&gp.capture_stream, cudaStreamNonBlocking));
    const long long l0 = g_vssr_launches;
    VSSR_CUDA(cudaStreamBeginCapture(gp.capture_stream, cudaStreamCaptureModeThreadLocal));
    rc = iteration(false, gp.capture_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ec = cudaStreamEndCapture(gp.capture_stream, &graph);
    const long long per_iteration = g_vssr_launches - l0;   // kernel nodes of one iteration
    g_vssr_launches = l0;
    if (rc || ec != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      return rc ? rc : (int)ec;
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) return (int)ei;
    for (int it = 1; it < relax_steps; ++it) {
      const cudaError_t el = cudaGraphLaunch(exec, st);
      if (el != cudaSuccess) { cudaGraphExecDestroy(exec); return (int)el; }
      g_vssr_launches += per_iteration;
      ++g_vssr_graph_launches;
    }
    if ((rc = gp.retire(exec, st))) return rc;
    if ((rc = iteration(true, stream))) return rc;
  }
  VSSR_PROF(VSSR_K_FIRE, st, relax_finalize_kernel<<<ceil_div(n_struct, 128), 128, 0, st>>>(w.e_mean, w.e_std, w.state, forces, atom_ptr, n_struct, out));
  return VSSR_OK;
}
