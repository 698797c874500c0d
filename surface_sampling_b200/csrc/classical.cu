// Tersoff and Stillinger-Weber many-body energy/forces + FIRE, one CTA per structure, the whole
// relaxation resident in shared memory (fp64).
// Replaces LAMMMPSCalc.run_lammps_energy / run_lammps_opt (mcmc/calculators/calculators.py:507-640,
// called from mcmc/dynamics.py:107-116 and calculators.py:685-688) for the GaN (pair_style tersoff)
// and Si (SW family) tutorials; functional forms restated in oracle/classical.py, FIRE in
// oracle/relax.py.
//
// Determinism: every output has exactly one writer.  Tersoff / SW: the directed pairs (centre i, neighbour j inside
// the cutoff) are compacted into a shared-memory pair table each step and a THREAD OWNS A PAIR -- it sums, in list
// order, every term of centre i's energy that pulls on neighbour j (pair term, the triplets where j is the bonded
// atom, the triplets where j is the third atom) into its private slot pG[pair]; EAM: thread i owns centre i and its
// slots G[i][slot].  A second phase lets every atom gather the slots that point at it through the reverse-edge map
// in fixed order.  No atomics anywhere.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int NT = 256;  // threads per CTA (512: Si +4 %, GaN -44 % -- one CTA per SM)
constexpr int NW = NT / 32;

struct Smem {
  double* x;      // [n_max][3] positions
  double* f;      // [n_max][3] forces (raw)
  double* v;      // [n_max][3] velocities
  double* x0;     // [n_max][3] positions at last list build
  double* G;      // [n_max][max_nbr][3]
  double* eat;    // [n_max]
  double* own;    // [n_max][3] centre-own gradient
  int* type;      // [n_max]
  int* cnt;       // [n_max]
  short* nj;      // [n_max][max_nbr] local neighbour index
  char4* ns;      // [n_max][max_nbr] shift
  unsigned char* rev;  // [n_max][max_nbr]
  unsigned char* fixed;  // [n_max]
  double* red;    // [NW] block reductions
  double* aux;    // [n_max] EAM: dF/drho of every atom between the two passes
  // Tersoff / SW: table of the directed pairs inside the cutoff, compacted per centre in skin-list order
  int* pcnt;             // [n_max] pairs of every centre
  int* pstart;           // [n_max + 1] first pair of every centre
  unsigned char* aidx;   // [n_max][max_nbr] skin slot -> index among the centre's pairs (255: outside the cutoff)
  short* pi;             // [pcap] centre
  unsigned char* pslot;  // [pcap] skin slot
  double* pu;            // [pcap][4] unit vector centre -> neighbour, distance
  double* pa;            // [pcap][2] SW: exp(gamma sigma / (r - a sigma)) and d/dr of its exponent
  double* pq;            // [pcap][3] energy share, dE/dr of the pair term, Tersoff: fA/2 db/dzeta | SW: 1 = inside cutoff
  double* pG;            // [pcap][3] gradient of the centre's energy w.r.t. the neighbour's position
  int pcap;
};
constexpr int PAIRS_PER_ATOM = 8;   // pair-table capacity per atom slot of the CTA (an average, not a per-atom limit)

__host__ __device__ inline size_t smem_layout(int kind, int n_max, int max_nbr, char* base, Smem* s) {
  const bool eam = kind == VSSR_POT_EAM;
  const int pcap = eam ? 0 : PAIRS_PER_ATOM * n_max;
  size_t off = 0;
  auto take = [&](size_t bytes) -> char* {
    char* p = base ? base + off : nullptr;
    off += (bytes + 15) / 16 * 16;
    return p;
  };
  double* x = (double*)take((size_t)n_max * 3 * 8);
  double* f = (double*)take((size_t)n_max * 3 * 8);
  double* v = (double*)take((size_t)n_max * 3 * 8);
  double* x0 = (double*)take((size_t)n_max * 3 * 8);
  double* G = (double*)take(eam ? (size_t)n_max * max_nbr * 3 * 8 : 0);
  double* eat = (double*)take((size_t)n_max * 8);
  double* own = (double*)take((size_t)n_max * 3 * 8);
  int* type = (int*)take((size_t)n_max * 4);
  int* cnt = (int*)take((size_t)n_max * 4);
  short* nj = (short*)take((size_t)n_max * max_nbr * 2);
  char4* ns = (char4*)take((size_t)n_max * max_nbr * 4);
  unsigned char* rev = (unsigned char*)take((size_t)n_max * max_nbr);
  unsigned char* fixed = (unsigned char*)take((size_t)n_max);
  double* red = (double*)take(16 * 8);
  double* aux = (double*)take(eam ? (size_t)n_max * 8 : 0);
  int* pcnt = (int*)take(eam ? 0 : (size_t)n_max * 4);
  int* pstart = (int*)take(eam ? 0 : (size_t)(n_max + 1) * 4);
  unsigned char* aidx = (unsigned char*)take(eam ? 0 : (size_t)n_max * max_nbr);
  short* pi = (short*)take((size_t)pcap * 2);
  unsigned char* pslot = (unsigned char*)take((size_t)pcap);
  double* pu = (double*)take((size_t)pcap * 4 * 8);
  double* pa = (double*)take((size_t)pcap * 2 * 8);
  double* pq = (double*)take((size_t)pcap * 3 * 8);
  double* pG = (double*)take((size_t)pcap * 3 * 8);
  if (s) {
    s->pcnt = pcnt; s->pstart = pstart; s->aidx = aidx; s->pi = pi; s->pslot = pslot; s->pu = pu; s->pa = pa; s->pq = pq;
    s->pG = pG; s->pcap = pcap;
    s->aux = aux;
    s->x = x; s->f = f; s->v = v; s->x0 = x0; s->G = G; s->eat = eat; s->own = own; s->type = type; s->cnt = cnt;
    s->nj = nj; s->ns = ns; s->rev = rev; s->fixed = fixed; s->red = red;
  }
  return off;
}

struct Cell64 {
  double c[9], inv[9], hinv[3];  // hinv = 1/perpendicular height
  bool pbc[3];
};

__device__ void load_cell64(const double* __restrict__ cell, const uint8_t* __restrict__ pbc, int b, Cell64& ci) {
  double* m = ci.c;
  for (int k = 0; k < 9; ++k) m[k] = cell[9 * b + k];
  double c0x = m[4] * m[8] - m[5] * m[7], c0y = m[5] * m[6] - m[3] * m[8], c0z = m[3] * m[7] - m[4] * m[6];
  double c1x = m[7] * m[2] - m[8] * m[1], c1y = m[8] * m[0] - m[6] * m[2], c1z = m[6] * m[1] - m[7] * m[0];
  double c2x = m[1] * m[5] - m[2] * m[4], c2y = m[2] * m[3] - m[0] * m[5], c2z = m[0] * m[4] - m[1] * m[3];
  double det = m[0] * c0x + m[1] * c0y + m[2] * c0z, idet = 1.0 / det;
  ci.inv[0] = c0x * idet; ci.inv[3] = c0y * idet; ci.inv[6] = c0z * idet;
  ci.inv[1] = c1x * idet; ci.inv[4] = c1y * idet; ci.inv[7] = c1z * idet;
  ci.inv[2] = c2x * idet; ci.inv[5] = c2y * idet; ci.inv[8] = c2z * idet;
  double vol = fabs(det);
  ci.hinv[0] = sqrt(c0x * c0x + c0y * c0y + c0z * c0z) / vol;
  ci.hinv[1] = sqrt(c1x * c1x + c1y * c1y + c1z * c1z) / vol;
  ci.hinv[2] = sqrt(c2x * c2x + c2y * c2y + c2z * c2z) / vol;
  for (int k = 0; k < 3; ++k) ci.pbc[k] = pbc[3 * b + k] != 0;
}

__device__ __forceinline__ void edge_vec(const Smem& s, const Cell64& ci, int i, int j, char4 sh, double& rx, double& ry,
                                         double& rz) {
  const double s0 = sh.x, s1 = sh.y, s2 = sh.z;
  rx = (s.x[3 * j] - s.x[3 * i]) + ((s0 * ci.c[0] + s1 * ci.c[3]) + s2 * ci.c[6]);
  ry = (s.x[3 * j + 1] - s.x[3 * i + 1]) + ((s0 * ci.c[1] + s1 * ci.c[4]) + s2 * ci.c[7]);
  rz = (s.x[3 * j + 2] - s.x[3 * i + 2]) + ((s0 * ci.c[2] + s1 * ci.c[5]) + s2 * ci.c[8]);
}

// Build the in-smem skin list (radius rl) + reverse map.  Returns status bits (block-uniform not required).
__device__ void build_list(const Smem& s, const Cell64& ci, int n, int max_nbr, double rl, int32_t* status) {
  const double rl2 = rl * rl;
  for (int i = threadIdx.x; i < n; i += NT) {
    int c = 0;
    bool over = false;
    const double xi = s.x[3 * i], yi = s.x[3 * i + 1], zi = s.x[3 * i + 2];
    for (int j = 0; j < n; ++j) {
      const double dx = s.x[3 * j] - xi, dy = s.x[3 * j + 1] - yi, dz = s.x[3 * j + 2] - zi;
      int lo[3], hi[3];
      for (int k = 0; k < 3; ++k) {
        if (ci.pbc[k]) {
          const double df = dx * ci.inv[k] + dy * ci.inv[3 + k] + dz * ci.inv[6 + k];
          const double w = rl * ci.hinv[k];
          lo[k] = (int)ceil(-df - w - 1e-9);
          hi[k] = (int)floor(-df + w + 1e-9);
        } else {
          lo[k] = hi[k] = 0;
        }
      }
      for (int s0 = lo[0]; s0 <= hi[0]; ++s0)
        for (int s1 = lo[1]; s1 <= hi[1]; ++s1)
          for (int s2 = lo[2]; s2 <= hi[2]; ++s2) {
            const double rx = dx + (((double)s0 * ci.c[0] + (double)s1 * ci.c[3]) + (double)s2 * ci.c[6]);
            const double ry = dy + (((double)s0 * ci.c[1] + (double)s1 * ci.c[4]) + (double)s2 * ci.c[7]);
            const double rz = dz + (((double)s0 * ci.c[2] + (double)s1 * ci.c[5]) + (double)s2 * ci.c[8]);
            const double d2 = rx * rx + ry * ry + rz * rz;
            if (d2 < rl2 && !(i == j && s0 == 0 && s1 == 0 && s2 == 0)) {
              if (c < max_nbr) {
                s.nj[i * max_nbr + c] = (short)j;
                s.ns[i * max_nbr + c] = make_char4((signed char)s0, (signed char)s1, (signed char)s2, 0);
                ++c;
              } else {
                over = true;
              }
            }
          }
    }
    s.cnt[i] = c;
    if (over) atomicOr(status, VSSR_STATUS_SLOT_OVERFLOW);
    s.x0[3 * i] = xi; s.x0[3 * i + 1] = yi; s.x0[3 * i + 2] = zi;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += NT) {
    for (int t = 0; t < s.cnt[i]; ++t) {
      const int j = s.nj[i * max_nbr + t];
      const char4 sh = s.ns[i * max_nbr + t];
      int r = 255;
      for (int u = 0; u < s.cnt[j]; ++u) {
        const char4 q = s.ns[j * max_nbr + u];
        if (s.nj[j * max_nbr + u] == i && q.x == -sh.x && q.y == -sh.y && q.z == -sh.z) { r = u; break; }
      }
      s.rev[i * max_nbr + t] = (unsigned char)r;
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------- Tersoff
struct TersP {
  double m, gamma, lam3, c, d, h, n, beta, lam2, B, R, D, lam1, A;
};
__device__ __forceinline__ TersP load_ters(const double* __restrict__ p) {
  TersP t;
  t.m = p[0]; t.gamma = p[1]; t.lam3 = p[2]; t.c = p[3]; t.d = p[4]; t.h = p[5]; t.n = p[6]; t.beta = p[7];
  t.lam2 = p[8]; t.B = p[9]; t.R = p[10]; t.D = p[11]; t.lam1 = p[12]; t.A = p[13];
  return t;
}
__device__ __forceinline__ void ters_fc(double r, double R, double D, double& fc, double& dfc) {
  if (r < R - D) { fc = 1.0; dfc = 0.0; }
  else if (r > R + D) { fc = 0.0; dfc = 0.0; }
  else {
    const double a = 1.5707963267948966 * (r - R) / D;
    fc = 0.5 * (1.0 - sin(a));
    dfc = -(0.7853981633974483 / D) * cos(a);
  }
}
// the four regime thresholds of b_ij (LAMMPS pair_tersoff c1..c4) depend on the parameter row only
__device__ __forceinline__ void ters_bij_consts(double n, double* c) {
  c[0] = pow(2.0 * n * 1.0e-16, -1.0 / n);
  c[1] = pow(2.0 * n * 1.0e-8, -1.0 / n);
  c[2] = 1.0 / c[1];
  c[3] = 1.0 / c[0];
}
__device__ __forceinline__ void ters_bij(double zeta, const TersP& p, const double* __restrict__ bc, double& b, double& db) {
  const double tmp = p.beta * zeta, n = p.n;
  double cl[4];
  if (bc) { cl[0] = bc[0]; cl[1] = bc[1]; cl[2] = bc[2]; cl[3] = bc[3]; }
  else ters_bij_consts(n, cl);
  const double c1 = cl[0], c2 = cl[1], c3 = cl[2], c4 = cl[3];
  double dbt;  // db/dtmp
  if (tmp > c1) { b = 1.0 / sqrt(tmp); dbt = -0.5 * b / tmp; }
  else if (tmp > c2) {
    const double tn = pow(tmp, -n), is = 1.0 / sqrt(tmp);
    b = (1.0 - tn / (2.0 * n)) * is;
    dbt = 0.5 * tn / tmp * is + (1.0 - tn / (2.0 * n)) * (-0.5 * is / tmp);
  }
  else if (tmp < c4) { b = 1.0; dbt = 0.0; }
  else if (tmp < c3) { b = 1.0 - pow(tmp, n) / (2.0 * n); dbt = -0.5 * pow(tmp, n - 1.0); }
  else {
    const double tn = pow(tmp, n);
    b = pow(1.0 + tn, -1.0 / (2.0 * n));
    dbt = -0.5 * pow(1.0 + tn, -1.0 / (2.0 * n) - 1.0) * tn / tmp;
  }
  db = p.beta * dbt;
}
// zeta term for (ij,k) and its derivatives w.r.t. r_ij, r_ik, cos
__device__ __forceinline__ void ters_zeta_term(const TersP& p, double rij, double rik, double cs, double& z, double& dz_drij,
                                               double& dz_drik, double& dz_dcos) {
  double fc, dfc;
  ters_fc(rik, p.R, p.D, fc, dfc);
  const double hc = p.h - cs, c2 = p.c * p.c, d2 = p.d * p.d;
  const double den = 1.0 / (d2 + hc * hc);
  const double g = p.gamma * (1.0 + c2 / d2 - c2 * den);
  const double dg = p.gamma * (-2.0 * c2 * hc) * den * den;
  const double dr = rij - rik;
  double arg, darg;  // arg and d arg / d rij
  if (p.m == 3.0) { const double t = p.lam3 * dr; arg = t * t * t; darg = 3.0 * t * t * p.lam3; }
  else { arg = p.lam3 * dr; darg = p.lam3; }
  double ex, dex;
  if (arg > 69.0776) { ex = 1.0e30; dex = 0.0; }
  else if (arg < -69.0776) { ex = 0.0; dex = 0.0; }
  else { ex = exp(arg); dex = ex * darg; }
  z = fc * g * ex;
  dz_drij = fc * g * dex;
  dz_drik = dfc * g * ex - fc * g * dex;
  dz_dcos = fc * dg * ex;
}

// ---------------------------------------------------------------------------------- pair table (Tersoff / SW)
// Directed pairs inside rcut, compacted per centre in skin-list order: count (thread per centre), warp scan, fill.
__device__ void build_pairs(const Smem& s, const Cell64& ci, int n, int max_nbr, double rcut, int32_t* status) {
  for (int i = threadIdx.x; i < n; i += NT) {
    const int cn = s.cnt[i];
    int c = 0;
    for (int t = 0; t < cn; ++t) {
      double x, y, z;
      edge_vec(s, ci, i, s.nj[i * max_nbr + t], s.ns[i * max_nbr + t], x, y, z);
      if (sqrt(x * x + y * y + z * z) < rcut) ++c;
    }
    s.pcnt[i] = c;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int base = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      const int c = i < n ? s.pcnt[i] : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (i < n) s.pstart[i] = base + incl - c;
      base += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
      s.pstart[n] = base;
      if (base > s.pcap) atomicOr(status, VSSR_STATUS_SLOT_OVERFLOW);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += NT) {
    const int cn = s.cnt[i], k0 = s.pstart[i];
    int c = 0;
    for (int t = 0; t < cn; ++t) {
      double x, y, z;
      edge_vec(s, ci, i, s.nj[i * max_nbr + t], s.ns[i * max_nbr + t], x, y, z);
      const double r = sqrt(x * x + y * y + z * z);
      unsigned char a = 255;
      if (r < rcut) {
        const int k = k0 + c;
        if (k < s.pcap && c < 255) {
          const double ir = 1.0 / r;
          s.pi[k] = (short)i; s.pslot[k] = (unsigned char)t;
          s.pu[4 * k] = x * ir; s.pu[4 * k + 1] = y * ir; s.pu[4 * k + 2] = z * ir; s.pu[4 * k + 3] = r;
          a = (unsigned char)c;
        }
        ++c;
      }
      s.aidx[i * max_nbr + t] = a;
    }
  }
  __syncthreads();
}
// pairs of centre i: [k0, k1) clipped to the table
__device__ __forceinline__ void pair_range(const Smem& s, int i, int& k0, int& k1) {
  k0 = min(s.pstart[i], s.pcap);
  k1 = min(s.pstart[i + 1], s.pcap);
}
// centre-own gradient and per-atom energy from the centre's pair slots (thread per centre, list order)
__device__ void pairs_to_centres(const Smem& s, int n) {
  for (int i = threadIdx.x; i < n; i += NT) {
    int k0, k1;
    pair_range(s, i, k0, k1);
    double gx = 0.0, gy = 0.0, gz = 0.0, e = 0.0;
    for (int k = k0; k < k1; ++k) {
      gx -= s.pG[3 * k]; gy -= s.pG[3 * k + 1]; gz -= s.pG[3 * k + 2];
      e += s.pq[3 * k];
    }
    s.own[3 * i] = gx; s.own[3 * i + 1] = gy; s.own[3 * i + 2] = gz;
    s.eat[i] = e;
  }
}

__device__ void tersoff_phase1(const Smem& s, const Cell64& ci, int n, int max_nbr, const double* __restrict__ params,
                               const double* __restrict__ bcs, int ntypes, double rcut, int32_t* status) {
  build_pairs(s, ci, n, max_nbr, rcut, status);
  const int np = min(s.pstart[n], s.pcap);
  // pass 1 (thread per pair ij): zeta_ij over the centre's other pairs, bond order, pair energy and its r-derivative
  for (int k = threadIdx.x; k < np; k += NT) {
    const int i = s.pi[k], ti = s.type[i];
    int k0, k1;
    pair_range(s, i, k0, k1);
    const int tj = s.type[s.nj[i * max_nbr + s.pslot[k]]];
    const int row = (ti * ntypes + tj) * ntypes + tj;
    const TersP pij = load_ters(params + (size_t)row * 14);
    const double ux = s.pu[4 * k], uy = s.pu[4 * k + 1], uz = s.pu[4 * k + 2], rij = s.pu[4 * k + 3];
    double e = 0.0, dEdr = 0.0, pref = 0.0;
    if (rij < pij.R + pij.D) {
      double zeta = 0.0;
      for (int q = k0; q < k1; ++q) {
        if (q == k) continue;
        const TersP pk = load_ters(params + (size_t)((ti * ntypes + tj) * ntypes + s.type[s.nj[i * max_nbr + s.pslot[q]]]) * 14);
        const double rik = s.pu[4 * q + 3];
        if (rik >= pk.R + pk.D) continue;
        const double cs = ux * s.pu[4 * q] + uy * s.pu[4 * q + 1] + uz * s.pu[4 * q + 2];
        double z, a1, a2, a3;
        ters_zeta_term(pk, rij, rik, cs, z, a1, a2, a3);
        zeta += z;
      }
      double fc, dfc, bb, db;
      ters_fc(rij, pij.R, pij.D, fc, dfc);
      ters_bij(zeta, pij, bcs ? bcs + 4 * row : nullptr, bb, db);
      const double er = pij.A * exp(-pij.lam1 * rij), ea = -pij.B * exp(-pij.lam2 * rij);
      const double fR = fc * er, dfR = er * (dfc - pij.lam1 * fc);
      const double fA = fc * ea, dfA = ea * (dfc - pij.lam2 * fc);
      e = 0.5 * (fR + bb * fA);
      dEdr = 0.5 * (dfR + bb * dfA);
      pref = 0.5 * fA * db;
    }
    s.pq[3 * k] = e; s.pq[3 * k + 1] = dEdr; s.pq[3 * k + 2] = pref;
  }
  __syncthreads();
  // pass 2 (thread per pair iu): everything of centre i's energy that pulls on neighbour u
  for (int k = threadIdx.x; k < np; k += NT) {
    const int i = s.pi[k], ti = s.type[i];
    int k0, k1;
    pair_range(s, i, k0, k1);
    const int tu = s.type[s.nj[i * max_nbr + s.pslot[k]]];
    const double ux = s.pu[4 * k], uy = s.pu[4 * k + 1], uz = s.pu[4 * k + 2], ru = s.pu[4 * k + 3];
    const double iru = 1.0 / ru;
    const double pref_u = s.pq[3 * k + 2];
    double gx = s.pq[3 * k + 1] * ux, gy = s.pq[3 * k + 1] * uy, gz = s.pq[3 * k + 1] * uz;
    for (int q = k0; q < k1; ++q) {
      if (q == k) continue;
      const double pref_q = s.pq[3 * q + 2];
      if (pref_u == 0.0 && pref_q == 0.0) continue;
      const int tq = s.type[s.nj[i * max_nbr + s.pslot[q]]];
      const double wx = s.pu[4 * q], wy = s.pu[4 * q + 1], wz = s.pu[4 * q + 2], rq = s.pu[4 * q + 3];
      const double cs = ux * wx + uy * wy + uz * wz;
      const double tx = (wx - cs * ux) * iru, ty = (wy - cs * uy) * iru, tz = (wz - cs * uz) * iru;   // d cos / d x_u
      if (pref_u != 0.0) {      // triplet (i; j = u, k = q): u is the bonded atom
        const TersP pk = load_ters(params + (size_t)((ti * ntypes + tu) * ntypes + tq) * 14);
        if (rq < pk.R + pk.D) {
          double z, dzj, dzk, dzc;
          ters_zeta_term(pk, ru, rq, cs, z, dzj, dzk, dzc);
          gx += pref_u * (dzj * ux + dzc * tx); gy += pref_u * (dzj * uy + dzc * ty); gz += pref_u * (dzj * uz + dzc * tz);
        }
      }
      if (pref_q != 0.0) {      // triplet (i; j = q, k = u): u is the third atom
        const TersP pk = load_ters(params + (size_t)((ti * ntypes + tq) * ntypes + tu) * 14);
        if (ru < pk.R + pk.D) {
          double z, dzj, dzk, dzc;
          ters_zeta_term(pk, rq, ru, cs, z, dzj, dzk, dzc);
          gx += pref_q * (dzk * ux + dzc * tx); gy += pref_q * (dzk * uy + dzc * ty); gz += pref_q * (dzk * uz + dzc * tz);
        }
      }
    }
    s.pG[3 * k] = gx; s.pG[3 * k + 1] = gy; s.pG[3 * k + 2] = gz;
  }
  __syncthreads();
  pairs_to_centres(s, n);
}

// ---------------------------------------------------------------------------------- SW
struct SWP {
  double eps, sigma, a, lam, gamma, cos0, A, B, p, q;
};
__device__ __forceinline__ SWP load_sw(const double* __restrict__ p) {
  SWP t;
  t.eps = p[0]; t.sigma = p[1]; t.a = p[2]; t.lam = p[3]; t.gamma = p[4]; t.cos0 = p[5]; t.A = p[6]; t.B = p[7];
  t.p = p[8]; t.q = p[9];
  return t;
}
// x^y for the exponents SW parameter files actually use (p = 4, q = 0), general pow otherwise
__device__ __forceinline__ double sw_pow(double x, double y) {
  if (y == 0.0) return 1.0;
  if (y == 4.0) { const double x2 = x * x; return x2 * x2; }
  return pow(x, y);
}

__device__ void sw_phase1(const Smem& s, const Cell64& ci, int n, int max_nbr, const double* __restrict__ params,
                          int ntypes, double rcut, int32_t* status) {
  build_pairs(s, ci, n, max_nbr, rcut, status);
  const int np = min(s.pstart[n], s.pcap);
  // pass 1 (thread per pair ij): two-body term (half per direction) and the pair's three-body exponential
  for (int k = threadIdx.x; k < np; k += NT) {
    const int i = s.pi[k], ti = s.type[i];
    const int tj = s.type[s.nj[i * max_nbr + s.pslot[k]]];
    const SWP pij = load_sw(params + (size_t)((ti * ntypes + tj) * ntypes + tj) * 10);
    const double cutij = pij.a * pij.sigma, rij = s.pu[4 * k + 3];
    double e = 0.0, dEdr = 0.0, live = 0.0, ex3 = 0.0, dex3 = 0.0;
    if (rij < cutij) {
      const double irij = 1.0 / rij;
      const double sr = pij.sigma * irij;
      const double srp = sw_pow(sr, pij.p), srq = sw_pow(sr, pij.q);
      const double rc = rij - cutij;
      const double ex = exp(pij.sigma / rc);
      const double pre = pij.A * pij.eps;
      const double poly = pij.B * srp - srq;
      const double phi = pre * poly * ex;
      const double dpoly = (-pij.p * pij.B * srp + pij.q * srq) * irij;
      const double dphi = pre * (dpoly * ex + poly * ex * (-pij.sigma / (rc * rc)));
      e = 0.5 * phi;
      dEdr = 0.5 * dphi;
      const double gs = pij.gamma * pij.sigma;
      ex3 = exp(gs / rc);
      dex3 = -gs / (rc * rc);
      live = 1.0;
    }
    s.pq[3 * k] = e; s.pq[3 * k + 1] = dEdr; s.pq[3 * k + 2] = live;
    s.pa[2 * k] = ex3; s.pa[2 * k + 1] = dex3;
  }
  __syncthreads();
  // pass 2 (thread per pair iu): pair term + the u-leg of every triplet (i; u, q) of the centre
  for (int k = threadIdx.x; k < np; k += NT) {
    double gx = 0.0, gy = 0.0, gz = 0.0, e3 = 0.0;
    if (s.pq[3 * k + 2] != 0.0) {
      const int i = s.pi[k], ti = s.type[i];
      int k0, k1;
      pair_range(s, i, k0, k1);
      const int tu = s.type[s.nj[i * max_nbr + s.pslot[k]]];
      const double ux = s.pu[4 * k], uy = s.pu[4 * k + 1], uz = s.pu[4 * k + 2], iru = 1.0 / s.pu[4 * k + 3];
      const double exu = s.pa[2 * k], dexu = s.pa[2 * k + 1];
      gx = s.pq[3 * k + 1] * ux; gy = s.pq[3 * k + 1] * uy; gz = s.pq[3 * k + 1] * uz;
      for (int q = k0; q < k1; ++q) {
        if (q == k || s.pq[3 * q + 2] == 0.0) continue;
        const int tq = s.type[s.nj[i * max_nbr + s.pslot[q]]];
        // LAMMPS orders a triplet by the neighbour list: (j, k) = (earlier, later) pair of the centre
        const double* pr = params + (size_t)((ti * ntypes + (k < q ? tu : tq)) * ntypes + (k < q ? tq : tu)) * 10;
        const double le = pr[3] * pr[0], cos0 = pr[5];
        const double wx = s.pu[4 * q], wy = s.pu[4 * q + 1], wz = s.pu[4 * q + 2];
        const double cs = ux * wx + uy * wy + uz * wz;
        const double dc = cs - cos0;
        const double ee = k < q ? exu * s.pa[2 * q] : s.pa[2 * q] * exu;
        const double h = le * dc * dc * ee;
        const double dh_dru = h * dexu, dh_dcos = 2.0 * le * dc * ee;
        gx += dh_dru * ux + dh_dcos * (wx - cs * ux) * iru;
        gy += dh_dru * uy + dh_dcos * (wy - cs * uy) * iru;
        gz += dh_dru * uz + dh_dcos * (wz - cs * uz) * iru;
        if (k < q) e3 += h;
      }
    }
    s.pG[3 * k] = gx; s.pG[3 * k + 1] = gy; s.pG[3 * k + 2] = gz;
    s.pq[3 * k] += e3;
  }
  __syncthreads();
  pairs_to_centres(s, n);
}

// ---------------------------------------------------------------------------------- EAM (funcfl, one element)
// LAMMPS `pair_style eam` (oracle/eam.py): the potential behind LAMMPSRunSurfCalc's Cu / Au toy
// runs (mcmc/calculators/calculators.py:755-811, tests/test_Cu.py, tests/test_Au.py).  `params` =
// [nrho, drho, nr, dr, rc, 0, 0, 0 | frho spline (nrho+1) x 7 | rhor spline (nr+1) x 7 | z2r spline (nr+1) x 7],
// the 7-coefficient splines of PairEAM::interpolate, built on the host (engine.eam_param_block).
constexpr int EAM_HDR = 8;
__device__ __forceinline__ void eam_lookup(const double* __restrict__ spl, double x, double rdx, int n, double& val,
                                           double& der) {
  double p = x * rdx + 1.0;
  int m = (int)p;
  m = m < 1 ? 1 : (m > n - 1 ? n - 1 : m);
  p -= (double)m;
  p = fmin(p, 1.0);
  const double* c = spl + (size_t)m * 7;
  val = ((__ldg(c + 3) * p + __ldg(c + 4)) * p + __ldg(c + 5)) * p + __ldg(c + 6);
  der = (__ldg(c) * p + __ldg(c + 1)) * p + __ldg(c + 2);
}

__device__ void eam_phase1(const Smem& s, const Cell64& ci, int n, int max_nbr, const double* __restrict__ params,
                           double rcut) {
  const int nrho = (int)params[0], nr = (int)params[2];
  const double rdrho = 1.0 / params[1], rdr = 1.0 / params[3];
  const double* frho = params + EAM_HDR;
  const double* rhor = frho + (size_t)(nrho + 1) * 7;
  const double* z2r = rhor + (size_t)(nr + 1) * 7;
  const double rhomax = (double)(nrho - 1) * params[1];
  // pass A: host electron density of every atom -> embedding energy F(rho_i) and F'(rho_i)
  for (int i = threadIdx.x; i < n; i += NT) {
    double rho = 0.0;
    const int cn = s.cnt[i];
    for (int t = 0; t < cn; ++t) {
      double x, y, z;
      edge_vec(s, ci, i, s.nj[i * max_nbr + t], s.ns[i * max_nbr + t], x, y, z);
      const double r = sqrt(x * x + y * y + z * z);
      if (r < rcut) {
        double v, d;
        eam_lookup(rhor, r, rdr, nr, v, d);
        rho += v;
      }
    }
    double F, dF;
    eam_lookup(frho, rho, rdrho, nrho, F, dF);
    if (rho > rhomax) F += dF * (rho - rhomax);      // linear extrapolation beyond the table
    s.eat[i] = F;
    s.aux[i] = dF;
  }
  __syncthreads();
  // pass B: pair term (half per direction) and the gradient of both terms along every directed edge i <- j
  for (int i = threadIdx.x; i < n; i += NT) {
    const int cn = s.cnt[i];
    const double dFi = s.aux[i];
    double gix = 0.0, giy = 0.0, giz = 0.0, ei = s.eat[i];
    double* Gi = s.G + (size_t)i * max_nbr * 3;
    for (int t = 0; t < cn; ++t) {
      const int j = s.nj[i * max_nbr + t];
      double x, y, z;
      edge_vec(s, ci, i, j, s.ns[i * max_nbr + t], x, y, z);
      const double r = sqrt(x * x + y * y + z * z);
      double gx = 0.0, gy = 0.0, gz = 0.0;
      if (r < rcut) {
        double rho_e, drho_e, z2, dz2;
        eam_lookup(rhor, r, rdr, nr, rho_e, drho_e);
        eam_lookup(z2r, r, rdr, nr, z2, dz2);
        const double ir = 1.0 / r;
        const double phi = z2 * ir, dphi = dz2 * ir - phi * ir;
        ei += 0.5 * phi;
        const double fpair = (dFi + s.aux[j]) * drho_e * 0.5 + 0.5 * dphi;
        gx = fpair * x * ir; gy = fpair * y * ir; gz = fpair * z * ir;
      }
      Gi[3 * t] = gx; Gi[3 * t + 1] = gy; Gi[3 * t + 2] = gz;
      gix -= gx; giy -= gy; giz -= gz;
    }
    s.own[3 * i] = gix; s.own[3 * i + 1] = giy; s.own[3 * i + 2] = giz;
    s.eat[i] = ei;
  }
}

// phase 1 (centre terms) + phase 2 (gather through reverse map) -> s.f = -dE/dx ; returns E (all threads)
__device__ double eval_forces(int kind, const Smem& s, const Cell64& ci, int n, int max_nbr,
                              const double* __restrict__ params, const double* __restrict__ bcs, int ntypes, double rcut,
                              int32_t* status) {
  if (kind == VSSR_POT_TERSOFF) tersoff_phase1(s, ci, n, max_nbr, params, bcs, ntypes, rcut, status);
  else if (kind == VSSR_POT_SW) sw_phase1(s, ci, n, max_nbr, params, ntypes, rcut, status);
  else eam_phase1(s, ci, n, max_nbr, params, rcut);
  __syncthreads();
  double e = 0.0, z0 = 0.0, z1 = 0.0;
  for (int j = threadIdx.x; j < n; j += NT) {
    double gx = s.own[3 * j], gy = s.own[3 * j + 1], gz = s.own[3 * j + 2];
    for (int t = 0; t < s.cnt[j]; ++t) {
      const int i = s.nj[j * max_nbr + t];
      const int r = s.rev[j * max_nbr + t];
      if (r == 255) continue;
      const double* g;
      if (kind == VSSR_POT_EAM) {
        g = s.G + ((size_t)i * max_nbr + r) * 3;
      } else {                        // j's slot in centre i's pair table, if j is inside i's cutoff right now
        const int a = s.aidx[i * max_nbr + r];
        if (a == 255) continue;
        g = s.pG + (size_t)(s.pstart[i] + a) * 3;
      }
      gx += g[0]; gy += g[1]; gz += g[2];
    }
    s.f[3 * j] = -gx; s.f[3 * j + 1] = -gy; s.f[3 * j + 2] = -gz;
    e += s.eat[j];
  }
  // fixed-order block sum of the energy
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  e = warp_sum(e);
  (void)z0; (void)z1;
  __syncthreads();
  if (lane == 0) s.red[wid] = e;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int q = 0; q < NW; ++q) tot += s.red[q];
  __syncthreads();
  return tot;
}

__device__ __forceinline__ double block_sum(double a, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  a = warp_sum(a);
  __syncthreads();
  if (lane == 0) red[wid] = a;
  __syncthreads();
  double r = 0.0;
#pragma unroll
  for (int q = 0; q < NW; ++q) r += red[q];
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_max(double a, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  a = warp_max(a);
  __syncthreads();
  if (lane == 0) red[wid] = a;
  __syncthreads();
  double r = red[0];
#pragma unroll
  for (int q = 1; q < NW; ++q) r = fmax(r, red[q]);
  __syncthreads();
  return r;
}

__device__ double max_cut(int kind, const double* __restrict__ params, int ntypes) {
  if (kind == VSSR_POT_EAM) return params[4];
  double c = 0.0;
  const int np = ntypes * ntypes * ntypes;
  for (int q = 0; q < np; ++q) {
    const double v = kind == VSSR_POT_TERSOFF ? params[q * 14 + 10] + params[q * 14 + 11] : params[q * 10 + 1] * params[q * 10 + 2];
    c = fmax(c, v);
  }
  return c;
}

// RELAX = false: single evaluation (list at the bare cutoff).  RELAX = true: FIRE loop.
template <bool RELAX>
__global__ void __launch_bounds__(NT, 2) classical_kernel(int kind, const double* __restrict__ params, int ntypes,
                                                       double* __restrict__ pos, const int32_t* __restrict__ types,
                                                       const uint8_t* __restrict__ fixed,
                                                       const int32_t* __restrict__ atom_ptr,
                                                       const double* __restrict__ cell, const uint8_t* __restrict__ pbc,
                                                       int n_max, int max_nbr, int relax_steps, double fmax_tol,
                                                       double skin, double* __restrict__ out_energy,
                                                       double* __restrict__ out8, double* __restrict__ out_forces,
                                                       double* __restrict__ out_eatom, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) char smem_raw[];
  Smem s;
  smem_layout(kind, n_max, max_nbr, smem_raw, &s);
  const int b = blockIdx.x;
  const int a0 = atom_ptr[b], a1 = atom_ptr[b + 1];
  int n = a1 - a0;
  if (n > n_max) {
    if (threadIdx.x == 0) atomicOr(status, VSSR_STATUS_NATOM_OVERFLOW);
    n = n_max;
  }
  Cell64 ci;
  load_cell64(cell, pbc, b, ci);
  for (int i = threadIdx.x; i < n; i += NT) {
    for (int c = 0; c < 3; ++c) { s.x[3 * i + c] = pos[3 * (a0 + i) + c]; s.v[3 * i + c] = 0.0; }
    s.type[i] = types[a0 + i];
    s.fixed[i] = RELAX ? fixed[a0 + i] : 0;
  }
  __syncthreads();
  // potential parameters: shared-memory copy when they fit (ntypes <= 3)
  __shared__ double sprm_buf[27 * 14];
  // (the EAM spline tables, ~84 KB, stay in global memory: read-only, L1/L2 resident)
  const int nprm = kind == VSSR_POT_EAM ? (1 << 30) : ntypes * ntypes * ntypes * (kind == VSSR_POT_TERSOFF ? 14 : 10);
  const double* sprm = params;
  if (nprm <= 27 * 14) {
    for (int q = threadIdx.x; q < nprm; q += NT) sprm_buf[q] = params[q];
    sprm = sprm_buf;
    __syncthreads();
  }
  // Tersoff: the regime thresholds of b_ij per parameter row (two pow each), formed once per CTA
  __shared__ double sbc_buf[27 * 4];
  const double* bcs = nullptr;
  if (kind == VSSR_POT_TERSOFF && ntypes <= 3) {
    for (int q = threadIdx.x; q < ntypes * ntypes * ntypes; q += NT) ters_bij_consts(sprm[q * 14 + 6], sbc_buf + 4 * q);
    bcs = sbc_buf;
    __syncthreads();
  }
  const double rcut = max_cut(kind, sprm, ntypes);
  const double rl = rcut + (RELAX ? skin : 0.0);
  build_list(s, ci, n, max_nbr, rl, status);

  double energy = eval_forces(kind, s, ci, n, max_nbr, sprm, bcs, ntypes, rcut, status);
  if (!RELAX) {
    for (int i = threadIdx.x; i < n; i += NT) {
      for (int c = 0; c < 3; ++c) out_forces[3 * (a0 + i) + c] = s.f[3 * i + c];
      if (out_eatom) out_eatom[a0 + i] = s.eat[i];
    }
    if (threadIdx.x == 0) out_energy[b] = energy;
    return;
  }

  // ---- FIRE (ASE defaults), whole-structure norms, fixed atoms masked ----
  double dt = 0.1, alpha = 0.1;
  int npos = 0, nsteps = 0;
  bool has_v = false, conv = false;
  double fabsmax = 0.0;
  const double half_skin2 = 0.25 * skin * skin;
  while (true) {
    double fm2 = 0.0, fam = 0.0;
    for (int i = threadIdx.x; i < n; i += NT) {
      const double fx = s.f[3 * i], fy = s.f[3 * i + 1], fz = s.f[3 * i + 2];
      fam = fmax(fam, fmax(fabs(fx), fmax(fabs(fy), fabs(fz))));
      if (!s.fixed[i]) fm2 = fmax(fm2, fx * fx + fy * fy + fz * fz);
    }
    fm2 = block_max(fm2, s.red);
    fabsmax = block_max(fam, s.red);
    conv = fm2 < fmax_tol * fmax_tol;
    if (conv || nsteps >= relax_steps) break;
    if (has_v) {
      double vf = 0.0, ff = 0.0, vv = 0.0;
      for (int i = threadIdx.x; i < n; i += NT) {
        if (s.fixed[i]) continue;
        for (int c = 0; c < 3; ++c) {
          const double f = s.f[3 * i + c], v = s.v[3 * i + c];
          vf += f * v; ff += f * f; vv += v * v;
        }
      }
      vf = block_sum(vf, s.red); ff = block_sum(ff, s.red); vv = block_sum(vv, s.red);
      if (vf > 0.0) {
        const double scale = alpha / sqrt(ff) * sqrt(vv);
        for (int i = threadIdx.x; i < n; i += NT)
          for (int c = 0; c < 3; ++c) {
            const double f = s.fixed[i] ? 0.0 : s.f[3 * i + c];
            s.v[3 * i + c] = (1.0 - alpha) * s.v[3 * i + c] + f * scale;
          }
        if (npos > 5) { dt = fmin(dt * 1.1, 1.0); alpha *= 0.99; }
        npos += 1;
      } else {
        for (int i = threadIdx.x; i < n; i += NT)
          for (int c = 0; c < 3; ++c) s.v[3 * i + c] = 0.0;
        alpha = 0.1; dt *= 0.5; npos = 0;
      }
    }
    has_v = true;
    double n2 = 0.0;
    for (int i = threadIdx.x; i < n; i += NT)
      for (int c = 0; c < 3; ++c) {
        const double f = s.fixed[i] ? 0.0 : s.f[3 * i + c];
        const double v = s.v[3 * i + c] + dt * f;
        s.v[3 * i + c] = v;
        const double dr = dt * v;
        n2 += dr * dr;
      }
    n2 = block_sum(n2, s.red);
    const double normdr = sqrt(n2);
    int moved = 0;
    for (int i = threadIdx.x; i < n; i += NT) {
      if (s.fixed[i]) continue;
      double m2 = 0.0;
      for (int c = 0; c < 3; ++c) {
        const double dr = dt * s.v[3 * i + c];
        const double x = s.x[3 * i + c] + (normdr > 0.2 ? 0.2 * dr / normdr : dr);
        s.x[3 * i + c] = x;
        const double dd = x - s.x0[3 * i + c];
        m2 += dd * dd;
      }
      if (m2 > half_skin2) moved = 1;
    }
    nsteps += 1;
    if (__syncthreads_or(moved)) build_list(s, ci, n, max_nbr, rl, status);
    energy = eval_forces(kind, s, ci, n, max_nbr, sprm, bcs, ntypes, rcut, status);
  }
  for (int i = threadIdx.x; i < n; i += NT)
    for (int c = 0; c < 3; ++c) {
      pos[3 * (a0 + i) + c] = s.x[3 * i + c];
      if (out_forces) out_forces[3 * (a0 + i) + c] = s.f[3 * i + c];
    }
  if (threadIdx.x == 0) {
    const bool oob = fabs(energy) > 1000.0 || fabsmax > 1000.0;  // mcmc/dynamics.py:159
    double* o = out8 + 8 * b;
    o[0] = oob ? 1000.0 : energy; o[1] = 0.0; o[2] = energy; o[3] = fabsmax; o[4] = (double)nsteps;
    o[5] = conv ? 1.0 : 0.0; o[6] = oob ? 1.0 : 0.0; o[7] = (double)(nsteps + 1);
  }
}

template <bool RELAX>
int set_smem(size_t bytes) {
  static size_t configured[64] = {};   // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  size_t& c = configured[dev & 63];
  if (bytes > c) {
    e = cudaFuncSetAttribute(classical_kernel<RELAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    c = bytes;
  }
  return 0;
}

}  // namespace

extern "C" size_t vssr_classical_smem_bytes(int32_t kind, int32_t n_max, int32_t max_nbr) {
  return smem_layout(kind, n_max, max_nbr, nullptr, nullptr);
}

extern "C" int vssr_classical_energy_forces(int32_t kind, const double* params, int32_t ntypes, const double* pos,
                                            const int32_t* types, const int32_t* atom_ptr, const double* cell,
                                            const uint8_t* pbc, int32_t n_struct, int32_t n_max, int32_t max_nbr,
                                            double* energy, double* forces, double* per_atom_energy, int32_t* status,
                                            void* stream) {
  if (!params || !pos || !types || !atom_ptr || !cell || !pbc || !energy || !forces || !status) return VSSR_ERR_ARG;
  if (kind != VSSR_POT_TERSOFF && kind != VSSR_POT_SW && kind != VSSR_POT_EAM) return VSSR_ERR_UNSUPPORTED;
  if (kind == VSSR_POT_EAM && ntypes != 1) return VSSR_ERR_UNSUPPORTED;   // funcfl: one element
  if (n_struct <= 0 || n_max <= 0 || n_max > 32767 || max_nbr <= 0 || max_nbr > 254) return VSSR_ERR_ARG;
  const size_t smem = smem_layout(kind, n_max, max_nbr, nullptr, nullptr);
  if (smem > 227 * 1024) return VSSR_ERR_ARG;
  int rc = set_smem<false>(smem);
  if (rc) return rc;
  VSSR_PROF(VSSR_K_CLASSICAL, (cudaStream_t)stream, classical_kernel<false><<<n_struct, NT, smem, (cudaStream_t)stream>>>(
      kind, params, ntypes, const_cast<double*>(pos), types, nullptr, atom_ptr, cell, pbc, n_max, max_nbr, 0, 0.0, 0.0,
      energy, nullptr, forces, per_atom_energy, status));
  return VSSR_OK;
}

extern "C" int vssr_classical_relax(int32_t kind, const double* params, int32_t ntypes, double* pos,
                                    const int32_t* types, const uint8_t* fixed, const int32_t* atom_ptr,
                                    const double* cell, const uint8_t* pbc, int32_t n_struct, int32_t n_max,
                                    int32_t max_nbr, int32_t relax_steps, double fmax, double skin, double* out,
                                    double* forces, int32_t* status, void* stream) {
  if (!params || !pos || !types || !fixed || !atom_ptr || !cell || !pbc || !out || !status) return VSSR_ERR_ARG;
  if (kind != VSSR_POT_TERSOFF && kind != VSSR_POT_SW && kind != VSSR_POT_EAM) return VSSR_ERR_UNSUPPORTED;
  if (kind == VSSR_POT_EAM && ntypes != 1) return VSSR_ERR_UNSUPPORTED;
  if (n_struct <= 0 || n_max <= 0 || n_max > 32767 || max_nbr <= 0 || max_nbr > 254 || relax_steps < 0 || skin < 0)
    return VSSR_ERR_ARG;
  const size_t smem = smem_layout(kind, n_max, max_nbr, nullptr, nullptr);
  if (smem > 227 * 1024) return VSSR_ERR_ARG;
  int rc = set_smem<true>(smem);
  if (rc) return rc;
  VSSR_PROF(VSSR_K_CLASSICAL, (cudaStream_t)stream, classical_kernel<true><<<n_struct, NT, smem, (cudaStream_t)stream>>>(kind, params, ntypes, pos, types, fixed, atom_ptr,
                                                                      cell, pbc, n_max, max_nbr, relax_steps, fmax, skin,
                                                                      nullptr, out, forces, nullptr, status));
  return VSSR_OK;
}

extern "C" int vssr_classical_relax_host(int32_t kind, const double* params, int32_t ntypes, double* pos,
                                         const int32_t* types, const uint8_t* fixed, const int32_t* atom_ptr,
                                         const double* cell, const uint8_t* pbc, int32_t n_struct, int32_t n_atoms,
                                         int32_t n_max, int32_t max_nbr, int32_t relax_steps, double fmax, double skin,
                                         double* out, double* forces, int32_t* status) {
  if (!params || !pos || !types || !fixed || !atom_ptr || !cell || !pbc || !out || !status) return VSSR_ERR_ARG;
  const int np = kind == VSSR_POT_EAM ? EAM_HDR + (((int)params[0] + 1) + 2 * ((int)params[2] + 1)) * 7
                                      : ntypes * ntypes * ntypes * (kind == VSSR_POT_TERSOFF ? 14 : 10);
  const size_t b_par = (size_t)np * 8, b_pos = (size_t)n_atoms * 24, b_typ = (size_t)n_atoms * 4, b_fix = n_atoms;
  const size_t b_ptr = (size_t)(n_struct + 1) * 4, b_cell = (size_t)n_struct * 72, b_pbc = (size_t)n_struct * 3;
  const size_t b_out = (size_t)n_struct * 64;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t total = al(b_par) + 2 * al(b_pos) + al(b_typ) + al(b_fix) + al(b_ptr) + al(b_cell) + al(b_pbc) + al(b_out) + 256;
  char* d = nullptr;
  VSSR_CUDA(cudaMalloc(&d, total));
  size_t off = 0;
  auto take = [&](size_t x) { char* p = d + off; off += al(x); return p; };
  double* d_par = (double*)take(b_par); double* d_pos = (double*)take(b_pos); double* d_f = (double*)take(b_pos);
  int32_t* d_typ = (int32_t*)take(b_typ); uint8_t* d_fix = (uint8_t*)take(b_fix); int32_t* d_ptr = (int32_t*)take(b_ptr);
  double* d_cell = (double*)take(b_cell); uint8_t* d_pbc = (uint8_t*)take(b_pbc); double* d_out = (double*)take(b_out);
  int32_t* d_status = (int32_t*)take(4);
  cudaStream_t st = 0;
  int rc = VSSR_OK;
  cudaError_t e;
#define H2D(dst, src, nbytes) if ((e = cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) { cudaFree(d); return (int)e; }
  H2D(d_par, params, b_par) H2D(d_pos, pos, b_pos) H2D(d_typ, types, b_typ) H2D(d_fix, fixed, b_fix)
  H2D(d_ptr, atom_ptr, b_ptr) H2D(d_cell, cell, b_cell) H2D(d_pbc, pbc, b_pbc)
#undef H2D
  cudaMemsetAsync(d_status, 0, 4, st);
  rc = vssr_classical_relax(kind, d_par, ntypes, d_pos, d_typ, d_fix, d_ptr, d_cell, d_pbc, n_struct, n_max, max_nbr,
                            relax_steps, fmax, skin, d_out, d_f, d_status, st);
  if (rc == VSSR_OK) {
    cudaMemcpyAsync(pos, d_pos, b_pos, cudaMemcpyDeviceToHost, st);
    if (forces) cudaMemcpyAsync(forces, d_f, b_pos, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(out, d_out, b_out, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(status, d_status, 4, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = (int)e;
  }
  cudaFree(d);
  return rc;
}
