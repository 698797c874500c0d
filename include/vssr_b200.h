/*
 * vssr_b200.h — C ABI of libvssr_b200.so, the B200 (sm_100a) energy/force + relaxation engine
 * for the VSSR-MC hot path of learningmatter-mit/surface-sampling.
 *
 * Every entry point is plain C: device pointers, sizes, a CUDA stream passed as void*.  No torch
 * types.  Nothing here allocates device memory: the caller (PyTorch, as plumbing) owns every
 * buffer, including the workspace whose size the *_workspace_bytes() queries return.  All work
 * is enqueued on `stream` and returns without synchronising; the int return value is 0 or a
 * negative VSSR_ERR_* / positive cudaError_t from the launch.  Conditions that can only be known
 * on the device (edge-capacity overflow, neighbour-slot overflow) are reported through the
 * caller-owned `status` word(s), never by exceptions.
 *
 * Reference interfaces these replace (paths relative to the reference tree):
 *   vssr_nbr_build            NFF AtomsBatch.update_nbr_list  <- mcmc/dynamics.py:129,
 *                             flags mcmc/utils/misc.py:34-42
 *   vssr_painn_energy_grad    NFF Painn.forward + energy_grad <- EnsembleNFF.calculate <-
 *                             mcmc/calculators/calculators.py:484 (one call per model there)
 *   vssr_ensemble_stats       EnsembleNFF mean/std + unit/offset conversion (same call site);
 *                             consumers mcmc/calculators/calculators.py:118-135,
 *                             mcmc/uncertainty/uncertainty.py:190-210
 *   vssr_fire_step            ase.optimize.FIRE.step + FixAtoms  <- mcmc/dynamics.py:127,133,141
 *   vssr_painn_relax          optimize_slab(optimizer="FIRE")    <- mcmc/dynamics.py:83-170
 *   vssr_classical_energy_forces / vssr_classical_relax
 *                             LAMMMPSCalc.run_lammps_energy / run_lammps_opt
 *                             <- mcmc/calculators/calculators.py:600-640, mcmc/dynamics.py:107-116
 *   vssr_system_reduce        get_system_val <- mcmc/uncertainty/prediction.py:181-223
 *
 * Batch layout ("ragged flat"): B structures (one per chain) are concatenated; structure b owns
 * atoms [atom_ptr[b], atom_ptr[b+1]).  A = atom_ptr[B] atoms in total.
 */
#ifndef VSSR_B200_H
#define VSSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSSR_OK 0
#define VSSR_ERR_ARG (-1)        /* bad argument (null pointer, size)        */
#define VSSR_ERR_WORKSPACE (-2)  /* workspace too small                      */
#define VSSR_ERR_UNSUPPORTED (-3)

/* status bits written by kernels into caller-owned int words */
#define VSSR_STATUS_EDGE_OVERFLOW 1   /* neighbour list needed more than e_cap edges      */
#define VSSR_STATUS_SLOT_OVERFLOW 2   /* classical kernel: > max_nbr neighbours per atom  */
#define VSSR_STATUS_NATOM_OVERFLOW 4  /* classical kernel: structure larger than n_max    */

int vssr_version(void);
/* compute capability of the current device * 10 (100 for B200); <0 on error */
int vssr_device_cc(void);

/* ------------------------------------------------------------------------------------------
 * Neighbour list (directed, periodic by explicit image enumeration, receiver-sorted CSR).
 * Edge (i <- j, S) exists iff  d2 < cutoff^2 and d2 != 0  with, in fp32 and without FMA,
 *   off = (S0*a + S1*b) + S2*c ;  r = (x_j - x_i) + off ;  d2 = (rx*rx + ry*ry) + rz*rz.
 * Rows are sorted by (j, S0, S1, S2).  `col` holds GLOBAL atom indices, `shift` 4 int8 per edge
 * (S0,S1,S2,0).  rowptr[A] is the total edge count (may exceed e_cap; then nothing beyond e_cap
 * is written and VSSR_STATUS_EDGE_OVERFLOW is or-ed into *status).
 * ---------------------------------------------------------------------------------------- */
int vssr_nbr_build(const float* pos /*[A,3]*/, const int32_t* atom_ptr /*[B+1]*/,
                   const float* cell /*[B,3,3] rows a,b,c*/, const uint8_t* pbc /*[B,3]*/,
                   int32_t n_struct, int32_t n_atoms, float cutoff,
                   int32_t* deg /*[A] scratch*/, int32_t* rowptr /*[A+1]*/, int32_t* col /*[e_cap]*/,
                   int8_t* shift /*[e_cap,4]*/, int64_t e_cap, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------
 * PaiNN ensemble (feat 128, 20 rbf, 3 conv, swish, cosine cutoff, excluded volume 12/1.5).
 * Packed weights: n_models blocks of vssr_painn_weight_floats() floats, laid out as documented in
 * surface_sampling_b200/csrc/painn_layout.h (Python packs them: engine.pack_painn_weights).
 * Outputs: energy[M,B] (fp64, kcal/mol, sum over atoms incl. excluded volume) and
 * grad[M,A,3] (fp32, dE/dx in kcal/mol/A).  One call = forward + hand-written backward for all
 * models and all structures.
 * ---------------------------------------------------------------------------------------- */
int64_t vssr_painn_weight_floats(void);
size_t vssr_painn_workspace_bytes(int32_t n_models, int32_t n_atoms, int64_t e_cap);
int vssr_painn_energy_grad(const float* weights, int32_t n_models,
                           const float* pos /*[A,3]*/, const int32_t* z /*[A]*/,
                           const int32_t* atom_ptr, const float* cell, int32_t n_struct,
                           int32_t n_atoms,
                           int32_t max_atoms_per_struct /* host-known max of atom_ptr[b+1]-atom_ptr[b] (> 0): sizes the
                              shared-memory staging of the message kernels; structures beyond one staging area are
                              covered by several sender-window launches */,
                           const int32_t* rowptr, const int32_t* col,
                           const int8_t* shift, int64_t e_cap, float cutoff,
                           const void* filter_cache /* from vssr_painn_filter_cache_build, or NULL */,
                           int32_t fc_n0, int64_t fc_e_cap0, int32_t fc_flags /* VSSR_FC_* */,
                           void* workspace, size_t workspace_bytes,
                           double* energy /*[M,B]*/, float* grad /*[M,A,3]*/,
                           float* embedding /*[M,A,128] or NULL: final scalar features*/,
                           void* stream);

/* Radial-filter memo for frozen pairs.  In VSSR-MC every chain shares one frozen bulk framework (the
 * first n0 atoms of every structure, FixAtoms / `group bulk`), so most edges keep a bit-identical
 * distance in every chain and at every FIRE step.  Their per-model, per-layer filter rows w(d) and
 * dw/dd are computed ONCE here from the framework alone; the evaluation kernels look an edge up by
 * (j_local, lattice shift) in the framework's CSR row and use the memo only when its fp32 distance is
 * bitwise equal, so the result never depends on a promise by the caller.  Synchronises the stream.
 * (No counterpart in the reference: it re-evaluates every filter on every call.)
 *
 * fc_flags of the evaluation calls:
 *   VSSR_FC_CONSTRAINED_GRAD  the caller applies FixAtoms to the framework's frozen atoms (ASE zeroes
 *       their forces before any optimiser sees them, ase Atoms.get_forces(apply_constraint=True) as used
 *       by mcmc/dynamics.py:25-141), so dE/dx of those atoms is not computed: a memoised edge joins two
 *       frozen atoms and its whole position-gradient branch is skipped.  Energies, and the gradient rows
 *       of every other atom, are unchanged; gradient rows of frozen framework atoms are returned as 0. */
#define VSSR_FC_CONSTRAINED_GRAD 1
#define VSSR_FC_NO_PAIR 2   /* debugging: do not use the two-structures-per-CTA memo kernels */
size_t vssr_painn_filter_cache_bytes(int32_t n_models, int32_t n0, int64_t e_cap0);
size_t vssr_painn_filter_cache_workspace_bytes(int32_t n0, int64_t e_cap0);   /* scratch for the build */
int vssr_painn_filter_cache_build(const float* weights, int32_t n_models, const float* pos0 /*[n0,3]*/,
                                  const float* cell /*[3,3]*/, const uint8_t* pbc /*[3]*/,
                                  const uint8_t* fixed0 /*[n0]*/, int32_t n0, float cutoff, float skin,
                                  int64_t e_cap0, void* cache, size_t cache_bytes, void* workspace,
                                  size_t workspace_bytes, int32_t* nslots_out /* host, may be NULL */,
                                  void* stream);

/* EnsembleNFF semantics: per model E_eV = E_kcal/23.06052 + offset_ev[b]; mean and population std
 * over models; forces = -mean(grad)/23.06052, forces_std = std(grad)/23.06052.               */
int vssr_ensemble_stats(const double* energy /*[M,B]*/, const float* grad /*[M,A,3]*/,
                        const double* offset_ev /*[B] or NULL*/, const int32_t* atom_ptr,
                        int32_t n_models, int32_t n_struct, int32_t n_atoms,
                        double* e_mean /*[B]*/, double* e_std /*[B]*/,
                        float* f_mean /*[A,3]*/, float* f_std /*[A,3] or NULL*/, void* stream);

/* Per-structure reductions of a per-atom quantity (ragged): out[B,6] =
 * sum, max, min, mean, mean_squared, rms  (get_system_val).                                    */
int vssr_system_reduce(const float* per_atom /*[A]*/, const int32_t* atom_ptr, int32_t n_struct,
                       float* out /*[B,6]*/, void* stream);
/* ||forces_std|| per atom (EnsembleUncertainty.get_forces_uncertainty, order="std")            */
int vssr_atom_norm(const float* vec /*[A,3]*/, int32_t n_atoms, float* out /*[A]*/, void* stream);

/* ------------------------------------------------------------------------------------------
 * Batched FIRE (ASE defaults dt=0.1 maxstep=0.2 dtmax=1 Nmin=5 finc=1.1 fdec=0.5 astart=0.1
 * fa=0.99), one structure per CTA, fp64 state.  `fire_state` is 8 doubles per structure:
 * [dt, a, n_pos_steps, nsteps, converged, has_velocity, fmax_masked, max_abs_raw_force].
 * vssr_fire_step does what one pass of ase Dynamics.irun does after a force evaluation:
 * convergence test on the masked forces (max_i |F_i| < fmax), and, if not converged and
 * nsteps < max_steps, one FIRE step (fixed atoms: force zeroed, position frozen).
 * ---------------------------------------------------------------------------------------- */
int vssr_fire_init(double* fire_state /*[B,8]*/, double* vel /*[A,3]*/, int32_t n_struct,
                   int32_t n_atoms, void* stream);
int vssr_fire_step(double* pos /*[A,3]*/, float* pos32 /*[A,3]*/, double* vel /*[A,3]*/,
                   const float* forces /*[A,3] eV/A raw*/, const uint8_t* fixed /*[A]*/,
                   const int32_t* atom_ptr, int32_t n_struct, double* fire_state,
                   int32_t max_steps, double fmax, void* stream);

/* optimize_slab(optimizer="FIRE") for the PaiNN ensemble, all structures at once, no host
 * round trip: neighbour list at cutoff+skin once, then <= relax_steps+1 ensemble evaluations.
 * Outputs per structure: out[B,8] = [energy_eV (mean, clamped to 1000 if oob), energy_std,
 * raw_energy, max|F| raw, nsteps, converged, energy_oob, n_evals].                              */
size_t vssr_painn_relax_workspace_bytes(int32_t n_models, int32_t n_atoms, int64_t e_cap);
int vssr_painn_relax(const float* weights, int32_t n_models, double* pos /*[A,3] in/out*/,
                     const int32_t* z, const uint8_t* fixed, const int32_t* atom_ptr,
                     const float* cell, const uint8_t* pbc, const double* offset_ev,
                     int32_t n_struct, int32_t n_atoms, int32_t max_atoms_per_struct, float cutoff, float skin,
                     int32_t relax_steps, double fmax, int64_t e_cap,
                     const void* filter_cache, int32_t fc_n0, int64_t fc_e_cap0, int32_t fc_flags,
                     void* workspace, size_t workspace_bytes, double* out /*[B,8]*/, float* forces /*[A,3]*/,
                     float* forces_std /*[A,3] or NULL*/, int32_t* status, void* stream);

/* Work actually done by the LAST evaluation that used a workspace (bench.py: executed-work roofline; no counterpart in
 * the reference).  out5 (device, 5 x int64) = [direct edges (filter evaluated), memoised edges, direct edges whose
 * receiver is a frozen framework atom, canonical structures (group memo kernels), edges of the cutoff+skin list].
 * `workspace` is the buffer given to vssr_painn_energy_grad / vssr_painn_relax with the SAME (n_models, n_atoms,
 * e_cap); `rowptr` the neighbour list used (the relax variant finds it inside its own workspace). */
int vssr_painn_edge_stats(const void* workspace, int32_t n_models, int32_t n_atoms, int64_t e_cap,
                          const int32_t* atom_ptr, int32_t n_struct, const int32_t* rowptr,
                          const void* filter_cache, int32_t fc_n0, int64_t fc_e_cap0, int64_t* out5, void* stream);
int vssr_painn_relax_edge_stats(const void* relax_workspace, int32_t n_models, int32_t n_atoms, int64_t e_cap,
                                const int32_t* atom_ptr, int32_t n_struct, const void* filter_cache,
                                int32_t fc_n0, int64_t fc_e_cap0, int64_t* out5, void* stream);

/* ------------------------------------------------------------------------------------------
 * Classical many-body potentials (fp64): LAMMPS `pair_style tersoff` and `pair_style sw` forms.
 * One CTA per structure; the structure (<= n_max atoms) lives in shared memory for the whole
 * relaxation.  `params`:
 *   Tersoff: [ntypes^3, 14] doubles in LAMMPS file order (m gamma lambda3 c d costheta0 n beta
 *            lambda2 B R D lambda1 A), index ((t_i*ntypes)+t_j)*ntypes+t_k.
 *   SW     : [ntypes^3, 10] doubles (epsilon sigma a lambda gamma costheta0 A B p q), same index.
 *   EAM    : [nrho, drho, nr, dr, rc, 0, 0, 0] then the three 7-coefficient spline tables of LAMMPS
 *            PairEAM::interpolate, rows 0..n (row 0 unused): F(rho) (nrho+1 rows), rho(r) (nr+1), r*phi(r) (nr+1).
 * ---------------------------------------------------------------------------------------- */
#define VSSR_POT_TERSOFF 0
#define VSSR_POT_SW 1
#define VSSR_POT_EAM 2   /* LAMMPS `pair_style eam`, single-element funcfl (LAMMPSRunSurfCalc: mcmc/calculators/
                            calculators.py:755-811, tests/test_Cu.py, tests/test_Au.py); ntypes must be 1 */
/* dynamic shared memory of one CTA (<= 227 KB): Tersoff / SW keep a table of 8 * n_max directed pairs, EAM per-slot
 * gradients [n_max][max_nbr] */
size_t vssr_classical_smem_bytes(int32_t kind, int32_t n_max, int32_t max_nbr);
int vssr_classical_energy_forces(int32_t kind, const double* params, int32_t ntypes,
                                 const double* pos /*[A,3]*/, const int32_t* types /*[A]*/,
                                 const int32_t* atom_ptr, const double* cell /*[B,3,3]*/,
                                 const uint8_t* pbc, int32_t n_struct, int32_t n_max,
                                 int32_t max_nbr, double* energy /*[B]*/, double* forces /*[A,3]*/,
                                 double* per_atom_energy /*[A] or NULL*/, int32_t* status,
                                 void* stream);
/* out[B,8] as in vssr_painn_relax (energy_std = 0). pos updated in place.                       */
int vssr_classical_relax(int32_t kind, const double* params, int32_t ntypes, double* pos,
                         const int32_t* types, const uint8_t* fixed, const int32_t* atom_ptr,
                         const double* cell, const uint8_t* pbc, int32_t n_struct, int32_t n_max,
                         int32_t max_nbr, int32_t relax_steps, double fmax, double skin,
                         double* out /*[B,8]*/, double* forces /*[A,3] or NULL*/,
                         int32_t* status, void* stream);

/* Host-buffer convenience entry (what a reference-side FFI stub binds for the LAMMPS-style
 * path): copies in, relaxes, copies out, synchronises.  All pointers are HOST memory.          */
int vssr_classical_relax_host(int32_t kind, const double* params, int32_t ntypes, double* pos,
                              const int32_t* types, const uint8_t* fixed, const int32_t* atom_ptr,
                              const double* cell, const uint8_t* pbc, int32_t n_struct,
                              int32_t n_atoms, int32_t n_max, int32_t max_nbr, int32_t relax_steps,
                              double fmax, double skin, double* out, double* forces,
                              int32_t* status);

/* number of kernels this library has enqueued since load (bench.py: gpu_launches); kernels replayed from the
 * relaxation's CUDA graph count once per replay.  vssr_graph_launch_count: cudaGraphLaunch calls issued (the
 * iterations 1 .. relax_steps-1 of vssr_painn_relax are one graph launch each; VSSR_NO_GRAPH=1 turns that off).  */
int64_t vssr_launch_count(void);
int64_t vssr_graph_launch_count(void);

/* Optional per-kernel-class profile (bench.py roofline): when enabled, every launch is bracketed
 * by a cudaEvent pair on its own stream.  Classes: 0 nbr, 1 edge geometry, 2 GEMM, 3 message fwd,
 * 4 message bwd, 5 elementwise, 6 readout, 7 ensemble stats, 8 FIRE, 9 classical.
 * vssr_profile_collect synchronises the device and returns summed milliseconds / launch counts. */
int vssr_kernel_class_count(void);
int vssr_profile_enable(int on);
int vssr_profile_collect(double* ms /*[n_class]*/, int64_t* launches /*[n_class]*/, int n_class);

#ifdef __cplusplus
}
#endif
#endif /* VSSR_B200_H */
