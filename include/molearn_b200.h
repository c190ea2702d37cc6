/* molearn_b200.h - C ABI of the B200-native molearn physics hot path.
 *
 * Every entry point takes plain pointers and sizes (no torch / C++ types) so it can be bound
 * with ctypes, cffi or any other FFI.  The reference has no native boundary for this path:
 * its TorchProteinEnergy is a Python object built from ATen ops.  The one native energy
 * boundary molearn does have - the OpenMM plugin call
 *   torchMultiStructureE(pos_ptr, force_ptr, energy_ptr, n_particles, batch_size)
 * at src/molearn/loss_functions/openmm_thread.py:221-241 - is the precedent followed here:
 * the caller owns every buffer, the library works in place on raw device pointers, energies
 * are per conformation, and backward multiplies a saved force by the upstream gradient
 * (openmm_thread.py:290-321).  Each declaration cites the reference code it replaces.
 *
 * Conventions
 *   - return value >= 0 = success (0 unless stated otherwise), negative = error (MLP_E_*); the message of
 *     the last error of the calling thread is returned by mlp_last_error().
 *   - "dev" pointers are device memory of the device the handle was created on; everything
 *     else is host memory.
 *   - no entry point allocates device memory, synchronises the device or copies to the host
 *     after mlp_energy_create(); all work is ordered on the cudaStream_t passed as `stream`.
 *   - no kernel uses atomics: every output is a fixed-order sum, bit-reproducible run to run.
 */
#ifndef MOLEARN_B200_H
#define MOLEARN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLP_ABI_VERSION 1

/* error codes */
#define MLP_OK            0
#define MLP_E_ARG        -1   /* bad argument (null pointer, negative size, ...)                */
#define MLP_E_IO         -2   /* parameter file missing / unreadable (reference: Exception)      */
#define MLP_E_KEY        -3   /* unknown residue / atom / force-field parameter (ref: KeyError)  */
#define MLP_E_PARSE      -4   /* malformed force-field line (reference: Exception)               */
#define MLP_E_CUDA       -5   /* CUDA runtime error                                              */
#define MLP_E_NOGPU      -6   /* no usable CUDA device - there is no CPU fallback                */
#define MLP_E_WORKSPACE  -7   /* workspace too small                                             */

/* mlp_topology_build flags */
#define MLP_TOPO_FIX_H    1u  /* get_convolutions(fix_h=True), torch_protein_energy_utils.py:434 */

/* energy terms: column order of every [B,4] array below */
#define MLP_TERM_BOND     1u
#define MLP_TERM_ANGLE    2u
#define MLP_TERM_TORSION  4u
#define MLP_TERM_NB       8u
#define MLP_TERM_BONDED   7u
#define MLP_TERM_ALL     15u

/* non-bonded functional form (TorchProteinEnergy(NB=...), torch_protein_energy.py:129-132) */
#define MLP_NB_REPULSIVE  0u  /* _cdist_nb      : A/w(d,1.9)^12 + qq/w(d,0.4)             :232-236 */
#define MLP_NB_FULL       1u  /* _cdist_nb_full : ... - B/w(d,1.9)^6                      :224-230 */

/* mlp_energy_eval flags */
#define MLP_EVAL_ACCUMULATE_GRAD 1u /* grad += instead of grad = (grad must hold valid data) */

typedef struct mlp_topology mlp_topology;
typedef struct mlp_energy mlp_energy;

/* Message of the last error raised on the calling thread ("" if none). */
const char* mlp_last_error(void);
int mlp_abi_version(void);

/* ---------------------------------------------------------------------------------------------
 * Host side: force-field parsing and topology construction.  Pure CPU, needs no GPU.
 * Replaces get_amber_parameters() (torch_protein_energy_utils.py:86-189) and the v=5 topology
 * part of get_convolutions() (:359-745, 813-836).  Index lists and their order are identical
 * to the reference's bond_idxs / angle_idxs / torsion_idxs; parameters are the same fp64 values.
 *
 * param_dir   directory holding amino12.lib, parm10.dat, frcmod.ff14SB (read at run time)
 * atom_names  [n_atoms] PDB atom names  (pdb_atom_names[:,0])
 * res_names   [n_atoms] residue names   (pdb_atom_names[:,1])
 * resids      [n_atoms] residue numbers (pdb_atom_names[:,2])
 * The inputs are not modified (the reference renames in place, utils.py:420-487; the
 * normalised names can be read back with mlp_topology_atoms()).
 */
int mlp_topology_build(const char* param_dir, int32_t n_atoms,
                       const char* const* atom_names, const char* const* res_names,
                       const int32_t* resids, uint32_t flags, mlp_topology** out);

/* counts[0..7] = n_atoms, n_bonds, n_angles, n_torsions, P (torsion series slots),
 *                n_excl (unique unordered excluded pairs), n_14, n_nb_pairs_lo32 (unused=0) */
int mlp_topology_counts(const mlp_topology* t, int64_t counts[8]);

/* Copy the topology out (any pointer may be NULL).  Shapes:
 *   bonds [n_bonds,2]  bond_para [n_bonds,2]=(r0,K)        (utils.py:705-707)
 *   angles [n_angles,3] angle_para [n_angles,2]=(theta0 rad,K)  (utils.py:708-710)
 *   torsions [n_torsions,4] torsion_para [n_torsions,4,P]=(factor,barrier,phase rad,|period|)
 *   excl [n_excl,2] unique i<j pairs zeroed at utils.py:818-830 (pair semantics, SURVEY NB-3)
 */
int mlp_topology_export(const mlp_topology* t, int32_t* bonds, double* bond_para,
                        int32_t* angles, double* angle_para,
                        int32_t* torsions, double* torsion_para, int32_t* excl);

/* Per-atom data after normalisation.  names/resnames/types are [n_atoms][8] NUL-padded. */
int mlp_topology_atoms(const mlp_topology* t, char* names, char* resnames, char* types,
                       double* charges, float* vdw_R, float* vdw_eps);

/* 1-4 pairs in torsion order [n_14,2] (bond_14_idxs, utils.py:687; computed, unused by E). */
int mlp_topology_pairs14(const mlp_topology* t, int32_t* pairs14);

/* The 1-4 scaled parameter arrays of utils.py:833-836, [n_14] fp32 each, in the reference's fp32 arithmetic:
 * vdw_14R = 0.5*(R_i+R_j), vdw_14e = sqrt(e_i+e_j) (sic), q1q2_14 = q_i*q_j.  The reference computes them and
 * never uses them in an energy; they are exported for callers that add a scaled 1-4 term.  Any pointer may be NULL. */
int mlp_topology_params14(const mlp_topology* t, float* vdw_14R, float* vdw_14e, float* q1q2_14);

void mlp_topology_destroy(mlp_topology* t);

/* ---------------------------------------------------------------------------------------------
 * Device side.
 */

/* Host-only inspection of the bonded kernel's work plan (no device needed; used by the CPU tests).
 * The bonded kernel evaluates WORK ITEMS: up to one torsion (i,j,k,l), one angle (i,j,k) and one bond
 * ((i,j) or (j,k)) on the same four atoms; every tile of `tile_atoms` consecutive atoms gets the
 * items that touch one of its atoms.  Rows of `items` ([capacity][9] int32):
 *   tile, i, j, k, l (global atom indices; unused roles repeat i), torsion id | -1, angle id | -1,
 *   bond id | -1, 1 if the bond sits on (j,k) else 0.
 * `*n_items` receives the number of rows (call with items = NULL to size the buffer);
 * stats[0] = tiles, [1] = max gradient contributions per atom, [2] = 1 if a torsion phase needs sine terms,
 * [3] = live (non-zero) torsions. */
int mlp_bonded_plan(const mlp_topology* t, int tile_atoms, int64_t* n_items, int32_t* items, int64_t capacity,
                    int64_t stats[4]);

/* Host only: the Lennard-Jones classes the repulsive pair kernel's type-pair table is built from.  Atoms with the same
 * (R, eps) share a class (numbered in order of first appearance); w2[a*n+b] = ((12 eps_ab)^(1/12) R_ab)^2 with the
 * reference's combination rules R_ab = (R_a+R_b)/2, eps_ab = sqrt(eps_a eps_b) (utils.py:814-815), so that
 * w2^6 / 12 = vdw_A[i,j] of utils.py:816 for any non-excluded pair of those classes.  atom_class [n_atoms] and w2
 * [capacity >= n^2] may be NULL (call once with w2 = NULL to size it); *n_classes receives n. */
int mlp_nb_pair_table(const mlp_topology* t, int32_t* atom_class, float* w2, int64_t capacity, int64_t* n_classes);

/* Number of CUDA devices visible (0 if none / no driver).  Never fails. */
int mlp_device_count(void);

/* Upload the topology to `device` and build the kernels' packed tables.  Replaces
 * TorchProteinEnergy._roll_init (torch_protein_energy.py:70-133): the roll/conv packing is an
 * implementation detail of the reference; the explicit lists are evaluated instead.  O(N)
 * device memory - the dense [N,N] vdw_A / vdw_B / q1q2 matrices (utils.py:813-830) are never
 * formed. */
int mlp_energy_create(const mlp_topology* t, int device, uint32_t nb_mode, mlp_energy** out);
void mlp_energy_destroy(mlp_energy* e);

/* Bytes of scratch device memory mlp_energy_eval needs for a batch of B conformations evaluated with
 * this term_mask, with (want_grad != 0) or without a gradient output.
 * The bulk is the pair term's partial forces: 3 KB per 128 x 128 tile pair and conformation (0.5 MB per MurD
 * conformation, 11 MB at 10 715 atoms).  It is bounded: a batch is evaluated in chunks of conformations whose
 * partial forces fit 384 MB (MLP_NB_WS_MB at create time; each chunk is one pair launch + one reduction), and a
 * system whose single conformation would need more than 64 MB (beyond ~26 k atoms) is evaluated band by band
 * through the rows of its tile-pair triangle (<= 64 MB per conformation and launch: four bands at 51 432 atoms;
 * MLP_NB_BANDS=0 selects block work items instead: 49 MB per conformation, 2 % slower).  Energy-only calls need
 * 8 bytes per tile pair and conformation. */
int64_t mlp_energy_workspace_bytes(const mlp_energy* e, int64_t B, uint32_t term_mask, int want_grad);

/* Packing facts, for benchmarks and tests: info[0..7] = bonded tiles, atoms per bonded tile, max
 * gradient contributions per atom, bonded shared memory bytes per CTA, NB tiles per edge, NB tile
 * pairs, masked NB tile pairs, padded atom count. */
int mlp_energy_info(const mlp_energy* e, int64_t info[8]);

/* Lennard-Jones classes (distinct (R, eps) of utils.py:814-815, plus one zero class for the padding atoms) when the
 * repulsive pair kernel reads w_ij^2 = ((12 eps_ij)^(1/12) R_ij)^2 from its type-pair table in shared memory, 0 when
 * it computes w_ij per pair (more than 7 classes, NB = 'full', or MLP_NB_LUT=0 in the environment at create time). */
int mlp_energy_nb_table_classes(const mlp_energy* e);

/* Energies and (optionally) gradient of B <= 65535 conformations.
 *
 *   x          dev, logical shape [B,3,N] fp32 addressed as x[b*sxb + c*sxc + n*sxn]
 *              (the reference layout [B,3,N] contiguous is sxb=3N,sxc=N,sxn=1; a decoder output
 *              [B,M,3] sliced to N atoms is sxb=3M,sxc=1,sxn=3).  sxc, sxn >= 0 and one conformation must
 *              span fewer than 2^31 elements ((N-1)*sxn + 2*sxc < 2^31; same for the gradient), else MLP_E_ARG.
 *   energies   dev [B,4] fp32 (bond, angle, torsion, nb) per conformation, or NULL.  Terms not
 *              in term_mask are written as 0.   (E-1..E-3, NB-2 of SURVEY 8a;
 *              torch_protein_energy.py:276-305, 224-239)
 *   grad       dev, same logical shape as x addressed with (sgb,sgc,sgn), or NULL for
 *              energy only.  grad = sum_t w[b,t] * dE_t/dx over the terms in term_mask.
 *   weights    dev [B,4] fp32 or NULL (= all ones).  The upstream gradient of a backward pass.
 *   workspace  dev, >= mlp_energy_workspace_bytes(e,B,term_mask,grad!=NULL) bytes, 16-byte aligned
 *   stream     cudaStream_t
 * Non-finite coordinates are not an error: they propagate into that conformation's outputs.
 * Returns the number of kernels it queued (>= 0) or a negative MLP_E_* code.
 *
 * Concurrency: work is ordered on `stream`; when both bonded and non-bonded terms are requested the bonded
 * kernel runs on a private stream of the handle that forks from and joins back into `stream` (so the call is
 * still stream-ordered for the caller and capturable into a CUDA graph).  One handle serves one host thread at
 * a time; use one handle per device and per concurrently calling thread.  Independent batches may alternate
 * over two streams through the same handle (calls from one thread): the next batch's pair kernel then fills
 * the last wave of the current one.
 */
int mlp_energy_eval(mlp_energy* e, const float* x, int64_t B,
                    int64_t sxb, int64_t sxc, int64_t sxn,
                    float* energies,
                    float* grad, int64_t sgb, int64_t sgc, int64_t sgn,
                    const float* weights, uint32_t term_mask, uint32_t flags,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* Number of kernel launches the last mlp_energy_eval on this handle issued. */
int mlp_energy_last_launches(const mlp_energy* e);

/* Optional per-kernel timing for benchmarks.  While enabled, mlp_energy_eval brackets each of
 * its kernels with CUDA events on the launching stream.  mlp_energy_profile_read waits for the
 * recorded events, returns the summed milliseconds and launch counts per kernel
 * (0 = unused, 1 = bonded, 2 = non-bonded pair kernel, 3 = non-bonded force reduction) and resets the totals. */
int mlp_energy_profile(mlp_energy* e, int enable);
int mlp_energy_profile_read(mlp_energy* e, double ms[4], int64_t counts[4]);

/* Per-conformation RMSD to K target structures (scan reducers; replaces the NumPy / biobox
 * loops of MolearnAnalysis.scan_error_from_target / scan_error / get_error, analyser.py:277-305,
 * 626-737, with the intended per-point formula sqrt(mean_atoms sum_xyz d^2), SURVEY S-3/S-4).
 *   x        dev [B,3,N] logical with strides as above;  value used is x*scale + shift
 *   targets  dev [K,N,3] fp32 contiguous (Angstrom)
 *   out      dev [B,K] fp32
 *   flags    MLP_RMSD_ALIGN: RMSD after optimal rigid superposition (Kabsch / Horn quaternion
 *            eigenvalue, fp64) - what biobox Molecule.rmsd computes at analyser.py:297, 671
 *            MLP_RMSD_PAIRED: K must equal B; structure b is scored against target b only and
 *            out is [B] (get_error: every structure against its own reconstruction,
 *            analyser.py:277-305, without the B x B matrix)
 */
#define MLP_RMSD_ALIGN 1u
#define MLP_RMSD_PAIRED 2u
int mlp_rmsd_eval(const float* x, int64_t B, int64_t N,
                  int64_t sxb, int64_t sxc, int64_t sxn, float scale, float shift,
                  const float* targets, int64_t K, float* out, uint32_t flags, void* stream);

/* Per-structure geometry statistics (scan reducers scan_bondlength / scan_ca_chirality and their
 * get_bondlengths / get_inversions helpers, analyser.py:340-473, 831-861, 980-1014).
 *   x          dev [B,3,N] logical with strides; lengths are computed on x*scale
 *   pairs      dev [n_pairs,2] atom indices; pair_group dev [n_pairs] group id in [0,n_groups), n_groups <= 8
 *   quads      dev [n_quads,4] = (N, CA, C, CB) indices
 *   out        dev [B, 2*n_groups+1] = mean and population std of the pair lengths of each group, then the
 *              number of quadruples with a negative triple product ((n-ca)x(c-ca)).(cb-ca)
 */
int mlp_geometry_eval(const float* x, int64_t B, int64_t sxb, int64_t sxc, int64_t sxn, float scale,
                      const int32_t* pairs, const int32_t* pair_group, int64_t n_pairs, int64_t n_groups,
                      const int32_t* quads, int64_t n_quads, float* out, void* stream);

/* Internal coordinates of every structure over index lists: the device form of MolearnAnalysis.get_bondlengths /
 * get_angles / get_dihedrals (analyser.py:418-566; formulas :791-861, the dihedral in its standard form).
 *   x         dev [B,3,N] logical with strides; B <= 65535
 *   pairs     dev [n_pairs,2]   -> distance |p0-p1| * scale
 *   triples   dev [n_triples,3] -> angle p0-p1-p2 in radians
 *   quads     dev [n_quads,4]   -> dihedral p0-p1-p2-p3 in radians (-pi, pi]
 *   out       dev [B, n_pairs + n_triples + n_quads] fp32: distances, then angles, then dihedrals
 */
int mlp_internal_coords_eval(const float* x, int64_t B, int64_t sxb, int64_t sxc, int64_t sxn, float scale,
                             const int32_t* pairs, int64_t n_pairs, const int32_t* triples, int64_t n_triples,
                             const int32_t* quads, int64_t n_quads, float* out, void* stream);

/* Measurement aid (bench.py): one launch of a pure FP32 FMA loop - FFMA (packed = 0) or FFMA2 / fma.rn.f32x2
 * (packed = 1) - on `stream`; the caller times it with events.  Returns the flops of the launch (> 0), or, when
 * scratch is NULL, the number of floats of device scratch it needs; negative = error.  The non-bonded roofline's
 * "measured FP32 peak" is this number, next to the nominal 148 SM x 128 lanes x 2 x clock. */
int64_t mlp_probe_fp32(float* scratch, int64_t scratch_floats, int packed, int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOLEARN_B200_H */
