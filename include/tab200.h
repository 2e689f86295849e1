/*
 * tab200.h -- C ABI of libtab200.so, the B200 (sm_100a) implementation of the
 * TensorAlloy energy / force / virial hot path.
 *
 * The reference (Bismarrck/tensoralloy) has NO C ABI: its boundary is the
 * Python operator surface plus one `sess.run(ops, feed_dict)` call
 * (tensoralloy/calculator.py:368-369).  Each entry point below names the
 * reference interface it replaces.  The Python mirror of the reference classes
 * (tensoralloy_b200/calculator.py, transformer/, nn/) binds these with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative TAB_E* code otherwise and
 *     never throws; tab_last_error() returns a thread-local message.
 *   - `d_` pointers are DEVICE pointers, `h_` pointers are HOST pointers; the
 *     caller owns every buffer.  The library owns only opaque handles.
 *   - all work is enqueued on the caller's CUDA stream (`stream` is a
 *     cudaStream_t passed as void*; NULL = legacy default stream).
 *   - one handle may be used from one host thread at a time (the reference
 *     calculator is not thread-safe either: calculator.py:368-370).
 *   - lattice matrices are 3x3 row-major with ROWS = lattice vectors (ASE / the
 *     reference, transformer/universal.py:446).
 *   - atom order of every input and output is the CALLER's order; the library's
 *     internal cell-sorted order never leaks.
 */
#ifndef TAB200_H
#define TAB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAB_OK            0
#define TAB_EINVAL       -1   /* bad argument */
#define TAB_ECUDA        -2   /* CUDA runtime error, see tab_last_error() */
#define TAB_ENOMEM       -3
#define TAB_EUNSUPPORTED -4
#define TAB_ESTATE       -5   /* call order violated (e.g. eval before build) */

/* precision of the model arithmetic: reference tensoralloy/precision.py:21-134
 * ('high' = float64, eps 1e-14; 'medium' = float32, eps 1e-8). */
#define TAB_PRECISION_HIGH   0
#define TAB_PRECISION_MEDIUM 1

typedef struct tab_nbr   tab_nbr;    /* cell list + neighbour lists of one structure */
typedef struct tab_model tab_model;  /* device-resident potential parameters        */

int         tab_version(void);
const char *tab_last_error(void);

/* ------------------------------------------------------------------------
 * Neighbour lists.
 * Replaces  ase.neighborlist.neighbor_list('ijSdD', atoms, rc)  as called by
 * get_radial_metadata (transformer/universal.py:58) and
 * find_neighbor_size_of_atoms (neighbor.py:84), plus the Python index-map loops
 * universal.py:69-106 (the slot of a pair is its position in the list).
 * ---------------------------------------------------------------------- */
int tab_nbr_create(tab_nbr **out);
int tab_nbr_free(tab_nbr *nbr);

/* Build the cell list, the periodic ghost images and the neighbour lists of
 * `n` atoms.  d_pos: [n,3] float64; d_types: [n] int32 element indices
 * (0..15) or NULL (all 0); h_cell: 9 doubles; h_pbc: 3 ints; rc: cutoff.
 * Membership is decided exactly like ASE: D = pos[j]-pos[i]+S.cell,
 * sqrt(D.D) < rc, float64, self pair excluded.  Synchronises the stream once
 * (list size read-back). */
int tab_nbr_build(tab_nbr *nbr, int32_t n, const double *d_pos,
                  const int32_t *d_types, const double *h_cell,
                  const int32_t *h_pbc, double rc, void *stream);

/* Domain-decomposed build (one rank of a spatial decomposition).  d_pos / d_types
 * hold n_owned owned atoms followed by n_halo HALO atoms received from other
 * ranks, already shifted into this rank's frame (so that D = pos[j] - pos[i]).
 * h_cell spans the local binning frame whose corner is h_origin (NULL = 0);
 * decomposed directions must be non-periodic in h_pbc (their images arrive as
 * halo atoms), the others get periodic images as usual.  Neighbour rows are built
 * for the owned atoms only; indices >= n_owned in exports refer to halo atoms.
 * The reference has no spatial decomposition. */
int tab_nbr_build_dd(tab_nbr *nbr, int32_t n_owned, int32_t n_halo,
                     const double *d_pos, const int32_t *d_types,
                     const double *h_cell, const double *h_origin,
                     const int32_t *h_pbc, double rc, void *stream);

/* Halo packing for the exchange of a spatial decomposition:
 * d_dst[k, :] = d_src[d_idx[k], :] (+ h_shift for 3-column position rows; the
 * sender applies the periodic shift).  d_idx: int64 row indices. */
int tab_pack_rows(const double *d_src, const int64_t *d_idx, int32_t m, int32_t ncol,
                  const double *h_shift, double *d_dst, void *stream);

/* Building blocks of a small all-reduce over NVLink peer memory (the 10-double
 * [E, virial] sum of the spatial decomposition; the reference has no such step):
 *   tab_peer_put   stores d_src[0..n) into slot `slot` (n doubles wide) of EACH of the
 *                  n_peers buffers whose device addresses are listed in d_peer_ptrs
 *                  (device array; the buffers are peer-mapped, e.g. torch symmetric
 *                  memory) -- one kernel, the stores travel over NVLink;
 *   tab_sum_slots  d_out[q] = sum_s d_slots[s * n + q]  (after a cross-rank barrier). */
int tab_peer_put(const double *d_src, int32_t n, const uint64_t *d_peer_ptrs,
                 int32_t n_peers, int32_t slot, void *stream);
int tab_sum_slots(const double *d_slots, int32_t n_slots, int32_t n, double *d_out,
                  void *stream);
/*   tab_reduce_slots  as tab_sum_slots for the entries [0, n_sum); the entries [n_sum, n) are
 *                     reduced with MAX (the largest displacement since the list build rides
 *                     along with the [E, virial] sums: one collective decides the rebuild). */
int tab_reduce_slots(const double *d_slots, int32_t n_slots, int32_t n, int32_t n_sum,
                     double *d_out, void *stream);

/* Rebuild-time kernels of the spatial decomposition (csrc/dd.cu; no reference counterpart).
 * Stable partitions (count -> scan -> scatter, no atomics: reproducible order).
 *   tab_dd_partition  migration: d_state [n, ncol] float64 rows, column 0 = x.  x is wrapped
 *                     into [0, lx); rows whose slab floor(x / width) is `rank` go to d_keep
 *                     (capacity n rows), rows of the left / right ring neighbour to
 *                     d_mail_left / d_mail_right = [count, rows ...] (float64; may be
 *                     peer-mapped buffers of the neighbour GPUs; capacity mail_cap rows).
 *                     d_counts int32 [8]: kept, to the left, to the right, lost (further than
 *                     an adjacent slab), 1 if a mailbox overflowed.  d_work: int32
 *                     [3 ceil(n / 256)] scratch.
 *   tab_dd_send_sets  halo send sets: indices (int64, ascending) of the atoms with
 *                     x < x_left_below -> d_idx_left, x >= x_right_from -> d_idx_right (an atom
 *                     may be in both); d_counts int32 [2]; d_work int32 [2 ceil(n / 256)]. */
int tab_dd_partition(const double *d_state, int32_t n, int32_t ncol, double lx, double width,
                     int32_t world, int32_t rank, double *d_keep, double *d_mail_left,
                     double *d_mail_right, int32_t mail_cap, int32_t *d_counts,
                     int32_t *d_work, void *stream);
int tab_dd_send_sets(const double *d_pos, int32_t n, double x_left_below, double x_right_from,
                     int64_t *d_idx_left, int64_t *d_idx_right, int32_t *d_counts,
                     int32_t *d_work, void *stream);

/* Batch of independent structures in ONE handle ("structure-parallel batches"; replaces
 * the padded [B, N+1, 3] / [B, nij_max, .] tensors of BatchUniversalTransformer,
 * transformer/universal.py:921-1388, and the per-structure ASE neighbour lists behind them).
 * Structure s owns atoms [h_offsets[s], h_offsets[s+1]) of d_pos / d_types; h_cells:
 * [n_struct, 9]; h_pbc: [n_struct, 3].  Meant for small structures (~100 atoms): the
 * candidates of an atom are all atoms and periodic images of its own structure.  Every
 * evaluation entry point accepts a batch handle; per-atom outputs cover all atoms of the
 * batch in caller order, and d_energy / d_virial become [n_struct] / [n_struct, 9]
 * (tab_atomic_jvp: d_A [n_struct, 9]).  Not supported on batch handles: tab_nbr_update,
 * tab_eam_hessian, the domain-decomposition passes. */
int tab_nbr_build_batch(tab_nbr *nbr, int32_t n_struct, const int32_t *h_offsets,
                        const double *d_pos, const int32_t *d_types, const double *h_cells,
                        const int32_t *h_pbc, double rc, void *stream);
int tab_nbr_batch_size(const tab_nbr *nbr);   /* 0 = single structure */

/* Keep the lists, refresh the positions (and optionally the cell): the MD step
 * between two rebuilds.  h_cell may be NULL (unchanged). */
int tab_nbr_update(tab_nbr *nbr, const double *d_pos, const double *h_cell,
                   void *stream);

/* MD-valid list reuse.  The reference rebuilds its lists on every call
 * (transformer/universal.py:58, one ase neighbor_list per get_np_feed_dict).  A handle with a
 * skin builds its lists with the radius rc + skin; the EAM / ADP pair kernels mask r >= rc
 * (entries beyond the model's cutoff contribute exactly 0; the symmetry-function kernels are
 * masked by their cutoff function), so that tab_nbr_update + evaluation equals a fresh build +
 * evaluation as long as no atom has moved further than skin / 2 since the build.
 *   tab_nbr_set_skin          sticky; takes effect at the next build (0 = exact lists)
 *   tab_nbr_max_displacement  largest |R - R_build| seen by the LAST tab_nbr_update (0 right
 *                             after a build); synchronises the stream (4-byte read-back);
 *                             *h_skin (may be NULL) receives the skin of the current lists.
 *                             A caller rebuilds when 2 * max_disp > skin.
 * Lists with a skin are refused by the consumers that hand list entries to the caller
 * (tab_nbr_export, tab_pairs_export, tab_eam_hessian): build those with skin = 0. */
int tab_nbr_set_skin(tab_nbr *nbr, double skin);
int tab_nbr_max_displacement(tab_nbr *nbr, double *h_max_disp, double *h_skin, void *stream);
/* the same quantity written to DEVICE memory (*d_out, Angstrom) without a synchronisation: the
 * spatial decomposition reduces it over the ranks inside the step (tab_reduce_slots). */
int tab_nbr_displacement_device(tab_nbr *nbr, double *d_out, void *stream);

/* Sizes, the quantities of neighbor.py:34-47 NeighborSize.  nij = number of
 * directed pairs; nnl_max = max neighbours of one atom (all species);
 * n_ext = owned + ghost atoms held on the device. */
int tab_nbr_sizes(const tab_nbr *nbr, int64_t *nij, int32_t *nnl_max,
                  int32_t *n_ext);

/* Per-atom neighbour counts in caller order, d_counts: [n] int32. */
int tab_nbr_counts(const tab_nbr *nbr, int32_t *d_counts, void *stream);

/* Export the list as the reference's (ilist, jlist, n1) arrays
 * (universal.py:58), caller atom indices, rows sorted by i, inside a row in the
 * library's deterministic cell order.  d_i, d_j: [nij] int32; d_S: [nij,3] int32. */
int tab_nbr_export(const tab_nbr *nbr, int32_t *d_i, int32_t *d_j, int32_t *d_S,
                   void *stream);

/* ------------------------------------------------------------------------
 * EAM / Finnis-Sinclair / ADP models.
 * Replaces the TF sub-graphs built by EamNN._get_model_outputs
 * (nn/eam/eam.py:495-570), EamAlloyNN._build_rho_nn (alloy.py:128-196),
 * EamFsNN._build_rho_nn (fs.py:146-203), AdpNN (adp.py:315-586), the potential
 * functions of nn/eam/potentials/*, and the autograd outputs of
 * BasicNN.build (nn/basic.py:276-354,679-787).
 * ---------------------------------------------------------------------- */
#define TAB_EAM_ALLOY 0   /* rho is a function of the neighbour element      */
#define TAB_EAM_FS    1   /* rho is a function of the ordered (centre,nbr) pair */
#define TAB_EAM_ADP   2   /* alloy + dipole + quadrupole terms               */

/* function kinds (one per rho / phi / embed / dipole / quadrupole slot) */
#define TAB_FN_ZERO          0
#define TAB_FN_ZHOU_RHO      1   /* zjw04.py:245-277   p = f_eq,beta,lamda,r_eq */
#define TAB_FN_ZHOU_PHI      2   /* zjw04.py:187-227   p = A,alpha,kappa,B,beta,lamda,r_eq */
#define TAB_FN_ZHOU_PHI_MIX  3   /* zjw04.py:229-243   p = [7 of a][4 rho of a][7 of b][4 rho of b] */
#define TAB_FN_ZHOU_EMBED    4   /* zjw04.py:279-389   p = Fn0..3,F0..3,eta,Fe,rho_e,rho_s */
#define TAB_FN_ZHOU_EMBED_XC 5   /* zjw04.py:440-550   same p, sigmoid blended */
#define TAB_FN_SUTTON_RHO     6   /* sutton90.py:61-78    p = a            (a/r)^6 */
#define TAB_FN_SUTTON_PHI     7   /* sutton90.py:43-59    p = b            (b/r)^12 */
#define TAB_FN_SQRT_EMBED     8   /* sutton90.py:80-97, grimmes.py:86-101  p = G   -G sqrt(rho) */
#define TAB_FN_AGRAWAL_RHO    9   /* agrawal.py:57-83     p = A,B,re,rc,m */
#define TAB_FN_AGRAWAL_PHI   10   /* agrawal.py:124-152   p = D,alpha,re,rc,m */
#define TAB_FN_AGRAWAL_EMBED 11   /* agrawal.py:85-122    p = F0,F1,beta,gamma */
#define TAB_FN_GRIMES_RHO    12   /* grimmes.py:62-84     p = n */
#define TAB_FN_GRIMES_PHI    13   /* grimmes.py:41-60     p = A,rho,C,D,gamma,r0 */
#define TAB_FN_MISHIN_EMBED  14   /* mishin.py:196-260    p = s1..s7,eps */
#define TAB_FN_MISHIN_POLAR  15   /* mishin.py:262-315, generic.py:52-84  p = p1,p2,p3,rc,h */
#define TAB_FN_SPLINE        16   /* cubic spline table (io/lammps.py:60-72 Spline; the
                                     reference's missing extension/interp CubicInterpolator):
                                     aux = offset (in intervals) into the model's
                                     coefficient pool, p = x0, 1/dx, n_intervals */
#define TAB_FN_MLP           17   /* 'nn' function (eam.py:174-190): MLP of the scalar
                                     argument, hidden layers with bias + activation, linear
                                     output without bias.  p = n_hidden (<= 4), TAB_ACT_*,
                                     widths (<= 64); weights in the coefficient pool at
                                     4*aux doubles: [w0, b0][W1, b1]...[w_out], W in-major */
#define TAB_FN_POWCUT_RHO    18   /* msah11.py:303-352  p = order, n, (f_i, rc_i) x n:
                                   sum f_i max(rc_i - r, 0)^order                      */
#define TAB_FN_MSAH_PHI      19   /* msah11.py:52-301   piecewise pair function in the
                                   coefficient pool (aux, units of 4 doubles):
                                   n_poly, first(lo, hi, c0, (b, c) x 4), second(lo, hi,
                                   c0..c3), then per tail: lo, hi, n_term, (a, order) x n */
#define TAB_FN_MSAH_EMBED_AL 20   /* msah11.py:400-411  p = c1, c2:
                                   -sqrt(x) + c1 x^2 - c2 x ln x  (x >= 1e-12, else 0)  */
#define TAB_FN_MSAH_EMBED_FE 21   /* msah11.py:412-420  p = c3, c4:
                                   -sqrt(x) - c3 x^2 + c4 x^4                          */
#define TAB_FN_MAX_PARAMS    32

typedef struct tab_fn {
    int32_t kind;
    int32_t aux;
    double  p[TAB_FN_MAX_PARAMS];
} tab_fn;

/* Create a model.  n_el <= 16 elements (indices = positions in the SORTED
 * element list, as in the reference's transformer.elements).
 *   rho   : n_el*n_el entries, rho[a*n_el+b] = density at a centre of element a
 *           from a neighbour of element b (alloy: independent of a)
 *   phi   : n_el*n_el entries (symmetric)
 *   embed : n_el entries
 *   dipole, quadrupole : n_el*n_el entries or NULL (ADP only)             */
int tab_eam_create(tab_model **out, int32_t kind, int32_t n_el,
                   const tab_fn *rho, const tab_fn *phi, const tab_fn *embed,
                   const tab_fn *dipole, const tab_fn *quadrupole);
int tab_model_free(tab_model *model);

/* Attach the coefficient pool used by TAB_FN_SPLINE (and TAB_FN_MLP) entries: interval k of a
 * table holds 4 doubles (c0, c1, c2, c3) of  f(x) = c0 + c1 t + c2 t^2 + c3 t^3,
 * t = x - (x0 + k dx).  h_coeffs: [n_intervals_total * 4] host doubles. */
int tab_eam_set_splines(tab_model *model, const double *h_coeffs, int64_t n_doubles);

/* One E + F + virial evaluation on the lists held by `nbr`.
 *   d_energy : [1]   total energy (eV)                      basic.py:742-787
 *   d_eatom  : [n]   per-atom energies or NULL              eam.py:287-298
 *   d_forces : [n,3] F = -dE/dR (eV/A) or NULL              basic.py:281-287
 *   d_virial : [9]   sum_pairs dE/dD (x) D  (eV) or NULL    basic.py:306-317
 * All outputs float64 in caller atom order.  `precision` selects the arithmetic
 * of the pair functions (TAB_PRECISION_*). */
int tab_eam_eval(tab_model *model, tab_nbr *nbr, int32_t precision,
                 double *d_energy, double *d_eatom, double *d_forces,
                 double *d_virial, void *stream);

/* The same evaluation split at the point where a spatial decomposition must
 * exchange F'(rho) of boundary atoms:
 *   pass1: rho_i, F(rho_i), F'(rho_i) of the owned atoms;
 *          d_fprime [n_owned] (caller order) receives F' (may be NULL)
 *   pass2: d_fprime_halo [n_halo] = F' of the halo atoms in caller order (NULL
 *          when there are none); then forces / energy / virial of the owned
 *          atoms.  Energy and virial are this rank's partial sums. */
int tab_eam_pass1(tab_model *model, tab_nbr *nbr, int32_t precision,
                  double *d_fprime, void *stream);
int tab_eam_pass2(tab_model *model, tab_nbr *nbr, int32_t precision,
                  const double *d_fprime_halo, double *d_energy, double *d_eatom,
                  double *d_forces, double *d_virial, void *stream);

/* Tabulate ONE function of the model on a grid: the evaluator behind `export_to_setfl`
 * (nn/eam/alloy.py:198-381, fs.py:205-..., adp.py:588-794, which run the TF graph on
 * r = k dr, rho = k drho).  which: 0 rho, 1 phi, 2 embed, 3 dipole, 4 quadrupole; index:
 * a * n_el + b (centre a, neighbour b) for pair functions, the element for the embedding.
 * d_x, d_y [n] float64, d_dy (derivative) may be NULL. */
int tab_eam_tabulate(tab_model *model, int32_t which, int32_t index, int32_t n,
                     const double *d_x, double *d_y, double *d_dy, void *stream);

/* Spatial decomposition without a second exchange: lists built by tab_nbr_build_dd over
 * [own atoms | inner halo (<= rc from the slab)] as the row-owning group + the outer halo
 * (rc .. 2 rc); rho, F', ADP moments and forces of the inner halo are recomputed on this
 * rank; d_mask [n_owned_group] (int32, caller order, 1 = own atom) keeps them out of the
 * partial energy / virial.  This is the decomposition path of ADP (no moment exchange
 * exists) and an alternative to tab_eam_pass1/pass2 for EAM / FS.  d_forces / d_eatom are
 * valid in the rows of the own atoms. */
int tab_eam_eval_dd(tab_model *model, tab_nbr *nbr, int32_t precision,
                    const int32_t *d_mask, double *d_energy, double *d_eatom,
                    double *d_forces, double *d_virial, void *stream);

/* Analytic Hessian d2E/dR_a dR_b of an EAM / FS model (float64 only):
 * d_hessian [n,3,n,3] in caller atom order.  Replaces tf.hessians(E, R) of
 * BasicNN._get_hessian_op (nn/basic.py:410-421). */
int tab_eam_hessian(tab_model *model, tab_nbr *nbr, double *d_hessian, void *stream);

/* Analytic elastic-constant op of an EAM / FS model (float64):
 *   C_ijkl = [ (d virial_ij / d h)^T h ]_kl   (eV),
 * the reference's definition, nn/constraint/elastic.py:24-91 (tf.gradients(virial[i, j], cell)
 * at fixed Cartesian positions), from the closed-form second derivatives of the pair and
 * embedding functions.  d_out [n, 36]: the share of every atom (caller order), rows = Voigt pair
 * (ij), columns = Voigt pair (kl), order xx yy zz yz xz xy; the caller sums over the atoms and
 * divides by V * GPa. */
int tab_eam_elastic(tab_model *model, tab_nbr *nbr, double *d_out, void *stream);

/* Host-buffer convenience: H2D of positions (+types), neighbour build, eval,
 * D2H of the results -- the whole of TensorAlloyCalculator.calculate
 * (calculator.py:335-370) in one call.  Host buffers should be pinned for full
 * copy speed.  rebuild != 0 rebuilds the lists, else tab_nbr_update is used. */
int tab_eam_compute_host(tab_model *model, tab_nbr *nbr, int32_t precision,
                         int32_t n, const double *h_pos, const int32_t *h_types,
                         const double *h_cell, const int32_t *h_pbc, double rc,
                         int32_t rebuild, double *h_energy, double *h_eatom,
                         double *h_forces, double *h_virial, void *stream);

/* ------------------------------------------------------------------------
 * AtomicNN: Behler symmetry functions (G2 + G4) + per-element MLPs.
 * Replaces SymmetryFunction.calculate (nn/atomic/sf.py:79-215), the triple
 * enumeration of get_angular_metadata (transformer/universal.py:115-233),
 * AtomicNN._get_model_outputs / _apply_minmax_normalization / _get_energy_ops
 * (nn/atomic/atomic.py:157-302), convolution1x1 (nn/convolutional.py:154-300) and
 * the autograd outputs of BasicNN.build (nn/basic.py:276-354).
 * ---------------------------------------------------------------------- */
typedef struct tab_atomic tab_atomic;

#define TAB_CUTOFF_COSINE     0   /* nn/cutoff.py:20-48 */
#define TAB_CUTOFF_POLYNOMIAL 1   /* nn/cutoff.py:51-85, gamma = 5 */

/* activation ids (nn/utils.py:50-74) */
#define TAB_ACT_SOFTPLUS   0
#define TAB_ACT_TANH       1
#define TAB_ACT_RELU       2
#define TAB_ACT_LEAKY_RELU 3
#define TAB_ACT_SIGMOID    4
#define TAB_ACT_SOFTSIGN   5
#define TAB_ACT_ELU        6
#define TAB_ACT_SQUAREPLUS 7

/* Symmetry-function hyper-parameters.  The radial sets are the (eta, omega)
 * pairs in sklearn ParameterGrid order (eta outer, omega inner), the angular sets
 * the (beta, gamma, zeta) triples (beta outer, gamma, zeta inner) -- sf.py:47-51.
 * Feature layout of a centre of element c: G2 blocks of the terms
 * [cc, c-x1, ...] then G4 blocks of c+(sorted pair j<=k) -- utils.py:262-286. */
#define TAB_RADIAL_SF      0
#define TAB_RADIAL_MORSE   1
#define TAB_RADIAL_DENSITY 2
#define TAB_RADIAL_PEXP    3
#define TAB_GRAP_SIGNED_SQRT_M0 1
#define TAB_GRAP_TRACELESS      2

typedef struct tab_sf_desc {
    int32_t n_el;
    int32_t cutoff;           /* TAB_CUTOFF_* */
    int32_t angular;
    int32_t n_r, n_a;
    double  rc, acut;
    const double *eta, *omega;             /* [n_r] */
    const double *beta, *gamma, *zeta;     /* [n_a] */
    /* GenericRadialAtomicPotential (nn/atomic/grap.py:121-466 legacy mode, :470-680 new mode): the
     * radial family and its multipole moments.  Parameters 1, 2 of set tau live in
     * eta[tau], omega[tau], parameter 3 in p3[tau]:
     *   TAB_RADIAL_SF      exp(-eta (r-omega)^2 / rc^2)          (= Behler G2)
     *   TAB_RADIAL_MORSE   D [e^{-2g(r-r0)} - 2 e^{-g(r-r0)}]    (D, gamma, r0)
     *   TAB_RADIAL_DENSITY A exp(-beta (r/re - 1))               (A, beta, re)
     *   TAB_RADIAL_PEXP    exp(-(r/rl)^pl)                       (rl, pl)
     * moments[0..n_moments): subset of {0, 1, 2, 3}; feature layout per term:
     * [tau][moment].  n_moments == 0 means {0} (plain symmetry functions).
     * grap_flags (new mode of the reference, grap.py:596-680):
     *   TAB_GRAP_SIGNED_SQRT_M0  the m = 0 entry is sign(P) sqrt(P^2 + 1e-16) (grap.py:667-676)
     *   TAB_GRAP_TRACELESS       `symmetric=True` multiplicity tensor (grap.py:485-494):
     *                            m = 2 minus P_0^2 / 3, m = 3 minus 3/5 sum_a P_a^2;
     *                            needs moments = 0..max */
    int32_t radial_kind;
    int32_t n_moments;
    int32_t moments[4];
    int32_t grap_flags;
    const double *p3;                      /* [n_r] or NULL */
} tab_sf_desc;

/* One element's MLP: sizes[0] = descriptor length, sizes[n_layers] = 1;
 * weights[l] is [sizes[l], sizes[l+1]] row-major (the reference's Conv1d kernel
 * [1, in, out]), biases[l] [sizes[l+1]] (NULL = zeros); layer n_layers-1 is the
 * linear 'Output' layer (bias used iff output_bias).  xlo/xhi: min-max
 * normalisation vectors (atomic.py:181-195) or NULL. */
typedef struct tab_mlp_desc {
    int32_t n_layers;
    int32_t sizes[9];
    int32_t activation;       /* TAB_ACT_* */
    int32_t use_resnet_dt;
    int32_t output_bias;
    const double *weights[8];
    const double *biases[8];
    const double *xlo, *xhi;
} tab_mlp_desc;

int tab_atomic_create(tab_atomic **out, const tab_sf_desc *sf, const tab_mlp_desc *mlps);
int tab_atomic_free(tab_atomic *model);
int tab_atomic_dim(const tab_atomic *model);   /* descriptor length per atom */

/* E, per-atom E, forces, virial as tab_eam_eval.  The lists must have been built
 * with a cutoff >= max(rc, acut). */
int tab_atomic_eval(tab_atomic *model, tab_nbr *nbr, int32_t precision,
                    double *d_energy, double *d_eatom, double *d_forces,
                    double *d_virial, void *stream);

/* Spatial decomposition of AtomicNN.  The force on an owned atom i holds
 * dE_j/dR_i of every centre j within rc, and with angular functions dE_j/dR_i needs the
 * whole neighbourhood of j: the lists are built (tab_nbr_build_dd) with
 *   "owned" group = the rank's own atoms followed by the INNER halo (atoms of other ranks
 *                   within rc of the slab), whose descriptors / MLP / per-pair gradients are
 *                   recomputed redundantly on this rank,
 *   halo group    = the OUTER halo (rc .. 2 rc), positions only,
 * so ONE position exchange per step suffices (no exchange of dE/dG or ghost forces).
 * d_mask [n_owned_group] int32, caller order: 1 = own atom, 0 = inner halo.  d_energy /
 * d_virial receive this rank's partial sums over the own atoms; d_forces / d_eatom
 * [n_owned_group, .] are valid in the rows of the own atoms. */
int tab_atomic_eval_dd(tab_atomic *model, tab_nbr *nbr, int32_t precision,
                       const int32_t *d_mask, double *d_energy, double *d_eatom,
                       double *d_forces, double *d_virial, void *stream);

/* Raw descriptors, d_desc [n, dim] float64 in caller atom order. */
int tab_atomic_descriptors(tab_atomic *model, tab_nbr *nbr, int32_t precision,
                           double *d_desc, void *stream);

/* Training support ("PyTorch custom ops where tensors cross into training").
 * With c = dE/dG [n, dim] (caller order) the forces and the virial are LINEAR in c:
 *   tab_atomic_forces : F = J(R)^T c , W = sum_p g_p (x) D_p        (the forward op)
 *   tab_atomic_jvp    : T = d/dc [ sum_a F_a.u_a + sum_ab A_ab W_ab ]  (its transpose;
 *                       d_u [n,3], d_A [9] on the device, d_out [n, dim])
 * so a force / stress loss back-propagates to the network parameters through c --
 * what the reference obtains from TF second-order autograd (nn/opt.py:132-157).  Symmetry
 * functions and the GRAP families with moments 0 .. 3 (the moment sums of the lists are
 * recomputed inside the call); single-structure and batch handles (d_virial / d_A per
 * structure). */
int tab_atomic_forces(tab_atomic *model, tab_nbr *nbr, int32_t precision,
                      const double *d_dedg, double *d_forces, double *d_virial,
                      void *stream);
int tab_atomic_jvp(tab_atomic *model, tab_nbr *nbr, int32_t precision,
                   const double *d_u, const double *d_A, double *d_out, void *stream);

/* Pair-level training operators of the EAM / ADP family ("PyTorch custom ops where
 * tensors cross into training").  The reference obtains d loss / d parameters of an energy
 * + force + stress loss from TF second-order autograd over its padded pair tensors
 * (nn/eam/eam.py:495-570, nn/basic.py:446-631, nn/opt.py:132-157).  Here the energy is a
 * function of the directed pair vectors D_p (one per list entry, rows sorted by centre
 * atom, tab_nbr_export order); the force / virial assembly is LINEAR in g_p = dE/dD_p:
 *   tab_pairs_export : d_i, d_j [nij] int32 (caller indices, either may be NULL),
 *                      d_D [nij,3] = R_j - R_i (+ image shift)
 *   tab_pair_forces  : F_i = sum_{p in row i} g_p - sum_{p -> i} g_p   [n,3]
 *                      W   = sum_p sym(g_p (x) D_p)   [9], batch handles [n_struct, 9]
 *   tab_pair_jvp     : the transpose, t_p = u_i - u_j + sym(A) D_p     [nij,3]
 * All arrays float64 on the device; single-structure and batch handles. */
int tab_pairs_export(tab_nbr *nbr, int32_t *d_i, int32_t *d_j, double *d_D, void *stream);
int tab_pair_forces(tab_nbr *nbr, const double *d_g, double *d_forces, double *d_virial,
                    void *stream);
int tab_pair_jvp(tab_nbr *nbr, const double *d_u, const double *d_A, double *d_t,
                 void *stream);

/* Per-kernel timing of tab_eam_eval with CUDA events recorded on the launching
 * stream (used by bench.py for the roofline figures; off by default).
 * tab_profile_read synchronises the device; ms[0..3] = mean milliseconds of the
 * rho pass, the F' spread, the force pass and the final reduction. */
int tab_profile_enable(int32_t on);
int tab_profile_read(double *ms, int32_t *calls);

/* number of kernel launches the library has enqueued since load / last reset
 * (feeds bench.py's `gpu_launches`). */
int64_t tab_launch_count(void);
void    tab_launch_count_reset(void);

/* Temperature-dependent heads of the finite-temperature AtomicNN in one kernel
 * (replaces the three 1x1-convolution chains and the tf.gradients call of
 * nn/atomic/finite_temperature.py:92-304; entropy models: 'default', 'Sommerfeld' :120-166,
 * and the beryllium free-electron form nn/atomic/special/beryllium.py:23-77).
 * Per element e: H[e] maps the descriptors (sizes[0] = dim; xlo / xhi of H[e] = min-max map)
 * to nH values through a linear last layer; S[e] and U[e] map [H, T] (nH + 1 values) to one
 * value.  Layer conventions as tab_mlp_desc (biases[n_layers - 1] + output_bias = bias of the
 * last layer).  algo: 0 default (S = s), 1 Sommerfeld (S = s T); special: 0 none, 1 Be. */
typedef struct tab_td tab_td;
typedef struct tab_td_desc {
    int32_t n_elements;
    int32_t dim;
    int32_t algo;
    int32_t special;
    const tab_mlp_desc *H, *S, *U;        /* [n_elements] each */
} tab_td_desc;

int tab_td_create(tab_td **out, const tab_td_desc *desc);
int tab_td_free(tab_td *td);

/* d_types int32 [n] (element index), d_G [n, dim] descriptors (tab_atomic_descriptors),
 * d_T [n] electron temperature of every atom (eV).  Outputs, float64 on the device, caller
 * atom order: d_U, d_S, d_F [n] (internal energy, electron entropy, free energy F = U - T S
 * per atom) and d_dFdG [n, dim] -- the input of tab_atomic_forces, which yields the forces and
 * the virial of the FREE energy (basic.py:190-202). */
int tab_td_eval(tab_td *td, int32_t n, const int32_t *d_types, const double *d_G,
                const double *d_T, int32_t precision, double *d_U, double *d_S, double *d_F,
                double *d_dFdG, void *stream);
/* 0 = every weight transfer (TMA bulk copy) of the evaluations so far completed; 1 = one
 * timed out and the outputs of that call hold NaN.  Synchronises the stream. */
int tab_td_status(tab_td *td, int32_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TAB200_H */
