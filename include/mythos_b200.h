/* mythos_b200.h -- plain C-ABI of the B200-native oxDNA-family energy / force / theta-gradient path.
 *
 * This is the drop-in boundary: every entry point takes device pointers, sizes and a CUDA stream, enqueues
 * work on that stream and returns; no host synchronisation, no allocation inside a call (scratch comes from a
 * caller-provided workspace), no torch / XLA types in any signature.  The XLA-FFI shim
 * (mythos_b200/csrc/xla_ffi_shim.cc) and the ctypes host layer (mythos_b200/_lib.py) are thin adapters over it.
 *
 * Reference interfaces each entry point stands in for (paths relative to the reference repo):
 *   mythos_b200_energy_*        ComposedEnergyFunction.compute_terms / __call__  mythos/energy/base.py:312-319
 *                               (all BaseEnergyFunction.compute_energy of dna1/dna2/rna2/na1),
 *                               EnergyFunction.map (batched frames)               mythos/energy/base.py:90-93,
 *                               and their jax.grad / jax.value_and_grad           simulators/jax_md/jaxmd.py:70-73,
 *                                                                                 optimization/objective.py:235
 *   mythos_b200_nl_build_*      get_neighbor_list_fn -> jax_md.partition.neighbor_list(OrderedSparse)
 *                                                                                 mythos/utils/neighbors.py:12-59
 *   mythos_b200_langevin_*      step_fn of the user-supplied simulator_init (jax_md.simulate.nvt_langevin on
 *                               RigidBody)                                        simulators/jax_md/jaxmd.py:73,82-94
 *   mythos_b200_difftre_*       energy_fn.map(reference_states) + value_and_grad(compute_loss) w.r.t. params
 *                                                                                 optimization/objective.py:224-235,345-364
 *   mythos_b200_weights_neff_*  compute_weights_and_neff                          optimization/objective.py:139-163
 *
 * Conventions: center is (F,N,3), quat is (F,N,4) = (w,x,y,z), row-major, contiguous; integer arrays are int32;
 * `pairs` is (2,capacity) with i = pairs[0][k] < j = pairs[1][k] and padding entries >= N (masked as the
 * reference does with `op_i < N`); per-term outputs use the oxDNA split_energy column order MB_TERM_*.
 */
#ifndef MYTHOS_B200_H
#define MYTHOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_ABI_VERSION 3

/* ---- status codes ------------------------------------------------------------------------------------- */
enum mb_status {
  MB_OK = 0,
  MB_EINVAL_SHAPE = 1, /* a size / pointer combination that cannot be a valid call          */
  MB_EINVAL_MODEL = 2, /* unknown model form / bank count                                    */
  MB_ECAPACITY = 3,    /* workspace or static capacity too small (data-independent)          */
  MB_ECUDA = 4         /* a CUDA runtime call failed; see mythos_b200_last_error()           */
};

/* ---- energy terms (oxDNA split_energy.dat column order) -------------------------------------------------- */
enum mb_term {
  MB_TERM_FENE = 0,
  MB_TERM_BEXC = 1,
  MB_TERM_STACK = 2,
  MB_TERM_UEXC = 3,
  MB_TERM_HB = 4,
  MB_TERM_CROSS = 5,
  MB_TERM_COAX = 6,
  MB_TERM_DEBYE = 7,
  MB_N_TERMS = 8
};
#define MB_BONDED_TERMS 0x07u
#define MB_UNBONDED_TERMS 0xF8u
#define MB_ALL_TERMS 0xFFu

/* ---- kernel-level parameter bank --------------------------------------------------------------------------
 * One bank = MB_P_COUNT reals in the order below ("<term>.<reference parameter name>").  These are the
 * *dependent-inclusive* parameters the reference's interaction functions receive: the
 * theta -> bank chain (init_params, smoothing solvers) stays on the host side of the call.
 * oxDNA1 / oxDNA2 / RNA2 use one bank; NA1 uses three (DNA, RNA, DNA-RNA hybrid) laid out back to back. */
#define MB_PARAM_LIST(X) \
  X(FENE_EPS, "fene.eps_backbone") \
  X(FENE_R0, "fene.r0_backbone") \
  X(FENE_DELTA, "fene.delta_backbone") \
  X(FENE_FMAX, "fene.fmax") \
  X(FENE_FINF, "fene.finf") \
  X(BEXC_EPS, "bonded_excluded_volume.eps_exc") \
  X(BEXC_BASE_RSTAR, "bonded_excluded_volume.dr_star_base") \
  X(BEXC_BASE_SIGMA, "bonded_excluded_volume.sigma_base") \
  X(BEXC_BASE_B, "bonded_excluded_volume.b_base") \
  X(BEXC_BASE_RC, "bonded_excluded_volume.dr_c_base") \
  X(BEXC_BACK_BASE_RSTAR, "bonded_excluded_volume.dr_star_back_base") \
  X(BEXC_BACK_BASE_SIGMA, "bonded_excluded_volume.sigma_back_base") \
  X(BEXC_BACK_BASE_B, "bonded_excluded_volume.b_back_base") \
  X(BEXC_BACK_BASE_RC, "bonded_excluded_volume.dr_c_back_base") \
  X(BEXC_BASE_BACK_RSTAR, "bonded_excluded_volume.dr_star_base_back") \
  X(BEXC_BASE_BACK_SIGMA, "bonded_excluded_volume.sigma_base_back") \
  X(BEXC_BASE_BACK_B, "bonded_excluded_volume.b_base_back") \
  X(BEXC_BASE_BACK_RC, "bonded_excluded_volume.dr_c_base_back") \
  X(UEXC_EPS, "unbonded_excluded_volume.eps_exc") \
  X(UEXC_BASE_RSTAR, "unbonded_excluded_volume.dr_star_base") \
  X(UEXC_BASE_SIGMA, "unbonded_excluded_volume.sigma_base") \
  X(UEXC_BASE_B, "unbonded_excluded_volume.b_base") \
  X(UEXC_BASE_RC, "unbonded_excluded_volume.dr_c_base") \
  X(UEXC_BACK_BASE_RSTAR, "unbonded_excluded_volume.dr_star_back_base") \
  X(UEXC_BACK_BASE_SIGMA, "unbonded_excluded_volume.sigma_back_base") \
  X(UEXC_BACK_BASE_B, "unbonded_excluded_volume.b_back_base") \
  X(UEXC_BACK_BASE_RC, "unbonded_excluded_volume.dr_c_back_base") \
  X(UEXC_BASE_BACK_RSTAR, "unbonded_excluded_volume.dr_star_base_back") \
  X(UEXC_BASE_BACK_SIGMA, "unbonded_excluded_volume.sigma_base_back") \
  X(UEXC_BASE_BACK_B, "unbonded_excluded_volume.b_base_back") \
  X(UEXC_BASE_BACK_RC, "unbonded_excluded_volume.dr_c_base_back") \
  X(UEXC_BACKBONE_RSTAR, "unbonded_excluded_volume.dr_star_backbone") \
  X(UEXC_BACKBONE_SIGMA, "unbonded_excluded_volume.sigma_backbone") \
  X(UEXC_BACKBONE_B, "unbonded_excluded_volume.b_backbone") \
  X(UEXC_BACKBONE_RC, "unbonded_excluded_volume.dr_c_backbone") \
  X(STACK_RLOW, "stacking.dr_low_stack") \
  X(STACK_RHIGH, "stacking.dr_high_stack") \
  X(STACK_RCLOW, "stacking.dr_c_low_stack") \
  X(STACK_RCHIGH, "stacking.dr_c_high_stack") \
  X(STACK_A, "stacking.a_stack") \
  X(STACK_R0, "stacking.dr0_stack") \
  X(STACK_RC, "stacking.dr_c_stack") \
  X(STACK_BLOW, "stacking.b_low_stack") \
  X(STACK_BHIGH, "stacking.b_high_stack") \
  X(STACK_T4_TH0, "stacking.theta0_stack_4") \
  X(STACK_T4_DSTAR, "stacking.delta_theta_star_stack_4") \
  X(STACK_T4_DC, "stacking.delta_theta_stack_4_c") \
  X(STACK_T4_A, "stacking.a_stack_4") \
  X(STACK_T4_B, "stacking.b_stack_4") \
  X(STACK_T5_TH0, "stacking.theta0_stack_5") \
  X(STACK_T5_DSTAR, "stacking.delta_theta_star_stack_5") \
  X(STACK_T5_DC, "stacking.delta_theta_stack_5_c") \
  X(STACK_T5_A, "stacking.a_stack_5") \
  X(STACK_T5_B, "stacking.b_stack_5") \
  X(STACK_T6_TH0, "stacking.theta0_stack_6") \
  X(STACK_T6_DSTAR, "stacking.delta_theta_star_stack_6") \
  X(STACK_T6_DC, "stacking.delta_theta_stack_6_c") \
  X(STACK_T6_A, "stacking.a_stack_6") \
  X(STACK_T6_B, "stacking.b_stack_6") \
  X(STACK_T9_TH0, "stacking.theta0_stack_9") \
  X(STACK_T9_DSTAR, "stacking.delta_theta_star_stack_9") \
  X(STACK_T9_DC, "stacking.delta_theta_stack_9_c") \
  X(STACK_T9_A, "stacking.a_stack_9") \
  X(STACK_T9_B, "stacking.b_stack_9") \
  X(STACK_T10_TH0, "stacking.theta0_stack_10") \
  X(STACK_T10_DSTAR, "stacking.delta_theta_star_stack_10") \
  X(STACK_T10_DC, "stacking.delta_theta_stack_10_c") \
  X(STACK_T10_A, "stacking.a_stack_10") \
  X(STACK_T10_B, "stacking.b_stack_10") \
  X(STACK_PHI1_XSTAR, "stacking.neg_cos_phi1_star_stack") \
  X(STACK_PHI1_XC, "stacking.neg_cos_phi1_c_stack") \
  X(STACK_PHI1_A, "stacking.a_stack_1") \
  X(STACK_PHI1_B, "stacking.b_neg_cos_phi1_stack") \
  X(STACK_PHI2_XSTAR, "stacking.neg_cos_phi2_star_stack") \
  X(STACK_PHI2_XC, "stacking.neg_cos_phi2_c_stack") \
  X(STACK_PHI2_A, "stacking.a_stack_2") \
  X(STACK_PHI2_B, "stacking.b_neg_cos_phi2_stack") \
  X(STACK_W00, "stacking.eps_stack[0,0]") \
  X(STACK_W01, "stacking.eps_stack[0,1]") \
  X(STACK_W02, "stacking.eps_stack[0,2]") \
  X(STACK_W03, "stacking.eps_stack[0,3]") \
  X(STACK_W10, "stacking.eps_stack[1,0]") \
  X(STACK_W11, "stacking.eps_stack[1,1]") \
  X(STACK_W12, "stacking.eps_stack[1,2]") \
  X(STACK_W13, "stacking.eps_stack[1,3]") \
  X(STACK_W20, "stacking.eps_stack[2,0]") \
  X(STACK_W21, "stacking.eps_stack[2,1]") \
  X(STACK_W22, "stacking.eps_stack[2,2]") \
  X(STACK_W23, "stacking.eps_stack[2,3]") \
  X(STACK_W30, "stacking.eps_stack[3,0]") \
  X(STACK_W31, "stacking.eps_stack[3,1]") \
  X(STACK_W32, "stacking.eps_stack[3,2]") \
  X(STACK_W33, "stacking.eps_stack[3,3]") \
  X(HB_RLOW, "hydrogen_bonding.dr_low_hb") \
  X(HB_RHIGH, "hydrogen_bonding.dr_high_hb") \
  X(HB_RCLOW, "hydrogen_bonding.dr_c_low_hb") \
  X(HB_RCHIGH, "hydrogen_bonding.dr_c_high_hb") \
  X(HB_A, "hydrogen_bonding.a_hb") \
  X(HB_R0, "hydrogen_bonding.dr0_hb") \
  X(HB_RC, "hydrogen_bonding.dr_c_hb") \
  X(HB_BLOW, "hydrogen_bonding.b_low_hb") \
  X(HB_BHIGH, "hydrogen_bonding.b_high_hb") \
  X(HB_T1_TH0, "hydrogen_bonding.theta0_hb_1") \
  X(HB_T1_DSTAR, "hydrogen_bonding.delta_theta_star_hb_1") \
  X(HB_T1_DC, "hydrogen_bonding.delta_theta_hb_1_c") \
  X(HB_T1_A, "hydrogen_bonding.a_hb_1") \
  X(HB_T1_B, "hydrogen_bonding.b_hb_1") \
  X(HB_T2_TH0, "hydrogen_bonding.theta0_hb_2") \
  X(HB_T2_DSTAR, "hydrogen_bonding.delta_theta_star_hb_2") \
  X(HB_T2_DC, "hydrogen_bonding.delta_theta_hb_2_c") \
  X(HB_T2_A, "hydrogen_bonding.a_hb_2") \
  X(HB_T2_B, "hydrogen_bonding.b_hb_2") \
  X(HB_T3_TH0, "hydrogen_bonding.theta0_hb_3") \
  X(HB_T3_DSTAR, "hydrogen_bonding.delta_theta_star_hb_3") \
  X(HB_T3_DC, "hydrogen_bonding.delta_theta_hb_3_c") \
  X(HB_T3_A, "hydrogen_bonding.a_hb_3") \
  X(HB_T3_B, "hydrogen_bonding.b_hb_3") \
  X(HB_T4_TH0, "hydrogen_bonding.theta0_hb_4") \
  X(HB_T4_DSTAR, "hydrogen_bonding.delta_theta_star_hb_4") \
  X(HB_T4_DC, "hydrogen_bonding.delta_theta_hb_4_c") \
  X(HB_T4_A, "hydrogen_bonding.a_hb_4") \
  X(HB_T4_B, "hydrogen_bonding.b_hb_4") \
  X(HB_T7_TH0, "hydrogen_bonding.theta0_hb_7") \
  X(HB_T7_DSTAR, "hydrogen_bonding.delta_theta_star_hb_7") \
  X(HB_T7_DC, "hydrogen_bonding.delta_theta_hb_7_c") \
  X(HB_T7_A, "hydrogen_bonding.a_hb_7") \
  X(HB_T7_B, "hydrogen_bonding.b_hb_7") \
  X(HB_T8_TH0, "hydrogen_bonding.theta0_hb_8") \
  X(HB_T8_DSTAR, "hydrogen_bonding.delta_theta_star_hb_8") \
  X(HB_T8_DC, "hydrogen_bonding.delta_theta_hb_8_c") \
  X(HB_T8_A, "hydrogen_bonding.a_hb_8") \
  X(HB_T8_B, "hydrogen_bonding.b_hb_8") \
  X(HB_W00, "hydrogen_bonding.eps_hb_weights[0,0]") \
  X(HB_W01, "hydrogen_bonding.eps_hb_weights[0,1]") \
  X(HB_W02, "hydrogen_bonding.eps_hb_weights[0,2]") \
  X(HB_W03, "hydrogen_bonding.eps_hb_weights[0,3]") \
  X(HB_W10, "hydrogen_bonding.eps_hb_weights[1,0]") \
  X(HB_W11, "hydrogen_bonding.eps_hb_weights[1,1]") \
  X(HB_W12, "hydrogen_bonding.eps_hb_weights[1,2]") \
  X(HB_W13, "hydrogen_bonding.eps_hb_weights[1,3]") \
  X(HB_W20, "hydrogen_bonding.eps_hb_weights[2,0]") \
  X(HB_W21, "hydrogen_bonding.eps_hb_weights[2,1]") \
  X(HB_W22, "hydrogen_bonding.eps_hb_weights[2,2]") \
  X(HB_W23, "hydrogen_bonding.eps_hb_weights[2,3]") \
  X(HB_W30, "hydrogen_bonding.eps_hb_weights[3,0]") \
  X(HB_W31, "hydrogen_bonding.eps_hb_weights[3,1]") \
  X(HB_W32, "hydrogen_bonding.eps_hb_weights[3,2]") \
  X(HB_W33, "hydrogen_bonding.eps_hb_weights[3,3]") \
  X(CROSS_RLOW, "cross_stacking.dr_low_cross") \
  X(CROSS_RHIGH, "cross_stacking.dr_high_cross") \
  X(CROSS_RCLOW, "cross_stacking.dr_c_low_cross") \
  X(CROSS_RCHIGH, "cross_stacking.dr_c_high_cross") \
  X(CROSS_K, "cross_stacking.k_cross") \
  X(CROSS_R0, "cross_stacking.r0_cross") \
  X(CROSS_RC, "cross_stacking.dr_c_cross") \
  X(CROSS_BLOW, "cross_stacking.b_low_cross") \
  X(CROSS_BHIGH, "cross_stacking.b_high_cross") \
  X(CROSS_T1_TH0, "cross_stacking.theta0_cross_1") \
  X(CROSS_T1_DSTAR, "cross_stacking.delta_theta_star_cross_1") \
  X(CROSS_T1_DC, "cross_stacking.delta_theta_cross_1_c") \
  X(CROSS_T1_A, "cross_stacking.a_cross_1") \
  X(CROSS_T1_B, "cross_stacking.b_cross_1") \
  X(CROSS_T2_TH0, "cross_stacking.theta0_cross_2") \
  X(CROSS_T2_DSTAR, "cross_stacking.delta_theta_star_cross_2") \
  X(CROSS_T2_DC, "cross_stacking.delta_theta_cross_2_c") \
  X(CROSS_T2_A, "cross_stacking.a_cross_2") \
  X(CROSS_T2_B, "cross_stacking.b_cross_2") \
  X(CROSS_T3_TH0, "cross_stacking.theta0_cross_3") \
  X(CROSS_T3_DSTAR, "cross_stacking.delta_theta_star_cross_3") \
  X(CROSS_T3_DC, "cross_stacking.delta_theta_cross_3_c") \
  X(CROSS_T3_A, "cross_stacking.a_cross_3") \
  X(CROSS_T3_B, "cross_stacking.b_cross_3") \
  X(CROSS_T4_TH0, "cross_stacking.theta0_cross_4") \
  X(CROSS_T4_DSTAR, "cross_stacking.delta_theta_star_cross_4") \
  X(CROSS_T4_DC, "cross_stacking.delta_theta_cross_4_c") \
  X(CROSS_T4_A, "cross_stacking.a_cross_4") \
  X(CROSS_T4_B, "cross_stacking.b_cross_4") \
  X(CROSS_T7_TH0, "cross_stacking.theta0_cross_7") \
  X(CROSS_T7_DSTAR, "cross_stacking.delta_theta_star_cross_7") \
  X(CROSS_T7_DC, "cross_stacking.delta_theta_cross_7_c") \
  X(CROSS_T7_A, "cross_stacking.a_cross_7") \
  X(CROSS_T7_B, "cross_stacking.b_cross_7") \
  X(CROSS_T8_TH0, "cross_stacking.theta0_cross_8") \
  X(CROSS_T8_DSTAR, "cross_stacking.delta_theta_star_cross_8") \
  X(CROSS_T8_DC, "cross_stacking.delta_theta_cross_8_c") \
  X(CROSS_T8_A, "cross_stacking.a_cross_8") \
  X(CROSS_T8_B, "cross_stacking.b_cross_8") \
  X(COAX_RLOW, "coaxial_stacking.dr_low_coax") \
  X(COAX_RHIGH, "coaxial_stacking.dr_high_coax") \
  X(COAX_RCLOW, "coaxial_stacking.dr_c_low_coax") \
  X(COAX_RCHIGH, "coaxial_stacking.dr_c_high_coax") \
  X(COAX_K, "coaxial_stacking.k_coax") \
  X(COAX_R0, "coaxial_stacking.dr0_coax") \
  X(COAX_RC, "coaxial_stacking.dr_c_coax") \
  X(COAX_BLOW, "coaxial_stacking.b_low_coax") \
  X(COAX_BHIGH, "coaxial_stacking.b_high_coax") \
  X(COAX_T4_TH0, "coaxial_stacking.theta0_coax_4") \
  X(COAX_T4_DSTAR, "coaxial_stacking.delta_theta_star_coax_4") \
  X(COAX_T4_DC, "coaxial_stacking.delta_theta_coax_4_c") \
  X(COAX_T4_A, "coaxial_stacking.a_coax_4") \
  X(COAX_T4_B, "coaxial_stacking.b_coax_4") \
  X(COAX_T1_TH0, "coaxial_stacking.theta0_coax_1") \
  X(COAX_T1_DSTAR, "coaxial_stacking.delta_theta_star_coax_1") \
  X(COAX_T1_DC, "coaxial_stacking.delta_theta_coax_1_c") \
  X(COAX_T1_A, "coaxial_stacking.a_coax_1") \
  X(COAX_T1_B, "coaxial_stacking.b_coax_1") \
  X(COAX_T5_TH0, "coaxial_stacking.theta0_coax_5") \
  X(COAX_T5_DSTAR, "coaxial_stacking.delta_theta_star_coax_5") \
  X(COAX_T5_DC, "coaxial_stacking.delta_theta_coax_5_c") \
  X(COAX_T5_A, "coaxial_stacking.a_coax_5") \
  X(COAX_T5_B, "coaxial_stacking.b_coax_5") \
  X(COAX_T6_TH0, "coaxial_stacking.theta0_coax_6") \
  X(COAX_T6_DSTAR, "coaxial_stacking.delta_theta_star_coax_6") \
  X(COAX_T6_DC, "coaxial_stacking.delta_theta_coax_6_c") \
  X(COAX_T6_A, "coaxial_stacking.a_coax_6") \
  X(COAX_T6_B, "coaxial_stacking.b_coax_6") \
  X(COAX_PHI3_XSTAR, "coaxial_stacking.cos_phi3_star_coax") \
  X(COAX_PHI3_XC, "coaxial_stacking.cos_phi3_c_coax") \
  X(COAX_PHI3_A, "coaxial_stacking.a_coax_3p") \
  X(COAX_PHI3_B, "coaxial_stacking.b_cos_phi3_coax") \
  X(COAX_PHI4_XSTAR, "coaxial_stacking.cos_phi4_star_coax") \
  X(COAX_PHI4_XC, "coaxial_stacking.cos_phi4_c_coax") \
  X(COAX_PHI4_A, "coaxial_stacking.a_coax_4p") \
  X(COAX_PHI4_B, "coaxial_stacking.b_cos_phi4_coax") \
  X(COAX_F6_A, "coaxial_stacking.a_coax_1_f6") \
  X(COAX_F6_B, "coaxial_stacking.b_coax_1_f6") \
  X(DEBYE_KAPPA, "debye.kappa") \
  X(DEBYE_PREF, "debye.prefactor") \
  X(DEBYE_SMOOTH, "debye.smoothing_coeff") \
  X(DEBYE_RCUT, "debye.r_cut") \
  X(DEBYE_RHIGH, "debye.r_high")

enum mb_param_index {
#define MB_X_ENUM(id, name) MB_P_##id,
  MB_PARAM_LIST(MB_X_ENUM)
#undef MB_X_ENUM
  MB_P_COUNT_RAW
};
#define MB_P_COUNT 232 /* MB_P_COUNT_RAW (231) padded to a multiple of 8 */
#define MB_MAX_BANKS 3
enum mb_bank { MB_BANK_DNA = 0, MB_BANK_RNA = 1, MB_BANK_DRH = 2 };

/* ---- model description (not differentiated; host memory, copied by value at launch) ---------------------- */
typedef struct mb_flavour_geom {
  double back[3];     /* backbone site  = c + back[0]*a1 + back[1]*a2 + back[2]*a3                          */
  double back_stack;  /* a1 offset of the backbone site the stacking term uses when use_back_stack (oxDNA2) */
  double stack;       /* stacking site  = c + stack*a1                                                      */
  double base;        /* base (HB) site = c + base*a1                                                       */
  double stack3[2];   /* RNA 3' stacking site = c + stack3[0]*a1 + stack3[1]*a2                             */
  double stack5[2];   /* RNA 5' stacking site                                                               */
  double p3[3];       /* RNA backbone direction vectors (a1,a2,a3 coefficients)                             */
  double p5[3];
  int32_t use_back_stack;
  int32_t _pad;
} mb_flavour_geom;

enum mb_stack_form { MB_STACK_DNA = 0, MB_STACK_RNA = 1 };  /* theta4,5,6 | theta5,6,9,10                   */
enum mb_cross_form { MB_CROSS_DNA1 = 0, MB_CROSS_RNA2 = 1 }; /* with | without the theta4 factor             */
enum mb_coax_form { MB_COAX_DNA1 = 0, MB_COAX_DNA2 = 1 };    /* phi3,phi4 + f4(2pi-theta1) | f6(theta1)      */

typedef struct mb_bank_forms {
  int32_t stack_form, cross_form, coax_form, has_debye;
} mb_bank_forms;

typedef struct mb_model {
  int32_t n_banks;            /* 1 (dna1, dna2, rna2) or 3 (na1: DNA, RNA, DRH)                              */
  int32_t half_charged_ends;  /* Debye: halve the charge of strand-end nucleotides                           */
  mb_flavour_geom geom[2];    /* [0] the only flavour, or DNA; [1] RNA (na1 only)                            */
  mb_bank_forms forms[MB_MAX_BANKS];
  double box[3];              /* periodic box; all zero = free space                                         */
} mb_model;

/* ---- energy / forces / parameter gradient ----------------------------------------------------------------- */
/* ---- per-frame structural observables -------------------------------------------------------
 * Stand in for mythos/observables/propeller.py:18-71 (PropellerTwist), rise.py:19-70 (Rise), pitch.py:32-88
 * (PitchAngle), diameter.py:21-76 (Diameter) and base.py:24-44 (local_helical_axis): per frame, the MEAN over the listed
 * base pairs / quartets of
 *   MB_OBS_PROPELLER    180 - acos(clamp(a3_i . a3_j)) * 180/pi            [degrees]       (base pairs)
 *   MB_OBS_DIAMETER     (|disp(back_i, back_j)| + sigma_backbone) * 8.518   [Angstrom]      (base pairs)
 *   MB_OBS_RISE         dot(dr, dr/|dr|) * 8.518, dr = disp(midp2, midp1)   [Angstrom]      (quartets)
 *   MB_OBS_PITCH_ANGLE  acos(clamp(p1 . p2)), p = backbone-backbone vector of a base pair with its component along the
 *                       local helical axis removed, normalised              [radians]       (quartets)
 * with midp = midpoint of the two base sites of a base pair; sites from the model's flavour geometry (nt_type picks the
 * flavour for the 3-bank model).  An empty list gives NaN in its columns (mean of nothing), as jnp.mean does. */
enum mb_observable { MB_OBS_PROPELLER = 0, MB_OBS_RISE = 1, MB_OBS_PITCH_ANGLE = 2, MB_OBS_DIAMETER = 3, MB_N_OBS = 4 };
typedef struct mb_observable_spec {
  const int32_t* base_pairs; /* (P,2) hydrogen-bonded nucleotide pairs, device pointer */
  int32_t n_base_pairs;
  int32_t n_quartets;
  const int32_t* quartets;   /* (Q,4) = (a1, b1, a2, b2): two adjacent base pairs (a1-b1), (a2-b2), device pointer */
  double sigma_backbone;     /* excluded-volume distance added to the diameter (diameter.py:40) */
} mb_observable_spec;
/* standalone: center (F,N,3), quat (F,N,4) -> out (F, MB_N_OBS); one block per frame, HBM-bound (reads only the listed
 * nucleotides) */
int mythos_b200_observables_f64(void* cuda_stream, const mb_model* model, int32_t n, int32_t n_frames, const void* center,
                                const void* quat, const int32_t* nt_type, const mb_observable_spec* spec, void* out);
int mythos_b200_observables_f32(void* cuda_stream, const mb_model* model, int32_t n, int32_t n_frames, const void* center,
                                const void* quat, const int32_t* nt_type, const mb_observable_spec* spec, void* out);

/* ---- probabilistic sequences ----------------------------------------------------------------
 * Stand in for `compute_seq_dep_weight` (mythos/energy/utils.py:45-132) as used by Stacking.pseq_weights
 * (dna1/stacking.py:261-287) and HydrogenBonding.weight (dna1/hydrogen_bonding.py:308-335): the sequence weight of a
 * pair is the expectation of the term's 4x4 table under a distribution in which nucleotides are independent except
 * the two members of one base pair.  The host reduces (unpaired_pseq, bp_pseq, SequenceConstraints) to per-nucleotide
 * marginals and per-base-pair same-pair expectations (differentiable there); the kernel evaluates
 *   w(i,j) = same_w[bp][within_i]                      if i and j are the two members of base pair bp
 *          = sum_ab pmarg[i][a] W[a][b] pmarg[j][b]    otherwise
 * and returns the gradients with respect to W (in d_params), pmarg and same_w.  Generic pair kernel only
 * (MB_FLAG_GENERIC_KERNEL, explicit pair list, single-bank models); gradient outputs are ACCUMULATED into (zero them). */
typedef struct mb_pseq {
  const void* pmarg;          /* (N,4) reals */
  const int32_t* bp_of;       /* (N) base-pair index, -1 = unpaired */
  const int32_t* within;      /* (N) 0/1 position inside its base pair */
  const void* same_w_stack;   /* (n_bp,2) reals, may be NULL if n_bp == 0 */
  const void* same_w_hb;      /* (n_bp,2) */
  void* d_pmarg;              /* out (N,4) or NULL */
  void* d_same_w_stack;       /* out (n_bp,2) or NULL */
  void* d_same_w_hb;          /* out (n_bp,2) or NULL */
  uint32_t terms;             /* which terms take their weight from the distribution: bits MB_TERM_STACK, MB_TERM_HB */
  uint32_t _pad;
} mb_pseq;

typedef struct mb_energy_args {
  const mb_model* model;
  int32_t n;              /* nucleotides per frame                                                           */
  int32_t n_frames;       /* F (1 for a single rigid body)                                                   */
  const void* center;     /* (F,N,3)                                                                         */
  const void* quat;       /* (F,N,4)                                                                         */
  const int32_t* seq;     /* (N) 0..3                                                                        */
  const int32_t* nt_type; /* (N) 1 = DNA, 2 = RNA; may be NULL when n_banks == 1                             */
  const int32_t* nt_type_stack; /* (N) nt_type seen by the stacking term; NULL = nt_type                     */
  const int32_t* is_end;  /* (N) strand-end flags; may be NULL when !half_charged_ends                       */
  const int32_t* bonded;  /* (B,2)                                                                           */
  int32_t n_bonded;
  const int32_t* pairs;   /* (2,capacity) per list                                                           */
  int64_t pair_capacity;
  int64_t pair_frame_stride; /* elements between the lists of consecutive frames; 0 = one list for all frames */
  const void* params;     /* (n_banks*MB_P_COUNT) reals                                                      */
  const void* cot;        /* (F,8) cotangent / weight of each term; NULL = ones; used by the gradient outputs */
  uint32_t term_mask;     /* which MB_TERM_* to evaluate                                                     */
  uint32_t flags;         /* MB_FLAG_*                                                                       */
  void* terms;            /* out (F,8), or NULL                                                              */
  void* d_center;         /* out (F,N,3) = d(sum_t cot_t E_t)/d center, or NULL                              */
  void* d_quat;           /* out (F,N,4), or NULL                                                            */
  void* d_params;         /* out (n_banks*MB_P_COUNT) per row, or NULL                                       */
  int64_t d_params_frame_stride; /* elements between rows of consecutive frames; 0 = one row summed over frames */
  const int32_t* pair_count; /* (F) number of valid pairs at the head of each frame's list (as mb_nl_args.count), or NULL */
  double all_pairs_cutoff;   /* > 0: all-pairs semantics (topology.unbonded_neighbors, topology.py:186-190): `pairs` is
                              * ignored and every non-bonded i<j whose centres are closer than this is evaluated, found by a
                              * shared-memory cell list inside the kernel.  The caller passes the interaction range of its
                              * parameters (all terms have compact support).  MB_ECAPACITY if the frame-resident kernel
                              * does not apply (3 banks, position gradients requested, frame too large for shared memory) */
  void* workspace;           /* scratch of mythos_b200_energy_workspace_bytes(): lets explicit pair lists run through the
                              * phase-queued list kernels (per-nucleotide records, backbone-site gradient buffer, short-range list;
                              * NULL or too small = the one-thread-per-pair kernels) and holds the frame-resident kernel's
                              * parameter-gradient images (REQUIRED when d_params is requested on that route: MB_ECAPACITY
                              * otherwise).  Needs no initialisation.                                                         */
  size_t workspace_bytes;
  const int32_t* pair_split; /* MB_FLAG_TAGGED_PAIRS with the list kernels: (F) entries of each list before this index are
                              * the short-range pairs, the rest (up to pair_count) the Debye pairs; else NULL              */
  const mb_observable_spec* observables; /* optional: also evaluate the per-frame observables of these frames in this call.
                              * On the frame-resident route they are an epilogue of the SAME kernel (the frame is already in
                              * shared memory); on the other routes the standalone kernel is enqueued behind the energy kernels */
  void* observables_out;     /* out (F, MB_N_OBS) when `observables` is set */
  const mb_pseq* pseq;       /* optional: probabilistic sequence weights for stacking and hydrogen bonding */
} mb_energy_args;
#define MB_FLAG_ACCUMULATE 0x1u /* add into the outputs instead of zeroing them first */
#define MB_FLAG_GENERIC_KERNEL 0x2u /* force the one-thread-per-pair kernels even where the frame-resident kernel applies */
#define MB_FLAG_TAGGED_PAIRS 0x8u /* `pairs` comes from mythos_b200_nl_build_* with MB_NL_TAG_SUPPORTS (frame-resident kernel, or the list kernels with pair_split) */
#define MB_FLAG_LIST_KERNEL 0x4u /* explicit lists with forces / several banks: use the phase-queued list kernels even for short lists (needs workspace) */

size_t mythos_b200_energy_workspace_bytes(int32_t n, int32_t n_frames, int64_t pair_capacity, int32_t real_bytes /* 4 or 8 */);
int mythos_b200_frame_kernel_fits(int32_t n, int32_t real_bytes, int32_t want_params); /* 1 if a frame of n nucleotides fits the
                                                  * frame-resident kernel's shared memory (single bank, explicit list) */
int mythos_b200_energy_f64(void* cuda_stream, const mb_energy_args* a);
int mythos_b200_energy_f32(void* cuda_stream, const mb_energy_args* a);

/* ---- cell-list neighbour build -------------------------------------------------------------------------------
 * Pair set (bit-exact contract): all i<j with i,j not listed in `bonded`, and
 *   d2 = dx*dx + dy*dy + dz*dz < (r_cutoff + dr_threshold)^2   (strict, evaluated in the positions' dtype,
 *   dx = xi - xj wrapped to [-L/2, L/2) when the box is periodic, no FMA contraction).
 * Output per frame: pairs (2,capacity), sorted by i then j, padded with N; count = pairs found (may exceed
 * capacity, in which case *overflow is set and the list is truncated -- jax_md's did_buffer_overflow). */
typedef struct mb_nl_args {
  int32_t n, n_frames;
  const void* center;     /* (F,N,3) */
  const int32_t* bonded;  /* (B,2) pairs to exclude */
  int32_t n_bonded;
  double box[3];
  double r_cutoff, dr_threshold;
  int32_t* pairs;         /* out (F,2,capacity) */
  int64_t capacity;
  int32_t* count;         /* out (F) */
  int32_t* overflow;      /* out (1), OR-ed */
  void* workspace;
  size_t workspace_bytes;
  uint32_t flags;         /* MB_NL_* */
  uint32_t _pad;
  int32_t* max_row;       /* rows mode: out (F) longest row found (to size the rows), or NULL */
  uint32_t tag_bits;      /* MB_NL_TAG_SUPPORTS: OR-ed into pairs[1][k] of every pair written (top 3 bits only)     */
  uint32_t _pad2;
  const int32_t* append_count; /* MB_NL_TAG_SUPPORTS: (F) entries already in each frame's list; this build appends after
                                * them and `count` receives the new totals (may alias append_count), or NULL      */
  int32_t lane_slots;     /* MB_NL_WARP_SLOTS: partners staged per nucleotide (<= 256)                              */
  int32_t _pad3;
  int64_t slot_base;      /* MB_NL_WARP_SLOTS: first entry of this build's slots in each frame's list               */
  int64_t slot_width;     /* MB_NL_WARP_SLOTS: entries per warp slot; slot_base + ceil(n/32)*slot_width <= capacity */
  /* Conditional rebuild on the device (the update() of jax_md's NeighborList, mythos/utils/neighbors.py:12-59 and
   * simulators/jax_md/jaxmd.py:82-94, without the host round trip): with `reference` the build of a frame happens only if
   * some nucleotide moved farther than move_threshold (= dr_threshold / 2) from its reference position; then
   * reference <- center and rebuilds[frame] += 1.  Otherwise the frame's list, count and max_row stay as they are.  One
   * launch, no host synchronisation, CUDA-graph capturable.  Free-space MB_NL_WARP_SLOTS builds on the frame-resident
   * route only (mythos_b200_nl_conditional_supported); anything else returns MB_EINVAL_SHAPE.                            */
  void* reference;        /* (F,N,3) in/out, dtype of center, or NULL (unconditional build)                          */
  double move_threshold;
  int32_t* rebuilds;      /* (F) or NULL                                                                              */
} mb_nl_args;
#define MB_NL_TAG_SUPPORTS 0x2u /* internal contract with mythos_b200_energy_* (MB_FLAG_TAGGED_PAIRS): `tag_bits` are OR-ed
                         * into pairs[1][k] (the index is pairs[1][k] & 0x1fffffff) and the build may append to a list.  The
                         * host builds the centres at the short-range cutoff (bit 30) and the backbone sites at the Debye
                         * cutoff (bit 29) into one list: only pairs inside the support of some term, labelled with it.  */
#define MB_NL_WARP_SLOTS 0x4u /* one-pass build: warp w (32 consecutive nucleotides in cell order) writes its pairs into entries
                         * [slot_base + w*slot_width, +slot_width) of the frame's list and pads the rest of the slot with n;
                         * `capacity` is the list stride; tag_bits are honoured; count accumulates over the builds sharing
                         * a list (zeroed when slot_base == 0); *overflow bit 0: a slot was too narrow, bit 2: a nucleotide
                         * had more than lane_slots partners; max_row (F,2) = longest nucleotide row, largest warp total.
                         * Consumers scan the whole capacity (pair_count = NULL).                                       */
#define MB_NL_ROWS 0x1u /* one-pass build: instead of a compact list, row k-major slots of width capacity / n per nucleotide
                         * (entry k * n + p = k-th partner of the p-th nucleotide in cell order), unused slots = n (the
                         * padding value); *overflow bit 0 is set if a row is too narrow; count = pairs found.  Consumers
                         * must scan the whole capacity (pair_count = NULL).                                            */
size_t mythos_b200_nl_workspace_bytes(int32_t n, int32_t n_frames);
#define MB_NL_PACKED_SLOTS 0x10u /* with MB_NL_WARP_SLOTS on the frame-resident route (mythos_b200_nl_conditional_supported): the
                         * warps' slots are packed back to back in warp order -- a compact list, count[f] valid entries at the
                         * head of the frame's list, nothing behind them written; slot_width is ignored, slot_base != 0 means
                         * "append behind the count[f] entries an earlier build left"; *overflow bit 0: capacity exceeded.
                         * Same pairs in the same order as the padded slots with the padding removed.                      */
#define MB_NL_REUSE_EXCLUSIONS 0x8u /* the exclusion table in `workspace` is that of an earlier build with the same bonded list
                         * (same workspace, same n): skip its two launches (per-step conditional rebuilds of an MD run)   */
/* 1 if a free-space MB_NL_WARP_SLOTS build of n nucleotides with these slots runs on the frame-resident route (one launch,
 * cell table and records in shared memory) on the current device, i.e. if mb_nl_args.reference may be used */
int mythos_b200_nl_conditional_supported(int32_t n, int32_t lane_slots, int32_t real_bytes /* 4 or 8 */);
int mythos_b200_nl_build_f64(void* cuda_stream, const mb_nl_args* a);
int mythos_b200_nl_build_f32(void* cuda_stream, const mb_nl_args* a);

/* backbone interaction sites (F*N,3) of (center, quat), the second point set of the support-tagged neighbour build;
 * nt_type (N) selects the flavour per nucleotide for the 3-bank model (NULL otherwise) */
int mythos_b200_backbone_sites_f64(void* cuda_stream, const mb_model* model, int64_t n_total, const void* center, const void* quat,
                                   void* out, const int32_t* nt_type, int32_t n);
int mythos_b200_backbone_sites_f32(void* cuda_stream, const mb_model* model, int64_t n_total, const void* center, const void* quat,
                                   void* out, const int32_t* nt_type, int32_t n);

/* the two point sets of the support-tagged build in one pass: centres and backbone sites of all frames as float32, brought
 * near the origin (model->box periodic: primary image; free space: relative to each frame's first nucleotide), and the
 * largest coordinate magnitude written (float, atomically maxed into *extent, which the caller zeroes): float32 builds
 * with cutoffs widened by 1e-3 are supersets of the float64 supports as long as *extent stays below ~1500 */
int mythos_b200_support_points_f64(void* cuda_stream, const mb_model* model, int32_t n, int32_t n_frames, const void* center,
                                   const void* quat, const int32_t* nt_type, void* out_center_f32, void* out_site_f32, void* extent_f32);
int mythos_b200_support_points_f32(void* cuda_stream, const mb_model* model, int32_t n, int32_t n_frames, const void* center,
                                   const void* quat, const int32_t* nt_type, void* out_center_f32, void* out_site_f32, void* extent_f32);

/* ---- rigid-body Langevin (BAOAB) step -------------------------------------------------------------------------
 * One call = B(dt/2) A(dt/2) O A(dt/2) on the state, i.e. everything of a step up to the force evaluation, or
 * the closing B(dt/2) kick; the caller evaluates the new force (mythos_b200_energy_*) in between. */
typedef struct mb_langevin_args {
  int32_t n;
  void* center;           /* (N,3) in/out */
  void* quat;             /* (N,4) in/out */
  void* p_center;         /* (N,3) linear momentum in/out */
  void* p_quat;           /* (N,4) quaternion-conjugate momentum in/out */
  const void* d_center;   /* (N,3) dE/dcenter at the current state */
  const void* d_quat;     /* (N,4) dE/dquat   at the current state */
  double dt, kT, gamma_center, gamma_quat, mass, inertia[3];
  double box[3];
  uint64_t seed, step;    /* counter-based RNG: noise = philox(seed, step, nucleotide) */
  const void* noise;      /* optional (N,6) standard normals (3 linear + 3 angular) overriding the RNG */
  int32_t phase;          /* 0 = B A O A (first part), 1 = closing B, 2 = closing B of the previous step fused with 0 */
  int32_t advance_step;   /* after the step, add 1 to *step_ptr (done inside the kernel by the last block to finish)   */
  const void* step_ptr;   /* optional device uint64[2]: [0] step counter (RNG counter and trajectory row; overrides
                           * `step`), [1] scratch that must be zero before the first call                            */
  void* traj_center;      /* optional out (traj_rows,N,3): positions after this call's drift at row *step_ptr      */
  void* traj_quat;        /* optional out (traj_rows,N,4)                                                          */
  int64_t traj_rows;
  int32_t zero_forces;    /* after the kick, zero d_center / d_quat (they are then written, not only read): the next
                           * mythos_b200_energy_* call can run with MB_FLAG_ACCUMULATE and no memset nodes           */
  int32_t _pad2;
} mb_langevin_args;
int mythos_b200_langevin_f64(void* cuda_stream, const mb_langevin_args* a);
int mythos_b200_langevin_f32(void* cuda_stream, const mb_langevin_args* a);

/* adjoint of one step: the vector-Jacobian product jax.grad computes through `step_fn` when a loss
 * is differentiated through the trajectory (mythos/simulators/jax_md/utils.py:174-193 checkpoint_scan; jaxmd.py:54-58,94).
 * The forces enter the step as an input: the kernel returns the cotangents of the PRE-step state (overwriting the
 * post-step cotangents it was given) and the cotangent of the force arrays; the force's own dependence on positions and
 * parameters is the caller's (two displaced mythos_b200_energy_* evaluations give the Hessian-vector product and the
 * mixed theta derivative).  `noise` (N,6) or the (seed, step) of the forward step; phase 0 (half kick) or 2 (full kick). */
typedef struct mb_langevin_adjoint_args {
  int32_t n;
  int32_t phase;
  const void* center;     /* pre-step state (N,3) */
  const void* quat;       /* (N,4) */
  const void* p_center;   /* (N,3) */
  const void* p_quat;     /* (N,4) */
  const void* d_center;   /* dE/dcenter at the pre-step positions (N,3) */
  const void* d_quat;     /* (N,4) */
  const void* noise;      /* (N,6) or NULL */
  void* lam_center;       /* in/out (N,3) */
  void* lam_quat;         /* in/out (N,4) */
  void* lam_p_center;     /* in/out (N,3) */
  void* lam_p_quat;       /* in/out (N,4) */
  void* lam_force_center; /* out (N,3) */
  void* lam_force_quat;   /* out (N,4) */
  double dt, kT, gamma_center, gamma_quat, mass;
  double inertia[3];
  uint64_t seed, step;
} mb_langevin_adjoint_args;
int mythos_b200_langevin_adjoint_f64(void* cuda_stream, const mb_langevin_adjoint_args* a);
int mythos_b200_langevin_adjoint_f32(void* cuda_stream, const mb_langevin_adjoint_args* a);

/* ---- DiffTRe reweighting ------------------------------------------------------------------------------------- */
typedef struct mb_weights_args {
  int32_t n_frames;
  const void* beta;        /* (F) 1/kT per frame */
  const void* e_new;       /* (F) */
  const void* e_ref;       /* (F) */
  void* weights;           /* out (F) softmax(-beta (e_new - e_ref)) */
  void* sums;              /* out (4): max exponent, sum exp, sum w ln w, n_eff */
} mb_weights_args;
int mythos_b200_weights_neff_f64(void* cuda_stream, const mb_weights_args* a);
int mythos_b200_weights_neff_f32(void* cuda_stream, const mb_weights_args* a);

/* ---- roofline denominators: FMA issue-rate micro-benchmark (blocks x 256 threads x iters x 8 FMA); time it with
 * CUDA events on `cuda_stream`; flops = blocks*256*iters*16.  scratch: blocks*256 reals. */
int mythos_b200_fma_peak_f64(void* cuda_stream, void* scratch, int blocks, int iters);
int mythos_b200_fma_peak_f32(void* cuda_stream, void* scratch, int blocks, int iters);
/* issue cost of the special functions the kernels call, same scheme: blocks x 256 threads x iters x 8 evaluations of
 * special(x)*a+b (one FMA per evaluation keeps the argument in range; subtract it).  kind: 0 div, 1 sqrt, 2 exp, 3 log,
 * 4 acos, 5 1/sqrt, 6 1/x, 7 cos, 8 fmod.  ops = blocks*256*iters*8.  These are the measured special-function weights of
 * the roofline work model. */
int mythos_b200_special_rate_f64(void* cuda_stream, void* scratch, int blocks, int iters, int kind);
int mythos_b200_special_rate_f32(void* cuda_stream, void* scratch, int blocks, int iters, int kind);

/* ---- trajectory ingest -----------------------------------------------------------------------
 * Stands in for `from_file` / `_read_file` (mythos/input/trajectory.py:192-320) and NucleotideState.quaternions
 * (trajectory.py:163-175 -> mythos/utils/math.py:9-65).  The caller copies the file's bytes to the device (16-byte
 * aligned), then:
 *   1. mythos_b200_traj_index(.., line_start = NULL, .., n_lines)   counts the lines (device int64 *n_lines)
 *   2. reads *n_lines, allocates line_start (n_lines + 1) and calls mythos_b200_traj_index again with it
 *   3. mythos_b200_traj_parse_*: F = n_lines / (N+3) states of "t = ..", "b = ..", "E = .." + N lines of 15 numbers; the
 *      first nine (centre, a1, a3) are converted with correct rounding (as strtod / np.fromstring do), the quaternion is
 *      the reference's axes -> Tait-Bryan ZYX -> quaternion formula; row r of a state goes to nucleotide dest[r] (the
 *      per-strand reversal of 5'->3' files, trajectory.py:291) or r.
 * status[0] = numbers that could not be converted (more than 19 significant digits, exponent outside +-64, malformed),
 * status[1] = header lines that are not where N nucleotide lines per state put them.  Both must be 0. */
typedef struct mb_traj_args {
  const void* text;          /* device copy of the file */
  int64_t n_bytes;
  const int64_t* line_start; /* (n_lines + 1) from mythos_b200_traj_index */
  int64_t n_lines;
  int32_t n, n_frames;
  const int32_t* dest;       /* (N) or NULL */
  const uint64_t* pow5;      /* (129,2) 128-bit truncated powers of five 5^-64..5^64, device (mythos_b200.input.trajectory) */
  void* center;              /* out (F,N,3) */
  void* quat;                /* out (F,N,4) */
  double* times;             /* out (F) */
  double* box;               /* out (F,3) */
  double* energies;          /* out (F,3) */
  int32_t* status;           /* out (4) */
} mb_traj_args;
size_t mythos_b200_traj_workspace_bytes(int64_t n_bytes);
int mythos_b200_traj_index(void* cuda_stream, const void* text, int64_t n_bytes, void* workspace, size_t workspace_bytes,
                           int64_t* line_start, int64_t line_capacity, int64_t* n_lines);
int mythos_b200_traj_parse_f64(void* cuda_stream, const mb_traj_args* a);
int mythos_b200_traj_parse_f32(void* cuda_stream, const mb_traj_args* a);

/* ---- theta -> parameter-bank chain (host side; no device work) ------------------------------------------------
 * Stands in for `BaseConfiguration.init_params` of every term (mythos/energy/configuration.py:110-113, e.g.
 * mythos/energy/dna1/stacking.py:120-183, hydrogen_bonding.py:148-223) and the smoothing solvers
 * (mythos/energy/dna1/base_smoothing_functions.py:48-142) as they run on every parameter update of a DiffTRe step
 * (optimization/objective.py:224 `energy_fn.with_params(opt_params)`), together with their reverse-mode derivative.
 * The chain is a straight-line scalar program (no data-dependent branches): node k computes op[k](node arg0[k],
 * node arg1[k]) with operands that precede it; `imm` holds constants (MB_TAPE_CONST) and exponents (MB_TAPE_POW);
 * MB_TAPE_INPUT reads inputs[arg0[k]].  out[o] = node whose value is parameter-bank slot o (-1: slot stays 0).
 * `values` / `adjoint` are caller-owned scratch of n_nodes doubles; forward fills `values`, the VJP reads them. */
enum mb_tape_op {
  MB_TAPE_CONST = 0, MB_TAPE_INPUT = 1, MB_TAPE_ADD = 2, MB_TAPE_SUB = 3, MB_TAPE_MUL = 4, MB_TAPE_DIV = 5,
  MB_TAPE_NEG = 6, MB_TAPE_RECIP = 7, MB_TAPE_EXP = 8, MB_TAPE_LOG = 9, MB_TAPE_SQRT = 10, MB_TAPE_POW = 11
};
typedef struct mb_theta_tape {
  int32_t n_nodes, n_inputs, n_outputs, _pad;
  const int32_t* op;    /* (n_nodes) mb_tape_op */
  const int32_t* arg0;  /* (n_nodes) */
  const int32_t* arg1;  /* (n_nodes) */
  const double* imm;    /* (n_nodes) */
  const int32_t* out;   /* (n_outputs) */
} mb_theta_tape;
int mythos_b200_theta_tape_forward(const mb_theta_tape* tape, const double* inputs, double* values, double* outputs);
int mythos_b200_theta_tape_vjp(const mb_theta_tape* tape, const double* values, const double* out_cot, double* adjoint,
                               double* in_grad);

/* ---- introspection ------------------------------------------------------------------------------------------- */
int mythos_b200_abi_version(void);
int mythos_b200_param_count(void);                 /* MB_P_COUNT                              */
const char* mythos_b200_param_name(int index);     /* "<term>.<name>" or NULL                 */
int mythos_b200_param_index(const char* name);     /* -1 if unknown                           */
const char* mythos_b200_last_error(void);          /* thread-local message of the last failure */
size_t mythos_b200_sizeof_model(void);
size_t mythos_b200_sizeof_energy_args(void);
size_t mythos_b200_sizeof_nl_args(void);

#ifdef __cplusplus
}
#endif
#endif /* MYTHOS_B200_H */
