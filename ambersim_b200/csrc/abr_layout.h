// abr_layout.h — device-side model blob + per-world shared-memory layout.
//
// The flattened model (AbrModelHost, include/abr.h) is repacked at abr_model_create into one
// float pool and one int pool ("blob") that every CTA stages into shared memory once; the
// Layout struct (all offsets/dims/options) travels by value as a __grid_constant__ kernel
// parameter, so every offset is a constant-bank operand.
#ifndef ABR_LAYOUT_H_
#define ABR_LAYOUT_H_

namespace abr {

// per-row constant block produced on the host from (solref, solimp, timestep, flags):
//   [0] k  [1] b  [2] dmin  [3] dmax  [4] 1/width  [5] mid  [6] power  [7] invweight
//   [8] 1/mid^(power-1)  [9] 1/(1-mid)^(power-1)
// contact pairs append: [10] invweight of tangent 2  [11] mu1  [12] mu2  [13] includemargin
constexpr int kRowPrm = 10;
constexpr int kConPrm = 14;
constexpr int kActPrm = 11;  // ctrlrange[2] forcerange[2] gainprm[3] biasprm[3] gear

struct Layout {
  // ---- dims
  int nq, nv, nu, nbody, njnt, ngeom, npair, nx;
  int ne, nl, ncon, nefc;      // static constraint sizes (MJX era: inactive rows are zeroed)
  int depth;                   // deepest body level (world = 0)
  int nroot;                   // kinematic roots (bodies whose parent is the world)
  int nmpair;                  // nonzero lower-triangular entries of the joint-space inertia
  int ntri;                    // nv*(nv+1)/2
  // ---- options
  int integrator, solver, iterations, ls_iterations, disableflags;
  float timestep, tolerance, ls_tolerance, meaninertia, impratio;
  float gravity[3];
  // ---- blob sizes
  int n_mf, n_mi;
  // ---- float pool offsets
  int f_body_pos, f_body_quat, f_body_ipos, f_body_iquat, f_body_mass, f_body_inertia;
  int f_jnt_pos, f_jnt_axis, f_jnt_range, f_jnt_margin, f_jnt_stiffness;
  int f_dof_armature, f_dof_damping;
  int f_qpos0, f_qpos_spring;
  int f_geom_size, f_geom_pos, f_geom_quat, f_vert, f_face_normal;
  int f_eq_prm, f_eq_data, f_lim_prm, f_con_prm, f_act_prm;
  // ---- int pool offsets
  int i_body_parent, i_body_jntadr, i_body_jntnum, i_body_dofadr, i_body_dofnum, i_body_rootslot;
  int i_body_childadr, i_body_childnum, i_child;
  int i_level_adr, i_level_body;
  int i_jnt_type, i_jnt_qposadr, i_jnt_dofadr, i_jnt_body;
  int i_dof_body, i_dof_jnt, i_dof_limrow, i_dof_actadr, i_dof_actnum, i_dof_act;
  int i_mpair;      // (i << 16) | j  for every nonzero lower entry of M
  int i_tri;        // (i << 16) | j  for every packed lower entry k
  int i_root_body;
  int i_lim_jnt, i_eq_j1, i_eq_j2;
  int i_con_pair, i_con_sub, i_con_row, i_con_condim, i_con_dofmask;
  int i_row_info;   // per efc row: kind (0 eq, 1 limit, 2 contact) | idx << 2 | sub << 20
  int i_pair_g1, i_pair_g2, i_pair_kind, i_geom_body, i_geom_vertadr, i_geom_vertnum;
  int i_geom_faceadr, i_geom_facenum, i_face_vertadr, i_face_vertnum, i_face_vert, i_geom_edgeadr, i_geom_edgenum, i_edge_vert;
  int i_act_jnt, i_act_flags;
  // ---- tree-sparse linear algebra (valid when every constraint couples dofs of one ancestor
  //      chain only, so H = M + J'DJ keeps M's branch-induced sparsity and L'DL has no fill-in)
  int sparse;       // 1 = tree-sparse path, 0 = dense packed-Cholesky fallback
  int nnz;          // entries (i,j), j ancestor-or-self of i, in row-chain order (== nmpair)
  int nstage;       // elimination stages: dofs of equal height in the dof tree
  int i_sp_adr;     // [nv+1] row start in the sparse array (diagonal first, then ancestors)
  int i_st_adr, i_st_dof;                         // stage -> dofs eliminated
  int i_fu_tadr, i_fu_tgt, i_fu_cadr, i_fu_src, i_fu_k;   // factor updates, gathered by target entry
  int i_sb_tadr, i_sb_tgt, i_sb_cadr, i_sb_src;   // back-substitution, gathered by target dof
  int i_mm_adr, i_mm_src;                         // M v: per row (entry << 16 | other dof)
  int i_hc_adr, i_hc_con;                         // Hessian: contacts touching both dofs of an entry
  int i_he_adr, i_he_eq;                          // Hessian: equality rows touching an entry (eq << 2 | which)
  int i_cd_adr, i_cd_dof;                         // contact -> dofs with a non-zero Jacobian column
  // ---- per-world shared-memory offsets (floats)
  int w_qpos, w_qvel, w_warm, w_ctrl, w_M, w_eqc, w_lims, w_B, w_D, w_aref, w_fs, w_as, w_H;
  int w_xpos, w_xquat, w_xipos, w_xanchor, w_xaxis, w_rootcom, w_cinert, w_cdof;
  int w_crb, w_buf, w_cvel, w_cacc, w_cdofdot, w_cdist, w_cpos, w_cframe, w_actf, w_bv;
  int w_a, w_Ma, w_grad, w_search, w_mv, w_fc, w_Jaref, w_jv, w_force, w_Fc, w_WB, w_y, w_invD;
  // ---- limb path (abr_limb.cuh): one lane per root-to-leaf path of a floating-base tree
  int limb_ok;      // model/options eligible for the limb kernels
  int lNL, lNC;     // compiled instantiation that serves this model (chain length, contacts per path)
  int lg2G;         // log2 of the lanes per world (paths padded to a power of two, <= 8)
  int l_mx;         // 2 bits per chain position: log2 of the largest lane group sharing a body there
  int f_ltab;       // float-pool offset of the per-lane table (limb::Map, row stride limb::kStride)
  float l_mass;     // total mass of the tree (subtree CoM denominator)
  int l_cb;         // every lane's contacts sit on the lane's leaf body and touch one plane (contact-body form)
  int l_hinge;      // every chain joint is a hinge (assumed by the fast variants)
  int l_pow2;       // every limit / contact row of the limb table has the default impedance power 2 (assumed by the fast variants)
  int hand_ok;      // fixed-base chains + joint equalities, no contacts: served by the hand kernels (abr_hand.cuh); uses lNL, lg2G
  int f_htab;       // float-pool offset of the per-lane table (hand::Map, row stride limb::kStride)
  int w_rk;         // RK4 save area: qpos0[nq] qvel0[nv] warm0[nv] sv[nv] sa[nv] kq[nv]
  int world_stride; // floats per world
};

}  // namespace abr
#endif
