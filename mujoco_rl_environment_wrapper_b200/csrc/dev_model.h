// dev_model.h — fp32 device image of a compiled model + the per-env shared-memory plan.
//
// Host side (build_dev_model): turns the fp64 blob (ModelView) + the env spec into ONE flat
// 4-byte-word image that the step kernel stages into shared memory with a single TMA bulk copy
// (cp.async.bulk, SASS UBLKCP), plus a small POD header (DevModel) passed as a kernel parameter.
// Device side: typed accessors over the image.
#pragma once
#include <stdint.h>

#ifndef MJB_HD
#if defined(__CUDACC__)
#define MJB_HD __host__ __device__ __forceinline__
#else
#define MJB_HD inline
#endif
#endif

namespace mjb {

// ---- constant image fields --------------------------------------------------------------------
// Moving bodies are renumbered ("kernel order") by tree depth so that one FK level is a contiguous
// lane range.  Static bodies do not exist on the device: their geoms carry precomputed world frames.
#define MJB_IMAGE_FIELDS(X)                                                                         \
  X(level_adr)      /* int [nlevel+1]  kernel-body range of every depth level                    */ \
  X(mb_parent)      /* int [nmb]       parent kernel body, -1 = static parent (pose folded in)   */ \
  X(mb_root)        /* int [nmb]       kernel body of the tree root                              */ \
  X(mb_jntadr)      /* int [nmb]                                                                 */ \
  X(mb_jntnum)      /* int [nmb]                                                                 */ \
  X(mb_dofadr)      /* int [nmb]       first dof, -1 if none                                     */ \
  X(mb_dofnum)      /* int [nmb]                                                                 */ \
  X(mb_childadr)    /* int [nmb+1]                                                               */ \
  X(mb_child)       /* int [nmb]       children lists                                            */ \
  X(mb_dofmask)     /* u32 [nmb*2]     dofs on the chain world->body (bit d), nv <= 64           */ \
  X(mb_pos)         /* f32 [nmb*3]     frame in parent (world frame when mb_parent == -1)        */ \
  X(mb_quat)        /* f32 [nmb*4]                                                               */ \
  X(mb_ipos)        /* f32 [nmb*3]                                                               */ \
  X(mb_iquat)       /* f32 [nmb*4]                                                               */ \
  X(mb_mass)        /* f32 [nmb]                                                                 */ \
  X(mb_inertia)     /* f32 [nmb*3]                                                               */ \
  X(mb_invweight)   /* f32 [nmb]       body_invweight0 (translational)                           */ \
  X(jnt_type)       /* int [njnt]                                                                */ \
  X(jnt_qposadr)    /* int [njnt]                                                                */ \
  X(jnt_dofadr)     /* int [njnt]                                                                */ \
  X(jnt_pos)        /* f32 [njnt*3]                                                              */ \
  X(jnt_axis)       /* f32 [njnt*3]                                                              */ \
  X(jnt_qpos0)      /* f32 [njnt]      reference angle (hinge / slide)                           */ \
  X(lim_dof)        /* int [nlim]      limited joints: dof, qpos address                         */ \
  X(lim_qposadr)    /* int [nlim]                                                                */ \
  X(lim_param)      /* f32 [nlim*12]   range lo, hi, margin, K, B, invweight, solimp[5], pad     */ \
  X(dof_mb)         /* int [nv]                                                                  */ \
  X(dof_parent)     /* int [nv]                                                                  */ \
  X(dof_kind)       /* int [nv]        0 hinge/slide axis in body, 1 free translation, 2 free rotation */ \
  X(dof_t0)         /* int [32]        first dof of the lane's kinematic tree (lane = dof)        */ \
  X(dof_t1)         /* int [32]        one past the last dof of that tree (0 for lanes >= nv)     */ \
  X(dof_lim)        /* int [32]        limited-joint slot of the dof (lane = dof), -1 = none (also lanes >= nv) */ \
  X(dof_armature)   /* f32 [nv]                                                                  */ \
  X(dof_damping)    /* f32 [nv]                                                                  */ \
  X(geom_type)      /* int [ngeom]                                                               */ \
  X(geom_mb)        /* int [ngeom]     kernel body, -1 static                                    */ \
  X(geom_slot)      /* int [ngeom]     slot in the per-env dynamic geom frames, -1 static        */ \
  X(geom_size)      /* f32 [ngeom*3]                                                             */ \
  X(geom_rbound)    /* f32 [ngeom]                                                               */ \
  X(geom_pos)       /* f32 [ngeom*3]   local (dynamic) or world (static)                         */ \
  X(geom_mat)       /* f32 [ngeom*9]   static: world rotation; dynamic: unused                   */ \
  X(geom_quat)      /* f32 [ngeom*4]   dynamic: local quaternion                                 */ \
  X(geom_ray)       /* int [ngeom]     1 if visible to rays (alpha != 0)                         */ \
  X(pair_pack)      /* u32 [npair]     g1 | g2 << 12 | class << 24  (sorted by type pair)        */ \
  X(pclass)         /* f32 [nclass*12] margin, includemargin, mu, K, B, solimp[5], condim, pad   */ \
  X(bp_block)       /* [nblock*4]      pair blocks (tree x tree | tree x static geom): see BP_*  */ \
  X(bp_passmask)    /* u32 [npass*2]   blocks present in each 32-pair pass of the broad phase    */ \
  X(site_mb)        /* int [nsite]                                                               */ \
  X(site_type)      /* int [nsite]                                                               */ \
  X(site_pos)       /* f32 [nsite*3]   local (or world when static)                              */ \
  X(site_quat)      /* f32 [nsite*4]                                                             */ \
  X(site_size)      /* f32 [nsite*3]                                                             */ \
  X(sensor_type)    /* int [nsensor]                                                             */ \
  X(sensor_site)    /* int [nsensor]                                                             */ \
  X(sensor_adr)     /* int [nsensor]                                                             */ \
  X(sensor_dim)     /* int [nsensor]                                                             */ \
  X(sensor_dtype)   /* int [nsensor]   0 real, 1 positive, 2 axis                                */ \
  X(sensor_cutoff)  /* f32 [nsensor]                                                             */ \
  X(act_dof)        /* int [nu]                                                                  */ \
  X(act_param)      /* f32 [nu*4]      gear, ctrllimited, lo, hi                                 */ \
  X(dof_actadr)     /* int [nv]        actuators acting on a dof: range in act_list              */ \
  X(dof_actnum)     /* int [nv]                                                                  */ \
  X(act_list)       /* int [nu]        actuator ids grouped by dof                               */ \
  X(probe_kind)     /* int [nprobe]    0 body xipos (kernel body), 1 geom xpos, 2 constant       */ \
  X(probe_id)       /* int [nprobe]                                                              */ \
  X(probe_const)    /* f32 [nprobe*3]  value for constant (static) probes                        */ \
  X(act_index)      /* int [n_agents*n_phys_act]                                                 */ \
  X(obs_index)      /* int [sum obs]   (kind << 24) | address                                    */ \
  X(tri_lut)        /* u32 [136]       pair p = r (r + 1) / 2 + c  ->  r | c << 8   (r >= c, r < 16)       */ \
  X(qpos0)          /* f32 [nq]                                                                  */

enum ImageField {
#define X(n) IF_##n,
  MJB_IMAGE_FIELDS(X)
#undef X
      IF_COUNT
};

// ---- per-env shared-memory scratch fields ---------------------------------------------------
#define MJB_SMEM_FIELDS(X)                                                                          \
  X(qpos) X(qvel) X(qacc) X(ctrl) X(qfrc) /* state, current accel iterate, smooth force         */ \
  X(xpos) X(xquat) X(xmat) X(xipos)       /* body frames                                         */ \
  X(cdof) X(cinert) X(crb) X(cvel) X(cacc) /* spatial quantities about the tree root origin      */ \
  X(gpos) X(gmat) X(spos) X(smat)         /* dynamic geom / site world frames                    */ \
  X(M) X(H)                               /* joint-space inertia, Newton Hessian / factors       */ \
  X(cand)                                 /* broadphase survivors (pair indices)                 */ \
  X(con)                                  /* contacts: dist, pos3, frame9, pair, mu, pad -> 16   */ \
  X(J)                                    /* contact Jacobians, 3 rows per contact (n, t1, t2)   */ \
  X(efcD) X(efcAref) X(efcJar) X(efcJv)   /* per-row: limits 2*nlim then 4 per contact           */ \
  X(vecA) X(vecB) X(vecC) X(vecD)         /* nv-sized temporaries (force, grad, search, spare)    */ \
  X(rk)                                   /* RK4 stage storage: q0, v0, dq, dv                    */ \
  X(sens)                                 /* sensordata                                          */

enum SmemField {
#define X(n) SF_##n,
  MJB_SMEM_FIELDS(X)
#undef X
      SF_COUNT
};

struct DevPlugin {
  int kind, act_lo, act_hi, n_obs;
  float param[4];
};

struct DevModel {
  // sizes
  int nq, nv, nu, nmb, njnt, nlim, ngeom, ngdyn, nsite, nsensor, nsensordata, npair, nclass, nlevel, nprobe;
  int maxcon, maxcand, maxefc, ldj, maxtree, nblock;
  // packing: `pack` real envs per warp (see replicate.h); the *1 sizes are those of ONE real env
  int pack, a1, t1, nq1, nv1, nu1, ns1, np1, ngeom1, maxcon1;
  int integrator, has_damping, need_acc_sensors;
  float timestep, gravity[3];
  double timestep_d;   // opt.timestep as compiled (fp64): the ant reward divides by it like the reference
  int solver_iterations, ls_iterations;
  float solver_tol;
  float reset_noise;   // 0: resets start exactly at qpos0 (reference behaviour)
  int njnt1;           // joints of one real env
  // env spec
  int n_agents, free_joint, skip_frames, max_steps, n_phys_act, act_dim;
  int obs_dim[8], obs_adr[9], agent_probe[8];
  int n_dynamics, n_rewards, n_dones, n_targets;
  DevPlugin dynamics[4], rewards[4], dones[4];
  int target_probe[16];
  unsigned long long seed;
  // layout
  int off[IF_COUNT];      // word offsets into the constant image
  int image_words;        // multiple of 4 (16 B) for the bulk copy
  int soff[SF_COUNT];     // word offsets into one env's scratch
  int env_words;          // scratch words per env (multiple of 4)
  // HBM strides
  int qpos_stride, qvel_stride, ctrl_stride, sensor_stride, act_stride, obs_stride, store_i32, store_f32;
};

// ---- camera rendering (render_kernel.cuh): a second, small constant table next to the image -------------
// cam record: kernel body (-1 = static: pose already in the world frame), mode (0 fixed), pos[3], quat[4],
// tan(fovy / 2), 2 pad words; then one clamped RGBA (4 floats) per geom
enum { CAM_MB = 0, CAM_MODE = 1, CAM_POS = 2, CAM_QUAT = 5, CAM_TANHALF = 9, CAM_STRIDE = 12 };
struct RenderHdr {
  int ncam, ngeom;
  int off_cam, off_rgba;   // word offsets into the table
  int words;               // multiple of 4
};
// per-env geom record the rays walk (built in shared memory after the kinematics pass)
enum { GL_POS = 0, GL_MAT = 3, GL_SIZE = 12, GL_TYPE = 15, GL_RBOUND = 16, GL_RGB = 17, GL_STRIDE = 20 };

// contact record layout inside SF_con (16 words per contact)
// (CON_DOFS: the dofs of the contact's mask as up to 16 packed bytes, ascending)
enum { CON_DIST = 0, CON_POS = 1, CON_FRAME = 4, CON_PAIR = 13, CON_MU = 14, CON_MASK = 15, CON_DOFS = 16, CON_STRIDE = 20 };
enum { LIM_LO = 0, LIM_HI, LIM_MARGIN, LIM_K, LIM_B, LIM_INVW, LIM_SOLIMP, LIM_STRIDE = 12 };
enum { PC_MARGIN = 0, PC_INCMARGIN, PC_MU, PC_K, PC_B, PC_SOLIMP, PC_CONDIM = 10, PC_STRIDE = 12 };
enum { DOF_AXIS = 0, DOF_FREE_TRANS = 1, DOF_FREE_ROT = 2 };
enum { PROBE_BODY = 0, PROBE_GEOM = 1, PROBE_CONST = 2 };
// pair block: root kernel body of tree A, root kernel body of tree B (or -1 - static geom id, or BP_ALWAYS),
// reach of A + reach of B (trees) / reach of A (static partner), largest pair margin in the block
enum { BP_ROOT_A = 0, BP_PARTNER = 1, BP_REACH = 2, BP_MARGIN = 3, BP_STRIDE = 4, BP_ALWAYS = 0x7fffffff };

}  // namespace mjb
