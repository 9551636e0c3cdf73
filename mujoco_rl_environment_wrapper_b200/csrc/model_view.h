// model_view.h — typed read-only view over a packed model blob (include/mjb_blob.h).
// Shared by the MJCF compiler's constant pass, the CUDA batch (fp32 device image
// builder) and the fp64 CPU oracle.  All pointers alias the blob.
#pragma once
#include <stdexcept>
#include <string>

#include "../../include/mjb_blob.h"

namespace mjb {

struct ModelView {
  const void* blob = nullptr;
  // dims
  int nq = 0, nv = 0, nu = 0, nbody = 0, njnt = 0, ngeom = 0, nsite = 0, nsensor = 0, nsensordata = 0, npair = 0;
  int integrator = 0, ntree = 0, maxdepth = 0, solver_iterations = 100;
  double timestep = 0.002, impratio = 1, tolerance = 1e-8;
  const double* gravity = nullptr;
  // body
  const int32_t *body_parentid, *body_rootid, *body_weldid, *body_jntnum, *body_jntadr, *body_dofnum, *body_dofadr,
      *body_geomnum, *body_geomadr, *body_depth, *body_treeid;
  const double *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass, *body_inertia, *body_subtreemass,
      *body_invweight0;
  // joint
  const int32_t *jnt_type, *jnt_bodyid, *jnt_qposadr, *jnt_dofadr, *jnt_limited;
  const double *jnt_pos, *jnt_axis, *jnt_range, *jnt_margin, *jnt_solref, *jnt_solimp;
  // dof
  const int32_t *dof_bodyid, *dof_jntid, *dof_parentid;
  const double *dof_armature, *dof_damping, *dof_invweight0;
  // geom
  const int32_t *geom_type, *geom_bodyid, *geom_contype, *geom_conaffinity, *geom_condim;
  const double *geom_size, *geom_pos, *geom_quat, *geom_friction, *geom_margin, *geom_gap, *geom_solmix, *geom_solref,
      *geom_solimp, *geom_rbound, *geom_rgba;
  // site
  const int32_t *site_bodyid, *site_type;
  const double *site_pos, *site_quat, *site_size;
  // sensor
  const int32_t *sensor_type, *sensor_objtype, *sensor_objid, *sensor_adr, *sensor_dim, *sensor_datatype;
  const double* sensor_cutoff;
  // actuator
  const int32_t *actuator_trnid, *actuator_ctrllimited;
  const double *actuator_gear, *actuator_ctrlrange;
  // collision pair table (static filters applied; geom1 has the lower type id)
  const int32_t *pair_geom1, *pair_geom2, *pair_condim;
  const double *pair_margin, *pair_includemargin, *pair_friction, *pair_solref, *pair_solimp;
  const double* qpos0;
  // cameras (optional in older blobs)
  int ncam = 0;
  const int32_t *cam_bodyid = nullptr, *cam_mode = nullptr;
  const double *cam_pos = nullptr, *cam_quat = nullptr, *cam_fovy = nullptr;

  const int32_t* I(const char* n, bool required = true) const {
    int c;
    const int32_t* p = mjb_blob_i32(blob, n, &c);
    if (c < 0 && required) throw std::runtime_error(std::string("model blob: missing int field ") + n);
    return p;
  }
  const double* F(const char* n, bool required = true) const {
    int c;
    const double* p = mjb_blob_f64(blob, n, &c);
    if (c < 0 && required) throw std::runtime_error(std::string("model blob: missing f64 field ") + n);
    return p;
  }
  int scalar(const char* n) const { return I(n)[0]; }

  explicit ModelView(const void* b, bool need_invweight = true) : blob(b) {
    const mjb_blob_header* h = (const mjb_blob_header*)b;
    if (!b || memcmp(h->magic, MJB_BLOB_MAGIC, 8) != 0) throw std::runtime_error("model blob: bad magic");
    nq = scalar("nq"); nv = scalar("nv"); nu = scalar("nu"); nbody = scalar("nbody"); njnt = scalar("njnt");
    ngeom = scalar("ngeom"); nsite = scalar("nsite"); nsensor = scalar("nsensor");
    nsensordata = scalar("nsensordata"); npair = scalar("npair"); integrator = scalar("opt_integrator");
    ntree = scalar("ntree"); maxdepth = scalar("maxdepth"); solver_iterations = scalar("opt_iterations");
    timestep = F("opt_timestep")[0]; impratio = F("opt_impratio")[0]; tolerance = F("opt_tolerance")[0];
    gravity = F("opt_gravity");
    body_parentid = I("body_parentid"); body_rootid = I("body_rootid"); body_weldid = I("body_weldid");
    body_jntnum = I("body_jntnum"); body_jntadr = I("body_jntadr"); body_dofnum = I("body_dofnum");
    body_dofadr = I("body_dofadr"); body_geomnum = I("body_geomnum"); body_geomadr = I("body_geomadr");
    body_depth = I("body_depth"); body_treeid = I("body_treeid");
    body_pos = F("body_pos"); body_quat = F("body_quat"); body_ipos = F("body_ipos"); body_iquat = F("body_iquat");
    body_mass = F("body_mass"); body_inertia = F("body_inertia"); body_subtreemass = F("body_subtreemass");
    body_invweight0 = F("body_invweight0", need_invweight);
    jnt_type = I("jnt_type"); jnt_bodyid = I("jnt_bodyid"); jnt_qposadr = I("jnt_qposadr");
    jnt_dofadr = I("jnt_dofadr"); jnt_limited = I("jnt_limited");
    jnt_pos = F("jnt_pos"); jnt_axis = F("jnt_axis"); jnt_range = F("jnt_range"); jnt_margin = F("jnt_margin");
    jnt_solref = F("jnt_solref"); jnt_solimp = F("jnt_solimp");
    dof_bodyid = I("dof_bodyid"); dof_jntid = I("dof_jntid"); dof_parentid = I("dof_parentid");
    dof_armature = F("dof_armature"); dof_damping = F("dof_damping");
    dof_invweight0 = F("dof_invweight0", need_invweight);
    geom_type = I("geom_type"); geom_bodyid = I("geom_bodyid"); geom_contype = I("geom_contype");
    geom_conaffinity = I("geom_conaffinity"); geom_condim = I("geom_condim");
    geom_size = F("geom_size"); geom_pos = F("geom_pos"); geom_quat = F("geom_quat");
    geom_friction = F("geom_friction"); geom_margin = F("geom_margin"); geom_gap = F("geom_gap");
    geom_solmix = F("geom_solmix"); geom_solref = F("geom_solref"); geom_solimp = F("geom_solimp");
    geom_rbound = F("geom_rbound"); geom_rgba = F("geom_rgba");
    site_bodyid = I("site_bodyid"); site_type = I("site_type");
    site_pos = F("site_pos"); site_quat = F("site_quat"); site_size = F("site_size");
    sensor_type = I("sensor_type"); sensor_objtype = I("sensor_objtype"); sensor_objid = I("sensor_objid");
    sensor_adr = I("sensor_adr"); sensor_dim = I("sensor_dim"); sensor_datatype = I("sensor_datatype");
    sensor_cutoff = F("sensor_cutoff");
    actuator_trnid = I("actuator_trnid"); actuator_ctrllimited = I("actuator_ctrllimited");
    actuator_gear = F("actuator_gear"); actuator_ctrlrange = F("actuator_ctrlrange");
    pair_geom1 = I("pair_geom1"); pair_geom2 = I("pair_geom2"); pair_condim = I("pair_condim");
    pair_margin = F("pair_margin"); pair_includemargin = F("pair_includemargin");
    pair_friction = F("pair_friction"); pair_solref = F("pair_solref"); pair_solimp = F("pair_solimp");
    qpos0 = F("qpos0");
    if (const int32_t* nc = I("ncam", false)) {
      ncam = nc[0];
      cam_bodyid = I("cam_bodyid"); cam_mode = I("cam_mode");
      cam_pos = F("cam_pos"); cam_quat = F("cam_quat"); cam_fovy = F("cam_fovy");
    }
  }
};

}  // namespace mjb
