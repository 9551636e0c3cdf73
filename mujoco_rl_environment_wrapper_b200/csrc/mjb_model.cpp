// mjb_model.cpp — host-only half of the C-ABI (include/mjb.h): MJCF -> model handle,
// dims, blob access and name lookups.  No CUDA in this translation unit.
#include <cstring>
#include <string>

#include "../../include/mjb.h"
#include "host_model.h"
#include "mjb_internal.h"

namespace mjb {
thread_local std::string g_last_error;
void set_error(const std::string& s) { g_last_error = s; }
}  // namespace mjb

extern "C" {

const char* mjb_last_error(void) { return mjb::g_last_error.c_str(); }
const char* mjb_version(void) { return "mjb 0.1 (sm_100a)"; }

int mjb_model_create(const char* mjcf_text, mjb_model** out) {
  if (!mjcf_text || !out) { mjb::set_error("mjb_model_create: null argument"); return MJB_ERR_ARG; }
  *out = nullptr;
  try {
    mjb_model* m = new mjb_model();
    try {
      mjb::compile_mjcf(mjcf_text, m->host);
    } catch (...) {
      delete m;
      throw;
    }
    *out = m;
    return MJB_OK;
  } catch (const std::exception& e) {
    mjb::set_error(e.what());
    return MJB_ERR_PARSE;
  }
}

void mjb_model_destroy(mjb_model* m) { delete m; }

int mjb_model_dims(const mjb_model* m, mjb_dims* out) {
  if (!m || !out) { mjb::set_error("mjb_model_dims: null argument"); return MJB_ERR_ARG; }
  try {
    const mjb::HostModel& h = m->host;
    memset(out, 0, sizeof(*out));
    out->nq = h.get_int("nq"); out->nv = h.get_int("nv"); out->nu = h.get_int("nu");
    out->nbody = h.get_int("nbody"); out->njnt = h.get_int("njnt"); out->ngeom = h.get_int("ngeom");
    out->nsite = h.get_int("nsite"); out->nsensor = h.get_int("nsensor");
    out->nsensordata = h.get_int("nsensordata"); out->npair = h.get_int("npair");
    out->integrator = h.get_int("opt_integrator"); out->ncam = h.get_int("ncam");
    out->timestep = h.Fv("opt_timestep")[0];
    return MJB_OK;
  } catch (const std::exception& e) {
    mjb::set_error(e.what());
    return MJB_ERR_ARG;
  }
}

const void* mjb_model_blob(const mjb_model* m, int64_t* nbytes) {
  if (!m) return nullptr;
  if (nbytes) *nbytes = (int64_t)m->host.blob.size();
  return m->host.blob.data();
}

int mjb_name2id(const mjb_model* m, int objtype, const char* name) {
  if (!m || !name || !*name) return -1;
  auto it = m->host.names.find(objtype);
  if (it == m->host.names.end()) return -1;
  for (size_t i = 0; i < it->second.size(); i++)
    if (it->second[i] == name) return (int)i;
  return -1;
}

const char* mjb_id2name(const mjb_model* m, int objtype, int id) {
  if (!m) return nullptr;
  auto it = m->host.names.find(objtype);
  if (it == m->host.names.end() || id < 0 || id >= (int)it->second.size()) return nullptr;
  return it->second[id].c_str();
}

}  // extern "C"
