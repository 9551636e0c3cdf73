// env_kernel.cuh — per-environment driver around the physics in step_kernel.cuh: state load /
// action scatter / substeps / state store, and the fused epilogue (observation gather, dynamics,
// reward, truncation, done — MuJoCo_Gym/mujoco_rl.py:243-289 and the reset path :291-331).
#pragma once
#include "step_kernel.cuh"

namespace mjb {

enum { MODE_STEP = 0, MODE_PHYSICS = 1, MODE_FORWARD = 2, MODE_RESET = 3 };

// counter-based draw replacing `random.randint` (README.md:154, Testing/Pick_Up_Dynamic.py:28,38)
MJB_HD uint32_t draw_u32(unsigned long long seed, uint32_t env, uint32_t agent, uint32_t counter) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (1ull + env) + 0xBF58476D1CE4E5B9ull * agent +
                         0x94D049BB133111EBull * counter;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}

MJB_DEV float probe_dist(const float* probe, int a, int b) {
  float dx = probe[4 * a] - probe[4 * b], dy = probe[4 * a + 1] - probe[4 * b + 1], dz = probe[4 * a + 2] - probe[4 * b + 2];
  return sqrtf(dx * dx + dy * dy + dz * dz);
}

// the dynamics / reward / done programme of one env, executed by one lane in the reference's order
// (dynamic outer, agent inner; then reward fn outer, agent inner; truncation; done fns with early exit)
MJB_DEV void run_plugins(const Ctx& c, const mjb_buffers& B, int env, bool is_reset, const float* probe, int* si, float* sf,
                         const float* act, int* ts_io) {
  const DevModel& dm = *c.dm;
  const int A = dm.n_agents;
  float* obs = B.obs + (size_t)env * A * dm.obs_stride;
  float* rew = B.reward + (size_t)env * A;
  uint8_t* term = B.term + (size_t)env * (A + 1);
  uint8_t* trunc = B.trunc + (size_t)env * (A + 1);
  float reward[MJB_MAX_AGENTS];
  bool done[MJB_MAX_AGENTS];
  int opos[MJB_MAX_AGENTS];
  if (is_reset)  // data_store = {agent: {}} (mujoco_rl.py:312); the draw counter is not part of the store
    MJB_NOUNROLL
    for (int a = 0; a < A; a++) {
      MJB_NOUNROLL
      for (int k = 0; k < dm.store_i32; k++) if (k != MJB_STORE_I_DRAWS) si[a * dm.store_i32 + k] = 0;
      MJB_NOUNROLL
      for (int k = 0; k < dm.store_f32; k++) sf[a * dm.store_f32 + k] = 0.f;
    }
  MJB_NOUNROLL
  for (int a = 0; a < A; a++) { reward[a] = 0.f; done[a] = false; opos[a] = dm.obs_adr[a + 1] - dm.obs_adr[a]; }
  MJB_NOUNROLL
  for (int p = 0; p < dm.n_dynamics; p++) {
    const DevPlugin& dyn = dm.dynamics[p];
    MJB_NOUNROLL
    for (int a = 0; a < A; a++) {
      int* sia = si + a * dm.store_i32;
      float* sfa = sf + a * dm.store_f32;
      float* oa = obs + a * dm.obs_stride;
      if (dyn.kind == MJB_DYN_LANGUAGE) {
        int utt = (int)act[a * dm.act_stride + dyn.act_lo];  // int(): truncation toward zero
        sia[MJB_STORE_I_UTTERANCE] = utt; sia[MJB_STORE_I_HAS_UTTERANCE] = 1;
        int other = a == 0 ? 1 : 0;
        float val = 0.f;
        if (other < A && si[other * dm.store_i32 + MJB_STORE_I_HAS_UTTERANCE]) val = (float)si[other * dm.store_i32 + MJB_STORE_I_UTTERANCE];
        oa[opos[a]++] = val;
      } else if (dyn.kind == MJB_DYN_PICKUP) {
        int tgt = sia[MJB_STORE_I_TARGET];
        if (tgt == 0 && dm.n_targets > 0) {
          tgt = 1 + (int)(draw_u32(dm.seed, env, a, sia[MJB_STORE_I_DRAWS]++) % (uint32_t)dm.n_targets);
          sia[MJB_STORE_I_TARGET] = tgt;
        }
        if (tgt > 0) {
          float d = probe_dist(probe, dm.agent_probe[a], dm.target_probe[tgt - 1]);
          if (d < dyn.param[0]) {
            sia[MJB_STORE_I_INVENTORY] ^= 1;
            reward[a] += 1.f;
            tgt = 1 + (int)(draw_u32(dm.seed, env, a, sia[MJB_STORE_I_DRAWS]++) % (uint32_t)dm.n_targets);
            sia[MJB_STORE_I_TARGET] = tgt;
            sfa[MJB_STORE_F_DISTANCE] = probe_dist(probe, dm.agent_probe[a], dm.target_probe[tgt - 1]);
          }
          const float* tp = probe + 4 * dm.target_probe[tgt - 1];
          oa[opos[a]] = tp[0]; oa[opos[a] + 1] = tp[1]; oa[opos[a] + 2] = tp[2];
        } else {
          oa[opos[a]] = 0.f; oa[opos[a] + 1] = 0.f; oa[opos[a] + 2] = 0.f;
        }
        oa[opos[a] + 3] = (float)sia[MJB_STORE_I_INVENTORY];
        opos[a] += 4;
      }
    }
  }
  if (is_reset) {
    // everything the dynamics wrote is discarded (mujoco_rl.py:326-328); rewards / dones are not run
    MJB_NOUNROLL
    for (int a = 0; a < A; a++) {
      MJB_NOUNROLL
      for (int k = 0; k < dm.store_i32; k++) if (k != MJB_STORE_I_DRAWS) si[a * dm.store_i32 + k] = 0;
      MJB_NOUNROLL
      for (int k = 0; k < dm.store_f32; k++) sf[a * dm.store_f32 + k] = 0.f;
      rew[a] = 0.f; term[a] = 0; trunc[a] = 0;
    }
    term[A] = 0; trunc[A] = 0;
    *ts_io = 0;
    return;
  }
  MJB_NOUNROLL
  for (int p = 0; p < dm.n_rewards; p++) {
    const DevPlugin& rf = dm.rewards[p];
    MJB_NOUNROLL
    for (int a = 0; a < A; a++) {
      int* sia = si + a * dm.store_i32;
      float* sfa = sf + a * dm.store_f32;
      if (rf.kind == MJB_REW_TAG_DISTANCE) {
        int tgt = sia[MJB_STORE_I_TARGET];
        if (dm.n_targets <= 0) continue;
        if (tgt == 0) {
          tgt = 1 + (int)(draw_u32(dm.seed, env, a, sia[MJB_STORE_I_DRAWS]++) % (uint32_t)dm.n_targets);
          sia[MJB_STORE_I_TARGET] = tgt;
          sfa[MJB_STORE_F_DISTANCE] = probe_dist(probe, dm.agent_probe[a], dm.target_probe[tgt - 1]);
        } else {
          float d = probe_dist(probe, dm.agent_probe[a], dm.target_probe[tgt - 1]);
          reward[a] += (sfa[MJB_STORE_F_DISTANCE] - d) * rf.param[0];
          sfa[MJB_STORE_F_DISTANCE] = d;
        }
      } else if (rf.kind == MJB_REW_ANT) {
        float x_after = probe[4 * dm.agent_probe[a]];
        if (!sia[MJB_STORE_I_HAS_XPOS]) {
          sia[MJB_STORE_I_HAS_XPOS] = 1;
        } else {
          float cc = 0.f;
          MJB_NOUNROLL
          for (int u = 0; u < dm.nu; u++) cc += SF(ctrl)[u] * SF(ctrl)[u];
          // contact cost term: cfrc_ext is zero on these models (no force/acc sensors), SURVEY Q11
          reward[a] += (x_after - sfa[MJB_STORE_F_XPOS_BEFORE]) / dm.timestep - 0.5f * cc;
        }
        sfa[MJB_STORE_F_XPOS_BEFORE] = x_after;
      }
    }
  }
  int ts = *ts_io;
  uint8_t tr = ts >= dm.max_steps ? 1 : 0;
  MJB_NOUNROLL
  for (int a = 0; a <= A; a++) trunc[a] = tr;
  bool all = false;
  MJB_NOUNROLL
  for (int p = 0; p < dm.n_dones && !all; p++) {
    const DevPlugin& df = dm.dones[p];
    MJB_NOUNROLL
    for (int a = 0; a < A; a++)
      if (df.kind == MJB_DONE_DISTANCE_LE) done[a] = done[a] || (sf[a * dm.store_f32 + MJB_STORE_F_DISTANCE] <= df.param[0]);
    MJB_NOUNROLL
    for (int a = 0; a < A; a++) all = all || done[a];
  }
  MJB_NOUNROLL
  for (int a = 0; a < A; a++) { rew[a] = reward[a]; term[a] = done[a] ? 1 : 0; }
  term[A] = all ? 1 : 0;
  *ts_io = ts + 1;
}

// whole per-env pipeline.  Every lane of the warp calls this with the same arguments.
template <bool PHYS>
MJB_DEV void run_env(const Ctx& c, const mjb_buffers& B, int env, int mode, int skip_frames, const uint8_t* reset_mask) {
  float* probe = c.probe;
  const DevModel& dm = *c.dm;
  const int lane = c.lane;
  if (mode == MODE_RESET && reset_mask && !reset_mask[env]) return;
  float *qpos = SF(qpos), *qvel = SF(qvel), *qacc = SF(qacc), *ctrl = SF(ctrl), *sens = SF(sens);
  float* g_qpos = B.qpos + (size_t)env * dm.qpos_stride;
  float* g_qvel = B.qvel + (size_t)env * dm.qvel_stride;
  float* g_ctrl = B.ctrl + (size_t)env * dm.ctrl_stride;
  float* g_warm = B.warmstart + (size_t)env * dm.qvel_stride;
  float* g_sens = B.sensordata + (size_t)env * dm.sensor_stride;
  float* g_probe = B.probe + (size_t)env * dm.nprobe * 4;
  if (mode == MODE_RESET) {
    MJB_NOUNROLL
    for (int i = lane; i < dm.nq; i += 32) qpos[i] = CF(qpos0)[i];
    MJB_NOUNROLL
    for (int i = lane; i < dm.nv; i += 32) { qvel[i] = 0.f; qacc[i] = 0.f; }
    MJB_NOUNROLL
    for (int i = lane; i < dm.nu; i += 32) ctrl[i] = 0.f;
  } else {
    MJB_NOUNROLL
    for (int i = lane; i < dm.nq; i += 32) qpos[i] = g_qpos[i];
    MJB_NOUNROLL
    for (int i = lane; i < dm.nv; i += 32) { qvel[i] = g_qvel[i]; qacc[i] = g_warm[i]; }
    MJB_NOUNROLL
    for (int i = lane; i < dm.nu; i += 32) ctrl[i] = g_ctrl[i];
  }
  MJB_NOUNROLL
  // the plugin store, this env's step counter and the dynamic actions are fetched now so that their
  // global-memory latency hides behind the physics; they are consumed by the epilogue
  const int A_ = dm.n_agents;
  const bool epi = (mode == MODE_STEP || mode == MODE_RESET);
  int pre_si[2] = {0, 0};
  float pre_sf = 0.f, pre_act[2] = {0.f, 0.f};
  int pre_ts = 0;
  if (epi) {
    const int* gsi = B.store_i + (size_t)env * A_ * dm.store_i32;
    const float* gsf = B.store_f + (size_t)env * A_ * dm.store_f32;
    const float* gact = B.actions + (size_t)env * A_ * dm.act_stride;
    for (int r = 0; r < 2; r++) {
      int i = lane + 32 * r;
      if (i < A_ * dm.store_i32) pre_si[r] = gsi[i];
      if (i < A_ * dm.act_stride) pre_act[r] = gact[i];
    }
    if (lane < A_ * dm.store_f32) pre_sf = gsf[lane];
    if (lane == 0) pre_ts = B.timestep[env];
  }
  for (int i = lane; i < dm.nsensordata; i += 32) sens[i] = mode == MODE_RESET ? 0.f : g_sens[i];
  MJB_NOUNROLL
  for (int i = lane; i < 4 * dm.nprobe; i += 32) probe[i] = g_probe[i];
  MJB_SYNC();
  if (mode == MODE_STEP || mode == MODE_PHYSICS) {
    // apply_action (mujoco_parent.py:316-332): overwrite qvel (freeJoint) or ctrl
    const float* act = B.actions + (size_t)env * dm.n_agents * dm.act_stride;
    MJB_NOUNROLL
    for (int i = lane; i < dm.n_agents * dm.n_phys_act; i += 32) {
      float v = act[(i / dm.n_phys_act) * dm.act_stride + (i % dm.n_phys_act)];
      int idx = CI(act_index)[i];
      if (dm.free_joint) qvel[idx] = v; else ctrl[idx] = v;
    }
    MJB_SYNC();
  }
  int ncon = -1;
  // mj_forward (reset / forward modes) is one pass of the same loop body without integration
  const bool integrate = !(mode == MODE_FORWARD || mode == MODE_RESET);
  const int passes = integrate ? skip_frames : 1;
  int niter = 0;
  MJB_NOUNROLL
  if (PHYS)
    for (int f = 0; f < passes; f++) ncon = substep(c, f == passes - 1, integrate, &niter);
  if (ncon >= 0) {
    if (B.ncon && lane == 0) B.ncon[env] = ncon;
    if (B.niter && lane == 0) B.niter[env] = niter;
    if (B.contact_geom) {
      const uint32_t* pairs = CU(pair_pack);
      MJB_NOUNROLL
      for (int k = lane; k < dm.maxcon; k += 32) {
        int g1 = -1, g2 = -1;
        float dist = 0.f;
        if (k < ncon) {
          const float* r = SF(con) + CON_STRIDE * k;
          uint32_t pk = pairs[((const int*)r)[CON_PAIR]];
          g1 = pk & 0xfff; g2 = (pk >> 12) & 0xfff; dist = r[CON_DIST];
        }
        B.contact_geom[((size_t)env * dm.maxcon + k) * 2] = g1;
        B.contact_geom[((size_t)env * dm.maxcon + k) * 2 + 1] = g2;
        if (B.contact_dist) B.contact_dist[(size_t)env * dm.maxcon + k] = dist;
      }
    }
  }
  MJB_SYNC();
  MJB_NOUNROLL
  for (int i = lane; i < dm.nq; i += 32) g_qpos[i] = qpos[i];
  MJB_NOUNROLL
  for (int i = lane; i < dm.nv; i += 32) { g_qvel[i] = qvel[i]; g_warm[i] = qacc[i]; }
  MJB_NOUNROLL
  for (int i = lane; i < dm.nu; i += 32) g_ctrl[i] = ctrl[i];
  MJB_NOUNROLL
  for (int i = lane; i < dm.nsensordata; i += 32) g_sens[i] = sens[i];
  MJB_NOUNROLL
  for (int i = lane; i < 4 * dm.nprobe; i += 32) g_probe[i] = probe[i];
  if (mode != MODE_STEP && mode != MODE_RESET) return;
  // ---- epilogue: get_observations (mujoco_parent.py:380-392): sensordata(t) ++ qpos(t+h) ++ qvel(t+h)
  MJB_NOUNROLL
  for (int a = 0; a < dm.n_agents; a++) {
    float* oa = B.obs + ((size_t)env * dm.n_agents + a) * dm.obs_stride;
    int n = dm.obs_adr[a + 1] - dm.obs_adr[a];
    MJB_NOUNROLL
    for (int i = lane; i < n; i += 32) {
      int e = CI(obs_index)[dm.obs_adr[a] + i], kind = e >> 24, adr = e & 0xffffff;
      oa[i] = kind == 0 ? sens[adr] : (kind == 1 ? qpos[adr] : qvel[adr]);
    }
  }
  // stage the prefetched rows in shared memory (the contact Jacobian scratch is dead by now)
  int* s_si = (int*)SF(J);
  float* s_sf = SF(J) + MJB_MAX_AGENTS * MJB_STORE_I_COUNT;
  float* s_act = s_sf + MJB_MAX_AGENTS * MJB_STORE_F_COUNT;
  int* s_ts = (int*)(s_act + 64);
  for (int r = 0; r < 2; r++) {
    int i = lane + 32 * r;
    if (i < A_ * dm.store_i32) s_si[i] = pre_si[r];
    if (i < A_ * dm.act_stride) s_act[i] = pre_act[r];
  }
  if (lane < A_ * dm.store_f32) s_sf[lane] = pre_sf;
  if (lane == 0) *s_ts = pre_ts;
  MJB_SYNC();
  if (lane == 0) run_plugins(c, B, env, mode == MODE_RESET, probe, s_si, s_sf, s_act, s_ts);
  MJB_SYNC();
  int* gsi = B.store_i + (size_t)env * A_ * dm.store_i32;
  float* gsf = B.store_f + (size_t)env * A_ * dm.store_f32;
  for (int i = lane; i < A_ * dm.store_i32; i += 32) gsi[i] = s_si[i];
  if (lane < A_ * dm.store_f32) gsf[lane] = s_sf[lane];
  if (lane == 0) B.timestep[env] = *s_ts;
}

}  // namespace mjb
