// env_kernel.cuh — per-environment driver around the physics in step_kernel.cuh: state load /
// action scatter / substeps / state store, and the fused epilogue (observation gather, dynamics,
// reward, truncation, done — MuJoCo_Gym/mujoco_rl.py:243-289 and the reset path :291-331).
#pragma once
#include "step_kernel.cuh"
#if defined(MJB_HOST_EMU)
#define MJB_IMG_WAIT(c) ((void)0)
#else
#include "tma_prims.cuh"
// the model image travels into shared memory by one bulk copy per launch; an env's row loads do not need it, so the
// first env of a warp requests its rows first and waits for the image after
#define MJB_IMG_WAIT(c) do { if ((c).img_bar) mbar_wait((c).img_bar, 0); } while (0)
#endif

namespace mjb {

enum { MODE_STEP = 0, MODE_PHYSICS = 1, MODE_FORWARD = 2, MODE_RESET = 3 };

// agent loops of run_plugins: unrolled when the bound is small (registers), rolled in the general form (code size)
#if defined(MJB_HOST_EMU)
#define MJB_AGENT_LOOP
#else
#define MJB_AGENT_LOOP _Pragma("unroll (AMAX <= 2 ? 2 : 1)")
#endif

// counter-based draw replacing `random.randint` (README.md:154, Testing/Pick_Up_Dynamic.py:28,38)
MJB_HD uint32_t draw_u32(unsigned long long seed, uint32_t env, uint32_t agent, uint32_t counter) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (1ull + env) + 0xBF58476D1CE4E5B9ull * agent +
                         0x94D049BB133111EBull * counter;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}

// Distances, the reward arithmetic and the done comparison run in fp64 on the fp32 positions (the reference computes
// math.dist on float64 numpy views, mujoco_parent.py:428-449): given identical xipos the reward, cast to fp32, and the
// done flag are those of the reference's own arithmetic.  The stored distance keeps its fp64 value in two store_f
// columns (MJB_STORE_F_DISTANCE64, +1) next to the fp32 copy that `data_store[agent]["distance"]` shows.
MJB_DEV double probe_dist(const float* probe, int a, int b) {
  const double dx = (double)probe[4 * a] - (double)probe[4 * b], dy = (double)probe[4 * a + 1] - (double)probe[4 * b + 1],
               dz = (double)probe[4 * a + 2] - (double)probe[4 * b + 2];
  // plain IEEE evaluation, no fused multiply-add: (dx*dx + dy*dy) + dz*dz, so that any fp64 host evaluation of the same
  // expression (numpy, torch) gives the same bits
#if defined(MJB_HOST_EMU)
  volatile double xx = dx * dx, yy = dy * dy, zz = dz * dz;
  return sqrt((xx + yy) + zz);
#else
  return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
#endif
}
MJB_DEV double ld_dist(const float* sfa) {
  union { double d; float f[2]; } u;
  u.f[0] = sfa[MJB_STORE_F_DISTANCE64]; u.f[1] = sfa[MJB_STORE_F_DISTANCE64 + 1];
  return u.d;
}
MJB_DEV void st_dist(float* sfa, double d) {
  union { double d; float f[2]; } u;
  u.d = d;
  sfa[MJB_STORE_F_DISTANCE64] = u.f[0]; sfa[MJB_STORE_F_DISTANCE64 + 1] = u.f[1];
  sfa[MJB_STORE_F_DISTANCE] = (float)d;
}

// the dynamics / reward / done programme of one env, executed by one lane in the reference's order
// (dynamic outer, agent inner; then reward fn outer, agent inner; truncation; done fns with early exit)
// `obs` / `rew` / `term` / `trunc` are the env's own result rows ([A, obs_stride], [A], [A + 1], [A + 1]; global or
// shared memory), `ctrl` the env's ctrl row (read by the ant reward), `probe` its exported positions.
// AMAX bounds the agent loops at compile time: with a small AMAX (the tile kernel uses 2) the per-agent accumulators
// live in registers and the loops unroll; MJB_MAX_AGENTS is the general form.
template <int AMAX>
MJB_DEV void run_plugins(const DevModel& dm, const float* ctrl, float* obs, float* rew, uint8_t* term, uint8_t* trunc, int env, int copy,
                         bool is_reset, const float* probe, int* si, float* sf, const float* act, int* ts_io, const double* dtab = nullptr) {
  // `env` is the REAL environment, `copy` its slot inside the (possibly packed) virtual env of this warp
  const int A = dm.a1, ab = copy * dm.a1, tb = copy * dm.t1, n_targets = dm.t1;
  // agent - target distance: from the table the warp filled lane-parallel (dist_table), or evaluated here
  auto pdist = [&](int a, int t) { return dtab ? dtab[a * n_targets + t] : probe_dist(probe, dm.agent_probe[ab + a], dm.target_probe[tb + t]); };
  double reward[AMAX];
  bool done[AMAX];
  int opos[AMAX];
  if (is_reset)  // data_store = {agent: {}} (mujoco_rl.py:312); the draw counter is not part of the store
    MJB_AGENT_LOOP
    for (int a = 0; a < AMAX; a++) if (a < A) {
      MJB_NOUNROLL
      for (int k = 0; k < dm.store_i32; k++) if (k != MJB_STORE_I_DRAWS) si[a * dm.store_i32 + k] = 0;
      MJB_NOUNROLL
      for (int k = 0; k < dm.store_f32; k++) sf[a * dm.store_f32 + k] = 0.f;
    }
  MJB_AGENT_LOOP
  for (int a = 0; a < AMAX; a++) if (a < A) { reward[a] = 0.0; done[a] = false; opos[a] = dm.obs_adr[ab + a + 1] - dm.obs_adr[ab + a]; }
  MJB_NOUNROLL
  for (int p = 0; p < dm.n_dynamics; p++) {
    const DevPlugin& dyn = dm.dynamics[p];
    MJB_AGENT_LOOP
    for (int a = 0; a < AMAX; a++) if (a < A) {
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
        if (tgt == 0 && n_targets > 0) {
          tgt = 1 + (int)(draw_u32(dm.seed, env, a, sia[MJB_STORE_I_DRAWS]++) % (uint32_t)n_targets);
          sia[MJB_STORE_I_TARGET] = tgt;
        }
        if (tgt > 0) {
          const double d = pdist(a, tgt - 1);
          if (d < (double)dyn.param[0]) {
            sia[MJB_STORE_I_INVENTORY] ^= 1;
            reward[a] += 1.0;
            tgt = 1 + (int)(draw_u32(dm.seed, env, a, sia[MJB_STORE_I_DRAWS]++) % (uint32_t)n_targets);
            sia[MJB_STORE_I_TARGET] = tgt;
            st_dist(sfa, pdist(a, tgt - 1));
          }
          const float* tp = probe + 4 * dm.target_probe[tb + tgt - 1];
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
    MJB_AGENT_LOOP
    for (int a = 0; a < AMAX; a++) if (a < A) {
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
    MJB_AGENT_LOOP
    for (int a = 0; a < AMAX; a++) if (a < A) {
      int* sia = si + a * dm.store_i32;
      float* sfa = sf + a * dm.store_f32;
      if (rf.kind == MJB_REW_TAG_DISTANCE) {
        int tgt = sia[MJB_STORE_I_TARGET];
        if (n_targets <= 0) continue;
        if (tgt == 0) {
          tgt = 1 + (int)(draw_u32(dm.seed, env, a, sia[MJB_STORE_I_DRAWS]++) % (uint32_t)n_targets);
          sia[MJB_STORE_I_TARGET] = tgt;
          st_dist(sfa, pdist(a, tgt - 1));
        } else {
          const double d = pdist(a, tgt - 1);
          reward[a] += (ld_dist(sfa) - d) * (double)rf.param[0];
          st_dist(sfa, d);
        }
      } else if (rf.kind == MJB_REW_ANT) {
        float x_after = probe[4 * dm.agent_probe[ab + a]];
        if (!sia[MJB_STORE_I_HAS_XPOS]) {
          sia[MJB_STORE_I_HAS_XPOS] = 1;
        } else {
          double cc = 0.0;
          MJB_NOUNROLL
          for (int u = 0; u < dm.nu1; u++) cc += (double)ctrl[u] * (double)ctrl[u];
          // contact cost term: cfrc_ext is zero on these models (no force / acc sensors); the Python
          // layer refuses to fuse this reward for models with accelerometers, where MuJoCo fills cfrc_ext
          reward[a] += ((double)x_after - (double)sfa[MJB_STORE_F_XPOS_BEFORE]) / dm.timestep_d - 0.5 * cc;
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
    MJB_AGENT_LOOP
    for (int a = 0; a < AMAX; a++) if (a < A)
      if (df.kind == MJB_DONE_DISTANCE_LE) done[a] = done[a] || (ld_dist(sf + a * dm.store_f32) <= (double)df.param[0]);
    MJB_AGENT_LOOP
    for (int a = 0; a < AMAX; a++) if (a < A) all = all || done[a];
  }
  MJB_AGENT_LOOP
  for (int a = 0; a < AMAX; a++) if (a < A) { rew[a] = (float)reward[a]; term[a] = done[a] ? 1 : 0; }
  term[A] = all ? 1 : 0;
  *ts_io = ts + 1;
}

// element i of a virtual per-copy array -> (copy, element); no division in the unpacked case
MJB_DEV void split_copy(int i, int n1, int K, int& k, int& j) {
  if (K == 1) { k = 0; j = i; }
  else { k = i / n1; j = i - k * n1; }
}

// whole per-env pipeline.  Every lane of the warp calls this with the same arguments.  `venv` is the virtual
// environment of this warp: dm.pack real environments (venv * pack + copy) presented to the physics as one
// model with `pack` independent copies (csrc/replicate.h); pack == 1 is the plain case.  Virtual state index i
// of a per-copy array of length n1 maps to copy i / n1, element i % n1 of that real env's HBM row.
template <bool PHYS, bool PACKED>
MJB_DEV void run_env(const Ctx& c_in, const mjb_buffers& B, int venv, int num_envs, int mode, int skip_frames,
                     const uint8_t* reset_mask) {
  Ctx c = c_in;
  float* probe = c.probe;
  const DevModel& dm = *c.dm;
  const int lane = c.lane, K = PACKED ? dm.pack : 1, e0 = venv * K;   // K == 1 folds all copy arithmetic away
  c.probe_quat = (!PACKED && B.probe_quat) ? B.probe_quat + (size_t)e0 * dm.np1 * 4 : nullptr;
  // which copies hold a real env, and which of them this launch updates
  uint32_t live = 0, upd = 0;
  for (int k = 0; k < K; k++)
    if (e0 + k < num_envs) {
      live |= 1u << k;
      if (!(mode == MODE_RESET && reset_mask && !reset_mask[e0 + k])) upd |= 1u << k;
    }
  if (!upd) return;
  // the loads below read the image only for copies that start from qpos0 or hold no env
  if (mode == MODE_RESET || live != (K >= 32 ? 0xffffffffu : (1u << K) - 1u)) MJB_IMG_WAIT(c);
  float *qpos = SF(qpos), *qvel = SF(qvel), *qacc = SF(qacc), *ctrl = SF(ctrl), *sens = SF(sens);
  auto fresh = [&](int k) { return mode == MODE_RESET && ((upd >> k) & 1u); };   // copy starts from qpos0
  auto has = [&](int k) { return ((live >> k) & 1u) != 0; };
  MJB_NOUNROLL
  for (int i = lane; i < dm.nq; i += 32) {
    int k, j;
    split_copy(i, dm.nq1, K, k, j);
    if (fresh(k) || !has(k)) qpos[i] = CF(qpos0)[i];
    else MJB_G2S(&qpos[i], &B.qpos[(size_t)(e0 + k) * dm.qpos_stride + j]);
  }
  MJB_NOUNROLL
  for (int i = lane; i < dm.nv; i += 32) {
    int k, j;
    split_copy(i, dm.nv1, K, k, j);
    bool z = fresh(k) || !has(k);
    if (z) { qvel[i] = 0.f; qacc[i] = 0.f; }
    else {
      MJB_G2S(&qvel[i], &B.qvel[(size_t)(e0 + k) * dm.qvel_stride + j]);
      MJB_G2S(&qacc[i], &B.warmstart[(size_t)(e0 + k) * dm.qvel_stride + j]);
    }
  }
  MJB_NOUNROLL
  for (int i = lane; i < dm.nu; i += 32) {
    int k, j;
    split_copy(i, dm.nu1, K, k, j);
    if (fresh(k) || !has(k)) ctrl[i] = 0.f;
    else MJB_G2S(&ctrl[i], &B.ctrl[(size_t)(e0 + k) * dm.ctrl_stride + j]);
  }
  MJB_NOUNROLL
  for (int i = lane; i < dm.nsensordata; i += 32) {
    int k, j;
    split_copy(i, dm.ns1, K, k, j);
    if (fresh(k) || !has(k)) sens[i] = 0.f;
    else MJB_G2S(&sens[i], &B.sensordata[(size_t)(e0 + k) * dm.sensor_stride + j]);
  }
  // exported positions: virtual probe order is [agents of every copy ..., targets of every copy ...]
  auto probe_slot = [&](int p, int& k, int& p1) {
    if (K == 1) { k = 0; p1 = p; }
    else if (p < K * dm.a1) { k = p / dm.a1; p1 = p - k * dm.a1; }
    else { int q = p - K * dm.a1; k = q / dm.t1; p1 = dm.a1 + (q - k * dm.t1); }
  };
  MJB_NOUNROLL
  for (int i = lane; i < 4 * dm.nprobe; i += 32) {
    int k, p1;
    probe_slot(i >> 2, k, p1);
    if (has(k)) MJB_G2S(&probe[i], &B.probe[((size_t)(e0 + k) * dm.np1 + p1) * 4 + (i & 3)]);
    else probe[i] = 0.f;
  }
  // unpacked case: fetch the plugin store rows, dynamic actions and step counter of the env now, so that their
  // global-memory latency overlaps the state loads / the physics; consumed by the epilogue
  const bool epi = (mode == MODE_STEP || mode == MODE_RESET);
  int pre_si[2] = {0, 0}, pre_ts = 0;
  float pre_sf = 0.f, pre_act[2] = {0.f, 0.f};
  if (epi && K == 1) {
    const int* gsi = B.store_i + (size_t)e0 * dm.a1 * dm.store_i32;
    const float* gsf = B.store_f + (size_t)e0 * dm.a1 * dm.store_f32;
    const float* gact = B.actions + (size_t)e0 * dm.a1 * dm.act_stride;
    for (int r = 0; r < 2; r++) {
      int i = lane + 32 * r;
      if (i < dm.a1 * dm.store_i32) pre_si[r] = gsi[i];
      if (i < dm.a1 * dm.act_stride) pre_act[r] = gact[i];
    }
    if (lane < dm.a1 * dm.store_f32) pre_sf = gsf[lane];
    if (lane == 0) pre_ts = B.timestep[e0];
  }
  MJB_IMG_WAIT(c);
  MJB_G2S_WAIT();
  if (MJB_UNLIKELY(mode == MODE_RESET && dm.reset_noise > 0.f)) {
    // optional decorrelated starts (off by default: the reference always restarts at qpos0).  The draw is keyed by
    // (seed, env, joint / dof) and by how the previous episode ended, so that it differs from reset to reset.
    MJB_SYNC();
    MJB_NOUNROLL
    for (int j = lane; j < dm.njnt; j += 32) {
      const int k = K == 1 ? 0 : j / dm.njnt1;
      if (!fresh(k) || !has(k) || CI(jnt_type)[j] == MJB_JNT_FREE) continue;
      const uint32_t env = (uint32_t)(e0 + k);
      const uint32_t ep = (uint32_t)B.timestep[env] ^ MJB_F2U(B.qpos[(size_t)env * dm.qpos_stride]) ^
                          (MJB_F2U(B.qvel[(size_t)env * dm.qvel_stride]) << 13);
      const float u = (float)(draw_u32(dm.seed ^ 0x5eedc0deULL, env, 0x4000u + (uint32_t)(j - k * dm.njnt1), ep) >> 8) * (1.f / 16777216.f);
      qpos[CI(jnt_qposadr)[j]] += dm.reset_noise * (2.f * u - 1.f);
    }
    MJB_NOUNROLL
    for (int i = lane; i < dm.nv; i += 32) {
      int k, j;
      split_copy(i, dm.nv1, K, k, j);
      if (!fresh(k) || !has(k)) continue;
      const uint32_t env = (uint32_t)(e0 + k);
      const uint32_t ep = (uint32_t)B.timestep[env] ^ MJB_F2U(B.qpos[(size_t)env * dm.qpos_stride]) ^
                          (MJB_F2U(B.qvel[(size_t)env * dm.qvel_stride]) << 13);
      const float u = (float)(draw_u32(dm.seed ^ 0x5eedc0deULL, env, 0x8000u + (uint32_t)j, ep) >> 8) * (1.f / 16777216.f);
      qvel[i] = dm.reset_noise * (2.f * u - 1.f);
    }
  }
  MJB_SYNC();
  if (mode == MODE_STEP || mode == MODE_PHYSICS) {
    // apply_action (mujoco_parent.py:316-332): overwrite qvel (freeJoint) or ctrl
    const int n_act = dm.n_agents * dm.n_phys_act;
    const bool from_regs = K == 1 && epi && dm.a1 * dm.act_stride <= 64;   // the action rows were prefetched above
    MJB_NOUNROLL
    for (int i0 = 0; i0 < n_act; i0 += 32) {
      const int i = i0 + lane;
      const bool on = i < n_act;
      int av = on ? i / dm.n_phys_act : 0, k, a1;
      split_copy(av, dm.a1, K, k, a1);
      const int off = a1 * dm.act_stride + (on ? i - av * dm.n_phys_act : 0);
      float v;
      if (from_regs) {
        const float lo = MJB_SHFL(pre_act[0], off & 31), hi = MJB_SHFL(pre_act[1], off & 31);
        v = off < 32 ? lo : hi;
      } else {
        v = (on && has(k)) ? B.actions[(size_t)(e0 + k) * dm.a1 * dm.act_stride + off] : 0.f;
      }
      if (on && has(k)) {
        int idx = CI(act_index)[i];
        if (dm.free_joint) qvel[idx] = v; else ctrl[idx] = v;
      }
    }
    MJB_SYNC();
  }
  MJB_PH(c, PH_LOAD);
  int ncon = -1;
  // mj_forward (reset / forward modes) is one pass of the same loop body without integration
  const bool integrate = !(mode == MODE_FORWARD || mode == MODE_RESET);
  const int passes = integrate ? skip_frames : 1;
  int niter = 0;
  if (PHYS) {
    int dropped[MJB_MAX_PACK] = {0, 0, 0, 0};
    MJB_NOUNROLL
    for (int f = 0; f < passes; f++) ncon = substep(c, f == passes - 1, integrate, &niter, dropped);
    if (B.ncon_dropped && lane == 0) {   // cumulative per real env: stays 0 while no contact was ever dropped
#pragma unroll
      for (int k = 0; k < MJB_MAX_PACK; k++)
        if (k < K && dropped[k] && ((upd >> k) & 1u)) B.ncon_dropped[e0 + k] += dropped[k];
    }
    if (integrate) {
      // MuJoCo's mj_checkPos / mj_checkVel: a non-finite or absurd state resets that env (here: that copy)
      MJB_NOUNROLL
      for (int k = 0; k < K; k++) {
        bool bad = false;
        for (int j = lane; j < dm.nq1; j += 32) { float x = qpos[k * dm.nq1 + j]; bad |= !(fabsf(x) < 1e10f); }
        for (int j = lane; j < dm.nv1; j += 32) { float x = qvel[k * dm.nv1 + j]; bad |= !(fabsf(x) < 1e10f); }
        if (MJB_UNLIKELY(MJB_BALLOT(bad))) {   // warp-uniform
          for (int j = lane; j < dm.nq1; j += 32) qpos[k * dm.nq1 + j] = CF(qpos0)[k * dm.nq1 + j];
          for (int j = lane; j < dm.nv1; j += 32) { qvel[k * dm.nv1 + j] = 0.f; qacc[k * dm.nv1 + j] = 0.f; }
          if (B.nreset && lane == 0 && ((upd >> k) & 1u)) B.nreset[e0 + k] += 1;
        }
      }
      MJB_SYNC();
    }
  }
  MJB_PH(c, PH_INTEGRATE);
  if (ncon >= 0 && (B.ncon || B.contact_geom) && K == 1) {
    const uint32_t* pairs = CU(pair_pack);
    if (B.ncon && lane == 0) B.ncon[e0] = ncon;
    if (B.niter && lane == 0) B.niter[e0] = niter;
    if (B.contact_geom) {
      MJB_NOUNROLL
      for (int k = lane; k < dm.maxcon; k += 32) {
        int g1 = -1, g2 = -1;
        float dist = 0.f;
        if (k < ncon) {
          const float* r = SF(con) + CON_STRIDE * k;
          uint32_t pk = pairs[((const int*)r)[CON_PAIR]];
          g1 = pk & 0xfff; g2 = (pk >> 12) & 0xfff; dist = r[CON_DIST];
        }
        B.contact_geom[((size_t)e0 * dm.maxcon + k) * 2] = g1;
        B.contact_geom[((size_t)e0 * dm.maxcon + k) * 2 + 1] = g2;
        if (B.contact_dist) B.contact_dist[(size_t)e0 * dm.maxcon + k] = dist;
      }
    }
  } else if (ncon >= 0 && (B.ncon || B.contact_geom)) {
    // packed: contact records per real env; a contact belongs to the copy its first geom lives in
    const uint32_t* pairs = CU(pair_pack);
    MJB_NOUNROLL
    for (int k = lane; k < K; k += 32) {
      if (!((upd >> k) & 1u)) continue;
      int n = 0;
      MJB_NOUNROLL
      for (int i = 0; i < ncon; i++) {
        const float* r = SF(con) + CON_STRIDE * i;
        uint32_t pk = pairs[((const int*)r)[CON_PAIR]];
        int g1 = pk & 0xfff, g2 = (pk >> 12) & 0xfff;
        if (g1 / dm.ngeom1 != k) continue;
        if (n < dm.maxcon1 && B.contact_geom) {
          size_t o = (size_t)(e0 + k) * dm.maxcon1 + n;
          B.contact_geom[2 * o] = g1 - k * dm.ngeom1; B.contact_geom[2 * o + 1] = g2 - k * dm.ngeom1;
          if (B.contact_dist) B.contact_dist[o] = r[CON_DIST];
        }
        n++;
      }
      if (B.contact_geom)
        for (int i = n; i < dm.maxcon1; i++) {
          size_t o = (size_t)(e0 + k) * dm.maxcon1 + i;
          B.contact_geom[2 * o] = -1; B.contact_geom[2 * o + 1] = -1;
          if (B.contact_dist) B.contact_dist[o] = 0.f;
        }
      if (B.ncon) B.ncon[e0 + k] = n < dm.maxcon1 ? n : dm.maxcon1;
      if (B.niter) B.niter[e0 + k] = niter;
    }
  }
  MJB_SYNC();
  auto writes = [&](int k) { return ((upd >> k) & 1u) != 0; };
  MJB_NOUNROLL
  for (int i = lane; i < dm.nq; i += 32) {
    int k, j;
    split_copy(i, dm.nq1, K, k, j);
    if (writes(k)) B.qpos[(size_t)(e0 + k) * dm.qpos_stride + j] = qpos[i];
  }
  MJB_NOUNROLL
  for (int i = lane; i < dm.nv; i += 32) {
    int k, j;
    split_copy(i, dm.nv1, K, k, j);
    if (writes(k)) { B.qvel[(size_t)(e0 + k) * dm.qvel_stride + j] = qvel[i]; B.warmstart[(size_t)(e0 + k) * dm.qvel_stride + j] = qacc[i]; }
  }
  MJB_NOUNROLL
  for (int i = lane; i < dm.nu; i += 32) {
    int k, j;
    split_copy(i, dm.nu1, K, k, j);
    if (writes(k)) B.ctrl[(size_t)(e0 + k) * dm.ctrl_stride + j] = ctrl[i];
  }
  MJB_NOUNROLL
  for (int i = lane; i < dm.nsensordata; i += 32) {
    int k, j;
    split_copy(i, dm.ns1, K, k, j);
    if (writes(k)) B.sensordata[(size_t)(e0 + k) * dm.sensor_stride + j] = sens[i];
  }
  MJB_NOUNROLL
  for (int i = lane; i < 4 * dm.nprobe; i += 32) {
    int k, p1;
    probe_slot(i >> 2, k, p1);
    if (writes(k)) B.probe[((size_t)(e0 + k) * dm.np1 + p1) * 4 + (i & 3)] = probe[i];
  }
  MJB_PH(c, PH_STORE);
  if (mode != MODE_STEP && mode != MODE_RESET) return;
  // ---- epilogue: get_observations (mujoco_parent.py:380-392): sensordata(t) ++ qpos(t+h) ++ qvel(t+h)
  MJB_NOUNROLL
  for (int av = 0; av < dm.n_agents; av++) {
    int k, a1;
    split_copy(av, dm.a1, K, k, a1);
    if (!writes(k)) continue;
    float* oa = B.obs + ((size_t)(e0 + k) * dm.a1 + a1) * dm.obs_stride;
    int n = dm.obs_adr[av + 1] - dm.obs_adr[av];
    MJB_NOUNROLL
    for (int i = lane; i < n; i += 32) {
      int e = CI(obs_index)[dm.obs_adr[av] + i], kind = e >> 24, adr = e & 0xffffff;
      oa[i] = kind == 0 ? sens[adr] : (kind == 1 ? qpos[adr] : qvel[adr]);
    }
  }
  MJB_PH(c, PH_EPI_OBS);
  // plugins per real env: its store rows, dynamic actions and step counter are staged in shared memory (the
  // contact Jacobian scratch is dead by now) so that one lane can run the reference-order programme on them
  int* s_si = (int*)SF(J);
  float* s_sf = SF(J) + MJB_MAX_AGENTS * MJB_STORE_I_COUNT;
  float* s_act = s_sf + MJB_MAX_AGENTS * MJB_STORE_F_COUNT;
  int* s_ts = (int*)(s_act + 64);
  double* s_dtab = (double*)(s_ts + 4);   // [a1, t1] agent - target distances of the copy being processed
  MJB_NOUNROLL
  for (int k = 0; k < K; k++) {
    if (!writes(k)) continue;
    const int env = e0 + k, A1 = dm.a1;
    int* gsi = B.store_i + (size_t)env * A1 * dm.store_i32;
    float* gsf = B.store_f + (size_t)env * A1 * dm.store_f32;
    const float* gact = B.actions + (size_t)env * A1 * dm.act_stride;
    MJB_SYNC();
    if (K == 1) {
      for (int r = 0; r < 2; r++) {
        int i = lane + 32 * r;
        if (i < A1 * dm.store_i32) s_si[i] = pre_si[r];
        if (i < A1 * dm.act_stride) s_act[i] = pre_act[r];
      }
      if (lane < A1 * dm.store_f32) s_sf[lane] = pre_sf;
      if (lane == 0) *s_ts = pre_ts;
    } else {
      for (int i = lane; i < A1 * dm.store_i32; i += 32) s_si[i] = gsi[i];
      for (int i = lane; i < A1 * dm.act_stride; i += 32) s_act[i] = gact[i];
      if (lane < A1 * dm.store_f32) s_sf[lane] = gsf[lane];
      if (lane == 0) *s_ts = B.timestep[env];
    }
    // every agent - target distance of this env, one pair per lane (fp64: run_plugins then only looks them up)
    MJB_NOUNROLL
    for (int i = lane; i < A1 * dm.t1; i += 32) {
      const int a = i / dm.t1, t = i - a * dm.t1;
      s_dtab[i] = probe_dist(probe, dm.agent_probe[k * A1 + a], dm.target_probe[k * dm.t1 + t]);
    }
    MJB_SYNC();
    MJB_PH(c, PH_EPI_STAGE);
    if (lane == 0) {
      // up to two agents (every level of the reference): the per-agent accumulators of the programme live in registers
      if (MJB_LIKELY(A1 <= 2))
        run_plugins<2>(dm, SF(ctrl) + k * dm.nu1, B.obs + (size_t)env * A1 * dm.obs_stride, B.reward + (size_t)env * A1, B.term + (size_t)env * (A1 + 1),
                       B.trunc + (size_t)env * (A1 + 1), env, k, mode == MODE_RESET, probe, s_si, s_sf, s_act, s_ts, s_dtab);
      else
        run_plugins<MJB_MAX_AGENTS>(dm, SF(ctrl) + k * dm.nu1, B.obs + (size_t)env * A1 * dm.obs_stride, B.reward + (size_t)env * A1,
                                    B.term + (size_t)env * (A1 + 1), B.trunc + (size_t)env * (A1 + 1), env, k, mode == MODE_RESET, probe, s_si, s_sf,
                                    s_act, s_ts, s_dtab);
    }
    MJB_SYNC();
    MJB_PH(c, PH_EPI_PLUG);
    for (int i = lane; i < A1 * dm.store_i32; i += 32) gsi[i] = s_si[i];
    if (lane < A1 * dm.store_f32) gsf[lane] = s_sf[lane];
    if (lane == 0) B.timestep[env] = *s_ts;
  }
  MJB_PH(c, PH_EPILOGUE);
}

}  // namespace mjb
