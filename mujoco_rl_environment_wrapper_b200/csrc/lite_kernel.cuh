// lite_kernel.cuh — MuJoCoRL.step WITHOUT physics: `skipFrames = 0`, the literal setting of the reference's own
// benchmarks (benchmarking/fps_gym/fps_custom_env.py:39-48).  apply_action scatters the actions into
// qvel (freeJoint) or ctrl and calls mj_step zero times (mujoco_parent.py:316-336); what remains is the observation
// gather (mujoco_parent.py:380-392) and the dynamics / reward / truncation / done loops (mujoco_rl.py:243-289).
//
// That step is pure data movement, so it is laid out for HBM bandwidth instead of for the physics:
//   * a CTA owns a TILE of consecutive envs; every array's rows of the tile are ONE contiguous block in HBM (env-major
//     rows), fetched by one 1-D TMA bulk copy each (cp.async.bulk -> shared memory, one mbarrier for the tile);
//   * actions scatter, observation gather and the plugin programme run out of shared memory (thread = env for the
//     scatter and the reference-order plugin programme, warp = env / lane = word for the gather);
//   * observations are gathered straight into HBM (consecutive threads write consecutive words of a row); the other
//     results leave as 128-bit stores of whole row blocks; only what the step changes is written (qvel or ctrl, obs,
//     reward, flags, data_store rows, step counter);
//   * no model image staging: the two index tables (built on the host) ride in on the first tile's bulk copies.
// Several small CTAs per SM keep tens of KB of loads in flight per SM.
#pragma once
#include <cstdio>
#include <vector>

#include "env_kernel.cuh"
#include "tma_prims.cuh"

namespace mjb {

#define LITE_THREADS 128

struct LiteLayout {
  int tile;                       // envs per CTA tile (multiple of 16)
  int use_ctrl, use_sens, use_probe;
  int per_obs;                    // floats of one env's obs rows = A * obs_stride
  // word offsets of the tile arrays in dynamic shared memory (after the 16-byte barrier slot)
  int o_qpos, o_qvel, o_ctrl, o_sens, o_act, o_probe, o_si, o_sf, o_ts, o_obs, o_rew, o_term, o_trunc;
  int o_gather, o_actidx;         // per-CTA copies of the two index tables (one entry per obs-row word / physical action)
  int table_words;                // gather table + action table, contiguous (o_gather .. ), multiple of 4
  int o_dm, dm_bytes;             // shared-memory copy of the DevModel header (the plugin programme reads it a lot)
  int A, qs, vs, cs, ss, as, ps, sis, sfs;   // agents and row strides in words (what thread 0 needs to issue the loads)
  int words;
  int profile;
};

// What the start of a tile needs, packed into the FIRST two cache lines of the kernel parameters: a cold launch then
// waits for one constant-cache fill before its loads are in flight, not for a dozen scattered parameter reads.
// Copy k: `fixed` bytes from `base` (index tables, DevModel header: first tile of a CTA only) or `row_bytes` per env from
// `base + env * row_bytes`, to shared-memory byte offset `dst_off`.
struct LiteCopy {
  const char* base;
  uint32_t row_bytes, fixed, dst_off, pad;
};
struct LiteIssue {
  uint32_t sum_row_bytes, first_bytes;
  int tile, active, env_base, ncopy;
  int profile, pad;               // MJB_LITE_PROFILE=1: CTA 0 prints the clock cycles of its phases (debug)
  LiteCopy c[10];
};

inline LiteLayout make_lite_layout(const DevModel& dm, int tile) {
  LiteLayout L{};
  const int A = dm.a1;
  L.tile = tile;
  L.use_ctrl = dm.nu > 0;
  L.use_sens = dm.nsensordata > 0;
  L.use_probe = dm.np1 > 0;
  L.per_obs = A * dm.obs_stride;
  int o = 0;
  auto take = [&](int words_per_env) { int at = o; o += ((tile * words_per_env + 3) / 4) * 4; return at; };
  L.o_qpos = take(dm.qpos_stride); L.o_qvel = take(dm.qvel_stride);
  L.o_ctrl = take(L.use_ctrl ? dm.ctrl_stride : 0); L.o_sens = take(L.use_sens ? dm.sensor_stride : 0);
  L.o_act = take(A * dm.act_stride); L.o_probe = take(L.use_probe ? dm.np1 * 4 : 0);
  L.o_si = take(A * dm.store_i32); L.o_sf = take(A * dm.store_f32); L.o_ts = take(1);
  L.o_obs = o; L.o_rew = take(A);   // observations are not staged: gathered straight into HBM
  L.o_term = o; o += ((tile * (A + 1) + 15) / 16) * 4;
  L.o_trunc = o; o += ((tile * (A + 1) + 15) / 16) * 4;
  L.o_gather = o; o += ((L.per_obs + 3) / 4) * 4;
  L.o_actidx = o; o += ((A * dm.n_phys_act + 3) / 4) * 4;
  L.table_words = o - L.o_gather;
  L.o_dm = o; L.dm_bytes = (int)((sizeof(DevModel) + 15) / 16 * 16); o += L.dm_bytes / 4;
  L.A = A; L.qs = dm.qpos_stride; L.vs = dm.qvel_stride; L.cs = dm.ctrl_stride; L.ss = dm.sensor_stride; L.as = A * dm.act_stride;
  L.ps = dm.np1 * 4; L.sis = A * dm.store_i32; L.sfs = A * dm.store_f32;
  L.words = o;
  return L;
}

__device__ __noinline__ void lite_report(const long long* t) {
  printf("k_lite cta0 cycles: tables+issue %lld | load wait %lld | scatter %lld | gather %lld | plugins %lld | stores %lld\n", t[1] - t[0], t[2] - t[1],
         t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5]);
}

// the two index tables in the kernel's shared-memory order: entry r of an env's obs rows -> (kind << 24 | address) or
// -1 for row padding, then the qvel / ctrl address of every physical action
inline std::vector<int> lite_tables(const DevModel& dm, const LiteLayout& L, const uint32_t* image) {
  std::vector<int> t(L.table_words, 0);
  const int* obs_index = reinterpret_cast<const int*>(image + dm.off[IF_obs_index]);
  const int* act_index = reinterpret_cast<const int*>(image + dm.off[IF_act_index]);
  for (int r = 0; r < L.per_obs; r++) {
    const int a = r / dm.obs_stride, k = r - a * dm.obs_stride;
    t[r] = k < dm.obs_adr[a + 1] - dm.obs_adr[a] ? obs_index[dm.obs_adr[a] + k] : -1;
  }
  for (int k = 0; k < dm.a1 * dm.n_phys_act; k++) t[(L.o_actidx - L.o_gather) + k] = act_index[k];
  return t;
}

// 128-bit copy of `words` floats (multiple of 4; both sides 16-byte aligned) from shared to global memory
__device__ __forceinline__ void tile_store(float* __restrict__ dst, const float* src, int words, int tid, int nthreads) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = tid; i < (words >> 2); i += nthreads) d4[i] = s4[i];
}

// Kernel parameters: the issue plan first (LiteIssue), then buffers and layout; the DevModel header travels through
// global memory into shared memory with the first tile.
__global__ void __launch_bounds__(LITE_THREADS, 8) k_lite(const __grid_constant__ LiteIssue I, const __grid_constant__ mjb_buffers B,
                                                         const __grid_constant__ LiteLayout L) {
  extern __shared__ __align__(128) uint32_t lite_smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(lite_smem);
  float* sm = reinterpret_cast<float*>(lite_smem + 4);
  const int tid = threadIdx.x, nthreads = blockDim.x, T = I.tile, active = I.active, env_base = I.env_base;
  long long tk[7], tx[4] = {0, 0, 0, 0};
  const bool prof = I.profile && blockIdx.x == 0 && tid == 0;
  if (prof) tk[0] = clock64();
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (prof) tx[0] = clock64();
  float *s_qpos = sm + L.o_qpos, *s_qvel = sm + L.o_qvel, *s_ctrl = sm + L.o_ctrl, *s_sens = sm + L.o_sens, *s_act = sm + L.o_act;
  float *s_probe = sm + L.o_probe, *s_sf = sm + L.o_sf, *s_rew = sm + L.o_rew;
  int *s_si = reinterpret_cast<int*>(sm + L.o_si), *s_ts = reinterpret_cast<int*>(sm + L.o_ts);
  uint8_t *s_term = reinterpret_cast<uint8_t*>(sm + L.o_term), *s_trunc = reinterpret_cast<uint8_t*>(sm + L.o_trunc);
  const DevModel& dm = *reinterpret_cast<const DevModel*>(sm + L.o_dm);   // valid after the first barrier wait
  int *s_gather = reinterpret_cast<int*>(sm + L.o_gather), *s_actidx = reinterpret_cast<int*>(sm + L.o_actidx);
  const int A = L.A;
  bool first = true;   // the first tile's barrier also covers the two index tables (built on the host, lite_tables())
  const int ntiles = (active + T - 1) / T;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, phase ^= 1u) {
    const int n = min(T, active - tile * T);          // envs of this tile
    const size_t e0 = (size_t)env_base + (size_t)tile * T;
    if (tid < 32) {
      // the previous tile's shared-memory reads / writes (generic proxy) are ordered before the bulk copies (async proxy);
      // lane k of warp 0 issues copy k, lane 0 also posts the byte count (complete_tx may legally run ahead of it)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (prof) tx[1] = clock64();
      if (tid == 0) mbar_expect_tx(bar, (uint32_t)n * I.sum_row_bytes + (first ? I.first_bytes : 0u));
      if (prof) tx[2] = clock64();
      if (tid < I.ncopy) {
        const LiteCopy cp = I.c[tid];
        const uint32_t nb = cp.fixed ? (first ? cp.fixed : 0u) : (uint32_t)n * cp.row_bytes;
        const char* src = cp.fixed ? cp.base : cp.base + e0 * cp.row_bytes;
        if (nb) bulk_g2s(reinterpret_cast<char*>(sm) + cp.dst_off, src, nb, bar);
      }
      if (prof) tx[3] = clock64();
    }
    for (int e = tid; e < n; e += nthreads) s_ts[e] = B.timestep[e0 + e];   // 4 bytes per env: not worth a bulk copy (alignment of odd ranges)
    first = false;
    if (prof) tk[1] = clock64();
    mbar_wait(bar, phase);
    if (prof) tk[2] = clock64();
    // apply_action (mujoco_parent.py:316-332): overwrite qvel (freeJoint) or ctrl; thread = env
    for (int e = tid; e < n; e += nthreads) {
      float* dst = dm.free_joint ? s_qvel + e * dm.qvel_stride : s_ctrl + e * dm.ctrl_stride;
      const float* src = s_act + e * A * dm.act_stride;
      for (int a = 0, k = 0; a < A; a++)
        for (int j = 0; j < dm.n_phys_act; j++, k++) dst[s_actidx[k]] = src[a * dm.act_stride + j];
    }
    __syncthreads();
    if (prof) tk[3] = clock64();
    // get_observations (mujoco_parent.py:380-392): thread = (word r of the obs rows, group of eight envs): one table read,
    // eight independent loads
    const int ngrp = (n + 7) >> 3;
    for (int i = tid; i < L.per_obs * ngrp; i += nthreads) {
      const int g = i / L.per_obs, r = i - g * L.per_obs, e = g << 3, m = min(8, n - e);
      const int ent = s_gather[r], kind = ent >> 24, adr = ent & 0xffffff;
      float* dst = B.obs + (e0 + e) * L.per_obs + r;   // consecutive threads -> consecutive words of a row: full-line stores
      if (ent < 0) {
        for (int u = 0; u < m; u++) dst[u * L.per_obs] = 0.f;
      } else {
        const int st = kind == 0 ? dm.sensor_stride : (kind == 1 ? dm.qpos_stride : dm.qvel_stride);
        const float* src = (kind == 0 ? s_sens : (kind == 1 ? s_qpos : s_qvel)) + adr + e * st;
        if (m == 8) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; u++) v[u] = src[u * st];
#pragma unroll
          for (int u = 0; u < 8; u++) dst[u * L.per_obs] = v[u];
        } else {
          for (int u = 0; u < m; u++) dst[u * L.per_obs] = src[u * st];
        }
      }
    }
    __syncthreads();
    if (prof) tk[4] = clock64();
    // dynamics / reward / truncation / done in the reference's order: thread = env, on its rows in shared memory
    for (int e = tid; e < n; e += nthreads) {
      if (A <= 2)
        run_plugins<2>(dm, s_ctrl + e * dm.ctrl_stride, B.obs + (e0 + e) * L.per_obs, s_rew + e * A, s_term + e * (A + 1), s_trunc + e * (A + 1),
                       (int)e0 + e, 0, false, s_probe + e * dm.np1 * 4, s_si + e * A * dm.store_i32, s_sf + e * A * dm.store_f32,
                       s_act + e * A * dm.act_stride, s_ts + e);
      else
      run_plugins<MJB_MAX_AGENTS>(dm, s_ctrl + e * dm.ctrl_stride, B.obs + (e0 + e) * L.per_obs, s_rew + e * A, s_term + e * (A + 1), s_trunc + e * (A + 1),
                  (int)e0 + e, 0, false, s_probe + e * dm.np1 * 4, s_si + e * A * dm.store_i32, s_sf + e * A * dm.store_f32,
                  s_act + e * A * dm.act_stride, s_ts + e);
    }
    __syncthreads();
    if (prof) tk[5] = clock64();
    // results: whole row blocks, 128-bit stores
    if (dm.free_joint) tile_store(B.qvel + e0 * dm.qvel_stride, s_qvel, n * dm.qvel_stride, tid, nthreads);
    else tile_store(B.ctrl + e0 * dm.ctrl_stride, s_ctrl, n * dm.ctrl_stride, tid, nthreads);
    tile_store(reinterpret_cast<float*>(B.store_i + e0 * A * dm.store_i32), reinterpret_cast<const float*>(s_si), n * A * dm.store_i32, tid, nthreads);
    tile_store(B.store_f + e0 * A * dm.store_f32, s_sf, n * A * dm.store_f32, tid, nthreads);
    for (int i = tid; i < n * A; i += nthreads) B.reward[e0 * A + i] = s_rew[i];
    for (int i = tid; i < n * (A + 1); i += nthreads) { B.term[e0 * (A + 1) + i] = s_term[i]; B.trunc[e0 * (A + 1) + i] = s_trunc[i]; }
    for (int e = tid; e < n; e += nthreads) B.timestep[e0 + e] = s_ts[e];
    __syncthreads();
    if (prof) { tk[6] = clock64(); lite_report(tk); printf("  start: init+sync %lld | fence %lld | switch+expect %lld | bulk issue %lld\n", tx[0] - tk[0], tx[1] - tx[0], tx[2] - tx[1], tx[3] - tx[2]); }
  }
}

}  // namespace mjb
