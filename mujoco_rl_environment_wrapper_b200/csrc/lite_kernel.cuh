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
//   * results leave as 128-bit coalesced stores of whole row blocks; only what the step changes is written
//     (qvel or ctrl, obs, reward, flags, data_store rows, step counter);
//   * no model image staging: the two index tables are copied into shared memory once per CTA, under the first load.
// Several small CTAs per SM keep tens of KB of loads in flight per SM.
#pragma once
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
  int words;
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
  L.o_obs = take(L.per_obs); L.o_rew = take(A);
  L.o_term = o; o += ((tile * (A + 1) + 15) / 16) * 4;
  L.o_trunc = o; o += ((tile * (A + 1) + 15) / 16) * 4;
  L.o_gather = o; o += ((L.per_obs + 3) / 4) * 4;
  L.o_actidx = o; o += ((A * dm.n_phys_act + 3) / 4) * 4;
  L.words = o;
  return L;
}

// 128-bit copy of `words` floats (multiple of 4; both sides 16-byte aligned) from shared to global memory
__device__ __forceinline__ void tile_store(float* __restrict__ dst, const float* src, int words, int tid, int nthreads) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = tid; i < (words >> 2); i += nthreads) d4[i] = s4[i];
}

__global__ void __launch_bounds__(LITE_THREADS, 8) k_lite(const __grid_constant__ DevModel dm, const uint32_t* __restrict__ image, const mjb_buffers B,
                                              const LiteLayout L, int active, int env_base) {
  extern __shared__ __align__(128) uint32_t lite_smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(lite_smem);
  float* sm = reinterpret_cast<float*>(lite_smem + 4);
  float *s_qpos = sm + L.o_qpos, *s_qvel = sm + L.o_qvel, *s_ctrl = sm + L.o_ctrl, *s_sens = sm + L.o_sens, *s_act = sm + L.o_act;
  float *s_probe = sm + L.o_probe, *s_sf = sm + L.o_sf, *s_obs = sm + L.o_obs, *s_rew = sm + L.o_rew;
  int *s_si = reinterpret_cast<int*>(sm + L.o_si), *s_ts = reinterpret_cast<int*>(sm + L.o_ts);
  uint8_t *s_term = reinterpret_cast<uint8_t*>(sm + L.o_term), *s_trunc = reinterpret_cast<uint8_t*>(sm + L.o_trunc);
  const int tid = threadIdx.x, nthreads = blockDim.x, T = L.tile, A = dm.a1;
  int *s_gather = reinterpret_cast<int*>(sm + L.o_gather), *s_actidx = reinterpret_cast<int*>(sm + L.o_actidx);
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  // the index tables, once per CTA: entry r of an env's obs rows -> (kind << 24 | address) or -1 for row padding
  {
    const int* act_index = reinterpret_cast<const int*>(image + dm.off[IF_act_index]);
    const int* obs_index = reinterpret_cast<const int*>(image + dm.off[IF_obs_index]);
    for (int r = tid; r < L.per_obs; r += nthreads) {
      const int a = r / dm.obs_stride, k = r - a * dm.obs_stride;
      s_gather[r] = k < dm.obs_adr[a + 1] - dm.obs_adr[a] ? __ldg(obs_index + dm.obs_adr[a] + k) : -1;
    }
    for (int k = tid; k < A * dm.n_phys_act; k += nthreads) s_actidx[k] = __ldg(act_index + k);
  }
  __syncthreads();
  const int ntiles = (active + T - 1) / T;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, phase ^= 1u) {
    const int n = min(T, active - tile * T);          // envs of this tile
    const size_t e0 = (size_t)env_base + (size_t)tile * T;
    if (tid == 0) {
      // the previous tile's shared-memory reads / writes (generic proxy) are ordered before the bulk copies (async proxy)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const uint32_t row = 4u * (uint32_t)n;
      uint32_t bytes = row * (dm.qpos_stride + dm.qvel_stride + A * dm.act_stride + A * dm.store_i32 + A * dm.store_f32);
      if (L.use_ctrl) bytes += row * dm.ctrl_stride;
      if (L.use_sens) bytes += row * dm.sensor_stride;
      if (L.use_probe) bytes += row * dm.np1 * 4;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(s_qpos, B.qpos + e0 * dm.qpos_stride, row * dm.qpos_stride, bar);
      bulk_g2s(s_qvel, B.qvel + e0 * dm.qvel_stride, row * dm.qvel_stride, bar);
      if (L.use_ctrl) bulk_g2s(s_ctrl, B.ctrl + e0 * dm.ctrl_stride, row * dm.ctrl_stride, bar);
      if (L.use_sens) bulk_g2s(s_sens, B.sensordata + e0 * dm.sensor_stride, row * dm.sensor_stride, bar);
      bulk_g2s(s_act, B.actions + e0 * A * dm.act_stride, row * A * dm.act_stride, bar);
      if (L.use_probe) bulk_g2s(s_probe, B.probe + e0 * dm.np1 * 4, row * dm.np1 * 4, bar);
      bulk_g2s(s_si, B.store_i + e0 * A * dm.store_i32, row * A * dm.store_i32, bar);
      bulk_g2s(s_sf, B.store_f + e0 * A * dm.store_f32, row * A * dm.store_f32, bar);
    }
    for (int e = tid; e < n; e += nthreads) s_ts[e] = B.timestep[e0 + e];   // 4 bytes per env: not worth a bulk copy (alignment of odd ranges)
    mbar_wait(bar, phase);
    // apply_action (mujoco_parent.py:316-332): overwrite qvel (freeJoint) or ctrl; thread = env
    for (int e = tid; e < n; e += nthreads) {
      float* dst = dm.free_joint ? s_qvel + e * dm.qvel_stride : s_ctrl + e * dm.ctrl_stride;
      const float* src = s_act + e * A * dm.act_stride;
      for (int a = 0, k = 0; a < A; a++)
        for (int j = 0; j < dm.n_phys_act; j++, k++) dst[s_actidx[k]] = src[a * dm.act_stride + j];
    }
    __syncthreads();
    // get_observations (mujoco_parent.py:380-392): warp = env, lane = word of its obs rows; padding is written as zero
    for (int e = warp; e < n; e += nwarps) {
      const float *q = s_qpos + e * dm.qpos_stride, *v = s_qvel + e * dm.qvel_stride, *sd = s_sens + e * dm.sensor_stride;
      for (int r = lane; r < L.per_obs; r += 32) {
        const int ent = s_gather[r], kind = ent >> 24, adr = ent & 0xffffff;
        s_obs[e * L.per_obs + r] = ent < 0 ? 0.f : (kind == 0 ? sd[adr] : (kind == 1 ? q[adr] : v[adr]));
      }
    }
    __syncthreads();
    // dynamics / reward / truncation / done in the reference's order: thread = env, on its rows in shared memory
    for (int e = tid; e < n; e += nthreads) {
      run_plugins(dm, s_ctrl + e * dm.ctrl_stride, s_obs + e * L.per_obs, s_rew + e * A, s_term + e * (A + 1), s_trunc + e * (A + 1),
                  (int)e0 + e, 0, false, s_probe + e * dm.np1 * 4, s_si + e * A * dm.store_i32, s_sf + e * A * dm.store_f32,
                  s_act + e * A * dm.act_stride, s_ts + e);
    }
    __syncthreads();
    // results: whole row blocks, 128-bit stores
    tile_store(B.obs + e0 * L.per_obs, s_obs, n * L.per_obs, tid, nthreads);
    if (dm.free_joint) tile_store(B.qvel + e0 * dm.qvel_stride, s_qvel, n * dm.qvel_stride, tid, nthreads);
    else tile_store(B.ctrl + e0 * dm.ctrl_stride, s_ctrl, n * dm.ctrl_stride, tid, nthreads);
    tile_store(reinterpret_cast<float*>(B.store_i + e0 * A * dm.store_i32), reinterpret_cast<const float*>(s_si), n * A * dm.store_i32, tid, nthreads);
    tile_store(B.store_f + e0 * A * dm.store_f32, s_sf, n * A * dm.store_f32, tid, nthreads);
    for (int i = tid; i < n * A; i += nthreads) B.reward[e0 * A + i] = s_rew[i];
    for (int i = tid; i < n * (A + 1); i += nthreads) { B.term[e0 * (A + 1) + i] = s_term[i]; B.trunc[e0 * (A + 1) + i] = s_trunc[i]; }
    for (int e = tid; e < n; e += nthreads) B.timestep[e0 + e] = s_ts[e];
    __syncthreads();
  }
}

}  // namespace mjb
