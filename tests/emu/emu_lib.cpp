// emu_lib.cpp — host build of the device step code on the SIMT emulator (test tooling only).
// Lets the exact kernel source run (slowly) in a container without a GPU, so that the warp-level
// algorithms can be diffed against the fp64 oracle before spending GPU time.
#define MJB_HOST_EMU 1
#include <cstring>
#include <string>
#include <vector>

#include "../../mujoco_rl_environment_wrapper_b200/csrc/dev_model_build.h"
#include "../../mujoco_rl_environment_wrapper_b200/csrc/env_kernel.cuh"

static std::string g_err;

extern "C" {

const char* emu_last_error() { return g_err.c_str(); }

int emu_layout(const void* blob, const mjb_env_spec* spec, int num_envs, mjb_layout* out) {
  try {
    mjb::ModelView mv(blob);
    mjb::DevImage img;
    mjb::build_dev_model(mv, *spec, img);
    mjb::fill_layout(img.dm, num_envs, *out);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// run `mode` for envs [0, num_envs) over HOST buffers; iters_out (optional) gets Newton iterations of the
// last forward of env 0
int emu_run(const void* blob, const mjb_env_spec* spec, int num_envs, const mjb_buffers* B, int mode, int skip_frames,
            const uint8_t* mask, int reverse) {
  try {
    mjb::ModelView mv(blob);
    mjb::DevImage img;
    mjb::build_dev_model(mv, *spec, img);
    std::vector<float> scratch(img.dm.env_words + 64, 0.f), probe(4 * img.dm.nprobe + 4, 0.f);
    for (int env = 0; env < num_envs; env++) {
      simt::run_warp([&]() {
        mjb::Ctx c{&img.dm, img.words.data(), scratch.data(), simt::lane(), probe.data(), 0, 0};
        mjb::run_env<true>(c, *B, env, mode, skip_frames, mask);
      }, reverse != 0);
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

}  // extern "C"
