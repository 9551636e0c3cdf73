// simt_emu.h — minimal single-warp SIMT emulator (test tooling, host only).
// 32 lanes are 32 ucontext fibres scheduled round-robin: a lane runs until it reaches a
// barrier / shuffle / ballot, then yields to the next lane; when control comes back every lane has
// arrived.  Works for kernels whose sync points sit in warp-uniform control flow (the same
// requirement CUDA has).  `reverse` runs the lanes in descending order, which exposes missing
// barriers in the opposite producer/consumer direction.
#pragma once
#include <stdint.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <stdexcept>
#include <vector>

namespace simt {

struct Warp {
  static const int W = 32;
  ucontext_t main_ctx, ctx[W];
  std::vector<char> stacks;
  int order[W];
  int cur = 0;        // position in `order`
  bool done[W];
  int ndone = 0;
  uint32_t exch[W];
  std::function<void()> body;
};

inline Warp*& current() {
  static thread_local Warp* w = nullptr;
  return w;
}
inline int lane() { return current()->order[current()->cur]; }

inline void yield_next() {
  Warp* w = current();
  int from = w->cur;
  for (int step = 1; step <= Warp::W; step++) {
    int nxt = (from + step) % Warp::W;
    if (!w->done[w->order[nxt]]) {
      if (nxt == from) return;
      w->cur = nxt;
      swapcontext(&w->ctx[w->order[from]], &w->ctx[w->order[nxt]]);
      return;
    }
  }
}
inline void barrier() { yield_next(); }

template <typename T>
inline T shfl(T v, int src) {
  static_assert(sizeof(T) == 4, "4-byte shuffles only");
  Warp* w = current();
  memcpy(&w->exch[lane()], &v, 4);
  barrier();
  T r;
  memcpy(&r, &w->exch[src & 31], 4);
  barrier();
  return r;
}
template <typename T>
inline T shfl_xor(T v, int m) { return shfl(v, lane() ^ m); }
inline uint32_t ballot(bool p) {
  Warp* w = current();
  w->exch[lane()] = p ? 1u : 0u;
  barrier();
  uint32_t r = 0;
  for (int i = 0; i < 32; i++) r |= (w->exch[i] & 1u) << i;
  barrier();
  return r;
}

inline void trampoline() {
  Warp* w = current();
  w->body();
  int me = w->order[w->cur];
  w->done[me] = true;
  w->ndone++;
  if (w->ndone == Warp::W) {
    setcontext(&w->main_ctx);
  } else {
    int from = w->cur;
    for (int step = 1; step <= Warp::W; step++) {
      int nxt = (from + step) % Warp::W;
      if (!w->done[w->order[nxt]]) {
        w->cur = nxt;
        setcontext(&w->ctx[w->order[nxt]]);
      }
    }
  }
}

// run `body` once per lane as a warp
inline void run_warp(const std::function<void()>& body, bool reverse = false, size_t stack_bytes = 256 * 1024) {
  Warp w;
  w.body = body;
  w.stacks.resize(stack_bytes * Warp::W);
  for (int i = 0; i < Warp::W; i++) {
    w.order[i] = reverse ? Warp::W - 1 - i : i;
    w.done[i] = false;
  }
  Warp* prev = current();
  current() = &w;
  for (int i = 0; i < Warp::W; i++) {
    getcontext(&w.ctx[i]);
    w.ctx[i].uc_stack.ss_sp = w.stacks.data() + (size_t)i * stack_bytes;
    w.ctx[i].uc_stack.ss_size = stack_bytes;
    w.ctx[i].uc_link = nullptr;
    makecontext(&w.ctx[i], (void (*)())trampoline, 0);
  }
  w.cur = 0;
  volatile bool started = false;
  getcontext(&w.main_ctx);
  if (!started) {
    started = true;
    setcontext(&w.ctx[w.order[0]]);
  }
  current() = prev;
}

}  // namespace simt
