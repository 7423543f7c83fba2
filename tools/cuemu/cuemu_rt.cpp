// cuemu_rt.cpp — fiber scheduler of the CPU emulation (see cuemu.h; tests only).
#include <time.h>

#include <random>
#include <vector>

#include "cuemu.h"

namespace cuemu {
namespace {

constexpr size_t STACK_BYTES = 512 * 1024;

// Minimal x86-64 System V context switch (callee-saved registers + stack pointer); swapcontext() costs a
// sigprocmask system call per switch, which dominated the run time of the emulated kernels.
#if !defined(__x86_64__)
#error "cuemu: the fiber switch is written for x86-64"
#endif
extern "C" void cuemu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl cuemu_switch
.type cuemu_switch,@function
cuemu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size cuemu_switch,.-cuemu_switch
)");

struct Fiber {
  void *sp = nullptr;
  char *stack = nullptr;
  bool done = false;
  ThreadCtx tc{};
};

struct Warp {
  int arrived = 0;
  unsigned gen = 0;
  uint64_t xbuf[2][32];
  unsigned ballot_bits[2] = {0, 0};
};

struct BlockState {
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  std::vector<int> order;
  void *sched_sp = nullptr;
  int current = -1;
  int nthreads = 0;
  int alive = 0;
  int bar_arrived = 0;
  unsigned bar_gen = 0;
  uint3 bidx{};
  dim3 bdim, gdim;
  void *smem = nullptr;
  const std::function<void()> *body = nullptr;
};

BlockState *B = nullptr;
long g_launches = 0;
std::vector<char *> g_stack_pool;

void yield() { cuemu_switch(&B->fibers[B->current].sp, B->sched_sp); }

void fiber_main() {
  (*B->body)();
  Fiber &f = B->fibers[B->current];
  f.done = true;
  B->alive--;
  // an exited thread no longer takes part in barriers: release one that only waited for it
  if (B->alive > 0 && B->bar_arrived >= B->alive) { B->bar_arrived = 0; B->bar_gen++; }
  cuemu_switch(&f.sp, B->sched_sp);
  abort();  // a finished fiber is never resumed
}

int warp_alive_in_mask(int w, unsigned mask) {
  int n = 0;
  for (int l = 0; l < 32; l++) {
    const int t = 32 * w + l;
    if (t < B->nthreads && ((mask >> l) & 1u) && !B->fibers[t].done) n++;
  }
  return n;
}

// one barrier among the live lanes of `mask` in the calling fiber's warp; returns the generation it closed
unsigned warp_barrier(unsigned mask) {
  const int w = B->current >> 5;
  Warp &W = B->warps[w];
  const unsigned my = W.gen;
  const int need = warp_alive_in_mask(w, mask);
  if (++W.arrived >= need) {
    W.arrived = 0;
    W.gen++;
  } else {
    while (W.gen == my) yield();
  }
  return my;
}

}  // namespace

const ThreadCtx &cur() { return B->fibers[B->current].tc; }
const uint3 &block_idx() { return B->bidx; }
const dim3 &block_dim() { return B->bdim; }
const dim3 &grid_dim() { return B->gdim; }
void *dyn_smem() { return B->smem; }
long launches() { return g_launches; }

void sync_block() {
  const unsigned my = B->bar_gen;
  if (++B->bar_arrived >= B->alive) {
    B->bar_arrived = 0;
    B->bar_gen++;
  } else {
    while (B->bar_gen == my) yield();
  }
}

void sync_warp(unsigned mask) { warp_barrier(mask); }

uint64_t shfl(unsigned mask, uint64_t v, int src) {
  const int w = B->current >> 5, lane = B->current & 31;
  Warp &W = B->warps[w];
  const int p = W.gen & 1;
  W.xbuf[p][lane] = v;
  warp_barrier(mask);
  // a source lane outside the mask / the block returns the caller's own value (undefined on hardware)
  if (src < 0 || src > 31 || !((mask >> src) & 1u) || 32 * w + src >= B->nthreads) return v;
  return W.xbuf[p][src];
}

unsigned ballot(unsigned mask, bool pred) {
  const int w = B->current >> 5, lane = B->current & 31;
  Warp &W = B->warps[w];
  const int p = W.gen & 1;
  if (W.arrived == 0) W.ballot_bits[p] = 0;
  if (pred) W.ballot_bits[p] |= 1u << lane;
  warp_barrier(mask);
  return W.ballot_bits[p] & mask;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
  g_launches++;
  if (B) { fprintf(stderr, "cuemu: nested launch\n"); abort(); }
  BlockState bs;
  B = &bs;
  const int nt = (int)(block.x * block.y * block.z);
  bs.nthreads = nt;
  bs.bdim = block;
  bs.gdim = grid;
  bs.body = &body;
  bs.fibers.resize(nt);
  bs.warps.resize((nt + 31) / 32);
  bs.order.resize(nt);
  while ((int)g_stack_pool.size() < nt) g_stack_pool.push_back((char *)malloc(STACK_BYTES));
  void *smem = nullptr;
  if (posix_memalign(&smem, 1024, smem_bytes + 1024) != 0) abort();
  bs.smem = smem;
  const char *shuf = getenv("CUEMU_SHUFFLE");
  std::mt19937 rng(shuf ? (unsigned)atoi(shuf) : 0u);
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        bs.bidx = uint3{bx, by, bz};
        bs.alive = nt;
        bs.bar_arrived = 0;
        memset(smem, 0xFF, smem_bytes);  // shared memory is not zero-initialised: NaN pattern
        for (auto &w : bs.warps) { w.arrived = 0; }
        for (int t = 0; t < nt; t++) {
          Fiber &f = bs.fibers[t];
          f.done = false;
          f.stack = g_stack_pool[t];
          f.tc.linear = t;
          f.tc.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
          // initial frame: six callee-saved registers, then the entry point as the return address of
          // cuemu_switch; rsp = 8 mod 16 at fiber_main's entry as after a call
          uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
          void **fr = (void **)(top - 64);
          for (int i = 0; i < 6; i++) fr[i] = nullptr;
          fr[6] = (void *)fiber_main;
          fr[7] = nullptr;
          f.sp = fr;
          bs.order[t] = t;
        }
        if (shuf) std::shuffle(bs.order.begin(), bs.order.end(), rng);
        long idle_rounds = 0;
        while (bs.alive > 0) {
          const int before = bs.alive;
          const unsigned g0 = bs.bar_gen;
          unsigned wg = 0;
          for (auto &w : bs.warps) wg += w.gen;
          for (int q = 0; q < nt; q++) {
            const int t = bs.order[q];
            if (bs.fibers[t].done) continue;
            bs.current = t;
            cuemu_switch(&bs.sched_sp, bs.fibers[t].sp);
          }
          unsigned wg1 = 0;
          for (auto &w : bs.warps) wg1 += w.gen;
          if (bs.alive == before && bs.bar_gen == g0 && wg1 == wg) {
            if (++idle_rounds > 4) {
              fprintf(stderr, "cuemu: deadlock in block (%u,%u,%u): %d threads alive, %d at the block barrier "
                      "(divergent __syncthreads or a collective some lanes never reach)\n", bx, by, bz, bs.alive, bs.bar_arrived);
              abort();
            }
          } else {
            idle_rounds = 0;
          }
        }
      }
  free(smem);
  B = nullptr;
}

}  // namespace cuemu

double cuemu_now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
