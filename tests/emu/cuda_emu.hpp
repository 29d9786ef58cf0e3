// cuda_emu.hpp -- TEST INFRASTRUCTURE ONLY.
//
// A tiny single-OS-thread SIMT emulator: every CUDA thread of a block is a ucontext fiber, so the
// kernel source in grsr_b200/csrc/grsr_core.cuh can be compiled with g++ and stepped through on a
// machine without a GPU.  It exists to debug kernel LOGIC (indexing, barrier placement, warp
// collectives) under ASAN/valgrind in the build container; it is never linked into the product
// library and nothing in grsr_b200/ can reach it.  Blocks run one after another; threads of a block
// are scheduled round-robin and only switch at barriers / warp collectives, so data races are NOT
// modelled (compute-sanitizer on the GPU box covers those).
//
// Supported: threadIdx/blockIdx/blockDim/gridDim (.x only), __syncthreads[_or|_count],
// __syncwarp, __shfl[_down|_up|_xor]_sync, __ballot_sync, __any/__all_sync, __match_any_sync,
// __reduce_{min,max,add,or}_sync, atomic{Add,Min,Max,Or,And,CAS,Exch}, __popc/__clz/__ffs,
// dynamic shared memory via emu::dyn_smem().
#pragma once
#include <ucontext.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

namespace emu {

struct Dim { unsigned x = 1, y = 1, z = 1; };

struct Fiber {
  ucontext_t ctx;
  std::vector<unsigned char> stack;
  bool done = false;
  // block barrier
  unsigned wait_gen = 0; bool at_block_bar = false;
  // warp barrier
  unsigned wwait_gen = 0; bool at_warp_bar = false;
  // cluster barrier
  unsigned cwait_gen = 0; bool at_cluster_bar = false;
};

struct Cluster {
  int nblocks = 1;
  int nthreads_total = 0;
  unsigned gen = 0; int arrived = 0;
};

struct Block {
  int nthreads = 0;
  std::vector<Fiber> fibers;
  ucontext_t sched_own;
  ucontext_t *sched = &sched_own;   // blocks of one cluster share a scheduler context
  Cluster *cluster = nullptr;
  int cluster_rank = 0;
  unsigned block_index = 0;
  int cur = -1;
  int live = 0;
  // block barrier state
  unsigned gen = 0; int arrived = 0; int or_acc = 0; int cnt_acc = 0; int or_res = 0; int cnt_res = 0;
  // warp state (up to 32 warps)
  unsigned wgen[32]; int warrived[32]; unsigned wexpect[32];
  uint64_t xchg[32][32];
  int xpred[32][32];
  std::vector<unsigned char> smem;
  const std::function<void()> *body = nullptr;
};

inline Block *&cur_block() { static Block *b = nullptr; return b; }
inline Dim &tidx() { static Dim d; return d; }
inline Dim &bidx() { static Dim d; return d; }
inline Dim &bdim() { static Dim d; return d; }
inline Dim &gdim() { static Dim d; return d; }

inline void yield_to_sched() {
  Block *b = cur_block();
  swapcontext(&b->fibers[b->cur].ctx, b->sched);
}

inline void die(const char *msg) { fprintf(stderr, "cuda_emu: %s\n", msg); abort(); }

inline void block_barrier(int pred, int *or_out, int *cnt_out) {
  Block *b = cur_block();
  Fiber &f = b->fibers[b->cur];
  if (b->live != b->nthreads) die("__syncthreads reached after a thread of the block exited");
  unsigned my_gen = b->gen;
  b->or_acc |= (pred != 0);
  b->cnt_acc += (pred != 0);
  if (++b->arrived == b->live) {
    b->or_res = b->or_acc; b->cnt_res = b->cnt_acc;
    b->or_acc = 0; b->cnt_acc = 0; b->arrived = 0; b->gen++;
  } else {
    f.at_block_bar = true; f.wait_gen = my_gen;
    while (b->gen == my_gen) yield_to_sched();
    f.at_block_bar = false;
  }
  if (or_out) *or_out = b->or_res;
  if (cnt_out) *cnt_out = b->cnt_res;
}

inline void warp_barrier(unsigned mask) {
  Block *b = cur_block();
  int w = b->cur >> 5;
  Fiber &f = b->fibers[b->cur];
  unsigned my_gen = b->wgen[w];
  if (b->warrived[w] == 0) b->wexpect[w] = mask;
  else if (b->wexpect[w] != mask) die("warp collective called with inconsistent masks");
  if (!((mask >> (b->cur & 31)) & 1)) die("warp collective called by a lane outside its mask");
  if (++b->warrived[w] == __builtin_popcount(mask)) {
    b->warrived[w] = 0; b->wgen[w]++;
  } else {
    f.at_warp_bar = true; f.wwait_gen = my_gen;
    while (b->wgen[w] == my_gen) yield_to_sched();
    f.at_warp_bar = false;
  }
}

inline void cluster_barrier() {
  Block *b = cur_block();
  Cluster *c = b->cluster;
  if (!c) die("cluster barrier outside a cluster launch");
  Fiber &f = b->fibers[b->cur];
  unsigned my_gen = c->gen;
  if (++c->arrived == c->nthreads_total) {
    c->arrived = 0; c->gen++;
  } else {
    f.at_cluster_bar = true; f.cwait_gen = my_gen;
    while (c->gen == my_gen) yield_to_sched();
    f.at_cluster_bar = false;
  }
}
inline unsigned cluster_rank() { return (unsigned)cur_block()->cluster_rank; }
inline unsigned cluster_size() { Block *b = cur_block(); return b->cluster ? (unsigned)b->cluster->nblocks : 1u; }

inline unsigned char *dyn_smem() { return cur_block()->smem.data(); }

inline void fiber_entry() {
  Block *b = cur_block();
  (*b->body)();
  b->fibers[b->cur].done = true;
  b->live--;
  yield_to_sched();
  die("resumed a finished fiber");
}

// Run `body` as a grid of `grid` blocks of `block` threads with `smem_bytes` of dynamic smem.
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body) {
  if (block == 0 || block > 1024) die("bad block size");
  gdim().x = grid; bdim().x = block;
  static const size_t kStack = 64 * 1024;
  for (unsigned bi = 0; bi < grid; bi++) {
    Block blk;
    blk.nthreads = (int)block; blk.live = (int)block; blk.body = &body;
    blk.smem.assign(smem_bytes + 16, 0xCD);
    blk.fibers.resize(block);
    memset(blk.wgen, 0, sizeof blk.wgen); memset(blk.warrived, 0, sizeof blk.warrived);
    cur_block() = &blk;
    bidx().x = bi;
    for (unsigned t = 0; t < block; t++) {
      Fiber &f = blk.fibers[t];
      f.stack.resize(kStack);
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack.data();
      f.ctx.uc_stack.ss_size = kStack;
      f.ctx.uc_link = nullptr;
      makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    int remaining = (int)block;
    while (remaining > 0) {
      int progressed = 0;
      for (unsigned t = 0; t < block; t++) {
        Fiber &f = blk.fibers[t];
        if (f.done) continue;
        if (f.at_block_bar && blk.gen == f.wait_gen) continue;
        if (f.at_warp_bar && blk.wgen[t >> 5] == f.wwait_gen) continue;
        blk.cur = (int)t; tidx().x = t;
        swapcontext(blk.sched, &f.ctx);
        progressed = 1;
        if (f.done) remaining--;
      }
      if (!progressed) die("deadlock: every live thread waits at a barrier (divergent barrier?)");
    }
    cur_block() = nullptr;
  }
}

// Run `body` as `nclusters` clusters of `cs` blocks of `block` threads; the blocks of a cluster run
// concurrently (their fibers share one scheduler) so that cluster barriers can be emulated.
inline void launch_clusters(unsigned nclusters, unsigned cs, unsigned block, size_t smem_bytes,
                            const std::function<void()> &body) {
  if (block == 0 || block > 1024 || cs == 0 || cs > 16) die("bad cluster launch");
  gdim().x = nclusters * cs; bdim().x = block;
  static const size_t kStack = 64 * 1024;
  for (unsigned ci = 0; ci < nclusters; ci++) {
    Cluster cl;
    cl.nblocks = (int)cs; cl.nthreads_total = (int)(cs * block);
    ucontext_t sched;
    std::vector<Block> blks(cs);
    for (unsigned r = 0; r < cs; r++) {
      Block &blk = blks[r];
      blk.nthreads = (int)block; blk.live = (int)block; blk.body = &body;
      blk.smem.assign(smem_bytes + 16, 0xCD);
      blk.fibers.resize(block);
      memset(blk.wgen, 0, sizeof blk.wgen); memset(blk.warrived, 0, sizeof blk.warrived);
      blk.sched = &sched; blk.cluster = &cl; blk.cluster_rank = (int)r; blk.block_index = ci * cs + r;
      for (unsigned t = 0; t < block; t++) {
        Fiber &f = blk.fibers[t];
        f.stack.resize(kStack);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.data();
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
      }
    }
    int remaining = (int)(cs * block);
    while (remaining > 0) {
      int progressed = 0;
      for (unsigned r = 0; r < cs; r++) {
        Block &blk = blks[r];
        for (unsigned t = 0; t < block; t++) {
          Fiber &f = blk.fibers[t];
          if (f.done) continue;
          if (f.at_block_bar && blk.gen == f.wait_gen) continue;
          if (f.at_warp_bar && blk.wgen[t >> 5] == f.wwait_gen) continue;
          if (f.at_cluster_bar && cl.gen == f.cwait_gen) continue;
          cur_block() = &blk;
          blk.cur = (int)t; tidx().x = t; bidx().x = blk.block_index;
          swapcontext(&sched, &f.ctx);
          progressed = 1;
          if (f.done) remaining--;
        }
      }
      if (!progressed) die("deadlock in cluster launch (divergent barrier?)");
    }
    cur_block() = nullptr;
  }
}

} // namespace emu

#define threadIdx (emu::tidx())
#define blockIdx  (emu::bidx())
#define blockDim  (emu::bdim())
#define gridDim   (emu::gdim())

inline void __syncthreads() { emu::block_barrier(0, nullptr, nullptr); }
inline int __syncthreads_or(int p) { int r; emu::block_barrier(p, &r, nullptr); return r; }
inline int __syncthreads_count(int p) { int r; emu::block_barrier(p, nullptr, &r); return r; }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_barrier(mask); }
inline void __threadfence() {}
inline void __threadfence_block() {}

namespace emu {
template <class T> inline T shfl_generic(unsigned mask, T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shfl type too wide");
  Block *b = cur_block();
  int w = b->cur >> 5, lane = b->cur & 31;
  uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
  b->xchg[w][lane] = raw;
  warp_barrier(mask);
  T out = v;
  if (src_lane >= 0 && src_lane < 32 && ((mask >> src_lane) & 1)) {
    uint64_t r = b->xchg[w][src_lane]; memcpy(&out, &r, sizeof(T));
  }
  warp_barrier(mask);
  return out;
}
inline int lane_id() { return cur_block()->cur & 31; }
} // namespace emu

template <class T> inline T __shfl_sync(unsigned m, T v, int src, int width = 32) {
  int lane = emu::lane_id();
  int base = lane & ~(width - 1);
  return emu::shfl_generic(m, v, base + (src & (width - 1)));
}
template <class T> inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) {
  int lane = emu::lane_id();
  int src = lane + (int)d;
  if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
  return emu::shfl_generic(m, v, src);
}
template <class T> inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) {
  int lane = emu::lane_id();
  int src = lane - (int)d;
  if (src < (lane & ~(width - 1))) src = lane;
  return emu::shfl_generic(m, v, src);
}
template <class T> inline T __shfl_xor_sync(unsigned m, T v, int x, int width = 32) {
  (void)width;
  return emu::shfl_generic(m, v, emu::lane_id() ^ x);
}
inline unsigned __ballot_sync(unsigned m, int pred) {
  emu::Block *b = emu::cur_block();
  int w = b->cur >> 5, lane = b->cur & 31;
  b->xpred[w][lane] = pred != 0;
  emu::warp_barrier(m);
  unsigned r = 0;
  for (int l = 0; l < 32; l++) if (((m >> l) & 1) && b->xpred[w][l]) r |= 1u << l;
  emu::warp_barrier(m);
  return r;
}
inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, p) == m; }
template <class T> inline unsigned __match_any_sync(unsigned m, T v) {
  emu::Block *b = emu::cur_block();
  int w = b->cur >> 5, lane = b->cur & 31;
  uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
  b->xchg[w][lane] = raw;
  emu::warp_barrier(m);
  unsigned r = 0;
  for (int l = 0; l < 32; l++) if (((m >> l) & 1) && b->xchg[w][l] == raw) r |= 1u << l;
  emu::warp_barrier(m);
  return r;
}
#define EMU_REDUCE(name, T, expr)                                              \
  inline T name(unsigned m, T v) {                                             \
    emu::Block *b = emu::cur_block();                                          \
    int w = b->cur >> 5, lane = b->cur & 31;                                   \
    b->xchg[w][lane] = (uint64_t)(int64_t)v;                                   \
    emu::warp_barrier(m);                                                      \
    bool first = true; T acc = 0;                                              \
    for (int l = 0; l < 32; l++) if ((m >> l) & 1) {                           \
      T x = (T)(int64_t)b->xchg[w][l];                                         \
      if (first) { acc = x; first = false; } else { acc = (expr); }            \
    }                                                                          \
    emu::warp_barrier(m);                                                      \
    return acc;                                                                \
  }
EMU_REDUCE(__reduce_min_sync, unsigned, (x < acc ? x : acc))
EMU_REDUCE(__reduce_max_sync, unsigned, (x > acc ? x : acc))
EMU_REDUCE(__reduce_add_sync, unsigned, (acc + x))
EMU_REDUCE(__reduce_or_sync, unsigned, (acc | x))
EMU_REDUCE(__reduce_min_sync, int, (x < acc ? x : acc))
EMU_REDUCE(__reduce_max_sync, int, (x > acc ? x : acc))
EMU_REDUCE(__reduce_add_sync, int, (acc + x))

template <class T> inline T atomicAdd(T *a, T v) { T o = *a; *a = o + v; return o; }
template <class T> inline T atomicMin(T *a, T v) { T o = *a; if (v < o) *a = v; return o; }
template <class T> inline T atomicMax(T *a, T v) { T o = *a; if (v > o) *a = v; return o; }
template <class T> inline T atomicOr(T *a, T v) { T o = *a; *a = o | v; return o; }
template <class T> inline T atomicAnd(T *a, T v) { T o = *a; *a = o & v; return o; }
template <class T> inline T atomicExch(T *a, T v) { T o = *a; *a = v; return o; }
template <class T> inline T atomicCAS(T *a, T cmp, T v) { T o = *a; if (o == cmp) *a = v; return o; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
template <class T> inline T __ldg(const T *p) { return *p; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
