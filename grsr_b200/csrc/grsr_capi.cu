// grsr_capi.cu -- the C ABI declared in include/grsr.h: argument checks, device buffers, copies,
// launches and the multi-GPU pair-list sharder.  All arithmetic of the path happens in the kernels
// of grsr_core.cuh / grsr_scenario.cuh; there is no host implementation of it here.
#include "../../include/grsr.h"
#include "grsr_core.cuh"
#include "grsr_scenario.cuh"
#include "grsr_stepwise.cuh"
#include "grsr_walk.cuh"
#include "grsr_unsigned.cuh"
#include "grsr_unsigned_host.hpp"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err = "";

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(GRSR_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));        \
  } while (0)

struct DevInfo { int sms = 0; int smem_optin = 0; bool ok = false; };

int dev_info(int dev, DevInfo &di) {
  static std::mutex mu;
  static std::vector<DevInfo> cache;
  std::lock_guard<std::mutex> lk(mu);
  if ((int)cache.size() <= dev) cache.resize(dev + 1);
  if (!cache[dev].ok) {
    CK(cudaDeviceGetAttribute(&cache[dev].sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&cache[dev].smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cache[dev].ok = true;
  }
  di = cache[dev];
  return GRSR_OK;
}

int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

// Per-device working set of the host-buffer entry points: one stream and grow-only buffers that
// survive between calls (a distance matrix call is ~20 ms of kernel time; allocating and freeing
// 20 MB around it would cost as much again).  One call at a time per device (mutex).
const int kCacheSlots = 11;  // 0 genomes, 1 inverses, 2 pairs, 3 results, 4-5 large-build scratch, 6 error word,
                             // 7 unused, 8 distance matrix, 9 stepwise state, 10 pair tickets
struct DevCache {
  std::recursive_mutex mu;
  cudaStream_t stream = nullptr;
  void *buf[kCacheSlots] = {};
  size_t cap[kCacheSlots] = {};
};
const int kMaxDevices = 64;
DevCache g_cache[kMaxDevices];

int cache_reserve(DevCache &c, int slot, size_t bytes, void **out) {
  if (c.cap[slot] < bytes) {
    if (c.buf[slot]) { cudaFree(c.buf[slot]); c.buf[slot] = nullptr; c.cap[slot] = 0; }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&c.buf[slot], want);
    if (e != cudaSuccess)
      return fail(GRSR_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    c.cap[slot] = want;
  }
  *out = c.buf[slot];
  return GRSR_OK;
}

struct Launch { int threads; int ept; int grid; size_t smem; };

// cudaFuncSetAttribute + cudaOccupancyMaxActiveBlocksPerMultiprocessor cost microseconds each and a
// distance call asks dozens of times (every candidate block shape of the walk kernels): remembered
// per (kernel, device, threads, dynamic shared memory).
struct OccKey {
  const void *fn; int dev, threads; size_t smem;
  bool operator<(const OccKey &o) const {
    if (fn != o.fn) return fn < o.fn;
    if (dev != o.dev) return dev < o.dev;
    if (threads != o.threads) return threads < o.threads;
    return smem < o.smem;
  }
};
std::mutex g_occ_mu;
std::map<OccKey, int> g_occ;
std::map<std::pair<const void *, int>, size_t> g_smem_set;   // largest dynamic smem opted in per (kernel, device)

template <class K>
int cached_occupancy(K kernel, int dev, int threads, size_t smem, int *occ) {
  std::lock_guard<std::mutex> lk(g_occ_mu);
  const void *fn = (const void *)kernel;
  OccKey key = {fn, dev, threads, smem};
  size_t &have = g_smem_set[std::make_pair(fn, dev)];
  auto it = g_occ.find(key);
  if (it != g_occ.end() && have >= smem) { *occ = it->second; return GRSR_OK; }
  if (have == 0)
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                            (int)cudaSharedmemCarveoutMaxShared));
  if (have < smem || have == 0) {
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    have = smem ? smem : 1;
  }
  int o = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, threads, smem));
  g_occ[key] = o;
  *occ = o;
  return GRSR_OK;
}

// Shape, shared memory and grid of a one-pair-per-block kernel instance `kernel` (already
// specialised for L.ept): opt in to the large carve-out, ask the occupancy calculator how many
// blocks stay resident per SM and launch a whole number of resident waves (148 SMs x occupancy).
template <class K>
int finish_plan(K kernel, long long units, Launch &L) {
  int dev;
  CK(cudaGetDevice(&dev));
  DevInfo di;
  int rc = dev_info(dev, di);
  if (rc) return rc;
  if (L.smem > (size_t)di.smem_optin)
    return fail(GRSR_ERR_TOOLARGE, "%zu bytes of shared memory per pair exceed the device limit %d",
                L.smem, di.smem_optin);
  int occ = 0;
  if ((rc = cached_occupancy(kernel, dev, L.threads, L.smem, &occ))) return rc;
  if (occ < 1) return fail(GRSR_ERR_CUDA, "kernel does not fit on an SM (smem %zu)", L.smem);
  int cap = env_int("GRSR_OCC", occ);
  if (cap >= 1 && cap < occ) occ = cap;
  long long g = (long long)di.sms * occ;
  if (g > units) g = units;
  if (g < 1) g = 1;
  L.grid = (int)g;
  return GRSR_OK;
}

// Plan of the warp-per-pair first tier (grsr_walk.cuh): warps per block chosen so that the most
// warps stay resident per SM; grid = whole resident waves.  shared = the kernel whose warps share one
// staged pi row per block (pair lists sorted by pi row).
struct WalkPlan { int shift; int wpb; int grid; size_t smem; int warps_per_sm; };

template <class K>
int plan_walk(K kernel, bool shared, int n, long long npairs, WalkPlan &P) {
  int dev;
  CK(cudaGetDevice(&dev));
  DevInfo di;
  int rc = dev_info(dev, di);
  if (rc) return rc;
  // The private-row kernel takes the shape with the most resident warps (few large blocks first: less
  // reserved shared memory).  In the shared-row kernel a block waits for its slowest warp at every run
  // of equal pi rows, so it takes the SMALLEST block that reaches 85 % of the best warp count
  // (measured on config 3: 5 blocks x 4 warps beat 3 x 7 and 1 x 22 by 3 % and 17 %).
  int best_wpb = 0, best_occ = 0;
  const int forced = env_int(shared ? "GRSR_WALK_SHARED_WPB" : "GRSR_WALK_WPB", 0);
  int occ_of[33] = {0};
  // opt in once to the largest shape so that the per-shape queries below never lower the attribute
  for (int k = 0; k < (shared ? 32 : 8); k++) {
    const int wpb = shared ? k + 1 : 8 - k;
    if (forced && wpb != forced) continue;
    const size_t smem = shared ? 16 + grsr::walk_tail_bytes(n) + grsr::walk_private_bytes(n) * wpb
                               : grsr::walk_warp_bytes(n) * wpb;
    if (smem > (size_t)di.smem_optin) continue;
    int occ = 0;
    if ((rc = cached_occupancy(kernel, dev, wpb * 32, smem, &occ))) return rc;
    occ_of[wpb] = occ;
    if (occ * wpb > best_occ * best_wpb) { best_wpb = wpb; best_occ = occ; }
  }
  if (shared && best_wpb)
    for (int wpb = 1; wpb < best_wpb; wpb++)
      if (occ_of[wpb] * wpb * 100 >= best_occ * best_wpb * 85) { best_wpb = wpb; best_occ = occ_of[wpb]; break; }
  if (!best_wpb) return fail(GRSR_ERR_TOOLARGE, "walk kernel does not fit on an SM (n=%d)", n);
  int cap = env_int("GRSR_WALK_OCC", best_occ);
  if (cap >= 1 && cap < best_occ) best_occ = cap;
  P.shift = grsr::walk_shift(n);
  int fs = env_int("GRSR_WALK_SHIFT", 0);
  if (fs >= P.shift && fs <= 10) P.shift = fs;
  int idle = env_int("GRSR_WALK_IDLE", grsr::kWalkIdle);
  if (idle < 1 || idle > 32) idle = grsr::kWalkIdle;
  P.shift = grsr::walk_shape_arg(P.shift, idle) | ((env_int("GRSR_WALK_GUIDE4", 0) & 0xFF) << 16);
  P.wpb = best_wpb;
  P.smem = shared ? 16 + grsr::walk_tail_bytes(n) + grsr::walk_private_bytes(n) * best_wpb
                  : grsr::walk_warp_bytes(n) * best_wpb;
  P.warps_per_sm = best_wpb * best_occ;
  long long g = (long long)di.sms * best_occ, need = (npairs + best_wpb - 1) / best_wpb;
  if (g > need) g = need;
  if (g < 1) g = 1;
  P.grid = (int)g;
  return GRSR_OK;
}

const int kMaxLargeGenes = 1000000;   // global-memory build: 31-bit ids, 21-bit packed counters

// which build serves num_genes = n: the shared-memory one while the pair's graph fits on chip,
// the global-memory one beyond (or when GRSR_FORCE_LARGE is set, for tests)
size_t smem_budget() {
  int dev = 0;
  DevInfo di;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 227u * 1024u; }
  if (dev_info(dev, di) != GRSR_OK || di.smem_optin <= 0) return 227u * 1024u;
  return (size_t)di.smem_optin;
}

bool small_fits(int n, bool scenario) {
  if (env_int("GRSR_FORCE_LARGE", 0)) return false;
  if (n > grsr::kMaxSmemGenes || 2 * n + 2 > 32767) return false;
  int t, e;
  grsr::pick_shape(n, env_int("GRSR_THREADS", 0), &t, &e);
  if (t * e < 2 * n + 2) return false;
  size_t smem = scenario ? grsr::work_bytes(n, t * e, true) + grsr::scen_extra_bytes(n)
                         : grsr::work_bytes(n, t * e, false) + grsr::stage_bytes(n);
  return smem <= smem_budget();
}

int check_genes_fit(int n) {
  if (n > kMaxLargeGenes)
    return fail(GRSR_ERR_TOOLARGE, "num_genes=%d: this build accepts at most %d genes", n,
                kMaxLargeGenes);
  return GRSR_OK;
}

// Plan of the global-memory build.  With many pairs every pair gets one 1024-thread block; with few
// pairs (fewer than SMs) each pair gets a thread-block cluster of cs = 2..16 blocks (SMs) instead, so
// that a single long scenario still uses a good part of the chip.  scratch / xch come from the
// device's buffer cache (slots 4 and 5).
struct LargePlan { int cs; int teams; int grid; void *scratch; void *xch; };

template <class KB, class KC>
int plan_large(KB block_kernel, KC cluster_kernel, long long units, size_t bytes_per_team, int dev,
               LargePlan &P) {
  DevInfo di;
  int rc = dev_info(dev, di);
  if (rc) return rc;
  if (dev < 0 || dev >= kMaxDevices) return fail(GRSR_ERR_NODEVICE, "device %d out of range", dev);
  int cs = 1;
  while (cs < 16 && (long long)cs * 2 * units <= di.sms) cs *= 2;
  int forced = env_int("GRSR_CLUSTER", 0);
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8 || forced == 16) cs = forced;
  P.cs = cs;
  if (cs == 1) {
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, block_kernel, 1024, 0));
    if (occ < 1) return fail(GRSR_ERR_CUDA, "large kernel does not fit on an SM");
    long long g = (long long)di.sms * occ;
    if (g > units) g = units;
    if (g < 1) g = 1;
    P.teams = (int)g; P.grid = (int)g;
  } else {
    if (cs > 8) CK(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(cs); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 34 * 8;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int maxc = 0;
    CK(cudaOccupancyMaxActiveClusters(&maxc, cluster_kernel, &cfg));
    if (maxc < 1) return fail(GRSR_ERR_CUDA, "a cluster of %d blocks does not fit on this device", cs);
    long long t = maxc;
    if (t > units) t = units;
    if (t < 1) t = 1;
    P.teams = (int)t; P.grid = (int)t * cs;
  }
  DevCache &c = g_cache[dev];
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  if ((rc = cache_reserve(c, 4, bytes_per_team * (size_t)P.teams, &P.scratch))) return rc;
  return cache_reserve(c, 5, sizeof(unsigned long long) * grsr::kXchWords * (size_t)P.teams, &P.xch);
}

template <class K, class... Args>
int launch_cluster(K kernel, const LargePlan &P, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = P.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(P.grid); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 34 * 8;
  cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, kernel, args...));
  return GRSR_OK;
}

// Upper triangle enumerated column by column -- (0,1), (0,2), (1,2), (0,3), ... -- so that
// consecutive pairs share their pi row (the one whose signed inverse the kernel keeps staged):
// pair p = (i, j) with p = j(j-1)/2 + i, i < j.
__global__ void triangle_pairs_kernel(int G, long long first, long long count, int32_t *pairs) {
  long long stride = (long long)gridDim.x * blockDim.x;
  (void)G;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < count; q += stride) {
    long long p = first + q;
    long long j = (long long)((1.0 + sqrt(1.0 + 8.0 * (double)p)) * 0.5);
    while (j * (j - 1) / 2 > p) j--;
    while ((j + 1) * j / 2 <= p) j++;
    long long i = p - j * (j - 1) / 2;
    pairs[2 * q] = (int32_t)i; pairs[2 * q + 1] = (int32_t)j;
  }
}

__global__ void square_pairs_kernel(int G, long long first, long long count, int32_t *pairs) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < count; q += stride) {
    long long p = first + q;
    pairs[2 * q] = (int32_t)(p / G); pairs[2 * q + 1] = (int32_t)(p % G);
  }
}

int check_device_range(int first_device, int ngpus) {
  int cnt = grsr_device_count();
  if (cnt <= 0)
    return fail(GRSR_ERR_NODEVICE, "no CUDA device is available; libgrsr_cuda has no CPU path");
  if (ngpus < 1 || first_device < 0 || first_device + ngpus > cnt)
    return fail(GRSR_ERR_NODEVICE, "devices %d..%d requested but %d CUDA device(s) present",
                first_device, first_device + ngpus - 1, cnt);
  return GRSR_OK;
}

// check_genes of the reference (G/mcread_input.c:83-135) on the device, fused with the signed
// inverse: pass 1 rejects zero / too-large genes and records, per gene, the first position holding
// it; pass 2 flags every later position holding the same gene; pass 3 turns positions into signed
// inverses.  The first offence in reading order (smallest genome, then smallest position -- the
// reference's order) lands in *err as genome << 34 | position << 2 | kind.
__global__ void validate_first_positions(const int32_t *__restrict__ genomes, int32_t *slot,
                                         long long total, int n, unsigned long long *err) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    long long row = e / n; int pos = (int)(e - row * n);
    long long g = genomes[e], a = g < 0 ? -g : g;
    if (a == 0) atomicMin(err, ((unsigned long long)row << 34) | ((unsigned long long)pos << 2) | 0ull);
    else if (a > n) atomicMin(err, ((unsigned long long)row << 34) | ((unsigned long long)pos << 2) | 1ull);
    else atomicMin(&slot[row * n + (a - 1)], pos);
  }
}

__global__ void validate_duplicates(const int32_t *__restrict__ genomes, const int32_t *__restrict__ slot,
                                    long long total, int n, unsigned long long *err) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    long long row = e / n; int pos = (int)(e - row * n);
    long long g = genomes[e], a = g < 0 ? -g : g;
    if (a >= 1 && a <= n && slot[row * n + (a - 1)] != pos)
      atomicMin(err, ((unsigned long long)row << 34) | ((unsigned long long)pos << 2) | 2ull);
  }
}

// Pass 3.  If an offence was recorded the table must not reach the distance kernels (they index
// shared memory with its genes): the DEVICE COPY is then overwritten with identity rows, so that the
// launches already queued behind this one run on harmless data while the host learns of the error
// from the word it reads back together with the results.
__global__ void positions_to_signed_inverse(int32_t *__restrict__ genomes, int32_t *slot,
                                            long long total, int n,
                                            const unsigned long long *__restrict__ err) {
  long long stride = (long long)gridDim.x * blockDim.x;
  const bool bad = (*err != ~0ull);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    long long row = e / n;
    if (bad) { int v = (int)(e - row * n) + 1; genomes[e] = v; slot[e] = v; continue; }
    int pos = slot[e];
    slot[e] = genomes[row * n + pos] > 0 ? (pos + 1) : -(pos + 1);
  }
}

// Validates the resident genome table and leaves its signed inverses in d_sinv.  Nothing here
// waits for the device: *d_err (~0 = table is fine) is read back by the caller with its results and
// turned into the reference's message by validation_error().
int validate_and_invert(int32_t *d_genomes, int G, int n, int32_t *d_sinv, unsigned long long *d_err,
                        cudaStream_t st) {
  long long total = (long long)G * n;
  int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  CK(cudaMemsetAsync(d_sinv, 0x7f, (size_t)total * sizeof(int32_t), st));
  CK(cudaMemsetAsync(d_err, 0xff, sizeof(unsigned long long), st));
  validate_first_positions<<<blocks, 256, 0, st>>>(d_genomes, d_sinv, total, n, d_err);
  validate_duplicates<<<blocks, 256, 0, st>>>(d_genomes, d_sinv, total, n, d_err);
  positions_to_signed_inverse<<<blocks, 256, 0, st>>>(d_genomes, d_sinv, total, n, d_err);
  CK(cudaGetLastError());
  return GRSR_OK;
}

int validation_error(unsigned long long err, const int32_t *genomes_host, int n, long long row0 = 0) {
  if (err == ~0ull) return GRSR_OK;
  long long row = row0 + (long long)(err >> 34);
  int pos = (int)((err >> 2) & 0xFFFFFFFFull), kind = (int)(err & 3ull);
  long long g = genomes_host[row * n + pos], a = g < 0 ? -g : g;
  if (kind == 0) return fail(GRSR_ERR_INVALID, "genome %lld: input gene == 0", row + 1);
  if (kind == 1)
    return fail(GRSR_ERR_INVALID, "genome %lld: input gene (%lld) larger than number of genes %d",
                row + 1, g, n);
  return fail(GRSR_ERR_INVALID, "genome %lld: duplicate gene (%lld)", row + 1, a);
}

// check_genes of the reference (G/mcread_input.c:83-135): every row a signed permutation of 1..n
// shape check only; the contents of a distance job's table are validated on the device
int check_table(const int32_t *genomes, long long rows, int n) {
  if (!genomes || n < 1 || rows < 1) return fail(GRSR_ERR_INVALID, "empty genome table");
  return check_genes_fit(n);
}

int check_genomes(const int32_t *genomes, long long rows, int n) {
  if (!genomes || n < 1 || rows < 1) return fail(GRSR_ERR_INVALID, "empty genome table");
  std::vector<unsigned char> seen((size_t)n + 1);
  for (long long r = 0; r < rows; r++) {
    memset(seen.data(), 0, seen.size());
    const int32_t *g = genomes + r * n;
    for (int i = 0; i < n; i++) {
      long long a = g[i] < 0 ? -(long long)g[i] : g[i];
      if (a == 0) return fail(GRSR_ERR_INVALID, "genome %lld: input gene == 0", r + 1);
      if (a > n)
        return fail(GRSR_ERR_INVALID, "genome %lld: input gene (%d) larger than number of genes %d",
                    r + 1, g[i], n);
      if (seen[a]) return fail(GRSR_ERR_INVALID, "genome %lld: duplicate gene (%lld)", r + 1, a);
      seen[a] = 1;
    }
  }
  return GRSR_OK;
}

// The device's default memory pool keeps what is freed (release threshold = everything), so that the
// per-call cudaMallocAsync / cudaFreeAsync of the distance path costs no driver allocation after the
// first call.
int pool_ready(int dev) {
  static std::mutex mu;
  static bool done[kMaxDevices] = {};
  std::lock_guard<std::mutex> lk(mu);
  if (dev < 0 || dev >= kMaxDevices || done[dev]) return GRSR_OK;
  cudaMemPool_t pool;
  CK(cudaDeviceGetDefaultMemPool(&pool, dev));
  unsigned long long keep = ~0ull;
  CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  done[dev] = true;
  return GRSR_OK;
}

enum PairSource { PAIRS_LIST, PAIRS_TRIANGLE, PAIRS_SQUARE };

int dist_pairs_dev_impl(const int32_t *d_genomes, const int32_t *d_sinv, int num_genes,
                        const int32_t *d_pairs, int64_t npairs, int32_t *d_out, int stride,
                        cudaStream_t st, int order_hint);

// The caller's current device is left as it was found (the host-buffer entry points switch devices).
struct DeviceGuard {
  int saved = -1;
  DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) { cudaGetLastError(); saved = -1; } }
  ~DeviceGuard() { if (saved >= 0) cudaSetDevice(saved); }
};

// Lets kernels on device `dev` store into / draw tickets from memory of device `owner` (NVLink peer
// mapping).  False when the two devices cannot do that.
bool peer_ok(int dev, int owner) {
  if (dev == owner) return true;
  static std::mutex mu;
  static signed char known[kMaxDevices][kMaxDevices] = {};   // 0 unknown, 1 enabled, -1 impossible
  if (dev < 0 || owner < 0 || dev >= kMaxDevices || owner >= kMaxDevices) return false;
  std::lock_guard<std::mutex> lk(mu);
  if (known[dev][owner]) return known[dev][owner] > 0;
  known[dev][owner] = -1;
  int can = 0;
  if (cudaDeviceCanAccessPeer(&can, dev, owner) != cudaSuccess || !can) { cudaGetLastError(); return false; }
  int cur = -1;
  cudaGetDevice(&cur);
  if (cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return false; }
  cudaError_t e = cudaDeviceEnablePeerAccess(owner, 0);
  if (cur >= 0) cudaSetDevice(cur);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); known[dev][owner] = 1; return true; }
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  known[dev][owner] = 1;
  return true;
}

// One device's share of a pair job: copy the genome table in, validate it and build the signed
// inverses, obtain the pair slice (copied or generated on the device), run the kernels, copy the
// results out -- one stream, one synchronisation at the end; the validation verdict travels back
// with the results.  stride 1 / 5: results go to out_host.  stride -G (matrix mode): the kernels
// scatter d into d_matrix[i][j] and [j][i] themselves; d_matrix may belong to another device.
int run_pairs_on_device(int dev, const int32_t *genomes, int G, int n, PairSource src,
                        const int32_t *pairs_host, long long first, long long count,
                        int32_t *out_host, int stride, int32_t *d_matrix) {
  if (count <= 0) return GRSR_OK;
  if (dev < 0 || dev >= kMaxDevices) return fail(GRSR_ERR_NODEVICE, "device %d out of range", dev);
  CK(cudaSetDevice(dev));
  DevCache &c = g_cache[dev];
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  if (!c.stream) CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  cudaStream_t st = c.stream;
  size_t gbytes = (size_t)G * n * sizeof(int32_t);
  void *dg, *ds, *dp, *dout = nullptr, *derr;
  int rc;
  if ((rc = cache_reserve(c, 0, gbytes, &dg)) || (rc = cache_reserve(c, 1, gbytes, &ds)) ||
      (rc = cache_reserve(c, 2, (size_t)count * 2 * sizeof(int32_t), &dp)) ||
      (rc = cache_reserve(c, 6, sizeof(unsigned long long), &derr)))
    return rc;
  if (stride > 0 && (rc = cache_reserve(c, 3, (size_t)count * stride * sizeof(int32_t), &dout))) return rc;
  // the pair slice does not depend on the table: it is generated while the table is still in flight
  if (src == PAIRS_LIST) {
    CK(cudaMemcpyAsync(dp, pairs_host + 2 * first, (size_t)count * 2 * sizeof(int32_t),
                       cudaMemcpyHostToDevice, st));
  } else if (src == PAIRS_TRIANGLE) {
    if ((rc = grsr_triangle_pairs_dev(G, first, count, (int32_t *)dp, st))) return rc;
  } else {
    int blocks = (int)((count + 255) / 256 < 1184 ? (count + 255) / 256 : 1184);
    square_pairs_kernel<<<blocks, 256, 0, st>>>(G, first, count, (int32_t *)dp);
    CK(cudaGetLastError());
  }
  CK(cudaMemcpyAsync(dg, genomes, gbytes, cudaMemcpyHostToDevice, st));
  if ((rc = validate_and_invert((int32_t *)dg, G, n, (int32_t *)ds, (unsigned long long *)derr, st)))
    return rc;
  // the triangle is sorted by pi row by construction: no order test on the device
  const int hint = (src == PAIRS_TRIANGLE) ? 1 : (src == PAIRS_SQUARE ? 0 : -1);
  if ((rc = dist_pairs_dev_impl((const int32_t *)dg, (const int32_t *)ds, n, (const int32_t *)dp, count,
                                stride > 0 ? (int32_t *)dout : d_matrix, stride, st, hint)))
    return rc;
  unsigned long long err = ~0ull;
  if (stride > 0)
    CK(cudaMemcpyAsync(out_host, dout, (size_t)count * stride * sizeof(int32_t),
                       cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&err, derr, sizeof err, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return validation_error(err, genomes, n);
}

// Persistent host threads for the multi-device calls: creating N - 1 threads per call costs ~25 us each,
// one after the other, which is a quarter of a millisecond before the last device even starts -- as much
// as a third of the kernel time of an 8-way matrix.  Workers sleep on a condition variable between calls.
struct WorkerPool {
  std::recursive_mutex call_mu;       // one multi-device call at a time; taken BEFORE any per-device lock
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  std::vector<std::thread> threads;
  std::function<void(int)> fn;
  int gen = 0, want = 0, left = 0;
  void start(int nworkers, std::function<void(int)> f) {
    std::unique_lock<std::mutex> lk(mu);
    while ((int)threads.size() < nworkers) {
      const int id = (int)threads.size() + 1;
      threads.emplace_back([this, id] { loop(id); });
    }
    fn = std::move(f); want = nworkers; left = nworkers; gen++;
    cv_go.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { return left == 0; });
  }
  void loop(int id) {
    int seen = 0;
    for (;;) {
      std::function<void(int)> f;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_go.wait(lk, [&] { return gen != seen; });
        seen = gen;
        if (id > want) continue;
        f = fn;
      }
      f(id);
      std::lock_guard<std::mutex> lk(mu);
      if (--left == 0) cv_done.notify_all();
    }
  }
};
WorkerPool &worker_pool() { static WorkerPool *p = new WorkerPool; return *p; }   // never destroyed: its threads outlive main

// Runs work(d) for d = 0 .. ngpus-1, device first_device + d: d = 0 on the calling thread, the others
// on the pool's threads; the first error (lowest device) is reported.
template <class F>
int on_devices(int first_device, int ngpus, F work) {
  std::vector<int> rcs(ngpus, 0);
  std::vector<std::string> errs(ngpus);
  if (ngpus > 1) {
    WorkerPool &pool = worker_pool();
    std::lock_guard<std::recursive_mutex> one_call(pool.call_mu);
    pool.start(ngpus - 1, [&](int d) { rcs[d] = work(d); if (rcs[d]) errs[d] = g_err; });
    rcs[0] = work(0);
    if (rcs[0]) errs[0] = g_err;
    pool.wait();
  } else {
    rcs[0] = work(0);
    if (rcs[0]) errs[0] = g_err;
  }
  for (int d = 0; d < ngpus; d++) {
    if (!rcs[d]) continue;
    if (ngpus == 1) return fail(rcs[d], "%s", errs[d].c_str());
    return fail(rcs[d], "device %d: %s", first_device + d, errs[d].c_str());
  }
  return GRSR_OK;
}

// Split [0, total) into ngpus contiguous slices, one host thread per device; no data-path collective.
int shard_pairs(const int32_t *genomes, int G, int n, PairSource src, const int32_t *pairs_host,
                long long total, int32_t *out_host, int stride, int first_device, int ngpus) {
  return on_devices(first_device, ngpus, [&](int d) {
    long long lo = total * d / ngpus, hi = total * (d + 1) / ngpus;
    return run_pairs_on_device(first_device + d, genomes, G, n, src, pairs_host, lo, hi - lo,
                               out_host + lo * stride, stride, nullptr);
  });
}

// (Optional, GRSR_GATHER=1.)  The table of a multi-device matrix call, the NVLink way: device d uploads, validates and inverts only
// ITS 1/N of the genome rows (one PCIe copy of the table in total instead of N), records an event, and
// every device then pulls the other slices -- genomes and signed inverses -- from their owners with peer
// copies.  State shared by the N host threads of one call.
struct TableGather {
  int ngpus = 0;
  void *dg[kMaxDevices] = {}, *ds[kMaxDevices] = {};
  cudaEvent_t ready[kMaxDevices] = {};
  std::atomic<int> arrived{0};
  std::atomic<int> failed{0};
};
cudaEvent_t g_gather_event[kMaxDevices] = {};

// One device's share of a matrix call with a gathered table (matrix mode only: results are scattered into
// d_matrix by the kernels).
int run_matrix_slice_gathered(TableGather &T, int d, int first_device, const int32_t *genomes, int G, int n,
                              long long first, long long count, int32_t *d_matrix) {
  const int dev = first_device + d, N = T.ngpus;
  auto arrive = [&](bool ok) {                          // every thread passes here exactly once, error or not
    if (!ok) T.failed.store(1);
    T.arrived.fetch_add(1);
    while (T.arrived.load() < N) std::this_thread::yield();
  };
  if (dev < 0 || dev >= kMaxDevices) { arrive(false); return fail(GRSR_ERR_NODEVICE, "device %d out of range", dev); }
  cudaError_t e0 = cudaSetDevice(dev);
  DevCache &c = g_cache[dev];
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  const size_t gbytes = (size_t)G * n * sizeof(int32_t);
  void *dg = nullptr, *ds = nullptr, *dp = nullptr, *derr = nullptr;
  int rc = GRSR_OK;
  const long long r0 = (long long)G * d / N, r1 = (long long)G * (d + 1) / N;
  bool ok = (e0 == cudaSuccess);
  if (ok && !c.stream) ok = cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) == cudaSuccess;
  if (ok && !g_gather_event[dev]) ok = cudaEventCreateWithFlags(&g_gather_event[dev], cudaEventDisableTiming) == cudaSuccess;
  cudaStream_t st = c.stream;
  if (ok) {
    rc = cache_reserve(c, 0, gbytes, &dg);
    if (!rc) rc = cache_reserve(c, 1, gbytes, &ds);
    if (!rc) rc = cache_reserve(c, 2, (size_t)(count > 0 ? count : 1) * 2 * sizeof(int32_t), &dp);
    if (!rc) rc = cache_reserve(c, 6, sizeof(unsigned long long), &derr);
    ok = (rc == GRSR_OK);
  }
  if (ok) {
    for (int q = 0; q < N; q++) peer_ok(dev, first_device + q);
    T.dg[d] = dg; T.ds[d] = ds; T.ready[d] = g_gather_event[dev];
    if (count > 0) ok = grsr_triangle_pairs_dev(G, first, count, (int32_t *)dp, st) == GRSR_OK;
    const size_t off = (size_t)r0 * n, sl = (size_t)(r1 - r0) * n;
    ok = ok && cudaMemcpyAsync((int32_t *)dg + off, genomes + off, sl * sizeof(int32_t), cudaMemcpyHostToDevice, st) == cudaSuccess;
    if (ok && r1 > r0)
      ok = validate_and_invert((int32_t *)dg + off, (int)(r1 - r0), n, (int32_t *)ds + off, (unsigned long long *)derr, st) == GRSR_OK;
    else if (ok)
      ok = cudaMemsetAsync(derr, 0xff, sizeof(unsigned long long), st) == cudaSuccess;
    ok = ok && cudaEventRecord(T.ready[d], st) == cudaSuccess;
  }
  arrive(ok);                                           // all slices are on their way (or somebody failed)
  if (!ok || T.failed.load()) {
    if (e0 == cudaSuccess && c.stream) cudaStreamSynchronize(c.stream);
    if (ok) return GRSR_OK;                             // another device reports the error
    if (rc) return rc;
    cudaError_t e = cudaGetLastError();
    return fail(GRSR_ERR_CUDA, "table upload failed on device %d: %s", dev, cudaGetErrorString(e != cudaSuccess ? e : e0));
  }
  unsigned long long err = ~0ull;
  const int rcb = [&]() -> int {
    for (int q = 0; q < N; q++) {                        // pull the other slices from their owners over NVLink
      if (q == d) continue;
      const long long q0 = (long long)G * q / N, q1 = (long long)G * (q + 1) / N;
      if (q1 <= q0) continue;
      const size_t off = (size_t)q0 * n, bytes = (size_t)(q1 - q0) * n * sizeof(int32_t);
      CK(cudaStreamWaitEvent(st, T.ready[q], 0));
      CK(cudaMemcpyPeerAsync((int32_t *)dg + off, dev, (int32_t *)T.dg[q] + off, first_device + q, bytes, st));
      CK(cudaMemcpyPeerAsync((int32_t *)ds + off, dev, (int32_t *)T.ds[q] + off, first_device + q, bytes, st));
    }
    if (count > 0) {
      int r = dist_pairs_dev_impl((const int32_t *)dg, (const int32_t *)ds, n, (const int32_t *)dp, count, d_matrix,
                                  -G, st, 1);
      if (r) return r;
    }
    CK(cudaMemcpyAsync(&err, derr, sizeof err, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return GRSR_OK;
  }();
  if (rcb) cudaStreamSynchronize(st);
  // nobody may reuse or free a slice while a peer still copies from it: leave together
  T.arrived.fetch_add(1);
  while (T.arrived.load() < 2 * N) std::this_thread::yield();
  if (rcb) return rcb;
  return validation_error(err, genomes, n, r0);          // the error word counts rows from the slice's first
}

// setinvmatrix on ngpus devices.  The G x G matrix lives on the first device; every device's kernels
// store their distances straight into it, both [i][j] and [j][i] (the other devices through NVLink
// peer mappings, 8 bytes per pair), the first device zeroes the diagonal, and ONE device-to-host copy
// delivers the matrix.  Without peer access between the devices: flat slices to the host and a host
// scatter.
int matrix_on_devices(const int32_t *genomes, int G, int n, int32_t *distmatrix, int first_device,
                      int ngpus) {
  const long long P = (long long)G * (G - 1) / 2;
  bool peers = true;
  for (int d = 1; d < ngpus; d++) peers = peers && peer_ok(first_device + d, first_device);
  if (env_int("GRSR_NO_PEER", 0)) peers = (ngpus == 1);
  if (!peers) {
    std::vector<int32_t> flat((size_t)P);
    int rc = shard_pairs(genomes, G, n, PAIRS_TRIANGLE, nullptr, P, flat.data(), 1, first_device, ngpus);
    if (rc) return rc;
    long long p = 0;
    for (long long j = 0; j < G; j++) {
      distmatrix[j * G + j] = 0;
      for (long long i = 0; i < j; i++, p++) distmatrix[i * G + j] = distmatrix[j * G + i] = flat[p];
    }
    return GRSR_OK;
  }
  // lock order of every multi-device call: the pool, then device caches (the first device's here, the
  // others' in their worker threads)
  std::unique_lock<std::recursive_mutex> pool_lock(worker_pool().call_mu, std::defer_lock);
  if (ngpus > 1) pool_lock.lock();
  CK(cudaSetDevice(first_device));
  DevCache &c = g_cache[first_device];
  std::lock_guard<std::recursive_mutex> lk(c.mu);       // held for the whole call: the matrix is ours
  if (!c.stream) CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  void *dmat;
  int rc = cache_reserve(c, 8, (size_t)G * G * sizeof(int32_t), &dmat);
  if (rc) return rc;
  CK(cudaMemset2DAsync(dmat, (size_t)(G + 1) * sizeof(int32_t), 0, sizeof(int32_t), (size_t)G, c.stream));
  // Measured at N = 8 on config 3 (gpurun_out/r2w_bench_n8*.json): 1.283 ms per call with the gathered
  // table against 1.228 ms with eight full uploads -- the 14 peer copies and the host rendezvous cost
  // more than the PCIe time they save (the eight links run in parallel).  Kept as an option, off by default.
  if (ngpus > 1 && G >= ngpus && env_int("GRSR_GATHER", 0)) {
    TableGather T;
    T.ngpus = ngpus;
    rc = on_devices(first_device, ngpus, [&](int d) {
      long long lo = P * d / ngpus, hi = P * (d + 1) / ngpus;
      return run_matrix_slice_gathered(T, d, first_device, genomes, G, n, lo, hi - lo, (int32_t *)dmat);
    });
  } else {
    rc = on_devices(first_device, ngpus, [&](int d) {
      long long lo = P * d / ngpus, hi = P * (d + 1) / ngpus;
      return run_pairs_on_device(first_device + d, genomes, G, n, PAIRS_TRIANGLE, nullptr, lo, hi - lo,
                                 nullptr, -G, (int32_t *)dmat);
    });
  }
  if (rc) return rc;
  CK(cudaSetDevice(first_device));
  CK(cudaMemcpyAsync(distmatrix, dmat, (size_t)G * G * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  return GRSR_OK;
}

// ---- scenarios ----------------------------------------------------------------------------------------

// Results of a scenario call are written by the kernels straight into page-locked host memory that
// every device maps (8 bytes per emitted step: nothing next to a step's cost), so that any device
// can serve any pair of the call's queue.  One grow-only arena per process; a scenario call owns it
// from start to end (scenario calls are serialised process-wide).
struct HostArena {
  std::mutex mu;
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes, void **out) {
    if (cap < bytes) {
      if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
      size_t want = bytes + bytes / 4 + 4096;
      cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable | cudaHostAllocMapped);
      if (e != cudaSuccess) {
        p = nullptr;
        return fail(GRSR_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", want, cudaGetErrorString(e));
      }
      cap = want;
    }
    *out = p;
    return GRSR_OK;
  }
};
HostArena g_arena;

struct ScenOut {                 // views into the arena
  int32_t *steps;                // npairs x max_steps x 2
  int32_t *nsteps, *status;      // npairs
  unsigned long long *evals;     // npairs
};

// Few pairs, on-chip sizes: while a pair's graph has hurdles (every breakpoint pair must be
// evaluated, hundreds to thousands per step) each step's evaluations are spread over the whole GPU
// (grsr_stepwise.cuh); the hurdle-free rest of every pair is left to the persistent kernel.
// d_src is the device copy of the sources (advanced in place).
int stepwise_prefix(DevCache &c, cudaStream_t st, int32_t *d_src, const int32_t *d_dst, int npairs, int n,
                    const ScenOut &O, int max_steps, const Launch &L) {
  const int cap = (2 * (n + 1) > 4096) ? 2 * (n + 1) : 4096;
  void *dstate, *dcands, *dd2;
  int r;
  if ((r = cache_reserve(c, 9, sizeof(grsr::StepState), &dstate)) ||
      (r = cache_reserve(c, 2, (size_t)cap * 2 * sizeof(int32_t), &dcands)) ||
      (r = cache_reserve(c, 3, (size_t)cap * sizeof(int32_t), &dd2)))
    return r;
  Launch LT;
  LT.threads = L.threads; LT.ept = L.ept;
  LT.smem = grsr::work_bytes(n, L.threads * L.ept, false) + 2 * grsr::scen_arr_bytes(n);
  GRSR_DISPATCH_SHAPE(L.threads, L.ept, r = finish_plan(grsr::trial_batch_kernel<kNt, kEpt>, cap, LT));
  if (r) return r;
  { Launch LPp = L;
    GRSR_DISPATCH_SHAPE(L.threads, L.ept, r = finish_plan(grsr::step_prepare_kernel<kNt, kEpt>, 1, LPp));
    if (r) return r; }
  const int32_t *ncand_dev = (const int32_t *)((const char *)dstate + offsetof(grsr::StepState, ncand));
  for (int p = 0; p < npairs; p++) {
    grsr::StepState hs;
    memset(&hs, 0, sizeof hs);
    hs.lastkey = ~0ull; hs.batch_end = ~0ull;
    CK(cudaMemcpyAsync(dstate, &hs, sizeof hs, cudaMemcpyHostToDevice, st));
    int32_t *cur_p = d_src + (size_t)p * n;
    const int32_t *dst_p = d_dst + (size_t)p * n;
    int32_t *steps_p = O.steps + 2 * (size_t)p * max_steps;
    // Eight batches are queued between two looks at the state: the three kernels are idempotent once the
    // pair is sorted, out of steps or stuck, and keep making correct (filtered, one-candidate) steps if the
    // last hurdle goes mid-way, so running a few rounds too many only costs their launch -- while a host
    // round trip per batch was most of a cheap step's time.
    const int kBatchesPerSync = 8;
    for (;;) {
      for (int rep = 0; rep < kBatchesPerSync; rep++) {
        GRSR_DISPATCH_SHAPE(L.threads, L.ept, (grsr::step_prepare_kernel<kNt, kEpt><<<1, L.threads, L.smem, st>>>(
            cur_p, dst_p, n, (grsr::StepState *)dstate, (int32_t *)dcands, cap)));
        GRSR_DISPATCH_SHAPE(L.threads, L.ept, (grsr::trial_batch_kernel<kNt, kEpt><<<LT.grid, LT.threads, LT.smem, st>>>(
            cur_p, dst_p, n, (const int32_t *)dcands, 0LL, (int32_t *)dd2, ncand_dev)));
        grsr::step_select_kernel<<<1, 256, 0, st>>>(cur_p, n, (grsr::StepState *)dstate,
                                                    (const int32_t *)dcands, (const int32_t *)dd2,
                                                    steps_p, max_steps);
      }
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(&hs, dstate, sizeof hs, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (hs.d == 0 || hs.status != GRSR_SCEN_OK || hs.h == 0) break;
    }
    O.nsteps[p] = hs.nsteps;
    O.evals[p] = hs.evals;
    O.status[p] = hs.status;      // OVERFLOW / STUCK end the pair here; the persistent kernel skips it
  }
  return GRSR_OK;
}

// One device's part of a scenario call: copy all pairs in (or, mapped = true, read them where they are:
// page-locked host memory every device maps), run the persistent kernel on the queue.
int scenario_on_device(int dev, const int32_t *src, const int32_t *dst, int npairs, int n,
                       const ScenOut &O, int max_steps, grsr::PairQueue queue, bool stepwise, bool mapped) {
  if (dev < 0 || dev >= kMaxDevices) return fail(GRSR_ERR_NODEVICE, "device %d out of range", dev);
  CK(cudaSetDevice(dev));
  DevCache &c = g_cache[dev];
  std::lock_guard<std::recursive_mutex> device_lock(c.mu);   // the global-memory build's workspace is per device
  if (!c.stream) CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  cudaStream_t st = c.stream;
  DevInfo di;
  int rc = dev_info(dev, di);
  if (rc) return rc;
  const bool small = small_fits(n, true);
  Launch L;
  L.threads = 1024; L.ept = 0; L.smem = 0; L.grid = 1;
  LargePlan LP = {1, 1, 1, nullptr, nullptr};
  if (small) {
    // latency shape (more threads per pair) when every pair of the call is resident at once
    grsr::pick_shape(n, env_int("GRSR_THREADS", 0), &L.threads, &L.ept, 8);
    L.smem = grsr::work_bytes(n, L.threads * L.ept, true) + grsr::scen_extra_bytes(n);
    GRSR_DISPATCH_SHAPE(L.threads, L.ept, rc = finish_plan(grsr::scenario_kernel<kNt, kEpt>, npairs, L));
    if (!rc && L.grid < npairs) {            // more pairs than resident blocks: throughput shape
      grsr::pick_shape(n, env_int("GRSR_THREADS", 0), &L.threads, &L.ept);
      L.smem = grsr::work_bytes(n, L.threads * L.ept, true) + grsr::scen_extra_bytes(n);
      GRSR_DISPATCH_SHAPE(L.threads, L.ept, rc = finish_plan(grsr::scenario_kernel<kNt, kEpt>, npairs, L));
    }
  } else {
    rc = plan_large(grsr::scenario_large_kernel, grsr::scenario_cluster_kernel, npairs,
                    grsr::scen_large_bytes(n), dev, LP);
  }
  if (rc) return rc;
  void *dsrc = (void *)src, *ddst = (void *)dst;
  if (!mapped) {
    size_t gbytes = (size_t)npairs * n * sizeof(int32_t);
    if ((rc = cache_reserve(c, 0, gbytes, &dsrc)) || (rc = cache_reserve(c, 1, gbytes, &ddst))) return rc;
    CK(cudaMemcpyAsync(dsrc, src, gbytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ddst, dst, gbytes, cudaMemcpyHostToDevice, st));
  }
  int resume = 0;
  if (stepwise && small) {
    if ((rc = stepwise_prefix(c, st, (int32_t *)dsrc, (const int32_t *)ddst, npairs, n, O, max_steps, L)))
      return rc;
    resume = 1;
  }
  if (small) {
    GRSR_DISPATCH_SHAPE(L.threads, L.ept, (grsr::scenario_kernel<kNt, kEpt><<<L.grid, L.threads, L.smem, st>>>(
        (const int32_t *)dsrc, (const int32_t *)ddst, npairs, n, O.steps, O.nsteps, max_steps, O.status,
        O.evals, queue, resume)));
  } else if (LP.cs == 1) {
    grsr::scenario_large_kernel<<<LP.grid, 1024, 0, st>>>(
        (const int32_t *)dsrc, (const int32_t *)ddst, npairs, n, O.steps, O.nsteps, max_steps, O.status,
        O.evals, queue, (unsigned char *)LP.scratch);
  } else {
    int r = launch_cluster(grsr::scenario_cluster_kernel, LP, st, (const int32_t *)dsrc,
                           (const int32_t *)ddst, npairs, n, O.steps, O.nsteps, max_steps, O.status,
                           O.evals, queue, (unsigned char *)LP.scratch, (unsigned long long *)LP.xch);
    if (r) return r;
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  return GRSR_OK;
}

// The scenario call behind grsr_scenario_batch / grsr_scenario_prefix.  All devices draw pairs from
// ONE ticket (a word in the first device's memory, reached by the others through NVLink peer mapping
// with system-scope atomics): per-pair work differs by orders of magnitude, so any static split would
// leave devices idle.  Without peer access: interleaved split (pair p on device p mod ngpus).
int scenario_call(const int32_t *src, const int32_t *dst, int npairs, int n, grsr_step_t *steps,
                  int32_t *nsteps, int max_steps, int32_t *complete, int64_t *evals, bool prefix,
                  int first_device, int ngpus) {
  int rc = check_genes_fit(n);
  if (rc) return rc;
  std::lock_guard<std::mutex> arena_lock(g_arena.mu);
  std::unique_lock<std::recursive_mutex> pool_lock(worker_pool().call_mu, std::defer_lock);
  if (ngpus > 1) pool_lock.lock();                      // before any per-device lock (see matrix_on_devices)
  const size_t ms = (size_t)(max_steps > 0 ? max_steps : 1);
  const size_t sbytes = (((size_t)npairs * ms * 2 * sizeof(int32_t)) + 15) & ~(size_t)15;
  const size_t ibytes = (((size_t)npairs * sizeof(int32_t)) + 15) & ~(size_t)15;
  // Several devices: a pair is read ONCE, by whichever device draws it, so the genomes are not copied to
  // every device (N x the whole batch through the host's staging buffers) but placed in the page-locked
  // arena, which every device maps, by the pool's threads in parallel; the kernels read them over PCIe
  // (12n bytes per pair against milliseconds of work).
  const bool stepwise_call = npairs <= env_int("GRSR_STEPWISE_MAX", 16);
  const bool mapped = ngpus > 1 && !stepwise_call && !env_int("GRSR_NO_MAPPED_INPUT", 0);
  const size_t obytes = sbytes + 2 * ibytes + ((((size_t)npairs * sizeof(unsigned long long)) + 15) & ~(size_t)15);
  const size_t gbytes_all = (((size_t)npairs * n * sizeof(int32_t)) + 15) & ~(size_t)15;
  void *base;
  if ((rc = g_arena.reserve(obytes + (mapped ? 2 * gbytes_all : 0), &base))) return rc;
  ScenOut O;
  O.steps = (int32_t *)base;
  O.nsteps = (int32_t *)((char *)base + sbytes);
  O.status = (int32_t *)((char *)base + sbytes + ibytes);
  O.evals = (unsigned long long *)((char *)base + sbytes + 2 * ibytes);
  memset(O.nsteps, 0, 2 * ibytes + (size_t)npairs * sizeof(unsigned long long));

  const int32_t *ksrc = src, *kdst = dst;
  if (mapped) {
    int32_t *msrc = (int32_t *)((char *)base + obytes), *mdst = (int32_t *)((char *)base + obytes + gbytes_all);
    on_devices(first_device, ngpus, [&](int d) -> int {
      const size_t lo = (size_t)npairs * d / ngpus * n, hi = (size_t)npairs * (d + 1) / ngpus * n;
      memcpy(msrc + lo, src + lo, (hi - lo) * sizeof(int32_t));
      memcpy(mdst + lo, dst + lo, (hi - lo) * sizeof(int32_t));
      return GRSR_OK;
    });
    ksrc = msrc; kdst = mdst;
  }
  CK(cudaSetDevice(first_device));
  const bool stepwise = stepwise_call;
  if (stepwise) ngpus = 1;                              // a handful of pairs: one device, whole-GPU steps
  bool peers = true;
  for (int d = 1; d < ngpus; d++) peers = peers && peer_ok(first_device + d, first_device);
  if (env_int("GRSR_NO_PEER", 0)) peers = (ngpus == 1);
  // tickets: word d of a small buffer on the first device (word 0 is the shared queue)
  DevCache &c0 = g_cache[first_device];
  void *dtick;
  {
    std::lock_guard<std::recursive_mutex> lk(c0.mu);
    if (!c0.stream) CK(cudaStreamCreateWithFlags(&c0.stream, cudaStreamNonBlocking));
    if ((rc = cache_reserve(c0, 10, 256, &dtick))) return rc;
    CK(cudaMemsetAsync(dtick, 0, 256, c0.stream));
    CK(cudaStreamSynchronize(c0.stream));               // zero before any device draws
  }
  rc = on_devices(first_device, ngpus, [&](int d) -> int {
    grsr::PairQueue q;
    if (peers) { q.ticket = (unsigned int *)dtick; q.first = 0; q.stride = 1; }
    else {
      // private queue per device (its own ticket word is allocated on that device)
      int dev = first_device + d, r;
      void *own;
      CK(cudaSetDevice(dev));
      { std::lock_guard<std::recursive_mutex> lk(g_cache[dev].mu);
        if (!g_cache[dev].stream) CK(cudaStreamCreateWithFlags(&g_cache[dev].stream, cudaStreamNonBlocking));
        if ((r = cache_reserve(g_cache[dev], 10, 256, &own))) return r;
        CK(cudaMemsetAsync(own, 0, 256, g_cache[dev].stream)); }
      q.ticket = (unsigned int *)own; q.first = d; q.stride = ngpus;
    }
    return scenario_on_device(first_device + d, ksrc, kdst, npairs, n, O, max_steps, q, stepwise, mapped);
  });
  if (rc) return rc;
  for (int p = 0; p < npairs; p++) {
    if (O.status[p] == GRSR_SCEN_STUCK)
      return fail(GRSR_ERR_SEARCH, "pair %d: reversals ended prematurely after %d steps", p, O.nsteps[p]);
    if (O.status[p] == GRSR_SCEN_OVERFLOW && !prefix)
      return fail(GRSR_ERR_INVALID, "pair %d: scenario needs more than max_steps=%d steps", p, max_steps);
  }
  for (int p = 0; p < npairs; p++) {
    nsteps[p] = O.nsteps[p];
    if (O.nsteps[p] > 0)
      memcpy(steps + (size_t)p * max_steps, O.steps + 2 * (size_t)p * ms, (size_t)O.nsteps[p] * sizeof(grsr_step_t));
    if (complete) complete[p] = (O.status[p] == GRSR_SCEN_OK) ? 1 : 0;
    if (evals) evals[p] = (int64_t)O.evals[p];
  }
  return GRSR_OK;
}

// The batched distance on resident data.  stride 1 / 5 / -G as in put_result (grsr_core.cuh).
// order_hint: 1 = the list is sorted by pi row (shared-row walk kernel), 0 = it is not (private rows),
// -1 = unknown: walk_order_kernel counts the row changes and both kernels are launched, the one the
// verdict does not name returns at once.
int dist_pairs_dev_impl(const int32_t *d_genomes, const int32_t *d_sinv, int num_genes,
                        const int32_t *d_pairs, int64_t npairs, int32_t *d_out, int stride,
                        cudaStream_t st, int order_hint) {
  if (npairs <= 0) return GRSR_OK;
  if (num_genes < 1) return fail(GRSR_ERR_INVALID, "num_genes must be positive");
  int rc = check_genes_fit(num_genes);
  if (rc) return rc;
  int dev0;
  CK(cudaGetDevice(&dev0));
  DevInfo di0;
  if ((rc = dev_info(dev0, di0))) return rc;
  // Two tiers when there are enough pairs to fill the chip with one warp per pair: the walk kernel
  // settles every pair without hurdle candidates, the block-per-pair kernel takes the deferred rest.
  const int walk_mode = env_int("GRSR_WALK", -1);
  if (small_fits(num_genes, false) && walk_mode != 0 && (walk_mode == 1 || npairs >= di0.sms) &&
      dev0 < kMaxDevices && npairs < (1ll << 31)) {
    int order_mode = env_int("GRSR_WALK_SHARED", -1);
    if (order_mode < 0) order_mode = order_hint;
    WalkPlan WP, WS;
    if (order_mode != 1 && (rc = plan_walk(grsr::dist_walk_kernel, false, num_genes, npairs, WP))) return rc;
    if (order_mode != 0 && (rc = plan_walk(grsr::dist_walk_shared_kernel, true, num_genes, npairs, WS))) return rc;
    Launch L;
    grsr::pick_shape(num_genes, env_int("GRSR_THREADS", 0), &L.threads, &L.ept);
    L.smem = grsr::work_bytes(num_genes, L.threads * L.ept, false) + grsr::stage_bytes(num_genes);
    GRSR_DISPATCH_SHAPE(L.threads, L.ept, rc = finish_plan(grsr::dist_deferred_kernel<kNt, kEpt>, npairs, L));
    if (rc) return rc;
    // Per-call scratch (deferred count, order verdict, tickets, deferred list) from the device's
    // stream-ordered pool: calls in flight on different streams of one device never share it.
    if ((rc = pool_ready(dev0))) return rc;
    void *dq;
    {
      cudaError_t e = cudaMallocAsync(&dq, 256 + (size_t)npairs * sizeof(int32_t), st);
      if (e != cudaSuccess)
        return fail(GRSR_ERR_NOMEM, "cudaMallocAsync(%zu) failed: %s", 256 + (size_t)npairs * sizeof(int32_t),
                    cudaGetErrorString(e));
    }
    unsigned *dcount = (unsigned *)dq;
    unsigned *dflag = (order_mode < 0) ? (unsigned *)((char *)dq + 64) : nullptr;
    unsigned *dtick_s = (unsigned *)((char *)dq + 128), *dtick_p = (unsigned *)((char *)dq + 192);
    int32_t *dlist = (int32_t *)((char *)dq + 256);
    CK(cudaMemsetAsync(dq, 0, 256, st));               // deferred count, pi-row change count, tickets
    if (dflag) {
      long long ob = (npairs + 255) / 256;
      grsr::walk_order_kernel<<<(int)(ob < 592 ? ob : 592), 256, 0, st>>>(d_pairs, (long long)npairs, dflag);
    }
    if (order_mode != 0)
      grsr::dist_walk_shared_kernel<<<WS.grid, WS.wpb * 32, WS.smem, st>>>(
          d_genomes, d_sinv, num_genes, d_pairs, (long long)npairs, d_out, stride, WS.shift, dlist, dcount, dflag, dtick_s);
    if (order_mode != 1)
      grsr::dist_walk_kernel<<<WP.grid, WP.wpb * 32, WP.smem, st>>>(
          d_genomes, d_sinv, num_genes, d_pairs, (long long)npairs, d_out, stride, WP.shift, dlist, dcount, dflag, dtick_p);
    CK(cudaGetLastError());
    GRSR_DISPATCH_SHAPE(L.threads, L.ept, (grsr::dist_deferred_kernel<kNt, kEpt><<<L.grid, L.threads, L.smem, st>>>(
        d_genomes, d_sinv, num_genes, d_pairs, dlist, dcount, d_out, stride)));
    CK(cudaGetLastError());
    CK(cudaFreeAsync(dq, st));
  } else if (small_fits(num_genes, false)) {
    Launch L;
    grsr::pick_shape(num_genes, env_int("GRSR_THREADS", 0), &L.threads, &L.ept, 8);
    L.smem = grsr::work_bytes(num_genes, L.threads * L.ept, false) + grsr::stage_bytes(num_genes);
    GRSR_DISPATCH_SHAPE(L.threads, L.ept, rc = finish_plan(grsr::dist_pairs_kernel<kNt, kEpt>, npairs, L));
    if (!rc && L.grid < npairs) {            // more pairs than resident blocks: throughput shape
      grsr::pick_shape(num_genes, env_int("GRSR_THREADS", 0), &L.threads, &L.ept);
      L.smem = grsr::work_bytes(num_genes, L.threads * L.ept, false) + grsr::stage_bytes(num_genes);
      GRSR_DISPATCH_SHAPE(L.threads, L.ept, rc = finish_plan(grsr::dist_pairs_kernel<kNt, kEpt>, npairs, L));
    }
    if (rc) return rc;
    GRSR_DISPATCH_SHAPE(L.threads, L.ept, (grsr::dist_pairs_kernel<kNt, kEpt><<<L.grid, L.threads, L.smem, st>>>(
        d_genomes, d_sinv, num_genes, d_pairs, npairs, d_out, stride)));
  } else {
    LargePlan LP;
    rc = plan_large(grsr::dist_pairs_large_kernel, grsr::dist_pairs_cluster_kernel, npairs,
                    grsr::large_work_bytes(num_genes, false), dev0, LP);
    if (rc) return rc;
    if (LP.cs == 1) {
      grsr::dist_pairs_large_kernel<<<LP.grid, 1024, 0, st>>>(
          d_genomes, d_sinv, num_genes, d_pairs, npairs, d_out, stride, (unsigned char *)LP.scratch);
    } else {
      rc = launch_cluster(grsr::dist_pairs_cluster_kernel, LP, st, d_genomes, d_sinv,
                          num_genes, d_pairs, (long long)npairs, d_out, stride,
                          (unsigned char *)LP.scratch, (unsigned long long *)LP.xch);
      if (rc) return rc;
    }
  }
  CK(cudaGetLastError());
  return GRSR_OK;
}

} // namespace

extern "C" {

int grsr_abi_version(void) { return 1; }

const char *grsr_last_error(void) { return g_err.c_str(); }

int grsr_device_count(void) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess) { cudaGetLastError(); return 0; }
  return cnt;
}

void grsr_release(void) {
  {
    std::lock_guard<std::mutex> lk(g_arena.mu);
    if (g_arena.p) { cudaFreeHost(g_arena.p); g_arena.p = nullptr; g_arena.cap = 0; }
  }
  DeviceGuard guard;
  for (int d = 0; d < kMaxDevices; d++) {
    DevCache &c = g_cache[d];
    std::lock_guard<std::recursive_mutex> lk(c.mu);
    bool any = c.stream != nullptr;
    for (int k = 0; k < kCacheSlots; k++) any = any || c.buf[k];
    if (!any) continue;
    if (cudaSetDevice(d) != cudaSuccess) { cudaGetLastError(); continue; }
    for (int k = 0; k < kCacheSlots; k++) { if (c.buf[k]) cudaFree(c.buf[k]); c.buf[k] = nullptr; c.cap[k] = 0; }
    if (c.stream) { cudaStreamDestroy(c.stream); c.stream = nullptr; }
  }
}

int grsr_max_genes(void) { return kMaxLargeGenes; }

int grsr_max_genes_on_chip(int scenario) {
  int n = grsr::kMaxSmemGenes;
  while (n > 1 && !small_fits(n, scenario != 0)) n--;
  return n;
}

int grsr_signed_inverse_dev(const int32_t *d_genomes, int num_genomes, int num_genes,
                            int32_t *d_sinv, void *stream) {
  long long total = (long long)num_genomes * num_genes;
  if (total <= 0) return fail(GRSR_ERR_INVALID, "empty genome table");
  int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  grsr::signed_inverse_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_genomes, d_sinv, total,
                                                                         num_genes);
  CK(cudaGetLastError());
  return GRSR_OK;
}

int grsr_triangle_pairs_dev(int num_genomes, int64_t first, int64_t count, int32_t *d_pairs,
                            void *stream) {
  if (count <= 0) return GRSR_OK;
  int blocks = (int)((count + 255) / 256 < 1184 ? (count + 255) / 256 : 1184);
  triangle_pairs_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(num_genomes, first, count, d_pairs);
  CK(cudaGetLastError());
  return GRSR_OK;
}

int grsr_dist_pairs_dev(const int32_t *d_genomes, const int32_t *d_sinv, int num_genes,
                        const int32_t *d_pairs, int64_t npairs, int32_t *d_out, int stride,
                        void *stream) {
  if (stride != 1 && stride != 5) return fail(GRSR_ERR_INVALID, "stride must be 1 or 5");
  return dist_pairs_dev_impl(d_genomes, d_sinv, num_genes, d_pairs, npairs, d_out, stride,
                             (cudaStream_t)stream, -1);
}

int grsr_dist_pairs_sorted_dev(const int32_t *d_genomes, const int32_t *d_sinv, int num_genes,
                               const int32_t *d_pairs, int64_t npairs, int32_t *d_out, int stride,
                               void *stream) {
  if (stride != 1 && stride != 5) return fail(GRSR_ERR_INVALID, "stride must be 1 or 5");
  return dist_pairs_dev_impl(d_genomes, d_sinv, num_genes, d_pairs, npairs, d_out, stride,
                             (cudaStream_t)stream, 1);
}

int grsr_dist_pairs(const int32_t *genomes, int num_genomes, int num_genes, const int32_t *pairs,
                    int64_t npairs, int32_t *out, int stride, int first_device, int ngpus) {
  int rc = check_device_range(first_device, ngpus);
  if (rc) return rc;
  if (stride != 1 && stride != 5) return fail(GRSR_ERR_INVALID, "stride must be 1 or 5");
  if (npairs < 0 || (npairs > 0 && (!pairs || !out))) return fail(GRSR_ERR_INVALID, "null buffer");
  if ((rc = check_table(genomes, num_genomes, num_genes))) return rc;
  for (int64_t p = 0; p < 2 * npairs; p++)
    if (pairs[p] < 0 || pairs[p] >= num_genomes)
      return fail(GRSR_ERR_INVALID, "pair %lld names genome %d of %d", (long long)(p / 2),
                  pairs[p], num_genomes);
  // an empty job still rejects a table that is not made of signed permutations (check_genes)
  if (npairs == 0) return check_genomes(genomes, num_genomes, num_genes);
  DeviceGuard guard;
  return shard_pairs(genomes, num_genomes, num_genes, PAIRS_LIST, pairs, npairs, out, stride,
                     first_device, ngpus);
}

int grsr_dist_matrix_shard(const int32_t *genomes, int num_genomes, int num_genes, int shard,
                           int num_shards, int32_t *out, int64_t *count, int device) {
  int rc = check_device_range(device, 1);
  if (rc) return rc;
  if (num_shards < 1 || shard < 0 || shard >= num_shards || !out || !count)
    return fail(GRSR_ERR_INVALID, "bad shard %d of %d", shard, num_shards);
  if ((rc = check_table(genomes, num_genomes, num_genes))) return rc;
  long long P = (long long)num_genomes * (num_genomes - 1) / 2;
  long long lo = P * shard / num_shards, hi = P * (shard + 1) / num_shards;
  *count = hi - lo;
  if (hi == lo) return check_genomes(genomes, num_genomes, num_genes);
  DeviceGuard guard;
  return run_pairs_on_device(device, genomes, num_genomes, num_genes, PAIRS_TRIANGLE, nullptr, lo,
                             hi - lo, out, 1, nullptr);
}

int grsr_dist_matrix(const int32_t *genomes, int num_genomes, int num_genes, int32_t *distmatrix,
                     int first_device, int ngpus) {
  int rc = check_device_range(first_device, ngpus);
  if (rc) return rc;
  if (!distmatrix) return fail(GRSR_ERR_INVALID, "null matrix");
  if ((rc = check_table(genomes, num_genomes, num_genes))) return rc;
  if (num_genomes == 1) {
    distmatrix[0] = 0;
    return check_genomes(genomes, num_genomes, num_genes);
  }
  DeviceGuard guard;
  return matrix_on_devices(genomes, num_genomes, num_genes, distmatrix, first_device, ngpus);
}

int grsr_stats_matrix(const int32_t *genomes, int num_genomes, int num_genes,
                      grsr_stats_t *statsmatrix, int first_device, int ngpus) {
  int rc = check_device_range(first_device, ngpus);
  if (rc) return rc;
  if (!statsmatrix) return fail(GRSR_ERR_INVALID, "null matrix");
  if ((rc = check_table(genomes, num_genomes, num_genes))) return rc;
  long long total = (long long)num_genomes * num_genomes;
  DeviceGuard guard;
  return shard_pairs(genomes, num_genomes, num_genes, PAIRS_SQUARE, nullptr, total,
                     (int32_t *)statsmatrix, 5, first_device, ngpus);
}

int grsr_trial_distances(const int32_t *src, const int32_t *dst, int num_genes, const int32_t *cands,
                         int64_t ncand, int32_t *out, int device) {
  int rc = check_device_range(device, 1);
  if (rc) return rc;
  if (ncand < 0 || (ncand > 0 && (!cands || !out))) return fail(GRSR_ERR_INVALID, "null buffer");
  if ((rc = check_genomes(src, 1, num_genes)) || (rc = check_genomes(dst, 1, num_genes))) return rc;
  for (int64_t c = 0; c < ncand; c++)
    if (cands[2 * c] < 0 || cands[2 * c] > cands[2 * c + 1] || cands[2 * c + 1] >= num_genes)
      return fail(GRSR_ERR_INVALID, "candidate %lld: bad positions %d..%d", (long long)c, cands[2 * c],
                  cands[2 * c + 1]);
  if (ncand == 0) return GRSR_OK;
  const int n = num_genes;
  Launch L;
  grsr::pick_shape(n, env_int("GRSR_THREADS", 0), &L.threads, &L.ept);
  L.smem = grsr::work_bytes(n, L.threads * L.ept, false) + 2 * grsr::scen_arr_bytes(n);
  if (n > grsr::kMaxSmemGenes || 2 * n + 2 > 32767 || L.threads * L.ept < 2 * n + 2 ||
      L.smem > smem_budget())
    return fail(GRSR_ERR_TOOLARGE, "num_genes=%d: candidate scoring keeps the graph on chip "
                "(at most %d genes)", n, grsr_max_genes_on_chip(0));
  if (device < 0 || device >= kMaxDevices) return fail(GRSR_ERR_NODEVICE, "device %d out of range", device);
  DeviceGuard guard;
  CK(cudaSetDevice(device));
  DevCache &c = g_cache[device];
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  if (!c.stream) CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  cudaStream_t st = c.stream;
  GRSR_DISPATCH_SHAPE(L.threads, L.ept, rc = finish_plan(grsr::trial_batch_kernel<kNt, kEpt>, ncand, L));
  if (rc) return rc;
  void *dsrc, *ddst, *dc, *dout;
  size_t gb = (size_t)n * sizeof(int32_t);
  if ((rc = cache_reserve(c, 0, gb, &dsrc)) || (rc = cache_reserve(c, 1, gb, &ddst)) ||
      (rc = cache_reserve(c, 2, (size_t)ncand * 2 * sizeof(int32_t), &dc)) ||
      (rc = cache_reserve(c, 3, (size_t)ncand * sizeof(int32_t), &dout)))
    return rc;
  CK(cudaMemcpyAsync(dsrc, src, gb, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ddst, dst, gb, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dc, cands, (size_t)ncand * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  GRSR_DISPATCH_SHAPE(L.threads, L.ept, (grsr::trial_batch_kernel<kNt, kEpt><<<L.grid, L.threads, L.smem, st>>>(
      (const int32_t *)dsrc, (const int32_t *)ddst, n, (const int32_t *)dc, (long long)ncand,
      (int32_t *)dout, (const int32_t *)nullptr)));
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, dout, (size_t)ncand * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return GRSR_OK;
}

int grsr_unsigned_dist(const int32_t *g1, const int32_t *g2, int num_genes, int circular,
                       int32_t *signed_g1, int32_t *signed_g2, grsr_unsigned_info_t *info, int device) {
  int rc = check_device_range(device, 1);
  if (rc) return rc;
  if (!info) return fail(GRSR_ERR_INVALID, "null info");
  if ((rc = check_genomes(g1, 1, num_genes)) || (rc = check_genomes(g2, 1, num_genes))) return rc;
  const int n = num_genes;
  Launch L;
  grsr::pick_shape(n, env_int("GRSR_THREADS", 0), &L.threads, &L.ept);
  L.smem = grsr::work_bytes(n, L.threads * L.ept, false) + grsr::unsigned_extra_bytes(n);
  if (n > grsr::kMaxSmemGenes || 2 * n + 2 > 32767 || L.threads * L.ept < 2 * n + 2 || L.smem > smem_budget())
    return fail(GRSR_ERR_TOOLARGE, "num_genes=%d: the unsigned distance keeps the graph on chip", n);
  if (device < 0 || device >= kMaxDevices) return fail(GRSR_ERR_NODEVICE, "device %d out of range", device);
  DeviceGuard guard;
  CK(cudaSetDevice(device));
  DevCache &c = g_cache[device];
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  if (!c.stream) CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  cudaStream_t st = c.stream;
  void *dg1, *dpi, *dsx, *ddb, *dbest;
  const size_t gb = (size_t)n * sizeof(int32_t);
  if ((rc = cache_reserve(c, 0, gb, &dg1)) || (rc = cache_reserve(c, 1, gb, &dpi)) ||
      (rc = cache_reserve(c, 2, gb, &dsx)) || (rc = cache_reserve(c, 3, 2 * gb, &ddb)) ||
      (rc = cache_reserve(c, 6, sizeof(unsigned long long), &dbest)))
    return rc;
  int32_t *demit = (int32_t *)ddb + n;

  // one pass of unsigned_mcdist's loop: all 2^k sign assignments on the device
  auto eval = [&](const int32_t *gamma, grsr::UnsignedPass &P, int k, int &best_d,
                  std::vector<int32_t> &best_pi) -> int {
    const int m = (int)P.doubletons.size();
    std::vector<int32_t> sidx((size_t)n, -1);
    for (size_t j = 0; j < P.singletons.size(); j++) sidx[(size_t)P.singletons[j]] = (int32_t)j;
    const unsigned long long leaves = 1ull << k;
    unsigned long long best = ~0ull;
    CK(cudaMemcpyAsync(dg1, gamma, gb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dpi, P.pi0.data(), gb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dsx, sidx.data(), gb, cudaMemcpyHostToDevice, st));
    if (m) CK(cudaMemcpyAsync(ddb, P.doubletons.data(), (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(dbest, 0xff, sizeof best, st));
    Launch LL = L;
    int r = 0;
    GRSR_DISPATCH_SHAPE(L.threads, L.ept, r = finish_plan(grsr::unsigned_signs_kernel<kNt, kEpt>, (long long)leaves, LL));
    if (r) return r;
    GRSR_DISPATCH_SHAPE(L.threads, L.ept, (grsr::unsigned_signs_kernel<kNt, kEpt><<<LL.grid, LL.threads, LL.smem, st>>>(
        (const int32_t *)dg1, (const int32_t *)dpi, (const int32_t *)dsx, (const int32_t *)ddb, n, k, m, 0ull, leaves,
        (unsigned long long *)dbest, (int32_t *)nullptr)));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&best, dbest, sizeof best, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int d = (int)(best >> 32);
    if (d < best_d) {
      const unsigned long long t = best & 0xFFFFFFFFull;
      GRSR_DISPATCH_SHAPE(L.threads, L.ept, (grsr::unsigned_signs_kernel<kNt, kEpt><<<1, LL.threads, LL.smem, st>>>(
          (const int32_t *)dg1, (const int32_t *)dpi, (const int32_t *)dsx, (const int32_t *)ddb, n, k, m, t, 1ull,
          (unsigned long long *)nullptr, demit)));
      CK(cudaGetLastError());
      best_pi.resize((size_t)n);
      CK(cudaMemcpyAsync(best_pi.data(), demit, gb, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      best_d = d;
    }
    return GRSR_OK;
  };
  grsr::UnsignedResult R;
  if ((rc = grsr::unsigned_distance(g1, g2, n, circular != 0, eval, R))) return rc;
  info->dist = R.dist; info->approx = R.approx; info->passes = R.passes;
  for (int q = 0; q < 2; q++) { info->num_sing[q] = R.num_sing[q]; info->num_doub[q] = R.num_doub[q]; }
  if (signed_g1) memcpy(signed_g1, R.g1.data(), gb);
  if (signed_g2) memcpy(signed_g2, R.g2.data(), gb);
  return GRSR_OK;
}

int grsr_scenario_batch(const int32_t *src, const int32_t *dst, int npairs, int num_genes,
                        grsr_step_t *steps, int32_t *nsteps, int max_steps, int first_device,
                        int ngpus) {
  int rc = check_device_range(first_device, ngpus);
  if (rc) return rc;
  if (npairs < 0 || max_steps < 0 || (npairs > 0 && (!steps || !nsteps)))
    return fail(GRSR_ERR_INVALID, "bad scenario arguments");
  if (npairs == 0) return GRSR_OK;
  if ((rc = check_genomes(src, npairs, num_genes))) return rc;
  if ((rc = check_genomes(dst, npairs, num_genes))) return rc;
  DeviceGuard guard;
  return scenario_call(src, dst, npairs, num_genes, steps, nsteps, max_steps, nullptr, nullptr, false,
                       first_device, ngpus);
}

int grsr_scenario_prefix(const int32_t *src, const int32_t *dst, int npairs, int num_genes,
                         grsr_step_t *steps, int32_t *nsteps, int max_steps, int32_t *complete,
                         int64_t *evals, int first_device, int ngpus) {
  int rc = check_device_range(first_device, ngpus);
  if (rc) return rc;
  if (npairs < 0 || max_steps < 0 || (npairs > 0 && (!steps || !nsteps)))
    return fail(GRSR_ERR_INVALID, "bad scenario arguments");
  if (npairs == 0) return GRSR_OK;
  if ((rc = check_genomes(src, npairs, num_genes))) return rc;
  if ((rc = check_genomes(dst, npairs, num_genes))) return rc;
  DeviceGuard guard;
  return scenario_call(src, dst, npairs, num_genes, steps, nsteps, max_steps, complete, evals, true,
                       first_device, ngpus);
}

#ifdef GRSR_PHASE_PROFILE
/* diagnostic build only (scripts/prof_phases.py): cycles and entry counts per phase id */
int grsr_prof_read(unsigned long long *cycles64, unsigned long long *counts64, int reset) {
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpyFromSymbol(cycles64, grsr::grsr_prof_cycles, 64 * sizeof(unsigned long long)));
  CK(cudaMemcpyFromSymbol(counts64, grsr::grsr_prof_counts, 64 * sizeof(unsigned long long)));
  if (reset) {
    unsigned long long z[64] = {0};
    CK(cudaMemcpyToSymbol(grsr::grsr_prof_cycles, z, sizeof z));
    CK(cudaMemcpyToSymbol(grsr::grsr_prof_counts, z, sizeof z));
  }
  return GRSR_OK;
}
#endif

} // extern "C"
