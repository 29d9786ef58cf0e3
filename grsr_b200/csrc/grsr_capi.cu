// grsr_capi.cu -- the C ABI declared in include/grsr.h: argument checks, device buffers, copies,
// launches and the multi-GPU pair-list sharder.  All arithmetic of the path happens in the kernels
// of grsr_core.cuh / grsr_scenario.cuh; there is no host implementation of it here.
#include "../../include/grsr.h"
#include "grsr_core.cuh"
#include "grsr_scenario.cuh"
#include "grsr_stepwise.cuh"
#include "grsr_walk.cuh"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
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

struct DevBuf {
  void *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) {
    if (p) { cudaFree(p); p = nullptr; }
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(GRSR_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    return GRSR_OK;
  }
};

// Per-device working set of the host-buffer entry points: one stream and grow-only buffers that
// survive between calls (a distance matrix call is ~20 ms of kernel time; allocating and freeing
// 20 MB around it would cost as much again).  One call at a time per device (mutex).
const int kCacheSlots = 8;   // 0 genomes, 1 inverses, 2 pairs, 3 results, 4-5 large-build scratch, 6 error word,
                             // 7 deferred-pair list of the two-tier distance path
struct DevCache {
  std::recursive_mutex mu;
  cudaStream_t stream = nullptr;
  void *buf[kCacheSlots] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t cap[kCacheSlots] = {0, 0, 0, 0, 0, 0, 0, 0};
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
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                          (int)cudaSharedmemCarveoutMaxShared));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, L.threads, L.smem));
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
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                          (int)cudaSharedmemCarveoutMaxShared));
  // The private-row kernel takes the shape with the most resident warps (few large blocks first: less
  // reserved shared memory).  In the shared-row kernel a block waits for its slowest warp at every run
  // of equal pi rows, so it takes the SMALLEST block that reaches 85 % of the best warp count
  // (measured on config 3: 5 blocks x 4 warps beat 3 x 7 and 1 x 22 by 3 % and 17 %).
  int best_wpb = 0, best_occ = 0;
  const int forced = env_int(shared ? "GRSR_WALK_SHARED_WPB" : "GRSR_WALK_WPB", 0);
  int occ_of[33] = {0};
  for (int k = 0; k < (shared ? 32 : 8); k++) {
    const int wpb = shared ? k + 1 : 8 - k;
    if (forced && wpb != forced) continue;
    const size_t smem = shared ? 16 + grsr::walk_tail_bytes(n) + grsr::walk_private_bytes(n) * wpb
                               : grsr::walk_warp_bytes(n) * wpb;
    if (smem > (size_t)di.smem_optin) continue;
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, wpb * 32, smem));
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
  P.shift = grsr::walk_shape_arg(P.shift, idle);
  P.wpb = best_wpb;
  P.smem = shared ? 16 + grsr::walk_tail_bytes(n) + grsr::walk_private_bytes(n) * best_wpb
                  : grsr::walk_warp_bytes(n) * best_wpb;
  P.warps_per_sm = best_wpb * best_occ;
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
  long long g = (long long)di.sms * best_occ, need = (npairs + best_wpb - 1) / best_wpb;
  if (g > need) g = need;
  if (g < 1) g = 1;
  P.grid = (int)g;
  return GRSR_OK;
}

const int kMaxLargeGenes = 1000000;   // global-memory build: 31-bit ids, 21-bit packed counters

// which build serves num_genes = n: the shared-memory one while the pair's graph fits on chip,
// the global-memory one beyond (or when GRSR_FORCE_LARGE is set, for tests)
bool small_fits(int n, bool scenario) {
  if (env_int("GRSR_FORCE_LARGE", 0)) return false;
  if (n > grsr::kMaxSmemGenes || 2 * n + 2 > 32767) return false;
  int t, e;
  grsr::pick_shape(n, env_int("GRSR_THREADS", 0), &t, &e);
  if (t * e < 2 * n + 2) return false;
  size_t smem = scenario ? grsr::work_bytes(n, t * e, true) + grsr::scen_extra_bytes(n)
                         : grsr::work_bytes(n, t * e, false) + grsr::stage_bytes(n);
  return smem <= 227u * 1024u;
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

__global__ void positions_to_signed_inverse(const int32_t *__restrict__ genomes, int32_t *slot,
                                            long long total, int n) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    long long row = e / n;
    int pos = slot[e];
    slot[e] = genomes[row * n + pos] > 0 ? (pos + 1) : -(pos + 1);
  }
}

// Validates the resident genome table and leaves its signed inverses in d_sinv.  One 8-byte
// device-to-host read; on an offence the message uses the reference's wording.
int validate_and_invert(const int32_t *genomes_host, const int32_t *d_genomes, int G, int n,
                        int32_t *d_sinv, unsigned long long *d_err, cudaStream_t st) {
  long long total = (long long)G * n;
  int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  unsigned long long err = ~0ull;
  CK(cudaMemsetAsync(d_sinv, 0x7f, (size_t)total * sizeof(int32_t), st));
  CK(cudaMemsetAsync(d_err, 0xff, sizeof(unsigned long long), st));
  validate_first_positions<<<blocks, 256, 0, st>>>(d_genomes, d_sinv, total, n, d_err);
  validate_duplicates<<<blocks, 256, 0, st>>>(d_genomes, d_sinv, total, n, d_err);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(&err, d_err, sizeof err, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (err != ~0ull) {
    long long row = (long long)(err >> 34);
    int pos = (int)((err >> 2) & 0xFFFFFFFFull), kind = (int)(err & 3ull);
    long long g = genomes_host[row * n + pos], a = g < 0 ? -g : g;
    if (kind == 0) return fail(GRSR_ERR_INVALID, "genome %lld: input gene == 0", row + 1);
    if (kind == 1)
      return fail(GRSR_ERR_INVALID, "genome %lld: input gene (%lld) larger than number of genes %d",
                  row + 1, g, n);
    return fail(GRSR_ERR_INVALID, "genome %lld: duplicate gene (%lld)", row + 1, a);
  }
  positions_to_signed_inverse<<<blocks, 256, 0, st>>>(d_genomes, d_sinv, total, n);
  CK(cudaGetLastError());
  return GRSR_OK;
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

enum PairSource { PAIRS_LIST, PAIRS_TRIANGLE, PAIRS_SQUARE };

// One device's share of a pair job: copy the genome table in, build the signed inverses, obtain the
// pair slice (copied or generated on the device), run the kernel, copy the results out.
int run_pairs_on_device(int dev, const int32_t *genomes, int G, int n, PairSource src,
                        const int32_t *pairs_host, long long first, long long count,
                        int32_t *out_host, int stride) {
  if (count <= 0) return GRSR_OK;
  if (dev < 0 || dev >= kMaxDevices) return fail(GRSR_ERR_NODEVICE, "device %d out of range", dev);
  CK(cudaSetDevice(dev));
  DevCache &c = g_cache[dev];
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  if (!c.stream) CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  cudaStream_t st = c.stream;
  size_t gbytes = (size_t)G * n * sizeof(int32_t);
  void *dg, *ds, *dp, *dout;
  int rc;
  if ((rc = cache_reserve(c, 0, gbytes, &dg)) || (rc = cache_reserve(c, 1, gbytes, &ds)) ||
      (rc = cache_reserve(c, 2, (size_t)count * 2 * sizeof(int32_t), &dp)) ||
      (rc = cache_reserve(c, 3, (size_t)count * stride * sizeof(int32_t), &dout)))
    return rc;
  void *derr;
  if ((rc = cache_reserve(c, 6, sizeof(unsigned long long), &derr))) return rc;
  CK(cudaMemcpyAsync(dg, genomes, gbytes, cudaMemcpyHostToDevice, st));
  if ((rc = validate_and_invert(genomes, (const int32_t *)dg, G, n, (int32_t *)ds,
                                (unsigned long long *)derr, st)))
    return rc;
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
  if ((rc = grsr_dist_pairs_dev((const int32_t *)dg, (const int32_t *)ds, n, (const int32_t *)dp,
                                count, (int32_t *)dout, stride, st)))
    return rc;
  CK(cudaMemcpyAsync(out_host, dout, (size_t)count * stride * sizeof(int32_t),
                     cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return GRSR_OK;
}

// Split [0, total) into ngpus contiguous slices, one host thread per device; no data-path collective.
int shard_pairs(const int32_t *genomes, int G, int n, PairSource src, const int32_t *pairs_host,
                long long total, int32_t *out_host, int stride, int first_device, int ngpus) {
  if (ngpus == 1)
    return run_pairs_on_device(first_device, genomes, G, n, src, pairs_host, 0, total, out_host,
                               stride);
  std::vector<int> rcs(ngpus, 0);
  std::vector<std::string> errs(ngpus);
  std::vector<std::thread> th;
  for (int d = 0; d < ngpus; d++) {
    long long lo = total * d / ngpus, hi = total * (d + 1) / ngpus;
    th.emplace_back([&, d, lo, hi]() {
      rcs[d] = run_pairs_on_device(first_device + d, genomes, G, n, src, pairs_host, lo, hi - lo,
                                   out_host + lo * stride, stride);
      if (rcs[d]) errs[d] = g_err;
    });
  }
  for (auto &t : th) t.join();
  for (int d = 0; d < ngpus; d++)
    if (rcs[d]) return fail(rcs[d], "device %d: %s", first_device + d, errs[d].c_str());
  return GRSR_OK;
}

// One device's share of a scenario batch: copy the pairs in, run the persistent scenario kernel,
// copy the step lists out.
int scenario_on_device(int dev, const int32_t *src, const int32_t *dst, int npairs, int n,
                       grsr_step_t *steps, int32_t *nsteps, int max_steps) {
  if (npairs <= 0) return GRSR_OK;
  if (dev < 0 || dev >= kMaxDevices) return fail(GRSR_ERR_NODEVICE, "device %d out of range", dev);
  CK(cudaSetDevice(dev));
  // one scenario call at a time per device: the global-memory build's workspace is shared
  std::lock_guard<std::recursive_mutex> device_lock(g_cache[dev].mu);
  DevInfo di;
  int rc = dev_info(dev, di);
  if (rc) return rc;
  if ((rc = check_genes_fit(n))) return rc;
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
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  DevBuf dsrc, ddst, dsteps, dns, dstat, dtick;
  size_t gbytes = (size_t)npairs * n * sizeof(int32_t);
  size_t sbytes = (size_t)npairs * (max_steps > 0 ? max_steps : 1) * sizeof(grsr_step_t);
  if ((rc = dsrc.alloc(gbytes)) || (rc = ddst.alloc(gbytes)) || (rc = dsteps.alloc(sbytes)) ||
      (rc = dns.alloc((size_t)npairs * 4)) || (rc = dstat.alloc((size_t)npairs * 4)) ||
      (rc = dtick.alloc(4))) {
    cudaStreamDestroy(st);
    return rc;
  }
  std::vector<int32_t> stat((size_t)npairs);
  auto body = [&]() -> int {
    CK(cudaMemcpyAsync(dsrc.p, src, gbytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ddst.p, dst, gbytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(dtick.p, 0, 4, st));
    int resume = 0;
    if (small && npairs <= env_int("GRSR_STEPWISE_MAX", 16)) {
      // Few pairs: while a pair's graph has hurdles (every breakpoint pair must be evaluated, hundreds
      // to thousands per step) spread each step's evaluations over the whole GPU; hand the
      // hurdle-free rest of every pair to the persistent kernel below (grsr_stepwise.cuh).
      const int cap = (2 * (n + 1) > 4096) ? 2 * (n + 1) : 4096;
      DevBuf dstate, dcands, dd2;
      int r;
      if ((r = dstate.alloc(sizeof(grsr::StepState))) || (r = dcands.alloc((size_t)cap * 2 * sizeof(int32_t))) ||
          (r = dd2.alloc((size_t)cap * sizeof(int32_t))))
        return r;
      Launch LT;
      LT.threads = L.threads; LT.ept = L.ept;
      LT.smem = grsr::work_bytes(n, L.threads * L.ept, false) + 2 * grsr::scen_arr_bytes(n);
      GRSR_DISPATCH_SHAPE(L.threads, L.ept, r = finish_plan(grsr::trial_batch_kernel<kNt, kEpt>, cap, LT));
      if (r) return r;
      { Launch LPp = L;
        GRSR_DISPATCH_SHAPE(L.threads, L.ept, r = finish_plan(grsr::step_prepare_kernel<kNt, kEpt>, 1, LPp));
        if (r) return r; }
      std::vector<int32_t> ns0((size_t)npairs, 0);
      const int32_t *ncand_dev = (const int32_t *)((const char *)dstate.p + offsetof(grsr::StepState, ncand));
      for (int p = 0; p < npairs; p++) {
        grsr::StepState hs;
        memset(&hs, 0, sizeof hs);
        hs.lastkey = ~0ull; hs.batch_end = ~0ull;
        CK(cudaMemcpyAsync(dstate.p, &hs, sizeof hs, cudaMemcpyHostToDevice, st));
        int32_t *cur_p = (int32_t *)dsrc.p + (size_t)p * n;
        const int32_t *dst_p = (const int32_t *)ddst.p + (size_t)p * n;
        int32_t *steps_p = (int32_t *)dsteps.p + 2 * (size_t)p * max_steps;
        for (;;) {
          GRSR_DISPATCH_SHAPE(L.threads, L.ept, (grsr::step_prepare_kernel<kNt, kEpt><<<1, L.threads, L.smem, st>>>(
              cur_p, dst_p, n, (grsr::StepState *)dstate.p, (int32_t *)dcands.p, cap)));
          GRSR_DISPATCH_SHAPE(L.threads, L.ept, (grsr::trial_batch_kernel<kNt, kEpt><<<LT.grid, LT.threads, LT.smem, st>>>(
              cur_p, dst_p, n, (const int32_t *)dcands.p, 0LL, (int32_t *)dd2.p, ncand_dev)));
          grsr::step_select_kernel<<<1, 256, 0, st>>>(cur_p, n, (grsr::StepState *)dstate.p,
                                                      (const int32_t *)dcands.p, (const int32_t *)dd2.p,
                                                      steps_p, max_steps);
          CK(cudaGetLastError());
          CK(cudaMemcpyAsync(&hs, dstate.p, sizeof hs, cudaMemcpyDeviceToHost, st));
          CK(cudaStreamSynchronize(st));
          if (hs.d == 0 || hs.status != GRSR_SCEN_OK || hs.h == 0) break;
        }
        ns0[p] = hs.nsteps;
        if (hs.status == GRSR_SCEN_OVERFLOW)
          return fail(GRSR_ERR_INVALID, "pair %d: scenario needs more than max_steps=%d steps", p, max_steps);
        if (hs.status == GRSR_SCEN_STUCK)
          return fail(GRSR_ERR_SEARCH, "pair %d: reversals ended prematurely after %d steps", p, hs.nsteps);
      }
      CK(cudaMemcpyAsync(dns.p, ns0.data(), (size_t)npairs * 4, cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));      // ns0 leaves scope
      resume = 1;
    }
    if (small) {
      GRSR_DISPATCH_SHAPE(L.threads, L.ept, (grsr::scenario_kernel<kNt, kEpt><<<L.grid, L.threads, L.smem, st>>>(
          (const int32_t *)dsrc.p, (const int32_t *)ddst.p, npairs, n, (int32_t *)dsteps.p,
          (int32_t *)dns.p, max_steps, (int32_t *)dstat.p, nullptr, (unsigned int *)dtick.p, resume)));
    } else if (LP.cs == 1) {
      grsr::scenario_large_kernel<<<LP.grid, 1024, 0, st>>>(
          (const int32_t *)dsrc.p, (const int32_t *)ddst.p, npairs, n, (int32_t *)dsteps.p,
          (int32_t *)dns.p, max_steps, (int32_t *)dstat.p, nullptr, (unsigned int *)dtick.p,
          (unsigned char *)LP.scratch);
    } else {
      int r = launch_cluster(grsr::scenario_cluster_kernel, LP, st,
                             (const int32_t *)dsrc.p, (const int32_t *)ddst.p, npairs, n,
                             (int32_t *)dsteps.p, (int32_t *)dns.p, max_steps, (int32_t *)dstat.p,
                             (unsigned long long *)nullptr, (unsigned int *)dtick.p,
                             (unsigned char *)LP.scratch, (unsigned long long *)LP.xch);
      if (r) return r;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(steps, dsteps.p, sbytes, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nsteps, dns.p, (size_t)npairs * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(stat.data(), dstat.p, (size_t)npairs * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return GRSR_OK;
  };
  rc = body();
  cudaStreamDestroy(st);
  if (rc) return rc;
  for (int p = 0; p < npairs; p++) {
    if (stat[p] == GRSR_SCEN_OVERFLOW)
      return fail(GRSR_ERR_INVALID, "pair %d: scenario needs more than max_steps=%d steps", p,
                  max_steps);
    if (stat[p] == GRSR_SCEN_STUCK)
      return fail(GRSR_ERR_SEARCH, "pair %d: reversals ended prematurely after %d steps", p,
                  nsteps[p]);
  }
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
  if (npairs <= 0) return GRSR_OK;
  if (stride != 1 && stride != 5) return fail(GRSR_ERR_INVALID, "stride must be 1 or 5");
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
    // order_mode: -1 = decide on the device (walk_order_kernel), 0 = private rows, 1 = shared rows
    const int order_mode = env_int("GRSR_WALK_SHARED", -1);
    WalkPlan WP, WS;
    if (order_mode != 1 && (rc = plan_walk(grsr::dist_walk_kernel, false, num_genes, npairs, WP))) return rc;
    if (order_mode != 0 && (rc = plan_walk(grsr::dist_walk_shared_kernel, true, num_genes, npairs, WS))) return rc;
    Launch L;
    grsr::pick_shape(num_genes, env_int("GRSR_THREADS", 0), &L.threads, &L.ept);
    L.smem = grsr::work_bytes(num_genes, L.threads * L.ept, false) + grsr::stage_bytes(num_genes);
    GRSR_DISPATCH_SHAPE(L.threads, L.ept, rc = finish_plan(grsr::dist_deferred_kernel<kNt, kEpt>, npairs, L));
    if (rc) return rc;
    void *dq;
    {
      DevCache &c = g_cache[dev0];
      std::lock_guard<std::recursive_mutex> lk(c.mu);
      if ((rc = cache_reserve(c, 7, 256 + (size_t)npairs * sizeof(int32_t), &dq))) return rc;
    }
    unsigned *dcount = (unsigned *)dq;
    unsigned *dflag = (order_mode < 0) ? (unsigned *)((char *)dq + 64) : nullptr;
    int32_t *dlist = (int32_t *)((char *)dq + 256);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(dq, 0, 128, st));               // deferred count and pi-row change count
    if (dflag) {
      long long ob = (npairs + 255) / 256;
      grsr::walk_order_kernel<<<(int)(ob < 592 ? ob : 592), 256, 0, st>>>(d_pairs, (long long)npairs, dflag);
    }
    if (order_mode != 0)
      grsr::dist_walk_shared_kernel<<<WS.grid, WS.wpb * 32, WS.smem, st>>>(
          d_genomes, d_sinv, num_genes, d_pairs, (long long)npairs, d_out, stride, WS.shift, dlist, dcount, dflag);
    if (order_mode != 1)
      grsr::dist_walk_kernel<<<WP.grid, WP.wpb * 32, WP.smem, st>>>(
          d_genomes, d_sinv, num_genes, d_pairs, (long long)npairs, d_out, stride, WP.shift, dlist, dcount, dflag);
    CK(cudaGetLastError());
    GRSR_DISPATCH_SHAPE(L.threads, L.ept, (grsr::dist_deferred_kernel<kNt, kEpt><<<L.grid, L.threads, L.smem, (cudaStream_t)stream>>>(
        d_genomes, d_sinv, num_genes, d_pairs, dlist, dcount, d_out, stride)));
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
    GRSR_DISPATCH_SHAPE(L.threads, L.ept, (grsr::dist_pairs_kernel<kNt, kEpt><<<L.grid, L.threads, L.smem, (cudaStream_t)stream>>>(
        d_genomes, d_sinv, num_genes, d_pairs, npairs, d_out, stride)));
  } else {
    int dev;
    CK(cudaGetDevice(&dev));
    LargePlan LP;
    rc = plan_large(grsr::dist_pairs_large_kernel, grsr::dist_pairs_cluster_kernel, npairs,
                    grsr::large_work_bytes(num_genes, false), dev, LP);
    if (rc) return rc;
    if (LP.cs == 1) {
      grsr::dist_pairs_large_kernel<<<LP.grid, 1024, 0, (cudaStream_t)stream>>>(
          d_genomes, d_sinv, num_genes, d_pairs, npairs, d_out, stride, (unsigned char *)LP.scratch);
    } else {
      rc = launch_cluster(grsr::dist_pairs_cluster_kernel, LP, (cudaStream_t)stream, d_genomes, d_sinv,
                          num_genes, d_pairs, (long long)npairs, d_out, stride,
                          (unsigned char *)LP.scratch, (unsigned long long *)LP.xch);
      if (rc) return rc;
    }
  }
  CK(cudaGetLastError());
  return GRSR_OK;
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
  if (npairs == 0) return GRSR_OK;
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
  return run_pairs_on_device(device, genomes, num_genomes, num_genes, PAIRS_TRIANGLE, nullptr, lo,
                             hi - lo, out, 1);
}

int grsr_dist_matrix(const int32_t *genomes, int num_genomes, int num_genes, int32_t *distmatrix,
                     int first_device, int ngpus) {
  int rc = check_device_range(first_device, ngpus);
  if (rc) return rc;
  if (!distmatrix) return fail(GRSR_ERR_INVALID, "null matrix");
  if ((rc = check_table(genomes, num_genomes, num_genes))) return rc;
  long long G = num_genomes, P = G * (G - 1) / 2;
  std::vector<int32_t> flat((size_t)(P > 0 ? P : 1));
  if (P > 0) {
    rc = shard_pairs(genomes, num_genomes, num_genes, PAIRS_TRIANGLE, nullptr, P, flat.data(), 1,
                     first_device, ngpus);
    if (rc) return rc;
  }
  // host gather of matrix rows: distmatrix[i][j] = distmatrix[j][i] (G/uniinvdist.c:84-88)
  long long p = 0;
  for (long long j = 0; j < G; j++) {
    distmatrix[j * G + j] = 0;
    for (long long i = 0; i < j; i++, p++)
      distmatrix[i * G + j] = distmatrix[j * G + i] = flat[p];
  }
  return GRSR_OK;
}

int grsr_stats_matrix(const int32_t *genomes, int num_genomes, int num_genes,
                      grsr_stats_t *statsmatrix, int first_device, int ngpus) {
  int rc = check_device_range(first_device, ngpus);
  if (rc) return rc;
  if (!statsmatrix) return fail(GRSR_ERR_INVALID, "null matrix");
  if ((rc = check_table(genomes, num_genomes, num_genes))) return rc;
  long long total = (long long)num_genomes * num_genomes;
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
      L.smem > 227u * 1024u)
    return fail(GRSR_ERR_TOOLARGE, "num_genes=%d: candidate scoring keeps the graph on chip "
                "(at most %d genes)", n, grsr_max_genes_on_chip(0));
  if (device < 0 || device >= kMaxDevices) return fail(GRSR_ERR_NODEVICE, "device %d out of range", device);
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
  std::vector<int> rcs(ngpus, 0);
  std::vector<std::string> errs(ngpus);
  auto work = [&](int d) {
    long long lo = (long long)npairs * d / ngpus, hi = (long long)npairs * (d + 1) / ngpus;
    rcs[d] = scenario_on_device(first_device + d, src + lo * num_genes, dst + lo * num_genes,
                                      (int)(hi - lo), num_genes, steps + lo * max_steps,
                                      nsteps + lo, max_steps);
    if (rcs[d]) errs[d] = g_err;
  };
  if (ngpus == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int d = 0; d < ngpus; d++) th.emplace_back(work, d);
    for (auto &t : th) t.join();
  }
  for (int d = 0; d < ngpus; d++)
    if (rcs[d]) return fail(rcs[d], "device %d: %s", first_device + d, errs[d].c_str());
  return GRSR_OK;
}

} // extern "C"
