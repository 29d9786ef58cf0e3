// startup_probe.cu -- where a one-shot `grimm` process spends its start-up on this box: CUDA
// initialisation, context creation on one device, first allocation, and loading libgrsr_cuda.so.
// Build: nvcc -O2 -o scripts/bin/startup_probe scripts/startup_probe.cu -ldl   (diagnostic only)
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <time.h>
static double now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(int argc, char **argv) {
  double t0 = now();
  int n = 0; cudaGetDeviceCount(&n);
  double t1 = now();
  cudaSetDevice(0); cudaFree(0);
  double t2 = now();
  void *p; cudaMalloc(&p, 1 << 20); cudaMemset(p, 0, 1 << 20); cudaDeviceSynchronize();
  double t3 = now();
  void *h = argc > 1 ? dlopen(argv[1], RTLD_NOW) : NULL;
  double t4 = now();
  printf("devices %d: cudaGetDeviceCount (driver init) %.1f ms, context on device 0 %.1f ms, first malloc+memset %.1f ms, dlopen(%s) %.1f ms %s\n",
         n, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), argc > 1 ? argv[1] : "-", 1e3 * (t4 - t3), h ? "ok" : "-");
  return 0;
}
