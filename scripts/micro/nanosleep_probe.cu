// How long does __nanosleep(t) really take?  (development aid for the event kernel's idle paths)
#include <cstdio>
__global__ void probe(unsigned ns, unsigned long long* out, int reps) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  long long c0 = clock64();
  for (int i = 0; i < reps; i++) __nanosleep(ns);
  long long c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (unsigned long long)(c1 - c0); }
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 16);
  unsigned long long h[2];
  for (int blocks : {1, 148}) for (int threads : {32, 512}) for (unsigned ns : {0u, 20u, 40u, 100u, 200u, 500u, 1000u, 4000u}) {
    const int reps = 200;
    probe<<<blocks, threads>>>(ns, d, reps); cudaDeviceSynchronize();
    probe<<<blocks, threads>>>(ns, d, reps); cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("blocks %3d threads %3d nanosleep(%4u): %.0f ns per call (globaltimer), %.0f cycles per call\n", blocks, threads, ns, (double)h[0] / reps, (double)h[1] / reps);
  }
  return 0;
}
