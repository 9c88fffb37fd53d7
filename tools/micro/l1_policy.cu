// Microbenchmark: does a global store keep / update / evict the line in the SM's L1?  One CTA, one warp measures.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* buf, long long* out, int n) {
  // warm: make sure TLB etc. are warm by touching the buffer far away
  const int tid = threadIdx.x;
  __shared__ long long res[8][64];
  double acc = 0;
  // case 0: cold load (first touch) then reload (L1 hit expected)
  {
    long long t0 = clock64(); double v = buf[tid * 16]; acc += v; long long t1 = clock64();
    if (acc == 123.456) buf[0] = acc;
    long long t2 = clock64(); double w = buf[tid * 16]; acc += w; long long t3 = clock64();
    if (acc == 123.456) buf[0] = acc;
    res[0][tid] = t1 - t0; res[1][tid] = t3 - t2;
  }
  __syncthreads();
  // case 1: line present in L1 (just loaded), another thread stores to it, then we load it again
  {
    buf[((tid + 1) % 32) * 16] = tid + 1.0;   // store into a neighbour's line (present in L1)
    __syncthreads();
    long long t0 = clock64(); double v = buf[tid * 16]; acc += v; long long t1 = clock64();
    if (acc == 123.456) buf[0] = acc;
    res[2][tid] = t1 - t0;
    long long t2 = clock64(); double w = buf[tid * 16]; acc += w; long long t3 = clock64();
    if (acc == 123.456) buf[0] = acc;
    res[3][tid] = t3 - t2;
  }
  __syncthreads();
  // case 2: store to a cold line (never loaded), then another thread loads it
  {
    double* cold = buf + 4096;
    cold[((tid + 1) % 32) * 16] = tid + 2.0;
    __syncthreads();
    long long t0 = clock64(); double v = cold[tid * 16]; acc += v; long long t1 = clock64();
    if (acc == 123.456) buf[0] = acc;
    res[4][tid] = t1 - t0;
    long long t2 = clock64(); double w = cold[tid * 16]; acc += w; long long t3 = clock64();
    if (acc == 123.456) buf[0] = acc;
    res[5][tid] = t3 - t2;
  }
  __syncthreads();
  // case 3: same thread stores then loads its own cold line
  {
    double* cold = buf + 8192;
    cold[tid * 16] = tid + 3.0;
    long long t0 = clock64(); double v = cold[tid * 16]; acc += v; long long t1 = clock64();
    if (acc == 123.456) buf[0] = acc;
    res[6][tid] = t1 - t0;
    long long t2 = clock64(); double w = cold[tid * 16 + 1]; acc += w; long long t3 = clock64();  // same sector, other word
    if (acc == 123.456) buf[0] = acc;
    res[7][tid] = t3 - t2;
  }
  __syncthreads();
  if (tid == 0) for (int c = 0; c < 8; ++c) out[c] = res[c][5];
  if (acc == 1.2345) out[9] = 1;
}
int main() {
  double* buf; long long* out;
  cudaMalloc(&buf, 1 << 20); cudaMemset(buf, 0, 1 << 20);
  cudaMalloc(&out, 128);
  for (int rep = 0; rep < 2; ++rep) {
    k<<<1, 32>>>(buf, out, 0);
    long long h[8]; cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
    printf("rep %d: cold load %lld | reload %lld || after neighbour store into present line: load %lld reload %lld || cold line stored by other: load %lld reload %lld || own store then load %lld, same sector other word %lld\n",
           rep, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
  }
  return 0;
}
