// Dev lab: does a warp whose halves waited on each other for a long time run converged afterwards?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned *out, const int *chase, int n_iter, int mode)
{ const int lane = threadIdx.x & 31, half = lane >> 4;
  unsigned m[8];
  int x = lane;
  m[0] = __activemask();
  if (half == 0)
    { if (mode == 3) { if ((lane & 15) == 0) for (int i = 0; i < n_iter; i++) x = chase[x]; }
      else           for (int i = 0; i < n_iter; i++) x = chase[x];
    }
  m[1] = __activemask();
  __syncwarp();
  m[2] = __activemask();
  x += __shfl_sync(0xffffffffu, x, 0);
  m[3] = __activemask();
  if (mode == 1) asm volatile("bar.sync 0;");
  m[4] = __activemask();
  // a second, short divergent region
  if (half == 1) x = chase[x & 1023];
  m[5] = __activemask();
  x += __shfl_sync(0xffffffffu, x, 1);
  m[6] = __activemask();
  if (lane == 0 || lane == 16)
    for (int i = 0; i < 7; i++) out[(blockIdx.x * 2 + half) * 8 + i] = m[i];
  if (x == -12345) out[0] = x;
}
int main()
{ int *chase; unsigned *out; const int N = 1 << 20;
  cudaMalloc(&chase, N * 4); cudaMalloc(&out, 4096);
  int *h = new int[N]; for (int i = 0; i < N; i++) h[i] = (int) (((long long) i * 7919 + 13) % N);
  cudaMemcpy(chase, h, N * 4, cudaMemcpyHostToDevice);
  for (int mode = 0; mode <= 3; mode += 1)
    for (int n = 1; n <= 100000; n *= 10)
      { cudaMemset(out, 0, 4096);
        k<<<1, 32>>>(out, chase, n, mode);
        unsigned r[16]; cudaMemcpy(r, out, 64, cudaMemcpyDeviceToHost);
        printf("mode %d n %6d | half0:", mode, n); for (int i = 0; i < 7; i++) printf(" %08x", r[i]);
        printf(" | half1:"); for (int i = 0; i < 7; i++) printf(" %08x", r[8 + i]); printf("\n");
      }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
