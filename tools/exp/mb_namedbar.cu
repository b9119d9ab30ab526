// mb_namedbar.cu -- what compute-sanitizer's synccheck accepts for named barriers used by a subset of a CTA's warps.
//   mode 0: warps 0,1 bar.sync 1,64 from two different code locations; warp 2 does not take part
//   mode 1: warp 0 bar.sync 1,64 / warp 1 bar.arrive 1,64 (producer/consumer form)
//   mode 2: warps 0,1 bar.sync 1,64 from the SAME code location; warp 2 does not take part
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/bin/mb_namedbar tools/exp/mb_namedbar.cu
#include <cstdio>
#include <cstdlib>
__global__ void k(int mode, int *out)
{
    const int w = threadIdx.x >> 5;
    __shared__ int s[96];
    s[threadIdx.x] = threadIdx.x;
    __syncthreads();
    int v = 0;
    if (mode == 0) {
        if (w == 0) {
            s[threadIdx.x] = 7;
            asm volatile("bar.sync 1, 64;" ::: "memory");
            v = s[threadIdx.x + 32];
        } else if (w == 1) {
            s[threadIdx.x] = 9;
            asm volatile("bar.sync 1, 64;" ::: "memory");
            v = s[threadIdx.x - 32] + 1;
        }
    } else if (mode == 1) {
        if (w == 0) {
            asm volatile("bar.sync 1, 64;" ::: "memory");
            v = s[threadIdx.x + 32];
        } else if (w == 1) {
            s[threadIdx.x] = 9;
            asm volatile("bar.arrive 1, 64;" ::: "memory");
        }
    } else {
        if (w < 2) {
            s[threadIdx.x] = 7 + w;
            asm volatile("bar.sync 1, 64;" ::: "memory");
            v = s[threadIdx.x ^ 32];
        }
    }
    __syncthreads();
    out[threadIdx.x] = v;
}
int main(int argc, char **argv)
{
    int *d;
    cudaMalloc(&d, 96 * sizeof(int));
    k<<<1, 96>>>(atoi(argv[1]), d);
    int h[96];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("mode %s: %d %d %d  %s\n", argv[1], h[0], h[32], h[64], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
