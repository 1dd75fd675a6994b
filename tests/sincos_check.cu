// TEST INFRASTRUCTURE (GPU box): the concentric map takes cos and sin of one angle (A10/code.cl:167-168); the CUDA
// path gets both from ONE sincos() call (rt_device.cuh, RT_SINCOS).  This program checks, for EVERY fp32 value of
// the angle's range [-pi/4, 3pi/4] (and a margin: |x| <= 4), that sincos() returns exactly the doubles sin() and
// cos() return -- so the fp32 results of the two forms are the same bits.  Prints the number of mismatches.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void check(unsigned lo, unsigned hi, unsigned long long* bad) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long n = (unsigned long long)hi - lo + 1;
    unsigned long long local = 0;
    for (; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        for (int sign = 0; sign < 2; sign++) {
            float x = __uint_as_float((unsigned)(lo + i) | (sign ? 0x80000000u : 0u));
            double s, c;
            sincos((double)x, &s, &c);
            double s1 = sin((double)x), c1 = cos((double)x);
            if (__double_as_longlong(s) != __double_as_longlong(s1) || __double_as_longlong(c) != __double_as_longlong(c1)) local++;
        }
    }
    if (local) atomicAdd(bad, local);
}

int main() {
    unsigned long long* bad;
    if (cudaMallocManaged(&bad, sizeof *bad) != cudaSuccess) { printf("no device\n"); return 2; }
    *bad = 0;
    const unsigned hi = 0x40800000u;   // 4.0f: all non-negative floats up to it, both signs
    check<<<148 * 16, 256>>>(0u, hi, bad);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 3; }
    printf("values %llu mismatches %llu\n", 2ull * ((unsigned long long)hi + 1), *bad);
    return *bad ? 1 : 0;
}
