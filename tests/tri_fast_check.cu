// TEST INFRASTRUCTURE (GPU box): interTriangleFast (rt_device.cuh) moves the two sign rejections of the
// reference's ray/triangle test (A10/code.cl:250-288) in front of the IEEE division.  That is only the same
// decision when `numerator * (1 / div)` cannot underflow to -0 (which `beta < 0` does not reject), so the early
// exit is guarded (|numerator| >= 2^-126, div <= 2^22).  This program compares the fast form with the reference
// order (interTrianglePre: same operations as interTriangle<true>, div supplied) on
//   class 0: ordinary geometry, div = dot(cross(e2, e1), d)                 (also checks interTriangle<true> itself)
//   class 1: numerators in the subnormal / tiny range against div in [2^20, 2^120] and +inf
//   class 2: tiny numerators against ordinary div
// and prints the number of decision or bit mismatches.  Compile with the library's arithmetic flags.
#include <cstdio>
#include <cuda_runtime.h>

#include "rt_device.cuh"

using namespace rt;

__device__ unsigned hash32(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return (unsigned)x;
}
// signed float with exponent drawn from [elo, ehi] (biased by 127 by the caller's choice of range) and random mantissa
__device__ float rnd(unsigned long long& st, int elo, int ehi) {
    unsigned h = hash32(st++);
    int e = elo + (int)(h % (unsigned)(ehi - elo + 1));
    unsigned mant = hash32(st++) & 0x7FFFFFu;
    unsigned sign = (hash32(st++) & 1u) << 31;
    if (e < -126) {   // subnormal: shift the mantissa down
        int sh = -126 - e;
        unsigned m = (0x800000u | mant) >> (sh > 24 ? 24 : sh);
        return __uint_as_float(sign | m);
    }
    return __uint_as_float(sign | ((unsigned)(e + 127) << 23) | mant);
}

__global__ void check(unsigned long long n, unsigned long long* bad, unsigned long long* accepted, unsigned long long* early_zero) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long local = 0, acc = 0, ez = 0;
    for (; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long st = i * 64 + 12345;
        const int cls = (int)(i % 3);
        f3 o, d, p0, e1, e2;
        float div;
        d = mk3(rnd(st, -4, 0), rnd(st, -4, 0), rnd(st, -4, 0));
        if (cls == 0) {
            o = mk3(rnd(st, -3, 3), rnd(st, -3, 3), rnd(st, -3, 3));
            p0 = mk3(rnd(st, -3, 3), rnd(st, -3, 3), rnd(st, -3, 3));
            e1 = mk3(rnd(st, -6, 2), rnd(st, -6, 2), rnd(st, -6, 2));
            e2 = mk3(rnd(st, -6, 2), rnd(st, -6, 2), rnd(st, -6, 2));
            div = dot(cross(e2, e1), d);
        } else {
            // s = o - p0 tiny: o = p0 + tiny is not representable next to an ordinary p0, so put p0 at the origin
            p0 = mk3(0.f, 0.f, 0.f);
            o = mk3(rnd(st, -149, -110), rnd(st, -149, -110), rnd(st, -149, -110));
            e1 = mk3(rnd(st, -10, 10), rnd(st, -10, 10), rnd(st, -10, 10));
            e2 = mk3(rnd(st, -10, 10), rnd(st, -10, 10), rnd(st, -10, 10));
            if (cls == 1) {
                unsigned h = hash32(st++);
                div = (h % 17u == 0) ? RT_INF : fabsf(rnd(st, 20, 120));
            } else {
                div = fabsf(rnd(st, -20, 20));
            }
        }
        float b0 = 0.f, g0 = 0.f, t0 = 0.f, b1 = 0.f, g1 = 0.f, t1 = 0.f;
        const float mint = 0.0f, maxt = RT_INF;
        bool v0 = interTrianglePre(o, d, mint, maxt, div, p0, e1, e2, b0, g0, t0);
        bool v1 = interTriangleFast(o, d, mint, maxt, div, p0, e1, e2, b1, g1, t1);
        bool same = v0 == v1 && (!v0 || (__float_as_uint(b0) == __float_as_uint(b1) && __float_as_uint(g0) == __float_as_uint(g1) &&
                                         __float_as_uint(t0) == __float_as_uint(t1)));
        if (cls == 0) {   // the precomputed form against the reference text's own form
            float b2 = 0.f, g2 = 0.f, t2 = 0.f;
            bool v2 = interTriangle<true>(o, d, mint, maxt, p0, p0 + e1, p0 + e2, b2, g2, t2);
            // p0 + e1 - p0 need not give e1 back: only compare when it does
            f3 r1 = (p0 + e1) - p0, r2 = (p0 + e2) - p0;
            bool exact = r1.x == e1.x && r1.y == e1.y && r1.z == e1.z && r2.x == e2.x && r2.y == e2.y && r2.z == e2.z;
            if (exact && (v2 != v0 || (v2 && (__float_as_uint(b2) != __float_as_uint(b0) || __float_as_uint(t2) != __float_as_uint(t0))))) same = false;
        }
        if (!same) local++;
        if (v0) acc++;
        // how often the corner is live: a negative numerator whose quotient is -0 (the reference does not reject on it)
        if (div > 0) {
            f3 s = o - p0;
            float nb = dot(cross(s, d), e2);
            float idiv = 1.0f / div;
            if (nb < 0.0f && nb * idiv == 0.0f) ez++;
        }
    }
    if (local) atomicAdd(bad, local);
    if (acc) atomicAdd(accepted, acc);
    if (ez) atomicAdd(early_zero, ez);
}

int main() {
    unsigned long long* c;
    if (cudaMallocManaged(&c, 3 * sizeof *c) != cudaSuccess) { printf("no device\n"); return 2; }
    c[0] = c[1] = c[2] = 0;
    const unsigned long long n = 3ull << 26;
    check<<<148 * 16, 256>>>(n, c, c + 1, c + 2);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 3; }
    printf("cases %llu accepted %llu underflow_corner %llu mismatches %llu\n", n, c[1], c[2], c[0]);
    return c[0] ? 1 : 0;
}
