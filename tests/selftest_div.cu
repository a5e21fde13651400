// selftest_div.cu — GPU self-test of csrc/ddgi_fastmath.cuh (TEST ONLY).
//   1. div_tenth(x) == x / 0.1f for ALL 2^32 float bit patterns (bitwise; NaN == NaN).
//   2. div_markstein(a, d, 1/d) == a / d on `pairs` random operand pairs drawn from the DDA
//      step's ranges: |a| in [2^-100, 1], |d| in [2^-60, 2], all mantissas, both signs.
//   3. rcp_regular(x) == __frcp_rn(x) == 1.0f / x for EVERY float with |x| in [2^-60, 2].
// Prints "tenth_mismatch=N markstein_mismatch=M rcp_mismatch=R pairs=K"; exit code 0 iff all are 0.
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#include "../dynamic-diffuse-global-illumination-minecraft_b200/csrc/ddgi_fastmath.cuh"

using namespace ddgi;

__device__ bool same(float a, float b)
{
    if (a != a && b != b) return true;
    return __float_as_uint(a) == __float_as_uint(b);
}

__global__ void tenth_all(unsigned long long* bad, uint32_t* first)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (uint64_t b = i; b < (1ull << 32); b += stride) {
        float x = __uint_as_float((uint32_t)b);
        float want = x / 0.1f;
        float got = div_tenth(x);
        if (!same(want, got)) {
            local++;
            atomicMin(first, (uint32_t)b);
        }
    }
    if (local) atomicAdd(bad, local);
}

// rcp_regular(x) == 1/x for every float with |x| in [2^-60, 2] (both signs)
__global__ void rcp_all(unsigned long long* bad)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t lo = 0x21800000u, hi = 0x40000000u;  // 2^-60 .. 2.0
    unsigned long long local = 0;
    for (uint64_t b = lo + i; b <= hi; b += stride) {
        for (uint32_t sgn = 0; sgn < 2; sgn++) {
            float x = __uint_as_float((uint32_t)b | (sgn << 31));
            if (!same(__frcp_rn(x), rcp_regular(x)) || !same(1.0f / x, rcp_regular(x))) local++;
        }
    }
    if (local) atomicAdd(bad, local);
}

__device__ uint32_t mix(uint64_t& s)
{
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    uint64_t z = s;
    z ^= z >> 33;
    z *= 0xff51afd7ed558ccdull;
    z ^= z >> 33;
    return (uint32_t)(z >> 16);
}

__global__ void markstein_random(uint64_t per_thread, unsigned long long* bad, float* ex)
{
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t s = 0x9E3779B97F4A7C15ull * (tid + 1);
    unsigned long long local = 0;
    for (uint64_t n = 0; n < per_thread; n++) {
        uint32_t r1 = mix(s), r2 = mix(s), r3 = mix(s);
        // every 4th pair: adversarial significands (all ones / one / near one) for both
        if ((n & 3) == 3) {
            uint32_t pat[4] = {0x7fffffu, 0x000000u, 0x000001u, 0x7ffffeu};
            r1 = (r1 & 0xff800000u) | (pat[(r3 >> 22) & 3] ^ ((r3 >> 26) & 3));
            r2 = (r2 & 0xff800000u) | (pat[(r3 >> 24) & 3] ^ ((r3 >> 28) & 3));
        }
        // a: sign | exponent in [27, 127] (2^-100 .. 1) biased towards [2^-24, 1] | mantissa
        uint32_t ea = (r3 & 7) ? 103 + (r3 >> 3) % 25 : 27 + (r3 >> 3) % 101;
        uint32_t abits = (r1 & 0x807fffffu) | (ea << 23);
        float a = __uint_as_float(abits);
        if (fabsf(a) > 1.0f) a = copysignf(1.0f, a);
        // d: sign | exponent in [67, 127] biased towards [2^-12, 1] | mantissa
        uint32_t ed = (r3 & 0x100) ? 115 + (r3 >> 9) % 13 : 67 + (r3 >> 9) % 61;
        float d = __uint_as_float((r2 & 0x807fffffu) | (ed << 23));
        float inv = 1.0f / d;
        float want = a / d;
        float got = div_markstein(a, d, inv);
        float got2 = div_markstein2(a, d, inv);
        if (!same(want, got) || !same(want, got2)) {
            if (!local) { ex[0] = a; ex[1] = d; }
            local++;
        }
    }
    if (local) atomicAdd(bad, local);
}

int main(int argc, char** argv)
{
    uint64_t log2_pairs = argc > 1 ? strtoull(argv[1], 0, 10) : 34;
    unsigned long long *bad, h[3] = {0, 0, 0};
    uint32_t* first;
    float* ex;
    cudaMalloc(&bad, 24);
    cudaMalloc(&first, 4);
    cudaMalloc(&ex, 8);
    cudaMemset(bad, 0, 24);
    cudaMemset(first, 0xff, 4);
    cudaMemset(ex, 0, 8);
    tenth_all<<<148 * 16, 256>>>(bad, first);
    uint64_t threads = 148ull * 16 * 256;
    uint64_t per_thread = ((1ull << log2_pairs) + threads - 1) / threads;
    markstein_random<<<148 * 16, 256>>>(per_thread, bad + 1, ex);
    rcp_all<<<148 * 16, 256>>>(bad + 2);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 2; }
    uint32_t hf;
    float hex[2];
    cudaMemcpy(h, bad, 24, cudaMemcpyDeviceToHost);
    cudaMemcpy(&hf, first, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hex, ex, 8, cudaMemcpyDeviceToHost);
    printf("tenth_mismatch=%llu markstein_mismatch=%llu rcp_mismatch=%llu pairs=%llu", h[0], h[1], h[2],
           (unsigned long long)(per_thread * threads));
    if (h[0]) printf(" first_tenth_bits=0x%08x", hf);
    if (h[1]) printf(" example a=%a d=%a", hex[0], hex[1]);
    printf("\n");
    return (h[0] || h[1] || h[2]) ? 1 : 0;
}
