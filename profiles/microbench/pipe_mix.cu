// micro-benchmark: can the FMA pipe (IMAD / IMAD.WIDE) take the shift work while the ALU pipe does LOP3?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed, uint32_t mul) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (i + 1) + threadIdx.x;
    uint32_t y = seed ^ blockIdx.x, z = seed + 0x9E3779B9u;
#pragma unroll 1
    for (int it = 0; it < 2048; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = __funnelshift_r(a[i], y, 7) ^ (z & a[(i + 1) & 7]);             // SHF + LOP3
                if (MODE == 1) { unsigned long long w = (unsigned long long)a[i] * mul; a[i] = (uint32_t)(w >> 32) ^ (uint32_t)w ^ a[(i+1)&7]; } // IMAD.WIDE + LOP3
                if (MODE == 2) a[i] = (a[i] * mul + z) ^ a[(i + 1) & 7];                                // IMAD + LOP3
                if (MODE == 3) a[i] = (a[i] ^ z) & (a[(i + 1) & 7] | y);                                // LOP3 + LOP3
                if (MODE == 4) { unsigned long long w = (unsigned long long)a[i] * mul + z; a[i] = (uint32_t)(w >> 32) + (uint32_t)w; } // IMAD.WIDE + IADD
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    if (s == 0x12345678u) out[0] = s;
}
template <int MODE> void run(const char *name, uint32_t *d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(d, 12345u + rep, 128u + rep * 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    double pairs = 148.0 * 8 * 256 * 2048 * 64;
    printf("%-28s %.3f ms  %.2f T pairs/s  (%.1f pair-lanes/clk/SM at 1.965 GHz)\n", name, best, pairs / best / 1e9, pairs / (best * 1e-3) / 148 / 1.965e9);
}
int main() {
    uint32_t *d; cudaMalloc(&d, 4);
    run<0>("SHF + LOP3", d); run<1>("IMAD.WIDE + LOP3", d); run<2>("IMAD + LOP3", d); run<3>("LOP3 + LOP3", d); run<4>("IMAD.WIDE + IADD", d);
    return 0;
}
