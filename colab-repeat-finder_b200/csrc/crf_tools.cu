// crf_tools.cu -- measurement helpers for bench.py (NOT part of the scan path, not in crf.h):
// an INT32 ALU-pipe peak micro-benchmark, so the integer roofline denominator in the bench line
// is measured on the box it runs on rather than assumed.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

// 8 independent chains per thread; every step is 1 SHF (funnel shift) + 1 LOP3 per chain, the
// same instruction mix the scan kernel's compare words are made of.
template <int STEPS>
__global__ void __launch_bounds__(256) alu_peak_kernel(uint32_t *out, uint32_t seed) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (i + 1) + threadIdx.x;
    uint32_t y = seed ^ blockIdx.x, z = seed + 0x9E3779B9u;
#pragma unroll 1
    for (int it = 0; it < STEPS; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __funnelshift_r(a[i], y, 7) ^ (z & a[(i + 1) & 7]);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    if (s == 0x12345678u) out[0] = s;  // keep the chains alive
}

}  // namespace

// Returns measured INT32 ALU ops/s (LOP3 + SHF counted as one op each) in *ops_per_s.
extern "C" int crf_tools_alu_peak(int device, double *ops_per_s, double *ms_out) {
    if (cudaSetDevice(device) != cudaSuccess) return 1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 1;
    uint32_t *d = nullptr;
    if (cudaMalloc(&d, 4) != cudaSuccess) return 1;
    constexpr int STEPS = 2048;
    const int blocks = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        alu_peak_kernel<STEPS><<<blocks, 256>>>(d, 12345u + rep);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return 1;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double ops = (double)blocks * 256.0 * STEPS * 8.0 * 8.0 * 2.0;
    *ops_per_s = ops / (best * 1e-3);
    if (ms_out) *ms_out = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return 0;
}
