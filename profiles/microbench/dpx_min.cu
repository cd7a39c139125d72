#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__global__ void k_min3(uint32_t *out, uint32_t a0, int iters) {
    uint32_t a = a0 + threadIdx.x, b = a0 * 3 + threadIdx.x, c = a0 * 7, d = a0 * 11;
    uint32_t e = a ^ 0x1234, f = b ^ 0x777, g = c ^ 0x999, h = d ^ 0xabc;
    for (int i = 0; i < iters; ++i) {
        a = __vimin3_u16x2(a, b, c); b = __vimin3_u16x2(b, c, d); c = __vimin3_u16x2(c, d, a); d = __vimin3_u16x2(d, a, b);
        e = __vimin3_u16x2(e, f, g); f = __vimin3_u16x2(f, g, h); g = __vimin3_u16x2(g, h, e); h = __vimin3_u16x2(h, e, f);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
__global__ void k_min2(uint32_t *out, uint32_t a0, int iters) {
    uint32_t a = a0 + threadIdx.x, b = a0 * 3 + threadIdx.x, c = a0 * 7, d = a0 * 11;
    uint32_t e = a ^ 0x1234, f = b ^ 0x777, g = c ^ 0x999, h = d ^ 0xabc;
    for (int i = 0; i < iters; ++i) {
        a = __vminu2(a, b); b = __vminu2(b, c); c = __vminu2(c, d); d = __vminu2(d, a);
        e = __vminu2(e, f); f = __vminu2(f, g); g = __vminu2(g, h); h = __vminu2(h, e);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
__global__ void k_lop(uint32_t *out, uint32_t a0, int iters) {
    uint32_t a = a0 + threadIdx.x, b = a0 * 3 + threadIdx.x, c = a0 * 7, d = a0 * 11;
    uint32_t e = a ^ 0x1234, f = b ^ 0x777, g = c ^ 0x999, h = d ^ 0xabc;
    for (int i = 0; i < iters; ++i) {
        a = (a ^ b) | c; b = (b ^ c) | d; c = (c ^ d) | a; d = (d ^ a) | b;
        e = (e ^ f) | g; f = (f ^ g) | h; g = (g ^ h) | e; h = (h ^ e) | f;
        a = __funnelshift_r(a, b, 3); e = __funnelshift_r(e, f, 5);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
template <typename K>
double run(K kern, const char *name, int ops_per_iter) {
    uint32_t *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    kern<<<148 * 8, 256>>>(out, 12345, 100);
    cudaEventRecord(e0);
    kern<<<148 * 8, 256>>>(out, 12345, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)148 * 8 * 256 * iters * ops_per_iter;
    printf("%s: %.3f ms, %.2f Tops/s\n", name, ms, ops / ms / 1e9);
    return ops / ms / 1e9;
}
int main() { run(k_lop, "lop3+shf", 10); run(k_min2, "vminu2", 8); run(k_min3, "vimin3_u16x2", 8); return 0; }
