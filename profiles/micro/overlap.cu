// Micro-benchmark: how well does a persistent streaming kernel overlap ALU work with the K1 traffic mix
// (1R in place + 2W, 12 B per element) on B200?  Sweeps ALU instructions per element, CTAs per SM,
// and software prefetch.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o overlap overlap.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int WORK, bool PREFETCH, int Q>
__global__ void __launch_bounds__(256) k(float4* a, float4* b, size_t n4, float c0, float c1) {
    const size_t chunk = 256 * Q;
    float4 v[Q], nx[Q];
    size_t base = (size_t)blockIdx.x * chunk;
    if (PREFETCH && base < n4) {
#pragma unroll
        for (int q = 0; q < Q; ++q) v[q] = a[base + q * 256 + threadIdx.x];
    }
    for (; base < n4; base += (size_t)gridDim.x * chunk) {
        if (PREFETCH) {
            const size_t nb = base + (size_t)gridDim.x * chunk;
            if (nb < n4) {
#pragma unroll
                for (int q = 0; q < Q; ++q) nx[q] = a[nb + q * 256 + threadIdx.x];
            }
        } else {
#pragma unroll
            for (int q = 0; q < Q; ++q) v[q] = a[base + q * 256 + threadIdx.x];
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            float4 w = v[q];
#pragma unroll
            for (int i = 0; i < WORK; ++i) {   // 4 independent FMA chains per float4
                w.x = fmaf(w.x, c0, c1); w.y = fmaf(w.y, c0, c1); w.z = fmaf(w.z, c0, c1); w.w = fmaf(w.w, c0, c1);
            }
            const size_t i = base + q * 256 + threadIdx.x;
            a[i] = w;
            b[i] = make_float4(w.x + 1.f, w.y + 1.f, w.z + 1.f, w.w + 1.f);
        }
        if (PREFETCH) {
#pragma unroll
            for (int q = 0; q < Q; ++q) v[q] = nx[q];
        }
    }
}

template <int WORK, bool PREFETCH, int Q> int run(float4* a, float4* b, size_t n4, int per_sm) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * per_sm;
    for (int i = 0; i < 3; ++i) k<WORK, PREFETCH, Q><<<grid, 256>>>(a, b, n4, 0.999f, 0.001f);
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int rep = 0; rep < 10; ++rep) {
        cudaEventRecord(e0); k<WORK, PREFETCH, Q><<<grid, 256>>>(a, b, n4, 0.999f, 0.001f); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // ALU-only time: WORK fma per element, 128 fma/clk/SM
    const double alu_ms = (double)WORK * n4 * 4 / (148.0 * 128 * 1.965e9) * 1e3;
    printf("work %3d fma/elem  prefetch %d  quads %d  ctas/sm %d : %.4f ms  (%.0f GB/s; pure ALU %.4f ms)\n", WORK, (int)PREFETCH, Q, per_sm, best,
           12.0 * n4 * 4 / best / 1e6, alu_ms);
    return 0;
}

int main() {
    const size_t n = size_t(1) << 26, n4 = n / 4;
    float4 *a, *b;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
    CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4));
    for (int per_sm : {4, 8}) {
        run<0, false, 4>(a, b, n4, per_sm);
        run<10, false, 4>(a, b, n4, per_sm);
        run<20, false, 4>(a, b, n4, per_sm);
        run<40, false, 4>(a, b, n4, per_sm);
        run<40, true, 4>(a, b, n4, per_sm);
        run<80, false, 4>(a, b, n4, per_sm);
        run<80, true, 4>(a, b, n4, per_sm);
        run<40, false, 1>(a, b, n4, per_sm);
        run<40, true, 1>(a, b, n4, per_sm);
    }
    return 0;
}
