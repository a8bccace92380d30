// Micro-benchmark: achievable HBM throughput for the access mixes of the SIR kernels on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rw_mix rw_mix.cu && ./rw_mix
// mixes: copy (1R:1W), K1-like (1R in place + 2W), read-only, write-only; grid-stride persistent vs one-shot.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>   // 0 copy a->b, 1 in-place a + write b (K1-like), 2 read a, 3 write b, 4 read a + read b -> write c (K3-like traffic)
__global__ void __launch_bounds__(256) k(float4* a, float4* b, float4* c, size_t n4, float* sink) {
    float acc = 0.f;
    const size_t chunk = 1024;   // float4 per CTA iteration = 4096 floats
    for (size_t base = (size_t)blockIdx.x * chunk; base < n4; base += (size_t)gridDim.x * chunk) {
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const size_t i = base + q * 256 + threadIdx.x;
            if (MODE == 0 || MODE == 1 || MODE == 2 || MODE == 4) v[q] = a[i]; else v[q] = make_float4(1, 2, 3, 4);
            if (MODE == 4) { float4 w = b[i]; v[q].x += w.x; v[q].y += w.y; v[q].z += w.z; v[q].w += w.w; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const size_t i = base + q * 256 + threadIdx.x;
            v[q].x = v[q].x * 0.95f + 0.1f; v[q].y = v[q].y * 0.95f + 0.1f; v[q].z = v[q].z * 0.95f + 0.1f; v[q].w = v[q].w * 0.95f + 0.1f;
            if (MODE == 0) b[i] = v[q];
            if (MODE == 1) { a[i] = v[q]; b[i] = v[q]; }
            if (MODE == 2) acc += v[q].x + v[q].y + v[q].z + v[q].w;
            if (MODE == 3) b[i] = v[q];
            if (MODE == 4) c[i] = v[q];
        }
    }
    if (MODE == 2 && acc == 123.456f) *sink = acc;
}

template <int MODE> int run(const char* name, float4* a, float4* b, float4* c, size_t n4, float* sink, int grid, double bytes_per_elem) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) k<MODE><<<grid, 256>>>(a, b, c, n4, sink);
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int rep = 0; rep < 10; ++rep) {
        cudaEventRecord(e0); k<MODE><<<grid, 256>>>(a, b, c, n4, sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-34s grid %6d  %.4f ms  %.0f GB/s\n", name, grid, best, bytes_per_elem * n4 * 4 / best / 1e6);
    return 0;
}

int main() {
    const size_t n = size_t(1) << 26, n4 = n / 4;
    float4 *a, *b, *c; float* sink;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&c, n * 4)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4)); CK(cudaMemset(c, 0, n * 4));
    for (int grid : {148 * 4, 148 * 8, 16384}) {
        run<0>("copy 1R:1W", a, b, c, n4, sink, grid, 8);
        run<1>("K1-like 1R (in place) : 2W", a, b, c, n4, sink, grid, 12);
        run<2>("read only", a, b, c, n4, sink, grid, 4);
        run<3>("write only", a, b, c, n4, sink, grid, 4);
        run<4>("K3-like 2R:1W", a, b, c, n4, sink, grid, 12);
    }
    return 0;
}
