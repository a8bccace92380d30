// Shared device/host helpers for libsmc (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/smc.h"

namespace smc {

// ----------------------------------------------------------------------------------------------
// error plumbing
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define SMC_CHECK_CUDA(expr)                                                                     \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            smc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return 2;                                                                            \
        }                                                                                        \
    } while (0)

#define SMC_REQUIRE(cond, ...)                                                                   \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            smc::set_error(__VA_ARGS__);                                                         \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

// ----------------------------------------------------------------------------------------------
// geometry constants
// ----------------------------------------------------------------------------------------------
constexpr int kThreads = 256;          // CTA size of every streaming kernel
constexpr int kWarps = kThreads / 32;
// K1 tuning knobs (overridable at build time for experiments; the defaults are the measured best)
#ifndef SMC_K1_QUADS
#define SMC_K1_QUADS 4
#endif
#ifndef SMC_K1_OCC
#define SMC_K1_OCC 4
#endif
#ifndef SMC_K1_PREFETCH
#define SMC_K1_PREFETCH 1
#endif
constexpr int kK1Quads = SMC_K1_QUADS; // quads (4 particles) per thread per K1 block
constexpr int kK1Block = kThreads * 4 * kK1Quads;   // particles per K1 chunk (4096)
constexpr int kK1Occ = SMC_K1_OCC;     // resident K1 CTAs per SM for the light (fp32, nx <= 2) functors
#ifndef SMC_K1_GRID_PER_SM
#define SMC_K1_GRID_PER_SM SMC_K1_OCC
#endif
constexpr int kK1MaxBlocks = 148 * SMC_K1_GRID_PER_SM;  // persistent K1 grid: CTAs loop over chunks (B200: 148 SMs)
constexpr int kMaxNx = 64;             // control block moment shift capacity
constexpr int kFusedNxMax = 4;         // K3 moves the payload itself up to this many dims
constexpr int kSyrkMaxBlocks = 148 * 2;  // grid of the high-dimensional moment kernel
constexpr int kRsumStride = 2048;        // entries per range-total buffer (two buffers, <= 148 x 12 ranges each)

// Fixed-point CDF: "index units" (N * w) with F = 62 - ceil(log2 N) fractional bits (>= 32);
// bit 63 of a published tile word is its "valid" flag.
constexpr unsigned long long kValidBit = 1ull << 63;

// ----------------------------------------------------------------------------------------------
// control block (device resident, first 1 KiB of the workspace)
// ----------------------------------------------------------------------------------------------
struct Ctrl {
    double log_norm;     // m + log(sum exp(logw - m)) of the last weighted step
    double ess;          // 1 / sum w^2 before resampling
    double u;            // systematic-resampling uniform of this step
    double shift[kMaxNx];// moment shift (last posterior mean)
    int32_t cur;         // which ping-pong buffer holds the current particles
    int32_t uniform;     // 1: current weights are exactly uniform (fresh / just resampled)
    int32_t do_resample; // decision of the last finalize
    int32_t frac_bits;   // F of the fixed-point CDF for this population size
    // resampler constants derived once per step from (u, log_norm, N) by the finalize kernel
    double lw_ref;                 // quantisation reference: fixed-point weight = 2^Fa exp(lw - lw_ref)
    unsigned long long uq;         // floor(u 2^F)
    unsigned long long margin;     // near-tie margin in fraction units
    // fp32 mode (see "fp32 CDF" in sir_kernels.cu)
    float kshift;                  // a = rni(2^(lw log2e + kshift)), kshift = Fa - lw_ref log2e
    int32_t rsum_valid;            // the kernel that wrote the log-weights also delivered the range totals
    int32_t rsum_sel;              // which of the two range-total buffers belongs to the current log-weights
};
static_assert(sizeof(Ctrl) <= 1024, "control block must fit its slot");

// Workspace layout (all offsets 256-B aligned), computed identically on host and device.
struct WsLayout {
    size_t ctrl, params, k1_parts, tile_agg, ticket, mom_parts, big_parts, board, total;   // tile_agg: 16-B {aggregate, inclusive} records
    int64_t n_k1_blocks, n_tiles;
};
constexpr int kBoardSms = 512;            // %smid is taken modulo this
constexpr int kBoardPieces = 1024;        // >= CTAs of a fused-step launch
constexpr int kParamSlots = 4;            // host->device staging slots for small matrices
constexpr int kParamSlotDoubles = 8192;   // 64 KiB each (up to 90 x 90)

__host__ __device__ inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// partial record: [0]=max, [1]=sum e, [2]=sum e^2, [3..3+nx) sum e (x-c), then nx(nx+1)/2 second moments
// (moments only when they are accumulated inline, i.e. nx <= kFusedNxMax; the layout reserves the
// wide stride for every nx <= kFusedNxMax and 3 doubles otherwise)
__host__ __device__ inline int partial_stride(int nx) {
    return nx <= kFusedNxMax ? 3 + nx + nx * (nx + 1) / 2 : 3;
}
// post-resample moment record per K3 tile: [0..nx) sum (x-c), then nx(nx+1)/2
__host__ __device__ inline int moment_stride(int nx) {
    return nx <= kFusedNxMax ? nx + nx * (nx + 1) / 2 : 1;
}

// particles per resampling tile (must match K3Geom in sir_kernels.cu: 256 x 16 fp32, 256 x 8 fp64)
__host__ __device__ inline int k3_tile_items(int dtype) { return dtype == SMC_F64 ? 2048 : 4096; }

inline WsLayout ws_layout(int64_t n, int nx, int dtype) {
    WsLayout L;
    L.n_k1_blocks = cdiv64(n, kK1Block);
    if (L.n_k1_blocks < 1) L.n_k1_blocks = 1;
    if (L.n_k1_blocks > kK1MaxBlocks) L.n_k1_blocks = kK1MaxBlocks;
    L.n_tiles = cdiv64(n, k3_tile_items(dtype));
    if (L.n_tiles < 1) L.n_tiles = 1;
    size_t off = 0;
    L.ctrl = off;      off = align256(off + 1024);
    L.params = off;    off = align256(off + size_t(kParamSlots) * kParamSlotDoubles * sizeof(double));
    L.k1_parts = off;  off = align256(off + size_t(L.n_k1_blocks) * partial_stride(nx) * sizeof(double));
    L.tile_agg = off;  off = align256(off + size_t(L.n_tiles) * 16 + 65536);   // 2 x kRsumStride range totals
    L.ticket = off;    off = align256(off + 256);
    L.mom_parts = off; off = align256(off + size_t(L.n_tiles) * moment_stride(nx) * sizeof(double));
    // per-CTA records of the X^T W X moment kernel (nx > kFusedNxMax): (1 + nx + nx^2) doubles each
    L.big_parts = off;
    if (nx > kFusedNxMax) off = align256(off + size_t(kSyrkMaxBlocks) * (1 + nx + size_t(nx) * nx) * sizeof(double));
    // placement board of the fused step (fused_step.cuh): kBoardSms rows of 8 words, kBoardPieces claim flags, one
    // slot counter; zero between launches
    L.board = off;     off = align256(off + size_t(kBoardSms * 8 + kBoardPieces + 8) * sizeof(uint32_t));
    L.total = off;
    return L;
}

// ----------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011); counter = (quad index lo/hi, t, stream), key = seed
// ----------------------------------------------------------------------------------------------
enum : uint32_t { kStreamProcess = 0, kStreamResample = 1, kStreamInit = 2, kStreamJitter = 3 };

__host__ __device__ inline void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
#else
    uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__host__ __device__ inline void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += W0; k1 += W1;
    }
}

__device__ inline void philox_quad(uint64_t seed, uint64_t quad, uint32_t t, uint32_t stream,
                                   uint32_t (&out)[4]) {
    out[0] = (uint32_t)quad; out[1] = (uint32_t)(quad >> 32); out[2] = t; out[3] = stream;
    philox4x32_10(out, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// The ten round keys, expanded once on the host and passed by value in the kernel arguments:
// inside the kernel they are constant-bank operands, so the key schedule costs no instructions.
struct PhiloxKeys { uint32_t k[20]; };
inline PhiloxKeys philox_expand(uint64_t seed) {
    PhiloxKeys K;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        K.k[2 * r] = k0; K.k[2 * r + 1] = k1;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return K;
}
__device__ inline void philox_quad_keys(const PhiloxKeys& K, uint64_t quad, uint32_t t, uint32_t stream,
                                        uint32_t (&out)[4]) {
    out[0] = (uint32_t)quad; out[1] = (uint32_t)(quad >> 32); out[2] = t; out[3] = stream;
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(out, K.k[2 * r], K.k[2 * r + 1]);
}

// single-MUFU approximations (flush-to-zero: no denormal fix-up code around them)
__device__ inline float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ inline float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ inline float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// four standard normals (fp32 Box-Muller on 32-bit uniforms, curand-style open interval)
__device__ inline void normals4_f32(const uint32_t (&r)[4], float (&z)[4]) {
    const float k2pow32inv = 2.3283064365386963e-10f;
    const float k2pi = 6.2831853071795865f;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        float u1 = fmaf((float)r[2 * p], k2pow32inv, 0.5f * k2pow32inv);
        float u2 = fmaf((float)r[2 * p + 1], k2pow32inv * k2pi, 0.5f * k2pow32inv * k2pi);
        // sqrt(-2 ln u1) = sqrt(-2 ln2 * log2 u1): two MUFU ops
        float rad = sqrt_approx(-1.3862943611198906f * lg2_approx(u1));
        float s, c;
        __sincosf(u2, &s, &c);
        z[2 * p] = rad * c;
        z[2 * p + 1] = rad * s;
    }
}

// two standard normals in fp64 from one Philox block (53-bit uniforms)
__device__ inline void normals2_f64(const uint32_t (&r)[4], double (&z)[2]) {
    const double k2pow53inv = 1.1102230246251565e-16;
    uint64_t a = ((uint64_t)r[0] << 21) ^ (uint64_t)(r[1] >> 11);
    uint64_t b = ((uint64_t)r[2] << 21) ^ (uint64_t)(r[3] >> 11);
    double u1 = ((double)a + 0.5) * k2pow53inv;
    double u2 = ((double)b + 0.5) * k2pow53inv;
    double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z[0] = rad * c;
    z[1] = rad * s;
}

// uniform in [0,1) with 53 bits from a Philox block
__host__ __device__ inline double uniform53(uint32_t a, uint32_t b) {
    uint64_t v = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11);
    return (double)v * 1.1102230246251565e-16;
}

// ----------------------------------------------------------------------------------------------
// warp / block reductions with a fixed tree (deterministic)
// ----------------------------------------------------------------------------------------------
template <typename T> __device__ inline T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T> __device__ inline T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// CTA-wide (256 threads) reductions through a kWarps-slot shared scratch; every thread gets the result
__device__ inline double block_sum_256(double v, double* s_tmp) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_tmp[warp] = v;
    __syncthreads();
    double r = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) r += s_tmp[w];
    return r;
}
__device__ inline double block_max_256(double v, double* s_tmp) {
    v = warp_max(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_tmp[warp] = v;
    __syncthreads();
    double r = s_tmp[0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) r = s_tmp[w] > r ? s_tmp[w] : r;
    return r;
}

// vector type helper
template <typename real> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

template <typename real> __device__ inline void load4(const real* p, real (&v)[4]);
template <> __device__ inline void load4<float>(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ inline void load4<double>(const double* p, double (&v)[4]) {
    double2 a = *reinterpret_cast<const double2*>(p);
    double2 b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename real> __device__ inline void store4(real* p, const real (&v)[4]);
template <> __device__ inline void store4<float>(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ inline void store4<double>(double* p, const double (&v)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}

__device__ inline float fast_exp(float x) { return ex2_approx(x * 1.4426950408889634f); }
__device__ inline double fast_exp(double x) { return exp(x); }

}  // namespace smc
