// Multinomial resampling on the device.
//
// Reference: ParticleFilter._multinomial_resample (models/particle_filter.py:173-186) is
// rng.choice(N, size=N, p=w), which NumPy evaluates as
//     cdf = cumsum(p); cdf /= cdf[-1]; idx = searchsorted(cdf, rng.random(N), side='right')
// i.e. idx_i = #{ j : cdf_j <= u_i } for N independent uniforms.
//
// Here the CDF is an exact 64-bit fixed-point prefix sum C_j = sum_{l<=j} round(w_l 2^62) (integer
// adds: independent of the summation order, identical on every grid shape) and the comparison
// cdf_j <= u_i is decided exactly as C_j <= floor(u_i * C_{N-1}) with a 128-bit product, so the only
// deviations from the reference are where the reference's own sequential fp64 cumsum rounds a CDF
// value across a uniform (the same "cumulative-weight tie" rule as the systematic resampler).
//
// Five launches: tile totals -> one-CTA scan of the totals -> tile scans (write C) -> guide table ->
// search.  The guide table (cut-point method) holds, for 2^L ~ N/8 equal cells of the CDF's value
// range, how many CDF entries lie below the cell: it is built by coherent (neighbouring threads ->
// neighbouring thresholds) binary searches, stays L2-resident (N/2 bytes), and leaves each of the N
// random-order outputs a search over ~8 adjacent CDF entries instead of log2 N scattered HBM probes.
#include "smc_common.cuh"

#include <string.h>

namespace smc {
namespace {

constexpr int kMnItems = 8;                     // weights per thread
constexpr int kMnTile = kThreads * kMnItems;    // 2048 weights per tile
constexpr int kMnScanThreads = 1024;
constexpr uint32_t kStreamMultinomial = 4;      // Philox stream of the on-chip uniforms

struct MnLayout {
    size_t cdf, tile_off, guide, total;
    int64_t n_tiles;
    int log2_cells;
};
inline MnLayout mn_layout(int64_t n) {
    MnLayout L;
    L.n_tiles = cdiv64(n, kMnTile);
    L.log2_cells = 0;
    while ((int64_t(8) << L.log2_cells) < n) ++L.log2_cells;     // 2^L cells ~ n / 8
    size_t off = 0;
    L.cdf = off;      off = align256(off + size_t(n) * 8);
    L.tile_off = off; off = align256(off + size_t(L.n_tiles + 1) * 8);
    L.guide = off;    off = align256(off + ((size_t(1) << L.log2_cells) + 2) * 4);
    L.total = off;
    return L;
}

// cell of a CDF value q < T: q >> s, with s such that T >> s < 2^L
__device__ __forceinline__ int guide_shift(unsigned long long T, int log2_cells) {
    const int bits = 64 - __clzll((long long)T);
    return bits > log2_cells ? bits - log2_cells : 0;
}

__device__ __forceinline__ unsigned long long quantise62(double w) {
    // round(w 2^62); NaN and non-positive weights carry no mass
    return w > 0.0 ? __double2ull_rn(fmin(w, 1.5) * 4611686018427387904.0) : 0ull;
}

// the thread's kMnItems consecutive weights (blocked layout), zero past the end
__device__ __forceinline__ void mn_load(const double* w, int64_t n, int64_t base, unsigned long long (&p)[kMnItems]) {
    if (base + kMnItems <= n) {
#pragma unroll
        for (int k = 0; k < kMnItems; k += 2) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(w + base + k));
            p[k] = quantise62(v.x);
            p[k + 1] = quantise62(v.y);
        }
    } else {
#pragma unroll
        for (int k = 0; k < kMnItems; ++k) p[k] = base + k < n ? quantise62(w[base + k]) : 0ull;
    }
}

__device__ __forceinline__ unsigned long long warp_incl_scan(unsigned long long v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(kThreads) k_mn_tile_totals(const double* w, int64_t n, unsigned long long* tile_off) {
    __shared__ unsigned long long s_w[kWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int64_t tile = blockIdx.x; tile * kMnTile < n; tile += gridDim.x) {
        unsigned long long p[kMnItems];
        mn_load(w, n, tile * kMnTile + (int64_t)tid * kMnItems, p);
        unsigned long long s = 0;
#pragma unroll
        for (int k = 0; k < kMnItems; ++k) s += p[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        __syncthreads();
        if (lane == 0) s_w[warp] = s;
        __syncthreads();
        if (tid == 0) {
            unsigned long long t = 0;
#pragma unroll
            for (int k = 0; k < kWarps; ++k) t += s_w[k];
            tile_off[tile] = t;
        }
    }
}

// one CTA: tile totals -> exclusive offsets in place, grand total in tile_off[n_tiles]
__global__ void __launch_bounds__(kMnScanThreads) k_mn_scan_tiles(unsigned long long* tile_off, int64_t n_tiles) {
    __shared__ unsigned long long s_w[kMnScanThreads / 32];
    __shared__ unsigned long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_tiles; base += kMnScanThreads) {
        const int64_t j = base + tid;
        const unsigned long long v = j < n_tiles ? tile_off[j] : 0ull;
        unsigned long long inc = warp_incl_scan(v, lane);
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long t = warp_incl_scan(s_w[lane], lane);
            s_w[lane] = t;
        }
        __syncthreads();
        const unsigned long long carry = s_carry;
        inc += carry + (warp > 0 ? s_w[warp - 1] : 0ull);
        if (j < n_tiles) tile_off[j] = inc - v;
        __syncthreads();
        if (tid == kMnScanThreads - 1) s_carry = inc;
        __syncthreads();
    }
    if (tid == 0) tile_off[n_tiles] = s_carry;
}

__global__ void __launch_bounds__(kThreads) k_mn_scan_write(const double* w, int64_t n, const unsigned long long* tile_off,
                                                            unsigned long long* cdf) {
    __shared__ unsigned long long s_w[kWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int64_t tile = blockIdx.x; tile * kMnTile < n; tile += gridDim.x) {
        const int64_t base = tile * kMnTile + (int64_t)tid * kMnItems;
        unsigned long long p[kMnItems];
        mn_load(w, n, base, p);
#pragma unroll
        for (int k = 1; k < kMnItems; ++k) p[k] += p[k - 1];
        const unsigned long long inc = warp_incl_scan(p[kMnItems - 1], lane);
        __syncthreads();
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        unsigned long long off = tile_off[tile] + inc - p[kMnItems - 1];
#pragma unroll
        for (int k = 0; k < kWarps; ++k)
            if (k < warp) off += s_w[k];
        if (base + kMnItems <= n) {
#pragma unroll
            for (int k = 0; k < kMnItems; k += 2)
                *reinterpret_cast<ulonglong2*>(cdf + base + k) = make_ulonglong2(off + p[k], off + p[k + 1]);
        } else {
#pragma unroll
            for (int k = 0; k < kMnItems; ++k)
                if (base + k < n) cdf[base + k] = off + p[k];
        }
    }
}

// floor(u T) for a double u in [0, 1) and T < 2^63, exactly (128-bit product of the significand)
__device__ __forceinline__ unsigned long long scale_uniform(double u, unsigned long long T) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(u);
    const int ex = (int)((bits >> 52) & 0x7ffull);
    if (!(u > 0.0) || ex == 0) return 0ull;                      // zero, negative, NaN, subnormal
    if (ex >= 1023) return T - 1;                                // u >= 1 is out of contract: last cell
    const unsigned long long m = (bits & ((1ull << 52) - 1)) | (1ull << 52);
    const int s = 1075 - ex;                                     // u = m 2^-s, s >= 53
    const unsigned long long hi = __umul64hi(m, T), lo = m * T;
    if (s >= 128) return 0ull;
    if (s >= 64) return hi >> (s - 64);
    return (hi << (64 - s)) | (lo >> s);
}

// count of C_j <= q over the non-decreasing array C[0, n)
__device__ __forceinline__ int64_t upper_bound_u64(const unsigned long long* C, int64_t n, unsigned long long q) {
    int64_t lo = 0, len = n;
    while (len > 0) {
        const int64_t half = len >> 1;
        const unsigned long long v = __ldg(C + lo + half);
        if (v <= q) { lo += half + 1; len -= half + 1; } else len = half;
    }
    return lo;
}

// guide[b] = #{ j : C_j < b 2^s } for b = 0 .. (T >> s) + 1: the answer for any q in cell b lies in
// [guide[b], guide[b + 1]]
__global__ void __launch_bounds__(kThreads) k_mn_guide(const unsigned long long* cdf, const unsigned long long* tile_off,
                                                       int64_t n_tiles, int64_t n, int log2_cells, int32_t* guide) {
    const unsigned long long T = tile_off[n_tiles];
    const int s = guide_shift(T, log2_cells);
    const unsigned long long cells = (T >> s) + 1;
    const unsigned long long b = (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
    if (b > cells) return;
    guide[b] = b == 0 ? 0 : (int32_t)upper_bound_u64(cdf, n, (b << s) - 1);
}

// two outputs per thread: one Philox block yields both uniforms in on-chip mode
__global__ void __launch_bounds__(kThreads) k_mn_search(const unsigned long long* cdf, const unsigned long long* tile_off,
                                                        int64_t n_tiles, int64_t n, int64_t n_out, const double* u_dev,
                                                        PhiloxKeys keys, uint32_t t, int log2_cells,
                                                        const int32_t* guide, int32_t* ancestors) {
    const unsigned long long T = tile_off[n_tiles];
    const int s = guide_shift(T, log2_cells);
    const int64_t pair = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t i0 = pair * 2;
    if (i0 >= n_out) return;
    double u[2];
    if (u_dev) {
        u[0] = u_dev[i0];
        u[1] = i0 + 1 < n_out ? u_dev[i0 + 1] : 0.0;
    } else {
        uint32_t r[4];
        philox_quad_keys(keys, (uint64_t)pair, t, kStreamMultinomial, r);
        u[0] = uniform53(r[0], r[1]);
        u[1] = uniform53(r[2], r[3]);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (i0 + k >= n_out) break;
        int64_t a = 0;
        if (T) {
            const unsigned long long q = scale_uniform(u[k], T);
            const unsigned long long b = q >> s;
            const int64_t lo = __ldg(guide + b), hi = __ldg(guide + b + 1);
            a = lo + upper_bound_u64(cdf + lo, hi - lo, q);
        }
        ancestors[i0 + k] = (int32_t)(a < n ? a : n - 1);
    }
}

}  // namespace
}  // namespace smc

using namespace smc;

extern "C" {

uint64_t smc_multinomial_workspace_bytes(int64_t n) { return n > 0 ? (uint64_t)mn_layout(n).total : 0; }

int smc_multinomial_ancestors(const double* w_dev, int64_t n, int64_t n_out, const double* u_dev, uint64_t seed,
                              int64_t t, int32_t* ancestors_dev, void* ws_dev, uint64_t ws_bytes, void* stream) {
    SMC_REQUIRE(n > 0 && n <= (int64_t)INT32_MAX && n_out > 0, "smc_multinomial_ancestors: bad size");
    SMC_REQUIRE(w_dev && ancestors_dev && ws_dev, "smc_multinomial_ancestors: null buffer");
    const MnLayout L = mn_layout(n);
    SMC_REQUIRE(ws_bytes >= L.total, "smc_multinomial_ancestors: workspace too small (need %zu bytes)", L.total);
    SMC_REQUIRE((reinterpret_cast<uintptr_t>(w_dev) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws_dev) & 255) == 0,
                "smc_multinomial_ancestors: weights must be 16-byte and the workspace 256-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(ws_dev);
    unsigned long long* cdf = reinterpret_cast<unsigned long long*>(ws + L.cdf);
    unsigned long long* tile_off = reinterpret_cast<unsigned long long*>(ws + L.tile_off);
    const unsigned grid = (unsigned)(L.n_tiles < 148 * 8 ? L.n_tiles : 148 * 8);
    k_mn_tile_totals<<<grid, kThreads, 0, st>>>(w_dev, n, tile_off);
    SMC_CHECK_CUDA(cudaGetLastError());
    k_mn_scan_tiles<<<1, kMnScanThreads, 0, st>>>(tile_off, L.n_tiles);
    SMC_CHECK_CUDA(cudaGetLastError());
    k_mn_scan_write<<<grid, kThreads, 0, st>>>(w_dev, n, tile_off, cdf);
    SMC_CHECK_CUDA(cudaGetLastError());
    int32_t* guide = reinterpret_cast<int32_t*>(ws + L.guide);
    const unsigned ggrid = (unsigned)cdiv64((int64_t(1) << L.log2_cells) + 1, kThreads);
    k_mn_guide<<<ggrid, kThreads, 0, st>>>(cdf, tile_off, L.n_tiles, n, L.log2_cells, guide);
    SMC_CHECK_CUDA(cudaGetLastError());
    const unsigned sgrid = (unsigned)cdiv64(cdiv64(n_out, 2), kThreads);
    k_mn_search<<<sgrid, kThreads, 0, st>>>(cdf, tile_off, L.n_tiles, n, n_out, u_dev, philox_expand(seed), (uint32_t)t,
                                            L.log2_cells, guide, ancestors_dev);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
