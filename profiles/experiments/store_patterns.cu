// Experiment (not part of the product): which store pattern reaches the write-stream peak for [n][311] int64 rows?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu && ./store_patterns
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int T = 128;
__global__ void __launch_bounds__(T) rowwise(uint64_t n, long long* __restrict__ out, int reps) {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t first = (uint64_t)blockIdx.x * T;
    for (int r = warp; r < T; r += T / 32) {
        uint64_t g = first + r;
        if (g >= n) break;
        long long* dst = out + g * 311;
        uint32_t w0 = (uint32_t)g * 2654435761u + lane, w1 = w0 >> 3;
#define CH(C, SH, M) dst[(C) * 62 + lane] = (long long)((w0 >> (SH)) & (M)); if (lane < 30) dst[(C) * 62 + 32 + lane] = (long long)((w1 >> (SH)) & (M));
        CH(0, 0, 63u) CH(1, 6, 63u) CH(2, 12, 7u) CH(3, 15, 15u) CH(4, 19, 3u)
        if (lane == 0) dst[310] = w0 & 3;
    }
}
__global__ void __launch_bounds__(T) dense16(uint64_t n, long long* __restrict__ out, int reps) {
    uint64_t first = (uint64_t)blockIdx.x * T;
    uint32_t rows = (uint32_t)min((uint64_t)T, n - first), total = rows * 311u;
    uint4* dst = reinterpret_cast<uint4*>(out + first * 311);
    for (uint32_t c = threadIdx.x; 2u * c + 1u < total; c += T) { uint32_t v = c * 2654435761u; dst[c] = make_uint4(v & 255u, 0u, (v >> 8) & 255u, 0u); }
}
__global__ void __launch_bounds__(T) dense8(uint64_t n, long long* __restrict__ out, int reps) {
    uint64_t first = (uint64_t)blockIdx.x * T;
    uint32_t rows = (uint32_t)min((uint64_t)T, n - first), total = rows * 311u;
    long long* dst = out + first * 311;
    for (uint32_t c = threadIdx.x; c < total; c += T) dst[c] = (long long)((c * 2654435761u) & 255u);
}
// grid-stride dense16 with few fat blocks (persistent style)
__global__ void __launch_bounds__(256) dense16_persistent(uint64_t n, long long* __restrict__ out, int reps) {
    uint64_t total = n * 311 / 2;
    uint4* dst = reinterpret_cast<uint4*>(out);
    for (uint64_t c = (uint64_t)blockIdx.x * 256 + threadIdx.x; c < total; c += (uint64_t)gridDim.x * 256) { uint32_t v = (uint32_t)c * 2654435761u; dst[c] = make_uint4(v & 255u, 0u, (v >> 8) & 255u, 0u); }
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}
int main() {
    uint64_t n = 1ull << 22; long long* out; cudaMalloc(&out, n * 311 * 8);
    double gb = n * 311 * 8 / 1e9; unsigned grid = (unsigned)(n / T);
    float t;
    t = timeit([&] { rowwise<<<grid, T>>>(n, out, 1); }); printf("rowwise_8B_256Bperinstr  %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
    t = timeit([&] { dense16<<<grid, T>>>(n, out, 1); }); printf("dense_16B_blocktile      %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
    t = timeit([&] { dense8<<<grid, T>>>(n, out, 1); }); printf("dense_8B_blocktile       %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
    t = timeit([&] { dense16_persistent<<<148 * 8, 256>>>(n, out, 1); }); printf("dense_16B_persistent     %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
    t = timeit([&] { cudaMemsetAsync(out, 1, n * 311 * 8); }); printf("cudaMemsetAsync          %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
    cudaError_t e = cudaDeviceSynchronize(); printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
