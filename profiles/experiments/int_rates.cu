// Microbenchmark (not part of the product): per-SM issue rates of the integer instructions the simulator kernels are made of
// ("INT32 lane throughput per SM and popc/brev/fns rates for sm_100 are not in MEASURED_PEAKS; the builder must
// microbenchmark them").  Every kernel runs ILP independent dependency chains of one PTX instruction per thread, 148 x 8 CTAs of 256
// threads (64 resident warps per SM), clock() deltas inside the kernel give cycles, so the rate is independent of the SM clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_rates int_rates.cu && ./int_rates
// The SASS of each loop body is checked with `cuobjdump -sass int_rates | grep -c <mnemonic>` (see profiles/r01_int_rates.md).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ILP = 8, UNROLL = 32, ITERS = 64, THREADS = 256;

// every warp records (start, end, SM id): clock64 is a per-SM counter, so max(end) - min(start) over the warps of one SM is the
// interval in which that SM executed its share of the work (warps do not progress evenly: the scheduler is not fair)
__device__ __forceinline__ void record_warp(unsigned long long* rec, long long t0, long long t1) {
    if ((threadIdx.x & 31) == 0) {
        uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long* r = rec + 3ull * (blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5));
        r[0] = (unsigned long long)t0; r[1] = (unsigned long long)t1; r[2] = smid;
    }
}
#define DEFINE_KERNEL(NAME, BODY)                                                                              \
    __global__ void __launch_bounds__(THREADS) NAME(uint32_t* out, unsigned long long* cycles, uint32_t seed) { \
        uint32_t x[ILP], y = seed | 1u, z = seed * 3u + 7u;                                                      \
        _Pragma("unroll") for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 2654435761u + j + seed;             \
        __syncthreads();                                                                                         \
        long long t0 = clock64();                                                                                \
        for (int it = 0; it < ITERS; ++it) {                                                                     \
            _Pragma("unroll") for (int u = 0; u < UNROLL; ++u) {                                                 \
                _Pragma("unroll") for (int j = 0; j < ILP; ++j) { BODY; }                                        \
            }                                                                                                    \
        }                                                                                                        \
        long long t1 = clock64();                                                                                \
        uint32_t acc = 0;                                                                                        \
        _Pragma("unroll") for (int j = 0; j < ILP; ++j) acc ^= x[j];                                             \
        out[blockIdx.x * THREADS + threadIdx.x] = acc + y + z;                                                   \
        record_warp(cycles, t0, t1);                                                                             \
    }

DEFINE_KERNEL(k_iadd3, asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(y)))
DEFINE_KERNEL(k_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(y), "r"(z)))
DEFINE_KERNEL(k_shf, asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[j]) : "r"(y)))
DEFINE_KERNEL(k_shl, asm volatile("shr.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(y)))
DEFINE_KERNEL(k_imad, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(y), "r"(z)))
DEFINE_KERNEL(k_mulhi, asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(y)))
DEFINE_KERNEL(k_popc, asm volatile("popc.b32 %0, %0;" : "+r"(x[j])))
DEFINE_KERNEL(k_brev, asm volatile("brev.b32 %0, %0;" : "+r"(x[j])))
DEFINE_KERNEL(k_clz, asm volatile("clz.b32 %0, %0;" : "+r"(x[j])))
DEFINE_KERNEL(k_bfind, asm volatile("bfind.u32 %0, %0;" : "+r"(x[j])))
DEFINE_KERNEL(k_prmt, asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(x[j]) : "r"(y)))
DEFINE_KERNEL(k_bfe, asm volatile("bfe.u32 %0, %0, 3, 27;" : "+r"(x[j])))
DEFINE_KERNEL(k_sel, asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(x[j]) : "r"(y), "r"(z)))
DEFINE_KERNEL(k_min, asm volatile("min.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(y)))
DEFINE_KERNEL(k_fns, asm volatile("fns.b32 %0, %0, 0, 2;" : "+r"(x[j])))
DEFINE_KERNEL(k_shfl, asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(x[j])))
DEFINE_KERNEL(k_vote, asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; vote.sync.ballot.b32 %0, p, 0xffffffff; }" : "+r"(x[j])))
// pipe pairing: one ALU-pipe op and one FMA-pipe op per slot (do they issue in the same cycles?)
DEFINE_KERNEL(k_lop3_imad, if (j & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(y), "r"(z));
                           else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(y), "r"(z)))
DEFINE_KERNEL(k_lop3_popc, if (j & 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(y), "r"(z));
                           else asm volatile("popc.b32 %0, %0;" : "+r"(x[j])))

// shared-memory table lookup (the card-attribute / rank-select LUTs): a chain of conflict-free LDS whose result is the next address
__global__ void __launch_bounds__(THREADS) k_lds(uint32_t* out, unsigned long long* cycles, uint32_t seed) {
    __shared__ uint32_t tab[1024];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    for (int i = threadIdx.x; i < 1024; i += THREADS) tab[i] = base + 4u * ((((i >> 5) * 13u + seed) & 31u) * 32u + (i & 31u));
    uint32_t x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = base + 4u * ((threadIdx.x & 31u) + 32u * j);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) asm volatile("ld.shared.u32 %0, [%0];" : "+r"(x[j]) : : "memory");
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc ^= x[j];
    out[blockIdx.x * THREADS + threadIdx.x] = acc;
    record_warp(cycles, t0, t1);
}

typedef void (*kern_t)(uint32_t*, unsigned long long*, uint32_t);
struct Entry { const char* name; const char* sass; kern_t fn; double sass_per_op; };

int main() {
    int dev = 0, sms = 0, clock_khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev);
    const int ctas_per_sm = 8, grid = sms * ctas_per_sm;
    uint32_t* out; unsigned long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * grid * THREADS);
    const int n_warps = grid * THREADS / 32;
    cudaMalloc(&cyc, sizeof(unsigned long long) * 3 * n_warps);
    unsigned long long* h = new unsigned long long[3 * n_warps];
    Entry es[] = {
        {"add.u32 (two fuse into one IADD3)", "IADD3", k_iadd3, 0.5}, {"lop3.b32", "LOP3.LUT", k_lop3, 1}, {"shf.l.wrap.b32", "SHF.L.W", k_shf, 1}, {"shr.u32", "SHF.R.U32.HI", k_shl, 1},
        {"mad.lo.u32", "IMAD", k_imad, 1}, {"mul.hi.u32", "IMAD.HI", k_mulhi, 1}, {"popc.b32", "POPC", k_popc, 1}, {"brev.b32", "BREV", k_brev, 1},
        {"clz.b32", "FLO+IADD3", k_clz, 2}, {"bfind.u32", "FLO", k_bfind, 1}, {"prmt.b32", "PRMT", k_prmt, 1}, {"bfe.u32", "SHF+SGXT", k_bfe, 2},
        {"setp+selp", "ISETP+SEL", k_sel, 2}, {"min.u32 (two fuse into one VIMNMX3)", "VIMNMX3", k_min, 0.5}, {"fns.b32", "(emulated)", k_fns, 1},
        {"shfl.sync.bfly", "SHFL.BFLY", k_shfl, 1}, {"vote.ballot", "ISETP+VOTE", k_vote, 2},
        {"lop3 | mad.lo 1:1", "LOP3+IMAD", k_lop3_imad, 1}, {"lop3 | popc 3:1", "LOP3+POPC", k_lop3_popc, 1}, {"ld.shared.u32", "LDS", k_lds, 1},
    };
    printf("{\"device_sms\": %d, \"sm_clock_khz_attr\": %d, \"resident_warps_per_sm\": %d, \"ilp\": %d, \"rates\": [\n", sms, clock_khz, ctas_per_sm * THREADS / 32, ILP);
    const int n = sizeof(es) / sizeof(es[0]);
    for (int e = 0; e < n; ++e) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        es[e].fn<<<grid, THREADS>>>(out, cyc, 1u);                     // warm-up
        cudaEventRecord(a);
        es[e].fn<<<grid, THREADS>>>(out, cyc, 12345u);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        cudaMemcpy(h, cyc, sizeof(unsigned long long) * 3 * n_warps, cudaMemcpyDeviceToHost);
        unsigned long long lo[256], hi[256]; int cnt[256];
        for (int i = 0; i < 256; ++i) { lo[i] = ~0ull; hi[i] = 0; cnt[i] = 0; }
        for (int w = 0; w < n_warps; ++w) {
            int sm = (int)h[3 * w + 2] & 255;
            if (h[3 * w] < lo[sm]) lo[sm] = h[3 * w];
            if (h[3 * w + 1] > hi[sm]) hi[sm] = h[3 * w + 1];
            cnt[sm]++;
        }
        double mean = 0, warps_per_sm = 0; int used = 0;
        for (int i = 0; i < 256; ++i) if (cnt[i]) { mean += (double)(hi[i] - lo[i]); warps_per_sm += cnt[i]; used++; }
        mean /= used; warps_per_sm /= used;
        // SM-wide work during the SM-wide interval
        double thread_ops_per_sm = warps_per_sm * 32.0 * ILP * UNROLL * ITERS;
        double lanes_per_clk = thread_ops_per_sm / mean;
        double total_ops = (double)grid * THREADS * ILP * UNROLL * ITERS;
        printf("  {\"ptx\": \"%s\", \"sass\": \"%s\", \"thread_ops_per_clk_per_sm\": %.1f, \"sass_inst_per_op\": %.1f, \"warp_inst_per_clk_per_smsp\": %.3f, \"cycles\": %.0f, "
               "\"event_ms\": %.4f, \"Gops_per_s_chip\": %.0f, \"implied_sm_mhz\": %.0f}%s\n",
               es[e].name, es[e].sass, lanes_per_clk, es[e].sass_per_op, lanes_per_clk * es[e].sass_per_op / 128.0, mean, ms, total_ops / ms / 1e6, mean / ms / 1e3, e + 1 < n ? "," : "");
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("], \"status\": \"%s\"}\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}
