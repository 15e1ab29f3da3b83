// cabi.cu — the extern "C" boundary of libdoko_cuda.so (include/doko_cuda.h).
// No torch types, no CPU fallback: every compute entry point launches sm_100a kernels or fails with a status code.
#include <cuda.h>   // CUtensorMap and its enums only: the encoder is fetched with cudaGetDriverEntryPoint, libcuda is not linked
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/doko_cuda.h"
#include "kernels.cuh"
#include "selfplay_kernels.cuh"

struct dk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;   // the context's own stream
    uint64_t launches = 0;
    std::string last_error;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    void* tmap_encode = nullptr;     // cuTensorMapEncodeTiled
    void* leaf_ws = nullptr;         // samples of dk_leaf_rollouts(determinize): K3's output for one chunk of leaves
    size_t leaf_ws_bytes = 0;
    void* pimc_ws = nullptr;         // workspace of dk_pimc_evaluate (post-action playout states per determinization and legal action)
    size_t pimc_ws_bytes = 0;
    bool fresh_smem_set = false;     // dynamic shared-memory opt-in of the fresh-game playout kernels done on this device
    size_t total_mem = 0;
    // scratch for the *_host entry points
    void* d_scratch = nullptr;
    size_t d_scratch_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t stream2 = nullptr;  // second compute stream of the *_host entry points (chunk kernels alternate, so a chunk's last wave overlaps the next chunk)
    cudaEvent_t h2d_done = nullptr;
    std::vector<cudaEvent_t> events;
    // NCCL (loaded lazily)
    // ln(N) table of the UCT search (host libm values, see uct.cuh)
    double* d_ln_table = nullptr;
    size_t ln_table_len = 0;
    // side streams of the UCT search (parts of a batch run phase-shifted) and the events that fork / join them
    cudaStream_t uct_streams[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t uct_fork = nullptr, uct_join[3] = {nullptr, nullptr, nullptr};
    void* nccl_lib = nullptr;
    void* nccl_comm = nullptr;
    int nccl_ranks = 0, nccl_rank = 0;
};

namespace {

dk_status fail(dk_ctx* ctx, dk_status st, const std::string& msg) {
    if (ctx) ctx->last_error = msg;
    return st;
}
#define DK_TRY(expr)                                                                                    \
    do {                                                                                                \
        dk_status st__ = (expr);                                                                        \
        if (st__ != DK_OK) return st__;                                                                 \
    } while (0)
#define DK_CUDA(ctx, call)                                                                              \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail((ctx), DK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));       \
    } while (0)

cudaStream_t pick_stream(dk_ctx* ctx, dk_stream s) { return s ? (cudaStream_t)s : ctx->stream; }

// The kernels use 16-byte vector accesses on the record arrays and on the outputs named in the header, 4-byte ones on packed words.
// A misaligned caller buffer (a sliced view, an offset Rust slice) would fault on the device — a sticky error for the whole CUDA
// context — so it is refused here.  NULL passes (nullable arguments are checked for presence separately).
bool aligned_to(const void* p, uintptr_t a) { return ((uintptr_t)p & (a - 1u)) == 0u; }
dk_status misaligned(dk_ctx* ctx, const char* what) { return fail(ctx, DK_ERR_INVALID_ARGUMENT, std::string(what) + ": misaligned pointer"); }
#define DK_ALIGNED(ctx, ptr, a)                                                                         \
    do {                                                                                                \
        if (!aligned_to((ptr), (a))) return misaligned((ctx), #ptr);                                    \
    } while (0)

// Tensor map of a record array ([n][128] bytes, tiles of up to 128 records, 128-byte swizzle) for the TMA kernels.  False when the
// driver entry point is missing or refuses the array (then the callers use their cooperative-copy kernels).
typedef CUresult (*dk_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
bool state_tensor_map(dk_ctx* ctx, const dk_state* states, size_t n, CUtensorMap* out) {
    if (getenv("DOKO_CUDA_NO_TMA") || ((uintptr_t)states & 15u) || n >= (1ull << 31)) return false;
    if (!ctx->tmap_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError();
            return false;
        }
        ctx->tmap_encode = fn;
    }
    const cuuint64_t dims[2] = {128u, (cuuint64_t)n}, strides[1] = {128u};
    const cuuint32_t box[2] = {128u, (cuuint32_t)(n < (size_t)dk::STATE_THREADS ? n : (size_t)dk::STATE_THREADS)}, elem[2] = {1u, 1u};
    return ((dk_tmap_encode_fn)ctx->tmap_encode)(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<dk_state*>(states), dims, strides, box, elem,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

dk::RngParams to_params(const dk_rng* r) {
    dk::RngParams p;
    p.seed_lo = (uint32_t)r->seed; p.seed_hi = (uint32_t)(r->seed >> 32); p.first_id = r->first_id; p.epoch = r->epoch; p.first_sub = r->first_sub;
    return p;
}

dk_status ensure_scratch(dk_ctx* ctx, size_t bytes) {
    if (ctx->d_scratch_bytes >= bytes) return DK_OK;
    if (ctx->d_scratch) { cudaFree(ctx->d_scratch); ctx->d_scratch = nullptr; ctx->d_scratch_bytes = 0; }
    DK_CUDA(ctx, cudaMalloc(&ctx->d_scratch, bytes));
    ctx->d_scratch_bytes = bytes;
    return DK_OK;
}

dk_status check_launch(dk_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DK_ERR_CUDA, std::string(what) + " launch: " + cudaGetErrorString(e));
    ctx->launches++;
    return DK_OK;
}

}  // namespace

extern "C" {

const char* dk_version(void) { return "doko_cuda 0.3 (sm_100a)"; }

dk_status dk_init(int device, dk_ctx** out) {
    if (!out) return DK_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return DK_ERR_NO_DEVICE;
    dk_ctx* ctx = new (std::nothrow) dk_ctx();
    if (!ctx) return DK_ERR_INVALID_ARGUMENT;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return DK_ERR_NO_DEVICE; }
    ctx->sm_count = prop.multiProcessorCount; ctx->cc_major = prop.major; ctx->cc_minor = prop.minor; ctx->total_mem = prop.totalGlobalMem;
    if (prop.major != 10) {   // the fatbin only holds sm_100a code
        delete ctx;
        return DK_ERR_NO_DEVICE;
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return DK_ERR_CUDA; }
    {   // the kernels' lookup tables: evaluated here, resident in this device's memory from now on
        std::vector<uint32_t> lut(dk::FULL_LUT_WORDS);
        for (uint32_t i = 0; i < dk::FULL_LUT_WORDS; ++i) lut[i] = dk::lut_word(i);
        for (uint32_t h = 0; h < dk::SEL12_WORDS / 2; ++h) { const uint64_t e = dk::sel12_entry(h); std::memcpy(&lut[dk::SEL12_LUT_BASE + 2u * h], &e, 8); }
        if (cudaMemcpyToSymbol(dk::g_lut, lut.data(), lut.size() * sizeof(uint32_t)) != cudaSuccess) { cudaStreamDestroy(ctx->stream); delete ctx; return DK_ERR_CUDA; }
    }
    *out = ctx;
    return DK_OK;
}

dk_status dk_destroy(dk_ctx* ctx) {
    if (!ctx) return DK_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    dk_comm_destroy(ctx);                                // a communicator the caller left open
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->pimc_ws) cudaFree(ctx->pimc_ws);
    if (ctx->leaf_ws) cudaFree(ctx->leaf_ws);
    if (ctx->d_ln_table) cudaFree(ctx->d_ln_table);
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    for (int i = 0; i < 3; ++i) { if (ctx->uct_streams[i]) cudaStreamDestroy(ctx->uct_streams[i]); if (ctx->uct_join[i]) cudaEventDestroy(ctx->uct_join[i]); }
    if (ctx->uct_fork) cudaEventDestroy(ctx->uct_fork);
    if (ctx->h2d_done) cudaEventDestroy(ctx->h2d_done);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return DK_OK;
}

const char* dk_last_error(const dk_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

dk_status dk_device_info(const dk_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
    if (!ctx) return DK_ERR_INVALID_ARGUMENT;
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (total_mem) *total_mem = ctx->total_mem;
    return DK_OK;
}

dk_status dk_synchronize(dk_ctx* ctx, dk_stream stream) {
    if (!ctx) return DK_ERR_INVALID_ARGUMENT;
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    DK_CUDA(ctx, cudaStreamSynchronize(pick_stream(ctx, stream)));
    return DK_OK;
}

uint64_t dk_launch_count(const dk_ctx* ctx) { return ctx ? ctx->launches : 0; }

// The fresh-game playout kernels use more than the 48 KB of shared memory a kernel gets by default: opt in once per device.
static dk_status fresh_smem_opt_in(dk_ctx* ctx) {
    if (ctx->fresh_smem_set) return DK_OK;
    const int bytes = (int)dk::FDO_FRESH_SMEM_BYTES;
    DK_CUDA(ctx, cudaFuncSetAttribute(dk::fdo_playout_fresh_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    DK_CUDA(ctx, cudaFuncSetAttribute(dk::fdo_playout_fresh_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    DK_CUDA(ctx, cudaFuncSetAttribute(dk::doko_playout_fresh_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    DK_CUDA(ctx, cudaFuncSetAttribute(dk::doko_playout_fresh_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    ctx->fresh_smem_set = true;
    return DK_OK;
}
static dk_status playout_launch(dk_ctx* ctx, int engine, uint32_t flags, size_t n, const dk_state* states, const dk::RngParams& rp,
                                void* points_out, void* steps_out, uint32_t mode, unsigned long long* stats, cudaStream_t s) {
    const bool with_ann = (flags & DK_PLAYOUT_WITH_ANNOUNCEMENTS) != 0;
    if (states == nullptr) {
        const unsigned fgrid = (unsigned)((n + dk::FDO_FRESH_THREADS - 1) / dk::FDO_FRESH_THREADS);
        DK_TRY(fresh_smem_opt_in(ctx));
        if (engine == DK_FDO) {
            if (with_ann) dk::fdo_playout_fresh_kernel<true><<<fgrid, dk::FDO_FRESH_THREADS, dk::FDO_FRESH_SMEM_BYTES, s>>>(rp, (uint64_t)n, points_out, steps_out, mode, stats);
            else dk::fdo_playout_fresh_kernel<false><<<fgrid, dk::FDO_FRESH_THREADS, dk::FDO_FRESH_SMEM_BYTES, s>>>(rp, (uint64_t)n, points_out, steps_out, mode, stats);
            return check_launch(ctx, "fdo_playout_fresh_kernel");
        }
        dk::doko_playout_fresh_kernel<false><<<fgrid, dk::FDO_FRESH_THREADS, dk::FDO_FRESH_SMEM_BYTES, s>>>(rp, (uint64_t)n, points_out, steps_out, mode, stats, nullptr, nullptr);
        return check_launch(ctx, "doko_playout_fresh_kernel");
    }
    unsigned grid = (unsigned)((n + dk::PLAYOUT_STATE_THREADS - 1) / dk::PLAYOUT_STATE_THREADS);
    if (engine == DK_FDO) {
        if (with_ann) dk::playout_state_kernel<DK_FDO, true><<<grid, dk::PLAYOUT_STATE_THREADS, 0, s>>>(rp, (uint64_t)n, states, 1u, points_out, steps_out, mode, stats);
        else dk::playout_state_kernel<DK_FDO, false><<<grid, dk::PLAYOUT_STATE_THREADS, 0, s>>>(rp, (uint64_t)n, states, 1u, points_out, steps_out, mode, stats);
    } else {
        dk::playout_state_kernel<DK_DOKO, false><<<grid, dk::PLAYOUT_STATE_THREADS, 0, s>>>(rp, (uint64_t)n, states, 1u, points_out, steps_out, mode, stats);
    }
    return check_launch(ctx, "playout_state_kernel");
}

dk_status dk_playout(dk_ctx* ctx, int engine, uint32_t flags, size_t n, const dk_state* states, const dk_rng* rng,
                     int32_t* points_out, uint32_t* steps_out, dk_stream stream) {
    if (!ctx || !rng || (engine != DK_FDO && engine != DK_DOKO)) return DK_ERR_INVALID_ARGUMENT;
    if (n == 0) return DK_OK;
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, points_out, 16); DK_ALIGNED(ctx, steps_out, 4);
    return playout_launch(ctx, engine, flags, n, states, to_params(rng), points_out, steps_out, dk::OUT_INT32, nullptr, pick_stream(ctx, stream));
}

dk_status dk_playout_summary(dk_ctx* ctx, int engine, uint32_t flags, size_t n, const dk_state* states, const dk_rng* rng, dk_playout_stats* stats,
                             int accumulate, dk_stream stream) {
    if (!ctx || !rng || !stats || (engine != DK_FDO && engine != DK_DOKO)) return DK_ERR_INVALID_ARGUMENT;
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, stats, 8);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = pick_stream(ctx, stream);
    if (!accumulate) DK_CUDA(ctx, cudaMemsetAsync(stats, 0, sizeof(dk_playout_stats), s));
    if (n == 0) return DK_OK;
    return playout_launch(ctx, engine, flags, n, states, to_params(rng), nullptr, nullptr, dk::OUT_INT32, (unsigned long long*)stats, s);
}

dk_status dk_playout_trace(dk_ctx* ctx, int engine, size_t n, const dk_rng* rng, int32_t* points_out, uint8_t* trace_out, uint32_t* aux_out,
                           dk_stream stream) {
    if (!ctx || !rng) return DK_ERR_INVALID_ARGUMENT;
    if (engine != DK_DOKO) return fail(ctx, DK_ERR_UNSUPPORTED, "dk_playout_trace: only DK_DOKO records traces");
    if (n == 0) return DK_OK;
    DK_ALIGNED(ctx, points_out, 16); DK_ALIGNED(ctx, aux_out, 16);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    DK_TRY(fresh_smem_opt_in(ctx));
    unsigned grid = (unsigned)((n + dk::FDO_FRESH_THREADS - 1) / dk::FDO_FRESH_THREADS);
    dk::doko_playout_fresh_kernel<true><<<grid, dk::FDO_FRESH_THREADS, dk::FDO_FRESH_SMEM_BYTES, pick_stream(ctx, stream)>>>(to_params(rng), (uint64_t)n, points_out, nullptr, dk::OUT_INT32, nullptr,
                                                                                                    trace_out, (uint4*)aux_out);
    return check_launch(ctx, "doko_playout_fresh_kernel<trace>");
}

dk_status dk_new_games(dk_ctx* ctx, int engine, size_t n, const dk_rng* rng, dk_state* out, dk_stream stream) {
    if (!ctx || !rng || !out || (engine != DK_FDO && engine != DK_DOKO)) return DK_ERR_INVALID_ARGUMENT;
    if (n == 0) return DK_OK;
    DK_ALIGNED(ctx, out, 16);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned grid = (unsigned)((n + dk::PLAYOUT_THREADS - 1) / dk::PLAYOUT_THREADS);
    CUtensorMap tmap;
    if (state_tensor_map(ctx, out, n, &tmap)) {
        dk::new_games_tma_kernel<<<grid, dk::PLAYOUT_THREADS, 0, pick_stream(ctx, stream)>>>(tmap, to_params(rng), (uint64_t)n);
        return check_launch(ctx, "new_games_tma_kernel");
    }
    dk::new_games_kernel<<<grid, dk::PLAYOUT_THREADS, 0, pick_stream(ctx, stream)>>>(to_params(rng), (uint64_t)n, out);
    return check_launch(ctx, "new_games_kernel");
}

dk_status dk_from_deals(dk_ctx* ctx, int engine, size_t n, const uint64_t* hands, const uint8_t* start, dk_state* out, dk_stream stream) {
    if (!ctx || !hands || !start || !out || (engine != DK_FDO && engine != DK_DOKO)) return DK_ERR_INVALID_ARGUMENT;
    if (n == 0) return DK_OK;
    DK_ALIGNED(ctx, out, 16); DK_ALIGNED(ctx, hands, 8);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned grid = (unsigned)((n + dk::STATE_THREADS - 1) / dk::STATE_THREADS);
    CUtensorMap tmap;
    if (!((uintptr_t)hands & 15u) && state_tensor_map(ctx, out, n, &tmap)) {
        dk::from_deals_tma_kernel<<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>(tmap, (uint64_t)n, hands, start);
        return check_launch(ctx, "from_deals_tma_kernel");
    }
    dk::from_deals_kernel<<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>((uint64_t)n, hands, start, out);
    return check_launch(ctx, "from_deals_kernel");
}

static dk_status legal_mask_launch(dk_ctx* ctx, int engine, size_t n, const dk_state* states, uint64_t* mask_out, uint64_t drop_mask, uint64_t drop_count,
                                   uint8_t* count_out, dk_stream stream) {
    if (n == 0) return DK_OK;
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, mask_out, 8);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned grid = (unsigned)((n + dk::STATE_THREADS - 1) / dk::STATE_THREADS);
    CUtensorMap tmap;
    if (state_tensor_map(ctx, states, n, &tmap)) {
        if (engine == DK_FDO) dk::legal_mask_tma_kernel<DK_FDO><<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>(tmap, (uint64_t)n, mask_out, drop_mask, drop_count, count_out);
        else dk::legal_mask_tma_kernel<DK_DOKO><<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>(tmap, (uint64_t)n, mask_out, drop_mask, drop_count, count_out);
        return check_launch(ctx, "legal_mask_tma_kernel");
    }
    if (engine == DK_FDO) dk::legal_mask_kernel<DK_FDO><<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>((uint64_t)n, states, mask_out, drop_mask, drop_count, count_out);
    else dk::legal_mask_kernel<DK_DOKO><<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>((uint64_t)n, states, mask_out, drop_mask, drop_count, count_out);
    return check_launch(ctx, "legal_mask_kernel");
}
dk_status dk_legal_mask(dk_ctx* ctx, int engine, size_t n, const dk_state* states, uint64_t* mask_out, dk_stream stream) {
    if (!ctx || !states || !mask_out || (engine != DK_FDO && engine != DK_DOKO)) return DK_ERR_INVALID_ARGUMENT;
    return legal_mask_launch(ctx, engine, n, states, mask_out, 0ull, 0ull, nullptr, stream);
}
dk_status dk_legal_mask_az(dk_ctx* ctx, size_t n, const dk_state* states, int is_secondary, uint64_t az_epoch, uint64_t* mask_out, uint8_t* n_allowed_out,
                           dk_stream stream) {
    if (!ctx || !states || (!mask_out && !n_allowed_out)) return DK_ERR_INVALID_ARGUMENT;
    const uint64_t calls = 0x1Full << 33;                                      // AnnouncementReContra .. AnnouncementBlack
    const bool young = az_epoch < DK_AZ_MIN_EPOCH;
    return legal_mask_launch(ctx, DK_FDO, n, states, mask_out, (is_secondary || young) ? calls : 0ull, young ? calls : 0ull, n_allowed_out, stream);
}
dk_status dk_random_action(dk_ctx* ctx, int engine, size_t n, const dk_state* states, const dk_rng* rng, uint32_t flags, uint8_t* action_out, dk_stream stream) {
    if (!ctx || !states || !rng || !action_out || (engine != DK_FDO && engine != DK_DOKO)) return DK_ERR_INVALID_ARGUMENT;
    if (n == 0) return DK_OK;
    DK_ALIGNED(ctx, states, 16);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned grid = (unsigned)((n + dk::STATE_THREADS - 1) / dk::STATE_THREADS);
    if (engine == DK_FDO) dk::random_action_kernel<DK_FDO><<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>(to_params(rng), (uint64_t)n, states, flags, action_out);
    else dk::random_action_kernel<DK_DOKO><<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>(to_params(rng), (uint64_t)n, states, flags, action_out);
    return check_launch(ctx, "random_action_kernel");
}
dk_status dk_state_id(dk_ctx* ctx, size_t n, const dk_state* states, const uint8_t* last_action, uint64_t* id_out, dk_stream stream) {
    if (!ctx || !states || !id_out) return DK_ERR_INVALID_ARGUMENT;
    if (n == 0) return DK_OK;
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, id_out, 8);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned grid = (unsigned)((n + dk::STATE_THREADS - 1) / dk::STATE_THREADS);
    dk::state_id_kernel<<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>((uint64_t)n, states, last_action, id_out);
    return check_launch(ctx, "state_id_kernel");
}

dk_status dk_apply(dk_ctx* ctx, int engine, size_t n, dk_state* states, const uint8_t* action_idx, uint32_t flags, uint8_t* err_out, dk_stream stream) {
    if (!ctx || !states || !action_idx || (engine != DK_FDO && engine != DK_DOKO)) return DK_ERR_INVALID_ARGUMENT;
    if (n == 0) return DK_OK;
    DK_ALIGNED(ctx, states, 16);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned grid = (unsigned)((n + dk::STATE_THREADS - 1) / dk::STATE_THREADS);
    CUtensorMap tmap;
    if (state_tensor_map(ctx, states, n, &tmap)) {
        const unsigned tgrid = (unsigned)((n + dk::STATE_THREADS * DK_APPLY_TILES - 1) / (dk::STATE_THREADS * DK_APPLY_TILES));
        if (engine == DK_FDO) dk::apply_tma_kernel<DK_FDO><<<tgrid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>(tmap, (uint64_t)n, action_idx, flags, err_out);
        else dk::apply_tma_kernel<DK_DOKO><<<tgrid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>(tmap, (uint64_t)n, action_idx, flags, err_out);
        return check_launch(ctx, "apply_tma_kernel");
    }
    if (engine == DK_FDO) dk::apply_kernel<DK_FDO><<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>((uint64_t)n, states, action_idx, flags, err_out);
    else dk::apply_kernel<DK_DOKO><<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>((uint64_t)n, states, action_idx, flags, err_out);
    return check_launch(ctx, "apply_kernel");
}

dk_status dk_terminal(dk_ctx* ctx, int engine, size_t n, const dk_state* states, uint8_t* done_out, int32_t* points_out, dk_stream stream) {
    if (!ctx || !states || (engine != DK_FDO && engine != DK_DOKO)) return DK_ERR_INVALID_ARGUMENT;
    if (n == 0) return DK_OK;
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, points_out, 16);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned grid = (unsigned)((n + dk::STATE_THREADS - 1) / dk::STATE_THREADS);
    dk::terminal_kernel<<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>((uint64_t)n, states, done_out, (int4*)points_out);
    return check_launch(ctx, "terminal_kernel");
}

dk_status dk_encode(dk_ctx* ctx, int layout, size_t n, const dk_state* states, int64_t* out, size_t row_stride, dk_stream stream) {
    if (!ctx || !states || !out) return DK_ERR_INVALID_ARGUMENT;
    size_t len = layout == DK_LAYOUT_FDO_PI311 ? 311 : (layout == DK_LAYOUT_DO114 ? 114 : (layout == DK_LAYOUT_DO110 ? 110 : 0));
    if (len == 0 || row_stride < len) return DK_ERR_INVALID_ARGUMENT;
    if (n == 0) return DK_OK;
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, out, 8);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = pick_stream(ctx, stream);
    unsigned grid = (unsigned)((n + dk::ENC_THREADS - 1) / dk::ENC_THREADS);
    if (layout == DK_LAYOUT_FDO_PI311) dk::encode_pi_kernel<<<grid, dk::ENC_THREADS, 0, s>>>((uint64_t)n, states, out, row_stride);
    else if (layout == DK_LAYOUT_DO114) dk::encode_kernel<DK_LAYOUT_DO114><<<grid, dk::ENC_THREADS, 0, s>>>((uint64_t)n, states, out, row_stride);
    else dk::encode_kernel<DK_LAYOUT_DO110><<<grid, dk::ENC_THREADS, 0, s>>>((uint64_t)n, states, out, row_stride);
    return check_launch(ctx, "encode_kernel");
}

dk_status dk_step_random_encode(dk_ctx* ctx, size_t n, dk_state* states, const dk_rng* rng, uint32_t flags, int64_t* obs_out, size_t row_stride,
                                uint8_t* action_out, dk_stream stream) {
    if (!ctx || !states || !rng || (obs_out && row_stride < 311)) return DK_ERR_INVALID_ARGUMENT;
    if (n == 0) return DK_OK;
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, obs_out, 8);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned grid = (unsigned)((n + dk::ENC_THREADS - 1) / dk::ENC_THREADS);
    CUtensorMap tmap;
    if (state_tensor_map(ctx, states, n, &tmap)) {
        dk::fdo_step_encode_tma_kernel<<<grid, dk::ENC_THREADS, 0, pick_stream(ctx, stream)>>>(tmap, to_params(rng), (uint64_t)n, flags, obs_out, row_stride, action_out);
        return check_launch(ctx, "fdo_step_encode_tma_kernel");
    }
    dk::fdo_step_encode_kernel<<<grid, dk::ENC_THREADS, 0, pick_stream(ctx, stream)>>>(to_params(rng), (uint64_t)n, states, flags, obs_out, row_stride, action_out);
    return check_launch(ctx, "fdo_step_encode_kernel");
}

dk_status dk_encode_narrow(dk_ctx* ctx, int layout, int elem_bytes, size_t n, const dk_state* states, void* out, dk_stream stream) {
    if (!ctx || !states || !out || (elem_bytes != 1 && elem_bytes != 4)) return DK_ERR_INVALID_ARGUMENT;
    if (layout != DK_LAYOUT_FDO_PI311 && layout != DK_LAYOUT_DO114 && layout != DK_LAYOUT_DO110) return DK_ERR_INVALID_ARGUMENT;
    if (n == 0) return DK_OK;
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, out, 32);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = pick_stream(ctx, stream);
    const unsigned grid = (unsigned)((n + dk::ENC_THREADS - 1) / dk::ENC_THREADS);
    const uint64_t m = (uint64_t)n;
    if (elem_bytes == 4) {
        int32_t* o = (int32_t*)out;
        if (layout == DK_LAYOUT_FDO_PI311) dk::encode_pi_narrow_kernel<int32_t><<<grid, dk::ENC_THREADS, 0, s>>>(m, states, o);
        else if (layout == DK_LAYOUT_DO114) dk::encode_narrow_kernel<DK_LAYOUT_DO114, int32_t><<<grid, dk::ENC_THREADS, 0, s>>>(m, states, o);
        else dk::encode_narrow_kernel<DK_LAYOUT_DO110, int32_t><<<grid, dk::ENC_THREADS, 0, s>>>(m, states, o);
    } else {
        uint8_t* o = (uint8_t*)out;
        if (layout == DK_LAYOUT_FDO_PI311) dk::encode_pi_narrow_kernel<uint8_t><<<grid, dk::ENC_THREADS, 0, s>>>(m, states, o);
        else if (layout == DK_LAYOUT_DO114) dk::encode_narrow_kernel<DK_LAYOUT_DO114, uint8_t><<<grid, dk::ENC_THREADS, 0, s>>>(m, states, o);
        else dk::encode_narrow_kernel<DK_LAYOUT_DO110, uint8_t><<<grid, dk::ENC_THREADS, 0, s>>>(m, states, o);
    }
    return check_launch(ctx, "encode_narrow_kernel");
}

dk_status dk_step_random_encode_narrow(dk_ctx* ctx, size_t n, dk_state* states, const dk_rng* rng, uint32_t flags, int elem_bytes, void* obs_out,
                                       uint8_t* action_out, dk_stream stream) {
    if (!ctx || !states || !rng || !obs_out || (elem_bytes != 1 && elem_bytes != 4)) return DK_ERR_INVALID_ARGUMENT;
    if (n == 0) return DK_OK;
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, obs_out, 32);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned grid = (unsigned)((n + dk::ENC_THREADS - 1) / dk::ENC_THREADS);
    CUtensorMap tmap;
    if (state_tensor_map(ctx, states, n, &tmap)) {
        if (elem_bytes == 4)
            dk::fdo_step_encode_narrow_tma_kernel<int32_t><<<grid, dk::ENC_THREADS, 0, pick_stream(ctx, stream)>>>(tmap, to_params(rng), (uint64_t)n, flags, (int32_t*)obs_out, action_out);
        else
            dk::fdo_step_encode_narrow_tma_kernel<uint8_t><<<grid, dk::ENC_THREADS, 0, pick_stream(ctx, stream)>>>(tmap, to_params(rng), (uint64_t)n, flags, (uint8_t*)obs_out, action_out);
        return check_launch(ctx, "fdo_step_encode_narrow_tma_kernel");
    }
    if (elem_bytes == 4)
        dk::fdo_step_encode_narrow_kernel<int32_t><<<grid, dk::ENC_THREADS, 0, pick_stream(ctx, stream)>>>(to_params(rng), (uint64_t)n, states, flags, (int32_t*)obs_out, action_out);
    else
        dk::fdo_step_encode_narrow_kernel<uint8_t><<<grid, dk::ENC_THREADS, 0, pick_stream(ctx, stream)>>>(to_params(rng), (uint64_t)n, states, flags, (uint8_t*)obs_out, action_out);
    return check_launch(ctx, "fdo_step_encode_narrow_kernel");
}

dk_status dk_determinize(dk_ctx* ctx, int engine, size_t n_info, size_t samples_per_info, const dk_state* states, const dk_rng* rng,
                         uint64_t* hands_out, uint8_t* reservations_out, uint8_t* status_out, dk_stream stream) {
    if (!ctx || !states || !rng) return DK_ERR_INVALID_ARGUMENT;
    if (engine != DK_FDO && engine != DK_DOKO) return DK_ERR_INVALID_ARGUMENT;
    if (n_info == 0 || samples_per_info == 0) return DK_OK;
    if (samples_per_info > 0xFFFFFFFFull || n_info > 0x7FFFFFFFull) return DK_ERR_INVALID_ARGUMENT;
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, hands_out, 16); DK_ALIGNED(ctx, reservations_out, 4);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    // Blocks per info-state: enough blocks for about 16 waves of 8 resident blocks per SM, at most one block per 128 samples.
    const uint64_t per_info_max = (samples_per_info + dk::MATCH_THREADS - 1) / dk::MATCH_THREADS, want = (uint64_t)ctx->sm_count * 128u;
    uint64_t splits = (want + n_info - 1) / n_info;
    if (splits > per_info_max) splits = per_info_max;
    if (splits < 1 || getenv("DOKO_CUDA_NO_SPLIT")) splits = 1;
    const unsigned grid = (unsigned)(n_info * splits);            // n_info <= 2^31 - 1 and splits > 1 only while n_info * splits <= want + n_info
    if (engine == DK_DOKO) {
        dk::doko_assign_kernel<<<grid, dk::MATCH_THREADS, 0, pick_stream(ctx, stream)>>>(to_params(rng), (uint64_t)n_info, (uint32_t)samples_per_info, (uint32_t)splits,
                                                                                   states, hands_out, reservations_out, status_out);
        return check_launch(ctx, "doko_assign_kernel");
    }
    dk::fdo_determinize_kernel<<<grid, dk::MATCH_THREADS, 0, pick_stream(ctx, stream)>>>(to_params(rng), (uint64_t)n_info, (uint32_t)samples_per_info, (uint32_t)splits,
                                                                                       states, hands_out, reservations_out, status_out);
    return check_launch(ctx, "fdo_determinize_kernel");
}

dk_status dk_leaf_rollouts(dk_ctx* ctx, size_t n_leaves, size_t rollouts_per_leaf, int determinize, const dk_state* states, const dk_rng* rng,
                           int64_t* point_sum_out, dk_stream stream) {
    if (!ctx || !states || !rng || !point_sum_out) return DK_ERR_INVALID_ARGUMENT;
    if (n_leaves == 0) return DK_OK;
    if (rollouts_per_leaf > 0x1000000ull || n_leaves > 0x7FFFFFFFull) return DK_ERR_INVALID_ARGUMENT;   // int32 block sums: |points| < 128
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, point_sum_out, 8);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = pick_stream(ctx, stream);
    // Fewer leaves than one wave of blocks: several blocks per leaf (at most one per 128 rollouts), sums combined by integer atomics.
    const uint64_t per_leaf_max = (rollouts_per_leaf + dk::MATCH_THREADS - 1) / dk::MATCH_THREADS, wave = (uint64_t)ctx->sm_count * DK_LEAF_BLOCKS;
    uint64_t splits = wave / n_leaves;
    if (splits > per_leaf_max) splits = per_leaf_max;
    if (splits < 1 || getenv("DOKO_CUDA_NO_SPLIT")) splits = 1;
    if (splits > 1) DK_CUDA(ctx, cudaMemsetAsync(point_sum_out, 0, n_leaves * 4 * sizeof(int64_t), s));
    const dk::RngParams rp = to_params(rng);
    if (!determinize || rollouts_per_leaf == 0) {
        dk::fdo_leaf_rollouts_kernel<false><<<(unsigned)(n_leaves * splits), dk::MATCH_THREADS, 0, s>>>(rp, 0ull, (uint64_t)n_leaves, (uint32_t)rollouts_per_leaf, (uint32_t)splits, states,
                                                                                                     nullptr, nullptr, nullptr, (long long*)point_sum_out);
        return check_launch(ctx, "fdo_leaf_rollouts_kernel");
    }
    // Determinized rollouts = K3 into a scratch buffer (hands 32 B + reservations 4 B + status 1 B per rollout, at most 256 MiB per
    // chunk of leaves) + the rollout kernel reading its samples: each kernel runs at its own occupancy.
    const size_t R = rollouts_per_leaf, per_leaf = ((R * 37 + 255) & ~(size_t)255) + 512;
    size_t chunk = ((size_t)256 << 20) / per_leaf;
    if (chunk < 1) chunk = 1;
    if (chunk > n_leaves) chunk = n_leaves;
    const size_t b_hands = (chunk * R * 32 + 255) & ~(size_t)255, b_res = (chunk * R * 4 + 255) & ~(size_t)255, b_status = (chunk * R + 255) & ~(size_t)255;
    if (ctx->leaf_ws_bytes < b_hands + b_res + b_status) {
        DK_CUDA(ctx, cudaStreamSynchronize(s));
        if (ctx->leaf_ws) { cudaFree(ctx->leaf_ws); ctx->leaf_ws = nullptr; ctx->leaf_ws_bytes = 0; }
        DK_CUDA(ctx, cudaMalloc(&ctx->leaf_ws, b_hands + b_res + b_status));
        ctx->leaf_ws_bytes = b_hands + b_res + b_status;
    }
    uint64_t* d_hands = (uint64_t*)ctx->leaf_ws;
    uint8_t* d_res = (uint8_t*)ctx->leaf_ws + b_hands;
    uint8_t* d_status = d_res + b_res;
    for (size_t l0 = 0; l0 < n_leaves; l0 += chunk) {
        const size_t nl = n_leaves - l0 < chunk ? n_leaves - l0 : chunk;
        dk::RngParams rc = rp;
        rc.first_id = rp.first_id + l0;                                               // sample unit = first_id + leaf
        const uint64_t det_per_max = (R + dk::MATCH_THREADS - 1) / dk::MATCH_THREADS, want = (uint64_t)ctx->sm_count * 128u;
        uint64_t dsplits = (want + nl - 1) / nl;
        if (dsplits > det_per_max) dsplits = det_per_max;
        if (dsplits < 1 || getenv("DOKO_CUDA_NO_SPLIT")) dsplits = 1;
        dk::fdo_determinize_kernel<<<(unsigned)(nl * dsplits), dk::MATCH_THREADS, 0, s>>>(rc, (uint64_t)nl, (uint32_t)R, (uint32_t)dsplits, states + l0, d_hands, d_res, d_status);
        DK_TRY(check_launch(ctx, "fdo_determinize_kernel"));
        dk::fdo_leaf_rollouts_kernel<true><<<(unsigned)(nl * splits), dk::MATCH_THREADS, 0, s>>>(rp, (uint64_t)l0, (uint64_t)nl, (uint32_t)R, (uint32_t)splits, states, d_hands, d_res, d_status,
                                                                                                (long long*)point_sum_out);
        DK_TRY(check_launch(ctx, "fdo_leaf_rollouts_kernel"));
    }
    return DK_OK;
}

dk_status dk_encode_ipi(dk_ctx* ctx, size_t n, const dk_state* states, const uint64_t* assumed_hands, const uint8_t* assumed_reservations,
                        const uint8_t* next_player, int64_t* out, size_t row_stride, uint8_t* err_out, dk_stream stream) {
    if (!ctx || !states || !assumed_hands || !assumed_reservations || !next_player || !out || row_stride < 311) return DK_ERR_INVALID_ARGUMENT;
    if (n == 0) return DK_OK;
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, assumed_hands, 16); DK_ALIGNED(ctx, assumed_reservations, 4); DK_ALIGNED(ctx, out, 8);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned grid = (unsigned)((n + dk::ENC_THREADS - 1) / dk::ENC_THREADS);
    dk::encode_ipi_kernel<<<grid, dk::ENC_THREADS, 0, pick_stream(ctx, stream)>>>((uint64_t)n, states, assumed_hands, assumed_reservations, next_player, out,
                                                                                 row_stride, err_out);
    return check_launch(ctx, "encode_ipi_kernel");
}

// ---- PIMC move decision ----------------------------------------------------------------------------------
dk_status dk_pimc_evaluate(dk_ctx* ctx, size_t n_roots, size_t n_det, size_t n_rollouts, const dk_state* states, const dk_rng* rng,
                           uint32_t* visits_out, int64_t* value_sum_out, uint8_t* status_out, dk_stream stream) {
    if (!ctx || !states || !rng || (!visits_out && !value_sum_out)) return DK_ERR_INVALID_ARGUMENT;
    if (n_roots == 0 || n_det == 0) return DK_OK;
    // unit_hi of rollout (d, r) is (first_sub + d) * n_rollouts + r: keep it inside 32 bits; int32 block sums: |points| < 128
    if (n_rollouts == 0 || n_rollouts > 0x1000000ull || n_det > 0xFFFFFFFFull ||
        ((uint64_t)rng->first_sub + n_det) * n_rollouts > 0xFFFFFFFFull) return DK_ERR_INVALID_ARGUMENT;
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, visits_out, 4); DK_ALIGNED(ctx, value_sum_out, 8);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = pick_stream(ctx, stream);
    // two kernels (kernels.cuh "N2 in two kernels"): prepare every (root, determinization) at full width into a workspace, then one thread
    // per (root, determinization, rollout); the outputs are accumulated with atomics, so they are zeroed first.  Roots go in chunks that
    // keep the workspace below 1 GiB.
    const size_t per_root = n_det * (dk::PIMC_MAX_LEGAL * sizeof(dk::PimcEntry) + sizeof(uint64_t));
    size_t chunk = ((size_t)1 << 30) / per_root;
    if (chunk < 1) chunk = 1;
    if (chunk > n_roots) chunk = n_roots;
    if (ctx->pimc_ws_bytes < chunk * per_root) {
        if (ctx->pimc_ws) { cudaFree(ctx->pimc_ws); ctx->pimc_ws = nullptr; ctx->pimc_ws_bytes = 0; }
        DK_CUDA(ctx, cudaMalloc(&ctx->pimc_ws, chunk * per_root));
        ctx->pimc_ws_bytes = chunk * per_root;
    }
    if (visits_out) DK_CUDA(ctx, cudaMemsetAsync(visits_out, 0, n_roots * n_det * dk::N_ACTIONS * sizeof(uint32_t), s));
    if (value_sum_out) DK_CUDA(ctx, cudaMemsetAsync(value_sum_out, 0, n_roots * n_det * dk::N_ACTIONS * sizeof(int64_t), s));
    const uint64_t prep_chunks = (n_det + dk::PIMC_PREP_THREADS - 1) / dk::PIMC_PREP_THREADS;
    const uint64_t roll_blocks = (n_det * n_rollouts + dk::PIMC_ROLL_THREADS - 1) / dk::PIMC_ROLL_THREADS;
    if (roll_blocks > 0x7FFFFFFFull) return DK_ERR_INVALID_ARGUMENT;
    for (size_t r0 = 0; r0 < n_roots; r0 += chunk) {
        const size_t nr = n_roots - r0 < chunk ? n_roots - r0 : chunk;
        if (nr * prep_chunks > 0x7FFFFFFFull || nr * roll_blocks > 0x7FFFFFFFull) return DK_ERR_INVALID_ARGUMENT;
        dk::PimcEntry* ws = (dk::PimcEntry*)ctx->pimc_ws;
        uint64_t* masks = (uint64_t*)((char*)ctx->pimc_ws + chunk * n_det * dk::PIMC_MAX_LEGAL * sizeof(dk::PimcEntry));
        dk::pimc_prepare_kernel<<<(unsigned)(nr * prep_chunks), dk::PIMC_PREP_THREADS, 0, s>>>(to_params(rng), (uint64_t)r0, (uint64_t)nr, (uint32_t)n_det, states, ws, masks,
                                                                                           status_out);
        dk_status st = check_launch(ctx, "pimc_prepare_kernel");
        if (st != DK_OK) return st;
        dk::pimc_rollout_kernel<<<(unsigned)(nr * roll_blocks), dk::PIMC_ROLL_THREADS, 0, s>>>(to_params(rng), (uint64_t)r0, (uint32_t)n_det, (uint32_t)n_rollouts,
                                                                                           (uint32_t)roll_blocks, states, ws, masks, visits_out,
                                                                                           (unsigned long long*)value_sum_out);
        st = check_launch(ctx, "pimc_rollout_kernel");
        if (st != DK_OK) return st;
    }
    return DK_OK;
}

dk_status dk_fuse(dk_ctx* ctx, int strategy, size_t n_roots, size_t n_rows, const uint32_t* visits, const uint8_t* status, const uint64_t* allowed,
                  uint8_t* action_out, uint32_t* n_success_out, dk_stream stream) {
    if (!ctx || !visits || !allowed || !action_out || (strategy != DK_FUSE_MAX_N && strategy != DK_FUSE_AVERAGE) || n_rows > 0xFFFFFFFFull)
        return DK_ERR_INVALID_ARGUMENT;
    if (n_roots == 0) return DK_OK;
    if (n_roots > 0x7FFFFFFFull) return DK_ERR_INVALID_ARGUMENT;                // one block per root
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    dk::fuse_kernel<<<(unsigned)n_roots, dk::FUSE_THREADS, 0, pick_stream(ctx, stream)>>>((uint32_t)strategy, (uint64_t)n_roots, (uint32_t)n_rows, visits, status, allowed,
                                                                             action_out, n_success_out);
    return check_launch(ctx, "fuse_kernel");
}

dk_status dk_pimc_root_stats(dk_ctx* ctx, size_t n_roots, size_t n_rows, const uint32_t* visits, const uint8_t* status, const uint64_t* allowed,
                             int64_t* stats, int accumulate, dk_stream stream) {
    if (!ctx || !visits || !allowed || !stats || n_rows > 0xFFFFFFFFull) return DK_ERR_INVALID_ARGUMENT;
    if (n_roots == 0) return DK_OK;
    if (n_roots > 0x7FFFFFFFull) return DK_ERR_INVALID_ARGUMENT;                // one block per root
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    dk::root_stats_kernel<<<(unsigned)n_roots, dk::FUSE_THREADS, 0, pick_stream(ctx, stream)>>>((uint64_t)n_roots, (uint32_t)n_rows, visits, status, allowed,
                                                                                   (long long*)stats, accumulate);
    return check_launch(ctx, "root_stats_kernel");
}

dk_status dk_pimc_pick(dk_ctx* ctx, int strategy, size_t n_roots, const int64_t* stats, const uint64_t* allowed, uint8_t* action_out, dk_stream stream) {
    if (!ctx || !stats || !allowed || !action_out || (strategy != DK_FUSE_MAX_N && strategy != DK_FUSE_AVERAGE)) return DK_ERR_INVALID_ARGUMENT;
    if (n_roots == 0) return DK_OK;
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned grid = (unsigned)((n_roots + dk::STATE_THREADS - 1) / dk::STATE_THREADS);
    dk::root_pick_kernel<<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>((uint32_t)strategy, (uint64_t)n_roots, (const long long*)stats, allowed, action_out);
    return check_launch(ctx, "root_pick_kernel");
}

// ---- AlphaZero self-play driver ----------------------------------------------------------------------------
struct dk_selfplay {
    dk_ctx* ctx = nullptr;
    size_t max_games = 0;
    dk::SpBuffers buf{};
    uint64_t* allowed = nullptr;               // [max_games]
    long long* rows = nullptr;                 // [max_games]
    uint8_t* flags = nullptr;                  // [max_games]
    uint32_t* block_counts = nullptr;          // [n_blocks]
    unsigned long long* block_offsets = nullptr;
    unsigned long long* counters = nullptr;    // count, dropped, unfinished, rows before the turn (two slots used alternately)
    int parity = 0;                            // which of counters[3], counters[4] holds the rows recorded before the next turn
    size_t turn_n = 0;
};

dk_status dk_sp_create(dk_ctx* ctx, size_t max_games, const dk_sp_buffers* bufs, dk_selfplay** out) {
    if (!ctx || !out || !bufs || max_games == 0 || max_games > 0x7FFFFFFFull * dk::SP_THREADS / 2) return DK_ERR_INVALID_ARGUMENT;
    if (!bufs->states || !bufs->policy || !bufs->value || !bufs->player || !bufs->game) return DK_ERR_INVALID_ARGUMENT;
    DK_ALIGNED(ctx, bufs->states, 8); DK_ALIGNED(ctx, bufs->policy, 4); DK_ALIGNED(ctx, bufs->value, 16); DK_ALIGNED(ctx, bufs->game, 4);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    dk_selfplay* sp = new dk_selfplay();
    sp->ctx = ctx; sp->max_games = max_games;
    sp->buf.states = (long long*)bufs->states; sp->buf.policy = bufs->policy; sp->buf.value = bufs->value;
    sp->buf.player = bufs->player; sp->buf.game = bufs->game; sp->buf.capacity = bufs->capacity;
    size_t nb = (max_games + dk::SP_THREADS - 1) / dk::SP_THREADS;
    cudaError_t e = cudaMalloc(&sp->allowed, max_games * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(&sp->rows, max_games * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc(&sp->flags, max_games);
    if (e == cudaSuccess) e = cudaMalloc(&sp->block_counts, nb * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&sp->block_offsets, nb * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&sp->counters, 5 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(sp->counters, 0, 5 * sizeof(unsigned long long));
    if (e != cudaSuccess) { dk_sp_destroy(sp); return fail(ctx, DK_ERR_CUDA, std::string("dk_sp_create: ") + cudaGetErrorString(e)); }
    *out = sp;
    return DK_OK;
}
dk_status dk_sp_destroy(dk_selfplay* sp) {
    if (!sp) return DK_OK;
    cudaSetDevice(sp->ctx->device);
    cudaFree(sp->allowed); cudaFree(sp->rows); cudaFree(sp->flags); cudaFree(sp->block_counts); cudaFree(sp->block_offsets); cudaFree(sp->counters);
    delete sp;
    return DK_OK;
}
dk_status dk_sp_reset(dk_selfplay* sp, dk_stream stream) {
    if (!sp) return DK_ERR_INVALID_ARGUMENT;
    dk_ctx* ctx = sp->ctx;
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    DK_CUDA(ctx, cudaMemsetAsync(sp->counters, 0, 5 * sizeof(unsigned long long), pick_stream(ctx, stream)));
    sp->parity = 0;
    return DK_OK;
}
dk_status dk_sp_begin_turn(dk_selfplay* sp, size_t n, const dk_state* states, uint64_t az_epoch, float keep_prob, uint32_t flags, const dk_rng* rng,
                           dk_stream stream) {
    if (!sp || !states || !rng || n > sp->max_games) return DK_ERR_INVALID_ARGUMENT;
    dk_ctx* ctx = sp->ctx;
    DK_ALIGNED(ctx, states, 16);
    sp->turn_n = n;
    if (n == 0) return DK_OK;
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = pick_stream(ctx, stream);
    unsigned nb = (unsigned)((n + dk::SP_THREADS - 1) / dk::SP_THREADS);
    // one pass: plan + deterministic row numbers (decoupled look-back over block_offsets, ticket in block_counts[0]) + encode into the rows
    DK_CUDA(ctx, cudaMemsetAsync(sp->block_counts, 0, sizeof(uint32_t), s));
    DK_CUDA(ctx, cudaMemsetAsync(sp->block_offsets, 0, nb * sizeof(unsigned long long), s));
    dk::sp_begin_kernel<<<nb, dk::SP_THREADS, 0, s>>>(to_params(rng), (uint64_t)n, states, az_epoch, keep_prob, (flags & DK_SP_SEARCH_FORCED) ? 1u : 0u,
                                                      sp->allowed, sp->flags, sp->rows, sp->buf, sp->block_counts, sp->block_offsets,
                                                      sp->counters + 3 + sp->parity, sp->counters + 3 + (1 - sp->parity), sp->counters,
                                                      sp->counters + 1, !((uintptr_t)sp->buf.states & 31u));
    sp->parity = 1 - sp->parity;
    return check_launch(ctx, "sp_begin_kernel");
}
dk_status dk_sp_turn_view(dk_selfplay* sp, const uint64_t** allowed, const uint8_t** flags, const int64_t** rows) {
    if (!sp) return DK_ERR_INVALID_ARGUMENT;
    if (allowed) *allowed = sp->allowed;
    if (flags) *flags = sp->flags;
    if (rows) *rows = (const int64_t*)sp->rows;
    return DK_OK;
}
dk_status dk_sp_uniform_search(dk_selfplay* sp, const dk_rng* rng, float* policy_out, uint8_t* action_out, dk_stream stream) {
    if (!sp || !rng || !policy_out || !action_out) return DK_ERR_INVALID_ARGUMENT;
    dk_ctx* ctx = sp->ctx;
    size_t n = sp->turn_n;
    if (n == 0) return DK_OK;
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned grid = (unsigned)((n + dk::STATE_THREADS - 1) / dk::STATE_THREADS);
    dk::sp_uniform_search_kernel<<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>(to_params(rng), (uint64_t)n, sp->allowed, sp->flags, policy_out, action_out);
    return check_launch(ctx, "sp_uniform_search_kernel");
}
dk_status dk_sp_end_turn(dk_selfplay* sp, dk_state* states, const float* policy, const uint8_t* action, uint8_t* err_out, dk_stream stream) {
    if (!sp || !states || !policy || !action) return DK_ERR_INVALID_ARGUMENT;
    dk_ctx* ctx = sp->ctx;
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, policy, 4);
    size_t n = sp->turn_n;
    if (n == 0) return DK_OK;
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned grid = (unsigned)((n + dk::STATE_THREADS - 1) / dk::STATE_THREADS);
    dk::sp_apply_kernel<<<grid, dk::STATE_THREADS, 0, pick_stream(ctx, stream)>>>((uint64_t)n, states, sp->allowed, sp->flags, sp->rows, policy, action, sp->buf, err_out);
    return check_launch(ctx, "sp_apply_kernel");
}
dk_status dk_sp_finalize(dk_selfplay* sp, const dk_state* states, dk_stream stream) {
    if (!sp || !states) return DK_ERR_INVALID_ARGUMENT;
    dk_ctx* ctx = sp->ctx;
    DK_ALIGNED(ctx, states, 16);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = pick_stream(ctx, stream);
    DK_CUDA(ctx, cudaMemsetAsync(sp->counters + 2, 0, sizeof(unsigned long long), s));
    dk::sp_finalize_kernel<<<ctx->sm_count * 8, dk::STATE_THREADS, 0, s>>>(sp->counters, states, sp->buf, sp->counters + 2);
    return check_launch(ctx, "sp_finalize_kernel");
}
dk_status dk_sp_counts(dk_selfplay* sp, uint64_t* rows, uint64_t* dropped, uint64_t* unfinished, dk_stream stream) {
    if (!sp) return DK_ERR_INVALID_ARGUMENT;
    dk_ctx* ctx = sp->ctx;
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long h[3];   // count, dropped, unfinished
    cudaStream_t s = pick_stream(ctx, stream);
    DK_CUDA(ctx, cudaMemcpyAsync(h, sp->counters, sizeof h, cudaMemcpyDeviceToHost, s));
    DK_CUDA(ctx, cudaStreamSynchronize(s));
    if (rows) *rows = h[0];
    if (dropped) *dropped = h[1];
    if (unfinished) *unfinished = h[2];
    return DK_OK;
}

dk_status dk_pack_replay_records(dk_ctx* ctx, size_t n_rows, const int64_t* states, const float* value, const float* policy, uint8_t* out, dk_stream stream) {
    if (!ctx || !states || !value || !policy || !out || ((uintptr_t)out & 3u)) return DK_ERR_INVALID_ARGUMENT;
    if (n_rows == 0) return DK_OK;
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    // 16-byte stores and TMA bulk copies need 16-byte aligned buffers; anything else goes word by word
    const bool aligned16 = (((uintptr_t)out | (uintptr_t)states | (uintptr_t)value | (uintptr_t)policy) & 15u) == 0u;
    unsigned long long want = aligned16 ? (n_rows + 3ull) / 4ull : ((unsigned long long)n_rows * dk::REPLAY_WORDS + dk::REPLAY_THREADS - 1) / dk::REPLAY_THREADS;
    unsigned long long cap = (unsigned long long)ctx->sm_count * 32ull;
    unsigned grid = (unsigned)(want < cap ? (want ? want : 1ull) : cap);
    dk::pack_replay_records_kernel<<<grid, dk::REPLAY_THREADS, 0, pick_stream(ctx, stream)>>>((unsigned long long)n_rows, (const long long*)states, value, policy,
                                                                                           (uint32_t*)out, aligned16);
    return check_launch(ctx, "pack_replay_records_kernel");
}

// ---- UCT search -------------------------------------------------------------------------------------------
size_t dk_uct_workspace_bytes(size_t n_trees, size_t iterations) { return (size_t)dk::uct_workspace_bytes(n_trees, iterations); }

dk_status dk_uct_search(dk_ctx* ctx, size_t n_roots, size_t trees_per_root, int determinize, size_t iterations, float uct_exploration_constant,
                        const dk_state* states, const dk_rng* rng, void* workspace, size_t workspace_bytes, uint32_t* visits_out, float* values_out,
                        uint8_t* action_out, uint8_t* status_out, dk_stream stream) {
    if (!ctx || !states || !rng || !workspace || trees_per_root == 0) return DK_ERR_INVALID_ARGUMENT;
    if (n_roots == 0) return DK_OK;
    const size_t n_trees = n_roots * trees_per_root;
    if (iterations == 0 || iterations > dk::UCT_MAX_ITERATIONS || n_trees > 0x7FFFFFFFull * dk::UCT_THREADS / 2 || trees_per_root > 0xFFFFFFFFull ||
        ((uint64_t)rng->first_sub + trees_per_root) * iterations > 0xFFFFFFFFull || workspace_bytes < dk_uct_workspace_bytes(n_trees, iterations))
        return DK_ERR_INVALID_ARGUMENT;
    DK_ALIGNED(ctx, states, 16); DK_ALIGNED(ctx, visits_out, 4); DK_ALIGNED(ctx, values_out, 4);
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = pick_stream(ctx, stream);
    if (ctx->ln_table_len < iterations + 1) {
        // ln via the HOST libm (what Rust's f64::ln calls), so that selection is bit-identical to the CPU reference path
        std::vector<double> t(iterations + 1);
        t[0] = 0.0;
        for (size_t n = 1; n <= iterations; ++n) t[n] = std::log((double)n);
        DK_CUDA(ctx, cudaStreamSynchronize(s));
        if (ctx->d_ln_table) cudaFree(ctx->d_ln_table);
        ctx->d_ln_table = nullptr; ctx->ln_table_len = 0;
        DK_CUDA(ctx, cudaMalloc(&ctx->d_ln_table, t.size() * sizeof(double)));
        DK_CUDA(ctx, cudaMemcpy(ctx->d_ln_table, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
        ctx->ln_table_len = t.size();
    }
    const dk::UctPool P = dk::uct_carve(workspace, n_trees, iterations, getenv("DOKO_CUDA_UCT_TREE_MAJOR") != nullptr);
    const dk::RngParams rp = to_params(rng);
    dk::UctTables T; T.ln = ctx->d_ln_table;
    const unsigned grid = (unsigned)((n_trees + dk::UCT_THREADS - 1) / dk::UCT_THREADS);
    const double c = (double)uct_exploration_constant;
    dk::uct_root_kernel<<<grid, dk::UCT_THREADS, 0, s>>>(rp, P, (uint32_t)trees_per_root, determinize, states, visits_out, values_out);
    DK_TRY(check_launch(ctx, "uct_root_kernel"));
    // One iteration of every tree = two launches (uct.cuh): all SMs run the same phase, each phase has its own register budget.
    // Large batches are cut into parts that iterate on their own streams: trees are independent, both phases wait on memory with a
    // third of the issue slots in use (profiles/r02_uct_*_v6_ncu_summary.json), so one part's walk fills the other's rollout and the
    // launch gaps disappear.  DOKO_CUDA_UCT_PARTS = 1 .. 4 overrides the choice.
    unsigned parts = n_trees >= 65536 ? 3u : 1u;
    if (const char* e = getenv("DOKO_CUDA_UCT_PARTS")) { const int v = atoi(e); if (v >= 1 && v <= 4) parts = (unsigned)v; }
    if ((size_t)parts * 256u > n_trees) parts = 1u;
    for (unsigned p = 1; p < parts; ++p) {
        if (!ctx->uct_streams[p - 1]) DK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->uct_streams[p - 1], cudaStreamNonBlocking));
        if (!ctx->uct_join[p - 1]) DK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->uct_join[p - 1], cudaEventDisableTiming));
    }
    if (parts > 1u) {
        if (!ctx->uct_fork) DK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->uct_fork, cudaEventDisableTiming));
        DK_CUDA(ctx, cudaEventRecord(ctx->uct_fork, s));
        for (unsigned p = 1; p < parts; ++p) DK_CUDA(ctx, cudaStreamWaitEvent(ctx->uct_streams[p - 1], ctx->uct_fork, 0));
    }
    const size_t per_part = ((n_trees + parts - 1) / parts + 255u) & ~(size_t)255u;      // whole blocks of either kernel
    for (uint32_t it = 0; it < (uint32_t)iterations; ++it) {
        for (unsigned p = 0; p < parts; ++p) {
            const size_t t0 = (size_t)p * per_part, t1 = t0 + per_part < n_trees ? t0 + per_part : n_trees;
            if (t0 >= t1) continue;
            cudaStream_t ps = p == 0 ? s : ctx->uct_streams[p - 1];
            const unsigned g_tree = (unsigned)((t1 - t0 + dk::UCT_THREADS - 1) / dk::UCT_THREADS);
            const unsigned g_roll = (unsigned)((t1 - t0 + dk::UCT_ROLLOUT_THREADS - 1) / dk::UCT_ROLLOUT_THREADS);
            dk::uct_tree_kernel<<<g_tree, dk::UCT_THREADS, 0, ps>>>(rp, P, (uint32_t)trees_per_root, (uint32_t)iterations, it, c, T, t0, t1);
            dk::uct_rollout_kernel<<<g_roll, dk::UCT_ROLLOUT_THREADS, 0, ps>>>(rp, P, (uint32_t)trees_per_root, (uint32_t)iterations, it, t0, t1);
            ctx->launches += 2;
        }
    }
    for (unsigned p = 1; p < parts; ++p) {
        DK_CUDA(ctx, cudaEventRecord(ctx->uct_join[p - 1], ctx->uct_streams[p - 1]));
        DK_CUDA(ctx, cudaStreamWaitEvent(s, ctx->uct_join[p - 1], 0));
    }
    dk::uct_moves_kernel<<<grid, dk::UCT_THREADS, 0, s>>>(P, visits_out, values_out, action_out, status_out);
    return check_launch(ctx, "uct kernels");
}

// ---- NCCL (dlopen; the only exchange step of the path) ---------------------------------------------------------------------
namespace {
struct NcclApi {
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, dk_nccl_id, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void* lib = nullptr;
} g_nccl;
std::mutex g_nccl_mu;
dk_status load_nccl(dk_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib) return DK_OK;
    const char* names[] = {getenv("DK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* nm : names) { if (nm && (lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break; }
    if (!lib) return fail(ctx, DK_ERR_NCCL, "libnccl.so.2 not found (set DK_NCCL_LIB)");
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, dk_nccl_id, int))dlsym(lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) { dlclose(lib); return fail(ctx, DK_ERR_NCCL, "NCCL symbols missing"); }
    g_nccl.lib = lib;
    return DK_OK;
}
dk_status nccl_fail(dk_ctx* ctx, const char* what, int rc) {
    return fail(ctx, DK_ERR_NCCL, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error"));
}
}  // namespace

dk_status dk_comm_unique_id(dk_ctx* ctx, dk_nccl_id* out) {
    if (!ctx || !out) return DK_ERR_INVALID_ARGUMENT;
    dk_status st = load_nccl(ctx);
    if (st != DK_OK) return st;
    int rc = g_nccl.GetUniqueId(out);
    return rc ? nccl_fail(ctx, "ncclGetUniqueId", rc) : DK_OK;
}
dk_status dk_comm_init(dk_ctx* ctx, int n_ranks, int rank, const dk_nccl_id* id) {
    if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return DK_ERR_INVALID_ARGUMENT;
    dk_status st = load_nccl(ctx);
    if (st != DK_OK) return st;
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = g_nccl.CommInitRank(&ctx->nccl_comm, n_ranks, *id, rank);
    if (rc) return nccl_fail(ctx, "ncclCommInitRank", rc);
    ctx->nccl_ranks = n_ranks; ctx->nccl_rank = rank;
    return DK_OK;
}
dk_status dk_comm_destroy(dk_ctx* ctx) {
    if (!ctx) return DK_ERR_INVALID_ARGUMENT;
    if (ctx->nccl_comm) { g_nccl.CommDestroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    return DK_OK;
}
// Sum of int64 root statistics over all ranks (ncclAllReduce, ncclInt64 = 4, ncclSum = 0), in place.
dk_status dk_allreduce_root_stats(dk_ctx* ctx, size_t n_values, int64_t* values, dk_stream stream) {
    if (!ctx || !values) return DK_ERR_INVALID_ARGUMENT;
    if (!ctx->nccl_comm) return fail(ctx, DK_ERR_NCCL, "dk_comm_init has not been called");
    if (n_values == 0) return DK_OK;
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = g_nccl.AllReduce(values, values, n_values, 4, 0, ctx->nccl_comm, pick_stream(ctx, stream));
    return rc ? nccl_fail(ctx, "ncclAllReduce", rc) : DK_OK;
}

// Host-buffer playouts.  The batch is cut into chunks; chunk c's kernel runs on the compute stream while chunk c-1's results are
// copied to the host on the copy stream, so the PCIe transfer overlaps the simulation (pinned host buffers make the copies truly
// asynchronous; pageable ones still work).
static dk_status playout_host_impl(dk_ctx* ctx, int engine, uint32_t flags, size_t n, const dk_state* states_host, const dk_rng* rng,
                                   void* points_out_host, void* steps_out_host, uint32_t mode) {
    if (!ctx || !rng || (engine != DK_FDO && engine != DK_DOKO)) return DK_ERR_INVALID_ARGUMENT;
    if (n == 0) return DK_OK;
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pb = mode == dk::OUT_INT32 ? 16 : (mode == dk::OUT_COMPACT ? 4 : 2), sb = mode == dk::OUT_INT32 ? 4 : 1;
    size_t b_states = states_host ? n * sizeof(dk_state) : 0, b_pts = n * pb, b_steps = n * sb;
    b_pts = (b_pts + 255) & ~(size_t)255;
    dk_status st = ensure_scratch(ctx, b_states + b_pts + b_steps);
    if (st != DK_OK) return st;
    if (!ctx->copy_stream) DK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->stream2) DK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    if (!ctx->h2d_done) DK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->h2d_done, cudaEventDisableTiming));
    char* base = (char*)ctx->d_scratch;
    dk_state* d_states = states_host ? (dk_state*)base : nullptr;
    char* d_pts = base + b_states;
    char* d_steps = base + b_states + b_pts;
    if (states_host) {
        DK_CUDA(ctx, cudaMemcpyAsync(d_states, states_host, b_states, cudaMemcpyHostToDevice, ctx->stream));
        DK_CUDA(ctx, cudaEventRecord(ctx->h2d_done, ctx->stream));
        DK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->h2d_done, 0));
    }
    // Chunks of 2^21 games so that the copy of chunk c overlaps the kernel of chunk c + 1; the last 2^21 games are cut into halves down
    // to 2^18, because the copy of the final chunk is the one transfer nothing overlaps.
    const size_t CH = (size_t)1 << 21, CH_MIN = (size_t)1 << 18;
    std::vector<size_t> sizes;
    {
        size_t rem = n;
        while (rem > CH) { sizes.push_back(CH); rem -= CH; }
        for (size_t piece = CH / 2; piece >= CH_MIN && rem > CH_MIN; piece /= 2)
            if (rem > piece) { sizes.push_back(piece); rem -= piece; }
        if (rem) sizes.push_back(rem);
    }
    const size_t n_chunks = sizes.size();
    if (ctx->events.size() < n_chunks) {
        size_t old = ctx->events.size();
        ctx->events.resize(n_chunks);
        for (size_t e = old; e < n_chunks; ++e) DK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->events[e], cudaEventDisableTiming));
    }
    dk::RngParams rp = to_params(rng);
    size_t off = 0;
    for (size_t c = 0; c < n_chunks; ++c) {
        const size_t cnt = sizes[c];
        dk::RngParams rc = rp;
        rc.first_id = rp.first_id + off;
        // chunk kernels alternate between two streams: the blocks of chunk c + 1 fill the SMs that the last wave of chunk c leaves idle
        cudaStream_t cs = (c & 1u) ? ctx->stream2 : ctx->stream;
        st = playout_launch(ctx, engine, flags, cnt, d_states ? d_states + off : nullptr, rc, points_out_host ? d_pts + off * pb : nullptr,
                            steps_out_host ? d_steps + off * sb : nullptr, mode, nullptr, cs);
        if (st != DK_OK) return st;
        DK_CUDA(ctx, cudaEventRecord(ctx->events[c], cs));
        DK_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->events[c], 0));
        if (points_out_host) DK_CUDA(ctx, cudaMemcpyAsync((char*)points_out_host + off * pb, d_pts + off * pb, cnt * pb, cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (steps_out_host) DK_CUDA(ctx, cudaMemcpyAsync((char*)steps_out_host + off * sb, d_steps + off * sb, cnt * sb, cudaMemcpyDeviceToHost, ctx->copy_stream));
        off += cnt;
    }
    DK_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    DK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    DK_CUDA(ctx, cudaStreamSynchronize(ctx->stream2));
    return DK_OK;
}

dk_status dk_playout_host(dk_ctx* ctx, int engine, uint32_t flags, size_t n, const dk_state* states_host, const dk_rng* rng,
                          int32_t* points_out_host, uint32_t* steps_out_host) {
    return playout_host_impl(ctx, engine, flags, n, states_host, rng, points_out_host, steps_out_host, dk::OUT_INT32);
}

dk_status dk_playout_host_compact(dk_ctx* ctx, int engine, uint32_t flags, size_t n, const dk_state* states_host, const dk_rng* rng,
                                  int8_t* points_out_host, uint8_t* steps_out_host) {
    return playout_host_impl(ctx, engine, flags, n, states_host, rng, points_out_host, steps_out_host, dk::OUT_COMPACT);
}

dk_status dk_playout_host_packed(dk_ctx* ctx, int engine, uint32_t flags, size_t n, const dk_state* states_host, const dk_rng* rng,
                                 uint16_t* points_packed_out_host, uint8_t* steps_out_host) {
    return playout_host_impl(ctx, engine, flags, n, states_host, rng, points_packed_out_host, steps_out_host, dk::OUT_PACKED);
}

dk_status dk_playout_summary_host(dk_ctx* ctx, int engine, uint32_t flags, size_t n, const dk_state* states_host, const dk_rng* rng,
                                  dk_playout_stats* stats_out_host) {
    if (!ctx || !rng || !stats_out_host || (engine != DK_FDO && engine != DK_DOKO)) return DK_ERR_INVALID_ARGUMENT;
    DK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t b_states = states_host ? n * sizeof(dk_state) : 0, b_stats = (sizeof(dk_playout_stats) + 255) & ~(size_t)255;
    DK_TRY(ensure_scratch(ctx, b_stats + b_states));
    dk_playout_stats* d_stats = (dk_playout_stats*)ctx->d_scratch;
    dk_state* d_states = states_host ? (dk_state*)((char*)ctx->d_scratch + b_stats) : nullptr;
    if (states_host && n) DK_CUDA(ctx, cudaMemcpyAsync(d_states, states_host, b_states, cudaMemcpyHostToDevice, ctx->stream));
    DK_TRY(dk_playout_summary(ctx, engine, flags, n, d_states, rng, d_stats, 0, ctx->stream));
    DK_CUDA(ctx, cudaMemcpyAsync(stats_out_host, d_stats, sizeof(dk_playout_stats), cudaMemcpyDeviceToHost, ctx->stream));
    DK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DK_OK;
}

}  // extern "C"
