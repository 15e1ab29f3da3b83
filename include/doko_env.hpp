// doko_env.hpp — C++ host-side mirror of the reference's game-state seam over the C ABI (doko_cuda.h).
//
// The reference's hosts are Rust traits over by-value states:
//   McEnvState  (rs-doko-mcts/src/env/env_state.rs:7-28)   current_player, is_terminal, allowed_actions, by_action,
//                                                           rewards_or_none, random_rollout
//   AzEnvState  (rs-doko-alpha-zero/src/env/env_state.rs:4-42) encode_into_memory, allowed_actions_by_action_index,
//                                                           take_action_by_action_index(action, skip_single, epoch), rewards_or_none
// Rust is not available in the build image, so this header gives the same surface in C++ for a BATCH of games that lives in
// device memory (one dk_state record per game).  Method names and argument meaning follow the traits; every method is one C-ABI
// call on caller-visible device buffers.  Errors: the reference panics; here a DokoError is thrown with dk_last_error().
// Header-only; link against libdoko_cuda.so and the CUDA runtime (for cudaMalloc/cudaMemcpy).
#pragma once
#include <cuda_runtime_api.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "doko_cuda.h"

namespace doko {

struct DokoError : std::runtime_error {
    dk_status status;
    DokoError(dk_status st, const std::string& what) : std::runtime_error(what), status(st) {}
};

class Context {
   public:
    explicit Context(int device = 0) {
        dk_status st = dk_init(device, &ctx_);
        if (st != DK_OK) throw DokoError(st, "dk_init failed: an sm_100 GPU is required (no CPU fallback)");
    }
    ~Context() { if (ctx_) dk_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    dk_ctx* get() const { return ctx_; }
    void check(dk_status st, const char* what) const {
        if (st != DK_OK) throw DokoError(st, std::string(what) + ": " + dk_last_error(ctx_));
    }

   private:
    dk_ctx* ctx_ = nullptr;
};

template <class T>
class DeviceBuffer {
   public:
    DeviceBuffer() = default;
    explicit DeviceBuffer(size_t n) { resize(n); }
    ~DeviceBuffer() { if (p_) cudaFree(p_); }
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    void resize(size_t n) {
        if (p_) cudaFree(p_);
        p_ = nullptr; n_ = n;
        if (n && cudaMalloc(&p_, n * sizeof(T)) != cudaSuccess) throw DokoError(DK_ERR_CUDA, "cudaMalloc failed");
    }
    T* data() const { return static_cast<T*>(p_); }
    size_t size() const { return n_; }
    std::vector<T> to_host() const {
        std::vector<T> h(n_);
        if (n_ && cudaMemcpy(h.data(), p_, n_ * sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess) throw DokoError(DK_ERR_CUDA, "cudaMemcpy D2H failed");
        return h;
    }
    void from_host(const std::vector<T>& h) {
        resize(h.size());
        if (n_ && cudaMemcpy(p_, h.data(), n_ * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) throw DokoError(DK_ERR_CUDA, "cudaMemcpy H2D failed");
    }

   private:
    void* p_ = nullptr;
    size_t n_ = 0;
};

// A batch of games in device memory — the batched counterpart of McFullDokoEnvState / FdoAzEnvState (engine DK_FDO)
// and McDokoEnvState (engine DK_DOKO).
class EnvBatch {
   public:
    EnvBatch(Context& ctx, int engine, size_t n) : ctx_(ctx), engine_(engine), states_(n) {}

    size_t len() const { return states_.size(); }
    dk_state* states() const { return states_.data(); }

    // FdoState::new_game / DoState::new_game for every game of the batch
    void new_games(const dk_rng& rng, dk_stream stream = nullptr) {
        ctx_.check(dk_new_games(ctx_.get(), engine_, len(), &rng, states(), stream), "dk_new_games");
    }
    // McEnvState::allowed_actions(first_expansion = true) / AzEnvState::allowed_actions_by_action_index as bit masks
    void allowed_actions(uint64_t* mask_out_dev, dk_stream stream = nullptr) const {
        ctx_.check(dk_legal_mask(ctx_.get(), engine_, len(), states(), mask_out_dev, stream), "dk_legal_mask");
    }
    // McEnvState::by_action / AzEnvState::take_action_by_action_index(action, skip_single, _), in place
    void take_action_by_action_index(const uint8_t* action_dev, bool skip_single, uint8_t* err_out_dev = nullptr, dk_stream stream = nullptr) {
        ctx_.check(dk_apply(ctx_.get(), engine_, len(), states(), action_dev, skip_single ? DK_APPLY_SKIP_SINGLE : 0u, err_out_dev, stream), "dk_apply");
    }
    // is_terminal + rewards_or_none (points; the MCTS env uses them as f64, the AZ env divides by 8)
    void rewards_or_none(uint8_t* done_out_dev, int32_t* points_out_dev, dk_stream stream = nullptr) const {
        ctx_.check(dk_terminal(ctx_.get(), engine_, len(), states(), done_out_dev, points_out_dev, stream), "dk_terminal");
    }
    // AzEnvState::encode_into_memory for the whole batch (row-major [n][row_stride] i64)
    void encode_into_memory(int layout, int64_t* out_dev, size_t row_stride, dk_stream stream = nullptr) const {
        ctx_.check(dk_encode(ctx_.get(), layout, len(), states(), out_dev, row_stride, stream), "dk_encode");
    }
    // McEnvState::random_rollout for every game (no-announcement policy unless with_announcements)
    void random_rollout(const dk_rng& rng, bool with_announcements, int32_t* points_out_dev, uint32_t* steps_out_dev = nullptr, dk_stream stream = nullptr) const {
        ctx_.check(dk_playout(ctx_.get(), engine_, with_announcements ? DK_PLAYOUT_WITH_ANNOUNCEMENTS : 0u, len(), states(), &rng, points_out_dev, steps_out_dev, stream),
                   "dk_playout");
    }
    // CAPSampling::sample (card_matching) — S hidden-hand samples per game
    void sample(const dk_rng& rng, size_t samples, uint64_t* hands_out_dev, uint8_t* reservations_out_dev, uint8_t* status_out_dev, dk_stream stream = nullptr) const {
        ctx_.check(dk_determinize(ctx_.get(), engine_, len(), samples, states(), &rng, hands_out_dev, reservations_out_dev, status_out_dev, stream), "dk_determinize");
    }

   private:
    Context& ctx_;
    int engine_;
    DeviceBuffer<dk_state> states_;
};

}  // namespace doko
