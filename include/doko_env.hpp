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

    // DefaultImpiPolicy::execute up to the fuse (compare_impi.rs:212-345): S determinizations x flat Monte-Carlo per-sample policy;
    // visits_out [n][S][39], status_out [n][S]
    void pimc_evaluate(const dk_rng& rng, size_t n_det, size_t n_rollouts, uint32_t* visits_out_dev, int64_t* value_sum_out_dev, uint8_t* status_out_dev,
                       dk_stream stream = nullptr) const {
        ctx_.check(dk_pimc_evaluate(ctx_.get(), len(), n_det, n_rollouts, states(), &rng, visits_out_dev, value_sum_out_dev, status_out_dev, stream), "dk_pimc_evaluate");
    }

    // EvFullDokoMCTSPolicy::evaluate for every game (mcts_policy.rs:76-118): one UCT tree per (game, d); determinize = CAPSampling first.
    // workspace: dk_uct_workspace_bytes(len() * trees_per_game, iterations) bytes of device memory.
    void monte_carlo_tree_search(const dk_rng& rng, size_t trees_per_game, bool determinize, size_t iterations, float uct_exploration_constant, void* workspace_dev,
                                 size_t workspace_bytes, uint32_t* visits_out_dev, float* values_out_dev, uint8_t* action_out_dev, uint8_t* status_out_dev,
                                 dk_stream stream = nullptr) const {
        ctx_.check(dk_uct_search(ctx_.get(), len(), trees_per_game, determinize ? 1 : 0, iterations, uct_exploration_constant, states(), &rng, workspace_dev,
                                 workspace_bytes, visits_out_dev, values_out_dev, action_out_dev, status_out_dev, stream), "dk_uct_search");
    }
    // encode_state_ipi for the whole batch (guessed hands / reservations by absolute seat, seat to guess for)
    void encode_ipi(const uint64_t* assumed_hands_dev, const uint8_t* assumed_reservations_dev, const uint8_t* next_player_dev, int64_t* out_dev, size_t row_stride,
                    uint8_t* err_out_dev = nullptr, dk_stream stream = nullptr) const {
        ctx_.check(dk_encode_ipi(ctx_.get(), len(), states(), assumed_hands_dev, assumed_reservations_dev, next_player_dev, out_dev, row_stride, err_out_dev, stream),
                   "dk_encode_ipi");
    }

   private:
    Context& ctx_;
    int engine_;
    DeviceBuffer<dk_state> states_;
};

// PolicyFusionFn::fuse (policy_fusion.rs:9-16) for a batch of roots; strategy = DK_FUSE_MAX_N (PolicyFusionMaxN) or DK_FUSE_AVERAGE
// (PolicyFusionAverageStrategy).  action_out[i] == DK_ACTION_NONE when no sample of root i succeeded.
inline void fuse(Context& ctx, int strategy, size_t n_roots, size_t n_rows, const uint32_t* visits_dev, const uint8_t* status_dev, const uint64_t* allowed_dev,
                 uint8_t* action_out_dev, uint32_t* n_success_out_dev = nullptr, dk_stream stream = nullptr) {
    ctx.check(dk_fuse(ctx.get(), strategy, n_roots, n_rows, visits_dev, status_dev, allowed_dev, action_out_dev, n_success_out_dev, stream), "dk_fuse");
}

// self_play (rs-doko-alpha-zero/src/alpha_zero/train/self_play.rs:19-207) for an EnvBatch in lock-step.  The experience buffers
// take the place of states_buffer / policy_targets_buffer / value_targets_buffer and stay in device memory.
class SelfPlay {
   public:
    SelfPlay(Context& ctx, size_t max_games, size_t capacity)
        : ctx_(ctx), states_(capacity * 311), policy_(capacity * DK_N_ACTIONS), value_(capacity * 4), player_(capacity), game_(capacity) {
        dk_sp_buffers b{states_.data(), policy_.data(), value_.data(), player_.data(), game_.data(), capacity};
        ctx_.check(dk_sp_create(ctx_.get(), max_games, &b, &sp_), "dk_sp_create");
    }
    ~SelfPlay() { dk_sp_destroy(sp_); }
    SelfPlay(const SelfPlay&) = delete;
    SelfPlay& operator=(const SelfPlay&) = delete;

    void reset(dk_stream stream = nullptr) { ctx_.check(dk_sp_reset(sp_, stream), "dk_sp_reset"); }
    // is_terminal / number_of_allowed_actions / encode_into_memory / current_player of every game; rows assigned
    void begin_turn(const EnvBatch& env, uint64_t az_epoch, float probability_of_keeping_experience, const dk_rng& rng, uint32_t flags = 0, dk_stream stream = nullptr) {
        ctx_.check(dk_sp_begin_turn(sp_, env.len(), env.states(), az_epoch, probability_of_keeping_experience, flags, &rng, stream), "dk_sp_begin_turn");
    }
    void turn_view(const uint64_t** allowed, const uint8_t** flags, const int64_t** rows) { ctx_.check(dk_sp_turn_view(sp_, allowed, flags, rows), "dk_sp_turn_view"); }
    // policy target + take_action_by_action_index(action, false, epoch)
    void end_turn(EnvBatch& env, const float* policy_dev, const uint8_t* action_dev, uint8_t* err_out_dev = nullptr, dk_stream stream = nullptr) {
        ctx_.check(dk_sp_end_turn(sp_, env.states(), policy_dev, action_dev, err_out_dev, stream), "dk_sp_end_turn");
    }
    void finalize(const EnvBatch& env, dk_stream stream = nullptr) { ctx_.check(dk_sp_finalize(sp_, env.states(), stream), "dk_sp_finalize"); }
    uint64_t rows(dk_stream stream = nullptr) { uint64_t r = 0; ctx_.check(dk_sp_counts(sp_, &r, nullptr, nullptr, stream), "dk_sp_counts"); return r; }
    const DeviceBuffer<int64_t>& states_buffer() const { return states_; }
    const DeviceBuffer<float>& policy_targets_buffer() const { return policy_; }
    const DeviceBuffer<float>& value_targets_buffer() const { return value_; }
    dk_selfplay* get() const { return sp_; }

   private:
    Context& ctx_;
    DeviceBuffer<int64_t> states_;
    DeviceBuffer<float> policy_, value_;
    DeviceBuffer<uint8_t> player_;
    DeviceBuffer<uint32_t> game_;
    dk_selfplay* sp_ = nullptr;
};

}  // namespace doko
