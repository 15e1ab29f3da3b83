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
#include "doko_state_view.hpp"

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
    // AzEnvState::allowed_actions_by_action_index(is_secondary, epoch) as masks and number_of_allowed_actions(epoch) (either may be null)
    void allowed_actions_by_action_index(bool is_secondary, uint64_t epoch, uint64_t* mask_out_dev, uint8_t* n_allowed_out_dev, dk_stream stream = nullptr) const {
        ctx_.check(dk_legal_mask_az(ctx_.get(), len(), states(), is_secondary ? 1 : 0, epoch, mask_out_dev, n_allowed_out_dev, stream), "dk_legal_mask_az");
    }
    // AzEnvState::id() of every game (last_action_dev may be null = None)
    void id(const uint8_t* last_action_dev, uint64_t* id_out_dev, dk_stream stream = nullptr) const {
        ctx_.check(dk_state_id(ctx_.get(), len(), states(), last_action_dev, id_out_dev, stream), "dk_state_id");
    }
    // FdoAllowedActions::random over the legal set of every game, without playing it
    void random_action(const dk_rng& rng, bool with_announcements, uint8_t* action_out_dev, dk_stream stream = nullptr) const {
        ctx_.check(dk_random_action(ctx_.get(), engine_, len(), states(), &rng, with_announcements ? DK_PLAYOUT_WITH_ANNOUNCEMENTS : 0u, action_out_dev, stream),
                   "dk_random_action");
    }
    // the batch played to the end, only its statistics kept (what an evaluator aggregates); stats_dev in device memory
    void playout_summary(const dk_rng& rng, bool with_announcements, dk_playout_stats* stats_dev, bool accumulate = false, dk_stream stream = nullptr) const {
        ctx_.check(dk_playout_summary(ctx_.get(), engine_, with_announcements ? DK_PLAYOUT_WITH_ANNOUNCEMENTS : 0u, len(), states(), &rng, stats_dev, accumulate ? 1 : 0,
                                      stream), "dk_playout_summary");
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
    // the same rows as int32 (what the reference's Python side narrows them to) or uint8: dense [n][len] (dk_encode_narrow)
    void encode_into_memory_i32(int layout, int32_t* out_dev, dk_stream stream = nullptr) const {
        ctx_.check(dk_encode_narrow(ctx_.get(), layout, 4, len(), states(), out_dev, stream), "dk_encode_narrow");
    }
    void encode_into_memory_u8(int layout, uint8_t* out_dev, dk_stream stream = nullptr) const {
        ctx_.check(dk_encode_narrow(ctx_.get(), layout, 1, len(), states(), out_dev, stream), "dk_encode_narrow");
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

// ONE game by value — the drop-in shape of the reference's env types (McFullDokoEnvState, rs-doko-mcts/src/env/envs/env_state_full_doko.rs:62-220;
// FdoAzEnvState, rs-doko-alpha-zero/src/env/envs/full_doko/full_doko.rs:16-172): a 128-byte record on the host plus the last action,
// every trait method = one C-ABI call on a one-record batch (record up, result down).  Correct but latency bound (≈ 10 µs per call): code that
// drives many games should hold them in an EnvBatch and call the same methods once per lock-step.  Copyable like the reference's states.
class FdoEnvState {
   public:
    FdoEnvState(Context& ctx, const dk_state& record, int last_action = DK_ACTION_NONE) : ctx_(&ctx), rec_(record), last_action_(last_action) {}
    // FdoState::new_game_from_hand_and_start_player (state.rs:125-166)
    static FdoEnvState new_game_from_hand_and_start_player(Context& ctx, const uint64_t hands[4], int start_player) {
        DeviceBuffer<uint64_t> h(4); DeviceBuffer<uint8_t> s(1); DeviceBuffer<dk_state> out(1);
        const uint8_t sp = (uint8_t)start_player;
        copy_in(h.data(), hands, 4 * sizeof(uint64_t)); copy_in(s.data(), &sp, 1);
        ctx.check(dk_from_deals(ctx.get(), DK_FDO, 1, h.data(), s.data(), out.data(), legacy()), "dk_from_deals");
        return FdoEnvState(ctx, out.to_host()[0]);
    }
    const dk_state& record() const { return rec_; }
    FdoStateView view() const { return from_record(rec_); }

    size_t current_player() const { return (rec_.meta >> 2) & 3u; }                       // BOTTOM when finished (full_doko.rs:46-48)
    bool is_terminal() const { return (rec_.meta & 3u) == DK_PHASE_FINISHED; }
    int last_action() const { return last_action_; }                                      // DK_ACTION_NONE = None
    // McEnvState::allowed_actions(first_expansion) (env_state_full_doko.rs:132-172): below the root no calls, and no solo / wedding
    // reservation once any seat has declared a solo
    uint64_t allowed_actions(bool first_expansion) const {
        uint64_t m = legal_mask(false, 1u << 20);
        if (!first_expansion) {
            for (int i = 0; i < rec_.n_reservations && i < 4; ++i) if (rec_.reservations[i] >= DK_RES_DIAMONDS_SOLO) m &= ~(0xFFull << 25);
            m &= ~(0x1Full << 33);
        }
        return m;
    }
    // AzEnvState::allowed_actions_by_action_index(is_secondary, epoch) as action indices; number_of_allowed_actions(epoch)
    std::vector<size_t> allowed_actions_by_action_index(bool is_secondary, uint64_t epoch) const {
        std::vector<size_t> idx;
        for (uint64_t m = legal_mask(is_secondary, epoch); m; m &= m - 1) idx.push_back((size_t)__builtin_ctzll(m));
        return idx;
    }
    size_t number_of_allowed_actions(uint64_t epoch) const { return (size_t)__builtin_popcountll(legal_mask(false, epoch)); }
    // McEnvState::by_action / AzEnvState::take_action_by_action_index(action, skip_single, epoch): a NEW state; an illegal action throws
    // (the reference panics)
    FdoEnvState take_action_by_action_index(size_t action, bool skip_single, uint64_t /*epoch*/ = 0) const {
        DeviceBuffer<dk_state> st(1); DeviceBuffer<uint8_t> a(1), err(1);
        const uint8_t ab = (uint8_t)action;
        copy_in(st.data(), &rec_, sizeof rec_); copy_in(a.data(), &ab, 1);
        ctx_->check(dk_apply(ctx_->get(), DK_FDO, 1, st.data(), a.data(), skip_single ? DK_APPLY_SKIP_SINGLE : 0u, err.data(), legacy()), "dk_apply");
        if (err.to_host()[0]) throw DokoError(DK_ERR_INVALID_ARGUMENT, "take_action_by_action_index: action is not allowed in this state");
        return FdoEnvState(*ctx_, st.to_host()[0], (int)action);
    }
    FdoEnvState by_action(size_t action) const { return take_action_by_action_index(action, false); }
    // rewards_or_none: player_points (McEnvState: as f64; AzEnvState: / 8 as f32); false when the game is not over
    bool rewards_or_none(double out[4]) const {
        if (!is_terminal()) return false;
        for (int p = 0; p < 4; ++p) out[p] = (double)rec_.points[p];
        return true;
    }
    bool rewards_or_none(float out[4]) const {
        if (!is_terminal()) return false;
        for (int p = 0; p < 4; ++p) out[p] = (float)rec_.points[p] / 8.0f;
        return true;
    }
    // McEnvState::random_rollout: the _no_announcement policy to the end of the game on the stream (rng.seed, rng.first_id, rng.epoch)
    void random_rollout(const dk_rng& rng, double out[4]) const {
        int32_t pts[4];
        ctx_->check(dk_playout_host(ctx_->get(), DK_FDO, 0u, 1, &rec_, &rng, pts, nullptr), "dk_playout_host");
        for (int p = 0; p < 4; ++p) out[p] = (double)pts[p];
    }
    // AzEnvState::encode_into_memory(&mut [i64; 311])
    void encode_into_memory(int64_t memory[DK_OBS_LEN_FDO_PI311]) const {
        DeviceBuffer<dk_state> st(1); DeviceBuffer<int64_t> obs(DK_OBS_LEN_FDO_PI311);
        copy_in(st.data(), &rec_, sizeof rec_);
        ctx_->check(dk_encode(ctx_->get(), DK_LAYOUT_FDO_PI311, 1, st.data(), obs.data(), DK_OBS_LEN_FDO_PI311, legacy()), "dk_encode");
        const std::vector<int64_t> h = obs.to_host();
        for (int i = 0; i < DK_OBS_LEN_FDO_PI311; ++i) memory[i] = h[i];
    }
    // AzEnvState::id()
    uint64_t id() const {
        DeviceBuffer<dk_state> st(1); DeviceBuffer<uint8_t> a(1); DeviceBuffer<uint64_t> out(1);
        const uint8_t ab = (uint8_t)last_action_;
        copy_in(st.data(), &rec_, sizeof rec_); copy_in(a.data(), &ab, 1);
        ctx_->check(dk_state_id(ctx_->get(), 1, st.data(), a.data(), out.data(), legacy()), "dk_state_id");
        return out.to_host()[0];
    }
    bool operator==(const FdoEnvState& o) const { return std::memcmp(&rec_, &o.rec_, sizeof rec_) == 0 && last_action_ == o.last_action_; }

   private:
    // All calls of this class run on the legacy default stream, the stream cudaMemcpy uses: record up, kernel, result down are ordered
    // without an explicit synchronisation.
    static dk_stream legacy() { return (dk_stream)cudaStreamLegacy; }
    static void copy_in(void* dst_dev, const void* src, size_t bytes) {
        if (cudaMemcpy(dst_dev, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) throw DokoError(DK_ERR_CUDA, "cudaMemcpy H2D failed");
    }
    uint64_t legal_mask(bool is_secondary, uint64_t epoch) const {
        DeviceBuffer<dk_state> st(1); DeviceBuffer<uint64_t> m(1);
        copy_in(st.data(), &rec_, sizeof rec_);
        ctx_->check(dk_legal_mask_az(ctx_->get(), 1, st.data(), is_secondary ? 1 : 0, epoch, m.data(), nullptr, legacy()), "dk_legal_mask_az");
        return m.to_host()[0];
    }
    Context* ctx_;
    dk_state rec_;
    int last_action_;
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
