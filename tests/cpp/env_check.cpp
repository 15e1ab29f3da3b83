// Drives include/doko_env.hpp (the C++ host-side mirror of the reference's env traits) on the GPU and writes what it saw to a file;
// tests/test_gpu_cpp_env.py replays the same games on the oracle and compares.  Built by __graft_entry__.build() (g++, links
// libdoko_cuda.so and libcudart).
//   env_check <out.bin> [n_games]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "doko_env.hpp"

template <class T> static void put(FILE* f, const std::vector<T>& v) { std::fwrite(v.data(), sizeof(T), v.size(), f); }

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const size_t n = argc > 2 ? (size_t)std::atol(argv[2]) : 4096;
    const uint64_t SEED = 0xD0C05EEDull;
    try {
        doko::Context ctx(0);
        // ---- batch: new_game, random_rollout (with announcements), encode_into_memory, allowed actions --------------------------
        doko::EnvBatch env(ctx, DK_FDO, n);
        const dk_rng deal{SEED, 1000, 3, 0}, roll{SEED, 1000, 4, 0};
        env.new_games(deal);
        doko::DeviceBuffer<int32_t> pts(n * 4);
        doko::DeviceBuffer<uint32_t> steps(n);
        doko::DeviceBuffer<int64_t> obs(n * 311);
        doko::DeviceBuffer<uint64_t> mask(n);
        doko::DeviceBuffer<dk_playout_stats> stats(1);
        env.random_rollout(roll, true, pts.data(), steps.data());
        env.encode_into_memory(DK_LAYOUT_FDO_PI311, obs.data(), 311);
        env.allowed_actions(mask.data());
        env.playout_summary(roll, true, stats.data());
        ctx.check(dk_synchronize(ctx.get(), nullptr), "dk_synchronize");
        FILE* f = std::fopen(argv[1], "wb");
        if (!f) return 2;
        const uint64_t hdr[2] = {n, 0};
        std::fwrite(hdr, sizeof hdr, 1, f);
        put(f, pts.to_host()); put(f, steps.to_host()); put(f, obs.to_host()); put(f, mask.to_host());
        std::vector<dk_state> host(n);
        if (cudaMemcpy(host.data(), env.states(), n * sizeof(dk_state), cudaMemcpyDeviceToHost) != cudaSuccess) return 3;
        put(f, host);
        put(f, stats.to_host());
        // ---- one game by value: the McEnvState / AzEnvState method set, action by action -------------------------------------
        doko::FdoEnvState s(ctx, host[0]);
        std::vector<int64_t> walk;       // per step: action, current_player before, number_of_allowed(epoch 0), number_of_allowed(epoch 10), id lo 32, id hi 32
        std::vector<dk_state> after;
        std::vector<int64_t> tokens;
        for (int k = 0; !s.is_terminal() && k < 400; ++k) {
            const std::vector<size_t> idx = s.allowed_actions_by_action_index(false, 10);
            const std::vector<size_t> idx_young = s.allowed_actions_by_action_index(false, 0);
            const size_t a = idx[(size_t)(k * 7 + 3) % idx.size()];
            const uint64_t id = s.id();
            walk.push_back((int64_t)a); walk.push_back((int64_t)s.current_player()); walk.push_back((int64_t)idx_young.size());
            walk.push_back((int64_t)s.number_of_allowed_actions(10)); walk.push_back((int64_t)(id & 0xFFFFFFFFull)); walk.push_back((int64_t)(id >> 32));
            walk.push_back((int64_t)s.allowed_actions(true)); walk.push_back((int64_t)s.allowed_actions(false));
            doko::FdoEnvState t = s.take_action_by_action_index(a, false, 10);
            if (!(t.last_action() == (int)a) || t == s) return 4;
            s = t;
            after.push_back(s.record());
            if (k % 9 == 0) { int64_t mem[311]; s.encode_into_memory(mem); tokens.insert(tokens.end(), mem, mem + 311); }
        }
        double rew[4] = {0, 0, 0, 0}; float rew8[4] = {0, 0, 0, 0};
        if (!s.rewards_or_none(rew) || !s.rewards_or_none(rew8)) return 5;
        // an illegal action must throw and leave the state usable
        bool threw = false;
        try { doko::FdoEnvState(ctx, host[1]).take_action_by_action_index(0, false); } catch (const doko::DokoError&) { threw = true; }
        if (!threw) return 6;
        double roll_out[4];
        doko::FdoEnvState(ctx, host[2]).random_rollout(dk_rng{SEED, 77, 5, 0}, roll_out);
        const uint64_t counts[3] = {walk.size() / 8, after.size(), tokens.size() / 311};
        std::fwrite(counts, sizeof counts, 1, f);
        put(f, walk); put(f, after); put(f, tokens);
        std::fwrite(rew, sizeof rew, 1, f); std::fwrite(rew8, sizeof rew8, 1, f); std::fwrite(roll_out, sizeof roll_out, 1, f);
        std::fclose(f);
        std::printf("ok %zu games, %zu by-value steps\n", n, after.size());
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "env_check: %s\n", e.what());
        return 1;
    }
}
