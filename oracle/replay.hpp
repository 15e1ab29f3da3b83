// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/rng.hpp header).
//
// The on-disk experience record of the reference's replay buffer:
//   rs-doko-alpha-zero/src/alpha_zero/net/experience_replay_buffer3.rs:11-20 (DBRecord), :94-121 (append_slice → bincode::serialize)
// DBRecord { state: heapless::Vec<i64, 311>, value: heapless::Vec<f32, 4>, policy: heapless::Vec<f32, 39> } serialized with
// bincode 1.3.3 (Cargo.lock:191-192; third-party, not under /root/reference).  `bincode::serialize` = DefaultOptions with
// fixed-width little-endian integers; a struct is its fields in order with no framing; heapless 0.8.0 (Cargo.lock:919-920)
// serializes a Vec as a sequence = u64 length, then the elements; f32 = its IEEE-754 bits, little endian.
// PARITY UNPINNED: the reference holds no golden bytes for this record; the layout follows the published bincode 1.x format.
// The sled key (a random u64, big endian, :114) and the sled store itself are storage-engine concerns and out of scope.
#pragma once
#include <cstdint>
#include <cstring>

namespace oracle {
namespace replay {

constexpr size_t STATE_DIM = 311, VALUE_DIM = 4, POLICY_DIM = 39;
constexpr size_t RECORD_BYTES = 8 + 8 * STATE_DIM + 8 + 4 * VALUE_DIM + 8 + 4 * POLICY_DIM;   // 2684

inline void put_u64(uint8_t*& p, uint64_t v) { for (int i = 0; i < 8; ++i) *p++ = (uint8_t)(v >> (8 * i)); }
inline void put_u32(uint8_t*& p, uint32_t v) { for (int i = 0; i < 4; ++i) *p++ = (uint8_t)(v >> (8 * i)); }

inline void serialize_record(const int64_t* state, const float* value, const float* policy, uint8_t* out) {
    uint8_t* p = out;
    put_u64(p, STATE_DIM);
    for (size_t i = 0; i < STATE_DIM; ++i) put_u64(p, (uint64_t)state[i]);
    put_u64(p, VALUE_DIM);
    for (size_t i = 0; i < VALUE_DIM; ++i) { uint32_t b; std::memcpy(&b, &value[i], 4); put_u32(p, b); }
    put_u64(p, POLICY_DIM);
    for (size_t i = 0; i < POLICY_DIM; ++i) { uint32_t b; std::memcpy(&b, &policy[i], 4); put_u32(p, b); }
}

}  // namespace replay
}  // namespace oracle
