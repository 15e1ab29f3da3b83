// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by, or
// called from the product path (master_doko_reinforcement_learning_b200/, include/).  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg use it.
//
// Random number sources for the CPU restatement of the reference's rules.
//
// The reference draws every random decision from `rand 0.9.0`'s `SmallRng`
// (Cargo.lock:1899-1900; e.g. rs-full-doko/src/state/state.rs:169-178).  That crate is a
// third-party dependency which is NOT under /root/reference, so two back-ends exist:
//
//  * SmallRngStream  — an emulation of rand 0.9.0 SmallRng (Xoshiro256++ seeded by SplitMix64,
//    the `UniformInt::sample_single_inclusive` widening-multiply sampler, and the 0.9
//    `IncreasingUniform` shuffle).  It exists ONLY to replay the reference's own SmallRng-seeded
//    golden vectors (rs-doko/src/hand/hand_random.rs:68-83, rs-full-doko/src/hand/hand.rs:559-585,
//    rs-doko/src/state/state.rs:426-447, rs-doko/src/util/bitflag/bitflag.rs:190-206,
//    rs-doko-embeddings/src/encode_state.rs:349-597).  Because those five vectors reproduce
//    (tests/test_oracle_smallrng.py) the emulation is pinned.
//
//  * PhiloxStream — the parity stream shared with the CUDA kernels.  Philox4x32-10, counter =
//    (unit_lo, unit_hi, site<<16 | block, epoch), key = 64-bit seed.  Every reference RNG call
//    site is a "site"; its k-th use inside one unit (game / sample / rollout) takes word k of
//    that site's stream and maps it with idx = (uint64(w) * n) >> 32 (announcement decisions, which
//    are two-way, take one bit each; the deal draws three times from a word and the four card draws
//    of a trick share the trick's word — `chain` below).  See DESIGN.md "Philox parity contract".
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>

namespace oracle {

// ---------------------------------------------------------------------------------------------
// Call sites of the reference that consume randomness (one Philox sub-stream each).
// ---------------------------------------------------------------------------------------------
enum Site : uint32_t {
    SITE_DEAL = 0,         // draw 0 = start player, draws 1..47 = the shuffle; draw s takes word min(s / 3, 11) for s <= 36 (chained)
                           //   rs-full-doko/src/state/state.rs:169-178, hand/hand.rs:174-188
                           //   rs-doko/src/state/state.rs:159-168,     hand/hand_random.rs:44-60
    SITE_RESERVATION = 1,  // allowed.random() in the reservation phase; word = #reservations made
    SITE_ANNOUNCEMENT = 2, // allowed.random() in the announcement phase; word = k-th decision of the call
    SITE_CARD = 3,         // allowed.random() in the card phase; word = card_index / 4 (the trick), its four draws chained
    SITE_MATCH_CARD = 4,   // card_matching rule 4 `.choose(rng)`  (card_matching.rs:182-186); k-th rule-4 use = word k / 3, chained
    SITE_MATCH_RESERVATION = 5, // hidden reservation `.choose(rng)` (card_matching.rs:450); word = seat index
    SITE_ASSIGN = 6,       // rs-doko-assignment random card / random player (assignment.rs:419-445): one word per card, the seat chained
    SITE_STEP = 7,         // lock-step env step (config 5): one decision per call, word 0
    SITE_KEEP = 8,         // self_play's `rng.gen::<f32>() < probability_of_keeping_experience` (self_play.rs:88); word 0 of the turn's epoch
    SITE_EXPAND = 9,       // MCTS expand_single's `unexpanded_actions.random(rng)` (rs-doko-mcts/src/mcts/mcts.rs:78-80); word 0 of the iteration's unit
    SITE_COUNT = 10,
};

// Abstract source.  `below(site, n)` returns a value in [0, n).
struct Rng {
    virtual ~Rng() = default;
    virtual uint32_t below(Site site, uint32_t n) = 0;
    // Shuffle of the 48-card deal array (reference: `cards_to_distribute.shuffle(&mut rng)`).
    virtual void shuffle48(uint8_t* cards) = 0;
    // `IteratorRandom::choose` over an iterator WITHOUT an exact size_hint (FdoHandIter,
    // rs-full-doko/src/matching/card_matching.rs:182-186).  `count` = number of elements.
    virtual uint32_t choose_unsized(Site site, uint32_t count) = 0;
    // Random start player: `rng.random_range(0..4)` / `gen_range(0..4)`.
    virtual uint32_t start_player() = 0;
    // A further draw from the word the LAST below(site, ..) took (chained draws, PhiloxStream::chain); a sequential generator just draws again.
    virtual uint32_t below_chained(Site site, uint32_t n) { return below(site, n); }
    // Position the sub-stream of `site` at word `ordinal` (no-op for a sequential generator).
    virtual void set_ordinal(Site, uint32_t) {}
    // Position the card-phase sub-stream at `card_index`; inside a trick `chain_mul` is the product of the numbers of legal card types
    // the plays already made in that trick chose from (State::card_chain_mul()) — where the trick's chained draws continue.
    virtual void set_card_position(uint32_t card_index, uint32_t chain_mul) { (void)chain_mul; set_ordinal(SITE_CARD, card_index); }
};

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants)
// ---------------------------------------------------------------------------------------------
struct Philox4x32 {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;

    static inline void block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
        uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
        uint32_t k0 = key[0], k1 = key[1];
        for (int r = 0; r < 10; ++r) {
            uint64_t p0 = (uint64_t)M0 * c0;
            uint64_t p1 = (uint64_t)M1 * c2;
            uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
            uint32_t n1 = (uint32_t)p1;
            uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
            uint32_t n3 = (uint32_t)p0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += W0; k1 += W1;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

// Word `ordinal` of sub-stream `site` for unit (unit_lo, unit_hi) in epoch `epoch`.
inline uint32_t philox_word(uint64_t seed, uint32_t unit_lo, uint32_t unit_hi, uint32_t epoch,
                            uint32_t site, uint32_t ordinal) {
    uint32_t ctr[4] = {unit_lo, unit_hi, (site << 16) | (ordinal >> 2), epoch};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    Philox4x32::block(ctr, key, out);
    return out[ordinal & 3];
}

inline uint32_t mul_shift(uint32_t w, uint32_t n) { return (uint32_t)(((uint64_t)w * n) >> 32); }

struct PhiloxStream final : Rng {
    uint64_t seed;
    uint32_t unit_lo, unit_hi, epoch;
    uint32_t ordinal[SITE_COUNT];
    // Chained draws.  (idx, rest) = (high, low) half of v * n: idx is the draw over n choices, the low half is the fractional part of
    // v * n / 2^32 — uniform again, on a grid of spacing n / 2^32 — and feeds the next draw of the chain.  After the draws n_0..n_{k-1}
    // the word has become word * n_0 * .. * n_{k-1} mod 2^32; the relative bias of draw k is below n_0 * .. * n_k / 2^32
    // (deal: 48*47*46 / 2^32 = 2.4e-5; a trick: 12^4 / 2^32 = 4.8e-6).
    uint32_t deal_word = 0xFFFFFFFFu, deal_v = 0;     // SITE_DEAL: word of the running chain and what is left of it
    uint32_t card_trick = 0xFFFFFFFFu, card_v = 0;    // SITE_CARD: trick of the running chain and what is left of its word
    uint32_t card_mul = 1;                            // factor for the first draw after set_card_position (resume inside a trick)
    uint32_t match_word = 0xFFFFFFFFu, match_v = 0;   // SITE_MATCH_CARD: three chained draws per word
    uint32_t last_v[SITE_COUNT] = {};                 // what the last one-word draw of a site left of its word (below_chained)
    static uint32_t chain(uint32_t& v, uint32_t n) { uint64_t p = (uint64_t)v * n; v = (uint32_t)p; return (uint32_t)(p >> 32); }

    PhiloxStream(uint64_t seed_, uint32_t unit_lo_, uint32_t unit_hi_ = 0, uint32_t epoch_ = 0)
        : seed(seed_), unit_lo(unit_lo_), unit_hi(unit_hi_), epoch(epoch_) {
        std::memset(ordinal, 0, sizeof ordinal);
    }
    uint32_t word(Site site) {
        return philox_word(seed, unit_lo, unit_hi, epoch, site, ordinal[site]++);
    }
    uint32_t below(Site site, uint32_t n) override {
        if (site == SITE_ANNOUNCEMENT && n <= 2) {
            // Announcement decisions are draws over {NoAnnouncement, call} (n = 2) or {NoAnnouncement} (n = 1): they take ONE BIT
            // each — decision k is bit (k & 31) of word (k >> 5) of the site's stream — instead of a whole word.
            uint32_t k = ordinal[site]++;
            uint32_t w = philox_word(seed, unit_lo, unit_hi, epoch, site, k >> 5);
            return n == 2 ? ((w >> (k & 31)) & 1u) : 0u;
        }
        if (site == SITE_DEAL) {
            // Draw s of the deal (0 = start seat, s = 1..47 the shuffle step i = 48 - s) takes word s / 3: three chained draws per
            // word.  The twelfth word (index 11) serves the four draws 33..36, so that the draws which decide the hands (the steps
            // i = 47..12; the steps 11..1 only permute seat 0's own positions) come from exactly three Philox blocks.
            uint32_t k = ordinal[site]++;
            uint32_t w = k < 33 ? k / 3 : (k < 37 ? 11u : 12u + (k - 37) / 3);
            if (w != deal_word) { deal_word = w; deal_v = philox_word(seed, unit_lo, unit_hi, epoch, site, w); }
            return chain(deal_v, n);
        }
        if (site == SITE_CARD) {
            // The draw at card_index ci belongs to trick ci / 4, which owns word ci / 4 of the site; the trick's draws are chained.
            uint32_t ci = ordinal[site]++;
            if ((ci >> 2) != card_trick) {
                card_trick = ci >> 2;
                card_v = philox_word(seed, unit_lo, unit_hi, epoch, site, ci >> 2) * card_mul;
            }
            card_mul = 1;
            return chain(card_v, n);
        }
        if (site == SITE_MATCH_CARD) {
            // rule 4's k-th use takes word k / 3: three chained draws per word (bias <= 36*35*34 / 2^32 = 1e-5)
            uint32_t k = ordinal[site]++;
            if (k / 3 != match_word) { match_word = k / 3; match_v = philox_word(seed, unit_lo, unit_hi, epoch, site, k / 3); }
            return chain(match_v, n);
        }
        last_v[site] = word(site);
        return chain(last_v[site], n);
    }
    uint32_t below_chained(Site site, uint32_t n) override { return chain(last_v[site], n); }
    void set_ordinal(Site site, uint32_t o) override {
        ordinal[site] = o;
        if (site == SITE_MATCH_CARD) match_word = 0xFFFFFFFFu;
        if (site == SITE_DEAL) deal_word = 0xFFFFFFFFu;
        if (site == SITE_CARD) {
            if (o & 3u) throw std::runtime_error("set_ordinal(SITE_CARD) inside a trick: use set_card_position");
            card_trick = 0xFFFFFFFFu; card_mul = 1;
        }
    }
    void set_card_position(uint32_t card_index, uint32_t chain_mul) override {
        ordinal[SITE_CARD] = card_index; card_trick = 0xFFFFFFFFu; card_mul = (card_index & 3u) ? chain_mul : 1u;
    }
    uint32_t start_player() override { return below(SITE_DEAL, 4); }
    // Durstenfeld, descending: for i in 47..1: j = draw(i+1); swap(i, j).
    void shuffle48(uint8_t* cards) override {
        for (uint32_t i = 47; i >= 1; --i) {
            uint32_t j = below(SITE_DEAL, i + 1);
            uint8_t t = cards[i]; cards[i] = cards[j]; cards[j] = t;
        }
    }
    // One draw over `count`, ascending-bit indexing.
    uint32_t choose_unsized(Site site, uint32_t count) override { return below(site, count); }
};

// ---------------------------------------------------------------------------------------------
// rand 0.9.0 SmallRng emulation (64-bit targets): Xoshiro256++
// ---------------------------------------------------------------------------------------------
struct SmallRngStream final : Rng {
    uint64_t s[4];

    explicit SmallRngStream(uint64_t seed) { seed_from_u64(seed); }

    // Xoshiro256PlusPlus::seed_from_u64: four SplitMix64 outputs.
    void seed_from_u64(uint64_t state) {
        for (int i = 0; i < 4; ++i) {
            state += 0x9E3779B97F4A7C15ull;
            uint64_t z = state;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z = z ^ (z >> 31);
            s[i] = z;
        }
    }
    static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next_u64() {
        uint64_t result = rotl(s[0] + s[3], 23) + s[0];
        uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    // "The lowest bits have some linear dependencies, so we use the upper bits instead."
    uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }

    // UniformInt<u32>::sample_single_inclusive(0, n-1) — widening multiply with one bias-reducing
    // extra draw (the default, non-"unbiased" path of rand 0.9).  Also what `usize` ranges that fit
    // in u32 use (UniformUsize).
    uint32_t range_u32(uint32_t n) {
        if (n == 0) throw std::runtime_error("empty range");
        uint64_t m = (uint64_t)next_u32() * n;
        uint32_t result = (uint32_t)(m >> 32);
        uint32_t lo_order = (uint32_t)m;
        if (lo_order > (uint32_t)(0u - n)) {
            uint64_t m2 = (uint64_t)next_u32() * n;
            uint32_t new_hi = (uint32_t)(m2 >> 32);
            uint64_t sum = (uint64_t)lo_order + new_hi;
            if (sum > 0xFFFFFFFFull) result += 1;
        }
        return result;
    }

    uint32_t below(Site, uint32_t n) override { return range_u32(n); }
    uint32_t start_player() override { return range_u32(4); }

    // rand 0.9 `SliceRandom::shuffle` = partial_shuffle(len): forward loop swap(i, idx_i) with
    // idx_i uniform in 0..=i, indices produced by `IncreasingUniform` (one random_range over a
    // product of consecutive bounds, peeled with % and /).
    static void calculate_bound_u32(uint32_t m, uint32_t& bound, uint32_t& count) {
        uint32_t product = m;
        uint32_t current = m + 1;
        for (;;) {
            uint64_t p = (uint64_t)product * current;
            if (p <= 0xFFFFFFFFull) { product = (uint32_t)p; current += 1; }
            else { bound = product; count = current - m; return; }
        }
    }
    void shuffle48(uint8_t* cards) override {
        const uint32_t len = 48;
        uint32_t n = 0;               // IncreasingUniform::new(rng, 0)
        uint32_t chunk = 0;
        uint32_t chunk_remaining = 1; // n == 0 → first index is always 0, no draw
        for (uint32_t i = 0; i < len; ++i) {
            uint32_t next_n = n + 1;
            uint32_t next_chunk_remaining;
            if (chunk_remaining >= 1) {
                next_chunk_remaining = chunk_remaining - 1;
            } else {
                uint32_t bound, remaining;
                calculate_bound_u32(next_n, bound, remaining);
                chunk = range_u32(bound);
                next_chunk_remaining = remaining - 1;
            }
            uint32_t index;
            if (next_chunk_remaining == 0) {
                index = chunk;
            } else {
                index = chunk % next_n;
                chunk /= next_n;
            }
            chunk_remaining = next_chunk_remaining;
            n = next_n;
            uint8_t t = cards[i]; cards[i] = cards[index]; cards[index] = t;
        }
    }
    // The reservoir/CoinFlipper path of IteratorRandom::choose is not pinned by any reference
    // golden vector (card_matching's tests are OS-seeded), so it is not emulated.
    uint32_t choose_unsized(Site, uint32_t) override {
        throw std::runtime_error("SmallRng reservoir choose is not emulated (no reference vector pins it)");
    }
};

}  // namespace oracle
