// Checks include/doko_state_view.hpp on records handed over in a file: to_record(from_record(r)) must reproduce r byte for byte, and the
// view's fields are dumped (int32 rows) for tests/test_state_view.py to compare with the oracle's view of the same states.
//   state_view_check <records.bin> <views.bin>     → prints the number of records whose round trip differs
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "doko_state_view.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<dk_state> recs;
    dk_state r;
    while (std::fread(&r, sizeof r, 1, f) == 1) recs.push_back(r);
    std::fclose(f);
    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 2;
    size_t bad = 0;
    for (const dk_state& rec : recs) {
        const doko::FdoStateView v = doko::from_record(rec);
        const dk_state back = doko::to_record(v);
        if (std::memcmp(&back, &rec, sizeof rec) != 0) bad++;
        int32_t row[18 + 12 * 4];
        int k = 0;
        row[k++] = v.current_phase; row[k++] = v.current_player; row[k++] = v.game_type; row[k++] = v.card_index; row[k++] = v.n_tricks;
        row[k++] = v.team_tag; row[k++] = v.wedding_player; row[k++] = v.solved_trick_index; row[k++] = (int32_t)v.re_players;
        row[k++] = v.re_lowest_announcement; row[k++] = v.contra_lowest_announcement; row[k++] = v.number_of_turns_without_announcement;
        row[k++] = v.announcement_starting_player; row[k++] = v.n_announcements; row[k++] = v.current_player_allowed_call;
        row[k++] = v.reservation_result; row[k++] = v.reservation_result_player; row[k++] = v.reservation_result_reservation;
        for (int t = 0; t < 12; ++t) { row[k++] = v.tricks[t].starting_player; row[k++] = v.tricks[t].winning_player; row[k++] = v.tricks[t].winning_card; row[k++] = v.tricks[t].n_cards; }
        std::fwrite(row, sizeof row, 1, o);
    }
    std::fclose(o);
    std::printf("%zu %zu\n", recs.size(), bad);
    return bad ? 1 : 0;
}
