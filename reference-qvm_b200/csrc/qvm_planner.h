// Host-side fusion planning (no device): which ops go into the next fused group.
//
// referenceqvm/_planner.py walks the op stream of a flush group by group; choosing ONE group is the inner loop -- a scan
// over the not-yet-run ops in program order that admits every op whose dependencies are met and whose dense targets
// still fit the tile -- and with the search below that scan runs ~50 times per group.  First fit lets whichever dense
// op comes first claim the tile's free slots; banning one or two of the qubits it admitted and scanning again often
// finds a tile that holds more of the stream (33-qubit depth-40 random circuits: 28 -> 25 passes, 30 qubits: 26 -> 23).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace qvm {

struct PlanStream {
    int n_ops;
    const uint8_t* dense;          // 1: the op has dense targets (tgt_off/tgt list them), 0: diagonal / control only
    const int32_t* tgt_off;        // n_ops + 1
    const int32_t* tgt;            // logical qubits (< 64)
    const int32_t* dep_off;        // n_ops + 1
    const int32_t* dep;            // indices of earlier ops that must have run (or be in the group)
    const uint8_t* done;
};

// One first-fit scan.  Returns the number of chosen ops; *tile_out = the group's qubits (seed included).
inline int plan_scan(const PlanStream& s, int first, int lookahead, int k, int max_new, uint64_t old_mask, uint64_t seed_mask,
                     int max_group_ops, uint64_t banned, std::vector<uint8_t>& in_group, int32_t* chosen, uint64_t* tile_out) {
    uint64_t tile = seed_mask;
    int new_count = __builtin_popcountll(tile & ~old_mask);
    int n_chosen = 0;
    const int stop = first + lookahead < s.n_ops ? first + lookahead : s.n_ops;
    for (int i = first; i < stop; ++i) {
        if (s.done[i]) continue;
        bool ready = true;
        for (int d = s.dep_off[i]; d < s.dep_off[i + 1]; ++d) {
            const int j = s.dep[d];
            if (!s.done[j] && !in_group[j]) {
                ready = false;
                break;
            }
        }
        if (!ready) continue;
        if (s.dense[i]) {
            uint64_t need = 0;
            for (int t = s.tgt_off[i]; t < s.tgt_off[i + 1]; ++t) need |= 1ull << s.tgt[t];
            need &= ~tile;
            if (need) {
                if (need & banned) continue;
                const int nn = __builtin_popcountll(need & ~old_mask);
                if (__builtin_popcountll(tile) + __builtin_popcountll(need) > k || new_count + nn > max_new) continue;
                tile |= need;
                new_count += nn;
            }
        }
        chosen[n_chosen++] = i;
        in_group[i] = 1;
        if (n_chosen >= max_group_ops) break;
    }
    for (int c = 0; c < n_chosen; ++c) in_group[chosen[c]] = 0;
    *tile_out = tile;
    return n_chosen;
}

// First fit, then up to `depth` levels of "ban one more of the qubits that candidate admitted and scan again"; the
// candidate holding the most ops wins (ties: the one found first, so depth 0 is plain first fit).
inline int plan_choose(const PlanStream& s, int first, int lookahead, int k, int max_new, uint64_t old_mask, uint64_t seed_mask,
                       int max_group_ops, int depth, int32_t* chosen, uint64_t* tile_out) {
    std::vector<uint8_t> in_group((size_t)s.n_ops, 0);
    std::vector<int32_t> scratch((size_t)max_group_ops);
    struct Cand { uint64_t banned, tile; };
    std::vector<Cand> frontier, next;
    std::vector<uint64_t> seen;
    uint64_t best_tile = 0;
    int best = plan_scan(s, first, lookahead, k, max_new, old_mask, seed_mask, max_group_ops, 0, in_group, chosen, &best_tile);
    frontier.push_back(Cand{0, best_tile});
    seen.push_back(0);
    for (int level = 0; level < depth; ++level) {
        next.clear();
        for (const Cand& c : frontier) {
            uint64_t free_qubits = c.tile & ~seed_mask & ~c.banned;
            while (free_qubits) {
                const uint64_t q = free_qubits & (~free_qubits + 1);
                free_qubits ^= q;
                const uint64_t banned = c.banned | q;
                bool dup = false;
                for (uint64_t b : seen) dup = dup || b == banned;
                if (dup) continue;
                seen.push_back(banned);
                uint64_t tile = 0;
                const int n = plan_scan(s, first, lookahead, k, max_new, old_mask, seed_mask, max_group_ops, banned, in_group,
                                        scratch.data(), &tile);
                next.push_back(Cand{banned, tile});
                if (n > best) {
                    best = n;
                    best_tile = tile;
                    memcpy(chosen, scratch.data(), (size_t)n * sizeof(int32_t));
                }
            }
        }
        frontier.swap(next);
    }
    *tile_out = best_tile;
    return best;
}

}  // namespace qvm
