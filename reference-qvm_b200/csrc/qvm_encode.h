// Host-side round planner + encoder of the register-tiled kernel (see qvm_regtile.cuh).
//
// Input: a fused group in physical coordinates (qvm_op_t list + matrix pool + the tile's bits).
// Output: everything one regtile_kernel launch reads -- round table, op headers, cold words, the
// tile offset table and the (possibly extended) matrix pool -- as plain host vectors, so the same
// function feeds the CUDA launch (qvm_capi.cu) and the CPU emulator used by the tests
// (tests/native/emu_regtile.cu).
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/qvm_b200.h"
#include "qvm_regtile.cuh"

namespace qvm {

struct LocalOp {
    int kind = 0, k = 0;
    int tgt[QVM_MAX_TARGETS];        // dense: tile-local bit; diag: tile-local bit, or -(1 + physical) when outer
    uint32_t ctrl_tile = 0;
    uint64_t ctrl_outer = 0;
    uint32_t mat_off = 0;
    uint32_t dense_mask = 0;         // tile-local bits touched densely
    uint32_t diag_mask = 0;          // tile-local bits touched diagonally (controls, diagonal targets)
};

struct PlannedRound {
    int slot_bit[kMaxRegBits] = {-1, -1, -1, -1, -1};      // tile-local bit held by register slot s
    std::vector<int> ops;                    // indices into the LocalOp list, execution order
    uint32_t regmask() const {
        uint32_t m = 0;
        for (int s = 0; s < kMaxRegBits; ++s)
            if (slot_bit[s] >= 0) m |= 1u << slot_bit[s];
        return m;
    }
    int slot_of(int bit) const {
        for (int s = 0; s < kMaxRegBits; ++s)
            if (slot_bit[s] == bit) return s;
        return -1;
    }
};

struct EncodedGroup {
    RegTileParams P{};                       // device pointers left null
    std::vector<RoundDesc> rounds;
    std::vector<uint4> hot;                  // the op stream, 16-byte words
    std::vector<ColdOp> cold;                // one per op (plain op or LAYER)
    std::vector<uint64_t> hi_off;
    std::vector<double> pool;                // interleaved re/im
    bool need_full = false;
    unsigned threads = 0;
    size_t smem = 0;
    uint64_t n_tiles = 0;
    int n_gate_ops = 0;                      // ops of the group that were encoded (NOPs excluded)
    uint32_t opc_hist[32] = {};
    int n_layers = 0;                        // LAYER dispatches emitted (several single-slot ops each)
    int n_dirty_rounds = 0;                  // rounds that apply a thread scalar (16 multiplies per thread)
    // In-tile relabelling: the qubit that entered the pass on tile bit b leaves it on tile bit sigma[b]
    // (identity unless the caller asked for qubits to be brought to the low positions).
    uint8_t sigma[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};
    bool relabelled = false;
};

enum { ENC_OK = 0, ENC_FALLBACK = 1, ENC_ERROR = 2 };

// Try to give the dense targets of `op` register slots in `rd`; `allowed` = tile bits that may
// become register bits in this round.  Dense2 targets must share a slot pair {0,1} or {2,3}.
inline bool place_dense(PlannedRound& rd, const LocalOp& op, uint32_t allowed, int R) {
    if (op.k == 1) {
        const int b = op.tgt[0];
        if (rd.slot_of(b) >= 0) return true;
        if (!((allowed >> b) & 1u)) return false;
        for (int s = R - 1; s >= 0; --s)        // the unpaired top slot first: pairs stay free for 2q gates
            if (rd.slot_bit[s] < 0) {
                rd.slot_bit[s] = b;
                return true;
            }
        return false;
    }
    const int a = op.tgt[0], b = op.tgt[1];
    const int sa = rd.slot_of(a), sb = rd.slot_of(b);
    if (sa >= 0 && sb >= 0) return (sa >> 1) == (sb >> 1);
    if (sa >= 0 || sb >= 0) {
        const int have = sa >= 0 ? sa : sb;
        if (have >= 4) return false;         // the fifth slot has no partner
        const int other_bit = sa >= 0 ? b : a;
        const int partner = have ^ 1;
        if (rd.slot_bit[partner] >= 0 || !((allowed >> other_bit) & 1u)) return false;
        rd.slot_bit[partner] = other_bit;
        return true;
    }
    if (!((allowed >> a) & 1u) || !((allowed >> b) & 1u)) return false;
    for (int p = 0; p < 2; ++p)
        if (rd.slot_bit[2 * p] < 0 && rd.slot_bit[2 * p + 1] < 0) {
            rd.slot_bit[2 * p + 1] = a;      // msb -> odd slot
            rd.slot_bit[2 * p] = b;
            return true;
        }
    return false;
}

// QVM_B200_ROUND_SEARCH=0: plain first-fit round planning (A/B runs)
inline bool round_search_enabled() {
    static const bool on = [] {
        const char* e = getenv("QVM_B200_ROUND_SEARCH");
        return !e || atoi(e) != 0;
    }();
    return on;
}

inline bool is_x_matrix(const double* m) {      // compared as numbers: conj(X) carries -0.0 imaginary parts
    return m[0] == 0.0 && m[1] == 0.0 && m[2] == 1.0 && m[3] == 0.0 && m[4] == 1.0 && m[5] == 0.0 && m[6] == 0.0 &&
           m[7] == 0.0;
}
inline bool is_h_matrix(const double* m) {      // s * [[1, 1], [1, -1]] with real s
    const double s = m[0];
    return s != 0.0 && m[1] == 0.0 && m[2] == s && m[3] == 0.0 && m[4] == s && m[5] == 0.0 && m[6] == -s && m[7] == 0.0;
}

// ENC_OK: `out` is ready to launch.  ENC_FALLBACK: group is valid-looking but outside what the
// register-tiled kernel encodes (the shared-memory kernel takes it and reports real errors).
inline int encode_group(int n_local, int n_global, uint64_t shard_index, int n_tile, const int32_t* tile_bits,
                        int n_ops, const qvm_op_t* ops, const double* mats, uint64_t mats_entries, EncodedGroup& out,
                        int R = 4, uint32_t low_in = 0, int n_low = 0) {
    if (n_tile < 9 || n_tile > 13) return ENC_FALLBACK;
    if (R != 4 && R != 5) return ENC_FALLBACK;
    if (R == 5 && n_tile < 10) R = 4;          // a 2^9 tile would leave half a warp per CTA
    const int npos = n_local + n_global;
    int8_t local_of[64];
    memset(local_of, -1, sizeof(local_of));
    RegTileParams& P = out.P;
    P = RegTileParams{};
    P.n_local = n_local;
    P.k = n_tile;
    P.R = R;
    P.shard_bits = shard_index << n_local;
    uint64_t tile_mask = 0;
    for (int j = 0; j < n_tile; ++j) {
        tile_mask |= 1ull << tile_bits[j];
        local_of[tile_bits[j]] = (int8_t)j;
    }

    // tile-relative view (anything unusual falls back to the shared-memory kernel and its checks)
    std::vector<LocalOp> lops;
    lops.reserve(n_ops);
    for (int i = 0; i < n_ops; ++i) {
        const qvm_op_t& op = ops[i];
        if (op.kind == QVM_OP_NOP) continue;
        if (op.kind != QVM_OP_DENSE && op.kind != QVM_OP_DIAG) return ENC_FALLBACK;
        if (op.k < 0 || op.k > QVM_MAX_TARGETS) return ENC_FALLBACK;
        if (op.kind == QVM_OP_DENSE && (op.k < 1 || op.k > 2)) return ENC_FALLBACK;
        const uint64_t want = op.kind == QVM_OP_DENSE ? (1ull << (2 * op.k)) : (1ull << op.k);
        if (op.mat_entries != want || (uint64_t)op.mat_offset + want > mats_entries) return ENC_FALLBACK;
        if (npos < 64 && (op.ctrl_mask >> npos)) return ENC_FALLBACK;
        LocalOp l;
        l.kind = op.kind;
        l.k = op.k;
        l.mat_off = op.mat_offset;
        uint64_t used = op.ctrl_mask;
        for (int p = 0; p < n_local; ++p)
            if (((op.ctrl_mask >> p) & 1ull) && local_of[p] >= 0) l.ctrl_tile |= 1u << local_of[p];
        l.ctrl_outer = op.ctrl_mask & ~tile_mask;
        l.diag_mask = l.ctrl_tile;
        for (int j = 0; j < op.k; ++j) {
            const int p = op.targets[j];
            if (p < 0 || p >= npos || ((used >> p) & 1ull)) return ENC_FALLBACK;
            used |= 1ull << p;
            const bool in_tile = p < n_local && local_of[p] >= 0;
            if (op.kind == QVM_OP_DENSE) {
                if (!in_tile) return ENC_FALLBACK;
                l.tgt[j] = local_of[p];
                l.dense_mask |= 1u << local_of[p];
            } else if (in_tile) {
                l.tgt[j] = local_of[p];
                l.diag_mask |= 1u << local_of[p];
            } else {
                l.tgt[j] = -(1 + p);
            }
        }
        lops.push_back(l);
    }
    out.n_gate_ops = (int)lops.size();
    out.rounds.clear();
    out.hot.clear();
    out.cold.clear();
    out.hi_off.clear();
    out.pool.clear();
    out.need_full = false;
    out.n_layers = 0;
    out.relabelled = false;
    for (int b = 0; b < 16; ++b) out.sigma[b] = (uint8_t)b;
    P.perm_round = -1;
    memset(P.perm_swz, 0, sizeof(P.perm_swz));

    // ---- in-tile relabelling request --------------------------------------------------------------
    // `low_in`: tile bits whose qubits the caller wants on the n_low lowest tile bits when the pass ends
    // (the next group's coalescing bits).  Those already there stay; each of the others trades places with
    // a low bit that is not wanted (sigma = disjoint transpositions).  The pairing is fixed further down,
    // once the thread bits of the writing round are known (shared-memory bank conflicts depend on it).
    int n_in = 0, n_vic = 0;
    int in_bit[8], vic_bit[8];
    if (low_in && n_low > 0 && n_low <= 5 && n_low < n_tile) {
        for (int b = n_low; b < n_tile && n_in < 8; ++b)
            if ((low_in >> b) & 1u) in_bit[n_in++] = b;
        for (int b = n_low - 1; b >= 0 && n_vic < n_in; --b)
            if (!((low_in >> b) & 1u)) vic_bit[n_vic++] = b;
        n_in = std::min(n_in, n_vic);            // more requests than low bits: the first ones win
    }
    const bool relabel = n_in > 0;
    // tile bits (old coordinates) whose qubits end on the five lowest tile bits: the last round keeps those
    // for its lanes, as round 0 keeps the five lowest old bits
    uint32_t low5_last = 0x1Fu;
    for (int i = 0; i < n_in; ++i) low5_last = (low5_last & ~(1u << vic_bit[i])) | (1u << in_bit[i]);
    if (lops.empty() && !relabel) return ENC_OK;

    // ---- rounds: greedy, commutation aware ------------------------------------------------------
    // Dense ops choose the rounds (each needs its targets in register slots).  Diagonal ops float:
    // they commute with everything except dense ops on their own qubits, so each is placed, inside
    // the window those neighbours leave, in the round where it touches the fewest register bits
    // (ideally none: then it is one scalar per thread instead of work on the register slots).
    const uint32_t all_bits = (1u << n_tile) - 1u;
    const uint32_t high_bits = all_bits & ~0x1Fu;          // lanes keep the five lowest tile bits in I/O rounds
    const int n = (int)lops.size();
    std::vector<int> dense_idx;
    for (int i = 0; i < n; ++i)
        if (lops[i].kind == QVM_OP_DENSE) dense_idx.push_back(i);
    std::vector<int> round_of(n, -1);          // earliest legal round of every op (all-ops greedy)
    std::vector<PlannedRound> rounds;
    {
        // One round, first fit: ops in program order; a dense op joins when its targets get register slots (bits in
        // `banned` may not become register bits), anything that cannot join blocks what depends on it.
        auto build_round = [&](const std::vector<char>& done, bool first_round, uint32_t banned, PlannedRound& rd,
                               std::vector<int>& placed) {
            const uint32_t allowed = (first_round ? high_bits : all_bits) & ~banned;
            uint32_t blocked_dense = 0, blocked_diag = 0;
            placed.clear();
            for (int i = 0; i < n; ++i) {
                if (done[i]) continue;
                const LocalOp& op = lops[i];
                const bool clash = (op.dense_mask & (blocked_dense | blocked_diag)) || (op.diag_mask & blocked_dense);
                bool ok = !clash;
                if (ok && op.kind == QVM_OP_DENSE) {
                    PlannedRound trial = rd;
                    ok = place_dense(trial, op, allowed, R);
                    if (ok) memcpy(rd.slot_bit, trial.slot_bit, sizeof(rd.slot_bit));
                }
                if (!ok) {
                    blocked_dense |= op.dense_mask;
                    blocked_diag |= op.diag_mask;
                    continue;
                }
                placed.push_back(i);
            }
        };
        // rounds the rest of the group takes with plain first fit from `done` on (store round included); -1: stuck
        auto rounds_to_finish = [&](std::vector<char> done, int remaining, int rounds_so_far, uint32_t last_mask) {
            std::vector<int> placed;
            int total = rounds_so_far;
            while (remaining > 0) {
                PlannedRound rd;
                build_round(done, total == 0, 0, rd, placed);
                if (placed.empty() && total > 0) return -1;
                for (int i : placed) done[i] = 1;
                remaining -= (int)placed.size();
                last_mask = rd.regmask();
                ++total;
            }
            if (total == 0) total = 1;
            if ((total > 1 && (last_mask & low5_last)) || (relabel && total == 1)) ++total;
            return total;
        };
        std::vector<char> done(n, 0);
        std::vector<int> placed, best_placed;
        int remaining = n;
        while (remaining > 0) {
            const bool first_round = rounds.empty();
            PlannedRound rd;
            build_round(done, first_round, 0, rd, placed);
            if (round_search_enabled() && remaining > (int)placed.size()) {
                // Which tile bits become register bits decides how many rounds the group needs.  First fit hands the
                // slots to whichever dense op comes first; try keeping one or two of those bits out of this round
                // and keep the variant with which the group finishes in the fewest rounds (then: most ops now).
                std::vector<char> trial_done = done;
                for (int i : placed) trial_done[i] = 1;
                int best_total = rounds_to_finish(trial_done, remaining - (int)placed.size(), (int)rounds.size() + 1, rd.regmask());
                if (best_total < 0) best_total = 1 << 20;
                size_t best_now = placed.size();
                PlannedRound best_rd = rd;
                best_placed = placed;
                const uint32_t base_mask = rd.regmask();
                std::vector<uint32_t> bans;
                for (int a = 0; a < n_tile; ++a)
                    if ((base_mask >> a) & 1u) {
                        bans.push_back(1u << a);
                        for (int b = a + 1; b < n_tile; ++b)
                            if ((base_mask >> b) & 1u) bans.push_back((1u << a) | (1u << b));
                    }
                for (uint32_t ban : bans) {
                    PlannedRound cand;
                    build_round(done, first_round, ban, cand, placed);
                    if (placed.empty() && !first_round) continue;
                    trial_done = done;
                    for (int i : placed) trial_done[i] = 1;
                    const int total = rounds_to_finish(trial_done, remaining - (int)placed.size(), (int)rounds.size() + 1, cand.regmask());
                    if (total < 0) continue;
                    if (total < best_total || (total == best_total && placed.size() > best_now)) {
                        best_total = total;
                        best_now = placed.size();
                        best_rd = cand;
                        best_placed = placed;
                    }
                }
                rd = best_rd;
                placed = best_placed;
            }
            for (int i : placed) {
                round_of[i] = (int)rounds.size();
                done[i] = 1;
            }
            remaining -= (int)placed.size();
            if (placed.empty() && !rounds.empty()) return ENC_ERROR;
            rounds.push_back(rd);
        }
    }
    if (rounds.empty()) rounds.push_back(PlannedRound());            // pure relabelling pass
    if ((rounds.size() > 1 && (rounds.back().regmask() & low5_last)) || (relabel && rounds.size() == 1))
        rounds.push_back(PlannedRound());   // store round
    if (rounds.size() > 4096) return ENC_FALLBACK;
    // Fill unused register slots.  A filler bit turns every control / diagonal factor on it into
    // register-slot work (pair swaps instead of a relabel, 8 multiplies instead of one per thread), so
    // pick the free bits that the ops of this round touch least; ties go to the highest bit.  I/O rounds
    // keep the five lowest tile bits for the lanes.
    for (size_t r = 0; r < rounds.size(); ++r) {
        PlannedRound& rd = rounds[r];
        int touch[16] = {};
        for (int i = 0; i < n; ++i)
            if (round_of[i] == (int)r)
                for (int b = 0; b < n_tile; ++b)
                    if ((lops[i].diag_mask >> b) & 1u) ++touch[b];
        const bool io_round = r == 0 || r + 1 == rounds.size();
        const uint32_t lanes = (r + 1 == rounds.size() && r > 0) ? low5_last : 0x1Fu;
        uint32_t used = rd.regmask();
        for (int s = 0; s < R; ++s) {
            if (rd.slot_bit[s] >= 0) continue;
            int best = -1;
            for (int pass = 0; pass < 2 && best < 0; ++pass)
                for (int b = n_tile - 1; b >= 0; --b) {
                    if ((used >> b) & 1u) continue;
                    if (pass == 0 && io_round && ((lanes >> b) & 1u)) continue;
                    if (best < 0 || touch[b] < touch[best]) best = b;
                }
            rd.slot_bit[s] = best;
            used |= 1u << best;
        }
    }

    // ---- fix the relabelling permutation ----------------------------------------------------------
    // Round 0 -- which reads HBM, not the tile, so no thread waits for another's slots -- writes the tile
    // in the new coordinates and every later round runs in them.  Its quarter-warps (lanes on its three
    // lowest thread bits) store 8 x 16 bytes at once: conflict-free when the swizzle patterns of the three
    // images under sigma are linearly independent.  Try every pairing of incoming and outgoing bits and
    // keep the best.
    uint8_t* sigma = out.sigma;
    if (relabel) {
        const PlannedRound& wr = rounds[0];
        int lane_bit[3], nl = 0;
        for (int b = 0; b < n_tile && nl < 3; ++b)
            if (wr.slot_of(b) < 0) lane_bit[nl++] = b;
        int perm[8], best_perm[8], best_rank = -1;
        for (int i = 0; i < n_in; ++i) perm[i] = best_perm[i] = i;
        do {
            uint8_t sg[16];
            for (int b = 0; b < 16; ++b) sg[b] = (uint8_t)b;
            for (int i = 0; i < n_in; ++i) {
                sg[in_bit[i]] = (uint8_t)vic_bit[perm[i]];
                sg[vic_bit[perm[i]]] = (uint8_t)in_bit[i];
            }
            uint32_t span[8] = {1, 0, 0, 0, 0, 0, 0, 0};      // span[v] = 1: v is a XOR of the patterns so far
            int rank = 0;
            for (int j = 0; j < nl; ++j) {
                const uint32_t pat = swz(1u << sg[lane_bit[j]]) & 7u;
                if (span[pat]) continue;
                ++rank;
                for (uint32_t v = 0; v < 8; ++v)
                    if (span[v] == 1) span[v ^ pat] = 2;
                for (uint32_t v = 0; v < 8; ++v)
                    if (span[v]) span[v] = 1;
            }
            if (rank > best_rank) {
                best_rank = rank;
                memcpy(best_perm, perm, sizeof(perm));
            }
        } while (best_rank < nl && std::next_permutation(perm, perm + n_in));
        for (int i = 0; i < n_in; ++i) {
            sigma[in_bit[i]] = (uint8_t)vic_bit[best_perm[i]];
            sigma[vic_bit[best_perm[i]]] = (uint8_t)in_bit[i];
        }
        out.relabelled = true;
        P.perm_round = 0;
        for (int b = 0; b < n_tile; ++b) P.perm_swz[b] = (uint16_t)swz(1u << sigma[b]);
    }

    // place the floating diagonal ops
    {
        const int n_rounds = (int)rounds.size();
        std::vector<uint32_t> regmask(n_rounds);
        std::vector<int> ts_ops(n_rounds, 0);     // thread-scalar ops already placed in each round
        for (int r = 0; r < n_rounds; ++r) regmask[r] = rounds[r].regmask();
        for (int i = 0; i < n; ++i) {
            if (lops[i].kind == QVM_OP_DENSE) continue;
            const uint32_t touched = lops[i].diag_mask;
            const int lo = round_of[i];                 // earliest round that respects every dependency
            int hi = n_rounds - 1;
            for (int j : dense_idx)
                if (j > i && (lops[j].dense_mask & touched)) hi = std::min(hi, round_of[j]);
            if (hi < lo) hi = lo;      // cannot happen: a later conflicting dense op waits for this op
            // fewest register bits touched first; among those, a round that already applies a thread
            // scalar (each such round costs 16 complex multiplies per thread once, however many ops
            // share it); then the latest round
            int best = lo, best_cost = 99, best_dirty = -1;
            for (int r = hi; r >= lo; --r) {
                const int cost = __builtin_popcount(touched & regmask[r]);
                const int dirty = cost == 0 ? ts_ops[r] : 0;
                if (cost < best_cost || (cost == best_cost && dirty > best_dirty)) {
                    best_cost = cost;
                    best_dirty = dirty;
                    best = r;
                }
            }
            if (best_cost == 0) ++ts_ops[best];
            round_of[i] = best;
        }
        for (int i = 0; i < n; ++i) rounds[round_of[i]].ops.push_back(i);     // ascending original index
    }

    // ---- encode ---------------------------------------------------------------------------------
    std::vector<double>& pool = out.pool;
    pool.assign(mats, mats + 2 * mats_entries);
    out.rounds.resize(rounds.size());
    std::vector<uint4>& hot = out.hot;
    std::vector<ColdOp>& cold = out.cold;
    std::vector<ColdOp> patches;             // appended to `cold` once all plain ops / layers are known
    double gs_re = 1.0, gs_im = 0.0;     // global complex scalar: folded H scales and RZ d0 phases
    const std::vector<LocalOp>& lops_old = lops;
    hot.reserve(n);
    cold.reserve(n);
    int dirty_round = -1, dirty_best = -1;
    std::vector<LocalOp> lops_new;       // the ops in the tile coordinates of rounds >= 1 (after sigma)
    if (relabel) {
        auto map_mask = [&](uint32_t m) {
            uint32_t o = 0;
            for (int b = 0; b < n_tile; ++b)
                if ((m >> b) & 1u) o |= 1u << sigma[b];
            return o;
        };
        lops_new = lops;
        for (LocalOp& l : lops_new) {
            for (int t = 0; t < l.k; ++t)
                if (l.tgt[t] >= 0) l.tgt[t] = sigma[l.tgt[t]];
            l.ctrl_tile = map_mask(l.ctrl_tile);
            l.dense_mask = map_mask(l.dense_mask);
            l.diag_mask = map_mask(l.diag_mask);
        }
    }
    for (size_t r = 0; r < rounds.size(); ++r) {
        const bool new_coords = relabel && r >= 1;
        if (new_coords)
            for (int s = 0; s < R; ++s) rounds[r].slot_bit[s] = sigma[rounds[r].slot_bit[s]];
        const std::vector<LocalOp>& lops = new_coords ? lops_new : lops_old;
        const PlannedRound& rd = rounds[r];
        RoundDesc& d = out.rounds[r];
        memset(&d, 0, sizeof(d));
        if (relabel && (int)r == P.perm_round)
            for (int s = 0; s < R; ++s) d.wr_swz[s] = (uint16_t)swz(1u << sigma[rd.slot_bit[s]]);
        d.scale = make_double2(1.0, 0.0);
        int8_t slot_of[16], thr_of[16];
        memset(slot_of, -1, sizeof(slot_of));
        memset(thr_of, -1, sizeof(thr_of));
        int sorted[kMaxRegBits];
        for (int s = 0; s < R; ++s) {
            d.reg_phys[s] = (uint8_t)tile_bits[rd.slot_bit[s]];
            d.swz_j[s] = (uint16_t)swz(1u << rd.slot_bit[s]);
            slot_of[rd.slot_bit[s]] = (int8_t)s;
            sorted[s] = rd.slot_bit[s];
        }
        for (int s = 1; s < R; ++s) {       // R <= kMaxRegBits: insertion sort
            const int v = sorted[s];
            int t = s;
            for (; t > 0 && sorted[t - 1] > v; --t) sorted[t] = sorted[t - 1];
            sorted[t] = v;
        }
        for (int s = 0; s < R; ++s) d.ins_mask[s] = (uint16_t)~((1u << sorted[s]) - 1u);
        int nt = 0;
        for (int b = 0; b < n_tile; ++b)
            if (slot_of[b] < 0) thr_of[b] = (int8_t)nt++;
        d.op_begin = (uint16_t)hot.size();

        // thread-scalar ops (no register-slot involvement in this round, at most two thread-bit
        // targets) commute with the whole round: emit them first, the kernel runs them as one batch
        auto thread_scalar = [&](const LocalOp& l) {
            if (l.kind != QVM_OP_DIAG) return false;
            int nthr = 0;
            for (int b = 0; b < n_tile; ++b)
                if (((l.ctrl_tile >> b) & 1u) && slot_of[b] >= 0) return false;
            for (int t = 0; t < l.k; ++t) {
                if (l.tgt[t] < 0) continue;
                if (slot_of[l.tgt[t]] >= 0) return false;
                ++nthr;
            }
            return nthr <= 2;
        };
        std::vector<int> order;
        order.reserve(rd.ops.size());
        for (int i : rd.ops)
            if (thread_scalar(lops[i])) order.push_back(i);
        const size_t n_ts = order.size();
        if ((int)n_ts > dirty_best) {
            dirty_best = (int)n_ts;
            dirty_round = (int)r;
        }
        for (int i : rd.ops)
            if (!thread_scalar(lops[i])) order.push_back(i);

        struct EncOp {
            uint4 h;
            ColdOp c;
            int lay_type;          // 0: not layerable; 1: S1, 2: H, 3: X relabel (position inside a layer slot)
            int lay_slot;
            double f_re, f_im;     // S1 factor (known on the host when layerable)
        };
        std::vector<EncOp> enc;
        enc.reserve(order.size());
        for (size_t oi_ = 0; oi_ < order.size(); ++oi_) {
            const LocalOp& l = lops[order[oi_]];
            const bool in_batch = oi_ < n_ts;
            ColdOp c;
            memset(&c, 0, sizeof(c));
            c.ctrl_outer = l.ctrl_outer;
            c.fsel = 0xFF;
            c.mat_off = l.mat_off;
            uint32_t mat_off = l.mat_off;
            uint32_t opc = 0, crm = 0, cthr = 0, xb = 0, s0 = 0, s1 = 0, p0 = 31, p1 = 31, shr = 0, flags = 0;
            for (int b = 0; b < n_tile; ++b)
                if ((l.ctrl_tile >> b) & 1u) {
                    if (slot_of[b] >= 0) crm |= 1u << slot_of[b];
                    else cthr |= 1u << thr_of[b];
                }
            if (l.kind == QVM_OP_DENSE && l.k == 1) {
                const uint32_t slot = (uint32_t)slot_of[l.tgt[0]];
                const double* m = &pool[2 * (size_t)l.mat_off];
                if (is_x_matrix(m)) {
                    if (crm == 0) {
                        opc = OPC_X_RELABEL;
                        xb = slot;
                    } else opc = OPC_XS_0 + slot;
                } else if (!l.ctrl_tile && !l.ctrl_outer && is_h_matrix(m)) {
                    opc = OPC_H0 + slot;
                    gs_re *= m[0];
                    gs_im *= m[0];
                } else {
                    opc = OPC_M1_0 + slot;
                    out.need_full = true;
                }
            } else if (l.kind == QVM_OP_DENSE) {
                const int sa = slot_of[l.tgt[0]];
                opc = OPC_M2_0 + (uint32_t)(sa >> 1);
                out.need_full = true;
                if (!(sa & 1)) {
                    // msb target sits in the even slot: swap the roles of the two index bits
                    static const int perm[4] = {0, 2, 1, 3};
                    const size_t off = pool.size() / 2;
                    pool.resize(pool.size() + 32);
                    for (int rr = 0; rr < 4; ++rr)
                        for (int cc = 0; cc < 4; ++cc) {
                            const size_t src = 2 * ((size_t)l.mat_off + perm[rr] * 4 + perm[cc]);
                            pool[2 * (off + rr * 4 + cc)] = pool[src];
                            pool[2 * (off + rr * 4 + cc) + 1] = pool[src + 1];
                        }
                    mat_off = (uint32_t)off;
                }
            } else {
                c.k = (uint8_t)l.k;
                int nreg = 0, nthr = 0;
                uint32_t reg_slot = 0;
                for (int t = 0; t < l.k; ++t) {
                    const int b = l.tgt[t];
                    const uint32_t wshift = (uint32_t)(l.k - 1 - t);
                    if (b < 0) c.loc[t] = (uint8_t)(LOC_OUT | (-(b + 1)));
                    else if (slot_of[b] >= 0) {
                        c.loc[t] = (uint8_t)(LOC_REG | slot_of[b]);
                        reg_slot = (uint32_t)slot_of[b];
                        shr = wshift;
                        ++nreg;
                    } else {
                        c.loc[t] = (uint8_t)(LOC_THR | thr_of[b]);
                        if (nthr == 0) { p0 = (uint32_t)thr_of[b]; s0 = wshift; }
                        else if (nthr == 1) { p1 = (uint32_t)thr_of[b]; s1 = wshift; }
                        ++nthr;
                    }
                }
                const int ncrm = __builtin_popcount(crm);
                if (in_batch) {
                    if (nthr <= 1) {
                        opc = OPC_T;
                        c.prefetch = 1;
                        if (nthr == 1) c.fsel = (uint8_t)s0;
                    } else opc = OPC_TG;
                } else if (nreg == 0 && nthr <= 1) {
                    // factor (by at most one thread bit) on the slots selected by register-slot controls
                    opc = ncrm == 1 ? OPC_S1_0 + (uint32_t)__builtin_ctz(crm) : OPC_DS;
                    c.prefetch = 1;
                    if (nthr == 1) {
                        opc = OPC_DST;
                        out.need_full = true;
                        c.fsel = (uint8_t)s0;
                        flags |= HDR_HAS_F1;
                    }
                } else if (nreg == 1 && nthr == 0) {
                    const double* dd = &pool[2 * (size_t)l.mat_off];
                    const bool plain_rz = l.k == 1 && !l.ctrl_tile && !l.ctrl_outer && (dd[0] != 0.0 || dd[1] != 0.0);
                    if (plain_rz) {
                        // diag(d0, d1) = d0 * diag(1, d1/d0): d0 joins the group's scalar, the rest touches
                        // only the 8 slots whose bit is 1
                        const double den = dd[0] * dd[0] + dd[1] * dd[1];
                        const double qr = (dd[2] * dd[0] + dd[3] * dd[1]) / den, qi = (dd[3] * dd[0] - dd[2] * dd[1]) / den;
                        const double nr = gs_re * dd[0] - gs_im * dd[1], ni = gs_re * dd[1] + gs_im * dd[0];
                        gs_re = nr;
                        gs_im = ni;
                        const size_t off = pool.size() / 2;
                        pool.push_back(qr);
                        pool.push_back(qi);
                        mat_off = (uint32_t)off;
                        c.mat_off = mat_off;
                        c.k = 0;                       // the pooled entry is a bare scalar now
                        memset(c.loc, 0, sizeof(c.loc));
                        opc = OPC_S1_0 + reg_slot;
                        c.prefetch = 1;
                    } else {
                        opc = OPC_D1_0 + reg_slot;
                        out.need_full = true;
                        c.prefetch = 1;
                        c.fsel = (uint8_t)shr;
                        flags |= HDR_HAS_F1;
                    }
                } else {
                    opc = OPC_DIAG_G;
                    out.need_full = true;
                }
            }
            EncOp e;
            e.h.x = opc | (crm << 8) | (xb << 13) | (p0 << 16) | (p1 << 24);
            e.h.y = cthr | flags | ((uint32_t)kOpWords << HDR_SIZE_SHIFT);
            e.h.z = mat_off;
            e.h.w = (s0 << 8) | (s1 << 16) | (shr << 24);
            e.c = c;
            e.lay_type = 0;
            e.lay_slot = 0;
            e.f_re = 1.0;
            e.f_im = 0.0;
            // single-slot ops share LAYER dispatches (outer dependence is patched in by the prologue)
            if (!in_batch) {
                if (opc >= OPC_H0 && opc < OPC_H0 + kMaxRegBits) {
                    e.lay_type = 2;
                    e.lay_slot = (int)(opc - OPC_H0);
                } else if (opc == OPC_X_RELABEL) {
                    e.lay_type = 3;
                    e.lay_slot = (int)xb;
                } else if (opc >= OPC_S1_0 && opc < OPC_S1_0 + kMaxRegBits) {
                    e.lay_type = 1;
                    e.lay_slot = (int)(opc - OPC_S1_0);
                    e.f_re = pool[2 * (size_t)mat_off];          // right when no outer target selects it
                    e.f_im = pool[2 * (size_t)mat_off + 1];
                }
            }
            out.opc_hist[opc & 31] += 1;
            enc.push_back(e);
        }

        // ---- emit: thread-scalar batch, then slot ops packed into layers -------------------------
        auto emit_plain = [&](EncOp& e) {
            e.c.n_words = (uint8_t)kOpWords;
            e.c.kind = COLD_PLAIN;
            e.c.word_off = (uint32_t)hot.size();
            if ((e.h.x & 0xFFu) == OPC_DIAG_G) e.h.w = (uint32_t)cold.size() << 8;      // its cold index
            hot.push_back(e.h);
            double f[2][2] = {{1.0, 0.0}, {1.0, 0.0}};
            if (e.c.prefetch) {
                bool outer_dep = false;
                for (int t = 0; t < e.c.k; ++t)
                    if ((e.c.loc[t] & LOC_CLASS) == LOC_OUT) outer_dep = true;
                if (!outer_dep) {           // factors known here: the prologue just copies them
                    const size_t m0 = (size_t)e.c.mat_off;
                    f[0][0] = pool[2 * m0];
                    f[0][1] = pool[2 * m0 + 1];
                    if (e.c.fsel != 0xFF) {
                        const size_t m1 = m0 + ((size_t)1 << e.c.fsel);
                        f[1][0] = pool[2 * m1];
                        f[1][1] = pool[2 * m1 + 1];
                    }
                    e.c.prefetch = 0;
                }
            }
            for (int i = 0; i < 2; ++i) {
                uint4 w;
                memcpy(&w.x, &f[i][0], 8);
                memcpy(&w.z, &f[i][1], 8);
                hot.push_back(w);
            }
            cold.push_back(e.c);
        };
        {
            unsigned n_tg = 0;
            for (size_t i = 0; i < n_ts; ++i)
                if ((enc[i].h.x & 0xFFu) == OPC_T) emit_plain(enc[i]);
            for (size_t i = 0; i < n_ts; ++i)
                if ((enc[i].h.x & 0xFFu) != OPC_T) {
                    emit_plain(enc[i]);
                    ++n_tg;
                }
            d.flags |= (uint16_t)(n_tg << 1);
        }
        d.dt_end = (uint16_t)hot.size();

        struct Layer {
            uint32_t s1 = 0, h = 0, x = 0;
            uint32_t cs[kMaxRegBits] = {0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu};      // 0xFFFF: absent
            uint32_t cx[kMaxRegBits] = {0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu};
            double f[kMaxRegBits][2] = {{1, 0}, {1, 0}, {1, 0}, {1, 0}, {1, 0}};
            int stage[kMaxRegBits] = {0, 0, 0, 0, 0};
            int count = 0;
            std::vector<ColdOp> pending;     // patches of sub-ops that depend on outer bits (word_off set at emit)
        };
        auto emit_layer = [&](Layer& L) {
            if (!L.count) return;
            ColdOp c;
            memset(&c, 0, sizeof(c));
            c.fsel = 0xFF;
            c.n_words = (uint8_t)layer_words(R);
            c.kind = COLD_LAYER;
            c.word_off = (uint32_t)hot.size();
            for (ColdOp& pc : L.pending) {
                pc.word_off = c.word_off;
                patches.push_back(pc);
            }
            uint4 h;
            h.x = OPC_LAYER | (L.h << 12);
            h.y = (uint32_t)layer_words(R) << HDR_SIZE_SHIFT;
            h.z = h.w = 0;
            hot.push_back(h);
            hot.push_back(make_uint4(L.cs[0] | (L.cs[1] << 16), L.cs[2] | (L.cs[3] << 16), L.cx[0] | (L.cx[1] << 16),
                                     L.cx[2] | (L.cx[3] << 16)));
            if (R == 5) hot.push_back(make_uint4(L.cs[4] | (L.cx[4] << 16), 0, 0, 0));
            for (int sl = 0; sl < R; ++sl) {
                uint4 w;
                memcpy(&w.x, &L.f[sl][0], 8);
                memcpy(&w.z, &L.f[sl][1], 8);
                hot.push_back(w);
            }
            cold.push_back(c);
            ++out.n_layers;
        };
        {
            // Sub-ops on different slots commute (their controls are thread bits, which no op of the round
            // targets), so each slot's sequence is packed independently into the run of layers.  An op that
            // is not layerable orders itself only against the slots it touches: it is emitted right after
            // the last layer those slots have used, and their later sub-ops start in the layer after it.
            std::vector<Layer> layers;
            std::vector<std::vector<size_t>> after(1);      // after[i]: plain ops emitted before layer i
            int last[kMaxRegBits] = {-1, -1, -1, -1, -1}, min_layer[kMaxRegBits] = {0, 0, 0, 0, 0};
            auto slots_of = [&](const EncOp& e) -> uint32_t {
                const uint32_t opc = e.h.x & 0xFFu, crm = (e.h.x >> 8) & 0x1Fu;
                if (opc >= OPC_XS_0 && opc < OPC_XS_0 + kMaxRegBits) return crm | (1u << (opc - OPC_XS_0));
                if (opc == OPC_DS || opc == OPC_DST) return crm;
                if (opc >= OPC_D1_0 && opc < OPC_D1_0 + kMaxRegBits) return crm | (1u << (opc - OPC_D1_0));
                return 0x1Fu;
            };
            for (size_t i = n_ts; i < enc.size(); ++i) {
                EncOp& e = enc[i];
                if (!e.lay_type) {
                    const uint32_t m = slots_of(e);
                    int pos = -1;
                    for (int sl = 0; sl < R; ++sl)
                        if ((m >> sl) & 1u) pos = std::max(pos, std::max(last[sl], min_layer[sl] - 1));
                    if ((int)after.size() < pos + 2) after.resize(pos + 2);
                    after[pos + 1].push_back(i);
                    for (int sl = 0; sl < R; ++sl)
                        if ((m >> sl) & 1u) {
                            min_layer[sl] = pos + 1;
                            last[sl] = -1;
                        }
                    continue;
                }
                const int sl = e.lay_slot;
                int li = std::max(last[sl], min_layer[sl]);
                if (li == last[sl] && layers[li].stage[sl] >= e.lay_type) ++li;
                if (li >= (int)layers.size()) layers.resize(li + 1);
                Layer& L = layers[li];
                L.stage[sl] = e.lay_type;
                ++L.count;
                const uint32_t cthr = e.h.y & 0xFFFFu;
                if (e.lay_type == 1) {
                    L.s1 |= 1u << sl;
                    L.cs[sl] = cthr;
                    L.f[sl][0] = e.f_re;
                    L.f[sl][1] = e.f_im;
                } else if (e.lay_type == 2) {
                    L.h |= 1u << sl;
                } else {
                    L.x |= 1u << sl;
                    L.cx[sl] = cthr;
                }
                if (e.lay_type != 2 && (e.c.ctrl_outer || (e.lay_type == 1 && e.c.k))) {
                    ColdOp pc = e.c;
                    pc.kind = e.lay_type == 1 ? COLD_PATCH_S1 : COLD_PATCH_X;
                    pc.slot = (uint8_t)sl;
                    pc.n_words = 0;
                    L.pending.push_back(pc);
                }
                last[sl] = li;
            }
            const size_t n_slots = std::max(layers.size() + 1, after.size());
            after.resize(n_slots);
            for (size_t li = 0; li < n_slots; ++li) {
                for (size_t i : after[li]) emit_plain(enc[i]);
                if (li < layers.size()) emit_layer(layers[li]);
            }
        }
        d.op_end = (uint16_t)hot.size();
    }
    if (gs_re != 1.0 || gs_im != 0.0) {
        RoundDesc& d = out.rounds[dirty_round < 0 ? (int)rounds.size() - 1 : dirty_round];
        d.scale = make_double2(gs_re, gs_im);
        d.flags |= 1u;
    }
    out.n_dirty_rounds = 0;
    for (const RoundDesc& rd : out.rounds)
        if (rd.dt_end > rd.op_begin || (rd.flags & 1u)) ++out.n_dirty_rounds;
    if (hot.size() > 0xFFFFu) return ENC_FALLBACK;

    // launch parameters that replace per-thread bit loops
    int low = 0;
    while (low < n_tile && tile_bits[low] == low) ++low;
    if (low > 5) low = 5;                       // keeps the offset table at >= 2^(k-5) entries, <= 2^13
    P.low = low;
    {
        int nr = 0;
        int p = 0;
        while (p < n_local) {
            if ((tile_mask >> p) & 1ull) { ++p; continue; }
            int e = p;
            while (e < n_local && !((tile_mask >> e) & 1ull)) ++e;
            if (nr >= 8) return ENC_FALLBACK;   // pathological tile shape: let the shared-memory kernel take it
            P.run_start |= (uint64_t)p << (8 * nr);
            P.run_len |= (uint64_t)(e - p) << (8 * nr);
            ++nr;
            p = e;
        }
        P.n_runs = nr;
    }
    out.hi_off.resize((size_t)1 << (n_tile - low));
    for (size_t j = 0; j < out.hi_off.size(); ++j) {
        uint64_t off = 0;
        for (int b = 0; b < n_tile - low; ++b)
            if ((j >> b) & 1u) off |= 1ull << tile_bits[low + b];
        out.hi_off[j] = off;
    }
    P.n_rounds = (int)out.rounds.size();
    {
        const RoundDesc& r0 = out.rounds[0];
        for (int sl = 0; sl < R; ++sl) {
            P.first_reg_phys[sl] = r0.reg_phys[sl];
            P.first_ins[sl] = r0.ins_mask[sl];
        }
    }
    P.n_ops = (int)cold.size();
    P.n_patches = (int)patches.size();
    cold.insert(cold.end(), patches.begin(), patches.end());
    P.n_words = (int)hot.size();
    out.smem = regtile_smem_bytes(P.k, P.low, P.n_rounds, P.n_words);
    if (out.smem > 227 * 1024) return ENC_FALLBACK;
    out.n_tiles = (1ull << n_local) >> n_tile;
    out.threads = 1u << (n_tile - R);
    return ENC_OK;
}

}  // namespace qvm
