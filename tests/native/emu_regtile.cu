// CPU emulator of regtile_kernel -- TEST INFRASTRUCTURE, not part of libqvm_b200.so.
//
// Runs the product's own encoder (csrc/qvm_encode.h) and the product's own __host__ __device__
// round body (csrc/qvm_regtile.cuh: rt_resolve_op, rt_thread_round) on the host: for every tile
// ("CTA") the threads of a round are executed one after another and the loop boundary between
// rounds plays the role of __syncthreads().  This lets the planner + encoder + interpreter be
// diffed against the oracle in the CPU test suite (tests/test_regtile_emulator.py).
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../reference-qvm_b200/csrc/qvm_encode.h"

using namespace qvm;

// one tile: every round runs all threads one after another (the loop boundary is the barrier)
template <int R>
static void run_tile(const RegTileParams& P, RtCta c, const double2* gbase, unsigned threads, bool full_variant) {
    constexpr int NA = 1 << R;
    std::vector<double2> snapshot;
    std::vector<double2> pre((size_t)threads * NA);          // round 0's amplitudes (loaded at kernel entry)
    for (uint32_t tid = 0; tid < threads; ++tid) {
        double2 a[NA];
        rt_first_load<R>(P, gbase, tid, a);
        memcpy(&pre[(size_t)tid * NA], a, sizeof(a));
    }
    for (int r = 0; r < P.n_rounds; ++r) {
        // the relabelling round writes other slots than it reads; on the device a barrier separates the
        // two phases, here the round reads a snapshot of the tile
        double2* const real_tile = c.tile_w;
        if (r == P.perm_round && r > 0) {
            snapshot.assign(real_tile, real_tile + ((size_t)1 << P.k));
            c.tile = snapshot.data();
        } else c.tile = real_tile;
        for (uint32_t tid = 0; tid < threads; ++tid) {
            double2 a[NA];
            if (r == 0) memcpy(a, &pre[(size_t)tid * NA], sizeof(a));
            if (full_variant) rt_thread_round<true, R>(c, r, tid, a);
            else rt_thread_round<false, R>(c, r, tid, a);
        }
    }
}

extern "C" int qvm_emu_apply_group(double* amps, int n_local, int n_global, uint64_t shard_index, int n_tile,
                                   const int32_t* tile_bits, int n_ops, const qvm_op_t* ops, const double* mats,
                                   uint64_t mats_entries, int force_full, int reg_bits, int* info /* 38 ints */,
                                   unsigned low_in /* tile-local mask */, int n_low, int* sigma /* 16 ints or null */) {
    EncodedGroup g;
    const int erc = encode_group(n_local, n_global, shard_index, n_tile, tile_bits, n_ops, ops, mats, mats_entries, g,
                                 reg_bits, low_in, n_low);
    if (sigma)
        for (int b = 0; b < 16; ++b) sigma[b] = g.sigma[b];
    if (info) {
        info[0] = (int)g.rounds.size();
        info[1] = g.need_full ? 1 : 0;
        info[2] = erc;
        info[3] = (int)g.cold.size();
        for (int i = 0; i < 32; ++i) info[4 + i] = (int)g.opc_hist[i];
        info[36] = g.n_layers;
        info[37] = g.n_dirty_rounds;
    }
    if (erc != ENC_OK) return erc;
    if (g.cold.empty() && !g.relabelled) return 0;
    RegTileParams P = g.P;
    P.rounds = g.rounds.data();
    P.hot = g.hot.data();
    P.cold = g.cold.data();
    P.hi_off = g.hi_off.data();
    P.mats = reinterpret_cast<const double2*>(g.pool.data());
    std::vector<double2> tile((size_t)1 << P.k);
    std::vector<OpWord> recs(P.n_words + 1);
    double2* state = reinterpret_cast<double2*>(amps);
    const bool full_variant = g.need_full || force_full;
    for (uint64_t blk = 0; blk < g.n_tiles; ++blk) {
        const uint64_t base = rt_tile_base(P, blk);
        const uint64_t full = base | P.shard_bits;
        for (int j = 0; j < P.n_ops; ++j) rt_resolve_op(P, (uint32_t)j, full, recs.data());
        for (int j = 0; j < P.n_patches; ++j) rt_patch_op(P, (uint32_t)(P.n_ops + j), full, recs.data());
        recs[P.n_words].u = make_uint4(0, 0, 0, 0);
        RtCta c;
        c.tile = c.tile_w = tile.data();
        c.perm_round = P.perm_round;
        c.perm_swz = P.perm_swz;
        c.ops = recs.data();
        c.rounds = g.rounds.data();
        c.hi_off = g.hi_off.data();
        c.gbase = state + base;
        c.gstore = rt_push_base(P.push, state, base);
        c.mats = P.mats;
        c.cold = P.cold;
        c.low = P.low;
        c.n_rounds = P.n_rounds;
        if (P.R == 5) run_tile<5>(P, c, state + base, g.threads, full_variant);
        else run_tile<4>(P, c, state + base, g.threads, full_variant);
    }
    return 0;
}
