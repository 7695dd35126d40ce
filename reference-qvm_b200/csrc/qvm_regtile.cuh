// Register-tiled tile kernel (the fast path of qvm_apply_group).
//
// A CTA owns a tile of 2^k amplitudes (k tile bits, 9 <= k <= 13) and runs 2^(k-4) threads; each
// thread keeps 16 amplitudes in registers.  The group's gates are split (host side) into ROUNDS.
// In a round, 4 of the tile bits are "register bits" (they index the 16 register slots of a
// thread) and the other k-4 are "thread bits".  A gate whose dense targets are register bits is
// applied with no addressing at all; controls and diagonal factors may sit on register, thread or
// outer bits (thread/outer ones cost a predicate or one per-thread scalar).  Changing the set of
// register bits is a shared-memory transposition (XOR-swizzled against bank conflicts).  The first
// round loads global -> registers directly and the last stores registers -> global, with the lane
// bits on the five lowest tile bits so every warp access is a run of contiguous 16-byte elements.
//
// HBM traffic per launch: one read + one write of the shard, independent of the number of gates.
//
// The op stream is interpreted, so instruction economy per op matters more than anything else
// (profiles/): one 16-byte "hot" word per op, a dense opcode for a jump table, ops that are
// inactive for this CTA encoded as an impossible thread mask, everything CTA-uniform precomputed by
// the prologue.
#pragma once
#include "qvm_kernels.cuh"

namespace qvm {

// Dense opcode space (host assigns; the switch compiles to a jump table).
enum : uint32_t {
    OPC_X_RELABEL = 0,    // X / CNOT with no register-slot control: flips a bit of the slot relabel mask
    OPC_H0 = 1,           // +slot: uncontrolled Hadamard butterfly (scale folded into the round scalar)
    OPC_D1_0 = 5,         // +slot: generic 2x2
    OPC_XS_0 = 9,         // +slot: X with a control on a register slot: real pair swaps
    OPC_D2_0 = 13,        // +pair: generic 4x4 on slots (1,0) / (3,2)
    OPC_DIAG_T = 15,      // diagonal, no register-slot involvement: one scalar for the thread
    OPC_DIAG_S = 16,      // per-thread scalar on the slots selected by register-slot controls
    OPC_DIAG1_0 = 17,     // +slot: one diagonal target on a register slot
    OPC_DIAG_G = 21,      // generic diagonal (several register-slot targets or > 2 thread targets)
    OPC_DIAG_S1_0 = 22,   // +slot: DIAG_S whose only register-slot control is that slot (8 multiplies)
    OPC_DIAG1NC_0 = 26,   // +slot: DIAG1 without register-slot controls (16 multiplies, no predicates)
    OPC_COUNT = 30,
};

// x: [7:0] opcode | [11:8] crm (register-slot control mask) | [15:12] b (slot for X_RELABEL)
//    | [23:16] s0 | [31:24] s1           (index weights, as shifts, of the two thread-bit targets)
// y: [15:0] cthr (thread-bit control mask; 0xFFFF = op inactive for this CTA)
//    | [23:16] p0 | [31:24] p1           (thread-bit positions of those targets; 31 = none)
// z: mat_off (double2 units)
// w: [7:0] oi (index bits from outer targets, patched per CTA) | [15:8] shr (DIAG1: weight shift of
//    the register-slot target)
struct __align__(16) HotOp {
    uint32_t x, y, z, w;
};

// Read only by the prologue (one thread per op) and by the generic diagonal path.
struct __align__(16) ColdOp {
    uint64_t ctrl_outer;      // physical control mask outside the tile (incl. global positions)
    uint8_t loc[6];           // diagonal targets MSB first: class | index
    uint8_t k;
    uint8_t pad[1];
    uint64_t pad2[2];
};
static_assert(sizeof(ColdOp) == 32, "ColdOp layout");

enum : uint8_t { LOC_REG = 0x00, LOC_THR = 0x40, LOC_OUT = 0x80, LOC_CLASS = 0xC0, LOC_INDEX = 0x3F };

struct __align__(16) RoundDesc {
    uint8_t reg_bit[4];       // tile-local bit of register slot s
    uint8_t reg_sorted[4];    // the same four bits, ascending (for the insert-zero thread mapping)
    uint8_t reg_phys[4];      // physical position of register slot s (global I/O rounds)
    uint16_t op_begin, op_end;
    double scale;             // real scalar applied in this round (product of folded H scales)
    uint16_t dt_end;          // ops [op_begin, dt_end) are DIAG_T (they commute with the whole round)
    uint8_t pad[6];
};
static_assert(sizeof(RoundDesc) == 32, "RoundDesc layout");

struct RegTileParams {
    int n_local;
    int k;
    int low;                  // tile_pos[j] == j for j < low
    int n_rounds;
    int n_ops;
    int n_runs;               // runs of consecutive non-tile positions
    uint64_t shard_bits;
    uint64_t run_start;       // 8 x 8 bits: start position of each run
    uint64_t run_len;         // 8 x 8 bits: length of each run
    const RoundDesc* rounds;
    const HotOp* hot;
    const ColdOp* cold;
    const double2* mats;
    const uint64_t* hi_off;   // [2^(k-low)] global offset contributed by tile-index bits >= low
};

// Bank-conflict swizzle for the double2 tile: XOR index bits 3..6 into bits 0..2 so that (almost)
// any three "lowest thread bits" land in distinct 16-byte bank groups.  Linear over XOR.
__device__ __forceinline__ uint32_t swz(uint32_t i) {
    uint32_t x = i;
    x ^= ((i >> 3) & 1u) * 3u;
    x ^= ((i >> 4) & 1u) * 6u;
    x ^= ((i >> 5) & 1u) * 5u;
    x ^= ((i >> 6) & 1u) * 7u;
    return x;
}

__device__ __forceinline__ double flip_sign(double v, uint32_t sign_bit) {
    return __hiloint2double(__double2hiint(v) ^ (int)sign_bit, __double2loint(v));
}

// Physical register p holds logical slot (p ^ xm): thread-uniform X / CNOT only flips bits of xm.
// A register-slot control mask crm is tested on the logical index:
//   ((p ^ xm) & crm) == crm  <=>  (p & crm) == (crm & ~xm) =: want.

template <int B>
__device__ __forceinline__ void reg_h(double2 (&a)[16], const uint32_t sign_bit) {
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const int i0 = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1));
        const int i1 = i0 | (1 << B);
        const double2 x0 = a[i0], x1 = a[i1];
        a[i0] = make_double2(x0.x + x1.x, x0.y + x1.y);
        // logical (l0 - l1): the physical difference, negated when slots 0/1 are relabelled
        a[i1] = make_double2(flip_sign(x0.x - x1.x, sign_bit), flip_sign(x0.y - x1.y, sign_bit));
    }
}

template <int B>
__device__ __forceinline__ void reg_dense1(double2 (&a)[16], const double2* __restrict__ mat, const uint32_t crm,
                                           const uint32_t xm) {
    const uint32_t flip = ((xm >> B) & 1u) ? 3u : 0u;      // relabelled slots: index-flipped matrix
    const double2 m00 = __ldg(&mat[0 ^ flip]), m01 = __ldg(&mat[1 ^ flip]), m10 = __ldg(&mat[2 ^ flip]),
                  m11 = __ldg(&mat[3 ^ flip]);
    const uint32_t want = crm & ~xm;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const int i0 = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1));
        const int i1 = i0 | (1 << B);
        if ((i0 & crm) == want) {
            const double2 x0 = a[i0], x1 = a[i1];
            double2 y0 = cmul(m00, x0), y1 = cmul(m10, x0);
            y0 = cfma(m01, x1, y0);
            y1 = cfma(m11, x1, y1);
            a[i0] = y0;
            a[i1] = y1;
        }
    }
}

template <int B>
__device__ __forceinline__ void reg_xswap(double2 (&a)[16], const uint32_t crm, const uint32_t xm) {
    const uint32_t want = crm & ~xm;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const int i0 = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1));
        const int i1 = i0 | (1 << B);
        if ((i0 & crm) == want) {
            const double2 x0 = a[i0];
            a[i0] = a[i1];
            a[i1] = x0;
        }
    }
}

// pair P: slots (msb = 2P+1, lsb = 2P)
template <int P>
__device__ __forceinline__ void reg_dense2(double2 (&a)[16], const double2* __restrict__ m, const uint32_t crm,
                                           const uint32_t xm) {
    constexpr int LO = 2 * P, HI = 2 * P + 1;
    const uint32_t x2 = (xm >> LO) & 3u;
    const uint32_t want = crm & ~xm;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i0 = (P == 0) ? (q << 2) : q;      // the two slot bits outside the pair
        if ((i0 & crm) == want) {
            const int idx[4] = {i0, i0 | (1 << LO), i0 | (1 << HI), i0 | (1 << LO) | (1 << HI)};
            const double2 x0 = a[idx[0]], x1 = a[idx[1]], x2v = a[idx[2]], x3 = a[idx[3]];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double2* row = m + ((r ^ x2) << 2);          // index-flipped matrix for relabelled slots
                double2 acc = cmul(__ldg(&row[0 ^ x2]), x0);
                acc = cfma(__ldg(&row[1 ^ x2]), x1, acc);
                acc = cfma(__ldg(&row[2 ^ x2]), x2v, acc);
                acc = cfma(__ldg(&row[3 ^ x2]), x3, acc);
                a[idx[r]] = acc;
            }
        }
    }
}

template <int B>
__device__ __forceinline__ void reg_diag1(double2 (&a)[16], double2 d0, double2 d1, const uint32_t crm,
                                          const uint32_t xm) {
    if ((xm >> B) & 1u) {
        const double2 t = d0;
        d0 = d1;
        d1 = t;
    }
    const uint32_t want = crm & ~xm;
#pragma unroll
    for (int p = 0; p < 16; ++p)
        if ((p & crm) == want) a[p] = cmul(((p >> B) & 1) ? d1 : d0, a[p]);
}

// scalar d on the 8 slots whose logical bit B is 1 (physical bit B is 0 when the slot is relabelled)
template <int B>
__device__ __forceinline__ void reg_diag_s1(double2 (&a)[16], const double2 d, const uint32_t xm) {
    if ((xm >> B) & 1u) {
#pragma unroll
        for (int p = 0; p < 16; ++p)
            if (!((p >> B) & 1)) a[p] = cmul(d, a[p]);
    } else {
#pragma unroll
        for (int p = 0; p < 16; ++p)
            if ((p >> B) & 1) a[p] = cmul(d, a[p]);
    }
}

// d0 / d1 on every slot by its logical bit B, no controls
template <int B>
__device__ __forceinline__ void reg_diag1_nc(double2 (&a)[16], double2 d0, double2 d1, const uint32_t xm) {
    if ((xm >> B) & 1u) {
        const double2 t = d0;
        d0 = d1;
        d1 = t;
    }
#pragma unroll
    for (int p = 0; p < 16; ++p) a[p] = cmul(((p >> B) & 1) ? d1 : d0, a[p]);
}

__host__ __device__ inline size_t regtile_smem_bytes(int k, int n_rounds, int n_ops) {
    size_t b = n_rounds > 1 ? ((size_t)16 << k) : 0;
    b += (size_t)n_ops * sizeof(HotOp);
    b += (size_t)n_rounds * sizeof(RoundDesc);
    return b;
}

// FULL = false drops the generic dense 2x2 / 4x4 and generic-diagonal handlers (the host launches
// that lean variant when a group has none of them): ~40% less code, friendlier to the I-cache.
template <int MAX_THREADS, int MIN_BLOCKS, bool FULL>
__global__ void __launch_bounds__(MAX_THREADS, MIN_BLOCKS)
regtile_kernel(double2* __restrict__ state, const RegTileParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const size_t tile_bytes = P.n_rounds > 1 ? ((size_t)16 << P.k) : 0;
    double2* tile = reinterpret_cast<double2*>(smem_raw);
    HotOp* hot = reinterpret_cast<HotOp*>(smem_raw + tile_bytes);
    RoundDesc* rounds = reinterpret_cast<RoundDesc*>(hot + P.n_ops);

    const uint32_t tid = threadIdx.x;
    const uint32_t nthreads = blockDim.x;

    // outer bits of this tile: deposit blockIdx into the runs of non-tile positions
    uint64_t base = 0;
    {
        uint64_t t = blockIdx.x;
        for (int r = 0; r < P.n_runs; ++r) {
            const uint32_t len = (uint32_t)(P.run_len >> (8 * r)) & 0xFFu;
            base |= (t & ((1ull << len) - 1ull)) << ((uint32_t)(P.run_start >> (8 * r)) & 0xFFu);
            t >>= len;
        }
    }
    const uint64_t full = base | P.shard_bits;

    // stage descriptors; one thread per op evaluates what is uniform over the CTA
    for (uint32_t j = tid; j < (uint32_t)P.n_ops; j += nthreads) {
        uint4 h = __ldg(reinterpret_cast<const uint4*>(&P.hot[j]));
        const ColdOp* c = &P.cold[j];
        const uint64_t co = __ldg(&c->ctrl_outer);
        if ((full & co) != co) h.y |= 0xFFFFu;           // inactive here: impossible thread mask
        const uint32_t opc = h.x & 0xFFu;
        if (opc >= OPC_DIAG_T) {
            const uint32_t kk = c->k;
            uint32_t oi = 0;
            for (uint32_t t = 0; t < kk; ++t) {
                const uint32_t loc = c->loc[t];
                if ((loc & LOC_CLASS) == LOC_OUT) oi |= (uint32_t)((full >> (loc & LOC_INDEX)) & 1ull) << (kk - 1 - t);
            }
            h.w |= oi;
        }
        *reinterpret_cast<uint4*>(&hot[j]) = h;
    }
    {
        const uint32_t rwords = (uint32_t)(P.n_rounds * (sizeof(RoundDesc) / 16));
        const uint4* rsrc = reinterpret_cast<const uint4*>(P.rounds);
        uint4* rdst = reinterpret_cast<uint4*>(rounds);
        for (uint32_t w = tid; w < rwords; w += nthreads) rdst[w] = __ldg(&rsrc[w]);
    }
    __syncthreads();

    double2 a[16];
    double2* gbase = state + base;
    const uint32_t low_mask = (1u << P.low) - 1u;

    for (int r = 0; r < P.n_rounds; ++r) {
        const uint4 rw = *reinterpret_cast<const uint4*>(&rounds[r]);      // reg_bit | reg_sorted | reg_phys | op range
        const uint2 rw2 = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(&rounds[r]) + 16);
        // tile index of this thread: tid with zeros inserted at the (ascending) register-bit positions
        uint32_t T = tid;
        T = insert_zero32(T, rw.y & 0xFFu);
        T = insert_zero32(T, (rw.y >> 8) & 0xFFu);
        T = insert_zero32(T, (rw.y >> 16) & 0xFFu);
        T = insert_zero32(T, rw.y >> 24);
        const uint32_t J0 = 1u << (rw.x & 0xFFu), J1 = 1u << ((rw.x >> 8) & 0xFFu), J2 = 1u << ((rw.x >> 16) & 0xFFu),
                       J3 = 1u << (rw.x >> 24);
        const uint32_t op_begin = rw.w & 0xFFFFu, op_end = rw.w >> 16;
        const uint32_t dt_end = *reinterpret_cast<const uint16_t*>(reinterpret_cast<const unsigned char*>(&rounds[r]) + 24);

        if (r == 0) {
            const uint64_t gT = (uint64_t)(T & low_mask) | __ldg(&P.hi_off[T >> P.low]);
            const uint64_t g0 = 1ull << (rw.z & 0xFFu), g1 = 1ull << ((rw.z >> 8) & 0xFFu),
                           g2 = 1ull << ((rw.z >> 16) & 0xFFu), g3 = 1ull << (rw.z >> 24);
            const double2* src = gbase + gT;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                a[j] = src[((j & 1) ? g0 : 0) | ((j & 2) ? g1 : 0) | ((j & 4) ? g2 : 0) | ((j & 8) ? g3 : 0)];
        } else {
            const uint32_t sT = swz(T), s0 = swz(J0), s1 = swz(J1), s2 = swz(J2), s3 = swz(J3);
#pragma unroll
            for (int j = 0; j < 16; ++j)
                a[j] = tile[sT ^ ((j & 1) ? s0 : 0) ^ ((j & 2) ? s1 : 0) ^ ((j & 4) ? s2 : 0) ^ ((j & 8) ? s3 : 0)];
        }

        const double rscale = __hiloint2double((int)rw2.y, (int)rw2.x);
        double2 ts = make_double2(rscale, 0.0);   // per-thread scalar: folded H scales and DIAG_T factors
        bool ts_dirty = rscale != 1.0;
        uint32_t xm = 0;                          // physical register p holds logical slot p ^ xm

        // DIAG_T ops touch no register qubit of this round, so they commute with all of its ops:
        // the host sorts them first and they run as one tight batch into the thread scalar.
#pragma unroll 1
        for (uint32_t o = op_begin; o < dt_end; ++o) {
            const uint4 w = *reinterpret_cast<const uint4*>(&hot[o]);
            const uint32_t cthr = w.y & 0xFFFFu;
            if ((tid & cthr) != cthr) continue;
            const uint32_t idx = (w.w & 0xFFu) | (((tid >> ((w.y >> 16) & 0xFFu)) & 1u) << ((w.x >> 16) & 0xFFu)) |
                                 (((tid >> (w.y >> 24)) & 1u) << (w.x >> 24));
            ts = cmul(__ldg(&P.mats[w.z + idx]), ts);
            ts_dirty = true;
        }

        if (ts_dirty) {          // scalars commute with everything: apply now, so ts is dead in the main loop
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = cmul(ts, a[j]);
        }

        for (uint32_t o = dt_end; o < op_end; ++o) {
            const uint4 w = *reinterpret_cast<const uint4*>(&hot[o]);
            const uint32_t cthr = w.y & 0xFFFFu;
            if ((tid & cthr) != cthr) continue;           // a control on a thread bit is 0, or op inactive here
            const uint32_t opc = w.x & 0xFFu;
            const uint32_t crm = (w.x >> 8) & 0xFu;
            const double2* mat = P.mats + w.z;
            if (opc == OPC_X_RELABEL) { xm ^= 1u << ((w.x >> 12) & 0xFu); continue; }
            if (opc < OPC_D1_0) {                 // Hadamard butterfly on slot opc - OPC_H0
                if (opc == OPC_H0 + 0) { reg_h<0>(a, (xm & 1u) << 31); xm &= ~1u; }
                else if (opc == OPC_H0 + 1) { reg_h<1>(a, (xm & 2u) << 30); xm &= ~2u; }
                else if (opc == OPC_H0 + 2) { reg_h<2>(a, (xm & 4u) << 29); xm &= ~4u; }
                else { reg_h<3>(a, (xm & 8u) << 28); xm &= ~8u; }
                continue;
            }
            if (FULL && opc >= OPC_D1_0 && opc < OPC_XS_0) {
                switch (opc - OPC_D1_0) {
                    case 0: reg_dense1<0>(a, mat, crm, xm); break;
                    case 1: reg_dense1<1>(a, mat, crm, xm); break;
                    case 2: reg_dense1<2>(a, mat, crm, xm); break;
                    default: reg_dense1<3>(a, mat, crm, xm); break;
                }
                continue;
            }
            if (FULL && (opc == OPC_D2_0 || opc == OPC_D2_0 + 1)) {
                if (opc == OPC_D2_0) reg_dense2<0>(a, mat, crm, xm);
                else reg_dense2<1>(a, mat, crm, xm);
                continue;
            }
            if (FULL && opc == OPC_DIAG_G) {
                const ColdOp* c = &P.cold[o];
                const uint32_t kk = c->k;
                uint32_t idx = w.w & 0xFFu;
                uint32_t rm0 = 0, rm1 = 0, rm2 = 0, rm3 = 0, rs0 = 0, rs1 = 0, rs2 = 0, rs3 = 0;
                for (uint32_t t = 0; t < kk; ++t) {
                    const uint32_t loc = c->loc[t];
                    const uint32_t cls = loc & LOC_CLASS;
                    const uint32_t sh = 1u << (kk - 1 - t);
                    if (cls == LOC_THR) idx |= ((tid >> (loc & LOC_INDEX)) & 1u) ? sh : 0u;
                    else if (cls == LOC_REG) {
                        const uint32_t slot = loc & LOC_INDEX;       // each slot appears at most once
                        if (slot == 0) { rm0 = 1u; rs0 = sh; }
                        else if (slot == 1) { rm1 = 2u; rs1 = sh; }
                        else if (slot == 2) { rm2 = 4u; rs2 = sh; }
                        else { rm3 = 8u; rs3 = sh; }
                    }
                }
                const uint32_t want = crm & ~xm;
#pragma unroll
                for (int p = 0; p < 16; ++p) {
                    if ((p & crm) == want) {
                        const uint32_t j = (uint32_t)p ^ xm;         // logical slot
                        const uint32_t ij = idx | ((j & rm0) ? rs0 : 0u) | ((j & rm1) ? rs1 : 0u) |
                                            ((j & rm2) ? rs2 : 0u) | ((j & rm3) ? rs3 : 0u);
                        a[p] = cmul(__ldg(&mat[ij]), a[p]);
                    }
                }
                continue;
            }
            switch (opc) {
                case OPC_XS_0 + 0: reg_xswap<0>(a, crm, xm); break;
                case OPC_XS_0 + 1: reg_xswap<1>(a, crm, xm); break;
                case OPC_XS_0 + 2: reg_xswap<2>(a, crm, xm); break;
                case OPC_XS_0 + 3: reg_xswap<3>(a, crm, xm); break;
                default: {      // thread-indexed diagonals: DIAG_T, DIAG_S, DIAG_S1 + slot, DIAG1 + slot, DIAG1NC + slot
                    // index bits from (at most two) thread-bit targets; position 31 means "none"
                    const uint32_t idx = (w.w & 0xFFu) | (((tid >> ((w.y >> 16) & 0xFFu)) & 1u) << ((w.x >> 16) & 0xFFu)) |
                                         (((tid >> (w.y >> 24)) & 1u) << (w.x >> 24));
                    const double2 d0 = __ldg(&mat[idx]);
                    if (opc >= OPC_DIAG_S1_0 && opc < OPC_DIAG_S1_0 + 4) {
                        switch (opc - OPC_DIAG_S1_0) {
                            case 0: reg_diag_s1<0>(a, d0, xm); break;
                            case 1: reg_diag_s1<1>(a, d0, xm); break;
                            case 2: reg_diag_s1<2>(a, d0, xm); break;
                            default: reg_diag_s1<3>(a, d0, xm); break;
                        }
                    } else if (opc >= OPC_DIAG1NC_0) {
                        const double2 d1 = __ldg(&mat[idx | (1u << ((w.w >> 8) & 0xFFu))]);
                        switch (opc - OPC_DIAG1NC_0) {
                            case 0: reg_diag1_nc<0>(a, d0, d1, xm); break;
                            case 1: reg_diag1_nc<1>(a, d0, d1, xm); break;
                            case 2: reg_diag1_nc<2>(a, d0, d1, xm); break;
                            default: reg_diag1_nc<3>(a, d0, d1, xm); break;
                        }
                    } else if (opc == OPC_DIAG_S) {
                        const uint32_t want = crm & ~xm;
#pragma unroll
                        for (int p = 0; p < 16; ++p)
                            if ((p & crm) == want) a[p] = cmul(d0, a[p]);
                    } else {
                        const double2 d1 = __ldg(&mat[idx | (1u << ((w.w >> 8) & 0xFFu))]);
                        switch (opc - OPC_DIAG1_0) {
                            case 0: reg_diag1<0>(a, d0, d1, crm, xm); break;
                            case 1: reg_diag1<1>(a, d0, d1, crm, xm); break;
                            case 2: reg_diag1<2>(a, d0, d1, crm, xm); break;
                            default: reg_diag1<3>(a, d0, d1, crm, xm); break;
                        }
                    }
                    break;
                }
            }
        }

        // recompute the thread mapping here rather than keeping it live across the op loop
        const uint4 sw_ = *reinterpret_cast<const uint4*>(&rounds[r]);
        uint32_t T2 = tid;
        T2 = insert_zero32(T2, sw_.y & 0xFFu);
        T2 = insert_zero32(T2, (sw_.y >> 8) & 0xFFu);
        T2 = insert_zero32(T2, (sw_.y >> 16) & 0xFFu);
        T2 = insert_zero32(T2, sw_.y >> 24);
        if (r == P.n_rounds - 1) {
            const uint64_t g0 = 1ull << (sw_.z & 0xFFu), g1 = 1ull << ((sw_.z >> 8) & 0xFFu),
                           g2 = 1ull << ((sw_.z >> 16) & 0xFFu), g3 = 1ull << (sw_.z >> 24);
            // register p holds logical slot p ^ xm; the slot -> address map is XOR-linear
            const uint64_t gT = ((uint64_t)(T2 & low_mask) | __ldg(&P.hi_off[T2 >> P.low])) ^
                                (((xm & 1) ? g0 : 0) | ((xm & 2) ? g1 : 0) | ((xm & 4) ? g2 : 0) | ((xm & 8) ? g3 : 0));
            double2* dst = gbase;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                dst[gT ^ (((j & 1) ? g0 : 0) | ((j & 2) ? g1 : 0) | ((j & 4) ? g2 : 0) | ((j & 8) ? g3 : 0))] = a[j];
        } else {
            // same mapping as this round's load: every thread rewrites exactly the slots it read
            const uint32_t s0 = swz(1u << (sw_.x & 0xFFu)), s1 = swz(1u << ((sw_.x >> 8) & 0xFFu)),
                           s2 = swz(1u << ((sw_.x >> 16) & 0xFFu)), s3 = swz(1u << (sw_.x >> 24));
            const uint32_t sT = swz(T2) ^ ((xm & 1) ? s0 : 0) ^ ((xm & 2) ? s1 : 0) ^ ((xm & 4) ? s2 : 0) ^ ((xm & 8) ? s3 : 0);
#pragma unroll
            for (int j = 0; j < 16; ++j)
                tile[sT ^ ((j & 1) ? s0 : 0) ^ ((j & 2) ? s1 : 0) ^ ((j & 4) ? s2 : 0) ^ ((j & 8) ? s3 : 0)] = a[j];
            __syncthreads();
        }
    }
}

}  // namespace qvm
