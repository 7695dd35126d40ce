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
#pragma once
#include "qvm_kernels.cuh"

namespace qvm {

enum : uint8_t {
    RK_DENSE1 = 1,   // generic 2x2 on register slot b0 (controls anywhere)
    RK_X = 2,        // swap the pair: free (slot relabelling) unless a control sits on a register slot
    RK_H = 3,        // uncontrolled (a+b, a-b) * s ; the scale goes into the thread scalar
    RK_DENSE2 = 4,   // generic 4x4 on slot pair b0: 0 -> (msb=1, lsb=0), 1 -> (msb=3, lsb=2)
    RK_DIAG_T = 5,   // diagonal with no register-slot involvement: one scalar for the whole thread
    RK_DIAG_S = 6,   // per-thread scalar applied to the slots selected by register-slot controls
    RK_DIAG1 = 7,    // exactly one diagonal target on register slot b0 (+ thread / outer targets)
    RK_DIAG_R = 8,   // generic: two or more diagonal targets on register slots
};

enum : uint8_t { LOC_REG = 0x00, LOC_THR = 0x40, LOC_OUT = 0x80, LOC_CLASS = 0xC0, LOC_INDEX = 0x3F };

// 32 bytes; the first 16 are everything the hot loop needs (one LDS.128).
struct __align__(16) RegOp {
    uint32_t w0;          // kind | b0 << 8 | k << 16 | ctrl_reg << 24
    uint32_t w1;          // ctrl_thr (16 bits) | loc[4] << 16 | loc[5] << 24
    uint32_t mat_off;     // double2 offset into mats
    uint32_t loc03;       // loc[0..3], one byte each: diagonal targets MSB first (class | index)
    uint64_t ctrl_outer;  // physical mask (non-tile local + global positions)
    uint64_t pad;
};
static_assert(sizeof(RegOp) == 32, "RegOp layout");

__host__ __device__ __forceinline__ uint32_t regop_loc(uint32_t loc03, uint32_t w1, int t) {
    return t < 4 ? ((loc03 >> (8 * t)) & 0xFFu) : ((w1 >> (16 + 8 * (t - 4))) & 0xFFu);
}

struct __align__(16) RoundDesc {
    uint8_t reg_bit[4];   // tile-local bit of register slot s
    uint8_t thr_bit[9];   // tile-local bit of thread bit b (ascending)
    uint8_t pad[3];
    uint16_t op_begin, op_end;
    uint8_t pad2[12];
};
static_assert(sizeof(RoundDesc) == 32, "RoundDesc layout");

struct RegTileParams {
    int n_local;
    int k;
    int n_rounds;
    int n_ops;
    uint64_t tile_mask;
    uint64_t shard_bits;
    uint8_t tile_pos[16];
    const RoundDesc* rounds;
    const RegOp* ops;
    const double2* mats;
};

// Bank-conflict swizzle for the double2 tile: XOR index bits 3..6 into bits 0..2 so that (almost)
// any three "lowest thread bits" land in distinct 16-byte bank groups.  Linear: sw(a|b) = sw(a)^sw(b)
// for disjoint a, b.
__device__ __forceinline__ uint32_t swz(uint32_t i) {
    uint32_t x = i;
    x ^= ((i >> 3) & 1u) * 3u;
    x ^= ((i >> 4) & 1u) * 6u;
    x ^= ((i >> 5) & 1u) * 5u;
    x ^= ((i >> 6) & 1u) * 7u;
    return x;
}

// Physical register p holds logical slot (p ^ xm): thread-uniform X / CNOT only flips bits of xm.
// A register-slot control mask crm is therefore tested on the logical index: ((p ^ xm) & crm) == crm.

template <int B>
__device__ __forceinline__ void reg_dense1(double2 (&a)[16], double2 m00, double2 m01, double2 m10, double2 m11,
                                           const uint32_t crm, const uint32_t xm) {
    if ((xm >> B) & 1u) {          // logical 0/1 live in physical 1/0: use the index-flipped matrix
        double2 t = m00; m00 = m11; m11 = t;
        t = m01; m01 = m10; m10 = t;
    }
    const uint32_t want = crm & ~xm;      // (p ^ xm) & crm == crm  <=>  p & crm == crm & ~xm
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
__device__ __forceinline__ void reg_x(double2 (&a)[16], const uint32_t crm, const uint32_t xm) {
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

// (l0 + l1, l0 - l1); afterwards logical 0 sits in physical 0 (the caller clears bit B of xm).
template <int B>
__device__ __forceinline__ void reg_h(double2 (&a)[16], const double sgn) {
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const int i0 = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1));
        const int i1 = i0 | (1 << B);
        const double2 x0 = a[i0], x1 = a[i1];
        a[i0] = make_double2(x0.x + x1.x, x0.y + x1.y);
        a[i1] = make_double2(sgn * (x0.x - x1.x), sgn * (x0.y - x1.y));
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

__host__ __device__ inline size_t regtile_smem_bytes(int k, int n_rounds, int n_ops) {
    size_t b = n_rounds > 1 ? ((size_t)16 << k) : 0;
    b += (size_t)n_ops * sizeof(RegOp);
    b += (size_t)n_rounds * sizeof(RoundDesc);
    b += (((size_t)n_ops * 2 + 15) / 16) * 16;
    return b;
}

template <int MAX_THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(MAX_THREADS, MIN_BLOCKS)
regtile_kernel(double2* __restrict__ state, const RegTileParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const size_t tile_bytes = P.n_rounds > 1 ? ((size_t)16 << P.k) : 0;
    double2* tile = reinterpret_cast<double2*>(smem_raw);
    RegOp* ops = reinterpret_cast<RegOp*>(smem_raw + tile_bytes);
    RoundDesc* rounds = reinterpret_cast<RoundDesc*>(ops + P.n_ops);
    uint8_t* op_active = reinterpret_cast<uint8_t*>(rounds + P.n_rounds);
    uint8_t* op_outer_idx = op_active + P.n_ops;

    const uint32_t tid = threadIdx.x;
    const uint32_t nthreads = blockDim.x;
    const int n_thr_bits = P.k - 4;

    uint64_t base = 0;
    {
        uint64_t t = blockIdx.x;
        for (int p = 0; p < P.n_local; ++p)
            if (!((P.tile_mask >> p) & 1ull)) {
                base |= (t & 1ull) << p;
                t >>= 1;
            }
    }
    const uint64_t full = base | P.shard_bits;

    {   // stage descriptors
        const uint32_t words = (uint32_t)(P.n_ops * (sizeof(RegOp) / 16));
        const uint4* src = reinterpret_cast<const uint4*>(P.ops);
        uint4* dst = reinterpret_cast<uint4*>(ops);
        for (uint32_t w = tid; w < words; w += nthreads) dst[w] = __ldg(&src[w]);
        const uint32_t rwords = (uint32_t)(P.n_rounds * (sizeof(RoundDesc) / 16));
        const uint4* rsrc = reinterpret_cast<const uint4*>(P.rounds);
        uint4* rdst = reinterpret_cast<uint4*>(rounds);
        for (uint32_t w = tid; w < rwords; w += nthreads) rdst[w] = __ldg(&rsrc[w]);
    }
    for (int j = tid; j < P.n_ops; j += nthreads) {
        const RegOp* g = &P.ops[j];
        const uint64_t co = g->ctrl_outer;
        op_active[j] = ((full & co) == co) ? 1 : 0;
        const uint32_t w0 = g->w0, w1 = g->w1, loc03 = g->loc03;
        const uint32_t kind = w0 & 0xFFu, k = (w0 >> 16) & 0xFFu;
        uint32_t oi = 0;
        if (kind >= RK_DIAG_T)
            for (uint32_t t = 0; t < k; ++t) {
                const uint32_t loc = regop_loc(loc03, w1, (int)t);
                if ((loc & LOC_CLASS) == LOC_OUT) oi |= (uint32_t)((full >> (loc & LOC_INDEX)) & 1ull) << (k - 1 - t);
            }
        op_outer_idx[j] = (uint8_t)oi;
    }
    __syncthreads();

    double2 a[16];
    double2* gbase = state + base;

    for (int r = 0; r < P.n_rounds; ++r) {
        const RoundDesc& rd = rounds[r];
        // tile-local index contributed by this thread, and by each register slot bit
        uint32_t T = 0;
        for (int b = 0; b < n_thr_bits; ++b)
            if ((tid >> b) & 1u) T |= 1u << rd.thr_bit[b];
        const uint32_t J0 = 1u << rd.reg_bit[0], J1 = 1u << rd.reg_bit[1], J2 = 1u << rd.reg_bit[2],
                       J3 = 1u << rd.reg_bit[3];

        if (r == 0) {
            uint64_t gT = 0;
            for (int b = 0; b < n_thr_bits; ++b)
                if ((tid >> b) & 1u) gT |= 1ull << P.tile_pos[rd.thr_bit[b]];
            const uint64_t g0 = 1ull << P.tile_pos[rd.reg_bit[0]], g1 = 1ull << P.tile_pos[rd.reg_bit[1]],
                           g2 = 1ull << P.tile_pos[rd.reg_bit[2]], g3 = 1ull << P.tile_pos[rd.reg_bit[3]];
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

        double2 ts = make_double2(1.0, 0.0);     // per-thread scalar accumulated by H scales / DIAG_T
        bool ts_dirty = false;
        uint32_t xm = 0;                          // physical register p holds logical slot p ^ xm

        for (int o = rd.op_begin; o < rd.op_end; ++o) {
            const uint4 w = *reinterpret_cast<const uint4*>(&ops[o]);
            const uint32_t kind = w.x & 0xFFu, b0 = (w.x >> 8) & 0xFFu, k = (w.x >> 16) & 0xFFu, crm = w.x >> 24;
            const uint32_t cthr = w.y & 0xFFFFu;
            const bool ok = op_active[o] && ((tid & cthr) == cthr);
            const double2* mat = P.mats + w.z;
            switch (kind) {
                case RK_X: {
                    if (crm == 0) {                      // thread-uniform: relabel, no data movement
                        if (ok) xm ^= 1u << b0;
                        break;
                    }
                    if (!ok) break;
                    switch (b0) {
                        case 0: reg_x<0>(a, crm, xm); break;
                        case 1: reg_x<1>(a, crm, xm); break;
                        case 2: reg_x<2>(a, crm, xm); break;
                        default: reg_x<3>(a, crm, xm); break;
                    }
                    break;
                }
                case RK_H: {
                    const double sgn = ((xm >> b0) & 1u) ? -1.0 : 1.0;
                    switch (b0) {
                        case 0: reg_h<0>(a, sgn); break;
                        case 1: reg_h<1>(a, sgn); break;
                        case 2: reg_h<2>(a, sgn); break;
                        default: reg_h<3>(a, sgn); break;
                    }
                    xm &= ~(1u << b0);
                    const double sc = __ldg(&mat[0]).x;
                    ts.x *= sc;
                    ts.y *= sc;
                    ts_dirty = true;
                    break;
                }
                case RK_DENSE1: {
                    if (!ok) break;
                    const double2 m00 = __ldg(&mat[0]), m01 = __ldg(&mat[1]), m10 = __ldg(&mat[2]), m11 = __ldg(&mat[3]);
                    switch (b0) {
                        case 0: reg_dense1<0>(a, m00, m01, m10, m11, crm, xm); break;
                        case 1: reg_dense1<1>(a, m00, m01, m10, m11, crm, xm); break;
                        case 2: reg_dense1<2>(a, m00, m01, m10, m11, crm, xm); break;
                        default: reg_dense1<3>(a, m00, m01, m10, m11, crm, xm); break;
                    }
                    break;
                }
                case RK_DENSE2: {
                    if (!ok) break;
                    if (b0 == 0) reg_dense2<0>(a, mat, crm, xm);
                    else reg_dense2<1>(a, mat, crm, xm);
                    break;
                }
                case RK_DIAG_T:
                case RK_DIAG_S:
                case RK_DIAG1: {
                    if (!ok) break;
                    uint32_t idx = op_outer_idx[o];
                    uint32_t sh1 = 0;                    // index weight of the register-slot target (DIAG1)
                    for (uint32_t t = 0; t < k; ++t) {
                        const uint32_t loc = regop_loc(w.w, w.y, (int)t);
                        const uint32_t bit = 1u << (k - 1 - t);
                        if ((loc & LOC_CLASS) == LOC_THR) idx |= ((tid >> (loc & LOC_INDEX)) & 1u) ? bit : 0u;
                        else if ((loc & LOC_CLASS) == LOC_REG) sh1 = bit;
                    }
                    if (kind == RK_DIAG_T) {
                        ts = cmul(__ldg(&mat[idx]), ts);
                        ts_dirty = true;
                        break;
                    }
                    const uint32_t want = crm & ~xm;
                    if (kind == RK_DIAG_S) {
                        const double2 d = __ldg(&mat[idx]);
#pragma unroll
                        for (int p = 0; p < 16; ++p)
                            if ((p & crm) == want) a[p] = cmul(d, a[p]);
                        break;
                    }
                    double2 d0 = __ldg(&mat[idx]), d1 = __ldg(&mat[idx | sh1]);
                    if ((xm >> b0) & 1u) {
                        const double2 t = d0; d0 = d1; d1 = t;
                    }
                    const uint32_t sm = 1u << b0;
#pragma unroll
                    for (int p = 0; p < 16; ++p)
                        if ((p & crm) == want) {
                            const bool hi = (p & sm) != 0;
                            const double2 d = make_double2(hi ? d1.x : d0.x, hi ? d1.y : d0.y);
                            a[p] = cmul(d, a[p]);
                        }
                    break;
                }
                case RK_DIAG_R: {
                    if (!ok) break;
                    uint32_t idx = op_outer_idx[o];
                    uint32_t rm0 = 0, rm1 = 0, rm2 = 0, rm3 = 0, rs0 = 0, rs1 = 0, rs2 = 0, rs3 = 0;
                    for (uint32_t t = 0; t < k; ++t) {
                        const uint32_t loc = regop_loc(w.w, w.y, (int)t);
                        const uint32_t cls = loc & LOC_CLASS;
                        const uint32_t sh = 1u << (k - 1 - t);
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
                    break;
                }
                default: break;
            }
        }

        if (ts_dirty) {
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = cmul(ts, a[j]);
        }

        if (r == P.n_rounds - 1) {
            uint64_t gT = 0;
            for (int b = 0; b < n_thr_bits; ++b)
                if ((tid >> b) & 1u) gT |= 1ull << P.tile_pos[rd.thr_bit[b]];
            const uint64_t g0 = 1ull << P.tile_pos[rd.reg_bit[0]], g1 = 1ull << P.tile_pos[rd.reg_bit[1]],
                           g2 = 1ull << P.tile_pos[rd.reg_bit[2]], g3 = 1ull << P.tile_pos[rd.reg_bit[3]];
            // register p holds logical slot p ^ xm; the slot -> address map is XOR-linear
            const uint64_t gx = ((xm & 1) ? g0 : 0) | ((xm & 2) ? g1 : 0) | ((xm & 4) ? g2 : 0) | ((xm & 8) ? g3 : 0);
            double2* dst = gbase + gT;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                dst[gx ^ (((j & 1) ? g0 : 0) | ((j & 2) ? g1 : 0) | ((j & 4) ? g2 : 0) | ((j & 8) ? g3 : 0))] = a[j];
        } else {
            // same mapping as this round's load: every thread rewrites exactly the slots it read
            const uint32_t s0 = swz(J0), s1 = swz(J1), s2 = swz(J2), s3 = swz(J3);
            const uint32_t sT = swz(T) ^ ((xm & 1) ? s0 : 0) ^ ((xm & 2) ? s1 : 0) ^ ((xm & 4) ? s2 : 0) ^ ((xm & 8) ? s3 : 0);
#pragma unroll
            for (int j = 0; j < 16; ++j)
                tile[sT ^ ((j & 1) ? s0 : 0) ^ ((j & 2) ? s1 : 0) ^ ((j & 4) ? s2 : 0) ^ ((j & 8) ? s3 : 0)] = a[j];
            __syncthreads();
        }
    }
}

}  // namespace qvm
