// Register-tiled tile kernel (the fast path of qvm_apply_group).
//
// A CTA owns a tile of 2^k amplitudes (k tile bits, 9 <= k <= 13) and runs 2^(k-R) threads; each
// thread keeps 2^R amplitudes in registers (R = 4: 16 amplitudes, the default; R = 5: 32).  The group's gates are split (host side, qvm_encode.h)
// into ROUNDS.  In a round, R of the tile bits are "register bits" (they index the 2^R register
// slots of a thread) and the other k-R are "thread bits".  A gate whose dense targets are register
// bits is applied with no addressing at all; controls and diagonal factors may sit on register,
// thread or outer bits (thread/outer ones cost a predicate or one per-thread scalar).  Changing the
// set of register bits is a shared-memory transposition (XOR-swizzled against bank conflicts).
// The first round loads global -> registers directly and the last stores registers -> global, with
// the lane bits on the five lowest tile bits so every warp access is a run of contiguous 16-byte
// elements.
//
// HBM traffic per launch: one read + one write of the shard, independent of the number of gates.
//
// The op stream is interpreted, and round-1 profiles (profiles/README.md) showed the interpreter,
// not HBM, bounds the kernel.  Hence this layout:
//   * the prologue resolves everything that is uniform over the CTA (outer controls, outer index
//     bits) and writes the op stream to shared memory as 16-byte words: a plain op is a header plus
//     the (at most two) complex factors it can need -- no global loads and no pointer chasing in
//     the op loop; generic ops (dense matrices, wide diagonals) still index the global matrix pool;
//   * single-slot ops (H, the 8-amplitude phase S1, X relabels) commute across slots and are packed
//     by the host into LAYER records: per slot [S1] -> [H] -> [X], up to 12 gates per dispatch, the
//     thread-control tests of all sub-ops done by two vectorised mask operations;
//   * the next op's header is fetched while the current op executes;
//   * handlers are in place on the register file: the Hadamard butterfly is a += b; b = a - 2b, the
//     complex multiply is inline PTX that reuses its operands' registers (ptxas otherwise rotates
//     the amplitudes through temporaries and pays two moves per multiply);
//   * all per-round address ingredients (insert masks, swizzled strides) come precomputed and the
//     global addresses of a thread's 16 amplitudes are built by an add tree;
//   * rare op shapes live in the FULL template variant only, the lean variant stays small for the
//     instruction cache.
//
// The per-thread round body is __host__ __device__: tests/native/emu_regtile.cu runs the very same
// code on the CPU (threads looped sequentially between barriers) so that the planner, the encoder
// and the interpreter can be diffed against the oracle without a GPU.  That emulator is test
// infrastructure and is not part of libqvm_b200.so.
#pragma once
#if defined(__CUDACC_RTC__)
// compiled at run time by NVRTC (qvm_jit.h): no host headers, the vector types are built in
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#endif

#if defined(__CUDACC__)
#define QVM_HD __host__ __device__ __forceinline__
#else
#define QVM_HD inline
#endif

// absent sub-ops are the common case: keep their tests fall-through and the bodies out of line
#define QVM_UNLIKELY(x) __builtin_expect(!!(x), 0)

#if defined(__CUDA_ARCH__)
#define QVM_LDG(p) __ldg(p)
#else
#define QVM_LDG(p) (*(p))
#endif

namespace qvm {

// Opcodes.  "+slot" families reserve 5 codes (R <= 5 register slots).
enum : uint32_t {
    // single-slot ops: the encoder always packs them into LAYERs; these codes only label them on the host
    OPC_X_RELABEL = 0,    // X / CNOT with no register-slot control: flips a bit of the slot relabel mask
    OPC_H0 = 1,           // +slot: uncontrolled Hadamard butterfly (scale folded into the round scalar)
    OPC_S1_0 = 6,         // +slot: factor f on the slots whose logical bit is 1
    OPC_XS_0 = 11,        // +slot: X on that slot with register-slot controls crm: pair swaps
    OPC_DS = 16,          // factor f on the slots selected by the register-slot control mask crm
    OPC_DST = 17,         // DS whose factor is f0 / f1 by thread bit p0 (FULL)
    OPC_D1_0 = 18,        // +slot: (d0, d1) by the logical bit of that slot, optional crm (FULL)
    OPC_DIAG_G = 23,      // generic diagonal through the matrix pool (FULL)
    OPC_M1_0 = 24,        // +slot: dense 2x2 from the pool (FULL)
    OPC_M2_0 = 29,        // +pair: dense 4x4 on slots (1,0) / (3,2) from the pool (FULL)
    OPC_T = 31,           // thread scalar: f0 / f1 by one thread bit (only in the thread-scalar batch)
    OPC_TG = 32,          // thread scalar through the pool: two thread-bit targets (thread-scalar batch)
    OPC_LAYER = 33,       // up to 3R single-slot ops in one dispatch: per slot [S1] -> [H] -> [X relabel]
    OPC_COUNT = 34,
};
constexpr int kMaxRegBits = 5;

// The op stream is a sequence of 16-byte words in shared memory.  A plain op is 3 words
// (header, f0, f1); a LAYER is header + thread-control masks (1 word for R = 4, 2 for R = 5) + one
// S1 factor per slot.
//
// Header word of an op.
// x: [7:0] opcode | [12:8] crm (register-slot control mask) | [15:13] xb (slot of X_RELABEL)
//    | [23:16] p0 | [31:24] p1     thread-bit positions of up to two thread targets (31 = none)
//    LAYER: [7:0] opcode | [12 + s] slot s has an H
// y: [15:0] cthr (thread-bit control mask; 0xFFFF = op inactive for this CTA) | [16] HAS_F1
//    | [23:20] size of this op in words
// z: mat_off (double2 units into the pool; generic ops)
// w: [7:0] oi (pool index bits from outer targets, patched per CTA) | [15:8] s0 | [23:16] s1
//    (index weights, as shifts, of the thread targets p0 / p1) | [31:24] shr (weight shift of the
//    register-slot target, D1 / DIAG_G)
// LAYER word 1: x = cthr of the S1 on slot 0 | slot 1 << 16, y = slots 2 | 3, z / w = the same for X;
// word 2 (R = 5 only): x = S1 mask of slot 4 | X mask of slot 4 << 16.  0xFFFF = no such sub-op.
constexpr uint32_t HDR_HAS_F1 = 1u << 16;
constexpr uint32_t HDR_SIZE_SHIFT = 20;
constexpr int kOpWords = 3;
__host__ __device__ constexpr int layer_words(int R) { return R == 5 ? 8 : 6; }
__host__ __device__ constexpr int layer_factor_word(int R) { return R == 5 ? 3 : 2; }

union __align__(16) OpWord {      // one 16-byte word of the op stream
    uint4 u;
    double2 d;
};
static_assert(sizeof(OpWord) == 16, "OpWord layout");

// Read only by the prologue (one thread per op) and by the generic diagonal path.
struct __align__(16) ColdOp {
    uint64_t ctrl_outer;      // physical control mask outside the tile (incl. global positions)
    uint8_t loc[6];           // diagonal targets MSB first: class | index
    uint8_t k;                // number of diagonal targets
    uint8_t fsel;             // weight shift of the one non-outer target that selects f0 / f1; 0xFF = none
    uint32_t mat_off;         // pool offset (double2 units) of the op's diagonal / matrix
    uint8_t prefetch;         // 1: the prologue copies pool[mat_off + oi (| 1 << fsel)] into f0 (f1)
    uint8_t n_words;          // words this op occupies in the stream
    uint8_t kind;             // COLD_*
    uint8_t slot;             // patches: the layer slot they belong to
    uint32_t word_off;        // first word of the op (patches: of the layer they patch)
    uint32_t pad1;
};
static_assert(sizeof(ColdOp) == 32, "ColdOp layout");

// plain op | layer (verbatim copy) | patches of a layer's S1 / X sub-op that depends on outer bits
enum : uint8_t { COLD_PLAIN = 0, COLD_LAYER = 1, COLD_PATCH_S1 = 2, COLD_PATCH_X = 3 };
enum : uint8_t { LOC_REG = 0x00, LOC_THR = 0x40, LOC_OUT = 0x80, LOC_CLASS = 0xC0, LOC_INDEX = 0x3F };

struct __align__(16) RoundDesc {
    double2 scale;            //  0: complex scalar applied in this round (folded H scales, RZ phases)
    uint16_t swz_j[5];        // 16: swz(1 << tile-local bit of slot s): shared-memory stride of slot s
    uint16_t op_begin, dt_end, op_end;   // 26: WORD offsets: [op_begin, dt_end) thread-scalar batch; [dt_end, op_end) slot ops
    uint16_t ins_mask[5];     // 32: ~((1 << b) - 1) for the register bits b ascending: x + (x & mask) inserts a zero bit
    uint16_t flags;           // 42: bit 0: scale != 1; bits 15..1: number of OPC_TG ops at the end of the batch
    uint8_t reg_phys[5];      // 44: physical position of register slot s (global I/O rounds)
    uint8_t pad0;
    uint16_t wr_swz[5];       // 50: relabelling round only: shared-memory stride of slot s on the WRITE side
    uint8_t pad[4];
};
static_assert(sizeof(RoundDesc) == 64, "RoundDesc layout");

// Fused global<->local qubit exchange (multi-GPU): the pass reads its tiles from this shard and STORES
// them straight into their owners' shards over NVLink peer memory.  Swapping the m global bits of the
// rank with the m local "victim" bits pos[] sends the amplitude (rank R, local index x) to the rank whose
// global bits equal x's victim bits, at local index x with the victim bits replaced by R's global bits:
// dst[L] is the (alternate) buffer of the rank with global bits L (dst[own bits] = this rank's own
// alternate buffer), mine_bits = R's global bits deposited at pos[].  The victim positions are outer
// bits of every tile, so a whole tile goes to one destination: one pointer select per CTA.  m == 0:
// the pass stores in place.
constexpr int kMaxPushBits = 3;
struct PushSpec {
    int m;
    uint8_t pos[4];
    uint64_t mask;
    uint64_t mine_bits;
    uint64_t rot;             // XORed into every tile's outer bits: rotates the order in which a rank visits its
                              // destinations (victim bits slowest in the tile index, rot = mine_bits: at any time
                              // every rank stores to a different peer), 0 = tiles in plain index order
    double2* dst[1 << kMaxPushBits];
};

struct RegTileParams {
    int n_local;
    int k;
    int R;                    // register bits per round (4 or 5)
    int low;                  // tile_pos[j] == j for j < low
    int n_rounds;
    int n_ops;                // cold entries that produce stream words (plain ops and layers)
    int n_patches;            // cold entries [n_ops, n_ops + n_patches) patch layer sub-ops afterwards
    int n_words;              // length of the op stream in 16-byte words
    int n_runs;               // runs of consecutive non-tile positions
    uint8_t first_reg_phys[8];   // round 0: physical positions of the register slots ...
    uint16_t first_ins[8];       // ... and its insert masks: the tile load needs nothing staged
    // In-tile qubit relabelling (qvm_encode.h): round `perm_round` (0, or -1: none) writes the amplitude
    // of tile index i to shared-memory slot swz(sigma(i)); the later rounds run in the new tile
    // coordinates, so the qubit of tile bit b leaves the pass on tile bit sigma(b).
    int perm_round;
    // The reset folded into the first pass: the shard is known to be all zeros (2: plus amp[0] = 1 on this shard),
    // so round 0 does not read HBM at all -- the reference's `wf = zeros; wf[0] = 1` (qvm_wavefunction.py:354-355)
    // followed by its first gate costs one write of the shard instead of a write, a read and a write.
    int from_zero;
    uint16_t perm_swz[16];    // swz(1 << sigma(b)) for every tile bit b
    PushSpec push;            // where the last round stores (fused exchange); m == 0: in place
    uint64_t shard_bits;
    uint64_t run_start;       // 8 x 8 bits: start position of each run
    uint64_t run_len;         // 8 x 8 bits: length of each run
    const RoundDesc* rounds;
    const uint4* hot;         // [n_words] the op stream as encoded on the host
    const ColdOp* cold;       // [n_ops + n_patches]
    const double2* mats;      // matrix pool
    const uint64_t* hi_off;   // [2^(k-low)] global offset contributed by tile-index bits >= low
};

#if defined(__CUDA_ARCH__)
// 16-byte asynchronous copy global -> shared (L2 only), used by the pipelined specialised kernels
__device__ __forceinline__ void cp_async16(double2* smem_dst, const double2* gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}

// ---- TMA bulk copies + mbarrier (the TMA-pipelined specialised kernels, qvm_jit.h) --------------------------
// A tile is 2^(k-low) runs of 2^low consecutive amplitudes (256 or 512 bytes): every thread moves ONE run with
// one cp.async.bulk.  Loads land in shared memory without occupying registers or the LSU pipe and signal an
// mbarrier by byte count; stores drain shared memory to global (or peer) memory asynchronously.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\t"
                 "bra WAIT_%=;\n\t"
                 "DONE_%=:\n\t"
                 "}" ::"r"(addr), "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void tma_load_run(double2* smem_dst, const double2* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_run(double2* gmem_dst, const double2* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// ---- small helpers ----------------------------------------------------------------------------
QVM_HD double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a <- f * a with two temporaries and no moves: the imaginary part is finished before the real part is
// overwritten
QVM_HD void cmul_ip(double2& a, const double2 f) {
#if defined(__CUDA_ARCH__)
    // same virtual register as source and destination: keeps ptxas from rotating the amplitudes
    // through temporaries (which costs two moves per multiply at every handler exit)
    asm("{\n\t"
        ".reg .f64 t;\n\t"
        "mul.rn.f64 t, %1, %3;\n\t"
        "mul.rn.f64 %1, %1, %2;\n\t"
        "fma.rn.f64 %1, %0, %3, %1;\n\t"
        "neg.f64 t, t;\n\t"
        "fma.rn.f64 %0, %0, %2, t;\n\t"
        "}"
        : "+d"(a.x), "+d"(a.y)
        : "d"(f.x), "d"(f.y));
#else
    const double t1 = a.y * f.y, t2 = a.y * f.x;
    a.y = fma(a.x, f.y, t2);
    a.x = fma(a.x, f.x, -t1);
#endif
}
QVM_HD double2 cfma(double2 m, double2 a, double2 acc) {
    acc.x = fma(m.x, a.x, acc.x);
    acc.x = fma(-m.y, a.y, acc.x);
    acc.y = fma(m.x, a.y, acc.y);
    acc.y = fma(m.y, a.x, acc.y);
    return acc;
}
QVM_HD uint32_t insert_zero32(uint32_t x, uint32_t pos) {
    const uint32_t lo = x & ((1u << pos) - 1u);
    return ((x >> pos) << (pos + 1)) | lo;
}
QVM_HD uint64_t insert_zero64(uint64_t x, uint32_t pos) {
    const uint64_t lo = x & ((1ull << pos) - 1ull);
    return ((x >> pos) << (pos + 1)) | lo;
}

// Bank-conflict swizzle for the double2 tile: XOR index bits >= 3 into bits 0..2 so that (almost)
// any three "lowest thread bits" land in distinct 16-byte bank groups.  Linear over XOR.  Bits 3..6 get
// the four remaining non-zero patterns.  Bits 7..12 only become lane bits on the write side of a
// relabelling round, next to the lowest old lane bits (patterns 1, 2): they get patterns with bit 2 set
// (any of those completes {1, 2} to a basis, and any three distinct ones are a basis themselves); the
// encoder picks the pairing.
QVM_HD uint32_t swz(uint32_t i) {
    uint32_t x = i;
    x ^= ((i >> 3) & 1u) * 3u;
    x ^= ((i >> 4) & 1u) * 6u;
    x ^= ((i >> 5) & 1u) * 5u;
    x ^= ((i >> 6) & 1u) * 7u;
    x ^= ((i >> 7) & 1u) * 4u;
    x ^= ((i >> 8) & 1u) * 6u;
    x ^= ((i >> 9) & 1u) * 5u;
    x ^= ((i >> 10) & 1u) * 7u;
    x ^= ((i >> 11) & 1u) * 4u;
    x ^= ((i >> 12) & 1u) * 6u;
    return x;
}

__host__ __device__ inline size_t regtile_smem_bytes(int k, int low, int n_rounds, int n_words) {
    size_t b = n_rounds > 1 ? ((size_t)16 << k) : 0;
    b += (size_t)(n_words + 1) * sizeof(OpWord);     // +1: the prefetch of the header after the last op
    b += (size_t)n_rounds * sizeof(RoundDesc);
    b += (size_t)8 << (k - low);
    return b;
}

// Per-CTA view handed to the round body.
struct RtCta {
    double2* tile;            // shared: 2^k amplitudes (unused when n_rounds == 1)
    double2* tile_w;          // where rounds write: the same array (the sequential CPU emulator alone reads the
                              // relabelling round from a snapshot, in place of that round's barrier)
    const OpWord* ops;        // shared: resolved op stream
    const RoundDesc* rounds;  // shared
    const uint64_t* hi_off;   // shared
    double2* gbase;           // state + outer-bit offset of this tile
    double2* gstore;          // where the last round stores this tile (== gbase unless the pass pushes, PushSpec)
    const double2* mats;      // global pool
    const ColdOp* cold;       // global
    int low;
    int n_rounds;
    int perm_round;           // see RegTileParams
    const uint16_t* perm_swz;
};

// ---- prologue: resolve op j for this CTA ---------------------------------------------------------
// `full` = physical index bits fixed for this CTA (outer tile position | shard bits).
QVM_HD void rt_resolve_op(const RegTileParams& P, uint32_t j, uint64_t full, OpWord* stream) {
    const ColdOp* c = &P.cold[j];
    const uint32_t w0 = c->word_off;
    OpWord* dst = stream + w0;
    if (c->kind == COLD_LAYER) {                   // outer-dependent sub-ops are patched afterwards
        const int nw = c->n_words;
        for (int i = 0; i < nw; ++i) dst[i].u = QVM_LDG(&P.hot[w0 + i]);
        return;
    }
    uint4 h = QVM_LDG(&P.hot[w0]);
    const uint64_t co = c->ctrl_outer;
    if ((full & co) != co) h.y |= 0xFFFFu;           // inactive here: impossible thread mask
    const uint32_t kk = c->k;
    uint32_t oi = 0;
    for (uint32_t t = 0; t < kk; ++t) {
        const uint32_t loc = c->loc[t];
        if ((loc & LOC_CLASS) == LOC_OUT) oi |= (uint32_t)((full >> (loc & LOC_INDEX)) & 1ull) << (kk - 1 - t);
    }
    h.w |= oi;
    // factors: written into the stream by the host when they do not depend on outer bits (no dependent
    // load here), fetched from the pool for this CTA's outer bits otherwise
    OpWord f0, f1;
    f0.u = QVM_LDG(&P.hot[w0 + 1]);
    f1.u = QVM_LDG(&P.hot[w0 + 2]);
    if (c->prefetch) {
        const double2* m = P.mats + c->mat_off;
        f0.d = QVM_LDG(&m[oi]);
        if (c->fsel != 0xFF) f1.d = QVM_LDG(&m[oi | (1u << c->fsel)]);
    }
    dst[0].u = h;
    dst[1] = f0;
    dst[2] = f1;
}

// Second prologue pass (after a barrier): a layer sub-op whose control sits outside the tile is
// switched off for CTAs where that control is 0 (impossible thread mask), and an S1 factor that
// depends on outer target bits is fetched for this CTA's bits.
QVM_HD void rt_patch_op(const RegTileParams& P, uint32_t j, uint64_t full, OpWord* stream) {
    const ColdOp* c = &P.cold[j];
    OpWord* layer = stream + c->word_off;
    const uint32_t slot = c->slot;
    // word 1: cs[0..3], cx[0..3]; word 2 (R = 5): cs[4], cx[4]
    uint16_t* m16 = reinterpret_cast<uint16_t*>(&layer[slot < 4 ? 1 : 2]);
    uint16_t* ms = slot < 4 ? &m16[slot] : &m16[0];
    uint16_t* mx = slot < 4 ? &m16[4 + slot] : &m16[1];
    const uint64_t co = c->ctrl_outer;
    const bool active = (full & co) == co;
    if (c->kind == COLD_PATCH_X) {
        if (!active) *mx = 0xFFFFu;
        return;
    }
    if (!active) {
        *ms = 0xFFFFu;
        return;
    }
    const uint32_t kk = c->k;
    if (kk) {
        uint32_t oi = 0;
        for (uint32_t t = 0; t < kk; ++t) {
            const uint32_t loc = c->loc[t];
            if ((loc & LOC_CLASS) == LOC_OUT) oi |= (uint32_t)((full >> (loc & LOC_INDEX)) & 1ull) << (kk - 1 - t);
        }
        layer[layer_factor_word(P.R) + slot].d = QVM_LDG(&P.mats[c->mat_off + oi]);
    }
}

// ---- slot handlers (all in place; NA = 2^R amplitudes per thread) ---------------------------------
// Physical register p holds logical slot (p ^ xm): thread-uniform X / CNOT only flips bits of xm.
// A register-slot control mask crm is tested on the logical index:
//   ((p ^ xm) & crm) == crm  <=>  (p & crm) == (crm & ~xm) =: want.

// Unnormalised Hadamard on slot B, in place and branch-free: a0 += a1; a1 = a0 - 2 a1 gives the
// physical (sum, difference).  If the slot is relabelled (physical 0 holds logical 1) the logical
// difference is the negated physical one: flip the sign bit of a1 and drop the relabel bit (the
// caller clears it) -- one integer op per double instead of a divergent pair of code variants.
QVM_HD double flip_sign(double v, uint32_t sign_bit) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) ^ (int)sign_bit, __double2loint(v));
#else
    return sign_bit ? -v : v;
#endif
}
template <int B, int NA>
QVM_HD void reg_h(double2 (&a)[NA], uint32_t& xm) {
    const uint32_t sign_bit = ((xm >> B) & 1u) << 31;
#pragma unroll
    for (int p = 0; p < NA / 2; ++p) {
        const int i0 = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1));
        const int i1 = i0 | (1 << B);
        a[i0].x += a[i1].x;
        a[i0].y += a[i1].y;
        a[i1].x = flip_sign(fma(-2.0, a[i1].x, a[i0].x), sign_bit);       // (x0 + x1) - 2 x1 = x0 - x1
        a[i1].y = flip_sign(fma(-2.0, a[i1].y, a[i0].y), sign_bit);
    }
    xm &= ~(1u << B);
}

template <int B, int NA>
QVM_HD void reg_xswap(double2 (&a)[NA], const uint32_t crm, const uint32_t xm) {
    const uint32_t want = crm & ~xm;
#pragma unroll
    for (int p = 0; p < NA / 2; ++p) {
        const int i0 = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1));
        const int i1 = i0 | (1 << B);
        if ((i0 & crm) == want) {
            const double2 x0 = a[i0];
            a[i0] = a[i1];
            a[i1] = x0;
        }
    }
}

// factor f on the NA / 2 slots whose logical bit B is 1
template <int B, int NA>
QVM_HD void reg_s1(double2 (&a)[NA], const double2 f, const uint32_t xm) {
    if ((xm >> B) & 1u) {
#pragma unroll
        for (int p = 0; p < NA; ++p)
            if (!((p >> B) & 1)) cmul_ip(a[p], f);
    } else {
#pragma unroll
        for (int p = 0; p < NA; ++p)
            if ((p >> B) & 1) cmul_ip(a[p], f);
    }
}

template <int B, int NA>
QVM_HD void reg_d1(double2 (&a)[NA], double2 d0, double2 d1, const uint32_t crm, const uint32_t xm) {
    if ((xm >> B) & 1u) {
        const double2 t = d0;
        d0 = d1;
        d1 = t;
    }
    const uint32_t want = crm & ~xm;
#pragma unroll
    for (int p = 0; p < NA; ++p)
        if ((p & crm) == want) cmul_ip(a[p], ((p >> B) & 1) ? d1 : d0);
}

template <int B, int NA>
QVM_HD void reg_dense1(double2 (&a)[NA], const double2* __restrict__ mat, const uint32_t crm, const uint32_t xm) {
    const uint32_t flip = ((xm >> B) & 1u) ? 3u : 0u;      // relabelled slots: index-flipped matrix
    const double2 m00 = QVM_LDG(&mat[0 ^ flip]), m01 = QVM_LDG(&mat[1 ^ flip]), m10 = QVM_LDG(&mat[2 ^ flip]),
                  m11 = QVM_LDG(&mat[3 ^ flip]);
    const uint32_t want = crm & ~xm;
#pragma unroll
    for (int p = 0; p < NA / 2; ++p) {
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

// pair P: slots (msb = 2P+1, lsb = 2P)
template <int P, int NA>
QVM_HD void reg_dense2(double2 (&a)[NA], const double2* __restrict__ m, const uint32_t crm, const uint32_t xm) {
    constexpr int LO = 2 * P, HI = 2 * P + 1;
    const uint32_t x2 = (xm >> LO) & 3u;
    const uint32_t want = crm & ~xm;
#pragma unroll
    for (int q = 0; q < NA / 4; ++q) {
        const int i0 = ((q >> LO) << (LO + 2)) | (q & ((1 << LO) - 1));      // the slot bits outside the pair
        if ((i0 & crm) == want) {
            const int idx[4] = {i0, i0 | (1 << LO), i0 | (1 << HI), i0 | (1 << LO) | (1 << HI)};
            const double2 x0 = a[idx[0]], x1 = a[idx[1]], x2v = a[idx[2]], x3 = a[idx[3]];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double2* row = m + ((r ^ x2) << 2);          // index-flipped matrix for relabelled slots
                double2 acc = cmul(QVM_LDG(&row[0 ^ x2]), x0);
                acc = cfma(QVM_LDG(&row[1 ^ x2]), x1, acc);
                acc = cfma(QVM_LDG(&row[2 ^ x2]), x2v, acc);
                acc = cfma(QVM_LDG(&row[3 ^ x2]), x3, acc);
                a[idx[r]] = acc;
            }
        }
    }
}

// XOR-combination of the strides selected by the bits of j (j is a compile-time constant at every use)
template <int R>
QVM_HD uint32_t xor_combo(const uint32_t (&st)[R], int j) {
    uint32_t x = 0;
#pragma unroll
    for (int s = 0; s < R; ++s)
        if ((j >> s) & 1) x ^= st[s];
    return x;
}

// Global addresses of a thread's 2^R amplitudes: register-bit positions carry zeros in the thread's
// own offset, so they are sums -- an add tree on byte offsets (two instructions per address).
template <int R, typename PTR>
QVM_HD void add_tree(PTR p0, const int64_t (&b)[R], PTR (&q)[1 << R]) {
    q[0] = p0;
#pragma unroll
    for (int s = 0; s < R; ++s)
#pragma unroll
        for (int j = 0; j < (1 << s); ++j) q[(1 << s) + j] = q[j] + b[s];
}

// Round 0's tile load: everything it needs comes from kernel parameters and one L2-resident table.
template <int R>
QVM_HD void rt_first_load(const RegTileParams& P, const double2* gbase, const uint32_t tid, double2 (&a)[1 << R],
                          const bool first_tile = false) {
    if (P.from_zero) {
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) a[j] = make_double2(0.0, 0.0);
        if (P.from_zero == 2 && first_tile && tid == 0) a[0].x = 1.0;       // thread 0, slot 0 of tile 0 = amplitude 0
        return;
    }
    uint32_t T = tid;
#pragma unroll
    for (int s = 0; s < R; ++s) T += T & P.first_ins[s];
    const uint64_t gT = (uint64_t)(T & ((1u << P.low) - 1u)) | QVM_LDG(&P.hi_off[T >> P.low]);
    int64_t b[R];
#pragma unroll
    for (int s = 0; s < R; ++s) b[s] = (int64_t)(16ull << P.first_reg_phys[s]);
    const char* q[1 << R];
    add_tree<R>(reinterpret_cast<const char*>(gbase + gT), b, q);
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) a[j] = *reinterpret_cast<const double2*>(q[j]);
}

// ---- one round of one thread ----------------------------------------------------------------------
// FULL = false drops the rare handlers (generic dense 2x2 / 4x4, generic and thread-selected
// diagonals; the host launches that lean variant when a group has none of them).
template <bool FULL, int R>
QVM_HD void rt_thread_round(const RtCta& c, const int r, const uint32_t tid, double2 (&a)[1 << R]) {
    constexpr int NA = 1 << R;
    const RoundDesc* rd = &c.rounds[r];
    const uint4 w1 = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(rd) + 16);
    const uint4 w2 = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(rd) + 32);
    // w1: swz_j[0..4] (x, y, z.lo) | op_begin (z.hi) | dt_end (w.lo) | op_end (w.hi)
    // w2: ins_mask[0..4] (x, y, z.lo) | flags (z.hi) | reg_phys[0..3] (w)
    const uint32_t op_begin = w1.z >> 16, dt_end = w1.w & 0xFFFFu, op_end = w1.w >> 16;
    const uint32_t flags = w2.z >> 16;
    const bool scaled = flags & 1u;
    const bool first = r == 0, last = r == c.n_rounds - 1;

    // tile index of this thread: tid with zeros inserted at the (ascending) register-bit positions
    uint32_t T = tid;
    T += T & (w2.x & 0xFFFFu);
    T += T & (w2.x >> 16);
    T += T & (w2.y & 0xFFFFu);
    T += T & (w2.y >> 16);
    if (R == 5) T += T & (w2.z & 0xFFFFu);
    uint32_t st[R];
    st[0] = w1.x & 0xFFFFu;
    st[1] = w1.x >> 16;
    st[2] = w1.y & 0xFFFFu;
    st[3] = w1.y >> 16;
    if (R == 5) st[R - 1] = w1.z & 0xFFFFu;
    const uint32_t sT = swz(T);
    const uint32_t low_mask = (1u << c.low) - 1u;

    if (!first) {        // (round 0's amplitudes were loaded from HBM by rt_first_load)
#pragma unroll
        for (int j = 0; j < NA; ++j) a[j] = c.tile[sT ^ xor_combo<R>(st, j)];
    }

    // Thread-scalar batch: these ops touch no register qubit of this round, so they commute with all
    // of its ops; the host sorts them first and they fold into one scalar per thread.
    double2 ts = rd->scale;
    bool ts_dirty = scaled;
    const uint32_t tg_begin = dt_end - kOpWords * (flags >> 1);          // the (rare) pool-indexed ops come last
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (const OpWord* op = &c.ops[op_begin]; op != &c.ops[tg_begin]; op += kOpWords) {
        const uint2 w = *reinterpret_cast<const uint2*>(op);         // OPC_T: f0 / f1 by thread bit p0
        if ((tid & w.y & 0xFFFFu) != (w.y & 0xFFFFu)) continue;
        cmul_ip(ts, op[1 + ((tid >> ((w.x >> 16) & 0xFFu)) & 1u)].d);
        ts_dirty = true;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (const OpWord* op = &c.ops[tg_begin]; op != &c.ops[dt_end]; op += kOpWords) {
        const uint4 w = op->u;                                       // OPC_TG: two thread-bit targets, pool factors
        if ((tid & w.y & 0xFFFFu) != (w.y & 0xFFFFu)) continue;
        const uint32_t idx = (w.w & 0xFFu) | (((tid >> ((w.x >> 16) & 0xFFu)) & 1u) << ((w.w >> 8) & 0xFFu)) |
                             (((tid >> (w.x >> 24)) & 1u) << ((w.w >> 16) & 0xFFu));
        cmul_ip(ts, QVM_LDG(&c.mats[w.z + idx]));
        ts_dirty = true;
    }
    if (ts_dirty) {          // scalars commute with everything: apply now, so ts is dead in the main loop
#pragma unroll
        for (int j = 0; j < NA; ++j) cmul_ip(a[j], ts);
    }

    uint32_t xm = 0;         // physical register p holds logical slot p ^ xm
    if (dt_end < op_end) {
        const OpWord* op = &c.ops[dt_end];
        const OpWord* const op_last = &c.ops[op_end];
        uint2 hn = *reinterpret_cast<const uint2*>(op);           // header words x, y of the NEXT op to run
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        while (op != op_last) {
            const OpWord* const cur = op;
            const uint32_t hx = hn.x, opc = hx & 0xFFu, cthr = hn.y & 0xFFFFu;
            op += (hn.y >> HDR_SIZE_SHIFT) & 0xFu;
            hn = *reinterpret_cast<const uint2*>(op);             // prefetch (one spare word follows the last op)
            if (opc == OPC_LAYER) {
                // per slot: [S1 factor] -> [H] -> [X relabel]; sub-ops on different slots commute.
                // (t & m) ^ m == 0  <=>  every thread-bit control of that sub-op is set; absent = 0xFFFF.
                constexpr int FW = layer_factor_word(R);
                const uint4 m = cur[1].u;
                const uint32_t t2 = tid | (tid << 16);
                const uint32_t s01 = (t2 & m.x) ^ m.x, s23 = (t2 & m.y) ^ m.y;
                const uint32_t x01 = (t2 & m.z) ^ m.z, x23 = (t2 & m.w) ^ m.w;
                if (QVM_UNLIKELY(!(s01 & 0xFFFFu))) reg_s1<0, NA>(a, cur[FW + 0].d, xm);
                if (QVM_UNLIKELY(hx & 0x1000u)) reg_h<0, NA>(a, xm);
                if (!(x01 & 0xFFFFu)) xm ^= 1u;
                if (QVM_UNLIKELY(!(s01 >> 16))) reg_s1<1, NA>(a, cur[FW + 1].d, xm);
                if (QVM_UNLIKELY(hx & 0x2000u)) reg_h<1, NA>(a, xm);
                if (!(x01 >> 16)) xm ^= 2u;
                if (QVM_UNLIKELY(!(s23 & 0xFFFFu))) reg_s1<2, NA>(a, cur[FW + 2].d, xm);
                if (QVM_UNLIKELY(hx & 0x4000u)) reg_h<2, NA>(a, xm);
                if (!(x23 & 0xFFFFu)) xm ^= 4u;
                if (QVM_UNLIKELY(!(s23 >> 16))) reg_s1<3, NA>(a, cur[FW + 3].d, xm);
                if (QVM_UNLIKELY(hx & 0x8000u)) reg_h<3, NA>(a, xm);
                if (!(x23 >> 16)) xm ^= 8u;
                if (R == 5) {
                    const uint32_t mb = cur[2].u.x;
                    const uint32_t sx4 = (t2 & mb) ^ mb;
                    if (QVM_UNLIKELY(!(sx4 & 0xFFFFu))) reg_s1<R - 1, NA>(a, cur[FW + R - 1].d, xm);
                    if (QVM_UNLIKELY(hx & 0x10000u)) reg_h<R - 1, NA>(a, xm);
                    if (!(sx4 >> 16)) xm ^= 16u;
                }
                continue;
            }
            if ((tid & cthr) != cthr) continue;   // a control on a thread bit is 0, or op inactive here
            double2 f = cur[1].d;
            const uint32_t crm = (hx >> 8) & 0x1Fu, p0 = (hx >> 16) & 0xFFu;
            if (opc >= OPC_XS_0 && opc < OPC_XS_0 + 5) {
                switch (opc - OPC_XS_0) {
                    case 0: reg_xswap<0, NA>(a, crm, xm); break;
                    case 1: reg_xswap<1, NA>(a, crm, xm); break;
                    case 2: reg_xswap<2, NA>(a, crm, xm); break;
                    case 3: reg_xswap<3, NA>(a, crm, xm); break;
                    default: if (R == 5) reg_xswap<R - 1, NA>(a, crm, xm); break;
                }
                continue;
            }
            if (opc == OPC_DS) {
                const uint32_t want = crm & ~xm;
#pragma unroll
                for (int p = 0; p < NA; ++p)
                    if ((p & crm) == want) cmul_ip(a[p], f);
                continue;
            }
            if (FULL) {
                // rare shapes live in the FULL variant only: the lean kernel stays small for the I-cache
                if (opc == OPC_DST) {
                    f = cur[1 + ((tid >> p0) & 1u)].d;
                    const uint32_t want = crm & ~xm;
#pragma unroll
                    for (int p = 0; p < NA; ++p)
                        if ((p & crm) == want) cmul_ip(a[p], f);
                    continue;
                }
                if (opc >= OPC_D1_0 && opc < OPC_D1_0 + 5) {
                    switch (opc - OPC_D1_0) {
                        case 0: reg_d1<0, NA>(a, f, cur[2].d, crm, xm); break;
                        case 1: reg_d1<1, NA>(a, f, cur[2].d, crm, xm); break;
                        case 2: reg_d1<2, NA>(a, f, cur[2].d, crm, xm); break;
                        case 3: reg_d1<3, NA>(a, f, cur[2].d, crm, xm); break;
                        default: if (R == 5) reg_d1<R - 1, NA>(a, f, cur[2].d, crm, xm); break;
                    }
                    continue;
                }
                const uint2 v = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(cur) + 8);
                const double2* mat = c.mats + v.x;
                if (opc == OPC_DIAG_G) {
                    const ColdOp* cd = &c.cold[(v.y >> 8) & 0xFFFFu];        // DIAG_G keeps its cold index in w[23:8]
                    const uint32_t kk = cd->k;
                    uint32_t idx = v.y & 0xFFu;
                    uint32_t rs[R];          // index weight of the target on register slot s (0: none)
#pragma unroll
                    for (int s = 0; s < R; ++s) rs[s] = 0;
                    for (uint32_t t = 0; t < kk; ++t) {
                        const uint32_t loc = cd->loc[t];
                        const uint32_t cls = loc & LOC_CLASS;
                        const uint32_t sh = 1u << (kk - 1 - t);
                        if (cls == LOC_THR) idx |= ((tid >> (loc & LOC_INDEX)) & 1u) ? sh : 0u;
                        else if (cls == LOC_REG) {
                            const uint32_t slot = loc & LOC_INDEX;       // each slot appears at most once
#pragma unroll
                            for (int s = 0; s < R; ++s)
                                if (slot == (uint32_t)s) rs[s] = sh;
                        }
                    }
                    const uint32_t want = crm & ~xm;
#pragma unroll
                    for (int p = 0; p < NA; ++p) {
                        if ((p & crm) == want) {
                            const uint32_t j = (uint32_t)p ^ xm;         // logical slot
                            uint32_t ij = idx;
#pragma unroll
                            for (int s = 0; s < R; ++s) ij |= ((j >> s) & 1u) ? rs[s] : 0u;
                            a[p] = cmul(QVM_LDG(&mat[ij]), a[p]);
                        }
                    }
                } else if (opc >= OPC_M1_0 && opc < OPC_M1_0 + 5) {
                    switch (opc - OPC_M1_0) {
                        case 0: reg_dense1<0, NA>(a, mat, crm, xm); break;
                        case 1: reg_dense1<1, NA>(a, mat, crm, xm); break;
                        case 2: reg_dense1<2, NA>(a, mat, crm, xm); break;
                        case 3: reg_dense1<3, NA>(a, mat, crm, xm); break;
                        default: if (R == 5) reg_dense1<R - 1, NA>(a, mat, crm, xm); break;
                    }
                } else if (opc == OPC_M2_0) reg_dense2<0, NA>(a, mat, crm, xm);
                else if (opc == OPC_M2_0 + 1) reg_dense2<1, NA>(a, mat, crm, xm);
            }
        }
    }

    if (last) {
        // register p holds logical slot p ^ xm: start from the address of logical slot xm and walk with
        // signed strides (a set xm bit turns +stride into -stride); sums again, no per-element masking
        const uint64_t gT = (uint64_t)(T & low_mask) | c.hi_off[T >> c.low];
        char* p0 = reinterpret_cast<char*>(c.gstore + gT);
        int64_t b[R];
#pragma unroll
        for (int s = 0; s < R; ++s) {
            b[s] = (int64_t)(16ull << rd->reg_phys[s]);
            if ((xm >> s) & 1u) {
                p0 += b[s];
                b[s] = -b[s];
            }
        }
        char* q[NA];
        add_tree<R>(p0, b, q);
#pragma unroll
        for (int j = 0; j < NA; ++j) *reinterpret_cast<double2*>(q[j]) = a[j];
    } else {
        // same mapping as this round's load: every thread rewrites exactly the slots it read ...
        uint32_t sX = sT;
        if (r == c.perm_round) {
            // ... except in the relabelling round: the tile is written in the next round's coordinates
            // (bit permutation sigma; swz and sigma are linear over XOR, so strides stay constants)
#if defined(__CUDA_ARCH__)
            if (!first) __syncthreads();     // every thread has read its slots before any slot is overwritten
#endif
            sX = 0;
            for (uint32_t b = 0, t = T; t; ++b, t >>= 1)
                if (t & 1u) sX ^= c.perm_swz[b];
            const uint4 w3 = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(rd) + 48);
            st[0] = w3.x >> 16;
            st[1] = w3.y & 0xFFFFu;
            st[2] = w3.y >> 16;
            st[3] = w3.z & 0xFFFFu;
            if (R == 5) st[R - 1] = w3.z >> 16;
        }
#pragma unroll
        for (int s = 0; s < R; ++s)
            if ((xm >> s) & 1u) sX ^= st[s];
#pragma unroll
        for (int j = 0; j < NA; ++j) c.tile_w[sX ^ xor_combo<R>(st, j)] = a[j];
    }
}

// outer bits of tile `tile_index`: deposit it into the runs of non-tile positions
QVM_HD uint64_t rt_tile_base(const RegTileParams& P, uint64_t tile_index) {
    uint64_t base = 0, t = tile_index;
    for (int r = 0; r < P.n_runs; ++r) {
        const uint32_t len = (uint32_t)(P.run_len >> (8 * r)) & 0xFFu;
        base |= (t & ((1ull << len) - 1ull)) << ((uint32_t)(P.run_start >> (8 * r)) & 0xFFu);
        t >>= len;
    }
    return base ^ P.push.rot;
}

// store base of the tile whose outer bits are `base` (see PushSpec)
QVM_HD double2* rt_push_base(const PushSpec& ps, double2* state, uint64_t base) {
    if (ps.m == 0) return state + base;
    uint32_t L = 0;
    for (int i = 0; i < ps.m; ++i) L |= (uint32_t)((base >> ps.pos[i]) & 1ull) << i;
    return ps.dst[L] + ((base & ~ps.mask) | ps.mine_bits);
}

#if defined(__CUDACC__)
template <int MAX_THREADS, int MIN_BLOCKS, bool FULL, int R>
__global__ void __launch_bounds__(MAX_THREADS, MIN_BLOCKS)
regtile_kernel(double2* __restrict__ state, const __grid_constant__ RegTileParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const size_t tile_bytes = P.n_rounds > 1 ? ((size_t)16 << P.k) : 0;
    OpWord* ops = reinterpret_cast<OpWord*>(smem_raw + tile_bytes);
    RoundDesc* rounds = reinterpret_cast<RoundDesc*>(ops + P.n_words + 1);
    uint64_t* hi_off = reinterpret_cast<uint64_t*>(rounds + P.n_rounds);

    const uint32_t tid = threadIdx.x;
    const uint32_t nthreads = blockDim.x;
    const uint64_t base = rt_tile_base(P, blockIdx.x);
    const uint64_t full = base | P.shard_bits;

    for (uint32_t j = tid; j < (uint32_t)P.n_ops; j += nthreads) rt_resolve_op(P, j, full, ops);
    if (tid == 0) ops[P.n_words].u = make_uint4(0, 0, 0, 0);          // the word the last prefetch reads
    if (P.n_patches) {
        __syncthreads();
        for (uint32_t j = tid; j < (uint32_t)P.n_patches; j += nthreads) rt_patch_op(P, P.n_ops + j, full, ops);
    }
    {
        const uint32_t rwords = (uint32_t)(P.n_rounds * (sizeof(RoundDesc) / 16));
        const uint4* rsrc = reinterpret_cast<const uint4*>(P.rounds);
        uint4* rdst = reinterpret_cast<uint4*>(rounds);
        for (uint32_t w = tid; w < rwords; w += nthreads) rdst[w] = __ldg(&rsrc[w]);
        const uint32_t n_hi = 1u << (P.k - P.low);
        for (uint32_t w = tid; w < n_hi; w += nthreads) hi_off[w] = __ldg(&P.hi_off[w]);
    }
    __syncthreads();

    RtCta c;
    c.tile = c.tile_w = reinterpret_cast<double2*>(smem_raw);
    c.perm_round = P.perm_round;
    c.perm_swz = P.perm_swz;
    c.ops = ops;
    c.rounds = rounds;
    c.hi_off = hi_off;
    c.gbase = state + base;
    c.gstore = rt_push_base(P.push, state, base);
    c.mats = P.mats;
    c.cold = P.cold;
    c.low = P.low;
    c.n_rounds = P.n_rounds;
    double2 a[1 << R];
    rt_first_load<R>(P, state + base, tid, a, base == 0);
    for (int r = 0; r < P.n_rounds; ++r) {
        rt_thread_round<FULL, R>(c, r, tid, a);
        if (r + 1 < P.n_rounds) __syncthreads();
    }
}
#endif  // __CUDACC__

}  // namespace qvm
