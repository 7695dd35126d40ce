// Device kernels of libqvm_b200 (sm_100a).  All arithmetic is complex128 held as double2.
//
// The workhorse is tile_kernel: one CTA stages the 2^k amplitudes addressed by an arbitrary set of
// k index bits ("tile bits", always containing the lowest `low` bits so global accesses are
// 16*2^low-byte contiguous runs) in shared memory, applies a whole run of gates to them and writes
// them back -- ONE read + ONE write of HBM for the run.  It implements the closed form of the
// reference's lifted operator (unitary_generator.py:266-323):
//     psi'[x] = sum_c G[r(x), c] * psi[x with bits a_j := c_(k-1-j)],   r(x)_(k-1-j) = x_(a_j).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "qvm_regtile.cuh"      // cmul / cfma / insert_zero helpers live there (shared with the CPU emulator)

namespace cg = cooperative_groups;

namespace qvm {

constexpr int kTileThreads = 256;
constexpr int kMaxFix = 13;

enum : uint8_t { DK_DENSE = 1, DK_DIAG = 2 };

// Per-op descriptor in tile-local coordinates (built on the host by build_dev_ops()).
struct __align__(16) DevOp {
    uint8_t kind;
    uint8_t k;                 // number of targets
    uint8_t nfix;              // dense: number of fixed tile positions (targets + tile controls)
    uint8_t pad0;
    uint8_t tgt[6];            // dense: tile-local target positions, MSB first
                               // diag : bit7 set -> physical position (outer / global), else tile-local
    uint8_t fix[kMaxFix];      // dense: fixed tile positions, ascending
    uint8_t pad1;
    uint32_t ctrl_tile;        // tile-local control mask
    uint32_t mat_off;          // offset into mats[], in double2
    uint64_t ctrl_outer;       // physical control mask outside the tile (incl. global positions)
};
static_assert(sizeof(DevOp) == 48, "DevOp layout");

struct TileParams {
    int n_local;
    int k;                     // tile bits
    int low;                   // tile_pos[0..low) == 0..low-1
    int n_ops;
    uint64_t tile_mask;        // physical mask of the tile bits
    uint64_t shard_bits;       // shard_index << n_local
    uint8_t tile_pos[16];      // physical position of tile bit j (ascending)
    const DevOp* ops;
    const double2* mats;
};

__device__ __forceinline__ uint32_t expand_fixed(uint32_t p, const uint8_t* fix, int nfix) {
    for (int j = 0; j < nfix; ++j) p = insert_zero32(p, fix[j]);
    return p;
}

// ---------------------------------------------------------------------------------------------
// Gate application on a tile resident in shared memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_dense1(double2* tile, const DevOp& op, const double2* mats, int k_tile,
                                            int tid, int nthreads) {
    const double2* m = mats + op.mat_off;
    const double2 m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
    const uint32_t tbit = 1u << op.tgt[0];
    const uint32_t cnt = 1u << (k_tile - op.nfix);
    for (uint32_t p = tid; p < cnt; p += nthreads) {
        const uint32_t i0 = expand_fixed(p, op.fix, op.nfix) | op.ctrl_tile;
        const uint32_t i1 = i0 | tbit;
        const double2 a0 = tile[i0], a1 = tile[i1];
        double2 b0 = cmul(m00, a0), b1 = cmul(m10, a0);
        b0 = cfma(m01, a1, b0);
        b1 = cfma(m11, a1, b1);
        tile[i0] = b0;
        tile[i1] = b1;
    }
}

__device__ __forceinline__ void tile_dense2(double2* tile, const DevOp& op, const double2* mats, int k_tile,
                                            int tid, int nthreads) {
    const double2* m = mats + op.mat_off;
    const uint32_t bhi = 1u << op.tgt[0], blo = 1u << op.tgt[1];
    const uint32_t cnt = 1u << (k_tile - op.nfix);
    for (uint32_t p = tid; p < cnt; p += nthreads) {
        const uint32_t i0 = expand_fixed(p, op.fix, op.nfix) | op.ctrl_tile;
        const uint32_t idx[4] = {i0, i0 | blo, i0 | bhi, i0 | bhi | blo};
        double2 a[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) a[c] = tile[idx[c]];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            double2 acc = cmul(m[r * 4 + 0], a[0]);
#pragma unroll
            for (int c = 1; c < 4; ++c) acc = cfma(m[r * 4 + c], a[c], acc);
            tile[idx[r]] = acc;
        }
    }
}

// 3 <= k <= 6: a group of min(32, 2^k) lanes owns one 2^k-amplitude block; lane r computes row r
// (and row r+32 when k == 6).  Inputs are read (broadcast) from shared memory, the matrix from
// global memory through L1.  Slow path: DefGates wider than two qubits only.
__device__ __forceinline__ void tile_densek(double2* tile, const DevOp& op, const double2* mats, int k_tile,
                                            int tid, int nthreads) {
    const int k = op.k;
    const int dim = 1 << k;
    const int lanes = dim < 32 ? dim : 32;
    const int per_warp = 32 / lanes;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int sub = lane / lanes, r0 = lane % lanes;
    const double2* m = mats + op.mat_off;
    const uint32_t cnt = 1u << (k_tile - op.nfix);
    uint32_t scat0 = 0, scat1 = 0;       // tile offsets of rows r0 and r0+32
    for (int j = 0; j < k; ++j) {
        if ((r0 >> (k - 1 - j)) & 1) scat0 |= 1u << op.tgt[j];
        if (((r0 + 32) >> (k - 1 - j)) & 1) scat1 |= 1u << op.tgt[j];
    }
    const uint32_t stride = (uint32_t)(nwarps * per_warp);
    const uint32_t rounds = (cnt + stride - 1) / stride;
    for (uint32_t it = 0; it < rounds; ++it) {
        const uint32_t g = it * stride + warp * per_warp + sub;
        const bool active = g < cnt;
        const uint32_t ib = active ? (expand_fixed(g, op.fix, op.nfix) | op.ctrl_tile) : 0u;
        double2 acc0 = make_double2(0.0, 0.0), acc1 = make_double2(0.0, 0.0);
        if (active) {
            for (int c = 0; c < dim; ++c) {
                uint32_t sc = 0;
                for (int j = 0; j < k; ++j)
                    if ((c >> (k - 1 - j)) & 1) sc |= 1u << op.tgt[j];
                const double2 a = tile[ib | sc];
                acc0 = cfma(__ldg(&m[r0 * dim + c]), a, acc0);
                if (dim == 64) acc1 = cfma(__ldg(&m[(r0 + 32) * dim + c]), a, acc1);
            }
        }
        __syncwarp();
        if (active) {
            tile[ib | scat0] = acc0;
            if (dim == 64) tile[ib | scat1] = acc1;
        }
        __syncwarp();
    }
}

// A run of consecutive DIAG ops [o, e): one read-modify-write of each amplitude for the whole run.
__device__ __forceinline__ void tile_diag_run(double2* tile, const DevOp* ops, int o, int e, const uint8_t* op_active,
                                              const uint8_t* op_outer_idx, const double2* mats, int k_tile, int tid,
                                              int nthreads) {
    const uint32_t tile_n = 1u << k_tile;
    for (uint32_t i = tid; i < tile_n; i += nthreads) {
        double2 v = tile[i];
        bool changed = false;
        for (int j = o; j < e; ++j) {
            if (!op_active[j]) continue;
            const DevOp& op = ops[j];
            if ((i & op.ctrl_tile) != op.ctrl_tile) continue;
            uint32_t idx = op_outer_idx[j];
            for (int t = 0; t < op.k; ++t) {
                const uint8_t loc = op.tgt[t];
                if (!(loc & 0x80)) idx |= ((i >> loc) & 1u) << (op.k - 1 - t);
            }
            v = cmul(__ldg(&mats[op.mat_off + idx]), v);
            changed = true;
        }
        if (changed) tile[i] = v;
    }
}

// Applies ops[0..n_ops) to the tile.  Caller guarantees the tile is loaded and synchronised; on
// return all writes are synchronised.
__device__ __forceinline__ void tile_apply_ops(double2* tile, const DevOp* ops, int n_ops, const uint8_t* op_active,
                                               const uint8_t* op_outer_idx, const double2* mats, int k_tile, int tid,
                                               int nthreads) {
    int o = 0;
    while (o < n_ops) {
        const DevOp& op = ops[o];
        if (op.kind == DK_DIAG) {
            int e = o + 1;
            while (e < n_ops && ops[e].kind == DK_DIAG) ++e;
            tile_diag_run(tile, ops, o, e, op_active, op_outer_idx, mats, k_tile, tid, nthreads);
            o = e;
            __syncthreads();
            continue;
        }
        if (op_active[o]) {                     // CTA-uniform
            if (op.k == 1) tile_dense1(tile, op, mats, k_tile, tid, nthreads);
            else if (op.k == 2) tile_dense2(tile, op, mats, k_tile, tid, nthreads);
            else tile_densek(tile, op, mats, k_tile, tid, nthreads);
            __syncthreads();
        }
        ++o;
    }
}

// Shared-memory carve-up: [tile: 2^k double2][ops: n_ops DevOp][active: n_ops u8][outer_idx: n_ops u8][hi_off]
__host__ __device__ inline size_t tile_smem_bytes(int k, int low, int n_ops) {
    size_t b = (size_t)16 << k;
    b += (size_t)n_ops * sizeof(DevOp);
    b += (((size_t)n_ops * 2 + 15) / 16) * 16;
    b += (size_t)8 << (k - low);
    return b;
}

__global__ void __launch_bounds__(kTileThreads) tile_kernel(double2* __restrict__ state, const TileParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* tile = reinterpret_cast<double2*>(smem_raw);
    DevOp* ops = reinterpret_cast<DevOp*>(smem_raw + ((size_t)16 << P.k));
    uint8_t* op_active = reinterpret_cast<uint8_t*>(ops + P.n_ops);
    uint8_t* op_outer_idx = op_active + P.n_ops;
    uint64_t* hi_off = reinterpret_cast<uint64_t*>(smem_raw + ((size_t)16 << P.k) + (size_t)P.n_ops * sizeof(DevOp) +
                                                   (((size_t)P.n_ops * 2 + 15) / 16) * 16);
    const int tid = threadIdx.x;
    const int k = P.k, low = P.low;

    // outer bits of this tile: deposit blockIdx into the non-tile positions
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

    const uint32_t n_hi = 1u << (k - low);
    for (uint32_t j = tid; j < n_hi; j += kTileThreads) {
        uint64_t off = 0;
        for (int b = 0; b < k - low; ++b)
            if ((j >> b) & 1u) off |= 1ull << P.tile_pos[low + b];
        hi_off[j] = off;
    }
    // stage op descriptors + CTA-uniform per-op facts
    {
        const uint32_t words = (uint32_t)(P.n_ops * (sizeof(DevOp) / 16));
        const uint4* src = reinterpret_cast<const uint4*>(P.ops);
        uint4* dst = reinterpret_cast<uint4*>(ops);
        for (uint32_t w = tid; w < words; w += kTileThreads) dst[w] = __ldg(&src[w]);
    }
    for (int j = tid; j < P.n_ops; j += kTileThreads) {
        const DevOp* g = &P.ops[j];
        const uint64_t co = g->ctrl_outer;
        op_active[j] = ((full & co) == co) ? 1 : 0;
        uint32_t oi = 0;
        if (g->kind == DK_DIAG)
            for (int t = 0; t < g->k; ++t) {
                const uint8_t loc = g->tgt[t];
                if (loc & 0x80) oi |= (uint32_t)((full >> (loc & 0x7f)) & 1ull) << (g->k - 1 - t);
            }
        op_outer_idx[j] = (uint8_t)oi;
    }
    __syncthreads();

    const uint32_t tile_n = 1u << k;
    const uint32_t low_mask = (1u << low) - 1u;
    double2* gbase = state + base;
    for (uint32_t i = tid; i < tile_n; i += kTileThreads) tile[i] = gbase[(i & low_mask) + hi_off[i >> low]];
    __syncthreads();

    tile_apply_ops(tile, ops, P.n_ops, op_active, op_outer_idx, P.mats, k, tid, kTileThreads);

    for (uint32_t i = tid; i < tile_n; i += kTileThreads) gbase[(i & low_mask) + hi_off[i >> low]] = tile[i];
}

// ---------------------------------------------------------------------------------------------
// Fused all-qubit dephasing on a vectorised rho: rho[i,j] *= f^popcount((i^j) & mask).
// ---------------------------------------------------------------------------------------------
__global__ void dephase_kernel(double2* __restrict__ state, uint64_t n_elems, int n_rho, uint64_t qubit_mask,
                               uint64_t shard_bits, const double* __restrict__ pow_table) {
    const uint64_t row_mask = (1ull << n_rho) - 1ull;
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < n_elems;
         x += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g = x | shard_bits;
        const uint64_t col = g & row_mask, row = (g >> n_rho) & row_mask;
        const int pc = __popcll((row ^ col) & qubit_mask);
        if (pc) {
            const double f = pow_table[pc];
            double2 v = state[x];
            v.x *= f;
            v.y *= f;
            state[x] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Reductions, MEASURE, collapse
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum (fixed tree); result valid in every thread.
__device__ __forceinline__ double block_sum(double v, double* scratch /* >= 33 doubles */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = lane < nw ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    const double r = scratch[32];
    __syncthreads();
    return r;
}

// mode 0: sum over amplitudes whose bit `qubit` is 0 of |amp|^2 (qubit < 0: all amplitudes)
// mode 1: vectorised rho with n_rho qubits: sum over local rows i with bit `qubit` == 0 of Re rho[i,i]
__device__ __forceinline__ double partial_prob(const double2* __restrict__ state, uint64_t n_elems, int n_local,
                                               int qubit, int mode, int n_rho, uint64_t shard_bits, uint64_t gtid,
                                               uint64_t gthreads) {
    double acc = 0.0;
    if (mode == 0) {
        if (qubit < 0 || qubit >= n_local) {
            // whole shard (caller handles a global qubit by looking at its shard bit)
            for (uint64_t x = gtid; x < n_elems; x += gthreads) {
                const double2 a = state[x];
                acc += a.x * a.x + a.y * a.y;
            }
        } else if (qubit < 3) {
            for (uint64_t x = gtid; x < n_elems; x += gthreads) {
                const double2 a = state[x];
                if (!((x >> qubit) & 1ull)) acc += a.x * a.x + a.y * a.y;
            }
        } else {
            const uint64_t half = n_elems >> 1;
            for (uint64_t h = gtid; h < half; h += gthreads) {
                const double2 a = state[insert_zero64(h, qubit)];
                acc += a.x * a.x + a.y * a.y;
            }
        }
    } else {
        // local rows: the local index is (row_local << n_rho | col) when n_local >= n_rho
        const uint64_t n_rows_local = n_elems >> n_rho;
        const uint64_t row_base = shard_bits >> n_rho;      // global row offset of this shard
        for (uint64_t r = gtid; r < n_rows_local; r += gthreads) {
            const uint64_t row = row_base | r;
            if (qubit >= 0 && ((row >> qubit) & 1ull)) continue;
            acc += state[(r << n_rho) | row].x;
        }
    }
    return acc;
}

__global__ void prob_partial_kernel(const double2* __restrict__ state, uint64_t n_elems, int n_local, int qubit,
                                    int mode, int n_rho, uint64_t shard_bits, double* __restrict__ partials) {
    __shared__ double scratch[33];
    const uint64_t gtid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t gthreads = (uint64_t)gridDim.x * blockDim.x;
    double acc = partial_prob(state, n_elems, n_local, qubit, mode, n_rho, shard_bits, gtid, gthreads);
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

__global__ void final_sum_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
    __shared__ double scratch[33];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) out[0] = acc;
}

__device__ __forceinline__ void collapse_range(double2* __restrict__ state, uint64_t n_elems, int qubit, int outcome,
                                               double scale, uint64_t gtid, uint64_t gthreads) {
    const double2 zero = make_double2(0.0, 0.0);
    for (uint64_t x = gtid; x < n_elems; x += gthreads) {
        if ((int)((x >> qubit) & 1ull) == outcome) {
            double2 a = state[x];
            a.x *= scale;
            a.y *= scale;
            state[x] = a;
        } else {
            state[x] = zero;
        }
    }
}

// scale < 0 means "zero the whole shard" (global qubit measured to the other value).
__global__ void collapse_kernel(double2* __restrict__ state, uint64_t n_elems, int n_local, int qubit, int outcome,
                                double scale) {
    const uint64_t gtid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t gthreads = (uint64_t)gridDim.x * blockDim.x;
    if (qubit >= n_local || qubit < 0) {
        for (uint64_t x = gtid; x < n_elems; x += gthreads) {
            double2 a = state[x];
            if (scale < 0.0) a = make_double2(0.0, 0.0);
            else { a.x *= scale; a.y *= scale; }
            state[x] = a;
        }
        return;
    }
    collapse_range(state, n_elems, qubit, outcome, scale, gtid, gthreads);
}

// MEASURE as ONE cooperative kernel: reduce -> grid.sync -> decide -> collapse + renormalise.
// result[0] = outcome, result[1] = p0 (as doubles).
__global__ void measure_kernel(double2* __restrict__ state, uint64_t n_elems, int n_local, int qubit, double r,
                               double* __restrict__ partials, double* __restrict__ result) {
    __shared__ double scratch[33];
    cg::grid_group grid = cg::this_grid();
    const uint64_t gtid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t gthreads = (uint64_t)gridDim.x * blockDim.x;
    double acc = partial_prob(state, n_elems, n_local, qubit, 0, 0, 0, gtid, gthreads);
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
    grid.sync();
    double tot = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) tot += partials[i];
    const double p0 = block_sum(tot, scratch);       // same order in every block -> identical value
    const int outcome = (r < p0) ? 0 : 1;
    const double scale = 1.0 / sqrt(outcome ? (1.0 - p0) : p0);
    collapse_range(state, n_elems, qubit, outcome, scale, gtid, gthreads);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        result[0] = (double)outcome;
        result[1] = p0;
    }
}

// Runs of MEASURE instructions (qvm_wavefunction.py:147-156 executed back to back): one pass builds the
// joint distribution of up to 5 "high" measured positions (>= 5) x all 32 values of the 5 lowest index bits,
// the host replays the reference's sequential decisions on it (one uniform per MEASURE, p0 conditional on
// the outcomes so far), and the NEXT pass applies the combined collapse + renormalisation while it builds
// the distribution for the following chunk.  A warp reads 32 consecutive amplitudes, so the high-bin is the
// same for all its lanes and every lane owns one column of its warp's histogram: no atomics, fixed
// summation order (deterministic to the last bit).
constexpr int kHistThreads = 128;          // 4 warps x 32 bins x 32 lanes x 8 B = 32 KB of shared memory
struct HistBits {
    int n_high;
    uint8_t pos[5];
};
__global__ void __launch_bounds__(kHistThreads)
measure_hist_kernel(double2* __restrict__ state, uint64_t n_elems, const HistBits hb, int do_collapse,
                    uint64_t keep_mask, uint64_t keep_value, double scale, double* __restrict__ partials) {
    __shared__ double acc[kHistThreads / 32][32][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_bins = 1 << hb.n_high;
    for (int b = 0; b < n_bins; ++b) acc[warp][b][lane] = 0.0;
    const uint64_t n_chunks = n_elems >> 5;
    const uint64_t warps_total = (uint64_t)gridDim.x * (kHistThreads / 32);
    for (uint64_t c = (uint64_t)blockIdx.x * (kHistThreads / 32) + warp; c < n_chunks; c += warps_total) {
        const uint64_t x = (c << 5) | (uint64_t)lane;
        double2 a = state[x];
        if (do_collapse) {
            if ((x & keep_mask) == keep_value) {
                a.x *= scale;
                a.y *= scale;
            } else {
                a = make_double2(0.0, 0.0);
            }
            state[x] = a;
        }
        int bin = 0;
        for (int i = 0; i < hb.n_high; ++i) bin |= (int)((x >> hb.pos[i]) & 1ull) << i;
        acc[warp][bin][lane] += a.x * a.x + a.y * a.y;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n_bins * 32; j += kHistThreads) {
        double t = 0.0;
        for (int w = 0; w < kHistThreads / 32; ++w) t += acc[w][j >> 5][j & 31];
        partials[(size_t)blockIdx.x * (n_bins * 32) + j] = t;
    }
}
// The same histogram over the amplitudes CONSISTENT with earlier outcomes only -- (x & keep_mask) == keep_value -- and
// without writing anything: after the first chunk of a MEASURE run has fixed ten or so bits, the following chunks read
// a thousandth of the state instead of rewriting all of it, and the run ends with ONE collapse pass.  Thread g walks
// the compact indices c = g, g + T, ... of the free positions (T a multiple of 32, so the bits of c that land on free
// positions among the 5 lowest stay the same for a thread: every thread still owns one lane value, several threads
// of a warp may share it) and deposits c into the free runs.  Fixed summation order, no atomics.
struct FilterBits {
    int n_runs;
    uint8_t start[20], len[20];        // runs of free positions, ascending
};
__global__ void __launch_bounds__(kHistThreads)
measure_hist_filter_kernel(const double2* __restrict__ state, int n_free, const HistBits hb, const FilterBits fb,
                           uint64_t keep_value, double* __restrict__ partials) {
    __shared__ double acc[kHistThreads / 32][32][32];
    __shared__ int lane_of[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_bins = 1 << hb.n_high;
    for (int b = 0; b < n_bins; ++b) acc[warp][b][lane] = 0.0;
    auto deposit = [&](uint64_t c) {
        uint64_t x = keep_value;
        for (int r = 0; r < fb.n_runs; ++r) {
            x |= (c & ((1ull << fb.len[r]) - 1ull)) << fb.start[r];
            c >>= fb.len[r];
        }
        return x;
    };
    if (threadIdx.x < 32) lane_of[threadIdx.x] = (int)(deposit((uint64_t)threadIdx.x) & 31ull);
    const uint64_t n = 1ull << n_free;
    const uint64_t threads_total = (uint64_t)gridDim.x * kHistThreads;
    for (uint64_t c = (uint64_t)blockIdx.x * kHistThreads + threadIdx.x; c < n; c += threads_total) {
        const uint64_t x = deposit(c);
        const double2 a = state[x];
        int bin = 0;
        for (int i = 0; i < hb.n_high; ++i) bin |= (int)((x >> hb.pos[i]) & 1ull) << i;
        acc[warp][bin][lane] += a.x * a.x + a.y * a.y;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n_bins * 32; j += kHistThreads) {
        double t = 0.0;
        for (int w = 0; w < kHistThreads / 32; ++w)
            for (int l = 0; l < 32; ++l)
                if (lane_of[l] == (j & 31)) t += acc[w][j >> 5][l];
        partials[(size_t)blockIdx.x * (n_bins * 32) + j] = t;
    }
}
// The run measured EVERY local position: one amplitude survives.  keep[0] = state[x0] * scale, then the shard is
// cleared (memset) and the survivor written back -- a write-only pass instead of read + write.
__global__ void survivor_fetch_kernel(const double2* __restrict__ state, uint64_t x0, double scale, double2* __restrict__ keep) {
    const double2 a = state[x0];
    keep[0] = make_double2(a.x * scale, a.y * scale);
}
__global__ void survivor_store_kernel(double2* __restrict__ state, uint64_t x0, const double2* __restrict__ keep) {
    state[x0] = keep[0];
}
// hist[j] = sum over blocks of partials[block][j], blocks in ascending order
__global__ void measure_hist_final_kernel(const double* __restrict__ partials, int n_blocks, int n_entries,
                                          double* __restrict__ hist) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_entries) return;
    double t = 0.0;
    for (int b = 0; b < n_blocks; ++b) t += partials[(size_t)b * n_entries + j];
    hist[j] = t;
}

// Batched trials: the shard holds 2^(n_local - n) independent n-qubit states ("segments", one per
// trial of run()); each block owns one segment.  Same arithmetic as measure_kernel, per segment, with
// that trial's own uniform: p0 -> outcome 0 iff r < p0 -> zero the other half, scale by 1/sqrt(p).
__global__ void batched_measure_kernel(double2* __restrict__ state, int n, int qubit, const double* __restrict__ r,
                                       int32_t* __restrict__ outcomes, double* __restrict__ p0_out) {
    __shared__ double scratch[33];
    const uint64_t seg = blockIdx.x;
    double2* s = state + (seg << n);
    const uint64_t n_elems = 1ull << n;
    double acc = partial_prob(s, n_elems, n, qubit, 0, 0, 0, threadIdx.x, blockDim.x);
    const double p0 = block_sum(acc, scratch);
    const int outcome = (r[seg] < p0) ? 0 : 1;
    const double scale = 1.0 / sqrt(outcome ? (1.0 - p0) : p0);
    collapse_range(s, n_elems, qubit, outcome, scale, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) {
        outcomes[seg] = outcome;
        p0_out[seg] = p0;
    }
}

// Batched trials under classical control: when the trials of a batch take different branches they are
// regrouped -- dst segment j := src segment index[j] (j < count) -- so that every batch again walks ONE
// instruction sequence.  One block per destination segment, 128-bit copies.
__global__ void gather_segments_kernel(double2* __restrict__ dst, const double2* __restrict__ src, int n,
                                       const uint32_t* __restrict__ index) {
    const uint64_t seg = blockIdx.x;
    const double2* s = src + ((uint64_t)index[seg] << n);
    double2* d = dst + (seg << n);
    const uint64_t n_elems = 1ull << n;
    for (uint64_t i = threadIdx.x; i < n_elems; i += blockDim.x) d[i] = s[i];
}

// every segment := |0...0>
__global__ void batched_reset_kernel(double2* __restrict__ state, uint64_t n_elems, int n) {
    const uint64_t mask = (1ull << n) - 1ull;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_elems; i += (uint64_t)gridDim.x * blockDim.x)
        state[i] = make_double2((i & mask) ? 0.0 : 1.0, 0.0);
}

// ---------------------------------------------------------------------------------------------
// Sampling: two-level CDF + per-shot binary search
// ---------------------------------------------------------------------------------------------
constexpr int kSampleBlockBits = 11;      // 2048 probabilities per CDF block

// mode 0: |amp_i|^2; mode 1: Re rho[i,i] of a vectorised rho (row-major, identity layout); mode 2: the buffer IS
// a plain array of probabilities (doubles): a diagonal gathered through a qubit layout / across shards
__device__ __forceinline__ double prob_at(const double2* __restrict__ state, uint64_t i, int mode, int n_rho,
                                          uint64_t row_base) {
    if (mode == 0) {
        const double2 a = state[i];
        return a.x * a.x + a.y * a.y;
    }
    if (mode == 2) return reinterpret_cast<const double*>(state)[i];
    return state[(i << n_rho) | (row_base | i)].x;
}

// diag(rho) through a qubit layout (density matrices under relabelling / sharding): entry i of the diagonal
// lives at the physical index whose bits pos_col[q] and pos_row[q] both equal bit q of i.  A shard writes the
// entries it owns and 0.0 elsewhere, so the shards' outputs add up to the full diagonal exactly.
struct DiagLayout {
    uint8_t pos_col[32];
    uint8_t pos_row[32];
};
// rho[i, i ^ col_xor] for every row i, through the layout (complex; owned entries, 0 elsewhere): the entries a
// Pauli string's trace tr(rho P) touches -- P maps |i> to a phase times |i ^ xmask>.  col_xor = 0: the diagonal.
__global__ void density_offdiag_kernel(const double2* __restrict__ state, int n_rho, const DiagLayout lay, int n_local,
                                       uint64_t shard_index, uint64_t col_xor, double2* __restrict__ out) {
    const uint64_t n = 1ull << n_rho;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t col = i ^ col_xor;
        uint64_t phys = 0;
        for (int q = 0; q < n_rho; ++q) {
            if ((i >> q) & 1ull) phys |= 1ull << lay.pos_row[q];
            if ((col >> q) & 1ull) phys |= 1ull << lay.pos_col[q];
        }
        out[i] = (phys >> n_local) == shard_index ? state[phys & ((1ull << n_local) - 1ull)] : make_double2(0.0, 0.0);
    }
}

// Expectation of a Pauli string on a state vector (reference: QVM expectation() is NotImplementedError,
// qvm_unitary.py:99-112; its meaning is <psi| tensor_up(P) |psi>, unitary_generator.py:366-411).
// P|x> = i^ny (-1)^popcount(x & zmask) |x ^ xmask> (xmask: X and Y qubits, zmask: Z and Y qubits), hence
// <psi|P|psi> = i^ny * sum_x conj(psi[x ^ xmask]) (-1)^popcount(x & zmask) psi[x]: one pass, the partial sums of
// the real and imaginary parts go to partials[2 * block], [2 * block + 1].  xmask is local; zmask may reach the
// shard bits (`full`).
__global__ void pauli_partial_kernel(const double2* __restrict__ state, uint64_t n_elems, uint64_t xmask, uint64_t zmask,
                                     uint64_t shard_bits, double* __restrict__ partials) {
    __shared__ double scratch[33];
    double re = 0.0, im = 0.0;
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < n_elems; x += (uint64_t)gridDim.x * blockDim.x) {
        const double2 a = state[x];
        const double2 b = state[x ^ xmask];
        const double s = (__popcll((x | shard_bits) & zmask) & 1) ? -1.0 : 1.0;
        re += s * (b.x * a.x + b.y * a.y);          // conj(b) * a
        im += s * (b.x * a.y - b.y * a.x);
    }
    re = block_sum(re, scratch);
    im = block_sum(im, scratch);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = re;
        partials[2 * blockIdx.x + 1] = im;
    }
}
// sum_x conj(a[x]) b[x] of two shards of equal shape (general operator programs: <psi|O psi>)
__global__ void inner_partial_kernel(const double2* __restrict__ a, const double2* __restrict__ b, uint64_t n_elems,
                                     double* __restrict__ partials) {
    __shared__ double scratch[33];
    double re = 0.0, im = 0.0;
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < n_elems; x += (uint64_t)gridDim.x * blockDim.x) {
        const double2 u = a[x], v = b[x];
        re += u.x * v.x + u.y * v.y;
        im += u.x * v.y - u.y * v.x;
    }
    re = block_sum(re, scratch);
    im = block_sum(im, scratch);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = re;
        partials[2 * blockIdx.x + 1] = im;
    }
}
// out[0] = sum of partials[2 b], out[1] = sum of partials[2 b + 1] (blocks in ascending order)
__global__ void final_sum2_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
    __shared__ double scratch[33];
    double re = 0.0, im = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        re += partials[2 * i];
        im += partials[2 * i + 1];
    }
    re = block_sum(re, scratch);
    im = block_sum(im, scratch);
    if (threadIdx.x == 0) {
        out[0] = re;
        out[1] = im;
    }
}
__global__ void density_diag_kernel(const double2* __restrict__ state, int n_rho, const DiagLayout lay, int n_local,
                                    uint64_t shard_index, double* __restrict__ out) {
    const uint64_t n = 1ull << n_rho;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t phys = 0;
        for (int q = 0; q < n_rho; ++q)
            if ((i >> q) & 1ull) phys |= (1ull << lay.pos_col[q]) | (1ull << lay.pos_row[q]);
        out[i] = (phys >> n_local) == shard_index ? state[phys & ((1ull << n_local) - 1ull)].x : 0.0;
    }
}

// block_sums[b] = sum of the probabilities in block b.  One warp per block.
__global__ void sample_block_sums_kernel(const double2* __restrict__ state, uint64_t n_probs, int mode, int n_rho,
                                         uint64_t row_base, double* __restrict__ block_sums, uint64_t n_blocks) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t b = warp; b < n_blocks; b += nwarps) {
        const uint64_t lo = b << kSampleBlockBits;
        uint64_t hi = lo + (1ull << kSampleBlockBits);
        if (hi > n_probs) hi = n_probs;
        double acc = 0.0;
        for (uint64_t i = lo + lane; i < hi; i += 32) acc += prob_at(state, i, mode, n_rho, row_base);
        acc = warp_sum(acc);
        if (lane == 0) block_sums[b] = acc;
    }
}

// In-place inclusive scan of v[0..n) by ONE block (deterministic); out_total[0] = v[n-1] after scan.
__global__ void scan_inclusive_kernel(double* __restrict__ v, uint64_t n, double* __restrict__ out_total) {
    __shared__ double chunk_tot[1024];
    const int t = threadIdx.x, nt = blockDim.x;
    const uint64_t per = (n + nt - 1) / nt;
    const uint64_t lo = (uint64_t)t * per;
    uint64_t hi = lo + per;
    if (hi > n) hi = n;
    double acc = 0.0;
    for (uint64_t i = lo; i < hi; ++i) acc += v[i];
    chunk_tot[t] = acc;
    __syncthreads();
    if (t == 0) {
        double run = 0.0;
        for (int i = 0; i < nt; ++i) {
            const double c = chunk_tot[i];
            chunk_tot[i] = run;
            run += c;
        }
        out_total[0] = run;
    }
    __syncthreads();
    double run = chunk_tot[t];
    for (uint64_t i = lo; i < hi; ++i) {
        run += v[i];
        v[i] = run;
    }
}

// Three-phase scan for long arrays (33 qubits: 2^22 block sums): chunks of kScanChunk entries, one CTA
// each.  (1) chunk totals, (2) the single-block scan above over the (few) totals, (3) every chunk scans
// itself and adds its offset.  Each thread owns kScanPerThread consecutive entries and sums them in
// order, thread totals are added in thread order, chunk offsets in chunk order: a fixed association,
// so results are deterministic to the last bit, and all global accesses are full 128-byte lines.
constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 16;
constexpr int kScanChunk = kScanThreads * kScanPerThread;
__device__ __forceinline__ double scan_thread_load(const double* __restrict__ v, uint64_t n, uint64_t base,
                                                   double (&x)[kScanPerThread]) {
    double run = 0.0;
#pragma unroll
    for (int j = 0; j < kScanPerThread; ++j) {
        x[j] = base + j < n ? v[base + j] : 0.0;
        run += x[j];
        x[j] = run;                       // inclusive within the thread
    }
    return run;
}
__global__ void __launch_bounds__(kScanThreads) scan_chunk_totals_kernel(const double* __restrict__ v, uint64_t n,
                                                                         double* __restrict__ chunk_tot) {
    __shared__ double tot[kScanThreads];
    double x[kScanPerThread];
    tot[threadIdx.x] = scan_thread_load(v, n, (uint64_t)blockIdx.x * kScanChunk + threadIdx.x * kScanPerThread, x);
    __syncthreads();
    if (threadIdx.x == 0) {
        double run = 0.0;
        for (int i = 0; i < kScanThreads; ++i) run += tot[i];
        chunk_tot[blockIdx.x] = run;
    }
}
__global__ void __launch_bounds__(kScanThreads) scan_chunks_kernel(double* __restrict__ v, uint64_t n,
                                                                   const double* __restrict__ chunk_cdf) {
    __shared__ double tot[kScanThreads];
    double x[kScanPerThread];
    const uint64_t base = (uint64_t)blockIdx.x * kScanChunk + threadIdx.x * kScanPerThread;
    tot[threadIdx.x] = scan_thread_load(v, n, base, x);
    __syncthreads();
    double before = blockIdx.x ? chunk_cdf[blockIdx.x - 1] : 0.0;      // everything in earlier chunks ...
    for (int i = 0; i < (int)threadIdx.x; ++i) before += tot[i];        // ... and in earlier threads of this one
#pragma unroll
    for (int j = 0; j < kScanPerThread; ++j)
        if (base + j < n) v[base + j] = before + x[j];
}

// One warp per shot.  target = u * total; block = first b with cdf[b] > target; then walk the block.
__global__ void sample_kernel(const double2* __restrict__ state, uint64_t n_probs, int mode, int n_rho,
                              uint64_t row_base, const double* __restrict__ block_cdf, uint64_t n_blocks,
                              const double* __restrict__ total_ptr, const double* __restrict__ uniforms,
                              uint64_t n_shots, unsigned long long* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const double total = total_ptr[0];
    for (uint64_t s = warp; s < n_shots; s += nwarps) {
        const double target = uniforms[s] * total;
        uint64_t lo = 0, hi = n_blocks;             // first b in [0, n_blocks) with cdf[b] > target
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (block_cdf[mid] > target) hi = mid; else lo = mid + 1;
        }
        uint64_t b = lo < n_blocks ? lo : n_blocks - 1;
        double run = b ? block_cdf[b - 1] : 0.0;
        const uint64_t first = b << kSampleBlockBits;
        uint64_t last = first + (1ull << kSampleBlockBits);
        if (last > n_probs) last = n_probs;
        uint64_t found = last - 1;
        bool done = false;
        uint64_t last_nonzero = first;
        for (uint64_t i0 = first; i0 < last && !done; i0 += 32) {
            const uint64_t i = i0 + lane;
            const double p = i < last ? prob_at(state, i, mode, n_rho, row_base) : 0.0;
            double incl = p;                         // warp inclusive scan
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            const double c = run + incl;
            const unsigned hit = __ballot_sync(0xffffffffu, (i < last) && (c > target));
            const unsigned nz = __ballot_sync(0xffffffffu, (i < last) && (p > 0.0));
            if (nz) last_nonzero = i0 + (31 - __clz(nz));
            if (hit) {
                found = i0 + (__ffs(hit) - 1);
                done = true;
            }
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (!done) found = last_nonzero;             // rounding at the very top of the block
        if (lane == 0) out[s] = found;
    }
}

// ---------------------------------------------------------------------------------------------
// Small utilities
// ---------------------------------------------------------------------------------------------
__global__ void fill_zero_kernel(double2* __restrict__ state, uint64_t n_elems, int set_amp0) {
    const double2 zero = make_double2(0.0, 0.0);
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < n_elems;
         x += (uint64_t)gridDim.x * blockDim.x)
        state[x] = (x == 0 && set_amp0) ? make_double2(1.0, 0.0) : zero;
}

__global__ void gather_strided_kernel(const double2* __restrict__ state, uint64_t offset, uint64_t stride,
                                      uint64_t count, double2* __restrict__ dst) {
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < count;
         j += (uint64_t)gridDim.x * blockDim.x)
        dst[j] = state[offset + j * stride];
}

// Readback through a qubit layout: element j of the result is the amplitude of LOGICAL index
// offset + j * stride, which lives at the physical index whose bit pos[q] is logical bit q.
struct LayoutPerm {
    uint8_t pos[64];
    int n;
};
__global__ void gather_layout_kernel(const double2* __restrict__ state, uint64_t offset, uint64_t stride,
                                     uint64_t count, const LayoutPerm perm, double2* __restrict__ dst) {
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < count;
         j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t logical = offset + j * stride;
        uint64_t phys = 0;
        for (int q = 0; q < perm.n; ++q) phys |= ((logical >> q) & 1ull) << perm.pos[q];
        dst[j] = state[phys];
    }
}

// half-shard pack / unpack (global<->local qubit exchange): element j of the half whose bit
// `local_bit` == bit_value lives at insert(j, local_bit, bit_value).
__global__ void pack_half_kernel(const double2* __restrict__ state, int local_bit, int bit_value, uint64_t off,
                                 uint64_t count, double2* __restrict__ dst) {
    const uint64_t vbit = (uint64_t)bit_value << local_bit;
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < count;
         j += (uint64_t)gridDim.x * blockDim.x)
        dst[j] = state[insert_zero64(off + j, local_bit) | vbit];
}
__global__ void unpack_half_kernel(double2* __restrict__ state, int local_bit, int bit_value, uint64_t off,
                                   uint64_t count, const double2* __restrict__ src) {
    const uint64_t vbit = (uint64_t)bit_value << local_bit;
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < count;
         j += (uint64_t)gridDim.x * blockDim.x)
        state[insert_zero64(off + j, local_bit) | vbit] = src[j];
}

// Global<->local qubit swap as ONE kernel over peer memory (NVLink P2P).  Swapping m global bits
// with m local bits sends amplitude (rank R, local victim bits L) to (rank L, bits R): rank R trades
// its sub-block L = P with peer P's sub-block L = R, for each of the 2^m - 1 peers P.  Element j of my
// sub-block (victim bits == peer_value) trades places with element j of the peer's sub-block (victim
// bits == mine_value).  The two ranks of a pair run this on disjoint halves of the j range, so every
// link direction carries reads issued by one side and writes issued by the other.  No staging
// buffers: HBM sees one read + one write per moved amplitude on each side.
struct ExchangeBits {
    int m;                  // victim bits
    uint8_t pos[8];         // their local positions, ascending
    uint64_t mine_bits;     // the victim bits set to the value selecting MY moving sub-block
    uint64_t peer_bits;     // ... and the PEER's moving sub-block
};
constexpr int kExchangeUnroll = 4;
__device__ __forceinline__ uint64_t exchange_expand(uint64_t j, const ExchangeBits& b) {
    for (int i = 0; i < b.m; ++i) j = insert_zero64(j, b.pos[i]);
    return j;
}
__global__ void __launch_bounds__(256) exchange_block_kernel(double2* __restrict__ mine, double2* __restrict__ peer,
                                                             const ExchangeBits b, uint64_t off, uint64_t count) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (; j + (kExchangeUnroll - 1) * stride < count; j += kExchangeUnroll * stride) {
        double2 x[kExchangeUnroll], y[kExchangeUnroll];
        uint64_t e[kExchangeUnroll];
#pragma unroll
        for (int u = 0; u < kExchangeUnroll; ++u) e[u] = exchange_expand(off + j + u * stride, b);
#pragma unroll
        for (int u = 0; u < kExchangeUnroll; ++u) y[u] = __ldcs(&peer[e[u] | b.peer_bits]);
#pragma unroll
        for (int u = 0; u < kExchangeUnroll; ++u) x[u] = __ldcs(&mine[e[u] | b.mine_bits]);
#pragma unroll
        for (int u = 0; u < kExchangeUnroll; ++u) __stcs(&mine[e[u] | b.mine_bits], y[u]);
#pragma unroll
        for (int u = 0; u < kExchangeUnroll; ++u) __stcs(&peer[e[u] | b.peer_bits], x[u]);
    }
    for (; j < count; j += stride) {
        const uint64_t e = exchange_expand(off + j, b);
        const double2 y = peer[e | b.peer_bits], x = mine[e | b.mine_bits];
        mine[e | b.mine_bits] = y;
        peer[e | b.peer_bits] = x;
    }
}


// Out-of-place global<->local qubit exchange: ONE pass that reads this shard and pushes every amplitude
// to its new owner (PushSpec, qvm_regtile.cuh) -- write-only NVLink traffic, no read of the peer before
// the write as in the in-place swap above, so each link direction carries plain posted stores.  Every
// rank runs it at the same time into the peers' ALTERNATE buffers; one barrier afterwards, then the
// buffers swap roles.  Runs of 2^pos[0] consecutive elements share a destination, so warps store
// contiguous 512-byte segments.  The fused form of the same exchange is the last round of
// regtile_kernel / spec_kernel storing through rt_push_base().
constexpr int kPushUnroll = 8;
// The loop index j is NOT the element index: its bits [c, c + m) (c = min(pos[0], 7)) are the victim bits, so
// consecutive 2^c-element runs go to different destinations in turn and the NVLink stores are spread over the
// whole duration of the kernel (walking the shard in address order would copy the local part first and leave
// the links idle meanwhile, then leave HBM idle while the links drain).
__device__ __forceinline__ uint64_t push_element(const PushSpec& ps, uint64_t j, uint32_t c, uint32_t& L) {
    const uint64_t lo = j & ((1ull << c) - 1ull);
    L = (uint32_t)(j >> c) & ((1u << ps.m) - 1u);
    uint64_t e = lo | ((j >> (c + ps.m)) << c);
    for (int i = 0; i < ps.m; ++i) e = insert_zero64(e, ps.pos[i]) | ((uint64_t)((L >> i) & 1u) << ps.pos[i]);
    return e;
}
__global__ void __launch_bounds__(256) exchange_push_kernel(const double2* __restrict__ src,
                                                            const __grid_constant__ PushSpec ps, uint64_t n_elems) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t c = ps.pos[0] < 7 ? ps.pos[0] : 7;
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (; j + (kPushUnroll - 1) * stride < n_elems; j += kPushUnroll * stride) {
        double2 x[kPushUnroll];
        double2* d[kPushUnroll];
#pragma unroll
        for (int u = 0; u < kPushUnroll; ++u) {
            uint32_t L;
            const uint64_t e = push_element(ps, j + u * stride, c, L);
            x[u] = __ldcs(&src[e]);
            d[u] = ps.dst[L] + ((e & ~ps.mask) | ps.mine_bits);
        }
#pragma unroll
        for (int u = 0; u < kPushUnroll; ++u) __stcs(d[u], x[u]);
    }
    for (; j < n_elems; j += stride) {
        uint32_t L;
        const uint64_t e = push_element(ps, j, c, L);
        ps.dst[L][(e & ~ps.mask) | ps.mine_bits] = src[e];
    }
}

}  // namespace qvm
