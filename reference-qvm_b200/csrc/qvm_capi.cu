// C-ABI entry points of libqvm_b200.so (see include/qvm_b200.h for the contract and the reference
// code each call replaces).  Host logic here is deliberately thin: validation, gate classification,
// tile-local op encoding and kernel launches.  No CPU fallback exists: without a CUDA device
// qvm_create fails.
#include "qvm_kernels.cuh"
#include "qvm_regtile.cuh"
#include "qvm_encode.h"
#include "qvm_jit.h"
#include "qvm_planner.h"
#include "qvm_regtile_src.inc"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <vector>

#include "../../include/qvm_b200.h"

using namespace qvm;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

constexpr size_t kRingBytes = 8u << 20;     // op-descriptor ring (pinned host + device mirror)
constexpr int kReduceThreads = 256;

}  // namespace

// One fused group ready to launch (see build_rec / launch_rec below).
enum { REC_NONE = 0, REC_REGTILE = 1, REC_TILE = 2 };
struct LaunchRec {
    int kind = REC_NONE;              // nothing to run | register-tiled (interpreter and / or specialised) | shared-memory kernel
    EncodedGroup g;                   // REC_REGTILE: what the interpreter reads
    bool has_spec = false;            // ... and, when the generator covers the group, its specialised form
    SpecSource spec;
    const CompiledSpec* cs = nullptr; // resolved lazily (compiled kernels live as long as the process)
    TileParams tp{};                  // REC_TILE
    std::vector<DevOp> dev;
    std::vector<double> mats;
    uint64_t executed = 0;
    uint64_t tile_mask = 0;
};

// A compiled program: the launches of one flush, kept for replay (qvm_record_begin / qvm_plan_run).
struct qvm_plan {
    int n_local = 0, n_global = 0;
    uint64_t shard_index = 0;
    std::vector<LaunchRec> recs;
    cudaGraphExec_t graph = nullptr;  // the whole sequence as ONE graph launch once every group has its kernel
    double2* graph_amps = nullptr;    // buffer the graph was captured on
    int graph_reset = 0;
    qvm_stats_t graph_stats{};        // what one graph launch adds to the handle's counters
    bool all_specialised = false;     // the last plain replay used a specialised kernel for every group
    bool graph_failed = false;
    uint64_t runs = 0;
};

struct qvm_state {
    int device = 0;
    int n_local = 0;
    int n_global = 0;
    uint64_t shard_index = 0;
    uint64_t n_elems = 0;
    double2* amps = nullptr;
    double2* alt = nullptr;           // second buffer of the same size (out-of-place exchange); roles swap at every push
    bool owns_amps = true;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    int tile_bits = 12;
    int low_bits = 5;
    int use_regtile = 1;
    int reg_bits = 4;                 // register bits per round of the register-tiled kernel (4 or 5)
    int sticky = 0;
    // op ring
    unsigned char* ring_host = nullptr;
    unsigned char* ring_dev = nullptr;
    size_t ring_pos = 0;
    // small scratch: reduction partials + results
    double* partials = nullptr;     // device, kPartialsCap doubles
    double* result_dev = nullptr;   // device, 8 doubles
    double* result_host = nullptr;  // pinned, 8 doubles
    double* pow_dev = nullptr;      // device, 65 doubles (dephasing powers)
    // sampling workspace
    double* cdf = nullptr;
    uint64_t cdf_blocks = 0;
    uint64_t cdf_cap = 0;
    int sample_mode = -1;
    int sample_n_rho = 0;
    uint64_t sample_n_probs = 0;
    double* total_dev = nullptr;
    // staging for host<->device traffic
    unsigned char* stage = nullptr;   // pinned
    size_t stage_bytes = 0;
    int measure_grid = 0;
    int zero_pending = 0;             // lazy reset: 1 = the shard is (virtually) all zeros, 2 = zeros with amp[0] = 1
    qvm_stats_t stats{};
    uint64_t opc_hist[32] = {};
    LaunchRec rec;                    // scratch: the group being built / launched (reused across launches)
    qvm_plan* recording = nullptr;    // non-null between qvm_record_begin and qvm_record_end
    int jit_policy = 1;               // 0 never, 1 compile a group structure at its second sighting, 2 at once
    // pacing (launch_rec): events after the last two gate passes, so a launch whose kernel is still compiling can wait
    // for it exactly as long as the GPU has earlier passes queued
    cudaEvent_t pace_ev[2] = {nullptr, nullptr};
    uint64_t pace_n = 0;
    bool capturing = false;
    bool exported = false;            // the buffer's IPC handle was given out: never recycled through the parking lot
};

namespace {

constexpr int kPartialsCap = 4096;

#define QVM_CUDA(q, expr)                                                                              \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            if (q) (q)->sticky = QVM_ERR_CUDA;                                                         \
            return fail(QVM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                                     \
        }                                                                                              \
    } while (0)

#define QVM_CHECK_HANDLE_LAZY(q)                                                   \
    do {                                                                           \
        if (!(q)) return fail(QVM_ERR_INVALID, "null handle");                     \
        if ((q)->sticky) return (q)->sticky; /* message of the original failure is kept */ \
        cudaError_t _e = cudaSetDevice((q)->device);                               \
        if (_e != cudaSuccess) return fail(QVM_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(_e)); \
    } while (0)

// Every entry point that reads or writes amplitudes first makes a pending reset real (materialize_zero below); only
// the fused-group launches consume it instead (their first round then reads nothing).
#define QVM_CHECK_HANDLE(q)                                                        \
    do {                                                                           \
        QVM_CHECK_HANDLE_LAZY(q);                                                  \
        if ((q)->zero_pending) {                                                   \
            const int _rc = materialize_zero(q);                                   \
            if (_rc != QVM_OK) return _rc;                                         \
        }                                                                          \
    } while (0)

int materialize_zero(qvm_state* q);

int grid_for(const qvm_state* q, uint64_t work_items, int threads, int per_sm = 8) {
    uint64_t blocks = (work_items + threads - 1) / threads;
    uint64_t cap = (uint64_t)q->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int total_positions(const qvm_state* q) { return q->n_local + q->n_global; }

// the reset that qvm_reset_zero only noted down, as a kernel
int materialize_zero(qvm_state* q) {
    if (!q->zero_pending) return QVM_OK;
    const int set_amp0 = q->zero_pending == 2;
    q->zero_pending = 0;
    fill_zero_kernel<<<grid_for(q, q->n_elems, 256), 256, 0, q->stream>>>(q->amps, q->n_elems, set_amp0);
    QVM_CUDA(q, cudaGetLastError());
    q->stats.kernel_launches += 1;
    q->stats.hbm_bytes += 16ull * q->n_elems;
    return QVM_OK;
}

// ---- gate classification -----------------------------------------------------------------------
struct cplx {
    double re, im;
};
inline bool is_zero(cplx a) { return a.re == 0.0 && a.im == 0.0; }
inline bool is_one(cplx a) { return a.re == 1.0 && a.im == 0.0; }

int classify(int k, const int32_t* qubits, const double* matrix, qvm_op_t* op, double* mat_out) {
    if (k < 1) return fail(QVM_ERR_INVALID, "gate must act on at least one qubit");
    if (k > 10) return fail(QVM_ERR_UNSUPPORTED, "gates on more than 10 qubits are not supported");
    const int dim = 1 << k;
    std::vector<cplx> m((size_t)dim * dim);
    for (size_t i = 0; i < m.size(); ++i) m[i] = cplx{matrix[2 * i], matrix[2 * i + 1]};
    std::vector<int> tq(qubits, qubits + k);      // remaining targets, MSB first
    uint64_t ctrl = 0;

    bool diag = true;
    for (int r = 0; r < dim && diag; ++r)
        for (int c = 0; c < dim; ++c)
            if (r != c && !is_zero(m[(size_t)r * dim + c])) {
                diag = false;
                break;
            }

    if (diag) {
        std::vector<cplx> d(dim);
        for (int r = 0; r < dim; ++r) d[r] = m[(size_t)r * dim + r];
        // a qubit whose 0-half is all ones is a control
        bool again = true;
        while (again && !tq.empty()) {
            again = false;
            const int kk = (int)tq.size();
            for (int j = 0; j < kk; ++j) {
                const int bit = kk - 1 - j;
                bool ctrl_like = true;
                for (int r = 0; r < (1 << kk); ++r)
                    if (!((r >> bit) & 1) && !is_one(d[r])) {
                        ctrl_like = false;
                        break;
                    }
                if (!ctrl_like) continue;
                std::vector<cplx> nd(1 << (kk - 1));
                for (int r = 0; r < (1 << (kk - 1)); ++r) nd[r] = d[insert_zero32((uint32_t)r, (uint32_t)bit) | (1u << bit)];
                d.swap(nd);
                ctrl |= 1ull << tq[j];
                tq.erase(tq.begin() + j);
                again = true;
                break;
            }
        }
        if (tq.empty() && is_one(d[0])) {
            op->kind = QVM_OP_NOP;
            op->k = 0;
            op->ctrl_mask = 0;
            op->mat_entries = 0;
            op->mat_offset = 0;
            for (int j = 0; j < QVM_MAX_TARGETS; ++j) op->targets[j] = -1;
            return QVM_OK;
        }
        if ((int)tq.size() > QVM_MAX_TARGETS)
            return fail(QVM_ERR_UNSUPPORTED, "diagonal gate on %d qubits exceeds %d", (int)tq.size(), QVM_MAX_TARGETS);
        op->kind = QVM_OP_DIAG;
        op->k = (int)tq.size();
        for (int j = 0; j < QVM_MAX_TARGETS; ++j) op->targets[j] = j < op->k ? tq[j] : -1;
        op->ctrl_mask = ctrl;
        op->mat_entries = (uint32_t)d.size();
        op->mat_offset = 0;
        for (size_t i = 0; i < d.size(); ++i) {
            mat_out[2 * i] = d[i].re;
            mat_out[2 * i + 1] = d[i].im;
        }
        return QVM_OK;
    }

    // dense: peel off controls (identity on the 0-half, no coupling between halves)
    bool again = true;
    while (again && tq.size() > 1) {
        again = false;
        const int kk = (int)tq.size();
        const int dd = 1 << kk;
        for (int j = 0; j < kk; ++j) {
            const int bit = kk - 1 - j;
            bool ctrl_like = true;
            for (int r = 0; r < dd && ctrl_like; ++r)
                for (int c = 0; c < dd; ++c) {
                    const int rb = (r >> bit) & 1, cb = (c >> bit) & 1;
                    if (rb && cb) continue;
                    const cplx v = m[(size_t)r * dd + c];
                    if (r == c ? !is_one(v) : !is_zero(v)) {
                        ctrl_like = false;
                        break;
                    }
                }
            if (!ctrl_like) continue;
            const int nd = dd >> 1;
            std::vector<cplx> nm((size_t)nd * nd);
            for (int r = 0; r < nd; ++r)
                for (int c = 0; c < nd; ++c) {
                    const uint32_t rr = insert_zero32((uint32_t)r, (uint32_t)bit) | (1u << bit);
                    const uint32_t cc = insert_zero32((uint32_t)c, (uint32_t)bit) | (1u << bit);
                    nm[(size_t)r * nd + c] = m[(size_t)rr * dd + cc];
                }
            m.swap(nm);
            ctrl |= 1ull << tq[j];
            tq.erase(tq.begin() + j);
            again = true;
            break;
        }
    }
    if ((int)tq.size() > QVM_MAX_TARGETS)
        return fail(QVM_ERR_UNSUPPORTED, "dense gate on %d qubits exceeds the supported %d", (int)tq.size(),
                    QVM_MAX_TARGETS);
    op->kind = QVM_OP_DENSE;
    op->k = (int)tq.size();
    for (int j = 0; j < QVM_MAX_TARGETS; ++j) op->targets[j] = j < op->k ? tq[j] : -1;
    op->ctrl_mask = ctrl;
    op->mat_entries = (uint32_t)m.size();
    op->mat_offset = 0;
    for (size_t i = 0; i < m.size(); ++i) {
        mat_out[2 * i] = m[i].re;
        mat_out[2 * i + 1] = m[i].im;
    }
    return QVM_OK;
}

int validate_qubits(const qvm_state* q, int k, const int32_t* qubits) {
    const int npos = total_positions(q);
    uint64_t seen = 0;
    for (int j = 0; j < k; ++j) {
        if (qubits[j] < 0 || qubits[j] >= npos)
            return fail(QVM_ERR_INVALID, "qubit index %d out of range [0, %d)", qubits[j], npos);
        if ((seen >> qubits[j]) & 1ull) return fail(QVM_ERR_INVALID, "duplicate qubit index %d", qubits[j]);
        seen |= 1ull << qubits[j];
    }
    return QVM_OK;
}

// ---- op ring -----------------------------------------------------------------------------------
int ring_alloc(qvm_state* q, size_t bytes, size_t* off) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes > kRingBytes) return fail(QVM_ERR_UNSUPPORTED, "op group of %zu bytes exceeds the descriptor ring", bytes);
    if (q->ring_pos + bytes > kRingBytes) {
        QVM_CUDA(q, cudaStreamSynchronize(q->stream));      // wrap: everything queued has consumed the ring
        q->ring_pos = 0;
    }
    *off = q->ring_pos;
    q->ring_pos += bytes;
    return QVM_OK;
}

// Relabelling request of one group: tile bits (tile-local mask) whose qubits should leave the pass on the
// n_low lowest tile bits; sigma[b] = tile bit on which the qubit of tile bit b does leave it.
struct RelabelReq {
    uint32_t low_in = 0;
    int n_low = 0;
    uint8_t sigma[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};
};

// ---- one fused group, ready to launch (and to launch again: compiled programs replay these) -----------
// build_rec() does everything that depends only on the group -- validation, tile-local encoding, the
// round plan, the specialised source -- and launch_rec() everything that touches the device: pick the
// specialised kernel when it is ready (else the interpreter), upload, launch.
int build_rec(qvm_state* q, int n_tile, const int32_t* tile_bits, int n_ops, const qvm_op_t* ops, const double* mats,
              uint64_t mats_entries, RelabelReq* relabel, LaunchRec& rec) {
    rec.kind = REC_NONE;
    rec.has_spec = false;
    rec.cs = nullptr;
    rec.tile_mask = 0;
    if (n_tile < 0 || n_tile > q->n_local || n_tile > QVM_MAX_TILE_BITS)
        return fail(QVM_ERR_INVALID, "tile of %d bits is outside [0, min(n_local=%d, %d)]", n_tile, q->n_local,
                    QVM_MAX_TILE_BITS);
    TileParams& P = rec.tp;
    P = TileParams{};
    P.n_local = q->n_local;
    P.k = n_tile;
    P.shard_bits = q->shard_index << q->n_local;
    int8_t local_of[64];
    memset(local_of, -1, sizeof(local_of));
    for (int j = 0; j < n_tile; ++j) {
        const int p = tile_bits[j];
        if (p < 0 || p >= q->n_local) return fail(QVM_ERR_INVALID, "tile bit %d is not a local position", p);
        if (j && p <= tile_bits[j - 1]) return fail(QVM_ERR_INVALID, "tile bits must be strictly ascending");
        P.tile_pos[j] = (uint8_t)p;
        P.tile_mask |= 1ull << p;
        local_of[p] = (int8_t)j;
    }
    rec.tile_mask = P.tile_mask;
    int low = 0;
    while (low < n_tile && tile_bits[low] == low) ++low;
    // keep the offset table small: at most 2^8 entries come from the high part
    if (n_tile - low > 10) return fail(QVM_ERR_INVALID, "tile needs at least %d contiguous low bits", n_tile - 10);
    P.low = low;

    if (q->use_regtile) {
        EncodedGroup& g = rec.g;
        const int erc = encode_group(q->n_local, q->n_global, q->shard_index, n_tile, tile_bits, n_ops, ops, mats,
                                     mats_entries, g, q->reg_bits, relabel ? relabel->low_in : 0u, relabel ? relabel->n_low : 0);
        if (erc == ENC_ERROR) return fail(QVM_ERR_CUDA, "round planner made no progress");
        if (erc == ENC_OK) {
            if (relabel) memcpy(relabel->sigma, g.sigma, sizeof(g.sigma));
            if (g.cold.empty() && !g.relabelled) return QVM_OK;        // nothing to run
            if (g.n_tiles > 0x7fffffffull) return fail(QVM_ERR_UNSUPPORTED, "too many tiles for one launch");
            for (int i = 0; i < 32; ++i) q->opc_hist[i] += g.opc_hist[i];
            rec.kind = REC_REGTILE;
            // a structure that comes back runs as a specialised straight-line kernel (qvm_jit.h)
            rec.has_spec = !g.need_full && g.P.R == 4 && generate_spec(g, rec.spec);
            return QVM_OK;
        }
        // ENC_FALLBACK: the shared-memory kernel below (or its error) handles it
    }
    // (the shared-memory kernel never relabels: relabel->sigma is still the identity)

    std::vector<DevOp>& dev = rec.dev;
    dev.clear();
    dev.reserve(n_ops);
    const int npos = total_positions(q);
    rec.executed = 0;
    for (int i = 0; i < n_ops; ++i) {
        const qvm_op_t& op = ops[i];
        if (op.kind == QVM_OP_NOP) continue;
        if (op.kind != QVM_OP_DENSE && op.kind != QVM_OP_DIAG) return fail(QVM_ERR_INVALID, "op %d: bad kind %d", i, op.kind);
        if (op.k < 0 || op.k > QVM_MAX_TARGETS) return fail(QVM_ERR_INVALID, "op %d: bad target count %d", i, op.k);
        const uint64_t want = op.kind == QVM_OP_DENSE ? (1ull << (2 * op.k)) : (1ull << op.k);
        if (op.mat_entries != want || (uint64_t)op.mat_offset + want > mats_entries)
            return fail(QVM_ERR_INVALID, "op %d: matrix pool reference out of range", i);
        if (npos < 64 && (op.ctrl_mask >> npos)) return fail(QVM_ERR_INVALID, "op %d: control outside the register", i);
        DevOp d{};
        d.kind = op.kind == QVM_OP_DENSE ? DK_DENSE : DK_DIAG;
        d.k = (uint8_t)op.k;
        d.mat_off = op.mat_offset;
        uint64_t used = op.ctrl_mask;
        for (int p = 0; p < q->n_local; ++p)
            if (((op.ctrl_mask >> p) & 1ull) && local_of[p] >= 0) d.ctrl_tile |= 1u << local_of[p];
        d.ctrl_outer = op.ctrl_mask & ~P.tile_mask;
        uint32_t fixmask = d.ctrl_tile;
        for (int j = 0; j < op.k; ++j) {
            const int p = op.targets[j];
            if (p < 0 || p >= npos) return fail(QVM_ERR_INVALID, "op %d: target %d out of range", i, p);
            if ((used >> p) & 1ull) return fail(QVM_ERR_INVALID, "op %d: qubit %d used twice", i, p);
            used |= 1ull << p;
            if (op.kind == QVM_OP_DENSE) {
                if (p >= q->n_local || local_of[p] < 0)
                    return fail(QVM_ERR_INVALID, "op %d: dense target %d is not a tile bit", i, p);
                d.tgt[j] = (uint8_t)local_of[p];
                fixmask |= 1u << local_of[p];
            } else {
                d.tgt[j] = (p < q->n_local && local_of[p] >= 0) ? (uint8_t)local_of[p] : (uint8_t)(0x80 | p);
            }
        }
        if (op.kind == QVM_OP_DENSE) {
            if (op.k < 1) return fail(QVM_ERR_INVALID, "op %d: dense op without targets", i);
            int nf = 0;
            for (int b = 0; b < n_tile; ++b)
                if ((fixmask >> b) & 1u) d.fix[nf++] = (uint8_t)b;
            d.nfix = (uint8_t)nf;
        }
        dev.push_back(d);
        ++rec.executed;
    }
    if (dev.empty()) return QVM_OK;
    rec.mats.assign(mats, mats + 2 * mats_entries);
    P.n_ops = (int)dev.size();
    const size_t smem = tile_smem_bytes(P.k, P.low, P.n_ops);
    if (smem > 227 * 1024) return fail(QVM_ERR_UNSUPPORTED, "group needs %zu bytes of shared memory (> 227 KiB)", smem);
    if ((q->n_elems >> n_tile) > 0x7fffffffull) return fail(QVM_ERR_UNSUPPORTED, "too many tiles for one launch");
    rec.kind = REC_TILE;
    return QVM_OK;
}

template <int T, int B, bool FULL, int R>
int launch_regtile(qvm_state* q, unsigned grid, unsigned threads, size_t smem, const RegTileParams& P) {
    static bool attr_set[64] = {};
    if (!attr_set[q->device & 63]) {
        QVM_CUDA(q, cudaFuncSetAttribute(regtile_kernel<T, B, FULL, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024));
        attr_set[q->device & 63] = true;
    }
    regtile_kernel<T, B, FULL, R><<<grid, threads, smem, q->stream>>>(q->amps, P);
    QVM_CUDA(q, cudaGetLastError());
    return QVM_OK;
}

// QVM_B200_FOLD_RESET=0: always reset with a kernel of its own (A/B runs)
bool fold_reset_enabled() {
    static const bool on = [] {
        const char* e = getenv("QVM_B200_FOLD_RESET");
        return !e || atoi(e) != 0;
    }();
    return on;
}

// QVM_B200_PATIENT=0: never wait for a kernel that is still compiling (A/B runs)
bool patient_enabled() {
    static const bool on = [] {
        const char* e = getenv("QVM_B200_PATIENT");
        return !e || atoi(e) != 0;
    }();
    return on;
}

// After a gate pass on a large shard while compilations are in flight: remember where the stream is.
void pace_mark(qvm_state* q) {
    if (q->n_local < 26 || q->capturing || q->jit_policy != 1 || !SpecCache::instance().busy()) return;
    cudaEvent_t& ev = q->pace_ev[q->pace_n & 1];
    if (!ev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
        ev = nullptr;
        cudaGetLastError();
        return;
    }
    if (cudaEventRecord(ev, q->stream) == cudaSuccess) ++q->pace_n;
    else cudaGetLastError();
}

// true while the pass before the last one is still unfinished, i.e. the GPU has a whole pass queued behind the one
// it is running: a launch may wait for its kernel without the device ever going idle
bool pace_backlog(qvm_state* q) {
    if (q->pace_n < 2) return false;
    cudaEvent_t ev = q->pace_ev[q->pace_n & 1];       // the older of the two
    if (!ev) return false;
    const cudaError_t e = cudaEventQuery(ev);
    if (e == cudaErrorNotReady) return true;
    if (e != cudaSuccess) cudaGetLastError();
    return false;
}

// The order in which a pushing pass walks its tiles (the runs of non-tile positions the block index is deposited
// into are launch parameters) decides when it stores where.  QVM_B200_PUSH_ORDER:
//   0 (default): the victim bits are the FASTEST-varying bits of the tile index -- consecutive CTAs store to
//      different destinations in turn, so the NVLink stores are spread over the whole kernel instead of all local
//      tiles running first and all remote ones after them;
//   1: the victim bits are the SLOWEST bits, XOR-rotated by the rank's own global bits (PushSpec::rot): the pass
//      goes through its destinations one after the other and at any time every rank stores to a different peer;
//   2: the lowest victim bit fastest, the others slowest and rotated: two destinations at a time.
// More than 8 runs: the order is left as it is.
int push_order() {
    static const int mode = [] {
        const char* e = getenv("QVM_B200_PUSH_ORDER");
        const int v = e ? atoi(e) : 0;
        return v < 0 || v > 2 ? 0 : v;
    }();
    return mode;
}

void runs_for_push(PushSpec& push, int& n_runs, uint64_t& run_start, uint64_t& run_len) {
    struct Run { int start, len; };
    std::vector<Run> runs, head, tail;
    const int mode = push.m >= 2 ? push_order() : 0;
    uint64_t rot_mask = 0;
    for (int r = 0; r < n_runs; ++r)
        runs.push_back(Run{(int)((run_start >> (8 * r)) & 0xFF), (int)((run_len >> (8 * r)) & 0xFF)});
    for (int i = 0; i < push.m; ++i) {
        const int p = push.pos[i];
        for (size_t r = 0; r < runs.size(); ++r) {
            const Run cur = runs[r];
            if (p < cur.start || p >= cur.start + cur.len) continue;
            std::vector<Run> parts;
            if (p > cur.start) parts.push_back(Run{cur.start, p - cur.start});
            if (p + 1 < cur.start + cur.len) parts.push_back(Run{p + 1, cur.start + cur.len - p - 1});
            runs.erase(runs.begin() + r);
            runs.insert(runs.begin() + r, parts.begin(), parts.end());
            if (mode == 1 || (mode == 2 && i > 0)) {
                tail.push_back(Run{p, 1});
                rot_mask |= 1ull << p;
            } else {
                head.push_back(Run{p, 1});
            }
            break;
        }
    }
    if ((head.empty() && tail.empty()) || head.size() + runs.size() + tail.size() > 8) return;
    head.insert(head.end(), runs.begin(), runs.end());
    head.insert(head.end(), tail.begin(), tail.end());
    n_runs = (int)head.size();
    run_start = run_len = 0;
    for (size_t r = 0; r < head.size(); ++r) {
        run_start |= (uint64_t)head[r].start << (8 * r);
        run_len |= (uint64_t)head[r].len << (8 * r);
    }
    push.rot = push.mine_bits & rot_mask;
}

// With `push` (a fused exchange, see PushSpec) the register-tiled kernels store the pass straight into the
// destination shards and *pushed is set; when the group cannot carry it (nothing to launch, shared-memory
// fallback kernel, victim bit inside the tile) the pass runs in place and the caller pushes separately.
int launch_rec(qvm_state* q, LaunchRec& rec, const PushSpec* push = nullptr, bool* pushed = nullptr) {
    if (rec.kind == REC_NONE) return QVM_OK;
    if (push && push->m && (rec.tile_mask & push->mask)) push = nullptr;     // a victim bit inside the tile: not fusable
    // a pending reset is folded into this pass when the register-tiled kernels run it (their round 0 then reads
    // nothing); the other kernels get a real reset first
    int from_zero = q->zero_pending;
    if (from_zero && (rec.kind != REC_REGTILE || (rec.has_spec && rec.spec.ctas_per_sm) || !fold_reset_enabled())) {
        const int rc0 = materialize_zero(q);
        if (rc0 != QVM_OK) return rc0;
        from_zero = 0;
    }
    if (rec.kind == REC_TILE) {
        const size_t ops_bytes = rec.dev.size() * sizeof(DevOp);
        const size_t mats_bytes = rec.mats.size() * sizeof(double);
        const size_t ops_pad = (ops_bytes + 255) & ~(size_t)255;
        size_t off = 0;
        int rc = ring_alloc(q, ops_pad + mats_bytes, &off);
        if (rc) return rc;
        memcpy(q->ring_host + off, rec.dev.data(), ops_bytes);
        memcpy(q->ring_host + off + ops_pad, rec.mats.data(), mats_bytes);
        QVM_CUDA(q, cudaMemcpyAsync(q->ring_dev + off, q->ring_host + off, ops_pad + mats_bytes, cudaMemcpyHostToDevice,
                                    q->stream));
        TileParams P = rec.tp;
        P.ops = reinterpret_cast<const DevOp*>(q->ring_dev + off);
        P.mats = reinterpret_cast<const double2*>(q->ring_dev + off + ops_pad);
        const size_t smem = tile_smem_bytes(P.k, P.low, P.n_ops);
        tile_kernel<<<(unsigned)(q->n_elems >> P.k), kTileThreads, smem, q->stream>>>(q->amps, P);
        QVM_CUDA(q, cudaGetLastError());
        q->stats.kernel_launches += 1;
        q->stats.gate_ops += rec.executed;
        q->stats.hbm_passes += 1;
        q->stats.hbm_bytes += 32ull * q->n_elems;
        q->stats.h2d_bytes += ops_pad + mats_bytes;
        return QVM_OK;
    }

    EncodedGroup& g = rec.g;
    if (rec.has_spec && q->jit_policy > 0) {
        // large states compile a structure at its first sighting (an interpreted pass there costs more than waiting
        // for a second call is worth); small ones at the second, so one-off programs do not keep NVRTC busy
        if (!rec.cs) rec.cs = SpecCache::instance().get(rec.spec.src, q->jit_policy, q->device, q->n_local >= 26);
        if (!rec.cs && q->n_local >= 26 && q->jit_policy == 1 && !q->capturing && patient_enabled()) {
            // The kernel of this structure is being compiled (ahead-of-time walk of the program, or queued just now).
            // The launches before this one are still on the stream: waiting costs no device time until the pass
            // before this one is the only work left, and a specialised pass is ~half an interpreted one.
            SpecCache& sc = SpecCache::instance();
            bool waited = false;
            while (sc.on_its_way(rec.spec.src) && pace_backlog(q)) {
                waited = true;
                usleep(200);
            }
            if (waited || !sc.on_its_way(rec.spec.src))
                rec.cs = sc.get(rec.spec.src, q->jit_policy, q->device, true);
            q->stats.paced_waits += waited ? 1 : 0;
        }
        q->stats.jit_compiles = (uint64_t)SpecCache::instance().compiles();
        if (const CompiledSpec* cs = rec.cs) {
            const SpecSource& ss = rec.spec;
            if (ss.smem > 48 * 1024 && const_cast<CompiledSpec*>(cs)->smem_attr < ss.smem) {
                QVM_CUDA(q, cudaFuncSetAttribute((const void*)cs->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)ss.smem));
                const_cast<CompiledSpec*>(cs)->smem_attr = ss.smem;
            }
            double2* amps = q->amps;
            SpecParams sp = ss.params;
            sp.from_zero = from_zero;
            if (push && push->m) {
                sp.push = *push;
                runs_for_push(sp.push, sp.n_runs, sp.run_start, sp.run_len);
                if (pushed) *pushed = true;
            }
            void* args[3] = {&amps, &sp, const_cast<double*>(ss.factors.data())};
            unsigned grid = (unsigned)g.n_tiles;
            if (ss.ctas_per_sm)      // persistent, pipelined variant
                grid = (unsigned)std::min<uint64_t>(g.n_tiles, (uint64_t)q->sm_count * ss.ctas_per_sm);
            QVM_CUDA(q, cudaLaunchKernel((const void*)cs->kernel, dim3(grid), dim3(ss.threads), args, ss.smem, q->stream));
            q->zero_pending = 0;
            q->stats.kernel_launches += 1;
            q->stats.jit_launches += 1;
            q->stats.gate_ops += (uint64_t)g.n_gate_ops;
            q->stats.hbm_passes += 1;
            q->stats.hbm_bytes += (from_zero ? 16ull : 32ull) * q->n_elems;
            q->stats.zero_fused_launches += from_zero ? 1 : 0;
            q->stats.h2d_bytes += ss.factors.size() * sizeof(double);
            q->stats.regtile_rounds += (uint64_t)g.rounds.size();
            pace_mark(q);
            return QVM_OK;
        }
    }

    const size_t rounds_bytes = (g.rounds.size() * sizeof(RoundDesc) + 255) & ~(size_t)255;
    const size_t hot_bytes = (g.hot.size() * sizeof(uint4) + 255) & ~(size_t)255;
    const size_t cold_bytes = (g.cold.size() * sizeof(ColdOp) + 255) & ~(size_t)255;
    const size_t hi_bytes = (g.hi_off.size() * sizeof(uint64_t) + 255) & ~(size_t)255;
    const size_t mats_bytes = g.pool.size() * sizeof(double);
    const size_t total_bytes = rounds_bytes + hot_bytes + cold_bytes + hi_bytes + mats_bytes;
    size_t off = 0;
    int rc = ring_alloc(q, total_bytes, &off);
    if (rc) return rc;
    unsigned char* hp = q->ring_host + off;
    memcpy(hp, g.rounds.data(), g.rounds.size() * sizeof(RoundDesc));
    memcpy(hp + rounds_bytes, g.hot.data(), g.hot.size() * sizeof(uint4));
    memcpy(hp + rounds_bytes + hot_bytes, g.cold.data(), g.cold.size() * sizeof(ColdOp));
    memcpy(hp + rounds_bytes + hot_bytes + cold_bytes, g.hi_off.data(), g.hi_off.size() * sizeof(uint64_t));
    memcpy(hp + rounds_bytes + hot_bytes + cold_bytes + hi_bytes, g.pool.data(), mats_bytes);
    QVM_CUDA(q, cudaMemcpyAsync(q->ring_dev + off, hp, total_bytes, cudaMemcpyHostToDevice, q->stream));
    unsigned char* dp = q->ring_dev + off;
    RegTileParams P = g.P;
    P.from_zero = from_zero;
    if (push && push->m) {
        P.push = *push;
        runs_for_push(P.push, P.n_runs, P.run_start, P.run_len);
        if (pushed) *pushed = true;
    }
    P.rounds = reinterpret_cast<const RoundDesc*>(dp);
    P.hot = reinterpret_cast<const uint4*>(dp + rounds_bytes);
    P.cold = reinterpret_cast<const ColdOp*>(dp + rounds_bytes + hot_bytes);
    P.hi_off = reinterpret_cast<const uint64_t*>(dp + rounds_bytes + hot_bytes + cold_bytes);
    P.mats = reinterpret_cast<const double2*>(dp + rounds_bytes + hot_bytes + cold_bytes + hi_bytes);

    // (capping registers to fit 5 or 6 CTAs of 128 threads per SM was measured slower: profiles/README.md)
    const unsigned grid = (unsigned)g.n_tiles;
    if (g.P.R == 5) {
        // 32 amplitudes per thread need ~190+ registers: 64-thread CTAs at 4 per SM, 128 at 2, 256 at 1.
        // Measured slower than R = 4 on B200 (fewer resident warps; profiles/README.md), kept selectable.
        if (g.threads > 128)
            rc = g.need_full ? launch_regtile<256, 1, true, 5>(q, grid, g.threads, g.smem, P)
                             : launch_regtile<256, 1, false, 5>(q, grid, g.threads, g.smem, P);
        else if (g.threads > 64)
            rc = g.need_full ? launch_regtile<128, 2, true, 5>(q, grid, g.threads, g.smem, P)
                             : launch_regtile<128, 2, false, 5>(q, grid, g.threads, g.smem, P);
        else
            rc = g.need_full ? launch_regtile<64, 4, true, 5>(q, grid, g.threads, g.smem, P)
                             : launch_regtile<64, 4, false, 5>(q, grid, g.threads, g.smem, P);
    } else if (g.threads > 256) {
        rc = g.need_full ? launch_regtile<512, 1, true, 4>(q, grid, g.threads, g.smem, P)
                         : launch_regtile<512, 1, false, 4>(q, grid, g.threads, g.smem, P);
    } else {
        rc = g.need_full ? launch_regtile<256, 2, true, 4>(q, grid, g.threads, g.smem, P)
                         : launch_regtile<256, 2, false, 4>(q, grid, g.threads, g.smem, P);
    }
    if (rc) return rc;
    q->zero_pending = 0;
    q->stats.kernel_launches += 1;
    q->stats.gate_ops += (uint64_t)g.n_gate_ops;
    q->stats.hbm_passes += 1;
    q->stats.hbm_bytes += (from_zero ? 16ull : 32ull) * q->n_elems;
    q->stats.zero_fused_launches += from_zero ? 1 : 0;
    q->stats.h2d_bytes += total_bytes;
    q->stats.regtile_rounds += (uint64_t)g.rounds.size();
    pace_mark(q);
    return QVM_OK;
}

// Encode ops in tile-local coordinates and launch one tile kernel; while a program is being compiled
// (qvm_record_begin) the launch is also kept for replay.
int launch_group(qvm_state* q, int n_tile, const int32_t* tile_bits, int n_ops, const qvm_op_t* ops, const double* mats,
                 uint64_t mats_entries, RelabelReq* relabel = nullptr, const PushSpec* push = nullptr,
                 bool* pushed = nullptr) {
    LaunchRec& rec = q->rec;
    int rc = build_rec(q, n_tile, tile_bits, n_ops, ops, mats, mats_entries, relabel, rec);
    if (rc != QVM_OK) return rc;
    rc = launch_rec(q, rec, push, pushed);
    if (rc != QVM_OK) return rc;
    if (q->recording && rec.kind != REC_NONE) {
        if (push && push->m) return fail(QVM_ERR_UNSUPPORTED, "exchanges cannot be part of a compiled program");
        q->recording->recs.push_back(rec);
    }
    return QVM_OK;
}

int reduce_partials(qvm_state* q, int n_partials, double* out_host) {
    final_sum_kernel<<<1, kReduceThreads, 0, q->stream>>>(q->partials, n_partials, q->result_dev);
    QVM_CUDA(q, cudaGetLastError());
    QVM_CUDA(q, cudaMemcpyAsync(q->result_host, q->result_dev, sizeof(double), cudaMemcpyDeviceToHost, q->stream));
    QVM_CUDA(q, cudaStreamSynchronize(q->stream));
    *out_host = q->result_host[0];
    q->stats.kernel_launches += 1;
    q->stats.d2h_bytes += sizeof(double);
    return QVM_OK;
}

int ensure_stage(qvm_state* q, size_t bytes) {
    if (q->stage_bytes >= bytes) return QVM_OK;
    if (q->stage) cudaFreeHost(q->stage);
    q->stage = nullptr;
    q->stage_bytes = 0;
    QVM_CUDA(q, cudaMallocHost(&q->stage, bytes));
    q->stage_bytes = bytes;
    return QVM_OK;
}

constexpr size_t kStageChunk = 64u << 20;   // 64 MiB pinned bounce buffer for pageable host arrays

}  // namespace

// =================================================================================================
extern "C" {

const char* qvm_version(void) { return "qvm_b200 0.1 (sm_100a)"; }
const char* qvm_last_error(void) { return g_err; }

int qvm_device_count(int* count) {
    if (!count) return fail(QVM_ERR_INVALID, "null pointer");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(QVM_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return QVM_OK;
}

// Handles of small and mid-sized states come and go at a high rate -- every density() call returns an independent
// snapshot (a cloned handle), every QVMConnection call on a new register a fresh one -- and a handle is a dozen
// allocations, 8 MiB of them pinned host memory: milliseconds, next to ~10 ms of device work for a whole 14-qubit
// noisy density program.  qvm_destroy therefore parks the complete resource set of a healthy, unshared handle (state
// buffers of at most 8 GiB, at most two sets per process) and qvm_create of the same shape on the same device takes
// it over.  A failed allocation anywhere in qvm_create empties the lot and retries.  QVM_B200_PARK=0 turns it off.
namespace {
struct Parked {
    int device, n_local, sm_count, measure_grid;
    double2* amps;
    unsigned char *ring_host, *ring_dev;
    double *partials, *result_dev, *result_host, *pow_dev, *total_dev;
    cudaStream_t stream;
};
std::mutex g_park_mu;
std::vector<Parked> g_parked;
constexpr uint64_t kParkMaxBytes = 8ull << 30;
constexpr size_t kParkMaxSets = 2;

bool park_enabled() {
    static const bool on = [] {
        const char* e = getenv("QVM_B200_PARK");
        return !e || atoi(e) != 0;
    }();
    return on;
}
void parked_free(const Parked& p) {
    cudaSetDevice(p.device);
    cudaFree(p.amps);
    cudaFreeHost(p.ring_host);
    cudaFree(p.ring_dev);
    cudaFree(p.partials);
    cudaFree(p.result_dev);
    cudaFreeHost(p.result_host);
    cudaFree(p.pow_dev);
    cudaFree(p.total_dev);
    cudaStreamDestroy(p.stream);
}
void parked_release_all() {
    std::vector<Parked> all;
    {
        std::lock_guard<std::mutex> lock(g_park_mu);
        all.swap(g_parked);
    }
    for (const Parked& p : all) parked_free(p);
}
}  // namespace

static int create_impl(int n_local, int device, void* external, qvm_t** out, bool retried = false);

static int create_impl(int n_local, int device, void* external, qvm_t** out, bool retried) {
    if (!out) return fail(QVM_ERR_INVALID, "null output pointer");
    *out = nullptr;
    if (n_local < 0 || n_local > 40) return fail(QVM_ERR_INVALID, "n_local=%d outside [0, 40]", n_local);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(QVM_ERR_CUDA, "no CUDA device available (%s); libqvm_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(QVM_ERR_INVALID, "device %d not in [0, %d)", device, count);
    qvm_state* q = new (std::nothrow) qvm_state();
    if (!q) return fail(QVM_ERR_NOMEM, "host allocation failed");
    q->device = device;
    q->n_local = n_local;
    q->n_elems = 1ull << n_local;
    if (const char* rb = getenv("QVM_B200_REG_BITS")) q->reg_bits = atoi(rb) == 5 ? 5 : 4;      // A/B runs
    if (const char* jp = getenv("QVM_B200_JIT")) q->jit_policy = std::max(0, std::min(2, atoi(jp)));
#define QVM_CREATE_CUDA(expr)                                                                          \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            int code = (_e == cudaErrorMemoryAllocation) ? QVM_ERR_NOMEM : QVM_ERR_CUDA;               \
            fail(code, "%s failed: %s", #expr, cudaGetErrorString(_e));                                \
            cudaGetLastError();                                                                        \
            q->sticky = code; /* (never parked) */                                                     \
            qvm_destroy(q);                                                                            \
            if (code == QVM_ERR_NOMEM && !retried) {                                                   \
                parked_release_all();                                                                  \
                return create_impl(n_local, device, external, out, true);                              \
            }                                                                                          \
            return code;                                                                               \
        }                                                                                              \
    } while (0)
    QVM_CREATE_CUDA(cudaSetDevice(device));
    if (!external && park_enabled()) {
        Parked p{};
        bool found = false;
        {
            std::lock_guard<std::mutex> lock(g_park_mu);
            for (size_t i = 0; i < g_parked.size(); ++i)
                if (g_parked[i].device == device && g_parked[i].n_local == n_local) {
                    p = g_parked[i];
                    g_parked.erase(g_parked.begin() + i);
                    found = true;
                    break;
                }
        }
        if (found) {
            q->sm_count = p.sm_count;
            q->measure_grid = p.measure_grid;
            q->amps = p.amps;
            q->ring_host = p.ring_host;
            q->ring_dev = p.ring_dev;
            q->partials = p.partials;
            q->result_dev = p.result_dev;
            q->result_host = p.result_host;
            q->pow_dev = p.pow_dev;
            q->total_dev = p.total_dev;
            q->own_stream = q->stream = p.stream;
            *out = q;
            return QVM_OK;
        }
    }
    cudaDeviceProp prop;
    QVM_CREATE_CUDA(cudaGetDeviceProperties(&prop, device));
    q->sm_count = prop.multiProcessorCount;
    {
        // stream-ordered scratch allocations (sampling, histograms, readback staging) come from the device's default
        // pool: keep what it has grown to instead of handing it back to the driver at every synchronisation
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            uint64_t keep = 1ull << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    QVM_CREATE_CUDA(cudaStreamCreateWithFlags(&q->own_stream, cudaStreamNonBlocking));
    q->stream = q->own_stream;
    if (external) {
        q->amps = static_cast<double2*>(external);
        q->owns_amps = false;
    } else {
        QVM_CREATE_CUDA(cudaMalloc(&q->amps, q->n_elems * sizeof(double2)));
    }
    QVM_CREATE_CUDA(cudaMallocHost(&q->ring_host, kRingBytes));
    QVM_CREATE_CUDA(cudaMalloc(&q->ring_dev, kRingBytes));
    QVM_CREATE_CUDA(cudaMalloc(&q->partials, kPartialsCap * sizeof(double)));
    QVM_CREATE_CUDA(cudaMalloc(&q->result_dev, 8 * sizeof(double)));
    QVM_CREATE_CUDA(cudaMallocHost(&q->result_host, 8 * sizeof(double)));
    QVM_CREATE_CUDA(cudaMalloc(&q->pow_dev, 65 * sizeof(double)));
    QVM_CREATE_CUDA(cudaMalloc(&q->total_dev, sizeof(double)));
    QVM_CREATE_CUDA(cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int per_sm = 0;
    QVM_CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, measure_kernel, kReduceThreads, 0));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    q->measure_grid = std::min(per_sm * q->sm_count, kPartialsCap);
#undef QVM_CREATE_CUDA
    *out = q;
    return QVM_OK;
}

int qvm_plan_choose_group(int n_ops, const uint8_t* dense, const int32_t* tgt_off, const int32_t* tgt, const int32_t* dep_off,
                          const int32_t* dep, const uint8_t* done, int first, int lookahead, int k, int max_new,
                          uint64_t old_mask, uint64_t seed_mask, int max_group_ops, int depth, int32_t* chosen,
                          int32_t* n_chosen, uint64_t* tile_mask) {
    if (!dense || !tgt_off || !dep_off || !done || !chosen || !n_chosen || !tile_mask || (tgt_off[n_ops] && !tgt) ||
        (dep_off[n_ops] && !dep))
        return fail(QVM_ERR_INVALID, "null pointer");
    if (n_ops < 0 || first < 0 || first > n_ops || lookahead < 1 || k < 1 || k > 64 || max_group_ops < 1 || depth < 0 || depth > 4)
        return fail(QVM_ERR_INVALID, "bad planning request");
    PlanStream s{n_ops, dense, tgt_off, tgt, dep_off, dep, done};
    *n_chosen = plan_choose(s, first, lookahead, k, max_new, old_mask, seed_mask, max_group_ops, depth, chosen, tile_mask);
    return QVM_OK;
}

int qvm_trim(void) {
    parked_release_all();
    return QVM_OK;
}

int qvm_create(int n_local, int device, qvm_t** out) { return create_impl(n_local, device, nullptr, out); }

int qvm_create_external(int n_local, int device, void* device_ptr, qvm_t** out) {
    if (!device_ptr) return fail(QVM_ERR_INVALID, "null device pointer");
    return create_impl(n_local, device, device_ptr, out);
}

int qvm_destroy(qvm_t* q) {
    if (!q) return QVM_OK;
    cudaSetDevice(q->device);
    if (q->stream) cudaStreamSynchronize(q->stream);
    if (park_enabled() && !q->sticky && q->owns_amps && q->amps && !q->alt && !q->exported && !q->n_global && !q->recording &&
        q->own_stream && q->stream == q->own_stream && q->ring_host && q->ring_dev && q->partials && q->result_dev &&
        q->result_host && q->pow_dev && q->total_dev && q->measure_grid && q->n_elems * sizeof(double2) <= kParkMaxBytes &&
        cudaGetLastError() == cudaSuccess) {
        std::unique_lock<std::mutex> lock(g_park_mu);
        uint64_t held = q->n_elems * sizeof(double2);
        for (const Parked& p : g_parked) held += (sizeof(double2) << p.n_local);
        if (g_parked.size() < kParkMaxSets && held <= kParkMaxBytes) {
            g_parked.push_back(Parked{q->device, q->n_local, q->sm_count, q->measure_grid, q->amps, q->ring_host, q->ring_dev,
                                      q->partials, q->result_dev, q->result_host, q->pow_dev, q->total_dev, q->own_stream});
            lock.unlock();
            for (cudaEvent_t ev : q->pace_ev)
                if (ev) cudaEventDestroy(ev);
            if (q->cdf) cudaFree(q->cdf);
            if (q->stage) cudaFreeHost(q->stage);
            delete q;
            return QVM_OK;
        }
    }
    if (q->amps && q->owns_amps) cudaFree(q->amps);
    if (q->alt) cudaFree(q->alt);
    for (cudaEvent_t ev : q->pace_ev)
        if (ev) cudaEventDestroy(ev);
    if (q->ring_host) cudaFreeHost(q->ring_host);
    if (q->ring_dev) cudaFree(q->ring_dev);
    if (q->partials) cudaFree(q->partials);
    if (q->result_dev) cudaFree(q->result_dev);
    if (q->result_host) cudaFreeHost(q->result_host);
    if (q->pow_dev) cudaFree(q->pow_dev);
    if (q->total_dev) cudaFree(q->total_dev);
    if (q->cdf) cudaFree(q->cdf);
    if (q->stage) cudaFreeHost(q->stage);
    if (q->own_stream) cudaStreamDestroy(q->own_stream);
    delete q;
    return QVM_OK;
}

int qvm_set_shard(qvm_t* q, int n_global, uint64_t shard_index) {
    QVM_CHECK_HANDLE_LAZY(q);
    if (n_global < 0 || q->n_local + n_global > 62) return fail(QVM_ERR_INVALID, "bad n_global=%d", n_global);
    if (n_global < 64 && (shard_index >> n_global)) return fail(QVM_ERR_INVALID, "shard index out of range");
    q->n_global = n_global;
    q->shard_index = shard_index;
    return QVM_OK;
}

int qvm_set_stream(qvm_t* q, void* cuda_stream) {
    QVM_CHECK_HANDLE(q);
    QVM_CUDA(q, cudaStreamSynchronize(q->stream));
    q->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : q->own_stream;
    return QVM_OK;
}

int qvm_set_tile_bits(qvm_t* q, int tile_bits, int low_bits) {
    QVM_CHECK_HANDLE_LAZY(q);
    if (tile_bits < 1 || tile_bits > QVM_MAX_TILE_BITS || low_bits < 0 || low_bits > tile_bits)
        return fail(QVM_ERR_INVALID, "bad tile configuration (%d, %d)", tile_bits, low_bits);
    q->tile_bits = tile_bits;
    q->low_bits = low_bits;
    return QVM_OK;
}

int qvm_spec_compile_check(int n_local, int n_global, uint64_t shard_index, int n_tile, const int32_t* tile_bits,
                           int n_ops, const qvm_op_t* ops, const double* mats, uint64_t mats_entries, char* log,
                           uint64_t log_cap, int n_low, int n_bring, const int32_t* bring_low) {
    if (log && log_cap) log[0] = 0;
    if ((n_tile && !tile_bits) || (n_ops && !ops) || (n_bring && !bring_low)) return fail(QVM_ERR_INVALID, "null pointer");
    uint32_t low_in = 0;
    for (int i = 0; i < n_bring; ++i)
        for (int j = 0; j < n_tile; ++j)
            if (tile_bits[j] == bring_low[i]) low_in |= 1u << j;
    EncodedGroup g;
    SpecSource ss;
    if (encode_group(n_local, n_global, shard_index, n_tile, tile_bits, n_ops, ops, mats, mats_entries, g, 4, low_in,
                     n_low) != ENC_OK ||
        !generate_spec(g, ss))
        return 1;
    std::vector<char> cubin;
    std::string err;
    if (!SpecCache::instance().to_cubin(ss.src, cubin, &err)) {
        if (log && log_cap) snprintf(log, (size_t)log_cap, "%s", err.c_str());
        return fail(QVM_ERR_UNSUPPORTED, "specialised kernel did not compile: %.300s", err.c_str());
    }
    return QVM_OK;
}

namespace {
// qvm_jit_share: which of this thread's ahead-of-time submissions are compiled here
struct ShareState {
    int stride = 1, offset = 0;
    uint64_t index = 0;
};
thread_local ShareState t_share;
}  // namespace

int qvm_spec_precompile(int n_local, int n_global, uint64_t shard_index, int n_tile, const int32_t* tile_bits, int n_ops,
                        const qvm_op_t* ops, const double* mats, uint64_t mats_entries, int n_low, int n_bring,
                        const int32_t* bring_low, int32_t* new_pos) {
    if ((n_tile && !tile_bits) || (n_ops && !ops) || (n_bring && !bring_low) || !new_pos) return fail(QVM_ERR_INVALID, "null pointer");
    uint32_t low_in = 0;
    for (int i = 0; i < n_bring; ++i)
        for (int j = 0; j < n_tile; ++j)
            if (tile_bits[j] == bring_low[i]) low_in |= 1u << j;
    EncodedGroup g;
    SpecSource ss;
    const int erc = encode_group(n_local, n_global, shard_index, n_tile, tile_bits, n_ops, ops, mats, mats_entries, g, 4,
                                 low_in, n_low);
    for (int j = 0; j < n_tile; ++j) new_pos[j] = tile_bits[erc == ENC_OK ? g.sigma[j] : j];
    if (erc != ENC_OK || (g.cold.empty() && !g.relabelled) || g.need_full || !generate_spec(g, ss)) return 1;
    std::string err;      // compiled on the background threads: qvm_jit_wait() returns when all cubins are on disk
    const bool foreign = t_share.stride > 1 && (int)(t_share.index++ % (uint64_t)t_share.stride) != t_share.offset;
    if (!SpecCache::instance().precompile_async(ss.src, &err, false, foreign))
        return fail(QVM_ERR_UNSUPPORTED, "ahead-of-time compile failed: %.300s", err.c_str());
    return QVM_OK;
}

// Planning aid (no device, nothing compiled): the rounds the encoder plans for one group and where the relabelling
// leaves every tile position -- what a planner needs to compare candidate groups by cost.
int qvm_group_probe(int n_local, int n_global, uint64_t shard_index, int n_tile, const int32_t* tile_bits, int n_ops,
                    const qvm_op_t* ops, const double* mats, uint64_t mats_entries, int n_low, int n_bring,
                    const int32_t* bring_low, int32_t* new_pos, int32_t* info4) {
    if ((n_tile && !tile_bits) || (n_ops && !ops) || (n_bring && !bring_low) || !new_pos || !info4) return fail(QVM_ERR_INVALID, "null pointer");
    uint32_t low_in = 0;
    for (int i = 0; i < n_bring; ++i)
        for (int j = 0; j < n_tile; ++j)
            if (tile_bits[j] == bring_low[i]) low_in |= 1u << j;
    EncodedGroup g;
    const int erc = encode_group(n_local, n_global, shard_index, n_tile, tile_bits, n_ops, ops, mats, mats_entries, g, 4,
                                 low_in, n_low);
    for (int j = 0; j < n_tile; ++j) new_pos[j] = tile_bits[erc == ENC_OK ? g.sigma[j] : j];
    info4[0] = erc == ENC_OK ? (int)g.rounds.size() : -1;
    info4[1] = erc == ENC_OK ? g.n_gate_ops : 0;
    info4[2] = erc == ENC_OK ? g.n_dirty_rounds : 0;
    info4[3] = erc == ENC_OK && !g.need_full ? 1 : 0;
    return QVM_OK;
}

int qvm_jit_counters(uint64_t* out4) {
    if (!out4) return fail(QVM_ERR_INVALID, "null pointer");
    SpecCache::instance().counters(out4);
    return QVM_OK;
}

int qvm_jit_share(int stride, int offset) {
    if (stride < 1 || offset < 0 || offset >= stride) return fail(QVM_ERR_INVALID, "bad share %d/%d", offset, stride);
    t_share.stride = stride;
    t_share.offset = offset;
    t_share.index = 0;
    return QVM_OK;
}

int qvm_jit_defer(int n) {
    SpecCache::instance().defer_next(n);
    return QVM_OK;
}

int qvm_jit_cache_dir(char* out, uint64_t cap) {
    if (!out || !cap) return fail(QVM_ERR_INVALID, "null pointer");
    snprintf(out, (size_t)cap, "%s", SpecCache::instance().cache_dir().c_str());
    return QVM_OK;
}

int qvm_jit_wait(void) {
    SpecCache::instance().wait_idle();
    return QVM_OK;
}

int qvm_jit_shutdown(void) {
    SpecCache::instance().shutdown();
    return QVM_OK;
}

int qvm_set_jit(qvm_t* q, int policy) {
    QVM_CHECK_HANDLE_LAZY(q);
    if (policy < 0 || policy > 2) return fail(QVM_ERR_INVALID, "JIT policy must be 0, 1 or 2");
    q->jit_policy = policy;
    return QVM_OK;
}

int qvm_set_kernel_path(qvm_t* q, int use_regtile) {
    QVM_CHECK_HANDLE_LAZY(q);
    q->use_regtile = use_regtile ? 1 : 0;
    return QVM_OK;
}

int qvm_debug_opcode_hist(qvm_t* q, uint64_t* out32) {
    if (!q || !out32) return fail(QVM_ERR_INVALID, "null pointer");
    memcpy(out32, q->opc_hist, sizeof(q->opc_hist));
    return QVM_OK;
}

int qvm_n_local(const qvm_t* q) { return q ? q->n_local : -1; }
void* qvm_device_ptr(qvm_t* q) {
    if (q && q->zero_pending) materialize_zero(q);            // whoever gets the raw pointer sees real zeros
    return q ? q->amps : nullptr;
}
void* qvm_stream(qvm_t* q) { return q ? q->stream : nullptr; }

int qvm_sync(qvm_t* q) {
    QVM_CHECK_HANDLE(q);
    QVM_CUDA(q, cudaStreamSynchronize(q->stream));
    return QVM_OK;
}

int qvm_stats(const qvm_t* q, qvm_stats_t* out) {
    if (!q || !out) return fail(QVM_ERR_INVALID, "null pointer");
    *out = q->stats;
    return QVM_OK;
}

int qvm_stats_reset(qvm_t* q) {
    if (!q) return fail(QVM_ERR_INVALID, "null handle");
    q->stats = qvm_stats_t{};
    return QVM_OK;
}

// ---- state I/O ------------------------------------------------------------------------------------
int qvm_reset_zero(qvm_t* q, int set_amp0) {
    QVM_CHECK_HANDLE_LAZY(q);
    // lazy: noted down, and folded into the next fused gate pass (whose first round then reads nothing) or made
    // real by whatever touches the amplitudes first
    q->zero_pending = set_amp0 ? 2 : 1;
    q->sample_mode = -1;
    return QVM_OK;
}

static int copy_chunked(qvm_t* q, double2* dev, double* host, uint64_t count, bool to_device) {
    int rc = ensure_stage(q, std::min<size_t>(kStageChunk, std::max<size_t>(count * 16, 16)));
    if (rc) return rc;
    uint64_t done = 0;
    while (done < count) {
        const uint64_t n = std::min<uint64_t>(count - done, q->stage_bytes / 16);
        if (to_device) {
            memcpy(q->stage, host + 2 * done, n * 16);
            QVM_CUDA(q, cudaMemcpyAsync(dev + done, q->stage, n * 16, cudaMemcpyHostToDevice, q->stream));
        } else {
            QVM_CUDA(q, cudaMemcpyAsync(q->stage, dev + done, n * 16, cudaMemcpyDeviceToHost, q->stream));
        }
        QVM_CUDA(q, cudaStreamSynchronize(q->stream));
        if (!to_device) memcpy(host + 2 * done, q->stage, n * 16);
        done += n;
    }
    return QVM_OK;
}

int qvm_set_state(qvm_t* q, uint64_t offset, uint64_t count, const double* re_im) {
    QVM_CHECK_HANDLE(q);
    if (!re_im && count) return fail(QVM_ERR_INVALID, "null host buffer");
    if (offset > q->n_elems || count > q->n_elems - offset) return fail(QVM_ERR_INVALID, "range outside the shard");
    q->sample_mode = -1;
    q->stats.h2d_bytes += count * 16;
    return copy_chunked(q, q->amps + offset, const_cast<double*>(re_im), count, true);
}

int qvm_get_state(qvm_t* q, uint64_t offset, uint64_t count, double* re_im) {
    QVM_CHECK_HANDLE(q);
    if (!re_im && count) return fail(QVM_ERR_INVALID, "null host buffer");
    if (offset > q->n_elems || count > q->n_elems - offset) return fail(QVM_ERR_INVALID, "range outside the shard");
    q->stats.d2h_bytes += count * 16;
    return copy_chunked(q, q->amps + offset, re_im, count, false);
}

int qvm_get_strided(qvm_t* q, uint64_t offset, uint64_t stride, uint64_t count, double* re_im) {
    QVM_CHECK_HANDLE(q);
    if (!count) return QVM_OK;
    if (!re_im) return fail(QVM_ERR_INVALID, "null host buffer");
    if (offset >= q->n_elems || (count - 1) > (q->n_elems - 1 - offset) / (stride ? stride : 1))
        return fail(QVM_ERR_INVALID, "strided range outside the shard");
    double2* tmp = nullptr;
    QVM_CUDA(q, cudaMallocAsync(&tmp, count * 16, q->stream));
    gather_strided_kernel<<<grid_for(q, count, 256), 256, 0, q->stream>>>(q->amps, offset, stride, count, tmp);
    QVM_CUDA(q, cudaGetLastError());
    q->stats.kernel_launches += 1;
    int rc = copy_chunked(q, tmp, re_im, count, false);
    cudaFreeAsync(tmp, q->stream);
    q->stats.d2h_bytes += count * 16;
    return rc;
}

int qvm_get_state_layout(qvm_t* q, uint64_t offset, uint64_t stride, uint64_t count, int n_pos, const int32_t* pos_of,
                         double* re_im) {
    QVM_CHECK_HANDLE(q);
    if (!count) return QVM_OK;
    if (!re_im || !pos_of) return fail(QVM_ERR_INVALID, "null pointer");
    if (n_pos != q->n_local) return fail(QVM_ERR_INVALID, "layout of %d positions on a %d-position shard", n_pos, q->n_local);
    if (offset >= q->n_elems || (count - 1) > (q->n_elems - 1 - offset) / (stride ? stride : 1))
        return fail(QVM_ERR_INVALID, "range outside the shard");
    LayoutPerm perm{};
    perm.n = n_pos;
    uint64_t seen = 0;
    for (int i = 0; i < n_pos; ++i) {
        if (pos_of[i] < 0 || pos_of[i] >= n_pos || ((seen >> pos_of[i]) & 1ull))
            return fail(QVM_ERR_INVALID, "layout is not a permutation of the local positions");
        seen |= 1ull << pos_of[i];
        perm.pos[i] = (uint8_t)pos_of[i];
    }
    const uint64_t chunk = std::min<uint64_t>(count, (uint64_t)1 << 24);       // 256 MiB of staging at most
    double2* tmp = nullptr;
    QVM_CUDA(q, cudaMallocAsync(&tmp, chunk * 16, q->stream));
    int rc = QVM_OK;
    for (uint64_t done = 0; done < count && rc == QVM_OK; done += chunk) {
        const uint64_t n = std::min<uint64_t>(chunk, count - done);
        gather_layout_kernel<<<grid_for(q, n, 256), 256, 0, q->stream>>>(q->amps, offset + done * stride, stride, n, perm, tmp);
        if (cudaGetLastError() != cudaSuccess) rc = fail(QVM_ERR_CUDA, "gather_layout_kernel launch failed");
        q->stats.kernel_launches += 1;
        if (rc == QVM_OK) rc = copy_chunked(q, tmp, re_im + 2 * done, n, false);
    }
    cudaFreeAsync(tmp, q->stream);
    q->stats.d2h_bytes += count * 16;
    return rc;
}

int qvm_copy_state(qvm_t* dst, qvm_t* src) {
    QVM_CHECK_HANDLE(dst);
    if (!src) return fail(QVM_ERR_INVALID, "null source");
    if (dst->n_local != src->n_local || dst->device != src->device)
        return fail(QVM_ERR_INVALID, "copy needs equal shapes on one device");
    dst->zero_pending = 0;                           // overwritten anyway
    if (src->zero_pending) {
        const int rc0 = materialize_zero(src);
        if (rc0 != QVM_OK) return rc0;
    }
    QVM_CUDA(src, cudaStreamSynchronize(src->stream));
    QVM_CUDA(dst, cudaMemcpyAsync(dst->amps, src->amps, dst->n_elems * 16, cudaMemcpyDeviceToDevice, dst->stream));
    dst->sample_mode = -1;
    dst->stats.hbm_bytes += 32ull * dst->n_elems;
    return QVM_OK;
}

int qvm_norm2(qvm_t* q, double* out) {
    QVM_CHECK_HANDLE(q);
    if (!out) return fail(QVM_ERR_INVALID, "null pointer");
    const int grid = std::min(grid_for(q, q->n_elems, kReduceThreads), kPartialsCap);
    prob_partial_kernel<<<grid, kReduceThreads, 0, q->stream>>>(q->amps, q->n_elems, q->n_local, -1, 0, 0, 0, q->partials);
    QVM_CUDA(q, cudaGetLastError());
    q->stats.kernel_launches += 1;
    q->stats.hbm_bytes += 16ull * q->n_elems;
    return reduce_partials(q, grid, out);
}

// ---- gates ----------------------------------------------------------------------------------------
int qvm_classify(int k, const int32_t* qubits, const double* matrix, qvm_op_t* op, double* mat_out) {
    if (!qubits || !matrix || !op || !mat_out) return fail(QVM_ERR_INVALID, "null pointer");
    return classify(k, qubits, matrix, op, mat_out);
}

int qvm_apply(qvm_t* q, int k, const int32_t* qubits, const double* matrix) {
    QVM_CHECK_HANDLE_LAZY(q);
    if (!qubits || !matrix) return fail(QVM_ERR_INVALID, "null pointer");
    if (k < 1 || k > 10) return fail(QVM_ERR_INVALID, "gate size k=%d outside [1, 10]", k);
    int rc = validate_qubits(q, k, qubits);
    if (rc) return rc;
    qvm_op_t op{};
    std::vector<double> mat((size_t)2 << (2 * k));
    rc = classify(k, qubits, matrix, &op, mat.data());
    if (rc) return rc;
    if (op.kind == QVM_OP_NOP) return QVM_OK;
    // tile = dense targets + lowest positions, up to the configured tile size
    const int want = std::min(q->n_local, q->tile_bits);
    uint64_t mask = 0;
    int cnt = 0;
    if (op.kind == QVM_OP_DENSE)
        for (int j = 0; j < op.k; ++j) {
            if (op.targets[j] >= q->n_local)
                return fail(QVM_ERR_INVALID, "dense target %d is a global position; remap it to a local one first",
                            op.targets[j]);
            mask |= 1ull << op.targets[j];
            ++cnt;
        }
    for (int p = 0; p < q->n_local && cnt < want; ++p)
        if (!((mask >> p) & 1ull)) {
            mask |= 1ull << p;
            ++cnt;
        }
    int32_t tile[QVM_MAX_TILE_BITS + QVM_MAX_TARGETS];
    int nt = 0;
    for (int p = 0; p < q->n_local; ++p)
        if ((mask >> p) & 1ull) tile[nt++] = p;
    q->sample_mode = -1;
    return launch_group(q, nt, tile, 1, &op, mat.data(), op.mat_entries);
}

int qvm_apply_group(qvm_t* q, int n_tile, const int32_t* tile_bits, int n_ops, const qvm_op_t* ops, const double* mats,
                    uint64_t mats_entries) {
    QVM_CHECK_HANDLE_LAZY(q);
    if (n_ops < 0 || (n_ops && !ops) || (n_tile && !tile_bits)) return fail(QVM_ERR_INVALID, "null pointer");
    if (mats_entries && !mats) return fail(QVM_ERR_INVALID, "null matrix pool");
    q->sample_mode = -1;
    return launch_group(q, n_tile, tile_bits, n_ops, ops, mats, mats_entries);
}

int qvm_apply_group_relabel(qvm_t* q, int n_tile, const int32_t* tile_bits, int n_ops, const qvm_op_t* ops,
                            const double* mats, uint64_t mats_entries, int n_low, int n_bring, const int32_t* bring_low,
                            int32_t* new_pos) {
    QVM_CHECK_HANDLE_LAZY(q);
    if (n_ops < 0 || (n_ops && !ops) || (n_tile && !tile_bits) || !new_pos || (n_bring && !bring_low))
        return fail(QVM_ERR_INVALID, "null pointer");
    if (mats_entries && !mats) return fail(QVM_ERR_INVALID, "null matrix pool");
    if (n_tile < 0 || n_tile > 16 || n_bring < 0 || n_low < 0) return fail(QVM_ERR_INVALID, "bad relabelling request");
    RelabelReq req;
    req.n_low = n_low;
    for (int i = 0; i < n_bring; ++i) {
        int j = 0;
        while (j < n_tile && tile_bits[j] != bring_low[i]) ++j;
        if (j == n_tile) return fail(QVM_ERR_INVALID, "position %d is not in the tile", bring_low[i]);
        req.low_in |= 1u << j;
    }
    q->sample_mode = -1;
    const int rc = launch_group(q, n_tile, tile_bits, n_ops, ops, mats, mats_entries, &req);
    if (rc != QVM_OK) return rc;
    for (int j = 0; j < n_tile; ++j) new_pos[j] = tile_bits[req.sigma[j]];
    return QVM_OK;
}

// ---- compiled programs: record the launches of a flush once, replay them on every later run ----------
int qvm_record_begin(qvm_t* q) {
    QVM_CHECK_HANDLE_LAZY(q);
    if (q->recording) return fail(QVM_ERR_INVALID, "a recording is already open on this handle");
    qvm_plan* plan = new (std::nothrow) qvm_plan();
    if (!plan) return fail(QVM_ERR_NOMEM, "host allocation failed");
    plan->n_local = q->n_local;
    plan->n_global = q->n_global;
    plan->shard_index = q->shard_index;
    q->recording = plan;
    return QVM_OK;
}

int qvm_record_end(qvm_t* q, qvm_plan_t** plan) {
    if (!q) return fail(QVM_ERR_INVALID, "null handle");
    if (!q->recording) return fail(QVM_ERR_INVALID, "no recording is open on this handle");
    qvm_plan* rec = q->recording;
    q->recording = nullptr;
    if (plan) *plan = rec;
    else delete rec;                       // aborted
    return QVM_OK;
}

int qvm_plan_launches(const qvm_plan_t* plan) { return plan ? (int)plan->recs.size() : -1; }

int qvm_plan_destroy(qvm_plan_t* plan) {
    if (!plan) return QVM_OK;
    if (plan->graph) cudaGraphExecDestroy(plan->graph);
    delete plan;
    return QVM_OK;
}

static int plan_launch_all(qvm_t* q, qvm_plan_t* plan, int reset_first, bool* all_specialised) {
    if (reset_first) q->zero_pending = 2;            // folded into the first launch below
    bool all = true;
    for (LaunchRec& rec : plan->recs) {
        const int rc = launch_rec(q, rec);
        if (rc != QVM_OK) return rc;
        all = all && rec.kind == REC_REGTILE && rec.cs != nullptr && q->jit_policy > 0;
    }
    if (q->zero_pending) {                           // a plan without launches: the reset is all there is
        const int rc = materialize_zero(q);
        if (rc != QVM_OK) return rc;
    }
    if (all_specialised) *all_specialised = all;
    return QVM_OK;
}

int qvm_plan_run(qvm_t* q, qvm_plan_t* plan, int reset_first) {
    QVM_CHECK_HANDLE_LAZY(q);
    if (!plan) return fail(QVM_ERR_INVALID, "null plan");
    if (q->recording) return fail(QVM_ERR_INVALID, "a recording is open on this handle");
    if (plan->n_local != q->n_local || plan->n_global != q->n_global || plan->shard_index != q->shard_index)
        return fail(QVM_ERR_INVALID, "plan was compiled for another shard shape");
    q->sample_mode = -1;
    ++plan->runs;
    if (!reset_first && q->zero_pending) {           // (a replay on top of a reset somebody else asked for)
        const int rc0 = materialize_zero(q);
        if (rc0 != QVM_OK) return rc0;
    }
    if (plan->graph && plan->graph_amps == q->amps && plan->graph_reset == reset_first) {
        QVM_CUDA(q, cudaGraphLaunch(plan->graph, q->stream));
        q->zero_pending = 0;                         // the captured first pass carries the reset
        q->stats.kernel_launches += plan->graph_stats.kernel_launches;
        q->stats.jit_launches += plan->graph_stats.jit_launches;
        q->stats.gate_ops += plan->graph_stats.gate_ops;
        q->stats.hbm_passes += plan->graph_stats.hbm_passes;
        q->stats.hbm_bytes += plan->graph_stats.hbm_bytes;
        q->stats.regtile_rounds += plan->graph_stats.regtile_rounds;
        q->stats.zero_fused_launches += plan->graph_stats.zero_fused_launches;
        q->stats.graph_launches += 1;
        return QVM_OK;
    }
    // Every group has its specialised kernel (no descriptor uploads left) and the program came back: capture the
    // whole sequence once -- launch-bound small registers then cost one cudaGraphLaunch per run.
    static const bool graphs_on = [] {
        const char* e = getenv("QVM_B200_GRAPHS");
        return !e || atoi(e) != 0;
    }();
    if (graphs_on && plan->runs >= 3 && plan->recs.size() + (reset_first ? 1 : 0) >= 2 && plan->all_specialised &&
        !plan->graph_failed) {
        if (plan->graph) {
            cudaGraphExecDestroy(plan->graph);
            plan->graph = nullptr;
        }
        const qvm_stats_t before = q->stats;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(q->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            bool all = false;
            q->capturing = true;
            const int rc = plan_launch_all(q, plan, reset_first, &all);
            q->capturing = false;
            ok = cudaStreamEndCapture(q->stream, &graph) == cudaSuccess && rc == QVM_OK && all && graph;
        }
        if (ok) ok = cudaGraphInstantiate(&plan->graph, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        const qvm_stats_t after = q->stats;
        q->stats = before;
        if (ok) {
            plan->graph_amps = q->amps;
            plan->graph_reset = reset_first;
            plan->graph_stats = qvm_stats_t{};
            plan->graph_stats.kernel_launches = after.kernel_launches - before.kernel_launches;
            plan->graph_stats.jit_launches = after.jit_launches - before.jit_launches;
            plan->graph_stats.gate_ops = after.gate_ops - before.gate_ops;
            plan->graph_stats.hbm_passes = after.hbm_passes - before.hbm_passes;
            plan->graph_stats.hbm_bytes = after.hbm_bytes - before.hbm_bytes;
            plan->graph_stats.regtile_rounds = after.regtile_rounds - before.regtile_rounds;
            plan->graph_stats.zero_fused_launches = after.zero_fused_launches - before.zero_fused_launches;
            return qvm_plan_run(q, plan, reset_first);        // (takes the graph branch above)
        }
        cudaGetLastError();
        plan->graph = nullptr;
        plan->graph_failed = true;
        q->sticky = 0;
    }
    bool all = false;
    const int rc = plan_launch_all(q, plan, reset_first, &all);
    plan->all_specialised = all;
    return rc;
}

int qvm_dephase_all(qvm_t* q, int n_rho, uint64_t qubit_mask, double factor) {
    QVM_CHECK_HANDLE(q);
    if (2 * n_rho != total_positions(q) || n_rho > 31) return fail(QVM_ERR_INVALID, "state is not a %d-qubit rho", n_rho);
    double table[65];
    table[0] = 1.0;
    for (int i = 1; i <= 64; ++i) table[i] = table[i - 1] * factor;
    size_t off = 0;
    int rc = ring_alloc(q, sizeof(table), &off);
    if (rc) return rc;
    memcpy(q->ring_host + off, table, sizeof(table));
    QVM_CUDA(q, cudaMemcpyAsync(q->ring_dev + off, q->ring_host + off, sizeof(table), cudaMemcpyHostToDevice, q->stream));
    dephase_kernel<<<grid_for(q, q->n_elems, 256, 16), 256, 0, q->stream>>>(
        q->amps, q->n_elems, n_rho, qubit_mask, q->shard_index << q->n_local,
        reinterpret_cast<const double*>(q->ring_dev + off));
    QVM_CUDA(q, cudaGetLastError());
    q->sample_mode = -1;
    q->stats.kernel_launches += 1;
    q->stats.hbm_passes += 1;
    q->stats.hbm_bytes += 32ull * q->n_elems;
    return QVM_OK;
}

// ---- MEASURE ----------------------------------------------------------------------------------------
int qvm_measure(qvm_t* q, int qubit, double r, int* outcome, double* p0) {
    QVM_CHECK_HANDLE(q);
    if (!outcome || !p0) return fail(QVM_ERR_INVALID, "null pointer");
    if (q->n_global) return fail(QVM_ERR_INVALID, "qvm_measure is single-shard; use qvm_prob_zero + qvm_collapse");
    if (qubit < 0 || qubit >= q->n_local) return fail(QVM_ERR_INVALID, "qubit %d out of range", qubit);
    void* args[] = {&q->amps, &q->n_elems, &q->n_local, &qubit, &r, &q->partials, &q->result_dev};
    QVM_CUDA(q, cudaLaunchCooperativeKernel((void*)measure_kernel, dim3(q->measure_grid), dim3(kReduceThreads), args, 0,
                                            q->stream));
    QVM_CUDA(q, cudaMemcpyAsync(q->result_host, q->result_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, q->stream));
    QVM_CUDA(q, cudaStreamSynchronize(q->stream));
    *outcome = (int)q->result_host[0];
    *p0 = q->result_host[1];
    q->sample_mode = -1;
    q->stats.kernel_launches += 1;
    q->stats.hbm_bytes += 16ull * q->n_elems + 24ull * q->n_elems;
    q->stats.d2h_bytes += 16;
    return QVM_OK;
}

// ---- runs of MEASURE instructions -----------------------------------------------------------------------
int qvm_measure_hist(qvm_t* q, int n_high, const int32_t* high_pos, int do_collapse, uint64_t keep_mask,
                     uint64_t keep_value, double scale, double* hist) {
    QVM_CHECK_HANDLE(q);
    if (q->n_local < 5) return fail(QVM_ERR_INVALID, "qvm_measure_hist needs at least 5 local positions");
    if (n_high < 0 || n_high > 5 || (n_high && !high_pos)) return fail(QVM_ERR_INVALID, "0..5 high positions");
    HistBits hb{};
    hb.n_high = n_high;
    for (int i = 0; i < n_high; ++i) {
        if (high_pos[i] < 5 || high_pos[i] >= q->n_local) return fail(QVM_ERR_INVALID, "high position %d out of range", high_pos[i]);
        hb.pos[i] = (uint8_t)high_pos[i];
    }
    if (keep_value & ~keep_mask) return fail(QVM_ERR_INVALID, "collapse value outside its mask");
    if (do_collapse < 0 || do_collapse > 2) return fail(QVM_ERR_INVALID, "do_collapse is 0, 1 or 2");
    const uint64_t all_local = q->n_elems - 1;
    if (keep_mask & ~all_local) return fail(QVM_ERR_INVALID, "collapse mask outside the shard");
    if (do_collapse == 2 && !hist) return fail(QVM_ERR_INVALID, "filter mode returns a histogram");
    const int n_entries = 32 << n_high;
    // final collapse of a run that measured every local position: one survivor, write-only
    if (do_collapse == 1 && !hist && keep_mask == all_local && q->n_elems > 1) {
        double2* keep = reinterpret_cast<double2*>(q->result_dev);
        if (scale != 0.0) survivor_fetch_kernel<<<1, 1, 0, q->stream>>>(q->amps, keep_value, scale, keep);
        QVM_CUDA(q, cudaGetLastError());
        QVM_CUDA(q, cudaMemsetAsync(q->amps, 0, q->n_elems * sizeof(double2), q->stream));
        if (scale != 0.0) survivor_store_kernel<<<1, 1, 0, q->stream>>>(q->amps, keep_value, keep);
        QVM_CUDA(q, cudaGetLastError());
        q->sample_mode = -1;
        q->stats.kernel_launches += scale != 0.0 ? 2 : 0;
        q->stats.hbm_bytes += 16ull * q->n_elems;
        return QVM_OK;
    }
    if (do_collapse == 2 && scale == 0.0) {          // (a shard whose global bits lost: nothing consistent here)
        memset(hist, 0, n_entries * sizeof(double));
        return QVM_OK;
    }
    const bool filter = do_collapse == 2 && keep_mask != 0;
    FilterBits fb{};
    int n_free = q->n_local;
    if (filter) {
        n_free = 0;
        for (int b = 0; b < q->n_local;) {
            if ((keep_mask >> b) & 1ull) {
                ++b;
                continue;
            }
            int e = b;
            while (e < q->n_local && !((keep_mask >> e) & 1ull)) ++e;
            if (fb.n_runs == 20) return fail(QVM_ERR_UNSUPPORTED, "collapse mask too fragmented for the filtered histogram");
            fb.start[fb.n_runs] = (uint8_t)b;
            fb.len[fb.n_runs] = (uint8_t)(e - b);
            ++fb.n_runs;
            n_free += e - b;
            b = e;
        }
    }
    const uint64_t n_items = filter ? (((1ull << n_free) + kHistThreads - 1) / kHistThreads)
                                    : (((q->n_elems >> 5) + (kHistThreads / 32) - 1) / (kHistThreads / 32));
    int grid = (int)std::min<uint64_t>(n_items, (uint64_t)q->sm_count * 8);
    if (grid < 1) grid = 1;
    double* dev = nullptr;
    QVM_CUDA(q, cudaMallocAsync(&dev, ((size_t)grid + 1) * n_entries * sizeof(double), q->stream));
    double* d_hist = dev + (size_t)grid * n_entries;
    int rc = QVM_OK;
    if (filter)
        measure_hist_filter_kernel<<<grid, kHistThreads, 0, q->stream>>>(q->amps, n_free, hb, fb, keep_value, dev);
    else
        measure_hist_kernel<<<grid, kHistThreads, 0, q->stream>>>(q->amps, q->n_elems, hb, do_collapse == 1 ? 1 : 0, keep_mask,
                                                                  keep_value, scale, dev);
    if (cudaGetLastError() != cudaSuccess) rc = fail(QVM_ERR_CUDA, "measure_hist_kernel launch failed");
    if (rc == QVM_OK && hist) {
        measure_hist_final_kernel<<<(n_entries + 255) / 256, 256, 0, q->stream>>>(dev, grid, n_entries, d_hist);
        std::vector<double> host(n_entries);
        if (cudaGetLastError() != cudaSuccess ||
            cudaMemcpyAsync(host.data(), d_hist, n_entries * sizeof(double), cudaMemcpyDeviceToHost, q->stream) != cudaSuccess ||
            cudaStreamSynchronize(q->stream) != cudaSuccess)
            rc = fail(QVM_ERR_CUDA, "histogram readback failed");
        else {
            memcpy(hist, host.data(), n_entries * sizeof(double));
            if (do_collapse == 2)                    // the projection is not applied yet: scale the probabilities instead
                for (int j = 0; j < n_entries; ++j) hist[j] *= scale * scale;
        }
        q->stats.kernel_launches += 1;
        q->stats.d2h_bytes += n_entries * sizeof(double);
    }
    cudaFreeAsync(dev, q->stream);
    if (rc != QVM_OK) return rc;
    if (do_collapse == 1) q->sample_mode = -1;
    q->stats.kernel_launches += 1;
    q->stats.hbm_bytes += filter ? (16ull << n_free) : (do_collapse == 1 ? 32ull : 16ull) * q->n_elems;
    return QVM_OK;
}

// ---- batched trials ------------------------------------------------------------------------------------
int qvm_reset_batched(qvm_t* q, int n_per_trial) {
    QVM_CHECK_HANDLE(q);
    if (q->n_global) return fail(QVM_ERR_INVALID, "batched trials are single-shard");
    if (n_per_trial < 1 || n_per_trial > q->n_local) return fail(QVM_ERR_INVALID, "bad qubits per trial %d", n_per_trial);
    batched_reset_kernel<<<grid_for(q, q->n_elems, 256, 4), 256, 0, q->stream>>>(q->amps, q->n_elems, n_per_trial);
    QVM_CUDA(q, cudaGetLastError());
    q->sample_mode = -1;
    q->stats.kernel_launches += 1;
    q->stats.hbm_bytes += 16ull * q->n_elems;
    return QVM_OK;
}

int qvm_gather_segments(qvm_t* dst, qvm_t* src, int n_per_trial, uint64_t count, const uint32_t* src_index) {
    QVM_CHECK_HANDLE(dst);
    if (!src || (count && !src_index)) return fail(QVM_ERR_INVALID, "null pointer");
    if (dst->device != src->device) return fail(QVM_ERR_INVALID, "segments are gathered on one device");
    if (n_per_trial < 1 || n_per_trial > dst->n_local || n_per_trial > src->n_local)
        return fail(QVM_ERR_INVALID, "bad qubits per trial %d", n_per_trial);
    const uint64_t dst_segs = 1ull << (dst->n_local - n_per_trial), src_segs = 1ull << (src->n_local - n_per_trial);
    if (count > dst_segs || count > (1ull << 20)) return fail(QVM_ERR_INVALID, "%llu segments do not fit", (unsigned long long)count);
    for (uint64_t j = 0; j < count; ++j)
        if (src_index[j] >= src_segs) return fail(QVM_ERR_INVALID, "source segment %u out of range", src_index[j]);
    if (!count) return QVM_OK;
    if (src->zero_pending) {
        const int rc0 = materialize_zero(src);
        if (rc0 != QVM_OK) return rc0;
    }
    QVM_CUDA(src, cudaStreamSynchronize(src->stream));
    uint32_t* dev = nullptr;
    QVM_CUDA(dst, cudaMallocAsync(&dev, count * sizeof(uint32_t), dst->stream));
    cudaError_t e = cudaMemcpyAsync(dev, src_index, count * sizeof(uint32_t), cudaMemcpyHostToDevice, dst->stream);
    if (e == cudaSuccess) {
        const int threads = n_per_trial >= 8 ? 256 : (n_per_trial >= 5 ? 32 : 32);
        gather_segments_kernel<<<(unsigned)count, threads, 0, dst->stream>>>(dst->amps, src->amps, n_per_trial, dev);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(dst->stream);       // src_index is the caller's (pageable) memory
    cudaFreeAsync(dev, dst->stream);
    QVM_CUDA(dst, e);
    dst->sample_mode = -1;
    dst->stats.kernel_launches += 1;
    dst->stats.hbm_bytes += (32ull << n_per_trial) * count;
    dst->stats.h2d_bytes += count * sizeof(uint32_t);
    return QVM_OK;
}

int qvm_measure_batched(qvm_t* q, int n_per_trial, int qubit, const double* uniforms, int32_t* outcomes, double* p0) {
    QVM_CHECK_HANDLE(q);
    if (!uniforms || !outcomes) return fail(QVM_ERR_INVALID, "null pointer");
    if (q->n_global) return fail(QVM_ERR_INVALID, "batched trials are single-shard");
    if (n_per_trial < 1 || n_per_trial > q->n_local) return fail(QVM_ERR_INVALID, "bad qubits per trial %d", n_per_trial);
    if (qubit < 0 || qubit >= n_per_trial) return fail(QVM_ERR_INVALID, "qubit %d out of range", qubit);
    const uint64_t segs = 1ull << (q->n_local - n_per_trial);
    if (segs > (1ull << 20)) return fail(QVM_ERR_UNSUPPORTED, "too many segments");
    const size_t bytes = (size_t)segs * (sizeof(double) * 2 + sizeof(int32_t));
    unsigned char* dev = nullptr;
    QVM_CUDA(q, cudaMallocAsync(&dev, bytes, q->stream));
    double* d_r = reinterpret_cast<double*>(dev);
    double* d_p0 = d_r + segs;
    int32_t* d_out = reinterpret_cast<int32_t*>(d_p0 + segs);
    int rc = QVM_OK;
    std::vector<double> host_p0(segs);
    if (cudaMemcpyAsync(d_r, uniforms, segs * sizeof(double), cudaMemcpyHostToDevice, q->stream) != cudaSuccess)
        rc = fail(QVM_ERR_CUDA, "upload of the uniforms failed");
    if (rc == QVM_OK) {
        const int threads = n_per_trial >= 12 ? 256 : (n_per_trial >= 8 ? 64 : 32);
        batched_measure_kernel<<<(unsigned)segs, threads, 0, q->stream>>>(q->amps, n_per_trial, qubit, d_r, d_out, d_p0);
        if (cudaGetLastError() != cudaSuccess) rc = fail(QVM_ERR_CUDA, "batched_measure_kernel launch failed");
    }
    if (rc == QVM_OK && (cudaMemcpyAsync(outcomes, d_out, segs * sizeof(int32_t), cudaMemcpyDeviceToHost, q->stream) != cudaSuccess ||
                         cudaMemcpyAsync(host_p0.data(), d_p0, segs * sizeof(double), cudaMemcpyDeviceToHost, q->stream) != cudaSuccess ||
                         cudaStreamSynchronize(q->stream) != cudaSuccess))
        rc = fail(QVM_ERR_CUDA, "readback of the outcomes failed");
    cudaFreeAsync(dev, q->stream);
    if (rc != QVM_OK) return rc;
    if (p0) memcpy(p0, host_p0.data(), segs * sizeof(double));
    q->sample_mode = -1;
    q->stats.kernel_launches += 1;
    q->stats.hbm_bytes += 40ull * q->n_elems;
    q->stats.h2d_bytes += segs * sizeof(double);
    q->stats.d2h_bytes += segs * (sizeof(double) + sizeof(int32_t));
    return QVM_OK;
}

int qvm_prob_zero(qvm_t* q, int qubit, double* p0_partial) {
    QVM_CHECK_HANDLE(q);
    if (!p0_partial) return fail(QVM_ERR_INVALID, "null pointer");
    if (qubit < 0 || qubit >= total_positions(q)) return fail(QVM_ERR_INVALID, "qubit %d out of range", qubit);
    if (qubit >= q->n_local && ((q->shard_index >> (qubit - q->n_local)) & 1ull)) {
        *p0_partial = 0.0;          // this shard holds only the |1> half of a global qubit
        return QVM_OK;
    }
    const int grid = std::min(grid_for(q, q->n_elems, kReduceThreads), kPartialsCap);
    prob_partial_kernel<<<grid, kReduceThreads, 0, q->stream>>>(q->amps, q->n_elems, q->n_local, qubit, 0, 0, 0,
                                                                q->partials);
    QVM_CUDA(q, cudaGetLastError());
    q->stats.kernel_launches += 1;
    q->stats.hbm_bytes += 16ull * q->n_elems;
    return reduce_partials(q, grid, p0_partial);
}

int qvm_density_prob_zero(qvm_t* q, int n_rho, int qubit, double* p0_partial) {
    QVM_CHECK_HANDLE(q);
    if (!p0_partial) return fail(QVM_ERR_INVALID, "null pointer");
    if (2 * n_rho != total_positions(q) || q->n_local < n_rho)
        return fail(QVM_ERR_INVALID, "state is not a %d-qubit rho sharded by rows", n_rho);
    if (qubit < -1 || qubit >= n_rho) return fail(QVM_ERR_INVALID, "qubit %d out of range", qubit);
    const uint64_t rows = q->n_elems >> n_rho;
    const int grid = std::min(grid_for(q, rows, kReduceThreads), kPartialsCap);
    prob_partial_kernel<<<grid, kReduceThreads, 0, q->stream>>>(q->amps, q->n_elems, q->n_local, qubit, 1, n_rho,
                                                                q->shard_index << q->n_local, q->partials);
    QVM_CUDA(q, cudaGetLastError());
    q->stats.kernel_launches += 1;
    return reduce_partials(q, grid, p0_partial);
}

int qvm_density_diag(qvm_t* q, int n_rho, const int32_t* pos_of, double* diag_host) {
    QVM_CHECK_HANDLE(q);
    if (!pos_of || !diag_host) return fail(QVM_ERR_INVALID, "null pointer");
    if (n_rho < 1 || n_rho > 31 || 2 * n_rho != total_positions(q))
        return fail(QVM_ERR_INVALID, "state is not a %d-qubit rho", n_rho);
    DiagLayout lay{};
    uint64_t seen = 0;
    for (int v = 0; v < 2 * n_rho; ++v) {
        const int p = pos_of[v];
        if (p < 0 || p >= 2 * n_rho || ((seen >> p) & 1ull)) return fail(QVM_ERR_INVALID, "layout is not a permutation");
        seen |= 1ull << p;
        (v < n_rho ? lay.pos_col[v] : lay.pos_row[v - n_rho]) = (uint8_t)p;
    }
    const uint64_t n = 1ull << n_rho;
    double* dev = nullptr;
    QVM_CUDA(q, cudaMallocAsync(&dev, n * sizeof(double), q->stream));
    density_diag_kernel<<<grid_for(q, n, 256), 256, 0, q->stream>>>(q->amps, n_rho, lay, q->n_local, q->shard_index, dev);
    cudaError_t e = cudaGetLastError();
    int rc = QVM_OK;
    if (e == cudaSuccess) {
        rc = ensure_stage(q, std::min<size_t>(kStageChunk, n * sizeof(double)));
        for (uint64_t done = 0; rc == QVM_OK && e == cudaSuccess && done < n;) {
            const uint64_t cnt = std::min<uint64_t>(n - done, q->stage_bytes / sizeof(double));
            e = cudaMemcpyAsync(q->stage, dev + done, cnt * sizeof(double), cudaMemcpyDeviceToHost, q->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(q->stream);
            if (e == cudaSuccess) memcpy(diag_host + done, q->stage, cnt * sizeof(double));
            done += cnt;
        }
    }
    cudaFreeAsync(dev, q->stream);
    if (rc != QVM_OK) return rc;
    QVM_CUDA(q, e);
    q->stats.kernel_launches += 1;
    q->stats.d2h_bytes += n * sizeof(double);
    return QVM_OK;
}

int qvm_density_offdiag(qvm_t* q, int n_rho, const int32_t* pos_of, uint64_t col_xor, double* re_im_host) {
    QVM_CHECK_HANDLE(q);
    if (!pos_of || !re_im_host) return fail(QVM_ERR_INVALID, "null pointer");
    if (n_rho < 1 || n_rho > 31 || 2 * n_rho != total_positions(q))
        return fail(QVM_ERR_INVALID, "state is not a %d-qubit rho", n_rho);
    if (col_xor >> n_rho) return fail(QVM_ERR_INVALID, "column mask outside the %d-qubit register", n_rho);
    DiagLayout lay{};
    uint64_t seen = 0;
    for (int v = 0; v < 2 * n_rho; ++v) {
        const int p = pos_of[v];
        if (p < 0 || p >= 2 * n_rho || ((seen >> p) & 1ull)) return fail(QVM_ERR_INVALID, "layout is not a permutation");
        seen |= 1ull << p;
        (v < n_rho ? lay.pos_col[v] : lay.pos_row[v - n_rho]) = (uint8_t)p;
    }
    const uint64_t n = 1ull << n_rho;
    double2* dev = nullptr;
    QVM_CUDA(q, cudaMallocAsync(&dev, n * sizeof(double2), q->stream));
    density_offdiag_kernel<<<grid_for(q, n, 256), 256, 0, q->stream>>>(q->amps, n_rho, lay, q->n_local, q->shard_index,
                                                                      col_xor, dev);
    int rc = cudaGetLastError() == cudaSuccess ? QVM_OK : fail(QVM_ERR_CUDA, "density_offdiag_kernel launch failed");
    if (rc == QVM_OK) rc = copy_chunked(q, dev, re_im_host, n, false);
    cudaFreeAsync(dev, q->stream);
    if (rc != QVM_OK) return rc;
    q->stats.kernel_launches += 1;
    q->stats.d2h_bytes += n * sizeof(double2);
    return QVM_OK;
}

static int reduce_partials2(qvm_state* q, int n_blocks, double* re_im) {
    final_sum2_kernel<<<1, kReduceThreads, 0, q->stream>>>(q->partials, n_blocks, q->result_dev);
    QVM_CUDA(q, cudaGetLastError());
    QVM_CUDA(q, cudaMemcpyAsync(q->result_host, q->result_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, q->stream));
    QVM_CUDA(q, cudaStreamSynchronize(q->stream));
    re_im[0] = q->result_host[0];
    re_im[1] = q->result_host[1];
    q->stats.kernel_launches += 2;
    q->stats.d2h_bytes += 2 * sizeof(double);
    return QVM_OK;
}

int qvm_pauli_expectation(qvm_t* q, uint64_t xmask, uint64_t zmask, double* re_im_partial) {
    QVM_CHECK_HANDLE(q);
    if (!re_im_partial) return fail(QVM_ERR_INVALID, "null pointer");
    const int npos = total_positions(q);
    if (xmask >> q->n_local) return fail(QVM_ERR_INVALID, "X / Y factors must sit on local positions (remap them first)");
    if (npos < 64 && (zmask >> npos)) return fail(QVM_ERR_INVALID, "Z mask outside the register");
    const int grid = std::min(grid_for(q, q->n_elems, kReduceThreads), kPartialsCap / 2);
    pauli_partial_kernel<<<grid, kReduceThreads, 0, q->stream>>>(q->amps, q->n_elems, xmask, zmask,
                                                                 q->shard_index << q->n_local, q->partials);
    QVM_CUDA(q, cudaGetLastError());
    q->stats.hbm_bytes += 32ull * q->n_elems;
    return reduce_partials2(q, grid, re_im_partial);
}

int qvm_inner_product(qvm_t* a, qvm_t* b, double* re_im_partial) {
    QVM_CHECK_HANDLE(a);
    if (!b || !re_im_partial) return fail(QVM_ERR_INVALID, "null pointer");
    if (a->n_local != b->n_local || a->device != b->device) return fail(QVM_ERR_INVALID, "shards of different shape");
    if (b->zero_pending) {
        const int rc0 = materialize_zero(b);
        if (rc0 != QVM_OK) return rc0;
    }
    QVM_CUDA(b, cudaStreamSynchronize(b->stream));
    const int grid = std::min(grid_for(a, a->n_elems, kReduceThreads), kPartialsCap / 2);
    inner_partial_kernel<<<grid, kReduceThreads, 0, a->stream>>>(a->amps, b->amps, a->n_elems, a->partials);
    QVM_CUDA(a, cudaGetLastError());
    a->stats.hbm_bytes += 32ull * a->n_elems;
    return reduce_partials2(a, grid, re_im_partial);
}

int qvm_collapse(qvm_t* q, int qubit, int outcome, double scale) {
    QVM_CHECK_HANDLE(q);
    if (qubit < 0 || qubit >= total_positions(q)) return fail(QVM_ERR_INVALID, "qubit %d out of range", qubit);
    if (outcome != 0 && outcome != 1) return fail(QVM_ERR_INVALID, "outcome must be 0 or 1");
    double s = scale;
    if (qubit >= q->n_local) {
        const int mine = (int)((q->shard_index >> (qubit - q->n_local)) & 1ull);
        if (mine != outcome) s = -1.0;      // zero the shard
    }
    collapse_kernel<<<grid_for(q, q->n_elems, 256, 16), 256, 0, q->stream>>>(q->amps, q->n_elems, q->n_local, qubit,
                                                                             outcome, s);
    QVM_CUDA(q, cudaGetLastError());
    q->sample_mode = -1;
    q->stats.kernel_launches += 1;
    q->stats.hbm_bytes += 24ull * q->n_elems;
    return QVM_OK;
}

// ---- sampling ---------------------------------------------------------------------------------------
int qvm_probs_prepare(qvm_t* q, int mode, int n_rho, double* total) {
    QVM_CHECK_HANDLE(q);
    if (mode < 0 || mode > 2) return fail(QVM_ERR_INVALID, "mode must be 0 (|psi|^2), 1 (diag rho) or 2 (plain probabilities)");
    uint64_t n_probs = q->n_elems;
    uint64_t row_base = 0;
    if (mode == 2) n_probs = 2 * q->n_elems;          // the buffer read as doubles
    if (mode == 1) {
        if (2 * n_rho != total_positions(q) || q->n_local < n_rho)
            return fail(QVM_ERR_INVALID, "state is not a %d-qubit rho sharded by rows", n_rho);
        n_probs = q->n_elems >> n_rho;
        row_base = (q->shard_index << q->n_local) >> n_rho;
    }
    const uint64_t n_blocks = (n_probs + (1ull << kSampleBlockBits) - 1) >> kSampleBlockBits;
    if (n_blocks > q->cdf_cap) {
        if (q->cdf) cudaFree(q->cdf);
        q->cdf = nullptr;
        q->cdf_cap = 0;
        if (cudaMalloc(&q->cdf, n_blocks * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            parked_release_all();
            cudaSetDevice(q->device);
            QVM_CUDA(q, cudaMalloc(&q->cdf, n_blocks * sizeof(double)));
        }
        q->cdf_cap = n_blocks;
    }
    const int threads = 256;
    const int grid = grid_for(q, n_blocks * 32, threads, 16);
    sample_block_sums_kernel<<<grid, threads, 0, q->stream>>>(q->amps, n_probs, mode, n_rho, row_base, q->cdf, n_blocks);
    QVM_CUDA(q, cudaGetLastError());
    if (n_blocks <= (uint64_t)kScanChunk) {
        scan_inclusive_kernel<<<1, 1024, 0, q->stream>>>(q->cdf, n_blocks, q->total_dev);
        QVM_CUDA(q, cudaGetLastError());
    } else {
        // long CDFs (>= 24 qubits): chunk totals -> scan of the totals -> every chunk scans itself
        const uint64_t n_chunks = (n_blocks + kScanChunk - 1) / kScanChunk;
        double* chunk_tot = nullptr;
        QVM_CUDA(q, cudaMallocAsync(&chunk_tot, n_chunks * sizeof(double), q->stream));
        scan_chunk_totals_kernel<<<(unsigned)n_chunks, kScanThreads, 0, q->stream>>>(q->cdf, n_blocks, chunk_tot);
        scan_inclusive_kernel<<<1, 1024, 0, q->stream>>>(chunk_tot, n_chunks, q->total_dev);
        scan_chunks_kernel<<<(unsigned)n_chunks, kScanThreads, 0, q->stream>>>(q->cdf, n_blocks, chunk_tot);
        const cudaError_t e = cudaGetLastError();
        cudaFreeAsync(chunk_tot, q->stream);
        QVM_CUDA(q, e);
        q->stats.kernel_launches += 2;
    }
    QVM_CUDA(q, cudaMemcpyAsync(q->result_host, q->total_dev, sizeof(double), cudaMemcpyDeviceToHost, q->stream));
    QVM_CUDA(q, cudaStreamSynchronize(q->stream));
    if (total) *total = q->result_host[0];
    q->cdf_blocks = n_blocks;
    q->sample_mode = mode;
    q->sample_n_rho = n_rho;
    q->sample_n_probs = n_probs;
    q->stats.kernel_launches += 2;
    q->stats.hbm_bytes += 16ull * (mode == 0 ? q->n_elems : n_probs);
    q->stats.d2h_bytes += 8;
    return QVM_OK;
}

int qvm_sample(qvm_t* q, uint64_t n_shots, const double* uniforms, uint64_t* out_index) {
    QVM_CHECK_HANDLE(q);
    if (q->sample_mode < 0) return fail(QVM_ERR_INVALID, "call qvm_probs_prepare first (state changed since)");
    if (!n_shots) return QVM_OK;
    if (!uniforms || !out_index) return fail(QVM_ERR_INVALID, "null pointer");
    const uint64_t row_base = q->sample_mode == 1 ? ((q->shard_index << q->n_local) >> q->sample_n_rho) : 0;
    double* u_dev = nullptr;
    unsigned long long* idx_dev = nullptr;
    const uint64_t chunk = 1ull << 22;
    int rc = ensure_stage(q, std::min<uint64_t>(n_shots, chunk) * 8);
    if (rc) return rc;
    QVM_CUDA(q, cudaMallocAsync(&u_dev, std::min(n_shots, chunk) * 8, q->stream));
    cudaError_t err = cudaMallocAsync(&idx_dev, std::min(n_shots, chunk) * 8, q->stream);
    // (from here on every exit path releases the two scratch allocations)
    for (uint64_t done = 0; err == cudaSuccess && done < n_shots; done += chunk) {
        const uint64_t n = std::min(chunk, n_shots - done);
        memcpy(q->stage, uniforms + done, n * 8);
        err = cudaMemcpyAsync(u_dev, q->stage, n * 8, cudaMemcpyHostToDevice, q->stream);
        if (err != cudaSuccess) break;
        const int threads = 128;
        const int grid = grid_for(q, n * 32, threads, 16);
        sample_kernel<<<grid, threads, 0, q->stream>>>(q->amps, q->sample_n_probs, q->sample_mode, q->sample_n_rho,
                                                       row_base, q->cdf, q->cdf_blocks, q->total_dev, u_dev, n, idx_dev);
        if ((err = cudaGetLastError()) != cudaSuccess) break;
        if ((err = cudaMemcpyAsync(q->stage, idx_dev, n * 8, cudaMemcpyDeviceToHost, q->stream)) != cudaSuccess) break;
        if ((err = cudaStreamSynchronize(q->stream)) != cudaSuccess) break;
        memcpy(out_index + done, q->stage, n * 8);
        q->stats.kernel_launches += 1;
        q->stats.h2d_bytes += n * 8;
        q->stats.d2h_bytes += n * 8;
    }
    if (u_dev) cudaFreeAsync(u_dev, q->stream);
    if (idx_dev) cudaFreeAsync(idx_dev, q->stream);
    if (err != cudaSuccess) {
        q->sticky = QVM_ERR_CUDA;
        return fail(QVM_ERR_CUDA, "qvm_sample: %s", cudaGetErrorString(err));
    }
    return QVM_OK;
}

// ---- shard exchange plumbing ---------------------------------------------------------------------------
int qvm_pack_half(qvm_t* q, int local_bit, int bit_value, uint64_t elem_offset, uint64_t count, void* device_dst) {
    QVM_CHECK_HANDLE(q);
    if (local_bit < 0 || local_bit >= q->n_local || (bit_value | 1) != 1) return fail(QVM_ERR_INVALID, "bad bit");
    if (!device_dst && count) return fail(QVM_ERR_INVALID, "null buffer");
    const uint64_t half = q->n_elems >> 1;
    if (elem_offset > half || count > half - elem_offset) return fail(QVM_ERR_INVALID, "range outside the half-shard");
    if (!count) return QVM_OK;
    pack_half_kernel<<<grid_for(q, count, 256, 16), 256, 0, q->stream>>>(q->amps, local_bit, bit_value, elem_offset, count,
                                                                         static_cast<double2*>(device_dst));
    QVM_CUDA(q, cudaGetLastError());
    q->stats.kernel_launches += 1;
    q->stats.hbm_bytes += 32ull * count;
    return QVM_OK;
}

int qvm_unpack_half(qvm_t* q, int local_bit, int bit_value, uint64_t elem_offset, uint64_t count,
                    const void* device_src) {
    QVM_CHECK_HANDLE(q);
    if (local_bit < 0 || local_bit >= q->n_local || (bit_value | 1) != 1) return fail(QVM_ERR_INVALID, "bad bit");
    if (!device_src && count) return fail(QVM_ERR_INVALID, "null buffer");
    const uint64_t half = q->n_elems >> 1;
    if (elem_offset > half || count > half - elem_offset) return fail(QVM_ERR_INVALID, "range outside the half-shard");
    if (!count) return QVM_OK;
    unpack_half_kernel<<<grid_for(q, count, 256, 16), 256, 0, q->stream>>>(q->amps, local_bit, bit_value, elem_offset,
                                                                           count, static_cast<const double2*>(device_src));
    QVM_CUDA(q, cudaGetLastError());
    q->sample_mode = -1;
    q->stats.kernel_launches += 1;
    q->stats.hbm_bytes += 32ull * count;
    return QVM_OK;
}

// ---- peer-memory exchange (one process per GPU: CUDA IPC over NVLink) -----------------------------------
int qvm_ipc_export(qvm_t* q, void* handle64) {
    QVM_CHECK_HANDLE(q);
    if (!handle64) return fail(QVM_ERR_INVALID, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    QVM_CUDA(q, cudaIpcGetMemHandle(&h, q->amps));
    q->exported = true;
    memcpy(handle64, &h, sizeof(h));
    return QVM_OK;
}

int qvm_ipc_open(qvm_t* q, const void* handle64, void** peer_ptr) {
    QVM_CHECK_HANDLE(q);
    if (!handle64 || !peer_ptr) return fail(QVM_ERR_INVALID, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    *peer_ptr = nullptr;
    QVM_CUDA(q, cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return QVM_OK;
}

int qvm_ipc_close(qvm_t* q, void* peer_ptr) {
    QVM_CHECK_HANDLE(q);
    if (!peer_ptr) return QVM_OK;
    QVM_CUDA(q, cudaStreamSynchronize(q->stream));
    QVM_CUDA(q, cudaIpcCloseMemHandle(peer_ptr));
    return QVM_OK;
}

int qvm_exchange_block(qvm_t* q, int m, const int32_t* local_bits, uint32_t mine_value, uint32_t peer_value,
                       void* peer_amps, uint64_t elem_offset, uint64_t count) {
    QVM_CHECK_HANDLE(q);
    if (m < 1 || m > 8 || m > q->n_local || !local_bits) return fail(QVM_ERR_INVALID, "bad victim bit count %d", m);
    if (!peer_amps) return fail(QVM_ERR_INVALID, "null peer pointer");
    if ((mine_value >> m) || (peer_value >> m) || mine_value == peer_value)
        return fail(QVM_ERR_INVALID, "sub-block selectors must be distinct %d-bit values", m);
    ExchangeBits b{};
    b.m = m;
    for (int i = 0; i < m; ++i) {
        const int p = local_bits[i];
        if (p < 0 || p >= q->n_local || (i && p <= local_bits[i - 1]))
            return fail(QVM_ERR_INVALID, "victim bits must be ascending local positions");
        b.pos[i] = (uint8_t)p;
        b.mine_bits |= (uint64_t)((mine_value >> i) & 1u) << p;
        b.peer_bits |= (uint64_t)((peer_value >> i) & 1u) << p;
    }
    const uint64_t block = q->n_elems >> m;
    if (elem_offset > block || count > block - elem_offset) return fail(QVM_ERR_INVALID, "range outside the sub-block");
    if (!count) return QVM_OK;
    const int grid = grid_for(q, (count + kExchangeUnroll - 1) / kExchangeUnroll, 256, 8);
    exchange_block_kernel<<<grid, 256, 0, q->stream>>>(q->amps, static_cast<double2*>(peer_amps), b, elem_offset, count);
    QVM_CUDA(q, cudaGetLastError());
    q->sample_mode = -1;
    q->stats.kernel_launches += 1;
    q->stats.hbm_bytes += 32ull * count;
    return QVM_OK;
}

// ---- out-of-place exchange: every rank pushes into its peers' alternate buffers ----------------------------
int qvm_alt_alloc(qvm_t* q) {
    QVM_CHECK_HANDLE(q);
    if (q->alt) return QVM_OK;
    if (!q->owns_amps) return fail(QVM_ERR_UNSUPPORTED, "handles on caller-owned memory have no alternate buffer");
    cudaError_t e = cudaMalloc(&q->alt, q->n_elems * sizeof(double2));
    if (e == cudaErrorMemoryAllocation) {            // whatever the parking lot holds is worth less than this buffer
        cudaGetLastError();
        parked_release_all();
        cudaSetDevice(q->device);
        e = cudaMalloc(&q->alt, q->n_elems * sizeof(double2));
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        q->alt = nullptr;
        return fail(e == cudaErrorMemoryAllocation ? QVM_ERR_NOMEM : QVM_ERR_CUDA, "alternate buffer of %llu bytes: %s",
                    (unsigned long long)(q->n_elems * sizeof(double2)), cudaGetErrorString(e));
    }
    return QVM_OK;
}

int qvm_alt_free(qvm_t* q) {
    QVM_CHECK_HANDLE(q);
    if (!q->alt) return QVM_OK;
    QVM_CUDA(q, cudaStreamSynchronize(q->stream));
    QVM_CUDA(q, cudaFree(q->alt));
    q->alt = nullptr;
    return QVM_OK;
}

void* qvm_alt_ptr(qvm_t* q) { return q ? q->alt : nullptr; }

int qvm_ipc_export_alt(qvm_t* q, void* handle64) {
    QVM_CHECK_HANDLE(q);
    if (!handle64) return fail(QVM_ERR_INVALID, "null pointer");
    if (!q->alt) return fail(QVM_ERR_INVALID, "no alternate buffer (qvm_alt_alloc)");
    cudaIpcMemHandle_t h;
    QVM_CUDA(q, cudaIpcGetMemHandle(&h, q->alt));
    memcpy(handle64, &h, sizeof(h));
    return QVM_OK;
}

static int make_push(qvm_t* q, int m, const int32_t* local_bits, uint32_t mine_value, void* const* dst_ptrs, PushSpec* ps) {
    if (m < 1 || m > kMaxPushBits || m > q->n_local || !local_bits || !dst_ptrs)
        return fail(QVM_ERR_INVALID, "push exchange takes 1..%d victim bits", kMaxPushBits);
    if (!q->alt) return fail(QVM_ERR_INVALID, "no alternate buffer (qvm_alt_alloc)");
    if (mine_value >> m) return fail(QVM_ERR_INVALID, "own global-bit value outside %d bits", m);
    *ps = PushSpec{};
    ps->m = m;
    for (int i = 0; i < m; ++i) {
        const int p = local_bits[i];
        if (p < 0 || p >= q->n_local || (i && p <= local_bits[i - 1]))
            return fail(QVM_ERR_INVALID, "victim bits must be ascending local positions");
        ps->pos[i] = (uint8_t)p;
        ps->mask |= 1ull << p;
        ps->mine_bits |= (uint64_t)((mine_value >> i) & 1u) << p;
    }
    for (uint32_t L = 0; L < (1u << m); ++L) {
        void* d = L == mine_value ? (dst_ptrs[L] ? dst_ptrs[L] : (void*)q->alt) : dst_ptrs[L];
        if (!d) return fail(QVM_ERR_INVALID, "null destination for victim value %u", L);
        if (d == (void*)q->amps) return fail(QVM_ERR_INVALID, "a push must not target the buffer it reads");
        ps->dst[L] = static_cast<double2*>(d);
    }
    return QVM_OK;
}

static int push_standalone(qvm_t* q, const PushSpec& ps) {
    const int grid = grid_for(q, (q->n_elems + kPushUnroll - 1) / kPushUnroll, 256, 8);
    exchange_push_kernel<<<grid, 256, 0, q->stream>>>(q->amps, ps, q->n_elems);
    QVM_CUDA(q, cudaGetLastError());
    q->stats.kernel_launches += 1;
    q->stats.hbm_bytes += 32ull * q->n_elems;
    q->stats.push_launches += 1;
    return QVM_OK;
}

static void push_done(qvm_t* q) {
    std::swap(q->amps, q->alt);
    q->sample_mode = -1;
}

int qvm_exchange_push(qvm_t* q, int m, const int32_t* local_bits, uint32_t mine_value, void* const* dst_ptrs) {
    QVM_CHECK_HANDLE(q);
    PushSpec ps;
    int rc = make_push(q, m, local_bits, mine_value, dst_ptrs, &ps);
    if (rc) return rc;
    rc = push_standalone(q, ps);
    if (rc) return rc;
    push_done(q);
    return QVM_OK;
}

int qvm_apply_group_push(qvm_t* q, int n_tile, const int32_t* tile_bits, int n_ops, const qvm_op_t* ops,
                         const double* mats, uint64_t mats_entries, int n_low, int n_bring, const int32_t* bring_low,
                         int32_t* new_pos, int m, const int32_t* local_bits, uint32_t mine_value, void* const* dst_ptrs,
                         int* fused) {
    QVM_CHECK_HANDLE_LAZY(q);
    if (n_ops < 0 || (n_ops && !ops) || (n_tile && !tile_bits) || !new_pos || (n_bring && !bring_low))
        return fail(QVM_ERR_INVALID, "null pointer");
    if (mats_entries && !mats) return fail(QVM_ERR_INVALID, "null matrix pool");
    if (n_tile < 0 || n_tile > 16 || n_bring < 0 || n_low < 0) return fail(QVM_ERR_INVALID, "bad relabelling request");
    PushSpec ps;
    int rc = make_push(q, m, local_bits, mine_value, dst_ptrs, &ps);
    if (rc) return rc;
    RelabelReq req;
    req.n_low = n_low;
    for (int i = 0; i < n_bring; ++i) {
        int j = 0;
        while (j < n_tile && tile_bits[j] != bring_low[i]) ++j;
        if (j == n_tile) return fail(QVM_ERR_INVALID, "position %d is not in the tile", bring_low[i]);
        req.low_in |= 1u << j;
    }
    bool pushed = false;
    rc = launch_group(q, n_tile, tile_bits, n_ops, ops, mats, mats_entries, &req, &ps, &pushed);
    if (rc != QVM_OK) return rc;
    for (int j = 0; j < n_tile; ++j) new_pos[j] = tile_bits[req.sigma[j]];
    if (pushed) q->stats.fused_push_launches += 1;
    else if ((rc = push_standalone(q, ps)) != QVM_OK) return rc;
    if (fused) *fused = pushed ? 1 : 0;
    push_done(q);
    return QVM_OK;
}

int qvm_exchange_half(qvm_t* q, int local_bit, int mine_value, void* peer_amps, uint64_t elem_offset, uint64_t count) {
    if ((mine_value | 1) != 1) return fail(QVM_ERR_INVALID, "bad bit");
    const int32_t bit = local_bit;
    return qvm_exchange_block(q, 1, &bit, (uint32_t)mine_value, (uint32_t)(1 - mine_value), peer_amps, elem_offset, count);
}

}  // extern "C"
