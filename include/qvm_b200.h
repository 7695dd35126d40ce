/*
 * qvm_b200.h -- C ABI of libqvm_b200.so, the B200 (sm_100a) gate-application engine that sits
 * behind referenceqvm.api.QVMConnection.
 *
 * The reference (rigetti/reference-qvm) is pure Python and has no FFI; each entry point below
 * names the reference code whose work it takes over (paths relative to the reference root).
 * Plain pointers and sizes only -- no torch / numpy types cross this boundary.  The Python host
 * (reference-qvm_b200/referenceqvm/_native.py) binds it with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - A handle owns one shard of a state vector: 2^n_local complex128 amplitudes resident in HBM,
 *     interleaved (re, im).  Amplitude index is little-endian: bit q of the index is qubit q
 *     (reference docs/source/walkthrough.rst:8-44).
 *   - "Physical position" p < n_local is bit p of the local index; n_local <= p < n_local+n_global
 *     is bit (p - n_local) of the shard index (multi-GPU: the top qubits select the GPU).
 *     Diagonal factors and controls may sit on any position; dense targets must be local.
 *   - k-qubit matrices are row-major 2^k x 2^k, interleaved (re, im); gate-local index bit
 *     (k-1-j) belongs to qubits[j] -- the first argument is the most significant bit, exactly the
 *     reference's apply_gate convention (referenceqvm/unitary_generator.py:266-323, pinned by
 *     referenceqvm/tests/test_unitary_generator.py:57-65).
 *   - Every function returns 0 (QVM_OK) or a negative code; qvm_last_error() gives the message
 *     (thread-local).  CUDA failures are sticky on the handle.  There is no CPU fallback.
 *   - One handle = one stream of work; calls on one handle must not overlap.  Host buffers are
 *     caller-owned and fully consumed/filled when the call returns.
 */
#ifndef QVM_B200_H
#define QVM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qvm_state qvm_t;
typedef struct qvm_plan qvm_plan_t;      /* a compiled program: the launches of one flush, kept for replay */

#define QVM_OK               0
#define QVM_ERR_INVALID     -1   /* bad argument (maps to ValueError / TypeError on the host)  */
#define QVM_ERR_CUDA        -2   /* CUDA runtime failure (sticky)                               */
#define QVM_ERR_NOMEM       -3   /* device or host allocation failed                            */
#define QVM_ERR_UNSUPPORTED -4   /* valid request outside what the kernels implement            */

#define QVM_OP_NOP   0           /* identity (classified away)                                  */
#define QVM_OP_DENSE 1           /* dense 2^k x 2^k block on k local targets, optional controls */
#define QVM_OP_DIAG  2           /* 2^k diagonal entries (k may be 0: a scalar), controls       */

#define QVM_MAX_TARGETS   6      /* dense / diagonal target qubits per op                        */
#define QVM_MAX_TILE_BITS 13     /* 2^13 amplitudes = 128 KiB of the 227 KiB shared memory / SM   */

/* One classified gate.  Produced by qvm_classify(), consumed by qvm_apply_group(). */
typedef struct qvm_op {
    int32_t  kind;                        /* QVM_OP_*                                            */
    int32_t  k;                           /* number of targets                                   */
    int32_t  targets[QVM_MAX_TARGETS];    /* physical positions, targets[0] = MSB of the block   */
    uint64_t ctrl_mask;                   /* physical positions that must be 1 for the op to act */
    uint32_t mat_offset;                  /* offset into the matrix pool, in complex entries     */
    uint32_t mat_entries;                 /* 4^k (dense) or 2^k (diag) complex entries           */
} qvm_op_t;

typedef struct qvm_stats {
    uint64_t kernel_launches;             /* launches of this library's kernels                  */
    uint64_t gate_ops;                    /* classified ops executed (NOPs excluded)             */
    uint64_t hbm_passes;                  /* full read+write sweeps of the shard                 */
    uint64_t hbm_bytes;                   /* algorithmic bytes moved by those kernels            */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t regtile_rounds;              /* register-tile rounds run by those launches            */
    uint64_t jit_launches;                /* launches that ran a specialised (NVRTC-built) kernel  */
    uint64_t jit_compiles;                /* group structures submitted for compilation (process)  */
    uint64_t push_launches;               /* stand-alone out-of-place exchange passes (qvm_exchange_push) */
    uint64_t fused_push_launches;         /* gate passes whose stores WERE the exchange (qvm_apply_group_push) */
    uint64_t graph_launches;              /* compiled programs replayed as one CUDA graph launch (qvm_plan_run) */
    uint64_t zero_fused_launches;         /* gate passes that absorbed the reset (read nothing: 16 B per amplitude) */
    uint64_t paced_waits;                 /* launches that waited for their kernel to finish compiling while the GPU
                                             still had earlier passes queued (no device time lost)                  */
} qvm_stats_t;

/* ---- lifetime ------------------------------------------------------------------------------- */
const char* qvm_version(void);
const char* qvm_last_error(void);
int qvm_device_count(int* count);

/* Allocate a shard of 2^n_local amplitudes on `device` (replaces the numpy allocation at
 * referenceqvm/qvm_wavefunction.py:354 and the csc_matrix at qvm_density.py:390-391). */
int qvm_create(int n_local, int device, qvm_t** out);
/* Same, but adopt caller-owned device memory (e.g. a torch CUDA tensor) of 16 * 2^n_local bytes. */
int qvm_create_external(int n_local, int device, void* device_ptr, qvm_t** out);
int qvm_destroy(qvm_t* q);
/* qvm_destroy parks the resources of a healthy, unshared handle whose state is at most 8 GiB (at most two sets, 8 GiB
 * in total) and qvm_create of the same shape on the same device takes them over: a dozen allocations, 8 MiB of them
 * pinned, are milliseconds next to a small program's device time (every density() call returns a cloned handle).  A
 * failed device allocation inside the library empties the lot and retries; qvm_trim does it on request.
 * QVM_B200_PARK=0 disables parking. */
int qvm_trim(void);

/* ---- host-side planning (no device) ----------------------------------------------------------- */
/* Which of the pending ops of a flush go into the next fused group (the inner loop of referenceqvm/_planner.py; the
 * reference has no counterpart: it applies one full-space operator per gate, qvm_wavefunction.py:158-162).  The op
 * stream is given in CSR form: dense[i] = op i has dense targets tgt[tgt_off[i] .. tgt_off[i+1]) (logical qubits
 * < 64), dep[dep_off[i] .. dep_off[i+1]) = earlier ops it does not commute with, done[i] = already run.  Scanning
 * from `first` over at most `lookahead` ops in program order, an op joins when its dependencies have run or joined and
 * its dense targets fit a tile of k qubits grown from seed_mask with at most max_new qubits outside old_mask (the
 * qubits that may be kept on the coalescing positions).  depth 0 = first fit; depth d > 0 also tries, d levels deep,
 * banning one more of the qubits a candidate admitted, and returns the candidate holding the most ops.
 * chosen[0 .. *n_chosen) (capacity max_group_ops) = op indices in program order, *tile_mask = the group's qubits. */
int qvm_plan_choose_group(int n_ops, const uint8_t* dense, const int32_t* tgt_off, const int32_t* tgt, const int32_t* dep_off,
                          const int32_t* dep, const uint8_t* done, int first, int lookahead, int k, int max_new,
                          uint64_t old_mask, uint64_t seed_mask, int max_group_ops, int depth, int32_t* chosen,
                          int32_t* n_chosen, uint64_t* tile_mask);
/* Multi-GPU: this handle is shard `shard_index` of 2^n_global (positions >= n_local). */
int qvm_set_shard(qvm_t* q, int n_global, uint64_t shard_index);
/* Run on a caller-owned cudaStream_t (0 restores the handle's own stream). */
int qvm_set_stream(qvm_t* q, void* cuda_stream);
int qvm_set_tile_bits(qvm_t* q, int tile_bits, int low_bits);   /* tuning (the host plans with 11 / 4) */
/* 1 (default): register-tiled kernel where eligible; 0: shared-memory kernel only (testing).   */
int qvm_set_kernel_path(qvm_t* q, int use_regtile);
/* Specialised kernels for recurring group structures (NVRTC, cached by structure; new angles reuse
 * them): 0 never, 1 (default) compile a structure in the background -- from its first sighting on shards of
 * >= 2^26 amplitudes, from the second on smaller ones -- or load its cubin from the disk cache at once when
 * an earlier process (or the ahead-of-time build) left it there, 2 at first sight, synchronously.  The
 * interpreted kernel runs whatever has no ready kernel.  Env QVM_B200_JIT overrides the default. */
int qvm_set_jit(qvm_t* q, int policy);
/* Block until every background compilation has finished (benchmarks: once, after warm-up). */
int qvm_jit_wait(void);
/* Process exit: drop queued compilations and wait for those in flight (also registered with atexit). */
int qvm_jit_shutdown(void);
/* Diagnostics (no device needed): encode the group, generate its specialised source and compile it to
 * an sm_100a cubin with NVRTC.  0 = compiled, 1 = this group is left to the interpreter, < 0 = the
 * generated source failed to compile (`log` receives the NVRTC log). */
int qvm_spec_compile_check(int n_local, int n_global, uint64_t shard_index, int n_tile, const int32_t* tile_bits,
                           int n_ops, const qvm_op_t* ops, const double* mats, uint64_t mats_entries, char* log,
                           uint64_t log_cap, int n_low, int n_bring, const int32_t* bring_low);
/* Ahead-of-time form of the same (no device needed): plan the rounds of one group exactly as a launch would,
 * report where the relabelling leaves every tile position (new_pos, as qvm_apply_group_relabel does) and put
 * the group's specialised cubin into the disk cache -- <directory of libqvm_b200.so>/_spec_cache unless
 * QVM_B200_CACHE_DIR names another (or "0": no disk cache).  A process that later meets the same structure
 * loads the cubin instead of interpreting the group while NVRTC works.  0 = cubin in the cache, 1 = group is
 * left to the interpreter. */
int qvm_spec_precompile(int n_local, int n_global, uint64_t shard_index, int n_tile, const int32_t* tile_bits,
                        int n_ops, const qvm_op_t* ops, const double* mats, uint64_t mats_entries, int n_low,
                        int n_bring, const int32_t* bring_low, int32_t* new_pos);
/* out4 = {structures submitted to NVRTC, cubins loaded from the disk cache, cubins written to it, kernels ready} */
/* Planning aid (no device, nothing compiled): what the encoder makes of one group -- info4 = {register-tile rounds
 * (-1: the group goes to the shared-memory kernel), ops encoded, rounds that apply a thread scalar, 1 if the group can
 * be specialised} -- and where the relabelling leaves every tile position (new_pos, as qvm_spec_precompile). */
int qvm_group_probe(int n_local, int n_global, uint64_t shard_index, int n_tile, const int32_t* tile_bits, int n_ops,
                    const qvm_op_t* ops, const double* mats, uint64_t mats_entries, int n_low, int n_bring,
                    const int32_t* bring_low, int32_t* new_pos, int32_t* info4);
int qvm_jit_counters(uint64_t* out4);
int qvm_jit_cache_dir(char* out, uint64_t cap);
/* The next n qvm_spec_precompile submissions are compiled after everything queued the normal way: the first passes
 * of a run that starts right away go through the interpreter before any compilation could finish, so the kernels
 * of the later passes should come first. */
int qvm_jit_defer(int n);
/* Several processes of one node (one per GPU) walking the same program share the disk cache: of the calling thread's
 * following qvm_spec_precompile submissions, number i is compiled here only when i % stride == offset; the others are
 * expected from the process with that offset (launches wait for their cubins like for locally queued ones, and a
 * cubin that has not appeared after 20 s is compiled locally after all).  (1, 0) = everything here (the default). */
int qvm_jit_share(int stride, int offset);
/* Debug: how many ops of each register-tile opcode were encoded so far (32 counters). */
int qvm_debug_opcode_hist(qvm_t* q, uint64_t* out32);
int qvm_n_local(const qvm_t* q);
void* qvm_device_ptr(qvm_t* q);         /* raw amplitudes, for NCCL/P2P plumbing on the host     */
void* qvm_stream(qvm_t* q);
int qvm_sync(qvm_t* q);
int qvm_stats(const qvm_t* q, qvm_stats_t* out);
int qvm_stats_reset(qvm_t* q);

/* ---- state I/O (wavefunction() / density() readback: qvm_wavefunction.py:361-367) ------------ */
/* all zeros; amp[0] = 1 if set_amp0 (:354-355).  Lazy: the next fused gate pass absorbs it (its first round reads
 * nothing, the pass costs one write of the shard); any other access to the amplitudes runs the reset kernel first. */
int qvm_reset_zero(qvm_t* q, int set_amp0);
int qvm_set_state(qvm_t* q, uint64_t offset, uint64_t count, const double* re_im);
int qvm_get_state(qvm_t* q, uint64_t offset, uint64_t count, double* re_im);
/* Strided gather of `count` amplitudes: element j is amp[offset + j*stride] (diag of rho). */
int qvm_get_strided(qvm_t* q, uint64_t offset, uint64_t stride, uint64_t count, double* re_im);
/* Readback through a qubit layout (see qvm_apply_group_relabel): re_im[j] = amplitude of the LOGICAL
 * index offset + j*stride, where logical bit i lives at physical position pos_of[i] (a permutation of
 * 0..n_local-1).  The reference has no layout: its self.wf is always in program order. */
int qvm_get_state_layout(qvm_t* q, uint64_t offset, uint64_t stride, uint64_t count, int n_pos,
                         const int32_t* pos_of, double* re_im);
int qvm_copy_state(qvm_t* dst, qvm_t* src);
int qvm_norm2(qvm_t* q, double* out);           /* sum |amp|^2 over the shard                    */

/* ---- gate application ------------------------------------------------------------------------- */
/* Classify a k-qubit matrix on `qubits` into one canonical op: exact-zero off-diagonals make it
 * DIAG, identity blocks make controls (CNOT -> control + X, CPHASE -> two controls + scalar).
 * Writes the reduced matrix/diagonal to `mat_out` (capacity 4^k complex) and its size to
 * op->mat_entries; op->mat_offset is left 0.  Takes over the name -> operator step of
 * tensor_gates / apply_gate (unitary_generator.py:326-363, :266-323).  Validation mirrors
 * apply_gate: non power-of-two / non-square / k<1 -> QVM_ERR_INVALID. */
int qvm_classify(int k, const int32_t* qubits, const double* matrix, qvm_op_t* op, double* mat_out);

/* psi <- lifted(G) psi for one gate: the kernel-side replacement of
 * `unitary = tensor_gates(...); self.wf = unitary.dot(self.wf)` (qvm_wavefunction.py:158-162).
 * Qubit indices are physical positions; out of range / duplicates -> QVM_ERR_INVALID
 * (reference: ValueError at unitary_generator.py:200-203). */
int qvm_apply(qvm_t* q, int k, const int32_t* qubits, const double* matrix);

/* Apply `n_ops` classified ops, in order, in ONE pass over HBM: each CTA stages the 2^n_tile
 * amplitudes addressed by `tile_bits` (sorted physical positions, must contain every dense target
 * of every op) in shared memory, runs all ops on them and writes them back.  `mats` is the
 * matrix pool the ops' mat_offset index (complex entries, interleaved re/im). */
int qvm_apply_group(qvm_t* q, int n_tile, const int32_t* tile_bits, int n_ops, const qvm_op_t* ops,
                    const double* mats, uint64_t mats_entries);

/* Same pass, plus an in-tile qubit relabelling on the way out: the qubits now at the positions
 * `bring_low[0..n_bring)` (members of `tile_bits`) are to leave the pass on positions < n_low -- the
 * coalescing bits every tile contains -- trading places with qubits there that were not asked for.  The
 * permutation costs no extra HBM traffic (the tile is transposed in shared memory between rounds anyway).
 * new_pos[j] (n_tile entries) = position after the pass of the qubit that was at tile_bits[j]; the
 * caller tracks the layout.  The library may relabel less than asked (new_pos says what happened).
 * There is no counterpart in the reference: it never reorders qubits (SWAP is an explicit operator,
 * unitary_generator.py:91-150); this only changes where amplitudes live between passes. */
int qvm_apply_group_relabel(qvm_t* q, int n_tile, const int32_t* tile_bits, int n_ops, const qvm_op_t* ops,
                            const double* mats, uint64_t mats_entries, int n_low, int n_bring,
                            const int32_t* bring_low, int32_t* new_pos);

/* ---- compiled programs (the caller side of qam.py:61-107 / :168-181: the reference re-walks, re-builds and
 * re-multiplies every operator on every call; here a program's gate stream is classified, fused, planned and
 * encoded ONCE) ------------------------------------------------------------------------------------------ */
/* Between qvm_record_begin and qvm_record_end every group launched on the handle (qvm_apply_group /
 * qvm_apply_group_relabel) is executed as usual AND kept, with everything a launch needs: the encoded round
 * plan, the specialised source and its launch-time factors.  qvm_record_end hands the recording over as a plan
 * (plan == NULL: discard it).  qvm_plan_run replays the launches in order on the handle's current state
 * (reset_first != 0: preceded by the reset to |0...0>): each group runs its specialised kernel as soon as
 * that is compiled, its interpreted form until then.  Once every group has its kernel and the plan has been
 * run a few times, the sequence is captured into a CUDA graph and a run costs one cudaGraphLaunch
 * (QVM_B200_GRAPHS=0 turns that off).  A plan is bound to the shard shape it was recorded on; the layout it
 * leaves the qubits in is what the caller observed while recording.  Exchanges cannot be recorded. */
int qvm_record_begin(qvm_t* q);
int qvm_record_end(qvm_t* q, qvm_plan_t** plan);
int qvm_plan_run(qvm_t* q, qvm_plan_t* plan, int reset_first);
int qvm_plan_launches(const qvm_plan_t* plan);
int qvm_plan_destroy(qvm_plan_t* plan);

/* rho[i,j] *= f^popcount((i xor j) & mask) on a 2n-qubit vectorised rho (n = n_total/2): the fused
 * form of `dephasing` Kraus pairs on every qubit in `mask` (qvm_density.py:298-313, :350-361). */
int qvm_dephase_all(qvm_t* q, int n_rho_qubits, uint64_t qubit_mask, double factor);

/* ---- MEASURE (qvm_wavefunction.py:96-129, :147-156; qvm_density.py:120-143, :150-156) --------- */
/* One cooperative kernel: p0 = sum_{bit=0} |amp|^2 (warp-shuffle + block reduce), outcome 0 iff
 * r < p0, zero the other half and scale by 1/sqrt(p_outcome).  Single-shard handles only. */
int qvm_measure(qvm_t* q, int qubit, double r, int* outcome, double* p0);
/* Runs of MEASURE instructions (the Measurement branch of qvm_wavefunction.py:147-156 executed back to
 * back).  One pass (a) optionally applies the combined collapse of the previous chunk -- amplitudes whose
 * index has (index & keep_mask) == keep_value are scaled by `scale`, the others zeroed -- and (b) returns
 * hist[bin][lane] (32 * 2^n_high doubles): the sum of |amp|^2 over the (collapsed) amplitudes whose bits at
 * high_pos[0..n_high) (each >= 5) spell `bin` (high_pos[i] = bit i) and whose 5 lowest index bits are
 * `lane`.  The caller replays the reference's sequential decisions on it -- p0 of each measured qubit given
 * the outcomes so far, one uniform each, outcome 0 iff r < p0, scale 1/sqrt(p0) or 1/sqrt(1 - p0) -- and
 * passes the combined projection to the next call (n_high = 0, hist = NULL for a final collapse only).
 * do_collapse = 2 ("filter"): nothing is written; the histogram is taken over the amplitudes with
 * (index & keep_mask) == keep_value only, times scale^2 -- only those amplitudes are read, so after the first
 * chunk of a run has fixed ten or so bits the following chunks cost next to nothing and the run ends with ONE
 * collapse call (which, when every local position was measured, is a write-only clear + one amplitude).
 * Summation order is fixed: results are deterministic to the last bit.  On a shard the positions are
 * local and the histogram covers the shard; the caller gathers the shards' histograms (a measured global
 * position is the same bit for a whole shard) and hands every shard its own projection. */
int qvm_measure_hist(qvm_t* q, int n_high, const int32_t* high_pos, int do_collapse, uint64_t keep_mask,
                     uint64_t keep_value, double scale, double* hist);
/* Batched trials (qvm_wavefunction.py:266-283: run() re-simulates the program once per trial).  The
 * handle's 2^n_local amplitudes are read as 2^(n_local - n_per_trial) independent n_per_trial-qubit
 * states, one per trial; gates applied to positions < n_per_trial act on all of them in the same pass.
 * qvm_reset_batched sets every trial to |0..0>; qvm_measure_batched measures `qubit` in every trial
 * with that trial's own uniform (uniforms / outcomes / p0: one entry per trial slot). */
int qvm_reset_batched(qvm_t* q, int n_per_trial);
/* Regroup trials (classical control, qvm_wavefunction.py:164-231: trials whose jumps go different ways): segment j of
 * `dst` := segment src_index[j] of `src` for j < count; both handles on one device, their other segments untouched. */
int qvm_gather_segments(qvm_t* dst, qvm_t* src, int n_per_trial, uint64_t count, const uint32_t* src_index);
int qvm_measure_batched(qvm_t* q, int n_per_trial, int qubit, const double* uniforms, int32_t* outcomes, double* p0);
/* The two halves of the above for sharded states (host all-reduces p0 in between). */
int qvm_prob_zero(qvm_t* q, int qubit, double* p0_partial);
int qvm_collapse(qvm_t* q, int qubit, int outcome, double scale);
/* Density flavour on a vectorised rho of n_rho_qubits: p0 = tr(P0 rho) over the local rows. */
int qvm_density_prob_zero(qvm_t* q, int n_rho_qubits, int qubit, double* p0_partial);

/* diag(rho) through a qubit layout (rho relabelled in its tiles and / or sharded by rows; replaces the Python
 * loop of sparse_trace, qvm_density.py:51-54, and the diagonal read at :444): pos_of[v] (2 n_rho entries) is
 * the physical position of vector qubit v -- v < n_rho: column qubit v, v >= n_rho: row qubit v - n_rho.
 * diag_host receives 2^n_rho doubles: Re rho[i,i] for the entries this shard owns, 0.0 for the others (the
 * shards' arrays add up to the full diagonal exactly). */
int qvm_density_diag(qvm_t* q, int n_rho_qubits, const int32_t* pos_of, double* diag_host);

/* ---- expectation values (the reference declares expectation() and leaves it NotImplementedError:
 * qvm_density.py:396-408, qvm_unitary.py:99-112; the operator it documents is tensor_up(PauliSum),
 * unitary_generator.py:366-411) ------------------------------------------------------------------------------ */
/* One Pauli string on a state vector.  With xmask = positions of its X and Y factors (local positions only) and
 * zmask = positions of its Z and Y factors (any position), P|x> = i^ny (-1)^popcount(x & zmask) |x ^ xmask>;
 * re_im_partial receives this shard's sum_x conj(psi[x ^ xmask]) (-1)^popcount(x & zmask) psi[x] (the caller adds
 * the shards and multiplies by i^ny and the term's coefficient).  One read pass. */
int qvm_pauli_expectation(qvm_t* q, uint64_t xmask, uint64_t zmask, double* re_im_partial);
/* rho[i, i ^ col_xor] for all 2^n_rho rows i through a layout (as qvm_density_diag, complex, 0 where not owned):
 * tr(rho P) = i^ny sum_i (-1)^popcount(i & zmask) rho[i, i ^ xmask]. */
int qvm_density_offdiag(qvm_t* q, int n_rho_qubits, const int32_t* pos_of, uint64_t col_xor, double* re_im_host);
/* sum_x conj(a[x]) b[x] over two shards of equal shape on one device (general operator programs: <psi|O psi>). */
int qvm_inner_product(qvm_t* a, qvm_t* b, double* re_im_partial);

/* ---- sampling (qvm_density.py:442-452; statistical meaning of qvm_wavefunction.py:266-321) ----- */
/* Build the two-level CDF of |amp|^2 (mode 0), of diag(rho) (mode 1, n_rho_qubits = n, identity layout, rows
 * sharded), or of a plain array of 2^(n_local+1) probabilities that the caller stored in the handle's buffer
 * as raw doubles (mode 2: a diagonal gathered by qvm_density_diag). */
int qvm_probs_prepare(qvm_t* q, int mode, int n_rho_qubits, double* total);
/* Inverse-CDF draws: index = first i with cdf[i] > uniforms[s] * total (numpy's searchsorted
 * 'right' on the normalised cdf).  Indices are local to the shard. */
int qvm_sample(qvm_t* q, uint64_t n_shots, const double* uniforms, uint64_t* out_index);

/* ---- shard exchange plumbing (global<->local qubit remap) --------------------------------------- */
/* Gather / scatter elements [elem_offset, elem_offset+count) of the half-shard whose local bit
 * `local_bit` equals `bit_value`, to / from a contiguous device buffer. */
int qvm_pack_half(qvm_t* q, int local_bit, int bit_value, uint64_t elem_offset, uint64_t count,
                  void* device_dst);
int qvm_unpack_half(qvm_t* q, int local_bit, int bit_value, uint64_t elem_offset, uint64_t count,
                    const void* device_src);

/* ---- peer-memory exchange: the same swap as ONE kernel over NVLink P2P, no staging ----------------- */
/* One process per GPU: export this shard's allocation as a 64-byte CUDA IPC handle / map a peer's. */
int qvm_ipc_export(qvm_t* q, void* handle64);
int qvm_ipc_open(qvm_t* q, const void* handle64, void** peer_device_ptr);
int qvm_ipc_close(qvm_t* q, void* peer_device_ptr);
/* For j in [elem_offset, elem_offset+count): swap my amp[insert(j, local_bit, mine_value)] with the
 * peer's amp[insert(j, local_bit, 1 - mine_value)].  The two ranks of a pair call it on disjoint j
 * ranges (each half of [0, 2^(n_local-1))); the caller brackets it with a cross-rank barrier on the
 * stream.  `peer_amps` may also be another shard on the same device (tests). */
int qvm_exchange_half(qvm_t* q, int local_bit, int mine_value, void* peer_amps, uint64_t elem_offset,
                      uint64_t count);
/* m victim bits at once (m global qubits swapped with m local ones in a single pass): my sub-block
 * "victim bits == mine_value" trades places with the peer's sub-block "victim bits == peer_value"
 * (bit i of the values belongs to local_bits[i], ascending).  A rank calls it once per peer of its
 * 2^m-rank subgroup with mine_value = the peer's global bits and peer_value = its own. */
int qvm_exchange_block(qvm_t* q, int m, const int32_t* local_bits, uint32_t mine_value, uint32_t peer_value,
                       void* peer_amps, uint64_t elem_offset, uint64_t count);


/* ---- out-of-place exchange: write-only pushes over NVLink, optionally fused into the last gate pass ---- */
/* The same global<->local swap of m (<= 3) qubit pairs as qvm_exchange_block, but out of place: a second
 * buffer of the shard's size ("alternate", qvm_alt_alloc; QVM_ERR_NOMEM when it does not fit -- 36 qubits
 * on 8 GPUs -- and the in-place exchange stays the path).  Every rank reads its current buffer ONCE and
 * stores each amplitude where it will live: dst_ptrs[L] (2^m entries) is the ALTERNATE buffer of the rank
 * whose global bits equal L (bit i of L belongs to local_bits[i], ascending local positions, and to the
 * i-th swapped global bit); dst_ptrs[mine_value] may be NULL (= this handle's own alternate buffer).  The
 * amplitude at local index x goes to dst_ptrs[victim bits of x] at index (x with the victim bits replaced
 * by mine_value).  On return (stream-ordered) the handle's current and alternate buffers have swapped
 * roles; the caller runs ONE cross-rank barrier on the stream before anyone reads its new current buffer
 * (the previous epoch's barrier already guarantees nobody still reads the buffer being written).
 * Peer pointers come from qvm_ipc_export / qvm_ipc_export_alt + qvm_ipc_open; all ranks push in the same
 * epochs, so "peer's alternate buffer" alternates in lock step with the caller's own.
 * No reference counterpart (the reference has no distributed path). */
int qvm_alt_alloc(qvm_t* q);
int qvm_alt_free(qvm_t* q);
void* qvm_alt_ptr(qvm_t* q);
int qvm_ipc_export_alt(qvm_t* q, void* handle64);
int qvm_exchange_push(qvm_t* q, int m, const int32_t* local_bits, uint32_t mine_value, void* const* dst_ptrs);
/* qvm_apply_group_relabel + qvm_exchange_push as ONE pass: the fused group's last round stores every tile
 * directly into its destination shard (local HBM or a peer over NVLink), so the exchange costs no pass of
 * its own.  Needs the victim bits outside the tile; otherwise (or when the group needs no launch / runs on
 * the shared-memory fallback kernel) the library runs the group in place and the push after it.  *fused
 * says which happened.  Buffers swap roles as in qvm_exchange_push. */
int qvm_apply_group_push(qvm_t* q, int n_tile, const int32_t* tile_bits, int n_ops, const qvm_op_t* ops,
                         const double* mats, uint64_t mats_entries, int n_low, int n_bring, const int32_t* bring_low,
                         int32_t* new_pos, int m, const int32_t* local_bits, uint32_t mine_value,
                         void* const* dst_ptrs, int* fused);

#ifdef __cplusplus
}
#endif
#endif /* QVM_B200_H */
