// Specialised kernels for fused groups that come back: the interpreter of qvm_regtile.cuh spends most
// of its time on control flow (profiles/README.md), so a group whose *structure* has been seen before
// is turned into straight-line CUDA source -- the very same device functions (reg_h, reg_s1,
// cmul_ip, add_tree ...) called with compile-time slots, masks and strides, the complex factors
// passed at launch time through the kernel-parameter constant bank -- compiled for sm_100a with
// NVRTC and cached by source text.  New angles on an old structure reuse the kernel.
//
// Policy (QVM_B200_JIT): 0 = never, 1 (default) = compile a structure in the background when it shows
// up for the second time (the interpreter keeps serving it until the kernel is ready), 2 = compile at
// first sight and wait for it.  The interpreted kernel stays the path for everything that
// is new, that NVRTC cannot build (library missing) or that the generator does not cover (FULL
// groups): this is a GPU path selection, never a CPU fallback.
#pragma once
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "qvm_encode.h"

namespace qvm {

struct SpecParams {
    int n_runs;
    uint64_t shard_bits, run_start, run_len;
    uint64_t n_tiles;                // pipelined variant: tiles are walked by a persistent grid
    PushSpec push;                   // fused exchange: where the last round stores (launch-time, m == 0: in place)
    int from_zero;                   // the reset folded into this pass (see RegTileParams::from_zero)
};

struct SpecSource {
    std::string src;                 // the whole translation unit (without the embedded header)
    std::vector<double> factors;     // launch-time values of F.v[]
    unsigned threads = 0;
    size_t smem = 0;
    uint64_t n_tiles = 0;
    unsigned ctas_per_sm = 0;        // > 0: persistent (pipelined) kernel, grid = min(n_tiles, SMs * ctas_per_sm)
    SpecParams params{};
};

namespace specgen {

inline std::string hexu(uint64_t v) {
    char buf[32];
    snprintf(buf, sizeof(buf), "0x%llxull", (unsigned long long)v);
    return buf;
}
inline std::string outer_cond(uint64_t co) { return co ? "((full & " + hexu(co) + ") == " + hexu(co) + ")" : ""; }
inline std::string thr_cond(uint32_t cthr) {
    char buf[64];
    if (!cthr) return "";
    snprintf(buf, sizeof(buf), "((tid & 0x%xu) == 0x%xu)", cthr, cthr);
    return buf;
}
inline std::string both(const std::string& a, const std::string& b) {
    if (a.empty()) return b;
    if (b.empty()) return a;
    return a + " && " + b;
}
inline std::string oi_expr(const ColdOp& c) {
    std::string e;
    for (int t = 0; t < c.k; ++t)
        if ((c.loc[t] & LOC_CLASS) == LOC_OUT) {
            char buf[96];
            snprintf(buf, sizeof(buf), "(((uint32_t)(full >> %d) & 1u) << %d)", c.loc[t] & LOC_INDEX, c.k - 1 - t);
            e += (e.empty() ? "" : " | ") + std::string(buf);
        }
    return e.empty() ? "0u" : "(" + e + ")";
}

}  // namespace specgen

// Resident CTAs per SM the specialised kernels are compiled for, in units of 128-thread CTAs
// (QVM_B200_SPEC_OCC): the register budget per thread follows from it (65536 / threads per SM).
inline unsigned spec_min_blocks(unsigned threads) {
    static const int occ = [] {
        const char* e = getenv("QVM_B200_SPEC_OCC");
        const int v = e ? atoi(e) : 4;
        return v < 1 ? 1 : (v > 8 ? 8 : v);
    }();
    const unsigned b = (unsigned)occ * 128u / (threads ? threads : 128u);
    return b < 1 ? 1 : b;
}

// QVM_B200_SPEC_STREAM=1: tile loads / stores with the evict-first (.cs) cache operator.
inline bool spec_streaming() {
    static const bool on = [] {
        const char* e = getenv("QVM_B200_SPEC_STREAM");
        return e && atoi(e) != 0;
    }();
    return on;
}

// QVM_B200_SPEC_PIPE: how a CTA overlaps its load, FP64 and store phases.
//   0  one tile per CTA: round 0 loads HBM -> registers, the last round stores registers -> HBM; only the other
//      resident CTAs hide one phase behind another
//   1  persistent CTAs that prefetch their next tile with cp.async into a second shared-memory buffer (measured
//      slower in round 1: the staged tile costs one more trip through the LSU pipe)
//   2  persistent CTAs with TMA: every thread moves one 256/512-byte run of the tile with cp.async.bulk -- the
//      next tile lands in the second buffer (mbarrier, no registers, no LSU traffic) while this one is computed,
//      and the finished tile drains to HBM (or to a peer over NVLink) asynchronously from shared memory
inline int spec_pipe_mode() {
    static const int mode = [] {
        const char* e = getenv("QVM_B200_SPEC_PIPE");
        const int v = e ? atoi(e) : 0;
        return v < 0 || v > 2 ? 0 : v;
    }();
    return mode;
}

// Source for one encoded group (lean R = 4 groups only).  Factors never appear as literals: F.v[i]
// indexes a by-value kernel argument -- first the encoder's matrix pool, then the factors the encoder
// inlined into the op stream, then the round scalars.
inline bool generate_spec(const EncodedGroup& g, SpecSource& out) {
    using namespace specgen;
    const RegTileParams& P = g.P;
    if ((g.cold.empty() && !g.relabelled) || g.need_full || P.R != 4) return false;
    std::vector<double>& F = out.factors;
    F.assign(g.pool.begin(), g.pool.end());
    auto fconst = [&](double re, double im) -> std::string {          // a new launch-time complex factor
        const size_t i = F.size();
        F.push_back(re);
        F.push_back(im);
        return "make_double2(F.v[" + std::to_string(i) + "], F.v[" + std::to_string(i + 1) + "])";
    };
    auto fword = [&](uint32_t w) -> std::string {                     // factor stored in stream word w
        double re, im;
        memcpy(&re, &g.hot[w].x, 8);
        memcpy(&im, &g.hot[w].z, 8);
        return fconst(re, im);
    };
    auto fpool = [&](const std::string& index) -> std::string {       // pool[index] with a run-time index
        return "make_double2(F.v[2 * (" + index + ")], F.v[2 * (" + index + ") + 1])";
    };
    std::ostringstream o;
    std::vector<int> cold_of_word(g.hot.size(), -1);
    for (int j = 0; j < P.n_ops; ++j) cold_of_word[g.cold[j].word_off] = j;
    auto find_patch = [&](uint32_t word, int slot, int kind) -> const ColdOp* {
        for (int j = P.n_ops; j < P.n_ops + P.n_patches; ++j)
            if (g.cold[j].word_off == word && g.cold[j].slot == slot && g.cold[j].kind == kind) return &g.cold[j];
        return nullptr;
    };
    const bool pipe = spec_pipe_mode() == 1 && P.k - 4 >= P.low;
    const bool tma = spec_pipe_mode() == 2 && P.low >= 4 && P.low <= 5;     // runs of 256 / 512 bytes, <= one per thread
    auto lin_of = [](uint16_t swz_stride) {                                  // tile-local bit of a slot: swz() keeps the top bit
        int b = 0;
        while ((swz_stride >> (b + 1)) != 0) ++b;
        return 1u << b;
    };
    std::ostringstream body;
    for (int r = 0; r < P.n_rounds; ++r) {
        const RoundDesc& rd = g.rounds[r];
        const bool first = r == 0, last = r == P.n_rounds - 1;
        body << "  {  // ---- round " << r << "\n    uint32_t T = tid;\n";
        for (int s = 0; s < 4; ++s) body << "    T += T & 0x" << std::hex << rd.ins_mask[s] << std::dec << "u;\n";
        body << "    const uint32_t sT = swz(T);\n    const uint32_t st[4] = {" << rd.swz_j[0] << "u, " << rd.swz_j[1] << "u, "
             << rd.swz_j[2] << "u, " << rd.swz_j[3] << "u};\n";
        if (first && tma) {
            // the tile sits in shared memory as TMA delivered it: linear, lanes on the lowest tile bits (conflict-free)
            body << "    const uint32_t lin[4] = {" << lin_of(rd.swz_j[0]) << "u, " << lin_of(rd.swz_j[1]) << "u, "
                 << lin_of(rd.swz_j[2]) << "u, " << lin_of(rd.swz_j[3]) << "u};\n"
                    "#pragma unroll\n    for (int j = 0; j < 16; ++j) a[j] = tile[T ^ xor_combo<4>(lin, j)];\n";
        } else if (first && !pipe) {
            body << "    if (SP.from_zero) {\n#pragma unroll\n      for (int j = 0; j < 16; ++j) a[j] = make_double2(0.0, 0.0);\n"
                    "      if (SP.from_zero == 2 && base == 0 && tid == 0) a[0].x = 1.0;\n    } else\n";
            body << "    {\n      const uint64_t gT = (uint64_t)(T & " << ((1u << P.low) - 1u) << "u) | hi_off[T >> " << P.low
                 << "];\n      const int64_t b[4] = {";
            for (int s = 0; s < 4; ++s) body << (s ? ", " : "") << "(int64_t)(16ull << " << (int)rd.reg_phys[s] << ")";
            body << "};\n      const char* q[16];\n      add_tree<4>(reinterpret_cast<const char*>(gbase + gT), b, q);\n"
                    "#pragma unroll\n      for (int j = 0; j < 16; ++j) a[j] = "
                 << (spec_streaming() ? "__ldcs(reinterpret_cast<const double2*>(q[j]))" : "*reinterpret_cast<const double2*>(q[j])")
                 << ";\n    }\n";
        } else {
            body << "#pragma unroll\n    for (int j = 0; j < 16; ++j) a[j] = tile[sT ^ xor_combo<4>(st, j)];\n";
        }
        if ((rd.flags & 1u) || rd.dt_end > rd.op_begin) {
            body << "    double2 ts = " << fconst(rd.scale.x, rd.scale.y) << ";\n";
            for (uint32_t w = rd.op_begin; w < rd.dt_end; w += kOpWords) {
                const uint4 h = g.hot[w];
                const ColdOp& c = g.cold[cold_of_word[w]];
                const uint32_t opc = h.x & 0xFFu, cthr = h.y & 0xFFFFu, p0 = (h.x >> 16) & 0xFFu, p1 = h.x >> 24;
                const std::string cond = both(outer_cond(c.ctrl_outer), thr_cond(cthr));
                std::string f;
                if (opc == OPC_T) {
                    if (c.prefetch) {
                        const std::string i0 = std::to_string(c.mat_off) + "u + " + oi_expr(c);
                        if (c.fsel != 0xFF)
                            f = fpool(i0 + " + ((((tid >> " + std::to_string(p0) + ") & 1u)) << " + std::to_string(c.fsel) + ")");
                        else f = fpool(i0);
                    } else if (p0 != 31 && c.fsel != 0xFF) {
                        const std::string f0 = fword(w + 1), f1 = fword(w + 2);
                        f = "(((tid >> " + std::to_string(p0) + ") & 1u) ? " + f1 + " : " + f0 + ")";
                    } else f = fword(w + 1);
                } else {      // OPC_TG
                    const uint32_t s0 = (h.w >> 8) & 0xFFu, s1 = (h.w >> 16) & 0xFFu;
                    f = fpool(std::to_string(h.z) + "u + (" + oi_expr(c) + " | (((tid >> " + std::to_string(p0) + ") & 1u) << " +
                              std::to_string(s0) + ") | (((tid >> " + std::to_string(p1) + ") & 1u) << " + std::to_string(s1) + "))");
                }
                body << "    " << (cond.empty() ? "" : "if (" + cond + ") ") << "cmul_ip(ts, " << f << ");\n";
            }
            body << "#pragma unroll\n    for (int j = 0; j < 16; ++j) cmul_ip(a[j], ts);\n";
        }
        body << "    uint32_t xm = 0;\n";
        for (uint32_t w = rd.dt_end; w < rd.op_end;) {
            const uint4 h = g.hot[w];
            const uint32_t opc = h.x & 0xFFu, size = (h.y >> HDR_SIZE_SHIFT) & 0xFu;
            if (opc == OPC_LAYER) {
                const uint4 m = g.hot[w + 1];
                const uint32_t cs[4] = {m.x & 0xFFFFu, m.x >> 16, m.y & 0xFFFFu, m.y >> 16};
                const uint32_t cx[4] = {m.z & 0xFFFFu, m.z >> 16, m.w & 0xFFFFu, m.w >> 16};
                for (int s = 0; s < 4; ++s) {
                    if (cs[s] != 0xFFFFu) {
                        const ColdOp* pc = find_patch(w, s, COLD_PATCH_S1);
                        const std::string f = (pc && pc->k) ? fpool(std::to_string(pc->mat_off) + "u + " + oi_expr(*pc))
                                                            : fword(w + 2 + s);
                        const std::string cond = both(pc ? outer_cond(pc->ctrl_outer) : "", thr_cond(cs[s]));
                        body << "    " << (cond.empty() ? "" : "if (" + cond + ") ") << "reg_s1<" << s << ", 16>(a, " << f << ", xm);\n";
                    }
                    if (h.x & (0x1000u << s)) body << "    reg_h<" << s << ", 16>(a, xm);\n";
                    if (cx[s] != 0xFFFFu) {
                        const ColdOp* pc = find_patch(w, s, COLD_PATCH_X);
                        const std::string cond = both(pc ? outer_cond(pc->ctrl_outer) : "", thr_cond(cx[s]));
                        body << "    " << (cond.empty() ? "" : "if (" + cond + ") ") << "xm ^= " << (1u << s) << "u;\n";
                    }
                }
            } else {
                const ColdOp& c = g.cold[cold_of_word[w]];
                const uint32_t cthr = h.y & 0xFFFFu, crm = (h.x >> 8) & 0x1Fu;
                const std::string cond = both(outer_cond(c.ctrl_outer), thr_cond(cthr));
                body << "    " << (cond.empty() ? "" : "if (" + cond + ") ") << "{\n";
                if (opc >= OPC_XS_0 && opc < OPC_XS_0 + 4) {
                    body << "      reg_xswap<" << (opc - OPC_XS_0) << ", 16>(a, " << crm << "u, xm);\n";
                } else if (opc == OPC_DS) {
                    const std::string f = c.prefetch ? fpool(std::to_string(c.mat_off) + "u + " + oi_expr(c)) : fword(w + 1);
                    body << "      const double2 f = " << f << ";\n      const uint32_t want = " << crm << "u & ~xm;\n"
                         << "#pragma unroll\n      for (int p = 0; p < 16; ++p) if ((p & " << crm << "u) == want) cmul_ip(a[p], f);\n";
                } else {
                    return false;      // not a lean op: leave the group to the interpreter
                }
                body << "    }\n";
            }
            w += size;
        }
        if (last && tma) {
            // back to the linear layout TMA stores from (the last round's lanes hold the lowest tile bits)
            if (!first) body << "    __syncthreads();\n";
            body << "    const uint32_t linw[4] = {" << lin_of(rd.swz_j[0]) << "u, " << lin_of(rd.swz_j[1]) << "u, "
                 << lin_of(rd.swz_j[2]) << "u, " << lin_of(rd.swz_j[3]) << "u};\n    uint32_t sX = T;\n"
                    "#pragma unroll\n    for (int s = 0; s < 4; ++s) if ((xm >> s) & 1u) sX ^= linw[s];\n"
                    "#pragma unroll\n    for (int j = 0; j < 16; ++j) tile[sX ^ xor_combo<4>(linw, j)] = a[j];\n";
        } else if (last) {
            body << "    {\n      const uint64_t gT = (uint64_t)(T & " << ((1u << P.low) - 1u) << "u) | hi_off[T >> " << P.low
                 << "];\n      char* p0 = reinterpret_cast<char*>(gstore + gT);\n      int64_t b[4] = {";
            for (int s = 0; s < 4; ++s) body << (s ? ", " : "") << "(int64_t)(16ull << " << (int)rd.reg_phys[s] << ")";
            body << "};\n#pragma unroll\n      for (int s = 0; s < 4; ++s) if ((xm >> s) & 1u) { p0 += b[s]; b[s] = -b[s]; }\n"
                    "      char* q[16];\n      add_tree<4>(p0, b, q);\n"
                    "#pragma unroll\n      for (int j = 0; j < 16; ++j) "
                 << (spec_streaming() ? "__stcs(reinterpret_cast<double2*>(q[j]), a[j])" : "*reinterpret_cast<double2*>(q[j]) = a[j]")
                 << ";\n    }\n";
        } else if (r == P.perm_round) {
            // relabelling round: the tile is written in the last round's coordinates (qvm_encode.h)
            if (!first || pipe || tma) body << "    __syncthreads();\n";
            body << "    uint32_t sX = 0u;\n";
            for (int b = 0; b < P.k; ++b)
                body << "    sX ^= (0u - ((T >> " << b << ") & 1u)) & " << P.perm_swz[b] << "u;\n";
            body << "    const uint32_t wst[4] = {" << rd.wr_swz[0] << "u, " << rd.wr_swz[1] << "u, " << rd.wr_swz[2] << "u, "
                 << rd.wr_swz[3] << "u};\n#pragma unroll\n    for (int s = 0; s < 4; ++s) if ((xm >> s) & 1u) sX ^= wst[s];\n"
                    "#pragma unroll\n    for (int j = 0; j < 16; ++j) tile[sX ^ xor_combo<4>(wst, j)] = a[j];\n";
        } else {
            if (first && tma) body << "    __syncthreads();\n";       // linear tile read by everyone before it is rewritten swizzled
            body << "    uint32_t sX = sT;\n#pragma unroll\n    for (int s = 0; s < 4; ++s) if ((xm >> s) & 1u) sX ^= st[s];\n"
                    "#pragma unroll\n    for (int j = 0; j < 16; ++j) tile[sX ^ xor_combo<4>(st, j)] = a[j];\n";
        }
        body << "  }\n";
        if (!last) body << "  __syncthreads();\n";
    }
    // The pass that absorbs the reset (SpecParams::from_zero): every tile but the one holding amplitude 0 is all zeros
    // and stays all zeros under linear gates -- those CTAs store zeros through the last round's addresses and leave.
    std::ostringstream zero_tile;
    {
        const RoundDesc& rl = g.rounds[P.n_rounds - 1];
        zero_tile << "  if (SP.from_zero && base != 0) {\n    uint32_t T = tid;\n";
        for (int s = 0; s < 4; ++s) zero_tile << "    T += T & 0x" << std::hex << rl.ins_mask[s] << std::dec << "u;\n";
        zero_tile << "    const uint64_t gT = (uint64_t)(T & " << ((1u << P.low) - 1u) << "u) | hi_off[T >> " << P.low
                  << "];\n    const int64_t b[4] = {";
        for (int s = 0; s < 4; ++s) zero_tile << (s ? ", " : "") << "(int64_t)(16ull << " << (int)rl.reg_phys[s] << ")";
        zero_tile << "};\n    char* q[16];\n    add_tree<4>(reinterpret_cast<char*>(gstore + gT), b, q);\n"
                     "#pragma unroll\n    for (int j = 0; j < 16; ++j) *reinterpret_cast<double2*>(q[j]) = make_double2(0.0, 0.0);\n"
                     "    return;\n  }\n";
    }
    if (F.empty()) F.push_back(0.0);
    if (F.size() > 3900) return false;          // kernel parameters are limited to 32 KB
    o << "#include \"qvm_regtile.cuh\"\nusing namespace qvm;\n"
         "struct SpecParams { int n_runs; uint64_t shard_bits, run_start, run_len, n_tiles; PushSpec push; int from_zero; };\n"
         "struct SpecFactors { double v[" << F.size() << "]; };\n"
         "__device__ const uint64_t hi_off[" << g.hi_off.size() << "] = {";
    for (size_t j = 0; j < g.hi_off.size(); ++j) o << (j ? ", " : "") << hexu(g.hi_off[j]);
    o << "};\n";
    const char* tile_base_fn =
        "__device__ __forceinline__ uint64_t spec_tile_base(const SpecParams& SP, uint64_t t) {\n  uint64_t base = 0;\n"
        "  for (int r = 0; r < SP.n_runs; ++r) {\n    const uint32_t len = (uint32_t)(SP.run_len >> (8 * r)) & 0xFFu;\n"
        "    base |= (t & ((1ull << len) - 1ull)) << ((uint32_t)(SP.run_start >> (8 * r)) & 0xFFu);\n    t >>= len;\n  }\n"
        "  return base ^ SP.push.rot;\n}\n";
    o << tile_base_fn;
    const unsigned ctas = (pipe || tma)
                              ? std::max(1u, std::min(spec_min_blocks(g.threads), (unsigned)((220u * 1024u) / ((size_t)32 << P.k))))
                              : spec_min_blocks(g.threads);
    if (tma) {
        const unsigned run_elems = 1u << P.low, n_runs_tile = 1u << (P.k - P.low), run_bytes = 16u << P.low;
        o << "__device__ __forceinline__ void spec_tma_load(double2* buf, unsigned long long* bar, const double2* gsrc, uint32_t tid) {\n"
             "  if (tid == 0) mbar_expect_tx(bar, " << ((size_t)16 << P.k) << "u);\n"
             "  if (tid < " << n_runs_tile << "u) tma_load_run(buf + tid * " << run_elems << "u, gsrc + hi_off[tid], " << run_bytes
          << "u, bar);\n}\n";
        o << "extern \"C\" __global__ void __launch_bounds__(" << g.threads << ", " << ctas
          << ") spec_kernel(double2* __restrict__ state, const __grid_constant__ SpecParams SP, const SpecFactors F) {\n"
             "  extern __shared__ __align__(128) unsigned char smem_raw[];\n"
             "  __shared__ __align__(8) unsigned long long mbar[2];\n"
             "  double2* const buf0 = reinterpret_cast<double2*>(smem_raw);\n  double2* const buf1 = buf0 + " << (1u << P.k) << ";\n"
             "  const uint32_t tid = threadIdx.x;\n  uint64_t t = blockIdx.x;\n  if (t >= SP.n_tiles) return;\n"
             "  if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_fence_init(); }\n  __syncthreads();\n"
             "  spec_tma_load(buf0, &mbar[0], state + spec_tile_base(SP, t), tid);\n  int cur = 0;\n  uint32_t ph0 = 0, ph1 = 0;\n"
             "  for (; t < SP.n_tiles; t += gridDim.x, cur ^= 1) {\n"
             "  double2* const tile = cur ? buf1 : buf0;\n"
             "  tma_wait_read_all();\n  __syncthreads();\n"                      // the other buffer's stores have left shared memory
             "  if (t + gridDim.x < SP.n_tiles) spec_tma_load(cur ? buf0 : buf1, &mbar[cur ^ 1], state + spec_tile_base(SP, t + gridDim.x), tid);\n"
             "  if (cur) { mbar_wait(&mbar[1], ph1); ph1 ^= 1u; } else { mbar_wait(&mbar[0], ph0); ph0 ^= 1u; }\n"
             "  const uint64_t base = spec_tile_base(SP, t);\n"
             "  const uint64_t full = base | SP.shard_bits;\n  double2* gbase = state + base;\n"
             "  double2* gstore = rt_push_base(SP.push, state, base);\n  double2 a[16];\n"
          << body.str()
          << "  fence_async_smem();\n  __syncthreads();\n"
             "  if (tid < " << n_runs_tile << "u) tma_store_run(gstore + hi_off[tid], tile + tid * " << run_elems << "u, " << run_bytes
          << "u);\n  tma_commit();\n  }\n  tma_wait_all();\n}\n";
    } else if (!pipe) {
        o << "extern \"C\" __global__ void __launch_bounds__(" << g.threads << ", " << ctas
          << ") spec_kernel(double2* __restrict__ state, const __grid_constant__ SpecParams SP, const SpecFactors F) {\n"
             "  extern __shared__ __align__(16) unsigned char smem_raw[];\n  double2* tile = reinterpret_cast<double2*>(smem_raw);\n"
             "  const uint32_t tid = threadIdx.x;\n  const uint64_t base = spec_tile_base(SP, blockIdx.x);\n"
             "  const uint64_t full = base | SP.shard_bits;\n  double2* gbase = state + base;\n"
             "  double2* gstore = rt_push_base(SP.push, state, base);\n  double2 a[16];\n"
          << zero_tile.str() << body.str() << "}\n";
    } else {
        // Persistent CTA, two tile buffers.  Thread `tid` fetches the tile-local indices tid + threads * j
        // (j = 0..15): consecutive lanes read consecutive amplitudes (256-byte runs), the global offset splits
        // into a per-thread part (low bits, and the tile bits fed by the upper bits of tid) and a constant per
        // j, and the shared-memory slot is swz(tid) ^ swz(threads * j) because the swizzle is linear.
        const uint32_t tbits = (uint32_t)P.k - 4u;                 // bits of the tile index that come from tid
        o << "__device__ __forceinline__ void spec_prefetch(double2* buf, const double2* gsrc, uint32_t tid) {\n"
             "  const uint32_t sD = swz(tid);\n"
             "  const double2* src = gsrc + ((uint64_t)(tid & " << ((1u << P.low) - 1u) << "u) | hi_off[tid >> " << P.low << "]);\n";
        for (uint32_t j = 0; j < 16; ++j) {
            const uint32_t i = j << tbits;
            o << "  cp_async16(&buf[sD ^ " << swz(i) << "u], src + " << hexu(g.hi_off[i >> P.low]) << ");\n";
        }
        o << "  asm volatile(\"cp.async.commit_group;\" ::: \"memory\");\n}\n";
        o << "extern \"C\" __global__ void __launch_bounds__(" << g.threads << ", " << ctas
          << ") spec_kernel(double2* __restrict__ state, const __grid_constant__ SpecParams SP, const SpecFactors F) {\n"
             "  extern __shared__ __align__(16) unsigned char smem_raw[];\n"
             "  double2* const buf0 = reinterpret_cast<double2*>(smem_raw);\n  double2* const buf1 = buf0 + " << (1u << P.k) << ";\n"
             "  const uint32_t tid = threadIdx.x;\n  uint64_t t = blockIdx.x;\n  if (t >= SP.n_tiles) return;\n"
             "  spec_prefetch(buf0, state + spec_tile_base(SP, t), tid);\n  int cur = 0;\n"
             "  for (; t < SP.n_tiles; t += gridDim.x, cur ^= 1) {\n"
             "  double2* const tile = cur ? buf1 : buf0;\n"
             "  asm volatile(\"cp.async.wait_all;\" ::: \"memory\");\n  __syncthreads();\n"
             "  if (t + gridDim.x < SP.n_tiles) spec_prefetch(cur ? buf0 : buf1, state + spec_tile_base(SP, t + gridDim.x), tid);\n"
             "  const uint64_t base = spec_tile_base(SP, t);\n"
             "  const uint64_t full = base | SP.shard_bits;\n  double2* gbase = state + base;\n"
             "  double2* gstore = rt_push_base(SP.push, state, base);\n  double2 a[16];\n"
          << body.str() << "  }\n}\n";
    }
    out.src = o.str();
    out.threads = g.threads;
    out.smem = (pipe || tma) ? ((size_t)32 << P.k) : (P.n_rounds > 1 ? ((size_t)16 << P.k) : 0);
    out.n_tiles = g.n_tiles;
    out.ctas_per_sm = (pipe || tma) ? ctas : 0;
    out.params = SpecParams{P.n_runs, P.shard_bits, P.run_start, P.run_len, g.n_tiles, PushSpec{}, 0};
    return true;
}

// ---- NVRTC + cudaLibrary plumbing -----------------------------------------------------------------
extern const char* const kRegtileHeaderSource;      // qvm_regtile.cuh, embedded by build.py

struct CompiledSpec {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    size_t smem_attr = 0;
};

// Process-wide cache of compiled structures.  Compilation (about a second of NVRTC per group) runs on
// background threads, so nobody waits for it: a structure keeps running through the interpreter until
// its kernel is ready.  qvm_jit_wait() drains the queue (benchmarks do that once after warm-up).
class SpecCache {
public:
    static SpecCache& instance() {
        static SpecCache* c = new SpecCache();       // never destroyed: workers may outlive static destructors
        return *c;
    }
    // nullptr: not compiled (yet) under the policy, still compiling, or compilation impossible
    // `eager`: compile at the FIRST sighting (large states: one interpreted pass costs more than the structure's
    // compile time is worth waiting for the second call).
    const CompiledSpec* get(const std::string& src, int policy, int device, bool eager = false) {
        std::unique_lock<std::mutex> lock(mu_);
        auto it = cache_.find(src);
        if (it != cache_.end()) return it->second->state == READY ? &it->second->cs : nullptr;
        if (stopping_) return nullptr;
        // bounded: a service that keeps meeting new structures stops compiling (the interpreter serves them)
        if (cache_.size() >= kMaxCompiledStructures) return nullptr;
        // a cubin of this very source on disk (an earlier process, or the ahead-of-time build): load it now,
        // the launch that asked gets the specialised kernel
        const std::string path = disk_path(src);
        if (!path.empty() && access(path.c_str(), R_OK) == 0) {
            Entry* e = new Entry();
            cache_.emplace(src, std::unique_ptr<Entry>(e));
            lock.unlock();
            std::string err;
            compile(src, e->cs, &err);
            lock.lock();
            e->state = e->cs.kernel ? READY : FAILED;
            if (!e->cs.kernel) last_error_ = err;
            return e->cs.kernel ? &e->cs : nullptr;
        }
        // an ahead-of-time job for this very source is on its way (precompile_async): its cubin will be on disk soon
        if (!path.empty() && policy < 2 && aot_queued_.count(path)) return nullptr;
        if (!path.empty() && policy < 2) {
            auto f = foreign_.find(path);
            if (f != foreign_.end()) {
                if (now_s() < f->second) return nullptr;
                foreign_.erase(f);
            }
        }
        if (seen_.size() > (1u << 20)) seen_.clear();
        if (policy < 2 && ++seen_[std::hash<std::string>()(src)] < (eager ? 1 : 2)) return nullptr;
        Entry* e = new Entry();
        cache_.emplace(src, std::unique_ptr<Entry>(e));
        ++compiles_;
        if (policy >= 2) {                          // synchronous (tests, ahead-of-time warm-up)
            lock.unlock();
            std::string err;
            compile(src, e->cs, &err);
            lock.lock();
            e->state = e->cs.kernel ? READY : FAILED;
            if (!e->cs.kernel) last_error_ = err;
            return e->cs.kernel ? &e->cs : nullptr;
        }
        jobs_.push_back(Job{src, e, device});
        ++outstanding_;
        start_workers();
        cv_.notify_one();
        return nullptr;
    }
    // A compilation of this source is queued or running: its kernel will be loadable soon.  launch_rec() waits for it
    // as long as the GPU still has earlier passes queued (a wait that costs no device time).
    bool on_its_way(const std::string& src) {
        std::lock_guard<std::mutex> lock(mu_);
        if (stopping_) return false;
        auto it = cache_.find(src);
        if (it != cache_.end()) return it->second->state == PENDING;
        const std::string path = disk_path(src);
        if (path.empty() || access(path.c_str(), R_OK) == 0) return false;
        if (aot_queued_.count(path)) return true;
        auto f = foreign_.find(path);
        if (f == foreign_.end()) return false;
        if (now_s() < f->second) return true;
        foreign_.erase(f);                 // the other process did not deliver
        return false;
    }
    bool busy() {
        std::lock_guard<std::mutex> lock(mu_);
        return outstanding_ > 0;
    }
    void defer_next(int n) {
        std::lock_guard<std::mutex> lock(mu_);
        defer_next_ = n > 0 ? n : 0;
    }
    void wait_idle() {
        std::unique_lock<std::mutex> lock(mu_);
        idle_cv_.wait(lock, [&] { return outstanding_ == 0; });
    }
    // Process exit: drop the queued jobs and wait for the compilations in flight -- a worker inside NVRTC
    // while the C++ runtime tears that library's statics down corrupts the heap.
    void shutdown() {
        std::unique_lock<std::mutex> lock(mu_);
        stopping_ = true;
        for (std::deque<Job>* jobs : {&jobs_, &late_jobs_}) {
            for (Job& job : *jobs) {
                if (job.entry) job.entry->state = FAILED;
                --outstanding_;
            }
            jobs->clear();
        }
        idle_cv_.wait(lock, [&] { return outstanding_ == 0; });
    }
    int compiles() {
        std::lock_guard<std::mutex> lock(mu_);
        return compiles_;
    }
    // {structures submitted to NVRTC, cubins loaded from the disk cache, cubins written to it, kernels ready}
    void counters(uint64_t out[4]) {
        std::lock_guard<std::mutex> lock(mu_);
        out[0] = (uint64_t)compiles_;
        out[1] = disk_hits_;
        out[2] = disk_stores_;
        uint64_t ready = 0;
        for (auto& kv : cache_) ready += kv.second->state == READY;
        out[3] = ready;
    }
    // Source -> cubin in the disk cache without a device (ahead-of-time builds); true when the cubin is there now.
    bool precompile(const std::string& src, std::string* err) {
        const std::string path = disk_path(src);
        if (path.empty()) {
            if (err) *err = "the disk cache is disabled or not writable";
            return false;
        }
        if (access(path.c_str(), R_OK) == 0) return true;
        std::vector<char> cubin;
        if (!to_cubin(src, cubin, err)) return false;
        return write_cubin(path, cubin);
    }
    // the same on the background threads (qvm_jit_wait() drains them): ahead-of-time builds of whole programs
    // `late`: behind every job queued the normal way (the first passes of a run that is about to start: they will have
    // gone through the interpreter before any compilation can finish, so the later passes' kernels come first).
    // `foreign`: another process sharing the disk cache compiles this one (ranks of one node split a program's
    // structures between them); here the cubin is only expected -- for kForeignPatienceS seconds, after which the
    // structure is compiled locally like any other.
    bool precompile_async(const std::string& src, std::string* err, bool late = false, bool foreign = false) {
        const std::string path = disk_path(src);
        if (path.empty()) {
            if (err) *err = "the disk cache is disabled or not writable";
            return false;
        }
        if (access(path.c_str(), R_OK) == 0) return true;
        std::unique_lock<std::mutex> lock(mu_);
        if (stopping_) return false;
        if (foreign) {
            if (!aot_queued_.count(path)) foreign_[path] = now_s() + kForeignPatienceS;
            return true;
        }
        foreign_.erase(path);
        if (!aot_queued_.insert(path).second) return true;          // already on its way
        if (defer_next_ > 0) {
            --defer_next_;
            late = true;
        }
        (late ? late_jobs_ : jobs_).push_back(Job{src, nullptr, -1});
        ++compiles_;
        ++outstanding_;
        start_workers();
        cv_.notify_one();
        return true;
    }
    std::string cache_dir() {
        std::lock_guard<std::mutex> lock(dir_mu_);
        return resolve_dir();
    }
    std::string last_error() {
        std::lock_guard<std::mutex> lock(mu_);
        return last_error_;
    }

    // source -> sm_100a cubin (needs no device: also used by the CPU test suite)
    bool to_cubin(const std::string& src, std::vector<char>& cubin, std::string* err) {
        if (const char* dir = getenv("QVM_B200_SPEC_DUMP")) {       // debugging aid: keep the generated sources
            char path[512];
            snprintf(path, sizeof(path), "%s/spec_%016zx.cu", dir, std::hash<std::string>()(src));
            if (FILE* fh = fopen(path, "w")) {
                fwrite(src.data(), 1, src.size(), fh);
                fclose(fh);
            }
        }
        if (!load_nvrtc(err)) return false;
        void* prog = nullptr;
        const char* headers[] = {kRegtileHeaderSource};
        const char* names[] = {"qvm_regtile.cuh"};
        if (create_(&prog, src.c_str(), "qvm_spec.cu", 1, headers, names) != 0) {
            if (err) *err = "nvrtcCreateProgram failed";
            return false;
        }
        const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-lineinfo"};
        const int rc = compile_(prog, 4, opts);
        if (rc != 0) {
            if (err) {
                size_t n = 0;
                if (log_size_ && log_ && log_size_(prog, &n) == 0 && n > 1) {
                    std::vector<char> log(n);
                    log_(prog, log.data());
                    *err = std::string("nvrtc: ") + log.data();
                } else *err = "nvrtcCompileProgram failed";
            }
            destroy_(&prog);
            return false;
        }
        size_t n = 0;
        cubin_size_(prog, &n);
        cubin.resize(n);
        cubin_(prog, cubin.data());
        destroy_(&prog);
        return n > 0;
    }

private:
    static constexpr size_t kMaxCompiledStructures = 4096;      // loaded modules kept for the life of the process
    enum { PENDING = 0, READY = 1, FAILED = 2 };
    struct Entry {
        int state = PENDING;
        CompiledSpec cs;
    };
    struct Job {
        std::string src;
        Entry* entry;
        int device;
    };
    typedef int (*fn_create)(void**, const char*, const char*, int, const char* const*, const char* const*);
    typedef int (*fn_compile)(void*, int, const char* const*);
    typedef int (*fn_size)(void*, size_t*);
    typedef int (*fn_get)(void*, char*);
    typedef int (*fn_destroy)(void**);

    void start_workers() {                       // mu_ held
        if (workers_started_) return;
        workers_started_ = true;
        load_nvrtc(nullptr);                      // before atexit: handlers run in reverse order of registration
        std::atexit([] { SpecCache::instance().shutdown(); });
        // NVRTC is the bottleneck of a first call (~0.5-1 s of front end per structure): all cores but two
        const unsigned hw = std::thread::hardware_concurrency();
        unsigned n = hw > 3 ? hw - 2 : 1;
        if (const char* e = getenv("QVM_B200_JIT_THREADS")) n = (unsigned)atoi(e);
        n = n < 1 ? 1 : (n > 32 ? 32 : n);
        for (unsigned i = 0; i < n; ++i) std::thread([this] { worker(); }).detach();
    }
    void worker() {
        for (;;) {
            Job job;
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_.wait(lock, [&] { return !jobs_.empty() || !late_jobs_.empty(); });      // (never returns once shutdown() emptied the queues)
                std::deque<Job>& from = jobs_.empty() ? late_jobs_ : jobs_;
                job = std::move(from.front());
                from.pop_front();
            }
            if (!job.entry) {                     // ahead-of-time: source -> cubin on disk, no device involved
                std::string err;
                const bool ok = precompile(job.src, &err);
                std::lock_guard<std::mutex> lock(mu_);
                if (!ok) {
                    last_error_ = err;
                    aot_queued_.erase(disk_path(job.src));        // nobody should wait for this cubin
                }
                if (--outstanding_ == 0) idle_cv_.notify_all();
                continue;
            }
            cudaSetDevice(job.device);
            CompiledSpec cs;
            std::string err;
            compile(job.src, cs, &err);
            {
                std::lock_guard<std::mutex> lock(mu_);
                job.entry->cs = cs;
                job.entry->state = cs.kernel ? READY : FAILED;
                if (!cs.kernel) last_error_ = err;
                if (--outstanding_ == 0) idle_cv_.notify_all();
            }
        }
    }
    bool load_nvrtc(std::string* err) {
        std::lock_guard<std::mutex> lock(nvrtc_mu_);
        if (tried_) {
            if (!nvrtc_ && err) *err = "libnvrtc not available";
            return nvrtc_ != nullptr;
        }
        tried_ = true;
        for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
            nvrtc_ = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (nvrtc_) break;
        }
        if (!nvrtc_) {
            if (err) *err = "libnvrtc not found";
            return false;
        }
        create_ = (fn_create)dlsym(nvrtc_, "nvrtcCreateProgram");
        compile_ = (fn_compile)dlsym(nvrtc_, "nvrtcCompileProgram");
        cubin_size_ = (fn_size)dlsym(nvrtc_, "nvrtcGetCUBINSize");
        cubin_ = (fn_get)dlsym(nvrtc_, "nvrtcGetCUBIN");
        log_size_ = (fn_size)dlsym(nvrtc_, "nvrtcGetProgramLogSize");
        log_ = (fn_get)dlsym(nvrtc_, "nvrtcGetProgramLog");
        destroy_ = (fn_destroy)dlsym(nvrtc_, "nvrtcDestroyProgram");
        if (!create_ || !compile_ || !cubin_size_ || !cubin_ || !destroy_) {
            nvrtc_ = nullptr;
            if (err) *err = "libnvrtc lacks the expected entry points";
            return false;
        }
        return true;
    }
    // ---- disk cache of cubins: <dir of libqvm_b200.so>/_spec_cache (QVM_B200_CACHE_DIR overrides, "0" disables),
    // keyed by a 128-bit hash of the embedded header, the generated source, the options and the NVRTC version
    std::string resolve_dir() {                   // dir_mu_ held
        if (dir_tried_) return dir_;
        dir_tried_ = true;
        std::string d;
        if (const char* e = getenv("QVM_B200_CACHE_DIR")) {
            if (!strcmp(e, "0") || !*e) return dir_;
            d = e;
        } else {
            Dl_info info;
            if (!dladdr((const void*)&kRegtileHeaderSource, &info) || !info.dli_fname) return dir_;
            d = info.dli_fname;
            const size_t slash = d.rfind('/');
            d = (slash == std::string::npos ? std::string(".") : d.substr(0, slash)) + "/_spec_cache";
        }
        mkdir(d.c_str(), 0755);
        if (access(d.c_str(), W_OK) != 0 && access(d.c_str(), R_OK) != 0) return dir_;
        dir_ = d;
        return dir_;
    }
    static uint64_t fnv(const char* p, size_t n, uint64_t h) {
        for (size_t i = 0; i < n; ++i) {
            h ^= (unsigned char)p[i];
            h *= 1099511628211ull;
        }
        return h;
    }
    std::string disk_path(const std::string& src) {
        std::string dir;
        {
            std::lock_guard<std::mutex> lock(dir_mu_);
            dir = resolve_dir();
        }
        if (dir.empty()) return "";
        static const std::string salt = [] {
            std::string s = "sm_100a|c++17|lineinfo|v2|";
            s += kRegtileHeaderSource;
            return s;
        }();
        uint64_t a = fnv(salt.data(), salt.size(), 14695981039346656037ull);
        a = fnv(src.data(), src.size(), a);
        uint64_t b = fnv(salt.data(), salt.size(), 0x9E3779B97F4A7C15ull);
        b = fnv(src.data(), src.size(), b);
        char name[80];
        snprintf(name, sizeof(name), "/spec_%016llx%016llx.cubin", (unsigned long long)a, (unsigned long long)b);
        return dir + name;
    }
    static bool read_cubin(const std::string& path, std::vector<char>& cubin) {
        FILE* fh = fopen(path.c_str(), "rb");
        if (!fh) return false;
        fseek(fh, 0, SEEK_END);
        const long n = ftell(fh);
        fseek(fh, 0, SEEK_SET);
        bool ok = n > 0;
        if (ok) {
            cubin.resize((size_t)n);
            ok = fread(cubin.data(), 1, (size_t)n, fh) == (size_t)n;
        }
        fclose(fh);
        return ok;
    }
    bool write_cubin(const std::string& path, const std::vector<char>& cubin) {
        char tmp[600];
        snprintf(tmp, sizeof(tmp), "%s.%d.tmp", path.c_str(), (int)getpid());
        FILE* fh = fopen(tmp, "wb");
        if (!fh) return false;
        const bool ok = fwrite(cubin.data(), 1, cubin.size(), fh) == cubin.size();
        fclose(fh);
        if (!ok || rename(tmp, path.c_str()) != 0) {
            unlink(tmp);
            return false;
        }
        std::lock_guard<std::mutex> lock(mu_count_);
        ++disk_stores_;
        return true;
    }
    void compile(const std::string& src, CompiledSpec& cs, std::string* err) {
        std::vector<char> cubin;
        const std::string path = disk_path(src);
        bool from_disk = !path.empty() && read_cubin(path, cubin);
        if (!from_disk) {
            if (!to_cubin(src, cubin, err)) return;
            if (!path.empty()) write_cubin(path, cubin);
        }
        cudaLibrary_t lib = nullptr;
        if (cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess) {
            cudaGetLastError();
            if (from_disk) {                     // a damaged file: drop it and build from source
                unlink(path.c_str());
                cubin.clear();
                from_disk = false;
                if (!to_cubin(src, cubin, err)) return;
                write_cubin(path, cubin);
                if (cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess) {
                    cudaGetLastError();
                    if (err) *err = "cudaLibraryLoadData failed";
                    return;
                }
            } else {
                if (err) *err = "cudaLibraryLoadData failed";
                return;
            }
        }
        if (from_disk) {
            std::lock_guard<std::mutex> lock(mu_count_);
            ++disk_hits_;
        }
        cudaKernel_t k = nullptr;
        if (cudaLibraryGetKernel(&k, lib, "spec_kernel") != cudaSuccess) {
            cudaGetLastError();
            cudaLibraryUnload(lib);
            if (err) *err = "spec_kernel not found in the compiled module";
            return;
        }
        cs.lib = lib;
        cs.kernel = k;
    }
    std::mutex mu_, nvrtc_mu_, dir_mu_, mu_count_;
    std::string dir_;
    bool dir_tried_ = false;
    uint64_t disk_hits_ = 0, disk_stores_ = 0;
    std::condition_variable cv_, idle_cv_;
    std::unordered_map<std::string, std::unique_ptr<Entry>> cache_;
    std::unordered_map<size_t, int> seen_;
    std::deque<Job> jobs_, late_jobs_;
    int defer_next_ = 0;
    static constexpr double kForeignPatienceS = 20.0;
    std::unordered_map<std::string, double> foreign_;       // cubin path -> until when another process is trusted with it
    static double now_s() {
        return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
    std::unordered_set<std::string> aot_queued_;      // disk paths of ahead-of-time jobs already queued
    int outstanding_ = 0, compiles_ = 0;
    bool workers_started_ = false, stopping_ = false;
    std::string last_error_;
    void* nvrtc_ = nullptr;
    bool tried_ = false;
    fn_create create_ = nullptr;
    fn_compile compile_ = nullptr;
    fn_size cubin_size_ = nullptr, log_size_ = nullptr;
    fn_get cubin_ = nullptr, log_ = nullptr;
    fn_destroy destroy_ = nullptr;
};

}  // namespace qvm
