#include "qvm_regtile.cuh"
using namespace qvm;
struct SpecParams { int n_runs; uint64_t shard_bits, run_start, run_len, n_tiles; PushSpec push; int from_zero; };
struct SpecFactors { double v[394]; };
__device__ const uint64_t hi_off[128] = {0x0ull, 0x200ull, 0x400ull, 0x600ull, 0x20000ull, 0x20200ull, 0x20400ull, 0x20600ull, 0x100000ull, 0x100200ull, 0x100400ull, 0x100600ull, 0x120000ull, 0x120200ull, 0x120400ull, 0x120600ull, 0x2000000ull, 0x2000200ull, 0x2000400ull, 0x2000600ull, 0x2020000ull, 0x2020200ull, 0x2020400ull, 0x2020600ull, 0x2100000ull, 0x2100200ull, 0x2100400ull, 0x2100600ull, 0x2120000ull, 0x2120200ull, 0x2120400ull, 0x2120600ull, 0x8000000ull, 0x8000200ull, 0x8000400ull, 0x8000600ull, 0x8020000ull, 0x8020200ull, 0x8020400ull, 0x8020600ull, 0x8100000ull, 0x8100200ull, 0x8100400ull, 0x8100600ull, 0x8120000ull, 0x8120200ull, 0x8120400ull, 0x8120600ull, 0xa000000ull, 0xa000200ull, 0xa000400ull, 0xa000600ull, 0xa020000ull, 0xa020200ull, 0xa020400ull, 0xa020600ull, 0xa100000ull, 0xa100200ull, 0xa100400ull, 0xa100600ull, 0xa120000ull, 0xa120200ull, 0xa120400ull, 0xa120600ull, 0x20000000ull, 0x20000200ull, 0x20000400ull, 0x20000600ull, 0x20020000ull, 0x20020200ull, 0x20020400ull, 0x20020600ull, 0x20100000ull, 0x20100200ull, 0x20100400ull, 0x20100600ull, 0x20120000ull, 0x20120200ull, 0x20120400ull, 0x20120600ull, 0x22000000ull, 0x22000200ull, 0x22000400ull, 0x22000600ull, 0x22020000ull, 0x22020200ull, 0x22020400ull, 0x22020600ull, 0x22100000ull, 0x22100200ull, 0x22100400ull, 0x22100600ull, 0x22120000ull, 0x22120200ull, 0x22120400ull, 0x22120600ull, 0x28000000ull, 0x28000200ull, 0x28000400ull, 0x28000600ull, 0x28020000ull, 0x28020200ull, 0x28020400ull, 0x28020600ull, 0x28100000ull, 0x28100200ull, 0x28100400ull, 0x28100600ull, 0x28120000ull, 0x28120200ull, 0x28120400ull, 0x28120600ull, 0x2a000000ull, 0x2a000200ull, 0x2a000400ull, 0x2a000600ull, 0x2a020000ull, 0x2a020200ull, 0x2a020400ull, 0x2a020600ull, 0x2a100000ull, 0x2a100200ull, 0x2a100400ull, 0x2a100600ull, 0x2a120000ull, 0x2a120200ull, 0x2a120400ull, 0x2a120600ull};
__device__ __forceinline__ uint64_t spec_tile_base(const SpecParams& SP, uint64_t t) {
  uint64_t base = 0;
  for (int r = 0; r < SP.n_runs; ++r) {
    const uint32_t len = (uint32_t)(SP.run_len >> (8 * r)) & 0xFFu;
    base |= (t & ((1ull << len) - 1ull)) << ((uint32_t)(SP.run_start >> (8 * r)) & 0xFFu);
    t >>= len;
  }
  return base ^ SP.push.rot;
}
extern "C" __global__ void __launch_bounds__(128, 4) spec_kernel(double2* __restrict__ state, const __grid_constant__ SpecParams SP, const SpecFactors F) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* tile = reinterpret_cast<double2*>(smem_raw);
  const uint32_t tid = threadIdx.x;
  const uint64_t base = spec_tile_base(SP, blockIdx.x);
  const uint64_t full = base | SP.shard_bits;
  double2* gbase = state + base;
  double2* gstore = rt_push_base(SP.push, state, base);
  double2 a[16];
  if (SP.from_zero && base != 0) {
    uint32_t T = tid;
    T += T & 0xffc0u;
    T += T & 0xff00u;
    T += T & 0xfe00u;
    T += T & 0xfc00u;
    const uint64_t gT = (uint64_t)(T & 15u) | hi_off[T >> 4];
    const int64_t b[4] = {(int64_t)(16ull << 29), (int64_t)(16ull << 27), (int64_t)(16ull << 25), (int64_t)(16ull << 17)};
    char* q[16];
    add_tree<4>(reinterpret_cast<char*>(gstore + gT), b, q);
#pragma unroll
    for (int j = 0; j < 16; ++j) *reinterpret_cast<double2*>(q[j]) = make_double2(0.0, 0.0);
    return;
  }
  {  // ---- round 0
    uint32_t T = tid;
    T += T & 0xffc0u;
    T += T & 0xff00u;
    T += T & 0xfe00u;
    T += T & 0xfc00u;
    const uint32_t sT = swz(T);
    const uint32_t st[4] = {1031u, 517u, 71u, 262u};
    if (SP.from_zero) {
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = make_double2(0.0, 0.0);
      if (SP.from_zero == 2 && base == 0 && tid == 0) a[0].x = 1.0;
    } else
    {
      const uint64_t gT = (uint64_t)(T & 15u) | hi_off[T >> 4];
      const int64_t b[4] = {(int64_t)(16ull << 29), (int64_t)(16ull << 27), (int64_t)(16ull << 17), (int64_t)(16ull << 25)};
      const char* q[16];
      add_tree<4>(reinterpret_cast<const char*>(gbase + gT), b, q);
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = *reinterpret_cast<const double2*>(q[j]);
    }
    double2 ts = make_double2(F.v[310], F.v[311]);
    if (((full & 0x20ull) == 0x20ull) && ((tid & 0x20u) == 0x20u)) cmul_ip(ts, make_double2(F.v[312], F.v[313]));
#pragma unroll
    for (int j = 0; j < 16; ++j) cmul_ip(a[j], ts);
    uint32_t xm = 0;
    reg_h<0, 16>(a, xm);
    if (((full & 0x8000ull) == 0x8000ull)) xm ^= 2u;
    if (((full & 0x2000ull) == 0x2000ull)) xm ^= 4u;
    if (((full & 0x100ull) == 0x100ull)) xm ^= 8u;
    if (((tid & 0x40u) == 0x40u)) reg_s1<1, 16>(a, make_double2(F.v[314], F.v[315]), xm);
    reg_h<1, 16>(a, xm);
    reg_s1<2, 16>(a, make_double2(F.v[316], F.v[317]), xm);
    if (((full & 0x40ull) == 0x40ull)) xm ^= 4u;
    reg_s1<3, 16>(a, make_double2(F.v[318], F.v[319]), xm);
    reg_h<3, 16>(a, xm);
    reg_h<2, 16>(a, xm);
    reg_h<3, 16>(a, xm);
    {
      reg_xswap<3, 16>(a, 1u, xm);
    }
    uint32_t sX = 0u;
    sX ^= (0u - ((T >> 0) & 1u)) & 1u;
    sX ^= (0u - ((T >> 1) & 1u)) & 2u;
    sX ^= (0u - ((T >> 2) & 1u)) & 4u;
    sX ^= (0u - ((T >> 3) & 1u)) & 132u;
    sX ^= (0u - ((T >> 4) & 1u)) & 22u;
    sX ^= (0u - ((T >> 5) & 1u)) & 37u;
    sX ^= (0u - ((T >> 6) & 1u)) & 71u;
    sX ^= (0u - ((T >> 7) & 1u)) & 11u;
    sX ^= (0u - ((T >> 8) & 1u)) & 262u;
    sX ^= (0u - ((T >> 9) & 1u)) & 517u;
    sX ^= (0u - ((T >> 10) & 1u)) & 1031u;
    const uint32_t wst[4] = {1031u, 517u, 71u, 262u};
#pragma unroll
    for (int s = 0; s < 4; ++s) if ((xm >> s) & 1u) sX ^= wst[s];
#pragma unroll
    for (int j = 0; j < 16; ++j) tile[sX ^ xor_combo<4>(wst, j)] = a[j];
  }
  __syncthreads();
  {  // ---- round 1
    uint32_t T = tid;
    T += T & 0xfffcu;
    T += T & 0xfff0u;
    T += T & 0xffe0u;
    T += T & 0xff80u;
    const uint32_t sT = swz(T);
    const uint32_t st[4] = {37u, 22u, 132u, 4u};
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = tile[sT ^ xor_combo<4>(st, j)];
    double2 ts = make_double2(F.v[320], F.v[321]);
    if (((tid & 0x5u) == 0x5u)) cmul_ip(ts, make_double2(F.v[322], F.v[323]));
    cmul_ip(ts, make_double2(F.v[2 * (13u + ((((uint32_t)(full >> 4) & 1u) << 0)))], F.v[2 * (13u + ((((uint32_t)(full >> 4) & 1u) << 0))) + 1]));
    if (((full & 0x80000040ull) == 0x80000040ull)) cmul_ip(ts, make_double2(F.v[324], F.v[325]));
    cmul_ip(ts, make_double2(F.v[2 * (29u + ((((uint32_t)(full >> 11) & 1u) << 0)))], F.v[2 * (29u + ((((uint32_t)(full >> 11) & 1u) << 0))) + 1]));
    if (((full & 0x5000ull) == 0x5000ull)) cmul_ip(ts, make_double2(F.v[326], F.v[327]));
    if (((full & 0x10080ull) == 0x10080ull)) cmul_ip(ts, make_double2(F.v[328], F.v[329]));
    cmul_ip(ts, make_double2(F.v[2 * (44u + ((((uint32_t)(full >> 19) & 1u) << 0)))], F.v[2 * (44u + ((((uint32_t)(full >> 19) & 1u) << 0))) + 1]));
    if (((full & 0x200080ull) == 0x200080ull)) cmul_ip(ts, make_double2(F.v[330], F.v[331]));
    if (((full & 0x404000ull) == 0x404000ull)) cmul_ip(ts, make_double2(F.v[332], F.v[333]));
    if (((full & 0x4000020ull) == 0x4000020ull)) cmul_ip(ts, make_double2(F.v[334], F.v[335]));
    cmul_ip(ts, make_double2(F.v[2 * (56u + ((((uint32_t)(full >> 28) & 1u) << 0)))], F.v[2 * (56u + ((((uint32_t)(full >> 28) & 1u) << 0))) + 1]));
    cmul_ip(ts, make_double2(F.v[2 * (62u + ((((uint32_t)(full >> 30) & 1u) << 0)))], F.v[2 * (62u + ((((uint32_t)(full >> 30) & 1u) << 0))) + 1]));
    if (((full & 0x30ull) == 0x30ull)) cmul_ip(ts, make_double2(F.v[336], F.v[337]));
    cmul_ip(ts, make_double2(F.v[2 * (78u + ((((uint32_t)(full >> 12) & 1u) << 0)))], F.v[2 * (78u + ((((uint32_t)(full >> 12) & 1u) << 0))) + 1]));
    if (((full & 0x100000010ull) == 0x100000010ull)) cmul_ip(ts, make_double2(F.v[338], F.v[339]));
    cmul_ip(ts, make_double2(F.v[2 * (105u + ((((uint32_t)(full >> 5) & 1u) << 0)))], F.v[2 * (105u + ((((uint32_t)(full >> 5) & 1u) << 0))) + 1]));
    if (((full & 0x8100ull) == 0x8100ull)) cmul_ip(ts, make_double2(F.v[340], F.v[341]));
    if (((full & 0x4800ull) == 0x4800ull)) cmul_ip(ts, make_double2(F.v[342], F.v[343]));
    if (((full & 0x110ull) == 0x110ull)) cmul_ip(ts, make_double2(F.v[344], F.v[345]));
    if (((tid & 0x48u) == 0x48u)) cmul_ip(ts, make_double2(F.v[346], F.v[347]));
    cmul_ip(ts, (((tid >> 4) & 1u) ? make_double2(F.v[350], F.v[351]) : make_double2(F.v[348], F.v[349])));
    cmul_ip(ts, (((tid >> 5) & 1u) ? make_double2(F.v[354], F.v[355]) : make_double2(F.v[352], F.v[353])));
    cmul_ip(ts, (((tid >> 3) & 1u) ? make_double2(F.v[358], F.v[359]) : make_double2(F.v[356], F.v[357])));
#pragma unroll
    for (int j = 0; j < 16; ++j) cmul_ip(a[j], ts);
    uint32_t xm = 0;
    reg_h<0, 16>(a, xm);
    reg_h<1, 16>(a, xm);
    reg_h<2, 16>(a, xm);
    reg_h<3, 16>(a, xm);
    reg_h<0, 16>(a, xm);
    if (((full & 0x40000ull) == 0x40000ull)) reg_s1<3, 16>(a, make_double2(F.v[360], F.v[361]), xm);
    {
      reg_xswap<3, 16>(a, 4u, xm);
    }
    {
      const double2 f = make_double2(F.v[362], F.v[363]);
      const uint32_t want = 10u & ~xm;
#pragma unroll
      for (int p = 0; p < 16; ++p) if ((p & 10u) == want) cmul_ip(a[p], f);
    }
    reg_s1<0, 16>(a, make_double2(F.v[364], F.v[365]), xm);
    if (((tid & 0x20u) == 0x20u)) xm ^= 2u;
    reg_h<2, 16>(a, xm);
    reg_h<3, 16>(a, xm);
    {
      const double2 f = make_double2(F.v[366], F.v[367]);
      const uint32_t want = 5u & ~xm;
#pragma unroll
      for (int p = 0; p < 16; ++p) if ((p & 5u) == want) cmul_ip(a[p], f);
    }
    {
      reg_xswap<0, 16>(a, 2u, xm);
    }
    uint32_t sX = sT;
#pragma unroll
    for (int s = 0; s < 4; ++s) if ((xm >> s) & 1u) sX ^= st[s];
#pragma unroll
    for (int j = 0; j < 16; ++j) tile[sX ^ xor_combo<4>(st, j)] = a[j];
  }
  __syncthreads();
  {  // ---- round 2
    uint32_t T = tid;
    T += T & 0xfffeu;
    T += T & 0xfff8u;
    T += T & 0xfe00u;
    T += T & 0xfc00u;
    const uint32_t sT = swz(T);
    const uint32_t st[4] = {1031u, 517u, 11u, 2u};
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = tile[sT ^ xor_combo<4>(st, j)];
    uint32_t xm = 0;
    reg_h<3, 16>(a, xm);
    {
      reg_xswap<2, 16>(a, 8u, xm);
    }
    if (((full & 0x80000ull) == 0x80000ull)) xm ^= 8u;
    reg_h<3, 16>(a, xm);
    uint32_t sX = sT;
#pragma unroll
    for (int s = 0; s < 4; ++s) if ((xm >> s) & 1u) sX ^= st[s];
#pragma unroll
    for (int j = 0; j < 16; ++j) tile[sX ^ xor_combo<4>(st, j)] = a[j];
  }
  __syncthreads();
  {  // ---- round 3
    uint32_t T = tid;
    T += T & 0xffc0u;
    T += T & 0xff00u;
    T += T & 0xfe00u;
    T += T & 0xfc00u;
    const uint32_t sT = swz(T);
    const uint32_t st[4] = {1031u, 517u, 262u, 71u};
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = tile[sT ^ xor_combo<4>(st, j)];
    double2 ts = make_double2(F.v[368], F.v[369]);
    if (((full & 0x10ull) == 0x10ull) && ((tid & 0x10u) == 0x10u)) cmul_ip(ts, make_double2(F.v[370], F.v[371]));
    cmul_ip(ts, (((tid >> 3) & 1u) ? make_double2(F.v[374], F.v[375]) : make_double2(F.v[372], F.v[373])));
    cmul_ip(ts, (((tid >> 2) & 1u) ? make_double2(F.v[378], F.v[379]) : make_double2(F.v[376], F.v[377])));
    cmul_ip(ts, (((tid >> 1) & 1u) ? make_double2(F.v[382], F.v[383]) : make_double2(F.v[380], F.v[381])));
    cmul_ip(ts, (((tid >> 4) & 1u) ? make_double2(F.v[386], F.v[387]) : make_double2(F.v[384], F.v[385])));
    if (((tid & 0x30u) == 0x30u)) cmul_ip(ts, make_double2(F.v[388], F.v[389]));
    cmul_ip(ts, (((tid >> 5) & 1u) ? make_double2(F.v[392], F.v[393]) : make_double2(F.v[390], F.v[391])));
#pragma unroll
    for (int j = 0; j < 16; ++j) cmul_ip(a[j], ts);
    uint32_t xm = 0;
    {
      const uint64_t gT = (uint64_t)(T & 15u) | hi_off[T >> 4];
      char* p0 = reinterpret_cast<char*>(gstore + gT);
      int64_t b[4] = {(int64_t)(16ull << 29), (int64_t)(16ull << 27), (int64_t)(16ull << 25), (int64_t)(16ull << 17)};
#pragma unroll
      for (int s = 0; s < 4; ++s) if ((xm >> s) & 1u) { p0 += b[s]; b[s] = -b[s]; }
      char* q[16];
      add_tree<4>(p0, b, q);
#pragma unroll
      for (int j = 0; j < 16; ++j) *reinterpret_cast<double2*>(q[j]) = a[j];
    }
  }
}
