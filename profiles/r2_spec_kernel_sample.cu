#include "qvm_regtile.cuh"
using namespace qvm;
struct SpecParams { int n_runs; uint64_t shard_bits, run_start, run_len, n_tiles; PushSpec push; };
struct SpecFactors { double v[582]; };
__device__ const uint64_t hi_off[128] = {0x0ull, 0x1000ull, 0x2000ull, 0x3000ull, 0x10000ull, 0x11000ull, 0x12000ull, 0x13000ull, 0x20000ull, 0x21000ull, 0x22000ull, 0x23000ull, 0x30000ull, 0x31000ull, 0x32000ull, 0x33000ull, 0x40000ull, 0x41000ull, 0x42000ull, 0x43000ull, 0x50000ull, 0x51000ull, 0x52000ull, 0x53000ull, 0x60000ull, 0x61000ull, 0x62000ull, 0x63000ull, 0x70000ull, 0x71000ull, 0x72000ull, 0x73000ull, 0x400000ull, 0x401000ull, 0x402000ull, 0x403000ull, 0x410000ull, 0x411000ull, 0x412000ull, 0x413000ull, 0x420000ull, 0x421000ull, 0x422000ull, 0x423000ull, 0x430000ull, 0x431000ull, 0x432000ull, 0x433000ull, 0x440000ull, 0x441000ull, 0x442000ull, 0x443000ull, 0x450000ull, 0x451000ull, 0x452000ull, 0x453000ull, 0x460000ull, 0x461000ull, 0x462000ull, 0x463000ull, 0x470000ull, 0x471000ull, 0x472000ull, 0x473000ull, 0x2000000ull, 0x2001000ull, 0x2002000ull, 0x2003000ull, 0x2010000ull, 0x2011000ull, 0x2012000ull, 0x2013000ull, 0x2020000ull, 0x2021000ull, 0x2022000ull, 0x2023000ull, 0x2030000ull, 0x2031000ull, 0x2032000ull, 0x2033000ull, 0x2040000ull, 0x2041000ull, 0x2042000ull, 0x2043000ull, 0x2050000ull, 0x2051000ull, 0x2052000ull, 0x2053000ull, 0x2060000ull, 0x2061000ull, 0x2062000ull, 0x2063000ull, 0x2070000ull, 0x2071000ull, 0x2072000ull, 0x2073000ull, 0x2400000ull, 0x2401000ull, 0x2402000ull, 0x2403000ull, 0x2410000ull, 0x2411000ull, 0x2412000ull, 0x2413000ull, 0x2420000ull, 0x2421000ull, 0x2422000ull, 0x2423000ull, 0x2430000ull, 0x2431000ull, 0x2432000ull, 0x2433000ull, 0x2440000ull, 0x2441000ull, 0x2442000ull, 0x2443000ull, 0x2450000ull, 0x2451000ull, 0x2452000ull, 0x2453000ull, 0x2460000ull, 0x2461000ull, 0x2462000ull, 0x2463000ull, 0x2470000ull, 0x2471000ull, 0x2472000ull, 0x2473000ull};
__device__ __forceinline__ uint64_t spec_tile_base(const SpecParams& SP, uint64_t t) {
  uint64_t base = 0;
  for (int r = 0; r < SP.n_runs; ++r) {
    const uint32_t len = (uint32_t)(SP.run_len >> (8 * r)) & 0xFFu;
    base |= (t & ((1ull << len) - 1ull)) << ((uint32_t)(SP.run_start >> (8 * r)) & 0xFFu);
    t >>= len;
  }
  return base;
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
  {  // ---- round 0
    uint32_t T = tid;
    T += T & 0xffe0u;
    T += T & 0xffc0u;
    T += T & 0xff80u;
    T += T & 0xfc00u;
    const uint32_t sT = swz(T);
    const uint32_t st[4] = {1031u, 71u, 132u, 37u};
    {
      const uint64_t gT = (uint64_t)(T & 15u) | hi_off[T >> 4];
      const int64_t b[4] = {(int64_t)(16ull << 25), (int64_t)(16ull << 16), (int64_t)(16ull << 17), (int64_t)(16ull << 13)};
      const char* q[16];
      add_tree<4>(reinterpret_cast<const char*>(gbase + gT), b, q);
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = *reinterpret_cast<const double2*>(q[j]);
    }
    uint32_t xm = 0;
    if (((tid & 0x4u) == 0x4u)) xm ^= 1u;
    if (((tid & 0x8u) == 0x8u)) xm ^= 2u;
    if (((full & 0x100000000ull) == 0x100000000ull)) xm ^= 4u;
    reg_h<3, 16>(a, xm);
    if (((full & 0x200ull) == 0x200ull)) reg_s1<1, 16>(a, make_double2(F.v[478], F.v[479]), xm);
    if (((tid & 0x10u) == 0x10u)) reg_s1<2, 16>(a, make_double2(F.v[480], F.v[481]), xm);
    reg_h<3, 16>(a, xm);
    reg_s1<1, 16>(a, make_double2(F.v[482], F.v[483]), xm);
    if (((tid & 0x1u) == 0x1u)) reg_s1<2, 16>(a, make_double2(F.v[484], F.v[485]), xm);
    if (((full & 0x80000000ull) == 0x80000000ull)) xm ^= 4u;
    if (((full & 0x4000ull) == 0x4000ull)) reg_s1<1, 16>(a, make_double2(F.v[486], F.v[487]), xm);
    reg_h<1, 16>(a, xm);
    if (((full & 0x100000ull) == 0x100000ull)) xm ^= 4u;
    if (((full & 0x80000000ull) == 0x80000000ull)) reg_s1<1, 16>(a, make_double2(F.v[488], F.v[489]), xm);
    if (((full & 0x4000ull) == 0x4000ull)) reg_s1<2, 16>(a, make_double2(F.v[490], F.v[491]), xm);
    {
      reg_xswap<2, 16>(a, 8u, xm);
    }
    reg_s1<1, 16>(a, make_double2(F.v[492], F.v[493]), xm);
    if (((full & 0x800ull) == 0x800ull)) reg_s1<1, 16>(a, make_double2(F.v[494], F.v[495]), xm);
    if (((full & 0x80000ull) == 0x80000ull)) reg_s1<1, 16>(a, make_double2(F.v[496], F.v[497]), xm);
    reg_h<1, 16>(a, xm);
    uint32_t sX = sT;
#pragma unroll
    for (int s = 0; s < 4; ++s) if ((xm >> s) & 1u) sX ^= st[s];
#pragma unroll
    for (int j = 0; j < 16; ++j) tile[sX ^ xor_combo<4>(st, j)] = a[j];
  }
  __syncthreads();
  {  // ---- round 1
    uint32_t T = tid;
    T += T & 0xffffu;
    T += T & 0xfffcu;
    T += T & 0xfff8u;
    T += T & 0xff00u;
    const uint32_t sT = swz(T);
    const uint32_t st[4] = {262u, 1u, 4u, 11u};
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = tile[sT ^ xor_combo<4>(st, j)];
    double2 ts = make_double2(F.v[498], F.v[499]);
    if (((tid & 0x41u) == 0x41u)) cmul_ip(ts, make_double2(F.v[500], F.v[501]));
    if (((full & 0x100000ull) == 0x100000ull) && ((tid & 0x4u) == 0x4u)) cmul_ip(ts, make_double2(F.v[502], F.v[503]));
    cmul_ip(ts, (((tid >> 4) & 1u) ? make_double2(F.v[506], F.v[507]) : make_double2(F.v[504], F.v[505])));
    if (((full & 0x400ull) == 0x400ull) && ((tid & 0x10u) == 0x10u)) cmul_ip(ts, make_double2(F.v[508], F.v[509]));
#pragma unroll
    for (int j = 0; j < 16; ++j) cmul_ip(a[j], ts);
    uint32_t xm = 0;
    {
      reg_xswap<2, 16>(a, 1u, xm);
    }
    if (((tid & 0x20u) == 0x20u)) xm ^= 1u;
    reg_h<1, 16>(a, xm);
    reg_h<2, 16>(a, xm);
    if (((full & 0x10ull) == 0x10ull)) xm ^= 8u;
    reg_h<1, 16>(a, xm);
    if (((full & 0x80000ull) == 0x80000ull)) xm ^= 2u;
    if (((full & 0x8000ull) == 0x8000ull)) reg_s1<2, 16>(a, make_double2(F.v[510], F.v[511]), xm);
    reg_h<2, 16>(a, xm);
    if (((full & 0x40000000ull) == 0x40000000ull)) xm ^= 8u;
    if (((full & 0x20000000ull) == 0x20000000ull)) reg_s1<2, 16>(a, make_double2(F.v[512], F.v[513]), xm);
    if (((full & 0x800000ull) == 0x800000ull)) xm ^= 4u;
    reg_h<3, 16>(a, xm);
    if (((full & 0x8000000ull) == 0x8000000ull)) xm ^= 8u;
    reg_h<2, 16>(a, xm);
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
    T += T & 0xfff0u;
    T += T & 0xff00u;
    T += T & 0xfe00u;
    const uint32_t sT = swz(T);
    const uint32_t st[4] = {262u, 517u, 22u, 2u};
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = tile[sT ^ xor_combo<4>(st, j)];
    double2 ts = make_double2(F.v[514], F.v[515]);
    if (((tid & 0x44u) == 0x44u)) cmul_ip(ts, make_double2(F.v[516], F.v[517]));
#pragma unroll
    for (int j = 0; j < 16; ++j) cmul_ip(a[j], ts);
    uint32_t xm = 0;
    {
      reg_xswap<1, 16>(a, 1u, xm);
    }
    if (((full & 0x1000000ull) == 0x1000000ull)) xm ^= 1u;
    if (((full & 0x80ull) == 0x80ull)) reg_s1<1, 16>(a, make_double2(F.v[518], F.v[519]), xm);
    reg_h<2, 16>(a, xm);
    reg_h<3, 16>(a, xm);
    if (((full & 0x100ull) == 0x100ull)) xm ^= 8u;
    reg_s1<1, 16>(a, make_double2(F.v[520], F.v[521]), xm);
    reg_h<1, 16>(a, xm);
    if (((full & 0x10000000ull) == 0x10000000ull)) reg_s1<2, 16>(a, make_double2(F.v[522], F.v[523]), xm);
    reg_h<1, 16>(a, xm);
    if (((full & 0x800000ull) == 0x800000ull)) reg_s1<2, 16>(a, make_double2(F.v[524], F.v[525]), xm);
    {
      const double2 f = make_double2(F.v[526], F.v[527]);
      const uint32_t want = 5u & ~xm;
#pragma unroll
      for (int p = 0; p < 16; ++p) if ((p & 5u) == want) cmul_ip(a[p], f);
    }
    if (((full & 0x20ull) == 0x20ull)) reg_s1<0, 16>(a, make_double2(F.v[528], F.v[529]), xm);
    reg_h<0, 16>(a, xm);
    reg_s1<2, 16>(a, make_double2(F.v[530], F.v[531]), xm);
    reg_h<2, 16>(a, xm);
    {
      const double2 f = make_double2(F.v[532], F.v[533]);
      const uint32_t want = 6u & ~xm;
#pragma unroll
      for (int p = 0; p < 16; ++p) if ((p & 6u) == want) cmul_ip(a[p], f);
    }
    reg_s1<0, 16>(a, make_double2(F.v[534], F.v[535]), xm);
    reg_h<0, 16>(a, xm);
    reg_h<2, 16>(a, xm);
    reg_h<0, 16>(a, xm);
    uint32_t sX = sT;
#pragma unroll
    for (int s = 0; s < 4; ++s) if ((xm >> s) & 1u) sX ^= st[s];
#pragma unroll
    for (int j = 0; j < 16; ++j) tile[sX ^ xor_combo<4>(st, j)] = a[j];
  }
  __syncthreads();
  {  // ---- round 3
    uint32_t T = tid;
    T += T & 0xfffeu;
    T += T & 0xfff8u;
    T += T & 0xffe0u;
    T += T & 0xfc00u;
    const uint32_t sT = swz(T);
    const uint32_t st[4] = {2u, 1031u, 11u, 37u};
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = tile[sT ^ xor_combo<4>(st, j)];
    double2 ts = make_double2(F.v[536], F.v[537]);
    cmul_ip(ts, (((tid >> 6) & 1u) ? make_double2(F.v[540], F.v[541]) : make_double2(F.v[538], F.v[539])));
    if (((full & 0x80000ull) == 0x80000ull) && ((tid & 0x20u) == 0x20u)) cmul_ip(ts, make_double2(F.v[542], F.v[543]));
    cmul_ip(ts, (((tid >> 5) & 1u) ? make_double2(F.v[546], F.v[547]) : make_double2(F.v[544], F.v[545])));
#pragma unroll
    for (int j = 0; j < 16; ++j) cmul_ip(a[j], ts);
    uint32_t xm = 0;
    {
      reg_xswap<3, 16>(a, 4u, xm);
    }
    {
      reg_xswap<1, 16>(a, 1u, xm);
    }
    if (((full & 0x10ull) == 0x10ull)) xm ^= 1u;
    if (((full & 0x40ull) == 0x40ull)) xm ^= 2u;
    if (((full & 0x200ull) == 0x200ull)) xm ^= 4u;
    if (((tid & 0x1u) == 0x1u)) reg_s1<3, 16>(a, make_double2(F.v[548], F.v[549]), xm);
    reg_s1<1, 16>(a, make_double2(F.v[550], F.v[551]), xm);
    reg_h<1, 16>(a, xm);
    reg_s1<2, 16>(a, make_double2(F.v[552], F.v[553]), xm);
    if (((full & 0x4000000ull) == 0x4000000ull)) xm ^= 4u;
    reg_s1<3, 16>(a, make_double2(F.v[554], F.v[555]), xm);
    reg_h<3, 16>(a, xm);
    uint32_t sX = sT;
#pragma unroll
    for (int s = 0; s < 4; ++s) if ((xm >> s) & 1u) sX ^= st[s];
#pragma unroll
    for (int j = 0; j < 16; ++j) tile[sX ^ xor_combo<4>(st, j)] = a[j];
  }
  __syncthreads();
  {  // ---- round 4
    uint32_t T = tid;
    T += T & 0xffffu;
    T += T & 0xff00u;
    T += T & 0xfe00u;
    T += T & 0xfc00u;
    const uint32_t sT = swz(T);
    const uint32_t st[4] = {1031u, 517u, 262u, 1u};
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = tile[sT ^ xor_combo<4>(st, j)];
    double2 ts = make_double2(F.v[556], F.v[557]);
    if (((tid & 0x44u) == 0x44u)) cmul_ip(ts, make_double2(F.v[558], F.v[559]));
#pragma unroll
    for (int j = 0; j < 16; ++j) cmul_ip(a[j], ts);
    uint32_t xm = 0;
    reg_h<3, 16>(a, xm);
    uint32_t sX = sT;
#pragma unroll
    for (int s = 0; s < 4; ++s) if ((xm >> s) & 1u) sX ^= st[s];
#pragma unroll
    for (int j = 0; j < 16; ++j) tile[sX ^ xor_combo<4>(st, j)] = a[j];
  }
  __syncthreads();
  {  // ---- round 5
    uint32_t T = tid;
    T += T & 0xff80u;
    T += T & 0xff00u;
    T += T & 0xfe00u;
    T += T & 0xfc00u;
    const uint32_t sT = swz(T);
    const uint32_t st[4] = {1031u, 517u, 262u, 132u};
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = tile[sT ^ xor_combo<4>(st, j)];
    double2 ts = make_double2(F.v[560], F.v[561]);
    cmul_ip(ts, (((tid >> 0) & 1u) ? make_double2(F.v[564], F.v[565]) : make_double2(F.v[562], F.v[563])));
    if (((tid & 0x6u) == 0x6u)) cmul_ip(ts, make_double2(F.v[566], F.v[567]));
    cmul_ip(ts, (((tid >> 3) & 1u) ? make_double2(F.v[570], F.v[571]) : make_double2(F.v[568], F.v[569])));
    if (((full & 0x800ull) == 0x800ull) && ((tid & 0x4u) == 0x4u)) cmul_ip(ts, make_double2(F.v[572], F.v[573]));
    cmul_ip(ts, (((tid >> 5) & 1u) ? make_double2(F.v[576], F.v[577]) : make_double2(F.v[574], F.v[575])));
#pragma unroll
    for (int j = 0; j < 16; ++j) cmul_ip(a[j], ts);
    uint32_t xm = 0;
    if (((full & 0x800ull) == 0x800ull)) reg_s1<0, 16>(a, make_double2(F.v[578], F.v[579]), xm);
    if (((full & 0x800000ull) == 0x800000ull)) reg_s1<0, 16>(a, make_double2(F.v[580], F.v[581]), xm);
    {
      const uint64_t gT = (uint64_t)(T & 15u) | hi_off[T >> 4];
      char* p0 = reinterpret_cast<char*>(gstore + gT);
      int64_t b[4] = {(int64_t)(16ull << 25), (int64_t)(16ull << 22), (int64_t)(16ull << 18), (int64_t)(16ull << 17)};
#pragma unroll
      for (int s = 0; s < 4; ++s) if ((xm >> s) & 1u) { p0 += b[s]; b[s] = -b[s]; }
      char* q[16];
      add_tree<4>(p0, b, q);
#pragma unroll
      for (int j = 0; j < 16; ++j) *reinterpret_cast<double2*>(q[j]) = a[j];
    }
  }
}
