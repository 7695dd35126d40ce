// NVLink peer-memory probe (2 GPUs, one process): what do SM-issued remote reads, remote writes and the
// in-place swap used by exchange_block_kernel reach, next to cudaMemcpyPeer?  Not part of the library.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o p2p_probe p2p_probe.cu && ./p2p_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int U>
__global__ void copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (; j + (U - 1) * stride < n; j += U * stride) {
        double2 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = __ldcs(&src[j + u * stride]);
#pragma unroll
        for (int u = 0; u < U; ++u) __stcs(&dst[j + u * stride], x[u]);
    }
}
template <int U>
__global__ void swap_kernel(double2* __restrict__ a, double2* __restrict__ b, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (; j + (U - 1) * stride < n; j += U * stride) {
        double2 x[U], y[U];
#pragma unroll
        for (int u = 0; u < U; ++u) y[u] = __ldcs(&b[j + u * stride]);
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = __ldcs(&a[j + u * stride]);
#pragma unroll
        for (int u = 0; u < U; ++u) __stcs(&a[j + u * stride], y[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) __stcs(&b[j + u * stride], x[u]);
    }
}

// 256-bit stores (sm_100: LDG/STG.E.ENL2.256): 32 bytes per lane, 1 KiB per warp instruction
template <int U>
__global__ void copy256_kernel(const double* __restrict__ src, double* __restrict__ dst, uint64_t n_quads) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (; j + (U - 1) * stride < n_quads; j += U * stride) {
        double a[U], b[U], c[U], d[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a[u]), "=d"(b[u]), "=d"(c[u]), "=d"(d[u]) : "l"(src + 4 * (j + u * stride)));
#pragma unroll
        for (int u = 0; u < U; ++u)
            asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(dst + 4 * (j + u * stride)), "d"(a[u]), "d"(b[u]), "d"(c[u]), "d"(d[u]) : "memory");
    }
}
// TMA bulk stores: a CTA stages 32 KiB in shared memory (plain loads), one thread pushes it with cp.async.bulk
__global__ void __launch_bounds__(256) bulk_push_kernel(const double2* __restrict__ src, double2* __restrict__ dst, uint64_t n) {
    extern __shared__ __align__(128) unsigned char smem[];
    double2* buf[2] = {reinterpret_cast<double2*>(smem), reinterpret_cast<double2*>(smem) + 2048};
    int cur = 0;
    for (uint64_t base = (uint64_t)blockIdx.x * 2048; base + 2048 <= n; base += (uint64_t)gridDim.x * 2048, cur ^= 1) {
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // the buffer two tiles back is free
        __syncthreads();
        for (int j = threadIdx.x; j < 2048; j += blockDim.x) buf[cur][j] = __ldcs(&src[base + j]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned s = (unsigned)__cvta_generic_to_shared(buf[cur]);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst + base), "r"(s), "r"(32768) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
    const uint64_t n = 1ull << 28;      // 4 GiB per buffer
    double2 *a0, *b0, *a1, *b1;
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&a0, n * 16)); CK(cudaMalloc(&b0, n * 16));
    CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0)); CK(cudaMalloc(&a1, n * 16)); CK(cudaMalloc(&b1, n * 16));
    cudaStream_t s0, s1;
    CK(cudaSetDevice(0)); CK(cudaStreamCreate(&s0)); CK(cudaMemsetAsync(a0, 1, n * 16, s0)); CK(cudaMemsetAsync(b0, 2, n * 16, s0));
    CK(cudaSetDevice(1)); CK(cudaStreamCreate(&s1)); CK(cudaMemsetAsync(a1, 3, n * 16, s1)); CK(cudaMemsetAsync(b1, 4, n * 16, s1));
    cudaEvent_t e0, e1;
    CK(cudaSetDevice(0)); CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto sync_all = [&]() { cudaSetDevice(0); cudaDeviceSynchronize(); cudaSetDevice(1); cudaDeviceSynchronize(); cudaSetDevice(0); };
    const int grids[] = {148 * 4, 148 * 8, 148 * 16};
    // mode: 0 remote read (GPU0 pulls from GPU1), 1 remote write (GPU0 pushes to GPU1), 2 both directions of each at once,
    //       3 in-place swap, each GPU one half (the product's exchange), 4 cudaMemcpyPeer both directions
    cudaSetDevice(0); cudaFuncSetAttribute(bulk_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaSetDevice(1); cudaFuncSetAttribute(bulk_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaSetDevice(0);
    for (int mode = 0; mode < 10; ++mode)
        for (int gi = 0; gi < 3; ++gi) {
            const int grid = grids[gi];
            if (mode == 6 && gi) continue;
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                sync_all();
                CK(cudaEventRecord(e0, s0));
                if (mode == 0) copy_kernel<4><<<grid, 256, 0, s0>>>(a1, a0, n);
                if (mode == 1) copy_kernel<4><<<grid, 256, 0, s0>>>(a0, a1, n);
                if (mode == 2) { copy_kernel<4><<<grid, 256, 0, s0>>>(a1, a0, n); cudaSetDevice(1); copy_kernel<4><<<grid, 256, 0, s1>>>(b0, b1, n); cudaSetDevice(0); }
                if (mode == 3) { copy_kernel<4><<<grid, 256, 0, s0>>>(a0, a1, n); cudaSetDevice(1); copy_kernel<4><<<grid, 256, 0, s1>>>(b1, b0, n); cudaSetDevice(0); }
                if (mode == 4) { swap_kernel<4><<<grid, 256, 0, s0>>>(a0, a1, n / 2); cudaSetDevice(1); swap_kernel<4><<<grid, 256, 0, s1>>>(a1 + n / 2, a0 + n / 2, n / 2); cudaSetDevice(0); }
                if (mode == 5) { swap_kernel<8><<<grid, 256, 0, s0>>>(a0, a1, n / 2); cudaSetDevice(1); swap_kernel<8><<<grid, 256, 0, s1>>>(a1 + n / 2, a0 + n / 2, n / 2); cudaSetDevice(0); }
                if (mode == 6) { cudaMemcpyPeerAsync(a1, 1, a0, 0, n * 16, s0); cudaSetDevice(1); cudaMemcpyPeerAsync(b0, 0, b1, 1, n * 16, s1); cudaSetDevice(0); }
                if (mode == 7) { copy_kernel<8><<<grid, 256, 0, s0>>>(a0, a1, n); cudaSetDevice(1); copy_kernel<8><<<grid, 256, 0, s1>>>(b1, b0, n); cudaSetDevice(0); }
                if (mode == 8) { copy256_kernel<4><<<grid, 256, 0, s0>>>((const double*)a0, (double*)a1, n / 2); cudaSetDevice(1); copy256_kernel<4><<<grid, 256, 0, s1>>>((const double*)b1, (double*)b0, n / 2); cudaSetDevice(0); }
                if (mode == 9) { bulk_push_kernel<<<grid / 2, 256, 65536, s0>>>(a0, a1, n); cudaSetDevice(1); bulk_push_kernel<<<grid / 2, 256, 65536, s1>>>(b1, b0, n); cudaSetDevice(0); }
                cudaSetDevice(1); cudaStreamSynchronize(s1); cudaSetDevice(0);
                CK(cudaEventRecord(e1, s0));
                CK(cudaEventSynchronize(e1));
                float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const char* names[] = {"remote read  (one direction)", "remote write (one direction)", "remote reads, both directions", "remote writes, both directions",
                                   "in-place swap, halves (unroll 4)", "in-place swap, halves (unroll 8)", "cudaMemcpyPeer, both directions",
                                   "remote writes both ways, unroll 8", "remote writes both ways, 256-bit", "remote writes both ways, TMA bulk"};
            // bytes per direction: modes 0/1: n*16 one way; 2/3/6: n*16 each way; 4/5: each direction carries n*16 (half read by one side + half written by the other)
            printf("%-34s grid %5d: %8.2f ms  %7.1f GB/s per direction\n", names[mode], grid, best, n * 16.0 / (best * 1e-3) / 1e9);
        }
    return 0;
}
