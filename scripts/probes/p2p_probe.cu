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
    for (int mode = 0; mode < 7; ++mode)
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
                cudaSetDevice(1); cudaStreamSynchronize(s1); cudaSetDevice(0);
                CK(cudaEventRecord(e1, s0));
                CK(cudaEventSynchronize(e1));
                float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const char* names[] = {"remote read  (one direction)", "remote write (one direction)", "remote reads, both directions", "remote writes, both directions",
                                   "in-place swap, halves (unroll 4)", "in-place swap, halves (unroll 8)", "cudaMemcpyPeer, both directions"};
            // bytes per direction: modes 0/1: n*16 one way; 2/3/6: n*16 each way; 4/5: each direction carries n*16 (half read by one side + half written by the other)
            printf("%-34s grid %5d: %8.2f ms  %7.1f GB/s per direction\n", names[mode], grid, best, n * 16.0 / (best * 1e-3) / 1e9);
        }
    return 0;
}
