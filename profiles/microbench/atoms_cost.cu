// Cost of shared-memory atomic increments (ATOMS.POPC.INC, what the choose kernels' histograms issue) as a function of
// the number of ACTIVE lanes per warp instruction and of the warps sharing an SM. One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_cost atoms_cost.cu && ./atoms_cost
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(int active, int iters, long long* out, uint32_t* sink) {
    __shared__ uint32_t hist[8 * 256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t* myh = hist + (wid & 7) * 256;
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
    const bool on = lane < active;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x = x * 1664525u + 1013904223u;
            if (on) atomicAdd(&myh[(x >> 13) & 255], 1u);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (threadIdx.x < 256) sink[blockIdx.x * 256 + threadIdx.x] = hist[threadIdx.x];
}

int main() {
    long long* out; uint32_t* sink;
    cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 148 * 256 * 4);
    const int iters = 256;
    printf("warps/SM active_lanes cycles_per_warp_instruction(SM-wide: cycles / (iters*8*warps)) cycles_per_active_lane\n");
    for (int warps : {1, 2, 8, 16, 24}) {
        for (int active : {1, 4, 8, 16, 32}) {
            k<<<148, warps * 32>>>(active, iters, out, sink);
            k<<<148, warps * 32>>>(active, iters, out, sink);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
            const double per_instr = avg / (iters * 8.0 * warps);
            printf("%2d %2d %8.2f %8.2f\n", warps, active, per_instr, per_instr / active);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
