// Shared helpers for the pcgnn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "pcgnn_b200.h"

#define PCG_FULL 0xffffffffu

// status / counter words (device int32 array, PCG_STATUS_WORDS long)
#define ST_SLOTS PCG_ST_SLOTS
#define ST_NSMALL 1      // items queued for the warp tier
#define ST_NMID 2        // ... the cta tier
#define ST_OVERFLOW PCG_ST_OVERFLOW
#define ST_NCL 4         // ... the cluster tier
#define ST_SMALL_CTR 5   // queue heads of the warp / cta tiers (device-side work distribution)
#define ST_MID_CTR 7
#define ST_NHUGE 8        // items queued for the huge tier (clusters of 8)
#define ST_NHUGE16 9      // ... for the clusters of 16
#define ST_NBIG 6        // ... the big tier

// Host-side caches (function attributes already set, side streams, probe results) are kept per device ordinal.
#define PCG_MAX_DEVICES 64
static inline int pcg_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= PCG_MAX_DEVICES) dev = 0;
    return dev;
}

void pcg_set_error(const char* fmt, ...);
int pcg_check_launch(const char* what);

// Programmatic dependent launch (PDL): when enabled (pcg_set_pdl), the kernels of the step's linear chain are
// launched with programmaticStreamSerializationAllowed, so a kernel's CTAs may become resident while the kernel in
// front of it on the stream drains; each of them executes griddepcontrol.wait before it touches that kernel's
// results (a no-op for launches without the attribute), and the kernels in front call
// griddepcontrol.launch_dependents as soon as they start.
int pcg_pdl_enabled();
__device__ __forceinline__ void pcg_grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pcg_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
static inline cudaError_t pcg_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

#define PCG_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            pcg_set_error(__VA_ARGS__);   \
            return (int)cudaErrorInvalidValue; \
        }                                 \
    } while (0)

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// |a - b| in fp32 as ordered bits: non-negative floats compare like unsigned ints, so the 32-bit
// pattern is the radix key (reference: torch.abs(center - neigh), src/layers.py:657).
__device__ __forceinline__ uint32_t dist_bits(float a, float b) {
    return __float_as_uint(fabsf(__fsub_rn(a, b)));
}

// Per-item sizes, in the reference's own double arithmetic:
//   c = math.ceil(d * threshold)   (src/layers.py:260)
//   keep c if d > c + 1 else all   (src/layers.py:662-672)
//   o = int(c * rho) for positive targets in train mode, at most P (src/layers.py:681, 690)
__device__ __forceinline__ void item_counts(int64_t d, double thr, double rho, bool positive, int P, int c_override,
                                            bool has_override, int& k, int& o) {
    int64_t c = has_override ? (int64_t)c_override : (int64_t)ceil((double)d * thr);
    k = (int)((d > c + 1) ? c : d);
    if (k < 0) k = 0;
    int64_t oo = positive ? (int64_t)((double)c * rho) : 0;
    if (oo > P) oo = P;
    if (oo < 0) oo = 0;
    o = (int)oo;
}

__device__ __forceinline__ float4 ld_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld_f4_cg(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void f4_add(float4& a, const float4& b) {
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}
