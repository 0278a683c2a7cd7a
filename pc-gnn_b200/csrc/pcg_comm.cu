// Gradient exchange + optimizer step of the data-parallel train step as ONE kernel over NVLink peer memory.
//
// The reference has no distributed code; its optimizer step is torch.optim.Adam on ~27 k
// parameters (/root/reference/src/model_handler.py:124, 153). Data parallel over the batch's targets needs one
// exchange per step: the sum of the small parameter gradient. With NCCL that is a separate, latency-bound
// collective between two CUDA graphs plus torch's multi-tensor Adam; here every rank
//   1. PUSHES its flat gradient into its slot of every peer's receive area (peer memory mapped by CUDA IPC: plain
//      16-byte stores over NVLink / NVSwitch, posted, no round trip). Every 8-byte half of a store is
//      {gradient word, step number}: the data carries its own arrival flag, so there is neither a system fence
//      (a round trip: all remote stores acknowledged) nor a separate flag store behind it,
//   2. polls the world's slots of its OWN receive area (local reads) until every word shows this step's number,
//      adds them in RANK ORDER (every rank computes bit-identical sums, so the replicas never drift) and divides by
//      the world size,
//   3. applies the Adam update (same formula as torch.optim.Adam with L2 weight decay) to its replica of the
//      parameters and clears the gradient for the next step,
// all inside the captured step graph. world == 1 runs step 3 only.
// History: round 1 PULLED (flag, then reads of every peer's send buffer, one NVLink round trip per peer): ~20 us at
// 8 GPUs against 5 us for the Adam part alone; early round 2 pushed the data, then fence.sys + a flag per CTA and peer
// (+15 us); flag-in-data removes the fence round trip and the flag hop.
// Areas are double buffered by step parity: a slot is rewritten two steps later, and a rank can only be two steps
// ahead of a peer that has not yet read if that peer had pushed twice in between, which it does after reading.
#include "pcg_common.cuh"

#define COMM_MAX_WORLD 8
#define COMM_NT 256
#define COMM_PER_CTA (COMM_NT * 4)
#define COMM_SPIN_LIMIT (1u << 26)        // polls before a rank gives up on a peer (seconds): NaN gradients, never a hang

struct CommP {
    float* grad;
    float* param;
    float* m;
    float* v;
    int n;
    int64_t n_pad;                        // floats per slot; a receive area is [2 parities][COMM_MAX_WORLD slots][2 * n_pad] words
    uint32_t* peer_recv[COMM_MAX_WORLD];  // every rank's receive area, as mapped HERE
    uint32_t* epoch;                      // steps completed so far (device counter, so graph replays advance it)
    int32_t* ticket;
    int rank, world;
    float lr, b1, b2, eps, wd;
    int do_adam;
};

__device__ __forceinline__ void st_sys_v4(uint32_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_sys_v4(const uint32_t* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(COMM_NT) k_allreduce_adam(CommP p) {
    __shared__ float s_c[2];              // step size, 1 / sqrt(bias correction 2)
    const int tid = threadIdx.x, c = blockIdx.x;
    pcg_grid_dependency_wait();           // the weight-gradient kernel in front of this one is complete
    const uint32_t e = *(volatile uint32_t*)p.epoch + 1u;      // this step's number (1-based)
    const int i0 = c * COMM_PER_CTA + tid * 4;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool in = i0 < p.n;             // n is padded to a multiple of 4 by the caller
    if (in) g = *reinterpret_cast<const float4*>(p.grad + i0);
    if (p.world > 1 && in) {
        // my slot of this step's parity, words {g.x, e, g.y, e} {g.z, e, g.w, e} at word offset 2 * i0
        const int64_t area = (int64_t)(e & 1u) * COMM_MAX_WORLD * 2 * p.n_pad;
        const int64_t mine = area + (int64_t)p.rank * 2 * p.n_pad + 2 * (int64_t)i0;
#pragma unroll
        for (int r = 0; r < COMM_MAX_WORLD; ++r)
            if (r < p.world && r != p.rank) {
                st_sys_v4(p.peer_recv[r] + mine, __float_as_uint(g.x), e, __float_as_uint(g.y), e);
                st_sys_v4(p.peer_recv[r] + mine + 4, __float_as_uint(g.z), e, __float_as_uint(g.w), e);
            }
        // every peer's words for my elements land in MY area: poll them (local reads), all slots in flight per round
        const uint32_t* base = p.peer_recv[p.rank] + area + 2 * (int64_t)i0;
        float4 x[COMM_MAX_WORLD];
        unsigned pending = 0;
#pragma unroll
        for (int r = 0; r < COMM_MAX_WORLD; ++r) {
            x[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < p.world && r != p.rank) pending |= 1u << r;
        }
        x[0] = p.rank == 0 ? g : x[0];
#pragma unroll
        for (int r = 1; r < COMM_MAX_WORLD; ++r)
            if (r == p.rank) x[r] = g;
        unsigned spins = 0;
        while (pending) {
            uint4 lo[COMM_MAX_WORLD], hi[COMM_MAX_WORLD];
#pragma unroll
            for (int r = 0; r < COMM_MAX_WORLD; ++r)
                if ((pending >> r) & 1u) {
                    lo[r] = ld_sys_v4(base + (int64_t)r * 2 * p.n_pad);
                    hi[r] = ld_sys_v4(base + (int64_t)r * 2 * p.n_pad + 4);
                }
#pragma unroll
            for (int r = 0; r < COMM_MAX_WORLD; ++r)
                if ((pending >> r) & 1u) {
                    if (lo[r].y == e && lo[r].w == e && hi[r].y == e && hi[r].w == e) {
                        x[r] = make_float4(__uint_as_float(lo[r].x), __uint_as_float(lo[r].z), __uint_as_float(hi[r].x),
                                           __uint_as_float(hi[r].z));
                        pending &= ~(1u << r);
                    }
                }
            if (++spins > COMM_SPIN_LIMIT) {          // a peer never arrived: poison the step instead of hanging
                x[0].x = __int_as_float(0x7fc00000);
                break;
            }
        }
        g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < COMM_MAX_WORLD; ++r)
            if (r < p.world) { g.x += x[r].x; g.y += x[r].y; g.z += x[r].z; g.w += x[r].w; }
        const float inv = 1.0f / (float)p.world;
        g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
    }
    if (p.do_adam) {
        if (tid == 0) {
            const double bc1 = 1.0 - pow((double)p.b1, (double)e), bc2 = 1.0 - pow((double)p.b2, (double)e);
            s_c[0] = (float)((double)p.lr / bc1);
            s_c[1] = (float)(1.0 / sqrt(bc2));
        }
        __syncthreads();
        if (in) {
            const float step = s_c[0], rs2 = s_c[1];
            float4 w = *reinterpret_cast<float4*>(p.param + i0);
            float4 m = *reinterpret_cast<float4*>(p.m + i0);
            float4 v = *reinterpret_cast<float4*>(p.v + i0);
#define PCG_ADAM1(W, M, V, G)                                           \
            {                                                           \
                const float gg = fmaf(p.wd, W, G);                      \
                M = fmaf(1.0f - p.b1, gg - M, M);                       \
                V = fmaf(1.0f - p.b2, gg * gg - V, V);                  \
                W -= step * M / (sqrtf(V) * rs2 + p.eps);               \
            }
            PCG_ADAM1(w.x, m.x, v.x, g.x) PCG_ADAM1(w.y, m.y, v.y, g.y)
            PCG_ADAM1(w.z, m.z, v.z, g.z) PCG_ADAM1(w.w, m.w, v.w, g.w)
#undef PCG_ADAM1
            *reinterpret_cast<float4*>(p.param + i0) = w;
            *reinterpret_cast<float4*>(p.m + i0) = m;
            *reinterpret_cast<float4*>(p.v + i0) = v;
            *reinterpret_cast<float4*>(p.grad + i0) = make_float4(0.f, 0.f, 0.f, 0.f);   // ready for the next backward
        }
    } else if (in) {
        *reinterpret_cast<float4*>(p.grad + i0) = g;
    }
    // last CTA out advances the step counter (every CTA has read it long before)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(p.ticket, 1) == (int)gridDim.x - 1) {
            *p.ticket = 0;
            *(volatile uint32_t*)p.epoch = e;
        }
    }
}

// ------------------------------------------------------------------------------------------- C ABI
extern "C" int pcg_comm_alloc(void** ptr, size_t bytes) {
    PCG_REQUIRE(ptr && bytes > 0, "pcg_comm_alloc: bad arguments");
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e == cudaSuccess) e = cudaMemset(*ptr, 0, bytes);
    if (e != cudaSuccess) { pcg_set_error("pcg_comm_alloc: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" int pcg_comm_free(void* ptr) {
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) { pcg_set_error("pcg_comm_free: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" int pcg_comm_export(void* ptr, unsigned char* handle64) {
    PCG_REQUIRE(ptr && handle64, "pcg_comm_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) { pcg_set_error("pcg_comm_export: %s", cudaGetErrorString(e)); return (int)e; }
    memcpy(handle64, &h, 64);
    return 0;
}

extern "C" int pcg_comm_import(const unsigned char* handle64, void** ptr) {
    PCG_REQUIRE(ptr && handle64, "pcg_comm_import: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { pcg_set_error("pcg_comm_import: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" int pcg_comm_unmap(void* ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) { pcg_set_error("pcg_comm_unmap: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" size_t pcg_comm_region_bytes(int64_t n_params) {
    const int64_t n_pad = (n_params + COMM_PER_CTA - 1) / COMM_PER_CTA * COMM_PER_CTA;
    const int64_t n_cta = n_pad / COMM_PER_CTA;
    (void)n_cta;
    return (size_t)(2 * COMM_MAX_WORLD * 2 * n_pad * 4 + 256);      // {word, step} pairs: two 32-bit words per gradient element
}

extern "C" int pcg_allreduce_adam(float* grad, float* param, float* m, float* v, int64_t n_params,
                                  void* const* peer_regions_host, int rank, int world, uint32_t* epoch, int32_t* ticket,
                                  float lr, float beta1, float beta2, float eps, float weight_decay, int do_adam,
                                  pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCG_REQUIRE(grad && epoch && ticket && n_params > 0 && n_params % 4 == 0, "pcg_allreduce_adam: bad arguments");
    PCG_REQUIRE(!do_adam || (param && m && v), "pcg_allreduce_adam: optimizer state missing");
    PCG_REQUIRE(world >= 1 && world <= COMM_MAX_WORLD && rank >= 0 && rank < world, "pcg_allreduce_adam: bad rank/world");
    PCG_REQUIRE(world == 1 || peer_regions_host, "pcg_allreduce_adam: peer regions missing");
    CommP p;
    p.grad = grad; p.param = param; p.m = m; p.v = v; p.n = (int)n_params;
    p.n_pad = (n_params + COMM_PER_CTA - 1) / COMM_PER_CTA * COMM_PER_CTA;
    for (int r = 0; r < COMM_MAX_WORLD; ++r) {
        p.peer_recv[r] = (world > 1 && r < world) ? (uint32_t*)peer_regions_host[r] : nullptr;
    }
    p.epoch = epoch; p.ticket = ticket; p.rank = rank; p.world = world;
    p.lr = lr; p.b1 = beta1; p.b2 = beta2; p.eps = eps; p.wd = weight_decay; p.do_adam = do_adam;
    cudaError_t le = pcg_launch(k_allreduce_adam, dim3((unsigned)(p.n_pad / COMM_PER_CTA)), dim3(COMM_NT), 0, stream,
                                (pcg_pdl_enabled() & 8) != 0, p);
    if (le != cudaSuccess) { pcg_set_error("pcg_allreduce_adam: launch: %s", cudaGetErrorString(le)); return (int)le; }
    return pcg_check_launch("pcg_allreduce_adam");
}

// ------------------------------------------------------------------------------------------------
// Score slice + halo exchange as one kernel (config C5: CSR rows, and with them the targets, are partitioned by
// node range; a target's neighbours may belong to any rank, so every rank needs every node's label score).
// Each rank scores only its own rows and writes the values straight into EVERY rank's score table (peer memory
// over NVLink, 128-byte coalesced stores), then bumps a counter in every peer; a one-warp wait kernel on each
// rank holds the stream until all peers' slices have landed. The reference computes these scores with
// label_clf over the batch's unique nodes (/root/reference/src/layers.py:231-237); NCCL's all-gather measured
// ~10 GB/s in this container (0.8 ms for 10 MB), which is what this replaces.
// Write-after-read: a rank starts step i+1 only after the gradient exchange of step i, which every rank joins
// after its own choose kernels, so no peer still reads the scores of step i (training steps only).
struct BcastP {
    float* table[COMM_MAX_WORLD];        // every rank's score table as mapped here
    uint32_t* counter[COMM_MAX_WORLD];   // every rank's arrival counters [COMM_MAX_WORLD]
    int rank, world;
};

__global__ void __launch_bounds__(256) k_score_bcast(const float* __restrict__ feat, int64_t n, int F, int64_t ldf,
                                                     const float* __restrict__ w, const float* __restrict__ b,
                                                     int64_t row_lo, BcastP p) {
    extern __shared__ float sw[];            // [ldf] weights, then 32 results
    float* out = sw + ldf;
    for (int c = threadIdx.x; c < ldf; c += blockDim.x) sw[c] = c < F ? w[c] : 0.f;
    const float bias = b ? b[0] : 0.f;
    __syncthreads();
    const int l = threadIdx.x & 7, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int V = (int)(ldf >> 2);
    for (int64_t base = (int64_t)blockIdx.x * 32; base < n; base += (int64_t)gridDim.x * 32) {
        const int64_t v = base + (threadIdx.x >> 3);
        float acc = 0.f;
        if (v < n) {
            const float* rowp = feat + v * ldf;
            for (int c = l; c < V; c += 8) {
                float4 x = ld_f4(rowp + 4 * c);
                const float4 ww = *reinterpret_cast<const float4*>(sw + 4 * c);
                acc = fmaf(x.x, ww.x, acc); acc = fmaf(x.y, ww.y, acc);
                acc = fmaf(x.z, ww.z, acc); acc = fmaf(x.w, ww.w, acc);
            }
        }
        acc += __shfl_xor_sync(PCG_FULL, acc, 4);
        acc += __shfl_xor_sync(PCG_FULL, acc, 2);
        acc += __shfl_xor_sync(PCG_FULL, acc, 1);
        if (l == 0) out[threadIdx.x >> 3] = acc + bias;
        __syncthreads();
        if (wid < p.world && base + lane < n) p.table[wid][row_lo + base + lane] = out[lane];   // warp r -> rank r
        __syncthreads();
    }
    if (threadIdx.x < p.world) {
        __threadfence_system();
        atomicAdd_system(p.counter[threadIdx.x] + p.rank, 1u);
    }
}

__global__ void k_score_wait(uint32_t* counters, int world, uint32_t n_cta, uint32_t* epoch) {
    const uint32_t e = *(volatile uint32_t*)epoch + 1u;
    if ((int)threadIdx.x < world)
        while ((int32_t)(ld_acquire_sys(counters + threadIdx.x) - e * n_cta) < 0) { }
    __syncwarp();
    if (threadIdx.x == 0) *(volatile uint32_t*)epoch = e;
}

extern "C" size_t pcg_score_region_bytes(int64_t n_global) {
    return (size_t)((n_global * 4 + 255) / 256 * 256 + 256);
}

extern "C" int pcg_score_bcast(const float* feat_rows, int64_t n_rows, int F, int64_t ldf, const float* w, const float* b,
                               int64_t row_lo, int64_t n_global, void* const* regions_host, int rank, int world,
                               uint32_t* epoch, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCG_REQUIRE(feat_rows && w && regions_host && epoch && n_rows > 0, "pcg_score_bcast: bad arguments");
    PCG_REQUIRE(world >= 1 && world <= COMM_MAX_WORLD && rank >= 0 && rank < world, "pcg_score_bcast: bad rank/world");
    PCG_REQUIRE(ldf % 4 == 0 && F <= ldf && ldf * 4 + 128 <= 48 * 1024, "pcg_score_bcast: bad row width");
    PCG_REQUIRE(((uintptr_t)feat_rows & 15) == 0, "pcg_score_bcast: feature rows must be 16-byte aligned");
    BcastP p;
    const size_t table_bytes = (size_t)((n_global * 4 + 255) / 256 * 256);
    for (int r = 0; r < COMM_MAX_WORLD; ++r) {
        char* base = r < world ? (char*)regions_host[r] : nullptr;
        p.table[r] = (float*)base;
        p.counter[r] = base ? (uint32_t*)(base + table_bytes) : nullptr;
    }
    p.rank = rank; p.world = world;
    int64_t blocks = (n_rows + 31) / 32;                  // must be the same on every rank (equal row ranges)
    const int64_t max_blocks = (int64_t)pcg_device_sms() * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    k_score_bcast<<<(unsigned)blocks, 256, (size_t)ldf * 4 + 128, stream>>>(feat_rows, n_rows, F, ldf, w, b, row_lo, p);
    k_score_wait<<<1, 32, 0, stream>>>(p.counter[rank], world, (uint32_t)blocks, epoch);
    return pcg_check_launch("pcg_score_bcast");
}
