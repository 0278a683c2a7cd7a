// The two heads around the hot path, each one forward and one backward kernel:
//
//  * label_clf similarity head on the batch:  center[i][c] = <feat[targets[i]], Wc[c]> + bc[c]
//    (reference: layers.py:200 nn.Linear(F, 2), applied at :236 and sliced at :243; its gradient is the
//    only path by which the similarity loss of model.py:54 reaches label_clf)
//  * PCALayer head + loss:  logits = W[2,E] @ combined[E,B]; loss = CE(logits) + lambda * CE(center)
//    (reference: model.py:38, :54-61; both cross-entropies are batch means)
//
// Reductions over the batch are two-stage: per-block partials, then the last block to arrive (ticket
// counter) adds them in block order, so results are deterministic run to run.
#include "pcg_common.cuh"

#define HEAD_BLOCK 64

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(PCG_FULL, v, off);
    return v;
}

// last block to arrive returns true (and the counter is reset for the next launch)
__device__ __forceinline__ bool last_block(int32_t* ticket) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        s_last = (t == (int)gridDim.x - 1);
        if (s_last) *ticket = 0;
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last;
}

// ------------------------------------------------------------------------------------ center scores
__global__ void __launch_bounds__(HEAD_BLOCK) k_center_fwd(const float* __restrict__ feat, int64_t ldf, int F,
                                                           const int32_t* __restrict__ targets, int B,
                                                           const float* __restrict__ w, const float* __restrict__ b,
                                                           float* __restrict__ out) {
    extern __shared__ float sw[];            // [2][ldf] zero padded
    for (int c = threadIdx.x; c < 2 * ldf; c += blockDim.x) {
        const int cls = c / (int)ldf, f = c - cls * (int)ldf;
        sw[c] = f < F ? w[cls * F + f] : 0.f;
    }
    __syncthreads();
    const int l = threadIdx.x & 7;
    const int V = (int)(ldf >> 2);
    const int i = blockIdx.x * (HEAD_BLOCK / 8) + (threadIdx.x >> 3);       // 8 lanes per target
    float a0 = 0.f, a1 = 0.f;
    if (i < B) {
        const float* rowp = feat + (int64_t)targets[i] * ldf;
        for (int c = l; c < V; c += 8) {
            const float4 x = ld_f4(rowp + 4 * c);
            const float4 w0 = *reinterpret_cast<const float4*>(sw + 4 * c);
            const float4 w1 = *reinterpret_cast<const float4*>(sw + ldf + 4 * c);
            a0 = fmaf(x.x, w0.x, a0); a0 = fmaf(x.y, w0.y, a0); a0 = fmaf(x.z, w0.z, a0); a0 = fmaf(x.w, w0.w, a0);
            a1 = fmaf(x.x, w1.x, a1); a1 = fmaf(x.y, w1.y, a1); a1 = fmaf(x.z, w1.z, a1); a1 = fmaf(x.w, w1.w, a1);
        }
    }
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
        a0 += __shfl_xor_sync(PCG_FULL, a0, off);
        a1 += __shfl_xor_sync(PCG_FULL, a1, off);
    }
    if (l == 0 && i < B) {
        out[2 * i] = a0 + b[0];
        out[2 * i + 1] = a1 + b[1];
    }
}

// dW[c][f] = sum_i g[i][c] * feat[targets[i]][f];  db[c] = sum_i g[i][c].
// Block `blk` owns CENTER_PER targets: thread (j, f) = (tid / 32, tid % 32 + 32*c) multiplies target j's row
// (one coalesced row read per warp, all warps' loads in flight together), the CENTER_PER partial rows are then
// added in target order by the first warps; the last block to arrive adds the blocks in block order.
#define CENTER_PER 32
#define CENTER_NT 1024
__global__ void __launch_bounds__(CENTER_NT) k_center_bwd(const float* __restrict__ feat, int64_t ldf, int F,
                                                          const int32_t* __restrict__ targets, int B,
                                                          const float* __restrict__ g, float* __restrict__ partial,
                                                          int32_t* ticket, float* __restrict__ dw,
                                                          float* __restrict__ db) {
    extern __shared__ float sp[];               // [CENTER_PER][2][ldf] products, then reused
    const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * CENTER_PER + j;
    const int stride = 2 * F + 2;
    float g0 = 0.f, g1 = 0.f;
    const float* rowp = feat;
    if (i < B) {
        g0 = g[2 * i];
        g1 = g[2 * i + 1];
        rowp = feat + (int64_t)targets[i] * ldf;
    }
    for (int f = lane; f < ldf; f += 32) {
        const float x = (i < B && f < F) ? __ldg(rowp + f) : 0.f;
        sp[(j * 2 + 0) * ldf + f] = g0 * x;
        sp[(j * 2 + 1) * ldf + f] = g1 * x;
    }
    __syncthreads();
    float* dst = partial + (int64_t)blockIdx.x * stride;
    for (int x = threadIdx.x; x < 2 * F; x += CENTER_NT) {
        const int c = x / F, f = x - c * F;
        float a = 0.f;
#pragma unroll 8
        for (int q = 0; q < CENTER_PER; ++q) a += sp[(q * 2 + c) * ldf + f];
        dst[x] = a;
    }
    if (threadIdx.x < 2) {                       // bias gradient of this block (target order)
        float a = 0.f;
        for (int q = 0; q < CENTER_PER; ++q) {
            const int t = blockIdx.x * CENTER_PER + q;
            if (t < B) a += g[2 * t + threadIdx.x];
        }
        dst[2 * F + threadIdx.x] = a;
    }
    if (!last_block(ticket)) return;
    for (int x = threadIdx.x; x < stride; x += CENTER_NT) {
        float s = 0.f;
        for (int q = 0; q < (int)gridDim.x; ++q) s += __ldcg(partial + (int64_t)q * stride + x);
        if (x < 2 * F) dw[x] = s; else db[x - 2 * F] = s;
    }
}

// ------------------------------------------------------------------------------------ head + loss
// HEAD_TB targets per block of 256 threads: thread (g, t) = (tid / 32, tid % 32) covers the embedding rows
// e = g, g + 8, ... of target t (coalesced along the batch, every load independent), the 8 partial sums are
// added in g order. p1[i] / q1[i] = softmax probability of class 1 (GNN head / label head), kept for the
// backward pass.
#define HEAD_TB 32
#define HEAD_NT 256
__global__ void __launch_bounds__(HEAD_NT) k_head_loss_fwd(const float* __restrict__ emb, int E, int B,
                                                           const float* __restrict__ w, const float* __restrict__ center,
                                                           const int64_t* __restrict__ labels, float lambda,
                                                           float* __restrict__ logits, float* __restrict__ p1,
                                                           float* __restrict__ q1, float* __restrict__ partial,
                                                           int32_t* ticket, float* __restrict__ loss) {
    extern __shared__ float sw[];            // [2][E]
    __shared__ float ps[HEAD_NT / 32][HEAD_TB][2];
    for (int c = threadIdx.x; c < 2 * E; c += blockDim.x) sw[c] = w[c];
    __syncthreads();
    const int g = threadIdx.x >> 5, t = threadIdx.x & 31;
    const int i = blockIdx.x * HEAD_TB + t;
    float g0 = 0.f, g1 = 0.f;
    if (i < B) {
#pragma unroll 8
        for (int e = g; e < E; e += HEAD_NT / 32) {
            const float x = __ldg(emb + (int64_t)e * B + i);
            g0 = fmaf(sw[e], x, g0);
            g1 = fmaf(sw[E + e], x, g1);
        }
    }
    ps[g][t][0] = g0;
    ps[g][t][1] = g1;
    __syncthreads();
    if (g == 0) {
        float lg = 0.f, ll = 0.f;
        if (i < B) {
            g0 = g1 = 0.f;
#pragma unroll
            for (int q = 0; q < HEAD_NT / 32; ++q) { g0 += ps[q][t][0]; g1 += ps[q][t][1]; }
            logits[2 * i] = g0;
            logits[2 * i + 1] = g1;
            const int y = labels[i] == 1 ? 1 : 0;
            {   // cross entropy = logsumexp - logit[y], computed like log_softmax (shift by the max)
                const float m = fmaxf(g0, g1);
                const float e0 = expf(g0 - m), e1 = expf(g1 - m);
                const float lse = m + logf(e0 + e1);
                lg = lse - (y ? g1 : g0);
                p1[i] = e1 / (e0 + e1);
            }
            {
                const float c0 = center[2 * i], c1 = center[2 * i + 1];
                const float m = fmaxf(c0, c1);
                const float e0 = expf(c0 - m), e1 = expf(c1 - m);
                const float lse = m + logf(e0 + e1);
                ll = lse - (y ? c1 : c0);
                q1[i] = e1 / (e0 + e1);
            }
        }
        lg = warp_sum(lg);
        ll = warp_sum(ll);
        if (t == 0) {
            partial[2 * blockIdx.x] = lg;
            partial[2 * blockIdx.x + 1] = ll;
        }
    }
    if (!last_block(ticket)) return;
    if (threadIdx.x == 0) {
        float a = 0.f, c = 0.f;
        for (int q = 0; q < (int)gridDim.x; ++q) { a += __ldcg(partial + 2 * q); c += __ldcg(partial + 2 * q + 1); }
        loss[0] = a / (float)B + lambda * (c / (float)B);       // model.py:54-61 (both CE are batch means)
    }
}

// d_emb[e][i] = sum_c W[c][e] * dl[i][c];  d_center[i][c] = lambda * (q - onehot) * s;  dW[c][e] = sum_i dl[i][c] * emb[e][i]
// with dl[i][c] = (p - onehot) * s and s = d_loss / B. Same (g, t) layout as the forward kernel: warp g owns the
// embedding rows e = g, g + 8, ...; lanes run along the block's HEAD_TB targets.
__global__ void __launch_bounds__(HEAD_NT) k_head_loss_bwd(const float* __restrict__ emb, int E, int B,
                                                           const float* __restrict__ w, const int64_t* __restrict__ labels,
                                                           const float* __restrict__ p1, const float* __restrict__ q1,
                                                           float lambda, const float* __restrict__ d_loss,
                                                           float* __restrict__ d_emb, float* __restrict__ d_center,
                                                           float* __restrict__ partial, int32_t* ticket,
                                                           float* __restrict__ dw) {
    extern __shared__ float sw[];            // [2][E]
    for (int c = threadIdx.x; c < 2 * E; c += blockDim.x) sw[c] = w[c];
    const int g = threadIdx.x >> 5, t = threadIdx.x & 31;
    const int i = blockIdx.x * HEAD_TB + t;
    const float s = d_loss[0] / (float)B;
    float dl0 = 0.f, dl1 = 0.f;
    if (i < B) {
        const int y = labels[i] == 1 ? 1 : 0;
        const float p = p1[i], q = q1[i];
        dl1 = (p - (float)y) * s;
        dl0 = -dl1;                                   // (1-p) - (1-y) = -(p - y)
        if (g == 0) {
            const float dc1 = lambda * (q - (float)y) * s;
            d_center[2 * i] = -dc1;
            d_center[2 * i + 1] = dc1;
        }
    }
    __syncthreads();
    float* dst = partial + (int64_t)blockIdx.x * 2 * E;
#pragma unroll 4
    for (int e = g; e < E; e += HEAD_NT / 32) {
        const float x = i < B ? __ldg(emb + (int64_t)e * B + i) : 0.f;
        if (i < B) d_emb[(int64_t)e * B + i] = fmaf(sw[e], dl0, sw[E + e] * dl1);
        const float a = warp_sum(dl0 * x), b = warp_sum(dl1 * x);
        if (t == 0) { dst[e] = a; dst[E + e] = b; }
    }
    if (!last_block(ticket)) return;
    for (int x = threadIdx.x; x < 2 * E; x += blockDim.x) {
        float a = 0.f;
        for (int q = 0; q < (int)gridDim.x; ++q) a += __ldcg(partial + (int64_t)q * 2 * E + x);
        dw[x] = a;
    }
}

// ------------------------------------------------------------------------------------------- C ABI
extern "C" size_t pcg_head_scratch_floats(int B, int F, int E) {
    const size_t blocks = (size_t)(B + HEAD_TB - 1) / HEAD_TB + 1;
    size_t a = ((size_t)(B + 31) / 32 + 1) * (size_t)(2 * F + 2);   // center bwd: one partial per 32 targets
    size_t b = blocks * 2 * (size_t)(E > 1 ? E : 1);
    return (a > b ? a : b) + 64;
}

extern "C" int pcg_center_fwd(const float* feat, int64_t ldf, int F, const int32_t* targets, int B, const float* w,
                              const float* b, float* center, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B == 0) return 0;
    PCG_REQUIRE(feat && targets && w && b && center, "pcg_center_fwd: null pointer");
    PCG_REQUIRE(ldf % 4 == 0 && F <= ldf && 2 * ldf * 4 <= 48 * 1024, "pcg_center_fwd: bad row width");
    k_center_fwd<<<(B + HEAD_BLOCK / 8 - 1) / (HEAD_BLOCK / 8), HEAD_BLOCK, (size_t)2 * ldf * 4, stream>>>(
        feat, ldf, F, targets, B, w, b, center);
    return pcg_check_launch("pcg_center_fwd");
}

extern "C" int pcg_center_bwd(const float* feat, int64_t ldf, int F, const int32_t* targets, int B,
                              const float* d_center, float* d_w, float* d_b, float* scratch, int32_t* ticket,
                              pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCG_REQUIRE(feat && targets && d_center && d_w && d_b && scratch && ticket, "pcg_center_bwd: null pointer");
    int blocks = (B + CENTER_PER - 1) / CENTER_PER;       // every block owns <= CENTER_PER targets
    if (blocks < 1) blocks = 1;
    const size_t smem = (size_t)CENTER_PER * 2 * ldf * 4;
    PCG_REQUIRE(smem <= 200 * 1024, "pcg_center_bwd: feature rows too wide (ldf=%lld)", (long long)ldf);
    static size_t configured_dev[PCG_MAX_DEVICES];
    size_t& configured = configured_dev[pcg_current_device()];
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_center_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { pcg_set_error("pcg_center_bwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        configured = smem;
    }
    k_center_bwd<<<blocks, CENTER_NT, smem, stream>>>(feat, ldf, F, targets, B, d_center, scratch, ticket, d_w, d_b);
    return pcg_check_launch("pcg_center_bwd");
}

extern "C" int pcg_head_loss_fwd(const float* emb, int E, int B, const float* w, const float* center,
                                 const int64_t* labels, float lambda, float* logits, float* p1, float* q1,
                                 float* loss, float* scratch, int32_t* ticket, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCG_REQUIRE(B > 0, "pcg_head_loss_fwd: empty batch");
    PCG_REQUIRE(emb && w && center && labels && logits && p1 && q1 && loss && scratch && ticket,
                "pcg_head_loss_fwd: null pointer");
    PCG_REQUIRE((size_t)2 * E * 4 <= 40 * 1024, "pcg_head_loss_fwd: embed dim too large");
    const int blocks = (B + HEAD_TB - 1) / HEAD_TB;
    k_head_loss_fwd<<<blocks, HEAD_NT, (size_t)2 * E * 4, stream>>>(emb, E, B, w, center, labels, lambda, logits, p1,
                                                                      q1, scratch, ticket, loss);
    return pcg_check_launch("pcg_head_loss_fwd");
}

extern "C" int pcg_head_loss_bwd(const float* emb, int E, int B, const float* w, const int64_t* labels, const float* p1,
                                 const float* q1, float lambda, const float* d_loss, float* d_emb, float* d_center,
                                 float* d_w, float* scratch, int32_t* ticket, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCG_REQUIRE(B > 0, "pcg_head_loss_bwd: empty batch");
    PCG_REQUIRE(emb && w && labels && p1 && q1 && d_loss && d_emb && d_center && d_w && scratch && ticket,
                "pcg_head_loss_bwd: null pointer");
    const size_t smem = (size_t)2 * E * 4;
    PCG_REQUIRE(smem <= 48 * 1024, "pcg_head_loss_bwd: embed dim too large");
    const int blocks = (B + HEAD_TB - 1) / HEAD_TB;
    k_head_loss_bwd<<<blocks, HEAD_NT, smem, stream>>>(emb, E, B, w, labels, p1, q1, lambda, d_loss, d_emb, d_center,
                                                          scratch, ticket, d_w);
    return pcg_check_launch("pcg_head_loss_bwd");
}
