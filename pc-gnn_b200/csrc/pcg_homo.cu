// Encoder of the GraphSAGE / GCN baselines: out[e][i] = relu(sum_f W[e][f] * X[i][f]) with X = the aggregated rows
// (GCN, and GraphSAGE with gcn=True) or [self | aggregate] (GraphSAGE, gcn=False), and its weight gradient.
//
// Reference (/root/reference/src/graphsage.py): `F.relu(self.weight.mm(combined.t()))` at :149 (Encoder) and :274
// (GCNEncoder) with weight [E, F] / [E, 2F]; the cat at :145; autograd's mm / relu backward. Round 1 left these to
// torch / cuBLAS; the shapes are tiny (C4: 1024 x 25 x 64 = 3 MFLOP), so each direction is one small kernel with the
// weight (forward) or the batch slice (backward) staged in shared memory. The head and the cross-entropy behind the
// encoder reuse pcg_head_loss_fwd / _bwd (lambda = 0).
#include "pcg_common.cuh"

#define ENC_TI 32            // targets per CTA
#define ENC_NT 256

struct EncP {
    const float* agg; int64_t lda;                 // [B, lda] aggregated rows
    const float* feat; int64_t ldf; const int32_t* targets;   // self rows (NULL: no self half)
    const float* w;                                // [E, Fin]
    int B, F, Fin, E;
    float* out;                                    // [E, B]
    const float* d_out;                            // [E, B]  (backward)
    float* partial; int32_t* ticket; float* d_w;   // (backward)
};

// X tile [ENC_TI][Fin + 1] into shared memory (zero rows beyond the batch)
__device__ __forceinline__ void load_x_tile(const EncP& p, int i0, float* xs, int ldx) {
    for (int idx = threadIdx.x; idx < ENC_TI * p.Fin; idx += ENC_NT) {
        const int i = idx / p.Fin, f = idx - i * p.Fin;
        float v = 0.f;
        if (i0 + i < p.B) {
            if (p.feat && f < p.F) v = __ldg(p.feat + (int64_t)__ldg(p.targets + i0 + i) * p.ldf + f);
            else v = __ldg(p.agg + (int64_t)(i0 + i) * p.lda + (p.feat ? f - p.F : f));
        }
        xs[i * ldx + f] = v;
    }
}

__global__ void __launch_bounds__(ENC_NT) k_encoder_fwd(EncP p) {
    extern __shared__ float sm[];
    const int ldx = p.Fin + 1;
    float* ws = sm;                                // [E][Fin]
    float* xs = sm + (size_t)p.E * p.Fin;          // [TI][Fin + 1]
    const int i0 = blockIdx.x * ENC_TI;
    for (int idx = threadIdx.x; idx < p.E * p.Fin; idx += ENC_NT) ws[idx] = __ldg(p.w + idx);
    load_x_tile(p, i0, xs, ldx);
    __syncthreads();
    const int i = threadIdx.x & (ENC_TI - 1);
    for (int e = threadIdx.x / ENC_TI; e < p.E; e += ENC_NT / ENC_TI) {     // a warp: one e, 32 targets
        const float* wr = ws + (size_t)e * p.Fin;
        const float* xr = xs + i * ldx;
        float a = 0.f;
        for (int f = 0; f < p.Fin; ++f) a = fmaf(wr[f], xr[f], a);
        if (i0 + i < p.B) p.out[(int64_t)e * p.B + i0 + i] = fmaxf(a, 0.f);
    }
}

// d_w[e][f] = sum_i (d_out[e][i] * (out[e][i] > 0)) * X[i][f]: every CTA reduces its ENC_TI targets, the last CTA to
// finish (ticket) adds the CTAs' partials in CTA order (deterministic).
__global__ void __launch_bounds__(ENC_NT) k_encoder_bwd(EncP p) {
    extern __shared__ float sm[];
    __shared__ int s_last;
    const int ldx = p.Fin + 1, ldg = ENC_TI + 1;
    float* gs = sm;                                // [E][TI + 1]
    float* xs = sm + (size_t)p.E * ldg;            // [TI][Fin + 1]
    const int i0 = blockIdx.x * ENC_TI;
    for (int idx = threadIdx.x; idx < p.E * ENC_TI; idx += ENC_NT) {
        const int e = idx / ENC_TI, i = idx - e * ENC_TI;
        float g = 0.f;
        if (i0 + i < p.B) {
            const int64_t at = (int64_t)e * p.B + i0 + i;
            g = __ldg(p.out + at) > 0.f ? __ldg(p.d_out + at) : 0.f;
        }
        gs[e * ldg + i] = g;
    }
    load_x_tile(p, i0, xs, ldx);
    __syncthreads();
    const int n = p.E * p.Fin;
    float* mine = p.partial + (size_t)blockIdx.x * n;
    for (int idx = threadIdx.x; idx < n; idx += ENC_NT) {
        const int e = idx / p.Fin, f = idx - e * p.Fin;
        float a = 0.f;
#pragma unroll 8
        for (int i = 0; i < ENC_TI; ++i) a = fmaf(gs[e * ldg + i], xs[i * ldx + f], a);
        mine[idx] = a;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(p.ticket, 1);
        s_last = (t == (int)gridDim.x - 1);
        if (s_last) *p.ticket = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int idx = threadIdx.x; idx < n; idx += ENC_NT) {
        float a = 0.f;
        for (int q0 = 0; q0 < (int)gridDim.x; q0 += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = q0 + u < (int)gridDim.x ? __ldcg(p.partial + (size_t)(q0 + u) * n + idx) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) a += v[u];
        }
        p.d_w[idx] = a;
    }
}

// ------------------------------------------------------------------------------------------- C ABI
extern "C" size_t pcg_encoder_scratch_floats(int B, int F_in, int E) {
    return (size_t)((B + ENC_TI - 1) / ENC_TI) * (size_t)E * F_in + 64;
}

static int enc_fill(EncP& p, const float* agg, int64_t lda, const float* feat, int64_t ldf, const int32_t* targets, int F,
                    const float* w, int B, int E, const char* who) {
    PCG_REQUIRE(agg && w && B > 0 && E > 0 && F > 0, "%s: bad arguments", who);
    PCG_REQUIRE((feat == nullptr) == (targets == nullptr), "%s: self rows need both the table and the target ids", who);
    p.agg = agg; p.lda = lda; p.feat = feat; p.ldf = ldf; p.targets = targets; p.w = w;
    p.B = B; p.F = F; p.Fin = feat ? 2 * F : F; p.E = E;
    return 0;
}

static int enc_smem(cudaError_t (*set)(size_t), size_t bytes, size_t& configured, const char* who) {
    PCG_REQUIRE(bytes <= 200 * 1024, "%s: weight / tile too large for shared memory (%zu bytes)", who, bytes);
    if (bytes > 48 * 1024 && bytes > configured) {
        cudaError_t e = set(bytes);
        if (e != cudaSuccess) { pcg_set_error("%s: smem attr: %s", who, cudaGetErrorString(e)); return (int)e; }
        configured = bytes;
    }
    return 0;
}

extern "C" int pcg_encoder_fwd(const float* agg, int64_t lda, const float* feat, int64_t ldf, const int32_t* targets, int F,
                               const float* w, int B, int E, float* out, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B == 0) return 0;
    EncP p = {};
    int rc = enc_fill(p, agg, lda, feat, ldf, targets, F, w, B, E, "pcg_encoder_fwd");
    if (rc) return rc;
    PCG_REQUIRE(out, "pcg_encoder_fwd: null output");
    p.out = out;
    const size_t smem = ((size_t)E * p.Fin + (size_t)ENC_TI * (p.Fin + 1)) * 4;
    static size_t configured_dev[PCG_MAX_DEVICES];
    size_t& configured = configured_dev[pcg_current_device()];
    rc = enc_smem([](size_t b) { return cudaFuncSetAttribute(k_encoder_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b); },
                  smem, configured, "pcg_encoder_fwd");
    if (rc) return rc;
    k_encoder_fwd<<<(B + ENC_TI - 1) / ENC_TI, ENC_NT, smem, stream>>>(p);
    return pcg_check_launch("pcg_encoder_fwd");
}

extern "C" int pcg_encoder_bwd(const float* agg, int64_t lda, const float* feat, int64_t ldf, const int32_t* targets, int F,
                               int B, int E, const float* out, const float* d_out, float* d_w, float* scratch,
                               int32_t* ticket, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    EncP p = {};
    int rc = enc_fill(p, agg, lda, feat, ldf, targets, F, agg /* unused */, B, E, "pcg_encoder_bwd");
    if (rc) return rc;
    PCG_REQUIRE(out && d_out && d_w && scratch && ticket, "pcg_encoder_bwd: null pointer");
    p.out = const_cast<float*>(out); p.d_out = d_out; p.d_w = d_w; p.partial = scratch; p.ticket = ticket;
    const size_t smem = ((size_t)E * (ENC_TI + 1) + (size_t)ENC_TI * (p.Fin + 1)) * 4;
    static size_t configured_dev[PCG_MAX_DEVICES];
    size_t& configured = configured_dev[pcg_current_device()];
    rc = enc_smem([](size_t b) { return cudaFuncSetAttribute(k_encoder_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b); },
                  smem, configured, "pcg_encoder_bwd");
    if (rc) return rc;
    k_encoder_bwd<<<(B + ENC_TI - 1) / ENC_TI, ENC_NT, smem, stream>>>(p);
    return pcg_check_launch("pcg_encoder_bwd");
}
