// Fused relation transforms + inter-relation combine (forward and backward), fp32 FFMA.
//
// Reference (/root/reference/src/layers.py): per relation r  to_feats_r = relu(cat(self, agg_r) @ W_r)
// (:616-629), then combined = relu(cat(self, to_feats_1..R) @ W).t() (:273-289), each a separate
// cat / mm / relu library call (and their autograd twins in backward). Here one kernel per direction
// keeps a tile of targets in shared memory across all R+1 small GEMMs, and the weight gradients are
// a split-K GEMM over the batch with a fixed-order reduction (deterministic).
//
// fp32 on the CUDA cores on purpose: the parity bar is 1e-5 relative, which tf32/bf16 tensor-core
// paths do not meet, and the whole dense part is ~50 MFLOP per step (microseconds).
#include "pcg_common.cuh"

#define DENSE_KC 32          // weight rows staged per chunk
#define DENSE_MAX_E 256      // embed dim limit (4 column chunks of 64)

struct DenseP {
    const float* feat;
    int64_t ldf;
    int F, B, R, E;
    const int32_t* targets;
    const float* agg;                 // [R*B, ldf]
    const float* w_intra[PCG_MAX_REL];  // each [2F, E]
    const float* w_inter;             // [F + R*E, E]
    float* cat;                       // [B, F + R*E]
    float* out;                       // [E, B]
};

// acc[c][q] += x * Ws[kk][c*64 + tx*4 + q] over the staged chunk
template <int CE>
__device__ __forceinline__ void fma_chunk(float (&acc)[CE][4], const float* __restrict__ xrow, int k0, int kend,
                                          const float* __restrict__ ws, int Ep, int tx) {
    for (int k = k0; k < kend; ++k) {
        const float x = xrow[k];
        const float* wr = ws + (k - k0) * Ep + tx * 4;
#pragma unroll
        for (int c = 0; c < CE; ++c) {
            const float4 w4 = *reinterpret_cast<const float4*>(wr + c * 64);
            acc[c][0] = fmaf(x, w4.x, acc[c][0]);
            acc[c][1] = fmaf(x, w4.y, acc[c][1]);
            acc[c][2] = fmaf(x, w4.z, acc[c][2]);
            acc[c][3] = fmaf(x, w4.w, acc[c][3]);
        }
    }
}

// stage rows [k0, k0+KC) of a [K, E] row-major weight into ws[KC][Ep] (zero padded)
__device__ __forceinline__ void stage_weight(const float* __restrict__ w, int K, int E, int k0, float* ws, int Ep,
                                             int tid, int nth) {
    for (int t = tid; t < DENSE_KC * Ep; t += nth) {
        const int kk = t / Ep, col = t - kk * Ep;
        const int k = k0 + kk;
        ws[t] = (k < K && col < E) ? __ldg(w + (int64_t)k * E + col) : 0.f;
    }
}

// TM targets per CTA, 16 threads per target row; thread (ty, tx) owns columns tx*4+q+64*c.
template <int TM, int CE>
__global__ void __launch_bounds__(TM * 16) k_dense_fwd(DenseP p) {
    extern __shared__ float sm[];
    const int F = p.F, E = p.E, R = p.R, B = p.B;
    const int K2 = F + R * E;
    const int K2p = K2 + 1, Fp = F + 1, Ep = CE * 64;
    float* cats = sm;                       // [TM][K2p]
    float* aggs = cats + TM * K2p;          // [TM][Fp]
    float* ws = sm + ((TM * K2p + TM * Fp + 3) & ~3);   // [KC][Ep], 16-byte aligned
    const int tid = threadIdx.x, nth = TM * 16;
    const int ty = tid >> 4, tx = tid & 15;
    const int i = blockIdx.x * TM + ty;
    const bool valid = i < B;
    const int32_t v = valid ? p.targets[i] : 0;
    float* crow = cats + ty * K2p;
    float* arow = aggs + ty * Fp;
    for (int f = tx; f < F; f += 16) crow[f] = valid ? __ldg(p.feat + (int64_t)v * p.ldf + f) : 0.f;
    for (int r = 0; r < R; ++r) {
        for (int f = tx; f < F; f += 16) arow[f] = valid ? __ldg(p.agg + ((int64_t)r * B + i) * p.ldf + f) : 0.f;
        float acc[CE][4];
#pragma unroll
        for (int c = 0; c < CE; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
        const float* w = p.w_intra[r];
        for (int k0 = 0; k0 < 2 * F; k0 += DENSE_KC) {
            __syncthreads();
            stage_weight(w, 2 * F, E, k0, ws, Ep, tid, nth);
            __syncthreads();
            const int kend = min(k0 + DENSE_KC, 2 * F);
            // rows k < F come from the self features, the rest from the aggregated neighbours
            const int ksplit = min(max(F, k0), kend);
            fma_chunk<CE>(acc, crow, k0, ksplit, ws, Ep, tx);
            if (ksplit < kend) fma_chunk<CE>(acc, arow - F, ksplit, kend, ws + (ksplit - k0) * Ep, Ep, tx);
        }
#pragma unroll
        for (int c = 0; c < CE; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int col = c * 64 + tx * 4 + q;
                if (col < E) crow[F + r * E + col] = fmaxf(acc[c][q], 0.f);
            }
    }
    float acc[CE][4];
#pragma unroll
    for (int c = 0; c < CE; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
    for (int k0 = 0; k0 < K2; k0 += DENSE_KC) {
        __syncthreads();
        stage_weight(p.w_inter, K2, E, k0, ws, Ep, tid, nth);
        __syncthreads();
        fma_chunk<CE>(acc, crow, k0, min(k0 + DENSE_KC, K2), ws, Ep, tx);
    }
    if (valid) {
#pragma unroll
        for (int c = 0; c < CE; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int col = c * 64 + tx * 4 + q;
                if (col < E) p.out[(int64_t)col * B + i] = fmaxf(acc[c][q], 0.f);
            }
        for (int k = tx; k < K2; k += 16) p.cat[(int64_t)i * K2 + k] = crow[k];
    }
}

// ---- backward, row part: dZ = d_out * (out > 0) [B,E];  dH = (dZ @ W[F:,:]^T) * (cat[:, F:] > 0) [B, R*E]
struct DenseBwdP {
    int F, B, R, E;
    const float* cat;
    const float* out;
    const float* d_out;
    const float* w_inter;
    float* dz;      // [B, E]
    float* dh;      // [B, R*E]
};

template <int TM>
__global__ void __launch_bounds__(TM * 16) k_dense_bwd_rows(DenseBwdP p) {
    extern __shared__ float sm[];
    const int F = p.F, E = p.E, R = p.R, B = p.B;
    const int K2 = F + R * E, RE = R * E, Ep1 = E + 1;
    float* dzs = sm;                 // [TM][Ep1]
    float* ws2 = dzs + TM * Ep1;     // [64][Ep1]   rows c of W[F + c, :]
    const int tid = threadIdx.x, nth = TM * 16;
    const int ty = tid >> 4, tx = tid & 15;
    const int i = blockIdx.x * TM + ty;
    const bool valid = i < B;
    for (int e = tx; e < E; e += 16) {
        float g = 0.f;
        if (valid) {
            const float o = p.out[(int64_t)e * B + i];
            g = o > 0.f ? p.d_out[(int64_t)e * B + i] : 0.f;
            p.dz[(int64_t)i * E + e] = g;
        }
        dzs[ty * Ep1 + e] = g;
    }
    for (int c0 = 0; c0 < RE; c0 += 64) {
        __syncthreads();
        for (int t = tid; t < 64 * E; t += nth) {
            const int cc = t / E, e = t - cc * E;
            const int c = c0 + cc;
            ws2[cc * Ep1 + e] = c < RE ? __ldg(p.w_inter + (int64_t)(F + c) * E + e) : 0.f;
        }
        __syncthreads();
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const float* zr = dzs + ty * Ep1;
        for (int e = 0; e < E; ++e) {
            const float z = zr[e];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fmaf(z, ws2[(tx + 16 * q) * Ep1 + e], acc[q]);
        }
        if (valid) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = c0 + tx + 16 * q;
                if (c < RE) {
                    const float h = p.cat[(int64_t)i * K2 + F + c];
                    p.dh[(int64_t)i * RE + c] = h > 0.f ? acc[q] : 0.f;
                }
            }
        }
    }
}

// ---- backward, weight part: part[job][split][m][n] = sum over the split's targets of A[i][m] * G[i][n]
//   job 0:      A = cat [B, K2],                        G = dZ [B, E]           -> d W_inter  [K2, E]
//   job 1 + r:  A = [cat[:, :F] | agg_r[:, :F]] [B,2F],  G = dH[:, rE:(r+1)E]    -> d W_r      [2F, E]
struct WgradP {
    int F, B, R, E, S;
    int64_t ldf;
    const float* cat;
    const float* agg;
    const float* dz;
    const float* dh;
    float* part;          // job-major: job j starts at part_off[j], layout [S][M_j][E]
    int64_t part_off[PCG_MAX_REL + 1];
};

template <int CE>
__global__ void __launch_bounds__(256) k_dense_wgrad(WgradP p) {
    __shared__ float As[32][33];
    extern __shared__ float Gs[];                       // [32][Ep]
    const int F = p.F, E = p.E, R = p.R, B = p.B;
    const int K2 = F + R * E, RE = R * E, Ep = CE * 64;
    const int job = blockIdx.y, split = blockIdx.z;
    const int M = job == 0 ? K2 : 2 * F;
    const int m0 = blockIdx.x * 32;
    if (m0 >= M) return;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int per = (B + p.S - 1) / p.S;
    const int ib = split * per, ie = min(B, ib + per);
    const int r = job - 1;
    float acc[2][CE][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < CE; ++c) acc[a][c][0] = acc[a][c][1] = acc[a][c][2] = acc[a][c][3] = 0.f;
    for (int i0 = ib; i0 < ie; i0 += 32) {
        __syncthreads();
        for (int t = tid; t < 32 * 32; t += 256) {      // A chunk: 32 targets x 32 rows of the weight
            const int ii = t >> 5, mm = t & 31;
            const int i = i0 + ii, m = m0 + mm;
            float a = 0.f;
            if (i < ie && m < M) {
                if (job == 0 || m < F) a = p.cat[(int64_t)i * K2 + m];
                else a = p.agg[((int64_t)r * B + i) * p.ldf + (m - F)];
            }
            As[ii][mm] = a;
        }
        for (int t = tid; t < 32 * Ep; t += 256) {      // G chunk: 32 targets x E columns
            const int ii = t / Ep, n = t - ii * Ep;
            const int i = i0 + ii;
            float g = 0.f;
            if (i < ie && n < E) g = job == 0 ? p.dz[(int64_t)i * E + n] : p.dh[(int64_t)i * RE + r * E + n];
            Gs[t] = g;
        }
        __syncthreads();
#pragma unroll 4
        for (int ii = 0; ii < 32; ++ii) {
            const float a0 = As[ii][ty * 2], a1 = As[ii][ty * 2 + 1];
            const float* gr = Gs + ii * Ep + tx * 4;
#pragma unroll
            for (int c = 0; c < CE; ++c) {
                const float4 g4 = *reinterpret_cast<const float4*>(gr + c * 64);
                acc[0][c][0] = fmaf(a0, g4.x, acc[0][c][0]); acc[0][c][1] = fmaf(a0, g4.y, acc[0][c][1]);
                acc[0][c][2] = fmaf(a0, g4.z, acc[0][c][2]); acc[0][c][3] = fmaf(a0, g4.w, acc[0][c][3]);
                acc[1][c][0] = fmaf(a1, g4.x, acc[1][c][0]); acc[1][c][1] = fmaf(a1, g4.y, acc[1][c][1]);
                acc[1][c][2] = fmaf(a1, g4.z, acc[1][c][2]); acc[1][c][3] = fmaf(a1, g4.w, acc[1][c][3]);
            }
        }
    }
    float* dst = p.part + p.part_off[job] + (int64_t)split * M * E;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int m = m0 + ty * 2 + a;
        if (m >= M) continue;
#pragma unroll
        for (int c = 0; c < CE; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int n = c * 64 + tx * 4 + q;
                if (n < E) dst[(int64_t)m * E + n] = acc[a][c][q];
            }
    }
}

// grad[j][x] = sum_s part[j][s][x], splits added in index order (deterministic)
struct ReduceP {
    int n_jobs, S;
    int64_t size[PCG_MAX_REL + 1];
    int64_t part_off[PCG_MAX_REL + 1];
    const float* part;
    float* grad[PCG_MAX_REL + 1];
};

__global__ void k_dense_reduce(ReduceP p) {
    const int job = blockIdx.y;
    const int64_t n = p.size[job];
    const float* src = p.part + p.part_off[job];
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int q = 0; q < p.S; ++q) s += src[(int64_t)q * n + x];
        p.grad[job][x] = s;
    }
}

// ------------------------------------------------------------------------------------------- C ABI
static int dense_splits(int B) { return B >= 4096 ? 16 : (B >= 512 ? 8 : (B >= 64 ? 2 : 1)); }

extern "C" size_t pcg_dense_bwd_scratch_floats(int B, int R, int F, int E) {
    const size_t K2 = (size_t)F + (size_t)R * E;
    const size_t S = dense_splits(B);
    return (size_t)B * E + (size_t)B * R * E + S * (K2 * E + (size_t)R * 2 * F * E) + 64;
}

template <int TM>
static int launch_fwd(const DenseP& p, cudaStream_t stream) {
    const int CE = (p.E + 63) / 64;
    const int K2 = p.F + p.R * p.E;
    size_t floats = (size_t)TM * (K2 + 1) + (size_t)TM * (p.F + 1);
    floats = (floats + 3) & ~(size_t)3;                        // weight chunk 16-byte aligned
    const size_t smem = (floats + (size_t)DENSE_KC * CE * 64) * 4;
    const int grid = (p.B + TM - 1) / TM;
#define PCG_FWD(CEv)                                                                                           \
    do {                                                                                                       \
        if (smem > 48 * 1024) {                                                                                \
            cudaError_t e = cudaFuncSetAttribute(k_dense_fwd<TM, CEv>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 (int)smem);                                                   \
            if (e != cudaSuccess) { pcg_set_error("pcg_dense_fwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; } \
        }                                                                                                      \
        k_dense_fwd<TM, CEv><<<grid, TM * 16, smem, stream>>>(p);                                              \
    } while (0)
    switch (CE) {
        case 1: PCG_FWD(1); break;
        case 2: PCG_FWD(2); break;
        case 3: PCG_FWD(3); break;
        default: PCG_FWD(4); break;
    }
#undef PCG_FWD
    return 0;
}

extern "C" int pcg_dense_fwd(const float* feat, int64_t ldf, int F, const int32_t* targets, int B, int R, int E,
                             const float* agg, const float* const* w_intra_host, const float* w_inter, float* cat,
                             float* out, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B == 0) return 0;
    PCG_REQUIRE(feat && targets && agg && w_intra_host && w_inter && cat && out, "pcg_dense_fwd: null pointer");
    PCG_REQUIRE(R >= 1 && R <= PCG_MAX_REL && E >= 1 && E <= DENSE_MAX_E && F >= 1, "pcg_dense_fwd: bad sizes R=%d E=%d F=%d",
                R, E, F);
    DenseP p;
    p.feat = feat; p.ldf = ldf; p.F = F; p.B = B; p.R = R; p.E = E; p.targets = targets; p.agg = agg;
    for (int r = 0; r < PCG_MAX_REL; ++r) p.w_intra[r] = r < R ? w_intra_host[r] : nullptr;
    p.w_inter = w_inter; p.cat = cat; p.out = out;
    // the tile (16 rows of F + R*E floats + a weight chunk) must fit shared memory
    const size_t need16 = ((size_t)16 * (F + R * E + 1) + 16 * (F + 1) + (size_t)DENSE_KC * ((E + 63) / 64) * 64) * 4;
    PCG_REQUIRE(need16 <= 200 * 1024, "pcg_dense_fwd: F + R*E = %d too wide for one shared-memory tile", F + R * E);
    int rc = B >= 2048 ? launch_fwd<16>(p, stream) : launch_fwd<8>(p, stream);
    if (rc) return rc;
    return pcg_check_launch("pcg_dense_fwd");
}

extern "C" int pcg_dense_bwd(int64_t ldf, int F, int B, int R, int E, const float* agg, const float* w_inter,
                             const float* cat, const float* out, const float* d_out, float* const* d_w_intra_host,
                             float* d_w_inter, float* scratch, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCG_REQUIRE(R >= 1 && R <= PCG_MAX_REL && E >= 1 && E <= DENSE_MAX_E && F >= 1, "pcg_dense_bwd: bad sizes");
    PCG_REQUIRE(agg && w_inter && cat && out && d_out && d_w_intra_host && d_w_inter && scratch,
                "pcg_dense_bwd: null pointer");
    const int K2 = F + R * E, RE = R * E, CE = (E + 63) / 64;
    const int S = dense_splits(B);
    float* dz = scratch;
    float* dh = dz + (size_t)B * E;
    float* part = dh + (size_t)B * RE;
    part = (float*)(((uintptr_t)part + 15) & ~(uintptr_t)15);
    if (B > 0) {
        DenseBwdP q;
        q.F = F; q.B = B; q.R = R; q.E = E; q.cat = cat; q.out = out; q.d_out = d_out; q.w_inter = w_inter;
        q.dz = dz; q.dh = dh;
        constexpr int TM = 16;
        const size_t smem = ((size_t)TM * (E + 1) + (size_t)64 * (E + 1)) * 4;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k_dense_bwd_rows<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { pcg_set_error("pcg_dense_bwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        }
        k_dense_bwd_rows<TM><<<(B + TM - 1) / TM, TM * 16, smem, stream>>>(q);
    }
    WgradP wp;
    wp.F = F; wp.B = B; wp.R = R; wp.E = E; wp.S = S; wp.ldf = ldf; wp.cat = cat; wp.agg = agg; wp.dz = dz; wp.dh = dh;
    wp.part = part;
    ReduceP rp;
    rp.n_jobs = R + 1; rp.S = S; rp.part = part;
    int64_t off = 0;
    for (int j = 0; j <= R; ++j) {
        const int64_t M = j == 0 ? K2 : 2 * F;
        wp.part_off[j] = off;
        rp.part_off[j] = off;
        rp.size[j] = M * E;
        rp.grad[j] = j == 0 ? d_w_inter : d_w_intra_host[j - 1];
        off += (int64_t)S * M * E;
    }
    const int mt = (max(K2, 2 * F) + 31) / 32;
    dim3 grid(mt, R + 1, S);
    const size_t gs = (size_t)32 * CE * 64 * 4;
    switch (CE) {
        case 1: k_dense_wgrad<1><<<grid, 256, gs, stream>>>(wp); break;
        case 2: k_dense_wgrad<2><<<grid, 256, gs, stream>>>(wp); break;
        case 3: k_dense_wgrad<3><<<grid, 256, gs, stream>>>(wp); break;
        default: k_dense_wgrad<4><<<grid, 256, gs, stream>>>(wp); break;
    }
    dim3 rg((unsigned)((max(K2, 2 * F) * E + 255) / 256), R + 1);
    k_dense_reduce<<<rg, 256, 0, stream>>>(rp);
    return pcg_check_launch("pcg_dense_bwd");
}
