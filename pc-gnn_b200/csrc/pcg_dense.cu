// Relation transforms + inter-relation combine (forward and weight-gradient backward), fp32 FFMA.
//
// Reference (/root/reference/src/layers.py): per relation r  to_feats_r = relu(cat(self, agg_r) @ W_r)
// (:616-629), then combined = relu(cat(self, to_feats_1..R) @ W).t() (:273-289), each a separate
// cat / mm / relu library call plus its autograd twin. Here every step is one tiled fp32 GEMM whose
// operand loaders do the concatenation / gathers / activation masks on the fly and whose epilogue does
// the ReLU, the transposed store or the split-K partial store:
//   forward   relations GEMM (grid.z = relation)  ->  combine GEMM                    (+ a self copy)
//   backward  dH GEMM (also emits dZ)  ->  weight-gradient GEMM for W and all W_r, split over the
//             batch  ->  fixed-order reduction of the partials (deterministic)
//
// fp32 on the CUDA cores on purpose: the parity bar is 1e-5 relative, which tf32/bf16 tensor-core
// paths do not meet, and the dense part is ~50 MFLOP (C2) to ~1 GFLOP (C3) per step.
#include "pcg_common.cuh"

#define DENSE_MAX_E 256

// ------------------------------------------------------------------------------------------------
// C[M,N] = A[M,K] * B[K,N] for the skinny shapes of this path (N = E or R*E, K <= a few hundred, M = the
// batch): TM x 64 output tiles (TM = 16 / 32 / 64 by batch size, so that a 1024-target batch still spreads
// over 64+ CTAs), K in chunks of 32 through a 3-stage cp.async pipeline. The operands are functors:
// addr(z, row, k) / addr(z, k, col) give the element's address (4-byte cp.async, zero fill out of range), so
// the concatenations, gathers and transposes of the reference's cat / index / .t() calls cost nothing; an
// A operand that must be COMPUTED (A_REG) is loaded through registers instead. A_MC / B_NC say which index
// is contiguous in memory so that the tile loads coalesce.
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool ok) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = ok ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int TM, bool A_REG, bool A_MC, bool B_NC, class AL, class BL, class EP>
__global__ void __launch_bounds__(256) k_gemm_p(int M, int N, int K, AL al, BL bl, EP ep) {
    constexpr int KC = 32, ST = 3, RPT = TM / 16;
    __shared__ __align__(16) float As[ST][KC][TM + 4];   // [k][m]
    __shared__ __align__(16) float Bs[ST][KC][64 + 4];   // [k][n]
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * 64, z = blockIdx.z;
    al.bind(z); bl.bind(z); ep.bind(z);  // resolve per-z pointers / offsets once (z is block-uniform)
    if (ep.skip(z, m0)) return;
    int kb, ke;
    ep.k_range(z, K, kb, ke);
    const int nt = (ke - kb + KC - 1) / KC;
    auto issue = [&](int t) {
        const int buf = t % ST, k0 = kb + t * KC;
#pragma unroll
        for (int j = 0; j < 8; ++j) {                    // B tile: 32 x 64
            const int nn = B_NC ? (tid & 63) : (tid >> 5) + 8 * j;
            const int kk = B_NC ? (tid >> 6) + 4 * j : (tid & 31);
            const bool ok = n0 + nn < N && k0 + kk < ke;
            cp_async4(&Bs[buf][kk][nn], ok ? bl.addr(z, k0 + kk, n0 + nn) : bl.any(), ok);
        }
#pragma unroll
        for (int j = 0; j < TM * KC / 256; ++j) {        // A tile: TM x 32
            const int mm = A_MC ? (tid % TM) : (tid >> 5) + 8 * j;
            const int kk = A_MC ? (tid / TM) + (256 / TM) * j : (tid & 31);
            const bool ok = m0 + mm < M && k0 + kk < ke;
            if constexpr (A_REG) As[buf][kk][mm] = ok ? al(z, m0 + mm, k0 + kk) : 0.f;
            else cp_async4(&As[buf][kk][mm], ok ? al.addr(z, m0 + mm, k0 + kk) : al.any(), ok);
        }
    };
    float acc[RPT][4];
#pragma unroll
    for (int a = 0; a < RPT; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll
    for (int s = 0; s < ST - 1; ++s) {
        if (s < nt) issue(s);
        cp_async_commit();
    }
    for (int t = 0; t < nt; ++t) {
        cp_async_wait<ST - 2>();          // this thread's copies of stage t have landed
        __syncthreads();                  // ... everybody's; and stage t-1 is no longer being read
        if (t + ST - 1 < nt) issue(t + ST - 1);
        cp_async_commit();                // (possibly empty: keeps the group count uniform)
        const int cur = t % ST;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
            float av[RPT];
#pragma unroll
            for (int a = 0; a < RPT; ++a) av[a] = As[cur][kk][ty * RPT + a];
#pragma unroll
            for (int a = 0; a < RPT; ++a) {
                acc[a][0] = fmaf(av[a], b4.x, acc[a][0]);
                acc[a][1] = fmaf(av[a], b4.y, acc[a][1]);
                acc[a][2] = fmaf(av[a], b4.z, acc[a][2]);
                acc[a][3] = fmaf(av[a], b4.w, acc[a][3]);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < RPT; ++a) {
        const int m = m0 + ty * RPT + a;
        if (m >= M) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int n = n0 + tx * 4 + b;
            if (n < N) ep(z, m, n, acc[a][b]);
        }
    }
}

// Variant for operands whose contiguous index is the one shared memory keeps contiguous (A: consecutive m,
// B: consecutive n) and whose rows are 16-byte aligned: 64 x 64 tiles, 16-byte cp.async (4 copies per thread and
// stage instead of 16 scalar loads), two stages. Used for the weight gradients when F and E are multiples of 4.
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, int bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}

template <class AL, class BL, class EP>
__global__ void __launch_bounds__(256) k_gemm_v(int M, int N, int K, AL al, BL bl, EP ep) {
    constexpr int TM = 64, KC = 32, ST = 2;
    __shared__ __align__(16) float As[ST][KC][TM + 4];   // [k][m]
    __shared__ __align__(16) float Bs[ST][KC][64 + 4];   // [k][n]
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * 64, z = blockIdx.z;
    al.bind(z); bl.bind(z); ep.bind(z);
    if (ep.skip(z, m0)) return;
    int kb, ke;
    ep.k_range(z, K, kb, ke);
    const int nt = (ke - kb + KC - 1) / KC;
    auto issue = [&](int t) {
        const int buf = t % ST, k0 = kb + t * KC;
#pragma unroll
        for (int j = 0; j < 2; ++j) {                    // 32 rows x 16 chunks of 4 floats, for A and for B
            const int c = tid + 256 * j;
            const int kk = c >> 4, q4 = (c & 15) * 4;
            const bool krow = k0 + kk < ke;
            const int ba = krow ? min(16, max(0, (M - (m0 + q4)) * 4)) : 0;
            const int bb = krow ? min(16, max(0, (N - (n0 + q4)) * 4)) : 0;
            cp_async16(&As[buf][kk][q4], ba > 0 ? al.addr(z, m0 + q4, k0 + kk) : al.any(), ba);
            cp_async16(&Bs[buf][kk][q4], bb > 0 ? bl.addr(z, k0 + kk, n0 + q4) : bl.any(), bb);
        }
    };
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    if (nt > 0) issue(0);
    cp_async_commit();
    for (int t = 0; t < nt; ++t) {
        if (t + 1 < nt) issue(t + 1);     // buffer (t+1)%2 was last read in iteration t-1, behind the barrier below
        cp_async_commit();
        cp_async_wait<1>();               // stage t has landed (this thread's copies)
        __syncthreads();
        const int cur = t % ST;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
            acc[0][0] = fmaf(a4.x, b4.x, acc[0][0]); acc[0][1] = fmaf(a4.x, b4.y, acc[0][1]);
            acc[0][2] = fmaf(a4.x, b4.z, acc[0][2]); acc[0][3] = fmaf(a4.x, b4.w, acc[0][3]);
            acc[1][0] = fmaf(a4.y, b4.x, acc[1][0]); acc[1][1] = fmaf(a4.y, b4.y, acc[1][1]);
            acc[1][2] = fmaf(a4.y, b4.z, acc[1][2]); acc[1][3] = fmaf(a4.y, b4.w, acc[1][3]);
            acc[2][0] = fmaf(a4.z, b4.x, acc[2][0]); acc[2][1] = fmaf(a4.z, b4.y, acc[2][1]);
            acc[2][2] = fmaf(a4.z, b4.z, acc[2][2]); acc[2][3] = fmaf(a4.z, b4.w, acc[2][3]);
            acc[3][0] = fmaf(a4.w, b4.x, acc[3][0]); acc[3][1] = fmaf(a4.w, b4.y, acc[3][1]);
            acc[3][2] = fmaf(a4.w, b4.z, acc[3][2]); acc[3][3] = fmaf(a4.w, b4.w, acc[3][3]);
        }
        __syncthreads();                  // everybody is done with buffer `cur` before iteration t+1 refills it
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int m = m0 + ty * 4 + a;
        if (m >= M) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int n = n0 + tx * 4 + b;
            if (n < N) ep(z, m, n, acc[a][b]);
        }
    }
}

// Variant of k_gemm_p for the forward GEMMs of small batches when rows are 16-byte aligned (F, E multiples of 4):
// A rows are contiguous along k (cat / feature / aggregate rows), so the A tile is kept m-major in shared
// memory and both operands arrive in 16-byte cp.async chunks: 16 x 64 output tiles, 3 stages over K.
template <class AL, class BL, class EP>
__global__ void __launch_bounds__(256) k_gemm_pv(int M, int N, int K, AL al, BL bl, EP ep) {
    constexpr int TM = 16, KC = 32, ST = 3;
    __shared__ __align__(16) float As[ST][TM][KC + 4];   // [m][k]
    __shared__ __align__(16) float Bs[ST][KC][64 + 4];   // [k][n]
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * 64, z = blockIdx.z;
    al.bind(z); bl.bind(z); ep.bind(z);
    if (ep.skip(z, m0)) return;
    int kb, ke;
    ep.k_range(z, K, kb, ke);
    const int nt = (ke - kb + KC - 1) / KC;
    auto issue = [&](int t) {
        const int buf = t % ST, k0 = kb + t * KC;
#pragma unroll
        for (int j = 0; j < 2; ++j) {                    // B: 32 k-rows x 16 chunks
            const int c = tid + 256 * j;
            const int kk = c >> 4, q4 = (c & 15) * 4;
            const int bb = k0 + kk < ke ? min(16, max(0, (N - (n0 + q4)) * 4)) : 0;
            cp_async16(&Bs[buf][kk][q4], bb > 0 ? bl.addr(z, k0 + kk, n0 + q4) : bl.any(), bb);
        }
        if (tid < 128) {                                 // A: 16 rows x 8 chunks
            const int mm = tid >> 3, q4 = (tid & 7) * 4;
            const int ba = m0 + mm < M ? min(16, max(0, (ke - (k0 + q4)) * 4)) : 0;
            cp_async16(&As[buf][mm][q4], ba > 0 ? al.addr(z, m0 + mm, k0 + q4) : al.any(), ba);
        }
    };
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < ST - 1; ++s) {
        if (s < nt) issue(s);
        cp_async_commit();
    }
    for (int t = 0; t < nt; ++t) {
        cp_async_wait<ST - 2>();
        __syncthreads();
        if (t + ST - 1 < nt) issue(t + ST - 1);
        cp_async_commit();
        const int cur = t % ST;
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[cur][ty][kk]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][kk + 1][tx * 4]);
            const float4 b2 = *reinterpret_cast<const float4*>(&Bs[cur][kk + 2][tx * 4]);
            const float4 b3 = *reinterpret_cast<const float4*>(&Bs[cur][kk + 3][tx * 4]);
            acc[0] = fmaf(a4.x, b0.x, acc[0]); acc[1] = fmaf(a4.x, b0.y, acc[1]); acc[2] = fmaf(a4.x, b0.z, acc[2]); acc[3] = fmaf(a4.x, b0.w, acc[3]);
            acc[0] = fmaf(a4.y, b1.x, acc[0]); acc[1] = fmaf(a4.y, b1.y, acc[1]); acc[2] = fmaf(a4.y, b1.z, acc[2]); acc[3] = fmaf(a4.y, b1.w, acc[3]);
            acc[0] = fmaf(a4.z, b2.x, acc[0]); acc[1] = fmaf(a4.z, b2.y, acc[1]); acc[2] = fmaf(a4.z, b2.z, acc[2]); acc[3] = fmaf(a4.z, b2.w, acc[3]);
            acc[0] = fmaf(a4.w, b3.x, acc[0]); acc[1] = fmaf(a4.w, b3.y, acc[1]); acc[2] = fmaf(a4.w, b3.z, acc[2]); acc[3] = fmaf(a4.w, b3.w, acc[3]);
        }
    }
    const int m = m0 + ty;
    if (m < M) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int n = n0 + tx * 4 + b;
            if (n < N) ep(z, m, n, acc[b]);
        }
    }
}

// rows of the batch per CTA: small batches need small tiles to fill the GPU
static int pick_tm(int M) { return M <= 2048 ? 16 : 64; }
#define PCG_GEMM(A_REG, A_MC, B_NC, tm, grid_m_rows, gy, gz, M, N, K, a, b, ep)                                   \
    do {                                                                                                          \
        const int tm_ = (tm);                                                                                     \
        dim3 g_((unsigned)(((grid_m_rows) + tm_ - 1) / tm_), (unsigned)(gy), (unsigned)(gz));                      \
        if (tm_ == 16) k_gemm_p<16, A_REG, A_MC, B_NC><<<g_, 256, 0, stream>>>(M, N, K, a, b, ep);                \
        else k_gemm64<A_MC, B_NC><<<g_, 256, 0, stream>>>(M, N, K, a, b, ep);                                     \
    } while (0)

// ------------------------------------------------------------------------------------------------
// C[M,N] = A[M,K] * B[K,N], 64x64x16 tiles, 256 threads, 4x4 outputs per thread, register-prefetched
// double buffering. A and B are functors (z, row, k) -> float / (z, k, col) -> float, only called in
// range; A_MC / B_NC say which index is contiguous in memory so the tile loads coalesce.
//   A_MC: consecutive m contiguous (else consecutive k);  B_NC: consecutive n contiguous (else k).
template <bool A_MC, bool B_NC, class AL, class BL, class EP>
__global__ void __launch_bounds__(256) k_gemm64(int M, int N, int K, AL al, BL bl, EP ep) {
    __shared__ __align__(16) float As[2][16][68];   // [k][m]
    __shared__ __align__(16) float Bs[2][16][68];   // [k][n]
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64, z = blockIdx.z;
    al.bind(z); bl.bind(z); ep.bind(z);  // resolve per-z pointers / offsets once (z is block-uniform)
    if (ep.skip(z, m0)) return;
    int kb, ke;
    ep.k_range(z, K, kb, ke);
    // this thread's 4 A and 4 B tile elements: fixed (row, k) offsets inside every k-tile
    const int a_mm = A_MC ? (tid & 63) : (tid >> 4);      // + 16*j when !A_MC
    const int a_kk = A_MC ? (tid >> 6) : (tid & 15);      // + 4*j  when  A_MC
    const int b_nn = B_NC ? (tid & 63) : (tid >> 4);
    const int b_kk = B_NC ? (tid >> 6) : (tid & 15);
    float ra0, ra1, ra2, ra3, rb0, rb1, rb2, rb3;
#define PCG_LOAD_A(j, dst)                                                     \
    {                                                                          \
        const int mm = A_MC ? a_mm : a_mm + 16 * (j);                          \
        const int k = k0 + (A_MC ? a_kk + 4 * (j) : a_kk);                     \
        dst = (m0 + mm < M && k < ke) ? al(z, m0 + mm, k) : 0.f;               \
    }
#define PCG_LOAD_B(j, dst)                                                     \
    {                                                                          \
        const int nn = B_NC ? b_nn : b_nn + 16 * (j);                          \
        const int k = k0 + (B_NC ? b_kk + 4 * (j) : b_kk);                     \
        dst = (n0 + nn < N && k < ke) ? bl(z, k, n0 + nn) : 0.f;               \
    }
#define PCG_LOAD_TILE(kstart)                                                  \
    {                                                                          \
        const int k0 = (kstart);                                               \
        PCG_LOAD_A(0, ra0) PCG_LOAD_A(1, ra1) PCG_LOAD_A(2, ra2) PCG_LOAD_A(3, ra3) \
        PCG_LOAD_B(0, rb0) PCG_LOAD_B(1, rb1) PCG_LOAD_B(2, rb2) PCG_LOAD_B(3, rb3) \
    }
#define PCG_STASH_A(j, src) As[buf][A_MC ? a_kk + 4 * (j) : a_kk][A_MC ? a_mm : a_mm + 16 * (j)] = src;
#define PCG_STASH_B(j, src) Bs[buf][B_NC ? b_kk + 4 * (j) : b_kk][B_NC ? b_nn : b_nn + 16 * (j)] = src;
#define PCG_STASH_TILE(bufidx)                                                 \
    {                                                                          \
        const int buf = (bufidx);                                              \
        PCG_STASH_A(0, ra0) PCG_STASH_A(1, ra1) PCG_STASH_A(2, ra2) PCG_STASH_A(3, ra3) \
        PCG_STASH_B(0, rb0) PCG_STASH_B(1, rb1) PCG_STASH_B(2, rb2) PCG_STASH_B(3, rb3) \
    }
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const int nt = (ke - kb + 15) / 16;
    if (nt > 0) {
        PCG_LOAD_TILE(kb)
        PCG_STASH_TILE(0)
    }
    __syncthreads();
    for (int t = 0; t < nt; ++t) {
        const int cur = t & 1;
        if (t + 1 < nt) PCG_LOAD_TILE(kb + (t + 1) * 16)
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
            acc[0][0] = fmaf(a4.x, b4.x, acc[0][0]); acc[0][1] = fmaf(a4.x, b4.y, acc[0][1]);
            acc[0][2] = fmaf(a4.x, b4.z, acc[0][2]); acc[0][3] = fmaf(a4.x, b4.w, acc[0][3]);
            acc[1][0] = fmaf(a4.y, b4.x, acc[1][0]); acc[1][1] = fmaf(a4.y, b4.y, acc[1][1]);
            acc[1][2] = fmaf(a4.y, b4.z, acc[1][2]); acc[1][3] = fmaf(a4.y, b4.w, acc[1][3]);
            acc[2][0] = fmaf(a4.z, b4.x, acc[2][0]); acc[2][1] = fmaf(a4.z, b4.y, acc[2][1]);
            acc[2][2] = fmaf(a4.z, b4.z, acc[2][2]); acc[2][3] = fmaf(a4.z, b4.w, acc[2][3]);
            acc[3][0] = fmaf(a4.w, b4.x, acc[3][0]); acc[3][1] = fmaf(a4.w, b4.y, acc[3][1]);
            acc[3][2] = fmaf(a4.w, b4.z, acc[3][2]); acc[3][3] = fmaf(a4.w, b4.w, acc[3][3]);
        }
        if (t + 1 < nt) PCG_STASH_TILE(cur ^ 1)
        __syncthreads();
    }
#undef PCG_LOAD_A
#undef PCG_LOAD_B
#undef PCG_LOAD_TILE
#undef PCG_STASH_A
#undef PCG_STASH_B
#undef PCG_STASH_TILE
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int m = m0 + ty * 4 + a;
        if (m >= M) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int n = n0 + tx * 4 + b;
            if (n < N) ep(z, m, n, acc[a][b]);
        }
    }
}

// ---------------------------------------------------------------------------------------- forward
struct FwdRelA {    // A(r, i, k) = [feat[targets[i]] | agg[rep(r*B + i)]][k]  (rep: row of the target's first occurrence)
    const float* feat; const float* agg; const int32_t* targets; const int32_t* rep; int64_t ldf; int B, F;
    __device__ void bind(int) {}
    __device__ const float* any() const { return feat; }
    __device__ const float* addr(int r, int i, int k) const {
        if (k < F) return feat + (int64_t)__ldg(targets + i) * ldf + k;
        const int64_t w = (int64_t)r * B + i;
        return agg + (rep ? (int64_t)__ldg(rep + w) : w) * ldf + (k - F);
    }
    __device__ float operator()(int r, int i, int k) const { return __ldg(addr(r, i, k)); }
};
struct FwdRelB {    // B(r, k, n) = W_r[k][n]
    const float* w[PCG_MAX_REL]; int E; const float* cur;
    __device__ void bind(int r) { cur = w[r]; }
    __device__ const float* any() const { return cur; }
    __device__ const float* addr(int, int k, int n) const { return cur + (int64_t)k * E + n; }
    __device__ float operator()(int r, int k, int n) const { return __ldg(addr(r, k, n)); }
};
struct FwdRelEp {   // cat[i][F + r*E + n] = relu(v); relation 0 also copies the self part cat[i][:F] when F <= E
    float* cat; int K2, F, E;
    const float* self_feat; const int32_t* targets; int64_t ldf;      // self_feat == NULL: k_copy_self does it
    __device__ void bind(int) {}
    __device__ void k_range(int, int K, int& kb, int& ke) const { kb = 0; ke = K; }
    __device__ bool skip(int, int) const { return false; }
    __device__ void operator()(int r, int i, int n, float v) const {
        cat[(int64_t)i * K2 + F + r * E + n] = fmaxf(v, 0.f);
        if (self_feat && r == 0 && n < F) cat[(int64_t)i * K2 + n] = __ldg(self_feat + (int64_t)__ldg(targets + i) * ldf + n);
    }
};
__global__ void k_copy_self(const float* __restrict__ feat, int64_t ldf, const int32_t* __restrict__ targets, int B,
                            int F, int K2, float* __restrict__ cat) {   // cat[i][:F] = feat[targets[i]][:F]
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * F) return;
    const int i = (int)(t / F), f = (int)(t - (int64_t)i * F);
    cat[(int64_t)i * K2 + f] = __ldg(feat + (int64_t)__ldg(targets + i) * ldf + f);
}
struct CombA {      // A(_, i, k) = cat[i][k]
    const float* cat; int K2;
    __device__ void bind(int) {}
    __device__ const float* any() const { return cat; }
    __device__ const float* addr(int, int i, int k) const { return cat + (int64_t)i * K2 + k; }
    __device__ float operator()(int z, int i, int k) const { return *addr(z, i, k); }
};
struct CombB {
    const float* w; int E;
    __device__ void bind(int) {}
    __device__ const float* any() const { return w; }
    __device__ const float* addr(int, int k, int n) const { return w + (int64_t)k * E + n; }
    __device__ float operator()(int z, int k, int n) const { return __ldg(addr(z, k, n)); }
};
struct CombEp {     // out[n][i] = relu(v)   ([E,B], the reference's transposed layout, layers.py:289)
    float* out; int B;
    __device__ void bind(int) {}
    __device__ void k_range(int, int K, int& kb, int& ke) const { kb = 0; ke = K; }
    __device__ bool skip(int, int) const { return false; }
    __device__ void operator()(int, int i, int n, float v) const { out[(int64_t)n * B + i] = fmaxf(v, 0.f); }
};

// --------------------------------------------------------------------------------------- backward
struct BwdHA {      // A(_, i, e) = dZ[i][e] = d_out[e][i] * (out[e][i] > 0); also written out as dz[i][e]
    const float* d_out; const float* out; float* dz; int B, E;
    __device__ void bind(int) {}
    __device__ float operator()(int, int i, int e) const {
        const float g = out[(int64_t)e * B + i] > 0.f ? d_out[(int64_t)e * B + i] : 0.f;
        if (blockIdx.y == 0) dz[(int64_t)i * E + e] = g;
        return g;
    }
};
struct BwdHB {      // B(_, e, c) = W[F + c][e]
    const float* w; int F, E;
    __device__ void bind(int) {}
    __device__ const float* any() const { return w; }
    __device__ const float* addr(int, int e, int c) const { return w + (int64_t)(F + c) * E + e; }
    __device__ float operator()(int z, int e, int c) const { return __ldg(addr(z, e, c)); }
};
struct BwdHEp {     // dH[i][c] = v * (cat[i][F + c] > 0)
    const float* cat; float* dh; int K2, F, RE;
    __device__ void bind(int) {}
    __device__ void k_range(int, int K, int& kb, int& ke) const { kb = 0; ke = K; }
    __device__ bool skip(int, int) const { return false; }
    __device__ void operator()(int, int i, int c, float v) const {
        dh[(int64_t)i * RE + c] = cat[(int64_t)i * K2 + F + c] > 0.f ? v : 0.f;
    }
};
// weight gradients: z = job * S + split; job 0 -> W (rows m < K2), job 1+r -> W_r (rows m < 2F)
struct WgA {        // A(z, m, i) = X_job[i][m]
    const float* cat; const float* agg; const int32_t* rep; int64_t ldf; int B, F, K2, S; int job;
    __device__ void bind(int z) { job = z / S; }
    __device__ const float* any() const { return cat; }
    __device__ const float* addr(int, int m, int i) const {      // rows beyond the job's matrix are dropped by WgEp
        if (job == 0) return m < K2 ? cat + (int64_t)i * K2 + m : cat;
        if (m >= 2 * F) return cat;
        if (m < F) return cat + (int64_t)i * K2 + m;
        const int64_t w = (int64_t)(job - 1) * B + i;
        return agg + (rep ? (int64_t)__ldg(rep + w) : w) * ldf + (m - F);
    }
    __device__ float operator()(int z, int m, int i) const {
        if ((job == 0 && m >= K2) || (job > 0 && m >= 2 * F)) return 0.f;
        return *addr(z, m, i);
    }
};
struct WgB {        // B(z, i, n) = dZ[i][n] (job 0) or dH[i][r*E + n]
    const float* dz; const float* dh; int E, RE, S; int job;
    __device__ void bind(int z) { job = z / S; }
    __device__ const float* any() const { return dz; }
    __device__ const float* addr(int, int i, int n) const {
        return job == 0 ? dz + (int64_t)i * E + n : dh + (int64_t)i * RE + (job - 1) * E + n;
    }
    __device__ float operator()(int z, int i, int n) const { return *addr(z, i, n); }
};
struct WgEp {       // part[job][split][m][n] = v over the split's slice of the batch
    float* part; int64_t off[PCG_MAX_REL + 1]; int rows[PCG_MAX_REL + 1]; int E, S, B;
    float* dst; int nrows, split;
    __device__ void bind(int z) {
        const int job = z / S;
        split = z % S;
        nrows = rows[job];
        dst = part + off[job] + (int64_t)split * nrows * E;
    }
    __device__ bool skip(int, int m0) const { return m0 >= nrows; }
    __device__ void k_range(int, int K, int& kb, int& ke) const {
        const int per = (B + S - 1) / S;
        kb = min(K, split * per);
        ke = min(K, kb + per);
    }
    __device__ void operator()(int, int m, int n, float v) const {
        if (m < nrows) dst[(int64_t)m * E + n] = v;
    }
};

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
// batch slices of the weight-gradient GEMM: enough CTAs to fill the GPU, slices of >= 64 targets
static int dense_splits(int B, int R, int F, int E) {
    const int Mw = F + R * E > 2 * F ? F + R * E : 2 * F;
    const int tiles = ((Mw + 63) / 64) * ((E + 63) / 64) * (R + 1);
    int s = (2 * 148 + tiles - 1) / tiles;
    const int max_s = (B + 63) / 64;
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    return s > 32 ? 32 : s;
}

extern "C" size_t pcg_dense_bwd_scratch_floats(int B, int R, int F, int E) {
    const size_t K2 = (size_t)F + (size_t)R * E;
    const size_t S = dense_splits(B, R, F, E);
    return (size_t)B * E + (size_t)B * R * E + S * (K2 * E + (size_t)R * 2 * F * E) + 64;
}

extern "C" int pcg_dense_fwd(const float* feat, int64_t ldf, int F, const int32_t* targets, int B, int R, int E,
                             const float* agg, const int32_t* agg_rep, const float* const* w_intra_host, const float* w_inter, float* cat,
                             float* out, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B == 0) return 0;
    PCG_REQUIRE(feat && targets && agg && w_intra_host && w_inter && cat && out, "pcg_dense_fwd: null pointer");
    PCG_REQUIRE(R >= 1 && R <= PCG_MAX_REL && E >= 1 && E <= DENSE_MAX_E && F >= 1, "pcg_dense_fwd: bad sizes R=%d E=%d F=%d",
                R, E, F);
    const int K2 = F + R * E;
    FwdRelA a{feat, agg, targets, agg_rep, ldf, B, F};
    FwdRelB b;
    for (int r = 0; r < PCG_MAX_REL; ++r) b.w[r] = r < R ? w_intra_host[r] : nullptr;
    b.E = E; b.cur = nullptr;
    const bool fold_self = F <= E;       // the relation-0 epilogue covers columns [0, E) of every target
    FwdRelEp ep{cat, K2, F, E, fold_self ? feat : nullptr, targets, ldf};
    if (!fold_self)
        k_copy_self<<<(unsigned)(((int64_t)B * F + 255) / 256), 256, 0, stream>>>(feat, ldf, targets, B, F, K2, cat);
    const int tm = pick_tm(B);
    // 16-byte cp.async operand tiles need every row start and every chunk 16-byte aligned
    bool vec = tm == 16 && F % 4 == 0 && E % 4 == 0 && ldf % 4 == 0 && ((uintptr_t)feat & 15) == 0 &&
               ((uintptr_t)agg & 15) == 0 && ((uintptr_t)cat & 15) == 0 && ((uintptr_t)w_inter & 15) == 0;
    for (int r = 0; r < R; ++r) vec = vec && ((uintptr_t)w_intra_host[r] & 15) == 0;
    if (vec) {
        dim3 g1((unsigned)((B + 15) / 16), (unsigned)((E + 63) / 64), (unsigned)R);
        k_gemm_pv<<<g1, 256, 0, stream>>>(B, E, 2 * F, a, b, ep);
    } else {
        PCG_GEMM(false, false, true, tm, B, (E + 63) / 64, R, B, E, 2 * F, a, b, ep);
    }
    CombA ca{cat, K2};
    CombB cb{w_inter, E};
    CombEp ce{out, B};
    if (vec) {
        dim3 g2((unsigned)((B + 15) / 16), (unsigned)((E + 63) / 64), 1);
        k_gemm_pv<<<g2, 256, 0, stream>>>(B, E, K2, ca, cb, ce);
    } else {
        PCG_GEMM(false, false, true, tm, B, (E + 63) / 64, 1, B, E, K2, ca, cb, ce);
    }
    return pcg_check_launch("pcg_dense_fwd");
}

extern "C" int pcg_dense_bwd(int64_t ldf, int F, int B, int R, int E, const float* agg, const int32_t* agg_rep,
                             const float* w_inter,
                             const float* cat, const float* out, const float* d_out, float* const* d_w_intra_host,
                             float* d_w_inter, float* scratch, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCG_REQUIRE(R >= 1 && R <= PCG_MAX_REL && E >= 1 && E <= DENSE_MAX_E && F >= 1, "pcg_dense_bwd: bad sizes");
    PCG_REQUIRE(agg && w_inter && cat && out && d_out && d_w_intra_host && d_w_inter && scratch,
                "pcg_dense_bwd: null pointer");
    const int K2 = F + R * E, RE = R * E;
    const int S = dense_splits(B, R, F, E);
    float* dz = scratch;
    float* dh = dz + (size_t)B * E;
    float* part = dh + (size_t)B * RE;
    if (B > 0) {
        BwdHA a{d_out, out, dz, B, E};
        BwdHB b{w_inter, F, E};
        BwdHEp ep{cat, dh, K2, F, RE};
        PCG_GEMM(true, true, false, pick_tm(B), B, (RE + 63) / 64, 1, B, RE, E, a, b, ep);
    }
    WgA wa{cat, agg, agg_rep, ldf, B, F, K2, S, 0};
    WgB wb{dz, dh, E, RE, S, 0};
    WgEp we;
    ReduceP rp;
    we.part = part; we.E = E; we.S = S; we.B = B; we.dst = nullptr; we.nrows = 0; we.split = 0;
    rp.n_jobs = R + 1; rp.S = S; rp.part = part;
    int64_t off = 0;
    for (int j = 0; j <= R; ++j) {
        const int64_t M = j == 0 ? K2 : 2 * F;
        we.off[j] = off; we.rows[j] = (int)M;
        rp.part_off[j] = off; rp.size[j] = M * E;
        rp.grad[j] = j == 0 ? d_w_inter : d_w_intra_host[j - 1];
        off += (int64_t)S * M * E;
    }
    const int Mw = K2 > 2 * F ? K2 : 2 * F;          // R == 1 with F > E: the relation weight has more rows
    if (F % 4 == 0 && E % 4 == 0 && ldf % 4 == 0 && ((uintptr_t)cat & 15) == 0 && ((uintptr_t)agg & 15) == 0 &&
        ((uintptr_t)scratch & 15) == 0) {
        dim3 gv((unsigned)((Mw + 63) / 64), (unsigned)((E + 63) / 64), (unsigned)((R + 1) * S));
        k_gemm_v<<<gv, 256, 0, stream>>>(Mw, E, B, wa, wb, we);
    } else {
        PCG_GEMM(false, true, true, 64, Mw, (E + 63) / 64, (R + 1) * S, Mw, E, B, wa, wb, we);
    }
    dim3 rg((unsigned)(((int64_t)Mw * E + 255) / 256), R + 1);
    k_dense_reduce<<<rg, 256, 0, stream>>>(rp);
    return pcg_check_launch("pcg_dense_bwd");
}
