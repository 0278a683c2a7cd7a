// Segmented neighbour aggregation (forward + feature-table backward).
//
// Replaces the reference's dense-mask matmul (/root/reference/src/layers.py:593-624: a [B,U] fp32
// mask built from Python index lists, copied to the GPU and multiplied with the [U,F] feature
// slice) by a gather-reduce over the selected id lists: one warp per slot of PCG_SLOT ids, feature
// rows read as float4 with several rows in flight per lane group. Items longer than one slot write
// per-slot partial sums; the last slot to finish (ticket counter) adds them up in slot order, so the
// result is deterministic.
#include "pcg_common.cuh"

template <int LPR>
__device__ __forceinline__ float4 group_reduce(float4 a) {   // sum across the 32/LPR row groups
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1) {
        a.x += __shfl_xor_sync(PCG_FULL, a.x, off);
        a.y += __shfl_xor_sync(PCG_FULL, a.y, off);
        a.z += __shfl_xor_sync(PCG_FULL, a.z, off);
        a.w += __shfl_xor_sync(PCG_FULL, a.w, off);
    }
    return a;
}

__device__ __forceinline__ float norm_scale(int n, int norm) {
    if (n <= 0) return 0.f;
    return norm == PCG_NORM_RSQRT ? 1.0f / sqrtf((float)n) : 1.0f / (float)n;
}

// LPR lanes share one feature row (float4 each); NV float4 per lane cover rows wider than LPR*4.
template <int LPR, int NV>
__global__ void __launch_bounds__(256) k_aggregate(const float* __restrict__ feat, int64_t ldf,
                                                   const int32_t* __restrict__ idx,
                                                   const int32_t* __restrict__ slot_item,
                                                   const int32_t* __restrict__ it_slot0,
                                                   const int32_t* __restrict__ it_m,
                                                   const int64_t* __restrict__ it_base,
                                                   const int32_t* __restrict__ it_extra,
                                                   const int32_t* __restrict__ it_rep, int n_items, int64_t cap_slots,
                                                   const int32_t* __restrict__ status, int norm, float* partial,
                                                   int32_t* it_done, float* __restrict__ agg) {
    constexpr int G = 32 / LPR;          // rows per warp-wide load
    constexpr int UN = (LPR == 32) ? 8 : 4;      // row loads in flight per lane group (16 in flight were measured no faster)
    const int lane = threadIdx.x & 31;
    pcg_grid_dependency_wait();          // (programmatic launch, pcg_set_pdl bit 16: resident under the selection kernels' tail)
    pcg_launch_dependents();             // the fused dense kernel behind may start streaming its weights
    int64_t n_slots = status[ST_SLOTS];
    if (n_slots > cap_slots) n_slots = cap_slots;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int V = (int)(ldf >> 2);       // float4 per row
    const int g = lane / LPR, l = lane % LPR;
    // items that own no slot (m == 0, no extra) never reach the slot loop below: zero their rows here
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_items; w += n_warps) {
        if (it_rep && it_rep[w] != w) continue;              // repeated targets have no row of their own
        if (it_m[w] > 0 || (it_extra && it_extra[w] >= 0)) continue;
        for (int c = lane; c < V; c += 32)
            *reinterpret_cast<float4*>(agg + w * ldf + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // grid-stride over the handed-out slots: the launch size does not depend on the capacity
    for (int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slots; s += n_warps) {
    const int w = slot_item[s];
    if (w < 0) continue;
    const int slot0 = it_slot0[w];
    const int m = it_m[w];
    const int c = (int)(s - slot0);
    const int len = min(PCG_SLOT, m - c * PCG_SLOT);
    const int32_t* __restrict__ ids = idx + it_base[w] + (int64_t)c * PCG_SLOT;

    // ids of this slot: two per lane, broadcast by shuffle
    int32_t id_lo = lane < len ? __ldg(ids + lane) : 0;
    int32_t id_hi = lane + 32 < len ? __ldg(ids + lane + 32) : 0;

    float4 acc[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int j0 = 0; j0 < len; j0 += G * UN) {
        float4 v[UN][NV];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int j = j0 + u * G + g;
            const int32_t id = __shfl_sync(PCG_FULL, (j & 32) ? id_hi : id_lo, j & 31);
            const float* rowp = feat + (int64_t)id * ldf;
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const int col = l + q * LPR;
                v[u][q] = (j < len && col < V) ? ld_f4(rowp + 4 * col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
            for (int q = 0; q < NV; ++q) f4_add(acc[q], v[u][q]);
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] = group_reduce<LPR>(acc[q]);

    const int nch = (m + PCG_SLOT - 1) / PCG_SLOT;
    const int extra = it_extra ? it_extra[w] : -1;
    if (nch <= 1) {
        const float sc = norm_scale(m + (extra >= 0 ? 1 : 0), norm);
        if (g == 0) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const int col = l + q * LPR;
                if (col < V) {
                    float4 a = acc[q];
                    if (extra >= 0) f4_add(a, ld_f4(feat + (int64_t)extra * ldf + 4 * col));
                    a.x *= sc; a.y *= sc; a.z *= sc; a.w *= sc;
                    *reinterpret_cast<float4*>(agg + (int64_t)w * ldf + 4 * col) = a;
                }
            }
        }
        continue;
    }
    // multi-slot item: publish the partial, last arriver reduces in slot order
    if (g == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int col = l + q * LPR;
            if (col < V) __stcg(reinterpret_cast<float4*>(partial + s * ldf + 4 * col), acc[q]);
        }
    }
    __threadfence();
    __syncwarp();
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(&it_done[w], 1);
    ticket = __shfl_sync(PCG_FULL, ticket, 0);
    if (ticket != nch - 1) continue;
    __threadfence();
    // Final sum of the item's nch partials by the whole warp: lane group g adds the partials g, g+G, g+2G, ... in that
    // order, UR loads in flight per lane (a hub row of C2 owns 100-190 slots: one lane per column adding them one
    // after the other was 10 us of dependent L2 round trips, the tail of the whole kernel), then the G group sums are
    // added by the same shuffle tree as the slots' rows. Fixed order: deterministic.
    constexpr int UR = NV == 1 ? 8 : (NV == 2 ? 4 : 2);
    float4 tot[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) tot[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c0 = g; c0 < nch; c0 += G * UR) {
        float4 v[UR][NV];
#pragma unroll
        for (int u = 0; u < UR; ++u) {
            const int cc = c0 + u * G;
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const int col = l + q * LPR;
                v[u][q] = (cc < nch && col < V) ? ld_f4_cg(partial + ((int64_t)slot0 + cc) * ldf + 4 * col)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < UR; ++u)
#pragma unroll
            for (int q = 0; q < NV; ++q) f4_add(tot[q], v[u][q]);
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) tot[q] = group_reduce<LPR>(tot[q]);
    const float sc = norm_scale(m + (extra >= 0 ? 1 : 0), norm);
    if (g == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int col = l + q * LPR;
            if (col < V) {
                float4 a = tot[q];
                if (extra >= 0) f4_add(a, ld_f4(feat + (int64_t)extra * ldf + 4 * col));
                a.x *= sc; a.y *= sc; a.z *= sc; a.w *= sc;
                *reinterpret_cast<float4*>(agg + (int64_t)w * ldf + 4 * col) = a;
            }
        }
    }
    if (lane == 0) it_done[w] = 0;          // every ticket of this item is taken: re-arm it for the next call on this selection
    }
}

// agg row of a duplicate item = agg row of its representative (see k_choose_prep).
__global__ void k_copy_dups(const int32_t* __restrict__ it_rep, int n_items, int64_t ldf, float* __restrict__ agg) {
    const int w = blockIdx.x;
    if (w >= n_items) return;
    const int rep = it_rep[w];
    if (rep == w) return;
    for (int64_t c = threadIdx.x * 4; c < ldf; c += blockDim.x * 4)
        *reinterpret_cast<float4*>(agg + (int64_t)w * ldf + c) = *reinterpret_cast<const float4*>(agg + (int64_t)rep * ldf + c);
}

// Backward w.r.t. the feature table: feat_grad[id] += d_agg[w] * scale for every id of item w.
template <int LPR, int NV>
__global__ void __launch_bounds__(256) k_aggregate_bwd(const float* __restrict__ d_agg, int64_t ldf,
                                                       const int32_t* __restrict__ idx,
                                                       const int32_t* __restrict__ slot_item,
                                                       const int32_t* __restrict__ it_slot0,
                                                       const int32_t* __restrict__ it_m,
                                                       const int64_t* __restrict__ it_base,
                                                       const int32_t* __restrict__ it_extra, int64_t cap_slots,
                                                       const int32_t* __restrict__ status, int norm,
                                                       float* feat_grad) {
    constexpr int G = 32 / LPR;
    const int lane = threadIdx.x & 31;
    int64_t n_slots = status[ST_SLOTS];
    if (n_slots > cap_slots) n_slots = cap_slots;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slots; s += n_warps) {
    const int w = slot_item[s];
    if (w < 0) continue;
    const int slot0 = it_slot0[w];
    const int m = it_m[w];
    const int c = (int)(s - slot0);
    const int len = min(PCG_SLOT, m - c * PCG_SLOT);
    const int32_t* __restrict__ ids = idx + it_base[w] + (int64_t)c * PCG_SLOT;
    const int V = (int)(ldf >> 2);
    const int g = lane / LPR, l = lane % LPR;
    const int extra = it_extra ? it_extra[w] : -1;
    const float sc = norm_scale(m + (extra >= 0 ? 1 : 0), norm);
    float4 gv[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const int col = l + q * LPR;
        gv[q] = col < V ? ld_f4(d_agg + (int64_t)w * ldf + 4 * col) : make_float4(0.f, 0.f, 0.f, 0.f);
        gv[q].x *= sc; gv[q].y *= sc; gv[q].z *= sc; gv[q].w *= sc;
    }
    int32_t id_lo = lane < len ? __ldg(ids + lane) : 0;
    int32_t id_hi = lane + 32 < len ? __ldg(ids + lane + 32) : 0;
    const int total = len + ((c == 0 && extra >= 0) ? 1 : 0);   // slot 0 also carries the extra (self) row
    for (int j0 = 0; j0 < total; j0 += G) {
        const int j = j0 + g;
        int32_t id = __shfl_sync(PCG_FULL, (j & 32) ? id_hi : id_lo, j & 31);
        if (j >= len) id = extra;
        if (j < total) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const int col = l + q * LPR;
                if (col < V) {
                    float* dst = feat_grad + (int64_t)id * ldf + 4 * col;
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(gv[q].x), "f"(gv[q].y),
                                 "f"(gv[q].z), "f"(gv[q].w)
                                 : "memory");
                }
            }
        }
    }
    }
}

template <int LPR, int NV>
static void launch_agg(bool bwd, const float* a0, int64_t ldf, const int32_t* idx, const int32_t* slot_item,
                       const int32_t* it_slot0, const int32_t* it_m, const int64_t* it_base, const int32_t* it_extra,
                       const int32_t* it_rep, int n_items, int64_t cap_slots, const int32_t* status, int norm, float* partial, int32_t* it_done, float* out,
                       cudaStream_t stream) {
    int64_t blocks64 = (cap_slots * 32 + 255) / 256;
    const int64_t max_blocks = (int64_t)pcg_device_sms() * 8;      // 8 resident CTAs of 8 warps per SM
    const int blocks = (int)(blocks64 < max_blocks ? blocks64 : max_blocks);
    if (!bwd)
        pcg_launch(k_aggregate<LPR, NV>, dim3(blocks), dim3(256), 0, stream, (pcg_pdl_enabled() & 16) != 0, a0, ldf, idx,
                   slot_item, it_slot0, it_m, it_base, it_extra, it_rep, n_items, cap_slots, status, norm, partial, it_done,
                   out);
    else
        k_aggregate_bwd<LPR, NV><<<blocks, 256, 0, stream>>>(a0, ldf, idx, slot_item, it_slot0, it_m, it_base, it_extra,
                                                             cap_slots, status, norm, out);
}

static int dispatch_agg(bool bwd, const float* a0, int64_t ldf, const int32_t* idx, const int32_t* slot_item,
                        const int32_t* it_slot0, const int32_t* it_m, const int64_t* it_base, const int32_t* it_extra,
                        const int32_t* it_rep, int n_items, int64_t cap_slots, const int32_t* status, int norm, float* partial, int32_t* it_done, float* out,
                        cudaStream_t stream) {
    const int64_t V = ldf / 4;
#define PCG_AGG(LPR, NV)                                                                                          \
    launch_agg<LPR, NV>(bwd, a0, ldf, idx, slot_item, it_slot0, it_m, it_base, it_extra, it_rep, n_items, cap_slots, \
                        status, norm, partial, it_done, out, stream)
    if (V <= 8) PCG_AGG(8, 1);
    else if (V <= 16) PCG_AGG(16, 1);
    else if (V <= 32) PCG_AGG(32, 1);
    else if (V <= 64) PCG_AGG(32, 2);
    else if (V <= 128) PCG_AGG(32, 4);
    else {
        pcg_set_error("pcg_aggregate: feature rows wider than 512 floats are not supported (ldf=%lld)", (long long)ldf);
        return (int)cudaErrorInvalidValue;
    }
#undef PCG_AGG
    return 0;
}

extern "C" int pcg_aggregate(const float* feat, int64_t ldf, const int32_t* idx, const int32_t* slot_item,
                             const int32_t* it_slot0, const int32_t* it_m, const int64_t* it_base,
                             const int32_t* it_extra, const int32_t* it_rep, int copy_dups, int n_items,
                             int64_t cap_slots, const int32_t* status, int norm, float* partial, int32_t* it_done,
                             float* agg, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_items == 0) return 0;
    PCG_REQUIRE(ldf > 0 && ldf % 4 == 0, "pcg_aggregate: ldf=%lld must be a positive multiple of 4", (long long)ldf);
    PCG_REQUIRE(feat && idx && slot_item && it_slot0 && it_m && it_base && status && partial && it_done && agg,
                "pcg_aggregate: null pointer");
    PCG_REQUIRE(((uintptr_t)feat & 15) == 0 && ((uintptr_t)agg & 15) == 0 && ((uintptr_t)partial & 15) == 0,
                "pcg_aggregate: feat/agg/partial must be 16-byte aligned");
    if (n_items == 0 || cap_slots == 0) return 0;
    int rc = dispatch_agg(false, feat, ldf, idx, slot_item, it_slot0, it_m, it_base, it_extra, it_rep, n_items, cap_slots,
                          status, norm, partial, it_done, agg, stream);
    if (rc) return rc;
    if (it_rep && copy_dups) k_copy_dups<<<n_items, 32, 0, stream>>>(it_rep, n_items, ldf, agg);
    return pcg_check_launch("pcg_aggregate");
}

extern "C" int pcg_aggregate_bwd(const float* d_agg, int64_t ldf, const int32_t* idx, const int32_t* slot_item,
                                 const int32_t* it_slot0, const int32_t* it_m, const int64_t* it_base,
                                 const int32_t* it_extra, int n_items, int64_t cap_slots, const int32_t* status,
                                 int norm, float* feat_grad, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_items == 0) return 0;
    PCG_REQUIRE(ldf > 0 && ldf % 4 == 0, "pcg_aggregate_bwd: ldf=%lld must be a positive multiple of 4", (long long)ldf);
    PCG_REQUIRE(d_agg && idx && slot_item && it_slot0 && it_m && it_base && status && feat_grad,
                "pcg_aggregate_bwd: null pointer");
    if (n_items == 0 || cap_slots == 0) return 0;
    int rc = dispatch_agg(true, d_agg, ldf, idx, slot_item, it_slot0, it_m, it_base, it_extra, nullptr, n_items, cap_slots,
                          status, norm, nullptr, nullptr, feat_grad, stream);
    if (rc) return rc;
    return pcg_check_launch("pcg_aggregate_bwd");
}
