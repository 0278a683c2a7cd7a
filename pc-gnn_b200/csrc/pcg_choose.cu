// Choose step on the GPU: label-aware top-k neighbour filtering + minority oversampling.
//
// Replaces the per-target Python loop of /root/reference/src/layers.py:633-738 (one torch.sort and
// two .tolist() host syncs per target per relation) by an exact k-smallest selection on the key
// (|s_v - s_u| as fp32 bits, position in the id-sorted row):
//   1. the k-th smallest distance T (and how many elements equal to T are still needed) is found by a
//      radix selection over keys held in REGISTERS: 8 key bits per step counted into per-warp shared-memory
//      histograms (the skewed exponent digit is peeled with ballots first), at most 4 steps, usually 2,
//   2. an ORDERED compaction writes every element with d < T plus the first `need` elements with
//      d == T (row order == id order, which is the reference's stable-sort tie rule),
//   3. for positive targets the o nearest train positives come from the score-sorted pool
//      (pcg_sort_pool): around lower_bound(s_v) the distances grow monotonically to both sides, so
//      the o-th smallest distance is a k-th-of-two-sorted-sequences search done with warp-wide
//      32-ary probes; pool members that are already kept (a per-item bitmap over pool positions,
//      filled during the compaction) are dropped: the set union of src/layers.py:690-694.
//
// Work decomposition (every thread holds at most a few row entries, so an item's latency is a handful of
// dependent memory round trips plus 2-4 selection steps, whatever its length):
//   prep     up to #SMs cooperative CTAs: folds repeated targets (pick_step samples with replacement), computes
//            every item's sizes, hands out the output slots by a prefix sum (deterministic layout, no atomics
//            on the item path) and sorts the items into four tier queues,
//   warp     d <= 256          one warp per item,           } one kernel (k_choose_small) drains both queues,
//   cta      256 < d <= 2048   one 256-thread CTA per item, } half of its CTAs starting on each
//   wide     2048 < d <= 16384 one thread-block CLUSTER of 8 x 256 threads per item (launched first: the longest
//            rows are the critical path); the CTAs' histograms and compaction totals are exchanged through
//            distributed shared memory, one cluster barrier per selection step,
//   big      d > 16384         one 1024-thread CTA, distances in shared memory / recomputed per pass (rare hubs).
// The tier kernels run side by side on forked streams (CUDA-graph capturable).
#include "pcg_common.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;




#define PCG_SMALL_MAX 256      // warp tier: <= 8 entries per lane
#define PCG_WARPS_PER_CTA 8    // warp kernel: 8 items in flight per CTA
#define PCG_GRP_NT 256         // cta tier: threads per CTA
#define PCG_CTA_MAX 2048       // cta tier: <= 8 entries per thread
#define PCG_LARGE_NT 1024      // wide and big tiers: threads per CTA (one CTA per SM)
#define PCG_CL 8               // wide tier: CTAs per cluster
#define PCG_CL_MAX 16384       // wide tier: 8 x 256 threads x <= 8 entries
#define PCG_HUGE_MAX 131072     // huge tier: 8 x 1024 threads x <= 16 entries (ids / pool positions stashed in shared memory)
#define PCG_HUGE16_MAX 262144   // ... on clusters of 16 CTAs (non-portable cluster size), where the device can place them
#define PCG_KB_WORDS 256       // kept-pool bitmap words per item (pools up to 8192 positives; else row-position bits)
#define PCG_CAND_CAP 2048      // big tier: candidate keys kept in shared memory once they fit
#define PCG_PREP_NT 1024

struct ChooseP {
    const int64_t* indptr;
    const int32_t* indices;
    const float* score;
    const float* entry_score;
    const float* center_score;
    const int32_t* targets;
    const int64_t* labels;
    const int32_t* k_override;
    const float* ps_score;      // pool scores ascending (ties by pool position)
    const int32_t* ps_pos;      // pool position of each sorted entry
    const int32_t* ps_id;       // node id of each sorted entry
    const int32_t* entry_pool_pos; // [nnz] pool position of every CSR entry's node, or -1 (NULL: binary-search fallback)
    int64_t n_nodes;            // rows per relation of THIS CSR (a row partition holds rows [row_lo, row_lo + n_nodes))
    int64_t row_lo;             // global id of local row 0 (0 for an unpartitioned graph)
    int R, B, P, train;
    double thresh[PCG_MAX_REL];
    double rho;
    int32_t* sel_idx;
    float* sel_dist;
    int64_t cap_slots;
    int32_t* slot_item;
    int32_t* it_slot0;
    int32_t* it_m;
    int64_t* it_base;
    int32_t* it_done;
    int32_t* it_rep;            // representative item of every item (itself unless an earlier target has the same id)
    int32_t* status;
    int32_t* first;             // [n_nodes] smallest batch index per node id (0x7f7f7f7f when absent); NULL: no folding
    int32_t* q_warp;
    int32_t* q_cta;
    int32_t* q_cl;
    int32_t* q_huge;
    int32_t* q_huge16;
    int huge16_ok;              // clusters of 16 x 1024 threads can be scheduled on this device
    int32_t* q_big;
    uint32_t* bits_slab;        // [grid_big, slab_words] kept-position bitmasks of the big tier
    int64_t slab_words;
    int32_t* sticky;            // workspace word that keeps every call's overflow / foreign-target flag until the host clears it
};

#ifdef PCG_TRACE
// Debug build only (make EXTRA=-DPCG_TRACE): per-item phase timestamps, PCG_TRACE_SLOTS int64 per item (slots 0-7 phases,
// 8-10 d / k / o, 11 selection steps, 12-27 selection sub-phases of the cluster tiers, 28-33 compaction / oversampling).
#define PCG_TRACE_SLOTS 40
__device__ long long* g_trace = nullptr;
__device__ __forceinline__ long long trace_now() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TRACE(slot) do { if (g_trace && tid == 0) g_trace[(int64_t)w * PCG_TRACE_SLOTS + (slot)] = trace_now(); } while (0)
#define TRACE_VAL(slot, val) do { if (g_trace && tid == 0) g_trace[(int64_t)w * PCG_TRACE_SLOTS + (slot)] = (val); } while (0)
#define PTRACE(slot) do { if (g_trace && threadIdx.x == 0) g_trace[(int64_t)p.R * p.B * PCG_TRACE_SLOTS + (slot)] = trace_now(); } while (0)
// cluster tiers, CTA 0 of the cluster only
#define TRACE0(slot) do { if (g_trace && tid == 0 && cg::this_cluster().block_rank() == 0) g_trace[(int64_t)w * PCG_TRACE_SLOTS + (slot)] = trace_now(); } while (0)
#define TRACE0_VAL(slot, val) do { if (g_trace && tid == 0 && cg::this_cluster().block_rank() == 0) g_trace[(int64_t)w * PCG_TRACE_SLOTS + (slot)] = (val); } while (0)
#else
#define PTRACE(slot) do {} while (0)
#define TRACE(slot) do {} while (0)
#define TRACE_VAL(slot, val) do {} while (0)
#define TRACE0(slot) do {} while (0)
#define TRACE0_VAL(slot, val) do {} while (0)
#endif
template <int NT>
__device__ __forceinline__ void grp_sync() {
    if (NT == 32) __syncwarp(); else __syncthreads();
}

// Exclusive prefix counts of two flags over the NT threads of the group (tile order == thread
// order), plus the totals. xw: >= NT/32 ints of group-private shared memory.
template <int NT>
__device__ __forceinline__ void grp_excl2(bool fa, bool fb, int tid, int* xw, int& ea, int& eb, int& ta, int& tb) {
    const unsigned lt = lanemask_lt();
    unsigned ma = __ballot_sync(PCG_FULL, fa), mb = __ballot_sync(PCG_FULL, fb);
    ea = __popc(ma & lt);
    eb = __popc(mb & lt);
    if (NT == 32) {
        ta = __popc(ma);
        tb = __popc(mb);
    } else {
        constexpr int NW = NT / 32;
        const int wid = tid >> 5, lane = tid & 31;
        if (lane == 0) xw[wid] = __popc(ma) | (__popc(mb) << 16);
        __syncthreads();
        // every warp scans the NW packed counts with shuffles (NW <= 32)
        int pk = lane < NW ? xw[lane] : 0;
        int incl = pk;
#pragma unroll
        for (int off = 1; off < NW; off <<= 1) {
            int t = __shfl_up_sync(PCG_FULL, incl, off);
            if (lane >= off) incl += t;
        }
        const int tot = __shfl_sync(PCG_FULL, incl, NW - 1);
        const int mine = __shfl_sync(PCG_FULL, incl - pk, wid);
        ea += mine & 0xffff;
        eb += mine >> 16;
        ta = tot & 0xffff;
        tb = tot >> 16;
        __syncthreads();
    }
}

// Group-wide min and max of get(0..n-1). xw[24..27] used as scratch for NT > 32.
template <int NT, class Get>
__device__ __forceinline__ void grp_minmax(Get get, int n, int tid, int* xw, uint32_t& kmin, uint32_t& kmax) {
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (int j = tid; j < n; j += NT) { const uint32_t x = get(j); lo = min(lo, x); hi = max(hi, x); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(PCG_FULL, lo, off));
        hi = max(hi, __shfl_xor_sync(PCG_FULL, hi, off));
    }
    if (NT > 32) {
        if (tid == 0) { xw[24] = (int)0xffffffffu; xw[25] = 0; }
        __syncthreads();
        if ((tid & 31) == 0) { atomicMin((uint32_t*)&xw[24], lo); atomicMax((uint32_t*)&xw[25], hi); }
        __syncthreads();
        lo = (uint32_t)xw[24];
        hi = (uint32_t)xw[25];
        __syncthreads();
    }
    kmin = lo;
    kmax = hi;
}
// k-th smallest (kth is 1-based) of get(0..n-1) as (T, need): T = that value, need = how many of
// the elements equal to T belong to the kth smallest (in position order).
// Digits are cut from the highest bit in which min and max differ (bits every key shares need no
// pass, and the first digit then spans exponent AND mantissa bits, which spreads the histogram).
template <int NT, class Get>
__device__ __forceinline__ void radix_select(Get get, int n, int kth, uint32_t* hist, int* xw, int tid, uint32_t& T,
                                             int& need) {
    uint32_t kmin, kmax;
    grp_minmax<NT>(get, n, tid, xw, kmin, kmax);
    if (kmin == kmax) { T = kmin; need = kth; return; }
    int hb = 31 - __clz(kmin ^ kmax);
    uint32_t mask = hb == 31 ? 0u : ~((2u << hb) - 1u);
    uint32_t prefix = kmin & mask;
    int remaining = kth;
    bool first = true;
    while (hb >= 0) {
        const int shift = max(hb - 7, 0);
        const uint32_t dmask = (1u << (hb - shift + 1)) - 1u;
        for (int b = tid; b < 256; b += NT) hist[b] = 0;
        grp_sync<NT>();
        const int n_up = (n + 31) & ~31;
        for (int j = tid; j < n_up; j += NT) {
            const bool valid = j < n;
            const uint32_t key = valid ? get(j) : 0u;
            const bool active = valid && (key & mask) == prefix;
            const uint32_t digit = (key >> shift) & dmask;
            if (first) {
                // the leading digit is the skewed one: peel the two most common values of each warp
                // with ballots so that one lane adds the whole count instead of 32 colliding atomics
                unsigned rem = __ballot_sync(PCG_FULL, active);
#pragma unroll 1
                for (int it = 0; it < 2 && rem; ++it) {
                    const int ldr = __ffs(rem) - 1;
                    const uint32_t d0 = __shfl_sync(PCG_FULL, digit, ldr);
                    const unsigned same = __ballot_sync(PCG_FULL, active && digit == d0);
                    if ((tid & 31) == ldr) atomicAdd(&hist[d0], (uint32_t)__popc(same));
                    rem &= ~same;
                }
                if ((rem >> (tid & 31)) & 1u) atomicAdd(&hist[digit], 1u);
            } else if (active) {
                atomicAdd(&hist[digit], 1u);
            }
        }
        grp_sync<NT>();
        if (tid < 32) {
            uint32_t c[8];
            int s = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) { c[b] = hist[tid * 8 + b]; s += (int)c[b]; }
            int incl = s;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                int t = __shfl_up_sync(PCG_FULL, incl, off);
                if (tid >= off) incl += t;
            }
            int excl = incl - s;
            if (excl < remaining && remaining <= incl) {
                int run = excl;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    if (run + (int)c[b] >= remaining) { xw[30] = tid * 8 + b; xw[31] = remaining - run; break; }
                    run += (int)c[b];
                }
            }
        }
        grp_sync<NT>();
        const uint32_t digit = (uint32_t)xw[30];
        remaining = xw[31];
        prefix |= digit << shift;
        mask |= dmask << shift;
        hb = shift - 1;
        first = false;
    }
    T = prefix;
    need = remaining;
}

// ---- bit-serial selection (two bits per step, counts by warp reduction): used by the big tier, whose keys
// live in shared memory or are recomputed per pass ----
// Decide the next digit of the k-th smallest key from the counts of the three lowest digit values.
__device__ __forceinline__ uint32_t pick_digit(int c0, int c1, int c2, int& remaining) {
    if (remaining <= c0) return 0u;
    remaining -= c0;
    if (remaining <= c1) return 1u;
    remaining -= c1;
    if (remaining <= c2) return 2u;
    remaining -= c2;
    return 3u;
}

// One CTA, keys behind get(0..n-1) (shared memory). Same two-bit steps; as soon as the keys still matching
// the prefix fit into `cand` they are copied there and the remaining steps only look at that list.
//   wsum: >= 3 * NT/32 ints of scratch, xw[24..28] scratch
template <int NT, class Get>
__device__ __forceinline__ void cta_bitselect(Get get, int n, int kth, uint32_t* cand, int cand_cap, int* wsum, int* xw,
                                              int tid, uint32_t& T, int& need) {
    constexpr int NW = NT / 32;
    const int wid = tid >> 5, lane = tid & 31;
    uint32_t kmin, kmax;
    grp_minmax<NT>(get, n, tid, xw, kmin, kmax);
    if (kmin == kmax) { T = kmin; need = kth; return; }
    int hb = 31 - __clz(kmin ^ kmax);
    uint32_t mask = hb == 31 ? 0u : ~((2u << hb) - 1u);
    uint32_t prefix = kmin & mask;
    int remaining = kth;
    bool listed = false;
    int n_list = 0;                // length of the candidate list once it exists
    int n_act = n;                 // keys matching the prefix
    while (hb >= 0) {
        const int shift = max(hb - 1, 0);
        const uint32_t dmask = hb >= 1 ? 3u : 1u;
        int c0 = 0, c1 = 0, c2 = 0;
        const int lim = listed ? n_list : n;
        for (int j = tid; j < lim; j += NT) {
            const uint32_t key = listed ? cand[j] : get(j);
            const bool active = ((key ^ prefix) & mask) == 0u;
            const uint32_t dg = (key >> shift) & dmask;
            c0 += active && dg == 0u;
            c1 += active && dg == 1u;
            c2 += active && dg == 2u;
        }
        c0 = __reduce_add_sync(PCG_FULL, c0);
        c1 = __reduce_add_sync(PCG_FULL, c1);
        c2 = __reduce_add_sync(PCG_FULL, c2);
        if (lane == 0) { wsum[wid * 3] = c0; wsum[wid * 3 + 1] = c1; wsum[wid * 3 + 2] = c2; }
        __syncthreads();
        c0 = lane < NW ? wsum[lane * 3] : 0;
        c1 = lane < NW ? wsum[lane * 3 + 1] : 0;
        c2 = lane < NW ? wsum[lane * 3 + 2] : 0;
        c0 = __reduce_add_sync(PCG_FULL, c0);
        c1 = __reduce_add_sync(PCG_FULL, c1);
        c2 = __reduce_add_sync(PCG_FULL, c2);
        const uint32_t dg = pick_digit(c0, c1, c2, remaining);
        const int cnt_dg = dg == 0u ? c0 : (dg == 1u ? c1 : (dg == 2u ? c2 : n_act - c0 - c1 - c2));
        prefix |= dg << shift;
        mask |= dmask << shift;
        hb = shift - 1;
        n_act = cnt_dg;
        __syncthreads();                       // wsum may be rewritten next step
        if (!listed && hb >= 0 && n_act <= cand_cap) {
            // copy the surviving keys into the candidate list (order is irrelevant for counting)
            if (tid == 0) xw[26] = 0;
            __syncthreads();
            const int n_up = (n + 31) & ~31;
            for (int j = tid; j < n_up; j += NT) {
                const uint32_t key = j < n ? get(j) : 0u;
                const bool active = j < n && ((key ^ prefix) & mask) == 0u;
                const unsigned m = __ballot_sync(PCG_FULL, active);
                int base = 0;
                if (lane == 0 && m) base = atomicAdd(&xw[26], __popc(m));
                base = __shfl_sync(PCG_FULL, base, 0);
                if (active) cand[base + __popc(m & lanemask_lt())] = key;
            }
            __syncthreads();
            listed = true;
            n_list = n_act;
        }
    }
    T = prefix;
    need = remaining;
}
// ---- histogram selection: 8 key bits per step, at most 4 steps (bit 31 of a distance is always 0) ----
// k-th smallest (kth is 1-based) of the keys held in registers by a group of NT threads (NT == 32: one warp,
// else a whole CTA), as (T, need). Every step counts the digit [shift, shift+8) of the keys that still match
// the prefix into shared-memory histograms (one per warp for NT <= 256, one per 4 warps above: no two lanes
// of a warp collide unless their digits are equal), the histograms are merged bin-per-thread, scanned, and the
// bin holding the k-th key extends the prefix. The first digit is the fp32 exponent, where most keys share two
// or three values: those are peeled off with ballots so one lane adds the whole count. The loop ends as soon
// as a single key matches the prefix (typically after two steps).
//   hist: NH*256 words, all ZERO on entry and on exit (NH = 1 for a warp, 8 for a CTA);  xw: >= 24 ints
// CL > 1: the group is a thread-block cluster of CL CTAs; every CTA publishes its merged histogram in its own
// shared memory (chist, parity double buffer `par`), one cluster barrier, then every thread adds up its bin
// over the CL CTAs through distributed shared memory: all CTAs scan the same totals and agree on the digit.
template <int NT, int NE, int CL = 1>
__device__ __forceinline__ void hist_select(const uint32_t (&key)[NE], uint32_t vmask, int kth, uint32_t* hist, int* xw,
                                            int tid, uint32_t& T, int& need, uint32_t* chist = nullptr,
                                            int* par = nullptr, int w = 0) {

    constexpr int NH = NT == 32 ? 1 : 8;
    const int lane = tid & 31, wid = tid >> 5;
    uint32_t* myh = hist + (NT == 32 ? 0 : (wid & (NH - 1)) * 256);
    uint32_t prefix = 0u, mask = 0u;
    int remaining = kth;
#pragma unroll 1
    for (int step = 0; step < 4; ++step) {
        const int shift = step == 3 ? 0 : 23 - 8 * step;
        const uint32_t dmask = step == 3 ? 0x7fu : 0xffu;
        if (step == 0) {
            // exponent digit: most keys of a warp share two or three values. The three most frequent digits of
            // the warp's first row are counted in registers over all NE rows and added by one lane each.
            uint32_t pd[3];
            {
                const bool a0 = vmask & 1u;
                const uint32_t dg0 = key[0] >> 23;
                unsigned rem = __ballot_sync(PCG_FULL, a0);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int ldr = rem ? __ffs(rem) - 1 : 0;
                    pd[q] = rem ? __shfl_sync(PCG_FULL, dg0, ldr) : 0xffffffffu;     // 0xffffffff never matches
                    rem &= ~__ballot_sync(PCG_FULL, a0 && dg0 == pd[q]);
                }
            }
            int pc0 = 0, pc1 = 0, pc2 = 0;
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                const bool active = (vmask >> e) & 1u;
                const uint32_t digit = key[e] >> 23;
                const bool m0 = active && digit == pd[0], m1 = active && digit == pd[1], m2 = active && digit == pd[2];
                pc0 += m0; pc1 += m1; pc2 += m2;
                if (active && !(m0 || m1 || m2)) atomicAdd(&myh[digit], 1u);
            }
            pc0 = __reduce_add_sync(PCG_FULL, pc0);
            pc1 = __reduce_add_sync(PCG_FULL, pc1);
            pc2 = __reduce_add_sync(PCG_FULL, pc2);
            if (lane == 0 && pc0) atomicAdd(&myh[pd[0]], (uint32_t)pc0);
            if (lane == 1 && pc1) atomicAdd(&myh[pd[1]], (uint32_t)pc1);
            if (lane == 2 && pc2) atomicAdd(&myh[pd[2]], (uint32_t)pc2);
        } else {
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                const bool active = ((vmask >> e) & 1u) && ((key[e] ^ prefix) & mask) == 0u;
                if (active) atomicAdd(&myh[(key[e] >> shift) & dmask], 1u);
            }
        }
        grp_sync<NT>();
        if (CL > 1) { TRACE0(12 + 4 * step); TRACE0_VAL(11, step + 1); }
        // merge (and clear) the histograms, scan, find the bin of the k-th key
        int dg = -1, rem_in = 0, cnt_in = 0;
        if (NT == 32) {
            uint32_t c[8];
            int s = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) { c[b] = hist[lane * 8 + b]; hist[lane * 8 + b] = 0u; s += (int)c[b]; }
            int incl = s;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(PCG_FULL, incl, off);
                if (lane >= off) incl += t;
            }
            int run = incl - s;
            if (run < remaining && remaining <= incl) {
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    if (dg < 0 && run + (int)c[b] >= remaining) { dg = lane * 8 + b; rem_in = remaining - run; cnt_in = (int)c[b]; }
                    run += (int)c[b];
                }
            }
            const unsigned who = __ballot_sync(PCG_FULL, dg >= 0);
            const int src = __ffs(who) - 1;
            dg = __shfl_sync(PCG_FULL, dg, src);
            rem_in = __shfl_sync(PCG_FULL, rem_in, src);
            cnt_in = __shfl_sync(PCG_FULL, cnt_in, src);
        } else {
            int c = 0, incl = 0;
            if (tid < 256) {
#pragma unroll
                for (int h = 0; h < NH; ++h) { c += (int)hist[h * 256 + tid]; hist[h * 256 + tid] = 0u; }
            }
            if (CL > 1) {
                cg::cluster_group cl = cg::this_cluster();
                uint32_t* mine_h = chist + *par * 256 + (tid & 255);
                if (tid < 256) *mine_h = (uint32_t)c;
                cl.sync();
                TRACE0(13 + 4 * step);
                if (tid < 256) {
                    c = 0;
#pragma unroll
                    for (int r = 0; r < CL; ++r) c += (int)*cl.map_shared_rank(mine_h, r);
                }
                *par ^= 1;
                TRACE0(14 + 4 * step);
            }
            if (tid < 256) {
                incl = c;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int t = __shfl_up_sync(PCG_FULL, incl, off);
                    if (lane >= off) incl += t;
                }
                if (lane == 31) xw[wid] = incl;
            }
            __syncthreads();
            if (tid < 256) {
                int base = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) base += q < wid ? xw[q] : 0;
                const int excl = base + incl - c;
                if (excl < remaining && remaining <= excl + c) { xw[16] = tid; xw[17] = remaining - excl; xw[18] = c; }
            }
            __syncthreads();
            dg = xw[16]; rem_in = xw[17]; cnt_in = xw[18];
            if (CL > 1) TRACE0(15 + 4 * step);
        }
        remaining = rem_in;
        prefix |= (uint32_t)dg << shift;
        mask |= dmask << shift;
        if (cnt_in == 1 && step < 3) {
            // a single key matches the prefix: it is the k-th
            uint32_t mine = 0u;
            bool have = false;
#pragma unroll
            for (int e = 0; e < NE; ++e)
                if (((vmask >> e) & 1u) && ((key[e] ^ prefix) & mask) == 0u) { mine = key[e]; have = true; }
            if (NT == 32) {
                const unsigned who = __ballot_sync(PCG_FULL, have);
                prefix = __shfl_sync(PCG_FULL, mine, __ffs(who) - 1);
            } else if (CL == 1) {
                if (have) xw[19] = (int)mine;
                __syncthreads();
                prefix = (uint32_t)xw[19];
            } else {
                cg::cluster_group cl = cg::this_cluster();
                if (have)
                    for (int r = 0; r < CL; ++r) *cl.map_shared_rank(&xw[19], r) = (int)mine;
                cl.sync();
                prefix = (uint32_t)xw[19];
            }
            break;
        }
    }
    T = prefix;
    need = remaining;
}

// First index in [lo, hi) where pred turns false (pred is true on a prefix), found with warp-wide
// 32-ary probes: ceil(log32(range)) rounds of one predicate evaluation per lane. Every warp of the
// group runs it redundantly (same addresses -> broadcast loads), so no block barrier is needed.
template <class Pred>
__device__ __forceinline__ int warp_partition_point(int lo, int hi, Pred pred) {
    const int lane = threadIdx.x & 31;
    while (hi > lo) {
        const int span = hi - lo;
        const int step = (span + 31) >> 5;
        const int idx = lo + lane * step;
        const bool t = idx < hi ? pred(idx) : false;
        const int cnt = __popc(__ballot_sync(PCG_FULL, t));
        if (cnt == 0) return lo;
        const int nlo = lo + (cnt - 1) * step + 1;
        hi = min(hi, lo + cnt * step);
        lo = nlo;
    }
    return lo;
}
// Per-item header shared by all tiers.
struct Item {
    int w, r, i, d, k, o, nslots, slot0;
    int32_t v;
    int64_t beg, off;
    float sv;
    bool use_kb, want_bits;
};

__device__ __forceinline__ void item_header(const ChooseP& p, int w, Item& it) {
    it.w = w;
    it.r = w / p.B;
    it.i = w - it.r * p.B;
    it.v = p.targets[it.i];
    const int64_t row = (int64_t)it.r * p.n_nodes + (it.v - p.row_lo);
    it.beg = p.indptr[row];
    it.d = (int)(p.indptr[row + 1] - it.beg);
    it.sv = p.center_score ? p.center_score[it.i] : p.score[it.v];
    const bool positive = p.train && p.labels && p.labels[it.i] == 1;
    item_counts(it.d, p.thresh[it.r], p.rho, positive, p.P, p.k_override ? p.k_override[w] : 0, p.k_override != nullptr,
                it.k, it.o);
    it.nslots = (it.k + it.o + PCG_SLOT - 1) / PCG_SLOT;
    it.slot0 = p.it_slot0[w];                       // handed out by k_choose_prep
    it.off = (int64_t)it.slot0 * PCG_SLOT;
    it.use_kb = it.o > 0 && p.entry_pool_pos != nullptr && p.P <= PCG_KB_WORDS * 32;
    it.want_bits = it.o > 0 && !it.use_kb;
}

// Minority oversampling for one positive item: the o nearest train positives (score-sorted pool)
// that are not already kept, appended behind the k kept neighbours. Returns how many were added.
//   kbits   bitmap over pool positions of kept neighbours (use_kb), else bits = kept row positions
template <int NT>
__device__ __forceinline__ int oversample(const ChooseP& p, const Item& it, int tid, const int32_t* __restrict__ nbr,
                                          const uint32_t* kbits, const uint32_t* bits, uint32_t* hist, int* xw) {
    const float* __restrict__ S = p.ps_score;
    const int32_t* __restrict__ SP = p.ps_pos;
    const int32_t* __restrict__ SI = p.ps_id;
    const int P = p.P, o = it.o, k = it.k, d = it.d;
    const float sv = it.sv;
    const int w = it.w;
    // split point: first sorted entry with score >= sv; A walks left from it, B right. The distances
    // grow monotonically along both, so the o-th smallest is a k-th-of-two-sorted-sequences search.
    // (first warp of the group searches, the result is broadcast through `hist`)
    int c = 0, a_less = 0, a_le = 0, b_less = 0, b_le = 0;
    uint32_t Tp = 0;
    if (NT == 32 || tid < 32) {
        c = warp_partition_point(0, P, [&](int q) { return __ldg(S + q) < sv; });
        const int nA0 = c, nB0 = P - c;
        auto A0 = [&](int q) -> uint32_t { return dist_bits(sv, __ldg(S + (c - 1 - q))); };
        auto B0 = [&](int q) -> uint32_t { return dist_bits(sv, __ldg(S + (c + q))); };
        const int ia = warp_partition_point(max(0, o - nB0), min(o, nA0), [&](int m) { return B0(o - m - 1) > A0(m); });
        const int ib = o - ia;
        if (ia > 0) Tp = A0(ia - 1);
        if (ib > 0) Tp = max(Tp, B0(ib - 1));
        a_less = warp_partition_point(0, nA0, [&](int q) { return A0(q) < Tp; });
        a_le = warp_partition_point(a_less, nA0, [&](int q) { return A0(q) <= Tp; });
        b_less = warp_partition_point(0, nB0, [&](int q) { return B0(q) < Tp; });
        b_le = warp_partition_point(b_less, nB0, [&](int q) { return B0(q) <= Tp; });
    }
    if (NT > 32) {
        if (tid == 0) {
            hist[64] = (uint32_t)c; hist[65] = (uint32_t)a_less; hist[66] = (uint32_t)a_le;
            hist[67] = (uint32_t)b_less; hist[68] = (uint32_t)b_le; hist[69] = Tp;
        }
        __syncthreads();
        c = (int)hist[64]; a_less = (int)hist[65]; a_le = (int)hist[66];
        b_less = (int)hist[67]; b_le = (int)hist[68]; Tp = hist[69];
        __syncthreads();
    }
    if (NT > 32) { TRACE(31); }
    const int cnt_less = a_less + b_less;
    const int tie_a = a_le - a_less, ties = tie_a + (b_le - b_less);
    const int needp = o - cnt_less;            // 1 <= needp <= ties
    auto tie_index = [&](int t) -> int { return t < tie_a ? c - 1 - (a_less + t) : c + b_less + (t - tie_a); };
    uint32_t Tpos = 0xffffffffu;
    if (needp < ties) {   // more equal-distance candidates than needed: smallest pool positions win
        auto getpos = [&](int t) -> uint32_t { return (uint32_t)__ldg(SP + tie_index(t)); };
        int unused;
        radix_select<NT>(getpos, ties, needp, hist, xw, tid, Tpos, unused);
    }
    TRACE(5);
    const int total = cnt_less + ties;
    int run_sel = 0, n_emit = 0;
    for (int base = 0; base < total; base += NT) {
        const int e = base + tid;
        const bool valid = e < total;
        int idx = 0, pos = 0;
        if (valid) {
            idx = e < a_less ? c - 1 - e : (e < cnt_less ? c + (e - a_less) : tie_index(e - cnt_less));
            pos = __ldg(SP + idx);
        }
        const bool selp = valid && (e < cnt_less || (uint32_t)pos <= Tpos);
        bool emit = false;
        int32_t id = 0;
        if (selp) {
            id = __ldg(SI + idx);
            bool dup;
            if (it.use_kb) {
                dup = (kbits[pos >> 5] >> (pos & 31)) & 1u;
            } else {
                int l = 0, h = d;            // lower_bound of id in the id-sorted row
                while (l < h) {
                    const int mid = (l + h) >> 1;
                    if (nbr[mid] < id) l = mid + 1; else h = mid;
                }
                dup = l < d && nbr[l] == id && ((k == d) || ((bits[l >> 5] >> (l & 31)) & 1u));
            }
            emit = !dup;
        }
        int es, ee, ts, te;
        grp_excl2<NT>(selp, emit, tid, xw, es, ee, ts, te);
        if (selp && p.sel_dist) p.sel_dist[it.off + k + run_sel + es] = __uint_as_float(dist_bits(sv, __ldg(S + idx)));
        if (emit) p.sel_idx[it.off + k + n_emit + ee] = id;
        run_sel += ts;
        n_emit += te;
    }
    return n_emit;
}

template <int NT>
__device__ __forceinline__ void item_finish(const ChooseP& p, const Item& it, int tid, int m) {
    if (tid == 0) {
        p.it_slot0[it.w] = it.slot0;
        p.it_m[it.w] = m;
        p.it_base[it.w] = it.off;
        p.it_done[it.w] = 0;
    }
    for (int c = tid; c < it.nslots; c += NT) p.slot_item[it.slot0 + c] = (c * PCG_SLOT < m) ? it.w : -1;
}
// --------------------------------------------------------------------------------- warp tier
struct WarpSmem {
    uint32_t shist[256];                      // selection histogram (zero between uses)
    uint32_t hist[256];                       // pool phase: radix histogram (equal-distance pool ties)
    uint32_t kbits[PCG_KB_WORDS];        // kept neighbours that are pool members, by pool position
    uint32_t bits[PCG_SMALL_MAX / 32];        // kept row positions (fallback membership test)
    int xw[32];
};

// One item handled by NT threads with the whole row in REGISTERS: NE ids/keys per thread, row position
// of slot e is e*NT + tid. Two memory round trips (ids, then scores), selection without histograms, kept
// list written in row order. kbits/bits: see oversample(); cand/wsum/xw: group scratch (CTA only).
template <int NT, int NE>
__device__ __forceinline__ void row_in_regs(const ChooseP& p, const Item& it, const int32_t* __restrict__ nbr,
                                            uint32_t* kbits, uint32_t* bits, uint32_t* hist, int* xw) {
    const int tid = NT == 32 ? (threadIdx.x & 31) : threadIdx.x;
    const int lane = tid & 31;
    const int w = it.w;
    const int d = it.d, k = it.k;
    const float* __restrict__ escore = p.entry_score ? p.entry_score + it.beg : nullptr;
    const int32_t* __restrict__ epp = it.use_kb ? p.entry_pool_pos + it.beg : nullptr;
    const bool all = k >= d;
    const bool need_dist = !all || p.sel_dist != nullptr;
    int32_t id[NE];
    int pp[NE];
    uint32_t key[NE];
    uint32_t vmask = 0;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int j = e * NT + tid;
        const bool valid = j < d;
        vmask |= (uint32_t)valid << e;
        id[e] = valid ? __ldg(nbr + j) : 0;
        pp[e] = (valid && epp) ? __ldg(epp + j) : -1;
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int j = e * NT + tid;
        uint32_t x = 0u;
        if (((vmask >> e) & 1u) && need_dist) x = dist_bits(it.sv, escore ? escore[j] : __ldg(p.score + id[e]));
        key[e] = x;
    }
    TRACE(2);
    uint32_t T = 0xffffffffu;
    int need = 0x7fffffff;
    if (!all) {
        if (k > 0) hist_select<NT, NE>(key, vmask, k, hist, xw, tid, T, need);
        else { T = 0; need = 0; }
    }
    TRACE(3);
    int run_less = 0, run_tie = 0;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e * NT < d) {                               // group-uniform
            const bool valid = (vmask >> e) & 1u;
            const bool less = valid && (all || key[e] < T);
            const bool tie = valid && !all && key[e] == T;
            int el, et, tl, tt;
            grp_excl2<NT>(less, tie, tid, xw, el, et, tl, tt);
            const int tie_before = run_tie + et;
            const bool sel = less || (tie && tie_before < need);
            if (sel) {
                const int64_t at = it.off + run_less + el + min(tie_before, need);
                p.sel_idx[at] = id[e];
                if (p.sel_dist) p.sel_dist[at] = __uint_as_float(key[e]);
                if (pp[e] >= 0) atomicOr(&kbits[pp[e] >> 5], 1u << (pp[e] & 31));
            }
            if (it.want_bits) {
                const unsigned sm = __ballot_sync(PCG_FULL, sel);
                if (lane == 0) bits[(e * NT + tid) >> 5] = sm;
            }
            run_less += tl;
            run_tie += tt;
        }
    }
    grp_sync<NT>();
    TRACE(4);
}
__device__ void choose_item_warp(const ChooseP& p, int w, WarpSmem& s) {
    const int lane = threadIdx.x & 31;
    const int tid = lane;
    Item it;
    item_header(p, w, it);
    TRACE(0); TRACE_VAL(8, it.d); TRACE_VAL(9, it.k); TRACE_VAL(10, it.o);
    TRACE(1);
    const int32_t* __restrict__ nbr = p.indices + it.beg;
    if (it.use_kb) {
        for (int q = lane; q < (p.P + 31) >> 5; q += 32) s.kbits[q] = 0u;
        __syncwarp();
    }
    const int d = it.d;
    if (d <= 32) row_in_regs<32, 1>(p, it, nbr, s.kbits, s.bits, s.shist, s.xw);
    else if (d <= 64) row_in_regs<32, 2>(p, it, nbr, s.kbits, s.bits, s.shist, s.xw);
    else if (d <= 128) row_in_regs<32, 4>(p, it, nbr, s.kbits, s.bits, s.shist, s.xw);
    else row_in_regs<32, 8>(p, it, nbr, s.kbits, s.bits, s.shist, s.xw);
    int m = it.k;
    if (it.o > 0) m += oversample<32>(p, it, lane, nbr, s.kbits, s.bits, s.hist, s.xw);
    TRACE(6);
    item_finish<32>(p, it, lane, m);
    __syncwarp();
    TRACE(7);
}


// ------------------------------------------------------------------------ CTA tiers
// One item handled by one CTA of NT threads (256 or 1024) with the row in REGISTERS: NE keys per thread,
// position of entry e = e*NT + tid. Two memory round trips (ids, then scores), histogram selection, then an
// ordered compaction whose offsets come from ONE scan over the per-(round, warp) counts.
template <int NT>
struct CtaSmem {
    uint32_t hist[8 * 256];                  // selection histograms (zero between uses)
    uint32_t chist[2 * 256];                 // cluster tier: this CTA's merged histogram (parity double buffer)
    int ctot;                                // cluster tier: this CTA's compaction totals
    int cnt[16 * (NT / 32)];                 // compaction: per (round, warp) counts, less | tie << 16
    int pre[16 * (NT / 32)];                 // their exclusive scan
    uint32_t ohist[256];                     // oversampling: radix histogram / broadcast words
    uint32_t kbits[PCG_KB_WORDS];            // kept neighbours that are pool members, by pool position
    uint32_t bits[PCG_CL_MAX / 32];          // kept row positions (fallback membership test; rank 0 of a cluster holds the row's)
    int xw[32];
};

// sid / spp: per-position stash of the ids and pool positions for the 1024-thread tier (each thread reads back
// only what it wrote: a register spill space that keeps the tier at 64 registers); NULL: kept in registers.
template <int NT, int NE, int CL>
__device__ __forceinline__ void cta_body(const ChooseP& p, const Item& it, CtaSmem<NT>& s, int32_t* sid, int16_t* spp,
                                         uint32_t* bits_local, int rank, int* par) {
    constexpr int NW = NT / 32;
    constexpr bool STASH = NT > PCG_GRP_NT;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int w = it.w;
    const int d = it.d, k = it.k;
    const int32_t* __restrict__ nbr = p.indices + it.beg;
    const float* __restrict__ escore = p.entry_score ? p.entry_score + it.beg : nullptr;
    const int32_t* __restrict__ epp = it.use_kb ? p.entry_pool_pos + it.beg : nullptr;
    const bool all = k >= d;
    const bool need_dist = !all || p.sel_dist != nullptr;
    const int pos0 = rank * (NE * NT) + tid;     // CTA `rank` of a cluster owns NE*NT consecutive positions
    uint32_t key[NE];
    int32_t id[NE];
    int pp[NE];
    uint32_t vmask = 0;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int j = pos0 + e * NT;
        const bool valid = j < d;
        vmask |= (uint32_t)valid << e;
        id[e] = valid ? __ldg(nbr + j) : 0;
        pp[e] = (valid && epp) ? __ldg(epp + j) : -1;
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int j = pos0 + e * NT;
        uint32_t x = 0u;
        if (((vmask >> e) & 1u) && need_dist) x = dist_bits(it.sv, escore ? escore[j] : __ldg(p.score + id[e]));
        key[e] = x;
        if (STASH) { sid[e * NT + tid] = id[e]; spp[e * NT + tid] = (int16_t)pp[e]; }   // CTA-local stash index
    }
    TRACE(2);
    uint32_t T = 0xffffffffu;
    int need = 0x7fffffff;
    if (!all) {
        if (k > 0) hist_select<NT, NE, CL>(key, vmask, k, s.hist, s.xw, tid, T, need, s.chist, par, w);
        else { T = 0; need = 0; }
    }
    TRACE(3);
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const bool valid = (vmask >> e) & 1u;
        const unsigned ml = __ballot_sync(PCG_FULL, valid && (all || key[e] < T));
        const unsigned mt = __ballot_sync(PCG_FULL, valid && !all && key[e] == T);
        if (lane == 0) s.cnt[e * NW + wid] = __popc(ml) | (__popc(mt) << 16);
    }
    __syncthreads();
    if (CL > 1) TRACE0(28);
    if (wid == 0) {
        constexpr int PER = (NE * NW + 31) / 32;       // consecutive entries per lane
        int v[PER], sum = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int at = lane * PER + q;
            v[q] = at < NE * NW ? s.cnt[at] : 0;
            sum += v[q];
        }
        int incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(PCG_FULL, incl, off);
            if (lane >= off) incl += t;
        }
        int run = incl - sum;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int at = lane * PER + q;
            if (at < NE * NW) s.pre[at] = run;
            run += v[q];
        }
        if (CL > 1 && lane == 31) s.ctot = incl;
    }
    int base_less = 0, base_tie = 0;
    uint32_t* kbits0 = s.kbits;
    uint32_t* bits0 = bits_local;
    if (CL > 1) {
        cg::cluster_group cl = cg::this_cluster();
        cl.sync();
        TRACE0(29);
        for (int c = 0; c < rank; ++c) {
            const int t = *cl.map_shared_rank(&s.ctot, c);
            base_less += t & 0xffff;
            base_tie += t >> 16;
        }
        kbits0 = cl.map_shared_rank(&s.kbits[0], 0);
        bits0 = cl.map_shared_rank(bits_local, 0);
    } else {
        __syncthreads();
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int j = pos0 + e * NT;
        const bool valid = (vmask >> e) & 1u;
        const bool less = valid && (all || key[e] < T);
        const bool tie = valid && !all && key[e] == T;
        const unsigned ml = __ballot_sync(PCG_FULL, less), mt = __ballot_sync(PCG_FULL, tie);
        const int pl = s.pre[e * NW + wid];
        const int tie_before = base_tie + (pl >> 16) + __popc(mt & lt);
        const bool sel = less || (tie && tie_before < need);
        if (sel) {
            const int64_t at = it.off + base_less + (pl & 0xffff) + __popc(ml & lt) + min(tie_before, need);
            p.sel_idx[at] = STASH ? sid[e * NT + tid] : id[e];
            if (p.sel_dist) p.sel_dist[at] = __uint_as_float(key[e]);
            const int q = STASH ? (int)spp[e * NT + tid] : pp[e];
            if (q >= 0) atomicOr(&kbits0[q >> 5], 1u << (q & 31));
        }
        if (it.want_bits) {
            const unsigned sm = __ballot_sync(PCG_FULL, sel);
            if (lane == 0 && j < d) bits0[j >> 5] = sm;
        }
    }
    TRACE(4);
}

template <int NT, int CL>
__device__ __forceinline__ void cta_item(const ChooseP& p, int w, CtaSmem<NT>& s, int32_t* sid, int16_t* spp,
                                         uint32_t* bits_local, int rank, int* par) {
    const int tid = threadIdx.x;
    Item it;
    item_header(p, w, it);
    TRACE(0); TRACE_VAL(8, it.d); TRACE_VAL(9, it.k); TRACE_VAL(10, it.o);
    if (rank == 0 && it.use_kb)
        for (int q = tid; q < (p.P + 31) >> 5; q += NT) s.kbits[q] = 0u;   // ordered before the ORs by the compaction's barriers
    TRACE(1);
    const int per = (it.d + CL * NT - 1) / (CL * NT);
    constexpr int NE_MAX = NT == PCG_GRP_NT ? 8 : 16;
    if (per <= 1) cta_body<NT, 1, CL>(p, it, s, sid, spp, bits_local, rank, par);
    else if (per <= 2) cta_body<NT, 2, CL>(p, it, s, sid, spp, bits_local, rank, par);
    else if (per <= 4 || NE_MAX == 4) cta_body<NT, 4, CL>(p, it, s, sid, spp, bits_local, rank, par);
    else if (per <= 8 || NE_MAX == 8) cta_body<NT, (NE_MAX < 8 ? NE_MAX : 8), CL>(p, it, s, sid, spp, bits_local, rank, par);
    else cta_body<NT, NE_MAX, CL>(p, it, s, sid, spp, bits_local, rank, par);
    if (CL > 1) {
        if (it.o > 0) cg::this_cluster().sync();      // kept bitmaps complete (remote ORs) before rank 0 reads them
        TRACE0(30);
        if (rank != 0) return;
    }
    __syncthreads();                      // kept bitmaps complete
    int m = it.k;
    if (it.o > 0) m += oversample<NT>(p, it, tid, p.indices + it.beg, s.kbits, bits_local, s.ohist, s.xw);
    TRACE(6);
    item_finish<NT>(p, it, tid, m);
    __syncthreads();
    TRACE(7);
}

// d <= 2048: one kernel of 256-thread CTAs serves both short tiers from two device-side queues: the rows of
// 256 < d <= 2048 (one per CTA) and the rows of d <= 256 (one per warp), so CTAs that get no
// (or short) CTA-tier rows take more of the warp-tier rows and no tier waits for the other's registers.
__global__ void __launch_bounds__(PCG_GRP_NT, 3) k_choose_small(ChooseP p) {
    constexpr size_t SM_BYTES = sizeof(CtaSmem<PCG_GRP_NT>) > sizeof(WarpSmem) * PCG_WARPS_PER_CTA
                                    ? sizeof(CtaSmem<PCG_GRP_NT>) : sizeof(WarpSmem) * PCG_WARPS_PER_CTA;
    __shared__ __align__(16) unsigned char raw[SM_BYTES];
    __shared__ int s_q;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // half of the CTAs start with the warp-tier queue, the other half with the cta-tier queue, so both drain
    // from the first microsecond; every CTA then helps with whatever is left of the other queue
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
        const bool cta_phase = ((blockIdx.x + phase) & 1) == 0;
        __syncthreads();                       // the two tiers lay the shared memory out differently
        if (cta_phase) {
            CtaSmem<PCG_GRP_NT>& s = *reinterpret_cast<CtaSmem<PCG_GRP_NT>*>(raw);
            for (int b = threadIdx.x; b < 8 * 256; b += PCG_GRP_NT) s.hist[b] = 0u;
            const int n = p.status[ST_NMID];
            for (;;) {
                __syncthreads();
                if (threadIdx.x == 0) s_q = atomicAdd(&p.status[ST_MID_CTR], 1);
                __syncthreads();
                const int q = s_q;
                if (q >= n) break;
                cta_item<PCG_GRP_NT, 1>(p, p.q_cta[q], s, nullptr, nullptr, s.bits, 0, nullptr);
            }
        } else {
            WarpSmem& ws = reinterpret_cast<WarpSmem*>(raw)[wid];
            for (int b = lane; b < 256; b += 32) ws.shist[b] = 0u;
            __syncwarp();
            const int n = p.status[ST_NSMALL];
            for (;;) {
                int q = 0;
                if (lane == 0) q = atomicAdd(&p.status[ST_SMALL_CTR], 1);
                q = __shfl_sync(PCG_FULL, q, 0);
                if (q >= n) break;
                choose_item_warp(p, p.q_warp[q], ws);
            }
        }
    }
}

// 2048 < d <= 16384: one item per CLUSTER of 8 CTAs x 256 threads (<= 8 entries per thread): the longest rows are
// the critical path of the step, so they get 8 SMs each and are launched first.
__global__ void __cluster_dims__(PCG_CL, 1, 1) __launch_bounds__(PCG_GRP_NT, 3) k_choose_wide(ChooseP p) {
    __shared__ CtaSmem<PCG_GRP_NT> s;
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    for (int b = threadIdx.x; b < 8 * 256; b += PCG_GRP_NT) s.hist[b] = 0u;
    __syncthreads();
    const int n = p.status[ST_NCL];
    const int n_cl = gridDim.x / PCG_CL, cid = blockIdx.x / PCG_CL;
    int par = 0;
    for (int q = cid; q < n; q += n_cl) cta_item<PCG_GRP_NT, PCG_CL>(p, p.q_cl[q], s, nullptr, nullptr, s.bits, rank, &par);
    cl.sync();          // nobody leaves while a peer may still address its shared memory
}

// 16384 < d <= 131072: one item per CLUSTER of 8 CTAs x 1024 threads, <= 16 keys per thread in registers, the ids and
// pool positions stashed in shared memory (power-law hubs of config C5; one CTA alone needs ~25 us per 10^4 entries).
#define PCG_HUGE_PER_CTA (PCG_HUGE_MAX / PCG_CL)
template <int CL>
__global__ void __launch_bounds__(PCG_LARGE_NT, 1) k_choose_huge(ChooseP p) {
    extern __shared__ uint32_t dyn_huge[];           // ids [PER_CTA] int32 | pool positions [PER_CTA] int16 | row bits
    int32_t* sid = reinterpret_cast<int32_t*>(dyn_huge);
    int16_t* spp = reinterpret_cast<int16_t*>(dyn_huge + PCG_HUGE_PER_CTA);
    uint32_t* bits = dyn_huge + PCG_HUGE_PER_CTA + PCG_HUGE_PER_CTA / 2;      // [CL * PER_CTA / 32], used on rank 0
    __shared__ CtaSmem<PCG_LARGE_NT> s;
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    for (int b = threadIdx.x; b < 8 * 256; b += PCG_LARGE_NT) s.hist[b] = 0u;
    __syncthreads();
    const int n = p.status[CL == 16 ? ST_NHUGE16 : ST_NHUGE];
    const int32_t* queue = CL == 16 ? p.q_huge16 : p.q_huge;
    const int n_cl = gridDim.x / CL, cid = blockIdx.x / CL;
    int par = 0;
    for (int q = cid; q < n; q += n_cl) cta_item<PCG_LARGE_NT, CL>(p, queue[q], s, sid, spp, bits, rank, &par);
    cl.sync();
}

// cluster launch with a run-time cluster size (16 is a non-portable size: opt-in attribute + occupancy check)
template <int CL>
static cudaError_t launch_huge(const ChooseP& p, int n_clusters, cudaStream_t stream, bool probe_only, int* max_clusters) {
    const size_t dyn = (size_t)PCG_HUGE_PER_CTA * 6 + (size_t)CL * PCG_HUGE_PER_CTA / 8;
    static bool configured_dev[PCG_MAX_DEVICES];      // function attributes are per device
    bool& configured = configured_dev[pcg_current_device()];
    cudaError_t e;
    if (!configured) {
        e = cudaFuncSetAttribute(k_choose_huge<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return e;
        if (CL > 8) {
            e = cudaFuncSetAttribute(k_choose_huge<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) return e;
        }
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_clusters * CL));
    cfg.blockDim = dim3(PCG_LARGE_NT);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (probe_only) return cudaOccupancyMaxActiveClusters(max_clusters, k_choose_huge<CL>, &cfg);
    return cudaLaunchKernelEx(&cfg, k_choose_huge<CL>, p);
}

// --------------------------------------------------------------------------------- big tier
// Rows beyond PCG_CL_MAX entries (rare hubs): one 1024-thread CTA per item; distances live in shared memory
// up to `sd_cap` entries and are recomputed per pass beyond; kept-position bits in a global slab.
__device__ void choose_item_big(const ChooseP& p, int w, uint32_t* sd, int sd_cap, uint32_t* hist, uint32_t* kbits,
                                uint32_t* cand, uint32_t* bits, int* xw) {
    constexpr int NT = PCG_LARGE_NT, NW = NT / 32;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    Item it;
    item_header(p, w, it);
    TRACE(0); TRACE_VAL(8, it.d); TRACE_VAL(9, it.k); TRACE_VAL(10, it.o);
    if (it.use_kb)
        for (int q = tid; q < (p.P + 31) >> 5; q += NT) kbits[q] = 0u;
    __syncthreads();
    TRACE(1);
    const int d = it.d, k = it.k;
    const float sv = it.sv;
    const int32_t* __restrict__ nbr = p.indices + it.beg;
    const float* __restrict__ escore = p.entry_score ? p.entry_score + it.beg : nullptr;
    const float* __restrict__ score = p.score;
    const bool cached = d <= sd_cap;
    if (cached) {
#pragma unroll 4
        for (int j = tid; j < d; j += NT) sd[j] = dist_bits(sv, escore ? escore[j] : __ldg(score + __ldg(nbr + j)));
        __syncthreads();
    }
    auto get = [&](int j) -> uint32_t {
        return cached ? sd[j] : dist_bits(sv, escore ? escore[j] : __ldg(score + __ldg(nbr + j)));
    };
    TRACE(2);
    uint32_t T = 0xffffffffu;
    int need = 0x7fffffff;
    if (k < d) {
        if (k > 0) cta_bitselect<NT>(get, d, k, cand, PCG_CAND_CAP, reinterpret_cast<int*>(hist), xw, tid, T, need);
        else { T = 0; need = 0; }
    }
    TRACE(3);
    // ---- ordered compaction: every warp owns a contiguous chunk of the row; count, one scan of the
    // per-warp counts, then each warp writes its chunk using ballots only ----
    const int chunk = (((d + NW - 1) / NW) + 31) & ~31;       // multiple of 32 positions per warp
    const int cb = min(d, wid * chunk), ce = min(d, cb + chunk);
    int cl = 0, ct = 0;
    if (k < d) {
        for (int base = cb; base < ce; base += 32) {
            const int j = base + lane;
            const uint32_t key = j < ce ? get(j) : 0xffffffffu;
            cl += __popc(__ballot_sync(PCG_FULL, j < ce && key < T));
            ct += __popc(__ballot_sync(PCG_FULL, j < ce && key == T));
        }
    } else {
        cl = ce - cb;
    }
    if (lane == 0) { xw[wid] = cl; hist[wid] = (uint32_t)ct; }   // hist is free here
    __syncthreads();
    int run_less = 0, run_tie = 0;
    for (int q = 0; q < wid; ++q) { run_less += xw[q]; run_tie += (int)hist[q]; }
    __syncthreads();
    const int32_t* __restrict__ epp = it.use_kb ? p.entry_pool_pos + it.beg : nullptr;
    int32_t id_n = cb + lane < ce ? __ldg(nbr + cb + lane) : 0;
    int pp_n = (epp && cb + lane < ce) ? __ldg(epp + cb + lane) : -1;
    for (int base = cb; base < ce; base += 32) {
        const int j = base + lane;
        const bool valid = j < ce;
        const int32_t id = id_n;
        const int pp = pp_n;
        if (j + 32 < ce) {                                   // prefetch the next tile
            id_n = __ldg(nbr + j + 32);
            pp_n = epp ? __ldg(epp + j + 32) : -1;
        }
        const uint32_t key = (valid && (k < d || p.sel_dist)) ? get(j) : 0u;
        const bool less = valid && (k >= d || key < T);
        const bool tie = valid && k < d && key == T;
        const unsigned lt = lanemask_lt();
        const unsigned ml = __ballot_sync(PCG_FULL, less), mt = __ballot_sync(PCG_FULL, tie);
        const int tie_before = run_tie + __popc(mt & lt);
        const bool sel = less || (tie && tie_before < need);
        if (sel) {
            const int64_t at = it.off + run_less + __popc(ml & lt) + min(tie_before, need);
            p.sel_idx[at] = id;
            if (p.sel_dist) p.sel_dist[at] = __uint_as_float(key);
            if (pp >= 0) atomicOr(&kbits[pp >> 5], 1u << (pp & 31));
        }
        if (it.want_bits) {
            const unsigned sm = __ballot_sync(PCG_FULL, sel);
            if (lane == 0) bits[base >> 5] = sm;
        }
        run_less += __popc(ml);
        run_tie += __popc(mt);
    }
    __syncthreads();   // kept bitmaps complete and visible to the whole CTA
    TRACE(4);
    int m = k;
    if (it.o > 0) m += oversample<NT>(p, it, tid, nbr, kbits, bits, hist, xw);
    TRACE(6);
    item_finish<NT>(p, it, tid, m);
    __syncthreads();
    TRACE(7);
}

// d > 32768: one item per 1024-thread CTA, one CTA per SM.
__global__ void __launch_bounds__(PCG_LARGE_NT, 1) k_choose_big(ChooseP p, int sd_cap) {
    extern __shared__ uint32_t dyn[];                // [sd_cap] distance cache
    __shared__ uint32_t hist[256];
    __shared__ uint32_t kbits[PCG_KB_WORDS];
    __shared__ uint32_t cand[PCG_CAND_CAP];
    __shared__ int xw[32];
    const int n = p.status[ST_NBIG];
    for (int q = blockIdx.x; q < n; q += gridDim.x)
        choose_item_big(p, p.q_big[q], dyn, sd_cap, hist, kbits, cand, p.bits_slab + (int64_t)blockIdx.x * p.slab_words, xw);
}

// ------------------------------------------------------------------------------------ prep
// (1) Batches drawn by pick_step are sampled with replacement in proportion to degree, so hub
// nodes appear several times: an item whose target id already occurred earlier in the batch is not
// processed, it shares the result of that earlier ("representative") item (same node, same relation,
// same label). first[v] = smallest batch index whose target is node v; the table lives in the workspace,
// is all 0x7f7f7f7f between calls and is restored before this kernel ends. (2) Sizes k, o of every
// representative item in the reference's arithmetic; its output slots by an exclusive prefix sum in
// item order (so the slot layout is deterministic); items that do not fit `cap_slots` are
// dropped and flagged. (3) Tier queues by row length.
// The items are split into contiguous chunks over up to #SMs co-resident CTAs which meet at two grid barriers
// (table complete / chunk totals published). Phase A computes the sizes of a CTA's items with four items per
// thread in flight and parks them in shared memory; phase B is a prefix sum over contiguous runs (slot order
// = item order w = r*B + i).
#define PCG_PREP_CHUNK 512       // items per CTA when the batch is small (more CTAs = shorter critical path)
#define PCG_PREP_ITEMS 8192      // most items one CTA stages (32 KB of shared memory)

// Barrier over the co-resident CTAs of the prep grid (grid <= #SMs, one CTA fits every SM): bar[0] counts
// arrivals of the current launch, the last CTA to leave the kernel resets the words (prep_exit).
__device__ __forceinline__ void prep_grid_barrier(int32_t* bar, int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&bar[0], 1);
        while (*(volatile int32_t*)&bar[0] < target) { }
        __threadfence();
    }
    __syncthreads();
}

// CLUSTER: the CTAs form ONE thread-block cluster and meet at cluster barriers (hardware barrier, and the kernel
// runs next to other kernels: the score table and the pool sort on the main stream); otherwise (more items than a
// cluster can stage) a cooperative grid with barriers through global memory.
template <bool CLUSTER>
__global__ void __launch_bounds__(PCG_PREP_NT, 1) k_choose_prep(ChooseP p, int chunk, int32_t* bar, int32_t* totals) {
    __shared__ int s_info[PCG_PREP_ITEMS];   // per item of the CTA: nslots << 3 | (tier + 1), 0 for repeated targets
    __shared__ int s_wsum[32];
    __shared__ int s_base;
    constexpr int NT = PCG_PREP_NT;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int B = p.B, W = p.R * p.B, G = gridDim.x;
    int32_t* first = p.first;
    PTRACE(0);
    if (first)
        for (int i = blockIdx.x * NT + tid; i < B; i += G * NT) {
            const int64_t lv = (int64_t)__ldg(p.targets + i) - p.row_lo;
            if (lv >= 0 && lv < p.n_nodes) atomicMin(&first[lv], i);
        }
    if (blockIdx.x == 0 && tid < PCG_STATUS_WORDS) p.status[tid] = 0;
    auto barrier = [&](int target) {
        if (CLUSTER) { __threadfence(); cg::this_cluster().sync(); }
        else prep_grid_barrier(bar, target);
    };
    barrier(G);                              // the first-occurrence table is complete, the status words are zero
    PTRACE(1);
    const int w0 = min(W, blockIdx.x * chunk), n_items = min(W, w0 + chunk) - w0;
    // ---- phase A: sizes of every item, four items per thread in flight (loads first, then the arithmetic)
    for (int q0 = tid; q0 < n_items; q0 += 4 * NT) {
        int rr[4], ii[4], rep[4];
        int64_t beg[4], end[4];
        bool pos[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = q0 + u * NT;
            const int w = w0 + q;
            rr[u] = w / B;
            ii[u] = w - rr[u] * B;
            rep[u] = -1;
            beg[u] = end[u] = 0;
            pos[u] = false;
            if (q < n_items) {
                const int64_t lv = (int64_t)__ldg(p.targets + ii[u]) - p.row_lo;
                if (lv >= 0 && lv < p.n_nodes) {
                    rep[u] = first ? __ldcg(first + lv) : ii[u];
                    const int64_t row = (int64_t)rr[u] * p.n_nodes + lv;
                    beg[u] = __ldg(p.indptr + row);
                    end[u] = __ldg(p.indptr + row + 1);
                    pos[u] = p.train && p.labels && __ldg(p.labels + ii[u]) == 1;
                } else {
                    rep[u] = -2;             // a target whose row this CSR does not hold: flagged below
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = q0 + u * NT;
            if (q < n_items) {
                const int w = w0 + q;
                int info = 0;
                p.it_rep[w] = rep[u] >= 0 ? rr[u] * B + rep[u] : w;
                if (rep[u] == -2) {          // not our row: empty item, error flag
                    p.it_slot0[w] = 0; p.it_m[w] = 0; p.it_base[w] = 0; p.it_done[w] = 0;
                    atomicExch(&p.status[ST_OVERFLOW], 2);
                    atomicOr(p.sticky, 2);
                }
                if (rep[u] == ii[u]) {
                    const int64_t d = end[u] - beg[u];
                    int k, o;
                    item_counts(d, p.thresh[rr[u]], p.rho, pos[u], p.P, p.k_override ? p.k_override[w] : 0,
                                p.k_override != nullptr, k, o);
                    const int tier = d <= PCG_SMALL_MAX ? 0 : (d <= PCG_CTA_MAX ? 1 : (d <= PCG_CL_MAX ? 2 : (d <= PCG_HUGE_MAX ? 3 :
                                     ((d <= PCG_HUGE16_MAX && p.huge16_ok) ? 4 : 5))));
                    info = (((k + o + PCG_SLOT - 1) / PCG_SLOT) << 3) | (tier + 1);
                }
                s_info[q] = info;
            }
        }
    }
    __syncthreads();
    PTRACE(2);
    // ---- phase B: exclusive prefix sum of the slot counts in item order; every thread owns a contiguous run
    const int per = (n_items + NT - 1) / NT;
    const int c0 = min(n_items, tid * per), c1 = min(n_items, c0 + per);
    int mine = 0;
    for (int q = c0; q < c1; ++q) mine += s_info[q] >> 3;
    int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(PCG_FULL, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) s_wsum[wid] = incl;
    __syncthreads();
    const int ws = s_wsum[lane];
    int wincl = ws;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(PCG_FULL, wincl, off);
        if (lane >= off) wincl += t;
    }
    const int cta_total = __shfl_sync(PCG_FULL, wincl, 31);
    if (tid == 0) totals[blockIdx.x] = cta_total;
    barrier(2 * G);                          // every CTA's total is published (and every read of `first` is done)
    if (wid == 0) {                          // slots of the CTAs before this one, added in CTA order
        int acc = 0;
        for (int c = lane; c < (int)blockIdx.x; c += 32) acc += __ldcg(totals + c);
        acc = __reduce_add_sync(PCG_FULL, acc);
        if (lane == 0) s_base = acc;
    }
    __syncthreads();
    int slot0 = s_base + __shfl_sync(PCG_FULL, wincl - ws, wid) + incl - mine;
    bool overflow = false;
    int32_t* const queues[6] = {p.q_warp, p.q_cta, p.q_cl, p.q_huge, p.q_huge16, p.q_big};
    const int counters[6] = {ST_NSMALL, ST_NMID, ST_NCL, ST_NHUGE, ST_NHUGE16, ST_NBIG};
    for (int c = 0; c < per; ++c) {                          // uniform trip count (ballots inside)
        const int q = c0 + c;
        int tier = -1;
        if (q < c1) {
            const int info = s_info[q];
            if (info) {
                const int w = w0 + q, nsl = info >> 3;
                tier = (info & 7) - 1;
                if ((int64_t)slot0 + nsl > p.cap_slots) {    // does not fit: flag, emit nothing for this item
                    overflow = true;
                    tier = -1;
                    p.it_slot0[w] = 0; p.it_m[w] = 0; p.it_base[w] = 0; p.it_done[w] = 0;
                    for (int x = 0; x < nsl; ++x)
                        if ((int64_t)slot0 + x < p.cap_slots) p.slot_item[slot0 + x] = -1;
                } else {
                    p.it_slot0[w] = slot0;
                }
                slot0 += nsl;
            }
        }
        // one queue reservation per (warp, tier): the up to six atomics of a round are issued together by six lanes
        // (one L2 round trip instead of one per tier present in the warp)
        unsigned tm[6];
#pragma unroll
        for (int t = 0; t < 6; ++t) tm[t] = __ballot_sync(PCG_FULL, tier == t);
        int qb = 0;
#pragma unroll
        for (int t = 0; t < 6; ++t)
            if (lane == t && tm[t]) qb = atomicAdd(&p.status[counters[t]], __popc(tm[t]));
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const int b = __shfl_sync(PCG_FULL, qb, t);
            if (tier == t) queues[t][b + __popc(tm[t] & lanemask_lt())] = w0 + q;
        }
    }
    PTRACE(3);
    if (__syncthreads_or(overflow) && tid == 0) { atomicExch(&p.status[ST_OVERFLOW], 1); atomicOr(p.sticky, 1); }
    if (first)                               // restore the table (all its reads are behind the second barrier)
        for (int i = blockIdx.x * NT + tid; i < B; i += G * NT) {
            const int64_t lv = (int64_t)__ldg(p.targets + i) - p.row_lo;
            if (lv >= 0 && lv < p.n_nodes) first[lv] = 0x7f7f7f7f;
        }
    if (tid == 0) {
        if (blockIdx.x == G - 1) p.status[ST_SLOTS] = s_base + cta_total;
        if (!CLUSTER) {
            __threadfence();
            if (atomicAdd(&bar[1], 1) == G - 1) { bar[0] = 0; bar[1] = 0; }   // last CTA out re-arms the barrier
        }
    }
    PTRACE(4);
}

// Select-all (GraphSAGE / GCN): the item list is the CSR row itself. One warp per item.
__global__ void k_select_all(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t n_nodes,
                             int R, const int32_t* __restrict__ targets, int B, int add_self, int64_t cap_slots,
                             int32_t* slot_item, int32_t* it_slot0, int32_t* it_m, int64_t* it_base, int32_t* it_extra,
                             int32_t* it_done, int32_t* status) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= R * B) return;
    const int r = w / B, i = w - r * B;
    const int32_t v = targets[i];
    const int64_t row = (int64_t)r * n_nodes + v;
    const int64_t beg = indptr[row];
    const int d = (int)(indptr[row + 1] - beg);
    int extra = -1;
    if (add_self) {   // graphsage.py:210: union with {self}
        int lo = 0, hi = d;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (indices[beg + mid] < v) lo = mid + 1; else hi = mid;
        }
        if (!(lo < d && indices[beg + lo] == v)) extra = v;
    }
    int nslots = (d + PCG_SLOT - 1) / PCG_SLOT;
    if (nslots == 0 && extra >= 0) nslots = 1;
    int slot0 = 0;
    if (lane == 0) slot0 = atomicAdd(&status[ST_SLOTS], nslots);
    slot0 = __shfl_sync(PCG_FULL, slot0, 0);
    const bool ovf = (int64_t)slot0 + nslots > cap_slots;
    if (lane == 0) {
        if (ovf) atomicExch(&status[ST_OVERFLOW], 1);
        it_slot0[w] = ovf ? 0 : slot0;
        it_m[w] = ovf ? 0 : d;
        it_base[w] = beg;
        it_extra[w] = ovf ? -1 : extra;
        it_done[w] = 0;
    }
    for (int c = lane; c < nslots; c += 32)
        if ((int64_t)slot0 + c < cap_slots) slot_item[slot0 + c] = ovf ? -1 : w;
}

// pool_pos_of[v] = position of node v in the pool, -1 for everyone else (built once per pool).
__global__ void k_fill_i32(int32_t* a, int64_t n, int32_t val) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = val;
}
__global__ void k_pool_positions(const int32_t* __restrict__ pool, int P, int32_t* __restrict__ pool_pos_of) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) pool_pos_of[pool[i]] = i;
}

__global__ void k_entry_pool_pos(const int32_t* __restrict__ indices, int64_t nnz,
                                 const int32_t* __restrict__ pool_pos_of, int32_t* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nnz) out[e] = pool_pos_of[indices[e]];
}
// ------------------------------------------------------------------------------------------- C ABI
struct WsLayout {
    size_t first, bar, totals, bits_slab, q_warp, q_cta, q_cl, q_huge, q_huge16, q_big, total;
    int64_t slab_words;
    int grid_big;
};

// `first` comes first: its place must not depend on the batch size (it carries state between calls)
static WsLayout ws_layout(int B, int R, int64_t max_degree, int64_t n_nodes, int sms) {
    WsLayout L;
    size_t W = (size_t)B * R;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    L.grid_big = sms;
    L.slab_words = max_degree > PCG_HUGE_MAX ? (max_degree + 31) / 32 : 0;      // (also when 16-CTA clusters take those rows)
    size_t o = 0;
    L.first = o; o = al(o + (size_t)n_nodes * 4);
    L.bar = o; o = al(o + 16);                   // prep grid barrier words (zero between calls) + the sticky overflow word
    L.totals = o; o = al(o + 1024 * 4);          // per-CTA slot totals of the prep kernel
    L.bits_slab = o; o = al(o + (size_t)L.grid_big * L.slab_words * 4);
    L.q_warp = o; o = al(o + W * 4);
    L.q_cta = o; o = al(o + W * 4);
    L.q_cl = o; o = al(o + W * 4);
    L.q_huge = o; o = al(o + W * 4);
    L.q_huge16 = o; o = al(o + W * 4);
    L.q_big = o; o = al(o + W * 4);
    L.total = o;
    return L;
}

// Host-side state of this file is kept PER DEVICE ordinal (side streams, fork / join events, probe results): a process
// that drives several GPUs gets a separate set for each (the calls themselves are not thread-safe per device).
static int g_sms[PCG_MAX_DEVICES];
static int device_sms() {
    const int dev = pcg_current_device();
    if (g_sms[dev] == 0) {
        if (cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_sms[dev] <= 0)
            g_sms[dev] = 148;
    }
    return g_sms[dev];
}

extern "C" int pcg_device_sms(void) { return device_sms(); }

#ifdef PCG_TRACE
extern "C" __attribute__((visibility("default"))) int pcg_debug_set_trace(long long* buf) {
    return (int)cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf));
}
#endif

extern "C" size_t pcg_choose_workspace_bytes(int B, int R, int64_t max_degree, int64_t n_nodes) {
    return ws_layout(B, R, max_degree, n_nodes, 148 * 2).total;   // sized for the largest grid we ever launch
}

extern "C" size_t pcg_choose_sticky_offset(int64_t n_nodes) { return ws_layout(1, 1, 0, n_nodes, 148).bar + 8; }

extern "C" int pcg_choose_workspace_init(void* workspace, size_t workspace_bytes, int64_t n_nodes, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    WsLayout L = ws_layout(1, 1, 0, n_nodes, 148);
    PCG_REQUIRE(workspace && workspace_bytes >= L.bar + 16, "pcg_choose_workspace_init: workspace too small");
    cudaError_t e = cudaMemsetAsync(workspace, 0x7f, (size_t)n_nodes * 4, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync((char*)workspace + L.bar, 0, 16, stream);
    if (e != cudaSuccess) { pcg_set_error("pcg_choose_workspace_init: memset: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" int pcg_pool_positions(const int32_t* pool, int P, int64_t n_nodes, int32_t* pool_pos_of,
                                  pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCG_REQUIRE(pool_pos_of && (P == 0 || pool), "pcg_pool_positions: null pointer");
    if (n_nodes > 0) k_fill_i32<<<(int)((n_nodes + 255) / 256), 256, 0, stream>>>(pool_pos_of, n_nodes, -1);
    if (P > 0) k_pool_positions<<<(P + 255) / 256, 256, 0, stream>>>(pool, P, pool_pos_of);
    return pcg_check_launch("pcg_pool_positions");
}

extern "C" int pcg_entry_pool_positions(const int32_t* indices, int64_t nnz, const int32_t* pool_pos_of,
                                        int32_t* entry_pool_pos, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCG_REQUIRE(nnz == 0 || (indices && pool_pos_of && entry_pool_pos), "pcg_entry_pool_positions: null pointer");
    if (nnz > 0)
        k_entry_pool_pos<<<(unsigned)((nnz + 255) / 256), 256, 0, stream>>>(indices, nnz, pool_pos_of, entry_pool_pos);
    return pcg_check_launch("pcg_entry_pool_positions");
}

// side streams + events so the tiers run side by side (fork/join; capturable)
#define PCG_N_SIDE 3
struct SideState {
    cudaStream_t side[PCG_N_SIDE];
    cudaEvent_t fork, join[PCG_N_SIDE];
    int huge16_state;            // can this device place clusters of 16 x 1024 threads? 0 = not probed, 1 = yes, 2 = no
    size_t big_smem;             // dynamic shared memory k_choose_big is configured for
};
static SideState g_dev[PCG_MAX_DEVICES];

extern "C" int pcg_choose(const int64_t* indptr, const int32_t* indices, int64_t n_nodes, int64_t row_lo, int R,
                          const float* score,
                          const float* entry_score, const float* center_score, const int32_t* targets,
                          const int64_t* labels, int B, const double* thresh_host, const int32_t* k_override,
                          double rho, const float* ps_score, const int32_t* ps_pos, const int32_t* ps_id,
                          const int32_t* entry_pool_pos, int P, int train,
                          int64_t max_degree, int32_t* sel_idx, float* sel_dist, int64_t cap_slots,
                          int32_t* slot_item, int32_t* it_slot0, int32_t* it_m, int64_t* it_base, int32_t* it_done,
                          int32_t* it_rep, void* workspace, size_t workspace_bytes, int32_t* status, int phases,
                          pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCG_REQUIRE(R >= 1 && R <= PCG_MAX_REL, "pcg_choose: R=%d outside [1,%d]", R, PCG_MAX_REL);
    PCG_REQUIRE(phases >= 1 && phases <= 3, "pcg_choose: phases must be 1 (prepare), 2 (select) or 3 (both)");
    PCG_REQUIRE(B >= 0, "pcg_choose: negative batch");
    if (B == 0) {   // empty batch: nothing to choose (pointers of empty buffers may be null)
        if (status) cudaMemsetAsync(status, 0, PCG_STATUS_WORDS * sizeof(int32_t), stream);
        return 0;
    }
    PCG_REQUIRE((phases & 2) == 0 || score || (entry_score && center_score), "pcg_choose: need a score table or explicit scores");
    PCG_REQUIRE(!(train && P > 0) || (ps_score && ps_pos && ps_id), "pcg_choose: sorted pool arrays missing");
    PCG_REQUIRE(indptr && indices && targets && sel_idx && slot_item && it_slot0 && it_m && it_base && it_done &&
                    it_rep && status,
                "pcg_choose: null pointer");
    const int sms = device_sms();
    WsLayout L = ws_layout(B, R, max_degree, n_nodes, sms);
    PCG_REQUIRE(workspace && workspace_bytes >= L.total, "pcg_choose: workspace too small (%zu < %zu)", workspace_bytes,
                L.total);
    cudaError_t e;
    ChooseP p;
    p.indptr = indptr; p.indices = indices; p.score = score; p.entry_score = entry_score;
    p.center_score = center_score; p.targets = targets; p.labels = labels; p.k_override = k_override;
    p.ps_score = ps_score; p.ps_pos = ps_pos; p.ps_id = ps_id; p.entry_pool_pos = entry_pool_pos;
    p.n_nodes = n_nodes; p.row_lo = row_lo; p.R = R; p.B = B;
    p.P = (train && ps_score) ? P : 0; p.train = train;
    for (int r = 0; r < PCG_MAX_REL; ++r) p.thresh[r] = r < R ? thresh_host[r] : 0.5;
    p.rho = rho; p.sel_idx = sel_idx; p.sel_dist = sel_dist; p.cap_slots = cap_slots; p.slot_item = slot_item;
    p.it_slot0 = it_slot0; p.it_m = it_m; p.it_base = it_base; p.it_done = it_done; p.it_rep = it_rep; p.status = status;
    char* ws = (char*)workspace;
    // duplicate targets are folded when ids come from a score table (explicit per-target lists are all distinct)
    p.first = entry_score == nullptr ? (int32_t*)(ws + L.first) : nullptr;
    p.q_warp = (int32_t*)(ws + L.q_warp);
    p.q_cta = (int32_t*)(ws + L.q_cta);
    p.q_cl = (int32_t*)(ws + L.q_cl);
    p.q_huge = (int32_t*)(ws + L.q_huge);
    p.q_huge16 = (int32_t*)(ws + L.q_huge16);
    p.q_big = (int32_t*)(ws + L.q_big);
    p.sticky = (int32_t*)(ws + L.bar) + 2;
    p.bits_slab = (uint32_t*)(ws + L.bits_slab);
    p.slab_words = L.slab_words;
    const int W = R * B;
    SideState& S = g_dev[pcg_current_device()];
    cudaStream_t* g_side = S.side;
    cudaEvent_t* g_join = S.join;
    cudaEvent_t& g_fork = S.fork;
    if (S.huge16_state == 0) {               // probed once per device
        int mc = 0;
        cudaError_t pe = launch_huge<16>(p, 1, stream, true, &mc);
        S.huge16_state = (pe == cudaSuccess && mc >= 1) ? 1 : 2;
        (void)cudaGetLastError();
    }
    p.huge16_ok = S.huge16_state == 1;
    if (phases & 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(PCG_PREP_NT);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (W <= 8 * PCG_PREP_ITEMS) {
            // one cluster of (up to) 8 CTAs, contiguous chunks of items
            int G = 8;
            while (G > 1 && W < G * 64) G >>= 1;
            const int chunk = (W + G - 1) / G;
            cfg.gridDim = dim3((unsigned)G);
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            e = cudaLaunchKernelEx(&cfg, k_choose_prep<true>, p, chunk, (int32_t*)(ws + L.bar), (int32_t*)(ws + L.totals));
        } else {
            // contiguous chunks of items over at most #SMs CTAs (they meet at grid barriers: all must be resident)
            int chunk = PCG_PREP_CHUNK;
            if ((W + chunk - 1) / chunk > sms) chunk = (W + sms - 1) / sms;
            PCG_REQUIRE(chunk <= PCG_PREP_ITEMS, "pcg_choose: batch too large (%d items; at most %d)", W, PCG_PREP_ITEMS * sms);
            cfg.gridDim = dim3((unsigned)((W + chunk - 1) / chunk));
            attr[0].id = cudaLaunchAttributeCooperative;
            attr[0].val.cooperative = 1;
            e = cudaLaunchKernelEx(&cfg, k_choose_prep<false>, p, chunk, (int32_t*)(ws + L.bar), (int32_t*)(ws + L.totals));
        }
        if (e != cudaSuccess) { pcg_set_error("pcg_choose: prep launch: %s", cudaGetErrorString(e)); return (int)e; }
    }
    if ((phases & 2) == 0) return pcg_check_launch("pcg_choose");
    const bool have_cta = max_degree > PCG_SMALL_MAX, have_cl = max_degree > PCG_CTA_MAX,
               have_huge = max_degree > PCG_CL_MAX,
               have_big = max_degree > (p.huge16_ok ? PCG_HUGE16_MAX : PCG_HUGE_MAX);
    if (have_cl && !g_fork) {
        // the long-row tiers are the critical path: their streams get the highest priority, so the block scheduler
        // places their CTAs (clusters of 8 need a whole GPC slot each) before the short-row kernel fills the SMs
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        for (int q = 0; q < PCG_N_SIDE; ++q)
            if ((e = cudaStreamCreateWithPriority(&g_side[q], cudaStreamNonBlocking, prio_hi)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&g_join[q], cudaEventDisableTiming)) != cudaSuccess) {
                pcg_set_error("pcg_choose: side stream: %s", cudaGetErrorString(e));
                return (int)e;
            }
        if ((e = cudaEventCreateWithFlags(&g_fork, cudaEventDisableTiming)) != cudaSuccess) {
            pcg_set_error("pcg_choose: event: %s", cudaGetErrorString(e));
            return (int)e;
        }
    }
    if (have_cl) cudaEventRecord(g_fork, stream);      // fork: the tiers run side by side, longest rows first
    if (have_big) {
        int64_t cap = max_degree > 49152 ? 49152 : (max_degree + 31) / 32 * 32;
        size_t dyn = (size_t)cap * 4;
        if (dyn > S.big_smem) {
            e = cudaFuncSetAttribute(k_choose_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            if (e != cudaSuccess) { pcg_set_error("pcg_choose: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
            S.big_smem = dyn;
        }
        cudaStreamWaitEvent(g_side[2], g_fork, 0);
        k_choose_big<<<W < L.grid_big ? W : L.grid_big, PCG_LARGE_NT, dyn, g_side[2]>>>(p, (int)cap);
        cudaEventRecord(g_join[2], g_side[2]);
    }
    if (have_huge) {
        cudaStreamWaitEvent(g_side[0], g_fork, 0);
        if (max_degree > PCG_HUGE_MAX && p.huge16_ok) {
            e = launch_huge<16>(p, W < 4 ? W : 4, g_side[0], false, nullptr);
            if (e != cudaSuccess) { pcg_set_error("pcg_choose: huge16 launch: %s", cudaGetErrorString(e)); return (int)e; }
        }
        e = launch_huge<8>(p, W < 16 ? W : 16, g_side[0], false, nullptr);
        if (e != cudaSuccess) { pcg_set_error("pcg_choose: huge launch: %s", cudaGetErrorString(e)); return (int)e; }
        cudaEventRecord(g_join[0], g_side[0]);
    }
    if (have_cl) {
        cudaStreamWaitEvent(g_side[1], g_fork, 0);
        const int n_cl = W < 56 ? W : 56;
        k_choose_wide<<<n_cl * PCG_CL, PCG_GRP_NT, 0, g_side[1]>>>(p);
        cudaEventRecord(g_join[1], g_side[1]);
    }
    {
        const int units = W / PCG_WARPS_PER_CTA + 1 + (have_cta ? W : 0);
        const int gs = units < sms * 3 ? units : sms * 3;     // (2 CTAs per SM, leaving room for the clusters, measured no better)
        k_choose_small<<<gs, PCG_GRP_NT, 0, stream>>>(p);
    }
    if (have_huge) cudaStreamWaitEvent(stream, g_join[0], 0);
    if (have_cl) cudaStreamWaitEvent(stream, g_join[1], 0);
    if (have_big) cudaStreamWaitEvent(stream, g_join[2], 0);
    return pcg_check_launch("pcg_choose");
}

extern "C" int pcg_select_all(const int64_t* indptr, const int32_t* indices, int64_t n_nodes, int R,
                              const int32_t* targets, int B, int add_self, int64_t cap_slots, int32_t* slot_item,
                              int32_t* it_slot0, int32_t* it_m, int64_t* it_base, int32_t* it_extra, int32_t* it_done,
                              int32_t* status, pcg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCG_REQUIRE(R >= 1 && B >= 0, "pcg_select_all: bad sizes");
    if (B == 0) {
        if (status) cudaMemsetAsync(status, 0, PCG_STATUS_WORDS * sizeof(int32_t), stream);
        return 0;
    }
    PCG_REQUIRE(indptr && indices && targets && slot_item && it_slot0 && it_m && it_base && it_extra && it_done && status,
                "pcg_select_all: null pointer");
    cudaError_t e = cudaMemsetAsync(status, 0, PCG_STATUS_WORDS * sizeof(int32_t), stream);
    if (e != cudaSuccess) { pcg_set_error("pcg_select_all: memset: %s", cudaGetErrorString(e)); return (int)e; }
    const int W = R * B;
    k_select_all<<<(W * 32 + 255) / 256, 256, 0, stream>>>(indptr, indices, n_nodes, R, targets, B, add_self, cap_slots,
                                                           slot_item, it_slot0, it_m, it_base, it_extra, it_done, status);
    return pcg_check_launch("pcg_select_all");
}
