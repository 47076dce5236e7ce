// pairgen.cu — seed-bucket pair-and-diagonal generator = minimap2's seeding + "-r 0" chaining.
//
// Restates (oracle/hlo_oracle.c:orc_candidates; minimap2 is absent from /root/reference, call site
// polyte.tune_params.py:510 "--sr -X ... -m 30 -n 2 -r 0"):
//   * index: minimizers bucketed by (cluster, hash); a bucket with more than mid_occ entries seeds nothing;
//   * -X: no self hits, a seed (q -> t) is skipped when strcmp(qname, tname) > 0;
//   * anchors of one (t, strand, diagonal) form one chain (bw = 0); chain score = k + sum(min(k, gap));
//     kept iff #anchors >= -n and score >= -m, which for -m <= 2k and -n <= 2 is
//     (max anchor pos - min anchor pos) >= max(m - k, n - 1, 0)  (DESIGN.md §pairgen).
//
// One warp per query read: lanes stream the read's buckets from the hash-sorted hit array (coalesced
// 8-byte entries) and fold every hit into a per-warp shared-memory hash table keyed by
// (t, strand, diagonal) that keeps min/max anchor position — the warp-level dedup of the north star.
// A read whose table overflows is redone exactly by pairgen_overflow_kernel with a table in HBM.
// Algorithmic bytes: 8 B per seed hit in + 16 B per candidate out.
#include "kernels.cuh"

namespace hlo {

#define HLO_EMPTY_KEY 0xffffffffffffffffULL

__global__ void hash_keys_kernel(const uint64_t *__restrict__ x, uint64_t n, uint64_t *__restrict__ keys) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = x[i] >> 8;
}
void launch_hash_keys(const uint64_t *mz_x, uint64_t n, uint64_t *keys, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    hash_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mz_x, n, keys);
    if (lc) lc->total++;
}

__global__ void mark_buckets_kernel(const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ svals,
                                    const uint32_t *__restrict__ mz_rid, const uint32_t *__restrict__ mz_rp,
                                    const uint32_t *__restrict__ cbase, uint64_t n, uint32_t *__restrict__ flags,
                                    uint2 *__restrict__ hits) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t idx = svals[i];
    const uint32_t rid = mz_rid[idx];
    hits[i] = make_uint2(rid, mz_rp[idx]);
    uint32_t f = 1;
    if (i > 0) {
        const uint32_t prid = mz_rid[svals[i - 1]];
        f = (skeys[i] != skeys[i - 1]) || (cbase[rid] != cbase[prid]);
    }
    flags[i] = f;
}
void launch_mark_buckets(const uint64_t *skeys, const uint32_t *svals, const uint32_t *mz_rid, const uint32_t *mz_rp,
                         const uint32_t *cbase, uint64_t n, uint32_t *flags, uint2 *hits, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    mark_buckets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(skeys, svals, mz_rid, mz_rp, cbase, n, flags, hits);
    if (lc) lc->total++;
}

// bid = exclusive scan of flags; bucket of element i = bid[i] + flags[i] - 1
__global__ void bucket_starts_kernel(const uint32_t *__restrict__ flags, const uint32_t *__restrict__ bid, uint64_t n,
                                     uint32_t *__restrict__ bstart) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = bid[i] + flags[i] - 1u;
    if (flags[i]) bstart[b] = (uint32_t)i;
    if (i == n - 1) bstart[b + 1] = (uint32_t)n;
}
void launch_bucket_starts(const uint32_t *flags, const uint32_t *bid, uint64_t n, uint32_t *bstart, cudaStream_t s,
                          LaunchCounter *lc) {
    if (n == 0) return;
    bucket_starts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(flags, bid, n, bstart);
    if (lc) lc->total++;
}

__global__ void bucket_ranges_kernel(const uint32_t *__restrict__ flags, const uint32_t *__restrict__ bid,
                                     const uint32_t *__restrict__ svals, const uint32_t *__restrict__ bstart, uint64_t n,
                                     uint2 *__restrict__ range) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = bid[i] + flags[i] - 1u;
    range[svals[i]] = make_uint2(bstart[b], bstart[b + 1]);
}
void launch_bucket_ranges(const uint32_t *flags, const uint32_t *bid, const uint32_t *svals, const uint32_t *bstart,
                          uint64_t n, uint2 *range, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    bucket_ranges_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(flags, bid, svals, bstart, n, range);
    if (lc) lc->total++;
}

// ------------------------------------------------------------------------------------------------
// table operations (generic pointers: shared memory in the fast kernel, HBM in the overflow kernel)

__device__ __forceinline__ uint32_t slot_hash(uint64_t key, uint32_t mask) {
    return (uint32_t)((key * 0x9E3779B97F4A7C15ULL) >> 40) & mask;
}

// returns false when no slot could be claimed
__device__ __forceinline__ bool table_insert(unsigned long long *keys, uint32_t *mn, uint32_t *mx, uint32_t mask, uint64_t key,
                                             uint32_t qp) {
    uint32_t slot = slot_hash(key, mask);
    for (uint32_t probe = 0; probe <= mask; ++probe) {
        const unsigned long long prev = atomicCAS(&keys[slot], HLO_EMPTY_KEY, (unsigned long long)key);
        if (prev == HLO_EMPTY_KEY || prev == key) {
            atomicMin(&mn[slot], qp);
            atomicMax(&mx[slot], qp);
            return true;
        }
        slot = (slot + 1) & mask;
    }
    return false;
}

__device__ __forceinline__ hlo_cand make_cand(uint32_t q, uint64_t key, uint32_t mn, uint32_t mx, int k) {
    hlo_cand c;
    c.q = q;
    c.t = (uint32_t)(key >> 32);
    const uint32_t lo = (uint32_t)key;
    c.diag = (int32_t)(lo & 0x7fffffffu) - HLO_DIAG_BIAS;
    c.strand_anchor = HLO_CAND_PACK(lo >> 31, mn - (uint32_t)k + 1u, mx + 1u);
    return c;
}

// fold the hits of query read q into the table; NT cooperating threads, this thread is `tid`.
// returns (per thread) the number of hits that survived -X, and sets *ovf on table overflow.
template <int NT_IS_WARP>
__device__ __forceinline__ uint32_t fold_read(const PairgenArgs &a, uint32_t q, unsigned long long *keys, uint32_t *mn,
                                              uint32_t *mx, uint32_t mask, uint32_t tid, uint32_t nt, volatile uint32_t *ovf) {
    const uint32_t ql = a.reads.len[q] & HLO_LEN_MASK;
    const uint32_t qid = a.reads.ids[q];
    const uint32_t m0 = a.mz_off[q], m1 = a.mz_off[q + 1];
    uint32_t nh = 0;
    for (uint32_t m = m0; m < m1; ++m) {
        const uint2 rg = a.range[m];
        if ((int)(rg.y - rg.x) > a.mid_occ) continue;
        if (*ovf) break;
        const uint32_t qrp = a.mz_rp[m];
        const uint32_t qpos = qrp >> 1;
        for (uint32_t i = rg.x + tid; i < rg.y; i += nt) {
            const uint2 h = a.hits[i];
            const uint32_t t = h.x;
            if (t == q) continue;                                // -X: no self hits
            if (!name_less(qid, a.reads.ids[t])) continue;       // -X: query = the smaller name
            ++nh;
            const uint32_t strand = (h.y ^ qrp) & 1u;
            const int32_t tp = (int32_t)(h.y >> 1);
            const int32_t qp = strand ? (int32_t)ql - ((int32_t)qpos + 1 - a.k) - 1 : (int32_t)qpos;
            const int32_t diag = tp - qp;
            const uint64_t key = (uint64_t)t << 32 | (uint64_t)(strand << 31) | (uint32_t)(diag + HLO_DIAG_BIAS);
            if (!table_insert(keys, mn, mx, mask, key, (uint32_t)qp)) *ovf = 1u;
        }
    }
    return nh;
}

__global__ void __launch_bounds__(128) pairgen_kernel(PairgenArgs a, int slots_log2) {
    extern __shared__ unsigned long long smem_u64[];
    const uint32_t S = 1u << slots_log2, mask = S - 1u;
    const uint32_t wp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    unsigned long long *keys = smem_u64 + (size_t)wp * S;
    uint32_t *mn = (uint32_t *)(smem_u64 + (size_t)nw * S) + (size_t)wp * 2 * S;
    uint32_t *mx = mn + S;
    __shared__ uint32_t ovf_flag[8];
    const unsigned lt_mask = (1u << lane) - 1u;
    for (;;) {
        unsigned long long qq = 0;
        if (lane == 0) qq = atomicAdd(&a.counters[3], 1ULL);
        qq = __shfl_sync(0xffffffffu, qq, 0);
        if (qq >= a.reads.n_reads) break;
        const uint32_t q = (uint32_t)qq;
        if (a.mz_off[q] == a.mz_off[q + 1]) continue;
        for (uint32_t i = lane; i < S; i += 32) keys[i] = HLO_EMPTY_KEY, mn[i] = 0xffffffffu, mx[i] = 0u;
        if (lane == 0) ovf_flag[wp] = 0u;
        __syncwarp();
        uint32_t nh = fold_read<1>(a, q, keys, mn, mx, mask, lane, 32, &ovf_flag[wp]);
        __syncwarp();
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nh += __shfl_xor_sync(0xffffffffu, nh, d);
        if (ovf_flag[wp]) {  // redo this read exactly in the overflow kernel (its hits are counted there)
            if (lane == 0) {
                const unsigned long long p = atomicAdd(&a.counters[2], 1ULL);
                if (p < a.overflow_cap) a.overflow_reads[p] = q;
            }
            __syncwarp();
            continue;
        }
        // count, reserve, write
        uint32_t cnt = 0;
        for (uint32_t i = lane; i < S; i += 32) cnt += (keys[i] != HLO_EMPTY_KEY && (int)(mx[i] - mn[i]) >= a.gap_min);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        unsigned long long base = 0;
        if (lane == 0) {
            if (cnt) base = atomicAdd(&a.counters[0], (unsigned long long)cnt);
            if (nh) atomicAdd(&a.counters[1], (unsigned long long)nh);
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (cnt) {
            uint32_t run = 0;
            for (uint32_t i0 = 0; i0 < S; i0 += 32) {
                const uint32_t i = i0 + lane;
                const unsigned long long key = keys[i];
                const uint32_t lo = mn[i], hi = mx[i];
                const bool ok = key != HLO_EMPTY_KEY && (int)(hi - lo) >= a.gap_min;
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    const unsigned long long p = base + run + __popc(bal & lt_mask);
                    if (p < a.cand_cap) a.cands[p] = make_cand(q, key, lo, hi, a.k);
                }
                run += __popc(bal);
            }
        }
        __syncwarp();
    }
}

void launch_pairgen(const PairgenArgs &a, int slots_log2, cudaStream_t s, LaunchCounter *lc) {
    if (a.reads.n_reads == 0) return;
    const int threads = 128, nw = threads / 32;
    const size_t smem = (size_t)nw * ((size_t)16 << slots_log2);
    static int configured_for = -1;
    if (configured_for != slots_log2) {
        cudaFuncSetAttribute(pairgen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
        configured_for = slots_log2;
    }
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pairgen_kernel, threads, smem);
    if (per_sm < 1) per_sm = 1;
    uint64_t want = ((uint64_t)a.reads.n_reads + nw - 1) / nw;
    uint64_t grid = (uint64_t)sms * per_sm;
    if (grid > want) grid = want;
    pairgen_kernel<<<(unsigned)grid, threads, smem, s>>>(a, slots_log2);
    if (lc) lc->total++;
}

// ------------------------------------------------------------------------------------------------
// overflow path: one block per overflowed read, table in HBM sized from the read's bucket sizes

__global__ void overflow_hit_counts_kernel(PairgenArgs a, uint32_t n_overflow, uint32_t *__restrict__ counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_overflow) return;
    const uint32_t q = a.overflow_reads[i];
    uint32_t c = 0;
    for (uint32_t m = a.mz_off[q]; m < a.mz_off[q + 1]; ++m) {
        const uint2 rg = a.range[m];
        if ((int)(rg.y - rg.x) <= a.mid_occ) c += rg.y - rg.x;
    }
    counts[i] = c;
}
void launch_overflow_hit_counts(const PairgenArgs &a, uint32_t n_overflow, uint32_t *counts, cudaStream_t s, LaunchCounter *lc) {
    if (n_overflow == 0) return;
    overflow_hit_counts_kernel<<<(n_overflow + 127) / 128, 128, 0, s>>>(a, n_overflow, counts);
    if (lc) lc->total++;
}

__global__ void __launch_bounds__(256) pairgen_overflow_kernel(PairgenArgs a, uint32_t n_overflow, unsigned long long *gkeys,
                                                               uint32_t *gmn, uint32_t *gmx, const uint32_t *__restrict__ goff) {
    __shared__ uint32_t ovf;
    __shared__ unsigned long long nh_tot;
    const uint32_t i = blockIdx.x;
    if (i >= n_overflow) return;
    const uint32_t q = a.overflow_reads[i];
    const uint32_t o0 = goff[i], S = goff[i + 1] - o0, mask = S - 1u;  // S is a power of two
    unsigned long long *keys = gkeys + o0;
    uint32_t *mn = gmn + o0, *mx = gmx + o0;
    if (threadIdx.x == 0) ovf = 0u, nh_tot = 0ULL;
    __syncthreads();
    const uint32_t nh = fold_read<0>(a, q, keys, mn, mx, mask, threadIdx.x, blockDim.x, &ovf);
    if (nh) atomicAdd(&nh_tot, (unsigned long long)nh);
    __syncthreads();
    if (threadIdx.x == 0 && nh_tot) atomicAdd(&a.counters[1], nh_tot);
    for (uint32_t j = threadIdx.x; j < S; j += blockDim.x) {
        const unsigned long long key = keys[j];
        const uint32_t lo = mn[j], hi = mx[j];
        if (key != HLO_EMPTY_KEY && (int)(hi - lo) >= a.gap_min) {
            const unsigned long long p = atomicAdd(&a.counters[0], 1ULL);
            if (p < a.cand_cap) a.cands[p] = make_cand(q, key, lo, hi, a.k);
        }
    }
}

void launch_pairgen_overflow(const PairgenArgs &a, uint32_t n_overflow, unsigned long long *gtable_keys, uint32_t *gtable_mn,
                             uint32_t *gtable_mx, const uint32_t *gtable_off, cudaStream_t s, LaunchCounter *lc) {
    if (n_overflow == 0) return;
    pairgen_overflow_kernel<<<n_overflow, 256, 0, s>>>(a, n_overflow, gtable_keys, gtable_mn, gtable_mx, gtable_off);
    if (lc) lc->total++;
}

}  // namespace hlo
