// sketch.cu — (w,k) minimizer sketch, bit-exact with minimap2's mm_sketch() (non-HPC branch).
//
// minimap2 is not in /root/reference (un-vendored dependency, call site polyte.tune_params.py:510,
// "-k 21 -w 11"); the algorithm restated here is the one in oracle/hlo_oracle.c:orc_sketch, and the
// parity tests compare the two tuple for tuple: x = hash64(canonical k-mer) << 8 | span,
// y = rid << 32 | last_pos << 1 | strand.
//
// One thread per read (taken in name order, see below) runs the sequential window automaton exactly (ties, the first-window special
// case, ambiguous-base resets and the end-of-read flush are all order dependent), reading the read's
// packed row; a count pass sizes the output, a write pass fills it, so the output is in minimap2's
// emission order.  Integer-pipe bound: ~150 ALU ops per base; algorithmic bytes per read =
// ceil(L/4) in + 16 B per minimizer out.
#include "kernels.cuh"

namespace hlo {

#define HLO_MAX_W HLO_MAX_W_PUBLIC

__device__ __forceinline__ uint64_t mm_hash64(uint64_t key, uint64_t mask) {
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

struct MzOut {
    uint64_t *x;
    uint32_t *rp;
    uint32_t *rid;
};

template <bool WRITE>
__device__ __forceinline__ void mz_push(uint32_t &n, uint64_t x, uint32_t rp, uint32_t rid, uint32_t base, const MzOut &o) {
    if (WRITE) {
        o.x[base + n] = x;
        o.rp[base + n] = rp;
        o.rid[base + n] = rid;
    }
    ++n;
}

template <bool WRITE>
__global__ void __launch_bounds__(128) sketch_kernel(DevReads r, int w, int k, uint32_t *__restrict__ count,
                                                     const uint32_t *__restrict__ off, MzOut o) {
    // thread x sketches read perm[x]: the minimizer arrays are laid out in the strcmp order of the read
    // names inside each cluster, so that after the stable hash sort every seed bucket is ordered the way
    // minimap2's -X rule compares names (pairgen.cu)
    const uint32_t xi = blockIdx.x * blockDim.x + threadIdx.x;
    if (xi >= r.n_reads) return;
    const uint32_t read = __ldg(r.perm + xi);
    const uint32_t lenf = r.len[read];
    const int len = (int)(lenf & HLO_LEN_MASK);
    const bool has_amb = (lenf & HLO_AMB_FLAG) != 0;
    const uint32_t *row = r.bases + (uint64_t)read * r.stride_w;
    const uint32_t *arow = r.amb + (uint64_t)read * r.stride_w;
    const uint64_t shift1 = 2 * (uint64_t)(k - 1), mask = (1ULL << 2 * k) - 1;
    const uint64_t INF = ~0ULL;
    uint64_t kmer0 = 0, kmer1 = 0;
    uint64_t buf_x[HLO_MAX_W];
    uint32_t buf_y[HLO_MAX_W];
    uint64_t min_x = INF;
    uint32_t min_y = 0xffffffffu;
    const uint32_t base = WRITE ? off[xi] : 0u;
    uint32_t n = 0;
    int l = 0, buf_pos = 0, min_pos = 0;
#pragma unroll
    for (int j = 0; j < HLO_MAX_W; ++j) buf_x[j] = INF, buf_y[j] = 0xffffffffu;
    uint32_t word = 0, aword = 0;
    for (int i = 0; i < len; ++i) {
        if ((i & 15) == 0) {
            word = __ldg(row + (i >> 4));
            aword = has_amb ? __ldg(arow + (i >> 4)) : 0u;
        }
        const uint32_t c = word & 3u, amb = aword & 1u;
        word >>= 2, aword >>= 2;
        uint64_t info_x = INF;
        uint32_t info_y = 0xffffffffu;
        if (!amb) {
            const int kmer_span = l + 1 < k ? l + 1 : k;
            kmer0 = (kmer0 << 2 | (uint64_t)c) & mask;
            kmer1 = (kmer1 >> 2) | (3ULL ^ (uint64_t)c) << shift1;
            if (kmer0 == kmer1) continue;  // strand-symmetric k-mer
            const uint32_t z = kmer0 < kmer1 ? 0u : 1u;
            ++l;
            if (l >= k) {
                info_x = mm_hash64(z ? kmer1 : kmer0, mask) << 8 | (uint64_t)kmer_span;
                info_y = (uint32_t)i << 1 | z;
            }
        } else
            l = 0;
        buf_x[buf_pos] = info_x, buf_y[buf_pos] = info_y;
        if (l == w + k - 1 && min_x != INF) {  // first full window: identical k-mers were not stored yet
            for (int j = buf_pos + 1; j < w; ++j)
                if (min_x == buf_x[j] && buf_y[j] != min_y) mz_push<WRITE>(n, buf_x[j], buf_y[j], read, base, o);
            for (int j = 0; j < buf_pos; ++j)
                if (min_x == buf_x[j] && buf_y[j] != min_y) mz_push<WRITE>(n, buf_x[j], buf_y[j], read, base, o);
        }
        if (info_x <= min_x) {  // a new minimum; write the old one
            if (l >= w + k && min_x != INF) mz_push<WRITE>(n, min_x, min_y, read, base, o);
            min_x = info_x, min_y = info_y, min_pos = buf_pos;
        } else if (buf_pos == min_pos) {  // the old minimum left the window
            if (l >= w + k - 1 && min_x != INF) mz_push<WRITE>(n, min_x, min_y, read, base, o);
            min_x = INF;
            for (int j = buf_pos + 1; j < w; ++j)
                if (min_x >= buf_x[j]) min_x = buf_x[j], min_y = buf_y[j], min_pos = j;  // >= : rightmost smallest
            for (int j = 0; j <= buf_pos; ++j)
                if (min_x >= buf_x[j]) min_x = buf_x[j], min_y = buf_y[j], min_pos = j;
            if (l >= w + k - 1 && min_x != INF) {
                for (int j = buf_pos + 1; j < w; ++j)
                    if (min_x == buf_x[j] && min_y != buf_y[j]) mz_push<WRITE>(n, buf_x[j], buf_y[j], read, base, o);
                for (int j = 0; j <= buf_pos; ++j)
                    if (min_x == buf_x[j] && min_y != buf_y[j]) mz_push<WRITE>(n, buf_x[j], buf_y[j], read, base, o);
            }
        }
        if (++buf_pos == w) buf_pos = 0;
    }
    if (min_x != INF) mz_push<WRITE>(n, min_x, min_y, read, base, o);
    if (!WRITE) count[xi] = n;
}

void launch_sketch_count(const DevReads &r, int w, int k, uint32_t *count, cudaStream_t s, LaunchCounter *lc) {
    if (r.n_reads == 0) return;
    MzOut o{nullptr, nullptr, nullptr};
    sketch_kernel<false><<<(r.n_reads + 127) / 128, 128, 0, s>>>(r, w, k, count, nullptr, o);
    if (lc) lc->total++;
}

void launch_sketch_write(const DevReads &r, int w, int k, const uint32_t *off, uint64_t *mz_x, uint32_t *mz_rp,
                         uint32_t *mz_rid, cudaStream_t s, LaunchCounter *lc) {
    if (r.n_reads == 0) return;
    MzOut o{mz_x, mz_rp, mz_rid};
    sketch_kernel<true><<<(r.n_reads + 127) / 128, 128, 0, s>>>(r, w, k, nullptr, off, o);
    if (lc) lc->total++;
}

__global__ void mz_assemble_kernel(const uint64_t *__restrict__ x, const uint32_t *__restrict__ rp,
                                   const uint32_t *__restrict__ rid, uint64_t n, hlo_mz *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    hlo_mz m;
    m.x = x[i];
    m.y = (uint64_t)rid[i] << 32 | rp[i];
    out[i] = m;
}

void launch_mz_assemble(const uint64_t *mz_x, const uint32_t *mz_rp, const uint32_t *mz_rid, uint64_t n, hlo_mz *out,
                        cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    mz_assemble_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mz_x, mz_rp, mz_rid, n, out);
    if (lc) lc->total++;
}

}  // namespace hlo
