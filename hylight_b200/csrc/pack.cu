// pack.cu — ASCII reads -> 2-bit packed rows, both strands, plus ambiguity planes.
//
// Replaces, for the hot path, what fastq2fasta.py:32-40 + minimap2's seq_nt4_table encoding do on the
// CPU (A/a=0 C/c=1 G/g=2 T/t/U/u=3, anything else ambiguous).  HBM-bound byte work:
// algorithmic bytes per read = L (ASCII in) + 2 strands * 2 planes * ceil(L/4) (out).
#include "kernels.cuh"

namespace hlo {

__device__ __forceinline__ uint32_t nt4_code(uint32_t ch) {
    const uint32_t u = ch & 0xDFu;  // fold case
    return u == 'A' ? 0u : u == 'C' ? 1u : u == 'G' ? 2u : (u == 'T' || u == 'U') ? 3u : 4u;
}

// One thread per (read, word).  Thread order is word-major inside a read so that a warp reads a
// contiguous stretch of ASCII and writes contiguous packed words.
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ ascii, const uint64_t *__restrict__ base_off,
                                                   DevReads r, uint32_t *__restrict__ amb_count) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t sw = r.stride_w;
    const uint64_t total = (uint64_t)r.n_reads * sw;
    if (gid >= total) return;
    const uint32_t read = (uint32_t)(gid / sw), wi = (uint32_t)(gid % sw);
    const uint64_t b0 = base_off[read];
    const int L = (int)(base_off[read + 1] - b0);
    const int p0 = (int)wi * 16;
    uint32_t fw = 0, fa = 0, rw = 0, ra = 0;
    if (p0 < L) {
        const uint8_t *src = ascii + b0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int p = p0 + i;
            if (p < L) {
                const uint32_t c = nt4_code(__ldg(src + p));
                fw |= (c & 3u) << (2 * i);
                fa |= (c >> 2) << (2 * i);
                const uint32_t d = nt4_code(__ldg(src + (L - 1 - p)));  // base p of the reverse complement
                rw |= (d > 3u ? 0u : 3u - d) << (2 * i);
                ra |= (d >> 2) << (2 * i);
            }
        }
        // ambiguous bases are stored as code 0 in the base plane
        fw &= ~(fa * 3u);
    }
    const uint64_t plane = (uint64_t)r.n_reads * sw;
    r.bases[gid] = fw;
    r.bases[plane + gid] = rw;
    if (r.amb) {
        r.amb[gid] = fa;
        r.amb[plane + gid] = ra;
    }
    if (fa) {
        atomicOr(&r.len[read], HLO_AMB_FLAG);
        atomicAdd(amb_count, (uint32_t)__popc(fa));
    }
}

// Per-read metadata for the plain case (names 0..N-1 in file order, rename_fas.py:48-57), derived on the
// device from the two offset arrays: length, cluster, name, numeric rank, and the name-order permutation.
//
// perm[cbase + x] = cbase + (the x-th integer of [0, N) in strcmp order of the decimal spellings), the order
// minimap2 -X compares names in ("0" < "1" < "10" < "100" < ... < "11" < ... < "2"): a preorder walk of the
// decimal trie, done per thread by skipping whole subtrees.
__device__ __forceinline__ uint32_t dec_subtree(uint64_t p, uint64_t N) {  // integers in [0, N) spelled with prefix p >= 1
    uint64_t c = 0, lo = p, span = 1;
    while (lo < N) {
        const uint64_t hi = lo + span < N ? lo + span : N;
        c += hi - lo;
        lo *= 10, span *= 10;
    }
    return (uint32_t)c;
}
__device__ __forceinline__ uint32_t dec_order_kth(uint32_t N, uint32_t x) {
    if (x == 0) return 0;  // "0" sorts first and has no children
    --x;
    uint64_t p = 1;
    for (;;) {
        const uint32_t c = dec_subtree(p, N);
        if (x < c) {
            if (x == 0) return (uint32_t)p;
            --x, p *= 10;  // into the subtree: first child
        } else {
            x -= c, ++p;   // next sibling
        }
    }
}

__global__ void __launch_bounds__(256) read_meta_kernel(const uint64_t *__restrict__ base_off, const uint32_t *__restrict__ cluster_off,
                                                        DevReads r) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r.n_reads) return;
    // cluster of read i: last c with cluster_off[c] <= i
    uint32_t lo = 0, hi = r.n_clusters;  // invariant: cluster_off[lo] <= i < cluster_off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(cluster_off + mid) <= i) lo = mid; else hi = mid;
    }
    const uint32_t cb = __ldg(cluster_off + lo), N = __ldg(cluster_off + lo + 1) - cb;
    r.len[i] = (uint32_t)(base_off[i + 1] - base_off[i]);
    r.cbase[i] = cb;
    r.cidx[i] = lo;
    r.ids[i] = i - cb;
    r.nrank[i] = i;
    r.perm[i] = cb + dec_order_kth(N, i - cb);
}

void launch_read_meta(const uint64_t *base_off, const uint32_t *cluster_off, DevReads r, cudaStream_t s, LaunchCounter *lc) {
    if (r.n_reads == 0) return;
    read_meta_kernel<<<(r.n_reads + 255) / 256, 256, 0, s>>>(base_off, cluster_off, r);
    if (lc) lc->total++;
}

void launch_pack(const uint8_t *ascii, const uint64_t *base_off, DevReads r, uint32_t *amb_count, cudaStream_t s,
                 LaunchCounter *lc) {
    const uint64_t total = (uint64_t)r.n_reads * r.stride_w;
    if (total == 0) return;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    pack_kernel<<<blocks, 256, 0, s>>>(ascii, base_off, r, amb_count);
    if (lc) lc->total++;
}

}  // namespace hlo
