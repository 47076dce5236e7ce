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

// ---- generic kernel: one thread per (read, word), byte loads; any read length -------------------------
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

// ---- tile kernel (reads up to PK_MAXLEN bases): a block stages the contiguous ASCII of PK_TR reads in shared
// memory with 16-byte loads, converts four characters at a time (SWAR), and derives the reverse-complement
// rows from the packed forward rows (bit reversal + complement) instead of reading the text twice.
constexpr int PK_TR = 32;        // reads per block
constexpr int PK_MAXLEN = 512;   // longest read the tile kernel takes
constexpr int PK_THREADS = 256;

// 0x80 in every byte of v that is zero (exact)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t v) {
    return ~(((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v | 0x7f7f7f7fu);
}
// four characters -> (8 bits of 2-bit codes, 8 bits of 01 ambiguity flags), character 0 in the low bits
__device__ __forceinline__ void pack4(uint32_t w, uint32_t &codes8, uint32_t &amb8) {
    const uint32_t v = w | 0x20202020u;  // fold case
    const uint32_t ok = zero_bytes(v ^ 0x61616161u) | zero_bytes(v ^ 0x63636363u) | zero_bytes(v ^ 0x67676767u) |
                        zero_bytes(v ^ 0x74747474u) | zero_bytes(v ^ 0x75757575u);  // a c g t u
    const uint32_t bad = (~ok & 0x80808080u) >> 7;                       // 1 per ambiguous character
    uint32_t c = ((w >> 1) ^ (w >> 2)) & 0x03030303u;                    // A=0 C=1 G=2 T/U=3
    c &= ~(bad * 3u);                                                    // ambiguous -> 0
    codes8 = (c * 0x01041040u) >> 24;                                    // gather the four 2-bit fields
    amb8 = (bad * 0x01041040u) >> 24;
}
// reverse the order of the 16 two-bit fields of a word
__device__ __forceinline__ uint32_t rev_pairs(uint32_t v) {
    const uint32_t r = __brev(v);
    return ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
}

__global__ void __launch_bounds__(PK_THREADS) pack_tile_kernel(const uint8_t *__restrict__ ascii, const uint64_t *__restrict__ base_off,
                                                               DevReads r, uint32_t *__restrict__ amb_count) {
    extern __shared__ uint32_t pk_smem[];
    const uint32_t sw = r.stride_w;
    uint32_t *s_fw = pk_smem;                      // [PK_TR][sw + 1] forward base words (+1 zero word)
    uint32_t *s_fa = s_fw + PK_TR * (sw + 1);      // [PK_TR][sw + 1] forward ambiguity words
    uint32_t *s_asc = s_fa + PK_TR * (sw + 1);     // staged text
    __shared__ uint64_t s_off[PK_TR + 1];
    __shared__ uint32_t s_namb;
    const uint32_t r0 = blockIdx.x * PK_TR;
    const uint32_t nr = min((uint32_t)PK_TR, r.n_reads - r0);
    if (threadIdx.x <= nr) s_off[threadIdx.x] = base_off[r0 + threadIdx.x];
    if (threadIdx.x == 0) s_namb = 0;
    __syncthreads();
    const uint64_t t0 = s_off[0], t1 = s_off[nr];
    const uint64_t a0 = t0 & ~15ull;                // 16-byte aligned start (the text buffer is 256-byte aligned)
    const uint32_t nvec = (uint32_t)((t1 - a0 + 15) >> 4);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(ascii + a0);
        uint4 *dst = reinterpret_cast<uint4 *>(s_asc);
        for (uint32_t v = threadIdx.x; v < nvec; v += PK_THREADS) dst[v] = __ldg(src + v);
    }
    __syncthreads();
    const uint64_t plane = (uint64_t)r.n_reads * sw;
    // forward rows
    for (uint32_t slot = threadIdx.x; slot < nr * (sw + 1); slot += PK_THREADS) {
        const uint32_t lr = slot / (sw + 1), wi = slot % (sw + 1);
        uint32_t fw = 0, fa = 0;
        if (wi < sw) {
            const uint32_t L = (uint32_t)(s_off[lr + 1] - s_off[lr]);
            const uint32_t p0 = wi * 16;
            if (p0 < L) {
                const uint32_t boff = (uint32_t)(s_off[lr] - a0) + p0;  // byte offset of the word's first character
                const uint32_t wq = boff >> 2, sh = (boff & 3u) * 8u;
                uint32_t prev = s_asc[wq];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t next = s_asc[wq + q + 1];
                    const uint32_t w = __funnelshift_r(prev, next, sh);  // characters 4q .. 4q+3 of the word
                    prev = next;
                    uint32_t c8, a8;
                    pack4(w, c8, a8);
                    fw |= c8 << (8 * q), fa |= a8 << (8 * q);
                }
                const uint32_t n = L - p0;  // characters of this word inside the read
                if (n < 16) {
                    const uint32_t m = (1u << (2 * n)) - 1u;
                    fw &= m, fa &= m;
                }
            }
            const uint64_t g = (uint64_t)(r0 + lr) * sw + wi;
            r.bases[g] = fw;
            r.amb[g] = fa;
            if (fa) {
                atomicOr(&r.len[r0 + lr], HLO_AMB_FLAG);
                atomicAdd(&s_namb, (uint32_t)__popc(fa));
            }
        }
        s_fw[slot] = fw, s_fa[slot] = fa;  // wi == sw: the zero word behind the row
    }
    __syncthreads();
    // reverse-complement rows from the packed forward rows
    for (uint32_t slot = threadIdx.x; slot < nr * sw; slot += PK_THREADS) {
        const uint32_t lr = slot / sw, wi = slot % sw;
        const int L = (int)(s_off[lr + 1] - s_off[lr]);
        uint32_t rw = 0, ra = 0;
        const int n = L - (int)wi * 16;  // bases of this word inside the read
        if (n > 0) {
            // RC bases 16wi .. 16wi+15 = forward bases L-16wi-16 .. L-16wi-1, reversed and complemented
            const int start = L - (int)wi * 16 - 16;
            const int j = start >> 4;               // floor; -1 when the window starts before base 0
            const uint32_t sh = (uint32_t)(start & 15) * 2u;
            const uint32_t *fwr = s_fw + lr * (sw + 1), *far_ = s_fa + lr * (sw + 1);
            const uint32_t lo = j >= 0 ? fwr[j] : 0u, hi = fwr[j + 1];
            const uint32_t alo = j >= 0 ? far_[j] : 0u, ahi = far_[j + 1];
            const uint32_t v = __funnelshift_r(lo, hi, sh), av = __funnelshift_r(alo, ahi, sh);
            const uint32_t m = n >= 16 ? 0xffffffffu : ((1u << (2 * n)) - 1u);
            ra = rev_pairs(av) & m;
            rw = ~rev_pairs(v) & m & ~(ra * 3u);
        }
        const uint64_t g = plane + (uint64_t)(r0 + lr) * sw + wi;
        r.bases[g] = rw;
        r.amb[g] = ra;
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_namb) atomicAdd(amb_count, s_namb);
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
    r.meta[i] = make_uint4((uint32_t)(base_off[i + 1] - base_off[i]), i - cb, lo, cb);
    r.perm[i] = cb + dec_order_kth(N, i - cb);
}

void launch_read_meta(const uint64_t *base_off, const uint32_t *cluster_off, DevReads r, cudaStream_t s, LaunchCounter *lc) {
    if (r.n_reads == 0) return;
    read_meta_kernel<<<(r.n_reads + 255) / 256, 256, 0, s>>>(base_off, cluster_off, r);
    if (lc) lc->total++;
}

// ---- packed input (hlo_batch.packed): the reads arrive 2-bit packed, word aligned, back to back --------------------
// words per read (for the scan that gives every read its first word) and the plain per-read metadata
__global__ void __launch_bounds__(256) packed_meta_kernel(const uint32_t *__restrict__ read_len, const uint32_t *__restrict__ cluster_off,
                                                          DevReads r, uint32_t *__restrict__ wcount) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > r.n_reads) return;
    if (i == r.n_reads) {
        wcount[i] = 0;
        return;
    }
    uint32_t lo = 0, hi = r.n_clusters;  // invariant: cluster_off[lo] <= i < cluster_off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(cluster_off + mid) <= i) lo = mid; else hi = mid;
    }
    const uint32_t cb = __ldg(cluster_off + lo), N = __ldg(cluster_off + lo + 1) - cb, L = read_len[i];
    wcount[i] = (L + 15u) >> 4;
    r.len[i] = L;
    r.cbase[i] = cb;
    r.cidx[i] = lo;
    r.ids[i] = i - cb;
    r.nrank[i] = i;
    r.meta[i] = make_uint4(L, i - cb, lo, cb);
    r.perm[i] = cb + dec_order_kth(N, i - cb);
}

// rows of both strands from the packed stream: the forward row is a copy, the reverse-complement row is derived from
// two forward words (bit reversal + complement), as pack_tile_kernel does from its staged rows
__global__ void __launch_bounds__(256) unpack2b_kernel(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ woff, DevReads r) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t sw = r.stride_w;
    if (gid >= (uint64_t)r.n_reads * sw) return;
    const uint32_t read = (uint32_t)(gid / sw), wi = (uint32_t)(gid % sw);
    const int L = (int)(r.len[read] & HLO_LEN_MASK);
    const uint32_t nw = ((uint32_t)L + 15u) >> 4;
    const uint32_t *src = packed + woff[read];
    uint32_t fw = 0, rw = 0;
    const int n = L - (int)wi * 16;  // bases of this word inside the read
    if (n > 0) {
        const uint32_t m = n >= 16 ? 0xffffffffu : ((1u << (2 * n)) - 1u);
        fw = __ldg(src + wi) & m;
        // RC bases 16wi .. 16wi+15 = forward bases L-16wi-16 .. L-16wi-1, reversed and complemented
        const int start = L - (int)wi * 16 - 16;
        const int j = start >> 4;  // floor; -1 when the window starts before base 0
        const uint32_t sh = (uint32_t)(start & 15) * 2u;
        const uint32_t lo = j >= 0 ? __ldg(src + j) : 0u, hi = (uint32_t)(j + 1) < nw ? __ldg(src + j + 1) : 0u;
        rw = ~rev_pairs(__funnelshift_r(lo, hi, sh)) & m;
    }
    const uint64_t plane = (uint64_t)r.n_reads * sw;
    r.bases[gid] = fw, r.bases[plane + gid] = rw;
    r.amb[gid] = 0u, r.amb[plane + gid] = 0u;
}

// the listed ambiguous bases: flag bits in both ambiguity planes, code 0 in both base planes
__global__ void __launch_bounds__(256) packed_amb_kernel(const unsigned long long *__restrict__ amb, uint64_t n_amb, DevReads r,
                                                         uint32_t *__restrict__ amb_count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_amb) return;
    const uint32_t read = (uint32_t)(amb[i] >> 32), pos = (uint32_t)amb[i];
    if (read >= r.n_reads) return;
    const uint32_t L = r.len[read] & HLO_LEN_MASK;
    if (pos >= L) return;
    const uint64_t plane = (uint64_t)r.n_reads * r.stride_w, row = (uint64_t)read * r.stride_w;
    const uint32_t rp = L - 1u - pos;
    atomicOr(&r.amb[row + (pos >> 4)], 1u << ((pos & 15u) * 2u));
    atomicOr(&r.amb[plane + row + (rp >> 4)], 1u << ((rp & 15u) * 2u));
    atomicAnd(&r.bases[row + (pos >> 4)], ~(3u << ((pos & 15u) * 2u)));
    atomicAnd(&r.bases[plane + row + (rp >> 4)], ~(3u << ((rp & 15u) * 2u)));
    atomicOr(&r.len[read], HLO_AMB_FLAG);
    atomicAdd(amb_count, 1u);
}

void launch_packed_meta(const uint32_t *read_len, const uint32_t *cluster_off, DevReads r, uint32_t *wcount, cudaStream_t s, LaunchCounter *lc) {
    packed_meta_kernel<<<(r.n_reads + 1 + 255) / 256, 256, 0, s>>>(read_len, cluster_off, r, wcount);
    if (lc) lc->total++;
}
void launch_unpack2b(const uint32_t *packed, const uint32_t *woff, const unsigned long long *amb, uint64_t n_amb, DevReads r,
                     uint32_t *amb_count, cudaStream_t s, LaunchCounter *lc) {
    const uint64_t total = (uint64_t)r.n_reads * r.stride_w;
    if (total == 0) return;
    unpack2b_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(packed, woff, r);
    if (lc) lc->total++;
    if (n_amb) {
        packed_amb_kernel<<<(unsigned)((n_amb + 255) / 256), 256, 0, s>>>(amb, n_amb, r, amb_count);
        if (lc) lc->total++;
    }
}

void launch_pack(const uint8_t *ascii, const uint64_t *base_off, DevReads r, uint32_t *amb_count, cudaStream_t s,
                 LaunchCounter *lc) {
    const uint64_t total = (uint64_t)r.n_reads * r.stride_w;
    if (total == 0) return;
    if (r.max_len <= PK_MAXLEN) {
        // staged text: PK_TR reads of at most max_len characters, plus alignment slack on both sides
        const size_t smem = ((size_t)2 * PK_TR * (r.stride_w + 1)) * 4 + (((size_t)PK_TR * r.max_len + 15 + 16 + 15) & ~15ull) + 16;
        cudaFuncSetAttribute(pack_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        pack_tile_kernel<<<(r.n_reads + PK_TR - 1) / PK_TR, PK_THREADS, smem, s>>>(ascii, base_off, r, amb_count);
    } else {
        const unsigned blocks = (unsigned)((total + 255) / 256);
        pack_kernel<<<blocks, 256, 0, s>>>(ascii, base_off, r, amb_count);
    }
    if (lc) lc->total++;
}

}  // namespace hlo
