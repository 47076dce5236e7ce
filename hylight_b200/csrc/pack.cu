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

void launch_pack(const uint8_t *ascii, const uint64_t *base_off, DevReads r, uint32_t *amb_count, cudaStream_t s,
                 LaunchCounter *lc) {
    const uint64_t total = (uint64_t)r.n_reads * r.stride_w;
    if (total == 0) return;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    pack_kernel<<<blocks, 256, 0, s>>>(ascii, base_off, r, amb_count);
    if (lc) lc->total++;
}

}  // namespace hlo
