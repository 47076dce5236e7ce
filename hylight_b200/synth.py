"""Seeded synthetic strain-mixture clusters (BASELINE.json configs 2-5).

One cluster mimics what get_fq_cluster.py:22-57 hands to POLYTE: P fragments -> `c.1.fq` (P reads)
followed by `c.2.fq` (P reads); after `cat p1 p2 | rename_fas.py` (polyte.tune_params.py:291-292)
read i of /1 gets the name i and read i of /2 the name P+i.  /2 is the reverse-complement mate.

Generator: one random ACGT genome of length G = N*L/(S*cov) (N = 2P reads), S strains = genome +
i.i.d. SNPs at `snp_rate`, fragments sampled uniformly with insert ~ N(insert_mean, insert_sd),
random strand per fragment, i.i.d. substitution errors at `err_rate`; no indels, no N.
"""
import numpy as np

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)


class ReadBatch:
    """Concatenated ASCII reads + offsets: the host-side layout the C ABI takes (hlo_pack_reads)."""

    def __init__(self, bases, base_off, cluster_off):
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8)
        self.base_off = np.ascontiguousarray(base_off, dtype=np.uint64)
        self.cluster_off = np.ascontiguousarray(cluster_off, dtype=np.uint32)

    @property
    def n_reads(self):
        return len(self.base_off) - 1

    @property
    def n_clusters(self):
        return len(self.cluster_off) - 1

    @property
    def n_bases(self):
        return int(self.base_off[-1])

    def read(self, i):
        return self.bases[int(self.base_off[i]):int(self.base_off[i + 1])].tobytes()

    def cluster_reads(self, c):
        return [self.read(i) for i in range(int(self.cluster_off[c]), int(self.cluster_off[c + 1]))]

    @classmethod
    def from_seqs(cls, seqs_per_cluster):
        flat, coff = [], [0]
        for cl in seqs_per_cluster:
            flat.extend(s.encode() if isinstance(s, str) else bytes(s) for s in cl)
            coff.append(len(flat))
        lens = np.fromiter((len(s) for s in flat), dtype=np.uint64, count=len(flat))
        off = np.zeros(len(flat) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        bases = np.frombuffer(b"".join(flat), dtype=np.uint8) if flat else np.zeros(0, dtype=np.uint8)
        return cls(bases, off, np.asarray(coff, dtype=np.uint32))

    @classmethod
    def concat(cls, batches):
        bases = np.concatenate([b.bases for b in batches])
        offs, coffs = [np.zeros(1, dtype=np.uint64)], [np.zeros(1, dtype=np.uint32)]
        bo, ro = 0, 0
        for b in batches:
            offs.append(b.base_off[1:] + np.uint64(bo))
            coffs.append(b.cluster_off[1:] + np.uint32(ro))
            bo += b.n_bases
            ro += b.n_reads
        return cls(bases, np.concatenate(offs), np.concatenate(coffs))


def _revcomp_codes(a):
    return (3 - a)[:, ::-1]


def make_cluster(n_fragments=15000, read_len=150, n_strains=5, strain_cov=10.0, snp_rate=0.01, err_rate=0.005,
                 insert_mean=450.0, insert_sd=27.0, seed=1, n_rate=0.0):
    """Returns (codes_1, codes_2): two (P, L) uint8 arrays of base codes 0..3 (4 = N)."""
    P, L, S = int(n_fragments), int(read_len), int(n_strains)
    N = 2 * P
    G = max(int(round(N * L / (S * strain_cov))), int(insert_mean + 6 * insert_sd) + L)
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, size=G, dtype=np.uint8)
    strains = np.empty((S, G), dtype=np.uint8)
    for s in range(S):
        r = np.random.default_rng(seed * 1000003 + 2 + s)
        g = genome.copy()
        snp = r.random(G) < snp_rate
        g[snp] = (g[snp] + r.integers(1, 4, size=int(snp.sum()), dtype=np.uint8)) % 4
        strains[s] = g
    fr = np.random.default_rng(seed * 1000003 + 50)
    strain_of = fr.integers(0, S, size=P)
    ins = np.clip(np.rint(fr.normal(insert_mean, insert_sd, size=P)).astype(np.int64), L, G)
    start = (fr.random(P) * (G - ins + 1)).astype(np.int64)
    minus = fr.random(P) < 0.5
    ar = np.arange(L, dtype=np.int64)
    left = strains[strain_of[:, None], start[:, None] + ar[None, :]]
    right = strains[strain_of[:, None], (start + ins - L)[:, None] + ar[None, :]]
    right_rc = _revcomp_codes(right)
    left_rc = _revcomp_codes(left)
    r1 = np.where(minus[:, None], right_rc, left)
    r2 = np.where(minus[:, None], left, right_rc)
    del left_rc
    er = np.random.default_rng(seed * 1000003 + 100)
    for arr in (r1, r2):
        e = er.random(arr.shape) < err_rate
        arr[e] = (arr[e] + er.integers(1, 4, size=int(e.sum()), dtype=np.uint8)) % 4
        if n_rate > 0:
            arr[er.random(arr.shape) < n_rate] = 4
    return r1, r2


_ASCII5 = np.frombuffer(b"ACGTN", dtype=np.uint8)


def cluster_batch(n_clusters=1, seed0=1, **kw):
    """`n_clusters` clusters with seeds seed0 + c, as one ReadBatch (p1 reads then p2 reads per cluster)."""
    parts = []
    coff = [0]
    lens = []
    for c in range(n_clusters):
        r1, r2 = make_cluster(seed=seed0 + c, **kw)
        codes = np.concatenate([r1, r2], axis=0)
        parts.append(_ASCII5[codes].reshape(-1))
        lens.append(np.full(codes.shape[0], codes.shape[1], dtype=np.uint64))
        coff.append(coff[-1] + codes.shape[0])
    bases = np.concatenate(parts)
    lens = np.concatenate(lens)
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    return ReadBatch(bases, off, np.asarray(coff, dtype=np.uint32))


def write_fastq_pair(batch, c, path1, path2, qual=b"I"):
    """Write cluster c as `X.1.fq` / `X.2.fq` the way get_fq_cluster.py:26-27 names them."""
    r0, r1 = int(batch.cluster_off[c]), int(batch.cluster_off[c + 1])
    P = (r1 - r0) // 2
    for path, lo, tag in ((path1, r0, b"/1"), (path2, r0 + P, b"/2")):
        with open(path, "wb") as f:
            for i in range(P):
                s = batch.read(lo + i)
                f.write(b"@frag%d" % i + tag + b"\n" + s + b"\n+\n" + qual * len(s) + b"\n")


def write_fasta(batch, c, path):
    """`s_p1_p2.fasta` of polyte.tune_params.py:491: names 0..N-1 in file order."""
    r0, r1 = int(batch.cluster_off[c]), int(batch.cluster_off[c + 1])
    with open(path, "wb") as f:
        for i in range(r0, r1):
            f.write(b">%d\n" % (i - r0) + batch.read(i) + b"\n")


# BASELINE.json configs as generator keyword sets
CONFIGS = {
    "config2": dict(n_fragments=15000, read_len=150, n_strains=5, strain_cov=10.0, snp_rate=0.01, err_rate=0.005),
    "config4": dict(n_fragments=15000, read_len=250, n_strains=20, strain_cov=10.0, snp_rate=0.01, err_rate=0.005),
    "config5": dict(n_fragments=100000, read_len=150, n_strains=5, strain_cov=10.0, snp_rate=0.01, err_rate=0.005),
}
