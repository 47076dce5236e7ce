"""ctypes binding of the CPU oracle (oracle/hlo_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  The product package (hylight_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

CAND_DT = np.dtype([("q", "<u4"), ("t", "<u4"), ("diag", "<i4"), ("strand_anchor", "<u4")])
REC_DT = np.dtype([("idA", "<u4"), ("idB", "<u4"), ("OHA", "<i4"), ("OHB", "<i4"), ("OLA", "<u4"), ("OLB", "<u4"),
                   ("K", "<u4"), ("ori", "<u4")])
MZ_DT = np.dtype([("x", "<u8"), ("y", "<u8")])
PAF_DT = np.dtype([("q", "<u4"), ("t", "<u4"), ("qid", "<u4"), ("tid", "<u4"), ("ql", "<i4"), ("qs", "<i4"),
                   ("qe", "<i4"), ("strand", "<i4"), ("tl", "<i4"), ("ts", "<i4"), ("te", "<i4"), ("matches", "<i4"),
                   ("blocklen", "<i4"), ("dp_max", "<i4"), ("clip_l", "<i4"), ("clip_r", "<i4"), ("n_ambi", "<i4"),
                   ("valid", "<i4"), ("zdrop_mid", "<i4")])


class Params(C.Structure):
    _fields_ = [("k", C.c_int32), ("w", C.c_int32), ("min_seeds", C.c_int32), ("min_chain_score", C.c_int32),
                ("mid_occ", C.c_int32), ("match_score", C.c_int32), ("mismatch_pen", C.c_int32),
                ("ambi_pen", C.c_int32), ("end_bonus", C.c_int32), ("min_dp_score", C.c_int32),
                ("min_ovlp_len", C.c_int32), ("max_overhang", C.c_int32), ("min_identity", C.c_double),
                ("overhang_ratio", C.c_double), ("zdrop", C.c_int32), ("pair_rule", C.c_int32)]


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "hlo_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "hlovlp.h")
    stale = (not os.path.exists(so)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(so) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return so


_ECREF = None


def ecref():
    """oracle/_ref/libecref.so: the reference's OWN EdgeCalculator::score / overlap_score
    (tools/HaploConduct/src/EdgeCalculator.cpp:25-140), compiled by oracle/Makefile where /root/reference is present
    (the build container; the file travels to the GPU box).  None when it was never built."""
    global _ECREF
    if _ECREF is None:
        so = os.path.join(_HERE, "_ref", "libecref.so")
        if not os.path.exists(so) and os.path.exists("/root/reference/tools/HaploConduct/src/EdgeCalculator.cpp"):
            subprocess.call(["make", "-C", _HERE, "_ref/libecref.so"], stdout=subprocess.DEVNULL)
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        L.ecref_overlap_score.restype = C.c_double
        L.ecref_overlap_score.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_uint,
                                          C.c_double, C.c_uint, C.POINTER(C.c_double)]
        L.ecref_phred_to_prob.restype = C.c_double
        L.ecref_phred_to_prob.argtypes = [C.c_int]
        _ECREF = L
    return _ECREF


_COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def ecref_score_record(b, quals, rec_row, cluster_base=0, mismatch_prob=0.0, min_read_len=0):
    """One record through the reference's own overlap_score, with the S-S orientation handling of compute_overlap
    (EdgeCalculator.cpp:199-221; Read.h:172-201 get_rev_comp / get_rev_phred) done here: (score, mismatch_rate)."""
    L = ecref()
    OHA = int(rec_row["OHA"])
    if OHA >= 0:
        id1, id2, pos1, fw1, fw2 = int(rec_row["idA"]), int(rec_row["idB"]), OHA, True, int(rec_row["ori"]) == 0
    else:
        id1, id2, pos1, fw1, fw2 = int(rec_row["idB"]), int(rec_row["idA"]), -OHA, int(rec_row["ori"]) == 0, True
    quals = np.ascontiguousarray(quals, dtype=np.uint8)

    def oriented(i, fw):
        lo, hi = int(b.base_off[cluster_base + i]), int(b.base_off[cluster_base + i + 1])
        s = bytes(b.bases[lo:hi]).upper()          # FastqStorage.cpp:123 upper-cases on load
        q = bytes(quals[lo:hi])
        return (s, q) if fw else (s.translate(_COMP)[::-1], q[::-1])

    s1, q1 = oriented(id1, fw1)
    s2, q2 = oriented(id2, fw2)
    mr = C.c_double()
    sc = L.ecref_overlap_score(s1, len(s1), s2, len(s2), q1, q2, pos1, float(mismatch_prob), int(min_read_len), C.byref(mr))
    return sc, mr.value


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_sketch.restype = C.c_uint64
        _LIB.orc_filter_chunked.restype = C.c_uint64
        _LIB.orc_filter_chunked_v3.restype = C.c_uint64
        _LIB.orc_paf_rows_to_recs.restype = C.c_uint64
        _LIB.orc_sort_uniq_recs.restype = C.c_uint64
        _LIB.orc_sizeof_paf.restype = C.c_size_t
        _LIB.orc_hash64.restype = C.c_uint64
        _LIB.orc_hash64.argtypes = [C.c_uint64, C.c_uint64]
        assert _LIB.orc_sizeof_paf() == PAF_DT.itemsize
    return _LIB


def default_params(**kw):
    p = Params()
    lib().orc_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Batch:
    """A batch of clusters in the layout the C ABI takes: concatenated ASCII bases + offsets."""

    def __init__(self, seqs_per_cluster, ids=None):
        # seqs_per_cluster: list of lists of str/bytes
        flat = []
        coff = [0]
        for cl in seqs_per_cluster:
            flat.extend(cl)
            coff.append(len(flat))
        flat = [s.encode() if isinstance(s, str) else bytes(s) for s in flat]
        lens = np.fromiter((len(s) for s in flat), dtype=np.uint64, count=len(flat))
        self.base_off = np.zeros(len(flat) + 1, dtype=np.uint64)
        np.cumsum(lens, out=self.base_off[1:])
        self.bases = np.frombuffer(b"".join(flat) + b"\0", dtype=np.uint8).copy()
        self.cluster_off = np.asarray(coff, dtype=np.uint32)
        self.n_reads = len(flat)
        self.n_clusters = len(seqs_per_cluster)
        self.ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint32)

    @classmethod
    def from_arrays(cls, bases, base_off, cluster_off, ids=None):
        b = cls.__new__(cls)
        b.bases = np.ascontiguousarray(bases, dtype=np.uint8)
        b.base_off = np.ascontiguousarray(base_off, dtype=np.uint64)
        b.cluster_off = np.ascontiguousarray(cluster_off, dtype=np.uint32)
        b.n_reads = len(b.base_off) - 1
        b.n_clusters = len(b.cluster_off) - 1
        b.ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint32)
        return b

    def read_cluster_base(self):
        out = np.zeros(self.n_reads, dtype=np.uint32)
        for c in range(self.n_clusters):
            out[self.cluster_off[c]:self.cluster_off[c + 1]] = self.cluster_off[c]
        return out


def sketch_one(seq, w=11, k=21, rid=0):
    seq = seq.encode() if isinstance(seq, str) else bytes(seq)
    out = np.zeros(len(seq) + 8, dtype=MZ_DT)
    n = lib().orc_sketch(seq, C.c_int(len(seq)), C.c_int(w), C.c_int(k), C.c_uint32(rid), _ptr(out), C.c_uint64(len(out)))
    return out[:n].copy()


def sketch_all(b, p=None):
    p = p or default_params()
    mz_p = C.c_void_p()
    roff = np.zeros(b.n_reads + 1, dtype=np.uint64)
    rc = lib().orc_sketch_all(_ptr(b.bases), _ptr(b.base_off), C.c_uint32(b.n_reads), C.c_int(p.w), C.c_int(p.k),
                              C.byref(mz_p), _ptr(roff))
    assert rc == 0, rc
    n = int(roff[-1])
    mz = np.ctypeslib.as_array(C.cast(mz_p, C.POINTER(C.c_uint8)), shape=(max(n, 1) * MZ_DT.itemsize,))[:n * MZ_DT.itemsize]
    mz = mz.view(MZ_DT).copy() if n else np.zeros(0, dtype=MZ_DT)
    lib().orc_free(mz_p)
    return mz, roff


def candidates(b, p=None):
    p = p or default_params()
    out_p = C.c_void_p()
    n = C.c_uint64()
    hits = C.c_uint64()
    rc = lib().orc_candidates(_ptr(b.bases), _ptr(b.base_off), C.c_uint32(b.n_reads), _ptr(b.cluster_off),
                              C.c_uint32(b.n_clusters), _ptr(b.ids), C.byref(p), C.byref(out_p), C.byref(n), C.byref(hits))
    assert rc == 0, rc
    nn = n.value
    if nn:
        arr = np.ctypeslib.as_array(C.cast(out_p, C.POINTER(C.c_uint8)), shape=(nn * CAND_DT.itemsize,)).view(CAND_DT).copy()
    else:
        arr = np.zeros(0, dtype=CAND_DT)
    lib().orc_free(out_p)
    return arr, hits.value


def mid_occ(b, p=None):
    """per-cluster occurrence cut-off (minimap2's data-derived mid_occ, or p.mid_occ when > 0)"""
    p = p or default_params()
    out = np.zeros(max(b.n_clusters, 1), dtype=np.uint32)
    rc = lib().orc_mid_occ(_ptr(b.bases), _ptr(b.base_off), C.c_uint32(b.n_reads), _ptr(b.cluster_off), C.c_uint32(b.n_clusters),
                           C.byref(p), _ptr(out))
    assert rc == 0, rc
    return out[:b.n_clusters]


def verify(b, cands, p=None, want_paf=False):
    """Returns (records, blocklen, accept_mask[, paf_rows]) for a candidate list, in candidate order."""
    p = p or default_params()
    cands = np.ascontiguousarray(cands, dtype=CAND_DT)
    n = len(cands)
    rec = np.zeros(max(n, 1), dtype=REC_DT)
    blk = np.zeros(max(n, 1), dtype=np.uint32)
    acc = np.zeros(max(n, 1), dtype=np.uint8)
    paf = np.zeros(max(n, 1), dtype=PAF_DT) if want_paf else None
    m = C.c_uint64()
    rc = lib().orc_verify(_ptr(b.bases), _ptr(b.base_off), C.c_uint32(b.n_reads), _ptr(b.cluster_off),
                          C.c_uint32(b.n_clusters), _ptr(b.ids), C.byref(p), _ptr(cands), C.c_uint64(n), _ptr(rec),
                          _ptr(blk), _ptr(acc), _ptr(paf), C.byref(m))
    assert rc == 0, rc
    out = (rec[:m.value].copy(), blk[:m.value].copy(), acc[:n].astype(bool))
    if want_paf:
        out = out + (paf[:n].copy(),)
    return out


def overlap_batch(b, p=None, dedup=True):
    """Full CPU chain.  Returns (records, blocklen, cluster_rec_off, stats dict)."""
    p = p or default_params()
    rec_p = C.c_void_p()
    blk_p = C.c_void_p()
    coff = np.zeros(b.n_clusters + 1, dtype=np.uint64)
    stats = np.zeros(4, dtype=np.uint64)
    rc = lib().orc_overlap_batch(_ptr(b.bases), _ptr(b.base_off), C.c_uint32(b.n_reads), _ptr(b.cluster_off),
                                 C.c_uint32(b.n_clusters), _ptr(b.ids), C.byref(p), C.c_int(1 if dedup else 0),
                                 C.byref(rec_p), C.byref(blk_p), _ptr(coff), _ptr(stats))
    assert rc == 0, rc
    n = int(coff[-1])
    if n:
        rec = np.ctypeslib.as_array(C.cast(rec_p, C.POINTER(C.c_uint8)), shape=(n * REC_DT.itemsize,)).view(REC_DT).copy()
        blk = np.ctypeslib.as_array(C.cast(blk_p, C.POINTER(C.c_uint32)), shape=(n,)).copy()
    else:
        rec = np.zeros(0, dtype=REC_DT)
        blk = np.zeros(0, dtype=np.uint32)
    lib().orc_free(rec_p)
    lib().orc_free(blk_p)
    st = dict(seed_hits=int(stats[0]), candidates=int(stats[1]), accepted=int(stats[2]), records=int(stats[3]))
    return rec, blk, coff, st


def filter_rows(paf, p=None):
    """rm_intermatch per row (filter_ovlp_inline.py:42-91): boolean mask."""
    p = p or default_params()
    paf = np.ascontiguousarray(paf, dtype=PAF_DT)
    return np.array([bool(lib().orc_filter_row(_ptr(paf[i:i + 1]), C.byref(p))) for i in range(len(paf))], dtype=bool)


def filter_chunked(paf, p=None):
    p = p or default_params()
    paf = np.ascontiguousarray(paf, dtype=PAF_DT)
    keep = np.zeros(max(len(paf), 1), dtype=np.uint64)
    n = lib().orc_filter_chunked(_ptr(paf), C.c_uint64(len(paf)), C.byref(p), _ptr(keep))
    return keep[:n].copy()


def filter_chunked_v3(paf, p=None):
    """filter_trans_ovlp_inline_v3.py on rows in the given order: indices of the surviving rows"""
    p = p or default_params()
    paf = np.ascontiguousarray(paf, dtype=PAF_DT)
    keep = np.zeros(max(len(paf), 1), dtype=np.uint64)
    n = lib().orc_filter_chunked_v3(_ptr(paf), C.c_uint64(len(paf)), C.byref(p), _ptr(keep))
    return keep[:n].copy()


def write_sfo_v3(paf, idx, path):
    paf = np.ascontiguousarray(paf, dtype=PAF_DT)
    idx = np.ascontiguousarray(idx, dtype=np.uint64)
    assert lib().orc_write_sfo_v3(_ptr(paf), _ptr(idx), C.c_uint64(len(idx)), path.encode()) == 0


def paf_rows_to_recs(paf, idx=None):
    paf = np.ascontiguousarray(paf, dtype=PAF_DT)
    n = len(paf) if idx is None else len(idx)
    rec = np.zeros(max(n, 1), dtype=REC_DT)
    if idx is not None:
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
    m = lib().orc_paf_rows_to_recs(_ptr(paf), _ptr(idx), C.c_uint64(n), _ptr(rec))
    return rec[:m].copy()


def write_paf(paf, path):
    paf = np.ascontiguousarray(paf, dtype=PAF_DT)
    assert lib().orc_write_paf(_ptr(paf), C.c_uint64(len(paf)), path.encode()) == 0


def write_sfo(rec, path):
    rec = np.ascontiguousarray(rec, dtype=REC_DT)
    assert lib().orc_write_sfo(_ptr(rec), C.c_uint64(len(rec)), path.encode()) == 0


def write_overlaps(rec, path):
    rec = np.ascontiguousarray(rec, dtype=REC_DT)
    assert lib().orc_write_overlaps(_ptr(rec), C.c_uint64(len(rec)), path.encode()) == 0


def overlap_line(rec_row):
    buf = C.create_string_buffer(160)
    r = np.ascontiguousarray(np.asarray(rec_row, dtype=REC_DT).reshape(1))
    lib().orc_format_overlap_line(_ptr(r), buf, C.c_size_t(160))
    return buf.value.decode()


def name_cmp(a, b):
    return lib().orc_name_cmp(C.c_uint32(a), C.c_uint32(b))


def score_records(b, quals, rec, rec_cluster=None, mismatch_prob=0.0, min_read_len=0):
    """HaploConduct's overlap_score (EdgeCalculator.cpp:67-139) per record: array (n, 2) of (score, mismatch_rate)."""
    rec = np.ascontiguousarray(rec, dtype=REC_DT)
    quals = np.ascontiguousarray(quals, dtype=np.uint8)
    rc = None if rec_cluster is None else np.ascontiguousarray(rec_cluster, dtype=np.uint32)
    if b.ids is not None:  # the consumer (FastqStorage.cpp:109-117) indexes its reads by name: names must be 0..N-1 in file order
        pos = np.arange(b.n_reads, dtype=np.int64) - np.repeat(np.asarray(b.cluster_off[:-1], dtype=np.int64), np.diff(b.cluster_off).astype(np.int64))
        if not np.array_equal(np.asarray(b.ids, dtype=np.int64), pos):
            raise ValueError("score_records: read names must be 0..N-1 in file order inside every cluster")
    out = np.zeros((max(len(rec), 1), 2), dtype=np.float64)
    lib().orc_score_records(_ptr(b.bases), _ptr(quals), _ptr(b.base_off), _ptr(b.cluster_off), _ptr(rec), _ptr(rc),
                            C.c_uint64(len(rec)), C.c_double(mismatch_prob), C.c_uint32(min_read_len), _ptr(out))
    return out[:len(rec)]
