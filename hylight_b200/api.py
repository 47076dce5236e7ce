"""ctypes binding of libhlovlp (include/hlovlp.h) — the only way Python reaches the CUDA path.

There is no CPU fallback: importing works without a GPU (so the symbol table can be checked), but
every compute entry point raises `HloError` when the library or a CUDA device is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# HLO_LIB: another build of the same library (the bounds-checking build `make -C hylight_b200/csrc debug`)
LIB_PATH = os.environ.get("HLO_LIB") or os.path.join(_HERE, "libhlovlp.so")

CAND_DT = np.dtype([("q", "<u4"), ("t", "<u4"), ("diag", "<i4"), ("strand_anchor", "<u4")])
REC_DT = np.dtype([("idA", "<u4"), ("idB", "<u4"), ("OHA", "<i4"), ("OHB", "<i4"), ("OLA", "<u4"), ("OLB", "<u4"),
                   ("K", "<u4"), ("ori", "<u4")])
REC16_DT = np.dtype([("idA", "<u4"), ("idB", "<u4"), ("OHA", "<i2"), ("OHB", "<i2"), ("OLA", "<u2"), ("K_ori", "<u2")])
REC8_DT = np.dtype([("idA", "<u2"), ("idB", "<u2"), ("OHA", "<i2"), ("K_ori", "<u2")])
SCORE_DT = np.dtype([("score", "<f8"), ("mismatch_rate", "<f8")])
AUX_DT = np.dtype([("cluster", "<u4"), ("blocklen", "<u4")])
MZ_DT = np.dtype([("x", "<u8"), ("y", "<u8")])

# every symbol include/hlovlp.h declares (tests/test_abi.py checks the header against this list too)
ABI_SYMBOLS = [
    "hlo_params_default", "hlo_abi_version", "hlo_strerror", "hlo_last_error", "hlo_device_count",
    "hlo_ctx_create", "hlo_ctx_destroy", "hlo_host_alloc", "hlo_host_free", "hlo_pack_reads", "hlo_reads_free",
    "hlo_sketch", "hlo_candidates", "hlo_verify", "hlo_overlap_batch", "hlo_overlap_host", "hlo_get_timing",
    "hlo_write_sfo", "hlo_write_overlaps", "hlo_format_overlap_line", "hlo_set_option", "hlo_stream", "hlo_score_records",
    "hlo_write_overlaps_scored", "hlo_write_overlaps16", "hlo_write_sfo16", "hlo_submit", "hlo_wait", "hlo_set_score_params",
    "hlo_write_overlaps8", "hlo_write_sfo8", "hlo_write_binary16", "hlo_pack2b",
    "hlo_debug_check_reads",
]


class HloError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("libhlovlp: %s (status %d)" % (msg, status))
        self.status = status


class Params(C.Structure):
    _fields_ = [("k", C.c_int32), ("w", C.c_int32), ("min_seeds", C.c_int32), ("min_chain_score", C.c_int32),
                ("mid_occ", C.c_int32), ("match_score", C.c_int32), ("mismatch_pen", C.c_int32),
                ("ambi_pen", C.c_int32), ("end_bonus", C.c_int32), ("min_dp_score", C.c_int32),
                ("min_ovlp_len", C.c_int32), ("max_overhang", C.c_int32), ("min_identity", C.c_double),
                ("overhang_ratio", C.c_double), ("zdrop", C.c_int32), ("pair_rule", C.c_int32)]


class Recs(C.Structure):
    _fields_ = [("rec", C.c_void_p), ("aux", C.c_void_p), ("n", C.c_uint64), ("cluster_off", C.c_void_p),
                ("n_clusters", C.c_uint32), ("rec16", C.c_void_p), ("rec8", C.c_void_p), ("score", C.c_void_p)]


class Batch(C.Structure):
    """hlo_batch: the input of hlo_submit"""
    _fields_ = [("bases", C.c_void_p), ("quals", C.c_void_p), ("base_off", C.c_void_p), ("n_reads", C.c_uint32),
                ("cluster_off", C.c_void_p), ("n_clusters", C.c_uint32), ("ids", C.c_void_p),
                ("packed", C.c_void_p), ("read_len", C.c_void_p), ("amb", C.c_void_p), ("n_amb", C.c_uint64)]


class Timing(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "pack_ms", "sketch_ms", "sort_ms", "index_ms", "pairgen_ms",
                                         "verify_ms", "dedup_ms", "d2h_ms", "total_ms", "verify_fast_ms", "sort_fallback")] + \
               [(n, C.c_uint64) for n in ("n_reads", "n_bases", "n_minimizers", "n_seed_hits", "n_candidates",
                                          "n_accepted", "n_records", "n_deferred", "n_overflow_reads",
                                          "verify_launches", "total_launches", "digest", "n_reruns", "n_syncs")] + \
               [("mid_occ_lo", C.c_uint32), ("mid_occ_hi", C.c_uint32), ("score_ms", C.c_float), ("n_sketch_reruns", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def lib():
    """Load libhlovlp.so (built in-tree by `make -C hylight_b200/csrc`); raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HloError(-2, "%s not found: build it with `make -C hylight_b200/csrc` "
                               "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.hlo_strerror.restype = C.c_char_p
        L.hlo_last_error.restype = C.c_char_p
        L.hlo_last_error.argtypes = [C.c_void_p]
        L.hlo_host_alloc.restype = C.c_void_p
        L.hlo_host_alloc.argtypes = [C.c_size_t]
        L.hlo_host_free.argtypes = [C.c_void_p]
        L.hlo_ctx_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.hlo_ctx_destroy.argtypes = [C.c_void_p]
        L.hlo_pack_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                     C.c_void_p, C.POINTER(C.c_void_p)]
        L.hlo_reads_free.argtypes = [C.c_void_p]
        L.hlo_sketch.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_uint64)]
        L.hlo_candidates.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.hlo_verify.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(Recs)]
        L.hlo_overlap_batch.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Recs)]
        L.hlo_overlap_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                       C.POINTER(Recs)]
        L.hlo_get_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
        L.hlo_write_sfo.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p]
        L.hlo_write_overlaps.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p]
        L.hlo_format_overlap_line.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.hlo_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        L.hlo_score_records.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_double,
                                        C.c_uint32, C.POINTER(C.c_void_p)]
        L.hlo_stream.restype = C.c_void_p
        L.hlo_stream.argtypes = [C.c_void_p]
        L.hlo_submit.argtypes = [C.c_void_p, C.POINTER(Batch)]
        L.hlo_wait.argtypes = [C.c_void_p, C.POINTER(Recs)]
        L.hlo_set_score_params.argtypes = [C.c_void_p, C.c_double, C.c_uint32]
        L.hlo_write_overlaps16.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p]
        L.hlo_write_sfo16.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p]
        L.hlo_write_overlaps8.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_char_p]
        L.hlo_write_sfo8.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_char_p]
        L.hlo_write_binary16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p]
        L.hlo_write_overlaps_scored.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p]
        _lib = L
    return _lib


def default_params(**kw):
    p = Params()
    lib().hlo_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def device_count():
    return int(lib().hlo_device_count())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _copy_out(ptr, n, dt):
    if not n:
        return np.zeros(0, dtype=dt)
    buf = (C.c_uint8 * (n * dt.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dt, count=n).copy()


class PinnedArray:
    """A numpy view over page-locked host memory from hlo_host_alloc (fast H2D for the e2e path)."""

    def __init__(self, nbytes):
        self.ptr = lib().hlo_host_alloc(max(int(nbytes), 1))
        if not self.ptr:
            raise HloError(-4, "hlo_host_alloc(%d) failed" % nbytes)
        self.nbytes = int(nbytes)
        self.array = np.frombuffer((C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr), dtype=np.uint8,
                                   count=self.nbytes)

    def free(self):
        if self.ptr:
            self.array = None
            lib().hlo_host_free(self.ptr)
            self.ptr = None


class Reads:
    def __init__(self, ctx, handle, n_reads, n_clusters):
        self.ctx, self.handle, self.n_reads, self.n_clusters = ctx, handle, n_reads, n_clusters

    def free(self):
        if self.handle:
            lib().hlo_reads_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One CUDA device + stream.  Single-threaded; create one per device/thread."""

    def __init__(self, device=0, params=None, **kw):
        self.params = params if params is not None else default_params(**kw)
        h = C.c_void_p()
        rc = lib().hlo_ctx_create(int(device), C.byref(self.params), C.byref(h))
        if rc:
            raise HloError(rc, lib().hlo_strerror(rc).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            lib().hlo_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise HloError(rc, "%s: %s" % (lib().hlo_strerror(rc).decode(), lib().hlo_last_error(self.h).decode()))

    def set_option(self, name, value):
        self._check(lib().hlo_set_option(self.h, name.encode(), int(value)))

    def pack(self, batch, ids=None):
        """batch: anything with .bases (u8), .base_off (u64), .cluster_off (u32)."""
        bases = np.ascontiguousarray(batch.bases, dtype=np.uint8)
        off = np.ascontiguousarray(batch.base_off, dtype=np.uint64)
        coff = np.ascontiguousarray(batch.cluster_off, dtype=np.uint32)
        if ids is None:
            ids = getattr(batch, "ids", None)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint32)
        h = C.c_void_p()
        self._check(lib().hlo_pack_reads(self.h, _ptr(bases), _ptr(off), len(off) - 1, _ptr(coff), len(coff) - 1,
                                         _ptr(ids), C.byref(h)))
        return Reads(self, h, len(off) - 1, len(coff) - 1)

    def sketch(self, reads):
        mz, ro, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._check(lib().hlo_sketch(self.h, reads.handle, C.byref(mz), C.byref(ro), C.byref(n)))
        return _copy_out(mz.value, n.value, MZ_DT), _copy_out(ro.value, reads.n_reads + 1, np.dtype("<u8"))

    def candidates(self, reads):
        p, n = C.c_void_p(), C.c_uint64()
        self._check(lib().hlo_candidates(self.h, reads.handle, C.byref(p), C.byref(n)))
        return _copy_out(p.value, n.value, CAND_DT)

    def _recs(self, r, copy=True):
        if not r.rec and not r.rec16 and not r.rec8:  # option "d2h" = 0: only the count came back
            return None, None, None
        if r.rec8:
            rec = _copy_out(r.rec8, r.n, REC8_DT)
        elif r.rec16:
            rec = _copy_out(r.rec16, r.n, REC16_DT)
        else:
            rec = _copy_out(r.rec, r.n, REC_DT)
        aux = _copy_out(r.aux, r.n, AUX_DT) if r.aux else None
        coff = _copy_out(r.cluster_off, r.n_clusters + 1, np.dtype("<u8")) if r.cluster_off else None
        self.last_score = _copy_out(r.score, r.n, SCORE_DT) if r.score else None
        return rec, aux, coff

    def verify(self, reads, cands):
        cands = np.ascontiguousarray(cands, dtype=CAND_DT)
        r = Recs()
        self._check(lib().hlo_verify(self.h, reads.handle, _ptr(cands), len(cands), C.byref(r)))
        return self._recs(r)[:2]

    def overlap_batch(self, reads):
        r = Recs()
        self._check(lib().hlo_overlap_batch(self.h, reads.handle, C.byref(r)))
        return self._recs(r)

    def overlap_host_raw(self, bases_ptr, off_ptr, n_reads, coff_ptr, n_clusters):
        """e2e entry with raw host pointers; returns the ctx-owned Recs struct without copying."""
        r = Recs()
        self._check(lib().hlo_overlap_host(self.h, bases_ptr, off_ptr, n_reads, coff_ptr, n_clusters, C.byref(r)))
        return r

    def overlap_host(self, batch):
        bases = np.ascontiguousarray(batch.bases, dtype=np.uint8)
        off = np.ascontiguousarray(batch.base_off, dtype=np.uint64)
        coff = np.ascontiguousarray(batch.cluster_off, dtype=np.uint32)
        r = self.overlap_host_raw(_ptr(bases), _ptr(off), len(off) - 1, _ptr(coff), len(coff) - 1)
        return self._recs(r)

    def submit(self, batch, quals=None, ids=None):
        """hlo_submit: enqueue upload + the whole chain, return at once.  The arrays are kept alive until wait()."""
        keep = [np.ascontiguousarray(batch.bases, dtype=np.uint8), np.ascontiguousarray(batch.base_off, dtype=np.uint64),
                np.ascontiguousarray(batch.cluster_off, dtype=np.uint32),
                None if quals is None else np.ascontiguousarray(quals, dtype=np.uint8),
                None if ids is None else np.ascontiguousarray(ids, dtype=np.uint32)]
        b = Batch(_ptr(keep[0]), _ptr(keep[3]), _ptr(keep[1]), len(keep[1]) - 1, _ptr(keep[2]), len(keep[2]) - 1, _ptr(keep[4]),
                  None, None, None, 0)
        self._inflight = keep
        self._check(lib().hlo_submit(self.h, C.byref(b)))

    def submit_raw(self, bases_ptr, off_ptr, n_reads, coff_ptr, n_clusters, quals_ptr=None):
        b = Batch(bases_ptr, quals_ptr, off_ptr, n_reads, coff_ptr, n_clusters, None, None, None, None, 0)
        self._check(lib().hlo_submit(self.h, C.byref(b)))

    def submit_packed(self, pk, cluster_off):
        """hlo_submit of a batch in the 2-bit packed form (pk = pack2b(batch)): a quarter of the upload"""
        words, lens, amb = pk
        coff = np.ascontiguousarray(cluster_off, dtype=np.uint32)
        self._inflight = [words, lens, amb, coff]
        b = Batch(None, None, None, len(lens), _ptr(coff), len(coff) - 1, None, _ptr(words), _ptr(lens), _ptr(amb) if len(amb) else None,
                  len(amb))
        self._check(lib().hlo_submit(self.h, C.byref(b)))

    def submit_packed_raw(self, words_ptr, len_ptr, n_reads, coff_ptr, n_clusters, amb_ptr=None, n_amb=0):
        b = Batch(None, None, None, n_reads, coff_ptr, n_clusters, None, words_ptr, len_ptr, amb_ptr, n_amb)
        self._check(lib().hlo_submit(self.h, C.byref(b)))

    def wait_raw(self):
        r = Recs()
        self._check(lib().hlo_wait(self.h, C.byref(r)))
        return r

    def wait(self):
        """hlo_wait: (records, aux, cluster offsets); fused scores, when asked for, in self.last_score"""
        r = self.wait_raw()
        self._inflight = None
        return self._recs(r)

    def check_canaries(self, reads=None):
        """hlo_debug_check_reads: raises when a kernel wrote into the padding around the packed rows"""
        L = lib()
        L.hlo_debug_check_reads.argtypes = [C.c_void_p, C.c_void_p]
        self._check(L.hlo_debug_check_reads(self.h, reads.handle if reads is not None else None))

    def set_score_params(self, mismatch_prob=0.0, min_read_len=0):
        self._check(lib().hlo_set_score_params(self.h, float(mismatch_prob), int(min_read_len)))

    def score_records(self, reads, quals, rec, rec_cluster=None, mismatch_prob=0.0, min_read_len=0):
        """HaploConduct's overlap_score per record (EdgeCalculator.cpp:67-139): array (n, 2) of (score, mismatch_rate)."""
        quals = np.ascontiguousarray(quals, dtype=np.uint8)
        rec = np.ascontiguousarray(rec, dtype=REC_DT)
        rc = None if rec_cluster is None else np.ascontiguousarray(rec_cluster, dtype=np.uint32)
        out = C.c_void_p()
        self._check(lib().hlo_score_records(self.h, reads.handle, _ptr(quals), _ptr(rec), _ptr(rc), len(rec), float(mismatch_prob),
                                            int(min_read_len), C.byref(out)))
        return _copy_out(out.value, len(rec), np.dtype([("score", "<f8"), ("mismatch_rate", "<f8")]))

    def stream_handle(self):
        """cudaStream_t of the ctx as an integer (wrap with torch.cuda.ExternalStream to record events on it)."""
        return int(lib().hlo_stream(self.h) or 0)

    def timing(self):
        t = Timing()
        self._check(lib().hlo_get_timing(self.h, C.byref(t)))
        return t.as_dict()


def expand_rec16(rec16):
    """hlo_rec16 array -> hlo_rec array (hlo_rec16_expand of the header)"""
    out = np.zeros(len(rec16), dtype=REC_DT)
    for f in ("idA", "idB", "OHA", "OHB", "OLA"):
        out[f] = rec16[f]
    out["OLB"] = rec16["OLA"]
    out["K"] = rec16["K_ori"] & 0x7fff
    out["ori"] = rec16["K_ori"] >> 15
    return out


def pack2b(batch):
    """hlo_pack2b: ASCII reads -> (packed words u32, read lengths u32, ambiguous positions u64 = read << 32 | pos)"""
    bases = np.ascontiguousarray(batch.bases, dtype=np.uint8)
    off = np.ascontiguousarray(batch.base_off, dtype=np.uint64)
    n = len(off) - 1
    lens = np.zeros(max(n, 1), dtype=np.uint32)
    words = np.zeros(int(((np.diff(off).astype(np.int64) + 15) // 16).sum()) + 1, dtype=np.uint32)
    amb = np.zeros(len(bases) + 1, dtype=np.uint64)
    nw, na = C.c_uint64(), C.c_uint64()
    L = lib()
    L.hlo_pack2b.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p, C.c_uint64,
                             C.POINTER(C.c_uint64)]
    rc = L.hlo_pack2b(_ptr(bases), _ptr(off), n, _ptr(words), C.byref(nw), _ptr(lens), _ptr(amb), len(amb), C.byref(na))
    if rc:
        raise HloError(rc, L.hlo_strerror(rc).decode())
    return words[:nw.value].copy(), lens[:n].copy(), amb[:na.value].copy()


def expand_rec8(rec8, len_by_id):
    """hlo_rec8 array + read lengths by name -> hlo_rec array (hlo_rec8_expand of the header)"""
    len_by_id = np.asarray(len_by_id, dtype=np.int64)
    out = np.zeros(len(rec8), dtype=REC_DT)
    a = rec8["OHA"].astype(np.int64)
    la, lb = len_by_id[rec8["idA"]], len_by_id[rec8["idB"]]
    out["idA"], out["idB"], out["OHA"], out["OHB"] = rec8["idA"], rec8["idB"], a, a + lb - la
    ol = np.where(a >= 0, np.minimum(la - a, lb), np.minimum(lb + a, la))
    out["OLA"], out["OLB"] = ol, ol
    out["K"] = rec8["K_ori"] & 0x7fff
    out["ori"] = rec8["K_ori"] >> 15
    return out


def _wr(fn, *args):
    rc = fn(*args)
    if rc:
        raise HloError(rc, lib().hlo_strerror(rc).decode())


def write_overlaps16(rec16, path):
    rec16 = np.ascontiguousarray(rec16, dtype=REC16_DT)
    _wr(lib().hlo_write_overlaps16, _ptr(rec16), len(rec16), os.fsencode(path))


def write_overlaps8(rec8, len_by_id, path):
    rec8 = np.ascontiguousarray(rec8, dtype=REC8_DT)
    lens = np.ascontiguousarray(len_by_id, dtype=np.uint32)
    _wr(lib().hlo_write_overlaps8, _ptr(rec8), len(rec8), _ptr(lens), len(lens), os.fsencode(path))


def write_sfo8(rec8, len_by_id, path):
    rec8 = np.ascontiguousarray(rec8, dtype=REC8_DT)
    lens = np.ascontiguousarray(len_by_id, dtype=np.uint32)
    _wr(lib().hlo_write_sfo8, _ptr(rec8), len(rec8), _ptr(lens), len(lens), os.fsencode(path))


def write_binary16(rec16, path, score=None):
    rec16 = np.ascontiguousarray(rec16, dtype=REC16_DT)
    sc = None if score is None else np.ascontiguousarray(score, dtype=SCORE_DT)
    _wr(lib().hlo_write_binary16, _ptr(rec16), _ptr(sc), len(rec16), os.fsencode(path))


def read_binary16(path):
    """reader of the binary hand-off file (the Python twin of the EdgeCalculator patch in INTEGRATION.md §5)"""
    raw = open(path, "rb").read()
    assert raw[:4] == b"HLOB", "not a hylight-ovlp binary overlaps file"
    version, rec_bytes, flags = np.frombuffer(raw, dtype="<u4", count=3, offset=4)
    n = int(np.frombuffer(raw, dtype="<u8", count=1, offset=16)[0])
    assert version == 1 and rec_bytes == REC16_DT.itemsize
    rec = np.frombuffer(raw, dtype=REC16_DT, count=n, offset=32).copy()
    score = np.frombuffer(raw, dtype=SCORE_DT, count=n, offset=32 + 16 * n).copy() if flags & 1 else None
    return rec, score


_M64 = (1 << 64) - 1


def _mix64(z):
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9))
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94d049bb133111eb))
    return z ^ (z >> np.uint64(31))


def records_digest(rec):
    """hlo_timing.digest recomputed on the host from hlo_rec records (records.cu:rec_digest): order-independent sum"""
    rec = np.ascontiguousarray(rec, dtype=REC_DT)
    with np.errstate(over="ignore"):
        w0 = rec["idA"].astype(np.uint64) | (rec["idB"].astype(np.uint64) << np.uint64(32))
        w1 = rec["OHA"].astype(np.int64).astype(np.uint64) & np.uint64(0xffffffff) | \
            ((rec["OHB"].astype(np.int64).astype(np.uint64) & np.uint64(0xffffffff)) << np.uint64(32))
        w2 = rec["OLA"].astype(np.uint64) | (rec["K"].astype(np.uint64) << np.uint64(32)) | (rec["ori"].astype(np.uint64) << np.uint64(63))
        h = _mix64(w0 ^ _mix64(w1 ^ _mix64(w2 + np.uint64(0x9e3779b97f4a7c15))))
        return int(h.sum(dtype=np.uint64))


def write_sfo(rec, path):
    rec = np.ascontiguousarray(rec, dtype=REC_DT)
    rc = lib().hlo_write_sfo(_ptr(rec), len(rec), os.fsencode(path))
    if rc:
        raise HloError(rc, lib().hlo_strerror(rc).decode())


def write_overlaps(rec, path):
    rec = np.ascontiguousarray(rec, dtype=REC_DT)
    rc = lib().hlo_write_overlaps(_ptr(rec), len(rec), os.fsencode(path))
    if rc:
        raise HloError(rc, lib().hlo_strerror(rc).decode())


def format_overlap_line(rec_row):
    r = np.ascontiguousarray(np.asarray(rec_row, dtype=REC_DT).reshape(1))
    buf = C.create_string_buffer(160)
    n = lib().hlo_format_overlap_line(_ptr(r), buf, 160)
    if n < 0:
        raise HloError(n, "hlo_format_overlap_line")
    return buf.value.decode()
