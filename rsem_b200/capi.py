"""ctypes mirror of include/rsem_b200.h (one Python method per C entry point, same names)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librsem_b200.so")


class RsemB200Error(RuntimeError):
    pass


class LenDist(C.Structure):
    _fields_ = [("lb", C.c_int32), ("ub", C.c_int32), ("span", C.c_int32),
                ("pdf", C.POINTER(C.c_double)), ("cdf", C.POINTER(C.c_double))]


class Model(C.Structure):
    _fields_ = [("model_type", C.c_int32), ("M", C.c_int32), ("seed_len", C.c_int32), ("est_rspd", C.c_int32),
                ("rspd_B", C.c_int32), ("has_mld", C.c_int32), ("pro_len", C.c_int32), ("reserved", C.c_int32),
                ("ori", C.c_double * 2), ("gld", LenDist), ("mld", LenDist),
                ("rspd_pdf", C.POINTER(C.c_double)), ("rspd_cdf", C.POINTER(C.c_double)),
                ("profile", C.POINTER(C.c_double)), ("noise_profile", C.POINTER(C.c_double)),
                ("mw", C.POINTER(C.c_double))]


class ModelStats(C.Structure):
    _fields_ = [("profile", C.POINTER(C.c_double)), ("noise_profile", C.POINTER(C.c_double)),
                ("gld_pdf", C.POINTER(C.c_double)), ("gld_lb", C.c_int32), ("gld_span", C.c_int32),
                ("rspd_pdf", C.POINTER(C.c_double))]


class RoundStats(C.Structure):
    _fields_ = [("sum", C.c_double), ("bchange", C.c_double), ("totnum", C.c_int64)]


class GibbsParams(C.Structure):
    _fields_ = [("M", C.c_int32), ("burnin", C.c_int32), ("gap", C.c_int32), ("n_chains", C.c_int32),
                ("chain_samples", C.POINTER(C.c_int32)), ("chain_seeds", C.POINTER(C.c_uint32)),
                ("n0", C.c_double), ("init_counts", C.POINTER(C.c_int32)), ("pseudo_counts", C.POINTER(C.c_double)),
                ("totc", C.c_double), ("eel", C.POINTER(C.c_double)), ("mw", C.POINTER(C.c_double)),
                ("n_genes", C.c_int32), ("gene_start", C.POINTER(C.c_int32))]


class GibbsOut(C.Structure):
    _fields_ = [("count_vectors", C.POINTER(C.c_int32)), ("sum_c", C.POINTER(C.c_double)),
                ("sum_c2", C.POINTER(C.c_double)), ("sum_tpm", C.POINTER(C.c_double)),
                ("sum_fpkm", C.POINTER(C.c_double)), ("sum_gene_c2", C.POINTER(C.c_double))]


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _p(a: Optional[np.ndarray], ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _arr(a, dtype) -> Optional[np.ndarray]:
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=dtype)


# every symbol include/rsem_b200.h declares (tests check that the .so exports all of them)
SYMBOLS = [
    "rsem_b200_version", "rsem_b200_last_error", "rsem_b200_device_count", "rsem_b200_ctx_create",
    "rsem_b200_ctx_destroy", "rsem_b200_ctx_set_stream", "rsem_b200_ctx_sync", "rsem_b200_ctx_device_bytes",
    "rsem_b200_comm_unique_id", "rsem_b200_comm_init", "rsem_b200_upload_hits", "rsem_b200_upload_conprb",
    "rsem_b200_download_conprb", "rsem_b200_adopt_device_matrix", "rsem_b200_upload_reads", "rsem_b200_upload_refs",
    "rsem_b200_set_model", "rsem_b200_calc_conprb", "rsem_b200_set_theta", "rsem_b200_get_theta",
    "rsem_b200_em_rounds", "rsem_b200_em_model_round", "rsem_b200_expected_weights", "rsem_b200_gibbs_upload",
    "rsem_b200_gibbs_run", "rsem_b200_launch_count", "rsem_b200_estep_timing", "rsem_b200_set_profiling",
    "rsem_b200_set_estep_variant", "rsem_b200_class_layout_info", "rsem_b200_shard_reads",
    "rsem_b200_estep_cta_times",
]


class Lib:
    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise RsemB200Error(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback)")
        self.path = path
        self.dll = C.CDLL(path)
        self.dll.rsem_b200_last_error.restype = C.c_char_p
        for s in SYMBOLS:
            getattr(self.dll, s)  # AttributeError if the library lacks a declared symbol

    def check(self, rc: int):
        if rc != 0:
            raise RsemB200Error(f"librsem_b200 error {rc}: {self.dll.rsem_b200_last_error().decode()}")

    def version(self) -> int:
        return self.dll.rsem_b200_version()

    def device_count(self) -> int:
        n = C.c_int(0)
        self.check(self.dll.rsem_b200_device_count(C.byref(n)))
        return n.value

    def shard_reads(self, row_ptr, n_shards: int):
        """-> [(first_read, last_read_exclusive)] per shard, the reference's thread-sharding rule (EM.cpp:135-157)"""
        row_ptr = _arr(row_ptr, np.uint64)
        bounds = np.zeros(n_shards + 1, np.uint64)
        self.check(self.dll.rsem_b200_shard_reads(C.c_uint64(len(row_ptr) - 1), _p(row_ptr, C.c_uint64), C.c_int32(n_shards),
                                                  _p(bounds, C.c_uint64)))
        return [(int(bounds[i]), int(bounds[i + 1])) for i in range(n_shards)]

    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self.check(self.dll.rsem_b200_comm_unique_id(buf))
        return buf.raw


_lib: Optional[Lib] = None


def load_library() -> Lib:
    global _lib
    if _lib is None:
        _lib = Lib()
    return _lib


class Context:
    """One GPU context (rsem_b200_ctx).  Methods are 1:1 with the C ABI."""

    def __init__(self, device: int = 0, lib: Optional[Lib] = None):
        self.lib = lib or load_library()
        self._h = C.c_void_p()
        self.lib.check(self.lib.dll.rsem_b200_ctx_create(C.c_int(device), C.byref(self._h)))
        self.N = self.H = self.M = 0
        self._keep = []

    def close(self):
        if self._h:
            self.lib.dll.rsem_b200_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing
    def set_stream(self, cuda_stream_ptr: int):
        self.lib.check(self.lib.dll.rsem_b200_ctx_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        self.lib.check(self.lib.dll.rsem_b200_ctx_sync(self._h))

    def device_bytes(self) -> int:
        b = C.c_uint64(0)
        self.lib.check(self.lib.dll.rsem_b200_ctx_device_bytes(self._h, C.byref(b)))
        return b.value

    def comm_init(self, unique_id: bytes, n_ranks: int, rank: int):
        self.lib.check(self.lib.dll.rsem_b200_comm_init(self._h, C.c_char_p(unique_id), n_ranks, rank))

    # ---- data
    def upload_hits(self, row_ptr, sid, M: int, pos=None, insertL=None):
        row_ptr = _arr(row_ptr, np.uint64)
        sid = _arr(sid, np.int32)
        pos = _arr(pos, np.int32)
        insertL = _arr(insertL, np.int32)
        N, H = len(row_ptr) - 1, len(sid)
        self.lib.check(self.lib.dll.rsem_b200_upload_hits(
            self._h, C.c_uint64(N), C.c_uint64(H), C.c_int32(M), _p(row_ptr, C.c_uint64), _p(sid, C.c_int32),
            _p(pos, C.c_int32), _p(insertL, C.c_int32)))
        self.N, self.H, self.M = N, H, M

    def upload_conprb(self, conprb, ncpv):
        conprb, ncpv = _arr(conprb, np.float64), _arr(ncpv, np.float64)
        assert len(conprb) == self.H and len(ncpv) == self.N
        self.lib.check(self.lib.dll.rsem_b200_upload_conprb(self._h, _dp(conprb), _dp(ncpv)))

    def upload_conprb_ptr(self, conprb_ptr: int, ncpv_ptr: int):
        """host pointers (e.g. pinned torch tensors' data_ptr())"""
        self.lib.check(self.lib.dll.rsem_b200_upload_conprb(self._h, C.c_void_p(conprb_ptr), C.c_void_p(ncpv_ptr)))

    def upload_hits_ptr(self, N: int, H: int, M: int, row_ptr_ptr: int, sid_ptr: int):
        self.lib.check(self.lib.dll.rsem_b200_upload_hits(
            self._h, C.c_uint64(N), C.c_uint64(H), C.c_int32(M), C.c_void_p(row_ptr_ptr), C.c_void_p(sid_ptr), None, None))
        self.N, self.H, self.M = N, H, M

    def download_conprb(self):
        conprb, ncpv = np.empty(self.H, np.float64), np.empty(self.N, np.float64)
        self.lib.check(self.lib.dll.rsem_b200_download_conprb(self._h, _dp(conprb), _dp(ncpv)))
        return conprb, ncpv

    def adopt_device_matrix(self, N: int, H: int, M: int, d_row_ptr: int, d_sid: int, d_conprb: int, d_ncpv: int):
        self.lib.check(self.lib.dll.rsem_b200_adopt_device_matrix(
            self._h, C.c_uint64(N), C.c_uint64(H), C.c_int32(M), C.c_void_p(d_row_ptr), C.c_void_p(d_sid),
            C.c_void_p(d_conprb), C.c_void_p(d_ncpv)))
        self.N, self.H, self.M = N, H, M

    def upload_reads(self, off1, base1, qual1, lowq, off2=None, base2=None, qual2=None):
        a = [_arr(off1, np.uint64), _arr(base1, np.uint8), _arr(qual1, np.uint8), _arr(off2, np.uint64),
             _arr(base2, np.uint8), _arr(qual2, np.uint8), _arr(lowq, np.uint8)]
        n_mates = 2 if off2 is not None else 1
        self.lib.check(self.lib.dll.rsem_b200_upload_reads(
            self._h, C.c_int32(n_mates), _p(a[0], C.c_uint64), _p(a[1], C.c_uint8), _p(a[2], C.c_uint8),
            _p(a[3], C.c_uint64), _p(a[4], C.c_uint8), _p(a[5], C.c_uint8), _p(a[6], C.c_uint8)))

    def upload_refs(self, seq_off, seq, full_len, tot_len, mask_off, mask_words):
        a = [_arr(seq_off, np.uint64), _arr(seq, np.uint8), _arr(full_len, np.int32), _arr(tot_len, np.int32),
             _arr(mask_off, np.uint64), _arr(mask_words, np.uint32)]
        M = len(a[2]) - 1
        self.lib.check(self.lib.dll.rsem_b200_upload_refs(
            self._h, C.c_int32(M), _p(a[0], C.c_uint64), _p(a[1], C.c_uint8), _p(a[2], C.c_int32), _p(a[3], C.c_int32),
            _p(a[4], C.c_uint64), _p(a[5], C.c_uint32)))

    def set_model(self, model: Model):
        self.lib.check(self.lib.dll.rsem_b200_set_model(self._h, C.byref(model)))

    # ---- compute
    def calc_conprb(self):
        self.lib.check(self.lib.dll.rsem_b200_calc_conprb(self._h))

    def set_theta(self, theta):
        theta = _arr(theta, np.float64)
        assert len(theta) == self.M + 1
        self.lib.check(self.lib.dll.rsem_b200_set_theta(self._h, _dp(theta)))

    def get_theta(self) -> np.ndarray:
        theta = np.empty(self.M + 1, np.float64)
        self.lib.check(self.lib.dll.rsem_b200_get_theta(self._h, _dp(theta)))
        return theta

    def em_rounds(self, first_round: int, max_rounds: int, min_round: int, max_round: int, n0: float):
        """returns (list of (sum, bchange, totnum), stopped)"""
        stats = (RoundStats * max(max_rounds, 1))()
        ran, stopped = C.c_int32(0), C.c_int32(0)
        self.lib.check(self.lib.dll.rsem_b200_em_rounds(
            self._h, C.c_int32(first_round), C.c_int32(max_rounds), C.c_int32(min_round), C.c_int32(max_round),
            C.c_double(n0), stats, C.byref(ran), C.byref(stopped)))
        return [(stats[i].sum, stats[i].bchange, stats[i].totnum) for i in range(ran.value)], bool(stopped.value)

    def em_model_round(self, n0: float, stats: ModelStats):
        rs = RoundStats()
        self.lib.check(self.lib.dll.rsem_b200_em_model_round(self._h, C.c_double(n0), C.byref(stats), C.byref(rs)))
        return rs.sum, rs.bchange, rs.totnum

    def expected_weights(self) -> np.ndarray:
        counts = np.empty(self.M + 1, np.float64)
        self.lib.check(self.lib.dll.rsem_b200_expected_weights(self._h, _dp(counts)))
        return counts

    def gibbs_upload(self, row_ptr, sid, conprb, M: int):
        row_ptr, sid, conprb = _arr(row_ptr, np.uint64), _arr(sid, np.int32), _arr(conprb, np.float64)
        self.lib.check(self.lib.dll.rsem_b200_gibbs_upload(
            self._h, C.c_uint64(len(row_ptr) - 1), C.c_uint64(len(sid)), C.c_int32(M), _p(row_ptr, C.c_uint64),
            _p(sid, C.c_int32), _dp(conprb)))

    def gibbs_run(self, params: GibbsParams, out: GibbsOut):
        self.lib.check(self.lib.dll.rsem_b200_gibbs_run(self._h, C.byref(params), C.byref(out)))

    # ---- instrumentation
    def launch_count(self) -> int:
        n = C.c_uint64(0)
        self.lib.check(self.lib.dll.rsem_b200_launch_count(self._h, C.byref(n)))
        return n.value

    def estep_timing(self, reset: bool = False):
        ms, n = C.c_double(0), C.c_uint64(0)
        self.lib.check(self.lib.dll.rsem_b200_estep_timing(self._h, C.byref(ms), C.byref(n), C.c_int32(int(reset))))
        return ms.value, n.value

    def set_profiling(self, on: bool):
        self.lib.check(self.lib.dll.rsem_b200_set_profiling(self._h, C.c_int32(int(on))))

    def class_layout_info(self) -> dict:
        out = (C.c_uint64 * 8)()
        self.lib.check(self.lib.dll.rsem_b200_class_layout_info(self._h, out))
        keys = ("built", "rows", "long_rows", "segments", "batches", "tiles", "vals", "ids")
        d = dict(zip(keys, (int(x) for x in out)))
        d["bytes_per_round"] = 8 * d["vals"] + 4 * d["ids"] + 16 * d["batches"]
        return d

    def estep_cta_times(self) -> np.ndarray:
        out = np.zeros(1024, np.uint64)
        n = C.c_int32(0)
        self.lib.check(self.lib.dll.rsem_b200_estep_cta_times(self._h, _p(out, C.c_uint64), C.c_int32(1024), C.byref(n)))
        return out[: n.value]

    def set_estep_variant(self, v: int):
        self.lib.check(self.lib.dll.rsem_b200_set_estep_variant(self._h, C.c_int32(v)))
