"""ctypes mirror of include/pedfac_cuda.h (the C ABI of libpedfac_cuda).

The library is the product; this module only loads it and marshals numpy arrays into the plain
C structs. There is no Python or CPU implementation of any entry point here: if the shared
library is missing, loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PF_OK = 0
PF_EINVAL, PF_ECUDA, PF_ENOMEM, PF_ENODEVICE, PF_ECYCLE = -1, -2, -3, -4, -5
PF_PROP_TRIO, PF_PROP_SELFING, PF_PROP_PRONG = 0, 1, 2

#: numpy dtype of `pf_proposal`
PROPOSAL_DTYPE = np.dtype([("kind", "<i4"), ("pa", "<i4"), ("ma", "<i4"), ("marr", "<i4")])

_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)


class PfConfig(C.Structure):
    _fields_ = [("n_loci", C.c_int32), ("locus_begin", C.c_int32), ("locus_end", C.c_int32),
                ("cap_indiv", C.c_int32), ("cap_marr", C.c_int32), ("n_chains", C.c_int32),
                ("device", C.c_int32), ("reserved", C.c_int32)]


class PfGraphView(C.Structure):
    _fields_ = [("n_indiv", C.c_int32), ("indiv", _i32p), ("indiv_cc", _i32p), ("indiv_founder", _u8p),
                ("n_marr", C.c_int32), ("marr", _i32p), ("marr_pa", _i32p), ("marr_ma", _i32p),
                ("marr_selfing", _u8p), ("marr_kid_begin", _i32p), ("marr_kids", _i32p),
                ("n_cc", C.c_int32)]


class PedfacError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def repo_root() -> Path:
    return Path(__file__).resolve().parent.parent


def cuda_library_path() -> Path:
    return Path(os.environ.get("PEDFAC_CUDA_LIB", repo_root() / "pedfac_b200" / "lib" / "libpedfac_cuda.so"))


#: every symbol include/pedfac_cuda.h declares (checked by tests/test_abi.py against the header)
EXPORTS = [
    "last_error", "create", "destroy", "set_stream", "set_locus_params", "set_genotypes_hard",
    "set_genotypes_soft", "set_genotypes_soft_f64", "set_genotypes_soft_dec16",
    "set_genotypes_unobserved", "sweep", "eval", "sweep_dev", "eval_dev", "sweep_batch",
    "eval_batch", "sweep_batch_dev", "eval_batch_dev", "sweep_eval", "eval_swaps", "eval_swaps_dev", "prefilter", "prefilter_scores_dev", "prefilter_topk_dev", "get_stub", "get_prong", "launch_count", "io_bytes", "comm_local_handle", "comm_connect", "comm_reduce_count", "comm_set_async", "comm_fence", "sweep_eval_dev", "sweep_eval_swaps", "deferred_results", "eval_swaps_batch", "profile", "profile_read", "plan_stats",
]


def bind(lib: C.CDLL, prefix: str) -> None:
    """Declare argument/return types of the `<prefix>*` entry points that `lib` exports."""
    vp = C.c_void_p

    def sig(name, restype, *argtypes):
        fn = getattr(lib, prefix + name, None)
        if fn is not None:
            fn.restype = restype
            fn.argtypes = list(argtypes)

    sig("last_error", C.c_char_p)
    sig("create", C.c_int, C.POINTER(vp), C.POINTER(PfConfig))
    sig("destroy", C.c_int, vp)
    sig("set_stream", C.c_int, vp, vp)
    sig("set_locus_params", C.c_int, vp, _f64p, _f64p)
    sig("set_genotypes_hard", C.c_int, vp, C.c_int32, _u8p)
    sig("set_genotypes_soft", C.c_int, vp, C.c_int32, C.POINTER(C.c_float))
    sig("set_genotypes_soft_f64", C.c_int, vp, C.c_int32, _f64p)
    sig("set_genotypes_soft_dec16", C.c_int, vp, C.c_int32, C.POINTER(C.c_uint16), C.c_uint32)
    sig("set_genotypes_unobserved", C.c_int, vp, C.c_int32)
    sig("sweep", C.c_int, vp, C.c_int32, C.POINTER(PfGraphView), _f64p)
    sig("eval", C.c_int, vp, C.c_int32, C.c_int32, C.c_int32, vp, _f64p)
    sig("sweep_dev", C.c_int, vp, C.c_int32, C.POINTER(PfGraphView), vp)
    sig("eval_dev", C.c_int, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp)
    sig("sweep_batch", C.c_int, vp, C.c_int32, _i32p, C.POINTER(PfGraphView), _f64p)
    sig("eval_batch", C.c_int, vp, C.c_int32, _i32p, _i32p, _i32p, vp, _f64p)
    sig("sweep_batch_dev", C.c_int, vp, C.c_int32, _i32p, C.POINTER(PfGraphView), vp)
    sig("eval_batch_dev", C.c_int, vp, C.c_int32, _i32p, _i32p, _i32p, vp, vp)
    sig("sweep_eval", C.c_int, vp, C.c_int32, C.POINTER(PfGraphView), _f64p, C.c_int32, C.c_int32, vp, _f64p)
    sig("eval_swaps", C.c_int, vp, C.c_int32, C.POINTER(PfGraphView), C.c_int32, vp, _f64p)
    sig("eval_swaps_dev", C.c_int, vp, C.c_int32, C.POINTER(PfGraphView), C.c_int32, vp, vp)
    sig("prefilter", C.c_int, vp, C.c_int32, C.c_int32, _i32p, _u8p, C.c_int32, _i32p, _f64p,
        C.c_int32, _i32p, _f64p)
    sig("prefilter_scores_dev", C.c_int, vp, C.c_int32, C.c_int32, _i32p, _u8p, C.c_int32, _i32p, _f64p, vp)
    sig("prefilter_topk_dev", C.c_int, vp, C.c_int32, C.c_int32, _i32p, C.c_int32, _i32p, vp, C.c_int32, _i32p, _f64p)
    sig("get_stub", C.c_int, vp, C.c_int32, C.c_int32, _f64p)
    sig("get_prong", C.c_int, vp, C.c_int32, C.c_int32, _f64p)
    sig("launch_count", C.c_int64, vp)
    sig("io_bytes", C.c_int, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64))
    sig("comm_local_handle", C.c_int, vp, C.c_int32, vp)
    sig("comm_connect", C.c_int, vp, C.c_int32, C.c_int32, vp)
    sig("comm_reduce_count", C.c_int64, vp)
    sig("comm_set_async", C.c_int, vp, C.c_int32)
    sig("comm_fence", C.c_int, vp)
    sig("sweep_eval_swaps", C.c_int, vp, C.c_int32, C.POINTER(PfGraphView), _f64p, C.c_int32, C.c_int32, vp, _f64p, C.POINTER(PfGraphView), C.c_int32, vp)
    sig("eval_swaps_batch", C.c_int, vp, C.c_int32, _i32p, C.POINTER(PfGraphView), _i32p, vp, _f64p)
    sig("deferred_results", C.c_int, vp, C.c_int32, _f64p)
    sig("sweep_eval_dev", C.c_int, vp, C.c_int32, C.POINTER(PfGraphView), C.c_int32, C.c_int32, vp, vp)
    sig("plan_stats", C.c_int, vp, _i32p)
    sig("profile", C.c_int, vp, C.c_int32)
    sig("profile_read", C.c_int, vp, _f64p, C.POINTER(C.c_int64))


_cuda_lib = None


def load_cuda_library() -> C.CDLL:
    """Load libpedfac_cuda.so (built by `__graft_entry__.build()` / `make -C pedfac_b200/csrc`)."""
    global _cuda_lib
    if _cuda_lib is None:
        path = cuda_library_path()
        if not path.exists():
            raise FileNotFoundError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "pedfac_b200 has no CPU fallback.")
        _cuda_lib = C.CDLL(str(path))
        bind(_cuda_lib, "pf_")
    return _cuda_lib


def as_i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


class GraphView:
    """Host arrays of one `pf_graph_view` plus the ctypes struct pointing into them."""

    def __init__(self, indiv, indiv_cc, indiv_founder, marr, marr_pa, marr_ma, marr_selfing,
                 marr_kid_begin, marr_kids, n_cc):
        self.indiv = as_i32(indiv)
        self.indiv_cc = as_i32(indiv_cc)
        self.indiv_founder = np.ascontiguousarray(indiv_founder, dtype=np.uint8)
        self.marr = as_i32(marr)
        self.marr_pa = as_i32(marr_pa)
        self.marr_ma = as_i32(marr_ma)
        self.marr_selfing = np.ascontiguousarray(marr_selfing, dtype=np.uint8)
        self.marr_kid_begin = as_i32(marr_kid_begin)
        self.marr_kids = as_i32(marr_kids)
        self.n_cc = int(n_cc)
        assert len(self.indiv) == len(self.indiv_cc) == len(self.indiv_founder)
        assert len(self.marr) == len(self.marr_pa) == len(self.marr_ma) == len(self.marr_selfing)
        assert len(self.marr_kid_begin) == len(self.marr) + 1
        self.struct = PfGraphView(
            len(self.indiv), ptr(self.indiv, _i32p), ptr(self.indiv_cc, _i32p), ptr(self.indiv_founder, _u8p),
            len(self.marr), ptr(self.marr, _i32p), ptr(self.marr_pa, _i32p), ptr(self.marr_ma, _i32p),
            ptr(self.marr_selfing, _u8p), ptr(self.marr_kid_begin, _i32p), ptr(self.marr_kids, _i32p),
            self.n_cc)


def make_proposals(rows) -> np.ndarray:
    """rows: iterable of (kind, pa, ma, marr) -> contiguous `pf_proposal` array."""
    arr = np.array([tuple(r) for r in rows], dtype=PROPOSAL_DTYPE) if not isinstance(rows, np.ndarray) else rows
    return np.ascontiguousarray(arr, dtype=PROPOSAL_DTYPE)


#: `pf_swap` of the header
SWAP_DTYPE = np.dtype([("kind", "<i4"), ("move_kids", "<i4"), ("x", "<i4"), ("z", "<i4"), ("marr_z", "<i4"),
                       ("marr_x", "<i4"), ("prev_pa", "<i4"), ("prev_ma", "<i4"), ("prev_marr", "<i4"),
                       ("reserved", "<i4")])


def make_swaps(rows) -> np.ndarray:
    """rows: iterable of (kind, move_kids, x, z, marr_z, marr_x, prev_pa, prev_ma, prev_marr)."""
    if isinstance(rows, np.ndarray) and rows.dtype == SWAP_DTYPE:
        return np.ascontiguousarray(rows)
    return np.array([tuple(r) + (0,) for r in rows], dtype=SWAP_DTYPE)


class CEngine:
    """Thin object wrapper over one engine handle of a library exporting `<prefix>*`.

    `pedfac_b200.Engine` is this class bound to libpedfac_cuda (`pf_`). The test suite binds the
    same wrapper to its CPU checker library so that both are driven identically.
    """

    def __init__(self, lib: C.CDLL, prefix: str, n_loci: int, cap_indiv: int, cap_marr: int,
                 locus_range=None, n_chains: int = 1, device: int = 0):
        self._lib, self._prefix = lib, prefix
        self.n_loci = int(n_loci)
        self.locus_begin, self.locus_end = (0, self.n_loci) if locus_range is None else map(int, locus_range)
        self.n_local = self.locus_end - self.locus_begin
        self.cap_indiv, self.cap_marr, self.n_chains = int(cap_indiv), int(cap_marr), int(n_chains)
        cfg = PfConfig(self.n_loci, self.locus_begin, self.locus_end, self.cap_indiv, self.cap_marr,
                       self.n_chains, int(device), 0)
        self._h = C.c_void_p()
        self._check(self._fn("create")(C.byref(self._h), C.byref(cfg)))

    def _fn(self, name):
        return getattr(self._lib, self._prefix + name)

    def _check(self, rc: int):
        if rc != PF_OK:
            raise PedfacError(rc, self._fn("last_error")().decode(errors="replace"))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._fn("destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- inputs -------------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        self._check(self._fn("set_stream")(self._h, C.c_void_p(cuda_stream)))

    def set_locus_params(self, eps, afreq):
        eps = np.ascontiguousarray(eps, dtype=np.float64)
        afreq = np.ascontiguousarray(afreq, dtype=np.float64)
        assert eps.shape == afreq.shape == (self.n_loci,)
        self._check(self._fn("set_locus_params")(self._h, ptr(eps, _f64p), ptr(afreq, _f64p)))

    def set_hard(self, slot: int, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        assert codes.shape == (self.n_loci,)
        self._check(self._fn("set_genotypes_hard")(self._h, slot, ptr(codes, _u8p)))

    def set_soft(self, slot: int, lik, fmt: str = "f32", denom: int = 10000):
        if fmt == "f32":
            a = np.ascontiguousarray(lik, dtype=np.float32)
            assert a.shape == (self.n_loci, 3)
            self._check(self._fn("set_genotypes_soft")(self._h, slot, ptr(a, C.POINTER(C.c_float))))
        elif fmt == "f64":
            a = np.ascontiguousarray(lik, dtype=np.float64)
            assert a.shape == (self.n_loci, 3)
            self._check(self._fn("set_genotypes_soft_f64")(self._h, slot, ptr(a, _f64p)))
        elif fmt == "dec16":
            a = np.ascontiguousarray(lik, dtype=np.uint16)
            assert a.shape == (self.n_loci, 3)
            self._check(self._fn("set_genotypes_soft_dec16")(self._h, slot, ptr(a, C.POINTER(C.c_uint16)), denom))
        else:
            raise ValueError(fmt)

    def set_unobserved(self, slot: int):
        self._check(self._fn("set_genotypes_unobserved")(self._h, slot))

    # -- the hot path --------------------------------------------------------------------------
    def sweep(self, view: GraphView, chain: int = 0) -> np.ndarray:
        out = np.empty(max(view.n_cc, 1), dtype=np.float64)
        self._check(self._fn("sweep")(self._h, chain, C.byref(view.struct), ptr(out, _f64p)))
        return out[:view.n_cc]

    def eval(self, kid: int, props, chain: int = 0) -> np.ndarray:
        props = make_proposals(props)
        out = np.empty(max(len(props), 1), dtype=np.float64)
        self._check(self._fn("eval")(self._h, chain, kid, len(props), C.c_void_p(props.ctypes.data), ptr(out, _f64p)))
        return out[:len(props)]

    def sweep_eval(self, view: GraphView, kid: int, props, chain: int = 0):
        """`pf_sweep_eval`: (ll_cc, lsum) of a sweep and an eval submitted together."""
        props = make_proposals(props)
        ll = np.empty(max(view.n_cc, 1), dtype=np.float64)
        out = np.empty(max(len(props), 1), dtype=np.float64)
        self._check(self._fn("sweep_eval")(self._h, chain, C.byref(view.struct), ptr(ll, _f64p), kid, len(props),
                                           C.c_void_p(props.ctypes.data), ptr(out, _f64p)))
        return ll[:view.n_cc], out[:len(props)]

    def sweep_eval_swaps(self, view: GraphView, kid: int, props, swap_view: GraphView, swaps, chain: int = 0):
        """`pf_sweep_eval_swaps`: (ll_cc, lsum) now; the swap values are deferred (`deferred_results`)."""
        props = make_proposals(props)
        swaps = make_swaps(swaps)
        ll = np.empty(max(view.n_cc, 1), dtype=np.float64)
        out = np.empty(max(len(props), 1), dtype=np.float64)
        self._check(self._fn("sweep_eval_swaps")(self._h, chain, C.byref(view.struct), ptr(ll, _f64p), kid, len(props),
                                                 C.c_void_p(props.ctypes.data), ptr(out, _f64p),
                                                 C.byref(swap_view.struct), len(swaps), C.c_void_p(swaps.ctypes.data)))
        return ll[:view.n_cc], out[:len(props)]

    def deferred_results(self, n: int) -> np.ndarray:
        out = np.empty(max(n, 1), dtype=np.float64)
        self._check(self._fn("deferred_results")(self._h, n, ptr(out, _f64p)))
        return out[:n]

    def eval_swaps(self, view: GraphView, swaps, chain: int = 0) -> np.ndarray:
        swaps = make_swaps(swaps)
        out = np.empty(max(len(swaps), 1), dtype=np.float64)
        self._check(self._fn("eval_swaps")(self._h, chain, C.byref(view.struct), len(swaps),
                                           C.c_void_p(swaps.ctypes.data), ptr(out, _f64p)))
        return out[:len(swaps)]

    def eval_swaps_batch(self, chains, views, swap_lists) -> np.ndarray:
        """`pf_eval_swaps_batch`: one list of swaps per (chain, view); values concatenated in list order."""
        chains = as_i32(chains)
        arr = (PfGraphView * len(views))(*[v.struct for v in views])
        lists = [make_swaps(s) for s in swap_lists]
        begin = as_i32(np.concatenate([[0], np.cumsum([len(s) for s in lists])]))
        allsw = np.concatenate(lists) if lists else make_swaps([])
        out = np.empty(max(len(allsw), 1), dtype=np.float64)
        self._check(self._fn("eval_swaps_batch")(self._h, len(views), ptr(chains, _i32p), arr, ptr(begin, _i32p),
                                                 C.c_void_p(allsw.ctypes.data), ptr(out, _f64p)))
        return out[:len(allsw)]

    def eval_swaps_dev(self, view: GraphView, swaps: np.ndarray, out_dev_ptr: int, chain: int = 0):
        self._check(self._fn("eval_swaps_dev")(self._h, chain, C.byref(view.struct), len(swaps),
                                               C.c_void_p(swaps.ctypes.data), C.c_void_p(out_dev_ptr)))

    def sweep_dev(self, view: GraphView, out_dev_ptr: int, chain: int = 0):
        self._check(self._fn("sweep_dev")(self._h, chain, C.byref(view.struct), C.c_void_p(out_dev_ptr)))

    def eval_dev(self, kid: int, props: np.ndarray, out_dev_ptr: int, chain: int = 0):
        self._check(self._fn("eval_dev")(self._h, chain, kid, len(props), C.c_void_p(props.ctypes.data),
                                         C.c_void_p(out_dev_ptr)))

    def sweep_batch(self, chains, views) -> np.ndarray:
        chains = as_i32(chains)
        arr = (PfGraphView * len(views))(*[v.struct for v in views])
        tot = sum(v.n_cc for v in views)
        out = np.empty(max(tot, 1), dtype=np.float64)
        self._check(self._fn("sweep_batch")(self._h, len(views), ptr(chains, _i32p), arr, ptr(out, _f64p)))
        return out[:tot]

    def eval_batch(self, chains, kids, prop_begin, props) -> np.ndarray:
        chains, kids, prop_begin = as_i32(chains), as_i32(kids), as_i32(prop_begin)
        props = make_proposals(props)
        assert len(prop_begin) == len(kids) + 1 and prop_begin[-1] == len(props)
        out = np.empty(max(len(props), 1), dtype=np.float64)
        self._check(self._fn("eval_batch")(self._h, len(kids), ptr(chains, _i32p), ptr(kids, _i32p),
                                           ptr(prop_begin, _i32p), C.c_void_p(props.ctypes.data), ptr(out, _f64p)))
        return out[:len(props)]

    def sweep_batch_dev(self, chains: np.ndarray, view_array, n: int, out_dev_ptr: int):
        """chains: int32 array; view_array: ctypes array of PfGraphView (kept alive by the caller)."""
        self._check(self._fn("sweep_batch_dev")(self._h, n, ptr(chains, _i32p), view_array, C.c_void_p(out_dev_ptr)))

    def eval_batch_dev(self, chains: np.ndarray, kids: np.ndarray, prop_begin: np.ndarray, props: np.ndarray, out_dev_ptr: int):
        self._check(self._fn("eval_batch_dev")(self._h, len(kids), ptr(chains, _i32p), ptr(kids, _i32p),
                                               ptr(prop_begin, _i32p), C.c_void_p(props.ctypes.data), C.c_void_p(out_dev_ptr)))

    def prefilter(self, kids, as_father, cands, cand_bias=None, top_k: int = 15, chain: int = 0):
        kids, cands = as_i32(kids), as_i32(cands)
        as_father = np.ascontiguousarray(as_father, dtype=np.uint8)
        bias = None if cand_bias is None else np.ascontiguousarray(cand_bias, dtype=np.float64)
        out_slot = np.empty((len(kids), top_k), dtype=np.int32)
        out_score = np.empty((len(kids), top_k), dtype=np.float64)
        self._check(self._fn("prefilter")(
            self._h, chain, len(kids), ptr(kids, _i32p), ptr(as_father, _u8p), len(cands), ptr(cands, _i32p),
            None if bias is None else ptr(bias, _f64p), top_k, ptr(out_slot, _i32p), ptr(out_score, _f64p)))
        return out_slot, out_score

    # -- read-back -----------------------------------------------------------------------------
    def prefilter_scores_dev(self, kids, as_father, cands, cand_bias, scores_ptr: int, chain: int = 0):
        """Dense partial scores [n_kids][n_cands] of this engine's loci into the buffer at `scores_ptr`
        (device memory for the CUDA engine)."""
        kids, cands = as_i32(kids), as_i32(cands)
        role = np.ascontiguousarray(as_father, dtype=np.uint8)
        bias = None if cand_bias is None else np.ascontiguousarray(cand_bias, dtype=np.float64)
        self._check(self._fn("prefilter_scores_dev")(self._h, chain, len(kids), ptr(kids, _i32p), ptr(role, _u8p), len(cands),
                                                     ptr(cands, _i32p), None if bias is None else ptr(bias, _f64p),
                                                     C.c_void_p(scores_ptr)))

    def prefilter_topk_dev(self, kids, cands, scores_ptr: int, top_k: int = 15, chain: int = 0):
        """Ranks dense (summed) scores in place; returns (slot[n_kids][top_k], score[n_kids][top_k])."""
        kids, cands = as_i32(kids), as_i32(cands)
        slot = np.full((len(kids), top_k), -1, dtype=np.int32)
        score = np.full((len(kids), top_k), -np.inf, dtype=np.float64)
        self._check(self._fn("prefilter_topk_dev")(self._h, chain, len(kids), ptr(kids, _i32p), len(cands), ptr(cands, _i32p),
                                                   C.c_void_p(scores_ptr), top_k, ptr(slot, _i32p), ptr(score, _f64p)))
        return slot, score

    def get_stub(self, slot: int, chain: int = 0) -> np.ndarray:
        out = np.empty((self.n_local, 3), dtype=np.float64)
        self._check(self._fn("get_stub")(self._h, chain, slot, ptr(out, _f64p)))
        return out

    def get_prong(self, marr: int, chain: int = 0) -> np.ndarray:
        out = np.empty((self.n_local, 3), dtype=np.float64)
        self._check(self._fn("get_prong")(self._h, chain, marr, ptr(out, _f64p)))
        return out

    def plan_stats(self) -> dict:
        out = np.zeros(8, dtype=np.int32)
        self._check(self._fn("plan_stats")(self._h, ptr(out, _i32p)))
        keys = ["levels", "groups", "tasks", "smem_bytes", "rows", "rows_smem", "up_levels", "tiles"]
        return dict(zip(keys, out.tolist()))

    def profile(self, enable: bool):
        self._check(self._fn("profile")(self._h, 1 if enable else 0))

    def profile_read(self):
        """-> ({kind: ms}, {kind: launches}) for kinds sweep / eval / reduce / prefilter."""
        ms = np.zeros(4, dtype=np.float64)
        cnt = np.zeros(4, dtype=np.int64)
        self._check(self._fn("profile_read")(self._h, ptr(ms, _f64p), ptr(cnt, C.POINTER(C.c_int64))))
        names = ["sweep", "eval", "reduce", "prefilter"]
        return dict(zip(names, ms.tolist())), dict(zip(names, cnt.tolist()))

    def launch_count(self) -> int:
        fn = getattr(self._lib, self._prefix + "launch_count", None)
        return int(fn(self._h)) if fn is not None else 0

    def io_bytes(self):
        h2d, d2h = C.c_uint64(0), C.c_uint64(0)
        self._check(self._fn("io_bytes")(self._h, C.byref(h2d), C.byref(d2h)))
        return int(h2d.value), int(d2h.value)

    # ---- locus shards on several GPUs (pf_comm_*): the library sums over the shards itself ----
    def comm_local_handle(self, world: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self._fn("comm_local_handle")(self._h, world, buf))
        return buf.raw

    def comm_connect(self, rank: int, world: int, handles: bytes) -> None:
        assert len(handles) == 64 * world
        self._check(self._fn("comm_connect")(self._h, rank, world, C.create_string_buffer(handles, len(handles))))

    def comm_set_async(self, enable: bool) -> None:
        self._check(self._fn("comm_set_async")(self._h, 1 if enable else 0))

    def comm_fence(self) -> None:
        self._check(self._fn("comm_fence")(self._h))

    def sweep_eval_dev(self, view, kid: int, props, out_ptr: int, chain: int = 0) -> None:
        """pf_sweep_eval_dev: [ll_cc | lsum] left at device address `out_ptr` (props: make_proposals array)."""
        self._check(self._fn("sweep_eval_dev")(self._h, chain, C.byref(view.struct), kid, len(props), C.c_void_p(props.ctypes.data), C.c_void_p(out_ptr)))

    def comm_reduce_count(self) -> int:
        return int(self._fn("comm_reduce_count")(self._h))

    def comm_init_from_torch(self, dist, rank: int, world: int) -> None:
        """Exchange the mailbox handles through an initialised torch.distributed group (plumbing only)."""
        mine = self.comm_local_handle(world)
        allh = [None] * world
        dist.all_gather_object(allh, mine)
        self.comm_connect(rank, world, b"".join(allh))


def compose_swap_ll(ll_tot, lsum, ll_cc_x, ll_cc_pa, use_ma, ll_cc_ma):
    """`pf_compose_swap_ll` of the header (_calculateBundlell's scalar order)."""
    a = (-ll_cc_x) + (-ll_cc_pa)
    return (a + float(-(1 if use_ma else 0)) * ll_cc_ma) + (ll_tot + lsum)


def compose_ll(ll_tot, lsum, ll_cc_kid, use_pa, ll_cc_pa, use_ma, ll_cc_ma):
    """`pf_compose_ll` of the header: the reference's scalar bookkeeping around the data term."""
    a = -ll_cc_kid
    a = a + (-ll_cc_pa if use_pa else 0.0)
    a = a + (-ll_cc_ma if use_ma else 0.0)
    return a + (ll_tot + lsum)
