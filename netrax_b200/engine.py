"""Python binding of the product: libnetrax_b200.so (C++ host mirror of NetRAX's likelihood API) on top of
libnrx_engine.so (sm_100a CUDA kernels).  There is no CPU fallback: importing works anywhere, but creating an
engine without the compiled extension or without a CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Optional, Sequence

import numpy as np

from ._capi import AVERAGE, BEST, LINKED, UNLINKED, FlatAPI, LikelihoodEngine, LikelihoodError, Partition
from .network_io import NetworkDesc

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_SO = os.path.join(HERE, "libnetrax_b200.so")
ENGINE_SO = os.path.join(HERE, "libnrx_engine.so")

_api: Optional[FlatAPI] = None


def load() -> FlatAPI:
    """Load the CUDA extension.  Fails loudly when it has not been built (run __graft_entry__.build())."""
    global _api
    if _api is None:
        for so in (ENGINE_SO, HOST_SO):
            if not os.path.exists(so):
                raise ImportError(f"netrax_b200: compiled extension {so} is missing — build it with "
                                  f"`make -C netrax_b200/csrc` (python -c 'import __graft_entry__ as g; g.build()'); "
                                  f"there is no CPU fallback")
        C.CDLL(ENGINE_SO, mode=C.RTLD_GLOBAL)
        lib = C.CDLL(HOST_SO)
        a = FlatAPI(lib, "nrxh_")
        lib.nrxh_set_eigen.restype = C.c_int
        lib.nrxh_set_eigen.argtypes = [C.c_void_p, C.c_uint] + [np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")] * 3
        lib.nrxh_eigen_decompose.restype = C.c_int
        lib.nrxh_eigen_decompose.argtypes = [C.c_uint] + [np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")] * 5
        lib.nrxh_comm_get_unique_id.restype = C.c_int
        lib.nrxh_comm_get_unique_id.argtypes = [C.c_char_p]
        lib.nrxh_comm_init.restype = C.c_int
        lib.nrxh_comm_init.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        lib.nrxh_launch_count.restype = C.c_ulonglong
        lib.nrxh_launch_count.argtypes = [C.c_void_p]
        lib.nrxh_num_slots.restype = C.c_uint
        lib.nrxh_num_slots.argtypes = [C.c_void_p]
        lib.nrxh_profile_enable.restype = C.c_int
        lib.nrxh_profile_enable.argtypes = [C.c_void_p, C.c_int]
        lib.nrxh_profile_read.restype = C.c_int
        lib.nrxh_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_double)] + [C.POINTER(C.c_ulonglong)] * 3
        lib.nrxh_profile_read_kind.restype = C.c_int
        lib.nrxh_profile_read_kind.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)] + [C.POINTER(C.c_ulonglong)] * 4
        lib.nrxh_persite_lnl.restype = C.c_int
        lib.nrxh_persite_lnl.argtypes = [C.c_void_p, C.c_uint, np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS"), C.c_uint]
        lib.nrxh_engine.restype = C.c_void_p
        lib.nrxh_engine.argtypes = [C.c_void_p]
        lib.nrxh_upload_alignment_u8.restype = C.c_int
        lib.nrxh_upload_alignment_u8.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]
        lib.nrxh_stage_alignment_u8.restype = C.c_int
        lib.nrxh_stage_alignment_u8.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]
        lib.nrxh_commit_staged_alignment.restype = C.c_int
        lib.nrxh_commit_staged_alignment.argtypes = [C.c_void_p]
        lib.nrxh_upload_alignment_codes.restype = C.c_int
        lib.nrxh_upload_alignment_codes.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS"), C.c_uint, C.c_void_p]
        lib.nrxh_compute_loglikelihood_batch.restype = C.c_int
        lib.nrxh_compute_loglikelihood_batch.argtypes = [C.POINTER(C.c_void_p), C.c_uint, C.c_int, C.c_int,
                                                         np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")]
        lib.nrxh_timer_start.restype = C.c_int
        lib.nrxh_timer_start.argtypes = [C.c_void_p]
        lib.nrxh_timer_stop.restype = C.c_int
        lib.nrxh_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        _api = a
    return _api


def eigen_decompose(states: int, freqs, subst):
    """The host library's own eigendecomposition (no device needed): (eigenvecs [S][SP], inv_eigenvecs [S][SP], eigenvals [SP])."""
    api = load()
    sp = (states + 3) & ~3
    ev, iev, evals = np.zeros(states * sp), np.zeros(states * sp), np.zeros(sp)
    api.check(api.lib.nrxh_eigen_decompose(states, np.ascontiguousarray(freqs, np.float64), np.ascontiguousarray(subst, np.float64),
                                           ev, iev, evals))
    return ev.reshape(states, sp), iev.reshape(states, sp), evals


def comm_unique_id() -> bytes:
    """128-byte NCCL unique id (call on rank 0, hand to the other ranks through the launcher's store)."""
    api = load()
    buf = C.create_string_buffer(128)
    api.check(api.lib.nrxh_comm_get_unique_id(buf))
    return buf.raw


def device_count() -> int:
    lib = C.CDLL(ENGINE_SO, mode=C.RTLD_GLOBAL)
    lib.nrx_device_count.restype = C.c_int
    return lib.nrx_device_count()


def compute_loglikelihood_batch(engines: Sequence["NetraxB200"], incremental: int = 1, update_pmatrices: int = 1) -> np.ndarray:
    """Batched scoring of candidate networks: every engine = its own CUDA stream; all evaluations are
    enqueued before the first result is collected.  out[i] == engines[i].computeLoglikelihood(incremental, update_pmatrices)."""
    api = load()
    hs = (C.c_void_p * len(engines))(*[e.h for e in engines])
    out = np.zeros(len(engines))
    api.check(api.lib.nrxh_compute_loglikelihood_batch(hs, len(engines), incremental, update_pmatrices, out))
    return out


class NetraxB200(LikelihoodEngine):
    """AnnotatedNetwork + likelihood state on one B200.  `reduce` (optional) is the parallel_reduce_cb of the
    reference: a callable that sums a float64 numpy array in place across all site shards (ranks)."""

    def __init__(self, net: NetworkDesc, partitions: Sequence[Partition], variant: int = AVERAGE, linkage: int = LINKED,
                 device: int = 0, plan_cache: bool = True, partition_brlens=None,
                 reduce: Optional[Callable[[np.ndarray], None]] = None, comm: Optional[tuple] = None):
        """`comm` = (unique_id_bytes, rank, nranks): attach an NCCL communicator so that the engine all-reduces its
        per-tree / per-pair partition sums on the device (site sharding across GPUs); `reduce` is then unused."""
        api = load()
        # LikelihoodEngine.__init__ runs nrxh_init; the callback must be installed before the first evaluation only
        super().__init__(api, net, partitions, variant=variant, linkage=linkage,
                         backend=f"device={device};plan_cache={1 if plan_cache else 0}", partition_brlens=partition_brlens)
        if reduce is not None:
            self.set_reduce(reduce)
        if comm is not None:
            uid, rank, nranks = comm
            api.check(api.lib.nrxh_comm_init(self.h, bytes(uid), int(rank), int(nranks)))

    def set_eigen(self, p: int, eigenvecs, inv_eigenvecs, eigenvals):
        """Test hook: inject an eigen-decomposition (e.g. the reference's) instead of the host Jacobi solver's."""
        self.api.check(self.api.lib.nrxh_set_eigen(self.h, p, np.ascontiguousarray(eigenvecs), np.ascontiguousarray(inv_eigenvecs),
                                                   np.ascontiguousarray(eigenvals)))

    def launch_count(self) -> int:
        return int(self.api.lib.nrxh_launch_count(self.h))

    def num_slots(self) -> int:
        return int(self.api.lib.nrxh_num_slots(self.h))

    def profile_enable(self, on: bool = True):
        self.api.check(self.api.lib.nrxh_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        ms = C.c_double()
        l, u, b = C.c_ulonglong(), C.c_ulonglong(), C.c_ulonglong()
        self.api.check(self.api.lib.nrxh_profile_read(self.h, C.byref(ms), C.byref(l), C.byref(u), C.byref(b)))
        return {"clv_ms": ms.value, "clv_launches": l.value, "clv_site_updates": u.value, "clv_bytes": b.value}

    PROF_KINDS = ("K2_clv_update", "K1_pmatrix", "K3_tree_lnl", "K3F_term_lnl_sum", "K4_edge_lnl", "K5_sumtable", "K6_derivatives",
                  "reduce_partials", "slot_copy", "K45_edge_lnl_sumtable")  # NRX_PROF_* of include/nrx_engine.h

    def profile_read_all(self):
        """{kernel family: {ms, launches, units, bytes, compulsory_bytes}} since profile_enable(True) (CUDA events on the engine
        stream).  bytes = algorithmic (SURVEY §8d per op), compulsory_bytes = each distinct operand once per launch."""
        out = {}
        for kind, name in enumerate(self.PROF_KINDS):
            ms = C.c_double()
            l, u, b, cb = C.c_ulonglong(), C.c_ulonglong(), C.c_ulonglong(), C.c_ulonglong()
            self.api.check(self.api.lib.nrxh_profile_read_kind(self.h, kind, C.byref(ms), C.byref(l), C.byref(u), C.byref(b), C.byref(cb)))
            out[name] = {"ms": ms.value, "launches": l.value, "units": u.value, "bytes": b.value, "compulsory_bytes": cb.value}
        return out

    def persite_lnl(self, tree: int) -> np.ndarray:
        stride = max(p.sites for p in self.partitions)
        out = np.zeros((self.P, stride))
        self.api.check(self.api.lib.nrxh_persite_lnl(self.h, tree, out.reshape(-1), stride))
        return out

    def upload_alignment_u8(self, p: int, tipchars_ptr: int, weights_ptr: int = 0):
        """Host -> device re-upload of one partition's alignment slice (pointers to pinned or pageable host memory)."""
        self.api.check(self.api.lib.nrxh_upload_alignment_u8(self.h, p, C.c_void_p(tipchars_ptr), C.c_void_p(weights_ptr) if weights_ptr else None))

    def stage_alignment_u8(self, p: int, tipchars_ptr: int, weights_ptr: int = 0):
        """Double-buffered upload (4-state partitions): the copy runs on a separate stream and overlaps with every evaluation enqueued
        after this call; commit_staged_alignment() makes it the live alignment."""
        self.api.check(self.api.lib.nrxh_stage_alignment_u8(self.h, p, C.c_void_p(tipchars_ptr) if tipchars_ptr else None, C.c_void_p(weights_ptr) if weights_ptr else None))

    def commit_staged_alignment(self):
        self.api.check(self.api.lib.nrxh_commit_staged_alignment(self.h))

    def upload_alignment_codes(self, p: int, codes_ptr: int, tipmap: np.ndarray, weights_ptr: int = 0):
        """Any alphabet: 1-byte codes [tips][patterns] (pointer to host memory) + tipmap[code] = state-set mask.  Asynchronous: the
        buffers must stay alive until the next evaluation has returned; an illegal code fails that evaluation."""
        tm = np.ascontiguousarray(tipmap, np.uint32)
        self.api.check(self.api.lib.nrxh_upload_alignment_codes(self.h, p, C.c_void_p(codes_ptr), tm, len(tm), C.c_void_p(weights_ptr) if weights_ptr else None))

    def timer_start(self):
        self.api.check(self.api.lib.nrxh_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        self.api.check(self.api.lib.nrxh_timer_stop(self.h, C.byref(ms)))
        return ms.value
