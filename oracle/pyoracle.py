"""ctypes access to the CPU checkers.  TEST INFRASTRUCTURE ONLY (see ksw_oracle.c).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never by genomicsbench_b200/.

  Oracle      oracle/libbsw_oracle.so   from-scratch restatement (always available; built by
                                        oracle/Makefile, travels to the GPU box prebuilt)
  Reference   oracle/_ref/libbswref.so  the unmodified reference class (bandedSWA.cpp) behind
                                        oracle/ref_shim.cpp; present when it was built from
                                        /root/reference in the build container
"""
from __future__ import annotations

import ctypes as C
from typing import Optional
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "libbsw_oracle.so"
REF_SO = HERE / "_ref" / "libbswref.so"
KSW_SO = HERE / "_ref" / "libkswref.so"


CHAIN_SEED_DTYPE = np.dtype([("rbeg", "<i8"), ("qbeg", "<i4"), ("len", "<i4"), ("score", "<i4"), ("pad", "<i4")])   # mem_seed_t
CHAIN_REG_DTYPE = np.dtype([("rb", "<i8"), ("re", "<i8"), ("qb", "<i4"), ("qe", "<i4"), ("score", "<i4"), ("truesc", "<i4"),
                            ("w", "<i4"), ("seedcov", "<i4"), ("seedlen0", "<i4"), ("pad", "<i4")])
CHAIN_REG_FIELDS = ("rb", "re", "qb", "qe", "score", "truesc", "w", "seedcov", "seedlen0")


class OracleParams(C.Structure):
    _fields_ = [(k, C.c_int32) for k in
                ("o_del", "e_del", "o_ins", "e_ins", "zdrop", "end_bonus", "match", "mismatch", "ambig",
                 "zdrop_mode")]


def make_params(o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100, end_bonus=5, match=1, mismatch=4,
                ambig=-1, zdrop_mode=0) -> OracleParams:
    return OracleParams(o_del, e_del, o_ins, e_ins, zdrop, end_bonus, match, mismatch, ambig, zdrop_mode)


def build(quiet: bool = True) -> None:
    """(Re)builds the checker libraries; the reference part only when /root/reference exists."""
    subprocess.run(["make", "-C", str(HERE)], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class Oracle:
    def __init__(self):
        if not ORACLE_SO.exists():
            build()
        self.lib = C.CDLL(str(ORACLE_SO))
        self.lib.bsw_oracle_batch.restype = C.c_int64
        self.lib.bsw_oracle_batch.argtypes = [C.POINTER(OracleParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_int64, C.c_int, C.c_int]
        self.lib.bsw_oracle_pair.restype = C.c_int64
        self.lib.bsw_oracle_pair.argtypes = [C.POINTER(OracleParams), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                             C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.bsw_oracle_row_trips.restype = C.c_int64
        self.lib.bsw_oracle_row_trips.argtypes = [C.POINTER(OracleParams), C.c_void_p, C.c_int, C.c_void_p,
                                                  C.c_int, C.c_int, C.c_int, C.c_void_p]
        self.lib.bsw_oracle_max_threads.restype = C.c_int
        self.lib.bsw_oracle_chain_window.restype = None
        self.lib.bsw_oracle_chain_window.argtypes = [C.POINTER(OracleParams), C.c_int, C.c_int64, C.c_void_p, C.c_int,
                                                     C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        self.lib.bsw_oracle_chain.restype = C.c_int
        self.lib.bsw_oracle_chain.argtypes = [C.POINTER(OracleParams), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p,
                                              C.c_void_p, C.c_int, C.c_void_p]

    def max_threads(self) -> int:
        return int(self.lib.bsw_oracle_max_threads())

    def batch(self, params: OracleParams, pairs: np.ndarray, seq_ref: np.ndarray, seq_qer: np.ndarray,
              w: int, nthreads: int = 0) -> int:
        """Writes the six result fields into pairs (in place); returns effective cells."""
        if nthreads <= 0:
            nthreads = self.max_threads()
        return int(self.lib.bsw_oracle_batch(C.byref(params), pairs.ctypes.data, seq_ref.ctypes.data,
                                             seq_qer.ctypes.data, len(pairs), w, nthreads))

    # ---- banded global alignment with traceback (global_oracle.c) ----------------------------------
    def global_align(self, params: OracleParams, query: np.ndarray, target: np.ndarray, w: int):
        """ksw_global2 (tools/bwa/ksw.c:502-606) -> (score, cigar uint32[n]: len << 4 | op, op 0 M / 1 I / 2 D)."""
        query = np.ascontiguousarray(query); target = np.ascontiguousarray(target)
        cig = np.zeros(len(query) + len(target) + 1, dtype=np.uint32)
        n = C.c_int(0)
        fn = self.lib.bsw_oracle_global
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(OracleParams), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        score = fn(C.byref(params), query.ctypes.data, len(query), target.ctypes.data, len(target), w, cig.ctypes.data, C.byref(n))
        return int(score), cig[:n.value].copy()

    # ---- seed -> pair construction / chain extension (chain_oracle.c) ----------------------------
    def chain_window(self, params: OracleParams, w: int, l_pac: int, seeds: np.ndarray, l_query: int):
        """Reference window [rmax0, rmax1) of a chain (tools/bwa/bwamem.c:643-659); seeds = CHAIN_SEED_DTYPE."""
        r0, r1 = C.c_int64(0), C.c_int64(0)
        seeds = np.ascontiguousarray(seeds)
        self.lib.bsw_oracle_chain_window(C.byref(params), w, C.c_int64(l_pac), seeds.ctypes.data, len(seeds), l_query,
                                         C.byref(r0), C.byref(r1))
        return int(r0.value), int(r1.value)

    def chain(self, params: OracleParams, w: int, pen_clip5: int, pen_clip3: int, max_band_try: int,
              query: np.ndarray, seeds: np.ndarray, rmax0: int, rmax1: int, rseq: np.ndarray,
              prior: Optional[np.ndarray] = None) -> np.ndarray:
        """mem_chain2aln of one chain (tools/bwa/bwamem.c:632-822) -> CHAIN_REG_DTYPE array, reference order.
        prior = the regions of the read's earlier chains (one shared mem_alnreg_v, bwamem.c:1105-1112)."""
        prior = np.zeros(0, dtype=CHAIN_REG_DTYPE) if prior is None else np.ascontiguousarray(prior, dtype=CHAIN_REG_DTYPE)
        seeds = np.ascontiguousarray(seeds)
        query = np.ascontiguousarray(query); rseq = np.ascontiguousarray(rseq)
        out = np.zeros(max(len(seeds), 1), dtype=CHAIN_REG_DTYPE)
        n = self.lib.bsw_oracle_chain(C.byref(params), w, pen_clip5, pen_clip3, max_band_try, len(query),
                                      query.ctypes.data, seeds.ctypes.data, len(seeds), C.c_int64(rmax0),
                                      C.c_int64(rmax1), rseq.ctypes.data, prior.ctypes.data, len(prior), out.ctypes.data)
        return out[:n]

    def band_retry(self, params: OracleParams, pairs: np.ndarray, seq_ref: np.ndarray, seq_qer: np.ndarray,
                   w: int, max_try: int = 2, prev_score=None) -> np.ndarray:
        """Restates the MAX_BAND_TRY loops of mem_chain2aln (tools/bwa/bwamem.c:630,723-753,770-800):
        for t in range(max_try): prev = score; w_t = w << t; extend; break if score == prev or
        max_off < (w_t >> 1) + (w_t >> 2).  Results in place; returns the band of each pair's last try."""
        n = len(pairs)
        prev = np.full(n, -1, dtype=np.int64) if prev_score is None else np.asarray(prev_score, dtype=np.int64).copy()
        band = np.zeros(n, dtype=np.int32)
        active = np.arange(n)
        for t in range(max_try):
            if len(active) == 0:
                break
            wt = w << t
            sub = pairs[active].copy()
            self.batch(params, sub, seq_ref, seq_qer, wt)
            for f in ("score", "qle", "tle", "gtle", "gscore", "max_off"):
                pairs[f][active] = sub[f]
            band[active] = wt
            stop = (sub["score"] == prev[active]) | (sub["max_off"] < (wt >> 1) + (wt >> 2))
            prev[active] = sub["score"]
            active = active[~stop]
        return band

    def pair(self, params: OracleParams, query: np.ndarray, target: np.ndarray, w: int, h0: int):
        out = np.zeros(6, dtype=np.int32)
        cells = self.lib.bsw_oracle_pair(C.byref(params), query.ctypes.data, len(query), target.ctypes.data,
                                         len(target), w, h0, out.ctypes.data, None)
        return dict(score=int(out[0]), qle=int(out[1]), tle=int(out[2]), gtle=int(out[3]),
                    gscore=int(out[4]), max_off=int(out[5]), cells=int(cells))

    def row_trips(self, params, query, target, w, h0) -> np.ndarray:
        trips = np.zeros(len(target), dtype=np.int32)
        self.lib.bsw_oracle_row_trips(C.byref(params), query.ctypes.data, len(query), target.ctypes.data,
                                      len(target), w, h0, trips.ctypes.data)
        return trips


class Reference:
    """The reference's own code (AVX2 getScores16 / scalarBandedSWA)."""

    def __init__(self):
        if not REF_SO.exists():
            raise FileNotFoundError(f"{REF_SO} absent (built only where /root/reference is mounted)")
        self.lib = C.CDLL(str(REF_SO))
        P = C.POINTER(OracleParams)
        self.lib.ref_getscores16.restype = C.c_double
        self.lib.ref_getscores16.argtypes = [P, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                             C.c_int32, C.c_int32]
        self.lib.ref_time_getscores16_inplace.restype = C.c_double
        self.lib.ref_time_getscores16_inplace.argtypes = self.lib.ref_getscores16.argtypes
        self.lib.ref_getscores16_solo.restype = None
        self.lib.ref_getscores16_solo.argtypes = [P, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        self.lib.ref_scalar.restype = C.c_double
        self.lib.ref_scalar.argtypes = [P, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
        self.lib.ref_max_threads.restype = C.c_int
        self.lib.ref_sizeof_seqpair.restype = C.c_int

    @staticmethod
    def available() -> bool:
        return REF_SO.exists()

    def max_threads(self) -> int:
        return int(self.lib.ref_max_threads())

    def getscores16(self, params, pairs, seq_ref, seq_qer, w, batch=512, nthreads=0) -> float:
        if nthreads <= 0:
            nthreads = self.max_threads()
        return float(self.lib.ref_getscores16(C.byref(params), pairs.ctypes.data, seq_ref.ctypes.data,
                                              seq_qer.ctypes.data, len(pairs), w, batch, nthreads))

    def time_getscores16_inplace(self, params, pairs_padded, n, seq_ref, seq_qer, w, batch=512, nthreads=0):
        """pairs_padded must have capacity roundup16(n)+2 (reference writes pads, reads 2 beyond)."""
        if nthreads <= 0:
            nthreads = self.max_threads()
        return float(self.lib.ref_time_getscores16_inplace(C.byref(params), pairs_padded.ctypes.data,
                                                           seq_ref.ctypes.data, seq_qer.ctypes.data, n, w,
                                                           batch, nthreads))

    def solo(self, params, pair_rec: np.ndarray, seq_ref, seq_qer, w) -> None:
        self.lib.ref_getscores16_solo(C.byref(params), pair_rec.ctypes.data, seq_ref.ctypes.data,
                                      seq_qer.ctypes.data, w)

    def scalar(self, params, pairs, seq_ref, seq_qer, w) -> float:
        return float(self.lib.ref_scalar(C.byref(params), pairs.ctypes.data, seq_ref.ctypes.data,
                                         seq_qer.ctypes.data, len(pairs), w))


class KswReference:
    """The canonical ksw_extend2 of the reference tree (tools/bwa/ksw.c:380-479), compiled unmodified
    into oracle/_ref/libkswref.so; used to pin the band-retry restatement with the loop exactly as
    mem_chain2aln writes it (tools/bwa/bwamem.c:723-753)."""

    def __init__(self):
        if not KSW_SO.exists():
            raise FileNotFoundError(f"{KSW_SO} absent (built only where /root/reference is mounted)")
        self.lib = C.CDLL(str(KSW_SO))
        I, PI = C.c_int, C.POINTER(C.c_int)
        self.lib.ksw_extend2.restype = I
        self.lib.ksw_extend2.argtypes = [I, C.c_void_p, I, C.c_void_p, I, C.c_void_p, I, I, I, I, I, I, I, I,
                                         PI, PI, PI, PI, PI]

    @staticmethod
    def available() -> bool:
        return KSW_SO.exists()

    def global_align(self, params, query: np.ndarray, target: np.ndarray, w: int):
        """The reference's ksw_global2 (tools/bwa/ksw.c:502-606) -> (score, cigar uint32[n])."""
        mat = self.scmat(params.match, params.mismatch, params.ambig)
        query = np.ascontiguousarray(query); target = np.ascontiguousarray(target)
        fn = self.lib.ksw_global2
        fn.restype = C.c_int
        fn.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                       C.c_int, C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_uint32))]
        n = C.c_int(0); cg = C.POINTER(C.c_uint32)()
        score = fn(len(query), query.ctypes.data, len(target), target.ctypes.data, 5, mat.ctypes.data, params.o_del,
                   params.e_del, params.o_ins, params.e_ins, w, C.byref(n), C.byref(cg))
        out = np.array([cg[i] for i in range(n.value)], dtype=np.uint32)
        if cg:
            C.CDLL(None).free(C.cast(cg, C.c_void_p))
        return int(score), out

    @staticmethod
    def scmat(match: int, mismatch: int, ambig: int) -> np.ndarray:
        """bwa_fill_scmat (benchmarks/bsw/main_banded.cpp:73-81)."""
        m = np.zeros(25, dtype=np.int8)
        k = 0
        for i in range(4):
            for j in range(4):
                m[k] = match if i == j else -mismatch
                k += 1
            m[k] = ambig
            k += 1
        m[20:25] = ambig
        return m

    def band_retry(self, params: OracleParams, pairs: np.ndarray, seq_ref: np.ndarray, seq_qer: np.ndarray,
                   w: int, max_try: int = 2, prev_score=None) -> np.ndarray:
        mat = self.scmat(params.match, params.mismatch, params.ambig)
        band = np.zeros(len(pairs), dtype=np.int32)
        out = [C.c_int(0) for _ in range(5)]
        for i in range(len(pairs)):
            q = np.ascontiguousarray(seq_qer[pairs["idq"][i]: pairs["idq"][i] + pairs["len2"][i]])
            t = np.ascontiguousarray(seq_ref[pairs["idr"][i]: pairs["idr"][i] + pairs["len1"][i]])
            score = -1 if prev_score is None else int(prev_score[i])
            for tr in range(max_try):                                   # bwamem.c:723 / :770
                prev = score                                            # :724 / :771
                aw = w << tr                                            # :725 / :772
                score = self.lib.ksw_extend2(len(q), q.ctypes.data, len(t), t.ctypes.data, 5, mat.ctypes.data,
                                             params.o_del, params.e_del, params.o_ins, params.e_ins, aw,
                                             params.end_bonus, params.zdrop, int(pairs["h0"][i]),
                                             *[C.byref(x) for x in out])  # :746 / :793
                band[i] = aw
                if score == prev or out[4].value < (aw >> 1) + (aw >> 2):   # :753 / :800
                    break
            pairs["score"][i] = score
            for f, x in zip(("qle", "tle", "gtle", "gscore", "max_off"), out):
                pairs[f][i] = x.value
        return band
