"""TEST INFRASTRUCTURE ONLY -- the parity oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``aligntools.c_b200``) never does: it fails loudly when its CUDA library is missing.

Two checkers live here (see oracle/Makefile):

* ``ref``  -- the UNMODIFIED reference DP core (``/root/reference/src/alignment.h``)
  compiled in-process into ``oracle/_ref/libaligntools_ref.so`` (binary only; it is
  built where the reference exists and travels to the GPU box with the snapshot).
* ``port`` -- ``oracle/at_oracle.c``, an integer restatement of SURVEY.md Appendix A
  (1 B/cell instead of the reference's 48 B/cell), pinned against ``ref`` and against
  the golden vectors in ``tests/golden``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libaligntools_ref.so")
REF_CLI = os.path.join(HERE, "_ref", "alignTools_ref")
PORT_SO = os.path.join(HERE, "libat_oracle.so")

MODES = {"global": 0, "local": 1, "fit": 2, "overlap": 3, "edit": 4}


def build(quiet: bool = True) -> None:
    """Compile the port (always) and oracle/_ref (only where /root/reference exists)."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


_port = None
_ref = None

_u8p = C.POINTER(C.c_uint8)


def _load_port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build()
        lib = C.CDLL(PORT_SO)
        lib.at_oracle_align.restype = C.c_int
        lib.at_oracle_align.argtypes = [
            C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_void_p, C.c_size_t, C.POINTER(C.c_int64), C.c_void_p,
            C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_size_t)]
        lib.at_oracle_batch.restype = C.c_int
        lib.at_oracle_batch.argtypes = [C.c_int] * 7 + [C.c_size_t] + [C.c_void_p] * 15 + [C.c_int]
        _port = lib
    return _port


def _load_ref():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            raise RuntimeError("oracle/_ref/libaligntools_ref.so missing (run `make -C oracle` where /root/reference exists)")
        lib = C.CDLL(REF_SO)
        lib.ref_align.restype = C.c_int
        lib.ref_align.argtypes = [
            C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.c_char_p, C.c_char_p,
            C.POINTER(C.c_size_t)]
        lib.ref_align_batch.restype = C.c_int
        lib.ref_align_batch.argtypes = [C.c_int] * 7 + [C.c_size_t] + [C.c_void_p] * 13 + [C.c_int]
        _ref = lib
    return _ref


@dataclass
class Params:
    """Scoring parameters; defaults are init_opt()'s (src/alignment.h:102-114)."""
    m: int = 1
    u: int = -2
    o: int = -5
    e: int = -1
    j: int = -10
    jump: bool = False


@dataclass
class Result:
    score: int
    r1: bytes
    r2: bytes
    ops: bytes | None = None         # per column: M / I (LOW) / D (UPP) / N (JUMP); port only
    coords: tuple | None = None      # (end_i, end_j, beg_i, beg_j); port only


def _mode(mode):
    return MODES[mode] if isinstance(mode, str) else int(mode)


def _sites_arr(sites):
    if sites is None:
        return None, 0
    a = np.ascontiguousarray(np.asarray(sites, dtype=np.int32))
    return a, a.size


def ref_align(mode, s1: bytes, s2: bytes, p: Params = Params(), sites=None) -> Result:
    """One pair through the compiled, unmodified reference."""
    lib = _load_ref()
    md = _mode(mode)
    if md == 2 and len(s1) > len(s2):
        raise ValueError("fit: first sequence must be shorter than the second (reference dies, :599)")
    cap = len(s1) + len(s2) + 1
    r1 = C.create_string_buffer(cap)
    r2 = C.create_string_buffer(cap)
    sc = C.c_double()
    al = C.c_size_t()
    sa, ns = _sites_arr(sites)
    rc = lib.ref_align(md, s1, len(s1), s2, len(s2), p.m, p.u, p.o, p.e, p.j, int(p.jump),
                       sa.ctypes.data if sa is not None and ns else None, ns,
                       C.byref(sc), r1, r2, C.byref(al))
    if rc:
        raise RuntimeError(f"ref_align rc={rc}")
    return Result(int(sc.value), r1.raw[:al.value], r2.raw[:al.value])


def port_align(mode, s1: bytes, s2: bytes, p: Params = Params(), sites=None) -> Result:
    """One pair through the integer restatement (oracle/at_oracle.c)."""
    lib = _load_port()
    md = _mode(mode)
    cap = len(s1) + len(s2) + 1
    r1 = C.create_string_buffer(cap)
    r2 = C.create_string_buffer(cap)
    ops = C.create_string_buffer(cap)
    sc = C.c_int64()
    al = C.c_size_t()
    coords = (C.c_int32 * 4)()
    sa, ns = _sites_arr(sites)
    rc = lib.at_oracle_align(md, s1, len(s1), s2, len(s2), p.m, p.u, p.o, p.e, p.j, int(p.jump),
                             sa.ctypes.data if sa is not None and ns else None, ns,
                             C.byref(sc), C.cast(coords, C.c_void_p), r1, r2, ops, C.byref(al))
    if rc:
        raise RuntimeError(f"at_oracle_align rc={rc}")
    return Result(int(sc.value), r1.raw[:al.value], r2.raw[:al.value], ops.raw[:al.value], tuple(coords))


def _ptr(a):
    return None if a is None else a.ctypes.data


class BatchOut:
    def __init__(self, n, q_len, t_len, want_aln=True, want_ops=False):
        cap = q_len.astype(np.uint64) + t_len.astype(np.uint64) + 1
        self.aln_off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(cap, out=self.aln_off[1:])
        tot = int(self.aln_off[-1])
        self.score = np.zeros(n, dtype=np.int64)
        self.coords = np.zeros((n, 4), dtype=np.int32)
        self.aln_len = np.zeros(n, dtype=np.uint32)
        self.r1 = np.zeros(tot, dtype=np.uint8) if want_aln else None
        self.r2 = np.zeros(tot, dtype=np.uint8) if want_aln else None
        self.ops = np.zeros(tot, dtype=np.uint8) if (want_aln and want_ops) else None

    def aln(self, p):
        o, n = int(self.aln_off[p]), int(self.aln_len[p])
        return self.r1[o:o + n].tobytes(), self.r2[o:o + n].tobytes()

    def op(self, p):
        o, n = int(self.aln_off[p]), int(self.aln_len[p])
        return self.ops[o:o + n].tobytes()


def port_batch(mode, p: Params, q, q_off, q_len, t, t_off, t_len, sites=None, site_off=None,
               want_aln=True, want_ops=False, threads=1) -> BatchOut:
    """Many pairs through the port.  q/t: uint8 arrays; *_off uint64; *_len uint32."""
    lib = _load_port()
    n = len(q_len)
    out = BatchOut(n, q_len, t_len, want_aln, want_ops)
    rc = lib.at_oracle_batch(_mode(mode), p.m, p.u, p.o, p.e, p.j, int(p.jump), n,
                             _ptr(q), _ptr(q_off), _ptr(q_len), _ptr(t), _ptr(t_off), _ptr(t_len),
                             _ptr(sites), _ptr(site_off), _ptr(out.score), _ptr(out.coords),
                             _ptr(out.r1), _ptr(out.r2), _ptr(out.ops), _ptr(out.aln_off),
                             _ptr(out.aln_len), threads)
    if rc:
        raise RuntimeError(f"at_oracle_batch rc={rc}")
    return out


def ref_batch(mode, p: Params, q, q_off, q_len, t, t_off, t_len, sites=None, site_off=None,
              want_aln=True, threads=1) -> BatchOut:
    """Many pairs through the compiled reference (its own alloc/fill/traceback/free per pair)."""
    lib = _load_ref()
    n = len(q_len)
    out = BatchOut(n, q_len, t_len, want_aln, False)
    score = np.zeros(n, dtype=np.float64)
    rc = lib.ref_align_batch(_mode(mode), p.m, p.u, p.o, p.e, p.j, int(p.jump), n,
                             _ptr(q), _ptr(q_off), _ptr(q_len), _ptr(t), _ptr(t_off), _ptr(t_len),
                             _ptr(sites), _ptr(site_off), _ptr(score),
                             _ptr(out.r1), _ptr(out.r2), _ptr(out.aln_off), _ptr(out.aln_len), threads)
    if rc:
        raise RuntimeError(f"ref_align_batch rc={rc}")
    out.score = score.astype(np.int64)
    return out


def run_ref_cli(args, cwd=None):
    """Run the compiled reference CLI; returns (rc, stdout bytes, stderr bytes)."""
    pr = subprocess.run([REF_CLI] + list(args), capture_output=True, cwd=cwd)
    return pr.returncode, pr.stdout, pr.stderr


def ref_kseq_dump(path: str) -> bytes | None:
    """Records of a FASTA/FASTQ(.gz) file as the reference's kseq parser sees them (one line per
    record: name, comment.s or "(null)", strlen(seq), seq); None when gzopen fails."""
    lib = _load_ref()
    lib.ref_kseq_dump.restype = C.c_long
    lib.ref_kseq_dump.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    cap = 8 * max(os.path.getsize(path), 1 << 12) + (1 << 16) if os.path.exists(path) else 1 << 12
    buf = C.create_string_buffer(cap)
    n = lib.ref_kseq_dump(path.encode(), buf, cap)
    return None if n < 0 else buf.raw[:n]
