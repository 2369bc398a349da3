"""TEST INFRASTRUCTURE ONLY -- never imported by the product package ``bioseq_b200``.

ctypes front-end of the plain-C restatement ``oracle/bsq_oracle.c`` plus a loader for the
reference's own compiled tokenizer (``oracle/_ref/ref_cbioseq*.so``, built by
``oracle/Makefile`` from the sources under /root/reference; present in the build
container and shipped to the GPU box as a prebuilt file, never rebuilt there).

Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl
reference legs).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KINDS = {"b": 0, "h": 1, "i": 2, "l": 3, "q": 3, "f": 4, "d": 5}
NP_OF_KIND = {0: np.int8, 1: np.int16, 2: np.int32, 3: np.uint64, 4: np.float32, 5: np.float64}

ERR_TOO_LONG = -2
ERR_BAD_TOKEN = -3


def build():
    """Compile oracle/bsq_oracle.c -> oracle/libbsq_oracle.so (gcc, ~1 s)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libbsq_oracle.so")
        src = os.path.join(HERE, "bsq_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        L = C.CDLL(path)
        L.bsqo_alphabet.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(C.c_int)]
        L.bsqo_alphabet_key.restype = C.c_char_p
        L.bsqo_ids.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int)] * 4
        L.bsqo_tokenize.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.POINTER(C.c_int64)]
        L.bsqo_onehot.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                  C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_void_p, C.POINTER(C.c_int64)]
        L.bsqo_decode.restype = C.c_int64
        L.bsqo_decode.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                  C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_uint32)]
        L.bsqo_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bsqo_philox4x32_10.restype = None
        L.bsqo_blosum62_probs.argtypes = [C.c_void_p]
        L.bsqo_blosum62_probs.restype = None
        L.bsqo_blosum62_thresholds.argtypes = [C.c_void_p]
        L.bsqo_blosum62_thresholds.restype = None
        L.bsqo_augment.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_uint64, C.c_int64]
        _LIB = L
    return _LIB


def alphabet_keys():
    L = lib()
    return [L.bsqo_alphabet_key(i).decode() for i in range(L.bsqo_alphabet_count())]


def pack(seqs):
    """list of str/bytes/bytearray -> (uint8 buffer, int64 offsets) like the pack layer."""
    bs = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in seqs]
    offs = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        np.cumsum([len(b) for b in bs], out=offs[1:])
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return buf, offs


class OracleTokenizer:
    """CPU restatement of cbioseq.Tokenizer (src/tokenize.cpp:22-112) over packed input."""

    def __init__(self, key, eos=False, bos=False, padchar=False):
        L = lib()
        self.key = key.upper()
        self.lut = np.zeros(256, dtype=np.int8)
        n = C.c_int()
        if L.bsqo_alphabet(key.encode(), self.lut.ctypes.data, C.byref(n)) != 0:
            raise RuntimeError("Invalid tokenizer type")
        self._nchars = n.value
        b, e, p, a = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.bsqo_ids(n.value, int(eos), int(bos), int(padchar), C.byref(b), C.byref(e), C.byref(p), C.byref(a))
        self._bos, self._eos, self._pad, self._size = b.value, e.value, p.value, a.value
        self._padchar = bool(padchar)

    def nchars(self): return self._nchars
    def bos(self): return self._bos
    def eos(self): return self._eos
    def pad(self): return self._pad
    def alphabet_size(self): return self._size
    def is_padded(self): return self._padchar

    @staticmethod
    def _packed(batch):
        if isinstance(batch, tuple) and len(batch) == 2 and isinstance(batch[0], np.ndarray):
            buf, offs = batch
            return np.ascontiguousarray(buf, dtype=np.uint8), np.ascontiguousarray(offs, dtype=np.int64)
        return pack(batch)

    def batch_tokenize(self, batch, padlen=-1, destchar="B", batch_first=False):
        if padlen <= 0:
            raise ValueError("batch tokenize requires padlen is provded.")
        kind = KINDS[destchar[0].lower()]
        buf, offs = self._packed(batch)
        n = len(offs) - 1
        out = np.empty((n, padlen) if batch_first else (padlen, n), dtype=NP_OF_KIND[kind])
        bad = C.c_int64()
        rc = lib().bsqo_tokenize(buf.ctypes.data, offs.ctypes.data, n, padlen, self.lut.ctypes.data,
                                 self._bos, self._eos, self._pad, int(self._padchar),
                                 int(batch_first), kind, out.ctypes.data, C.byref(bad))
        if rc == ERR_TOO_LONG:
            raise RuntimeError(f"seq len + bos + eos > padlen: {bad.value}, vs padlen {padlen}")
        assert rc == 0, rc
        return out

    def batch_onehot_encode(self, batch, padlen=-1, destchar="B", mask=None):
        if padlen <= 0:
            raise ValueError("batch tokenize requires padlen is provded.")
        kind = KINDS[destchar[0].lower()]
        buf, offs = self._packed(batch)
        n = len(offs) - 1
        m = None
        if mask is not None:
            m = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.uint8) for x in mask])
                                     if len(mask) else np.zeros(0, np.uint8))
            assert m.size == buf.size
        out = np.empty((padlen, n, self._size), dtype=NP_OF_KIND[kind])
        bad = C.c_int64()
        rc = lib().bsqo_onehot(buf.ctypes.data, offs.ctypes.data, m.ctypes.data if m is not None else None,
                               n, padlen, self.lut.ctypes.data, self._bos, self._eos, self._pad,
                               int(self._padchar), self._size, kind, out.ctypes.data, C.byref(bad))
        if rc == ERR_TOO_LONG:
            raise ValueError(f"seq len + bos + eos > padlen: {bad.value}, vs padlen {padlen}")
        assert rc == 0, rc
        return out

    def decode_tokens(self, arr):
        arr = np.asarray(arr)
        if arr.ndim not in (1, 2):
            raise ValueError("Currently supported: 1 or 2 dimensions for decoding tokens.")
        a2 = arr.reshape(1, -1) if arr.ndim == 1 else arr
        rows, cols = a2.shape
        if arr.ndim == 1:
            rs, cs = 0, arr.strides[0]
        else:
            rs, cs = arr.strides
        cap = rows * cols * 5 + 1
        chars = np.empty(cap, dtype=np.uint8)
        offs = np.empty(rows + 1, dtype=np.int64)
        bad = C.c_uint32()
        rc = lib().bsqo_decode(arr.ctypes.data, arr.itemsize, rows, cols, rs, cs, self.lut.ctypes.data,
                               self._bos, self._eos, self._pad, int(self._padchar),
                               chars.ctypes.data, cap, offs.ctypes.data, C.byref(bad))
        if rc == ERR_BAD_TOKEN:
            raise RuntimeError(f"Unexpected/invalid token {bad.value}")
        assert rc >= 0, rc
        raw = chars.tobytes()
        out = [raw[offs[i]:offs[i + 1]].decode("latin-1") for i in range(rows)]
        return out[0] if arr.ndim == 1 else out


# ---------------------------------------------------------------------------------------------
# BLOSUM62 augmentation (bioseq/blosum.py:36-87) -- see the block comment in bsq_oracle.c.
# ---------------------------------------------------------------------------------------------
BLOSUM_ORDER = "ARNDCQEGHILKMFPSTWYV"


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32).copy()
    k = np.asarray(key, dtype=np.uint32).copy()
    out = np.zeros(4, dtype=np.uint32)
    lib().bsqo_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def blosum62_probs():
    """normrows of bioseq/blosum.py:43 (21 x 20: ARNDCQEGHILKMFPSTWYV + X rows)."""
    p = np.zeros((21, 20), dtype=np.float64)
    lib().bsqo_blosum62_probs(p.ctypes.data)
    return p


def blosum62_thresholds():
    t = np.zeros((21, 19), dtype=np.uint32)
    lib().bsqo_blosum62_thresholds(t.ctypes.data)
    return t


def augment(buf, offs, chain_len=1, augment_frac=1.0, seed=0, seq_index_base=0):
    """augment_seq over a packed batch (returns a mutated copy of ``buf``)."""
    out = np.ascontiguousarray(buf, dtype=np.uint8).copy()
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    rc = lib().bsqo_augment(out.ctypes.data, offs.ctypes.data, len(offs) - 1, chain_len, float(augment_frac),
                            seed & 0xFFFFFFFFFFFFFFFF, seq_index_base)
    assert rc == 0, rc
    return out


def load_ref(opt="O3"):
    """Import the reference's own compiled tokenizer if the prebuilt file is present.

    Returns the module (``ref_cbioseq`` / ``ref_cbioseq_O0``) or None.  Only one of the two
    can live in a process (both register the C++ type ``Tokenizer`` with pybind11)."""
    d = os.path.join(HERE, "_ref")
    name = "ref_cbioseq" if opt == "O3" else "ref_cbioseq_O0"
    if not os.path.isdir(d):
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        return __import__(name)
    except ImportError:
        return None


# ---------------------------------------------------------------------------------------------
# FlatFile (src/fxstats.cpp:26-134): restatement of the FASTA/FASTQ -> flat file conversion.
# Pinned against the compiled reference (oracle/_ref) by tests/test_flatfile.py and against
# tests/golden/flatfile.json (generated from the reference by oracle/make_golden_flatfile.py).
# ---------------------------------------------------------------------------------------------
_SPACE = b" \t\n\v\f\r"


def parse_fastx(data):
    """Sequences of a FASTA/FASTQ byte string under the rules of the reference's vendored
    kseq.h (src/kseq.h:173-216) as driven by FlatFile::make (src/fxstats.cpp:40-49):
    kseq_read until the first negative return."""
    n, p = len(data), 0
    seqs = []
    last_char = 0

    def line_end(q):
        e = data.find(b"\n", q)
        return n if e < 0 else e

    while True:
        if last_char == 0:                                  # kseq.h:178-182 jump to the next header
            while p < n and data[p] not in b">@":
                p += 1
            if p >= n:
                break
            last_char = data[p]
            p += 1
        if p >= n:                                          # :184 name read hits EOF
            break
        i = p
        while i < n and data[i] not in _SPACE:
            i += 1
        delim = data[i] if i < n else 0
        p = i + 1 if i < n else n
        if delim != 0x0A:                                   # :185 comment = rest of the line
            e = line_end(p)
            p = e + 1 if e < n else n
        seq = bytearray()
        c = -1
        while True:                                         # :190-194 sequence lines
            if p >= n:
                c = -1
                break
            c = data[p]
            p += 1
            if c in (0x3E, 0x2B, 0x40):
                break
            if c == 0x0A:
                continue
            seq.append(c)
            if p < n:
                e = line_end(p)
                seq += data[p:e]
                p = e + 1 if e < n else n
                if len(seq) > 1 and seq[-1] == 0x0D:        # :138 one trailing CR per line
                    seq.pop()
        if c in (0x3E, 0x40):
            last_char = c
        if c == 0x2B:                                       # '+': FASTQ quality block :203-214
            e = data.find(b"\n", p)
            if e < 0:
                break                                       # -2: no quality string
            p = e + 1
            qual = bytearray()
            while p < n:
                e = line_end(p)
                qual += data[p:e]
                p = e + 1 if e < n else n
                if len(qual) > 1 and qual[-1] == 0x0D:
                    qual.pop()
                if len(qual) >= len(seq):
                    break
            last_char = 0
            if len(qual) != len(seq):
                break                                       # -2: quality of a different length
        seqs.append(bytes(seq))
    return seqs


def flatfile_image(seqs):
    """Bytes of the flat file for a list of sequences (layout src/fxstats.cpp:50-59)."""
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        np.cumsum([len(s) for s in seqs], out=offs[1:])
    return np.uint64(len(seqs)).tobytes() + offs.tobytes() + b"".join(seqs)
