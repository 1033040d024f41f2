"""ctypes mirror of the reference's handle types (``/root/reference/include/mosfhet.h:22-133``).

The reference's API is a flat C namespace over pointer-to-struct handles.  This module gives the
host side the *same* handle layouts so that Python callers (and the parity tests) can pass
ciphertexts and keys to ``libmosfhet_b200.so`` exactly as a C caller of the reference would:

* ``TLWE``  = ``struct _TLWE {Torus *a, b; int n;} *``                 (mosfhet.h:51-54)
* ``TRLWE`` = ``struct _TRLWE {TorusPolynomial *a, b; int k;} *``      (mosfhet.h:73-76)
* ``TRGSW_DFT``, ``Bootstrap_Key``, ``TLWE_KS_Key``                    (mosfhet.h:111-114, 129-133, 62-65)

``Host*`` helpers build such pointer trees on top of numpy buffers (which they keep alive) and
flatten trees produced by a C library back into numpy arrays in the flat layouts documented in
``include/mosfhet_b200.h`` section 3.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

Torus = C.c_uint64


class TorusPolynomialS(C.Structure):          # mosfhet.h:32-35
    _fields_ = [("coeffs", C.POINTER(Torus)), ("N", C.c_int)]


class DFTPolynomialS(C.Structure):            # mosfhet.h:37-40
    _fields_ = [("coeffs", C.POINTER(C.c_double)), ("N", C.c_int)]


class TLWES(C.Structure):                     # mosfhet.h:51-54
    _fields_ = [("a", C.POINTER(Torus)), ("b", Torus), ("n", C.c_int)]


class TLWEKeyS(C.Structure):                  # mosfhet.h:56-60
    _fields_ = [("s", C.POINTER(Torus)), ("n", C.c_int), ("sigma", C.c_double)]


TLWE = C.POINTER(TLWES)


class TLWEKSKeyS(C.Structure):                # mosfhet.h:62-65
    _fields_ = [("s", C.POINTER(C.POINTER(C.POINTER(TLWE)))),
                ("base_bit", C.c_int), ("t", C.c_int), ("n", C.c_int)]


TorusPolynomial = C.POINTER(TorusPolynomialS)
DFTPolynomial = C.POINTER(DFTPolynomialS)


class TRLWES(C.Structure):                    # mosfhet.h:73-76
    _fields_ = [("a", C.POINTER(TorusPolynomial)), ("b", TorusPolynomial), ("k", C.c_int)]


class TRLWEDFTS(C.Structure):                 # mosfhet.h:78-81
    _fields_ = [("a", C.POINTER(DFTPolynomial)), ("b", DFTPolynomial), ("k", C.c_int)]


TRLWE = C.POINTER(TRLWES)
TRLWE_DFT = C.POINTER(TRLWEDFTS)


class TRLWEKeyS(C.Structure):                 # mosfhet.h:83-88
    _fields_ = [("s", C.POINTER(TorusPolynomial)), ("s_dft", C.POINTER(DFTPolynomial)),
                ("k", C.c_int), ("sigma", C.c_double)]


class TRGSWS(C.Structure):                    # mosfhet.h:106-109
    _fields_ = [("samples", C.POINTER(TRLWE)), ("l", C.c_int), ("Bg_bit", C.c_int)]


class TRGSWDFTS(C.Structure):                 # mosfhet.h:111-114
    _fields_ = [("samples", C.POINTER(TRLWE_DFT)), ("l", C.c_int), ("Bg_bit", C.c_int)]


TRGSW = C.POINTER(TRGSWS)
TRGSW_DFT = C.POINTER(TRGSWDFTS)


class TRGSWKeyS(C.Structure):                 # mosfhet.h:116-119
    _fields_ = [("trlwe_key", C.POINTER(TRLWEKeyS)), ("l", C.c_int), ("Bg_bit", C.c_int)]


class GenericKSKeyS(C.Structure):             # mosfhet.h:100-103
    _fields_ = [("s", C.POINTER(C.POINTER(C.POINTER(TRLWE)))),
                ("base_bit", C.c_int), ("t", C.c_int), ("n", C.c_int), ("include_b", C.c_int)]


class TRLWEKSKeyS(C.Structure):               # mosfhet.h:90-93
    _fields_ = [("s", C.POINTER(C.POINTER(TRLWE_DFT))), ("base_bit", C.c_int), ("t", C.c_int), ("k", C.c_int)]


class BootstrapKeyS(C.Structure):             # mosfhet.h:129-133
    _fields_ = [("s", C.POINTER(TRGSW_DFT)), ("su", C.POINTER(TRGSW)),
                ("n", C.c_int), ("k", C.c_int), ("N", C.c_int), ("Bg_bit", C.c_int),
                ("l", C.c_int), ("unfolding", C.c_int)]


TLWE_Key = C.POINTER(TLWEKeyS)
TLWE_KS_Key = C.POINTER(TLWEKSKeyS)
TRLWE_Key = C.POINTER(TRLWEKeyS)
TRGSW_Key = C.POINTER(TRGSWKeyS)
Bootstrap_Key = C.POINTER(BootstrapKeyS)
Generic_KS_Key = C.POINTER(GenericKSKeyS)
TRLWE_KS_Key = C.POINTER(TRLWEKSKeyS)


class ParamsS(C.Structure):                   # include/mosfhet_b200.h: mb200_params
    _fields_ = [("n", C.c_int), ("N", C.c_int), ("k", C.c_int), ("l", C.c_int),
                ("Bg_bit", C.c_int), ("t", C.c_int), ("base_bit", C.c_int)]


def aligned_empty(shape, dtype, align: int = 64) -> np.ndarray:
    """numpy array whose data pointer is `align`-byte aligned: the reference's AVX-512 build casts
    polynomial storage to ``__m512i *`` / ``__m512d *`` (trlwe.c:394-411, polynomial.c:381-394) and
    its allocator hands out 64-byte aligned blocks (misc.c:115-128)."""
    dtype = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    raw = np.empty(nbytes + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off: off + nbytes].view(dtype).reshape(shape)


def aligned_copy(arr, dtype) -> np.ndarray:
    arr = np.asarray(arr, dtype=dtype)
    out = aligned_empty(arr.shape, dtype)
    out[...] = arr
    return out


def _u64ptr(arr: np.ndarray):
    return arr.ctypes.data_as(C.POINTER(Torus))


def _f64ptr(arr: np.ndarray):
    return arr.ctypes.data_as(C.POINTER(C.c_double))


# ---------------------------------------------------------------------------------------------
# Host-side builders: numpy buffer  ->  reference-layout pointer tree
# ---------------------------------------------------------------------------------------------
class HostTLWE:
    """A ``TLWE`` handle over a flat ``[n+1]`` uint64 array (a[0..n), b)."""

    def __init__(self, flat: np.ndarray):
        flat = np.ascontiguousarray(flat, dtype=np.uint64)
        self.n = flat.shape[0] - 1
        self._a = aligned_copy(flat[: self.n], np.uint64) if self.n else np.zeros(1, np.uint64)
        self.struct = TLWES(_u64ptr(self._a), int(flat[self.n]), self.n)
        self.handle = C.pointer(self.struct)

    @classmethod
    def zeros(cls, n: int) -> "HostTLWE":
        return cls(np.zeros(n + 1, np.uint64))

    def flat(self) -> np.ndarray:
        return np.concatenate([self._a[: self.n], np.array([self.struct.b], np.uint64)])


class HostTRLWE:
    """A ``TRLWE`` handle over a ``[(k+1), N]`` uint64 array (a[0..k) then b)."""

    def __init__(self, polys: np.ndarray):
        self.polys = aligned_copy(polys, np.uint64)
        self.k = self.polys.shape[0] - 1
        self.N = self.polys.shape[1]
        self._ps = [TorusPolynomialS(_u64ptr(self.polys[i]), self.N) for i in range(self.k + 1)]
        self._a = (TorusPolynomial * max(self.k, 1))(*[C.pointer(p) for p in self._ps[: self.k]])
        self.struct = TRLWES(C.cast(self._a, C.POINTER(TorusPolynomial)), C.pointer(self._ps[self.k]), self.k)
        self.handle = C.pointer(self.struct)

    @classmethod
    def zeros(cls, k: int, N: int) -> "HostTRLWE":
        return cls(np.zeros((k + 1, N), np.uint64))

    def flat(self) -> np.ndarray:
        return self.polys


class HostTRLWEDFT:
    """A ``TRLWE_DFT`` handle over a ``[(k+1), N]`` float64 array (Re|Im halves per polynomial)."""

    def __init__(self, polys: np.ndarray):
        self.polys = aligned_copy(polys, np.float64)
        self.k = self.polys.shape[0] - 1
        self.N = self.polys.shape[1]
        self._ps = [DFTPolynomialS(_f64ptr(self.polys[i]), self.N) for i in range(self.k + 1)]
        self._a = (DFTPolynomial * max(self.k, 1))(*[C.pointer(p) for p in self._ps[: self.k]])
        self.struct = TRLWEDFTS(C.cast(self._a, C.POINTER(DFTPolynomial)), C.pointer(self._ps[self.k]), self.k)
        self.handle = C.pointer(self.struct)

    @classmethod
    def zeros(cls, k: int, N: int) -> "HostTRLWEDFT":
        return cls(np.zeros((k + 1, N), np.float64))


class HostTRGSWDFT:
    """A ``TRGSW_DFT`` handle over a ``[(k+1)*l, (k+1), N]`` float64 array (host slot order)."""

    def __init__(self, rows: np.ndarray, l: int, Bg_bit: int):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        self.rows = [HostTRLWEDFT(rows[r]) for r in range(rows.shape[0])]
        self._samples = (TRLWE_DFT * len(self.rows))(*[r.handle for r in self.rows])
        self.struct = TRGSWDFTS(C.cast(self._samples, C.POINTER(TRLWE_DFT)), l, Bg_bit)
        self.handle = C.pointer(self.struct)


class HostBootstrapKey:
    """A ``Bootstrap_Key`` (unfolding == 1) over a ``[n, (k+1)*l, (k+1), N]`` float64 array."""

    def __init__(self, bsk: np.ndarray, k: int, l: int, Bg_bit: int):
        bsk = np.ascontiguousarray(bsk, dtype=np.float64)
        self.n, self.N = bsk.shape[0], bsk.shape[3]
        self.trgsw = [HostTRGSWDFT(bsk[i], l, Bg_bit) for i in range(self.n)]
        self._s = (TRGSW_DFT * self.n)(*[g.handle for g in self.trgsw])
        self.struct = BootstrapKeyS(C.cast(self._s, C.POINTER(TRGSW_DFT)), None,
                                    self.n, k, self.N, Bg_bit, l, 1)
        self.handle = C.pointer(self.struct)


class HostUnfoldedBootstrapKey:
    """A ``Bootstrap_Key`` with ``unfolding`` > 1 over a torus ``[n/u * 2^u, (k+1)*l, (k+1), N]`` uint64 array
    (``->su``, bootstrap.c:23-48)."""

    def __init__(self, su: np.ndarray, n: int, unfolding: int, k: int, l: int, Bg_bit: int):
        su = np.ascontiguousarray(su, dtype=np.uint64)
        assert su.shape[0] == (n // unfolding) << unfolding
        self.n, self.N = n, su.shape[3]
        self.trgsw = [HostTRGSW(su[i], l, Bg_bit) for i in range(su.shape[0])]
        self.su_array = (TRGSW * len(self.trgsw))(*[g.handle for g in self.trgsw])
        self.struct = BootstrapKeyS(None, C.cast(self.su_array, C.POINTER(TRGSW)), n, k, self.N, Bg_bit, l, unfolding)
        self.handle = C.pointer(self.struct)


class HostKSKey:
    """A ``TLWE_KS_Key`` over a ``[N_in, t, 2^base_bit-1, n_out+1]`` uint64 array."""

    def __init__(self, ksk: np.ndarray, base_bit: int):
        ksk = np.ascontiguousarray(ksk, dtype=np.uint64)
        n_in, t, bm1, w = ksk.shape
        self.rows = [[[HostTLWE(ksk[i, j, d]) for d in range(bm1)] for j in range(t)] for i in range(n_in)]
        self._lvl2 = [[(TLWE * bm1)(*[r.handle for r in self.rows[i][j]]) for j in range(t)] for i in range(n_in)]
        self._lvl1 = [(C.POINTER(TLWE) * t)(*[C.cast(a, C.POINTER(TLWE)) for a in self._lvl2[i]]) for i in range(n_in)]
        self._lvl0 = (C.POINTER(C.POINTER(TLWE)) * n_in)(*[C.cast(a, C.POINTER(C.POINTER(TLWE))) for a in self._lvl1])
        self.struct = TLWEKSKeyS(C.cast(self._lvl0, C.POINTER(C.POINTER(C.POINTER(TLWE)))), base_bit, t, n_in)
        self.handle = C.pointer(self.struct)


class HostGenericKSKey:
    """A ``Generic_KS_Key`` over a ``[n + include_b, t, 2^base_bit-1, k+1, N]`` uint64 array (uncompressed rows)."""

    def __init__(self, ksk: np.ndarray, base_bit: int, include_b: int):
        ksk = np.ascontiguousarray(ksk, dtype=np.uint64)
        ne, t, bm1 = ksk.shape[:3]
        self.rows = [[[HostTRLWE(ksk[i, j, d]) for d in range(bm1)] for j in range(t)] for i in range(ne)]
        self._lvl2 = [[(TRLWE * bm1)(*[r.handle for r in self.rows[i][j]]) for j in range(t)] for i in range(ne)]
        self._lvl1 = [(C.POINTER(TRLWE) * t)(*[C.cast(a, C.POINTER(TRLWE)) for a in self._lvl2[i]]) for i in range(ne)]
        self._lvl0 = (C.POINTER(C.POINTER(TRLWE)) * ne)(*[C.cast(a, C.POINTER(C.POINTER(TRLWE))) for a in self._lvl1])
        self.struct = GenericKSKeyS(C.cast(self._lvl0, C.POINTER(C.POINTER(C.POINTER(TRLWE)))), base_bit, t,
                                    ne - include_b, include_b)
        self.handle = C.pointer(self.struct)


class HostTRLWEKSKey:
    """A ``TRLWE_KS_Key`` over a ``[k_in, t, k_out+1, N]`` float64 array (host slot order)."""

    def __init__(self, ksk: np.ndarray, base_bit: int):
        ksk = np.ascontiguousarray(ksk, dtype=np.float64)
        k_in, t = ksk.shape[:2]
        self.rows = [[HostTRLWEDFT(ksk[i, j]) for j in range(t)] for i in range(k_in)]
        self._lvl1 = [(TRLWE_DFT * t)(*[r.handle for r in self.rows[i]]) for i in range(k_in)]
        self._lvl0 = (C.POINTER(TRLWE_DFT) * k_in)(*[C.cast(a, C.POINTER(TRLWE_DFT)) for a in self._lvl1])
        self.struct = TRLWEKSKeyS(C.cast(self._lvl0, C.POINTER(C.POINTER(TRLWE_DFT))), base_bit, t, k_in)
        self.handle = C.pointer(self.struct)


class HostTRLWEKSKeyPair:
    """The ``TRLWE_KS_Key[2]`` of ``trlwe_new_priv_KS_key`` over a ``[2, t, 2, N]`` float64 array."""

    def __init__(self, ksk2: np.ndarray, base_bit: int):
        self.keys = [HostTRLWEKSKey(ksk2[i][None], base_bit) for i in range(2)]
        self._arr = (TRLWE_KS_Key * 2)(*[k.handle for k in self.keys])
        self.handle = C.cast(self._arr, C.POINTER(TRLWE_KS_Key))


class HostTRGSW:
    """A torus-domain ``TRGSW`` handle over a ``[(k+1)*l, (k+1), N]`` uint64 array."""

    def __init__(self, rows: np.ndarray, l: int, Bg_bit: int):
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        self.rows = [HostTRLWE(rows[r]) for r in range(rows.shape[0])]
        self._samples = (TRLWE * len(self.rows))(*[r.handle for r in self.rows])
        self.struct = TRGSWS(C.cast(self._samples, C.POINTER(TRLWE)), l, Bg_bit)
        self.handle = C.pointer(self.struct)

    def flat(self) -> np.ndarray:
        return np.stack([r.polys for r in self.rows])


def generic_ks_key_to_flat(h) -> np.ndarray:
    """Uncompressed Generic_KS_Key -> [n + include_b, t, 2^base_bit-1, k+1, N]."""
    s = h.contents
    bm1 = (1 << s.base_bit) - 1
    ne = s.n + s.include_b
    return np.stack([np.stack([np.stack([trlwe_to_flat(s.s[i][j][d]) for d in range(bm1)]) for j in range(s.t)])
                     for i in range(ne)])


def trlwe_ks_key_to_flat(h) -> np.ndarray:
    """TRLWE_KS_Key -> [k_in, t, k_out+1, N] float64 (host slot order)."""
    s = h.contents
    return np.stack([np.stack([trlwe_dft_to_flat(s.s[i][j]) for j in range(s.t)]) for i in range(s.k)])


def handle_array(handles, ctype):
    """ctypes array of handles (``TLWE *`` etc.) from a list of ``Host*`` objects or raw handles."""
    hs = [getattr(h, "handle", h) for h in handles]
    return (ctype * len(hs))(*hs)


# ---------------------------------------------------------------------------------------------
# Flatteners: pointer tree produced by a C library  ->  numpy (copies)
# ---------------------------------------------------------------------------------------------
def tlwe_to_flat(h) -> np.ndarray:
    s = h.contents
    out = np.empty(s.n + 1, np.uint64)
    out[: s.n] = np.ctypeslib.as_array(s.a, shape=(s.n,))
    out[s.n] = s.b
    return out


def trlwe_to_flat(h) -> np.ndarray:
    s = h.contents
    N = s.b.contents.N
    out = np.empty((s.k + 1, N), np.uint64)
    for i in range(s.k):
        out[i] = np.ctypeslib.as_array(s.a[i].contents.coeffs, shape=(N,))
    out[s.k] = np.ctypeslib.as_array(s.b.contents.coeffs, shape=(N,))
    return out


def trlwe_dft_to_flat(h) -> np.ndarray:
    s = h.contents
    N = s.b.contents.N
    out = np.empty((s.k + 1, N), np.float64)
    for i in range(s.k):
        out[i] = np.ctypeslib.as_array(s.a[i].contents.coeffs, shape=(N,))
    out[s.k] = np.ctypeslib.as_array(s.b.contents.coeffs, shape=(N,))
    return out


def trgsw_dft_to_flat(h, k: int) -> np.ndarray:
    s = h.contents
    rows = (k + 1) * s.l
    return np.stack([trlwe_dft_to_flat(s.samples[r]) for r in range(rows)])


def trgsw_to_flat(h, k: int) -> np.ndarray:
    s = h.contents
    rows = (k + 1) * s.l
    return np.stack([trlwe_to_flat(s.samples[r]) for r in range(rows)])


def bootstrap_key_to_flat(h) -> np.ndarray:
    s = h.contents
    return np.stack([trgsw_dft_to_flat(s.s[i], s.k) for i in range(s.n)])


def ks_key_to_flat(h) -> np.ndarray:
    s = h.contents
    bm1 = (1 << s.base_bit) - 1
    n_out = s.s[0][0][0].contents.n
    out = np.empty((s.n, s.t, bm1, n_out + 1), np.uint64)
    for i in range(s.n):
        for j in range(s.t):
            for d in range(bm1):
                out[i, j, d] = tlwe_to_flat(s.s[i][j][d])
    return out
