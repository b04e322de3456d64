"""Oracle (TEST INFRASTRUCTURE): an independent reader of the Kaldi binary model files the reference loads
(SURVEY.md Appendix B).  Deliberately NOT the product's `fakebob_b200.kaldi_io`: the oracle must not inherit a reader bug.

Layout restated from upstream Kaldi's io-funcs / matrix WriteBinary conventions: a binary file starts with "\\0B"; tokens are
ASCII words followed by one space; basic ints are written as a size byte (4) followed by the little-endian value; a float
vector is "FV " + int32 dim + data, a float matrix "FM " + rows + cols + row-major data, doubles "DV " / "DM ", packed
symmetric matrices "FP " / "DP " + dim + the lower triangle row by row.
  DiagGmm          <DiagGMM> <GCONSTS> FV <WEIGHTS> FV <MEANS_INVVARS> FM <INV_VARS> FM </DiagGMM>
  FullGmm          <FullGMM> <GCONSTS> FV <WEIGHTS> FV <MEANS_INVCOVARS> FM <INV_COVARS> C x FP </FullGMM>
  IvectorExtractor <IvectorExtractor> <w> DM <w_vec> DV <M> int32 C, C x DM <SigmaInv> C x DP <IvectorOffset> double </IvectorExtractor>
  Plda             <Plda> DV mean, DM transform, DV psi </Plda>
"""
import struct

import numpy as np


class _Stream:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        self.i = 0
        if self.b[:2] != b"\0B":
            raise ValueError("%s is not a Kaldi binary file" % path)
        self.i = 2

    def token(self):
        j = self.b.index(b" ", self.i)
        t = self.b[self.i:j].decode("ascii")
        self.i = j + 1
        return t

    def expect(self, tok):
        t = self.token()
        if t != tok:
            raise ValueError("expected %s, found %s" % (tok, t))

    def peek(self):
        j = self.b.index(b" ", self.i)
        return self.b[self.i:j].decode("ascii")

    def int32(self):
        if self.b[self.i] != 4:
            raise ValueError("expected a 4-byte integer")
        v = struct.unpack_from("<i", self.b, self.i + 1)[0]
        self.i += 5
        return v

    def float64(self):
        if self.b[self.i] != 8:
            raise ValueError("expected an 8-byte float")
        v = struct.unpack_from("<d", self.b, self.i + 1)[0]
        self.i += 9
        return v

    def _data(self, n, kind):
        dt = np.dtype("<f4") if kind == "F" else np.dtype("<f8")
        a = np.frombuffer(self.b, dtype=dt, count=n, offset=self.i).copy()
        self.i += n * dt.itemsize
        return a

    def vector(self):
        t = self.token()
        if t not in ("FV", "DV"):
            raise ValueError("expected a vector, found %s" % t)
        return self._data(self.int32(), t[0])

    def matrix(self):
        t = self.token()
        if t not in ("FM", "DM"):
            raise ValueError("expected a matrix, found %s" % t)
        r, c = self.int32(), self.int32()
        return self._data(r * c, t[0]).reshape(r, c)

    def packed(self):
        t = self.token()
        if t not in ("FP", "DP"):
            raise ValueError("expected a packed matrix, found %s" % t)
        d = self.int32()
        tri = self._data(d * (d + 1) // 2, t[0])
        full = np.zeros((d, d), dtype=tri.dtype)
        full[np.tril_indices(d)] = tri
        return full + np.tril(full, -1).T


def read_diag_gmm(path):
    s = _Stream(path)
    s.expect("<DiagGMM>")
    out = {}
    s.expect("<GCONSTS>"); out["gconsts"] = s.vector()
    s.expect("<WEIGHTS>"); out["weights"] = s.vector()
    s.expect("<MEANS_INVVARS>"); out["means_invvars"] = s.matrix()
    s.expect("<INV_VARS>"); out["inv_vars"] = s.matrix()
    s.expect("</DiagGMM>")
    return out


def read_full_gmm(path):
    s = _Stream(path)
    s.expect("<FullGMM>")
    out = {}
    s.expect("<GCONSTS>"); out["gconsts"] = s.vector()
    s.expect("<WEIGHTS>"); out["weights"] = s.vector()
    s.expect("<MEANS_INVCOVARS>"); out["means_invcovars"] = s.matrix()
    s.expect("<INV_COVARS>")
    out["inv_covars"] = np.stack([s.packed() for _ in range(out["weights"].shape[0])])
    s.expect("</FullGMM>")
    return out


def read_ivector_extractor(path):
    s = _Stream(path)
    s.expect("<IvectorExtractor>")
    s.expect("<w>"); w = s.matrix()
    s.expect("<w_vec>"); w_vec = s.vector()
    s.expect("<M>")
    C = s.int32()
    M = np.stack([s.matrix() for _ in range(C)])
    s.expect("<SigmaInv>")
    sig = np.stack([s.packed() for _ in range(C)])
    s.expect("<IvectorOffset>")
    off = s.float64()
    s.expect("</IvectorExtractor>")
    return {"w": w, "w_vec": w_vec, "M": M, "sigma_inv": sig, "prior_offset": off}


def read_plda(path):
    s = _Stream(path)
    s.expect("<Plda>")
    mean, transform, psi = s.vector(), s.matrix(), s.vector()
    s.expect("</Plda>")
    return {"mean": mean, "transform": transform, "psi": psi}


def read_vector(path):
    return _Stream(path).vector()


def read_matrix(path):
    return _Stream(path).matrix()
