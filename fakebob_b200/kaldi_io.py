"""Readers / writers for the Kaldi on-disk objects the reference's model files use.

The reference never parses these itself -- it hands the paths to Kaldi binaries
(``gmm_ubm_kaldiHelper.py:206-208``: ``<spk>-identity.gmm`` / ``final.dubm``;
``ivector_PLDA_kaldiHelper.py:202,262-266``: ``final.ubm``, ``final.ie``, ``plda``,
``mean.vec``, ``transform.mat``; enrolled i-vectors via scp targets written at
``build_spk_models.py:141-149``).  The B200 path loads them once at construction
and keeps the parameters resident in HBM, so it needs its own parser
(SURVEY.md Appendix B).  Writers exist so synthetic ``pre-models/`` trees can be
produced with the same format real Kaldi would write.
"""
import io
import struct

import numpy as np


class KaldiFormatError(ValueError):
    pass


# ----------------------------------------------------------------------------- low level
class _In:
    def __init__(self, f):
        self.f = f
        self.binary = False

    def read_header(self):
        pos = self.f.tell()
        h = self.f.read(2)
        if h == b"\0B":
            self.binary = True
        else:
            self.binary = False
            self.f.seek(pos)
        return self.binary

    def peek(self, n=1):
        pos = self.f.tell()
        b = self.f.read(n)
        self.f.seek(pos)
        return b

    def _skip_ws(self):
        while True:
            c = self.f.read(1)
            if not c:
                return
            if c not in b" \t\r\n":
                self.f.seek(-1, io.SEEK_CUR)
                return

    def token(self):
        self._skip_ws()
        out = bytearray()
        while True:
            c = self.f.read(1)
            if not c or c in b" \t\r\n":
                break
            out += c
        return out.decode("ascii")

    def expect(self, *toks):
        t = self.token()
        if t not in toks:
            raise KaldiFormatError("expected %s, got %r" % ("/".join(toks), t))
        return t

    def int32(self):
        if self.binary:
            sz = self.f.read(1)
            if sz != b"\x04":
                raise KaldiFormatError("bad int32 size byte %r" % sz)
            return struct.unpack("<i", self.f.read(4))[0]
        return int(self.token())

    def float_basic(self):
        """WriteBasicType(float|double)."""
        if self.binary:
            sz = self.f.read(1)[0]
            if sz == 4:
                return struct.unpack("<f", self.f.read(4))[0]
            if sz == 8:
                return struct.unpack("<d", self.f.read(8))[0]
            raise KaldiFormatError("bad float size byte %d" % sz)
        return float(self.token())

    def _dtype(self, tok):
        if tok[0] == "F":
            return np.dtype("<f4")
        if tok[0] == "D":
            return np.dtype("<f8")
        raise KaldiFormatError("bad type token %r" % tok)

    def vector(self):
        if self.binary:
            tok = self.expect("FV", "DV")
            dt = self._dtype(tok)
            n = self.int32()
            return np.frombuffer(self.f.read(n * dt.itemsize), dtype=dt).copy()
        self.expect("[")
        vals = []
        while True:
            t = self.token()
            if t == "]":
                break
            if t == "":
                raise KaldiFormatError("unterminated text vector")
            vals.append(float(t))
        return np.asarray(vals, dtype=np.float64)

    def matrix(self):
        if self.binary:
            tok = self.expect("FM", "DM")
            dt = self._dtype(tok)
            r = self.int32()
            c = self.int32()
            return np.frombuffer(self.f.read(r * c * dt.itemsize), dtype=dt).reshape(r, c).copy()
        rows = self._text_rows()
        if not rows:
            return np.zeros((0, 0), dtype=np.float64)
        return np.asarray(rows, dtype=np.float64)

    def _text_rows(self):
        """Text matrix body: '[' then rows separated by newlines, terminated by ']'."""
        self.expect("[")
        buf = bytearray()
        while True:
            c = self.f.read(1)
            if not c:
                raise KaldiFormatError("unterminated text matrix")
            if c == b"]":
                break
            buf += c
        rows = []
        for line in buf.decode("ascii").split("\n"):
            vals = line.split()
            if vals:
                rows.append([float(v) for v in vals])
        return rows

    def sp_matrix(self):
        """Packed symmetric matrix -> full (n, n)."""
        if self.binary:
            tok = self.expect("FP", "DP")
            dt = self._dtype(tok)
            n = self.int32()
            packed = np.frombuffer(self.f.read(n * (n + 1) // 2 * dt.itemsize), dtype=dt)
        else:
            m = self.matrix_ragged()
            n = len(m)
            packed = np.concatenate([np.asarray(r, dtype=np.float64) for r in m]) if n else np.zeros(0)
        full = np.zeros((n, n), dtype=packed.dtype)
        il = np.tril_indices(n)
        full[il] = packed
        full.T[il] = packed
        return full

    def matrix_ragged(self):
        return self._text_rows()


class _Out:
    def __init__(self, f, binary=True):
        self.f = f
        self.binary = binary
        if binary:
            f.write(b"\0B")

    def token(self, t):
        self.f.write(t.encode("ascii") + b" ")

    def int32(self, v):
        if self.binary:
            self.f.write(b"\x04" + struct.pack("<i", int(v)))
        else:
            self.f.write(("%d " % v).encode())

    def double(self, v):
        if self.binary:
            self.f.write(b"\x08" + struct.pack("<d", float(v)))
        else:
            self.f.write(("%.17g " % v).encode())

    def vector(self, v, double=False):
        dt = np.dtype("<f8") if double else np.dtype("<f4")
        v = np.ascontiguousarray(v, dtype=dt).reshape(-1)
        if self.binary:
            self.token("DV" if double else "FV")
            self.int32(v.shape[0])
            self.f.write(v.tobytes())
        else:
            self.f.write(b" [ " + " ".join(_fmt(x) for x in v).encode() + b" ]\n")

    def matrix(self, m, double=False):
        dt = np.dtype("<f8") if double else np.dtype("<f4")
        m = np.ascontiguousarray(m, dtype=dt)
        if m.size == 0:
            m = m.reshape(0, 0)
        if self.binary:
            self.token("DM" if double else "FM")
            self.int32(m.shape[0])
            self.int32(m.shape[1])
            self.f.write(m.tobytes())
        else:
            self.f.write(b" [")
            for r in m:
                self.f.write(b"\n  " + " ".join(_fmt(x) for x in r).encode())
            self.f.write(b" ]\n")

    def sp_matrix(self, m, double=False):
        dt = np.dtype("<f8") if double else np.dtype("<f4")
        m = np.asarray(m)
        n = m.shape[0]
        packed = np.ascontiguousarray(m[np.tril_indices(n)], dtype=dt)
        if not self.binary:
            raise NotImplementedError("text SpMatrix writer")
        self.token("DP" if double else "FP")
        self.int32(n)
        self.f.write(packed.tobytes())


def _fmt(x):
    """Kaldi text output uses ostream precision(7)."""
    return "%.7g" % float(x)


# ----------------------------------------------------------------------------- objects
def read_diag_gmm(path):
    """-> dict(weights, means_invvars, inv_vars, gconsts) float32."""
    with open(path, "rb") as f:
        s = _In(f)
        s.read_header()
        s.expect("<DiagGMM>", "<DiagGMMBegin>")
        out = {}
        t = s.token()
        if t == "<GCONSTS>":
            out["gconsts"] = s.vector().astype(np.float32)
            t = s.token()
        if t != "<WEIGHTS>":
            raise KaldiFormatError("expected <WEIGHTS>, got %r" % t)
        out["weights"] = s.vector().astype(np.float32)
        s.expect("<MEANS_INVVARS>")
        out["means_invvars"] = s.matrix().astype(np.float32)
        s.expect("<INV_VARS>")
        out["inv_vars"] = s.matrix().astype(np.float32)
        s.expect("</DiagGMM>", "<DiagGMMEnd>")
    if "gconsts" not in out:
        out["gconsts"] = diag_gconsts(out["weights"], out["means_invvars"], out["inv_vars"])
    return out


def diag_gconsts(weights, means_invvars, inv_vars):
    iv = np.asarray(inv_vars, dtype=np.float64)
    miv = np.asarray(means_invvars, dtype=np.float64)
    D = iv.shape[1]
    gc = np.log(np.asarray(weights, dtype=np.float64)) - 0.5 * D * np.log(2 * np.pi)
    gc = gc + (0.5 * np.log(iv) - 0.5 * miv * miv / iv).sum(axis=1)
    return gc.astype(np.float32)


def write_diag_gmm(path, weights, means_invvars, inv_vars, gconsts=None, binary=True):
    if gconsts is None:
        gconsts = diag_gconsts(weights, means_invvars, inv_vars)
    with open(path, "wb") as f:
        o = _Out(f, binary)
        o.token("<DiagGMM>")
        o.token("<GCONSTS>"); o.vector(gconsts)
        o.token("<WEIGHTS>"); o.vector(weights)
        o.token("<MEANS_INVVARS>"); o.matrix(means_invvars)
        o.token("<INV_VARS>"); o.matrix(inv_vars)
        o.token("</DiagGMM>")


def read_full_gmm(path):
    """-> dict(weights (C), means_invcovars (C,D), inv_covars (C,D,D), gconsts (C)) float32."""
    with open(path, "rb") as f:
        s = _In(f)
        s.read_header()
        s.expect("<FullGMM>", "<FullGMMBegin>")
        out = {}
        t = s.token()
        if t == "<GCONSTS>":
            out["gconsts"] = s.vector().astype(np.float32)
            t = s.token()
        if t != "<WEIGHTS>":
            raise KaldiFormatError("expected <WEIGHTS>, got %r" % t)
        out["weights"] = s.vector().astype(np.float32)
        s.expect("<MEANS_INVCOVARS>")
        out["means_invcovars"] = s.matrix().astype(np.float32)
        s.expect("<INV_COVARS>")
        C = out["weights"].shape[0]
        out["inv_covars"] = np.stack([s.sp_matrix().astype(np.float32) for _ in range(C)])
        s.expect("</FullGMM>", "<FullGMMEnd>")
    return out


def write_full_gmm(path, weights, means_invcovars, inv_covars, gconsts):
    with open(path, "wb") as f:
        o = _Out(f, True)
        o.token("<FullGMM>")
        o.token("<GCONSTS>"); o.vector(gconsts)
        o.token("<WEIGHTS>"); o.vector(weights)
        o.token("<MEANS_INVCOVARS>"); o.matrix(means_invcovars)
        o.token("<INV_COVARS>")
        for ic in inv_covars:
            o.sp_matrix(ic)
        o.token("</FullGMM>")


def read_ivector_extractor(path):
    """-> dict(w (C,R)|empty, w_vec (C), M (C,D,R), sigma_inv (C,D,D), prior_offset) float64."""
    with open(path, "rb") as f:
        s = _In(f)
        s.read_header()
        s.expect("<IvectorExtractor>")
        s.expect("<w>")
        w = s.matrix()
        s.expect("<w_vec>")
        w_vec = s.vector()
        s.expect("<M>")
        C = s.int32()
        M = np.stack([s.matrix().astype(np.float64) for _ in range(C)])
        s.expect("<SigmaInv>")
        sig = np.stack([s.sp_matrix().astype(np.float64) for _ in range(C)])
        s.expect("<IvectorOffset>")
        off = s.float_basic()
        s.expect("</IvectorExtractor>")
    return {"w": w, "w_vec": w_vec.astype(np.float64), "M": M, "sigma_inv": sig, "prior_offset": float(off)}


def write_ivector_extractor(path, w_vec, M, sigma_inv, prior_offset):
    with open(path, "wb") as f:
        o = _Out(f, True)
        o.token("<IvectorExtractor>")
        o.token("<w>"); o.matrix(np.zeros((0, 0)), double=True)
        o.token("<w_vec>"); o.vector(w_vec, double=True)
        o.token("<M>"); o.int32(len(M))
        for m in M:
            o.matrix(m, double=True)
        o.token("<SigmaInv>")
        for sinv in sigma_inv:
            o.sp_matrix(sinv, double=True)
        o.token("<IvectorOffset>"); o.double(prior_offset)
        o.token("</IvectorExtractor>")


def read_plda(path):
    """-> dict(mean (L), transform (L,L), psi (L)) float64."""
    with open(path, "rb") as f:
        s = _In(f)
        s.read_header()
        s.expect("<Plda>")
        mean = s.vector().astype(np.float64)
        transform = s.matrix().astype(np.float64)
        psi = s.vector().astype(np.float64)
        s.expect("</Plda>")
    return {"mean": mean, "transform": transform, "psi": psi}


def write_plda(path, mean, transform, psi):
    with open(path, "wb") as f:
        o = _Out(f, True)
        o.token("<Plda>")
        o.vector(mean, double=True)
        o.matrix(transform, double=True)
        o.vector(psi, double=True)
        o.token("</Plda>")


def read_vector(path, offset=None):
    """A Kaldi vector file, or the vector at ``offset`` inside an ark (scp target 'path:offset')."""
    if offset is None and ":" in path:
        head, tail = path.rsplit(":", 1)
        if tail.isdigit():
            path, offset = head, int(tail)
    with open(path, "rb") as f:
        if offset:
            f.seek(offset)
        s = _In(f)
        s.read_header()
        return s.vector()


def write_vector(path, v, binary=True, double=False):
    with open(path, "wb") as f:
        _Out(f, binary).vector(v, double=double)


def read_matrix(path):
    with open(path, "rb") as f:
        s = _In(f)
        s.read_header()
        return s.matrix()


def write_matrix(path, m, binary=True, double=False):
    with open(path, "wb") as f:
        _Out(f, binary).matrix(m, double=double)


def write_text_vector_ark(path, items):
    """'ark,t' vector table: 'utt  [ v0 v1 ... ]\\n'.  Returns {utt: 'path:offset'} scp targets."""
    targets = {}
    with open(path, "wb") as f:
        for utt, v in items:
            f.write(utt.encode("ascii") + b" ")
            targets[utt] = "%s:%d" % (path, f.tell())
            f.write(b" [ " + " ".join(_fmt(x) for x in np.asarray(v).reshape(-1)).encode() + b" ]\n")
    return targets


def parse_conf(path):
    """Kaldi --config file: one '--name=value' per line, '#' comments."""
    out = {}
    with open(path, "r") as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            if not line:
                continue
            for tok in line.split():
                if not tok.startswith("--"):
                    raise KaldiFormatError("bad option %r in %s" % (tok, path))
                k, _, v = tok[2:].partition("=")
                out[k.replace("_", "-")] = v
    return out
