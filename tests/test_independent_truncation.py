"""What the reference's iterator does with TRUNCATED streams, restated a second time in pure Python with the exact
error plumbing of the Go code -- BytesReader64 word reads (x/xio/reader64.go:40-81), IStream.ReadBits / PeekBits
(encoding/istream.go:71-125: a failed read still consumes the reader's last word), the swallowed errors of
readDeltaOfDelta (m3tsz/timestamp_iterator.go:265-302), `res, it.err = ReadBits(..)` overwriting an earlier error
(m3tsz/iterator.go:221-224), Next() = hasNext() after the value read (iterator.go:81-106) -- and compared with the
C oracle's iterator on EVERY byte-truncation of golden, fixture and generated streams, both modes: number of
datapoints returned, their values, and the final Err()."""
import base64
import json
import os
import random
import struct

import numpy as np
import pytest

import oracle_lib as O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "m3tsz_goldens.json")))
SEC = 10 ** 9
EOF, NO_SCHEME, INVALID_MULT, ANN_LEN, ANN_SHORT, VARINT_OVERFLOW, UNEXPECTED_EOF = 1, 5, 7, 8, 9, 11, 12
M64 = (1 << 64) - 1
UNIT_NANOS = {1: 10 ** 9, 2: 10 ** 6, 3: 10 ** 3, 4: 1}
DEFAULT_BITS = {1: 32, 2: 32, 3: 64, 4: 64}


class IStream:
    def __init__(self, data):
        self.data, self.index = data, 0
        self.current, self.remaining = 0, 0

    def _read64(self):
        d, i = self.data, self.index
        if i + 8 <= len(d):
            self.index += 8
            return int.from_bytes(d[i:i + 8], "big"), 8, 0
        if i >= len(d):
            return 0, 0, EOF
        n = len(d) - i
        self.index = len(d)
        return int.from_bytes(d[i:], "big") << (64 - 8 * n), n, 0

    def _peek64(self):
        d, i = self.data, self.index
        if i + 8 <= len(d):
            return int.from_bytes(d[i:i + 8], "big"), 8, 0
        if i >= len(d):
            return 0, 0, EOF
        n = len(d) - i
        return int.from_bytes(d[i:], "big") << (64 - 8 * n), n, 0

    def read_bits(self, n):
        res = (self.current >> (64 - n)) if n else 0
        rem = self.remaining
        if n <= rem:
            self.current = (self.current << n) & M64
            self.remaining -= n
            return res, 0
        need = n - rem
        cur, nb, err = self._read64()
        if err:
            return 0, err
        nb *= 8
        if nb < need:
            return 0, EOF  # (the word just read is gone)
        self.current = (cur << need) & M64
        self.remaining = nb - need
        return res | (cur >> (64 - need)), 0

    def peek_bits(self, n):
        res = (self.current >> (64 - n)) if n else 0
        if n <= self.remaining:
            return res, 0
        need = n - self.remaining
        nxt, nb, err = self._peek64()
        if err:
            return 0, err
        if 8 * nb < need:
            return 0, EOF
        return res | (nxt >> (64 - need)), 0


def _sx(v, bits):
    return v - (1 << bits) if (v >> (bits - 1)) & 1 else v


class GoIterator:
    def __init__(self, data, int_opt, default_unit=1):
        self.s = IStream(bytes(data))
        self.int_opt, self.default_unit = int_opt, default_unit
        self.err, self.done = 0, False
        self.prev_time = self.prev_delta = 0
        self.unit, self.unit_changed, self.has_scheme = 0, False, False
        self.prev_bits = self.prev_xor = 0
        self.int_val, self.mult, self.sig, self.is_float = 0.0, 0, 0, False
        self.cur = None

    # ---- TimestampIterator ----
    def _read_varint(self):
        x, s = 0, 0
        for i in range(10):
            b, err = self.s.read_bits(8)
            if err:
                if i > 0 and err == EOF:
                    err = UNEXPECTED_EOF
                return 0, err
            if b < 0x80:
                if i == 9 and b > 1:
                    return 0, VARINT_OVERFLOW
                ux = x | (b << s)
                v = ux >> 1
                return (~v if ux & 1 else v), 0
            x |= (b & 0x7F) << s
            s += 7
        return 0, VARINT_OVERFLOW

    def _read_annotation(self):
        n, err = self._read_varint()
        if err:
            return err
        n += 1
        if n <= 0:
            return ANN_LEN
        for _ in range(n):
            _, err = self.s.read_bits(8)
            if err:
                return err
        return 0

    def _read_time_unit(self):
        tu, err = self.s.read_bits(8)
        if err:
            return err
        if 1 <= tu <= 8 and tu != self.unit:
            self.unit_changed = True
            self.has_scheme = True
        self.unit = tu
        return 0

    def _marker_or_dod(self):
        pk, err = self.s.peek_bits(11)
        if not err and (pk >> 2) == 0x100:
            mv = pk & 3
            if mv == 0:
                _, err = self.s.read_bits(11)
                if err:
                    return 0, err
                self.done = True
                return 0, 0
            if mv == 1:
                _, err = self.s.read_bits(11)
                if err:
                    return 0, err
                err = self._read_annotation()
                if err:
                    return 0, err
                return self._marker_or_dod()
            if mv == 2:
                _, err = self.s.read_bits(11)
                if err:
                    return 0, err
                err = self._read_time_unit()
                if err:
                    return 0, err
                return self._marker_or_dod()
        return self._dod()

    def _dod(self):
        if self.unit_changed:
            if not 1 <= self.unit <= 8:
                return 0, NO_SCHEME
            v, err = self.s.read_bits(64)
            if err:
                return 0, err
            return _sx(v, 64), 0
        if not self.has_scheme:
            return 0, NO_SCHEME
        assert self.unit in DEFAULT_BITS, "test streams use s / ms / us / ns"
        cb, err = self.s.read_bits(1)
        if err:
            return 0, err
        if cb == 0:
            return 0, 0
        for opcode, vbits in ((0b10, 7), (0b110, 9), (0b1110, 12)):
            nxt, err = self.s.read_bits(1)
            if err:
                return 0, 0  # swallowed (timestamp_iterator.go:271-274)
            cb = (cb << 1) | nxt
            if cb == opcode:
                v, err = self.s.read_bits(vbits)
                if err:
                    return 0, err
                return _sx(v, vbits) * UNIT_NANOS[self.unit], 0
        vbits = DEFAULT_BITS[self.unit]
        v, err = self.s.read_bits(vbits)
        if err:
            return 0, err
        return _sx(v, vbits) * UNIT_NANOS[self.unit], 0

    def _read_timestamp(self):
        first = False
        if self.prev_time != 0:
            dod, err = self._marker_or_dod()
            if not err:
                self.prev_delta += dod
                self.prev_time += self.prev_delta
        else:
            first = True
            nt, err = self.s.read_bits(64)
            if not err:
                nt = _sx(nt, 64)
                if self.unit == 0:
                    du = self.default_unit
                    self.unit = du if (du in UNIT_NANOS and nt % UNIT_NANOS[du] == 0) else 0
                self.has_scheme = 1 <= self.unit <= 8
                dod, err = self._marker_or_dod()
                if not err:
                    self.prev_delta += dod
                    self.prev_time += self.prev_delta  # readNextTimestamp
                    self.prev_time = nt + self.prev_delta
        if err:
            return False, False, err
        if self.unit_changed:
            self.prev_delta, self.unit_changed = 0, False
        return first, self.done, 0

    # ---- readerIterator ----
    def _rb(self, n):
        res, self.err = self.s.read_bits(n)  # overwrites an earlier error
        return res

    def _full_float(self):
        v, err = self.s.read_bits(64)
        if err:
            self.err = err
            return
        self.prev_bits = self.prev_xor = v

    def _next_float(self):
        cb, err = self.s.read_bits(1)
        if err:
            self.err = err
            return
        if cb == 0:
            self.prev_xor = 0
            return
        nxt, err = self.s.read_bits(1)
        if err:
            self.err = err
            return
        if nxt == 0:
            x = self.prev_xor
            pl, pt = (64, 0) if x == 0 else (64 - x.bit_length(), (x & -x).bit_length() - 1)
            m, err = self.s.read_bits(64 - pl - pt)
            if err:
                self.err = err
                return
            self.prev_xor = (m << pt) & M64
            self.prev_bits ^= self.prev_xor
            return
        h, err = self.s.read_bits(12)
        if err:
            self.err = err
            return
        lead, nb = (h & 4032) >> 6, (h & 63) + 1
        m, err = self.s.read_bits(nb)
        if err:
            self.err = err
            return
        self.prev_xor = (m << (64 - lead - nb)) & M64 if 64 - lead - nb >= 0 else 0
        self.prev_bits ^= self.prev_xor

    def _sig_mult(self):
        if self._rb(1) == 1:
            if self._rb(1) == 0:
                self.sig = 0
            else:
                self.sig = (self._rb(6) + 1) & 0xFF
        if self._rb(1) == 1:
            self.mult = self._rb(3)
            if self.mult > 6:
                self.err = INVALID_MULT

    def _int_diff(self):
        if self.sig == 64:
            sign = 1.0 if self._rb(1) == 1 else -1.0
            self.int_val += sign * float(self._rb(64))
            return
        bits = self._rb(self.sig + 1)
        sign = -1.0
        if (bits >> self.sig) == 1:
            sign = 1.0
            bits ^= 1 << self.sig
        self.int_val += sign * float(bits)

    def next(self):
        if self.err or self.done:
            return False
        first, done, err = self._read_timestamp()
        if err or done:
            self.err = err
            return False
        if not self.int_opt:
            self._full_float() if first else self._next_float()
        elif first:
            if self._rb(1) == 1:
                self._full_float()
                self.is_float = True
            else:
                self._sig_mult()
                self._int_diff()
        else:
            if self._rb(1) == 0:
                if self._rb(1) == 1:
                    pass
                elif self._rb(1) == 1:
                    self._full_float()
                    self.is_float = True
                else:
                    self._sig_mult()
                    self._int_diff()
                    self.is_float = False
            elif self.is_float:
                self._next_float()
            else:
                self._int_diff()
        if not self.int_opt or self.is_float:
            vb = self.prev_bits
        else:
            assert self.mult <= 6 or self.err
            v = self.int_val if self.mult == 0 or self.mult > 6 else self.int_val / 10.0 ** self.mult
            vb = struct.unpack("<Q", struct.pack("<d", v))[0]
        self.cur = (self.prev_time, vb)
        return not self.err and not self.done


def _go_decode(data, int_opt):
    it = GoIterator(data, int_opt)
    out = []
    while it.next():
        out.append(it.cur)
    return out, it.err


def _oracle_decode(data, int_opt):
    dps, err = O.decode_all(data, int_opt)
    return [(d[0], struct.unpack("<Q", struct.pack("<d", d[1]))[0]) for d in dps], err


def _streams(int_opt):
    out = []
    if not int_opt:
        out += [bytes.fromhex(s["bytes"]) for s in G["streams"]]
    else:
        out += [base64.b64decode(b) for b in G["fixtures_b64"]["streams"][:4]]
        out.append(base64.b64decode(G["regression_b64"]["stream"]))
    r = random.Random(17 + int(int_opt))
    rng = np.random.default_rng(23 + int(int_opt))
    start = 1599955200 * SEC
    for k in range(10):
        P = 40
        ts = start + np.cumsum(rng.choice([1, 10, 60, 300, 4000, 10 ** 6], size=P)).astype(np.int64) * SEC
        if k % 3 == 0:
            vals = 100.0 + np.cumsum(rng.normal(size=P))
        elif k % 3 == 1:
            vals = np.round(rng.normal(size=P) * 10.0 ** r.randrange(0, 8), r.randrange(0, 5))
        else:
            vals = np.round(rng.normal(size=P) * 5, 2)
            vals[::7] = [np.nan, 2.0 ** 63, -2.0 ** 63, 1e300, 0.5, 12.0][: len(vals[::7])]
        out.append(O.encode_series(ts, vals, start, O.UNIT_S, int_opt))
    # with annotations and unit changes
    e = O.Encoder(1427162400 * SEC, int_opt)
    t = 1427162462 * SEC
    for i in range(30):
        unit = O.UNIT_MS if i == 0 else (O.UNIT_US if i == 10 else O.UNIT_S)
        ann = b"foo" if i < 5 else (b"bar" if i < 7 else (b"x" * 200 if i == 10 else b""))
        assert e.encode(t, float(r.randrange(1000)) / 4, unit, ann) == 0
        t += SEC * r.randrange(1, 500)
    out.append(e.stream())
    return out


@pytest.mark.parametrize("int_opt", [False, True])
def test_every_truncation_matches_oracle_iterator(int_opt):
    n_cases = n_err = 0
    for stream in _streams(int_opt):
        full, err = _go_decode(stream, int_opt)
        assert err == 0 and _oracle_decode(stream, int_opt) == (full, 0)
        for cut in range(len(stream)):
            got = _go_decode(stream[:cut], int_opt)
            exp = _oracle_decode(stream[:cut], int_opt)
            assert got == exp, (cut, len(stream), got[1], exp[1], len(got[0]), len(exp[0]))
            n_cases += 1
            n_err += got[1] != 0
    assert n_cases > 3000 and n_err > 0.5 * n_cases
