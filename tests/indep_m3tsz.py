"""A SECOND, independent restatement of the M3TSZ decoder in pure Python -- test infrastructure only.

The CPU oracle (oracle/m3tsz_oracle.c) and the CUDA kernels were written by the same hands; this decoder was
written separately, straight from the reference's iterator (m3tsz/iterator.go:81-219,
m3tsz/timestamp_iterator.go:80-361, m3tsz/float_encoder_iterator.go:105-165, scheme.go:40-242,
x/time/unit.go) without looking at the C code, and is used to cross-check the oracle's ENCODER and DECODER on
streams no golden vector of the reference covers (int-optimised encodes of arbitrary series).  Well-formed
streams only: a short read raises EOFError instead of reproducing the iterator's error swallowing."""
import struct

UNIT_NANOS = {1: 10 ** 9, 2: 10 ** 6, 3: 10 ** 3, 4: 1, 5: 60 * 10 ** 9, 6: 3600 * 10 ** 9, 7: 86400 * 10 ** 9,
              8: 365 * 86400 * 10 ** 9}  # x/time/unit.go:30-60 (second .. year)
# scheme.go:40-75: bucket value bits [7, 9, 12]; default bucket 32 bits for s / ms, 64 for us / ns
DEFAULT_BUCKET_BITS = {1: 32, 2: 32, 3: 64, 4: 64}
MARKER_OPCODE, MARKER_OPCODE_BITS, MARKER_VALUE_BITS = 0x100, 9, 2  # scheme.go:198-242
MARKER_EOS, MARKER_ANNOTATION, MARKER_TIME_UNIT = 0, 1, 2


class Bits:
    def __init__(self, data):
        self.v = int.from_bytes(data, "big")
        self.n = 8 * len(data)
        self.pos = 0

    def remaining(self):
        return self.n - self.pos

    def peek(self, k):
        if self.remaining() < k:
            return None
        return (self.v >> (self.n - self.pos - k)) & ((1 << k) - 1) if k else 0

    def read(self, k):
        r = self.peek(k)
        if r is None:
            raise EOFError("need %d bits, %d left" % (k, self.remaining()))
        self.pos += k
        return r


def _sign_extend(v, bits):
    return v - (1 << bits) if v >> (bits - 1) else v


def _initial_time_unit(start_ns, default_unit):  # timestamp_encoder.go:248-259
    if default_unit in UNIT_NANOS and start_ns % UNIT_NANOS[default_unit] == 0:
        return default_unit
    return 0


def _lz_tz(x):  # encoding.go:29-49: (64, 0) for zero
    if x == 0:
        return 64, 0
    return 64 - x.bit_length(), (x & -x).bit_length() - 1


def _read_varint(b):  # encoding/binary.ReadVarint over whole bytes read through the bit stream
    ux, shift = 0, 0
    while True:
        byte = b.read(8)
        ux |= (byte & 0x7F) << shift
        if byte < 0x80:
            break
        shift += 7
    x = ux >> 1
    return ~x if ux & 1 else x


def decode(data, int_optimized, default_unit=1):
    """-> list of (timestamp_ns, value_bits_as_uint64, unit, annotation bytes or b"")"""
    b = Bits(bytes(data))
    out = []
    prev_time = prev_delta = 0
    unit, unit_changed = 0, False
    prev_bits = prev_xor = 0
    int_val, mult, sig, is_float = 0.0, 0, 0, False
    if b.n == 0:
        return out

    def dod_or_marker():
        """timestamp_iterator.go:175-302 -> (dod_ns, done, annotation)"""
        nonlocal unit, unit_changed
        ann = b""
        while True:
            pk = b.peek(MARKER_OPCODE_BITS + MARKER_VALUE_BITS)
            if pk is not None and (pk >> MARKER_VALUE_BITS) == MARKER_OPCODE:
                mv = pk & 3
                if mv == MARKER_EOS:
                    b.read(11)
                    return 0, True, ann
                if mv == MARKER_ANNOTATION:
                    b.read(11)
                    n = _read_varint(b) + 1
                    assert n > 0
                    ann = bytes(b.read(8) for _ in range(n))
                    continue
                if mv == MARKER_TIME_UNIT:
                    b.read(11)
                    tu = b.read(8)
                    if 1 <= tu <= 8 and tu != unit:
                        unit_changed = True
                    unit = tu
                    continue
            break
        if unit_changed:  # readFullTimestamp: 64 bits of nanoseconds
            return _sign_extend(b.read(64), 64), False, ann
        assert unit in DEFAULT_BUCKET_BITS, "no time encoding scheme for unit %d" % unit
        if b.read(1) == 0:
            return 0, False, ann
        for ones, vbits in ((1, 7), (2, 9), (3, 12)):  # opcodes 10, 110, 1110
            if b.read(1) == 0:
                return _sign_extend(b.read(vbits), vbits) * UNIT_NANOS[unit], False, ann
        vbits = DEFAULT_BUCKET_BITS[unit]
        return _sign_extend(b.read(vbits), vbits) * UNIT_NANOS[unit], False, ann

    def read_full_float():
        nonlocal prev_bits, prev_xor
        prev_bits = prev_xor = b.read(64)

    def read_next_float():  # float_encoder_iterator.go:117-165
        nonlocal prev_bits, prev_xor
        if b.read(1) == 0:
            prev_xor = 0
            return
        if b.read(1) == 0:  # '10' contained
            pl, pt = _lz_tz(prev_xor)
            prev_xor = b.read(64 - pl - pt) << pt
        else:  # '11' uncontained
            hdr = b.read(12)
            lead, nbits = hdr >> 6, (hdr & 63) + 1
            prev_xor = b.read(nbits) << (64 - lead - nbits)
        prev_bits ^= prev_xor

    def read_sig_mult():  # iterator.go:178-193
        nonlocal sig, mult
        if b.read(1) == 1:
            sig = 0 if b.read(1) == 0 else b.read(6) + 1
        if b.read(1) == 1:
            mult = b.read(3)
            assert mult <= 6, "invalid multiplier"

    def read_int_diff():  # iterator.go:195-219 (sign bit 1 = add)
        nonlocal int_val
        sign = 1.0 if b.read(1) == 1 else -1.0
        int_val += sign * float(b.read(sig))

    first = True
    while True:
        if first:
            nt = _sign_extend(b.read(64), 64)
            if unit == 0:
                unit = _initial_time_unit(nt, default_unit)
            dod, done, ann = dod_or_marker()
            if done:
                break
            prev_delta += dod
            prev_time = nt + prev_delta
        else:
            dod, done, ann = dod_or_marker()
            if done:
                break
            prev_delta += dod
            prev_time += prev_delta
        if unit_changed:
            prev_delta, unit_changed = 0, False
        # ---- value ----
        if not int_optimized:
            read_full_float() if first else read_next_float()
        elif first:
            if b.read(1) == 1:
                read_full_float()
                is_float = True
            else:
                read_sig_mult()
                read_int_diff()
        else:
            if b.read(1) == 0:  # update
                if b.read(1) == 1:
                    pass  # repeat
                elif b.read(1) == 1:
                    read_full_float()
                    is_float = True
                else:
                    read_sig_mult()
                    read_int_diff()
                    is_float = False
            elif is_float:
                read_next_float()
            else:
                read_int_diff()
        if not int_optimized or is_float:
            vbits = prev_bits
        else:
            v = int_val if mult == 0 else int_val / (10.0 ** mult)
            vbits = struct.unpack("<Q", struct.pack("<d", v))[0]
        out.append((prev_time, vbits, unit, ann))
        first = False
        # the reference keys "first" on PrevTime == 0 (timestamp_iterator.go:89); same thing for real timestamps
        assert prev_time != 0
    return out
