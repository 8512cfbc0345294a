"""Fiat-Shamir transcript of the prover (host side of the device-resident prover).

Mirrors ``src/transcript.rs:8-86`` of the reference: a merlin transcript with the domain separator
``b"plonk"`` (src/prover.rs:112), points appended as 48-byte compressed G1, scalars as 32 little-endian
bytes, challenges drawn 32 bytes at a time and rejected until they decode to a non-zero canonical
Scalar, then appended back under the same label.

merlin 3.0.0 is a Cargo dependency of the reference, not part of its tree; what is implemented here is its
published construction: STROBE-128 (rate 166, operations meta-AD / AD / PRF) over keccak-f[1600].
A few hundred bytes per proof pass through this, so it stays on the host.  The permutation itself is the library's
``bpk_keccak_f1600`` (C++, host): interpreted, the ~25 permutations of a proof cost ~10 ms, a tenth of a 2^20-gate
proof.  ``keccak_f1600`` below is the same permutation in plain Python, kept as the cross-check of the native one.
"""
from __future__ import annotations

import ctypes

FR_MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

_MASK = 0xFFFFFFFFFFFFFFFF
_ROUND_CONSTANTS = []
_RHO_OFFSETS = [0] * 25   # rotation of lane x + 5 y
_PI_DEST = [0] * 25       # lane x + 5 y moves to _PI_DEST[x + 5 y]


def _init_tables():
    # round constants from the degree-8 LFSR x^8 + x^6 + x^5 + x^4 + 1 (FIPS 202, algorithm 5)
    lfsr = 1
    for _ in range(24):
        rc = 0
        for j in range(7):
            if lfsr & 1:
                rc |= 1 << ((1 << j) - 1)
            lfsr <<= 1
            if lfsr & 0x100:
                lfsr ^= 0x171
        _ROUND_CONSTANTS.append(rc)
    # rho offsets: walk (x, y) -> (y, 2x + 3y) from (1, 0); offset t-th step = (t+1)(t+2)/2
    x, y = 1, 0
    for t in range(24):
        _RHO_OFFSETS[x + 5 * y] = ((t + 1) * (t + 2) // 2) % 64
        x, y = y, (2 * x + 3 * y) % 5
    for x in range(5):
        for y in range(5):
            _PI_DEST[x + 5 * y] = y + 5 * ((2 * x + 3 * y) % 5)


_init_tables()


def keccak_f1600(lanes: list) -> list:
    """the 24-round permutation on 25 little-endian 64-bit lanes, lane index x + 5 y"""
    s = list(lanes)
    for rc in _ROUND_CONSTANTS:
        col = [s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20] for x in range(5)]
        for x in range(5):
            r = col[(x + 1) % 5]
            d = col[(x + 4) % 5] ^ (((r << 1) | (r >> 63)) & _MASK)
            for y in range(0, 25, 5):
                s[x + y] ^= d
        moved = [0] * 25
        for i in range(25):
            k = _RHO_OFFSETS[i]
            v = s[i]
            moved[_PI_DEST[i]] = ((v << k) | (v >> (64 - k))) & _MASK if k else v
        for y in range(0, 25, 5):
            row = moved[y:y + 5]
            for x in range(5):
                s[x + y] = row[x] ^ ((~row[(x + 1) % 5]) & _MASK & row[(x + 2) % 5])
        s[0] ^= rc
    return s


_native = None


def _native_keccak():
    """bpk_keccak_f1600 from libbpk.so (dlopen works without a GPU); None if the library is not built"""
    global _native
    if _native is None:
        try:
            from . import load_library
            _native = load_library().bpk_keccak_f1600
        except Exception:
            _native = False
    return _native or None


class _Strobe:
    """STROBE-128/1600 duplex restricted to the three operations merlin uses"""
    RATE = 166
    I, A, C, T, M, K = 1, 2, 4, 8, 16, 32

    def __init__(self, protocol: bytes):
        self.buf = bytearray(200)
        self.buf[:18] = bytes([1, self.RATE + 2, 1, 0, 1, 96]) + b"STROBEv1.0.2"
        self._permute()
        self.pos = 0
        self.begin = 0
        self.flags = 0
        self.operate(self.M | self.A, protocol)

    def _permute(self):
        fn = _native_keccak()
        if fn is not None:
            fn((ctypes.c_uint8 * 200).from_buffer(self.buf))
            return
        lanes = [int.from_bytes(self.buf[8 * i:8 * i + 8], "little") for i in range(25)]
        self.buf = bytearray(b"".join(v.to_bytes(8, "little") for v in keccak_f1600(lanes)))

    def _end_block(self):
        self.buf[self.pos] ^= self.begin
        self.buf[self.pos + 1] ^= 0x04
        self.buf[self.RATE + 1] ^= 0x80
        self._permute()
        self.pos = 0
        self.begin = 0

    def _duplex(self, data: bytes, squeeze: int = 0) -> bytes:
        out = bytearray()
        for k in range(squeeze if squeeze else len(data)):
            if squeeze:
                out.append(self.buf[self.pos])
                self.buf[self.pos] = 0
            else:
                self.buf[self.pos] ^= data[k]
            self.pos += 1
            if self.pos == self.RATE:
                self._end_block()
        return bytes(out)

    def operate(self, flags: int, data: bytes = b"", more: bool = False, squeeze: int = 0) -> bytes:
        if more:
            if flags != self.flags:
                raise ValueError("continued STROBE operation with different flags")
        else:
            previous = self.begin
            self.begin = self.pos + 1
            self.flags = flags
            self._duplex(bytes([previous, flags]))
            if flags & (self.C | self.K) and self.pos:
                self._end_block()
        return self._duplex(data, squeeze)


class Transcript:
    """merlin::Transcript::new(label) with append_message / challenge_bytes"""

    def __init__(self, label: bytes):
        self._strobe = _Strobe(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def _framed(self, label: bytes, length: int):
        s = self._strobe
        s.operate(s.M | s.A, label)
        s.operate(s.M | s.A, length.to_bytes(4, "little"), more=True)

    def append_message(self, label: bytes, message: bytes):
        self._framed(label, len(message))
        self._strobe.operate(self._strobe.A, message)

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        self._framed(label, n)
        s = self._strobe
        return s.operate(s.I | s.A | s.C, squeeze=n)


class PlonkTranscript(Transcript):
    """src/transcript.rs:65-86"""

    def __init__(self):
        super().__init__(b"plonk")  # src/prover.rs:112

    def append_point(self, label: bytes, compressed48: bytes):
        assert len(compressed48) == 48
        self.append_message(label, compressed48)

    def append_scalar(self, label: bytes, value: int):
        self.append_message(label, int(value).to_bytes(32, "little"))

    def get_and_append_challenge(self, label: bytes) -> int:
        while True:
            raw = self.challenge_bytes(label, 32)
            v = int.from_bytes(raw, "little")
            if 0 < v < FR_MODULUS:   # Scalar::from_bytes(..).is_some() && != 0
                self.append_message(label, raw)
                return v
