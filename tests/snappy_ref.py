"""TEST INFRASTRUCTURE: a small pure-Python writer of the Snappy framing format, written from the
published format descriptions, used to make `map.collated.rad.sz` inputs for the host's decoder
(no Snappy library exists in this image). Greedy hash-table matcher: emits literals and all three
copy element types, so the decoder's every branch is exercised."""
import struct

_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = (c >> 8) ^ _CRC_TABLE[(c ^ b) & 0xFF]
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(n):
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def _literal(chunk: bytes) -> bytes:
    n = len(chunk) - 1
    if n < 60:
        return bytes([n << 2]) + chunk
    nb = (n.bit_length() + 7) // 8
    return bytes([(59 + nb) << 2]) + n.to_bytes(nb, "little") + chunk


def _copy(offset: int, length: int, force4: bool) -> bytes:
    out = bytearray()
    while length > 0:
        l = min(length, 64)
        if length - l in (1, 2, 3):      # never leave a tail shorter than the 1-byte-offset form's minimum
            l -= 4
        if force4:
            out += bytes([((l - 1) << 2) | 3]) + struct.pack("<I", offset)
        elif 4 <= l <= 11 and offset < 2048:
            out += bytes([((offset >> 8) << 5) | ((l - 4) << 2) | 1, offset & 0xFF])
        else:
            out += bytes([((l - 1) << 2) | 2]) + struct.pack("<H", offset)
        length -= l
    return bytes(out)


def compress_block(data: bytes, force4: bool = False) -> bytes:
    """one raw snappy block (<= 65536 input bytes)"""
    out = bytearray(_varint(len(data)))
    table = {}
    i = lit = 0
    n = len(data)
    while i + 4 <= n:
        key = data[i:i + 4]
        j = table.get(key)
        table[key] = i
        if j is not None and i - j <= 65535:
            m = 4
            while i + m < n and data[j + m] == data[i + m]:     # may run past i: overlapping copy (a run)
                m += 1
            if lit < i:
                out += _literal(data[lit:i])
            out += _copy(i - j, m, force4)
            i += m
            lit = i
        else:
            i += 1
    if lit < n:
        out += _literal(data[lit:])
    return bytes(out)


def frame(data: bytes, block: int = 65536, store_every: int = 0, force4: bool = False, padding: bool = False) -> bytes:
    """Snappy framing format: stream identifier, then one data chunk per `block` bytes (type 0x00
    compressed, or 0x01 stored for every `store_every`-th chunk), optional padding / skippable chunks"""
    out = bytearray(b"\xff\x06\x00\x00sNaPpY")
    k = 0
    for o in range(0, len(data), block):
        chunk = data[o:o + block]
        k += 1
        if store_every and k % store_every == 0:
            body, typ = chunk, 1
        else:
            body, typ = compress_block(chunk, force4), 0
        payload = struct.pack("<I", masked_crc(chunk)) + body
        out += bytes([typ]) + len(payload).to_bytes(3, "little") + payload
        if padding and k % 3 == 0:
            out += b"\xfe\x05\x00\x00" + b"\x00" * 5 + b"\x80\x02\x00\x00zz"
    return bytes(out)
