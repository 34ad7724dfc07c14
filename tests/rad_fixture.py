"""A collated-RAD writer built ONLY from byte-level statements in the reference tree (test infrastructure) — independent
of this repo's C++ writer (afqh_write_collated_rad), so that the reader is checked against the reference's layout:

  * header: is_paired u8, ref_count u64, names as u16 length + bytes, num_chunks u64 — the LAST 8 bytes of the header
    (src/convert.rs:247-254: `end_header_pos = stream_position - size_of::<u64>()`, back-patched at :585-590)
  * three tag sections (file, read, alignment) written in that order (src/convert.rs:280-359); a section is a u16 tag
    count followed by (u16 name length, name bytes, u8 type id) per tag — libradicl's TagSection::write, type ids per
    the RAD specification (bool 0, u8 1, u16 2, u32 3, u64 4, f32 5, f64 6, array 7 (+ length-type u8, element-type u8), string 8)
  * file-tag VALUES right behind the sections, in tag order: cblen u16, ulen u16 (src/convert.rs:361-369); a string value
    is u16 length + bytes (the `known_rad_type` tag of tests/multi_barcode_integration.rs:72-75, 106-110)
  * barcode / UMI integer widths by length: 1-4 bases u8, 5-8 u16, 9-16 u32, 17-32 u64 (src/convert.rs:322-343)
  * a chunk: nbytes u32 (INCLUDING these 8 header bytes: `nbytes = data.get_ref().len()` over a buffer that starts with the two
    u32 placeholders, src/convert.rs:380-383, 472-480), nrec u32, then the records
  * a record (write_list, src/convert.rs:122-144): na u32, barcode, UMI, then na x u32 reference ids, bit 31 set = forward
    orientation (src/convert.rs:438-445)
"""
import struct

TYPE_ID = {"bool": 0, "u8": 1, "u16": 2, "u32": 3, "u64": 4, "f32": 5, "f64": 6, "array": 7, "string": 8}
FMT = {"bool": "<B", "u8": "<B", "u16": "<H", "u32": "<I", "u64": "<Q", "f32": "<f", "f64": "<d"}


def width_type(n_bases):
    if 1 <= n_bases <= 4: return "u8"
    if n_bases <= 8: return "u16"
    if n_bases <= 16: return "u32"
    if n_bases <= 32: return "u64"
    raise ValueError(n_bases)


def _s16(s):
    b = s.encode()
    return struct.pack("<H", len(b)) + b


def tag_section(tags):
    """tags: [(name, type)] or (name, 'array', length type, element type)."""
    out = struct.pack("<H", len(tags))
    for t in tags:
        out += _s16(t[0]) + struct.pack("<B", TYPE_ID[t[1]])
        if t[1] == "array":
            out += struct.pack("<BB", TYPE_ID[t[2]], TYPE_ID[t[3]])
    return out


def tag_value(t, v):
    if t[1] == "string":
        return _s16(v)
    if t[1] == "array":
        return struct.pack(FMT[t[2]], len(v)) + b"".join(struct.pack(FMT[t[3]], x) for x in v)
    return struct.pack(FMT[t[1]], v)


def write_collated_rad(path, ref_names, cells, bc_len=16, umi_len=12, extra_file_tags=(), extra_read_tags=(), extra_aln_tags=(),
                       extra_first=False, with_ulen=True):
    """cells: [(barcode, [(umi, [ref ids], [orientation forward?])...])]; one chunk per cell (collated).
    extra_*_tags: [(tag descriptor, value or per-record/per-alignment constant)] placed after (or, extra_first, before)
    the standard tags. Returns the number of bytes written."""
    bt, ut = width_type(bc_len), width_type(umi_len)
    out = struct.pack("<BQ", 0, len(ref_names)) + b"".join(_s16(n) for n in ref_names) + struct.pack("<Q", len(cells))
    std_file = [(("cblen", "u16"), bc_len)] + ([(("ulen", "u16"), umi_len)] if with_ulen else [])
    file_tags = (list(extra_file_tags) + std_file) if extra_first else (std_file + list(extra_file_tags))
    std_read = [(("b", bt), None), (("u", ut), None)]
    read_tags = (list(extra_read_tags) + std_read) if extra_first else (std_read + list(extra_read_tags))
    std_aln = [(("compressed_ori_refid", "u32"), None)]
    aln_tags = (list(extra_aln_tags) + std_aln) if extra_first else (std_aln + list(extra_aln_tags))
    out += tag_section([t for t, _ in file_tags]) + tag_section([t for t, _ in read_tags]) + tag_section([t for t, _ in aln_tags])
    for t, v in file_tags:
        out += tag_value(t, v)
    for bc, recs in cells:
        body = b""
        for rec in recs:
            umi, refs = rec[0], rec[1]
            fw = rec[2] if len(rec) > 2 else [True] * len(refs)
            body += struct.pack("<I", len(refs))
            for t, v in read_tags:
                body += struct.pack(FMT[t[1]], bc if t[0] == "b" else umi if t[0] == "u" else v)
            for r, f in zip(refs, fw):
                for t, v in aln_tags:
                    body += struct.pack(FMT[t[1]], (r | (0x80000000 if f else 0)) if t[0] == "compressed_ori_refid" else v)
        out += struct.pack("<II", len(body) + 8, len(recs)) + body
    with open(path, "wb") as f:
        f.write(out)
    return len(out)


def fnv(values):
    h = 0xCBF29CE484222325
    for v in values:
        h = ((h ^ v) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h
