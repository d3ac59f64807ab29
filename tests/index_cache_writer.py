"""Writer of kimchi's proving-key cache format ("MINAPK01", format version 3) — test infrastructure restating the reference's
serializer, kimchi/src/cached_prover_index.rs:560-760 (write_preamble, ScalarHeader::write, SectionEntry::write) and the layout
comment at :1-56: fixed preamble, fixed ScalarHeader, section table, then 32-byte aligned payload sections of raw Montgomery limbs.
The reference tree ships no cache file (they are produced at run time), so the ingestion test builds one with this writer."""
import struct

FILE_MAGIC = b"MINAPK01"
FORMAT_VERSION = 3
ARK_FF_VERSION = b"ark-ff-0.5"
ARK_FF_VERSION_MAX_LEN, IDENTIFIER_MAX_LEN, SECTION_ALIGNMENT, PERMUTS, COLUMNS = 32, 512, 32, 7, 15
PREAMBLE_SIZE = 8 + 4 + 4 + ARK_FF_VERSION_MAX_LEN + 4 + IDENTIFIER_MAX_LEN + 4
SCALAR_HEADER_SIZE = 4 + 4 + 8 + 8 + 1 + 7 + 8 + 4 + 4 + 4 + 1 + 3 + 32 + 32 * PERMUTS + 32
SECTION_ENTRY_SIZE = 4 + 8 + 8 + 4 + 4


def align_up(n):
    return (n + SECTION_ALIGNMENT - 1) & ~(SECTION_ALIGNMENT - 1)


def write_cache(identifier: str, header: dict, sections: list, version: int = FORMAT_VERSION, magic: bytes = FILE_MAGIC) -> bytes:
    """sections: [(tag, payload bytes, elem_domain_size)]; header: the ScalarHeader fields (limb lists for endo / shift / digest)"""
    ident = identifier.encode()
    assert len(ident) <= IDENTIFIER_MAX_LEN
    out = bytearray()
    out += magic + struct.pack("<II", version, 0) + ARK_FF_VERSION.ljust(ARK_FF_VERSION_MAX_LEN, b"\0")
    out += struct.pack("<I", len(ident)) + ident.ljust(IDENTIFIER_MAX_LEN, b"\0") + struct.pack("<I", len(sections))
    assert len(out) == PREAMBLE_SIZE
    out += struct.pack("<IIQQB7xQIIIB3x", header["public"], header["prev_challenges"], header["zk_rows"], header["max_poly_size"],
                       int(header.get("disable_gates_checks", False)), header["domain_d1_size"], header.get("feature_flags", 0),
                       header.get("optional_selectors_present", 0), header.get("lookup_selectors_present", 0),
                       int(header.get("has_verifier_index_digest", False)))
    out += struct.pack("<4Q", *header["endo_limbs"])
    for row in header["shift_limbs"]:
        out += struct.pack("<4Q", *row)
    out += struct.pack("<4Q", *header.get("verifier_index_digest_limbs", [0, 0, 0, 0]))
    assert len(out) == PREAMBLE_SIZE + SCALAR_HEADER_SIZE
    table_off = len(out)
    off = align_up(table_off + SECTION_ENTRY_SIZE * len(sections))
    entries, body = bytearray(), bytearray()
    for tag, payload, dom in sections:
        entries += struct.pack("<IQQII", tag, off, len(payload), dom, 0)
        body += bytes(payload)
        pad = align_up(len(payload)) - len(payload)
        body += bytes(pad)
        off += len(payload) + pad
    out += entries
    out += bytes(align_up(len(out)) - len(out))
    out += body
    return bytes(out)
