"""A small, independent encoder of the bincode 1.x wire format (the crate the reference serializes its replay records with:
bincode 1.3.3, Cargo.lock:191-192), written from the format's specification — NOT from oracle/replay.hpp or the CUDA packer:

    `bincode::serialize(&T)` = DefaultOptions().with_fixint_encoding().allow_trailing_bytes(), little endian:
      * integers: fixed width, little endian two's complement;  f32 / f64: IEEE-754 bits, little endian;  bool: one byte
      * a struct / tuple: its fields in declaration order, nothing in between (no names, no tags, no padding)
      * a sequence (serde `serialize_seq`: Vec, slices, heapless::Vec): u64 length, then the elements
      * Option: one tag byte (0 / 1) then the value;  enum variant: u32 index then the fields;  String: u64 byte length + UTF-8

The serde data model of a value is given as nested Python tuples: ("struct", [fields...]), ("seq", elem_type, [values...]),
("i64", v), ("u64", v), ("u32", v), ("f32", v), ("option", None | typed value).  heapless 0.8.0 `Vec<T, N>` implements
Serialize through `serialize_seq(Some(len))` (heapless/src/ser.rs), i.e. it is a "seq".
"""
import struct

_SCALARS = {"i8": "<b", "u8": "<B", "i16": "<h", "u16": "<H", "i32": "<i", "u32": "<I", "i64": "<q", "u64": "<Q", "f32": "<f", "f64": "<d"}


def encode(v):
    kind = v[0]
    if kind in _SCALARS:
        return struct.pack(_SCALARS[kind], v[1])
    if kind == "bool":
        return b"\x01" if v[1] else b"\x00"
    if kind == "struct":
        return b"".join(encode(f) for f in v[1])
    if kind == "seq":
        elem, values = v[1], v[2]
        return struct.pack("<Q", len(values)) + b"".join(encode((elem, x)) for x in values)
    if kind == "option":
        return b"\x00" if v[1] is None else b"\x01" + encode(v[1])
    if kind == "string":
        raw = v[1].encode("utf-8")
        return struct.pack("<Q", len(raw)) + raw
    raise ValueError(kind)


def db_record(state, value, policy):
    """DBRecord { state: heapless::Vec<i64, 311>, value: heapless::Vec<f32, 4>, policy: heapless::Vec<f32, 39> }
    (rs-doko-alpha-zero/src/alpha_zero/net/experience_replay_buffer3.rs:11-20) as the bytes `bincode::serialize(&record)` produces (:109)."""
    return encode(("struct", [("seq", "i64", [int(x) for x in state]), ("seq", "f32", [float(x) for x in value]), ("seq", "f32", [float(x) for x in policy])]))
