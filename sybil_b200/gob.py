"""Go `encoding/gob` reader / writer — the envelope of every file either side of the hot path.

sybil stores each column of a block (`int_<col>.db`, `str_<col>.db`), the block's `info.db` and the
results it exchanges between nodes as gob streams (`src/lib/file_decoder.go:27-81`,
`column_store_io.go:132-134`, `printer.go:272-297`).  The reference host is Go and calls
`gob.Decode`; this module is what a host WITHOUT Go uses to read real sybil files and hand the flat
arrays to the C ABI, plus a writer for the same structs (N2) so blocks generated
here can be laid out the way sybil lays them out.

The decoder is generic: it reads the type definitions that precede the values in the stream and binds
struct fields by the transmitted names, so it decodes any gob stream made of bools, ints, uints,
floats, strings, byte slices, slices, arrays, maps, structs and interface values.  It is pinned to Go's
own output by the reference's golden files (`testdata/TestDecodeGoldenFiles/*.golden.gob`, fixtures of
`src/lib/decoding_test.go:20-74`): `tests/test_gob.py` decodes them and compares with the JSON
rendering of the same values that the reference keeps next to them.

Wire format (Go documentation of encoding/gob):
  stream   = { uvarint byte-count, message }
  message  = int type-id ; type-id < 0: definition of type -id (a wireType struct)
                           type-id > 0: a value of that type (non-struct: preceded by a 0 byte)
  uint     = one byte < 128, or a byte holding 256-n followed by n big-endian bytes
  int      = uint u with bit 0 as complement flag: u & 1 ? ~(u >> 1) : u >> 1
  float    = IEEE-754 bits, byte-reversed, as uint
  string   = uint length, bytes        slice/array = uint count, elements
  map      = uint count, key/value pairs
  struct   = { uint field-number delta, value } 0      (zero-valued fields are not sent)
  interface= string name ; empty = nil ; else { type definitions } int type-id, uint byte-count, value
"""
import struct as _struct

# bootstrap type ids (encoding/gob/type.go)
T_BOOL, T_INT, T_UINT, T_FLOAT, T_BYTES, T_STRING, T_COMPLEX, T_INTERFACE = 1, 2, 3, 4, 5, 6, 7, 8
T_WIRETYPE, T_ARRAYTYPE, T_COMMONTYPE, T_SLICETYPE, T_STRUCTTYPE, T_FIELDTYPE, T_FIELDTYPE_SLICE, T_MAPTYPE = 16, 17, 18, 19, 20, 21, 22, 23
T_GOBENC, T_BINMARSH, T_TEXTMARSH = 24, 25, 26
FIRST_USER_ID = 65


class GobError(ValueError):
    pass


# the self-describing part: how a wireType itself is laid out (fields in declaration order)
_BOOT = {
    T_WIRETYPE: ("struct", [("ArrayT", T_ARRAYTYPE), ("SliceT", T_SLICETYPE), ("StructT", T_STRUCTTYPE), ("MapT", T_MAPTYPE),
                            ("GobEncoderT", T_GOBENC), ("BinaryMarshalerT", T_BINMARSH), ("TextMarshalerT", T_TEXTMARSH)]),
    T_ARRAYTYPE: ("struct", [("CommonType", T_COMMONTYPE), ("Elem", T_INT), ("Len", T_INT)]),
    T_COMMONTYPE: ("struct", [("Name", T_STRING), ("Id", T_INT)]),
    T_SLICETYPE: ("struct", [("CommonType", T_COMMONTYPE), ("Elem", T_INT)]),
    T_STRUCTTYPE: ("struct", [("CommonType", T_COMMONTYPE), ("Field", T_FIELDTYPE_SLICE)]),
    T_FIELDTYPE: ("struct", [("Name", T_STRING), ("Id", T_INT)]),
    T_FIELDTYPE_SLICE: ("slice", T_FIELDTYPE),
    T_MAPTYPE: ("struct", [("CommonType", T_COMMONTYPE), ("Key", T_INT), ("Elem", T_INT)]),
    T_GOBENC: ("struct", [("CommonType", T_COMMONTYPE)]),
    T_BINMARSH: ("struct", [("CommonType", T_COMMONTYPE)]),
    T_TEXTMARSH: ("struct", [("CommonType", T_COMMONTYPE)]),
}


class Decoder:
    """Decoder(data).decode() -> next top-level value, or raises EOFError at the end of the stream."""

    def __init__(self, data):
        self.b = memoryview(bytes(data))
        self.p = 0
        self.types = dict(_BOOT)

    # ---- primitives ------------------------------------------------------------------------
    def _uint(self):
        b = self.b
        if self.p >= len(b):
            raise EOFError
        c = b[self.p]
        self.p += 1
        if c < 128:
            return c
        n = 256 - c
        if n > 8 or self.p + n > len(b):
            raise GobError("bad uint at %d" % (self.p - 1))
        v = int.from_bytes(b[self.p:self.p + n], "big")
        self.p += n
        return v

    def _int(self):
        u = self._uint()
        return ~(u >> 1) if u & 1 else u >> 1

    def _float(self):
        u = self._uint()
        return _struct.unpack("<d", u.to_bytes(8, "big"))[0]  # byte-reversed bits

    def _bytes(self):
        n = self._uint()
        if self.p + n > len(self.b):
            raise GobError("bad length at %d" % self.p)
        v = bytes(self.b[self.p:self.p + n])
        self.p += n
        return v

    # ---- values ----------------------------------------------------------------------------
    def _value(self, tid):
        if tid == T_BOOL:
            return self._uint() != 0
        if tid == T_INT:
            return self._int()
        if tid == T_UINT:
            return self._uint()
        if tid == T_FLOAT:
            return self._float()
        if tid == T_BYTES:
            return self._bytes()
        if tid == T_STRING:
            return self._bytes().decode("utf-8", "surrogateescape")
        if tid == T_COMPLEX:
            return complex(self._float(), self._float())
        if tid == T_INTERFACE:
            return self._interface()
        t = self.types.get(tid)
        if t is None:
            raise GobError("value of undefined type %d" % tid)
        kind = t[0]
        if kind == "struct":
            out, fields, i = {}, t[1], -1
            while True:
                d = self._uint()
                if d == 0:
                    return out
                i += d
                if i >= len(fields):
                    raise GobError("field %d out of range for %r" % (i, fields))
                out[fields[i][0]] = self._value(fields[i][1])
        if kind == "slice" or kind == "array":
            n = self._uint()
            et = t[1]
            if et == T_UINT:
                return [self._uint() for _ in range(n)]
            if et == T_INT:
                return [self._int() for _ in range(n)]
            return [self._value(et) for _ in range(n)]
        if kind == "map":
            n = self._uint()
            out = {}
            for _ in range(n):
                k = self._value(t[1])
                out[k] = self._value(t[2])
            return out
        if kind == "opaque":  # GobEncoder / BinaryMarshaler / TextMarshaler payloads
            return self._bytes()
        raise GobError("unknown kind %r" % (kind,))

    def _define(self, tid):
        w = self._value(T_WIRETYPE)
        if "StructT" in w:
            s = w["StructT"]
            self.types[tid] = ("struct", [(f.get("Name", ""), f.get("Id", 0)) for f in s.get("Field", [])], s.get("CommonType", {}).get("Name", ""))
        elif "SliceT" in w:
            self.types[tid] = ("slice", w["SliceT"].get("Elem", 0))
        elif "ArrayT" in w:
            self.types[tid] = ("array", w["ArrayT"].get("Elem", 0))
        elif "MapT" in w:
            self.types[tid] = ("map", w["MapT"].get("Key", 0), w["MapT"].get("Elem", 0))
        elif "GobEncoderT" in w or "BinaryMarshalerT" in w or "TextMarshalerT" in w:
            self.types[tid] = ("opaque",)
        else:
            raise GobError("empty wireType for %d" % tid)

    def _is_struct(self, tid):
        t = self.types.get(tid)
        return t is not None and t[0] == "struct"

    def _interface(self):
        name = self._bytes().decode("utf-8", "surrogateescape")
        if not name:
            return None
        # type definitions the receiver may not have yet, each followed by the length of the next chunk
        while True:
            tid = self._int()
            if tid >= 0:
                break
            self._define(-tid)
            self._uint()
        self._uint()  # byte count of the value
        if not self._is_struct(tid):
            if self._uint() != 0:
                raise GobError("interface value: expected the 0 marker of a non-struct")
        v = self._value(tid)
        if isinstance(v, dict):
            v.setdefault("__type__", name)
            return v
        return {"__type__": name, "value": v}

    def decode(self):
        while True:
            self._uint()  # message length: with interface values it covers only the first fragment — parse linearly
            tid = self._int()
            if tid < 0:
                self._define(-tid)
                continue
            if not self._is_struct(tid):
                if self._uint() != 0:
                    raise GobError("top-level non-struct value: expected the 0 marker")
            return self._value(tid)


def decode(data):
    """The first top-level value of a gob stream."""
    return Decoder(data).decode()


# ---------------------------------------------------------------------------------------------
# writer: values described by explicit Go-like type descriptors
#   "bool" "int" "uint" "float" "string" "bytes"
#   ("slice", elem)  ("map", key, elem)  ("struct", GoTypeName, [(FieldName, type), ...])
#   "interface": the value is (registered name, concrete type descriptor, concrete value) or None.
#       The concrete types' definitions are sent ahead of the value message (Encoder.encode collects
#       them from the value), so an interface value is name, type id, byte count, value — the form
#       Go's decodeTypeSequence accepts when it has nothing to define in line.
# ---------------------------------------------------------------------------------------------
_BASIC = {"bool": T_BOOL, "int": T_INT, "uint": T_UINT, "float": T_FLOAT, "bytes": T_BYTES, "string": T_STRING,
          "interface": T_INTERFACE}


def _enc_uint(v):
    if v < 0:
        raise GobError("negative uint")
    if v < 128:
        return bytes([v])
    n = (v.bit_length() + 7) // 8
    return bytes([256 - n]) + v.to_bytes(n, "big")


def _enc_int(v):
    return _enc_uint((~v << 1) | 1 if v < 0 else v << 1)


class Encoder:
    """Encoder().encode(value, type) -> bytes of one gob stream holding that value."""

    def __init__(self):
        self.ids = {}      # descriptor key -> type id
        self.defs = []     # (id, wireType bytes) in the order Go would send them
        self.next_id = FIRST_USER_ID

    def _key(self, t):
        if isinstance(t, str):
            return t
        if t[0] == "struct":
            return ("struct", t[1])
        return (t[0],) + tuple(self._key(x) for x in t[1:])

    def _type_id(self, t):
        if isinstance(t, str):
            return _BASIC[t]
        k = self._key(t)
        if k in self.ids:
            return self.ids[k]
        tid = self.next_id
        self.next_id += 1
        self.ids[k] = tid
        # Go registers a composite type before its element types get ids, but SENDS the elements'
        # definitions first; ids are assigned on first sight, depth first
        slot = len(self.defs)
        self.defs.append(None)
        if t[0] == "slice":
            e = self._type_id(t[1])
            body = _enc_uint(2) + self._common("[]" + self._name(t[1]), tid) + _enc_uint(1) + _enc_int(e) + b"\x00" + b"\x00"
        elif t[0] == "map":
            kk, e = self._type_id(t[1]), self._type_id(t[2])
            body = (_enc_uint(4) + self._common("map[%s]%s" % (self._name(t[1]), self._name(t[2])), tid) + _enc_uint(1) + _enc_int(kk) +
                    _enc_uint(1) + _enc_int(e) + b"\x00" + b"\x00")
        elif t[0] == "struct":
            fields = b""
            for fname, ft in t[2]:
                fid = self._type_id(ft)
                fields += _enc_uint(1) + _enc_uint(len(fname.encode())) + fname.encode() + _enc_uint(1) + _enc_int(fid) + b"\x00"
            body = (_enc_uint(3) + self._common(t[1], tid) + (_enc_uint(1) + _enc_uint(len(t[2])) + fields if t[2] else b"") + b"\x00" + b"\x00")
        else:
            raise GobError("unknown descriptor %r" % (t,))
        self.defs[slot] = (tid, body)
        return tid

    def _name(self, t):
        if isinstance(t, str):
            return {"bytes": "[]uint8", "float": "float64", "interface": "interface {}"}.get(t, t)
        if t[0] == "struct":
            return t[1]
        if t[0] == "slice":
            return "[]" + self._name(t[1])
        return "map[%s]%s" % (self._name(t[1]), self._name(t[2]))

    @staticmethod
    def _common(name, tid):
        # field 1 of ArrayT/SliceT/StructT/MapT: CommonType{Name, Id}
        nb = name.encode()
        return _enc_uint(1) + _enc_uint(1) + _enc_uint(len(nb)) + nb + _enc_uint(1) + _enc_int(tid) + b"\x00"

    def _zero(self, v, t):
        if t == "interface":
            return v is None
        if isinstance(t, str):
            return v in (0, 0.0, False, "", b"", None)
        if t[0] == "struct":
            return False
        return v is None or len(v) == 0

    def _val(self, v, t):
        if t == "bool":
            return _enc_uint(1 if v else 0)
        if t == "int":
            return _enc_int(int(v))
        if t == "uint":
            return _enc_uint(int(v))
        if t == "float":
            return _enc_uint(int.from_bytes(_struct.pack("<d", float(v)), "big"))
        if t == "string":
            b = v.encode("utf-8", "surrogateescape")
            return _enc_uint(len(b)) + b
        if t == "bytes":
            return _enc_uint(len(v)) + bytes(v)
        if t == "interface":
            if v is None:
                return _enc_uint(0)
            name, ct, cv = v
            nb = name.encode()
            body = self._val(cv, ct)
            if isinstance(ct, str) or ct[0] != "struct":
                body = b"\x00" + body
            return _enc_uint(len(nb)) + nb + _enc_int(self._type_id(ct)) + _enc_uint(len(body)) + body
        if t[0] == "slice":
            et = t[1]
            if et == "uint":
                return _enc_uint(len(v)) + b"".join(map(_enc_uint, map(int, v)))
            if et == "int":
                return _enc_uint(len(v)) + b"".join(map(_enc_int, map(int, v)))
            return _enc_uint(len(v)) + b"".join(self._val(x, et) for x in v)
        if t[0] == "map":
            return _enc_uint(len(v)) + b"".join(self._val(k, t[1]) + self._val(x, t[2]) for k, x in v.items())
        if t[0] == "struct":
            out, last = b"", -1
            for i, (fname, ft) in enumerate(t[2]):
                fv = v.get(fname)
                if fv is None or self._zero(fv, ft):
                    continue
                out += _enc_uint(i - last) + self._val(fv, ft)
                last = i
            return out + b"\x00"
        raise GobError("unknown descriptor %r" % (t,))

    def _register_concrete(self, v, t):
        """Assign ids to (and queue the definitions of) the concrete types inside interface values."""
        if t == "interface":
            if v is not None:
                self._type_id(v[1])
                self._register_concrete(v[2], v[1])
        elif isinstance(t, str) or v is None:
            return
        elif t[0] == "slice":
            if not isinstance(t[1], str) or t[1] == "interface":
                for x in v:
                    self._register_concrete(x, t[1])
        elif t[0] == "map":
            if not isinstance(t[2], str) or t[2] == "interface":
                for x in v.values():
                    self._register_concrete(x, t[2])
        elif t[0] == "struct":
            for fname, ft in t[2]:
                if fname in v:
                    self._register_concrete(v[fname], ft)

    def encode(self, value, t):
        tid = self._type_id(t)
        self._register_concrete(value, t)
        out = b""
        for did, body in self.defs:
            msg = _enc_int(-did) + body
            out += _enc_uint(len(msg)) + msg
        self.defs = []
        body = self._val(value, t)
        msg = _enc_int(tid) + (body if not isinstance(t, str) and t[0] == "struct" else b"\x00" + body)
        return out + _enc_uint(len(msg)) + msg


def encode(value, t):
    return Encoder().encode(value, t)
