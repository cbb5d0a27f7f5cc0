"""Wire schema of the detector stage, rebuilt from its field table.  TEST INFRASTRUCTURE.

The reference ships the schema only as a serialized descriptor inside
/root/reference/container/fluentout/schemas_pb.rb:8 (proto3 file ``schemas.proto``:
Schema, LogSchema, ParserSchema, DetectorSchema, OutputSchema); the Python wrappers
(``detectmatelibrary.schemas``) are not vendored.  This module restates the three
messages on the hot path as a FileDescriptorProto built field by field (SURVEY.md
Appendix A) and lets protobuf's upb runtime do the encoding -- an independent codec
from the product's hand-written one in ``detectmateservice_b200/wire.py``.
``tests/test_wire_schema.py`` pins it against the 201-byte ParserSchema fixture and, when
/root/reference is present, against the reference descriptor itself.
"""
from __future__ import annotations

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto
_STR, _I32, _FLT, _MSG = _F.TYPE_STRING, _F.TYPE_INT32, _F.TYPE_FLOAT, _F.TYPE_MESSAGE
_OPT, _REP = _F.LABEL_OPTIONAL, _F.LABEL_REPEATED

# (name, number, type, label, proto3_optional)
_LOG = [("__version__", 1, _STR, _OPT), ("logID", 2, _STR, _OPT), ("log", 3, _STR, _OPT),
        ("logSource", 4, _STR, _OPT), ("hostname", 5, _STR, _OPT)]
_PARSER = [("__version__", 1, _STR, _OPT), ("parserType", 2, _STR, _OPT), ("parserID", 3, _STR, _OPT),
           ("EventID", 4, _I32, _OPT), ("template", 5, _STR, _OPT), ("variables", 6, _STR, _REP),
           ("parsedLogID", 7, _STR, _OPT), ("logID", 8, _STR, _OPT), ("log", 9, _STR, _OPT),
           ("logFormatVariables", 10, "map", _REP), ("receivedTimestamp", 11, _I32, _OPT),
           ("parsedTimestamp", 12, _I32, _OPT)]
_DETECTOR = [("__version__", 1, _STR, _OPT), ("detectorID", 2, _STR, _OPT), ("detectorType", 3, _STR, _OPT),
             ("alertID", 4, _STR, _OPT), ("detectionTimestamp", 5, _I32, _OPT), ("logIDs", 6, _STR, _REP),
             ("score", 8, _FLT, _OPT), ("extractedTimestamps", 9, _I32, _REP), ("description", 10, _STR, _OPT),
             ("receivedTimestamp", 11, _I32, _OPT), ("alertsObtain", 12, "map", _REP)]


def _camel(name: str) -> str:
    return name[0].upper() + name[1:] + "Entry"


def _add_message(fd: descriptor_pb2.FileDescriptorProto, name: str, fields) -> None:
    m = fd.message_type.add()
    m.name = name
    n_oneof = 0
    for fname, num, ftype, label in fields:
        f = m.field.add()
        f.name, f.number, f.label = fname, num, label
        if ftype == "map":
            entry = m.nested_type.add()
            entry.name = _camel(fname)
            entry.options.map_entry = True
            for en, enum_ in (("key", 1), ("value", 2)):
                ef = entry.field.add()
                ef.name, ef.number, ef.type, ef.label = en, enum_, _STR, _OPT
            f.type = _MSG
            f.type_name = f".{name}.{entry.name}"
        else:
            f.type = ftype
            if label == _OPT:
                # proto3 `optional`: explicit presence via a synthetic oneof
                f.proto3_optional = True
                f.oneof_index = n_oneof
                n_oneof += 1
    for fname, num, ftype, label in fields:
        if ftype != "map" and label == _OPT:
            m.oneof_decl.add().name = "_" + fname


def build_file_descriptor() -> descriptor_pb2.FileDescriptorProto:
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "dm_oracle_schemas.proto"
    fd.syntax = "proto3"
    _add_message(fd, "LogSchema", _LOG)
    _add_message(fd, "ParserSchema", _PARSER)
    _add_message(fd, "DetectorSchema", _DETECTOR)
    return fd


_POOL = descriptor_pool.DescriptorPool()
_POOL.Add(build_file_descriptor())
LogSchema = message_factory.GetMessageClass(_POOL.FindMessageTypeByName("LogSchema"))
ParserSchema = message_factory.GetMessageClass(_POOL.FindMessageTypeByName("ParserSchema"))
DetectorSchema = message_factory.GetMessageClass(_POOL.FindMessageTypeByName("DetectorSchema"))


def parser_schema_from_dict(d: dict) -> "ParserSchema":
    """Same construction the reference fixtures use: ``ParserSchema(dict)``
    (tests/library_integration/library_integration_base_fixtures.py:80-83)."""
    m = ParserSchema()
    m.__setattr__("__version__", d.get("__version__", "1.0.0"))
    for k, v in d.items():
        if k == "__version__":
            continue
        if k == "variables":
            m.variables.extend(v)
        elif k == "logFormatVariables":
            for kk, vv in v.items():
                m.logFormatVariables[kk] = vv
        else:
            setattr(m, k, v)
    return m
