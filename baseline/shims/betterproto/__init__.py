"""Minimal stand-in for ``betterproto`` 2.0.0b7 — ONLY for the benchmark's reference arm.

The unmodified reference (``baseline/_ref/pytensor_federated``) imports ``betterproto`` and
``grpclib``; neither wheel exists in this offline image.  This shim implements the small API
surface the reference touches (dataclass message fields, ``bytes(msg)``, ``Msg().parse``,
``ServiceStub``) with the standard proto3 wire format, so the reference's own classes, codec and
client/server logic run as written.  It is not part of the product package.
"""
from __future__ import annotations

import dataclasses
import struct
import sys
import typing
from typing import Any, Dict, Optional

_PLACEHOLDER = object()


def _field(number: int, kind: str):
    return dataclasses.field(default=_PLACEHOLDER, metadata={"pb": (number, kind)})


def bytes_field(number: int, **_): return _field(number, "bytes")
def string_field(number: int, **_): return _field(number, "string")
def int64_field(number: int, **_): return _field(number, "int64")
def int32_field(number: int, **_): return _field(number, "int32")
def float_field(number: int, **_): return _field(number, "float")
def message_field(number: int, **_): return _field(number, "message")


def _varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while v > 0x7F:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _read_varint(buf, pos):
    res = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        res |= (b & 0x7F) << shift
        if not b & 0x80:
            return res, pos
        shift += 7


def _signed(v: int, bits: int) -> int:
    v &= (1 << bits) - 1
    return v - (1 << bits) if v >> (bits - 1) else v


class Message:
    """Base of generated message dataclasses."""

    def __post_init__(self) -> None:
        hints = self._hints()
        for f in dataclasses.fields(self):
            if getattr(self, f.name) is _PLACEHOLDER:
                setattr(self, f.name, self._default(f, hints))

    @classmethod
    def _hints(cls) -> Dict[str, Any]:
        cached = cls.__dict__.get("_hint_cache")
        if cached is None:
            mod = sys.modules[cls.__module__]
            cached = typing.get_type_hints(cls, vars(mod))
            cls._hint_cache = cached
        return cached

    @staticmethod
    def _is_repeated(hint) -> bool:
        return typing.get_origin(hint) in (list, typing.List)

    def _default(self, f, hints):
        _, kind = f.metadata["pb"]
        if self._is_repeated(hints[f.name]):
            return []
        return {"bytes": b"", "string": "", "int64": 0, "int32": 0, "float": 0.0, "message": None}[kind]

    def __bytes__(self) -> bytes:
        hints = self._hints()
        out = bytearray()
        for f in sorted(dataclasses.fields(self), key=lambda f: f.metadata["pb"][0]):
            number, kind = f.metadata["pb"]
            value = getattr(self, f.name)
            repeated = self._is_repeated(hints[f.name])
            if kind in ("int64", "int32"):
                if repeated:
                    if value:
                        body = b"".join(_varint(int(v)) for v in value)
                        out += _varint(number << 3 | 2) + _varint(len(body)) + body
                elif value:
                    out += _varint(number << 3 | 0) + _varint(int(value))
            elif kind == "float":
                if value:
                    out += _varint(number << 3 | 5) + struct.pack("<f", value)
            else:
                items = value if repeated else ([] if value in (None, b"", "") else [value])
                for item in items:
                    body = item.encode("utf-8") if kind == "string" else bytes(item)
                    out += _varint(number << 3 | 2) + _varint(len(body)) + body
        return bytes(out)

    SerializeToString = __bytes__

    def parse(self, data):
        hints = self._hints()
        by_number = {f.metadata["pb"][0]: f for f in dataclasses.fields(self)}
        for f in by_number.values():
            setattr(self, f.name, self._default(f, hints))
        buf = memoryview(bytes(data))
        pos, end = 0, len(buf)
        while pos < end:
            key, pos = _read_varint(buf, pos)
            number, wt = key >> 3, key & 7
            if wt == 0:
                raw, pos = _read_varint(buf, pos)
            elif wt == 2:
                n, pos = _read_varint(buf, pos)
                raw = buf[pos : pos + n]
                pos += n
            elif wt == 5:
                raw = bytes(buf[pos : pos + 4])
                pos += 4
            elif wt == 1:
                raw = bytes(buf[pos : pos + 8])
                pos += 8
            else:
                raise ValueError(f"unsupported wire type {wt}")
            f = by_number.get(number)
            if f is None:
                continue
            _, kind = f.metadata["pb"]
            hint = hints[f.name]
            repeated = self._is_repeated(hint)
            if kind in ("int64", "int32"):
                bits = 64 if kind == "int64" else 32
                if wt == 2:
                    p, vals = 0, []
                    while p < len(raw):
                        v, p = _read_varint(raw, p)
                        vals.append(_signed(v, bits))
                    getattr(self, f.name).extend(vals) if repeated else setattr(self, f.name, vals[-1])
                else:
                    v = _signed(raw, bits)
                    getattr(self, f.name).append(v) if repeated else setattr(self, f.name, v)
            elif kind == "float":
                setattr(self, f.name, struct.unpack("<f", raw)[0])
            elif kind == "bytes":
                setattr(self, f.name, bytes(raw))
            elif kind == "string":
                setattr(self, f.name, bytes(raw).decode("utf-8"))
            else:
                cls = typing.get_args(hint)[0] if repeated else hint
                msg = cls().parse(raw)
                getattr(self, f.name).append(msg) if repeated else setattr(self, f.name, msg)
        return self

    @classmethod
    def FromString(cls, data):
        return cls().parse(data)


class ServiceStub:
    """Client stub base: unary and streaming calls over a (shimmed) grpclib channel."""

    def __init__(self, channel, *, timeout: Optional[float] = None, deadline=None, metadata=None) -> None:
        self.channel = channel
        self.timeout = timeout
        self.deadline = deadline
        self.metadata = metadata

    async def _unary_unary(self, route, request, response_type, *, timeout=None, deadline=None, metadata=None):
        return await self.channel._unary_unary(route, request, response_type, timeout if timeout is not None else self.timeout)

    async def _stream_stream(self, route, request_iterator, request_type, response_type, *, timeout=None, deadline=None, metadata=None):
        from grpclib.const import Cardinality

        async with self.channel.request(route, Cardinality.STREAM_STREAM, request_type, response_type, timeout=timeout) as stream:
            await stream.send_request()
            if hasattr(request_iterator, "__aiter__"):
                async for item in request_iterator:
                    await stream.send_message(item)
                    yield await stream.recv_message()
            else:
                for item in request_iterator:
                    await stream.send_message(item)
                    yield await stream.recv_message()
            await stream.end()
