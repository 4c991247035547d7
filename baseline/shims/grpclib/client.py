"""``Channel`` / ``Stream`` with grpclib's method names, implemented with ``grpc.aio``."""
from __future__ import annotations

import asyncio
from typing import Generic, Optional

import grpc
import grpc.aio

from .exceptions import StreamTerminatedError
from .stream import _RecvType, _SendType

_OPTIONS = (("grpc.max_send_message_length", 1 << 30), ("grpc.max_receive_message_length", 1 << 30))


class Stream(Generic[_SendType, _RecvType]):
    def __init__(self, channel: "Channel", route: str, request_type, reply_type, timeout) -> None:
        self._channel = channel
        self._route = route
        self._reply_type = reply_type
        self._timeout = timeout
        self._call = None

    async def __aenter__(self):
        multicallable = self._channel._aio.stream_stream(
            self._route, request_serializer=bytes, response_deserializer=lambda b: self._reply_type().parse(b)
        )
        self._call = multicallable(timeout=self._timeout)
        return self

    async def __aexit__(self, *exc) -> None:
        if self._call is not None:
            self._call.cancel()

    async def send_request(self) -> None:
        try:
            await asyncio.wait_for(self._call.wait_for_connection(), 10)
        except grpc.aio.AioRpcError as ex:
            if ex.code() == grpc.StatusCode.UNAVAILABLE:
                raise ConnectionRefusedError(str(ex.details())) from ex
            raise StreamTerminatedError(str(ex)) from ex

    async def send_message(self, message, *, end: bool = False) -> None:
        try:
            await self._call.write(message)
            if end:
                await self._call.done_writing()
        except (grpc.aio.AioRpcError, asyncio.InvalidStateError) as ex:
            raise StreamTerminatedError(str(ex)) from ex

    async def recv_message(self):
        try:
            response = await self._call.read()
        except grpc.aio.AioRpcError as ex:
            raise StreamTerminatedError(str(ex)) from ex
        if response is grpc.aio.EOF:
            raise StreamTerminatedError("Connection lost")
        return response

    async def end(self) -> None:
        try:
            await self._call.done_writing()
        except Exception:
            pass

    async def cancel(self) -> None:
        self._call.cancel()


class Channel:
    def __init__(self, host: Optional[str] = None, port: Optional[int] = None, **_) -> None:
        self._host = host
        self._port = port
        self._aio = grpc.aio.insecure_channel(f"{host}:{port}", options=_OPTIONS)
        self._protocol = object()

    def request(self, name, cardinality, request_type, reply_type, *, timeout=None, deadline=None, metadata=None):
        return Stream(self, name, request_type, reply_type, timeout)

    async def _unary_unary(self, route, request, response_type, timeout):
        call = self._aio.unary_unary(route, request_serializer=bytes, response_deserializer=lambda b: response_type().parse(b))
        try:
            return await call(request, timeout=timeout)
        except grpc.aio.AioRpcError as ex:
            if ex.code() == grpc.StatusCode.UNAVAILABLE:
                raise ConnectionRefusedError(str(ex.details())) from ex
            if ex.code() == grpc.StatusCode.DEADLINE_EXCEEDED:
                raise asyncio.TimeoutError() from ex
            raise StreamTerminatedError(str(ex)) from ex

    def close(self) -> None:
        if self._protocol is None:
            return
        self._protocol = None
        coro = self._aio.close(None)
        try:
            loop = asyncio.get_running_loop()
            loop.create_task(coro)
        except RuntimeError:
            try:
                asyncio.get_event_loop().run_until_complete(coro)
            except Exception:
                coro.close()
