"""``Server`` with grpclib's lifecycle, serving betterproto-style handlers through ``grpc.aio``."""
from __future__ import annotations

import grpc
import grpc.aio

_OPTIONS = (("grpc.max_send_message_length", 1 << 30), ("grpc.max_receive_message_length", 1 << 30))


class Stream:
    """What a grpclib handler receives: ``recv_message`` / ``send_message`` / async iteration."""

    def __init__(self, context) -> None:
        self._context = context

    async def recv_message(self):
        msg = await self._context.read()
        return None if msg is grpc.aio.EOF else msg

    async def send_message(self, message) -> None:
        await self._context.write(message)

    def __aiter__(self):
        return self._iterate()

    async def _iterate(self):
        while True:
            msg = await self._context.read()
            if msg is grpc.aio.EOF:
                return
            yield msg


class Server:
    def __init__(self, handlers) -> None:
        self._handlers = list(handlers)
        self._server = None

    async def start(self, host=None, port=None) -> None:
        self._server = grpc.aio.server(options=_OPTIONS)
        by_service = {}
        for service in self._handlers:
            for route, h in service.__mapping__().items():
                _, svc, method = route.split("/")

                def make(func):
                    async def call(request_iterator, context):
                        await func(Stream(context))

                    return call

                by_service.setdefault(svc, {})[method] = grpc.stream_stream_rpc_method_handler(
                    make(h.func),
                    request_deserializer=lambda b, t=h.request_type: t().parse(b),
                    response_serializer=bytes,
                )
        self._server.add_generic_rpc_handlers(
            tuple(grpc.method_handlers_generic_handler(svc, methods) for svc, methods in by_service.items())
        )
        bound = self._server.add_insecure_port(f"{host}:{port}")
        if bound == 0:
            raise OSError(f"cannot bind {host}:{port}")
        await self._server.start()

    async def wait_closed(self) -> None:
        await self._server.wait_for_termination()

    def close(self) -> None:
        pass
