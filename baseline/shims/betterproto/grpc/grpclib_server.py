"""``ServiceBase`` of betterproto's grpclib server glue (shim, reference arm only)."""


class ServiceBase:
    async def _call_rpc_handler_server_stream(self, handler, stream, request) -> None:
        async for response in handler(request):
            await stream.send_message(response)
