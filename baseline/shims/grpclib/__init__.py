"""Minimal stand-in for ``grpclib`` on top of ``grpc.aio`` — ONLY for the benchmark's reference
arm (see ``baseline/shims/betterproto/__init__.py`` for the rationale)."""
from . import client, const, exceptions, metadata, server, stream  # noqa: F401
from .exceptions import GRPCError  # noqa: F401
