import collections
import enum


class Cardinality(enum.Enum):
    UNARY_UNARY = "unary_unary"
    UNARY_STREAM = "unary_stream"
    STREAM_UNARY = "stream_unary"
    STREAM_STREAM = "stream_stream"


class Status(enum.Enum):
    OK = 0
    CANCELLED = 1
    UNKNOWN = 2
    DEADLINE_EXCEEDED = 4
    UNIMPLEMENTED = 12
    UNAVAILABLE = 14


Handler = collections.namedtuple("Handler", "func, cardinality, request_type, reply_type")
