from typing import TypeVar

_SendType = TypeVar("_SendType")
_RecvType = TypeVar("_RecvType")
