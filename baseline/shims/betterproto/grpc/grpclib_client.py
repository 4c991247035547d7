from typing import Any

MetadataLike = Any
