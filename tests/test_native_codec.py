"""The C++ message codec must be byte-identical to the Python codec (which is pinned to the schema)."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st
from hypothesis.extra import numpy as hnp

from pytensor_federated_b200.npproto import native_codec
from pytensor_federated_b200.npproto.utils import ndarray_from_numpy, ndarray_to_numpy
from pytensor_federated_b200.rpc import InputArrays, OutputArrays

pytestmark = pytest.mark.skipif(not native_codec.available(), reason="libb200fed.so not built")

CASES = [
    [np.arange(5), np.random.default_rng(0).uniform(size=(2, 3)), np.array(5)],
    [np.array(["hello", "world"]), np.datetime64("2022-06"), np.zeros((0, 4), dtype=np.float32)],
    [np.arange(24, dtype=np.int16).reshape(2, 3, 4)[:, ::2], np.asfortranarray(np.arange(6.0).reshape(2, 3))],
    [],
]


@pytest.mark.parametrize("arrays", CASES, ids=["numeric", "strings-dates-empty", "non-contiguous", "no-items"])
def test_native_and_python_codecs_agree(arrays):
    uuid = "3e0f3b6c-0d4d-4d55-8a72-6d4c0c6c3f11"
    python_bytes = bytes(InputArrays(items=[ndarray_from_numpy(a) for a in arrays], uuid=uuid))
    native_bytes = native_codec.encode_arrays(arrays, uuid)
    assert native_bytes == python_bytes
    decoded, got_uuid = native_codec.decode_arrays(python_bytes)
    assert got_uuid == uuid and len(decoded) == len(arrays)
    for got, want in zip(decoded, arrays):
        want = np.asarray(want)
        assert got.dtype == want.dtype and got.shape == want.shape
        np.testing.assert_array_equal(got, want)
        assert not got.flags.writeable
    # and the Python parser reads what the native encoder wrote
    parsed = OutputArrays().parse(native_bytes)
    for item, want in zip(parsed.items, arrays):
        np.testing.assert_array_equal(ndarray_to_numpy(item), want)


def test_native_decoder_rejects_garbage_and_skips_unknown_fields():
    good = native_codec.encode_arrays([np.arange(3)], "u")
    with pytest.raises(ValueError):
        native_codec.decode_arrays(good[:-2] + b"\xff")
    with pytest.raises(ValueError):
        native_codec.decode_arrays(b"\x0a\x7f\x00")
    arrays, uuid = native_codec.decode_arrays(good + b"\x48\x05")  # unknown varint field 9
    assert uuid == "u" and arrays[0].tolist() == [0, 1, 2]


@settings(max_examples=50, deadline=None)
@given(
    st.lists(
        hnp.arrays(
            dtype=st.sampled_from([np.float64, np.float32, np.int64, np.uint8, np.bool_]),
            shape=hnp.array_shapes(min_dims=0, max_dims=4, min_side=0, max_side=4),
        ),
        max_size=4,
    ),
    st.text(alphabet="abcdef0123456789-", max_size=36),
)
def test_fuzz_against_python_codec(arrays, uuid):
    python_bytes = bytes(InputArrays(items=[ndarray_from_numpy(a) for a in arrays], uuid=uuid))
    assert native_codec.encode_arrays(arrays, uuid) == python_bytes
    decoded, got_uuid = native_codec.decode_arrays(python_bytes)
    assert got_uuid == uuid
    for got, want in zip(decoded, arrays):
        np.testing.assert_array_equal(got, want)


def test_short_data_field_is_rejected_not_read_from_neighbouring_fields():
    from pytensor_federated_b200 import _pb
    from pytensor_federated_b200.npproto import Ndarray

    # shape says 4 float64 (32 bytes), the data field holds 8: must not decode the dtype/shape bytes behind it
    item = Ndarray(data=np.arange(1.0, 2.0).tobytes(), dtype="float64", shape=[4], strides=[8])
    message = _pb.enc_len_field(1, bytes(item)) + _pb.enc_len_field(2, b"u")
    with pytest.raises(ValueError, match="malformed"):
        native_codec.decode_arrays(message)
    with pytest.raises(Exception):
        ndarray_to_numpy(item)   # the Python codec agrees


def test_more_than_16_dimensions_round_trip_through_the_fallback():
    a = np.arange(2.0**18).reshape((2,) * 18)
    wire = bytes(InputArrays.from_arrays([a], uuid="deep"))
    with pytest.raises(TypeError):          # the fast path declines ...
        native_codec.decode_arrays(wire)
    parsed = InputArrays.FromString(wire)   # ... and the message class falls back to the general decoder
    assert parsed.uuid == "deep"
    (got,) = parsed.arrays
    assert got.shape == a.shape
    np.testing.assert_array_equal(got, a)
