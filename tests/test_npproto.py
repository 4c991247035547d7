"""Array codec and wire-format tests (CPU).  Mirrors the intent of the
reference's ``test_npproto.py:11-31`` and adds the cases it never covered
(non-contiguous input, byte-level schema pinning, fuzzing)."""
from datetime import datetime

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st
from hypothesis.extra import numpy as hnp

from pytensor_federated_b200 import _pb, npproto
from pytensor_federated_b200.npproto import utils
from pytensor_federated_b200.rpc import GetLoadParams, GetLoadResult, InputArrays, OutputArrays

ROUNDTRIP_CASES = [
    np.arange(5),
    np.random.default_rng(0).uniform(size=(2, 3)),
    np.array(5),
    np.array(["hello", "world"]),
    np.array([datetime(2020, 3, 4, 5, 6, 7, 8), datetime(2020, 3, 4, 5, 6, 7, 9)]),
    np.datetime64("2022-06"),
    np.array([(1, 2), (3, 2, 1)], dtype=object),
    np.zeros((0, 4), dtype=np.float32),
    np.array(1.5, dtype=np.float16),
    np.arange(24, dtype=np.int16).reshape(2, 3, 4),
    np.array([True, False]),
    np.array([1 + 2j, 3 - 4j]),
]


@pytest.mark.parametrize("arr", ROUNDTRIP_CASES, ids=lambda a: f"{np.asarray(a).dtype}-{np.asarray(a).shape}")
def test_roundtrip_through_bytes(arr, monkeypatch):
    if np.asarray(arr).dtype.hasobject:
        with pytest.raises(TypeError, match="Refusing to unpickle"):     # opt-in only: bytes from a peer
            utils.ndarray_to_numpy(utils.ndarray_from_numpy(arr))
        monkeypatch.setenv("B200FED_ALLOW_PICKLE", "1")
    nda = utils.ndarray_from_numpy(arr)
    enc = bytes(nda)
    dec = npproto.Ndarray().parse(enc)
    assert isinstance(dec.data, bytes)
    result = utils.ndarray_to_numpy(dec)
    assert result.dtype == np.asarray(arr).dtype
    assert result.shape == np.asarray(arr).shape
    np.testing.assert_array_equal(result, arr)


@pytest.mark.parametrize(
    "make",
    [
        lambda a: a.T,
        lambda a: np.asfortranarray(a),
        lambda a: a[::2],
        lambda a: a[:, ::-1],
        lambda a: a[1:, 1:3],
    ],
    ids=["transposed", "fortran", "strided", "negative-stride", "window"],
)
def test_non_contiguous_inputs_are_canonicalised(make):
    """The reference corrupts or rejects these; we must not."""
    base = np.arange(20, dtype=np.float64).reshape(4, 5)
    view = make(base)
    dec = npproto.Ndarray().parse(bytes(utils.ndarray_from_numpy(view)))
    np.testing.assert_array_equal(utils.ndarray_to_numpy(dec), view)
    # what is on the wire is C-ordered and self-consistent
    assert list(dec.strides) == list(np.ascontiguousarray(view).strides)


def test_decoded_arrays_are_readonly_views():
    arr = np.arange(6.0).reshape(2, 3)
    out = utils.ndarray_to_numpy(npproto.Ndarray().parse(bytes(utils.ndarray_from_numpy(arr))))
    assert not out.flags.writeable
    with pytest.raises(ValueError):
        out[0, 0] = 1.0


def test_wire_bytes_are_pinned():
    """Tags and packing must equal the reference schema (ndarray.proto:7-12)."""
    arr = np.array([[1, 2, 3]], dtype=np.int8)
    enc = bytes(utils.ndarray_from_numpy(arr))
    expected = (
        b"\x0a\x03\x01\x02\x03"  # field 1 (bytes) len 3
        b"\x12\x04int8"  # field 2 (string)
        b"\x1a\x02\x01\x03"  # field 3 packed shape [1, 3]
        b"\x22\x02\x03\x01"  # field 4 packed strides [3, 1]
    )
    assert enc == expected
    # 0-d: shape/strides are elided (proto3 defaults)
    enc0 = bytes(utils.ndarray_from_numpy(np.array(7, dtype=np.uint8)))
    assert enc0 == b"\x0a\x01\x07\x12\x05uint8"


def test_unpacked_repeated_and_unknown_fields_are_accepted():
    # shape sent un-packed (two varint fields) + an unknown field 9 — legal proto3
    raw = b"\x0a\x02\x01\x02" + b"\x12\x04int8" + b"\x18\x02" + b"\x18\x01" + b"\x48\x05"
    nda = npproto.Ndarray().parse(raw)
    assert nda.shape == [2, 1]
    np.testing.assert_array_equal(utils.ndarray_to_numpy(nda), np.array([[1], [2]], dtype=np.int8))


def test_negative_int64_varint():
    enc = _pb.enc_packed_int64(4, [-8, 8])
    out = []
    ((field, wt, value),) = list(_pb.iter_fields(enc))
    _pb.dec_packed_int64(value, wt, out)
    assert field == 4 and out == [-8, 8]
    assert len(_pb.encode_varint(-1)) == 10


def test_truncated_messages_raise():
    enc = bytes(utils.ndarray_from_numpy(np.arange(10)))
    with pytest.raises(ValueError):
        npproto.Ndarray().parse(enc[:10])  # cut inside the 80-byte data field
    with pytest.raises(ValueError):
        list(_pb.iter_fields(b"\x0a\x7f\x00"))


def test_service_messages_roundtrip():
    a = utils.ndarray_from_numpy(np.array([1.0, 2.0]))
    b = utils.ndarray_from_numpy(np.array(3))
    msg = InputArrays(items=[a, b], uuid="abc-123")
    dec = InputArrays().parse(bytes(msg))
    assert dec == msg and dec.uuid == "abc-123" and len(dec.items) == 2
    out = OutputArrays.FromString(bytes(OutputArrays(items=[b], uuid="u")))
    assert out.items[0] == b
    assert bytes(GetLoadParams()) == b""
    load = GetLoadResult(n_clients=3, percent_cpu=12.5, percent_ram=50.0)
    enc = bytes(load)
    assert enc[0] == 0x08 and enc[2] == 0x15 and enc[7] == 0x1D
    assert GetLoadResult.FromString(enc) == load
    assert bytes(GetLoadResult()) == b""
    assert GetLoadResult.FromString(b"") == GetLoadResult(0, 0.0, 0.0)


def test_interop_with_google_protobuf_runtime():
    """Cross-check the hand-written codec against protobuf's own encoder using a
    descriptor built at runtime (protoc is not available in this image)."""
    pytest.importorskip("google.protobuf")
    from pytensor_federated_b200.protocol import build_message_classes

    classes = build_message_classes()
    arr = np.arange(12, dtype=np.float32).reshape(3, 4)
    mine = utils.ndarray_from_numpy(arr)
    theirs = classes["npproto.ndarray"](
        data=mine.data, dtype=mine.dtype, shape=mine.shape, strides=mine.strides
    )
    assert theirs.SerializeToString() == bytes(mine)
    msg = classes["InputArrays"](uuid="u-1")
    msg.items.append(theirs)
    parsed = InputArrays().parse(msg.SerializeToString())
    np.testing.assert_array_equal(utils.ndarray_to_numpy(parsed.items[0]), arr)
    res = classes["GetLoadResult"](n_clients=7, percent_cpu=1.25, percent_ram=99.5)
    assert GetLoadResult.FromString(res.SerializeToString()) == GetLoadResult(7, 1.25, 99.5)
    assert bytes(GetLoadResult(7, 1.25, 99.5)) == res.SerializeToString()


@settings(max_examples=60, deadline=None)
@given(
    hnp.arrays(
        dtype=st.sampled_from([np.float64, np.float32, np.int64, np.int32, np.uint8, np.bool_]),
        shape=hnp.array_shapes(min_dims=0, max_dims=4, min_side=0, max_side=5),
    )
)
def test_fuzz_roundtrip(arr):
    dec = npproto.Ndarray().parse(bytes(utils.ndarray_from_numpy(arr)))
    out = utils.ndarray_to_numpy(dec)
    assert out.dtype == arr.dtype and out.shape == arr.shape
    np.testing.assert_array_equal(out, arr)


def test_torch_helpers():
    torch = pytest.importorskip("torch")
    t = torch.arange(6, dtype=torch.bfloat16).reshape(2, 3)
    nda = utils.ndarray_from_tensor(t)
    assert nda.dtype == "float32"
    back = utils.ndarray_to_tensor(npproto.Ndarray().parse(bytes(nda)))
    assert torch.equal(back, t.float())
