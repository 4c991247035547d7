"""TLS on the gRPC path: server identity, client verification, mutual TLS, plaintext clients rejected."""
import asyncio
import datetime
import ipaddress

import numpy as np
import pytest

from pytensor_federated_b200 import ArraysToArraysServiceClient, service
from pytensor_federated_b200.config import TlsConfig, tls_from_env
from pytensor_federated_b200.rpc import Server
from pytensor_federated_b200.utils import get_useful_event_loop

x509 = pytest.importorskip("cryptography.x509")
pytestmark = pytest.mark.timeout(120)


def _make_cert(common_name, issuer=None, is_ca=False):
    """(cert_pem, key_pem, cert, key); signed by ``issuer = (cert, key)`` or self-signed."""
    from cryptography.hazmat.primitives import hashes, serialization
    from cryptography.hazmat.primitives.asymmetric import ec
    from cryptography.x509.oid import NameOID

    key = ec.generate_private_key(ec.SECP256R1())
    name = x509.Name([x509.NameAttribute(NameOID.COMMON_NAME, common_name)])
    now = datetime.datetime.now(datetime.timezone.utc)
    builder = (
        x509.CertificateBuilder()
        .subject_name(name)
        .issuer_name(issuer[0].subject if issuer else name)
        .public_key(key.public_key())
        .serial_number(x509.random_serial_number())
        .not_valid_before(now - datetime.timedelta(minutes=5))
        .not_valid_after(now + datetime.timedelta(days=1))
        .add_extension(x509.BasicConstraints(ca=is_ca, path_length=None), critical=True)
        .add_extension(
            x509.SubjectAlternativeName([x509.DNSName("localhost"), x509.IPAddress(ipaddress.ip_address("127.0.0.1"))]),
            critical=False,
        )
    )
    cert = builder.sign(issuer[1] if issuer else key, hashes.SHA256())
    key_pem = key.private_bytes(serialization.Encoding.PEM, serialization.PrivateFormat.PKCS8, serialization.NoEncryption())
    return cert.public_bytes(serialization.Encoding.PEM), key_pem, cert, key


@pytest.fixture()
def pki():
    ca_pem, _, ca_cert, ca_key = _make_cert("test-ca", is_ca=True)
    server_pem, server_key, _, _ = _make_cert("localhost", issuer=(ca_cert, ca_key))
    client_pem, client_key, _, _ = _make_cert("client", issuer=(ca_cert, ca_key))
    yield {"ca": ca_pem, "server": (server_pem, server_key), "client": (client_pem, client_key)}
    service.set_default_tls(None)
    service._default_tls_set = False


def _serve_and_call(server_tls, client_tls, monkeypatch):
    monkeypatch.setenv("B200FED_CONNECT_SLEEP", "0,0")
    loop = get_useful_event_loop()
    server = Server([service.ArraysToArraysService(lambda a, b: [a * b])], tls=server_tls)
    port = loop.run_until_complete(server.start("127.0.0.1", 0))
    service.set_default_tls(client_tls)
    client = ArraysToArraysServiceClient("127.0.0.1", port)
    try:
        (out,) = client.evaluate(np.array(6.0), np.array(7.0), retries=0)
        load = loop.run_until_complete(service.get_load_async("127.0.0.1", port, timeout=2))
        return out, load
    finally:
        del client
        loop.run_until_complete(server.close(None))


def test_tls_roundtrip_and_load_probe(pki, monkeypatch):
    cert, key = pki["server"]
    out, load = _serve_and_call(TlsConfig(cert=cert, key=key), TlsConfig(ca=pki["ca"]), monkeypatch)
    assert out == 42.0 and load is not None and load.n_clients == 1


def test_plaintext_client_cannot_talk_to_a_tls_node(pki, monkeypatch):
    cert, key = pki["server"]
    with pytest.raises((service.StreamTerminatedError, TimeoutError)):
        _serve_and_call(TlsConfig(cert=cert, key=key), None, monkeypatch)


def test_untrusted_server_certificate_is_rejected(pki, monkeypatch):
    other_ca, _, _, _ = _make_cert("someone-else", is_ca=True)
    cert, key = pki["server"]
    with pytest.raises((service.StreamTerminatedError, TimeoutError)):
        _serve_and_call(TlsConfig(cert=cert, key=key), TlsConfig(ca=other_ca), monkeypatch)


def test_mutual_tls_requires_a_client_certificate(pki, monkeypatch):
    cert, key = pki["server"]
    server_tls = TlsConfig(ca=pki["ca"], cert=cert, key=key, mutual=True)
    ccert, ckey = pki["client"]
    out, _ = _serve_and_call(server_tls, TlsConfig(ca=pki["ca"], cert=ccert, key=ckey), monkeypatch)
    assert out == 42.0
    with pytest.raises((service.StreamTerminatedError, TimeoutError)):
        _serve_and_call(server_tls, TlsConfig(ca=pki["ca"]), monkeypatch)


def test_tls_from_env_reads_pem_files(pki, tmp_path, monkeypatch):
    assert tls_from_env() is None
    (tmp_path / "ca.pem").write_bytes(pki["ca"])
    monkeypatch.setenv("B200FED_TLS_CA", str(tmp_path / "ca.pem"))
    monkeypatch.setenv("B200FED_TLS_SERVER_NAME", "localhost")
    cfg = tls_from_env()
    assert cfg.ca == pki["ca"] and cfg.cert is None and cfg.server_name == "localhost" and not cfg.mutual


def test_incomplete_tls_material_fails_closed(pki, tmp_path, monkeypatch):
    """Partial TLS material must raise for the role, never degrade to plaintext (server and client)."""
    import asyncio

    from pytensor_federated_b200.config import TlsConfigError
    from pytensor_federated_b200.rpc import Server

    cert, key = pki["server"]
    # server: certificate without key, and mutual TLS without a CA
    for bad in (TlsConfig(cert=cert), TlsConfig(ca=pki["ca"]), TlsConfig(cert=cert, key=key, mutual=True)):
        with pytest.raises(TlsConfigError, match="server"):
            asyncio.new_event_loop().run_until_complete(Server([], tls=bad).start("127.0.0.1", 0))
    # the environment counts: a lone B200FED_TLS_CERT is a request for TLS
    (tmp_path / "c.pem").write_bytes(cert)
    monkeypatch.setenv("B200FED_TLS_CERT", str(tmp_path / "c.pem"))
    with pytest.raises(TlsConfigError, match="key"):
        asyncio.new_event_loop().run_until_complete(Server([]).start("127.0.0.1", 0))
    # ... unless plaintext is asked for explicitly
    loop = asyncio.new_event_loop()
    srv = Server([], tls=False)
    assert loop.run_until_complete(srv.start("127.0.0.1", 0)) > 0
    loop.run_until_complete(srv.close())
    monkeypatch.delenv("B200FED_TLS_CERT")
    # client: identity without the CA, or half an identity
    with pytest.raises(TlsConfigError, match="client"):
        TlsConfig(cert=cert, key=key).check_client()
    with pytest.raises(TlsConfigError, match="key"):
        TlsConfig(ca=pki["ca"], cert=cert).check_client()
    service.set_default_tls(TlsConfig(cert=cert, key=key))
    try:
        with pytest.raises(TlsConfigError):
            service._Channel("127.0.0.1", 1)
    finally:
        service.set_default_tls(None)
