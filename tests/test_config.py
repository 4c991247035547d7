from pytensor_federated_b200.config import FederationConfig, get_config


def test_defaults_and_env_overrides(monkeypatch):
    for k in ("B200FED_COMM", "B200FED_TIMEOUT", "B200FED_CONNECT_SLEEP", "B200FED_NO_MULTICAST", "B200FED_SPECULATIVE_US"):
        monkeypatch.delenv(k, raising=False)
    cfg = get_config()
    assert cfg == FederationConfig()
    monkeypatch.setenv("B200FED_COMM", "ipc")
    monkeypatch.setenv("B200FED_TIMEOUT", "3.5")
    monkeypatch.setenv("B200FED_CONNECT_SLEEP", "0,0")
    monkeypatch.setenv("B200FED_NO_MULTICAST", "1")
    monkeypatch.setenv("B200FED_SERVE_AHEAD", "not-a-number")
    monkeypatch.setenv("B200FED_SPECULATIVE_US", "750")
    cfg = get_config()
    assert (cfg.comm, cfg.timeout, cfg.connect_sleep, cfg.multicast, cfg.serve_ahead) == ("ipc", 3.5, (0.0, 0.0), False, 8)
    assert cfg.speculative_us == 750.0 and FederationConfig().speculative_us == 0.0   # off unless asked for


def test_operator_cli_info(capsys):
    import json

    from pytensor_federated_b200.__main__ import main

    assert main(["info"]) == 0
    info = json.loads(capsys.readouterr().out)
    assert info["version"] and info["graph_backend"] in ("builtin", "pytensor")
    assert set(info["config"]) >= {"comm", "timeout", "serve_ahead", "speculative_us"} and info["tls"] is None
    assert "built" in info["native_library"]
