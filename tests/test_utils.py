import asyncio

from pytensor_federated_b200 import utils
from pytensor_federated_b200.rpc import GetLoadResult


def test_argmin_none_or_func_skips_missing_entries():
    # nothing to choose from
    assert utils.argmin_none_or_func((), abs) is None
    assert utils.argmin_none_or_func(iter([None] * 3), abs) is None
    # plain numbers, unreachable servers in between
    values = [None, -7.5, None, None, 2.0, -1.0]
    assert utils.argmin_none_or_func(values, abs) == 5
    assert utils.argmin_none_or_func(values, lambda v: v) == 1
    assert utils.argmin_none_or_func(values, lambda v: -v) == 4
    # the first of several equal candidates wins; generators are fine
    assert utils.argmin_none_or_func((v for v in [None, 4, 4, 9]), float) == 1


def test_argmin_none_or_func_on_load_reports():
    # a fleet of six nodes, two of them down; each criterion picks a different survivor
    fleet = {
        0: GetLoadResult(n_clients=12, percent_cpu=81.0, percent_ram=35.5),
        1: None,
        2: GetLoadResult(n_clients=4, percent_cpu=97.5, percent_ram=12.25),
        3: GetLoadResult(n_clients=9, percent_cpu=3.5, percent_ram=64.0),
        4: None,
        5: GetLoadResult(n_clients=0, percent_cpu=44.0, percent_ram=90.0),
    }
    reports = [fleet[i] for i in range(6)]
    picked = {
        field: utils.argmin_none_or_func(reports, lambda rep, f=field: getattr(rep, f))
        for field in ("n_clients", "percent_cpu", "percent_ram")
    }
    assert picked == {"n_clients": 5, "percent_cpu": 3, "percent_ram": 2}


def test_event_loop_is_reentrant():
    assert asyncio._get_running_loop() is None
    loop = utils.get_useful_event_loop()
    assert isinstance(loop, asyncio.AbstractEventLoop)
    assert not loop.is_running()
    assert utils.get_useful_event_loop() is loop

    async def nested():
        assert loop.is_running()
        inner = utils.get_useful_event_loop()
        assert inner is loop
        assert hasattr(inner, "_nest_patched")
        # the whole point: running something *inside* a running loop
        inner.run_until_complete(asyncio.sleep(0.01))
        return 7

    assert loop.run_until_complete(nested()) == 7


def test_event_loop_in_fresh_thread():
    import threading

    seen = {}

    def work():
        loop = utils.get_useful_event_loop()
        seen["ok"] = loop.run_until_complete(asyncio.sleep(0, result=3))
        loop.close()

    t = threading.Thread(target=work)
    t.start()
    t.join()
    assert seen["ok"] == 3
