import asyncio

from pytensor_federated_b200 import utils
from pytensor_federated_b200.rpc import GetLoadResult


def test_argmin_ignores_none():
    assert utils.argmin_none_or_func([], float) is None
    assert utils.argmin_none_or_func([None, None], float) is None
    loads = [
        None,
        GetLoadResult(3, 0.5, 0.2),
        None,
        GetLoadResult(2, 0.05, 0.4),
        GetLoadResult(1, 0.1, 0.6),
    ]
    assert utils.argmin_none_or_func(loads, lambda l: l.percent_ram) == 1
    assert utils.argmin_none_or_func(loads, lambda l: l.percent_cpu) == 3
    assert utils.argmin_none_or_func(loads, lambda l: l.n_clients) == 4
    # ties resolve to the first occurrence
    assert utils.argmin_none_or_func([None, 2, 1, 1], float) == 2


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
