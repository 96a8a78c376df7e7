"""rayfed_b200 — a Blackwell-native cross-party federated execution engine.

Drop-in API of ray-project/rayfed (``import rayfed_b200 as fed``)::

    fed.init(addresses=..., party="alice")
    @fed.remote
    def f(x): ...
    y = f.party("bob").remote(x)      # runs in bob only; alice pushes x
    fed.get(y)                        # collective: every party gets the value
    fed.shutdown()

plus the B200-native federated-learning layer (``fed.fl``), the NVLink
transport and the sm_100a kernels (``fed.ops``).
"""
import os as _os

# Streams are multiplexed onto a small number of GPU front-end queues (default 8);
# kernels that wait for a peer must not sit in front of unrelated local work.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# A receive parks a stream on a peer's flag (cuStreamWaitValue32).  With CUDA's default LAZY
# module loading, the *first* launch of any kernel has to load its module, which synchronises
# with the context — i.e. blocks the launching host thread for as long as a wait is parked
# (measured on B200: profiles/README.md "stream memory operations").  Between two parties that is
# a deadlock in the first round.  EAGER loading removes the hazard but costs tens of seconds of
# start-up (every cubin of torch / cuDNN / cuBLAS is loaded up front), so it is switched on only
# where a multi-party device job is evident — a torchrun / launcher environment with more than one
# rank — or when asked for (RAYFED_B200_EAGER_MODULES=1; =0 never).  Whatever the outcome, the
# device channel asks the driver which mode it really got and, under LAZY loading, gates each
# receive on the host (it polls the flag and only then enqueues the — already satisfied — stream
# wait), so no wait is ever parked and nothing can deadlock.
def _want_eager_modules() -> bool:
    flag = _os.environ.get("RAYFED_B200_EAGER_MODULES")
    if flag is not None:
        return flag != "0"
    try:
        return int(_os.environ.get("WORLD_SIZE", "1")) > 1 or bool(_os.environ.get("RAYFED_B200_PARTY"))
    except ValueError:
        return False


if _want_eager_modules():
    _os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

from .api import get, init, kill, remote, shutdown  # noqa: E402,F401
from .exceptions import FedRemoteError, FedTaskError  # noqa: E402,F401
from .fed_object import FedObject  # noqa: E402,F401
from .proxy.barriers import recv, send  # noqa: E402,F401



def aggregate(fed_objects, weights=None, fused=None):
    """Cross-party (weighted) mean of one value per party — see :mod:`rayfed_b200.collective`."""
    from .collective import aggregate as _aggregate

    return _aggregate(fed_objects, weights=weights, fused=fused)


def init_from_env(**overrides):
    """``fed.init`` with addresses / party / device / transport taken from the environment set
    by ``python -m rayfed_b200.tool.launch`` (keyword arguments override, ``config`` is merged)."""
    import json

    from .tool import launch as _launch

    env = _os.environ
    if _launch.ENV_PARTY not in env or _launch.ENV_ADDRESSES not in env:
        raise RuntimeError("init_from_env: not started by `python -m rayfed_b200.tool.launch` "
                           f"({_launch.ENV_PARTY} / {_launch.ENV_ADDRESSES} are not set)")
    config = {"cross_silo_comm": {"transport": env.get(_launch.ENV_TRANSPORT, "auto")},
              "barrier_on_initializing": True}
    if env.get(_launch.ENV_DEVICE):
        config["device"] = env[_launch.ENV_DEVICE]
    user_cfg = dict(overrides.pop("config", None) or {})
    comm = dict(config["cross_silo_comm"])
    comm.update(user_cfg.pop("cross_silo_comm", None) or {})
    config.update(user_cfg)
    config["cross_silo_comm"] = comm
    kwargs = dict(addresses=json.loads(env[_launch.ENV_ADDRESSES]), party=env[_launch.ENV_PARTY], config=config,
                  job_name=env.get(_launch.ENV_JOB))
    kwargs.update(overrides)
    return init(**kwargs)


__all__ = [
    "aggregate",
    "init_from_env",
    "get",
    "init",
    "kill",
    "remote",
    "shutdown",
    "recv",
    "send",
    "FedObject",
    "FedRemoteError",
    "FedTaskError",
]

__version__ = "0.1.0"
