"""Address validation & misc utils (mirrors fed/tests/without_ray_tests/test_utils.py)."""
import pytest

import rayfed_b200 as fed
import rayfed_b200.utils as fed_utils


@pytest.mark.parametrize(
    "address, valid",
    [
        ("192.168.0.1:8080", True),
        ("sa127032as:80", True),
        ("https://www.example.com", True),
        ("http://www.example.com", True),
        ("local", True),
        ("localhost", True),
        (None, False),
        ("invalid_string", False),
        ("http", False),
        ("example.com", False),
    ],
)
def test_validate_address(address, valid):
    if valid:
        fed_utils.validate_address(address)
    else:
        with pytest.raises(ValueError):
            fed_utils.validate_address(address)


def test_dict2tuple():
    assert fed_utils.dict2tuple({"k": "v", "a": 1}) == (("k", "v"), ("a", 1))
    assert fed_utils.dict2tuple(None) is None
    assert fed_utils.dict2tuple((("k", "v"),)) == (("k", "v"),)


def test_exports():
    for name in ("init", "remote", "get", "kill", "shutdown", "send", "recv", "FedObject", "FedRemoteError"):
        assert hasattr(fed, name)


def test_remote_error_str():
    assert str(fed.FedRemoteError("alice")) == "FedRemoteError occurred at alice"
    assert "caused by boom" in str(fed.FedRemoteError("alice", ValueError("boom")))
