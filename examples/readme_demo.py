"""The reference README demo (README.md:24-60 of ray-project/rayfed), unchanged
except for the import: two parties, an actor each, a cross-party aggregate.

    python examples/readme_demo.py alice &  python examples/readme_demo.py bob
"""
import sys

import rayfed_b200 as fed


@fed.remote
class MyActor:
    def __init__(self, value):
        self.value = value

    def inc(self, num):
        self.value = self.value + num
        return self.value


@fed.remote
def aggregate(val1, val2):
    return val1 + val2


def main(party):
    addresses = {"alice": "127.0.0.1:11012", "bob": "127.0.0.1:11011"}
    fed.init(addresses=addresses, party=party)
    actor_alice = MyActor.party("alice").remote(1)
    actor_bob = MyActor.party("bob").remote(1)
    val_alice = actor_alice.inc.remote(1)
    val_bob = actor_bob.inc.remote(2)
    sum_val_obj = aggregate.party("bob").remote(val_alice, val_bob)
    result = fed.get(sum_val_obj)
    print(f"The result in party {party} is {result}")
    fed.shutdown()


if __name__ == "__main__":
    assert len(sys.argv) == 2, "Please run this script with party."
    main(sys.argv[1])
