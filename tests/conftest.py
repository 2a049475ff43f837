import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "emu")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# whole-trainer tests (PPO.train / run_experiment train): run after every kernel / parity / data-path test, so that with -x a
# problem in the trainer's host code cannot hide the state of the device path
_TRAINER_TESTS = ("test_graph_replayed_update_matches_the_eager_update", "test_ppo_update_changes_weights_and_returns_seven_scalars",
                  "test_same_seed_gives_bit_identical_weights", "test_ppo_trains_on_", "test_train_then_eval",
                  "test_n_rank_ppo_replicas_stay_identical", "test_uneven_shards_are_rejected")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("LHW_TEST_NATURAL_ORDER", "0") != "1":
        items.sort(key=lambda it: next((k + 1 for k, name in enumerate(_TRAINER_TESTS) if name in it.nodeid), 0))   # stable
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_tight():
    from oracle.oracle import Oracle
    return Oracle(tolerance=1e-14)
