"""Staged rewards against goldens produced by EXECUTING the reference's own get_reward methods (tools/gen_reward_golden.py;
reference env.py:425-472, 546-589, 640-690, 738-779, 820-863) on explicit contact lists: integer work, bit-exact.

Pins both the reward staging (incl. SewNeedle's latched `_threaded_needle`) and the model compiler's geom -> class assignment
against the reference's string predicates (`== "stick"`, `startswith("right")`, `startswith("hole-")`, ...), for the fp64 oracle
and for the CUDA `stage_reward` source run by the host emulator.  (The GPU kernels get their contact lists from the collision
stage only; tests/test_gpu_tasks.py compares their rewards with the oracle's on real contact sets.)
"""
import json
import os

import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reward_golden.json")
TASKS = ["insert_peg", "slot_insertion", "sew_needle", "tube_transfer", "hook_package"]


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as fh:
        return json.load(fh)


@pytest.mark.parametrize("task", TASKS)
def test_oracle_rewards_match_the_reference(gold, task):
    from av_aloha_b200 import model_io
    from oracle.oracle import OracleEnv, OracleModel
    o = OracleEnv(OracleModel(model_io.model_path(task, 3)))
    assert o.model.ngeom == gold[task]["ngeom"]
    seen = set()
    for flat, latch, reward, latch_out in gold[task]["cases"]:
        assert o.reward_from_pairs(flat, latch) == (reward, latch_out), (flat, latch)
        seen.add(reward)
    assert len(seen) >= 4


@pytest.mark.parametrize("task", TASKS)
def test_kernel_source_rewards_match_the_reference(gold, task):
    from av_aloha_b200 import model_io
    from tests.emu import emu
    eb = emu.EmuBatch(model_io.model_path(task, 3), 1)
    for flat, latch, reward, latch_out in gold[task]["cases"]:
        assert emu.emu_reward_from_pairs(eb, flat, latch) == (reward, latch_out), (flat, latch)


def test_two_arm_models_share_the_class_assignment(gold):
    """the 2-arm variants park the middle arm but keep geom ids and names (reference env.py:60-62): same rewards"""
    from av_aloha_b200 import model_io
    from oracle.oracle import OracleEnv, OracleModel
    for task in ("insert_peg", "hook_package"):
        assert model_io.load_names(task, 2)["geom"] == model_io.load_names(task, 3)["geom"]
        o = OracleEnv(OracleModel(model_io.model_path(task, 2)))
        for flat, latch, reward, latch_out in gold[task]["cases"][::7]:
            assert o.reward_from_pairs(flat, latch) == (reward, latch_out)
