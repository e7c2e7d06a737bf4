"""Generate tests/golden/reward_golden.json by EXECUTING the reference's own get_reward methods.

dm_control / MuJoCo are absent, so the task classes cannot be instantiated; but each `get_reward`
(gym_guided_vision/gym_guided_vision/env.py:425-472, 546-589, 640-690, 738-779, 820-863) only reads
`self._physics.data.ncon`, `self._physics.data.contact[i].geom1 / .geom2`, `self._physics.model.id2name(id, 'geom')` and -- for
SewNeedle -- the latched attribute `self._threaded_needle`.  This script cuts the method out of the reference source with `ast`,
executes it unmodified against a duck-typed `self` whose geom ids / names are those of OUR compiled model (unnamed geoms are ''
as in dm_control), and records the reward for: every unordered pair of "interesting" geoms as a single contact (all geoms with a
reward class + three unnamed ones), 250 random contact lists of 2-6 pairs over those, and 250 denser lists over the geoms the predicates name.  So the fixture pins both the staged-reward logic
and the model compiler's geom -> class assignment against the reference's string predicates.

    python tools/gen_reward_golden.py
"""
import ast
import json
import os
import sys
import textwrap
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from av_aloha_b200 import model_io  # noqa: E402

REF = "/root/reference/gym_guided_vision/gym_guided_vision/env.py"
TASKS = {"InsertPegEnv": "insert_peg", "SlotInsertionEnv": "slot_insertion", "SewNeedleEnv": "sew_needle",
         "TubeTransferEnv": "tube_transfer", "HookPackageEnv": "hook_package"}


def reference_get_reward(src, tree, cls):
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls)
    fn = next(n for n in node.body if isinstance(n, ast.FunctionDef) and n.name == "get_reward")
    ns = {}
    exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
    return ns["get_reward"]


def fake_self(names, pairs, latch):
    contacts = [SimpleNamespace(geom1=a, geom2=b) for a, b in pairs]
    model = SimpleNamespace(id2name=lambda i, kind: names[i])
    return SimpleNamespace(_physics=SimpleNamespace(data=SimpleNamespace(ncon=len(contacts), contact=contacts), model=model),
                           _threaded_needle=bool(latch))


def main():
    src = open(REF).read()
    tree = ast.parse(src)
    rng = np.random.default_rng(2024)
    out = {}
    for cls, task in TASKS.items():
        fn = reference_get_reward(src, tree, cls)
        names = model_io.load_names(task, 3)["geom"]
        classes = model_io.load_avm(model_io.model_path(task, 3))["geom_class"]
        ids = [i for i, c in enumerate(classes) if c] + [i for i, c in enumerate(classes) if not c][:3]
        lists = [[(a, b) if rng.random() < 0.5 else (b, a)] for k, a in enumerate(ids) for b in ids[k + 1:]]
        lists.append([])
        for _ in range(250):
            n = int(rng.integers(2, 7))
            lists.append([tuple(int(x) for x in rng.choice(ids, 2, replace=False)) for _ in range(n)])
        # denser lists over the geoms the predicates name (task objects, table, two pads per hand): reaches the rewards that need
        # several simultaneous contacts (e.g. both hands touching while the object still rests on the table)
        core = [i for i, c in enumerate(classes) if c and not names[i].startswith(("left", "right"))]
        core += [i for i, n_ in enumerate(names) if n_.startswith("left")][:2] + [i for i, n_ in enumerate(names) if n_.startswith("right")][:2]
        for _ in range(250):
            n = int(rng.integers(2, 9))
            lists.append([tuple(int(x) for x in rng.choice(core, 2, replace=False)) for _ in range(n)])
        cases = []
        for pairs in lists:
            for latch in ((0, 1) if task == "sew_needle" else (0,)):
                s = fake_self(names, pairs, latch)
                r = fn(s)
                cases.append([[int(x) for p in pairs for x in p], latch, int(r), int(bool(s._threaded_needle))])
        out[task] = {"ngeom": len(names), "cases": cases}
        print(task, len(cases), "cases, rewards", sorted({c[2] for c in cases}))
    dst = os.path.join(ROOT, "tests", "golden", "reward_golden.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
