"""Generate tests/golden/reset_golden.json by EXECUTING the reference's own reset code.

MuJoCo / dm_control are absent, so GuidedVisionEnv cannot be instantiated; but the object-placement part of every
task's reset() (gym_guided_vision/gym_guided_vision/env.py:474-501, 513-543, 604-637, 705-735, 792-818) is plain numpy
on the GLOBAL np.random state.  This script cuts those statements out of the reference source text (from
"# reset physics" to the first `self._physics.bind`), executes them unmodified under np.random.seed(s) and records,
per task, which `<name>_position` ends up in which free joint (read from the bind lines that follow).  Dead draws
(e.g. `peg_position` at env.py:525) therefore advance the RNG exactly as they do in the reference.
"""
import json
import os
import re

import numpy as np

REF = "/root/reference/gym_guided_vision/gym_guided_vision/env.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TASKS = {"InsertPegEnv": "insert_peg", "SlotInsertionEnv": "slot_insertion", "SewNeedleEnv": "sew_needle",
         "TubeTransferEnv": "tube_transfer", "HookPackageEnv": "hook_package"}


def main():
    src = open(REF).read()
    out = {}
    for cls, task in TASKS.items():
        body = src[src.index(f"class {cls}("):]
        body = body[body.index("    def reset(self"):]
        start = body.index("# reset physics")
        end = body.index("self._physics.bind")
        block = "\n".join(l[8:] if l.startswith("        ") else l.strip() for l in body[start:end].split("\n"))
        binds = re.findall(r"self\._physics\.bind\(self\._(\w+)\)\.qpos = np\.concatenate\(\[(\w+), (\w+)\]\)",
                           body[end:body.index("self._physics.forward()")])
        cases = []
        for seed in (1000, 1234, 7):
            np.random.seed(seed)
            eps = []
            for _ in range(3):                      # three consecutive resets from one RNG stream
                ns = {"np": np}
                exec(block, ns)
                eps.append({j: [float(x) for x in ns[p]] + [float(x) for x in ns[q]] for j, p, q in binds})
            cases.append({"seed": seed, "episodes": eps})
        out[task] = {"joints": [j for j, _, _ in binds], "cases": cases}
    dst = os.path.join(ROOT, "tests", "golden", "reset_golden.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main()
