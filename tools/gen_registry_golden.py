"""tests/golden/registry_golden.json: what the reference's gym_guided_vision/__init__.py registers (ids, entry points, kwargs),
recorded by EXECUTING that file with a recording stand-in for gymnasium's `register`."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
calls = []
reg = types.ModuleType("gymnasium.envs.registration")
reg.register = lambda **kw: calls.append(kw)
for name in ("gymnasium", "gymnasium.envs"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["gymnasium.envs.registration"] = reg
src = open("/root/reference/gym_guided_vision/gym_guided_vision/__init__.py").read()
exec(compile(src, "reference/__init__.py", "exec"), {"__name__": "gym_guided_vision"})
with open(os.path.join(ROOT, "tests", "golden", "registry_golden.json"), "w") as fh:
    json.dump(calls, fh, indent=1)
print(len(calls), "registrations recorded")
