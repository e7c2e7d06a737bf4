"""GPU forward pass (Newton) on the fixture's contact-rich states vs the fp64 oracle's Newton (1e-13):
 (a) oracle on the GPU's contact list (the solve alone), (b) oracle with its own narrowphase (end to end).
Prints the distribution of |dqacc|inf / max(1, |qacc|inf) and of the energy-norm error sqrt(da' M da) / sqrt(1 + a' M a)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from oracle.oracle import OracleModel
from test_solver_newton import TASKS, _gpu_forward, _oracle, _rel


def q(x):
    x = np.asarray(x)
    return "median %.2e p90 %.2e p99 %.2e max %.2e  frac>1e-2 %.3f" % (np.median(x), np.quantile(x, .9), np.quantile(x, .99), x.max(), np.mean(x > 1e-2))


for task in TASKS:
    path, st, g = _gpu_forward(task)
    om = OracleModel(path)
    inj, own, en, worst_dof = [], [], [], []
    for e in range(len(st["qpos"])):
        c = g["contacts"][e][: g["ncon"][e]]
        o = _oracle(om, st, e, contacts=np.concatenate([c[:, 0:7], c[:, 7:9]], axis=1))
        d = g["qacc"][e] - o.qacc
        inj.append(_rel(g["qacc"][e], o.qacc))
        worst_dof.append(int(np.abs(d).argmax()))
        M = o.M
        en.append(float(np.sqrt(d @ M @ d) / np.sqrt(1.0 + o.qacc @ M @ o.qacc)))
        o2 = _oracle(om, st, e)
        own.append(_rel(g["qacc"][e], o2.qacc))
    print(task, "N", len(inj), "newton iters mean %.2f max %.0f  scaled grad max %.1e" % (g["stat"][:, 0].mean(), g["stat"][:, 0].max(), g["stat"][:, 1].max()))
    print("   same contacts, max norm   :", q(inj))
    print("   same contacts, energy norm:", q(en))
    print("   own narrowphase, max norm :", q(own))
    bad = np.nonzero(np.array(inj) > 1e-2)[0]
    print("   states > 1e-2 (same contacts):", [(int(e), "%.1e" % inj[e], "dof", worst_dof[e]) for e in bad][:12])
