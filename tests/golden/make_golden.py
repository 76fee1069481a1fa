"""Generates tests/golden/cfg1_and_batch.npz.

The reference's own arithmetic (CasADi 3.7.1 -> IPOPT) cannot run in this image, so these vectors are
NOT reference outputs: they are the oracle's answers (dense Bunch-Kaufman path) for BASELINE cfg 1 plus 15 seeded
instances, together with the answers of an independent solver (SciPy SLSQP with analytic derivatives, polished from
the oracle's point).  PARITY UNPINNED stays true; the fixture pins the oracle against regressions and against SLSQP.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kiss_mpc_b200.synthetic import cfg1_instance, make_batch  # noqa: E402
from oracle import oracle as ok  # noqa: E402
from oracle.nlp_numpy import NLP, slsqp_polish  # noqa: E402

cfg = ok.OracleConfig(linsolve="dense")
xc, gl = cfg1_instance()
b = make_batch(15, seed=1001)
x_cur = np.concatenate([xc, b["x_cur"]]); goal = np.concatenate([gl, b["goal"]])
r = ok.solve(cfg, x_cur, goal)
Us, fs = [], []
for i in range(len(x_cur)):
    Xp, Up, fp = slsqp_polish(NLP(cfg, x_cur[i], goal[i]), r.X[i], r.U[i])
    Us.append(Up); fs.append(fp)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cfg1_and_batch.npz"), x_cur=x_cur, goal=goal, X=r.X, U=r.U,
                    obj=r.obj, status=r.status, iters=r.iters, U_slsqp=np.array(Us), obj_slsqp=np.array(fs))
print("objective cfg1:", r.obj[0], "iters:", r.iters, "max|U-U_slsqp|:", np.abs(np.array(Us) - r.U).max())
