#!/usr/bin/env python
"""Oracle pins at full count: per problem form, >= 1,000 instances each get
  * the first-order certificate on the unscaled problem from the oracle's own multipliers,
  * the second-order check (reduced Hessian of the Lagrangian on the null space of the active constraints),
  * a SciPy SLSQP polish (independent solver, analytic derivatives) started from the oracle's point,
and the dense LDL^T path is compared with the Riccati path.  CPU only; writes profiles/r02_oracle_pins.json.
    python scripts/pin_oracle.py [count] [workers] [form,form,...]    (results are merged into the existing record)
tests/test_pins_cpu.py runs the same checks on smaller samples in the routine suite."""
import json
import multiprocessing as mp
import os
import sys
import time
from dataclasses import replace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as ok  # noqa: E402
from oracle.nlp_numpy import NLP, dual_certificate, reduced_hessian_min_eig, slsqp_polish  # noqa: E402
from test_pins_cpu import FORMS, form_batch  # noqa: E402


def one(args):
    name, i, x, g, obs, X, U, duals, df, obj = args
    cfg, _ = form_batch(ok, name, 1)
    yc, zL, zU, s, yd, vL = ok.split_duals(cfg, duals)
    nlp = NLP(cfg, x, g, obs=obs)
    c = dual_certificate(nlp, X, U, yc, zL, zU, df, yd=yd if cfg.O else None, vL=vL if cfg.O else None)
    eig = reduced_hessian_min_eig(nlp, X, U, yc, df, yd=yd if cfg.O else None)
    Xp, Up, fp = slsqp_polish(nlp, X, U)
    return dict(stat=c["stationarity"] * df, comp=c["complementarity"] * df, primal=c["primal"], eig=eig,
                dobj=abs(fp - obj) / abs(obj), dU=float(np.abs(Up - U).max()))


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else os.cpu_count()
    names = sys.argv[3].split(",") if len(sys.argv) > 3 else list(FORMS)
    path = os.path.join(ROOT, "profiles", "r02_oracle_pins.json")
    out = json.load(open(path)) if os.path.exists(path) else {"count_per_form": count, "forms": {}}
    for name in names:
        t0 = time.time()
        cfg, b = form_batch(ok, name, count)
        r = ok.solve(cfg, b["x_cur"], b["goal"], obs=b["obs"], want_duals=True)
        conv = np.where(r.status == 0)[0]
        jobs = [(name, int(i), b["x_cur"][i], b["goal"][i], None if not cfg.O else b["obs"][i], r.X[i], r.U[i], r.duals[i], float(r.meta["df"][i]), float(r.obj[i]))
                for i in conv]
        with mp.Pool(workers) as pool:
            res = pool.map(one, jobs, chunksize=4)
        nd = min(count, 128 if cfg.O else (256 if cfg.N > 30 else 1000))
        rd = ok.solve(replace(cfg, linsolve="dense"), b["x_cur"][:nd], b["goal"][:nd], obs=None if not cfg.O else b["obs"][:nd])
        cd = (rd.status == 0) & (r.status[:nd] == 0)
        agg = lambda k, f: float(f([q[k] for q in res]))
        out["forms"][name] = {
            "instances": count, "converged": int(len(conv)), "status_counts": {int(k): int(v) for k, v in zip(*np.unique(r.status, return_counts=True))},
            "scaled_stationarity_max": agg("stat", max), "scaled_complementarity_max": agg("comp", max), "primal_max": agg("primal", max),
            "reduced_hessian_min_eig_min": agg("eig", min), "slsqp_rel_objective_diff_max": agg("dobj", max),
            "slsqp_control_diff_max": agg("dU", max), "slsqp_control_diff_p99": float(np.percentile([q["dU"] for q in res], 99)),
            "dense_vs_riccati": {"instances": nd, "status_equal": float((rd.status == r.status[:nd]).mean()),
                                 "iterations_equal": float((rd.iters == r.iters[:nd]).mean()),
                                 "max_control_diff": float(np.abs(rd.U - r.U[:nd])[cd].max())},
            "seconds": time.time() - t0}
        print(name, json.dumps(out["forms"][name]), flush=True)
        json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
