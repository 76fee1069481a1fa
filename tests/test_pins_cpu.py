"""CPU-side pins of the oracle beyond its golden vectors: real IPOPT is not installable here, so the
restatement is held against everything that IS available without it --
  * first-order certificates on the UNSCALED problem from the oracle's own multipliers (>= 1,000 instances per problem form),
  * second-order sufficiency (reduced Hessian positive definite) on a sample of each form,
  * an independent solver (SciPy SLSQP with analytic derivatives) polished from the oracle's answer,
  * the dense LDL^T-with-inertia linear algebra (what mirrors IPOPT + MUMPS) against the stage-wise Riccati solve,
  * per-class obstacle radii (optimizer.py:231-250) and the reference's own mpc/agent.py driven through the drop-in.
scripts/pin_oracle.py runs the same checks at 1,000 SLSQP polishes per form and keeps the record under profiles/."""
import importlib.util
import os
import sys
import types
from dataclasses import replace

import numpy as np
import pytest

from kiss_mpc_b200.synthetic import make_batch, make_tracks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

FORMS = {
    "box": dict(kw=dict(), seed=2101, O=0),
    "N50": dict(kw=dict(N=50), seed=2102, O=0),
    "O10": dict(kw=dict(O=10), seed=2103, O=10),
    "literal": dict(kw=dict(cost_mode="code_literal", goal_range="code", y_bounds=(-1e20, 1e20)), seed=2104, O=0),
    "tracks": dict(kw=dict(O=4, obs_stagewise=True), seed=2105, O=4),
}


def form_batch(ok, name, B, linsolve="riccati"):
    f = FORMS[name]
    cfg = ok.OracleConfig(linsolve=linsolve, **f["kw"])
    b = make_batch(B, seed=f["seed"], O=f["O"])
    if name == "tracks":
        b["obs"] = make_tracks(b["obs"], cfg.N, seed=17)
    return cfg, b


def certificates(ok, cfg, b, r, idx):
    from oracle.nlp_numpy import NLP, dual_certificate
    out = []
    for i in idx:
        yc, zL, zU, s, yd, vL = ok.split_duals(cfg, r.duals[i])
        nlp = NLP(cfg, b["x_cur"][i], b["goal"][i], obs=None if not cfg.O else b["obs"][i])
        out.append(dual_certificate(nlp, r.X[i], r.U[i], yc, zL, zU, r.meta["df"][i], yd=yd if cfg.O else None, vL=vL if cfg.O else None))
    return out


@pytest.mark.parametrize("name", list(FORMS))
def test_first_order_certificates_1024(oracle_mod, name):
    """Every converged answer of 1,024 fresh instances is a KKT point of the UNSCALED reference NLP to IPOPT's own tolerances:
    stationarity / complementarity are bounded by tol / df (the scaled tolerance 1e-8 undone), primal feasibility by 1e-8."""
    cfg, b = form_batch(oracle_mod, name, 1024)
    r = oracle_mod.solve(cfg, b["x_cur"], b["goal"], obs=b["obs"], want_duals=True)
    conv = np.where(r.status == 0)[0]
    assert len(conv) >= (1024 if name != "literal" else 1000)        # (the literal cost has a few max-iter instances)
    cs = certificates(oracle_mod, cfg, b, r, conv)
    df = r.meta["df"][conv]
    stat = np.array([c["stationarity"] for c in cs]); comp = np.array([c["complementarity"] for c in cs])
    assert (stat * df <= 1.01e-8 * 100).all()          # s_d <= s_max-scaled: E_0 <= tol on the scaled problem
    assert (comp * df <= 1.01e-8 * 100).all()
    assert max(c["primal"] for c in cs) <= 1e-8
    assert max(c["bound_violation"] for c in cs) <= 1e-7 * 20 and max(c["dual_sign"] for c in cs) == 0.0


@pytest.mark.parametrize("name", list(FORMS))
def test_second_order_and_slsqp_sample(oracle_mod, name):
    """Second-order sufficiency on 48 instances per form and an independent solver (SLSQP) polished from the oracle's point on
    6: objective within 1e-6 relative, controls within 1e-5 (the parity bar of north_star) wherever the minimiser is strict.
    (scripts/pin_oracle.py does the same on 1,000 / 250 per form: profiles/r02_oracle_pins.json.)"""
    from oracle.nlp_numpy import NLP, reduced_hessian_min_eig, slsqp_polish
    cfg, b = form_batch(oracle_mod, name, 48)
    r = oracle_mod.solve(cfg, b["x_cur"], b["goal"], obs=b["obs"], want_duals=True)
    conv = np.where(r.status == 0)[0]
    eigs = []
    for i in conv:
        yc, zL, zU, s, yd, vL = oracle_mod.split_duals(cfg, r.duals[i])
        nlp = NLP(cfg, b["x_cur"][i], b["goal"][i], obs=None if not cfg.O else b["obs"][i])
        eigs.append(reduced_hessian_min_eig(nlp, r.X[i], r.U[i], yc, r.meta["df"][i], yd=yd if cfg.O else None))
    eigs = np.array(eigs)
    if name == "literal":
        # goal cost over k = 1..N-1 and a linear v penalty: v_{N-1} has no curvature -> semi-definite is the most there is
        assert (eigs >= -1e-7).all()
    else:
        assert (eigs > 1e-6).all()
    for i in conv[:6]:
        nlp = NLP(cfg, b["x_cur"][i], b["goal"][i], obs=None if not cfg.O else b["obs"][i])
        Xp, Up, fp = slsqp_polish(nlp, r.X[i], r.U[i])
        assert abs(fp - r.obj[i]) <= 1e-6 * abs(r.obj[i])
        if name != "literal":
            # (SLSQP stops at its own accuracy on the UNRELAXED bounds while the interior-point answer sits on the 1e-8-relaxed ones
            #  at mu ~ 1e-9; with nearly active distance rows that difference reaches 1e-5, so those forms get 3e-5)
            assert np.abs(Up - r.U[i]).max() <= (1e-5 if not cfg.O else 3e-5)


@pytest.mark.parametrize("name,B", [("box", 1024), ("literal", 512), ("N50", 256), ("O10", 40), ("tracks", 96)])
def test_dense_ldl_vs_riccati(oracle_mod, name, B):
    """The linear algebra that mirrors IPOPT (full augmented system, Bunch-Kaufman LDL^T, inertia read off D) against the
    stage-wise Riccati solve the GPU kernels share: same statuses, same iteration counts, same controls."""
    cfg, b = form_batch(oracle_mod, name, B, linsolve="dense")
    rd = oracle_mod.solve(cfg, b["x_cur"], b["goal"], obs=b["obs"])
    rr = oracle_mod.solve(replace(cfg, linsolve="riccati"), b["x_cur"], b["goal"], obs=b["obs"])
    assert (rd.status == rr.status).all()
    conv = rd.status == 0
    assert np.abs(rd.U - rr.U)[conv].max() <= (1e-8 if not cfg.O else 1e-6)
    assert (np.abs(rd.obj - rr.obj) / np.abs(rd.obj))[conv].max() <= 1e-9
    assert (rd.iters == rr.iters).mean() >= (0.99 if not cfg.O else 0.9)


# ---------------- per-class obstacle radii (optimizer.py:231-250) ----------------
def test_per_class_radii_oracle_and_solver_sources(oracle_mod):
    """Static columns use the first static radius, dynamic columns the first dynamic radius: a radius per slot in the oracle, in
    the thread-solver source and in the warp-solver source (g++ builds), and the distances it enforces."""
    import emul
    from oracle.nlp_numpy import NLP, kkt_certificate
    O, B = 6, 24
    cfg = oracle_mod.OracleConfig(linsolve="riccati", O=O, inflation=0.5)
    b = make_batch(B, seed=77, O=O)
    rad = np.tile(np.array([0.1, 0.1, 0.1, 0.1, 0.45, 0.45]), (B, 1))          # 4 static slots, 2 dynamic slots
    ref = oracle_mod.solve(cfg, b["x_cur"], b["goal"], obs=b["obs"], obs_rad=rad)
    uni = oracle_mod.solve(cfg, b["x_cur"], b["goal"], obs=b["obs"])
    assert np.abs(ref.U - uni.U).max() > 1e-3                                      # the radii matter
    conv = ref.status == 0
    assert conv.mean() >= 0.9
    d = np.linalg.norm(ref.X[:, None, :2, 1:] - b["obs"][:, :, :, None], axis=2) - rad[:, :, None]
    assert d[conv].min() >= cfg.inflation - 1e-6
    for i in np.where(conv)[0][:6]:
        cert = kkt_certificate(NLP(cfg, b["x_cur"][i], b["goal"][i], obs=b["obs"][i], obs_rad=rad[i]), ref.X[i], ref.U[i])
        assert cert["primal"] < 1e-8 and cert["obstacle_violation"] < 1e-6 and cert["stationarity_rel"] < 1e-5
    for warp in (False, True):
        X, U, obj, st, it, tp = emul.solve(cfg, b["x_cur"], b["goal"], obs=b["obs"], obs_rad=rad, warp=warp)
        assert (st == ref.status).all()
        assert np.abs(U - ref.U)[conv].max() <= 1e-6
    # dense augmented system agrees too
    rd = oracle_mod.solve(replace(cfg, linsolve="dense"), b["x_cur"][:6], b["goal"][:6], obs=b["obs"][:6], obs_rad=rad[:6])
    assert (rd.status == ref.status[:6]).all() and np.abs(rd.U - ref.U[:6])[rd.status == 0].max() <= 1e-6


# ---------------- the reference's own mpc/agent.py, unmodified, on the drop-in ----------------
class _FakeBatched:
    """Stands in for BatchedMotionPlanner on a GPU-less machine: same ``solve`` signature, answers from the oracle.  Everything
    above it -- kiss_mpc_b200.MotionPlanner.solve's handling of the reference's keyword arguments -- is the real product code."""
    oracle = None
    calls = []

    def __init__(self, config, max_batch=1, device=0, layout="instance_major"):
        self.config = config

    def close(self):
        pass

    def solve(self, x, g, X0, U0, obstacles, obstacle_radius, inflation):
        ok, c = _FakeBatched.oracle, self.config
        O = 0 if obstacles is None else obstacles.shape[1]
        cfg = ok.OracleConfig(N=c.N, T=c.T, linsolve="dense", x_bounds=c.x_bounds, y_bounds=c.y_bounds, v_bounds=c.v_bounds,
                              w_bounds=c.w_bounds, O=O, inflation=inflation,
                              cost_mode=c.cost_mode, goal_range=c.goal_range)
        rad = None if not O else np.broadcast_to(np.asarray(obstacle_radius, float), (1, O))
        r = ok.solve(cfg, x, g, X0=X0, U0=U0, obs=obstacles, obs_rad=rad)
        _FakeBatched.calls.append(dict(O=O, rad=None if rad is None else rad.copy(), inflation=inflation, x=x.copy()))
        from kiss_mpc_b200 import SolveResult
        return SolveResult(r.X, r.U, r.obj, r.status, r.iters)


def load_reference_agent(planner_cls):
    """Import /root/reference/mpc/agent.py as it is, with `mpc.optimizer` and `obstacle_handling.geometry` replaced by stubs that
    need no casadi: MotionPlanner = the drop-in, Circle = the data holder with the reference's (literal) distance formula."""
    class Circle:   # geometry.py:25-44 without the casadi import
        def __init__(self, center, radius):
            self.radius = radius
            self.center = np.array(center, dtype=np.float64)

        @property
        def location(self):
            return tuple(self.center)

        @location.setter
        def location(self, value):
            self.center += np.array(value) - self.center

        def calculate_distance(self, distance_to, custom_self_location=None):
            center = np.array(custom_self_location) if custom_self_location is not None else self.center
            return np.linalg.norm(np.array(distance_to[:2] - center) - self.radius)

    saved = {k: sys.modules.get(k) for k in ("mpc", "mpc.optimizer", "obstacle_handling", "obstacle_handling.geometry", "mpc.agent")}
    try:
        for name in ("mpc", "obstacle_handling"):
            m = types.ModuleType(name); m.__path__ = []
            sys.modules[name] = m
        mo = types.ModuleType("mpc.optimizer"); mo.MotionPlanner = planner_cls
        mg = types.ModuleType("obstacle_handling.geometry"); mg.Circle = Circle
        sys.modules["mpc.optimizer"], sys.modules["obstacle_handling.geometry"] = mo, mg
        spec = importlib.util.spec_from_file_location("mpc.agent", os.path.join(REF, "mpc", "agent.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod, Circle


class _Ob:
    def __init__(self, geometry):
        self.geometry = geometry


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "mpc", "agent.py")), reason="reference tree not present (GPU box)")
def test_reference_agent_unmodified_on_drop_in(oracle_mod, monkeypatch):
    """agent.py:62 constructs the planner, agent.py:139-155 calls it with the reference's keywords and applies the answer --
    through kiss_mpc_b200.MotionPlanner, with static and dynamic obstacles of DIFFERENT radii (0.1 vs the hard-coded 0.3 of
    dynamic_obstacle.py:9), checked against direct oracle solves with one radius per class."""
    import kiss_mpc_b200.planner as P
    _FakeBatched.oracle, _FakeBatched.calls = oracle_mod, []
    monkeypatch.setattr(P, "BatchedMotionPlanner", _FakeBatched)
    mod, Circle = load_reference_agent(P.MotionPlanner)
    ego = mod.EgoAgent(id=1, radius=0.4, initial_position=(0.0, 0.0), initial_orientation=np.pi / 2, planning_time_step=0.1, horizon=12,
                       goal_position=(1.5, 1.0))
    assert isinstance(ego.planner, P.MotionPlanner) and ego.planner.horizon == 12 and ego.planner.time_step == 0.1
    ego.update_goal(np.array([1.5, 1.0, 0.0]))
    stat = [_Ob(Circle((0.8, 0.9), 0.1)), _Ob(Circle((3.0, 3.0), 0.25))]     # second static radius is ignored by the reference
    dyn = [_Ob(Circle((0.5, -1.0), 0.3))]
    X0, U0, x0 = ego.states_matrix.copy(), ego.controls_matrix.copy(), ego.state.copy()
    ego.step(static_obstacles=stat, dynamic_obstacles=dyn)
    call = _FakeBatched.calls[-1]
    assert call["O"] == 3 and np.allclose(call["rad"], [[0.1, 0.1, 0.3]]) and call["inflation"] == pytest.approx(0.5)
    cfg = oracle_mod.OracleConfig(N=12, T=0.1, linsolve="dense", O=3, inflation=0.5)
    ref = oracle_mod.solve(cfg, x0[None], np.array([[1.5, 1.0, 0.0]]), X0=X0[None], U0=U0[None],
                           obs=np.array([[(0.8, 0.9), (3.0, 3.0), (0.5, -1.0)]]), obs_rad=np.array([[0.1, 0.1, 0.3]]))
    assert ego.planner.last_status == int(ref.status[0]) == 0
    assert ego.states_matrix.shape == (3, 13) and ego.controls_matrix.shape == (2, 12)
    assert np.array_equal(ego.states_matrix, ref.X[0]) and np.array_equal(ego.controls_matrix, ref.U[0])
    # agent.py:153-155: hand-off of the state and the applied control
    assert np.allclose(ego.geometry.center, ref.X[0][:2, 1]) and ego.linear_velocity == ref.U[0][0, 0] and ego.angular_velocity == ref.U[0][1, 0]
    # second step: warm start = previous solution unshifted, current state = X[:, 1] (agent.py:70-72, :139-145)
    ego.step(static_obstacles=stat, dynamic_obstacles=dyn)
    assert np.array_equal(_FakeBatched.calls[-1]["x"][0], ref.X[0][:, 1])
    # a failed solve is surfaced, not swallowed (the reference never reads IPOPT's status, optimizer.py:375-400)
    far = mod.EgoAgent(id=2, radius=0.4, initial_position=(25.0, 0.0), initial_orientation=0.0, planning_time_step=0.1, horizon=12,
                       goal_position=(26.0, 0.0))
    far.update_goal(np.array([26.0, 0.0, 0.0]))
    with pytest.warns(RuntimeWarning, match="solver status"):
        far.step()
    assert far.planner.last_status != 0
    far2 = mod.EgoAgent(id=3, radius=0.4, initial_position=(25.0, 0.0), initial_orientation=0.0, planning_time_step=0.1, horizon=12,
                        goal_position=(26.0, 0.0))
    far2.planner.on_failure = "raise"
    with pytest.raises(P.KmpcError):
        far2.step()


# ---------------- occupancy map -> circles (obstacle_handling/static_obstacle.py:12-56) ----------------
def _synthetic_map(seed, h=240, w=320):
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 254, np.uint8)                      # free space
    for _ in range(14):                                       # dark blobs: walls, pillars, clutter (some cut by the border)
        y, x = int(rng.integers(-10, h)), int(rng.integers(-10, w))
        hh, ww = int(rng.integers(3, 60)), int(rng.integers(3, 60))
        img[max(y, 0):y + hh, max(x, 0):x + ww] = int(rng.integers(0, 120))
    img[rng.random((h, w)) < 0.002] = 0                       # speckle
    img[rng.random((h, w)) < 0.01] = 205                      # "unknown" grey of a ROS map: free for the script (> 127)
    return img


def test_map_to_circles_equals_reference_script():
    """kmpc_map_to_circles against the reference script run with its own OpenCV calls (oracle/map_circles_cv2.py): the same circles
    in the same order, on synthetic maps and -- where the reference tree is present -- on its rrc_lab.pgm (27,262 circles)."""
    cv2 = pytest.importorskip("cv2")
    import ctypes as C
    from kiss_mpc_b200 import _lib, map_to_circles, read_pgm
    from oracle.map_circles_cv2 import circles_cv2
    images = [_synthetic_map(s) for s in (1, 2, 3)]
    # (an all-occupied map has no free pixel to measure from: the script's own cv2.circle call rejects the resulting radius)
    images.append(np.full((40, 50), 255, np.uint8))           # all free: no circle
    real = os.path.join(REF, "obstacle_handling", "rrc_lab.pgm")
    if os.path.exists(real):
        img = read_pgm(real)
        assert np.array_equal(img, cv2.imread(real, cv2.IMREAD_GRAYSCALE))
        images.append(img)
    L = _lib.load()
    for img in images:
        cen, rad = map_to_circles(img)
        c2, r2, _ = circles_cv2(img)
        assert len(cen) == len(c2) and np.array_equal(cen, c2) and np.array_equal(rad, r2)
        h, w = img.shape
        d = np.empty((h, w), np.float32)
        assert L.kmpc_map_distance(img.ctypes.data, w, h, 127, d.ctypes.data) == 0
        _, binary = cv2.threshold(img, 127, 255, cv2.THRESH_BINARY)
        assert np.abs(d - cv2.distanceTransform(cv2.bitwise_not(binary), cv2.DIST_L2, 5)).max() <= 1e-5
    # metres: resolution / origin of a map_server YAML, y up; truncation of the list
    img = images[0]
    cen, rad = map_to_circles(img)
    cm, rm = map_to_circles(img, resolution=0.05, origin=(-8.0, -6.0))
    assert np.allclose(cm[:, 0], -8.0 + (cen[:, 0] + 0.5) * 0.05) and np.allclose(cm[:, 1], -6.0 + (img.shape[0] - cen[:, 1] - 0.5) * 0.05)
    assert np.allclose(rm, rad * 0.05)
    c5, r5 = map_to_circles(img, max_circles=5)
    assert np.array_equal(c5, cen[:5]) and np.array_equal(r5, rad[:5])
    assert (np.diff(rad) <= 0).all()                           # largest first
    with pytest.raises(ValueError):
        map_to_circles(np.zeros((2, 3, 4), np.uint8))
