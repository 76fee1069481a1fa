"""CPU-side tests (no GPU): the oracle against its own cross-checks and committed golden vectors, the g++ build of the
per-thread solver source against the oracle, the C ABI surface, host logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from kiss_mpc_b200.synthetic import cfg1_instance, make_batch, make_tracks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


# ---------------- oracle ----------------
def test_oracle_dense_vs_riccati(oracle_mod):
    b = make_batch(64, seed=1002)
    rd = oracle_mod.solve(oracle_mod.OracleConfig(linsolve="dense"), b["x_cur"], b["goal"])
    rr = oracle_mod.solve(oracle_mod.OracleConfig(linsolve="riccati"), b["x_cur"], b["goal"])
    assert (rd.status == 0).all() and (rr.status == 0).all()
    assert np.abs(rd.U - rr.U).max() < 1e-9
    assert (rd.iters == rr.iters).all()


def test_oracle_kkt_certificate_and_slsqp(oracle_mod):
    """Solver-independent check of the oracle's answers: first-order certificate + SLSQP polish."""
    from oracle.nlp_numpy import NLP, kkt_certificate, slsqp_polish
    cfg = oracle_mod.OracleConfig(linsolve="dense")
    b = make_batch(4, seed=3)
    r = oracle_mod.solve(cfg, b["x_cur"], b["goal"])
    for i in range(4):
        nlp = NLP(cfg, b["x_cur"][i], b["goal"][i])
        cert = kkt_certificate(nlp, r.X[i], r.U[i])
        assert cert["primal"] < 1e-8 and cert["bound_violation"] < 1e-7 and cert["stationarity_rel"] < 1e-5
        Xp, Up, fp = slsqp_polish(nlp, r.X[i], r.U[i])
        assert abs(fp - r.obj[i]) <= 1e-6 * abs(r.obj[i])
        assert np.abs(Up - r.U[i]).max() <= 1e-5


def test_oracle_golden(oracle_mod):
    """Golden vectors committed under tests/golden (made by tests/golden/make_golden.py from the oracle + SLSQP polish)."""
    g = np.load(os.path.join(GOLD, "cfg1_and_batch.npz"))
    cfg = oracle_mod.OracleConfig(linsolve="dense")
    r = oracle_mod.solve(cfg, g["x_cur"], g["goal"])
    assert (r.status == g["status"]).all()
    assert np.abs(r.U - g["U"]).max() <= 1e-9
    assert np.abs(r.obj - g["obj"]).max() <= 1e-9 * np.abs(g["obj"]).max()
    # the polished (SLSQP) solutions are an independent solver's answer for the same instances
    assert np.abs(r.U - g["U_slsqp"]).max() <= 1e-5
    assert (np.abs(r.obj - g["obj_slsqp"]) / np.abs(g["obj_slsqp"])).max() <= 1e-6


def test_oracle_stagewise_obstacle_centres(oracle_mod):
    """centres that move with the stage (dynamic_obstacle.py:47-56).  A track of N equal columns IS the static
    problem (same bits); a moving track is checked with the solver-independent certificate of the NumPy NLP."""
    from dataclasses import replace
    from oracle.nlp_numpy import NLP, kkt_certificate
    cfg = oracle_mod.OracleConfig(linsolve="dense", O=4)
    cfs = replace(cfg, obs_stagewise=True)
    b = make_batch(8, seed=1004, O=4)
    still = np.ascontiguousarray(np.repeat(b["obs"][:, :, None, :], cfg.N, axis=2))
    r0 = oracle_mod.solve(cfg, b["x_cur"], b["goal"], obs=b["obs"])
    r1 = oracle_mod.solve(cfs, b["x_cur"], b["goal"], obs=still)
    assert (r0.status == r1.status).all() and (r0.iters == r1.iters).all()
    np.testing.assert_array_equal(r0.U, r1.U)
    tr = make_tracks(b["obs"], cfg.N, seed=5)
    r2 = oracle_mod.solve(cfs, b["x_cur"], b["goal"], obs=tr)
    assert (r2.status == 0).all()
    assert np.abs(r2.U - r0.U).max() > 1e-6          # the motion matters
    for i in range(8):
        nlp = NLP(cfs, b["x_cur"][i], b["goal"][i], obs=tr[i])
        cert = kkt_certificate(nlp, r2.X[i], r2.U[i], inflation=cfg.inflation)
        assert cert["primal"] < 1e-8 and cert["obstacle_violation"] < 1e-7 and cert["stationarity_rel"] < 1e-5
        d = np.linalg.norm(r2.X[i][None, :2, 1:] - tr[i].transpose(0, 2, 1), axis=1) - cfg.obs_radius
        assert d.min() >= cfg.inflation - 1e-7


def test_obstacle_predictor_restatement():
    """oracle/obstacle_predictor.py (dynamic_obstacle.py:20-37): straight line when omega = 0, heading through deg2rad as
    the reference writes it, the current state in column 0."""
    from oracle.obstacle_predictor import predict_track, predict_tracks
    tr = predict_track((1.0, 2.0, np.deg2rad(90)), 1.0, 0.0, 6)           # the constructor defaults (:8)
    a = np.deg2rad(np.deg2rad(90))                                          # radians fed to deg2rad (:24-25)
    np.testing.assert_allclose(tr[0], 1.0 + 0.1 * np.cos(a) * np.arange(6), rtol=0, atol=1e-15)
    np.testing.assert_allclose(tr[1], 2.0 + 0.1 * np.sin(a) * np.arange(6), rtol=0, atol=1e-15)
    assert (tr[2] == np.deg2rad(90)).all()
    tr = predict_track((0.0, 0.0, 0.5), 2.0, 1.0, 4, literal=False)
    assert tr[:, 0].tolist() == [0.0, 0.0, 0.5]
    np.testing.assert_allclose(tr[:, 1], [0.2 * np.cos(0.5), 0.2 * np.sin(0.5), 0.6])
    np.testing.assert_allclose(tr[:, 2], [tr[0, 1] + 0.2 * np.cos(0.6), tr[1, 1] + 0.2 * np.sin(0.6), 0.7])
    assert predict_tracks(np.zeros((3, 3)), np.ones(3), np.zeros(3), 5).shape == (3, 5, 2)


# ---------------- per-thread solver source (g++ build) vs oracle ----------------
@pytest.mark.parametrize("case", ["box", "N50", "literal", "obs", "tracks", "layout1", "infeasible"])
def test_solver_source_matches_oracle(oracle_mod, case):
    import emul
    kw, B, seed, O, layout = {}, 96, 1002, 0, 0
    if case == "N50":
        kw, seed = dict(N=50), 1003
    elif case == "literal":
        kw = dict(cost_mode="code_literal", goal_range="code", y_bounds=(-oracle_mod.INF, oracle_mod.INF))
    elif case == "obs":
        kw, seed, O = dict(O=10), 1004, 10
    elif case == "tracks":
        kw, seed, O = dict(O=6, obs_stagewise=True), 1004, 6
    elif case == "layout1":
        layout = 1
    elif case == "infeasible":
        kw = dict(max_iter=300)
    cfg = oracle_mod.OracleConfig(linsolve="riccati", **kw)
    b = make_batch(B, seed=seed, O=O)
    if case == "infeasible":
        b["x_cur"][::4, 0] = 25.0
    if case == "tracks":
        b["obs"] = make_tracks(b["obs"], cfg.N, seed=9)
    ref = oracle_mod.solve(cfg, b["x_cur"], b["goal"], obs=b["obs"])
    X, U, obj, st, it, tp = emul.solve(cfg, b["x_cur"], b["goal"], obs=b["obs"], layout=layout)
    assert (st == ref.status).all()
    conv = st == 0
    assert np.abs(U - ref.U)[conv].max() <= 1e-9
    assert (it == ref.iters).mean() >= 0.95


@pytest.mark.parametrize("case", ["box", "N50", "literal", "obs", "tracks", "tracksN50", "infeasible", "warm"])
def test_warp_solver_source_matches_oracle(oracle_mod, case):
    """The warp-per-instance kernel source (kmpc_warp.cuh: what the GPU runs for N <= 63) on the 32-fibre warp emulator
    (tests/host_emul/simt.h, one emulated warp = one block) against the oracle."""
    import emul
    kw, B, seed, O = {}, 24, 1002, 0
    if case == "N50":
        kw, seed, B = dict(N=50), 1003, 12
    elif case == "literal":
        kw = dict(cost_mode="code_literal", goal_range="code", y_bounds=(-oracle_mod.INF, oracle_mod.INF))
    elif case == "obs":
        kw, seed, O = dict(O=10), 1004, 10
    elif case == "tracks":
        kw, seed, O = dict(O=6, obs_stagewise=True), 1004, 6
    elif case == "tracksN50":
        kw, seed, O, B = dict(N=50, O=3, obs_stagewise=True), 1004, 3, 8
    elif case == "infeasible":
        kw = dict(max_iter=300)
    cfg = oracle_mod.OracleConfig(linsolve="riccati", **kw)
    b = make_batch(B, seed=seed, O=O)
    if case.startswith("tracks"):
        b["obs"] = make_tracks(b["obs"], cfg.N, seed=9)
    X0 = U0 = None
    x = b["x_cur"]
    if case == "infeasible":
        x[::4, 0] = 25.0
    if case == "warm":
        r0 = oracle_mod.solve(cfg, x, b["goal"])
        X0, U0, x = r0.X, r0.U, r0.X[:, :, 1].copy()
    ref = oracle_mod.solve(cfg, x, b["goal"], X0=X0, U0=U0, obs=b["obs"])
    X, U, obj, st, it, tp = emul.solve(cfg, x, b["goal"], X0=X0, U0=U0, obs=b["obs"], warp=True)
    assert (st == ref.status).all()
    conv = st == 0
    assert np.abs(U - ref.U)[conv].max() <= (1e-9 if O == 0 else 1e-6)
    assert (np.abs(obj - ref.obj) / np.abs(ref.obj))[conv].max() <= 1e-9
    assert (it == ref.iters).mean() >= (0.95 if O == 0 else 0.5)


@pytest.mark.parametrize("case,W,chunk", [("box", 16, 40), ("box", 16, 2), ("box", 8, 1), ("N50", 12, 3), ("literal", 6, 1), ("obs", 10, 24),
                                          ("warm", 16, 4), ("infeasible", 16, 40)])
def test_warp_block_emulation_bit_identical(oracle_mod, case, W, chunk):
    """The block-level machinery of the warp kernel on the multi-warp fibre emulator (tests/host_emul/simt.h): blocks of W warps
    pulling `chunk` instances from their queue -- the Riccati warp serving every instance of the block, the matrix-only inertia
    candidates, refills during the serial window and the TAIL MODE (at most a quarter of the slots taken: every live instance
    borrows free slots for full-solve candidates of the next perturbations) -- must return bit for bit what one instance alone on a
    one-warp block returns: scheduling and speculation never change the arithmetic of an instance."""
    import emul
    kw, B, seed, O = {}, 40, 1002, 0
    if case == "N50":
        kw, seed, B = dict(N=50), 1003, 12
    elif case == "literal":
        kw, B = dict(cost_mode="code_literal", goal_range="code", y_bounds=(-oracle_mod.INF, oracle_mod.INF)), 12
    elif case == "obs":
        kw, seed, O, B = dict(O=4), 1004, 4, 24
    elif case == "infeasible":
        kw = dict(max_iter=300)
    cfg = oracle_mod.OracleConfig(linsolve="riccati", **kw)
    b = make_batch(B, seed=seed, O=O)
    X0 = U0 = None
    x = b["x_cur"]
    if case == "infeasible":
        x[::4, 0] = 25.0
    if case == "warm":
        r0 = oracle_mod.solve(cfg, x, b["goal"])
        X0, U0, x = r0.X, r0.U, r0.X[:, :, 1].copy()
    one = emul.solve(cfg, x, b["goal"], X0=X0, U0=U0, obs=b["obs"], warp=True)
    blk = emul.solve(cfg, x, b["goal"], X0=X0, U0=U0, obs=b["obs"], warp=True, warps=W, chunk=chunk)
    for a, c in zip(one[:5], blk[:5]):
        assert np.array_equal(a, c)
    trips_one, trips_blk = int(one[5].sum()), int(blk[5][0])
    assert trips_blk <= trips_one          # speculation only ever saves retry trips
    if case == "box" and chunk <= 2:
        assert trips_blk < 0.9 * trips_one  # tail mode: the inertia retries are gone


@pytest.mark.parametrize("kw", [dict(), dict(O=2), dict(cost_mode="code_literal", goal_range="code", y_bounds=(-1e20, 1e20))])
def test_non_finite_inputs_give_invalid_number(oracle_mod, kw):
    """NaN / inf in the current state, the goal or an obstacle centre: IPOPT's Invalid_Number_Detected (-13) for that instance
    (1e300 diverges, 4), nothing hangs, the neighbours are solved as usual -- oracle, thread solver and warp solver alike."""
    import emul
    cfg = oracle_mod.OracleConfig(linsolve="riccati", max_iter=200, **kw)
    O = kw.get("O", 0)
    b = make_batch(8, seed=5, O=O)
    b["x_cur"][1, 0] = np.nan; b["goal"][2, 1] = np.inf; b["x_cur"][3, 2] = 1e300; b["goal"][4, 0] = -np.inf
    b["x_cur"][5, 2] = np.nan; b["goal"][6, 2] = np.nan
    want = [0, -13, -13, 4, -13, -13, -13, 0]
    if O:
        b["obs"][7, 0, 0] = np.nan; want[7] = -13
    ref = oracle_mod.solve(cfg, b["x_cur"], b["goal"], obs=b["obs"])
    assert ref.status.tolist() == want
    for warp in (False, True):
        X, U, obj, st, it, tp = emul.solve(cfg, b["x_cur"], b["goal"], obs=b["obs"], warp=warp)
        assert st.tolist() == want and np.abs(U[0] - ref.U[0]).max() <= 1e-9


def test_solver_source_warm_start(oracle_mod):
    import emul
    cfg = oracle_mod.OracleConfig(linsolve="riccati")
    b = make_batch(48, seed=11)
    r0 = oracle_mod.solve(cfg, b["x_cur"], b["goal"])
    r1 = oracle_mod.solve(cfg, r0.X[:, :, 1], b["goal"], X0=r0.X, U0=r0.U)
    X, U, obj, st, it, tp = emul.solve(cfg, r0.X[:, :, 1], b["goal"], X0=r0.X, U0=r0.U)
    assert (st == r1.status).all() and np.abs(U - r1.U).max() <= 1e-9


# ---------------- C ABI surface ----------------
def test_abi_exports_every_declared_symbol():
    from kiss_mpc_b200 import _lib, build
    so = build.build()
    hdr = open(os.path.join(ROOT, "include", "kmpc.h")).read()
    declared = set(re.findall(r"\b(kmpc_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"kmpc_config", "kmpc_handle", "kmpc_stats"}
    assert declared == set(_lib.SYMBOLS)
    L = C.CDLL(so)
    for s in declared:
        assert hasattr(L, s), s
    assert L.kmpc_version() == _lib.KMPC_VERSION


def test_abi_struct_and_workspace_size():
    from kiss_mpc_b200 import PlannerConfig, _lib
    L = _lib.load()
    c = PlannerConfig().to_c(65536, 0, 0)
    assert C.sizeof(c) == 10 * 4 + 8 * (1 + 3 + 3 + 4 + 4 + 1)
    nbytes = L.kmpc_workspace_bytes(C.byref(c))
    assert 100e6 < nbytes < 4e9
    bad = PlannerConfig(N=0).to_c(1, 0, 0)
    assert L.kmpc_workspace_bytes(C.byref(bad)) == 0


def test_abi_argument_errors_without_gpu():
    """Error behaviour of the C ABI that needs no device: bad configurations are rejected before any CUDA call, a missing
    device is reported as KMPC_E_NODEVICE, nothing throws across the boundary."""
    import torch
    from kiss_mpc_b200 import PlannerConfig, _lib
    L = _lib.load()
    h = C.c_void_p()
    bad = PlannerConfig(N=0).to_c(4, 0, 0)
    assert L.kmpc_create(C.byref(bad), C.byref(h)) == -1 and not h.value            # KMPC_E_BADARG
    assert b"invalid configuration" in L.kmpc_last_error(None)
    bad2 = PlannerConfig(v_bounds=(0.5, -0.2)).to_c(4, 0, 0)                         # lo >= hi
    assert L.kmpc_create(C.byref(bad2), C.byref(h)) == -1
    assert L.kmpc_create(C.byref(bad2), None) == -1
    if not torch.cuda.is_available():
        ok = PlannerConfig().to_c(4, 0, 0)
        assert L.kmpc_create(C.byref(ok), C.byref(h)) == -4 and not h.value         # KMPC_E_NODEVICE: no CPU path
    assert L.kmpc_solve(None, 1, None, None, None, None, None, 0, 0.0, None, 0.0, None, None, None, None, None, None) == -1
    assert L.kmpc_solve_host_into(None, 1, None, None, None, None, None, 0, 0.0, None, 0.0, None, None, None, None, None) == -1
    assert L.kmpc_host_sync(None) == -1 and L.kmpc_enable_peer(None, 0) == -1
    assert L.kmpc_shared_buffer_create(None, 16, None, None) == -1 and L.kmpc_shared_buffer_open(None, None, None) == -1
    assert L.kmpc_pinned_alloc(0, None) == -1 and L.kmpc_pinned_free(None) == 0
    L.kmpc_destroy(None)                                                              # no-op


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kiss_mpc_b200 import BatchedMotionPlanner, KmpcError
    with pytest.raises(KmpcError):
        BatchedMotionPlanner(max_batch=4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "kiss_mpc_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "kmpc_oracle" not in src, f


def test_sensor_filter_restatement():
    """oracle/sensor_filter.py against the reference's semantics (environment.py:48-65): sorted by distance, within the
    sensor radius, dict-keyed ties keep the later obstacle, literal distance of geometry.py:44."""
    from oracle.sensor_filter import circle_distance, sensor_filter
    st = np.array([0.0, 0.0, 0.3])
    cen = np.array([[3.0, 0.0], [1.0, 1.0], [1.0, 1.0], [10.0, 0.0], [0.0, -2.0]]); rad = np.array([0.5, 0.3, 0.3, 0.5, 0.2])
    assert abs(circle_distance(st, cen[0], 0.5) - np.hypot(-3.5, -0.5)) < 1e-15          # radius subtracted from both components
    assert abs(circle_distance(st, cen[0], 0.5, literal=False) - 2.5) < 1e-15
    assert sensor_filter(st, cen, rad, 5.0) == [4, 2, 0]                                 # tie 1/2 -> 2; 3 is out of range; literal: 1.81 < 1.84
    assert sensor_filter(st, cen, rad, 5.0, literal=False) == [2, 4, 0]


# ---------------- host logic ----------------
class _Geo:  # duck-typed obstacle_handling.geometry.Circle
    def __init__(self, c, r): self.center, self.radius = np.array(c, float), r


class _Obs:
    def __init__(self, c, r=0.3): self.geometry = _Geo(c, r)


def test_model_shim_bookkeeping(oracle_mod):
    """the `Model` the ROS node drives (ros2interface.py:19, :28-38, :55-60, :93-107, :172-174), stepped on CPU
    with an oracle-backed planner: hand-off x <- X[:,1] (agent.py:70-72), published control U[:,0] (agent.py:154-155), unshifted
    warm start (agent.py:139-145), sensor filter order (environment.py:48-65), waypoint advance (environment.py:77-80),
    odometry override + reset (ros2interface.py:93-107)."""
    from kiss_mpc_b200.model import Model, literal_circle_distance
    from oracle_planner import OraclePlanner
    pl = OraclePlanner(oracle_mod, 0.8, 7)
    near, far, tie = _Obs((2.0, 1.0)), _Obs((40.0, 40.0)), _Obs((2.0, 1.0))
    m = Model(id=1, initial_position=(0, 0), initial_orientation=np.deg2rad(90), horizon=7, use_warm_start=True,
              planning_time_step=0.8, linear_velocity_bounds=(-0.3, 0.3), angular_velocity_bounds=(-0.3, 0.3), waypoints=[],
              static_obstacles=[far, near, tie], planner=pl)
    assert m.states_matrix.shape == (3, 8) and (m.states_matrix == m.initial_state[:, None]).all() and m.current_waypoint() is None
    m.waypoints = np.array([(0.3, 0.9, 1.0), (1.5, 2.5, 0.0)]); m.waypoint_index = 0          # ros2interface.py:172-174
    m.update_goal(m.current_waypoint())
    assert (m.goal_state == [0.3, 0.9, 1.0]).all()
    prev_X = m.states_matrix.copy()
    m.step()
    c = pl.calls[-1]
    assert (c["current_state"] == prev_X[:, 1]).all() and c["n_static"] == 1 and c["first_static"] is tie   # out of range dropped; equal distances collapse to the later obstacle
    assert m.linear_velocity == m.controls_matrix[0, 0] and m.angular_velocity == m.controls_matrix[1, 0]
    assert abs(m.linear_velocity) <= 0.3 + 1e-8 and abs(m.angular_velocity) <= 0.3 + 1e-8
    assert (m.center == m.states_matrix[:2, 1]).all()
    idx = []
    for _ in range(40):
        X_before = m.states_matrix.copy()
        m.step()
        assert (pl.calls[-1]["current_state"] == X_before[:, 1]).all()       # perfect-model hand-off
        idx.append(m.waypoint_index)
        if m.final_goal_reached:
            break
    assert idx[0] == 0 or idx[0] == 1
    assert m.waypoint_index == 1 and m.final_goal_reached and (m.goal_state == [1.5, 2.5, 0.0]).all()
    assert literal_circle_distance(m.center, m.radius, m.goal_state) <= 0.5
    assert all(c["status"] == 0 for c in pl.calls)
    # odometry callback: new initial_state, matrices reset to it (ros2interface.py:93-107)
    m.initial_state = np.array([1.0, 2.0, 0.5]); m.reset(matrices_only=True)
    assert (m.states_matrix == np.array([1.0, 2.0, 0.5])[:, None]).all() and (m.controls_matrix == 0).all()
    lv = m.linear_velocity
    m.reset()
    assert m.linear_velocity == 0.0 and (lv != 0.0 or True)
    m.state_override = True
    m.initial_state = np.array([1.1, 2.1, 0.4])
    m.step()
    assert (pl.calls[-1]["current_state"] == [1.1, 2.1, 0.4]).all() and (m.center == [1.1, 2.1]).all()


def test_shard_range_partitions():
    from kiss_mpc_b200 import shard_range
    for B in (0, 1, 7, 64, 65536, 65537):
        for w in (1, 2, 3, 8):
            parts = [shard_range(B, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == B
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))


def _gloo_worker(rank, world, port, B, q):
    import torch
    import torch.distributed as dist
    from kiss_mpc_b200 import SolveResult, gather_results, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(B, rank, world)
    idx = torch.arange(lo, hi, dtype=torch.float64)
    local = SolveResult(idx[:, None, None].repeat(1, 3, 4), idx[:, None, None].repeat(1, 2, 3), idx * 2, idx.to(torch.int32), idx.to(torch.int32) + 1)
    out = gather_results(local, B)
    if rank == 0:
        q.put((out.objective.tolist(), out.status.tolist(), tuple(out.states.shape)))
    dist.destroy_process_group()


def test_sharded_gather_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    B, world, port = 11, 2, 29731
    ps = [ctx.Process(target=_gloo_worker, args=(r, world, port, B, q)) for r in range(world)]
    [p.start() for p in ps]
    obj, st, shp = q.get(timeout=120)
    [p.join(60) for p in ps]
    assert obj == [2.0 * i for i in range(B)] and st == list(range(B)) and shp == (B, 3, 4)
