"""GPU tests of the SURVEY 8f rows built in round 2, through the C ABI:
row 2 - the control-step epilogue (mppi_step_epilogue: env.step, collision_check, get_top_samples by radix select),
        against vectors recorded from the live reference (tests/golden/epilogue_*.npz) and the oracle;
row 4 - the device map rasteriser (mppi_raster_map) against the reference's own grids;
plus occupancy grids larger than shared memory (global-memory instantiation of the solve kernel) and top samples
merged across sample shards."""
import json
import os

import numpy as np
import pytest
import torch

from engine_util import ParityStats, assert_parity, build_engine, build_oracle
from oracle import fixtures as fx
from oracle import mppi_oracle as mo
from test_oracle_epilogue_maps import navigation_obstacle_raster, racing_lane_raster, racing_obstacle_raster

pytestmark = pytest.mark.gpu


def _expected_order(costs: torch.Tensor, n: int):
    """The select's contract: the n smallest (cost, sample id) pairs = a stable ascending sort of the costs."""
    c = costs.detach().cpu()
    order = torch.sort(c, stable=True).indices[:n]
    return order.numpy(), c[order].numpy()


@pytest.mark.parametrize("name", ["epilogue_racing", "epilogue_navigation2d"])
def test_step_epilogue_matches_reference_control_loop(name):
    """forward -> env.step(action_seq[0]) -> env.collision_check(state_seq) -> get_top_samples, as
    example/racing.py:229-237 / example/navigation2d.py:36-44 run them, on the reference's recorded loop."""
    z = np.load(os.path.join(fx.GOLDEN_DIR, f"{name}.npz"))
    cfg = json.loads(str(z["cfg"]))
    model, solver = build_engine(cfg)
    T, ds = cfg["horizon"], model.dim_state
    goal, thr = z["goal"], float(z["goal_threshold"])
    n_top = z["top_w"].shape[1]
    omodel = fx.oracle_model(cfg["model"])
    obstacle = omodel.obstacle if cfg["model"] == "racing" else omodel.grid
    for s in range(len(z["state"])):
        if "refpath" in z.files:
            model.reference_path_tensor = torch.from_numpy(z["refpath"][s])
        state = torch.from_numpy(z["state"][s])
        action, states = solver.forward(state, noise=torch.from_numpy(z["noise"][s]))
        nxt, reached, coll, (traj, w) = solver.step_epilogue(action, states, state=state, goal=goal,
                                                             goal_threshold=thr, top_n=n_top)
        assert solver._lib.mppi_last_epilogue_launches(solver._h) == 2  # K <= 8192: step + flags + select | re-roll
        # the engine's action_seq differs from the reference's by the parity tolerance; the step itself is exact
        onxt, oreached = mo.env_step(omodel, state, action[0].cpu(), goal, thr)
        np.testing.assert_allclose(nxt.cpu().numpy(), onxt.numpy(), rtol=0, atol=2e-6)
        np.testing.assert_allclose(nxt.cpu().numpy(), z["next_state"][s], rtol=0, atol=5e-4)
        assert bool(reached) == oreached == bool(z["is_goal"][s])
        assert tuple(coll.shape) == (1, T + 1)
        ocoll = mo.collision_check(obstacle, states.cpu())  # exact on the engine's own predicted trajectory
        np.testing.assert_array_equal(coll.cpu().numpy(), ocoll.numpy())
        # and the reference's flags (its trajectory differs by the parity tolerance: at most a cell-edge flip)
        assert int((coll.cpu().numpy() != z["collisions"][s]).sum()) <= 1
        # top samples: same winners as the reference where its weights are distinct and non-zero
        w = w.cpu().numpy()
        assert np.all(np.diff(w) <= 0)
        np.testing.assert_allclose(w, z["top_w"][s], rtol=5e-3, atol=1e-7)
        ref_w = z["top_w"][s]
        gaps = np.abs(np.diff(ref_w)) > 1e-2 * ref_w[:-1]
        stable = np.concatenate([[True], gaps]) & np.concatenate([gaps, [True]]) & (ref_w > 1e-30)
        np.testing.assert_allclose(traj.cpu().numpy()[stable], z["top_traj"][s][stable], rtol=0, atol=5e-3)
        # and exactly what get_top_samples returns
        t2, w2 = solver.get_top_samples(n_top)
        assert torch.equal(t2, traj) and torch.equal(w2, torch.from_numpy(w).to(w2.device))
        solver._previous_action_seq = torch.from_numpy(z["action_seq"][s])


@pytest.mark.parametrize("name", ["epilogue_racing", "epilogue_navigation2d"])
def test_step_epilogue_probes_goal_flags_clamps_and_border(name):
    """Recorded probes of env.step (states around the goal, actions beyond the env bounds) and of collision_check
    (positions over obstacles and beyond the map border) - reference outputs, bit for bit on the flags."""
    z = np.load(os.path.join(fx.GOLDEN_DIR, f"{name}.npz"))
    cfg = json.loads(str(z["cfg"]))
    model, solver = build_engine(cfg)
    T, ds, du = cfg["horizon"], model.dim_state, model.dim_control
    if "refpath" in z.files:
        model.reference_path_tensor = torch.from_numpy(z["refpath"][0])
    solver.forward(torch.from_numpy(z["state"][0]))  # the epilogue belongs to a solve
    goal, thr = z["goal"], float(z["goal_threshold"])
    probes = z["coll_probe_in"][0]
    rows = 0
    for i, (st, act) in enumerate(zip(z["probe_state"], z["probe_action"])):
        seq = torch.zeros(T + 1, ds)
        chunk = probes[rows: rows + T + 1]
        seq[: len(chunk)] = torch.from_numpy(chunk)
        actions = torch.zeros(T, du)
        actions[0] = torch.from_numpy(act)
        nxt, reached, coll, top = solver.step_epilogue(actions, seq.view(1, T + 1, ds), state=torch.from_numpy(st),
                                                       goal=goal, goal_threshold=thr)
        assert top is None
        np.testing.assert_allclose(nxt.cpu().numpy(), z["probe_next"][i], rtol=0, atol=2e-6)
        assert bool(reached) == bool(z["probe_goal"][i]), (i, st, z["probe_next"][i])
        np.testing.assert_array_equal(coll.cpu().numpy()[0, : len(chunk)], z["coll_probe_out"][0, rows: rows + len(chunk)])
        rows = (rows + T + 1) % max(1, len(probes) - T - 1)
    assert z["coll_probe_out"].sum() > 0


TOP_CASES = [
    (dict(model="racing", horizon=80, num_samples=65536, sigmas=[0.5, 0.1], lambda_=1.0, use_sg_filter=True), 300, 3),
    (dict(model="racing", horizon=25, num_samples=4000, sigmas=[0.5, 0.1], lambda_=1.0), 1024, 2),
    (dict(model="navigation2d", horizon=30, num_samples=70001, sigmas=[0.5, 0.5], lambda_="ESSPS"), 300, 3),
    (dict(model="cartpole", horizon=50, num_samples=1048576, u_min=[-3.0], u_max=[3.0], sigmas=[1.0], lambda_=0.001,
          state0=[0.0, 0.0, 0.05, 0.0]), 300, 4),
    (dict(model="pendulum", horizon=20, num_samples=200, u_min=[-2.0], u_max=[2.0], sigmas=[1.0], lambda_=1.0,
          state0=[3.0, 0.0]), 200, 2),
    (dict(model="mountaincar", horizon=40, num_samples=300, u_min=[-1.0], u_max=[1.0], sigmas=[1.0], lambda_=0.1,
          state0=[-0.5, 0.0]), 1, 2),
]


@pytest.mark.parametrize("cfg,n,launches", TOP_CASES, ids=lambda v: f"{v['model']}-K{v['num_samples']}" if isinstance(v, dict) else str(v))
def test_top_select_equals_stable_sort_of_all_costs(cfg, n, launches):
    """The radix select returns exactly the first n entries of a stable ascending sort of the K costs (ids and
    costs bit for bit - ties by the lower sample id), the winners' weights are softmax(-c / lambda) and their
    trajectories are what rolling their controls gives; launches = select levels + select/step + re-roll."""
    import ctypes as C

    from mppi_playground_b200 import _capi

    model, solver = build_engine(cfg)
    K, T = cfg["num_samples"], cfg["horizon"]
    state = torch.tensor(cfg["state0"]) if "state0" in cfg else (
        fx.load_env_racing() if cfg["model"] == "racing" else fx.load_env_navigation2d()).start_state.clone()
    if cfg["model"] == "racing":
        import mppi_playground_b200 as eng

        env = fx.load_env_racing()
        model.reference_path_tensor, _ = eng.racing_reference_path(state, env.center_path, 0, T, v_max=env.v_max)
    noise = solver.sampler_noise() if K <= 70001 else None
    solver.forward(state)
    costs = solver._costs
    want_ids, want_costs = _expected_order(costs, n)
    a = _capi.MppiStepEpilogue()
    traj = torch.empty(n, T + 1, model.dim_state, device="cuda")
    w = torch.empty(n, device="cuda")
    tc = torch.empty(n, device="cuda")
    ti = torch.empty(n, device="cuda", dtype=torch.int32)
    a.top_n, a.d_top_traj, a.d_top_w, a.d_top_cost, a.d_top_id = n, traj.data_ptr(), w.data_ptr(), tc.data_ptr(), ti.data_ptr()
    _capi.check(solver._lib.mppi_step_epilogue(solver._h, C.byref(a), None))
    torch.cuda.synchronize()
    assert solver._lib.mppi_last_epilogue_launches(solver._h) == launches
    np.testing.assert_array_equal(ti.cpu().numpy(), want_ids)
    np.testing.assert_array_equal(tc.cpu().numpy(), want_costs)
    weights = solver._weights
    np.testing.assert_allclose(w.cpu().numpy(), weights[torch.from_numpy(want_ids).cuda().long()].cpu().numpy(),
                               rtol=1e-6, atol=0)
    # candidates API (the per-rank half of a sharded get_top_samples) agrees
    cc, ci = solver.top_candidates(n)
    assert torch.equal(ci, ti) and torch.equal(cc, tc)
    # trajectories: roll the winners' clamped controls through the engine's own rollout entry point
    if noise is not None:
        prev = torch.zeros(T, model.dim_control, device="cuda")  # first solve: zero warm start
        u = torch.clamp(prev + noise[torch.from_numpy(want_ids).cuda().long()], solver._u_min, solver._u_max).contiguous()
        rolled = torch.empty(n, T + 1, model.dim_state, device="cuda")
        _capi.check(solver._lib.mppi_rollout_actions(solver._h, solver._device_state(state).data_ptr(), u.data_ptr(), n,
                                                     rolled.data_ptr(), None))
        torch.cuda.synchronize()
        assert torch.equal(rolled, traj)  # (mountaincar: both entry points store the in-place-mutated states)
    # get_top_samples goes through the same path
    t2, w2 = solver.get_top_samples(n)
    assert torch.equal(t2, traj) and torch.equal(w2, w)


def test_top_select_tie_break_is_by_sample_id():
    """Cartpole's costs tie exactly (bang-bang force: many samples share a trajectory): the select keeps the lowest
    ids among equal costs, like the stable sort."""
    cfg = dict(model="cartpole", horizon=6, num_samples=5000, u_min=[-3.0], u_max=[3.0], sigmas=[1.0], lambda_=0.001,
               state0=[0.0, 0.0, 0.05, 0.0])
    model, solver = build_engine(cfg)
    solver.forward(torch.tensor(cfg["state0"]))
    costs = solver._costs.cpu()
    assert costs.unique().numel() <= 64  # 2^6 force patterns at most
    for n in (1, 7, 300, 1024):
        want_ids, want_costs = _expected_order(costs, n)
        cc, ci = solver.top_candidates(n)
        np.testing.assert_array_equal(ci.cpu().numpy(), want_ids)
        np.testing.assert_array_equal(cc.cpu().numpy(), want_costs)


def test_large_n_falls_back_to_the_full_sort():
    cfg = dict(model="pendulum", horizon=20, num_samples=4096, u_min=[-2.0], u_max=[2.0], sigmas=[1.0], lambda_=1.0)
    model, solver = build_engine(cfg)
    solver.forward(torch.tensor([3.0, 0.0]))
    traj, w = solver.get_top_samples(2000)
    t1, w1 = solver.get_top_samples(1024)
    assert torch.equal(traj[:1024], t1) and torch.equal(w[:1024], w1)
    assert bool((w[:-1] >= w[1:]).all())


@pytest.mark.parametrize("cfg", [
    dict(model="racing", horizon=40, num_samples=6000, sigmas=[0.5, 0.1], lambda_=1.0, use_sg_filter=True),
    dict(model="navigation2d", horizon=30, num_samples=3001, sigmas=[0.5, 0.5], lambda_="ESSPS"),
], ids=["racing-fixed", "navigation2d-ESSPS"])
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("inject", [False, True], ids=["sampler", "injected"])
def test_sharded_top_samples_equal_single_solver(cfg, world, inject):
    """get_top_samples on a sample-sharded solver (VERDICT r1 missing #4): every shard's n best candidates are
    gathered, each shard merges them and re-rolls the global winners by global sample id - same trajectories and
    weights as the unsharded solver (here the shards share one GPU and the gather is a concatenation)."""
    import mppi_playground_b200 as eng
    from mppi_playground_b200.mppi import solve_shards_inprocess, top_samples_inprocess

    model, single = build_engine(cfg)
    shards = [build_engine(cfg, shard=(r, world)) for r in range(world)]
    state = (fx.load_env_racing() if cfg["model"] == "racing" else fx.load_env_navigation2d()).start_state.clone()
    if cfg["model"] == "racing":
        env = fx.load_env_racing()
        ref, _ = eng.racing_reference_path(state, env.center_path, 0, cfg["horizon"], v_max=env.v_max)
        for m in [model] + [m for m, _ in shards]:
            m.reference_path_tensor = ref
    noise = single.sampler_noise() if inject else None
    single.forward(state, noise=noise)
    solve_shards_inprocess([sv for _, sv in shards], state, noise=noise)
    n = 300
    traj, w = single.get_top_samples(n)
    outs = top_samples_inprocess([sv for _, sv in shards], n)
    for t, ww in outs:
        assert torch.equal(t, traj)  # same sampler keys / noise rows, same arithmetic
        np.testing.assert_allclose(ww.cpu().numpy(), w.cpu().numpy(), rtol=2e-5, atol=1e-12)  # S: summation order


def _fixture_grids():
    env, nav = fx.load_env_racing(), fx.load_env_navigation2d()
    return env, nav


def test_device_rasteriser_reproduces_the_reference_grids():
    """mppi_raster_map against the grids the reference painted (18 602 / 445 529 / 5 019 occupied cells)."""
    env, nav = _fixture_grids()
    model, solver = build_engine(dict(model="racing", horizon=25, num_samples=1024, sigmas=[0.5, 0.1], lambda_=1.0))
    g_obs = solver.rasterise_map(0, racing_obstacle_raster(), want_grid=True)
    g_lane = solver.rasterise_map(1, racing_lane_raster(), want_grid=True)
    np.testing.assert_array_equal(g_obs.cpu().numpy(), env.obstacle)
    np.testing.assert_array_equal(g_lane.cpu().numpy(), env.lane)
    assert int(g_obs.sum()) == 18602 and int(g_lane.sum()) == 445529
    nmodel, nsolver = build_engine(dict(model="navigation2d", horizon=30, num_samples=512, sigmas=[0.5, 0.5], lambda_=1.0))
    g_nav = nsolver.rasterise_map(0, navigation_obstacle_raster(), want_grid=True)
    np.testing.assert_array_equal(g_nav.cpu().numpy(), nav.obstacle)
    assert int(g_nav.sum()) == 5019


def test_solve_on_rasterised_maps_equals_solve_on_uploaded_maps():
    """The rasteriser writes the packed layout the kernels read: a solve on device-built maps is bit-identical to
    one on the uploaded fp32 grids, the bounded fast path stays enabled, and collision flags see the new map."""
    import ctypes as C

    import mppi_playground_b200 as eng
    from mppi_playground_b200 import maps

    env = fx.load_env_racing()
    cfg = dict(model="racing", horizon=80, num_samples=8192, sigmas=[0.5, 0.1], lambda_=1.0, use_sg_filter=True)
    m1, uploaded = build_engine(cfg)
    m2, rastered = build_engine(cfg)
    ref, _ = eng.racing_reference_path(env.start_state, env.center_path, 0, 80, v_max=env.v_max)
    m1.reference_path_tensor = m2.reference_path_tensor = ref
    a1, s1 = uploaded.forward(env.start_state)
    rastered2 = rastered
    rastered2.rasterise_map(0, racing_obstacle_raster())
    rastered2.rasterise_map(1, racing_lane_raster())
    a2, s2 = rastered2.forward(env.start_state)
    assert torch.equal(a1, a2) and torch.equal(s1, s2) and torch.equal(uploaded._costs, rastered2._costs)
    fast, mism, flags = C.c_int32(), C.c_uint64(), C.c_int32()
    rastered2._lib.mppi_map_info(rastered2._h, 0, C.byref(fast), C.byref(mism), C.byref(flags))
    assert fast.value == 1 and (flags.value & 4), "bounded fast path must survive a re-rasterised map"
    # dynamic obstacle: a disc dropped onto the predicted trajectory shows up in the flags and in the costs
    r = racing_obstacle_raster()
    x, y = s2[0, 10, 0].item(), s2[0, 10, 1].item()
    r.add_circle_obstacle(np.array([x, y]), 1.0)
    rastered2.rasterise_map(0, r)
    _, _, coll, _ = rastered2.step_epilogue(a2, s2, state=env.start_state)
    assert coll[0, 10].item() == 1.0
    _, _, coll0, _ = uploaded.step_epilogue(a1, s1, state=env.start_state)
    assert coll0[0, 10].item() == 0.0
    a3, s3 = rastered2.forward(env.start_state)  # and the next solve runs on the new map
    assert torch.isfinite(a3).all() and torch.isfinite(s3).all()
    assert float(rastered2._costs.max()) >= 10000.0


def test_maps_larger_than_shared_memory_use_the_global_path():
    """VERDICT r1 missing #5: the reference's lookup has no size limit (obstacle_map_2d.py:168-200). Two
    2000 x 2000 grids (2 x 500 kB packed) cannot be staged; the global-memory instantiation must match the oracle
    at the usual bars, and the staged geometry of a normal racing solver must be untouched."""
    import mppi_playground_b200 as eng

    rng = np.random.default_rng(3)
    W = 2000
    cell, origin, lim = 0.05, (1000, 1000), (-50.0, 50.0, -50.0, 50.0)
    obstacle = np.zeros((W, W), dtype=np.float32)
    lane = np.zeros((W, W), dtype=np.float32)
    for _ in range(400):
        cx, cy, r = rng.integers(0, W), rng.integers(0, W), rng.integers(5, 40)
        obstacle[max(cx - r, 0): cx + r, max(cy - r, 0): cy + r] = 1.0
    lane[:, : W // 2 - 300] = 1.0
    lane[:, W // 2 + 300:] = 1.0
    env = fx.load_env_racing()
    q = env.Q
    model = eng.RacingModel(obstacle, lane, cell_size=(cell, cell), origin=(origin, origin), u_min=env.u_min,
                            u_max=env.u_max, wheelbase=env.wheelbase, v_max=env.v_max, lim=lim, Qc=q[0], Ql=q[1],
                            Qv=q[2], Qo=q[3], Qin=q[4], Qdin=q[5])
    solver = eng.MPPI(horizon=40, num_samples=4096, dim_state=4, dim_control=2, dynamics=model.dynamics,
                      cost_func=model.cost_func, u_min=model.u_min, u_max=model.u_max, sigmas=torch.tensor([0.5, 0.1]),
                      lambda_=1.0)
    omodel = mo.RacingModel(mo.GridMap(torch.from_numpy(obstacle), cell, origin), mo.GridMap(torch.from_numpy(lane), cell, origin),
                            u_min=env.u_min, u_max=env.u_max, wheelbase=env.wheelbase, v_max=env.v_max, lim=lim,
                            Qc=q[0], Ql=q[1], Qv=q[2], Qo=q[3], Qin=q[4], Qdin=q[5])
    oracle = mo.OracleMPPI(horizon=40, num_samples=4096, dim_state=4, dim_control=2, dynamics=omodel.dynamics,
                           cost_func=omodel.cost, u_min=env.u_min, u_max=env.u_max, sigmas=[0.5, 0.1], lambda_=1.0,
                           burn_constructor_draw=False)
    state = torch.tensor([0.5, -3.0, 0.3, 2.0])
    ref = torch.zeros(41, 4)
    ref[:, 0] = torch.linspace(0.5, 12.0, 41)
    ref[:, 1] = -3.0
    ref[:, 3] = 4.0
    model.reference_path_tensor, omodel.reference_path = ref, ref
    for s in range(2):
        noise = solver.sampler_noise().cpu()
        action, states = solver.forward(state)
        info = solver.launch_info()
        assert info["smem_bytes"] < 60000, info  # nothing staged
        tr = oracle.forward(state, noise=noise)
        st = ParityStats(solver._costs.cpu().numpy(), tr.costs.numpy(), action.cpu().numpy(), tr.action_seq.numpy(),
                         states.cpu().numpy(), tr.state_seq.numpy(), 1.0, 1.0)
        assert_parity(st)
        assert (tr.costs > 9000).float().mean() > 0.01  # obstacles are actually hit
        oracle.prev_action_seq = action.cpu().clone()
        state = states[0, 1].cpu().clone()
    # a normal solver keeps the staged, paired geometry
    _, normal = build_engine(dict(model="racing", horizon=80, num_samples=65536, sigmas=[0.5, 0.1], lambda_=1.0))
    assert normal.launch_info()["smem_bytes"] > 150000


def test_peer_barrier_between_inprocess_shards():
    """mppi_p2p_barrier: every shard flags every peer's mailbox and waits for all flags in its own; shards driven
    by one process on separate streams pass it together, repeatedly (double-buffered by the sequence parity), and
    the fused solves keep working afterwards."""
    from mppi_playground_b200.mppi import connect_shards_inprocess, solve_fused_shards_inprocess

    cfg = dict(model="cartpole", horizon=20, num_samples=4096, u_min=[-3.0], u_max=[3.0], sigmas=[1.0], lambda_=0.001)
    world = 3
    shards = [build_engine(cfg, shard=(r, world))[1] for r in range(world)]
    connect_shards_inprocess(shards)
    streams = [torch.cuda.Stream() for _ in range(world)]
    state = torch.tensor([0.0, 0.0, 0.05, 0.0])
    for _ in range(5):
        for sv, st in zip(shards, streams):
            with torch.cuda.stream(st):
                sv.peer_barrier()
        outs = solve_fused_shards_inprocess(shards, state, streams)
    for st in streams:
        st.synchronize()
    for sv in shards:
        sv.check_exchange()
    assert all(torch.equal(o[0], outs[0][0]) for o in outs)
    _, single = build_engine(cfg)
    for _ in range(5):
        a1, _ = single.forward(state)
    np.testing.assert_allclose(outs[0][0].cpu().numpy(), a1.cpu().numpy(), rtol=2e-5, atol=2e-6)

