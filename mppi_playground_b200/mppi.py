"""``MPPI`` - drop-in for ``pi_mpc.mppi.MPPI`` of kohonda/mppi_playground.

Same constructor keywords, defaults, ``forward`` / ``__call__`` / ``reset`` /
``get_top_samples`` / ``get_samples_from_posterior`` surface and error
behaviour as the reference class (src/pi_mpc/mppi.py:16-620); the solve itself
is one launch of the sm_100a rollout kernel in libmppi_b200.so, reached
through the C ABI in include/mppi_b200.h. torch is used for device memory and
streams only.

Differences a caller can observe (all documented in DESIGN.md):
  * no CPU fallback: the reference silently drops to CPU when CUDA is missing
    (mppi.py:102-105); this class raises instead;
  * ``dynamics`` / ``cost_func`` must resolve to a built-in device model
    (``mppi_playground_b200.models.resolve``); ``info`` is accepted and
    ignored (the built-in costs take what they need from the model object);
  * the sampler is the engine's counter-based Philox, not torch's global
    generator, so native-mode noise differs from the reference's draw by
    draw; ``forward(state, noise=...)`` injects a reference noise tensor;
  * ``_state_seq_batch`` is not materialised; ``get_top_samples`` re-rolls.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _capi, models


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class MPPI(nn.Module):
    """Model Predictive Path Integral solver (Williams et al., T-RO 2017) on B200."""

    def __init__(
        self,
        horizon: int,
        num_samples: int,
        dim_state: int,
        dim_control: int,
        dynamics: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
        cost_func: Callable[[torch.Tensor, torch.Tensor, Dict], torch.Tensor],
        u_min: torch.Tensor,
        u_max: torch.Tensor,
        sigmas: torch.Tensor,
        lambda_,
        lbps_delta: float = 0.01,
        essps_target_ess: Optional[float] = None,
        lambda_min: float = 0.01,
        lambda_max: float = 10.0,
        exploration: float = 0.0,
        use_sg_filter: bool = False,
        sg_window_size: int = 5,
        sg_poly_order: int = 3,
        device=torch.device("cuda"),
        dtype=torch.float32,
        seed: int = 42,
        *,
        shard: Optional[Tuple[int, int]] = None,
        process_group=None,
        block_size: int = 0,
        fused_exchange: bool = True,
        samples_per_thread: int = 0,
    ) -> None:
        """Arguments up to ``seed`` are the reference's (mppi.py:24-47).

        Keyword-only extensions: ``shard=(rank, world)`` / ``process_group``
        split the K samples over one process per GPU (``num_samples`` stays the
        GLOBAL count); ``block_size`` overrides the launch geometry; ``fused_exchange`` (default) wires the
        ranks' mailboxes over CUDA IPC so that a fixed-lambda / MPO sharded solve is one kernel launch per GPU
        with the partials exchanged by peer stores over NVLink (False: NCCL all-gather + a finish kernel);
        ``samples_per_thread`` (0: engine picks from the sample count; 1 / 2: force the single / paired-sample
        loop of the racing and navigation2d rollouts - tests and tuning).
        """
        super().__init__()
        u_min, u_max, sigmas = (torch.as_tensor(x) for x in (u_min, u_max, sigmas))
        assert u_min.shape == (dim_control,)  # mppi.py:96-98
        assert u_max.shape == (dim_control,)
        assert sigmas.shape == (dim_control,)
        if not torch.cuda.is_available():
            raise RuntimeError("mppi_playground_b200.MPPI needs a CUDA device (B200, sm_100a); there is no CPU "
                               "fallback - the reference's own pi_mpc.MPPI is the CPU implementation")
        if dtype != torch.float32:
            raise NotImplementedError("the device models are fp32, like the reference's default dtype")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(f"device={device}: the engine only runs on CUDA devices")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self._device, self._dtype = device, dtype
        self._lib = _capi.load()

        self._horizon, self._num_samples = int(horizon), int(num_samples)
        self._dim_state, self._dim_control = int(dim_state), int(dim_control)
        self._dynamics, self._cost_func = dynamics, cost_func
        self._u_min = u_min.clone().detach().to(device, dtype)  # mppi.py:115-117
        self._u_max = u_max.clone().detach().to(device, dtype)
        self._sigmas = sigmas.clone().detach().to(device, dtype)
        self._exploration = exploration
        self._use_sg_filter, self._sg_window_size, self._sg_poly_order = use_sg_filter, sg_window_size, sg_poly_order
        self._lbps_delta = lbps_delta
        self._lambda_min, self._lambda_max = lambda_min, lambda_max
        self._lambda_init = lambda_

        # sample sharding (one process per GPU)
        self._pg = process_group
        if shard is None and process_group is not None:
            import torch.distributed as dist

            shard = (dist.get_rank(process_group), dist.get_world_size(process_group))
        self._rank, self._world = shard if shard is not None else (0, 1)

        hs = host_setup(horizon=horizon, num_samples=num_samples, dim_state=dim_state, dim_control=dim_control,
                        dynamics=dynamics, cost_func=cost_func, u_min=u_min, u_max=u_max, sigmas=sigmas,
                        lambda_=lambda_, lbps_delta=lbps_delta, essps_target_ess=essps_target_ess,
                        lambda_min=lambda_min, lambda_max=lambda_max, exploration=exploration,
                        use_sg_filter=use_sg_filter, sg_window_size=sg_window_size, sg_poly_order=sg_poly_order,
                        seed=seed, device_index=device.index, shard=(self._rank, self._world),
                        block_size=block_size, samples_per_thread=samples_per_thread)
        self._auto_lambda, self._binding, self._coeffs = hs.auto_lambda, hs.binding, hs.coeffs
        self._shard_lo, self._local_samples = hs.shard_lo, hs.local_samples
        self._essps_target_ess = hs.essps_target_ess
        self._params_cache = hs.params
        cfg = hs.cfg
        h = C.c_void_p()
        with torch.cuda.device(device):
            _capi.check(self._lib.mppi_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._map_identity = None
        self._bind_maps(required=False)
        self._noise_keepalive = None
        self._gathered = None
        self._fused = False
        if self._world > 1 and process_group is not None and fused_exchange and self._auto_lambda in (None, "MPO"):
            self._connect_peers(process_group)

    def _connect_peers(self, group) -> None:
        """Exchange the CUDA IPC handles of the shard mailboxes (64 bytes per rank, once) and connect."""
        import torch.distributed as dist

        mine = (C.c_uint8 * 64)()
        with torch.cuda.device(self._device):
            _capi.check(self._lib.mppi_p2p_export(self._h, mine))
        handles = [None] * self._world
        dist.all_gather_object(handles, bytes(mine), group=group)
        flat = (C.c_uint8 * (64 * self._world)).from_buffer_copy(b"".join(handles))
        with torch.cuda.device(self._device):
            _capi.check(self._lib.mppi_p2p_connect(self._h, flat, self._world, self._rank))
        self._fused = True

    # ------------------------------------------------------------------ plumbing
    def __del__(self):
        d = self.__dict__  # plain dict access: nn.Module.__setattr__ is unusable during interpreter shutdown
        h, lib = d.get("_h"), d.get("_lib")
        d["_h"] = None
        if h and lib is not None:
            try:
                lib.mppi_destroy(h)
            except Exception:
                pass

    def _bind_maps(self, required: bool) -> None:
        """(Re)upload the occupancy grids when the model's map objects appear or change
        (example/racing.py:227 sets them after the solver is constructed)."""
        ident_fn = getattr(self._binding, "map_identity", None)
        ident = ident_fn() if ident_fn else "static"
        if ident == self._map_identity:
            return
        try:
            specs = self._binding.maps()
        except ValueError:
            if required:
                raise
            return
        for slot, (grid, cell, ox, oy) in enumerate(specs):
            g = grid.detach().to(torch.float32).contiguous()
            on_dev = int(g.is_cuda)
            if g.is_cuda and g.device != self._device:
                g, on_dev = g.to(self._device), 1
            # the device lookup is a bit per cell (the reference's grids are exactly 0/1:
            # obstacle_map_2d.py:79,123,159, lane_map_2d.py:83); a grid carrying other values would silently
            # lose them in the packing, so it is refused
            if not bool(((g == 0) | (g == 1)).all()):
                raise ValueError(f"occupancy grid of map slot {slot} holds values other than 0 and 1; the device "
                                 "lookup is bit-packed and cannot represent it")
            if g.is_cuda:  # mppi_set_map packs on the legacy default stream: the producer stream must be done
                torch.cuda.current_stream(g.device).synchronize()
            with torch.cuda.device(self._device):
                _capi.check(self._lib.mppi_set_map(self._h, slot, g.data_ptr(), on_dev, g.shape[0], g.shape[1],
                                                   float(cell), float(ox), float(oy)))
        self._map_identity = ident

    def rasterise_map(self, slot: int, raster, want_grid: bool = False) -> Optional[torch.Tensor]:
        """Build the occupancy grid of map ``slot`` on the device from a shape list (``maps.ObstacleRaster`` /
        ``maps.LaneRaster``) - the device twin of ObstacleMap.add_*_obstacle / LaneMap.populate_map
        (obstacle_map_2d.py:103-160, lane_map_2d.py:68-88) - and bind it to the solver without an fp32 grid
        round trip. ``want_grid`` also returns the reference's ``_map_torch`` ([W,H] fp32 0/1)."""
        from .maps import flatten_shapes

        discs, rects = flatten_shapes(raster)
        grid = torch.empty(raster.width, raster.height, device=self._device) if want_grid else None
        i32p = C.POINTER(C.c_int32)
        with torch.cuda.device(self._device):
            _capi.check(self._lib.mppi_raster_map(
                self._h, slot, raster.mode, raster.width, raster.height, float(raster.cell_size),
                float(raster.origin[0]), float(raster.origin[1]), discs.ctypes.data_as(i32p), len(discs),
                rects.ctypes.data_as(i32p), len(rects), _ptr(grid)))
        # the grids now live in the engine: stop mirroring the binding's map objects for this solver
        ident_fn = getattr(self._binding, "map_identity", None)
        self._map_identity = ident_fn() if ident_fn else "static"
        return grid

    def _refresh_params(self) -> None:
        p = self._binding.params(strict=True)
        if p != self._params_cache:
            arr = (C.c_float * len(p))(*p)
            _capi.check(self._lib.mppi_set_model_params(self._h, arr, len(p)))
            self._params_cache = list(p)

    def _device_state(self, state) -> torch.Tensor:
        assert tuple(state.shape) == (self._dim_state,)  # mppi.py:247
        if not torch.is_tensor(state):
            return torch.tensor(np.asarray(state), device=self._device, dtype=self._dtype)  # mppi.py:249-250
        return state.detach().to(self._device, self._dtype).contiguous()

    def _device_refpath(self) -> Optional[torch.Tensor]:
        if self._binding.model_id != _capi.MODEL_RACING:
            return None
        ref = self._binding.reference_path()
        if ref is None:
            raise ValueError("reference path, obstacle map, and lane map must be set before calling solve method.")
        ref = torch.as_tensor(ref)
        if tuple(ref.shape) != (self._horizon + 1, 4):
            raise ValueError(f"reference_path must be [{self._horizon + 1}, 4], got {tuple(ref.shape)}")
        return ref.detach().to(self._device, torch.float32).contiguous()

    # ------------------------------------------------------------------ solve
    def forward(self, state, info: Dict = {}, *, noise: Optional[torch.Tensor] = None):
        """One MPPI solve (mppi.py:223-460).

        Returns ``(action_seq [T, du], state_seq [1, T+1, ds])`` on the solver's device.
        ``noise`` ([K,T,du], sigma already applied - what the reference keeps in
        ``_action_noises``) replaces the in-kernel sampler for parity tests.
        """
        st = self._device_state(state)
        self._bind_maps(required=True)
        self._refresh_params()
        ref = self._device_refpath()
        T, du, ds = self._horizon, self._dim_control, self._dim_state
        if noise is not None:
            noise = torch.as_tensor(noise).detach().to(self._device, torch.float32).contiguous()
            lo = self._shard_lo
            self.__dict__["_noise_full"] = None
            if tuple(noise.shape) == (self._num_samples, T, du) and self._world > 1:
                self.__dict__["_noise_full"] = noise  # get_top_samples re-rolls other ranks' winners from it
                noise = noise[lo:lo + self._local_samples].contiguous()
            assert tuple(noise.shape) == (self._local_samples, T, du)
        else:
            self.__dict__["_noise_full"] = None
        self.__dict__["_noise_keepalive"] = (noise, st, ref)  # plain dict write: nn.Module.__setattr__ is slow
        action = torch.empty(T, du, device=self._device, dtype=torch.float32)
        states = torch.empty(T + 1, ds, device=self._device, dtype=torch.float32)
        s = _stream_ptr(self._device)
        if self._fused:
            # a peer that never arrived leaves NaN outputs and sets a flag in mapped host memory: report the
            # failure of an EARLIER solve before launching the next one (no synchronisation needed to read it)
            self.check_exchange(synchronize=False)
        if self._world == 1 or self._fused:
            _capi.check(self._lib.mppi_solve(self._h, st.data_ptr(), _ptr(ref), _ptr(noise), action.data_ptr(),
                                             states.data_ptr(), s))
        else:
            self._sharded_solve(st, ref, noise, action, states, s)
        return action, states.view(1, T + 1, ds)

    def peer_barrier(self) -> None:
        """Device-side barrier across the ranks of a fused sharded solver on the current stream (``mppi_p2p_barrier``):
        all GPUs pass it within one NVLink latency of each other."""
        _capi.check(self._lib.mppi_p2p_barrier(self._h, _stream_ptr(self._device)))

    def check_exchange(self, synchronize: bool = True) -> None:
        """Raise if the fused peer exchange of a sharded solve timed out (a rank's kernel was not running within
        ~2 s of the others: the outputs of that solve are NaN on this rank and the carried state was not updated).
        ``synchronize=True`` waits for the solves launched so far first; ``forward`` calls it without waiting, so a
        failure surfaces at the latest on the next solve."""
        if not self._fused:
            return
        if synchronize:
            torch.cuda.current_stream(self._device).synchronize()
        flag = C.c_int32()
        _capi.check(self._lib.mppi_p2p_status(self._h, C.byref(flag)))
        if flag.value:
            raise RuntimeError(f"fused shard exchange timed out in solve #{flag.value} on rank {self._rank}: a peer's "
                               "solve kernel was not running concurrently (outputs of that solve are NaN)")

    def _sharded_solve(self, st, ref, noise, action, states, s) -> None:
        """K split over ranks: roll the shard, exchange the shard partials
        ((8 + T*du) floats per rank), finish redundantly on every rank."""
        import torch.distributed as dist

        self._shard_stage_rollout(st, ref, noise, s)
        if self._auto_lambda in ("LBPS", "ESSPS"):
            all_costs = gather_shards(self._wrap_costs(), self._num_samples, self._world, self._pg)
            self._shard_stage_lambda(all_costs, s)
        part = self._wrap_partial()
        gathered = torch.empty(self._world * part.numel(), device=self._device, dtype=torch.float32)
        dist.all_gather_into_tensor(gathered, part, group=self._pg)
        self._shard_stage_finish(gathered, st, action, states, s)

    # the three stages of a sharded solve (mppi_shard_* of the C ABI); also driven in-process by
    # solve_shards_inprocess() when one process owns several shards
    def _shard_stage_rollout(self, st, ref, noise, s) -> None:
        _capi.check(self._lib.mppi_shard_rollout(self._h, st.data_ptr(), _ptr(ref), _ptr(noise), s))

    def _shard_stage_lambda(self, all_costs: torch.Tensor, s) -> None:
        self._gathered_costs = all_costs
        _capi.check(self._lib.mppi_shard_lambda(self._h, all_costs.data_ptr(), s))

    def _shard_stage_finish(self, gathered: torch.Tensor, st, action, states, s) -> None:
        self._gathered = gathered
        n = gathered.numel() // self._lib.mppi_partial_floats(self._h)
        _capi.check(self._lib.mppi_shard_finish(self._h, gathered.data_ptr(), n, st.data_ptr(), action.data_ptr(),
                                                states.data_ptr(), s))

    def _wrap(self, ptr: int, n: int) -> torch.Tensor:
        return wrap_device_memory(ptr, n, self._device)

    def _wrap_costs(self) -> torch.Tensor:
        p = C.c_void_p()
        _capi.check(self._lib.mppi_costs_ptr(self._h, C.byref(p)))
        return self._wrap(p.value, self._local_samples)

    def _wrap_partial(self) -> torch.Tensor:
        p = C.c_void_p()
        _capi.check(self._lib.mppi_partial_ptr(self._h, C.byref(p)))
        return self._wrap(p.value, self._lib.mppi_partial_floats(self._h))

    def solve_host(self, state: np.ndarray, reference_path: Optional[np.ndarray] = None):
        """End-to-end solve with HOST buffers through ``mppi_solve_host`` (H2D of the
        inputs, the solve, D2H of both outputs, synchronised). numpy in, numpy out."""
        self._bind_maps(required=True)
        self._refresh_params()
        state = np.ascontiguousarray(state, dtype=np.float32)
        assert state.shape == (self._dim_state,)
        ref_p = None
        if self._binding.model_id == _capi.MODEL_RACING:
            if reference_path is None:
                r = self._binding.reference_path()
                reference_path = None if r is None else torch.as_tensor(r).detach().cpu().numpy()
            if reference_path is None:
                raise ValueError("reference path must be set before calling solve method.")
            reference_path = np.ascontiguousarray(reference_path, dtype=np.float32)
            assert reference_path.shape == (self._horizon + 1, 4)
            ref_p = reference_path.ctypes.data
        action = np.empty((self._horizon, self._dim_control), dtype=np.float32)
        states = np.empty((1, self._horizon + 1, self._dim_state), dtype=np.float32)
        _capi.check(self._lib.mppi_solve_host(self._h, state.ctypes.data, ref_p, action.ctypes.data,
                                              states.ctypes.data))
        return action, states

    def reset(self):
        """Zero the warm start and the SG history (mppi.py:212-221)."""
        _capi.check(self._lib.mppi_reset(self._h, _stream_ptr(self._device)))

    # ------------------------------------------------------------------ inspection
    def get_top_samples(self, num_samples: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Highest-weight samples of the last solve, weight-descending (mppi.py:462-487):
        ``([n, T+1, ds], [n])``. Re-rolled on demand instead of stored."""
        assert num_samples <= self._num_samples
        if self._world > 1:
            _, _, _, top = self.step_epilogue(top_n=num_samples)
            return top
        traj = torch.empty(num_samples, self._horizon + 1, self._dim_state, device=self._device)
        w = torch.empty(num_samples, device=self._device)
        _capi.check(self._lib.mppi_top_samples(self._h, num_samples, traj.data_ptr(), w.data_ptr(),
                                               _stream_ptr(self._device)))
        return traj, w

    def top_candidates(self, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """This rank's ``n`` best samples of the last solve as (cost [n] ascending, GLOBAL sample id [n] int32):
        the per-rank half of ``get_top_samples`` on a sharded solver (padding: cost inf, id -1)."""
        cost = torch.empty(n, device=self._device, dtype=torch.float32)
        ids = torch.empty(n, device=self._device, dtype=torch.int32)
        _capi.check(self._lib.mppi_top_candidates(self._h, n, cost.data_ptr(), ids.data_ptr(),
                                                  _stream_ptr(self._device)))
        return cost, ids

    def _gather_candidates(self, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """One all-gather of every rank's n best (cost, id) pairs (2n words per rank)."""
        cost, ids = self.top_candidates(n)
        return gather_candidates(cost, ids, self._world, self._pg)

    def step_epilogue(self, action_seq: Optional[torch.Tensor] = None, state_seq: Optional[torch.Tensor] = None,
                      state=None, goal=None, goal_threshold: float = 0.0, top_n: int = 0, candidates=None):
        """The control-step epilogue of the reference's loops (example/racing.py:233-237,
        example/navigation2d.py:39-44) in one launch: ``env.step(action_seq[0])`` (racing_env.py:142-163),
        ``env.collision_check(state_seq)`` (:374-384) and ``solver.get_top_samples(top_n)`` (mppi.py:462-487).

        Returns ``(next_state [ds], is_goal_reached (0-dim bool), is_collisions [1, T+1], (top_traj, top_w))``;
        groups that were not asked for are ``None``. ``state`` defaults to the state of the last solve;
        ``goal`` = (x, y) with ``goal_threshold`` as in the env's ``step``. On a sharded solver the top samples
        are merged across ranks (one all-gather of each rank's ``top_n`` best); ``candidates`` = (costs, ids)
        replaces that gather (shards driven by one process)."""
        T, ds = self._horizon, self._dim_state
        a = _capi.MppiStepEpilogue()
        keep = []
        want_step = action_seq is not None
        if want_step:
            act = action_seq.detach().to(self._device, torch.float32).contiguous()
            assert tuple(act.shape) == (T, self._dim_control)
            nxt = torch.empty(ds, device=self._device)
            a.d_action_seq, a.d_next_state = act.data_ptr(), nxt.data_ptr()
            keep += [act]
            if state is not None:
                st = self._device_state(state)
                a.d_state = st.data_ptr()
                keep += [st]
            flags = torch.empty(T + 2, device=self._device)
            if state_seq is not None:
                seq = state_seq.detach().to(self._device, torch.float32).reshape(T + 1, ds).contiguous()
                a.d_state_seq, a.d_flags = seq.data_ptr(), flags.data_ptr()
                keep += [seq]
            elif goal is not None:
                seq = torch.zeros(T + 1, ds, device=self._device)
                a.d_state_seq, a.d_flags = seq.data_ptr(), flags.data_ptr()
                keep += [seq]
            if goal is not None:
                a.goal_x, a.goal_y, a.goal_threshold = float(goal[0]), float(goal[1]), float(goal_threshold)
        else:
            assert state_seq is None and goal is None, "collision flags / goal test come with the env step"
        top = None
        if top_n:
            assert top_n <= self._num_samples  # mppi.py:476
            if top_n > _capi.TOP_MAX:
                raise ValueError(f"step_epilogue selects at most {_capi.TOP_MAX} top samples; use get_top_samples")
            traj = torch.empty(top_n, T + 1, ds, device=self._device)
            w = torch.empty(top_n, device=self._device)
            a.top_n, a.d_top_traj, a.d_top_w = top_n, traj.data_ptr(), w.data_ptr()
            if candidates is None and self._world > 1:
                if self._pg is None:
                    raise ValueError("a sharded solver without a process group needs candidates=(costs, ids)")
                candidates = self._gather_candidates(top_n)
            if candidates is not None:
                cc = candidates[0].detach().to(self._device, torch.float32).contiguous()
                ci = candidates[1].detach().to(self._device, torch.int32).contiguous()
                assert cc.numel() == ci.numel()
                a.d_cand_cost, a.d_cand_id, a.n_cand = cc.data_ptr(), ci.data_ptr(), cc.numel()
                keep += [cc, ci]
                full = self.__dict__.get("_noise_full")
                if full is not None:
                    a.d_noise_global = full.data_ptr()
            top = (traj, w)
        _capi.check(self._lib.mppi_step_epilogue(self._h, C.byref(a), _stream_ptr(self._device)))
        self.__dict__["_epilogue_keepalive"] = keep
        nxt_out = nxt if want_step else None
        reached = flags[0] > 0.5 if (want_step and a.d_flags) else None
        coll = flags[1:].view(1, T + 1) if (want_step and state_seq is not None) else None
        return nxt_out, reached, coll, top

    def get_samples_from_posterior(self, optimal_solution: torch.Tensor, state: torch.Tensor, num_samples: int):
        """mppi.py:489-506: sample action sequences from N(optimal_solution, diag sigma^2)
        and roll them out. Returns ``(samples [n,T,du], states [n,T+1,ds])``."""
        assert num_samples <= self._num_samples
        st = self._device_state(state)
        mean = optimal_solution.detach().to(self._device, torch.float32)
        eps = torch.randn(num_samples, self._horizon, self._dim_control, device=self._device)
        samples = (mean + eps * self._sigmas).contiguous()
        traj = torch.empty(num_samples, self._horizon + 1, self._dim_state, device=self._device)
        _capi.check(self._lib.mppi_rollout_actions(self._h, st.data_ptr(), samples.data_ptr(), num_samples,
                                                   traj.data_ptr(), _stream_ptr(self._device)))
        return samples, traj

    @property
    def _weights(self) -> torch.Tensor:
        """softmax(-costs / lambda) of the last solve (mppi.py:376), this rank's shard."""
        w = torch.empty(self._local_samples, device=self._device)
        _capi.check(self._lib.mppi_weights(self._h, w.data_ptr(), _stream_ptr(self._device)))
        return w

    @property
    def _costs(self) -> torch.Tensor:
        return self._wrap_costs().clone()

    def _lambdas(self) -> Tuple[float, float]:
        used, nxt = C.c_double(), C.c_double()
        _capi.check(self._lib.mppi_get_lambda(self._h, C.byref(used), C.byref(nxt), _stream_ptr(self._device)))
        return used.value, nxt.value

    @property
    def _lambda(self) -> float:
        """Temperature the next solve starts from (the reference's ``_lambda`` after ``forward``)."""
        if self._auto_lambda is None:
            return self._lambda_init
        used, nxt = self._lambdas()
        return nxt if self._auto_lambda == "MPO" else used

    @property
    def _previous_action_seq(self) -> torch.Tensor:
        out = torch.empty(self._horizon, self._dim_control, device=self._device)
        _capi.check(self._lib.mppi_get_carry(self._h, out.data_ptr(), None, _stream_ptr(self._device)))
        return out

    @_previous_action_seq.setter
    def _previous_action_seq(self, value: torch.Tensor) -> None:
        v = value.detach().to(self._device, torch.float32).contiguous()
        assert tuple(v.shape) == (self._horizon, self._dim_control)
        _capi.check(self._lib.mppi_set_carry(self._h, v.data_ptr(), None, _stream_ptr(self._device)))
        torch.cuda.current_stream(self._device).synchronize()

    @property
    def _actions_history_for_sg(self) -> torch.Tensor:
        out = torch.empty(self._horizon - 1, self._dim_control, device=self._device)
        _capi.check(self._lib.mppi_get_carry(self._h, None, out.data_ptr(), _stream_ptr(self._device)))
        return out

    @_actions_history_for_sg.setter
    def _actions_history_for_sg(self, value: torch.Tensor) -> None:
        v = value.detach().to(self._device, torch.float32).contiguous()
        assert tuple(v.shape) == (self._horizon - 1, self._dim_control)
        _capi.check(self._lib.mppi_set_carry(self._h, None, v.data_ptr(), _stream_ptr(self._device)))
        torch.cuda.current_stream(self._device).synchronize()

    def sampler_noise(self, solve_index: Optional[int] = None) -> torch.Tensor:
        """sigma*eps the in-kernel sampler uses for solve ``solve_index`` (default: the
        next solve), [K_local, T, du] - so a test can hand the oracle the same noise."""
        idx = self._lib.mppi_solve_index(self._h) if solve_index is None else solve_index
        out = torch.empty(self._local_samples, self._horizon, self._dim_control, device=self._device)
        _capi.check(self._lib.mppi_sample_noise(self._h, idx, out.data_ptr(), _stream_ptr(self._device)))
        return out

    def launch_info(self) -> Dict[str, int]:
        g, b, s = C.c_int32(), C.c_int32(), C.c_int32()
        _capi.check(self._lib.mppi_launch_info(self._h, C.byref(g), C.byref(b), C.byref(s)))
        return {"grid": g.value, "block": b.value, "smem_bytes": s.value,
                "launches_last_solve": self._lib.mppi_last_launch_count(self._h)}

    def kernel_timing(self, enable: bool) -> None:
        _capi.check(self._lib.mppi_kernel_timing(self._h, int(enable)))

    def kernel_time_ms(self) -> Tuple[float, int]:
        ms, n = C.c_double(), C.c_int32()
        _capi.check(self._lib.mppi_kernel_time_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ------------------------------------------------------------------ reference helpers kept verbatim in behaviour
    def _savitzky_golay_coeffs(self, window_size: int, poly_order: int) -> torch.Tensor:
        return savitzky_golay_coeffs(window_size, poly_order)


class HostSetup:
    """Everything ``MPPI.__init__`` derives on the host before the first CUDA call."""

    __slots__ = ("auto_lambda", "binding", "coeffs", "shard_lo", "local_samples", "essps_target_ess", "params", "cfg")


def host_setup(*, horizon, num_samples, dim_state, dim_control, dynamics, cost_func, u_min, u_max, sigmas, lambda_,
               lbps_delta=0.01, essps_target_ess=None, lambda_min=0.01, lambda_max=10.0, exploration=0.0,
               use_sg_filter=False, sg_window_size=5, sg_poly_order=3, seed=42, device_index=0, shard=(0, 1),
               block_size=0, samples_per_thread=0) -> HostSetup:
    """The host half of ``MPPI.__init__`` (mppi.py:24-210): lambda-mode dispatch, callable -> device-model
    resolution, the ``MppiConfig`` of the C ABI. Needs no GPU, so the CPU tests drive it with the very objects
    example/*.py build (tests/test_reference_live.py) - including the solver-before-attributes order of
    example/racing.py:24-58."""
    hs = HostSetup()
    # lambda mode dispatch, mppi.py:183-210
    if lambda_ == "MPO":
        hs.auto_lambda, mode, lam0 = "MPO", _capi.LAMBDA_MPO, 1.0
    elif lambda_ == "LBPS":
        hs.auto_lambda, mode, lam0 = "LBPS", _capi.LAMBDA_LBPS, 1.0
    elif lambda_ == "ESSPS":
        hs.auto_lambda, mode, lam0 = "ESSPS", _capi.LAMBDA_ESSPS, 1.0
    elif isinstance(lambda_, float):
        hs.auto_lambda, mode, lam0 = None, _capi.LAMBDA_FIXED, lambda_
    else:
        raise ValueError("lambda_ must be 'MPO', 'LBPS', 'ESSPS', or a float value.")
    hs.binding = models.resolve(dynamics, cost_func, int(dim_state), int(dim_control))
    rank, world = shard
    lo, hi = shard_bounds(int(num_samples), world, rank)
    hs.shard_lo, hs.local_samples = lo, hi - lo
    hs.essps_target_ess = essps_target_ess if essps_target_ess is not None else num_samples / 10

    cfg = _capi.MppiConfig()
    cfg.abi_version = _capi.ABI_VERSION
    cfg.model = hs.binding.model_id
    cfg.horizon, cfg.num_samples = int(horizon), hs.local_samples
    cfg.dim_state, cfg.dim_control = int(dim_state), int(dim_control)
    for d in range(int(dim_control)):
        cfg.u_min[d], cfg.u_max[d], cfg.sigmas[d] = float(u_min[d]), float(u_max[d]), float(sigmas[d])
    cfg.lambda_mode, cfg.lambda_ = mode, float(lam0)
    cfg.lbps_delta, cfg.essps_target_ess = float(lbps_delta), float(hs.essps_target_ess)
    cfg.lambda_min, cfg.lambda_max = float(lambda_min), float(lambda_max)
    cfg.exploration = float(exploration)
    cfg.use_sg_filter, cfg.sg_window_size, cfg.sg_poly_order = int(bool(use_sg_filter)), sg_window_size, sg_poly_order
    # SG coefficients exactly as the reference forms them (fp32 pinv of the Vandermonde matrix)
    hs.coeffs = savitzky_golay_coeffs(sg_window_size, sg_poly_order)
    if sg_window_size <= _capi.MAX_SG:
        cfg.sg_coeffs_given = 1
        for i, c in enumerate(hs.coeffs.tolist()):
            cfg.sg_coeffs[i] = c
    cfg.seed, cfg.device = int(seed) & (2**64 - 1), int(device_index)
    cfg.sample_offset, cfg.total_samples = lo, int(num_samples)
    # strict=False: example/racing.py assigns the cost weights after it built the solver; forward() re-reads
    p = hs.binding.params(strict=False)
    cfg.num_model_params = len(p)
    for i, v in enumerate(p):
        cfg.model_params[i] = float(v)
    cfg.block_size = block_size
    if samples_per_thread not in (0, 1, 2):
        raise ValueError("samples_per_thread must be 0 (engine picks), 1 or 2")
    cfg.flags = {0: 0, 1: _capi.CFG_FORCE_SINGLE, 2: _capi.CFG_FORCE_PAIRED}[samples_per_thread]
    hs.params, hs.cfg = list(p), cfg
    return hs


def savitzky_golay_coeffs(window_size: int, poly_order: int) -> torch.Tensor:
    """First row of pinv(vander(-h..h)) in fp32 (mppi.py:568-596), on the host."""
    if window_size % 2 == 0 or window_size <= poly_order:
        raise ValueError("window_size must be odd and greater than poly_order.")
    half = (window_size - 1) // 2
    idx = torch.arange(-half, half + 1, dtype=torch.float32)
    return torch.linalg.pinv(torch.vander(idx, N=poly_order + 1, increasing=True))[0]


def connect_shards_inprocess(solvers) -> None:
    """Wire the fused peer exchange between shard solvers that live in ONE process (``shard=(r, world)``,
    fixed lambda / MPO). Afterwards every shard's ``forward`` is a single launch, but the shards wait for
    each other inside the kernel, so they must be launched concurrently: one CUDA stream (or GPU) each,
    see ``solve_fused_shards_inprocess``."""
    world = len(solvers)
    ptrs = (C.c_uint64 * world)()
    scratch = (C.c_uint8 * 64)()
    for r, sv in enumerate(solvers):
        with torch.cuda.device(sv._device):
            _capi.check(sv._lib.mppi_p2p_export(sv._h, scratch))
        v = C.c_uint64()
        _capi.check(sv._lib.mppi_p2p_mailbox_ptr(sv._h, C.byref(v)))
        ptrs[r] = v.value
    for r, sv in enumerate(solvers):
        _capi.check(sv._lib.mppi_p2p_connect_local(sv._h, ptrs, world, r))
        sv._fused = True


def solve_fused_shards_inprocess(solvers, state, streams, noise: Optional[torch.Tensor] = None):
    """Launch one fused sharded solve per solver, each on its own stream, then wait for all."""
    out = []
    for sv, st in zip(solvers, streams):
        with torch.cuda.stream(st):
            out.append(sv.forward(state, noise=noise))
    for st in streams:
        st.synchronize()
    for sv in solvers:
        flag = C.c_int32()
        _capi.check(sv._lib.mppi_p2p_status(sv._h, C.byref(flag)))
        if flag.value:
            raise RuntimeError("fused shard exchange timed out (a shard's solve was not running concurrently)")
    return out


def solve_shards_inprocess(solvers, state, noise: Optional[torch.Tensor] = None):
    """Drive ``world`` shard solvers (``MPPI(..., shard=(r, world))``) that live in ONE process - one
    per GPU, or several on one GPU - through a sharded solve: stage 1 on every shard, concatenate the
    partials (device-to-device copies instead of a collective), finish on every shard. Returns the
    per-shard ``(action_seq, state_seq)`` list; all entries agree."""
    world = len(solvers)
    ctx = []
    for sv in solvers:
        assert sv._world == world
        st = sv._device_state(state)
        sv._bind_maps(required=True)
        sv._refresh_params()
        ref = sv._device_refpath()
        nz = None
        sv.__dict__["_noise_full"] = None
        if noise is not None:
            full = torch.as_tensor(noise).detach().to(sv._device, torch.float32).contiguous()
            sv.__dict__["_noise_full"] = full
            nz = full[sv._shard_lo: sv._shard_lo + sv._local_samples].contiguous()
        sv._noise_keepalive = (nz, st, ref)
        s = _stream_ptr(sv._device)
        sv._shard_stage_rollout(st, ref, nz, s)
        ctx.append((st, s))
    if solvers[0]._auto_lambda in ("LBPS", "ESSPS"):
        for sv in solvers:
            torch.cuda.current_stream(sv._device).synchronize()
        for sv, (st, s) in zip(solvers, ctx):
            all_costs = torch.cat([o._wrap_costs().to(sv._device) for o in solvers]).contiguous()
            sv._shard_stage_lambda(all_costs, s)
    for sv in solvers:
        torch.cuda.current_stream(sv._device).synchronize()
    out = []
    for sv, (st, s) in zip(solvers, ctx):
        gathered = torch.cat([o._wrap_partial().to(sv._device) for o in solvers]).contiguous()
        action = torch.empty(sv._horizon, sv._dim_control, device=sv._device)
        states = torch.empty(sv._horizon + 1, sv._dim_state, device=sv._device)
        sv._shard_stage_finish(gathered, st, action, states, s)
        out.append((action, states.view(1, sv._horizon + 1, sv._dim_state)))
    return out


def top_samples_inprocess(solvers, n: int):
    """``get_top_samples(n)`` for shard solvers that live in ONE process: every shard's n best candidates are
    concatenated (device copies instead of the all-gather) and each shard merges and re-rolls the global winners.
    Returns the per-shard ``(top_traj, top_w)`` list; all entries agree."""
    cands = [sv.top_candidates(n) for sv in solvers]
    for sv in solvers:
        torch.cuda.current_stream(sv._device).synchronize()
    out = []
    for sv in solvers:
        cost = torch.cat([c.to(sv._device) for c, _ in cands])
        ids = torch.cat([i.to(sv._device) for _, i in cands])
        out.append(sv.step_epilogue(top_n=n, candidates=(cost, ids))[3])
    return out


# ---------------------------------------------------------------------- sharding helpers (host logic, CPU-testable)


def shard_bounds(num_samples: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous global sample ids [lo, hi) owned by ``rank``; sizes differ by at most 1."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(num_samples, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_shards(local: torch.Tensor, total: int, world: int, group=None) -> torch.Tensor:
    """all-gather of per-rank vectors whose lengths follow ``shard_bounds`` (differ by <= 1)."""
    import torch.distributed as dist

    sizes = [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]
    width = max(sizes)
    padded = local.new_zeros(width)
    padded[: local.numel()] = local
    buf = local.new_empty(world * width)
    if hasattr(dist, "all_gather_into_tensor") and local.is_cuda:
        dist.all_gather_into_tensor(buf, padded, group=group)
    else:
        chunks = [local.new_empty(width) for _ in range(world)]
        dist.all_gather(chunks, padded, group=group)
        buf = torch.cat(chunks)
    return torch.cat([buf[r * width: r * width + sizes[r]] for r in range(world)]).contiguous()


def gather_candidates(cost: torch.Tensor, ids: torch.Tensor, world: int, group=None):
    """All-gather of per-rank top-n candidate lists: ``cost`` [n] fp32 and ``ids`` [n] int32 (global sample ids,
    -1 = padding) travel as ONE buffer of 2n 32-bit words per rank (the ids bit-reinterpreted, never converted).
    Returns ``(costs [world*n], ids [world*n])`` in rank order."""
    import torch.distributed as dist

    n = cost.numel()
    assert ids.numel() == n and cost.dtype == torch.float32 and ids.dtype == torch.int32
    mine = torch.cat([cost.view(torch.int32), ids])  # integer words: no NaN canonicalisation anywhere on the way
    buf = torch.empty(world * 2 * n, device=cost.device, dtype=torch.int32)
    if hasattr(dist, "all_gather_into_tensor") and cost.is_cuda:
        dist.all_gather_into_tensor(buf, mine, group=group)
    else:
        chunks = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(chunks, mine, group=group)
        buf = torch.cat(chunks)
    buf = buf.view(world, 2, n)
    return buf[:, 0].reshape(-1).contiguous().view(torch.float32), buf[:, 1].reshape(-1).contiguous()


class _CudaArrayView:
    """Minimal __cuda_array_interface__ carrier so torch can alias engine-owned memory."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def wrap_device_memory(ptr: int, n: int, device: torch.device) -> torch.Tensor:
    return torch.as_tensor(_CudaArrayView(ptr, n), device=device)
