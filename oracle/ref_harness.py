"""TEST INFRASTRUCTURE ONLY - loader for the *live* reference (kohonda/mppi_playground).

This module imports the unmodified reference from ``/root/reference`` so that
``oracle/gen_golden.py`` can record golden vectors from it and so that the CPU
tests in this container can cross-check ``oracle/mppi_oracle.py`` against it.
It never copies reference sources; it only imports them where they lie.

``/root/reference`` does not exist on the GPU box: nothing under ``-m gpu``,
``__graft_entry__.smoke()`` or ``bench.py`` may import this file.

What has to be stubbed (render / CLI only imports of the reference):
  matplotlib(.pyplot)            src/envs/racing_env.py:12, obstacle_map_2d.py:13,
                                 circuit_generator/path_generate.py:7
  moviepy.video.io.ImageSequenceClip   src/envs/racing_env.py:13
  fire, gymnasium, tqdm          example/*.py (CLI + simulator)
The reference reads ``src/envs/circuit_generator/circuit.csv`` relative to the
cwd (src/envs/racing_env.py:47-49) so env construction runs with
cwd=/root/reference.
"""
from __future__ import annotations

import ast
import contextlib
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("MPPI_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "pi_mpc", "mppi.py"))


_STUBS = [
    "matplotlib",
    "matplotlib.pyplot",
    "matplotlib.patches",
    "moviepy",
    "moviepy.video",
    "moviepy.video.io",
    "moviepy.video.io.ImageSequenceClip",
    "fire",
    "gymnasium",
]


def _install_stubs() -> None:
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = mock.MagicMock(name=name)
    try:
        import tqdm  # noqa: F401
    except Exception:
        sys.modules["tqdm"] = mock.MagicMock(name="tqdm")


def _install_paths() -> None:
    sys.dont_write_bytecode = True  # the reference tree is read-only
    for sub in ("src", "example"):
        p = os.path.join(REFERENCE_ROOT, sub)
        if p not in sys.path:
            sys.path.insert(0, p)


@contextlib.contextmanager
def _cwd(path: str):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def load_reference():
    """Return a namespace with the reference's MPPI class and env classes."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    _install_paths()
    # our drop-in also calls itself pi_mpc; make sure the reference's wins here
    for k in [k for k in sys.modules if k == "pi_mpc" or k.startswith("pi_mpc.")]:
        mod = sys.modules[k]
        if not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
            del sys.modules[k]
    ns = types.SimpleNamespace()
    from pi_mpc.mppi import MPPI  # type: ignore

    ns.MPPI = MPPI
    from envs.navigation_2d import Navigation2DEnv  # type: ignore
    from envs.racing_env import RacingEnv  # type: ignore

    ns.Navigation2DEnv = Navigation2DEnv
    ns.RacingEnv = RacingEnv
    return ns


def make_racing():
    """(env, controller) exactly as example/racing.py:221-227 builds them (CPU)."""
    ns = load_reference()
    import racing as racing_example  # type: ignore  (example/racing.py)

    with _cwd(REFERENCE_ROOT):
        env = ns.RacingEnv()
        controller = racing_example.racing_controller(env, debug=False)
    controller.set_cost_map(env._obstacle_map, env._lane_map)
    return env, controller, ns


def make_navigation2d():
    ns = load_reference()
    with _cwd(REFERENCE_ROOT):
        env = ns.Navigation2DEnv()
    return env, ns


def extract_closures(example: str, names):
    """Pull nested function defs (dynamics / cost closures) out of example/<example>.py.

    The pendulum / cartpole / mountaincar models are closures nested inside
    ``main()`` (example/pendulum.py:17-47, cartpole.py:17-81,
    mountaincar.py:17-55). They are compiled *from the reference file where it
    lies* with the ``@torch.jit.script`` decorator stripped (eager and scripted
    execution dispatch the same ATen ops).
    """
    import torch

    path = os.path.join(REFERENCE_ROOT, "example", f"{example}.py")
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    env = {"torch": torch}
    found = {}
    # module-level helpers first (angle_normalize)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name != "main":
            node.decorator_list = []
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, env)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            node.decorator_list = []
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, env)
            found[node.name] = env[node.name]
    missing = [n for n in names if n not in found]
    if missing:
        raise KeyError(f"{missing} not found in {path}")
    return [found[n] for n in names]


def load_example_with_solver(example: str, solver_cls):
    """Import example/<example>.py *where it lies* with ``pi_mpc.mppi.MPPI`` replaced by ``solver_cls``
    (how a user swaps the engine in: another ``pi_mpc`` first on PYTHONPATH). Returns the module object;
    the reference's own ``pi_mpc`` is restored in ``sys.modules`` afterwards."""
    import importlib.util

    load_reference()
    shim = types.ModuleType("pi_mpc")
    shim_mppi = types.ModuleType("pi_mpc.mppi")
    shim.MPPI = shim_mppi.MPPI = solver_cls
    shim.mppi = shim_mppi
    saved = {k: sys.modules[k] for k in list(sys.modules) if k == "pi_mpc" or k.startswith("pi_mpc.")}
    for k in saved:
        del sys.modules[k]
    sys.modules["pi_mpc"], sys.modules["pi_mpc.mppi"] = shim, shim_mppi
    try:
        path = os.path.join(REFERENCE_ROOT, "example", f"{example}.py")
        spec = importlib.util.spec_from_file_location(f"_dropin_{example}", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        del sys.modules["pi_mpc"], sys.modules["pi_mpc.mppi"]
        sys.modules.update(saved)
    return mod
