"""Registered device functors: the equations, detectors and callbacks of the
reference's shipped examples and tests (``include/diffurch_b200.h: dfb_system_id``).

A Rust host closure cannot run on the GPU, so ``Solver.equation`` takes one of
these descriptors plus the per-trajectory values its closure captures.
"""
from dataclasses import dataclass, field
from typing import Dict, Tuple

from . import _abi as abi


@dataclass(frozen=True)
class System:
    name: str
    id: int
    n_state: int
    params: Tuple[str, ...] = ()
    n_aux: int = 0
    detectors: Dict[str, int] = field(default_factory=dict)
    callbacks: Dict[str, int] = field(default_factory=dict)
    uses_history: bool = False
    reference: str = ""

    @property
    def n_params(self):
        return len(self.params)


Lorenz = System("lorenz", abi.SYS_LORENZ, 3, ("sigma", "rho", "beta"),
                reference="examples/ode-chaos-lorenz.rs:19-22")
LorenzVariational = System("lorenz_variational", abi.SYS_LORENZ_VARIATIONAL, 12,
                           ("sigma", "rho", "beta"), n_aux=3, callbacks={"qr": 1},
                           reference="tests/lyapunov_exponents.rs:264-296")
Linear1 = System("linear1", abi.SYS_LINEAR1, 1, ("lambda",), n_aux=1, callbacks={"renormalize": 1},
                 reference="tests/lyapunov_exponents.rs:17-31")
Linear1Sin = System("linear1_sin", abi.SYS_LINEAR1_SIN, 1, ("lambda",), n_aux=1,
                    callbacks={"renormalize": 1}, reference="tests/lyapunov_exponents.rs:55")
Linear3 = System("linear3", abi.SYS_LINEAR3, 3, tuple(f"a{i}{j}" for i in range(3) for j in range(3)),
                 reference="examples/ode-linear-3-fibonacci.rs:11-20")
Linear3Matrix = System("linear3_matrix", abi.SYS_LINEAR3_MATRIX, 9,
                       tuple(f"a{i}{j}" for i in range(3) for j in range(3)), n_aux=3,
                       callbacks={"qr_ln": 1, "qr_ln_abs": 2},
                       reference="tests/lyapunov_exponents.rs:107-131")
Harmonic = System("harmonic", abi.SYS_HARMONIC, 2, reference="tests/equation_types.rs:89-92")
PowerDecay = System("power_decay", abi.SYS_POWER_DECAY, 1, reference="tests/equation_types.rs:107")
Constant3 = System("constant3", abi.SYS_CONSTANT3, 3, reference="tests/equation_types.rs:14")
Time3 = System("time3", abi.SYS_TIME3, 3, reference="tests/equation_types.rs:31")
Zero1 = System("zero1", abi.SYS_ZERO1, 1, detectors={"half_time": 1},
               reference="tests/delay_propagation.rs:13,86")
DdeLinear = System("dde_linear", abi.SYS_DDE_LINEAR, 1, ("a", "b", "tau", "k"), uses_history=True,
                   reference="tests/equation_types.rs:116-136")
NddeLinear = System("ndde_linear", abi.SYS_NDDE_LINEAR, 1, ("a", "b", "tau", "k"), uses_history=True,
                    reference="tests/equation_types.rs:139-158")
BouncingBall = System("bouncing_ball", abi.SYS_BOUNCING_BALL, 2, ("g", "k"),
                      detectors={"dx": 1, "x": 2}, callbacks={"bounce": 1},
                      reference="examples/bouncing-ball.rs:21-33")
EggBilliard = System("egg_billiard", abi.SYS_EGG_BILLIARD, 4, ("max_reflections",), n_aux=1,
                     detectors={"boundary": 1}, callbacks={"reflect": 1},
                     reference="examples/egg-billiard.rs:12-40")
RelayMsign = System("relay_msign", abi.SYS_RELAY_MSIGN, 2, detectors={"x": 1},
                    reference="examples/ode-2-relay-msign.rs:17-20")
DdeRelayClamp = System("dde_relay_clamp", abi.SYS_DDE_RELAY_CLAMP, 2, ("alpha", "sigma", "T"),
                       detectors={"abs_delayed_x_minus_1": 1}, uses_history=True,
                       reference="examples/dde-relay-clamp.rs:24-35")

ALL = [Lorenz, LorenzVariational, Linear1, Linear1Sin, Linear3, Linear3Matrix, Harmonic,
       PowerDecay, Constant3, Time3, Zero1, DdeLinear, NddeLinear, BouncingBall, EggBilliard,
       RelayMsign, DdeRelayClamp]
BY_ID = {s.id: s for s in ALL}


def register(system):
    """Add a user system (diffurch_b200.user.load_user_system) to the id registry."""
    BY_ID[system.id] = system
    return system
