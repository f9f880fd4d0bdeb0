"""Drop-in for the reference's vendored ``torchdiffeq`` package (ode_net/code/torchdiffeq/__init__.py:1-3), restricted
to the PHOENIX hot path: ``odeint`` / ``odeint_adjoint`` on an ``odenet.ODENet`` right-hand side, solved on a B200 by
libphoenix_b200.so.  Same signatures, defaults and error behaviour as torchdiffeq 0.1.1 as vendored by the reference."""
from ._api import odeint, odeint_adjoint, odeint_adjoint_many, set_deferred_adjoint

__version__ = "0.1.1"
__all__ = ["odeint", "odeint_adjoint", "odeint_adjoint_many", "set_deferred_adjoint"]
