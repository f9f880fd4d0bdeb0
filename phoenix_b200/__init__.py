"""phoenix_b200 — B200-native (sm_100a) implementation of the PHOENIX NeuralODE hot path.

``phoenix_b200.odenet.ODENet`` and ``phoenix_b200.torchdiffeq.odeint / odeint_adjoint`` mirror the reference's
``odenet`` and vendored ``torchdiffeq`` modules (QuackenbushLab/phoenix, ode_net/code); all arithmetic runs in
hand-written CUDA kernels behind the C ABI of ``include/phoenix_b200.h`` (libphoenix_b200.so).  No CPU fallback.
"""
from . import engine
from .engine import check_errors, last_status, last_step_log, set_precision, set_step_logging, set_sync_errors
from .odenet import ODENet, LogShiftedSoftSignMod, SoftsignMod
from .prior import prior_grad_from_matrix, prior_loss
from .torchdiffeq import odeint, odeint_adjoint, odeint_adjoint_many, set_deferred_adjoint

__all__ = ["ODENet", "SoftsignMod", "LogShiftedSoftSignMod", "odeint", "odeint_adjoint", "odeint_adjoint_many", "prior_loss",
           "prior_grad_from_matrix", "engine", "check_errors",
           "last_status", "last_step_log", "set_precision", "set_step_logging", "set_sync_errors", "set_deferred_adjoint"]
