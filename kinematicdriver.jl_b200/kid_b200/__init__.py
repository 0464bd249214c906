"""kid_b200 — host-side mirror of KinematicDriver.jl's API surface for the device hot path.

Sub-modules: parameters (ClimaParams plumbing), equation_types (style tags + factories),
initial_condition, model (Case assembly), lib (ctypes binding of libkid_cuda.so),
solve (ODE.solve replacement), drivers (run_KiD_simulation / run_box_simulation / 2-D).
"""
from . import equation_types, initial_condition, lib, model, parameters  # noqa: F401
