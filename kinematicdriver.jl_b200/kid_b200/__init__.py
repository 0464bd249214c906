"""kid_b200 — host-side mirror of KinematicDriver.jl's API surface for the device hot path (libkid_cuda.so).

Sub-modules
  parameters         ClimaParams-named defaults, override_toml_dict, the flat parameter table of include/kid_params.h
  equation_types     moisture / precipitation style tags and their factories (get_moisture_type, get_precipitation_type)
  initial_condition  ShipwayHill2012 sounding, hydrostatic profile (ρ_ivp), initial_condition_1d / _0d (host side)
  model              Case assembly: k1d_case, col_sed_case, box_case, ensembles (host- or device-built initial condition)
  k2d                K2DModel cases, x-decomposition, edge exchange rings (local / torch.distributed)
  lib                ctypes binding of the C ABI (include/kid_cuda.h): Handle
  solve              ODE.solve replacement, the four drivers, OutputPipeline, PipelinedEnsemble
  calibration        make_filter_props, device-side G vectors, run_dyn_model / run_ensembles (KiD, KiD_col_sed, Box)
"""
from . import equation_types, initial_condition, lib, model, parameters  # noqa: F401
