"""K2DModel host side: case assembly (run_kinematic2d_simulations.jl:15-171) and the x-decomposed stepping
with a neighbour exchange of the shared GLL edge columns (the cross-rank part of CC.Spaces.weighted_dss!,
src/K2DModel/tendency.jl:70,106-108,156-157,206-207).

Three ways to run:
  * one handle owning the whole periodic domain: `handle.step(t0, n)` (the wrap is inside the kernel);
  * one PROCESS driving several sub-domain handles (`LocalRing`): kid_k2d_exchange_local peer copies;
  * one process per GPU, exchange inside the library (`LibraryRing`, kid_k2d_comm_init): kid_step issues one grouped
    ncclSend/ncclRecv of the packed edge columns per stage on its own stream — what a Julia host would use;
  * one process per GPU (`DistributedRing`): the edge buffers are exchanged with torch.distributed
    send/recv (NCCL over NVLink; gloo in the CPU tests of the plumbing).
"""
from __future__ import annotations

import numpy as np

from . import equation_types as ET
from . import initial_condition as IC
from . import parameters as PR
from .lib import COL_Q_SURF, MODEL_K2D, Handle, make_config
from .model import Case, TimeStepping, float_enum, initialise_state, make_function_space

K2D_DEFAULTS = dict(
    moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation1M", rain_formation_choice=None,
    sedimentation_choice=None, w1=3.0, t1=1800.0, p0=100000.0, precip_sources=True, precip_sinks=True,
    prescribed_Nd=1e8, r_dry=0.04e-6, std_dry=1.4, kappa=0.9, domain_width=3000.0, domain_height=3000.0,
    nx=32, nz=32, npoly=4, dt=1.0, dt_output=120.0, t_ini=0.0, t_end=3600.0,
    open_system_activation=False, local_activation=False,
)


def partition_elements(helem_global: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block of whole spectral elements owned by `rank`: (offset, count)."""
    base, rem = divmod(helem_global, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def k2d_case(FT=np.float64, rank: int = 0, world: int = 1, device: int = 0, **options) -> Case:
    unknown = set(options) - set(K2D_DEFAULTS)
    if unknown:
        raise KeyError(f"unknown K2D option(s): {sorted(unknown)}")
    o = dict(K2D_DEFAULTS)
    o.update(options)
    FT = np.dtype(FT).type
    td = PR.override_toml_dict(
        FT=FT, w1=o["w1"], t1=o["t1"], p0=o["p0"], precip_sources=int(o["precip_sources"]),
        precip_sinks=int(o["precip_sinks"]), qtot_flux_correction=1, prescribed_Nd=o["prescribed_Nd"],
        r_dry=o["r_dry"], std_dry=o["std_dry"], kappa=o["kappa"],
        open_system_activation=int(o["open_system_activation"]), local_activation=int(o["local_activation"]))
    common_params = PR.create_common_parameters(td)
    kid_params = PR.create_kid_parameters(td)
    thermo_params = PR.create_thermodynamics_parameters(td)
    moisture = ET.get_moisture_type(o["moisture_choice"], td)
    precip = ET.get_precipitation_type(o["precipitation_choice"], td, rain_formation_choice=o["rain_formation_choice"],
                                       sedimentation_choice=o["sedimentation_choice"])
    nz, npoly, helem_global = int(o["nz"]), int(o["npoly"]), int(o["nx"])
    if o["domain_height"] > kid_params.z_2:
        raise ValueError("domain_height cannot be higher than z_2")
    offset, helem = partition_elements(helem_global, world, rank)
    if helem < 1:
        raise ValueError("more ranks than horizontal elements")
    nc = helem * (npoly + 1)
    zc, _ = make_function_space(FT, 0.0, o["domain_height"], nz)
    ρ_profile = IC.ρ_ivp(FT, kid_params, thermo_params, "ShipwayHill2012")
    ip = IC.initial_condition_1d(FT, common_params, kid_params, thermo_params, ρ_profile, zc, "ShipwayHill2012")
    q_surf = IC.init_profile(FT, kid_params, thermo_params, FT(0), "ShipwayHill2012")["qv"]
    Y_col = initialise_state(moisture, precip, ip).astype(FT)
    cfg = make_config(
        model=MODEL_K2D, moisture=moisture.kid_enum, precip=precip.kid_enum, rain_formation=precip.rain_formation,
        sedimentation=precip.sedimentation, float_type=float_enum(FT), nz=nz, nc=nc, helem=helem, npoly=npoly,
        helem_global=helem_global, helem_offset=offset,
        precip_sources=int(common_params.precip_sources), precip_sinks=int(common_params.precip_sinks),
        open_system_activation=int(common_params.open_system_activation),
        local_activation=int(common_params.local_activation), qtot_flux_correction=1, device=device,
        z_min=0.0, z_max=float(o["domain_height"]), dt=float(o["dt"]),
        domain_width=float(o["domain_width"]), domain_height=float(o["domain_height"]))
    return Case(cfg=cfg, params=PR.flatten_params(td, precip.rain_formation), ρ_dry=ip["ρ_dry"].astype(FT),
                T=ip["T"].astype(FT), Y0=np.broadcast_to(Y_col, (nc,) + Y_col.shape).copy(),
                var_names=ET.state_variable_names(moisture, precip),
                col_scalars={COL_Q_SURF: np.full(nc, q_surf, dtype=FT)}, z_centers=zc,
                meta=dict(opts=o, toml=td, moisture=moisture, precip=precip, rank=rank, world=world,
                          TS=TimeStepping(o["dt"], o["dt_output"], o["t_end"])))


class LocalRing:
    """Several sub-domain handles (a periodic ring in x) driven by one process."""

    def __init__(self, handles: list[Handle]):
        self.handles = handles

    def step(self, t0: float, nsteps: int):
        hs = self.handles
        dt = hs[0].cfg.dt
        for s in range(nsteps):
            t = t0 + s * dt
            for stage in range(3):
                for h in hs:
                    h.k2d_stage_begin(t, stage)
                for i, h in enumerate(hs):
                    h.k2d_exchange_local(hs[(i + 1) % len(hs)])
                for h in hs:
                    h.k2d_stage_end(stage)

    def get_state(self) -> np.ndarray:
        return np.concatenate([h.get_state() for h in self.handles], axis=0)


class _DevArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, n: int, dtype):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": np.dtype(dtype).str, "data": (int(ptr), False),
                                         "version": 2}


def ring_neighbours(rank: int, world: int) -> tuple[int, int]:
    return (rank - 1) % world, (rank + 1) % world


def exchange_edges(dist, send_left, send_right, recv_left, recv_right, rank: int, world: int):
    """Periodic ring exchange of the packed edge columns with torch.distributed point-to-point ops.
    recv_left <- left neighbour's send_right, recv_right <- right neighbour's send_left.
    The receive order (right first) makes the pairwise in-order matching correct for world == 2, where both
    neighbours are the same peer."""
    if world == 1:
        recv_left.copy_(send_right)
        recv_right.copy_(send_left)
        return
    left, right = ring_neighbours(rank, world)
    ops = [dist.P2POp(dist.isend, send_left, left), dist.P2POp(dist.isend, send_right, right),
           dist.P2POp(dist.irecv, recv_right, right), dist.P2POp(dist.irecv, recv_left, left)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()


class LibraryRing:
    """One process per GPU, exchange INSIDE the library (kid_k2d_comm_init): `dist` — any initialised torch.distributed
    backend — only carries the 128-byte NCCL id from rank 0 to the others; afterwards `step` is one kid_step call."""

    def __init__(self, handle: Handle, dist, rank: int, world: int):
        from .lib import nccl_unique_id
        box = [nccl_unique_id(handle.lib) if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        self.h = handle
        handle.k2d_comm_init(rank, world, box[0])

    def step(self, t0: float, nsteps: int):
        self.h.step(t0, nsteps)


class DistributedRing:
    """One process per GPU; `dist` is an initialised torch.distributed module (backend nccl)."""

    def __init__(self, handle: Handle, dist, rank: int, world: int):
        import torch
        self.h, self.dist, self.rank, self.world, self.torch = handle, dist, rank, world, torch
        ptrs, n = handle.k2d_edge_buffers()
        dev = torch.device("cuda", handle.cfg.device)
        self.bufs = [torch.as_tensor(_DevArray(p, n, handle.dtype), device=dev) for p in ptrs]
        self.stream = torch.cuda.ExternalStream(handle.stream, device=dev)

    def step(self, t0: float, nsteps: int):
        h, torch = self.h, self.torch
        dt = h.cfg.dt
        sl, sr, rl, rr = self.bufs
        with torch.cuda.stream(self.stream):  # NCCL orders itself after the pack kernel on the library's stream
            for s in range(nsteps):
                t = t0 + s * dt
                for stage in range(3):
                    h.k2d_stage_begin(t, stage)
                    exchange_edges(self.dist, sl, sr, rl, rr, self.rank, self.world)
                    h.k2d_stage_end(stage)
