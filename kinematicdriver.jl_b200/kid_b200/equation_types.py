"""Equation-style tags and the option -> type factories.

Mirrors src/Common/equation_types.jl:23-61 and src/Common/helper_functions.jl:4-93.
The styles select the kernel variant of libkid_cuda.so and carry the toml dict the
parameter table is flattened from.  Styles the device path does not implement
(CloudyMoisture/CloudyPrecip, MoistureP3/PrecipitationP3) raise: there is no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass

# enum values of include/kid_cuda.h
MOIST_EQ, MOIST_NONEQ = 0, 1
PRECIP_NONE, PRECIP_0M, PRECIP_1M, PRECIP_2M = 0, 1, 2, 3
RF = {"CliMA_1M": 0, "KK2000": 1, "B1994": 2, "TC1980": 3, "LD2004": 4, "VarTimeScaleAcnv": 5,
      "SB2006": 6, "SB2006NL": 7}
SED = {"CliMA_1M": 0, "SB2006": 1, "Chen2022": 2}


class AbstractStyle:
    pass


class AbstractMoistureStyle(AbstractStyle):
    pass


class AbstractPrecipitationStyle(AbstractStyle):
    pass


@dataclass(frozen=True)
class EquilibriumMoisture(AbstractMoistureStyle):
    kid_enum: int = MOIST_EQ


@dataclass(frozen=True)
class NonEquilibriumMoisture(AbstractMoistureStyle):
    liquid: object = None  # CMP.CloudLiquid(toml_dict): the toml dict stands in for it
    ice: object = None
    kid_enum: int = MOIST_NONEQ


@dataclass(frozen=True)
class NoPrecipitation(AbstractPrecipitationStyle):
    kid_enum: int = PRECIP_NONE
    rain_formation: int = 0
    sedimentation: int = 0


@dataclass(frozen=True)
class Precipitation0M(AbstractPrecipitationStyle):
    params: object = None
    kid_enum: int = PRECIP_0M
    rain_formation: int = 0
    sedimentation: int = 0


@dataclass(frozen=True)
class Precipitation1M(AbstractPrecipitationStyle):
    rain_formation_name: str = "CliMA_1M"
    sedimentation_name: str = "CliMA_1M"
    kid_enum: int = PRECIP_1M

    @property
    def rain_formation(self):
        return RF[self.rain_formation_name]

    @property
    def sedimentation(self):
        return SED[self.sedimentation_name]


@dataclass(frozen=True)
class Precipitation2M(AbstractPrecipitationStyle):
    rain_formation_name: str = "SB2006"
    sedimentation_name: str = "SB2006"
    kid_enum: int = PRECIP_2M

    @property
    def rain_formation(self):
        return RF[self.rain_formation_name]

    @property
    def sedimentation(self):
        return SED[self.sedimentation_name]


def get_moisture_type(moisture_choice: str, toml_dict=None):
    """src/Common/helper_functions.jl:4-17"""
    if moisture_choice == "EquilibriumMoisture":
        return EquilibriumMoisture()
    if moisture_choice == "NonEquilibriumMoisture":
        return NonEquilibriumMoisture(toml_dict, toml_dict)
    if moisture_choice in ("CloudyMoisture", "MoistureP3"):
        raise NotImplementedError(
            f"{moisture_choice} is outside the device hot path of libkid_cuda.so and there is no CPU fallback")
    raise ValueError(f"Invalid moisture choice: {moisture_choice}")


def get_precipitation_type(precipitation_choice: str, toml_dict=None, *, rain_formation_choice=None,
                           sedimentation_choice=None, boundary=None):
    """src/Common/helper_functions.jl:22-93"""
    if precipitation_choice == "NoPrecipitation":
        return NoPrecipitation()
    if precipitation_choice == "Precipitation0M":
        return Precipitation0M(toml_dict)
    if precipitation_choice == "Precipitation1M":
        if sedimentation_choice in ("CliMA_1M", None):
            st = "CliMA_1M"
        elif sedimentation_choice == "Chen2022":
            # CMP.Chen2022VelType (helper_functions.jl:36-37): rain by Chen et al. (2022) Table B1; the snow / ice tables
            # (B2-B4) are not part of the parameter table — snow keeps the Blk1M law (warm columns carry no snow)
            st = "Chen2022"
        else:
            raise ValueError(f"Invalid sedimentation choice: {sedimentation_choice}")
        rf = "CliMA_1M" if rain_formation_choice is None else rain_formation_choice
        if rf not in ("CliMA_1M", "KK2000", "B1994", "TC1980", "LD2004", "VarTimeScaleAcnv"):
            raise ValueError(f"Invalid rain formation choice: {rain_formation_choice}")
        return Precipitation1M(rf, st)
    if precipitation_choice == "Precipitation2M":
        if sedimentation_choice in ("SB2006", None):
            st = "SB2006"
        elif sedimentation_choice == "Chen2022":
            st = "Chen2022"  # CMP.Chen2022VelTypeRain (helper_functions.jl:69-70)
        else:
            raise ValueError(f"Invalid sedimentation choice: {sedimentation_choice}")
        rf = "SB2006" if rain_formation_choice is None else rain_formation_choice
        if rf not in ("SB2006", "SB2006NL"):
            raise ValueError(f"Invalid rain formation choice: {rain_formation_choice}")
        return Precipitation2M(rf, st)
    if precipitation_choice in ("CloudyPrecip", "PrecipitationP3"):
        raise NotImplementedError(
            f"{precipitation_choice} is outside the device hot path of libkid_cuda.so and there is no CPU fallback")
    raise ValueError(f"Invalid precipitation choice: {precipitation_choice}")


def state_variable_names(moisture, precip) -> tuple[str, ...]:
    """initialise_state field sets, src/Common/ode_utils.jl:23-54"""
    names = ["ρq_tot"]
    if moisture.kid_enum == MOIST_NONEQ:
        names += ["ρq_liq", "ρq_ice"]
    if precip.kid_enum in (PRECIP_1M, PRECIP_2M):
        names += ["ρq_rai", "ρq_sno"]
    if precip.kid_enum == PRECIP_2M:
        names += ["N_liq", "N_rai", "N_aer"]
    return tuple(names)


def state_variable_names_by_enum(moisture_enum: int, precip_enum: int) -> tuple[str, ...]:
    """Same, from the C-ABI enums of a kid_config_t."""
    from types import SimpleNamespace
    return state_variable_names(SimpleNamespace(kid_enum=moisture_enum), SimpleNamespace(kid_enum=precip_enum))
