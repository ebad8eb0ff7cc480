"""RenderAlgorithm enum + traits, kept verbatim as the dispatch key.

Values and order follow ``RenderAlgorithmEnum`` (FractalSharkLib/RenderAlgorithm.h:81-159); the traits
(MainType / LAv2 mode / PerturbExtras) follow the compile-time table at RenderAlgorithm.h:1042-1066.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass


class Numeric(enum.IntEnum):
    """``enum fs_numeric`` of include/fs_gpu.h."""
    F32 = 0
    F64 = 1
    X2_32 = 2
    HDR32 = 3
    HDR64 = 4
    HDR2X32 = 5
    X2_64 = 6
    X4_32 = 7
    X4_64 = 8


class LAv2Mode(enum.IntEnum):
    Invalid = 0
    Full = 1
    PO = 2
    LAO = 3


class PerturbExtras(enum.IntEnum):
    Disable = 0
    Bad = 1
    SimpleCompression = 2


_NAMES = """CpuHigh Cpu64 CpuHDR32 CpuHDR64 Cpu64PerturbedBLA Cpu32PerturbedBLAHDR Cpu64PerturbedBLAHDR
Cpu32PerturbedBLAV2HDR Cpu64PerturbedBLAV2HDR Cpu32PerturbedRCBLAV2HDR Cpu64PerturbedRCBLAV2HDR
Gpu1x32 Gpu2x32 Gpu4x32 Gpu1x64 Gpu2x64 Gpu4x64 GpuHDRx32
Gpu1x32PerturbedScaled Gpu2x32PerturbedScaled GpuHDRx32PerturbedScaled
Gpu1x64PerturbedBLA GpuHDRx32PerturbedBLA GpuHDRx64PerturbedBLA
Gpu1x32PerturbedLAv2 Gpu1x32PerturbedLAv2PO Gpu1x32PerturbedLAv2LAO
Gpu1x32PerturbedRCLAv2 Gpu1x32PerturbedRCLAv2PO Gpu1x32PerturbedRCLAv2LAO
Gpu2x32PerturbedLAv2 Gpu2x32PerturbedLAv2PO Gpu2x32PerturbedLAv2LAO
Gpu2x32PerturbedRCLAv2 Gpu2x32PerturbedRCLAv2PO Gpu2x32PerturbedRCLAv2LAO
Gpu1x64PerturbedLAv2 Gpu1x64PerturbedLAv2PO Gpu1x64PerturbedLAv2LAO
Gpu1x64PerturbedRCLAv2 Gpu1x64PerturbedRCLAv2PO Gpu1x64PerturbedRCLAv2LAO
GpuHDRx32PerturbedLAv2 GpuHDRx32PerturbedLAv2PO GpuHDRx32PerturbedLAv2LAO
GpuHDRx32PerturbedRCLAv2 GpuHDRx32PerturbedRCLAv2PO GpuHDRx32PerturbedRCLAv2LAO
GpuHDRx2x32PerturbedLAv2 GpuHDRx2x32PerturbedLAv2PO GpuHDRx2x32PerturbedLAv2LAO
GpuHDRx2x32PerturbedRCLAv2 GpuHDRx2x32PerturbedRCLAv2PO GpuHDRx2x32PerturbedRCLAv2LAO
GpuHDRx64PerturbedLAv2 GpuHDRx64PerturbedLAv2PO GpuHDRx64PerturbedLAv2LAO
GpuHDRx64PerturbedRCLAv2 GpuHDRx64PerturbedRCLAv2PO GpuHDRx64PerturbedRCLAv2LAO
AUTO MAX""".split()

RenderAlgorithm = enum.IntEnum("RenderAlgorithm", {n: i for i, n in enumerate(_NAMES)})


@dataclass(frozen=True)
class Traits:
    family: str          # "cpu" | "direct" | "scaled" | "bla" | "lav2" | "meta"
    numeric: Numeric | None
    mode: LAv2Mode = LAv2Mode.Invalid
    pextras: PerturbExtras = PerturbExtras.Disable


_DIRECT = {"Gpu1x32": Numeric.F32, "Gpu2x32": Numeric.X2_32, "Gpu4x32": Numeric.X4_32, "Gpu1x64": Numeric.F64,
           "Gpu2x64": Numeric.X2_64, "Gpu4x64": Numeric.X4_64, "GpuHDRx32": Numeric.HDR64}  # T of Render<IterType,T> (Fractal.cpp:1228-1253)
_PREFIX = {"Gpu1x32": Numeric.F32, "Gpu2x32": Numeric.X2_32, "Gpu1x64": Numeric.F64, "GpuHDRx32": Numeric.HDR32,
           "GpuHDRx2x32": Numeric.HDR2X32, "GpuHDRx64": Numeric.HDR64}


def traits(alg: "RenderAlgorithm") -> Traits:
    """Family / numeric type / LAv2 mode / PerturbExtras of an algorithm."""
    name = RenderAlgorithm(alg).name
    if name in ("AUTO", "MAX"):
        return Traits("meta", None)
    if name.startswith("Cpu"):
        return Traits("cpu", None)
    if name in _DIRECT:
        return Traits("direct", _DIRECT[name])
    for suffix, fam in (("PerturbedScaled", "scaled"), ("PerturbedBLA", "bla")):
        if name.endswith(suffix):
            if fam == "scaled":
                # T of RenderPerturbBLAScaled<IterType,T> (Fractal.cpp:1254-1261): double for Gpu1x32PerturbedScaled,
                # HDRFloat<float> for GpuHDRx32PerturbedScaled; Gpu2x32PerturbedScaled is not wired in the reference
                num = {"Gpu1x32": Numeric.F64, "GpuHDRx32": Numeric.HDR32}.get(name[: -len(suffix)])
                return Traits(fam, num, pextras=PerturbExtras.Bad)
            return Traits(fam, _PREFIX[name[: -len(suffix)]],
                          pextras=PerturbExtras.Bad if fam == "scaled" else PerturbExtras.Disable)
    for tail, mode in (("LAv2PO", LAv2Mode.PO), ("LAv2LAO", LAv2Mode.LAO), ("LAv2", LAv2Mode.Full)):
        if name.endswith(tail):
            head = name[: -len(tail)]
            rc = head.endswith("PerturbedRC")
            head = head[: -len("PerturbedRC" if rc else "Perturbed")]
            return Traits("lav2", _PREFIX[head], mode,
                          PerturbExtras.SimpleCompression if rc else PerturbExtras.Disable)
    raise ValueError(name)
