"""interpax_b200 -- the query-evaluation path of interpax (interp1d/2d/3d, Interpolator1D/2D/3D,
approx_df, and the PPoly / CubicSpline family) as hand-written CUDA for B200 (sm_100a) behind a C ABI (include/ipx.h).

Importing the package does not need a GPU; calling any function does, and raises if the CUDA
library or a CUDA device is missing (there is no CPU fallback).
"""

from .api import (
    CUBIC_METHODS,
    METHODS_1D,
    METHODS_2D,
    METHODS_3D,
    OTHER_METHODS,
    Interpolator1D,
    Interpolator2D,
    Interpolator3D,
    approx_df,
    approx_df_vjp,
    bin_index,
    interp1d,
    interp2d,
    interp3d,
    jvp_knots,
    vjp_f,
    vjp_knots,
    vjp_tables,
)
from .ppoly import Akima1DInterpolator, CubicHermiteSpline, CubicSpline, PchipInterpolator, PPoly

__all__ = [
    "Akima1DInterpolator",
    "CubicHermiteSpline",
    "CubicSpline",
    "PchipInterpolator",
    "PPoly",
    "approx_df",
    "approx_df_vjp",
    "bin_index",
    "Interpolator1D",
    "Interpolator2D",
    "Interpolator3D",
    "interp1d",
    "interp2d",
    "interp3d",
    "jvp_knots",
    "vjp_f",
    "vjp_knots",
    "vjp_tables",
]

__version__ = "0.1.0"
