"""Mirror of the reference package src/splines (same module and class names)."""
from .quintic_hermite_spline import QuinticHermiteSpline  # noqa: F401
from .spline import Spline  # noqa: F401
from .spline_manager import PathLookupTable, QuinticHermiteSplineManager  # noqa: F401
