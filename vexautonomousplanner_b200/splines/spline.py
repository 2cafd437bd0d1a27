"""Mirror of src/splines/spline.py: abstract base class with heading / curvature helpers.

`get_heading` / `get_curvature` (spline.py:48-80) are served by the CUDA evaluation kernel (vap_eval, which = 3)
in the concrete subclass; the polyline `get_arc_length` of the base class (spline.py:82-105) is shadowed by
QuinticHermiteSpline.get_arc_length exactly as in the reference and is kept only for API completeness.
"""
from abc import ABC, abstractmethod
from typing import Dict, Optional, Tuple

import numpy as np


class Spline(ABC):
    """Abstract base class for spline curves"""

    def __init__(self):
        self.x_points: Optional[np.ndarray] = None
        self.y_points: Optional[np.ndarray] = None
        self._length_cache: Dict[Tuple[int, int], float] = {}

    def fit(self, x: np.ndarray, y: np.ndarray) -> bool:
        if len(x) != len(y) or len(x) < 2:
            return False
        self.x_points = np.array(x)
        self.y_points = np.array(y)
        return True

    @abstractmethod
    def get_point(self, t: float) -> np.ndarray:
        """Get point on spline at parameter t"""

    @abstractmethod
    def get_derivative(self, t: float) -> np.ndarray:
        """Get first derivative at parameter t"""

    @abstractmethod
    def get_second_derivative(self, t: float) -> np.ndarray:
        """Get second derivative at parameter t"""

    @abstractmethod
    def _heading_curvature(self, t: float):
        """(heading, curvature) at parameter t"""

    def get_heading(self, t: float) -> float:
        """Heading angle in radians, atan2(dy, dx) (spline.py:48-59)."""
        return self._heading_curvature(t)[0]

    def get_curvature(self, t: float) -> float:
        """Curvature (dx*ddy - dy*ddx) / (dx^2 + dy^2)^(3/2), 0 when |den| < 1e-10 (spline.py:61-80)."""
        return self._heading_curvature(t)[1]
