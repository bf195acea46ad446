"""Initial position / velocity distributions (reference distributions.py:12-185).

`draw(n)` keeps the reference's contract -- an (3, n) float64 array sampled on
the host from NumPy's global RNG -- and is the injection seam for verification
(the simulator replays whatever a custom Distribution draws).  For the built-in
classes below the simulator does not call draw(): the same distributions are
sampled on the GPU by a counter-based Philox4x32-10 generator
(csrc/cmt_device.cuh: draw), indexed by the global molecule number.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from pathlib import Path

import numpy as np

__all__ = [
    "Distribution", "GaussianDistribution", "CeNTREXVelocityDistribution",
    "CeNTREXPositionDistribution", "GaussianPositionDistribution",
    "StandardVelocityDistribution", "StandardPositionDistribution",
]


class Distribution(ABC):
    @abstractmethod
    def draw(self, n: int) -> np.ndarray:
        """n samples, shape (3, n) for position/velocity distributions."""

    @abstractmethod
    def save_to_hdf(self, filepath: Path, run_name: str, group_name: str = None):
        ...


def _normal(mean, sigma, n):
    try:
        from scipy.stats import norm

        return norm.rvs(loc=mean, scale=sigma, size=n)
    except ImportError:  # pragma: no cover
        return np.random.normal(mean, sigma, size=n)


def _uniform(n):
    try:
        from scipy.stats import uniform

        return uniform.rvs(size=n)
    except ImportError:  # pragma: no cover
        return np.random.uniform(size=n)


@dataclass
class GaussianDistribution(Distribution):
    mean: float
    sigma: float

    def draw(self, n: int) -> np.ndarray:
        return _normal(self.mean, self.sigma, n)

    def save_to_hdf(self, filepath: Path, run_name: str, group_name: str = None):
        raise NotImplementedError("Saving GaussianDistribution to hdf is not implemented. Save child class instead.")


@dataclass
class CeNTREXVelocityDistribution(Distribution):
    """Independent Gaussians per axis; forward velocity 184 +- 16 m/s (distributions.py:54-76)."""

    vx: float = 0.0
    sigmax: float = 39.5
    vy: float = 0.0
    sigmay: float = 39.5
    vz: float = 184.0
    sigmaz: float = 16.0

    def draw(self, n: int) -> np.ndarray:
        return np.array((_normal(self.vx, self.sigmax, n), _normal(self.vy, self.sigmay, n),
                         _normal(self.vz, self.sigmaz, n)))

    def save_to_hdf(self, filepath: Path, run_name: str, group_name: str = "velocity_distribution"):
        from ._hdf import save_distribution

        save_distribution(self, filepath, run_name, "velocity_distribution")


@dataclass
class CeNTREXPositionDistribution(Distribution):
    """Uniform on a disc of diameter d at height z (distributions.py:101-119)."""

    d: float = 0.02
    z: float = 0.25 * 0.0254

    def draw(self, n: int) -> np.ndarray:
        theta = _uniform(n) * 2 * np.pi
        r = np.sqrt(_uniform(n)) * self.d / 2
        return np.array((r * np.cos(theta), r * np.sin(theta), np.full(n, self.z)))

    def save_to_hdf(self, filepath: Path, run_name: str, group_name: str = "position_distribution"):
        from ._hdf import save_distribution

        save_distribution(self, filepath, run_name, "position_distribution")


@dataclass
class GaussianPositionDistribution(Distribution):
    """Gaussian in x and y at height z (distributions.py:144-162)."""

    sigmax: float = 0.25 * 25.4 / 5 * 3.8e-3
    sigmay: float = 0.25 * 25.4 / 5 * 3.8e-3
    z: float = 0.25 * 0.0254

    def draw(self, n: int) -> np.ndarray:
        return np.vstack((_normal(0, self.sigmax, n), _normal(0, self.sigmay, n), np.full(n, self.z)))

    def save_to_hdf(self, filepath: Path, run_name: str, group_name: str = "position_distribution"):
        from ._hdf import save_distribution

        save_distribution(self, filepath, run_name, "position_distribution")


# names used by the reference's README (README.md:44); the code there calls them CeNTREX*
StandardVelocityDistribution = CeNTREXVelocityDistribution
StandardPositionDistribution = CeNTREXPositionDistribution
