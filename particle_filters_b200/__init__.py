"""particle_filters_b200 -- B200-native Sequential-Monte-Carlo engine.

Hot path of liyiyang-amber/Particle_filters (SIR particle-filter step, LEDH
particle-flow update) behind the reference's own class API; see DESIGN.md.
Importing the package does not need a GPU; running a filter does.
"""

from .particle_filter import DevicePFState, ParticleFilter, PFState  # noqa: F401
from .ssm import (AcousticMeasurement, ExpMeasurement, LinearGaussianSSM, LinearMeasurement,  # noqa: F401
                  Lorenz96SSM, StochVolSSM)
from .ledh_particle_filter import LEDHConfig, LEDHFlowPF  # noqa: F401

__all__ = ["ParticleFilter", "PFState", "DevicePFState", "StochVolSSM", "LinearGaussianSSM", "Lorenz96SSM", "LEDHFlowPF",
           "LEDHConfig", "LinearMeasurement", "AcousticMeasurement", "ExpMeasurement"]
