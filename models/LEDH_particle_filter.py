"""Drop-in for the reference's ``models/LEDH_particle_filter.py`` (namespace-package overlay,
see models/particle_filter.py).  Exports the same symbols as the reference module."""

from particle_filters_b200.ledh_particle_filter import (  # noqa: F401
    Array,
    GaussianTracker,
    GFn,
    HFn,
    JacobianHFn,
    LEDHConfig,
    LEDHFlowPF,
    LogLikePdf,
    LogTransPdf,
    PFState,
    effective_sample_size,
    systematic_resample,
)
