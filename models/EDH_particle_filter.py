"""Drop-in for the reference's ``models/EDH_particle_filter.py`` (namespace-package overlay,
see models/particle_filter.py).  Exports the same symbols as the reference module."""

from particle_filters_b200.edh_particle_filter import (  # noqa: F401
    Array,
    EDHConfig,
    EDHFlowPF,
    EKFTracker,
    GaussianTracker,
    GFn,
    HFn,
    JacobianHFn,
    LogLikePdf,
    LogTransPdf,
    PFState,
    UKFTracker,
    effective_sample_size,
    rk4_step,
    systematic_resample,
)
