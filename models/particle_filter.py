"""Drop-in for the reference's ``models/particle_filter.py`` (namespace-package overlay).

``models/`` has no ``__init__.py`` here or in the reference, so with this repo
ahead of the reference on ``sys.path`` this module shadows the reference's
``models.particle_filter`` while every other ``models.*`` / ``simulator.*``
import still resolves to the reference.  Exports the same
symbols: ``Array, GFn, HFn, PFState, ParticleFilter``.
"""

from particle_filters_b200.particle_filter import (  # noqa: F401
    Array,
    GFn,
    HFn,
    PFState,
    ParticleFilter,
)
