"""Structural observables of stored trajectories, evaluated on the device (``mythos/observables``: propeller twist,
rise, pitch, diameter -- the four the DiffTRe examples optimise against)."""

from mythos_b200.observables.base import BaseObservable, ObservableSet, get_duplex_quartets
from mythos_b200.observables.diameter import Diameter
from mythos_b200.observables.pitch import PitchAngle, compute_pitch
from mythos_b200.observables.propeller import PropellerTwist
from mythos_b200.observables.rise import Rise

__all__ = ["BaseObservable", "Diameter", "ObservableSet", "PitchAngle", "PropellerTwist", "Rise", "compute_pitch", "get_duplex_quartets"]
