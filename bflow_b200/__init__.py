"""bflow_b200 — B200-native (sm_100a) RAFT-spline inference hot path of uzh-rpg/bflow.

Public surface = the reference's own: ``RAFTSpline`` and ``BezierCurves`` (models/raft_spline/raft.py).  The operator-level
mirrors of models/raft_utils/{corr,utils}.py live in :mod:`bflow_b200.ops`, the event representation of
data/utils/representations.py in :mod:`bflow_b200.events`, the flow metrics of utils/metrics.py in :mod:`bflow_b200.metrics`.
"""
from .bezier import BezierCurves
from .raft import RAFTSpline
from . import config

__all__ = ['RAFTSpline', 'BezierCurves', 'config']
__version__ = '0.1.0'
