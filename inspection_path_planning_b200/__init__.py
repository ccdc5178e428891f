"""B200-native plan-fitness evaluator behind the API of kaiolae/Inspection-Path-Planning's SWIG module.

    from inspection_path_planning_b200 import cpp_binding
"""
from . import cpp_binding  # noqa: F401
from .cpp_binding import PlanCoverageEstimator, circulating, image, normal, nothing  # noqa: F401
