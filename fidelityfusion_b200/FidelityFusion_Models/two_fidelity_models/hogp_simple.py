"""Stands in for FidelityFusion_Models/two_fidelity_models/hogp_simple.py (reference :21-126), the copy of HOGP_simple
that FidelityFusion_Models/GAR.py:6 imports: predictive-variance factor (K* K_0^-1 U_0)^2 (:68), y_var ignored (:91-94)."""
from ...GaussianProcess.hogp_simple import HOGP_simple_ffm as HOGP_simple, eigen_pairs  # noqa: F401
