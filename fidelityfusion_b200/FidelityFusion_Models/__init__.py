"""Drop-in modules for the operator-level pieces that live under the reference's FidelityFusion_Models/ package:
`two_fidelity_models/hogp_simple.py` (the HOGP_simple copy that GAR.py:6 imports) and `MF_data.py` (data manager:
subset / overlap matching, normalisation, non-subset fill).  The L4 model classes (AR, CIGAR, GAR, ...) are NOT
mirrored: the reference's own files run unmodified on these modules (fidelityfusion_b200.binding.install())."""
