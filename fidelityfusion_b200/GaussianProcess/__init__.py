"""Drop-in replacements for the reference's GaussianProcess/ operator modules (same names, signatures,
state_dict keys); the math runs in libffgp's sm_100a kernels."""
