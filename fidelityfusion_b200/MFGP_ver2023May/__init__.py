"""Drop-in replacements for the gen-2023 operator modules of the reference (MFGP_ver2023May/)."""
from .base_gp.cigp import CIGP
from .base_gp.hogp import HOGP
from .base_gp.fides import FIDES
