from MFGP_ver2023May.AR_AutoRegression import AR
from MFGP_ver2023May.CIGAR_ConditionalIndependentGAR import CIGAR
from MFGP_ver2023May.GAR_GeneralizedAutoAR import GAR
from MFGP_ver2023May.CAR_ContinuAR import CAR
from MFGP_ver2023May.NAR_NonlinearAR import NAR
from MFGP_ver2023May.ResGP import ResGP

from MFGP_ver2023May.base_gp.cigp import CIGP
from MFGP_ver2023May.base_gp.hogp import HOGP